from .painn import PaiNN, PaiNNInteraction, PaiNNMixing
from .schnet import SchNet, SchNetInteraction

__all__ = ["PaiNN", "PaiNNInteraction", "PaiNNMixing", "SchNet", "SchNetInteraction"]
