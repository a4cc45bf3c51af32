"""ctypes binding of the C-ABI library ``csrc/libspk_b200.so`` (declared in ``include/spk_b200.h``).

There is NO fallback: if the library is missing (not built) importing any op raises, and every op raises when it is
handed a non-CUDA tensor.  torch is used only as the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPK_B200_LIB") or os.path.join(HERE, "csrc", "libspk_b200.so")   # override: A/B builds

SPK_OK = 0
ACT_NONE, ACT_SILU, ACT_SSP, ACT_GIVEN = 0, 1, 2, 3
SAVE_DERIV = 0x10
RBF_GAUSSIAN, RBF_BESSEL = 0, 1
GEO_STRIDE = 8

P = c_void_p  # device pointers and the stream travel as void*

CHAIN_MAX_STEPS = 8
CHAIN_GEMM, CHAIN_MIX_CTX, CHAIN_MIX_UPDATE, CHAIN_MIX_UPDATE_BWD, CHAIN_MIX_CTX_BWD = 0, 1, 2, 3, 4


class ChainStep(Structure):
    """spk_chain_step_t of include/spk_b200.h."""
    _fields_ = [("kind", c_int32), ("rows_per_atom", c_int32), ("K", c_int32), ("N", c_int32), ("act", c_int32),
                ("F", c_int32), ("eps", c_float), ("reserved", c_int32), ("lda", c_int64), ("ldy", c_int64),
                ("ld_add", c_int64), ("A", c_void_p), ("a_pre", c_void_p), ("Wp", c_void_p), ("bias", c_void_p),
                ("addend", c_void_p), ("Y", c_void_p), ("y_pre", c_void_p), ("g0", c_void_p), ("g1", c_void_p),
                ("g2", c_void_p), ("g3", c_void_p), ("o0", c_void_p), ("o1", c_void_p)]


# name -> argument ctypes (return type is int unless listed in _RESTYPE)
SIGNATURES = {
    "spk_version": [],
    "spk_graph_workspace_bytes": [c_int64, c_int64],
    "spk_graph_build": [P, P, c_int64, c_int64, P, P, P, P, P, P, P, P, c_size_t, P],
    "spk_graph_build_active": [P, P, P, c_float, c_int64, c_int64, P, P, P, P, P, P, P, P, c_size_t, P],
    "spk_segment_ptr": [P, c_int64, c_int64, P, P],
    "spk_pairwise_fwd": [P, P, P, P, c_int64, c_int64, P, P],
    "spk_pairwise_bwd": [P, P, P, P, P, c_int64, c_float, P, P],
    "spk_edge_geometry": [P, P, c_int64, c_int, c_int, P, P, c_float, P, P, P, P, P],
    "spk_rbf_fwd": [P, c_int64, c_int, c_int, P, P, P, P, P],
    "spk_cosine_cutoff_fwd": [P, c_int64, c_float, P, P, P],
    "spk_act_fwd": [P, c_int64, c_int, P, P, P],
    "spk_embedding": [P, P, c_int64, c_int, c_int, P, P],
    "spk_segment_sum": [P, P, P, c_int64, c_int, P, P],
    "spk_dense": [P, c_int64, c_int, c_int64, P, c_int, P, c_int, P, c_int, P, c_int64, P, c_int64, P, P],
    "spk_tc_packed_floats": [c_int, c_int],
    "spk_tc_packed_floats_tn": [c_int, c_int, c_int],
    "spk_atom_chain_workspace_ints": [c_int, c_int64],
    "spk_atom_chain": [POINTER(ChainStep), c_int, c_int64, P, c_size_t, c_int, P],
    "spk_atom_chain_debug": [POINTER(ChainStep), c_int, c_int64, P, c_size_t, c_int, P, P],
    "spk_tc_pack_weight": [P, c_int, c_int, P, P],
    "spk_dense_tc": [P, c_int64, c_int, c_int64, P, c_int, P, c_int, P, c_int, P, c_int64, P, c_int64, P, P],
    "spk_mlp2_tc": [P, c_int64, c_int, c_int64, P, P, c_int, P, c_int, P, P, c_int64, P, c_int64, P, P],
    "spk_painn_edge_fwd": [P, P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, c_int, P, P, P],
    "spk_painn_edge_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, c_int, P, P, P, c_int, P],
    "spk_painn_filter_packed_floats": [],
    "spk_painn_pack_filter": [P, P, c_int, c_int, P, P],
    "spk_painn_edge_fwd_tc": [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, c_int, P, P, P],
    "spk_painn_edge_bwd_tc": [P, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, c_int, P, P, P, c_int, P],
    "spk_painn_edge_wij_fwd": [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, P, P, P],
    "spk_painn_edge_wij_bwd": [P, P, P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, P, P, P, P, P],
    "spk_painn_mix_ctx": [P, P, c_int64, c_int, c_float, P, P],
    "spk_painn_mix_update": [P, P, P, P, c_int64, c_int, P, P, P],
    "spk_painn_mix_update_bwd": [P, P, P, P, c_int64, c_int, P, P, P],
    "spk_painn_mix_ctx_bwd": [P, P, P, c_int64, c_int, c_float, P, P, P],
    "spk_cfconv_fwd": [P, P, P, P, P, c_int64, c_int64, c_int, P, P],
    "spk_schnet_filter_packed_floats": [],
    "spk_schnet_pack_filter": [P, P, P, c_int, c_int, P, P],
    "spk_schnet_cfconv_fwd_tc": [P, P, P, P, P, P, P, c_int, c_int64, c_int64, c_int, c_int, P, P],
    "spk_cfconv_bwd": [P, P, P, P, P, P, P, c_int64, c_int64, c_int, P, P, P, P],
    "spk_radial_bwd": [P, P, P, P, P, c_int64, c_int, P, c_int, P],
    "spk_atomwise_out": [P, P, P, P, c_int64, c_int64, c_int, P, P, P],
    "spk_atomwise_out_bwd": [P, P, P, c_int64, c_int, P, P],
    "spk_add": [P, P, c_int64, P, P],
    "spk_halo_pull": [P, P, P, P, c_int64, c_int, P],
    "spk_halo_pull_add": [P, P, P, P, P, P, c_int64, c_int, P],
    "spk_md_velocity_verlet": [P, P, P, P, P, c_int64, c_float, c_float, c_float, c_int, P],
    "spk_neighbor_list_workspace_bytes": [c_int64, c_int64],
    "spk_neighbor_list": [P, P, P, P, c_int64, c_int64, c_float, c_int64, c_int, P, P, P, P, P, P, c_size_t, P],
}
_RESTYPE = {"spk_graph_workspace_bytes": c_size_t, "spk_tc_packed_floats": c_size_t,
            "spk_tc_packed_floats_tn": c_size_t, "spk_atom_chain_workspace_ints": c_size_t,
            "spk_schnet_filter_packed_floats": c_size_t,
            "spk_painn_filter_packed_floats": c_size_t, "spk_neighbor_list_workspace_bytes": c_size_t}

_lib = None


class SpkLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises SpkLibraryError if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpkLibraryError(
            f"{LIB_PATH} not found: build it with `python -m schnetpack_b200.build` (needs nvcc). "
            "schnetpack_b200 has no CPU fallback."
        )
    h = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(h, name)  # AttributeError if the .so is stale / symbol missing
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, c_int)
    _lib = h
    return h


def check(rc: int, name: str):
    if rc == SPK_OK:
        return
    if rc == -1:
        raise ValueError(f"{name}: invalid argument (SPK_ERR_ARG)")
    if rc == -2:
        raise NotImplementedError(f"{name}: shape outside the compiled sm_100a templates (SPK_ERR_UNSUPPORTED)")
    raise RuntimeError(f"{name}: CUDA error {-rc - 1000}")


# kernels launched per C-ABI call (default 1); bench.py reports the running total as ``gpu_launches``
LAUNCHES = {"spk_graph_build": 8, "spk_graph_build_active": 8, "spk_atomwise_out": 2, "spk_neighbor_list": 6}
launch_count = 0


def call(name: str, *args):
    global launch_count
    rc = getattr(lib(), name)(*args)
    check(rc, name)
    launch_count += LAUNCHES.get(name, 1)
