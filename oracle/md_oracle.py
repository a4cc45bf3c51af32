"""ORACLE -- test infrastructure, NOT product code: CPU restatement (torch float64) of the reference's NVE loop for the
"next" row f2 of SURVEY.md section 8 -- the simulator / integrator / calculator arithmetic that ``schnetpack_b200.md.DeviceMD``
replaces with device kernels.

Parity: UNPINNED by import -- ``schnetpack.md`` needs ``ase`` (units) at import time, which this image does not have, so the
three formulas are restated from the cited lines (they are one line each) and anchored on the reference's own structure:

  * ``md/simulator.py:118-144``   calculate -> { half_step, main_step, calculate, half_step } per step;
  * ``md/integrators.py:59-70``   half_step:  momenta += 0.5 * forces * time_step;
  * ``md/integrators.py:97-110``  VelocityVerlet._main_step:  positions += time_step * momenta / masses;
  * ``md/calculators/base_calculator.py:85-98,120-152,160-175``  the model sees positions * position_conversion, its energy is
    multiplied by energy_conversion and its forces by force_conversion = energy_conversion / position_conversion;
  * neighbour list per step: every pair within the model cutoff (``md/neighborlist_md.py`` keeps a list with cutoff + skin
    and rebuilds when an atom moved more than skin / 2 (``:55-98``); the model's cutoff function removes the pairs beyond the
    cutoff, so the forces equal those of an exact list -- which is what is built here, with oracle/nl_oracle.py).
The energy / force engine is oracle/spk_oracle.py (pinned to the reference).
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import nl_oracle as NL
from oracle import spk_oracle as O


def nve_trajectory(spec, params, batch, masses, momenta, time_step, n_steps, position_conversion=1.0,
                   energy_conversion=1.0):
    """Returns MD-unit positions, momenta and (energy in MD units, forces in MD units) after ``n_steps``.  ``batch`` holds ONE
    periodic or open system with ``_positions`` / ``_cell`` in MODEL units (as the CUDA driver takes them)."""
    f_conv = energy_conversion / position_conversion
    cell = np.asarray(batch["_cell"], dtype=np.float64).reshape(3, 3)
    pbc = np.asarray(batch["_pbc"]).reshape(3)
    x = torch.as_tensor(np.asarray(batch["_positions"], dtype=np.float64)) / position_conversion        # MD units
    p = torch.as_tensor(np.asarray(momenta, dtype=np.float64)).clone()
    m = torch.as_tensor(np.asarray(masses, dtype=np.float64)).reshape(-1, 1)

    def calculate(x_md):
        R = (x_md * position_conversion).numpy()                                     # base_calculator.py:160-175
        ii, jj, ss, off = NL.neighbor_list(R, cell, pbc, spec["cutoff"])
        b = dict(batch)
        b.update({"_positions": R, "_idx_i": ii, "_idx_j": jj, "_offsets": off})
        out = O.energy_forces(spec, params, b, dtype=torch.float64)
        return out["energy"] * energy_conversion, out["forces"] * f_conv             # base_calculator.py:96,137-152

    e, f = calculate(x)                                                              # simulator.py:118
    for _ in range(n_steps):
        p = p + 0.5 * f * time_step                                                  # integrators.py:70
        x = x + time_step * p / m                                                    # integrators.py:108
        e, f = calculate(x)                                                          # simulator.py:137
        p = p + 0.5 * f * time_step                                                  # simulator.py:144
    return x.numpy(), p.numpy(), e.numpy(), f.numpy()
