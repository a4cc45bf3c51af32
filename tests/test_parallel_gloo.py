"""CPU, world_size 2 over gloo: the N>1 host logic -- sharding a collated batch by systems (index re-basing) and
gathering the results -- reproduces the single-process result.  The per-rank compute engine here is the oracle (CPU);
on GPUs the same functions feed the CUDA modules (bench.py --gpus N uses per-rank batches, no data-path collective)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import spk_oracle as O
    from schnetpack_b200 import parallel as P
    from schnetpack_b200 import synthetic as S

    spec = S.model_spec("painn", n_atom_basis=64, n_interactions=2)
    params = S.init_params(spec, seed=21)
    full = S.qm9like_batch(11, seed=22)          # ragged systems, 11 does not divide by 2
    shard = P.shard_batch(full, rank, world)
    res = O.energy_forces(spec, params, shard, dtype=torch.float64)
    local = {"energy": res["energy"], "forces": res["forces"]}
    gathered = P.gather_results(local, int(full["_n_atoms"].shape[0]), int(full["_atomic_numbers"].shape[0]),
                                int(shard["_atomic_numbers"].shape[0]))
    if rank == 0:
        ref = O.energy_forces(spec, params, full, dtype=torch.float64)
        np.savez(out_path, e=gathered["energy"].numpy(), f=gathered["forces"].numpy(), e_ref=ref["energy"].numpy(),
                 f_ref=ref["forces"].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_batch_covers_everything_once():
    sys.path.insert(0, ROOT)
    from schnetpack_b200 import parallel as P
    from schnetpack_b200 import synthetic as S

    full = S.qm9like_batch(13, seed=3)
    for world in (1, 2, 3, 4, 8):
        tot_a = tot_e = tot_m = 0
        for r in range(world):
            sh = P.shard_batch(full, r, world)
            na = sh["_atomic_numbers"].shape[0]
            assert int(sh["_n_atoms"].sum()) == na
            if sh["_idx_i"].size:
                assert sh["_idx_i"].min() >= 0 and sh["_idx_i"].max() < na and sh["_idx_j"].max() < na
                # geometry preserved: r_ij identical to the full batch's
            tot_a += na
            tot_e += sh["_idx_i"].shape[0]
            tot_m += sh["_n_atoms"].shape[0]
        assert tot_a == full["_atomic_numbers"].shape[0] and tot_e == full["_idx_i"].shape[0]
        assert tot_m == full["_n_atoms"].shape[0]


@pytest.mark.timeout(300)
def test_two_rank_sharded_evaluation_matches_single_process(tmp_path):
    out = str(tmp_path / "res.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    np.testing.assert_allclose(z["e"], z["e_ref"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(z["f"], z["f_ref"], rtol=1e-10, atol=1e-12)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm) needs no GPU and prints exactly one JSON
    line with the contract keys; under N > 1 only rank 0 would print (exercised here at N = 1)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "cfg1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline",
              "e2e"):
        assert k in d, k
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
