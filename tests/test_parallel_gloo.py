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
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1    # "reference" when oracle/_ref or /root/reference is present


# ------------------------------------------------------------------------------------ one large system over several ranks
def test_partition_graph_covers_every_edge_once_and_exchange_lists_match():
    sys.path.insert(0, ROOT)
    from schnetpack_b200 import parallel as P
    from schnetpack_b200 import synthetic as S

    _, box = S.make_config("cfg4", n_atoms_total=300)
    R, ii, jj = box["_positions"], box["_idx_i"], box["_idx_j"]
    for world in (2, 3, 5):
        owner = P.slab_owners(R, world)
        assert np.bincount(owner, minlength=world).min() >= R.shape[0] // world
        plans = [P.partition_graph(owner, ii, jj, r, world) for r in range(world)]
        assert sorted(np.concatenate([p.edge_ids for p in plans]).tolist()) == list(range(ii.shape[0]))
        for p in plans:
            glob = np.concatenate([p.owned, p.ghosts])
            assert (glob[p.idx_i] == ii[p.edge_ids]).all() and (glob[p.idx_j] == jj[p.edge_ids]).all()
            assert (p.idx_i < p.n_owned).all()
            for peer, (a, b) in p.recv.items():              # what I expect from peer == what peer plans to send me
                sent = plans[peer].owned[plans[peer].send[p.rank]]
                assert (sent == p.ghosts[a:b]).all()


def _halo_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import dist_oracle as D
    from oracle import spk_oracle as O
    from schnetpack_b200 import parallel as P
    from schnetpack_b200 import synthetic as S

    spec, box = S.make_config("cfg4", n_atoms_total=240)
    spec = dict(spec, n_atom_basis=32, n_interactions=2)
    params = S.init_params(spec, seed=5)
    owner = P.slab_owners(box["_positions"], world)
    plan = P.partition_graph(owner, box["_idx_i"], box["_idx_j"], rank, world)
    e_part, f_own = D.painn_energy_forces(spec, params, box, plan)
    dist.all_reduce(e_part)                                   # total energy = sum of the ranks' partial sums
    # gather the owned force blocks on rank 0
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([plan.n_owned]))
    mx = int(max(s.item() for s in sizes))
    pad_f = torch.zeros((mx, 3), dtype=f_own.dtype)
    pad_f[: plan.n_owned] = f_own
    pad_i = torch.full((mx,), -1, dtype=torch.int64)
    pad_i[: plan.n_owned] = torch.as_tensor(plan.owned)
    fs = [torch.empty_like(pad_f) for _ in range(world)]
    ids = [torch.empty_like(pad_i) for _ in range(world)]
    dist.all_gather(fs, pad_f)
    dist.all_gather(ids, pad_i)
    if rank == 0:
        forces = torch.zeros((box["_positions"].shape[0], 3), dtype=f_own.dtype)
        for f, i in zip(fs, ids):
            ok = i >= 0
            forces[i[ok]] = f[ok]
        ref = O.energy_forces(spec, params, box, dtype=torch.float64)
        np.savez(out_path, e=e_part.numpy(), f=forces.numpy(), e_ref=ref["energy"].numpy(), f_ref=ref["forces"].numpy(),
                 n_ghost=plan.n_ghost)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_one_system_over_several_ranks_with_halo_exchange(tmp_path, world):
    """Graph partition + autograd-aware halo exchange (positions, x, mu forward; their gradients backward) reproduce the
    single-process energy and forces of one periodic box."""
    out = str(tmp_path / "halo.npz")
    port = _free_port()
    mp.spawn(_halo_worker, args=(world, port, out), nprocs=world, join=True)
    z = np.load(out)
    assert int(z["n_ghost"]) > 0
    np.testing.assert_allclose(z["e"], z["e_ref"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(z["f"], z["f_ref"], rtol=1e-9, atol=1e-11)


def test_peer_halo_index_lists_address_the_right_rows():
    """The peer-memory halo (csrc/halo.cu) reads ghost rows out of their owners' tables: ``ghost_rank`` / ``ghost_row`` must
    address exactly the global rows ``ghosts``, and the reverse lists (``send``, ``send_pos``) must address, in every peer's
    ghost block, exactly the copies of this rank's rows.  Emulated with numpy over all ranks of 2-, 3- and 5-way partitions."""
    sys.path.insert(0, ROOT)
    from schnetpack_b200 import parallel as P
    from schnetpack_b200 import synthetic as S

    box = S.periodic_box(400, seed=31)
    n = box["_positions"].shape[0]
    X = np.random.default_rng(0).normal(size=(n, 5))
    for world in (2, 3, 5):
        owner = P.slab_owners(box["_positions"], world)
        plans = [P.partition_graph(owner, box["_idx_i"], box["_idx_j"], r, world) for r in range(world)]
        tables = [X[pl.owned] for pl in plans]                                    # every rank's owned rows, local order
        for r, pl in enumerate(plans):
            pulled = np.stack([tables[p][row] for p, row in zip(pl.ghost_rank, pl.ghost_row)]) if pl.n_ghost else np.zeros((0, 5))
            assert np.array_equal(pulled, X[pl.ghosts])                           # forward pull
            # reverse: the gradient of my owned row = sum over the peers' ghost copies, located by (send[p], send_pos[p])
            G = [np.random.default_rng(10 + q).normal(size=(plans[q].n_ghost, 5)) for q in range(world)]
            acc = np.zeros((pl.n_owned, 5))
            for p in sorted(pl.send):
                k = np.arange(len(pl.send[p]))
                assert np.array_equal(plans[p].ghosts[pl.send_pos[p] + k], pl.owned[pl.send[p]])   # same atoms
                np.add.at(acc, pl.send[p], G[p][pl.send_pos[p] + k])
            want = np.zeros((n, 5))
            for q in range(world):
                if q != r and plans[q].n_ghost:
                    np.add.at(want, plans[q].ghosts, G[q])
            assert np.allclose(acc, want[pl.owned])
