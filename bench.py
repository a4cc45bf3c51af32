#!/usr/bin/env python
"""Benchmark of the SchNetPack message-passing hot path on B200 (driver contract: one JSON line on rank 0).

    python bench.py --gpus 1 --steps 20 --warmup 5                  # CUDA path (this repo)
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 1  # reference arm: CPU path on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...            # weak scaling: one batch per rank, no collective

Workload (BASELINE.json configs[1]): MD17 aspirin x 256, PaiNN F=128 T=3 (20 Gaussian RBF, cosine cutoff 5 A),
energy + forces.  One *step* = one full evaluation ``model(inputs)`` of the whole batch (256 molecule-evals).
``value`` = molecule-evals/s with inputs resident in HBM; ``e2e`` = the same call fed from pinned HOST buffers with the
H2D copies of the batch and the D2H read of energy+forces inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from schnetpack_b200 import synthetic as S  # noqa: E402

METRIC = "energy+force evals/s"
UNIT = "molecule-evals/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload(name: str, rank: int, batch: int | None):
    kw = {}
    if name in ("cfg1", "cfg2", "cfg3") and batch is not None:
        kw["batch"] = batch
    if name in ("cfg2", "cfg3"):
        kw["seed"] = rank
    spec, data = S.make_config(name, **kw)
    return spec, data


def n_systems(data):
    return int(data[S.n_atoms].shape[0])


# ------------------------------------------------------------------------------------------------- algorithmic bytes
def edge_kernel_bytes(E, N, F, has_mu, backward):
    """SURVEY.md §8(d) 'no-reuse gather model' per launch of the fused PaiNN edge kernel.
    forward : E*(8+8+12 + 4*3F [x_j] + 4*3F [mu_j]) + N*(4*4F read q,mu + 4*4F write)      = E*3100 + N*4096 (F=128)
    backward: E*(28 + 2*4*3F regather + 2*4*3F sender-grad accumulate + 12) + N*(2*4*4F)   = E*6184 + N*4096
    first block (mu == 0): no mu_j gather and no mu-dependent third."""
    if not backward:
        per_e = 28 + 4 * 3 * F + (4 * 3 * F if has_mu else 0)
        per_n = 4 * 4 * F * 2 if has_mu else 4 * F + 4 * 4 * F
    else:
        per_e = 28 + 2 * 4 * 3 * F + (2 * 4 * 3 * F if has_mu else 0) + 12
        per_n = 2 * 4 * 4 * F
    return E * per_e + N * per_n


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- reference (CPU) arm
def oracle_eval_time(spec, params, data, steps, warmup, threads):
    from oracle import spk_oracle as O

    torch.set_num_threads(threads)
    ts = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.energy_forces(spec, params, data, dtype=torch.float32)
        dt = time.perf_counter() - t0
        if it >= warmup:
            ts.append(dt)
    return ts


def best_thread_count(spec, params, data):
    """The reference's eager CPU path scales badly past a few dozen threads on these small ops (128 threads were 13x
    slower than 8 on the round-1 box), so give it the thread count it is fastest with: probe 8/16/32/64/all once each
    and keep the best.  Returns (threads, seconds_per_eval_at_that_count)."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores} or {cores})
    best = None
    first = True
    for c in cands:
        t = min(oracle_eval_time(spec, params, data, 1, 1 if first else 0, c))
        first = False
        if best is None or t < best[1]:
            best = (c, t)
        elif t > 1.5 * best[1]:
            break
    return best


def run_reference(args, rank, world):
    """The reference's CPU path = the same ATen op sequence, restated in oracle/spk_oracle.py (kind 'port': the Python
    reference cannot travel to the GPU box), timed on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    spec, data = workload(args.config, 0, args.batch)
    params = S.init_params(spec, seed=0)
    cores, _ = best_thread_count(spec, params, data)
    ts = oracle_eval_time(spec, params, data, args.steps, args.warmup, cores)
    B = n_systems(data)
    E = int(data[S.idx_i].shape[0])
    ms = 1e3 * float(np.mean(ts))
    v = B / float(np.mean(ts))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {S.CONFIGS[args.config]['desc']}", "systems": B,
                   "atoms": int(data[S.Z].shape[0]), "edges": E},
        "edge_msgs_per_s": E * spec["n_interactions"] / float(np.mean(ts)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full-batch evals ({B} systems) after {args.warmup} warm-up, "
                                   f"oracle/spk_oracle.py fp32 (same ATen op sequence as the reference), "
                                   f"torch.set_num_threads({cores}) = fastest of 8/16/32/64/{os.cpu_count()} probed; median {1e3 * float(np.median(ts)):.1f} ms, "
                                   f"min {1e3 * float(np.min(ts)):.1f} ms"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args, rank, world, local_rank):
    import schnetpack_b200 as sb
    from schnetpack_b200 import _lib, ops
    from schnetpack_b200.model import batch_to_device, from_spec

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL's version banner goes to stdout: keep the ONE JSON line clean
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)

    spec, data = workload(args.config, rank, args.batch)
    params = S.init_params(spec, seed=0)
    model = from_spec(spec, params, dev)
    B, N, E = n_systems(data), int(data[S.Z].shape[0]), int(data[S.idx_i].shape[0])
    F, T = spec["n_atom_basis"], spec["n_interactions"]
    want_forces = bool(spec.get("forces", True))
    padded = S.Rij in data

    def evaluate(x):
        if padded:
            x = model.representation(x)
            return model.output_modules[0](x)
        return model(x)

    resident = batch_to_device(data, dev)

    def fresh(x):
        # the model writes into the dict and sets requires_grad on positions: hand it a shallow copy each step
        y = dict(x)
        ops._GRAPH_CACHE.clear()  # every step rebuilds the CSR/sender views from idx_i/idx_j (no cached work)
        if S.R in y:
            y[S.R] = y[S.R].detach()
        return y

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush_buf.fill_(1)

    # ---- per-kernel live timing of the fused edge kernels (CUDA events on the launch stream) -------------------
    # With the CUDA-graph path the events are "external" events: recorded inside the capture they become event-record
    # nodes of the graph, so every replay of the timed region re-records them around the edge-kernel nodes.
    from schnetpack_b200.model import GraphedPotential

    use_graph = not (padded or args.no_graph)
    ev = {"fwd": [], "bwd": []}
    orig_fwd, orig_bwd = ops.painn_edge_fwd, ops.painn_edge_bwd
    timing = {"on": False}

    def new_event():
        return torch.cuda.Event(enable_timing=True, external=True) if use_graph else torch.cuda.Event(True)

    def timed_fwd(x, mu, *a, **k):
        if not timing["on"]:
            return orig_fwd(x, mu, *a, **k)
        s, e = new_event(), new_event()
        s.record()
        r = orig_fwd(x, mu, *a, **k)
        e.record()
        ev["fwd"].append((s, e, mu is not None))
        return r

    def timed_bwd(x, mu, *a, **k):
        if not timing["on"]:
            return orig_bwd(x, mu, *a, **k)
        s, e = new_event(), new_event()
        s.record()
        r = orig_bwd(x, mu, *a, **k)
        e.record()
        ev["bwd"].append((s, e, mu is not None))
        return r

    ops.painn_edge_fwd, ops.painn_edge_bwd = timed_fwd, timed_bwd

    # ---- device-resident timing -----------------------------------------------------------------------------------
    # One step = one E+F evaluation of the resident batch, graph-view build included.  Default: replay of the captured
    # evaluation (GraphedPotential, the product's steady-state path -- the eager Python loop is launch-bound on the host
    # at this kernel granularity); --no-graph times eager model(inputs) calls instead.
    kernel_ms = {"fwd": [], "bwd": []}          # (ms, has_mu) per timed edge-kernel launch
    evaluate(fresh(resident))                   # one-time work (weight packing, kernel attributes) outside the counts
    torch.cuda.synchronize()
    timing["on"] = spec["kind"] == "painn"
    if use_graph:
        g_res = GraphedPotential(model)
        c0 = _lib.launch_count
        g_res(resident)                          # 2 eager warm-ups + capture
        launches_per_step = (_lib.launch_count - c0) // (g_res.warmup + 1)
        n_bwd = len(ev["bwd"]) // (g_res.warmup + 1)
        n_fwd = len(ev["fwd"]) // (g_res.warmup + 1)
        pairs = {"fwd": ev["fwd"][-n_fwd:] if n_fwd else [], "bwd": ev["bwd"][-n_bwd:] if n_bwd else []}

        def step():
            return g_res.replay()
    else:
        launches_per_step = None
        pairs = None

        def step():
            return evaluate(fresh(resident))

    sampler = ClockSampler(local_rank) if rank == 0 else None   # polls every 100 ms from the warm-up on (GPU under load)
    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    ev["fwd"].clear()
    ev["bwd"].clear()
    if dist is not None:
        dist.barrier()
    launches0 = _lib.launch_count
    step_ev = []
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        out = step()
        e.record()
        step_ev.append((s, e))
        if use_graph:                           # the external events are re-recorded by every replay: read them now
            torch.cuda.synchronize()
            for kind in ("fwd", "bwd"):
                kernel_ms[kind] += [(a.elapsed_time(b), hm) for a, b, hm in pairs[kind]]
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = launches_per_step * args.steps if use_graph else _lib.launch_count - launches0
    timing["on"] = False
    if not use_graph:
        for kind in ("fwd", "bwd"):
            kernel_ms[kind] = [(a.elapsed_time(b), hm) for a, b, hm in ev[kind]]
    clocks = sampler.stop() if sampler is not None else None
    dev_ms = sum(s.elapsed_time(e) for s, e in step_ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- end-to-end: pinned host buffers -> H2D -> model -> D2H(energy, forces) -------------------------------------
    host = {}
    for k, v in data.items():
        tt = torch.as_tensor(v)
        if tt.is_floating_point():
            tt = tt.float()
        host[k] = tt.pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    e_host = torch.empty(B, dtype=torch.float32).pin_memory()
    f_host = torch.empty((N, 3), dtype=torch.float32).pin_memory() if want_forces else None
    d2h = e_host.numel() * 4 + (f_host.numel() * 4 if f_host is not None else 0)

    graphed = GraphedPotential(model) if use_graph else None

    def e2e_step():
        # public API call a user makes: host batch in, energy/forces out.  GraphedPotential copies the pinned host
        # tensors into its static device buffers (H2D, non-blocking) and replays the captured evaluation.
        if graphed is not None:
            o = graphed(host)
        else:
            x = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            o = evaluate(x)
        e_host.copy_(o["energy"], non_blocking=True)
        if f_host is not None:
            f_host.copy_(o["forces"], non_blocking=True)

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e2e_ev = []
    for _ in range(args.steps):
        flush_l2()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        e2e_step()
        e.record()
        e2e_ev.append((s, e))
    torch.cuda.synchronize()
    t2 = torch.tensor([sum(s.elapsed_time(e) for s, e in e2e_ev)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.barrier()
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t2.item()) * 1e-3)
    e2e_steps_ms = sorted(s.elapsed_time(e) for s, e in e2e_ev)

    if rank != 0:
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    roof = None
    roof_all = {}
    if kernel_ms["fwd"] or kernel_ms["bwd"]:
        for kind in ("fwd", "bwd"):
            if not kernel_ms[kind]:
                continue
            tot_ms = sum(ms for ms, _ in kernel_ms[kind])
            tot_bytes = sum(edge_kernel_bytes(E, N, F, hm, kind == "bwd") for _, hm in kernel_ms[kind])
            n = len(kernel_ms[kind])
            ach = tot_bytes / (tot_ms * 1e-3) / 1e9
            roof_all[kind] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "traffic": None,
                              "kernel": f"k_painn_edge_{kind}" + ("_tc" if ops.edge_tc_ok(F, spec.get("n_rbf", 20), E) else ""),
                              "launches_timed": n,
                              "avg_us": 1e3 * tot_ms / n, "algorithmic_bytes_per_launch": tot_bytes / n,
                              "share_of_step": tot_ms / dev_ms, "peak_source": peak_src,
                              "model": "algorithmic bytes = no-reuse gather model of SURVEY 8(d): every gathered row counted as "
                                       "HBM traffic; rows served by the 126 MB L2 let frac approach or exceed 1 -- `traffic` is "
                                       "the DRAM bytes ncu measured",
                              "timing": "CUDA events around the kernel, inside the timed region"
                                        + (" (external event nodes of the replayed graph)" if use_graph else "")}
        dom = max(roof_all, key=lambda k: roof_all[k]["share_of_step"])
        roof = dict(roof_all[dom])
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr) and args.config == "cfg2" and args.batch in (None, 256):   # the ncu capture is of this workload
            try:
                roof["traffic"] = json.load(open(tr)).get(roof["kernel"])
            except Exception:
                pass

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) --------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores, _ = best_thread_count(spec, params, data)
        ts = oracle_eval_time(spec, params, data, 3, 0, cores)
        cpu = {"value": B / float(np.mean(ts)), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 full-batch evals ({B} systems, {E} edges) of oracle/spk_oracle.py fp32 (the reference's "
                         f"ATen op sequence) on {cores} threads (fastest of 8/16/32/64/{os.cpu_count()} probed on this "
                         f"host); mean {1e3 * float(np.mean(ts)):.0f} ms"}

    # ---- the reference's eager single-GPU path (oracle port = same ATen op sequence) on this B200, for the >=5x target ----
    eager = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import spk_oracle as O

            p_dev = O.to_torch(params, torch.float32, dev)
            x_dev = O.to_torch(data, torch.float32, dev)
            for _ in range(3):
                O.energy_forces(spec, p_dev, x_dev, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_it = 10
            for _ in range(n_it):
                O.energy_forces(spec, p_dev, x_dev, device=dev)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n_it
            eager = {"value": B / dt, "unit": UNIT, "ms_per_step": 1e3 * dt,
                     "kind": "port: oracle/spk_oracle.py (the reference modules' eager ATen op sequence, cuBLAS SGEMM, "
                             "index_select / index_add_ atomics, autograd backward) on the same B200, inputs resident"}
        except Exception as exc:  # pragma: no cover
            eager = {"error": repr(exc)[:200]}

    md_info = None
    if args.md and world == 1 and S.cell in data and not padded:
        # device-resident MD step (row f1 + f2): velocity Verlet + neighbour list rebuilt on the device + E/F, one CUDA graph
        from schnetpack_b200.md import DeviceMD
        from schnetpack_b200.neighbors import neighbor_list

        masses = torch.ones(N, device=dev)
        md = DeviceMD(model, resident, masses, time_step=1e-4, cutoff=spec["cutoff"], capacity=int(E * 1.15))
        md.run(5)
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(True), torch.cuda.Event(True)
        s_.record()
        md.run(50)
        e_.record()
        torch.cuda.synchronize()
        md_ms = s_.elapsed_time(e_) / 50
        nat = resident[S.n_atoms]
        for _ in range(3):
            neighbor_list(resident[S.R], resident.get(S.cell), resident.get(S.pbc), nat, spec["cutoff"], capacity=int(E * 1.15), pad=True)
        torch.cuda.synchronize()
        s_.record()
        for _ in range(20):
            neighbor_list(resident[S.R], resident.get(S.cell), resident.get(S.pbc), nat, spec["cutoff"], capacity=int(E * 1.15), pad=True)
        e_.record()
        torch.cuda.synchronize()
        md_info = {"ms_per_md_step": md_ms, "neighbor_list_ms": s_.elapsed_time(e_) / 20, "capacity": int(E * 1.15),
                   "pairs": int(md.n_pairs[0]), "overflow": int(md.n_pairs[1]),
                   "what": "velocity Verlet + device cell-list rebuild (every step, no skin) + E/F, CUDA-graph replay"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {S.CONFIGS[args.config]['desc']}", "systems_per_gpu": B, "atoms": N,
                   "edges": E, "n_atom_basis": F, "n_interactions": T, "parallelism": f"batch-sharded x{world}",
                   "l2": "256 MiB device memset between timed steps (outside the per-step event intervals)",
                   "step": ("CUDA-graph replay of model(inputs) on the resident batch (GraphedPotential.replay)" if use_graph
                            else "eager model(inputs) on the resident batch"),
                   "weights": "seeded xavier-uniform (synthetic.init_params)"},
        "edge_msgs_per_s": world * E * T * args.steps / (total_ms * 1e-3),
        "wall_ms_per_step": 1e3 * t_wall / args.steps,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "eager model(inputs)" if graphed is None else "GraphedPotential(model)(host_batch): CUDA-graph replay",
                "ms_per_step_median": e2e_steps_ms[len(e2e_steps_ms) // 2], "ms_per_step_max": e2e_steps_ms[-1]},
        "gpu_launches": launches,
        "roofline": roof, "roofline_all": roof_all, "cpu_baseline": cpu, "eager_gpu_baseline": eager,
        "impl_switches": {"dense": ops.DENSE_IMPL, "edge": ops.EDGE_IMPL},
    }
    if md_info is not None:
        line["md"] = md_info
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(S.CONFIGS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--md", action="store_true", help="also time the device-resident MD step (rows f1+f2); not part of the metric")
    ap.add_argument("--no-graph", action="store_true", help="time eager model(inputs) calls (value and e2e) instead of CUDA-graph replays")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    run_cuda(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
