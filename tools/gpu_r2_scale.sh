#!/bin/bash
# round 2, scaling run on N GPUs of one box: usage  bash tools/gpu_r2_scale.sh N
# (1) the default driver command (cfg2 batch-sharded, + the cfg5 strong-scaling leg under key "spatial"), (2) cfg5 alone with
# the halo over NCCL point-to-point for comparison
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -$N
run() { # name, extra env, args
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N $3 > gpurun_out/r2s_$1_n$N.json 2> gpurun_out/r2s_$1_n$N.err
  echo "$1 N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2s_$1_n$N.json"))
    sp = d.get("spatial") or (d if d.get("scaling") == "strong" else None)
    print("  value", round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2))
    if sp:
        print("  spatial:", sp.get("error") or (round(sp["ms_per_step"], 2), "ms/step", sp["rank0"], sp["halo"]["transport"][:24], "halo ms", round(sp["halo"]["ms_per_step_max_over_ranks"], 3),
              "share", round(sp["halo"]["share_of_step"], 4), {k: (round(v["avg_us"], 1), round(v["frac"], 3)) for k, v in sp["roofline_all"].items()}))
except Exception as e:
    print("  failed:", e); print(open("gpurun_out/r2s_$1_n$N.err").read()[-1200:])
PY
}
run default "SPK_B200_HALO=peer" "--steps 50 --warmup 5 --no-cpu-baseline"
run cfg5nccl "SPK_B200_HALO=nccl" "--config cfg5 --steps 6 --warmup 3 --no-cpu-baseline"
