#!/bin/bash
# Multi-GPU bench runs (one box, torchrun): usage scripts/gpu_multi.sh <N> <config> [extra bench args]
N=$1; CFG=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $CFG --steps 10 --warmup 3 --no-torch-gpu --cpu-baseline-steps 0 --no-pruned "$@" > gpurun_out/bench_cfg${CFG}_N${N}.json 2> gpurun_out/bench_cfg${CFG}_N${N}.err
echo "cfg $CFG N $N rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg${CFG}_N${N}.json"))
print({k:d[k] for k in ("metric","value","n_gpus","ms_per_step","scaling")}, d["e2e"]["value"], d["config"]["per_gpu_batch"], d["config"]["global_batch"], d.get("strong"), d["metric_reduction"])
PY
tail -2 gpurun_out/bench_cfg${CFG}_N${N}.err
