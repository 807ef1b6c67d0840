#!/bin/bash
# Multi-GPU validation: bench.py at N GPUs (weak scaling + multi_gpu_selfcheck) and the raw sweep split over the ranks.
#   gpurun --gpus 4 --timeout 2400 -- "bash tools/gpu_multi.sh 4"
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/multi_bench_${N}gpu.json 2> gpurun_out/multi_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
b=json.load(open("gpurun_out/multi_bench_${N}gpu.json"))
print("value",b["value"],"e2e",b["e2e"]["value"],"verified",b["verified"],"/",b["verified_of"])
print(json.dumps(b["multi_gpu_selfcheck"]))
PY
tail -3 gpurun_out/multi_bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/sweep_raw.py --msm 20,22,24,26 --variable-base --ntt 20 --ntt-dist 24,26 --out gpurun_out/r02_sweep_${N}gpu.jsonl > gpurun_out/multi_sweep_${N}gpu.log 2>&1; echo "sweep rc=$?"; tail -4 gpurun_out/multi_sweep_${N}gpu.log | cut -c1-300
