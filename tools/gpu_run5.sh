#!/bin/bash
# round 2, GPU run 5: priority-stream split (latency-bound kernels ahead of queued accumulation CTAs), lanes, suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="--timeout 240 --timeout-method=thread"
timeout 500 python tools/variant_sweep.py --n 1024 ZKB_PRIO_SPLIT=0,1 ZKB_LANES=4,6 > gpurun_out/r5_sweep.log 2>&1; cat gpurun_out/r5_sweep.log
timeout 1200 python -m pytest tests/test_gpu_prover.py tests/test_gpu_round2.py -m gpu -q $T > gpurun_out/r5_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r5_pytest.log
tail -8 gpurun_out/r5_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo "bench rc=$?" >> gpurun_out/r5_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r5_bench.json'))
print("value",b['value'],"e2e",b['e2e']['value'],"verified",b['verified'],"lat",b['latency_ms'])
PY
tail -3 gpurun_out/r5_bench.err
