#!/bin/bash
# round 2, GPU run 3: full GPU suite after the memory fixes, pair-tree tuning sweep at the bench's shape, bench N=1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="--timeout 200 --timeout-method=thread"
timeout 1300 python -m pytest tests -m gpu -q $T > gpurun_out/r3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r3_pytest.log
tail -12 gpurun_out/r3_pytest.log
timeout 420 python tools/variant_sweep.py --n 1024 ZKB_AFFINE=0 > gpurun_out/r3_sweep_a.log 2>&1; cat gpurun_out/r3_sweep_a.log
timeout 600 python tools/variant_sweep.py --n 1024 ZKB_AFFINE_LEVELS=2,3 ZKB_AFFINE_GROUP=192,256,384,512 > gpurun_out/r3_sweep_b.log 2>&1; cat gpurun_out/r3_sweep_b.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "bench rc=$?" >> gpurun_out/r3_bench.err
tail -c 7000 gpurun_out/r3_bench.json; tail -5 gpurun_out/r3_bench.err
