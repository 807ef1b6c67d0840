#!/bin/bash
# round 2, GPU run 1: parity of everything new, G2 variant sweep, full bench at N=1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
timeout 600 python tools/variant_sweep.py ZKB_ACC_VARIANT_G2=2,1,0,3,4,5,6,7 > gpurun_out/r1_g2_sweep.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?" >> gpurun_out/r1_bench.err
tail -3 gpurun_out/r1_pytest.log; cat gpurun_out/r1_g2_sweep.log; tail -c 3000 gpurun_out/r1_bench.json; tail -5 gpurun_out/r1_bench.err
