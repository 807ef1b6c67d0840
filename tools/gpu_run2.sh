#!/bin/bash
# round 2, GPU run 2: quick kernel parity first (abort early when it fails), then the full suite, variant sweep, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="--timeout 150 --timeout-method=thread"
timeout 420 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q $T > gpurun_out/r2_kernels.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r2_kernels.log
tail -5 gpurun_out/r2_kernels.log
if [ $rc -ne 0 ]; then echo "kernel tests failed - stopping"; exit 0; fi
timeout 300 python tools/variant_sweep.py --n 512 ZKB_AFFINE=1,0 ZKB_ACC_VARIANT_G2=0,2 > gpurun_out/r2_sweep_a.log 2>&1; cat gpurun_out/r2_sweep_a.log
timeout 1100 python -m pytest tests -m gpu -q $T --deselect tests/test_gpu_kernels.py > gpurun_out/r2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
timeout 420 python tools/variant_sweep.py --n 512 ZKB_ACC_VARIANT_G2=3,4,5,6,7 > gpurun_out/r2_sweep_b.log 2>&1; cat gpurun_out/r2_sweep_b.log
timeout 300 python tools/variant_sweep.py --n 512 ZKB_AFFINE_LEVELS=2,4 ZKB_AFFINE_GROUP=256,1024 > gpurun_out/r2_sweep_c.log 2>&1; cat gpurun_out/r2_sweep_c.log
timeout 480 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err
tail -c 6000 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
