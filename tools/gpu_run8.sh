#!/bin/bash
# round 2, GPU run 8: second form of the pair tree (separate index / divisor / inverse / add kernels)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="--timeout 240 --timeout-method=thread"
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -q $T -k "pair_tree" > gpurun_out/r8_pytest.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r8_pytest.log
tail -5 gpurun_out/r8_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python tools/variant_sweep.py --n 1024 ZKB_AFFINE=0,2 > gpurun_out/r8_sweep_a.log 2>&1; cat gpurun_out/r8_sweep_a.log
ZKB_AFFINE=2 timeout 900 python tools/variant_sweep.py --n 1024 ZKB_AFFINE_LEVELS=2,3,4 ZKB_AFFINE_GROUP=128,256,512 > gpurun_out/r8_sweep_b.log 2>&1; cat gpurun_out/r8_sweep_b.log
