#!/bin/bash
# round 2, GPU run 4 (2 GPUs): full GPU suite incl. the 2-GPU tests, bench at N=1 and N=2 (multi-GPU self-check)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="--timeout 240 --timeout-method=thread"
nvidia-smi -L > gpurun_out/r4_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q $T > gpurun_out/r4_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4_pytest.log
tail -12 gpurun_out/r4_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r4_bench1.json 2> gpurun_out/r4_bench1.err; echo "bench rc=$?" >> gpurun_out/r4_bench1.err
tail -c 3000 gpurun_out/r4_bench1.json; tail -3 gpurun_out/r4_bench1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r4_bench2.json 2> gpurun_out/r4_bench2.err; echo "bench rc=$?" >> gpurun_out/r4_bench2.err
tail -c 5000 gpurun_out/r4_bench2.json; tail -5 gpurun_out/r4_bench2.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r4_ref.json 2> gpurun_out/r4_ref.err; tail -c 600 gpurun_out/r4_ref.json
