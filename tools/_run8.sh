cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/multi_bench_8gpu.json 2> gpurun_out/multi_bench_8gpu.err; echo "bench rc=$?"
python - <<PY
import json
b=json.load(open("gpurun_out/multi_bench_8gpu.json"))
print("value",b["value"],"e2e",b["e2e"]["value"],"verified",b["verified"],"/",b["verified_of"])
print(json.dumps(b["multi_gpu_selfcheck"]))
PY
tail -3 gpurun_out/multi_bench_8gpu.err
