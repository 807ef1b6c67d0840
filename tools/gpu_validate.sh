#!/bin/bash
# Full validation on one B200 (what profiles/r02_* were made with): GPU test suite, bench, raw sweep with the
# variable-base column, ncu launch list + full captures of the accumulation / NTT kernels.
#   gpurun --timeout 4200 -- "bash tools/gpu_validate.sh"
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="--timeout 240 --timeout-method=thread"
timeout 1300 python -m pytest tests -m gpu -q $T > gpurun_out/r6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r6_pytest.log
tail -6 gpurun_out/r6_pytest.log
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; echo "bench rc=$?" >> gpurun_out/r6_bench.err
tail -c 1500 gpurun_out/r6_bench.json; tail -3 gpurun_out/r6_bench.err
timeout 600 python tools/sweep_raw.py --msm 16,18,20,22,24 --variable-base --ntt 16,18,20,22,24 --ntt-dist 24 --out gpurun_out/r02_sweep_1gpu.jsonl > gpurun_out/r6_sweep.log 2>&1; tail -3 gpurun_out/r6_sweep.log | cut -c1-400
# ncu: launch list of a bench run (no extras), then full captures of the top kernels on one 128-proof chunk
export ZKB_SKIP_DEPTHS=1 ZKB_CPU_SAMPLE=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r6_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_accumulate' -c 12 -f -o gpurun_out/r02_acc python tools/prof_driver.py 128 1 > gpurun_out/r6_ncu_acc.log 2>&1; echo "ncu acc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ntt_pass|k_digits|k_reduce1' -s 20 -c 8 -f -o gpurun_out/r02_ntt python tools/prof_driver.py 128 1 > gpurun_out/r6_ncu_ntt.log 2>&1; echo "ncu ntt rc=$?"
ls -la gpurun_out/*.ncu-rep
