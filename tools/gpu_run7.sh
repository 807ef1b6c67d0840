#!/bin/bash
# round 2, GPU run 7: what bounds the pair-tree kernels (ncu), the software-pipelined tape evaluator
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export ZKB_AFFINE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_affine_level|k_accumulate_pts' -c 4 -f -o gpurun_out/r02_tree python tools/prof_driver.py 128 1 > gpurun_out/r7_ncu_tree.log 2>&1; echo "ncu tree rc=$?"
unset ZKB_AFFINE
ncu -i gpurun_out/r02_tree.ncu-rep --page raw --csv > gpurun_out/r02_tree_raw.csv 2>/dev/null; rm -f gpurun_out/r02_tree.ncu-rep
T="--timeout 240 --timeout-method=thread"
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q $T -k "generic" > gpurun_out/r7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r7_pytest.log; tail -3 gpurun_out/r7_pytest.log
timeout 300 python - <<'PY' > gpurun_out/r7_generic.log 2>&1
import os, sys, json, time
sys.path.insert(0, os.getcwd())
import numpy as np
from zk_franchise_proof_circuit_b200 import prover, census_tree
art = "artifacts/zkCensus/dev/160"
t0 = time.perf_counter()
c = prover.load(open(art + "/proving_key.zkey", "rb").read(), open(art + "/circuit.wasm", "rb").read(), generic=True)
print("load_s", time.perf_counter() - t0)
c0 = prover.load(open(art + "/proving_key.zkey", "rb").read(), open(art + "/circuit.wasm", "rb").read())
vs = census_tree.gen_census(c0, 256, seed=1)
docs = [json.dumps(v).encode() for v in vs]
c.fullprove_batch(docs[:32])
for n in (128, 256):
    t0 = time.perf_counter(); p, q, s = c.fullprove_batch(docs[:n]); dt = time.perf_counter() - t0
    print(n, "proofs", dt, "s", n / dt, "proofs/s", sum(s))
# witness only: time 128 witnesses via the batch path's first stage is not exposed; use zkb_witness for one
t0 = time.perf_counter(); c.witness(docs[0]); print("one witness ms", (time.perf_counter() - t0) * 1e3)
PY
cat gpurun_out/r7_generic.log
