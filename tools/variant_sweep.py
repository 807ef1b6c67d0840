#!/usr/bin/env python3
"""Kernel-variant sweep (development aid): runs the resident census pipeline once per environment-variable setting in
a fresh process (the variants are read once per process), prints the eight stage times per 512 proofs, the overall
device-resident proofs/s and a digest of the proofs (pinned r, s: every variant must give the same digest).

    python tools/variant_sweep.py ZKB_ACC_VARIANT_G2=0,1,2,3 [ZKB_ACC_VARIANT=0,1] [--n 512]
"""
import hashlib
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("witness", "build_abc", "ntt_join", "msm_sort", "msm_acc_g1", "msm_acc_g2", "msm_reduce", "finalize")


def child(n):
    sys.path.insert(0, ROOT)
    import time
    import numpy as np
    from zk_franchise_proof_circuit_b200 import prover, census_tree
    art = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
    c = prover.load(open(art + "/proving_key.zkey", "rb").read(), open(art + "/circuit.wasm", "rb").read())
    vs = census_tree.gen_census(c, 128, seed=0xC0FFEE)
    packed = np.stack([prover.pack_inputs(vs[i % 128]) for i in range(n)])
    c.set_blinding(1234567, 7654321)
    c.set_inputs(packed)
    c.prove_resident(n)
    t0 = time.perf_counter()
    c.prove_resident(n)
    c.prove_resident(n)
    dt = (time.perf_counter() - t0) / 2
    st = c.prove_resident(n, stages=True)
    proofs, pubs, status = c.get_results(n)
    assert (status == 0).all()
    print(json.dumps({"proofs_per_s": n / dt, "stage_ms": {k: round(float(v), 2) for k, v in zip(NAMES, st)},
                      "digest": hashlib.sha256(proofs.tobytes()).hexdigest()[:16]}))


def main():
    if sys.argv[1:2] == ["--child"]:
        return child(int(sys.argv[2]))
    n = 512
    axes = []
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == "--n":
            n = int(args.pop(0))
        else:
            k, v = a.split("=")
            axes.append([(k, x) for x in v.split(",")])
    for combo in itertools.product(*axes):
        env = dict(os.environ)
        env.update(dict(combo))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n)], env=env, capture_output=True, text=True)
        last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-400:]
        print(" ".join(f"{k}={v}" for k, v in combo), last, flush=True)


if __name__ == "__main__":
    main()
