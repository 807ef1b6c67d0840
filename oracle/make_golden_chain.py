#!/usr/bin/env python3
"""Regenerates tests/golden/chain40_kat.json and chain600_kat.json (TEST INFRASTRUCTURE): the oracle's proof of the
synthetic Poseidon-shaped chain circuit (BASELINE.json configs[3] shape, setup seed 7) with r = 1234567, s = 7654321
at 40 links (17,523 wires) and 600 links (262,803 wires, domain 2^19).  They pin the oracle's circuit generator + dev
setup + prover against accidental change; the GPU tests compare the product's proof of the same key and witness with
them, and `bench.py --gpus N` (N > 1) compares the proof made through zkb_load_circuit_shard across all ranks with the
600-link vector.

    python oracle/make_golden_chain.py            # rewrite both vectors
    python oracle/make_golden_chain.py --artifacts # (build step) only make sure artifacts/chain600/ exists and matches
"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import numpy as np          # noqa: E402
import oracle_lib as O      # noqa: E402

R, S, SEED = 1234567, 7654321, 7


def make(links, d):
    n_wires, n_cons, domain = O.chain_artifacts(links, SEED, d, check=links <= 40)
    zkey = open(os.path.join(d, "proving_key.zkey"), "rb").read()
    wtns = open(os.path.join(d, "witness.wtns"), "rb").read()
    w = np.frombuffer(wtns[-n_wires * 32:], dtype=np.uint8).reshape(n_wires, 32)
    proof = O.ZKeyRef(zkey).prove(w, R, S)
    return {"links": links, "setup_seed": SEED, "r": R, "s": S, "n_wires": n_wires, "n_constraints": n_cons,
            "domain": domain, "zkey_sha256": hashlib.sha256(zkey).hexdigest(),
            "wtns_sha256": hashlib.sha256(wtns).hexdigest(),
            "public": [str(int.from_bytes(w[1].tobytes(), "little"))], "proof": O.proof_json(proof)}


def ensure_chain600_artifacts():
    """artifacts/chain600/{proving_key.zkey, witness.wtns, verification_key.json} (git-ignored; travels to the GPU box):
    regenerated unless the files already match the committed vector."""
    d = os.path.join(ROOT, "artifacts", "chain600")
    kat_path = os.path.join(ROOT, "tests", "golden", "chain600_kat.json")
    kat = json.load(open(kat_path)) if os.path.exists(kat_path) else None
    zp, wp = os.path.join(d, "proving_key.zkey"), os.path.join(d, "witness.wtns")
    if kat and os.path.exists(zp) and os.path.exists(wp) and \
            hashlib.sha256(open(zp, "rb").read()).hexdigest() == kat["zkey_sha256"] and \
            hashlib.sha256(open(wp, "rb").read()).hexdigest() == kat["wtns_sha256"]:
        print("chain600 artifacts up to date:", d)
        return
    O.lib().orc_set_threads(len(os.sched_getaffinity(0)))
    O.chain_artifacts(600, SEED, d, check=False)
    if kat:
        assert hashlib.sha256(open(zp, "rb").read()).hexdigest() == kat["zkey_sha256"], "generator changed: chain600 key"
    print("chain600 artifacts written:", d)


if __name__ == "__main__":
    if "--artifacts" in sys.argv:
        ensure_chain600_artifacts()
        sys.exit(0)
    for links in (40, 600):
        d = os.path.join(ROOT, "artifacts", "chain600") if links == 600 else tempfile.mkdtemp(prefix=f"chain{links}_")
        kat = make(links, d)
        out = os.path.join(ROOT, "tests", "golden", f"chain{links}_kat.json")
        json.dump(kat, open(out, "w"), indent=1)
        print(out, kat["zkey_sha256"][:16], kat["wtns_sha256"][:16])
