#!/usr/bin/env python3
"""Regenerates tests/golden/chain40_kat.json (TEST INFRASTRUCTURE): the oracle's proof of the synthetic Poseidon-shaped
chain circuit (BASELINE.json configs[3] shape, 40 links, setup seed 7) with r = 1234567, s = 7654321.  It pins the
oracle's circuit generator + dev setup + prover against accidental change; the GPU test compares the product's proof
of the same key and witness with it."""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np          # noqa: E402
import oracle_lib as O      # noqa: E402

d = tempfile.mkdtemp(prefix="chain40_")
n_wires, n_cons, domain = O.chain_artifacts(40, 7, d)
zkey = open(os.path.join(d, "proving_key.zkey"), "rb").read()
wtns = open(os.path.join(d, "witness.wtns"), "rb").read()
w = np.frombuffer(wtns[-n_wires * 32:], dtype=np.uint8).reshape(n_wires, 32)
proof = O.ZKeyRef(zkey).prove(w, 1234567, 7654321)
kat = {"links": 40, "setup_seed": 7, "r": 1234567, "s": 7654321, "n_wires": n_wires, "n_constraints": n_cons,
       "domain": domain, "zkey_sha256": hashlib.sha256(zkey).hexdigest(), "wtns_sha256": hashlib.sha256(wtns).hexdigest(),
       "public": [str(int.from_bytes(w[1].tobytes(), "little"))], "proof": O.proof_json(proof)}
out = os.path.join(os.path.dirname(HERE), "tests", "golden", "chain40_kat.json")
json.dump(kat, open(out, "w"), indent=1)
print(out, kat["zkey_sha256"][:16], kat["wtns_sha256"][:16])
