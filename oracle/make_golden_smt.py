#!/usr/bin/env python3
"""Regenerates tests/golden/smt_verifier_kat.json (TEST INFRASTRUCTURE): witnesses of the SMTVerifier program
(oracle/make_smt_wasm.py: the reference wasm with SMTVerifier(160) as main) computed by running that wasm NATIVELY
(oracle/_ref/smt, `make -C oracle ref_smt`) on four Merkle proofs: the reference fixture's census path, two synthetic
voters and the depth-160 voter."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H            # noqa: E402
import make_smt_wasm as SW      # noqa: E402
import ref_witness as RW        # noqa: E402

ref = RW.RefWasm()
assert ref.available(), "run `make -C oracle ref_smt` first"
wasm = SW.patch(open(os.path.join(H.ART, "circuit.wasm"), "rb").read())
voters = [H.fixture_inputs()] + list(H.voters(2)) + [H.deep_voters()[0]]
shas = []
for v in voters:
    code, w = ref.witness(SW.smt_inputs(v))
    assert code == 0
    shas.append(hashlib.sha256(w.tobytes()).hexdigest())
kat = {"what": "SMTVerifier(160) as main component of the reference wasm (oracle/make_smt_wasm.py); witnesses by the wasm "
               "transpiled to C and run natively", "wasm_sha256": hashlib.sha256(wasm).hexdigest(),
       "cases": ["inputs_example.json census path", "synthetic voter 0", "synthetic voter 1", "depth-160 voter"],
       "witness_sha256": shas}
json.dump(kat, open(os.path.join(ROOT, "tests", "golden", "smt_verifier_kat.json"), "w"), indent=1)
print(kat)
