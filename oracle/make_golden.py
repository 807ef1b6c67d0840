#!/usr/bin/env python3
"""Regenerates tests/golden/ from the mounted reference (TEST INFRASTRUCTURE)."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_witness as RW  # noqa: E402

REF = "/root/reference/artifacts/zkCensus/dev/160"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

for f in ("inputs_example.json", "proof.json", "signals.json", "verification_key.json"):
    shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
inp = json.load(open(os.path.join(REF, "inputs_example.json")))
code, w = RW.witness(inp)
assert code == 0
val = lambda i: str(int.from_bytes(w[i].tobytes(), "little"))
kat = {"sha256": hashlib.sha256(w.tobytes()).hexdigest(), "n_wires": int(w.shape[0]),
       "wires": {str(i): val(i) for i in list(range(0, 13)) + [333, 334, 335, 336, 1000, 41138, 82753]}}
for bad, name in (({"voteWeight": "11"}, "voteWeight_gt_available"), ({"nullifier": "5"}, "wrong_nullifier")):
    b = dict(inp)
    b.update(bad)
    kat[f"exception_code_{name}"] = RW.witness(b)[0]
json.dump(kat, open(os.path.join(OUT, "witness_kat.json"), "w"), indent=1)
print(kat["sha256"])
