#!/usr/bin/env python3
"""tests/golden/ops_program_kat.json (TEST INFRASTRUCTURE): known answers of the OpsTest program (make_ops_wasm.py).

The expected witnesses come from a model of the circom operators written here over Python integers, from the language
documentation (circom 2 "Basic operators": field, relational on the signed representative val(z) = z - p for
z > p / 2, shifts that reverse for k > p / 2, bit operations masked to 254 bits and reduced, integer division and
remainder on the representatives in [0, p)) - independent of the wasm and of the product.  tests/test_witness_program.py
checks that the wasm itself, run concretely, gives the same witnesses; the product's extracted program must too.

    python oracle/make_golden_ops.py          # rewrites the JSON
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import make_ops_wasm as OW      # noqa: E402

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
HALF = P // 2
MASK = (1 << 254) - 1


def val(z):
    return z - P if z > HALF else z


def shr(x, k):
    if k <= HALF:
        return x >> k if k < 254 else 0
    return shl(x, P - k)


def shl(x, k):
    if k <= HALF:
        return ((x << k) & MASK) % P if k < 254 else 0
    return shr(x, P - k)


def model(x):
    """(status, witness) of OpsTest on inputs x[0..2] in [0, p)"""
    x0, x1, x2 = x
    if x1 == 0:
        return 4, None                       # integer division by zero: the wasm traps
    b = lambda t: 1 if t else 0
    inv = lambda z: pow(z, P - 2, P)
    o = [0] * OW.N_OUT
    o[0], o[1], o[2], o[3] = x0 // x1, x0 % x1, pow(x0, x2, P), x0 * inv(x1) % P
    o[4], o[5], o[6] = ((x0 | x1) & MASK) % P, ((x0 ^ x1) & MASK) % P, (~x0 & MASK) % P
    o[7], o[8] = shl(x0, x2), shr(x0, x2)
    o[9], o[10], o[11], o[12] = b(val(x0) < val(x1)), b(val(x0) > val(x1)), b(val(x0) <= val(x1)), b(val(x0) >= val(x1))
    o[13], o[14] = b(x0 == x1), b(x0 != x1)
    o[15], o[16], o[17] = b(x0 != 0 and x2 != 0), b(x0 != 0 or x2 != 0), b(x2 == 0)
    o[18], o[19], o[20] = (P - x0) % P, ((x0 & x1) & MASK) % P, inv(x1)
    o[21], o[22] = (x0 * x1 + x2) % P, (x0 - x1) % P
    o[23] = (x0 * x2 % P if x2 != 0 else x1) if val(x0) < val(x1) else (x0 - x1) % P
    o[24] = x1 if val(x0) > val(x1) else x0
    o[25] = x2 * x2 % P
    status = 0 if val(x2) <= val(x1) else 4
    return status, [1] + o + [x0, x1, x2]


def cases():
    rnd = random.Random(0x0B5)
    out = [(7, 3, 2), (5, 9, 0), (9, 5, 1), (4, 4, 4), (1 << 200, 12345, 300), (0, 1, 0), (P - 1, 1, 1), (1, HALF, 253),
           (HALF + 1, HALF, 254), (P - 1, HALF, 255), (MASK % P, HALF - 1, 31), (12345, 77, P - 3), (12345, 77, P - 254)]
    for _ in range(12):
        out.append((rnd.randrange(P), rnd.randrange(1, HALF), rnd.choice([0, 1, 2, 17, 64, 200, 253, rnd.randrange(1 << 60)])))
    for _ in range(4):
        out.append((rnd.randrange(P), rnd.randrange(HALF + 1, P), rnd.randrange(P)))      # x1 negative
    out += [(3, 0, 0), (P - 1, P - 2, 253), (1, 2, 3)]       # zero divisor; assert fails (253 > -2); assert fails (3 > 2)
    return out


def main():
    wasm = OW.build(open(os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160", "circuit.wasm"), "rb").read())
    kat = {"what": "OpsTest program (oracle/make_ops_wasm.py): inputs x[3], status, witness (30 values) by the Python "
                   "model of the circom operators in oracle/make_golden_ops.py",
           "wasm_sha256": hashlib.sha256(wasm).hexdigest(), "cases": []}
    for x in cases():
        st, w = model(x)
        kat["cases"].append({"x": [str(v) for v in x], "status": st, "witness": [str(v) for v in w] if w and st == 0 else None})
    path = os.path.join(ROOT, "tests", "golden", "ops_program_kat.json")
    with open(path, "w") as f:
        json.dump(kat, f, indent=0)
        f.write("\n")
    print(path, len(kat["cases"]), "cases,", sum(c["status"] == 0 for c in kat["cases"]), "valid")


if __name__ == "__main__":
    main()
