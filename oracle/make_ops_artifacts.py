#!/usr/bin/env python3
"""artifacts/opsTest/dev/1/ (TEST INFRASTRUCTURE, git-ignored; travels to the GPU box): a third circuit in the
reference's on-disk layout (`artifacts/<name>/<env>/<nLevels>/`, zk_census_test.go:78):

  circuit.wasm            the OpsTest program of make_ops_wasm.py (every runtime operation, nested conditionals, an assert)
  circuit.r1cs            30 wires, 26 public outputs, 3 private inputs; one real constraint (x[2] * x[2] = o[25]) and empty
                          rows up to 3,000 (domain 2^12).  The other outputs are `<--` hints, unconstrained as in circom.
  proving_key.zkey        dev setup of that system (known toxic waste)
  verification_key.json   its verification key

What it is for: a SMALL non-census circuit through `zkb_load_circuit(zkey, wasm)` -> generic witness path -> proof."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import make_dev_artifacts as MDA    # noqa: E402
import make_ops_wasm as OW          # noqa: E402

OUT = os.path.join(ROOT, "artifacts", "opsTest", "dev", "1")
SEED = 0x5EED0095
N_ROWS = 3000


def main(force=False):
    os.makedirs(OUT, exist_ok=True)
    paths = [os.path.join(OUT, n) for n in ("circuit.wasm", "circuit.r1cs", "proving_key.zkey", "verification_key.json")]
    src = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160", "circuit.wasm")
    if not force and all(os.path.exists(p) for p in paths) and os.path.getmtime(paths[0]) >= os.path.getmtime(src) \
            and os.path.getmtime(paths[0]) >= os.path.getmtime(os.path.join(HERE, "make_ops_wasm.py")):
        print("opsTest artifacts up to date:", OUT)
        return
    open(paths[0], "wb").write(OW.build(open(src, "rb").read()))
    x2, o25 = 1 + OW.N_OUT + 2, 1 + 25
    rows = [({x2: 1}, {x2: 1}, {o25: 1})]
    rows += [({}, {}, {})] * (N_ROWS - len(rows))
    MDA.write_r1cs(paths[1], rows, OW.N_WIRES, OW.N_OUT, OW.N_IN)
    import oracle_lib as O
    O.setup(paths[1], paths[2], paths[3], SEED)
    print("opsTest artifacts written:", OUT, os.path.getsize(paths[2]) >> 10, "KiB zkey")


if __name__ == "__main__":
    main(force="--force" in sys.argv)
