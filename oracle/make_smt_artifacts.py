#!/usr/bin/env python3
"""artifacts/smtVerifier/dev/160/ (TEST INFRASTRUCTURE, git-ignored; travels to the GPU box): a second circuit in the
reference's on-disk layout (`artifacts/<name>/<env>/<nLevels>/`, zk_census_test.go:78):

  circuit.wasm            the SMTVerifier(160) program of make_smt_wasm.py
  circuit.r1cs            a stand-in constraint system over its 82,754 wires: `enabled` and `fnc` boolean, 5,000 rows in
                          all (the rest empty) so that the domain is 2^13 - the circuit's real constraints are not needed
                          to exercise the proving path, and no circom compiler exists here to emit them
  proving_key.zkey        dev setup of that system (known toxic waste), 2 public signals (enabled, root)
  verification_key.json   its verification key

What it is for: `zkb_load_circuit(zkey, wasm)` with a wasm that is NOT census.circom -> generic witness path -> proof."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import make_dev_artifacts as MDA    # noqa: E402
import make_smt_wasm as SW          # noqa: E402

OUT = os.path.join(ROOT, "artifacts", "smtVerifier", "dev", "160")
SEED = 0x5EED0161
N_WIRES, N_PUBLIC, N_ROWS = 82754, 2, 5000


def main(force=False):
    os.makedirs(OUT, exist_ok=True)
    paths = [os.path.join(OUT, n) for n in ("circuit.wasm", "circuit.r1cs", "proving_key.zkey", "verification_key.json")]
    src = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160", "circuit.wasm")
    if not force and all(os.path.exists(p) for p in paths) and os.path.getmtime(paths[0]) >= os.path.getmtime(src):
        print("smtVerifier artifacts up to date:", OUT)
        return
    open(paths[0], "wb").write(SW.patch(open(src, "rb").read()))
    one = {0: 1}
    rows = [({1: 1}, {1: 1}, {1: 1}),            # enabled * enabled = enabled
            ({169: 1}, {169: 1}, {169: 1}),      # fnc * fnc = fnc
            ({1: 1}, one, {1: 1})]               # enabled * 1 = enabled
    rows += [({}, {}, {})] * (N_ROWS - len(rows))
    MDA.write_r1cs(paths[1], rows, N_WIRES, N_PUBLIC, SW.N_IN - N_PUBLIC)
    import oracle_lib as O
    O.lib().orc_set_threads(len(os.sched_getaffinity(0)))
    O.setup(paths[1], paths[2], paths[3], SEED)
    print("smtVerifier artifacts written:", OUT, os.path.getsize(paths[2]) >> 20, "MiB zkey")


if __name__ == "__main__":
    main(force="--force" in sys.argv)
