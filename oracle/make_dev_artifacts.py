#!/usr/bin/env python3
"""Generate the dev artifacts directory the product and the tests read (TEST INFRASTRUCTURE).

Mirrors the reference's on-disk contract `artifacts/<name>/<env>/<nLevels>/` (`zk_census_test.go:78`)
under <repo>/artifacts/zkCensus/dev/160/ (git-ignored):

  circuit.wasm            copied from the reference (the product's load_circuit consumes it as-is)
  inputs_example.json     copied from the reference fixture
  circuit.r1cs            RECONSTRUCTED: census_model.py R1CS reduced onto the wasm's wires (iden3 format)
  proving_key.zkey        SYNTHESISED dev key: liboracle orc_setup with known toxic waste (seed below)
  verification_key.json   vkey matching that dev key (snarkjs layout)

The reference's real circuit.r1cs / proving_key.zkey are missing blobs (`.MISSING_LARGE_BLOBS:1-3`)
and its real ceremony (`circuit/circuit-compiler.sh:52-136`) needs circom + snarkjs, which are not in
this image.  The product parses the zkey generically; a production key loads unchanged.
"""
import json
import os
import shutil
import struct
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import wasm_tools as W          # noqa: E402
import census_model as M        # noqa: E402

REF = os.environ.get("ZKB_REFERENCE", "/root/reference")
REF_DIR = os.path.join(REF, "artifacts/zkCensus/dev/160")
OUT_DIR = os.path.join(os.path.dirname(HERE), "artifacts/zkCensus/dev/160")
SETUP_SEED = 0x5EED0160
N_PUBLIC = 8


def write_r1cs(path, rows, n_wires, n_pub_in, n_prv_in):
    """iden3 r1cs binary: header(1), constraints(2), wire2label(3)."""
    P = M.P
    hdr = struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack(
        "<IIIIQI", n_wires, 0, n_pub_in, n_prv_in, n_wires, len(rows))
    cons = bytearray()
    for A, B, C in rows:
        for X in (A, B, C):
            cons += struct.pack("<I", len(X))
            for w in sorted(X):
                cons += struct.pack("<I", w) + X[w].to_bytes(32, "little")
    w2l = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    with open(path, "wb") as f:
        f.write(b"r1cs" + struct.pack("<II", 1, 3))
        for sid, body in ((1, hdr), (2, bytes(cons)), (3, w2l)):
            f.write(struct.pack("<IQ", sid, len(body)))
            f.write(body)


def main(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    wasm_src = os.path.join(REF_DIR, "circuit.wasm")
    wasm_dst = os.path.join(OUT_DIR, "circuit.wasm")
    if os.path.exists(wasm_src):
        shutil.copyfile(wasm_src, wasm_dst)
        shutil.copyfile(os.path.join(REF_DIR, "inputs_example.json"), os.path.join(OUT_DIR, "inputs_example.json"))
    if not os.path.exists(wasm_dst):
        raise SystemExit("no circuit.wasm available (reference not mounted and no previous copy)")
    r1cs = os.path.join(OUT_DIR, "circuit.r1cs")
    zkey = os.path.join(OUT_DIR, "proving_key.zkey")
    vkey = os.path.join(OUT_DIR, "verification_key.json")
    if not force and all(os.path.exists(p) for p in (r1cs, zkey, vkey)):
        print("dev artifacts up to date:", OUT_DIR)
        return
    t0 = time.time()
    mod = W.Module(open(wasm_dst, "rb").read())
    tables = W.poseidon_tables(mod)
    wmap, _ = W.witness_map(mod)
    inputs = json.load(open(os.path.join(OUT_DIR, "inputs_example.json")))
    c = M.build(tables, inputs, emit=True)
    rows, n_elim = M.reduce_r1cs(c, wmap, verbose=True)
    n_wires = len(wmap)
    n_inputs = W.getter_const(mod, "getInputSize")
    write_r1cs(r1cs, rows, n_wires, N_PUBLIC, n_inputs - N_PUBLIC)
    print(f"r1cs: {len(rows)} constraints over {n_wires} wires ({n_elim} signals eliminated) "
          f"in {time.time() - t0:.1f}s")
    import oracle_lib as O
    t0 = time.time()
    O.setup(r1cs, zkey, vkey, SETUP_SEED)
    print(f"dev setup (toxic seed {SETUP_SEED:#x}): {os.path.getsize(zkey) / 2**20:.1f} MiB zkey "
          f"in {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main(force="--force" in sys.argv)
