#!/usr/bin/env python3
"""Damaged witness calculators through the generic extractor (csrc/wasm_symexec.cc) under AddressSanitizer.

`zkb_load_circuit` hands the caller's .wasm to the extractor: truncations, bit flips in section headers, random bytes in
the code, data and name sections must end in a message (ZKB_UNSUPPORTED_CIRCUIT in the product) or in a valid program,
never in a memory error.  The harness (tests/host_emul/tape_host.cc) is built with -fsanitize=address and run in child
processes, 12 mutants each.  Round 2 found and fixed: type indices of functions / imports not checked against the type
section, call targets checked after the callee's type was looked up, runtime functions called with fewer arguments than
their name implies, element-segment offsets, unbounded vector counts, size arithmetic that could wrap.

    python tools/fuzz_wasm_extractor.py [n_seeds] [first_seed]          # needs artifacts/opsTest (build())
"""
import ctypes
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WASM = os.path.join(ROOT, "artifacts", "opsTest", "dev", "1", "circuit.wasm")


def build_harness(out_dir):
    so = os.path.join(out_dir, "libtape_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-o", so, os.path.join(ROOT, "tests", "host_emul", "tape_host.cc")])
    return so


def sections(wasm):
    p, secs = 8, []
    while p < len(wasm):
        sid, q, size, sh = wasm[p], p + 1, 0, 0
        while True:
            b = wasm[q]
            q += 1
            size |= (b & 0x7F) << sh
            sh += 7
            if not b & 0x80:
                break
        secs.append((sid, p, q, size))
        p = q + size
    return secs


def mutant(wasm, secs, rnd):
    z = bytearray(wasm)
    kind = rnd.randrange(5)
    if kind == 0:
        return bytes(z[:rnd.randrange(8, len(z))])
    if kind == 1:                                   # section id / size / first bytes (counts, type indices)
        sid, p0, q0, size = rnd.choice(secs)
        for _ in range(rnd.randrange(1, 4)):
            z[rnd.randrange(p0, min(q0 + 64, len(z)))] ^= 1 << rnd.randrange(8)
    elif kind == 2:                                 # code
        sid, p0, q0, size = rnd.choice([s for s in secs if s[0] == 10])
        for _ in range(rnd.randrange(1, 6)):
            z[rnd.randrange(q0, q0 + size)] = rnd.randrange(256)
    elif kind == 3:                                 # data / name
        sid, p0, q0, size = rnd.choice([s for s in secs if s[0] in (11, 0)])
        for _ in range(rnd.randrange(1, 6)):
            z[rnd.randrange(q0, min(q0 + 8192, q0 + size))] = rnd.randrange(256)
    else:                                           # type / import / function / element sections, anywhere
        sid, p0, q0, size = rnd.choice([s for s in secs if s[0] in (1, 2, 3, 9)])
        for _ in range(rnd.randrange(1, 4)):
            z[rnd.randrange(q0, q0 + size)] = rnd.randrange(256)
    return bytes(z)


def child(so, seed, n_mutants=12):
    wasm = open(WASM, "rb").read()
    L = ctypes.CDLL(so)
    L.tape_host_build.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    rnd = random.Random(seed)
    secs = sections(wasm)
    ok = bad = 0
    for _ in range(n_mutants):
        z = mutant(wasm, secs, rnd)
        buf = (ctypes.c_char * len(z)).from_buffer_copy(z)
        info = (ctypes.c_uint32 * 8)()
        err = ctypes.create_string_buffer(512)
        rc = L.tape_host_build(ctypes.addressof(buf), len(z), info, err, 512)
        assert rc == 0 or err.value, "rejected without a message"
        ok += rc == 0
        bad += rc != 0
    print("seed", seed, "ok", ok, "rejected", bad)


def run(n_seeds, first_seed=0, n_mutants=12, timeout=900):
    """returns (clean runs, failures as (seed, stderr tail))"""
    gcc_file = lambda n: subprocess.check_output(["gcc", "-print-file-name=" + n]).decode().strip()
    with tempfile.TemporaryDirectory() as d:
        so = build_harness(d)
        # python is not linked against libasan / libstdc++: both are preloaded so that the interceptors resolve
        env = dict(os.environ, LD_PRELOAD=gcc_file("libasan.so") + " " + gcc_file("libstdc++.so.6"), ASAN_OPTIONS="detect_leaks=0")
        clean, failures = 0, []
        for seed in range(first_seed, first_seed + n_seeds):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", so, str(seed), str(n_mutants)], env=env,
                                   capture_output=True, text=True, timeout=timeout)
            except subprocess.TimeoutExpired:
                failures.append((seed, "timeout"))
                continue
            if r.returncode == 0:
                clean += 1
                print(r.stdout.strip().splitlines()[-1], flush=True)
            else:
                failures.append((seed, r.stderr[-3000:]))
                print("seed", seed, "FAILED", flush=True)
    return clean, failures


if __name__ == "__main__":
    if sys.argv[1:2] == ["--child"]:
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
        first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        clean, failures = run(n, first)
        for seed, tail in failures:
            print("---- seed", seed, "----\n" + tail)
        print(clean, "clean,", len(failures), "failed")
        sys.exit(1 if failures else 0)
