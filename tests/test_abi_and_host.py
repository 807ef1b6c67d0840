"""CPU tests: the C-ABI library loads and exports every symbol include/zkcensus_b200.h declares; host-side
logic (multi-process sharding over gloo, JSON mirror) - no compute calls without a GPU."""
import ctypes
import json
import os
import re
import subprocess
import sys
import pytest

import helpers as H

HEADER = os.path.join(H.ROOT, "include", "zkcensus_b200.h")
LIB = os.path.join(H.ROOT, "zk_franchise_proof_circuit_b200", "libzkcensus_b200.so")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+|groth16_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    L = ctypes.CDLL(LIB)
    syms = _declared_symbols()
    assert len(syms) >= 20 and "groth16_prover" in syms
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/zkcensus_b200.h but not exported"


def test_no_gpu_means_error_not_fallback():
    """Without a device every compute entry point fails loudly (there is no CPU path in the product)."""
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    L = ctypes.CDLL(LIB)
    if L.zkb_device_count() > 0:
        pytest.skip("a GPU is visible")
    L.zkb_last_error.restype = ctypes.c_char_p
    ctx = ctypes.c_void_p()
    assert L.zkb_ctx_create(0, ctypes.byref(ctx)) == 1
    assert b"no CPU fallback" in L.zkb_last_error()
    buf = (ctypes.c_uint8 * 64)()
    assert L.zkb_raw_field_op(0, 0, buf, buf, buf, ctypes.c_size_t(2)) == 1
    # the raw sessions and the verifier are compute entry points too
    h = ctypes.c_void_p()
    assert L.zkb_msm_session_create(0, 16, 0, 1, ctypes.c_uint64(1), 16, ctypes.byref(h)) == 1
    assert L.zkb_ntt_dist_create(0, 24, 0, 1, ctypes.c_uint64(1), ctypes.byref(h)) == 1
    f = ctypes.c_float()
    assert L.zkb_ntt_bench(0, 16, 1, 1, ctypes.byref(f), ctypes.byref(f)) == 1
    vk = open(H.GOLDEN + "/verification_key.json", "rb").read()
    pf = open(H.GOLDEN + "/proof.json", "rb").read()
    pub = open(H.GOLDEN + "/signals.json", "rb").read()
    L.zkb_verify.argtypes = [ctypes.c_char_p, ctypes.c_size_t] * 3
    assert L.zkb_verify(vk, len(vk), pub, len(pub), pf, len(pf)) == 1     # error, not "valid" and not a CPU pairing


def test_product_does_not_import_oracle():
    """Nothing under the package (or bench's product arm) may reference oracle/."""
    pkg = os.path.join(H.ROOT, "zk_franchise_proof_circuit_b200")
    for dp, _, fs in os.walk(pkg):
        if "build" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "ref_witness" not in txt and "census_model" not in txt, f
                assert not re.search(r"#include\s+\"[^\"]*oracle/", txt), f


def test_go_shaped_proof_roundtrip():
    from zk_franchise_proof_circuit_b200 import prover
    pdata = open(H.GOLDEN + "/proof.json", "rb").read()
    psig = open(H.GOLDEN + "/signals.json", "rb").read()
    p = prover.parse_proof(pdata, psig)
    out_p, out_s = p.bytes()
    assert out_p == pdata.strip() and out_s == psig.strip()        # byte-identical re-marshal
    with pytest.raises(ValueError):
        prover.parse_proof(b'{"pi_a":[]}', psig)


def test_pack_inputs_order():
    from zk_franchise_proof_circuit_b200 import prover
    inp = H.fixture_inputs()
    p = prover.pack_inputs(inp)
    assert p.shape == (334, 32)
    val = lambda i: int.from_bytes(p[i].tobytes(), "little")
    # main-signal order: wasm hashmap positions, signal k at row k-1
    assert val(0) == int(inp["electionId"][0]) and val(2) == int(inp["nullifier"])
    assert val(7) == int(inp["voteWeight"]) and val(8) == int(inp["availableWeight"])
    assert val(9) == int(inp["address"]) and val(12) == int(inp["censusSiblings"][0])
    assert val(173) == int(inp["sikSiblings"][0])


def test_shard_plan_two_ranks_gloo(tmp_path):
    """bench.py's sharding (contiguous proof ranges per rank, max-over-ranks timing) on 2 CPU ranks."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, json\n"
        f"sys.path.insert(0, {H.ROOT!r})\n"
        "import torch, torch.distributed as dist\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "lo, hi = bench.shard_range(1000, r, w)\n"
        "t = bench.max_over_ranks(float(10 + r), 'cpu')\n"
        "tot = bench.sum_over_ranks(float(hi - lo), 'cpu')\n"
        f"open(os.path.join({str(tmp_path)!r}, 'rank%d.json' % r), 'w').write("
        "json.dumps({'rank': r, 'lo': lo, 'hi': hi, 'tmax': t, 'tot': tot}))\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]   # per-rank files: stdout can interleave
    assert [(r["lo"], r["hi"]) for r in rows] == [(0, 500), (500, 1000)]
    assert all(r["tmax"] == 11.0 and r["tot"] == 1000.0 for r in rows)


def test_census_tree_builder_matches_oracle_generator(art_dir):
    """census_tree.SparseMerkleTree (batched, level by level) == the oracle's recursive arbo restatement."""
    import census_gen as G
    from zk_franchise_proof_circuit_b200 import census_tree as CT
    P = G.Poseidon(H.poseidon_tables())
    hasher = lambda rows: [P(list(r)) for r in rows]
    import random
    rnd = random.Random(5)
    leaves = {rnd.getrandbits(160): rnd.getrandbits(200) for _ in range(37)}
    a, b = CT.SparseMerkleTree(hasher, leaves), G.SMT(P, leaves)
    assert a.root == b.root
    for k in list(leaves)[:10]:
        assert a.siblings(k) == b.siblings(k)
    assert CT.SparseMerkleTree(hasher, {}).root == 0
    one = CT.SparseMerkleTree(hasher, {5: 7})
    assert one.root == P([5, 7, 1]) and one.siblings(5) == []
    assert CT.bytes_to_arbo(b"x") == G.bytes_to_arbo(b"x")


def test_host_build_of_pairing_verifies_reference_fixture(tmp_path):
    """csrc/pairing.cuh compiled for the host (portable path): the reference's proof.json verifies under its
    verification_key.json; tampering is rejected (the same code runs per thread in zkb_verify on the GPU)."""
    import oracle_lib as O
    src = os.path.join(H.ROOT, "tests", "host_emul", "pairing_host.cc")
    so = str(tmp_path / "libpairing_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    L = ctypes.CDLL(so)
    vk = json.load(open(H.GOLDEN + "/verification_key.json"))
    pf = json.load(open(H.GOLDEN + "/proof.json"))
    pub = json.load(open(H.GOLDEN + "/signals.json"))
    vkb, ic, n = O.vkey_bin(vk)
    call = lambda pub_, pf_: L.host_groth16_verify(O._buf(vkb), O._buf(ic), n, O._buf(O.pub_bin(pub_)), O._buf(O.proof_bin(pf_)))
    assert call(pub, pf) == 1
    bad = list(pub)
    bad[0] = "1"
    assert call(bad, pf) == 0
    pf2 = json.loads(json.dumps(pf))
    pf2["pi_c"] = pf["pi_a"]
    assert call(pub, pf2) == 0


def test_export_vkey_matches_setup_vkeys(art_dir):
    """zkb_export_vkey == `snarkjs zkey export verificationkey` (circuit/circuit-compiler.sh:128-134): byte-identical to
    the verification keys the dev setup wrote for the census key and the chain key (whose vk_alphabeta_12 computation is
    pinned against the REFERENCE's verification_key.json by test_vk_alphabeta_12_reproduced).  Host arithmetic: runs
    without a GPU."""
    from zk_franchise_proof_circuit_b200 import prover
    for d in (art_dir, os.path.join(H.ROOT, "artifacts", "chain600")):
        if not os.path.exists(os.path.join(d, "proving_key.zkey")):
            continue
        got = prover.export_vkey(open(os.path.join(d, "proving_key.zkey"), "rb").read())
        assert got == open(os.path.join(d, "verification_key.json"), "rb").read()
    from zk_franchise_proof_circuit_b200._native import NativeError
    with pytest.raises(NativeError):
        prover.export_vkey(b"zkey" + bytes(100))


def test_json_formatter_under_address_sanitizer(tmp_path):
    """u256_to_dec overflowed its stack buffer for values >= 10^72.  csrc/json_io.cc built with
    -fsanitize=address,undefined formats r-1, q-1, 2^256-1, 0, 1 and a proof / public-signal document."""
    src = os.path.join(H.ROOT, "tests", "host_emul", "json_asan.cc")
    exe = str(tmp_path / "json_asan")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                           "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "json_asan ok" in out.stdout, out.stdout + out.stderr


def test_census_of_chosen_depth_passes_the_reference_wasm(art_dir):
    """census_tree.gen_census_depth (bench.py's depth sweep) with the oracle's Poseidon as the hasher: tree_depth is
    what was asked for and the reference wasm accepts the inputs (depth 1, the reference's own 10-leaf shape 4, and
    the circuit's maximum 160)."""
    import census_gen as G
    import ref_witness as RW
    from zk_franchise_proof_circuit_b200 import census_tree as CT
    if not RW.available():
        pytest.skip("oracle/_ref not built")

    class Hasher:
        P = G.Poseidon(H.poseidon_tables())

        def poseidon(self, rows):
            return [self.P(list(r)) for r in rows]
    for depth in (1, 4, 160):
        v = CT.gen_census_depth(Hasher(), 1, depth, seed=3)[0]
        assert CT.tree_depth(v) == depth
        assert RW.witness(v)[0] == 0


def test_reference_arm_under_torchrun_two_ranks(art_dir):
    """`bench.py --impl reference --gpus 2` launched like the driver launches it: rank 0 alone times the CPU
    implementation and prints ONE JSON line with the contract's keys, rank 1 exits 0 without work."""
    import ref_witness as RW
    if not RW.available():
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29741", ZKB_REF_SAMPLE="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29741", os.path.join(H.ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["unit"] == "proofs/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("census.circom") and "tree_depth" in d["config"]


def test_zkey_parser_survives_damaged_keys(art_dir):
    """truncations and bit flips in the section table / header of a proving key are rejected (or accepted) without
    reading outside the buffer: zkb_export_vkey and groth16_public_size_for_zkey_buf parse on the host, no GPU needed"""
    import random
    L = ctypes.CDLL(LIB)
    L.groth16_public_size_for_zkey_buf.argtypes = [ctypes.c_void_p, ctypes.c_ulong, ctypes.POINTER(ctypes.c_ulong),
                                                  ctypes.c_void_p, ctypes.c_ulong]
    L.zkb_export_vkey.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
    zkey = open(os.path.join(H.ROOT, "artifacts", "smtVerifier", "dev", "160", "proving_key.zkey"), "rb").read() \
        if os.path.exists(os.path.join(H.ROOT, "artifacts", "smtVerifier", "dev", "160", "proving_key.zkey")) \
        else open(art_dir + "/proving_key.zkey", "rb").read()
    rnd = random.Random(11)
    err = ctypes.create_string_buffer(256)
    n = ctypes.c_ulong(0)
    ok = bad = 0
    for trial in range(150):
        z = bytearray(zkey)
        if trial % 3 == 0:
            z = z[:rnd.randrange(0, min(len(z), 4096))]                    # truncated inside the header sections
        elif trial % 3 == 1:
            z = z[:rnd.randrange(len(z) // 2, len(z))]                      # truncated inside a point section
        else:
            for _ in range(4):
                z[rnd.randrange(0, 2048)] ^= 1 << rnd.randrange(8)         # section ids, lengths, nVars, domain size
        buf = (ctypes.c_char * max(len(z), 1)).from_buffer_copy(bytes(z) or b"\0")
        rc = L.groth16_public_size_for_zkey_buf(ctypes.addressof(buf), len(z), ctypes.byref(n), err, 256)
        assert rc in (0, 1)
        out = ctypes.create_string_buffer(1 << 16)
        m = ctypes.c_size_t(1 << 16)
        rc2 = L.zkb_export_vkey(ctypes.addressof(buf), len(z), out, ctypes.byref(m))
        assert rc2 in (0, 1, 2)
        ok += rc == 0
        bad += rc == 1
    assert bad > 50 and ok + bad == 150


def test_generated_squaring_is_current_and_exact():
    """csrc/fp_sqr.inc (the 36-limb-product squaring the device code includes) is what tools/gen_sqr.py emits, and the
    instruction list it was generated from squares every carry pattern exactly (emulated instruction by instruction;
    a carry that would be dropped, or one that would leave an asm block, raises)."""
    import importlib.util
    import random
    spec = importlib.util.spec_from_file_location("gen_sqr", os.path.join(H.ROOT, "tools", "gen_sqr.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    blocks = G.program()
    inc = open(os.path.join(H.ROOT, "zk_franchise_proof_circuit_b200", "csrc", "fp_sqr.inc")).read()
    assert inc == G.emit(blocks)
    assert sum(1 for b in blocks for i in b if ".lo" in i[0]) == 36        # limb products (mul.lo / mad.lo / madc.lo)
    rnd = random.Random(5)
    M = (1 << 256) - 1
    vals = [0, 1, M, 1 << 255, H.census_gen.P - 1, 0xFFFFFFFF, 0xFFFFFFFF << 224]
    vals += [M ^ (0xFFFFFFFF << (32 * i)) for i in range(8)] + [rnd.getrandbits(256) for _ in range(500)]
    for v in vals:
        assert G.emulate(blocks, v) == v * v


def test_ntt_pass_model_radix4_equals_radix2_and_the_dft():
    """tools/ntt_pass_model.py restates the index / twiddle logic of csrc/ntt.cu's passes (k_ntt_pass, k_ntt_pass4) over
    a small prime: the radix-4 pass == the radix-2 pass for every pass shape of 2^11..2^13 (DIF and DIT), and the DIF
    passes of a 2^11 vector are the DFT in bit-reversed order."""
    import importlib.util
    import random
    spec = importlib.util.spec_from_file_location("ntt_pass_model", os.path.join(H.ROOT, "tools", "ntt_pass_model.py"))
    M = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(M)
    M.main((11, 12, 13))
    logn, n, P = 11, 1 << 11, M.P
    w = pow(M.GEN, (P - 1) // n, P)
    tw = [pow(w, e, P) for e in range(n // 2)]
    rnd = random.Random(2)
    x = [rnd.randrange(P) for _ in range(n)]
    got = M.run(logn, 0, 11, True, True, x, tw)
    pw = [pow(w, e, P) for e in range(n)]
    for k in rnd.sample(range(n), 64):
        want = sum(x[j] * pw[(j * k) % n] for j in range(n)) % P
        assert got[int(format(k, "011b")[::-1], 2)] == want


def test_host_parsers_on_damaged_documents_under_address_sanitizer(tmp_path):
    """tests/host_emul/parsers_asan.cc: the JSON parsers, the census-wasm recogniser and the .zkey reader built with
    -fsanitize=address on truncated / spliced / bit-flipped documents (held in exact-size heap buffers): they accept or
    reject, never read outside the buffer.  (The generic extractor has its own run in test_witness_program.py.)"""
    import random
    exe = str(tmp_path / "parsers_asan")
    csrc = os.path.join(H.ROOT, "zk_franchise_proof_circuit_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address", "-std=c++17", "-o", exe,
                           os.path.join(H.ROOT, "tests", "host_emul", "parsers_asan.cc")] +
                          [os.path.join(csrc, n) for n in ("json_io.cc", "wasm_circuit.cc", "zkey.cc")])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    rnd = random.Random(9)
    m = str(tmp_path / "m.bin")

    def run(kind, data):
        open(m, "wb").write(bytes(data))
        r = subprocess.run([exe, kind, m], capture_output=True, text=True, env=env, timeout=120)
        assert r.returncode == 0, (kind, r.stderr[:2000])
        return int(r.stdout.strip())
    docs = [open(os.path.join(H.GOLDEN, n), "rb").read() for n in ("inputs_example.json", "proof.json", "verification_key.json", "signals.json")]
    assert run("json", docs[0]) & 1 and run("json", docs[3]) & 4
    specials = [b"", b"{", b"[", b"\"", b"{\"a\":", b"{\"a\":[", b"{\"a\":\"", b"[\"1\",", b"[" * 5000, b"{\"a\":" + b"9" * 5000 + b"}",
                b"{\"a\":\"" + b"9" * 5000 + b"\"}", b"[-1]", b"{\"a\":1e5}", b"{\"a\":[1,2,],}", b"\xff\xfe{}"]
    for d in specials:
        run("json", d)
    for t in range(120):
        d = bytearray(rnd.choice(docs))
        k = t % 4
        if k == 0:
            d = d[:rnd.randrange(0, len(d))]
        elif k == 1:
            for _ in range(rnd.randrange(1, 6)):
                d[rnd.randrange(len(d))] = rnd.choice(b"{}[]\",:0123456789 \n\\-eE.x\x00\xff")
        elif k == 2:
            i = rnd.randrange(len(d))
            d[i:i] = bytes(rnd.choice(b"{}[]\",:") for _ in range(rnd.randrange(1, 30)))
        else:
            i = rnd.randrange(len(d))
            del d[i:min(len(d), i + rnd.randrange(1, 200))]
        run("json", d)
    art = os.path.join(H.ROOT, "artifacts", "zkCensus", "dev", "160")
    if not os.path.exists(art + "/circuit.wasm"):
        return
    wasm = open(art + "/circuit.wasm", "rb").read()
    assert run("wasm", wasm) == 1
    for t in range(24):
        d = bytearray(wasm)
        if t % 3 == 0:
            d = d[:rnd.randrange(8, len(d))]
        elif t % 3 == 1:
            for _ in range(rnd.randrange(1, 6)):
                d[rnd.randrange(0, 4096)] ^= 1 << rnd.randrange(8)
        else:
            for _ in range(rnd.randrange(1, 12)):
                d[rnd.randrange(len(d) - 600000, len(d))] = rnd.randrange(256)
        run("wasm", d)
    zpath = os.path.join(H.ROOT, "artifacts", "opsTest", "dev", "1", "proving_key.zkey")
    if not os.path.exists(zpath):
        return
    zkey = open(zpath, "rb").read()
    assert run("zkey", zkey) == 1
    offs, p = [], 12
    while p + 12 <= len(zkey):
        offs.append(p)
        p += 12 + int.from_bytes(zkey[p + 4:p + 12], "little")
    for t in range(80):
        z = bytearray(zkey)
        if t % 4 == 0:
            z = z[:rnd.randrange(0, len(z))]
        elif t % 4 == 1:
            for _ in range(rnd.randrange(1, 5)):
                z[rnd.randrange(0, 600)] ^= 1 << rnd.randrange(8)
        elif t % 4 == 2:
            for _ in range(rnd.randrange(1, 5)):
                z[rnd.randrange(0, len(z))] = rnd.randrange(256)
        else:
            o = rnd.choice(offs)                    # a section id / length / first count field
            for _ in range(rnd.randrange(1, 3)):
                z[o + rnd.randrange(0, 24)] = rnd.randrange(256)
        run("zkey", z)
