"""GPU parity tests of the raw kernels (field arithmetic, NTT, MSM) against the CPU oracle.
Integer work: the bar is bit-exact equality."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def rand_field(rng, p, n):
    vals = [int.from_bytes(rng.bytes(40), "little") % p for _ in range(n)]
    return vals


def to_limbs(vals):
    return np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals), dtype=np.uint32).reshape(-1, 8).copy()


def from_limbs(a):
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]


@pytest.mark.parametrize("field,p", [("fq", Q), ("fr", R)])
def test_field_ops(field, p):
    from zk_franchise_proof_circuit_b200 import raw
    rng = np.random.default_rng(7)
    n = 4096
    a = rand_field(rng, p, n)
    b = rand_field(rng, p, n)
    # edge cases
    a[:6] = [0, 1, p - 1, p - 1, 0, p - 2]
    b[:6] = [0, p - 1, p - 1, 1, p - 1, p - 2]
    # carry patterns of the dedicated squaring (fp_sqr.inc): saturated limbs, single limbs, the sign boundary
    a[6:16] = [(1 << 224) - 1, (1 << 253) - 1, p // 2, (1 << 32) - 1, (1 << 64) - 1, ((1 << 128) - 1) << 64,
               0xFFFFFFFF << 192, (1 << 253) + (1 << 32) - 1, ((1 << 253) - 1) ^ (0xFFFFFFFF << 96), p - (1 << 200)]
    rinv = pow(1 << 256, -1, p)
    A, B = to_limbs(a), to_limbs(b)
    assert from_limbs(raw.field_op(field, "mul", A, B)) == [x * y * rinv % p for x, y in zip(a, b)]
    assert from_limbs(raw.field_op(field, "sqr", A, B)) == [x * x * rinv % p for x in a]
    assert from_limbs(raw.field_op(field, "add", A, B)) == [(x + y) % p for x, y in zip(a, b)]
    assert from_limbs(raw.field_op(field, "sub", A, B)) == [(x - y) % p for x, y in zip(a, b)]
    assert from_limbs(raw.field_op(field, "neg", A, B)) == [(-x) % p for x in a]
    assert from_limbs(raw.field_op(field, "to_mont", A, B)) == [x * (1 << 256) % p for x in a]
    assert from_limbs(raw.field_op(field, "from_mont", A, B)) == [x * rinv % p for x in a]
    # Montgomery inverse: inv(aR) = a^-1 R  <=>  out = (x)^-1 * R^2
    got = from_limbs(raw.field_op(field, "inv", A[:64], B[:64]))
    exp = [(pow(x, -1, p) * (1 << 512)) % p if x else 0 for x in a[:64]]
    assert got == exp


@pytest.mark.parametrize("logn", [11, 12, 13, 17, 19])
def test_ntt_matches_oracle(logn):
    from zk_franchise_proof_circuit_b200 import raw
    rng = np.random.default_rng(logn)
    n = 1 << logn
    vals = rand_field(rng, R, 64)
    v = np.frombuffer(b"".join(vals[i % 64].to_bytes(32, "little") for i in range(n)), dtype=np.uint8).reshape(n, 32).copy()
    # make it non-periodic
    v[:, 0] ^= (np.arange(n) & 0xFF).astype(np.uint8)
    v[:, 1] ^= ((np.arange(n) >> 8) & 0xFF).astype(np.uint8)
    fwd = raw.ntt(v)
    assert np.array_equal(fwd, O.ntt(v))
    inv = raw.ntt(fwd, inverse=True)
    assert np.array_equal(inv, v)            # round trip
    assert np.array_equal(raw.ntt(v, inverse=True), O.ntt(v, inverse=True))


def test_ntt_batched_and_coset():
    from zk_franchise_proof_circuit_b200 import raw
    logn, nvec = 13, 3
    n = 1 << logn
    rng = np.random.default_rng(99)
    v = rng.integers(0, 256, size=(nvec, n, 32), dtype=np.uint8)
    v[:, :, 31] &= 0x0F
    out = raw.ntt(v)
    for k in range(nvec):
        assert np.array_equal(out[k], O.ntt(v[k]))
    # coset chain: iNTT, multiply coefficient i by w_{2n}^i, NTT
    w2n = pow(5, (R - 1) >> (logn + 1), R)
    co = raw.coset_ntt(v)
    for k in range(nvec):
        coeffs = O.ntt(v[k], inverse=True)
        c = [int.from_bytes(coeffs[i].tobytes(), "little") for i in range(n)]
        acc = 1
        sh = []
        for i in range(n):
            sh.append(c[i] * acc % R)
            acc = acc * w2n % R
        shb = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in sh), dtype=np.uint8).reshape(n, 32)
        assert np.array_equal(co[k], O.ntt(shb))


def _g1_points(n, seed):
    """n distinct G1 points k_i * G (oracle fixed-base), canonical affine bytes."""
    rng = np.random.default_rng(seed)
    import ctypes
    out = np.zeros((n, 64), dtype=np.uint8)
    L = O.lib()
    for i in range(n):
        k = int.from_bytes(rng.bytes(31), "little") + 1
        buf = (ctypes.c_uint8 * 64)()
        L.orc_g1_mul_gen(O._buf(O.le32(k)), buf)
        out[i] = np.frombuffer(bytes(buf), dtype=np.uint8)
    return out


def _scalars(rng, shape_n, special=True):
    s = rng.integers(0, 256, size=shape_n + (32,), dtype=np.uint8)
    s[..., 31] &= 0x1F      # < 2^253 < r
    if special:
        flat = s.reshape(-1, 32)
        flat[0] = 0
        flat[1] = 0
        flat[1, 0] = 1
        flat[2] = np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint8)
        flat[3] = 0xFF
        flat[3, 31] = 0x0F
        flat[4] = 0
        flat[4, 1] = 0x80      # 2^15: the signed-digit boundary
        flat[5] = 0
        flat[5, 0] = 1
        flat[5, 1] = 0x80      # 2^15 + 1
    return s


@pytest.mark.parametrize("window", [16, 13, 12])
def test_msm_g1_small_matches_oracle(window, monkeypatch):
    from zk_franchise_proof_circuit_b200 import raw
    monkeypatch.setenv("ZKB_RAW_MSM_C", str(window))       # signed-digit window size (bits)
    n = 700
    bases = _g1_points(n, 1)
    bases[10] = 0                     # point at infinity
    bases[11] = bases[12]             # repeated base (exercises the doubling branch)
    rng = np.random.default_rng(5)
    sc = _scalars(rng, (3, n))
    sc[1, 11] = sc[1, 12]             # same digits on equal bases -> P + P inside one bucket
    out = raw.msm_g1(bases, sc)
    for b in range(3):
        assert out[b].tobytes() == O.msm_g1(bases, sc[b])


@pytest.mark.parametrize("levels,group", [(1, 8), (2, 8), (3, 16), (4, 8), (3, 512)])
def test_msm_g1_pair_tree_matches_oracle(levels, group, monkeypatch):
    """The batched-affine pair tree (the prover's H MSM in batch shape) on degenerate inputs: a base at infinity, equal
    bases with equal digits (P + P inside a bucket -> doubling), equal bases with opposite scalars (P - P -> the
    point at infinity flowing through later levels), tiny inversion groups so that thread ranges straddle buckets."""
    from zk_franchise_proof_circuit_b200 import raw
    monkeypatch.setenv("ZKB_RAW_MSM_C", "12")              # 2,048 buckets, 23 windows: lists of ~8 at n = 700
    n = 700
    bases = _g1_points(n, 1)
    bases[10] = 0
    bases[11] = bases[12]
    bases[20] = bases[21]
    bases[30] = bases[31] = bases[32] = bases[33]
    rng = np.random.default_rng(5)
    sc = _scalars(rng, (3, n))
    sc[1, 11] = sc[1, 12]                                  # doubling
    neg = (R - int.from_bytes(sc[1, 20].tobytes(), "little")) % R
    sc[1, 21] = np.frombuffer(neg.to_bytes(32, "little"), dtype=np.uint8)     # P - P in every window
    sc[2, 30] = sc[2, 31] = sc[2, 32] = sc[2, 33]          # four equal entries: doubling at two levels
    sc[2, 40:60] = 0
    out = raw.msm_g1(bases, sc, pair_tree=True, levels=levels, group=group)
    for b in range(3):
        assert out[b].tobytes() == O.msm_g1(bases, sc[b]), f"batch item {b}"


@pytest.mark.parametrize("window", [16, 12])
def test_msm_g1_variable_base_matches_oracle(window, monkeypatch):
    """No window table: one bucket set per window, window sums combined by Horner's rule.  Same degenerate inputs as the
    fixed-base test (infinity, repeated bases, signed-digit boundary scalars)."""
    from zk_franchise_proof_circuit_b200 import raw
    monkeypatch.setenv("ZKB_RAW_MSM_C", str(window))
    n = 700
    bases = _g1_points(n, 1)
    bases[10] = 0
    bases[11] = bases[12]
    rng = np.random.default_rng(5)
    sc = _scalars(rng, (3, n))
    sc[1, 11] = sc[1, 12]
    out = raw.msm_g1(bases, sc, variable_base=True)
    for b in range(3):
        assert out[b].tobytes() == O.msm_g1(bases, sc[b]), f"batch item {b}"


def test_msm_g1_pair_tree_dense_buckets(monkeypatch):
    """H-MSM shape scaled down: 2^14 random scalars, c = 12 -> ~180 entries per bucket, default tree (3 levels, 512)."""
    from zk_franchise_proof_circuit_b200 import raw
    monkeypatch.setenv("ZKB_RAW_MSM_C", "12")
    n = 1 << 14
    base = _g1_points(64, 9)
    bases = np.tile(base, (n // 64, 1))                    # 64 distinct points, each 256 times (plenty of doublings)
    rng = np.random.default_rng(6)
    sc = _scalars(rng, (2, n), special=False)
    out, kms, _ = raw.msm_g1(bases, sc, timing=True, pair_tree=True)
    ref, kms0, _ = raw.msm_g1(bases, sc, timing=True)
    print(f"pair tree {kms:.3f} ms, XYZZ {kms0:.3f} ms")
    assert np.array_equal(out, ref)
    assert out[0].tobytes() == O.msm_g1(bases, sc[0])


def test_msm_on_zkey_bases(art_dir):
    """A (G1) and B2 (G2) sections of the dev zkey with a random scalar vector, full size."""
    from zk_franchise_proof_circuit_b200 import raw
    zk = O.ZKeyRef(open(art_dir + "/proving_key.zkey", "rb").read())
    rng = np.random.default_rng(11)
    n = zk.n_vars
    sc = _scalars(rng, (n,))
    a = zk.points(5)
    out, kms, tms = raw.msm_g1(a, sc, timing=True)
    print(f"G1 MSM n={n}: kernels {kms:.3f} ms, table build {tms:.1f} ms")
    assert out[0].tobytes() == O.msm_g1(a, sc)
    b2 = zk.points(7)
    out2, kms2, tms2 = raw.msm_g2(b2, sc, timing=True)
    print(f"G2 MSM n={n}: kernels {kms2:.3f} ms, table build {tms2:.1f} ms")
    assert out2[0].tobytes() == O.msm_g2(b2, sc)


def test_modmul_throughput_reported():
    from zk_franchise_proof_circuit_b200 import raw
    for bps in (2, 4, 8):
        rate, ms = raw.bench_modmul("fq", 2048, bps)
        print(f"modmul microbench blocks/SM={bps}: {rate / 1e9:.1f} G modmul/s ({ms:.2f} ms)")
    assert rate > 1e9
