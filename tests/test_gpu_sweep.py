"""GPU parity tests of the device-resident raw sessions (include/zkcensus_b200.h, zkb_msm_session_* / zkb_ntt_bench):
a G1 MSM split by point range into sub-MSMs and across ranks, partial sums pushed into rank 0's exchange buffer and
combined there.  Bit-exact against the CPU oracle at small sizes; at large sizes the split-invariance
property (any number of ranks / sub-MSMs gives the same point) is the check."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _run_group(logn, nranks, devices, window=16, seed=1):
    from zk_franchise_proof_circuit_b200 import raw
    ss = [raw.MsmSession(logn, r, nranks, device=devices[r % len(devices)], seed=seed, window=window) for r in range(nranks)]
    for s in ss[1:]:
        s.attach_local(ss[0])
    for s in ss:
        s.run()
    point, ms = ss[0].combine(nranks)
    return ss, point, ms


@pytest.mark.parametrize("window", [16, 13])
def test_session_single_rank_matches_oracle(window):
    ss, point, _ = _run_group(12, 1, [0], window)
    bases, scalars = ss[0].read()
    assert point == O.msm_g1(bases, scalars)
    # a second step (next epoch) gives the same point
    ss[0].run()
    assert ss[0].combine(1)[0] == point


def test_session_four_ranks_one_device_matches_oracle():
    ss, point, _ = _run_group(13, 4, [0])
    parts = [s.read() for s in ss]
    bases = np.concatenate([p[0] for p in parts])
    scalars = np.concatenate([p[1] for p in parts])
    assert point == O.msm_g1(bases, scalars)
    # the same global index space: one rank over all 2^13 points produces the same bases and the same sum
    one, point1, _ = _run_group(13, 1, [0])
    b1, s1 = one[0].read()
    assert np.array_equal(b1, bases) and np.array_equal(s1, scalars)
    assert point1 == point


def test_session_variable_base_matches_fixed_base():
    """The variable-base form of the session (no window table) gives the same point as the fixed-base one, on one rank
    and split over four; and the oracle's at 2^12."""
    from zk_franchise_proof_circuit_b200 import raw
    def run(logn, nranks, vb):
        ss = [raw.MsmSession(logn, r, nranks, device=0, seed=1, window=16, variable_base=vb) for r in range(nranks)]
        for s in ss[1:]:
            s.attach_local(ss[0])
        for s in ss:
            s.run()
        point, _ = ss[0].combine(nranks)
        return ss, point
    ss, p_vb = run(12, 1, True)
    bases, scalars = ss[0].read()
    assert p_vb == O.msm_g1(bases, scalars)
    _, p_fixed = run(20, 1, False)
    _, p_vb1 = run(20, 1, True)
    _, p_vb4 = run(20, 4, True)
    assert p_fixed == p_vb1 == p_vb4 and p_fixed != bytes(64)


def test_session_split_invariance_large():
    """2^20 points: 8 sub-MSMs of 2^17 on one rank == 2 ranks x 4 sub-MSMs == 16 ranks of 2^16 (no oracle at this size)."""
    _, p1, _ = _run_group(20, 1, [0])
    _, p2, _ = _run_group(20, 2, [0])
    _, p16, _ = _run_group(20, 16, [0])
    assert p1 == p2 == p16
    assert p1 != bytes(64)


def test_session_two_gpus_peer_push():
    """Rank 1 on a second GPU writes its partial sum into GPU 0's exchange buffer over NVLink (peer mapping)."""
    from zk_franchise_proof_circuit_b200 import _native
    if _native.lib().zkb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _, p_two, _ = _run_group(18, 2, [0, 1])
    _, p_one, _ = _run_group(18, 1, [0])
    assert p_two == p_one


def test_ntt_bench_runs():
    from zk_franchise_proof_circuit_b200 import raw
    dif, dit = raw.ntt_bench(17, nvec=8, iters=2)
    assert dif > 0 and dit > 0


def _bitrev_perm(logn):
    n = 1 << logn
    idx = np.arange(n, dtype=np.uint32)
    rev = np.zeros(n, dtype=np.uint32)
    for b in range(logn):
        rev |= ((idx >> b) & 1) << (logn - 1 - b)
    return rev


def test_dif_forward_is_bitreversed_ntt():
    """ties the parity aid used below to the oracle-checked natural-order transform"""
    from zk_franchise_proof_circuit_b200 import raw
    logn = 13
    rng = np.random.default_rng(3)
    v = rng.integers(0, 256, size=(1 << logn, 32), dtype=np.uint8)
    v[:, 31] &= 0x0F
    nat = O.ntt(v)
    assert np.array_equal(raw.ntt_dif_forward(v), nat[_bitrev_perm(logn)])


def _ntt_dist_group(logn, nranks, devices):
    from zk_franchise_proof_circuit_b200 import raw
    ss = [raw.NttDist(logn, r, nranks, device=devices[r % len(devices)]) for r in range(nranks)]
    for a in ss:
        for b in ss:
            if a is not b:
                a.attach_local(b.rank, b)
    for s in ss:
        s.fill()
    n2 = 1 << (logn // 2)
    n1 = 1 << (logn - logn // 2)
    cols = np.concatenate([s.read(0) for s in ss]).reshape(n2, n1, 32)       # B[n2][n1]
    x = np.ascontiguousarray(cols.transpose(1, 0, 2)).reshape(n1 * n2, 32)     # x[n1 * N2 + n2]
    for s in ss:
        s.run()
    times = [s.sync() for s in ss]
    out = np.concatenate([s.read(1) for s in ss])
    for s in ss:
        s.close()
    return x, out, times


@pytest.mark.parametrize("logn,nranks", [(24, 1), (24, 4), (22, 2), (23, 4)])
def test_ntt_dist_matches_single_gpu_plan(logn, nranks):
    """2^22 (the smallest: 2^11 x 2^11, one tile per column), 2^23 (odd split) and 2^24 values, ranks as streams of one
    GPU: concatenated output rows == the single-GPU DIF transform."""
    from zk_franchise_proof_circuit_b200 import raw
    x, out, _ = _ntt_dist_group(logn, nranks, [0])
    assert np.array_equal(out, raw.ntt_dif_forward(x))


def test_ntt_dist_two_gpus():
    from zk_franchise_proof_circuit_b200 import raw, _native
    if _native.lib().zkb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, out, times = _ntt_dist_group(24, 2, [0, 1])
    assert np.array_equal(out, raw.ntt_dif_forward(x))
    print("4-step NTT 2^24 on 2 GPUs, ms (total, columns, exchange, rows):", times)


def test_session_argument_errors():
    """Bad shapes are errors with a message, not crashes or silent fallbacks."""
    from zk_franchise_proof_circuit_b200 import raw
    from zk_franchise_proof_circuit_b200._native import NativeError
    with pytest.raises(NativeError):
        raw.MsmSession(30)                      # larger than the entry index allows
    with pytest.raises(NativeError):
        raw.MsmSession(16, rank=0, nranks=3)    # ranks must be a power of two
    with pytest.raises(NativeError):
        raw.NttDist(20)                         # the 4-step split needs 2^24 or more
    s = raw.MsmSession(12)
    s.run()
    with pytest.raises(NativeError):
        s.combine(0)
    assert s.combine(1)[0] != bytes(64)
