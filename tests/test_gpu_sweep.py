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
