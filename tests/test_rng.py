"""CPU: host mirror of the device Philox streams."""
import numpy as np

from sdim_b200.rng import (CHANNEL_D, CHANNEL_F, CHANNEL_P, measurement_draws, noise_draws, philox4x32,
                           prob_to_thresh24)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [int(x) for x in philox4x32(0, 0, 0, 0, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert [int(x) for x in philox4x32(f, f, f, f, f, f)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert [int(x) for x in philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)] == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_thresholds():
    assert prob_to_thresh24(0.0) == 1 << 24 and prob_to_thresh24(1.0) == 0
    assert prob_to_thresh24(0.5) == 1 << 23


def test_measurement_draws_uniform_and_shard_invariant():
    d = 5
    a = measurement_draws(9, d, np.arange(20000), 3)
    assert a.shape == (20000, 3) and a.max() == d - 1
    hist = np.bincount(a[:, 1], minlength=d) / 20000
    assert np.abs(hist - 1 / d).max() < 0.01
    b = measurement_draws(9, d, np.arange(10000, 20000), 3)
    assert np.array_equal(a[10000:], b)                    # counters use the global shot id
    big = measurement_draws(9, d, np.array([2 ** 33 + 5], dtype=np.uint64), 3)
    assert not np.array_equal(big[0], a[5])


def test_noise_draw_distribution():
    """Distribution of sdim/program.py:486-507."""
    d, shots = 5, 200000
    th = [prob_to_thresh24(p) for p in (0.3, 0.3, 0.3, 0.0, 1.0)]
    ch = [CHANNEL_D, CHANNEL_F, CHANNEL_P, CHANNEL_D, CHANNEL_D]
    ab = noise_draws(3, d, np.arange(shots), th, ch)
    fired = (ab.sum(-1) > 0).mean(0)
    assert np.abs(fired[:3] - 0.3).max() < 0.005 and fired[3] == 0 and fired[4] == 1.0
    dd = ab[:, 4].astype(int)
    r = dd[:, 0] + d * dd[:, 1]
    assert r.min() == 1 and r.max() == d * d - 1
    assert np.abs(np.bincount(r, minlength=d * d)[1:] / shots - 1 / (d * d - 1)).max() < 0.003
    f = ab[:, 1]
    assert (f[:, 1] == 0).all() and set(np.unique(f[:, 0])) == set(range(d))
    pz = ab[:, 2]
    assert (pz[:, 0] == 0).all() and set(np.unique(pz[:, 1])) == set(range(d))
