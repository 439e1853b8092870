import numpy as np

from oracle import philox as P


def test_random123_known_answers():
    out = P.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    out = P.philox4x32(f, f, f, f, f, f)
    assert [int(x) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = P.philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_ranges():
    idx = np.arange(100000)
    u = P.uniform_f64(123, P.STREAM_BELIEF, 3, 7, idx)
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    v = P.uniform_f32(123, P.STREAM_ACTION, 3, 7, idx)
    assert v.dtype == np.float32 and v.min() >= 0 and v.max() < 1 and abs(v.mean() - 0.5) < 0.01


def test_belief_stream_shares_one_block_between_four_slots():
    rows = np.arange(6).reshape(2, 3) + 40
    for K in (1, 2, 5, 9):
        u = P.belief_uniforms(9, 2, 11, rows, K)
        assert u.shape == (2, 3, K) and u.min() > 0 and u.max() < 1
        kq = (K + 3) // 4
        for jj in range(K):
            x = P.draw(9, P.STREAM_BELIEF, 2, 11, rows * kq + jj // 4)
            want = (x[jj % 4].astype(np.float64) + 0.5) * 2.0 ** -32
            assert np.array_equal(u[..., jj], want)
    assert len(np.unique(P.belief_uniforms(9, 2, 11, np.arange(1000), 7))) == 7000
    # the centred 32-bit value is exact in float64 and survives the float32 screen of the many-agent kernel within 2^-23
    w = np.array([0, 1, 2 ** 31, 2 ** 32 - 1], dtype=np.uint64)
    u = (w.astype(np.float64) + 0.5) * 2.0 ** -32
    assert np.array_equal(u * 2.0 ** 33, 2.0 * w.astype(np.float64) + 1.0)
