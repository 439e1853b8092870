"""Oracle pin: oracle/belief.py against vectors recorded from the real BeliefFilter class."""
import numpy as np
import pytest

from oracle import belief as B


@pytest.mark.parametrize("tag", ["rand5", "rand3", "rand5x5", "org_known", "hvt_known"])
def test_belief_vectors(golden, tag):
    g = golden("belief_vectors.npz")
    fa = g[f"{tag}/filterAction"]
    n_bad_ap = 0
    for t in range(g[f"{tag}/obs"].shape[0]):
        ap, bprime, pred = B.belief_update(fa, g[f"{tag}/obs"][t], g[f"{tag}/prev"][t], g[f"{tag}/u"][t])
        assert np.array_equal(bprime, g[f"{tag}/bprime"][t]), (tag, t)
        np.testing.assert_allclose(pred, g[f"{tag}/prediction"][t], rtol=2e-15, atol=0)
        n_bad_ap += int((ap != g[f"{tag}/ap"][t]).sum())
    assert n_bad_ap == 0


def test_prior_and_likelihood(golden):
    g = golden("belief_vectors.npz")
    assert np.array_equal(B.uniform_prior(256, 5), g["rand5/prior0"])
    assert np.array_equal(B.uniform_prior(64, 3), g["rand3/prior0"])
    lik = B.likelihood_from_action(g["rand5/other_action"][0], 3)
    assert np.array_equal(lik, g["rand5/obs"][0])


def test_hundredths_lossless(golden):
    g = golden("belief_vectors.npz")
    b = g["rand5/bprime"]
    assert np.array_equal(B.from_hundredths(B.to_hundredths(b)), b)


def test_ia2c_tape_beliefs(golden):
    g = golden("ia2c_E64.npz")
    for f, actor in (("bf0", "act2"), ("bf1", "act1")):
        fa = g[f"{f}/filterAction"][0]
        lik = B.likelihood_from_action(g[f"{actor}/sampled"], 3)
        assert np.array_equal(lik, g[f"{f}/obs"])
        for t in range(lik.shape[0]):
            ap, bprime, _ = B.belief_update(fa, lik[t], g[f"{f}/prev"][t], g[f"{f}/u"][t])
            assert np.array_equal(ap, g[f"{f}/ap"][t])
            assert np.array_equal(bprime, g[f"{f}/bprime"][t])
