"""Pins oracle/descriptors_np.py and oracle/whiten_np.py against outputs of the unmodified reference modules."""
import numpy as np

from oracle import descriptors_np as D
from oracle import whiten_np as WN
from tests.util import golden

RTOL = 1e-5  # north_star: descriptors within 1e-5 relative of the reference fp32 path


def _cases(g):
    ci = 0
    while "c%d_p" % ci in g.files:
        fm = []
        while "c%d_fmap%d" % (ci, len(fm)) in g.files:
            fm.append(g["c%d_fmap%d" % (ci, len(fm))])
        yield ci, fm, float(g["c%d_p" % ci]), g["c%d_P" % ci], g["c%d_m" % ci], int(g["c%d_dim" % ci])
        ci += 1


def test_gem_l2n_multiscale_whiten_match_reference():
    g = golden("descriptors.npz")
    seen = 0
    for ci, fm, p, P, m, dim in _cases(g):
        plain = D.descriptor_pipeline(fm[:1], p=p)
        np.testing.assert_allclose(plain, g["c%d_plain" % ci], rtol=RTOL, atol=1e-7)
        agg = D.descriptor_pipeline(fm, p=p, aggregate=True, msp_is_p=True)
        np.testing.assert_allclose(agg, g["c%d_agg" % ci], rtol=RTOL, atol=1e-7)
        wh = D.descriptor_pipeline(fm, p=p, aggregate=True, msp_is_p=True, P=P, m=m, dimensions=dim)
        np.testing.assert_allclose(wh, g["c%d_whiten" % ci], rtol=2e-5, atol=2e-6)
        assert wh.shape[1] == dim
        seen += 1
    assert seen == 3


def test_whitenlearn_matches_reference():
    g = golden("whiten.npz")
    m, P = WN.whitenlearn(g["X"], list(g["qidxs"]), list(g["pidxs"]))
    np.testing.assert_allclose(m, g["m"], rtol=1e-12, atol=1e-14)
    # eigenvectors are defined up to sign: compare rows up to sign
    sign = np.sign((P * g["P"]).sum(axis=1, keepdims=True))
    np.testing.assert_allclose(P * sign, g["P"], rtol=1e-6, atol=1e-8)
    Y = WN.whitenapply(g["X"][:, :50], m, P, dimensions=16)
    np.testing.assert_allclose(np.abs(Y), np.abs(g["Y"]), rtol=1e-6, atol=1e-8)
