"""K2 parity (through the C ABI) against the oracle and the reference's golden outputs; tolerance 1e-5 relative
(north_star), plus an absolute floor of 1e-7 for near-zero components."""
import numpy as np
import pytest
import torch

from oracle import descriptors_np as D
from tests.util import golden

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-7
# Whitened descriptors (3xTF32 projection, tcgen05 or mma.sync): north_star's 1e-5 relative, plus an absolute floor of
# 5e-7 for the small components of a UNIT vector (5e-7 of the norm; bench.py reports the measured maximum, ~1e-7 at
# 2048 -> 2048: `roofline_k2.error_vs_float64`)
W_RTOL, W_ATOL = 1e-5, 5e-7


def _run(fmaps, p, **kw):
    from gandtr_b200 import _lib
    fm = [torch.from_numpy(np.ascontiguousarray(f)).cuda() for f in fmaps]
    pt = torch.tensor([p], dtype=torch.float32, device="cuda")
    P, m, split = kw.pop("P", None), kw.pop("m", None), kw.pop("split", False)
    P_split = None
    if P is not None:
        P = torch.tensor(P, dtype=torch.float32, device="cuda")
        m = torch.tensor(m, dtype=torch.float32, device="cuda").reshape(-1)
        if split:
            P_split = _lib.whiten_prepare(P[:kw.get("dim") or P.shape[0]].contiguous())
    out = _lib.gem_whiten(fm, pt, 1e-6, P=P, m=m, P_split=P_split, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_reference_golden():
    g = golden("descriptors.npz")
    ci = 0
    while "c%d_p" % ci in g.files:
        fm = []
        while "c%d_fmap%d" % (ci, len(fm)) in g.files:
            fm.append(g["c%d_fmap%d" % (ci, len(fm))])
        p = float(g["c%d_p" % ci])
        np.testing.assert_allclose(_run(fm[:1], p), g["c%d_plain" % ci], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(_run(fm, p, aggregate=True, msp_is_p=True), g["c%d_agg" % ci], rtol=RTOL, atol=ATOL)
        wh = _run(fm, p, aggregate=True, msp_is_p=True, P=g["c%d_P" % ci], m=g["c%d_m" % ci], dim=int(g["c%d_dim" % ci]))
        np.testing.assert_allclose(wh, g["c%d_whiten" % ci], rtol=2e-5, atol=2e-6)
        ci += 1
    assert ci == 3


@pytest.mark.parametrize("n,c,sizes,p", [(3, 512, [(48, 64)], 3.0), (2, 2048, [(24, 32), (17, 23), (12, 16)], 3.0),
                                         (2, 512, [(48, 64), (33, 45), (24, 32)], 2.92), (1, 96, [(5, 7)], 1.0),
                                         (4, 64, [(3, 3)], 2.0), (2, 40, [(1, 1)], 3.0)])
def test_oracle_shapes(n, c, sizes, p):
    rs = np.random.RandomState(n * 1000 + c)
    fm = [(np.abs(rs.normal(0, 1, (n, c, h, w))) * (rs.rand(n, c, 1, 1) > 0.1)).astype(np.float32) for h, w in sizes]
    multi = len(sizes) > 1
    P = (rs.normal(0, 1, (c, c)) / np.sqrt(c))
    m = 0.05 * rs.rand(c, 1)
    np.testing.assert_allclose(_run(fm[:1], p), D.descriptor_pipeline(fm[:1], p=p), rtol=RTOL, atol=ATOL)
    agg = _run(fm, p, aggregate=True, msp_is_p=multi)
    np.testing.assert_allclose(agg, D.descriptor_pipeline(fm, p=p, aggregate=True, msp_is_p=multi), rtol=RTOL, atol=ATOL)
    dim = c - 8
    wh = _run(fm, p, aggregate=True, msp_is_p=multi, P=P, m=m, dim=dim)
    ref = D.descriptor_pipeline(fm, p=p, aggregate=True, msp_is_p=multi, P=P, m=m, dimensions=dim)
    np.testing.assert_allclose(wh, ref, rtol=W_RTOL, atol=W_ATOL)
    np.testing.assert_allclose(np.linalg.norm(wh, axis=1), 1.0, atol=1e-5)


def test_unaligned_rows_and_offsets():
    """Feature-map rows that do not start on 16-byte boundaries (odd h*w) and a sliced storage offset."""
    rs = np.random.RandomState(5)
    base = torch.from_numpy(np.abs(rs.normal(0, 1, (1 + 2 * 40 * 7 * 9,))).astype(np.float32)).cuda()
    fm = base[1:].view(2, 40, 7, 9)
    from gandtr_b200 import _lib
    with pytest.raises(_lib.GdtError):
        _lib.gem_whiten([fm.permute(0, 1, 3, 2)], torch.tensor([3.0], device="cuda"))   # non-contiguous
    out = _lib.gem_whiten([fm], torch.tensor([3.0], device="cuda")).cpu().numpy()
    np.testing.assert_allclose(out, D.descriptor_pipeline([fm.cpu().numpy()], p=3.0), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("n,c,hw,dim", [(128, 2048, (24, 32), 2048), (130, 512, (6, 8), 512), (5, 2048, (3, 4), 1000),
                                        (64, 96, (4, 4), 96), (33, 512, (12, 16), 128), (1, 256, (2, 2), 256)])
def test_whitening_tcgen05_path(n, c, hw, dim):
    """The tcgen05 (kind::tf32, 3xTF32) projection selected by gdt_whiten_prepare against the float64 oracle and against
    the mma.sync kernel: full / partial 128-row image tiles, partial 128-column output tiles, short K."""
    rs = np.random.RandomState(n + c + dim)
    fm = [(np.abs(rs.normal(0, 1, (n, c) + hw)) * (rs.rand(n, c, 1, 1) > 0.1)).astype(np.float32)]
    P = rs.normal(0, 1, (c, c)) / np.sqrt(c)
    m = 0.05 * rs.rand(c, 1)
    ref = D.descriptor_pipeline(fm, p=3.0, aggregate=True, P=P, m=m, dimensions=dim)
    tc = _run(fm, 3.0, aggregate=True, P=P, m=m, dim=dim, split=True)
    simt = _run(fm, 3.0, aggregate=True, P=P, m=m, dim=dim)
    np.testing.assert_allclose(tc, ref, rtol=W_RTOL, atol=W_ATOL)
    np.testing.assert_allclose(tc, simt, rtol=W_RTOL, atol=W_ATOL)
    # the contract's 1e-5 read against the vector's norm (unit vectors: SURVEY App. C item 9): measured ~1e-7
    assert np.abs(tc - ref).max() < 1e-6 * np.linalg.norm(ref, axis=1).max()
    np.testing.assert_allclose(np.linalg.norm(tc, axis=1), 1.0, atol=1e-5)
