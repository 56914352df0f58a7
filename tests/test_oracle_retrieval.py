"""Pins oracle/retrieval_np.py against the unmodified reference's compute_map / compute_map_and_print outputs."""
import numpy as np

from oracle import retrieval_np as R
from tests.util import golden


def _gnd(g):
    out = []
    for e, h, j in zip(g["easy"], g["hard"], g["junk"]):
        out.append({"easy": e[e >= 0], "hard": h[h >= 0], "junk": j[j >= 0]})
    return out


def test_map_protocols_match_reference_on_reference_ranks():
    g = golden("map_eval.npz")
    gnd = _gnd(g)
    avg, per, mpr = R.compute_map_protocols("roxford5k", g["ranks"], gnd)
    for key in ("easy", "medium", "hard"):
        assert avg["map_" + key] == float(g["map_" + key])            # float64 bit-equal
        np.testing.assert_array_equal(per["ap_" + key], g["ap_" + key])
    np.testing.assert_array_equal(mpr["mpr_medium"], g["mprM"])
    gnd_old = [{"ok": np.concatenate([x["easy"], x["hard"]]), "junk": x["junk"]} for x in gnd]
    gnd_old[3]["ok"] = np.array([], dtype=np.int64)
    avg_o, per_o, _ = R.compute_map_protocols("tokyo", g["ranks"], gnd_old)
    assert avg_o["map"] == float(g["old_map"])
    np.testing.assert_array_equal(per_o["ap"], g["old_ap"])
    assert np.isnan(per_o["ap"][3])


def test_exact_ranking_agrees_with_reference_ranking_modulo_near_ties():
    g = golden("map_eval.npz")
    s = R.scores_exact(g["q"], g["db"])                   # [nq, ndb]
    ref_s = R.scores_reference(g["q"], g["db"])           # [ndb, nq] fp32 sgemm
    assert np.abs(s - ref_s.T).max() < 5e-7
    mine = R.full_ranks(s)
    ref = g["ranks"]
    diff = np.argwhere(mine != ref)
    for pos, qi in diff:                                   # any disagreement must be a near-tie in the reference
        assert abs(float(ref_s[mine[pos, qi], qi]) - float(ref_s[ref[pos, qi], qi])) < 2e-7
    avg, _, _ = R.compute_map_protocols("roxford5k", mine, _gnd(g))
    assert abs(avg["map_medium"] - float(g["map_medium"])) < 1e-6


def test_topk_padding_and_tie_rule():
    s = np.array([[0.5, 0.7, 0.7, -0.0, 0.0]], dtype=np.float32)
    ts, ti = R.topk(s, 7)
    assert ti[0].tolist() == [1, 2, 0, 3, 4, -1, -1]
    assert np.isneginf(ts[0, 5:]).all()
