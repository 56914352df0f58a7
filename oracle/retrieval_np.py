"""ORACLE (test infrastructure, never imported by gandtr_b200/): NumPy restatement of scoring, ranking and mAP.

Follows
  * mdir/components/optim/score/cirscore.py:71-72      scores = vecs.T @ qvecs ; ranks = argsort(-scores, axis=0)
  * mdir/external/cirtorch/utils/evaluate.py:3-37      compute_ap
  * mdir/external/cirtorch/utils/evaluate.py:39-111    compute_map
  * mdir/external/cirtorch/utils/evaluate.py:114-152   compute_map_and_print (old / revisited protocol split)
Parity pin: tests/test_oracle_retrieval.py checks compute_map* against the unmodified reference functions on
seeded synthetic ground truth (fixtures tests/golden/map_*.npz from tools/gen_golden.py).

Ranking definition used by the product (and by `topk` here): score descending, ties -> lower index first, with
the score of a (query, row) pair defined as the float64-accumulated dot product rounded once to float32. The
reference's np.argsort on a float32 sgemm result is unstable on ties and differs from this by ~1e-7 in score; the
tests compare modulo those near-ties.
"""
import numpy as np


def scores_exact(q, db):
    """q: [nq, d], db: [ndb, d] float32 -> [nq, ndb] float32 (float64 accumulation, one rounding)."""
    return (np.asarray(q, dtype=np.float64) @ np.asarray(db, dtype=np.float64).T).astype(np.float32)


def scores_reference(q, db):
    """The reference's own arithmetic: float32 sgemm, db-major layout (cirscore.py:71) -> [ndb, nq]."""
    return np.dot(np.asarray(db, dtype=np.float32), np.asarray(q, dtype=np.float32).T)


def topk(scores, k, index_base=0):
    """scores: [nq, ndb] -> (top scores [nq, k], idx [nq, k] int64); (-inf, -1) padding when ndb < k."""
    nq, ndb = scores.shape
    s = scores + np.float32(0.0)
    order = np.lexsort((np.broadcast_to(np.arange(ndb), s.shape), -s), axis=1)[:, :k]
    out_s = np.full((nq, k), -np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    kk = min(k, ndb)
    out_s[:, :kk] = np.take_along_axis(s, order, axis=1)[:, :kk]
    out_i[:, :kk] = order[:, :kk] + index_base
    return out_s, out_i


def full_ranks(scores):
    """[nq, ndb] -> ranks [ndb, nq] int64 under the product's total order (the layout compute_map expects)."""
    nq, ndb = scores.shape
    s = scores + np.float32(0.0)
    return np.lexsort((np.broadcast_to(np.arange(ndb), s.shape), -s), axis=1).T.astype(np.int64)


def compute_ap(ranks, nres):
    """Trapezoidal AP from 0-based positive positions (evaluate.py:3-37), same operation order."""
    ap = 0.0
    recall_step = 1.0 / nres
    for j, rank in enumerate(ranks):
        rank = int(rank)
        precision_0 = 1.0 if rank == 0 else float(j) / rank
        precision_1 = float(j + 1) / (rank + 1)
        ap += (precision_0 + precision_1) * recall_step / 2.0
    return ap


def compute_map(ranks, gnd, kappas=()):
    """evaluate.py:39-111. ranks: [ndb, nq]; gnd: list of {'ok': ids, 'junk': ids}."""
    nq = len(gnd)
    aps = np.zeros(nq)
    prs = np.zeros((nq, len(kappas)))
    pr = np.zeros(len(kappas))
    total, nempty = 0.0, 0
    for i in range(nq):
        ok = np.asarray(gnd[i]["ok"])
        if ok.shape[0] == 0:
            aps[i] = np.nan
            prs[i, :] = np.nan
            nempty += 1
            continue
        junk_ids = np.asarray(gnd[i].get("junk", []))
        col = ranks[:, i]
        pos = np.flatnonzero(np.isin(col, ok))
        junk = np.flatnonzero(np.isin(col, junk_ids))
        if len(junk):
            # each positive moves up by the number of junk entries ranked before it (evaluate.py:83-94)
            pos = pos - np.searchsorted(junk, pos, side="left")
        ap = compute_ap(pos, len(ok))
        total += ap
        aps[i] = ap
        pos1 = pos + 1
        for j, kap in enumerate(kappas):
            kq = min(int(pos1.max()), kap)
            prs[i, j] = (pos1 <= kq).sum() / kq
        pr = pr + prs[i, :]
    return total / (nq - nempty), aps, pr / (nq - nempty), prs


def compute_map_protocols(dataset, ranks, gnd, kappas=(1, 5, 10)):
    """evaluate.py:114-152 without the printing: returns (averages, per-query scores, mean P@k per protocol)."""
    if "ok" in gnd[0]:
        m, aps, _, _ = compute_map(ranks, gnd)
        return {"map": m}, {"ap": aps}, {}
    if not (dataset.startswith("roxford5k") or dataset.startswith("rparis6k")):
        return None
    out_avg, out_aps, out_pr = {}, {}, {}
    for name, ok_keys, junk_keys in (("easy", ("easy",), ("junk", "hard")),
                                     ("medium", ("easy", "hard"), ("junk",)),
                                     ("hard", ("hard",), ("junk", "easy"))):
        g = [{"ok": np.concatenate([np.asarray(x[k_], dtype=np.int64) for k_ in ok_keys]),
              "junk": np.concatenate([np.asarray(x[k_], dtype=np.int64) for k_ in junk_keys])} for x in gnd]
        m, aps, mpr, _ = compute_map(ranks, g, kappas)
        out_avg["map_" + name] = m
        out_aps["ap_" + name] = aps
        out_pr["mpr_" + name] = mpr
    return out_avg, out_aps, out_pr
