"""K3 parity (through the C ABI): exact CUDA-core path and tcgen05 path against the oracle ranking
(score desc, index asc; score = float64-accumulated dot rounded to float32)."""
import numpy as np
import pytest
import torch

from oracle import retrieval_np as R
from tests.util import golden, unit_rows

pytestmark = pytest.mark.gpu


def _check_lists(s, i, os_, oi, q, db):
    """Index lists identical; scores bit-equal except for rare 1-ulp fp64-summation-order differences, in which
    case the two orderings may swap neighbours whose scores are within 1 ulp."""
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert s.shape == os_.shape and i.shape == oi.shape
    ulp = np.abs(s.view(np.int32).astype(np.int64) - os_.view(np.int32).astype(np.int64))
    finite = np.isfinite(os_)
    assert (ulp[finite] <= 1).all(), "scores differ by more than 1 ulp"
    assert np.array_equal(np.isfinite(s), finite)
    bad = np.argwhere(i != oi)
    for r, c in bad:
        # permitted only between entries whose oracle scores are within 1 ulp of each other
        lo, hi = max(c - 1, 0), min(c + 1, s.shape[1] - 1)
        assert np.abs(os_[r, lo:hi + 1].view(np.int32).astype(np.int64) - int(os_[r, c].view(np.int32))).min() <= 1
        assert i[r, c] in oi[r, lo:hi + 1]
    assert len(bad) <= max(2, s.size // 1000)


@pytest.mark.parametrize("nq,ndb,d,k", [(5, 1000, 64, 10), (70, 5000, 512, 100), (3, 50, 32, 100), (17, 3001, 130, 7),
                                        (1, 1, 8, 1), (9, 777, 2048, 33)])
def test_exact_path(nq, ndb, d, k):
    from gandtr_b200 import _lib
    rs = np.random.RandomState(nq + ndb)
    q, db = unit_rows(rs, nq, d), unit_rows(rs, ndb, d)
    s, i = _lib.score_topk_exact(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), k, index_base=1000)
    os_, oi = R.topk(R.scores_exact(q, db), k, index_base=1000)
    _check_lists(s, i, os_, oi, q, db)


def test_exact_path_ties_lower_index_first():
    from gandtr_b200 import _lib
    rs = np.random.RandomState(3)
    base = unit_rows(rs, 40, 64)
    db = np.concatenate([base, base, base[:20]])          # every row has duplicates -> exact score ties
    q = unit_rows(rs, 6, 64)
    s, i = _lib.score_topk_exact(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), 25)
    os_, oi = R.topk(R.scores_exact(q, db), 25)
    assert np.array_equal(i.cpu().numpy(), oi)
    assert np.array_equal(s.cpu().numpy(), os_)


def _tc(q, db, k, index_base=0):
    from gandtr_b200 import _lib
    qd, dbd = torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda()
    shadow, nmax = _lib.db_prepare(dbd)
    s, i, st = _lib.score_topk(qd, dbd, shadow, nmax, k, index_base=index_base)
    torch.cuda.synchronize()
    return s, i, st.cpu().numpy()


@pytest.mark.parametrize("nq,ndb,d,k", [(64, 4096, 512, 100), (130, 20000, 512, 100), (70, 10000, 2048, 100),
                                        (5, 300, 64, 10), (200, 7001, 136, 50), (33, 100000, 512, 100)])
def test_tcgen05_path_matches_oracle(nq, ndb, d, k):
    rs = np.random.RandomState(nq * 7 + d)
    q, db = unit_rows(rs, nq, d), unit_rows(rs, ndb, d)
    s, i, st = _tc(q, db, k, index_base=5)
    assert st[0] == 0, "candidate overflow: status %s" % st
    os_, oi = R.topk(R.scores_exact(q, db), k, index_base=5)
    _check_lists(s, i, os_, oi, q, db)


def test_tcgen05_planted_neighbours_and_exact_cross_check():
    """Clustered data (queries are noisy copies of database rows) and on-device cross-check vs the exact kernel."""
    from gandtr_b200 import _lib
    rs = np.random.RandomState(11)
    nq, ndb, d, k = 256, 50000, 512, 100
    db = unit_rows(rs, ndb, d)
    src = rs.randint(0, ndb, nq)
    q = db[src] + 0.3 * rs.normal(0, 1, (nq, d)).astype(np.float32) / np.sqrt(d)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    s, i, st = _tc(q, db, k)
    assert st[0] == 0
    assert (i[:, 0].cpu().numpy() == src).all()
    se, ie = _lib.score_topk_exact(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), k)
    assert torch.equal(i, ie) and torch.equal(s, se)      # same summation order on both paths: bit-identical


def test_tcgen05_overflow_is_reported_not_hidden():
    """Degenerate database (all rows identical): every row ties, candidates overflow, the query is flagged."""
    rs = np.random.RandomState(2)
    row = unit_rows(rs, 1, 64)
    db = np.repeat(row, 20000, axis=0)
    q = unit_rows(rs, 4, 64)
    q[0] = row[0]
    s, i, st = _tc(q, db, 10)
    assert st[0] == -7 and st[2] >= 1
    assert (i[:, 0].cpu().numpy() == -2).any()


def test_topk_merge_equals_global_topk():
    from gandtr_b200 import _lib
    rs = np.random.RandomState(5)
    nq, ndb, d, k, g = 20, 4000, 64, 50, 4
    q, db = unit_rows(rs, nq, d), unit_rows(rs, ndb, d)
    qd = torch.from_numpy(q).cuda()
    parts_s, parts_i = [], []
    bounds = [0, 900, 2100, 2130, ndb]                     # ragged shards, one smaller than k
    for a, b in zip(bounds[:-1], bounds[1:]):
        s, i = _lib.score_topk_exact(qd, torch.from_numpy(db[a:b]).cuda(), k, index_base=a)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = _lib.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    os_, oi = R.topk(R.scores_exact(q, db), k)
    _check_lists(ms, mi, os_, oi, q, db)


def test_golden_reference_ranks_top100():
    """Top-100 of the reference's own argsort (float32 sgemm) vs ours, modulo reference near-ties (< 2e-7)."""
    from gandtr_b200 import _lib
    g = golden("map_eval.npz")
    q, db, ranks = g["q"], g["db"], g["ranks"]
    s, i = _lib.score_topk_exact(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), 100)
    i = i.cpu().numpy()
    ref_s = R.scores_reference(q, db)
    for qi in range(q.shape[0]):
        for pos in np.flatnonzero(i[qi] != ranks[:100, qi]):
            assert abs(float(ref_s[i[qi, pos], qi]) - float(ref_s[ranks[pos, qi], qi])) < 2e-7


def test_tcgen05_edge_shapes():
    """k larger than the shard (padding), k = 1, a single query against a long shard, and a ragged last tile."""
    for nq, ndb, d, k in [(128, 600, 2048, 1000), (300, 70000, 64, 1), (1, 200001, 128, 100), (129, 257, 512, 257)]:
        rs = np.random.RandomState(nq + ndb)
        q, db = unit_rows(rs, nq, d), unit_rows(rs, ndb, d)
        s, i, st = _tc(q, db, k)
        assert st[0] == 0, st
        os_, oi = R.topk(R.scores_exact(q, db), k)
        _check_lists(s, i, os_, oi, q, db)
        if k > ndb:
            assert (i[:, ndb:].cpu().numpy() == -1).all() and np.isneginf(s[:, ndb:].cpu().numpy()).all()


def test_index_repairs_overflowed_queries_exactly():
    """Massive exact ties (duplicated rows) overflow the candidate segments of some queries; ShardedIndex must hand back
    the exact ranking for them (index asc among equal scores)."""
    from gandtr_b200.retrieval import ShardedIndex
    rs = np.random.RandomState(8)
    base = unit_rows(rs, 50, 64)
    db = np.concatenate([np.repeat(base[:1], 30000, axis=0), unit_rows(rs, 40000, 64)])
    q = unit_rows(rs, 6, 64)
    q[0] = base[0]
    index = ShardedIndex(torch.from_numpy(db).cuda())
    s, i = index.search(torch.from_numpy(q).cuda(), 50)
    os_, oi = R.topk(R.scores_exact(q, db), 50)
    _check_lists(s, i, os_, oi, q, db)
    assert index.shard.last_status[0] == -7 and (i[0].cpu().numpy() == np.arange(50)).all()


def test_index_repairs_duplicate_clusters_on_the_tensor_path(monkeypatch):
    """A few thousand exact duplicates overflow the per-stripe candidate segments at k = 50; the retry with a larger k has
    room for them, so the exact CUDA-core scan is never needed."""
    from gandtr_b200 import retrieval
    rs = np.random.RandomState(9)
    base = unit_rows(rs, 3, 128)
    db = np.concatenate([unit_rows(rs, 60000, 128), np.repeat(base[:1], 2500, axis=0), unit_rows(rs, 60000, 128)])
    q = unit_rows(rs, 5, 128)
    q[2] = base[0]
    calls = []
    orig = retrieval.CudaOps.exact_topk
    monkeypatch.setattr(retrieval.CudaOps, "exact_topk", lambda self, *a, **k: calls.append(1) or orig(self, *a, **k))
    monkeypatch.setattr(retrieval, "TC_MIN_WORK", 1)
    index = retrieval.ShardedIndex(torch.from_numpy(db).cuda())
    s, i = index.search(torch.from_numpy(q).cuda(), 50)
    os_, oi = R.topk(R.scores_exact(q, db), 50)
    _check_lists(s, i, os_, oi, q, db)
    assert index.shard.last_status[0] == -7 and not calls and (i[2].cpu().numpy() == 60000 + np.arange(50)).all()


def test_unnormalised_and_tiny_magnitude_vectors():
    """Power-of-two scaling keeps the fp16 shadow in range: rows with norms ~1e4 and ~1e-4 rank exactly."""
    rs = np.random.RandomState(12)
    for scale in (1.0e4, 1.0e-4):
        q = (unit_rows(rs, 40, 256) * np.float32(scale) * rs.uniform(0.5, 2.0, (40, 1)).astype(np.float32)).astype(np.float32)
        db = (unit_rows(rs, 30000, 256) * np.float32(scale) * rs.uniform(0.2, 3.0, (30000, 1)).astype(np.float32)).astype(np.float32)
        s, i, st = _tc(q, db, 100)
        assert st[0] == 0
        os_, oi = R.topk(R.scores_exact(q, db), 100)
        _check_lists(s, i, os_, oi, q, db)


def test_packed_merge_rank_path_and_sort_fallback():
    """gdt_topk_merge_packed: sorted per-shard lists take the rank-by-binary-search path, anything else the bitonic sort;
    both must equal a plain descending sort of the valid keys (short lists, padding, a duplicated entry, unsorted input)."""
    from gandtr_b200 import _lib
    rs = np.random.RandomState(5)
    g, nq, k = 8, 37, 100
    S = np.full((g, nq, k), -np.inf, np.float32)
    I = np.full((g, nq, k), -1, np.int64)
    for s in range(g):
        for q in range(nq):
            m = int(rs.randint(0, k + 1)) if q % 3 else k          # ragged valid counts, some lists empty
            sc = np.sort(rs.standard_normal(m).astype(np.float32))[::-1]
            if q == 5 and m > 10:
                sc[3:9] = sc[3]                                     # equal scores inside one list: index breaks the tie
            ids = s * 1000 + np.sort(rs.choice(1000, m, replace=False))
            # order inside a list: score desc, index asc
            o = np.lexsort((ids, -sc.astype(np.float64)))
            S[s, q, :m], I[s, q, :m] = sc[o], ids[o]
    S[1, 7], I[1, 7] = S[0, 7], I[0, 7]                             # the same list twice: duplicates across lists
    unsorted_q = 11
    S[2, unsorted_q, :k] = S[2, unsorted_q, :k][::-1].copy(); I[2, unsorted_q, :k] = I[2, unsorted_q, :k][::-1].copy()
    St, It = torch.from_numpy(S).cuda(), torch.from_numpy(I).cuda()
    ms, mi = _lib.topk_merge_packed(_lib.topk_pack(St, It))
    ms, mi = ms.cpu().numpy(), mi.cpu().numpy()
    for q in range(nq):
        valid = I[:, q, :] >= 0
        sc, ids = S[:, q, :][valid], I[:, q, :][valid]
        o = np.lexsort((ids, -sc.astype(np.float64)))[:k]
        n = len(o)
        assert np.array_equal(mi[q, :n], ids[o]), q
        assert np.array_equal(ms[q, :n], sc[o]), q
        assert (mi[q, n:] == -1).all() and np.isneginf(ms[q, n:]).all(), q


def test_merge_to_keys_and_unpack_round_trip():
    """gdt_topk_merge_packed_keys + gdt_topk_unpack (the query-sharded merge's two halves) == gdt_topk_merge_packed,
    overflow marker included."""
    from gandtr_b200 import _lib
    rs = np.random.RandomState(9)
    g, nq, k = 4, 19, 50
    S = np.sort(rs.standard_normal((g, nq, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    I = np.stack([s * 1000 + np.sort(rs.choice(1000, (nq, k)), axis=1) for s in range(g)]).astype(np.int64)
    S[1, 3, 20:], I[1, 3, 20:] = -np.inf, -1                      # a short list
    I[2, 5, 0] = -2                                               # shard 2 overflowed on query 5
    keys = _lib.topk_pack(torch.from_numpy(S).cuda(), torch.from_numpy(I).cuda())
    ms, mi = _lib.topk_merge_packed(keys)
    us, ui = _lib.topk_unpack(_lib.topk_merge_packed_keys(keys))
    assert int(mi[5, 0]) == -2 and int(ui[5, 0]) == -2
    ok = [q for q in range(nq) if q != 5]
    assert torch.equal(mi[ok], ui[ok]) and torch.equal(ms[ok], us[ok])
