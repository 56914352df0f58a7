"""K4 parity (through the C ABI): ranks of ground-truth ids and AP / P@k against the reference's golden numbers."""
import numpy as np
import pytest
import torch

from oracle import retrieval_np as R
from tests.util import golden

pytestmark = pytest.mark.gpu


def _pad(lists, fill=-1):
    n = max(1, max(len(x) for x in lists))
    return np.array([np.pad(np.asarray(x, dtype=np.int64), (0, n - len(x)), constant_values=fill) for x in lists])


def _ap_on_gpu(q, db, ok, junk, kappas, shards=1):
    from gandtr_b200 import _lib
    nq = q.shape[0]
    probes = _pad([np.concatenate([o, j]) for o, j in zip(ok, junk)])
    pd = torch.from_numpy(probes).cuda()
    qd = torch.from_numpy(q).cuda()
    bounds = np.linspace(0, db.shape[0], shards + 1).astype(int)
    dbs = [torch.from_numpy(db[a:b]).cuda() for a, b in zip(bounds[:-1], bounds[1:])]
    ps = torch.zeros(probes.shape, dtype=torch.float32, device="cuda")
    for a, dbd in zip(bounds[:-1], dbs):
        _lib.probe_scores(qd, dbd, pd, index_base=int(a), out=ps)
    before = torch.zeros(probes.shape, dtype=torch.int64, device="cuda")
    for a, dbd in zip(bounds[:-1], dbs):
        _lib.rank_counts(qd, dbd, pd, ps, index_base=int(a), out=before)
    before = before.cpu().numpy()
    pos = _pad([before[i, :len(o)] for i, o in enumerate(ok)], 0)
    jnk = _pad([before[i, len(o):len(o) + len(j)] for i, (o, j) in enumerate(zip(ok, junk))], 0)
    npos = torch.tensor([len(o) for o in ok], dtype=torch.int32, device="cuda")
    njunk = torch.tensor([len(j) for j in junk], dtype=torch.int32, device="cuda")
    ap, prk = _lib.map_eval(torch.from_numpy(pos).cuda(), torch.from_numpy(jnk).cuda(), npos, njunk, kappas)
    return ap.cpu().numpy(), prk.cpu().numpy(), before


@pytest.mark.parametrize("shards", [1, 3])
def test_revisited_protocol_matches_reference_golden(shards):
    g = golden("map_eval.npz")
    q, db = g["q"], g["db"]
    easy = [x[x >= 0] for x in g["easy"]]
    hard = [x[x >= 0] for x in g["hard"]]
    junk = [x[x >= 0] for x in g["junk"]]
    # our total order vs the oracle's full ranking
    ranks = R.full_ranks(R.scores_exact(q, db))
    for name, ok, jk in (("easy", easy, [np.concatenate([j, h]) for j, h in zip(junk, hard)]),
                         ("medium", [np.concatenate([e, h]) for e, h in zip(easy, hard)], junk),
                         ("hard", hard, [np.concatenate([j, e]) for j, e in zip(junk, easy)])):
        ap, prk, before = _ap_on_gpu(q, db, ok, jk, [1, 5, 10], shards)
        m, aps, mpr, prs = R.compute_map(ranks, [{"ok": o, "junk": j} for o, j in zip(ok, jk)], [1, 5, 10])
        np.testing.assert_array_equal(ap, aps)                 # float64, same op order: bit-equal incl. NaN
        np.testing.assert_array_equal(prk, prs)
        # and against the unmodified reference (its ranks come from a float32 sgemm: allow near-tie swaps)
        np.testing.assert_allclose(ap, g["ap_" + name], rtol=0, atol=1e-6, equal_nan=True)
        valid = ~np.isnan(ap)
        assert abs(ap[valid].sum() / valid.sum() - float(g["map_" + name])) < 1e-6
        # positions themselves
        for i, o in enumerate(ok):
            inv = np.empty(db.shape[0], dtype=np.int64)
            inv[ranks[:, i]] = np.arange(db.shape[0])
            assert np.array_equal(before[i, :len(o)], inv[o])


def test_old_protocol_with_empty_query():
    g = golden("map_eval.npz")
    q, db = g["q"], g["db"]
    ok = [np.concatenate([e[e >= 0], h[h >= 0]]) for e, h in zip(g["easy"], g["hard"])]
    junk = [x[x >= 0] for x in g["junk"]]
    ok[3] = np.array([], dtype=np.int64)
    ap, _, _ = _ap_on_gpu(q, db, ok, junk, [])
    assert np.isnan(ap[3])
    np.testing.assert_allclose(ap, g["old_ap"], rtol=0, atol=1e-6, equal_nan=True)
    valid = ~np.isnan(ap)
    assert abs(ap[valid].sum() / valid.sum() - float(g["old_map"])) < 1e-6


def _roxford_shaped(d=512, seed=1):
    """4 993 'real' + 100 000 distractor unit vectors, 70 queries around planted positives, per-query
    easy / hard / junk id lists (some queries without hard positives)."""
    rs = np.random.RandomState(seed)
    ndb, nq = 4993 + 100000, 70
    db = rs.standard_normal((ndb, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    rs2 = np.random.RandomState(seed + 1)
    gnd, q = [], []
    for i in range(nq):
        ids = rs2.permutation(4993)[:110]
        ne, nh, nj = rs2.randint(10, 31), (0 if i % 9 == 8 else rs2.randint(10, 41)), rs2.randint(10, 41)
        easy, hard, junk = ids[:ne], ids[ne:ne + nh], ids[ne + nh:ne + nh + nj]
        centre = db[np.concatenate([easy, hard])].mean(0)
        v = centre / np.linalg.norm(centre) + 0.35 * rs2.standard_normal(d) / np.sqrt(d)
        # pull the positives towards the query so that they rank high among the distractors
        db[easy] = db[easy] + rs2.uniform(-0.02, 0.14, (len(easy), 1)).astype(np.float32) * v / np.linalg.norm(v)
        db[hard] = db[hard] + rs2.uniform(-0.06, 0.08, (len(hard), 1)).astype(np.float32) * v / np.linalg.norm(v)
        q.append((v / np.linalg.norm(v)).astype(np.float32))
        gnd.append({"bbx": None, "easy": easy, "hard": hard, "junk": junk})
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    return np.stack(q), db.astype(np.float32), gnd


def test_roxford_shaped_eval_full_size_matches_oracle():
    """BASELINE config 3 at full size: ranking + mAP (Easy / Medium / Hard) + mP@k on the GPU against the oracle's
    compute_map_and_print restatement on the full 104 993 x 70 ranking."""
    from gandtr_b200.retrieval import ShardedIndex, compute_map_and_print
    q, db, gnd = _roxford_shaped()
    index = ShardedIndex(torch.from_numpy(db).cuda())
    qd = torch.from_numpy(q).cuda()
    s, i = index.search(qd, 100)
    os_, oi = R.topk(R.scores_exact(q, db), 100)
    assert np.array_equal(i.cpu().numpy(), oi)
    avg, per = compute_map_and_print("roxford5k", index, qd, gnd, printer=lambda *_: None)
    ranks = R.full_ranks(R.scores_exact(q, db))
    oavg, oaps, _ = R.compute_map_protocols("roxford5k", ranks, gnd)
    for name in ("easy", "medium", "hard"):
        assert abs(avg["map_" + name] - oavg["map_" + name]) < 1e-12
        np.testing.assert_array_equal(per["ap_" + name], oaps["ap_" + name])
    assert np.isnan(per["ap_hard"][8]) and 0.2 < avg["map_medium"] <= 1.0


def test_duplicated_and_foreign_ground_truth_ids_follow_in1d_semantics():
    """`np.in1d(ranks[:, i], qgnd)` (evaluate.py:75-76) finds a duplicated id once and an id outside the database never,
    while `compute_ap(pos, len(qgnd))` keeps the list length as given: evaluate_map must do the same."""
    from gandtr_b200.retrieval import ShardedIndex, evaluate_map, compute_map_and_print
    from tests.util import unit_rows
    rs = np.random.RandomState(17)
    ndb, nq, d = 3000, 12, 64
    db, q = unit_rows(rs, ndb, d), unit_rows(rs, nq, d)
    gnd = []
    for i in range(nq):
        ok = rs.permutation(ndb)[:9]
        junk = rs.permutation(ndb)[:5]
        if i % 3 == 0:
            ok = np.concatenate([ok, ok[:3]])                           # duplicates
        if i % 4 == 1:
            ok = np.concatenate([ok, [ndb + 5, ndb + 77]])              # ids no shard owns
            junk = np.concatenate([junk, junk[:2], [ndb + 1]])
        gnd.append({"ok": ok, "junk": junk})
    gnd[7]["ok"] = np.array([], dtype=np.int64)                         # empty query: NaN, excluded
    index = ShardedIndex(torch.from_numpy(db).cuda())
    qd = torch.from_numpy(q).cuda()
    m, aps, mpr, prs = evaluate_map(index, qd, gnd, [1, 5, 10])
    mo, apo, mpro, prso = R.compute_map(R.full_ranks(R.scores_exact(q, db)), gnd, [1, 5, 10])
    assert np.array_equal(aps, apo, equal_nan=True) and np.isnan(aps[7])
    assert abs(m - mo) < 1e-15
    np.testing.assert_array_equal(prs, prso)
    np.testing.assert_allclose(mpr, mpro, rtol=0, atol=1e-15)
    # the three revisited protocols in one pass (one rank_counts + one map_eval launch) equal three separate evaluations
    rg = [{"bbx": None, "easy": g["ok"][:4], "hard": g["ok"][4:], "junk": g["junk"]} for g in gnd]
    avg, per = compute_map_and_print("roxford5k", index, qd, rg, printer=lambda *_: None)
    oavg, oper, _ = R.compute_map_protocols("roxford5k", R.full_ranks(R.scores_exact(q, db)), rg)
    for name in ("easy", "medium", "hard"):
        assert np.array_equal(per["ap_" + name], oper["ap_" + name], equal_nan=True)
        assert abs(avg["map_" + name] - oavg["map_" + name]) < 1e-15
    # ground truth prepared once (validation loops evaluate the same dataset again and again): same numbers
    from gandtr_b200.retrieval import PreparedGroundTruth
    prep = PreparedGroundTruth("roxford5k", rg, index.n_total, "cuda")
    for _ in range(2):
        avg2, per2 = compute_map_and_print("roxford5k", index, qd, prep, printer=lambda *_: None)
        assert avg2 == avg and all(np.array_equal(per2[k], per[k], equal_nan=True) for k in per)


@pytest.mark.parametrize("d,ndb", [(64, 5000), (512, 20001), (2048, 3000), (130, 4000)])
def test_rank_counts_fp32_bound_path_equals_exact_rescoring(d, ndb):
    """gdt_rank_counts ranks a row from its fp32 score when the rigorous error interval contains no probe key, and
    re-scores exactly otherwise. Counts must equal (a) the same kernel with every pair re-scored exactly
    (gdt_debug_k4_exact) and (b) the oracle's positions -- including exact ties (duplicated rows: index order decides) and
    near-ties a few float32 ulps apart."""
    from gandtr_b200 import _lib
    from tests.util import unit_rows
    lib = _lib.load()
    rs = np.random.RandomState(d + ndb)
    nq, pmax = 19, 37
    db = unit_rows(rs, ndb, d)
    q = unit_rows(rs, nq, d)
    probes = np.stack([rs.permutation(ndb)[:pmax] for _ in range(nq)]).astype(np.int64)
    probes[:, -3:] = -1                                                   # padding
    for i in range(nq):                                                   # exact ties and near-ties around the probes
        db[(probes[i, 0] + 1) % ndb] = db[probes[i, 0]]
        db[(probes[i, 1] + 2) % ndb] = db[probes[i, 1]] * np.float32(1 + 2e-7)
    qd, dbd, pd = torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), torch.from_numpy(probes).cuda()
    ps = _lib.probe_scores(qd, dbd, pd)
    fast = _lib.rank_counts(qd, dbd, pd, ps)
    try:
        _lib.check(lib.gdt_debug_k4_exact(1), "exact")
        exact = _lib.rank_counts(qd, dbd, pd, ps)
    finally:
        _lib.check(lib.gdt_debug_k4_exact(0), "exact")
    assert torch.equal(fast, exact)
    ranks = R.full_ranks(R.scores_exact(q, db))                           # [ndb, nq]
    inv = np.empty_like(ranks)
    for i in range(nq):
        inv[ranks[:, i], i] = np.arange(ndb)
    got = fast.cpu().numpy()
    for i in range(nq):
        valid = probes[i] >= 0
        assert np.array_equal(got[i, valid], inv[probes[i, valid], i]), "query %d" % i
        assert (got[i, ~valid] == 0).all()
    # accumulation across shards: two calls with different index_base add up to the single-shard counts
    half = ndb // 2
    acc = torch.zeros_like(fast)
    ps2 = torch.zeros_like(ps)
    _lib.probe_scores(qd, dbd[:half].contiguous(), pd, index_base=0, out=ps2)
    _lib.probe_scores(qd, dbd[half:].contiguous(), pd, index_base=half, out=ps2)
    _lib.rank_counts(qd, dbd[:half].contiguous(), pd, ps2, index_base=0, out=acc)
    _lib.rank_counts(qd, dbd[half:].contiguous(), pd, ps2, index_base=half, out=acc)
    assert torch.equal(acc, fast)
