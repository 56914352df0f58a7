"""The row-sharded (two-phase, histogram-exchange) K3 path on ONE GPU -- the headline multi-GPU path of BASELINE config 4,
exercised without NCCL so that a single-GPU box produces parity evidence for it:

  * by hand through the C ABI: gdt_db_prepare_norm / _convert with the statistics made common across the shards,
    gdt_score_topk_filter per shard, the [nq, 256] histograms summed, gdt_score_topk_finalize per shard against the GLOBAL
    threshold, gdt_topk_pack + gdt_topk_merge_packed (and the unpacked gdt_topk_merge) -- against the oracle ranking of the
    whole matrix;
  * through `ShardedIndex` with the ranks emulated as threads (tests/util.py ThreadComm): the very code the NCCL ranks
    run -- including the empty-shard branch, the deferred overflow check, the local repair and the mAP position exchange;
  * the measured error of the tensor-core coarse scores against the bound E_q the filter's exactness rests on.
The NCCL variants of the first two live in tests/test_gpu_multi.py (2 GPUs)."""
import numpy as np
import pytest
import torch

from oracle import retrieval_np as R
from tests.test_gpu_topk import _check_lists
from tests.util import ThreadComm, run_ranks, unit_rows

pytestmark = pytest.mark.gpu


def _prepare_shards(parts):
    """db_prepare_sharded for every shard of one GPU, with the all-reduce(MAX) done by hand: phase 1 (norms) for all
    shards, maximum, phase 2 (convert) for all shards, maximum."""
    from gandtr_b200 import _lib
    lib = _lib.load()
    stats = [torch.zeros(4, dtype=torch.float32, device="cuda") for _ in parts]
    shadows = [torch.empty(p.shape, dtype=torch.float16, device="cuda") for p in parts]
    for p, st in zip(parts, stats):
        if p.shape[0]:
            _lib.check(lib.gdt_db_prepare_norm(_lib._ptr(p), p.shape[0], p.shape[1], _lib._ptr(st), _lib._stream()), "norm")
    gmax = torch.stack(stats).max(0).values
    for st in stats:
        st[0] = gmax[0]
    for p, sh, st in zip(parts, shadows, stats):
        if p.shape[0]:
            _lib.check(lib.gdt_db_prepare_convert(_lib._ptr(p), p.shape[0], p.shape[1], _lib._ptr(sh), _lib._ptr(st),
                                                  _lib._stream()), "convert")
    gmax = torch.stack(stats).max(0).values
    for st in stats:
        st[1:4] = gmax[1:4]
    return shadows, stats


def _two_phase_by_hand(q, db, bounds, k, exchange=True):
    from gandtr_b200 import _lib
    qd = torch.from_numpy(q).cuda()
    parts = [torch.from_numpy(db[a:b]).cuda().contiguous() for a, b in zip(bounds[:-1], bounds[1:])]
    shadows, stats = _prepare_shards(parts)
    live = [j for j, p in enumerate(parts) if p.shape[0]]
    states = {j: _lib.score_topk_filter(qd, shadows[j], stats[j], k) for j in live}
    if exchange:
        total = torch.stack([states[j].hist for j in live]).sum(0).to(torch.int32)
        for j in live:
            states[j].hist.copy_(total)                      # what the all-reduce(SUM) leaves on every rank
    lists, survivors = [], []
    for j in range(len(parts)):
        if j in states:
            s, i, st = _lib.score_topk_finalize(qd, parts[j], states[j], index_base=bounds[j])
            survivors.append(int(st[1]))
        else:
            s = torch.full((q.shape[0], k), float("-inf"), device="cuda")
            i = torch.full((q.shape[0], k), -1, dtype=torch.int64, device="cuda")
        lists.append((s, i))
    S, I = torch.stack([x[0] for x in lists]), torch.stack([x[1] for x in lists])
    merged = _lib.topk_merge(S, I)
    packed = _lib.topk_merge_packed(_lib.topk_pack(S, I))
    assert torch.equal(merged[0], packed[0]) and torch.equal(merged[1], packed[1])    # both exchange formats agree
    return merged, lists, survivors


@pytest.mark.parametrize("nq,ndb,d,k,bounds", [
    (150, 60001, 256, 100, [0, 20000, 40001, 60001]),
    (70, 30000, 512, 100, [0, 29000, 29000, 30000]),          # an empty shard and a shard smaller than the seed range
    (129, 90000, 128, 10, [0, 45000, 90000]),
    (64, 40000, 2048, 100, [0, 10000, 20000, 30000, 40000]),
])
def test_two_phase_through_the_c_abi_equals_the_oracle(nq, ndb, d, k, bounds):
    rs = np.random.RandomState(nq + ndb + d)
    db = unit_rows(rs, ndb, d)
    src = rs.randint(0, ndb, nq)
    q = db[src] + 0.4 * rs.normal(0, 1, (nq, d)).astype(np.float32) / np.sqrt(d)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    (s, i), lists, survivors = _two_phase_by_hand(q, db, bounds, k)
    os_, oi = R.topk(R.scores_exact(q, db), k)
    _check_lists(s, i, os_, oi, q, db)
    assert not (i[:, 0] == -2).any()
    # the exchange is what makes a shard re-score only ITS members of the global top k: every per-shard list holds fewer
    # valid entries than k (unless it owns them all), and together they hold just over k
    valid = torch.stack([(x[1] >= 0).sum(1) for x in lists])               # [shards, nq]
    assert int(valid.sum(0).min()) >= k
    assert max(survivors) < k + 100, survivors
    # without the exchange every shard returns a full local top k (the merge is still exact)
    (s2, i2), lists2, _ = _two_phase_by_hand(q, db, bounds, k, exchange=False)
    _check_lists(s2, i2, os_, oi, q, db)
    nonempty = [j for j in range(len(bounds) - 1) if bounds[j + 1] - bounds[j] >= k]
    assert all(int((lists2[j][1] >= 0).sum(1).min()) == k for j in nonempty)


def test_two_phase_near_ties_and_duplicates_overflow_then_repair():
    """Dense exact ties (duplicated rows) in one shard overflow that shard's candidate segments: the marker must survive
    pack / merge, and the repaired result must be the exact ranking (ties -> lower index first)."""
    from gandtr_b200 import _lib
    from gandtr_b200.retrieval import CudaOps, DatabaseShard
    rs = np.random.RandomState(21)
    d, k = 64, 50
    base = unit_rows(rs, 4, d)
    near = base[1] + 1e-6 * rs.normal(0, 1, (3000, d)).astype(np.float32)             # near-ties around another row
    db = np.concatenate([unit_rows(rs, 25000, d), np.repeat(base[:1], 30000, axis=0), near.astype(np.float32),
                         unit_rows(rs, 22000, d)])
    bounds = [0, 25000, 58000, 80000]
    q = unit_rows(rs, 9, d)
    q[0], q[5] = base[0], base[1]
    (s, i), lists, _ = _two_phase_by_hand(q, db, bounds, k)
    flagged = torch.nonzero(i[:, 0] == -2).flatten().tolist()
    assert 0 in flagged                                                 # 30 000 identical rows cannot fit any segment
    assert (lists[1][1][0, 0] == -2) and not (lists[0][1][:, 0] == -2).any()
    os_, oi = R.topk(R.scores_exact(q, db), k)
    ok = [r for r in range(q.shape[0]) if r not in flagged]
    _check_lists(s[ok], i[ok], os_[ok], oi[ok], q[ok], db)             # unflagged queries are already exact
    # repair exactly as ShardedIndex.search does: exact local lists of the flagged queries from every shard, merged
    ops = CudaOps()
    qd = torch.from_numpy(q).cuda()[flagged].contiguous()
    rep = [ops.repair_topk(qd, DatabaseShard(torch.from_numpy(db[a:b]).cuda(), index_base=a), k)
           for a, b in zip(bounds[:-1], bounds[1:])]
    rs_, ri = _lib.topk_merge_packed(_lib.topk_pack(torch.stack([x[0] for x in rep]), torch.stack([x[1] for x in rep])))
    s[flagged], i[flagged] = rs_, ri
    _check_lists(s, i, os_, oi, q, db)
    assert (i[0].cpu().numpy() == 25000 + np.arange(k)).all()


def _planted(rs, nq, ndb, d):
    db = unit_rows(rs, ndb, d)
    src = rs.randint(0, ndb, nq)
    q = db[src] + 0.4 * rs.normal(0, 1, (nq, d)).astype(np.float32) / np.sqrt(d)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32), db, src


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_index_with_emulated_ranks(world):
    """ShardedIndex.search / evaluate_map exactly as the NCCL ranks run them, the ranks being threads on one GPU."""
    from gandtr_b200.retrieval import ShardedIndex, evaluate_map
    rs = np.random.RandomState(7)
    nq, ndb, d, k = 151, 60001, 256, 100              # 151 queries: the query-sharded merge pads its last slice
    q, db, src = _planted(rs, nq, ndb, d)
    os_, oi = R.topk(R.scores_exact(q, db), k)
    gnd = [{"ok": np.array([src[j]]), "junk": np.array([(src[j] + 1) % ndb])} for j in range(nq)]
    mo, apo, _, _ = R.compute_map(R.full_ranks(R.scores_exact(q, db)), gnd)
    dbd = torch.from_numpy(db).cuda()

    def rank_fn(rank, comm):
        index = ShardedIndex.from_full(dbd, comm=comm)
        assert "comm" in index.shard.aux and index.world == world
        qd = torch.from_numpy(q).cuda() if rank == 0 else torch.zeros((nq, d), device="cuda")
        s, i = index.search(qd, k, broadcast=True)
        _check_lists(s, i, os_, oi, q, db)
        s2, i2 = index.search_from_host(torch.from_numpy(q).pin_memory(), k)     # sliced upload + all-gather of the queries
        assert torch.equal(i2, i) and torch.equal(s2, s)
        st = index.shard.last_status
        assert st[0] == 0 and st[1] < k + 40, st              # only this shard's share of the global top k was re-scored
        m, aps, _, _ = evaluate_map(index, qd, gnd)
        assert abs(m - mo) < 1e-12 and np.array_equal(aps, apo)
        return comm.calls
    calls = run_ranks(world, rank_fn)
    # per search: 1 broadcast, 1 histogram all-reduce, 1 all-to-all of the packed lists (query-sharded merge), 1 packed
    # all-gather of the merged slices; index construction: 2 all-reduce(MAX)
    # (+ the second search through search_from_host: one more all-gather for the queries, one for its merged slices)
    assert calls[0]["all_gather"] == 3 and calls[0]["all_reduce_max"] == 2 and calls[0]["broadcast"] == 1
    assert calls[0]["all_to_all"] == 2


def test_sharded_index_emulated_ranks_empty_shard_and_overflow_repair():
    """Custom partition with an EMPTY last shard (it still takes part in every collective) and a shard full of duplicated
    rows whose candidate lists overflow: the marker travels through pack / all-gather / merge, every rank takes the same
    repair decision after the one end-of-search check, and the result is the exact ranking."""
    from gandtr_b200 import retrieval
    rs = np.random.RandomState(33)
    d, k = 64, 50
    base = unit_rows(rs, 2, d)
    db = np.concatenate([unit_rows(rs, 30000, d), np.repeat(base[:1], 30000, axis=0), unit_rows(rs, 20000, d)])
    bounds = [0, 30000, 80000, 80000]
    q = unit_rows(rs, 6, d)
    q[0] = base[0]
    os_, oi = R.topk(R.scores_exact(q, db), k)
    dbd = torch.from_numpy(db).cuda()
    old = retrieval.TC_MIN_WORK
    retrieval.TC_MIN_WORK = 1
    try:
        def rank_fn(rank, comm):
            index = retrieval.ShardedIndex(dbd[bounds[rank]:bounds[rank + 1]].contiguous(), n_total=len(db),
                                           index_base=bounds[rank], comm=comm)
            s, i = index.search(torch.from_numpy(q).cuda(), k)
            _check_lists(s, i, os_, oi, q, db)
            assert (i[0].cpu().numpy() == 30000 + np.arange(k)).all()
            return comm.calls["all_gather"]
        gathers = run_ranks(3, rank_fn)
    finally:
        retrieval.TC_MIN_WORK = old
    assert gathers == [2, 2, 2]                                  # the search itself + the repaired queries


@pytest.mark.parametrize("kind,nq,ndb,d", [("gauss", 96, 30000, 512), ("positive", 64, 20000, 2048), ("gauss", 40, 6000, 8192),
                                           ("positive", 40, 6000, 8192), ("near_ties", 64, 20000, 512),
                                           ("wide_norms", 64, 20000, 256)])
def test_k3_coarse_score_error_stays_inside_the_filter_bound(kind, nq, ndb, d, record_property):
    """The tcgen05 pass only filters, and the filter is exact iff |coarse - sq*sx*<q,x>| <= E_q for every pair (header of
    score_topk_sm100.cu). Measured here against fp64 on the raw tensor-core scores (gdt_debug_k3_coarse_scores): random,
    all-positive (every product has the same sign: worst case for accumulation error), near-tie, d = 8192 and
    wide-dynamic-range data. The accumulation part is also isolated (coarse vs the fp64 dot of the SAME fp16 operands) and
    compared with its allowance acc(d) = (d/16 + 16) * 2^-20, the term that was chosen without a hardware specification."""
    from gandtr_b200 import _lib
    rs = np.random.RandomState(d + ndb)
    if kind == "gauss":
        q, db = unit_rows(rs, nq, d), unit_rows(rs, ndb, d)
    elif kind == "positive":
        q = np.abs(unit_rows(rs, nq, d)) + 0.02
        db = np.abs(unit_rows(rs, ndb, d)) + 0.02
        q, db = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32), (db / np.linalg.norm(db, axis=1, keepdims=True)).astype(np.float32)
    elif kind == "near_ties":
        centre = unit_rows(rs, 8, d)
        db = np.repeat(centre, ndb // 8, axis=0) + 1e-4 * rs.normal(0, 1, (ndb // 8 * 8, d)).astype(np.float32)
        db = (db / np.linalg.norm(db, axis=1, keepdims=True)).astype(np.float32)
        q = np.repeat(centre, nq // 8, axis=0).astype(np.float32)
        ndb = db.shape[0]
    else:
        q = (unit_rows(rs, nq, d) * rs.uniform(1e-3, 1e3, (nq, 1))).astype(np.float32)
        db = (unit_rows(rs, ndb, d) * np.exp(rs.uniform(-6, 6, (ndb, 1)))).astype(np.float32)
    qd, dbd = torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda()
    shadow, stats = _lib.db_prepare(dbd)
    coarse, meta = _lib.debug_k3_coarse_scores(qd, shadow, stats, 100)
    torch.cuda.synchronize()
    assert not torch.isnan(coarse).any()                                  # every (query, row) pair was produced
    sx = float(stats[1])
    sq = meta[:, 3].double()
    exact = qd.double() @ dbd.double().t()                                # products of fp32 values are exact in fp64
    err = (coarse.double() - exact * sq[:, None] * sx).abs().max(1).values
    e_q = meta[:, 2].double() / 2
    ratio = float((err / e_q).max())
    # accumulation error alone: same fp16 operands, fp64 accumulation
    q16 = (qd * meta[:, 3:4]).half().double()
    c16 = q16 @ shadow.double().t()
    acc_err = (coarse.double() - c16).abs().max(1).values
    acc_allow = (d / 16 + 16) * 2.0 ** -20 * q16.norm(dim=1) * float(stats[3])
    acc_ratio = float((acc_err / acc_allow).max())
    record_property("bound_ratio", ratio)
    record_property("acc_ratio", acc_ratio)
    print("K3 bound check %-10s d=%-5d max|coarse-exact|/E_q = %.4f   accumulation / allowance = %.4f" % (kind, d, ratio, acc_ratio))
    assert ratio < 1.0, "coarse scores leave the bound the filter relies on"
    assert acc_ratio < 0.5, "the accumulation allowance (2x an estimate) has less than 2x head-room"
    # and the end result on the same data is the oracle ranking. Scores that crowd into one histogram bin (near-ties;
    # all-positive vectors, whose cosines sit within a few percent of each other) overflow the candidate segments: that
    # is reported, never hidden, and ShardedIndex repairs those queries exactly
    from gandtr_b200.retrieval import ShardedIndex
    s, i, st = _lib.score_topk(qd, dbd, shadow, stats, 100)
    os_, oi = R.topk(R.scores_exact(q, db), 100)
    if int(st[0]) == 0:
        _check_lists(s, i, os_, oi, q, db)
    else:
        assert kind in ("near_ties", "positive") and int(st[0]) == -7
        s, i = ShardedIndex(dbd).search(qd, 100)
        _check_lists(s, i, os_, oi, q, db)
