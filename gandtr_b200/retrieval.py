"""Query x database scoring, ranking and mAP on row-sharded databases.

Replaces the tail of `CirDatasetAp.__call__` (mdir/components/optim/score/cirscore.py:66-73):
    scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0); compute_map_and_print(dataset, ranks, gnd)
and the consumers of `ranks` inside mdir/external/cirtorch/utils/evaluate.py:39-152.

Layout: the database is sharded ROW-WISE over the ranks of a torch.distributed group (rank r owns the global rows
[lo_r, hi_r) of `shard_bounds`), queries are replicated (broadcast from rank 0), every rank runs the fused
score + top-k kernel on its shard, the per-shard (score, global index) lists are exchanged with ONE all_gather of
[nq, k] and merged by `gdt_topk_merge`. mAP needs the full-ranking positions of the ground-truth ids only: per-shard
"rows that sort before" counts are summed with one all_reduce and fed to `gdt_map_eval`. The full ndb x nq score /
rank matrices of the reference are never materialised.

The compute callables are injectable (`ops=`) so the collective plumbing can be exercised on CPU/gloo in the tests;
the default `CudaOps` calls libgandtr_b200.so and raises when it is missing -- there is no CPU fallback in the
product.
"""
import os
import weakref

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "DistComm", "CudaOps", "DatabaseShard", "ShardedIndex", "evaluate_map", "evaluate_protocols",
           "compute_map_and_print", "PreparedProtocols", "PreparedGroundTruth", "TC_MIN_WORK", "EXACT_MAX_K", "MERGE_MAX_ENTRIES"]

# below this many multiply-adds the exact CUDA-core kernel is used instead of the tcgen05 pipeline
TC_MIN_WORK = 1 << 24
# kernel limits (score_exact_sm100.cu): lists longer than EXACT_MAX_K take the full-sort path, merges of more than
# MERGE_MAX_ENTRIES (= shards x k) entries run in rounds
EXACT_MAX_K = 4096
MERGE_MAX_ENTRIES = 16384
# the exact fallback streams the database in row blocks so that its score workspace stays below this many floats
EXACT_WS_FLOATS = 1 << 26


def shard_bounds(n_total, world_size, rank):
    """Contiguous block partition of `n_total` rows: rank r owns [lo, hi). Sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class DistComm:
    """The collectives of the sharded path over one torch.distributed process group (NCCL over NVLink on GPUs, gloo in
    the CPU tests). Tests may inject another object with the same five methods (tests/util.py `ThreadComm` emulates
    several ranks on ONE GPU)."""

    def __init__(self, group=None):
        self.group = group
        self.world, self.rank = _world(group)

    def _nccl(self):
        return self.world > 1 and dist.get_backend(self.group) == "nccl"

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_gather(self, t):
        """-> [world, *t.shape], rank-major."""
        t = t.contiguous()
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        if self.world == 1:
            out[0] = t
        elif self._nccl():
            dist.all_gather_into_tensor(out, t, group=self.group)      # one collective, no staging copies
        else:
            dist.all_gather(list(out.unbind(0)), t, group=self.group)   # gloo (CPU tests)
        return out

    def all_to_all(self, t):
        """t: [world, m, ...], chunk r goes to rank r -> [world, m, ...], chunk r came from rank r."""
        if self.world == 1:
            return t.clone()
        t = t.contiguous()
        if not self._nccl():       # gloo (CPU tests of the host logic): the same exchange through an all-gather
            return self.all_gather(t)[:, self.rank].contiguous()
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=self.group)
        return out

    def broadcast(self, t, src=0):
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def sum_int(self, v):
        if self.world == 1:
            return int(v)
        t = torch.tensor([int(v)], dtype=torch.int64)
        if self._nccl():
            t = t.cuda()
        dist.all_reduce(t, group=self.group)
        return int(t.item())


class CudaOps:
    """The product's compute: every method is one or more launches of libgandtr_b200.so kernels."""

    def prepare(self, db, comm=None):
        from . import _lib
        d = db.shape[1]
        if d % 8 != 0 or d > 8192:
            return {}
        if comm is not None and comm.world > 1:
            # common scale / error bound on every rank, so that the per-query score histograms can be summed
            shadow, stats = _lib.db_prepare_sharded(db, comm.all_reduce_max)
            return {"shadow": shadow, "norm_max": stats, "comm": comm}
        if db.shape[0] > 0:
            shadow, norm_max = _lib.db_prepare(db)
            return {"shadow": shadow, "norm_max": norm_max}
        return {}

    @staticmethod
    def _empty_lists(nq, k, device):
        return (torch.full((nq, k), float("-inf"), dtype=torch.float32, device=device),
                torch.full((nq, k), -1, dtype=torch.int64, device=device))

    def local_topk(self, q, shard, k):
        """(scores [nq,k], global idx [nq,k]) of one shard; exact ordering (score desc, index asc).
        Single shard: complete and repaired. On a sharded index the lists hold this shard's members of the GLOBAL top k
        (fewer than k valid entries, (-inf, -1) padding) and are returned WITHOUT a host sync: a query whose candidate
        lists overflowed carries idx[q, 0] == -2, which the merge propagates and `ShardedIndex.search` repairs after one
        check at the very end of the search."""
        if "comm" in shard.aux:
            return self._local_topk_exchanged(q, shard, k)
        from . import _lib
        db, aux, base = shard.db, shard.aux, shard.index_base
        nq, d = q.shape
        ndb = db.shape[0]
        if ndb == 0:
            return self._empty_lists(nq, k, q.device)
        use_tc = "shadow" in aux and k <= 1024 and nq * ndb * d >= TC_MIN_WORK and base + ndb <= 0xFFFFFFFF
        if not use_tc:
            return self.exact_topk(q, shard, k)
        s, i, st = _lib.score_topk(q, db, aux["shadow"], aux["norm_max"], k, index_base=base)
        status = st.cpu()  # one small sync per search: the status word decides whether a repair pass is needed
        shard.last_status = status.tolist()
        if int(status[0]) != 0:
            bad = torch.nonzero(i[:, 0] == -2).flatten()
            if bad.numel():
                s[bad], i[bad] = self.repair_topk(q[bad].contiguous(), shard, k)
        return s, i

    def repair_topk(self, q, shard, k):
        """Exact local lists of queries whose candidate segments overflowed (dense near-ties, duplicated rows): first a
        retry with a much larger k -- that multiplies the candidate and survivor capacities and spreads the shard over
        more stripes -- and only what still overflows goes to the exact CUDA-core scan. No collective in here."""
        from . import _lib
        db, aux, base = shard.db, shard.aux, shard.index_base
        if db.shape[0] == 0:
            return self._empty_lists(q.shape[0], k, q.device)
        kk = 1024              # 8192-entry candidate segments per stripe, 8192 survivors
        if kk > k and "shadow" in aux:
            s2, i2, _ = _lib.score_topk(q, db, aux["shadow"], aux["norm_max"], kk, index_base=base)
            s, i = s2[:, :k].contiguous(), i2[:, :k].contiguous()
            bad = torch.nonzero(i2[:, 0] == -2).flatten()
            if bad.numel():
                s[bad], i[bad] = self.exact_topk(q[bad].contiguous(), shard, k)
            return s, i
        return self.exact_topk(q, shard, k)

    def _local_topk_exchanged(self, q, shard, k):
        """Row-sharded search with the histogram exchange (include/gandtr_b200.h, two-phase form). Every rank takes the
        same code path (tcgen05 vs exact is decided from global quantities) so the collectives line up."""
        from . import _lib
        db, aux, base = shard.db, shard.aux, shard.index_base
        nq, d = q.shape
        comm = aux["comm"]
        ndb = db.shape[0]
        use_tc = k <= 1024 and nq * shard.n_total * d >= TC_MIN_WORK * comm.world and shard.n_total <= 0xFFFFFFFF
        if not use_tc:
            return self.exact_topk(q, shard, k)
        if ndb == 0:
            hist = torch.zeros((nq, 256), dtype=torch.int32, device=q.device)
            comm.all_reduce_sum(hist)                    # take part in the histogram exchange with an empty contribution
            return self._empty_lists(nq, k, q.device)
        state = _lib.score_topk_filter(q, aux["shadow"], aux["norm_max"], k)
        comm.all_reduce_sum(state.hist)
        s, i, st = _lib.score_topk_finalize(q, db, state, index_base=base)
        shard.last_status = st                           # device tensor: read lazily (bench / diagnostics), no sync here
        return s, i

    def exact_topk(self, q, shard, k):
        """CUDA-core exact lists of one shard (fallback and small problems). The database is streamed in row blocks so the
        score workspace stays bounded (EXACT_WS_FLOATS), block lists are merged on the device."""
        from . import _lib
        db, base = shard.db, shard.index_base
        nq, ndb = q.shape[0], db.shape[0]
        if ndb == 0:
            return self._empty_lists(nq, k, q.device)
        if k > EXACT_MAX_K:
            return self._full_sort_topk(q, db, k, base)
        rows = min(ndb, max(4 * k, 1 << 20))
        qstep = max(1, min(nq, EXACT_WS_FLOATS // rows))
        out_s, out_i = [], []
        for a in range(0, nq, qstep):
            qa = q[a:a + qstep].contiguous()
            parts = [_lib.score_topk_exact(qa, db[r0:r0 + rows], k, index_base=base + r0) for r0 in range(0, ndb, rows)]
            if len(parts) == 1:
                s, i = parts[0]
            else:
                s, i = self.merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
            out_s.append(s)
            out_i.append(i)
        return (out_s[0], out_i[0]) if len(out_s) == 1 else (torch.cat(out_s), torch.cat(out_i))

    @staticmethod
    def _full_sort_topk(q, db, k, base):
        """k > EXACT_MAX_K (deep mining walks): outside the hot path -- fp64 library GEMM rounded once to fp32, stable
        descending sort (ties -> lower index first), query blocks bounded like the exact kernel's workspace."""
        nq, ndb = q.shape[0], db.shape[0]
        kk = min(k, ndb)
        s = torch.full((nq, k), float("-inf"), dtype=torch.float32, device=q.device)
        i = torch.full((nq, k), -1, dtype=torch.int64, device=q.device)
        step = max(1, EXACT_WS_FLOATS // (2 * max(ndb, 1)))
        dbd = db.double()
        for a in range(0, nq, step):
            sc = (q[a:a + step].double() @ dbd.t()).float()
            v, o = torch.sort(sc, dim=1, descending=True, stable=True)
            s[a:a + step, :kk], i[a:a + step, :kk] = v[:, :kk], o[:, :kk] + base
        return s, i

    def merge(self, scores, idx):
        """[g, nq, k] lists -> [nq, k]; more than MERGE_MAX_ENTRIES entries per query are merged in rounds."""
        from . import _lib
        scores, idx = scores.contiguous(), idx.contiguous()
        g, nq, k = scores.shape
        if k > MERGE_MAX_ENTRIES // 2:
            raise _lib.GdtError("top-k lists longer than %d entries cannot be merged on the device" % (MERGE_MAX_ENTRIES // 2))
        gmax = max(2, MERGE_MAX_ENTRIES // k)
        while g > gmax:
            parts = [_lib.topk_merge(scores[a:a + gmax].contiguous(), idx[a:a + gmax].contiguous()) for a in range(0, g, gmax)]
            scores, idx = torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts])
            g = scores.shape[0]
        return _lib.topk_merge(scores, idx)

    def pack(self, scores, idx):
        from . import _lib
        return _lib.topk_pack(scores.contiguous(), idx.contiguous())

    def merge_packed_keys(self, keys):
        from . import _lib
        return _lib.topk_merge_packed_keys(keys.contiguous())

    def unpack(self, keys):
        from . import _lib
        return _lib.topk_unpack(keys.contiguous())

    def merge_packed(self, keys):
        from . import _lib
        g, nq, k = keys.shape
        if g * k > MERGE_MAX_ENTRIES:
            raise _lib.GdtError("packed merge of %d x %d entries exceeds %d" % (g, k, MERGE_MAX_ENTRIES))
        return _lib.topk_merge_packed(keys.contiguous())

    def probe_scores(self, q, shard, probe_idx, out):
        from . import _lib
        if shard.db.shape[0]:
            _lib.probe_scores(q, shard.db, probe_idx, index_base=shard.index_base, out=out)
        return out

    def rank_counts(self, q, shard, probe_idx, probe_score, out):
        from . import _lib
        if shard.db.shape[0]:
            _lib.rank_counts(q, shard.db, probe_idx, probe_score, index_base=shard.index_base, out=out)
        return out

    def map_eval(self, pos_rank, junk_rank, npos, njunk, kappas, nres=None):
        from . import _lib
        return _lib.map_eval(pos_rank, junk_rank, npos, njunk, kappas, nres=nres)


class DatabaseShard:
    """Rows [index_base, index_base + n) of the database, resident in HBM: fp32 rows + the fp16 shadow used by the
    tcgen05 coarse pass (6 bytes per element in total)."""

    def __init__(self, db, index_base=0, ops=None, comm=None, n_total=None):
        self.ops = ops or CudaOps()
        self.db = db.contiguous()
        self.index_base = int(index_base)
        self.n_total = int(n_total if n_total is not None else db.shape[0])
        try:
            self.aux = self.ops.prepare(self.db, comm=comm)
        except TypeError:          # injected ops with the single-argument signature (tests)
            self.aux = self.ops.prepare(self.db)
        self._last_status = None

    @property
    def rows(self):
        return self.db.shape[0]

    @property
    def last_status(self):
        """[overflow status, max survivors re-scored per query, overflowed queries, max candidates per query] of the last
        tcgen05 search as a list (reads the device words on demand)."""
        st = self._last_status
        return st.tolist() if isinstance(st, torch.Tensor) else st

    @last_status.setter
    def last_status(self, value):
        self._last_status = value


class ShardedIndex:
    """Row-sharded database over a process group. `search` is the drop-in for
    `np.argsort(-np.dot(vecs.T, qvecs), axis=0)[:k]` (cirscore.py:71-72), returned query-major."""

    def __init__(self, local_db, n_total=None, group=None, ops=None, index_base=None, comm=None):
        self.comm = comm if comm is not None else DistComm(group)
        self.group = group
        self.world, self.rank = self.comm.world, self.comm.rank
        self.ops = ops or CudaOps()
        if n_total is None:
            n_total = self.comm.sum_int(local_db.shape[0])
        self.n_total = int(n_total)
        if index_base is None:
            index_base = shard_bounds(self.n_total, self.world, self.rank)[0]
            if self.world > 1:
                assert local_db.shape[0] == shard_bounds(self.n_total, self.world, self.rank)[1] - index_base, \
                    "local shard does not match shard_bounds(); pass index_base explicitly for custom partitions"
        self.shard = DatabaseShard(local_db, index_base, self.ops, comm=self.comm, n_total=self.n_total)

    @classmethod
    def from_full(cls, db, group=None, ops=None, comm=None):
        """Every rank holds (or can produce) the full [ndb, d] matrix: keep only this rank's rows."""
        comm = comm if comm is not None else DistComm(group)
        lo, hi = shard_bounds(db.shape[0], comm.world, comm.rank)
        return cls(db[lo:hi].contiguous(), n_total=db.shape[0], group=group, ops=ops, index_base=lo, comm=comm)

    def broadcast_queries(self, q, src=0):
        return self.comm.broadcast(q, src=src)

    def _exchange_and_merge(self, s, i):
        """Per-shard lists -> the same merged [nq, k] lists on every rank. With the product's ops the lists travel packed
        (8 bytes per entry, ONE all-gather); injected test ops without `pack` use the two-tensor form."""
        if hasattr(self.ops, "pack") and s.shape[1] * self.world <= MERGE_MAX_ENTRIES:
            keys = self.ops.pack(s, i)
            merged = self._merge_query_sharded(keys) if hasattr(self.ops, "merge_packed_keys") else None
            return merged if merged is not None else self.ops.merge_packed(self.comm.all_gather(keys))
        return self.ops.merge(self.comm.all_gather(s), self.comm.all_gather(i))

    def _merge_query_sharded(self, keys):
        """Query-sharded merge: rank r merges the lists of queries [r * m, (r + 1) * m) only. The per-shard lists reach it by
        ONE all-to-all (1 / world of an all-gather's bytes), the merged slices travel packed in ONE all-gather of
        nq * k * 8 bytes in total, and every rank unpacks the same [nq, k] result. None when the communicator has no
        all-to-all or there are fewer queries than ranks."""
        nq, k = keys.shape
        w = self.world
        if w == 1 or nq < w or not hasattr(self.comm, "all_to_all"):
            return None
        m = -(-nq // w)
        if m * w != nq:                                   # pad with empty lists (key 0 = padding)
            keys = torch.cat([keys, torch.zeros((m * w - nq, k), dtype=keys.dtype, device=keys.device)])
        recv = self.comm.all_to_all(keys.view(w, m, k))   # [source rank][my m queries][k]
        if recv is None:
            return None
        mine = self.ops.merge_packed_keys(recv)           # [m, k] merged keys of my query slice
        s, i = self.ops.unpack(self.comm.all_gather(mine).view(w * m, k))
        return s[:nq], i[:nq]

    def search(self, q, k, broadcast=False):
        """q: [nq, d] float32 (identical on every rank, or rank 0's copy with broadcast=True).
        Returns (scores [nq, k], global indices [nq, k] int64) on every rank.
        Sharded: filter -> histogram all-reduce -> finalize -> pack -> all-gather -> merge are enqueued back to back; the
        only host read is the overflow check on the MERGED lists at the end (identical on every rank, so every rank takes
        the same repair decision)."""
        if broadcast:
            q = self.broadcast_queries(q)
        s, i = self.ops.local_topk(q, self.shard, k)
        if self.world == 1:
            return s, i
        s, i = self._exchange_and_merge(s, i)
        if hasattr(self.ops, "repair_topk"):
            bad = i[:, 0] == -2
            if bool(bad.any()):                      # the search's single host sync
                rows = torch.nonzero(bad).flatten()
                rs, ri = self.ops.repair_topk(q[rows].contiguous(), self.shard, k)     # exact local lists, no collective
                rs, ri = self._exchange_and_merge(rs, ri)
                s[rows], i[rows] = rs, ri
        return s, i

    def search_from_host(self, q_host, k):
        """The same search for queries that still sit in (pinned) host memory, identical on every rank: instead of every
        rank pulling the whole [nq, d] matrix through the box's shared host links (world copies of the same bytes), rank r
        uploads rows [r * m, (r + 1) * m) only and ONE all-gather over NVLink completes the matrix on every device."""
        dev = self.shard.db.device
        w = self.world
        if w == 1:
            return self.search(q_host.to(dev, non_blocking=True), k)
        nq, d = q_host.shape
        m = -(-nq // w)
        lo, hi = min(self.rank * m, nq), min((self.rank + 1) * m, nq)
        mine = torch.zeros((m, d), dtype=q_host.dtype, device=dev)
        if hi > lo:
            mine[:hi - lo].copy_(q_host[lo:hi], non_blocking=True)
        full = self.comm.all_gather(mine).view(w * m, d)[:nq]
        return self.search(full.contiguous() if not full.is_contiguous() else full, k)

    def positions(self, q, probe_idx):
        """0-based position every probe id would take in the full descending ranking of its query
        (= what `np.flatnonzero(np.in1d(ranks[:, i], ids))` reads, evaluate.py:75-76). probe_idx: [nq, pmax] int64, -1 pad."""
        ps = torch.zeros(probe_idx.shape, dtype=torch.float32, device=q.device)
        self.ops.probe_scores(q, self.shard, probe_idx, ps)
        self.comm.all_reduce_sum(ps)                    # non-owners contributed exact zeros
        before = torch.zeros(probe_idx.shape, dtype=torch.int64, device=q.device)
        self.ops.rank_counts(q, self.shard, probe_idx, ps, before)
        self.comm.all_reduce_sum(before)
        return before


def _pad_rows(lists, fill=-1):
    n = max(1, max((len(x) for x in lists), default=1))
    arr = np.full((len(lists), n), fill, dtype=np.int64)
    for r, x in enumerate(lists):
        arr[r, :len(x)] = np.asarray(x, dtype=np.int64)
    return arr


def _to_device(arr, device):
    t = torch.from_numpy(arr)
    if torch.device(device).type == "cuda":
        t = t.pin_memory().to(device, non_blocking=True)
    return t


class PreparedProtocols:
    """The ground truth of a dataset turned, ONCE, into what the device needs: the probe ids per query, the columns of
    every protocol's positives / junk in the probe matrix, and the recall denominators. A dataset's ground truth does
    not change between evaluations (validation every few epochs), so the set logic of evaluate.py:75-80 (`np.in1d`,
    duplicated / foreign ids) need not be redone per call; `evaluate()` is then a handful of launches and one read-back.
    groups: per query {name: ids}; protocols: {protocol: (ok group names, junk group names)}."""

    def __init__(self, groups, protocols, n_total, device):
        names = sorted({n for ok, jk in protocols.values() for n in tuple(ok) + tuple(jk)})
        self.nq = nq = len(groups)
        self.n_total = int(n_total)
        raw = [{n: np.asarray(g.get(n, []), dtype=np.int64).reshape(-1) for n in names} for g in groups]
        # one probe column per distinct in-range id of a query, whatever groups it appears in
        ids = []
        for r in raw:
            allv = np.concatenate([r[n] for n in names]) if names else np.zeros(0, np.int64)
            ids.append(np.unique(allv[(allv >= 0) & (allv < self.n_total)]))
        self.prots = list(protocols.keys())
        ok_cols, junk_cols, nres = [], [], []

        def cols(r, uniq, group_names):
            given = np.concatenate([r[n] for n in group_names]) if group_names else np.zeros(0, np.int64)
            found = np.unique(given[(given >= 0) & (given < self.n_total)])       # np.in1d: a set, foreign ids never found
            return np.searchsorted(uniq, found), len(given)
        for prot in self.prots:
            ok_names, junk_names = protocols[prot]
            for r, uniq in zip(raw, ids):
                oc, n_given = cols(r, uniq, ok_names)
                jc, _ = cols(r, uniq, junk_names)
                ok_cols.append(oc)
                junk_cols.append(jc)
                nres.append(n_given)
        self.nres = nres
        counts = np.stack([[len(c) for c in ok_cols], [len(c) for c in junk_cols], nres]).astype(np.int32)
        self.probe_idx = _to_device(_pad_rows(ids), device)                       # [nq, pu]
        self.okm = _to_device(_pad_rows(ok_cols, fill=0), device)                 # [nprot * nq, max positives]
        self.jkm = _to_device(_pad_rows(junk_cols, fill=0), device)
        cnt = _to_device(counts, device)
        self.npos, self.njunk, self.nres_dev = cnt[0].contiguous(), cnt[1].contiguous(), cnt[2].contiguous()
        self._graphs, self._kappa_dev, self._graph_failed, self._evals = {}, {}, False, 0

    def _device_part(self, index, q, kappas):
        """Everything of an evaluation that runs on the device -> [nprot * nq, 1 + max(nk, 1)] float64 (AP, P@k)."""
        before = index.positions(q, self.probe_idx)                               # [nq, pu]
        rep = before.repeat(len(self.prots), 1) if len(self.prots) > 1 else before
        pos_rank = torch.gather(rep, 1, self.okm).contiguous()
        junk_rank = torch.gather(rep, 1, self.jkm).contiguous()
        try:
            ap, prk = index.ops.map_eval(pos_rank, junk_rank, self.npos, self.njunk, kappas, nres=self.nres_dev)
        except TypeError:                  # injected test ops with the five-argument signature
            ap, prk = index.ops.map_eval(pos_rank, junk_rank, self.npos, self.njunk, kappas)
        return torch.cat([ap.reshape(-1, 1), prk.reshape(ap.shape[0], -1)], dim=1)

    def _graphed(self, index, q, kap_dev, key):
        """An evaluation is ~25 launches of a few microseconds each (latency-bound: 1.1 ms, of which the kernels take
        0.7): on a single-process index the whole device part is captured ONCE per (index, query shape, kappas) into a
        CUDA graph and replayed with the queries copied into its static input. None = not applicable (sharded index:
        collectives in the middle; injected test ops; capture failed once)."""
        if (index.world != 1 or not isinstance(index.ops, CudaOps) or not q.is_cuda or self._graph_failed
                or os.environ.get("GANDTR_B200_NO_GRAPH")):
            return None
        self._evals += 1
        if self._evals < 2:                 # a one-shot evaluation (ground truth prepared per call) is not worth a capture
            return None
        ent = self._graphs.get(key)
        if ent is not None and (ent["index"]() is not index or ent["db_ptr"] != index.shard.db.data_ptr()):
            ent = None
        if ent is None:
            try:
                static_q = q.clone()
                side = torch.cuda.Stream(device=q.device)
                side.wait_stream(torch.cuda.current_stream(q.device))
                with torch.cuda.stream(side):              # warm-up outside the capture: lazy attributes, allocator pools
                    for _ in range(2):
                        self._device_part(index, static_q, kap_dev)
                torch.cuda.current_stream(q.device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._device_part(index, static_q, kap_dev)
                ent = {"graph": graph, "q": static_q, "out": out, "index": weakref.ref(index), "db_ptr": index.shard.db.data_ptr()}
                if len(self._graphs) >= 4:
                    self._graphs.clear()
                self._graphs[key] = ent
            except Exception:                              # noqa: BLE001 -- an optimisation only: fall back to eager launches
                self._graph_failed = True
                torch.cuda.synchronize(q.device)
                return None
        ent["q"].copy_(q)
        ent["graph"].replay()
        return ent["out"]

    def evaluate(self, index, q, kappas=()):
        """-> {protocol: (map, aps, mean P@k, P@k)}: ONE pass over the database, ONE evaluation launch, ONE read-back."""
        assert index.n_total == self.n_total, "ground truth was prepared for another database size"
        nq, nk = self.nq, len(kappas)
        kappas = tuple(int(k) for k in kappas)
        both = None
        if q.is_cuda and isinstance(index.ops, CudaOps):
            kap = self._kappa_dev.get(kappas)          # device copy made once: no host -> device copy inside a capture
            if kap is None:
                kap = self._kappa_dev[kappas] = torch.tensor(list(kappas), dtype=torch.int32, device=q.device)
            both = self._graphed(index, q, kap, (id(index), tuple(q.shape), kappas))
            if both is None:
                both = self._device_part(index, q, kap)
        else:
            both = self._device_part(index, q, list(kappas))
        both = both.cpu().numpy()
        ap, prk = both[:, 0], both[:, 1:1 + nk]
        out = {}
        for pi, prot in enumerate(self.prots):
            a, pk = ap[pi * nq:(pi + 1) * nq], prk[pi * nq:(pi + 1) * nq]
            # same accumulation order as the reference loop (evaluate.py:98,106,108-109): sequential, empty queries skipped
            total, pr, nempty = 0.0, np.zeros(nk), 0
            for i in range(nq):
                if self.nres[pi * nq + i] == 0:
                    nempty += 1
                    continue
                total = total + a[i]
                pr = pr + pk[i, :]
            nvalid = nq - nempty
            out[prot] = ((total / nvalid if nvalid else float("nan")), a, (pr / nvalid if nvalid else pr * np.nan), pk)
        return out


def evaluate_protocols(index, q, groups, protocols, kappas=()):
    """Several (ok, junk) partitions of the same id groups with ONE pass over the database and ONE evaluation launch.
    groups: per query {name: ids}; protocols: {protocol: (ok group names, junk group names)}.
    Returns {protocol: (map, aps, mean P@k, P@k)}.
    Ids follow `np.in1d(ranks[:, i], ids)` (evaluate.py:75-76): set semantics -- a duplicated id is found once, an id
    outside the database never -- while the recall step stays 1 / len(list as given) (evaluate.py:77-78)."""
    return PreparedProtocols(groups, protocols, index.n_total, q.device).evaluate(index, q, kappas)


def evaluate_map(index, q, gnd, kappas=()):
    """compute_map (evaluate.py:39-111) on the GPU. gnd: list of {'ok': ids, 'junk': ids}.
    Returns (map, aps [nq] float64, mean P@k, P@k [nq, nk]) as NumPy, NaN rows for queries without positives."""
    groups = [{"ok": g["ok"], "junk": g.get("junk", [])} for g in gnd]
    return evaluate_protocols(index, q, groups, {"map": (("ok",), ("junk",))}, kappas)["map"]


REVISITED_PROTOCOLS = {"easy": (("easy",), ("junk", "hard")),          # evaluate.py:125-131
                       "medium": (("easy", "hard"), ("junk",)),         # :133-139
                       "hard": (("hard",), ("junk", "easy"))}           # :141-147


class PreparedGroundTruth:
    """`gnd` of `compute_map_and_print(dataset, ranks, gnd)` prepared once for repeated evaluations of the same dataset."""

    def __init__(self, dataset, gnd, n_total, device):
        self.old = "ok" in gnd[0]                                      # old protocol (Oxford/Paris/Tokyo), :117-120
        if self.old:
            groups = [{"ok": g["ok"], "junk": g.get("junk", [])} for g in gnd]
            self.prepared = PreparedProtocols(groups, {"map": (("ok",), ("junk",))}, n_total, device)
        else:
            if not (dataset.startswith("roxford5k") or dataset.startswith("rparis6k")):
                raise ValueError("Unsupported ground-truth format for dataset %s" % dataset)
            self.prepared = PreparedProtocols(gnd, REVISITED_PROTOCOLS, n_total, device)


def compute_map_and_print(dataset, index, q, gnd, kappas=(1, 5, 10), printer=print):
    """Same protocol split, rounding and output dictionaries as evaluate.py:114-152, with (index, q) in place of
    the precomputed `ranks` matrix. `gnd`: the reference's list of dicts, or a PreparedGroundTruth of it."""
    prep = gnd if isinstance(gnd, PreparedGroundTruth) else PreparedGroundTruth(dataset, gnd, index.n_total, q.device)
    if prep.old:
        m, aps, _, _ = prep.prepared.evaluate(index, q)["map"]
        printer(">> {}: mAP {:.2f}".format(dataset, np.around(m * 100, decimals=2)))
        return {"map": m}, {"ap": aps}
    out_avg, out_aps, mprs = {}, {}, {}
    res = prep.prepared.evaluate(index, q, kappas)
    for name in ("easy", "medium", "hard"):
        m, aps, mpr, _ = res[name]
        out_avg["map_" + name], out_aps["ap_" + name], mprs[name] = m, aps, mpr
    printer(">> {}: mAP E: {}, M: {}, H: {}".format(dataset, *[np.around(out_avg["map_" + n] * 100, decimals=2)
                                                               for n in ("easy", "medium", "hard")]))
    printer(">> {}: mP@k{} E: {}, M: {}, H: {}".format(dataset, list(kappas), *[np.around(mprs[n] * 100, decimals=2)
                                                                                   for n in ("easy", "medium", "hard")]))
    return out_avg, out_aps
