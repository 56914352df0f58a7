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
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "CudaOps", "DatabaseShard", "ShardedIndex", "evaluate_map", "evaluate_protocols", "compute_map_and_print",
           "TC_MIN_WORK"]

# below this many multiply-adds the exact CUDA-core kernel is used instead of the tcgen05 pipeline
TC_MIN_WORK = 1 << 24


def shard_bounds(n_total, world_size, rank):
    """Contiguous block partition of `n_total` rows: rank r owns [lo, hi). Sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class CudaOps:
    """The product's compute: every method is one or more launches of libgandtr_b200.so kernels."""

    def prepare(self, db, group=None, world=1):
        from . import _lib
        d = db.shape[1]
        if d % 8 != 0 or d > 8192:
            return {}
        if world > 1:
            # common scale / error bound on every rank, so that the per-query score histograms can be summed
            shadow, stats = _lib.db_prepare_sharded(db, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group))
            return {"shadow": shadow, "norm_max": stats, "group": group}
        if db.shape[0] > 0:
            shadow, norm_max = _lib.db_prepare(db)
            return {"shadow": shadow, "norm_max": norm_max}
        return {}

    def local_topk(self, q, shard, k):
        """(scores [nq,k], global idx [nq,k]) of one shard; exact ordering (score desc, index asc). On a sharded index
        the lists hold this shard's members of the GLOBAL top k (fewer than k valid entries, (-inf, -1) padding)."""
        if "group" in shard.aux:
            return self._local_topk_exchanged(q, shard, k)
        from . import _lib
        db, aux, base = shard.db, shard.aux, shard.index_base
        nq, d = q.shape
        ndb = db.shape[0]
        if ndb == 0:
            return (torch.full((nq, k), float("-inf"), dtype=torch.float32, device=q.device),
                    torch.full((nq, k), -1, dtype=torch.int64, device=q.device))
        use_tc = "shadow" in aux and k <= 1024 and nq * ndb * d >= TC_MIN_WORK and base + ndb <= 0xFFFFFFFF
        if not use_tc:
            return self._exact_chunked(q, db, k, base)
        s, i, st = _lib.score_topk(q, db, aux["shadow"], aux["norm_max"], k, index_base=base)
        status = st.cpu()  # one small sync per search: the status word decides whether a repair pass is needed
        shard.last_status = status.tolist()
        if int(status[0]) != 0:
            self._repair(q, shard, k, s, i)
        return s, i

    def _repair(self, q, shard, k, s, i):
        """Candidate overflow (dense near-ties, duplicated rows): first retry the flagged queries alone with a much larger
        k -- that multiplies the candidate and survivor capacities and spreads the shard over more stripes -- and only
        what still overflows goes to the exact CUDA-core scan."""
        from . import _lib
        db, aux, base = shard.db, shard.aux, shard.index_base
        bad = torch.nonzero(i[:, 0] == -2).flatten()
        if not bad.numel():
            return
        kk = 1024              # 8192-entry candidate segments per stripe, 8192 survivors
        if kk > k:
            s2, i2, st2 = _lib.score_topk(q[bad].contiguous(), db, aux["shadow"], aux["norm_max"], kk, index_base=base)
            ok = i2[:, 0] != -2
            s[bad[ok]] = s2[ok, :k]
            i[bad[ok]] = i2[ok, :k]
            bad = bad[~ok]
        if bad.numel():
            se, ie = self._exact_chunked(q[bad].contiguous(), db, k, base)
            s[bad] = se
            i[bad] = ie

    def _local_topk_exchanged(self, q, shard, k):
        """Row-sharded search with the histogram exchange (include/gandtr_b200.h, two-phase form). Every rank takes the
        same code path (tcgen05 vs exact is decided from global quantities) so the collectives line up."""
        from . import _lib
        db, aux, base = shard.db, shard.aux, shard.index_base
        nq, d = q.shape
        group = aux["group"]
        ndb = db.shape[0]
        use_tc = k <= 1024 and nq * shard.n_total * d >= TC_MIN_WORK * dist.get_world_size(group) and shard.n_total <= 0xFFFFFFFF
        if not use_tc:
            if ndb == 0:
                return (torch.full((nq, k), float("-inf"), dtype=torch.float32, device=q.device),
                        torch.full((nq, k), -1, dtype=torch.int64, device=q.device))
            return self._exact_chunked(q, db, k, base)
        if ndb == 0:
            hist = torch.zeros((nq, 256), dtype=torch.int32, device=q.device)
            dist.all_reduce(hist, group=group)          # take part in the histogram exchange with an empty contribution
            return (torch.full((nq, k), float("-inf"), dtype=torch.float32, device=q.device),
                    torch.full((nq, k), -1, dtype=torch.int64, device=q.device))
        s, i, st = _lib.score_topk_two_phase(q, db, aux["shadow"], aux["norm_max"], k, index_base=base,
                                             exchange=lambda h: dist.all_reduce(h, group=group))
        status = st.cpu()
        shard.last_status = status.tolist()
        if int(status[0]) != 0:
            self._repair(q, shard, k, s, i)      # local, no collective: the repaired lists are supersets of what is needed
        return s, i

    @staticmethod
    def _exact_chunked(q, db, k, base):
        from . import _lib
        # the exact kernel materialises [nq_chunk, ndb] scores in its workspace: bound it to ~1 GiB
        ndb = db.shape[0]
        step = max(1, min(q.shape[0], (1 << 28) // max(ndb, 1)))
        outs = [_lib.score_topk_exact(q[a:a + step].contiguous(), db, k, index_base=base)
                for a in range(0, q.shape[0], step)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    def merge(self, scores, idx):
        from . import _lib
        return _lib.topk_merge(scores.contiguous(), idx.contiguous())

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

    def map_eval(self, pos_rank, junk_rank, npos, njunk, kappas):
        from . import _lib
        return _lib.map_eval(pos_rank, junk_rank, npos, njunk, kappas)


class DatabaseShard:
    """Rows [index_base, index_base + n) of the database, resident in HBM: fp32 rows + the fp16 shadow used by the
    tcgen05 coarse pass (6 bytes per element in total)."""

    def __init__(self, db, index_base=0, ops=None, group=None, world=1, n_total=None):
        self.ops = ops or CudaOps()
        self.db = db.contiguous()
        self.index_base = int(index_base)
        self.n_total = int(n_total if n_total is not None else db.shape[0])
        try:
            self.aux = self.ops.prepare(self.db, group=group, world=world)
        except TypeError:          # injected ops with the single-argument signature (tests)
            self.aux = self.ops.prepare(self.db)
        self.last_status = None

    @property
    def rows(self):
        return self.db.shape[0]


class ShardedIndex:
    """Row-sharded database over a process group. `search` is the drop-in for
    `np.argsort(-np.dot(vecs.T, qvecs), axis=0)[:k]` (cirscore.py:71-72), returned query-major."""

    def __init__(self, local_db, n_total=None, group=None, ops=None, index_base=None):
        self.group = group
        self.world, self.rank = _world(group)
        self.ops = ops or CudaOps()
        if n_total is None:
            n_total = self._sum_int(local_db.shape[0])
        self.n_total = int(n_total)
        if index_base is None:
            index_base = shard_bounds(self.n_total, self.world, self.rank)[0]
            if self.world > 1:
                assert local_db.shape[0] == shard_bounds(self.n_total, self.world, self.rank)[1] - index_base, \
                    "local shard does not match shard_bounds(); pass index_base explicitly for custom partitions"
        self.shard = DatabaseShard(local_db, index_base, self.ops, group=group, world=self.world, n_total=self.n_total)

    @classmethod
    def from_full(cls, db, group=None, ops=None):
        """Every rank holds (or can produce) the full [ndb, d] matrix: keep only this rank's rows."""
        world, rank = _world(group)
        lo, hi = shard_bounds(db.shape[0], world, rank)
        return cls(db[lo:hi].contiguous(), n_total=db.shape[0], group=group, ops=ops, index_base=lo)

    def _sum_int(self, v):
        if self.world == 1:
            return int(v)
        t = torch.tensor([int(v)], dtype=torch.int64)
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def broadcast_queries(self, q, src=0):
        if self.world > 1:
            dist.broadcast(q, src=src, group=self.group)
        return q

    def search(self, q, k, broadcast=False):
        """q: [nq, d] float32 (identical on every rank, or rank 0's copy with broadcast=True).
        Returns (scores [nq, k], global indices [nq, k] int64) on every rank."""
        if broadcast:
            q = self.broadcast_queries(q)
        s, i = self.ops.local_topk(q, self.shard, k)
        if self.world == 1:
            return s, i
        all_s = torch.empty((self.world,) + tuple(s.shape), dtype=s.dtype, device=s.device)
        all_i = torch.empty((self.world,) + tuple(i.shape), dtype=i.dtype, device=i.device)
        # list-of-views form: one collective on NCCL (equal sizes), and also supported by gloo (CPU tests)
        dist.all_gather(list(all_s.unbind(0)), s.contiguous(), group=self.group)
        dist.all_gather(list(all_i.unbind(0)), i.contiguous(), group=self.group)
        return self.ops.merge(all_s, all_i)

    def positions(self, q, probe_idx):
        """0-based position every probe id would take in the full descending ranking of its query
        (= what `np.flatnonzero(np.in1d(ranks[:, i], ids))` reads, evaluate.py:75-76). probe_idx: [nq, pmax] int64, -1 pad."""
        ps = torch.zeros(probe_idx.shape, dtype=torch.float32, device=q.device)
        self.ops.probe_scores(q, self.shard, probe_idx, ps)
        if self.world > 1:
            dist.all_reduce(ps, group=self.group)       # non-owners contributed exact zeros
        before = torch.zeros(probe_idx.shape, dtype=torch.int64, device=q.device)
        self.ops.rank_counts(q, self.shard, probe_idx, ps, before)
        if self.world > 1:
            dist.all_reduce(before, group=self.group)
        return before


def _pad_ids(lists, device, fill=-1):
    n = max(1, max((len(x) for x in lists), default=1))
    arr = np.full((len(lists), n), fill, dtype=np.int64)
    for r, x in enumerate(lists):
        arr[r, :len(x)] = np.asarray(x, dtype=np.int64)
    return torch.from_numpy(arr).to(device)


def _map_from_positions(ops, before, ok_cols, junk_cols, kappas, device):
    """before: [nq, pu] positions of the probe ids; ok_cols / junk_cols: per query, the columns of `before` that hold its
    positives / junk ids. Runs gdt_map_eval and averages like the reference loop."""
    nq = len(ok_cols)
    npos_h = [len(c) for c in ok_cols]
    njunk_h = [len(c) for c in junk_cols]
    pos_rank = torch.gather(before, 1, _pad_ids(ok_cols, device, fill=0)).contiguous()
    junk_rank = torch.gather(before, 1, _pad_ids(junk_cols, device, fill=0)).contiguous()
    npos = torch.tensor(npos_h, dtype=torch.int32, device=device)
    njunk = torch.tensor(njunk_h, dtype=torch.int32, device=device)
    ap, prk = ops.map_eval(pos_rank, junk_rank, npos, njunk, list(kappas))
    ap, prk = ap.cpu().numpy(), prk.cpu().numpy()
    # same accumulation order as the reference loop (evaluate.py:98,106,108-109): sequential, empty queries skipped
    total, pr, nempty = 0.0, np.zeros(len(kappas)), 0
    for i in range(nq):
        if npos_h[i] == 0:
            nempty += 1
            continue
        total = total + ap[i]
        pr = pr + prk[i, :]
    nvalid = nq - nempty
    return (total / nvalid if nvalid else float("nan")), ap, (pr / nvalid if nvalid else pr * np.nan), prk


def evaluate_protocols(index, q, groups, protocols, kappas=()):
    """Several (ok, junk) partitions of the same id groups with ONE pass over the database.
    groups: per query {name: ids}; protocols: {protocol: (ok group names, junk group names)}.
    Returns {protocol: (map, aps, mean P@k, P@k)}."""
    names = sorted({n for ok, jk in protocols.values() for n in tuple(ok) + tuple(jk)})
    ids, spans = [], []
    for g in groups:
        off, span, parts = 0, {}, []
        for n in names:
            a = np.asarray(g.get(n, []), dtype=np.int64).reshape(-1)
            span[n] = (off, off + len(a))
            off += len(a)
            parts.append(a)
        ids.append(np.concatenate(parts) if parts else np.zeros(0, np.int64))
        spans.append(span)
    before = index.positions(q, _pad_ids(ids, q.device))
    out = {}
    for prot, (ok_names, junk_names) in protocols.items():
        ok_cols = [np.concatenate([np.arange(*sp[n]) for n in ok_names]) if ok_names else np.zeros(0, np.int64) for sp in spans]
        junk_cols = [np.concatenate([np.arange(*sp[n]) for n in junk_names]) if junk_names else np.zeros(0, np.int64) for sp in spans]
        out[prot] = _map_from_positions(index.ops, before, ok_cols, junk_cols, kappas, q.device)
    return out


def evaluate_map(index, q, gnd, kappas=()):
    """compute_map (evaluate.py:39-111) on the GPU. gnd: list of {'ok': ids, 'junk': ids}.
    Returns (map, aps [nq] float64, mean P@k, P@k [nq, nk]) as NumPy, NaN rows for queries without positives."""
    groups = [{"ok": g["ok"], "junk": g.get("junk", [])} for g in gnd]
    return evaluate_protocols(index, q, groups, {"map": (("ok",), ("junk",))}, kappas)["map"]


def compute_map_and_print(dataset, index, q, gnd, kappas=(1, 5, 10), printer=print):
    """Same protocol split, rounding and output dictionaries as evaluate.py:114-152, with (index, q) in place of
    the precomputed `ranks` matrix."""
    if "ok" in gnd[0]:                                            # old protocol (Oxford/Paris/Tokyo), :117-120
        m, aps, _, _ = evaluate_map(index, q, gnd)
        printer(">> {}: mAP {:.2f}".format(dataset, np.around(m * 100, decimals=2)))
        return {"map": m}, {"ap": aps}
    if not (dataset.startswith("roxford5k") or dataset.startswith("rparis6k")):
        raise ValueError("Unsupported ground-truth format for dataset %s" % dataset)
    out_avg, out_aps, mprs = {}, {}, {}
    res = evaluate_protocols(index, q, gnd, {"easy": (("easy",), ("junk", "hard")),          # :125-131
                                             "medium": (("easy", "hard"), ("junk",)),         # :133-139
                                             "hard": (("hard",), ("junk", "easy"))}, kappas)  # :141-147
    for name in ("easy", "medium", "hard"):
        m, aps, mpr, _ = res[name]
        out_avg["map_" + name], out_aps["ap_" + name], mprs[name] = m, aps, mpr
    printer(">> {}: mAP E: {}, M: {}, H: {}".format(dataset, *[np.around(out_avg["map_" + n] * 100, decimals=2)
                                                               for n in ("easy", "medium", "hard")]))
    printer(">> {}: mP@k{} E: {}, M: {}, H: {}".format(dataset, list(kappas), *[np.around(mprs[n] * 100, decimals=2)
                                                                                   for n in ("easy", "medium", "hard")]))
    return out_avg, out_aps
