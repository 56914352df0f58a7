"""Host logic of the row-sharded retrieval path on CPU: world_size-2 gloo, with the oracle standing in for the
CUDA kernels (the product's CudaOps needs a GPU; here only the sharding / broadcast / all_gather / all_reduce /
packed query-sharded merge plumbing of gandtr_b200.retrieval is under test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retrieval_np as R
from tests.util import golden, unit_rows


class OracleOps:
    """Same interface as gandtr_b200.retrieval.CudaOps, NumPy oracle arithmetic."""

    def prepare(self, db):
        return {}

    def local_topk(self, q, shard, k):
        s, i = R.topk(R.scores_exact(q.numpy(), shard.db.numpy()), k, index_base=shard.index_base)
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge(self, scores, idx):
        g, nq, k = scores.shape
        s = scores.permute(1, 0, 2).reshape(nq, g * k).numpy()
        i = idx.permute(1, 0, 2).reshape(nq, g * k).numpy()
        key_i = np.where(i < 0, np.iinfo(np.int64).max, i)
        order = np.lexsort((key_i, -s), axis=1)[:, :k]
        return torch.from_numpy(np.take_along_axis(s, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))

    # the packed exchange format and the query-sharded merge (ShardedIndex._merge_query_sharded), restated in NumPy:
    # key = order-preserving bits of the fp32 score << 32 | ~index (larger sorts first), 0 = padding
    @staticmethod
    def _keys(s, i):
        s = np.ascontiguousarray(s, dtype=np.float32) + np.float32(0)
        u = s.view(np.uint32).astype(np.uint64)
        ob = np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xffffffff), u | np.uint64(0x80000000))
        key = (ob << np.uint64(32)) | (~i.astype(np.uint64) & np.uint64(0xffffffff))
        return np.where(i >= 0, key, np.uint64(0))

    def pack(self, scores, idx):
        return torch.from_numpy(self._keys(scores.numpy(), idx.numpy()).view(np.int64))

    def merge_packed_keys(self, keys):
        self.slice_merges = getattr(self, "slice_merges", 0) + 1
        g, nq, k = keys.shape
        u = keys.numpy().view(np.uint64).transpose(1, 0, 2).reshape(nq, g * k)
        return torch.from_numpy(np.sort(u, axis=1)[:, ::-1][:, :k].copy().view(np.int64))

    def merge_packed(self, keys):
        return self.unpack(self.merge_packed_keys(keys))

    def unpack(self, keys):
        u = keys.numpy().view(np.uint64)
        ob = (u >> np.uint64(32)).astype(np.uint32)
        bits = np.where(ob & np.uint32(0x80000000), ob & np.uint32(0x7fffffff), ~ob)
        s = np.where(u != 0, bits.view(np.float32), np.float32(-np.inf)).astype(np.float32)
        i = np.where(u != 0, (~u & np.uint64(0xffffffff)).astype(np.int64), np.int64(-1))
        return torch.from_numpy(s), torch.from_numpy(i)

    def probe_scores(self, q, shard, probe_idx, out):
        qn, dbn, pi = q.numpy().astype(np.float64), shard.db.numpy().astype(np.float64), probe_idx.numpy()
        for r in range(pi.shape[0]):
            for c in range(pi.shape[1]):
                loc = pi[r, c] - shard.index_base
                if pi[r, c] >= 0 and 0 <= loc < dbn.shape[0]:
                    out[r, c] = float(np.float32(qn[r] @ dbn[loc]))
        return out

    def rank_counts(self, q, shard, probe_idx, probe_score, out):
        s = R.scores_exact(q.numpy(), shard.db.numpy())
        ids = np.arange(s.shape[1]) + shard.index_base
        pi, ps = probe_idx.numpy(), probe_score.numpy()
        for r in range(pi.shape[0]):
            for c in range(pi.shape[1]):
                if pi[r, c] >= 0:
                    out[r, c] += int(((s[r] > ps[r, c]) | ((s[r] == ps[r, c]) & (ids < pi[r, c]))).sum())
        return out

    def map_eval(self, pos_rank, junk_rank, npos, njunk, kappas):
        nq = pos_rank.shape[0]
        ap = np.full(nq, np.nan)
        prk = np.full((nq, len(kappas)), np.nan)
        for i in range(nq):
            if npos[i] == 0:
                continue
            pos = np.sort(pos_rank[i, :npos[i]].numpy())
            junk = np.sort(junk_rank[i, :njunk[i]].numpy())
            pos = pos - np.searchsorted(junk, pos, side="left")
            ap[i] = R.compute_ap(pos, int(npos[i]))
            for j, kap in enumerate(kappas):
                kq = min(int(pos.max()) + 1, kap)
                prk[i, j] = ((pos + 1) <= kq).sum() / kq
        return torch.from_numpy(ap), torch.from_numpy(prk)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gandtr_b200.retrieval import ShardedIndex, evaluate_map, shard_bounds
        g = golden("map_eval.npz")
        q, db = torch.from_numpy(g["q"]), torch.from_numpy(g["db"])
        index = ShardedIndex.from_full(db, ops=OracleOps())
        lo, hi = shard_bounds(db.shape[0], world, rank)
        assert index.shard.index_base == lo and index.shard.rows == hi - lo and index.n_total == db.shape[0]
        # queries live on rank 0 only and are broadcast
        qr = q.clone() if rank == 0 else torch.zeros_like(q)
        s, i = index.search(qr, 100, broadcast=True)
        os_, oi = R.topk(R.scores_exact(g["q"], g["db"]), 100)
        assert np.array_equal(i.numpy(), oi) and np.array_equal(s.numpy(), os_)
        # queries in host memory on every rank: each rank contributes its slice, one all-gather completes them
        s2, i2 = index.search_from_host(q, 100)
        assert torch.equal(i2, i) and torch.equal(s2, s)
        assert index.ops.slice_merges == 2                    # both searches went through the query-sharded packed merge
        # mAP (medium protocol) through the sharded positions + all_reduce
        ok = [np.concatenate([e[e >= 0], h[h >= 0]]) for e, h in zip(g["easy"], g["hard"])]
        junk = [x[x >= 0] for x in g["junk"]]
        m, aps, mpr, prs = evaluate_map(index, qr, [{"ok": o, "junk": j} for o, j in zip(ok, junk)], [1, 5, 10])
        np.testing.assert_allclose(aps, g["ap_medium"], rtol=0, atol=1e-6, equal_nan=True)
        assert abs(m - float(g["map_medium"])) < 1e-6
        np.testing.assert_allclose(mpr, g["mprM"], rtol=0, atol=1e-9)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_search_and_map_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_bounds_partition():
    from gandtr_b200.retrieval import shard_bounds
    for n, w in [(10, 3), (7, 8), (0, 2), (1000000, 8), (5, 5)]:
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_single_process_index_with_oracle_ops_ragged():
    from gandtr_b200.retrieval import ShardedIndex
    rs = np.random.RandomState(1)
    q, db = unit_rows(rs, 7, 32), unit_rows(rs, 45, 32)
    idx = ShardedIndex(torch.from_numpy(db), ops=OracleOps())
    s, i = idx.search(torch.from_numpy(q), 60)          # k > ndb: (-inf, -1) tail
    os_, oi = R.topk(R.scores_exact(q, db), 60)
    assert np.array_equal(i.numpy(), oi) and np.array_equal(s.numpy(), os_)
    assert (i.numpy()[:, 45:] == -1).all()
