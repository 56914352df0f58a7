"""Per-phase timing of the row-sharded search (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/search_breakdown.py [rows d nq]
Prints, per rank 0 and as the max over ranks, the CUDA-event time of: filter (query preparation + seed + main tcgen05
pass), histogram all-reduce, finalize (exact re-score), pack, all-to-all of the packed lists (query-sharded merge), merge
of this rank's query slice, all-gather of the merged slices, unpack, and the
overflow-flag read (the search's one host round trip). With one process it times one shard of the given size.
Answers "where do the ~2.3 ms of fixed cost per search go" (lead 3)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from bench import synth_db_rows
from gandtr_b200 import _lib
from gandtr_b200.retrieval import shard_bounds


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
    k = 100
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(rows, world, rank)
    db = synth_db_rows(lo, hi, d, dev)
    g = torch.Generator(device=dev).manual_seed(3)
    q = torch.randn((nq, d), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True)
    if world > 1:
        shadow, stats = _lib.db_prepare_sharded(db, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX))
    else:
        shadow, stats = _lib.db_prepare(db)
    ndb = db.shape[0]
    m = -(-nq // world)
    recv = torch.empty((world, m, k), dtype=torch.int64, device=dev)
    all_k = torch.empty((world, m, k), dtype=torch.int64, device=dev)
    names = ["filter", "hist_allreduce", "finalize", "pack", "all_to_all", "merge_slice", "all_gather", "unpack", "flag_check", "total"]
    acc = {n: 0.0 for n in names}
    iters, warm = 8, 3
    for it in range(warm + iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        ev[0].record()
        state = _lib.score_topk_filter(q, shadow, stats, k)
        ev[1].record()
        if world > 1:
            dist.all_reduce(state.hist)
        ev[2].record()
        scores, idx, status = _lib.score_topk_finalize(q, db, state, index_base=lo)
        ev[3].record()
        keys = _lib.topk_pack(scores, idx)
        if m * world != nq:
            keys = torch.cat([keys, torch.zeros((m * world - nq, k), dtype=keys.dtype, device=dev)])
        ev[4].record()
        if world > 1:
            dist.all_to_all_single(recv, keys.view(world, m, k))
        else:
            recv.copy_(keys.view(world, m, k))
        ev[5].record()
        mine = _lib.topk_merge_packed_keys(recv)
        ev[6].record()
        if world > 1:
            dist.all_gather_into_tensor(all_k, mine)
        else:
            all_k[0].copy_(mine)
        ev[7].record()
        ms, mi = _lib.topk_unpack(all_k.view(world * m, k))
        ev[8].record()
        bad = bool((mi[:nq, 0] == -2).any())              # the search's single host read
        ev[9].record()
        torch.cuda.synchronize()
        st = status.cpu()
        if it >= warm:
            for n, (a, b) in zip(names[:9], [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9)]):
                acc[n] += ev[a].elapsed_time(ev[b]) / iters
            acc["total"] += ev[0].elapsed_time(ev[9]) / iters
    t = torch.tensor([acc[n] for n in names], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("search breakdown: %d rows x %d, %d queries, k = %d, %d GPU(s); status %s" % (rows, d, nq, k, world, st.tolist()))
        for n, a, b in zip(names, t.tolist(), tmax.tolist()):
            print("  %-16s rank0 %8.3f ms   max over ranks %8.3f ms" % (n, a, b))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
