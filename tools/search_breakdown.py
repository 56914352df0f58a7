"""Per-phase timing of the row-sharded search (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/search_breakdown.py [rows d nq]
Prints, per rank 0 and as the max over ranks, the CUDA-event time of: filter (query preparation + seed + main tcgen05
pass), histogram all-reduce, finalize (exact re-score), pack, all_gather of the packed [nq, k] lists, merge, and the
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
    all_k = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
    names = ["filter", "hist_allreduce", "finalize", "pack", "all_gather", "merge", "flag_check", "total"]
    acc = {n: 0.0 for n in names}
    iters, warm = 8, 3
    for it in range(warm + iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        ev[0].record()
        state = _lib.score_topk_filter(q, shadow, stats, k)
        ev[1].record()
        if world > 1:
            dist.all_reduce(state.hist)
        ev[2].record()
        scores, idx, status = _lib.score_topk_finalize(q, db, state, index_base=lo)
        ev[3].record()
        keys = _lib.topk_pack(scores, idx)
        ev[4].record()
        if world > 1:
            dist.all_gather_into_tensor(all_k, keys)
        else:
            all_k[0].copy_(keys)
        ev[5].record()
        ms, mi = _lib.topk_merge_packed(all_k)
        ev[6].record()
        bad = bool((mi[:, 0] == -2).any())                # the search's single host read
        ev[7].record()
        torch.cuda.synchronize()
        st = status.cpu()
        if it >= warm:
            for n, (a, b) in zip(names[:7], [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7)]):
                acc[n] += ev[a].elapsed_time(ev[b]) / iters
            acc["total"] += ev[0].elapsed_time(ev[7]) / iters
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
