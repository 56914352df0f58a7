"""Per-phase timing of the row-sharded search (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/search_breakdown.py [rows d nq]
Prints, per rank 0 and as the max over ranks, the CUDA-event time of: filter (seed + main tcgen05 pass), histogram
all-reduce, finalize (exact re-score), status read-back (host round trip), all_gather of the [nq, k] lists, merge.
Answers "where do the ~2.3 ms of fixed cost per search go" (lead 3)."""
import ctypes
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
    lib = _lib.load()
    ndb = db.shape[0]
    ws = torch.empty(lib.gdt_score_topk_workspace_bytes(nq, ndb, d, k), dtype=torch.uint8, device=dev)
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    status = torch.empty(4, dtype=torch.int32, device=dev)
    off, nbytes = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.check(lib.gdt_score_topk_exchange_layout(nq, ndb, d, k, ctypes.byref(off), ctypes.byref(nbytes)), "layout")
    hist = ws[off.value:off.value + nbytes.value].view(torch.int32).view(nq, 256)
    all_s = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
    all_i = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    names = ["filter", "hist_allreduce", "finalize", "status_readback", "all_gather", "merge", "total"]
    acc = {n: 0.0 for n in names}
    iters, warm = 8, 3
    for it in range(warm + iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        ev[0].record()
        _lib.check(lib.gdt_score_topk_filter(P(q), P(shadow), P(stats), nq, ndb, d, k, P(status), P(ws), ws.numel(), stream()), "filter")
        ev[1].record()
        if world > 1:
            dist.all_reduce(hist)
        ev[2].record()
        _lib.check(lib.gdt_score_topk_finalize(P(q), P(db), nq, ndb, d, k, lo, P(scores), P(idx), P(status), P(ws), ws.numel(), stream()), "finalize")
        ev[3].record()
        t0 = time.perf_counter()
        st = status.cpu()
        t_host = (time.perf_counter() - t0) * 1e3        # includes waiting for everything queued before it
        ev[4].record()
        if world > 1:
            dist.all_gather_into_tensor(all_s.view(-1), scores.view(-1))
            dist.all_gather_into_tensor(all_i.view(-1), idx.view(-1))
        else:
            all_s[0].copy_(scores)
            all_i[0].copy_(idx)
        ev[5].record()
        _lib.topk_merge(all_s, all_i)
        ev[6].record()
        torch.cuda.synchronize()
        if it >= warm:
            for n, (a, b) in zip(names[:6], [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6)]):
                acc[n] += ev[a].elapsed_time(ev[b]) / iters
            acc["total"] += ev[0].elapsed_time(ev[6]) / iters
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
