"""Host -> device copy ceiling of the box, nvbandwidth-style (one process per GPU under torchrun, or a single process):
every rank copies its own pinned 302 MB buffer (the e2e arm's per-step upload: 128 uint8 images of 1024x768) to its GPU,
all ranks at once, timed with CUDA events; prints GB/s per rank and the aggregate, with and without the NUMA binding the
bench applies (bench.bind_host_near_gpu). If the aggregate here equals what the e2e arm moves, the e2e arm is at the
box's host-memory / PCIe ceiling and not limited by this code.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist


def measure(dev, nbytes, iters=20):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dst.copy_(host, non_blocking=True)
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dst.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if dist.is_initialized():
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return nbytes / (ms * 1e-3) / 1e9


def main():
    from bench import bind_host_near_gpu
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 128 * 768 * 1024 * 3
    unbound = measure(dev, nbytes)
    info = bind_host_near_gpu(local)
    bound = measure(dev, nbytes)
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes,
                          "h2d_GBps_per_rank_slowest": {"default_placement": unbound, "numa_bound": bound},
                          "h2d_GBps_aggregate": {"default_placement": unbound * world, "numa_bound": bound * world},
                          "binding": infos}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
