"""Shared helpers of the test-suite: synthetic inputs and golden-fixture access."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LUT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gandtr_b200", "data", "rgb2lab_lut_s16.bin")
MEAN = [0.485, 0.456, 0.406]
STD = [0.229, 0.224, 0.225]


def golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def load_lut():
    return np.fromfile(LUT_PATH, dtype="<i2").reshape(33, 33, 33, 3)


def synth_image(seed, h, w, kind):
    rs = np.random.RandomState(seed)
    if kind == "noise":
        return rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([128 + 70 * (np.sin(xx / (37.0 + 5 * c)) + np.cos(yy / (23.0 + 3 * c))) for c in range(3)], -1)
    img = img + rs.normal(0, 8, img.shape)
    if kind == "dark":
        img = 255.0 * (np.clip(img, 0, 255) / 255.0) ** 2.8
    return np.clip(img, 0, 255).astype(np.uint8)


def unit_rows(rs, n, d):
    x = rs.normal(0, 1, (n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


class ThreadComm:
    """Several ranks of the row-sharded path emulated as THREADS of one process on ONE GPU: the collectives of
    gandtr_b200.retrieval.DistComm with a host-side barrier instead of NCCL. Every rank runs the product's real kernels
    (CudaOps) on its own CUDA stream; no kernel ever waits on another kernel, only host threads wait on each other, so a
    single-GPU box can execute the whole filter -> histogram exchange -> finalize -> pack -> gather -> merge sequence.

        shared = ThreadComm.Shared(world); comms = [ThreadComm(shared, r) for r in range(world)]
    """

    class Shared:
        def __init__(self, world):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world
        self.calls = {"all_reduce_sum": 0, "all_reduce_max": 0, "all_gather": 0, "broadcast": 0}

    def _exchange(self, t):
        import torch
        if t.is_cuda:
            torch.cuda.current_stream().synchronize()          # my contribution is complete
        self.shared.slots[self.rank] = t
        self.shared.barrier.wait()
        stacked = torch.stack([x.clone() for x in self.shared.slots])
        if t.is_cuda:
            torch.cuda.current_stream().synchronize()          # I am done reading the peers' tensors
        self.shared.barrier.wait()
        return stacked

    def all_reduce_sum(self, t):
        self.calls["all_reduce_sum"] += 1
        t.copy_(self._exchange(t).sum(0))
        return t

    def all_reduce_max(self, t):
        self.calls["all_reduce_max"] += 1
        t.copy_(self._exchange(t).max(0).values)
        return t

    def all_gather(self, t):
        self.calls["all_gather"] += 1
        return self._exchange(t.contiguous())

    def all_to_all(self, t):
        self.calls["all_to_all"] = self.calls.get("all_to_all", 0) + 1
        return self._exchange(t.contiguous())[:, self.rank].contiguous()       # chunk `rank` of every peer's tensor

    def broadcast(self, t, src=0):
        self.calls["broadcast"] += 1
        t.copy_(self._exchange(t)[src])
        return t

    def sum_int(self, v):
        import torch
        return int(self._exchange(torch.tensor([int(v)], dtype=torch.int64)).sum())


def run_ranks(world, fn):
    """Run fn(rank, comm) on `world` threads (each on its own CUDA stream); re-raises the first failure."""
    import threading
    import torch
    shared = ThreadComm.Shared(world)
    errors, results = [], [None] * world

    def body(rank):
        try:
            if torch.cuda.is_available():
                with torch.cuda.stream(torch.cuda.Stream()):
                    results[rank] = fn(rank, ThreadComm(shared, rank))
                    torch.cuda.current_stream().synchronize()
            else:
                results[rank] = fn(rank, ThreadComm(shared, rank))
        except BaseException as e:          # noqa: BLE001 -- a dead rank must not leave the others at the barrier
            errors.append(e)
            shared.barrier.abort()
    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
    if real or errors:
        raise (real or errors)[0]
    return results
