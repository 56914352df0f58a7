"""Sharded on-disk descriptor store -- SURVEY 8(f) row N4.

The reference keeps descriptors as an in-memory float64 `n x D` NumPy array (`EmbeddingOutput`,
mdir/components/data/output.py:118-156) and whitening as a `{'m', 'P'}` pickle (mdir/stages/whiten.py:75). For
databases of 10 M images that does not persist or shard. Layout here (one directory):

    manifest.json                     {"format": 1, "dim": D, "dtype": "float32" | "bfloat16", "rows": N,
                                       "shards": [{"file", "lo", "hi"}...], "whitening": "whitening.pkl" | null}
    shard_00000.npy ...               [hi - lo, D] row blocks (global rows lo..hi), np.load(mmap_mode='r')-able: float32, or
                                      the bfloat16 bit patterns as uint16 (half the bytes; round-to-nearest-even, i.e. ~3
                                      significant digits: enough to rank, NOT bit-exact -- the fp32 payload is the default)
    ids.txt                           optional: one image identifier per global row
    whitening.pkl                     optional: the reference's {'m': D x 1, 'P': D x D} pickle, unchanged

`load_shard(path, rank, world)` returns the rows `retrieval.shard_bounds(N, world, rank)` of the database as a float32 CUDA
tensor (reading only the files that overlap), ready for `ShardedIndex(rows, n_total=N, index_base=lo)`. The read is
pipelined: file blocks are copied into two pinned staging buffers in turn and uploaded on a copy stream, so the disk /
page-cache read of block i + 1 overlaps the PCIe transfer of block i.
"""
import json
import os
import pickle

import numpy as np
import torch

from .retrieval import shard_bounds

__all__ = ["save", "load_manifest", "load_shard", "load_whitening"]


def save(path, descriptors, ids=None, whitening=None, rows_per_file=1 << 20, lo=0, n_total=None, dtype="float32"):
    """Write rows [lo, lo + len(descriptors)) of an N-row database. Every rank of a sharded extraction calls this with its
    own block; the rank that holds row 0 also writes ids / whitening and the manifest (pass n_total)."""
    if dtype not in ("float32", "bfloat16"):
        raise ValueError("descriptor store payload must be float32 or bfloat16, got %r" % (dtype,))
    os.makedirs(path, exist_ok=True)
    t = descriptors.detach().to(torch.float32) if isinstance(descriptors, torch.Tensor) else \
        torch.from_numpy(np.asarray(descriptors, dtype=np.float32))
    if dtype == "bfloat16":
        x = t.to(torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16)      # round-to-nearest-even
    else:
        x = t.cpu().numpy()
    n, d = x.shape
    n_total = int(n_total if n_total is not None else lo + n)
    shards = []
    for a in range(0, n, rows_per_file):
        b = min(n, a + rows_per_file)
        name = "shard_%012d.npy" % (lo + a)
        np.save(os.path.join(path, name), x[a:b])
        shards.append({"file": name, "lo": lo + a, "hi": lo + b})
    with open(os.path.join(path, "part_%012d.json" % lo), "w") as f:
        json.dump(shards, f)
    if lo == 0:
        if ids is not None:
            with open(os.path.join(path, "ids.txt"), "w") as f:
                f.write("\n".join(str(i) for i in ids))
        if whitening is not None:
            with open(os.path.join(path, "whitening.pkl"), "wb") as f:
                pickle.dump({"m": np.asarray(whitening["m"]), "P": np.asarray(whitening["P"])}, f)
        with open(os.path.join(path, "manifest.json"), "w") as f:
            json.dump({"format": 1, "dim": d, "dtype": dtype, "rows": n_total,
                       "whitening": "whitening.pkl" if whitening is not None else None}, f)
    return shards


def load_manifest(path):
    with open(os.path.join(path, "manifest.json")) as f:
        man = json.load(f)
    shards = []
    for name in sorted(os.listdir(path)):
        if name.startswith("part_") and name.endswith(".json"):
            with open(os.path.join(path, name)) as f:
                shards += json.load(f)
    man["shards"] = sorted(shards, key=lambda s: s["lo"])
    covered = 0
    for s in man["shards"]:
        if s["lo"] != covered:
            raise ValueError("descriptor store %s: rows %d..%d are missing" % (path, covered, s["lo"]))
        covered = s["hi"]
    if covered != man["rows"]:
        raise ValueError("descriptor store %s: %d of %d rows present" % (path, covered, man["rows"]))
    return man


def load_shard(path, rank=0, world_size=1, device=None, block_rows=None):
    """-> (rows [hi - lo, D] float32 on `device`, lo, N). On a CUDA device the upload is double-buffered through pinned
    staging memory (block_rows rows per transfer, default ~32 MiB)."""
    man = load_manifest(path)
    lo, hi = shard_bounds(man["rows"], world_size, rank)
    device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
    bf16 = man.get("dtype", "float32") == "bfloat16"
    np_dtype, t_dtype = (np.uint16, torch.int16) if bf16 else (np.float32, torch.float32)
    dim = man["dim"]
    out = torch.empty((hi - lo, dim), dtype=torch.float32, device=device)
    pieces = []                                   # (file, first row in the file, first row in `out`, rows)
    step = int(block_rows or max(1, (32 << 20) // (dim * np.dtype(np_dtype).itemsize)))
    for s in man["shards"]:
        a, b = max(lo, s["lo"]), min(hi, s["hi"])
        for r in range(a, b, step):
            pieces.append((s["file"], r - s["lo"], r - lo, min(step, b - r)))
    if device.type != "cuda":
        for name, f0, o0, n in pieces:
            block = np.ascontiguousarray(np.load(os.path.join(path, name), mmap_mode="r")[f0:f0 + n])
            t = torch.from_numpy(block.view(np.int16) if bf16 else block)
            out[o0:o0 + n] = t.view(torch.bfloat16).to(torch.float32) if bf16 else t
        return out, lo, man["rows"]
    copy_stream = torch.cuda.Stream(device)
    copy_stream.wait_stream(torch.cuda.current_stream(device))     # `out` may be a recycled block still in use there
    stage = [torch.empty((step, dim), dtype=t_dtype).pin_memory() for _ in range(2)]
    done = [None, None]
    files = {}
    for j, (name, f0, o0, n) in enumerate(pieces):
        buf = stage[j % 2]
        if done[j % 2] is not None:
            done[j % 2].synchronize()             # the upload that last used this staging buffer has finished
        if name not in files:
            files[name] = np.load(os.path.join(path, name), mmap_mode="r")
        np.copyto(buf.numpy()[:n].view(np_dtype), files[name][f0:f0 + n])      # disk / page cache -> pinned memory
        with torch.cuda.stream(copy_stream):
            src = buf[:n].to(device, non_blocking=True)
            out[o0:o0 + n] = src.view(torch.bfloat16).to(torch.float32) if bf16 else src
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        done[j % 2] = ev
    torch.cuda.current_stream(device).wait_stream(copy_stream)
    return out, lo, man["rows"]


def load_whitening(path):
    man = load_manifest(path)
    if not man.get("whitening"):
        return None
    with open(os.path.join(path, man["whitening"]), "rb") as f:
        return pickle.load(f)
