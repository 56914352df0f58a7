"""Sharded on-disk descriptor store -- SURVEY 8(f) row N4.

The reference keeps descriptors as an in-memory float64 `n x D` NumPy array (`EmbeddingOutput`,
mdir/components/data/output.py:118-156) and whitening as a `{'m', 'P'}` pickle (mdir/stages/whiten.py:75). For
databases of 10 M images that does not persist or shard. Layout here (one directory):

    manifest.json                     {"format": 1, "dim": D, "dtype": "float32", "rows": N, "shards": [{"file", "lo", "hi"}...],
                                       "whitening": "whitening.pkl" | null}
    shard_00000.npy ...               float32 [hi - lo, D] row blocks (global rows lo..hi), np.load(mmap_mode='r')-able
    ids.txt                           optional: one image identifier per global row
    whitening.pkl                     optional: the reference's {'m': D x 1, 'P': D x D} pickle, unchanged

`load_shard(path, rank, world)` returns the rows `retrieval.shard_bounds(N, world, rank)` of the database as a CUDA tensor
(reading only the files that overlap), ready for `ShardedIndex(rows, n_total=N, index_base=lo)`.
"""
import json
import os
import pickle

import numpy as np
import torch

from .retrieval import shard_bounds

__all__ = ["save", "load_manifest", "load_shard", "load_whitening"]


def save(path, descriptors, ids=None, whitening=None, rows_per_file=1 << 20, lo=0, n_total=None):
    """Write rows [lo, lo + len(descriptors)) of an N-row database. Every rank of a sharded extraction calls this with its
    own block; the rank that holds row 0 also writes ids / whitening and the manifest (pass n_total)."""
    os.makedirs(path, exist_ok=True)
    x = descriptors.detach().to("cpu", torch.float32).numpy() if isinstance(descriptors, torch.Tensor) else \
        np.asarray(descriptors, dtype=np.float32)
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
            json.dump({"format": 1, "dim": d, "dtype": "float32", "rows": n_total,
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


def load_shard(path, rank=0, world_size=1, device=None):
    """-> (rows [hi - lo, D] float32 on `device`, lo, N)."""
    man = load_manifest(path)
    lo, hi = shard_bounds(man["rows"], world_size, rank)
    device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
    out = torch.empty((hi - lo, man["dim"]), dtype=torch.float32, device=device)
    for s in man["shards"]:
        a, b = max(lo, s["lo"]), min(hi, s["hi"])
        if a < b:
            block = np.load(os.path.join(path, s["file"]), mmap_mode="r")[a - s["lo"]:b - s["lo"]]
            out[a - lo:b - lo] = torch.from_numpy(np.ascontiguousarray(block)).to(device, non_blocking=False)
    return out, lo, man["rows"]


def load_whitening(path):
    man = load_manifest(path)
    if not man.get("whitening"):
        return None
    with open(os.path.join(path, man["whitening"]), "rb") as f:
        return pickle.load(f)
