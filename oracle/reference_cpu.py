"""ORACLE / CPU BASELINE (test infrastructure, never imported by gandtr_b200/).

The reference's own CPU implementation of the hot path, restated call for call with the SAME third-party
libraries the reference uses (OpenCV, torch CPU ops, NumPy/OpenBLAS), because the reference package itself
(/root/reference) cannot travel to the GPU box. This is what `bench.py --impl reference` and `cpu_baseline`
time; `tests/test_oracle_reference_cpu.py` pins it bit-for-bit to the golden fixtures produced by the
unmodified reference (tools/gen_golden.py) and to the NumPy oracle.

  transform_cv2        mdir/components/data/transform/core_transforms.py:76-100 (Pil2Numpy),
                       photometric_transforms.py:28-36 (ApplyClahe), functional.py:28-35,55-63,81-85,140-161,
                       core_transforms.py:35-70 (ToTensor, Normalize)
  gem_l2n_torch        mdir/external/cirtorch/layers/functional.py:21-22,130-131
  aggregate_torch      mdir/components/data/wrapper.py:235-245
  whiten_torch         mdir/components/data/wrapper.py:320-322
  rank_numpy           mdir/components/optim/score/cirscore.py:71-72
"""
import numpy as np

_clahe_cache = {}


def transform_cv2(img_u8, mean, std, clip_limit=1.0, grid=8):
    """u8 HWC RGB -> float32 CHW, `pil2np | apply_clahe:clip | totensor | normalize` exactly as the reference."""
    import cv2
    img = img_u8.astype(np.float32) / 255.0                                       # core_transforms.py:83
    spc = (cv2.cvtColor(img, cv2.COLOR_RGB2LAB) + np.array([0, 128, 128], dtype=np.float32)) / \
        np.array([100.0, 255.0, 255.0], dtype=np.float32)                         # functional.py:35
    key = (float(clip_limit), int(grid))
    if key not in _clahe_cache:
        _clahe_cache[key] = cv2.createCLAHE(clipLimit=clip_limit, tileGridSize=(int(grid), int(grid)))  # :145
    chan = spc[:, :, 0]
    spc[:, :, 0] = _clahe_cache[key].apply((chan * 255).astype(np.uint8)).astype(np.float32) / 255.0   # :148
    rgb = cv2.cvtColor((spc * np.array([100.0, 255.0, 255.0], dtype=np.float32)) -
                       np.array([0, 128, 128], dtype=np.float32), cv2.COLOR_LAB2RGB)                    # :63
    chw = np.ascontiguousarray(rgb.transpose(2, 0, 1))                             # core_transforms.py:35-44
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    chw -= m                                                                       # F.normalize: sub_ then div_
    chw /= s
    return chw


def gem_l2n_torch(fmap, p, eps=1e-6):
    """[n,c,h,w] torch CPU tensor -> [n,c]: LF.gem then LF.l2n."""
    import torch
    import torch.nn.functional as F
    x = F.avg_pool2d(fmap.clamp(min=eps).pow(p), (fmap.size(-2), fmap.size(-1))).pow(1.0 / p)
    x = x / (torch.norm(x, p=2, dim=1, keepdim=True) + eps).expand_as(x)
    return x.squeeze(-1).squeeze(-1)


def aggregate_torch(descs, msp):
    """list of [n,c] per-scale descriptors -> [n,c] (wrapper.py:238-243, per image)."""
    import torch
    v = torch.zeros_like(descs[0])
    for d in descs:
        v += d.pow(msp)
    v = (v / len(descs)).pow(1.0 / msp)
    return v / v.norm(dim=1, keepdim=True)


def whiten_torch(v, P, m, dimensions=None):
    """[n,c] -> [n,dim] (wrapper.py:320-322 applied per image)."""
    import torch
    dim = dimensions or P.shape[0]
    X = torch.mm(P[:dim, :], v.t() - m.reshape(-1, 1))
    X = X / (torch.norm(X, p=2, dim=0, keepdim=True) + 1e-6)
    return X.t().contiguous()


def rank_numpy(vecs_dxn, qvecs_dxq, k=None):
    """cirscore.py:71-72: scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)."""
    scores = np.dot(vecs_dxn.T, qvecs_dxq)
    ranks = np.argsort(-scores, axis=0)
    return (scores, ranks) if k is None else (scores, ranks[:k])
