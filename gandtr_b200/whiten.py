"""Learned whitening: apply and learn -- mirrors mdir/external/cirtorch/utils/whiten.py:4-70 and the stage functions of
mdir/stages/whiten.py:10-107. Float64 like the reference (it is not throughput-critical: it runs once and produces the
P, m consumed by K2). The O(D^2 n) work -- the three Gram products `df df^T`, `(P(X-m)) (P(X-m))^T`, `Xc Xc^T` over all
descriptors -- runs in this library's own fp64 kernel (gdt_syrk_f64: triangular tiles, fixed-order split-n reduction,
bit-reproducible); the O(D^3) factorisations (Cholesky, inverse, symmetric eigendecomposition) go through torch.linalg
(cuSOLVER) as SURVEY 8(f) N2 prescribes. Descriptors never leave the device. Matrices follow the reference layout: X is
D x n, m is D x 1, P is D x D.

Eigenvectors are defined up to sign: rows of P may differ from NumPy's `eig` by a factor -1, which leaves every inner
product between whitened vectors -- hence every ranking -- unchanged.
"""
import sys
import time

import numpy as np
import torch

__all__ = ["whitenapply", "whitenlearn", "pcawhitenlearn", "cholesky", "whiten", "learn_lw_whitening", "learn_pca_whitening"]


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("gandtr_b200 has no CPU path: whitening needs a CUDA device")
    return torch.device("cuda", torch.cuda.current_device())


def _t(x):
    return x.to(torch.float64) if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=torch.float64, device=_dev())


def _gram(A, alpha=1.0):
    """alpha * A A^T for a D x n float64 matrix: gdt_syrk_f64 (symmetric by construction)."""
    from . import _lib
    return _lib.syrk_f64(A if A.stride(1) == 1 else A.contiguous(), alpha)


def whitenapply(X, m, P, dimensions=None):
    X, m, P = _t(X), _t(m), _t(P)
    if not dimensions:
        dimensions = P.shape[0]
    X = P[:dimensions, :] @ (X - m)
    return X / (torch.linalg.norm(X, ord=2, dim=0, keepdim=True) + 1e-6)


def cholesky(S):
    """Cholesky with diagonal loading until positive definite (whiten.py:52-70)."""
    alpha = 0
    eye = torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
    while 1:
        L, info = torch.linalg.cholesky_ex(S + alpha * eye)
        if int(info) == 0:
            return L
        alpha = 1e-10 if alpha == 0 else alpha * 10
        print(">>>> whiten.py::cholesky: Matrix is not positive definite, adding {:.0e} on the diagonal".format(alpha))


def _eig_desc(D):
    eigval, eigvec = torch.linalg.eigh((D + D.t()) / 2)
    order = torch.argsort(eigval, descending=True)
    return eigval[order], eigvec[:, order]


def whitenlearn(X, qidxs, pidxs):
    """Supervised whitening from (query, positive) pairs (whiten.py:37-50)."""
    X = _t(X)
    qidxs = torch.as_tensor(np.asarray(qidxs), dtype=torch.long, device=X.device)
    pidxs = torch.as_tensor(np.asarray(pidxs), dtype=torch.long, device=X.device)
    m = X[:, qidxs].mean(dim=1, keepdim=True)
    df = X[:, qidxs] - X[:, pidxs]
    S = _gram(df, 1.0 / df.shape[1])
    P = torch.linalg.inv(cholesky(S))
    df = P @ (X - m)
    D = _gram(df)
    _, eigvec = _eig_desc(D)
    return m, eigvec.t() @ P


def pcawhitenlearn(X, shrink=None):
    X = _t(X)
    N = X.shape[1]
    m = X.mean(dim=1, keepdim=True)
    Xc = X - m
    Xcov = _gram(Xc, 1.0 / N)                  # symmetric by construction: (Xcov + Xcov.T) / (2 N) of the reference
    eigval, eigvec = _eig_desc(Xcov)
    if shrink:
        b = eigval[shrink - 1]
        eigval = (1 - b) * eigval + b
    return m, torch.diag(1.0 / torch.sqrt(eigval)) @ eigvec.t()


# ---- stage functions: f(params, data) -> (metadata, *data)  (mdir/examples/perform_scenario.py:126-130) ----

def whiten(params, data):
    """Apply pre-computed whitening (stages/whiten.py:10-27). values: n x D."""
    dimensions = params.pop("dimensions", None) or None
    assert not params, params.keys()
    whitening, names, values = data
    assert len(names) == len(values)
    if not whitening:
        return {"status": "No whitening applied"}, names, values
    time0 = time.time()
    whitened = whitenapply(_t(values).t(), whitening["m"], whitening["P"], dimensions)
    out = whitened.t().contiguous()
    return {"timings": {"whitening_apply": round(time.time() - time0, 2)}}, names, (out if isinstance(values, torch.Tensor) else out.cpu().numpy())


def learn_lw_whitening(params, data):
    """stages/whiten.py:30-75, incl. the retry on shrinking random subsets when the matrix is not positive definite."""
    assert not params
    names, values, queries, positives = data
    assert len(names) == len(values) and len(queries) == len(positives)
    if not len(names) and not len(queries):
        return {"status": "Empty whitening produced"}, None
    X = _t(values).t()
    name_index = {x: i for i, x in enumerate(names)}
    qidxs = np.array([name_index[x] for x in queries])
    pidxs = np.array([name_index[x] for x in positives])
    time0 = time.time()
    max_trials, max_excluded, trial = 100, 0.95, 0
    while True:
        if trial == 0:
            qwhit, pwhit = qidxs, pidxs
        else:
            idxs = np.random.permutation(len(qidxs))[:int(len(qidxs) * (1 - trial / max_trials * max_excluded))]
            print("Using subset of queries (%s/%s) trial %s" % (len(idxs), len(qidxs), trial), file=sys.stderr)
            qwhit, pwhit = qidxs[idxs], pidxs[idxs]
        try:
            m, P = whitenlearn(X, qwhit, pwhit)
            if not bool(torch.isfinite(P).all()):
                raise torch.linalg.LinAlgError("Matrix is not positive definite")
            break
        except torch.linalg.LinAlgError:
            if trial >= max_trials - 1:
                raise
            trial += 1
    metadata = {"stats": {"failed_times": trial, "vectors_used": round(len(qwhit) / float(len(qidxs)), 2),
                          "vectors_total": len(qidxs)},
                "timings": {"whitening_learn": round(time.time() - time0, 2)}}
    return metadata, {"m": m.cpu().numpy(), "P": P.cpu().numpy()}


def learn_pca_whitening(params, data):
    shrink = params.pop("shrink", None) or None
    assert not params
    values, = data
    if not len(values):
        return {"status": "Empty whitening produced"}, None
    time0 = time.time()
    m, P = pcawhitenlearn(_t(values).t(), shrink)
    return {"timings": {"whitening_learn": round(time.time() - time0, 2)}}, {"m": m.cpu().numpy(), "P": P.cpu().numpy()}
