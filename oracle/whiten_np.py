"""ORACLE (test infrastructure, never imported by gandtr_b200/): NumPy restatement of learned whitening.

Follows mdir/external/cirtorch/utils/whiten.py:4-12 (whitenapply), :37-53 (whitenlearn), :55-70 (cholesky with
diagonal loading). float64 like the reference. Parity pin: tests/test_oracle_whiten.py against the unmodified
reference functions (fixture tests/golden/whiten_*.npz from tools/gen_golden.py). Eigenvector signs are not
unique: comparisons are made on |P| rows / on whitened descriptors up to sign.
"""
import numpy as np


def cholesky_loaded(S):
    alpha = 0.0
    while True:
        try:
            return np.linalg.cholesky(S + alpha * np.eye(*S.shape))
        except np.linalg.LinAlgError:
            alpha = 1e-10 if alpha == 0 else alpha * 10


def whitenlearn(X, qidxs, pidxs):
    """X: [D, n] float64; (qidxs, pidxs) matching pairs -> (m [D,1], P [D,D])."""
    m = X[:, qidxs].mean(axis=1, keepdims=True)
    df = X[:, qidxs] - X[:, pidxs]
    S = df @ df.T / df.shape[1]
    P = np.linalg.inv(cholesky_loaded(S))
    df = P @ (X - m)
    D = df @ df.T
    eigval, eigvec = np.linalg.eig(D)
    order = eigval.argsort()[::-1]
    eigvec = eigvec[:, order]
    return m, eigvec.T @ P


def whitenapply(X, m, P, dimensions=None):
    dimensions = dimensions or P.shape[0]
    Y = P[:dimensions, :] @ (X - m)
    return Y / (np.linalg.norm(Y, ord=2, axis=0, keepdims=True) + 1e-6)
