"""Problem constructors returning GPU-resident (f, h, L, x0).

Signature-compatible with accbpg/applications.py:17-206: the random instances draw from the
legacy global NumPy RNG in the same order, so a seeded call builds the same H / A / b as the
reference; f and h are the device operators of this package, L and x0 are host values exactly
as in the reference (x0 is uploaded by the driver).
"""
import os

import numpy as np

from .objectives import SparseDOptimalObj, DOptimalObj, PoissonRegression, KLdivRegression
from .bregman import BurgEntropySimplex, BurgEntropyL1, BurgEntropyL2, ShannonEntropyL1


def _parse_libsvm(filename, zero_based="auto"):
    """Rows of a LIBSVM / svmlight text file as lists of (feature index, value), the labels, the index shift and the
    feature count.  Same conventions as accbpg/utils.py:22-95 ('#' comments, sorted unique 1-based indices by default,
    feature count from the largest index)."""
    rows, labels = [], []
    opener = open
    if filename.endswith(".gz"):
        import gzip
        opener = lambda p: gzip.open(p, "rt")
    elif filename.endswith(".bz2"):
        import bz2
        opener = lambda p: bz2.open(p, "rt")
    min_idx, max_idx = None, -1
    with opener(filename) as fh:
        for line in fh:
            line = line.split("#", 1)[0].split()
            if not line:
                continue
            labels.append(float(line[0]))
            feats = []
            prev = -1
            for tok in line[1:]:
                k, v = tok.split(":", 1)
                k = int(k)
                if k < 0 or (zero_based is False and k == 0):
                    raise ValueError("Invalid index {0:d} in LibSVM data file.".format(k))
                if k <= prev:
                    raise ValueError("Feature indices in LibSVM data file should be sorted and unique.")
                prev = k
                feats.append((k, float(v)))
                min_idx = k if min_idx is None else min(min_idx, k)
                max_idx = max(max_idx, k)
            rows.append(feats)
    shift = 1 if (zero_based is False or (zero_based == "auto" and min_idx is not None and min_idx > 0)) else 0
    return rows, np.array(labels), shift, max_idx + 1 - shift


def load_libsvm_dense(filename, zero_based="auto"):
    """Parse a LIBSVM / svmlight text file into a dense (n_samples, n_features) float64 array and labels
    (accbpg/utils.py:22-95 followed by .toarray, as accbpg/applications.py:21-25 does)."""
    rows, labels, shift, nfeat = _parse_libsvm(filename, zero_based)
    X = np.zeros((len(rows), nfeat))
    for i, feats in enumerate(rows):
        for k, v in feats:
            X[i, k - shift] = v
    return X, labels


def load_libsvm_sparse(filename, zero_based="auto"):
    """The same file as compressed sparse ROWS of X (= compressed columns of H = X^T, one column per sample, in file
    order): (indptr int64 [n_samples + 1], indices int32 [nnz], values float64 [nnz], n_features, labels).  Explicit
    zeros of the file are kept as stored entries, as scipy's csr_matrix keeps them."""
    rows, labels, shift, nfeat = _parse_libsvm(filename, zero_based)
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    for i, feats in enumerate(rows):
        indptr[i + 1] = indptr[i] + len(feats)
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    values = np.empty(int(indptr[-1]), dtype=np.float64)
    q = 0
    for feats in rows:
        for k, v in feats:
            indices[q] = k - shift
            values[q] = v
            q += 1
    return indptr, indices, values, nfeat, labels


def D_opt_libsvm(filename, device=None, sparse=False):
    """D-optimal design instance from a LIBSVM regression data set.   applications.py:17-33.
    sparse=True keeps H = X^T in compressed-column form on the device (SparseDOptimalObj) instead of densifying it as the
    reference does with .toarray('C'); it needs more samples than features (otherwise the dense path is taken)."""
    if sparse:
        indptr, indices, values, nfeat, _ = load_libsvm_sparse(filename)
        n = indptr.size - 1
        if n > nfeat:
            f = SparseDOptimalObj(indptr, indices, values, nfeat, device=device)
            return f, BurgEntropySimplex(device=device), 1.0, (1.0 / n) * np.ones(n)
    X, _ = load_libsvm_dense(filename)
    H = np.ascontiguousarray(X.T) if X.shape[0] > X.shape[1] else np.ascontiguousarray(X)
    n = H.shape[1]
    return DOptimalObj(H, device=device), BurgEntropySimplex(device=device), 1.0, (1.0 / n) * np.ones(n)


def D_opt_design(m, n, randseed=-1, device=None):
    """Random D-optimal design instance, H = randn(m, n).   applications.py:36-56."""
    if randseed > 0:
        np.random.seed(randseed)
    H = np.random.randn(m, n)
    return DOptimalObj(H, device=device), BurgEntropySimplex(device=device), 1.0, (1.0 / n) * np.ones(n)


def _poisson_data(m, n, noise, randseed, normalizeA):
    if randseed > 0:
        np.random.seed(randseed)
    A = np.random.rand(m, n)
    if normalizeA:
        A = A / A.sum(axis=0)
    x = np.random.rand(n) / n
    xavg = x.sum() / x.size
    x = np.maximum(x - xavg, 0) * 10
    b = np.dot(A, x) + noise * (np.random.rand(m) - 0.5)
    assert b.min() > 0, "need b > 0 for nonnegative regression."
    return A, b


def Poisson_regrL1(m, n, noise=0.01, lamda=0, randseed=-1, normalizeA=True, device=None):
    """min D_KL(b, Ax) + lamda*||x||_1 over x >= 0.   applications.py:98-134."""
    A, b = _poisson_data(m, n, noise, randseed, normalizeA)
    return (PoissonRegression(A, b, device=device), BurgEntropyL1(lamda, device=device), b.sum(),
            (1.0 / n) * np.ones(n) * 10)


def Poisson_regrL2(m, n, noise=0.01, lamda=0, randseed=-1, normalizeA=True, device=None):
    """min D_KL(b, Ax) + (lamda/2)*||x||_2^2 over x >= 0.   applications.py:137-172."""
    A, b = _poisson_data(m, n, noise, randseed, normalizeA)
    return (PoissonRegression(A, b, device=device), BurgEntropyL2(lamda, device=device), b.sum(),
            (1.0 / n) * np.ones(n))


def KL_nonneg_regr(m, n, noise=0.01, lamdaL1=0, randseed=-1, normalizeA=True, device=None):
    """min D_KL(Ax, b) + lamda*||x||_1 over x >= 0.   applications.py:175-206."""
    if randseed > 0:
        np.random.seed(randseed)
    A = np.random.rand(m, n)
    if normalizeA:
        A = A / A.sum(axis=0)
    x = np.random.rand(n)
    b = np.dot(A, x) + noise * (np.random.rand(m) - 0.5)
    assert b.min() > 0, "need b > 0 for nonnegative regression."
    L = max(A.sum(axis=0))
    return KLdivRegression(A, b, device=device), ShannonEntropyL1(lamdaL1, device=device), L, 0.5 * np.ones(n)


def D_opt_KYinit(V, device=None):
    """Kumar-Yildirim sparse starting point (applications.py:59-95).

    m Gram-Schmidt sweeps, each with one product q^T V over all of V (8mn bytes: the whole cost, m passes over the
    design matrix).  That pass, the arg-max / arg-min over its n results and the gather of the two chosen columns run
    on the device (K6 `rmatvec`, K14 `argext`); the m x m Gram-Schmidt bookkeeping on m-vectors between the passes is
    host scalar control in the reference's exact order, and the random directions come from the legacy NumPy stream as
    in the reference.  Returns a host NumPy vector like the reference."""
    import torch
    from . import _native as nat
    from .runtime import Runtime
    lib = nat.lib
    rt = Runtime.get(device)
    Vd = rt.to_device(V)
    m, n = int(Vd.shape[0]), int(Vd.shape[1])
    if n <= 2 * m:
        return (1.0 / n) * np.ones(n)
    ws = rt.workspace(("linreg", m, n), lib.accbpg_linreg_workspace_bytes(m, n))
    qV = rt.empty(n)
    host_V = None if isinstance(V, torch.Tensor) else np.asarray(V)
    picked = []
    Q = np.zeros((m, m))
    s0 = rt.S_TMP
    for i in range(m):
        b = np.random.rand(m)
        q = b.copy()
        for j in range(i):
            q = q - np.dot(Q[:, j], b) * Q[:, j]
        qd = rt.to_device(q)
        nat.check(lib.accbpg_linreg_rmatvec(rt.ctx, rt.stream, Vd.data_ptr(), m, n, Vd.stride(0), qd.data_ptr(),
                                            ws.data_ptr(), qV.data_ptr()))
        nat.check(lib.accbpg_vec_argext(rt.ctx, rt.stream, n, qV.data_ptr(), 1, rt.slot(s0)))
        nat.check(lib.accbpg_vec_argext(rt.ctx, rt.stream, n, qV.data_ptr(), 0, rt.slot(s0 + 2)))
        vals = rt.read(s0, 4)
        kmax, kmin = int(vals[1]), int(vals[3])
        picked += [kmax, kmin]
        if host_V is not None:
            v = host_V[:, kmin] - host_V[:, kmax]
        else:
            cols = Vd[:, [kmin, kmax]].cpu().numpy()          # 2m doubles back (data movement only)
            v = cols[:, 0] - cols[:, 1]
        q = v.copy()
        for j in range(i):
            q = q - np.dot(Q[:, j], v) * Q[:, j]
        Q[:, i] = q / np.linalg.norm(q)
    x0 = np.zeros(n)
    x0[picked] = np.ones(len(picked)) / len(picked)
    return x0 / x0.sum()