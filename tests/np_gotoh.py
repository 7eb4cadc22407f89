"""Independent numpy formulation of the frozen Gotoh spec, used to pin the
oracle.  Deliberately NOT the oracle's algorithm: rows are vectorised and the horizontal gap
state is obtained by a prefix maximum, E[j] = max_{k<j}(H^[k] + k*ge) - go - j*ge with
H^ = max(diag + S, F), which is valid because opening a second gap directly after a gap never
beats extending it when go >= 0.
"""
import numpy as np


def gotoh_np(a, b, mat, go, ge):
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    mat = np.asarray(mat, dtype=np.int64)
    m, n = len(a), len(b)
    if m == 0 and n == 0:
        return 0
    if m == 0:
        return -(go + n * ge)
    if n == 0:
        return -(go + m * ge)
    NEG = -(1 << 40)
    j = np.arange(n + 1, dtype=np.int64)
    H = -(go + j * ge)
    H[0] = 0
    F = np.full(n + 1, NEG, dtype=np.int64)
    for i in range(1, m + 1):
        F = np.maximum(F - ge, H - go - ge)
        Hh = np.empty(n + 1, dtype=np.int64)
        Hh[0] = -(go + i * ge)
        Hh[1:] = np.maximum(H[:-1] + mat[a[i - 1], b], F[1:])
        P = np.maximum.accumulate(Hh[:-1] + j[:-1] * ge)
        E = P - go - j[1:] * ge
        Hn = Hh.copy()
        Hn[1:] = np.maximum(Hh[1:], E)
        H = Hn
    return int(H[n])


def gotoh_py(a, b, mat, go, ge):
    """Third, naive full-matrix version (pure Python; tiny inputs only)."""
    m, n = len(a), len(b)
    if m == 0 and n == 0:
        return 0
    if m == 0:
        return -(go + n * ge)
    if n == 0:
        return -(go + m * ge)
    NEG = -10 ** 12
    H = [[NEG] * (n + 1) for _ in range(m + 1)]
    E = [[NEG] * (n + 1) for _ in range(m + 1)]
    F = [[NEG] * (n + 1) for _ in range(m + 1)]
    H[0][0] = 0
    for i in range(1, m + 1):
        H[i][0] = -(go + i * ge)
    for jj in range(1, n + 1):
        H[0][jj] = -(go + jj * ge)
    for i in range(1, m + 1):
        for jj in range(1, n + 1):
            E[i][jj] = max(E[i][jj - 1] - ge, H[i][jj - 1] - go - ge)
            F[i][jj] = max(F[i - 1][jj] - ge, H[i - 1][jj] - go - ge)
            H[i][jj] = max(H[i - 1][jj - 1] + int(mat[a[i - 1]][b[jj - 1]]), E[i][jj], F[i][jj])
    return H[m][n]
