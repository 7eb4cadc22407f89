"""Independent pure-Python statement of the progressive-alignment spec (tests only, tiny inputs).

Deliberately different from both oracle/gotoh_oracle.c:tsq_oracle_msa (column counts) and the CUDA
kernels (profiles + column maps): the column score is the literal sum over member PAIRS, clusters are
lists of gapped strings, the DP uses Python integers in dictionaries."""
from __future__ import annotations

NEG = -(1 << 80)


def sub_score(colx, coly, S):
    """Sum over residue pairs (x in colx, y in coly) of S[x][y]; '-' scores 0 against anything."""
    return sum(S[x][y] for x in colx if x != "-" for y in coly if y != "-")


def align_profiles(X, Y, S, go, ge):
    """X, Y: lists of equal-length gapped strings.  Returns (rows of X + rows of Y aligned, score)."""
    m = len(X[0]) if X else 0
    q = len(Y[0]) if Y else 0
    w = len(X) * len(Y)
    GO, GE = w * go, w * ge
    GOE = GO + GE
    colx = [[r[i] for r in X] for i in range(m)]
    coly = [[r[j] for r in Y] for j in range(q)]
    H, E, F, SUB = {}, {}, {}, {}
    H[0, 0], E[0, 0], F[0, 0] = 0, NEG, NEG
    for j in range(1, q + 1):
        H[0, j] = E[0, j] = -GO - j * GE
        F[0, j] = NEG
    for i in range(1, m + 1):
        H[i, 0] = F[i, 0] = -GO - i * GE
        E[i, 0] = NEG
        for j in range(1, q + 1):
            SUB[i, j] = sub_score(colx[i - 1], coly[j - 1], S)
            E[i, j] = max(E[i, j - 1] - GE, H[i, j - 1] - GOE)
            F[i, j] = max(F[i - 1, j] - GE, H[i - 1, j] - GOE)
            H[i, j] = max(H[i - 1, j - 1] + SUB[i, j], E[i, j], F[i, j])
    # walk back: diagonal, then gap in X (E), then gap in Y (F); open rather than extend
    path = []
    i, j, state = m, q, 0
    while i > 0 or j > 0:
        if i == 0:
            path.append((-1, j - 1)); j -= 1; continue
        if j == 0:
            path.append((i - 1, -1)); i -= 1; continue
        if state == 0:
            if H[i, j] == H[i - 1, j - 1] + SUB[i, j]:
                path.append((i - 1, j - 1)); i -= 1; j -= 1
            else:
                state = 1 if H[i, j] == E[i, j] else 2
        elif state == 1:
            path.append((-1, j - 1))
            if E[i, j] == H[i, j - 1] - GOE:
                state = 0
            j -= 1
        else:
            path.append((i - 1, -1))
            if F[i, j] == H[i - 1, j] - GOE:
                state = 0
            i -= 1
    path.reverse()
    rows = ["".join("-" if xi < 0 else r[xi] for xi, _ in path) for r in X]
    rows += ["".join("-" if yj < 0 else r[yj] for _, yj in path) for r in Y]
    return rows, H[m, q]


def progressive(seqs, left, right, S, go, ge):
    """seqs: ungapped strings over the matrix alphabet; merges (left[t], right[t]), node t = n + t.
    Returns (rows in input order, [score of every merge])."""
    n = len(seqs)
    if n == 0:
        return [], []
    members = {r: [r] for r in range(n)}
    rows = {r: [seqs[r]] for r in range(n)}
    scores = []
    for t, (l, r) in enumerate(zip(left, right)):
        new, sc = align_profiles(rows[l], rows[r], S, go, ge)
        rows[n + t] = new
        members[n + t] = members[l] + members[r]
        scores.append(sc)
    root = 0 if n == 1 else 2 * n - 2
    out = [None] * n
    for who, row in zip(members[root], rows[root]):
        out[who] = row
    return out, scores

