"""Seeded synthetic inputs of BASELINE.json's configurations."""
from __future__ import annotations

import numpy as np

SEED0 = 20261017
AA20 = "ARNDCQEGHILKMFPSTWYV"
# Robinson & Robinson (1991) background frequencies, order of AA20
_RR = np.array([0.07805, 0.05129, 0.04487, 0.05364, 0.01925, 0.04264, 0.06295, 0.07377, 0.02199,
                0.05142, 0.09019, 0.05744, 0.02243, 0.03856, 0.05203, 0.07120, 0.05841, 0.01330,
                0.03216, 0.06441])
_RR = _RR / _RR.sum()
_AA = np.frombuffer(AA20.encode(), dtype=np.uint8)
_NT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rng(config: int):
    return np.random.Generator(np.random.PCG64(SEED0 + config))


def protein(n: int, length, config: int = 2, family: bool = False) -> list[str]:
    """n protein sequences; length = int (fixed) or (lo, hi, mean, sd) for clipped normal."""
    rng = _rng(config)
    if isinstance(length, int):
        lens = np.full(n, length)
    else:
        lo, hi, mean, sd = length
        lens = np.clip(np.rint(rng.normal(mean, sd, n)), lo, hi).astype(int)
    if not family:
        return [_AA[rng.choice(20, int(l), p=_RR)].tobytes().decode() for l in lens]
    root = rng.choice(20, int(lens.max()), p=_RR)
    out = []
    for l in lens:
        s = root[: int(l)].copy()
        rate = rng.uniform(0.10, 0.60)
        mut = rng.random(len(s)) < rate
        s[mut] = rng.choice(20, int(mut.sum()), p=_RR)
        if len(s) > 20 and rng.random() < 0.7:   # one indel
            a = int(rng.integers(1, len(s) - 10))
            k = int(rng.integers(1, 9))
            s = np.concatenate([s[:a], s[a + k:], rng.choice(20, k, p=_RR)])
        out.append(_AA[s].tobytes().decode())
    return out


def nucleotide(n: int, lo: int, hi: int, config: int = 4, family: bool = False) -> list[str]:
    rng = _rng(config)
    lens = rng.integers(lo, hi + 1, n)
    if not family:
        return [_NT[rng.integers(0, 4, int(l))].tobytes().decode() for l in lens]
    root = rng.integers(0, 4, int(lens.max()))
    out = []
    for l in lens:
        s = root[: int(l)].copy()
        mut = rng.random(len(s)) < rng.uniform(0.02, 0.3)
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        out.append(_NT[s].tobytes().decode())
    return out


def config(idx: int, scale: float = 1.0):
    """(alphabet, sequences) of BASELINE.json configs[idx-1]; scale shrinks n for twins."""
    if idx == 1:
        return 0, protein(max(2, int(100 * scale)), (200, 400, 300, 30), 1)
    if idx == 2:
        return 0, protein(max(2, int(1000 * scale)), 300, 2)
    if idx == 3:
        return 0, protein(max(2, int(10000 * scale)), 400, 3)
    if idx == 4:
        return 1, nucleotide(max(2, int(500 * scale)), 10000, 30000, 4)
    if idx == 5:
        return 0, protein(max(2, int(100000 * scale)), 150, 5)
    raise ValueError(idx)


def total_cells(seqs) -> int:
    l = np.array([len(s) for s in seqs], dtype=np.int64)
    return int((l.sum() ** 2 - (l * l).sum()) // 2)
