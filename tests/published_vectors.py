"""Global alignments whose optimal score is PUBLISHED: the only anchors of the Gotoh arithmetic that lie outside this
repository (the reference holds no alignment code or vectors: DESIGN.md section 3).  Both use linear gaps, i.e. gap
open 0 in the spec's cost go + k * ge.

1. R. Durbin, S. Eddy, A. Krogh, G. Mitchison, "Biological sequence analysis" (1998), section 2.3, figure 2.5: HEAGAWGHEE
   against PAWHEAE, BLOSUM50, gap penalty d = 8 per residue: the global dynamic-programming matrix ends in F = 1.
   (The BLOSUM50 entries below are the ones of the book's figure 2.2 for the letters involved.)
2. The worked example of the Needleman-Wunsch algorithm's encyclopedia entry: GATTACA against GCATGCU, match +1,
   mismatch -1, indel -1: best score 0.
3. An AFFINE gap with end gaps penalised: the Biopython Tutorial's pairwise2 example
   globalms("ACCGT", "ACG", 2, -1, -.5, -.1) -> score 5 (match 2, mismatch -1, a gap of k residues costs
   0.5 + 0.1 (k - 1)).  Scaled by 10 to integers: match 20, mismatch -10, gap of k residues 4 + k, i.e. go 4, ge 1: 50.
4. Rosalind problem GLOB ("Global Alignment with Scoring Matrix"), sample dataset: PLEASANTLY against MEANLY, BLOSUM62,
   linear gap penalty 5: maximum alignment score 8.
5. Rosalind problem GAFF ("Global Alignment with Scoring Matrix and Affine Gap Penalty"), sample dataset: PRTEINS against
   PRTWPSEIN, BLOSUM62, gap opening 11 and extension 1 (a gap of k residues costs 11 + (k - 1), i.e. go 10, ge 1 here):
   maximum alignment score 8 (PRT---EINS / PRTWPSEIN-: a three-residue gap and a penalised end gap).
6. Rosalind problem GCON ("Global Alignment with Constant Gap Penalty"), sample dataset: PLEASANTLY against MEANLY,
   BLOSUM62, every gap costs 5 whatever its length (go 5, ge 0): 13.
7. Rosalind problem EDIT, sample dataset: the edit distance of PLEASANTLY and MEANLY is 5 -- as an alignment with match 0,
   mismatch -1 and indel -1 the optimum is -5.
4, 5 and 6 use the library's DEFAULT protein matrix (matrix None below): they anchor the BLOSUM62 table of
Consensus.cpp:34-59 as the library holds it, too.
"""
import numpy as np

PROTEIN_ORDER = "ARNDCQEGHILKMFPSTWYVBZX"

_B50 = {("A", "A"): 5, ("A", "E"): -1, ("A", "G"): 0, ("A", "H"): -2, ("A", "P"): -1, ("A", "W"): -3,
        ("E", "E"): 6, ("E", "G"): -3, ("E", "H"): 0, ("E", "P"): -1, ("E", "W"): -3,
        ("G", "G"): 8, ("G", "H"): -2, ("G", "P"): -2, ("G", "W"): -3,
        ("H", "H"): 10, ("H", "P"): -2, ("H", "W"): -3, ("P", "P"): 10, ("P", "W"): -4, ("W", "W"): 15}


def blosum50_subset() -> np.ndarray:
    """23 x 23 in the library's symbol order; only the entries among A, E, G, H, P, W are BLOSUM50, the rest 0."""
    m = np.zeros((23, 23), dtype=np.int8)
    for (a, b), v in _B50.items():
        i, j = PROTEIN_ORDER.index(a), PROTEIN_ORDER.index(b)
        m[i, j] = m[j, i] = v
    return m


def edit_distance_matrix() -> np.ndarray:
    m = np.full((23, 23), -1, dtype=np.int8)
    np.fill_diagonal(m, 0)
    return m


def unit_nucleotide(match: int = 1, mismatch: int = -1) -> np.ndarray:
    m = np.full((5, 5), mismatch, dtype=np.int8)
    np.fill_diagonal(m, match)
    return m


# (name, alphabet, sequence a, sequence b, matrix or None = the library's default, gap open, gap extend, published score)
VECTORS = [
    ("Durbin et al. 1998, fig. 2.5", 0, "HEAGAWGHEE", "PAWHEAE", blosum50_subset(), 0, 8, 1),
    ("Needleman-Wunsch worked example", 1, "GATTACA", "GCATGCU", unit_nucleotide(), 0, 1, 0),
    ("Biopython tutorial globalms, x10", 1, "ACCGT", "ACG", unit_nucleotide(20, -10), 4, 1, 50),
    ("Rosalind GLOB sample", 0, "PLEASANTLY", "MEANLY", None, 0, 5, 8),
    ("Rosalind GAFF sample", 0, "PRTEINS", "PRTWPSEIN", None, 10, 1, 8),
    ("Rosalind GCON sample", 0, "PLEASANTLY", "MEANLY", None, 5, 0, 13),
    ("Rosalind EDIT sample", 0, "PLEASANTLY", "MEANLY", edit_distance_matrix(), 0, 1, -5),
]
