"""UPGMA guide tree from the distance matrix.  CPU leg: the oracle against scipy's
independent average-linkage implementation and closed-form cases; gpu leg: the CUDA tree bit-exact
against the oracle (merge order, node ids, heights) and the Newick file."""
import os

import numpy as np
import pytest

from oracle import pyoracle as o


def _packed_random(n, seed):
    rng = np.random.default_rng(seed)
    return rng.random(n * (n - 1) // 2) + 0.05          # no ties with probability 1


def _clusters_from_merges(left, right, n):
    members = {i: frozenset([i]) for i in range(n)}
    out = []
    for t in range(n - 1):
        s = members[int(left[t])] | members[int(right[t])]
        members[n + t] = s
        out.append(s)
    return out


def test_upgma_three_leaf_closed_form():
    # d(0,1)=2, d(0,2)=6, d(1,2)=10 -> merge (0,1) at height 1, then with 2 at (6+10)/2/2 = 4
    l, r, h = o.upgma(np.array([2.0, 6.0, 10.0]), 3)
    assert l.tolist() == [0, 3] and r.tolist() == [1, 2] and h.tolist() == [1.0, 4.0]
    assert o.newick(l, r, h, ["a", "b", "c"]) == "((a:1.000000,b:1.000000):3.000000,c:4.000000);"


def test_upgma_tie_break_is_smallest_slot_pair():
    # all distances equal: merges (0,1), then slot 0 with 2, then with 3
    l, r, h = o.upgma(np.full(6, 1.0), 4)
    assert l.tolist() == [0, 4, 5] and r.tolist() == [1, 2, 3] and h.tolist() == [0.5, 0.5, 0.5]


@pytest.mark.parametrize("n,seed", [(2, 1), (5, 2), (40, 3), (150, 4)])
def test_upgma_oracle_matches_scipy_average_linkage(n, seed):
    from scipy.cluster.hierarchy import linkage
    d = _packed_random(n, seed)
    l, r, h = o.upgma(d, n)
    Z = linkage(d, method="average")                     # condensed order == our packed order
    ours = _clusters_from_merges(l, r, n)
    theirs = _clusters_from_merges(Z[:, 0].astype(int), Z[:, 1].astype(int), n)
    assert set(ours) == set(theirs)
    # heights: scipy reports the cluster distance, ours is half of it; compare per cluster
    hz = {c: Z[t, 2] / 2 for t, c in enumerate(theirs)}
    for t, c in enumerate(ours):
        assert abs(h[t] - hz[c]) <= 1e-12 * max(1.0, abs(hz[c]))
    assert (np.diff(h) >= -1e-15).all()                   # UPGMA heights never decrease


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 3, 33, 200, 700])
def test_gpu_guide_tree_is_bit_exact(n, tmp_path):
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    seqs = synth.protein(n, (40, 160, 100, 30), 1, family=(n % 2 == 0))
    with t.Context() as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        d = ctx.distances()
        l, r, h = ctx.guide_tree()
        path = str(tmp_path / "tree.dnd")
        labels = [f"q{k}" for k in range(n)]
        ctx.write_newick(path, labels)
    ol, orr, oh = o.upgma(d, n)
    assert (l == ol).all() and (r == orr).all()
    assert h.tobytes() == oh.tobytes()
    assert open(path).read().strip() == o.newick(ol, orr, oh, labels)


@pytest.mark.gpu
def test_gpu_guide_tree_with_many_ties_and_duplicates():
    import tweakseq_b200 as t
    base = ["ACDEFGHIKLMNPQRSTVWY" * 3, "WWWWWCCCCCHHHHH" * 4, "ACDEFGHIKL" * 6]
    seqs = [base[k % 3] for k in range(60)] + ["", "A"]      # exact duplicates: many zero distances / ties
    with t.Context() as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        d = ctx.distances()
        l, r, h = ctx.guide_tree()
    ol, orr, oh = o.upgma(d, len(seqs))
    assert (l == ol).all() and (r == orr).all() and h.tobytes() == oh.tobytes()


@pytest.mark.gpu
def test_run_fasta_writes_the_guide_tree_next_to_the_matrix(tmp_path):
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    from tweakseq_b200.fasta import write_fasta
    seqs = synth.protein(25, (50, 120, 80, 20), 1)
    labels = [f"p{k}" for k in range(25)]
    fin, fout = str(tmp_path / "in.fa"), str(tmp_path / "out.mat")
    write_fasta(fin, labels, seqs)
    tool = t.B200Gotoh()
    tool.align = False
    assert tool.run(fin, fout) == 0
    txt = open(fout + ".dnd").read().strip()
    assert txt.endswith(";") and txt.count("(") == 24 and all(f"{l}:" in txt for l in labels)


@pytest.mark.gpu
def test_newick_labels_with_structure_characters_are_quoted_not_rewritten(tmp_path):
    """The tree file must name the leaves as the matrix file and the FASTA headers do (clustalo reads both): a label
    holding Newick structure characters goes in single quotes, a quote inside doubled."""
    import tweakseq_b200 as t
    with t.Context() as ctx:
        ctx.set_sequences(["ACDEFGHIK", "ACDEFGHIR", "WWWWWWW"])
        ctx.run()
        path = str(tmp_path / "t.dnd")
        ctx.write_newick(path, ["sp|P1|A:1", "b (x)", "c;d'e"])
    txt = open(path).read().strip()
    assert "'sp|P1|A:1':" in txt and "'b (x)':" in txt and "'c;d''e':" in txt
    bare = txt.replace("'sp|P1|A:1'", "a").replace("'b (x)'", "b").replace("'c;d''e'", "c")
    assert bare.count("(") == 2 and bare.count(")") == 2 and bare.endswith(";") and bare.count(";") == 1
    with t.Context() as ctx:
        ctx.set_sequences(["ACDEFGHIK", "ACDEFGHIR", "WWWWWWW"])
        ctx.run()
        ctx.write_newick(path, ["plain_1", "sp|P2|B", "x.y-z"])
    assert "'" not in open(path).read()          # ordinary labels stay bare
