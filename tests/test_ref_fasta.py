"""The FASTA wire format either side of the tool against the reference's OWN code:
oracle/_ref/libref_fasta.so is tweakseq/Core/FASTAFile.cpp + SequenceFile.cpp compiled where they lie
(oracle/Makefile, functional Qt stand-ins in oracle/ref_shim/).  CPU leg: the host-side restatement
(tweakseq_b200/fasta.py) reads and writes exactly what the reference's reader / writer do, quirks included.
The GPU leg (tests/test_zz_aligner_cli.py) feeds the product files written by the reference's writer and parses
the product's output with the reference's reader."""
import numpy as np
import pytest

from oracle import pyoracle as o
from tweakseq_b200 import fasta

pytestmark = pytest.mark.skipif(not o.ref_fasta_available(), reason="oracle/_ref/libref_fasta.so not built "
                                                                    "(needs /root/reference at build time)")

AA = "ARNDCQEGHILKMFPSTWYV"


def test_reader_restatement_equals_the_references_reader_on_awkward_files(tmp_path):
    rng = np.random.default_rng(8)
    path = str(tmp_path / "t.fa")
    fixed = [
        ">a desc one\nMKV\n\n  LLA  \n>b\n;skipped comment\nAC-D\r\n>c x\nWW\n",
        ";old style header\nMKV\n>b two words\nAC\nDE\n",
        "junk before any header\n>a\nMK\n",
        ">a\n>b\nMK\n",                       # a header right after a header is read as RESIDUES by the reference
        ">only\n",
        "",
        "\n\n>a  \n  M K V  \n",
        ">a\tTabbed description\nMKV\n>b\nmkv*\n",
    ]
    for text in fixed:
        open(path, "w", newline="").write(text)
        assert fasta.read_fasta(path, strict=True) == o.ref_fasta_read(path), repr(text)
    for trial in range(60):                   # random mixtures of the same ingredients
        lines = []
        for _ in range(int(rng.integers(0, 25))):
            u = rng.random()
            if u < 0.25:
                lines.append(">" + "".join(rng.choice(list("abcXYZ_|1 "), int(rng.integers(0, 12)))))
            elif u < 0.32:
                lines.append(";" + "".join(rng.choice(list("abc "), int(rng.integers(0, 8)))))
            elif u < 0.42:
                lines.append("")
            else:
                lines.append(" " * int(rng.integers(0, 3)) + "".join(rng.choice(list(AA + "-"), int(rng.integers(1, 90)))) +
                             " " * int(rng.integers(0, 3)))
        eol = "\r\n" if trial % 3 == 0 else "\n"
        open(path, "w", newline="").write(eol.join(lines) + (eol if trial % 2 else ""))
        assert fasta.read_fasta(path, strict=True) == o.ref_fasta_read(path), (trial, lines)


def test_writer_restatement_equals_the_references_writer(tmp_path):
    rng = np.random.default_rng(9)
    a, b = str(tmp_path / "ours.fa"), str(tmp_path / "ref.fa")
    for trial in range(30):
        n = int(rng.integers(0, 8))
        lens = [int(x) for x in rng.choice([0, 1, 79, 80, 81, 159, 160, 161, 240, 333], n)]
        seqs = ["".join(rng.choice(list(AA + "-"), l)) for l in lens]
        labels = [f"s{k}" for k in range(n)]
        comments = [f">{l} some description {k}" for k, l in enumerate(labels)]
        fasta.write_fasta(a, labels, seqs, comments)
        o.ref_fasta_write(b, labels, seqs, comments)
        assert open(a, "rb").read() == open(b, "rb").read(), (trial, lens)
    # and what the reference writes, the reference reads back (non-empty sequences; an empty one makes its
    # reader take the next header for residues -- the hazard INTEGRATION.md lists)
    seqs = ["".join(rng.choice(list(AA), int(l))) for l in (5, 80, 200)]
    o.ref_fasta_write(b, ["x", "y", "z"], seqs, [">x", ">y d", ">z"])
    assert o.ref_fasta_read(b) == (["x", "y", "z"], seqs, [">x", ">y d", ">z"])


@pytest.mark.skipif(not o.ref_sequence_available(), reason="oracle/_ref/libref_sequence.so not built")
def test_cell_filter_restatements_equal_the_references_sequence_filter():
    """Sequence::filter (Sequence.cpp:57-69) is what exportFASTA applies before the aligner sees a sequence: the
    Python restatement and the C++ adapter's filterCells (same rule, host/B200Gotoh.cpp) against the reference's
    own compiled code on cells carrying its flag bits (Sequence.h:36-39)."""
    rng = np.random.default_rng(12)
    for trial in range(200):
        n = int(rng.integers(0, 60))
        cells = []
        for _ in range(n):
            v = ord(str(rng.choice(list(AA + "-acx"))))
            u = rng.random()
            if u < 0.2:
                v |= fasta.EXCLUDE_CELL
            if rng.random() < 0.2:
                v |= fasta.HIGHLIGHT_CELL
            if rng.random() < 0.05:
                v |= 0x8000                      # any higher bit is a flag too (REMOVE_FLAGS = 0x007F)
            cells.append(v)
        for apply in (True, False):
            want = "".join(chr(v) for v in o.ref_filter(cells, apply))
            assert fasta.filter_cells(cells, apply) == want, (trial, apply)
