"""The FASTA wire format either side of the tool, as tweakseq reads and writes it.

Restates tweakseq/Core/FASTAFile.cpp:71-187 (reader state machine, 80-column writer, label
rule) and the residue-cell filter of tweakseq/Core/Sequence.cpp:57-69 -- host-side text
handling only; nothing here touches scores.
"""
from __future__ import annotations

EXCLUDE_CELL = 0x0080    # Sequence.h:36
HIGHLIGHT_CELL = 0x0100  # Sequence.h:37
REMOVE_FLAGS = 0x007F    # Sequence.h:39


def filter_cells(cells, applyExclusions: bool = True) -> str:
    """Sequence::filter (Sequence.cpp:57-69): drop excluded cells, strip flag bits; '-' stays."""
    out = []
    for c in cells:
        u = ord(c) if isinstance(c, str) else int(c)
        if (u & EXCLUDE_CELL) and applyExclusions:
            continue
        out.append(chr(u & REMOVE_FLAGS))
    return "".join(out)


def parse_comment(s: str) -> str:
    """FASTAFile::parseComment (FASTAFile.cpp:177-187): header[1 : first space after pos 1]."""
    idx = s.find(" ", 1)
    return s[1:] if idx == -1 else s[1:idx]


def read_fasta(path: str, strict: bool = False):
    """FASTAFile::read (FASTAFile.cpp:71-147).  Returns (labels, sequences, comments).

    strict=True is the reference's reader to the letter (checked against its own compiled code in
    tests/test_ref_fasta.py): a final header without residues leaves `sequences` one entry short.  The default
    pads that entry with "" so that the three lists always line up."""
    SEEKING, COMMENT, SEQ = 0, 1, 2
    labels, seqs, comments = [], [], []
    state = SEEKING
    with open(path, "r", encoding="latin-1") as f:
        for raw in f:
            s = raw.strip()
            if not s:
                continue
            first = s[0]
            hdr = first in ";>"
            if state == SEEKING:
                if hdr:
                    state = COMMENT
                    comments.append(s)
                    labels.append(parse_comment(s))
            elif state == COMMENT:
                if first == ";":
                    continue
                state = SEQ
                seqs.append(s)
            else:
                if hdr:
                    state = COMMENT
                    comments.append(s)
                    labels.append(parse_comment(s))
                else:
                    seqs[-1] += s
    while not strict and len(seqs) < len(labels):   # a trailing header with no residues
        seqs.append("")
    return labels, seqs, comments


def write_fasta(path: str, labels, seqs, comments=None):
    """FASTAFile::write (FASTAFile.cpp:149-171): the comment line, then 80-column lines."""
    if comments is None:
        comments = [">" + l for l in labels]
    with open(path, "w", encoding="latin-1") as f:
        for i in range(len(labels)):
            f.write(comments[i] + "\n")
            s = seqs[i]
            for j in range(0, len(s), 80):
                f.write(s[j:j + 80] + "\n")


def read_distmat(path: str):
    """Reads the PHYLIP-style square matrix tsq_run_fasta writes. Returns (labels, rows)."""
    with open(path) as f:
        n = int(f.readline().split()[0])
        labels, rows = [], []
        for _ in range(n):
            parts = f.readline().split()
            labels.append(parts[0])
            rows.append([float(x) for x in parts[1:]])
    return labels, rows
