"""Host-side mirror of tweakseq's alignment-tool wrapper interface for the in-process backend.

``AlignmentTool`` restates tweakseq/Core/AlignmentTool.h:36-71 without Qt (same method names,
argument meaning and defaults: AlignmentTool.cpp:43-63), extended with the two members an
in-process backend needs: ``inProcess()`` and ``run()``.
``B200Gotoh`` is the new tool, shaped like tweakseq/Core/ClustalO.cpp:48-111: same settings
element (``<alignment_tool><name/><path/><preferred/>`` -- ClustalO.cpp:54-61), same
break-on-name-mismatch parse (ClustalO.cpp:63-86).  The C++ twin a maintainer would compile
into tweakseq is host/B200Gotoh.{h,cpp}.
"""
from __future__ import annotations

import ctypes as C
import xml.etree.ElementTree as ET

import numpy as np

from . import capi
from .fasta import filter_cells, read_fasta


class AlignmentTool:
    """Qt-free restatement of tweakseq/Core/AlignmentTool.h:36-71."""

    def __init__(self):
        self.name_ = ""
        self.version_ = ""
        self.executable_ = ""
        self.preferred_ = False   # AlignmentTool.cpp:59-63
        self.usesStdOut_ = False

    def name(self): return self.name_
    def version(self): return self.version_
    def executable(self): return self.executable_
    def setExecutable(self, e): self.executable_ = e
    def setPreferred(self, pref): self.preferred_ = bool(pref)
    def preferred(self): return self.preferred_
    def usesStdOut(self): return self.usesStdOut_

    # virtuals; the base class versions are no-ops (AlignmentTool.cpp:43-53)
    def makeCommand(self, fin, fout):
        """Returns (exec, arglist); the reference fills two out-parameters."""
        return "", []

    def writeSettings(self, parent: ET.Element): pass
    def readSettings(self, doc: ET.Element): pass

    # extension for in-process tools
    def inProcess(self): return False

    def run(self, fin, fout, log=None, cancel=None):
        raise NotImplementedError


class B200Gotoh(AlignmentTool):
    """The in-process tool: all-vs-all Gotoh scores, guide-tree distances, tree and the progressive alignment on a B200."""

    def __init__(self):
        super().__init__()
        self.name_ = "b200gotoh"
        self.executable_ = capi.library_path()   # "path" = the shared library (ClustalO.cpp:96)
        self.alphabet = capi.PROTEIN
        self.gap_open = -1
        self.gap_extend = -1
        self.device = 0
        self.devices = 1           # B200s of the box one job uses (tsq_params.n_devices; -1 = all of them)
        self.identity = False      # ClustalW-style identity distance instead of the score distance
        self.kimura = False        # ... Kimura-corrected, -ln(1 - D - D^2/5); a pair with D >= 0.75 is an error (implies identity)
        self.align = True          # run(): fout = the multiple alignment readNewAlignment ingests; False: the distance matrix
        self.keep_distmat = False  # with align: also write <fout>.distmat
        self.keep_tree = False     # with align: also write <fout>.dnd (without align the tree always accompanies the matrix)
        self.last_stats: dict = {}

    def inProcess(self): return True

    def makeCommand(self, fin, fout):
        # What startAlignment() would exec for an external tool (ClustalO.cpp:48-52).  For the
        # in-process tool this is informational.
        return "", ["--in-process", "-i", fin, "--outfmt=fa" if self.align else "--distmat-out", fout]

    def writeSettings(self, parent: ET.Element):
        e = ET.SubElement(parent, "alignment_tool")           # ClustalO.cpp:54-61
        ET.SubElement(e, "name").text = self.name()
        ET.SubElement(e, "path").text = self.executable()
        ET.SubElement(e, "preferred").text = "yes" if self.preferred() else "no"
        ET.SubElement(e, "gap_open").text = str(self.gap_open)
        ET.SubElement(e, "gap_extend").text = str(self.gap_extend)
        ET.SubElement(e, "device").text = str(self.device)
        ET.SubElement(e, "devices").text = str(self.devices)
        ET.SubElement(e, "alphabet").text = {capi.ALPHABET_AUTO: "auto", capi.NUCLEOTIDE: "nucleotide"}.get(self.alphabet, "protein")
        ET.SubElement(e, "align_in_process").text = "yes" if self.align else "no"

    def readSettings(self, doc: ET.Element):
        for node in doc.iter("alignment_tool"):                # ClustalO.cpp:63-86
            for elem in list(node):
                if elem.tag == "name" and (elem.text or "") != self.name_:
                    break
                if elem.tag == "path":
                    self.executable_ = elem.text or ""
                if elem.tag == "preferred":
                    self.setPreferred((elem.text or "") == "yes")
                if elem.tag == "gap_open":
                    self.gap_open = int(elem.text)
                if elem.tag == "gap_extend":
                    self.gap_extend = int(elem.text)
                if elem.tag == "device":
                    self.device = int(elem.text)
                if elem.tag == "devices":
                    self.devices = int(elem.text)
                if elem.tag == "alphabet":
                    self.alphabet = {"auto": capi.ALPHABET_AUTO, "nucleotide": capi.NUCLEOTIDE}.get(elem.text or "", capi.PROTEIN)
                if elem.tag == "align_in_process":
                    self.align = (elem.text or "") == "yes"
        self.getVersion()

    def getVersion(self):
        # ClustalO.cpp:100-111 runs `clustalo --version`; here the library reports it.
        try:
            self.version_ = capi.load_library().tsq_version_string().decode()
        except capi.TsqError:
            self.version_ = ""
        return self.version_

    # ---- the in-process path -------------------------------------------------------------
    def run(self, fin, fout, log=None, cancel: C.c_int | None = None) -> int:
        """FASTA file in (what Project::exportFASTA wrote); out, with ``align`` set (the default), the
        multiple alignment itself (FASTA, tree order): the file Project::readNewAlignment
        (Project.cpp:908-1032) reads back, no external aligner involved -- with ``keep_tree`` the tree goes to
        <fout>.dnd and with ``keep_distmat`` the matrix to <fout>.distmat.  ``alphabet`` may be ALPHABET_AUTO here
        (decided from the file's residues).  With ``align`` off, fout is the PHYLIP distance matrix for clustalo.

        Returns the exit status startAlignment()/alignmentFinished() would see (0 = success:
        SeqEditMainWin.cpp:836-861)."""
        flags = ((capi.FLAG_IDENTITY if (self.identity or self.kimura) else 0) | (capi.FLAG_KIMURA if self.kimura else 0) | (capi.FLAG_MSA_OUT if self.align else 0) |
                 (capi.FLAG_KEEP_DISTMAT if self.keep_distmat else 0) | (capi.FLAG_KEEP_TREE if self.keep_tree else 0))
        return capi.run_fasta(fin, fout, log=log, cancel=cancel, alphabet=self.alphabet,
                              gap_open=self.gap_open, gap_extend=self.gap_extend, device=self.device, flags=flags,
                              n_devices=self.devices)

    def multiple_alignment(self, residues):
        """Distances, UPGMA guide tree and the progressive alignment along it, in memory:
        (rows in submitted order, tree order of the rows)."""
        with capi.Context(alphabet=self.alphabet, gap_open=self.gap_open, gap_extend=self.gap_extend,
                          device=self.device, n_devices=self.devices, flags=capi.FLAG_IDENTITY if self.identity else 0) as ctx:
            ctx.set_sequences(residues)
            ctx.run()
            rows, order = ctx.msa()
            self.last_stats = ctx.stats()
            return rows, order

    def distance_matrix(self, residues, labels=None, progress=None, cancel=None, flags: int = 0):
        """Scores and distances for in-memory residues (what Sequence::filter(true) returns).

        Returns (scores int32 packed, distances float64 packed)."""
        if self.identity or self.kimura:
            flags |= capi.FLAG_IDENTITY
        if self.kimura:
            flags |= capi.FLAG_KIMURA
        with capi.Context(alphabet=self.alphabet, gap_open=self.gap_open, gap_extend=self.gap_extend,
                          device=self.device, n_devices=self.devices, flags=flags) as ctx:
            ctx.set_sequences(residues)
            ctx.run(progress=progress, cancel=cancel)
            self.last_stats = ctx.stats()
            d = None if flags & capi.FLAG_NO_DISTANCES else ctx.distances()
            return ctx.scores(), d

    def guide_tree(self, residues, labels=None, newick_path=None):
        """Distances + UPGMA guide tree.  Returns (left, right, height) merge arrays and
        writes Newick to newick_path if given (what clustalo takes as --guidetree-in)."""
        with capi.Context(alphabet=self.alphabet, gap_open=self.gap_open, gap_extend=self.gap_extend,
                          device=self.device, n_devices=self.devices, flags=capi.FLAG_IDENTITY if self.identity else 0) as ctx:
            ctx.set_sequences(residues)
            ctx.run()
            tree = ctx.guide_tree()
            if newick_path:
                ctx.write_newick(newick_path, labels)
            self.last_stats = ctx.stats()
            return tree

    def pairwise_alignment(self, residues_a, residues_b):
        """One optimal global alignment of two sequences with its path: (row_a, row_b, score).
        What running the external tool on a two-record FASTA would hand back, in process."""
        with capi.Context(alphabet=self.alphabet, gap_open=self.gap_open, gap_extend=self.gap_extend,
                          device=self.device) as ctx:
            ctx.set_sequences([residues_a, residues_b])
            ctx.upload()
            return ctx.align_pair(0, 1)

    def consensus(self, aligned_rows, plurality: float = -1.0) -> str:
        """Consensus annotation of an alignment: Consensus::calculate (Consensus.cpp:80-161) on the GPU."""
        with capi.Context(device=self.device) as ctx:
            return ctx.consensus(aligned_rows, plurality)

    def distance_matrix_from_cells(self, cell_rows, applyExclusions=True, **kw):
        """Rows of 16-bit residue cells with tweakseq's flag bits (Sequence.h:36-39)."""
        return self.distance_matrix([filter_cells(r, applyExclusions) for r in cell_rows], **kw)

    def distance_matrix_from_fasta(self, path, **kw):
        labels, seqs, _ = read_fasta(path)
        s, d = self.distance_matrix(seqs, **kw)
        return labels, s, d


def square(packed: np.ndarray, n: int, diag=0):
    """Packed upper triangle -> full symmetric n x n matrix."""
    m = np.full((n, n), diag, dtype=packed.dtype)
    iu = np.triu_indices(n, 1)
    m[iu] = packed
    m[(iu[1], iu[0])] = packed
    return m
