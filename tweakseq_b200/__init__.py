"""tweakseq_b200 -- B200-native all-vs-all Gotoh distance-matrix backend for tweakseq.

The product is ``libtsqb200.so`` (C ABI: include/tsq_b200.h; CUDA kernels for sm_100a in
csrc/).  This package is the thin host-side mirror of tweakseq's alignment-tool wrapper
interface (tweakseq/Core/AlignmentTool.h:36-71) on top of that ABI.  There is no CPU compute
path: importing works anywhere, computing requires the built library and a B200.
"""
from .capi import (TsqError, Context, Params, Stats, PROTEIN, NUCLEOTIDE, FLAG_FORCE_S32,
                   FLAG_NO_DISTANCES, FLAG_NO_WAVE16, FLAG_IDENTITY, FLAG_MSA_OUT, FLAG_KEEP_DISTMAT, FLAG_INPUT_ORDER, FLAG_KEEP_TREE, FLAG_KIMURA, FLAG_SCORES_I16, ALPHABET_AUTO,
                   library_path, load_library, pair_index)
from .backend import AlignmentTool, B200Gotoh

__all__ = ["TsqError", "Context", "Params", "Stats", "PROTEIN", "NUCLEOTIDE", "FLAG_FORCE_S32",
           "FLAG_NO_DISTANCES", "FLAG_NO_WAVE16", "FLAG_IDENTITY", "FLAG_MSA_OUT", "FLAG_KEEP_DISTMAT", "FLAG_INPUT_ORDER", "FLAG_KEEP_TREE", "FLAG_KIMURA", "FLAG_SCORES_I16", "ALPHABET_AUTO", "library_path", "load_library", "pair_index", "AlignmentTool",
           "B200Gotoh"]
__version__ = "0.2"
