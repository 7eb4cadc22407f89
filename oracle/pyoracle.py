"""ctypes binding of the CPU oracle (oracle/gotoh_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (tweakseq_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtsq_oracle.so")

PROTEIN, NUCLEOTIDE = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gotoh_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtsq_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_REF_SO = os.path.join(_HERE, "_ref", "libref_consensus.so")
_ref = None


def ref_consensus_available() -> bool:
    """oracle/_ref/libref_consensus.so: the reference's OWN Consensus.cpp, compiled where it lies by
    oracle/Makefile (only possible where /root/reference exists; the built file travels with the snapshot)."""
    if not os.path.exists(_REF_SO) and os.path.exists("/root/reference/tweakseq/Core/Annotations/Consensus.cpp"):
        subprocess.call(["make", "-C", _HERE, "_ref/libref_consensus.so"], stdout=subprocess.DEVNULL)
    return os.path.exists(_REF_SO)


_REF_FASTA_SO = os.path.join(_HERE, "_ref", "libref_fasta.so")
_ref_fasta = None


def ref_fasta_available() -> bool:
    """oracle/_ref/libref_fasta.so: the reference's OWN FASTAFile.cpp + SequenceFile.cpp, compiled where they lie."""
    if not os.path.exists(_REF_FASTA_SO) and os.path.exists("/root/reference/tweakseq/Core/FASTAFile.cpp"):
        subprocess.call(["make", "-C", _HERE, "_ref/libref_fasta.so"], stdout=subprocess.DEVNULL)
    return os.path.exists(_REF_FASTA_SO)


def _ref_fasta_lib():
    global _ref_fasta
    if _ref_fasta is None:
        _ref_fasta = C.CDLL(_REF_FASTA_SO)
        _ref_fasta.tsq_ref_fasta_read.restype = C.c_int
        _ref_fasta.tsq_ref_fasta_read.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_ulong]
        _ref_fasta.tsq_ref_fasta_write.restype = C.c_int
        _ref_fasta.tsq_ref_fasta_write.argtypes = [C.c_char_p] * 4
    return _ref_fasta


def ref_fasta_read(path: str):
    """FASTAFile::read of the reference itself (tweakseq/Core/FASTAFile.cpp:71-147): (labels, sequences, comments)."""
    cap = max(os.path.getsize(path) * 2, 1 << 12) + 64
    bufs = [C.create_string_buffer(cap) for _ in range(3)]
    n = _ref_fasta_lib().tsq_ref_fasta_read(path.encode(), bufs[0], bufs[1], bufs[2], cap)
    if n < 0:
        raise RuntimeError(f"tsq_ref_fasta_read: {n}")
    out = []
    for b in bufs:
        text = b.value.decode("latin-1")
        out.append(text.split("\n")[:-1] if text else [])
    return tuple(out)


def ref_fasta_write(path: str, labels, seqs, comments) -> None:
    """FASTAFile::write of the reference itself (tweakseq/Core/FASTAFile.cpp:149-171)."""
    j = lambda l: ("".join(x + "\n" for x in l)).encode("latin-1")
    if _ref_fasta_lib().tsq_ref_fasta_write(path.encode(), j(labels), j(seqs), j(comments)) != 0:
        raise RuntimeError("tsq_ref_fasta_write failed")


_REF_SEQ_SO = os.path.join(_HERE, "_ref", "libref_sequence.so")
_ref_seq = None


def ref_sequence_available() -> bool:
    """oracle/_ref/libref_sequence.so: the reference's OWN Sequence.cpp (Sequence::filter), compiled where it lies."""
    if not os.path.exists(_REF_SEQ_SO) and os.path.exists("/root/reference/tweakseq/Core/Sequence.cpp"):
        subprocess.call(["make", "-C", _HERE, "_ref/libref_sequence.so"], stdout=subprocess.DEVNULL)
    return os.path.exists(_REF_SEQ_SO)


def ref_filter(cells, apply_exclusions: bool) -> list[int]:
    """Sequence::filter of the reference itself (tweakseq/Core/Sequence.cpp:57-69) on 16-bit residue cells."""
    global _ref_seq
    if _ref_seq is None:
        _ref_seq = C.CDLL(_REF_SEQ_SO)
        _ref_seq.tsq_ref_filter.restype = C.c_int
        _ref_seq.tsq_ref_filter.argtypes = [C.POINTER(C.c_uint16), C.c_uint, C.c_int, C.POINTER(C.c_uint16)]
    a = np.ascontiguousarray(np.array(list(cells), dtype=np.uint16))
    out = np.zeros(max(len(a), 1), dtype=np.uint16)
    k = _ref_seq.tsq_ref_filter(a.ctypes.data_as(C.POINTER(C.c_uint16)) if len(a) else out.ctypes.data_as(C.POINTER(C.c_uint16)),
                                len(a), 1 if apply_exclusions else 0, out.ctypes.data_as(C.POINTER(C.c_uint16)))
    return [int(v) for v in out[:k]]


_REF_TOOLS_SO = os.path.join(_HERE, "_ref", "libref_tools.so")
_ref_tools = None
REF_TOOLS = {"clustalo": 0, "muscle": 1, "mafft": 2}


def ref_tools_available() -> bool:
    """oracle/_ref/libref_tools.so: the reference's OWN ClustalO / Muscle / MAFFT wrappers, compiled where they lie."""
    if not os.path.exists(_REF_TOOLS_SO) and os.path.exists("/root/reference/tweakseq/Core/ClustalO.cpp"):
        subprocess.call(["make", "-C", _HERE, "_ref/libref_tools.so"], stdout=subprocess.DEVNULL)
    return os.path.exists(_REF_TOOLS_SO)


def _ref_tools_lib():
    global _ref_tools
    if _ref_tools is None:
        _ref_tools = C.CDLL(_REF_TOOLS_SO)
        _ref_tools.tsq_ref_tool_command.restype = C.c_int
        _ref_tools.tsq_ref_tool_command.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_ulong,
                                                    C.POINTER(C.c_int)]
        _ref_tools.tsq_ref_tool_version.restype = C.c_int
        _ref_tools.tsq_ref_tool_version.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_ulong]
    return _ref_tools


def ref_tool_command(tool: str, fin: str, fout: str, exe: str = ""):
    """What the reference's own wrapper hands to QProcess::start: (name, executable, argv list, uses_stdout)."""
    buf = C.create_string_buffer(1 << 14)
    so = C.c_int()
    rc = _ref_tools_lib().tsq_ref_tool_command(REF_TOOLS[tool], exe.encode(), fin.encode(), fout.encode(), buf, len(buf), C.byref(so))
    if rc != 0:
        raise RuntimeError(f"tsq_ref_tool_command: {rc}")
    parts = buf.value.decode().split("\n")[:-1]
    return parts[0], parts[1], parts[2:], bool(so.value)


def ref_tool_version(tool: str, exe: str) -> str:
    """The wrapper's own getVersion() (a real fork + exec of `exe` with its version flag), as version() reports it."""
    buf = C.create_string_buffer(1 << 12)
    rc = _ref_tools_lib().tsq_ref_tool_version(REF_TOOLS[tool], exe.encode(), buf, len(buf))
    if rc != 0:
        raise RuntimeError(f"tsq_ref_tool_version: {rc}")
    return buf.value.decode()


_REF_QT_SO = os.path.join(_HERE, "_ref", "libref_qt_adapter.so")
_ref_qt = None


def ref_qt_adapter_available() -> bool:
    """oracle/_ref/libref_qt_adapter.so: host/qt/B200GotohTool.cpp (the real Qt adapter) compiled against the
    reference's AlignmentTool.h / ClustalO.cpp and the functional Qt stand-ins; links libtsqb200.so."""
    if not os.path.exists(_REF_QT_SO) and os.path.exists("/root/reference/tweakseq/Core/ClustalO.cpp"):
        subprocess.call(["make", "-C", _HERE, "_ref/libref_qt_adapter.so"], stdout=subprocess.DEVNULL)
    return os.path.exists(_REF_QT_SO)


def _ref_qt_lib():
    global _ref_qt
    if _ref_qt is None:
        _ref_qt = C.CDLL(_REF_QT_SO)
        _ref_qt.tsq_qt_settings_round_trip.argtypes = [C.c_char_p, C.c_ulong]
        _ref_qt.tsq_qt_worker_run.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                              C.c_char_p, C.c_ulong]
        _ref_qt.tsq_qt_worker_run_in_memory.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                        C.c_char_p, C.c_ulong]
    return _ref_qt


def qt_settings_round_trip() -> dict:
    """ClustalO (reference) and B200GotohTool (host/qt) write into one settings element and read it back."""
    buf = C.create_string_buffer(1 << 14)
    if _ref_qt_lib().tsq_qt_settings_round_trip(buf, len(buf)) != 0:
        raise RuntimeError("tsq_qt_settings_round_trip failed")
    return dict(line.split("=", 1) for line in buf.value.decode().splitlines())


def qt_worker_run(fin: str, fout: str, align_in_process: bool = True):
    """B200GotohWorker::start() as startAlignment would use it: (exit code, exit status, log lines)."""
    buf = C.create_string_buffer(1 << 16)
    ec, es = C.c_int(), C.c_int()
    if _ref_qt_lib().tsq_qt_worker_run(fin.encode(), fout.encode(), 1 if align_in_process else 0, C.byref(ec), C.byref(es),
                                       buf, len(buf)) != 0:
        raise RuntimeError("tsq_qt_worker_run failed")
    return ec.value, es.value, buf.value.decode().splitlines()


def qt_worker_run_in_memory(labels, residues, fout: str):
    """The in-memory worker: (label, Sequence::filter(true)) pairs as startAlignment takes them from the model;
    fout receives the alignment with ">label" headers.  (exit code, exit status, log lines)."""
    buf = C.create_string_buffer(1 << 16)
    ec, es = C.c_int(), C.c_int()
    lab = "".join(l + "\n" for l in labels).encode()
    res = "".join(r + "\n" for r in residues).encode()
    if _ref_qt_lib().tsq_qt_worker_run_in_memory(lab, res, fout.encode(), C.byref(ec), C.byref(es), buf, len(buf)) != 0:
        raise RuntimeError("tsq_qt_worker_run_in_memory failed")
    return ec.value, es.value, buf.value.decode().splitlines()


def ref_consensus(cell_rows, plurality: float = -1.0) -> str:
    """Consensus::calculate of the reference itself (tweakseq/Core/Annotations/Consensus.cpp:80-161).
    cell_rows: equal-length sequences of 16-bit residue cells (str, bytes or ints, flag bits allowed)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_SO)
        _ref.tsq_ref_consensus.restype = C.c_int
        _ref.tsq_ref_consensus.argtypes = [C.POINTER(C.c_uint16), C.c_uint, C.c_uint, C.c_double, C.c_char_p]
    rows = [[ord(ch) if isinstance(ch, str) else int(ch) for ch in r] for r in cell_rows]
    ncols = len(rows[0]) if rows else 0
    cells = np.ascontiguousarray(np.array(rows, dtype=np.uint16).reshape(len(rows), ncols))
    out = C.create_string_buffer(ncols + 1)
    rc = _ref.tsq_ref_consensus(cells.ctypes.data_as(C.POINTER(C.c_uint16)), len(rows), ncols, plurality, out)
    if rc != 0:
        raise RuntimeError(f"tsq_ref_consensus: {rc}")
    return out.raw[:ncols].decode("latin-1")


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, u32p, u64p, i32p, i8p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_int8))
        L.tsq_oracle_nsym.restype = C.c_int
        L.tsq_oracle_matrix.restype = i8p
        L.tsq_oracle_encode.restype = C.c_size_t
        L.tsq_oracle_encode.argtypes = [C.c_int, C.c_char_p, C.c_size_t, u8p]
        L.tsq_oracle_gotoh.restype = C.c_int32
        L.tsq_oracle_gotoh.argtypes = [u8p, C.c_int, u8p, C.c_int, i8p, C.c_int, C.c_int, C.c_int]
        L.tsq_oracle_self_score.restype = C.c_int32
        L.tsq_oracle_self_score.argtypes = [u8p, C.c_int, i8p, C.c_int]
        L.tsq_oracle_distance.restype = C.c_double
        L.tsq_oracle_distance.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.tsq_oracle_pair_index.restype = C.c_uint64
        L.tsq_oracle_pair_index.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.tsq_oracle_all_pairs.restype = C.c_uint64
        L.tsq_oracle_all_pairs.argtypes = [u8p, u64p, u32p, C.c_uint32, i8p, C.c_int, C.c_int, C.c_int,
                                           C.c_uint64, C.c_uint64, i32p, C.c_int]
        L.tsq_oracle_pair_list.restype = C.c_uint64
        L.tsq_oracle_pair_list.argtypes = [u8p, u64p, u32p, i8p, C.c_int, C.c_int, C.c_int, u32p, u32p,
                                           C.c_uint64, i32p, C.c_int]
        L.tsq_oracle_gotoh_id.restype = None
        L.tsq_oracle_gotoh_id.argtypes = [u8p, C.c_int, u8p, C.c_int, i8p, C.c_int, C.c_int, C.c_int, i32p, i32p]
        L.tsq_oracle_consensus.restype = None
        L.tsq_oracle_consensus.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_double, C.c_char_p]
        L.tsq_oracle_traceback.restype = C.c_uint32
        L.tsq_oracle_traceback.argtypes = [u8p, C.c_int, u8p, C.c_int, i8p, C.c_int, C.c_int, C.c_int, u8p, u8p, i32p]
        L.tsq_oracle_upgma.restype = None
        L.tsq_oracle_upgma.argtypes = [C.POINTER(C.c_double), C.c_uint32, u32p, u32p, C.POINTER(C.c_double)]
        L.tsq_oracle_msa.restype = C.c_int
        L.tsq_oracle_msa.argtypes = [u8p, u64p, u32p, C.c_uint32, i8p, C.c_int, C.c_int, C.c_int, u32p, u32p,
                                     C.POINTER(u8p), u32p, C.POINTER(C.c_int64)]
        L.tsq_oracle_free.restype = None
        L.tsq_oracle_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def nsym(alphabet: int) -> int:
    return lib().tsq_oracle_nsym(alphabet)


def matrix(alphabet: int) -> np.ndarray:
    n = nsym(alphabet)
    lib().tsq_oracle_matrix.argtypes = [C.c_int]
    ptr = lib().tsq_oracle_matrix(alphabet)
    return np.ctypeslib.as_array(ptr, shape=(n * n,)).reshape(n, n).copy()


def encode(seq: str | bytes, alphabet: int = PROTEIN) -> np.ndarray:
    raw = seq.encode("latin-1", "replace") if isinstance(seq, str) else bytes(seq)
    out = np.empty(max(len(raw), 1), dtype=np.uint8)
    k = lib().tsq_oracle_encode(alphabet, raw, len(raw), _p(out, C.c_uint8))
    return out[:k].copy()


def gotoh(a: np.ndarray, b: np.ndarray, mat: np.ndarray, go: int, ge: int) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    return int(lib().tsq_oracle_gotoh(_p(a, C.c_uint8), len(a), _p(b, C.c_uint8), len(b), _p(m8, C.c_int8),
                                      m8.shape[0], go, ge))


def score_str(a: str, b: str, alphabet: int = PROTEIN, go: int | None = None, ge: int = 1) -> int:
    if go is None:
        go = 11 if alphabet == PROTEIN else 10
    return gotoh(encode(a, alphabet), encode(b, alphabet), matrix(alphabet), go, ge)


def self_score(a: np.ndarray, mat: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    return int(lib().tsq_oracle_self_score(_p(a, C.c_uint8), len(a), _p(m8, C.c_int8), m8.shape[0]))


def distance(sij: int, sii: int, sjj: int) -> float:
    return float(lib().tsq_oracle_distance(sij, sii, sjj))


def pair_index(i: int, j: int, n: int) -> int:
    return int(lib().tsq_oracle_pair_index(i, j, n))


def _pack(encoded: list[np.ndarray]):
    lens = np.array([len(s) for s in encoded], dtype=np.uint32)
    offs = np.zeros(len(encoded), dtype=np.uint64)
    if len(encoded) > 1:
        offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    flat = np.concatenate(encoded).astype(np.uint8) if lens.sum() else np.zeros(1, np.uint8)
    return np.ascontiguousarray(flat), offs, lens


def all_pairs(encoded: list[np.ndarray], mat: np.ndarray, go: int, ge: int, nthreads: int = 1,
              pair_begin: int = 0, pair_end: int | None = None):
    """Scores of packed pairs [pair_begin, pair_end); returns (int32 array, cells)."""
    n = len(encoded)
    total = n * (n - 1) // 2
    if pair_end is None:
        pair_end = total
    flat, offs, lens = _pack(encoded)
    out = np.zeros(max(pair_end - pair_begin, 0), dtype=np.int32)
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    cells = lib().tsq_oracle_all_pairs(_p(flat, C.c_uint8), _p(offs, C.c_uint64), _p(lens, C.c_uint32), n,
                                       _p(m8, C.c_int8), m8.shape[0], go, ge, pair_begin, pair_end,
                                       _p(out, C.c_int32), nthreads)
    return out, int(cells)


def pair_list(encoded: list[np.ndarray], pi: np.ndarray, pj: np.ndarray, mat: np.ndarray, go: int, ge: int,
              nthreads: int = 1):
    flat, offs, lens = _pack(encoded)
    pi = np.ascontiguousarray(pi, dtype=np.uint32)
    pj = np.ascontiguousarray(pj, dtype=np.uint32)
    out = np.zeros(len(pi), dtype=np.int32)
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    cells = lib().tsq_oracle_pair_list(_p(flat, C.c_uint8), _p(offs, C.c_uint64), _p(lens, C.c_uint32),
                                       _p(m8, C.c_int8), m8.shape[0], go, ge, _p(pi, C.c_uint32),
                                       _p(pj, C.c_uint32), len(pi), _p(out, C.c_int32), nthreads)
    return out, int(cells)


def rows_simd(encoded: list[np.ndarray], mat: np.ndarray, go: int, ge: int, nthreads: int = 1,
              row_begin: int = 0, row_end: int | None = None):
    """The SIMD CPU baseline kernel (gotoh_simd.c): all pairs (i, j), row_begin <= i < row_end, j > i.
    Returns (int32 scores of packed indices [packed(row_begin, row_begin+1), packed(row_end, row_end+1)), cells)."""
    n = len(encoded)
    if row_end is None or row_end > n - 1:
        row_end = max(n - 1, 0)
    flat, offs, lens = _pack(encoded)
    first = pair_index(row_begin, row_begin + 1, n) if row_begin + 1 < n else n * (n - 1) // 2
    last = pair_index(row_end, row_end + 1, n) if row_end + 1 < n else n * (n - 1) // 2
    out = np.zeros(max(last - first, 0), dtype=np.int32)
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    L = lib()
    L.tsq_oracle_rows_simd.restype = C.c_uint64
    L.tsq_oracle_rows_simd.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint32,
                                       C.POINTER(C.c_int8), C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                       C.POINTER(C.c_int32), C.c_int]
    cells = L.tsq_oracle_rows_simd(_p(flat, C.c_uint8), _p(offs, C.c_uint64), _p(lens, C.c_uint32), n, _p(m8, C.c_int8),
                                   m8.shape[0], go, ge, row_begin, row_end, _p(out, C.c_int32), nthreads)
    return out, int(cells)


def distances(scores: np.ndarray, selfs: np.ndarray) -> np.ndarray:
    """Packed fp64 distances from packed int32 scores and per-sequence self scores."""
    n = len(selfs)
    iu, ju = np.triu_indices(n, 1)
    mn = np.minimum(selfs[iu], selfs[ju]).astype(np.int64)
    d = np.ones(len(scores), dtype=np.float64)
    ok = mn > 0
    q = scores[ok].astype(np.float64) / mn[ok].astype(np.float64)
    d[ok] = 1.0 - q
    return d


def upgma(packed: np.ndarray, n: int):
    """UPGMA merges (left, right, height) of a packed fp64 distance matrix (oracle, O(n^3))."""
    d = np.ascontiguousarray(packed, dtype=np.float64)
    left = np.zeros(max(n - 1, 0), dtype=np.uint32)
    right = np.zeros(max(n - 1, 0), dtype=np.uint32)
    height = np.zeros(max(n - 1, 0), dtype=np.float64)
    if n >= 2:
        lib().tsq_oracle_upgma(_p(d, C.c_double), n, _p(left, C.c_uint32), _p(right, C.c_uint32), _p(height, C.c_double))
    return left, right, height


def newick(left, right, height, labels) -> str:
    """Newick text from UPGMA merges, same format as tsq_write_newick (branch lengths %.6f)."""
    n = len(labels)
    if n == 0:
        return ";"
    if n == 1:
        return labels[0] + ";"
    hs = lambda i: 0.0 if i < n else float(height[i - n])

    def rec(i, ph):
        if i < n:
            return "%s:%.6f" % (labels[i], ph - 0.0)
        l, r = int(left[i - n]), int(right[i - n])
        inner = "(" + rec(l, hs(i)) + "," + rec(r, hs(i)) + ")"
        return inner if ph < 0 else inner + ":%.6f" % (ph - hs(i))
    import sys
    sys.setrecursionlimit(max(10000, 4 * n))
    return rec(2 * n - 2, -1.0) + ";"


def gotoh_id(a: np.ndarray, b: np.ndarray, mat: np.ndarray, go: int, ge: int):
    """(score, identities): the Gotoh score and the most identities among the optimal alignments."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    sc, nid = C.c_int32(), C.c_int32()
    lib().tsq_oracle_gotoh_id(_p(a, C.c_uint8), len(a), _p(b, C.c_uint8), len(b), _p(m8, C.c_int8), m8.shape[0], go, ge,
                              C.byref(sc), C.byref(nid))
    return sc.value, nid.value


def traceback(a: np.ndarray, b: np.ndarray, mat: np.ndarray, go: int, ge: int, alphabet: int = PROTEIN):
    """(row_a, row_b, score): one optimal alignment as two gapped strings of canonical symbols."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    oa = np.empty(len(a) + len(b) + 1, dtype=np.uint8)
    ob = np.empty(len(a) + len(b) + 1, dtype=np.uint8)
    sc = C.c_int32()
    k = lib().tsq_oracle_traceback(_p(a, C.c_uint8), len(a), _p(b, C.c_uint8), len(b), _p(m8, C.c_int8), m8.shape[0],
                                   go, ge, _p(oa, C.c_uint8), _p(ob, C.c_uint8), C.byref(sc))
    letters = "ACGTN" if alphabet == NUCLEOTIDE else "ARNDCQEGHILKMFPSTWYVBZX"
    dec = lambda v: "".join("-" if x == 0xff else letters[x] for x in v[:k])
    return dec(oa), dec(ob), sc.value


def ln(x: float) -> float:
    """The spec's own logarithm (tsq_oracle_ln): fixed sequence of IEEE double operations."""
    L = lib()
    L.tsq_oracle_ln.restype = C.c_double
    L.tsq_oracle_ln.argtypes = [C.c_double]
    return L.tsq_oracle_ln(x)


def kimura(identities: int, min_len: int):
    """(Kimura-corrected identity distance, ok); ok False: D >= 0.75, the formula does not apply."""
    L = lib()
    L.tsq_oracle_kimura.restype = C.c_double
    L.tsq_oracle_kimura.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int)]
    ok = C.c_int()
    d = L.tsq_oracle_kimura(identities, min_len, C.byref(ok))
    return d, bool(ok.value)


def all_pairs_id(encoded, mat, go, ge):
    """Packed (scores, identities, clustalw distances) by the identity-aware oracle (single thread)."""
    n = len(encoded)
    sc, nid, d = [], [], []
    for i in range(n):
        for j in range(i + 1, n):
            s, k = gotoh_id(encoded[i], encoded[j], mat, go, ge)
            ml = min(len(encoded[i]), len(encoded[j]))
            sc.append(s); nid.append(k); d.append(1.0 - k / ml if ml > 0 else 1.0)
    return np.array(sc, np.int32), np.array(nid, np.int32), np.array(d, np.float64)


def consensus(rows, plurality: float | None = None) -> str:
    """Consensus::calculate restated (oracle, O(cols * rows^2)); plurality defaults to rows/2."""
    raw = [r.encode("latin-1", "replace") if isinstance(r, str) else bytes(r) for r in rows]
    ncols = len(raw[0]) if raw else 0
    arr = (C.c_char_p * max(len(raw), 1))(*raw) if raw else (C.c_char_p * 1)()
    out = C.create_string_buffer(ncols + 1)
    lib().tsq_oracle_consensus(arr, len(raw), ncols, len(raw) / 2.0 if plurality is None else plurality, out)
    return out.raw[:ncols].decode("latin-1")


def msa(encoded, mat, go: int, ge: int, left, right, alphabet: int = PROTEIN):
    """Progressive alignment along the merges (left, right): (rows as gapped strings in submitted
    order, int64 profile-alignment score of every merge)."""
    n = len(encoded)
    flat, offs, lens = _pack(encoded) if n else (np.zeros(1, np.uint8), np.zeros(1, np.uint64), np.zeros(1, np.uint32))
    m8 = np.ascontiguousarray(mat, dtype=np.int8)
    left = np.ascontiguousarray(left, dtype=np.uint32)
    right = np.ascontiguousarray(right, dtype=np.uint32)
    lp = left if len(left) else np.zeros(1, np.uint32)
    rp = right if len(right) else np.zeros(1, np.uint32)
    sc = np.zeros(max(n - 1, 1), dtype=np.int64)
    rows, ncols = C.POINTER(C.c_uint8)(), C.c_uint32()
    rc = lib().tsq_oracle_msa(_p(flat, C.c_uint8), _p(offs, C.c_uint64), _p(lens, C.c_uint32), n, _p(m8, C.c_int8),
                              m8.shape[0], go, ge, _p(lp, C.c_uint32), _p(rp, C.c_uint32), C.byref(rows),
                              C.byref(ncols), _p(sc, C.c_int64))
    assert rc == 0
    letters = "ACGTN" if alphabet == NUCLEOTIDE else "ARNDCQEGHILKMFPSTWYVBZX"
    out = []
    if n:
        arr = np.ctypeslib.as_array(rows, shape=(n * max(ncols.value, 1),))
        for r in range(n):
            v = arr[r * ncols.value:(r + 1) * ncols.value]
            out.append("".join("-" if x == 0xff else letters[x] for x in v))
        lib().tsq_oracle_free(rows)
    return out, sc[:max(n - 1, 0)].copy()
