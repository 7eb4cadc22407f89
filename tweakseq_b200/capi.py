"""ctypes binding of libtsqb200.so (include/tsq_b200.h) -- no compute happens in Python.

Fails loudly (``TsqError``) when the CUDA library is missing: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PROTEIN, NUCLEOTIDE = 0, 1
FLAG_FORCE_S32, FLAG_NO_DISTANCES, FLAG_NO_WAVE16, FLAG_IDENTITY, FLAG_MSA_OUT, FLAG_KEEP_DISTMAT, FLAG_INPUT_ORDER = 1, 2, 4, 8, 16, 32, 64
FLAG_KEEP_TREE = 128
FLAG_KIMURA = 256
FLAG_SCORES_I16 = 512
ALPHABET_AUTO = -1

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtsqb200.so")

STATUS = {0: "TSQ_OK", -1: "TSQ_ERR_INVALID", -2: "TSQ_ERR_NO_DEVICE", -3: "TSQ_ERR_CUDA",
          -4: "TSQ_ERR_NOMEM", -5: "TSQ_ERR_CANCELLED", -6: "TSQ_ERR_STATE", -7: "TSQ_ERR_IO",
          -8: "TSQ_ERR_MATRIX", -9: "TSQ_ERR_RANGE"}


class TsqError(RuntimeError):
    def __init__(self, status: int, message: str = ""):
        self.status = status
        super().__init__(f"{STATUS.get(status, status)}: {message}" if message else STATUS.get(status, str(status)))


class Params(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("alphabet", C.c_int32), ("gap_open", C.c_int32),
                ("gap_extend", C.c_int32), ("matrix", C.POINTER(C.c_int8)), ("device", C.c_int32),
                ("part_rank", C.c_int32), ("part_world", C.c_int32), ("flags", C.c_uint32),
                ("n_devices", C.c_int32)]


class Merge(C.Structure):
    _fields_ = [("left", C.c_uint32), ("right", C.c_uint32), ("height", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("n_sequences", C.c_uint64), ("n_pairs", C.c_uint64), ("cells", C.c_uint64),
                ("cells_s16", C.c_uint64), ("cells_s32", C.c_uint64), ("kernel_ms", C.c_double),
                ("upload_ms", C.c_double), ("download_ms", C.c_double), ("gcups_kernel", C.c_double),
                ("launches", C.c_uint32), ("sm_count", C.c_uint32), ("strip_width", C.c_uint32),
                ("upload_launches", C.c_uint32), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("tree_ms", C.c_double), ("msa_ms", C.c_double), ("encode_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Limits(C.Structure):
    _fields_ = [("max_len_packed", C.c_uint32), ("max_len_inter", C.c_uint32), ("inter_is_32bit", C.c_int32),
                ("wave_packed", C.c_int32), ("wave_window", C.c_int64), ("wave_window_max", C.c_int64),
                ("delta", C.c_int32), ("bias_at_limit", C.c_int32)]


class PipeRates(C.Structure):
    _fields_ = [("dpx_per_clk_sm", C.c_double), ("issue_per_clk_sm", C.c_double),
                ("mix_packed_cells_per_clk_sm", C.c_double), ("sm_mhz", C.c_double)]


PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_char_p)
LOG_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)

# every symbol include/tsq_b200.h declares (tests check the library exports each one)
SYMBOLS = ["tsq_version", "tsq_version_string", "tsq_status_string", "tsq_device_count", "tsq_default_params",
           "tsq_create", "tsq_destroy", "tsq_last_error", "tsq_set_sequences", "tsq_set_sequences_flat", "tsq_upload", "tsq_compute",
           "tsq_download", "tsq_set_stream", "tsq_synchronize", "tsq_run", "tsq_scores", "tsq_distances",
           "tsq_self_scores", "tsq_identities", "tsq_device_scores", "tsq_partition", "tsq_finalize", "tsq_device_results",
           "tsq_get_stats", "tsq_measure_dpx_rate", "tsq_run_fasta", "tsq_plan_partition", "tsq_guide_tree",
           "tsq_write_newick", "tsq_consensus", "tsq_align_pair", "tsq_partition_of", "tsq_msa", "tsq_write_msa_fasta", "tsq_write_distmat",
           "tsq_device_slab", "tsq_results_sharded", "tsq_set_result_buffers", "tsq_get_device_stats", "tsq_get_limits", "tsq_measure_pipe_rates", "tsq_detect_alphabet", "tsq_stream_results", "tsq_encode", "tsq_scores16"]

_lib = None


def library_path() -> str:
    return _SO


def build_library(force: bool = False) -> str:
    """Compile csrc/ into libtsqb200.so with nvcc (sm_100a).  Used by __graft_entry__.build()."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return _SO


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise TsqError(-2, f"{_SO} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback")
    L = C.CDLL(_SO)
    vp, i32p, u64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
    L.tsq_version.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tsq_version_string.restype = C.c_char_p
    L.tsq_status_string.restype = C.c_char_p
    L.tsq_status_string.argtypes = [C.c_int]
    L.tsq_default_params.argtypes = [C.POINTER(Params)]
    L.tsq_default_params.restype = None
    L.tsq_create.argtypes = [C.POINTER(vp), C.POINTER(Params)]
    L.tsq_destroy.argtypes = [vp]
    L.tsq_last_error.argtypes = [vp]
    L.tsq_last_error.restype = C.c_char_p
    L.tsq_set_sequences.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_uint32]
    L.tsq_set_sequences_flat.argtypes = [vp, C.c_char_p, u64p, C.c_uint32]
    for f in ("tsq_upload", "tsq_compute", "tsq_download", "tsq_synchronize", "tsq_finalize"):
        getattr(L, f).argtypes = [vp]
    L.tsq_set_stream.argtypes = [vp, vp]
    L.tsq_stream_results.argtypes = [vp, C.c_int]
    L.tsq_run.argtypes = [vp, PROGRESS_CB, vp, C.POINTER(C.c_int)]
    L.tsq_scores.argtypes = [vp, C.POINTER(i32p), u64p]
    L.tsq_scores16.argtypes = [vp, C.POINTER(C.POINTER(C.c_int16)), u64p]
    L.tsq_distances.argtypes = [vp, C.POINTER(C.POINTER(C.c_double)), u64p]
    L.tsq_self_scores.argtypes = [vp, C.POINTER(i32p), C.POINTER(C.c_uint32)]
    L.tsq_identities.argtypes = [vp, C.POINTER(i32p), u64p]
    L.tsq_device_scores.argtypes = [vp, C.POINTER(vp), u64p]
    L.tsq_partition.argtypes = [vp, u64p, u64p]
    L.tsq_partition_of.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.tsq_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), u64p]
    L.tsq_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.tsq_get_device_stats.argtypes = [vp, C.c_int32, C.POINTER(Stats)]
    L.tsq_get_limits.argtypes = [vp, C.POINTER(Limits)]
    L.tsq_detect_alphabet.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_uint32]
    L.tsq_encode.argtypes = [C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    L.tsq_measure_pipe_rates.argtypes = [vp, C.POINTER(PipeRates)]
    L.tsq_device_slab.argtypes = [vp, C.POINTER(vp), u64p, u64p]
    L.tsq_results_sharded.argtypes = [vp, C.POINTER(C.c_int)]
    L.tsq_set_result_buffers.argtypes = [vp, vp, vp, C.c_uint64]
    L.tsq_guide_tree.argtypes = [vp, C.POINTER(C.POINTER(Merge)), C.POINTER(C.c_uint32)]
    L.tsq_write_newick.argtypes = [vp, C.POINTER(C.c_char_p), C.c_char_p]
    L.tsq_consensus.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_double, C.c_char_p]
    L.tsq_align_pair.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_int32)]
    L.tsq_msa.argtypes = [vp, C.POINTER(C.POINTER(C.c_char)), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                          C.POINTER(C.POINTER(C.c_uint32))]
    L.tsq_write_msa_fasta.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_char_p, C.c_int]
    L.tsq_write_distmat.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_double)]
    L.tsq_plan_partition.argtypes = [C.POINTER(Params), C.POINTER(C.c_uint32), C.c_uint32, C.c_int32, u64p, u64p]
    L.tsq_measure_dpx_rate.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tsq_run_fasta.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Params), LOG_CB, vp, C.POINTER(C.c_int)]
    _lib = L
    return L


def pair_index(i: int, j: int, n: int) -> int:
    """Packed upper-triangle index of (i, j), i < j (include/tsq_b200.h)."""
    return i * n - i * (i + 1) // 2 + (j - i - 1)


def flatten(seqs):
    """list of str/bytes -> (contiguous bytes, uint64 offsets[n+1]) for set_sequences_flat."""
    raw = [s.encode("latin-1", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(raw) + 1, dtype=np.uint64)
    if raw:
        offs[1:] = np.cumsum([len(r) for r in raw], dtype=np.uint64)
    return b"".join(raw), offs


def detect_alphabet(seqs) -> int:
    """tsq_detect_alphabet(): NUCLEOTIDE when at least 90 % of the letters are ACGTUN, else PROTEIN (host only)."""
    L = load_library()
    raw = [s.encode("latin-1", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * max(len(raw), 1))(*raw) if raw else (C.c_char_p * 1)()
    lens = (C.c_uint32 * max(len(raw), 1))(*[len(r) for r in raw]) if raw else (C.c_uint32 * 1)()
    return L.tsq_detect_alphabet(arr, lens, len(raw))


def encode(seq, alphabet: int = PROTEIN) -> tuple[np.ndarray, int]:
    """tsq_encode(): the symbols tsq_set_sequences would store for `seq`, and the sum of S(x, x) under the
    alphabet's default matrix (host only)."""
    L = load_library()
    raw = seq.encode("latin-1", "replace") if isinstance(seq, str) else bytes(seq)
    out = np.empty(max(len(raw), 1), dtype=np.uint8)
    n, self_score = C.c_uint64(0), C.c_int64(0)
    rc = L.tsq_encode(alphabet, raw, len(raw), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n), C.byref(self_score))
    if rc != 0:
        raise TsqError(rc, "tsq_encode")
    return out[:n.value].copy(), int(self_score.value)


def plan_partition(lengths, world: int, **kw) -> list[tuple[int, int]]:
    """tsq_plan_partition(): per-rank [begin, end) slabs of the sorted packed index space (host only)."""
    L = load_library()
    p = Params()
    L.tsq_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    lens = np.ascontiguousarray(lengths, dtype=np.uint32)
    b = (C.c_uint64 * world)()
    e = (C.c_uint64 * world)()
    rc = L.tsq_plan_partition(C.byref(p), lens.ctypes.data_as(C.POINTER(C.c_uint32)), len(lens), world, b, e)
    if rc != 0:
        raise TsqError(rc, "tsq_plan_partition")
    return [(int(b[r]), int(e[r])) for r in range(world)]


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch/cupy can wrap library-owned memory."""

    def __init__(self, ptr: int, count: int, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}
        self._owner = owner


class Context:
    """One tsq_ctx.  Methods mirror the C ABI one to one."""

    def __init__(self, alphabet: int = PROTEIN, gap_open: int = -1, gap_extend: int = -1,
                 matrix: np.ndarray | None = None, device: int = 0, part_rank: int = 0,
                 part_world: int = 1, flags: int = 0, n_devices: int = 1):
        self._L = load_library()
        p = Params()
        self._L.tsq_default_params(C.byref(p))
        p.alphabet, p.gap_open, p.gap_extend = alphabet, gap_open, gap_extend
        p.device, p.part_rank, p.part_world, p.flags = device, part_rank, part_world, flags
        p.n_devices = n_devices
        self._matrix = None
        if matrix is not None:
            self._matrix = np.ascontiguousarray(matrix, dtype=np.int8)
            p.matrix = self._matrix.ctypes.data_as(C.POINTER(C.c_int8))
        self._h = C.c_void_p()
        rc = self._L.tsq_create(C.byref(self._h), C.byref(p))
        if rc != 0:
            self._h = None
            raise TsqError(rc, self._L.tsq_status_string(rc).decode())
        self.n = 0
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.tsq_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc: int):
        if rc != 0:
            raise TsqError(rc, self._L.tsq_last_error(self._h).decode())

    def set_sequences(self, seqs):
        """Per-sequence pointer form of the ABI (tsq_set_sequences)."""
        raw = [s.encode("latin-1", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
        n = len(raw)
        arr = (C.c_char_p * max(n, 1))(*raw) if n else (C.c_char_p * 1)()
        lens = (C.c_uint32 * max(n, 1))(*[len(r) for r in raw]) if n else (C.c_uint32 * 1)()
        self._keep = (raw, arr, lens)
        self._ck(self._L.tsq_set_sequences(self._h, arr, lens, n))
        self.n = n
        self._pair_capacity = 2 * max((len(r) for r in raw), default=0) + 1

    def set_sequences_flat(self, buf: bytes, offsets: np.ndarray):
        """One contiguous host buffer + n+1 offsets (tsq_set_sequences_flat)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        self._keep = (buf, offsets)
        self._ck(self._L.tsq_set_sequences_flat(self._h, buf, offsets.ctypes.data_as(C.POINTER(C.c_uint64)), n))
        self.n = n
        self._pair_capacity = 2 * int(np.diff(offsets).max(initial=0)) + 1

    def upload(self):
        self._ck(self._L.tsq_upload(self._h))

    def compute(self):
        self._ck(self._L.tsq_compute(self._h))

    def finalize(self):
        self._ck(self._L.tsq_finalize(self._h))

    def download(self):
        self._ck(self._L.tsq_download(self._h))

    def synchronize(self):
        self._ck(self._L.tsq_synchronize(self._h))

    def stream_results(self, enable: bool = True):
        """Staged calls: let compute() send finished row ranges to the host behind its launches (tsq_stream_results)."""
        self._ck(self._L.tsq_stream_results(self._h, 1 if enable else 0))

    def set_stream(self, cuda_stream_ptr: int | None):
        self._ck(self._L.tsq_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def run(self, progress=None, cancel: C.c_int | None = None):
        cb = PROGRESS_CB(lambda u, f, m: progress(f, m.decode() if m else "")) if progress else PROGRESS_CB(0)
        self._ck(self._L.tsq_run(self._h, cb, None, C.byref(cancel) if cancel is not None else None))

    @property
    def npairs(self) -> int:
        return self.n * (self.n - 1) // 2 if self.n >= 2 else 0

    def scores(self, copy: bool = True) -> np.ndarray:
        """Packed int32 scores (copy=False: a view of the library's buffer, valid until the next run -- what a
        20 GB result wants).  A rank of a sharded partition without result buffers sees its own slab only."""
        p, cnt = C.POINTER(C.c_int32)(), C.c_uint64()
        self._ck(self._L.tsq_scores(self._h, C.byref(p), C.byref(cnt)))
        if cnt.value == 0:
            return np.zeros(0, np.int32)
        a = np.ctypeslib.as_array(p, shape=(cnt.value,))
        return a.copy() if copy else a

    def scores16(self, copy: bool = True) -> np.ndarray:
        """Packed int16 scores of a context created with FLAG_SCORES_I16 (tsq_scores16)."""
        p, cnt = C.POINTER(C.c_int16)(), C.c_uint64()
        self._ck(self._L.tsq_scores16(self._h, C.byref(p), C.byref(cnt)))
        if cnt.value == 0:
            return np.zeros(0, np.int16)
        a = np.ctypeslib.as_array(p, shape=(cnt.value,))
        return a.copy() if copy else a

    def distances(self, copy: bool = True) -> np.ndarray:
        p, cnt = C.POINTER(C.c_double)(), C.c_uint64()
        self._ck(self._L.tsq_distances(self._h, C.byref(p), C.byref(cnt)))
        if cnt.value == 0:
            return np.zeros(0, np.float64)
        a = np.ctypeslib.as_array(p, shape=(cnt.value,))
        return a.copy() if copy else a

    def identities(self) -> np.ndarray:
        """FLAG_IDENTITY: identical residue pairs on the chosen optimal alignment, packed like scores()."""
        p, cnt = C.POINTER(C.c_int32)(), C.c_uint64()
        self._ck(self._L.tsq_identities(self._h, C.byref(p), C.byref(cnt)))
        if cnt.value == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(p, shape=(cnt.value,)).copy()

    def self_scores(self) -> np.ndarray:
        p, n = C.POINTER(C.c_int32)(), C.c_uint32()
        self._ck(self._L.tsq_self_scores(self._h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def partition(self) -> tuple[int, int]:
        b, e = C.c_uint64(), C.c_uint64()
        self._ck(self._L.tsq_partition(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def partition_of(self, rank: int) -> tuple[int, int]:
        """[begin, end) slab of any rank of this context's partition (no communication needed)."""
        b, e = C.c_uint64(), C.c_uint64()
        self._ck(self._L.tsq_partition_of(self._h, rank, C.byref(b), C.byref(e)))
        return b.value, e.value

    def device_scores(self) -> _DevArray:
        """Sorted-order packed int32 score buffer on the device (for the NCCL gather)."""
        d, cnt = C.c_void_p(), C.c_uint64()
        self._ck(self._L.tsq_device_scores(self._h, C.byref(d), C.byref(cnt)))
        return _DevArray(d.value or 0, cnt.value, "<i4", self)

    def device_slab(self):
        """(device array, first packed index) of the sorted-order scores this context really holds: the whole
        triangle, or a partitioned rank's own slab (tsq_device_slab)."""
        d, first, cnt = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._ck(self._L.tsq_device_slab(self._h, C.byref(d), C.byref(first), C.byref(cnt)))
        return _DevArray(d.value or 0, cnt.value, "<i4", self), first.value

    def results_sharded(self) -> bool:
        """True: every rank of the partition finalizes and downloads its own slab (fixed-length input);
        False: the slabs must be gathered into rank 0 first (tsq_results_sharded)."""
        v = C.c_int()
        self._ck(self._L.tsq_results_sharded(self._h, C.byref(v)))
        return bool(v.value)

    def set_result_buffers(self, scores: np.ndarray | None, distances: np.ndarray | None = None):
        """Caller-owned host arrays (int32 / float64, n*(n-1)/2 each) that receive the results, e.g. views of one
        shared-memory segment mapped by every rank (tsq_set_result_buffers).  None restores the library's own."""
        if scores is None:
            self._ext = None
            self._ck(self._L.tsq_set_result_buffers(self._h, None, None, 0))
            return
        assert scores.dtype == np.int32 and scores.flags.c_contiguous
        assert distances is None or (distances.dtype == np.float64 and distances.flags.c_contiguous and len(distances) == len(scores))
        self._ext = (scores, distances)
        self._ck(self._L.tsq_set_result_buffers(self._h, scores.ctypes.data, distances.ctypes.data if distances is not None else None,
                                                len(scores)))

    def device_stats(self, index: int) -> dict:
        """Figures of one device of a multi-device context (tsq_get_device_stats)."""
        st = Stats()
        self._ck(self._L.tsq_get_device_stats(self._h, index, C.byref(st)))
        return st.as_dict()

    def device_results(self):
        ds, dd, cnt = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._ck(self._L.tsq_device_results(self._h, C.byref(ds), C.byref(dd), C.byref(cnt)))
        s = _DevArray(ds.value or 0, cnt.value, "<i4", self)
        d = _DevArray(dd.value, cnt.value, "<f8", self) if dd.value else None
        return s, d

    def guide_tree(self):
        """UPGMA merges of the last run: (left[n-1], right[n-1], height[n-1]) numpy arrays."""
        p, cnt = C.POINTER(Merge)(), C.c_uint32()
        self._ck(self._L.tsq_guide_tree(self._h, C.byref(p), C.byref(cnt)))
        k = cnt.value
        left = np.array([p[i].left for i in range(k)], dtype=np.uint32)
        right = np.array([p[i].right for i in range(k)], dtype=np.uint32)
        height = np.array([p[i].height for i in range(k)], dtype=np.float64)
        return left, right, height

    def write_newick(self, path: str, labels=None):
        arr = None
        if labels is not None:
            raw = [l.encode() for l in labels]
            arr = (C.c_char_p * len(raw))(*raw)
        self._ck(self._L.tsq_write_newick(self._h, arr, path.encode()))

    def msa(self):
        """Progressive multiple alignment along the guide tree of the last run: (rows, tree_order) --
        equal-length gapped strings in submitted order and the leaf order of the tree."""
        rows, order = C.POINTER(C.c_char)(), C.POINTER(C.c_uint32)()
        n, cols = C.c_uint32(), C.c_uint32()
        self._ck(self._L.tsq_msa(self._h, C.byref(rows), C.byref(n), C.byref(cols), C.byref(order)))
        raw = C.string_at(rows, n.value * cols.value) if n.value * cols.value else b""
        out = [raw[r * cols.value:(r + 1) * cols.value].decode("ascii") for r in range(n.value)]
        return out, [int(order[i]) for i in range(n.value)]

    def write_msa_fasta(self, path: str, headers=None, residues=None, tree_order: bool = True):
        """tsq_write_msa_fasta(): the alignment as FASTA; residues = the submitted spelling to echo."""
        harr = rarr = larr = None
        if headers is not None:
            hraw = [h.encode("latin-1", "replace") for h in headers]
            harr = (C.c_char_p * max(len(hraw), 1))(*hraw)
        if residues is not None:
            rraw = [s.encode("latin-1", "replace") if isinstance(s, str) else bytes(s) for s in residues]
            rarr = (C.c_char_p * max(len(rraw), 1))(*rraw)
            larr = (C.c_uint32 * max(len(rraw), 1))(*[len(r) for r in rraw])
        self._ck(self._L.tsq_write_msa_fasta(self._h, harr, rarr, larr, path.encode(), 1 if tree_order else 0))

    def consensus(self, rows, plurality: float = -1.0) -> str:
        """Consensus annotation of equal-length aligned rows (Consensus.cpp:80-161); '?' = no plurality."""
        raw = [r.encode("latin-1", "replace") if isinstance(r, str) else bytes(r) for r in rows]
        ncols = len(raw[0]) if raw else 0
        if any(len(r) != ncols for r in raw):
            raise ValueError("alignment rows differ in length")
        arr = (C.c_char_p * max(len(raw), 1))(*raw) if raw else (C.c_char_p * 1)()
        out = C.create_string_buffer(ncols + 1)
        self._ck(self._L.tsq_consensus(self._h, arr, len(raw), ncols, plurality, out))
        return out.raw[:ncols].decode("latin-1")

    def align_pair(self, i: int, j: int) -> tuple[str, str, int]:
        """One optimal global alignment of submitted sequences i and j: (row_i, row_j, score)."""
        cap = self._pair_capacity
        a, b = C.create_string_buffer(cap), C.create_string_buffer(cap)
        cols, score = C.c_uint32(), C.c_int32()
        self._ck(self._L.tsq_align_pair(self._h, i, j, a, b, cap, C.byref(cols), C.byref(score)))
        return a.raw[:cols.value].decode("ascii"), b.raw[:cols.value].decode("ascii"), score.value

    def stats(self) -> dict:
        st = Stats()
        self._ck(self._L.tsq_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def limits(self) -> dict:
        """The length limits that select a kernel under this context's matrix and gap model (tsq_get_limits)."""
        lim = Limits()
        self._ck(self._L.tsq_get_limits(self._h, C.byref(lim)))
        return {k: getattr(lim, k) for k, _ in lim._fields_}

    def measure_pipe_rates(self) -> dict:
        """Live roofline denominators: DPX rate, issue ceiling, the inner loop's mix in isolation (tsq_measure_pipe_rates)."""
        r = PipeRates()
        self._ck(self._L.tsq_measure_pipe_rates(self._h, C.byref(r)))
        return {k: getattr(r, k) for k, _ in r._fields_}

    def measure_dpx_rate(self) -> tuple[float, float]:
        ops, mhz = C.c_double(), C.c_double()
        self._ck(self._L.tsq_measure_dpx_rate(self._h, C.byref(ops), C.byref(mhz)))
        return ops.value, mhz.value


def write_distmat(path: str, labels, packed) -> None:
    """tsq_write_distmat(): PHYLIP-style square matrix file from packed upper-triangle distances (host only)."""
    L = load_library()
    raw = [l.encode("latin-1", "replace") for l in labels]
    arr = (C.c_char_p * max(len(raw), 1))(*raw) if raw else (C.c_char_p * 1)()
    d = np.ascontiguousarray(packed, dtype=np.float64)
    rc = L.tsq_write_distmat(path.encode(), arr, len(raw), d.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise TsqError(rc, "tsq_write_distmat")


def run_fasta(fasta_in: str, distmat_out: str, log=None, cancel: C.c_int | None = None, **kw) -> int:
    """tsq_run_fasta(): FASTA file in, PHYLIP-style distance matrix out (flags=FLAG_MSA_OUT: the
    multiple alignment out, matrix in <out>.distmat).  Returns the status."""
    L = load_library()
    p = Params()
    L.tsq_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    cb = LOG_CB(lambda u, m: log(m.decode())) if log else LOG_CB(0)
    return L.tsq_run_fasta(fasta_in.encode(), distmat_out.encode(), C.byref(p), cb, None,
                           C.byref(cancel) if cancel is not None else None)
