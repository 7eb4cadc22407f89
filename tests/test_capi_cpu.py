"""CPU-side checks of the C ABI: the library builds, loads, exports every symbol the header
declares, and fails loudly (no fallback) without a B200.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tweakseq_b200 as t
from tweakseq_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tsq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsq_[a-z0-9_]+)\s*\(", src)) - {"tsq_progress_cb", "tsq_log_cb"})


def test_library_exports_every_declared_symbol():
    L = t.load_library()
    syms = _header_symbols()
    assert len(syms) >= 25
    assert sorted(capi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(L, s), s


def test_version_and_status_strings():
    L = t.load_library()
    ma, mi = C.c_int(), C.c_int()
    assert L.tsq_version(C.byref(ma), C.byref(mi)) == 0
    assert (ma.value, mi.value) == (0, 2)
    assert b"sm_100a" in L.tsq_version_string()
    assert L.tsq_status_string(0) == b"ok"
    assert L.tsq_status_string(-2) == b"no sm_100 CUDA device"


def test_default_params():
    p = capi.Params()
    t.load_library().tsq_default_params(C.byref(p))
    assert p.struct_size == C.sizeof(capi.Params)
    assert (p.alphabet, p.gap_open, p.gap_extend, p.part_rank, p.part_world, p.flags) == (0, -1, -1, 0, 1, 0)
    assert p.n_devices == 1


def test_an_older_callers_shorter_parameter_block_is_accepted():
    """struct_size lets tsq_params grow: a caller compiled against 0.1 (no n_devices) still gets through the
    parameter checks (and then, here, fails on the missing device, not on the block)."""
    L = t.load_library()
    h = C.c_void_p()
    p = capi.Params()
    L.tsq_default_params(C.byref(p))
    p.struct_size = capi.Params.n_devices.offset
    p.n_devices = 99                     # beyond struct_size: must be ignored
    assert L.tsq_create(C.byref(h), C.byref(p)) in (0, -2)
    if h.value:
        L.tsq_destroy(h)
    p.struct_size = C.sizeof(capi.Params) + 8
    assert L.tsq_create(C.byref(h), C.byref(p)) == -1


def test_parameter_validation_happens_before_device_probe():
    L = t.load_library()
    h = C.c_void_p()
    p = capi.Params()
    L.tsq_default_params(C.byref(p))
    p.alphabet = 7
    assert L.tsq_create(C.byref(h), C.byref(p)) == -1
    L.tsq_default_params(C.byref(p))
    p.part_rank, p.part_world = 3, 2
    assert L.tsq_create(C.byref(h), C.byref(p)) == -1
    L.tsq_default_params(C.byref(p))
    m = np.zeros((23, 23), dtype=np.int8)
    m[0, 1] = 3                      # asymmetric
    p.matrix = m.ctypes.data_as(C.POINTER(C.c_int8))
    assert L.tsq_create(C.byref(h), C.byref(p)) == -8
    assert L.tsq_create(None, None) == -1


def test_no_cpu_fallback_without_a_device():
    if t.load_library().tsq_device_count() > 0:
        return                       # on the GPU box the gpu-marked tests cover creation
    try:
        t.Context()
    except t.TsqError as e:
        assert e.status == -2
    else:
        raise AssertionError("tsq_create must fail without an sm_100 device")


def test_null_context_calls_are_errors_not_crashes():
    L = t.load_library()
    for f in ("tsq_upload", "tsq_compute", "tsq_download", "tsq_synchronize", "tsq_finalize"):
        assert getattr(L, f)(None) == -1
    assert L.tsq_destroy(None) == 0
    assert L.tsq_last_error(None) == b"null context"
    assert L.tsq_run_fasta(None, None, None, capi.LOG_CB(0), None, None) == -1
    assert L.tsq_msa(None, None, None, None, None) == -1
    assert L.tsq_write_msa_fasta(None, None, None, None, b"/tmp/x.fa", 1) == -1
    assert L.tsq_guide_tree(None, None, None) == -1
    assert L.tsq_write_distmat(None, None, 0, None) == -1
    assert L.tsq_write_distmat(b"/nonexistent-dir/x.dist", None, 0, None) == -7


def test_distmat_writer_prints_what_printf_would(tmp_path):
    """tsq_write_distmat (host only): same digits as "%.6f", including halves, tiny, negative and > 1 values."""
    import numpy as np
    from tweakseq_b200 import capi
    from tweakseq_b200.fasta import read_distmat
    rng = np.random.default_rng(4)
    n = 37
    d = rng.random(n * (n - 1) // 2)
    d[:12] = [0.0, 1.0, 0.5, 0.0000005, 0.0000015, 0.9999995, 1.0000005, -0.25, 1e-9, 2.5, 0.1234565, 0.1234575]
    labels = [f"seq{k}" for k in range(n)]
    path = str(tmp_path / "m.dist")
    capi.write_distmat(path, labels, d)
    lines = open(path).read().split("\n")
    assert lines[0] == str(n) and lines[-1] == ""
    k = 0
    for i in range(n):
        parts = lines[1 + i].split(" ")
        assert parts[0] == labels[i] and len(parts) == n + 1
        for j in range(n):
            v = 0.0 if i == j else d[capi.pair_index(min(i, j), max(i, j), n)]
            assert parts[1 + j] == "%.6f" % v, (i, j, v)
    lab, rows = read_distmat(path)
    assert lab == labels and rows[3][3] == 0.0
    capi.write_distmat(str(tmp_path / "empty.dist"), [], np.zeros(0))
    assert open(tmp_path / "empty.dist").read() == "0\n"


# ---- tsq_encode: the (vectorised) encoder of tsq_set_sequences against the oracle's ---------------------------
@pytest.mark.parametrize("alphabet", [0, 1])
def test_encode_matches_the_oracle_on_every_byte_and_length(alphabet, monkeypatch):
    """encode_simd.cpp maps 32 bytes per step with two 16-entry shuffles; the oracle walks its 256-entry table.
    Every byte value, every length around the block size, runs of dropped bytes ('-', '.', whitespace) at block
    boundaries, and the self score S(x, x) summed on the way."""
    from tweakseq_b200 import capi
    from oracle import pyoracle as o
    rng = np.random.default_rng(4100 + alphabet)
    mat = o.matrix(alphabet)
    letters = np.frombuffer(b"ARNDCQEGHILKMFPSTWYVBZXJOUacgtun-. \t\n\r*1", dtype=np.uint8)
    cases = [bytes(range(256)), bytes(range(256)) * 3, b"", b"-", b"A", b"-" * 70, b"A" * 31 + b"-" + b"C" * 32 + b"." + b"G" * 33]
    for l in list(range(0, 100)) + [127, 128, 129, 1000, 4097]:
        cases.append(rng.choice(letters, l).tobytes())
        cases.append(rng.integers(0, 256, l, dtype=np.uint8).tobytes())
    for hook in (False, True):
        if hook:
            monkeypatch.setenv("TSQ_ENCODE_SCALAR", "1")
        for raw in cases:
            got, self_score = capi.encode(raw, alphabet)
            want = o.encode(raw, alphabet)
            assert got.tobytes() == want.tobytes(), raw[:80]
            assert self_score == o.self_score(want, mat)


def test_encode_rejects_bad_arguments():
    from tweakseq_b200 import capi
    with pytest.raises(capi.TsqError):
        capi.encode(b"ACGT", 7)
