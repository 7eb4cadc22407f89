"""The C++ adapter (host/): builds against the C ABI, mirrors AlignmentTool, refuses to run without
a B200 (CPU leg) and runs end to end with one (gpu leg)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tweakseq_b200", "csrc")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    return os.path.join(HOST, "selftest")


def test_cpp_adapter_builds_and_mirrors_the_interface():
    exe = _build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host selftest ok" in out.stdout


def test_adapter_headers_cite_the_reference_interface():
    src = open(os.path.join(HOST, "AlignmentTool.h")).read()
    for member in ("name()", "version()", "executable()", "setExecutable", "setPreferred", "preferred()",
                   "usesStdOut()", "makeCommand", "writeSettings", "readSettings", "inProcess", "run("):
        assert member in src
    assert "AlignmentTool.h:36-71" in src


@pytest.mark.gpu
def test_cpp_adapter_runs_fasta_end_to_end(tmp_path):
    from tweakseq_b200 import synth
    from tweakseq_b200.fasta import read_distmat, write_fasta
    from oracle import pyoracle as o
    exe = _build()
    _, seqs = synth.config(1, 0.2)
    labels = [f"s{k}" for k in range(len(seqs))]
    fin, fout, faln = str(tmp_path / "in.fa"), str(tmp_path / "out.mat"), str(tmp_path / "out.aln.fa")
    write_fasta(fin, labels, seqs)
    out = subprocess.run([exe, fin, fout, faln], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "B200 path" in out.stdout, out.stdout + out.stderr
    # alignInProcess: the third file is the multiple alignment, every input row back under its label
    from tweakseq_b200.fasta import read_fasta
    alab, arows, _ = read_fasta(faln)
    assert sorted(alab) == sorted(labels) and len({len(r) for r in arows}) == 1
    for l, r in zip(alab, arows):
        assert r.replace("-", "") == seqs[labels.index(l)]
    lab, rows = read_distmat(fout)
    assert lab == labels
    enc = [o.encode(s) for s in seqs]
    ref, _ = o.all_pairs(enc, o.matrix(0), 11, 1, nthreads=os.cpu_count() or 1)
    selfs = np.array([o.self_score(e, o.matrix(0)) for e in enc], dtype=np.int32)
    d = o.distances(ref, selfs)
    n = len(seqs)
    k = 0
    for i in range(n):
        for j in range(i + 1, n):
            assert abs(rows[i][j] - d[k]) < 5e-7
            k += 1
