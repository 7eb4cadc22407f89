"""host/tsq-aligner: the backend behind the boundary tweakseq already has -- a child process started with the
argv of its tool wrappers (ClustalO.cpp:51, Muscle.cpp:52, MAFFT.cpp:52; version probes ClustalO.cpp:100-111,
Muscle.cpp:103).  CPU leg: argv handling, alphabet detection, loud failure without a B200.  GPU leg: the three
argv conventions end to end against the in-memory alignment."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "tsq-aligner")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tweakseq_b200", "csrc")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")], stdout=subprocess.DEVNULL)
    return EXE


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)


def write(path, labels, seqs):
    with open(path, "w") as f:
        for l, s in zip(labels, seqs):
            f.write(f">{l} a description\n{s}\n")


def test_version_probes_of_all_three_wrappers(exe):
    for flag in ("--version", "-version"):
        out = run(exe, flag)
        assert out.returncode == 0 and out.stdout.strip().startswith("tsq-b200")


def test_the_three_argv_conventions_are_understood(exe, tmp_path):
    prot, nuc = str(tmp_path / "p.fa"), str(tmp_path / "n.fa")
    write(prot, ["a", "b"], ["MKTAYIAKQR", "MKTAYIAKQK"])
    write(nuc, ["a", "b"], ["ACGTACGTNNACGU", "ACGTTTGA"])
    fout = str(tmp_path / "o.fa")
    # ClustalO.cpp:51
    out = run(exe, "--force", "-v", "--outfmt=fa", "--output-order=tree-order", "-i", prot, "-o", fout, "--dry-run")
    assert out.returncode == 0 and f"in={prot} out={fout} alphabet=protein" in out.stdout and "output=alignment" in out.stdout
    # Muscle.cpp:52
    out = run(exe, "-in", nuc, "-out", fout, "--dry-run")
    assert out.returncode == 0 and f"in={nuc} out={fout} alphabet=nucleotide" in out.stdout
    # MAFFT.cpp:52 (alignment on stdout, MAFFT.cpp:98)
    out = run(exe, "--auto", "--thread", "-1", prot, "--dry-run")
    assert out.returncode == 0 and f"in={prot} out=<stdout> alphabet=protein" in out.stdout
    # explicit type beats detection; extras
    out = run(exe, "-i", nuc, "-o", fout, "--seqtype=Protein", "--gap-open", "7", "--gap-extend=2", "--matrix-only", "--dry-run")
    assert "alphabet=protein gap_open=7 gap_extend=2" in out.stdout and "output=matrix" in out.stdout and "order=tree" in out.stdout
    out = run(exe, "-i", prot, "-o", fout, "--output-order=input-order", "--dry-run")
    assert out.returncode == 0 and "order=input" in out.stdout
    assert run(exe, "-i", prot, "-o", fout, "--output-order=random", "--dry-run").returncode == 2


def test_usage_errors_and_missing_files(exe, tmp_path):
    assert run(exe).returncode == 2
    assert run(exe, "--bogus").returncode == 2
    assert run(exe, "-i").returncode == 2
    assert run(exe, "-i", "x.fa", "-o", "y.fa", "--outfmt=clu").returncode == 2          # FASTA only
    out = run(exe, "-i", str(tmp_path / "missing.fa"), "-o", str(tmp_path / "o.fa"))
    assert out.returncode == 1 and "cannot open" in out.stderr


def test_without_a_b200_it_fails_loudly(exe, tmp_path):
    import tweakseq_b200 as t
    if t.load_library().tsq_device_count() > 0:
        pytest.skip("a B200 is present")
    fin = str(tmp_path / "p.fa")
    write(fin, ["a", "b"], ["MKTAYIAKQR", "MKTAYIAKQK"])
    out = run(exe, "-i", fin, "-o", str(tmp_path / "o.fa"))
    assert out.returncode == 1 and "no CPU path" in out.stderr and not os.path.exists(tmp_path / "o.fa.dnd")


@pytest.mark.gpu
def test_gpu_cli_end_to_end_all_conventions(exe, tmp_path):
    import tweakseq_b200 as t
    from tweakseq_b200.fasta import read_fasta
    rng = np.random.default_rng(90)
    root = rng.choice(list("ARNDCQEGHILKMFPSTWYV"), 90)
    seqs = ["".join(c if rng.random() > 0.2 else rng.choice(list("ARNDCQEGHILKMFPSTWYV")) for c in root if rng.random() > 0.05)
            for _ in range(12)]
    labels = [f"s{k}" for k in range(12)]
    fin = str(tmp_path / "in.fa")
    write(fin, labels, seqs)
    want, order = t.B200Gotoh().multiple_alignment(seqs)
    expect = [(labels[r], want[r]) for r in order]                      # tree order, as --output-order=tree-order

    fout = str(tmp_path / "clustalo.fa")
    out = run(exe, "--force", "-v", "--outfmt=fa", "--output-order=tree-order", "-i", fin, "-o", fout)
    assert out.returncode == 0 and "progressive alignment" in out.stdout, out.stdout + out.stderr
    lab, rows, _ = read_fasta(fout)
    assert list(zip(lab, rows)) == expect

    fout = str(tmp_path / "muscle.fa")
    assert run(exe, "-in", fin, "-out", fout).returncode == 0
    lab, rows, _ = read_fasta(fout)
    assert list(zip(lab, rows)) == expect

    out = run(exe, "--auto", "--thread", "-1", fin)                     # mafft: the alignment IS stdout
    assert out.returncode == 0
    so = str(tmp_path / "mafft.fa")
    open(so, "w").write(out.stdout)
    lab, rows, _ = read_fasta(so)
    assert list(zip(lab, rows)) == expect


@pytest.mark.gpu
def test_gpu_round_trip_through_the_references_own_fasta_code(exe, tmp_path):
    """The drop-in property itself: the input file is written by the reference's OWN FASTAFile::write (what
    Project::exportFASTA calls, Project.cpp:870-881) and the tool's output is parsed by the reference's OWN
    FASTAFile::read (what Project::readNewAlignment calls, Project.cpp:908-915) -- both compiled from the
    reference sources into oracle/_ref/libref_fasta.so.  Labels must come back one to one and the rows must be
    the in-memory alignment."""
    from oracle import pyoracle as o
    if not o.ref_fasta_available():
        pytest.skip("oracle/_ref/libref_fasta.so not built (needs /root/reference at build time)")
    import tweakseq_b200 as t
    rng = np.random.default_rng(91)
    root = rng.choice(list("ARNDCQEGHILKMFPSTWYV"), 170)
    seqs = ["".join(c if rng.random() > 0.25 else rng.choice(list("ARNDCQEGHILKMFPSTWYV")) for c in root if rng.random() > 0.06)
            for _ in range(15)]
    labels = [f"seq_{k}" for k in range(15)]
    comments = [f">{l} exported by tweakseq" for l in labels]
    fin, fout = str(tmp_path / "tweakseq.in.fa"), str(tmp_path / "tweakseq.out.fa")
    o.ref_fasta_write(fin, labels, seqs, comments)                       # 80-column lines, comment line verbatim
    out = run(exe, "--force", "-v", "--outfmt=fa", "--output-order=tree-order", "-i", fin, "-o", fout)
    assert out.returncode == 0, out.stdout + out.stderr
    got_labels, got_rows, got_comments = o.ref_fasta_read(fout)
    want, order = t.B200Gotoh().multiple_alignment(seqs)
    assert got_labels == [labels[r] for r in order]                      # every label, once, in tree order
    assert got_rows == [want[r] for r in order]
    assert got_comments == [comments[r] for r in order]                  # header lines come back verbatim
    assert len({len(r) for r in got_rows}) == 1 and all(r.replace("-", "") == seqs[labels.index(l)] for l, r in zip(got_labels, got_rows))


@pytest.mark.gpu
@pytest.mark.parametrize("nrows,ncols", [(2, 33), (17, 64), (100, 301)])
def test_gpu_consensus_equals_the_references_own_code(nrows, ncols):
    """tsq_consensus against Consensus::calculate of the reference itself (oracle/_ref/libref_consensus.so =
    tweakseq/Core/Annotations/Consensus.cpp compiled where it lies), not only against the restatement."""
    from oracle import pyoracle as o
    if not o.ref_consensus_available():
        pytest.skip("oracle/_ref/libref_consensus.so not built (needs /root/reference at build time)")
    import tweakseq_b200 as t
    from test_consensus import _random_alignment
    rng = np.random.default_rng(nrows * 7 + ncols)
    rows = _random_alignment(rng, nrows, ncols)
    with t.Context() as ctx:
        assert ctx.consensus(rows) == o.ref_consensus(rows)
        assert ctx.consensus(rows, plurality=nrows * 0.8) == o.ref_consensus(rows, nrows * 0.8)


# ---- driven by the reference's OWN wrappers (oracle/_ref/libref_tools.so = tweakseq/Core/ClustalO.cpp, Muscle.cpp,
# MAFFT.cpp, AlignmentTool.cpp compiled where they lie; QProcess stand-in that really forks and execs) -------------

def _ref_tools():
    from oracle import pyoracle as o
    if not o.ref_tools_available():
        pytest.skip("oracle/_ref/libref_tools.so not built (needs /root/reference at build time)")
    return o


def test_the_references_own_wrappers_probe_and_drive_the_binary(exe, tmp_path):
    o = _ref_tools()
    # getVersion() of each wrapper, run on tsq-aligner: what the editor would show as the tool's version
    assert o.ref_tool_version("clustalo", exe).startswith("tsq-b200")          # ClustalO.cpp:100-111: stdout, trimmed
    assert o.ref_tool_version("muscle", exe) == "0.1"                          # Muscle.cpp:100-112: second word of stdout
    assert o.ref_tool_version("mafft", exe).startswith("tsq-b200")             # MAFFT.cpp:103-116: stderr
    # makeCommand() of each wrapper: the argv is accepted and means what the wrapper means by it
    fin, fout = str(tmp_path / "in.fa"), str(tmp_path / "out.fa")
    write(fin, ["a", "b"], ["MKTAYIAKQR", "MKTAYIAKQK"])
    for tool in ("clustalo", "muscle", "mafft"):
        name, _, args, uses_stdout = o.ref_tool_command(tool, fin, fout, exe)
        out = run(exe, *args, "--dry-run")
        assert out.returncode == 0, (tool, args, out.stderr)
        assert f"in={fin} " in out.stdout and "output=alignment" in out.stdout
        assert ("out=<stdout>" in out.stdout) == uses_stdout                   # MAFFT.cpp:98: the alignment is stdout
        if not uses_stdout:
            assert f"out={fout} " in out.stdout


@pytest.mark.gpu
def test_gpu_editor_round_trip_with_reference_code_on_both_sides(exe, tmp_path):
    """startAlignment() to readNewAlignment() with the reference's own code wherever it has any: its writer makes the
    input file, its wrapper makes the argv (stdout redirected to the output file when the wrapper says so,
    SeqEditMainWin.cpp:1655-1657), this repository's binary aligns on the B200, its reader parses the result."""
    o = _ref_tools()
    if not o.ref_fasta_available():
        pytest.skip("oracle/_ref/libref_fasta.so not built")
    import tweakseq_b200 as t
    rng = np.random.default_rng(92)
    root = rng.choice(list("ARNDCQEGHILKMFPSTWYV"), 120)
    seqs = ["".join(c if rng.random() > 0.2 else rng.choice(list("ARNDCQEGHILKMFPSTWYV")) for c in root if rng.random() > 0.05)
            for _ in range(10)]
    labels = [f"p{k}" for k in range(10)]
    comments = [f">{l} from the editor" for l in labels]
    want, order = t.B200Gotoh().multiple_alignment(seqs)
    for tool in ("clustalo", "muscle", "mafft"):
        fin, fout = str(tmp_path / f"{tool}.in.fa"), str(tmp_path / f"{tool}.out.fa")
        o.ref_fasta_write(fin, labels, seqs, comments)
        _, exec_, args, uses_stdout = o.ref_tool_command(tool, fin, fout, exe)
        out = subprocess.run([exec_, *args], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, (tool, out.stdout, out.stderr)
        if uses_stdout:
            open(fout, "w").write(out.stdout)
        got_labels, got_rows, _ = o.ref_fasta_read(fout)
        assert got_labels == [labels[r] for r in order] and got_rows == [want[r] for r in order], tool


@pytest.mark.gpu
def test_gpu_qt_worker_runs_the_alignment_in_process(tmp_path):
    """host/qt/B200GotohTool.cpp itself (compiled against the reference's AlignmentTool.h over functional Qt
    stand-ins): B200GotohWorker::start() -> tool->run() -> tsq_run_fasta on the B200 -> finished(0, NormalExit), and
    the output file read by the reference's own FASTA reader is the in-memory alignment."""
    from oracle import pyoracle as o
    if not (o.ref_qt_adapter_available() and o.ref_fasta_available()):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import tweakseq_b200 as t
    rng = np.random.default_rng(93)
    root = rng.choice(list("ARNDCQEGHILKMFPSTWYV"), 100)
    seqs = ["".join(c if rng.random() > 0.2 else rng.choice(list("ARNDCQEGHILKMFPSTWYV")) for c in root if rng.random() > 0.05)
            for _ in range(9)]
    labels = [f"q{k}" for k in range(9)]
    fin, fout = str(tmp_path / "in.fa"), str(tmp_path / "out.fa")
    o.ref_fasta_write(fin, labels, seqs, [f">{l}" for l in labels])
    code, status, log = o.qt_worker_run(fin, fout, align_in_process=True)
    assert (code, status) == (0, 0), log
    assert any("progressive alignment" in l for l in log)
    want, order = t.B200Gotoh().multiple_alignment(seqs)
    got_labels, got_rows, _ = o.ref_fasta_read(fout)
    assert got_labels == [labels[r] for r in order] and got_rows == [want[r] for r in order]


@pytest.mark.gpu
def test_gpu_qt_worker_in_memory_route_keeps_the_labels(tmp_path):
    """The hazard of the file route: Project::exportFASTA writes `comment` as the header
    (Project.cpp:876-880), so a renamed sequence or a PDB import (comment without '>') comes back under a label
    readNewAlignment cannot match.  The in-memory worker takes (label, filter(true)) from the model and writes
    ">label": the reference's own reader must find exactly the labels that went in, nucleotides detected."""
    from oracle import pyoracle as o
    if not (o.ref_qt_adapter_available() and o.ref_fasta_available()):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import tweakseq_b200 as t
    rng = np.random.default_rng(94)
    seqs = ["".join(rng.choice(list("ACGT"), int(l))) for l in rng.integers(60, 110, 7)]
    labels = ["renamed_1", "1abc_A", "chr1:100-200", "s3", "s4", "s5", "s6"]      # none of them is its comment's first word
    fout = str(tmp_path / "out.fa")
    code, status, log = o.qt_worker_run_in_memory(labels, seqs, fout)
    assert (code, status) == (0, 0), log
    assert any("7 sequences from the project (nucleotide)" in l for l in log), log
    tool = t.B200Gotoh()
    tool.alphabet = t.NUCLEOTIDE
    want, order = tool.multiple_alignment(seqs)
    got_labels, got_rows, _ = o.ref_fasta_read(fout)
    assert got_labels == [labels[r] for r in order] and got_rows == [want[r] for r in order]
