"""the Qt adapter (host/qt/B200GotohTool.{h,cpp}) cannot be built here (no Qt5), but it can
be type-checked: g++ -fsyntax-only against the REAL reference headers it derives from
(tweakseq/Core/AlignmentTool.h:36-71, Core/XMLHelper.h) plus declaration-only Qt stubs (tests/qt_stub/).
Skipped where /root/reference is absent (the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CORE = "/root/reference/tweakseq/Core"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_CORE, "AlignmentTool.h")), reason="reference sources not present")
def test_qt_adapter_type_checks_against_the_reference_headers():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter",
           "-I", os.path.join(ROOT, "tests", "qt_stub"), "-I", REF_CORE, "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "host", "qt", "B200GotohTool.cpp")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr


def test_qt_adapter_overrides_exactly_the_reference_virtuals():
    """Every virtual of the reference class is re-declared with the same parameter list, plus the two additions."""
    src = open(os.path.join(ROOT, "host", "qt", "B200GotohTool.h")).read()
    for decl in ("virtual void makeCommand(QString &, QString &, QString &, QStringList &) override;",
                 "virtual void writeSettings(QDomDocument &, QDomElement &) override;",
                 "virtual void readSettings(QDomDocument &) override;",
                 "virtual bool inProcess() TSQ_OVERRIDE {return true;}",
                 "virtual int run(const QString &fin, const QString &fout, QObject *logReceiver, volatile int *cancel) TSQ_OVERRIDE;"):
        assert decl in src, decl


# ---- and EXECUTED: oracle/_ref/libref_qt_adapter.so is host/qt/B200GotohTool.cpp itself, compiled against the
# reference's real AlignmentTool.h next to the reference's own ClustalO.cpp / AlignmentTool.cpp, over the functional
# Qt stand-ins of oracle/ref_shim (DOM, QThread that runs in place, recorded queued invocations) ------------------

def _qt():
    from oracle import pyoracle as o
    if not o.ref_qt_adapter_available():
        pytest.skip("oracle/_ref/libref_qt_adapter.so not built (needs /root/reference at build time)")
    return o


def test_qt_adapter_settings_live_in_one_document_with_the_references_clustalo():
    """Project::writeSettings / readAlignmentToolSettings (Project.cpp:853-862, 1161-1182): every tool writes its own
    <alignment_tool> element and skips the others' by name (ClustalO.cpp:71-74).  Neither side picks up the other's."""
    r = _qt().qt_settings_round_trip()
    assert r["elements"] == "2"
    assert r["clustalo.path"] == "/opt/somewhere/clustalo" and r["clustalo.preferred"] == "no"      # not b200's "yes"
    assert r["b200.name"] == "b200gotoh" and r["b200.path"] == "libtsqb200.so" and r["b200.preferred"] == "yes"
    assert (r["b200.gap_open"], r["b200.gap_extend"], r["b200.device"], r["b200.align_in_process"]) == ("9", "2", "1", "no")
    assert r["b200.version"].startswith("tsq-b200") and r["b200.in_process"] == "yes" and r["b200.argc"] == "8"


def test_qt_worker_reports_like_qprocess_finished_and_refuses_without_a_b200(tmp_path):
    import tweakseq_b200 as t
    o = _qt()
    if t.load_library().tsq_device_count() > 0:
        pytest.skip("a B200 is present: the gpu leg covers the run")
    fin = str(tmp_path / "in.fa")
    open(fin, "w").write(">a x\nMKTAYIAK\n>b\nMKTAYIAR\n")
    code, status, log = o.qt_worker_run(fin, str(tmp_path / "out.fa"))
    assert (code, status) == (-2, 0)                       # TSQ_ERR_NO_DEVICE through finished(int, int): no fallback
    assert any("read 2 sequences" in l for l in log)       # log lines arrive as queued MessageWin::addMessage calls


# ---- the registration edits as an artifact: host/qt/tweakseq_registration.patch -----------------
PATCH = os.path.join(ROOT, "host", "qt", "tweakseq_registration.patch")
REF_ROOT = "/root/reference"
PATCHED_FILES = ["tweakseq/Core/AlignmentTool.h", "tweakseq/Core/Project.h", "tweakseq/Core/Project.cpp",
                 "tweakseq/Core/Application.h", "tweakseq/Core/Application.cpp", "tweakseq/UI/SeqEditMainWin.h",
                 "tweakseq/UI/SeqEditMainWin.cpp", "tweakseq/tweakseq.pro"]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_CORE, "AlignmentTool.h")), reason="reference sources not present")
def test_registration_patch_applies_to_the_reference_and_the_adapter_overrides_the_patched_virtuals(tmp_path):
    import shutil
    # the committed patch is what the generator produces from the reference as it is
    regen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_qt_patch.py"), REF_ROOT, str(tmp_path / "regen.patch")],
                           capture_output=True, text=True, timeout=120)
    assert regen.returncode == 0, regen.stderr
    assert open(tmp_path / "regen.patch", encoding="latin-1").read() == open(PATCH, encoding="latin-1").read()
    # dry run against the reference tree itself (read-only: nothing is written) ...
    dry = subprocess.run(["patch", "-p1", "--dry-run", "-d", REF_ROOT, "-i", PATCH], capture_output=True, text=True, timeout=120)
    assert dry.returncode == 0, dry.stdout + dry.stderr
    assert dry.stdout.count("checking file") == len(PATCHED_FILES) and "FAILED" not in dry.stdout and "fuzz" not in dry.stdout
    # ... and for real on a copy of the eight files
    work = tmp_path / "tree"
    for f in PATCHED_FILES:
        os.makedirs(work / os.path.dirname(f), exist_ok=True)
        shutil.copy(os.path.join(REF_ROOT, f), work / f)
    out = subprocess.run(["patch", "-p1", "-d", str(work), "-i", PATCH], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and not list(work.rglob("*.rej")), out.stdout + out.stderr
    txt = {f: open(work / f, encoding="latin-1").read() for f in PATCHED_FILES}
    assert "virtual bool inProcess(){return false;}" in txt["tweakseq/Core/AlignmentTool.h"]
    assert txt["tweakseq/Core/Project.cpp"].count("b200GotohTool_") >= 9 and "Core/B200GotohTool.cpp" in txt["tweakseq/tweakseq.pro"]
    cpp = txt["tweakseq/UI/SeqEditMainWin.cpp"]
    assert "SLOT(alignmentFinishedInProcess(int,int))" in cpp and "SLOT(alignmentMessage(QString))" in cpp
    assert "->filter(true)" in cpp and "alignmentWorker_->requestCancel()" in cpp
    # the slots the worker's signals are connected to exist with exactly those signatures
    assert "void alignmentFinishedInProcess(int,int);" in txt["tweakseq/UI/SeqEditMainWin.h"]
    assert "void alignmentMessage(const QString &);" in txt["tweakseq/UI/SeqEditMainWin.h"]
    # against the PATCHED AlignmentTool.h the adapter's inProcess()/run() are checked overrides (TSQ_OVERRIDE = override)
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror=suggest-override",
           "-I", os.path.join(ROOT, "tests", "qt_stub"), "-I", str(work / "tweakseq" / "Core"), "-I", REF_CORE,
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "host", "qt", "B200GotohTool.cpp")]
    chk = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert chk.returncode == 0, chk.stderr
