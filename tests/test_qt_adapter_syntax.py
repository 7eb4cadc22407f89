"""the Qt adapter (host/qt/B200GotohTool.{h,cpp}) cannot be built here (no Qt5), but it can
be type-checked: g++ -fsyntax-only against the REAL reference headers it derives from
(tweakseq/Core/AlignmentTool.h:36-71, Core/XMLHelper.h) plus declaration-only Qt stubs (tests/qt_stub/).
Skipped where /root/reference is absent (the GPU box)."""
import os
import subprocess

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
    for decl in ("virtual void makeCommand(QString &, QString &, QString &, QStringList &);",
                 "virtual void writeSettings(QDomDocument &, QDomElement &);",
                 "virtual void readSettings(QDomDocument &);",
                 "virtual bool inProcess(){return true;}",
                 "virtual int run(const QString &fin, const QString &fout, QObject *logReceiver, volatile int *cancel);"):
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
