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
