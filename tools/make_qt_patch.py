#!/usr/bin/env python
"""Generates host/qt/tweakseq_registration.patch: the mechanical edits (INTEGRATION.md
section 4) that register the in-process B200 tool in groundstate/tweakseq, as a unified diff against the reference
tree (apply with `patch -p1` from the repository root of tweakseq).

    python tools/make_qt_patch.py [/root/reference [out.patch]]

Every edit is an exact-string replacement anchored on reference text that must occur exactly once; the script fails
loudly when the reference moved.  Nothing of the reference is copied into this repository: the patch holds only
the changed hunks with their context lines.  host/qt/B200GotohTool.{h,cpp} are the two new files the patched
tree compiles (copy them to tweakseq/Core/)."""
import difflib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "host", "qt", "tweakseq_registration.patch")

EDITS = {}


def edit(path, old, new, count=1):
    EDITS.setdefault(path, []).append((old, new, count))


# ---- Core/AlignmentTool.h: the two virtuals an in-process tool needs -------------------------------
edit("tweakseq/Core/AlignmentTool.h",
     "class QDomDocument;\nclass QDomElement;\n",
     "class QDomDocument;\nclass QDomElement;\nclass QObject;\n\n"
     "#define TWEAKSEQ_ALIGNMENTTOOL_INPROCESS 1 // AlignmentTool has inProcess()/run(): tools may run inside the editor\n")
edit("tweakseq/Core/AlignmentTool.h",
     "\t\tvirtual void readSettings(QDomDocument &);\n\t\n\tprotected:",
     "\t\tvirtual void readSettings(QDomDocument &);\n"
     "\t\t\n"
     "\t\t// an in-process tool computes in run() (on a worker thread) instead of handing an argv to QProcess;\n"
     "\t\t// log lines go to logReceiver's message(QString) signal, *cancel != 0 asks it to stop; returns the exit code\n"
     "\t\tvirtual bool inProcess(){return false;}\n"
     "\t\tvirtual int  run(const QString &,const QString &,QObject *,volatile int *){return -1;}\n"
     "\t\n\tprotected:")

# ---- Core/Project.{h,cpp}: own the tool, select it, persist it ---------------------------------------------------
edit("tweakseq/Core/Project.h",
     "AlignmentTool *alignmentTool_,*mafftTool_,*clustalOTool_,*muscleTool_;",
     "AlignmentTool *alignmentTool_,*mafftTool_,*clustalOTool_,*muscleTool_,*b200GotohTool_;")
edit("tweakseq/Core/Project.cpp",
     '#include "ClustalO.h"\n',
     '#include "B200GotohTool.h"\n#include "ClustalO.h"\n')
edit("tweakseq/Core/Project.cpp",
     "\tif (mafftTool_) delete mafftTool_;\n}",
     "\tif (mafftTool_) delete mafftTool_;\n\tif (b200GotohTool_) delete b200GotohTool_;\n}")
edit("tweakseq/Core/Project.cpp",
     "\telse if (atool == \"MAFFT\" && mafftTool_)\n\t\talignmentTool_=mafftTool_;\n}",
     "\telse if (atool == \"MAFFT\" && mafftTool_)\n\t\talignmentTool_=mafftTool_;\n"
     "\telse if (atool == \"b200gotoh\" && b200GotohTool_)\n\t\talignmentTool_=b200GotohTool_;\n}")
edit("tweakseq/Core/Project.cpp",
     "\tif (mafftTool_)\n\t\tmafftTool_->writeSettings(doc,root);;\n}",
     "\tif (mafftTool_)\n\t\tmafftTool_->writeSettings(doc,root);;\n"
     "\tif (b200GotohTool_)\n\t\tb200GotohTool_->writeSettings(doc,root);\n}")
edit("tweakseq/Core/Project.cpp",
     "\tclustalOTool_ = NULL;\n\tif (app->alignmentToolAvailable(\"clustalo\"))\n\t\tclustalOTool_ = new ClustalO();\n",
     "\tclustalOTool_ = NULL;\n\tif (app->alignmentToolAvailable(\"clustalo\"))\n\t\tclustalOTool_ = new ClustalO();\n"
     "\t\n"
     "\tb200GotohTool_ = NULL; // in-process: available when the library finds a B200 (there is no CPU fallback)\n"
     "\tif (app->alignmentToolAvailable(\"b200gotoh\"))\n\t\tb200GotohTool_ = new B200GotohTool();\n")
edit("tweakseq/Core/Project.cpp",
     "\tif (mafftTool_){\n\t\tmafftTool_->readSettings(doc);\n\t\tif (mafftTool_->preferred())\n\t\t\talignmentTool_=mafftTool_;\n\t}\n}",
     "\tif (mafftTool_){\n\t\tmafftTool_->readSettings(doc);\n\t\tif (mafftTool_->preferred())\n\t\t\talignmentTool_=mafftTool_;\n\t}\n"
     "\t\n"
     "\tif (b200GotohTool_){\n\t\tb200GotohTool_->readSettings(doc);\n\t\tif (b200GotohTool_->preferred() || NULL == alignmentTool_)\n"
     "\t\t\talignmentTool_=b200GotohTool_;\n\t}\n}")

# ---- Core/Application.{h,cpp}: availability + default settings -----------------------------------------------------
edit("tweakseq/Core/Application.h",
     "bool clustaloConfigured_,muscleConfigured_,mafftConfigured_;",
     "bool clustaloConfigured_,muscleConfigured_,mafftConfigured_,b200gotohConfigured_;")
edit("tweakseq/Core/Application.cpp",
     '#include "ClustalO.h"\n',
     '#include "B200GotohTool.h"\n#include "ClustalO.h"\n#include "tsq_b200.h"\n')
edit("tweakseq/Core/Application.cpp",
     "\tif (mafftConfigured_){\n\t\tMAFFT atool;\n\t\tatool.setPreferred(preferredTool == \"MAFFT\");\n\t\tatool.setExecutable(mafft);\n"
     "\t\tatool.writeSettings(saveDoc,root);\n\t}\n",
     "\tif (mafftConfigured_){\n\t\tMAFFT atool;\n\t\tatool.setPreferred(preferredTool == \"MAFFT\");\n\t\tatool.setExecutable(mafft);\n"
     "\t\tatool.writeSettings(saveDoc,root);\n\t}\n"
     "\t\n"
     "\tif (tsq_device_count() > 0){ // nothing to locate: the tool is libtsqb200.so, linked in\n"
     "\t\tB200GotohTool atool;\n\t\tatool.setPreferred(preferredTool == \"b200gotoh\");\n\t\tatool.writeSettings(saveDoc,root);\n\t}\n")
edit("tweakseq/Core/Application.cpp",
     "\telse if (toolName == \"MAFFT\")\n\t\treturn mafftConfigured_;\n",
     "\telse if (toolName == \"MAFFT\")\n\t\treturn mafftConfigured_;\n"
     "\telse if (toolName == \"b200gotoh\")\n\t\treturn b200gotohConfigured_ || tsq_device_count() > 0;\n")
edit("tweakseq/Core/Application.cpp",
     "clustaloConfigured_=muscleConfigured_=mafftConfigured_=false;",
     "clustaloConfigured_=muscleConfigured_=mafftConfigured_=b200gotohConfigured_=false;")
edit("tweakseq/Core/Application.cpp",
     "\t\t\t\t\telse if (elem.text()==\"MAFFT\")\n\t\t\t\t\t\tmafftConfigured_=true;\n",
     "\t\t\t\t\telse if (elem.text()==\"MAFFT\")\n\t\t\t\t\t\tmafftConfigured_=true;\n"
     "\t\t\t\t\telse if (elem.text()==\"b200gotoh\")\n\t\t\t\t\t\tb200gotohConfigured_=(tsq_device_count() > 0);\n")

# ---- UI/SeqEditMainWin.{h,cpp}: menu item, launch branch, stop, the two slots the worker's signals reach ----------------
edit("tweakseq/UI/SeqEditMainWin.h",
     "\tvoid alignmentFinished(int,QProcess::ExitStatus);\n",
     "\tvoid alignmentFinished(int,QProcess::ExitStatus);\n"
     "\tvoid alignmentFinishedInProcess(int,int); // B200GotohWorker::finished: no QProcess behind it\n"
     "\tvoid alignmentMessage(const QString &);   // B200GotohWorker::message -> mw->addMessage\n")
edit("tweakseq/UI/SeqEditMainWin.h",
     "\tvoid settingsAlignmentToolMAFFT();\n",
     "\tvoid settingsAlignmentToolMAFFT();\n\tvoid settingsAlignmentToolB200Gotoh();\n")
edit("tweakseq/UI/SeqEditMainWin.h",
     "*settingsAlignmentToolMAFFTAction,*settingsAlignmentToolMUSCLEAction,*settingsAlignmentToolClustalOAction;",
     "*settingsAlignmentToolMAFFTAction,*settingsAlignmentToolMUSCLEAction,*settingsAlignmentToolClustalOAction,"
     "*settingsAlignmentToolB200GotohAction;")
edit("tweakseq/UI/SeqEditMainWin.h",
     "\tQProcess *alignmentProc_;\n",
     "\tQProcess *alignmentProc_;\n\tclass B200GotohWorker *alignmentWorker_; // the in-process counterpart of alignmentProc_\n")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     '#include "ClustalO.h"\n',
     '#include "B200GotohTool.h"\n#include "ClustalO.h"\n')
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "\talignmentProc_=NULL;\n\talignmentFileOut_=alignmentFileIn_=NULL;\n",
     "\talignmentProc_=NULL;\n\talignmentWorker_=NULL;\n\talignmentFileOut_=alignmentFileIn_=NULL;\n")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "\tif (NULL != alignmentProc_){\n\t\talignStopAction->setEnabled(alignmentProc_->state() == QProcess::Running);\n\t}\n",
     "\tif (NULL != alignmentProc_){\n\t\talignStopAction->setEnabled(alignmentProc_->state() == QProcess::Running);\n\t}\n"
     "\tif (NULL != alignmentWorker_ && alignmentWorker_->isRunning())\n\t\talignStopAction->setEnabled(true);\n")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "\tif (NULL != alignmentProc_){\n\t\talignmentProc_->kill();\n",
     "\tif (NULL != alignmentWorker_ && alignmentWorker_->isRunning()){\n"
     "\t\talignmentWorker_->requestCancel(); // the kernels poll the flag; finished(9,0) follows: \"user interrupted\"\n"
     "\t\treturn;\n\t}\n"
     "\tif (NULL != alignmentProc_){\n\t\talignmentProc_->kill();\n")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "void SeqEditMainWin::setupColourMapMenu()\n{",
     "void SeqEditMainWin::alignmentMessage(const QString &msg)\n{\n\tmw->addMessage(msg);\n}\n\n"
     "void SeqEditMainWin::alignmentFinishedInProcess(int exitCode,int exitStatus)\n{\n"
     "\tqDebug() << trace.header(__PRETTY_FUNCTION__) << \" exitCode=\" << exitCode << \" exitStatus=\" << exitStatus;\n"
     "\tQFile f(alignmentFileOut_->fileName());\n"
     "\tif (exitStatus == 0 && exitCode == 0 && f.exists()){\n"
     "\t\tstatusBar()->showMessage(\"Alignment finished\");\n"
     "\t\treadNewAlignment(alignAll);\n"
     "\t}\n"
     "\telse{\n"
     "\t\tQString msg = \"Alignment not completed\";\n"
     "\t\tif (exitCode == 9)\n\t\t\tmsg += \" (user interrupted)\";\n"
     "\t\tstatusBar()->showMessage(msg);\n"
     "\t}\n"
     "\talignAllAction->setEnabled(true);\n\talignStopAction->setEnabled(false);\n}\n\n"
     "void SeqEditMainWin::setupColourMapMenu()\n{")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "void SeqEditMainWin::settingsAlignmentToolProperties()\n{",
     "void SeqEditMainWin::settingsAlignmentToolB200Gotoh()\n{\n"
     "\tif (project_->alignmentTool()->name() != \"b200gotoh\"){\n"
     "\t\tproject_->setAlignmentTool(\"b200gotoh\");\n"
     "\t\tsettingsAlignmentToolPropertiesAction->setText(\"b200gotoh\");\n\t}\n}\n\n"
     "void SeqEditMainWin::settingsAlignmentToolProperties()\n{")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "\tQActionGroup *ag = new QActionGroup(this);\n\tag->setExclusive(true);\n\tag->addAction(settingsAlignmentToolClustalOAction);\n",
     "\tsettingsAlignmentToolB200GotohAction = new QAction( tr(\"B200 Gotoh (in process)\"), this);\n"
     "\tsettingsAlignmentToolB200GotohAction->setStatusTip(tr(\"Select the in-process B200 aligner\"));\n"
     "\taddAction(settingsAlignmentToolB200GotohAction);\n"
     "\tconnect(settingsAlignmentToolB200GotohAction, SIGNAL(triggered()), this, SLOT(settingsAlignmentToolB200Gotoh()));\n"
     "\tsettingsAlignmentToolB200GotohAction->setCheckable(true);\n"
     "\tsettingsAlignmentToolB200GotohAction->setChecked(project_->alignmentTool()->name()==\"b200gotoh\");\n"
     "\tsettingsAlignmentToolB200GotohAction->setEnabled(app->alignmentToolAvailable(\"b200gotoh\"));\n"
     "\t\n"
     "\tQActionGroup *ag = new QActionGroup(this);\n\tag->setExclusive(true);\n\tag->addAction(settingsAlignmentToolClustalOAction);\n"
     "\tag->addAction(settingsAlignmentToolB200GotohAction);\n")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "\talignmentToolMenu->addAction(settingsAlignmentToolMAFFTAction);\n",
     "\talignmentToolMenu->addAction(settingsAlignmentToolMAFFTAction);\n"
     "\talignmentToolMenu->addAction(settingsAlignmentToolB200GotohAction);\n")
edit("tweakseq/UI/SeqEditMainWin.cpp",
     "\tQString fin  = alignmentFileIn_->fileName();\n\tQString fout = alignmentFileOut_->fileName();\n",
     "\tQString fin  = alignmentFileIn_->fileName();\n\tQString fout = alignmentFileOut_->fileName();\n"
     "\t\n"
     "\tif (project_->alignmentTool()->inProcess()){\n"
     "\t\t// no child process and no input file: (label, filter(true)) straight from the model, so that the rows come\n"
     "\t\t// back under the labels readNewAlignment matches by -- exportFASTA writes `comment` as the header instead\n"
     "\t\tQStringList labels,residues;\n"
     "\t\tif (alignAll){\n"
     "\t\t\tfor (int s=0;s<project_->sequences.size();s++){\n"
     "\t\t\t\tlabels.append(project_->sequences.sequences().at(s)->label);\n"
     "\t\t\t\tresidues.append(project_->sequences.sequences().at(s)->filter(true));\n"
     "\t\t\t}\n"
     "\t\t}\n"
     "\t\telse{\n"
     "\t\t\tfor (int s=0;s<project_->sequenceSelection->size();s++){\n"
     "\t\t\t\tlabels.append(project_->sequenceSelection->itemAt(s)->label);\n"
     "\t\t\t\tresidues.append(project_->sequenceSelection->itemAt(s)->filter(true));\n"
     "\t\t\t}\n"
     "\t\t}\n"
     "\t\tif (NULL != alignmentWorker_){\n\t\t\talignmentWorker_->wait();\n\t\t\tdelete alignmentWorker_;\n\t\t}\n"
     "\t\talignmentWorker_ = new B200GotohWorker(static_cast<B200GotohTool *>(project_->alignmentTool()),labels,residues,fout,this);\n"
     "\t\tconnect(alignmentWorker_,SIGNAL(message(QString)),this,SLOT(alignmentMessage(QString)));\n"
     "\t\tconnect(alignmentWorker_,SIGNAL(finished(int,int)),this,SLOT(alignmentFinishedInProcess(int,int)));\n"
     "\t\talignmentStarted();\n"
     "\t\talignmentWorker_->start();\n"
     "\t\treturn;\n"
     "\t}\n"
     "\t\n")

# ---- tweakseq.pro ------------------------------------------------------------------------------------------------------
edit("tweakseq/tweakseq.pro",
     "\t\t\t\t\t\t\t\t include/ClustalO.h \\\n",
     "\t\t\t\t\t\t\t\t include/B200GotohTool.h \\\n\t\t\t\t\t\t\t\t include/ClustalO.h \\\n")
edit("tweakseq/tweakseq.pro",
     "\t\t\t\t\t\t\t\t\tCore/ClustalO.cpp \\\n",
     "\t\t\t\t\t\t\t\t\tCore/B200GotohTool.cpp \\\n\t\t\t\t\t\t\t\t\tCore/ClustalO.cpp \\\n")
edit("tweakseq/tweakseq.pro",
     "QT           += core gui xml widgets printsupport\n",
     "QT           += core gui xml widgets printsupport\n\n"
     "# the in-process B200 aligner: include/tsq_b200.h and libtsqb200.so of the tsq-b200 repository\n"
     "TSQ_B200_DIR = $$(TSQ_B200_DIR)\n"
     "INCLUDEPATH += $$TSQ_B200_DIR/include\n"
     "LIBS        += -L$$TSQ_B200_DIR/tweakseq_b200 -ltsqb200 -Wl,-rpath,$$TSQ_B200_DIR/tweakseq_b200\n")


def main():
    chunks = []
    for path, edits in EDITS.items():
        src = open(os.path.join(REF, path), encoding="latin-1").read()
        dst = src
        for old, new, count in edits:
            n = dst.count(old)
            if n != count:
                raise SystemExit(f"{path}: anchor occurs {n} times, expected {count}:\n{old!r}")
            dst = dst.replace(old, new)
        diff = difflib.unified_diff(src.splitlines(keepends=True), dst.splitlines(keepends=True),
                                    "a/" + path, "b/" + path, n=3)
        chunks.append("".join(diff))
    header = ("Registers the in-process B200 aligner (host/qt/B200GotohTool.{h,cpp} of tsq-b200, copied to tweakseq/Core/\n"
              "with a link or copy of the header in tweakseq/include/) in groundstate/tweakseq.\n"
              "Generated by tools/make_qt_patch.py; apply from the repository root with  patch -p1 < tweakseq_registration.patch\n"
              "Build with  TSQ_B200_DIR=/path/to/tsq-b200 qmake && make.\n\n")
    with open(OUT, "w", encoding="latin-1") as f:
        f.write(header + "".join(chunks))
    print(f"wrote {OUT}: {sum(len(v) for v in EDITS.values())} edits in {len(EDITS)} files")


if __name__ == "__main__":
    main()
