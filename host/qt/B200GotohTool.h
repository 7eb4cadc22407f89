// B200GotohTool.h -- the adapter written against the REAL tweakseq types (Qt5).  No Qt5 exists in this repository's
// build image, so the file is type-checked against the reference's own headers (original and patched) over Qt stubs
// and executed over functional Qt stand-ins (tests/test_qt_adapter_syntax.py, oracle/ref_shim/); host/B200Gotoh.{h,cpp}
// is its Qt-free twin.  Drop both files into tweakseq/Core/ and apply host/qt/tweakseq_registration.patch
// (INTEGRATION.md section 4).
#ifndef __B200_GOTOH_TOOL_H_
#define __B200_GOTOH_TOOL_H_

#include <QObject>
#include <QThread>
#include <QStringList>

#include "AlignmentTool.h"   // tweakseq/Core/AlignmentTool.h; the patch adds inProcess()/run() as virtuals

// with the patched header the two additions are checked overrides; against the original they are new virtuals
#ifdef TWEAKSEQ_ALIGNMENTTOOL_INPROCESS
#define TSQ_OVERRIDE override
#else
#define TSQ_OVERRIDE
#endif

class B200GotohTool : public AlignmentTool
{
	public:
		B200GotohTool();
		virtual ~B200GotohTool();

		virtual void makeCommand(QString &, QString &, QString &, QStringList &) override;
		virtual void writeSettings(QDomDocument &, QDomElement &) override;
		virtual void readSettings(QDomDocument &) override;

		virtual bool inProcess() TSQ_OVERRIDE {return true;}
		// fin: FASTA written by Project::exportFASTA; fout: the alignment (alignInProcess) or the distance matrix for
		// clustalo --distmat-in.  logReceiver: an object with an invokable message(QString) -- the worker below -- or 0.
		virtual int run(const QString &fin, const QString &fout, QObject *logReceiver, volatile int *cancel) TSQ_OVERRIDE;
		// The in-memory route: (label, Sequence::filter(true)) of every sequence, taken from the model on the GUI
		// thread.  The rows written to fout carry ">label", the key Project::readNewAlignment matches by
		// (FASTAFile.cpp:177-187) -- exportFASTA writes `comment` as the header instead (Project.cpp:876-880), which
		// loses renamed sequences and PDB imports.  fout: the multiple alignment, FASTA, tree order.
		int run(const QStringList &labels, const QStringList &residues, const QString &fout, QObject *logReceiver, volatile int *cancel);

		int gapOpen, gapExtend, device;
		int devices;   // B200s of the box one job uses (tsq_params.n_devices; -1 = all of them)
		int alphabet;  // TSQ_ALPHABET_AUTO (default: decided from the residues), TSQ_PROTEIN, TSQ_NUCLEOTIDE;
		               // Project::sequenceDataType() may set it: SequenceFile::DNA -> TSQ_NUCLEOTIDE, ::Proteins -> TSQ_PROTEIN
		bool alignInProcess; // fout = the multiple alignment readNewAlignment ingests (no clustalo needed); else the matrix

	private:
		void init();
		void getVersion();
};

// Runs the tool off the GUI thread and reports like QProcess: finished(exit code, exit status), log lines as
// message(QString) -- both emitted from the worker thread, delivered queued to the main window's slots
// (alignmentFinishedInProcess / alignmentMessage in the patch).
class B200GotohWorker : public QThread
{
	Q_OBJECT
	public:
		B200GotohWorker(B200GotohTool *t, const QString &fin, const QString &fout, QObject *parent = 0);
		B200GotohWorker(B200GotohTool *t, const QStringList &labels, const QStringList &residues, const QString &fout, QObject *parent = 0);
		volatile int cancel;
	public slots:
		void requestCancel(){cancel=1;} // what alignmentStop() (SeqEditMainWin.cpp:803-812) triggers instead of kill()
	signals:
		void message(const QString &);
		void finished(int exitCode, int exitStatus);
	protected:
		void run() override;
	private:
		B200GotohTool *tool;
		QString fin_, fout_;
		QStringList labels_, residues_;
		bool inMemory_;
};

#endif
