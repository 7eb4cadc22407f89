// B200GotohTool.h -- the adapter written against the REAL tweakseq types (Qt5).  Not compiled in
// this repository (no Qt5 in the build image); host/B200Gotoh.{h,cpp} is its Qt-free twin and is
// what the tests exercise.  Drop both files into tweakseq/Core/, add them to tweakseq.pro
// (HEADERS/SOURCES, LIBS += -ltsqb200) and apply the edits listed in INTEGRATION.md.
#ifndef __B200_GOTOH_TOOL_H_
#define __B200_GOTOH_TOOL_H_

#include <QObject>
#include <QThread>

#include "AlignmentTool.h"   // tweakseq/Core/AlignmentTool.h, with inProcess()/run() added

class B200GotohTool : public AlignmentTool
{
	public:
		B200GotohTool();
		virtual ~B200GotohTool();

		virtual void makeCommand(QString &, QString &, QString &, QStringList &);
		virtual void writeSettings(QDomDocument &, QDomElement &);
		virtual void readSettings(QDomDocument &);

		virtual bool inProcess(){return true;}
		// fin: FASTA written by Project::exportFASTA; fout: the alignment (alignInProcess) or the distance matrix for clustalo --distmat-in
		virtual int run(const QString &fin, const QString &fout, QObject *logReceiver, volatile int *cancel);

		int gapOpen, gapExtend, device;
		int devices;   // B200s of the box one job uses (tsq_params.n_devices; -1 = all of them)
		int alphabet;  // TSQ_ALPHABET_AUTO (default: decided from the exported residues), TSQ_PROTEIN, TSQ_NUCLEOTIDE;
		               // Project::sequenceDataType() may set it: SequenceFile::DNA -> TSQ_NUCLEOTIDE, ::Proteins -> TSQ_PROTEIN
		bool alignInProcess; // fout = the multiple alignment readNewAlignment ingests (no clustalo needed); else the matrix

	private:
		void init();
		void getVersion();
};

// Runs tool->run() off the GUI thread and reports like QProcess::finished(int, ExitStatus)
class B200GotohWorker : public QThread
{
	Q_OBJECT
	public:
		B200GotohWorker(B200GotohTool *t, const QString &fin, const QString &fout, QObject *parent = 0);
		volatile int cancel;
	public slots:
		void requestCancel(){cancel=1;} // what alignmentStop() (SeqEditMainWin.cpp:803-812) triggers instead of kill()
	signals:
		void message(const QString &);
		void finished(int exitCode, int exitStatus);
	protected:
		void run();
	private:
		B200GotohTool *tool;
		QString fin_, fout_;
};

#endif
