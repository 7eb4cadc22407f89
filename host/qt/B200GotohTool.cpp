// B200GotohTool.cpp -- see B200GotohTool.h.  Follows tweakseq/Core/ClustalO.cpp:48-111.
#include <QDomDocument>
#include <QMetaObject>

#include "B200GotohTool.h"
#include "XMLHelper.h"
#include "tsq_b200.h"

B200GotohTool::B200GotohTool()
{
	init();
}

B200GotohTool::~B200GotohTool()
{
}

void B200GotohTool::makeCommand(QString &fin, QString &fout, QString &exec, QStringList &arglist)
{
	// only used when the matrix is handed on to clustalo (ClustalO.cpp:51 + --distmat-in)
	exec = executable_;
	arglist << "--force" << "-v" << "--outfmt=fa" << "--output-order=tree-order" << "-i" << fin << "--distmat-in" << fout;
}

void B200GotohTool::writeSettings(QDomDocument &doc, QDomElement &parentElem)
{
	QDomElement pelem = doc.createElement("alignment_tool");
	parentElem.appendChild(pelem);
	XMLHelper::addElement(doc, pelem, "name", name());
	XMLHelper::addElement(doc, pelem, "path", executable());
	XMLHelper::addElement(doc, pelem, "preferred", (preferred() ? "yes" : "no"));
	XMLHelper::addElement(doc, pelem, "gap_open", QString::number(gapOpen));
	XMLHelper::addElement(doc, pelem, "gap_extend", QString::number(gapExtend));
	XMLHelper::addElement(doc, pelem, "device", QString::number(device));
	XMLHelper::addElement(doc, pelem, "devices", QString::number(devices));
	XMLHelper::addElement(doc, pelem, "alphabet", (alphabet == TSQ_NUCLEOTIDE ? "nucleotide" : alphabet == TSQ_PROTEIN ? "protein" : "auto"));
	XMLHelper::addElement(doc, pelem, "align_in_process", (alignInProcess ? "yes" : "no"));
}

void B200GotohTool::readSettings(QDomDocument &doc)
{
	QDomNodeList nl = doc.elementsByTagName("alignment_tool");
	for (int i = 0; i < nl.count(); ++i){
		QDomElement elem = nl.item(i).firstChildElement();
		while (!elem.isNull()){
			if (elem.tagName() == "name" && elem.text() != name_)
				break;
			if (elem.tagName() == "path") executable_ = elem.text();
			if (elem.tagName() == "preferred") setPreferred(elem.text() == "yes");
			if (elem.tagName() == "gap_open") gapOpen = elem.text().toInt();
			if (elem.tagName() == "gap_extend") gapExtend = elem.text().toInt();
			if (elem.tagName() == "device") device = elem.text().toInt();
			if (elem.tagName() == "devices") devices = elem.text().toInt();
			if (elem.tagName() == "alphabet") alphabet = (elem.text() == "nucleotide" ? TSQ_NUCLEOTIDE : elem.text() == "protein" ? TSQ_PROTEIN : TSQ_ALPHABET_AUTO);
			if (elem.tagName() == "align_in_process") alignInProcess = (elem.text() == "yes");
			elem = elem.nextSiblingElement();
		}
	}
	getVersion();
}

static void forwardLog(void *user, const char *line)
{
	QObject *rcv = static_cast<QObject *>(user);
	if (rcv) // MessageWin::addMessage(QString) lives on the GUI thread
		QMetaObject::invokeMethod(rcv, "addMessage", Qt::QueuedConnection, Q_ARG(QString, QString::fromUtf8(line)));
}

int B200GotohTool::run(const QString &fin, const QString &fout, QObject *logReceiver, volatile int *cancel)
{
	tsq_params p;
	tsq_default_params(&p);
	p.gap_open = gapOpen;
	p.gap_extend = gapExtend;
	p.device = device;
	p.n_devices = devices;
	p.alphabet = alphabet; // auto: tsq_run_fasta decides from the residues, as clustalo does without --seqtype
	if (alignInProcess) p.flags |= TSQ_FLAG_MSA_OUT;
	return tsq_run_fasta(fin.toLocal8Bit().constData(), fout.toLocal8Bit().constData(), &p, forwardLog, logReceiver, cancel);
}

void B200GotohTool::init()
{
	name_ = "b200gotoh";
	version_ = "";
	executable_ = "libtsqb200.so";
	gapOpen = -1;
	gapExtend = -1;
	device = 0;
	devices = 1;
	alphabet = TSQ_ALPHABET_AUTO;
	alignInProcess = true;
}

void B200GotohTool::getVersion()
{
	version_ = QString(tsq_version_string());
}

B200GotohWorker::B200GotohWorker(B200GotohTool *t, const QString &fin, const QString &fout, QObject *parent)
	: QThread(parent), cancel(0), tool(t), fin_(fin), fout_(fout)
{
}

void B200GotohWorker::run()
{
	int rc = tool->run(fin_, fout_, parent(), &cancel);
	emit finished(rc == TSQ_ERR_CANCELLED ? 9 : rc, 0 /* QProcess::NormalExit */);
}
