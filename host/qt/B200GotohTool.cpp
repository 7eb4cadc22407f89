// B200GotohTool.cpp -- see B200GotohTool.h.  Follows tweakseq/Core/ClustalO.cpp:48-111.
#include <QDomDocument>
#include <QMetaObject>
#include <QByteArray>

#include <cstdio>
#include <vector>

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
	// MessageWin::addMessage is a plain method (UI/MessageWin.h:42), not a slot: the line leaves as the worker's
	// message(QString) SIGNAL (signals are invokable), and the main window's alignmentMessage slot, connected to
	// it, calls mw->addMessage on the GUI thread
	QObject *rcv = static_cast<QObject *>(user);
	if (rcv)
		QMetaObject::invokeMethod(rcv, "message", Qt::DirectConnection, Q_ARG(QString, QString::fromUtf8(line)));
}

static void fillParams(tsq_params &p, const B200GotohTool &t)
{
	tsq_default_params(&p);
	p.gap_open = t.gapOpen;
	p.gap_extend = t.gapExtend;
	p.device = t.device;
	p.n_devices = t.devices;
	p.alphabet = t.alphabet;
}

int B200GotohTool::run(const QString &fin, const QString &fout, QObject *logReceiver, volatile int *cancel)
{
	tsq_params p;
	fillParams(p, *this); // alphabet auto: tsq_run_fasta decides from the residues, as clustalo does without --seqtype
	if (alignInProcess) p.flags |= TSQ_FLAG_MSA_OUT;
	return tsq_run_fasta(fin.toLocal8Bit().constData(), fout.toLocal8Bit().constData(), &p, forwardLog, logReceiver, cancel);
}

int B200GotohTool::run(const QStringList &labels, const QStringList &residues, const QString &fout, QObject *logReceiver, volatile int *cancel)
{
	const int n = residues.size();
	if (labels.size() != n) return TSQ_ERR_INVALID;
	std::vector<QByteArray> res, hdr;
	for (int i = 0; i < n; ++i){
		res.push_back(residues.at(i).toLocal8Bit());
		hdr.push_back((QString(">") + labels.at(i)).toLocal8Bit()); // the label, not the comment (Project.cpp:876-880)
	}
	std::vector<const char *> rp, hp;
	std::vector<uint32_t> rl;
	for (int i = 0; i < n; ++i){ // pointers only once the vectors no longer grow
		rp.push_back(res[i].constData());
		rl.push_back((uint32_t) res[i].size());
		hp.push_back(hdr[i].constData());
	}
	char line[256];
	tsq_params p;
	fillParams(p, *this);
	if (p.alphabet == TSQ_ALPHABET_AUTO)
		p.alphabet = tsq_detect_alphabet(rp.data(), rl.data(), (uint32_t) n);
	tsq_ctx *c = 0;
	int rc = tsq_create(&c, &p);
	if (rc != TSQ_OK){
		forwardLog(logReceiver, tsq_status_string(rc));
		return rc;
	}
	snprintf(line, sizeof line, "tsq-b200: %d sequences from the project (%s)", n, p.alphabet == TSQ_NUCLEOTIDE ? "nucleotide" : "protein");
	forwardLog(logReceiver, line);
	rc = tsq_set_sequences(c, rp.data(), rl.data(), (uint32_t) n);
	if (rc == TSQ_OK) rc = tsq_run(c, 0, 0, cancel);
	if (rc == TSQ_OK) rc = tsq_write_msa_fasta(c, hp.data(), rp.data(), rl.data(), fout.toLocal8Bit().constData(), 1);
	if (rc == TSQ_OK){
		tsq_stats st;
		tsq_get_stats(c, &st);
		snprintf(line, sizeof line, "tsq-b200: %llu pairs, kernel %.3f ms (%.1f GCUPS), alignment %.1f ms", (unsigned long long) st.n_pairs,
			st.kernel_ms, st.gcups_kernel, st.msa_ms);
		forwardLog(logReceiver, line);
	}
	else
		forwardLog(logReceiver, tsq_last_error(c));
	tsq_destroy(c);
	return rc;
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
	: QThread(parent), cancel(0), tool(t), fin_(fin), fout_(fout), inMemory_(false)
{
}

B200GotohWorker::B200GotohWorker(B200GotohTool *t, const QStringList &labels, const QStringList &residues, const QString &fout, QObject *parent)
	: QThread(parent), cancel(0), tool(t), fout_(fout), labels_(labels), residues_(residues), inMemory_(true)
{
}

void B200GotohWorker::run()
{
	int rc = inMemory_ ? tool->run(labels_, residues_, fout_, this, &cancel) : tool->run(fin_, fout_, this, &cancel);
	emit finished(rc == TSQ_ERR_CANCELLED ? 9 : rc, 0 /* QProcess::NormalExit */); // 9: "user interrupted", SeqEditMainWin.cpp:852-857
}
