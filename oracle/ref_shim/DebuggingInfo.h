/* stands in for tweakseq/Core/DebuggingInfo.h: Consensus.cpp only prints trace.header(...) to qDebug() */
#ifndef TSQ_REF_DEBUGGINGINFO_H
#define TSQ_REF_DEBUGGINGINFO_H
struct DebuggingInfo { const char* header(const char* s = "") { return s; } };
[[maybe_unused]] static DebuggingInfo trace;
#endif
