/* stands in for tweakseq/Core/Sequences.h: Consensus.cpp calls Sequences::sequences() */
#ifndef TSQ_REF_SEQUENCES_H
#define TSQ_REF_SEQUENCES_H
#include "Sequence.h"
class Sequences {
 public:
  QList<Sequence*>& sequences() { return list_; }
  QList<Sequence*> list_;
};
#endif
