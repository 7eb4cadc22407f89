/* stands in for tweakseq/Core/Sequence.h: Consensus.cpp reads Sequence::residues (16-bit cells, Sequence.h:36-39) */
#ifndef TSQ_REF_SEQUENCE_H
#define TSQ_REF_SEQUENCE_H
#include "../qt_min.h"
class Sequence { public: QString residues; };
#endif
