//  TEST INFRASTRUCTURE (oracle/) -- not part of the shipped product.
//
//  Stand-in for the reference's BAM/CRAM reader so that the reference `meryl`
//  binary links without the vendored htslib (which needs bzlib/lzma/curl
//  headers that this image does not have).  Every method reports "cannot
//  open", so openSequenceFile() (utility/src/sequence/dnaSeqFile-v1.H:257-275),
//  which tries bufSeqFile first, never reaches it for FASTA/FASTQ input.
//  Class declaration being satisfied: utility/src/sequence/dnaSeqFile-v1.H:221-252.

#include "sequence.H"

namespace merylutil::inline sequence::inline v1 {

bool   htsSeqFile::open(char const *, bool)  { return false; }
bool   htsSeqFile::close(void)               { return false; }
bool   htsSeqFile::reopen(void)              { return false; }
bool   htsSeqFile::loadIndex(bool)           { return false; }
void   htsSeqFile::unloadIndex(void)         { }
void   htsSeqFile::createIndex(void)         { }
void   htsSeqFile::destroyIndex(void)        { }
bool   htsSeqFile::findSequence(uint64)      { return false; }
uint64 htsSeqFile::sequenceLength(uint64)    { return 0; }

bool   htsSeqFile::loadSequence(char *&, uint32 &, char *&, uint8 *&, uint64 &, uint64 &, uint32 &) { return false; }
bool   htsSeqFile::loadBases(char *, uint64, uint64 &, bool &) { return false; }

}
