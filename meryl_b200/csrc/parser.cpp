//  P1/P2: FASTA/FASTQ -> base stream, on the host.
//
//  A byte-level state machine with the behaviour of the reference's reader stack
//    bufSeqFile::open (format sniffing)      src/utility/src/sequence/bufSeqFile-v1.C:54-104
//    bufSeqFile::loadBases                   src/utility/src/sequence/bufSeqFile-v1.C:554-595
//    readBuffer::skip / copy / skipLine      src/utility/src/files/buffered-v1-reading.H:232-322
//    loadBases() work-buffer builder         src/meryl/merylOp-countThreads.C:139-231
//  but restartable at any byte, so a file can be streamed through it in pieces of any size.
//  What it emits is the reference's logical work buffer: the visible characters of each
//  record's sequence lines, a '.' after every record and at the end of every file.  (The
//  reference cuts that stream into 2 MB buffers with a k-1 carry; the cut does not change the
//  k-mers.)  Runs of '.' may differ in length from the reference's; k-mers cannot.
//
//  States (c = next byte):
//    A     between records, skipping white space     '+' -> P1, '>'/'@' -> H, visible -> W
//    P1    in the '+' line;  P2  in the quality line (any content, ends at CR, LF or CRLF)
//    Q     first byte after the quality line         '>'/'@' -> H, otherwise as L
//    H     in a name line
//    L     before the first word of a record         every visible byte is sequence
//    W     inside a word                             white space -> G
//    G     between words                             '>','@','+' end the record, then as A
//  *CR    saw CR at the end of a skipped line; an LF is swallowed, anything else is re-read.
//  A byte that is neither visible ('!'..'~') nor white space inside a sequence makes the
//  reference spin forever (copyVisible stops, skipWhitespace does not advance); here it is
//  an MB_EFORMAT error.

#include <string.h>

#include "common.h"

namespace {

inline bool isWs (uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }   //  SP HT LF VT FF CR
inline bool isVis(uint8_t c) { return c >= '!' && c <= '~'; }

//  First CR or LF in [p, e), or e.
inline const uint8_t *findEol(const uint8_t *p, const uint8_t *e) {
  const uint8_t *lf = (const uint8_t *)memchr(p, '\n', e - p);
  const uint8_t *lim = lf ? lf : e;
  const uint8_t *cr = (const uint8_t *)memchr(p, '\r', lim - p);
  return cr ? cr : lim;
}

}  //  namespace


bool SeqParser::sniff(const uint8_t *d, size_t n) {
  size_t p = 0;
  while (p < n && isWs(d[p])) p++;
  if (p == n) return true;                                 //  empty: opens, yields nothing
  if (d[p] != '>' && d[p] != '@') return false;
  if (d[p] == '>') return true;

  const uint8_t *e = d + n;                                //  '@': FASTQ or SAM?
  const uint8_t *q = findEol(d + p, e);
  if (q < e) { if (*q == '\r' && q + 1 < e && q[1] == '\n') q += 2; else q += 1; }
  p = q - d;
  while (p < n && isWs(d[p])) p++;
  if (p < n && d[p] == '@') return false;                  //  second header line: SAM
  while (p < n && d[p] != '\n' && p < 1000) {
    if (d[p] == '\t') return false;                        //  tab in the 2nd line: SAM
    p++;
  }
  while (p < n && isWs(d[p])) p++;
  return p == 1000 || (p < n && d[p] == '+');
}


int SeqParser::feed(const uint8_t *d, size_t n) {
  const uint8_t *p = d, *e = d + n;

  while (p < e) {
    switch (_st) {

      case A:
        while (p < e && isWs(*p)) p++;
        if (p == e) break;
        if      (*p == '+')               { _st = P1; p++; }
        else if (*p == '>' || *p == '@')  { _st = H;  p++; }
        else                                _st = L;
        break;

      case P1: case P2: case H: {
        const uint8_t *q = findEol(p, e);
        p = q;
        if (p == e) break;
        const State next = (_st == P1) ? P2 : (_st == P2) ? Q : L;
        if (*p == '\r') _st = (_st == P1) ? P1CR : (_st == P2) ? P2CR : HCR;
        else            _st = next;
        p++;
        break;
      }

      case P1CR: if (*p == '\n') p++;  _st = P2;  break;
      case P2CR: if (*p == '\n') p++;  _st = Q;   break;
      case HCR:  if (*p == '\n') p++;  _st = L;   break;

      case Q:
        if (*p == '>' || *p == '@') { _st = H; p++; }
        else                          _st = L;
        break;

      case L:
        while (p < e && isWs(*p)) p++;
        if (p == e) break;
        if (!isVis(*p)) return mb_fail(MB_EFORMAT, "byte 0x%02x inside a sequence (the reference does not terminate on it)", *p);
        _st = W;
        break;

      case W: {
        const uint8_t *q = p;
        while (q < e && isVis(*q)) q++;
        MB_TRY(emit(p, q - p));
        p = q;
        if (p == e) break;
        if (!isWs(*p)) return mb_fail(MB_EFORMAT, "byte 0x%02x inside a sequence (the reference does not terminate on it)", *p);
        _st = G;
        break;
      }

      case G:
        while (p < e && isWs(*p)) p++;
        if (p == e) break;
        if (*p == '>' || *p == '@' || *p == '+') { MB_TRY(brk()); _st = A; }
        else if (isVis(*p))                         _st = W;
        else return mb_fail(MB_EFORMAT, "byte 0x%02x inside a sequence (the reference does not terminate on it)", *p);
        break;
    }
  }
  return MB_OK;
}


int SeqParser::finish(void) {
  _st = A;
  return brk();                                            //  merylOp-countThreads.C:196
}
