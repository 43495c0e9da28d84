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

#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

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
        //  Fast path for the common line of bases: everything up to the next LF is visible (checked 64 bytes at a
        //  time in a loop the compiler vectorises); anything else -- CR, blanks inside the line, a stray byte -- stops
        //  the fast path in front of it and the byte-wise scan takes over.
        {
          const uint8_t *lf = (const uint8_t *)memchr(q, '\n', e - q);
          const uint8_t *lim = lf ? lf : e;
          while (lim - q >= 64) {
            uint8_t bad = 0;
            for (int i = 0; i < 64; i++) bad |= (uint8_t)((uint8_t)(q[i] - '!') > (uint8_t)('~' - '!'));
            if (bad) break;
            q += 64;
          }
        }
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


//  Parse one block of a file on several host threads.  The block is cut where a record seems to start (a line that
//  begins with '>' or '@'; for '@' the line after next must begin with '+', because a quality line may begin with '@'
//  too); every piece but the first is parsed by its own SeqParser from state A into its own buffer.  Afterwards the
//  pieces are stitched in order and the guess is VERIFIED: the state the machine really has at a cut must be one from
//  which a '>' / '@' starts a record (A, or Q = behind a quality line, or G = behind a sequence line, which first ends
//  the record); if it is not, that piece is parsed again, sequentially, from the true state.  The result is therefore
//  exactly what one sequential pass produces.  `ps` carries the state from block to block.
int mb_parse_parallel(SeqParser::Sink out, void *outArg, SeqParser &ps, const uint8_t *d, size_t n) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  size_t minPiece = (size_t)4 << 20;
  if (getenv("MB_PARSE_THREADS"))   nt = (unsigned)atoi(getenv("MB_PARSE_THREADS"));          //  testing knobs
  if (getenv("MB_PARSE_MIN_PIECE")) minPiece = (size_t)strtoull(getenv("MB_PARSE_MIN_PIECE"), nullptr, 10);
  if (minPiece < 16) minPiece = 16;
  if (nt < 2 || n < 2 * minPiece) return ps.feed(d, n);
  if (n / nt < minPiece) nt = (unsigned)(n / minPiece);

  //  cut points
  std::vector<size_t> cut(1, 0);
  for (unsigned i = 1; i < nt; i++) {
    size_t p = n / nt * i;
    if (p <= cut.back()) continue;
    bool found = false;
    for (int tries = 0; tries < 64 && !found; tries++) {          //  next line start that looks like a record start
      const uint8_t *lf = (const uint8_t *)memchr(d + p, '\n', n - p);
      if (!lf || (size_t)(lf - d) + 1 >= n) break;
      p = (size_t)(lf - d) + 1;
      if (d[p] == '>') found = true;
      else if (d[p] == '@') {
        const uint8_t *l1 = (const uint8_t *)memchr(d + p, '\n', n - p);
        const uint8_t *l2 = l1 ? (const uint8_t *)memchr(l1 + 1, '\n', n - (size_t)(l1 + 1 - d)) : nullptr;
        if (l2 && (size_t)(l2 + 1 - d) < n && l2[1] == '+') found = true;
      }
    }
    if (found && p > cut.back()) cut.push_back(p);
  }
  cut.push_back(n);
  const size_t np = cut.size() - 1;
  if (np < 2) return ps.feed(d, n);

  struct Piece { std::string *outp = nullptr; int rc = MB_OK; SeqParser::State end = SeqParser::A; std::string err; };
  std::vector<Piece> pc(np);
  if (ps.pool().size() < np) ps.pool().resize(np);
  for (size_t i = 0; i < np; i++) { pc[i].outp = &ps.pool()[i]; pc[i].outp->clear(); }
  auto sink = [](void *a, const char *b, size_t m) -> int { ((std::string *)a)->append(b, m); return 0; };
  const SeqParser::State st0 = ps.state();
  std::vector<std::thread> th;
  for (size_t i = 0; i < np; i++)
    th.emplace_back([&, i]() {
      Piece &P = pc[i];
      P.outp->reserve((cut[i + 1] - cut[i]) / 2 + 4096);
      SeqParser q(sink, P.outp);
      q.setState(i == 0 ? st0 : SeqParser::A);
      P.rc = q.feed(d + cut[i], cut[i + 1] - cut[i]);
      if (P.rc != MB_OK) P.err = mb_last_error();
      P.end = q.state();
    });
  for (auto &x : th) x.join();

  //  stitch + verify
  SeqParser::State cur = st0;
  for (size_t i = 0; i < np; i++) {
    bool ok = (i == 0);
    if (i > 0) {
      if (cur == SeqParser::G) { MB_TRY(out(outArg, ".", 1)); cur = SeqParser::A; }     //  what G does on '>' / '@'
      ok = (cur == SeqParser::A || cur == SeqParser::Q);
    }
    if (ok) {
      if (pc[i].rc != MB_OK) return mb_fail(pc[i].rc, "%s", pc[i].err.c_str());
      //  hand the piece on in bounded slices (the sink copies into the pinned staging buffers)
      for (size_t at = 0; at < pc[i].outp->size(); at += (size_t)64 << 20)
        MB_TRY(out(outArg, pc[i].outp->data() + at, std::min<size_t>((size_t)64 << 20, pc[i].outp->size() - at)));
      cur = pc[i].end;
    } else {                                                 //  the cut was not at a record start: sequentially from the true state
      SeqParser q(out, outArg);
      q.setState(cur);
      MB_TRY(q.feed(d + cut[i], cut[i + 1] - cut[i]));
      cur = q.state();
    }
  }
  ps.setState(cur);
  return MB_OK;
}

