//  meryl-lookup on a B200: `meryl-lookup-b200 <report> -sequence <s1> [s2] -mers <db> [db ...] -output <o1> [o2]`
//  (also `meryl-b200 lookup ...`).
//
//  The reports of src/meryl-lookup that count k-mers per sequence:
//      -existence   one line per sequence: name, its k-mers, and per database the database's k-mers and how many of the
//                   sequence's k-mers it holds (existence.C:59-116)
//      -include     the sequences (or pairs) with at least one k-mer in the database, as FASTA / FASTQ records named
//      -exclude     "<ident> nKmers=<found>"; the ones with none (include-exclude.C:62-134); -10x: the k-mers starting in the
//                   first 23 positions of the first file's reads are not looked up
//  with the reference's options -min / -max (value window of the table), -threads / -loadthreads / -memory / -V (accepted;
//  they size host resources of the reference that do not exist here).  The per-position reports (-bed, -bed-runs, -wig-count,
//  -wig-depth: dump.C) are written from mb_position_lookup's per-position flags and values.  Option handling and the usage errors follow meryl-lookup.C:155-368.
//
//  The sequences are read on the host the way the reference's loader reads them (bufSeqFile::loadSequence / loadFASTA /
//  loadFASTQ, src/utility/src/sequence/bufSeqFile-v1.C:291-549; names: dnaSeq::findNameAndFlags, dnaSeq-v1.C:89-101); the
//  counting is kernel K10 (csrc/union.cu) against the database kept resident on the device (mb_table_open / mb_table_query),
//  the sequences passing in batches.

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <string>
#include <vector>

#include "../../include/meryl_b200.h"

namespace {

struct Sequences {                                      //  one input file
  std::vector<std::string> ident;
  std::string              bases, quals;                //  all records back to back (quals as in the file; empty for FASTA)
  std::vector<uint64_t>    seqStart, qualStart;         //  n + 1 entries
  Sequences() : seqStart(1, 0), qualStart(1, 0) {}
  size_t size(void) const { return ident.size(); }
};

inline bool isWhite(char c)   { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }   //  types-v1.H:258-263
inline bool isVisible(char c) { return c >= '!' && c <= '~'; }

std::string shellQuoted(const char *path) {              //  inside single quotes, a quote in it as '\''
  std::string q = "'";
  for (const char *c = path; *c; c++) { if (*c == '\'') q += "'\\''"; else q += *c; }
  return q + "'";
}

bool readAll(const char *path, std::vector<char> &out) {
  const size_t L = strlen(path);
  const char *tool = nullptr;
  if      (L > 3 && !strcmp(path + L - 3, ".gz"))   tool = "gzip -dc";
  else if (L > 4 && !strcmp(path + L - 4, ".bz2"))  tool = "bzip2 -dc";
  else if (L > 3 && !strcmp(path + L - 3, ".xz"))   tool = "xz -dc";
  FILE *f = nullptr;
  if (tool) {
    if (access(path, R_OK) != 0) return false;
    std::string cmd = std::string(tool) + " " + shellQuoted(path);
    f = popen(cmd.c_str(), "r");
  } else f = !strcmp(path, "-") ? stdin : fopen(path, "rb");
  if (!f) return false;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.insert(out.end(), buf, buf + n);
  if (tool) pclose(f); else if (f != stdin) fclose(f);
  return true;
}

//  bufSeqFile::loadSequence over a whole file in memory.  A NUL byte ends the input, as in the reference's buffer.
void parseSequences(const std::vector<char> &d, Sequences &S) {
  size_t pos = 0;
  const size_t n = d.size();
  auto peek = [&]() -> char { return pos < n ? d[pos] : 0; };
  auto next = [&]() -> char { const char c = peek(); if (pos < n) pos++; return c; };
  for (;;) {
    while (isWhite(peek())) next();
    while (peek() != '>' && peek() != '@' && peek() != 0) next();          //  lost sync: on to the next record start
    if (peek() == 0) break;
    const bool fastq = (peek() == '@');
    std::string name, seq, qlt;
    bool ok = true;
    char ch = next();                                                       //  the '>' or '@'
    for (ch = next(); ch != '\n' && ch != 0; ch = next()) if (ch != '\r') name.push_back(ch);
    if (!fastq) {
      for (ch = peek(); ch != '>' && ch != '@' && ch != 0; ch = peek()) {
        ch = next();
        if (!isWhite(ch)) seq.push_back(ch);
      }
    } else {
      while (!name.empty() && isWhite(name.back())) name.pop_back();
      while (isWhite(ch)) ch = next();
      for (; ch != '\n' && ch != 0; ch = next()) if (!isWhite(ch)) seq.push_back(ch);
      while (isWhite(ch)) ch = next();
      if (ch != '+') ok = false;
      else {
        for (ch = next(); ch != '\n' && ch != 0; ch = next()) { }
        while (isWhite(ch)) ch = next();
        for (; ch != '\n' && ch != 0; ch = next()) if (!isWhite(ch)) qlt.push_back(ch);
        while (isWhite(peek())) next();
      }
    }
    if (!ok) { seq.clear(); qlt.clear(); }                                  //  the name stays, the sequence is empty
    size_t h = 0;
    while (h < name.size() && isWhite(name[h])) h++;
    size_t e = h;
    while (e < name.size() && isVisible(name[e])) e++;
    S.ident.push_back(name.substr(h, e - h));
    S.bases += seq;  S.quals += qlt;
    S.seqStart.push_back(S.bases.size());
    S.qualStart.push_back(S.quals.size());
  }
}

FILE *openOut(const char *name, bool &piped) {
  piped = false;
  if (!strcmp(name, "-")) return stdout;
  const size_t L = strlen(name);
  if (L > 3 && !strcmp(name + L - 3, ".gz")) {
    std::string cmd = std::string("gzip -1c > ") + shellQuoted(name);
    piped = true;
    return popen(cmd.c_str(), "w");
  }
  return fopen(name, "wb");
}

int closeOut(FILE *f, bool piped) {
  if (!f || f == stdout) return fflush(stdout);
  return piped ? pclose(f) : fclose(f);
}

//  include-exclude.C:100-111 with fasta-fastq-v1.C:28-138: FASTQ when the first quality value is above '!', else FASTA on one line
void writeRecord(FILE *o, const Sequences &S, size_t i, uint64_t nFound) {
  if (!o) return;
  const char  *b  = S.bases.data() + S.seqStart[i];
  const size_t sl = (size_t)(S.seqStart[i + 1] - S.seqStart[i]);
  const char  *q  = S.quals.data() + S.qualStart[i];
  const size_t ql = (size_t)(S.qualStart[i + 1] - S.qualStart[i]);
  if (ql == 0 || q[0] == '!') {
    fprintf(o, ">%s nKmers=%llu\n", S.ident[i].c_str(), (unsigned long long)nFound);
    if (sl) { fwrite(b, 1, sl, o); fputc('\n', o); }
  } else {
    fprintf(o, "@%s nKmers=%llu\n", S.ident[i].c_str(), (unsigned long long)nFound);
    fwrite(b, 1, sl, o);
    fputs("\n+\n", o);
    for (size_t j = 0; j < sl; j++) fputc(j < ql ? q[j] : '!', o);
    fputc('\n', o);
  }
}

//  One database against all sequences of a file: the database is decoded once and stays on the device (mb_table_open) while
//  the sequences go by in batches of ~256 MB of letters (MB_LOOKUP_BATCH: testing knob); a database too large to stay
//  resident falls back to the one-shot calls, which decode it group of files by group of files for every batch.
int askDatabase(const char *db, uint32_t minV, uint32_t maxV, const Sequences &S, uint32_t skipFirst,
                uint64_t *total, uint64_t *found, uint64_t *dbKmers, uint8_t *flags, uint32_t *values) {
  size_t batch = (size_t)256 << 20;
  if (getenv("MB_LOOKUP_BATCH")) batch = (size_t)strtoull(getenv("MB_LOOKUP_BATCH"), nullptr, 10);
  mb_table *T = nullptr;
  const int orc = getenv("MB_LOOKUP_NO_TABLE") ? MB_ENOMEM : mb_table_open(db, minV, maxV, -1, 0, &T);     //  (testing knob: the fallback)
  if (orc != MB_OK && orc != MB_ENOMEM) return orc;
  if (T && dbKmers) *dbKmers = mb_table_kmers(T);
  int rc = MB_OK;
  const size_t n = S.size();
  std::vector<uint64_t> rel;
  bool first = true;
  for (size_t i0 = 0; (i0 < n || first) && rc == MB_OK; first = false) {
    size_t i1 = i0;
    while (i1 < n && (i1 == i0 || S.seqStart[i1 + 1] - S.seqStart[i0] <= batch)) i1++;
    rel.assign(i1 - i0 + 1, 0);
    for (size_t i = i0; i <= i1; i++) rel[i - i0] = S.seqStart[i] - S.seqStart[i0];
    const char *b = S.bases.data() + S.seqStart[i0];
    uint64_t *t = total ? total + i0 : nullptr, *f = found ? found + i0 : nullptr;
    uint8_t  *fl = flags ? flags + S.seqStart[i0] : nullptr;
    uint32_t *va = values ? values + S.seqStart[i0] : nullptr;
    if (T) rc = mb_table_query(T, b, rel.data(), i1 - i0, skipFirst, t, f, fl, va);
    else if (flags) rc = mb_position_lookup(db, minV, maxV, b, rel.data(), i1 - i0, fl, va, -1, 0);
    else            rc = mb_existence(db, minV, maxV, b, rel.data(), i1 - i0, skipFirst, t, f, dbKmers, -1, 0);
    i0 = i1;
  }
  mb_table_close(T);
  return rc;
}

enum Report { NONE, EXISTENCE, INCLUDE, EXCLUDE, BED, WIGCOUNT, WIGDEPTH, UNSUPPORTED };

const char *reportName(Report r) {                       //  toString(lookupOp), meryl-lookup.H:38-50
  switch (r) { case BED: return "-bed"; case WIGCOUNT: return "-wig-count"; case WIGDEPTH: return "-wig-depth"; case EXISTENCE: return "-existence"; case INCLUDE: return "-include"; case EXCLUDE: return "-exclude"; default: return "(not supplied)"; }
}

}  //  namespace


int mbLookupMain(int argc, char **argv) {
  const char *seq1 = nullptr, *seq2 = nullptr, *out1 = "-", *out2 = nullptr, *unsupported = nullptr;    //  (output defaults to stdout, meryl-lookup.H:89)
  std::vector<const char *> dbs, labels;
  uint32_t minV = 0, maxV = 0xffffffffu;
  Report report = NONE;
  bool is10x = false, bedRuns = false;
  std::vector<std::string> err;
  for (int a = 1; a < argc; a++) {
    const char *w = argv[a];
    auto more = [&]() { return a + 1 < argc && argv[a + 1][0] != '-'; };
    if      (!strcmp(w, "lookup") && a == 1) { }
    else if (!strcmp(w, "-sequence") && a + 1 < argc) { seq1 = argv[++a]; if (more()) seq2 = argv[++a]; }
    else if (!strcmp(w, "-mers"))     { while (more()) dbs.push_back(argv[++a]); }
    else if (!strcmp(w, "-labels"))   { while (more()) labels.push_back(argv[++a]); }
    else if (!strcmp(w, "-output") && a + 1 < argc)   { out1 = argv[++a]; if (more()) out2 = argv[++a]; }
    else if (!strcmp(w, "-min") && a + 1 < argc)      { minV = (uint32_t)strtoul(argv[++a], 0, 10); }
    else if (!strcmp(w, "-max") && a + 1 < argc)      { maxV = (uint32_t)strtoul(argv[++a], 0, 10); }
    else if ((!strcmp(w, "-threads") || !strcmp(w, "-loadthreads") || !strcmp(w, "-memory")) && a + 1 < argc) { ++a; }
    else if (!strcmp(w, "-existence")) { report = EXISTENCE; }
    else if (!strcmp(w, "-include"))   { report = INCLUDE; }
    else if (!strcmp(w, "-exclude"))   { report = EXCLUDE; }
    else if (!strcmp(w, "-bed"))       { report = BED; }
    else if (!strcmp(w, "-bed-runs"))  { report = BED; bedRuns = true; }
    else if (!strcmp(w, "-wig-count")) { report = WIGCOUNT; }
    else if (!strcmp(w, "-wig-depth")) { report = WIGDEPTH; }
    else if (!strcmp(w, "-estimate"))  { report = UNSUPPORTED; unsupported = w; }
    else if (!strcmp(w, "-10x"))       { is10x = true; }
    else if (!strcmp(w, "-V"))         { }
    else if (!strcmp(w, "-help"))      { err.push_back(""); }               //  (usage and exit code 1, as the reference)
    else err.push_back(std::string("Unknown option '") + w + "'.\n");
  }
  if (report == UNSUPPORTED) {
    const char *ref = getenv("MERYL_LOOKUP_REFERENCE");
    if (ref) { execv(ref, argv); perror("execv"); return 1; }
    fprintf(stderr, "ERROR: '%s' (the memory estimate of the reference's host table) has no meaning here; set MERYL_LOOKUP_REFERENCE=<path to\n", unsupported);
    fprintf(stderr, "       meryl-lookup> to have such command lines passed to the reference binary.\n");
    return 1;
  }
  //  lookupGlobal::checkInvalid, meryl-lookup.C:291-368
  if (report == NONE) err.push_back("No report-type (-bed, -wig-count, -wig-depth, -existence, -include, -exclude) supplied.\n");
  else {
    if (!seq1)        err.push_back("No input sequences (-sequence) supplied.\n");
    if (dbs.empty())  err.push_back("No meryl database (-mers) supplied.\n");
    if (!out1)        err.push_back("No output file (-output) supplied.\n");
    if (report != INCLUDE && report != EXCLUDE) {
      if (seq2) err.push_back(std::string("Only one input sequence (-sequence) supported for ") + reportName(report) + ".\n");
      if (out2) err.push_back(std::string("Only one output file (-output) supported for ") + reportName(report) + ".\n");
    } else {
      if (seq2 && !out2) err.push_back("No second output file (-output) supplied for second input (-input) file.\n");
      if (!seq2 && out2) err.push_back("No second input file (-input) supplied for second output (-output) file.\n");
      if (dbs.size() > 1) err.push_back(std::string("Only one meryl database (-mers) supported for ") + reportName(report) + ".\n");
    }
    if (!labels.empty() && report != BED) err.push_back(std::string("Labels (-labels) not supported for ") + reportName(report) + ".\n");
  }
  if (!err.empty()) {
    fprintf(stderr, "usage: %s <report-type> -sequence <input1.fasta> [<input2.fasta>] -mers <input1.meryl> [<input2.meryl> ...] -output <output1> [<output2>]\n", argv[0]);
    for (const std::string &e : err) fputs(e.c_str(), stderr);
    return 1;
  }

  Sequences S1, S2;
  {
    std::vector<char> raw;
    if (!readAll(seq1, raw)) { fprintf(stderr, "ERROR: cannot read '%s'.\n", seq1); return 1; }
    parseSequences(raw, S1);
  }
  if (seq2) {
    std::vector<char> raw;
    if (!readAll(seq2, raw)) { fprintf(stderr, "ERROR: cannot read '%s'.\n", seq2); return 1; }
    parseSequences(raw, S2);
  }

  bool piped1 = false, piped2 = false;
  FILE *o1 = openOut(out1, piped1), *o2 = out2 ? openOut(out2, piped2) : nullptr;
  if (!o1 || (out2 && !o2)) { fprintf(stderr, "ERROR: cannot open the output.\n"); return 1; }

  if (report == BED || report == WIGCOUNT || report == WIGDEPTH) {
    //  dump.C:86-240 (what is kept per position) and :249-412 (what is written), sequence by sequence
    const size_t n = S1.size(), nBases = S1.bases.size();
    const size_t nDb = (report == WIGDEPTH) ? 1 : dbs.size();           //  (the depth report looks at the first database only)
    const bool   perDb = (report == BED) && !labels.empty();            //  with labels a BED line says which database; without, one line per k-mer
    uint32_t k = 0;
    uint64_t hu = 0, hd = 0, ht = 0, hn = 0;
    if (mb_db_histogram(dbs[0], &k, &hu, &hd, &ht, nullptr, nullptr, 0, &hn) != MB_OK) { fprintf(stderr, "ERROR: %s\n", mb_last_error()); return 1; }
    std::vector<std::vector<uint8_t>> found(perDb ? nDb : 1, std::vector<uint8_t>(nBases, 0));
    std::vector<uint8_t>  flags(nBases, 0), valid(nBases, 0);
    std::vector<uint32_t> count(report == WIGCOUNT ? nBases : 0, 0), vals(report == WIGCOUNT ? nBases : 0, 0);
    for (size_t d = 0; d < nDb; d++) {
      if (askDatabase(dbs[d], minV, maxV, S1, 0, nullptr, nullptr, nullptr, flags.data(), report == WIGCOUNT ? vals.data() : nullptr) != MB_OK) {
        fprintf(stderr, "ERROR: %s\n", mb_last_error());
        return 1;
      }
      std::vector<uint8_t> &f = found[perDb ? d : 0];
      for (size_t p = 0; p < nBases; p++) { f[p] |= (flags[p] >> 1) & 1; valid[p] |= flags[p] & 1; }
      if (report == WIGCOUNT) for (size_t p = 0; p < nBases; p++) count[p] += vals[p];
    }
    char num[64];
    for (size_t i = 0; i < n; i++) {
      const size_t b = (size_t)S1.seqStart[i], len = (size_t)S1.seqStart[i + 1] - b;
      const char *id = S1.ident[i].c_str();
      if (report == BED) {
        size_t maxP = 0;                                              //  one past the last k-mer found in any database
        for (size_t t = 0; t < found.size(); t++) for (size_t p = len; p > maxP; p--) if (found[t][b + p - 1]) { maxP = p; break; }
        if (!bedRuns) {
          for (size_t p = 0; p < maxP; p++)
            for (size_t d = 0; d < found.size(); d++) {
              if (!found[d][b + p]) continue;
              fprintf(o1, "%s\t%zu\t%zu", id, p, p + k);
              if (d < labels.size()) fprintf(o1, "\t%s", labels[d]);
              fputc('\n', o1);
            }
        } else {
          std::vector<size_t> bgn(found.size(), (size_t)-1);
          for (size_t p = 0; p <= maxP; p++)
            for (size_t d = 0; d < found.size(); d++) {
              const bool bit = p < maxP && found[d][b + p];
              if (bit) { if (bgn[d] == (size_t)-1) bgn[d] = p; continue; }
              if (bgn[d] == (size_t)-1) continue;
              fprintf(o1, "%s\t%zu\t%zu", id, bgn[d], p + k);         //  (dump.C:344: the end is written as p + k)
              bgn[d] = (size_t)-1;
              if (d < labels.size()) fprintf(o1, "\t%s", labels[d]);
              fputc('\n', o1);
            }
        }
      } else {
        fprintf(o1, "variableStep chrom=%s\n", id);
        if (report == WIGCOUNT) {
          for (size_t p = 0; p < len; p++) if (valid[b + p] && count[b + p]) fprintf(o1, "%zu\t%u\n", p + 1, count[b + p]);
        } else {
          //  depth of the found k-mers over every base: +1 where one starts, -1 k bases on, running sum (dump.C:219-238)
          std::vector<int32_t> diff(len + k + 1, 0);
          size_t maxP = 0;
          for (size_t p = 0; p < len; p++) if (found[0][b + p]) { diff[p]++; diff[p + k]--; maxP = p + k; }
          int32_t dsum = 0;
          for (size_t p = 0; p < maxP; p++) { dsum += diff[p]; if (dsum) fprintf(o1, "%zu\t%d\n", p + 1, dsum); }
        }
      }
    }
    (void)num;
  } else if (report == EXISTENCE) {
    const size_t n = S1.size();
    std::vector<std::vector<uint64_t>> found(dbs.size(), std::vector<uint64_t>(n, 0));
    std::vector<uint64_t> total(n, 0), dbKmers(dbs.size(), 0);
    for (size_t d = 0; d < dbs.size(); d++)
      if (askDatabase(dbs[d], minV, maxV, S1, 0, total.data(), found[d].data(), &dbKmers[d], nullptr, nullptr) != MB_OK) {
        fprintf(stderr, "ERROR: %s\n", mb_last_error());
        return 1;
      }
    for (size_t i = 0; i < n; i++) {                       //  existence.C:91-116
      fprintf(o1, "%s\t%llu", S1.ident[i].c_str(), (unsigned long long)total[i]);
      for (size_t d = 0; d < dbs.size(); d++) fprintf(o1, "\t%llu\t%llu", (unsigned long long)dbKmers[d], (unsigned long long)found[d][i]);
      fputc('\n', o1);
    }
  } else {
    const size_t n1 = S1.size(), n2 = S2.size(), n = n1 > n2 ? n1 : n2;
    std::vector<uint64_t> t1(n1, 0), f1(n1, 0), t2(n2, 0), f2(n2, 0);
    if (askDatabase(dbs[0], minV, maxV, S1, is10x ? 23 : 0, t1.data(), f1.data(), nullptr, nullptr, nullptr) != MB_OK ||
        (seq2 && askDatabase(dbs[0], minV, maxV, S2, 0, t2.data(), f2.data(), nullptr, nullptr, nullptr) != MB_OK)) {
      fprintf(stderr, "ERROR: %s\n", mb_last_error());
      return 1;
    }
    uint64_t kept = 0;
    for (size_t i = 0; i < n; i++) {                       //  include-exclude.C:114-134
      const uint64_t nf = (i < n1 ? f1[i] : 0) + (i < n2 ? f2[i] : 0);
      if ((nf > 0) != (report == INCLUDE)) continue;
      kept++;
      if (i < n1) writeRecord(o1, S1, i, nf);
      if (i < n2) writeRecord(o2, S2, i, nf);
    }
    fprintf(stderr, "\nIncluding %llu reads (or read pairs) out of %llu.\n", (unsigned long long)kept, (unsigned long long)n);
  }
  if (closeOut(o1, piped1) != 0 || (o2 && closeOut(o2, piped2) != 0)) { fprintf(stderr, "ERROR: cannot write the output.\n"); return 1; }
  return 0;
}
