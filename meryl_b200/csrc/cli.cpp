//  meryl-b200: the `meryl count` command line on a B200.
//
//  Keeps the count-relevant subset of the reference's word grammar
//  (src/meryl/merylCommandBuilder.C:188-326 options, :346-348 operations, :441-461 output,
//  :538-549 sequence inputs; usage text src/meryl/meryl.C:96-171):
//
//      meryl-b200 [-V|-Q|-P|-C] [k=K] [n=N] [memory=GB] [threads=T] [gpus=N]
//                 count|count-forward|count-reverse [compress] <seq files...> output <db.meryl>
//
//  Words are order-free like the reference's.  Errors are a message on stderr and exit(1)
//  (missing k: merylOp-count.C:311-312; missing output: :314-315).  The set operations and filters on databases
//  (csrc/union.cu), `histogram`, `statistics` and `print <db>` are here too.  Any other word (compare, lookup tools, ...)
//  is not part of this tool: if $MERYL_REFERENCE names the reference
//  binary the whole command line is handed to it unchanged, otherwise it is an error.

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <string>
#include <vector>

#include "../../include/meryl_b200.h"

static uint64_t physicalMemory(void) {          //  getMaxMemoryAllowed(), src/utility/src/system/system-v1.C:216-238
  uint64_t m = (uint64_t)sysconf(_SC_PHYS_PAGES) * (uint64_t)sysconf(_SC_PAGESIZE);
  const char *cpu = getenv("SLURM_JOB_CPUS_PER_NODE"), *env = getenv("SLURM_MEM_PER_CPU");
  if (env && cpu) { uint64_t v = strtoull(cpu, 0, 10) * strtoull(env, 0, 10) * 1024 * 1024; if (v < m) m = v; }
  if ((env = getenv("SLURM_MEM_PER_NODE"))) { uint64_t v = strtoull(env, 0, 10) * 1024 * 1024; if (v < m) m = v; }
  if ((env = getenv("PBS_RESC_MEM")))       { uint64_t v = strtoull(env, 0, 10); if (v < m) m = v; }
  return m ? m : UINT64_MAX;
}

static bool isCountWord(const char *w, uint32_t *mode) {
  if (!strcmp(w, "count"))         { *mode = MB_MODE_CANONICAL; return true; }
  if (!strcmp(w, "count-forward")) { *mode = MB_MODE_FORWARD;   return true; }
  if (!strcmp(w, "count-reverse")) { *mode = MB_MODE_REVERSE;   return true; }
  return false;
}

//  meryl-b200 [threads=T] [memory=GB] union|union-min|union-max|union-sum|intersect|intersect-min|intersect-max|intersect-sum|
//             subtract|difference|symmetric-difference <db.meryl> <db.meryl> ... output <db.meryl>
//  (merylCommandBuilder.C:333-435: the operation word; database inputs are directories holding a merylIndex)
static const char *kSetOps[] = { "union", "union-min", "union-max", "union-sum", "intersect", "intersect-min", "intersect-max",
                                 "intersect-sum", "subtract", "difference", "symmetric-difference",
                                 "less-than", "greater-than", "at-least", "at-most", "equal-to", "not-equal-to",
                                 "increase", "decrease", "multiply", "divide", "divide-round", "modulo" };
static int setOpOf(const char *w) {
  for (int i = 0; i < 23; i++) if (!strcmp(w, kSetOps[i])) return i;
  return -1;
}

static int unionSumMain(int argc, char **argv) {
  std::vector<const char *> dbs;
  int op = -1;
  bool haveParam = false;
  uint64_t param = 0;
  double fracDistinct = -1, wordFrequency = -1;
  const char *outDir = nullptr;
  bool quiet = false;
  for (int i = 1; i < argc; i++) {
    const char *w = argv[i];
    if (setOpOf(w) >= 0 && op < 0) { op = setOpOf(w); continue; }
    if (!strncmp(w, "threads=", 8) || !strncmp(w, "memory=", 7) || !strncmp(w, "k=", 2) || !strncmp(w, "-V", 2) || !strcmp(w, "-P") ||
        !strcmp(w, "[") || !strcmp(w, "]")) continue;
    if (!strcmp(w, "-Q")) { quiet = true; continue; }
    if (!strcmp(w, "output")) {
      if (i + 1 >= argc) { fprintf(stderr, "ERROR: 'output' needs a database name.\n"); return 1; }
      outDir = argv[++i];
      continue;
    }
    if (op >= MB_OP_LESS_THAN && !haveParam && w[0] >= '0' && w[0] <= '9' && w[strspn(w, "0123456789")] == 0) {   //  merylCommandBuilder.C:225-233
      param = strtoull(w, 0, 10);  haveParam = true;
      continue;
    }
    //  thresholds of the filters: threshold= / t=, distinct= / d=, word-frequency= / f=  (merylCommandBuilder.C:277-296)
    if (!strncmp(w, "threshold=", 10) || !strncmp(w, "t=", 2)) { param = strtoull(strchr(w, '=') + 1, 0, 10);  haveParam = true;  continue; }
    if (!strncmp(w, "distinct=", 9) || !strncmp(w, "d=", 2))   { fracDistinct = strtod(strchr(w, '=') + 1, 0);  continue; }
    if (!strncmp(w, "word-frequency=", 15) || !strncmp(w, "f=", 2)) { wordFrequency = strtod(strchr(w, '=') + 1, 0);  continue; }
    const std::string idx = std::string(w) + "/merylIndex";
    if (access(idx.c_str(), R_OK) != 0) { fprintf(stderr, "ERROR: '%s' is not a meryl database (no readable merylIndex).\n", w); return 1; }
    dbs.push_back(w);
  }
  if (dbs.empty()) { fprintf(stderr, "ERROR: %s needs at least one input database.\n", kSetOps[op]); return 1; }
  if (!outDir)     { fprintf(stderr, "ERROR: No output specified for %s operation.\n", kSetOps[op]); return 1; }
  mb_union_stats st;
  int rc;
  if (op >= MB_OP_LESS_THAN) {
    if (dbs.size() != 1) { fprintf(stderr, "ERROR: %s takes exactly one input database.\n", kSetOps[op]); return 1; }
    if ((fracDistinct >= 0 || wordFrequency >= 0) && op <= MB_OP_NOT_EQUAL_TO) {       //  merylOperation::initializeThreshold
      if (mb_db_threshold(dbs[0], fracDistinct, wordFrequency, &param) != MB_OK) { fprintf(stderr, "ERROR: %s\n", mb_last_error()); return 1; }
      haveParam = true;
    }
    if (!haveParam)      { fprintf(stderr, "ERROR: %s needs a number (e.g. '%s 5').\n", kSetOps[op], kSetOps[op]); return 1; }
    rc = mb_value_operation((uint32_t)op, param, dbs[0], outDir, -1, 0, &st);
  } else
    rc = mb_set_operation((uint32_t)op, dbs.data(), (uint32_t)dbs.size(), outDir, -1, 0, &st);
  if (rc != MB_OK) {
    fprintf(stderr, "ERROR: %s\n", mb_last_error());
    return 1;
  }
  if (!quiet) {
    fprintf(stderr, "%s of %u databases (k=%u): %llu k-mers in, %llu distinct out (%llu unique), %u group%s, %.1f ms\n",
            kSetOps[op], st.nInputs, st.k, (unsigned long long)st.nInputKmers, (unsigned long long)st.nOutputKmers, (unsigned long long)st.nUnique,
            st.nGroups, st.nGroups == 1 ? "" : "s", st.msTotal);
    fprintf(stderr, "  read+decode %.2f  merge %.2f  encode+write %.2f ms;  %llu bytes in, %llu bytes out\n",
            st.msDecode, st.msMerge, st.msEncode, (unsigned long long)st.inBytes, (unsigned long long)st.outBytes);
    fprintf(stderr, "\nBye.\n");
  }
  return 0;
}

//  gpus=N: one process per GPU (forked before anything touches CUDA), rank r = device r.  Every rank ingests shard r of every
//  input (mb_add_file_shard: the reads are split evenly, cut at record starts), the library routes the k-mers to the GPU that
//  owns their data file, every rank writes its own files into the output directory, rank 0 the master index.
//  The NCCL id travels through a page of shared memory.
struct SharedPage {
  volatile int idReady;                 //  1: id[] is valid, -1: rank 0 failed before it had one
  uint8_t      id[128];
  mb_stats     st[32];
  char         err[32][512];
};

static int multiGpuCount(uint32_t nGpus, mb_config cfg, const std::vector<const char *> &inputs, const std::vector<char> &inputCompress, const char *outDir, mb_stats *total) {
  SharedPage *sh = (SharedPage *)mmap(nullptr, sizeof(SharedPage), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (sh == MAP_FAILED) { perror("mmap"); return 1; }
  memset((void *)sh, 0, sizeof(SharedPage));
  std::vector<pid_t> pids(nGpus, -1);
  for (uint32_t r = 0; r < nGpus; r++) {
    pids[r] = fork();
    if (pids[r] < 0) { perror("fork"); return 1; }
    if (pids[r] == 0) {                                   //  rank r
      int rc = MB_OK;
      if (r == 0) {
        rc = mb_comm_unique_id(sh->id);
        __sync_synchronize();
        sh->idReady = (rc == MB_OK) ? 1 : -1;
      } else {
        while (sh->idReady == 0) sched_yield();
        if (sh->idReady < 0) _exit(2);
      }
      mb_ctx *c = nullptr;
      cfg.device = (int32_t)r;  cfg.rank = r;  cfg.worldSize = nGpus;
      if (rc == MB_OK) rc = mb_create(&cfg, &c);
      if (rc == MB_OK) rc = mb_comm_init(c, sh->id);
      if (rc == MB_OK) rc = mb_set_output_dir(c, outDir);
      for (size_t i = 0; rc == MB_OK && i < inputs.size(); i++) {
        rc = mb_set_input_compression(c, inputCompress[i]);
        if (rc == MB_OK) rc = mb_add_file_shard(c, inputs[i], r, nGpus);
      }
      if (rc == MB_OK) rc = mb_count(c);
      if (rc == MB_OK) rc = mb_write_db(c, outDir);
      if (rc == MB_OK) rc = mb_get_stats(c, &sh->st[r]);
      if (rc != MB_OK) { strncpy(sh->err[r], mb_last_error(), sizeof(sh->err[r]) - 1); _exit(1); }   //  (the peers will fail in their collectives)
      mb_destroy(c);
      _exit(0);
    }
  }
  int bad = 0;
  for (uint32_t r = 0; r < nGpus; r++) {
    int status = 0;
    if (waitpid(pids[r], &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) {
      if (!bad) fprintf(stderr, "ERROR: rank %u (GPU %u): %s\n", r, r, sh->err[r][0] ? sh->err[r] : "process ended abnormally");
      bad = 1;
      for (uint32_t q = 0; q < nGpus; q++) if (q != r && pids[q] > 0) kill(pids[q], SIGTERM);     //  do not leave the others waiting in a collective
    }
  }
  if (!bad) {
    *total = sh->st[0];
    for (uint32_t r = 1; r < nGpus; r++) {
      total->nKmers += sh->st[r].nKmers;  total->nDistinct += sh->st[r].nDistinct;  total->nUnique += sh->st[r].nUnique;
      total->nBases += sh->st[r].nBases;  total->dbBytes += sh->st[r].dbBytes;
      if (sh->st[r].msTotal > total->msTotal) total->msTotal = sh->st[r].msTotal;
    }
  }
  munmap((void *)sh, sizeof(SharedPage));
  return bad;
}

static int flatMain(int argc, char **argv);

//  Operation trees (merylCommandBuilder.C:441-563: '[' opens an operation whose result is an input of the operation around
//  it, ']' closes it).  The reference streams a sub-operation's k-mers into its parent; here every operation is one call on
//  the device, so a sub-operation is run first, into its own `output` if it names one, else into a temporary database next
//  to the root's output, and the parent reads that directory.  The result is the same function of the inputs.  k= is global
//  as in the reference (kmerTiny::setSize); the other options of the outermost level are handed down to every operation.
struct TreeNode {
  std::vector<std::string> words;       //  this level's words in order; "\x01<i>" stands for sub-operation i
  std::vector<TreeNode>    subs;
};

static bool parseTree(char **argv, int argc, int &i, TreeNode &node, bool inBracket) {
  for (; i < argc; i++) {
    if (!strcmp(argv[i], "[")) {
      i++;
      node.subs.emplace_back();
      node.words.push_back(std::string("\x01") + std::to_string(node.subs.size() - 1));
      if (!parseTree(argv, argc, i, node.subs.back(), true)) return false;
    } else if (!strcmp(argv[i], "]")) {
      if (!inBracket) { fprintf(stderr, "ERROR: ']' without a '['.\n"); return false; }
      return true;
    } else node.words.push_back(argv[i]);
  }
  if (inBracket) { fprintf(stderr, "ERROR: '[' without a ']'.\n"); return false; }
  return true;
}

static bool isOptionWord(const std::string &w) {
  static const char *opts[] = { "k=", "memory=", "threads=", "gpus=", "-V", "-Q", "-P" };
  for (const char *o : opts) if (!strncmp(w.c_str(), o, strlen(o))) return true;
  return false;
}

//  Runs the operation of `node` (after its sub-operations) and returns its database in `outDb`.
static int runTree(const char *prog, TreeNode &node, const std::vector<std::string> &inherited, const std::string &tmpBase, int &tmpSeq,
                   bool isRoot, std::string &outDb, std::vector<std::string> &temps) {
  std::vector<std::string> subDb(node.subs.size());
  for (size_t j = 0; j < node.subs.size(); j++)
    if (runTree(prog, node.subs[j], inherited, tmpBase, tmpSeq, false, subDb[j], temps) != 0) return 1;
  std::vector<std::string> cmd(1, prog);
  for (const std::string &w : inherited) {                 //  an option of its own overrides the inherited one
    const std::string key = w.substr(0, w.find('=') == std::string::npos ? w.size() : w.find('=') + 1);
    bool own = false;
    for (const std::string &x : node.words) if (!strncmp(x.c_str(), key.c_str(), key.size())) own = true;
    if (!own) cmd.push_back(w);
  }
  bool haveOutput = false;
  for (size_t j = 0; j < node.words.size(); j++) {
    const std::string &w = node.words[j];
    if (w[0] == '\x01') { cmd.push_back(subDb[(size_t)atoi(w.c_str() + 1)]); continue; }
    if (w == "output" && j + 1 < node.words.size()) { haveOutput = true; outDb = node.words[j + 1]; }
    cmd.push_back(w);
  }
  bool report = false;                                     //  print: the result is text, not a database
  for (const std::string &w : node.words) if (w == "print") report = true;
  if (!haveOutput && !(isRoot && report)) {
    if (isRoot) { fprintf(stderr, "ERROR: No output specified for the outermost operation.\n"); return 1; }
    outDb = tmpBase + ".tmp-" + std::to_string(tmpSeq++);
    temps.push_back(outDb);
    cmd.push_back("output");  cmd.push_back(outDb);
  }
  if (getenv("MB_CLI_DRY")) {                              //  testing: print the flat command lines instead of running them
    for (size_t j = 1; j < cmd.size(); j++) printf("%s%s", j > 1 ? " " : "", cmd[j].c_str());
    printf("\n");
    return 0;
  }
  std::vector<char *> av;
  for (std::string &c : cmd) av.push_back(&c[0]);
  av.push_back(nullptr);
  return flatMain((int)cmd.size(), av.data());
}

static void removeTempDb(const std::string &dir) {
  char name[64];
  for (unsigned f = 0; f < 64; f++)
    for (int isIdx = 0; isIdx < 2; isIdx++) {
      char bits[7];
      for (int b = 0; b < 6; b++) bits[b] = ((f >> (5 - b)) & 1) ? '1' : '0';
      bits[6] = 0;
      snprintf(name, sizeof(name), "/0x%s.%s", bits, isIdx ? "merylIndex" : "merylData");      //  constructBlockName, kmers.C:31-56
      unlink((dir + name).c_str());
    }
  unlink((dir + "/merylIndex").c_str());
  rmdir(dir.c_str());
}

int mbLookupMain(int argc, char **argv);                 //  lookup_cli.cpp: meryl-lookup's -existence / -include / -exclude

int main(int argc, char **argv) {
  const char *self = strrchr(argv[0], '/');
  self = self ? self + 1 : argv[0];
  if (!strncmp(self, "meryl-lookup", 12) || (argc > 1 && !strcmp(argv[1], "lookup"))) return mbLookupMain(argc, argv);
  bool tree = false;
  for (int i = 1; i < argc; i++) if (!strcmp(argv[i], "[")) tree = true;
  if (!tree) return flatMain(argc, argv);

  TreeNode root;
  int i = 1;
  if (!parseTree(argv, argc, i, root, false)) return 1;
  //  `meryl [ op ... ]`: the outermost level holds nothing but options and one bracket -- that bracket is the root
  std::vector<std::string> inherited;
  for (const std::string &w : root.words) if (isOptionWord(w)) inherited.push_back(w);
  TreeNode *top = &root;
  bool onlyOptions = root.subs.size() == 1;
  for (const std::string &w : root.words) if (!isOptionWord(w) && w[0] != '\x01') onlyOptions = false;
  if (onlyOptions) top = &root.subs[0];
  else {                                                   //  the outermost level is itself an operation: its options stay with it
    std::vector<std::string> keep;
    for (const std::string &w : root.words) if (!isOptionWord(w)) keep.push_back(w);
    root.words.swap(keep);
  }
  //  k= is global wherever it is written
  std::vector<TreeNode *> stack(1, &root);
  while (!stack.empty()) {
    TreeNode *n = stack.back();  stack.pop_back();
    for (const std::string &w : n->words)
      if (!strncmp(w.c_str(), "k=", 2)) { bool have = false; for (const std::string &x : inherited) if (!strncmp(x.c_str(), "k=", 2)) have = true; if (!have) inherited.push_back(w); }
    for (TreeNode &c : n->subs) stack.push_back(&c);
  }
  std::string rootOut;
  for (size_t j = 0; j + 1 < top->words.size(); j++) if (top->words[j] == "output") rootOut = top->words[j + 1];
  bool rootReports = false;                                //  `print [ ... ]`: the sub-operation runs into a temporary database and is printed
  for (const std::string &w : top->words)                  //  from that (`histogram` / `statistics` of a tree: the reference refuses, and so does this)
    if (w == "print") rootReports = true;
  if (rootOut.empty() && rootReports) {
    const char *td = getenv("TMPDIR");
    rootOut = std::string(td && td[0] ? td : "/tmp") + "/meryl-b200-" + std::to_string((long)getpid());
  }
  if (rootOut.empty()) { fprintf(stderr, "ERROR: No output specified for the outermost operation.\n"); return 1; }
  int tmpSeq = 0;
  std::vector<std::string> temps;
  std::string outDb;
  const int rc = runTree(argv[0], *top, inherited, rootOut, tmpSeq, true, outDb, temps);
  for (const std::string &t : temps) removeTempDb(t);
  return rc;
}

//  `meryl histogram <db>` and `meryl statistics <db>` (merylOp-histogram.C:20-100): printed from the master index.
static int reportMain(int argc, char **argv, bool statistics) {
  const char *db = nullptr;
  for (int i = 1; i < argc; i++) {
    const std::string idx = std::string(argv[i]) + "/merylIndex";
    if (access(idx.c_str(), R_OK) == 0) {
      if (db) { fprintf(stderr, "ERROR: told to dump a histogram for more than one input!\n"); return 1; }
      db = argv[i];
    }
  }
  if (!db) { fprintf(stderr, "ERROR: no meryl database given to %s.\n", statistics ? "statistics" : "histogram"); return 1; }
  uint32_t k = 0;
  uint64_t nUnique = 0, nDistinct = 0, nTotal = 0, n = 0;
  if (mb_db_histogram(db, &k, &nUnique, &nDistinct, &nTotal, nullptr, nullptr, 0, &n) != MB_OK) { fprintf(stderr, "ERROR: %s\n", mb_last_error()); return 1; }
  std::vector<uint64_t> val(n + 1), occ(n + 1);
  if (mb_db_histogram(db, nullptr, nullptr, nullptr, nullptr, val.data(), occ.data(), n, &n) != MB_OK) { fprintf(stderr, "ERROR: %s\n", mb_last_error()); return 1; }
  if (!statistics) {
    for (uint64_t i = 0; i < n; i++) fprintf(stdout, "%llu\t%llu\n", (unsigned long long)val[i], (unsigned long long)occ[i]);
    return 0;
  }
  const uint64_t nUniverse = (2 * k >= 64) ? 0 : ((uint64_t)1 << (2 * k));        //  buildLowBitMask<uint64>(2k) + 1, wraps for k >= 32
  fprintf(stdout, "Number of %u-mers that are:\n", k);
  fprintf(stdout, "  unique   %20llu  (exactly one instance of the kmer is in the input)\n", (unsigned long long)nUnique);
  fprintf(stdout, "  distinct %20llu  (non-redundant kmer sequences in the input)\n", (unsigned long long)nDistinct);
  fprintf(stdout, "  present  %20llu  (...)\n", (unsigned long long)nTotal);
  fprintf(stdout, "  missing  %20llu  (non-redundant kmer sequences not in the input)\n", (unsigned long long)(nUniverse - nDistinct));
  fprintf(stdout, "\n");
  fprintf(stdout, "             number of   cumulative   cumulative     presence\n");
  fprintf(stdout, "              distinct     fraction     fraction   in dataset\n");
  fprintf(stdout, "frequency        kmers     distinct        total       (1e-6)\n");
  fprintf(stdout, "--------- ------------ ------------ ------------ ------------\n");
  uint64_t sDistinct = 0, sTotal = 0;
  for (uint64_t i = 0; i < n; i++) {
    sDistinct += occ[i];
    sTotal    += occ[i] * val[i];
    fprintf(stdout, "%9llu %12llu %12.4f %12.4f %12.6f\n", (unsigned long long)val[i], (unsigned long long)occ[i],
            (double)sDistinct / nDistinct, (double)sTotal / nTotal, (double)val[i] / nTotal * 1000000.0);
  }
  return 0;
}

//  `meryl print <db>` (merylOp-nextMer.C:646-667: "<k-mer>\t<value>" per line): the blocks are decoded on the device
//  (mb_for_each_kmer), the text is made here.  Ascending k-mer order (the reference's order with threads=1).
static int printPiece(void *arg, uint32_t k, const uint64_t *lo, const uint64_t *hi, const uint32_t *val, uint64_t n) {
  FILE *out = (FILE *)arg;
  static const char letter[4] = { 'A', 'C', 'T', 'G' };    //  kmers-tiny.H:86
  char line[128];
  for (uint64_t i = 0; i < n; i++) {
    for (uint32_t b = 0; b < k; b++) {
      const uint32_t sh = 2 * (k - 1 - b);
      const uint64_t w = sh >= 64 ? hi[i] >> (sh - 64) : lo[i] >> sh;
      line[b] = letter[w & 3];
    }
    const int m = snprintf(line + k, sizeof(line) - k, "\t%u\n", val[i]);
    if (fwrite(line, 1, k + (size_t)m, out) != k + (size_t)m) return 1;
  }
  return 0;
}

static int printMain(int argc, char **argv) {
  const char *db = nullptr;
  for (int i = 1; i < argc; i++) {
    const std::string idx = std::string(argv[i]) + "/merylIndex";
    if (access(idx.c_str(), R_OK) == 0) {
      if (db) { fprintf(stderr, "ERROR: print of more than one database: combine them with an operation first.\n"); return 1; }
      db = argv[i];
    }
  }
  if (!db) { fprintf(stderr, "ERROR: no meryl database given to print.\n"); return 1; }
  if (mb_for_each_kmer(db, printPiece, stdout, -1, 0) != MB_OK) { fprintf(stderr, "ERROR: %s\n", mb_last_error()); return 1; }
  return fflush(stdout) == 0 ? 0 : 1;
}

static int flatMain(int argc, char **argv) {
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "histogram"))  return reportMain(argc, argv, false);
    if (!strcmp(argv[i], "statistics")) return reportMain(argc, argv, true);
    if (!strcmp(argv[i], "print"))      return printMain(argc, argv);
  }
  for (int i = 1; i < argc; i++)
    if (setOpOf(argv[i]) >= 0) return unionSumMain(argc, argv);

  uint32_t k = 0, mode = 0, nGpus = 1;
  uint64_t nExpected = 0, memBytes = physicalMemory();
  bool haveOp = false, onlyConfig = false, quiet = false, compress = false;
  const char *outDir = nullptr, *countSuffix = nullptr;
  std::string countSuffixWord;
  std::vector<const char *> inputs;
  std::vector<char> inputCompress;                     //  per input: was `compress` in force when it was named?

  for (int i = 1; i < argc; i++) {
    const char *w = argv[i];
    if      (!strncmp(w, "-V", 2))            { }
    else if (!strcmp(w, "-Q"))                { quiet = true; }
    else if (!strcmp(w, "-P"))                { }
    else if (!strcmp(w, "-C"))                { onlyConfig = true; }
    else if (!strncmp(w, "k=", 2))            { k = (uint32_t)strtoul(w + 2, 0, 10); }
    else if (!strncmp(w, "n=", 2))            { nExpected = strtoull(w + 2, 0, 10); }
    else if (!strncmp(w, "memory=", 7))       { memBytes = (uint64_t)(strtod(w + 7, 0) * 1024 * 1024 * 1024); }
    else if (!strncmp(w, "threads=", 8))      { }                         //  no CPU worker threads to size
    else if (!strncmp(w, "gpus=", 5))         { nGpus = (uint32_t)strtoul(w + 5, 0, 10); }     //  (not a word of the reference)
    else if (isCountWord(w, &mode))           { haveOp = true; }
    else if (!strcmp(w, "compress"))          { compress = true; }           //  merylCommandBuilder.C:237-240: for the inputs that follow
    else if (!strncmp(w, "count-suffix=", 13)) {                               //  merylCommandBuilder.C:271-274
      //  The reference builds the suffix with kmerTiny::addR at this point of the command line (merylOp.H:139-147); in front
      //  of `k=` the k-mer masks are still zero and every addR drops what came before: only the last base survives, the test
      //  becomes "ends with A...A<last base>".  Reproduced (the databases are compared byte for byte), and said aloud.
      countSuffixWord = w + 13;
      if (k == 0 && countSuffixWord.size() > 1) {
        countSuffixWord = std::string(countSuffixWord.size() - 1, 'A') + countSuffixWord.back();
        fprintf(stderr, "NOTE: count-suffix= in front of k=: like the reference, only its last base counts -- keeping the k-mers that end with '%s'.\n",
                countSuffixWord.c_str());
      }
      countSuffix = countSuffixWord.c_str();
    }
    else if (!strcmp(w, "output")) {
      if (i + 1 >= argc) { fprintf(stderr, "ERROR: 'output' needs a database name.\n"); return 1; }
      outDir = argv[++i];
    }
    else if (!strcmp(w, "[") || !strcmp(w, "]")) { }
    else if (access(w, R_OK) == 0 || !strcmp(w, "-")) { inputs.push_back(w); inputCompress.push_back(compress ? 1 : 0); }
    else {
      //  not a count command line: the reference handles everything else
      const char *ref = getenv("MERYL_REFERENCE");
      if (ref) { execv(ref, argv); perror("execv"); return 1; }
      fprintf(stderr, "ERROR: word '%s' is not part of `meryl count` (operation, option or readable input file).\n", w);
      fprintf(stderr, "       meryl-b200 implements count, count-forward and count-reverse; set MERYL_REFERENCE=<path to meryl>\n");
      fprintf(stderr, "       to have every other command line passed to the reference binary.\n");
      return 1;
    }
  }

  if (!haveOp) {
    const char *ref = getenv("MERYL_REFERENCE");
    if (ref) { execv(ref, argv); perror("execv"); return 1; }
    fprintf(stderr, "usage: %s [k=K] [n=N] [memory=GB] count|count-forward|count-reverse <sequence files> output <db.meryl>\n", argv[0]);
    return 1;
  }
  if (k == 0)                    { fprintf(stderr, "ERROR: Kmer size not supplied with modifier k=<kmer-size>.\n"); return 1; }
  if (!outDir && !onlyConfig)    { fprintf(stderr, "ERROR: No output specified for count operation.\n"); return 1; }
  if (inputs.empty())            { fprintf(stderr, "ERROR: No sequence inputs for count operation.\n"); return 1; }
  if (nGpus == 0 || nGpus > 32 || 64 % nGpus) { fprintf(stderr, "ERROR: gpus=%u: the number of GPUs must divide 64 (1, 2, 4, 8, 16, 32).\n", nGpus); return 1; }

  if (nExpected == 0)
    for (const char *p : inputs) nExpected += mb_estimate_kmers_in_file(p);

  uint32_t wPrefix = 0, nBatches = 0;  int simple = 0;  uint64_t memUsed = 0;
  if (mb_plan(k, nExpected, memBytes, &wPrefix, &simple, &memUsed, &nBatches) != MB_OK) {
    fprintf(stderr, "ERROR: %s\n", mb_last_error());
    return 1;
  }
  if (!quiet) {
    fprintf(stderr, "\nCounting %llu (estimated) %s %u-mers from %zu input file%s.\n", (unsigned long long)nExpected,
            mode == MB_MODE_CANONICAL ? "canonical" : mode == MB_MODE_FORWARD ? "forward" : "reverse",
            k, inputs.size(), inputs.size() == 1 ? "" : "s");
    //  This line is parsed by Canu (merylOp-count.C:398-401); keep its shape.
    fprintf(stderr, "Configured %s mode for %.3f GB memory per batch, and up to %u batch%s.\n",
            simple ? "simple" : "complex", (memUsed < memBytes ? memUsed : memBytes) / 1024.0 / 1024.0 / 1024.0,
            nBatches, nBatches == 1 ? "" : "es");
    fprintf(stderr, "Database prefix width %u bits (%u file bits + %u block bits), suffix %u bits.\n\n", wPrefix, 6, wPrefix - 6, 2 * k - wPrefix);
  }
  if (onlyConfig) return 0;

  mb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.k = k;  cfg.mode = mode;  cfg.wPrefix = wPrefix;  cfg.device = -1;
  //  (`compress` holds for the inputs named after it: every input carries its own flag)

  //  count-suffix=: the full count goes to a temporary database, mb_keep_suffix writes the wanted k-mers (csrc/union.cu)
  const std::string finalDir = outDir;
  const std::string countDir = countSuffix ? finalDir + ".tmp-count" : finalDir;
  mb_stats st;
  if (nGpus > 1) {
    if (multiGpuCount(nGpus, cfg, inputs, inputCompress, countDir.c_str(), &st) != 0) return 1;
  } else {
    //  what mb_count_files does, with the per-input compression flag
    mb_ctx *c = nullptr;
    int rc = mb_create(&cfg, &c);
    if (rc == MB_OK) rc = mb_set_output_dir(c, countDir.c_str());
    for (size_t i = 0; rc == MB_OK && i < inputs.size(); i++) {
      rc = mb_set_input_compression(c, inputCompress[i]);
      if (rc == MB_OK) rc = mb_add_file(c, inputs[i]);
    }
    if (rc == MB_OK) rc = mb_count(c);
    if (rc == MB_OK) rc = mb_write_db(c, countDir.c_str());
    if (rc == MB_OK) rc = mb_get_stats(c, &st);
    if (rc != MB_OK) { fprintf(stderr, "ERROR: %s\n", mb_last_error()); if (c) mb_destroy(c); return 1; }
    mb_destroy(c);
  }
  if (countSuffix) {
    mb_union_stats us;
    const int rc = mb_keep_suffix(countDir.c_str(), countSuffix, finalDir.c_str(), -1, 0, &us);
    if (rc != MB_OK) fprintf(stderr, "ERROR: %s\n", mb_last_error());
    removeTempDb(countDir);
    if (rc != MB_OK) return 1;
    if (!quiet) fprintf(stderr, "Kept the %llu distinct %u-mers that end with '%s'.\n", (unsigned long long)us.nOutputKmers, k, countSuffix);
  }
  if (!quiet) {
    fprintf(stderr, "Counted %llu %u-mers (%llu distinct, %llu unique) in %u pass%s, %.1f ms on the device.\n",
            (unsigned long long)st.nKmers, k, (unsigned long long)st.nDistinct, (unsigned long long)st.nUnique,
            st.nPasses, st.nPasses == 1 ? "" : "es", st.msTotal);
    fprintf(stderr, "  histogram %.2f  extract %.2f  sort %.2f+%.2f  count %.2f  encode %.2f ms;  db %llu bytes\n",
            st.msHistogram, st.msExtract, st.msSortHist, st.msSort, st.msCount, st.msEncode, (unsigned long long)st.dbBytes);
    fprintf(stderr, "\nFinished counting.\n");
  }
  return 0;
}
