/*  TEST / BENCH INFRASTRUCTURE (oracle/) -- not part of the shipped product.
 *
 *  The bench's seeded synthetic reads (reads of a virtual uniform-random genome, both
 *  strands, substitution errors) written by the CPU, so that the reference arm of bench.py and the
 *  reference-binary parity tests can make their input WITHOUT loading the CUDA library.  Same recipe,
 *  same bytes as the device generator (meryl_b200/csrc/synth.cu; pinned by tests/test_gpu_count.py::
 *  test_cpu_generator_equals_device_generator):
 *     base g of the genome = mix64(seed ^ g * C1) & 3 in ACGT order; read r starts at mix64(h) % (G - L + 1) and is
 *     reverse-complemented iff h >> 63, h = mix64((seed + C2) ^ r * C3); position j of read r is substituted iff
 *     (u32) e < err * 2^32, e = mix64((seed + C4) ^ (r L + j) * C5), by base (c + 1 + (e >> 32) % 3) & 3.
 *  format 0: "<bases>."   format 1: "@r%012llu\n<bases>\n+\n<I x L>\n"
 *
 *  usage: synth_reads <out file> <first read> <n reads> <read len> <genome len> <err> <seed> <format>
 *  Built by oracle/Makefile: gcc -O3 -fopenmp.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

static void one_read(unsigned char *o, uint64_t r, uint32_t L, uint64_t G, uint32_t thr, uint64_t seed, int format) {
  const uint64_t h     = mix64((seed + 0x7f4a7c15) ^ (r * 0xd6e8feb86659fd93ull));
  const uint64_t start = mix64(h) % (G - L + 1);
  const int      rev   = (h >> 63) != 0;
  if (format == 1) {
    char name[32];
    snprintf(name, sizeof(name), "@r%012llu\n", (unsigned long long)r);
    memcpy(o, name, 15);
    o += 15;
  }
  for (uint32_t j = 0; j < L; j++) {
    const uint64_t g = rev ? start + (L - 1 - j) : start + j;
    uint32_t c = (uint32_t)mix64(seed ^ (g * 0x2545f4914f6cdd1dull)) & 3;
    if (rev) c = 3 - c;
    const uint64_t e = mix64((seed + 0x51ed27) ^ ((r * L + j) * 0x9e3779b97f4a7c15ull));
    if ((uint32_t)e < thr) c = (c + 1 + (uint32_t)((e >> 32) % 3)) & 3;
    o[j] = "ACGT"[c];
  }
  o += L;
  if (format == 0) { *o = '.'; return; }
  *o++ = '\n';  *o++ = '+';  *o++ = '\n';
  memset(o, 'I', L);
  o[L] = '\n';
}

int main(int argc, char **argv) {
  if (argc != 9) { fprintf(stderr, "usage: %s out first n readLen genomeLen err seed format\n", argv[0]); return 2; }
  const uint64_t first = strtoull(argv[2], 0, 10), n = strtoull(argv[3], 0, 10);
  const uint32_t L = (uint32_t)strtoul(argv[4], 0, 10);
  const uint64_t G = strtoull(argv[5], 0, 10);
  const double   err = atof(argv[6]);
  const uint64_t seed = strtoull(argv[7], 0, 10);
  const int      format = atoi(argv[8]);
  const uint32_t thr = err <= 0 ? 0 : err >= 1 ? 0xffffffffu : (uint32_t)(err * 4294967296.0);
  const uint64_t rec = format == 0 ? (uint64_t)L + 1 : 2 * (uint64_t)L + 19;
  if (L == 0 || G < L || format < 0 || format > 1) { fprintf(stderr, "bad arguments\n"); return 2; }
  FILE *fh = fopen(argv[1], "wb");
  if (!fh) { perror(argv[1]); return 1; }
  uint64_t chunk = ((uint64_t)256 << 20) / rec;                   /*  reads per 256 MB piece  */
  if (chunk < 1) chunk = 1;
  unsigned char *buf = (unsigned char *)malloc(chunk * rec);
  if (!buf) { fprintf(stderr, "out of memory\n"); return 1; }
  for (uint64_t r0 = 0; r0 < n; r0 += chunk) {
    const uint64_t m = n - r0 < chunk ? n - r0 : chunk;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < m; i++) one_read(buf + i * rec, first + r0 + i, L, G, thr, seed, format);
    if (fwrite(buf, rec, m, fh) != m) { perror("write"); return 1; }
  }
  free(buf);
  if (fclose(fh) != 0) { perror("close"); return 1; }
  return 0;
}
