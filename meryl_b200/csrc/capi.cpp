//  C-ABI glue that needs no CUDA: error string, one-shot mb_count_files.
#include <stdio.h>
#include <string.h>

#include "common.h"

static thread_local char g_err[1024];

int mb_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void mb_clear_error(void) { g_err[0] = 0; }

extern "C" const char *mb_last_error(void) { return g_err; }

extern "C" int mb_count_files(const mb_config *cfg, const char **paths, uint32_t nPaths, const char *outDir, mb_stats *st) {
  mb_clear_error();
  if (!cfg || (!paths && nPaths) || !outDir) return mb_fail(MB_EINVAL, "mb_count_files: NULL argument");
  mb_ctx *c = nullptr;
  int rc = mb_create(cfg, &c);
  if (rc != MB_OK) return rc;
  if (rc == MB_OK) rc = mb_set_output_dir(c, outDir);       //  the data files are written while the count runs
  for (uint32_t i = 0; rc == MB_OK && i < nPaths; i++) rc = mb_add_file(c, paths[i]);
  if (rc == MB_OK) rc = mb_count(c);
  if (rc == MB_OK) rc = mb_write_db(c, outDir);
  if (rc == MB_OK && st) rc = mb_get_stats(c, st);
  char keep[1024];
  strncpy(keep, mb_last_error(), sizeof(keep) - 1);  keep[sizeof(keep) - 1] = 0;
  mb_destroy(c);
  if (rc != MB_OK) mb_fail(rc, "%s", keep);
  return rc;
}
