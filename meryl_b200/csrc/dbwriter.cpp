//  W5/W6: the host-side tail of the db writer.  The 64 data files and their block indexes are
//  produced on the device (encode.cu); what is left for the host is
//    - the master `merylIndex`  (~merylFileWriter  src/utility/src/kmers-v1/kmers-writer.C:118-152,
//                                merylHistogram::dump src/utility/src/kmers-v1/kmers-histogram.C:63-95)
//      a single stuffedBits stream of 32768-bit sub-blocks, serialised as in
//      stuffedBits::dump (src/utility/src/bits/stuffedBits-v1.C:65-111);
//    - file naming              (constructBlockName src/utility/src/kmers-v1/kmers.C:31-56)
//    - writing the images to a directory.

#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include "common.h"

namespace {

//  Minimal MSB-first bit stream with the reference's sub-block rule: a field never straddles
//  two sub-blocks; if it does not fit, the current sub-block is closed where it is.
class FieldStream {
public:
  explicit FieldStream(u64 maxBits) : _max((maxBits + 63) & ~(u64)63) { open(); }

  void put(u32 width, u64 value) {
    if (width == 0) return;
    if (_pos + width > _max) { _len.back() = _pos; open(); }
    std::vector<u64> &w = _words.back();
    if (width < 64) value &= ((u64)1 << width) - 1;
    const u64 idx  = _pos >> 6;
    const u32 room = 64 - (u32)(_pos & 63);
    if (w.size() < idx + 2) w.resize(idx + 2, 0);
    if (width <= room)  w[idx] |= value << (room - width);
    else              { w[idx] |= value >> (width - room);  w[idx + 1] |= value << (64 - (width - room)); }
    _pos += width;
    _len.back() = _pos;
  }

  void serialise(std::vector<uint8_t> &out) const {
    const u32 nSub   = (u32)_len.size();
    const u32 subMax = 32 * ((nSub - 1) / 32 + 1);          //  growth rule of the sub-block table
    auto push = [&](const void *p, size_t n) { const uint8_t *b = (const uint8_t *)p; out.insert(out.end(), b, b + n); };
    push(&_max, 8);
    push(&nSub, 4);
    push(&subMax, 4);
    u64 bgn = 0;
    for (u32 i = 0; i < nSub; i++) { push(&bgn, 8); bgn += _len[i]; }
    for (u32 i = 0; i < nSub; i++)   push(&_len[i], 8);
    for (u32 i = 0; i < nSub; i++)   push(_words[i].data(), 8 * ((_len[i] + 63) / 64));
  }

private:
  void open(void) { _words.emplace_back(); _len.push_back(0); _pos = 0; }

  u64                            _max;
  u64                            _pos = 0;
  std::vector<u64>               _len;
  std::vector<std::vector<u64>>  _words;
};

}  //  namespace


void mb_build_master_index(u32 k, u32 wPrefix, const MbHistogram &h, std::vector<uint8_t> &out) {
  FieldStream s(32 * 1024);
  s.put(64, 0x646e496c7972656dull);            //  "merylInd"
  s.put(64, 0x33302e765f5f7865ull);            //  "ex__v.03"
  s.put(32, wPrefix);
  s.put(32, 2 * k - wPrefix);
  s.put(32, 6);
  s.put(32, wPrefix - 6);
  s.put(32, 0);                                //  flags: not a multiset
  s.put(64, h.nUnique);
  s.put(64, h.nDistinct);
  s.put(64, h.nTotal);
  s.put(64, h.values.size());
  for (size_t i = 0; i < h.values.size(); i++) {
    s.put(64, h.values[i]);
    s.put(64, h.occ[i]);
  }
  out.clear();
  s.serialise(out);
}


void mb_file_name(char *out, size_t cap, const char *dir, u32 file, bool isIndex) {
  char bits[7];
  for (int i = 0; i < 6; i++) bits[i] = ((file >> (5 - i)) & 1) ? '1' : '0';
  bits[6] = 0;
  snprintf(out, cap, "%s/0x%s.%s", dir, bits, isIndex ? "merylIndex" : "merylData");
}


static int writeWhole(const char *path, const uint8_t *p, size_t n) {
  FILE *f = fopen(path, "wb");
  if (!f) return mb_fail(MB_EIO, "cannot create '%s': %s", path, strerror(errno));
  if (n && fwrite(p, 1, n, f) != n) {
    fclose(f);
    return mb_fail(MB_EIO, "short write to '%s': %s", path, strerror(errno));
  }
  if (fclose(f) != 0) return mb_fail(MB_EIO, "cannot close '%s': %s", path, strerror(errno));
  return MB_OK;
}


int mb_write_db_dir(const MbDbImages &db, const char *dir, bool writeMaster) {
  if (mkdir(dir, 0777) != 0 && errno != EEXIST)
    return mb_fail(MB_EIO, "cannot create directory '%s': %s", dir, strerror(errno));
  char path[4096];
  for (u32 f = 0; f < 64; f++) {
    if (!db.owned[f]) continue;
    mb_file_name(path, sizeof(path), dir, f, false);
    if (!db.streamed[f]) MB_TRY(writeWhole(path, db.dataPtr[f], db.dataLen[f]));
    mb_file_name(path, sizeof(path), dir, f, true);
    MB_TRY(writeWhole(path, db.index[f].data(), db.index[f].size()));
  }
  if (writeMaster) {
    snprintf(path, sizeof(path), "%s/merylIndex", dir);
    MB_TRY(writeWhole(path, db.master.data(), db.master.size()));
  }
  return MB_OK;
}
