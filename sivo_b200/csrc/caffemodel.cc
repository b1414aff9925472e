#include "caffemodel.h"

#include <cstring>
#include <fstream>

#include "common.h"

namespace sivo {
namespace {

struct Span {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
};

uint64_t varint(Span& s) {
  uint64_t v = 0;
  int shift = 0;
  while (true) {
    if (s.p >= s.end || shift > 63) fail(SIVO_EFORMAT, "caffemodel: truncated varint");
    uint8_t b = *s.p++;
    v |= static_cast<uint64_t>(b & 0x7F) << shift;
    if (!(b & 0x80)) return v;
    shift += 7;
  }
}

struct Field {
  uint32_t num;
  uint32_t wire;
  uint64_t scalar;  // wire 0 / 1 / 5
  Span bytes;       // wire 2
};

bool next_field(Span& s, Field& f) {
  if (s.done()) return false;
  uint64_t key = varint(s);
  f.num = static_cast<uint32_t>(key >> 3);
  f.wire = static_cast<uint32_t>(key & 7);
  switch (f.wire) {
    case 0: f.scalar = varint(s); break;
    case 1:
      if (s.end - s.p < 8) fail(SIVO_EFORMAT, "caffemodel: truncated fixed64");
      memcpy(&f.scalar, s.p, 8);
      s.p += 8;
      break;
    case 5: {
      if (s.end - s.p < 4) fail(SIVO_EFORMAT, "caffemodel: truncated fixed32");
      uint32_t v;
      memcpy(&v, s.p, 4);
      f.scalar = v;
      s.p += 4;
      break;
    }
    case 2: {
      uint64_t n = varint(s);
      if (static_cast<uint64_t>(s.end - s.p) < n) fail(SIVO_EFORMAT, "caffemodel: truncated length-delimited field");
      f.bytes = {s.p, s.p + n};
      s.p += n;
      break;
    }
    default: fail(SIVO_EFORMAT, "caffemodel: unsupported wire type %u", f.wire);
  }
  return true;
}

Blob parse_blob(Span s) {
  Blob b;
  int64_t legacy[4] = {1, 1, 1, 1};
  bool has_shape = false, has_legacy = false;
  Field f;
  while (next_field(s, f)) {
    if (f.num == 7 && f.wire == 2) {  // BlobShape
      has_shape = true;
      Span t = f.bytes;
      Field g;
      while (next_field(t, g)) {
        if (g.num == 1 && g.wire == 2) {
          Span u = g.bytes;
          while (!u.done()) b.shape.push_back(static_cast<int64_t>(varint(u)));
        } else if (g.num == 1 && g.wire == 0) {
          b.shape.push_back(static_cast<int64_t>(g.scalar));
        }
      }
    } else if (f.num == 5 && f.wire == 2) {  // packed float data
      size_t n = static_cast<size_t>(f.bytes.end - f.bytes.p) / 4;
      size_t old = b.data.size();
      b.data.resize(old + n);
      memcpy(b.data.data() + old, f.bytes.p, n * 4);
    } else if (f.num == 5 && f.wire == 5) {  // unpacked float
      uint32_t v = static_cast<uint32_t>(f.scalar);
      float x;
      memcpy(&x, &v, 4);
      b.data.push_back(x);
    } else if (f.num >= 1 && f.num <= 4 && f.wire == 0) {
      legacy[f.num - 1] = static_cast<int64_t>(f.scalar);
      has_legacy = true;
    } else if (f.num == 8) {
      fail(SIVO_EFORMAT, "caffemodel: double_data blobs are not supported");
    }
  }
  if (!has_shape) {
    if (!has_legacy && b.data.empty()) return b;
    b.shape.assign(legacy, legacy + 4);
  }
  if (b.count() != b.data.size())
    fail(SIVO_EFORMAT, "caffemodel: blob shape holds %zu values but data has %zu", b.count(), b.data.size());
  return b;
}

}  // namespace

WeightMap read_caffemodel(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) fail(SIVO_ENOENT, "cannot open caffemodel '%s'", path.c_str());
  std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const char kLfs[] = "version https://git-lfs";
  if (raw.size() >= sizeof(kLfs) - 1 && memcmp(raw.data(), kLfs, sizeof(kLfs) - 1) == 0)
    fail(SIVO_EFORMAT, "'%s' is a Git-LFS pointer stub, not a caffemodel", path.c_str());
  WeightMap out;
  Span s{raw.data(), raw.data() + raw.size()};
  Field f1;
  while (next_field(s, f1)) {
    if (f1.num == 2 && f1.wire == 2) fail(SIVO_EFORMAT, "caffemodel: V1LayerParameter files are not supported");
    if (f1.num != 100 || f1.wire != 2) continue;
    Span ls = f1.bytes;
    Field g;
    std::string name;
    std::vector<Blob> blobs;
    while (next_field(ls, g)) {
      if (g.num == 1 && g.wire == 2) name.assign(reinterpret_cast<const char*>(g.bytes.p), g.bytes.end - g.bytes.p);
      else if (g.num == 7 && g.wire == 2) blobs.push_back(parse_blob(g.bytes));
    }
    if (!blobs.empty()) out[name] = std::move(blobs);
  }
  return out;
}

}  // namespace sivo
