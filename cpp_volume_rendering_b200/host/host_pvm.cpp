// host_pvm.cpp -- .pvm volumes, plain and "DDS v3d" / "DDS v3e" compressed.
//
// The reference reads these through Pvm / DDSV3 (libs/file_utils/pvm.cpp:22-41,191-309,382-572; reader.cpp:100-160).
// That decoder is GPL code and is NOT reproduced here: this file is an independent implementation of the FILE FORMAT
// (the facts below), written around a byte-wise bit reader and whole-buffer passes.  tests/test_pvm_dds.py pins it
// against the reference's own decoder compiled in place (oracle/_ref, test-only) on encoder output and on random
// streams, byte for byte.
//
// Differential data stream, as stored on disk:
//   8 bytes  "DDS v3d\n" (byte shuffle over the whole stream) or "DDS v3e\n" (shuffle per block of skip * 2^24 bytes)
//   bit stream, most significant bit first (the reference pulls big-endian 32-bit words; reading past the end yields 0):
//     2 bits   skip - 1      bytes per sample: the payload was split into `skip` byte planes before coding
//    16 bits   strip - 1     row length in bytes used by the predictor
//     repeated runs: 7 bits count (0 ends the stream), 3 bits width code c (width = 0 for c = 0, else c + 1),
//                    then `count` residuals of `width` bits, each stored with a bias of 2^width / 2
//   value[n] = value[n-1] + residual                                       when strip == 1 or n <= strip
//            = value[n-1] + value[n-strip] - value[n-strip-1] + residual   otherwise            (all modulo 256)
//   after decoding, plane-major bytes are put back in sample order: plane i of a block holds bytes i, i+skip, ...
#include "vrbhost.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <string>
#include <vector>

namespace vis {
bool vrb_volume_dims_ok(long long w, long long h, long long d);      // host_data.cpp
namespace dds {

static const size_t kShuffleBlockV3e = (size_t)1 << 24;
static const unsigned kMaxRun = 127;

// ---- bit stream, MSB first, zero tail
class BitSource {
 public:
  BitSource(const uint8_t* p, size_t n) : m_p(p), m_n(n), m_pos(0), m_acc(0), m_have(0) {}
  uint32_t Take(unsigned bits) {   // bits <= 16
    while (m_have < bits) {
      m_acc = (m_acc << 8) | (m_pos < m_n ? m_p[m_pos] : 0u);
      ++m_pos;
      m_have += 8;
    }
    m_have -= bits;
    return (uint32_t)((m_acc >> m_have) & ((1ull << bits) - 1ull));
  }
 private:
  const uint8_t* m_p; size_t m_n, m_pos; uint64_t m_acc; unsigned m_have;
};

class BitSink {
 public:
  explicit BitSink(std::vector<uint8_t>& out) : m_out(out), m_acc(0), m_have(0) {}
  void Put(uint32_t v, unsigned bits) {
    m_acc = (m_acc << bits) | (v & ((1ull << bits) - 1ull));
    m_have += bits;
    while (m_have >= 8) { m_have -= 8; m_out.push_back((uint8_t)(m_acc >> m_have)); }
  }
  void Flush() { if (m_have) { m_out.push_back((uint8_t)(m_acc << (8 - m_have))); m_have = 0; } }
 private:
  std::vector<uint8_t>& m_out; uint64_t m_acc; unsigned m_have;
};

static inline unsigned WidthOfCode(unsigned c) { return c == 0 ? 0u : c + 1u; }
static inline unsigned CodeOfWidth(unsigned w) { return w == 0 ? 0u : w - 1u; }

// Plane-major <-> sample-major byte order, block by block (block_samples == 0: one block).
static void Shuffle(std::vector<uint8_t>& buf, unsigned skip, size_t block_samples, bool to_samples) {
  if (skip <= 1 || buf.empty()) return;
  const size_t n = buf.size();
  const size_t blk = block_samples == 0 ? n : (size_t)skip * block_samples;
  std::vector<uint8_t> tmp(blk < n ? blk : n);
  for (size_t base = 0; base < n; base += blk) {
    const size_t len = n - base < blk ? n - base : blk;
    uint8_t* p = buf.data() + base;
    size_t seq = 0;
    for (unsigned plane = 0; plane < skip; ++plane)
      for (size_t j = plane; j < len; j += skip, ++seq) {
        if (to_samples) tmp[j] = p[seq]; else tmp[seq] = p[j];
      }
    std::memcpy(p, tmp.data(), len);
  }
}

// Largest decoded image accepted: the biggest volume the readers take (vrb_volume_dims_ok: 4096^3 voxels, two bytes each)
// plus a generous text header.  A run of zero-width residuals expands ~100x, so a crafted stream must not be allowed to
// grow the output without bound before the PVM header has even been seen.
static const size_t kMaxDecodedBytes = (size_t)4096 * 4096 * 4096 * 2 + (1u << 20);

bool Decode(const uint8_t* stream, size_t size, size_t block_samples, std::vector<uint8_t>& out, size_t size_hint) {
  BitSource in(stream, size);
  const unsigned skip = in.Take(2) + 1u;
  const size_t strip = (size_t)in.Take(16) + 1u;
  out.clear();
  if (size_hint) out.reserve(size_hint);
  unsigned prev = 0;
  for (;;) {   // ends: past the last byte every field reads 0, and a zero count closes the stream
    const unsigned count = in.Take(7);
    if (count == 0) break;
    const unsigned width = WidthOfCode(in.Take(3));
    const int bias = (1 << width) / 2;
    if (out.size() + count > kMaxDecodedBytes) return false;
    for (unsigned k = 0; k < count; ++k) {
      const size_t n = out.size();
      int v = (int)prev + (int)in.Take(width) - bias;
      if (strip != 1 && n > strip) v += (int)out[n - strip] - (int)out[n - strip - 1];
      prev = (unsigned)v & 255u;
      out.push_back((uint8_t)prev);
    }
  }
  Shuffle(out, skip, block_samples, true);
  return true;
}

// Smallest residual width that holds d (d already reduced to [-128, 127]).
static inline unsigned WidthFor(int d) {
  if (d == 0) return 0;
  for (unsigned w = 2; w < 8; ++w) { int h = 1 << (w - 1); if (d >= -h && d < h) return w; }
  return 8;
}

void Encode(const uint8_t* data, size_t n, unsigned skip, size_t strip, size_t block_samples, std::vector<uint8_t>& out) {
  if (skip < 1) skip = 1;
  if (skip > 4) skip = 4;
  if (strip < 1) strip = 1;
  if (strip > 65536) strip = 65536;
  std::vector<uint8_t> planes(data, data + n);
  Shuffle(planes, skip, block_samples, false);
  // residuals and their widths
  std::vector<int8_t> res(n);
  std::vector<uint8_t> need(n);
  for (size_t i = 0; i < n; ++i) {
    int pred = i ? planes[i - 1] : 0;
    if (strip != 1 && i > strip) pred += (int)planes[i - strip] - (int)planes[i - strip - 1];
    int d = ((int)planes[i] - pred) & 255;
    if (d >= 128) d -= 256;
    res[i] = (int8_t)d;
    need[i] = (uint8_t)WidthFor(d);
  }
  BitSink bits(out);
  bits.Put(skip - 1, 2);
  bits.Put((uint32_t)(strip - 1), 16);
  // Runs: a run keeps the width of its widest member; it is cut when it is full, when a wider residual arrives after
  // a stretch that would have paid for a new header, or when the next few residuals are all narrower.
  const size_t kLook = 6;
  size_t i = 0;
  while (i < n) {
    unsigned w = need[i];
    size_t j = i + 1;
    while (j < n && j - i < kMaxRun) {
      if (need[j] > w) {
        // widening costs (j - i) * (need[j] - w) extra bits for what is already in the run; a new run costs 10
        if ((j - i) * (need[j] - w) > 10) break;
        w = need[j];
      } else if (need[j] < w) {
        size_t m = j, lim = j + kLook < n ? j + kLook : n;
        while (m < lim && need[m] < w) ++m;
        if (m - j == kLook) break;
      }
      ++j;
    }
    // the widest member decides (the loop may have widened past early members only)
    unsigned wr = 0;
    for (size_t k = i; k < j; ++k) if (need[k] > wr) wr = need[k];
    bits.Put((uint32_t)(j - i), 7);
    bits.Put(CodeOfWidth(wr), 3);
    if (wr) { const int bias = (1 << wr) / 2; for (size_t k = i; k < j; ++k) bits.Put((uint32_t)(res[k] + bias), wr); }
    i = j;
  }
  bits.Put(0, 7);
  bits.Flush();
}

static bool ReadFile(const std::string& path, std::vector<uint8_t>& bytes) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  if (std::fseek(f, 0, SEEK_END) != 0) { std::fclose(f); return false; }
  const long n = std::ftell(f);
  if (n < 0 || std::fseek(f, 0, SEEK_SET) != 0) { std::fclose(f); return false; }
  bytes.resize((size_t)n);
  size_t got = bytes.empty() ? 0 : std::fread(bytes.data(), 1, bytes.size(), f);
  std::fclose(f);
  return got == bytes.size();
}

// 0: not a DDS file; 1: "DDS v3d"; 2: "DDS v3e"
int Version(const uint8_t* p, size_t n) {
  if (n >= 8 && std::memcmp(p, "DDS v3d\n", 8) == 0) return 1;
  if (n >= 8 && std::memcmp(p, "DDS v3e\n", 8) == 0) return 2;
  return 0;
}

bool DecodeFileImage(const std::vector<uint8_t>& file, std::vector<uint8_t>& out) {
  int v = Version(file.data(), file.size());
  if (!v) return false;
  return Decode(file.data() + 8, file.size() - 8, v == 1 ? 0 : kShuffleBlockV3e, out, std::min(file.size() * 2, kMaxDecodedBytes));
}

void EncodeFileImage(const uint8_t* data, size_t n, unsigned skip, size_t strip, int version, std::vector<uint8_t>& file) {
  const char* id = version == 2 ? "DDS v3e\n" : "DDS v3d\n";
  file.assign(id, id + 8);
  Encode(data, n, skip, strip, version == 2 ? kShuffleBlockV3e : 0, file);
}

}  // namespace dds

// ---- PVM container (format notes libs/file_utils/pvm.h:10-76; header walk of pvm.cpp:218-256: one line each for the
// magic, "W H D", the spacing (PVM2/PVM3 only) and the component count; '#' comment lines after a plain "PVM" magic).
namespace {
struct Cursor {
  const uint8_t* p; size_t n, pos;
  bool Line(std::string& s) {
    if (pos >= n) return false;
    size_t e = pos;
    while (e < n && p[e] != '\n') ++e;
    if (e >= n) return false;               // header lines end with '\n'
    s.assign((const char*)p + pos, e - pos);
    pos = e + 1;
    return true;
  }
};
}  // namespace

static StructuredGridVolume* ReadPvmImpl(const std::string& filepath);
StructuredGridVolume* VolumeReader::readpvm(std::string filepath) {
  try {
    return ReadPvmImpl(filepath);
  } catch (const std::bad_alloc&) {                  // a huge (or hostile) file must not take the host process down
    vrb::SetError("readpvm: out of memory while reading " + filepath);
    return nullptr;
  }
}
static StructuredGridVolume* ReadPvmImpl(const std::string& filepath) {
  std::vector<uint8_t> file, unpacked;
  if (!dds::ReadFile(filepath, file)) { vrb::SetError("readpvm: cannot open " + filepath); return nullptr; }
  const std::vector<uint8_t>* img = &file;
  if (file.size() >= 4 && std::memcmp(file.data(), "DDS ", 4) == 0) {
    if (!dds::Version(file.data(), file.size())) { vrb::SetError("readpvm: unknown DDS stream version in " + filepath); return nullptr; }
    if (!dds::DecodeFileImage(file, unpacked)) { vrb::SetError("readpvm: oversized DDS stream in " + filepath); return nullptr; }
    img = &unpacked;
  }
  Cursor c{img->data(), img->size(), 0};
  std::string line;
  if (!c.Line(line)) { vrb::SetError("readpvm: bad magic in " + filepath); return nullptr; }
  while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
  const int version = line == "PVM" ? 1 : line == "PVM2" ? 2 : line == "PVM3" ? 3 : 0;
  if (!version) { vrb::SetError("readpvm: bad magic in " + filepath); return nullptr; }
  int w = 0, h = 0, d = 0, comp = 0;
  float sx = 1.0f, sy = 1.0f, sz = 1.0f;   // parsed as float and widened, like Pvm::GetScale(double*) (pvm.cpp:61-66)
  bool ok = c.Line(line);
  while (ok && version == 1 && !line.empty() && line[0] == '#') ok = c.Line(line);
  ok = ok && std::sscanf(line.c_str(), "%d %d %d", &w, &h, &d) == 3;
  if (ok && version >= 2) ok = c.Line(line) && std::sscanf(line.c_str(), "%g %g %g", &sx, &sy, &sz) == 3;
  ok = ok && c.Line(line) && std::sscanf(line.c_str(), "%d", &comp) == 1;
  if (!ok || w < 1 || h < 1 || d < 1 || sx <= 0.0f || sy <= 0.0f || sz <= 0.0f) { vrb::SetError("readpvm: bad header in " + filepath); return nullptr; }
  if (!vrb_volume_dims_ok(w, h, d)) { vrb::SetError("readpvm: implausible sizes in " + filepath); return nullptr; }
  if (comp != 1 && comp != 2) { vrb::SetError("readpvm: only 1- and 2-component volumes are supported (Pvm::PostProcessData): " + filepath); return nullptr; }
  const size_t n = (size_t)w * h * d;
  // the reference insists on an exact size (payload + the four PVM3 strings); here only a short payload is an error
  if (c.n - c.pos < n * comp) { vrb::SetError("readpvm: truncated payload in " + filepath); return nullptr; }
  const uint8_t* raw = c.p + c.pos;
  StructuredGridVolume* vol = new StructuredGridVolume(filepath, w, h, d);
  vol->SetScale((double)sx, (double)sy, (double)sz);
  if (comp == 1) {
    unsigned char* data = new unsigned char[n];
    std::memcpy(data, raw, n);
    vol->SetArrayData(data, _8_BITS);
  } else {
    unsigned short* data = new unsigned short[n];   // second byte is the high byte (pvm.cpp:91-104)
    for (size_t i = 0; i < n; ++i) data[i] = (unsigned short)(raw[2 * i + 1] * 256 + raw[2 * i]);
    vol->SetArrayData(data, _16_BITS);
  }
  return vol;
}

// Writer (the reference ships none that compiles; used by tools and tests): PVM3 header, payload, four empty strings;
// dds_version 0 = plain, 1 = "DDS v3d", 2 = "DDS v3e".  16-bit voxels are stored low byte first.
bool WritePvm(const std::string& path, const void* voxels, int w, int h, int d, int bytes_per_voxel, const double scale[3], int dds_version) {
  if (w < 1 || h < 1 || d < 1 || (bytes_per_voxel != 1 && bytes_per_voxel != 2)) { vrb::SetError("WritePvm: bad arguments"); return false; }
  char head[256];
  int hl = std::snprintf(head, sizeof head, "PVM3\n%d %d %d\n%.9g %.9g %.9g\n%d\n", w, h, d, scale[0], scale[1], scale[2], bytes_per_voxel);
  const size_t n = (size_t)w * h * d;
  std::vector<uint8_t> img(head, head + hl);
  img.resize(hl + n * bytes_per_voxel + 4, 0);
  uint8_t* pay = img.data() + hl;
  if (bytes_per_voxel == 1) std::memcpy(pay, voxels, n);
  else {
    const unsigned short* v = (const unsigned short*)voxels;
    for (size_t i = 0; i < n; ++i) { pay[2 * i] = (uint8_t)(v[i] & 255u); pay[2 * i + 1] = (uint8_t)(v[i] >> 8); }
  }
  std::vector<uint8_t> file;
  if (dds_version == 0) file.swap(img);
  else dds::EncodeFileImage(img.data(), img.size(), (unsigned)bytes_per_voxel, (size_t)w * bytes_per_voxel, dds_version, file);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { vrb::SetError("WritePvm: cannot create " + path); return false; }
  size_t put = std::fwrite(file.data(), 1, file.size(), f);
  std::fclose(f);
  if (put != file.size()) { vrb::SetError("WritePvm: short write to " + path); return false; }
  return true;
}

}  // namespace vis

// ---- C entry points for tests / tools (ctypes)
extern "C" {
// whole DDS file image -> unpacked bytes; returns the unpacked size, -1 when `file` is not a DDS stream, -2 when cap is short
long long vrbh_dds_decode(const unsigned char* file, unsigned long long n, unsigned char* out, unsigned long long cap) {
  std::vector<uint8_t> in(file, file + n), res;
  if (!vis::dds::DecodeFileImage(in, res)) return -1;
  if (res.size() > cap) return -2;
  if (!res.empty()) std::memcpy(out, res.data(), res.size());
  return (long long)res.size();
}
// bytes -> whole DDS file image (version 1 = v3d, 2 = v3e); returns the size needed; writes only when it fits
long long vrbh_dds_encode(const unsigned char* data, unsigned long long n, unsigned skip, unsigned strip, int version,
                          unsigned char* out, unsigned long long cap) {
  std::vector<uint8_t> file;
  vis::dds::EncodeFileImage(data, n, skip, strip, version, file);
  if (file.size() <= cap) std::memcpy(out, file.data(), file.size());
  return (long long)file.size();
}
int vrbh_pvm_write(const char* path, const void* voxels, int w, int h, int d, int bytes_per_voxel, const double* scale, int dds_version) {
  return vis::WritePvm(path, voxels, w, h, d, bytes_per_voxel, scale, dds_version) ? 0 : -1;
}
}
