// Corpus file IO of the feature pipeline, host side: the formats on either side of the GPU batch.
//
// The reference reads one wav (soundfile) and writes four .npz files (numpy.savez) per utterance from a Python loop
// (idiaptts/src/data_preparation/world/WorldFeatLabelGen.py:996-1013 -> :1121-1172, LabelGen.py:63-101) and reads them back one
// array at a time (WorldFeatLabelGen.py:459-567).  Once extraction runs at tens of thousands of audio-seconds per second that loop
// IS the wall time, so the file side works on whole shards: a pool of threads reads PCM data straight into the packed (pinned)
// sample buffer the kernels consume and writes / reads .npz archives from / into the packed feature matrix.  No device code here.
//
// Formats (bit-compatible with the Python side):
//   * wav: RIFF/WAVE, 'fmt ' tag 1 (PCM) or 0xFFFE (extensible, PCM sub-format); the packed reader takes 16-bit mono, everything
//     else is reported by the probe so that the caller can take its general path.
//   * npz: ZIP archive, method 0 (stored), one member "<key>.npy" per array; .npy version 1.0 header, little-endian float32, C order,
//     shape (rows, cols) -- what numpy.savez writes and numpy.load reads.  numpy writes ZIP64 local headers (force_zip64) with the
//     true sizes in the central directory; the reader takes sizes and offsets from the central directory (ZIP64 extra fields
//     included), so both kinds load.  Compressed members (savez_compressed) are refused with a message: the caller falls back.
#include <fcntl.h>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b2w {
namespace {

// ---- a pool of threads over file indices; the first error message wins -------------------------------------------------------
struct Errors {
  std::mutex mu;
  std::string first;
  void put(const std::string& s) {
    std::lock_guard<std::mutex> g(mu);
    if (first.empty()) first = s;
  }
};

template <typename F>
int for_each_file(int n, int threads, const char* what, F&& fn) {
  Errors err;
  std::atomic<int> next{0};
  std::atomic<bool> failed{false};
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if (nt > n) nt = n;
  auto work = [&]() {
    std::string msg;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || failed.load(std::memory_order_relaxed)) break;
      msg.clear();
      if (!fn(i, msg)) {
        failed.store(true);
        err.put(msg);
      }
    }
  };
  if (nt <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  }
  if (failed.load()) {
    set_error("%s: %s", what, err.first.c_str());
    return -1;
  }
  return 0;
}

// ---- small file helpers ---------------------------------------------------------------------------------------------------
struct Fd {
  int fd = -1;
  ~Fd() {
    if (fd >= 0) ::close(fd);
  }
};

bool read_at(int fd, void* dst, size_t n, int64_t off) {
  char* p = static_cast<char*>(dst);
  while (n > 0) {
    const ssize_t r = ::pread(fd, p, n, off);
    if (r <= 0) return false;
    p += r;
    off += r;
    n -= (size_t)r;
  }
  return true;
}

bool write_all(int fd, const void* src, size_t n) {
  const char* p = static_cast<const char*>(src);
  while (n > 0) {
    const ssize_t r = ::write(fd, p, n);
    if (r <= 0) return false;
    p += r;
    n -= (size_t)r;
  }
  return true;
}

inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const unsigned char* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
inline void wr16(std::vector<unsigned char>& v, uint32_t x) {
  v.push_back((unsigned char)(x & 255));
  v.push_back((unsigned char)((x >> 8) & 255));
}
inline void wr32(std::vector<unsigned char>& v, uint32_t x) {
  wr16(v, x & 0xffff);
  wr16(v, x >> 16);
}

// ---- CRC-32 (ZIP, polynomial 0xEDB88320), eight bytes per step ----------------------------------------------------------------
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 255];
  }
};
const CrcTables& crc_tables() {
  static const CrcTables tables;
  return tables;
}
#if defined(__x86_64__) && defined(__GNUC__)
#define B2W_CRC_CLMUL 1
// Carry-less-multiplication folding (V. Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction",
// Intel 2009; constants of the reflected CRC-32 polynomial as zlib-family libraries use them): folds 64 bytes per step.
// `state` is the running (inverted) register; n >= 64 and a multiple of 16.  Checked against the table code in the tests.
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(uint32_t state, const unsigned char* buf, size_t n) {
  alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
  alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
  alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
  alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x00));
  x2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x10));
  x3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x20));
  x4 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)state));
  x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(k1k2));
  buf += 64;
  n -= 64;
  while (n >= 64) {  // four independent 128-bit lanes
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
    x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
    x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x00));
    y6 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x10));
    y7 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x20));
    y8 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
    x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
    x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64;
    n -= 64;
  }
  x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(k3k4));  // four lanes into one
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (n >= 16) {  // remaining whole 16-byte blocks
    x2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf));
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16;
    n -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);  // 128 -> 64 bits
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(k5k0));
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(poly));  // Barrett reduction to 32 bits
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
bool have_clmul() {
  static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
  return ok;
}
#endif

uint32_t crc32_update(uint32_t crc, const void* data, size_t n, bool allow_clmul = true) {
  const CrcTables& T = crc_tables();
  const unsigned char* p = static_cast<const unsigned char*>(data);
  crc = ~crc;
#ifdef B2W_CRC_CLMUL
  if (allow_clmul && n >= 64 && have_clmul()) {
    const size_t body = n & ~(size_t)15;
    crc = crc32_clmul(crc, p, body);
    p += body;
    n -= body;
  }
#endif
  while (n >= 8) {
    uint32_t a, b;
    memcpy(&a, p, 4);
    memcpy(&b, p + 4, 4);
    a ^= crc;
    crc = T.t[7][a & 255] ^ T.t[6][(a >> 8) & 255] ^ T.t[5][(a >> 16) & 255] ^ T.t[4][a >> 24] ^ T.t[3][b & 255] ^
          T.t[2][(b >> 8) & 255] ^ T.t[1][(b >> 16) & 255] ^ T.t[0][b >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = T.t[0][(crc ^ *p++) & 255] ^ (crc >> 8);
  return ~crc;
}

// ---- wav ------------------------------------------------------------------------------------------------------------------
struct WavInfo {
  int64_t num_samples = 0, data_off = 0;
  int fs = 0, bits = 0, channels = 0;
};

bool wav_probe_one(const char* path, WavInfo& w, std::string& msg) {
  Fd f;
  f.fd = ::open(path, O_RDONLY);
  if (f.fd < 0) {
    msg = std::string("cannot open ") + path;
    return false;
  }
  struct stat st;
  if (fstat(f.fd, &st) != 0) {
    msg = std::string("cannot stat ") + path;
    return false;
  }
  unsigned char h[12];
  if (!read_at(f.fd, h, 12, 0) || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) {
    msg = std::string(path) + " is not a RIFF/WAVE file";
    return false;
  }
  int64_t pos = 12;
  bool have_fmt = false;
  int block_align = 0;
  while (pos + 8 <= (int64_t)st.st_size) {
    unsigned char c[8];
    if (!read_at(f.fd, c, 8, pos)) break;
    const int64_t size = rd32(c + 4);
    if (memcmp(c, "fmt ", 4) == 0) {
      unsigned char b[40] = {0};
      const size_t nb = (size_t)(size < 40 ? size : 40);
      if (size < 16 || !read_at(f.fd, b, nb, pos + 8)) {
        msg = std::string(path) + ": truncated fmt chunk";
        return false;
      }
      int tag = rd16(b);
      w.channels = rd16(b + 2);
      w.fs = (int)rd32(b + 4);
      block_align = rd16(b + 12);
      w.bits = rd16(b + 14);
      if (tag == 0xFFFE && size >= 26) tag = rd16(b + 24);  // extensible: the sub-format GUID starts with the format tag
      if (tag != 1) {
        msg = std::string(path) + ": not a PCM wav (format tag " + std::to_string(tag) + ")";
        return false;
      }
      have_fmt = true;
    } else if (memcmp(c, "data", 4) == 0) {
      if (!have_fmt || block_align <= 0) {
        msg = std::string(path) + ": data chunk before fmt chunk";
        return false;
      }
      int64_t bytes = size;
      if (pos + 8 + bytes > (int64_t)st.st_size) bytes = (int64_t)st.st_size - pos - 8;  // streamed files leave the size open
      w.data_off = pos + 8;
      w.num_samples = bytes / block_align;
      return true;
    }
    pos += 8 + size + (size & 1);
  }
  msg = std::string(path) + ": no data chunk";
  return false;
}

// ---- npy / npz --------------------------------------------------------------------------------------------------------------
// .npy 1.0 header of a C-ordered little-endian float32 matrix, padded so that the data start at a multiple of 64
std::string npy_header(int64_t rows, int cols) {
  std::string d = "{'descr': '<f4', 'fortran_order': False, 'shape': (" + std::to_string(rows) + ", " + std::to_string(cols) + "), }";
  size_t total = 10 + d.size() + 1;
  const size_t pad = (64 - total % 64) % 64;
  d.append(pad, ' ');
  d.push_back('\n');
  std::string h("\x93NUMPY\x01\x00", 8);
  h.push_back((char)(d.size() & 255));
  h.push_back((char)(d.size() >> 8));
  return h + d;
}

struct Member {
  std::string name;
  uint32_t crc = 0;
  uint64_t size = 0, offset = 0;
  int method = 0;
};

// central directory of a ZIP archive (ZIP64 end record and extra fields understood)
bool zip_directory(int fd, const char* path, std::vector<Member>& out, std::string& msg) {
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 22) {
    msg = std::string(path) + " is not a zip archive";
    return false;
  }
  const int64_t fsize = st.st_size;
  const int64_t tail = fsize < 65557 ? fsize : 65557;
  std::vector<unsigned char> buf((size_t)tail);
  if (!read_at(fd, buf.data(), (size_t)tail, fsize - tail)) {
    msg = std::string("cannot read ") + path;
    return false;
  }
  int64_t e = -1;
  for (int64_t i = tail - 22; i >= 0; --i)
    if (rd32(&buf[(size_t)i]) == 0x06054b50u) {
      e = i;
      break;
    }
  if (e < 0) {
    msg = std::string(path) + ": no end-of-central-directory record";
    return false;
  }
  uint64_t n = rd16(&buf[(size_t)e + 10]), cd_size = rd32(&buf[(size_t)e + 12]), cd_off = rd32(&buf[(size_t)e + 16]);
  if ((n == 0xffff || cd_size == 0xffffffffu || cd_off == 0xffffffffu) && e >= 20 && rd32(&buf[(size_t)e - 20]) == 0x07064b50u) {
    const uint64_t z64 = rd64(&buf[(size_t)e - 20 + 8]);
    unsigned char r[56];
    if (!read_at(fd, r, 56, (int64_t)z64) || rd32(r) != 0x06064b50u) {
      msg = std::string(path) + ": bad ZIP64 end record";
      return false;
    }
    n = rd64(r + 32);
    cd_size = rd64(r + 40);
    cd_off = rd64(r + 48);
  }
  if (cd_off + cd_size > (uint64_t)fsize) {
    msg = std::string(path) + ": central directory out of range";
    return false;
  }
  std::vector<unsigned char> cd((size_t)cd_size);
  if (cd_size && !read_at(fd, cd.data(), (size_t)cd_size, (int64_t)cd_off)) {
    msg = std::string("cannot read ") + path;
    return false;
  }
  size_t p = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (p + 46 > cd.size() || rd32(&cd[p]) != 0x02014b50u) {
      msg = std::string(path) + ": bad central directory entry";
      return false;
    }
    Member m;
    m.method = rd16(&cd[p + 10]);
    m.crc = rd32(&cd[p + 16]);
    uint64_t csize = rd32(&cd[p + 20]);
    m.size = rd32(&cd[p + 24]);
    const size_t nl = rd16(&cd[p + 28]), xl = rd16(&cd[p + 30]), cl = rd16(&cd[p + 32]);
    m.offset = rd32(&cd[p + 42]);
    if (p + 46 + nl + xl + cl > cd.size()) {
      msg = std::string(path) + ": bad central directory entry";
      return false;
    }
    m.name.assign(reinterpret_cast<const char*>(&cd[p + 46]), nl);
    // ZIP64 extra field: the values whose 32-bit slots are saturated, in the order usize, csize, offset
    size_t x = p + 46 + nl;
    const size_t xe = x + xl;
    while (x + 4 <= xe) {
      const int id = rd16(&cd[x]);
      const size_t len = rd16(&cd[x + 2]);
      if (id == 1) {
        size_t q = x + 4;
        if (m.size == 0xffffffffu && q + 8 <= xe) { m.size = rd64(&cd[q]); q += 8; }
        if (csize == 0xffffffffu && q + 8 <= xe) { csize = rd64(&cd[q]); q += 8; }
        if (m.offset == 0xffffffffu && q + 8 <= xe) { m.offset = rd64(&cd[q]); q += 8; }
      }
      x += 4 + len;
    }
    (void)csize;
    out.push_back(std::move(m));
    p += 46 + nl + xl + cl;
  }
  return true;
}

// where the member's bytes start (after its local header)
bool zip_data_offset(int fd, const char* path, const Member& m, uint64_t& data_off, std::string& msg) {
  unsigned char h[30];
  if (!read_at(fd, h, 30, (int64_t)m.offset) || rd32(h) != 0x04034b50u) {
    msg = std::string(path) + ": bad local header of " + m.name;
    return false;
  }
  data_off = m.offset + 30 + rd16(h + 26) + rd16(h + 28);
  return true;
}

struct NpyShape {
  int64_t rows = 0;
  int cols = 0;
  uint64_t data_off = 0;  // within the file
};

// parse the .npy header of member m: float32, C order, 1-D (cols = 1) or 2-D
bool npy_open(int fd, const char* path, const Member& m, NpyShape& s, std::string& msg) {
  if (m.method != 0) {
    msg = std::string(path) + ": member " + m.name + " is compressed (numpy.savez_compressed); only stored archives are read natively";
    return false;
  }
  uint64_t off;
  if (!zip_data_offset(fd, path, m, off, msg)) return false;
  unsigned char pre[12];
  if (m.size < 10 || !read_at(fd, pre, 12 <= m.size ? 12 : 10, (int64_t)off) || memcmp(pre, "\x93NUMPY", 6) != 0) {
    msg = std::string(path) + ": member " + m.name + " is not an .npy array";
    return false;
  }
  size_t hlen, hoff;
  if (pre[6] == 1) {
    hlen = rd16(pre + 8);
    hoff = 10;
  } else {
    hlen = rd32(pre + 8);
    hoff = 12;
  }
  if (hoff + hlen > m.size || hlen > (1u << 20)) {
    msg = std::string(path) + ": bad .npy header in " + m.name;
    return false;
  }
  std::string d(hlen, '\0');
  if (!read_at(fd, &d[0], hlen, (int64_t)(off + hoff))) {
    msg = std::string("cannot read ") + path;
    return false;
  }
  auto value_after = [&](const char* key) -> size_t {
    const size_t k = d.find(key);
    if (k == std::string::npos) return k;
    const size_t c = d.find(':', k);
    return c == std::string::npos ? c : d.find_first_not_of(' ', c + 1);
  };
  const size_t pd = value_after("'descr'"), pf = value_after("'fortran_order'"), ps = value_after("'shape'");
  if (pd == std::string::npos || pf == std::string::npos || ps == std::string::npos || d[ps] != '(') {
    msg = std::string(path) + ": unreadable .npy header in " + m.name;
    return false;
  }
  if (d.compare(pd, 5, "'<f4'") != 0 || d.compare(pf, 5, "False") != 0) {
    msg = std::string(path) + ": member " + m.name + " is not a C-ordered little-endian float32 array";
    return false;
  }
  const size_t pe = d.find(')', ps);
  std::vector<int64_t> dims;
  size_t q = ps + 1;
  while (q < pe) {
    while (q < pe && (d[q] == ' ' || d[q] == ',')) ++q;
    if (q >= pe) break;
    int64_t v = 0;
    bool any = false;
    while (q < pe && d[q] >= '0' && d[q] <= '9') {
      v = v * 10 + (d[q] - '0');
      ++q;
      any = true;
    }
    if (!any) {
      msg = std::string(path) + ": unreadable shape in " + m.name;
      return false;
    }
    dims.push_back(v);
  }
  if (dims.size() == 1) {
    s.rows = dims[0];
    s.cols = 1;
  } else if (dims.size() == 2 && dims[1] <= 0x7fffffff) {
    s.rows = dims[0];
    s.cols = (int)dims[1];
  } else {
    msg = std::string(path) + ": member " + m.name + " is not a 1-D or 2-D array";
    return false;
  }
  s.data_off = off + hoff + hlen;
  if ((uint64_t)s.rows * (uint64_t)s.cols * 4 + hoff + hlen != m.size) {
    msg = std::string(path) + ": size of member " + m.name + " does not match its shape";
    return false;
  }
  return true;
}

const Member* find_member(const std::vector<Member>& dir, const char* key) {
  const std::string want = std::string(key) + ".npy";
  for (const Member& m : dir)
    if (m.name == want) return &m;
  return nullptr;
}

}  // namespace
}  // namespace b2w

using namespace b2w;

extern "C" {

int b2w_wav_probe(const char* const* paths, int32_t num_files, int64_t* num_samples, int32_t* fs, int32_t* bits, int32_t* channels,
                  int64_t* data_offset, int32_t threads) {
  B2W_REQUIRE(num_files >= 0 && (num_files == 0 || (paths && num_samples && fs && bits && channels && data_offset)),
              "b2w_wav_probe: null argument");
  return for_each_file(num_files, threads, "b2w_wav_probe", [&](int i, std::string& msg) {
    WavInfo w;
    if (!wav_probe_one(paths[i], w, msg)) return false;
    num_samples[i] = w.num_samples;
    fs[i] = w.fs;
    bits[i] = w.bits;
    channels[i] = w.channels;
    data_offset[i] = w.data_off;
    return true;
  });
}

int b2w_wav_read_i16(const char* const* paths, int32_t num_files, const int64_t* data_offset, const int64_t* utt_sample_offset,
                     int16_t* samples, int32_t threads) {
  B2W_REQUIRE(num_files >= 0 && (num_files == 0 || (paths && data_offset && utt_sample_offset && samples)),
              "b2w_wav_read_i16: null argument");
  return for_each_file(num_files, threads, "b2w_wav_read_i16", [&](int i, std::string& msg) {
    Fd f;
    f.fd = ::open(paths[i], O_RDONLY);
    const int64_t n = utt_sample_offset[i + 1] - utt_sample_offset[i];
    if (f.fd < 0 || n < 0 || !read_at(f.fd, samples + utt_sample_offset[i], (size_t)n * 2, data_offset[i])) {
      msg = std::string("cannot read ") + std::to_string((long long)n) + " samples from " + paths[i];
      return false;
    }
    return true;
  });
}

int b2w_wav_write_pcm16(const char* const* paths, int32_t num_files, const int64_t* utt_sample_offset, const float* samples, int32_t fs,
                        int32_t threads) {
  B2W_REQUIRE(num_files >= 0 && fs > 0 && (num_files == 0 || (paths && utt_sample_offset && samples)), "b2w_wav_write_pcm16: null argument");
  return for_each_file(num_files, threads, "b2w_wav_write_pcm16", [&](int i, std::string& msg) {
    const int64_t n = utt_sample_offset[i + 1] - utt_sample_offset[i];
    if (n < 0 || n * 2 + 36 > 0xffffffffll) {
      msg = std::string(paths[i]) + ": sample count outside what a RIFF file holds";
      return false;
    }
    std::vector<unsigned char> out;
    out.reserve(44 + (size_t)n * 2);
    const char* riff = "RIFF";
    out.insert(out.end(), riff, riff + 4);
    wr32(out, (uint32_t)(36 + n * 2));
    const char* wavefmt = "WAVEfmt ";
    out.insert(out.end(), wavefmt, wavefmt + 8);
    wr32(out, 16);
    wr16(out, 1);  // PCM
    wr16(out, 1);  // mono
    wr32(out, (uint32_t)fs);
    wr32(out, (uint32_t)fs * 2);
    wr16(out, 2);
    wr16(out, 16);
    const char* data = "data";
    out.insert(out.end(), data, data + 4);
    wr32(out, (uint32_t)(n * 2));
    out.resize(44 + (size_t)n * 2);
    int16_t* pcm = reinterpret_cast<int16_t*>(&out[44]);
    const float* src = samples + utt_sample_offset[i];
    for (int64_t k = 0; k < n; ++k) {
      double v = nearbyint((double)src[k] * 32767.0);  // round half to even, as numpy.round
      v = v < -32768.0 ? -32768.0 : (v > 32767.0 ? 32767.0 : v);
      pcm[k] = (int16_t)(v == v ? v : 0.0);
    }
    Fd f;
    f.fd = ::open(paths[i], O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (f.fd < 0 || !write_all(f.fd, out.data(), out.size())) {
      msg = std::string("cannot write ") + paths[i];
      return false;
    }
    return true;
  });
}

int b2w_npz_write_f32(const char* const* paths, int32_t num_files, const char* const* keys, int32_t num_keys, const int32_t* col_offset,
                      const int32_t* cols, const int64_t* utt_frame_offset, const float* feats, int64_t feat_stride, int32_t threads) {
  B2W_REQUIRE(num_files >= 0 && num_keys > 0 && paths && keys && col_offset && cols && utt_frame_offset && feats,
              "b2w_npz_write_f32: null argument");
  for (int k = 0; k < num_keys; ++k)
    B2W_REQUIRE(cols[k] > 0 && col_offset[k] >= 0 && col_offset[k] + (int64_t)cols[k] <= feat_stride && strlen(keys[k]) < 200,
                "b2w_npz_write_f32: key %d: columns [%d, %d) outside a row of %lld", k, col_offset[k], col_offset[k] + cols[k],
                (long long)feat_stride);
  return for_each_file(num_files, threads, "b2w_npz_write_f32", [&](int i, std::string& msg) {
    const int64_t r0 = utt_frame_offset[i], rows = utt_frame_offset[i + 1] - r0;
    if (rows < 0) {
      msg = "negative row count";
      return false;
    }
    // the whole archive is assembled in memory (a few hundred KB, uninitialised buffer sized up front) and written with one call
    std::vector<std::string> heads(num_keys), names(num_keys);
    size_t total = 22;
    for (int k = 0; k < num_keys; ++k) {
      names[k] = std::string(keys[k]) + ".npy";
      heads[k] = npy_header(rows, cols[k]);
      const uint64_t size = heads[k].size() + (uint64_t)rows * cols[k] * 4;
      total += 30 + 46 + 2 * names[k].size() + (size_t)size;
      if (size >= 0xffffffffull || total >= 0xffffffffull) {
        msg = std::string(paths[i]) + ": arrays of 4 GiB and more need ZIP64";
        return false;
      }
    }
    std::unique_ptr<unsigned char[]> buf(new unsigned char[total]);
    unsigned char* const out = buf.get();
    size_t pos = 0;
    auto put16 = [&](uint32_t x) {
      out[pos++] = (unsigned char)(x & 255);
      out[pos++] = (unsigned char)((x >> 8) & 255);
    };
    auto put32 = [&](uint32_t x) {
      put16(x & 0xffff);
      put16(x >> 16);
    };
    auto put = [&](const void* p, size_t n) {
      memcpy(out + pos, p, n);
      pos += n;
    };
    std::vector<unsigned char> central;
    for (int k = 0; k < num_keys; ++k) {
      const std::string& name = names[k];
      const std::string& head = heads[k];
      const size_t bytes = (size_t)rows * cols[k] * 4;
      const uint32_t size = (uint32_t)(head.size() + bytes);
      const uint32_t local_off = (uint32_t)pos;
      put32(0x04034b50u);
      put16(20);
      put16(0);
      put16(0);
      put16(0);
      put16(0x21);  // 1980-01-01
      const size_t crc_pos = pos;
      put32(0);
      put32(size);
      put32(size);
      put16((uint32_t)name.size());
      put16(0);
      put(name.data(), name.size());
      const size_t data_pos = pos;
      put(head.data(), head.size());
      const float* src = feats + r0 * feat_stride + col_offset[k];
      if (cols[k] == feat_stride) {
        put(src, bytes);
      } else if (cols[k] <= 4) {  // narrow column blocks (lf0, vuv, bap): element loop instead of a memcpy call per row
        float* dst = reinterpret_cast<float*>(out + pos);  // (the member's offset in the archive need not be a multiple of 4)
        const int c = cols[k];
        if ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
          for (int64_t r = 0; r < rows; ++r)
            for (int e = 0; e < c; ++e) dst[r * c + e] = src[r * feat_stride + e];
        } else {
          for (int64_t r = 0; r < rows; ++r) memcpy(out + pos + (size_t)r * c * 4, src + r * feat_stride, (size_t)c * 4);
        }
        pos += bytes;
      } else {
        const size_t rb = (size_t)cols[k] * 4;
        for (int64_t r = 0; r < rows; ++r) memcpy(out + pos + r * rb, src + r * feat_stride, rb);
        pos += bytes;
      }
      const uint32_t crc = crc32_update(0, out + data_pos, (size_t)size);
      out[crc_pos] = (unsigned char)(crc & 255);
      out[crc_pos + 1] = (unsigned char)((crc >> 8) & 255);
      out[crc_pos + 2] = (unsigned char)((crc >> 16) & 255);
      out[crc_pos + 3] = (unsigned char)(crc >> 24);
      wr32(central, 0x02014b50u);
      wr16(central, 20);
      wr16(central, 20);
      wr16(central, 0);
      wr16(central, 0);
      wr16(central, 0);
      wr16(central, 0x21);
      wr32(central, crc);
      wr32(central, size);
      wr32(central, size);
      wr16(central, (uint32_t)name.size());
      wr16(central, 0);
      wr16(central, 0);
      wr16(central, 0);
      wr16(central, 0);
      wr32(central, 0);
      wr32(central, local_off);
      central.insert(central.end(), name.begin(), name.end());
    }
    const uint32_t cd_off = (uint32_t)pos;
    put(central.data(), central.size());
    put32(0x06054b50u);
    put16(0);
    put16(0);
    put16((uint32_t)num_keys);
    put16((uint32_t)num_keys);
    put32((uint32_t)central.size());
    put32(cd_off);
    put16(0);
    Fd f;
    f.fd = ::open(paths[i], O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (f.fd < 0 || pos != total || !write_all(f.fd, out, pos)) {
      msg = std::string("cannot write ") + paths[i];
      return false;
    }
    return true;
  });
}

int b2w_npz_probe(const char* const* paths, int32_t num_files, const char* key, int64_t* rows, int32_t* cols, int32_t threads) {
  B2W_REQUIRE(num_files >= 0 && key && (num_files == 0 || (paths && rows && cols)), "b2w_npz_probe: null argument");
  return for_each_file(num_files, threads, "b2w_npz_probe", [&](int i, std::string& msg) {
    Fd f;
    f.fd = ::open(paths[i], O_RDONLY);
    if (f.fd < 0) {
      msg = std::string("cannot open ") + paths[i];
      return false;
    }
    std::vector<Member> dir;
    if (!zip_directory(f.fd, paths[i], dir, msg)) return false;
    const Member* m = find_member(dir, key);
    if (!m) {
      msg = std::string(paths[i]) + ": no array '" + key + "'";
      return false;
    }
    NpyShape s;
    if (!npy_open(f.fd, paths[i], *m, s, msg)) return false;
    rows[i] = s.rows;
    cols[i] = s.cols;
    return true;
  });
}

int b2w_npz_read_f32(const char* const* paths, int32_t num_files, const char* const* keys, int32_t num_keys, const int32_t* col_offset,
                     const int32_t* cols, const int64_t* utt_frame_offset, float* feats, int64_t feat_stride, int32_t verify_crc,
                     int32_t threads) {
  B2W_REQUIRE(num_files >= 0 && num_keys > 0 && paths && keys && col_offset && cols && utt_frame_offset && feats,
              "b2w_npz_read_f32: null argument");
  for (int k = 0; k < num_keys; ++k)
    B2W_REQUIRE(cols[k] > 0 && col_offset[k] >= 0 && col_offset[k] + (int64_t)cols[k] <= feat_stride,
                "b2w_npz_read_f32: key %d: columns [%d, %d) outside a row of %lld", k, col_offset[k], col_offset[k] + cols[k],
                (long long)feat_stride);
  return for_each_file(num_files, threads, "b2w_npz_read_f32", [&](int i, std::string& msg) {
    Fd f;
    f.fd = ::open(paths[i], O_RDONLY);
    if (f.fd < 0) {
      msg = std::string("cannot open ") + paths[i];
      return false;
    }
    std::vector<Member> dir;
    if (!zip_directory(f.fd, paths[i], dir, msg)) return false;
    const int64_t r0 = utt_frame_offset[i], rows = utt_frame_offset[i + 1] - r0;
    std::vector<float> tmp;
    for (int k = 0; k < num_keys; ++k) {
      const Member* m = find_member(dir, keys[k]);
      if (!m) {
        msg = std::string(paths[i]) + ": no array '" + keys[k] + "'";
        return false;
      }
      NpyShape s;
      if (!npy_open(f.fd, paths[i], *m, s, msg)) return false;
      if (s.rows != rows || s.cols != cols[k]) {
        msg = std::string(paths[i]) + ": array '" + keys[k] + "' has shape (" + std::to_string((long long)s.rows) + ", " +
              std::to_string(s.cols) + "), expected (" + std::to_string((long long)rows) + ", " + std::to_string(cols[k]) + ")";
        return false;
      }
      float* dst = feats + r0 * feat_stride + col_offset[k];
      const size_t bytes = (size_t)rows * cols[k] * 4;
      const float* got;
      if (cols[k] == feat_stride) {
        if (!read_at(f.fd, dst, bytes, (int64_t)s.data_off)) {
          msg = std::string("cannot read ") + paths[i];
          return false;
        }
        got = dst;
      } else {
        tmp.resize((size_t)rows * cols[k]);
        if (bytes && !read_at(f.fd, tmp.data(), bytes, (int64_t)s.data_off)) {
          msg = std::string("cannot read ") + paths[i];
          return false;
        }
        const size_t rb = (size_t)cols[k] * 4;
        for (int64_t r = 0; r < rows; ++r) memcpy(dst + r * feat_stride, tmp.data() + r * cols[k], rb);
        got = tmp.data();
      }
      if (verify_crc) {
        // the checksum covers the .npy header too
        const size_t hbytes = (size_t)(m->size - bytes);
        std::vector<unsigned char> head(hbytes);
        if (!read_at(f.fd, head.data(), hbytes, (int64_t)(s.data_off - hbytes))) {
          msg = std::string("cannot read ") + paths[i];
          return false;
        }
        uint32_t crc = crc32_update(0, head.data(), hbytes);
        crc = crc32_update(crc, got, bytes);
        if (crc != m->crc) {
          msg = std::string(paths[i]) + ": bad CRC-32 in array '" + keys[k] + "'";
          return false;
        }
      }
    }
    return true;
  });
}

/* development aid (tests): CRC-32 of a host buffer by the table code (variant 0) or with carry-less folding where the CPU has it
 * (variant 1) */
uint32_t b2w_crc32(const void* data, int64_t n, int32_t variant) { return crc32_update(0, data, (size_t)n, variant != 0); }

}  // extern "C"
