// vo_io.cu -- input staging for the frame loop (VO.m:16-17 imageDatastore, VO.m:71-72 readimage;
// SURVEY.md 8f row N1).  At >= 1000 frames/s the 8-bit grayscale PNGs of a KITTI sequence have to be
// inflated and un-filtered by several host threads straight into the (pinned) batch buffer that
// vo_frames uploads.  Host code only: zlib inflate + the five PNG row filters, one file per worker.
// Scope: what KITTI odometry ships -- 8-bit grayscale, non-interlaced.  Anything else is an error,
// not a silent conversion.
#include "vo_internal.h"
#include <zlib.h>
#include <atomic>
#include <thread>

namespace vo {

static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct PngHeader { int rows, cols, depth, color, interlace; };

static int png_header(const uint8_t* f, size_t n, PngHeader* h) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (n < 33 || memcmp(f, sig, 8) != 0) { set_error("png: bad signature"); return VO_ERR_ARG; }
  if (be32(f + 8) != 13 || memcmp(f + 12, "IHDR", 4) != 0) { set_error("png: IHDR is not the first chunk"); return VO_ERR_ARG; }
  h->cols = (int)be32(f + 16); h->rows = (int)be32(f + 20);
  h->depth = f[24]; h->color = f[25]; h->interlace = f[28];
  if (h->rows <= 0 || h->cols <= 0) { set_error("png: empty image"); return VO_ERR_ARG; }
  return VO_OK;
}

static inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// file bytes -> out[rows][ld] (8-bit gray).  scratch: rows * (cols + 1) bytes.
static int png_decode(const uint8_t* f, size_t n, uint8_t* out, int ld, int rows, int cols, std::vector<uint8_t>& idat,
                      std::vector<uint8_t>& raw) {
  PngHeader h;
  VO_TRY(png_header(f, n, &h));
  if (h.rows != rows || h.cols != cols) { set_error("png: image is %d x %d, expected %d x %d", h.rows, h.cols, rows, cols); return VO_ERR_ARG; }
  if (h.depth != 8 || h.color != 0 || h.interlace != 0) {
    set_error("png: only 8-bit grayscale non-interlaced files are supported (depth %d, colour type %d, interlace %d)", h.depth, h.color, h.interlace);
    return VO_ERR_ARG;
  }
  idat.clear();
  size_t pos = 8;
  bool end = false;
  while (pos + 12 <= n && !end) {
    const uint32_t len = be32(f + pos);
    if (pos + 12 + (size_t)len > n) { set_error("png: truncated chunk"); return VO_ERR_ARG; }
    if (memcmp(f + pos + 4, "IDAT", 4) == 0) idat.insert(idat.end(), f + pos + 8, f + pos + 8 + len);
    else if (memcmp(f + pos + 4, "IEND", 4) == 0) end = true;
    pos += 12 + (size_t)len;
  }
  const size_t stride = (size_t)cols + 1;
  raw.resize(stride * rows);
  uLongf got = (uLongf)raw.size();
  const int zr = uncompress(raw.data(), &got, idat.data(), (uLong)idat.size());
  if (zr != Z_OK || got != raw.size()) { set_error("png: inflate failed (zlib %d, %lu of %zu bytes)", zr, (unsigned long)got, raw.size()); return VO_ERR_ARG; }
  for (int y = 0; y < rows; ++y) {
    const uint8_t* src = raw.data() + stride * y;
    uint8_t* dst = out + (size_t)ld * y;
    const uint8_t* up = y ? out + (size_t)ld * (y - 1) : nullptr;
    const int ft = src[0];
    ++src;
    switch (ft) {
      case 0: memcpy(dst, src, cols); break;
      case 1: { int a = 0; for (int x = 0; x < cols; ++x) { a = (src[x] + a) & 255; dst[x] = (uint8_t)a; } break; }
      case 2: for (int x = 0; x < cols; ++x) dst[x] = (uint8_t)(src[x] + (up ? up[x] : 0)); break;
      case 3: { int a = 0; for (int x = 0; x < cols; ++x) { a = (src[x] + ((a + (up ? up[x] : 0)) >> 1)) & 255; dst[x] = (uint8_t)a; } break; }
      case 4: {
        int a = 0, c = 0;
        for (int x = 0; x < cols; ++x) {
          const int b = up ? up[x] : 0;
          a = (src[x] + paeth(a, b, c)) & 255;
          dst[x] = (uint8_t)a;
          c = b;
        }
        break;
      }
      default: set_error("png: bad filter type %d in row %d", ft, y); return VO_ERR_ARG;
    }
  }
  return VO_OK;
}

static int read_file(const char* path, std::vector<uint8_t>& buf) {
  FILE* fp = fopen(path, "rb");
  if (!fp) { set_error("cannot open %s", path); return VO_ERR_ARG; }
  fseek(fp, 0, SEEK_END);
  const long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  buf.resize(sz > 0 ? (size_t)sz : 0);
  const size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, fp) : 0;
  fclose(fp);
  if (got != buf.size()) { set_error("short read on %s", path); return VO_ERR_ARG; }
  return VO_OK;
}

}  // namespace vo

using namespace vo;

extern "C" {

int vo_png_info(const uint8_t* file, size_t n_bytes, int* rows, int* cols, int* bit_depth, int* color_type) {
  VO_CHECK_ARG(file && rows && cols, "null argument");
  PngHeader h;
  VO_TRY(png_header(file, n_bytes, &h));
  *rows = h.rows; *cols = h.cols;
  if (bit_depth) *bit_depth = h.depth;
  if (color_type) *color_type = h.color;
  return VO_OK;
}

int vo_png_decode_gray8(const uint8_t* file, size_t n_bytes, uint8_t* out, int ld, int rows, int cols) {
  VO_CHECK_ARG(file && out && ld >= cols, "null argument or ld < cols");
  std::vector<uint8_t> idat, raw;
  return png_decode(file, n_bytes, out, ld, rows, cols, idat, raw);
}

int vo_png_read_batch(const char* const* paths, int n, int rows, int cols, uint8_t* out, int n_threads) {
  VO_CHECK_ARG(paths && out && n >= 0 && rows > 0 && cols > 0, "bad argument");
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads > n) n_threads = n;
  if (n_threads < 1) n_threads = 1;
  std::atomic<int> next(0), rc(VO_OK);
  std::vector<std::string> errs(n_threads);
  auto work = [&](int t) {
    std::vector<uint8_t> file, idat, raw;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || rc.load() != VO_OK) break;
      int r = read_file(paths[i], file);
      if (r == VO_OK) r = png_decode(file.data(), file.size(), out + (size_t)i * rows * cols, cols, rows, cols, idat, raw);
      if (r != VO_OK) { errs[t] = std::string(paths[i]) + ": " + vo_last_error(); rc.store(r); break; }   // error text is thread-local
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  if (rc.load() != VO_OK)
    for (auto& e : errs)
      if (!e.empty()) { set_error("%s", e.c_str()); break; }
  return rc.load();
}

}  // extern "C"
