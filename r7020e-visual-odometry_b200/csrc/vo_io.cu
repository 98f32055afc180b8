// vo_io.cu -- input staging for the frame loop (VO.m:16-17 imageDatastore, VO.m:71-72 readimage;
// SURVEY.md 8f row N1).  At >= 1000 frames/s the 8-bit grayscale PNGs of a KITTI sequence have to be
// inflated and un-filtered by several host threads straight into the (pinned) batch buffer that
// vo_frames uploads.  Host code only: inflate (vo_inflate.cu; zlib gives the verdict on anything that
// decoder rejects) + the five PNG row filters, one file per worker.  The Paeth filter is a serial
// recurrence along a row (each pixel needs its left neighbour), so runs of Paeth rows are un-filtered
// four rows at a time as a skewed wavefront: four independent dependency chains per iteration.
// Scope: what KITTI odometry ships -- 8-bit grayscale, non-interlaced.  Anything else is an error,
// not a silent conversion.
#include "vo_internal.h"
#include "vo_inflate.h"
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

// Paeth predictor, branch-free: pa = |b - c|, pb = |a - c|, pc = |a + b - 2c|
static inline int paeth(int a, int b, int c) {
  const int p = b - c, q = a - c;
  const int pa = abs(p), pb = abs(q), pc = abs(p + q);
  const int bc = pb <= pc ? b : c;
  return (pa <= pb && pa <= pc) ? a : bc;
}

// K consecutive Paeth rows (src: filtered bytes, dst: output rows, up: the finished row above dst[0]
// or zeros).  Row k works on column t - k at step t: its "up" pixel is what row k-1 produced one step
// earlier, so the K recurrences of a step are independent of each other.
template <int K>
static void unfilter_paeth_rows(const uint8_t* const* src, uint8_t* const* dst, const uint8_t* up, int cols) {
  int a[K], c[K];
  for (int k = 0; k < K; ++k) a[k] = c[k] = 0;
  auto step = [&](int t, auto checked) {
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
      const int x = t - k;
      if (decltype(checked)::value && (x < 0 || x >= cols)) continue;
      const int b = k ? a[k - 1] : up[x];          // a[k-1] is still row k-1's value at column x
      const int v = (src[k][x] + paeth(a[k], b, c[k])) & 255;
      c[k] = b; a[k] = v;
      dst[k][x] = (uint8_t)v;
    }
  };
  // rows k >= 1 read a[k-1] before row k-1 overwrites it (k runs downwards); at column 0 the left and
  // up-left neighbours are 0, and row k-1 has just finished column x when row k needs it
  int t = 0;
  for (; t < K - 1 && t < cols + K - 1; ++t) step(t, std::true_type{});
  for (; t < cols; ++t) step(t, std::false_type{});
  for (; t < cols + K - 1; ++t) step(t, std::true_type{});
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
  const size_t stride = (size_t)cols + 1, n_raw = stride * rows, n_idat = idat.size();
  raw.resize(n_raw + INFLATE_OUT_SLACK);
  idat.resize(n_idat + INFLATE_IN_SLACK, 0);
  if (!inflate_zlib_fast(idat.data(), n_idat, raw.data(), n_raw)) {   // zlib decides what is wrong (or that nothing is)
    uLongf got = (uLongf)n_raw;
    const int zr = uncompress(raw.data(), &got, idat.data(), (uLong)n_idat);
    if (zr != Z_OK || got != n_raw) { set_error("png: inflate failed (zlib %d, %lu of %zu bytes)", zr, (unsigned long)got, n_raw); return VO_ERR_ARG; }
  }
  std::vector<uint8_t> zero_row;
  for (int y = 0; y < rows; ++y) {
    const uint8_t* src = raw.data() + stride * y;
    uint8_t* dst = out + (size_t)ld * y;
    const uint8_t* up = y ? out + (size_t)ld * (y - 1) : nullptr;
    const int ft = src[0];
    ++src;
    if (ft == 4) {   // a run of Paeth rows: four (or two) at a time as a wavefront
      int run = 1;
      while (run < 4 && y + run < rows && raw[stride * (y + run)] == 4) ++run;
      if (run >= 2) {
        if (run == 3) run = 2;
        if (!up) { zero_row.assign((size_t)cols, 0); up = zero_row.data(); }
        const uint8_t* s[4]; uint8_t* d[4];
        for (int k = 0; k < run; ++k) { s[k] = raw.data() + stride * (y + k) + 1; d[k] = out + (size_t)ld * (y + k); }
        if (run == 4) unfilter_paeth_rows<4>(s, d, up, cols); else unfilter_paeth_rows<2>(s, d, up, cols);
        y += run - 1;
        continue;
      }
    }
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

int vo_inflate_zlib(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out) {
  VO_CHECK_ARG(in && (out || n_out == 0), "null argument");
  std::vector<uint8_t> src(n_in + INFLATE_IN_SLACK, 0), dst(n_out + INFLATE_OUT_SLACK);
  memcpy(src.data(), in, n_in);
  if (!inflate_zlib_fast(src.data(), n_in, dst.data(), n_out)) { set_error("inflate: malformed zlib stream, or it does not inflate to exactly %zu bytes", n_out); return VO_ERR_ARG; }
  if (n_out) memcpy(out, dst.data(), n_out);
  return VO_OK;
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
