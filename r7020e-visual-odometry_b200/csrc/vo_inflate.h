// vo_inflate.h -- fast zlib-stream decoder used by the PNG input stage (vo_io.cu).
#pragma once
#include <cstddef>
#include <cstdint>

namespace vo {

constexpr size_t INFLATE_IN_SLACK = 32;   // readable bytes required after the stream
constexpr size_t INFLATE_OUT_SLACK = 8;   // writable bytes required after the output

// Inflates the zlib stream in[0, n_in) to exactly n_out bytes and checks the Adler-32 trailer.
// `in` must be followed by INFLATE_IN_SLACK readable bytes, `out` must have room for
// n_out + INFLATE_OUT_SLACK bytes.  Returns false on any mismatch or malformed input (never reads or
// writes outside those bounds); the caller decides what to report.
bool inflate_zlib_fast(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out);

}  // namespace vo
