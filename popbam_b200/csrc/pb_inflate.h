// pb_inflate.h -- raw DEFLATE decoder for BGZF blocks (see pb_inflate.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

namespace pbio {
// Inflates one complete raw DEFLATE stream of known output size.  Returns false on any malformed input or when
// the stream does not produce exactly out_len bytes.  Thread-safe (thread-local tables).
bool inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);
}  // namespace pbio
