// Host-side rANS entropy coder of libstemb200 (C ABI): 64-bit state, 32-bit renormalisation words written backwards,
// 16-bit probability precision, per-symbol CDF selection by index, and 4-bit "bypass" escape coding for values
// outside a CDF's support. Byte-compatible with the reference's pybind11 module compressai.ans
// (compressai/cpp_exts/rans/rans_interface.cpp:99-275 over third_party/ryg_rans/rans64.h:59-135); restated here
// over flat int32 arrays so that symbols and indexes can come straight from pinned copies of the GPU kernels'
// outputs instead of Python lists.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/stemb200.h"
#include "internal.h"

namespace {

constexpr uint32_t kPrecision = 16;       // probability resolution
constexpr uint32_t kBypassBits = 4;       // escape nibbles
constexpr uint32_t kBypassMax = (1u << kBypassBits) - 1;
constexpr uint64_t kLowerBound = 1ull << 31;  // normalisation interval [L, L * 2^32)

struct Sym {
  uint16_t start, range;
  uint8_t bypass;
};

inline void put_word_if_needed(uint64_t& x, uint64_t x_max, uint32_t*& ptr) {
  if (x >= x_max) {
    *--ptr = static_cast<uint32_t>(x);
    x >>= 32;
  }
}

}  // namespace

extern "C" int64_t stemb200_rans_encode_host(const int32_t* symbols, const int32_t* indexes, int64_t n,
                                             const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride,
                                             const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out,
                                             int64_t out_capacity) {
  if (!symbols || !indexes || !cdfs || !cdf_sizes || !offsets || !out || n < 0 || n_cdfs < 1 || cdf_stride < 2)
    return stem::set_error("rans_encode: bad argument");
  std::vector<Sym> syms;
  syms.reserve(static_cast<size_t>(n) + 16);
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return stem::set_error("rans_encode: CDF index out of range");
    const int32_t* cdf = cdfs + static_cast<int64_t>(ci) * cdf_stride;
    const int32_t max_value = cdf_sizes[ci] - 2;
    if (max_value < 0 || max_value + 1 >= cdf_stride) return stem::set_error("rans_encode: bad CDF size");
    int32_t value = symbols[i] - offsets[ci];
    uint32_t raw = 0;
    if (value < 0) {
      raw = static_cast<uint32_t>(-2 * value - 1);
      value = max_value;
    } else if (value >= max_value) {
      raw = static_cast<uint32_t>(2 * (value - max_value));
      value = max_value;
    }
    syms.push_back({static_cast<uint16_t>(cdf[value]), static_cast<uint16_t>(cdf[value + 1] - cdf[value]), 0});
    if (value == max_value) {
      // escape: number of nibbles (unary in chunks of 15), then the nibbles, least significant first
      int32_t n_nib = 0;
      while ((raw >> (n_nib * kBypassBits)) != 0) ++n_nib;
      int32_t v = n_nib;
      while (v >= static_cast<int32_t>(kBypassMax)) {
        syms.push_back({static_cast<uint16_t>(kBypassMax), static_cast<uint16_t>(kBypassMax + 1), 1});
        v -= kBypassMax;
      }
      syms.push_back({static_cast<uint16_t>(v), static_cast<uint16_t>(v + 1), 1});
      for (int32_t j = 0; j < n_nib; ++j) {
        const uint32_t nib = (raw >> (j * kBypassBits)) & kBypassMax;
        syms.push_back({static_cast<uint16_t>(nib), static_cast<uint16_t>(nib + 1), 1});
      }
    }
  }
  // rANS is last-in first-out: encode from the last symbol to the first, writing words backwards
  std::vector<uint32_t> buf(syms.size() + 4);
  uint32_t* const end = buf.data() + buf.size();
  uint32_t* ptr = end;
  uint64_t x = kLowerBound;
  for (size_t k = syms.size(); k-- > 0;) {
    const Sym s = syms[k];
    if (!s.bypass) {
      const uint64_t x_max = ((kLowerBound >> kPrecision) << 32) * s.range;
      put_word_if_needed(x, x_max, ptr);
      x = ((x / s.range) << kPrecision) + (x % s.range) + s.start;
    } else {
      const uint64_t x_max = ((kLowerBound >> 16) << 32) * (1u << (16 - kBypassBits));
      put_word_if_needed(x, x_max, ptr);
      x = (x << kBypassBits) | s.start;
    }
  }
  ptr -= 2;
  ptr[0] = static_cast<uint32_t>(x);
  ptr[1] = static_cast<uint32_t>(x >> 32);
  const int64_t nbytes = static_cast<int64_t>(end - ptr) * 4;
  if (nbytes > out_capacity) return stem::set_error("rans_encode: output buffer too small");
  memcpy(out, ptr, static_cast<size_t>(nbytes));
  return nbytes;
}

extern "C" int stemb200_rans_decode_host(const uint8_t* stream, int64_t nbytes, const int32_t* indexes, int64_t n,
                                         const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride,
                                         const int32_t* cdf_sizes, const int32_t* offsets, int32_t* symbols_out) {
  if (!stream || !indexes || !cdfs || !cdf_sizes || !offsets || !symbols_out || n < 0 || nbytes < 8 || (nbytes & 3))
    return stem::set_error("rans_decode: bad argument");
  std::vector<uint32_t> words(static_cast<size_t>(nbytes / 4) + 2, 0u);  // aligned copy (+2 guard words)
  memcpy(words.data(), stream, static_cast<size_t>(nbytes));
  const uint32_t* ptr = words.data();
  const uint32_t* const last = words.data() + words.size();
  uint64_t x = static_cast<uint64_t>(ptr[0]) | (static_cast<uint64_t>(ptr[1]) << 32);
  ptr += 2;
  auto get_bits = [&](uint32_t nb) -> uint32_t {
    const uint32_t val = static_cast<uint32_t>(x & ((1u << nb) - 1));
    x >>= nb;
    if (x < kLowerBound && ptr < last) x = (x << 32) | *ptr++;
    return val;
  };
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return stem::set_error("rans_decode: CDF index out of range");
    const int32_t* cdf = cdfs + static_cast<int64_t>(ci) * cdf_stride;
    const int32_t size = cdf_sizes[ci];
    const int32_t max_value = size - 2;
    if (max_value < 0 || size > cdf_stride) return stem::set_error("rans_decode: bad CDF size");
    const uint32_t cum = static_cast<uint32_t>(x & ((1u << kPrecision) - 1));
    // first entry greater than cum, minus one (the table is increasing: binary search)
    int32_t lo = 0, hi = size;  // search in [0, size)
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (static_cast<uint32_t>(cdf[mid]) > cum) hi = mid;
      else lo = mid + 1;
    }
    const int32_t s = lo - 1;
    if (s < 0 || s + 1 >= size) return stem::set_error("rans_decode: corrupt stream");
    const uint32_t start = static_cast<uint32_t>(cdf[s]), freq = static_cast<uint32_t>(cdf[s + 1] - cdf[s]);
    x = freq * (x >> kPrecision) + (x & ((1ull << kPrecision) - 1)) - start;
    if (x < kLowerBound && ptr < last) x = (x << 32) | *ptr++;
    int32_t value = s;
    if (value == max_value) {
      int32_t v = static_cast<int32_t>(get_bits(kBypassBits));
      int32_t n_nib = v;
      while (v == static_cast<int32_t>(kBypassMax)) {
        v = static_cast<int32_t>(get_bits(kBypassBits));
        n_nib += v;
      }
      uint32_t raw = 0;
      for (int32_t j = 0; j < n_nib; ++j) raw |= get_bits(kBypassBits) << (j * kBypassBits);
      value = static_cast<int32_t>(raw >> 1);
      if (raw & 1) value = -value - 1;
      else value += max_value;
    }
    symbols_out[i] = value + offsets[ci];
  }
  return 0;
}
