// Packed text in HBM: the staging layout every kernel reads.
//
// The input bytes are remapped to dense codes that preserve the reference's comparison
// order — `signed char` (reference src/Suffix_Array.cpp:77,289; include/Suffix_Array.hpp:29),
// i.e. bytes 0x80..0xFF sort before 0x00..0x7F — and packed MSB-first into 64-bit words,
// `bits` in {1,2,4,8} per symbol (2 for DNA, 8 for general bytes).  MSB-first means that
// comparing two 64-bit windows as unsigned integers compares 64/bits symbols
// lexicographically, and clz(a ^ b) / bits is their common-prefix length.  Two zero words
// follow the text so a window may start at any symbol of the text.
//
// End of text: windows are zero-padded.  Code 0 is the smallest symbol, so a padded window
// never sorts after the window of a longer suffix it is a prefix of; the remaining ties
// (a short suffix against a longer one that continues with code-0 symbols) are resolved by
// the rank-refinement rounds, which order "beyond the end" before every real rank
// (reference rule: the shorter suffix sorts first, src/Suffix_Array.cpp:71,76).
#pragma once

#include <cstdint>

namespace capsb {

struct PackedText {
  const uint64_t* words;  // ceil(n*bits/64) + 2 words, tail zero
  uint64_t n;             // symbols
  unsigned log2_bits;     // 0..3  (bits per symbol = 1 << log2_bits)

  __host__ __device__ unsigned bits() const { return 1u << log2_bits; }
  __host__ __device__ unsigned syms_per_word() const { return 64u >> log2_bits; }

#ifdef __CUDACC__
  // 64-bit window (64/bits symbols) starting at symbol i; i may be anywhere in [0, n].
  __device__ __forceinline__ uint64_t window(uint64_t i) const {
    const uint64_t bitpos = i << log2_bits;
    const uint64_t w = bitpos >> 6;
    const unsigned off = static_cast<unsigned>(bitpos & 63u);
    const uint64_t hi = __ldg(words + w);
    if (off == 0) return hi;
    const uint64_t lo = __ldg(words + w + 1);
    return (hi << off) | (lo >> (64u - off));
  }

  // Code of the symbol at position i (i < n).
  __device__ __forceinline__ unsigned symbol(uint64_t i) const {
    const uint64_t bitpos = i << log2_bits;
    const uint64_t w = __ldg(words + (bitpos >> 6));
    const unsigned b = bits();
    const unsigned shift = 64u - b - static_cast<unsigned>(bitpos & 63u);
    return static_cast<unsigned>((w >> shift) & ((1ull << b) - 1ull));
  }

  // Length of the common prefix of suffixes i and j, in symbols, given that the first
  // `known` symbols already agree; exact (bounded by the shorter suffix), at most `limit`
  // further windows are inspected.  Returns true when the answer is final.
  __device__ __forceinline__ bool common_prefix(uint64_t i, uint64_t j, uint64_t known,
                                                unsigned limit, uint64_t* out) const {
    const uint64_t shorter = n - (i > j ? i : j);
    const unsigned spw = syms_per_word();
    uint64_t l = known;
    for (unsigned step = 0; step < limit; ++step) {
      if (l >= shorter) {
        *out = shorter;
        return true;
      }
      const uint64_t x = window(i + l) ^ window(j + l);
      if (x != 0) {
        l += static_cast<uint64_t>(__clzll(static_cast<long long>(x))) >> log2_bits;
        *out = l < shorter ? l : shorter;
        return true;
      }
      l += spw;
    }
    if (l >= shorter) {
      *out = shorter;
      return true;
    }
    *out = l;
    return false;
  }
#endif
};

}  // namespace capsb
