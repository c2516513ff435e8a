// TEST INFRASTRUCTURE ONLY — not product code.
//
// C entry points around the UNMODIFIED reference class
// (CaPS_SA::Suffix_Array<idx_t>, /root/reference/include/Suffix_Array.hpp:22-181),
// so that tests and bench.py's CPU baseline can call the reference in-process
// through ctypes.  Built only into oracle/_ref/libcaps_sa_ref.so by oracle/Makefile.
#include "Suffix_Array.hpp"

#include <chrono>
#include <cstdint>
#include <cstring>

namespace {

template <class idx_t>
double run_reference(const char* text, uint64_t n, uint64_t subproblems, uint64_t ctx,
                     idx_t* sa_out, idx_t* lcp_out) {
  CaPS_SA::Suffix_Array<idx_t> suf(text, static_cast<idx_t>(n), static_cast<idx_t>(subproblems),
                                   static_cast<idx_t>(ctx));
  const auto t0 = std::chrono::steady_clock::now();
  suf.construct();
  const auto t1 = std::chrono::steady_clock::now();
  std::memcpy(sa_out, suf.SA(), n * sizeof(idx_t));
  std::memcpy(lcp_out, suf.LCP(), n * sizeof(idx_t));
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

extern "C" {

// Returns construct() seconds (the region the reference itself times at
// src/Suffix_Array.cpp:466-494).  n must be >= 16 (reference SIGFPEs below that).
double caps_sa_ref_construct_u32(const char* text, uint64_t n, uint64_t subproblems, uint64_t ctx,
                                 uint32_t* sa_out, uint32_t* lcp_out) {
  return run_reference<uint32_t>(text, n, subproblems, ctx, sa_out, lcp_out);
}

double caps_sa_ref_construct_u64(const char* text, uint64_t n, uint64_t subproblems, uint64_t ctx,
                                 uint64_t* sa_out, uint64_t* lcp_out) {
  return run_reference<uint64_t>(text, n, subproblems, ctx, sa_out, lcp_out);
}

}  // extern "C"
