// CaPS_SA::Suffix_Array<idx_t> — B200-native drop-in for the reference's construction class.
//
// Same public surface as the reference class (reference include/Suffix_Array.hpp:22-181):
//   Suffix_Array(const char* T, idx_t n, idx_t subproblem_count = 0, idx_t max_context = 0)  (:155)
//   T(), n(), SA(), LCP()                                                                    (:165-174)
//   construct()                                                                              (:177)
//   dump(std::ofstream&)                                                                     (:180)
//   non-copyable, non-movable                                                                (:157-160)
// with explicit instantiations for uint32_t and uint64_t (reference src/Suffix_Array.cpp:543-544).
//
// What differs is everything behind construct(): the ParlayLib/AVX2 samplesort is replaced by
// hand-written sm_100a CUDA kernels reached through the C-ABI in caps_sa_gpu.h.  The text is
// borrowed; SA()/LCP() are host arrays (pinned) owned by the object and valid until it is
// destroyed.  Fatal conditions print to std::cerr and exit(EXIT_FAILURE), as in the reference
// (src/Suffix_Array.cpp:33-37); there is no CPU fallback.
#ifndef CAPS_SA_B200_SUFFIX_ARRAY_HPP
#define CAPS_SA_B200_SUFFIX_ARRAY_HPP

#include <cstddef>
#include <cstdint>
#include <fstream>

namespace CaPS_SA
{

template <typename T_idx_>
class Suffix_Array
{
public:

    typedef T_idx_ idx_t;

    // `T` must outlive the object.  `subproblem_count` is accepted for compatibility and used
    // only as a tuning hint (the output never depended on it); with a bounded `max_context` the
    // suffix array stays exact and LCP entries are capped at `max_context` (caps_sa_gpu.h).
    Suffix_Array(const char* T, idx_t n, idx_t subproblem_count = 0, idx_t max_context = 0);

    Suffix_Array(const Suffix_Array&) = delete;
    Suffix_Array& operator=(const Suffix_Array&) = delete;
    Suffix_Array(Suffix_Array&&) = delete;
    Suffix_Array& operator=(Suffix_Array&&) = delete;

    ~Suffix_Array();

    const char* T() const { return T_; }

    idx_t n() const { return n_; }

    const idx_t* SA() const { return SA_; }

    const idx_t* LCP() const { return LCP_; }

    // Builds SA and LCP on the GPU; results are in host memory when it returns.
    void construct();

    // Writes `size_t n || SA[n] || LCP[n]` (reference src/Suffix_Array.cpp:497-509).
    void dump(std::ofstream& output);

    // Not in the reference: the same bytes written to `path` by several threads with pwrite (the
    // 24.8 GB dump of a 3.1 Gbp text through one ofstream is the slowest step of the CLI).
    // Returns false if the file cannot be written.
    bool dump(const char* path);

private:

    const char* const T_;
    const idx_t n_;
    idx_t* const SA_;
    idx_t* const LCP_;
    const idx_t subproblem_hint_;
    const idx_t max_context_;
    bool constructed_;
};

}

#endif
