/* TEST INFRASTRUCTURE ONLY — not product code.
 *
 * CPU oracle for the CaPS-SA construction path.  Two independent pieces:
 *   caps_port.c : plain-C restatement of the reference's samplesort-with-LCP-merge
 *                 (each function cites the reference file:line it follows);
 *   sa_check.c  : algorithm-independent validator (suffix order via ISA, LCP via Kasai).
 * PARITY PINNING: the restatement is checked in tests/test_oracle.py against
 *   (1) the reference's only runnable fixture, data/simpletest2 (SURVEY.md Appendix A1,
 *       committed under tests/golden/), and
 *   (2) outputs of the unmodified reference compiled here (oracle/_ref, see Makefile)
 *       on seeded random / periodic / Fibonacci / byte-alphabet inputs (tests/golden/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
 */
#ifndef CAPS_ORACLE_PORT_H
#define CAPS_ORACLE_PORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Restatement of Suffix_Array<idx_t>::construct() (reference src/Suffix_Array.cpp:466-494).
 * text: n bytes compared as signed char; subproblems/max_context: 0 = reference defaults.
 * Outputs widened to 64-bit.  Returns 0, or -1 if n < 16 (the reference divides by zero
 * there, src/Suffix_Array.cpp:24,27) or on allocation failure. */
int caps_port_construct(const char* text, uint64_t n, uint64_t subproblems, uint64_t max_context,
                        uint64_t* sa_out, uint64_t* lcp_out);

/* Same, narrowing the outputs (n must fit). */
int caps_port_construct_u32(const char* text, uint64_t n, uint64_t subproblems,
                            uint64_t max_context, uint32_t* sa_out, uint32_t* lcp_out);

/* Byte mapping of the CLI (reference src/main.cpp:61-70), in place. */
void caps_port_map_acgt(char* text, uint64_t n);

/* Independent validator.  idx_bytes is 4 or 8.  Returns 0 when sa is the suffix array of
 * text under signed-char order with shorter-suffix-first and lcp is its LCP array
 * (lcp[0] = 0); otherwise a positive code (1 = not a permutation, 2 = order violated,
 * 3 = LCP mismatch, 4 = bad args / out of memory) and, if bad_pos != NULL, the first
 * offending SA position. */
int caps_check_sa_lcp(const char* text, uint64_t n, const void* sa, const void* lcp,
                      int idx_bytes, uint64_t* bad_pos);

/* Number of OpenMP threads the multi-threaded checkers use from now on (launchers like torchrun
 * export OMP_NUM_THREADS=1). */
void caps_oracle_set_threads(int threads);

/* caps_check_sa_lcp with OpenMP loops (Kasai's walk restarted per range of text positions) and
 * an inverse permutation at the input's index width: the form bench.py applies to the 3.1 G
 * suffix results.  Same codes. */
int caps_check_sa_lcp_mt(const char* text, uint64_t n, const void* sa, const void* lcp, int idx_bytes,
                         uint64_t* bad_pos);
/* ... with the number of ranges Kasai's walk is cut into bounded (highly repetitive texts: one per thread). */
int caps_check_sa_lcp_mt_pieces(const char* text, uint64_t n, const void* sa, const void* lcp, int idx_bytes,
                                uint64_t max_pieces, uint64_t* bad_pos);

/* The same validation for a text that is its first `period` (<= 4096) bytes repeated, with OpenMP
 * loops and the LCP from the closed form for periodic texts (sa_check.c) — for the 1 Gbp
 * periodic text of BASELINE config 4, where the sequential walk above takes minutes.
 * Extra code: 5 = the text is not period-periodic. */
int caps_check_sa_lcp_periodic(const char* text, uint64_t n, uint64_t period, const void* sa, const void* lcp,
                               int idx_bytes, uint64_t* bad_pos);

/* Naive O(n^2 log n) SA + direct LCP for tiny inputs (third, trivially-correct oracle;
 * same role as the reference's chatgpt_baseline.py:5-28). */
int caps_naive_sa_lcp(const char* text, uint64_t n, uint64_t* sa_out, uint64_t* lcp_out);

#ifdef __cplusplus
}
#endif
#endif
