// rs_chien_bitsliced.h -- Chien search of the RS(120,110) decoder without table lookups.
//
// The reference evaluates lambda(alpha^i) for i = 1..255 one position at a time (rschecksf.cpp:296-320).  The
// table-driven kernel version did the same with one shared-memory lookup per term and position, and ncu showed
// the kernel bound by exactly those lookups (LSU wavefronts at 84 % of peak, half of them the Chien search).
// Here 32 positions are evaluated at once: a "slice" is 32 field elements stored as 8 bit planes (bit p of
// plane b = bit b of the element at position p), so
//   * the term lambda_j * alpha^(j i) for the positions i = 1..32 is an XOR of constant slices selected by
//     the bits of lambda_j (kChienInit),
//   * moving a term 32 positions ahead multiplies it by the constant alpha^(32 j), which is a fixed XOR
//     network on the 8 planes (kChienAdvance),
//   * the sum of the terms is a plane-wise XOR and "sum == 0" is a NOR of the 8 planes: one 32-bit word whose
//     set bits are the roots of the block, in ascending position order.
// Every table entry is used as a template constant, so the code is pure LOP3 arithmetic with immediates.
// The same field elements are computed as in the reference, only 32 at a time; roots come out in the same
// (ascending i) order.  Host-compilable: tests/test_rs_bitslice.py checks it against a plain Chien search.
#pragma once
#include <cstdint>
#include <utility>

#include "rs_bitslice_tables.h"

#if defined(__CUDACC__)
#define RSB_HD __host__ __device__ __forceinline__
#else
#define RSB_HD inline
#endif

namespace fec {
namespace rsbits {

using Seq8 = std::make_integer_sequence<int, 8>;

// the table entries as scalar compile-time constants (arrays of the host side cannot be read in device code,
// scalar constexpr values can)
template <int J, int BO, int BI>
constexpr bool kAdvanceBit = ((kChienAdvance[J - 1][BO] >> BI) & 1) != 0;
template <int J, int B, int PL>
constexpr uint32_t kInitWord = kChienInit[J - 1][B][PL];

// output plane BO of (slice * alpha^(32 J)): XOR of the input planes named by the mask
template <int J, int BO, int... BI>
RSB_HD uint32_t advance_plane(const uint32_t (&t)[8], std::integer_sequence<int, BI...>) {
    uint32_t acc = 0;
    ((acc ^= (kAdvanceBit<J, BO, BI> ? t[BI] : 0u)), ...);
    return acc;
}

template <int J, int... BO>
RSB_HD void advance_term(uint32_t (&t)[8], std::integer_sequence<int, BO...>) {
    const uint32_t n[8] = {advance_plane<J, BO>(t, Seq8{})...};
    ((t[BO] = n[BO]), ...);
}

template <int J, int B, int... PL>
RSB_HD void init_bit(uint32_t (&t)[8], uint32_t m, std::integer_sequence<int, PL...>) {
    ((t[PL] ^= m & kInitWord<J, B, PL>), ...);
}

// slice of coeff * alpha^(J i), i = 1..32
template <int J, int... B>
RSB_HD void init_term(uint32_t (&t)[8], uint32_t coeff, std::integer_sequence<int, B...>) {
    ((t[B] = 0u), ...);
    (init_bit<J, B>(t, 0u - ((coeff >> B) & 1u), Seq8{}), ...);
}

template <int D>
struct ChienSlices {
    uint32_t t[D][8];

    template <int... J>
    RSB_HD void init(const uint32_t* lam_poly, std::integer_sequence<int, J...>) {
        (init_term<J + 1>(t[J], lam_poly[J + 1], Seq8{}), ...);
    }
    template <int... J>
    RSB_HD void advance(std::integer_sequence<int, J...>) {
        (advance_term<J + 1>(t[J], Seq8{}), ...);
    }
    // bit p set <=> 1 + sum_j term_j == 0 at position p of the current block
    RSB_HD uint32_t zeros() const {
        uint32_t nz = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int pl = 0; pl < 8; pl++) {
            uint32_t x = (pl == 0) ? 0xFFFFFFFFu : 0u;  // the constant term 1
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 0; j < D; j++) x ^= t[j][pl];
            nz |= x;
        }
        return ~nz;
    }
};

// lam_poly[0..D]: lambda in polynomial form (lam_poly[0] == 1), coefficients above the lane's own degree zero.
// Records up to `deg` roots (values i in 1..255, ascending) in root[] and returns their number.
// `more(need)` is the early-exit vote: it returns whether any lane of the warp still needs roots (on the host:
// the identity).
template <int D, class Vote>
RSB_HD int chien_bitsliced(const uint32_t* lam_poly, uint32_t (&root)[11], int deg, Vote more) {
    ChienSlices<D> s;
    s.init(lam_poly, std::make_integer_sequence<int, D>{});
    int count = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int blk = 0; blk < 8; blk++) {
        uint32_t hits = s.zeros();
        if (blk == 7) hits &= 0x7FFFFFFFu;  // position 256 does not exist
        while (hits) {                      // a degree-d polynomial has at most d roots
#if defined(__CUDA_ARCH__)
            const uint32_t p = (uint32_t)__ffs((int)hits) - 1u;
#else
            const uint32_t p = (uint32_t)__builtin_ctz(hits);
#endif
            hits &= hits - 1;
            if (count < deg) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int c = 0; c < 10; c++)  // root[] lives in registers: no dynamic indexing
                    if (c == count) root[c] = (uint32_t)(32 * blk + 1) + p;
                count++;
            }
        }
        if (blk == 7 || !more(count < deg)) break;
        s.advance(std::make_integer_sequence<int, D>{});
    }
    return count;
}

}  // namespace rsbits
}  // namespace fec
