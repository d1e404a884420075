// rs_decode.h -- the per-codeword RS(120,110) decoder of rs_kernels.cu, written against a small policy class so
// that the very same code runs in the kernel (tables in shared memory, warp votes) and on the host
// (tests/host/rs_decode_check.cpp: plain arrays, a "warp" of one), where the CPU suite compares it with the
// oracle on a million codewords.
//
// Policy P:  static uint32_t ato(uint32_t i)   alpha^(i mod 255), i < 768            (RS_ato_mod, dllmain.cpp:145-146)
//            static uint32_t iof(uint32_t v)   log_alpha v, iof(0) = 255             (RS_iof, dllmain.cpp:131-143)
//            static Row      lfsr(uint32_t c)  c * (g(x) - x^10), coefficients x^0..x^9 in the bytes of .x .y .z
//            static bool     any(unsigned mask, bool p)      does any lane of the warp have p
//            static unsigned max(unsigned mask, unsigned v)  largest v in the warp
#pragma once
#include <cstddef>
#include <cstdint>

#include "rs_chien_bitsliced.h"

namespace fec {
namespace rsdec {

constexpr int NN = 255, NROOTS = 10, PAD = 135, CW = 120, DATA = 110;

// the decoder itself stays an out-of-line function in the kernel (as it was before it moved here)
#if defined(__CUDACC__)
#define RSB_HD_CALL __host__ __device__ __noinline__
#else
#define RSB_HD_CALL inline
#endif

// rschecksf.cpp:50-52
RSB_HD uint32_t mod255(uint32_t x) { return (x * 0x1010102u) >> 24; }

// Chien search (rschecksf.cpp:296-320): find the roots alpha^i, i = 1..255, of lambda, in ascending i, stopping
// once deg(lambda) of them are found.  Bit-sliced, 32 positions at a time, no table lookups: see
// rs_chien_bitsliced.h.  D is the largest degree in the warp (warp-uniform, so the warp runs ONE instantiation
// instead of serialising one search per distinct degree); coefficients above a lane's own degree are zero and
// contribute nothing.  A degree-d polynomial has at most d roots, so running past a lane's own early-exit point
// cannot change its count.
#pragma nv_exec_check_disable
template <class P, int D>
RSB_HD int chien(const uint32_t (&lam_poly)[NROOTS + 1], uint32_t (&root)[NROOTS + 1], int deg,
                                     unsigned mask) {
    return rsbits::chien_bitsliced<D>(lam_poly, root, deg, [mask](bool need) { return P::any(mask, need); });
}

// Decode one codeword stored at col[k * stride], k = 0..119, in place.  Returns the number of
// roots found (= corrected symbols as the reference counts them), 0 for a clean word, -1 if
// uncorrectable.  Called by all lanes of `mask` together: control flow is kept warp-uniform (clean
// lanes ride along with all-zero syndromes, which Berlekamp-Massey turns into lambda = 1, degree 0,
// zero roots, return value 0 -- exactly the reference's early return).
#pragma nv_exec_check_disable
template <class P>
RSB_HD_CALL int rs_decode_column(uint8_t* col, uint32_t stride, unsigned mask) {
    // ---- remainder of cw(x) mod g(x); cw[0] is the highest-degree coefficient -----------------
    uint32_t r0 = 0, r1 = 0, r2 = 0;  // coefficients x^0..x^3 | x^4..x^7 | x^8,x^9
#pragma unroll 4
    for (int k = 0; k < CW; k++) {
        const uint32_t c = r2 >> 8;  // coefficient of x^9 moves to x^10 and is reduced away
        const auto row = P::lfsr(c);
        r2 = ((r2 << 8) & 0xFF00u) | (r1 >> 24);
        r1 = (r1 << 8) | (r0 >> 24);
        r0 = (r0 << 8) | col[(size_t)k * stride];
        r0 ^= row.x;
        r1 ^= row.y;
        r2 ^= row.z;
    }
    // all syndromes zero <=> remainder zero (rschecksf.cpp:224-230); skip the rest if the whole warp is clean
    if (!P::any(mask, (r0 | r1 | r2) != 0u)) return 0;

    // ---- syndromes S_i = rem(alpha^i), then index form (rschecksf.cpp:232-233) ------------------
    uint32_t syn[NROOTS];  // 32-bit holders: byte arrays make the compiler pack/unpack registers
    {
        uint32_t lg[NROOTS];
#pragma unroll
        for (int k = 0; k < NROOTS; k++) {
            const uint32_t v = ((k < 4 ? r0 : k < 8 ? r1 : r2) >> (8 * (k & 3))) & 0xFFu;
            lg[k] = P::iof(v);
        }
#pragma unroll
        for (int i = 0; i < NROOTS; i++) {
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < NROOTS; k++)
                if (lg[k] != NN) acc ^= P::ato(lg[k] + i * k);  // <= 254 + 81
            syn[i] = P::iof(acc);
        }
    }

    // ---- Berlekamp-Massey (rschecksf.cpp:236-284): lambda polynomial form, b / syn index form ---
    uint32_t lam[NROOTS + 1], b[NROOTS + 1], nxt[NROOTS + 1];
#pragma unroll
    for (int i = 0; i <= NROOTS; i++) {
        lam[i] = (i == 0) ? 1 : 0;
        b[i] = (i == 0) ? 0 : NN;
    }
    int el = 0;
#pragma unroll
    for (int r = 1; r <= NROOTS; r++) {
        uint32_t discr = 0;
#pragma unroll
        for (int i = 0; i < r; i++)
            if (lam[i] != 0 && syn[r - i - 1] != NN) discr ^= P::ato(P::iof(lam[i]) + syn[r - i - 1]);
        discr = P::iof(discr);
        if (discr == NN) {
#pragma unroll
            for (int i = NROOTS; i > 0; i--) b[i] = b[i - 1];
            b[0] = NN;
        } else {
            nxt[0] = lam[0];
#pragma unroll
            for (int i = 0; i < NROOTS; i++) {
                nxt[i + 1] = lam[i + 1];
                if (b[i] != NN) nxt[i + 1] ^= P::ato(discr + b[i]);
            }
            if (2 * el <= r - 1) {
                el = r - el;
#pragma unroll
                for (int i = 0; i <= NROOTS; i++)
                    b[i] = (lam[i] == 0) ? (uint32_t)NN : mod255(P::iof(lam[i]) - discr + NN);
            } else {
#pragma unroll
                for (int i = NROOTS; i > 0; i--) b[i] = b[i - 1];
                b[0] = NN;
            }
#pragma unroll
            for (int i = 0; i <= NROOTS; i++) lam[i] = nxt[i];
        }
    }

    uint32_t lam_poly[NROOTS + 1];
    int deg_lambda = 0;
#pragma unroll
    for (int i = 0; i <= NROOTS; i++) {
        lam_poly[i] = lam[i];
        lam[i] = P::iof(lam[i]);
        if (lam[i] != NN) deg_lambda = i;
    }

    // ---- Chien search, one instantiation per warp (largest degree present) ---------------------------
    uint32_t root[NROOTS + 1];
    int count = 0;
    switch (P::max(mask, (unsigned)deg_lambda)) {
        case 1: count = chien<P, 1>(lam_poly, root, deg_lambda, mask); break;
        case 2: count = chien<P, 2>(lam_poly, root, deg_lambda, mask); break;
        case 3: count = chien<P, 3>(lam_poly, root, deg_lambda, mask); break;
        case 4: count = chien<P, 4>(lam_poly, root, deg_lambda, mask); break;
        case 5: count = chien<P, 5>(lam_poly, root, deg_lambda, mask); break;
        case 6: count = chien<P, 6>(lam_poly, root, deg_lambda, mask); break;
        case 7: count = chien<P, 7>(lam_poly, root, deg_lambda, mask); break;
        case 8: count = chien<P, 8>(lam_poly, root, deg_lambda, mask); break;
        case 9: count = chien<P, 9>(lam_poly, root, deg_lambda, mask); break;
        case 10: count = chien<P, 10>(lam_poly, root, deg_lambda, mask); break;
        default: break;  // every lane has degree 0: nothing to search
    }
    if (deg_lambda != count) return -1;  // rschecksf.cpp:325-326

    // ---- omega(x) = syn(x) lambda(x) mod x^10, index form (rschecksf.cpp:331-341) ----------------
    const int deg_omega = deg_lambda - 1;
    uint32_t om[NROOTS];
#pragma unroll
    for (int i = 0; i < NROOTS; i++) {
        uint32_t tmp = 0;
#pragma unroll
        for (int j = 0; j <= i; j++)
            if (syn[i - j] != NN && lam[j] != NN) tmp ^= P::ato(syn[i - j] + lam[j]);
        om[i] = (i <= deg_omega) ? P::iof(tmp) : (uint32_t)NN;
    }

    // ---- Forney (rschecksf.cpp:346-374) ---------------------------------------------------------
#pragma unroll
    for (int c = NROOTS - 1; c >= 0; c--) {
        if (c >= count) continue;
        const uint32_t rj = root[c];
        if (rj < PAD + 1) continue;  // root in the virtual padding: counted, not applied
        uint32_t num1 = 0;
#pragma unroll
        for (int i = NROOTS - 1; i >= 0; i--)
            if (i <= deg_omega && om[i] != NN) num1 ^= P::ato(mod255(om[i] + (uint32_t)i * rj));
        if (!num1) continue;
        const uint32_t num2 = P::ato(NN - rj);
        uint32_t den = 0;
        const int top = (deg_lambda < NROOTS - 1 ? deg_lambda : NROOTS - 1) & ~1;
#pragma unroll
        for (int i = 8; i >= 0; i -= 2)
            if (i <= top && lam[i + 1] != NN) den ^= P::ato(mod255(lam[i + 1] + (uint32_t)i * rj));
        // exponent used unreduced: the table has 768 entries (viterbi.h:101, rschecksf.cpp:366-370)
        col[(size_t)(rj - 1 - PAD) * stride] ^= P::ato(P::iof(num1) + P::iof(num2) + (NN - P::iof(den)));
    }
    return count;
}


}  // namespace rsdec
}  // namespace fec
