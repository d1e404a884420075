// viterbi_pair_core.h -- the arithmetic of the two-frames-per-thread Viterbi kernel (viterbi_kernels.cu): branch
// metrics, one ACS step on the 64 packed path metrics, renormalisation, decision-bit layout, one traceback step.
// Host-compilable: on the device the packed 16-bit operations are single DPX instructions, on the host they are
// emulated half by half, so tests/host/viterbi_pair_check.cpp can run the very same step functions over whole
// frames and compare the decoded bits with the oracle in the CPU suite.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define VPC_HD __host__ __device__ __forceinline__
#else
#define VPC_HD inline
struct uint4 {
    unsigned int x, y, z, w;
};
inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
#endif

namespace fec {

namespace vpc {  // device intrinsics and their host emulations

// per 16-bit half: min(a + b, c), unsigned
VPC_HD uint32_t addmin_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __viaddmin_u16x2(a, b, c);
#else
    const uint32_t lo = (a + b) & 0xFFFFu, hi = ((a >> 16) + (b >> 16)) & 0xFFFFu;
    return (lo < (c & 0xFFFFu) ? lo : (c & 0xFFFFu)) | ((hi < (c >> 16) ? hi : (c >> 16)) << 16);
#endif
}

// per 16-bit half: max(max(a + b, c), 0), signed
VPC_HD uint32_t addmax_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2_relu(a, b, c);
#else
    auto half = [](uint32_t x, uint32_t y, uint32_t z) {
        int v = (int16_t)(uint16_t)(x + y), w = (int16_t)(uint16_t)z;
        v = v > w ? v : w;
        return (uint32_t)(v > 0 ? v : 0);
    };
    return half(a & 0xFFFFu, b & 0xFFFFu, c & 0xFFFFu) | (half(a >> 16, b >> 16, c >> 16) << 16);
#endif
}

VPC_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    const uint64_t v = ((uint64_t)y << 32) | x;  // selectors used here never set the sign-replicate bit
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}

VPC_HD uint32_t funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, shift);
#else
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << (shift & 31u)) >> 32);
#endif
}

VPC_HD uint32_t brev(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
#endif
}

}  // namespace vpc

namespace {

// ---- code structure -------------------------------------------------------------------------
// Encoder polynomials in the register orientation of viterbi-benchmark.cpp:64.
VPC_HD constexpr unsigned kPoly(int k) { return k == 0 ? 109u : k == 1 ? 79u : k == 2 ? 83u : 109u; }
VPC_HD constexpr unsigned parity8(unsigned v) {
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    return v & 1u;
}
// Expected code bit k on the branch old-state i -> new-state 2i (the bytes of const.asm:35-49).
VPC_HD constexpr unsigned tbit(int i, int k) { return parity8((2u * (unsigned)i) & kPoly(k)); }
// Branch-metric pattern of butterfly i: bits (T0,T1,T2); T3 == T0 because polys 0 and 3 coincide.
VPC_HD constexpr int pattern(int i) { return (int)(tbit(i, 0) | (tbit(i, 1) << 1) | (tbit(i, 2) << 2)); }

// Where the decision of new state s lives in a frame's 64-bit decision word {lo, hi}.  The traceback
// keeps the 6-bit state bit-reversed in the low bits of its history register h (newest decoded bit at
// bit 0), so r = brev6(s).  word = r >> 5 (= s & 1, known one step early), and the bit sits at 31 - (r & 31)
// so that `word << (h & 31)` brings it to bit 31, ready to be funnel-shifted into h.
VPC_HD constexpr unsigned brev6(unsigned s) {
    return ((s & 1u) << 5) | ((s & 2u) << 3) | ((s & 4u) << 1) | ((s & 8u) >> 1) | ((s & 16u) >> 3) | ((s & 32u) >> 5);
}
VPC_HD constexpr int dec_word(int s) { return (int)(brev6((unsigned)s) >> 5); }
VPC_HD constexpr uint32_t dec_bit(int s) { return 1u << (31u - (brev6((unsigned)s) & 31u)); }

constexpr uint32_t kSat = 0x0FF00FF0u;    // 255 * 16 per half: paddusb ceiling
constexpr uint32_t kM63 = 0x03F003F0u;    // 63 * 16 per half
constexpr uint32_t kEven = 0xFFFEFFFEu;

// ne = min(t, m0) per 16-bit half, and for each half whose minimum is t (ties included) add `bit` to
// that frame's decision word.  ptxas fuses the min + setp pattern into one VIMNMX.U16x2 with two
// predicate outputs (the same pattern __vibmin_u16x2 uses) and emits the adds as predicated VIADD,
// which does not occupy the ALU pipe the min / add-min instructions run on (measured:
// VIADD + VIADDMNMX pairs issue at 1.0 instruction/clk per SM sub-partition, profiles/intbench_r01b.jsonl;
// a predicated IMAD in the same place was 12-15 % slower end to end).
VPC_HD uint32_t min_decide(uint32_t t, uint32_t m0, uint32_t& decA, uint32_t& decB, uint32_t bit) {
#if defined(__CUDA_ARCH__)
    uint32_t ne;
    asm("{.reg .pred pu, pv; .reg .u16 rs0, rs1, rs2, rs3;\n\t"
        "min.u16x2 %0, %3, %4;\n\t"
        "mov.b32 {rs0, rs1}, %0;\n\t"
        "mov.b32 {rs2, rs3}, %3;\n\t"
        "setp.eq.u16 pv, rs0, rs2;\n\t"
        "setp.eq.u16 pu, rs1, rs3;\n\t"
        "@pv add.u32 %1, %1, %5;\n\t"
        "@pu add.u32 %2, %2, %5;}\n\t"
        : "=r"(ne), "+r"(decA), "+r"(decB)
        : "r"(t), "r"(m0), "r"(bit));
#else
    const uint32_t lo = (t & 0xFFFFu) < (m0 & 0xFFFFu) ? (t & 0xFFFFu) : (m0 & 0xFFFFu);
    const uint32_t hi = (t >> 16) < (m0 >> 16) ? (t >> 16) : (m0 >> 16);
    const uint32_t ne = lo | (hi << 16);
    if (lo == (t & 0xFFFFu)) decA += bit;  // ties included: the minimum is t
    if (hi == (t >> 16)) decB += bit;
#endif
    return ne;
}

// Branch metrics for two frames at once (deconvolve.cpp:334-349 restated):
//   x_k = y_k ^ T_k,  m = avg(avg(x0,x1), avg(x2,x3)) >> 2  with avg(a,b) = (a+b+1)>>1.
// With a = (x0+x1+1)>>1 and b = (x2+x3+1)>>1:  16*m = (2a + 2b + 2) & 0x3F0.
// x ^ 0xFF = 255 - x, so the four (T0,T1) cases of x0+x1+1 are linear in y0+y1 or y0-y1.
// wA / wB: the four soft symbols of one trellis step of frame A / frame B.
VPC_HD void branch_metrics(uint32_t wA, uint32_t wB, uint32_t (&bm)[8], uint32_t (&bmm)[8]) {
    const uint32_t p01 = vpc::byte_perm(wA, wB, 0x5140);  // A0 B0 A1 B1
    const uint32_t p23 = vpc::byte_perm(wA, wB, 0x7362);  // A2 B2 A3 B3
    const uint32_t y0 = vpc::byte_perm(p01, 0u, 0x4140), y1 = vpc::byte_perm(p01, 0u, 0x4342);
    const uint32_t y2 = vpc::byte_perm(p23, 0u, 0x4140), y3 = vpc::byte_perm(p23, 0u, 0x4342);
    const uint32_t u01 = y0 + y1, v01 = y0 - y1 + 0x01020102u;  // y0-y1+256 (+2 rounding term)
    const uint32_t u23 = y2 + y3, v23 = y2 - y3 + 0x01000100u;
    uint32_t e[2][2], f[2][2];  // e[T0][T1] = 2a+2, f[T2][T0] = 2b
    e[0][0] = (u01 + 0x00030003u) & kEven;
    e[1][1] = (0x02010201u - u01) & kEven;
    e[0][1] = v01 & kEven;
    e[1][0] = (0x02040204u - v01) & kEven;
    f[0][0] = (u23 + 0x00010001u) & kEven;
    f[1][1] = (0x01FF01FFu - u23) & kEven;
    f[0][1] = v23 & kEven;
    f[1][0] = (0x02000200u - v23) & kEven;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int t0 = p & 1, t1 = (p >> 1) & 1, t2 = (p >> 2) & 1;
        bm[p] = (e[t0][t1] + f[t2][t0]) & kM63;
        bmm[p] = kM63 - bm[p];
    }
}

// Renormalize256 (deconvolve.cpp:407-412): after every second step, if metric[state 0] > 150, 63 is
// subtracted with saturation at 0 from all 64 metrics -- decided per frame, i.e. per 16-bit half.
// Returns the per-half addend: -1008 (= -63 * 16) where the frame renormalises, else 0.
VPC_HD uint32_t renorm_addend(uint32_t m0) {
    // bit 15 of (m0 + 0x7FFF - 2400) is set iff m0 > 2400 (= 150 * 16)
    const uint32_t hit = ((m0 + 0x769F769Fu) >> 15) & 0x00010001u;
    return hit * 0xFC10u;
}

// One trellis step for two frames: old metrics M -> new metrics N, 4 decision words.
// Per butterfly i (old states i, i+32 -> new states 2i, 2i+1), with m = bm[pattern(i)]:
//   t1 = min(M[i+32] + (63-m), 255)        candidate through the upper branch, saturated
//   N[2i] = min(t1, M[i] + m)               == min(sat(M[i]+m), sat(M[i+32]+63-m)) because t1 <= 255
//   decision(2i) = (t1 <= M[i] + m)         == (N[2i] == t1): ties choose predecessor i+32
// and symmetrically for 2i+1 (deconvolve.cpp:352-359).
// kRenorm: the renormalisation that the reference applies after the previous (odd) step is folded
// into the operand fetch of this step: M' = relu(M + neg), relu(max(M + neg, neg)) == max(M-63*16, 0).
// Decision words: x,y = frame A {lo, hi}; z,w = frame B; the decision of new state s is bit dec_bit(s) of
// word dec_word(s) (a permutation of the reference's decision_t, viterbi.h:90-92, which is never exported).
//
// The file is compiled with ptxas -O1, which keeps this source order, so the loop is software
// pipelined by hand: the add / add-min stage of butterfly i+1 (2 add-mins + 2 adds) is issued between
// the two select stages of butterfly i (2 mins + 4 predicated adds), which keeps the ALU pipe and the
// pipe the plain adds use both fed and puts 4+ independent instructions between every producer and
// its consumer.
template <bool kRenorm>
VPC_HD uint4 acs_step(const uint32_t (&M)[64], uint32_t (&N)[64], uint32_t wA, uint32_t wB,
                                          uint32_t neg) {
    constexpr int kPat[32] = {pattern(0),  pattern(1),  pattern(2),  pattern(3),  pattern(4),  pattern(5),  pattern(6),
                              pattern(7),  pattern(8),  pattern(9),  pattern(10), pattern(11), pattern(12), pattern(13),
                              pattern(14), pattern(15), pattern(16), pattern(17), pattern(18), pattern(19), pattern(20),
                              pattern(21), pattern(22), pattern(23), pattern(24), pattern(25), pattern(26), pattern(27),
                              pattern(28), pattern(29), pattern(30), pattern(31)};
    uint32_t bm[8], bmm[8];
    branch_metrics(wA, wB, bm, bmm);
    uint32_t dA[2] = {0u, 0u}, dB[2] = {0u, 0u};
    uint32_t t1[2], t3[2], m0[2], m2[2];
    {
        uint32_t a = M[0], b = M[32];
        if (kRenorm) {
            a = vpc::addmax_s16x2_relu(a, neg, neg);
            b = vpc::addmax_s16x2_relu(b, neg, neg);
        }
        t1[0] = vpc::addmin_u16x2(b, bmm[kPat[0]], kSat);
        m0[0] = a + bm[kPat[0]];
        t3[0] = vpc::addmin_u16x2(b, bm[kPat[0]], kSat);
        m2[0] = a + bmm[kPat[0]];
    }
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const int c = i & 1, n = c ^ 1;
        uint32_t a = 0, b = 0;
        if (i + 1 < 32) {
            a = M[i + 1], b = M[i + 33];
            if (kRenorm) a = vpc::addmax_s16x2_relu(a, neg, neg);
        }
        N[2 * i] = min_decide(t1[c], m0[c], dA[dec_word(2 * i)], dB[dec_word(2 * i)], dec_bit(2 * i));
        if (i + 1 < 32) {
            if (kRenorm) b = vpc::addmax_s16x2_relu(b, neg, neg);
            t1[n] = vpc::addmin_u16x2(b, bmm[kPat[(i + 1) & 31]], kSat);
            m0[n] = a + bm[kPat[(i + 1) & 31]];
        }
        N[2 * i + 1] = min_decide(t3[c], m2[c], dA[dec_word(2 * i + 1)], dB[dec_word(2 * i + 1)], dec_bit(2 * i + 1));
        if (i + 1 < 32) {
            t3[n] = vpc::addmin_u16x2(b, bm[kPat[(i + 1) & 31]], kSat);
            m2[n] = a + bmm[kPat[(i + 1) & 31]];
        }
    }
    return make_uint4(dA[0], dA[1], dB[0], dB[1]);
}

// ChainBack (deconvolve.cpp:416-435) for the two frames of this thread.
// The reference keeps state << 2 in an 8-bit register (es = (es >> 1) | (k << 7), state = es >> 2).  Here
// the history register runs the other way, h = (h << 1) | k: its low six bits are the state bit-reversed,
// and after 32 steps it holds 32 decoded bits (bit j = time 32m + j).  With the decision layout of
// dec_word / dec_bit one step is two dependent funnel shifts per frame,
//     x = word << (h & 31)          the decision bit of the current state arrives at bit 31
//     h = (h << 1) | (x >> 31)
// and `word` (lo / hi = bit 5 of the bit-reversed state) is selected during the PREVIOUS step, off the
// critical path, because bit 5 of the next h is bit 4 of the current one.
struct TraceState {
    uint32_t hA = 0, hB = 0;  // history registers
    uint32_t sA = 0, sB = 0;  // decision word already selected for the step about to run
};

// w_next: the decision record of the step that runs after this one (time t - 1).
VPC_HD void trace_step(TraceState& st, const uint4& w_next) {
    const uint32_t xA = vpc::funnelshift_l(0u, st.sA, st.hA);
    const uint32_t xB = vpc::funnelshift_l(0u, st.sB, st.hB);
    st.sA = (st.hA & 16u) ? w_next.y : w_next.x;
    st.sB = (st.hB & 16u) ? w_next.w : w_next.z;
    st.hA = vpc::funnelshift_l(xA, st.hA, 1);
    st.hB = vpc::funnelshift_l(xB, st.hB, 1);
}

// 32 decoded bits (bit j = time 32m + j) -> 4 output bytes, MSB-first within each byte.
VPC_HD uint32_t trace_word(uint32_t h) { return vpc::byte_perm(vpc::brev(h), 0u, 0x0123); }

}  // namespace

}  // namespace fec
