// viterbi_kernels.cu -- sm_100a kernels for the DAB mother-code Viterbi decoder.
//
// Replaces the hot loops of the reference: Butterfly256 / Renormalize256 / ChainBack
// (deconvolve.cpp:334-387, 407-412, 416-435) for whole batches of frames.  Bit-exact with
// the reference's 8-bit saturating arithmetic; the layout is redesigned for the B200 integer
// pipes instead of 256-bit CPU vectors:
//
//   * one THREAD decodes TWO frames.  A 32-bit register holds the same trellis state of frame A
//     (low half) and frame B (high half) as unsigned 16-bit lanes, so every packed min / add-min
//     (VIADDMNMX.U16x2, full rate on the ALU pipe -- profiles/intbench_r01.jsonl) advances two
//     frames.  64 registers hold the 64 path metrics; the butterfly network is pure register
//     renaming (no shuffles, no shared memory, no permutes).
//   * path metrics are kept scaled by 16 (value*16 fits 16 bits: 255*16 = 4080).  All reference
//     operations (saturating add at 255, saturating subtract of 63, threshold 150) are linear
//     in the metric, so scaling is exact, and it lets the branch metric be formed as
//     (e + f) & 0x03F0 without a shift.
//   * decisions: 64 bits per step per frame.  The survivor select is a packed min that also
//     returns one predicate per frame (VIMNMX.U16x2 with predicate outputs); each predicate adds
//     its decision bit to a per-frame 64-bit word with a predicated VIADD, which issues beside the
//     ALU-pipe min/add-min work instead of competing with it.  The words are streamed to a per-warp
//     scratch area in global memory, fully coalesced (512 B per warp per step).
//     Shared memory cannot hold them: 8 B x 3078 steps = 24.6 KB per frame would cap an SM at
//     9 frames.
//   * traceback runs in the same kernel, by the same thread, reading its own scratch back through a
//     32-step ring of registers (the registers the path metrics no longer need), so that the loads of
//     28 steps are in flight while the serial state recursion runs.  The position of each decision bit
//     inside the 64-bit word is free (it is a compile-time constant of the ACS code), so it is chosen
//     to make that recursion two dependent funnel shifts per step (see dec_word / dec_bit).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "fec_internal.h"
#include "viterbi_pair_core.h"

namespace fec {

namespace {


#ifndef VIT_SYM_PREFETCH
#define VIT_SYM_PREFETCH 1  // 0: none, 1: L1 prefetch instruction, 2: second register stage
#endif
#ifndef VIT_SYM_PREFETCH_PAIRS
#define VIT_SYM_PREFETCH_PAIRS 6
#endif
#ifndef VIT_TRACE_PF_BLOCKS
#define VIT_TRACE_PF_BLOCKS 4
#endif
constexpr uint32_t kSymPrefetchPairs = VIT_SYM_PREFETCH_PAIRS;  // ACS loop: L1 prefetch distance for the symbol rows (8 B pairs)
constexpr int kTraceRing = 32;      // steps held in registers (one output word)
constexpr int kTraceSub = 4;        // steps per reload group
constexpr int kTracePrefetchBlocks = VIT_TRACE_PF_BLOCKS;  // 32-step blocks ahead that are pulled into L2

// Pull the 16 KB decision block starting at `blk` (warp base, lane 0) into L2: 128 lines, 4 per lane.
__device__ __forceinline__ void trace_prefetch_l2(const uint4* blk, uint32_t lane) {
    const char* q = reinterpret_cast<const char*>(blk) + lane * 128u;
#pragma unroll
    for (int k = 0; k < 4; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + k * 4096));
}

// The state recursion is serial, but the decision records it consumes do not depend on it.  They were
// written by this same thread and have mostly been evicted to HBM by now, so they come back in two
// stages: an L2 prefetch kTracePrefetchBlocks blocks ahead, and a ring of kTraceRing records in registers
// (the registers the path metrics no longer need): as soon as a group of kTraceSub records has been
// consumed it is reloaded with the records of the next block, 28 steps before they are needed.
// (ptxas puts all ring loads on one scoreboard and waits for it once per block, so the window does not roll
// across block boundaries.  A cp.async ring in shared memory, whose counted wait does roll, was measured:
// FIC 112.8 vs 112.1 Gbit/s, but MSC 120 vs 141 because 192 KB of rings per SM leave the symbol loads no L1.)
template <bool kWordStores>
__device__ __forceinline__ void traceback(const uint4* __restrict__ dec, uint32_t lane, uint32_t framebits,
                                          uint8_t* outA, uint8_t* outB, bool liveA, bool liveB) {
    constexpr int kBlk = kTraceRing * 32;  // uint4 elements per 32-step block of one warp
    TraceState st;
    const int nblk = (int)(framebits / kTraceRing);
    int t = (int)framebits - 1;
    // record of time 32 (nblk - 1), this lane
    const uint4* p = dec + (size_t)((nblk > 0 ? nblk - 1 : 0) * kTraceRing + 6) * 32;
    for (int d = 1; d < kTracePrefetchBlocks; d++)
        if (nblk - 1 - d >= 0) trace_prefetch_l2(p - lane - (size_t)d * kBlk, lane);
    uint4 R[kTraceRing];
    if (nblk > 0) {
#pragma unroll
        for (int j = 0; j < kTraceRing; j++) R[j] = p[j * 32];
    } else {  // defined on every path: an undefined ring would be live across the whole kernel for ptxas
#pragma unroll
        for (int j = 0; j < kTraceRing; j++) R[j] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (kWordStores) {  // start state 0: word lo of the last record
        st.sA = R[kTraceRing - 1].x, st.sB = R[kTraceRing - 1].z;
    } else {
        const uint4 w = dec[(size_t)(t + 6) * 32];
        st.sA = w.x, st.sB = w.z;
    }
    if (!kWordStores) {  // ragged top: framebits % 32 steps, then the partial output word byte by byte
        const int head = (int)(framebits % kTraceRing);
        for (int i = 0; i < head; i++, t--) trace_step(st, dec[(size_t)(t > 0 ? t + 5 : 6) * 32]);
        const uint32_t vA = __brev(st.hA), vB = __brev(st.hB);
        for (int b = 0; b < (head + 7) / 8; b++) {
            if (liveA) outA[nblk * 4 + b] = (uint8_t)(vA >> (24 - 8 * b));
            if (liveB) outB[nblk * 4 + b] = (uint8_t)(vB >> (24 - 8 * b));
        }
    }
    for (int m = nblk - 1; m >= 0; m--) {
        const uint4* pn = m > 0 ? p - kBlk : p;  // last block: reload in place (values unused)
        if (m >= kTracePrefetchBlocks) trace_prefetch_l2(p - lane - (size_t)kTracePrefetchBlocks * kBlk, lane);
#pragma unroll
        for (int j = kTraceRing - 1; j >= 0; j--) {
            trace_step(st, R[j > 0 ? j - 1 : kTraceRing - 1]);  // at j == 0 R[31] already holds the next block
            if (j % kTraceSub == 0) {
#pragma unroll
                for (int k = 0; k < kTraceSub; k++) R[j + k] = pn[(j + k) * 32];
            }
        }
        if (kWordStores) {
            if (liveA) *reinterpret_cast<uint32_t*>(outA + 4 * m) = trace_word(st.hA);
            if (liveB) *reinterpret_cast<uint32_t*>(outB + 4 * m) = trace_word(st.hB);
        } else {
            const uint32_t vA = __brev(st.hA), vB = __brev(st.hB);
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (liveA) outA[4 * m + b] = (uint8_t)(vA >> (24 - 8 * b));
                if (liveB) outB[4 * m + b] = (uint8_t)(vB >> (24 - 8 * b));
            }
        }
        p = pn;
    }
}

}  // namespace

// Throughput kernel.  One warp per block; a warp decodes groups of 64 frames (lane L: frames
// 64g+L and 64g+32+L).  The grid is persistent (as many warps as fit on the device) and groups are
// handed out dynamically through a ticket counter so that uneven progress does not leave SM
// sub-partitions idle at the tail.  scratch: [ticket counter, 256 B][per warp: steps x 32 uint4].
template <bool kWordStores>
__global__ void __launch_bounds__(kVitThreads, kVitMinBlocks)
viterbi_pair_kernel(const uint8_t* __restrict__ syms, uint8_t* __restrict__ out, uint8_t* __restrict__ scratch,
                    unsigned long long nframes, uint32_t framebits) {
    const uint32_t steps = framebits + 6;  // framebits is even: 2 * ((F + 6) / 2) == F + 6
    const size_t rowbytes = (size_t)4 * steps, outbytes = (framebits + 7) / 8;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long warp = blockIdx.x, nwarps = gridDim.x;
    const unsigned long long ngroups = (nframes + 63) / 64;
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(scratch);
    uint4* dec = reinterpret_cast<uint4*>(scratch + kVitScratchHeader) + warp * (size_t)steps * 32 + lane;

    unsigned long long g = warp;  // first group is static, later ones come from the ticket counter
    while (g < ngroups) {
        const unsigned long long fA = g * 64 + lane, fB = fA + 32;
        const bool liveA = fA < nframes, liveB = fB < nframes;
        const uint2* rowA = reinterpret_cast<const uint2*>(syms + (liveA ? fA : nframes - 1) * rowbytes);
        const uint2* rowB = reinterpret_cast<const uint2*>(syms + (liveB ? fB : nframes - 1) * rowbytes);

        uint32_t X[64], Y[64];
        X[0] = 0u;  // Locals256: start state 0 has metric 0, all others 63 (deconvolve.cpp:130-132)
#pragma unroll
        for (int s = 1; s < 64; s++) X[s] = kM63;

        // Two steps per iteration (the reference's Butterfly256 granularity): even step X -> Y with the
        // pending renormalisation folded in, odd step Y -> X, then the renormalisation test on the new
        // metric of state 0.  The 8 symbol bytes of the next iteration are fetched while this one runs; every
        // fourth such load starts a new 32-byte sector, and one iteration of lead did not cover an L2 / HBM
        // round trip (ncu: 9 % of the loop's samples waited on it), so the sector kSymPrefetchPairs
        // iterations ahead is prefetched into L1 (no registers: the loop body has none to spare).
        // (A 6-step body would make the register renaming of the butterfly network close on itself
        // and save ~30 moves per step, but it overflows the instruction cache once warps are in
        // different phases: measured 94 vs 118 Gbit/s on the MSC batch.)
        uint32_t neg = 0u;
        uint2 a0 = __ldg(rowA), b0 = __ldg(rowB);  // steps >= 8: the first pairs always exist
#if VIT_SYM_PREFETCH == 2
        uint2 a1 = __ldg(rowA + 1), b1 = __ldg(rowB + 1);
#endif
        const uint32_t last_pair = steps / 2 - 1;
        for (uint32_t t = 0; t < steps; t += 2) {
#if VIT_SYM_PREFETCH == 2
            uint2 a2 = a1, b2 = b1;
            if (t + 4 < steps) a2 = __ldg(rowA + (t >> 1) + 2), b2 = __ldg(rowB + (t >> 1) + 2);
#else
            uint2 na0 = a0, nb0 = b0;
            if (t + 2 < steps) na0 = __ldg(rowA + (t >> 1) + 1), nb0 = __ldg(rowB + (t >> 1) + 1);
#endif
#if VIT_SYM_PREFETCH == 1
            {
                const uint32_t ahead = min((t >> 1) + kSymPrefetchPairs, last_pair);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rowA + ahead));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rowB + ahead));
            }
#endif
            dec[(size_t)(t + 0) * 32] = acs_step<true>(X, Y, a0.x, b0.x, neg);
            dec[(size_t)(t + 1) * 32] = acs_step<false>(Y, X, a0.y, b0.y, 0u);
            neg = renorm_addend(X[0]);
#if VIT_SYM_PREFETCH == 2
            a0 = a1, b0 = b1, a1 = a2, b1 = b2;
#else
            a0 = na0, b0 = nb0;
#endif
        }
        (void)last_pair;
        traceback<kWordStores>(dec, lane, framebits, out + fA * outbytes, out + fB * outbytes, liveA, liveB);

        unsigned long long next = 0;
        if (lane == 0) next = nwarps + atomicAdd(ticket, 1ull);
        g = __shfl_sync(0xffffffffu, next, 0);
    }
}

// ---------------------------------------------------------------------------------------------------
// Latency / small-batch kernel: one warp per frame, two path metrics per lane, survivor decisions as warp
// ballots in shared memory, traceback out of shared memory -- decisions never leave the SM.  This is the
// layout of the design brief, tuned for the length of the per-step dependency chain, which is all that matters
// when one warp runs alone on an SM sub-partition (the single-frame drop-in call):
//
//   * ONE shuffle per trellis step.  The 6 state bits are spread over 5 lane bits and 1 "slot" bit (which of the
//     lane's two registers).  A butterfly needs the two old states that differ in state bit 5 in one lane and
//     produces the two new states that differ in bit 0, so after a butterfly the slot holds bit 0 and some lane
//     bit holds the new bit 5.  Instead of restoring a fixed layout (two shuffles of a packed pair plus
//     pack/unpack, as the first version of this kernel did), the assignment of state bits to lane bits rotates:
//     one shfl.xor across exactly that lane bit swaps it with the slot -- each lane keeps one of its two new
//     metrics and trades the other with its partner.  Before butterfly t lane bit l holds state bit
//     (l + t) mod 5, the exchange after it crosses lane bit 4 - (t mod 5), the period is 5 steps.
//     The butterfly index of a lane is rotl5(lane, t mod 5), so its branch mask is one of five per-lane constants.
//   * The serial chain per step is select -> shuffle -> select -> add-min -> min (about 40 cycles); the branch
//     metric of the step (a dot-product form of the reference's two-level pavgb, see branch_metric_dp), the
//     ballots, the renormalisation test and the decision store hang off it.
//   * Traceback runs in "lane coordinates" (which lane / slot holds the current state), where one step is a
//     rotate and a bit-select: the decision words are stored pre-rotated so that the decision bit lands on the
//     lane bit it replaces.  The decoded bits are collected afterwards by the whole warp.
// ~34 warp-instructions per trellis step against 6.5 per frame-step for the pair kernel, so it is used where the
// pair kernel cannot fill the machine: the single-frame drop-in call and batches below kVitWarpKernelMaxFrames.
// ---------------------------------------------------------------------------------------------------
namespace {

// Branch metric of deconvolve.cpp:338-349, m = avg(avg(x0,x1), avg(x2,x3)) >> 2 with avg(a,b) = (a+b+1) >> 1,
// without the two-level rounding: with S = x0+x1+x2+x3 and q = lsb(x0^x1) + lsb(x2^x3) it equals
// (S + 2 + q) >> 4 (each first-level average rounds up exactly when its two bytes differ in parity), which is
// two byte dot-products (tests/test_device_code_on_host.py checks the identity exhaustively per byte pair).
__device__ __forceinline__ uint32_t branch_metric_dp(uint32_t w, uint32_t xmask) {
    const uint32_t x = w ^ xmask;
    const uint32_t s = __dp4a(x, 0x01010101u, 2u);
    const uint32_t par = (x ^ (x >> 8)) & 0x00010001u;
    return __dp4a(par, 0x01010101u, s) >> 4;
}

__device__ __forceinline__ uint32_t rotl5(uint32_t v, uint32_t r) { return ((v << r) | (v >> (5u - r))) & 31u; }

// One trellis step of the warp kernel.  kPhase = t mod 5, kOdd = t & 1 (the renormalisation test follows odd
// steps: deconvolve.cpp:407-412).  A / B: the lane's old states with state bit 5 = 0 / 1.
template <int kPhase, bool kOdd>
__device__ __forceinline__ void warp_step(uint32_t& A, uint32_t& B, uint32_t w, uint32_t xmask, uint32_t lane,
                                          uint2* dec_slot) {
    constexpr uint32_t kFull = 0xffffffffu;
    constexpr int kLaneBit = 4 - kPhase;            // lane bit that holds the new state bit 5
    constexpr int kPrevBit = (5 - kPhase) % 5;      // lane bit the exchange before this butterfly crossed
    const uint32_t m = branch_metric_dp(w, xmask), mm = 63u - m;
    // ACS (deconvolve.cpp:352-359); ties choose the upper predecessor (decision = 1)
    bool pe, po;
    const uint32_t ne = __vibmin_u32(__viaddmin_u32(B, mm, 255u), A + m, &pe);
    const uint32_t no = __vibmin_u32(__viaddmin_u32(B, m, 255u), A + mm, &po);
    const bool up = (lane >> kLaneBit) & 1u;
    const uint32_t recv = __shfl_xor_sync(kFull, up ? ne : no, 1u << kLaneBit);
    const uint32_t keep = up ? no : ne;
    uint32_t n0 = 0;
    if (kOdd) n0 = __shfl_sync(kFull, ne, 0);  // new state 0 always lives in lane 0, slot 0
    const uint32_t be = __ballot_sync(kFull, pe), bo = __ballot_sync(kFull, po);
    A = up ? recv : keep;
    B = up ? keep : recv;
    if (kOdd) {  // Renormalize256: metric[state 0] > 150 -> all metrics -= 63, clamped at 0
        const int neg = n0 > 150u ? -63 : 0;
        A = (uint32_t)__viaddmax_s32_relu((int)A, neg, 0);
        B = (uint32_t)__viaddmax_s32_relu((int)B, neg, 0);
    }
    // decision of the new state in (lane l, slot e/o) is bit l of be / bo; stored rotated left by kPrevBit so that
    // the traceback's rotate-right by its lane number drops the bit onto lane bit kPrevBit
    if (lane == 0) *dec_slot = make_uint2(__funnelshift_l(be, be, kPrevBit), __funnelshift_l(bo, bo, kPrevBit));
}

// One traceback step in lane coordinates: (ln, w) = lane number of the current state and the decision word of
// its slot.  The predecessor keeps the lane (a butterfly is lane-local) with slot = the decision; undoing the
// exchange before the butterfly swaps that slot with lane bit kPrevBit.
template <int kPrevBit>
__device__ __forceinline__ void warp_trace_step(uint32_t& ln, uint32_t& w, const uint2 next, uint32_t* x_slot) {
    const uint32_t x = __funnelshift_r(w, w, ln);       // decision bit of lane ln -> bit kPrevBit
    w = (ln >> kPrevBit) & 1u ? next.y : next.x;         // slot of the predecessor = the lane bit it replaces
    *x_slot = x;
    ln = (ln & ~(1u << kPrevBit)) | (x & (1u << kPrevBit));
}

}  // namespace

// kU32: the symbols arrive in QIRX's one-uint32-per-symbol layout (low byte used, deconvolve.cpp:219-228) and are
// compacted while they are staged.  The decoded bytes are collected in shared memory and written out by the
// whole warp, so both ends work on host-mapped (pinned) memory as well: the single-frame drop-in call runs this
// kernel straight on the caller's bounce buffer, with no copy operations around it.
template <bool kU32>
__global__ void __launch_bounds__(32) viterbi_warp_kernel(const void* __restrict__ syms_any, uint8_t* __restrict__ out,
                                                          unsigned long long nframes, uint32_t framebits) {
    extern __shared__ __align__(16) uint8_t wsmem[];
    const uint32_t steps = framebits + 6, lane = threadIdx.x;
    uint32_t* s_sym = reinterpret_cast<uint32_t*>(wsmem);         // [steps] 4 symbols per step; reused by the traceback
    uint2* s_dec = reinterpret_cast<uint2*>(wsmem + 4 * (size_t)steps);  // [steps] rotated {even, odd} ballots
    const size_t outbytes = (framebits + 7) / 8;

    // branch masks (const.asm:35-49 restated: 0xFF where the expected code bit is 1) of the five butterflies
    // this lane runs in turn: index rotl5(lane, t mod 5)
    uint32_t xm[5];
#pragma unroll
    for (uint32_t r = 0; r < 5; r++) {
        const uint32_t i = r ? rotl5(lane, r) : lane;
        xm[r] = (parity8((2u * i) & kPoly(0)) ? 0xFF0000FFu : 0u) |  // polys 0 and 3 coincide
                (parity8((2u * i) & kPoly(1)) ? 0x0000FF00u : 0u) | (parity8((2u * i) & kPoly(2)) ? 0x00FF0000u : 0u);
    }

    for (unsigned long long f = blockIdx.x; f < nframes; f += gridDim.x) {
        __syncwarp();
        if (kU32) {
            const uint4* row = reinterpret_cast<const uint4*>(syms_any) + f * (size_t)steps;  // one step per uint4
            // eight loads in flight per lane: over PCIe (host-mapped input) each round trip costs ~1.5 us
            for (uint32_t i0 = lane; i0 < steps; i0 += 32 * 8) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (i0 + 32 * u < steps) v[u] = __ldg(row + i0 + 32 * u);
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (i0 + 32 * u < steps)
                        s_sym[i0 + 32 * u] = (v[u].x & 0xFFu) | ((v[u].y & 0xFFu) << 8) | ((v[u].z & 0xFFu) << 16) | (v[u].w << 24);
            }
        } else {
            const uint2* row = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(syms_any) + f * 4 * (size_t)steps);
            for (uint32_t i0 = lane; i0 < steps / 2; i0 += 32 * 8) {
                uint2 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (i0 + 32 * u < steps / 2) v[u] = __ldg(row + i0 + 32 * u);
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (i0 + 32 * u < steps / 2) reinterpret_cast<uint2*>(s_sym)[i0 + 32 * u] = v[u];
            }
        }
        __syncwarp();

        // ---- forward pass: ten steps per iteration (period 5 of the layout x period 2 of the renormalisation) ----
        uint32_t A = (lane == 0) ? 0u : 63u, B = 63u;  // Locals256: M[0] = 0, others 63 (deconvolve.cpp:130-132)
        uint32_t t = 0;
        for (; t + 10 <= steps; t += 10) {
            uint32_t w[10];
#pragma unroll
            for (int u = 0; u < 10; u++) w[u] = s_sym[t + u];
            warp_step<0, false>(A, B, w[0], xm[0], lane, s_dec + t + 0);
            warp_step<1, true>(A, B, w[1], xm[1], lane, s_dec + t + 1);
            warp_step<2, false>(A, B, w[2], xm[2], lane, s_dec + t + 2);
            warp_step<3, true>(A, B, w[3], xm[3], lane, s_dec + t + 3);
            warp_step<4, false>(A, B, w[4], xm[4], lane, s_dec + t + 4);
            warp_step<0, true>(A, B, w[5], xm[0], lane, s_dec + t + 5);
            warp_step<1, false>(A, B, w[6], xm[1], lane, s_dec + t + 6);
            warp_step<2, true>(A, B, w[7], xm[2], lane, s_dec + t + 7);
            warp_step<3, false>(A, B, w[8], xm[3], lane, s_dec + t + 8);
            warp_step<4, true>(A, B, w[9], xm[4], lane, s_dec + t + 9);
        }
        // the remaining 0, 2, ... 8 steps (steps is even): t is a multiple of 10 here
        if (t < steps) {
            warp_step<0, false>(A, B, s_sym[t + 0], xm[0], lane, s_dec + t + 0);
            warp_step<1, true>(A, B, s_sym[t + 1], xm[1], lane, s_dec + t + 1);
        }
        if (t + 2 < steps) {
            warp_step<2, false>(A, B, s_sym[t + 2], xm[2], lane, s_dec + t + 2);
            warp_step<3, true>(A, B, s_sym[t + 3], xm[3], lane, s_dec + t + 3);
        }
        if (t + 4 < steps) {
            warp_step<4, false>(A, B, s_sym[t + 4], xm[4], lane, s_dec + t + 4);
            warp_step<0, true>(A, B, s_sym[t + 5], xm[0], lane, s_dec + t + 5);
        }
        if (t + 6 < steps) {
            warp_step<1, false>(A, B, s_sym[t + 6], xm[1], lane, s_dec + t + 6);
            warp_step<2, true>(A, B, s_sym[t + 7], xm[2], lane, s_dec + t + 7);
        }
        __syncwarp();

        // ---- ChainBack (deconvolve.cpp:416-435) by lane 0, from state 0 (lane 0, slot 0) after the last step;
        // the rotated decision word of every step replaces the symbol word of that step in s_sym -----------------
        if (lane == 0 && framebits > 0) {
            uint32_t ln = 0;
            int tau = (int)steps - 1;  // decisions of steps 6 .. F+5 are consumed (t = tau - 6)
            uint32_t w = s_dec[tau].x;
            auto generic = [&](int tt) {  // one step with a run-time phase
                const uint32_t pb = (5u - (uint32_t)tt % 5u) % 5u;
                const uint2 next = s_dec[tt > 6 ? tt - 1 : tt];
                const uint32_t x = __funnelshift_r(w, w, ln);
                w = (ln >> pb) & 1u ? next.y : next.x;
                s_sym[tt] = x;
                ln = (ln & ~(1u << pb)) | (x & (1u << pb));
            };
            for (; tau >= 6 && tau % 5 != 4; tau--) generic(tau);
            for (; tau >= 10; tau -= 5) {  // tau % 5 == 4: the five records do not depend on the state
                uint2 nx[5];
#pragma unroll
                for (int j = 0; j < 5; j++) nx[j] = s_dec[tau - 1 - j];
                warp_trace_step<1>(ln, w, nx[0], s_sym + tau);
                warp_trace_step<2>(ln, w, nx[1], s_sym + tau - 1);
                warp_trace_step<3>(ln, w, nx[2], s_sym + tau - 2);
                warp_trace_step<4>(ln, w, nx[3], s_sym + tau - 3);
                warp_trace_step<0>(ln, w, nx[4], s_sym + tau - 4);
            }
            for (; tau >= 6; tau--) generic(tau);
        }
        __syncwarp();
        // decoded bit t = the decision consumed at step t + 6 = bit (5 - (t+6) % 5) % 5 of its rotated word;
        // output byte n holds bits 8n .. 8n+7, MSB first (missing bits of a ragged last byte stay 0)
        for (uint32_t n = lane; n < outbytes; n += 32) {
            uint32_t v = 0;
#pragma unroll
            for (uint32_t j = 0; j < 8; j++) {
                const uint32_t tb = 8 * n + j;
                if (tb < framebits) {
                    const uint32_t tt = tb + 6;
                    v |= ((s_sym[tt] >> ((5u - tt % 5u) % 5u)) & 1u) << (7 - j);
                }
            }
            out[f * outbytes + n] = (uint8_t)v;
        }
    }
}

// u32 -> u8 compaction for the QIRX one-word-per-symbol layout (low byte only, deconvolve.cpp:219-228)
__global__ void __launch_bounds__(256) compact_symbols_kernel(const uint4* __restrict__ in, uint32_t* __restrict__ outw,
                                                              size_t nquads) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquads; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(in + i);
        outw[i] = (v.x & 0xFFu) | ((v.y & 0xFFu) << 8) | ((v.z & 0xFFu) << 16) | (v.w << 24);
    }
}

// Depuncturing front end: the receiver transmits only the code bits its puncturing
// vector keeps; the step before deconvolve() puts them back at their positions in the rate-1/4 mother-code
// layout and fills the punctured positions with the erasure value (the soft-symbol midpoint).  idx[p] is the
// position of mother-code symbol p inside a received row, or -1 when it was punctured.  One thread builds the
// four symbols of one trellis step (one aligned 32-bit store).
__global__ void __launch_bounds__(256) depuncture_kernel(const uint8_t* __restrict__ rx, size_t rx_per_frame,
                                                         const int4* __restrict__ idx, uint32_t steps,
                                                         uint32_t erasure, uint32_t* __restrict__ out, size_t n) {
    const size_t total = n * (size_t)steps;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / steps;
        const uint32_t q = (uint32_t)(i - f * steps);
        const int4 ix = __ldg(idx + q);
        const uint8_t* row = rx + f * rx_per_frame;
        const uint32_t b0 = ix.x >= 0 ? __ldg(row + ix.x) : erasure;
        const uint32_t b1 = ix.y >= 0 ? __ldg(row + ix.y) : erasure;
        const uint32_t b2 = ix.z >= 0 ? __ldg(row + ix.z) : erasure;
        const uint32_t b3 = ix.w >= 0 ? __ldg(row + ix.w) : erasure;
        out[i] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
}

size_t viterbi_scratch_bytes(int grid_blocks, uint32_t framebits) {
    return kVitScratchHeader + (size_t)grid_blocks * (size_t)(framebits + 6) * 32 * sizeof(uint4);
}

// Persistent grid: as many one-warp blocks as are resident at once, fewer when the batch is small, and capped so
// that the decision scratch (512 B per block per trellis step) stays within a byte budget: at F = 9216 the full
// grid would take 11 GB per call (VITERBI_B200_SCRATCH_MB overrides the 6 GiB default).
int viterbi_grid_blocks(int num_sms, unsigned long long nframes, uint32_t framebits) {
    static const unsigned long long budget = [] {
        const char* env = getenv("VITERBI_B200_SCRATCH_MB");
        const long long mb = (env && *env) ? atoll(env) : 6144;
        return (unsigned long long)(mb >= 16 ? mb : 6144) << 20;
    }();
    const unsigned long long groups = (nframes + 63) / 64;
    unsigned long long blocks = (unsigned long long)num_sms * kVitMinBlocks;
    const unsigned long long fit = budget / ((unsigned long long)(framebits + 6) * 32 * sizeof(uint4));
    if (blocks > fit) blocks = fit > 0 ? fit : 1;
    return (int)(groups < blocks ? groups : blocks);
}

cudaError_t launch_viterbi_pair(const uint8_t* d_syms, uint8_t* d_out, void* d_scratch, unsigned long long nframes,
                                uint32_t framebits, int grid_blocks, cudaStream_t stream) {
    if (nframes == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, kVitScratchHeader, stream);  // ticket counter
    if (e != cudaSuccess) return e;
    // 32-bit output stores need F % 32 == 0 (whole words per row) and a 4-byte aligned d_out
    if (framebits % 32 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3) == 0)
        viterbi_pair_kernel<true><<<grid_blocks, kVitThreads, 0, stream>>>(d_syms, d_out, (uint8_t*)d_scratch, nframes,
                                                                            framebits);
    else
        viterbi_pair_kernel<false><<<grid_blocks, kVitThreads, 0, stream>>>(d_syms, d_out, (uint8_t*)d_scratch, nframes,
                                                                             framebits);
    count_launch();
    return cudaGetLastError();
}

size_t viterbi_warp_smem_bytes(uint32_t framebits) { return 12 * (size_t)(framebits + 6) + 16; }

// The warp kernel needs more than 48 KB of dynamic shared memory above F = 4090.  The opt-in is a per-device
// function attribute, so it is raised once per device to the worst case (F = 9216) from the device
// initialisation in fec_api.cu -- not lazily per launch, where concurrent callers with different frame sizes
// would lower each other's limit.
cudaError_t viterbi_configure_device() {
    const int worst = (int)viterbi_warp_smem_bytes(kMaxFramebits);
    cudaError_t e = cudaFuncSetAttribute(viterbi_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, worst);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(viterbi_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, worst);
}

template <bool kU32>
static cudaError_t launch_viterbi_warp_t(const void* d_syms, uint8_t* d_out, unsigned long long nframes, uint32_t framebits,
                                         int num_sms, cudaStream_t stream) {
    if (nframes == 0) return cudaSuccess;
    const size_t smem = viterbi_warp_smem_bytes(framebits);
    const unsigned long long cap = (unsigned long long)num_sms * 32;
    const unsigned grid = (unsigned)(nframes < cap ? nframes : cap);
    viterbi_warp_kernel<kU32><<<grid, 32, smem, stream>>>(d_syms, d_out, nframes, framebits);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_viterbi_warp(const uint8_t* d_syms, uint8_t* d_out, unsigned long long nframes, uint32_t framebits,
                                int num_sms, cudaStream_t stream) {
    return launch_viterbi_warp_t<false>(d_syms, d_out, nframes, framebits, num_sms, stream);
}

cudaError_t launch_viterbi_warp_u32(const uint32_t* d_syms, uint8_t* d_out, unsigned long long nframes, uint32_t framebits,
                                    int num_sms, cudaStream_t stream) {
    return launch_viterbi_warp_t<true>(d_syms, d_out, nframes, framebits, num_sms, stream);
}

cudaError_t launch_depuncture(const uint8_t* d_rx, size_t rx_per_frame, const int32_t* d_idx, uint32_t framebits,
                              uint32_t erasure, uint8_t* d_syms, size_t nframes, int num_sms, cudaStream_t stream) {
    if (nframes == 0) return cudaSuccess;
    const uint32_t steps = framebits + 6;
    const size_t total = nframes * (size_t)steps;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms * 16;
    if (blocks > cap) blocks = cap;
    depuncture_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_rx, rx_per_frame, (const int4*)d_idx, steps, erasure & 0xFFu,
                                                          (uint32_t*)d_syms, nframes);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_compact_symbols(const uint32_t* d_in, uint8_t* d_out, size_t nsymbols, int num_sms,
                                   cudaStream_t stream) {
    if (nsymbols == 0) return cudaSuccess;
    const size_t nquads = nsymbols / 4;  // symbol count per frame is a multiple of 4
    size_t blocks = (nquads + 255) / 256;
    const size_t cap = (size_t)num_sms * 16;
    if (blocks > cap) blocks = cap;
    compact_symbols_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint4*)d_in, (uint32_t*)d_out, nquads);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fec
