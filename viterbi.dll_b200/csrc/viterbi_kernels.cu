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
// Latency / small-batch kernel: one warp per frame, lane L holds states L and L+32.
//
// This is the layout of the design brief: the butterfly L (old states L, L+32 -> new states 2L, 2L+1)
// needs no communication; the next step then needs new states L and L+32, which lanes L>>1 and
// 16+(L>>1) hold, so each step exchanges one packed {even, odd} pair through two shuffles.  Soft
// symbols are staged into shared memory with coalesced 8-byte loads and read back as one broadcast word
// per step; the 64 decision bits of a step are two warp ballots (even / odd new states) stored in shared
// memory, and lane 0 runs the traceback out of shared memory -- decisions never leave the SM.
// It costs ~30 warp-instructions per trellis step (the pair kernel: 416 per 64 frame-steps = 6.5), so
// it is used where the pair kernel cannot fill the machine: the single-frame drop-in call and batches
// below kVitWarpKernelMaxFrames.
// ---------------------------------------------------------------------------------------------------
// kU32: the symbols arrive in QIRX's one-uint32-per-symbol layout (low byte used, deconvolve.cpp:219-228) and are
// compacted while they are staged.  The decoded bytes are collected in shared memory and written out by the
// whole warp, so both ends work on host-mapped (pinned) memory as well: the single-frame drop-in call runs this
// kernel straight on the caller's bounce buffer, with no copy operations around it.
template <bool kU32>
__global__ void __launch_bounds__(32) viterbi_warp_kernel(const void* __restrict__ syms_any, uint8_t* __restrict__ out,
                                                          unsigned long long nframes, uint32_t framebits) {
    extern __shared__ __align__(16) uint8_t wsmem[];
    const uint32_t steps = framebits + 6, lane = threadIdx.x;
    uint32_t* s_sym = reinterpret_cast<uint32_t*>(wsmem);         // [steps] 4 symbols per step
    uint2* s_dec = reinterpret_cast<uint2*>(wsmem + 4 * (size_t)steps);  // [steps] {even, odd} ballots
    uint8_t* s_out = wsmem + 12 * (size_t)steps;                  // [ceil(F/8)] decoded bytes
    const size_t outbytes = (framebits + 7) / 8;

    // branch masks of butterfly `lane` (const.asm:35-49 restated): 0xFF where the expected code bit is 1
    const uint32_t xmask = (parity8((2u * lane) & kPoly(0)) ? 0xFF0000FFu : 0u) |  // polys 0 and 3 coincide
                           (parity8((2u * lane) & kPoly(1)) ? 0x0000FF00u : 0u) |
                           (parity8((2u * lane) & kPoly(2)) ? 0x00FF0000u : 0u);
    const uint32_t half_sel = (lane & 1u) ? 0x4432u : 0x4410u;  // pick the odd / even half of a packed pair
    const uint32_t srcA = lane >> 1, srcB = 16u + (lane >> 1);

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

        uint32_t A = (lane == 0) ? 0u : 63u, B = 63u;  // Locals256: M[0] = 0, others 63 (deconvolve.cpp:130-132)
        // branch metric (deconvolve.cpp:338-349): avg(avg(x0,x1), avg(x2,x3)) >> 2, pavgb rounds up.
        // It does not depend on the path metrics, so the metric of step t+1 is computed while the
        // shuffles of step t are in flight: the serial chain per step is add-min, add-min, pack,
        // shuffle, unpack.
        auto branch_metric = [xmask](uint32_t w) {
            const uint32_t x = w ^ xmask;
            const uint32_t s = (x & 0x00FF00FFu) + ((x >> 8) & 0x00FF00FFu) + 0x00010001u;  // x0+x1+1 | x2+x3+1
            const uint32_t ab = (s >> 1) & 0x00FF00FFu;
            return ((ab & 0xFFFFu) + (ab >> 16) + 1u) >> 3;
        };
        uint32_t m = branch_metric(s_sym[0]);
        for (uint32_t t = 0; t < steps; t++) {
            const uint32_t wnext = s_sym[t + 1 < steps ? t + 1 : t];
            const uint32_t mm = 63u - m;
            // ACS (deconvolve.cpp:352-359); ties choose the upper predecessor (decision = 1)
            const uint32_t t1 = __viaddmin_u32(B, mm, 255u), ne = __viaddmin_u32(A, m, t1);
            const uint32_t t3 = __viaddmin_u32(B, m, 255u), no = __viaddmin_u32(A, mm, t3);
            const uint32_t pair = __byte_perm(ne, no, 0x5410);  // {N[2L], N[2L+1]} as 16-bit halves
            const uint32_t pa = __shfl_sync(0xffffffffu, pair, srcA), pb = __shfl_sync(0xffffffffu, pair, srcB);
            const uint32_t be = __ballot_sync(0xffffffffu, ne == t1), bo = __ballot_sync(0xffffffffu, no == t3);
            // Renormalize256 (deconvolve.cpp:407-412): new state 0 is `ne` of lane 0
            const bool renorm = (t & 1u) && (__ballot_sync(0xffffffffu, ne > 150u) & 1u);
            if (lane == 0) s_dec[t] = make_uint2(be, bo);
            m = branch_metric(wnext);
            A = __byte_perm(pa, 0u, half_sel);
            B = __byte_perm(pb, 0u, half_sel);
            if (renorm) {
                A = A > 63u ? A - 63u : 0u;
                B = B > 63u ? B - 63u : 0u;
            }
        }
        __syncwarp();
        // ChainBack (deconvolve.cpp:416-435) by lane 0, same 32-bit state register as the pair kernel:
        // state = h >> 26; its decision is bit (state >> 1) of the even / odd ballot word.
        if (lane == 0) {
            uint8_t* o = s_out;
            uint32_t h = 0;
            int t = (int)framebits - 1;
            auto step = [&](const uint2 w, int tt) {
                const uint32_t x = __funnelshift_r((h & 0x04000000u) ? w.y : w.x, 0u, h >> 27);
                h = __funnelshift_r(h, x, 1);
                if ((tt & 7) == 0) o[tt >> 3] = (uint8_t)(h >> 24);
            };
            for (; (t & 7) != 7 && t >= 0; t--) step(s_dec[t + 6], t);  // ragged top (framebits % 8 != 0)
            for (; t >= 7; t -= 8) {  // 8 steps per iteration: the loads do not depend on the state
                uint2 w[8];
#pragma unroll
                for (int j = 0; j < 8; j++) w[j] = s_dec[t - j + 6];
#pragma unroll
                for (int j = 0; j < 8; j++) step(w[j], t - j);
            }
        }
        __syncwarp();
        for (uint32_t i = lane; i < outbytes; i += 32) out[f * outbytes + i] = s_out[i];
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

int viterbi_grid_blocks(int num_sms, unsigned long long nframes) {
    const unsigned long long groups = (nframes + 63) / 64;
    const unsigned long long resident = (unsigned long long)num_sms * kVitMinBlocks;
    return (int)(groups < resident ? groups : resident);
}

cudaError_t launch_viterbi_pair(const uint8_t* d_syms, uint8_t* d_out, void* d_scratch, unsigned long long nframes,
                                uint32_t framebits, int grid_blocks, cudaStream_t stream) {
    if (nframes == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, kVitScratchHeader, stream);  // ticket counter
    if (e != cudaSuccess) return e;
    if (framebits % 32 == 0)
        viterbi_pair_kernel<true><<<grid_blocks, kVitThreads, 0, stream>>>(d_syms, d_out, (uint8_t*)d_scratch, nframes,
                                                                            framebits);
    else
        viterbi_pair_kernel<false><<<grid_blocks, kVitThreads, 0, stream>>>(d_syms, d_out, (uint8_t*)d_scratch, nframes,
                                                                             framebits);
    count_launch();
    return cudaGetLastError();
}

size_t viterbi_warp_smem_bytes(uint32_t framebits) { return 12 * (size_t)(framebits + 6) + ((framebits + 7) / 8 + 15) / 16 * 16; }

template <bool kU32>
static cudaError_t launch_viterbi_warp_t(const void* d_syms, uint8_t* d_out, unsigned long long nframes, uint32_t framebits,
                                         int num_sms, cudaStream_t stream) {
    if (nframes == 0) return cudaSuccess;
    const size_t smem = viterbi_warp_smem_bytes(framebits);
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(viterbi_warp_kernel<kU32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
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
