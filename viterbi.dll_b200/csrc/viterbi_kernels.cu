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
//
// kPunct: the depuncturing front end fused into the symbol fetch.  `syms` then holds only
// the transmitted soft symbols, rx_per_frame bytes per frame, and ptab one entry per loop iteration (two trellis
// steps = eight mother-code symbols): .x = offset of the iteration's first transmitted symbol inside a row, .y =
// 8 x (transmitted symbols among the first four), .z = two byte-permute selectors that spread the run over the
// two step words and put the erasure byte everywhere else.  A thread reads the three aligned words that hold its
// run of at most eight bytes (rows are dense, so a row starts at any byte), funnel-shifts them into place and
// expands: ~22 more instructions per iteration instead of a separate pass that writes and re-reads the expanded
// 4(F+6) bytes per frame.  Measured on device-resident input (FIC 65,536 / MSC 262,144 frames, half of the symbols
// punctured): 1.25x / 1.12x the unpunctured time, against 1.44x / 1.52x with the separate depuncturing kernel in
// front.  What is left is the load/store unit: three uncoalesced 4-byte loads per frame and iteration (96 sectors
// per warp instruction pair against 32 for the plain row fetch).  The last row is read from a padded copy when the
// buffer may end right behind it (last_row).
template <bool kWordStores, bool kPunct>
__global__ void __launch_bounds__(kVitThreads, kVitMinBlocks)
viterbi_pair_kernel(const uint8_t* __restrict__ syms, uint8_t* __restrict__ out, uint8_t* __restrict__ scratch,
                    unsigned long long nframes, uint32_t framebits, const uint4* __restrict__ ptab, uint32_t rx_per_frame,
                    uint32_t erasure_word, const uint8_t* __restrict__ last_row) {
    const uint32_t steps = framebits + 6;  // framebits is even: 2 * ((F + 6) / 2) == F + 6
    const size_t rowbytes = kPunct ? (size_t)rx_per_frame : (size_t)4 * steps, outbytes = (framebits + 7) / 8;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long warp = blockIdx.x, nwarps = gridDim.x;
    const unsigned long long ngroups = (nframes + 63) / 64;
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(scratch);
    uint4* dec = reinterpret_cast<uint4*>(scratch + kVitScratchHeader) + warp * (size_t)steps * 32 + lane;

    unsigned long long g = warp;  // first group is static, later ones come from the ticket counter
    while (g < ngroups) {
        const unsigned long long fA = g * 64 + lane, fB = fA + 32;
        const bool liveA = fA < nframes, liveB = fB < nframes;
        // kPunct: the fetch may touch a few bytes past a row, which is the next row except for the very last one:
        // that row is read from a padded copy when the caller's buffer has no slack (last_row != nullptr)
        const unsigned long long rA = liveA ? fA : nframes - 1, rB = liveB ? fB : nframes - 1;
        const uint8_t* rawA = (kPunct && last_row != nullptr && rA == nframes - 1) ? last_row : syms + rA * rowbytes;
        const uint8_t* rawB = (kPunct && last_row != nullptr && rB == nframes - 1) ? last_row : syms + rB * rowbytes;
        const uint2* rowA = reinterpret_cast<const uint2*>(rawA);
        const uint2* rowB = reinterpret_cast<const uint2*>(rawB);
        // kPunct: the row's aligned base and its misalignment (0..3)
        const uint32_t misA = (uint32_t)(reinterpret_cast<uintptr_t>(rawA) & 3), misB = (uint32_t)(reinterpret_cast<uintptr_t>(rawB) & 3);
        const uint8_t* alA = rawA - misA;
        const uint8_t* alB = rawB - misB;
        auto fetch_run = [](const uint8_t* al, uint32_t mis, uint32_t off, uint32_t (&w)[3]) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(al + ((mis + off) & ~3u));
            w[0] = __ldg(p), w[1] = __ldg(p + 1), w[2] = __ldg(p + 2);
        };
        auto expand = [erasure_word](const uint32_t (&w)[3], uint32_t mis, const uint4 e) {
            const uint32_t sh = ((mis + e.x) & 3u) * 8u;
            const uint32_t lo = __funnelshift_r(w[0], w[1], sh), hi = __funnelshift_r(w[1], w[2], sh);  // the run, byte 0 first
            const uint32_t q = __funnelshift_rc(lo, hi, e.y);  // ... from the second step's first transmitted symbol on
            return make_uint2(__byte_perm(lo, erasure_word, e.z & 0xFFFFu), __byte_perm(q, erasure_word, e.z >> 16));
        };

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
        uint2 a0, b0;
        if (kPunct) {
            const uint4 e = __ldg(ptab);
            uint32_t wa[3], wb[3];
            fetch_run(alA, misA, e.x, wa);
            fetch_run(alB, misB, e.x, wb);
            a0 = expand(wa, misA, e), b0 = expand(wb, misB, e);
        } else {
            a0 = __ldg(rowA), b0 = __ldg(rowB);  // steps >= 8: the first pairs always exist
        }
        const uint32_t last_pair = steps / 2 - 1;
        for (uint32_t t = 0; t < steps; t += 2) {
            uint2 na0 = a0, nb0 = b0;
            uint4 en = make_uint4(0u, 0u, 0u, 0u);
            uint32_t wa[3] = {0u, 0u, 0u}, wb[3] = {0u, 0u, 0u};
            if (kPunct) {
                // (loading the table entry one iteration earlier, so that the fetch addresses never wait for it, was
                // measured and was slower: 7.04 vs 6.30 ms on the MSC batch -- four more registers live across the steps)
                if (t + 2 < steps) {
                    en = __ldg(ptab + (t >> 1) + 1);
                    fetch_run(alA, misA, en.x, wa);
                    fetch_run(alB, misB, en.x, wb);
                    // the rows advance by at most 8 bytes per iteration: pull the line ~8 iterations ahead into L1
                    const uint32_t ahead = min(en.x + 64u, rx_per_frame - 1u);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rawA + ahead));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rawB + ahead));
                }
            } else {
                if (t + 2 < steps) na0 = __ldg(rowA + (t >> 1) + 1), nb0 = __ldg(rowB + (t >> 1) + 1);
                const uint32_t ahead = min((t >> 1) + kSymPrefetchPairs, last_pair);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rowA + ahead));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rowB + ahead));
            }
            dec[(size_t)(t + 0) * 32] = acs_step<true>(X, Y, a0.x, b0.x, neg);
            dec[(size_t)(t + 1) * 32] = acs_step<false>(Y, X, a0.y, b0.y, 0u);
            neg = renorm_addend(X[0]);
            if (kPunct && t + 2 < steps) na0 = expand(wa, misA, en), nb0 = expand(wb, misB, en);
            a0 = na0, b0 = nb0;
        }
        traceback<kWordStores>(dec, lane, framebits, out + fA * outbytes, out + fB * outbytes, liveA, liveB);

        unsigned long long next = 0;
        if (lane == 0) next = nwarps + atomicAdd(ticket, 1ull);
        g = __shfl_sync(0xffffffffu, next, 0);
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

// Energy-dispersal descrambler (the step a DAB receiver performs between deconvolve() and
// RScheckSuperframe(); ETSI EN 300 401 clause 10): every decoded logical frame is XORed with the PRBS of
// X^9 + X^5 + 1 restarted (all ones) at the start of the frame.  The sequence does not depend on the frame length,
// so one table of kMaxFramebits / 8 bytes serves every F.  An element-wise pass over F/8 bytes per frame (0.4 % of
// the bytes the Viterbi kernel moves), only used by the chained dabplus_* calls when the option is on.
__constant__ uint32_t c_prbs_words[kMaxFramebits / 32];

__global__ void __launch_bounds__(256) descramble_kernel(uint8_t* __restrict__ bits, size_t nframes, uint32_t bytes_per_frame) {
    const size_t total = nframes * bytes_per_frame;
    const uint8_t* prbs = reinterpret_cast<const uint8_t*>(c_prbs_words);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        bits[i] ^= prbs[i % bytes_per_frame];
}

cudaError_t descramble_upload_table() {
    static uint32_t words[kMaxFramebits / 32];
    uint8_t* bytes = reinterpret_cast<uint8_t*>(words);
    unsigned reg = 0x1FF;  // nine ones; bit 8 = oldest
    for (uint32_t i = 0; i < kMaxFramebits; i++) {
        const unsigned b = ((reg >> 8) ^ (reg >> 4)) & 1u;
        reg = ((reg << 1) | b) & 0x1FF;
        if ((i & 7) == 0) bytes[i >> 3] = 0;
        bytes[i >> 3] |= (uint8_t)(b << (7 - (i & 7)));  // MSB first, like the decoder's output bytes
    }
    return cudaMemcpyToSymbol(c_prbs_words, words, sizeof words);
}

cudaError_t launch_descramble(uint8_t* d_bits, size_t nframes, uint32_t framebits, int num_sms, cudaStream_t stream) {
    if (nframes == 0 || framebits < 8) return cudaSuccess;
    const size_t total = nframes * (framebits / 8);
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms * 16;
    if (blocks > cap) blocks = cap;
    descramble_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_bits, nframes, framebits / 8);
    count_launch();
    return cudaGetLastError();
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
        viterbi_pair_kernel<true, false><<<grid_blocks, kVitThreads, 0, stream>>>(d_syms, d_out, (uint8_t*)d_scratch, nframes,
                                                                                   framebits, nullptr, 0u, 0u, nullptr);
    else
        viterbi_pair_kernel<false, false><<<grid_blocks, kVitThreads, 0, stream>>>(d_syms, d_out, (uint8_t*)d_scratch, nframes,
                                                                                    framebits, nullptr, 0u, 0u, nullptr);
    count_launch();
    return cudaGetLastError();
}

// Punctured input decoded in one pass (kPunct).  d_rx: nframes dense rows of rx_per_frame transmitted symbols;
// d_ptab: punct_table() on the device; d_last_row: nullptr when d_rx is followed by at least kPunctSlackBytes
// readable bytes, else a copy of the last row that is.
cudaError_t launch_viterbi_pair_punctured(const uint8_t* d_rx, uint32_t rx_per_frame, const void* d_ptab, uint32_t erasure,
                                          uint8_t* d_out, void* d_scratch, unsigned long long nframes, uint32_t framebits,
                                          int grid_blocks, cudaStream_t stream, const uint8_t* d_last_row) {
    if (nframes == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, kVitScratchHeader, stream);
    if (e != cudaSuccess) return e;
    const uint32_t ew = (erasure & 0xFFu) * 0x01010101u;
    if (framebits % 32 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3) == 0)
        viterbi_pair_kernel<true, true><<<grid_blocks, kVitThreads, 0, stream>>>(d_rx, d_out, (uint8_t*)d_scratch, nframes, framebits,
                                                                                  (const uint4*)d_ptab, rx_per_frame, ew, d_last_row);
    else
        viterbi_pair_kernel<false, true><<<grid_blocks, kVitThreads, 0, stream>>>(d_rx, d_out, (uint8_t*)d_scratch, nframes, framebits,
                                                                                   (const uint4*)d_ptab, rx_per_frame, ew, d_last_row);
    count_launch();
    return cudaGetLastError();
}

// keep[4*(F+6)] (non-zero = transmitted) -> the per-iteration table of the kPunct kernel: (F+6)/2 entries of 4 words
void punct_table(uint32_t framebits, const uint8_t* keep, uint32_t* table) {
    const uint32_t iters = (framebits + 6) / 2;
    uint32_t off = 0;
    for (uint32_t it = 0; it < iters; it++) {
        const uint8_t* k = keep + 8 * (size_t)it;
        uint32_t sel[2] = {0u, 0u}, c[2] = {0u, 0u};
        for (int w = 0; w < 2; w++)
            for (int j = 0; j < 4; j++)  // selector nibble: byte c of the (shifted) run, or 4 = byte 0 of the erasure word
                sel[w] |= (k[4 * w + j] ? c[w]++ : 4u) << (4 * j);
        table[4 * it + 0] = off;
        table[4 * it + 1] = 8u * c[0];
        table[4 * it + 2] = sel[0] | (sel[1] << 16);
        table[4 * it + 3] = 0u;
        off += c[0] + c[1];
    }
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
