// viterbi_warp_kernel.cu -- the warp-per-frame Viterbi kernel (latency / small-batch path) of libviterbi_b200.so.
//
// Replaces Butterfly256 / Renormalize256 / ChainBack (deconvolve.cpp:334-387, 407-412, 416-435) for one frame per
// warp with the survivor decisions in shared memory.  It lives in its own translation unit because it wants the
// default ptxas optimisation level (the list scheduler issues the shuffle of a step ahead of the independent
// work that fills its latency), while viterbi_kernels.cu is built with ptxas -O1 for the pair kernel.
#include <cuda_runtime.h>

#include <cstdint>

#include "fec_internal.h"
#include "viterbi_pair_core.h"

namespace fec {

// ---------------------------------------------------------------------------------------------------
// Latency / small-batch kernel: one warp per frame, two path metrics per lane, survivor decisions as warp
// ballots in shared memory, traceback out of shared memory -- decisions never leave the SM.  This is the
// layout of the design brief, tuned for the length of the per-step dependency chain, which is all that matters
// when one warp runs alone on an SM sub-partition (the single-frame drop-in call):
//
//   * ONE shuffle per trellis step.  The 6 state bits are spread over 5 lane bits and 1 "slot" bit (which of the
//     lane's two metrics).  A butterfly needs the two old states that differ in state bit 5 in one lane and
//     produces the two new states that differ in bit 0, so after a butterfly the slot holds bit 0 and some lane
//     bit holds the new bit 5.  Instead of restoring a fixed layout (two shuffles of a packed pair plus
//     pack/unpack, as the first version of this kernel did), the assignment of state bits to lane bits rotates:
//     one shfl.xor across exactly that lane bit swaps it with the slot -- each lane keeps one of its two new
//     metrics and takes the other from its partner.  Before butterfly t lane bit l holds state bit
//     (l + t) mod 5, the exchange after it crosses lane bit 4 - (t mod 5), the period is 5 steps.
//     The butterfly index of a lane is rotl5(lane, t mod 5), so its branch mask is one of five per-lane constants.
//   * The butterfly is three packed instructions: the lane keeps {A, A} and {B, B} as u16x2 pairs, so
//     {sat(B+mm), sat(B+m)} is one VIADDMNMX.U16x2, {A+m, A+mm} one add, and the survivor select one
//     VIMNMX.U16x2 whose two predicate outputs are the two decisions (ballots).  The exchange sends the packed
//     result {N[2i], N[2i+1]} whole; a byte permute with a per-lane selector picks the kept and the received
//     half and duplicates them for the next step.
//   * The serial chain per step is add-min -> min -> shuffle -> permute (-> relu-add after odd steps); the branch
//     metric (a dot-product form of the reference's two-level pavgb, see branch_metric_dp), the ballots, the
//     renormalisation test and the decision store hang off it.
//   * Traceback runs in "lane coordinates" (which lane / slot holds the current state), where one step is a
//     rotate and a bit-select on decision words that the whole warp pre-rotates first, so that the decision bit
//     lands on the lane bit it replaces.  The decoded bits are collected afterwards by the whole block.
// The block has four warps: all of them stage the symbols (over PCIe when the input is the caller's pinned bounce
// buffer: enough loads in flight for one round trip) and collect the output; warp 0 alone runs the trellis.
// ~25 warp-instructions per trellis step against 6.5 per frame-step for the pair kernel, so it is used where the
// pair kernel cannot fill the machine: the single-frame drop-in call and batches below vit_warp_kernel_max_frames().
// ---------------------------------------------------------------------------------------------------
namespace {

// Branch metric of deconvolve.cpp:338-349, m = avg(avg(x0,x1), avg(x2,x3)) >> 2 with avg(a,b) = (a+b+1) >> 1,
// without the two-level rounding: with S = x0+x1+x2+x3 and q = lsb(x0^x1) + lsb(x2^x3) it equals
// (S + 2 + q) >> 4 (each first-level average rounds up exactly when its two bytes differ in parity), which is
// two byte dot-products.  Returned packed for the butterfly: m_lo = {m, 63 - m}, m_hi = {63 - m, m}.
// Written as one volatile asm block: this file is built with ptxas -O1, which keeps the order of the PTX, and
// volatile statements keep their source order in the PTX -- so where this call stands in warp_step() is where its
// eight instructions issue (in the shadow of the shuffle of the step before).
__device__ __forceinline__ void branch_metric_pair(uint32_t w, uint32_t xmask, uint32_t& m_lo, uint32_t& m_hi) {
    asm volatile("{.reg .b32 x, s, p;\n\t"
                 "xor.b32 x, %2, %3;\n\t"
                 "dp4a.u32.u32 s, x, 0x01010101, 2;\n\t"
                 "shr.u32 p, x, 8;\n\t"
                 "lop3.b32 p, p, 0x00010001, x, 0x48;\n\t"  // (p ^ x) & 0x00010001
                 "dp4a.u32.u32 s, p, 0x01010101, s;\n\t"
                 "shr.u32 s, s, 4;\n\t"
                 "mad.lo.u32 %0, s, 0xFFFF0001, 0x003F0000;\n\t"
                 "sub.u32 %1, 0x003F003F, %0;}\n\t"
                 : "=r"(m_lo), "=r"(m_hi)
                 : "r"(w), "r"(xmask));
}

constexpr uint32_t kTraceWarmup = 96;  // steps a speculative traceback walks before its segment (survivors merge within ~5 K = 35)

__device__ __forceinline__ uint32_t rotl5(uint32_t v, uint32_t r) { return ((v << r) | (v >> (5u - r))) & 31u; }

// The survivor select and the exchange of one step, in issue order: N = min(t, m) per 16-bit half with the two
// "the minimum is t" predicates (ties included; the same min + setp pattern as min_decide in viterbi_pair_core.h,
// one VIMNMX.U16x2 with two predicate outputs), the butterfly shuffle of N across lane bit kLaneBit, on odd
// steps the broadcast of lane 0's N (new state 0, for the renormalisation test), and the two ballots.
template <int kLaneBit, bool kOdd>
__device__ __forceinline__ void select_exchange(uint32_t t, uint32_t m, uint32_t& N, uint32_t& R, uint32_t& N0, uint32_t& b_lo,
                                                uint32_t& b_hi) {
    if (kOdd)
        asm volatile("{.reg .pred pu, pv; .reg .u16 rs0, rs1, rs2, rs3;\n\t"
                     "min.u16x2 %0, %5, %6;\n\t"
                     "mov.b32 {rs0, rs1}, %0;\n\t"
                     "mov.b32 {rs2, rs3}, %5;\n\t"
                     "setp.eq.u16 pv, rs0, rs2;\n\t"
                     "setp.eq.u16 pu, rs1, rs3;\n\t"
                     "shfl.sync.bfly.b32 %1, %0, %7, 0x1f, 0xffffffff;\n\t"
                     "shfl.sync.idx.b32 %2, %0, 0, 0x1f, 0xffffffff;\n\t"
                     "vote.sync.ballot.b32 %3, pv, 0xffffffff;\n\t"
                     "vote.sync.ballot.b32 %4, pu, 0xffffffff;}\n\t"
                     : "=r"(N), "=r"(R), "=r"(N0), "=r"(b_lo), "=r"(b_hi)
                     : "r"(t), "r"(m), "n"(1 << kLaneBit));
    else
        asm volatile("{.reg .pred pu, pv; .reg .u16 rs0, rs1, rs2, rs3;\n\t"
                     "min.u16x2 %0, %4, %5;\n\t"
                     "mov.b32 {rs0, rs1}, %0;\n\t"
                     "mov.b32 {rs2, rs3}, %4;\n\t"
                     "setp.eq.u16 pv, rs0, rs2;\n\t"
                     "setp.eq.u16 pu, rs1, rs3;\n\t"
                     "shfl.sync.bfly.b32 %1, %0, %6, 0x1f, 0xffffffff;\n\t"
                     "vote.sync.ballot.b32 %2, pv, 0xffffffff;\n\t"
                     "vote.sync.ballot.b32 %3, pu, 0xffffffff;}\n\t"
                     : "=r"(N), "=r"(R), "=r"(b_lo), "=r"(b_hi)
                     : "r"(t), "r"(m), "n"(1 << kLaneBit));
}

// One trellis step of the warp kernel.  kPhase = t mod 5, kOdd = t & 1 (the renormalisation test follows odd
// steps: deconvolve.cpp:407-412).  AA = {A, A}, BB = {B, B}: the lane's old states with state bit 5 = 0 / 1.
// m_lo / m_hi: this step's packed branch metrics on entry, the NEXT step's (from w_next, xmask_next) on return --
// they are computed between the shuffle and its first consumer, where the warp would otherwise wait.
// selA / selB: byte-permute selectors of this lane for this phase (see the kernel).
template <int kPhase, bool kOdd>
__device__ __forceinline__ void warp_step(uint32_t& AA, uint32_t& BB, uint32_t& m_lo, uint32_t& m_hi, uint32_t w_next,
                                          uint32_t xmask_next, uint32_t selA, uint32_t selB, volatile uint2* dec_slot) {
    // ACS (deconvolve.cpp:352-359); ties choose the upper predecessor (decision = 1)
    const uint32_t T = __viaddmin_u16x2(BB, m_hi, 0x00FF00FFu);  // {sat(B + mm), sat(B + m)}
    const uint32_t M = AA + m_lo;                                 // {A + m, A + mm}
    uint32_t N, R, N0 = 0, be, bo;
    select_exchange<4 - kPhase, kOdd>(T, M, N, R, N0, be, bo);  // N = {N[2i], N[2i+1]}, R = the partner's
    branch_metric_pair(w_next, xmask_next, m_lo, m_hi);
    // decision of the new state in (lane l, low / high half) is bit l of be / bo (all lanes store the same words)
    *reinterpret_cast<volatile unsigned long long*>(dec_slot) = (unsigned long long)be | ((unsigned long long)bo << 32);
    AA = __byte_perm(N, R, selA);
    BB = __byte_perm(N, R, selB);
    if (kOdd) {  // Renormalize256: metric[state 0] > 150 -> all metrics -= 63, clamped at 0
        const uint32_t neg = (N0 & 0xFFFFu) > 150u ? 0xFFC1FFC1u : 0u;  // new state 0 lives in lane 0, low half
        AA = __viaddmax_s16x2_relu(AA, neg, 0u);
        BB = __viaddmax_s16x2_relu(BB, neg, 0u);
    }
}

// One traceback step in lane coordinates: (ln, w) = lane number of the current state and the decision word of
// its slot, pre-rotated left by kPrevBit.  The predecessor keeps the lane (a butterfly is lane-local) with slot =
// the decision; undoing the exchange before the butterfly swaps that slot with lane bit kPrevBit.
template <int kPrevBit>
__device__ __forceinline__ void warp_trace_step(uint32_t& ln, uint32_t& w, const uint2 next, uint32_t* x_slot) {
    const uint32_t x = __funnelshift_r(w, w, ln);       // decision bit of lane ln -> bit kPrevBit
    w = (ln >> kPrevBit) & 1u ? next.y : next.x;         // slot of the predecessor = the lane bit it replaces
    *x_slot = x;
    ln = (ln & ~(1u << kPrevBit)) | (x & (1u << kPrevBit));
}

}  // namespace

// done_flag: when not null (single-frame drop-in on the caller's pinned bounce buffer) the block writes 1 there
// after its output bytes are visible system-wide; the host polls the flag instead of synchronising the stream.
// kThreads: 128 (latency shape: four warps stage / trace / collect) or 32 (throughput shape: four times as many
// frames resident per SM).  A template parameter, not blockDim.x: with a run-time block size ptxas scheduled the
// forward loop differently and the single-frame call lost 12 % (36.6 -> 40.1 us at F = 768).
template <int kThreads>
__global__ void __launch_bounds__(kThreads) viterbi_warp_kernel(const uint8_t* __restrict__ syms, uint8_t* __restrict__ out,
                                                                  unsigned long long nframes, uint32_t framebits,
                                                                  uint32_t* done_flag) {
    extern __shared__ __align__(16) uint8_t wsmem[];
    constexpr uint32_t nthreads = kThreads;
    const uint32_t steps = framebits + 6, tid = threadIdx.x, lane = tid & 31u;
    uint32_t* s_sym = reinterpret_cast<uint32_t*>(wsmem);  // [steps + 2] 4 symbols per step (+ a readable pad word, never used); reused by the traceback
    uint2* s_dec = reinterpret_cast<uint2*>(wsmem + 4 * (size_t)(steps + 2));  // [steps] {even, odd} ballots
    const size_t outbytes = (framebits + 7) / 8;
    __shared__ uint32_t s_state[kThreads + 1];  // end state of every traceback segment (0 = no segment)
    if (tid == 0) s_state[kThreads] = 0;

    // per-lane constants of the five phases: the branch mask (const.asm:35-49 restated: 0xFF where the expected
    // code bit is 1) of butterfly rotl5(lane, phase), and the selectors that build {A, A} / {B, B} from the lane's
    // own result N and the partner's R: lane bit 4 - phase clear -> A = N.lo (kept), B = R.lo; set -> A = R.hi, B = N.hi
    uint32_t xm[5], sa[5], sb[5];
#pragma unroll
    for (uint32_t r = 0; r < 5; r++) {
        const uint32_t i = r ? rotl5(lane, r) : lane;
        xm[r] = (parity8((2u * i) & kPoly(0)) ? 0xFF0000FFu : 0u) |  // polys 0 and 3 coincide
                (parity8((2u * i) & kPoly(1)) ? 0x0000FF00u : 0u) | (parity8((2u * i) & kPoly(2)) ? 0x00FF0000u : 0u);
        const bool up = (lane >> (4u - r)) & 1u;
        sa[r] = up ? 0x7676u : 0x1010u;
        sb[r] = up ? 0x3232u : 0x5454u;
    }

    for (unsigned long long f = blockIdx.x; f < nframes; f += gridDim.x) {
        __syncthreads();
        {   // stage the frame: 8-byte loads, eight in flight per thread (over PCIe each round trip costs ~1.5 us)
            const uint2* row = reinterpret_cast<const uint2*>(syms + f * 4 * (size_t)steps);
            for (uint32_t i0 = tid; i0 < steps / 2; i0 += nthreads * 8) {
                uint2 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (i0 + nthreads * u < steps / 2) v[u] = __ldg(row + i0 + nthreads * u);
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (i0 + nthreads * u < steps / 2) reinterpret_cast<uint2*>(s_sym)[i0 + nthreads * u] = v[u];
            }
        }
        __syncthreads();

        if (tid < 32) {
            // ---- forward pass: ten steps per iteration (period 5 of the layout x period 2 of the renormalisation) ----
            // Locals256: M[0] = 0, others 63 (deconvolve.cpp:130-132)
            uint32_t AA = (lane == 0) ? 0u : 0x003F003Fu, BB = 0x003F003Fu;
            volatile uint2* dec = s_dec;
            uint32_t ml, mh;
            branch_metric_pair(s_sym[0], xm[0], ml, mh);
            uint32_t t = 0;
            for (; t + 10 <= steps; t += 10) {
                uint32_t w[10];  // the symbol words of steps t+1 .. t+10 (the last one belongs to the next group)
#pragma unroll
                for (int u = 0; u < 10; u++) w[u] = s_sym[t + 1 + u];
                warp_step<0, false>(AA, BB, ml, mh, w[0], xm[1], sa[0], sb[0], dec + t + 0);
                warp_step<1, true>(AA, BB, ml, mh, w[1], xm[2], sa[1], sb[1], dec + t + 1);
                warp_step<2, false>(AA, BB, ml, mh, w[2], xm[3], sa[2], sb[2], dec + t + 2);
                warp_step<3, true>(AA, BB, ml, mh, w[3], xm[4], sa[3], sb[3], dec + t + 3);
                warp_step<4, false>(AA, BB, ml, mh, w[4], xm[0], sa[4], sb[4], dec + t + 4);
                warp_step<0, true>(AA, BB, ml, mh, w[5], xm[1], sa[0], sb[0], dec + t + 5);
                warp_step<1, false>(AA, BB, ml, mh, w[6], xm[2], sa[1], sb[1], dec + t + 6);
                warp_step<2, true>(AA, BB, ml, mh, w[7], xm[3], sa[2], sb[2], dec + t + 7);
                warp_step<3, false>(AA, BB, ml, mh, w[8], xm[4], sa[3], sb[3], dec + t + 8);
                warp_step<4, true>(AA, BB, ml, mh, w[9], xm[0], sa[4], sb[4], dec + t + 9);
            }
            // the remaining 0, 2, ... 8 steps (steps is even): t is a multiple of 10 here
            auto sym_at = [&](uint32_t i) { return s_sym[i]; };  // i <= steps: the pad word at most
            if (t < steps) {
                warp_step<0, false>(AA, BB, ml, mh, sym_at(t + 1), xm[1], sa[0], sb[0], dec + t + 0);
                warp_step<1, true>(AA, BB, ml, mh, sym_at(t + 2), xm[2], sa[1], sb[1], dec + t + 1);
            }
            if (t + 2 < steps) {
                warp_step<2, false>(AA, BB, ml, mh, sym_at(t + 3), xm[3], sa[2], sb[2], dec + t + 2);
                warp_step<3, true>(AA, BB, ml, mh, sym_at(t + 4), xm[4], sa[3], sb[3], dec + t + 3);
            }
            if (t + 4 < steps) {
                warp_step<4, false>(AA, BB, ml, mh, sym_at(t + 5), xm[0], sa[4], sb[4], dec + t + 4);
                warp_step<0, true>(AA, BB, ml, mh, sym_at(t + 6), xm[1], sa[0], sb[0], dec + t + 5);
            }
            if (t + 6 < steps) {
                warp_step<1, false>(AA, BB, ml, mh, sym_at(t + 7), xm[2], sa[1], sb[1], dec + t + 6);
                warp_step<2, true>(AA, BB, ml, mh, sym_at(t + 8), xm[3], sa[2], sb[2], dec + t + 7);
            }
        }
        __syncthreads();
        // rotate the decision words left by the lane bit their step's preceding exchange crossed, (5 - t % 5) % 5,
        // so that the traceback's rotate-right by its lane number drops the decision onto that lane bit
        for (uint32_t tt = 6 + tid; tt < steps; tt += nthreads) {
            const uint32_t pb = (5u - tt % 5u) % 5u;
            const uint2 d = s_dec[tt];
            s_dec[tt] = make_uint2(__funnelshift_l(d.x, d.x, pb), __funnelshift_l(d.y, d.y, pb));
        }
        __syncthreads();

        // ---- ChainBack (deconvolve.cpp:416-435), from state 0 (lane 0, slot 0) after the last step; the rotated
        // decision word of every step replaces the symbol word of that step in s_sym.
        // The recursion is serial (9 ticks per step: a third of the frame's latency if one thread walks it all),
        // but survivor paths merge: a walk started kTraceWarmup steps further up from ANY state has almost always
        // joined the true path by the time it reaches its segment.  So every thread walks one short segment
        // speculatively, and the speculation is then VERIFIED end to end: the state a segment started from must be
        // the state the segment above it ended in, and the topmost segments start from the true end state.  Only
        // if every link holds is the result used; otherwise thread 0 walks the whole frame serially (inputs whose
        // survivors do not merge, e.g. constant symbols).  Either way the output is exactly the serial one.
        bool walked = false;
        if (framebits > 0) {
            // segment length: odd, so that the threads' 8-byte reads / 4-byte writes fall into distinct banks
            const uint32_t seg = ((framebits + nthreads - 1) / nthreads) | 1u;
            const uint32_t lo = 6 + tid * seg, hi = min(lo + seg, steps);  // this thread's steps [lo, hi)
            uint32_t st_in = 0, st_out = 0;
            if (lo < steps) {
                auto step = [&](uint32_t tt, uint32_t& ln, uint32_t& sl, uint32_t& w, bool store) {
                    const uint32_t pb = (5u - tt % 5u) % 5u;
                    const uint2 next = s_dec[tt > 6 ? tt - 1 : tt];
                    const uint32_t x = __funnelshift_r(w, w, ln);
                    sl = (ln >> pb) & 1u;
                    w = sl ? next.y : next.x;
                    if (store) s_sym[tt] = x;
                    ln = (ln & ~(1u << pb)) | (x & (1u << pb));
                };
                uint32_t tt = min(hi - 1 + kTraceWarmup, steps - 1);
                uint32_t ln = 0, sl = 0, w = s_dec[tt].x;  // exact when tt == steps - 1, a guess otherwise
                for (; tt >= hi; tt--) step(tt, ln, sl, w, false);
                st_in = ln | (sl << 5) | 0x100u;
                for (; tt >= lo; tt--) step(tt, ln, sl, w, true);
                st_out = ln | (sl << 5) | 0x100u;
            }
            s_state[tid] = st_out;
            __syncthreads();
            // the segment above (tid + 1) ended where this one began?  (segments that reach the top are exact)
            const bool last = lo >= steps || lo + seg >= steps;
            const bool ok = lo >= steps || last || hi - 1 + kTraceWarmup >= steps - 1 || s_state[tid + 1] == st_in;
            walked = __syncthreads_and(ok);
        }
        if (!walked && tid == 0 && framebits > 0) {
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
        __syncthreads();
        // decoded bit t = the decision consumed at step t + 6 = bit (5 - (t+6) % 5) % 5 of its rotated word;
        // output byte n holds bits 8n .. 8n+7, MSB first (missing bits of a ragged last byte stay 0)
        for (uint32_t n = tid; n < outbytes; n += nthreads) {
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
    if (done_flag != nullptr) {
        __threadfence_system();  // this thread's output bytes are visible to the host ...
        __syncthreads();         // ... and so are everybody else's, before the flag goes up
        if (tid == 0) {
            *reinterpret_cast<volatile uint32_t*>(done_flag) = 1u;
            __threadfence_system();
        }
    }
}

size_t viterbi_warp_smem_bytes(uint32_t framebits) { return 12 * (size_t)(framebits + 6) + 8 + 16; }

// The warp kernel needs more than 48 KB of dynamic shared memory above F = 4090.  The opt-in is a per-device
// function attribute, so it is raised once per device to the worst case (F = 9216) from the device
// initialisation in fec_api.cu -- not lazily per launch, where concurrent callers with different frame sizes
// would lower each other's limit.
cudaError_t viterbi_configure_device() {
    const int worst = (int)viterbi_warp_smem_bytes(kMaxFramebits);
    cudaError_t e = cudaFuncSetAttribute(viterbi_warp_kernel<kVitWarpThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, worst);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(viterbi_warp_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, worst);
}

cudaError_t launch_viterbi_warp(const uint8_t* d_syms, uint8_t* d_out, unsigned long long nframes, uint32_t framebits,
                                int num_sms, cudaStream_t stream, uint32_t* done_flag) {
    if (nframes == 0) return cudaSuccess;
    const size_t smem = viterbi_warp_smem_bytes(framebits);
    // Latency shape (a few frames): four warps per block, so that staging, the speculative traceback and the output
    // pass are spread wide.  Throughput shape (more frames than two per SM): one warp per block -- only warp 0 runs the
    // trellis, and a 32-thread block lets four times as many frames be resident per SM.
    const unsigned long long cap = (unsigned long long)num_sms * 32;
    const unsigned grid = (unsigned)(nframes < cap ? nframes : cap);
    if (nframes <= (unsigned long long)num_sms * 2)
        viterbi_warp_kernel<kVitWarpThreads><<<grid, kVitWarpThreads, smem, stream>>>(d_syms, d_out, nframes, framebits, done_flag);
    else
        viterbi_warp_kernel<32><<<grid, 32, smem, stream>>>(d_syms, d_out, nframes, framebits, done_flag);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fec
