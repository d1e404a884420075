// rs_kernels.cu -- sm_100a kernel for the DAB+ superframe Reed-Solomon check.
//
// Replaces RScheckSuperframe / DECODE_RS (rschecksf.cpp:65-93, 199-377) for batches of
// superframes: RS(120,110) = RS(255,245) shortened by 135, GF(256) poly 0x11D, roots a^0..a^9,
// errors-only Berlekamp-Massey / Chien / Forney, including the reference's quirks (roots inside
// the virtual padding are counted but not applied, unreduced Forney exponent, no den==0 test,
// "first failing column aborts the superframe" partial-write rule).
//
// Layout: a block stages a tile of whole superframes in shared memory with coalesced 16-byte
// loads (this is also the column de-interleave: codeword j byte k sits at tile[k*s + j]), one
// thread decodes one codeword in place, a shared-memory atomicMin finds the first failing
// column of each superframe, and the 110*s data bytes are written back coalesced, masked by
// that column.
//
// The syndrome stage differs from the reference's 10 Horner chains (1190 table steps per
// codeword): the codeword is reduced modulo the generator polynomial with a byte-wide LFSR
// (one 16-byte table row per data byte).  The remainder is zero iff all ten syndromes are zero,
// so clean codewords stop there; otherwise S_i = rem(a^i) (100 table steps).  The syndromes are
// the same field elements either way, so every later stage sees the reference's values.
//
// The Chien search uses no tables at all: it is bit-sliced over 32 positions (rs_chien_bitsliced.h).
#include <cuda_runtime.h>

#include <cstdint>

#include "fec_internal.h"
#include "rs_decode.h"

namespace fec {

namespace {

using rsdec::CW;
using rsdec::DATA;
using rsdec::NN;
using rsdec::NROOTS;

struct RsTables {
    uint8_t ato[768];   // alpha^(i mod 255), dllmain.cpp:145-146
    uint8_t iof[256];   // log, log(0) = 255, dllmain.cpp:131-143
    uint4 lfsr[256];    // c * (g(x) - x^10): coefficients x^0..x^9 in bytes 0..9
};

// In global memory, not __constant__: a block stages them with tid-strided reads, which are coalesced loads
// here but 32 serialised fetches per warp from the constant cache.
__device__ RsTables c_tables;

// Block-shared copies of the tables.  They are file-scope __shared__ objects (not pointers handed down
// through function arguments) so that every lookup compiles to an LDS with an immediate offset.
__shared__ uint4 s_lfsr[256];
__shared__ uint8_t s_ato[768];
__shared__ uint8_t s_iof[256];
__shared__ int s_nfail;  // failing superframes of the current tile (host path that merges the caller's bytes)
// The per-codeword decoder lives in rs_decode.h (shared with the host check of the CPU suite); this is its device
// policy: tables in shared memory, warp votes.
struct DevicePolicy {
    static __device__ __forceinline__ uint32_t ato(uint32_t i) { return s_ato[i]; }
    static __device__ __forceinline__ uint32_t iof(uint32_t v) { return s_iof[v]; }
    static __device__ __forceinline__ uint4 lfsr(uint32_t c) { return s_lfsr[c]; }
    static __device__ __forceinline__ bool any(unsigned mask, bool p) { return __any_sync(mask, p) != 0; }
    static __device__ __forceinline__ unsigned max(unsigned mask, unsigned v) { return __reduce_max_sync(mask, v); }
};

}  // namespace

// Extra destinations of a launch: the kernel stores every result byte and return value it produces into each of
// them as well (same layout as out / ret).  With peer pointers this IS the gather of the result bitstreams
//  -- done by the producing kernel's own stores over NVLink, tile by tile, instead of by a
// collective that follows it.
struct RsCopies {
    uint8_t* out[kRsMaxCopies];
    int32_t* ret[kRsMaxCopies];
    int n;
};

// One block = `sf_per_block` whole superframes; static shared memory holds the tables (5 KB), dynamic:
//   [first_fail int x sf_per_block] [sum int x sf_per_block] [tile]
// 8 blocks (32 warps) per SM: the launch bound holds the kernel to 64 registers.  Left alone the bit-sliced Chien
// search takes 96 (degree <= 5) to 125 registers, i.e. 5 or 4 blocks; the spills the bound causes sit almost
// entirely in the rare high-degree searches.  Measured on the bench mix: 337 M superframes/s unbounded, 455 M at 7
// blocks, 464 M at 8, 457 M at 9, 400 M at 10.
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 8
#endif
// kCopies: the launch has extra destinations (RsCopies); the plain instantiation carries none of that code.
template <bool kCopies>
__global__ void __launch_bounds__(kRsThreads, RS_MIN_BLOCKS)
rs_superframe_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int32_t* __restrict__ ret,
                     const uint8_t* __restrict__ orig, unsigned long long nsf, uint32_t s, uint32_t sf_per_block,
                     const RsCopies copies) {
    extern __shared__ __align__(16) uint8_t smem[];
    int* s_fail = reinterpret_cast<int*>(smem);
    int* s_sum = s_fail + sf_per_block;
    uint8_t* tile = reinterpret_cast<uint8_t*>(s_sum + sf_per_block);
    tile += (16 - (reinterpret_cast<uintptr_t>(tile) & 15)) & 15;

    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < 256; i += blockDim.x) s_lfsr[i] = c_tables.lfsr[i];
    for (uint32_t i = tid; i < 768; i += blockDim.x) s_ato[i] = c_tables.ato[i];
    for (uint32_t i = tid; i < 256; i += blockDim.x) s_iof[i] = c_tables.iof[i];

    const size_t sf_in = (size_t)CW * s, sf_out = (size_t)DATA * s;
    const uint32_t inv_s = (uint32_t)((0x100000000ull + s - 1) / s);
    const unsigned long long nblk = (nsf + sf_per_block - 1) / sf_per_block;
    for (unsigned long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const unsigned long long sf0 = blk * sf_per_block;
        const uint32_t nloc = (uint32_t)((nsf - sf0 < sf_per_block) ? (nsf - sf0) : sf_per_block);
        const size_t bytes = sf_in * nloc;
        const uint8_t* src = in + sf0 * sf_in;
        __syncthreads();  // previous tile fully written out / tables loaded
        // ---- stage the tile: coalesced 8-byte loads (120*s*sf0 is always a multiple of 8) ----------
        if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
            const size_t nvec = bytes / 8;  // bytes is a multiple of 8 as well
            for (size_t i = tid; i < nvec; i += blockDim.x)
                reinterpret_cast<uint2*>(tile)[i] = __ldg(reinterpret_cast<const uint2*>(src) + i);
        } else {
            for (size_t i = tid; i < bytes; i += blockDim.x) tile[i] = __ldg(src + i);
        }
        for (uint32_t i = tid; i < nloc; i += blockDim.x) {
            s_fail[i] = (int)s;  // "no column failed"
            s_sum[i] = 0;
        }
        __syncthreads();
        // ---- decode: one thread per codeword, corrections applied in place in the tile ------------
        const uint32_t ncw = nloc * s;
        for (uint32_t c0 = tid & ~31u; c0 < ncw; c0 += blockDim.x) {  // warp-uniform trip count
            const uint32_t c = c0 + (tid & 31u);
            const unsigned mask = __ballot_sync(0xffffffffu, c < ncw);
            if (c < ncw) {
                const uint32_t n = c / s, j = c - n * s;
                const int r = rsdec::rs_decode_column<DevicePolicy>(tile + n * sf_in + j, s, mask);
                if (r < 0)
                    atomicMin(&s_fail[n], (int)j);
                else if (r > 0)
                    atomicAdd(&s_sum[n], r);
            }
        }
        __syncthreads();
        // ---- return values: sum of the per-column counts, or -1 (rschecksf.cpp:80-88) -------------
        for (uint32_t n = tid; n < nloc; n += blockDim.x) {
            const int32_t r = (s_fail[n] < (int)s) ? -1 : s_sum[n];
            ret[sf0 + n] = r;
            for (int c = 0; kCopies && c < copies.n; c++) copies.ret[c][sf0 + n] = r;
        }
        // ---- write back the 110*s data bytes; columns >= first failure stay untouched -------------
        uint8_t* dst = out + sf0 * sf_out;
        for (uint32_t n = 0; n < nloc; n++) {
            const uint32_t fail = (uint32_t)s_fail[n];
            const uint8_t* t = tile + n * sf_in;
            uint8_t* d = dst + n * sf_out;
            if (fail == s) {
                // contiguous copy: byte head up to 4-byte alignment of d, then 32-bit words whose
                // source bytes are gathered from two aligned shared-memory words
                const uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(d) & 3)) & 3);
                const uint32_t nhead = head < sf_out ? head : (uint32_t)sf_out;
                const size_t doff = (size_t)(d - out);  // the same bytes go to every extra copy (peer buffers over NVLink)
                if (tid < nhead) {
                    d[tid] = t[tid];
                    for (int c = 0; kCopies && c < copies.n; c++) copies.out[c][doff + tid] = t[tid];
                }
                const size_t nwords = (sf_out - nhead) / 4;
                const uint32_t toff = (uint32_t)(reinterpret_cast<uintptr_t>(t + nhead) & 3);
                const uint32_t* tw = reinterpret_cast<const uint32_t*>(t + nhead - toff);
                const uint32_t sel = 0x3210u + 0x1111u * toff;
                for (size_t w = tid; w < nwords; w += blockDim.x) {
                    const uint32_t v = __byte_perm(tw[w], tw[w + 1], sel);
                    reinterpret_cast<uint32_t*>(d + nhead)[w] = v;
                    for (int c = 0; kCopies && c < copies.n; c++) reinterpret_cast<uint32_t*>(copies.out[c] + doff + nhead)[w] = v;
                }
                for (size_t i = nhead + nwords * 4 + tid; i < sf_out; i += blockDim.x) {
                    d[i] = t[i];
                    for (int c = 0; kCopies && c < copies.n; c++) copies.out[c][doff + i] = t[i];
                }
            } else if (orig == nullptr) {
                // i % s through a multiply-high (s is launch-uniform; exact for i < 2^17, s <= 1024)
                if (fail > 0) {
                    const size_t doff = (size_t)(d - out);
                    for (uint32_t i = tid; i < (uint32_t)sf_out; i += blockDim.x)
                        if (i - __umulhi(i, inv_s) * s < fail) {
                            d[i] = t[i];
                            for (int c = 0; kCopies && c < copies.n; c++) copies.out[c][doff + i] = t[i];
                        }
                }
            }
        }
        if (orig != nullptr) {
            // Host path without an upload of the caller's outVector: `out` is a staging buffer that is copied
            // back whole, so every byte of a failing superframe's row is produced here -- decoded columns from
            // the tile, the untouched ones (rschecksf.cpp:85-88) from the caller's own bytes, read through the
            // device mapping of the caller's pinned buffer (aligned 32-bit loads: a warp asks for whole 128-byte
            // lines across PCIe).  All failing superframes of the tile are handled in ONE flat loop over
            // (superframe, 4-byte slot) pairs, four slots per thread at a time with their loads issued together:
            // a PCIe round trip costs ~2 us, and one at a time per superframe they would serialise.
            __syncthreads();  // s_sum has been consumed by the return values: it now holds the list of failing superframes
            if (tid == 0) s_nfail = 0;
            __syncthreads();
            for (uint32_t n = tid; n < nloc; n += blockDim.x)
                if ((uint32_t)s_fail[n] < s) s_sum[atomicAdd(&s_nfail, 1)] = (int)n;
            __syncthreads();
            const uint32_t slots = (uint32_t)sf_out / 4 + 2;  // per superframe: head bytes, the aligned words, tail bytes
            const uint32_t inv_slots = (uint32_t)((0x100000000ull + slots - 1) / slots);
            const uint32_t total = (uint32_t)s_nfail * slots;
            auto col = [&](uint32_t i) { return i - __umulhi(i, inv_s) * s; };
            constexpr int kBatch = 4;
            for (uint32_t f0 = tid; f0 < total; f0 += blockDim.x * kBatch) {
                uint32_t sn[kBatch], sj[kBatch], o0[kBatch], o1[kBatch];
#pragma unroll
                for (int q = 0; q < kBatch; q++) {  // stage 1: which slot, and its load(s) of the caller's bytes
                    const uint32_t fl = f0 + q * blockDim.x;
                    o0[q] = o1[q] = 0u;
                    sn[q] = 0xFFFFFFFFu;
                    if (fl < total) {
                        const uint32_t li = __umulhi(fl, inv_slots);
                        const uint32_t n = (uint32_t)s_sum[li], j = fl - li * slots;
                        const uint8_t* dd = dst + n * sf_out;
                        const uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dd) & 3)) & 3);
                        const uint32_t nhead = head < sf_out ? head : (uint32_t)sf_out;
                        const uint32_t nwords = (uint32_t)((sf_out - nhead) / 4);
                        sn[q] = n, sj[q] = j;
                        if (j >= 1 && j <= nwords) {
                            const uint8_t* o = orig + (sf0 + n) * sf_out + nhead;
                            const uint32_t ooff = (uint32_t)(reinterpret_cast<uintptr_t>(o) & 3);
                            const uint32_t* ow = reinterpret_cast<const uint32_t*>(o - ooff) + (j - 1);
                            // the second word is only touched when it holds a byte of this row (never past the caller's array)
                            o0[q] = __ldg(ow);
                            if (ooff) o1[q] = __ldg(ow + 1);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < kBatch; q++) {  // stage 2: merge and store
                    if (sn[q] == 0xFFFFFFFFu) continue;
                    const uint32_t n = sn[q], j = sj[q], fail = (uint32_t)s_fail[n];
                    const uint8_t* t = tile + n * sf_in;
                    uint8_t* dd = dst + n * sf_out;
                    const uint8_t* o = orig + (sf0 + n) * sf_out;
                    const uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dd) & 3)) & 3);
                    const uint32_t nhead = head < sf_out ? head : (uint32_t)sf_out;
                    const uint32_t nwords = (uint32_t)((sf_out - nhead) / 4);
                    if (j == 0) {
                        for (uint32_t i = 0; i < nhead; i++) dd[i] = col(i) < fail ? t[i] : __ldg(o + i);
                    } else if (j <= nwords) {
                        const uint32_t w = j - 1, i0 = nhead + 4 * w;
                        const uint32_t toff = (uint32_t)(reinterpret_cast<uintptr_t>(t + nhead) & 3);
                        const uint32_t* tw = reinterpret_cast<const uint32_t*>(t + nhead - toff);
                        const uint32_t ooff = (uint32_t)(reinterpret_cast<uintptr_t>(o + nhead) & 3);
                        const uint32_t tv = __byte_perm(tw[w], tw[w + 1], 0x3210u + 0x1111u * toff);
                        const uint32_t ov = __byte_perm(o0[q], o1[q], 0x3210u + 0x1111u * ooff);
                        const uint32_t keep = (col(i0) < fail ? 0x000000FFu : 0u) | (col(i0 + 1) < fail ? 0x0000FF00u : 0u) |
                                              (col(i0 + 2) < fail ? 0x00FF0000u : 0u) | (col(i0 + 3) < fail ? 0xFF000000u : 0u);
                        reinterpret_cast<uint32_t*>(dd + nhead)[w] = (tv & keep) | (ov & ~keep);
                    } else if (j == nwords + 1) {
                        for (uint32_t i = nhead + nwords * 4; i < (uint32_t)sf_out; i++) dd[i] = col(i) < fail ? t[i] : __ldg(o + i);
                    }
                }
            }
        }
    }
}

size_t rs_smem_bytes(uint32_t s, uint32_t sf_per_block) {
    // + 16 alignment slack + 8 so the word gather may read one aligned word past the tile
    return 2 * sizeof(int) * (size_t)sf_per_block + 16 + (size_t)CW * s * sf_per_block + 8;
}

uint32_t rs_superframes_per_block(uint32_t s) {
    uint32_t n = kRsThreads / s;
    if (n >= 2) n &= ~1u;  // keep tiles 16-byte aligned in global memory
    return n ? n : 1;
}

cudaError_t rs_upload_tables() {
    static RsTables h;
    uint8_t alpha[255];
    unsigned sr = 1;
    h.iof[0] = NN;
    for (unsigned i = 0; i < NN; i++) {
        h.iof[sr] = (uint8_t)i;
        alpha[i] = (uint8_t)sr;
        sr <<= 1;
        if (sr & 0x100u) sr ^= 0x11Du;
    }
    for (unsigned i = 0; i < 768; i++) h.ato[i] = alpha[i % NN];
    // generator g(x) = prod (x - alpha^i), i = 0..9; g[k] = coefficient of x^k
    uint8_t g[NROOTS + 1] = {1};
    auto mul = [&](uint8_t a, uint8_t b) -> uint8_t {
        if (!a || !b) return 0;
        return alpha[(h.iof[a] + h.iof[b]) % NN];
    };
    for (int i = 0; i < NROOTS; i++) {
        uint8_t nx[NROOTS + 1] = {0};
        for (int k = 0; k <= i; k++) {
            nx[k + 1] ^= g[k];
            nx[k] ^= mul(g[k], alpha[i]);
        }
        for (int k = 0; k <= i + 1; k++) g[k] = nx[k];
    }
    for (unsigned c = 0; c < 256; c++) {
        uint8_t row[16] = {0};
        for (int k = 0; k < NROOTS; k++) row[k] = mul((uint8_t)c, g[k]);
        uint32_t w[4];
        for (int q = 0; q < 4; q++)
            w[q] = row[4 * q] | (row[4 * q + 1] << 8) | (row[4 * q + 2] << 16) | ((uint32_t)row[4 * q + 3] << 24);
        h.lfsr[c] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return cudaMemcpyToSymbol(c_tables, &h, sizeof(h));
}

// opt-in to the largest tile (one superframe of kRsMaxDims codewords) once per device, see viterbi_configure_device()
cudaError_t rs_configure_device() {
    const int worst = (int)rs_smem_bytes(kRsMaxDims, 1);
    cudaError_t e = cudaFuncSetAttribute(rs_superframe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, worst);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(rs_superframe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, worst);
}

cudaError_t launch_rs_superframes(const uint8_t* d_in, uint8_t* d_out, int32_t* d_ret, const uint8_t* d_orig,
                                  unsigned long long nsf, uint32_t s, int num_sms, cudaStream_t stream,
                                  uint8_t* const* extra_out, int32_t* const* extra_ret, int nextra) {
    if (nsf == 0) return cudaSuccess;
    const uint32_t spb = rs_superframes_per_block(s);
    const size_t smem = rs_smem_bytes(s, spb);
    unsigned long long nblk = (nsf + spb - 1) / spb;
    // persistent grid: exactly the blocks that are resident at once, so each block stages the tables once
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rs_superframe_kernel<false>, kRsThreads, smem) != cudaSuccess ||
        per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 4;
    }
    const unsigned long long cap = (unsigned long long)num_sms * (unsigned)per_sm;
    if (nblk > cap) nblk = cap;
    RsCopies copies;
    copies.n = 0;
    for (int c = 0; c < nextra && c < kRsMaxCopies; c++) {
        copies.out[copies.n] = extra_out[c];
        copies.ret[copies.n] = extra_ret[c];
        copies.n++;
    }
    if (copies.n > 0)
        rs_superframe_kernel<true><<<(unsigned)nblk, kRsThreads, smem, stream>>>(d_in, d_out, d_ret, d_orig, nsf, s, spb, copies);
    else
        rs_superframe_kernel<false><<<(unsigned)nblk, kRsThreads, smem, stream>>>(d_in, d_out, d_ret, d_orig, nsf, s, spb, copies);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fec
