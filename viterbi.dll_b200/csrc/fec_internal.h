// fec_internal.h -- declarations shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace fec {

// Viterbi throughput kernel launch shape: one warp per block (64 frames in flight per block), kVitMinBlocks resident
// blocks per SM.  The launch bound of 16 makes ptxas fit the ACS loop into 128 registers (one spill; the traceback's
// 32-record register ring spills a little instead), i.e. 4 warps per SM sub-partition; at 12 the kernel takes 155
// registers.  Measured 16 vs 12: FIC 114.0 vs 112.2, MSC 142.8 vs 141.7 Gbit/s.
constexpr int kVitThreads = 32;
#ifndef VIT_MIN_BLOCKS
#define VIT_MIN_BLOCKS 16
#endif
constexpr int kVitMinBlocks = VIT_MIN_BLOCKS;
constexpr size_t kVitScratchHeader = 256;  // ticket counter, keeps the decision area 256-byte aligned

// Batches below this size go to the warp-per-frame kernel (the pair kernel needs 64 frames per warp and takes a flat
// ~0.33 us x (F+6) for anything up to ~37,000 frames; the warp kernel sustains 24 Gbit/s at F=768 and 18.5 Gbit/s at
// F=3072).  Measured crossover (profiles/kernel_crossover_r02.jsonl): 8,192 frames at F=768, 6,144 at F=3072.
constexpr unsigned long long vit_warp_kernel_max_frames(uint32_t framebits) { return framebits <= 1536 ? 8192 : 6144; }
constexpr int kVitWarpThreads = 128;  // warp-per-frame kernel: warp 0 runs the trellis, all four stage / collect
constexpr int kRsThreads = 128;
constexpr uint32_t kMaxFramebits = 9216;  // decision array bound of the reference, deconvolve.cpp:127
constexpr int kRsMaxCopies = 15;           // extra destinations of one RS launch (peers of a 16-GPU box)
constexpr uint32_t kRsMaxDims = 1024;     // one superframe must fit a shared-memory tile (120 KB)

void count_launch();

size_t viterbi_scratch_bytes(int grid_blocks, uint32_t framebits);
int viterbi_grid_blocks(int num_sms, unsigned long long nframes, uint32_t framebits);
size_t viterbi_warp_smem_bytes(uint32_t framebits);
cudaError_t viterbi_configure_device();
cudaError_t launch_viterbi_pair(const uint8_t* d_syms, uint8_t* d_out, void* d_scratch, unsigned long long nframes,
                                uint32_t framebits, int grid_blocks, cudaStream_t stream);
// done_flag (optional, host-mapped): set to 1 by the kernel once its output is visible to the host
cudaError_t launch_viterbi_warp(const uint8_t* d_syms, uint8_t* d_out, unsigned long long nframes, uint32_t framebits,
                                int num_sms, cudaStream_t stream, uint32_t* done_flag = nullptr);
constexpr size_t kPunctSlackBytes = 16;  // readable bytes the fused depuncturing fetch may touch past the last row
cudaError_t launch_viterbi_pair_punctured(const uint8_t* d_rx, uint32_t rx_per_frame, const void* d_ptab, uint32_t erasure,
                                          uint8_t* d_out, void* d_scratch, unsigned long long nframes, uint32_t framebits,
                                          int grid_blocks, cudaStream_t stream, const uint8_t* d_last_row);
void punct_table(uint32_t framebits, const uint8_t* keep, uint32_t* table);  // (F+6)/2 entries of 4 words
cudaError_t launch_depuncture(const uint8_t* d_rx, size_t rx_per_frame, const int32_t* d_idx, uint32_t framebits,
                              uint32_t erasure, uint8_t* d_syms, size_t nframes, int num_sms, cudaStream_t stream);
cudaError_t descramble_upload_table();
cudaError_t launch_descramble(uint8_t* d_bits, size_t nframes, uint32_t framebits, int num_sms, cudaStream_t stream);
cudaError_t launch_compact_symbols(const uint32_t* d_in, uint8_t* d_out, size_t nsymbols, int num_sms,
                                   cudaStream_t stream);

size_t rs_smem_bytes(uint32_t s, uint32_t sf_per_block);
uint32_t rs_superframes_per_block(uint32_t s);
cudaError_t rs_upload_tables();
cudaError_t rs_configure_device();
// d_orig: nullptr = in-place semantics (d_out already holds the caller's bytes; untouched columns are not written);
// else every byte of d_out is written, untouched columns copied from d_orig (same layout as d_out).
// extra_out / extra_ret [nextra <= kRsMaxCopies]: further destinations (e.g. peer buffers) that receive the same
// stores as d_out / d_ret; their addresses must be congruent to d_out modulo 4.
cudaError_t launch_rs_superframes(const uint8_t* d_in, uint8_t* d_out, int32_t* d_ret, const uint8_t* d_orig,
                                  unsigned long long nsf, uint32_t s, int num_sms, cudaStream_t stream,
                                  uint8_t* const* extra_out = nullptr, int32_t* const* extra_ret = nullptr, int nextra = 0);

}  // namespace fec
