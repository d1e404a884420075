/* viterbi_b200.h -- C ABI of libviterbi_b200.so, a B200 (sm_100a) replacement for the FEC hot
 * path of Drehrumbum/viterbi.dll.
 *
 * Two groups of entry points:
 *   1. the reference's export surface (viterbi.def:4-8), same names, argument meaning and return
 *      conventions, so a host that does LoadLibrary/GetProcAddress (viterbi-benchmark.cpp:201-229)
 *      or dlopen/dlsym can switch libraries without code changes;
 *   2. batched entry points (new): many DAB frames / DAB+ superframes per launch, with host-pointer
 *      and device-pointer flavours.  This is the path that is measured.
 *
 * All functions are thread-safe.  There is no CPU fallback: every decode runs on the selected
 * CUDA device, and failures are reported through the return value (see "save mode").
 */
#ifndef VITERBI_B200_H
#define VITERBI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITERBI_B200_MAX_FRAMEBITS 9216u /* decision array bound of the reference, deconvolve.cpp:127 */

/* ---------------------------------------------------------------------------------------------
 * 1. Drop-in surface (viterbi.def:4-8)
 * ------------------------------------------------------------------------------------------- */

/* Replaces deconvolve(), deconvolve.cpp:551-554 (caller typedef viterbi-benchmark.cpp:72-73).
 * Hard-output Viterbi decode of one K=7 rate-1/4 DAB frame.
 *   framebits   info bits F (even, <= 9216); the frame carries F+6 trellis steps.  (The reference computes
 *               nbits = (F+6)/2 two-step iterations, deconvolve.cpp:126, so an odd F silently decodes F-1
 *               steps' worth of symbols and reads decisions it never wrote; its decision array ends at
 *               F = 9216, deconvolve.cpp:127.  Both cases are rejected here with return value 1, without
 *               entering save mode.)
 *   piData      4*(F+6) words, one soft symbol per word; only the low byte is used
 *               (deconvolve.cpp:219-228; README.md:19 out-of-range symbols)
 *   inputLength ignored, as in the reference
 *   output      ceil(F/8) bytes, MSB first
 * Returns 0; returns 1 on bad arguments / device failure and from then on until initialize()
 * ("save mode": exc_handler.cpp:204,214, viterbi_helpers.asm:183-186). */
int deconvolve(unsigned int framebits, unsigned int* piData, int inputLength, unsigned char* output);

/* Replaces RScheckSuperframe(), rschecksf.cpp:65-93 (caller typedef viterbi-benchmark.cpp:76-77).
 * DAB+ superframe: RSDims column-interleaved RS(120,110) codewords, byte k of codeword j at
 * p[j + k*RSDims].  Writes the 110 data bytes of each codeword to outVector with the same striding
 * and returns the total number of corrected symbols; on the first uncorrectable codeword returns -1
 * and leaves that column and all later ones untouched (rschecksf.cpp:80-88).  startIx is unused
 * (rschecksf.cpp:69).  Returns -1 on device failure as well (exc_handler.cpp:116-124,208-211).
 * p and outVector may be the same buffer (the reference copies each column before it writes:
 * rschecksf.cpp:75-84).  RSDims is limited to 1..1024 (one superframe must fit a 120 KB shared-memory
 * tile; DAB+ uses at most 24) -- larger values, which the reference's loop would accept, return -1;
 * RSDims == 0 returns 0 like the reference's empty loop. */
int RScheckSuperframe(unsigned char* p, int startIx, unsigned int RSDims, unsigned char* outVector);

/* Same function under the spelling BASELINE.json uses. */
int RSCheckSuperframe(unsigned char* p, int startIx, unsigned int RSDims, unsigned char* outVector);

/* Replaces initialize(), dllmain.cpp:156-160: called on every receiver start.  Re-reads the
 * configuration (here: environment VITERBI_B200_DEVICE, VITERBI_B200_LOG), clears save mode and
 * probes the device: after a sticky CUDA error the context is reset (cudaDeviceReset) and set up
 * again, and every thread's staging state of the old context is discarded on its next call -- the
 * recover-on-initialize contract of exc_handler.cpp:214 / dllmain.cpp:156.  Returns non-zero (true)
 * on success like the reference. */
int initialize(void);

/* Replaces GetCPUCaps(), viterbi_helpers.asm:48-157 / getcpucaps.h:27-38.  The CPU dispatcher is
 * replaced by device selection: returns 0 (no CPU decoder variants exist in this library). */
int GetCPUCaps(void);

/* Replaces WakeUpYMM(), dllmain.cpp:54-56: no-op. */
void WakeUpYMM(void);

/* ---------------------------------------------------------------------------------------------
 * 2. Batched entry points (new)
 * ------------------------------------------------------------------------------------------- */

/* Return codes of the batched API */
#define FEC_OK 0
#define FEC_ERR_ARG 1    /* NULL pointer, odd framebits, framebits > 9216, RSDims == 0 ... */
#define FEC_ERR_DEVICE 2 /* CUDA failure; fec_last_error() has the text */

/* n frames, one byte per soft symbol: syms [n][4*(F+6)], out [n][ceil(F/8)].  Host pointers
 * (pageable or pinned; pinned memory from fec_host_alloc() makes the copies asynchronous). */
int viterbi_deconvolve_batch(unsigned int framebits, const uint8_t* syms, size_t n, uint8_t* out);

/* Same, QIRX layout: one uint32 per soft symbol, low byte used. */
int viterbi_deconvolve_batch_u32(unsigned int framebits, const uint32_t* syms, size_t n, uint8_t* out);

/* Device-pointer flavours: buffers already in HBM, work enqueued on `stream` (a cudaStream_t, NULL =
 * default stream), no synchronisation.  Alignment: d_syms 8 bytes (u8 layout) or 16 bytes (u32 layout: one
 * trellis step per 16-byte load); d_out any (a 4-byte aligned d_out gets 32-bit stores when F % 32 == 0).
 * Rows are dense, so with F % 2 == 0 every row of an aligned d_syms is aligned too. */
int viterbi_deconvolve_batch_device(unsigned int framebits, const uint8_t* d_syms, size_t n, uint8_t* d_out,
                                    void* stream);
int viterbi_deconvolve_batch_u32_device(unsigned int framebits, const uint32_t* d_syms, size_t n,
                                        uint8_t* d_out, void* stream);

/* Depuncturing front end (the step a receiver performs immediately before
 * deconvolve(), outside viterbi.dll): rx [n][rx_per_frame] holds only the soft symbols that were transmitted;
 * keep [4*(F+6)] (host memory, one byte per mother-code symbol, non-zero = transmitted) is the puncturing
 * pattern of a whole frame, the same for every frame of the batch; punctured positions are filled with
 * `erasure` (0..255, normally the midpoint 128) on the device and the result is decoded exactly like
 * viterbi_deconvolve_batch() would decode the expanded symbols.  The number of non-zero bytes in keep must
 * equal rx_per_frame.  Up to 4x fewer bytes cross PCIe than with the expanded layout. */
int viterbi_deconvolve_batch_punctured(unsigned int framebits, const uint8_t* rx, size_t rx_per_frame,
                                       const uint8_t* keep, unsigned int erasure, size_t n, uint8_t* out);
/* d_rx / d_out in HBM, keep still a host pointer (it is turned into an index table and uploaded on `stream`). */
int viterbi_deconvolve_batch_punctured_device(unsigned int framebits, const uint8_t* d_rx, size_t rx_per_frame,
                                              const uint8_t* keep, unsigned int erasure, size_t n, uint8_t* d_out,
                                              void* stream);

/* n superframes with the same RSDims: in [n][120*RSDims], out [n][110*RSDims], ret [n].
 * Per superframe identical to RScheckSuperframe(), including the partial-write rule: bytes of out
 * belonging to the first failing column and later ones are not written (they keep the caller's values).
 * in and out must not overlap (n > 1 superframes have different strides in the two arrays).
 * Host flavour: when out is pinned memory (fec_host_alloc, cudaMallocHost, cudaHostRegister) the caller's bytes
 * of failing superframes are fetched by the kernel through the buffer's device mapping; a pageable out is
 * uploaded first (both give identical results; the pinned path moves about half the bytes). */
int rs_check_superframe_batch(const uint8_t* in, unsigned int RSDims, size_t n, uint8_t* out, int32_t* ret);
int rs_check_superframe_batch_device(const uint8_t* d_in, unsigned int RSDims, size_t n, uint8_t* d_out,
                                     int32_t* d_ret, void* stream);

/* DAB+ audio pipeline on the device: QIRX decodes five logical frames with
 * deconvolve() and then hands the 5 * F/8 = 120 * s bytes to RScheckSuperframe() (exc_handler.cpp:34-36).
 * These calls do both steps for nsf superframes without the host round trip in between:
 * syms [nsf * 5][4*(F+6)] soft symbols -> Viterbi -> [nsf][120*s] superframes (s = F/192) -> RS check
 * -> out [nsf][110*s] (same partial-write rule as RScheckSuperframe), ret [nsf].  framebits must be a
 * multiple of 192.  (The DAB energy-dispersal descrambler that sits between the two calls in a full
 * receiver is not part of viterbi.dll; it is applied only when fec_set_energy_dispersal(1) was called.) */
int dabplus_decode_superframes(unsigned int framebits, const uint8_t* syms, size_t nsf, uint8_t* out, int32_t* ret);
int dabplus_decode_superframes_device(unsigned int framebits, const uint8_t* d_syms, size_t nsf, uint8_t* d_out,
                                      int32_t* d_ret, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 3. Multi-device calls (new): one host batch decoded on all selected GPUs from ONE process
 * -------------------------------------------------------------------------------------------
 * Frames and superframes are independent (deconvolve.cpp:116-132 re-initialises the path metrics on every
 * call, rschecksf.cpp:72 keeps its scratch on the stack), so the batch is cut into contiguous shards (64-frame /
 * whole-superframe aligned), one per device of fec_set_devices(), and every shard runs the single-device
 * host-pointer path on a persistent worker thread of that device: own streams, own staging buffers, no
 * exchange between devices.  Results are identical to the single-device calls.  Pinned buffers recommended. */
int viterbi_deconvolve_batch_multi(unsigned int framebits, const uint8_t* syms, size_t n, uint8_t* out);
int rs_check_superframe_batch_multi(const uint8_t* in, unsigned int RSDims, size_t n, uint8_t* out, int32_t* ret);
int dabplus_decode_superframes_multi(unsigned int framebits, const uint8_t* syms, size_t nsf, uint8_t* out, int32_t* ret);

/* Devices used by the *_multi calls and fec_allgather_device(): `count` ordinals (distinct), or count == 0 for
 * "all visible devices" (the default). */
int fec_set_devices(const int* ordinals, int count);
/* Copies up to `capacity` selected ordinals to `ordinals` (may be NULL) and returns how many are selected. */
int fec_get_devices(int* ordinals, int capacity);

/* The one exchange of the design: gather device-resident result arrays over NVLink, for
 * hosts that keep the shards on the GPUs.  Single process.  Shard i (bytes_per_shard bytes at d_shard[i], resident on
 * selected device i) arrives at offset i * bytes_per_shard of every d_all[j] (each d_all[j] holds
 * count * bytes_per_shard bytes on device j; d_shard[i] may be that very place in d_all[i]).  Enqueued on streams[i]
 * (cudaStream_t of device i; streams == NULL or an entry NULL = that device's default stream); the result is
 * complete once the caller has synchronised ALL of them.
 * Default transport: the copy engines (cudaMemcpyPeerAsync, peer access enabled on first use) -- no kernel, no SM
 * taken from the decoders, no dependency on NCCL.  VITERBI_B200_GATHER=nccl uses ncclAllGather instead (one
 * communicator per selected device, ncclCommInitAll on first use; NCCL is loaded at run time from libnccl.so.2 or
 * $VITERBI_B200_NCCL_LIB -- the library has no link-time dependency on it). */
int fec_allgather_device(const void* const* d_shard, void* const* d_all, size_t bytes_per_shard, void* const* streams);

/* Energy dispersal (ETSI EN 300 401 clause 10): a DAB transmitter XORs every logical frame with the PRBS of
 * X^9 + X^5 + 1 (restarted with all ones per frame) in front of the convolutional encoder, and a receiver removes it
 * between deconvolve() and RScheckSuperframe() -- QIRX does that itself, it is not part of viterbi.dll.  With this
 * option on (process-wide, default off) the chained dabplus_decode_superframes* calls remove it on the device, so
 * the chain can be fed real sub-channel symbols.  deconvolve / the batch Viterbi calls are never affected. */
int fec_set_energy_dispersal(int on);

/* Gather fused into the producing kernel: the same calls as rs_check_superframe_batch_device() /
 * dabplus_decode_superframes_device(), and every result byte and return value the RS kernel stores into d_out /
 * d_ret is ALSO stored into ncopies (<= 15) further buffers of the same layout.  With buffers of the peer GPUs
 * (single process: cudaMalloc + fec_enable_peer_access(); one process per GPU: CUDA IPC mappings) these stores
 * travel over NVLink while the kernel is still decoding, tile by tile, so every GPU ends up holding every GPU's
 * results without a collective after the kernel.  Each copy must be congruent to d_out modulo 4 bytes; copies
 * follow the partial-write rule like d_out itself (pre-fill them the same way).  Visibility on the peers follows
 * the usual rule: after this stream's work has completed (and the peers have synchronised with it). */
int rs_check_superframe_batch_device_bcast(const uint8_t* d_in, unsigned int RSDims, size_t n, uint8_t* d_out,
                                           int32_t* d_ret, uint8_t* const* d_out_copies, int32_t* const* d_ret_copies,
                                           int ncopies, void* stream);
int dabplus_decode_superframes_device_bcast(unsigned int framebits, const uint8_t* d_syms, size_t nsf, uint8_t* d_out,
                                            int32_t* d_ret, uint8_t* const* d_out_copies, int32_t* const* d_ret_copies,
                                            int ncopies, void* stream);
/* cudaDeviceEnablePeerAccess between every pair of the selected devices (single-process hosts). */
int fec_enable_peer_access(void);
/* One process per GPU: export a buffer from fec_device_alloc() as a 64-byte CUDA IPC handle, map a peer process's
 * buffer into this process for the calling thread's device (with peer access; returns NULL on failure), unmap it.
 * The exporter keeps the buffer allocated while peers have it mapped. */
#define FEC_IPC_HANDLE_BYTES 64
int fec_ipc_export(const void* d_ptr, unsigned char* handle /* [FEC_IPC_HANDLE_BYTES] */);
void* fec_ipc_import(const unsigned char* handle);
int fec_ipc_close(void* d_ptr);

/* ---------------------------------------------------------------------------------------------
 * Device selection and utilities (replace getcpucaps/setupdll per the design brief)
 * ------------------------------------------------------------------------------------------- */
int fec_device_count(void);
int fec_set_device(int ordinal); /* selects the CUDA device used by the calling process */
int fec_get_device(void);
/* Per-thread override of fec_set_device() (-1 = none): a host that drives several GPUs from several threads
 * with the device-pointer calls binds each thread to its GPU with this. */
int fec_set_thread_device(int ordinal);
int fec_in_save_mode(void);
const char* fec_last_error(void); /* thread-local text of the last failure, "" if none */

/* pinned host memory for asynchronous staging */
void* fec_host_alloc(size_t bytes);
void fec_host_free(void* p);

/* plain device memory helpers for hosts without their own CUDA runtime binding */
void* fec_device_alloc(size_t bytes);
void fec_device_free(void* p);
int fec_memcpy_h2d(void* d_dst, const void* src, size_t bytes);
int fec_memcpy_d2h(void* dst, const void* d_src, size_t bytes);
/* Asynchronous device-to-device copy on `stream` (cudaMemcpyAsync, unified addressing): source and destination may
 * live on different GPUs -- cudaMalloc'ed buffers of one process with fec_enable_peer_access(), or a peer process's
 * buffer mapped with fec_ipc_import().  The copy engines move the bytes over NVLink without occupying an SM, so a
 * host can push each shard of results into the peers' arrays while the next batch is being decoded: a third way to
 * gather (beside fec_allgather_device and the *_bcast calls), and the one that disturbs the decode kernels least. */
int fec_memcpy_d2d_async(void* d_dst, const void* d_src, size_t bytes, void* stream);
int fec_device_synchronize(void);

/* Viterbi kernel selection: 0 = automatic (warp-per-frame kernel below 8,192 frames per launch for F <= 1536 and
 * below 6,144 frames for longer frames -- the measured crossover -- the two-frames-per-thread throughput kernel
 * above), 1 = always the throughput kernel, 2 = always the warp-per-frame kernel.  Both are bit-exact; this exists
 * for tests and measurements. */
#define FEC_VITERBI_AUTO 0
#define FEC_VITERBI_PAIR 1
#define FEC_VITERBI_WARP 2
int fec_set_viterbi_kernel(int mode);

/* number of kernels this library has launched since load (for benchmark accounting) */
unsigned long long fec_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* VITERBI_B200_H */
