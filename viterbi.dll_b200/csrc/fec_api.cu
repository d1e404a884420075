// fec_api.cu -- the C ABI of libviterbi_b200.so (include/viterbi_b200.h) over the CUDA runtime.
//
// Mirrors the reference's export surface (viterbi.def:4-8) and its failure convention
// (exc_handler.cpp:204-214: after a fault deconvolve returns 1 and RScheckSuperframe -1 until
// initialize() is called), and adds the batched entry points.  The CPU dispatcher / ini file
// (setupdll.cpp) is replaced by device selection.  No CPU decode path exists here.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/viterbi_b200.h"
#include "fec_internal.h"

namespace fec {

namespace {

constexpr int kMaxDevices = 64;
constexpr int kPipe = 3;  // host-path pipeline depth (streams / staging slots)
constexpr size_t kBounceBytes = 256 * 1024;  // calls moving less than this bounce through pinned memory
constexpr size_t kTrimBytes = (size_t)768 << 20;  // staging buffers above this are released when the call returns
constexpr float kRsFetchMaxFailFrac = 0.55f;  // above this share of failing superframes outVector is uploaded instead
constexpr int kGraphCache = 4;  // instantiated single-kernel graphs kept per calling thread (drop-in path)

std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_save_mode{0};
std::atomic<int> g_device{-1};  // -1: use the calling thread's current device
std::atomic<int> g_vit_kernel{FEC_VITERBI_AUTO};
std::atomic<int> g_descramble{0};  // dabplus_* calls: remove the DAB energy dispersal between Viterbi and RS
std::atomic<unsigned> g_generation{1};  // bumped when initialize() had to reset a device: staging state of older generations is dead
thread_local std::string t_error;
thread_local int t_device = -1;  // per-thread device (fec_set_thread_device, multi-device workers); wins over g_device

struct DeviceState {
    std::mutex mu;
    bool ready = false;
    int num_sms = 0;
};
DeviceState g_dev[kMaxDevices];

// One staging slot of the host-pointer pipeline: device input / output / scratch on its own stream.
struct Slot {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* d_aux = nullptr;  // u32 / punctured staging (viterbi) or ret (rs)
    void* d_scratch = nullptr;
    void* h_pin = nullptr;  // pinned bounce buffer for the single-call drop-in path
    size_t in_cap = 0, out_cap = 0, aux_cap = 0, scratch_cap = 0, pin_cap = 0;
};

// A single-kernel CUDA graph of the drop-in decode on this thread's bounce buffer: the kernel
// arguments never change between calls of the same shape, so the launch is one cudaGraphLaunch of a
// pre-instantiated graph.
struct DropinGraph {
    cudaGraphExec_t exec = nullptr;
    unsigned framebits = 0;
    const void* in = nullptr;
    void* out = nullptr;
    void* flag = nullptr;
    unsigned long long last_use = 0;
};

// Staging state of the host-pointer calls.  One per calling thread: QIRX >= 4.0 calls deconvolve() from several
// threads at once (README.md:56), and with per-thread streams and buffers those calls overlap on the device
// instead of queueing behind one lock.  Released when the thread exits.
struct HostPipe {
    int device = -1;
    unsigned generation = 0;
    Slot slot[kPipe];
    void* d_idx = nullptr;  // depuncturing tables of the call in progress: [index table][per-iteration table]
    size_t idx_cap = 0;
    cudaEvent_t idx_ready = nullptr;
    DropinGraph graph[kGraphCache];
    unsigned long long graph_clock = 0;
    float rs_fail_frac = -1.0f;  // running failure fraction of this thread's RS batches (-1: nothing observed yet)
    ~HostPipe();
};
thread_local HostPipe g_pipe;

bool fail(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return false;
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    t_error = buf;
    (void)cudaGetLastError();  // clear the sticky-less error state
    return true;
}

int bad_arg(const char* what) {
    t_error = what;
    return FEC_ERR_ARG;
}

// Per-device one-time setup: GF tables, shared-memory opt-ins of both kernels (per-device function attributes,
// set here for the worst case so that concurrent callers never lower each other's limit), scratch pool.
// Not cached on failure: the next call (or initialize()) tries again.
bool init_device(DeviceState* st, int dev) {
    cudaDeviceProp prop;
    if (fail(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return false;
    st->num_sms = prop.multiProcessorCount;
    if (fail(rs_upload_tables(), "RS table upload") || fail(descramble_upload_table(), "PRBS table upload") || fail(viterbi_configure_device(), "viterbi kernel attributes") ||
        fail(rs_configure_device(), "rs kernel attributes"))
        return false;
    // the table upload ran on the legacy default stream; the kernels run on non-blocking streams, which do not
    // order themselves behind it
    if (fail(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return false;
    // keep stream-ordered scratch allocations cached in the pool between calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
    return true;
}

// Resolve the device this call runs on and make sure its per-device state exists.  The calling thread's current
// device is only changed when a device was selected explicitly and differs from it.
DeviceState* device_state(int* ordinal_out = nullptr) {
    int dev = t_device >= 0 ? t_device : g_device.load();
    int cur = -1;
    if (fail(cudaGetDevice(&cur), "cudaGetDevice")) return nullptr;
    if (dev < 0) dev = cur;
    if (dev < 0 || dev >= kMaxDevices) {
        t_error = "device ordinal out of range";
        return nullptr;
    }
    if (dev != cur && fail(cudaSetDevice(dev), "cudaSetDevice")) return nullptr;
    DeviceState* st = &g_dev[dev];
    {
        std::lock_guard<std::mutex> lock(st->mu);
        if (!st->ready) st->ready = init_device(st, dev);
        if (!st->ready) return nullptr;
    }
    if (ordinal_out) *ordinal_out = dev;
    return st;
}

bool grow(void** p, size_t* cap, size_t need, bool pinned = false) {
    if (need <= *cap) return true;
    if (*p) {
        if (pinned)
            cudaFreeHost(*p);
        else
            cudaFree(*p);
        *p = nullptr;
        *cap = 0;
    }
    const size_t want = need + need / 4;
    cudaError_t e = pinned ? cudaMallocHost(p, want) : cudaMalloc(p, want);
    if (fail(e, pinned ? "cudaMallocHost" : "cudaMalloc")) return false;
    *cap = want;
    return true;
}

void drop_graphs(HostPipe& pipe, bool destroy) {
    for (DropinGraph& g : pipe.graph) {
        if (g.exec && destroy) cudaGraphExecDestroy(g.exec);
        g = DropinGraph();
    }
}

// destroy = false: the context the handles belonged to is gone (device reset); just forget them
void release_pipe(HostPipe& pipe, bool destroy = true) {
    drop_graphs(pipe, destroy);
    for (Slot& s : pipe.slot) {
        if (destroy) {
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.d_aux) cudaFree(s.d_aux);
            if (s.d_scratch) cudaFree(s.d_scratch);
            if (s.h_pin) cudaFreeHost(s.h_pin);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        s = Slot();
    }
    if (destroy) {
        if (pipe.d_idx) cudaFree(pipe.d_idx);
        if (pipe.idx_ready) cudaEventDestroy(pipe.idx_ready);
    }
    pipe.d_idx = nullptr;
    pipe.idx_ready = nullptr;
    pipe.idx_cap = 0;
    pipe.device = -1;
    (void)cudaGetLastError();  // a thread that outlives the CUDA context frees nothing; that is fine
}

HostPipe::~HostPipe() {
    if (device < 0) return;
    if (generation != g_generation.load())
        release_pipe(*this, false);
    else if (cudaSetDevice(device) == cudaSuccess)
        release_pipe(*this);
}

bool prepare_pipe(int dev) {
    const unsigned gen = g_generation.load();
    if (g_pipe.device >= 0 && g_pipe.generation != gen) release_pipe(g_pipe, false);
    if (g_pipe.device != dev) {
        if (g_pipe.device >= 0) {
            cudaSetDevice(g_pipe.device);
            release_pipe(g_pipe);
            cudaSetDevice(dev);
        }
        g_pipe.device = dev;
        g_pipe.generation = gen;
    }
    for (Slot& s : g_pipe.slot)
        if (!s.stream && fail(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    if (!g_pipe.idx_ready && fail(cudaEventCreateWithFlags(&g_pipe.idx_ready, cudaEventDisableTiming), "cudaEventCreate"))
        return false;
    return true;
}

// Release staging buffers that an oversized call left behind (F = 9216 batches): they would otherwise stay
// allocated until the calling thread exits.
void trim_pipe() {
    for (Slot& s : g_pipe.slot) {
        auto trim = [](void** p, size_t* cap) {
            if (*cap > kTrimBytes) {
                cudaFree(*p);
                *p = nullptr;
                *cap = 0;
            }
        };
        trim(&s.d_in, &s.in_cap);
        trim(&s.d_out, &s.out_cap);
        trim(&s.d_aux, &s.aux_cap);
        trim(&s.d_scratch, &s.scratch_cap);
    }
}

// Wait for a stream with a short spin on cudaStreamQuery before blocking: the drop-in call lasts tens of
// microseconds, about what a blocking synchronise adds on its own.
bool wait_stream(cudaStream_t stream, const char* what) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t e = cudaStreamQuery(stream);
        if (e == cudaSuccess) return true;
        if (e != cudaErrorNotReady) return !fail(e, what);
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) break;
    }
    return !fail(cudaStreamSynchronize(stream), what);
}

// Host-path chunk size in frames per pipeline stage.  The throughput kernel takes a flat ~0.38 us x (F+6)
// for anything up to ~37,000 frames, the H2D copy of n frames takes n x 4(F+6) B / 55 GB/s, so the
// kernel hides behind the copy once a chunk holds more than ~5,300 frames whatever F is; 12,288 leaves a 2x margin
// at the FIC size (profiles/e2e_chunk_sweep.py).  Larger frames get proportionally fewer frames per chunk (about
// 38 MB of symbols, never below 2,048 frames) so that the staging buffers and the decision scratch of a slot stay
// bounded at F = 9216.  VITERBI_B200_CHUNK_FRAMES overrides it.
size_t host_chunk_frames(unsigned framebits) {
    static const long forced = [] {
        const char* env = getenv("VITERBI_B200_CHUNK_FRAMES");
        return (env && *env) ? atol(env) : 0L;
    }();
    if (forced >= 64) return (size_t)forced & ~(size_t)63;
    const size_t row = 4 * ((size_t)framebits + 6);
    size_t frames = ((size_t)12288 * 3096 / row) & ~(size_t)63;
    if (frames > 12288) frames = 12288;
    if (frames < 2048) frames = 2048;
    return frames;
}

// Host-path chunk size of the RS calls in input bytes per pipeline stage (VITERBI_B200_RS_CHUNK_MB overrides).
size_t rs_chunk_bytes() {
    static const size_t bytes = [] {
        const char* env = getenv("VITERBI_B200_RS_CHUNK_MB");
        const long v = (env && *env) ? atol(env) : 16;  // 32 / 16 / 8 / 4 MB measured: 41.3 / 42.9 / 41.8 / 40.2 M superframes/s
        return (size_t)(v >= 1 ? v : 16) << 20;
    }();
    return bytes;
}

bool vit_args_ok(unsigned framebits) { return !(framebits & 1u) && framebits <= VITERBI_B200_MAX_FRAMEBITS; }

// Does a launch of n frames go to the two-frames-per-thread throughput kernel (else: warp-per-frame kernel)?
// VITERBI_B200_PUNCT_SEPARATE=1: punctured input is expanded by the separate depuncturing kernel even where the
// throughput kernel could do it in its symbol fetch (A/B measurements)
bool punct_fused() {
    static const bool separate = [] {
        const char* env = getenv("VITERBI_B200_PUNCT_SEPARATE");
        return env && *env == '1';
    }();
    return !separate;
}

bool uses_pair_kernel(size_t n, unsigned framebits) {
    const int mode = g_vit_kernel.load();
    return mode == FEC_VITERBI_PAIR || (mode == FEC_VITERBI_AUTO && n >= vit_warp_kernel_max_frames(framebits));
}

// Enqueue one batch that is already in device memory.  Scratch is stream-ordered.
int vit_device(DeviceState* st, unsigned framebits, const uint8_t* d_syms, size_t n, uint8_t* d_out,
               cudaStream_t stream, void* scratch, size_t scratch_cap) {
    if (n == 0 || framebits == 0) return FEC_OK;
    if (!uses_pair_kernel(n, framebits))
        // latency / small-batch path: decisions stay in shared memory
        return fail(launch_viterbi_warp(d_syms, d_out, n, framebits, st->num_sms, stream), "viterbi warp kernel launch")
                   ? FEC_ERR_DEVICE
                   : FEC_OK;
    const int blocks = viterbi_grid_blocks(st->num_sms, n, framebits);
    const size_t need = viterbi_scratch_bytes(blocks, framebits);
    void* ws = scratch;
    const bool own = (ws == nullptr) || scratch_cap < need;
    if (own && fail(cudaMallocAsync(&ws, need, stream), "cudaMallocAsync(scratch)")) return FEC_ERR_DEVICE;
    cudaError_t e = launch_viterbi_pair(d_syms, d_out, ws, n, framebits, blocks, stream);
    if (own) (void)cudaFreeAsync(ws, stream);
    return fail(e, "viterbi kernel launch") ? FEC_ERR_DEVICE : FEC_OK;
}

// Punctured rows that are already in device memory (followed by kPunctSlackBytes readable bytes): the throughput
// kernel expands them in its symbol fetch.  d_ptab: punct_table() on the device.
int vit_device_punctured(DeviceState* st, unsigned framebits, const uint8_t* d_rx, size_t rx_per_frame, const void* d_ptab,
                         unsigned erasure, size_t n, uint8_t* d_out, cudaStream_t stream, void* scratch, size_t scratch_cap,
                         const uint8_t* d_last_row = nullptr) {
    const int blocks = viterbi_grid_blocks(st->num_sms, n, framebits);
    const size_t need = viterbi_scratch_bytes(blocks, framebits);
    void* ws = scratch;
    const bool own = (ws == nullptr) || scratch_cap < need;
    if (own && fail(cudaMallocAsync(&ws, need, stream), "cudaMallocAsync(scratch)")) return FEC_ERR_DEVICE;
    cudaError_t e = launch_viterbi_pair_punctured(d_rx, (uint32_t)rx_per_frame, d_ptab, erasure, d_out, ws, n, framebits, blocks, stream,
                                                  d_last_row);
    if (own) (void)cudaFreeAsync(ws, stream);
    return fail(e, "viterbi kernel launch") ? FEC_ERR_DEVICE : FEC_OK;
}

enum class SymFormat { U8, U32, Punctured };

// keep[4*(F+6)] (non-zero = transmitted) -> idx[p] = position of mother-code symbol p in a received row, -1 if
// punctured.  Returns false when the pattern does not account for exactly rx_per_frame symbols.
bool puncture_index(unsigned framebits, const uint8_t* keep, size_t rx_per_frame, std::vector<int32_t>& idx) {
    const size_t nsym = 4 * ((size_t)framebits + 6);
    idx.resize(nsym);
    int32_t next = 0;
    for (size_t p = 0; p < nsym; p++) idx[p] = keep[p] ? next++ : -1;
    return (size_t)next == rx_per_frame;
}

// Wait for the completion flag the warp kernel raises in the pinned bounce buffer (a PCIe write the host sees about a
// microsecond after the kernel's last store -- several microseconds earlier than a stream synchronise returns).  The
// stream is queried now and then so that a failed launch cannot leave the caller spinning.
bool wait_flag(volatile uint32_t* flag, cudaStream_t stream, const char* what) {
    for (unsigned spins = 1;; spins++) {
        if (*flag != 0u) {
            std::atomic_thread_fence(std::memory_order_acquire);
            return true;
        }
        if ((spins & 0x3FFu) == 0) {
            const cudaError_t e = cudaStreamQuery(stream);
            if (e == cudaSuccess) return true;  // kernel done: its stores, the flag included, are visible
            if (e != cudaErrorNotReady) return !fail(e, what);
        }
    }
}

// The single-frame drop-in decode on this thread's bounce buffer: one graph launch + one wait on the flag.
int dropin_launch(DeviceState* st, unsigned framebits, const uint8_t* in, uint8_t* out, uint32_t* flag, size_t n) {
    Slot& s0 = g_pipe.slot[0];
    DropinGraph* hit = nullptr;
    DropinGraph* victim = &g_pipe.graph[0];
    if (n == 1) {
        for (DropinGraph& g : g_pipe.graph) {
            if (g.exec && g.framebits == framebits && g.in == in && g.out == out && g.flag == flag) hit = &g;
            if (g.last_use < victim->last_use) victim = &g;
        }
    }
    *reinterpret_cast<volatile uint32_t*>(flag) = 0u;
    if (hit) {
        hit->last_use = ++g_pipe.graph_clock;
        count_launch();
        if (fail(cudaGraphLaunch(hit->exec, s0.stream), "cudaGraphLaunch")) return FEC_ERR_DEVICE;
        return wait_flag(flag, s0.stream, "drop-in decode") ? FEC_OK : FEC_ERR_DEVICE;
    }
    auto launch = [&]() { return launch_viterbi_warp(in, out, n, framebits, st->num_sms, s0.stream, flag); };
    if (n == 1) {
        // first call of this shape: capture the launch into a graph and keep the executable
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool ok = cudaStreamBeginCapture(s0.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const cudaError_t le = launch();
            const cudaError_t ce = cudaStreamEndCapture(s0.stream, &graph);
            ok = le == cudaSuccess && ce == cudaSuccess && graph != nullptr &&
                 cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
        }
        (void)cudaGetLastError();
        if (ok) {
            if (victim->exec) cudaGraphExecDestroy(victim->exec);
            *victim = DropinGraph{exec, framebits, in, out, flag, ++g_pipe.graph_clock};
            if (fail(cudaGraphLaunch(exec, s0.stream), "cudaGraphLaunch")) return FEC_ERR_DEVICE;
            return wait_flag(flag, s0.stream, "drop-in decode") ? FEC_OK : FEC_ERR_DEVICE;
        }
        // capture not possible (e.g. the host application is capturing globally): plain launch below
    }
    if (fail(launch(), "viterbi warp kernel launch")) return FEC_ERR_DEVICE;
    // a batch raises the flag once per block, so only the stream tells when all of them are done
    if (n > 1) return wait_stream(s0.stream, "drop-in decode") ? FEC_OK : FEC_ERR_DEVICE;
    return wait_flag(flag, s0.stream, "drop-in decode") ? FEC_OK : FEC_ERR_DEVICE;
}

// Host-pointer batch: chunks pipelined over kPipe streams (H2D | kernel | D2H overlap).
int vit_host(unsigned framebits, const void* syms, SymFormat fmt, size_t n, uint8_t* out, const uint8_t* keep = nullptr,
             size_t rx_per_frame = 0, unsigned erasure = 0) {
    if (!vit_args_ok(framebits)) return bad_arg("framebits must be even and <= 9216");
    if (n == 0 || framebits == 0) return FEC_OK;
    if (!syms || !out) return bad_arg("null pointer");
    const bool is_u32 = fmt == SymFormat::U32, punct = fmt == SymFormat::Punctured;
    std::vector<int32_t> idx;
    if (punct) {
        if (!keep) return bad_arg("null pointer");
        if (erasure > 255) return bad_arg("erasure must be 0..255");
        if (!puncture_index(framebits, keep, rx_per_frame, idx)) return bad_arg("keep pattern does not match rx_per_frame");
    }
    int dev;
    DeviceState* st = device_state(&dev);
    if (!st) return FEC_ERR_DEVICE;
    if (!prepare_pipe(dev)) return FEC_ERR_DEVICE;

    const size_t nsym = 4 * ((size_t)framebits + 6), nout = (framebits + 7) / 8;
    const size_t in_row = punct ? rx_per_frame : nsym * (is_u32 ? 4 : 1);
    const size_t idx_bytes = nsym * sizeof(int32_t), ptab_bytes = ((size_t)framebits + 6) / 2 * 16;
    std::vector<uint32_t> ptab;
    if (punct) {
        // both table forms go up on slot 0's stream; the other slots wait for the event (the vectors outlive the
        // call's final synchronise, so the pageable sources are safe)
        ptab.resize(ptab_bytes / 4);
        punct_table(framebits, keep, ptab.data());
        if (!grow(&g_pipe.d_idx, &g_pipe.idx_cap, idx_bytes + ptab_bytes) ||
            fail(cudaMemcpyAsync(g_pipe.d_idx, idx.data(), idx_bytes, cudaMemcpyHostToDevice, g_pipe.slot[0].stream), "H2D index table") ||
            fail(cudaMemcpyAsync((uint8_t*)g_pipe.d_idx + idx_bytes, ptab.data(), ptab_bytes, cudaMemcpyHostToDevice, g_pipe.slot[0].stream),
                 "H2D puncturing table") ||
            fail(cudaEventRecord(g_pipe.idx_ready, g_pipe.slot[0].stream), "cudaEventRecord"))
            return FEC_ERR_DEVICE;
        for (int k = 1; k < kPipe; k++)
            if (fail(cudaStreamWaitEvent(g_pipe.slot[k].stream, g_pipe.idx_ready, 0), "cudaStreamWaitEvent")) return FEC_ERR_DEVICE;
    }
    // Small calls (the single-frame drop-in above all) bounce through this thread's pinned buffer: the driver's
    // pageable-copy path serialises concurrent callers (4 threads at F=3072 were slower than 1), a 50 KB memcpy
    // into pinned memory does not.  QIRX's one-word-per-symbol layout is compacted to bytes by that very copy
    // (only the low byte counts: deconvolve.cpp:219-228), so a quarter of the bytes cross PCIe.
    uint8_t* bounce_out = nullptr;
    uint8_t* const user_out = out;
    const size_t in_bytes = n * in_row, out_bytes = n * nout;
    if (in_bytes + out_bytes <= kBounceBytes) {
        Slot& s0 = g_pipe.slot[0];
        const size_t in_keep = (is_u32 ? n * nsym : in_bytes), in_pad = (in_keep + 255) & ~(size_t)255;
        const size_t out_pad = (out_bytes + 3) & ~(size_t)3;
        if (s0.pin_cap < in_pad + out_pad + 4) drop_graphs(g_pipe, true);  // the graphs point into the old buffer
        if (!grow(&s0.h_pin, &s0.pin_cap, in_pad + out_pad + 4, true)) return FEC_ERR_DEVICE;
        if (is_u32) {
            const uint32_t* src = static_cast<const uint32_t*>(syms);
            uint8_t* dst = static_cast<uint8_t*>(s0.h_pin);
            for (size_t i = 0; i < n * nsym; i++) dst[i] = (uint8_t)src[i];
        } else if (syms != s0.h_pin) {  // (the u32 case below re-enters with the compacted buffer itself)
            memcpy(s0.h_pin, syms, in_bytes);
        }
        syms = s0.h_pin;
        bounce_out = (uint8_t*)s0.h_pin + in_pad;
        out = bounce_out;
        // ... and when the warp-per-frame kernel would decode them anyway, it runs directly on the bounce buffer:
        // pinned memory is mapped into the device's address space, the kernel stages the symbols into shared
        // memory itself, writes the decoded bytes back through the mapping and raises a completion flag behind
        // them, so the call is one (graph) launch and one poll -- no copy operations, no synchronise.
        if (!punct && !uses_pair_kernel(n, framebits)) {
            const int rc = dropin_launch(st, framebits, (const uint8_t*)syms, bounce_out,
                                         reinterpret_cast<uint32_t*>(bounce_out + out_pad), n);
            if (rc == FEC_OK) memcpy(user_out, bounce_out, out_bytes);
            return rc;
        }
        if (is_u32) {  // compacted already: the rest of the path sees the byte layout
            return vit_host(framebits, syms, SymFormat::U8, n, user_out);
        }
    }
    // chunks are pipelined over kPipe streams: the H2D copy of chunk k+1, the kernel of chunk k and the
    // D2H copy of chunk k-1 overlap
    size_t chunk = host_chunk_frames(framebits);
    if (chunk > n) chunk = n;
    int rc = FEC_OK;
    size_t done = 0;
    for (int k = 0; done < n && rc == FEC_OK; k++) {
        Slot& s = g_pipe.slot[k % kPipe];
        const size_t m = (n - done < chunk) ? n - done : chunk;
        // the slot's previous chunk must have left its buffers
        if (fail(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize")) { rc = FEC_ERR_DEVICE; break; }
        const int blocks = viterbi_grid_blocks(st->num_sms, m, framebits);
        if (!grow(&s.d_in, &s.in_cap, m * nsym) || !grow(&s.d_out, &s.out_cap, m * nout) ||
            !grow(&s.d_scratch, &s.scratch_cap, viterbi_scratch_bytes(blocks, framebits)) ||
            ((is_u32 || punct) && !grow(&s.d_aux, &s.aux_cap, m * in_row + 16))) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        const uint8_t* src = (const uint8_t*)syms + done * in_row;
        if (punct) {
            if (in_row && fail(cudaMemcpyAsync(s.d_aux, src, m * in_row, cudaMemcpyHostToDevice, s.stream), "H2D")) {
                rc = FEC_ERR_DEVICE;
                break;
            }
            if (punct_fused() && uses_pair_kernel(m, framebits)) {
                // the throughput kernel expands the rows in its symbol fetch (d_aux carries kPunctSlackBytes of slack)
                rc = vit_device_punctured(st, framebits, (const uint8_t*)s.d_aux, rx_per_frame, (const uint8_t*)g_pipe.d_idx + idx_bytes,
                                          erasure, m, (uint8_t*)s.d_out, s.stream, s.d_scratch, s.scratch_cap);
                if (rc != FEC_OK) break;
                if (fail(cudaMemcpyAsync(out + done * nout, s.d_out, m * nout, cudaMemcpyDeviceToHost, s.stream), "D2H")) {
                    rc = FEC_ERR_DEVICE;
                    break;
                }
                done += m;
                continue;
            }
            if (fail(launch_depuncture((const uint8_t*)s.d_aux, rx_per_frame, (const int32_t*)g_pipe.d_idx, framebits, erasure,
                                       (uint8_t*)s.d_in, m, st->num_sms, s.stream),
                     "depuncture kernel")) {
                rc = FEC_ERR_DEVICE;
                break;
            }
        } else if (is_u32) {
            if (fail(cudaMemcpyAsync(s.d_aux, src, m * in_row, cudaMemcpyHostToDevice, s.stream), "H2D") ||
                fail(launch_compact_symbols((const uint32_t*)s.d_aux, (uint8_t*)s.d_in, m * nsym, st->num_sms, s.stream),
                     "compact kernel")) {
                rc = FEC_ERR_DEVICE;
                break;
            }
        } else if (fail(cudaMemcpyAsync(s.d_in, src, m * nsym, cudaMemcpyHostToDevice, s.stream), "H2D")) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        rc = vit_device(st, framebits, (const uint8_t*)s.d_in, m, (uint8_t*)s.d_out, s.stream, s.d_scratch, s.scratch_cap);
        if (rc != FEC_OK) break;
        if (fail(cudaMemcpyAsync(out + done * nout, s.d_out, m * nout, cudaMemcpyDeviceToHost, s.stream), "D2H")) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        done += m;
    }
    for (Slot& s : g_pipe.slot)
        if (s.stream && fail(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize") && rc == FEC_OK) rc = FEC_ERR_DEVICE;
    if (bounce_out && rc == FEC_OK) memcpy(user_out, bounce_out, out_bytes);
    trim_pipe();
    return rc;
}

// Device pointer through which a kernel can read the caller's host buffer directly, or nullptr when the buffer is
// ordinary pageable memory (pinned memory -- fec_host_alloc(), cudaMallocHost, cudaHostRegister -- is mapped).
const uint8_t* mapped_host_pointer(const void* p) {
    static const bool never = [] {  // VITERBI_B200_RS_UPLOAD=1: always upload outVector (A/B measurements)
        const char* env = getenv("VITERBI_B200_RS_UPLOAD");
        return env && *env == '1';
    }();
    if (never) return nullptr;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return attr.type == cudaMemoryTypeHost ? (const uint8_t*)attr.devicePointer : nullptr;
}

int rs_host(const uint8_t* in, unsigned s, size_t n, uint8_t* out, int32_t* ret) {
    if (s == 0 || s > kRsMaxDims) return bad_arg("RSDims must be 1..1024");
    if (n == 0) return FEC_OK;
    if (!in || !out || !ret) return bad_arg("null pointer");
    int dev;
    DeviceState* st = device_state(&dev);
    if (!st) return FEC_ERR_DEVICE;
    if (!prepare_pipe(dev)) return FEC_ERR_DEVICE;
    const size_t in_row = 120 * (size_t)s, out_row = 110 * (size_t)s;
    const size_t in_bytes = n * in_row, out_bytes = n * out_row, ret_bytes = n * sizeof(int32_t);
    const size_t in_pad = (in_bytes + 255) & ~(size_t)255, out_pad = (out_bytes + 255) & ~(size_t)255;
    if (in_bytes + out_bytes <= kBounceBytes) {
        // Small calls (the single-superframe drop-in): the kernel runs directly on this thread's pinned bounce
        // buffer through its device mapping -- it stages the superframe from there, applies the partial-write rule
        // to the copy of the caller's outVector in place and writes the return values next to it.  One kernel
        // launch, one wait, no copy operations.  p == outVector (legal in the reference, which copies each column
        // to rsBlock first: rschecksf.cpp:75-84) is safe: the input is copied before the output is touched.
        Slot& s0 = g_pipe.slot[0];
        if (s0.pin_cap < in_pad + out_pad + ret_bytes) drop_graphs(g_pipe, true);
        if (!grow(&s0.h_pin, &s0.pin_cap, in_pad + out_pad + ret_bytes, true)) return FEC_ERR_DEVICE;
        uint8_t* bounce = (uint8_t*)s0.h_pin;
        memcpy(bounce, in, in_bytes);
        memcpy(bounce + in_pad, out, out_bytes);  // the partial-write rule keeps the caller's bytes
        int32_t* b_ret = reinterpret_cast<int32_t*>(bounce + in_pad + out_pad);
        if (fail(launch_rs_superframes(bounce, bounce + in_pad, b_ret, nullptr, n, s, st->num_sms, s0.stream), "rs kernel launch") ||
            !wait_stream(s0.stream, "rs superframe check"))
            return FEC_ERR_DEVICE;
        memcpy(out, bounce + in_pad, out_bytes);
        memcpy(ret, b_ret, ret_bytes);
        return FEC_OK;
    }
    // The partial-write rule (rschecksf.cpp:80-88) leaves the columns from the first failing one on untouched, so
    // the result rows are a merge of decoded bytes and the caller's current outVector bytes.  Two ways to get the
    // caller's bytes to the device: upload outVector ahead of the kernel (DMA, all rows), or -- when outVector is
    // pinned -- let the kernel fetch the rows of FAILING superframes itself through the buffer's device mapping.
    // Measured (profiles/rs_e2e_ab.py): with no failures the second is 1.56x faster end to end (67 vs 43 M
    // superframes/s on the s = 1..8 mix: 0.46x the bytes cross PCIe), but kernel reads reach only ~85 % of the DMA
    // rate, so above ~55 % failing superframes the upload wins.  The choice is made per chunk from the failure
    // fraction of the chunks already returned (carried over between calls of the same thread).
    const uint8_t* d_orig_base = mapped_host_pointer(out);
    size_t chunk = rs_chunk_bytes() / in_row;
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    int rc = FEC_OK;
    size_t done = 0;
    size_t slot_first[kPipe] = {0}, slot_count[kPipe] = {0};  // superframes whose return values a slot last produced
    auto observe = [&](int k) {  // the slot's stream has been synchronised: its return values are in `ret`
        if (!slot_count[k]) return;
        size_t failed = 0;
        for (size_t i = 0; i < slot_count[k]; i++) failed += ret[slot_first[k] + i] < 0;
        const float frac = (float)failed / (float)slot_count[k];
        g_pipe.rs_fail_frac = g_pipe.rs_fail_frac < 0 ? frac : 0.5f * g_pipe.rs_fail_frac + 0.5f * frac;
        slot_count[k] = 0;
    };
    for (int k = 0; done < n && rc == FEC_OK; k++) {
        Slot& sl = g_pipe.slot[k % kPipe];
        const size_t m = (n - done < chunk) ? n - done : chunk;
        if (fail(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize")) { rc = FEC_ERR_DEVICE; break; }
        observe(k % kPipe);
        if (!grow(&sl.d_in, &sl.in_cap, m * in_row) || !grow(&sl.d_out, &sl.out_cap, m * out_row) ||
            !grow(&sl.d_aux, &sl.aux_cap, m * sizeof(int32_t))) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        const bool fetch = d_orig_base && g_pipe.rs_fail_frac < kRsFetchMaxFailFrac;
        const uint8_t* d_orig = fetch ? d_orig_base + done * out_row : nullptr;
        if (fail(cudaMemcpyAsync(sl.d_in, in + done * in_row, m * in_row, cudaMemcpyHostToDevice, sl.stream), "H2D") ||
            (!d_orig && fail(cudaMemcpyAsync(sl.d_out, out + done * out_row, m * out_row, cudaMemcpyHostToDevice, sl.stream), "H2D out")) ||
            fail(launch_rs_superframes((const uint8_t*)sl.d_in, (uint8_t*)sl.d_out, (int32_t*)sl.d_aux, d_orig, m, s,
                                       st->num_sms, sl.stream),
                 "rs kernel launch") ||
            fail(cudaMemcpyAsync(out + done * out_row, sl.d_out, m * out_row, cudaMemcpyDeviceToHost, sl.stream), "D2H") ||
            fail(cudaMemcpyAsync(ret + done, sl.d_aux, m * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream), "D2H ret")) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        slot_first[k % kPipe] = done;
        slot_count[k % kPipe] = m;
        done += m;
    }
    for (int k = 0; k < kPipe; k++) {
        Slot& sl = g_pipe.slot[k];
        if (sl.stream && fail(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize") && rc == FEC_OK) rc = FEC_ERR_DEVICE;
        if (rc == FEC_OK) observe(k);
    }
    trim_pipe();
    return rc;
}

// Viterbi -> superframe -> RS with host buffers, one device (the calling thread's).
int dabplus_host(unsigned framebits, const uint8_t* syms, size_t nsf, uint8_t* out, int32_t* ret);

// ---- multi-device host calls --------------------------------------------------------------------------------
// Frames and superframes are independent (deconvolve.cpp:116-132 re-initialises the metrics per call,
// rschecksf.cpp:72 uses stack scratch), so one host batch is cut into contiguous shards, one per selected
// device, and each shard runs the single-device host path on its own worker thread (own streams, own staging
// buffers) -- no exchange between devices.  Workers are persistent so that their staging state survives calls.
struct Worker {
    int dev = -1;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = FEC_OK;
    std::string error;

    void loop() {
        t_device = dev;
        std::unique_lock<std::mutex> lock(mu);
        for (;;) {
            cv.wait(lock, [this] { return has_job || quit; });
            if (quit) return;
            lock.unlock();
            t_error.clear();
            const int r = job();
            lock.lock();
            rc = r;
            error = t_error;
            has_job = false;
            done = true;
            cv.notify_all();
        }
    }
};

struct WorkerPool {
    std::mutex mu;  // serialises multi-device calls (each already uses every selected device)
    std::vector<int> devices;  // empty = all visible devices
    std::unique_ptr<Worker> worker[kMaxDevices];

    ~WorkerPool() {
        for (auto& w : worker) {
            if (!w) continue;
            {
                std::lock_guard<std::mutex> lock(w->mu);
                w->quit = true;
            }
            w->cv.notify_all();
            if (w->th.joinable()) w->th.join();
        }
    }

    Worker* get(int dev) {
        if (!worker[dev]) {
            worker[dev].reset(new Worker());
            worker[dev]->dev = dev;
            worker[dev]->th = std::thread([w = worker[dev].get()] { w->loop(); });
        }
        return worker[dev].get();
    }
};
WorkerPool g_pool;

// devices of the multi-device calls (pool mutex held)
bool selected_devices(std::vector<int>& devs) {
    devs = g_pool.devices;
    if (devs.empty()) {
        int n = 0;
        if (fail(cudaGetDeviceCount(&n), "cudaGetDeviceCount") || n < 1) {
            if (t_error.empty()) t_error = "no CUDA device";
            return false;
        }
        for (int i = 0; i < n && i < kMaxDevices; i++) devs.push_back(i);
    }
    return true;
}

// Cut `units` work units into one contiguous range per device (boundaries on multiples of `align` units) and run
// fn(lo, hi) for each non-empty range on that device's worker.  Returns the first failure.
int run_sharded(size_t units, size_t align, const std::function<int(size_t, size_t)>& fn) {
    std::lock_guard<std::mutex> lock(g_pool.mu);
    std::vector<int> devs;
    if (!selected_devices(devs)) return FEC_ERR_DEVICE;
    const size_t world = devs.size(), blocks = (units + align - 1) / align;
    std::vector<Worker*> used;
    for (size_t r = 0; r < world; r++) {
        size_t lo = blocks * r / world * align, hi = blocks * (r + 1) / world * align;
        if (lo > units) lo = units;
        if (hi > units) hi = units;
        if (lo == hi) continue;
        Worker* w = g_pool.get(devs[r]);
        {
            std::lock_guard<std::mutex> wl(w->mu);
            w->job = [&fn, lo, hi] { return fn(lo, hi); };
            w->has_job = true;
            w->done = false;
        }
        w->cv.notify_all();
        used.push_back(w);
    }
    int rc = FEC_OK;
    for (Worker* w : used) {
        std::unique_lock<std::mutex> wl(w->mu);
        w->cv.wait(wl, [w] { return w->done; });
        if (w->rc != FEC_OK && rc == FEC_OK) {
            rc = w->rc;
            t_error = "device " + std::to_string(w->dev) + ": " + w->error;
        }
    }
    return rc;
}

// ---- NCCL, loaded at run time (the library has no link-time dependency on it): the one collective of the design,
// the gather of result arrays over NVLink, for hosts that keep the shards on the devices.
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    void* handle = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::vector<int> devs;
    std::vector<ncclComm_t> comms;
};
NcclApi g_nccl;

bool nccl_load() {
    if (g_nccl.handle) return true;
    const char* env = getenv("VITERBI_B200_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names)
        if (nm && *nm && (h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) {
        t_error = "NCCL not found (libnccl.so.2; set VITERBI_B200_NCCL_LIB)";
        return false;
    }
    auto sym = [h](const char* n) { return dlsym(h, n); };
    g_nccl.CommInitAll = (int (*)(ncclComm_t*, int, const int*))sym("ncclCommInitAll");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))sym("ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.AllGather) {
        t_error = "NCCL library lacks a required symbol";
        dlclose(h);
        return false;
    }
    g_nccl.handle = h;
    return true;
}

bool nccl_fail(int r, const char* what) {
    if (r == 0) return false;
    t_error = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return true;
}

// ---- optional call log (the run-time counterpart of the reference's VIT_WRITE_LOGFILE build,
// deconvolve.cpp:568-621 / rschecksf.cpp:103-185): one line per API call with the call index, wall-clock entry
// time, time since the previous call, thread id, call duration, the number of calls in flight when this one
// returned ("ReE", the reference's re-entrancy counter), the frame / superframe shape and the buffer addresses.
// Enabled by VITERBI_B200_LOG=<file>, read by initialize() and on first use; off by default. ------------------
struct CallLogFile {
    std::mutex mu;
    FILE* fp = nullptr;
    std::string path;
    unsigned long long counter = 0;
    std::chrono::steady_clock::time_point last{};
    bool have_last = false;
};
CallLogFile g_log;
std::atomic<int> g_log_on{-1};  // -1 not configured yet, 0 off, 1 on
std::atomic<int> g_calls_in_flight{0};

void configure_call_log() {
    const char* env = getenv("VITERBI_B200_LOG");
    std::lock_guard<std::mutex> lock(g_log.mu);
    const std::string want = (env && *env) ? env : "";
    if (want != g_log.path || g_log_on.load() < 0) {
        if (g_log.fp) fclose(g_log.fp);
        g_log.fp = want.empty() ? nullptr : fopen(want.c_str(), "a");
        g_log.path = want;
        g_log.have_last = false;
    }
    g_log_on.store(g_log.fp ? 1 : 0);
}

class CallLog {
  public:
    CallLog(const char* fn, unsigned shape, size_t n, const void* in, const void* out)
        : fn_(fn), shape_(shape), n_(n), in_(in), out_(out) {
        if (g_log_on.load(std::memory_order_relaxed) < 0) configure_call_log();
        on_ = g_log_on.load(std::memory_order_relaxed) == 1;
        if (!on_) return;
        g_calls_in_flight.fetch_add(1);
        wall_ = std::chrono::system_clock::now();
        t0_ = std::chrono::steady_clock::now();
    }
    void result(int rc) { rc_ = rc; }
    ~CallLog() {
        if (!on_) return;
        const auto t1 = std::chrono::steady_clock::now();
        const int others = g_calls_in_flight.fetch_sub(1) - 1;
        const double us = std::chrono::duration<double, std::micro>(t1 - t0_).count();
        const auto since_epoch = std::chrono::duration_cast<std::chrono::microseconds>(wall_.time_since_epoch()).count();
        const time_t secs = (time_t)(since_epoch / 1000000);
        struct tm tmv;
        localtime_r(&secs, &tmv);
        const size_t tid = std::hash<std::thread::id>()(std::this_thread::get_id()) % 100000;
        std::lock_guard<std::mutex> lock(g_log.mu);
        if (!g_log.fp) return;
        const double dt_ms = g_log.have_last ? std::chrono::duration<double, std::milli>(t0_ - g_log.last).count() : 0.0;
        g_log.last = t0_;
        g_log.have_last = true;
        fprintf(g_log.fp, "%6llu  %02d:%02d:%02d.%06lld  dT: %8.3f ms  TID: %5zu  %s: %9.1f us  ReE: %d  shape: %4u  n: %zu  "
                          "In: %p  Out: %p  rc: %d\n",
                g_log.counter++, tmv.tm_hour, tmv.tm_min, tmv.tm_sec, (long long)(since_epoch % 1000000), dt_ms, tid, fn_, us,
                others, shape_, n_, in_, out_, rc_);
        fflush(g_log.fp);
    }

  private:
    const char* fn_;
    unsigned shape_;
    size_t n_;
    const void *in_, *out_;
    bool on_ = false;
    int rc_ = 0;
    std::chrono::system_clock::time_point wall_;
    std::chrono::steady_clock::time_point t0_;
};

}  // namespace

namespace {

int dabplus_host(unsigned framebits, const uint8_t* syms, size_t nsf, uint8_t* out, int32_t* ret) {
    if (!vit_args_ok(framebits) || framebits == 0 || framebits % 192u) return bad_arg("framebits must be a multiple of 192");
    if (nsf == 0) return FEC_OK;
    if (!syms || !out || !ret) return bad_arg("null pointer");
    int dev;
    DeviceState* st = device_state(&dev);
    if (!st) return FEC_ERR_DEVICE;
    if (!prepare_pipe(dev)) return FEC_ERR_DEVICE;
    const unsigned rsdims = framebits / 192u;
    const size_t nsym = 4 * ((size_t)framebits + 6), in_row = 5 * nsym, out_row = 110 * (size_t)rsdims;
    const uint8_t* d_orig_base = mapped_host_pointer(out);  // see rs_host: no upload of a pinned `out`
    size_t chunk = host_chunk_frames(framebits) / 5;
    if (chunk < 1) chunk = 1;
    if (chunk > nsf) chunk = nsf;
    int rc = FEC_OK;
    size_t done = 0;
    for (int k = 0; done < nsf && rc == FEC_OK; k++) {
        Slot& sl = g_pipe.slot[k % kPipe];
        const size_t m = (nsf - done < chunk) ? nsf - done : chunk;
        if (fail(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize")) { rc = FEC_ERR_DEVICE; break; }
        const int blocks = viterbi_grid_blocks(st->num_sms, m * 5, framebits);
        // d_scratch: [decoded frames = superframes, 120*s bytes each][Viterbi decision scratch]
        const size_t bits_bytes = (m * 120 * (size_t)rsdims + 255) & ~(size_t)255;
        const size_t scratch_bytes = viterbi_scratch_bytes(blocks, framebits);
        if (!grow(&sl.d_in, &sl.in_cap, m * in_row) || !grow(&sl.d_out, &sl.out_cap, m * out_row) ||
            !grow(&sl.d_aux, &sl.aux_cap, m * sizeof(int32_t)) || !grow(&sl.d_scratch, &sl.scratch_cap, bits_bytes + scratch_bytes)) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        uint8_t* d_bits = (uint8_t*)sl.d_scratch;
        const uint8_t* d_orig = d_orig_base ? d_orig_base + done * out_row : nullptr;
        if (fail(cudaMemcpyAsync(sl.d_in, syms + done * in_row, m * in_row, cudaMemcpyHostToDevice, sl.stream), "H2D") ||
            (!d_orig && fail(cudaMemcpyAsync(sl.d_out, out + done * out_row, m * out_row, cudaMemcpyHostToDevice, sl.stream), "H2D out"))) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        rc = vit_device(st, framebits, (const uint8_t*)sl.d_in, m * 5, d_bits, sl.stream, d_bits + bits_bytes,
                        sl.scratch_cap - bits_bytes);
        if (rc != FEC_OK) break;
        if (g_descramble.load() && fail(launch_descramble(d_bits, m * 5, framebits, st->num_sms, sl.stream), "descramble kernel")) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        if (fail(launch_rs_superframes(d_bits, (uint8_t*)sl.d_out, (int32_t*)sl.d_aux, d_orig, m, rsdims, st->num_sms, sl.stream),
                 "rs kernel launch") ||
            fail(cudaMemcpyAsync(out + done * out_row, sl.d_out, m * out_row, cudaMemcpyDeviceToHost, sl.stream), "D2H") ||
            fail(cudaMemcpyAsync(ret + done, sl.d_aux, m * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream), "D2H ret")) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        done += m;
    }
    for (Slot& sl : g_pipe.slot)
        if (sl.stream && fail(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize") && rc == FEC_OK) rc = FEC_ERR_DEVICE;
    trim_pipe();
    return rc;
}

}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace fec

using namespace fec;

// the library is built with -fvisibility=hidden; only the C ABI below is exported
#pragma GCC visibility push(default)
extern "C" {

// ---------------------------------------------------------------------------------------------
// drop-in surface
// ---------------------------------------------------------------------------------------------
int deconvolve(unsigned int framebits, unsigned int* piData, int inputLength, unsigned char* output) {
    (void)inputLength;
    CallLog log("deco", framebits, 1, piData, output);
    if (g_save_mode.load()) {  // decon_savemode, viterbi_helpers.asm:183-186
        log.result(1);
        return 1;
    }
    const int rc = vit_host(framebits, piData, SymFormat::U32, 1, output);
    log.result(rc == FEC_OK ? 0 : 1);
    if (rc == FEC_OK) return 0;
    // The reference latches save mode after a fault inside the decoder (NULL buffers give an
    // access violation there, viterbi-benchmark.cpp:457-464); a device failure is our equivalent.
    // An unsupported framebits value never faults in the reference, so it does not latch.
    if (rc == FEC_ERR_DEVICE || !piData || !output) g_save_mode.store(1);  // exc_handler.cpp:214
    return 1;
}

int RScheckSuperframe(unsigned char* p, int startIx, unsigned int RSDims, unsigned char* outVector) {
    (void)startIx;
    if (RSDims == 0) return 0;  // the reference's column loop does not execute
    CallLog log("rssf", RSDims, 1, p, outVector);
    int32_t ret = -1;
    const int rc = rs_host(p, RSDims, 1, outVector, &ret);
    if (rc != FEC_OK) ret = -1;  // exc_handler.cpp:208-211
    log.result(ret);
    return ret;
}

int RSCheckSuperframe(unsigned char* p, int startIx, unsigned int RSDims, unsigned char* outVector) {
    return RScheckSuperframe(p, startIx, RSDims, outVector);
}

int initialize(void) {
    g_save_mode.store(0);  // dllmain.cpp:157
    t_error.clear();
    configure_call_log();  // the reference re-reads its configuration here (dllmain.cpp:158 -> SetupDLL)
    const char* env = getenv("VITERBI_B200_DEVICE");
    if (env && *env) g_device.store(atoi(env));
    DeviceState* st = device_state();
    // The reference's initialize() is how a host recovers after a fault (exc_handler.cpp:214 -> dllmain.cpp:156):
    // probe the context, and after a sticky CUDA error tear it down, declare all staging state of the old context
    // dead (threads drop theirs on their next call) and set the device up again.
    if (st && cudaDeviceSynchronize() == cudaSuccess) return 1;
    (void)cudaGetLastError();
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur < 0 || cur >= kMaxDevices) {
        (void)cudaGetLastError();
        return 0;
    }
    if (cudaDeviceReset() != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    g_generation.fetch_add(1);
    {
        std::lock_guard<std::mutex> lock(g_dev[cur].mu);
        g_dev[cur].ready = false;
    }
    t_error.clear();
    return device_state() != nullptr;
}

int GetCPUCaps(void) { return 0; }

void WakeUpYMM(void) {}

// ---------------------------------------------------------------------------------------------
// batched API
// ---------------------------------------------------------------------------------------------
int viterbi_deconvolve_batch(unsigned int framebits, const uint8_t* syms, size_t n, uint8_t* out) {
    CallLog log("deco_batch", framebits, n, syms, out);
    const int rc = vit_host(framebits, syms, SymFormat::U8, n, out);
    log.result(rc);
    return rc;
}

int viterbi_deconvolve_batch_u32(unsigned int framebits, const uint32_t* syms, size_t n, uint8_t* out) {
    CallLog log("deco_batch_u32", framebits, n, syms, out);
    const int rc = vit_host(framebits, syms, SymFormat::U32, n, out);
    log.result(rc);
    return rc;
}

int viterbi_deconvolve_batch_device(unsigned int framebits, const uint8_t* d_syms, size_t n, uint8_t* d_out,
                                    void* stream) {
    if (!vit_args_ok(framebits)) return bad_arg("framebits must be even and <= 9216");
    if (n == 0 || framebits == 0) return FEC_OK;
    if (!d_syms || !d_out) return bad_arg("null pointer");
    if (reinterpret_cast<uintptr_t>(d_syms) & 7) return bad_arg("d_syms must be 8-byte aligned");
    DeviceState* st = device_state();
    if (!st) return FEC_ERR_DEVICE;
    return vit_device(st, framebits, d_syms, n, d_out, (cudaStream_t)stream, nullptr, 0);
}

int viterbi_deconvolve_batch_u32_device(unsigned int framebits, const uint32_t* d_syms, size_t n, uint8_t* d_out,
                                        void* stream) {
    if (!vit_args_ok(framebits)) return bad_arg("framebits must be even and <= 9216");
    if (n == 0 || framebits == 0) return FEC_OK;
    if (!d_syms || !d_out) return bad_arg("null pointer");
    if (reinterpret_cast<uintptr_t>(d_syms) & 15) return bad_arg("d_syms must be 16-byte aligned (one trellis step per uint4)");
    DeviceState* st = device_state();
    if (!st) return FEC_ERR_DEVICE;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nsym = 4 * ((size_t)framebits + 6) * n;
    void* d_u8 = nullptr;
    if (fail(cudaMallocAsync(&d_u8, nsym, s), "cudaMallocAsync(u8 symbols)")) return FEC_ERR_DEVICE;
    int rc = fail(launch_compact_symbols(d_syms, (uint8_t*)d_u8, nsym, st->num_sms, s), "compact kernel") ? FEC_ERR_DEVICE
                                                                                                          : FEC_OK;
    if (rc == FEC_OK) rc = vit_device(st, framebits, (const uint8_t*)d_u8, n, d_out, s, nullptr, 0);
    (void)cudaFreeAsync(d_u8, s);
    return rc;
}

int viterbi_deconvolve_batch_punctured(unsigned int framebits, const uint8_t* rx, size_t rx_per_frame, const uint8_t* keep,
                                       unsigned int erasure, size_t n, uint8_t* out) {
    CallLog log("deco_batch_punct", framebits, n, rx, out);
    const int rc = vit_host(framebits, rx, SymFormat::Punctured, n, out, keep, rx_per_frame, erasure);
    log.result(rc);
    return rc;
}

int viterbi_deconvolve_batch_punctured_device(unsigned int framebits, const uint8_t* d_rx, size_t rx_per_frame,
                                              const uint8_t* keep, unsigned int erasure, size_t n, uint8_t* d_out,
                                              void* stream) {
    if (!vit_args_ok(framebits)) return bad_arg("framebits must be even and <= 9216");
    if (n == 0 || framebits == 0) return FEC_OK;
    if (!d_rx || !d_out || !keep) return bad_arg("null pointer");
    if (erasure > 255) return bad_arg("erasure must be 0..255");
    std::vector<int32_t> idx;
    if (!puncture_index(framebits, keep, rx_per_frame, idx)) return bad_arg("keep pattern does not match rx_per_frame");
    DeviceState* st = device_state();
    if (!st) return FEC_ERR_DEVICE;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nsym = 4 * ((size_t)framebits + 6);
    if (punct_fused() && uses_pair_kernel(n, framebits)) {
        // Fused path: the throughput kernel expands the rows in its symbol fetch.  Its fetch may touch up to
        // kPunctSlackBytes past a row, which for every row but the last one is simply the next row; the last row
        // is read from a padded copy, so nothing is assumed about what follows the caller's buffer.
        const size_t ptab_bytes = ((size_t)framebits + 6) / 2 * 16, last_bytes = rx_per_frame + kPunctSlackBytes;
        std::vector<uint32_t> ptab(ptab_bytes / 4);
        punct_table(framebits, keep, ptab.data());
        void* d_tmp = nullptr;  // [table][padded last row]
        if (fail(cudaMallocAsync(&d_tmp, ptab_bytes + last_bytes, s), "cudaMallocAsync(puncturing table)")) return FEC_ERR_DEVICE;
        uint8_t* d_last = (uint8_t*)d_tmp + ptab_bytes;
        // ptab is pageable: cudaMemcpyAsync returns once it has been staged, so the vector may die with this call
        int rc = (fail(cudaMemcpyAsync(d_tmp, ptab.data(), ptab_bytes, cudaMemcpyHostToDevice, s), "H2D puncturing table") ||
                  fail(cudaMemsetAsync(d_last + rx_per_frame, 0, kPunctSlackBytes, s), "cudaMemsetAsync") ||
                  (rx_per_frame && fail(cudaMemcpyAsync(d_last, d_rx + (n - 1) * rx_per_frame, rx_per_frame, cudaMemcpyDeviceToDevice, s),
                                        "D2D last row")))
                     ? FEC_ERR_DEVICE
                     : FEC_OK;
        if (rc == FEC_OK) rc = vit_device_punctured(st, framebits, d_rx, rx_per_frame, d_tmp, erasure, n, d_out, s, nullptr, 0, d_last);
        (void)cudaFreeAsync(d_tmp, s);
        return rc;
    }
    void *d_idx = nullptr, *d_syms = nullptr;
    if (fail(cudaMallocAsync(&d_idx, nsym * sizeof(int32_t), s), "cudaMallocAsync(index table)")) return FEC_ERR_DEVICE;
    if (fail(cudaMallocAsync(&d_syms, nsym * n, s), "cudaMallocAsync(expanded symbols)")) {
        (void)cudaFreeAsync(d_idx, s);
        return FEC_ERR_DEVICE;
    }
    // idx is pageable: cudaMemcpyAsync returns once it has been staged, so the vector may die with this call
    int rc = (fail(cudaMemcpyAsync(d_idx, idx.data(), nsym * sizeof(int32_t), cudaMemcpyHostToDevice, s), "H2D index table") ||
              fail(launch_depuncture(d_rx, rx_per_frame, (const int32_t*)d_idx, framebits, erasure, (uint8_t*)d_syms, n,
                                     st->num_sms, s),
                   "depuncture kernel"))
                 ? FEC_ERR_DEVICE
                 : FEC_OK;
    if (rc == FEC_OK) rc = vit_device(st, framebits, (const uint8_t*)d_syms, n, d_out, s, nullptr, 0);
    (void)cudaFreeAsync(d_syms, s);
    (void)cudaFreeAsync(d_idx, s);
    return rc;
}

int rs_check_superframe_batch(const uint8_t* in, unsigned int RSDims, size_t n, uint8_t* out, int32_t* ret) {
    CallLog log("rssf_batch", RSDims, n, in, out);
    const int rc = rs_host(in, RSDims, n, out, ret);
    log.result(rc);
    return rc;
}

namespace {
bool copies_ok(const uint8_t* d_out, uint8_t* const* outs, int32_t* const* rets, int ncopies) {
    if (ncopies < 0 || ncopies > kRsMaxCopies || (ncopies > 0 && (!outs || !rets))) return false;
    for (int c = 0; c < ncopies; c++)
        if (!outs[c] || !rets[c] || ((reinterpret_cast<uintptr_t>(outs[c]) ^ reinterpret_cast<uintptr_t>(d_out)) & 3)) return false;
    return true;
}
}  // namespace

int rs_check_superframe_batch_device_bcast(const uint8_t* d_in, unsigned int RSDims, size_t n, uint8_t* d_out, int32_t* d_ret,
                                           uint8_t* const* d_out_copies, int32_t* const* d_ret_copies, int ncopies, void* stream) {
    if (RSDims == 0 || RSDims > kRsMaxDims) return bad_arg("RSDims must be 1..1024");
    if (!copies_ok(d_out, d_out_copies, d_ret_copies, ncopies)) return bad_arg("bad copy list (at most 15, non-null, same alignment modulo 4 as d_out)");
    if (n == 0) return FEC_OK;
    if (!d_in || !d_out || !d_ret) return bad_arg("null pointer");
    DeviceState* st = device_state();
    if (!st) return FEC_ERR_DEVICE;
    return fail(launch_rs_superframes(d_in, d_out, d_ret, nullptr, n, RSDims, st->num_sms, (cudaStream_t)stream, d_out_copies,
                                      d_ret_copies, ncopies),
                "rs kernel launch")
               ? FEC_ERR_DEVICE
               : FEC_OK;
}

int rs_check_superframe_batch_device(const uint8_t* d_in, unsigned int RSDims, size_t n, uint8_t* d_out,
                                     int32_t* d_ret, void* stream) {
    return rs_check_superframe_batch_device_bcast(d_in, RSDims, n, d_out, d_ret, nullptr, nullptr, 0, stream);
}

// Viterbi -> superframe -> RS on the device.  Five consecutive decoded frames ARE one superframe
// ([nsf*5][F/8] == [nsf][120*s] with s = F/192), so no regrouping pass is needed between the kernels.
int dabplus_decode_superframes_device_bcast(unsigned int framebits, const uint8_t* d_syms, size_t nsf, uint8_t* d_out, int32_t* d_ret,
                                            uint8_t* const* d_out_copies, int32_t* const* d_ret_copies, int ncopies, void* stream) {
    if (!vit_args_ok(framebits) || framebits == 0 || framebits % 192u) return bad_arg("framebits must be a multiple of 192");
    if (!copies_ok(d_out, d_out_copies, d_ret_copies, ncopies)) return bad_arg("bad copy list (at most 15, non-null, same alignment modulo 4 as d_out)");
    if (nsf == 0) return FEC_OK;
    if (!d_syms || !d_out || !d_ret) return bad_arg("null pointer");
    if (reinterpret_cast<uintptr_t>(d_syms) & 7) return bad_arg("d_syms must be 8-byte aligned");
    DeviceState* st = device_state();
    if (!st) return FEC_ERR_DEVICE;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned rsdims = framebits / 192u;
    void* d_bits = nullptr;
    if (fail(cudaMallocAsync(&d_bits, nsf * 120 * (size_t)rsdims, s), "cudaMallocAsync(decoded frames)")) return FEC_ERR_DEVICE;
    int rc = vit_device(st, framebits, d_syms, nsf * 5, (uint8_t*)d_bits, s, nullptr, 0);
    if (rc == FEC_OK && g_descramble.load() && fail(launch_descramble((uint8_t*)d_bits, nsf * 5, framebits, st->num_sms, s), "descramble kernel"))
        rc = FEC_ERR_DEVICE;
    if (rc == FEC_OK && fail(launch_rs_superframes((const uint8_t*)d_bits, d_out, d_ret, nullptr, nsf, rsdims, st->num_sms, s,
                                                   d_out_copies, d_ret_copies, ncopies),
                             "rs kernel launch"))
        rc = FEC_ERR_DEVICE;
    (void)cudaFreeAsync(d_bits, s);
    return rc;
}

int dabplus_decode_superframes_device(unsigned int framebits, const uint8_t* d_syms, size_t nsf, uint8_t* d_out,
                                      int32_t* d_ret, void* stream) {
    return dabplus_decode_superframes_device_bcast(framebits, d_syms, nsf, d_out, d_ret, nullptr, nullptr, 0, stream);
}

int dabplus_decode_superframes(unsigned int framebits, const uint8_t* syms, size_t nsf, uint8_t* out, int32_t* ret) {
    CallLog log("dabplus", framebits, nsf, syms, out);
    const int rc = dabplus_host(framebits, syms, nsf, out, ret);
    log.result(rc);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// multi-device host calls: one batch, all selected devices, one process
// ---------------------------------------------------------------------------------------------
int viterbi_deconvolve_batch_multi(unsigned int framebits, const uint8_t* syms, size_t n, uint8_t* out) {
    CallLog log("deco_multi", framebits, n, syms, out);
    int rc;
    if (!vit_args_ok(framebits))
        rc = bad_arg("framebits must be even and <= 9216");
    else if (n == 0 || framebits == 0)
        rc = FEC_OK;
    else if (!syms || !out)
        rc = bad_arg("null pointer");
    else {
        const size_t row = 4 * ((size_t)framebits + 6), nout = (framebits + 7) / 8;
        rc = run_sharded(n, 64, [=](size_t lo, size_t hi) {  // 64 frames = one warp group of the kernel
            return vit_host(framebits, syms + lo * row, SymFormat::U8, hi - lo, out + lo * nout);
        });
    }
    log.result(rc);
    return rc;
}

int rs_check_superframe_batch_multi(const uint8_t* in, unsigned int RSDims, size_t n, uint8_t* out, int32_t* ret) {
    CallLog log("rssf_multi", RSDims, n, in, out);
    int rc;
    if (RSDims == 0 || RSDims > kRsMaxDims)
        rc = bad_arg("RSDims must be 1..1024");
    else if (n == 0)
        rc = FEC_OK;
    else if (!in || !out || !ret)
        rc = bad_arg("null pointer");
    else
        rc = run_sharded(n, 1, [=](size_t lo, size_t hi) {
            return rs_host(in + lo * 120 * (size_t)RSDims, RSDims, hi - lo, out + lo * 110 * (size_t)RSDims, ret + lo);
        });
    log.result(rc);
    return rc;
}

int dabplus_decode_superframes_multi(unsigned int framebits, const uint8_t* syms, size_t nsf, uint8_t* out, int32_t* ret) {
    CallLog log("dabplus_multi", framebits, nsf, syms, out);
    int rc;
    if (!vit_args_ok(framebits) || framebits == 0 || framebits % 192u)
        rc = bad_arg("framebits must be a multiple of 192");
    else if (nsf == 0)
        rc = FEC_OK;
    else if (!syms || !out || !ret)
        rc = bad_arg("null pointer");
    else {
        const size_t in_row = 5 * 4 * ((size_t)framebits + 6), out_row = 110 * (size_t)(framebits / 192u);
        rc = run_sharded(nsf, 1, [=](size_t lo, size_t hi) {  // whole superframes: the 5 frames of one stay together
            return dabplus_host(framebits, syms + lo * in_row, hi - lo, out + lo * out_row, ret + lo);
        });
    }
    log.result(rc);
    return rc;
}

int fec_set_devices(const int* ordinals, int count) {
    if (count < 0 || count > kMaxDevices || (count > 0 && !ordinals)) return bad_arg("bad device list");
    int ndev = 0;
    if (fail(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount")) return FEC_ERR_DEVICE;
    std::vector<int> devs;
    for (int i = 0; i < count; i++) {
        if (ordinals[i] < 0 || ordinals[i] >= ndev || ordinals[i] >= kMaxDevices) return bad_arg("device ordinal out of range");
        for (int d : devs)
            if (d == ordinals[i]) return bad_arg("duplicate device ordinal");
        devs.push_back(ordinals[i]);
    }
    std::lock_guard<std::mutex> lock(g_pool.mu);
    if (devs != g_pool.devices && !g_nccl.comms.empty()) {  // communicators belong to the old device list
        for (ncclComm_t c : g_nccl.comms) g_nccl.CommDestroy(c);
        g_nccl.comms.clear();
        g_nccl.devs.clear();
    }
    g_pool.devices = devs;
    return FEC_OK;
}

// cudaDeviceEnablePeerAccess between every pair of `devs` (pool mutex held)
static int enable_peer_access_among(const std::vector<int>& devs) {
    int cur = -1;
    (void)cudaGetDevice(&cur);
    int rc = FEC_OK;
    for (int a : devs)
        for (int b : devs) {
            if (a == b) continue;
            int can = 0;
            if (fail(cudaDeviceCanAccessPeer(&can, a, b), "cudaDeviceCanAccessPeer") || !can) {
                if (t_error.empty()) t_error = "no peer access between devices " + std::to_string(a) + " and " + std::to_string(b);
                rc = FEC_ERR_DEVICE;
                continue;
            }
            if (cudaSetDevice(a) != cudaSuccess) { rc = FEC_ERR_DEVICE; continue; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                fail(e, "cudaDeviceEnablePeerAccess");
                rc = FEC_ERR_DEVICE;
            }
            (void)cudaGetLastError();
        }
    if (cur >= 0) (void)cudaSetDevice(cur);
    return rc;
}

int fec_enable_peer_access(void) {
    std::lock_guard<std::mutex> lock(g_pool.mu);
    std::vector<int> devs;
    if (!selected_devices(devs)) return FEC_ERR_DEVICE;
    return enable_peer_access_among(devs);
}

// ---- CUDA IPC: one process per GPU hosts map each other's result buffers, so that the *_bcast kernels can store
// into them over NVLink.  The import happens with the CALLING thread's device current and lazy peer access, which is
// what makes the mapping usable from kernels of that device (a mapping opened under the exporting device's ordinal
// is not, even with peer access enabled afterwards).
int fec_ipc_export(const void* d_ptr, unsigned char* handle) {
    if (!d_ptr || !handle) return bad_arg("null pointer");
    if (!device_state()) return FEC_ERR_DEVICE;
    static_assert(sizeof(cudaIpcMemHandle_t) == FEC_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    if (fail(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)), "cudaIpcGetMemHandle")) return FEC_ERR_DEVICE;
    memcpy(handle, &h, sizeof h);
    return FEC_OK;
}

void* fec_ipc_import(const unsigned char* handle) {
    if (!handle) {
        bad_arg("null pointer");
        return nullptr;
    }
    if (!device_state()) return nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    if (fail(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
    return p;
}

int fec_ipc_close(void* d_ptr) {
    if (!d_ptr) return FEC_OK;
    return fail(cudaIpcCloseMemHandle(d_ptr), "cudaIpcCloseMemHandle") ? FEC_ERR_DEVICE : FEC_OK;
}

int fec_get_devices(int* ordinals, int capacity) {
    std::lock_guard<std::mutex> lock(g_pool.mu);
    std::vector<int> devs;
    if (!selected_devices(devs)) return 0;
    for (int i = 0; i < (int)devs.size() && i < capacity; i++)
        if (ordinals) ordinals[i] = devs[i];
    return (int)devs.size();
}

int fec_set_thread_device(int ordinal) {
    if (ordinal < -1 || ordinal >= kMaxDevices) return bad_arg("device ordinal out of range");
    t_device = ordinal;
    if (ordinal < 0) return FEC_OK;
    return device_state() ? FEC_OK : FEC_ERR_DEVICE;
}

// ncclAllGather over the selected devices from this one process (communicators from ncclCommInitAll, created on
// first use): shard i (bytes_per_shard bytes at d_shard[i], on device i of the list) ends up at offset
// i * bytes_per_shard of every d_all[j].  Enqueued on streams[i] (NULL array or entries = default stream).
int fec_allgather_device(const void* const* d_shard, void* const* d_all, size_t bytes_per_shard, void* const* streams) {
    if (!d_shard || !d_all) return bad_arg("null pointer");
    std::lock_guard<std::mutex> lock(g_pool.mu);
    std::vector<int> devs;
    if (!selected_devices(devs)) return FEC_ERR_DEVICE;
    for (size_t i = 0; i < devs.size(); i++)
        if (!d_shard[i] || !d_all[i]) return bad_arg("null pointer");
    if (bytes_per_shard == 0) return FEC_OK;
    int cur = -1;
    (void)cudaGetDevice(&cur);
    // Default: the copy engines push every shard into every device's array (cudaMemcpyPeerAsync over NVLink, no
    // kernel, no dependency on NCCL).  VITERBI_B200_GATHER=nccl selects the collective instead.
    static const bool use_nccl = [] {
        const char* env = getenv("VITERBI_B200_GATHER");
        return env && strcmp(env, "nccl") == 0;
    }();
    if (!use_nccl) {
        static std::vector<int> peers_enabled_for;
        if (peers_enabled_for != devs) {
            (void)enable_peer_access_among(devs);  // best effort: without it the copies are staged through the host
            t_error.clear();
            peers_enabled_for = devs;
        }
        int rc = FEC_OK;
        for (size_t i = 0; i < devs.size() && rc == FEC_OK; i++) {
            if (fail(cudaSetDevice(devs[i]), "cudaSetDevice")) { rc = FEC_ERR_DEVICE; break; }
            cudaStream_t st = streams ? (cudaStream_t)streams[i] : nullptr;
            for (size_t j = 0; j < devs.size() && rc == FEC_OK; j++) {
                uint8_t* dst = static_cast<uint8_t*>(d_all[j]) + i * bytes_per_shard;
                if (dst == d_shard[i]) continue;  // in-place shard
                if (fail(cudaMemcpyPeerAsync(dst, devs[j], d_shard[i], devs[i], bytes_per_shard, st), "cudaMemcpyPeerAsync"))
                    rc = FEC_ERR_DEVICE;
            }
        }
        if (cur >= 0) (void)cudaSetDevice(cur);
        return rc;
    }
    if (!nccl_load()) return FEC_ERR_DEVICE;
    if (g_nccl.devs != devs) {
        for (ncclComm_t c : g_nccl.comms) g_nccl.CommDestroy(c);
        g_nccl.comms.assign(devs.size(), nullptr);
        if (nccl_fail(g_nccl.CommInitAll(g_nccl.comms.data(), (int)devs.size(), devs.data()), "ncclCommInitAll")) {
            g_nccl.comms.clear();
            g_nccl.devs.clear();
            return FEC_ERR_DEVICE;
        }
        g_nccl.devs = devs;
    }
    int rc = FEC_OK;
    if (nccl_fail(g_nccl.GroupStart(), "ncclGroupStart")) return FEC_ERR_DEVICE;
    for (size_t i = 0; i < devs.size() && rc == FEC_OK; i++) {
        if (fail(cudaSetDevice(devs[i]), "cudaSetDevice") ||
            nccl_fail(g_nccl.AllGather(d_shard[i], d_all[i], bytes_per_shard, /*ncclUint8*/ 1, g_nccl.comms[i],
                                       streams ? (cudaStream_t)streams[i] : nullptr),
                      "ncclAllGather"))
            rc = FEC_ERR_DEVICE;
    }
    if (nccl_fail(g_nccl.GroupEnd(), "ncclGroupEnd")) rc = FEC_ERR_DEVICE;
    count_launch();
    if (cur >= 0) (void)cudaSetDevice(cur);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// device selection and utilities
// ---------------------------------------------------------------------------------------------
int fec_device_count(void) {
    int n = 0;
    if (fail(cudaGetDeviceCount(&n), "cudaGetDeviceCount")) return 0;
    return n;
}

int fec_set_device(int ordinal) {
    if (ordinal < 0 || ordinal >= kMaxDevices) return bad_arg("device ordinal out of range");
    if (fail(cudaSetDevice(ordinal), "cudaSetDevice")) return FEC_ERR_DEVICE;
    g_device.store(ordinal);
    return device_state() ? FEC_OK : FEC_ERR_DEVICE;
}

int fec_get_device(void) {
    int dev = g_device.load();
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return -1;
    return dev;
}

int fec_in_save_mode(void) { return g_save_mode.load(); }

const char* fec_last_error(void) { return t_error.c_str(); }

void* fec_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (fail(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost")) return nullptr;
    return p;
}

void fec_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

void* fec_device_alloc(size_t bytes) {
    if (!device_state()) return nullptr;
    void* p = nullptr;
    if (fail(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc")) return nullptr;
    return p;
}

void fec_device_free(void* p) {
    if (p) cudaFree(p);
}

int fec_memcpy_h2d(void* d_dst, const void* src, size_t bytes) {
    return fail(cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice), "cudaMemcpy H2D") ? FEC_ERR_DEVICE : FEC_OK;
}

int fec_memcpy_d2h(void* dst, const void* d_src, size_t bytes) {
    return fail(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy D2H") ? FEC_ERR_DEVICE : FEC_OK;
}

// Device-to-device copy by the copy engines (peer buffers included: the gather of result arrays without a kernel).
int fec_memcpy_d2d_async(void* d_dst, const void* d_src, size_t bytes, void* stream) {
    if (bytes == 0) return FEC_OK;
    if (!d_dst || !d_src) return bad_arg("null pointer");
    return fail(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream), "cudaMemcpyAsync D2D") ? FEC_ERR_DEVICE
                                                                                                                     : FEC_OK;
}

int fec_device_synchronize(void) { return fail(cudaDeviceSynchronize(), "cudaDeviceSynchronize") ? FEC_ERR_DEVICE : FEC_OK; }

int fec_set_energy_dispersal(int on) {
    g_descramble.store(on ? 1 : 0);
    return FEC_OK;
}

int fec_set_viterbi_kernel(int mode) {
    if (mode < FEC_VITERBI_AUTO || mode > FEC_VITERBI_WARP) return bad_arg("unknown kernel mode");
    g_vit_kernel.store(mode);
    return FEC_OK;
}

unsigned long long fec_kernel_launches(void) { return g_launches.load(); }

}  // extern "C"
#pragma GCC visibility pop
