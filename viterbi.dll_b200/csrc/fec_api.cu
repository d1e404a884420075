// fec_api.cu -- the C ABI of libviterbi_b200.so (include/viterbi_b200.h) over the CUDA runtime.
//
// Mirrors the reference's export surface (viterbi.def:4-8) and its failure convention
// (exc_handler.cpp:204-214: after a fault deconvolve returns 1 and RScheckSuperframe -1 until
// initialize() is called), and adds the batched entry points.  The CPU dispatcher / ini file
// (setupdll.cpp) is replaced by device selection.  No CPU decode path exists here.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
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

std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_save_mode{0};
std::atomic<int> g_device{-1};  // -1: use the calling thread's current device
std::atomic<int> g_vit_kernel{FEC_VITERBI_AUTO};
thread_local std::string t_error;

struct DeviceState {
    std::once_flag once;
    cudaError_t init_status = cudaSuccess;
    int num_sms = 0;
};
DeviceState g_dev[kMaxDevices];

// One staging slot of the host-pointer pipeline: device input / output / scratch on its own stream.
struct Slot {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* d_aux = nullptr;  // u32 staging (viterbi) or ret (rs)
    void* d_scratch = nullptr;
    void* h_pin = nullptr;  // pinned bounce buffer for the single-call drop-in path
    size_t in_cap = 0, out_cap = 0, aux_cap = 0, scratch_cap = 0, pin_cap = 0;
};

// Staging state of the host-pointer calls.  One per calling thread: QIRX >= 4.0 calls deconvolve() from several
// threads at once (README.md:56), and with per-thread streams and buffers those calls overlap on the device
// instead of queueing behind one lock.  Released when the thread exits.
struct HostPipe {
    int device = -1;
    Slot slot[kPipe];
    void* d_idx = nullptr;  // depuncturing index table of the call in progress
    size_t idx_cap = 0;
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

// Resolve the device this call runs on and make sure its per-device state exists.
DeviceState* device_state(int* ordinal_out = nullptr) {
    int dev = g_device.load();
    if (dev >= 0) {
        if (fail(cudaSetDevice(dev), "cudaSetDevice")) return nullptr;
    } else if (fail(cudaGetDevice(&dev), "cudaGetDevice")) {
        return nullptr;
    }
    if (dev < 0 || dev >= kMaxDevices) {
        t_error = "device ordinal out of range";
        return nullptr;
    }
    DeviceState* st = &g_dev[dev];
    std::call_once(st->once, [st, dev] {
        cudaDeviceProp prop;
        st->init_status = cudaGetDeviceProperties(&prop, dev);
        if (st->init_status != cudaSuccess) return;
        st->num_sms = prop.multiProcessorCount;
        st->init_status = rs_upload_tables();
        if (st->init_status != cudaSuccess) return;
        // keep stream-ordered scratch allocations cached in the pool between calls
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    });
    if (fail(st->init_status, "device init")) return nullptr;
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

void release_pipe(HostPipe& pipe) {
    for (Slot& s : pipe.slot) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_aux) cudaFree(s.d_aux);
        if (s.d_scratch) cudaFree(s.d_scratch);
        if (s.h_pin) cudaFreeHost(s.h_pin);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = Slot();
    }
    if (pipe.d_idx) cudaFree(pipe.d_idx);
    pipe.d_idx = nullptr;
    pipe.idx_cap = 0;
    pipe.device = -1;
    (void)cudaGetLastError();  // a thread that outlives the CUDA context frees nothing; that is fine
}

HostPipe::~HostPipe() {
    if (device >= 0 && cudaSetDevice(device) == cudaSuccess) release_pipe(*this);
}

bool prepare_pipe(int dev) {
    if (g_pipe.device != dev) {
        if (g_pipe.device >= 0) {
            cudaSetDevice(g_pipe.device);
            release_pipe(g_pipe);
            cudaSetDevice(dev);
        }
        g_pipe.device = dev;
    }
    for (Slot& s : g_pipe.slot)
        if (!s.stream && fail(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    return true;
}

// Host-path chunk size in frames per pipeline stage.  The throughput kernel takes a flat ~0.38 us x (F+6)
// for anything up to ~37,000 frames, the H2D copy of n frames takes n x 4(F+6) B / 55 GB/s, so the
// kernel hides behind the copy once a chunk holds more than ~5,300 frames whatever F is; 12,288
// leaves a 2x margin (profiles/e2e_chunk_sweep.py).  VITERBI_B200_CHUNK_FRAMES overrides it.
size_t host_chunk_frames() {
    static const size_t frames = [] {
        const char* env = getenv("VITERBI_B200_CHUNK_FRAMES");
        const long v = (env && *env) ? atol(env) : 12288;
        return (size_t)(v >= 64 ? v : 12288) & ~(size_t)63;
    }();
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

// Enqueue one batch that is already in device memory.  Scratch is stream-ordered.
int vit_device(DeviceState* st, unsigned framebits, const uint8_t* d_syms, size_t n, uint8_t* d_out,
               cudaStream_t stream, void* scratch, size_t scratch_cap) {
    if (n == 0 || framebits == 0) return FEC_OK;
    const int mode = g_vit_kernel.load();
    if (mode == FEC_VITERBI_WARP || (mode == FEC_VITERBI_AUTO && n < kVitWarpKernelMaxFrames))
        // latency / small-batch path: decisions stay in shared memory
        return fail(launch_viterbi_warp(d_syms, d_out, n, framebits, st->num_sms, stream), "viterbi warp kernel launch")
                   ? FEC_ERR_DEVICE
                   : FEC_OK;
    const int blocks = viterbi_grid_blocks(st->num_sms, n);
    const size_t need = viterbi_scratch_bytes(blocks, framebits);
    void* ws = scratch;
    const bool own = (ws == nullptr) || scratch_cap < need;
    if (own && fail(cudaMallocAsync(&ws, need, stream), "cudaMallocAsync(scratch)")) return FEC_ERR_DEVICE;
    cudaError_t e = launch_viterbi_pair(d_syms, d_out, ws, n, framebits, blocks, stream);
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
    if (punct && (!grow(&g_pipe.d_idx, &g_pipe.idx_cap, nsym * sizeof(int32_t)) ||
                  fail(cudaMemcpy(g_pipe.d_idx, idx.data(), nsym * sizeof(int32_t), cudaMemcpyHostToDevice), "H2D index table")))
        return FEC_ERR_DEVICE;
    // Small calls (the single-frame drop-in above all) bounce through this thread's pinned buffer: the driver's
    // pageable-copy path serialises concurrent callers (4 threads at F=3072 were slower than 1), a 50 KB memcpy
    // into pinned memory does not.
    uint8_t* bounce_out = nullptr;
    uint8_t* const user_out = out;
    const size_t in_bytes = n * in_row, out_bytes = n * nout;
    if (in_bytes + out_bytes <= kBounceBytes) {
        Slot& s0 = g_pipe.slot[0];
        const size_t in_pad = (in_bytes + 255) & ~(size_t)255;
        if (!grow(&s0.h_pin, &s0.pin_cap, in_pad + out_bytes, true)) return FEC_ERR_DEVICE;
        memcpy(s0.h_pin, syms, in_bytes);
        syms = s0.h_pin;
        bounce_out = (uint8_t*)s0.h_pin + in_pad;
        out = bounce_out;
    }
    // ... and when the warp-per-frame kernel would decode them anyway, it runs directly on the bounce buffer:
    // pinned memory is mapped into the device's address space, the kernel stages the symbols into shared memory
    // itself (compacting the u32 layout on the way) and writes the decoded bytes back through the mapping, so
    // the call is one kernel launch and one synchronise -- no copy operations, no compaction kernel.
    if (bounce_out && !punct && n < kVitWarpKernelMaxFrames && g_vit_kernel.load() != FEC_VITERBI_PAIR) {
        Slot& s0 = g_pipe.slot[0];
        const cudaError_t e =
            is_u32 ? launch_viterbi_warp_u32((const uint32_t*)syms, bounce_out, n, framebits, st->num_sms, s0.stream)
                   : launch_viterbi_warp((const uint8_t*)syms, bounce_out, n, framebits, st->num_sms, s0.stream);
        if (fail(e, "viterbi warp kernel launch") || fail(cudaStreamSynchronize(s0.stream), "cudaStreamSynchronize"))
            return FEC_ERR_DEVICE;
        memcpy(user_out, bounce_out, out_bytes);
        return FEC_OK;
    }
    // chunks are pipelined over kPipe streams: the H2D copy of chunk k+1, the kernel of chunk k and the
    // D2H copy of chunk k-1 overlap
    size_t chunk = host_chunk_frames();
    if (chunk > n) chunk = n;
    int rc = FEC_OK;
    size_t done = 0;
    for (int k = 0; done < n && rc == FEC_OK; k++) {
        Slot& s = g_pipe.slot[k % kPipe];
        const size_t m = (n - done < chunk) ? n - done : chunk;
        // the slot's previous chunk must have left its buffers
        if (fail(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize")) { rc = FEC_ERR_DEVICE; break; }
        const int blocks = viterbi_grid_blocks(st->num_sms, m);
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
    return rc;
}

int rs_host(const uint8_t* in, unsigned s, size_t n, uint8_t* out, int32_t* ret) {
    if (s == 0 || s > 1024) return bad_arg("RSDims must be 1..1024");
    if (n == 0) return FEC_OK;
    if (!in || !out || !ret) return bad_arg("null pointer");
    int dev;
    DeviceState* st = device_state(&dev);
    if (!st) return FEC_ERR_DEVICE;
    if (!prepare_pipe(dev)) return FEC_ERR_DEVICE;
    const size_t in_row = 120 * (size_t)s, out_row = 110 * (size_t)s;
    // small calls (the single-superframe drop-in): bounce through this thread's pinned buffer, see vit_host
    uint8_t* bounce = nullptr;
    uint8_t* const user_out = out;
    int32_t* const user_ret = ret;
    const size_t in_bytes = n * in_row, out_bytes = n * out_row, ret_bytes = n * sizeof(int32_t);
    const size_t in_pad = (in_bytes + 255) & ~(size_t)255, out_pad = (out_bytes + 255) & ~(size_t)255;
    if (in_bytes + out_bytes <= kBounceBytes) {
        Slot& s0 = g_pipe.slot[0];
        if (!grow(&s0.h_pin, &s0.pin_cap, in_pad + out_pad + ret_bytes, true)) return FEC_ERR_DEVICE;
        bounce = (uint8_t*)s0.h_pin;
        memcpy(bounce, in, in_bytes);
        memcpy(bounce + in_pad, out, out_bytes);  // the partial-write rule keeps the caller's bytes
        in = bounce;
        out = bounce + in_pad;
        ret = reinterpret_cast<int32_t*>(bounce + in_pad + out_pad);
    }
    size_t chunk = rs_chunk_bytes() / in_row;
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    int rc = FEC_OK;
    size_t done = 0;
    for (int k = 0; done < n && rc == FEC_OK; k++) {
        Slot& sl = g_pipe.slot[k % kPipe];
        const size_t m = (n - done < chunk) ? n - done : chunk;
        if (fail(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize")) { rc = FEC_ERR_DEVICE; break; }
        if (!grow(&sl.d_in, &sl.in_cap, m * in_row) || !grow(&sl.d_out, &sl.out_cap, m * out_row) ||
            !grow(&sl.d_aux, &sl.aux_cap, m * sizeof(int32_t))) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        // the partial-write rule needs the caller's current output bytes on the device
        if (fail(cudaMemcpyAsync(sl.d_in, in + done * in_row, m * in_row, cudaMemcpyHostToDevice, sl.stream), "H2D") ||
            fail(cudaMemcpyAsync(sl.d_out, out + done * out_row, m * out_row, cudaMemcpyHostToDevice, sl.stream), "H2D out") ||
            fail(launch_rs_superframes((const uint8_t*)sl.d_in, (uint8_t*)sl.d_out, (int32_t*)sl.d_aux, m, s, st->num_sms,
                                       sl.stream),
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
    if (bounce && rc == FEC_OK) {
        memcpy(user_out, out, out_bytes);
        memcpy(user_ret, ret, ret_bytes);
    }
    return rc;
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
    if (reinterpret_cast<uintptr_t>(d_syms) & 15) return bad_arg("d_syms must be 16-byte aligned");
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

int rs_check_superframe_batch_device(const uint8_t* d_in, unsigned int RSDims, size_t n, uint8_t* d_out,
                                     int32_t* d_ret, void* stream) {
    if (RSDims == 0 || RSDims > 1024) return bad_arg("RSDims must be 1..1024");
    if (n == 0) return FEC_OK;
    if (!d_in || !d_out || !d_ret) return bad_arg("null pointer");
    DeviceState* st = device_state();
    if (!st) return FEC_ERR_DEVICE;
    return fail(launch_rs_superframes(d_in, d_out, d_ret, n, RSDims, st->num_sms, (cudaStream_t)stream), "rs kernel launch")
               ? FEC_ERR_DEVICE
               : FEC_OK;
}

// Viterbi -> superframe -> RS on the device.  Five consecutive decoded frames ARE one superframe
// ([nsf*5][F/8] == [nsf][120*s] with s = F/192), so no regrouping pass is needed between the kernels.
int dabplus_decode_superframes_device(unsigned int framebits, const uint8_t* d_syms, size_t nsf, uint8_t* d_out,
                                      int32_t* d_ret, void* stream) {
    if (!vit_args_ok(framebits) || framebits == 0 || framebits % 192u) return bad_arg("framebits must be a multiple of 192");
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
    if (rc == FEC_OK && fail(launch_rs_superframes((const uint8_t*)d_bits, d_out, d_ret, nsf, rsdims, st->num_sms, s),
                             "rs kernel launch"))
        rc = FEC_ERR_DEVICE;
    (void)cudaFreeAsync(d_bits, s);
    return rc;
}

int dabplus_decode_superframes(unsigned int framebits, const uint8_t* syms, size_t nsf, uint8_t* out, int32_t* ret) {
    if (!vit_args_ok(framebits) || framebits == 0 || framebits % 192u) return bad_arg("framebits must be a multiple of 192");
    if (nsf == 0) return FEC_OK;
    if (!syms || !out || !ret) return bad_arg("null pointer");
    int dev;
    DeviceState* st = device_state(&dev);
    if (!st) return FEC_ERR_DEVICE;
    if (!prepare_pipe(dev)) return FEC_ERR_DEVICE;
    const unsigned rsdims = framebits / 192u;
    const size_t nsym = 4 * ((size_t)framebits + 6), in_row = 5 * nsym, out_row = 110 * (size_t)rsdims;
    size_t chunk = host_chunk_frames() / 5;
    if (chunk < 1) chunk = 1;
    if (chunk > nsf) chunk = nsf;
    int rc = FEC_OK;
    size_t done = 0;
    for (int k = 0; done < nsf && rc == FEC_OK; k++) {
        Slot& sl = g_pipe.slot[k % kPipe];
        const size_t m = (nsf - done < chunk) ? nsf - done : chunk;
        if (fail(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize")) { rc = FEC_ERR_DEVICE; break; }
        if (!grow(&sl.d_in, &sl.in_cap, m * in_row) || !grow(&sl.d_out, &sl.out_cap, m * out_row) ||
            !grow(&sl.d_aux, &sl.aux_cap, m * sizeof(int32_t))) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        if (fail(cudaMemcpyAsync(sl.d_in, syms + done * in_row, m * in_row, cudaMemcpyHostToDevice, sl.stream), "H2D") ||
            fail(cudaMemcpyAsync(sl.d_out, out + done * out_row, m * out_row, cudaMemcpyHostToDevice, sl.stream), "H2D out")) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        rc = dabplus_decode_superframes_device(framebits, (const uint8_t*)sl.d_in, m, (uint8_t*)sl.d_out, (int32_t*)sl.d_aux,
                                               sl.stream);
        if (rc != FEC_OK) break;
        if (fail(cudaMemcpyAsync(out + done * out_row, sl.d_out, m * out_row, cudaMemcpyDeviceToHost, sl.stream), "D2H") ||
            fail(cudaMemcpyAsync(ret + done, sl.d_aux, m * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream), "D2H ret")) {
            rc = FEC_ERR_DEVICE;
            break;
        }
        done += m;
    }
    for (Slot& sl : g_pipe.slot)
        if (sl.stream && fail(cudaStreamSynchronize(sl.stream), "cudaStreamSynchronize") && rc == FEC_OK) rc = FEC_ERR_DEVICE;
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

int fec_device_synchronize(void) { return fail(cudaDeviceSynchronize(), "cudaDeviceSynchronize") ? FEC_ERR_DEVICE : FEC_OK; }

int fec_set_viterbi_kernel(int mode) {
    if (mode < FEC_VITERBI_AUTO || mode > FEC_VITERBI_WARP) return bad_arg("unknown kernel mode");
    g_vit_kernel.store(mode);
    return FEC_OK;
}

unsigned long long fec_kernel_launches(void) { return g_launches.load(); }

}  // extern "C"
#pragma GCC visibility pop
