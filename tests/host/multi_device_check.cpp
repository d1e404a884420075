// multi_device_check.cpp -- a plain C++ host (no torch, no CUDA headers) that drives libviterbi_b200.so the way
// a native receiver would: dlopen, fec_set_devices, ONE host batch per call decoded on every selected GPU of the
// box (viterbi_deconvolve_batch_multi / rs_check_superframe_batch_multi / dabplus_decode_superframes_multi),
// then the NCCL gather of device-resident shards (fec_allgather_device).  Every result is compared bit for bit
// with the CPU checker (oracle/_ref = the reference's own deconvolve.cpp / rschecksf.cpp when present, else the
// C port) loaded from the path given on the command line.  TEST CODE: the checker is only the checker.
//
//   multi_device_check <libviterbi_b200.so> <checker.so> [frames_per_device=65536] [max_devices=0 (all)]
//
// Prints one JSON line; exit code 0 = everything bit-exact.
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

template <class T>
T must(void* h, const char* name) {
    void* p = dlsym(h, name);
    if (!p) {
        fprintf(stderr, "missing symbol %s\n", name);
        exit(2);
    }
    return reinterpret_cast<T>(p);
}

struct Rng {  // splitmix64
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    double gauss() {  // Box-Muller
        const double u = uniform() + 1e-300, v = uniform();
        return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v);
    }
};

inline unsigned parity(unsigned v) { return (unsigned)__builtin_popcount(v) & 1u; }

// random bits -> K=7 rate-1/4 encoder (polys 109,79,83,109: viterbi-benchmark.cpp:64,304-311) -> AWGN -> u8
// (viterbi-benchmark.cpp:658-670)
void make_frame(Rng& rng, unsigned framebits, double amp, uint8_t* syms) {
    static const unsigned poly[4] = {109, 79, 83, 109};
    unsigned sr = 0;
    for (unsigned t = 0; t < framebits + 6; t++) {
        const unsigned bit = t < framebits ? (unsigned)(rng.next() >> 63) : 0u;
        sr = (sr << 1) | bit;
        for (int j = 0; j < 4; j++) {
            const double v = 127.5 + 32.0 * ((parity(sr & poly[j]) ? amp : -amp) + rng.gauss());
            syms[4 * t + j] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : (int)v);
        }
    }
}

struct Gf {
    uint8_t exp[512], log[256], gen[11];
    Gf() {
        unsigned sr = 1;
        for (int i = 0; i < 255; i++) {
            exp[i] = exp[i + 255] = (uint8_t)sr;
            log[sr] = (uint8_t)i;
            sr <<= 1;
            if (sr & 0x100) sr ^= 0x11D;
        }
        log[0] = 0;
        memset(gen, 0, sizeof gen);
        gen[0] = 1;  // g(x) = prod (x - a^i), i = 0..9, coefficient of x^k in gen[k]
        for (int i = 0; i < 10; i++) {
            uint8_t nx[11] = {0};
            for (int k = 0; k <= i; k++) {
                nx[k + 1] ^= gen[k];
                nx[k] ^= mul(gen[k], exp[i]);
            }
            memcpy(gen, nx, 11);
        }
    }
    uint8_t mul(uint8_t a, uint8_t b) const { return (a && b) ? exp[log[a] + log[b]] : 0; }
    void encode(const uint8_t* msg, uint8_t* cw) const {  // cw[0..109] = msg, cw[110..119] = parity
        uint8_t rem[10] = {0};
        for (int k = 0; k < 110; k++) {
            const uint8_t fb = msg[k] ^ rem[9];
            for (int j = 9; j > 0; j--) rem[j] = rem[j - 1] ^ mul(fb, gen[j]);
            rem[0] = mul(fb, gen[0]);
        }
        memcpy(cw, msg, 110);
        for (int j = 0; j < 10; j++) cw[110 + j] = rem[9 - j];
    }
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s libviterbi_b200.so checker.so [frames_per_device] [max_devices]\n", argv[0]);
        return 2;
    }
    const size_t per_dev = argc > 3 ? (size_t)atoll(argv[3]) : 65536;
    const int max_dev = argc > 4 ? atoi(argv[4]) : 0;
    void* lib = dlopen(argv[1], RTLD_NOW);
    void* chk = dlopen(argv[2], RTLD_NOW);
    if (!lib || !chk) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    // ---- the product's C ABI (include/viterbi_b200.h) ----
    auto initialize = must<int (*)()>(lib, "initialize");
    auto fec_device_count = must<int (*)()>(lib, "fec_device_count");
    auto fec_set_devices = must<int (*)(const int*, int)>(lib, "fec_set_devices");
    auto fec_get_devices = must<int (*)(int*, int)>(lib, "fec_get_devices");
    auto fec_last_error = must<const char* (*)()>(lib, "fec_last_error");
    auto fec_host_alloc = must<void* (*)(size_t)>(lib, "fec_host_alloc");
    auto fec_host_free = must<void (*)(void*)>(lib, "fec_host_free");
    auto vit_multi = must<int (*)(unsigned, const uint8_t*, size_t, uint8_t*)>(lib, "viterbi_deconvolve_batch_multi");
    auto vit_single = must<int (*)(unsigned, const uint8_t*, size_t, uint8_t*)>(lib, "viterbi_deconvolve_batch");
    auto rs_multi = must<int (*)(const uint8_t*, unsigned, size_t, uint8_t*, int32_t*)>(lib, "rs_check_superframe_batch_multi");
    auto dab_multi = must<int (*)(unsigned, const uint8_t*, size_t, uint8_t*, int32_t*)>(lib, "dabplus_decode_superframes_multi");
    auto fec_set_thread_device = must<int (*)(int)>(lib, "fec_set_thread_device");
    auto fec_device_alloc = must<void* (*)(size_t)>(lib, "fec_device_alloc");
    auto fec_device_free = must<void (*)(void*)>(lib, "fec_device_free");
    auto fec_memcpy_h2d = must<int (*)(void*, const void*, size_t)>(lib, "fec_memcpy_h2d");
    auto fec_memcpy_d2h = must<int (*)(void*, const void*, size_t)>(lib, "fec_memcpy_d2h");
    auto fec_device_synchronize = must<int (*)()>(lib, "fec_device_synchronize");
    auto fec_allgather_device = must<int (*)(const void* const*, void* const*, size_t, void* const*)>(lib, "fec_allgather_device");
    auto fec_kernel_launches = must<unsigned long long (*)()>(lib, "fec_kernel_launches");
    // ---- the checker: compiled reference (ref_*) or C port (oracle_*) ----
    const bool is_ref = dlsym(chk, "ref_deconvolve_batch_u8") != nullptr;
    auto chk_vit = must<int (*)(unsigned, const uint8_t*, size_t, uint8_t*, int)>(
        chk, is_ref ? "ref_deconvolve_batch_u8" : "oracle_deconvolve_batch_u8");
    auto chk_rs = must<int (*)(const uint8_t*, unsigned, size_t, uint8_t*, int32_t*, int)>(
        chk, is_ref ? "ref_rs_check_superframe_batch" : "oracle_rs_check_superframe_batch");
    const int cores = (int)std::thread::hardware_concurrency();

    if (!initialize()) {
        fprintf(stderr, "initialize failed: %s\n", fec_last_error());
        return 3;
    }
    int ndev = fec_device_count();
    if (max_dev > 0 && ndev > max_dev) ndev = max_dev;
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; i++) devs[i] = i;
    if (fec_set_devices(devs.data(), ndev) != 0 || fec_get_devices(nullptr, 0) != ndev) {
        fprintf(stderr, "fec_set_devices failed: %s\n", fec_last_error());
        return 3;
    }

    const unsigned F = 768;  // FIC block; also a DAB+ frame size (F = 192 * 4)
    const size_t n = per_dev * (size_t)ndev + 37;  // ragged tail: the last shard is not 64-aligned
    const size_t row = 4 * (F + 6), nout = F / 8;
    uint8_t* syms = (uint8_t*)fec_host_alloc(n * row);
    uint8_t* out = (uint8_t*)fec_host_alloc(n * nout);
    if (!syms || !out) {
        fprintf(stderr, "fec_host_alloc failed: %s\n", fec_last_error());
        return 3;
    }
    {
        const double amp = 1.0 / std::sqrt(0.5 / std::pow(10.0, (3.0 + 10.0 * std::log10(0.25)) / 10.0));  // Eb/N0 = 3 dB
        std::vector<std::thread> th;
        for (int t = 0; t < cores; t++)
            th.emplace_back([&, t] {
                for (size_t f = n * t / cores; f < n * (t + 1) / cores; f++) {
                    Rng rng{0x1234 + f};
                    make_frame(rng, F, amp, syms + f * row);
                }
            });
        for (auto& x : th) x.join();
    }
    long long bad = 0;
    // ---- 1. Viterbi: one batch over all devices -------------------------------------------------------------
    memset(out, 0xAA, n * nout);
    if (vit_multi(F, syms, n, out) != 0) {  // warm-up: contexts, worker threads, staging buffers
        fprintf(stderr, "viterbi_deconvolve_batch_multi failed: %s\n", fec_last_error());
        return 4;
    }
    const unsigned long long l0 = fec_kernel_launches();
    double t0 = now();
    const int reps = 5;
    for (int r = 0; r < reps; r++)
        if (vit_multi(F, syms, n, out) != 0) return 4;
    const double vit_s = (now() - t0) / reps;
    const unsigned long long launches = fec_kernel_launches() - l0;
    std::vector<uint8_t> want(n * nout);
    t0 = now();
    chk_vit(F, syms, n, want.data(), cores);
    const double cpu_s = now() - t0;
    long long vit_bad = 0;
    for (size_t f = 0; f < n; f++) vit_bad += memcmp(out + f * nout, want.data() + f * nout, nout) != 0;
    bad += vit_bad;
    // the single-device call on the same buffers gives the same bytes
    memset(out, 0x55, n * nout);
    t0 = now();
    if (vit_single(F, syms, n, out) != 0) return 4;
    const double single_s = now() - t0;
    bad += memcmp(out, want.data(), n * nout) != 0;

    // ---- 2. DAB+ chain: five decoded frames = one superframe (s = 4) -> RS check -------------------------------
    const size_t nsf = n / 5;
    const unsigned s = F / 192;
    std::vector<uint8_t> dab_want(nsf * 110 * s, 0xEE);
    std::vector<int32_t> dab_ret_want(nsf), dab_ret(nsf, 12345);
    chk_rs(want.data(), s, nsf, dab_want.data(), dab_ret_want.data(), cores);
    uint8_t* dab_out = (uint8_t*)fec_host_alloc(nsf * 110 * s);
    memset(dab_out, 0xEE, nsf * 110 * s);
    if (dab_multi(F, syms, nsf, dab_out, dab_ret.data()) != 0) {
        fprintf(stderr, "dabplus_decode_superframes_multi failed: %s\n", fec_last_error());
        return 5;
    }
    const long long dab_bad = (memcmp(dab_out, dab_want.data(), dab_want.size()) != 0) +
                              (memcmp(dab_ret.data(), dab_ret_want.data(), nsf * sizeof(int32_t)) != 0);
    bad += dab_bad;

    // ---- 3. RS: encoded superframes with 0..7 byte errors per codeword, pinned and pageable outVector ----------
    const unsigned rs_s = 5;
    const size_t rs_n = (per_dev / 8) * (size_t)ndev + 3;
    uint8_t* rs_in = (uint8_t*)fec_host_alloc(rs_n * 120 * rs_s);
    uint8_t* rs_out = (uint8_t*)fec_host_alloc(rs_n * 110 * rs_s);
    {
        Gf gf;
        Rng rng{77};
        uint8_t msg[110], cw[120];
        for (size_t i = 0; i < rs_n; i++)
            for (unsigned j = 0; j < rs_s; j++) {
                for (auto& b : msg) b = (uint8_t)rng.next();
                gf.encode(msg, cw);
                const int nerr = (int)(rng.next() % 8);
                for (int e = 0; e < nerr; e++) cw[rng.next() % 120] ^= (uint8_t)(1 + rng.next() % 255);
                for (int k = 0; k < 120; k++) rs_in[i * 120 * rs_s + j + (size_t)k * rs_s] = cw[k];
            }
    }
    std::vector<uint8_t> rs_want(rs_n * 110 * rs_s, 0xEE), rs_pageable(rs_n * 110 * rs_s, 0xEE);
    std::vector<int32_t> rs_ret_want(rs_n), rs_ret(rs_n, 777);
    chk_rs(rs_in, rs_s, rs_n, rs_want.data(), rs_ret_want.data(), cores);
    memset(rs_out, 0xEE, rs_n * 110 * rs_s);
    if (rs_multi(rs_in, rs_s, rs_n, rs_out, rs_ret.data()) != 0) {
        fprintf(stderr, "rs_check_superframe_batch_multi failed: %s\n", fec_last_error());
        return 6;
    }
    long long rs_bad = (memcmp(rs_out, rs_want.data(), rs_want.size()) != 0) +
                       (memcmp(rs_ret.data(), rs_ret_want.data(), rs_n * sizeof(int32_t)) != 0);
    std::fill(rs_ret.begin(), rs_ret.end(), 777);
    if (rs_multi(rs_in, rs_s, rs_n, rs_pageable.data(), rs_ret.data()) != 0) return 6;
    rs_bad += (rs_pageable != rs_want) + (memcmp(rs_ret.data(), rs_ret_want.data(), rs_n * sizeof(int32_t)) != 0);
    bad += rs_bad;
    long long rs_failed = 0;
    for (int32_t r : rs_ret_want) rs_failed += r < 0;

    // ---- 4. NCCL gather of device-resident shards, from this one process -----------------------------------------
    long long gather_bad = 0;
    double gather_ms = -1;
    {
        const size_t shard = per_dev * nout;  // one device's decoded bytes
        std::vector<void*> d_shard(ndev), d_all(ndev);
        for (int i = 0; i < ndev; i++) {
            fec_set_thread_device(devs[i]);
            d_shard[i] = fec_device_alloc(shard);
            d_all[i] = fec_device_alloc(shard * ndev);
            if (!d_shard[i] || !d_all[i] || fec_memcpy_h2d(d_shard[i], want.data() + (size_t)i * shard, shard) != 0) {
                fprintf(stderr, "device staging failed: %s\n", fec_last_error());
                return 7;
            }
        }
        const int rc = fec_allgather_device(d_shard.data(), d_all.data(), shard, nullptr);
        if (rc != 0) {
            fprintf(stderr, "fec_allgather_device failed: %s\n", fec_last_error());
            return 7;
        }
        for (int i = 0; i < ndev; i++) {
            fec_set_thread_device(devs[i]);
            fec_device_synchronize();
        }
        t0 = now();
        fec_allgather_device(d_shard.data(), d_all.data(), shard, nullptr);
        for (int i = 0; i < ndev; i++) {
            fec_set_thread_device(devs[i]);
            fec_device_synchronize();
        }
        gather_ms = (now() - t0) * 1e3;
        std::vector<uint8_t> back(shard * ndev);
        for (int i = 0; i < ndev; i++) {
            fec_set_thread_device(devs[i]);
            fec_memcpy_d2h(back.data(), d_all[i], shard * ndev);
            gather_bad += memcmp(back.data(), want.data(), shard * ndev) != 0;
            fec_device_free(d_shard[i]);
            fec_device_free(d_all[i]);
        }
        fec_set_thread_device(-1);
        bad += gather_bad;
    }

    printf("{\"devices\": %d, \"frames\": %zu, \"framebits\": %u, \"viterbi_multi_gbit_per_s\": %.3f, "
           "\"viterbi_single_device_gbit_per_s\": %.3f, \"cpu_checker_gbit_per_s\": %.3f, \"checker\": \"%s\", "
           "\"cpu_threads\": %d, \"kernel_launches_per_call\": %.1f, \"viterbi_mismatched_frames\": %lld, "
           "\"dabplus_superframes\": %zu, \"dabplus_mismatch\": %lld, \"rs_superframes\": %zu, \"rs_failed\": %lld, "
           "\"rs_mismatch\": %lld, \"allgather_bytes_per_device\": %zu, \"allgather_ms\": %.3f, \"allgather_mismatch\": %lld, "
           "\"ok\": %s}\n",
           ndev, n, F, n * (double)F / vit_s / 1e9, n * (double)F / single_s / 1e9, n * (double)F / cpu_s / 1e9,
           is_ref ? "reference" : "port", cores, launches / (double)reps, vit_bad, nsf, dab_bad, rs_n, rs_failed, rs_bad,
           per_dev * nout * (size_t)ndev, gather_ms, gather_bad, bad == 0 ? "true" : "false");
    fec_host_free(syms);
    fec_host_free(out);
    fec_host_free(dab_out);
    fec_host_free(rs_in);
    fec_host_free(rs_out);
    return bad == 0 ? 0 : 1;
}
