// latbench.cpp -- BASELINE configs[0], the reference's own speed test (viterbi-benchmark.cpp:332-348): repeated
// single-frame deconvolve() calls from one host thread at F = 768 / 1536 / 2304 / 3072, here for the drop-in
// library and, side by side, for the CPU checker (oracle/_ref = the reference's deconvolve.cpp), plus the call
// rate from 8 concurrent threads (README.md:56) and the single-superframe RScheckSuperframe() latency.
//
//   g++ -O2 -std=c++17 -pthread -o latbench latbench.cpp -ldl
//   ./latbench <libviterbi_b200.so> [checker.so] [calls=2000]        -> one JSON object per line
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

typedef int (*deconvolve_t)(unsigned, unsigned*, int, unsigned char*);
typedef int (*rs_t)(unsigned char*, int, unsigned, unsigned char*);

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Stats { double mean, median, p99, min; };
static Stats stats(std::vector<double>& v) {
    std::sort(v.begin(), v.end());
    double s = 0;
    for (double x : v) s += x;
    return {s / v.size(), v[v.size() / 2], v[(size_t)(v.size() * 0.99)], v[0]};
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    void* lib = dlopen(argv[1], RTLD_NOW);
    void* chk = argc > 2 && argv[2][0] ? dlopen(argv[2], RTLD_NOW) : nullptr;
    const int calls = argc > 3 ? atoi(argv[3]) : 2000;
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    auto initialize = (int (*)())dlsym(lib, "initialize");
    auto deco = (deconvolve_t)dlsym(lib, "deconvolve");
    auto rs = (rs_t)dlsym(lib, "RScheckSuperframe");
    deconvolve_t ref = chk ? (deconvolve_t)dlsym(chk, "ref_deconvolve") : nullptr;
    if (chk && !ref) ref = (deconvolve_t)dlsym(chk, "oracle_deconvolve");
    rs_t ref_rs = chk ? (rs_t)dlsym(chk, "ref_rs_check_superframe") : nullptr;
    if (chk && !ref_rs) ref_rs = (rs_t)dlsym(chk, "oracle_rs_check_superframe");
    if (!initialize()) { fprintf(stderr, "initialize failed\n"); return 3; }
    uint64_t seed = 88172645463325252ull;
    auto rnd = [&seed] { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; };
    const unsigned sizes[4] = {768, 1536, 2304, 3072};
    for (unsigned F : sizes) {
        const size_t nsym = 4 * (F + 6);
        const int nbuf = 16;  // distinct frames, not one cache-resident buffer
        std::vector<std::vector<unsigned>> sym(nbuf, std::vector<unsigned>(nsym));
        for (auto& s : sym) {  // random bits -> K=7 rate-1/4 encoder -> AWGN at Eb/N0 = 3 dB -> u8 (viterbi-benchmark.cpp:304-311,658-670)
            static const unsigned poly[4] = {109, 79, 83, 109};
            const double amp = 1.0 / std::sqrt(0.5 / std::pow(10.0, (3.0 + 10.0 * std::log10(0.25)) / 10.0));
            unsigned sr = 0;
            for (unsigned t = 0; t < F + 6; t++) {
                sr = (sr << 1) | (t < F ? (unsigned)(rnd() >> 63) : 0u);
                for (int j = 0; j < 4; j++) {
                    const double u1 = ((rnd() >> 11) + 1) * (1.0 / 9007199254740993.0), u2 = (rnd() >> 11) * (1.0 / 9007199254740992.0);
                    const double g = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
                    const double v = 127.5 + 32.0 * ((__builtin_popcount(sr & poly[j]) & 1 ? amp : -amp) + g);
                    s[4 * t + j] = (unsigned)(v < 0 ? 0 : v > 255 ? 255 : (int)v) | 0xABCD0000u;  // upper bytes are ignored (README.md:19)
                }
            }
        }
        std::vector<unsigned char> out(F / 8), want(F / 8);
        int bad = 0;
        for (int i = 0; i < 50; i++) {
            if (deco(F, sym[i % nbuf].data(), 0, out.data()) != 0) bad++;
            if (ref) { ref(F, sym[i % nbuf].data(), 0, want.data()); bad += memcmp(out.data(), want.data(), F / 8) != 0; }
        }
        std::vector<double> t(calls);
        for (int i = 0; i < calls; i++) {
            const double a = now_us();
            deco(F, sym[i % nbuf].data(), 0, out.data());
            t[i] = now_us() - a;
        }
        Stats g = stats(t);
        Stats c = {0, 0, 0, 0};
        if (ref) {
            for (int i = 0; i < calls; i++) {
                const double a = now_us();
                ref(F, sym[i % nbuf].data(), 0, want.data());
                t[i] = now_us() - a;
            }
            c = stats(t);
        }
        // 8 concurrent callers
        const int nth = 8, per = calls / 4;
        std::vector<std::thread> th;
        const double a8 = now_us();
        for (int k = 0; k < nth; k++)
            th.emplace_back([&, k] {
                std::vector<unsigned char> o(F / 8);
                for (int i = 0; i < per; i++) deco(F, sym[(i + k) % nbuf].data(), 0, o.data());
            });
        for (auto& x : th) x.join();
        const double rate8 = nth * (double)per / ((now_us() - a8) * 1e-6);
        printf("{\"framebits\": %u, \"calls\": %d, \"dropin_us_mean\": %.2f, \"dropin_us_median\": %.2f, \"dropin_us_p99\": %.2f, "
               "\"dropin_us_min\": %.2f, \"cpu_reference_us_mean\": %.2f, \"cpu_reference_us_median\": %.2f, "
               "\"calls_per_s_8_threads\": %.0f, \"mismatches\": %d}\n",
               F, calls, g.mean, g.median, g.p99, g.min, c.mean, c.median, rate8, bad);
        fflush(stdout);
    }
    for (unsigned s : {1u, 8u, 16u}) {
        std::vector<unsigned char> p(120 * s), o(110 * s, 0xEE), o2(110 * s, 0xEE);
        for (auto& b : p) b = (unsigned char)rnd();  // garbage: the expensive (uncorrectable) path
        std::vector<unsigned char> clean(120 * s, 0);  // the all-zero word is a codeword: the cheap path
        int bad = 0;
        for (int i = 0; i < 20; i++) {
            const int r = rs(p.data(), 0, s, o.data());
            if (ref_rs) bad += r != ref_rs(p.data(), 0, s, o2.data()) || memcmp(o.data(), o2.data(), o.size()) != 0;
        }
        std::vector<double> t(calls), tc(calls);
        for (int i = 0; i < calls; i++) {
            double a = now_us();
            rs(p.data(), 0, s, o.data());
            t[i] = now_us() - a;
            a = now_us();
            rs(clean.data(), 0, s, o.data());
            tc[i] = now_us() - a;
        }
        Stats g = stats(t), gc = stats(tc);
        Stats c = {0, 0, 0, 0}, cc = {0, 0, 0, 0};
        if (ref_rs) {
            for (int i = 0; i < calls; i++) {
                double a = now_us();
                ref_rs(p.data(), 0, s, o2.data());
                t[i] = now_us() - a;
                a = now_us();
                ref_rs(clean.data(), 0, s, o2.data());
                tc[i] = now_us() - a;
            }
            c = stats(t), cc = stats(tc);
        }
        printf("{\"rs_dims\": %u, \"dropin_us_median_garbage\": %.2f, \"dropin_us_median_clean\": %.2f, "
               "\"cpu_reference_us_median_garbage\": %.2f, \"cpu_reference_us_median_clean\": %.2f, \"mismatches\": %d}\n",
               s, g.median, gc.median, c.median, cc.median, bad);
    }
    return 0;
}
