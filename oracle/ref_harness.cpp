// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Glue that lets the reference's own hot-path translation units
// (/root/reference/deconvolve.cpp and /root/reference/rschecksf.cpp, compiled
// UNMODIFIED from where they lie by oracle/Makefile) run on Linux as
// oracle/_ref/libviterbi_ref_*.so.  It supplies only what the reference takes
// from its Windows/MASM side:
//   * the decoder constants the MASM data file exports (const.asm:17-63),
//     generated here from the code polynomials instead of being transcribed;
//   * the GF(256)/symbol lookup tables built at DLL attach (dllmain.cpp:124-150);
//   * the dispatcher pointer `deconJumpTarget` (setupdll.cpp:39);
//   * extern "C" entry points + multi-threaded batch loops for timing.
// Nothing here is linked into, or called by, the product library.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

// ---- symbols the reference expects from const.asm (matched by name only) ----
extern "C" {
alignas(64) unsigned char m256_63_0[32];  // m128_63_0 is the same label (const.asm:19-22)
alignas(16) unsigned char m128_63[16];
alignas(32) unsigned char m256_XOR_0_3_4_7[32];
alignas(16) unsigned char m128_1st_XOR_0_3_4_7[16];
alignas(16) unsigned char m128_2nd_XOR_0_3_4_7[16];
alignas(16) unsigned char m128_XOR_1_5[16];
alignas(16) unsigned char m128_XOR_2_6[16];
alignas(32) unsigned char m256_XOR_1_5[32];
alignas(32) unsigned char m256_XOR_2_6[32];
alignas(16) unsigned char m128_16X_0x1[16];
// alias: the 128-bit init vector shares storage with the 256-bit one
extern unsigned char m128_63_0[16] __attribute__((alias("m256_63_0")));
int* symbols32LUT;  // dllmain.cpp:60
}

// same layout as viterbi.h:101-105
struct RS_LookUp {
    unsigned char RS_ato_mod[768];
    unsigned char RS_iof[256];
};
RS_LookUp* rsLUT;  // dllmain.cpp:43

typedef int DECON(unsigned int, unsigned int*, int, unsigned char*);
DECON* deconJumpTarget;  // setupdll.cpp:39

// the reference's entry points (C++ linkage, named by viterbi.def)
int deconvolve(unsigned int framebits, unsigned int* piData, int inputLength, unsigned char* output);
int RScheckSuperframe(unsigned char* p, int startIx, unsigned int RSDims, unsigned char* outVector);
extern "C" {
DECON decon_sse2_lut32, decon_ssse3, decon_avx, decon_avx2, decon_avx5;
}

namespace {

int parity8(unsigned v) { return __builtin_parity(v & 0xFFu); }

// branch-mask byte for butterfly i and code polynomial k (viterbi-benchmark.cpp:64)
unsigned char tmask(int i, int k) {
    static const int poly[4] = {109, 79, 83, 109};
    return parity8((2 * i) & poly[k]) ? 0xFF : 0x00;
}

// 256-bit constants are stored lane-permuted [0-7][16-23][8-15][24-31] (const.asm:7-8)
int perm256(int b) {
    static const int base[4] = {0, 16, 8, 24};
    return base[b >> 3] + (b & 7);
}

struct Init {
    Init() {
        m256_63_0[0] = 0;
        for (int i = 1; i < 32; i++) m256_63_0[i] = 63;  // bytes 16..31 == m128_63 in the asm layout
        for (int i = 0; i < 16; i++) {
            m128_63[i] = 63;
            m128_16X_0x1[i] = 1;
            m128_1st_XOR_0_3_4_7[i] = tmask(i, 0);
            m128_2nd_XOR_0_3_4_7[i] = tmask(i + 16, 0);
            m128_XOR_1_5[i] = tmask(i, 1);  // period 16 in i
            m128_XOR_2_6[i] = tmask(i, 2);
        }
        for (int b = 0; b < 32; b++) {
            m256_XOR_0_3_4_7[b] = tmask(perm256(b), 0);
            m256_XOR_1_5[b] = tmask(perm256(b), 1);
            m256_XOR_2_6[b] = tmask(perm256(b), 2);
        }
        // tables of dllmain.cpp:124-150
        static RS_LookUp lut;
        static int sym32[256];
        unsigned char alpha_to[256];
        lut.RS_iof[0] = 255;
        alpha_to[255] = 0;
        int sr = 1;
        for (int i = 0; i < 255; i++) {
            lut.RS_iof[sr] = (unsigned char)i;
            alpha_to[i] = (unsigned char)sr;
            sr <<= 1;
            if (sr & 256) sr ^= 285;
            sr &= 255;
        }
        for (int i = 0; i < 768; i++) lut.RS_ato_mod[i] = alpha_to[i % 255];
        for (int i = 0; i < 256; i++) sym32[i] = (int)((unsigned)i * 0x01010101u);
        rsLUT = &lut;
        symbols32LUT = sym32;
        deconJumpTarget = decon_avx2;
    }
} g_init;

template <class F>
void run_sharded(size_t n, int nthreads, F&& fn) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        size_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        th.emplace_back([=] { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

// 0 sse2_lut32, 1 ssse3, 2 avx, 3 avx2, 4 avx5 (same numbering as the ini digit, inifiletext.h:12-31)
int ref_select(int which) {
    DECON* tab[5] = {decon_sse2_lut32, decon_ssse3, decon_avx, decon_avx2, decon_avx5};
    if (which < 0 || which > 4) return -1;
    deconJumpTarget = tab[which];
    return 0;
}

int ref_deconvolve(unsigned framebits, unsigned* piData, int inputLength, unsigned char* output) {
    return deconvolve(framebits, piData, inputLength, output);
}

int ref_rs_check_superframe(unsigned char* p, int startIx, unsigned rsdims, unsigned char* outVector) {
    return RScheckSuperframe(p, startIx, rsdims, outVector);
}

// n frames in the reference's one-uint32-per-symbol layout, [n][4*(F+6)]; out [n][ceil(F/8)]
int ref_deconvolve_batch_u32(unsigned framebits, const uint32_t* syms, size_t n, uint8_t* out, int nthreads) {
    const size_t nsym = 4 * ((size_t)framebits + 6), nout = (framebits + 7) / 8;
    run_sharded(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t f = lo; f < hi; f++)
            deconvolve(framebits, const_cast<uint32_t*>(syms) + f * nsym, 0, out + f * nout);
    });
    return 0;
}

// convenience for parity tests: u8 symbols are widened per frame, then decoded
int ref_deconvolve_batch_u8(unsigned framebits, const uint8_t* syms, size_t n, uint8_t* out, int nthreads) {
    const size_t nsym = 4 * ((size_t)framebits + 6), nout = (framebits + 7) / 8;
    run_sharded(n, nthreads, [=](size_t lo, size_t hi) {
        std::vector<uint32_t> w(nsym);
        for (size_t f = lo; f < hi; f++) {
            for (size_t i = 0; i < nsym; i++) w[i] = syms[f * nsym + i];
            deconvolve(framebits, w.data(), 0, out + f * nout);
        }
    });
    return 0;
}

// n superframes of s columns: in [n][120*s], out [n][110*s] (caller pre-fills out), ret [n]
int ref_rs_check_superframe_batch(const uint8_t* in, unsigned s, size_t n, uint8_t* out, int32_t* ret,
                                  int nthreads) {
    run_sharded(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t f = lo; f < hi; f++)
            ret[f] = RScheckSuperframe(const_cast<uint8_t*>(in) + f * 120 * s, 0, s, out + f * 110 * s);
    });
    return 0;
}

}  // extern "C"
