// Host check of csrc/viterbi_pair_core.h -- the arithmetic of the two-frames-per-thread Viterbi kernel -- against the
// oracle (oracle/fec_oracle.c, the plain-C restatement of deconvolve.cpp:85-435).  The same step functions the kernel
// inlines (branch metrics in packed u16 arithmetic, ACS on metrics scaled by 16, renormalisation folded into the
// operand fetch, permuted decision bits, two-shift traceback) are run here over whole frames, two frames per "thread"
// as on the device, with the packed DPX instructions emulated half by half.
// usage: viterbi_pair_check [pairs_per_size]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../viterbi.dll_b200/csrc/viterbi_pair_core.h"

extern "C" int oracle_deconvolve(unsigned framebits, const unsigned* piData, int inputLength, unsigned char* output);

using namespace fec;

// the kernel's per-group loop (viterbi_pair_kernel) and traceback, without the memory pipeline around them
static void decode_pair(unsigned F, const uint8_t* symA, const uint8_t* symB, uint8_t* outA, uint8_t* outB) {
    const unsigned steps = F + 6;
    std::vector<uint4> dec(steps);
    uint32_t X[64], Y[64];
    X[0] = 0u;
    for (int s = 1; s < 64; s++) X[s] = kM63;
    uint32_t neg = 0u;
    auto word = [](const uint8_t* p) { return (uint32_t)p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); };
    for (unsigned t = 0; t < steps; t += 2) {
        dec[t] = acs_step<true>(X, Y, word(symA + 4 * t), word(symB + 4 * t), neg);
        dec[t + 1] = acs_step<false>(Y, X, word(symA + 4 * t + 4), word(symB + 4 * t + 4), 0u);
        neg = renorm_addend(X[0]);
    }
    TraceState st;
    int t = (int)F - 1;
    st.sA = dec[t + 6].x, st.sB = dec[t + 6].z;  // start state 0: word lo of the last record
    const int nblk = (int)(F / 32), head = (int)(F % 32);
    for (int i = 0; i < head; i++, t--) trace_step(st, dec[t > 0 ? t + 5 : 6]);
    if (head) {
        const uint32_t vA = vpc::brev(st.hA), vB = vpc::brev(st.hB);
        for (int b = 0; b < (head + 7) / 8; b++) {
            outA[nblk * 4 + b] = (uint8_t)(vA >> (24 - 8 * b));
            outB[nblk * 4 + b] = (uint8_t)(vB >> (24 - 8 * b));
        }
    }
    for (int m = nblk - 1; m >= 0; m--) {
        for (int j = 31; j >= 0; j--, t--) trace_step(st, dec[t > 0 ? t + 5 : 6]);
        const uint32_t wA = trace_word(st.hA), wB = trace_word(st.hB);
        memcpy(outA + 4 * m, &wA, 4);
        memcpy(outB + 4 * m, &wB, 4);
    }
}

// One ACS step against the scalar 8-bit model of SURVEY.md section 8(a) (deconvolve.cpp:334-387, 407-412), on
// random path metrics over the whole 0..255 range: exercises the saturation at 255 and the clamp at 0 of the
// renormalisation directly (whole frames rarely drive the worst states that far).
static bool check_steps(int trials) {
    static const unsigned POLY[4] = {0x6D, 0x4F, 0x53, 0x6D};
    for (int tr = 0; tr < trials; tr++) {
        uint32_t M[64], N[64];
        uint8_t mA[64], mB[64], y[2][4];
        for (int s = 0; s < 64; s++) {
            mA[s] = (tr & 1) ? (uint8_t)(rand() % 256) : (uint8_t)(180 + rand() % 76);
            mB[s] = (tr & 2) ? (uint8_t)(rand() % 256) : (uint8_t)(rand() % 90);
            M[s] = ((uint32_t)mA[s] * 16u) | (((uint32_t)mB[s] * 16u) << 16);
        }
        for (int f = 0; f < 2; f++)
            for (int k = 0; k < 4; k++) y[f][k] = (uint8_t)rand();
        const bool renA = rand() & 1, renB = rand() & 1;
        const uint32_t neg = (renA ? 0xFC10u : 0u) | ((renB ? 0xFC10u : 0u) << 16);
        auto word = [](const uint8_t* p) { return (uint32_t)p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); };
        const uint4 d = acs_step<true>(M, N, word(y[0]), word(y[1]), neg);
        for (int f = 0; f < 2; f++) {
            const uint8_t* m = f ? mB : mA;
            const bool ren = f ? renB : renA;
            for (int i = 0; i < 32; i++) {
                unsigned x[4];
                for (int k = 0; k < 4; k++) x[k] = y[f][k] ^ (__builtin_parity((2u * i) & POLY[k]) ? 0xFFu : 0u);
                const unsigned bm = ((((x[0] + x[1] + 1) >> 1) + ((x[2] + x[3] + 1) >> 1) + 1) >> 1) >> 2, bmm = 63 - bm;
                auto fetch = [&](int s) { return ren ? (m[s] > 63 ? m[s] - 63u : 0u) : (unsigned)m[s]; };
                auto sat = [](unsigned v) { return v > 255 ? 255u : v; };
                const unsigned m0 = sat(fetch(i) + bm), m1 = sat(fetch(i + 32) + bmm);
                const unsigned m2 = sat(fetch(i) + bmm), m3 = sat(fetch(i + 32) + bm);
                const unsigned want[2] = {m1 <= m0 ? m1 : m0, m3 <= m2 ? m3 : m2}, wdec[2] = {m1 <= m0, m3 <= m2};
                for (int o = 0; o < 2; o++) {
                    const int snew = 2 * i + o;
                    const unsigned got = ((N[snew] >> (16 * f)) & 0xFFFFu);
                    const uint32_t words[2] = {f ? d.z : d.x, f ? d.w : d.y};
                    const unsigned gdec = (words[dec_word(snew)] & dec_bit(snew)) != 0;
                    if (got != want[o] * 16 || gdec != wdec[o]) {
                        printf("STEP MISMATCH trial %d frame %d state %d: metric %u (want %u) decision %u (want %u)\n", tr, f, snew,
                               got, want[o] * 16, gdec, wdec[o]);
                        return false;
                    }
                }
            }
        }
    }
    return true;
}

int main(int argc, char** argv) {
    const int pairs = argc > 1 ? atoi(argv[1]) : 40;
    srand(4242);
    if (!check_steps(20000)) return 1;
    long frames = 0;
    const unsigned sizes[] = {768, 3072, 2, 10, 34, 100, 770, 1536};
    for (unsigned F : sizes) {
        const unsigned nsym = 4 * (F + 6), nout = (F + 7) / 8;
        for (int p = 0; p < pairs; p++) {
            std::vector<uint8_t> sym[2] = {std::vector<uint8_t>(nsym), std::vector<uint8_t>(nsym)};
            for (int f = 0; f < 2; f++) {
                const int kind = (p + f) % 8;
                if (kind >= 5) {
                    // encoded traffic (viterbi-benchmark.cpp:304-311 restated: sr = (sr << 1) | bit, code bit j =
                    // parity(sr & poly_j), 6 zero tail bits) + noise.  State 0 stays near the best path, so the
                    // renormalisation (which watches state 0 only) rarely fires and the other metrics saturate at
                    // 255 -- the case random symbols hardly reach.  kind 7: the all-zero message.
                    static const unsigned poly[4] = {109, 79, 83, 109};
                    const int noise = kind == 5 ? 96 : 48;
                    unsigned sr = 0;
                    for (unsigned t = 0; t < F + 6; t++) {
                        const unsigned bit = (t < F && kind != 7) ? (rand() & 1) : 0;
                        sr = (sr << 1) | bit;
                        for (int j = 0; j < 4; j++) {
                            const int level = __builtin_parity(sr & poly[j]) ? 200 : 56;
                            int v = level + rand() % (2 * noise + 1) - noise;
                            sym[f][4 * t + j] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                        }
                    }
                    continue;
                }
                for (unsigned i = 0; i < nsym; i++) {
                    const int r = rand();
                    sym[f][i] = kind == 0   ? (uint8_t)r                             // uniform bytes
                                : kind == 1 ? (uint8_t)((r & 1) ? 255 : 0)           // saturation / clamp heavy
                                : kind == 2 ? (uint8_t)(120 + r % 16)                // tie heavy
                                : kind == 3 ? (uint8_t)128                           // all ties
                                            : (uint8_t)((r % 3) ? 40 + r % 60 : 160 + r % 60);
                }
            }
            std::vector<uint8_t> got[2] = {std::vector<uint8_t>(nout + 4), std::vector<uint8_t>(nout + 4)};
            decode_pair(F, sym[0].data(), sym[1].data(), got[0].data(), got[1].data());
            for (int f = 0; f < 2; f++) {
                std::vector<unsigned> s32(sym[f].begin(), sym[f].end());
                std::vector<uint8_t> want(nout);
                oracle_deconvolve(F, s32.data(), 0, want.data());
                if (memcmp(want.data(), got[f].data(), nout) != 0) {
                    printf("MISMATCH F=%u pair=%d frame=%d\n", F, p, f);
                    return 1;
                }
                frames++;
            }
        }
    }
    printf("ok: 20000 single steps and %ld frames agree\n", frames);
    return 0;
}
