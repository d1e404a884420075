// Host check of csrc/rs_chien_bitsliced.h against a plain Chien search (rschecksf.cpp:296-320 restated):
// random and constructed locator polynomials of every degree 1..10.  Exit code 0 = all agree.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../viterbi.dll_b200/csrc/rs_chien_bitsliced.h"

static uint8_t EXP[512], LOG[256];

static uint8_t mul(uint8_t a, uint8_t b) { return (a && b) ? EXP[LOG[a] + LOG[b]] : 0; }

static int plain_chien(const uint32_t* lam, int top, uint32_t* root, int deg) {
    int count = 0;
    for (int i = 1; i <= 255; i++) {
        uint8_t q = 1;
        for (int j = 1; j <= top; j++) q ^= mul((uint8_t)lam[j], EXP[(j * i) % 255]);
        if (q == 0 && count < deg) root[count++] = (uint32_t)i;
    }
    return count;
}

template <int D>
static int run(const uint32_t* lam, uint32_t (&root)[11], int deg) {
    return fec::rsbits::chien_bitsliced<D>(lam, root, deg, [](bool need) { return need; });
}

static int bitsliced(int D, const uint32_t* lam, uint32_t (&root)[11], int deg) {
    switch (D) {
        case 1: return run<1>(lam, root, deg);
        case 2: return run<2>(lam, root, deg);
        case 3: return run<3>(lam, root, deg);
        case 4: return run<4>(lam, root, deg);
        case 5: return run<5>(lam, root, deg);
        case 6: return run<6>(lam, root, deg);
        case 7: return run<7>(lam, root, deg);
        case 8: return run<8>(lam, root, deg);
        case 9: return run<9>(lam, root, deg);
        default: return run<10>(lam, root, deg);
    }
}

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 20000;
    unsigned sr = 1;
    for (int i = 0; i < 255; i++) {
        EXP[i] = (uint8_t)sr;
        LOG[sr] = (uint8_t)i;
        sr <<= 1;
        if (sr & 0x100) sr ^= 0x11D;
    }
    for (int i = 255; i < 512; i++) EXP[i] = EXP[i - 255];
    srand(12345);
    long checked = 0, with_roots = 0;
    for (int D = 1; D <= 10; D++) {
        for (int t = 0; t < trials; t++) {
            uint32_t lam[11] = {1};
            int mode = t % 3;
            if (mode == 0) {  // random coefficients, lane degree <= D (the warp runs the largest degree present)
                int own = 1 + rand() % D;
                for (int j = 1; j <= own; j++) lam[j] = rand() & 0xFF;
            } else {  // product of (1 + X_k x) with distinct roots alpha^{i_k}: fully splitting locator
                int nroots = (mode == 1) ? D : 1 + rand() % D;
                bool used[256] = {false};
                uint8_t poly[11] = {1};
                for (int k = 0; k < nroots; k++) {
                    int i;
                    do i = 1 + rand() % 255; while (used[i]);
                    used[i] = true;
                    uint8_t X = EXP[(255 - i) % 255];  // alpha^{-i}
                    for (int j = k + 1; j >= 1; j--) poly[j] ^= mul(poly[j - 1], X);
                }
                for (int j = 0; j <= 10; j++) lam[j] = poly[j];
            }
            int deg = 0;
            for (int j = 0; j <= 10; j++)
                if (lam[j]) deg = j;
            for (int use_deg : {deg, deg > 1 ? deg - 1 : deg}) {  // the count < deg guard as well
                uint32_t r0[11] = {0}, r1[11] = {0};
                int c0 = plain_chien(lam, D, r0, use_deg);
                int c1 = bitsliced(D, lam, r1, use_deg);
                bool same = c0 == c1;
                for (int k = 0; k < c0 && same; k++) same = r0[k] == r1[k];
                if (!same) {
                    printf("MISMATCH D=%d trial=%d deg=%d plain=%d bitsliced=%d\n", D, t, use_deg, c0, c1);
                    return 1;
                }
                checked++;
                with_roots += c0 > 0;
            }
        }
    }
    printf("ok: %ld searches agree (%ld with roots)\n", checked, with_roots);
    return 0;
}
