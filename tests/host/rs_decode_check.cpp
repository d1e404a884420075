// Host check of csrc/rs_decode.h -- the per-codeword RS(120,110) decoder the kernel runs -- against the oracle
// (oracle/fec_oracle.c, the plain-C restatement of rschecksf.cpp:199-377): same return value, same corrected bytes,
// on encoded codewords with 0..8 byte errors and on random garbage.  usage: rs_decode_check [trials]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../viterbi.dll_b200/csrc/rs_decode.h"

extern "C" int oracle_rs_check_superframe(const unsigned char* p, int startIx, unsigned RSDims, unsigned char* outVector);

struct Row {
    uint32_t x, y, z;
};
static uint8_t ATO[768], IOF[256], ALPHA[255], G[11];
static Row LFSR[256];

struct HostPolicy {
    static uint32_t ato(uint32_t i) { return ATO[i]; }
    static uint32_t iof(uint32_t v) { return IOF[v]; }
    static Row lfsr(uint32_t c) { return LFSR[c]; }
    static bool any(unsigned, bool p) { return p; }
    static unsigned max(unsigned, unsigned v) { return v; }
};

static uint8_t mul(uint8_t a, uint8_t b) { return (a && b) ? ALPHA[(IOF[a] + IOF[b]) % 255] : 0; }

static void build_tables() {  // the construction of rs_upload_tables() in rs_kernels.cu
    unsigned sr = 1;
    IOF[0] = 255;
    for (unsigned i = 0; i < 255; i++) {
        IOF[sr] = (uint8_t)i;
        ALPHA[i] = (uint8_t)sr;
        sr <<= 1;
        if (sr & 0x100u) sr ^= 0x11Du;
    }
    for (unsigned i = 0; i < 768; i++) ATO[i] = ALPHA[i % 255];
    memset(G, 0, sizeof G);
    G[0] = 1;
    for (int i = 0; i < 10; i++) {
        uint8_t nx[11] = {0};
        for (int k = 0; k <= i; k++) {
            nx[k + 1] ^= G[k];
            nx[k] ^= mul(G[k], ALPHA[i]);
        }
        memcpy(G, nx, sizeof nx);
    }
    for (unsigned c = 0; c < 256; c++) {
        uint8_t row[12] = {0};
        for (int k = 0; k < 10; k++) row[k] = mul((uint8_t)c, G[k]);
        LFSR[c].x = row[0] | (row[1] << 8) | (row[2] << 16) | ((uint32_t)row[3] << 24);
        LFSR[c].y = row[4] | (row[5] << 8) | (row[6] << 16) | ((uint32_t)row[7] << 24);
        LFSR[c].z = row[8] | (row[9] << 8);
    }
}

static void encode(const uint8_t* msg, uint8_t* cw) {  // systematic: cw = msg || remainder of msg(x) x^10 mod g
    uint8_t reg[10] = {0};                              // reg[0] = coefficient of x^9
    for (int k = 0; k < 110; k++) {
        const uint8_t fb = msg[k] ^ reg[0];
        for (int j = 0; j < 9; j++) reg[j] = reg[j + 1] ^ mul(fb, G[9 - j]);
        reg[9] = mul(fb, G[0]);
    }
    memcpy(cw, msg, 110);
    memcpy(cw + 110, reg, 10);
}

int main(int argc, char** argv) {
    const long trials = argc > 1 ? atol(argv[1]) : 200000;
    build_tables();
    srand(777);
    long hist[3] = {0, 0, 0};  // clean, corrected, rejected
    for (long t = 0; t < trials; t++) {
        uint8_t cw[120];
        if (t % 16 == 15) {
            for (int k = 0; k < 120; k++) cw[k] = rand() & 0xFF;  // garbage: high locator degrees
        } else {
            uint8_t msg[110];
            for (int k = 0; k < 110; k++) msg[k] = rand() & 0xFF;
            encode(msg, cw);
            const int nerr = rand() % 9;
            bool used[120] = {false};
            for (int e = 0; e < nerr; e++) {
                int pos;
                do pos = rand() % 120; while (used[pos]);
                used[pos] = true;
                cw[pos] ^= (uint8_t)(1 + rand() % 255);
            }
        }
        uint8_t want[110];
        memset(want, 0xEE, sizeof want);
        const int want_ret = oracle_rs_check_superframe(cw, 0, 1, want);
        uint8_t buf[120];
        memcpy(buf, cw, 120);
        const int got_ret = fec::rsdec::rs_decode_column<HostPolicy>(buf, 1, 1u);
        const bool same = got_ret == want_ret && (want_ret < 0 || memcmp(buf, want, 110) == 0);
        if (!same) {
            printf("MISMATCH trial %ld: oracle %d, rs_decode.h %d\n", t, want_ret, got_ret);
            return 1;
        }
        hist[want_ret < 0 ? 2 : want_ret > 0 ? 1 : 0]++;
    }
    printf("ok: %ld codewords agree (%ld clean, %ld corrected, %ld rejected)\n", trials, hist[0], hist[1], hist[2]);
    return 0;
}
