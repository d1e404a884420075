/* oracle/fec_oracle.c -- plain-C restatement of the reference's FEC hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see fec_oracle.h).  Written from the semantics of
 * the reference (natural state order, scalar loops); every function cites the
 * reference lines it follows.  Parity PINNED: tests/test_oracle.py checks this
 * file against tests/golden/kat.json and, when present, oracle/_ref.
 */
#include "fec_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Viterbi: K=7, rate 1/4, DAB mother code                                   */
/* ------------------------------------------------------------------------- */

#define NSTATES 64
#define RENORM_THRESHOLD 150 /* viterbi.h:86 */
#define MAX_FRAMEBITS 9216   /* decision array 384*24+6, deconvolve.cpp:127 */

/* code polynomials as used by the encoder, viterbi-benchmark.cpp:64 */
static const unsigned k_poly[4] = {109, 79, 83, 109};

/* Branch mask for butterfly i, polynomial k: 0xFF where the encoder output for
 * the transition old-state i -> new-state 2i is 1 (const.asm:35-49 holds these
 * bytes as data; here they follow from the polynomial). */
static uint8_t branch_mask(unsigned i, unsigned k) {
    return __builtin_parity((2u * i) & k_poly[k]) ? 0xFF : 0x00;
}

static inline unsigned avg_up(unsigned a, unsigned b) { return (a + b + 1u) >> 1; } /* pavgb */
static inline unsigned sat255(unsigned x) { return x > 255u ? 255u : x; }           /* paddusb */

/* One trellis step (one half of Butterfly256, deconvolve.cpp:334-387):
 * old metrics M -> new metrics N, returns the 64 decision bits, bit s = decision
 * of new state s (decision_t layout, viterbi.h:90-92). */
static uint64_t acs_step(const uint8_t M[NSTATES], uint8_t N[NSTATES], const uint8_t y[4],
                         const uint8_t T[4][32]) {
    uint64_t dec = 0;
    for (unsigned i = 0; i < 32; i++) {
        unsigned x0 = y[0] ^ T[0][i], x1 = y[1] ^ T[1][i], x2 = y[2] ^ T[2][i], x3 = y[3] ^ T[3][i];
        unsigned m = avg_up(avg_up(x0, x1), avg_up(x2, x3)) >> 2; /* 0..63, deconvolve.cpp:338-349 */
        unsigned mm = 63u - m;
        unsigned m0 = sat255(M[i] + m), m1 = sat255(M[i + 32] + mm);
        unsigned m2 = sat255(M[i] + mm), m3 = sat255(M[i + 32] + m);
        /* min + cmpeq(survivor, upper-branch): ties choose the i+32 predecessor */
        N[2 * i] = (uint8_t)(m1 <= m0 ? m1 : m0);
        N[2 * i + 1] = (uint8_t)(m3 <= m2 ? m3 : m2);
        dec |= (uint64_t)(m1 <= m0) << (2 * i);
        dec |= (uint64_t)(m3 <= m2) << (2 * i + 1);
    }
    return dec;
}

static int deconvolve_core(unsigned framebits, const uint8_t* s8, const uint32_t* s32, uint8_t* output) {
    if ((framebits & 1u) || framebits > MAX_FRAMEBITS) return -2;
    uint8_t T[4][32];
    for (unsigned k = 0; k < 4; k++)
        for (unsigned i = 0; i < 32; i++) T[k][i] = branch_mask(i, k);

    const unsigned steps = 2u * ((framebits + 6u) / 2u); /* nbits iterations of 2 steps, deconvolve.cpp:126 */
    uint64_t* D = (uint64_t*)malloc(sizeof(uint64_t) * (steps ? steps : 1));
    if (!D) return -3;

    uint8_t A[NSTATES], B[NSTATES];
    A[0] = 0; /* Locals256: metric 0 for the start state, 63 elsewhere (deconvolve.cpp:130-132) */
    for (unsigned s = 1; s < NSTATES; s++) A[s] = 63;

    uint8_t *M = A, *N = B;
    for (unsigned t = 0; t < steps; t++) {
        uint8_t y[4];
        for (unsigned k = 0; k < 4; k++) /* only the low byte of each word counts (deconvolve.cpp:219-228) */
            y[k] = s8 ? s8[4u * t + k] : (uint8_t)(s32[4u * t + k] & 0xFFu);
        D[t] = acs_step(M, N, y, T);
        uint8_t* tmp = M;
        M = N;
        N = tmp;
        /* Renormalize256 runs once per loop iteration = after every second step (deconvolve.cpp:407-412) */
        if ((t & 1u) && M[0] > RENORM_THRESHOLD)
            for (unsigned s = 0; s < NSTATES; s++) M[s] = (uint8_t)(M[s] > 63 ? M[s] - 63 : 0); /* psubusb */
    }

    /* ChainBack (deconvolve.cpp:416-435): start in state 0, skip the 6 tail steps.
     * es keeps state<<2 in an 8-bit register; the byte written last (t%8==0) holds
     * bits 8n..8n+7 MSB-first. */
    unsigned es = 0;
    for (unsigned t = framebits; t-- > 0;) {
        unsigned k = (unsigned)(D[t + 6] >> (es >> 2)) & 1u;
        es = (es >> 1) | (k << 7);
        output[t >> 3] = (uint8_t)es;
    }
    free(D);
    return 0;
}

int oracle_deconvolve(unsigned framebits, const uint32_t* piData, int inputLength, uint8_t* output) {
    (void)inputLength; /* ignored by the reference too */
    return deconvolve_core(framebits, NULL, piData, output);
}

int oracle_deconvolve_u8(unsigned framebits, const uint8_t* syms, uint8_t* output) {
    return deconvolve_core(framebits, syms, NULL, output);
}

/* ------------------------------------------------------------------------- */
/* Reed-Solomon RS(120,110) = RS(255,245) shortened by 135, GF(256)/0x11D     */
/* ------------------------------------------------------------------------- */

#define NN 255
#define NROOTS 10
#define PAD 135 /* rschecksf.cpp:45 */

static uint8_t g_ato[768], g_iof[256];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* dllmain.cpp:124-150: log table with log(0)=255, antilog table of alpha^(i mod 255), i<768 */
static void build_tables(void) {
    uint8_t alpha[255];
    unsigned sr = 1;
    g_iof[0] = NN;
    for (unsigned i = 0; i < NN; i++) {
        g_iof[sr] = (uint8_t)i;
        alpha[i] = (uint8_t)sr;
        sr <<= 1;
        if (sr & 0x100u) sr ^= 0x11Du;
    }
    for (unsigned i = 0; i < 768; i++) g_ato[i] = alpha[i % NN];
}

void oracle_rs_tables(uint8_t ato_mod[768], uint8_t index_of[256]) {
    pthread_once(&g_once, build_tables);
    memcpy(ato_mod, g_ato, 768);
    memcpy(index_of, g_iof, 256);
}

/* rschecksf.cpp:50-52: u32 wrap-around multiply; equals x % 255 for 0 <= x < 66299 */
static inline unsigned mod255(unsigned x) { return (x * 0x1010102u) >> 24; }

int oracle_rs_decode_codeword(uint8_t cw[120]) {
    pthread_once(&g_once, build_tables);
    const uint8_t *ato = g_ato, *iof = g_iof;
    uint8_t syn[NROOTS + 1], lam[NROOTS + 1], bpoly[NROOTS + 1], nxt[NROOTS + 1], root[NROOTS + 1];

    /* syndromes S_i = cw(alpha^i) by Horner, cw[0] highest degree (rschecksf.cpp:208-219) */
    unsigned any = 0;
    for (unsigned i = 0; i < NROOTS; i++) {
        unsigned acc = cw[0];
        for (unsigned j = 1; j < NN - PAD; j++) acc = cw[j] ^ (acc ? ato[iof[acc] + i] : 0u);
        syn[i] = (uint8_t)acc;
        any |= acc;
    }
    if (!any) return 0; /* rschecksf.cpp:224-230 */
    for (unsigned i = 0; i < NROOTS; i++) syn[i] = iof[syn[i]]; /* to index form (232-233) */

    /* Berlekamp-Massey, lambda in polynomial form, b and syn in index form (236-284) */
    memset(lam, 0, sizeof lam);
    lam[0] = 1;
    memset(bpoly, NN, sizeof bpoly);
    bpoly[0] = 0;
    int el = 0;
    for (int r = 1; r <= NROOTS; r++) {
        unsigned discr = 0;
        for (int i = 0; i < r; i++)
            if (lam[i] != 0 && syn[r - i - 1] != NN) discr ^= ato[iof[lam[i]] + syn[r - i - 1]];
        discr = iof[discr];
        if (discr == NN) {
            memmove(bpoly + 1, bpoly, NROOTS); /* B(x) <- x B(x) */
            bpoly[0] = NN;
            continue;
        }
        nxt[0] = lam[0]; /* T(x) = lambda(x) - discr * x * B(x) */
        for (int i = 0; i < NROOTS; i++) {
            nxt[i + 1] = lam[i + 1];
            if (bpoly[i] != NN) nxt[i + 1] ^= ato[discr + bpoly[i]];
        }
        if (2 * el <= r - 1) {
            el = r - el;
            for (int i = 0; i <= NROOTS; i++) /* B(x) <- lambda(x) / discr */
                bpoly[i] = (uint8_t)(lam[i] == 0 ? NN : mod255(iof[lam[i]] - discr + NN));
        } else {
            memmove(bpoly + 1, bpoly, NROOTS);
            bpoly[0] = NN;
        }
        memcpy(lam, nxt, sizeof lam);
    }

    int deg_lambda = 0;
    for (int i = 0; i <= NROOTS; i++) {
        lam[i] = iof[lam[i]];
        if (lam[i] != NN) deg_lambda = i;
    }

    /* Chien search over i = 1..255 with cumulative exponent update and early exit (296-320) */
    memcpy(bpoly, lam, sizeof lam);
    int count = 0;
    for (int i = 1; i <= NN; i++) {
        unsigned q = 1;
        for (int j = deg_lambda; j > 0; j--)
            if (bpoly[j] != NN) {
                bpoly[j] = (uint8_t)mod255(bpoly[j] + j);
                q ^= ato[bpoly[j]];
            }
        if (q) continue;
        root[count] = (uint8_t)i;
        if (++count == deg_lambda) break;
    }
    if (deg_lambda != count) return -1; /* 325-326 */

    /* omega(x) = syn(x) lambda(x) mod x^nroots, index form, kept in bpoly (331-341) */
    int deg_omega = deg_lambda - 1;
    for (int i = 0; i <= deg_omega; i++) {
        unsigned tmp = 0;
        for (int j = i; j >= 0; j--)
            if (syn[i - j] != NN && lam[j] != NN) tmp ^= ato[syn[i - j] + lam[j]];
        bpoly[i] = iof[tmp];
    }

    /* Forney (346-374).  Roots inside the virtual padding are counted, not applied;
     * the error-value exponent is used unreduced (table has 768 entries); no den==0 test. */
    for (int j = count - 1; j >= 0; j--) {
        unsigned rj = root[j];
        if (rj < PAD + 1) continue;
        unsigned num1 = 0;
        for (int i = deg_omega; i >= 0; i--)
            if (bpoly[i] != NN) num1 ^= ato[mod255(bpoly[i] + (unsigned)i * rj)];
        if (!num1) continue;
        unsigned num2 = ato[NN - rj];
        unsigned den = 0;
        int top = deg_lambda < NROOTS - 1 ? deg_lambda : NROOTS - 1;
        for (int i = top & ~1; i >= 0; i -= 2)
            if (lam[i + 1] != NN) den ^= ato[mod255(lam[i + 1] + (unsigned)i * rj)];
        cw[rj - 1 - PAD] ^= ato[iof[num1] + iof[num2] + (NN - iof[den])];
    }
    return count;
}

int oracle_rs_check_superframe(const uint8_t* p, int startIx, unsigned rsdims, uint8_t* outVector) {
    (void)startIx; /* unused in the reference (rschecksf.cpp:69) */
    int errors = 0;
    uint8_t cw[120];
    for (size_t j = 0; j < rsdims; j++) {
        for (size_t k = 0; k < 120; k++) cw[k] = p[j + k * rsdims]; /* column gather (75-76) */
        int r = oracle_rs_decode_codeword(cw);
        if (r == -1) return -1; /* this and all later columns stay untouched (85-88) */
        errors += r;
        for (size_t k = 0; k < 110; k++) outVector[j + k * rsdims] = cw[k];
    }
    return errors;
}

/* ------------------------------------------------------------------------- */
/* batch helpers (pthreads)                                                   */
/* ------------------------------------------------------------------------- */

typedef struct {
    int kind; /* 0 viterbi u8, 1 rs */
    unsigned param;
    const uint8_t* in;
    uint8_t* out;
    int32_t* ret;
    size_t lo, hi;
    int status;
} job_t;

static void* job_main(void* arg) {
    job_t* j = (job_t*)arg;
    if (j->kind == 0) {
        const size_t nsym = 4 * ((size_t)j->param + 6), nout = (j->param + 7) / 8;
        for (size_t f = j->lo; f < j->hi; f++) {
            int r = oracle_deconvolve_u8(j->param, j->in + f * nsym, j->out + f * nout);
            if (r) j->status = r;
        }
    } else {
        const size_t s = j->param;
        for (size_t f = j->lo; f < j->hi; f++)
            j->ret[f] = oracle_rs_check_superframe(j->in + f * 120 * s, 0, j->param, j->out + f * 110 * s);
    }
    return NULL;
}

static int run_jobs(int kind, unsigned param, const uint8_t* in, size_t n, uint8_t* out, int32_t* ret,
                    int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)nthreads);
    int status = 0;
    for (int t = 0; t < nthreads; t++) {
        job_t j = {kind, param, in, out, ret, n * (size_t)t / (size_t)nthreads,
                   n * ((size_t)t + 1) / (size_t)nthreads, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].status) status = jobs[t].status;
    }
    free(th);
    free(jobs);
    return status;
}

int oracle_deconvolve_batch_u8(unsigned framebits, const uint8_t* syms, size_t n, uint8_t* out, int nthreads) {
    if ((framebits & 1u) || framebits > MAX_FRAMEBITS) return -2;
    return run_jobs(0, framebits, syms, n, out, NULL, nthreads);
}

int oracle_rs_check_superframe_batch(const uint8_t* in, unsigned s, size_t n, uint8_t* out, int32_t* ret,
                                     int nthreads) {
    return run_jobs(1, s, in, n, out, ret, nthreads);
}
