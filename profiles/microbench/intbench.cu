// profiles/microbench/intbench.cu -- integer-pipe throughput probe for B200 (sm_100a).
//
// MEASURED_PEAKS.json only records HBM and bf16 tensor peaks; the Viterbi ACS kernel is
// bound by the integer ALU / issue rate, so the roofline denominator P_int has to be
// measured on the box.  Each test runs UNROLL independent
// dependency chains per thread of one instruction class (or a mix), on every SM at
// full occupancy, and reports lane-ops/s and lane-ops per SM per clock.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o intbench intbench.cu
//   ./intbench [iters]            -> one JSON object per line
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

constexpr int CHAINS = 8;

enum Op {
    OP_IADD,       // add.u32 (ptxas picks IADD3 / IMAD.IADD)
    OP_LOP3,       // 3-input logic
    OP_IMAD,       // mad.lo
    OP_SHF,        // funnel shift
    OP_PRMT,       // byte permute
    OP_MNMX32,     // min.u32
    OP_VIADD16,    // add.u16x2           (VIADD.16x2?)
    OP_VMNMX16,    // min.u16x2           (VIMNMX.U16x2)
    OP_VMNMX3_16,  // min3 u16x2          (VIMNMX3.U16x2)
    OP_VADDMNMX16, // min(a+b,c) u16x2    (VIADDMNMX.U16x2)
    OP_VADDMNMX32, // min(a+b,c) u32      (VIADDMNMX.U32)
    OP_MIX_IADD_LOP,      // alternate IADD / LOP3
    OP_MIX_VADDMNMX_IMAD, // alternate VIADDMNMX / IMAD
    OP_MIX_VADDMNMX_IADD, // alternate VIADDMNMX / IADD3 (a+b+c form)
    OP_MIX_VADDMNMX_LOP,  // alternate VIADDMNMX / LOP3
    OP_ACS_CORE,          // the 3-op ACS body: t=min(B+mm,255); n=min(A+m,t); d=n+K-t
    OP_SHFL,
    OP_BALLOT,
    OP_HSET2,        // set.eq.f16x2.f16x2 (HSET2.BF)
    OP_HFMA2,        // fma.rn.f16x2
    OP_ACS_FP16FLAG, // proposed ACS core: 2x VIADDMNMX (ALU) + HSET2.BF.EQ + HFMA2 flag accumulate
    OP_ACS_PREDIMAD, // current ACS core: VIADDMNMX + IMAD + VIMNMX(preds) + 2 predicated IMAD
    OP_ACS_PREDFADD, // variant: decision bits accumulated with predicated FADD (fmalite-capable)
    OP_FADD,
    OP_VIMNMX_PRED,       // VIMNMX.U16x2 with predicate outputs only
    OP_ADDMIN_MINPRED,    // VIADDMNMX + VIMNMX.P (2 ALU), predicates dropped
    OP_PRED_IMAD_REG,     // 2 ALU + 2 predicated IMAD with register weight
    OP_PRED_IADD,         // 2 ALU + 2 predicated add with immediate (ptxas picks the pipe)
    OP_VIADD_IMM,         // add.u32 with immediate (VIADD)
    OP_VIADD_IMM_PLUS_ALU,// VIADD + VIADDMNMX
    OP_VIADD_IMM_PLUS_IMAD,// VIADD + IMAD
    OP_COUNT
};

static const char* op_name[OP_COUNT] = {
    "iadd", "lop3", "imad", "shf", "prmt", "mnmx_u32", "viadd_16x2", "vimnmx_u16x2", "vimnmx3_u16x2",
    "viaddmnmx_u16x2", "viaddmnmx_u32", "mix_iadd_lop3", "mix_viaddmnmx_imad", "mix_viaddmnmx_iadd3",
    "mix_viaddmnmx_lop3", "acs_core_3op", "shfl_xor", "ballot", "hset2_bf_eq", "hfma2", "acs_fp16flag_4op",
    "acs_predimad_4op", "acs_predfadd_4op", "fadd", "vimnmx_pred", "addmin_minpred_2op", "acs_predimad_reg_4op",
    "acs_prediadd_4op", "viadd_imm", "viadd_imm+viaddmnmx", "viadd_imm+imad"};
// SASS instructions per chain-iteration (ptxas fuses two dependent add/min steps into one
// IADD3 / VIMNMX3, hence 0.5 for those three)
static const double op_count[OP_COUNT] = {0.5, 1, 1, 1, 1, 0.5, 1, 0.5, 1, 1, 1, 2, 2, 2, 2, 3, 1, 3, 1, 1, 4, 4, 4, 1, 1, 2, 4, 4, 1, 2, 2};

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t x, uint32_t a, uint32_t b) {
    // asm volatile keeps nvcc from folding the repeated chain step algebraically
    if (OP == OP_IADD) { asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a)); return x; }
    if (OP == OP_LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(a), "r"(b)); return x; }
    if (OP == OP_IMAD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b)); return x; }
    if (OP == OP_SHF) return __funnelshift_l(x, a, 7);
    if (OP == OP_PRMT) return __byte_perm(x, a, 0x2103);
    if (OP == OP_MNMX32) { asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(a)); return x; }
    if (OP == OP_VIADD16) return __vadd2(x, a);
    if (OP == OP_VMNMX16) { asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(a)); return x; }
    if (OP == OP_VMNMX3_16) return __vimin3_u16x2(x, a, b);
    if (OP == OP_VADDMNMX16) return __viaddmin_u16x2(x, a, b);
    if (OP == OP_VADDMNMX32) return __viaddmin_u32(x, a, b);
    if (OP == OP_MIX_IADD_LOP) {
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(a), "r"(b));
        return x;
    }
    if (OP == OP_MIX_VADDMNMX_IMAD) {
        x = __viaddmin_u16x2(x, a, b);
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
        return x;
    }
    if (OP == OP_MIX_VADDMNMX_IADD) return __viaddmin_u16x2(x, a, b) + a + b;
    if (OP == OP_MIX_VADDMNMX_LOP) {
        x = __viaddmin_u16x2(x, a, b);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(a), "r"(b));
        return x;
    }
    if (OP == OP_ACS_CORE) {
        uint32_t t = __viaddmin_u16x2(x, b, 0x00FF00FFu);
        uint32_t n = __viaddmin_u16x2(a, b, t);
        return n + 0x80008000u - t;
    }
    if (OP == OP_HSET2) { asm volatile("set.eq.f16x2.f16x2 %0, %0, %1;" : "+r"(x) : "r"(a)); return x; }
    if (OP == OP_HFMA2) { asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b)); return x; }
    if (OP == OP_VIADD_IMM) { asm volatile("add.u32 %0, %0, 0x12345;" : "+r"(x)); return x; }
    if (OP == OP_VIADD_IMM_PLUS_ALU) { asm volatile("add.u32 %0, %0, 0x12345;" : "+r"(x)); return __viaddmin_u16x2(x, a, b); }
    if (OP == OP_VIADD_IMM_PLUS_IMAD) {
        asm volatile("add.u32 %0, %0, 0x12345;" : "+r"(x));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
        return x;
    }
    if (OP == OP_SHFL) return __shfl_xor_sync(0xffffffffu, x, 1);
    if (OP == OP_BALLOT) return __ballot_sync(0xffffffffu, (int)x < 0) + a;
    return x;
}

__device__ __forceinline__ uint32_t acs_fp16flag(uint32_t x, uint32_t a, uint32_t b, uint32_t& acc) {
    uint32_t t = __viaddmin_u16x2(x, b, 0x0FF00FF0u);
    uint32_t n = __viaddmin_u16x2(a, b, t);
    uint32_t f;
    asm volatile("set.eq.f16x2.f16x2 %0, %1, %2;" : "=r"(f) : "r"(n), "r"(t));
    asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(f), "r"(0x40004000u));
    return n;
}
__device__ __forceinline__ uint32_t acs_predimad(uint32_t x, uint32_t a, uint32_t b, uint32_t& accA, uint32_t& accB,
                                                 uint32_t one) {
    uint32_t t = __viaddmin_u16x2(x, b, 0x0FF00FF0u);
    uint32_t m0, n;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(m0) : "r"(a), "r"(one), "r"(b));
    asm volatile("{.reg .pred pu, pv; .reg .u16 rs0, rs1, rs2, rs3;\n\t"
                 "min.u16x2 %0, %3, %4;\n\t"
                 "mov.b32 {rs0, rs1}, %0;\n\t"
                 "mov.b32 {rs2, rs3}, %3;\n\t"
                 "setp.eq.u16 pv, rs0, rs2;\n\t"
                 "setp.eq.u16 pu, rs1, rs3;\n\t"
                 "@pv mad.lo.u32 %1, %5, 2, %1;\n\t"
                 "@pu mad.lo.u32 %2, %5, 2, %2;}\n\t"
                 : "=r"(n), "+r"(accA), "+r"(accB)
                 : "r"(t), "r"(m0), "r"(one));
    return n;
}

__device__ __forceinline__ uint32_t acs_predfadd(uint32_t x, uint32_t a, uint32_t b, float& accA, float& accB,
                                                 uint32_t one) {
    uint32_t t = __viaddmin_u16x2(x, b, 0x0FF00FF0u);
    uint32_t m0, n;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(m0) : "r"(a), "r"(one), "r"(b));
    asm volatile("{.reg .pred pu, pv; .reg .u16 rs0, rs1, rs2, rs3;\n\t"
                 "min.u16x2 %0, %3, %4;\n\t"
                 "mov.b32 {rs0, rs1}, %0;\n\t"
                 "mov.b32 {rs2, rs3}, %3;\n\t"
                 "setp.eq.u16 pv, rs0, rs2;\n\t"
                 "setp.eq.u16 pu, rs1, rs3;\n\t"
                 "@pv add.rn.f32 %1, %1, 0f40000000;\n\t"
                 "@pu add.rn.f32 %2, %2, 0f40000000;}\n\t"
                 : "=r"(n), "+f"(accA), "+f"(accB)
                 : "r"(t), "r"(m0));
    return n;
}

#define MINPRED_HEAD                                         \
    "{.reg .pred pu, pv; .reg .u16 rs0, rs1, rs2, rs3;\n\t" \
    "min.u16x2 %0, %3, %4;\n\t"                             \
    "mov.b32 {rs0, rs1}, %0;\n\t"                           \
    "mov.b32 {rs2, rs3}, %3;\n\t"                           \
    "setp.eq.u16 pv, rs0, rs2;\n\t"                         \
    "setp.eq.u16 pu, rs1, rs3;\n\t"
template <int MODE>
__device__ __forceinline__ uint32_t acs_generic(uint32_t x, uint32_t a, uint32_t b, uint32_t& accA, uint32_t& accB,
                                                uint32_t one, uint32_t wreg) {
    uint32_t n;
    if (MODE == 0) {  // VIMNMX.P only; each predicate has one cheap consumer so it is not optimised away
        asm volatile(MINPRED_HEAD "@pv mov.u32 %1, %3;\n\t@pu mov.u32 %2, %3;}\n\t"
                     : "=r"(n), "+r"(accA), "+r"(accB) : "r"(x), "r"(a));
        return n;
    }
    uint32_t t = __viaddmin_u16x2(x, b, 0x0FF00FF0u);
    if (MODE == 1) {
        asm volatile(MINPRED_HEAD "}\n\t" : "=r"(n), "+r"(accA), "+r"(accB) : "r"(t), "r"(a));
    } else if (MODE == 2) {
        asm volatile(MINPRED_HEAD "@pv mad.lo.u32 %1, %5, %6, %1;\n\t@pu mad.lo.u32 %2, %5, %6, %2;}\n\t"
                     : "=r"(n), "+r"(accA), "+r"(accB) : "r"(t), "r"(a), "r"(one), "r"(wreg));
    } else {
        asm volatile(MINPRED_HEAD "@pv add.u32 %1, %1, 2;\n\t@pu add.u32 %2, %2, 2;}\n\t"
                     : "=r"(n), "+r"(accA), "+r"(accB) : "r"(t), "r"(a));
    }
    return n;
}

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, uint32_t seed_a, uint32_t seed_b, int iters,
                                             long long* cycles) {
    uint32_t x[CHAINS];
    uint32_t accs[CHAINS], accs2[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) accs[c] = 0x64006400u, accs2[c] = 0;
    const uint32_t one = (seed_a >> 1);  // seed_a == 3 -> 1, opaque to the compiler
    float fa[CHAINS], fb[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) fa[c] = 8388608.0f, fb[c] = 8388608.0f;
    uint32_t a = seed_a + threadIdx.x, b = seed_b ^ (threadIdx.x * 2654435761u);
    if (OP == OP_MNMX32 || OP == OP_VMNMX16 || OP == OP_VMNMX3_16) { a |= 0xFFF0FFF0u; b |= 0xFF00FF00u; }
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = blockIdx.x * 977u + threadIdx.x * 31u + c * 0x01010101u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (OP == OP_ACS_FP16FLAG) x[c] = acs_fp16flag(x[c], a, b, accs[c]);
                else if (OP == OP_ACS_PREDIMAD) x[c] = acs_predimad(x[c], a, b, accs[c], accs2[c], one);
                else if (OP == OP_ACS_PREDFADD) x[c] = acs_predfadd(x[c], a, b, fa[c], fb[c], one);
                else if (OP == OP_VIMNMX_PRED) x[c] = acs_generic<0>(x[c], a, b, accs[c], accs2[c], one, seed_b);
                else if (OP == OP_ADDMIN_MINPRED) x[c] = acs_generic<1>(x[c], a, b, accs[c], accs2[c], one, seed_b);
                else if (OP == OP_PRED_IMAD_REG) x[c] = acs_generic<2>(x[c], a, b, accs[c], accs2[c], one, seed_b);
                else if (OP == OP_PRED_IADD) x[c] = acs_generic<3>(x[c], a, b, accs[c], accs2[c], one, seed_b);
                else if (OP == OP_FADD) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fa[c]) : "f"(fb[c])); }
                else x[c] = step<OP>(x[c], a, b);
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= x[c] ^ accs[c] ^ accs2[c] ^ __float_as_uint(fa[c]) ^ __float_as_uint(fb[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    (void)cycles;
}

static double wall_now() {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

// Each op runs back to back for at least 0.3 s of wall time so that an NVML sampler running beside this process
// (profiles/int_peaks.py) sees the SM clock and the throttle reasons UNDER THIS LOAD; t_start / t_end let it pick
// the samples that belong to the op.  The rate is wall-clock (CUDA events), per-clock figures are derived by the
// sampler from the NVML clock -- the in-kernel clock64 estimate this file used to print was inconsistent with the
// wall-clock rate and is gone.
template <int OP>
void run(int nsm, int iters, uint32_t* d_out, long long* d_cyc, long long* h_cyc) {
    (void)h_cyc;
    const int blocks = nsm * 8, threads = 256;  // 8 x 256 = 2048 threads/SM = full occupancy
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    bench<OP><<<blocks, threads>>>(d_out, 3, 5, iters / 8 + 1, d_cyc);  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    int reps = 0;
    const double w0 = wall_now();
    do {
        CK(cudaEventRecord(e0));
        bench<OP><<<blocks, threads>>>(d_out, 3, 5, iters, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
        reps++;
    } while (reps < 3 || wall_now() - w0 < 0.3);
    const double w1 = wall_now();
    const double laneops = (double)blocks * threads * (double)iters * 8.0 * CHAINS * op_count[OP];
    const double tops = laneops / (best * 1e-3) / 1e12;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"reps\": %d, \"tera_laneops_per_s\": %.3f, \"t_start\": %.6f, \"t_end\": %.6f}\n",
           op_name[OP], best, reps, tops, w0, w1);
    fflush(stdout);
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, nsm, prop.clockRate);
    uint32_t* d_out;
    long long *d_cyc, *h_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * nsm * 8 * 256));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * nsm * 8));
    h_cyc = (long long*)malloc(sizeof(long long) * nsm * 8);
    run<OP_IADD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_LOP3>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_IMAD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_SHF>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_PRMT>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_MNMX32>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VIADD16>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VMNMX16>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VMNMX3_16>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VADDMNMX16>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VADDMNMX32>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_MIX_IADD_LOP>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_MIX_VADDMNMX_IMAD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_MIX_VADDMNMX_IADD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_MIX_VADDMNMX_LOP>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_ACS_CORE>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_SHFL>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_BALLOT>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_HSET2>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_HFMA2>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_ACS_FP16FLAG>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_ACS_PREDIMAD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_ACS_PREDFADD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_FADD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VIMNMX_PRED>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_ADDMIN_MINPRED>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_PRED_IMAD_REG>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_PRED_IADD>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VIADD_IMM>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VIADD_IMM_PLUS_ALU>(nsm, iters, d_out, d_cyc, h_cyc);
    run<OP_VIADD_IMM_PLUS_IMAD>(nsm, iters, d_out, d_cyc, h_cyc);
    return 0;
}
