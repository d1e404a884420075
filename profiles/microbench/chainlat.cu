// chainlat.cu -- dependent-chain latencies of the instructions on the critical path of the warp-per-frame Viterbi
// kernel (one warp alone on an SM sub-partition, the situation of the single-frame drop-in call).  Cycles per
// link from clock64 around N dependent repetitions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chainlat chainlat.cu && ./chainlat
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

constexpr int N = 4096;

template <int OP>
__global__ void __launch_bounds__(32) chain(uint32_t* out, long long* cyc, uint32_t a, uint32_t b) {
    __shared__ uint32_t ring[64];
    for (int i = threadIdx.x; i < 64; i += 32) ring[i] = (i + 1) & 63;
    __syncwarp();
    uint32_t x = threadIdx.x + a, y = b;
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __shfl_xor_sync(0xffffffffu, x, 16);
        if (OP == 1) x = __viaddmin_u16x2(x, a, b);
        if (OP == 2) x = __byte_perm(x, a, 0x5410 + (x & 1));
        if (OP == 3) { asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(b)); }
        if (OP == 4) x = ring[x & 63];
        if (OP == 5) x = __ballot_sync(0xffffffffu, x & 1) + threadIdx.x;
        if (OP == 6) {  // the forward chain of the packed warp kernel: add-min, min, shuffle, permute
            x = __viaddmin_u16x2(x, a, 0x00FF00FFu);
            asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            const uint32_t r = __shfl_xor_sync(0xffffffffu, x, 8);
            x = __byte_perm(x, r, 0x5410);
        }
        if (OP == 7) {  // same plus the renormalisation link (relu add) that follows odd steps
            x = __viaddmin_u16x2(x, a, 0x00FF00FFu);
            asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            const uint32_t r = __shfl_xor_sync(0xffffffffu, x, 8);
            x = __byte_perm(x, r, 0x5410);
            x = __viaddmax_s16x2_relu(x, b, 0u);
        }
        if (OP == 8) x = __funnelshift_r(x, x, a) | b;  // traceback link 1 (shf) + link 2 (lop3)
        if (OP == 9) x = x * a + b;                       // IMAD
        if (OP == 10) x = __dp4a(x, a, b);
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int links, uint32_t* d_out, long long* d_cyc) {
    chain<OP><<<1, 32>>>(d_out, d_cyc, 3, 0x00400040u);
    chain<OP><<<1, 32>>>(d_out, d_cyc, 3, 0x00400040u);
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    printf("{\"chain\": \"%s\", \"cycles_per_iteration\": %.2f, \"links\": %d}\n", name, (double)c / N, links);
}

int main() {
    uint32_t* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, 128);
    cudaMalloc(&d_cyc, 8);
    run<0>("shfl.bfly", 1, d_out, d_cyc);
    run<1>("viaddmnmx.u16x2", 1, d_out, d_cyc);
    run<2>("prmt", 1, d_out, d_cyc);
    run<3>("vimnmx.u16x2", 1, d_out, d_cyc);
    run<4>("lds (pointer chase)", 1, d_out, d_cyc);
    run<5>("vote.ballot + iadd", 2, d_out, d_cyc);
    run<6>("addmin -> min -> shfl -> prmt", 4, d_out, d_cyc);
    run<7>("addmin -> min -> shfl -> prmt -> relu-add", 5, d_out, d_cyc);
    run<8>("shf.r.w -> lop3", 2, d_out, d_cyc);
    run<9>("imad", 1, d_out, d_cyc);
    run<10>("idp.4a", 1, d_out, d_cyc);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
