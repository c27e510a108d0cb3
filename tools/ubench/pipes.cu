// Pipe-throughput probes for the density scan on sm_100a: scalar FFMA / FADD vs the packed FFMA2 / FADD2 forms,
// and the scan's instruction mix.  Prints warp-instructions (and fp32 lane-ops) per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_exp/pipes tools/ubench/pipes.cu && build_exp/pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ITERS = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, float seed) {
    float a[8];
    f32x2 p[8];
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = seed + k + threadIdx.x; p[k] = pack2(a[k], a[k] + 0.5f); }
    const float c = seed * 0.999f;
    const f32x2 c2 = pack2(c, c);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) a[k] = fmaf(a[k], c, a[k]);                 // 8 FFMA
            if (MODE == 1) p[k] = fma2(p[k], c2, p[k]);                // 8 FFMA2
            if (MODE == 2) a[k] = a[k] + c;                            // 8 FADD
            if (MODE == 3) p[k] = add2(p[k], c2);                      // 8 FADD2
            if (MODE == 4) {                                           // scan mix, scalar: 3 FADD + 3 FFMA + 1 SHF per candidate
                float rx = c - a[k], ry = a[(k + 1) & 7] - c, rz = c - a[(k + 2) & 7];
                float d = fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, -seed)));
                m = __funnelshift_l(__float_as_uint(d), m, 1);
                a[k] = d * 1e-30f + a[k];
            }
            if (MODE == 5) {                                           // scan mix, packed: 3 FADD2 + 3 FFMA2 + 2 SHF per 2 candidates
                f32x2 rx = add2(p[k], c2), ry = add2(p[(k + 1) & 7], c2), rz = add2(p[(k + 2) & 7], c2);
                f32x2 d = fma2(rz, rz, fma2(ry, ry, fma2(rx, rx, c2)));
                m = __funnelshift_l((uint32_t)d, m, 1);
                m = __funnelshift_l((uint32_t)(d >> 32), m, 1);
                p[k] = fma2(d, pack2(1e-30f, 1e-30f), p[k]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + __uint_as_float((uint32_t)p[k]) + __uint_as_float((uint32_t)(p[k] >> 32));
    if (s == 123.456f || m == 0x12345u) out[0] = s;
}
template <int MODE>
void run(const char *name, double inst_per_iter, double flops_per_iter) {
    float *out;
    cudaMalloc(&out, 4);
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;
    probe<MODE><<<blocks, 256>>>(out, 1.0f);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)blocks * 8;
    double winst = warps * ITERS * inst_per_iter;
    double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-28s %8.3f ms  %6.3f warp-inst/cycle/SM  %7.1f lane-flop/cycle/SM (clock %d MHz nominal)\n", name, ms,
           winst / cycles / sms, warps * ITERS * flops_per_iter * 32 / cycles / sms, khz / 1000);
    cudaFree(out);
}
int main() {
    run<0>("FFMA x8", 8, 16);
    run<1>("FFMA2 x8", 8, 32);
    run<2>("FADD x8", 8, 8);
    run<3>("FADD2 x8", 8, 16);
    run<4>("scan mix scalar (8 cand)", 8 * 8, 8 * 9);
    run<5>("scan mix packed (16 cand)", 8 * 9, 16 * 9);
    return 0;
}
