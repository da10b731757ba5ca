// Micro-benchmark: is packed fp32 FMA (fma.rn.f32x2 / FFMA2) faster than scalar FFMA on B200, and does it free issue slots?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>  // 0: scalar FFMA, 1: FFMA2, 2: FFMA + 1 LDS per 8 FMA, 3: FFMA2 + 1 LDS per 4 FFMA2
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = a; sm[threadIdx.x + 256] = b; sm[threadIdx.x + 512] = a; sm[threadIdx.x + 768] = b;
    __syncthreads();
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = i * 0.001f + threadIdx.x;
    float x = a, y = b;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 2 || MODE == 3) { x = sm[(threadIdx.x + it) & 1023]; }
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], x, y);
        } else {
            float2 xx = make_float2(x, x), yy = make_float2(y, y);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float2 t = __ffma2_rn(make_float2(acc[i], acc[i + 1]), xx, yy);
                    acc[i] = t.x; acc[i + 1] = t.y;
                }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int warps_per_sm) {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
    int iters = 20000; int blocks = 148 * warps_per_sm / 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 100, 0.999f, 0.001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 128 * iters * 256.0 * blocks;
    printf("%-28s warps/SM %2d  %.2f TFLOP/s  (%.3f ms)\n", name, warps_per_sm, fl / ms / 1e9, ms);
    cudaFree(out);
}
int main() {
    for (int w : {8, 16, 32}) {
        run<0>("FFMA", w); run<1>("FFMA2", w); run<2>("FFMA + LDS/128fma", w); run<3>("FFMA2 + LDS/128fma", w);
    }
    return 0;
}
