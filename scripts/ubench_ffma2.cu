// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) throughput on sm_100a, register-resident.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_ffma2 scripts/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

template <bool PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s0) {
    float a[8], b[8], acc[8][8];
    for (int i = 0; i < 8; ++i) { a[i] = s0 + threadIdx.x * 1e-3f + i; b[i] = s0 * 0.5f + i * 0.25f; }
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) {
                const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    float2 r = fma2(aa, make_float2(b[j], b[j + 1]), make_float2(acc[i][j], acc[i][j + 1]));
                    acc[i][j] = r.x; acc[i][j + 1] = r.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 1e-7f; }   // keep the loop from being hoisted
    }
    float t = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) t += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 2, iters = 20000;
    float* out; cudaMalloc(&out, grid * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (packed) k<true><<<grid, 256>>>(out, iters, 1.f); else k<false><<<grid, 256>>>(out, iters, 1.f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double flop = 2.0 * 64 * iters * (double)grid * 256;
            if (rep == 2) printf("%s: %.3f ms  %.1f TFLOP/s fp32\n", packed ? "FFMA2 (f32x2)" : "FFMA scalar", ms, flop / ms * 1e-9);
        }
    }
    return 0;
}
