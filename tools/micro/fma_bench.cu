// Developer micro-benchmark: FFMA vs FFMA2 issue throughput on sm_100a (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int CH>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[CH * 2];
#pragma unroll
    for (int i = 0; i < CH * 2; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) {
                x[2 * i] = fmaf(x[2 * i], a, b);
                x[2 * i + 1] = fmaf(x[2 * i + 1], a, b);
            } else {
                unsigned long long v, aa, bb;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH * 2; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int CH>
void run(const char* name, int ctas_per_sm) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, CH><<<148 * ctas_per_sm, 256>>>(out, 100, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE, CH><<<148 * ctas_per_sm, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)148 * ctas_per_sm * 256 * iters * CH * 2;
    printf("%-28s ctas/SM=%d chains=%2d : %.3f ms  %.1f TFLOP/s  (%.1f lane-FMA/clk/SM at 1.9 GHz)\n", name, ctas_per_sm, CH * 2, ms,
           2 * fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main() {
    run<0, 4>("FFMA", 4); run<0, 8>("FFMA", 4); run<0, 8>("FFMA", 8); run<0, 4>("FFMA", 2);
    run<1, 4>("FFMA2", 4); run<1, 8>("FFMA2", 4); run<1, 8>("FFMA2", 8); run<1, 4>("FFMA2", 2);
    return 0;
}
