// Dependent-issue latencies of the operations the sweep kernel's phase B chains (B200, one warp per block).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/lat_bench scripts/lat_bench.cu && gpurun_out/lat_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double tiny_to_abs(double f) { return fabs(f) < 1e-10 ? fabs(f) : f; }
template <int MODE>
__global__ void k(double* out, long long* cyc, double a, double c, int n) {
    double F = out[threadIdx.x], G = out[threadIdx.x + 32];
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) F = __fma_rn(a, F, c);
            if (MODE == 1) F = tiny_to_abs(__fma_rn(a, F, c));
            if (MODE == 2) F = __shfl_down_sync(0xffffffffu, F, 1, 32);
            if (MODE == 3) { F = __fma_rn(a, F, c); G = __fma_rn(a, G, c); }  // two independent chains
            if (MODE == 4) { double o = __shfl_down_sync(0xffffffffu, F, 1, 32); double p = __shfl_down_sync(0xffffffffu, G, 1, 32);
                             if ((threadIdx.x & 31) + 1 < 25) { G = __fma_rn(F, p, G); F = F * o; } }  // one scan step
            if (MODE == 5) F = F * a;
            if (MODE == 6) F = F + a;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = F + G;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 64 * 32 * 8 * 8); cudaMemset(out, 0, 64 * 32 * 8 * 8); cudaMalloc(&cyc, 8 * 1024);
    const int n = 256;
    k<MODE><<<148, 32 * warps>>>(out, cyc, 0.999, 1e-3, n);
    k<MODE><<<148, 32 * warps>>>(out, cyc, 0.999, 1e-3, n);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s warps/SM %2d: %.1f cycles per step\n", name, warps, (double)h[0] / (n * 16));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 16}) {
        if (w == 1) { run<0>("DFMA dependent", 1); run<1>("DFMA + tiny_to_abs dependent", 1); run<2>("SHFL f64 dependent", 1);
            run<3>("2 independent DFMA chains (per pair)", 1); run<4>("scan step (2 shfl f64 + DFMA + DMUL)", 1); run<5>("DMUL dependent", 1); run<6>("DADD dependent", 1); }
        if (w == 4) { run<0>("DFMA dependent", 4); run<1>("DFMA + tiny_to_abs dependent", 4); run<4>("scan step", 4); }
        if (w == 16) { run<0>("DFMA dependent", 16); run<1>("DFMA + tiny_to_abs dependent", 16); run<2>("SHFL f64 dependent", 16); run<4>("scan step", 16); }
    }
    return 0;
}
