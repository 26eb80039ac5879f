// Throughput peaks of the pipes the RT hot path leans on (B200): fp64 FMA / MUL / ADD, fp64 divide, exp, sqrt, warp
// shuffles, shared-memory loads.  Independent operations, many resident warps -- the denominators of the `bound: "fp64"`
// rooflines in bench.py (MEASURED_PEAKS.json only holds the HBM copy and the bf16 GEMM figures).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/peaks_bench scripts/peaks_bench.cu
//   gpurun_out/peaks_bench > profiles/r2_measured_fp64_peaks.json
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>

constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double a, double c, int n) {
    __shared__ double sh[256 * 4];
    double v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) v[j] = out[threadIdx.x] + j * 1e-3 + 1.0;
    sh[threadIdx.x] = v[0];
    sh[threadIdx.x + 256] = v[1];
    sh[threadIdx.x + 512] = v[2];
    sh[threadIdx.x + 768] = v[3];
    __syncthreads();
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            if (MODE == 0) v[j] = __fma_rn(a, v[j], c);
            if (MODE == 1) v[j] = __dmul_rn(a, v[j]);
            if (MODE == 2) v[j] = __dadd_rn(c, v[j]);
            if (MODE == 3) v[j] = __ddiv_rn(c, v[j] + 1.5);
            if (MODE == 4) v[j] = exp(-v[j] * 1e-3);
            if (MODE == 5) v[j] = sqrt(v[j] + 2.0);
            if (MODE == 6) v[j] = __shfl_down_sync(0xffffffffu, v[j], 1, 32);
            if (MODE == 7) {
                int w = __double2loint(v[j]);
                w = __shfl_down_sync(0xffffffffu, w, 1, 32);
                v[j] = __hiloint2double(__double2hiint(v[j]), w);
            }
            if (MODE == 8) v[j] += sh[(threadIdx.x + 32 * j + it) & 1023];
            if (MODE == 9) {  // the sweep's mix: one 64-bit shuffle per two DFMA
                if (j % 3 == 0) v[j] = __shfl_down_sync(0xffffffffu, v[j], 1, 32);
                else v[j] = __fma_rn(a, v[j], c);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int nsm, int n, double ops_per_thread_iter) {
    double* out;
    const int blocks = nsm * 8;  // 8 x 256 threads = 2048 threads per SM
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaMemset(out, 0, (size_t)blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 0.999, 1e-3, n / 8);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, 0.999, 1e-3, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    return (double)blocks * 256 * n * ops_per_thread_iter / (best * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int nsm = p.multiProcessorCount;
    const double ghz = clk * 1e-6;
    const int n = 4096;
    const double dfma = run<0>(nsm, n, ILP), dmul = run<1>(nsm, n, ILP), dadd = run<2>(nsm, n, ILP);
    const double ddiv = run<3>(nsm, n / 8, ILP), dexp = run<4>(nsm, n / 8, ILP), dsqrt = run<5>(nsm, n / 8, ILP);
    const double shfl64 = run<6>(nsm, n, ILP), shfl32 = run<7>(nsm, n, ILP), lds = run<8>(nsm, n, ILP);
    const double mix = run<9>(nsm, n, ILP);
    auto per_sm_clk = [&](double ops) { return ops / (nsm * ghz * 1e9); };
    printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"sm_clock_ghz_nominal\": %.3f,\n", p.name, nsm, ghz);
    printf(" \"how\": \"scripts/peaks_bench.cu: 2048 threads per SM, %d independent chains per thread, best of 3, CUDA events\",\n", ILP);
    printf(" \"fp64_fma_tflops\": %.2f, \"dfma_per_clk_per_sm\": %.1f,\n", 2 * dfma * 1e-12, per_sm_clk(dfma));
    printf(" \"dmul_per_clk_per_sm\": %.1f, \"dadd_per_clk_per_sm\": %.1f,\n", per_sm_clk(dmul), per_sm_clk(dadd));
    printf(" \"ddiv_rn_gops\": %.1f, \"ddiv_per_clk_per_sm\": %.2f,\n", ddiv * 1e-9, per_sm_clk(ddiv));
    printf(" \"exp_f64_gops\": %.1f, \"exp_per_clk_per_sm\": %.2f,\n", dexp * 1e-9, per_sm_clk(dexp));
    printf(" \"sqrt_f64_gops\": %.1f, \"sqrt_per_clk_per_sm\": %.2f,\n", dsqrt * 1e-9, per_sm_clk(dsqrt));
    printf(" \"shfl_f64_per_clk_per_sm\": %.1f, \"shfl_b32_per_clk_per_sm\": %.1f,\n", per_sm_clk(shfl64), per_sm_clk(shfl32));
    printf(" \"lds_f64_per_clk_per_sm\": %.1f,\n", per_sm_clk(lds));
    printf(" \"mix_2dfma_1shfl64_ops_per_clk_per_sm\": %.1f\n}\n", per_sm_clk(mix));
    return 0;
}
