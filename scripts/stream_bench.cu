// microbenchmark: how fast can the tile access pattern of k_fband_wp phase A stream from HBM?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/stream_bench scripts/stream_bench.cu ; gpurun -- "./scripts/stream_bench 1; ./scripts/stream_bench 16"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct Ptrs { const double* p[24]; };

// block: NCOLS columns x (THREADS/NCOLS) rows per round; tile = NCOLS columns x nlay layers
template <int NCOLS, int NARR, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_stream(Ptrs a, double* out, int ncol, int nlay) {
    const int c = threadIdx.x % NCOLS, r = threadIdx.x / NCOLS;
    constexpr int ROWS = THREADS / NCOLS;
    const int ntile = ncol / NCOLS;
    double acc = 0.0;
    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int col = tile * NCOLS + c;
        for (int i0 = r; i0 < nlay; i0 += ROWS * UNROLL) {
            double q[UNROLL][NARR];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int i = i0 + u * ROWS;
                if (i < nlay) {
                    const size_t e = col + (size_t)ncol * i;
#pragma unroll
                    for (int z = 0; z < NARR; z++) q[u][z] = a.p[z][e];
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int i = i0 + u * ROWS;
                if (i < nlay) {
#pragma unroll
                    for (int z = 0; z < NARR; z++) acc += q[u][z];
                }
            }
        }
    }
    if (acc == 1.2345e300) out[0] = acc;
}

template <int NCOLS, int NARR, int THREADS, int UNROLL>
void run(const char* name, Ptrs a, double* out, int ncol, int nlay, int ctas_per_sm, void* flushbuf, size_t flushbytes) {
    int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9, sum = 0; int n = 8;
    for (int it = 0; it < n + 2; it++) {
        CK(cudaMemset(flushbuf, 0, flushbytes));
        CK(cudaEventRecord(e0));
        k_stream<NCOLS, NARR, THREADS, UNROLL><<<grid, THREADS>>>(a, out, ncol, nlay);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    double bytes = (double)NARR * ncol * nlay * 8;
    printf("%-46s grid %4d  avg %7.1f us  best %7.1f us  -> %6.0f GB/s (avg)\n", name, grid, sum / n * 1e3, best * 1e3, bytes / (sum / n * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
    const int nbatch = argc > 1 ? atoi(argv[1]) : 1;
    const int ncol = 7700 * nbatch / 32 * 32, nlay = 100;   // multiple of 32 columns
    Ptrs a; double* out;
    for (int z = 0; z < 24; z++) { double* p; CK(cudaMalloc(&p, (size_t)ncol * (nlay + 1) * 8)); CK(cudaMemset(p, 0, (size_t)ncol * (nlay + 1) * 8)); a.p[z] = p; }
    CK(cudaMalloc(&out, 64));
    void* fb; size_t fbytes = 256u << 20; CK(cudaMalloc(&fb, fbytes));
    printf("ncol %d nlay %d (%d atmospheres)\n", ncol, nlay, nbatch);
    run<8, 24, 256, 1>("NCOLS 8, 24 arrays, 256 thr, 1 cell in flight", a, out, ncol, nlay, 2, fb, fbytes);
    run<8, 24, 256, 1>("same, 4 CTAs/SM", a, out, ncol, nlay, 4, fb, fbytes);
    run<8, 24, 256, 2>("NCOLS 8, 24 arrays, 2 cells in flight", a, out, ncol, nlay, 2, fb, fbytes);
    run<16, 24, 256, 1>("NCOLS 16, 24 arrays, 1 cell", a, out, ncol, nlay, 2, fb, fbytes);
    run<16, 24, 256, 2>("NCOLS 16, 24 arrays, 2 cells", a, out, ncol, nlay, 2, fb, fbytes);
    run<32, 24, 256, 1>("NCOLS 32, 24 arrays, 1 cell", a, out, ncol, nlay, 2, fb, fbytes);
    run<32, 24, 256, 2>("NCOLS 32, 24 arrays, 2 cells", a, out, ncol, nlay, 2, fb, fbytes);
    run<32, 24, 256, 2>("NCOLS 32, 24 arrays, 2 cells, 4 CTAs/SM", a, out, ncol, nlay, 4, fb, fbytes);
    run<32, 24, 512, 1>("NCOLS 32, 24 arrays, 512 thr, 2 CTAs/SM", a, out, ncol, nlay, 2, fb, fbytes);
    run<32, 24, 1024, 1>("NCOLS 32, 24 arrays, 1024 thr, 2 CTAs/SM", a, out, ncol, nlay, 2, fb, fbytes);
    run<16, 11, 256, 1>("NCOLS 16, 11 arrays, 1 cell", a, out, ncol, nlay, 2, fb, fbytes);
    run<16, 11, 256, 4>("NCOLS 16, 11 arrays, 4 cells", a, out, ncol, nlay, 2, fb, fbytes);
    run<16, 7, 256, 4>("NCOLS 16, 7 arrays, 4 cells", a, out, ncol, nlay, 2, fb, fbytes);
    return 0;
}
