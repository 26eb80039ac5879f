// Shared definitions for libhelios_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "../../include/helios_b200.h"

// physical constants: the literals of the reference device code (K:36-41), so that both sides
// evaluate with identical scalars.
namespace hc {
constexpr double PI = 3.141592653589793;
constexpr double HCONST = 6.62607004e-27;
constexpr double CSPEED = 29979245800.0;
constexpr double KBOLTZMANN = 1.38064852e-16;
constexpr double STEFANBOLTZMANN = 5.6703669999999995e-5;
constexpr double AMU = 1.6605390666e-24;
}  // namespace hc

struct helios_comm_state;

// Batched atmospheres (helios_ctx_set_batch): nbatch atmospheres of identical shape go through every
// per-iteration kernel in ONE launch.  Per-atmosphere arrays hold nbatch consecutive single-atmosphere
// arrays; their strides follow from the shape by the reference's allocation sizes (Q:400-409, 411-461,
// 613-665), collected here once.
struct BatchDesc {
    bool active = false;  // batch mode on (nbatch may be 1: a single atmosphere with the on-device bookkeeping)
    int nbatch = 1;
    int nlayer = 0, nbin = 0, ny = 0;
    const int* table_index = nullptr;     // device [nbatch]: which opacity table an atmosphere reads (null: 0)
    size_t ktable_stride = 0, cross_stride = 0, mmass_stride = 0;  // doubles between consecutive tables
    const double* g = nullptr;            // device [nbatch]: surface gravity (null: the scalar argument)
    const double* planck_star = nullptr;  // device [nbatch][nbin]: stellar Planck row of each atmosphere
    int* done = nullptr;                  // device [nbatch]: converged flags, owned by the context
    int* converged_at = nullptr;          // device [nbatch]: iteration count at which done[b] latched
    int* iter_dev = nullptr;              // device [1]: iteration counter (rad_temp_iter reads it when `use_iter_dev`)
    bool use_iter_dev = false;
    unsigned* ticket = nullptr;           // device [1]: blocks finished (helios_rad_temp_iter_latched)
    __host__ __device__ int nint() const { return nlayer + 1; }
    __host__ __device__ size_t wg() const { return (size_t)(nlayer + 1) * nbin * ny; }  // every [i][x][y] array (Q:407)
    __host__ __device__ size_t band_lay() const { return (size_t)nlayer * nbin; }
    __host__ __device__ size_t band_int() const { return (size_t)(nlayer + 1) * nbin; }
    __host__ __device__ size_t planck_lay() const { return (size_t)(nlayer + 2) * nbin; }
};

struct helios_plan_info {
    const void* ptr = nullptr;
    size_t bytes = 0;
    int kind = 0;  // 1 = isothermal, 2 = non-isothermal
    int nobeam = 0, nint = 0, ncol = 0, nbatch = 0;
};

struct helios_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    size_t l2_bytes = 0;
    size_t total_mem = 0;
    unsigned long long launches = 0;
    int fband_mode = 0;
    size_t bytes_allocated = 0;
    std::unordered_map<void*, size_t> allocs;
    // small per-context scratch (reductions, flags)
    double* scratch = nullptr;
    size_t scratch_bytes = 0;
    helios_comm_state* comm = nullptr;
    BatchDesc batch;
    bool capturing = false;  // helios_graph_begin .. helios_graph_end
    std::vector<void*> deferred_free;  // buffers released while capturing (a free synchronises: illegal inside a capture)
    unsigned long long capture_launches0 = 0;
    // Direct-beam arrays known to hold only (signed) zeros: written by fdir_* with dir_beam == 0 and not
    // touched since (every write to device memory goes through this library).  fband_* then skips loading
    // them and the G+/- coefficient arrays that only multiply them.  [0] = F_dir_wg, [1] = Fc_dir_wg.
    const void* zero_beam[2] = {nullptr, nullptr};
    size_t zero_beam_bytes = 0;
    // sweep plans built by this context (fband_plan.cu): the layout of a plan depends on how it was built (beam rows
    // present or not), so the planned sweeps only accept plans recorded here and not overwritten since
    std::unordered_map<const void*, helios_plan_info> plans;
    unsigned* integ_ticket = nullptr;  // integrate_flux: blocks finished per (atmosphere, interface)
    size_t integ_ticket_n = 0;
    void* flush_buf = nullptr;  // helios_l2_flush
    size_t flush_bytes = 0;
    // pow(epsi,-2), pow(mu_star,-2) of calc_trans_*, keyed on (epsi, mu_star): trans.cu
    double trans_cache[4] = {0, 0, 0, 0};
    bool trans_cache_valid = false;
    std::mutex mu;
};

struct helios_event {
    cudaEvent_t ev = nullptr;
    int device = 0;
};

void helios_set_error(const char* fmt, ...);
int helios_fail_cuda(cudaError_t e, const char* what, const char* file, int line);
// grows ctx->scratch to at least nbytes (stream-ordered with respect to ctx->stream users)
int helios_ctx_scratch(helios_ctx* ctx, size_t nbytes, double** out);

#define HCUDA(call)                                                            \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return helios_fail_cuda(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define HARG(cond)                                                             \
    do {                                                                       \
        if (!(cond)) {                                                         \
            helios_set_error("%s: invalid argument: %s", __func__, #cond);     \
            return HELIOS_ERR_ARG;                                             \
        }                                                                      \
    } while (0)

#define HCTX(ctx)                                                              \
    do {                                                                       \
        if ((ctx) == nullptr) {                                                \
            helios_set_error("%s: null context", __func__);                    \
            return HELIOS_ERR_ARG;                                             \
        }                                                                      \
        HCUDA(cudaSetDevice((ctx)->device));                                   \
    } while (0)

// after a <<<>>> launch
#define HLAUNCHED(ctx)                                                         \
    do {                                                                       \
        (ctx)->launches++;                                                     \
        HCUDA(cudaGetLastError());                                             \
    } while (0)

// entry points that have no batched form refuse to run in batch mode instead of silently doing one atmosphere
#define HNOBATCH(ctx)                                                          \
    do {                                                                       \
        if ((ctx)->batch.active) {                                         \
            helios_set_error("%s: not available in batch mode", __func__);     \
            return HELIOS_ERR_STATE;                                           \
        }                                                                      \
    } while (0)

// batch mode: the call's dimensions must be the ones the batch was declared with
#define HBATCHDIMS(ctx, cond)                                                  \
    do {                                                                       \
        if ((ctx)->batch.active && !(cond)) {                              \
            helios_set_error("%s: dimensions differ from helios_ctx_set_batch: %s", __func__, #cond); \
            return HELIOS_ERR_ARG;                                             \
        }                                                                      \
    } while (0)

// a buffer range is about to be overwritten from outside the kernels: forget what was known about it
static inline void helios_note_write(helios_ctx* ctx, const void* p, size_t nbytes) {
    const char* lo = static_cast<const char*>(p);
    for (int k = 0; k < 2; k++) {
        const char* z = static_cast<const char*>(ctx->zero_beam[k]);
        if (z != nullptr && lo < z + ctx->zero_beam_bytes && z < lo + nbytes) ctx->zero_beam[k] = nullptr;
    }
    for (auto it = ctx->plans.begin(); it != ctx->plans.end();) {
        const char* z = static_cast<const char*>(it->second.ptr);
        if (lo < z + it->second.bytes && z < lo + nbytes) it = ctx->plans.erase(it);
        else ++it;
    }
}

static inline void helios_plan_record(helios_ctx* ctx, const void* ptr, size_t bytes, int kind, int nobeam, int nint,
                                      int ncol) {
    helios_plan_info& pi = ctx->plans[ptr];
    pi.ptr = ptr;
    pi.bytes = bytes;
    pi.kind = kind;
    pi.nobeam = nobeam;
    pi.nint = nint;
    pi.ncol = ncol;
    pi.nbatch = ctx->batch.nbatch;
}

static inline const helios_plan_info* helios_plan_lookup(const helios_ctx* ctx, const void* ptr, int kind, int nint,
                                                         int ncol) {
    auto it = ctx->plans.find(ptr);
    if (it == ctx->plans.end()) return nullptr;
    const helios_plan_info& pi = it->second;
    return (pi.kind == kind && pi.nint == nint && pi.ncol == ncol && pi.nbatch == ctx->batch.nbatch) ? &pi : nullptr;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device helpers ------------

// fitting function for the E parameter (K:109-124)
__device__ __forceinline__ double E_parameter(double w0, double g0, double i2s_transition) {
    if (w0 > i2s_transition && g0 >= 0.0) {
        return fmax(1.0, 1.225 - 0.1582 * g0 - 0.1777 * w0 - 0.07465 * (g0 * g0) + 0.2351 * w0 * g0 -
                             0.05582 * (w0 * w0));
    }
    return 1.0;
}
