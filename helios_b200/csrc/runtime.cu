// Context, buffer and event management of libhelios_b200.so.
// Replaces pycuda.autoinit / gpuarray.to_gpu / cuda.mem_alloc / .get() / cuda.Event of the
// reference (C:24, Q:463-665, C:838-841).  The library owns every device allocation.
#include "common.cuh"

static thread_local char g_err[1024] = "";

void helios_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int helios_fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    helios_set_error("CUDA error %d (%s) in `%s` at %s:%d", (int)e, cudaGetErrorString(e), what, file,
                     line);
    // clear the sticky-free error state so that later calls report their own failures
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? HELIOS_ERR_NOMEM : HELIOS_ERR_CUDA;
}

// read sweep used by helios_l2_flush
__global__ void k_read_sweep(const double4* __restrict__ src, size_t n4, double* __restrict__ sink) {
    double acc = 0.0;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n4; k += (size_t)gridDim.x * blockDim.x) {
        const double4 v = src[k];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 1.2345e300) sink[0] = acc;  // never true for the zero-filled buffer; keeps the loads alive
}

extern "C" {

int helios_abi_version(void) { return HELIOS_ABI_VERSION; }

const char* helios_last_error(void) { return g_err; }

int helios_device_count(int* count) {
    HARG(count != nullptr);
    HCUDA(cudaGetDeviceCount(count));
    return HELIOS_OK;
}

int helios_ctx_create(int device, helios_ctx** out) {
    HARG(out != nullptr);
    int n = 0;
    HCUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        helios_set_error("helios_ctx_create: device %d out of range (%d visible)", device, n);
        return HELIOS_ERR_ARG;
    }
    HCUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HCUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        helios_set_error("helios_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
        return HELIOS_ERR_STATE;
    }
    helios_ctx* ctx = new helios_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    ctx->total_mem = prop.totalGlobalMem;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return helios_fail_cuda(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return HELIOS_OK;
}

int helios_comm_destroy(helios_ctx* ctx);
static void batch_release(helios_ctx* ctx);

int helios_ctx_destroy(helios_ctx* ctx) {
    if (ctx == nullptr) return HELIOS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) helios_comm_destroy(ctx);
    for (auto& kv : ctx->allocs) cudaFree(kv.first);
    ctx->allocs.clear();
    if (ctx->scratch) cudaFree(ctx->scratch);
    batch_release(ctx);
    if (ctx->integ_ticket) cudaFree(ctx->integ_ticket);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return HELIOS_OK;
}

int helios_ctx_set_stream(helios_ctx* ctx, void* cuda_stream) {
    HCTX(ctx);
    HCUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return HELIOS_OK;
}

int helios_ctx_get_stream(helios_ctx* ctx, void** cuda_stream) {
    HCTX(ctx);
    HARG(cuda_stream != nullptr);
    *cuda_stream = (void*)ctx->stream;
    return HELIOS_OK;
}

int helios_ctx_sync(helios_ctx* ctx) {
    HCTX(ctx);
    HCUDA(cudaStreamSynchronize(ctx->stream));
    return HELIOS_OK;
}

int helios_ctx_device_info(helios_ctx* ctx, int* num_sms, size_t* l2_bytes, size_t* total_mem) {
    HCTX(ctx);
    if (num_sms) *num_sms = ctx->num_sms;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    if (total_mem) *total_mem = ctx->total_mem;
    return HELIOS_OK;
}

int helios_ctx_launch_count(helios_ctx* ctx, unsigned long long* count) {
    HCTX(ctx);
    HARG(count != nullptr);
    *count = ctx->launches;
    return HELIOS_OK;
}

static void batch_release(helios_ctx* ctx) {
    if (ctx->batch.done) cudaFree(ctx->batch.done);
    ctx->batch = BatchDesc();
}

int helios_ctx_set_batch(helios_ctx* ctx, int nbatch, int nlayer, int nbin, int ny, const int* table_index,
                         size_t ktable_stride, size_t crosstable_stride, size_t meanmass_stride, const double* g,
                         const double* planck_star) {
    HCTX(ctx);
    HARG(nbatch >= 0);
    HCUDA(cudaStreamSynchronize(ctx->stream));
    batch_release(ctx);
    if (nbatch == 0) return HELIOS_OK;
    HARG(nlayer > 1 && nbin > 0 && ny > 0);
    BatchDesc b;
    b.active = true;
    b.nbatch = nbatch;
    b.nlayer = nlayer;
    b.nbin = nbin;
    b.ny = ny;
    b.table_index = table_index;
    b.ktable_stride = ktable_stride;
    b.cross_stride = crosstable_stride;
    b.mmass_stride = meanmass_stride;
    b.g = g;
    b.planck_star = planck_star;
    // done[nbatch] | converged_at[nbatch] | iteration counter | ticket
    const size_t n = 2 * (size_t)nbatch + 2;
    HCUDA(cudaMalloc((void**)&b.done, sizeof(int) * n));
    HCUDA(cudaMemset(b.done, 0, sizeof(int) * n));
    b.converged_at = b.done + nbatch;
    b.iter_dev = b.done + 2 * nbatch;
    b.ticket = reinterpret_cast<unsigned*>(b.done + 2 * nbatch + 1);
    ctx->batch = b;
    return HELIOS_OK;
}

int helios_ctx_batch_state(helios_ctx* ctx, int* done_host, int* converged_at_host, int* iter_host, int reset) {
    HCTX(ctx);
    if (!ctx->batch.active) {
        helios_set_error("helios_ctx_batch_state: not in batch mode");
        return HELIOS_ERR_STATE;
    }
    const size_t nb = (size_t)ctx->batch.nbatch;
    if (reset) HCUDA(cudaMemsetAsync(ctx->batch.done, 0, sizeof(int) * (2 * nb + 2), ctx->stream));
    if (done_host)
        HCUDA(cudaMemcpyAsync(done_host, ctx->batch.done, sizeof(int) * nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (converged_at_host)
        HCUDA(cudaMemcpyAsync(converged_at_host, ctx->batch.converged_at, sizeof(int) * nb, cudaMemcpyDeviceToHost,
                              ctx->stream));
    if (iter_host)
        HCUDA(cudaMemcpyAsync(iter_host, ctx->batch.iter_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (done_host || converged_at_host || iter_host) HCUDA(cudaStreamSynchronize(ctx->stream));
    return HELIOS_OK;
}

int helios_ctx_batch_device_iteration(helios_ctx* ctx, int enable) {
    HCTX(ctx);
    if (!ctx->batch.active) {
        helios_set_error("helios_ctx_batch_device_iteration: not in batch mode");
        return HELIOS_ERR_STATE;
    }
    ctx->batch.use_iter_dev = enable != 0;
    return HELIOS_OK;
}

// ---------------------------------------------------------------- CUDA graphs
struct helios_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int device = 0;
    unsigned long long launches = 0;  // kernel launches recorded (bench bookkeeping)
};

static void graph_release(helios_graph* g) {
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
}

// Graph objects released while some context of the process is capturing (a host garbage collector picks its own moment)
// are destroyed when that capture ends: destroying a graph inside a capture invalidates it.
static std::mutex g_graph_mu;
static int g_capturing = 0;
static std::vector<helios_graph*> g_graph_late;

static void graph_capture_mark(int delta) {
    std::vector<helios_graph*> late;
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        g_capturing += delta;
        if (g_capturing == 0) late.swap(g_graph_late);
    }
    for (helios_graph* g : late) graph_release(g);
}

int helios_graph_begin(helios_ctx* ctx) {
    HCTX(ctx);
    if (ctx->capturing) {
        helios_set_error("helios_graph_begin: already capturing");
        return HELIOS_ERR_STATE;
    }
    HCUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    graph_capture_mark(+1);
    ctx->capture_launches0 = ctx->launches;
    return HELIOS_OK;
}

int helios_graph_end(helios_ctx* ctx, helios_graph** out) {
    HCTX(ctx);
    HARG(out != nullptr);
    if (!ctx->capturing) {
        helios_set_error("helios_graph_end: no capture in progress");
        return HELIOS_ERR_STATE;
    }
    ctx->capturing = false;
    helios_graph* g = new helios_graph();
    g->device = ctx->device;
    g->launches = ctx->launches - ctx->capture_launches0;
    ctx->launches = ctx->capture_launches0;  // recorded, not executed
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
    graph_capture_mark(-1);
    {
        std::vector<void*> late;
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            late.swap(ctx->deferred_free);
        }
        if (!late.empty()) {
            cudaStreamSynchronize(ctx->stream);
            for (void* p : late) cudaFree(p);
        }
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
        return helios_fail_cuda(e, "graph capture / instantiate", __FILE__, __LINE__);
    }
    *out = g;
    return HELIOS_OK;
}

int helios_graph_launch(helios_ctx* ctx, helios_graph* g) {
    HCTX(ctx);
    HARG(g != nullptr && g->exec != nullptr);
    HCUDA(cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->launches;
    return HELIOS_OK;
}

int helios_graph_destroy(helios_graph* g) {
    if (g == nullptr) return HELIOS_OK;
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (g_capturing > 0) {
            g_graph_late.push_back(g);
            return HELIOS_OK;
        }
    }
    graph_release(g);
    return HELIOS_OK;
}

int helios_l2_flush(helios_ctx* ctx, int mode) {
    HCTX(ctx);
    HARG(mode == 0 || mode == 1);
    const size_t bytes = ctx->l2_bytes * 2 > ((size_t)256 << 20) ? ctx->l2_bytes * 2 : ((size_t)256 << 20);
    if (ctx->flush_buf == nullptr) {
        HCUDA(cudaMalloc(&ctx->flush_buf, bytes + 64));
        ctx->flush_bytes = bytes;
    }
    HCUDA(cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    if (mode == 1) {
        // the memset leaves L2 full of DIRTY lines, whose write-back would be charged to whatever runs next;
        // a read sweep over the same buffer replaces them with clean lines (cold for every other address)
        double* sink = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->flush_buf) + ctx->flush_bytes);
        k_read_sweep<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(reinterpret_cast<const double4*>(ctx->flush_buf),
                                                               ctx->flush_bytes / sizeof(double4), sink);
        HCUDA(cudaGetLastError());
    }
    return HELIOS_OK;
}

int helios_ctx_set_fband_mode(helios_ctx* ctx, int mode) {
    HCTX(ctx);
    HARG(mode >= 0 && mode <= 2);
    ctx->fband_mode = mode;
    return HELIOS_OK;
}

int helios_ctx_bytes_allocated(helios_ctx* ctx, size_t* nbytes) {
    HCTX(ctx);
    HARG(nbytes != nullptr);
    *nbytes = ctx->bytes_allocated;
    return HELIOS_OK;
}

int helios_buf_alloc(helios_ctx* ctx, size_t nbytes, void** dptr) {
    HCTX(ctx);
    HARG(dptr != nullptr);
    void* p = nullptr;
    // zero-sized arrays exist in the reference (e.g. empty entr_* tables, Q:474-475)
    size_t n = nbytes ? nbytes : 8;
    HCUDA(cudaMalloc(&p, n));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->allocs[p] = n;
    ctx->bytes_allocated += n;
    *dptr = p;
    return HELIOS_OK;
}

int helios_buf_free(helios_ctx* ctx, void* dptr) {
    HCTX(ctx);
    helios_note_write(ctx, dptr, 1);
    if (dptr == nullptr) return HELIOS_OK;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        auto it = ctx->allocs.find(dptr);
        if (it == ctx->allocs.end()) {
            helios_set_error("helios_buf_free: %p was not allocated by this context", dptr);
            return HELIOS_ERR_ARG;
        }
        ctx->bytes_allocated -= it->second;
        ctx->allocs.erase(it);
    }
    if (ctx->capturing) {
        // (a host-side garbage collector may release an array at any time) synchronising or freeing would invalidate the
        // capture: the buffer is released when the capture ends
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->deferred_free.push_back(dptr);
        return HELIOS_OK;
    }
    // kernels still in flight on the stream may use the buffer
    HCUDA(cudaStreamSynchronize(ctx->stream));
    HCUDA(cudaFree(dptr));
    return HELIOS_OK;
}

int helios_buf_h2d(helios_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes) {
    HCTX(ctx);
    if (nbytes == 0) return HELIOS_OK;
    HARG(dst_dev != nullptr && src_host != nullptr);
    helios_note_write(ctx, dst_dev, nbytes);
    HCUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    // the caller may reuse/free the (pageable) host array right away, as with gpuarray.to_gpu
    HCUDA(cudaStreamSynchronize(ctx->stream));
    return HELIOS_OK;
}

int helios_buf_d2h(helios_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes) {
    HCTX(ctx);
    if (nbytes == 0) return HELIOS_OK;
    HARG(dst_host != nullptr && src_dev != nullptr);
    HCUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    HCUDA(cudaStreamSynchronize(ctx->stream));
    return HELIOS_OK;
}

int helios_buf_h2d_async(helios_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes) {
    HCTX(ctx);
    if (nbytes == 0) return HELIOS_OK;
    HARG(dst_dev != nullptr && src_host != nullptr);
    helios_note_write(ctx, dst_dev, nbytes);
    HCUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    return HELIOS_OK;
}

int helios_buf_d2h_async(helios_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes) {
    HCTX(ctx);
    if (nbytes == 0) return HELIOS_OK;
    HARG(dst_host != nullptr && src_dev != nullptr);
    HCUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    return HELIOS_OK;
}

int helios_buf_d2d(helios_ctx* ctx, void* dst_dev, const void* src_dev, size_t nbytes) {
    HCTX(ctx);
    if (nbytes == 0) return HELIOS_OK;
    HARG(dst_dev != nullptr && src_dev != nullptr);
    helios_note_write(ctx, dst_dev, nbytes);
    HCUDA(cudaMemcpyAsync(dst_dev, src_dev, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return HELIOS_OK;
}

int helios_buf_zero(helios_ctx* ctx, void* dptr, size_t nbytes) {
    HCTX(ctx);
    if (nbytes == 0) return HELIOS_OK;
    HARG(dptr != nullptr);
    helios_note_write(ctx, dptr, nbytes);
    HCUDA(cudaMemsetAsync(dptr, 0, nbytes, ctx->stream));
    return HELIOS_OK;
}

int helios_host_alloc(size_t nbytes, void** hptr) {
    HARG(hptr != nullptr);
    HCUDA(cudaMallocHost(hptr, nbytes ? nbytes : 8));
    return HELIOS_OK;
}

int helios_host_free(void* hptr) {
    if (hptr == nullptr) return HELIOS_OK;
    HCUDA(cudaFreeHost(hptr));
    return HELIOS_OK;
}

int helios_event_create(helios_ctx* ctx, helios_event** ev) {
    HCTX(ctx);
    HARG(ev != nullptr);
    helios_event* e = new helios_event();
    e->device = ctx->device;
    cudaError_t err = cudaEventCreate(&e->ev);
    if (err != cudaSuccess) {
        delete e;
        return helios_fail_cuda(err, "cudaEventCreate", __FILE__, __LINE__);
    }
    *ev = e;
    return HELIOS_OK;
}

int helios_event_destroy(helios_event* ev) {
    if (ev == nullptr) return HELIOS_OK;
    cudaSetDevice(ev->device);
    cudaEventDestroy(ev->ev);
    delete ev;
    return HELIOS_OK;
}

int helios_event_record(helios_ctx* ctx, helios_event* ev) {
    HCTX(ctx);
    HARG(ev != nullptr);
    HCUDA(cudaEventRecord(ev->ev, ctx->stream));
    return HELIOS_OK;
}

int helios_event_synchronize(helios_event* ev) {
    HARG(ev != nullptr);
    HCUDA(cudaSetDevice(ev->device));
    HCUDA(cudaEventSynchronize(ev->ev));
    return HELIOS_OK;
}

int helios_event_elapsed_ms(helios_event* start, helios_event* stop, float* ms) {
    HARG(start != nullptr && stop != nullptr && ms != nullptr);
    HCUDA(cudaSetDevice(start->device));
    HCUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    return HELIOS_OK;
}

}  // extern "C"

int helios_ctx_scratch(helios_ctx* ctx, size_t nbytes, double** out) {
    if (nbytes > ctx->scratch_bytes) {
        if (ctx->scratch) {
            HCUDA(cudaStreamSynchronize(ctx->stream));
            HCUDA(cudaFree(ctx->scratch));
            ctx->scratch = nullptr;
            ctx->scratch_bytes = 0;
        }
        size_t n = nbytes < (1u << 20) ? (1u << 20) : nbytes;
        HCUDA(cudaMalloc((void**)&ctx->scratch, n));
        ctx->scratch_bytes = n;
    }
    *out = ctx->scratch;
    return HELIOS_OK;
}
