// One-shot peer-memory all-reduce for wavelength sharding.
//
// The only exchange of a wavelength-sharded RT iteration is the sum over ranks of the per-interface
// flux totals: 2*ninterface doubles (~1.6 kB), pure latency.  Instead of a ring/tree collective each
// rank STOREs its partial vector straight into a slot of every peer's mailbox over NVLink (the
// mailboxes are cudaIpc-mapped into every process), raises a sequence flag, waits for the peers'
// flags and adds the world's slots in rank order.  One kernel, one NVLink round trip, and every rank
// ends up with bitwise-identical totals (fixed summation order), so the replicated temperature update
// cannot drift between ranks.  The reference has no counterpart (single GPU).
#include "common.cuh"
#include "comm.cuh"

// vec_a[0..n) and (optionally) vec_b[0..n) are reduced in one exchange; `net` (optional) receives a - b.
__global__ void __launch_bounds__(256)
k_peer_allreduce(CommPeers peers, double* __restrict__ vec_a, double* __restrict__ vec_b,
                 double* __restrict__ net, int n, int rank, int world, int slot, size_t data_bytes,
                 unsigned long long* __restrict__ seq_dev) {
    const unsigned long long seq = *seq_dev + 1ull;  // this round (one block: everybody reads before thread 0 advances it)
    const int bank = (int)(seq & 1ull);
    const int nn = vec_b ? 2 * n : n;
    // 1. scatter my partials into slot `rank` of every mailbox (my own included)
    for (int k = threadIdx.x; k < nn * world; k += blockDim.x) {
        const int r = k / nn, t = k - r * nn;
        double* data = reinterpret_cast<double*>(peers.p[r]);
        data[((size_t)bank * world + rank) * slot + t] = t < n ? vec_a[t] : vec_b[t - n];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag[rank] of every mailbox <- seq
    if ((int)threadIdx.x < world) {
        unsigned long long* flags =
            reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peers.p[threadIdx.x]) + data_bytes);
        st_flag(flags + 16 * rank, seq);
    }
    // 3. wait until every rank has published this round into MY mailbox
    if ((int)threadIdx.x < world) {
        const unsigned long long* flags =
            reinterpret_cast<const unsigned long long*>(reinterpret_cast<char*>(peers.p[rank]) + data_bytes);
        while (ld_flag(flags + 16 * threadIdx.x) < seq) {
            __nanosleep(64);
        }
    }
    __syncthreads();
    __threadfence_system();
    // 4. add the world's slots in rank order
    const volatile double* mine = reinterpret_cast<const volatile double*>(peers.p[rank]);
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < world; r++) a += mine[((size_t)bank * world + r) * slot + t];
        vec_a[t] = a;
        if (vec_b) {
            for (int r = 0; r < world; r++) b += mine[((size_t)bank * world + r) * slot + n + t];
            vec_b[t] = b;
            if (net) net[t] = a - b;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *seq_dev = seq;
}

static int comm_launch(helios_ctx* ctx, const char* who, double* a, double* b, double* net, int n) {
    helios_comm_state* c = ctx->comm;
    if (!c) {
        helios_set_error("%s: no communicator", who);
        return HELIOS_ERR_STATE;
    }
    if (a == nullptr || n <= 0 || (b ? 2 * n : n) > c->slot) {
        helios_set_error("%s: invalid argument (n = %d, slot = %d doubles)", who, n, c->slot);
        return HELIOS_ERR_ARG;
    }
    for (int r = 0; r < c->world; r++) {
        if (c->peers[r] == nullptr) {
            helios_set_error("%s: peer %d not connected", who, r);
            return HELIOS_ERR_STATE;
        }
    }
    CommPeers p;
    for (int r = 0; r < COMM_MAX_WORLD; r++) p.p[r] = c->peers[r];
    k_peer_allreduce<<<1, 256, 0, ctx->stream>>>(p, a, b, net, n, c->rank, c->world, c->slot, c->data_bytes, c->seq_dev);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_comm_fused_next(helios_ctx* ctx, int n, FusedComm* fc) {
    helios_comm_state* c = ctx->comm;
    if (!c || !c->fused) {
        fc->world = 0;
        return HELIOS_OK;
    }
    if (2 * n > c->slot) {
        helios_set_error("fused flux all-reduce: %d interfaces do not fit the mailbox slot (%d doubles)", n, c->slot);
        return HELIOS_ERR_ARG;
    }
    for (int r = 0; r < c->world; r++) {
        if (c->peers[r] == nullptr) {
            helios_set_error("fused flux all-reduce: peer %d not connected", r);
            return HELIOS_ERR_STATE;
        }
    }
    if (c->fused_ticket == nullptr) {
        HCUDA(cudaMalloc((void**)&c->fused_ticket, sizeof(unsigned)));
        HCUDA(cudaMemsetAsync(c->fused_ticket, 0, sizeof(unsigned), ctx->stream));
    }
    for (int r = 0; r < COMM_MAX_WORLD; r++) fc->peers.p[r] = c->peers[r];
    fc->rank = c->rank;
    fc->world = c->world;
    fc->slot = c->slot;
    fc->data_bytes = c->data_bytes;
    fc->ll_off = c->ll_off;
    fc->seq_dev = c->seq_dev;
    fc->ticket = c->fused_ticket;
    return HELIOS_OK;
}

extern "C" {

int helios_comm_set_fused(helios_ctx* ctx, int on) {
    HCTX(ctx);
    if (!ctx->comm) {
        helios_set_error("helios_comm_set_fused: no communicator");
        return HELIOS_ERR_STATE;
    }
    ctx->comm->fused = on != 0;
    return HELIOS_OK;
}

int helios_comm_create(helios_ctx* ctx, int rank, int world, int slot_doubles, unsigned char* handle_out) {
    HCTX(ctx);
    HARG(handle_out != nullptr && world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world &&
         slot_doubles > 0);
    static_assert(sizeof(cudaIpcMemHandle_t) == HELIOS_IPC_HANDLE_BYTES, "IPC handle size");
    if (ctx->comm) {
        helios_set_error("helios_comm_create: communicator already exists");
        return HELIOS_ERR_STATE;
    }
    helios_comm_state* c = new helios_comm_state();
    c->rank = rank;
    c->world = world;
    c->slot = slot_doubles;
    c->data_bytes = (size_t)2 * world * slot_doubles * sizeof(double);
    c->data_bytes = (c->data_bytes + 127) / 128 * 128;
    c->ll_off = c->data_bytes + (size_t)world * 128;
    const size_t total = c->ll_off + (size_t)2 * world * slot_doubles * 16;
    cudaError_t e = cudaMalloc(&c->own, total);
    if (e != cudaSuccess) {
        delete c;
        return helios_fail_cuda(e, "cudaMalloc(mailbox)", __FILE__, __LINE__);
    }
    e = cudaMemset(c->own, 0, total);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->seq_dev, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c->seq_dev, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->own);
    if (e != cudaSuccess) {
        cudaFree(c->own);
        delete c;
        return helios_fail_cuda(e, "mailbox setup", __FILE__, __LINE__);
    }
    memcpy(handle_out, &h, sizeof(h));
    c->peers[rank] = c->own;
    ctx->comm = c;
    return HELIOS_OK;
}

int helios_comm_connect(helios_ctx* ctx, const unsigned char* handles) {
    HCTX(ctx);
    HARG(handles != nullptr);
    helios_comm_state* c = ctx->comm;
    if (!c) {
        helios_set_error("helios_comm_connect: call helios_comm_create first");
        return HELIOS_ERR_STATE;
    }
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * HELIOS_IPC_HANDLE_BYTES, sizeof(h));
        HCUDA(cudaIpcOpenMemHandle(&c->peers[r], h, cudaIpcMemLazyEnablePeerAccess));
        c->opened[r] = true;
    }
    return HELIOS_OK;
}

int helios_comm_allreduce_sum(helios_ctx* ctx, double* vec, int n) {
    HCTX(ctx);
    return comm_launch(ctx, "helios_comm_allreduce_sum", vec, nullptr, nullptr, n);
}

int helios_comm_allreduce_flux_totals(helios_ctx* ctx, double* F_up_tot, double* F_down_tot, double* F_net,
                                      int numinterfaces) {
    HCTX(ctx);
    HARG(F_down_tot != nullptr && F_net != nullptr);
    return comm_launch(ctx, "helios_comm_allreduce_flux_totals", F_up_tot, F_down_tot, F_net, numinterfaces);
}

int helios_comm_destroy(helios_ctx* ctx) {
    if (ctx == nullptr || ctx->comm == nullptr) return HELIOS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    helios_comm_state* c = ctx->comm;
    for (int r = 0; r < c->world; r++)
        if (c->opened[r] && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
    if (c->own) cudaFree(c->own);
    if (c->fused_ticket) cudaFree(c->fused_ticket);
    if (c->seq_dev) cudaFree(c->seq_dev);
    delete c;
    ctx->comm = nullptr;
    return HELIOS_OK;
}

}  // extern "C"
