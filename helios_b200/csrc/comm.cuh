// Peer-memory mailbox of the wavelength-sharded all-reduce (comm.cu), shared with the fused form inside
// k_band_integrate (flux.cu).
#pragma once
#include "common.cuh"

#define COMM_MAX_WORLD 16

struct helios_comm_state {
    int rank = 0, world = 1, slot = 0;
    // exchange sequence number, ON THE DEVICE: every exchanging kernel reads it at its start (round = *seq_dev + 1) and its
    // last block advances it, so that a recorded launch (CUDA graph) can be replayed -- a by-value sequence number would
    // be frozen into the graph.  All ranks make the same sequence of exchanges, so their counters stay in step.
    unsigned long long* seq_dev = nullptr;
    // own mailbox: [2 banks][world][slot] doubles, then [world] flags (u64), flags padded to 128 B
    void* own = nullptr;
    void* peers[COMM_MAX_WORLD] = {nullptr};
    bool opened[COMM_MAX_WORLD] = {false};
    void** peers_dev = nullptr;
    size_t data_bytes = 0;
    bool fused = false;            // helios_comm_set_fused: integrate_flux_double performs the exchange in its own launch
    unsigned* fused_ticket = nullptr;  // device: interfaces finished (fused form)
};

struct CommPeers {
    void* p[COMM_MAX_WORLD];
};

// what k_band_integrate needs to run the exchange in its epilogue (world == 0: off)
struct FusedComm {
    CommPeers peers;
    int rank = 0, world = 0, slot = 0;
    size_t data_bytes = 0;
    unsigned long long* seq_dev = nullptr;
    unsigned* ticket = nullptr;
};

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// fills `fc` for an exchange inside k_band_integrate; HELIOS_OK or an error status
int helios_comm_fused_next(helios_ctx* ctx, int n, FusedComm* fc);
