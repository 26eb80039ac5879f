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
    // own mailbox: [2 banks][world][slot] doubles, then [world] flags (u64), flags padded to 128 B (the stand-alone
    // kernel); then, at ll_off, [2 banks][world][slot] 16-byte packet pairs (the fused form, no flags)
    void* own = nullptr;
    size_t ll_off = 0;
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
    size_t data_bytes = 0, ll_off = 0;
    unsigned long long* seq_dev = nullptr;
    unsigned* ticket = nullptr;
};

// One double as two self-validating 8-byte packets {32 data bits, round}: the receiver needs no flag and the sender no
// fence -- a packet that carries the current round number has arrived whole (8-byte stores are not torn), whatever the
// order in which NVLink delivers the stores.  A slot is reused every second round (two banks) and starts out zero, so a
// stale packet never carries the round that is being waited for (rounds count from 1).
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned round) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(round), "r"(hi), "r"(round)
                 : "memory");
}
__device__ __forceinline__ double ll_load(const uint4* p, unsigned round) {
    unsigned lo, f0, hi, f1;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
    } while (f0 != round || f1 != round);
    return __hiloint2double((int)hi, (int)lo);
}

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
