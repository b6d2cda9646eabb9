// Pieces of the gradient bucket (amp_bucket.cu) that other translation units need: the peer tables, the flag protocol and
// the handle itself.  amp_disc_train.cu uses them to run the exchange inside its last kernel (finalize_exchange_kernel).
#pragma once

#include <cuda.h>

#include "amp_internal.h"

namespace amp {
namespace bucket {

constexpr int kMaxWorld = 16;
constexpr int kThreads = 256;
// Bounded spins: a rank that waits longer than this for a peer gives up (default 30 s of SM clock at ~2 GHz; NCCL would block
// and then abort).  A timeout is LOUD: the failing rank poisons the start of the requested range of its own bucket with NaN
// (so the optimiser step that follows cannot silently apply un-reduced gradients), raises a bit in a host-visible status word
// that the next amp_bucket_allreduce_mean call on that handle reports as an error, and still walks the arrival counter so the
// handle stays consistent.  AMP_B200_BUCKET_TIMEOUT_MS (read once in amp_bucket_create) overrides the limit.
constexpr long long kDefaultSpinLimitCycles = 60000000000LL;

struct Control {            // device words behind the flag array
    unsigned int arrivals;  // CTAs of the running call that have finished their slice
    uint32_t status;        // sticky failure bits (1: barrier A timed out, 2: barrier B timed out, 4: a bulk load never landed)
    uint32_t epoch;         // last epoch used; bumped by 2 by the last CTA of every call (device state: graph replays stay in step)
    unsigned int pushed;    // finalize_exchange_kernel (amp_disc_train.cu): CTAs that have pushed their share into the peers' staging
};

struct Peers {
    float *data[kMaxWorld];
    uint32_t *flags[kMaxWorld];  // flags[p] = rank p's flag array (kMaxWorld words: word q is written by rank q)
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Data moves with ordinary (weak) accesses that bypass L1 -- .cg loads are served by the owning GPU's L2, so a line of peer
// memory is never read from a stale local L1 -- and is ordered against the flags by __threadfence_system() + the
// release / acquire pair of the barriers.  (sys-scoped relaxed accesses for the payload ran at 210 GB/s per direction.)
__device__ __forceinline__ float4 ld_peer(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float4 *p, float4 v) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// wait until every rank's word in the LOCAL flag array has reached `epoch`; returns false on timeout
__device__ __forceinline__ bool wait_all(const uint32_t *local_flags, int world, uint32_t epoch, long long spin_limit) {
    const long long t0 = clock64();
    for (int p = 0; p < world; ++p) {
        while ((int32_t)(ld_acquire_sys(local_flags + p) - epoch) < 0) {
            if (clock64() - t0 > spin_limit) return false;
            __nanosleep(64);
        }
    }
    return true;
}

// failure path shared by both kernels: sticky device bit + host-visible word + NaN poison of the local range
__device__ __forceinline__ void report_failure(Control *ctl, volatile uint32_t *host_status, uint32_t bit, float *local, long long offset,
                                               long long count) {
    atomicOr(&ctl->status, bit);
    if (host_status) host_status[0] = bit | 0x80000000u;  // plain store into mapped pinned memory; the host ORs what it sees
    const float nan = __int_as_float(0x7fc00000);
    for (long long i = 0; i < min(count, 4LL); ++i) local[offset + i] = nan;
    __threadfence_system();
}

}  // namespace bucket
}  // namespace amp

struct amp_bucket {
    int world, rank, device;
    int64_t floats;
    float *data;          // local bucket (cudaMalloc, IPC-exported): `floats` of gradients, then `stage_floats` of staging that the
                          // peers' finalize_exchange_kernel pushes their partial gradients into (rank r's copy of the local slice)
    int64_t stage_floats;
    uint32_t *flags;      // local flag array [kMaxWorld] (IPC-exported) ... + arrivals + status words behind it
    unsigned long long *timing;  // globaltimer stamps of the last call: start, barrier A passed, slice published, barrier B passed
    amp::bucket::Peers peers;
    void *opened[2 * amp::bucket::kMaxWorld];
    int n_opened;
    int64_t user_floats;  // size the caller asked for (floats is rounded up to a whole quad)
    long long spin_limit; // SM cycles a rank waits for its peers before it gives up (loudly)
    volatile uint32_t *host_status;  // mapped pinned word the kernels write on failure; read without a sync by the next call
    uint32_t *host_status_dev;       // its device alias
    bool bulk;            // AMP_B200_BUCKET_BULK, read once at create
    bool connected;
    // ---- shared form (amp_bucket_create_shared): the data is a VMM allocation bound to an NVSwitch multicast object ----
    bool vmm;                               // data / peers.data[] are cuMemMap mappings, not cudaMalloc / legacy IPC
    size_t map_bytes;                       // size of every mapping (data + staging, rounded to the granularities)
    CUmemGenericAllocationHandle mem;       // local physical allocation
    CUmemGenericAllocationHandle peer_mem[amp::bucket::kMaxWorld];  // imported peers' allocations (0: none)
    CUmemGenericAllocationHandle mc;        // the multicast object (created by rank 0, imported elsewhere); 0: none
    bool joined;                            // this rank's device has been added to the multicast team
    bool mc_bound;
    bool in_switch;                         // AMP_B200_BUCKET_IN_SWITCH (default on), read once at create
    float *mc_data;                         // multicast mapping of all ranks' data: multimem.ld_reduce / multimem.st address
};
