// The one exchange step of the AMP path: the gradient all-reduce.
//
//   amp_bucket_*  <- skrl Model.reduce_parameters (upstream skrl >= 1.4.3; enabled by the reference at train.py:53-58,
//                    184-196): all-reduce(SUM) of the flattened gradients, divided by the world size, per mini-batch.
//
// One process per GPU.  Each rank owns a "bucket": ONE flat fp32 device buffer that the gradient producers write into
// directly (amp_disc_train_step's outputs are views of it), exported to the other ranks of the node with CUDA IPC.  The
// all-reduce is ONE kernel per rank working on peer memory over NVLink / NVSwitch, two-shot and in place:
//
//   barrier A   every rank tells every peer "my bucket holds this step's gradients" (a flag written into the peer's memory)
//   reduce      rank r owns slice r of the bucket: it loads slice r of EVERY rank (peer loads, fixed rank order, so the sum
//               is deterministic and identical everywhere), scales by 1/world, and stores the result into slice r of
//               EVERY rank (peer stores)
//   barrier B   "my stores are out" -- when a rank has seen that flag from all peers its whole bucket is final
//
// so each rank moves 2 (W-1)/W of the bucket over NVLink, all links busy in both directions at once, with no staging
// copy and no second launch.  Two kernels implement the middle step: allreduce_mean_bulk_kernel (default, <= 8 ranks) moves
// the payload with the bulk async-copy engine, allreduce_mean_kernel with per-thread 16-byte peer loads and stores.  Flags are monotonically increasing epochs (two per call, kept in device memory so a
// captured launch replays correctly); spins are bounded and a timeout is loud (see kDefaultSpinLimitCycles).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>

#include "amp_bucket.cuh"

namespace amp {
namespace bucket {

// count % 4 == 0, base pointers 16-byte aligned.  Slice r = quads [r * per, min((r + 1) * per, quads)).
template <int MAXW, int U>
__global__ void __launch_bounds__(kThreads) allreduce_mean_kernel(Peers peers, int rank, int world, long long offset, long long count,
                                                                  Control *ctl, volatile uint32_t *host_status, long long spin_limit,
                                                                  unsigned long long *timing) {
    __shared__ bool ok;
    __shared__ uint32_t s_epoch;
    unsigned long long t_start = 0, t_a = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    uint32_t *local_flags = peers.flags[rank];
    // the epoch lives in device memory and is advanced by the last CTA of every call (after ALL CTAs have read it: they pass
    // the arrival counter first), so a CUDA-graph replay of this launch uses fresh epochs like an eager call does
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t *>(&ctl->epoch) + 1;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    // ---- barrier A: announce (block 0) and wait (every block, on the local flag array) ----
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x == 0) {
        ok = wait_all(local_flags, world, epoch, spin_limit);
        if (!ok && blockIdx.x == 0) report_failure(ctl, host_status, 1u, peers.data[rank], offset, count);
    }
    __syncthreads();
    if (ok && blockIdx.x == 0 && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_a));
        timing[0] = t_start;
        timing[1] = t_a;
    }
    // ---- reduce slice `rank` and publish it to everyone ----
    const long long quads = count / 4;
    const long long per = (quads + world - 1) / world;
    const long long q0 = (long long)rank * per, q1 = min(quads, q0 + per);
    const float inv = 1.0f / (float)world;
    // NVLink round trips (~2 us) are the cost here, not bandwidth, so what matters is bytes in flight: every thread issues
    // the loads of U quads from ALL peers (fully unrolled over up to MAXW ranks) before the first add.
    const long long stride = (long long)gridDim.x * blockDim.x;
    // a rank whose barrier A timed out skips the exchange (its bucket is poisoned) but still walks barrier B below
    for (long long q = q0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; ok && q < q1; q += U * stride) {
        float4 v[MAXW][U];
#pragma unroll
        for (int p = 0; p < MAXW; ++p) {
            if (p < world) {
                const float4 *src = reinterpret_cast<const float4 *>(peers.data[p] + offset);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (q + u * stride < q1) v[p][u] = ld_peer(src + q + u * stride);
            }
        }
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < MAXW; ++p) {  // fixed rank order: the sum is the same on every rank
            if (p < world) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (q + u * stride < q1) { acc[u].x += v[p][u].x; acc[u].y += v[p][u].y; acc[u].z += v[p][u].z; acc[u].w += v[p][u].w; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc[u].x *= inv; acc[u].y *= inv; acc[u].z *= inv; acc[u].w *= inv; }
#pragma unroll
        for (int p = 0; p < MAXW; ++p) {
            if (p < world) {
                float4 *dst = reinterpret_cast<float4 *>(peers.data[p] + offset);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (q + u * stride < q1) st_peer(dst + q + u * stride, acc[u]);
            }
        }
    }
    // ---- barrier B: the last block of this rank to finish announces; it also waits, so the kernel (and with it the stream)
    // completes only when every peer's stores into the local bucket have been announced ----
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(&ctl->arrivals, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        ctl->arrivals = 0;       // ready for the next call (stream-ordered)
        ctl->epoch = epoch + 1;  // two epochs per call
    }
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, epoch + 1);
    }
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        timing[2] = t;  // this rank's reduce + publish is complete
        if (ok && !wait_all(local_flags, world, epoch + 1, spin_limit)) report_failure(ctl, host_status, 2u, peers.data[rank], offset, count);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        timing[3] = t;
    }
}

// ---- bulk-copy variant --------------------------------------------------------------------------------------------------------
// Same two-shot algorithm, but the payload moves with the bulk async-copy engine (cp.async.bulk, the non-tensor TMA path):
// per 4 KB chunk of its slice a CTA pulls the chunk of EVERY rank into shared memory (W bulk loads on one mbarrier), its 256
// threads add the W copies (fixed rank order) and scale, and W bulk stores push the result back to every rank.  Per-thread
// 16-byte peer loads topped out near 300-350 GB/s per direction on NVLink; bulk requests keep far more bytes in flight per
// issued instruction.  kBulkStages chunks are in flight per CTA.
constexpr int kChunkBytes = 4096;
constexpr int kChunkQuads = kChunkBytes / 16;
constexpr int kBulkStages = 4;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, long long spin_limit) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > spin_limit) return false;
    }
    return true;
}

template <int MAXW>
__global__ void __launch_bounds__(kThreads) allreduce_mean_bulk_kernel(Peers peers, int rank, int world, long long offset, long long count,
                                                                       Control *ctl, volatile uint32_t *host_status, long long spin_limit,
                                                                       unsigned long long *timing) {
    extern __shared__ __align__(128) unsigned char bulk_smem[];  // [stage][MAXW + 1][kChunkBytes]: W inputs + 1 output per stage
    __shared__ __align__(8) unsigned long long full_bar[kBulkStages];
    __shared__ bool ok;
    __shared__ uint32_t s_epoch;
    unsigned long long t_start = 0, t_a = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    uint32_t *local_flags = peers.flags[rank];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kBulkStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full_bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_epoch = *reinterpret_cast<volatile uint32_t *>(&ctl->epoch) + 1;  // device-resident epoch, see allreduce_mean_kernel
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    // ---- barrier A ----
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x == 0) {
        ok = wait_all(local_flags, world, epoch, spin_limit);
        if (!ok && blockIdx.x == 0) report_failure(ctl, host_status, 1u, peers.data[rank], offset, count);
    }
    __syncthreads();
    if (ok && blockIdx.x == 0 && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_a));
        timing[0] = t_start;
        timing[1] = t_a;
    }
    // ---- slice `rank`, chunk by chunk ----
    const long long quads = count / 4;
    const long long per = (quads + world - 1) / world;
    const long long q0 = (long long)rank * per, q1 = min(quads, q0 + per);
    // a rank whose barrier A timed out skips the exchange (its bucket is poisoned) but still walks barrier B below
    const long long n_chunks = (ok && q1 > q0) ? (q1 - q0 + kChunkQuads - 1) / kChunkQuads : 0;
    const float inv = 1.0f / (float)world;
    auto stage_base = [&](int st) { return bulk_smem + (size_t)st * (MAXW + 1) * kChunkBytes; };
    auto chunk_bytes = [&](long long c) { return (uint32_t)(min((long long)kChunkQuads, q1 - (q0 + c * kChunkQuads)) * 16); };
    auto issue_loads = [&](long long c, int st) {  // thread 0 only
        const uint32_t bytes = chunk_bytes(c);
        const uint32_t bar = smem_addr(&full_bar[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * (uint32_t)world) : "memory");
        for (int p = 0; p < world; ++p) {
            const float *src = peers.data[p] + offset + (q0 + c * kChunkQuads) * 4;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(stage_base(st) + (size_t)p * kChunkBytes)),
                         "l"(src), "r"(bytes), "r"(bar)
                         : "memory");
        }
    };
    bool good = true;
    long long k = 0;  // this CTA's k-th chunk = chunk blockIdx.x + k * gridDim.x; it lives in stage k % kBulkStages
    if (threadIdx.x == 0) {
        for (int d = 0; d < kBulkStages - 1; ++d) {  // prologue: kBulkStages - 1 chunks in flight
            const long long c = blockIdx.x + (long long)d * gridDim.x;
            if (c < n_chunks) issue_loads(c, d);
        }
    }
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++k) {
        const int st = (int)(k % kBulkStages);
        const long long ahead = c + (long long)(kBulkStages - 1) * gridDim.x;
        if (threadIdx.x == 0) {
            // this trip writes out[st], last read by the bulk stores of trip k - kBulkStages: at most the kBulkStages - 1
            // most recent store groups may still be reading.  Then prefetch chunk k + kBulkStages - 1 into the stage whose
            // inputs were consumed one trip ago.
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBulkStages - 1) : "memory");
            if (ahead < n_chunks) issue_loads(ahead, (int)((k + kBulkStages - 1) % kBulkStages));
        }
        if (!mbar_wait_bounded(smem_addr(&full_bar[st]), (uint32_t)((k / kBulkStages) & 1), spin_limit)) good = false;
        const uint32_t bytes = chunk_bytes(c);
        const float4 *in = reinterpret_cast<const float4 *>(stage_base(st));
        float4 *out = reinterpret_cast<float4 *>(stage_base(st) + (size_t)MAXW * kChunkBytes);
        if (threadIdx.x * 16u < bytes) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < MAXW; ++p) {
                if (p < world) {
                    const float4 v = in[p * kChunkQuads + threadIdx.x];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
            out[threadIdx.x] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of `out` -> visible to the bulk engine
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int p = 0; p < world; ++p) {
                float *dst = peers.data[p] + offset + (q0 + c * kChunkQuads) * 4;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(out)), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every store of this CTA has completed
    if (!good && threadIdx.x == 0) report_failure(ctl, host_status, 4u, peers.data[rank], offset, count);
    // ---- barrier B ----
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(&ctl->arrivals, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        ctl->arrivals = 0;
        ctl->epoch = epoch + 1;
    }
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, epoch + 1);
    }
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        timing[2] = t;
        if (ok && !wait_all(local_flags, world, epoch + 1, spin_limit)) report_failure(ctl, host_status, 2u, peers.data[rank], offset, count);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        timing[3] = t;
    }
}

// ---- in-switch variant (shared buckets) -------------------------------------------------------------------------------------
// The data of every rank is bound to ONE NVSwitch multicast object; `mc` is its mapping on this device.  A multimem.ld_reduce
// on mc + i makes the switch fetch element i from every rank and return the sum; a multimem.st on mc + i writes element i on
// every rank.  Rank r does that for slice r, between the same two flag barriers as above: each GPU sends its bucket once and
// receives it once (about half the two-shot kernel's bytes), and nothing is staged.  The sum is the switch's -- the same value
// on every rank, but not the rank-ordered fp32 sum of the kernels above.
template <int U>
__global__ void __launch_bounds__(kThreads) allreduce_mean_switch_kernel(Peers peers, float *mc, int rank, int world, long long offset, long long count,
                                                                         Control *ctl, volatile uint32_t *host_status, long long spin_limit,
                                                                         unsigned long long *timing) {
    __shared__ bool ok, last;
    __shared__ uint32_t s_epoch;
    unsigned long long t_start = 0, t_a = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    uint32_t *local_flags = peers.flags[rank];
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t *>(&ctl->epoch) + 1;  // device-resident epoch, see allreduce_mean_kernel
    __syncthreads();
    const uint32_t epoch = s_epoch;
    // ---- barrier A ----
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x == 0) {
        ok = wait_all(local_flags, world, epoch, spin_limit);
        if (!ok && blockIdx.x == 0) report_failure(ctl, host_status, 1u, peers.data[rank], offset, count);
    }
    __syncthreads();
    if (ok && blockIdx.x == 0 && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_a));
        timing[0] = t_start;
        timing[1] = t_a;
    }
    // ---- slice `rank`: reduce in the switch, broadcast through the switch ----
    const long long quads = count / 4;
    const long long per = (quads + world - 1) / world;
    const long long q0 = (long long)rank * per, q1 = min(quads, q0 + per);
    const float inv = 1.0f / (float)world;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float *base = mc + offset;
    for (long long q = q0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; ok && q < q1; q += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (q + u * stride < q1)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "l"(base + 4 * (q + u * stride))
                             : "memory");
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (q + u * stride < q1)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * (q + u * stride)), "f"(v[u].x * inv),
                             "f"(v[u].y * inv), "f"(v[u].z * inv), "f"(v[u].w * inv)
                             : "memory");
    }
    // ---- barrier B: one system fence per CTA, by the thread that signals (the CTA barrier orders the others' stores before it) ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(&ctl->arrivals, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        ctl->arrivals = 0;
        ctl->epoch = epoch + 1;
    }
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, epoch + 1);
    }
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        timing[2] = t;
        if (ok && !wait_all(local_flags, world, epoch + 1, spin_limit)) report_failure(ctl, host_status, 2u, peers.data[rank], offset, count);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        timing[3] = t;
    }
}

// ---- driver entry points of the virtual-memory-management / multicast API, resolved through the runtime so that the library
// keeps loading on hosts without libcuda.so ----
struct DriverApi {
    decltype(&cuDeviceGet) DeviceGet;
    decltype(&cuDeviceGetAttribute) DeviceGetAttribute;
    decltype(&cuMemCreate) MemCreate;
    decltype(&cuMemRelease) MemRelease;
    decltype(&cuMemAddressReserve) MemAddressReserve;
    decltype(&cuMemAddressFree) MemAddressFree;
    decltype(&cuMemMap) MemMap;
    decltype(&cuMemUnmap) MemUnmap;
    decltype(&cuMemSetAccess) MemSetAccess;
    decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity;
    decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle;
    decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle;
    decltype(&cuMulticastCreate) MulticastCreate;
    decltype(&cuMulticastAddDevice) MulticastAddDevice;
    decltype(&cuMulticastBindMem) MulticastBindMem;
    decltype(&cuMulticastGetGranularity) MulticastGetGranularity;
    decltype(&cuGetErrorString) GetErrorString;
    bool ok;
};

static const DriverApi &driver() {
    static DriverApi api = [] {
        DriverApi a{};
        bool all = true;
        auto get = [&](const char *name, auto &fn) {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && p)
                fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
            else
                all = false;
        };
        get("cuDeviceGet", a.DeviceGet);
        get("cuDeviceGetAttribute", a.DeviceGetAttribute);
        get("cuMemCreate", a.MemCreate);
        get("cuMemRelease", a.MemRelease);
        get("cuMemAddressReserve", a.MemAddressReserve);
        get("cuMemAddressFree", a.MemAddressFree);
        get("cuMemMap", a.MemMap);
        get("cuMemUnmap", a.MemUnmap);
        get("cuMemSetAccess", a.MemSetAccess);
        get("cuMemGetAllocationGranularity", a.MemGetAllocationGranularity);
        get("cuMemExportToShareableHandle", a.MemExportToShareableHandle);
        get("cuMemImportFromShareableHandle", a.MemImportFromShareableHandle);
        get("cuMulticastCreate", a.MulticastCreate);
        get("cuMulticastAddDevice", a.MulticastAddDevice);
        get("cuMulticastBindMem", a.MulticastBindMem);
        get("cuMulticastGetGranularity", a.MulticastGetGranularity);
        get("cuGetErrorString", a.GetErrorString);
        a.ok = all;
        return a;
    }();
    return api;
}

static int cu_fail(CUresult r, const char *what) {
    const char *msg = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &msg);
    return fail(r == CUDA_ERROR_NOT_SUPPORTED || r == CUDA_ERROR_NOT_PERMITTED ? AMP_ENOTSUP : AMP_ECUDA, "%s failed: %s (CUresult %d)", what,
                msg ? msg : "?", (int)r);
}
#define AMP_CU_TRY(expr)                                    \
    do {                                                    \
        CUresult _r = (expr);                               \
        if (_r != CUDA_SUCCESS) return cu_fail(_r, #expr);  \
    } while (0)

static CUmemAllocationProp shared_alloc_prop(int device) {
    CUmemAllocationProp ap{};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return ap;
}

// reserve an address range on the current device and map `handle` (a local / imported allocation or the multicast object) there
static int map_handle(CUmemGenericAllocationHandle handle, size_t bytes, size_t align, int device, float **out) {
    const DriverApi &cu = driver();
    CUdeviceptr va = 0;
    AMP_CU_TRY(cu.MemAddressReserve(&va, bytes, align, 0, 0));
    CUresult r = cu.MemMap(va, bytes, 0, handle, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = cu.MemSetAccess(va, bytes, &acc, 1);
        if (r != CUDA_SUCCESS) cu.MemUnmap(va, bytes);
    }
    if (r != CUDA_SUCCESS) {
        cu.MemAddressFree(va, bytes);
        return cu_fail(r, "cuMemMap / cuMemSetAccess");
    }
    *out = reinterpret_cast<float *>(va);
    return AMP_OK;
}

static void unmap(float *p, size_t bytes) {
    if (!p) return;
    driver().MemUnmap(reinterpret_cast<CUdeviceptr>(p), bytes);
    driver().MemAddressFree(reinterpret_cast<CUdeviceptr>(p), bytes);
}

}  // namespace bucket
}  // namespace amp


using namespace amp;
using namespace amp::bucket;

extern "C" {

int amp_bucket_destroy(amp_bucket_t *b) {
    if (!b) return AMP_OK;
    for (int i = 0; i < b->n_opened; ++i)
        if (b->opened[i]) cudaIpcCloseMemHandle(b->opened[i]);
    if (b->vmm) {
        cudaDeviceSynchronize();  // nothing may still be running on mappings that are about to go away
        unmap(b->mc_data, b->map_bytes);
        for (int p = 0; p < b->world; ++p) {
            if (p != b->rank) unmap(b->peers.data[p], b->map_bytes);
            if (b->peer_mem[p]) driver().MemRelease(b->peer_mem[p]);
        }
        unmap(b->data, b->map_bytes);
        if (b->mc) driver().MemRelease(b->mc);  // the team's memory unbinds when the last reference to the object goes
        if (b->mem) driver().MemRelease(b->mem);
        b->data = nullptr;
    }
    if (b->data) cudaFree(b->data);
    if (b->flags) cudaFree(b->flags);
    if (b->timing) cudaFree(b->timing);
    if (b->host_status) cudaFreeHost(const_cast<uint32_t *>(b->host_status));
    delete b;
    return AMP_OK;
}

static int create_impl(int64_t floats, int32_t world, int32_t rank, bool shared, amp_bucket_t **out) {
    AMP_REQUIRE(out, "amp_bucket_create: NULL out");
    *out = nullptr;
    AMP_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "amp_bucket_create: bad world %d / rank %d (max %d ranks)",
                world, rank, kMaxWorld);
    AMP_REQUIRE(floats >= 1, "amp_bucket_create: empty bucket");
    amp_bucket *b = new (std::nothrow) amp_bucket();
    if (!b) return fail(AMP_ENOMEM, "amp_bucket_create: host allocation failed");
    std::memset(b, 0, sizeof(*b));
    b->world = world;
    b->rank = rank;
    b->floats = (floats + 3) / 4 * 4;
    b->user_floats = floats;
    b->spin_limit = kDefaultSpinLimitCycles;
    if (const char *ms = getenv("AMP_B200_BUCKET_TIMEOUT_MS")) b->spin_limit = std::max(1LL, atoll(ms)) * 2000000LL;  // ~2 GHz
    const char *bulk_env = getenv("AMP_B200_BUCKET_BULK");  // 0 selects the load/store kernel; read once per handle
    b->bulk = !(bulk_env && bulk_env[0] == '0') && world <= 8;
    const char *switch_env = getenv("AMP_B200_BUCKET_IN_SWITCH");  // 0: even a shared bucket keeps the peer-memory kernels; read once per handle
    b->in_switch = !(switch_env && switch_env[0] == '0');
    AMP_CUDA_TRY(cudaGetDevice(&b->device));
    // staging behind the bucket: `world` copies of the largest slice any range of the bucket can have (a whole quad more per
    // rank for the rounding of the slice length)
    b->stage_floats = world > 1 ? b->floats + 4 * (int64_t)world : 0;
    cudaError_t e = cudaSuccess;
    if (shared && world > 1) {
        // a virtual-memory-management allocation that can travel as a POSIX fd and be bound to a multicast object
        const DriverApi &cu = driver();
        auto bail = [&](int rc) {
            amp_bucket_destroy(b);
            return rc;
        };
        if (!cu.ok) return bail(fail(AMP_ENOTSUP, "amp_bucket_create_shared: this driver has no virtual-memory-management / multicast entry points"));
        CUdevice dev;
        int mc_ok = 0, fd_ok = 0;
        if (cu.DeviceGet(&dev, b->device) != CUDA_SUCCESS || cu.DeviceGetAttribute(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS ||
            cu.DeviceGetAttribute(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) != CUDA_SUCCESS || !mc_ok || !fd_ok)
            return bail(fail(AMP_ENOTSUP, "amp_bucket_create_shared: device %d has no NVSwitch multicast (%d) or no POSIX-fd memory handles (%d)", b->device,
                             mc_ok, fd_ok));
        const CUmemAllocationProp ap = shared_alloc_prop(b->device);
        CUmulticastObjectProp mp{};
        mp.numDevices = (unsigned)world;
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        mp.size = (size_t)(b->floats + b->stage_floats) * 4;
        size_t g_mem = 0, g_mc = 0;
        CUresult r = cu.MemGetAllocationGranularity(&g_mem, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
        if (r == CUDA_SUCCESS) r = cu.MulticastGetGranularity(&g_mc, &mp, CU_MULTICAST_GRANULARITY_MINIMUM);
        if (r != CUDA_SUCCESS) return bail(cu_fail(r, "cuMemGetAllocationGranularity / cuMulticastGetGranularity"));
        const size_t g = std::max(g_mem, g_mc);  // both are powers of two
        b->map_bytes = ((size_t)(b->floats + b->stage_floats) * 4 + g - 1) / g * g;
        b->vmm = true;
        r = cu.MemCreate(&b->mem, b->map_bytes, &ap, 0);
        if (r != CUDA_SUCCESS) return bail(cu_fail(r, "cuMemCreate"));
        if (int rc = map_handle(b->mem, b->map_bytes, g, b->device, &b->data)) return bail(rc);
        e = cudaMemset(b->data, 0, b->map_bytes);
    } else {
        e = cudaMalloc((void **)&b->data, (size_t)(b->floats + b->stage_floats) * 4);
        if (e == cudaSuccess) e = cudaMemset(b->data, 0, (size_t)(b->floats + b->stage_floats) * 4);
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&b->flags, kMaxWorld * sizeof(uint32_t) + sizeof(Control));
    if (e == cudaSuccess) e = cudaMemset(b->flags, 0, kMaxWorld * sizeof(uint32_t) + sizeof(Control));
    if (e == cudaSuccess) {
        void *hs = nullptr;
        e = cudaHostAlloc(&hs, sizeof(uint32_t), cudaHostAllocMapped);
        if (e == cudaSuccess) {
            b->host_status = static_cast<volatile uint32_t *>(hs);
            *b->host_status = 0;
            e = cudaHostGetDevicePointer((void **)&b->host_status_dev, hs, 0);
        }
    }
    if (e == cudaSuccess && world <= 8) {  // opt-in shared memory of the bulk kernels: once per handle, not per call
        e = cudaFuncSetAttribute(allreduce_mean_bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkStages * 3 * kChunkBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(allreduce_mean_bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkStages * 5 * kChunkBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(allreduce_mean_bulk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkStages * 9 * kChunkBytes);
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&b->timing, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(b->timing, 0, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        amp_bucket_destroy(b);
        return cuda_fail(e, "cudaMalloc(amp_bucket_create)");
    }
    if (world == 1) {  // nothing to exchange: the "all-reduce" is the identity
        b->peers.data[0] = b->data;
        b->peers.flags[0] = b->flags;
        b->connected = true;
    }
    *out = b;
    return AMP_OK;
}

int amp_bucket_create(int64_t floats, int32_t world, int32_t rank, amp_bucket_t **out) { return create_impl(floats, world, rank, false, out); }

int amp_bucket_create_shared(int64_t floats, int32_t world, int32_t rank, amp_bucket_t **out) { return create_impl(floats, world, rank, true, out); }

int amp_bucket_in_switch(const amp_bucket_t *b) { return b && b->mc_data && b->connected && b->in_switch ? 1 : 0; }

int amp_bucket_export_shared(amp_bucket_t *b, void *flags_handle64, int32_t *data_fd, int32_t *multicast_fd) {
    AMP_REQUIRE(b && flags_handle64 && data_fd && multicast_fd, "amp_bucket_export_shared: NULL argument");
    AMP_REQUIRE(b->vmm, "amp_bucket_export_shared: the bucket was not made by amp_bucket_create_shared (or has one rank)");
    const DriverApi &cu = driver();
    cudaIpcMemHandle_t h;
    AMP_CUDA_TRY(cudaIpcGetMemHandle(&h, b->flags));
    std::memcpy(flags_handle64, &h, 64);
    int fd = -1;
    AMP_CU_TRY(cu.MemExportToShareableHandle(&fd, b->mem, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *data_fd = fd;
    *multicast_fd = -1;
    if (b->rank == 0) {  // one rank creates the multicast object; the others import it
        CUmulticastObjectProp mp{};
        mp.numDevices = (unsigned)b->world;
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        mp.size = b->map_bytes;
        if (!b->mc) AMP_CU_TRY(cu.MulticastCreate(&b->mc, &mp));
        AMP_CU_TRY(cu.MemExportToShareableHandle(&fd, b->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        *multicast_fd = fd;
    }
    return AMP_OK;
}

int amp_bucket_join_shared(amp_bucket_t *b, int32_t multicast_fd) {
    AMP_REQUIRE(b, "amp_bucket_join_shared: NULL handle");
    AMP_REQUIRE(b->vmm, "amp_bucket_join_shared: the bucket was not made by amp_bucket_create_shared (or has one rank)");
    AMP_REQUIRE(!b->joined, "amp_bucket_join_shared: already joined");
    const DriverApi &cu = driver();
    if (b->rank != 0) {
        AMP_REQUIRE(multicast_fd >= 0, "amp_bucket_join_shared: no file descriptor for the multicast object");
        AMP_CU_TRY(cu.MemImportFromShareableHandle(&b->mc, (void *)(uintptr_t)multicast_fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    }
    AMP_REQUIRE(b->mc, "amp_bucket_join_shared: rank 0 has to call amp_bucket_export_shared first (it creates the multicast object)");
    CUdevice dev;
    AMP_CU_TRY(cu.DeviceGet(&dev, b->device));
    AMP_CU_TRY(cu.MulticastAddDevice(b->mc, dev));
    b->joined = true;
    return AMP_OK;
}

int amp_bucket_connect_shared(amp_bucket_t *b, const void *all_flags_handles, const int32_t *data_fds) {
    AMP_REQUIRE(b && all_flags_handles && data_fds, "amp_bucket_connect_shared: NULL argument");
    AMP_REQUIRE(b->vmm, "amp_bucket_connect_shared: the bucket was not made by amp_bucket_create_shared (or has one rank)");
    AMP_REQUIRE(b->joined, "amp_bucket_connect_shared: amp_bucket_join_shared has not been called");
    AMP_REQUIRE(!b->connected, "amp_bucket_connect_shared: already connected");
    const DriverApi &cu = driver();
    const cudaIpcMemHandle_t *fh = static_cast<const cudaIpcMemHandle_t *>(all_flags_handles);
    for (int p = 0; p < b->world; ++p) {
        if (p == b->rank) {
            b->peers.data[p] = b->data;
            b->peers.flags[p] = b->flags;
            continue;
        }
        AMP_REQUIRE(data_fds[p] >= 0, "amp_bucket_connect_shared: no file descriptor for rank %d", p);
        AMP_CU_TRY(cu.MemImportFromShareableHandle(&b->peer_mem[p], (void *)(uintptr_t)data_fds[p], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        if (int rc = map_handle(b->peer_mem[p], b->map_bytes, 1 << 21, b->device, &b->peers.data[p])) return rc;
        void *f = nullptr;
        AMP_CUDA_TRY(cudaIpcOpenMemHandle(&f, fh[p], cudaIpcMemLazyEnablePeerAccess));
        b->opened[b->n_opened++] = f;
        b->peers.flags[p] = static_cast<uint32_t *>(f);
    }
    // would block until every rank of the team has added its device -- which is why joining is a call of its own: the caller
    // gets here only after EVERY rank has reported a successful join
    AMP_CU_TRY(cu.MulticastBindMem(b->mc, 0, b->mem, 0, b->map_bytes, 0));
    b->mc_bound = true;
    if (int rc = map_handle(b->mc, b->map_bytes, 1 << 21, b->device, &b->mc_data)) return rc;
    b->connected = true;
    return AMP_OK;
}

int64_t amp_bucket_floats(const amp_bucket_t *b) { return b ? b->floats : 0; }

float *amp_bucket_data(amp_bucket_t *b) { return b ? b->data : nullptr; }

int amp_bucket_export(amp_bucket_t *b, void *handles128) {
    AMP_REQUIRE(b && handles128, "amp_bucket_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h[2];
    AMP_CUDA_TRY(cudaIpcGetMemHandle(&h[0], b->data));
    AMP_CUDA_TRY(cudaIpcGetMemHandle(&h[1], b->flags));
    std::memcpy(handles128, h, 128);
    return AMP_OK;
}

int amp_bucket_connect(amp_bucket_t *b, const void *all_handles) {
    AMP_REQUIRE(b && all_handles, "amp_bucket_connect: NULL argument");
    AMP_REQUIRE(!b->connected, "amp_bucket_connect: already connected");
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int p = 0; p < b->world; ++p) {
        if (p == b->rank) {
            b->peers.data[p] = b->data;
            b->peers.flags[p] = b->flags;
            continue;
        }
        void *d = nullptr, *f = nullptr;
        AMP_CUDA_TRY(cudaIpcOpenMemHandle(&d, h[2 * p], cudaIpcMemLazyEnablePeerAccess));
        b->opened[b->n_opened++] = d;
        AMP_CUDA_TRY(cudaIpcOpenMemHandle(&f, h[2 * p + 1], cudaIpcMemLazyEnablePeerAccess));
        b->opened[b->n_opened++] = f;
        b->peers.data[p] = static_cast<float *>(d);
        b->peers.flags[p] = static_cast<uint32_t *>(f);
    }
    b->connected = true;
    return AMP_OK;
}

int amp_bucket_allreduce_mean(amp_bucket_t *b, int64_t offset_floats, int64_t count, void *stream) {
    AMP_REQUIRE(b, "amp_bucket_allreduce_mean: NULL handle");
    AMP_REQUIRE(b->connected, "amp_bucket_allreduce_mean: amp_bucket_connect has not been called");
    AMP_REQUIRE(offset_floats >= 0 && count >= 0 && offset_floats % 4 == 0 && offset_floats + count <= b->floats,
                "amp_bucket_allreduce_mean: range [%lld, +%lld) outside the bucket of %lld floats or not 16-byte aligned",
                (long long)offset_floats, (long long)count, (long long)b->floats);
    if (const uint32_t failed = *b->host_status)  // written by an EARLIER call's kernel on this handle; no synchronisation needed
        return fail(AMP_ECUDA,
                    "amp_bucket_allreduce_mean: an earlier all-reduce on this bucket timed out waiting for a peer (status bits 0x%x): the "
                    "replicas have diverged; the local gradient range was poisoned with NaN", failed & 0x7fffffffu);
    if (count == 0 || b->world == 1) return AMP_OK;
    // whole quads move; a ragged count is only accepted when the range ends at the end of the bucket (whose padding floats are
    // zero on every rank) -- inside the bucket the floats after a ragged end may be another producer's unfinished gradients
    AMP_REQUIRE(count % 4 == 0 || offset_floats + count == b->user_floats,
                "amp_bucket_allreduce_mean: count %lld is not a multiple of 4 and the range does not end at the bucket's end (%lld floats)",
                (long long)count, (long long)b->user_floats);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    (void)cudaStreamIsCapturing(as_stream(stream), &cap);  // capture is fine: the epoch is device state (see the kernels)
    const long long padded = (count + 3) / 4 * 4;
    const long long per = (padded / 4 + b->world - 1) / b->world;
    Control *ctl = reinterpret_cast<Control *>(b->flags + kMaxWorld);
    auto launch = [&](auto kern, int u) {
        const int grid = (int)std::max<long long>(1, std::min<long long>((per + (long long)u * kThreads - 1) / ((long long)u * kThreads), 8LL * sm_count()));
        kern<<<grid, kThreads, 0, as_stream(stream)>>>(b->peers, b->rank, b->world, offset_floats, padded, ctl, b->host_status_dev, b->spin_limit,
                                                        b->timing);
    };
    // default: the bulk-copy kernel (2 GPUs, 2.65 M floats: 34.8 us against 39.3 us with per-thread peer loads / stores and
    // 51.5 us for NCCL all-reduce + divide); AMP_B200_BUCKET_BULK=0 selects the load/store kernel, which also serves > 8 ranks
    if (b->mc_data && b->in_switch) {  // shared bucket: the NVSwitch reduces and broadcasts
        constexpr int U = 2;
        const int grid = (int)std::max<long long>(1, std::min<long long>((per + (long long)U * kThreads - 1) / ((long long)U * kThreads), 4LL * sm_count()));
        allreduce_mean_switch_kernel<U><<<grid, kThreads, 0, as_stream(stream)>>>(b->peers, b->mc_data, b->rank, b->world, offset_floats, padded, ctl,
                                                                                   b->host_status_dev, b->spin_limit, b->timing);
    } else if (b->bulk) {
        const long long chunks = (per + kChunkQuads - 1) / kChunkQuads;
        const int grid = (int)std::max<long long>(1, std::min<long long>(chunks, 4LL * sm_count()));
        auto launch_bulk = [&](auto kern, int maxw) {
            const size_t smem = (size_t)kBulkStages * (maxw + 1) * kChunkBytes;
            kern<<<grid, kThreads, smem, as_stream(stream)>>>(b->peers, b->rank, b->world, offset_floats, padded, ctl, b->host_status_dev,
                                                               b->spin_limit, b->timing);
        };
        if (b->world <= 2) launch_bulk(allreduce_mean_bulk_kernel<2>, 2);
        else if (b->world <= 4) launch_bulk(allreduce_mean_bulk_kernel<4>, 4);
        else launch_bulk(allreduce_mean_bulk_kernel<8>, 8);
    } else if (b->world <= 2) launch(allreduce_mean_kernel<2, 4>, 4);
    else if (b->world <= 4) launch(allreduce_mean_kernel<4, 4>, 4);
    else if (b->world <= 8) launch(allreduce_mean_kernel<8, 2>, 2);
    else launch(allreduce_mean_kernel<kMaxWorld, 1>, 1);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_bucket_last_timing(amp_bucket_t *b, void *stream, uint64_t *stamps4) {
    AMP_REQUIRE(b && stamps4, "amp_bucket_last_timing: NULL argument");
    AMP_CUDA_TRY(cudaMemcpyAsync(stamps4, b->timing, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, as_stream(stream)));
    AMP_CUDA_TRY(cudaStreamSynchronize(as_stream(stream)));
    return AMP_OK;
}

int amp_bucket_poll_status(amp_bucket_t *b, void *stream, uint32_t *status) {
    AMP_REQUIRE(b && status, "amp_bucket_poll_status: NULL argument");
    AMP_CUDA_TRY(cudaMemcpyAsync(status, b->flags + kMaxWorld + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, as_stream(stream)));  // Control::status
    AMP_CUDA_TRY(cudaStreamSynchronize(as_stream(stream)));
    return AMP_OK;
}

}  // extern "C"
