// K3 of the AMP hot path: discriminator forward + style reward on the 5th-generation tensor cores.
//
//   amp_disc_style_reward  <- skrl AMP._update style-reward block (upstream skrl >= 1.4.3; configured by the reference at
//                             agents/skrl_g1_dance_amp_cfg.yaml:31-39 (MLP 1024-512-1, ReLU), :80 (RunningStandardScaler),
//                             :95 (discriminator_reward_scale 2.0))
//
//   x_hat  = clamp((x - mean) / (sqrt(var) + 1e-8), -5, 5)                       fp32, then rounded to bf16
//   h1     = relu(x_hat W1^T + b1)        tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), stored bf16
//   h2     = relu(h1 W2^T + b2)           tcgen05.mma, kept in TMEM
//   logit  = h2 . w3 + b3                 fp32, folded into the TMEM read-out of layer 2 (one thread owns one row)
//   reward = -log(max(1 - 1/(1+exp(-logit)), 1e-4)) * scale
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of the A (activations) and B (weights) K-blocks into a
//               4-stage 128B-swizzled shared-memory ring, completion on mbarriers
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma (M=128, N=256, K=16) and commits to the
//               ring's "empty" barriers and to the accumulator "full" barrier
//   warps 2..5  epilogue: tcgen05.ld of the fp32 accumulator (each warp owns its 32-lane TMEM quarter, each thread one
//               output row), bias + ReLU, then either bf16 store of h1 or the running dot product with w3
//   Two 256-column accumulator stages (all 512 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Rows are processed in chunks sized so that the bf16 activations of a chunk (x_hat and h1) stay resident in the
// 126 MB L2 between the three launches of a chunk; HBM sees x once and the rewards once.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <new>

#include "amp_internal.h"
#include "amp_math.cuh"

namespace amp {
namespace disc {

constexpr int BM = 128;  // rows per tile = TMEM lanes
constexpr int BN = 256;  // accumulator columns per tile = max UMMA N
constexpr int BK = 64;   // bf16 per K-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_THREADS = 128;
constexpr int TMEM_COLS = 512;
// EPI_STORE keeps a bf16 staging area for TMA stores: per epilogue warp two 32-row x 64-column slabs (4 KiB each)
constexpr int STORE_SLAB_BYTES = 32 * BK * 2;
constexpr int STORE_STAGING_BYTES = 4 * 2 * STORE_SLAB_BYTES;  // 32 KiB
__host__ __device__ constexpr int stages_for(int epi) { return epi == 0 ? 3 : 4; }
__host__ __device__ constexpr int smem_bytes_for(int epi) {
    return stages_for(epi) * (A_STAGE_BYTES + B_STAGE_BYTES) + (epi == 0 ? STORE_STAGING_BYTES : 0) + 256 /*barriers*/ +
           1024 /*alignment slack*/;
}

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand in the canonical 128B-swizzle layout TMA produces for a
// [rows x 64 bf16] box: row r at byte r*128, 16-byte chunks XOR-swizzled with (r % 8); 8-row groups 1024 B apart.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   bits [32,46) stride byte offset >> 4 = 64 (1024 B)    bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor for kind::f16: D = fp32 (c_format 1 @ bit 4), A = B = bf16 (format 1 @ bits 7, 10), both K-major
// (bits 15, 16 = 0), N >> 3 @ bits [17,23), M >> 4 @ bits [24,29).
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct EpilogueParams {
    const float *bias;      // [N]
    __nv_bfloat16 *out;     // EPI_STORE: (M, N) row-major bf16
    const float *w3;        // EPI_REWARD: [N]
    const float *b3;        // EPI_REWARD: device scalar
    float scale;            // EPI_REWARD
    float *reward;          // EPI_REWARD: [M]
    float *logits;          // EPI_REWARD: [M] or NULL
};

enum { EPI_STORE = 0, EPI_REWARD = 1 };

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
disc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, int M, int N, int num_k_blocks, EpilogueParams ep) {
    constexpr int STAGES = stages_for(EPI);
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B needs 1024-byte aligned stage buffers
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + STAGES * A_STAGE_BYTES;
    const uint32_t staging = smem_b + STAGES * B_STAGE_BYTES;  // EPI_STORE only
    const uint32_t bars = staging + (EPI == EPI_STORE ? STORE_STAGING_BYTES : 0);
    const uint32_t full_bar = bars;                    // STAGES x 8 B
    const uint32_t empty_bar = bars + 8 * STAGES;      // STAGES x 8 B
    const uint32_t tmem_full_bar = bars + 16 * STAGES; // 2 x 8 B
    const uint32_t tmem_empty_bar = tmem_full_bar + 16;// 2 x 8 B
    const uint32_t tmem_slot = tmem_empty_bar + 16;    // 4 B
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m_tiles = (M + BM - 1) / BM, num_n_tiles = N / BN;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        if (EPI == EPI_STORE) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full_bar + 8 * i, 1);
            mbar_init(empty_bar + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmem_full_bar + 8 * i, 1);
            mbar_init(tmem_empty_bar + 8 * i, NUM_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // whole warp: allocate all 512 TMEM columns, publish the base address through shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
                for (int nt = 0; nt < num_n_tiles; ++nt) {
                    for (int kb = 0; kb < num_k_blocks; ++kb) {
                        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                        mbar_arrive_expect_tx(full_bar + 8 * stage, A_STAGE_BYTES + B_STAGE_BYTES);
                        tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, kb * BK, mt * BM, full_bar + 8 * stage);
                        tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_b, kb * BK, nt * BN, full_bar + 8 * stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
                for (int nt = 0; nt < num_n_tiles; ++nt) {
                    mbar_wait(tmem_empty_bar + 8 * acc, acc_phase ^ 1);  // epilogue drained this accumulator
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < num_k_blocks; ++kb) {
                        mbar_wait(full_bar + 8 * stage, phase);          // TMA bytes landed
                        tcgen05_fence_after();
                        const uint64_t a0 = make_kmajor_sw128_desc(smem_a + stage * A_STAGE_BYTES);
                        const uint64_t b0 = make_kmajor_sw128_desc(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
                            umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, kInstrDesc, (uint32_t)((kb | k) != 0));
                        }
                        umma_commit(empty_bar + 8 * stage);              // frees the smem slot once these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(tmem_full_bar + 8 * acc);                // accumulator complete -> epilogue
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else {
        // ================= epilogue warps 2..5 =================
        const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t store_it = 0;  // EPI_STORE: running count of TMA stores issued by this warp (selects the slab)
        for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
            const int row = mt * BM + quarter * 32 + lane;
            float dot = 0.0f;
            for (int nt = 0; nt < num_n_tiles; ++nt) {
                mbar_wait(tmem_full_bar + 8 * acc, acc_phase);
                tcgen05_fence_after();
                if constexpr (EPI == EPI_STORE) {
                    // bias + ReLU -> bf16, staged per warp as a [32 rows x 64 cols] slab in the 128B-swizzled layout and
                    // written with one TMA store per slab (full 128-byte lines; rows past M are clipped by the tensor map)
#pragma unroll 1
                    for (int cb = 0; cb < BN / BK; ++cb) {
                        const uint32_t slab = staging + (uint32_t)((quarter * 2 + (store_it & 1)) * STORE_SLAB_BYTES);
                        if (lane == 0) bulk_wait_read<1>();  // the store that last read this slab has drained it
                        __syncwarp();
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t v[32];
                            tmem_ld_32x32(tmem_base + lane_base + (uint32_t)(acc * BN + cb * BK + half * 32), v);
                            tmem_ld_wait();
                            const float4 *bias4 = reinterpret_cast<const float4 *>(ep.bias + nt * BN + cb * BK + half * 32);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {  // 16-byte chunk (half*4 + j) of this thread's 128-byte row
                                const float4 b0 = __ldg(bias4 + 2 * j), b1 = __ldg(bias4 + 2 * j + 1);
                                const __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, 0.0f),
                                                                                fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, 0.0f));
                                const __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, 0.0f),
                                                                                fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, 0.0f));
                                const __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, 0.0f),
                                                                                fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, 0.0f));
                                const __nv_bfloat162 p3 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, 0.0f),
                                                                                fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, 0.0f));
                                const int chunk = (half * 4 + j) ^ (lane & 7);  // SWIZZLE_128B: chunk index XOR (row % 8)
                                st_shared_v4(slab + (uint32_t)(lane * 128 + chunk * 16), *reinterpret_cast<const uint32_t *>(&p0),
                                             *reinterpret_cast<const uint32_t *>(&p1), *reinterpret_cast<const uint32_t *>(&p2),
                                             *reinterpret_cast<const uint32_t *>(&p3));
                            }
                        }
                        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmap_out, slab, nt * BN + cb * BK, mt * BM + quarter * 32);
                            bulk_commit();
                        }
                        ++store_it;
                    }
                } else {
#pragma unroll 1
                    for (int chunk = 0; chunk < BN / 32; ++chunk) {
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + lane_base + (uint32_t)(acc * BN + chunk * 32), v);
                        tmem_ld_wait();
                        const int col0 = nt * BN + chunk * 32;
                        const float4 *bias4 = reinterpret_cast<const float4 *>(ep.bias + col0);
                        const float4 *w4 = reinterpret_cast<const float4 *>(ep.w3 + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bb = __ldg(bias4 + j), ww = __ldg(w4 + j);
                            dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 0]) + bb.x, 0.0f), ww.x, dot);
                            dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.0f), ww.y, dot);
                            dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.0f), ww.z, dot);
                            dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.0f), ww.w, dot);
                        }
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(tmem_empty_bar + 8 * acc);  // this thread is done reading the accumulator stage
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            if constexpr (EPI == EPI_REWARD) {
                if (row < M) {
                    const float logit = dot + __ldg(ep.b3);
                    if (ep.logits) ep.logits[row] = logit;
                    ep.reward[row] = style_reward(logit, ep.scale);
                }
            }
        }
        if constexpr (EPI == EPI_STORE) {
            if (lane == 0) bulk_wait_all();  // all TMA stores of this warp have completed before the CTA exits
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- RunningStandardScaler (eval) + bf16 cast: x (M, in) fp32 -> x_hat (M, Kp) bf16, zero padded -----------------------
// One warp per row; lane l owns the column pairs (2(l + 32 j), +1), j < NB = Kp / 64, of EVERY row it visits, so the
// scaler statistics of its columns live in registers and a row costs NB 8-byte loads and NB 4-byte stores per lane, all
// fully coalesced.  VEC = rows are 8-byte aligned (even stride), else scalar loads.
template <int NB, bool VEC>
__global__ void __launch_bounds__(256) normalise_cast_kernel(const float *__restrict__ x, int64_t x_stride, int64_t M,
                                                              int in_features, const float *__restrict__ mean,
                                                              const float *__restrict__ denom,
                                                              __nv_bfloat16 *__restrict__ out) {
    constexpr int Kp = NB * 64;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float2 mu[NB], dn[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int c = 2 * (lane + 32 * j);
        mu[j] = make_float2(c < in_features ? mean[c] : 0.0f, c + 1 < in_features ? mean[c + 1] : 0.0f);
        dn[j] = make_float2(c < in_features ? denom[c] : 1.0f, c + 1 < in_features ? denom[c + 1] : 1.0f);
    }
    for (int64_t r = warp; r < M; r += nwarps) {
        const float *xr = x + r * x_stride;
        __nv_bfloat162 *orow = reinterpret_cast<__nv_bfloat162 *>(out + r * Kp);
        float2 v[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int c = 2 * (lane + 32 * j);
            v[j] = make_float2(0.0f, 0.0f);
            if (VEC && c + 1 < in_features) {
                v[j] = __ldcs(reinterpret_cast<const float2 *>(xr + c));
            } else {
                if (c < in_features) v[j].x = __ldcs(xr + c);
                if (c + 1 < in_features) v[j].y = __ldcs(xr + c + 1);
            }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int c = 2 * (lane + 32 * j);
            // clamp((x - mean) / (sqrt(var) + 1e-8), -5, 5); padding columns stay exactly 0
            float a = fminf(fmaxf(__fdiv_rn(__fsub_rn(v[j].x, mu[j].x), dn[j].x), -5.0f), 5.0f);
            float b = fminf(fmaxf(__fdiv_rn(__fsub_rn(v[j].y, mu[j].y), dn[j].y), -5.0f), 5.0f);
            if (c >= in_features) a = 0.0f;
            if (c + 1 >= in_features) b = 0.0f;
            orow[lane + 32 * j] = __floats2bfloat162_rn(a, b);
        }
    }
}

template <bool VEC>
static int launch_normalise_cast(int nb, int grid, cudaStream_t st, const float *x, int64_t x_stride, int64_t rows, int in_features,
                                 const float *mean, const float *denom, __nv_bfloat16 *out) {
#define AMP_CAST_CASE(NBV)                                                                                            \
    case NBV:                                                                                                         \
        normalise_cast_kernel<NBV, VEC><<<grid, 256, 0, st>>>(x, x_stride, rows, in_features, mean, denom, out);      \
        break
    switch (nb) {
        AMP_CAST_CASE(1); AMP_CAST_CASE(2); AMP_CAST_CASE(3); AMP_CAST_CASE(4); AMP_CAST_CASE(5); AMP_CAST_CASE(6);
        AMP_CAST_CASE(7); AMP_CAST_CASE(8); AMP_CAST_CASE(9); AMP_CAST_CASE(10); AMP_CAST_CASE(11); AMP_CAST_CASE(12);
        AMP_CAST_CASE(13); AMP_CAST_CASE(14); AMP_CAST_CASE(15); AMP_CAST_CASE(16);
        default: return fail(AMP_EINVAL, "discriminator input wider than 1024 columns is not supported (got %d K-blocks)", nb);
    }
#undef AMP_CAST_CASE
    return AMP_OK;
}

// fp32 master (rows, cols) -> bf16 (rows, cols_padded), zero padded
__global__ void cast_weight_kernel(const float *__restrict__ w, int rows, int cols, int cols_padded,
                                   __nv_bfloat16 *__restrict__ out) {
    const int64_t total = (int64_t)rows * cols_padded;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols_padded), c = (int)(e - (int64_t)r * cols_padded);
        out[e] = __float2bfloat16_rn(c < cols ? w[(int64_t)r * cols + c] : 0.0f);
    }
}

// skrl scaler statistics are float64 buffers used as .float(): mean_f = (float)mean, denom = sqrt((float)var) + 1e-8
__global__ void scaler_stats_kernel(const double *__restrict__ mean, const double *__restrict__ var, int n,
                                    float *__restrict__ mean_f, float *__restrict__ denom_f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        mean_f[i] = __double2float_rn(mean[i]);
        denom_f[i] = __fadd_rn(__fsqrt_rn(__double2float_rn(var[i])), 1e-8f);
    }
}

__global__ void reward_from_logits_kernel(const float *__restrict__ logits, int64_t M, float scale, float *__restrict__ reward) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
        reward[i] = style_reward(logits[i], scale);
}

// ---- host side: tensor maps --------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved through the runtime so the library carries no link-time dependency on libcuda.so
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// (rows, cols) bf16 row-major with pitch `pitch` elements; box = [box_rows x 64 cols], 128B swizzle, OOB rows read as 0
static int make_tmap(CUtensorMap *map, const void *ptr, int64_t rows, int64_t cols, int64_t pitch, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(AMP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AMP_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return AMP_OK;
}

}  // namespace disc
}  // namespace amp

struct amp_disc {
    int in_features, Kp, h1, h2;
    int64_t chunk_rows;
    int device;
    __nv_bfloat16 *W1, *W2;       // (h1, Kp), (h2, h1) bf16
    float *b1, *b2, *w3, *b3;     // fp32
    float *mean, *denom;          // fp32 [in_features]
    __nv_bfloat16 *xhat, *hid;    // workspaces (chunk_rows, Kp), (chunk_rows, h1)
    CUtensorMap tmap_w1, tmap_w2; // weights never move: encoded once
    bool loaded;
};

using namespace amp;
using namespace amp::disc;

extern "C" {

int amp_disc_create(int32_t in_features, int32_t h1, int32_t h2, int64_t max_rows, void *stream, amp_disc_t **out) {
    (void)stream;
    AMP_REQUIRE(out, "amp_disc_create: NULL out");
    *out = nullptr;
    AMP_REQUIRE(in_features >= 1 && h1 >= BN && h2 >= BN && h1 % BN == 0 && h2 % BN == 0,
                "amp_disc_create: hidden sizes must be multiples of %d (got %d, %d), in_features >= 1 (got %d)", BN, h1, h2,
                in_features);
    AMP_REQUIRE(max_rows >= 1, "amp_disc_create: max_rows must be positive");
    int dev = 0, major = 0;
    AMP_CUDA_TRY(cudaGetDevice(&dev));
    AMP_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(AMP_ENODEV, "amp_disc_create: tcgen05 kernels need an sm_100 device (found compute capability %d.x)", major);

    amp_disc *d = new (std::nothrow) amp_disc();
    if (!d) return fail(AMP_ENOMEM, "amp_disc_create: host allocation failed");
    std::memset(d, 0, sizeof(*d));
    d->device = dev;
    d->in_features = in_features;
    d->Kp = (in_features + BK - 1) / BK * BK;
    d->h1 = h1;
    d->h2 = h2;
    // two 128-row tiles per SM per chunk: full waves, and x_hat + h1 of a chunk (~80 MB at h1 = 1024) stay in L2
    const int64_t wave_rows = (int64_t)sm_count() * BM * 2;
    d->chunk_rows = std::min<int64_t>((max_rows + BM - 1) / BM * BM, wave_rows);

    auto alloc = [&](void **p, size_t bytes) { return cudaMalloc(p, bytes); };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = alloc((void **)&d->W1, (size_t)h1 * d->Kp * 2);
    if (e == cudaSuccess) e = alloc((void **)&d->W2, (size_t)h2 * h1 * 2);
    if (e == cudaSuccess) e = alloc((void **)&d->b1, (size_t)h1 * 4);
    if (e == cudaSuccess) e = alloc((void **)&d->b2, (size_t)h2 * 4);
    if (e == cudaSuccess) e = alloc((void **)&d->w3, (size_t)h2 * 4);
    if (e == cudaSuccess) e = alloc((void **)&d->b3, 4);
    if (e == cudaSuccess) e = alloc((void **)&d->mean, (size_t)in_features * 4);
    if (e == cudaSuccess) e = alloc((void **)&d->denom, (size_t)in_features * 4);
    if (e == cudaSuccess) e = alloc((void **)&d->xhat, (size_t)d->chunk_rows * d->Kp * 2);
    if (e == cudaSuccess) e = alloc((void **)&d->hid, (size_t)d->chunk_rows * h1 * 2);
    if (e != cudaSuccess) {
        amp_disc_destroy(d);
        return cuda_fail(e, "cudaMalloc(amp_disc_create)");
    }
    int rc = make_tmap(&d->tmap_w1, d->W1, h1, d->Kp, d->Kp, BN);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_w2, d->W2, h2, h1, h1, BN);
    if (rc == AMP_OK) {
        e = cudaFuncSetAttribute(disc_gemm_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(EPI_STORE));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(disc_gemm_kernel<EPI_REWARD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smem_bytes_for(EPI_REWARD));
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaFuncSetAttribute(disc_gemm_kernel)");
    }
    if (rc != AMP_OK) {
        amp_disc_destroy(d);
        return rc;
    }
    *out = d;
    return AMP_OK;
}

int amp_disc_destroy(amp_disc_t *d) {
    if (!d) return AMP_OK;
    void *ptrs[] = {d->W1, d->W2, d->b1, d->b2, d->w3, d->b3, d->mean, d->denom, d->xhat, d->hid};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete d;
    return AMP_OK;
}

int64_t amp_disc_chunk_rows(const amp_disc_t *d) { return d ? d->chunk_rows : 0; }

int amp_disc_load(amp_disc_t *d, const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                  const float *b3, const double *running_mean, const double *running_variance, void *stream) {
    AMP_REQUIRE(d && W1 && b1 && W2 && b2 && W3 && b3 && running_mean && running_variance, "amp_disc_load: NULL argument");
    cudaStream_t st = as_stream(stream);
    const int blocks = sm_count() * 4;
    cast_weight_kernel<<<blocks, 256, 0, st>>>(W1, d->h1, d->in_features, d->Kp, d->W1);
    cast_weight_kernel<<<blocks, 256, 0, st>>>(W2, d->h2, d->h1, d->h1, d->W2);
    AMP_CUDA_TRY(cudaGetLastError());
    AMP_CUDA_TRY(cudaMemcpyAsync(d->b1, b1, (size_t)d->h1 * 4, cudaMemcpyDeviceToDevice, st));
    AMP_CUDA_TRY(cudaMemcpyAsync(d->b2, b2, (size_t)d->h2 * 4, cudaMemcpyDeviceToDevice, st));
    AMP_CUDA_TRY(cudaMemcpyAsync(d->w3, W3, (size_t)d->h2 * 4, cudaMemcpyDeviceToDevice, st));
    AMP_CUDA_TRY(cudaMemcpyAsync(d->b3, b3, 4, cudaMemcpyDeviceToDevice, st));
    scaler_stats_kernel<<<(d->in_features + 255) / 256, 256, 0, st>>>(running_mean, running_variance, d->in_features,
                                                                      d->mean, d->denom);
    AMP_CUDA_TRY(cudaGetLastError());
    d->loaded = true;
    return AMP_OK;
}

int amp_disc_style_reward(amp_disc_t *d, const float *x, int64_t x_stride, int64_t M, float reward_scale, float *reward,
                          float *logits, void *stream) {
    AMP_REQUIRE(d && M >= 0, "amp_disc_style_reward: bad handle or negative size");
    AMP_REQUIRE(d->loaded, "amp_disc_style_reward: amp_disc_load has not been called");
    if (M == 0) return AMP_OK;
    AMP_REQUIRE(x && reward, "amp_disc_style_reward: NULL buffer");
    AMP_REQUIRE(x_stride >= d->in_features, "amp_disc_style_reward: x_stride %lld < in_features %d", (long long)x_stride,
                d->in_features);
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    for (int64_t r0 = 0; r0 < M; r0 += d->chunk_rows) {
        const int64_t rows = std::min<int64_t>(d->chunk_rows, M - r0);
        const int m_tiles = (int)((rows + BM - 1) / BM);
        const int grid = std::min(m_tiles, sms);
        const float *xc = x + r0 * x_stride;
        const bool vec = (x_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(xc) & 7u) == 0);
        const int cast_grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
        int rc = vec ? launch_normalise_cast<true>(d->Kp / BK, cast_grid, st, xc, x_stride, rows, d->in_features, d->mean, d->denom, d->xhat)
                     : launch_normalise_cast<false>(d->Kp / BK, cast_grid, st, xc, x_stride, rows, d->in_features, d->mean, d->denom, d->xhat);
        if (rc != AMP_OK) return rc;
        AMP_CUDA_TRY(cudaGetLastError());

        CUtensorMap tm_x, tm_h, tm_h_store;
        rc = make_tmap(&tm_x, d->xhat, rows, d->Kp, d->Kp, BM);
        if (rc != AMP_OK) return rc;
        rc = make_tmap(&tm_h, d->hid, rows, d->h1, d->h1, BM);
        if (rc != AMP_OK) return rc;
        rc = make_tmap(&tm_h_store, d->hid, rows, d->h1, d->h1, 32);  // epilogue slabs: 32 rows x 64 columns
        if (rc != AMP_OK) return rc;

        EpilogueParams e1{};
        e1.bias = d->b1;
        e1.out = d->hid;
        disc_gemm_kernel<EPI_STORE><<<grid, NUM_THREADS, smem_bytes_for(EPI_STORE), st>>>(tm_x, d->tmap_w1, tm_h_store, (int)rows,
                                                                                          d->h1, d->Kp / BK, e1);
        AMP_CUDA_TRY(cudaGetLastError());

        EpilogueParams e2{};
        e2.bias = d->b2;
        e2.w3 = d->w3;
        e2.b3 = d->b3;
        e2.scale = reward_scale;
        e2.reward = reward + r0;
        e2.logits = logits ? logits + r0 : nullptr;
        disc_gemm_kernel<EPI_REWARD><<<grid, NUM_THREADS, smem_bytes_for(EPI_REWARD), st>>>(tm_h, d->tmap_w2, tm_h, (int)rows, d->h2,
                                                                                            d->h1 / BK, e2);
        AMP_CUDA_TRY(cudaGetLastError());
    }
    return AMP_OK;
}

int amp_style_reward_from_logits(const float *logits, int64_t M, float reward_scale, float *reward, void *stream) {
    AMP_REQUIRE(M >= 0, "amp_style_reward_from_logits: negative size");
    if (M == 0) return AMP_OK;
    AMP_REQUIRE(logits && reward, "amp_style_reward_from_logits: NULL buffer");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((M + 255) / 256, (int64_t)sm_count() * 8));
    reward_from_logits_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, M, reward_scale, reward);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

}  // extern "C"
