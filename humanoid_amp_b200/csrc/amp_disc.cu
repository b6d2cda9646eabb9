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
// Structure: see disc_fused_kernel below (one persistent CTA per SM, 320 threads, warp-specialised: TMA producer, single
// thread tcgen05.mma issuer, two sets of four epilogue warps).  Rows are processed in chunks whose bf16 x_hat stays
// L2-resident between the scaler/cast kernel and the fused kernel; the cast of chunk c+1 runs on a side stream under the
// fused kernel of chunk c.
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
constexpr int NUM_EPI_THREADS = 128;
constexpr int TMEM_COLS = 512;
// bf16 staging area for the h1 TMA stores: per E1 warp two 32-row x 64-column slabs (4 KiB each)
constexpr int STORE_SLAB_BYTES = 32 * BK * 2;
constexpr int STORE_STAGING_BYTES = 4 * 2 * STORE_SLAB_BYTES;  // 32 KiB

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

// =====================================================================================================================
// Fused two-layer kernel: layer 1 and layer 2 share ONE persistent launch and keep the tensor pipe busy.
//
// Why not keep h1 on chip: one 128-row tile needs h1 = 128 x 1024 bf16 = 256 KB (> 227 KB smem) for the K loop of layer 2,
// and its accumulator D2 = 128 x 512 fp32 already fills all 512 TMEM columns.  So the accumulators are halved
// (D1 = columns [0,256), D2 = [256,512)) and h1 makes a round trip through a per-CTA, double-buffered workspace that
// never leaves the L2 (148 CTAs x 2 slots x 256 KB = 77.6 MB), written with TMA stores and read back with TMA loads.
//
// The single MMA-issuing thread interleaves the two GEMMs in a static software pipeline over the CTA's row tiles:
//     prologue   G1(t0, nt = 0..3)
//     iteration i, group g = 0..3:   G1(t[i+1], nt = g)  then  G2(t[i]) units [8g, 8g+8)        (unit = one 64-wide K block
//                                                                                                of one 256-column N tile)
// so while the epilogue warps E1 drain D1 (bias + ReLU + bf16 + swizzled st.shared + TMA store; ~3k cycles, the slow part of
// the split version) the tensor pipe runs 8 units = 4096 cycles of layer-2 MMAs, and while E2 drains D2 it runs a layer-1
// tile.  Per row tile the pipe is busy 4*1536 + 32*512 = 22528 cycles, the MMA-bound floor of the two layers together.
//
// Warp roles (320 threads): 0 TMA producer, 1 TMEM alloc + MMA issuer, 2..5 E1 (D1 -> h1), 6..9 E2 (D2 -> logit, reward).
// Producer and issuer walk the same schedule (walk_schedule); E1/E2 just follow their accumulator barriers.
// =====================================================================================================================
constexpr int FUSED_STAGES = 4;
constexpr int FUSED_THREADS = 320;
constexpr int FUSED_SMEM_BYTES = FUSED_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + STORE_STAGING_BYTES + 256 + 1024;
constexpr uint32_t D1_COL = 0, D2_COL = 256;

template <class G1, class G2>
__device__ __forceinline__ void walk_schedule(int T, int n1_tiles, int units, G1 &&g1, G2 &&g2) {
    if (T <= 0) return;
    for (int nt = 0; nt < n1_tiles; ++nt) g1(0, nt);
    const int per_group = (units + n1_tiles - 1) / n1_tiles;
    for (int i = 0; i < T; ++i) {
        int u = 0;
        for (int g = 0; g < n1_tiles; ++g) {
            if (i + 1 < T) g1(i + 1, g);
            const int end = min(units, u + per_group);
            for (; u < end; ++u) g2(i, u);
        }
    }
}

struct FusedParams {
    int M;          // rows of this launch (x_hat rows)
    int kb1;        // K blocks of layer 1 (Kp / 64)
    int n1_tiles;   // h1 / 256
    int n2_tiles;   // h2 / 256
    const float *b1, *b2, *w3, *b3;
    float scale;
    float *reward;  // [M]
    float *logits;  // [M] or NULL
};

__global__ void __launch_bounds__(FUSED_THREADS, 1)
disc_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                  const __grid_constant__ CUtensorMap tmap_h_load, const __grid_constant__ CUtensorMap tmap_h_store,
                  const __grid_constant__ CUtensorMap tmap_w2, FusedParams p) {
    constexpr int STAGES = FUSED_STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + STAGES * A_STAGE_BYTES;
    const uint32_t staging = smem_b + STAGES * B_STAGE_BYTES;
    const uint32_t bars = staging + STORE_STAGING_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
    const uint32_t d1_full = bars + 16 * STAGES, d1_empty = d1_full + 8, d2_full = d1_full + 16, d2_empty = d1_full + 24;
    const uint32_t h1_ready = d1_full + 32;  // 2 x 8 B
    const uint32_t tmem_slot = h1_ready + 16;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m_tiles = (p.M + BM - 1) / BM;
    const int T = ((int)blockIdx.x < num_m_tiles) ? (num_m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int kb2 = 4 * p.n1_tiles;            // K blocks of layer 2 = h1 / 64
    const int units = p.n2_tiles * kb2;        // layer-2 units per row tile
    const int slot_row0 = (int)blockIdx.x * 2 * BM;  // first row of this CTA's two h1 slots in the workspace

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_h_load) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_h_store) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w2) : "memory");
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full_bar + 8 * i, 1);
            mbar_init(empty_bar + 8 * i, 1);
        }
        mbar_init(d1_full, 1);
        mbar_init(d1_empty, NUM_EPI_THREADS);
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, NUM_EPI_THREADS);
        mbar_init(h1_ready, 4);
        mbar_init(h1_ready + 8, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
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
            auto load_pair = [&](const CUtensorMap *ma, int a_col, int a_row, const CUtensorMap *mb, int b_col, int b_row) {
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                mbar_arrive_expect_tx(full_bar + 8 * stage, A_STAGE_BYTES + B_STAGE_BYTES);
                tma_load_2d(smem_a + stage * A_STAGE_BYTES, ma, a_col, a_row, full_bar + 8 * stage);
                tma_load_2d(smem_b + stage * B_STAGE_BYTES, mb, b_col, b_row, full_bar + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            };
            walk_schedule(
                T, p.n1_tiles, units,
                [&](int ti, int nt) {
                    const int m = (int)blockIdx.x + ti * (int)gridDim.x;
                    for (int kb = 0; kb < p.kb1; ++kb) load_pair(&tmap_x, kb * BK, m * BM, &tmap_w1, kb * BK, nt * BN);
                },
                [&](int ti, int u) {
                    const int n2 = u / kb2, kb = u - n2 * kb2;
                    if (u == 0) {  // h1 of this row tile has been written (all four E1 warps' TMA stores completed)
                        mbar_wait(h1_ready + 8 * (ti & 1), (uint32_t)((ti >> 1) & 1));
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                    load_pair(&tmap_h_load, kb * BK, slot_row0 + (ti & 1) * BM, &tmap_w2, kb * BK, n2 * BN);
                });
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, c1 = 0, c2 = 0;
            auto mma_block = [&](uint32_t d_tmem, bool first) {
                mbar_wait(full_bar + 8 * stage, phase);
                tcgen05_fence_after();
                const uint64_t a0 = make_kmajor_sw128_desc(smem_a + stage * A_STAGE_BYTES);
                const uint64_t b0 = make_kmajor_sw128_desc(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, kInstrDesc, (uint32_t)(!first || k != 0));
                umma_commit(empty_bar + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            };
            walk_schedule(
                T, p.n1_tiles, units,
                [&](int, int) {
                    mbar_wait(d1_empty, (c1 & 1) ^ 1);  // E1 has drained D1
                    tcgen05_fence_after();
                    for (int kb = 0; kb < p.kb1; ++kb) mma_block(tmem_base + D1_COL, kb == 0);
                    umma_commit(d1_full);
                    ++c1;
                },
                [&](int, int u) {
                    const int kb = u % kb2;
                    if (kb == 0) {
                        mbar_wait(d2_empty, (c2 & 1) ^ 1);  // E2 has drained D2
                        tcgen05_fence_after();
                    }
                    mma_block(tmem_base + D2_COL, kb == 0);
                    if (kb == kb2 - 1) {
                        umma_commit(d2_full);
                        ++c2;
                    }
                });
        }
    } else if (warp < 6) {
        // ================= E1: D1 -> bias + ReLU -> bf16 h1 slabs -> TMA store into the CTA's L2-resident slot =================
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t c1 = 0, store_it = 0;
        for (int ti = 0; ti < T; ++ti) {
            const int row0 = slot_row0 + (ti & 1) * BM + quarter * 32;
            for (int nt = 0; nt < p.n1_tiles; ++nt) {
                mbar_wait(d1_full, c1 & 1);
                tcgen05_fence_after();
#pragma unroll 1
                for (int cb = 0; cb < BN / BK; ++cb) {
                    const uint32_t slab = staging + (uint32_t)((quarter * 2 + (store_it & 1)) * STORE_SLAB_BYTES);
                    if (lane == 0) bulk_wait_read<1>();
                    __syncwarp();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + lane_base + D1_COL + (uint32_t)(cb * BK + half * 32), v);
                        tmem_ld_wait();
                        const float4 *bias4 = reinterpret_cast<const float4 *>(p.b1 + nt * BN + cb * BK + half * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b0 = __ldg(bias4 + 2 * j), b1v = __ldg(bias4 + 2 * j + 1);
                            const __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, 0.0f),
                                                                            fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, 0.0f));
                            const __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, 0.0f),
                                                                            fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, 0.0f));
                            const __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 4]) + b1v.x, 0.0f),
                                                                            fmaxf(__uint_as_float(v[8 * j + 5]) + b1v.y, 0.0f));
                            const __nv_bfloat162 p3 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 6]) + b1v.z, 0.0f),
                                                                            fmaxf(__uint_as_float(v[8 * j + 7]) + b1v.w, 0.0f));
                            const int chunk = (half * 4 + j) ^ (lane & 7);
                            st_shared_v4(slab + (uint32_t)(lane * 128 + chunk * 16), *reinterpret_cast<const uint32_t *>(&p0),
                                         *reinterpret_cast<const uint32_t *>(&p1), *reinterpret_cast<const uint32_t *>(&p2),
                                         *reinterpret_cast<const uint32_t *>(&p3));
                        }
                    }
                    if (cb == BN / BK - 1) {  // last TMEM read of this accumulator: release D1 to the issuer early
                        tcgen05_fence_before();
                        mbar_arrive(d1_empty);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_h_store, slab, nt * BN + cb * BK, row0);
                        bulk_commit();
                    }
                    ++store_it;
                }
                ++c1;
            }
            if (lane == 0) {  // this warp's 32 rows of h1 are complete in the workspace
                bulk_wait_all();
                mbar_arrive(h1_ready + 8 * (ti & 1));
            }
        }
    } else {
        // ================= E2: D2 -> bias + ReLU -> running dot with w3 -> logit -> reward =================
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t c2 = 0;
        for (int ti = 0; ti < T; ++ti) {
            const int m = (int)blockIdx.x + ti * (int)gridDim.x;
            const int row = m * BM + quarter * 32 + lane;
            float dot = 0.0f;
            for (int n2 = 0; n2 < p.n2_tiles; ++n2) {
                mbar_wait(d2_full, c2 & 1);
                tcgen05_fence_after();
#pragma unroll 1
                for (int chunk = 0; chunk < BN / 32; ++chunk) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + lane_base + D2_COL + (uint32_t)(chunk * 32), v);
                    tmem_ld_wait();
                    const int col0 = n2 * BN + chunk * 32;
                    const float4 *bias4 = reinterpret_cast<const float4 *>(p.b2 + col0);
                    const float4 *w4 = reinterpret_cast<const float4 *>(p.w3 + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = __ldg(bias4 + j), ww = __ldg(w4 + j);
                        dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 0]) + bb.x, 0.0f), ww.x, dot);
                        dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.0f), ww.y, dot);
                        dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.0f), ww.z, dot);
                        dot = fmaf(fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.0f), ww.w, dot);
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(d2_empty);
                ++c2;
            }
            if (row < p.M) {
                const float logit = dot + __ldg(p.b3);
                if (p.logits) p.logits[row] = logit;
                p.reward[row] = style_reward(logit, p.scale);
            }
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
    // (x - mean) / denom is evaluated as (x - mean) * (1 / denom) with an IEEE reciprocal held in registers: at most 1 ulp
    // (fp32) from the division, and the result is rounded to bf16 (8-bit mantissa) right after, so the two agree except
    // when the quotient sits within 2^-16 relative of a bf16 rounding boundary.  Padding columns get rcp = 0 -> exactly 0.
    float2 mu[NB], rc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int c = 2 * (lane + 32 * j);
        mu[j] = make_float2(c < in_features ? mean[c] : 0.0f, c + 1 < in_features ? mean[c + 1] : 0.0f);
        rc[j] = make_float2(c < in_features ? __frcp_rn(denom[c]) : 0.0f, c + 1 < in_features ? __frcp_rn(denom[c + 1]) : 0.0f);
    }
    auto load_row = [&](int64_t r, float2(&v)[NB]) {
        const float *xr = x + r * x_stride;
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
    };
    auto store_row = [&](int64_t r, const float2(&v)[NB]) {
        __nv_bfloat162 *orow = reinterpret_cast<__nv_bfloat162 *>(out + r * Kp);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const float a = fminf(fmaxf(__fmul_rn(__fsub_rn(v[j].x, mu[j].x), rc[j].x), -5.0f), 5.0f);
            const float b = fminf(fmaxf(__fmul_rn(__fsub_rn(v[j].y, mu[j].y), rc[j].y), -5.0f), 5.0f);
            orow[lane + 32 * j] = __floats2bfloat162_rn(a, b);
        }
    };
    int64_t r = warp;
    if constexpr (NB <= 4) {  // two rows in flight per warp while the register budget allows it
        for (; r + nwarps < M; r += 2 * nwarps) {
            float2 v0[NB], v1[NB];
            load_row(r, v0);
            load_row(r + nwarps, v1);
            store_row(r, v0);
            store_row(r + nwarps, v1);
        }
    }
    for (; r < M; r += nwarps) {
        float2 v0[NB];
        load_row(r, v0);
        store_row(r, v0);
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
    int ws_ctas;  // persistent CTAs the h1 workspace was sized for
    int device;
    __nv_bfloat16 *W1, *W2;       // (h1, Kp), (h2, h1) bf16
    float *b1, *b2, *w3, *b3;     // fp32
    float *mean, *denom;          // fp32 [in_features]
    __nv_bfloat16 *xhat[2], *hid; // workspaces 2 x (chunk_rows, Kp) (double buffer) and (ws_ctas * 2 * 128, h1)
    cudaStream_t side;            // the scaler/cast of chunk i+1 runs here, under the fused kernel of chunk i
    cudaEvent_t ev_start, ev_ready[2], ev_free[2];
    CUtensorMap tmap_w1, tmap_w2; // weights never move: encoded once
    CUtensorMap tmap_h_load, tmap_h_store;  // h1 workspace: 128-row loads, 32-row epilogue slab stores
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
    // Rows per launch ("chunk"): the bf16 x_hat of a chunk should stay L2-resident between the cast kernel that writes it
    // and the fused kernel that reads it (<= ~64 MB), in whole waves of 128-row tiles: 8 tiles per SM at K*A = 166.
    const int64_t per_wave_bytes = (int64_t)sm_count() * BM * d->Kp * 2;
    const int64_t tiles_per_cta = std::max<int64_t>(1, std::min<int64_t>(8, ((int64_t)64 << 20) / per_wave_bytes));
    d->chunk_rows = std::min<int64_t>((max_rows + BM - 1) / BM * BM, (int64_t)sm_count() * BM * tiles_per_cta);
    d->ws_ctas = sm_count();  // h1 workspace: two 128-row slots per persistent CTA, L2-resident
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
    if (e == cudaSuccess) e = alloc((void **)&d->xhat[0], (size_t)d->chunk_rows * d->Kp * 2);
    if (e == cudaSuccess && max_rows > d->chunk_rows) e = alloc((void **)&d->xhat[1], (size_t)d->chunk_rows * d->Kp * 2);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) {
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_ready[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_free[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = alloc((void **)&d->hid, (size_t)d->ws_ctas * 2 * BM * h1 * 2);
    if (e != cudaSuccess) {
        amp_disc_destroy(d);
        return cuda_fail(e, "cudaMalloc(amp_disc_create)");
    }
    int rc = make_tmap(&d->tmap_w1, d->W1, h1, d->Kp, d->Kp, BN);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_w2, d->W2, h2, h1, h1, BN);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_h_load, d->hid, (int64_t)d->ws_ctas * 2 * BM, h1, h1, BM);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_h_store, d->hid, (int64_t)d->ws_ctas * 2 * BM, h1, h1, 32);
    if (rc == AMP_OK) {
        e = cudaFuncSetAttribute(disc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaFuncSetAttribute(disc_fused_kernel)");
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
    void *ptrs[] = {d->W1, d->W2, d->b1, d->b2, d->w3, d->b3, d->mean, d->denom, d->xhat[0], d->xhat[1], d->hid};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (d->side) cudaStreamDestroy(d->side);
    if (d->ev_start) cudaEventDestroy(d->ev_start);
    for (int i = 0; i < 2; ++i) {
        if (d->ev_ready[i]) cudaEventDestroy(d->ev_ready[i]);
        if (d->ev_free[i]) cudaEventDestroy(d->ev_free[i]);
    }
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
    const int64_t n_chunks = (M + d->chunk_rows - 1) / d->chunk_rows;
    AMP_REQUIRE(n_chunks == 1 || d->xhat[1], "amp_disc_style_reward: %lld rows exceed the max_rows given to amp_disc_create",
                (long long)M);
    // With several chunks the scaler/cast of chunk c+1 runs on the handle's side stream underneath the fused kernel of
    // chunk c (the fused kernel is operand-delivery bound and leaves HBM and most issue slots idle; the cast kernel uses no
    // shared memory, so its CTAs co-reside with the persistent fused CTAs).  A single chunk stays on the caller's stream.
    const bool overlap = n_chunks > 1;
    auto issue_cast = [&](int64_t c) -> int {
        const int b = (int)(c & 1);
        const int64_t r0 = c * d->chunk_rows, rows = std::min<int64_t>(d->chunk_rows, M - r0);
        cudaStream_t cs = overlap ? d->side : st;
        if (overlap && c >= 2) AMP_CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_free[b], 0));  // fused(c-2) is done with xhat[b]
        const float *xc = x + r0 * x_stride;
        const bool vec = (x_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(xc) & 7u) == 0);
        const int cast_grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
        int rc = vec ? launch_normalise_cast<true>(d->Kp / BK, cast_grid, cs, xc, x_stride, rows, d->in_features, d->mean, d->denom, d->xhat[b])
                     : launch_normalise_cast<false>(d->Kp / BK, cast_grid, cs, xc, x_stride, rows, d->in_features, d->mean, d->denom, d->xhat[b]);
        if (rc != AMP_OK) return rc;
        AMP_CUDA_TRY(cudaGetLastError());
        if (overlap) AMP_CUDA_TRY(cudaEventRecord(d->ev_ready[b], cs));
        return AMP_OK;
    };
    if (overlap) {  // the side stream starts after everything already queued on the caller's stream (x may still be in flight)
        AMP_CUDA_TRY(cudaEventRecord(d->ev_start, st));
        AMP_CUDA_TRY(cudaStreamWaitEvent(d->side, d->ev_start, 0));
    }
    int rc = issue_cast(0);
    if (rc != AMP_OK) return rc;
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int b = (int)(c & 1);
        const int64_t r0 = c * d->chunk_rows, rows = std::min<int64_t>(d->chunk_rows, M - r0);
        if (c + 1 < n_chunks && (rc = issue_cast(c + 1)) != AMP_OK) return rc;
        if (overlap) AMP_CUDA_TRY(cudaStreamWaitEvent(st, d->ev_ready[b], 0));
        const int m_tiles = (int)((rows + BM - 1) / BM);
        const int grid = std::min(m_tiles, std::min(sms, d->ws_ctas));
        CUtensorMap tm_x;
        rc = make_tmap(&tm_x, d->xhat[b], rows, d->Kp, d->Kp, BM);
        if (rc != AMP_OK) return rc;
        FusedParams fp{};
        fp.M = (int)rows;
        fp.kb1 = d->Kp / BK;
        fp.n1_tiles = d->h1 / BN;
        fp.n2_tiles = d->h2 / BN;
        fp.b1 = d->b1;
        fp.b2 = d->b2;
        fp.w3 = d->w3;
        fp.b3 = d->b3;
        fp.scale = reward_scale;
        fp.reward = reward + r0;
        fp.logits = logits ? logits + r0 : nullptr;
        disc_fused_kernel<<<grid, FUSED_THREADS, FUSED_SMEM_BYTES, st>>>(tm_x, d->tmap_w1, d->tmap_h_load, d->tmap_h_store, d->tmap_w2, fp);
        AMP_CUDA_TRY(cudaGetLastError());
        if (overlap) AMP_CUDA_TRY(cudaEventRecord(d->ev_free[b], st));
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
