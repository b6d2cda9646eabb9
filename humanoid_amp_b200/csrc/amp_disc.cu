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
// Structure: see disc_fused_kernel below -- a persistent kernel (one CTA per SM, 512 threads, warp-specialised: TMA producer,
// single-thread tcgen05.mma issuer, eight read-out warps, six scaler/cast warps).  For large batches of narrow rows the scaler
// and the bf16 cast run INSIDE the kernel: the cast warps read the fp32 rows of the CTA's next row tile straight from the
// caller's buffer, normalise, round to bf16 and park the tile in a per-CTA, L2-resident scratch slot that the TMA producer
// then streams into the operand ring -- ONE launch, no x_hat workspace.  Wide rows (K*A = 830), batches of at most eight row
// tiles per SM and gathered batches run normalise_cast_kernel first (style_reward_impl explains each with its measurement).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "amp_internal.h"
#include "amp_math.cuh"
#include "amp_tc.cuh"

namespace amp {
namespace disc {

using namespace amp::tc;  // PTX wrappers + tensor-map helpers (amp_tc.cuh)

constexpr int BK = kBlockK;  // bf16 per K-block = one 128-byte swizzle row
constexpr int BM = 128;  // rows per tile = TMEM lanes
constexpr int BN = 256;  // accumulator columns per tile = max UMMA N
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int NUM_EPI_THREADS = 256;  // eight epilogue warps
constexpr int TMEM_COLS = 512;
// bf16 staging area for the h1 TMA stores: every epilogue warp owns two 32-row x 32-column slabs (2 KiB each, 64-byte rows
// in the 64B-swizzled layout) -- small enough that a fourth operand stage still fits next to them
constexpr int SLAB_COLS = 32;
constexpr int STORE_SLAB_BYTES = 32 * SLAB_COLS * 2;
constexpr int STORE_STAGING_BYTES = 8 * 2 * STORE_SLAB_BYTES;  // 32 KiB


// Instruction descriptor for kind::f16: D = fp32 (c_format 1 @ bit 4), A = B = bf16 (format 1 @ bits 7, 10), both K-major
// (bits 15, 16 = 0), N >> 3 @ bits [17,23), M >> 4 @ bits [24,29).
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
// same for cta_group::2: M = 256 (128 rows per CTA of the pair)
constexpr uint32_t kInstrDescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

// =====================================================================================================================
// Fused two-layer kernel: layer 1 and layer 2 share ONE persistent launch and keep the tensor pipe busy.
//
// Why not keep h1 on chip: one 128-row tile needs h1 = 128 x 1024 bf16 = 256 KB (> 227 KB smem) for the K loop of layer 2,
// and a layer-2 accumulator of 128 x 512 fp32 alone fills all 512 TMEM columns.  So accumulator tiles are 128 x 256 (two
// TMEM regions) and h1 makes a round trip through a per-CTA, double-buffered workspace that stays in the 126 MB L2
// (148 CTAs x 2 slots x 256 KB = 77.6 MB), written with TMA stores and read back with TMA loads.
//
// The single MMA-issuing thread runs a static software pipeline over the CTA's row tiles (walk_schedule below): the four
// layer-1 tiles of row tile i+1, then the two layer-2 tiles of row tile i; accumulator tiles ping-pong between the two
// TMEM regions so every drain (bias + ReLU + bf16 + swizzled st.shared + TMA store for layer 1; bias + ReLU + dot with w3
// for layer 2) overlaps the MMAs of the next tile.  Per row tile the pipe has 4*12 + 2*64 = 176 MMAs = 22.5 k cycles of
// work; measured (in-kernel cycle counters, AMP_DISC_PROFILE build) ~35 k cycles per tile: ~10 k waiting for operands
// (TMA delivery + shared-memory bandwidth: a 128x256x16 MMA reads 12 KB of smem per 128 cycles while TMA refills the ring
// at the same rate) and ~2 k for accumulator hand-offs.
//
// Warp roles (512 threads): 0 TMA producer, 1 TMEM alloc + MMA issuer, 4..11 epilogue (warps w and w+4 share a TMEM lane
// quarter and split the accumulator columns), 2-3 and 12-15 scaler + bf16 cast of the next row tile.  Producer, issuer and
// epilogue warps all walk the same static schedule.
// =====================================================================================================================
// PAIR = true runs the same pipeline on a CTA pair (cluster of 2, tcgen05 cta_group::2): one MMA covers 256 rows (128 per
// CTA) and each CTA stages only ITS half of the 256-row weight block, so the operand bytes delivered per SM per MMA drop
// from 48 KB to 32 KB per K block.  ncu on the single-CTA version showed TMA loads pinned at 40 % of the xbar->L1 peak on
// every SM (2.57 GB per 151 552 rows, ~51 B/cycle/SM) with the tensor pipe 55 % active: operand delivery, not math, is
// the limiter, and a CTA pair is the only way to shrink it without more TMEM.
// Warp roles by warpgroup (setmaxnreg moves registers between them: the TMEM read-out needs ~168, everything else < 88):
//   WG0  warp 0 TMA producer, warp 1 tcgen05.mma issuer, warps 2-3 scaler/cast    (88 registers)
//   WG1-2 warps 4-11 epilogue                                                      (168 registers)
//   WG3  warps 12-15 scaler/cast                                                   (88 registers)
constexpr int NUM_CONV_WARPS = 6;
constexpr int FUSED_THREADS = 512;
constexpr int EPI_WARP0 = 4;
constexpr int REGS_LIGHT = 88, REGS_EPI = 168;  // 128 * (88 + 168 + 168 + 88) = 65536
constexpr uint32_t ACC_COLS = 256;  // two accumulator regions: TMEM columns [0,256) and [256,512)
__host__ __device__ constexpr int fused_stages(bool pair) { return pair ? 6 : 4; }
__host__ __device__ constexpr int fused_b_bytes(bool pair) { return pair ? B_STAGE_BYTES / 2 : B_STAGE_BYTES; }
__host__ __device__ constexpr int fused_smem_bytes(bool pair) {
    return fused_stages(pair) * (A_STAGE_BYTES + fused_b_bytes(pair)) + STORE_STAGING_BYTES + 256 /*barriers*/ +
           1024 /*partial dots*/ + 1024 /*alignment slack*/;
}

// Schedule shared by producer, issuer and epilogue warps.  Every accumulator tile is told which of the two 256-column
// TMEM regions it uses; issuer and epilogue keep one use counter per region (mbarrier parity = counter & 1).
//     prologue      G1(t0): N1 tiles 0..3, alternating regions
//     iteration i   G1(t[i+1]) tile 0 -> region 0,  8 units of G2(t[i]) -> region 1,  G1 tile 1,  8 units, ...
// i.e. the short layer-1 tiles (12 MMAs = 1.5 k cycles each, ~2 k cycles to drain) are spread between the halves of the two
// long layer-2 tiles (64 MMAs = 8.2 k cycles each): a layer-1 drain has 4 k cycles of layer-2 MMAs to hide under, a
// layer-2 drain has the next layer-1 tile, and h1(t[i]) is complete 5.6 k cycles before its first layer-2 unit needs it.
// (Back-to-back layer-1 tiles -- the previous schedule -- left the issuer waiting ~20 % of its time for drains, in-kernel
// counters of the AMP_DISC_PROFILE build; so did the very first version whose 4-warp drains took longer than 8 units.)
template <class G1, class G2>
__device__ __forceinline__ void walk_schedule(int T, int n1_tiles, int units, G1 &&g1, G2 &&g2) {
    if (T <= 0) return;
    const int seg = (units + n1_tiles - 1) / n1_tiles;
    // i = -1 is the prologue (layer 1 of the first row tile only); one call site per functor keeps the kernel small -- every
    // role inlines its functors here and the roles share the instruction cache
#pragma unroll 1
    for (int i = -1; i < T; ++i) {
        int u = 0;
#pragma unroll 1
        for (int nt = 0; nt < n1_tiles; ++nt) {
            if (i + 1 < T) g1(i + 1, nt, i < 0 ? (nt & 1) : 0);
            if (i >= 0) {
#pragma unroll 1
                for (const int e = min(units, u + seg); u < e; ++u) g2(i, u, 1);
            }
        }
    }
}

struct FusedParams {
    int64_t M;      // rows of this call
    // scaler + cast stage (converter warps): source rows, statistics, the CTA-private bf16 scratch slots
    const float *x;            // (rows, in_features) fp32, row pitch x_stride floats (gathered batches never get here: they
                               // take the cast-kernel path, see style_reward_impl)
    int64_t x_stride;
    const float *mean, *denom; // fp32 [in_features]: (float)running_mean, sqrt((float)running_variance) + 1e-8
    __nv_bfloat16 *xs;         // (gridDim.x * 2 * 128, Kp) bf16: two x_hat row-tile slots per CTA, L2-resident
    int in_features, Kp;
    int x_vec;                 // rows are 8-byte aligned: float2 loads
    int xhat_rows_in_order;    // wide inputs (Kp > 256): tmap_x covers a bf16 x_hat of the whole batch written by
                               // normalise_cast_kernel, the converter warps idle
    int kb1;        // K blocks of layer 1 (Kp / 64)
    int ksteps1_last;  // 16-wide K steps the LAST layer-1 block really needs: ceil((in_features + 2 - 64 (kb1 - 1)) / 16), 1..4
                       // (b1 travels inside the layer-1 product: two padding columns of x_hat hold 1.0, W1 holds b1 there)
    int n1_tiles;   // h1 / 256
    int n2_tiles;   // h2 / 256
    const float *b2, *w3, *b3;
    float scale;
    float *reward;  // [M]
    float *logits;  // [M] or NULL
    long long *prof;  // AMP_DISC_PROFILE builds only: per-CTA cycle counters of the producer / issuer waits, else NULL
    int prof_ctas;    // AMP_DISC_PROFILE builds only: CTAs the counter buffer is laid out for (issuer block, then read-out block)
    int prof_mode;    // AMP_DISC_PROFILE builds only: bit 0 = skip the h1 staging + TMA store (timing experiment, wrong results)
};

#ifdef AMP_DISC_PROFILE
#define AMP_PROF_T0 const long long _t0 = clock64()
#define AMP_PROF_ADD(var) var += clock64() - _t0
#else
#define AMP_PROF_T0
#define AMP_PROF_ADD(var)
#endif

template <bool PAIR>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
disc_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                  const __grid_constant__ CUtensorMap tmap_h_load, const __grid_constant__ CUtensorMap tmap_h_store,
                  const __grid_constant__ CUtensorMap tmap_w2, FusedParams p) {
    constexpr int STAGES = fused_stages(PAIR);
    constexpr int B_BYTES = fused_b_bytes(PAIR);
    constexpr int B_ROWS = PAIR ? BN / 2 : BN;  // weight rows staged by this CTA per K block
    constexpr uint32_t STAGE_TX = PAIR ? 2u * (A_STAGE_BYTES + B_BYTES) : (uint32_t)(A_STAGE_BYTES + B_BYTES);
    // one arrival per epilogue warp (lane 0, after the warp's TMEM reads are fenced and the warp has synchronised): with a
    // CTA pair the peer's arrivals are remote mbarrier operations, and 256 of them per drain serialised for thousands of cycles
    constexpr uint32_t EPI_ARRIVALS = (PAIR ? 2 : 1) * (NUM_EPI_THREADS / 32);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + STAGES * A_STAGE_BYTES;
    const uint32_t staging = smem_b + STAGES * B_BYTES;
    const uint32_t bars = staging + STORE_STAGING_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
    const uint32_t acc_full = bars + 16 * STAGES, acc_empty = acc_full + 16;  // 2 x 8 B each, one per accumulator region
    const uint32_t h1_ready = acc_full + 32;  // 2 x 8 B
    const uint32_t tmem_slot = h1_ready + 16;
    const uint32_t xready = bars + 160, xfree = bars + 176;  // 2 x 8 B each: x_hat scratch slot filled / fully consumed
    const uint32_t part_smem = bars + 256;  // 2 x 4 x 32 floats
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef AMP_DISC_PROFILE
    unsigned long long ns_entry = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_entry));
#endif
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int num_m_tiles = (int)((p.M + BM - 1) / BM);
    // work unit = one 128-row tile per CTA; a pair takes two consecutive tiles (a 256-row block) so both CTAs always walk
    // the same schedule (the second tile of the last block may lie entirely past M: TMA zero-fills, stores are clipped)
    const int group = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
    const int num_groups = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
    const int num_blocks = PAIR ? (num_m_tiles + 1) / 2 : num_m_tiles;
    const int T = group < num_blocks ? (num_blocks - 1 - group) / num_groups + 1 : 0;
    auto tile_of = [&](int ti) { return PAIR ? 2 * (group + ti * num_groups) + (int)rank : group + ti * num_groups; };
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
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, EPI_ARRIVALS);
        }
        mbar_init(h1_ready, 8);  // one arrival per epilogue warp
        mbar_init(h1_ready + 8, 8);
        for (int i = 0; i < 2; ++i) {
            mbar_init(xready + 8 * i, NUM_CONV_WARPS);  // one arrival per converter warp
            mbar_init(xfree + 8 * i, 1);                // the producer, once the slot's last TMA read has landed
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs of a pair allocates (all 512 columns) and later frees
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
#ifdef AMP_DISC_PROFILE
    if (threadIdx.x == 0 && p.prof) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        long long *e = p.prof + (size_t)p.prof_ctas * 8 + blockIdx.x * 8;
        e[7] = (long long)(ns - ns_entry);  // entry -> setup done (barrier init, TMEM alloc, first sync)
        p.prof[blockIdx.x * 8 + 7] = (long long)ns_entry;
    }
#endif
    const bool is_epilogue_warp = warp >= EPI_WARP0 && warp < EPI_WARP0 + NUM_EPI_THREADS / 32;
    // accumulator-drained barriers live in the leader CTA; the epilogue threads of both CTAs arrive there
    const uint32_t acc_empty_at_leader = PAIR ? mapa_rank(acc_empty, 0) : acc_empty;  // + 8 * region
    auto arrive_drained = [&](uint32_t addr) {
        if constexpr (PAIR) mbar_arrive_cluster(addr); else mbar_arrive(addr);
    };

    // setmaxnreg sits at the top of each warpgroup's own branch so that ptxas budgets the code it dominates accordingly
    if (!is_epilogue_warp) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_LIGHT));
    if (is_epilogue_warp) {
        // handled in the last branch below
    } else if (warp == 0) {
        // ================= TMA producer (every CTA loads its own rows of A and its own part of the weight block) =========
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] long long w_empty = 0, w_h1 = 0;
            // weights and h1 slots: evict_last.  x_hat: normal priority while its row tile is being re-read (a 48 KB tile comes
            // back within a few microseconds), evict_first on its last pass -- marking it evict_last made 192 MB of x_hat
            // per launch compete with the 77.6 MB of h1 slots for the protected part of L2.
            const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first(), normal = l2_policy_evict_normal();
            // x_hat scratch slots: slot s is handed back to the converter warps once every TMA read of it has landed, which
            // is known STAGES ring acquisitions after its last load was issued (the MMA that freed that stage had waited
            // for the load, and for every load before it)
            long long loads_issued = 0, free_at = 0;
            int free_slot = -1;
            auto load_pair = [&](const CUtensorMap *ma, int a_col, int a_row, uint64_t a_policy, const CUtensorMap *mb, int b_col,
                                 int b_row) {
                {
                    AMP_PROF_T0;
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    AMP_PROF_ADD(w_empty);
                }
                const uint32_t fb = full_bar + 8 * stage;
#ifdef AMP_DISC_PROFILE
                if (p.prof_mode & 16) {  // timing experiment: no TMA loads, the MMAs run on whatever the stage holds
                    if (!PAIR || leader) mbar_arrive(fb);
                } else
#endif
                if constexpr (PAIR) {
                    if (leader) mbar_arrive_expect_tx(fb, STAGE_TX);  // bytes of both CTAs are credited to the leader's barrier
                    const uint32_t fb_leader = mapa_rank(fb, 0);
                    tma_load_2d_pair_hint(smem_a + stage * A_STAGE_BYTES, ma, a_col, a_row, fb_leader, a_policy);
                    tma_load_2d_pair_hint(smem_b + stage * B_BYTES, mb, b_col, b_row + (int)rank * B_ROWS, fb_leader, keep);
                } else {
                    mbar_arrive_expect_tx(fb, STAGE_TX);
                    tma_load_2d_hint(smem_a + stage * A_STAGE_BYTES, ma, a_col, a_row, fb, a_policy);
                    tma_load_2d_hint(smem_b + stage * B_BYTES, mb, b_col, b_row, fb, keep);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                ++loads_issued;
                if (free_slot >= 0 && loads_issued >= free_at) {
                    mbar_arrive(xfree + 8 * free_slot);
                    free_slot = -1;
                }
            };
            walk_schedule(
                T, p.n1_tiles, units,
                [&](int ti, int nt, int) {
                    if (p.xhat_rows_in_order) {
                        // x_hat of the whole batch was prepared by the cast kernel (wide inputs): plain row-tile reads, streamed
#pragma unroll 1
                        for (int kb = 0; kb < p.kb1; ++kb)
                            load_pair(&tmap_x, kb * BK, tile_of(ti) * BM, nt == p.n1_tiles - 1 ? stream : normal, &tmap_w1, kb * BK, nt * BN);
                        return;
                    }
                    if (nt == 0) {  // the converter warps have parked x_hat of this row tile in slot ti & 1
                        AMP_PROF_T0;
                        mbar_wait(xready + 8 * (ti & 1), (uint32_t)((ti >> 1) & 1));
                        AMP_PROF_ADD(w_h1);
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                    // the slot is re-read by the four N1 tiles of this row tile (keep it in L2) and dead after the last one
#pragma unroll 1
                    for (int kb = 0; kb < p.kb1; ++kb)
                        load_pair(&tmap_x, kb * BK, slot_row0 + (ti & 1) * BM, nt == p.n1_tiles - 1 ? stream : keep, &tmap_w1, kb * BK, nt * BN);
                    if (nt == p.n1_tiles - 1) {
                        free_slot = ti & 1;
                        free_at = loads_issued + STAGES;
                    }
                },
                [&](int ti, int u, int) {
                    const int n2 = u / kb2, kb = u - n2 * kb2;
                    if (u == 0) {  // h1 of this row tile has been written (all four E1 warps' TMA stores completed)
                        AMP_PROF_T0;
                        mbar_wait(h1_ready + 8 * (ti & 1), (uint32_t)((ti >> 1) & 1));
                        AMP_PROF_ADD(w_h1);
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                    // h1 is dead after its last read (the second N2 tile): let L2 evict those lines before live ones
                    load_pair(&tmap_h_load, kb * BK, slot_row0 + (ti & 1) * BM, n2 == p.n2_tiles - 1 ? stream : keep, &tmap_w2, kb * BK, n2 * BN);
                });
#ifdef AMP_DISC_PROFILE
            if (p.prof) {
                p.prof[blockIdx.x * 8 + 4] = w_empty;
                p.prof[blockIdx.x * 8 + 5] = w_h1;
            }
#endif
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer (one thread; for a pair, one thread of the leader CTA drives both SMs) =============
        if (lane == 0 && leader) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] long long w_full = 0, w_d1 = 0, w_d2 = 0;
#ifdef AMP_DISC_PROFILE
            const long long t_begin = clock64();
            unsigned long long ns_begin;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
#endif
            auto commit = [&](uint32_t bar) {
                if constexpr (PAIR) umma_commit_pair(bar, (uint16_t)0x3); else umma_commit(bar);
            };
            // ksteps < 4 only for the last layer-1 block: x_hat is zero-padded to Kp, so the trailing all-zero K steps add
            // exact zeros to the accumulator and can be skipped (166 inputs: 11 MMAs per layer-1 tile instead of 12)
            auto mma_block = [&](uint32_t d_tmem, bool first, int ksteps) {
                {
                    AMP_PROF_T0;
                    mbar_wait(full_bar + 8 * stage, phase);
                    AMP_PROF_ADD(w_full);
                }
                tcgen05_fence_after();
                const uint64_t a0 = make_kmajor_sw128_desc(smem_a + stage * A_STAGE_BYTES);
                const uint64_t b0 = make_kmajor_sw128_desc(smem_b + stage * B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    if (k >= ksteps) break;
#ifdef AMP_DISC_PROFILE
                    if (p.prof_mode & 8) break;  // timing experiment: operands are delivered but no MMA is issued
#endif
                    if constexpr (PAIR) umma_bf16_pair(d_tmem, a0 + 2 * k, b0 + 2 * k, kInstrDescPair, (uint32_t)(!first || k != 0));
                    else umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, kInstrDesc, (uint32_t)(!first || k != 0));
                }
                commit(empty_bar + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            };
            uint32_t uses = 0;  // bit r = parity of the number of accumulator tiles issued so far into TMEM region r
            auto acquire_acc = [&](int r, long long &w) {
                AMP_PROF_T0;
                mbar_wait(acc_empty + 8 * r, ((uses >> r) & 1) ^ 1);  // the epilogue (of both CTAs) has drained its last tile
                AMP_PROF_ADD(w);
                tcgen05_fence_after();
            };
            walk_schedule(
                T, p.n1_tiles, units,
                [&](int, int, int r) {
                    acquire_acc(r, w_d1);
#pragma unroll 1
                    for (int kb = 0; kb < p.kb1; ++kb)
                        mma_block(tmem_base + r * ACC_COLS, kb == 0, kb == p.kb1 - 1 ? p.ksteps1_last : BK / UMMA_K);
                    commit(acc_full + 8 * r);
                    uses ^= 1u << r;
                },
                [&](int, int u, int r) {
                    const int kb = u % kb2;
                    if (kb == 0) acquire_acc(r, w_d2);
                    mma_block(tmem_base + r * ACC_COLS, kb == 0, BK / UMMA_K);
                    if (kb == kb2 - 1) {
                        commit(acc_full + 8 * r);
                        uses ^= 1u << r;
                    }
                });
#ifdef AMP_DISC_PROFILE
            if (p.prof) {
                p.prof[blockIdx.x * 8 + 0] = clock64() - t_begin;
                p.prof[blockIdx.x * 8 + 1] = w_full;
                p.prof[blockIdx.x * 8 + 2] = w_d1;
                p.prof[blockIdx.x * 8 + 3] = w_d2;
                unsigned long long ns_end;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
                p.prof[blockIdx.x * 8 + 6] = (long long)(ns_end - ns_begin);
            }
#endif
        }
        __syncwarp();
    } else {
        // ================= converter warps 2, 3, 12..15: RunningStandardScaler (eval form) + bf16 cast of the CTA's row tiles =====
        //   x_hat = clamp((x - mean) * (1 / denom), -5, 5) -> bf16, zero padded to Kp (ones in the two bias columns)
        // One pass per K-block (64 columns): lane l owns the column pair (2l, 2l+1) of the block, so a load is a coalesced
        // 8-byte-per-lane row segment, a store a coalesced 4-byte-per-lane one, and the block's statistics are four
        // registers.  Eight rows are in flight per warp.  (x - mean) / denom is evaluated with an IEEE reciprocal: at most
        // 1 ulp (fp32) from the division and rounded to bf16 right after.  Rows past M are written as zeros.  Tile ti goes to
        // slot ti & 1 of this CTA's scratch (two slots: the tile the layer-1 MMAs are reading and the one being prepared);
        // the slot comes back through xfree.  The loop nest is deliberately NOT unrolled beyond the eight rows: this code
        // shares the instruction cache with the TMEM read-out (ncu: no_instruction was the top stall of both).
        const int cw = warp < EPI_WARP0 ? warp - 2 : warp - (EPI_WARP0 + NUM_EPI_THREADS / 32) + 2;
        constexpr int RIF = 8;  // rows in flight per warp
        __nv_bfloat16 *const my_slots = p.xs + (size_t)slot_row0 * p.Kp;
        for (int ti = 0; ti < (p.xhat_rows_in_order ? 0 : T); ++ti) {
            const int s = ti & 1;
            if (ti >= 2) mbar_wait(xfree + 8 * s, (uint32_t)(((ti >> 1) - 1) & 1));
            const int64_t row_base = (int64_t)tile_of(ti) * BM;
            __nv_bfloat16 *const dst = my_slots + (size_t)s * BM * p.Kp;
            // (A bulk L2 prefetch of the next tile's fp32 rows was measured here: it shortens the converter's load latency but
            // costs the fused kernel 3.4 % -- the extra L2 fills evict h1 / weight lines -- and the converter is not critical.)
#ifdef AMP_DISC_PROFILE
            if (p.prof_mode & 4) {  // timing experiment: no conversion work at all (wrong results)
                __syncwarp();
                if (lane == 0) mbar_arrive(xready + 8 * s);
                continue;
            }
#endif
#pragma unroll 1
            for (int kb = 0; kb < p.kb1; ++kb) {
                const int c = 2 * (lane + 32 * kb);
                const bool in0 = c < p.in_features, in1 = c + 1 < p.in_features;
                const bool wide = p.x_vec && in1;
                float2 mu = make_float2(in0 ? __ldg(p.mean + c) : 0.0f, in1 ? __ldg(p.mean + c + 1) : 0.0f);
                float2 rc = make_float2(in0 ? __frcp_rn(__ldg(p.denom + c)) : 0.0f, in1 ? __frcp_rn(__ldg(p.denom + c + 1)) : 0.0f);
                // bias columns: x is read as 0 there, (0 - (-1)) * 1 = 1
                if (!in0 && c < p.in_features + 2) { mu.x = -1.0f; rc.x = 1.0f; }
                if (!in1 && c + 1 < p.in_features + 2) { mu.y = -1.0f; rc.y = 1.0f; }
                auto source_row = [&](int64_t grow) -> const float * {  // first float of this lane's pair in batch row grow
                    return p.x + grow * p.x_stride + c;
                };
                auto emit = [&](int r, float2 x, bool live) {
                    const float a = clamp_nan(__fmul_rn(__fsub_rn(x.x, mu.x), rc.x), -5.0f, 5.0f);
                    const float b = clamp_nan(__fmul_rn(__fsub_rn(x.y, mu.y), rc.y), -5.0f, 5.0f);
                    reinterpret_cast<__nv_bfloat162 *>(dst + (size_t)r * p.Kp)[lane + 32 * kb] =
                        live ? __floats2bfloat162_rn(a, b) : __floats2bfloat162_rn(0.0f, 0.0f);
                };
                if (wide || !in0) {  // 8-byte loads (or a pure padding pair: nothing to load), eight rows in flight
#pragma unroll 1
                    for (int r0 = cw; r0 < BM; r0 += NUM_CONV_WARPS * RIF) {
                        float2 v[RIF];
#pragma unroll
                        for (int q = 0; q < RIF; ++q) {
                            const int r = r0 + q * NUM_CONV_WARPS;
                            const int64_t grow = row_base + r;
                            v[q] = make_float2(0.0f, 0.0f);
                            if (wide && r < BM && grow < p.M) v[q] = __ldcs(reinterpret_cast<const float2 *>(source_row(grow)));
                        }
#pragma unroll
                        for (int q = 0; q < RIF; ++q) {
                            const int r = r0 + q * NUM_CONV_WARPS;
                            if (r < BM) emit(r, v[q], row_base + r < p.M);
                        }
                    }
                } else {  // odd pitch / unaligned base / the last column of an odd width: 4-byte loads, one row at a time
#pragma unroll 1
                    for (int r = cw; r < BM; r += NUM_CONV_WARPS) {
                        const int64_t grow = row_base + r;
                        float2 x = make_float2(0.0f, 0.0f);
                        if (grow < p.M) {
                            const float *xr = source_row(grow);
                            x.x = __ldcs(xr);
                            if (in1) x.y = __ldcs(xr + 1);
                        }
                        emit(r, x, grow < p.M);
                    }
                }
            }
            asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy global writes -> visible to the TMA loads
            __syncwarp();
            if (lane == 0) mbar_arrive(xready + 8 * s);
        }
    }
    if (is_epilogue_warp) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
        // ================= epilogue warps 4..11: both accumulators, columns split between the two warps of a TMEM quarter ======
        // Warp w and warp w+4 own the same 32 TMEM lanes (rows); w takes columns [0,128) of every 256-column accumulator and
        // w+4 takes [128,256), so each drain is half as long as with one warp per quarter.  The warps follow the issuer's
        // schedule: D1 tiles (ReLU -> bf16 -> swizzled slab -> TMA store of h1; the bias rides inside the product) and, after
        // the last K block of each layer-2 N tile, D2 (bias + ReLU -> dot with w3).  tcgen05.ld of the next 32 columns is in
        // flight while the current 32 are processed.  Both read-out loops are unrolled by two only (the double-buffered TMEM
        // registers need static indices): fully unrolled they were a third of a 113 KB kernel and instruction-cache misses
        // (ncu: stall_no_inst) were the top stall reason of these warps.
        const int ew = warp - EPI_WARP0;
        const int quarter = warp & 3;
        const int colhalf = ew >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t col_base = (uint32_t)(colhalf * (BN / 2));
        float *part = reinterpret_cast<float *>(smem_raw + (part_smem - smem_u32(smem_raw)));  // [2][4][32] partial dots
        uint32_t uses = 0, store_it = 0;  // uses bit r: parity of the tiles drained so far from region r (as the issuer counts)
        const uint64_t h1_keep = l2_policy_evict_last();
        float dot[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        [[maybe_unused]] long long e_wait1 = 0, e_drain1 = 0, e_post1 = 0, e_wait2 = 0, e_drain2 = 0, e_post2 = 0, e_slab = 0;
        walk_schedule(
            T, p.n1_tiles, units,
            [&](int ti, int nt, int r) {
                // ---- D1 (ti, nt): this warp's 128 columns = four 32-column slabs ----
                const int row0 = slot_row0 + (ti & 1) * BM + quarter * 32;
                const uint32_t acc = tmem_base + lane_base + r * ACC_COLS + col_base;
#ifdef AMP_DISC_PROFILE
                const long long e_t0 = clock64();
                long long e_t2 = 0;
#endif
                mbar_wait(acc_full + 8 * r, (uses >> r) & 1);
#ifdef AMP_DISC_PROFILE
                const long long e_t1 = clock64();
#endif
                tcgen05_fence_after();
                uint32_t v[2][32];
                tmem_ld_32x32(acc, v[0]);
#pragma unroll 1
                for (int hp = 0; hp < 4; hp += 2) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {  // four steps of 32 columns = four slabs
                        const int h = hp + hh;
                        const int col = nt * BN + (int)col_base + h * 32;  // h1 column of this step
                        const uint32_t slab = staging + (uint32_t)((ew * 2 + (store_it & 1)) * STORE_SLAB_BYTES);
                        {
                            AMP_PROF_T0;
                            if (lane == 0) bulk_wait_read<1>();  // the store that last read this slab has drained it
                            __syncwarp();
                            AMP_PROF_ADD(e_slab);
                        }
                        tmem_ld_wait();  // v[hh] has landed
                        if (h + 1 < 4) {
                            tmem_ld_32x32(acc + (uint32_t)((h + 1) * 32), v[hh ^ 1]);
                        } else {  // last TMEM read of this accumulator by this warp: release the region to the issuer now
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) arrive_drained(acc_empty_at_leader + 8 * r);
#ifdef AMP_DISC_PROFILE
                            e_t2 = clock64();
#endif
                        }
                        const uint32_t(&cur)[32] = v[hh];
#ifdef AMP_DISC_PROFILE
                        if (p.prof_mode & 1) { ++store_it; continue; }
#endif
                        // the accumulator already holds x W1^T + b1 (the bias rides in two padding columns): round, then ReLU
                        // on the packed pair (rounding is monotonic and keeps the sign, so max(round(x), 0) == round(max(x, 0)));
                        // the NaN-propagating max: torch's relu(NaN) is NaN
                        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // 16-byte chunk j of this thread's 64-byte slab row
                            const __nv_bfloat162 p0 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 0]), __uint_as_float(cur[8 * j + 1])), zero2);
                            const __nv_bfloat162 p1 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 2]), __uint_as_float(cur[8 * j + 3])), zero2);
                            const __nv_bfloat162 p2 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 4]), __uint_as_float(cur[8 * j + 5])), zero2);
                            const __nv_bfloat162 p3 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 6]), __uint_as_float(cur[8 * j + 7])), zero2);
                            // SWIZZLE_64B: 16-byte chunk index XOR address bits [7,9) = (row >> 1) & 3 (rows are 64 bytes)
                            const int chunk = j ^ ((lane >> 1) & 3);
                            st_shared_v4(slab + (uint32_t)(lane * 64 + chunk * 16), *reinterpret_cast<const uint32_t *>(&p0),
                                         *reinterpret_cast<const uint32_t *>(&p1), *reinterpret_cast<const uint32_t *>(&p2),
                                         *reinterpret_cast<const uint32_t *>(&p3));
                        }
                        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d_hint(&tmap_h_store, slab, col, row0, h1_keep);  // the slot is overwritten in place: keep it in L2
                            bulk_commit();
                        }
                        ++store_it;
                    }
                }
#ifdef AMP_DISC_PROFILE
                e_wait1 += e_t1 - e_t0;
                e_drain1 += e_t2 - e_t1;
                e_post1 += clock64() - e_t2;
#endif
                uses ^= 1u << r;
                if (nt == p.n1_tiles - 1 && lane == 0) {  // this warp's part of h1(ti) is complete in the workspace
                    bulk_wait_all();
                    mbar_arrive(h1_ready + 8 * (ti & 1));
                }
            },
            [&](int ti, int u, int r) {
                const int n2 = u / kb2;
                if (u - n2 * kb2 != kb2 - 1) return;  // D2 (ti, n2) is complete after the last K block of the N tile
                // ---- D2 (ti, n2): this warp's 128 columns ----
                const uint32_t acc = tmem_base + lane_base + r * ACC_COLS + col_base;
#ifdef AMP_DISC_PROFILE
                const long long e_t0 = clock64();
                long long e_t2 = 0;
#endif
                mbar_wait(acc_full + 8 * r, (uses >> r) & 1);
#ifdef AMP_DISC_PROFILE
                const long long e_t1 = clock64();
#endif
                tcgen05_fence_after();
                uint32_t v[2][32];
                tmem_ld_32x32(acc, v[0]);
#pragma unroll 1
                for (int cp = 0; cp < 4; cp += 2) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int chunk = cp + cc;
                        const int col0 = n2 * BN + (int)col_base + chunk * 32;
                        const float4 *bias4 = reinterpret_cast<const float4 *>(p.b2 + col0);
                        const float4 *w4 = reinterpret_cast<const float4 *>(p.w3 + col0);
                        float4 bb[4], ww[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            bb[j] = __ldg(bias4 + j);
                            ww[j] = __ldg(w4 + j);
                        }
                        tmem_ld_wait();  // v[cc] has landed
                        if (chunk + 1 < 4) {
                            tmem_ld_32x32(acc + (uint32_t)((chunk + 1) * 32), v[cc ^ 1]);
                        } else {  // last TMEM read: hand the region back before finishing the arithmetic
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) arrive_drained(acc_empty_at_leader + 8 * r);
#ifdef AMP_DISC_PROFILE
                            e_t2 = clock64();
#endif
                        }
                        const uint32_t(&cur)[32] = v[cc];
#pragma unroll
                        for (int half = 0; half < 2; ++half) {  // 16 columns at a time: the second half's constants load under the first
                            if (half == 1) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    bb[j] = __ldg(bias4 + 4 + j);
                                    ww[j] = __ldg(w4 + 4 + j);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int e = 16 * half + 4 * j;
                                dot[0] = fmaf(max_nan(__uint_as_float(cur[e + 0]) + bb[j].x, 0.0f), ww[j].x, dot[0]);
                                dot[1] = fmaf(max_nan(__uint_as_float(cur[e + 1]) + bb[j].y, 0.0f), ww[j].y, dot[1]);
                                dot[2] = fmaf(max_nan(__uint_as_float(cur[e + 2]) + bb[j].z, 0.0f), ww[j].z, dot[2]);
                                dot[3] = fmaf(max_nan(__uint_as_float(cur[e + 3]) + bb[j].w, 0.0f), ww[j].w, dot[3]);
                            }
                        }
                    }
                }
#ifdef AMP_DISC_PROFILE
                e_wait2 += e_t1 - e_t0;
                e_drain2 += e_t2 - e_t1;
                e_post2 += clock64() - e_t2;
#endif
                uses ^= 1u << r;
                if (n2 == p.n2_tiles - 1) {
                    // combine the two column halves of this row: the upper-half warp hands its partial sum over through
                    // shared memory (double-buffered by tile parity), a 64-thread named barrier orders the exchange
                    const float mine = (dot[0] + dot[1]) + (dot[2] + dot[3]);
                    float *slot = part + ((ti & 1) * 4 + quarter) * 32 + lane;
                    if (colhalf == 1) *slot = mine;
                    asm volatile("bar.sync %0, 64;" ::"r"(8 + quarter) : "memory");
                    if (colhalf == 0) {
                        const int64_t row = (int64_t)tile_of(ti) * BM + quarter * 32 + lane;
                        if (row < p.M) {
                            const float logit = (mine + *slot) + __ldg(p.b3);
                            if (p.logits) p.logits[row] = logit;
                            p.reward[row] = style_reward(logit, p.scale);
                        }
                    }
                    dot[0] = dot[1] = dot[2] = dot[3] = 0.0f;
                }
            });
#ifdef AMP_DISC_PROFILE
        if (p.prof && ew == 0 && lane == 0) {
            long long *e = p.prof + (size_t)p.prof_ctas * 8 + blockIdx.x * 8;
            e[0] = e_wait1; e[1] = e_drain1; e[2] = e_post1; e[3] = e_wait2; e[4] = e_drain2; e[5] = e_post2; e[6] = e_slab;
        }
#endif
    }

    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
#ifdef AMP_DISC_PROFILE
    if (threadIdx.x == 0 && p.prof) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.prof[blockIdx.x * 8 + 7] = (long long)ns - p.prof[blockIdx.x * 8 + 7];  // entry -> all roles done
    }
#endif
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Small batches (at most one row tile per two SMs, e.g. the 4096 environments of the reference configs): ONE launch, a cluster
// of two CTAs per 128-row tile.  With one tile per CTA nothing overlaps -- the fused kernel above spends its time in a serial
// chain of 172 MMAs (16 us) behind a separate cast launch.  Here the chain is split over the pair and the cast is folded in:
//
//   both CTAs    all warps: scaler + bf16 cast of the tile straight into shared memory in the UMMA operand layout (resident
//                A operand of layer 1; the two CTAs convert the same rows, L2 serves the second one)
//   CTA r        layer 1, N tiles 2r and 2r+1 (22 MMAs)  ->  ReLU -> bf16 -> TMA store into the tile's h1 scratch rows
//   cluster barrier (h1 complete in L2)
//   CTA r        layer 2, N tile r (64 MMAs), h1 by TMA
//   CTA 0        relu(acc + b2) . w3 over ITS 256 columns, the four partial sums of every thread go to the same thread of CTA 1
//   cluster barrier
//   CTA 1        continues the SAME accumulation chains with its columns, combines the two column halves, emits logit + reward
//
// Every arithmetic step and its order is the one of disc_fused_kernel (same x_hat, same MMAs per accumulator, the layer-3 dot is
// the fused kernel's per-thread chain handed from CTA 0 to CTA 1), so results are bit-identical with the large-batch paths.
// 320 threads: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..9 read-out (as above).
// =====================================================================================================================
constexpr int SMALL_THREADS = 320;
constexpr int SMALL_STAGES = 3;
constexpr int SMALL_STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // 48 KiB
constexpr int SMALL_STAGING_BYTES = 8 * STORE_SLAB_BYTES;        // one 32 x 32 slab per read-out warp
__host__ __device__ constexpr int small_smem_bytes(int kb1) {
    return kb1 * A_STAGE_BYTES + SMALL_STAGES * SMALL_STAGE_BYTES + SMALL_STAGING_BYTES + 256 /*barriers*/ + 4096 /*chain hand-off*/ +
           1024 /*partial dots*/ + 1024 /*alignment slack*/;
}

struct SmallParams {
    int64_t M;
    const float *x;
    int64_t x_stride;
    const int64_t *row_index;  // NULL or gathered rows (x holds `capacity` rows)
    int64_t capacity;
    uint32_t *flags;
    const float *mean, *denom;
    int in_features, kb1, ksteps1_last, x_vec;
    int n1_tiles, n2_tiles;  // 4, 2 (the kernel requires h1 = 1024, h2 = 512: two N tiles of each layer per CTA / one per CTA)
    const float *b2, *w3, *b3;
    float scale;
    float *reward, *logits;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SMALL_THREADS, 1)
disc_small_kernel(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_h_load,
                  const __grid_constant__ CUtensorMap tmap_h_store, const __grid_constant__ CUtensorMap tmap_w2, SmallParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a1 = base;                                    // kb1 K-blocks of the resident x_hat tile
    const uint32_t ring = smem_a1 + p.kb1 * A_STAGE_BYTES;            // stage: A 16 KiB | B 32 KiB
    const uint32_t staging = ring + SMALL_STAGES * SMALL_STAGE_BYTES;
    const uint32_t bars = staging + SMALL_STAGING_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * SMALL_STAGES;
    const uint32_t acc_full = bars + 64, acc_empty = bars + 80;        // 2 x 8 B each (TMEM regions 0 / 1)
    const uint32_t tmem_slot = bars + 96;
    const uint32_t chain_smem = bars + 256;                            // 256 threads x 4 floats from CTA 0
    const uint32_t part_smem = chain_smem + 4096;                      // 4 x 32 floats
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float *chain_in = reinterpret_cast<float *>(smem_raw + (chain_smem - smem_u32(smem_raw)));
    float *part = reinterpret_cast<float *>(smem_raw + (part_smem - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int tile = (int)blockIdx.x >> 1;
    const int64_t row_base = (int64_t)tile * BM;
    const int h1_row0 = tile * BM;  // this tile's rows in the h1 scratch
    const int kb2 = 4 * p.n1_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_h_load) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_h_store) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w2) : "memory");
        for (int i = 0; i < SMALL_STAGES; ++i) {
            mbar_init(full_bar + 8 * i, 1);
            mbar_init(empty_bar + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, NUM_EPI_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- scaler + bf16 cast of the tile, straight into the 128B-swizzled K-major operand layout (all ten warps) -----------
    // row r of K-block kb at kb * 16 KiB + r * 128 B, 16-byte chunk index XOR (r & 7); lane l owns the column pair (2l, 2l+1)
    // of every K-block.  All K-blocks of seven rows are in flight per warp (the tile is converted in two load rounds: every
    // round is a DRAM round trip on the critical path of a kernel that is nothing but latency).
    {
        constexpr int RIF = 7, KB = 3;  // small_ok: kb1 <= 3
        float2 mu[KB], rc[KB];  // rc holds the denominators until the first round of row loads is in flight (issue is in order:
        bool in0[KB], in1[KB];  // taking the reciprocals here would hold the row loads back by one DRAM round trip)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const int c = 2 * (lane + 32 * kb);
            in0[kb] = kb < p.kb1 && c < p.in_features;
            in1[kb] = kb < p.kb1 && c + 1 < p.in_features;
            mu[kb] = make_float2(in0[kb] ? __ldg(p.mean + c) : 0.0f, in1[kb] ? __ldg(p.mean + c + 1) : 0.0f);
            rc[kb] = make_float2(in0[kb] ? __ldg(p.denom + c) : 0.0f, in1[kb] ? __ldg(p.denom + c + 1) : 0.0f);
        }
        bool have_rcp = false;
#pragma unroll 1
        for (int r0 = warp; r0 < BM; r0 += (SMALL_THREADS / 32) * RIF) {
            float2 v[RIF][KB];
#pragma unroll
            for (int q = 0; q < RIF; ++q) {
                const int r = r0 + q * (SMALL_THREADS / 32);
                const int64_t grow = row_base + r;
                const bool live = r < BM && grow < p.M;
                int64_t sr = grow;
                if (live && p.row_index) {
                    sr = __ldg(p.row_index + grow);
                    if (sr < 0 || sr >= p.capacity) {
                        if (lane == 0 && p.flags) atomicOr(p.flags, 2u);
                        sr = 0;
                    }
                }
                const float *xr = p.x + sr * p.x_stride + 2 * lane;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    v[q][kb] = make_float2(0.0f, 0.0f);
                    if (live && in0[kb]) {
                        if (p.x_vec && in1[kb]) {
                            v[q][kb] = __ldcs(reinterpret_cast<const float2 *>(xr + 64 * kb));
                        } else {
                            v[q][kb].x = __ldcs(xr + 64 * kb);
                            if (in1[kb]) v[q][kb].y = __ldcs(xr + 64 * kb + 1);
                        }
                    }
                }
            }
            if (!have_rcp) {
                have_rcp = true;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    const int c = 2 * (lane + 32 * kb);
                    rc[kb] = make_float2(in0[kb] ? __frcp_rn(rc[kb].x) : 0.0f, in1[kb] ? __frcp_rn(rc[kb].y) : 0.0f);
                    if (!in0[kb] && c < p.in_features + 2) { mu[kb].x = -1.0f; rc[kb].x = 1.0f; }  // bias columns: (0 - (-1)) * 1 = 1
                    if (!in1[kb] && c + 1 < p.in_features + 2) { mu[kb].y = -1.0f; rc[kb].y = 1.0f; }
                }
            }
#pragma unroll
            for (int q = 0; q < RIF; ++q) {
                const int r = r0 + q * (SMALL_THREADS / 32);
                if (r >= BM) continue;
                const bool live = row_base + r < p.M;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    if (kb >= p.kb1) continue;
                    const float a = clamp_nan(__fmul_rn(__fsub_rn(v[q][kb].x, mu[kb].x), rc[kb].x), -5.0f, 5.0f);
                    const float b = clamp_nan(__fmul_rn(__fsub_rn(v[q][kb].y, mu[kb].y), rc[kb].y), -5.0f, 5.0f);
                    const __nv_bfloat162 o = live ? __floats2bfloat162_rn(a, b) : __floats2bfloat162_rn(0.0f, 0.0f);
                    const uint32_t addr = smem_a1 + (uint32_t)(kb * A_STAGE_BYTES + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + ((lane & 3) << 2));
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t *>(&o)) : "memory");
                }
            }
        }
        fence_proxy_async_smem();  // generic-proxy shared writes -> visible to the tensor core (async proxy)
    }
    tcgen05_fence_before();
    __syncthreads();  // barriers initialised, TMEM allocated, x_hat in place (the peer is first touched after cluster barrier #1)
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer ==================================================================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
            auto acquire = [&]() {
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                return full_bar + 8 * stage;
            };
            auto advance = [&]() {
                if (++stage == SMALL_STAGES) { stage = 0; phase ^= 1; }
            };
            for (int j = 0; j < 2; ++j) {  // layer 1: only the weight block moves, x_hat is resident
                const int nt = 2 * (int)rank + j;
                for (int kb = 0; kb < p.kb1; ++kb) {
                    const uint32_t fb = acquire();
                    mbar_arrive_expect_tx(fb, (uint32_t)B_STAGE_BYTES);
                    tma_load_2d_hint(ring + stage * SMALL_STAGE_BYTES + A_STAGE_BYTES, &tmap_w1, kb * BK, nt * BN, fb, keep);
                    advance();
                }
            }
            // layer 2, N tile `rank`: the weight blocks of the first stages do not depend on h1 and are requested before the
            // cluster barrier; the h1 blocks follow once both CTAs have stored their halves
            int b_issued = 0;
            int b_stage[SMALL_STAGES];
            uint32_t b_full[SMALL_STAGES];
            for (; b_issued < SMALL_STAGES && b_issued < kb2; ++b_issued) {
                const uint32_t fb = acquire();
                mbar_arrive_expect_tx(fb, (uint32_t)SMALL_STAGE_BYTES);
                tma_load_2d_hint(ring + stage * SMALL_STAGE_BYTES + A_STAGE_BYTES, &tmap_w2, b_issued * BK, (int)rank * BN, fb, keep);
                b_stage[b_issued] = stage;
                b_full[b_issued] = fb;
                advance();
            }
            cluster_sync_all();  // #1: h1 of the tile is complete (both CTAs' TMA stores have been waited for)
            asm volatile("fence.proxy.async.global;" ::: "memory");
            for (int kb = 0; kb < b_issued; ++kb)
                tma_load_2d_hint(ring + b_stage[kb] * SMALL_STAGE_BYTES, &tmap_h_load, kb * BK, h1_row0, b_full[kb], rank == 1 ? stream : keep);
            for (int kb = b_issued; kb < kb2; ++kb) {
                const uint32_t fb = acquire();
                mbar_arrive_expect_tx(fb, (uint32_t)SMALL_STAGE_BYTES);
                tma_load_2d_hint(ring + stage * SMALL_STAGE_BYTES, &tmap_h_load, kb * BK, h1_row0, fb, stream);
                tma_load_2d_hint(ring + stage * SMALL_STAGE_BYTES + A_STAGE_BYTES, &tmap_w2, kb * BK, (int)rank * BN, fb, keep);
                advance();
            }
            cluster_sync_all();  // #2
        } else {  // the idle lanes only take part in the two cluster barriers (they must not wait for lane 0 in between)
            cluster_sync_all();  // #1
            cluster_sync_all();  // #2
        }
    } else if (warp == 1) {
        // ================= MMA issuer =====================================================================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&]() {
                if (++stage == SMALL_STAGES) { stage = 0; phase ^= 1; }
            };
            for (int j = 0; j < 2; ++j) {  // layer 1 into TMEM region j
                for (int kb = 0; kb < p.kb1; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tcgen05_fence_after();
                    const uint64_t a0 = make_kmajor_sw128_desc(smem_a1 + kb * A_STAGE_BYTES);
                    const uint64_t b0 = make_kmajor_sw128_desc(ring + stage * SMALL_STAGE_BYTES + A_STAGE_BYTES);
                    const int ksteps = kb == p.kb1 - 1 ? p.ksteps1_last : BK / UMMA_K;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (k >= ksteps) break;
                        umma_bf16(tmem_base + j * ACC_COLS, a0 + 2 * k, b0 + 2 * k, kInstrDesc, (uint32_t)(kb != 0 || k != 0));
                    }
                    umma_commit(empty_bar + 8 * stage);
                    advance();
                }
                umma_commit(acc_full + 8 * j);
            }
            cluster_sync_all();  // #1 (the layer-2 operands below only arrive after it)
            mbar_wait(acc_empty, 0u);  // layer-1 tile 0 has been read out of region 0
            tcgen05_fence_after();
            for (int kb = 0; kb < kb2; ++kb) {  // layer 2 into region 0
                mbar_wait(full_bar + 8 * stage, phase);
                tcgen05_fence_after();
                const uint64_t a0 = make_kmajor_sw128_desc(ring + stage * SMALL_STAGE_BYTES);
                const uint64_t b0 = make_kmajor_sw128_desc(ring + stage * SMALL_STAGE_BYTES + A_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tmem_base, a0 + 2 * k, b0 + 2 * k, kInstrDesc, (uint32_t)(kb != 0 || k != 0));
                umma_commit(empty_bar + 8 * stage);
                advance();
            }
            umma_commit(acc_full);  // second completion of region 0's barrier: parity 1
            cluster_sync_all();  // #2
        } else {
            cluster_sync_all();  // #1
            cluster_sync_all();  // #2
        }
    } else {
        // ================= read-out warps 2..9 (quarter = TMEM lane quarter, colhalf = column half of the accumulator) ========
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int colhalf = ew >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t col_base = (uint32_t)(colhalf * (BN / 2));
        const uint64_t h1_keep = l2_policy_evict_last();
        const uint32_t slab = staging + (uint32_t)(ew * STORE_SLAB_BYTES);
        for (int j = 0; j < 2; ++j) {  // ---- D1: ReLU -> bf16 -> slab -> TMA store into the h1 scratch (as in the fused kernel) ----
            const int nt = 2 * (int)rank + j;
            const uint32_t acc = tmem_base + lane_base + j * ACC_COLS + col_base;
            mbar_wait(acc_full + 8 * j, 0u);
            tcgen05_fence_after();
            uint32_t v[2][32];
            tmem_ld_32x32(acc, v[0]);
#pragma unroll 1
            for (int hp = 0; hp < 4; hp += 2) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int h = hp + hh;
                    const int col = nt * BN + (int)col_base + h * 32;
                    if (lane == 0) bulk_wait_read<0>();  // the previous store has drained the slab
                    __syncwarp();
                    tmem_ld_wait();
                    if (h + 1 < 4) {
                        tmem_ld_32x32(acc + (uint32_t)((h + 1) * 32), v[hh ^ 1]);
                    } else {
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + 8 * j);
                    }
                    const uint32_t(&cur)[32] = v[hh];
                    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const __nv_bfloat162 p0 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * q + 0]), __uint_as_float(cur[8 * q + 1])), zero2);
                        const __nv_bfloat162 p1 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * q + 2]), __uint_as_float(cur[8 * q + 3])), zero2);
                        const __nv_bfloat162 p2 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * q + 4]), __uint_as_float(cur[8 * q + 5])), zero2);
                        const __nv_bfloat162 p3 = __hmax2_nan(__floats2bfloat162_rn(__uint_as_float(cur[8 * q + 6]), __uint_as_float(cur[8 * q + 7])), zero2);
                        const int chunk = q ^ ((lane >> 1) & 3);  // SWIZZLE_64B
                        st_shared_v4(slab + (uint32_t)(lane * 64 + chunk * 16), *reinterpret_cast<const uint32_t *>(&p0),
                                     *reinterpret_cast<const uint32_t *>(&p1), *reinterpret_cast<const uint32_t *>(&p2),
                                     *reinterpret_cast<const uint32_t *>(&p3));
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d_hint(&tmap_h_store, slab, col, h1_row0 + quarter * 32, h1_keep);
                        bulk_commit();
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_all();  // this warp's part of h1 is in L2
        __syncwarp();
        cluster_sync_all();  // #1
        // ---- D2: this CTA's 256 columns of h2 (N tile `rank`), the fused kernel's per-thread chain -----------------------------
        float dot[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const int slot4 = (ew * 32 + lane) * 4;
        auto chain_step = [&]() {
            const uint32_t acc = tmem_base + lane_base + col_base;
            mbar_wait(acc_full, 1u);
            tcgen05_fence_after();
            uint32_t v[2][32];
            tmem_ld_32x32(acc, v[0]);
#pragma unroll 1
            for (int cp = 0; cp < 4; cp += 2) {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int chunk = cp + cc;
                    const int col0 = (int)rank * BN + (int)col_base + chunk * 32;
                    const float4 *bias4 = reinterpret_cast<const float4 *>(p.b2 + col0);
                    const float4 *w4 = reinterpret_cast<const float4 *>(p.w3 + col0);
                    float4 bb[4], ww[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        bb[q] = __ldg(bias4 + q);
                        ww[q] = __ldg(w4 + q);
                    }
                    tmem_ld_wait();
                    if (chunk + 1 < 4) tmem_ld_32x32(acc + (uint32_t)((chunk + 1) * 32), v[cc ^ 1]);
                    const uint32_t(&cur)[32] = v[cc];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (half == 1) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                bb[q] = __ldg(bias4 + 4 + q);
                                ww[q] = __ldg(w4 + 4 + q);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int e = 16 * half + 4 * q;
                            dot[0] = fmaf(max_nan(__uint_as_float(cur[e + 0]) + bb[q].x, 0.0f), ww[q].x, dot[0]);
                            dot[1] = fmaf(max_nan(__uint_as_float(cur[e + 1]) + bb[q].y, 0.0f), ww[q].y, dot[1]);
                            dot[2] = fmaf(max_nan(__uint_as_float(cur[e + 2]) + bb[q].z, 0.0f), ww[q].z, dot[2]);
                            dot[3] = fmaf(max_nan(__uint_as_float(cur[e + 3]) + bb[q].w, 0.0f), ww[q].w, dot[3]);
                        }
                    }
                }
            }
        };
        if (rank == 0) {
            chain_step();
            // hand the four running sums to the same thread of CTA 1 (its shared memory, through the cluster window)
            const uint32_t remote = mapa_rank(chain_smem + (uint32_t)slot4 * 4u, 1);
            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(dot[0]), "f"(dot[1]), "f"(dot[2]), "f"(dot[3]) : "memory");
            cluster_sync_all();  // #2 (release: the stores above are visible to CTA 1 after its wait)
        } else {
            cluster_sync_all();  // #2
            const float4 in = *reinterpret_cast<const float4 *>(chain_in + slot4);
            dot[0] = in.x; dot[1] = in.y; dot[2] = in.z; dot[3] = in.w;
            chain_step();
            const float mine = (dot[0] + dot[1]) + (dot[2] + dot[3]);
            float *slot = part + quarter * 32 + lane;
            if (colhalf == 1) *slot = mine;
            asm volatile("bar.sync %0, 64;" ::"r"(8 + quarter) : "memory");
            if (colhalf == 0) {
                const int64_t row = row_base + quarter * 32 + lane;
                if (row < p.M) {
                    const float logit = (mine + *slot) + __ldg(p.b3);
                    if (p.logits) p.logits[row] = logit;
                    p.reward[row] = style_reward(logit, p.scale);
                }
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
                                                              __nv_bfloat16 *__restrict__ out,
                                                              const int64_t *__restrict__ row_index, int64_t capacity,
                                                              uint32_t *__restrict__ flags, int ones_cols) {
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
        // the first `ones_cols` padding columns hold 1.0 (the layer-1 bias rides in the matching columns of W1, see
        // cast_weight_kernel): x is read as 0 there, so (0 - (-1)) * 1 = 1
        if (c >= in_features && c < in_features + ones_cols) { mu[j].x = -1.0f; rc[j].x = 1.0f; }
        if (c + 1 >= in_features && c + 1 < in_features + ones_cols) { mu[j].y = -1.0f; rc[j].y = 1.0f; }
    }
    // row_index != NULL: destination row r reads source row row_index[r] (RandomMemory.sample_by_index fused into the
    // preprocessor); an index outside [0, capacity) reads row 0 and raises bit 1 of *flags
    auto load_row = [&](int64_t r, float2(&v)[NB]) {
        int64_t sr = r;
        if (row_index) {
            sr = __ldg(row_index + r);
            if (sr < 0 || sr >= capacity) {
                if (lane == 0 && flags) atomicOr(flags, 2u);
                sr = 0;
            }
        }
        const float *xr = x + sr * x_stride;
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
            const float a = clamp_nan(__fmul_rn(__fsub_rn(v[j].x, mu[j].x), rc[j].x), -5.0f, 5.0f);
            const float b = clamp_nan(__fmul_rn(__fsub_rn(v[j].y, mu[j].y), rc[j].y), -5.0f, 5.0f);
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
                                 const float *mean, const float *denom, __nv_bfloat16 *out, const int64_t *row_index,
                                 int64_t capacity, uint32_t *flags, int ones_cols = 0) {
#define AMP_CAST_CASE(NBV)                                                                                            \
    case NBV:                                                                                                         \
        normalise_cast_kernel<NBV, VEC><<<grid, 256, 0, st>>>(x, x_stride, rows, in_features, mean, denom, out,       \
                                                              row_index, capacity, flags, ones_cols);                 \
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

// fp32 master (rows, cols) -> bf16 (rows, cols_padded), zero padded.  bias != NULL: the layer's bias is folded into the
// product -- padding columns `cols` and `cols + 1` receive bias[r] split into a bf16 head and a bf16 tail (hi + lo carries
// 16 mantissa bits; the input holds 1.0 in those two columns), so the accumulator already contains x W^T + b.
__global__ void cast_weight_kernel(const float *__restrict__ w, int rows, int cols, int cols_padded,
                                   __nv_bfloat16 *__restrict__ out, const float *__restrict__ bias) {
    const int64_t total = (int64_t)rows * cols_padded;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols_padded), c = (int)(e - (int64_t)r * cols_padded);
        float v = c < cols ? w[(int64_t)r * cols + c] : 0.0f;
        if (bias && c >= cols && c < cols + 2) {
            const float b = bias[r];
            const float hi = __bfloat162float(__float2bfloat16_rn(b));
            v = c == cols ? hi : b - hi;
        }
        out[e] = __float2bfloat16_rn(v);
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

}  // namespace disc

int scaler_stats_to_f32(const double *mean, const double *var, int n, float *mean_f, float *denom_f, cudaStream_t st) {
    disc::scaler_stats_kernel<<<(n + 255) / 256, 256, 0, st>>>(mean, var, n, mean_f, denom_f);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

}  // namespace amp

// Batches of at most this many row tiles per SM take the cast-kernel path (see style_reward_impl)
constexpr int kCastPathWaves = 8;

struct amp_disc {
    int in_features, Kp, h1, h2;
    int ws_ctas;  // persistent CTAs the h1 / x_hat scratch slots were sized for
    int device;
    __nv_bfloat16 *W1, *W2;       // (h1, Kp), (h2, h1) bf16
    float *b1, *b2, *w3, *b3;     // fp32
    float *mean, *denom;          // fp32 [in_features]
    __nv_bfloat16 *xs, *hid;      // per-CTA scratch slots, L2-resident: x_hat (ws_ctas * 2 * 128, Kp) and h1 (ws_ctas * 2 * 128, h1)
    // Wide inputs (Kp > 256, e.g. K*A = 830): the in-kernel converter warps cannot keep enough bytes in flight for 3.3 KB
    // rows (measured 0.35 ms vs 0.24 ms per 65 536 x 830 rows), so x_hat of a whole chunk is prepared by normalise_cast_kernel
    // in `xs` (xs_rows x Kp) and the fused kernel streams it by row tile.  Narrow inputs: one launch, converter warps.
    bool wide;
    int64_t xs_rows;              // rows of `xs`: ws_ctas * kCastPathWaves * 128 (narrow) or the chunk capacity (wide)
    CUtensorMap tmap_x;           // x_hat: 128-row loads
    CUtensorMap tmap_w1, tmap_w2; // weights never move: encoded once
    CUtensorMap tmap_w1_half, tmap_w2_half; // 128-row boxes: each CTA of a pair stages its half of a 256-row weight block
    CUtensorMap tmap_h_load, tmap_h_store;  // h1 scratch: 128-row loads, 32-row epilogue slab stores
    long long *prof;                        // AMP_DISC_PROFILE builds: device counters (ws_ctas x 16)
    bool small_ok;                          // the two-CTA-per-tile kernel applies (1024-512 hidden sizes, K*A <= 190)
    bool use_pair;                          // CTA-pair (cta_group::2) kernel (default; AMP_B200_DISC_PAIR=0 at create selects single)
    int prof_mode;                          // AMP_DISC_PROFILE builds: AMP_DISC_PROF_MODE read once at create
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
    AMP_REQUIRE(in_features <= 1022, "amp_disc_create: discriminator input wider than 1022 columns is not supported (got %d)",
                in_features);
    AMP_REQUIRE(max_rows >= 1, "amp_disc_create: max_rows must be positive");  // a sizing hint only: nothing scales with it
    int dev = 0, major = 0;
    AMP_CUDA_TRY(cudaGetDevice(&dev));
    AMP_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(AMP_ENODEV, "amp_disc_create: tcgen05 kernels need an sm_100 device (found compute capability %d.x)", major);

    amp_disc *d = new (std::nothrow) amp_disc();
    if (!d) return fail(AMP_ENOMEM, "amp_disc_create: host allocation failed");
    std::memset(d, 0, sizeof(*d));
    d->device = dev;
    d->in_features = in_features;
    // b1 rides inside the layer-1 product: x_hat carries 1.0 in its first two padding columns and W1 carries b1 there as a
    // bf16 head + tail, so Kp always leaves at least two padding columns
    d->Kp = (in_features + 2 + BK - 1) / BK * BK;
    d->h1 = h1;
    d->h2 = h2;
    d->ws_ctas = sm_count();  // two 128-row slots of x_hat and of h1 per persistent CTA
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
    d->wide = d->Kp > 256;
    // narrow inputs: the in-kernel converter needs two 128-row slots per CTA; the same buffer, sized for eight persistent waves
    // (151 552 rows, 58 MB at K*A = 166), is the x_hat of a whole batch on the cast-kernel path below the crossover
    d->xs_rows = (int64_t)d->ws_ctas * kCastPathWaves * BM;
    if (d->wide)  // one chunk = the caller's max_rows, capped at 512 MB of x_hat; larger batches are cut into chunks
        d->xs_rows = std::max<int64_t>(BM, std::min<int64_t>((max_rows + BM - 1) / BM * BM, (((int64_t)512 << 20) / (d->Kp * 2)) / BM * BM));
    if (e == cudaSuccess) e = alloc((void **)&d->xs, (size_t)d->xs_rows * d->Kp * 2);
    if (e == cudaSuccess) e = alloc((void **)&d->hid, (size_t)d->ws_ctas * 2 * BM * h1 * 2);
    if (e == cudaSuccess) e = cudaMemset(d->xs, 0, (size_t)d->xs_rows * d->Kp * 2);
    if (e != cudaSuccess) {
        amp_disc_destroy(d);
        return cuda_fail(e, "cudaMalloc(amp_disc_create)");
    }
    int rc = make_tmap(&d->tmap_w1, d->W1, h1, d->Kp, d->Kp, BN);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_w2, d->W2, h2, h1, h1, BN);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_w1_half, d->W1, h1, d->Kp, d->Kp, BN / 2);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_w2_half, d->W2, h2, h1, h1, BN / 2);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_x, d->xs, d->xs_rows, d->Kp, d->Kp, BM);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_h_load, d->hid, (int64_t)d->ws_ctas * 2 * BM, h1, h1, BM);
    // epilogue slabs: 32 rows x 32 columns (64-byte rows, 64B swizzle)
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_h_store, d->hid, (int64_t)d->ws_ctas * 2 * BM, h1, h1, 32, SLAB_COLS, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc == AMP_OK) {
        const char *small_env = getenv("AMP_B200_DISC_SMALL");  // 0 disables the small-batch kernel (tests compare the paths)
        d->small_ok = h1 == 4 * BN && h2 == 2 * BN && d->Kp / BK <= 3 && !(small_env && small_env[0] == '0');
        if (d->small_ok) {
            // the attribute belongs to the function, not to the handle: always the largest footprint (kb1 = 3), or a later handle
            // with a narrower input would lower the limit under an earlier one
            e = cudaFuncSetAttribute(disc_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem_bytes(3));
            if (e != cudaSuccess) rc = cuda_fail(e, "cudaFuncSetAttribute(disc_small_kernel)");
        }
        e = cudaFuncSetAttribute(disc_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem_bytes(false));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(disc_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem_bytes(true));
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaFuncSetAttribute(disc_fused_kernel)");
#ifdef AMP_DISC_PROFILE
        if (cudaMalloc((void **)&d->prof, (size_t)d->ws_ctas * 16 * sizeof(long long)) != cudaSuccess) d->prof = nullptr;
        if (const char *m = getenv("AMP_DISC_PROF_MODE")) d->prof_mode = atoi(m);
        {  // timing experiments that skip the converter still feed the MMAs realistic operands (tensor-core power depends on data)
            const size_t n = (size_t)d->xs_rows * d->Kp;
            std::vector<uint16_t> h(n);
            uint32_t st = 12345u;
            for (size_t i = 0; i < n; ++i) {
                st = st * 1664525u + 1013904223u;
                h[i] = (uint16_t)(((st >> 16) & 0x807f) | (0x7e + ((st >> 8) & 1)) << 7);  // +-[0.5, 2) bf16
            }
            cudaMemcpy(d->xs, h.data(), n * 2, cudaMemcpyHostToDevice);
        }
#endif
        // The CTA pair (cta_group::2) is the default: each SM stages only half of every weight block (a third less L2 -> SM
        // traffic per row).  Measured, 1 M rows x 166: pair 1.35 ms, single CTA 1.49 ms; 65 536 x 830: 0.218 vs 0.232 ms.  (Round 1
        // measured the pair slower -- its accumulator hand-off used mbarrier.arrive.release.cluster, which ptxas expands to
        // MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR: see mbar_arrive_cluster.)  AMP_B200_DISC_PAIR=0 selects the single-CTA kernel;
        // both are covered by the GPU tests.
        const char *pair = getenv("AMP_B200_DISC_PAIR");
        d->use_pair = !(pair && pair[0] == '0');
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
    void *ptrs[] = {d->W1, d->W2, d->b1, d->b2, d->w3, d->b3, d->mean, d->denom, d->xs, d->hid, d->prof};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete d;
    return AMP_OK;
}

// rows one persistent wave covers (one 128-row tile per CTA)
int64_t amp_disc_chunk_rows(const amp_disc_t *d) { return d ? (int64_t)d->ws_ctas * BM : 0; }

// narrow inputs (K*A <= 254): ONE launch for batches beyond eight persistent waves, cast + fused kernel below that (too
// little to hide the first tile's conversion under); wide inputs: cast + fused kernel per chunk of xs_rows rows
int64_t amp_disc_launch_count(const amp_disc_t *d, int64_t M) {
    if (!d || M <= 0) return 0;
    if (d->wide) return 2 * ((M + d->xs_rows - 1) / d->xs_rows);
    if (d->small_ok && (M + BM - 1) / BM <= sm_count() / 2) return 1;  // the two-CTA-per-tile kernel
    return M <= d->xs_rows ? 2 : 1;
}

int amp_disc_load(amp_disc_t *d, const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                  const float *b3, const double *running_mean, const double *running_variance, void *stream) {
    AMP_REQUIRE(d && W1 && b1 && W2 && b2 && W3 && b3 && running_mean && running_variance, "amp_disc_load: NULL argument");
    cudaStream_t st = as_stream(stream);
    const int blocks = sm_count() * 4;
    cast_weight_kernel<<<blocks, 256, 0, st>>>(W1, d->h1, d->in_features, d->Kp, d->W1, b1);
    cast_weight_kernel<<<blocks, 256, 0, st>>>(W2, d->h2, d->h1, d->h1, d->W2, nullptr);
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

}  // extern "C"

// x rows are taken in order (row_index == NULL) or gathered: row r of the batch = x[row_index[r]] with x holding `capacity` rows.
// ONE launch: scaler + cast + both layers + reward; the grid is min(row tiles, SMs) persistent CTAs.
static int style_reward_impl(amp_disc_t *d, const float *x, int64_t x_stride, const int64_t *row_index, int64_t capacity,
                             uint32_t *flags, int64_t M, float reward_scale, float *reward, float *logits, void *stream) {
    AMP_REQUIRE(d && M >= 0, "amp_disc_style_reward: bad handle or negative size");
    AMP_REQUIRE(d->loaded, "amp_disc_style_reward: amp_disc_load has not been called");
    if (M == 0) return AMP_OK;
    AMP_REQUIRE(x && reward, "amp_disc_style_reward: NULL buffer");
    AMP_REQUIRE(x_stride >= d->in_features, "amp_disc_style_reward: x_stride %lld < in_features %d", (long long)x_stride,
                d->in_features);
    AMP_REQUIRE(M <= ((int64_t)1 << 31) - BM, "amp_disc_style_reward: %lld rows exceed the 2^31 row limit of one call", (long long)M);
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    // Small batches (at most one row tile per two SMs): ONE launch of the two-CTA-per-tile kernel, cast folded in
    if (d->small_ok && (M + BM - 1) / BM <= sms / 2) {
        SmallParams sp{};
        sp.M = M;
        sp.x = x;
        sp.x_stride = x_stride;
        sp.row_index = row_index;
        sp.capacity = capacity;
        sp.flags = flags;
        sp.mean = d->mean;
        sp.denom = d->denom;
        sp.in_features = d->in_features;
        sp.kb1 = d->Kp / BK;
        sp.ksteps1_last = (d->in_features + 2 - BK * (sp.kb1 - 1) + UMMA_K - 1) / UMMA_K;
        sp.x_vec = ((x_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 7u) == 0)) ? 1 : 0;
        sp.n1_tiles = d->h1 / BN;
        sp.n2_tiles = d->h2 / BN;
        sp.b2 = d->b2;
        sp.w3 = d->w3;
        sp.b3 = d->b3;
        sp.scale = reward_scale;
        sp.reward = reward;
        sp.logits = logits;
        const int m_tiles = (int)((M + BM - 1) / BM);
        disc_small_kernel<<<2 * m_tiles, SMALL_THREADS, small_smem_bytes(sp.kb1), st>>>(d->tmap_w1, d->tmap_h_load, d->tmap_h_store,
                                                                                        d->tmap_w2, sp);
        AMP_CUDA_TRY(cudaGetLastError());
        return AMP_OK;
    }
    // Gathered batches (row_index) always take the cast-kernel path, in chunks of the scratch: random 664-byte rows out of a
    // multi-GB memory are pure DRAM latency, which six converter warps per SM cannot cover (measured 1 M sampled rows x 166:
    // 1.99 ms in-kernel against 1.43 ms with the gather done by the cast kernel at full-chip parallelism)
    const bool always_cast = d->wide || row_index != nullptr;
    const int64_t chunk = always_cast ? d->xs_rows : M;
    for (int64_t r0 = 0; r0 < M; r0 += chunk) {
        const int64_t rows = std::min(chunk, M - r0);
        const int m_tiles = (int)((rows + BM - 1) / BM);
        const float *xc = row_index ? x : x + r0 * x_stride;
        const int64_t *ic = row_index ? row_index + r0 : nullptr;
        const bool vec = (x_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(xc) & 7u) == 0);
        // Small and medium batches have little to hide the first tile's conversion under: the converter warps of a CTA run 9
        // dependent load rounds (~17 us at 65 536 rows) before its first MMA.  Measured cast path vs in-kernel conversion, x 166:
        // 4096 rows 39 vs 58 us, 65 536 rows 111 vs 125 us, 131 072 rows 177 vs 182 us, 262 144 rows equal, 1 M rows 1.29 vs 1.25 ms
        // (and no 384 MB workspace).  Crossover = eight row tiles per SM; the scratch holds exactly that many rows.
        const bool external = always_cast || rows <= d->xs_rows;
        if (external) {
            const int cast_grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
            int rc = vec ? launch_normalise_cast<true>(d->Kp / BK, cast_grid, st, xc, x_stride, rows, d->in_features, d->mean, d->denom,
                                                       d->xs, ic, capacity, flags, 2)
                         : launch_normalise_cast<false>(d->Kp / BK, cast_grid, st, xc, x_stride, rows, d->in_features, d->mean, d->denom,
                                                        d->xs, ic, capacity, flags, 2);
            if (rc != AMP_OK) return rc;
            AMP_CUDA_TRY(cudaGetLastError());
        }
        FusedParams fp{};
        fp.M = rows;
        fp.x = xc;
        fp.x_stride = x_stride;
        fp.mean = d->mean;
        fp.denom = d->denom;
        fp.xs = d->xs;
        fp.in_features = d->in_features;
        fp.Kp = d->Kp;
        fp.x_vec = vec ? 1 : 0;
        fp.xhat_rows_in_order = external ? 1 : 0;
        fp.kb1 = d->Kp / BK;
        fp.ksteps1_last = (d->in_features + 2 - BK * (fp.kb1 - 1) + UMMA_K - 1) / UMMA_K;
        fp.n1_tiles = d->h1 / BN;
        fp.n2_tiles = d->h2 / BN;
        fp.b2 = d->b2;
        fp.w3 = d->w3;
        fp.b3 = d->b3;
        fp.scale = reward_scale;
        fp.reward = reward + r0;
        fp.logits = logits ? logits + r0 : nullptr;
        fp.prof = d->prof;
        fp.prof_mode = d->prof_mode;
        fp.prof_ctas = d->ws_ctas;
#ifdef AMP_DISC_PROFILE
        if (d->prof) cudaMemsetAsync(d->prof, 0, (size_t)d->ws_ctas * 16 * sizeof(long long), st);  // CTAs outside this grid read 0
#endif
        // CTA pairs once there is more than one row tile per SM; below that single CTAs spread the tiles over twice as many
        // independent pipelines (4096 rows: 45 us single, 47-51 us pair)
        if (d->use_pair && m_tiles > sms) {
            // CTA pairs: an even grid of at most ws_ctas CTAs, launched as clusters of 2
            const int pair_blocks = (m_tiles + 1) / 2;
            const int clusters = std::min(pair_blocks, std::min(sms, d->ws_ctas) / 2);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(2 * clusters);
            cfg.blockDim = dim3(FUSED_THREADS);
            cfg.dynamicSmemBytes = fused_smem_bytes(true);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            AMP_CUDA_TRY(cudaLaunchKernelEx(&cfg, disc_fused_kernel<true>, d->tmap_x, d->tmap_w1_half, d->tmap_h_load, d->tmap_h_store,
                                            d->tmap_w2_half, fp));
        } else {
            const int grid = std::min(m_tiles, std::min(sms, d->ws_ctas));
            disc_fused_kernel<false><<<grid, FUSED_THREADS, fused_smem_bytes(false), st>>>(d->tmap_x, d->tmap_w1, d->tmap_h_load,
                                                                                           d->tmap_h_store, d->tmap_w2, fp);
        }
        AMP_CUDA_TRY(cudaGetLastError());
    }
#ifdef AMP_DISC_PROFILE
    if (d->prof) {  // developer build: dump the wait breakdown of this launch (synchronises!)
        static long long host[1024 * 16];
        cudaStreamSynchronize(st);
        const int n = std::min(d->ws_ctas, 1024);
        cudaMemcpy(host, d->prof, (size_t)n * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
        {  // epilogue warp 2 of every CTA (for a pair: even CTAs = leaders, odd = peers); assumes the grid used all ws_ctas CTAs
            double ea[2][8] = {{0}};
            for (int i = 0; i < n; ++i)
                for (int k = 0; k < 8; ++k) ea[d->use_pair ? (i & 1) : 0][k] += (double)host[n * 8 + i * 8 + k];
            int ran = 0;
            for (int i = 0; i < n; ++i) ran += (host[n * 8 + i * 8 + 1] + host[n * 8 + i * 8 + 4]) > 0;
            const double div = std::max(1.0, d->use_pair ? ran / 2.0 : (double)ran);
            for (int rk = 0; rk < (d->use_pair ? 2 : 1); ++rk)
                fprintf(stderr, "[amp_disc profile] epilogue rank %d: D1 wait=%.0f drain=%.0f post=%.0f slab_wait=%.0f | D2 wait=%.0f drain=%.0f post=%.0f\n", rk,
                        ea[rk][0] / div, ea[rk][1] / div, ea[rk][2] / div, ea[rk][6] / div, ea[rk][3] / div, ea[rk][4] / div, ea[rk][5] / div);
        }
        double acc[8] = {0};
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            if (host[i * 8 + 0] <= 0) continue;
            ++cnt;
            for (int k = 0; k < 8; ++k) acc[k] += (double)host[i * 8 + k];
        }
        if (cnt)
            fprintf(stderr, "[amp_disc profile] CTA wall: entry->setup done %.2f us, entry->all roles done %.2f us (mean over CTAs that ran)\n",
                    [&] { double a = 0; int c = 0; for (int i = 0; i < n; ++i) if (host[i * 8 + 0] > 0) { a += (double)host[n * 8 + i * 8 + 7]; ++c; } return c ? a / c * 1e-3 : 0.0; }(),
                    [&] { double a = 0; int c = 0; for (int i = 0; i < n; ++i) if (host[i * 8 + 0] > 0) { a += (double)host[i * 8 + 7]; ++c; } return c ? a / c * 1e-3 : 0.0; }());
            {
                double mn = 1e30, mx = 0; long long e0 = 0x7fffffffffffffffLL, e1 = 0;
                for (int i = 0; i < n; ++i) if (host[i * 8 + 0] > 0) { mn = std::min(mn, (double)host[i * 8 + 7]); mx = std::max(mx, (double)host[i * 8 + 7]); }
                fprintf(stderr, "[amp_disc profile] CTA wall min %.2f us max %.2f us\n", mn * 1e-3, mx * 1e-3);
            }
            fprintf(stderr, "[amp_disc profile] issuer CTAs=%d total=%.0f wait_full=%.0f wait_d1_empty=%.0f wait_d2_empty=%.0f | producer(all) wait_empty=%.0f wait_h1+xhat=%.0f (cycles, mean per CTA) | issuer wall %.1f us => SM clock %.0f MHz\n",
                    cnt, acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[4] / n, acc[5] / n, acc[6] / cnt * 1e-3,
                    acc[0] / acc[6] * 1e3);
    }
#endif
    return AMP_OK;
}

extern "C" {

int amp_disc_style_reward(amp_disc_t *d, const float *x, int64_t x_stride, int64_t M, float reward_scale, float *reward,
                          float *logits, void *stream) {
    return style_reward_impl(d, x, x_stride, nullptr, 0, nullptr, M, reward_scale, reward, logits, stream);
}

int amp_disc_style_reward_indexed(amp_disc_t *d, const float *memory, int64_t memory_stride, int64_t capacity,
                                  const int64_t *row_index, int64_t M, float reward_scale, float *reward, float *logits,
                                  uint32_t *flags, void *stream) {
    AMP_REQUIRE(M == 0 || (row_index && capacity >= 1), "amp_disc_style_reward_indexed: NULL row_index or empty memory");
    return style_reward_impl(d, memory, memory_stride, row_index, capacity, flags, M, reward_scale, reward, logits, stream);
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
