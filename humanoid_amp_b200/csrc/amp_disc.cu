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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

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
// Warp roles (320 threads): 0 TMA producer, 1 TMEM alloc + MMA issuer, 2..9 epilogue (warps w and w+4 share a TMEM lane
// quarter and split the accumulator columns).  Producer, issuer and epilogue warps all walk the same static schedule.
// =====================================================================================================================
// PAIR = true runs the same pipeline on a CTA pair (cluster of 2, tcgen05 cta_group::2): one MMA covers 256 rows (128 per
// CTA) and each CTA stages only ITS half of the 256-row weight block, so the operand bytes delivered per SM per MMA drop
// from 48 KB to 32 KB per K block.  ncu on the single-CTA version showed TMA loads pinned at 40 % of the xbar->L1 peak on
// every SM (2.57 GB per 151 552 rows, ~51 B/cycle/SM) with the tensor pipe 55 % active: operand delivery, not math, is
// the limiter, and a CTA pair is the only way to shrink it without more TMEM.
constexpr int FUSED_THREADS = 320;
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
    for (int nt = 0; nt < n1_tiles; ++nt) g1(0, nt, nt & 1);
    const int seg = (units + n1_tiles - 1) / n1_tiles;
    for (int i = 0; i < T; ++i) {
        int u = 0;
        for (int nt = 0; nt < n1_tiles; ++nt) {
            if (i + 1 < T) g1(i + 1, nt, 0);
            for (const int e = min(units, u + seg); u < e; ++u) g2(i, u, 1);
        }
    }
}

struct FusedParams {
    int M;          // rows of this launch (x_hat rows)
    int kb1;        // K blocks of layer 1 (Kp / 64)
    int ksteps1_last;  // 16-wide K steps the LAST layer-1 block really needs: ceil((in_features [+ 2] - 64 (kb1 - 1)) / 16), 1..4
    int fold_bias1;    // b1 travels inside the layer-1 product (two padding columns of x_hat / W1): D1 is cvt + max only
    int n1_tiles;   // h1 / 256
    int n2_tiles;   // h2 / 256
    const float *b1, *b2, *w3, *b3;
    float scale;
    float *reward;  // [M]
    float *logits;  // [M] or NULL
    long long *prof;  // AMP_DISC_PROFILE builds only: per-CTA cycle counters of the producer / issuer waits, else NULL
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
    const uint32_t part_smem = bars + 256;  // 2 x 4 x 32 floats
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int num_m_tiles = (p.M + BM - 1) / BM;
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
    // accumulator-drained barriers live in the leader CTA; the epilogue threads of both CTAs arrive there
    const uint32_t acc_empty_at_leader = PAIR ? mapa_rank(acc_empty, 0) : acc_empty;  // + 8 * region
    auto arrive_drained = [&](uint32_t addr) {
        if constexpr (PAIR) mbar_arrive_cluster(addr); else mbar_arrive(addr);
    };

    if (warp == 0) {
        // ================= TMA producer (every CTA loads its own rows of A and its own part of the weight block) =========
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] long long w_empty = 0, w_h1 = 0;
            // weights and h1 slots: evict_last.  x_hat: normal priority while its row tile is being re-read (a 48 KB tile comes
            // back within a few microseconds), evict_first on its last pass -- marking it evict_last made 192 MB of x_hat
            // per launch compete with the 77.6 MB of h1 slots for the protected part of L2.
            const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first(), normal = l2_policy_evict_normal();
            auto load_pair = [&](const CUtensorMap *ma, int a_col, int a_row, uint64_t a_policy, const CUtensorMap *mb, int b_col,
                                 int b_row) {
                {
                    AMP_PROF_T0;
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    AMP_PROF_ADD(w_empty);
                }
                const uint32_t fb = full_bar + 8 * stage;
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
            };
            walk_schedule(
                T, p.n1_tiles, units,
                [&](int ti, int nt, int) {
                    const int m = tile_of(ti);
                    // x_hat is read by the four N1 tiles of this row tile and never again: stream it on the last pass
                    for (int kb = 0; kb < p.kb1; ++kb)
                        load_pair(&tmap_x, kb * BK, m * BM, nt == p.n1_tiles - 1 ? stream : normal, &tmap_w1, kb * BK, nt * BN);
                },
                [&](int ti, int u, int) {
                    const int n2 = u / kb2, kb = u - n2 * kb2;
                    if (u == 0) {  // h1 of this row tile has been written (all four E1 warps' TMA stores completed)
                        AMP_PROF_T0;
                        mbar_wait(h1_ready + 8 * (ti & 1), (uint32_t)((ti >> 1) & 1));
                        AMP_PROF_ADD(w_h1);
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                    load_pair(&tmap_h_load, kb * BK, slot_row0 + (ti & 1) * BM, keep, &tmap_w2, kb * BK, n2 * BN);
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
        // ================= epilogue warps 2..9: both accumulators, columns split between the two warps of a TMEM quarter ======
        // Warp w and warp w+4 own the same 32 TMEM lanes (rows); w takes columns [0,128) of every 256-column accumulator and
        // w+4 takes [128,256), so each drain is half as long as with one warp per quarter -- the issuer's in-kernel cycle
        // counters showed it waiting ~20 % of the time for D2 and ~12 % for D1 to be drained.  The warps follow the issuer's
        // schedule: D1 tiles (bias + ReLU -> bf16 -> swizzled slab -> TMA store of h1) and, after the last K block of each
        // layer-2 N tile, D2 (bias + ReLU -> dot with w3).  tcgen05.ld of the next 32 columns is in flight while the current
        // 32 are processed.
        const int ew = warp - 2;
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
#pragma unroll
                for (int h = 0; h < 4; ++h) {  // four steps of 32 columns = four slabs
                    const int col = nt * BN + (int)col_base + h * 32;  // h1 column of this step
                    const float4 *bias4 = reinterpret_cast<const float4 *>(p.b1 + col);
                    float4 bb[8];
                    if (!p.fold_bias1) {  // launch-uniform
#ifdef AMP_DISC_PROFILE
                        if (p.prof_mode & 2) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        } else
#endif
                        {
#pragma unroll
                            for (int j = 0; j < 8; ++j) bb[j] = __ldg(bias4 + j);
                        }
                    }
                    const uint32_t slab = staging + (uint32_t)((ew * 2 + (store_it & 1)) * STORE_SLAB_BYTES);
                    {
                        AMP_PROF_T0;
                        if (lane == 0) bulk_wait_read<1>();  // the store that last read this slab has drained it
                        __syncwarp();
                        AMP_PROF_ADD(e_slab);
                    }
                    tmem_ld_wait();  // v[h & 1] has landed
                    if (h + 1 < 4) {
                        tmem_ld_32x32(acc + (uint32_t)((h + 1) * 32), v[(h + 1) & 1]);
                    } else {  // last TMEM read of this accumulator by this warp: release the region to the issuer now
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_drained(acc_empty_at_leader + 8 * r);
#ifdef AMP_DISC_PROFILE
                        e_t2 = clock64();
#endif
                    }
                    const uint32_t(&cur)[32] = v[h & 1];
#ifdef AMP_DISC_PROFILE
                    if (p.prof_mode & 1) { ++store_it; continue; }
#endif
#pragma unroll
                    for (int j = 0; j < 4; ++j) {  // 16-byte chunk j of this thread's 64-byte slab row
                        __nv_bfloat162 p0, p1, p2, p3;
                        if (p.fold_bias1) {
                            // the accumulator already holds x W1^T + b1: round, then ReLU on the packed pair (rounding is
                            // monotonic and keeps the sign, so max(round(x), 0) == round(max(x, 0))): 1 instruction per element
                            const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
                            p0 = __hmax2(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 0]), __uint_as_float(cur[8 * j + 1])), zero2);
                            p1 = __hmax2(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 2]), __uint_as_float(cur[8 * j + 3])), zero2);
                            p2 = __hmax2(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 4]), __uint_as_float(cur[8 * j + 5])), zero2);
                            p3 = __hmax2(__floats2bfloat162_rn(__uint_as_float(cur[8 * j + 6]), __uint_as_float(cur[8 * j + 7])), zero2);
                        } else {
                            const float4 b0 = bb[2 * j], b1v = bb[2 * j + 1];
                            p0 = __floats2bfloat162_rn(fmaxf(__uint_as_float(cur[8 * j + 0]) + b0.x, 0.0f),
                                                       fmaxf(__uint_as_float(cur[8 * j + 1]) + b0.y, 0.0f));
                            p1 = __floats2bfloat162_rn(fmaxf(__uint_as_float(cur[8 * j + 2]) + b0.z, 0.0f),
                                                       fmaxf(__uint_as_float(cur[8 * j + 3]) + b0.w, 0.0f));
                            p2 = __floats2bfloat162_rn(fmaxf(__uint_as_float(cur[8 * j + 4]) + b1v.x, 0.0f),
                                                       fmaxf(__uint_as_float(cur[8 * j + 5]) + b1v.y, 0.0f));
                            p3 = __floats2bfloat162_rn(fmaxf(__uint_as_float(cur[8 * j + 6]) + b1v.z, 0.0f),
                                                       fmaxf(__uint_as_float(cur[8 * j + 7]) + b1v.w, 0.0f));
                        }
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
#pragma unroll
                for (int chunk = 0; chunk < 4; ++chunk) {
                    const int col0 = n2 * BN + (int)col_base + chunk * 32;
                    const float4 *bias4 = reinterpret_cast<const float4 *>(p.b2 + col0);
                    const float4 *w4 = reinterpret_cast<const float4 *>(p.w3 + col0);
                    float4 bb[8], ww[8];
#ifdef AMP_DISC_PROFILE
                    if (p.prof_mode & 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) bb[j] = ww[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    } else
#endif
                    {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            bb[j] = __ldg(bias4 + j);
                            ww[j] = __ldg(w4 + j);
                        }
                    }
                    tmem_ld_wait();  // v[chunk & 1] has landed
                    if (chunk + 1 < 4) {
                        tmem_ld_32x32(acc + (uint32_t)((chunk + 1) * 32), v[(chunk + 1) & 1]);
                    } else {  // last TMEM read: hand the region back before finishing the arithmetic
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_drained(acc_empty_at_leader + 8 * r);
#ifdef AMP_DISC_PROFILE
                        e_t2 = clock64();
#endif
                    }
                    const uint32_t(&cur)[32] = v[chunk & 1];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        dot[0] = fmaf(fmaxf(__uint_as_float(cur[4 * j + 0]) + bb[j].x, 0.0f), ww[j].x, dot[0]);
                        dot[1] = fmaf(fmaxf(__uint_as_float(cur[4 * j + 1]) + bb[j].y, 0.0f), ww[j].y, dot[1]);
                        dot[2] = fmaf(fmaxf(__uint_as_float(cur[4 * j + 2]) + bb[j].z, 0.0f), ww[j].z, dot[2]);
                        dot[3] = fmaf(fmaxf(__uint_as_float(cur[4 * j + 3]) + bb[j].w, 0.0f), ww[j].w, dot[3]);
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
                        const int row = tile_of(ti) * BM + quarter * 32 + lane;
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
            long long *e = p.prof + (size_t)gridDim.x * 8 + blockIdx.x * 8;
            e[0] = e_wait1; e[1] = e_drain1; e[2] = e_post1; e[3] = e_wait2; e[4] = e_drain2; e[5] = e_post2; e[6] = e_slab;
        }
#endif
    }

    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
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

int normalise_cast_rows(const float *x, int64_t x_stride, int64_t rows, int in_features, int Kp, const float *mean_f,
                        const float *denom_f, void *out_bf16, cudaStream_t st) {
    if (rows <= 0) return AMP_OK;
    const bool vec = (x_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 7u) == 0);
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 8);
    __nv_bfloat16 *out = static_cast<__nv_bfloat16 *>(out_bf16);
    int rc = vec ? disc::launch_normalise_cast<true>(Kp / disc::BK, grid, st, x, x_stride, rows, in_features, mean_f, denom_f, out,
                                                     nullptr, 0, nullptr)
                 : disc::launch_normalise_cast<false>(Kp / disc::BK, grid, st, x, x_stride, rows, in_features, mean_f, denom_f, out,
                                                      nullptr, 0, nullptr);
    if (rc != AMP_OK) return rc;
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

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
    CUtensorMap tmap_w1_half, tmap_w2_half; // 128-row boxes: each CTA of a pair stages its half of a 256-row weight block
    CUtensorMap tmap_h_load, tmap_h_store;  // h1 workspace: 128-row loads, 32-row epilogue slab stores
    long long *prof;                        // AMP_DISC_PROFILE builds: device counters (ws_ctas x 8)
    bool use_pair;                          // CTA-pair (cta_group::2) kernel, opt-in with AMP_B200_DISC_PAIR=1
    bool fold_bias1;                        // Kp - in_features >= 2: b1 rides in two padding columns of W1 (x_hat holds 1.0 there)
    bool loaded;
};

using namespace amp;
using namespace amp::disc;

struct ChunkPlan;
static ChunkPlan plan_chunks(const amp_disc *d, int64_t M);
static int64_t chunk_count(const amp_disc *d, int64_t M);

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
    d->fold_bias1 = d->Kp - in_features >= 2 && !(getenv("AMP_B200_DISC_NO_FOLD") && getenv("AMP_B200_DISC_NO_FOLD")[0] == '1');
    d->h1 = h1;
    d->h2 = h2;
    // Rows per launch ("chunk").  Measured (tools/sweep_disc.py, 1 M rows): 2 / 4 / 8 / 16 / 32 tiles per CTA per launch give
    // 1.83 / 1.63 / 1.51 / 1.46 / 1.43 ms -- every launch pays an un-overlapped layer-1 prologue and a layer-2 tail, while
    // keeping x_hat L2-resident buys nothing (the fused kernel is not DRAM-bound).  So chunks are as large as a 256 MB x_hat
    // buffer allows, but a big batch is still cut in two so the cast of the second half overlaps the first fused launch.
    const int64_t wave_rows = (int64_t)sm_count() * BM;
    const int64_t per_wave_bytes = wave_rows * d->Kp * 2;
    int64_t tiles_per_cta = std::max<int64_t>(1, std::min<int64_t>(32, ((int64_t)256 << 20) / per_wave_bytes));
    if (const char *t = getenv("AMP_B200_DISC_TILES_PER_CTA")) tiles_per_cta = std::max(1, atoi(t));  // tuning knob
    const int64_t rows_padded = (max_rows + BM - 1) / BM * BM;
    d->chunk_rows = std::min<int64_t>(rows_padded, wave_rows * tiles_per_cta);
    if (rows_padded >= 4 * wave_rows && rows_padded <= 2 * d->chunk_rows)  // two balanced chunks instead of one (+ a sliver)
        d->chunk_rows = ((rows_padded + 1) / 2 + wave_rows - 1) / wave_rows * wave_rows;
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
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_w1_half, d->W1, h1, d->Kp, d->Kp, BN / 2);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_w2_half, d->W2, h2, h1, h1, BN / 2);
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_h_load, d->hid, (int64_t)d->ws_ctas * 2 * BM, h1, h1, BM);
    // epilogue slabs: 32 rows x 32 columns (64-byte rows, 64B swizzle)
    if (rc == AMP_OK) rc = make_tmap(&d->tmap_h_store, d->hid, (int64_t)d->ws_ctas * 2 * BM, h1, h1, 32, SLAB_COLS, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc == AMP_OK) {
        e = cudaFuncSetAttribute(disc_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem_bytes(false));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(disc_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem_bytes(true));
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaFuncSetAttribute(disc_fused_kernel)");
    #ifdef AMP_DISC_PROFILE
        if (cudaMalloc((void **)&d->prof, (size_t)d->ws_ctas * 16 * sizeof(long long)) != cudaSuccess) d->prof = nullptr;
#endif
        // Measured on the 1 M-row bench (B200, sustained, sw_power_cap active): single-CTA 1.61 ms, CTA pair 1.68 ms.  The pair
        // halves the weight bytes each SM ingests but every accumulator hand-off crosses the cluster twice; it is kept as
        // an opt-in (AMP_B200_DISC_PAIR=1) and covered by the GPU tests.
        const char *pair = getenv("AMP_B200_DISC_PAIR");
        d->use_pair = pair && pair[0] == '1';
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

int64_t amp_disc_launch_count(const amp_disc_t *d, int64_t M) { return (d && M > 0) ? 2 * chunk_count(d, M) : 0; }

int amp_disc_load(amp_disc_t *d, const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                  const float *b3, const double *running_mean, const double *running_variance, void *stream) {
    AMP_REQUIRE(d && W1 && b1 && W2 && b2 && W3 && b3 && running_mean && running_variance, "amp_disc_load: NULL argument");
    cudaStream_t st = as_stream(stream);
    const int blocks = sm_count() * 4;
    cast_weight_kernel<<<blocks, 256, 0, st>>>(W1, d->h1, d->in_features, d->Kp, d->W1, d->fold_bias1 ? b1 : nullptr);
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

// Chunk plan of an M-row batch.  One chunk when M fits a workspace buffer, else equal big chunks; optionally a short LEAD-IN
// chunk first (AMP_B200_DISC_LEAD_TILES row tiles per CTA): the scaler/cast of chunk c+1 hides under the fused kernel of
// chunk c but the cast of chunk 0 has nothing to hide under.  Measured on the 1 M-row bench: 0 / 2 / 4 / 8 lead-in tiles ->
// 1.396 / 1.469 / 1.415 / 1.401 ms -- the extra launch (prologue + tail) costs what the smaller first cast saves, so the
// default is no lead-in.
struct ChunkPlan {
    int64_t n, lead, big;  // number of chunks, rows of the lead-in chunk (0 = none), rows of every following chunk
    int64_t start(int64_t c) const { return c == 0 ? 0 : (lead ? lead + (c - 1) * big : c * big); }
};
static ChunkPlan plan_chunks(const amp_disc *d, int64_t M) {
    ChunkPlan p{1, 0, d->chunk_rows};
    if (M <= d->chunk_rows) return p;
    const int64_t wave_rows = (int64_t)d->ws_ctas * BM;
    int64_t lead_tiles = 0;
    if (const char *t = getenv("AMP_B200_DISC_LEAD_TILES")) lead_tiles = std::max(0, atoi(t));  // tuning knob, 0 = no lead-in
    p.lead = std::min<int64_t>(lead_tiles * wave_rows, d->chunk_rows / 4) / wave_rows * wave_rows;
    const int64_t rest = M - p.lead;
    const int64_t k = (rest + d->chunk_rows - 1) / d->chunk_rows;
    p.big = std::min<int64_t>(d->chunk_rows, ((rest + k - 1) / k + wave_rows - 1) / wave_rows * wave_rows);
    p.n = (p.lead ? 1 : 0) + (rest + p.big - 1) / p.big;
    return p;
}

static int64_t chunk_count(const amp_disc *d, int64_t M) { return plan_chunks(d, M).n; }

// x rows are taken in order (row_index == NULL) or gathered: row r of the batch = x[row_index[r]] with x holding `capacity` rows
static int style_reward_impl(amp_disc_t *d, const float *x, int64_t x_stride, const int64_t *row_index, int64_t capacity,
                             uint32_t *flags, int64_t M, float reward_scale, float *reward, float *logits, void *stream) {
    AMP_REQUIRE(d && M >= 0, "amp_disc_style_reward: bad handle or negative size");
    AMP_REQUIRE(d->loaded, "amp_disc_style_reward: amp_disc_load has not been called");
    if (M == 0) return AMP_OK;
    AMP_REQUIRE(x && reward, "amp_disc_style_reward: NULL buffer");
    AMP_REQUIRE(x_stride >= d->in_features, "amp_disc_style_reward: x_stride %lld < in_features %d", (long long)x_stride,
                d->in_features);
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    const ChunkPlan plan = plan_chunks(d, M);
    const int64_t n_chunks = plan.n;
    AMP_REQUIRE(n_chunks == 1 || d->xhat[1], "amp_disc_style_reward: %lld rows exceed the max_rows given to amp_disc_create",
                (long long)M);
    // With several chunks the scaler/cast of chunk c+1 runs on the handle's side stream underneath the fused kernel of
    // chunk c (the fused kernel is operand-delivery bound and leaves HBM and most issue slots idle; the cast kernel uses no
    // shared memory, so its CTAs co-reside with the persistent fused CTAs).  A single chunk stays on the caller's stream.
    const bool overlap = n_chunks > 1;
    auto issue_cast = [&](int64_t c) -> int {
        const int b = (int)(c & 1);
        const int64_t r0 = plan.start(c), rows = std::min<int64_t>(plan.start(c + 1), M) - r0;
        cudaStream_t cs = overlap ? d->side : st;
        if (overlap && c >= 2) AMP_CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_free[b], 0));  // fused(c-2) is done with xhat[b]
        const float *xc = row_index ? x : x + r0 * x_stride;
        const int64_t *ic = row_index ? row_index + r0 : nullptr;
        const bool vec = (x_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(xc) & 7u) == 0);
        const int cast_grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
        const int ones = d->fold_bias1 ? 2 : 0;
        int rc = vec ? launch_normalise_cast<true>(d->Kp / BK, cast_grid, cs, xc, x_stride, rows, d->in_features, d->mean, d->denom,
                                                   d->xhat[b], ic, capacity, flags, ones)
                     : launch_normalise_cast<false>(d->Kp / BK, cast_grid, cs, xc, x_stride, rows, d->in_features, d->mean, d->denom,
                                                    d->xhat[b], ic, capacity, flags, ones);
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
        const int64_t r0 = plan.start(c), rows = std::min<int64_t>(plan.start(c + 1), M) - r0;
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
        fp.fold_bias1 = d->fold_bias1 ? 1 : 0;
        fp.ksteps1_last = (d->in_features + (d->fold_bias1 ? 2 : 0) - BK * (fp.kb1 - 1) + UMMA_K - 1) / UMMA_K;
        fp.n1_tiles = d->h1 / BN;
        fp.n2_tiles = d->h2 / BN;
        fp.b1 = d->b1;
        fp.b2 = d->b2;
        fp.w3 = d->w3;
        fp.b3 = d->b3;
        fp.scale = reward_scale;
        fp.reward = reward + r0;
        fp.logits = logits ? logits + r0 : nullptr;
        fp.prof = d->prof;
#ifdef AMP_DISC_PROFILE
        if (const char *m = getenv("AMP_DISC_PROF_MODE")) fp.prof_mode = atoi(m);
#endif
        if (d->use_pair && m_tiles >= 2) {
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
            AMP_CUDA_TRY(cudaLaunchKernelEx(&cfg, disc_fused_kernel<true>, tm_x, d->tmap_w1_half, d->tmap_h_load, d->tmap_h_store,
                                            d->tmap_w2_half, fp));
        } else {
            disc_fused_kernel<false><<<grid, FUSED_THREADS, fused_smem_bytes(false), st>>>(tm_x, d->tmap_w1, d->tmap_h_load,
                                                                                           d->tmap_h_store, d->tmap_w2, fp);
        }
        AMP_CUDA_TRY(cudaGetLastError());
        if (overlap) AMP_CUDA_TRY(cudaEventRecord(d->ev_free[b], st));
    }
#ifdef AMP_DISC_PROFILE
    if (d->prof) {  // developer build: dump the wait breakdown of the LAST chunk (synchronises!)
        static long long host[1024 * 16];
        cudaStreamSynchronize(st);
        const int n = std::min(d->ws_ctas, 1024);
        cudaMemcpy(host, d->prof, (size_t)n * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
        {  // epilogue warp 2 of every CTA (for a pair: even CTAs = leaders, odd = peers); assumes the grid used all ws_ctas CTAs
            double ea[2][8] = {{0}};
            for (int i = 0; i < n; ++i)
                for (int k = 0; k < 8; ++k) ea[d->use_pair ? (i & 1) : 0][k] += (double)host[n * 8 + i * 8 + k];
            const double div = d->use_pair ? n / 2.0 : (double)n;
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
            fprintf(stderr, "[amp_disc profile] issuer CTAs=%d total=%.0f wait_full=%.0f wait_d1_empty=%.0f wait_d2_empty=%.0f | producer(all) wait_empty=%.0f wait_h1=%.0f (cycles, mean per CTA) | issuer wall %.1f us => SM clock %.0f MHz\n",
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
