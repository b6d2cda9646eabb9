// Discriminator LOSS + GRADIENTS of the AMP update on the 5th-generation tensor cores.
//
//   amp_disc_train_*  <- skrl AMP._update, "compute discriminator loss" block + the backward pass of that loss (upstream
//                        skrl >= 1.4.3, not vendored; configured by the reference at agents/skrl_g1_dance_amp_cfg.yaml:31-39
//                        (MLP 1024-512-1, ReLU), :89 (loss scale 5), :94 (discriminator_batch_size 4096), :96-98 (logit
//                        regularisation 0.05, gradient penalty 5, weight decay 1e-4))
//
// The three batches (agent rollout, replay buffer, motion dataset; B rows each) sit in ONE row space of four blocks of
// Bp = B rounded up to 128 rows:   [agent | replay | motion | gradient-penalty rows of the motion block].
// Every matrix below has 4*Bp rows, so the two weight-gradient products absorb the gradient-penalty terms as extra K:
//
//   X   (4Bp, Kp)  bf16   rows <3Bp: normalised states x_hat          rows >=3Bp: G  = dL/dg (g = input gradient of d)
//   A1  (4Bp, h1)  bf16   rows <3Bp: a1 = relu(x_hat W1^T + b1)       rows >=3Bp: q1 = (G W1^T) * m1
//   A2  (3Bp, h2)  bf16   a2 = relu(a1 W2^T + b2)
//   E2  (4Bp, h2)  bf16   rows <3Bp: dz2 = dd * w3 * m2               rows >=3Bp: u2 = w3 * m2
//   E1  (4Bp, h1)  bf16   rows <3Bp: dz1 = (dz2 W2) * m1              rows >=3Bp: v1 = (u2 W2) * m1
//
//   dL/dW2 = E2^T A1          dL/dW1 = E1^T X          (both contract over all 4Bp rows; derivation: oracle/disc_train_oracle.py)
//
// Launch sequence of one step (9 GEMM launches of ONE templated tcgen05 kernel + 6 small kernels):
//   cast_transpose x2   W1, W2 -> bf16 (row-major and transposed) + sum of squares
//   GEMM  F1   A1[:3Bp] = relu(X[:3Bp] W1^T + b1)                     NT, bias+ReLU epilogue
//   GEMM  F2   A2       = relu(A1[:3Bp] W2^T + b2)                    NT, bias+ReLU
//   head       d = a2.w3 + b3, BCE terms, dd, E2 (dz2 and u2), dL/db2, dL/dw3 (BCE part), dL/db3
//   GEMM  B1   E1 = (E2 W2) * m1            (all 4Bp rows: dz1 and v1 in one launch)    NT on W2^T, mask epilogue
//   GEMM  GP2  G = c * (v1 W1), sum g^2     -> X[3Bp:]                                 NT on W1^T, scale+sumsq epilogue
//   GEMM  GP3  q1 = (G W1^T) * m1           -> A1[3Bp:]                                NT, mask epilogue
//   GEMM  GP6  S = (q1 W2^T) * m2                                                      NT, mask epilogue
//   GEMM  dW2  slices of E2^T A1            both operands MN-major (the batch dimension is K), split-K into slices
//   GEMM  dW1  slices of E1^T X
//   colsum x2  dL/db1 = colsum(E1[:3Bp]),  dL/dw3 += colsum(S)
//   finalize   sum the split-K slices, add weight decay / logit regularisation, write the fp32 gradients and loss terms
//
// Rows >= B inside a block are padding: the head kernel writes zero rows there, and zero rows of E2 stay zero through every
// later product, so no GEMM needs a row bound.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "amp_bucket.cuh"
#include "amp_internal.h"
#include "amp_math.cuh"
#include "amp_tc.cuh"

namespace amp {
namespace train {

using namespace amp::tc;

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = kBlockK;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int MN_BOX_BYTES = BK * 64 * 2;   // one MN-major TMA box: 64 K-rows x 64 columns (128-byte swizzled rows) = 8 KiB
constexpr int GEMM_THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int SLAB_BYTES = 32 * 64 * 2;     // per epilogue warp: 32 rows x 64 bf16 columns, staged for coalesced stores
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 8 * SLAB_BYTES + 256 + 1024;

enum Epilogue { EPI_BIAS_RELU = 0, EPI_MASK = 1, EPI_SCALE_SUMSQ = 2, EPI_SLICE = 3 };

// MN-major operand in the canonical 128B-swizzle layout: a [64 K-rows x 64 MN-elements] TMA box puts K-row k at byte k*128
// with the 16-byte chunks XOR-swizzled by (k % 8).  Canonical form (CUTLASS make_umma_desc<Major::MN>, SW128, units of
// 16 bytes): ((8,n),(8,k)) : ((1,LBO),(8,SBO)) -- 64 MN-elements contiguous, 8 K-rows 128 B apart form one swizzle atom,
// SBO = distance between 8-row K groups (1024 B), LBO = distance between 64-element MN blocks (one box = 8192 B).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(MN_BOX_BYTES >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct GemmParams {
    int m_tiles, n_tiles, splits, kblocks;  // work items = splits x m_tiles x n_tiles; kblocks = K / 64 over all splits
    int n_valid;                            // output columns >= n_valid do not exist (multiple of 64)
    const float *bias;                      // EPI_BIAS_RELU
    const float *dot_w;                     // EPI_BIAS_RELU, optional: dot_out[row] += sum_col relu(acc + bias)[col] * dot_w[col]
    float *dot_out;                         //   (fp32 atomics, one per epilogue warp-row and tile; the last layer of the MLP)
    __nv_bfloat16 *out;                     // bf16 output (all but EPI_SLICE)
    long long out_ld;
    const __nv_bfloat16 *mask;              // EPI_MASK: out = mask[mrow, col] != 0 ? alpha * acc : 0
    long long mask_ld;
    int mask_split, mask_shift;             // mrow = row < mask_split ? row : row - mask_shift
    float alpha;                            // EPI_MASK / EPI_SCALE_SUMSQ
    float *sumsq;                           // EPI_SCALE_SUMSQ: atomicAdd(sumsq, sum acc^2)
    float *slices;                          // EPI_SLICE: fp32 (splits, m_tiles*128, slice_ld)
    long long slice_ld, slice_stride;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&v);
}

// One persistent launch: D[128 x 256] tiles of  A (M x K)  times  B (N x K)^T, bf16 operands, fp32 accumulation in TMEM.
//   A_MN / B_MN = false: operand stored K-major (row = M/N index, K contiguous)   -> TMA box [128|256 rows x 64 K]
//               = true : operand stored MN-major (row = K index, M/N contiguous)  -> TMA boxes [64 K-rows x 64 columns]
// Warp roles (320 threads): 0 TMA producer, 1 TMEM alloc + single-thread MMA issuer, 2..9 epilogue (warps w and w+4 share a
// TMEM lane quarter and split the 256 accumulator columns).  Accumulators ping-pong between the two halves of TMEM so the
// drain of item i overlaps the MMAs of item i+1.
template <bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
train_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + STAGES * A_STAGE_BYTES;
    const uint32_t slabs = smem_b + STAGES * B_STAGE_BYTES;
    const uint32_t bars = slabs + 8 * SLAB_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
    const uint32_t acc_full = bars + 16 * STAGES, acc_empty = acc_full + 16;
    const uint32_t tmem_slot = acc_full + 32;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.m_tiles * p.n_tiles;
    const int items = tiles * p.splits;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full_bar + 8 * i, 1);
            mbar_init(empty_bar + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, 8);  // one arrival per epilogue warp
        }
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
            const uint64_t keep = l2_policy_evict_last();  // every operand of the step is re-read by other CTAs / later launches
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int s = item / tiles, t = item - s * tiles;
                const int m = t / p.n_tiles, n = t - m * p.n_tiles;
                const int kb0 = (int)((long long)s * p.kblocks / p.splits), kb1 = (int)((long long)(s + 1) * p.kblocks / p.splits);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t fb = full_bar + 8 * stage;
                    mbar_arrive_expect_tx(fb, A_STAGE_BYTES + B_STAGE_BYTES);
                    const uint32_t sa = smem_a + stage * A_STAGE_BYTES, sb = smem_b + stage * B_STAGE_BYTES;
                    if constexpr (A_MN) {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d_hint(sa + j * MN_BOX_BYTES, &tmap_a, m * BM + j * 64, kb * BK, fb, keep);
                    } else {
                        tma_load_2d_hint(sa, &tmap_a, kb * BK, m * BM, fb, keep);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_2d_hint(sb + j * MN_BOX_BYTES, &tmap_b, n * BN + j * 64, kb * BK, fb, keep);
                    } else {
                        tma_load_2d_hint(sb, &tmap_b, kb * BK, n * BN, fb, keep);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // kind::f16 instruction descriptor: D fp32 (bit 4), A/B bf16 (bits 7, 10), A/B major (bits 15, 16; 1 = MN-major),
            // N >> 3 at bits [17,23), M >> 4 at bits [24,29)
            constexpr uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                                            ((uint32_t)(BM >> 4) << 24);
            // descriptor start-address step (units of 16 B) per UMMA_K = 16: K-major 32 B inside the swizzled row;
            // MN-major 16 K-rows x 128 B = 2048 B
            constexpr uint32_t a_step = A_MN ? (16 * 128) >> 4 : 2;
            constexpr uint32_t b_step = B_MN ? (16 * 128) >> 4 : 2;
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
                const int s = item / tiles, n = (item - s * tiles) % p.n_tiles;
                const int kb0 = (int)((long long)s * p.kblocks / p.splits), kb1 = (int)((long long)(s + 1) * p.kblocks / p.splits);
                // the last N tile may be narrower than 256 columns (n_valid = 832: 256, 256, 256, 64): issue only the columns
                // that exist (UMMA N is any multiple of 16); the TMA boxes past n_valid are zero-filled and simply not read
                const uint32_t idesc = idesc_base | ((uint32_t)(min(BN, p.n_valid - n * BN) >> 3) << 17);
                const uint32_t r = it & 1, use = it >> 1;
                mbar_wait(acc_empty + 8 * r, (use & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + r * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_a + stage * A_STAGE_BYTES, sb = smem_b + stage * B_STAGE_BYTES;
                    const uint64_t a0 = A_MN ? make_mnmajor_sw128_desc(sa) : make_kmajor_sw128_desc(sa);
                    const uint64_t b0 = B_MN ? make_mnmajor_sw128_desc(sb) : make_kmajor_sw128_desc(sb);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(d_tmem, a0 + a_step * k, b0 + b_step * k, idesc, (uint32_t)(kb != kb0 || k != 0));
                    umma_commit(empty_bar + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(acc_full + 8 * r);
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps 2..9 =================
        const int quarter = warp & 3;
        const int colhalf = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        [[maybe_unused]] const uint32_t slab = slabs + (uint32_t)((warp - 2) * SLAB_BYTES);
        uint32_t it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
            const int s = item / tiles, t = item - s * tiles;
            const int m = t / p.n_tiles, n = t - m * p.n_tiles;
            const uint32_t r = it & 1, use = it >> 1;
            [[maybe_unused]] const long long row = (long long)m * BM + quarter * 32 + lane;
            const int col0 = n * BN + colhalf * (BN / 2);
            const uint32_t acc = tmem_base + lane_base + r * BN + (uint32_t)(colhalf * (BN / 2));
            const int chunks = min(4, max(0, (p.n_valid - col0 + 31) / 32));  // warp-uniform; n_valid % 64 == 0: 0, 2 or 4
            // bf16 epilogues go through a per-warp shared-memory slab (32 rows x 64 columns, 16-byte pieces XOR-swizzled by
            // row): the TMEM read-out gives a thread one ROW, but global memory wants 8 lanes on the 128 contiguous bytes of a
            // row.  In the second pass lane l handles 16-byte piece (l & 7) of rows (l >> 3) + 4 i: every row segment is
            // one full 128-byte line (4 lines per instruction instead of 32).  (Direct stores from the row-per-thread layout:
            // the B1 launch spent 8 us per tile in this epilogue against 2 us of MMAs.)
            // EPI_MASK: the mask (activations of an earlier launch) does not depend on this accumulator -- fetch it in the
            // second-pass layout before waiting for the MMAs, so the global-memory latency hides under them.
            const long long row_base = (long long)m * BM + quarter * 32;
            const int sub_row = lane >> 3, piece = lane & 7;
            [[maybe_unused]] uint4 mk[2][8];
            if constexpr (EPI == EPI_MASK) {
                const long long mbase = row_base < p.mask_split ? row_base : row_base - p.mask_shift;  // tile-uniform side
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (2 * g >= chunks) break;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        mk[g][i] = __ldg(reinterpret_cast<const uint4 *>(p.mask + (mbase + sub_row + 4 * i) * p.mask_ld + col0 + g * 64) + piece);
                }
            }
            mbar_wait(acc_full + 8 * r, use & 1);
            tcgen05_fence_after();
            [[maybe_unused]] float ss = 0.0f;
            [[maybe_unused]] float dot = 0.0f;
            [[maybe_unused]] const bool with_dot = EPI == EPI_BIAS_RELU && p.dot_w != nullptr;  // launch-uniform
            uint32_t v[2][32];
            if (chunks > 0) tmem_ld_32x32(acc, v[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c >= chunks) break;
                const int col = col0 + c * 32;
                tmem_ld_wait();
                if (c + 1 < chunks) tmem_ld_32x32(acc + (uint32_t)((c + 1) * 32), v[(c + 1) & 1]);
                const uint32_t(&cur)[32] = v[c & 1];
                if constexpr (EPI == EPI_SLICE) {
                    float4 *dst = reinterpret_cast<float4 *>(p.slices + (long long)s * p.slice_stride + row * p.slice_ld + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1]), __uint_as_float(cur[4 * j + 2]),
                                             __uint_as_float(cur[4 * j + 3]));
                } else {
                    uint32_t o[16];
                    if constexpr (EPI == EPI_BIAS_RELU) {
                        const float4 *bias4 = reinterpret_cast<const float4 *>(p.bias + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(bias4 + j);
                            const float h0 = max_nan(__uint_as_float(cur[4 * j]) + b.x, 0.0f), h1 = max_nan(__uint_as_float(cur[4 * j + 1]) + b.y, 0.0f);
                            const float h2 = max_nan(__uint_as_float(cur[4 * j + 2]) + b.z, 0.0f), h3 = max_nan(__uint_as_float(cur[4 * j + 3]) + b.w, 0.0f);
                            o[2 * j] = pack_bf16(h0, h1);
                            o[2 * j + 1] = pack_bf16(h2, h3);
                            if (with_dot) {
                                const float4 w = __ldg(reinterpret_cast<const float4 *>(p.dot_w + col) + j);
                                dot = fmaf(h0, w.x, dot); dot = fmaf(h1, w.y, dot); dot = fmaf(h2, w.z, dot); dot = fmaf(h3, w.w, dot);
                            }
                        }
                    } else {  // EPI_MASK (mask applied in the second pass), EPI_SCALE_SUMSQ
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float a = __uint_as_float(cur[2 * j]), b = __uint_as_float(cur[2 * j + 1]);
                            if constexpr (EPI == EPI_SCALE_SUMSQ) {
                                ss = fmaf(a, a, ss);
                                ss = fmaf(b, b, ss);
                            }
                            o[j] = pack_bf16(a * p.alpha, b * p.alpha);
                        }
                    }
                    // first pass: this thread's 32 columns = pieces (c & 1) * 4 + j of its row in the slab
                    const uint32_t my_row = slab + (uint32_t)(lane * 128);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4(my_row + (uint32_t)((((c & 1) * 4 + j) ^ (lane & 7)) * 16), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    if (c & 1) {  // the 64-column group is complete: second pass
                        __syncwarp();
                        const int g = c >> 1;
                        __nv_bfloat16 *dst = p.out + (row_base + sub_row) * p.out_ld + col0 + g * 64 + piece * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = sub_row + 4 * i;
                            uint4 q;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                                         : "r"(slab + (uint32_t)(rr * 128 + ((piece ^ (rr & 7)) * 16))));
                            if constexpr (EPI == EPI_MASK) {
                                const uint4 mm = mk[g][i];
                                q.x &= ((mm.x & 0xffffu) ? 0xffffu : 0u) | ((mm.x >> 16) ? 0xffff0000u : 0u);
                                q.y &= ((mm.y & 0xffffu) ? 0xffffu : 0u) | ((mm.y >> 16) ? 0xffff0000u : 0u);
                                q.z &= ((mm.z & 0xffffu) ? 0xffffu : 0u) | ((mm.z >> 16) ? 0xffff0000u : 0u);
                                q.w &= ((mm.w & 0xffffu) ? 0xffffu : 0u) | ((mm.w >> 16) ? 0xffff0000u : 0u);
                            }
                            *reinterpret_cast<uint4 *>(dst + (long long)(4 * i) * p.out_ld) = q;
                        }
                        __syncwarp();  // the slab is rewritten by the next group
                    }
                }
            }
            // all TMEM reads of this accumulator by this warp have completed (tcgen05.wait::ld above): release the region
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8 * r);
            if constexpr (EPI == EPI_SCALE_SUMSQ) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (lane == 0 && chunks > 0) atomicAdd(p.sumsq, ss);
            }
            if constexpr (EPI == EPI_BIAS_RELU) {
                if (with_dot && chunks > 0) atomicAdd(p.dot_out + row, dot);
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

// fp32 master (R, C) -> bf16 (R, Cp) zero padded, its transpose bf16 (Cp, R), and sum of squares (atomicAdd into *sumsq).
// 32 x 32 tiles through shared memory: both stores are coalesced.  R % 32 == 0, Cp % 32 == 0.
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float *__restrict__ w, int R, int C, int Cp,
                                                              __nv_bfloat16 *__restrict__ out, __nv_bfloat16 *__restrict__ out_t,
                                                              float *__restrict__ sumsq) {
    __shared__ float tile[32][33];
    __shared__ float part[8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        const float x = c < C ? w[(long long)r * C + c] : 0.0f;
        ss = fmaf(x, x, ss);
        tile[ty + 8 * i][tx] = x;
        out[(long long)r * Cp + c] = __float2bfloat16_rn(x);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        out_t[(long long)c * R + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (tx == 0) part[ty] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int i = 0; i < 8; ++i) t += part[i];
        atomicAdd(sumsq, t);
    }
}

// Column sums arrive by fp32 atomics from a few hundred CTAs at once; same-address atomics serialise in L2, so each sum has
// REPL copies (CTA b adds into copy b % REPL) that finalize_kernel folds.
constexpr int REPL = 8;

// scalar accumulators at the head of the fp32 workspace (zeroed at the start of every step)
enum Acc { ACC_BCE_CAT = 0, ACC_BCE_MOTION, ACC_W3_SQ, ACC_GP_SQ, ACC_W_SQ, ACC_GB3, ACC_COUNT = 8 };

__device__ __forceinline__ float softplus_f(float x) { return max_nan(x, 0.0f) + log1pf(__expf(-fabsf(x))); }

// After the second layer: logits, BCE terms, dL/dd, and everything that hangs off it row by row.
//   A block takes 32 rows: the a2 tile (32 x h2 bf16) is read ONCE into shared memory; phase A: one lane per row ->
//   d = dotsum + b3 (dotsum = a2 . w3 from the layer-2 epilogue) and dd; phase B: one thread per column pair walks the 32 rows: E2[row] = dd * w3 * m2 (and, for motion
//   rows, E2[row + Bp] = w3 * m2), column sums dL/db2 and sum dd * a2.  Dynamic shared memory: 32 * h2 * 2 bytes.
__global__ void __launch_bounds__(256) head_kernel(const __nv_bfloat16 *__restrict__ a2, const float *__restrict__ dotsum, int h2,
                                                   int Bp, int B, const float *__restrict__ w3, const float *__restrict__ b3, float loss_scale,
                                                   __nv_bfloat16 *__restrict__ e2, float *__restrict__ acc, float *__restrict__ gw3,
                                                   float *__restrict__ gb2, float *__restrict__ logits) {
    extern __shared__ __align__(16) unsigned char head_smem[];
    __nv_bfloat16 *tile = reinterpret_cast<__nv_bfloat16 *>(head_smem);
    __shared__ float s_dd[32];
    __shared__ float s_loss[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row0 = (long long)blockIdx.x * 32;
    const int block_src = (int)(row0 / Bp);  // 32 | Bp: a block never straddles two sources
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a2 + row0 * h2);
        uint4 *dst = reinterpret_cast<uint4 *>(tile);
        const int quads = 32 * h2 / 8;  // h2 % 256 == 0: a multiple of 4 * 256
        for (int i = threadIdx.x; i < quads; i += 4 * 256) {  // four loads in flight per thread
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = __ldg(src + i + u * 256);
#pragma unroll
            for (int u = 0; u < 4; ++u) dst[i + u * 256] = q[u];
        }
    }
    __syncthreads();
    if (warp == 0) {  // the logits were accumulated (fp32, before the bf16 rounding of a2) by the epilogue of the layer-2 GEMM
        const long long row = row0 + lane;
        const bool valid = (int)(row - (long long)block_src * Bp) < B;
        const float d = dotsum[row] + __ldg(b3);
        if (logits) logits[row] = d;
        const float sig = 1.0f / (1.0f + __expf(-d));
        float dd, loss;
        if (block_src < 2) {  // agent / replay: target 0, mean over 2B rows
            dd = 0.5f * loss_scale * sig / (2.0f * (float)B);
            loss = softplus_f(d);
        } else {              // motion: target 1, mean over B rows
            dd = 0.5f * loss_scale * (sig - 1.0f) / (float)B;
            loss = softplus_f(-d);
        }
        s_dd[lane] = valid ? dd : 0.0f;
        s_loss[lane] = valid ? loss : 0.0f;
    }
    __syncthreads();
    if (warp == 0) {
        float l = s_loss[lane], g = s_dd[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l += __shfl_xor_sync(0xffffffffu, l, o);
            g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        if (lane == 0) {
            atomicAdd(acc + (block_src < 2 ? ACC_BCE_CAT : ACC_BCE_MOTION), l);
            atomicAdd(acc + ACC_GB3, g);
        }
    }
    const bool motion = block_src == 2;
    gb2 += (blockIdx.x % REPL) * h2;
    gw3 += (blockIdx.x % REPL) * h2;
    for (int c = threadIdx.x * 2; c < h2; c += 512) {
        const float2 w = __ldg(reinterpret_cast<const float2 *>(w3 + c));
        float sb0 = 0.0f, sb1 = 0.0f, sw0 = 0.0f, sw1 = 0.0f;
#pragma unroll 8
        for (int rl = 0; rl < 32; ++rl) {
            const long long row = row0 + rl;
            const float dd = s_dd[rl];
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(tile + rl * h2 + c));
            const float m0 = a.x > 0.0f ? 1.0f : 0.0f, m1 = a.y > 0.0f ? 1.0f : 0.0f;
            const float z0 = dd * w.x * m0, z1 = dd * w.y * m1;
            sb0 += z0; sb1 += z1;
            sw0 = fmaf(dd, a.x, sw0); sw1 = fmaf(dd, a.y, sw1);
            *reinterpret_cast<uint32_t *>(e2 + row * h2 + c) = pack_bf16(z0, z1);
            if (motion) {
                const bool valid = (int)(row - 2LL * Bp) < B;
                *reinterpret_cast<uint32_t *>(e2 + (row + Bp) * h2 + c) = valid ? pack_bf16(w.x * m0, w.y * m1) : 0u;
            }
        }
        atomicAdd(gb2 + c, sb0); atomicAdd(gb2 + c + 1, sb1);
        atomicAdd(gw3 + c, sw0); atomicAdd(gw3 + c + 1, sw1);
    }
}

// Staging of one source: RunningStandardScaler (eval form, skrl) + bf16 cast, one thread per output column pair -- the
// batches of a discriminator update are a few thousand rows, so the row-per-warp kernel of the style-reward path (built for
// 1 M rows, statistics cached in registers) would be one long latency chain here.  Same arithmetic as that kernel:
// clamp((x - (float)mean) * rcp(sqrt((float)var) + 1e-8), -5, 5).  mean == NULL: x is already normalised.
// Column `in_features` (when Kp > in_features) is set to 1: with that "ones" column the product E1^T X delivers the bias
// gradient dL/db1 = colsum(E1) as column in_features of dL/dW1 (W1's own padding columns are zero, so nothing else sees it).
constexpr int STAGE_ROWS = 32;  // rows per CTA of stage_cast_kernel
__global__ void __launch_bounds__(256) stage_cast_kernel(const float *__restrict__ x, long long x_stride, int rows, int in_features,
                                                         int Kp, const double *__restrict__ mean, const double *__restrict__ var,
                                                         __nv_bfloat16 *__restrict__ out) {
    // blockIdx.x: strip of 512 columns (one column pair per thread); blockIdx.y: strip of STAGE_ROWS rows.  The float64
    // statistics of the thread's two columns are narrowed ONCE (fp64 conversions run at a small fraction of the fp32 rate:
    // converting per element made this kernel 4x slower than its memory traffic).
    const int c = blockIdx.x * 512 + threadIdx.x * 2;
    if (c >= Kp) return;
    float mu[2], rc[2];
    bool live[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int cc = c + j;
        live[j] = cc < in_features;
        mu[j] = live[j] && mean ? __double2float_rn(__ldg(mean + cc)) : 0.0f;
        rc[j] = live[j] ? __frcp_rn(var ? __fadd_rn(__fsqrt_rn(__double2float_rn(__ldg(var + cc))), 1e-8f) : 1.0f) : 0.0f;
    }
    const float pad0 = c == in_features ? 1.0f : 0.0f, pad1 = c + 1 == in_features ? 1.0f : 0.0f;
    const int r0 = blockIdx.y * STAGE_ROWS, r1 = min(rows, r0 + STAGE_ROWS);
    // eight rows of loads are issued before the first is consumed (the ld.global.cs intrinsics keep program order)
    for (int rb = r0; rb < r1; rb += 8) {
        float x0[8], x1[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float *xr = x + (long long)min(rb + u, r1 - 1) * x_stride + c;
            x0[u] = live[0] ? __ldcs(xr) : 0.0f;
            x1[u] = live[1] ? __ldcs(xr + 1) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (rb + u >= r1) break;
            const float v0 = live[0] ? clamp_nan(__fmul_rn(__fsub_rn(x0[u], mu[0]), rc[0]), -5.0f, 5.0f) : pad0;
            const float v1 = live[1] ? clamp_nan(__fmul_rn(__fsub_rn(x1[u], mu[1]), rc[1]), -5.0f, 5.0f) : pad1;
            *reinterpret_cast<uint32_t *>(out + (long long)(rb + u) * Kp + c) = pack_bf16(v0, v1);
        }
    }
}

// out[b % REPL][c] += sum over rows of src[r, c] (bf16 -> fp32): block b takes 16 rows x 512 columns, one thread per column pair
constexpr int COLSUM_ROWS = 16;
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16 *__restrict__ src, long long ld, int rows, int cols,
                                                     float *__restrict__ out) {
    const int c = blockIdx.x * 512 + threadIdx.x * 2;
    if (c >= cols) return;
    const int r0 = blockIdx.y * COLSUM_ROWS, r1 = min(rows, r0 + COLSUM_ROWS);
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 16
    for (int r = r0; r < r1; ++r) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(src + (long long)r * ld + c));
        s0 += a.x; s1 += a.y;
    }
    out += (blockIdx.y % REPL) * (long long)cols;
    atomicAdd(out + c, s0);
    atomicAdd(out + c + 1, s1);
}

struct FinalizeParams {
    const float *W1, *W2, *w3;     // fp32 masters
    const float *slices1, *slices2;
    int splits1, splits2, in_features, Kp, h1, h2, B;
    const float *acc;              // scalar accumulators
    const float *ws_gw3, *ws_gb2, *ws_gb1;
    float loss_scale, c_reg, c_gp, c_wd;
    float *gW1, *gb1, *gW2, *gb2, *gW3, *gb3, *terms;
};

// Deterministic split-K reduction + the closed-form regularisation gradients; writes the six fp32 gradient tensors (torch
// layout) and terms[6] = {bce_agent_replay, bce_motion, logit_regularization, gradient_penalty, weight_decay, loss}.
__global__ void __launch_bounds__(256) finalize_kernel(FinalizeParams p) {
    const long long n1 = (long long)p.h1 * p.in_features, n2 = (long long)p.h2 * p.h1;
    const float cw = 2.0f * p.loss_scale * p.c_wd;
    const long long stride = (long long)gridDim.x * blockDim.x, tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long s1 = (long long)p.h1 * p.Kp, s2 = n2;
    if ((p.in_features & 1) == 0) {  // rows of gW1 / W1 are 8-byte aligned: two columns per thread
        const int half = p.in_features / 2;
        for (long long e = tid; e < n1 / 2; e += stride) {
            const int r = (int)(e / half), c = 2 * (int)(e - (long long)r * half);
            float2 g = make_float2(0.0f, 0.0f);
            for (int s = 0; s < p.splits1; ++s) {
                const float2 v = *reinterpret_cast<const float2 *>(p.slices1 + s * s1 + (long long)r * p.Kp + c);
                g.x += v.x; g.y += v.y;
            }
            const float2 w = *reinterpret_cast<const float2 *>(p.W1 + 2 * e);
            *reinterpret_cast<float2 *>(p.gW1 + 2 * e) = make_float2(fmaf(cw, w.x, g.x), fmaf(cw, w.y, g.y));
        }
    } else {
        for (long long e = tid; e < n1; e += stride) {
            const int r = (int)(e / p.in_features), c = (int)(e - (long long)r * p.in_features);
            float g = 0.0f;
            for (int s = 0; s < p.splits1; ++s) g += p.slices1[s * s1 + (long long)r * p.Kp + c];
            p.gW1[e] = fmaf(cw, p.W1[e], g);
        }
    }
    for (long long e = tid; e < n2 / 4; e += stride) {  // h1 % 256 == 0: every row of W2 is 16-byte aligned
        float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int s = 0; s < p.splits2; ++s) {
            const float4 v = *reinterpret_cast<const float4 *>(p.slices2 + s * s2 + 4 * e);
            g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
        }
        const float4 w = *reinterpret_cast<const float4 *>(p.W2 + 4 * e);
        *reinterpret_cast<float4 *>(p.gW2 + 4 * e) = make_float4(fmaf(cw, w.x, g.x), fmaf(cw, w.y, g.y), fmaf(cw, w.z, g.z), fmaf(cw, w.w, g.w));
    }
    for (long long e = tid; e < p.h2; e += stride) {
        float gw = 0.0f, gb = 0.0f;
#pragma unroll
        for (int k = 0; k < REPL; ++k) {
            gw += p.ws_gw3[k * p.h2 + e];
            gb += p.ws_gb2[k * p.h2 + e];
        }
        p.gW3[e] = fmaf(cw + 2.0f * p.loss_scale * p.c_reg, p.w3[e], gw);
        p.gb2[e] = gb;
    }
    for (long long e = tid; e < p.h1; e += stride) {
        if (p.Kp > p.in_features) {  // the "ones" column of X (stage_cast_kernel): column in_features of E1^T X is colsum(E1)
            float g = 0.0f;
            for (int s = 0; s < p.splits1; ++s) g += p.slices1[s * s1 + e * p.Kp + p.in_features];
            p.gb1[e] = g;
        } else {
            float g = 0.0f;
#pragma unroll
            for (int k = 0; k < REPL; ++k) g += p.ws_gb1[k * p.h1 + e];
            p.gb1[e] = g;
        }
    }
    if (tid == 0) {
        p.gb3[0] = p.acc[ACC_GB3];
        const float bce_cat = p.acc[ACC_BCE_CAT] / (2.0f * (float)p.B), bce_motion = p.acc[ACC_BCE_MOTION] / (float)p.B;
        const float reg = p.acc[ACC_W3_SQ], gp = p.acc[ACC_GP_SQ] / (float)p.B, wd = p.acc[ACC_W_SQ] + p.acc[ACC_W3_SQ];
        if (p.terms) {
            p.terms[0] = bce_cat; p.terms[1] = bce_motion; p.terms[2] = reg; p.terms[3] = gp; p.terms[4] = wd;
            p.terms[5] = p.loss_scale * (0.5f * (bce_cat + bce_motion) + p.c_reg * reg + p.c_gp * gp + p.c_wd * wd);
        }
    }
}

// ---- finalize + gradient exchange in ONE kernel (second half) ------------------------------------
// With more than one rank the six gradient tensors live side by side in the rank's gradient bucket (amp_bucket.cu) and what the
// optimiser needs is their MEAN over the ranks (skrl Model.reduce_parameters, enabled by the reference at train.py:53-58,
// 184-196).  finalize_kernel followed by the all-reduce kernel writes the local gradients to the bucket, reads them back and
// pulls every peer's copy over NVLink (a round trip per load).  This kernel is the reduce-scatter / all-gather pair with the
// split-K reduction as its first stage:
//
//   push       every rank evaluates quad q of its local gradients (the arithmetic of finalize_kernel, same order) and STORES it
//              straight into the staging area of the rank that owns q (peer stores, fire and forget) -- the local bucket is not
//              written, nothing is read back
//   barrier A  "all my pushes are out" (last CTA of the rank -> one flag per peer); wait for every rank's flag
//   reduce     the owner adds the W staged copies of its slice (local loads, fixed rank order -> the same bits on every rank,
//              and the same bits as finalize + all-reduce), scales by 1/W and stores the result into every rank's bucket
//   barrier B  as in the all-reduce kernels: the stream continues only when every peer's stores have been announced
//
// The quads of the exchanged range map to (tensor, element) through `seg`: the six tensors may sit in any order inside the
// range as long as each starts on a quad (only the one-element dL/db3 is ragged, so it must come last).  All CTAs of the grid
// must be resident at once (every CTA waits for flags that the peers' LAST CTAs send): the host sizes the grid from the
// occupancy calculator.
#ifndef AMP_EXCHANGE_STAMPS
#define AMP_EXCHANGE_STAMPS 0  // developer builds: 1 = stamps inside the push phase, 2 = inside the reduce phase (tools/bench_fused_exchange.py)
#endif
struct ExchangeParams {
    bucket::Peers peers;
    float *mc;             // multicast mapping of every rank's bucket (shared buckets whose all-reduce runs in the switch), or NULL
    int rank, world;
    long long hull;        // float offset of the exchanged range inside the bucket (multiple of 4)
    long long count;       // floats in the range (all six tensors)
    long long stage;       // float offset of the staging area from the bucket base
    long long seg[6];      // range-relative float offset of gW1, gb1, gW2, gb2, gW3, gb3
    bucket::Control *ctl;
    volatile uint32_t *host_status;
    long long spin_limit;
    unsigned long long *timing;
};

// quad `i4 / 4` of the local gradients; i4 = range-relative float index of its first element (a multiple of 4)
__device__ __forceinline__ float4 gradient_quad(const FinalizeParams &p, const ExchangeParams &x, long long i4) {
    const long long n1 = (long long)p.h1 * p.in_features, n2 = (long long)p.h2 * p.h1;
    const long long s1 = (long long)p.h1 * p.Kp, s2 = n2;
    const float cw = 2.0f * p.loss_scale * p.c_wd;
    float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    long long e = i4 - x.seg[0];
    if (e >= 0 && e < n1) {  // dL/dW1: split-K slices (h1, Kp) -> torch layout (h1, in_features); a quad may cross a row end
        int r = (int)e / p.in_features, c = (int)e - r * p.in_features;  // n1 <= 2048 * 1024: 32-bit arithmetic
        if ((p.in_features & 1) == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float2 g = make_float2(0.0f, 0.0f);
                for (int s = 0; s < p.splits1; ++s) {
                    const float2 t = *reinterpret_cast<const float2 *>(p.slices1 + s * s1 + (long long)r * p.Kp + c);
                    g.x += t.x; g.y += t.y;
                }
                const float2 w = *reinterpret_cast<const float2 *>(p.W1 + e + 2 * h);
                v[2 * h] = fmaf(cw, w.x, g.x);
                v[2 * h + 1] = fmaf(cw, w.y, g.y);
                c += 2;
                if (c >= p.in_features) { c = 0; ++r; }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float g = 0.0f;
                for (int s = 0; s < p.splits1; ++s) g += p.slices1[s * s1 + (long long)r * p.Kp + c];
                v[j] = fmaf(cw, p.W1[e + j], g);
                if (++c >= p.in_features) { c = 0; ++r; }
            }
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    }
    e = i4 - x.seg[2];
    if (e >= 0 && e < n2) {  // dL/dW2
        float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int s = 0; s < p.splits2; ++s) {
            const float4 t = *reinterpret_cast<const float4 *>(p.slices2 + s * s2 + e);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        const float4 w = *reinterpret_cast<const float4 *>(p.W2 + e);
        return make_float4(fmaf(cw, w.x, g.x), fmaf(cw, w.y, g.y), fmaf(cw, w.z, g.z), fmaf(cw, w.w, g.w));
    }
    e = i4 - x.seg[1];
    if (e >= 0 && e < p.h1) {  // dL/db1
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g = 0.0f;
            if (p.Kp > p.in_features) {
                for (int s = 0; s < p.splits1; ++s) g += p.slices1[s * s1 + (e + j) * p.Kp + p.in_features];
            } else {
#pragma unroll
                for (int k = 0; k < REPL; ++k) g += p.ws_gb1[k * p.h1 + e + j];
            }
            v[j] = g;
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    }
    e = i4 - x.seg[3];
    if (e >= 0 && e < p.h2) {  // dL/db2
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g = 0.0f;
#pragma unroll
            for (int k = 0; k < REPL; ++k) g += p.ws_gb2[k * p.h2 + e + j];
            v[j] = g;
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    }
    e = i4 - x.seg[4];
    if (e >= 0 && e < p.h2) {  // dL/dw3
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g = 0.0f;
#pragma unroll
            for (int k = 0; k < REPL; ++k) g += p.ws_gw3[k * p.h2 + e + j];
            v[j] = fmaf(cw + 2.0f * p.loss_scale * p.c_reg, p.w3[e + j], g);
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    }
    return make_float4(p.acc[ACC_GB3], 0.0f, 0.0f, 0.0f);  // dL/db3 (the ragged last quad)
}

// SWITCH = true (shared buckets): the push phase writes the local bucket instead, and the reduce phase is the in-switch pair of
// allreduce_mean_switch_kernel (multimem.ld_reduce / multimem.st on the multicast mapping) -- finalize_kernel and that kernel in
// one launch, with the switch's sum.
template <int MAXW, bool SWITCH>
__global__ void __launch_bounds__(256, 4) finalize_exchange_kernel(FinalizeParams p, ExchangeParams x) {
    using namespace amp::bucket;
    __shared__ bool ok, last;
    __shared__ uint32_t s_epoch;
    const int rank = x.rank, world = x.world;
    uint32_t *local_flags = x.peers.flags[rank];
    unsigned long long t_start = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    // the epoch is device state advanced by the last CTA of every call on this bucket (graph replays stay in step)
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t *>(&x.ctl->epoch) + 1;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const long long quads = (x.count + 3) / 4, per = (quads + world - 1) / world;
    const long long stride = (long long)gridDim.x * blockDim.x, tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // ---- push: local gradients -> the owners' staging (slot `rank` of the owner's W slots of `per` quads) ----
    // No barrier in front: the peers stopped reading their staging before they passed barrier B of the previous call, and this
    // rank passed that barrier before its stream reached this launch.
    float *bucket_local = x.peers.data[rank];
    for (long long q = tid; q < quads; q += stride) {
        const float4 g = gradient_quad(p, x, 4 * q);
        if (SWITCH) {
            float *dst = bucket_local + x.hull + 4 * q;
            if (x.count - 4 * q >= 4) *reinterpret_cast<float4 *>(dst) = g;
            else dst[0] = g.x;  // the ragged last quad holds dL/db3 alone
        } else {
            const int owner = (int)q / (int)per;  // the range is a few million floats: 32-bit arithmetic
            st_peer(reinterpret_cast<float4 *>(x.peers.data[owner] + x.stage) + (long long)rank * per + (q - owner * per), g);
        }
    }
#if AMP_EXCHANGE_STAMPS == 1
    if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); x.timing[0] = t_start; x.timing[1] = t; }
#endif
    if (tid == 0 && p.terms) {  // the loss terms are local (skrl logs them per rank)
        const float bce_cat = p.acc[ACC_BCE_CAT] / (2.0f * (float)p.B), bce_motion = p.acc[ACC_BCE_MOTION] / (float)p.B;
        const float reg = p.acc[ACC_W3_SQ], gp = p.acc[ACC_GP_SQ] / (float)p.B, wd = p.acc[ACC_W_SQ] + p.acc[ACC_W3_SQ];
        p.terms[0] = bce_cat; p.terms[1] = bce_motion; p.terms[2] = reg; p.terms[3] = gp; p.terms[4] = wd;
        p.terms[5] = p.loss_scale * (0.5f * (bce_cat + bce_motion) + p.c_reg * reg + p.c_gp * gp + p.c_wd * wd);
    }
    // ---- barrier A: the last CTA to finish its pushes announces; every CTA waits for all ranks ----
    // ONE system-scope fence per CTA, by the thread that then signals: the CTA barrier orders every thread's stores before it and
    // the fence is cumulative (the cooperative-groups grid barrier is built the same way).  A fence per thread -- 32 warps per SM
    // each draining the SM's peer stores -- cost more than the exchange itself.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(&x.ctl->pushed, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
#if AMP_EXCHANGE_STAMPS == 1
        if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); x.timing[2] = t; }
#endif
        if (threadIdx.x == 0) x.ctl->pushed = 0;  // every CTA has passed the counter
        if (threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(x.peers.flags[threadIdx.x] + rank, epoch);
        }
    }
    if (threadIdx.x == 0) {
        ok = wait_all(local_flags, world, epoch, x.spin_limit);
        if (!ok && blockIdx.x == 0) report_failure(x.ctl, x.host_status, 1u, bucket_local, x.hull, x.count);
    }
    __syncthreads();
    if (ok && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t_a;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_a));
#if AMP_EXCHANGE_STAMPS == 1
        x.timing[3] = t_a;
#elif AMP_EXCHANGE_STAMPS == 2
        x.timing[0] = t_a;
#else
        x.timing[0] = t_start;
        x.timing[1] = t_a;
#endif
    }
    // ---- reduce the own slice (local loads) and publish the mean to every rank's bucket ----
    const long long q0 = (long long)rank * per, q1 = min(quads, q0 + per);
    const float inv = 1.0f / (float)world;
    const float4 *staged = reinterpret_cast<const float4 *>(bucket_local + x.stage);
    for (long long q = q0 + tid; SWITCH && ok && q < q1; q += stride) {
        float *at = x.mc + x.hull + 4 * q;
        if (x.count - 4 * q >= 4) {
            float4 v;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(at) : "memory");
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(at), "f"(v.x * inv), "f"(v.y * inv), "f"(v.z * inv), "f"(v.w * inv) : "memory");
        } else {  // dL/db3: one float, the rest of the quad belongs to somebody else
            float v;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(at) : "memory");
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(at), "f"(v * inv) : "memory");
        }
    }
    for (long long q = q0 + tid; !SWITCH && ok && q < q1; q += stride) {
        float4 v[MAXW];
#pragma unroll
        for (int r = 0; r < MAXW; ++r)
            if (r < world) v[r] = ld_peer(staged + (long long)r * per + (q - q0));  // .cg: written by the peers, never in this L1
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < MAXW; ++r)  // fixed rank order, starting from zero: the all-reduce kernels' sum
            if (r < world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        const int valid = (int)min(4LL, x.count - 4 * q);  // < 4 only in the ragged last quad: its other floats are not ours
#pragma unroll
        for (int r = 0; r < MAXW; ++r) {
            if (r < world) {
                float *dst = x.peers.data[r] + x.hull + 4 * q;
                if (valid == 4) {
                    st_peer(reinterpret_cast<float4 *>(dst), acc);
                } else {
                    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
                    for (int j = 0; j < valid; ++j) asm volatile("st.global.cg.f32 [%0], %1;" ::"l"(dst + j), "f"(a[j]) : "memory");
                }
            }
        }
    }
#if AMP_EXCHANGE_STAMPS == 2
    if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); x.timing[1] = t; }
#endif
    // ---- barrier B (amp_bucket.cu) ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(&x.ctl->arrivals, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        x.ctl->arrivals = 0;
        x.ctl->epoch = epoch + 1;
    }
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(x.peers.flags[threadIdx.x] + rank, epoch + 1);
    }
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#if AMP_EXCHANGE_STAMPS != 1
        x.timing[2] = t;
#endif
        if (ok && !wait_all(local_flags, world, epoch + 1, x.spin_limit)) report_failure(x.ctl, x.host_status, 2u, bucket_local, x.hull, x.count);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#if AMP_EXCHANGE_STAMPS != 1
        x.timing[3] = t;
#endif
    }
}

__global__ void sumsq_kernel(const float *__restrict__ x, int n, float *__restrict__ out) {
    float s = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(x[i], x[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const GemmParams &p, cudaStream_t st) {
    auto kern = train_gemm_kernel<A_MN, B_MN, EPI>;  // dynamic shared memory opt-in: configure_gemm_kernels(), at create
    const int items = p.m_tiles * p.n_tiles * p.splits;
    if (items <= 0) return AMP_OK;
    const int grid = std::min(items, sm_count());
    kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(ta, tb, p);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

// every instantiation the step launches, for the CURRENT device (the attribute is per device; amp_disc_train_create calls
// this, so a captured CUDA graph never contains a first-use configuration call)
static cudaError_t configure_gemm_kernels() {
    cudaError_t e = cudaFuncSetAttribute(train_gemm_kernel<false, false, EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(train_gemm_kernel<false, false, EPI_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(train_gemm_kernel<false, false, EPI_SCALE_SUMSQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(train_gemm_kernel<true, true, EPI_SLICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    return e;
}

}  // namespace train
}  // namespace amp

struct amp_disc_train {
    int in_features, Kp, h1, h2, device;
    int64_t max_batch, Bp_max;
    __nv_bfloat16 *X, *A1, *A2, *E1, *E2, *S;   // activations (row counts: 4Bp, 4Bp, 3Bp, 4Bp, 4Bp, Bp)
    __nv_bfloat16 *W1b, *W1t, *W2b, *W2t;       // bf16 weights: (h1,Kp), (Kp,h1), (h2,h1), (h1,h2)
    float *ws;                                  // fp32: [ACC_COUNT scalars | gw3 h2 | gb2 h2 | gb1 h1]
    float *slices1, *slices2;                   // split-K slices of dW1 (splits1, h1, Kp) and dW2 (splits2, h2, h1)
    int splits1, splits2;
    int64_t staged_rows[3];                     // rows staged per source since the last step (-1: none)
    int maps_Bp;                                // padded batch size the tensor maps below were encoded for (0: none yet)
    CUtensorMap tX_k, tX_mn, tA1_k, tA1_mn, tE2_k, tE2_mn, tE1_mn, tW1b, tW1t, tW2b, tW2t, tXgp_k, tA1gp_k, tE1gp_k;
};

using namespace amp;
using namespace amp::train;

extern "C" {

int amp_disc_train_destroy(amp_disc_train_t *t) {
    if (!t) return AMP_OK;
    void *ptrs[] = {t->X, t->A1, t->A2, t->E1, t->E2, t->S, t->W1b, t->W1t, t->W2b, t->W2t, t->ws, t->slices1, t->slices2};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete t;
    return AMP_OK;
}

int amp_disc_train_create(int32_t in_features, int32_t h1, int32_t h2, int64_t max_batch_rows, void *stream, amp_disc_train_t **out) {
    AMP_REQUIRE(out, "amp_disc_train_create: NULL out");
    *out = nullptr;
    AMP_REQUIRE(in_features >= 1 && in_features <= 1024 && h1 >= BN && h2 >= BN && h1 % BN == 0 && h2 % BN == 0 && h2 <= 2048,
                "amp_disc_train_create: hidden sizes must be multiples of %d, h2 <= 2048 (got %d, %d), 1 <= in_features <= 1024 (got %d)", BN,
                h1, h2, in_features);
    AMP_REQUIRE(max_batch_rows >= 1 && max_batch_rows <= (1 << 22), "amp_disc_train_create: max_batch_rows must be in [1, 4194304]");
    int dev = 0, major = 0;
    AMP_CUDA_TRY(cudaGetDevice(&dev));
    AMP_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(AMP_ENODEV, "amp_disc_train_create: tcgen05 kernels need an sm_100 device (found compute capability %d.x)", major);
    amp_disc_train *t = new (std::nothrow) amp_disc_train();
    if (!t) return fail(AMP_ENOMEM, "amp_disc_train_create: host allocation failed");
    std::memset(t, 0, sizeof(*t));
    t->device = dev;
    t->in_features = in_features;
    t->Kp = (in_features + BK - 1) / BK * BK;
    t->h1 = h1;
    t->h2 = h2;
    t->max_batch = max_batch_rows;
    t->Bp_max = (max_batch_rows + BM - 1) / BM * BM;
    for (int i = 0; i < 3; ++i) t->staged_rows[i] = -1;
    const int sms = sm_count();
    // split-K so that the weight-gradient products fill the machine: items = splits x (M/128) x ceil(N/256) ~ one wave
    const int tiles1 = (h1 / BM) * ((t->Kp + BN - 1) / BN), tiles2 = (h2 / BM) * (h1 / BN);
    t->splits1 = std::max(1, std::min(16, sms / tiles1));
    t->splits2 = std::max(1, std::min(16, sms / tiles2));
    const size_t Bp = (size_t)t->Bp_max;
    struct { void **p; size_t bytes; } allocs[] = {
        {(void **)&t->X, 4 * Bp * t->Kp * 2},   {(void **)&t->A1, 4 * Bp * h1 * 2},     {(void **)&t->A2, 3 * Bp * h2 * 2},
        {(void **)&t->E1, 4 * Bp * h1 * 2},     {(void **)&t->E2, 4 * Bp * h2 * 2},     {(void **)&t->S, Bp * h2 * 2},
        {(void **)&t->W1b, (size_t)h1 * t->Kp * 2}, {(void **)&t->W1t, (size_t)t->Kp * h1 * 2},
        {(void **)&t->W2b, (size_t)h2 * h1 * 2},    {(void **)&t->W2t, (size_t)h1 * h2 * 2},
        {(void **)&t->ws, ((size_t)ACC_COUNT + (size_t)REPL * (2 * h2 + h1) + 3 * Bp) * 4},
        {(void **)&t->slices1, (size_t)t->splits1 * h1 * t->Kp * 4}, {(void **)&t->slices2, (size_t)t->splits2 * h2 * h1 * 4},
    };
    cudaStream_t st = as_stream(stream);
    for (auto &a : allocs) {
        cudaError_t e = cudaMalloc(a.p, a.bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(*a.p, 0, a.bytes, st);  // padding rows must hold finite values from the start
        if (e != cudaSuccess) {
            amp_disc_train_destroy(t);
            return cuda_fail(e, "cudaMalloc(amp_disc_train_create)");
        }
    }
    {
        cudaError_t e = configure_gemm_kernels();
        if (e != cudaSuccess) {
            amp_disc_train_destroy(t);
            return cuda_fail(e, "cudaFuncSetAttribute(train_gemm_kernel)");
        }
    }
    if ((size_t)32 * h2 * 2 > 48 * 1024) {  // head_kernel keeps a 32-row tile of a2 in shared memory
        // per function, not per handle: always the limit of the widest head create accepts (h2 = 2048), so that a narrower
        // handle created later cannot lower it under a wider one
        cudaError_t e = cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 2048 * 2);
        if (e != cudaSuccess) {
            amp_disc_train_destroy(t);
            return cuda_fail(e, "cudaFuncSetAttribute(head_kernel)");
        }
    }
    *out = t;
    return AMP_OK;
}

int amp_disc_train_stage(amp_disc_train_t *t, int32_t source, const float *x, int64_t x_stride, int64_t batch_rows,
                         const double *running_mean, const double *running_variance, void *stream) {
    const int64_t rows = batch_rows;
    AMP_REQUIRE(t && source >= 0 && source < 3, "amp_disc_train_stage: bad handle or source %d (0 agent, 1 replay, 2 motion)", source);
    AMP_REQUIRE(batch_rows >= 1 && batch_rows <= t->max_batch, "amp_disc_train_stage: batch_rows %lld outside [1, %lld]",
                (long long)batch_rows, (long long)t->max_batch);
    AMP_REQUIRE(x && x_stride >= t->in_features, "amp_disc_train_stage: NULL x or x_stride < in_features");
    AMP_REQUIRE((running_mean == nullptr) == (running_variance == nullptr), "amp_disc_train_stage: give both scaler buffers or neither");
    cudaStream_t st = as_stream(stream);
    const int64_t Bp = (batch_rows + BM - 1) / BM * BM;
    stage_cast_kernel<<<dim3((t->Kp + 511) / 512, (unsigned)((rows + STAGE_ROWS - 1) / STAGE_ROWS)), 256, 0, st>>>(
        x, x_stride, (int)rows, t->in_features, t->Kp, running_mean, running_variance, t->X + (size_t)source * Bp * t->Kp);
    AMP_CUDA_TRY(cudaGetLastError());
    t->staged_rows[source] = rows;
    return AMP_OK;
}

}  // extern "C"

// the exchanged range of a fused step: the six gradient tensors must tile one contiguous, quad-aligned range of the bucket
static int plan_exchange(amp_disc_train_t *t, amp_bucket_t *bk, float *const g[6], ExchangeParams *x) {
    const long long n[6] = {(long long)t->h1 * t->in_features, t->h1, (long long)t->h2 * t->h1, t->h2, t->h2, 1};
    long long off[6];
    int order[6] = {0, 1, 2, 3, 4, 5};
    for (int k = 0; k < 6; ++k) {
        off[k] = g[k] - bk->data;
        AMP_REQUIRE(g[k] >= bk->data && off[k] + n[k] <= bk->floats, "amp_disc_train_step_exchange: gradient tensor %d is not inside the bucket", k);
    }
    std::sort(order, order + 6, [&](int a, int c) { return off[a] < off[c]; });
    AMP_REQUIRE(off[order[0]] % 4 == 0, "amp_disc_train_step_exchange: the gradient range must start on a 16-byte boundary of the bucket");
    for (int k = 0; k + 1 < 6; ++k)
        AMP_REQUIRE(off[order[k]] + n[order[k]] == off[order[k + 1]],
                    "amp_disc_train_step_exchange: the six gradient tensors must lie side by side in the bucket (dL/db3, the one ragged "
                    "tensor, last)");
    AMP_REQUIRE(order[5] == 5, "amp_disc_train_step_exchange: dL/db3 must be the last tensor of the range");
    x->hull = off[order[0]];
    x->count = off[5] + 1 - x->hull;
    AMP_REQUIRE(x->count < (1LL << 31), "amp_disc_train_step_exchange: gradient range of %lld floats is too long (32-bit index arithmetic)", x->count);
    for (int k = 0; k < 6; ++k) x->seg[k] = off[k] - x->hull;
    const long long quads = (x->count + 3) / 4, per = (quads + bk->world - 1) / bk->world;
    AMP_REQUIRE(bk->world == 1 || per * 4 * bk->world <= bk->stage_floats, "amp_disc_train_step_exchange: staging area too small");  // sized from the bucket
    x->peers = bk->peers;
    x->mc = bk->connected && bk->in_switch ? bk->mc_data : nullptr;
    x->rank = bk->rank;
    x->world = bk->world;
    x->stage = bk->floats;
    x->ctl = reinterpret_cast<bucket::Control *>(bk->flags + bucket::kMaxWorld);
    x->host_status = bk->host_status_dev;
    x->spin_limit = bk->spin_limit;
    x->timing = bk->timing;
    return AMP_OK;
}

static int train_step_impl(amp_disc_train_t *t, const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                           const float *b3, int64_t batch_rows, float loss_scale, float logit_regularization_scale,
                           float gradient_penalty_scale, float weight_decay_scale, float *gW1, float *gb1, float *gW2, float *gb2,
                           float *gW3, float *gb3, float *terms, float *logits, amp_bucket_t *bucket, void *stream) {
    AMP_REQUIRE(t && W1 && b1 && W2 && b2 && W3 && b3 && gW1 && gb1 && gW2 && gb2 && gW3 && gb3, "amp_disc_train_step: NULL argument");
    ExchangeParams ex{};
    const bool exchange = bucket && bucket->world > 1;
    if (bucket) {  // the placement rules hold for a world of one too (same call, same errors on every world size)
        AMP_REQUIRE(bucket->connected, "amp_disc_train_step_exchange: amp_bucket_connect has not been called");
        AMP_REQUIRE(bucket->world <= 8, "amp_disc_train_step_exchange: at most 8 ranks (got %d)", bucket->world);
        if (const uint32_t failed = *bucket->host_status)
            return fail(AMP_ECUDA, "amp_disc_train_step_exchange: an earlier exchange on this bucket timed out waiting for a peer (status bits 0x%x)",
                        failed & 0x7fffffffu);
        float *const g6[6] = {gW1, gb1, gW2, gb2, gW3, gb3};
        if (int rc = plan_exchange(t, bucket, g6, &ex)) return rc;
    }
    AMP_REQUIRE(batch_rows >= 1 && batch_rows <= t->max_batch, "amp_disc_train_step: batch_rows %lld outside [1, %lld]",
                (long long)batch_rows, (long long)t->max_batch);
    for (int i = 0; i < 3; ++i)
        AMP_REQUIRE(t->staged_rows[i] == batch_rows, "amp_disc_train_step: source %d has %lld staged rows, the step needs %lld", i,
                    (long long)t->staged_rows[i], (long long)batch_rows);
    cudaStream_t st = as_stream(stream);
    const int B = (int)batch_rows;
    const int Bp = (B + BM - 1) / BM * BM;
    const int Kp = t->Kp, h1 = t->h1, h2 = t->h2;
    float *acc = t->ws, *ws_gw3 = t->ws + ACC_COUNT, *ws_gb2 = ws_gw3 + REPL * h2, *ws_gb1 = ws_gb2 + REPL * h2;
    float *ws_dot = ws_gb1 + REPL * h1;  // [3Bp] a2 . w3, accumulated by the layer-2 epilogue
    AMP_CUDA_TRY(cudaMemsetAsync(t->ws, 0, ((size_t)ACC_COUNT + (size_t)REPL * (2 * h2 + h1) + 3 * (size_t)Bp) * 4, st));

    // ---- weights: bf16 copies (both orientations) + sums of squares ----
    cast_transpose_kernel<<<dim3(Kp / 32, h1 / 32), 256, 0, st>>>(W1, h1, t->in_features, Kp, t->W1b, t->W1t, acc + ACC_W_SQ);
    cast_transpose_kernel<<<dim3(h1 / 32, h2 / 32), 256, 0, st>>>(W2, h2, h1, h1, t->W2b, t->W2t, acc + ACC_W_SQ);
    sumsq_kernel<<<1, 256, 0, st>>>(W3, h2, acc + ACC_W3_SQ);
    AMP_CUDA_TRY(cudaGetLastError());

    // ---- tensor maps: the buffers never move, only the row counts depend on the batch -> encoded once per padded batch size
    // (fifteen cuTensorMapEncodeTiled calls were ~15 us of host time on every step) ----
    if (t->maps_Bp != Bp) {
        int rc = AMP_OK;
        auto mk = [&](CUtensorMap *m, const void *ptr, int64_t rows, int64_t cols, int box_rows, int box_cols) {
            if (rc == AMP_OK) rc = make_tmap(m, ptr, rows, cols, cols, box_rows, box_cols);
        };
        mk(&t->tX_k, t->X, 4LL * Bp, Kp, BM, BK);
        mk(&t->tX_mn, t->X, 4LL * Bp, Kp, BK, 64);
        mk(&t->tA1_k, t->A1, 4LL * Bp, h1, BM, BK);
        mk(&t->tA1_mn, t->A1, 4LL * Bp, h1, BK, 64);
        mk(&t->tE2_k, t->E2, 4LL * Bp, h2, BM, BK);
        mk(&t->tE2_mn, t->E2, 4LL * Bp, h2, BK, 64);
        mk(&t->tE1_mn, t->E1, 4LL * Bp, h1, BK, 64);
        mk(&t->tW1b, t->W1b, h1, Kp, BN, BK);
        mk(&t->tW1t, t->W1t, Kp, h1, BN, BK);
        mk(&t->tW2b, t->W2b, h2, h1, BN, BK);
        mk(&t->tW2t, t->W2t, h1, h2, BN, BK);
        // row-block views: a TMA coordinate is relative to the map's base, so sub-ranges get their own maps
        mk(&t->tXgp_k, t->X + (size_t)3 * Bp * Kp, Bp, Kp, BM, BK);
        mk(&t->tA1gp_k, t->A1 + (size_t)3 * Bp * h1, Bp, h1, BM, BK);
        mk(&t->tE1gp_k, t->E1 + (size_t)3 * Bp * h1, Bp, h1, BM, BK);
        if (rc != AMP_OK) return rc;
        t->maps_Bp = Bp;
    }
    const CUtensorMap &tX_k = t->tX_k, &tX_mn = t->tX_mn, &tA1_k = t->tA1_k, &tA1_mn = t->tA1_mn, &tE2_k = t->tE2_k, &tE2_mn = t->tE2_mn,
                      &tE1_mn = t->tE1_mn, &tW1b = t->tW1b, &tW1t = t->tW1t, &tW2b = t->tW2b, &tW2t = t->tW2t,
                      &tXgp_k = t->tXgp_k, &tA1gp_k = t->tA1gp_k, &tE1gp_k = t->tE1gp_k;
    int rc = AMP_OK;

    GemmParams g{};
    // F1: A1[:3Bp] = relu(X[:3Bp] W1^T + b1)
    g = GemmParams{};
    g.m_tiles = 3 * Bp / BM; g.n_tiles = h1 / BN; g.splits = 1; g.kblocks = Kp / BK; g.n_valid = h1;
    g.bias = b1; g.out = t->A1; g.out_ld = h1;
    if ((rc = launch_gemm<false, false, EPI_BIAS_RELU>(tX_k, tW1b, g, st)) != AMP_OK) return rc;
    // F2: A2 = relu(A1[:3Bp] W2^T + b2)
    g = GemmParams{};
    g.m_tiles = 3 * Bp / BM; g.n_tiles = h2 / BN; g.splits = 1; g.kblocks = h1 / BK; g.n_valid = h2;
    g.bias = b2; g.out = t->A2; g.out_ld = h2; g.dot_w = W3; g.dot_out = ws_dot;
    if ((rc = launch_gemm<false, false, EPI_BIAS_RELU>(tA1_k, tW2b, g, st)) != AMP_OK) return rc;
    // head: logits, BCE, dd, E2 (dz2 | u2), dL/db2, dL/dw3 (first part), dL/db3
    head_kernel<<<3 * Bp / 32, 256, (size_t)32 * h2 * 2, st>>>(t->A2, ws_dot, h2, Bp, B, W3, b3, loss_scale, t->E2, acc, ws_gw3, ws_gb2, logits);
    AMP_CUDA_TRY(cudaGetLastError());
    // B1: E1 = (E2 W2) * m1 -- rows < 3Bp use a1 of the same row, rows >= 3Bp (u2 -> v1) the a1 of the motion block
    g = GemmParams{};
    g.m_tiles = 4 * Bp / BM; g.n_tiles = h1 / BN; g.splits = 1; g.kblocks = h2 / BK; g.n_valid = h1;
    g.out = t->E1; g.out_ld = h1; g.mask = t->A1; g.mask_ld = h1; g.mask_split = 3 * Bp; g.mask_shift = Bp; g.alpha = 1.0f;
    if ((rc = launch_gemm<false, false, EPI_MASK>(tE2_k, tW2t, g, st)) != AMP_OK) return rc;
    // GP2: g = v1 W1 (motion rows); X[3Bp:] = G = (2 s c_gp / B) g; sum g^2
    g = GemmParams{};
    g.m_tiles = Bp / BM; g.n_tiles = (Kp + BN - 1) / BN; g.splits = 1; g.kblocks = h1 / BK; g.n_valid = Kp;
    g.out = t->X + (size_t)3 * Bp * Kp; g.out_ld = Kp; g.alpha = 2.0f * loss_scale * gradient_penalty_scale / (float)B;
    g.sumsq = acc + ACC_GP_SQ;
    if ((rc = launch_gemm<false, false, EPI_SCALE_SUMSQ>(tE1gp_k, tW1t, g, st)) != AMP_OK) return rc;
    // GP3: A1[3Bp:] = q1 = (G W1^T) * m1(motion)
    g = GemmParams{};
    g.m_tiles = Bp / BM; g.n_tiles = h1 / BN; g.splits = 1; g.kblocks = Kp / BK; g.n_valid = h1;
    g.out = t->A1 + (size_t)3 * Bp * h1; g.out_ld = h1; g.mask = t->A1 + (size_t)2 * Bp * h1; g.mask_ld = h1;
    g.mask_split = 1 << 30; g.mask_shift = 0; g.alpha = 1.0f;
    if ((rc = launch_gemm<false, false, EPI_MASK>(tXgp_k, tW1b, g, st)) != AMP_OK) return rc;
    // GP6: S = (q1 W2^T) * m2(motion)
    g = GemmParams{};
    g.m_tiles = Bp / BM; g.n_tiles = h2 / BN; g.splits = 1; g.kblocks = h1 / BK; g.n_valid = h2;
    g.out = t->S; g.out_ld = h2; g.mask = t->A2 + (size_t)2 * Bp * h2; g.mask_ld = h2; g.mask_split = 1 << 30; g.mask_shift = 0;
    g.alpha = 1.0f;
    if ((rc = launch_gemm<false, false, EPI_MASK>(tA1gp_k, tW2b, g, st)) != AMP_OK) return rc;
    // dW2 slices = E2^T A1 over all 4Bp rows (K = rows: both operands MN-major)
    g = GemmParams{};
    g.m_tiles = h2 / BM; g.n_tiles = h1 / BN; g.splits = std::min(t->splits2, 4 * Bp / BK); g.kblocks = 4 * Bp / BK; g.n_valid = h1;
    g.slices = t->slices2; g.slice_ld = h1; g.slice_stride = (long long)h2 * h1;
    const int used2 = g.splits;
    if ((rc = launch_gemm<true, true, EPI_SLICE>(tE2_mn, tA1_mn, g, st)) != AMP_OK) return rc;
    // dW1 slices = E1^T X
    g = GemmParams{};
    g.m_tiles = h1 / BM; g.n_tiles = (Kp + BN - 1) / BN; g.splits = std::min(t->splits1, 4 * Bp / BK); g.kblocks = 4 * Bp / BK; g.n_valid = Kp;
    g.slices = t->slices1; g.slice_ld = Kp; g.slice_stride = (long long)h1 * Kp;
    const int used1 = g.splits;
    if ((rc = launch_gemm<true, true, EPI_SLICE>(tE1_mn, tX_mn, g, st)) != AMP_OK) return rc;
    // bias gradient of layer 1 and the gradient-penalty part of dL/dw3
    if (Kp == t->in_features)  // no spare column for the "ones" trick: explicit column sum
        colsum_kernel<<<dim3((h1 + 511) / 512, (3 * Bp + COLSUM_ROWS - 1) / COLSUM_ROWS), 256, 0, st>>>(t->E1, h1, 3 * Bp, h1, ws_gb1);
    colsum_kernel<<<dim3((h2 + 511) / 512, (Bp + COLSUM_ROWS - 1) / COLSUM_ROWS), 256, 0, st>>>(t->S, h2, Bp, h2, ws_gw3);
    AMP_CUDA_TRY(cudaGetLastError());

    FinalizeParams f{};
    f.W1 = W1; f.W2 = W2; f.w3 = W3; f.slices1 = t->slices1; f.slices2 = t->slices2; f.splits1 = used1; f.splits2 = used2;
    f.in_features = t->in_features; f.Kp = Kp; f.h1 = h1; f.h2 = h2; f.B = B; f.acc = acc; f.ws_gw3 = ws_gw3; f.ws_gb2 = ws_gb2;
    f.ws_gb1 = ws_gb1; f.loss_scale = loss_scale; f.c_reg = logit_regularization_scale; f.c_gp = gradient_penalty_scale;
    f.c_wd = weight_decay_scale; f.gW1 = gW1; f.gb1 = gb1; f.gW2 = gW2; f.gb2 = gb2; f.gW3 = gW3; f.gb3 = gb3; f.terms = terms;
    if (exchange) {
        // every CTA waits for flags the peers send from their LAST CTA: the whole grid has to be resident
        auto launch = [&](auto kern) -> int {
            int per_sm = 0;
            AMP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
            AMP_REQUIRE(per_sm >= 1, "amp_disc_train_step_exchange: the exchange kernel does not fit on an SM");
            const long long quads = (ex.count + 3) / 4;
            const int grid = (int)std::max<long long>(1, std::min<long long>((quads + 255) / 256, (long long)std::min(per_sm, 4) * sm_count()));
            kern<<<grid, 256, 0, st>>>(f, ex);
            return AMP_OK;
        };
        int rc = ex.mc ? launch(finalize_exchange_kernel<2, true>)
                 : bucket->world <= 2 ? launch(finalize_exchange_kernel<2, false>)
                 : bucket->world <= 4 ? launch(finalize_exchange_kernel<4, false>)
                                      : launch(finalize_exchange_kernel<8, false>);
        if (rc != AMP_OK) return rc;
    } else {
        finalize_kernel<<<sm_count() * 4, 256, 0, st>>>(f);
    }
    AMP_CUDA_TRY(cudaGetLastError());
    for (int i = 0; i < 3; ++i) t->staged_rows[i] = -1;
    return AMP_OK;
}

extern "C" {

int amp_disc_train_step(amp_disc_train_t *t, const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                        const float *b3, int64_t batch_rows, float loss_scale, float logit_regularization_scale,
                        float gradient_penalty_scale, float weight_decay_scale, float *gW1, float *gb1, float *gW2, float *gb2,
                        float *gW3, float *gb3, float *terms, float *logits, void *stream) {
    return train_step_impl(t, W1, b1, W2, b2, W3, b3, batch_rows, loss_scale, logit_regularization_scale, gradient_penalty_scale,
                           weight_decay_scale, gW1, gb1, gW2, gb2, gW3, gb3, terms, logits, nullptr, stream);
}

int amp_disc_train_step_exchange(amp_disc_train_t *t, const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                                 const float *b3, int64_t batch_rows, float loss_scale, float logit_regularization_scale,
                                 float gradient_penalty_scale, float weight_decay_scale, float *gW1, float *gb1, float *gW2, float *gb2,
                                 float *gW3, float *gb3, float *terms, float *logits, amp_bucket_t *bucket, void *stream) {
    AMP_REQUIRE(bucket, "amp_disc_train_step_exchange: NULL bucket");
    return train_step_impl(t, W1, b1, W2, b2, W3, b3, batch_rows, loss_scale, logit_regularization_scale, gradient_penalty_scale,
                           weight_decay_scale, gW1, gb1, gW2, gb2, gW3, gb3, terms, logits, bucket, stream);
}

}  // extern "C"
