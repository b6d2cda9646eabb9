// K1 / K2 of the AMP hot path: motion sampling (frame index + blend, lerp, slerp) and AMP-observation building.
//
//   amp_frame_blend        <- MotionLoader._compute_frame_blend      (reference motions/motion_loader.py:281-307)
//   amp_sample_full        <- MotionLoader.sample                    (:368-390)
//   amp_lerp / amp_slerp   <- MotionLoader._interpolate / _slerp     (:211-215, :242-279)
//   amp_collect_reference  <- G1AmpEnv.collect_reference_motions     (reference g1_amp_env.py:445-486)
//   amp_compute_obs        <- compute_obs                            (:535-561)
//   amp_tangent_normal     <- quaternion_to_tangent_and_normal       (:489-497)
//   amp_obs_step           <- G1AmpEnv._get_observations, AMP part   (:176-193)
//
// All of it is gather / transform / stream-out work bounded by HBM bandwidth, not by math: no tensor cores here.
// Design for B200: one warp owns a run of consecutive destination frames, consecutive lanes own consecutive
// observation columns (so every global load and store of a warp is one contiguous 128-byte run), the float64 index
// math and the root slerp are done once per frame by one lane and staged through shared memory, and the grid is a
// persistent multiple of the SM count.
//
// Compile with -fmad=false (see amp_math.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "amp_internal.h"
#include "amp_math.cuh"

namespace amp {

struct KeyList {
    int32_t k[kMaxKeyBodies];
};

// ------------------------------------------------------------------------------------------------------------------
// Packed AMP row table, built once per library.  Row f (R = A rounded up to 4 floats, 16-byte aligned) holds exactly
// the clip columns compute_obs consumes, already permuted into robot dof order and laid out at the observation's own
// column offsets so that output column c is interpolated from packed column c:
//   [0,D)        dof_pos[dof_indexes]          [D,2D)      dof_vel[dof_indexes]
//   2D           root z                        2D+1..2D+4  root quaternion wxyz   (tangent/normal columns of the obs)
//   2D+5, 2D+6   root x, y                     2D+7..2D+9  root linear velocity   2D+10..2D+12 root angular velocity
//   2D+13+3j+a   key body j position, axis a (absolute; the root position is subtracted after interpolation)
// This folds the column gathers of g1_amp_env.py:478-484 into the staging and cuts the per-frame read from
// (2*D_clip + 13*B) floats to R floats (G1_dance: 2260 B -> 336 B).
// ------------------------------------------------------------------------------------------------------------------
__global__ void pack_rows_kernel(LibView v, const int32_t *__restrict__ dof_idx, int32_t ref, KeyList keys,
                                 float *__restrict__ packed) {
    const int64_t total = v.num_frames * (int64_t)v.row_floats;
    const int D = v.obs_dofs, D2 = 2 * D, B = v.num_bodies;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = e / v.row_floats;
        const int c = (int)(e - f * v.row_floats);
        float val = 0.0f;
        if (c < D) {
            val = v.dof_pos[f * v.num_dofs + dof_idx[c]];
        } else if (c < D2) {
            val = v.dof_vel[f * v.num_dofs + dof_idx[c - D]];
        } else if (c == D2) {
            val = v.body_pos[(f * B + ref) * 3 + 2];
        } else if (c < D2 + 5) {
            val = v.body_rot[(f * B + ref) * 4 + (c - D2 - 1)];
        } else if (c < D2 + 7) {
            val = v.body_pos[(f * B + ref) * 3 + (c - D2 - 5)];
        } else if (c < D2 + 10) {
            val = v.body_lin[(f * B + ref) * 3 + (c - D2 - 7)];
        } else if (c < D2 + 13) {
            val = v.body_ang[(f * B + ref) * 3 + (c - D2 - 10)];
        } else if (c < v.obs_width) {
            const int j = (c - D2 - 13) / 3, a = (c - D2 - 13) % 3;
            val = v.body_pos[(f * B + keys.k[j]) * 3 + a];
        }
        packed[e] = val;
    }
}

// Looks up the trajectory of a sample and evaluates the float64 frame/blend math; clamps for memory safety and raises
// the sticky flags where the reference would raise IndexError / propagate NaN.
__device__ __forceinline__ FrameBlend lookup_frame(const LibView &v, double t, int64_t id) {
    if (id < 0 || id >= v.num_traj) {
        atomicOr(v.flags, 1u);
        id = id < 0 ? 0 : v.num_traj - 1;
    }
    const int64_t start = v.starts[id], end = v.ends[id];
    FrameBlend fb = frame_blend(t, v.durations[id], start, end, v.dt);
    if (fb.i0 < start || fb.i0 > end) {  // only reachable with a NaN time (or a zero-length clip): rint(NaN) -> INT64_MIN
        atomicOr(v.flags, 2u);
        fb.i0 = fb.i0 < start ? start : end;
        fb.i1 = fb.i0;
    }
    return fb;
}

// ------------------------------------------------------------------------------------------------------------------
// amp_frame_blend
// ------------------------------------------------------------------------------------------------------------------
__global__ void frame_blend_kernel(LibView v, const double *__restrict__ times, const int64_t *__restrict__ ids,
                                   int64_t S, int64_t *__restrict__ idx0, int64_t *__restrict__ idx1,
                                   float *__restrict__ blend32, double *__restrict__ blend64) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t id = ids ? ids[i] : 0;
        if (id < 0 || id >= v.num_traj) {
            atomicOr(v.flags, 1u);
            id = id < 0 ? 0 : v.num_traj - 1;
        }
        // the raw (unclamped) result is reported, exactly what the reference method returns
        const FrameBlend fb = frame_blend(times[i], v.durations[id], v.starts[id], v.ends[id], v.dt);
        idx0[i] = fb.i0;
        idx1[i] = fb.i1;
        if (blend32) blend32[i] = __double2float_rn(fb.blend);
        if (blend64) blend64[i] = fb.blend;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// amp_sample_full: a warp owns 32 consecutive output frames.
//   phase 1  one lane per frame does the float64 index math (instead of every lane repeating it for one frame);
//   phase 2  for each of the six tensors the warp walks the tile's elements FLAT (element e -> frame e / width, column
//            e % width, by multiply-high with a precomputed reciprocal), so all 32 lanes are busy whatever the row width
//            (a 33-float body_positions row would otherwise leave half of the second pass idle, an 11-quaternion
//            body_rotations row two thirds of the slerp lanes) and the stores of a tile are one contiguous run.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSampleWarps = 8;

struct SampleDivisors {
    uint32_t dof, body3, body;  // ceil(2^32 / width) for width = D, 3B, B
};

__device__ __forceinline__ void lerp_tile(const float *__restrict__ src, float *__restrict__ dst, int width, uint32_t magic,
                                          int nf, const int *s_i0, const int *s_i1, const float *s_b, int lane) {
    const int total = nf * width;
#pragma unroll 4
    for (int e = lane; e < total; e += 32) {
        const int fl = (int)__umulhi((uint32_t)e, magic);  // e / width, exact for e < 2^32 / width
        const int c = e - fl * width;
        const float b = s_b[fl];
        const float a0 = __ldg(src + (int64_t)s_i0[fl] * width + c), a1 = __ldg(src + (int64_t)s_i1[fl] * width + c);
        __stcs(dst + e, lerp(b, a0, a1));
    }
}

__global__ void __launch_bounds__(kSampleWarps * 32) sample_full_kernel(LibView v, SampleDivisors dv,
                                                                         const double *__restrict__ times,
                                                                         const int64_t *__restrict__ ids, int64_t S,
                                                                         float *__restrict__ dof_pos, float *__restrict__ dof_vel,
                                                                         float *__restrict__ body_pos, float *__restrict__ body_rot,
                                                                         float *__restrict__ body_lin, float *__restrict__ body_ang) {
    __shared__ int s_i0_all[kSampleWarps][32], s_i1_all[kSampleWarps][32];
    __shared__ float s_b_all[kSampleWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *s_i0 = s_i0_all[warp], *s_i1 = s_i1_all[warp];
    float *s_b = s_b_all[warp];
    const int D = v.num_dofs, B = v.num_bodies;
    const int64_t num_tiles = (S + 31) / 32;
    for (int64_t tile = blockIdx.x * (int64_t)kSampleWarps + warp; tile < num_tiles; tile += (int64_t)gridDim.x * kSampleWarps) {
        const int64_t f0 = tile * 32;
        const int nf = (int)min((int64_t)32, S - f0);
        if (lane < nf) {
            const FrameBlend fb = lookup_frame(v, times[f0 + lane], ids ? ids[f0 + lane] : 0);
            s_i0[lane] = (int)fb.i0;
            s_i1[lane] = (int)fb.i1;
            s_b[lane] = __double2float_rn(fb.blend);
        }
        __syncwarp();
        if (dof_pos) lerp_tile(v.dof_pos, dof_pos + f0 * D, D, dv.dof, nf, s_i0, s_i1, s_b, lane);
        if (dof_vel) lerp_tile(v.dof_vel, dof_vel + f0 * D, D, dv.dof, nf, s_i0, s_i1, s_b, lane);
        if (body_pos) lerp_tile(v.body_pos, body_pos + f0 * B * 3, B * 3, dv.body3, nf, s_i0, s_i1, s_b, lane);
        if (body_lin) lerp_tile(v.body_lin, body_lin + f0 * B * 3, B * 3, dv.body3, nf, s_i0, s_i1, s_b, lane);
        if (body_ang) lerp_tile(v.body_ang, body_ang + f0 * B * 3, B * 3, dv.body3, nf, s_i0, s_i1, s_b, lane);
        if (body_rot) {
            const float4 *rot = reinterpret_cast<const float4 *>(v.body_rot);
            float4 *o = reinterpret_cast<float4 *>(body_rot) + f0 * B;
            const int total = nf * B;
#pragma unroll 2
            for (int e = lane; e < total; e += 32) {
                const int fl = (int)__umulhi((uint32_t)e, dv.body);
                const int j = e - fl * B;
                __stcs(o + e, slerp(__ldg(rot + (int64_t)s_i0[fl] * B + j), __ldg(rot + (int64_t)s_i1[fl] * B + j), s_b[fl]));
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Row-table variant of amp_sample_full (used when the library owns a "lerp row" table, i.e. for MotionLoader.sample):
// the five linearly interpolated tensors of a frame are staged side by side as ONE row
//   [dof_pos D | dof_vel D | body_pos 3B | body_lin 3B | body_ang 3B]      (W = 2D + 9B floats)
// so the lerp part becomes the same "lane = column, two row reads, one store" loop as the fused AMP kernel: per frame
// ceil(W/32) slots instead of a per-element divide + three shared-memory look-ups.  Where a column goes (which output
// tensor, which offset, which row pitch) is frame-invariant and hoisted.  Rotations keep the flat (frame, body) walk.
// ------------------------------------------------------------------------------------------------------------------
__global__ void pack_lerp_rows_kernel(LibView v, float *__restrict__ rows) {
    const int D = v.num_dofs, B3 = v.num_bodies * 3, W = v.lerp_width, Ws = v.lerp_stride;
    const int64_t total = v.num_frames * (int64_t)Ws;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = e / Ws;
        const int c = (int)(e - f * Ws);
        float val = 0.0f;
        if (c < D) val = v.dof_pos[f * D + c];
        else if (c < 2 * D) val = v.dof_vel[f * D + c - D];
        else if (c < 2 * D + B3) val = v.body_pos[f * B3 + c - 2 * D];
        else if (c < 2 * D + 2 * B3) val = v.body_lin[f * B3 + c - 2 * D - B3];
        else if (c < W) val = v.body_ang[f * B3 + c - 2 * D - 2 * B3];
        rows[e] = val;
    }
}

template <int NSLOT>
__global__ void __launch_bounds__(kSampleWarps * 32) sample_rows_kernel(LibView v, uint32_t magic_body,
                                                                         const double *__restrict__ times,
                                                                         const int64_t *__restrict__ ids, int64_t S,
                                                                         float *__restrict__ dof_pos, float *__restrict__ dof_vel,
                                                                         float *__restrict__ body_pos, float *__restrict__ body_rot,
                                                                         float *__restrict__ body_lin, float *__restrict__ body_ang) {
    __shared__ int s_i0_all[kSampleWarps][32], s_i1_all[kSampleWarps][32];
    __shared__ float s_b_all[kSampleWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *s_i0 = s_i0_all[warp], *s_i1 = s_i1_all[warp];
    float *s_b = s_b_all[warp];
    const int D = v.num_dofs, B = v.num_bodies, B3 = 3 * B, W = v.lerp_width, Ws = v.lerp_stride;
    const float *__restrict__ rows = v.lerp_rows;

    float *dst[NSLOT];   // where column lane + 32 s of frame 0 goes
    int pitch[NSLOT];    // floats between consecutive frames in that output tensor
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int c = lane + 32 * s;
        dst[s] = nullptr;
        pitch[s] = 0;
        if (c < D) { dst[s] = dof_pos + c; pitch[s] = D; }
        else if (c < 2 * D) { dst[s] = dof_vel + (c - D); pitch[s] = D; }
        else if (c < 2 * D + B3) { dst[s] = body_pos + (c - 2 * D); pitch[s] = B3; }
        else if (c < 2 * D + 2 * B3) { dst[s] = body_lin + (c - 2 * D - B3); pitch[s] = B3; }
        else if (c < W) { dst[s] = body_ang + (c - 2 * D - 2 * B3); pitch[s] = B3; }
    }

    const int64_t num_tiles = (S + 31) / 32;
    for (int64_t tile = blockIdx.x * (int64_t)kSampleWarps + warp; tile < num_tiles; tile += (int64_t)gridDim.x * kSampleWarps) {
        const int64_t f0 = tile * 32;
        const int nf = (int)min((int64_t)32, S - f0);
        if (lane < nf) {
            const FrameBlend fb = lookup_frame(v, times[f0 + lane], ids ? ids[f0 + lane] : 0);
            s_i0[lane] = (int)fb.i0;
            s_i1[lane] = (int)fb.i1;
            s_b[lane] = __double2float_rn(fb.blend);
        }
        __syncwarp();
        constexpr int G = NSLOT < 4 ? NSLOT : 4;  // slots per batch: all loads of a batch are issued before its stores
#pragma unroll 1
        for (int f = 0; f < nf; ++f) {
            const float b = s_b[f], omb = __fsub_rn(1.0f, b);
            const float *p0 = rows + (int64_t)s_i0[f] * Ws + lane, *p1 = rows + (int64_t)s_i1[f] * Ws + lane;
#pragma unroll
            for (int g = 0; g < NSLOT; g += G) {  // the table is padded: rows may be over-read, only stores are guarded
                float a0[G], a1[G];
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    if (g + t < NSLOT) {
                        a0[t] = __ldg(p0 + 32 * (g + t));
                        a1[t] = __ldg(p1 + 32 * (g + t));
                    }
                }
#pragma unroll
                for (int t = 0; t < G; ++t)
                    if (g + t < NSLOT && dst[g + t]) __stcs(dst[g + t] + (f0 + f) * pitch[g + t], lerp_w(omb, b, a0[t], a1[t]));
            }
        }
        {
            const float4 *rot = reinterpret_cast<const float4 *>(v.body_rot);
            float4 *o = reinterpret_cast<float4 *>(body_rot) + f0 * B;
            const int total = nf * B;
#pragma unroll 2
            for (int e = lane; e < total; e += 32) {
                const int fl = (int)__umulhi((uint32_t)e, magic_body);
                const int j = e - fl * B;
                __stcs(o + e, slerp(__ldg(rot + (int64_t)s_i0[fl] * B + j), __ldg(rot + (int64_t)s_i1[fl] * B + j), s_b[fl]));
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// amp_lerp / amp_slerp with explicit end points
// ------------------------------------------------------------------------------------------------------------------
__global__ void lerp_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ blend,
                            int64_t total, int64_t inner, float *__restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = lerp(blend[e / inner], a[e], b[e]);
}

__global__ void slerp_kernel(const float4 *__restrict__ q0, const float4 *__restrict__ q1,
                             const float *__restrict__ blend, int64_t total, int64_t bodies, float4 *__restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = slerp(q0[e], q1[e], blend[e / bodies]);
}

// ------------------------------------------------------------------------------------------------------------------
// amp_collect_reference: fused history times -> frame/blend -> gather + lerp -> root slerp -> compute_obs -> stacked row.
//
// A warp owns a tile of consecutive samples (tile_samples * K <= tile_cap frames).
//   phase 1  one lane per frame: float64 index math, root quaternion slerp, tangent/normal, root position, and the
//            destination offset; staged as a 64-byte FrameMeta record in the warp's slice of shared memory;
//   phase 2  the warp walks its frames; lane l produces columns l, l+32, l+64, ... so the two packed-row reads and the
//            store are contiguous 128-byte runs.  Which lanes hold the tangent/normal and key-body columns does not
//            depend on the frame, so that classification is hoisted out of the loop and the loop body is branch-free.
// Only __syncwarp is needed inside the loop: warps never share data.
//
// Two variants (template SMEM_TABLE):
//   true   the whole packed table (G1_walk 134 KB, G1_dance 202 KB) is copied once per CTA into shared memory and one
//          persistent CTA per SM (up to 32 warps) streams rows out of it: the gathers become conflict-free LDS with
//          32-bit addresses and L1/L2 only carry the output stream.  ncu on the global-table version showed the L1 data
//          pipe at 84 % (two tag wavefronts per misaligned 128-byte row read) -- that is the limiter this removes.
//   false  table stays in global memory (L1/L2 resident): for pooled libraries that do not fit in 227 KB and for
//          small batches where copying the table per CTA would cost more than the gathers.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kCollectWarps = 8;     // warps per CTA of the global-table variant
constexpr int kMaxHistory = 64;      // largest num_amp_observations the fused kernel accepts
constexpr int kTileFrames = 32;      // frames per warp tile (one phase-1 pass) when K <= 32
constexpr int kMaxSmemOptin = 232448;  // 227 KiB per CTA on sm_100

// 16 floats per frame; indices into the record viewed as float[16]
struct __align__(16) FrameMeta {
    int32_t off0, off1;  // [0] [1]   float offsets of the two packed rows
    float b, omb;        // [2] [3]   blend, 1 - blend
    float tn[6];         // [4..9]    tangent, normal of the slerped root rotation
    float root[3];       // [10..12]  interpolated root position (x, y, z)
    float zero;          // [13]      0.0f: the "nothing to subtract" operand of the branch-free column fix-up
    int64_t out;         // [14] [15] float offset of this frame's A columns in the destination
};
static_assert(sizeof(FrameMeta) == 64, "FrameMeta must stay one 64-byte record");

// Table reads.  The shared-memory variant addresses the table with explicit 32-bit shared-space addresses: through a generic
// pointer nvcc 12.9 rebuilds the shared window base (S2R SR_CgaCtaId + MOV + LEA) in front of every group of loads -- ncu
// showed those S2R among the most-stalled instructions of the row loop.
template <bool SMEM_TABLE>
__device__ __forceinline__ float table_ld(const float *p) {
    if constexpr (SMEM_TABLE) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)(uintptr_t)p));
        return v;
    } else {
        return __ldg(p);
    }
}

#ifdef AMP_COLLECT_PROFILE
// developer builds only: 0 = normal, 1 = phase 1 only, 2 = phase 2 only (cheap metadata), 3 = phase 1 without the slerp
__constant__ int c_collect_mode;
#endif

template <int NSLOT, bool SMEM_TABLE>
__global__ void __launch_bounds__(SMEM_TABLE ? 1024 : kCollectWarps * 32, SMEM_TABLE ? 1 : 4)
collect_reference_kernel(LibView v, const double *__restrict__ cur_times, const int64_t *__restrict__ ids, int64_t n,
                         int K, int tile_samples, int tile_cap, float *__restrict__ out, int64_t row_stride,
                         int64_t capacity, int64_t start_row, const int64_t *__restrict__ row_index, int64_t num_tiles) {
    extern __shared__ __align__(16) unsigned char collect_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const int D2 = 2 * v.obs_dofs, A = v.obs_width, R = v.row_floats;

    const float *tab = v.packed;
    unsigned char *meta_base = collect_smem;
    const int Rs = R;  // row stride of the table the kernel reads (an odd smem stride was tried: no gain, fewer warps fit)
    if constexpr (SMEM_TABLE) {
        const int quads = (int)(v.num_frames * R / 4);
        float4 *dst = reinterpret_cast<float4 *>(collect_smem);
        const float4 *src = reinterpret_cast<const float4 *>(v.packed);
        for (int i = threadIdx.x; i < quads; i += blockDim.x) dst[i] = __ldg(src + i);
        // NOT a dereferenceable pointer: the 32-bit shared-space address of the table, carried in a pointer-typed variable so
        // that both variants share the offset arithmetic below (table_ld<true> reads it with ld.shared)
        tab = reinterpret_cast<const float *>((uintptr_t)__cvta_generic_to_shared(collect_smem));
        // the last row may be over-read by 32*NSLOT - R floats (values discarded): keep that slack inside the allocation
        meta_base = collect_smem + (size_t)quads * 16 + 512;
        __syncthreads();
    }
    FrameMeta *meta = reinterpret_cast<FrameMeta *>(meta_base) + warp * tile_cap;

    // Frame-invariant classification of this lane's columns (column c = lane + 32*s):
    //   plain columns      value = lerp - 0.0f                 (x - 0.0f == x bit for bit)
    //   key-body columns   value = lerp - root[axis]           (g1_amp_env.py:552, interpolated key minus interpolated root)
    //   tangent/normal     value = tn[i]
    // One shared-memory operand w = meta[fix_idx] per special slot serves all three; tn_mask selects w itself.
    int fix_idx[NSLOT];
    uint32_t tn_mask[NSLOT];
    bool special[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int c = lane + 32 * s;
        const bool is_tn = c > D2 && c < D2 + 7;
        const bool is_key = c >= D2 + 13 && c < A;
        fix_idx[s] = is_tn ? 4 + c - (D2 + 1) : (is_key ? 10 + (c - (D2 + 13)) % 3 : 13);
        tn_mask[s] = is_tn ? 0xffffffffu : 0u;
        special[s] = __any_sync(0xffffffffu, is_tn || is_key);  // warp-uniform: does this slot need the fix-up at all?
    }
    const bool last_active = lane + 32 * (NSLOT - 1) < A;  // only the last slot can have lanes past the row end

    for (int64_t tile = blockIdx.x * (int64_t)warps + warp; tile < num_tiles; tile += (int64_t)gridDim.x * warps) {
        const int64_t s0 = tile * tile_samples;
        const int ns = (int)min((int64_t)tile_samples, n - s0);
        const int nf = ns * K;

        // ---- phase 1: per-frame scalars ------------------------------------------------------------------------
        for (int fl = lane; fl < nf; fl += 32) {
            const int si = fl / K, k = fl - si * K;
            const int64_t sample = s0 + si;
            // g1_amp_env.py:454-457  t - dt*k, float64, never clamped (negative -> extrapolation)
            const double t = __dsub_rn(cur_times[sample], __dmul_rn(v.dt, (double)k));
#ifdef AMP_COLLECT_PROFILE
            if (c_collect_mode == 2) {
                FrameMeta m;
                m.off0 = (int32_t)(((sample * 7 + 1 - k) % (v.num_frames - 1)) * Rs);
                m.off1 = m.off0 + Rs;
                m.b = 0.25f; m.omb = 0.75f;
                for (int i = 0; i < 6; ++i) m.tn[i] = 0.5f;
                m.root[0] = m.root[1] = m.root[2] = 0.1f;
                m.zero = 0.0f;
                m.out = (start_row + sample) * row_stride + (int64_t)k * A;
                meta[fl] = m;
                continue;
            }
#endif
            const FrameBlend fb = lookup_frame(v, t, ids ? ids[sample] : 0);
            const float b = __double2float_rn(fb.blend), omb = __fsub_rn(1.0f, b);
            FrameMeta m;
            m.off0 = (int32_t)(fb.i0 * Rs);
            m.off1 = (int32_t)(fb.i1 * Rs);
            m.b = b;
            m.omb = omb;
            const float *r0 = tab + m.off0, *r1 = tab + m.off1;
            const float4 q0 = make_float4(table_ld<SMEM_TABLE>(r0 + D2 + 1), table_ld<SMEM_TABLE>(r0 + D2 + 2),
                                          table_ld<SMEM_TABLE>(r0 + D2 + 3), table_ld<SMEM_TABLE>(r0 + D2 + 4));
            const float4 q1 = make_float4(table_ld<SMEM_TABLE>(r1 + D2 + 1), table_ld<SMEM_TABLE>(r1 + D2 + 2),
                                          table_ld<SMEM_TABLE>(r1 + D2 + 3), table_ld<SMEM_TABLE>(r1 + D2 + 4));
#ifdef AMP_COLLECT_PROFILE
            if (c_collect_mode == 3) tangent_normal(make_float4(q0.x * omb, q1.y * b, q0.z, q1.w), m.tn);
            else
#endif
            tangent_normal(slerp(q0, q1, b), m.tn);
            m.root[0] = lerp_w(omb, b, table_ld<SMEM_TABLE>(r0 + D2 + 5), table_ld<SMEM_TABLE>(r1 + D2 + 5));
            m.root[1] = lerp_w(omb, b, table_ld<SMEM_TABLE>(r0 + D2 + 6), table_ld<SMEM_TABLE>(r1 + D2 + 6));
            m.root[2] = lerp_w(omb, b, table_ld<SMEM_TABLE>(r0 + D2), table_ld<SMEM_TABLE>(r1 + D2));
            m.zero = 0.0f;
            int64_t row;
            if (row_index) {
                row = row_index[sample];
            } else {
                row = start_row + sample;
                if (capacity > 0) row %= capacity;
            }
            m.out = row * row_stride + (int64_t)k * A;
            meta[fl] = m;
        }
        __syncwarp();
#ifdef AMP_COLLECT_PROFILE
        if (c_collect_mode == 1 || c_collect_mode == 3) {
            if (lane < nf) out[meta[lane].out] = meta[lane].tn[0] + meta[lane].root[2];
            __syncwarp();
            continue;
        }
#endif

        // ---- phase 2: stream the rows out ----------------------------------------------------------------------
        // Lane l reads table[off + l + 32 s] unconditionally (the allocation is padded so the last row may be
        // over-read) and only the store of the last slot is predicated.
        // Consecutive history frames of a sample are one frame period apart, so frame k+1 normally interpolates rows
        // (i0-1, i0) where frame k used (i0, i0+1): the lower row of the previous frame is kept in registers and becomes
        // the upper row (11 row reads instead of 20 per sample at K = 10).  The test is warp-uniform (the offsets come
        // from the broadcast FrameMeta read) and only skips loads -- the values are the same either way.
        float lo[NSLOT], hi[NSLOT];
        int prev_off0 = -1, prev_off1 = -1;
#pragma unroll 2
        for (int f = 0; f < nf; ++f) {
            const float *mf = reinterpret_cast<const float *>(&meta[f]);
            const int4 head = *reinterpret_cast<const int4 *>(mf);  // off0, off1, b, omb in one broadcast read
            const float b = __int_as_float(head.z), omb = __int_as_float(head.w);
            float *o = out + *reinterpret_cast<const int64_t *>(mf + 14) + lane;
            if (head.y == prev_off0) {
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) hi[s] = lo[s];
            } else if (head.y != prev_off1) {
                const float *p1 = tab + head.y + lane;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) hi[s] = table_ld<SMEM_TABLE>(p1 + 32 * s);
            }
            if (head.x != prev_off0) {
                const float *p0 = tab + head.x + lane;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) lo[s] = table_ld<SMEM_TABLE>(p0 + 32 * s);
            }
            prev_off0 = head.x;
            prev_off1 = head.y;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) {
                float val = lerp_w(omb, b, lo[s], hi[s]);
                if (special[s]) {  // warp-uniform branch
                    const float w = mf[fix_idx[s]];
                    const float d = __fsub_rn(val, w);
                    val = __uint_as_float((__float_as_uint(w) & tn_mask[s]) | (__float_as_uint(d) & ~tn_mask[s]));
                }
                if (s < NSLOT - 1 || last_active) __stcs(o + 32 * s, val);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// compute_obs on caller tensors (one warp per row) and the per-step env variant with in-place history shift.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pick6(const float tn[6], int i) {
    // register array indexed by a lane-dependent value: select chain instead of local memory
    float r = tn[0];
    r = i == 1 ? tn[1] : r;
    r = i == 2 ? tn[2] : r;
    r = i == 3 ? tn[3] : r;
    r = i == 4 ? tn[4] : r;
    r = i == 5 ? tn[5] : r;
    return r;
}

// compute_obs on seven contiguous caller tensors: one warp per row, lane l owns columns l, l+32, ...; where a column
// comes from is row-invariant, so every slot gets a (pointer, per-row stride) pair once and the row loop is branch-free.
template <int NSLOT>
__global__ void __launch_bounds__(256) compute_obs_kernel(const float *__restrict__ dof_pos,
                                                           const float *__restrict__ dof_vel,
                                                           const float *__restrict__ root_pos,
                                                           const float *__restrict__ root_rot,
                                                           const float *__restrict__ root_lin,
                                                           const float *__restrict__ root_ang,
                                                           const float *__restrict__ key_pos, int64_t n, int D, int Kb,
                                                           float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int D2 = 2 * D, A = D2 + 13 + 3 * Kb;
    const float *src[NSLOT], *sub[NSLOT];
    int stride[NSLOT], tn_idx[NSLOT];
    bool active[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int c = lane + 32 * s;
        active[s] = c < A;
        tn_idx[s] = -1;
        sub[s] = nullptr;
        src[s] = root_pos;
        stride[s] = 0;
        if (c < D) { src[s] = dof_pos + c; stride[s] = D; }
        else if (c < D2) { src[s] = dof_vel + (c - D); stride[s] = D; }
        else if (c == D2) { src[s] = root_pos + 2; stride[s] = 3; }
        else if (c < D2 + 7) { tn_idx[s] = c - D2 - 1; }
        else if (c < D2 + 10) { src[s] = root_lin + (c - D2 - 7); stride[s] = 3; }
        else if (c < D2 + 13) { src[s] = root_ang + (c - D2 - 10); stride[s] = 3; }
        else if (c < A) {
            const int e = c - D2 - 13;
            src[s] = key_pos + e;
            stride[s] = 3 * Kb;
            sub[s] = root_pos + e % 3;
        }
    }
    for (int64_t i = warp; i < n; i += nwarps) {
        const float4 q = make_float4(__ldg(root_rot + i * 4), __ldg(root_rot + i * 4 + 1), __ldg(root_rot + i * 4 + 2),
                                     __ldg(root_rot + i * 4 + 3));
        float val[NSLOT], minus[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            val[s] = (active[s] && tn_idx[s] < 0) ? __ldg(src[s] + i * stride[s]) : 0.0f;
            minus[s] = sub[s] ? __ldg(sub[s] + i * 3) : 0.0f;
        }
        float tn[6];
        tangent_normal(q, tn);
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            if (!active[s]) continue;
            float r = sub[s] ? __fsub_rn(val[s], minus[s]) : val[s];
            if (tn_idx[s] >= 0) r = pick6(tn, tn_idx[s]);
            __stcs(out + i * A + lane + 32 * s, r);
        }
    }
}

__global__ void tangent_normal_kernel(const float *__restrict__ q, int64_t n, float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float tn[6];
        tangent_normal(make_float4(q[i * 4], q[i * 4 + 1], q[i * 4 + 2], q[i * 4 + 3]), tn);
#pragma unroll
        for (int j = 0; j < 6; ++j) out[i * 6 + j] = tn[j];
    }
}

// One warp per env, lane l owns columns l, l+32, ...  Which simulator tensor feeds a column does not depend on the env, so
// every slot gets a (pointer, per-env stride) pair once and the env loop is: issue all state loads and all history loads,
// then all stores.  History: slot s -> s+1 for s = K-2 .. 0 (oldest first, eight slots per trip: nothing is overwritten
// before it is read; a lane touches only its own columns, so there is no cross-lane hazard), then slot 0 = the new row.
// Reference g1_amp_env.py:176-193.
// Register budget: at 64 registers (4 CTAs/SM) the 3-slot instantiation (A = 81 / 83, every shipped robot) spilled its loaded
// values to local memory inside the env loop -- ncu: 30 % of all stall samples on one STL waiting for the load it spills, i.e.
// one load in flight at a time.  128 registers (2 CTAs/SM) keep every load of an env in flight together.
template <int NSLOT>
__global__ void __launch_bounds__(256, 2)
obs_step_kernel(const float *__restrict__ joint_pos, const float *__restrict__ joint_vel,
                const float *__restrict__ body_pos, const float *__restrict__ body_quat,
                const float *__restrict__ body_lin, const float *__restrict__ body_ang, int64_t N, int D, int Bsim,
                int ref, KeyList keys, int Kb, int K, float *__restrict__ amp_buf, float *__restrict__ policy_obs,
                int64_t policy_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int D2 = 2 * D, A = D2 + 13 + 3 * Kb, policy_width = A - 3 * Kb;

    const float *src[NSLOT], *sub[NSLOT];  // value = src[i*stride] (- sub[i*3*Bsim] for key-body columns)
    int stride[NSLOT];
    int tn_idx[NSLOT];                     // >= 0: the column is tangent/normal component tn_idx
    bool active[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int c = lane + 32 * s;
        active[s] = c < A;
        tn_idx[s] = -1;
        sub[s] = nullptr;
        src[s] = joint_pos;
        stride[s] = 0;
        if (c < D) { src[s] = joint_pos + c; stride[s] = D; }
        else if (c < D2) { src[s] = joint_vel + (c - D); stride[s] = D; }
        else if (c == D2) { src[s] = body_pos + ref * 3 + 2; stride[s] = Bsim * 3; }
        else if (c < D2 + 7) { tn_idx[s] = c - D2 - 1; }
        else if (c < D2 + 10) { src[s] = body_lin + ref * 3 + (c - D2 - 7); stride[s] = Bsim * 3; }
        else if (c < D2 + 13) { src[s] = body_ang + ref * 3 + (c - D2 - 10); stride[s] = Bsim * 3; }
        else if (c < A) {
            const int e = c - D2 - 13, j = e / 3, a = e - 3 * j;
            src[s] = body_pos + keys.k[j] * 3 + a;
            sub[s] = body_pos + ref * 3 + a;
            stride[s] = Bsim * 3;
        }
    }

    constexpr int HB = 4;   // history slots moved per trip, all loads ahead of the stores
    // envs in flight per warp: the loads of both are issued before the first store (memory-level parallelism is what bounds
    // this kernel: ~10 narrow loads per env, DRAM latency each); wide observations (> 128 columns) keep one env per trip,
    // two would not fit the register file
    constexpr bool TWO = NSLOT <= 4;
    constexpr int EPW = TWO ? 2 : 1;
    auto load_inputs = [&](int64_t i, float4 &q, float(&val)[NSLOT], float(&minus)[NSLOT]) {
        q = __ldg(reinterpret_cast<const float4 *>(body_quat) + i * Bsim + ref);
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            val[s] = (active[s] && tn_idx[s] < 0) ? __ldg(src[s] + i * stride[s]) : 0.0f;
            minus[s] = sub[s] ? __ldg(sub[s] + i * stride[s]) : 0.0f;
        }
    };
    auto load_history = [&](const float *env, int hi, int lo, float(&h)[HB][NSLOT]) {
#pragma unroll
        for (int t = 0; t < HB; ++t)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (hi - t >= lo && active[s]) h[t][s] = env[(int64_t)(hi - t) * A + 32 * s];
    };
    auto store_history = [&](float *env, int hi, int lo, const float(&h)[HB][NSLOT]) {
#pragma unroll
        for (int t = 0; t < HB; ++t)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (hi - t >= lo && active[s]) env[(int64_t)(hi - t + 1) * A + 32 * s] = h[t][s];
    };
    auto write_slot0 = [&](int64_t i, float *env, const float4 &q, const float(&val)[NSLOT], const float(&minus)[NSLOT]) {
        float tn[6];
        tangent_normal(q, tn);
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            if (!active[s]) continue;
            float r = sub[s] ? __fsub_rn(val[s], minus[s]) : val[s];
            if (tn_idx[s] >= 0) r = pick6(tn, tn_idx[s]);
            env[32 * s] = r;
            if (policy_obs && lane + 32 * s < policy_width) policy_obs[i * policy_stride + lane + 32 * s] = r;
        }
    };

    for (int64_t i0 = warp; i0 < N; i0 += EPW * nwarps) {
        const int64_t i1 = i0 + nwarps;
        const bool two = TWO && i1 < N;
        const int64_t j1 = two ? i1 : i0;  // second env of the trip (aliases the first when there is none: loads only)
        float4 q0, q1;
        float val0[NSLOT], minus0[NSLOT], val1[NSLOT], minus1[NSLOT];
        load_inputs(i0, q0, val0, minus0);
        if constexpr (TWO) load_inputs(j1, q1, val1, minus1);
        float *env0 = amp_buf + i0 * (int64_t)K * A + lane;
        float *env1 = amp_buf + j1 * (int64_t)K * A + lane;
        // history shift, oldest slots first.  The first trip of both envs is loaded before anything is stored (K <= 5: the
        // only trip); longer histories continue env by env.
        int hi = K - 2;
        if (hi >= 0) {
            const int lo = max(hi - (HB - 1), 0);
            float h0[HB][NSLOT], h1[HB][NSLOT];
            load_history(env0, hi, lo, h0);
            if constexpr (TWO) load_history(env1, hi, lo, h1);
            store_history(env0, hi, lo, h0);
            if constexpr (TWO) {
                if (two) store_history(env1, hi, lo, h1);
            }
            for (hi -= HB; hi >= 0; hi -= HB) {
                const int lo2 = max(hi - (HB - 1), 0);
                load_history(env0, hi, lo2, h0);
                if constexpr (TWO) load_history(env1, hi, lo2, h1);
                store_history(env0, hi, lo2, h0);
                if constexpr (TWO) {
                    if (two) store_history(env1, hi, lo2, h1);
                }
            }
        }
        write_slot0(i0, env0, q0, val0, minus0);
        if constexpr (TWO) {
            if (two) write_slot0(i1, env1, q1, val1, minus1);
        }
    }
}

// Actor observation + its history (g1_amp_env.py:195-242): one warp per env, lanes stride the columns.
__global__ void __launch_bounds__(256)
actor_obs_kernel(const float *__restrict__ amp_buf, int64_t N, int K, int A, int base, const float *__restrict__ last_actions,
                 int act, const float *__restrict__ command, int cmd, int n_hist /* n - 1 */, int inc_act, int inc_cmd,
                 float *__restrict__ hist_buf, uint8_t *__restrict__ just_reset, float *__restrict__ actor_obs,
                 int64_t actor_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int cur = base + act + cmd;
    const int hact = inc_act ? act : 0, hcmd = inc_cmd ? cmd : 0, P = base + hact + hcmd;
    for (int64_t i = warp; i < N; i += nwarps) {
        const float *obs0 = amp_buf + i * (int64_t)K * A;  // slot 0 = newest
        float *arow = actor_obs + i * actor_stride;
        for (int c = lane; c < cur; c += 32) {
            float val;
            if (c < base) val = obs0[c];
            else if (c < base + act) val = __ldg(last_actions + i * act + (c - base));
            else val = __ldg(command + i * cmd + (c - base - act));
            arow[c] = val;
        }
        if (n_hist <= 0) continue;
        const bool reset = just_reset && just_reset[i] != 0;
        float *hist = hist_buf + i * (int64_t)n_hist * P;
        for (int pc = lane; pc < P; pc += 32) {
            float frame;
            if (pc < base) frame = obs0[pc];
            else if (pc < base + hact) frame = __ldg(last_actions + i * act + (pc - base));
            else frame = __ldg(command + i * cmd + (pc - base - hact));
            // slots from the oldest down: slot s takes slot s-1 (or the new frame everywhere after a reset)
            for (int sl = n_hist - 1; sl >= 1; --sl) {
                const float moved = reset ? frame : hist[(int64_t)(sl - 1) * P + pc];
                hist[(int64_t)sl * P + pc] = moved;
                arow[cur + (int64_t)sl * P + pc] = moved;
            }
            hist[pc] = frame;
            arow[cur + pc] = frame;
        }
        __syncwarp();
        if (reset && lane == 0) just_reset[i] = 0;
    }
}

// Task reward (g1_amp_env.py:246-288, 500-532, 564-606): one warp per env; the per-joint sums are lane-strided partial sums
// combined with a fixed shuffle tree (deterministic), lane 0 finishes the scalar part.
struct RewardScales {
    float termination, action_l2, joint_pos_limits, joint_acc_l2, joint_vel_l2, track_vel;
    float exp_at_floor, linear_slope;  // weight*exp(-floor), weight/sigma^2*exp(-floor), evaluated in double on the host
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// the per-joint sums of one env (whole warp): lane-strided partial sums combined with a fixed shuffle tree.  Loads only --
// the fused env-step kernel issues them at the top of an env's trip, together with the observation loads.
struct RewardSums {
    float act, lim, acc, vel;
};
__device__ __forceinline__ RewardSums task_reward_sums(int64_t i, int lane, const float *__restrict__ actions, int act,
                                                       const float *__restrict__ joint_pos, const float *__restrict__ limits,
                                                       const float *__restrict__ joint_acc, const float *__restrict__ joint_vel, int D) {
    float s_act = 0.0f, s_lim = 0.0f, s_acc = 0.0f, s_vel = 0.0f;
    for (int j = lane; j < act; j += 32) {
        const float a = __ldg(actions + i * act + j);
        s_act += a * a;
    }
    for (int j = lane; j < D; j += 32) {
        const float p = __ldg(joint_pos + i * D + j);
        const float2 lim = __ldg(reinterpret_cast<const float2 *>(limits) + i * D + j);
        s_lim += -fminf(p - lim.x, 0.0f) + fmaxf(p - lim.y, 0.0f);
        const float a = __ldg(joint_acc + i * D + j), w = __ldg(joint_vel + i * D + j);
        s_acc += a * a;
        s_vel += w * w;
    }
    RewardSums r;
    r.act = warp_sum(s_act);
    r.lim = warp_sum(s_lim);
    r.acc = warp_sum(s_acc);
    r.vel = warp_sum(s_vel);
    return r;
}

__device__ __forceinline__ void task_reward_finish(const RewardScales &sc, const RewardSums &sums, int64_t i, int lane,
                                                   const uint8_t *__restrict__ terminated, const float *__restrict__ body_lin,
                                                   const float *__restrict__ body_quat, int Bsim, int ref,
                                                   const float *__restrict__ command, float *__restrict__ total,
                                                   float *__restrict__ terms, float *__restrict__ track_err);

// one env, whole warp: sums, then lane 0 finishes
__device__ __forceinline__ void task_reward_env(const RewardScales &sc, int64_t i, int lane, const uint8_t *__restrict__ terminated,
                                                const float *__restrict__ actions, int act, const float *__restrict__ joint_pos,
                                                const float *__restrict__ limits, const float *__restrict__ joint_acc,
                                                const float *__restrict__ joint_vel, int D, const float *__restrict__ body_lin,
                                                const float *__restrict__ body_quat, int Bsim, int ref,
                                                const float *__restrict__ command, float *__restrict__ total,
                                                float *__restrict__ terms, float *__restrict__ track_err) {
    const RewardSums sums = task_reward_sums(i, lane, actions, act, joint_pos, limits, joint_acc, joint_vel, D);
    task_reward_finish(sc, sums, i, lane, terminated, body_lin, body_quat, Bsim, ref, command, total, terms, track_err);
}

__device__ __forceinline__ void task_reward_finish(const RewardScales &sc, const RewardSums &sums, int64_t i, int lane,
                                                   const uint8_t *__restrict__ terminated, const float *__restrict__ body_lin,
                                                   const float *__restrict__ body_quat, int Bsim, int ref,
                                                   const float *__restrict__ command, float *__restrict__ total,
                                                   float *__restrict__ terms, float *__restrict__ track_err) {
    const float s_act = sums.act, s_lim = sums.lim, s_acc = sums.acc, s_vel = sums.vel;
    if (lane == 0) {
        const float r_term = sc.termination * (terminated[i] ? 1.0f : 0.0f);
        const float r_act = sc.action_l2 * s_act, r_lim = sc.joint_pos_limits * s_lim;
        const float r_acc = sc.joint_acc_l2 * s_acc, r_vel = sc.joint_vel_l2 * s_vel;
        float r_track = 0.0f, err = 0.0f;
        if (sc.track_vel > 0.0f) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(body_quat) + i * Bsim + ref);  // w x y z
            const float *vp = body_lin + (i * Bsim + ref) * 3;
            const float vx = __ldg(vp), vy = __ldg(vp + 1), vz = __ldg(vp + 2);
            // quat_rotate_inverse: v*(2w^2-1) - 2w (xyz x v) + 2 xyz (xyz . v); only x and y are needed
            const float k = 2.0f * (q.x * q.x) - 1.0f;
            const float cx = q.z * vz - q.w * vy, cy = q.w * vx - q.y * vz;  // (xyz x v).x, .y
            const float dotv = q.y * vx + q.z * vy + q.w * vz;
            const float bx = vx * k - cx * q.x * 2.0f + q.y * dotv * 2.0f;
            const float by = vy * k - cy * q.x * 2.0f + q.z * dotv * 2.0f;
            const float dx = bx - __ldg(command + i * 2), dy = by - __ldg(command + i * 2 + 1);
            err = sqrtf(dx * dx + dy * dy);
            const float e2 = err * err;
            // exp_reward_with_floor(e2, weight, sigma = 0.5, floor = 4.0): threshold = floor * sigma^2 = 1.0
            r_track = e2 > 1.0f ? sc.exp_at_floor - sc.linear_slope * (e2 - 1.0f) : sc.track_vel * expf(-e2 / 0.25f);
        }
        total[i] = ((((r_term + r_act) + r_lim) + r_acc) + r_vel) + r_track;
        if (terms) {
            float *t = terms + i * 6;
            t[0] = r_term; t[1] = r_act; t[2] = r_lim; t[3] = r_acc; t[4] = r_vel; t[5] = r_track;
        }
        if (track_err) track_err[i] = err;
    }
}

__global__ void __launch_bounds__(256)
task_reward_kernel(RewardScales sc, const uint8_t *__restrict__ terminated, const float *__restrict__ actions, int act,
                   const float *__restrict__ joint_pos, const float *__restrict__ limits, const float *__restrict__ joint_acc,
                   const float *__restrict__ joint_vel, int D, const float *__restrict__ body_lin, const float *__restrict__ body_quat,
                   int Bsim, int ref, const float *__restrict__ command, int64_t N, float *__restrict__ total,
                   float *__restrict__ terms, float *__restrict__ track_err) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < N; i += nwarps)
        task_reward_env(sc, i, lane, terminated, actions, act, joint_pos, limits, joint_acc, joint_vel, D, body_lin, body_quat, Bsim,
                        ref, command, total, terms, track_err);
}

// ------------------------------------------------------------------------------------------------------------------
// amp_env_step: the whole per-step env path in ONE launch: compute_obs from simulator state, AMP
// history shift + slot 0 (g1_amp_env.py:176-193), actor observation + its history with warm start (:195-242) and,
// optionally, the task reward (:246-319) -- one warp per env, lane l owns observation columns l, l+32, ...; the new
// observation row lives in registers and feeds the AMP buffer, the actor observation and the actor history without being
// read back; every simulator tensor is fetched from DRAM once per env.
// ------------------------------------------------------------------------------------------------------------------
struct EnvStepParams {
    const float *joint_pos, *joint_vel, *body_pos, *body_quat, *body_lin, *body_ang;
    int64_t N;
    int D, Bsim, ref, Kb, K;
    KeyList keys;
    float *amp_buf;
    // actor observation
    const float *last_actions, *command;
    int act, cmd, n_hist, inc_act, inc_cmd;
    float *hist_buf;
    uint8_t *just_reset;
    float *actor_obs;
    int64_t actor_stride;
    // task reward (total == NULL: skipped)
    RewardScales sc;
    const uint8_t *terminated;
    const float *actions, *limits, *joint_acc;
    float *total, *terms, *track_err;
};

template <int NSLOT>
__global__ void __launch_bounds__(256, NSLOT <= 3 ? 3 : 2) env_step_kernel(const __grid_constant__ EnvStepParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int D = p.D, D2 = 2 * D, Bsim = p.Bsim, ref = p.ref, K = p.K;
    const int A = D2 + 13 + 3 * p.Kb, base = A - 3 * p.Kb;

    const float *src[NSLOT], *sub[NSLOT];
    int stride[NSLOT], tn_idx[NSLOT];
    bool active[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int c = lane + 32 * s;
        active[s] = c < A;
        tn_idx[s] = -1;
        sub[s] = nullptr;
        src[s] = p.joint_pos;
        stride[s] = 0;
        if (c < D) { src[s] = p.joint_pos + c; stride[s] = D; }
        else if (c < D2) { src[s] = p.joint_vel + (c - D); stride[s] = D; }
        else if (c == D2) { src[s] = p.body_pos + ref * 3 + 2; stride[s] = Bsim * 3; }
        else if (c < D2 + 7) { tn_idx[s] = c - D2 - 1; }
        else if (c < D2 + 10) { src[s] = p.body_lin + ref * 3 + (c - D2 - 7); stride[s] = Bsim * 3; }
        else if (c < D2 + 13) { src[s] = p.body_ang + ref * 3 + (c - D2 - 10); stride[s] = Bsim * 3; }
        else if (c < A) {
            const int e = c - D2 - 13, j = e / 3, a = e - 3 * j;
            src[s] = p.body_pos + p.keys.k[j] * 3 + a;
            sub[s] = p.body_pos + ref * 3 + a;
            stride[s] = Bsim * 3;
        }
    }
    const int cur = base + p.act + p.cmd;
    const int hact = p.inc_act ? p.act : 0, hcmd = p.inc_cmd ? p.cmd : 0, P = base + hact + hcmd;
    constexpr int HB = 4;

    for (int64_t i = warp; i < p.N; i += nwarps) {
        // ---- all simulator-state loads of this env, then the history loads, before the first store ------------------
        // (the kernel's pointers may alias as far as the compiler knows, so nothing below a store is hoisted for us: the reward's
        // loads and the first slot of last_actions | command are issued here by hand, one DRAM round trip for the whole env)
        RewardSums sums{};
        if (p.total) sums = task_reward_sums(i, lane, p.actions, p.act, p.joint_pos, p.limits, p.joint_acc, p.joint_vel, D);
        float extra0 = 0.0f;  // column base + lane of the actor observation
        if (p.actor_obs && lane < p.act + p.cmd)
            extra0 = lane < p.act ? __ldg(p.last_actions + i * p.act + lane) : __ldg(p.command + i * p.cmd + (lane - p.act));
        const float4 q = __ldg(reinterpret_cast<const float4 *>(p.body_quat) + i * Bsim + ref);
        float val[NSLOT], minus[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            val[s] = (active[s] && tn_idx[s] < 0) ? __ldg(src[s] + i * stride[s]) : 0.0f;
            minus[s] = sub[s] ? __ldg(sub[s] + i * stride[s]) : 0.0f;
        }
        float *env = p.amp_buf + i * (int64_t)K * A + lane;
        for (int hi = K - 2; hi >= 0; hi -= HB) {  // AMP history, oldest slots first, HB slots per trip
            const int lo = max(hi - (HB - 1), 0);
            float h[HB][NSLOT];
#pragma unroll
            for (int t = 0; t < HB; ++t)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (hi - t >= lo && active[s]) h[t][s] = env[(int64_t)(hi - t) * A + 32 * s];
#pragma unroll
            for (int t = 0; t < HB; ++t)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (hi - t >= lo && active[s]) env[(int64_t)(hi - t + 1) * A + 32 * s] = h[t][s];
        }
        float tn[6];
        tangent_normal(q, tn);
        float row[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            float r = sub[s] ? __fsub_rn(val[s], minus[s]) : val[s];
            if (tn_idx[s] >= 0) r = pick6(tn, tn_idx[s]);
            row[s] = r;
            if (active[s]) env[32 * s] = r;
        }
        // ---- actor observation: [base | last_actions | command | history frames] ----------------------------------------
        if (p.actor_obs) {
            float *arow = p.actor_obs + i * p.actor_stride;
            const bool reset = p.n_hist > 0 && p.just_reset && p.just_reset[i] != 0;
            float *hist = p.n_hist > 0 ? p.hist_buf + i * (int64_t)p.n_hist * P : nullptr;
            auto history_column = [&](int pc, float frame) {  // slot s takes slot s-1 (or the new frame everywhere after a reset)
                for (int sl = p.n_hist - 1; sl >= 1; --sl) {
                    const float moved = reset ? frame : hist[(int64_t)(sl - 1) * P + pc];
                    hist[(int64_t)sl * P + pc] = moved;
                    arow[cur + (int64_t)sl * P + pc] = moved;
                }
                hist[pc] = frame;
                arow[cur + pc] = frame;
            };
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) {
                const int c = lane + 32 * s;
                if (c < base) {
                    arow[c] = row[s];
                    if (p.n_hist > 0) history_column(c, row[s]);
                }
            }
            for (int e = lane; e < p.act + p.cmd; e += 32) {
                const bool is_act = e < p.act;
                const float v = e == lane ? extra0 : (is_act ? __ldg(p.last_actions + i * p.act + e) : __ldg(p.command + i * p.cmd + (e - p.act)));
                arow[base + e] = v;
                if (p.n_hist > 0) {
                    if (is_act && p.inc_act) history_column(base + e, v);
                    if (!is_act && p.inc_cmd) history_column(base + hact + (e - p.act), v);
                }
            }
            if (reset) {
                __syncwarp();
                if (lane == 0) p.just_reset[i] = 0;
            }
        }
        // ---- task reward on the same state ----------------------------------------------------------------------------
        if (p.total)
            task_reward_finish(p.sc, sums, i, lane, p.terminated, p.body_lin, p.body_quat, Bsim, ref, p.command, p.total, p.terms, p.track_err);
    }
}

// ---- host helpers -------------------------------------------------------------------------------------------------
static int grid_for(int64_t items, int per_block, int ctas_per_sm) {
    const int64_t want = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
static int upload(amp_lib *lib, const T *host, size_t count, cudaStream_t st, T **dev) {
    void *p = nullptr;
    AMP_CUDA_TRY(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    lib->owned[lib->n_owned++] = p;
    if (count) AMP_CUDA_TRY(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *dev = static_cast<T *>(p);
    return AMP_OK;
}

}  // namespace amp

using namespace amp;

extern "C" {

int amp_lib_create(const amp_lib_desc_t *d, void *stream, amp_lib_t **out) {
    AMP_REQUIRE(d && out, "amp_lib_create: NULL argument");
    *out = nullptr;
    AMP_REQUIRE(d->num_frames > 0 && d->num_dofs > 0 && d->num_bodies > 0 && d->num_trajectories > 0,
                "amp_lib_create: empty library (F=%lld D=%d B=%d T=%d)", (long long)d->num_frames, d->num_dofs,
                d->num_bodies, d->num_trajectories);
    AMP_REQUIRE(d->traj_starts && d->traj_ends && d->durations, "amp_lib_create: NULL trajectory tables");
    AMP_REQUIRE(d->dof_positions && d->dof_velocities && d->body_positions && d->body_rotations &&
                    d->body_linear_velocities && d->body_angular_velocities,
                "amp_lib_create: NULL clip tensor");
    AMP_REQUIRE(aligned16(d->body_rotations), "amp_lib_create: body_rotations must be 16-byte aligned");
    AMP_REQUIRE(d->dt > 0.0, "amp_lib_create: dt must be positive");
    for (int t = 0; t < d->num_trajectories; ++t)
        AMP_REQUIRE(d->traj_starts[t] >= 0 && d->traj_ends[t] >= d->traj_starts[t] && d->traj_ends[t] < d->num_frames,
                    "amp_lib_create: trajectory %d spans [%lld, %lld] outside [0, %lld)", t,
                    (long long)d->traj_starts[t], (long long)d->traj_ends[t], (long long)d->num_frames);
    const bool with_env = d->num_obs_dofs > 0;
    if (with_env) {
        AMP_REQUIRE(d->dof_indexes && d->num_key_bodies >= 0 && d->num_key_bodies <= kMaxKeyBodies &&
                        (d->num_key_bodies == 0 || d->key_body_indexes),
                    "amp_lib_create: bad env selection (Kb=%d, max %d)", d->num_key_bodies, kMaxKeyBodies);
        AMP_REQUIRE(d->ref_body_index >= 0 && d->ref_body_index < d->num_bodies, "amp_lib_create: ref body %d out of range",
                    d->ref_body_index);
        for (int i = 0; i < d->num_obs_dofs; ++i)
            AMP_REQUIRE(d->dof_indexes[i] >= 0 && d->dof_indexes[i] < d->num_dofs, "amp_lib_create: dof index %d out of range",
                        d->dof_indexes[i]);
        for (int i = 0; i < d->num_key_bodies; ++i)
            AMP_REQUIRE(d->key_body_indexes[i] >= 0 && d->key_body_indexes[i] < d->num_bodies,
                        "amp_lib_create: key body index %d out of range", d->key_body_indexes[i]);
    }

    amp_lib *lib = new (std::nothrow) amp_lib();
    if (!lib) return fail(AMP_ENOMEM, "amp_lib_create: host allocation failed");
    std::memset(lib, 0, sizeof(*lib));
    cudaStream_t st = as_stream(stream);
    int rc = AMP_OK;
    auto bail = [&](int code) {
        amp_lib_destroy(lib);
        return code;
    };
    if (cudaGetDevice(&lib->device) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaGetDevice"));

    LibView &v = lib->v;
    v.num_frames = d->num_frames;
    v.num_dofs = d->num_dofs;
    v.num_bodies = d->num_bodies;
    v.num_traj = d->num_trajectories;
    v.dt = d->dt;
    v.dof_pos = d->dof_positions;
    v.dof_vel = d->dof_velocities;
    v.body_pos = d->body_positions;
    v.body_rot = d->body_rotations;
    v.body_lin = d->body_linear_velocities;
    v.body_ang = d->body_angular_velocities;

    int64_t *starts = nullptr, *ends = nullptr;
    double *durs = nullptr;
    if ((rc = upload(lib, d->traj_starts, (size_t)d->num_trajectories, st, &starts))) return bail(rc);
    if ((rc = upload(lib, d->traj_ends, (size_t)d->num_trajectories, st, &ends))) return bail(rc);
    if ((rc = upload(lib, d->durations, (size_t)d->num_trajectories, st, &durs))) return bail(rc);
    v.starts = starts;
    v.ends = ends;
    v.durations = durs;
    uint32_t zero = 0, *flags = nullptr;
    if ((rc = upload(lib, &zero, 1, st, &flags))) return bail(rc);
    v.flags = flags;

    if (with_env) {
        v.obs_dofs = d->num_obs_dofs;
        v.num_keys = d->num_key_bodies;
        v.obs_width = 2 * v.obs_dofs + 13 + 3 * v.num_keys;
        v.row_floats = (v.obs_width + 3) & ~3;
        if (v.num_frames * (int64_t)v.row_floats >= (int64_t)1 << 31) {
            set_error("amp_lib_create: packed table of %lld x %d floats exceeds 32-bit row offsets", (long long)v.num_frames,
                      v.row_floats);
            return bail(AMP_EINVAL);
        }
        int32_t *dof_idx = nullptr;
        if ((rc = upload(lib, d->dof_indexes, (size_t)v.obs_dofs, st, &dof_idx))) return bail(rc);
        void *packed = nullptr;
        // + 1 KiB: the fused kernel reads up to 32*NSLOT floats from a row start without a bounds check
        cudaError_t e = cudaMalloc(&packed, (size_t)v.num_frames * v.row_floats * sizeof(float) + 1024);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(packed rows)"));
        lib->owned[lib->n_owned++] = packed;
        v.packed = static_cast<float *>(packed);
        KeyList keys{};
        for (int i = 0; i < v.num_keys; ++i) keys.k[i] = d->key_body_indexes[i];
        const int64_t total = v.num_frames * (int64_t)v.row_floats;
        pack_rows_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(v, dof_idx, d->ref_body_index, keys,
                                                                  static_cast<float *>(packed));
        e = cudaGetLastError();
        if (e != cudaSuccess) return bail(cuda_fail(e, "pack_rows_kernel launch"));
    }
    if (with_env) {
        // the shared-memory-table variant needs the opt-in limit; set once per handle, here, not on the launch path
        // (cudaFuncSetAttribute is per device and idempotent, so concurrent handles need no guard)
        cudaError_t ea = cudaSuccess;
#define AMP_COLLECT_ATTR(NS)                                                                                            \
    if (ea == cudaSuccess)                                                                                              \
        ea = cudaFuncSetAttribute(collect_reference_kernel<NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin)
        AMP_COLLECT_ATTR(1); AMP_COLLECT_ATTR(2); AMP_COLLECT_ATTR(3); AMP_COLLECT_ATTR(4);
        AMP_COLLECT_ATTR(5); AMP_COLLECT_ATTR(6); AMP_COLLECT_ATTR(7); AMP_COLLECT_ATTR(8);
#undef AMP_COLLECT_ATTR
        if (ea != cudaSuccess) return bail(cuda_fail(ea, "cudaFuncSetAttribute(collect_reference_kernel)"));
        const char *e = getenv("AMP_B200_COLLECT_TABLE");  // read once per handle; amp_lib_set_option overrides it
        lib->collect_table = !e ? 0 : (e[0] == 'g' ? 1 : (e[0] == 's' ? 2 : 0));
    }
    if (!with_env) {
        // plain handle (MotionLoader.sample): stage the lerp row table for the row-table sample kernel
        v.lerp_width = 2 * v.num_dofs + 9 * v.num_bodies;
        v.lerp_stride = (v.lerp_width + 3) & ~3;
        if (v.lerp_width <= 32 * 16 && v.num_frames * (int64_t)v.lerp_stride < ((int64_t)1 << 40)) {
            void *rows = nullptr;
            // + 2 KiB: a row is read in 32-float slots without a bounds check
            cudaError_t e2 = cudaMalloc(&rows, (size_t)v.num_frames * v.lerp_stride * sizeof(float) + 2048);
            if (e2 != cudaSuccess) return bail(cuda_fail(e2, "cudaMalloc(lerp rows)"));
            lib->owned[lib->n_owned++] = rows;
            const int64_t total = v.num_frames * (int64_t)v.lerp_stride;
            pack_lerp_rows_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(v, static_cast<float *>(rows));
            e2 = cudaGetLastError();
            if (e2 != cudaSuccess) return bail(cuda_fail(e2, "pack_lerp_rows_kernel launch"));
            v.lerp_rows = static_cast<float *>(rows);
        }
    }
    // host arrays of the descriptor are read by the async copies above: finish them before returning
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaStreamSynchronize(amp_lib_create)"));
    *out = lib;
    return AMP_OK;
}

int amp_lib_set_option(amp_lib_t *lib, int32_t option, int64_t value) {
    AMP_REQUIRE(lib, "amp_lib_set_option: NULL handle");
    switch (option) {
        case AMP_OPT_COLLECT_TABLE:
            AMP_REQUIRE(value >= 0 && value <= 2, "amp_lib_set_option: AMP_OPT_COLLECT_TABLE takes 0 (auto), 1 (global) or 2 (shared)");
            lib->collect_table = (int)value;
            return AMP_OK;
        default:
            return fail(AMP_EINVAL, "amp_lib_set_option: unknown option %d", option);
    }
}

int amp_lib_destroy(amp_lib_t *lib) {
    if (!lib) return AMP_OK;
    for (int i = 0; i < lib->n_owned; ++i) cudaFree(lib->owned[i]);
    delete lib;
    return AMP_OK;
}

int amp_lib_obs_width(const amp_lib_t *lib) { return lib ? lib->v.obs_width : 0; }

int amp_lib_poll_flags(amp_lib_t *lib, void *stream, uint32_t *flags) {
    AMP_REQUIRE(lib && flags, "amp_lib_poll_flags: NULL argument");
    cudaStream_t st = as_stream(stream);
    AMP_CUDA_TRY(cudaMemcpyAsync(flags, lib->v.flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    AMP_CUDA_TRY(cudaMemsetAsync(lib->v.flags, 0, sizeof(uint32_t), st));
    AMP_CUDA_TRY(cudaStreamSynchronize(st));
    return AMP_OK;
}

int amp_frame_blend(amp_lib_t *lib, const double *times, const int64_t *ids, int64_t S, int64_t *idx0, int64_t *idx1,
                    float *blend32, double *blend64, void *stream) {
    AMP_REQUIRE(lib && S >= 0, "amp_frame_blend: bad handle or negative size");
    if (S == 0) return AMP_OK;
    AMP_REQUIRE(times && idx0 && idx1, "amp_frame_blend: NULL buffer");
    frame_blend_kernel<<<grid_for(S, 256, 8), 256, 0, as_stream(stream)>>>(lib->v, times, ids, S, idx0, idx1, blend32,
                                                                           blend64);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_sample_full(amp_lib_t *lib, const double *times, const int64_t *ids, int64_t S, float *dof_pos, float *dof_vel,
                    float *body_pos, float *body_rot, float *body_lin, float *body_ang, void *stream) {
    AMP_REQUIRE(lib && S >= 0, "amp_sample_full: bad handle or negative size");
    if (S == 0) return AMP_OK;
    AMP_REQUIRE(times, "amp_sample_full: NULL times");
    AMP_REQUIRE(!body_rot || aligned16(body_rot), "amp_sample_full: body_rot output must be 16-byte aligned");
    AMP_REQUIRE(lib->v.num_frames < ((int64_t)1 << 31), "amp_sample_full: more than 2^31 frames");
    AMP_REQUIRE(lib->v.num_bodies * 3 < 4096 && lib->v.num_dofs < 4096, "amp_sample_full: row wider than 4096 floats");
    auto magic = [](int w) { return (uint32_t)((((uint64_t)1 << 32) + (uint64_t)w - 1) / (uint64_t)w); };
    SampleDivisors dv{magic(lib->v.num_dofs), magic(lib->v.num_bodies * 3), magic(lib->v.num_bodies)};
    const int grid = grid_for((S + 31) / 32, kSampleWarps, 8);
    cudaStream_t st = as_stream(stream);
    const bool all_outputs = dof_pos && dof_vel && body_pos && body_rot && body_lin && body_ang;
    if (lib->v.lerp_rows && all_outputs) {
#define AMP_LAUNCH_ROWS(NS)                                                                                             \
    case NS:                                                                                                            \
        sample_rows_kernel<NS><<<grid, kSampleWarps * 32, 0, st>>>(lib->v, dv.body, times, ids, S, dof_pos, dof_vel, body_pos, \
                                                                   body_rot, body_lin, body_ang);                      \
        break
        switch ((lib->v.lerp_width + 31) / 32) {
            AMP_LAUNCH_ROWS(1); AMP_LAUNCH_ROWS(2); AMP_LAUNCH_ROWS(3); AMP_LAUNCH_ROWS(4); AMP_LAUNCH_ROWS(5); AMP_LAUNCH_ROWS(6);
            AMP_LAUNCH_ROWS(7); AMP_LAUNCH_ROWS(8); AMP_LAUNCH_ROWS(9); AMP_LAUNCH_ROWS(10); AMP_LAUNCH_ROWS(11); AMP_LAUNCH_ROWS(12);
            AMP_LAUNCH_ROWS(13); AMP_LAUNCH_ROWS(14); AMP_LAUNCH_ROWS(15); AMP_LAUNCH_ROWS(16);
            default: return fail(AMP_EINVAL, "amp_sample_full: internal: lerp row width %d", lib->v.lerp_width);
        }
#undef AMP_LAUNCH_ROWS
    } else {
        sample_full_kernel<<<grid, kSampleWarps * 32, 0, st>>>(lib->v, dv, times, ids, S, dof_pos, dof_vel, body_pos, body_rot,
                                                               body_lin, body_ang);
    }
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_lerp(const float *a, const float *b, const float *blend, int64_t n, int64_t inner, float *out, void *stream) {
    AMP_REQUIRE(n >= 0 && inner >= 1, "amp_lerp: bad sizes");
    if (n == 0) return AMP_OK;
    AMP_REQUIRE(a && b && blend && out, "amp_lerp: NULL buffer");
    lerp_kernel<<<grid_for(n * inner, 256, 8), 256, 0, as_stream(stream)>>>(a, b, blend, n * inner, inner, out);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_slerp(const float *q0, const float *q1, const float *blend, int64_t n, int64_t bodies, float *out, void *stream) {
    AMP_REQUIRE(n >= 0 && bodies >= 1, "amp_slerp: bad sizes");
    if (n == 0) return AMP_OK;
    AMP_REQUIRE(q0 && q1 && blend && out, "amp_slerp: NULL buffer");
    AMP_REQUIRE(aligned16(q0) && aligned16(q1) && aligned16(out), "amp_slerp: quaternion buffers must be 16-byte aligned");
    slerp_kernel<<<grid_for(n * bodies, 256, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4 *>(q0), reinterpret_cast<const float4 *>(q1), blend, n * bodies, bodies,
        reinterpret_cast<float4 *>(out));
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_collect_reference(amp_lib_t *lib, const double *cur_times, const int64_t *ids, int64_t n, int32_t K, float *out,
                          int64_t row_stride, int64_t capacity_rows, int64_t start_row, const int64_t *row_index,
                          void *stream) {
    AMP_REQUIRE(lib && n >= 0, "amp_collect_reference: bad handle or negative size");
    AMP_REQUIRE(lib->v.obs_width > 0, "amp_collect_reference: library was created without the env selection");
    AMP_REQUIRE(K >= 1 && K <= kMaxHistory, "amp_collect_reference: num_amp_observations %d outside [1, %d]", K,
                kMaxHistory);
    if (n == 0) return AMP_OK;
    const int A = lib->v.obs_width;
    AMP_REQUIRE(cur_times && out, "amp_collect_reference: NULL buffer");
    AMP_REQUIRE(row_stride >= (int64_t)K * A, "amp_collect_reference: row_stride %lld < K*A = %d", (long long)row_stride,
                K * A);
    AMP_REQUIRE(start_row >= 0, "amp_collect_reference: negative start_row");

    // tile = consecutive samples of one warp (one phase-1 pass of 32 frames; a whole sample when K > 32)
    const int tile_cap = K <= kTileFrames ? kTileFrames : kMaxHistory;
    const int max_tile = std::max(1, tile_cap / K);
    const int nslot = (A + 31) / 32;
    const int sms = sm_count();
    cudaStream_t st = as_stream(stream);

    // shared-memory-table variant: table + >= 8 warps of frame records must fit in 227 KiB, and every warp of the
    // persistent grid should get at least two full tiles (otherwise copying the table per CTA is not worth it)
    const size_t table_bytes = (size_t)lib->v.num_frames * lib->v.row_floats * sizeof(float) + 512;
    const size_t per_warp = (size_t)tile_cap * sizeof(FrameMeta);
    int smem_warps = table_bytes + 8 * per_warp <= (size_t)kMaxSmemOptin ? (int)std::min<size_t>(32, (kMaxSmemOptin - table_bytes) / per_warp) : 0;
    const int64_t full_tiles = (n + max_tile - 1) / max_tile;
    const int force_variant = lib->collect_table;  // AMP_OPT_COLLECT_TABLE (amp_lib_set_option / AMP_B200_COLLECT_TABLE at create)
    bool use_smem = smem_warps >= 8 && full_tiles >= (int64_t)2 * sms * smem_warps;
    if (force_variant == 1) use_smem = false;
    if (force_variant == 2 && smem_warps >= 1) use_smem = true;

    int tile_samples, grid, threads;
    int64_t tiles;
    size_t smem;
    if (use_smem) {
        tile_samples = max_tile;
        tiles = full_tiles;
        threads = smem_warps * 32;
        grid = (int)std::min<int64_t>(sms, (tiles + smem_warps - 1) / smem_warps);
        smem = table_bytes + (size_t)smem_warps * per_warp;
    } else {
        // shrink the tile when the batch is small so the work still covers the chip
        const int64_t target_warps = (int64_t)sms * kCollectWarps * 4;
        tile_samples = (int)std::min<int64_t>(max_tile, std::max<int64_t>(1, (n + target_warps - 1) / target_warps));
        tiles = (n + tile_samples - 1) / tile_samples;
        threads = kCollectWarps * 32;
        grid = grid_for(tiles, kCollectWarps, 4);
        smem = (size_t)kCollectWarps * per_warp;
    }
#define AMP_LAUNCH_COLLECT(NS)                                                                                          \
    do {                                                                                                                \
        if (use_smem) {                                                                                                 \
            collect_reference_kernel<NS, true><<<grid, threads, smem, st>>>(lib->v, cur_times, ids, n, K, tile_samples, \
                                                                            tile_cap, out, row_stride, capacity_rows,   \
                                                                            start_row, row_index, tiles);               \
        } else {                                                                                                        \
            collect_reference_kernel<NS, false><<<grid, threads, smem, st>>>(lib->v, cur_times, ids, n, K,              \
                                                                             tile_samples, tile_cap, out, row_stride,   \
                                                                             capacity_rows, start_row, row_index,       \
                                                                             tiles);                                    \
        }                                                                                                               \
    } while (0)
#ifdef AMP_COLLECT_PROFILE
    {
        const char *e = getenv("AMP_COLLECT_MODE");
        const int mode = e ? atoi(e) : 0;
        AMP_CUDA_TRY(cudaMemcpyToSymbolAsync(c_collect_mode, &mode, sizeof(int), 0, cudaMemcpyHostToDevice, st));
    }
#endif
    switch (nslot) {
        case 1: AMP_LAUNCH_COLLECT(1); break;
        case 2: AMP_LAUNCH_COLLECT(2); break;
        case 3: AMP_LAUNCH_COLLECT(3); break;
        case 4: AMP_LAUNCH_COLLECT(4); break;
        case 5: AMP_LAUNCH_COLLECT(5); break;
        case 6: AMP_LAUNCH_COLLECT(6); break;
        case 7: AMP_LAUNCH_COLLECT(7); break;
        case 8: AMP_LAUNCH_COLLECT(8); break;
        default: return fail(AMP_EINVAL, "amp_collect_reference: observation width %d > 256 is not supported", A);
    }
#undef AMP_LAUNCH_COLLECT
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_compute_obs(const float *dof_pos, const float *dof_vel, const float *root_pos, const float *root_rot,
                    const float *root_lin, const float *root_ang, const float *key_pos, int64_t n, int32_t D, int32_t Kb,
                    float *out, void *stream) {
    AMP_REQUIRE(n >= 0 && D >= 0 && Kb >= 0, "amp_compute_obs: bad sizes");
    if (n == 0) return AMP_OK;
    AMP_REQUIRE((D == 0 || (dof_pos && dof_vel)) && root_pos && root_rot && root_lin && root_ang && (Kb == 0 || key_pos) && out,
                "amp_compute_obs: NULL buffer");
    const int A = 2 * D + 13 + 3 * Kb;
    const int grid = grid_for(n, 8, 8);
    cudaStream_t st = as_stream(stream);
#define AMP_LAUNCH_OBS(NS)                                                                                              \
    case NS:                                                                                                            \
        compute_obs_kernel<NS><<<grid, 256, 0, st>>>(dof_pos, dof_vel, root_pos, root_rot, root_lin, root_ang, key_pos, n, D, \
                                                     Kb, out);                                                          \
        break
    switch ((A + 31) / 32) {
        AMP_LAUNCH_OBS(1); AMP_LAUNCH_OBS(2); AMP_LAUNCH_OBS(3); AMP_LAUNCH_OBS(4); AMP_LAUNCH_OBS(5); AMP_LAUNCH_OBS(6);
        AMP_LAUNCH_OBS(7); AMP_LAUNCH_OBS(8);
        default: return fail(AMP_EINVAL, "amp_compute_obs: observation width %d > 256 is not supported", A);
    }
#undef AMP_LAUNCH_OBS
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_tangent_normal(const float *q, int64_t n, float *out, void *stream) {
    AMP_REQUIRE(n >= 0, "amp_tangent_normal: negative size");
    if (n == 0) return AMP_OK;
    AMP_REQUIRE(q && out, "amp_tangent_normal: NULL buffer");
    tangent_normal_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(q, n, out);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_obs_step(const float *joint_pos, const float *joint_vel, const float *body_pos_w, const float *body_quat_w,
                 const float *body_lin_vel_w, const float *body_ang_vel_w, int64_t N, int32_t D, int32_t Bsim,
                 int32_t ref_body, const int32_t *key_bodies, int32_t Kb, int32_t K, float *amp_buf, float *policy_obs,
                 int64_t policy_stride, void *stream) {
    AMP_REQUIRE(N >= 0 && D >= 1 && Bsim >= 1 && K >= 1, "amp_obs_step: bad sizes");
    AMP_REQUIRE(Kb >= 0 && Kb <= kMaxKeyBodies && (Kb == 0 || key_bodies), "amp_obs_step: bad key body list");
    AMP_REQUIRE(ref_body >= 0 && ref_body < Bsim, "amp_obs_step: ref body out of range");
    if (N == 0) return AMP_OK;
    AMP_REQUIRE(joint_pos && joint_vel && body_pos_w && body_quat_w && body_lin_vel_w && body_ang_vel_w && amp_buf,
                "amp_obs_step: NULL buffer");
    AMP_REQUIRE(aligned16(body_quat_w), "amp_obs_step: body_quat_w must be 16-byte aligned");
    KeyList keys{};
    for (int i = 0; i < Kb; ++i) {
        AMP_REQUIRE(key_bodies[i] >= 0 && key_bodies[i] < Bsim, "amp_obs_step: key body %d out of range", key_bodies[i]);
        keys.k[i] = key_bodies[i];
    }
    const int A = 2 * D + 13 + 3 * Kb;
    AMP_REQUIRE(!policy_obs || policy_stride >= A - 3 * Kb, "amp_obs_step: policy_stride too small");
    const int grid = grid_for(N, 8, 8);
    cudaStream_t st = as_stream(stream);
#define AMP_LAUNCH_STEP(NS)                                                                                              \
    obs_step_kernel<NS><<<grid, 256, 0, st>>>(joint_pos, joint_vel, body_pos_w, body_quat_w, body_lin_vel_w, body_ang_vel_w, \
                                              N, D, Bsim, ref_body, keys, Kb, K, amp_buf, policy_obs, policy_stride)
    switch ((A + 31) / 32) {
        case 1: AMP_LAUNCH_STEP(1); break;
        case 2: AMP_LAUNCH_STEP(2); break;
        case 3: AMP_LAUNCH_STEP(3); break;
        case 4: AMP_LAUNCH_STEP(4); break;
        case 5: AMP_LAUNCH_STEP(5); break;
        case 6: AMP_LAUNCH_STEP(6); break;
        case 7: AMP_LAUNCH_STEP(7); break;
        case 8: AMP_LAUNCH_STEP(8); break;
        default: return fail(AMP_EINVAL, "amp_obs_step: observation width %d > 256 is not supported", A);
    }
#undef AMP_LAUNCH_STEP
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_actor_obs_step(const float *amp_buf, int64_t N, int32_t K, int32_t A, int32_t base_width, const float *last_actions,
                       int32_t act, const float *command, int32_t cmd, int32_t num_actor_observations,
                       int32_t hist_include_actions, int32_t hist_include_command, float *hist_buf, uint8_t *just_reset,
                       float *actor_obs, int64_t actor_stride, void *stream) {
    AMP_REQUIRE(N >= 0 && K >= 1 && A >= 1 && base_width >= 1 && base_width <= A && act >= 0 && cmd >= 0,
                "amp_actor_obs_step: bad sizes");
    AMP_REQUIRE(num_actor_observations >= 1, "amp_actor_obs_step: num_actor_observations must be >= 1");
    if (N == 0) return AMP_OK;
    AMP_REQUIRE(amp_buf && actor_obs && (act == 0 || last_actions) && (cmd == 0 || command), "amp_actor_obs_step: NULL buffer");
    const int n_hist = num_actor_observations - 1;
    AMP_REQUIRE(n_hist == 0 || hist_buf, "amp_actor_obs_step: history buffer required when num_actor_observations > 1");
    const int P = base_width + (hist_include_actions ? act : 0) + (hist_include_command ? cmd : 0);
    AMP_REQUIRE(actor_stride >= base_width + act + cmd + (int64_t)n_hist * P, "amp_actor_obs_step: actor_stride too small");
    actor_obs_kernel<<<grid_for(N, 8, 8), 256, 0, as_stream(stream)>>>(amp_buf, N, K, A, base_width, last_actions, act, command, cmd,
                                                                       n_hist, hist_include_actions, hist_include_command, hist_buf,
                                                                       just_reset, actor_obs, actor_stride);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

static int make_reward_scales(const float *scales, RewardScales *out) {
    RewardScales sc{};
    sc.termination = scales[0];
    sc.action_l2 = scales[1];
    sc.joint_pos_limits = scales[2];
    sc.joint_acc_l2 = scales[3];
    sc.joint_vel_l2 = scales[4];
    sc.track_vel = scales[5];
    if (sc.track_vel > 0.0f) {
        const double w = (double)scales[5], sigma_sq = 0.25, floor_v = 4.0;  // the scripted reference evaluates these in double
        sc.exp_at_floor = (float)(w * std::exp(-floor_v));
        sc.linear_slope = (float)(w / sigma_sq * std::exp(-floor_v));
    }
    *out = sc;
    return AMP_OK;
}

int amp_env_step(const amp_env_step_t *a, void *stream) {
    AMP_REQUIRE(a, "amp_env_step: NULL argument block");
    const int64_t N = a->num_envs;
    const int D = a->num_dofs, Bsim = a->num_sim_bodies, Kb = a->num_key_bodies, K = a->num_amp_observations;
    AMP_REQUIRE(N >= 0 && D >= 1 && Bsim >= 1 && K >= 1, "amp_env_step: bad sizes");
    AMP_REQUIRE(Kb >= 0 && Kb <= kMaxKeyBodies && (Kb == 0 || a->key_bodies), "amp_env_step: bad key body list");
    AMP_REQUIRE(a->ref_body >= 0 && a->ref_body < Bsim, "amp_env_step: ref body out of range");
    if (N == 0) return AMP_OK;
    AMP_REQUIRE(a->joint_pos && a->joint_vel && a->body_pos_w && a->body_quat_w && a->body_lin_vel_w && a->body_ang_vel_w && a->amp_buf,
                "amp_env_step: NULL buffer");
    AMP_REQUIRE(aligned16(a->body_quat_w), "amp_env_step: body_quat_w must be 16-byte aligned");
    EnvStepParams p{};
    for (int i = 0; i < Kb; ++i) {
        AMP_REQUIRE(a->key_bodies[i] >= 0 && a->key_bodies[i] < Bsim, "amp_env_step: key body %d out of range", a->key_bodies[i]);
        p.keys.k[i] = a->key_bodies[i];
    }
    const int A = 2 * D + 13 + 3 * Kb, base = A - 3 * Kb;
    p.joint_pos = a->joint_pos; p.joint_vel = a->joint_vel; p.body_pos = a->body_pos_w; p.body_quat = a->body_quat_w;
    p.body_lin = a->body_lin_vel_w; p.body_ang = a->body_ang_vel_w;
    p.N = N; p.D = D; p.Bsim = Bsim; p.ref = a->ref_body; p.Kb = Kb; p.K = K;
    p.amp_buf = a->amp_buf;
    p.act = a->action_size; p.cmd = a->command_size;
    AMP_REQUIRE(p.act >= 0 && p.cmd >= 0, "amp_env_step: negative action / command size");
    if (a->actor_obs) {
        AMP_REQUIRE(a->num_actor_observations >= 1, "amp_env_step: num_actor_observations must be >= 1");
        AMP_REQUIRE((p.act == 0 || a->last_actions) && (p.cmd == 0 || a->command), "amp_env_step: NULL last_actions / command");
        p.n_hist = a->num_actor_observations - 1;
        p.inc_act = a->hist_include_actions ? 1 : 0;
        p.inc_cmd = a->hist_include_command ? 1 : 0;
        AMP_REQUIRE(p.n_hist == 0 || a->hist_buf, "amp_env_step: history buffer required when num_actor_observations > 1");
        const int P = base + (p.inc_act ? p.act : 0) + (p.inc_cmd ? p.cmd : 0);
        AMP_REQUIRE(a->actor_stride >= base + p.act + p.cmd + (int64_t)p.n_hist * P, "amp_env_step: actor_stride too small");
        p.last_actions = a->last_actions; p.command = a->command; p.hist_buf = a->hist_buf; p.just_reset = a->just_reset;
        p.actor_obs = a->actor_obs; p.actor_stride = a->actor_stride;
    }
    if (a->reward_total) {
        AMP_REQUIRE(a->reward_scales && a->reset_terminated && (p.act == 0 || a->actions) && a->soft_limits && a->joint_acc,
                    "amp_env_step: NULL reward input");
        make_reward_scales(a->reward_scales, &p.sc);
        AMP_REQUIRE(p.sc.track_vel <= 0.0f || (a->command && p.cmd == 2), "amp_env_step: velocity tracking needs the (N, 2) command");
        AMP_REQUIRE((reinterpret_cast<uintptr_t>(a->soft_limits) & 7u) == 0, "amp_env_step: soft_limits must be 8-byte aligned");
        p.terminated = a->reset_terminated; p.actions = a->actions; p.limits = a->soft_limits; p.joint_acc = a->joint_acc;
        p.command = a->command;
        p.total = a->reward_total; p.terms = a->reward_terms; p.track_err = a->track_err;
    }
    const int grid = grid_for(N, 8, 8);
    cudaStream_t st = as_stream(stream);
#define AMP_LAUNCH_ENV_STEP(NS) case NS: env_step_kernel<NS><<<grid, 256, 0, st>>>(p); break
    switch ((A + 31) / 32) {
        AMP_LAUNCH_ENV_STEP(1); AMP_LAUNCH_ENV_STEP(2); AMP_LAUNCH_ENV_STEP(3); AMP_LAUNCH_ENV_STEP(4);
        AMP_LAUNCH_ENV_STEP(5); AMP_LAUNCH_ENV_STEP(6); AMP_LAUNCH_ENV_STEP(7); AMP_LAUNCH_ENV_STEP(8);
        default: return fail(AMP_EINVAL, "amp_env_step: observation width %d > 256 is not supported", A);
    }
#undef AMP_LAUNCH_ENV_STEP
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_task_reward(const float *scales, const uint8_t *reset_terminated, const float *actions, int32_t act, const float *joint_pos,
                    const float *soft_limits, const float *joint_acc, const float *joint_vel, int32_t D, const float *body_lin_vel_w,
                    const float *body_quat_w, int32_t Bsim, int32_t ref_body, const float *command, int64_t N, float *total,
                    float *terms, float *track_err, void *stream) {
    AMP_REQUIRE(scales && N >= 0 && act >= 0 && D >= 0, "amp_task_reward: bad arguments");
    if (N == 0) return AMP_OK;
    AMP_REQUIRE(reset_terminated && total && (act == 0 || actions) && (D == 0 || (joint_pos && soft_limits && joint_acc && joint_vel)),
                "amp_task_reward: NULL buffer");
    RewardScales sc{};
    make_reward_scales(scales, &sc);
    if (sc.track_vel > 0.0f) {
        AMP_REQUIRE(body_lin_vel_w && body_quat_w && command && Bsim >= 1 && ref_body >= 0 && ref_body < Bsim,
                    "amp_task_reward: velocity tracking needs body_lin_vel_w, body_quat_w, command and a valid reference body");
        AMP_REQUIRE(aligned16(body_quat_w), "amp_task_reward: body_quat_w must be 16-byte aligned");
    }
    AMP_REQUIRE(D == 0 || (reinterpret_cast<uintptr_t>(soft_limits) & 7u) == 0, "amp_task_reward: soft_limits must be 8-byte aligned");
    task_reward_kernel<<<grid_for(N, 8, 8), 256, 0, as_stream(stream)>>>(sc, reset_terminated, actions, act, joint_pos, soft_limits,
                                                                         joint_acc, joint_vel, D, body_lin_vel_w, body_quat_w, Bsim,
                                                                         ref_body, command, N, total, terms, track_err);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

}  // extern "C"
