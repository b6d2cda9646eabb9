// AMP memories and the state preprocessor's statistics (the rows either side of the style
// reward): skrl RandomMemory.sample_by_index gathers and RunningStandardScaler (train = True update, eval apply).
// All three are HBM-bound row streams.
#include <algorithm>
#include <cmath>

#include "amp_internal.h"
#include "amp_math.cuh"

namespace amp {

// ------------------------------------------------------------------------------------------------------------------
// out[r, :] = src[row_index[r], :]   (skrl memories/torch/base.py sample_by_index: tensors_view[name][indexes])
// One warp per destination row, VEC floats per lane and load; ROWS rows per trip so that ROWS * ceil(W / (32 VEC)) loads
// are in flight per lane before the first store (a lone 664-byte row per warp left HBM latency exposed).
// ------------------------------------------------------------------------------------------------------------------
template <int VEC>
struct VecT;
template <>
struct VecT<1> { using type = float; };
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<4> { using type = float4; };

template <int VEC, int ROWS, int NSLOT>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ src, int64_t src_stride, int64_t capacity,
                                                           const int64_t *__restrict__ row_index, int64_t M, int Wv,
                                                           float *__restrict__ out, int64_t out_stride,
                                                           uint32_t *__restrict__ flags) {
    using V = typename VecT<VEC>::type;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * ROWS; r0 < M; r0 += nwarps * ROWS) {
        int64_t idx[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const int64_t r = r0 + j;
            int64_t i = r < M ? __ldg(row_index + r) : 0;
            if (i < 0 || i >= capacity) {  // skrl would raise an IndexError: flag it, emit a zero row
                if (r < M && lane == 0 && flags) atomicOr(flags, 2u);
                i = -1;
            }
            idx[j] = i;
        }
        for (int c0 = 0; c0 < Wv; c0 += 32 * NSLOT) {  // NSLOT * 32 vectors of every row per trip (one trip up to 512 floats)
            V v[ROWS][NSLOT];
#pragma unroll
            for (int j = 0; j < ROWS; ++j)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    const int c = c0 + lane + 32 * s;
                    V z{};
                    v[j][s] = (idx[j] >= 0 && c < Wv) ? __ldg(reinterpret_cast<const V *>(src + idx[j] * src_stride) + c) : z;
                }
#pragma unroll
            for (int j = 0; j < ROWS; ++j)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    const int c = c0 + lane + 32 * s;
                    if (r0 + j < M && c < Wv) __stcs(reinterpret_cast<V *>(out + (r0 + j) * out_stride) + c, v[j][s]);
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// RunningStandardScaler statistics (skrl resources/preprocessors/torch/running_standard_scaler.py, _parallel_variance):
//   pass 1  per-CTA column sums and sums of squares in float64 (thread = column, rows strided over the grid, 8 loads
//           in flight per thread) -> scratch[cta][2][W]
//   pass 2  fixed-order reduction over the CTAs (deterministic), batch mean / unbiased variance rounded to fp32 (they are
//           fp32 tensors in the reference), then the float64 parallel-variance merge into the running buffers, in place.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kStatThreads = 256;
constexpr int kStatUnroll = 8;

__global__ void __launch_bounds__(kStatThreads) scaler_partial_kernel(const float *__restrict__ x, int64_t x_stride, int64_t M,
                                                                       int W, double *__restrict__ scratch) {
    const int c = blockIdx.y * kStatThreads + threadIdx.x;
    double s = 0.0, ss = 0.0;
    if (c < W) {
        const float *col = x + c;
        int64_t r = blockIdx.x;
        const int64_t step = gridDim.x;
        for (; r + (kStatUnroll - 1) * step < M; r += kStatUnroll * step) {
            float v[kStatUnroll];
#pragma unroll
            for (int u = 0; u < kStatUnroll; ++u) v[u] = __ldcs(col + (r + u * step) * x_stride);
#pragma unroll
            for (int u = 0; u < kStatUnroll; ++u) {
                const double d = (double)v[u];
                s += d;
                ss = fma(d, d, ss);
            }
        }
        for (; r < M; r += step) {
            const double d = (double)__ldcs(col + r * x_stride);
            s += d;
            ss = fma(d, d, ss);
        }
        double *slot = scratch + (size_t)blockIdx.x * 2 * W;
        slot[c] = s;
        slot[W + c] = ss;
    }
}

// 32 columns per block; the 32 warps split the partial sums (warp w takes parts w, w+32, ...: at most a handful each, all
// loads in flight together) and meet in shared memory in a fixed order, so the result does not depend on scheduling.
constexpr int kMergeWarps = 32;
__global__ void __launch_bounds__(32 * kMergeWarps) scaler_merge_kernel(const double *__restrict__ scratch, int parts, int64_t M, int W,
                                                                        double *__restrict__ running_mean, double *__restrict__ running_var,
                                                                        const double *__restrict__ count_in) {
    __shared__ double sh[2][kMergeWarps][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double s = 0.0, ss = 0.0;
    if (c < W) {
#pragma unroll 8
        for (int p = ty; p < parts; p += kMergeWarps) {
            s += scratch[(size_t)p * 2 * W + c];
            ss += scratch[(size_t)p * 2 * W + W + c];
        }
    }
    sh[0][ty][tx] = s;
    sh[1][ty][tx] = ss;
    __syncthreads();
    if (ty != 0 || c >= W) return;
    s = ss = 0.0;
#pragma unroll
    for (int w = 0; w < kMergeWarps; ++w) {
        s += sh[0][w][tx];
        ss += sh[1][w][tx];
    }
    const double n = (double)M;
    // torch.mean / torch.var(unbiased) of an fp32 batch are fp32 values; M == 1 gives NaN like torch.var
    const double mean_b = (double)(float)(s / n);
    const double var_b = (double)(float)((ss - s * s / n) / (n - 1.0));
    const double cur = *count_in;
    const double total = cur + n;
    const double rm = running_mean[c];
    const double delta = mean_b - rm;
    const double m2 = running_var[c] * cur + var_b * n + delta * delta * cur * n / total;
    running_mean[c] = rm + delta * n / total;
    running_var[c] = m2 / total;
}

__global__ void scaler_count_kernel(double *count, int64_t M) { *count += (double)M; }

// eval: clamp((x - mean.float()) / (sqrt(var.float()) + eps), -clip, clip), every operation rounded like torch's fp32 ops.
// The fp32 statistics are staged once per CTA in shared memory; one warp per row, ROWS rows in flight per trip.
template <int VEC>
__global__ void __launch_bounds__(256) scaler_apply_kernel(const float *__restrict__ x, int64_t x_stride, int64_t M, int W,
                                                            const double *__restrict__ mean, const double *__restrict__ var,
                                                            float eps, float clip, float *__restrict__ out, int64_t out_stride) {
    using V = typename VecT<VEC>::type;
    extern __shared__ __align__(16) float stat_smem[];  // [W] mean, then [W] denominators (W rounded up to 4)
    const int Wp = (W + 3) & ~3;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        stat_smem[c] = __double2float_rn(mean[c]);
        stat_smem[Wp + c] = __fadd_rn(__fsqrt_rn(__double2float_rn(var[c])), eps);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int Wv = W / VEC;
    constexpr int ROWS = 2;
    for (int64_t r0 = warp * ROWS; r0 < M; r0 += nwarps * ROWS) {
        for (int c = lane; c < Wv; c += 32) {
            V v[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j)
                if (r0 + j < M) v[j] = __ldcs(reinterpret_cast<const V *>(x + (r0 + j) * x_stride) + c);
            const V m = reinterpret_cast<const V *>(stat_smem)[c], d = reinterpret_cast<const V *>(stat_smem + Wp)[c];
            const float *mf = reinterpret_cast<const float *>(&m), *df = reinterpret_cast<const float *>(&d);
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                if (r0 + j >= M) continue;
                float *f = reinterpret_cast<float *>(&v[j]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) f[e] = clamp_nan(__fdiv_rn(__fsub_rn(f[e], mf[e]), df[e]), -clip, clip);
                __stcs(reinterpret_cast<V *>(out + (r0 + j) * out_stride) + c, v[j]);
            }
        }
    }
}

static int pick_vec(const void *a, int64_t sa, const void *b, int64_t sb, int W) {
    auto ok = [&](int v) {
        return W % v == 0 && sa % v == 0 && sb % v == 0 && (reinterpret_cast<uintptr_t>(a) % (4 * v)) == 0 &&
               (reinterpret_cast<uintptr_t>(b) % (4 * v)) == 0;
    };
    return ok(4) ? 4 : (ok(2) ? 2 : 1);
}

}  // namespace amp

using namespace amp;

extern "C" {

int amp_gather_rows(const float *src, int64_t src_stride, int64_t capacity, const int64_t *row_index, int64_t M, int32_t W,
                    float *out, int64_t out_stride, uint32_t *flags, void *stream) {
    AMP_REQUIRE(M >= 0 && W >= 1 && capacity >= 0, "amp_gather_rows: bad sizes");
    if (M == 0) return AMP_OK;
    AMP_REQUIRE(src && row_index && out, "amp_gather_rows: NULL buffer");
    AMP_REQUIRE(src_stride >= W && out_stride >= W, "amp_gather_rows: row strides must be >= W");
    cudaStream_t st = as_stream(stream);
    const int vec = pick_vec(src, src_stride, out, out_stride, W);
    const int Wv = W / vec;
    constexpr int ROWS = 4;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((M + 8 * ROWS - 1) / (8 * ROWS), (int64_t)sm_count() * 8));
    if (vec == 4) gather_rows_kernel<4, ROWS, 4><<<grid, 256, 0, st>>>(src, src_stride, capacity, row_index, M, Wv, out, out_stride, flags);
    else if (vec == 2) gather_rows_kernel<2, ROWS, 4><<<grid, 256, 0, st>>>(src, src_stride, capacity, row_index, M, Wv, out, out_stride, flags);
    else gather_rows_kernel<1, ROWS, 4><<<grid, 256, 0, st>>>(src, src_stride, capacity, row_index, M, Wv, out, out_stride, flags);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

static int stat_parts(int64_t M) { return (int)std::max<int64_t>(1, std::min<int64_t>((M + 15) / 16, (int64_t)sm_count() * 4)); }

int64_t amp_scaler_scratch_bytes(int32_t W) { return W < 1 ? 0 : (int64_t)sm_count() * 4 * 2 * W * (int64_t)sizeof(double); }

int amp_scaler_update(const float *x, int64_t x_stride, int64_t M, int32_t W, double *running_mean, double *running_variance,
                      double *current_count, void *scratch, int64_t scratch_bytes, void *stream) {
    AMP_REQUIRE(M >= 0 && W >= 1, "amp_scaler_update: bad sizes");
    if (M == 0) return AMP_OK;
    AMP_REQUIRE(x && running_mean && running_variance && current_count && scratch, "amp_scaler_update: NULL buffer");
    AMP_REQUIRE(x_stride >= W, "amp_scaler_update: x_stride %lld < W %d", (long long)x_stride, W);
    const int parts = stat_parts(M);
    AMP_REQUIRE(scratch_bytes >= (int64_t)parts * 2 * W * (int64_t)sizeof(double),
                "amp_scaler_update: scratch too small (%lld bytes; amp_scaler_scratch_bytes(W) gives the size)", (long long)scratch_bytes);
    cudaStream_t st = as_stream(stream);
    const int col_blocks = (W + kStatThreads - 1) / kStatThreads;
    scaler_partial_kernel<<<dim3(parts, col_blocks), kStatThreads, 0, st>>>(x, x_stride, M, W, static_cast<double *>(scratch));
    scaler_merge_kernel<<<(W + 31) / 32, 32 * kMergeWarps, 0, st>>>(static_cast<const double *>(scratch), parts, M, W, running_mean,
                                                             running_variance, current_count);
    scaler_count_kernel<<<1, 1, 0, st>>>(current_count, M);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_scaler_apply(const float *x, int64_t x_stride, int64_t M, int32_t W, const double *running_mean,
                     const double *running_variance, float epsilon, float clip, float *out, int64_t out_stride, void *stream) {
    AMP_REQUIRE(M >= 0 && W >= 1, "amp_scaler_apply: bad sizes");
    if (M == 0) return AMP_OK;
    AMP_REQUIRE(x && running_mean && running_variance && out, "amp_scaler_apply: NULL buffer");
    AMP_REQUIRE(x_stride >= W && out_stride >= W, "amp_scaler_apply: row strides must be >= W");
    cudaStream_t st = as_stream(stream);
    const int vec = pick_vec(x, x_stride, out, out_stride, W);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((M + 15) / 16, (int64_t)sm_count() * 8));
    AMP_REQUIRE(W <= 4096, "amp_scaler_apply: W %d > 4096 is not supported", W);
    const size_t smem = (size_t)2 * ((W + 3) & ~3) * sizeof(float);
    if (vec == 4) scaler_apply_kernel<4><<<grid, 256, smem, st>>>(x, x_stride, M, W, running_mean, running_variance, epsilon, clip, out, out_stride);
    else if (vec == 2) scaler_apply_kernel<2><<<grid, 256, smem, st>>>(x, x_stride, M, W, running_mean, running_variance, epsilon, clip, out, out_stride);
    else scaler_apply_kernel<1><<<grid, 256, smem, st>>>(x, x_stride, M, W, running_mean, running_variance, epsilon, clip, out, out_stride);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

}  // extern "C"
