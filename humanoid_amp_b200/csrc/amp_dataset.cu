// Offline dataset pipeline: the reference's motions/data_convert.py:161-379 on the GPU.
//
//   CSV rows at 30 fps (root xyz, root quat xyzw, D joint angles; float32)
//     -> 2N-1 frames at 60 fps: scipy interp1d(kind="linear") of root position / joint angles, scipy Slerp of the root rotation
//     -> forward kinematics over the URDF tree (Pinocchio forwardKinematics + updateFramePlacements), Eigen matrix -> quaternion
//     -> central differences + gaussian_filter1d(sigma=1) for joint / body linear velocities
//     -> quaternion-log angular velocities of adjacent frames, averaged, gaussian_filter1d(sigma=1)
//
// Every frame is independent (one thread per output frame for interpolation + FK, one thread per output element for the
// velocity stages); all float64 steps of the reference are float64 here.  Per-INPUT-frame quantities of scipy's Slerp (the
// normalised key quaternions and the rotation vectors between neighbours) are prepared by slerp_keys_kernel.
// Compiled with -fmad=false: the float32 steps of the reference (numpy) round every operation.
#include <algorithm>
#include <cstdint>

#include "amp_internal.h"

namespace amp {
namespace dataset {

constexpr int kMaxJoints = 64;

struct Tree {
    int n_joints;
    const int32_t *parent;   // [J] joint whose child link is this joint's parent link, -1 = root link
    const int32_t *qidx;     // [J] column of the joint-angle array, -1 = fixed joint
    const double *origin_xyz;  // (J, 3)
    const double *origin_rot;  // (J, 9) row-major, from the URDF rpy
    const double *axis;        // (J, 3) unit
};

__device__ __forceinline__ void mat3_mul(const double *a, const double *b, double *c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

// scipy Rotation.from_quat / _normalize: q / sqrt(x^2 + y^2 + z^2 + w^2)
__device__ __forceinline__ void quat_normalise(double *q) {
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// scipy _compose_quat(p, q), both (x, y, z, w)
__device__ __forceinline__ void quat_compose(const double *p, const double *q, double *o) {
    const double cx = p[1] * q[2] - p[2] * q[1], cy = p[2] * q[0] - p[0] * q[2], cz = p[0] * q[1] - p[1] * q[0];
    o[0] = p[3] * q[0] + q[3] * p[0] + cx;
    o[1] = p[3] * q[1] + q[3] * p[1] + cy;
    o[2] = p[3] * q[2] + q[3] * p[2] + cz;
    o[3] = p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2];
}

// Per input frame i: key_quat[i] = normalised root quaternion (scipy Rotation.from_quat); for i < n_in - 1,
// rotvec[i] = (key[i].inv() * key[i+1]).as_rotvec()  (scipy Slerp.__init__)
__global__ void slerp_keys_kernel(const float *__restrict__ rows, int n_in, int n_cols, double *__restrict__ key_quat,
                                  double *__restrict__ rotvec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    double q0[4], q1[4];
    for (int k = 0; k < 4; ++k) q0[k] = (double)rows[(int64_t)i * n_cols + 3 + k];
    quat_normalise(q0);
    for (int k = 0; k < 4; ++k) key_quat[4 * (int64_t)i + k] = q0[k];
    if (i + 1 >= n_in) return;
    for (int k = 0; k < 4; ++k) q1[k] = (double)rows[(int64_t)(i + 1) * n_cols + 3 + k];
    quat_normalise(q1);
    const double inv[4] = {-q0[0], -q0[1], -q0[2], q0[3]};
    double r[4];
    quat_compose(inv, q1, r);
    quat_normalise(r);
    if (r[3] < 0.0) { r[0] = -r[0]; r[1] = -r[1]; r[2] = -r[2]; r[3] = -r[3]; }  // as_rotvec: w >= 0 (rotation angle <= pi)
    const double angle = 2.0 * atan2(sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]), r[3]);
    double scale;
    if (angle <= 1e-3) {
        const double a2 = angle * angle;
        scale = 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0;
    } else {
        scale = angle / sin(angle / 2.0);
    }
    for (int k = 0; k < 3; ++k) rotvec[3 * (int64_t)i + k] = scale * r[k];
}

struct FkArgs {
    const float *rows;  // (n_in, n_cols)
    int n_in, n_cols, n_out, n_dofs, n_bodies;
    const double *t_orig, *t_new;  // [n_in], [n_out]
    const int32_t *lerp_lo;        // [n_out] lower knot of interp1d
    const int32_t *slerp_ind;      // [n_out]
    const double *slerp_alpha;     // [n_out]
    const double *key_quat, *rotvec;
    Tree tree;
    const int32_t *body_joint;  // [B] joint whose child link is the recorded body, -1 = root link
    double *dof_positions;      // (n_out, D) f64
    float *body_positions;      // (n_out, B, 3)
    float *body_rotations;      // (n_out, B, 4) wxyz
    double *root_pose;          // (n_out, 7) xyz + xyzw, or NULL
};

// Eigen Quaternion(Matrix3) -> (w, x, y, z) float32 (data_convert.py:340-343)
__device__ __forceinline__ void store_quat_from_matrix(const double *m, float *out) {
    double t = m[0] + m[4] + m[8];
    double q[4];  // x, y, z, w
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t;
        q[1] = (m[2] - m[6]) * t;
        q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
        q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
    out[0] = (float)q[3]; out[1] = (float)q[0]; out[2] = (float)q[1]; out[3] = (float)q[2];
}

__global__ void __launch_bounds__(128) interp_fk_kernel(FkArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.n_out) return;
    // ---- scipy interp1d, kind="linear" (_call_linear of scipy >= 1.10; the reference pins no version, this is the one in the
    //      image): y = ((x - x_lo) / (x_hi - x_lo)) y_hi + ((x_hi - x) / (x_hi - x_lo)) y_lo, float64 weights times the float32
    //      samples widened exactly
    const int lo = a.lerp_lo[n], hi = lo + 1;
    const double x_lo = a.t_orig[lo], x_hi = a.t_orig[hi], x = a.t_new[n];
    const float *r_lo = a.rows + (int64_t)lo * a.n_cols, *r_hi = a.rows + (int64_t)hi * a.n_cols;
    const double w_hi = (x - x_lo) / (x_hi - x_lo), w_lo = (x_hi - x) / (x_hi - x_lo);
    auto lerp = [&](int col) { return w_hi * (double)r_hi[col] + w_lo * (double)r_lo[col]; };
    double root_p[3] = {lerp(0), lerp(1), lerp(2)};
    double *dof = a.dof_positions + (int64_t)n * a.n_dofs;
    for (int d = 0; d < a.n_dofs; ++d) dof[d] = lerp(7 + d);
    // ---- scipy Slerp.__call__: rotations[ind] * Rotation.from_rotvec(rotvecs[ind] * alpha)
    const int ind = a.slerp_ind[n];
    const double alpha = a.slerp_alpha[n];
    const double rv[3] = {a.rotvec[3 * (int64_t)ind] * alpha, a.rotvec[3 * (int64_t)ind + 1] * alpha, a.rotvec[3 * (int64_t)ind + 2] * alpha};
    const double angle = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    double scale;
    if (angle <= 1e-3) {
        const double a2 = angle * angle;
        scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
    } else {
        scale = sin(angle / 2.0) / angle;
    }
    double dq[4] = {scale * rv[0], scale * rv[1], scale * rv[2], cos(angle / 2.0)};
    double q[4];
    quat_compose(a.key_quat + 4 * (int64_t)ind, dq, q);
    quat_normalise(q);
    if (a.root_pose) {
        double *rp = a.root_pose + 7 * (int64_t)n;
        rp[0] = root_p[0]; rp[1] = root_p[1]; rp[2] = root_p[2];
        rp[3] = q[0]; rp[4] = q[1]; rp[5] = q[2]; rp[6] = q[3];
    }
    // ---- forward kinematics (Pinocchio: free-flyer normalises its quaternion; oMi[child] = oMi[parent] * (origin * R(axis, q)))
    quat_normalise(q);
    double R0[9];
    {
        const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
        const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1],
                     tyz = tz * q[1], tzz = tz * q[2];
        R0[0] = 1 - (tyy + tzz); R0[1] = txy - twz;        R0[2] = txz + twy;
        R0[3] = txy + twz;       R0[4] = 1 - (txx + tzz);  R0[5] = tyz - twx;
        R0[6] = txz - twy;       R0[7] = tyz + twx;        R0[8] = 1 - (txx + tyy);
    }
    double R[kMaxJoints][9], P[kMaxJoints][3];
    const Tree &t = a.tree;
    for (int j = 0; j < t.n_joints; ++j) {
        const int pj = t.parent[j];
        const double *Rp = pj < 0 ? R0 : R[pj], *Pp = pj < 0 ? root_p : P[pj];
        double Rl[9];
        for (int k = 0; k < 9; ++k) Rl[k] = t.origin_rot[9 * j + k];
        const int qi = t.qidx[j];
        if (qi >= 0) {  // Rodrigues about the joint axis (Eigen AngleAxis::toRotationMatrix)
            const double th = dof[qi], c = cos(th), s = sin(th), C = 1 - c;
            const double ax = t.axis[3 * j], ay = t.axis[3 * j + 1], az = t.axis[3 * j + 2];
            const double Rj[9] = {c + ax * ax * C, ax * ay * C - az * s, ax * az * C + ay * s, ay * ax * C + az * s, c + ay * ay * C,
                                  ay * az * C - ax * s, az * ax * C - ay * s, az * ay * C + ax * s, c + az * az * C};
            double tmp[9];
            mat3_mul(Rl, Rj, tmp);
            for (int k = 0; k < 9; ++k) Rl[k] = tmp[k];
        }
        mat3_mul(Rp, Rl, R[j]);
        const double ox = t.origin_xyz[3 * j], oy = t.origin_xyz[3 * j + 1], oz = t.origin_xyz[3 * j + 2];
        for (int k = 0; k < 3; ++k) P[j][k] = Pp[k] + (Rp[3 * k] * ox + Rp[3 * k + 1] * oy + Rp[3 * k + 2] * oz);
    }
    for (int b = 0; b < a.n_bodies; ++b) {
        const int j = a.body_joint[b];
        const double *Rb = j < 0 ? R0 : R[j], *Pb = j < 0 ? root_p : P[j];
        float *bp = a.body_positions + ((int64_t)n * a.n_bodies + b) * 3;
        bp[0] = (float)Pb[0]; bp[1] = (float)Pb[1]; bp[2] = (float)Pb[2];
        store_quat_from_matrix(Rb, a.body_rotations + ((int64_t)n * a.n_bodies + b) * 4);
    }
}

// scipy.ndimage "reflect" boundary: (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

// gaussian_filter1d(sigma=1, truncate=4): radius 4, weights exp(-x^2/2) / sum; scipy's correlate1d takes the symmetric
// branch: out = in[l] w[4] + sum_{ll=-4..-1} (in[l+ll] + in[l-ll]) w[ll+4], accumulated in float64
// The five distinct weights are computed by the caller exactly as scipy does (numpy exp and its pairwise sum) and passed in.
struct GaussW {
    double w[5];  // w[0] = centre, w[k] = offset +-k
};

// Raw derivative of frame m for column c of a float64 (N, C) array: data_convert.py:292-295
__device__ __forceinline__ double cdiff_f64(const double *p, int m, int n, int C, int c, double dt) {
    if (n == 1) return 0.0;
    if (m == 0) return (p[(int64_t)C + c] - p[c]) / dt;
    if (m == n - 1) return (p[(int64_t)(n - 1) * C + c] - p[(int64_t)(n - 2) * C + c]) / dt;
    return (p[(int64_t)(m + 1) * C + c] - p[(int64_t)(m - 1) * C + c]) / (2 * dt);
}

// the same for a float32 array: numpy keeps float32 (the python-float divisor is a weak scalar): f32(a - b) / f32(2 dt)
__device__ __forceinline__ float cdiff_f32(const float *p, int m, int n, int C, int c, double dt) {
    if (n == 1) return 0.0f;
    if (m == 0) return __fdiv_rn(__fsub_rn(p[(int64_t)C + c], p[c]), (float)dt);
    if (m == n - 1) return __fdiv_rn(__fsub_rn(p[(int64_t)(n - 1) * C + c], p[(int64_t)(n - 2) * C + c]), (float)dt);
    return __fdiv_rn(__fsub_rn(p[(int64_t)(m + 1) * C + c], p[(int64_t)(m - 1) * C + c]), (float)(2 * dt));
}

__global__ void velocity_f64_kernel(const double *__restrict__ p, int n, int C, double dt, double *__restrict__ out, GaussW g) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)n * C; e += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(e / C), c = (int)(e - (int64_t)l * C);
        double acc = cdiff_f64(p, l, n, C, c, dt) * g.w[0];
        for (int ll = -4; ll <= -1; ++ll)
            acc += (cdiff_f64(p, reflect(l + ll, n), n, C, c, dt) + cdiff_f64(p, reflect(l - ll, n), n, C, c, dt)) * g.w[-ll];
        out[e] = acc;
    }
}

__global__ void velocity_f32_kernel(const float *__restrict__ p, int n, int C, double dt, float *__restrict__ out, GaussW g) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)n * C; e += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(e / C), c = (int)(e - (int64_t)l * C);
        double acc = (double)cdiff_f32(p, l, n, C, c, dt) * g.w[0];
        for (int ll = -4; ll <= -1; ++ll)
            acc += ((double)cdiff_f32(p, reflect(l + ll, n), n, C, c, dt) + (double)cdiff_f32(p, reflect(l - ll, n), n, C, c, dt)) * g.w[-ll];
        out[e] = (float)acc;
    }
}

// compute_angular_velocity (data_convert.py:87-108) of one pair of (w, x, y, z) float32 quaternions, component k.
// The reference evaluates this in float32, where 2 acos(w) of two nearly equal rotations is ill-conditioned (one ulp of w is
// ~0.04 rad/s at 60 fps) and the result depends on the BLAS behind np.linalg.norm; here the SAME expression is evaluated in
// float64 on the same float32 inputs, i.e. the value the reference's float32 evaluation scatters around.
__device__ __forceinline__ void ang_vel_pair(const float *qa, const float *qb, double dt, double *out) {
    const double w1 = qa[0], x1 = qa[1], y1 = qa[2], z1 = qa[3];
    double nsq = w1 * w1 + x1 * x1 + y1 * y1 + z1 * z1;
    if (nsq < 1e-8) nsq = 1e-8;
    const double iw = w1 / nsq, ix = -x1 / nsq, iy = -y1 / nsq, iz = -z1 / nsq;
    const double w2 = qb[0], x2 = qb[1], y2 = qb[2], z2 = qb[3];
    double r[4] = {iw * w2 - ix * x2 - iy * y2 - iz * z2, iw * x2 + ix * w2 + iy * z2 - iz * y2, iw * y2 - ix * z2 + iy * w2 + iz * x2,
                   iw * z2 + ix * y2 - iy * x2 + iz * w2};
    out[0] = out[1] = out[2] = 0.0;
    const double nrm = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    if (nrm < 1e-8) return;
    for (int k = 0; k < 4; ++k) r[k] /= nrm;
    if (r[0] < 0.0)
        for (int k = 0; k < 4; ++k) r[k] = -r[k];
    const double w = fmin(fmax(r[0], -1.0), 1.0);
    const double angle = 2.0 * acos(w), sin_half = sqrt(1.0 - w * w);
    if (sin_half < 1e-8) return;
    for (int k = 0; k < 3; ++k) out[k] = (angle / dt) * (r[1 + k] / sin_half);
}

// raw angular velocity of frame m, body b (float32 array in the reference): data_convert.py:357-365
__device__ __forceinline__ void ang_vel_raw(const float *rot, int m, int n, int B, int b, double dt, float *out) {
    out[0] = out[1] = out[2] = 0.0f;
    if (n < 2) return;
    auto q = [&](int f) { return rot + ((int64_t)f * B + b) * 4; };
    double a[3], c[3];
    if (m == 0) {
        ang_vel_pair(q(0), q(1), dt, a);
        for (int k = 0; k < 3; ++k) out[k] = (float)a[k];
    } else if (m == n - 1) {
        ang_vel_pair(q(n - 2), q(n - 1), dt, a);
        for (int k = 0; k < 3; ++k) out[k] = (float)a[k];
    } else {
        ang_vel_pair(q(m - 1), q(m), dt, a);
        ang_vel_pair(q(m), q(m + 1), dt, c);
        for (int k = 0; k < 3; ++k) out[k] = (float)(0.5 * (a[k] + c[k]));
    }
}

__global__ void ang_vel_raw_kernel(const float *__restrict__ rot, int n, int B, double dt, float *__restrict__ raw) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)n * B; e += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(e / B), b = (int)(e - (int64_t)m * B);
        ang_vel_raw(rot, m, n, B, b, dt, raw + e * 3);
    }
}

// gaussian_filter1d(sigma=1, axis=0) of a float32 (n, C) array (float64 accumulation, float32 result)
__global__ void gaussian_f32_kernel(const float *__restrict__ in, int n, int C, float *__restrict__ out, GaussW g) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)n * C; e += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(e / C), c = (int)(e - (int64_t)l * C);
        double acc = (double)in[e] * g.w[0];
        for (int ll = -4; ll <= -1; ++ll)
            acc += ((double)in[(int64_t)reflect(l + ll, n) * C + c] + (double)in[(int64_t)reflect(l - ll, n) * C + c]) * g.w[-ll];
        out[e] = (float)acc;
    }
}

}  // namespace dataset
}  // namespace amp

using namespace amp;
using namespace amp::dataset;

extern "C" {

int64_t amp_dataset_scratch_bytes(int32_t n_in, int32_t n_out, int32_t n_bodies) {
    // key quaternions (n_in, 4) + rotation vectors (n_in, 3) in float64, raw angular velocities (n_out, B, 3) in float32
    return (int64_t)n_in * 7 * 8 + (int64_t)n_out * n_bodies * 3 * 4 + 256;
}

int amp_dataset_interp_fk(const amp_dataset_desc_t *d, double *dof_positions, float *body_positions, float *body_rotations,
                          double *root_pose, void *scratch, int64_t scratch_bytes, void *stream) {
    AMP_REQUIRE(d && dof_positions && body_positions && body_rotations, "amp_dataset_interp_fk: NULL argument");
    AMP_REQUIRE(d->n_in >= 2 && d->n_out >= 1 && d->n_cols == 7 + d->n_dofs && d->n_dofs >= 0 && d->n_bodies >= 1,
                "amp_dataset_interp_fk: need at least two input frames and n_cols == 7 + n_dofs (got n_in %d, n_cols %d, n_dofs %d)",
                d->n_in, d->n_cols, d->n_dofs);
    AMP_REQUIRE(d->n_joints >= 0 && d->n_joints <= kMaxJoints, "amp_dataset_interp_fk: at most %d URDF joints are supported (got %d)",
                kMaxJoints, d->n_joints);
    AMP_REQUIRE(scratch && scratch_bytes >= amp_dataset_scratch_bytes(d->n_in, d->n_out, d->n_bodies),
                "amp_dataset_interp_fk: scratch buffer too small");
    cudaStream_t st = as_stream(stream);
    double *key_quat = static_cast<double *>(scratch);
    double *rotvec = key_quat + (int64_t)d->n_in * 4;
    slerp_keys_kernel<<<(d->n_in + 127) / 128, 128, 0, st>>>(d->rows, d->n_in, d->n_cols, key_quat, rotvec);
    AMP_CUDA_TRY(cudaGetLastError());
    FkArgs a{};
    a.rows = d->rows;
    a.n_in = d->n_in;
    a.n_cols = d->n_cols;
    a.n_out = d->n_out;
    a.n_dofs = d->n_dofs;
    a.n_bodies = d->n_bodies;
    a.t_orig = d->t_orig;
    a.t_new = d->t_new;
    a.lerp_lo = d->lerp_lo;
    a.slerp_ind = d->slerp_ind;
    a.slerp_alpha = d->slerp_alpha;
    a.key_quat = key_quat;
    a.rotvec = rotvec;
    a.tree = Tree{d->n_joints, d->joint_parent, d->joint_qidx, d->joint_origin_xyz, d->joint_origin_rot, d->joint_axis};
    a.body_joint = d->body_joint;
    a.dof_positions = dof_positions;
    a.body_positions = body_positions;
    a.body_rotations = body_rotations;
    a.root_pose = root_pose;
    interp_fk_kernel<<<(d->n_out + 127) / 128, 128, 0, st>>>(a);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

int amp_dataset_velocities(int32_t n_out, int32_t n_dofs, int32_t n_bodies, double dt, const double *gauss_w, const double *dof_positions,
                           const float *body_positions, const float *body_rotations, double *dof_velocities,
                           float *body_linear_velocities, float *body_angular_velocities, void *scratch, int64_t scratch_bytes,
                           void *stream) {
    AMP_REQUIRE(n_out >= 1 && n_dofs >= 0 && n_bodies >= 1 && dt > 0.0, "amp_dataset_velocities: bad sizes");
    AMP_REQUIRE(gauss_w, "amp_dataset_velocities: NULL gaussian weights");
    GaussW g;
    for (int k = 0; k < 5; ++k) g.w[k] = gauss_w[k];
    AMP_REQUIRE(dof_positions && body_positions && body_rotations && dof_velocities && body_linear_velocities && body_angular_velocities,
                "amp_dataset_velocities: NULL argument");
    AMP_REQUIRE(scratch && scratch_bytes >= (int64_t)n_out * n_bodies * 3 * 4, "amp_dataset_velocities: scratch buffer too small");
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    auto grid = [&](int64_t elems) { return (int)std::max<int64_t>(1, std::min<int64_t>((elems + 255) / 256, (int64_t)sms * 8)); };
    if (n_dofs > 0) velocity_f64_kernel<<<grid((int64_t)n_out * n_dofs), 256, 0, st>>>(dof_positions, n_out, n_dofs, dt, dof_velocities, g);
    velocity_f32_kernel<<<grid((int64_t)n_out * n_bodies * 3), 256, 0, st>>>(body_positions, n_out, n_bodies * 3, dt, body_linear_velocities, g);
    float *raw = static_cast<float *>(scratch);
    ang_vel_raw_kernel<<<grid((int64_t)n_out * n_bodies), 256, 0, st>>>(body_rotations, n_out, n_bodies, dt, raw);
    gaussian_f32_kernel<<<grid((int64_t)n_out * n_bodies * 3), 256, 0, st>>>(raw, n_out, n_bodies * 3, body_angular_velocities, g);
    AMP_CUDA_TRY(cudaGetLastError());
    return AMP_OK;
}

}  // extern "C"
