// Device math of the AMP path, written with explicitly rounded intrinsics (__fmul_rn / __dadd_rn ...) so that no
// multiply-add is ever contracted into an FMA: the reference evaluates every torch / numpy op with its own rounding and
// two hard thresholds in the slerp make the result sensitive to 1-ulp changes.  The file is also
// compiled with -fmad=false as a second line of defence.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace amp {

// ---- float64 frame index / blend : reference motions/motion_loader.py:294-305 -------------------------------------
struct FrameBlend {
    int64_t i0, i1;  // GLOBAL frame indices (trajectory start + local index)
    double blend;    // float64 blend as the reference method returns it
};

// np.clip(x, 0, 1): NaN propagates (np.minimum/np.maximum semantics), unlike fmin/fmax.
__device__ __forceinline__ double clip01(double x) {
    x = (x < 0.0) ? 0.0 : x;
    x = (x > 1.0) ? 1.0 : x;
    return x;
}

// phase = clip(t/dur, 0, 1); l0 = rint(phase*span) (half-even); l1 = min(l0+1, span);
// blend = rint(((t - l0*dt)/dt) * 1e5) / 1e5      (numpy round(decimals=5) == multiply, rint, divide)
__device__ __forceinline__ FrameBlend frame_blend(double t, double dur, int64_t start, int64_t end, double dt) {
    const int64_t span = end - start;
    const double phase = clip01(__ddiv_rn(t, dur));
    const int64_t l0 = __double2ll_rn(__dmul_rn(phase, (double)span));  // cvt.rni: round-half-even like np.rint
    const int64_t l1 = (l0 + 1 < span) ? (l0 + 1) : span;
    double b = __ddiv_rn(__dsub_rn(t, __dmul_rn((double)l0, dt)), dt);
    b = __ddiv_rn(rint(__dmul_rn(b, 1e5)), 1e5);
    FrameBlend r;
    r.i0 = start + l0;
    r.i1 = start + l1;
    r.blend = b;
    return r;
}

// ---- fp32 lerp : motion_loader.py:215   (1.0 - blend) * a + blend * b ------------------------------------------------
__device__ __forceinline__ float lerp_w(float one_minus_blend, float blend, float a, float b) {
    return __fadd_rn(__fmul_rn(one_minus_blend, a), __fmul_rn(blend, b));
}
__device__ __forceinline__ float lerp(float blend, float a, float b) {
    return lerp_w(__fsub_rn(1.0f, blend), blend, a, b);
}

// ---- float64 acos / sin for the slerp --------------------------------------------------------------------------------
// The reference evaluates torch.acos / torch.sin in fp32 (SLEEF, <= 1 ulp).  Here each is evaluated in float64 and rounded
// ONCE to fp32, i.e. the correctly rounded fp32 value (verified on 3.1 M / 4.5 M fp32 arguments against numpy float64:
// zero mismatches after rounding), with branch-free polynomials instead of libdevice's general-purpose acos()/sin()
// (whose slow paths cost ~2000 instructions per frame and most of the small-batch latency).
//   acos(c), c in [0, 1]:   c <= 0.5: pi/2 - asin(c);   c > 0.5: 2 asin(sqrt((1 - c)/2));   asin(s) = s + s z Q(z), z = s^2,
//                           Q = degree-10 minimax fit on z in [0, 0.25], relative error 1e-15
//   sin(x), |x| < 1e6:      3-term Cody-Waite reduction by pi/2, Taylor kernels of degree 15 / 16 on |r| <= pi/4
__device__ __forceinline__ double asin_kernel(double s, double z) {
    double q = 0.02797051111814061;
    q = fma(q, z, -0.006944262741785383);
    q = fma(q, z, 0.01550894988389679);
    q = fma(q, z, 0.010271259188591278);
    q = fma(q, z, 0.01414412466480873);
    q = fma(q, z, 0.017336854258322456);
    q = fma(q, z, 0.022373031645160257);
    q = fma(q, z, 0.030381916716773705);
    q = fma(q, z, 0.044642857591684446);
    q = fma(q, z, 0.07499999999719918);
    q = fma(q, z, 0.16666666666666943);
    return fma(s * z, q, s);
}

// c >= 0 (the caller has taken |c|); c > 1 yields NaN like acos
__device__ __forceinline__ double acos_unit(double c) {
    if (c > 1.0) return __longlong_as_double(0x7ff8000000000000LL);
    if (c <= 0.5) return 0x1.921fb54442d18p+0 - asin_kernel(c, c * c);
    const double z = (1.0 - c) * 0.5;
    return 2.0 * asin_kernel(sqrt(z), z);
}

__device__ __forceinline__ double sin_poly(double r, double r2) {  // sin(r), |r| <= pi/4
    double ps = -7.647163731819816e-13;
    ps = fma(ps, r2, 1.6059043836821613e-10);
    ps = fma(ps, r2, -2.505210838544172e-08);
    ps = fma(ps, r2, 2.7557319223985893e-06);
    ps = fma(ps, r2, -0.0001984126984126984);
    ps = fma(ps, r2, 0.008333333333333333);
    ps = fma(ps, r2, -0.16666666666666666);
    return fma(ps * r2, r, r);
}

__device__ __forceinline__ double sin_reduced(double x) {
    // Warp-uniform fast path: consecutive frames differ by small rotations, so for interpolation every lane's argument
    // is within pi/4 and neither the range reduction nor the cosine kernel is needed.  (For |x| <= pi/4 the general path
    // also gets k = 0, r = x and the same polynomial, so which path a warp takes does not change the value -- only within
    // an ulp of the boundary could k be 1, where the two kernels agree to 1e-16, far below the fp32 rounding that follows.)
    if (__all_sync(__activemask(), fabs(x) <= 0x1.921fb54442d18p-1)) return sin_poly(x, x * x);
    if (!(fabs(x) < 1.0e6)) return sin(x);  // absurd blends (and NaN / inf): libdevice's full-range path
    const double k = rint(x * 0x1.45f306dc9c883p-1);  // x * 2/pi
    double r = fma(-k, 0x1.921fb54400000p+0, x);
    r = fma(-k, 0x1.0b4611a600000p-34, r);
    r = fma(-k, 0x1.3198a2e037073p-69, r);
    const double r2 = r * r;
    const int q = (int)(long long)k & 3;
    double v;
    if (q & 1) {
        double pc = 4.779477332387385e-14;
        pc = fma(pc, r2, -1.1470745597729725e-11);
        pc = fma(pc, r2, 2.08767569878681e-09);
        pc = fma(pc, r2, -2.755731922398589e-07);
        pc = fma(pc, r2, 2.48015873015873e-05);
        pc = fma(pc, r2, -0.001388888888888889);
        pc = fma(pc, r2, 0.041666666666666664);
        pc = fma(pc, r2, -0.5);
        v = fma(pc, r2, 1.0);
    } else {
        v = sin_poly(r, r2);
    }
    return (q & 2) ? -v : v;
}

// ---- shortest-arc slerp, wxyz : motion_loader.py:247-279 -------------------------------------------------------------
// float4 holds (w, x, y, z) in (.x, .y, .z, .w).
__device__ __forceinline__ float4 slerp(float4 q0, float4 q1, float blend) {
    // ((w0*w1 + x0*x1) + y0*y1) + z0*z1, every product and sum rounded on its own (:248-253)
    float c = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0.x, q1.x), __fmul_rn(q0.y, q1.y)), __fmul_rn(q0.z, q1.z)),
                        __fmul_rn(q0.w, q1.w));
    if (c < 0.0f) {  // q1[neg_mask] = -q1[neg_mask] (:255-257); NaN compares false, as in torch
        q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z; q1.w = -q1.w;
    }
    c = fabsf(c);
    // acos / sin: float64 evaluation rounded once to fp32 (see above) = the correctly rounded value of each fp32 op, which
    // torch's SLEEF kernels match in all but their <= 1 ulp cases.  libdevice acosf/sinf (up to 2 ulp off) would add a
    // second, independent error that extrapolation (|blend| up to K-1) amplifies past the 1e-6 absolute bar.
    const float half = __double2float_rn(acos_unit((double)c));       // NaN for c > 1, masked below
    const float s = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(c, c)));     // sqrt(1.0 - c*c), unfused (:262)
    const float ra = __fdiv_rn(__double2float_rn(sin_reduced((double)__fmul_rn(__fsub_rn(1.0f, blend), half))), s);
    const float rb = __fdiv_rn(__double2float_rn(sin_reduced((double)__fmul_rn(blend, half))), s);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(ra, q0.x), __fmul_rn(rb, q1.x));
    o.y = __fadd_rn(__fmul_rn(ra, q0.y), __fmul_rn(rb, q1.y));
    o.z = __fadd_rn(__fmul_rn(ra, q0.z), __fmul_rn(rb, q1.z));
    o.w = __fadd_rn(__fmul_rn(ra, q0.w), __fmul_rn(rb, q1.w));
    if (fabsf(s) < 0.001f) {  // midpoint fallback, ignores blend (:275-277); false for NaN s
        o.x = __fadd_rn(__fmul_rn(0.5f, q0.x), __fmul_rn(0.5f, q1.x));
        o.y = __fadd_rn(__fmul_rn(0.5f, q0.y), __fmul_rn(0.5f, q1.y));
        o.z = __fadd_rn(__fmul_rn(0.5f, q0.z), __fmul_rn(0.5f, q1.z));
        o.w = __fadd_rn(__fmul_rn(0.5f, q0.w), __fmul_rn(0.5f, q1.w));
    }
    if (c >= 1.0f) o = q0;  // identity fallback, also masks the NaNs of acos(c > 1) (:278)
    return o;
}

// ---- quaternion_to_tangent_and_normal : g1_amp_env.py:489-497 -------------------------------------------------------
// quat_apply(q, v) = v + w*t + xyz x t with t = 2*(xyz x v), specialised for v = x-hat and v = z-hat keeping the
// rounding sequence of the generic expression (products rounded separately, then added).
__device__ __forceinline__ void tangent_normal(float4 q, float tn[6]) {
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    const float x2 = __fmul_rn(x, __fmul_rn(2.0f, x));
    const float y2 = __fmul_rn(y, __fmul_rn(2.0f, y));
    const float z2 = __fmul_rn(z, __fmul_rn(2.0f, z));
    const float wx = __fmul_rn(w, __fmul_rn(2.0f, x));
    const float wy = __fmul_rn(w, __fmul_rn(2.0f, y));
    const float wz = __fmul_rn(w, __fmul_rn(2.0f, z));
    const float xy = __fmul_rn(x, __fmul_rn(2.0f, y));
    const float xz = __fmul_rn(x, __fmul_rn(2.0f, z));
    const float yz = __fmul_rn(y, __fmul_rn(2.0f, z));
    tn[0] = __fadd_rn(1.0f, __fsub_rn(-y2, z2));  // tangent = R(q) x-hat
    tn[1] = __fadd_rn(wz, xy);
    tn[2] = __fadd_rn(-wy, xz);
    tn[3] = __fadd_rn(wy, xz);                    // normal = R(q) z-hat
    tn[4] = __fadd_rn(-wx, yz);
    tn[5] = __fadd_rn(1.0f, __fsub_rn(-x2, y2));
}

// ---- torch.clamp(x, lo, hi) / torch.maximum(x, c): NaN propagates (fminf / fmaxf would return the bound) -------------------
// (max.NaN / min.NaN are single FMNMX instructions)
__device__ __forceinline__ float max_nan(float x, float c) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(c));
    return r;
}
__device__ __forceinline__ float min_nan(float x, float c) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(c));
    return r;
}
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) { return min_nan(max_nan(x, lo), hi); }

// ---- skrl AMP style reward : -log(max(1 - 1/(1+exp(-d)), 1e-4)) * scale ------------------------------------------------
__device__ __forceinline__ float style_reward(float logit, float scale) {
    const float e = expf(-logit);
    const float p = __fsub_rn(1.0f, __fdiv_rn(1.0f, __fadd_rn(1.0f, e)));
    return __fmul_rn(-logf(max_nan(p, 0.0001f)), scale);  // torch.maximum: a NaN logit gives a NaN reward, as in the reference
}

}  // namespace amp
