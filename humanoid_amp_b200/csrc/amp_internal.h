// Internal declarations shared by the translation units of libamp_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "amp_b200.h"

namespace amp {

// Thread-local error text behind amp_last_error().
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define AMP_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return ::amp::cuda_fail(_e, #expr); \
    } while (0)

#define AMP_REQUIRE(cond, ...)                                        \
    do {                                                              \
        if (!(cond)) return ::amp::fail(AMP_EINVAL, __VA_ARGS__);     \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached multiprocessor count of the current device (148 on B200)

constexpr int kMaxKeyBodies = 16;

// Device-resident view of a staged motion library; passed BY VALUE to kernels.
struct LibView {
    int64_t num_frames;
    int32_t num_dofs;    // D_clip
    int32_t num_bodies;  // B
    int32_t num_traj;
    int32_t obs_dofs;    // D (robot dofs in the observation); 0 when no env selection was given
    int32_t num_keys;    // Kb
    int32_t obs_width;   // A = 2D + 13 + 3Kb
    int32_t row_floats;  // R = A rounded up to a multiple of 4 (16-byte rows)
    int32_t _pad;
    double dt;
    const int64_t *starts;   // [num_traj]
    const int64_t *ends;     // [num_traj]
    const double *durations; // [num_traj]
    const float *dof_pos, *dof_vel, *body_pos, *body_rot, *body_lin, *body_ang;
    const float *packed;     // (F, R) packed AMP rows, see amp_motion.cu
    const float *lerp_rows;  // (F, lerp_stride) all linearly interpolated columns of a frame side by side, or NULL
    int32_t lerp_width;      // W = 2*D_clip + 9*B
    int32_t lerp_stride;     // W rounded up to a multiple of 4
    uint32_t *flags;         // sticky error bits
};

// amp_disc.cu helper shared with amp_disc_train.cu (all pointers device; no allocation):
//   scaler_stats_to_f32   mean_f = (float)mean, denom_f = sqrt((float)var) + 1e-8 (skrl RunningStandardScaler, eval form)
int scaler_stats_to_f32(const double *mean, const double *var, int n, float *mean_f, float *denom_f, cudaStream_t st);

}  // namespace amp

struct amp_lib {
    amp::LibView v;
    int device;
    void *owned[8];  // device allocations freed by amp_lib_destroy
    int n_owned;
    int collect_table;  // AMP_OPT_COLLECT_TABLE: 0 auto, 1 global-memory table, 2 shared-memory table
};
