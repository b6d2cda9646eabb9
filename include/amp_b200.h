/*
 * amp_b200.h -- C ABI of the B200-native AMP hot path (libamp_b200.so, sm_100a).
 *
 * The reference (zhoushanghai/humanoid_amp) has no FFI for this path: its boundary is a Python call surface
 * (SURVEY.md section 8b).  Every entry point below names the reference function it replaces (file:line into the
 * reference tree).  The Python shims in humanoid_amp_b200/ keep the reference signatures and call these through ctypes.
 *
 * Conventions
 *   - every pointer marked "device" is a caller-owned CUDA device pointer on the handle's device; "host" pointers are
 *     read during the call only;
 *   - every call enqueues on the passed cudaStream_t (void*; NULL = legacy default stream) and returns without
 *     synchronising; no allocation happens after *_create;
 *   - return value: 0 on success, a negative AMP_E* code otherwise; amp_last_error() gives a thread-local message;
 *   - a handle belongs to one device; it is not safe for concurrent calls from several host threads, different
 *     handles are independent;
 *   - quaternions are wxyz, all float data is fp32, times are fp64, frame indices / motion ids are int64.
 */
#ifndef AMP_B200_H
#define AMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMP_B200_ABI_VERSION 1

/* The library is built with -fvisibility=hidden; only the entry points below are exported. */
#if defined(__GNUC__)
#define AMP_API __attribute__((visibility("default")))
#else
#define AMP_API
#endif

enum {
    AMP_OK = 0,
    AMP_EINVAL = -1,   /* bad argument (NULL pointer, negative size, misaligned buffer, ...) */
    AMP_ECUDA = -2,    /* a CUDA runtime call or launch failed; amp_last_error() carries cudaGetErrorString */
    AMP_ENODEV = -3,   /* no CUDA device / device is not sm_100 */
    AMP_ERANGE = -4,   /* a motion id was outside [0, num_trajectories) or a time was NaN (see amp_lib_poll_flags) */
    AMP_ENOMEM = -5,
    AMP_ENOTSUP = -6   /* the platform lacks a capability the call needs (NVSwitch multicast, POSIX-fd memory sharing) */
};

typedef struct amp_lib amp_lib_t;   /* a motion library staged on one device */
typedef struct amp_disc amp_disc_t; /* discriminator weights staged for the tensor-core kernel */
typedef struct amp_disc_train amp_disc_train_t; /* workspaces of the discriminator loss + gradient step */
typedef struct amp_bucket amp_bucket_t; /* one rank's flat gradient bucket, shared with the node's other ranks over NVLink */

/* Description of a loaded motion library: what the reference MotionLoader.__init__ leaves behind
 * (motions/motion_loader.py:98-164) plus the env's index lists (g1_amp_env.py:40-60). */
typedef struct {
    int64_t num_frames;        /* F, frames of all clips concatenated */
    int32_t num_dofs;          /* D_clip, columns of dof_positions / dof_velocities */
    int32_t num_bodies;        /* B, bodies per frame in the clip */
    int32_t num_trajectories;  /* clips concatenated */
    int32_t _pad0;
    double dt;                 /* 1.0 / fps of the FIRST clip (motion_loader.py:122) */
    const int64_t *traj_starts; /* host [num_trajectories] first global frame of each clip (:132) */
    const int64_t *traj_ends;   /* host [num_trajectories] last global frame of each clip, inclusive (:134) */
    const double *durations;    /* host [num_trajectories] dt * (frames - 1) (:135) */
    /* the six fp32 tensors of the loader, device, C-contiguous (:141-158) */
    const float *dof_positions;            /* (F, D_clip) */
    const float *dof_velocities;           /* (F, D_clip) */
    const float *body_positions;           /* (F, B, 3) */
    const float *body_rotations;           /* (F, B, 4) wxyz */
    const float *body_linear_velocities;   /* (F, B, 3) */
    const float *body_angular_velocities;  /* (F, B, 3) */
    /* env-side selection used by the fused AMP-observation path; all host arrays, may be NULL/0 when only
     * amp_frame_blend / amp_sample_full are needed */
    const int32_t *dof_indexes;  /* host [num_obs_dofs] clip column of each robot dof = motion_dof_indexes (g1_amp_env.py:52-54) */
    int32_t num_obs_dofs;        /* D, robot dofs in the observation */
    int32_t ref_body_index;      /* motion_ref_body_index (:55-57) */
    const int32_t *key_body_indexes; /* host [num_key_bodies] motion_key_body_indexes (:58-60) */
    int32_t num_key_bodies;      /* Kb */
    int32_t _pad1;
} amp_lib_desc_t;

/* ---- library -------------------------------------------------------------------------------------------------- */
AMP_API int amp_b200_abi_version(void);
AMP_API const char *amp_last_error(void);
/* Makes `device` current for the calling thread in the library's (statically linked) CUDA runtime.  The Python shims
 * call it before every entry point so a process driving several GPUs always launches on the tensor's device. */
AMP_API int amp_set_device(int device);
/* SM count, major, minor of the current device (any pointer may be NULL). */
AMP_API int amp_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Stage a motion library: copies the trajectory tables and builds the packed AMP row table
 * (one row per frame holding both interpolation end points of the columns compute_obs consumes).
 * Replaces the device side of MotionLoader.__init__ (motion_loader.py:141-158) + the index gathers of
 * collect_reference_motions (g1_amp_env.py:478-484) which are folded into the packing. */
AMP_API int amp_lib_create(const amp_lib_desc_t *desc, void *stream, amp_lib_t **out);
AMP_API int amp_lib_destroy(amp_lib_t *lib);
/* AMP observation width A = 2*D + 13 + 3*Kb (0 if the library was created without the env selection). */
AMP_API int amp_lib_obs_width(const amp_lib_t *lib);
/* Reads and clears the device-side sticky flags (bit0: motion id out of range, bit1: NaN time). Synchronises the
 * stream.  The kernels clamp such inputs for memory safety; the reference raises IndexError instead. */
AMP_API int amp_lib_poll_flags(amp_lib_t *lib, void *stream, uint32_t *flags);
/* Per-handle tuning knobs (never change results, only which kernel variant runs).  Options:
 *   AMP_OPT_COLLECT_TABLE  where amp_collect_reference keeps the packed row table: 0 = choose by size (default),
 *                          1 = global memory (L1/L2), 2 = copied into shared memory once per persistent CTA (falls back to 1
 *                          when the table does not fit in 227 KB).  Initial value: environment variable
 *                          AMP_B200_COLLECT_TABLE = global | smem, read ONCE in amp_lib_create. */
enum { AMP_OPT_COLLECT_TABLE = 1 };
AMP_API int amp_lib_set_option(amp_lib_t *lib, int32_t option, int64_t value);

/* MotionLoader._compute_frame_blend (motion_loader.py:281-307), float64 on device, bit-exact.
 *   times  device f64[S]; motion_ids device i64[S] or NULL (= all zeros, the reference default at :366)
 *   idx0, idx1 device i64[S] global frame indices; blend32 device f32[S] or NULL (the cast at :371);
 *   blend64 device f64[S] or NULL (the value the reference method returns). */
AMP_API int amp_frame_blend(amp_lib_t *lib, const double *times, const int64_t *motion_ids, int64_t S, int64_t *idx0,
                    int64_t *idx1, float *blend32, double *blend64, void *stream);

/* MotionLoader.sample with given times (motion_loader.py:368-390): frame/blend + 5 lerps + slerp over ALL bodies.
 * Outputs device fp32, C-contiguous: dof_pos (S,D_clip) dof_vel (S,D_clip) body_pos (S,B,3) body_rot (S,B,4)
 * body_lin_vel (S,B,3) body_ang_vel (S,B,3); any output pointer may be NULL to skip it. */
AMP_API int amp_sample_full(amp_lib_t *lib, const double *times, const int64_t *motion_ids, int64_t S, float *dof_pos,
                    float *dof_vel, float *body_pos, float *body_rot, float *body_lin_vel, float *body_ang_vel,
                    void *stream);

/* MotionLoader._interpolate / _slerp with explicit end points (motion_loader.py:211-215, :242-279).
 *   a, b, out device f32[n * inner]; blend device f32[n] broadcast over inner.
 *   q0, q1, out device f32[n * bodies * 4]; blend device f32[n] broadcast over bodies. */
AMP_API int amp_lerp(const float *a, const float *b, const float *blend, int64_t n, int64_t inner, float *out, void *stream);
AMP_API int amp_slerp(const float *q0, const float *q1, const float *blend, int64_t n, int64_t bodies, float *out,
              void *stream);

/* env.collect_reference_motions (g1_amp_env.py:445-486) fused end to end: history times t - dt*k (:454-457),
 * frame/blend, gather+lerp of the consumed columns only, root slerp, compute_obs (:535-561), written as the
 * history-stacked row (slot 0 = newest).
 *   cur_times device f64[n]; motion_ids device i64[n] or NULL (= zeros, :462)
 *   out device f32: row r of the destination starts at out + r*row_stride and receives K*A floats.
 *   Destination row of sample i:  row_index ? row_index[i] : (start_row + i) % capacity_rows
 *   (capacity_rows <= 0 means no wrap).  This is how the result lands directly in amp_observation_buffer[env_ids]
 *   (:417-419) or in a skrl RandomMemory ring. */
AMP_API int amp_collect_reference(amp_lib_t *lib, const double *cur_times, const int64_t *motion_ids, int64_t n, int32_t K,
                          float *out, int64_t row_stride, int64_t capacity_rows, int64_t start_row,
                          const int64_t *row_index, void *stream);

/* Free function compute_obs (g1_amp_env.py:535-561) on contiguous inputs:
 *   dof_pos, dof_vel (n,D); root_pos (n,3); root_rot (n,4); root_lin_vel, root_ang_vel (n,3); key_pos (n,Kb,3)
 *   out (n, 2D+13+3Kb). */
AMP_API int amp_compute_obs(const float *dof_pos, const float *dof_vel, const float *root_pos, const float *root_rot,
                    const float *root_lin_vel, const float *root_ang_vel, const float *key_pos, int64_t n, int32_t D,
                    int32_t Kb, float *out, void *stream);
/* quaternion_to_tangent_and_normal (g1_amp_env.py:489-497): q (n,4) -> out (n,6). */
AMP_API int amp_tangent_normal(const float *q, int64_t n, float *out, void *stream);

/* Per-step env path, G1AmpEnv._get_observations AMP part (g1_amp_env.py:176-193): compute_obs from simulator state,
 * shift the history (slot i -> i+1), write slot 0, all in place in amp_buf (N,K,A).
 *   joint_pos, joint_vel (N,D); body_* (N,Bsim,3|4) as Isaac Lab's robot.data lays them out; ref_body and
 *   key_bodies (host [Kb]) index Bsim.  policy_obs: optional (N, policy_width) output receiving obs[:, :A-3Kb]
 *   (the "base actor obs" slice, :196) at row stride policy_stride; NULL to skip. */
AMP_API int amp_obs_step(const float *joint_pos, const float *joint_vel, const float *body_pos_w, const float *body_quat_w,
                 const float *body_lin_vel_w, const float *body_ang_vel_w, int64_t N, int32_t D, int32_t Bsim,
                 int32_t ref_body, const int32_t *key_bodies, int32_t Kb, int32_t K, float *amp_buf, float *policy_obs,
                 int64_t policy_stride, void *stream);

/* Policy ("actor") observation of G1AmpEnv._get_observations (g1_amp_env.py:195-242) -- SURVEY.md section 8f item 1.
 *   amp_buf (N,K,A): slot 0 holds this step's compute_obs row (written by amp_obs_step); base = its first `base_width`
 *   columns (A - 3*Kb).  last_actions (N, act).  command (N, cmd) or NULL with cmd = 0 (cfg.rew_track_vel <= 0).
 *   n = num_actor_observations.  n == 1: actor_obs = [base | last_actions | command].
 *   n  > 1: current = [base | last_actions | command]; history frame = [base | last_actions? | command?] (the two include
 *   flags); hist_buf (N, n-1, P) is updated in place: envs whose just_reset byte is set get every slot = the new history
 *   frame (warm start) and the byte is cleared, the others shift slot i -> i+1 and take the new frame in slot 0;
 *   actor_obs = [current | hist_buf flattened].  actor_obs rows are actor_stride floats apart. */
AMP_API int amp_actor_obs_step(const float *amp_buf, int64_t N, int32_t K, int32_t A, int32_t base_width,
                               const float *last_actions, int32_t act, const float *command, int32_t cmd,
                               int32_t num_actor_observations, int32_t hist_include_actions, int32_t hist_include_command,
                               float *hist_buf, uint8_t *just_reset, float *actor_obs, int64_t actor_stride, void *stream);

/* Task reward of G1AmpEnv._get_rewards (g1_amp_env.py:246-288) with compute_rewards (:564-606) and
 * exp_reward_with_floor (:500-532, sigma 0.5, floor 4.0) -- SURVEY.md section 8f item 1.
 *   scales host [6]: rew_termination, rew_action_l2, rew_joint_pos_limits, rew_joint_acc_l2, rew_joint_vel_l2, rew_track_vel
 *   reset_terminated (N) bytes; actions (N,act); joint_pos / joint_acc / joint_vel (N,D); soft_limits (N,D,2);
 *   body_lin_vel_w (N,Bsim,3), body_quat_w (N,Bsim,4), command (N,2): only read when rew_track_vel > 0.
 *   total (N); terms (N,6) or NULL (the six terms in the order of `scales`); track_err (N) or NULL. */
AMP_API int amp_task_reward(const float *scales, const uint8_t *reset_terminated, const float *actions, int32_t act,
                            const float *joint_pos, const float *soft_limits, const float *joint_acc, const float *joint_vel,
                            int32_t D, const float *body_lin_vel_w, const float *body_quat_w, int32_t Bsim, int32_t ref_body,
                            const float *command, int64_t N, float *total, float *terms, float *track_err, void *stream);

/* The whole per-step env path in ONE launch: G1AmpEnv._get_observations (g1_amp_env.py:175-242: compute_obs, AMP history
 * shift + slot 0, actor observation + its history with warm start) and, when reward_total != NULL, _get_rewards (:246-319)
 * evaluated on the SAME simulator state (valid when no reset happens between the two in the caller's step, e.g. play.py's
 * loop; otherwise call amp_task_reward before the resets and this entry point without the reward part after them, as
 * DirectRLEnv.step orders them).  Same results as amp_obs_step + amp_actor_obs_step + amp_task_reward; each simulator tensor
 * is read once.  Field meanings are those of the three entry points above. */
typedef struct {
    const float *joint_pos, *joint_vel;                                        /* device (N, D) */
    const float *body_pos_w, *body_quat_w, *body_lin_vel_w, *body_ang_vel_w;   /* device (N, Bsim, 3|4) */
    int64_t num_envs;
    int32_t num_dofs, num_sim_bodies, ref_body, num_key_bodies;
    const int32_t *key_bodies;        /* host [num_key_bodies] */
    int32_t num_amp_observations;     /* K */
    int32_t _pad0;
    float *amp_buf;                   /* device (N, K, A), in place */
    /* actor observation (actor_obs == NULL: skipped) */
    const float *last_actions;        /* device (N, action_size) */
    const float *command;             /* device (N, command_size) or NULL with command_size = 0 */
    int32_t action_size, command_size, num_actor_observations, hist_include_actions, hist_include_command, _pad1;
    float *hist_buf;                  /* device (N, n-1, P) or NULL when n == 1 */
    uint8_t *just_reset;              /* device (N) or NULL */
    float *actor_obs;                 /* device, rows actor_stride floats apart */
    int64_t actor_stride;
    /* task reward (reward_total == NULL: skipped) */
    const float *reward_scales;       /* host [6], as amp_task_reward */
    const uint8_t *reset_terminated;  /* device (N) */
    const float *actions;             /* device (N, action_size) */
    const float *soft_limits;         /* device (N, D, 2) */
    const float *joint_acc;           /* device (N, D) */
    float *reward_total, *reward_terms, *track_err; /* device (N), (N,6) or NULL, (N) or NULL */
} amp_env_step_t;
AMP_API int amp_env_step(const amp_env_step_t *args, void *stream);

/* ---- AMP memories and the state preprocessor (upstream skrl >= 1.4.3, not vendored) ------------------ */
/* skrl memories/torch/base.py Memory.sample_by_index (RandomMemory.sample draws the indexes): out[r,:] = src[row_index[r],:].
 *   src device f32 (capacity, W) with row stride src_stride floats; row_index device i64[M]; out device f32 (M, W).
 *   An index outside [0, capacity) (an IndexError in skrl) yields a zero row and sets bit 1 of *flags (device word or NULL). */
AMP_API int amp_gather_rows(const float *src, int64_t src_stride, int64_t capacity, const int64_t *row_index, int64_t M, int32_t W,
                            float *out, int64_t out_stride, uint32_t *flags, void *stream);
/* skrl RunningStandardScaler, train = True (_parallel_variance): merge the batch x (M, W) into the float64 running
 * statistics in place.  running_mean / running_variance device f64[W], current_count device f64[1]; scratch is a device
 * buffer of at least amp_scaler_scratch_bytes(W) bytes (no allocation inside the call). */
AMP_API int64_t amp_scaler_scratch_bytes(int32_t W);
AMP_API int amp_scaler_update(const float *x, int64_t x_stride, int64_t M, int32_t W, double *running_mean, double *running_variance,
                              double *current_count, void *scratch, int64_t scratch_bytes, void *stream);
/* skrl RunningStandardScaler, train = False: out = clamp((x - mean.float()) / (sqrt(var.float()) + epsilon), -clip, clip). */
AMP_API int amp_scaler_apply(const float *x, int64_t x_stride, int64_t M, int32_t W, const double *running_mean,
                             const double *running_variance, float epsilon, float clip, float *out, int64_t out_stride, void *stream);

/* ---- discriminator style reward (skrl AMP._update; cfg agents/skrl_g1_dance_amp_cfg.yaml:31-39, 80, 94-95) ------ */
/* Network Linear(in,h1)-ReLU-Linear(h1,h2)-ReLU-Linear(h2,1) on RunningStandardScaler-normalised input.
 * h1, h2 must be multiples of 256 (reference: 1024, 512); in_features any value >= 1 (padded to 64 internally). */
AMP_API int amp_disc_create(int32_t in_features, int32_t h1, int32_t h2, int64_t max_rows, void *stream, amp_disc_t **out);
AMP_API int amp_disc_destroy(amp_disc_t *d);
/* Rows one persistent wave covers (one 128-row tile per SM); 0 for a NULL handle.  max_rows of amp_disc_create is a hint
 * only: the scratch (two 128-row x_hat and h1 slots per SM, L2-resident) does not scale with the batch. */
AMP_API int64_t amp_disc_chunk_rows(const amp_disc_t *d);
/* Kernel launches amp_disc_style_reward issues for a batch of M rows (0 for M == 0): 1 for at most 9472 rows (two-CTA-per-tile
 * kernel) and for more than eight row tiles per SM (persistent kernel with in-kernel scaler + cast), else 2 (cast kernel + fused
 * kernel; per chunk of max_rows rows for inputs wider than 254 columns). */
AMP_API int64_t amp_disc_launch_count(const amp_disc_t *d, int64_t M);
/* Refresh the staged bf16 weights / fp32 biases / scaler statistics from the fp32 masters the trainer owns
 * (device pointers; W row-major (out,in) as torch.nn.Linear stores them; mean/var are the scaler's float64 buffers). */
AMP_API int amp_disc_load(amp_disc_t *d, const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                  const float *b3, const double *running_mean, const double *running_variance, void *stream);
/* x device f32 (M, in_features) with row stride x_stride floats; reward device f32[M];
 * logits device f32[M] or NULL.  reward = -log(max(1 - 1/(1+exp(-logit)), 1e-4)) * reward_scale. */
AMP_API int amp_disc_style_reward(amp_disc_t *d, const float *x, int64_t x_stride, int64_t M, float reward_scale,
                          float *reward, float *logits, void *stream);
/* The same on rows gathered from a memory: batch row r = memory[row_index[r], :] (skrl Memory.sample_by_index fused into the
 * preprocessor + discriminator, no (M, in_features) intermediate).  memory device f32 (capacity, in_features), row stride
 * memory_stride floats; row_index device i64[M]; an index outside [0, capacity) reads row 0 and sets bit 1 of *flags
 * (device word or NULL). */
AMP_API int amp_disc_style_reward_indexed(amp_disc_t *d, const float *memory, int64_t memory_stride, int64_t capacity,
                                          const int64_t *row_index, int64_t M, float reward_scale, float *reward, float *logits,
                                          uint32_t *flags, void *stream);
/* Standalone epilogue (logits -> reward) for callers that own the discriminator forward. */
AMP_API int amp_style_reward_from_logits(const float *logits, int64_t M, float reward_scale, float *reward, void *stream);

/* ---- discriminator loss + gradients (skrl AMP._update "compute discriminator loss" block and its backward
 *      pass; cfg agents/skrl_g1_dance_amp_cfg.yaml:89, 94, 96-98) ------------------------------------------------------- */
/* Workspaces for batches of up to max_batch_rows rows PER SOURCE (skrl: discriminator_batch_size, or the mini-batch length
 * when that is 0).  Same network shape constraints as amp_disc_create; in_features <= 1024. */
AMP_API int amp_disc_train_create(int32_t in_features, int32_t h1, int32_t h2, int64_t max_batch_rows, void *stream,
                                  amp_disc_train_t **out);
AMP_API int amp_disc_train_destroy(amp_disc_train_t *t);
/* Stage one of the three batches of the step: source 0 = agent AMP states (rollout memory), 1 = replay buffer, 2 = motion
 * dataset.  x device f32 (batch_rows, in_features), row stride x_stride floats; the three sources of a step have the same
 * batch_rows.
 * running_mean / running_variance: the amp_state_preprocessor's float64 buffers AFTER its train-mode update for this batch
 * (skrl calls the preprocessor with train=True on each batch in turn; use amp_scaler_update for that) -- the rows are
 * normalised, clipped to +-5 and rounded to bf16; both NULL = x is already normalised. */
AMP_API int amp_disc_train_stage(amp_disc_train_t *t, int32_t source, const float *x, int64_t x_stride, int64_t batch_rows,
                                 const double *running_mean, const double *running_variance, void *stream);
/* Loss and gradients of  loss_scale * (0.5 (BCE(cat(agent, replay), 0) + BCE(motion, 1)) + logit_reg * |W3|^2
 *                                      + grad_penalty * mean_rows |d logit / d motion_state|^2 + weight_decay * sum |W|^2)
 * for the three staged batches.  W1 (h1,in) b1 W2 (h2,h1) b2 W3 (1,h2) b3: fp32 masters, device, torch.nn.Linear layout.
 * gW1 .. gb3: device f32, same shapes, OVERWRITTEN with d loss / d parameter (they may point into the trainer's flat
 * all-reduce bucket).  terms: device f32[6] or NULL = {bce_agent_replay, bce_motion, logit_regularization (sum W3^2),
 * gradient_penalty, weight_decay (sum of all W^2), loss}.  logits: device f32[3 * Bp] or NULL, Bp = batch_rows rounded up to
 * 128; the logit of row r of source s is logits[s * Bp + r].  bf16 tensor-core operands, fp32 accumulation. */
AMP_API int amp_disc_train_step(amp_disc_train_t *t, const float *W1, const float *b1, const float *W2, const float *b2,
                                const float *W3, const float *b3, int64_t batch_rows, float loss_scale,
                                float logit_regularization_scale, float gradient_penalty_scale, float weight_decay_scale,
                                float *gW1, float *gb1, float *gW2, float *gb2, float *gW3, float *gb3, float *terms,
                                float *logits, void *stream);

/* ---- gradient all-reduce over peer memory (skrl Model.reduce_parameters, enabled by the
 *      reference at train.py:53-58, 184-196) ------------------------------------------------------------------------ */
/* One process per GPU, all ranks on one node.  A bucket is ONE flat fp32 device buffer (rounded up to 4 floats, zero
 * initialised) that gradient producers write into directly; amp_bucket_allreduce_mean averages it over the ranks in place
 * with one kernel per rank that loads / stores peer memory over NVLink (two-shot: rank r reduces slice r of every rank
 * and writes the result back to every rank).  world == 1 needs no export / connect and the all-reduce is a no-op.
 *   create  -> export (128 bytes: two CUDA IPC handles) -> [exchange the blobs of all ranks, rank order] -> connect.
 * Every rank must call amp_bucket_allreduce_mean the same number of times with the same range (like any collective). */
AMP_API int amp_bucket_create(int64_t floats, int32_t world, int32_t rank, amp_bucket_t **out);
AMP_API int amp_bucket_destroy(amp_bucket_t *b);
AMP_API int64_t amp_bucket_floats(const amp_bucket_t *b);
AMP_API float *amp_bucket_data(amp_bucket_t *b);
AMP_API int amp_bucket_export(amp_bucket_t *b, void *handles128);
AMP_API int amp_bucket_connect(amp_bucket_t *b, const void *all_handles);
/* In place: bucket[offset, offset+count) = mean over ranks.  offset % 4 == 0.  Enqueued on `stream`; spins on peers are
 * bounded (~2 s) and a timeout is reported by amp_bucket_poll_status (bit 0: a peer never announced its data, bit 1: a
 * peer never announced its stores), not by hanging the device. */
AMP_API int amp_bucket_allreduce_mean(amp_bucket_t *b, int64_t offset_floats, int64_t count, void *stream);
AMP_API int amp_bucket_poll_status(amp_bucket_t *b, void *stream, uint32_t *status);
/* %globaltimer stamps (ns) of the last all-reduce on this rank: kernel start, barrier A passed, own slice published,
 * barrier B passed.  Synchronises the stream.  For tools/bench_allreduce.py. */
AMP_API int amp_bucket_last_timing(amp_bucket_t *b, void *stream, uint64_t *stamps4);

/* Shared form of the bucket: the same object, but its memory is a virtual-memory-management allocation that every rank maps
 * (unicast, for the kernels above) AND that is bound to one NVSwitch multicast object, so that amp_bucket_allreduce_mean runs
 * in the switch: rank r reduces slice r with multimem.ld_reduce (the switch adds the W copies) and broadcasts the mean with
 * multimem.st -- every rank moves about half the bytes of the two-shot kernel and the sum is the switch's, identical on all
 * ranks.  Everything else (flags, epochs, timeouts, amp_disc_train_step_exchange) is unchanged.
 *   create_shared -> export_shared (64-byte IPC handle of the flags, a POSIX fd of the data, and on rank 0 the fd of the
 *   multicast object; the caller owns the fds and closes them after connect) -> [pass the fds between the processes, e.g.
 *   SCM_RIGHTS over a unix socket] -> join_shared (adds this rank's device to the multicast team) -> [make sure EVERY rank
 *   has joined: binding memory blocks until the team is complete] -> connect_shared (maps the peers, binds, maps the team).
 * AMP_ENOTSUP when the device has no multicast support or no POSIX-fd handles: use amp_bucket_create then.  A rank that fails
 * at any step destroys its handle; when the ranks agree on that after each step, nobody is left waiting. */
AMP_API int amp_bucket_create_shared(int64_t floats, int32_t world, int32_t rank, amp_bucket_t **out);
AMP_API int amp_bucket_export_shared(amp_bucket_t *b, void *flags_handle64, int32_t *data_fd, int32_t *multicast_fd);
AMP_API int amp_bucket_join_shared(amp_bucket_t *b, int32_t multicast_fd);
AMP_API int amp_bucket_connect_shared(amp_bucket_t *b, const void *all_flags_handles, const int32_t *data_fds);
/* 1 when amp_bucket_allreduce_mean on this bucket runs in the switch (shared form, connected), else 0. */
AMP_API int amp_bucket_in_switch(const amp_bucket_t *b);

/* amp_disc_train_step with the gradient exchange fused into its last kernel (replaces the discriminator's
 * share of skrl Model.reduce_parameters, train.py:184-196): gW1 .. gb3 must be six views of `bucket` that lie side by side
 * (any order, each starting on a multiple of 4 floats, the one-element gb3 last); on return of the stream they hold the MEAN
 * over the ranks of d loss / d parameter -- the reduction amp_bucket_allreduce_mean over that range would have made after
 * amp_disc_train_step (same summation: fixed rank order over peer memory, the switch's sum on a shared bucket), identical on
 * every rank.  The kernel that sums the split-K slices pushes each quad straight into the owning rank's staging area (peer
 * stores), the owner reduces and publishes; on a shared bucket it writes the local bucket and reduces its slice in the
 * switch.  No second launch.  terms / logits stay per rank.  A collective: every rank calls it at the same point of its bucket's call
 * sequence.  bucket world == 1: identical to amp_disc_train_step.  At most 8 ranks. */
AMP_API int amp_disc_train_step_exchange(amp_disc_train_t *t, const float *W1, const float *b1, const float *W2, const float *b2,
                                         const float *W3, const float *b3, int64_t batch_rows, float loss_scale,
                                         float logit_regularization_scale, float gradient_penalty_scale,
                                         float weight_decay_scale, float *gW1, float *gb1, float *gW2, float *gb2, float *gW3,
                                         float *gb3, float *terms, float *logits, amp_bucket_t *bucket, void *stream);

/* ---- offline dataset pipeline (reference motions/data_convert.py:161-379) ----------------------------- */
/* CSV rows at 30 fps -> 2N-1 frames at 60 fps (scipy interp1d / Slerp semantics) -> forward kinematics over the URDF tree
 * (Pinocchio forwardKinematics + updateFramePlacements, Eigen matrix -> quaternion) -> velocities.  All pointers device. */
typedef struct amp_dataset_desc {
    int32_t n_in, n_cols, n_out, n_dofs, n_bodies, n_joints;
    const float *rows;          /* (n_in, n_cols): root xyz, root quat xyzw, n_dofs joint angles (data_convert.py:178-183) */
    const double *t_orig;       /* [n_in]  np.linspace(0, (n_in-1)/30, n_in)            (:188) */
    const double *t_new;        /* [n_out] np.linspace(0, (n_in-1)/30, 2 n_in - 1)      (:193) */
    const int32_t *lerp_lo;     /* [n_out] lower knot of scipy interp1d: clip(searchsorted(t_orig, t_new), 1, n_in-1) - 1 */
    const int32_t *slerp_ind;   /* [n_out] scipy Slerp: searchsorted(t_orig, t_new) - 1, 0 where t_new == t_orig[0] */
    const double *slerp_alpha;  /* [n_out] (t_new - t_orig[ind]) / (t_orig[ind+1] - t_orig[ind]) */
    const int32_t *joint_parent;    /* [n_joints] topological order; joint whose child link is this joint's parent link, -1 = root */
    const int32_t *joint_qidx;      /* [n_joints] column of the joint-angle block driving the joint, -1 = fixed */
    const double *joint_origin_xyz; /* (n_joints, 3) */
    const double *joint_origin_rot; /* (n_joints, 9) row-major Rz(yaw) Ry(pitch) Rx(roll) of the URDF origin */
    const double *joint_axis;       /* (n_joints, 3) unit axis */
    const int32_t *body_joint;      /* [n_bodies] joint whose child link is the recorded body, -1 = the root link (data_convert.py:301-327) */
} amp_dataset_desc_t;
AMP_API int64_t amp_dataset_scratch_bytes(int32_t n_in, int32_t n_out, int32_t n_bodies);
/* data_convert.py:196-215 (interpolation) + :331-347 (FK).  dof_positions f64 (n_out, n_dofs); body_positions f32 (n_out, B, 3);
 * body_rotations f32 (n_out, B, 4) wxyz; root_pose f64 (n_out, 7) xyz + quat xyzw or NULL. */
AMP_API int amp_dataset_interp_fk(const amp_dataset_desc_t *d, double *dof_positions, float *body_positions, float *body_rotations,
                                  double *root_pose, void *scratch, int64_t scratch_bytes, void *stream);
/* data_convert.py:290-297 (joint velocities), :349-355 (body linear velocities), :357-371 (body angular velocities); each is the
 * raw difference followed by scipy gaussian_filter1d(sigma=1, axis=0) with reflect boundaries.  gauss_w host [5]: the kernel
 * weights for offsets 0..4 as scipy computes them.  The angular velocity evaluates compute_angular_velocity (:87-108) in
 * float64 on the float32 rotations (the reference's float32 evaluation is ill-conditioned, see csrc/amp_dataset.cu). */
AMP_API int amp_dataset_velocities(int32_t n_out, int32_t n_dofs, int32_t n_bodies, double dt, const double *gauss_w,
                                   const double *dof_positions, const float *body_positions, const float *body_rotations,
                                   double *dof_velocities, float *body_linear_velocities, float *body_angular_velocities,
                                   void *scratch, int64_t scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AMP_B200_H */
