/*
 * quadgym.h -- C ABI of libquadgym.so: the B200 (sm_100a) batched replacement for the physics +
 * epilogue of QuadrupedEnv.step() in antopio26/quadruped-gym.
 *
 * The reference has no native FFI of its own: its hot path sits behind five `mujoco` entry points
 * and the Gymnasium reset/step methods.  Each entry point below names the reference interface it
 * replaces (paths are /root/reference/...):
 *
 *   qg_model_load      <- mujoco.MjModel.from_xml_path            src/envs/quadruped.py:59
 *   qg_batch_create    <- mujoco.MjData(model)                     src/envs/quadruped.py:60
 *   qg_reset           <- mujoco.mj_resetData + ctrl default       src/envs/quadruped.py:115-139
 *                         (+ random yaw, src/envs/walking_quad.py:68-75)
 *   qg_step            <- clip, frame_skip x mujoco.mj_step,       src/envs/quadruped.py:153-182
 *                         sensordata copy, reward_fns sum,
 *                         termination_fns any
 *   qg_get_state /     <- env.data.{qpos,qvel,act,ctrl,time,       src/envs/quadruped.py:164,
 *   qg_set_state          qacc_warmstart} attribute access         src/envs/walking_quad.py:74,136
 *   qg_set_reward_table<- env.reward_fns dict                      src/envs/quadruped.py:97,170-175
 *   qg_set_options     <- max_time / termination_fns               src/envs/quadruped.py:98-100,149-151
 *                                                                  src/envs/walking_quad.py:152-162
 *
 * Conventions: every function returns 0 on success or a negative QG_E* code and never throws;
 * qg_last_error() gives a thread-local message.  All *_dev pointers are device memory owned by the
 * caller (torch tensors); the library owns only its internal state planes.  All work is enqueued
 * on the caller's stream (`stream` is a cudaStream_t passed as void*; NULL = legacy default
 * stream) with no hidden synchronisation, except the *_host entry points, which synchronise the
 * stream before returning.  A batch handle is not thread-safe; use one handle per GPU / process.
 * There is no CPU fallback: without a CUDA device every batch call fails with QG_ECUDA.
 */
#ifndef QUADGYM_H
#define QUADGYM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QG_OK 0
#define QG_EINVAL (-1)  /* bad argument */
#define QG_EBLOB (-2)   /* malformed or truncated model blob */
#define QG_EMODEL (-3)  /* model outside the supported class (free base + 4 legs x 3 z-hinges ...) */
#define QG_ECUDA (-4)   /* CUDA runtime error (no device, launch failure, out of memory) */
#define QG_ENOMEM (-5)

#define QG_NQ 19
#define QG_NV 18
#define QG_NU 12
#define QG_NSENSORDATA 33
#define QG_MAX_TERMS 16

/* fused reward terms (qg_set_reward_table). "sensor" = value in this step's sensordata (lags the
 * state by one physics step exactly as in the reference), "state" = mjData after the step. */
enum {
    QG_TERM_ALIVE = 0,          /* 1.0                                   walking_quad.py:286-290 */
    QG_TERM_CTRL_SQ = 1,        /* sum(ctrl^2)                           README.md:68-69 */
    QG_TERM_QVEL_X = 2,         /* state qvel[0]                         README.md:65-66 */
    QG_TERM_FORWARD = 3,        /* sensor linvel_x * pos_x               dummy_walking_quad.py:11-13 */
    QG_TERM_DRIFT = 4,          /* |sensor linvel_y * pos_y|             dummy_walking_quad.py:15-17 */
    QG_TERM_CONTROL_COST = 5,   /* alpha*first_cost+(1-alpha)*|dctrl|^2  walking_quad.py:255-270 (param = alpha) */
    QG_TERM_ORIENTATION = 6,    /* sensor zaxis_z                        walking_quad.py:237-241 */
    QG_TERM_HEIGHT_COST = 7,    /* |sensor pos_z - param|                walking_quad.py:243-247 */
    QG_TERM_POSTURE_COST = 8,   /* ||(ctrl - centres)/nu||               walking_quad.py:249-253 */
    QG_TERM_EXP_ORIENTATION = 9,/* exp(zaxis_z) - 1                      walking_quad.py:368, math_utils.py:4-5 */
    QG_TERM_EXP_HEIGHT = 10,    /* exp(|pos_z - param|) - 1              walking_quad.py:369 */
    QG_NUM_TERMS = 11
};

typedef struct qg_model qg_model;
typedef struct qg_batch qg_batch;

/* Sums over all environments and physics steps since the last qg_get_counters(reset=1). */
typedef struct qg_counters {
    unsigned long long physics_steps;   /* env * substeps */
    unsigned long long contacts;        /* active contacts */
    unsigned long long efc_rows;        /* constraint rows (limits + 4 per contact) */
    unsigned long long newton_iters;
    unsigned long long ls_evals;        /* line-search cost evaluations */
    unsigned long long verts_tested;    /* hull vertices visited by the support search */
    unsigned long long diverged;        /* envs reset by the non-finite / >1e10 guard */
    unsigned long long contact_overflow;/* contacts dropped because a lane's table was full */
    unsigned long long episodes;        /* terminations seen by qg_step */
    unsigned long long active_rows;     /* rows with non-zero force at the solver's final point */
} qg_counters;

const char* qg_last_error(void);
const char* qg_version(void);

/* --- model ------------------------------------------------------------------------------- */
int qg_model_load(const void* blob, size_t nbytes, qg_model** out);
void qg_model_destroy(qg_model* m);
/* sizes[8] = nq nv nu nbody njnt ngeom nmesh nsensordata */
int qg_model_info(const qg_model* m, int* sizes, double* timestep);

/* --- batch of environments ---------------------------------------------------------------- */
int qg_batch_create(const qg_model* m, int n_envs, int device, qg_batch** out);
void qg_batch_destroy(qg_batch* b);
int qg_batch_num_envs(const qg_batch* b);

/* max_time: episode limit in seconds (time >= max_time -> terminated, quadruped.py:151);
 * flip_termination: also terminate when sensordata[29] < 0 (walking_quad.py:152-156);
 * auto_reset: reset terminated envs inside qg_step (obs_dev then holds the reset observation,
 *             i.e. zeros, and terminal_obs_dev the last one);
 * solver_iterations / ls_iterations <= 0 keep the model's values capped at the device defaults. */
int qg_set_options(qg_batch* b, double max_time, int flip_termination, int auto_reset,
                   int solver_iterations, int ls_iterations);

/* reward = sum_k weights[k] * term(term_ids[k]; params[k]); terms_dev (if given) receives the
 * weighted values [N, n_terms].  n_terms = 0 gives the reference's default reward 0.0
 * (quadruped.py:145-147). */
int qg_set_reward_table(qg_batch* b, int n_terms, const int* term_ids, const double* weights,
                        const double* params);

/* mask_dev: [N] uint8 (NULL = all).  random_yaw != 0 draws the base yaw ~ U(0, 2pi) from a
 * counter-based generator keyed on (seed, global env id = env_offset + i, episode index). */
int qg_reset(qg_batch* b, const uint8_t* mask_dev, uint64_t seed, int random_yaw,
             long long env_offset, void* stream);

/* action_dev [N,12] f32; obs_dev [N,33] f32; reward_dev [N] f32; terms_dev [N,n_terms] f32 or
 * NULL; terminated_dev [N] u8; terminal_obs_dev [N,33] f32 or NULL.  One launch runs the whole
 * frame_skip loop and the sensor / reward / termination / auto-reset epilogue. */
int qg_step(qg_batch* b, const float* action_dev, int frame_skip, float* obs_dev, float* reward_dev,
            float* terms_dev, uint8_t* terminated_dev, float* terminal_obs_dev, void* stream);

/* Same call with HOST buffers -- the end-to-end path.  The batch is cut into segments of contiguous environments
 * (4 for >= 32,768 envs, 2 for >= 8,192, else 1), each on its own internal stream: H2D of the segment's actions, the
 * step kernel over the segment, D2H of its obs / reward / terminated (and, if the pointers are given, of its reward terms
 * [N,n_terms] and terminal observations [N,33]) straight from / into the caller's buffers (page-locked buffers are DMA-ed
 * without staging), so the copies of one segment run under the kernels of the others.  Results are identical to
 * qg_step's.  qg_step_host_async orders the work after the caller's earlier work on `stream` and returns without
 * waiting; qg_host_wait blocks until the outputs are in the host buffers; qg_step_host = both. */
int qg_step_host_async(qg_batch* b, const float* action_host, int frame_skip, float* obs_host, float* reward_host,
                       float* terms_host, uint8_t* terminated_host, float* terminal_obs_host, void* stream);
int qg_host_wait(qg_batch* b, void* stream);
int qg_step_host(qg_batch* b, const float* action_host, int frame_skip, float* obs_host, float* reward_host,
                 float* terms_host, uint8_t* terminated_host, float* terminal_obs_host, void* stream);

/* state in MuJoCo's conventions: qpos [N,19], qvel [N,18], act [N,12], qacc_warmstart [N,18],
 * time [N] f64, ctrl [N,12]; any pointer may be NULL to skip that field. */
int qg_get_state(qg_batch* b, float* qpos_dev, float* qvel_dev, float* act_dev, float* warm_dev,
                 double* time_dev, float* ctrl_dev, void* stream);
int qg_set_state(qg_batch* b, const float* qpos_dev, const float* qvel_dev, const float* act_dev,
                 const float* warm_dev, const double* time_dev, const float* ctrl_dev, void* stream);

/* One physics step with stage outputs for parity tests (any output pointer may be NULL):
 * qacc [N,18] and qacc_smooth [N,18] in MuJoCo's convention (base linear part in the world
 * frame); qfrc_bias [N,18] and M [N,18,18] in the kernel's "B form" (the three base-linear
 * dofs expressed in the base body frame: M_B = T^T M T, T = diag(R_base, I_15));
 * counts [N,4] = ncon, nefc, newton iterations, line-search evaluations; sensordata [N,33].
 * The state advances exactly as in qg_step with frame_skip = 1 and ctrl_dev [N,12] written to
 * data.ctrl unclipped (mj_step semantics, not env.step).  Synchronises the stream. */
int qg_debug_step(qg_batch* b, const float* ctrl_dev, float* qacc_dev, float* qacc_smooth_dev,
                  float* qfrc_bias_dev, float* M_dev, int* counts_dev, float* sensordata_dev,
                  void* stream);

int qg_get_counters(qg_batch* b, qg_counters* out_host, int reset, void* stream);

/* --- WalkingQuadrupedEnv reward stack (src/envs/walking_quad.py:9-428) -------------------------------------
 * qg_walk_enable allocates the per-env walking state: command inputs (control_inputs.py:9-12), ideal position
 * (walking_quad.py:88-94), control-cost memory (:255-270), derivative memory (:388-396) and the online
 * frequency / amplitude estimator (math_utils.py:11-133; `window` = ceil(2/(min_freq*dt)) samples x 12 channels).
 * dt = timestep*frame_skip as ONE float64 product (how the reference computes it); settling_time masks actions to
 * the joint centres while data.time < settling_time (walking_quad.py:142-143).  sample_opts = {min_speed,
 * max_speed, fixed_heading_angle, fixed_velocity_angle, fixed_speed}, sample_has = which of the three fixed_*
 * are given (control_inputs.py:88-116); random_controls resamples the command at every reset (:121-122). */
#define QG_WALK_NTERMS 11   /* WalkingQuadrupedEnv.reward_keys, walking_quad.py:331-350 */
int qg_walk_enable(qg_batch* b, int window, double dt, double timestep, int frame_skip, double settling_time,
                   int random_controls, const double* sample_opts, const int* sample_has);
/* options of control_inputs.sample for the NEXT resets (reset(options=...), walking_quad.py:100-103,121-122):
 * same layout as in qg_walk_enable; NULL pointers = the sampler's defaults (speed U(0,1), angles U(-pi,pi)). */
int qg_walk_set_sample_options(qg_batch* b, const double* sample_opts, const int* sample_has);
/* reset() bookkeeping of WalkingQuadrupedEnv (walking_quad.py:96-126); the command sampler is keyed on
 * (seed, env_offset + env, episode) from THIS call; hard != 0 also clears the state that survives reset() in the
 * reference (estimator, first control cost) and restarts the episode counters. */
int qg_walk_reset(qg_batch* b, const uint8_t* mask_dev, int hard, uint64_t seed, long long env_offset, void* stream);
/* set_velocity_speed_alpha + set_orientation (control_inputs.py:36-51): [N,3] = speed, alpha, theta (float64) */
int qg_walk_set_commands(qg_batch* b, const double* speed_alpha_theta_dev, const uint8_t* mask_dev, void* stream);
/* any pointer may be NULL: velocity/heading/global_velocity/ideal_position [N,3], f_est/a_est [12,N] (float64) */
int qg_walk_get_commands(qg_batch* b, double* velocity_dev, double* heading_dev, double* global_velocity_dev,
                         double* ideal_position_dev, double* f_est_dev, double* a_est_dev, void* stream);
/* One WalkingQuadrupedEnv.step() of bookkeeping after the physics launch (run qg_step with auto_reset = 0):
 * ideal position, estimator.update(previous ctrl), the 11 reward terms and their sum (float64 arithmetic;
 * float32 and/or float64 outputs), then for terminated envs the walking reset() bookkeeping, the zero reset
 * observation and the terminal observation.  ctrl_dev NULL = the batch's data.ctrl.  The caller then resets the
 * physics of the terminated envs with qg_reset(mask = terminated_dev). */
int qg_walk_step(qg_batch* b, float* obs_dev, const float* ctrl_dev, const uint8_t* terminated_dev,
                 float* terminal_obs_dev, float* reward_dev, float* terms_dev, double* reward64_dev,
                 double* terms64_dev, int auto_reset, void* stream);

/* --- POWalkingQuadrupedEnv observation (src/envs/po_walking_quad.py:8-90) -----------------------------------
 * 26 values per frame (gyro, accel, Madgwick-filter Euler angles, body_vel xy, ctrl, command velocity xy, heading
 * angle), FIFO-stacked over obs_window frames -> stacked_dev [N, 26*obs_window] f32, oldest frame first.
 * Dt = timestep*frame_skip (po_walking_quad.py:18), beta = Madgwick IMU gain (ahrs default 0.033).
 * Call order of one step: qg_step, then qg_po_observe(is_reset_call = 0), then qg_walk_step, then the masked qg_reset
 * (qg_walk_step zeroes the observation of terminated envs and resamples their commands, so the frame must be built
 * before it, exactly as the reference builds it inside QuadrupedEnv.step, quadruped.py:167);
 * sensordata_dev is that step's sensordata (the terminal one for terminated envs).  is_reset_call = 1 fills the
 * stack of the masked envs (terminated_dev as mask, NULL = all) with the reset frame (po_walking_quad.py:59-70). */
#define QG_PO_FRAME 26
int qg_po_enable(qg_batch* b, int obs_window, double Dt, double beta, double settling_time);
int qg_po_observe(qg_batch* b, const float* sensordata_dev, const uint8_t* terminated_dev, float* stacked_dev,
                  float* terminal_stacked_dev, int auto_reset, int is_reset_call, void* stream);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
unsigned long long qg_launch_count(void);

/* FP32 FFMA peak microbenchmark (roofline denominator): runs `iters` dependent-chain
 * FFMA rounds on every SM and returns the achieved TFLOP/s. */
int qg_fp32_peak(int device, int iters, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif
