// qg_walk.cuh -- WalkingQuadrupedEnv's reward stack on the device (one thread per environment, float64).
//
// Restates /root/reference/src/envs/walking_quad.py:128-148 (step order), :166-290 (terms), :352-422
// (input_control_reward), /root/reference/src/envs/math_utils.py:4-8,11-133 (exp_dist, unit, online
// frequency / amplitude estimator) and /root/reference/src/envs/control_inputs.py:14-116 (command state and
// sampler).  All arithmetic is float64 in the reference's operation order, on the float32 sensordata / ctrl the
// step kernel produced, so per-term values match numpy to the last bits (exp / sqrt / BLAS dot aside).
#pragma once
#include "qg_kernels.cuh"

#define QG_WALK_NTERMS 11

struct QgWalkState {
    int n, window;
    double dt;              // timestep * frame_skip (one Python float product, as the reference computes it)
    double timestep;
    int frame_skip;
    double ema_alpha;       // 0.8 (walking_quad.py:57)
    // command inputs (control_inputs.py:9-12), [N,3] each
    double *velocity, *heading, *global_velocity;
    double* ideal_position;      // [N,3]
    double* prev_derive;         // [N]  previous_rewards_to_derive
    double* first_ctrl_cost;     // [N]  previous_ctrl_cost (set once, never updated)
    double* prev_ctrl;           // [N,12]
    // estimator (math_utils.py:30-52); rings are [window][12][N].  The amplitude needs max - min over the window: the ring
    // is cut into blocks of `blk` samples whose extrema are kept ([nblk][12][N]); a step rescans only the block it writes
    // into and combines the block extrema (2 sqrt(window) loads per channel instead of `window`; exact).
    int blk, nblk;
    float *blk_max, *blk_min;
    float* signal_ring;
    unsigned char* cross_ring;
    int* cross_count;            // [12][N]
    double *prev_sample, *f_est, *a_est;  // [12][N]
    signed char* prev_sign;      // [12][N]
    int *buffer_index, *sample_count;     // [N]
    int* flags;                  // [N] bit0 prev_derive set, bit1 first_ctrl_cost set, bit2 prev_sample set, bit3 prev_sign set
    int* episode;                // [N]
};

struct QgWalkOpts {
    int random_controls, auto_reset;
    unsigned long long seed;
    long long env_offset;
    // control_inputs.sample options (control_inputs.py:88-92); has_* = "fixed" value given
    double min_speed, max_speed, fixed_heading, fixed_velocity_angle, fixed_speed;
    int has_heading, has_velocity_angle, has_speed;
    float joint_centers[12];
};

DI double np_norm2(double a, double b) { return sqrt(a * a + b * b); }

__device__ void walk_sample_commands(const QgWalkState& W, const QgWalkOpts& o, int e, int episode) {
    unsigned long long gid = (unsigned long long)(o.env_offset + e);
    uint4 r = philox4x32(make_uint2((unsigned)o.seed, (unsigned)(o.seed >> 32)),
                         make_uint4((unsigned)gid, (unsigned)(gid >> 32), (unsigned)episode, 0x57414c4bu));
    const double PI = 3.141592653589793;
    double u0 = (r.x + 0.5) / 4294967296.0, u1 = (r.y + 0.5) / 4294967296.0, u2 = (r.z + 0.5) / 4294967296.0;
    double theta = o.has_heading ? o.fixed_heading : (-PI + 2 * PI * u0);
    double alpha = o.has_velocity_angle ? o.fixed_velocity_angle : (-PI + 2 * PI * u1);
    double speed = o.has_speed ? o.fixed_speed : (o.min_speed + (o.max_speed - o.min_speed) * u2);
    double* v = W.velocity + 3 * (size_t)e;
    double* hd = W.heading + 3 * (size_t)e;
    double* gv = W.global_velocity + 3 * (size_t)e;
    hd[0] = cos(theta); hd[1] = sin(theta);                       // set_orientation
    v[0] = speed * cos(alpha); v[1] = speed * sin(alpha);         // set_velocity_speed_alpha
    gv[0] = hd[0] * v[0] - hd[1] * v[1];                          // update_global_velocity
    gv[1] = hd[1] * v[0] + hd[0] * v[1];
    gv[2] = 0.0;
}

// One WalkingQuadrupedEnv.step() worth of bookkeeping, run AFTER the physics launch:
//   ideal position += global_velocity*timestep*frame_skip (:133), estimator.update(previous ctrl) (:136),
//   reward = input_control_reward() on the new sensordata / ctrl (:352-422), then reset() bookkeeping
//   (:96-126) for terminated environments when auto_reset is on.
// Thread layout: a block handles 32 environments with 12 x 32 threads.  Phase 1: thread (channel k, env) updates the
// estimator of its channel (every array is [.][12][N]: a warp touches 32 consecutive environments of one channel).
// Phase 2: the 32 threads of channel 0 evaluate the reward terms of their environment in float64, in the reference's
// order of operations.
#define QG_WALK_ENVS_PER_BLOCK 32
__global__ void __launch_bounds__(12 * QG_WALK_ENVS_PER_BLOCK)
qg_walk_kernel(QgWalkState W, QgWalkOpts o, float* __restrict__ obs, const float* __restrict__ ctrl,
               const float4* __restrict__ S, const unsigned char* __restrict__ terminated,
               float* __restrict__ terminal_obs, float* __restrict__ reward, float* __restrict__ terms,
               double* __restrict__ reward64, double* __restrict__ terms64) {
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int e = blockIdx.x * QG_WALK_ENVS_PER_BLOCK + lane;
    const int N = W.n;
    const bool live = e < N;
    int flags = 0, idx = 0, sc = 0;
    const int win = W.window;
    if (live) {
        flags = W.flags[e];
        idx = W.buffer_index[e];
        sc = W.sample_count[e];
        // ---- OnlineFrequencyAmplitudeEstimation.update(previous data.ctrl), channel k  (math_utils.py:57-133)
        const double c_prev = W.prev_ctrl[(size_t)e * 12 + k];
        const size_t ke = (size_t)k * N + e;
        if (!(flags & 4)) {     // very first call: store the sample, no estimate yet
            W.prev_sample[ke] = c_prev;
            W.signal_ring[((size_t)idx * 12 + k) * N + e] = (float)c_prev;
            W.blk_max[((size_t)(idx / W.blk) * 12 + k) * N + e] = (float)c_prev;
            W.blk_min[((size_t)(idx / W.blk) * 12 + k) * N + e] = (float)c_prev;
        } else {
            const int scn = sc < win ? sc + 1 : sc;
            const double dur = scn * W.dt;
            const int lim = scn < win ? scn : win;
            double diff = c_prev - W.prev_sample[ke];
            int sg = (diff > 0.0) - (diff < 0.0);
            int crossing = 0;
            if (flags & 8) {
                int ps = W.prev_sign[ke];
                if (sg == 0) sg = ps;
                crossing = (sg != ps) ? 1 : 0;
            }
            const size_t ri = ((size_t)idx * 12 + k) * N + e;
            int cc = W.cross_count[ke] - (int)W.cross_ring[ri] + crossing;
            W.cross_ring[ri] = (unsigned char)crossing;
            W.cross_count[ke] = cc;
            W.signal_ring[ri] = (float)c_prev;
            W.prev_sample[ke] = c_prev;
            W.prev_sign[ke] = (signed char)sg;
            double f_cur = (cc / 2.0) / dur;
            W.f_est[ke] = W.ema_alpha * W.f_est[ke] + (1 - W.ema_alpha) * f_cur;
            // amplitude = max - min over the filled part of the ring: rescan the block just written, then the block extrema
            const int b = idx / W.blk, lo = b * W.blk;
            int hi = lo + W.blk;
            hi = hi < lim ? hi : lim;
            float mx = -3.0e38f, mn = 3.0e38f;
            for (int w = lo; w < hi; ++w) {
                float v = W.signal_ring[((size_t)w * 12 + k) * N + e];
                mx = fmaxf(mx, v);
                mn = fminf(mn, v);
            }
            W.blk_max[((size_t)b * 12 + k) * N + e] = mx;
            W.blk_min[((size_t)b * 12 + k) * N + e] = mn;
            const int nvalid = (lim + W.blk - 1) / W.blk;
            for (int j = 0; j < nvalid; ++j) {
                if (j == b) continue;
                mx = fmaxf(mx, W.blk_max[((size_t)j * 12 + k) * N + e]);
                mn = fminf(mn, W.blk_min[((size_t)j * 12 + k) * N + e]);
            }
            W.a_est[ke] = W.ema_alpha * W.a_est[ke] + (1 - W.ema_alpha) * ((double)mx - (double)mn);
        }
    }
    __syncthreads();       // f_est / a_est of all 12 channels are written before the reward reads them
    __shared__ unsigned char s_term[QG_WALK_ENVS_PER_BLOCK];
    if (k == 0) s_term[lane] = (live && terminated && terminated[e]) ? 1 : 0;
    if (k == 0 && live) {
    if (!(flags & 4)) { sc = 1; flags |= 4; }
    else { if (sc < win) sc++; flags |= 8; }
    W.buffer_index[e] = (idx + 1) % win;
    W.sample_count[e] = sc;

    // ---- compute_ideal_position (walking_quad.py:88-94)
    double* ip = W.ideal_position + 3 * (size_t)e;
    const double* gv = W.global_velocity + 3 * (size_t)e;
#pragma unroll
    for (int j = 0; j < 3; ++j) ip[j] += gv[j] * W.timestep * W.frame_skip;  // (gv*timestep)*frame_skip, left to right

    double c_new[12], c_prev[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) c_prev[j] = W.prev_ctrl[(size_t)e * 12 + j];
    if (ctrl) {
#pragma unroll
        for (int j = 0; j < 12; ++j) c_new[j] = (double)ctrl[(size_t)e * 12 + j];
    } else {  // data.ctrl of the batch (the step kernel ran without auto-reset, so this is the applied control)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            float4 c = S[(size_t)(QG_PL_LEG0 + 4 * l + 3) * N + e];
            c_new[3 * l] = c.x; c_new[3 * l + 1] = c.y; c_new[3 * l + 2] = c.z;
        }
    }

    // ---- input_control_reward (walking_quad.py:352-422)
    const float* s = obs + (size_t)e * 33;
    const double* vel = W.velocity + 3 * (size_t)e;
    const double* hd = W.heading + 3 * (size_t)e;
    double v[QG_WALK_NTERMS];
    v[0] = 10.0 * 1;                                                            // alive_bonus
    {                                                                           // control_cost (:255-270)
        double sq[12];
        for (int k = 0; k < 12; ++k) { double d = c_new[k] - c_prev[k]; sq[k] = d * d; }
        double cost = np_sum12(sq);
        if (!(flags & 2)) { W.first_ctrl_cost[e] = cost; flags |= 2; }
        v[1] = -2.0 * (0.8 * W.first_ctrl_cost[e] + (1 - 0.8) * cost);
    }
    {                                                                           // progress_direction_reward_local (:198-202)
        double bx = s[30], by = s[31], nb = np_norm2(bx, by), nc = np_norm2(vel[0], vel[1]);
        v[2] = 10.0 * ((bx / nb) * (vel[0] / nc) + (by / nb) * (vel[1] / nc));
        double d = nb - nc;                                                     // progress_speed_cost_local (:213-219)
        v[3] = -50.0 * (d * d);
    }
    v[4] = 10.0 * (exp((double)s[24] * hd[0] + (double)s[25] * hd[1]) - 1);     // heading_reward (:231-235)
    v[5] = 10.0 * (exp((double)s[29]) - 1);                                     // orientation_reward (:237-241)
    v[6] = -50.0 * (exp(fabs((double)s[20] - 0.13)) - 1);                       // body_height_cost(0.13) (:243-247)
    {
        double ss = 0.0;                                                        // joint_posture_cost (:249-253)
        for (int k = 0; k < 12; ++k) { double d = (c_new[k] - (double)o.joint_centers[k]) / 12; ss += d * d; }
        v[7] = -1.0 * sqrt(ss);
        double sa = 0.0, sf = 0.0;                                              // control_amplitude/frequency_cost (:272-284)
        for (int k = 0; k < 12; ++k) {
            const double ta = (k % 3 == 0) ? 1.5 : (k % 3 == 1 ? 0.5 : 0.0), tf = (k % 3 == 2) ? 0.0 : 1.0;
            double da = (W.a_est[(size_t)k * N + e] - ta) / 12, df = (W.f_est[(size_t)k * N + e] - tf) / 12;
            sa += da * da;
            sf += df * df;
        }
        v[8] = -2.5 * sqrt(sa);
        v[9] = -8.0 * sqrt(sf);
    }
    {                                                                           // d/dt(-20 * ideal_position_cost) (:388-396)
        double dx = (double)s[18] - ip[0], dy = (double)s[19] - ip[1];
        double r = -20.0 * np_norm2(dx, dy);
        double prev = (flags & 1) ? W.prev_derive[e] : r;
        v[10] = (r - prev) / W.dt;
        W.prev_derive[e] = r;
        flags |= 1;
    }
    double total = 0.0;
    for (int k = 0; k < QG_WALK_NTERMS; ++k) total += v[k];                     // Python sum(values): sequential
    if (reward) reward[e] = (float)total;
    if (reward64) reward64[e] = total;
    for (int k = 0; k < QG_WALK_NTERMS; ++k) {
        if (terms) terms[(size_t)e * QG_WALK_NTERMS + k] = (float)v[k];
        if (terms64) terms64[(size_t)e * QG_WALK_NTERMS + k] = v[k];
    }
    for (int k = 0; k < 12; ++k) W.prev_ctrl[(size_t)e * 12 + k] = c_new[k];    // control_cost's side effect

    // ---- reset() bookkeeping of WalkingQuadrupedEnv (:96-126) for terminated environments
    const bool term = terminated && terminated[e];
    if (o.auto_reset && term) {
        ip[0] = ip[1] = ip[2] = 0.0;
        for (int k = 0; k < 12; ++k) W.prev_ctrl[(size_t)e * 12 + k] = (double)o.joint_centers[k];
        flags &= ~1;  // previous_rewards_to_derive = None ; previous_ctrl_cost and the estimator survive
        int ep = ++W.episode[e];
        if (o.random_controls) walk_sample_commands(W, o, e, ep);
    }
    W.flags[e] = flags;
    }
    // ---- the returned observation of a terminated environment becomes the reset observation (zeros), the last one is
    //      kept as terminal observation (zeros elsewhere): the block's 32 rows are contiguous, all 384 threads move them
    __syncthreads();       // the reward threads are done reading obs
    const int e0 = blockIdx.x * QG_WALK_ENVS_PER_BLOCK;
    const int nrow = (N - e0 < QG_WALK_ENVS_PER_BLOCK) ? N - e0 : QG_WALK_ENVS_PER_BLOCK;
    for (int i = threadIdx.x; i < nrow * 33; i += blockDim.x) {
        const bool t = s_term[i / 33] != 0;
        const size_t g = (size_t)e0 * 33 + i;
        if (terminal_obs) terminal_obs[g] = t ? obs[g] : 0.f;
        if (t && o.auto_reset) obs[g] = 0.f;
    }
}

__global__ void qg_walk_reset_kernel(QgWalkState W, QgWalkOpts o, const unsigned char* __restrict__ mask, int hard) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= W.n) return;
    if (mask && !mask[e]) return;
    double* ip = W.ideal_position + 3 * (size_t)e;
    ip[0] = ip[1] = ip[2] = 0.0;
    for (int k = 0; k < 12; ++k) W.prev_ctrl[(size_t)e * 12 + k] = (double)o.joint_centers[k];
    int flags = hard ? 0 : (W.flags[e] & ~1);
    W.flags[e] = flags;
    int ep = hard ? 0 : W.episode[e] + 1;
    W.episode[e] = ep;
    if (o.random_controls) walk_sample_commands(W, o, e, ep);
}

__global__ void qg_walk_set_commands_kernel(QgWalkState W, const double* __restrict__ speed_alpha_theta,
                                            const unsigned char* __restrict__ mask) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= W.n) return;
    if (mask && !mask[e]) return;
    double speed = speed_alpha_theta[3 * (size_t)e], alpha = speed_alpha_theta[3 * (size_t)e + 1], theta = speed_alpha_theta[3 * (size_t)e + 2];
    double* v = W.velocity + 3 * (size_t)e;
    double* hd = W.heading + 3 * (size_t)e;
    double* gv = W.global_velocity + 3 * (size_t)e;
    hd[0] = cos(theta); hd[1] = sin(theta);
    v[0] = speed * cos(alpha); v[1] = speed * sin(alpha);
    gv[0] = hd[0] * v[0] - hd[1] * v[1];
    gv[1] = hd[1] * v[0] + hd[0] * v[1];
    gv[2] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// POWalkingQuadrupedEnv observation (/root/reference/src/envs/po_walking_quad.py:29-90): Madgwick IMU filter ->
// Euler angles, 26 values per frame (gyro 3, accel 3, euler 3, body_vel xy 2, ctrl 12, command velocity xy 2,
// heading angle 1), FIFO-stacked over obs_window frames (oldest first).
// The filter lives in the third-party package `ahrs` (not vendored, not installable here): updateIMU and
// Quaternion.to_angles are restated from its published algorithm -- PARITY UNPINNED.
#define QG_PO_FRAME 26

struct QgPoState {
    int n, window;
    double Dt, beta, settle_half;   // timestep*frame_skip, Madgwick gain (0.033 for the IMU variant), settling_time/2
    double* q;                      // [N,4] computed_orientation
    int* is_view;                   // [N] computed_orientation aliases data.qpos[3:7] (po_walking_quad.py:68)
    float* ring;                    // [N][window][26] frames; slot head_ctr[0] is the oldest of every environment
    int* head_ctr;                  // [2] device-side ring head and blocks-done counter (no host state: graph-capture safe)
};

DI void madgwick_update_imu(double* q, const double* g, const double* a, double Dt, double beta) {
    double gn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    if (gn == 0.0) return;
    double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    // qDot = 0.5 * q (x) (0, gyr)
    double dw = 0.5 * (-qx * g[0] - qy * g[1] - qz * g[2]);
    double dx = 0.5 * (qw * g[0] + qy * g[2] - qz * g[1]);
    double dy = 0.5 * (qw * g[1] - qx * g[2] + qz * g[0]);
    double dz = 0.5 * (qw * g[2] + qx * g[1] - qy * g[0]);
    double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (an > 0.0) {
        double ax = a[0] / an, ay = a[1] / an, az = a[2] / an;
        double qn = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
        double w = qw / qn, x = qx / qn, y = qy / qn, z = qz / qn;
        double f0 = 2.0 * (x * z - w * y) - ax, f1 = 2.0 * (w * x + y * z) - ay, f2 = 2.0 * (0.5 - x * x - y * y) - az;
        if (sqrt(f0 * f0 + f1 * f1 + f2 * f2) > 0.0) {
            // gradient = J^T f
            double g0 = -2.0 * y * f0 + 2.0 * x * f1;
            double g1 = 2.0 * z * f0 + 2.0 * w * f1 - 4.0 * x * f2;
            double g2 = -2.0 * w * f0 + 2.0 * z * f1 - 4.0 * y * f2;
            double g3 = 2.0 * x * f0 + 2.0 * y * f1;
            double gnorm = sqrt(g0 * g0 + g1 * g1 + g2 * g2 + g3 * g3);
            dw -= beta * g0 / gnorm; dx -= beta * g1 / gnorm; dy -= beta * g2 / gnorm; dz -= beta * g3 / gnorm;
        }
    }
    qw += dw * Dt; qx += dx * Dt; qy += dy * Dt; qz += dz * Dt;
    double n = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    q[0] = qw / n; q[1] = qx / n; q[2] = qy / n; q[3] = qz / n;
}

DI void po_frame(const QgPoState& P, const QgWalkState& W, int e, const float* s, const double* ctrl, const double* q, float* out) {
    // Quaternion.to_angles(): roll, pitch, yaw
    double w = q[0], x = q[1], y = q[2], z = q[3];
    double roll = atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y));
    double sp = 2.0 * (w * y - z * x);
    double pitch = asin(fmin(1.0, fmax(-1.0, sp)));
    double yaw = atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
    out[0] = s[15]; out[1] = s[16]; out[2] = s[17];          // gyro
    out[3] = s[12]; out[4] = s[13]; out[5] = s[14];          // accel
    out[6] = (float)roll; out[7] = (float)pitch; out[8] = (float)yaw;
    out[9] = s[30]; out[10] = s[31];                         // body_vel xy (optical flow)
    for (int k = 0; k < 12; ++k) out[11 + k] = (float)ctrl[k];
    const double* v = W.velocity + 3 * (size_t)e;
    const double* hd = W.heading + 3 * (size_t)e;
    out[23] = (float)v[0]; out[24] = (float)v[1];
    out[25] = (float)atan2(hd[1], hd[0]);                    // get_heading_theta
}

// One POWalkingQuadrupedEnv observation per step (qg_step -> THIS -> qg_walk_step -> masked qg_reset):
//   sens = sensordata of this step (terminal one for terminated envs), S = state planes (post-step, pre-reset).
// The frames live in a ring [N][window][26] whose oldest slot is `head` for every environment (all environments push one
// frame per step; a reset overwrites all slots of that environment, which is independent of the head).  A block handles
// 32 environments: 32 threads run the filter and build the new (and, for finished environments, the reset) frame in
// shared memory, then ALL threads write the block's contiguous rows: the new ring slot, the terminal stack of finished
// environments, the reset fill, and the oldest-first stacked output [N, 26*window] (po_walking_quad.py:59-90) -- rows of
// consecutive environments are contiguous, so every access is coalesced; nothing is shifted in memory.
#define QG_PO_ENVS_PER_BLOCK 32
__global__ void __launch_bounds__(256)
qg_po_kernel(QgPoState P, QgWalkState W, QgWalkOpts o, const float* __restrict__ sens, const float4* __restrict__ S,
             const unsigned char* __restrict__ terminated, float* __restrict__ stacked, float* __restrict__ terminal_stacked,
             int auto_reset, int is_reset_call) {
    const int head = P.head_ctr[0];     // the same for every block: only the last block to finish advances it
    __shared__ __align__(8) float s_new[QG_PO_ENVS_PER_BLOCK][QG_PO_FRAME], s_rst[QG_PO_ENVS_PER_BLOCK][QG_PO_FRAME];
    __shared__ unsigned char s_term[QG_PO_ENVS_PER_BLOCK];
    const int N = P.n, Wn = P.window;
    const int e0 = blockIdx.x * QG_PO_ENVS_PER_BLOCK;
    const int nrow = (N - e0 < QG_PO_ENVS_PER_BLOCK) ? N - e0 : QG_PO_ENVS_PER_BLOCK;
    if (threadIdx.x < nrow) {
        const int e = e0 + threadIdx.x;
        double* q = P.q + 4 * (size_t)e;
        const bool term = is_reset_call ? (!terminated || terminated[e]) : (terminated && terminated[e]);
        s_term[threadIdx.x] = term ? 1 : 0;
        if (!is_reset_call) {
            const float* s = sens + (size_t)e * 33;
            float4 tq = S[(size_t)QG_PL_TIME * N + e];
            double time = __hiloint2double(__float_as_int(tq.y), __float_as_int(tq.x));
            float4 bq = S[(size_t)QG_PL_QUAT * N + e];
            double ctrl[12];
            for (int l = 0; l < 4; ++l) {
                float4 c = S[(size_t)(QG_PL_LEG0 + 4 * l + 3) * N + e];
                ctrl[3 * l] = c.x; ctrl[3 * l + 1] = c.y; ctrl[3 * l + 2] = c.z;
            }
            if (P.is_view[e]) { q[0] = bq.x; q[1] = bq.y; q[2] = bq.z; q[3] = bq.w; }
            if (time > P.settle_half) {
                double g[3] = {s[15], s[16], s[17]}, a[3] = {s[12], s[13], s[14]};
                madgwick_update_imu(q, g, a, P.Dt, P.beta);
                P.is_view[e] = 0;   // updateIMU returns a new array: the alias to qpos is gone
            }
            po_frame(P, W, e, s, ctrl, q, s_new[threadIdx.x]);
        }
        if (term && (auto_reset || is_reset_call)) {
            // reset(): sensordata zero, ctrl = joint centres, orientation still the stale filter state, commands not yet resampled
            float zero[33];
            for (int j = 0; j < 33; ++j) zero[j] = 0.f;
            double c0[12];
            for (int j = 0; j < 12; ++j) c0[j] = (double)o.joint_centers[j];
            po_frame(P, W, e, zero, c0, q, s_rst[threadIdx.x]);
            P.is_view[e] = 1;   // computed_orientation = data.qpos[3:7]
        }
    }
    __syncthreads();
    // All copies below move float2 (a frame is 13 of them, rows start 8-byte aligned) and walk the block's rows with
    // incremental (environment r, frame j, pair p) indices: no division per element.
    constexpr int FP = QG_PO_FRAME / 2;
    static_assert(QG_PO_FRAME % 2 == 0, "frames are moved as float2");
    const int row2 = FP * Wn, total2 = nrow * row2;
    float2* ring2 = reinterpret_cast<float2*>(P.ring) + (size_t)e0 * row2;
    float2* out2 = reinterpret_cast<float2*>(stacked) + (size_t)e0 * row2;
    float2* tout2 = terminal_stacked ? reinterpret_cast<float2*>(terminal_stacked) + (size_t)e0 * row2 : nullptr;
    const float2(*new2)[FP] = reinterpret_cast<const float2(*)[FP]>(s_new);
    const float2(*rst2)[FP] = reinterpret_cast<const float2(*)[FP]>(s_rst);
    const int head_new = is_reset_call ? head : (head + 1) % Wn;     // oldest slot after this push
    const int r0 = threadIdx.x / row2, m0 = threadIdx.x - r0 * row2, j0 = m0 / FP, p0 = m0 - j0 * FP;
    const int dr = (int)blockDim.x / row2, dm = (int)blockDim.x - dr * row2, dj = dm / FP, dp = dm - dj * FP;
    auto for_each = [&](auto&& body) {          // body(i, r, j, p): element i = pair p of logical frame j of environment r
        int r = r0, j = j0, p = p0;
        for (int i = threadIdx.x; i < total2; i += blockDim.x) {
            body(i, r, j, p);
            p += dp; j += dj; r += dr;
            if (p >= FP) { p -= FP; j++; }
            if (j >= Wn) { j -= Wn; r++; }
        }
    };
    // logical frame j (oldest first) of environment r after the push, BEFORE any reset fill
    auto pushed = [&](int r, int j, int p) -> float2 {
        if (!is_reset_call && j == Wn - 1) return new2[r][p];
        int slot = head_new + j;
        if (slot >= Wn) slot -= Wn;
        return ring2[(size_t)r * row2 + slot * FP + p];
    };
    if (!is_reset_call && tout2)      // terminal stacks of the finished environments only
        for_each([&](int i, int r, int j, int p) { if (s_term[r]) tout2[i] = pushed(r, j, p); });
    for_each([&](int i, int r, int j, int p) {
        const bool fill = s_term[r] && (auto_reset || is_reset_call);
        if (is_reset_call && !fill) return;           // reset() touches the masked environments only
        out2[i] = fill ? rst2[r][p] : pushed(r, j, p);
    });
    __syncthreads();       // every read of the old ring contents is done
    for_each([&](int i, int r, int j, int p) {       // here j is the PHYSICAL slot of element i
        const bool fill = s_term[r] && (auto_reset || is_reset_call);
        if (fill) ring2[i] = rst2[r][p];
        else if (!is_reset_call && j == head) ring2[i] = new2[r][p];
    });
    // the slot just written was the oldest: the last block to leave advances the head (every block has read it by then)
    if (!is_reset_call && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(P.head_ctr + 1, 1) == (int)gridDim.x - 1) {
            P.head_ctr[1] = 0;
            P.head_ctr[0] = head_new;
        }
    }
}
