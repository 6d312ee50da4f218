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
    // estimator (math_utils.py:30-52); rings are [window][12][N]
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
__global__ void qg_walk_kernel(QgWalkState W, QgWalkOpts o, float* __restrict__ obs, const float* __restrict__ ctrl,
                               const float4* __restrict__ S, const unsigned char* __restrict__ terminated,
                               float* __restrict__ terminal_obs, float* __restrict__ reward, float* __restrict__ terms,
                               double* __restrict__ reward64, double* __restrict__ terms64) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = W.n;
    if (e >= N) return;
    int flags = W.flags[e];
    // ---- compute_ideal_position (walking_quad.py:88-94)
    double* ip = W.ideal_position + 3 * (size_t)e;
    const double* gv = W.global_velocity + 3 * (size_t)e;
#pragma unroll
    for (int k = 0; k < 3; ++k) ip[k] += gv[k] * W.timestep * W.frame_skip;  // (gv*timestep)*frame_skip, left to right

    // ---- OnlineFrequencyAmplitudeEstimation.update(previous data.ctrl)  (math_utils.py:57-133)
    double c_new[12], c_prev[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) c_prev[k] = W.prev_ctrl[(size_t)e * 12 + k];
    if (ctrl) {
#pragma unroll
        for (int k = 0; k < 12; ++k) c_new[k] = (double)ctrl[(size_t)e * 12 + k];
    } else {  // data.ctrl of the batch (the step kernel ran without auto-reset, so this is the applied control)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            float4 c = S[(size_t)(QG_PL_LEG0 + 4 * l + 3) * N + e];
            c_new[3 * l] = c.x; c_new[3 * l + 1] = c.y; c_new[3 * l + 2] = c.z;
        }
    }
    int idx = W.buffer_index[e], sc = W.sample_count[e];
    const int win = W.window;
    if (!(flags & 4)) {
        for (int k = 0; k < 12; ++k) {
            W.prev_sample[(size_t)k * N + e] = c_prev[k];
            W.signal_ring[((size_t)idx * 12 + k) * N + e] = (float)c_prev[k];
        }
        sc = 1;
        idx = (idx + 1) % win;
        flags |= 4;
    } else {
        if (sc < win) sc++;
        const double dur = sc * W.dt;
        const int lim = sc < win ? sc : win;
        for (int k = 0; k < 12; ++k) {
            const size_t ke = (size_t)k * N + e;
            double diff = c_prev[k] - W.prev_sample[ke];
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
            W.signal_ring[ri] = (float)c_prev[k];
            W.prev_sample[ke] = c_prev[k];
            W.prev_sign[ke] = (signed char)sg;
            double f_cur = (cc / 2.0) / dur;
            W.f_est[ke] = W.ema_alpha * W.f_est[ke] + (1 - W.ema_alpha) * f_cur;
            float mx = -3.0e38f, mn = 3.0e38f;
            for (int w = 0; w < lim; ++w) {
                float v = W.signal_ring[((size_t)w * 12 + k) * N + e];
                mx = fmaxf(mx, v);
                mn = fminf(mn, v);
            }
            W.a_est[ke] = W.ema_alpha * W.a_est[ke] + (1 - W.ema_alpha) * ((double)mx - (double)mn);
        }
        idx = (idx + 1) % win;
        flags |= 8;
    }
    W.buffer_index[e] = idx;
    W.sample_count[e] = sc;

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

    // ---- reset() bookkeeping of WalkingQuadrupedEnv (:96-126) for terminated environments; the returned
    //      observation becomes the reset observation (zeros), the last one is kept as terminal observation
    const bool term = terminated && terminated[e];
    if (terminal_obs)
        for (int k = 0; k < 33; ++k) terminal_obs[(size_t)e * 33 + k] = term ? obs[(size_t)e * 33 + k] : 0.f;
    if (o.auto_reset && term) {
        for (int k = 0; k < 33; ++k) obs[(size_t)e * 33 + k] = 0.f;
        ip[0] = ip[1] = ip[2] = 0.0;
        for (int k = 0; k < 12; ++k) W.prev_ctrl[(size_t)e * 12 + k] = (double)o.joint_centers[k];
        flags &= ~1;  // previous_rewards_to_derive = None ; previous_ctrl_cost and the estimator survive
        int ep = ++W.episode[e];
        if (o.random_controls) walk_sample_commands(W, o, e, ep);
    }
    W.flags[e] = flags;
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

// After the physics + walking launches of one step (and BEFORE the masked physics reset):
//   sens = sensordata of this step (terminal one for terminated envs), S = state planes (post-step, pre-reset),
//   stacked [N, 26*window] is shifted by one frame and the new frame appended; terminated envs get their
//   terminal stack copied to terminal_stacked and are re-filled with the reset frame (po_walking_quad.py:59-70).
__global__ void qg_po_kernel(QgPoState P, QgWalkState W, QgWalkOpts o, const float* __restrict__ sens, const float4* __restrict__ S,
                             const unsigned char* __restrict__ terminated, float* __restrict__ stacked,
                             float* __restrict__ terminal_stacked, int auto_reset, int is_reset_call) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = P.n;
    if (e >= N) return;
    const int F = QG_PO_FRAME, Wn = P.window;
    float* st = stacked + (size_t)e * F * Wn;
    double* q = P.q + 4 * (size_t)e;
    const float* s = sens + (size_t)e * 33;
    float4 tq = S[(size_t)QG_PL_TIME * N + e];
    double time = __hiloint2double(__float_as_int(tq.y), __float_as_int(tq.x));
    float4 bq = S[(size_t)QG_PL_QUAT * N + e];
    double ctrl[12];
    for (int l = 0; l < 4; ++l) {
        float4 c = S[(size_t)(QG_PL_LEG0 + 4 * l + 3) * N + e];
        ctrl[3 * l] = c.x; ctrl[3 * l + 1] = c.y; ctrl[3 * l + 2] = c.z;
    }
    float frame[QG_PO_FRAME];
    if (!is_reset_call) {
        if (P.is_view[e]) { q[0] = bq.x; q[1] = bq.y; q[2] = bq.z; q[3] = bq.w; }
        if (time > P.settle_half) {
            double g[3] = {s[15], s[16], s[17]}, a[3] = {s[12], s[13], s[14]};
            madgwick_update_imu(q, g, a, P.Dt, P.beta);
            P.is_view[e] = 0;   // updateIMU returns a new array: the alias to qpos is gone
        }
        po_frame(P, W, e, s, ctrl, q, frame);
        for (int k = 0; k < F * (Wn - 1); ++k) st[k] = st[k + F];
        for (int k = 0; k < F; ++k) st[F * (Wn - 1) + k] = frame[k];
    }
    const bool term = is_reset_call ? (!terminated || terminated[e]) : (terminated && terminated[e]);
    if (!is_reset_call && terminal_stacked)
        for (int k = 0; k < F * Wn; ++k) terminal_stacked[(size_t)e * F * Wn + k] = term ? st[k] : 0.f;
    if (term && (auto_reset || is_reset_call)) {
        // reset(): sensordata zero, ctrl = joint centres, orientation still the stale filter state, commands not yet resampled
        float zero[33];
        for (int k = 0; k < 33; ++k) zero[k] = 0.f;
        double c0[12];
        for (int k = 0; k < 12; ++k) c0[k] = (double)o.joint_centers[k];
        po_frame(P, W, e, zero, c0, q, frame);
        for (int w = 0; w < Wn; ++w)
            for (int k = 0; k < F; ++k) st[w * F + k] = frame[k];
        P.is_view[e] = 1;   // computed_orientation = data.qpos[3:7]
    }
}
