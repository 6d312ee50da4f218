// qg_kernels.cuh -- global kernels: fused env.step (frame_skip loop + sensor/reward/termination/
// auto-reset epilogue), reset, state pack/unpack, FP32 peak microbenchmark.
//
// qg_step_kernel replaces the body of QuadrupedEnv.step (/root/reference/src/envs/quadruped.py:153-182):
//   :160 action clip, :163-165 frame_skip x mj_step, :167 sensordata copy, :170-175 reward sum,
//   :178 termination any, plus the SB3-style auto-reset that SubprocVecEnv performs around it
//   (/root/reference/src/train_quadruped.py:50) and mj_resetData + default ctrl (quadruped.py:120-124).
#pragma once
#ifndef QG_BLOCKSYNC
#define QG_BLOCKSYNC 5
#endif
#include "qg_step.cuh"

#ifndef QG_BLOCK
#define QG_BLOCK 256
#endif
#ifndef QG_MINBLOCKS
#define QG_MINBLOCKS 1
#endif
#ifndef QG_BLOCKSYNC
#define QG_BLOCKSYNC 5
#endif

struct QgCounters {
    unsigned long long physics_steps, contacts, efc_rows, newton_iters, ls_evals, verts_tested, diverged,
        contact_overflow, episodes, active_rows;
};

DI float4 ldS(const float4* S, int plane, int N, int env) { return S[(size_t)plane * N + env]; }
DI void stS(float4* S, int plane, int N, int env, float4 v) { S[(size_t)plane * N + env] = v; }

// Philox4x32-10 counter-based generator (per-env streams keyed on seed / env id / episode)
DI uint4 philox4x32(uint2 key, uint4 c) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return c;
}

DI void load_lane(const float4* __restrict__ S, int N, int env, int leg, LaneState& L, int& episode, double& first_cc,
                  int& flags) {
    float4 p = ldS(S, QG_PL_POS, N, env), q = ldS(S, QG_PL_QUAT, N, env), v = ldS(S, QG_PL_VLIN, N, env);
    float4 o = ldS(S, QG_PL_VANG, N, env), a = ldS(S, QG_PL_WLIN, N, env), b = ldS(S, QG_PL_WANG, N, env);
    float4 t = ldS(S, QG_PL_TIME, N, env), x = ldS(S, QG_PL_AUX, N, env);
    L.pb = V3(p.x, p.y, p.z);
    L.qw = q.x; L.qx = q.y; L.qy = q.z; L.qz = q.w;
    L.vw = V3(v.x, v.y, v.z);
    L.om = V3(o.x, o.y, o.z);
    L.wl = V3(a.x, a.y, a.z);
    L.wa = V3(b.x, b.y, b.z);
    L.time = __hiloint2double(__float_as_int(t.y), __float_as_int(t.x));
    episode = __float_as_int(t.z);
    flags = __float_as_int(x.x);
    first_cc = __hiloint2double(__float_as_int(x.z), __float_as_int(x.y));
    int pl = QG_PL_LEG0 + 4 * leg;
    float4 l0 = ldS(S, pl, N, env), l1 = ldS(S, pl + 1, N, env), l2 = ldS(S, pl + 2, N, env), l3 = ldS(S, pl + 3, N, env);
    L.q[0] = l0.x; L.q[1] = l0.y; L.q[2] = l0.z; L.qd[0] = l0.w;
    L.qd[1] = l1.x; L.qd[2] = l1.y; L.act[0] = l1.z; L.act[1] = l1.w;
    L.act[2] = l2.x; L.wj[0] = l2.y; L.wj[1] = l2.z; L.wj[2] = l2.w;
    L.ctrl[0] = l3.x; L.ctrl[1] = l3.y; L.ctrl[2] = l3.z;
}

DI void store_lane(float4* __restrict__ S, int N, int env, int leg, const LaneState& L, int episode, double first_cc,
                   int flags) {
    if (leg == 0) {
        stS(S, QG_PL_POS, N, env, make_float4(L.pb.x, L.pb.y, L.pb.z, 0.f));
        stS(S, QG_PL_QUAT, N, env, make_float4(L.qw, L.qx, L.qy, L.qz));
    } else if (leg == 1) {
        stS(S, QG_PL_VLIN, N, env, make_float4(L.vw.x, L.vw.y, L.vw.z, 0.f));
        stS(S, QG_PL_VANG, N, env, make_float4(L.om.x, L.om.y, L.om.z, 0.f));
    } else if (leg == 2) {
        stS(S, QG_PL_WLIN, N, env, make_float4(L.wl.x, L.wl.y, L.wl.z, 0.f));
        stS(S, QG_PL_WANG, N, env, make_float4(L.wa.x, L.wa.y, L.wa.z, 0.f));
    } else {
        stS(S, QG_PL_TIME, N, env, make_float4(__int_as_float(__double2loint(L.time)), __int_as_float(__double2hiint(L.time)),
                                                __int_as_float(episode), 0.f));
        stS(S, QG_PL_AUX, N, env, make_float4(__int_as_float(flags), __int_as_float(__double2loint(first_cc)),
                                               __int_as_float(__double2hiint(first_cc)), 0.f));
    }
    int pl = QG_PL_LEG0 + 4 * leg;
    stS(S, pl, N, env, make_float4(L.q[0], L.q[1], L.q[2], L.qd[0]));
    stS(S, pl + 1, N, env, make_float4(L.qd[1], L.qd[2], L.act[0], L.act[1]));
    stS(S, pl + 2, N, env, make_float4(L.act[2], L.wj[0], L.wj[1], L.wj[2]));
    stS(S, pl + 3, N, env, make_float4(L.ctrl[0], L.ctrl[1], L.ctrl[2], 0.f));
}

// mj_resetData + the env's default ctrl (quadruped.py:120-124); optional random yaw (walking_quad.py:68-75)
DI void reset_lane(const QgModelC& P, LaneState& L, int leg, const QgStepOpts& o, int env, int episode) {
    L.pb = V3(P.qpos0[0], P.qpos0[1], P.qpos0[2]);
    L.qw = P.qpos0[3]; L.qx = P.qpos0[4]; L.qy = P.qpos0[5]; L.qz = P.qpos0[6];
    if (o.random_yaw) {
        unsigned long long gid = (unsigned long long)(o.env_offset + env);
        uint4 r = philox4x32(make_uint2((unsigned)o.seed, (unsigned)(o.seed >> 32)),
                             make_uint4((unsigned)gid, (unsigned)(gid >> 32), (unsigned)episode, 0x59415721u));
        float ang = (r.x + 0.5f) * (6.283185307179586f / 4294967296.f);
        float sn, cs;
        sincosf(0.5f * ang, &sn, &cs);
        L.qw = cs; L.qx = 0.f; L.qy = 0.f; L.qz = sn;
    }
    L.vw = L.om = L.wl = L.wa = V3(0, 0, 0);
    L.time = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        L.q[k] = P.qpos0[7 + 3 * leg + k];
        L.qd[k] = L.act[k] = L.wj[k] = 0.f;
        L.ctrl[k] = o.reset_ctrl[3 * leg + k];
    }
}

// state planes (HBM) -> the block's shared-memory state: every lane moves its 4 leg planes, the 8 base planes are split
// over the quad's lanes (2 each); the caller syncs the quad before anyone reads another lane's words
DI void planes_to_shared(const float4* __restrict__ S, int N, int env, int leg, const StateRef& R) {
    const int pl = QG_PL_LEG0 + 4 * leg;
    const float4 l0 = ldS(S, pl, N, env), l1 = ldS(S, pl + 1, N, env), l2 = ldS(S, pl + 2, N, env), l3 = ldS(S, pl + 3, N, env);
    const float4 b0 = ldS(S, 2 * leg, N, env), b1 = ldS(S, 2 * leg + 1, N, env);
    SLW(R, SL_Q) = l0.x; SLW(R, SL_Q + 1) = l0.y; SLW(R, SL_Q + 2) = l0.z; SLW(R, SL_QD) = l0.w;
    SLW(R, SL_QD + 1) = l1.x; SLW(R, SL_QD + 2) = l1.y; SLW(R, SL_ACT) = l1.z; SLW(R, SL_ACT + 1) = l1.w;
    SLW(R, SL_ACT + 2) = l2.x; SLW(R, SL_WJ) = l2.y; SLW(R, SL_WJ + 1) = l2.z; SLW(R, SL_WJ + 2) = l2.w;
    SLW(R, SL_CTRL) = l3.x; SLW(R, SL_CTRL + 1) = l3.y; SLW(R, SL_CTRL + 2) = l3.z;
    SLW(R, SL_PCTRL) = l3.x; SLW(R, SL_PCTRL + 1) = l3.y; SLW(R, SL_PCTRL + 2) = l3.z;
    if (leg == 0) {          // QG_PL_POS, QG_PL_QUAT
        SBW(R, SB_POS) = b0.x; SBW(R, SB_POS + 1) = b0.y; SBW(R, SB_POS + 2) = b0.z;
        SBW(R, SB_QUAT) = b1.x; SBW(R, SB_QUAT + 1) = b1.y; SBW(R, SB_QUAT + 2) = b1.z; SBW(R, SB_QUAT + 3) = b1.w;
    } else if (leg == 1) {   // QG_PL_VLIN, QG_PL_VANG
        SBW(R, SB_VLIN) = b0.x; SBW(R, SB_VLIN + 1) = b0.y; SBW(R, SB_VLIN + 2) = b0.z;
        SBW(R, SB_VANG) = b1.x; SBW(R, SB_VANG + 1) = b1.y; SBW(R, SB_VANG + 2) = b1.z;
    } else if (leg == 2) {   // QG_PL_WLIN, QG_PL_WANG
        SBW(R, SB_WLIN) = b0.x; SBW(R, SB_WLIN + 1) = b0.y; SBW(R, SB_WLIN + 2) = b0.z;
        SBW(R, SB_WANG) = b1.x; SBW(R, SB_WANG + 1) = b1.y; SBW(R, SB_WANG + 2) = b1.z;
    } else {                 // QG_PL_TIME (time lo/hi, episode), QG_PL_AUX (flags, first control cost lo/hi)
        SBW(R, SB_TIME) = b0.x; SBW(R, SB_TIME + 1) = b0.y; SBW(R, SB_EPISODE) = b0.z;
        SBW(R, SB_FLAGS) = b1.x; SBW(R, SB_FCC) = b1.y; SBW(R, SB_FCC + 1) = b1.z;
    }
}

DI void shared_to_planes(float4* __restrict__ S, int N, int env, int leg, const StateRef& R) {
    const int pl = QG_PL_LEG0 + 4 * leg;
    stS(S, pl, N, env, make_float4(SLW(R, SL_Q), SLW(R, SL_Q + 1), SLW(R, SL_Q + 2), SLW(R, SL_QD)));
    stS(S, pl + 1, N, env, make_float4(SLW(R, SL_QD + 1), SLW(R, SL_QD + 2), SLW(R, SL_ACT), SLW(R, SL_ACT + 1)));
    stS(S, pl + 2, N, env, make_float4(SLW(R, SL_ACT + 2), SLW(R, SL_WJ), SLW(R, SL_WJ + 1), SLW(R, SL_WJ + 2)));
    stS(S, pl + 3, N, env, make_float4(SLW(R, SL_CTRL), SLW(R, SL_CTRL + 1), SLW(R, SL_CTRL + 2), 0.f));
    if (leg == 0) {
        stS(S, QG_PL_POS, N, env, make_float4(SBW(R, SB_POS), SBW(R, SB_POS + 1), SBW(R, SB_POS + 2), 0.f));
        stS(S, QG_PL_QUAT, N, env, make_float4(SBW(R, SB_QUAT), SBW(R, SB_QUAT + 1), SBW(R, SB_QUAT + 2), SBW(R, SB_QUAT + 3)));
    } else if (leg == 1) {
        stS(S, QG_PL_VLIN, N, env, make_float4(SBW(R, SB_VLIN), SBW(R, SB_VLIN + 1), SBW(R, SB_VLIN + 2), 0.f));
        stS(S, QG_PL_VANG, N, env, make_float4(SBW(R, SB_VANG), SBW(R, SB_VANG + 1), SBW(R, SB_VANG + 2), 0.f));
    } else if (leg == 2) {
        stS(S, QG_PL_WLIN, N, env, make_float4(SBW(R, SB_WLIN), SBW(R, SB_WLIN + 1), SBW(R, SB_WLIN + 2), 0.f));
        stS(S, QG_PL_WANG, N, env, make_float4(SBW(R, SB_WANG), SBW(R, SB_WANG + 1), SBW(R, SB_WANG + 2), 0.f));
    } else {
        stS(S, QG_PL_TIME, N, env, make_float4(SBW(R, SB_TIME), SBW(R, SB_TIME + 1), SBW(R, SB_EPISODE), 0.f));
        stS(S, QG_PL_AUX, N, env, make_float4(SBW(R, SB_FLAGS), SBW(R, SB_FCC), SBW(R, SB_FCC + 1), 0.f));
    }
}

// reset_lane into the shared state (lane `leg` writes its own leg words, lane 0 the base words); keep_ctrl: the
// mid-step blow-up guard keeps data.ctrl
DI void reset_shared(const QgModelC& P, const StateRef& R, int leg, const QgStepOpts& o, int env, int episode, bool keep_ctrl) {
    LaneState L;
    reset_lane(P, L, leg, o, env, episode);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        SLW(R, SL_Q + k) = L.q[k]; SLW(R, SL_QD + k) = 0.f; SLW(R, SL_ACT + k) = 0.f; SLW(R, SL_WJ + k) = 0.f;
        if (!keep_ctrl) SLW(R, SL_CTRL + k) = L.ctrl[k];
    }
    if (leg == 0) {
        SBW(R, SB_POS) = L.pb.x; SBW(R, SB_POS + 1) = L.pb.y; SBW(R, SB_POS + 2) = L.pb.z;
        SBW(R, SB_QUAT) = L.qw; SBW(R, SB_QUAT + 1) = L.qx; SBW(R, SB_QUAT + 2) = L.qy; SBW(R, SB_QUAT + 3) = L.qz;
#pragma unroll
        for (int k = 0; k < 12; ++k) SBW(R, SB_VLIN + k) = 0.f;     // vlin, vang, wlin, wang
        sb_set_time(R, 0.0);
    }
}

DI bool shared_bad(const StateRef& R) {
    bool ok = true;
#pragma unroll
    for (int w = SB_POS; w < SB_WLIN; ++w) ok = ok && fabsf(SBW(R, w)) < 1e10f;
#pragma unroll
    for (int k = 0; k < 6; ++k) ok = ok && fabsf(SLW(R, SL_Q + k)) < 1e10f;     // q, qd
    return !ok;
}

DI bool lane_bad(const LaneState& L) {
    bool ok = fabsf(L.pb.x) < 1e10f && fabsf(L.pb.y) < 1e10f && fabsf(L.pb.z) < 1e10f && fabsf(L.qw) < 1e10f &&
              fabsf(L.qx) < 1e10f && fabsf(L.qy) < 1e10f && fabsf(L.qz) < 1e10f && fabsf(L.vw.x) < 1e10f &&
              fabsf(L.vw.y) < 1e10f && fabsf(L.vw.z) < 1e10f && fabsf(L.om.x) < 1e10f && fabsf(L.om.y) < 1e10f &&
              fabsf(L.om.z) < 1e10f;
#pragma unroll
    for (int k = 0; k < 3; ++k) ok = ok && fabsf(L.q[k]) < 1e10f && fabsf(L.qd[k]) < 1e10f;
    return !ok;
}

// numpy's pairwise summation order for 12 float64 values (np.sum of a length-12 array)
DI double np_sum12(const double* v) {
    double r = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    r += v[8]; r += v[9]; r += v[10]; r += v[11];
    return r;
}

template <bool DEBUG, int CONE>
__global__ void __launch_bounds__(QG_BLOCK, QG_MINBLOCKS)
qg_step_kernel(const QgModelC* __restrict__ gm, const float4* __restrict__ gverts, const int4* __restrict__ adj4,
               const int4* __restrict__ cadj4, float4* __restrict__ S, int N, const float* __restrict__ action,
               int clip_action, int frame_skip, float* __restrict__ obs, float* __restrict__ reward,
               float* __restrict__ terms, unsigned char* __restrict__ terminated, float* __restrict__ terminal_obs,
               QgStepOpts opts, QgCounters* __restrict__ ctr, QgDebugOut dbg, const int* __restrict__ perm,
               unsigned char* __restrict__ bin_key, int* __restrict__ chunk_ctr, int nchunks, int env0, int cnt) {
    extern __shared__ __align__(16) unsigned char smem[];
    QgModelC& P = *reinterpret_cast<QgModelC*>(smem);
    float4* sverts = reinterpret_cast<float4*>(smem + ((sizeof(QgModelC) + 15) & ~size_t(15)));
    // per-warp scratch of the quad all-reduce, behind the vertex table
    static_assert(QG_CQ_FLOATS <= QG_QR_SLOTS * 32, "the collision queue aliases the reduction rows");
    float* sred = reinterpret_cast<float*>(sverts + gm->nvert) + (threadIdx.x >> 5) * (QG_QR_SLOTS * 32);
    // The collision queue of this warp ALIASES its reduction rows: the two are never live together (block barriers
    // separate the collision phase from the reductions before and after it), and 27 KB less shared memory per block is
    // 27 KB more L1 for the thread-local contact tables and link frames.
    const WarpQueue wq = warp_queue(sred);
    // resident state of the block's environments (qg_step.cuh, StateRef), behind the reduction rows of all warps
    float* sbase = reinterpret_cast<float*>(sverts + gm->nvert) + (blockDim.x >> 5) * (QG_QR_SLOTS * 32);
    float* sleg = sbase + SB_NWORDS * (blockDim.x >> 2);
    {
        const int4* src = reinterpret_cast<const int4*>(gm);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(QgModelC) / 16); i += blockDim.x) dst[i] = src[i];
        int nv = gm->nvert;
        for (int i = threadIdx.x; i < nv; i += blockDim.x) sverts[i] = gverts[i];
    }
    // Persistent blocks (one per SM): the model tables are staged once, then the block pulls chunks of
    // blockDim.x / 4 environments from a device counter until the batch is done (dynamic: chunks differ in cost).
    __shared__ int s_chunk;
    __shared__ unsigned s_ctr[QG_BLOCK / 32][QG_C_COUNT];   // per-warp counters, flushed once when the block is done
    if ((threadIdx.x & 31) < QG_C_COUNT) s_ctr[threadIdx.x >> 5][threadIdx.x & 31] = 0u;
    const QgDebugOut dbg_in = dbg;
#pragma unroll 1
    for (;;) {
    __syncthreads();   // staging done / previous chunk finished with s_chunk
    if (threadIdx.x == 0) s_chunk = atomicAdd(chunk_ctr, 1);
    __syncthreads();
    if (s_chunk >= nchunks) break;
    // chunks are handed out from the END of the slot order: the binning permutation sorts the environments by the solver
    // effort of their last step, ascending, so the expensive chunks start first and the cheap ones fill the tail of the
    // launch (longest-processing-time-first on the SMs)
    const int chunk = nchunks - 1 - s_chunk;
    dbg = dbg_in;
    const int t = chunk * blockDim.x + threadIdx.x;
    const int leg = t & 3;
    // One launch covers the environments [env0, env0 + cnt) of the batch (the whole batch, or one segment of the
    // pipelined host path); N stays the plane stride.  Quads past the end shadow the last environment (same reads, no
    // writes) so that every thread of the block reaches the same barriers.
    const bool valid = (t >> 2) < cnt;
    // `perm` (optional, this launch's slice) maps quad slots to environments: environments with similar contact / solver
    // effort in the previous launch share warps (less SIMT divergence); results per environment do not depend on the slot
    const int slot = valid ? (t >> 2) : cnt - 1;
    const int env = perm ? perm[slot] : env0 + slot;
    if (!valid) { dbg.qacc = dbg.qacc_smooth = dbg.qfrc_bias = dbg.M = dbg.sensordata = nullptr; dbg.counts = nullptr; }
    const unsigned qm = 0xFu << (threadIdx.x & 28);
    WarpCounters wc;
    wc.row = s_ctr[threadIdx.x >> 5];
    wc.count = valid;
    wc.lane = threadIdx.x & 31;
    QuadRed qr;
    qr.s = sred;
    qr.lane = threadIdx.x & 31;
    qr.qm = qm;

    {
        StateRef SR;
        SR.nb = blockDim.x >> 2;
        SR.nl = blockDim.x;
        SR.b = sbase + (threadIdx.x >> 2);
        SR.l = sleg + threadIdx.x;
        planes_to_shared(S, N, env, leg, SR);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = action[(size_t)env * 12 + 3 * leg + k];
            if (clip_action) a = fminf(fmaxf(a, -1.f), 1.f);
            SLW(SR, SL_CTRL + k) = a;
        }
        __syncwarp(qm);
        if (sb_time(SR) < opts.settling_time) {
#pragma unroll
            for (int k = 0; k < 3; ++k) SLW(SR, SL_CTRL + k) = opts.reset_ctrl[3 * leg + k];
        }

        StepStats st;
        st.ncon = st.nefc = st.niter = st.nls = 0;
        st.last_ls = 0;
        int diverged = 0;
        SensorOut so;
        Contacts C;
        const int max_iter = opts.max_iter, ls_iter = opts.ls_iter;
#pragma unroll 1
        for (int s = 0; s < frame_skip; ++s) {
#if QG_BLOCKSYNC == 8 || QG_BLOCKSYNC == 9
            qg_pair_sync();
#elif QG_BLOCKSYNC && QG_BLOCKSYNC != 6
            __syncthreads();
#endif
            if (qsumi(shared_bad(SR) ? 1 : 0, qm) > 0) {  // mj_checkPos / mj_checkVel
                __syncwarp(qm);
                reset_shared(P, SR, leg, opts, env, __float_as_int(SBW(SR, SB_EPISODE)), true);
                __syncwarp(qm);
                diverged += (leg == 0);
            }
            physics_step<DEBUG, CONE>(P, sverts, adj4, cadj4, SR, leg, qr, wq, max_iter, ls_iter, s == frame_skip - 1, so,
                                st, wc, C, dbg, env);
        }
        // the epilogue's view of the state after the step
        int episode = __float_as_int(SBW(SR, SB_EPISODE)), flags = __float_as_int(SBW(SR, SB_FLAGS));
        double first_cc = __hiloint2double(__float_as_int(SBW(SR, SB_FCC + 1)), __float_as_int(SBW(SR, SB_FCC)));
        const double time_now = sb_time(SR);
        const float vw_x = SBW(SR, SB_VLIN);

        // ---- reward terms (float64 from the float32 sensordata / ctrl / state, reference formulas)
        double total = 0.0;
        if (opts.n_terms > 0) {
            const float* q0 = sleg + (threadIdx.x & ~3);     // leg words of the quad's lane 0
            double call[12], pall[12];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    call[3 * j + k] = (double)q0[(SL_CTRL + k) * SR.nl + j];
                    pall[3 * j + k] = (double)q0[(SL_PCTRL + k) * SR.nl + j];
                }
            for (int i = 0; i < opts.n_terms; ++i) {
                double v = 0.0, p = opts.term_p[i];
                switch (opts.term_id[i]) {
                    case 0: v = 1.0; break;
                    case 1: { double sq[12]; for (int k = 0; k < 12; ++k) sq[k] = call[k] * call[k]; v = np_sum12(sq); } break;
                    case 2: v = (double)vw_x; break;
                    case 3: v = (double)so.linvel.x * (double)so.pos.x; break;
                    case 4: v = fabs((double)so.linvel.y * (double)so.pos.y); break;
                    case 5: {
                        double sq[12];
                        for (int k = 0; k < 12; ++k) { double d = call[k] - pall[k]; sq[k] = d * d; }
                        double cost = np_sum12(sq);
                        if (!(flags & 1)) { first_cc = cost; flags |= 1; }
                        v = p * first_cc + (1.0 - p) * cost;
                    } break;
                    case 6: v = (double)so.zaxis.z; break;
                    case 7: v = fabs((double)so.pos.z - p); break;
                    case 8: {
                        double sq = 0.0;
                        for (int k = 0; k < 12; ++k) { double d = (call[k] - (double)opts.reset_ctrl[k]) / 12.0; sq += d * d; }
                        v = sqrt(sq);
                    } break;
                    case 9: v = exp((double)so.zaxis.z) - 1.0; break;
                    case 10: v = exp(fabs((double)so.pos.z - p)) - 1.0; break;
                    default: v = 0.0;
                }
                double wv = opts.term_w[i] * v;
                total += wv;
                if (terms && leg == 0 && valid) terms[(size_t)env * opts.n_terms + i] = (float)wv;
            }
        }

        // ---- termination (time limit is `terminated`, never truncated: quadruped.py:149-151,178-179)
        const bool term = (time_now >= opts.max_time) || (opts.flip_termination && so.zaxis.z < 0.f);
        if (leg == 0 && valid) {
            reward[env] = (float)total;
            terminated[env] = term ? 1 : 0;
        }

        // ---- sensordata of the last forward pass (lags the state by one physics step, as in the reference).
        //      After an auto-reset the observation is the reset observation: zeros (quadruped.py:120,138).
        auto write_obs = [&](float* o, float gate) {
            o[3 * leg] = gate * so.jq[0]; o[3 * leg + 1] = gate * so.jq[1]; o[3 * leg + 2] = gate * so.jq[2];
            if (leg == 0) { o[12] = gate * so.acc.x; o[13] = gate * so.acc.y; o[14] = gate * so.acc.z; o[15] = gate * so.gyro.x; o[16] = gate * so.gyro.y; o[17] = gate * so.gyro.z; }
            else if (leg == 1) { o[18] = gate * so.pos.x; o[19] = gate * so.pos.y; o[20] = gate * so.pos.z; o[21] = gate * so.linvel.x; o[22] = gate * so.linvel.y; o[23] = gate * so.linvel.z; }
            else if (leg == 2) { o[24] = gate * so.xaxis.x; o[25] = gate * so.xaxis.y; o[26] = gate * so.xaxis.z; o[27] = gate * so.zaxis.x; o[28] = gate * so.zaxis.y; o[29] = gate * so.zaxis.z; }
            else { o[30] = gate * so.vel.x; o[31] = gate * so.vel.y; o[32] = gate * so.vel.z; }
        };
        const bool do_reset = term && opts.auto_reset;
        if (valid) {
        if (do_reset) {
            float* o = obs + (size_t)env * 33;
            for (int k = leg; k < 33; k += 4) o[k] = 0.f;
        } else {
            write_obs(obs + (size_t)env * 33, 1.f);
        }
        if (terminal_obs) {
            if (term) write_obs(terminal_obs + (size_t)env * 33, 1.f);
            else {
                float* o = terminal_obs + (size_t)env * 33;
                for (int k = leg; k < 33; k += 4) o[k] = 0.f;
            }
        }
        if (DEBUG && dbg.sensordata) write_obs(dbg.sensordata + (size_t)env * 33, 1.f);
        if (DEBUG && dbg.counts) {
            int nc = qsumi(st.ncon, qm), ne = qsumi(st.nefc, qm);
            if (leg == 0) {
                dbg.counts[env * 4] = nc; dbg.counts[env * 4 + 1] = ne;
                dbg.counts[env * 4 + 2] = st.niter; dbg.counts[env * 4 + 3] = st.nls;
            }
        }
        }
        __syncwarp(qm);          // all four lanes are done reading the shared state
        if (leg == 3) {          // episode counter and reward memory back to the shared words the store below reads
            SBW(SR, SB_FLAGS) = __int_as_float(flags);
            SBW(SR, SB_FCC) = __int_as_float(__double2loint(first_cc));
            SBW(SR, SB_FCC + 1) = __int_as_float(__double2hiint(first_cc));
            SBW(SR, SB_EPISODE) = __int_as_float(do_reset ? episode + 1 : episode);
        }
        if (do_reset) reset_shared(P, SR, leg, opts, env, episode + 1, false);
        __syncwarp(qm);
        if (valid) shared_to_planes(S, N, env, leg, SR);

        wc_add(wc, QG_C_STEPS, leg == 0 ? frame_skip : 0);
        wc_add(wc, QG_C_DIVERGED, diverged);
        wc_add(wc, QG_C_EPISODES, (leg == 0 && term) ? 1 : 0);
        if (bin_key) {   // key of the next launch's binning: largest per-leg contact count x line-search evaluations of the last physics step
            int mnc = max(C.n, __shfl_xor_sync(qm, C.n, 1));
            mnc = max(mnc, __shfl_xor_sync(qm, mnc, 2));
            if (leg == 0 && valid) bin_key[env] = (unsigned char)(min(mnc, 3) * 16 + min(st.last_ls, 15));
        }
    }

    }   // chunk loop
    // ---- counters: one atomic per warp and counter for the whole launch
    __syncwarp();
    if ((threadIdx.x & 31) < QG_C_COUNT) {
        unsigned v = s_ctr[threadIdx.x >> 5][threadIdx.x & 31];
        if (v) atomicAdd(reinterpret_cast<unsigned long long*>(ctr) + (threadIdx.x & 31), (unsigned long long)v);
    }
    // the last block to leave re-arms the chunk counter for the next launch (no host-side state: graph-capture safe)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(chunk_ctr + 1, 1) == (int)gridDim.x - 1) { chunk_ctr[0] = 0; chunk_ctr[1] = 0; }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void qg_reset_kernel(const QgModelC* __restrict__ gm, float4* __restrict__ S, int N,
                                const unsigned char* __restrict__ mask, QgStepOpts opts, int clear_env_state) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int env = t >> 2, leg = t & 3;
    if (env >= N) return;
    if (mask && !mask[env]) return;
    LaneState L;
    int episode, flags;
    double first_cc;
    load_lane(S, N, env, leg, L, episode, first_cc, flags);
    if (clear_env_state) { episode = 0; flags = 0; first_cc = 0.0; }
    else episode++;
    reset_lane(*gm, L, leg, opts, env, episode);
    store_lane(S, N, env, leg, L, episode, first_cc, flags);
}

// planes <-> MuJoCo-layout arrays
__global__ void qg_get_state_kernel(const float4* __restrict__ S, int N, float* qpos, float* qvel, float* act,
                                    float* warm, double* time, float* ctrl) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int env = t >> 2, leg = t & 3;
    if (env >= N) return;
    LaneState L;
    int episode, flags;
    double first_cc;
    load_lane(S, N, env, leg, L, episode, first_cc, flags);
    for (int k = 0; k < 3; ++k) {
        if (qpos) qpos[(size_t)env * 19 + 7 + 3 * leg + k] = L.q[k];
        if (qvel) qvel[(size_t)env * 18 + 6 + 3 * leg + k] = L.qd[k];
        if (act) act[(size_t)env * 12 + 3 * leg + k] = L.act[k];
        if (warm) warm[(size_t)env * 18 + 6 + 3 * leg + k] = L.wj[k];
        if (ctrl) ctrl[(size_t)env * 12 + 3 * leg + k] = L.ctrl[k];
    }
    if (leg == 0) {
        if (qpos) {
            float* q = qpos + (size_t)env * 19;
            q[0] = L.pb.x; q[1] = L.pb.y; q[2] = L.pb.z; q[3] = L.qw; q[4] = L.qx; q[5] = L.qy; q[6] = L.qz;
        }
        if (qvel) {
            float* v = qvel + (size_t)env * 18;
            v[0] = L.vw.x; v[1] = L.vw.y; v[2] = L.vw.z; v[3] = L.om.x; v[4] = L.om.y; v[5] = L.om.z;
        }
        if (warm) {
            float* v = warm + (size_t)env * 18;
            v[0] = L.wl.x; v[1] = L.wl.y; v[2] = L.wl.z; v[3] = L.wa.x; v[4] = L.wa.y; v[5] = L.wa.z;
        }
        if (time) time[env] = L.time;
    }
}

__global__ void qg_set_state_kernel(float4* __restrict__ S, int N, const float* qpos, const float* qvel,
                                    const float* act, const float* warm, const double* time, const float* ctrl) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int env = t >> 2, leg = t & 3;
    if (env >= N) return;
    LaneState L;
    int episode, flags;
    double first_cc;
    load_lane(S, N, env, leg, L, episode, first_cc, flags);
    for (int k = 0; k < 3; ++k) {
        if (qpos) L.q[k] = qpos[(size_t)env * 19 + 7 + 3 * leg + k];
        if (qvel) L.qd[k] = qvel[(size_t)env * 18 + 6 + 3 * leg + k];
        if (act) L.act[k] = act[(size_t)env * 12 + 3 * leg + k];
        if (warm) L.wj[k] = warm[(size_t)env * 18 + 6 + 3 * leg + k];
        if (ctrl) L.ctrl[k] = ctrl[(size_t)env * 12 + 3 * leg + k];
    }
    if (qpos) {
        const float* q = qpos + (size_t)env * 19;
        L.pb = V3(q[0], q[1], q[2]);
        L.qw = q[3]; L.qx = q[4]; L.qy = q[5]; L.qz = q[6];
    }
    if (qvel) {
        const float* v = qvel + (size_t)env * 18;
        L.vw = V3(v[0], v[1], v[2]);
        L.om = V3(v[3], v[4], v[5]);
    }
    if (warm) {
        const float* v = warm + (size_t)env * 18;
        L.wl = V3(v[0], v[1], v[2]);
        L.wa = V3(v[3], v[4], v[5]);
    }
    if (time) L.time = time[env];
    store_lane(S, N, env, leg, L, episode, first_cc, flags);
}

// FP32 FFMA peak: 8 independent dependent-chains per thread, 2 flops per FFMA
__global__ void qg_ffma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}


// ---------------------------------------------------------------------------------------------
// Counting sort of the environments by the key the step kernel left (64 bins: largest per-leg contact count x line-search
// evaluations of the last physics step -- the trip counts a warp pays the maximum of; evaluations separate the
// environments slightly better than Newton iterations).  Key = min(contacts, 3) * 16 + min(evaluations, 15); measured
// alternatives: contacts 0..7 x evaluations 0..7 1.391 ms, this one 1.381 ms, evaluations only 1.49 ms, env-total contacts 1.41+.  Two tiny launches per env.step(); the order inside a bin is irrelevant.
#define QG_NBINS 64
__global__ void qg_bin_hist_kernel(const unsigned char* __restrict__ key, int N, int* __restrict__ count) {
    __shared__ int sc[QG_NBINS];
    if (threadIdx.x < QG_NBINS) sc[threadIdx.x] = 0;
    __syncthreads();
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N; e += gridDim.x * blockDim.x) atomicAdd(&sc[key[e] & (QG_NBINS - 1)], 1);
    __syncthreads();
    if (threadIdx.x < QG_NBINS && sc[threadIdx.x]) atomicAdd(&count[threadIdx.x], sc[threadIdx.x]);
}
__global__ void qg_bin_scatter_kernel(const unsigned char* __restrict__ key, int N, const int* __restrict__ count,
                                      int* __restrict__ cursor, int* __restrict__ perm, int env0) {
    __shared__ int base[QG_NBINS], sc[QG_NBINS], sbase[QG_NBINS];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < QG_NBINS; ++b) { base[b] = acc; acc += count[b]; }
    }
    if (threadIdx.x < QG_NBINS) sc[threadIdx.x] = 0;
    __syncthreads();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0, local = 0;
    if (e < N) { b = key[e] & (QG_NBINS - 1); local = atomicAdd(&sc[b], 1); }
    __syncthreads();
    if (threadIdx.x < QG_NBINS && sc[threadIdx.x]) sbase[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], sc[threadIdx.x]);
    __syncthreads();
    if (e < N) perm[base[b] + sbase[b] + local] = env0 + e;   // key / perm point at the slice, ids are absolute
}
