// qg_step.cuh -- one physics step (mj_step restated B200-first) for ONE LEG PER LANE.
//
// Replaces mujoco.mj_step at /root/reference/src/envs/quadruped.py:165 for the model class of
// /root/reference/src/models/quadruped/quadruped.xml.
//
// Mapping: 4 consecutive lanes (a "quad") own one environment; lane l owns leg l (3 hinge links,
// their geoms, joint limits and contacts) and carries a replicated copy of the free-base state.
// All dynamics are written in the BASE BODY frame "B" (origin = base position, axes = base axes):
//   * leg kinematics, inertias and Jacobians do not depend on the base pose at all;
//   * the generalised coordinates differ from MuJoCo's only in that the three linear base dofs are
//     expressed in B instead of the world (a_B = R^T a_world); because damping and armature of
//     those dofs are isotropic this is an exact change of variables for M, bias, J and the solver;
//   * every contact is plane-vs-robot, so its Jacobian touches only the 6 base dofs and the 3 dofs
//     of the owning leg.  M and the Newton Hessian H = M + J^T D J therefore share one "arrow"
//     sparsity pattern: a 6x6 base block, four 6x3 couplings and four 3x3 leg blocks.  Each lane
//     eliminates its own 3x3 block in registers; the 6x6 Schur complement is summed across the
//     quad through a shared-memory all-reduce (QuadRed) and factorised redundantly by all four lanes.
// The kernel is latency bound (255 registers -> 8 warps per SM), so the code is shaped by
// three rules: keep the SASS small and shared between warps (ONE inlined arrow_solve inside a phase machine serves
// the three SPD solves of a step; rolled loops; the warps of a block walk the solver loop together through
// __syncthreads_or so that one instruction-cache fill serves all of them), keep serial dependency chains short
// (MUFU reciprocals / square roots, list offsets fetched with the vertex, 4 neighbours per trip), and keep lanes
// busy (the collision narrow phase is balanced over the warp through a shared-memory queue).
#pragma once
#include "qg_math.cuh"
#include "qg_model.h"

struct LaneState {
    v3 pb;                 // base position (world)
    float qw, qx, qy, qz;  // base quaternion
    v3 vw, om;             // base linear velocity (world), angular velocity (body)
    v3 wl, wa;             // qacc_warmstart of the base (world lin, body ang)
    double time;
    float q[3], qd[3], act[3], wj[3], ctrl[3];
};

// The state of the environments a block is stepping lives in SHARED MEMORY for the whole env.step(), de-replicated: the
// base words once per environment, the leg words once per lane.  physics_step reads what a stage needs when it needs it
// (nothing of the state is carried in registers across the solver), and writes the integrated state back at the end.
// Word w of environment slot e: base[w * nb + e]; word w of lane t: leg[w * nl + t]  (a warp's 8 environments read 8
// consecutive words with a 4-lane broadcast each; leg words are consecutive over the lanes: no bank conflicts).
enum { SB_POS = 0, SB_QUAT = 3, SB_VLIN = 7, SB_VANG = 10, SB_WLIN = 13, SB_WANG = 16, SB_TIME = 19 /* lo, hi */,
       SB_EPISODE = 21, SB_FLAGS = 22, SB_FCC = 23 /* first control cost: lo, hi */, SB_NWORDS = 25 };
enum { SL_Q = 0, SL_QD = 3, SL_ACT = 6, SL_WJ = 9, SL_CTRL = 12, SL_PCTRL = 15 /* data.ctrl before this env.step() */, SL_NWORDS = 18 };
struct StateRef {
    float* b;   // this environment's base word 0
    float* l;   // this lane's leg word 0
    int nb, nl; // strides between words (environments / threads per block)
};
DI float& SBW(const StateRef& r, int w) { return r.b[w * r.nb]; }
DI float& SLW(const StateRef& r, int w) { return r.l[w * r.nl]; }
DI v3 sb3(const StateRef& r, int w) { return V3(SBW(r, w), SBW(r, w + 1), SBW(r, w + 2)); }
DI double sb_time(const StateRef& r) { return __hiloint2double(__float_as_int(SBW(r, SB_TIME + 1)), __float_as_int(SBW(r, SB_TIME))); }
DI void sb_set_time(const StateRef& r, double t) {
    SBW(r, SB_TIME) = __int_as_float(__double2loint(t));
    SBW(r, SB_TIME + 1) = __int_as_float(__double2hiint(t));
}
// pose and velocities of the base and the lane's joint state (what the forward pass reads); the warm start is read where
// the solver needs it
DI void state_load(const StateRef& r, LaneState& S) {
    S.pb = sb3(r, SB_POS);
    S.qw = SBW(r, SB_QUAT); S.qx = SBW(r, SB_QUAT + 1); S.qy = SBW(r, SB_QUAT + 2); S.qz = SBW(r, SB_QUAT + 3);
    S.vw = sb3(r, SB_VLIN);
    S.om = sb3(r, SB_VANG);
#pragma unroll
    for (int k = 0; k < 3; ++k) { S.q[k] = SLW(r, SL_Q + k); S.qd[k] = SLW(r, SL_QD + k); S.act[k] = SLW(r, SL_ACT + k); S.ctrl[k] = SLW(r, SL_CTRL + k); }
}

// experiment (QG_BLOCKSYNC 8 / 9): lockstep domain = a PAIR of warps instead of the block.  8: warps w and w + 4 (the two
// warps of one SM sub-partition when warps are dealt round-robin), 9: warps 2k and 2k + 1 (control).
DI void qg_pair_sync() {
#if QG_BLOCKSYNC == 9
    const int id = 1 + ((threadIdx.x >> 6) & 3);
#else
    const int id = 1 + ((threadIdx.x >> 5) & 3);
#endif
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

struct SensorOut {
    float jq[3];
    v3 acc, gyro, pos, linvel, xaxis, zaxis, vel;
};

// Statistics.  The batch counters (qg_get_counters) are kept per WARP in shared memory: at a few warp-converged points
// of a physics step the lanes' values are summed with one REDUX instruction and lane 0 adds the result to its warp's row;
// nothing is carried in registers across the step (nine per-lane accumulators cost 2.4 % of the step time through the
// spills they caused).  Per-environment counts exist only in the DEBUG instantiation (qg_debug_step).
enum { QG_C_STEPS = 0, QG_C_NCON, QG_C_NEFC, QG_C_NITER, QG_C_NLS, QG_C_NVERT, QG_C_DIVERGED, QG_C_OVERFLOW, QG_C_EPISODES,
       QG_C_NACT, QG_C_COUNT };
struct WarpCounters {
    unsigned* row;   // this warp's QG_C_COUNT counters (shared memory)
    bool count;      // false for the shadow quads past the end of the batch
    int lane;
};
DI void wc_add(const WarpCounters& w, int i, int v) {   // all 32 lanes must call it together
    unsigned t = __reduce_add_sync(0xffffffffu, (unsigned)(w.count ? v : 0));
    if (w.lane == 0) w.row[i] += t;
}
struct StepStats {
    int ncon, nefc, niter, nls;   // per lane, DEBUG only
    int last_ls;                  // line-search evaluations of the most recent physics step: binning key of the next launch
};

// per-lane contact table (thread-local memory; only the first `nc` slots are ever touched).  Four float4 per contact
// so that every access is one 128-bit local load / store instead of four 32-bit ones.
struct Contacts {
    float4 geo[QG_MAXCON_LANE];   // contact point x, y, z (base frame) and the row stiffness D
    float4 par[QG_MAXCON_LANE];   // mu, reference-acceleration damping B, K * imp * r, link level (int bits)
    float4 jar[QG_MAXCON_LANE];   // J a - aref of the contact's rows (3 used with the elliptic cone)
    float4 jv[QG_MAXCON_LANE];    // J v of the current search direction
    int n;
};
DI void ld4(float4 v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
DI float4 st4(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }

__device__ __noinline__ float impedance(float r, float d0, float dmax, float width, float mid, float power) {
    float x = fabsf(r) / fmaxf(1e-15f, width), y;
    if (x >= 1.f) y = 1.f;
    else if (x <= 0.f) y = 0.f;
    else if (power == 2.f) y = (x <= mid) ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
    else if (power == 1.f) y = x;
    else y = (x <= mid) ? powf(x / mid, power) * mid : 1.f - powf((1.f - x) / (1.f - mid), power) * (1.f - mid);
    float v = fmaf(y, dmax - d0, d0);
    return fminf(fmaxf(v, 1e-4f), 0.9999f);
}

// ---------------------------------------------------------------------------------------------
// arrow-structured SPD solve  [Abb Abl; Alb All] [xb; xl] = [rb; rl]
//   All (3x3 sym: 00 01 02 11 12 22) and Abl (6x3, [r*3+c]) are lane-local,
//   Abb = Acommon (identical on the 4 lanes) + sum over lanes of Alocal,
//   rb is identical on all lanes.
DI void arrow_solve(const float* All, const float* Abl, const float* Alocal, const float* Acommon,
                    const float* rb, const float* rl, const QuadRed& qr, float* xb, float* xl) {
    // LDL^T of the leg block
    float d0 = fmaxf(All[0], 1e-12f), i0 = 1.f / d0;
    float l10 = All[1] * i0, l20 = All[2] * i0;
    float d1 = fmaxf(All[3] - l10 * All[1], 1e-12f), i1 = 1.f / d1;
    float l21 = (All[4] - l20 * All[1]) * i1;
    float d2 = fmaxf(All[5] - l20 * All[2] - l21 * (All[4] - l20 * All[1]), 1e-12f), i2 = 1.f / d2;
#define LEG_SOLVE(b0, b1, b2, o0, o1, o2)            \
    {                                                 \
        float y0 = (b0), y1 = (b1)-l10 * y0;          \
        float y2 = (b2)-l20 * y0 - l21 * y1;          \
        o2 = y2 * i2;                                 \
        o1 = y1 * i1 - l21 * o2;                      \
        o0 = y0 * i0 - l10 * o1 - l20 * o2;           \
    }
    float Y[18];  // Y[r*3+c] = (All^-1 Abl[r,:]^T)[c]
#pragma unroll
    for (int r = 0; r < 6; ++r) LEG_SOLVE(Abl[r * 3], Abl[r * 3 + 1], Abl[r * 3 + 2], Y[r * 3], Y[r * 3 + 1], Y[r * 3 + 2]);
    float t0, t1, t2;
    LEG_SOLVE(rl[0], rl[1], rl[2], t0, t1, t2);
    float A[21], b[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            float s = fmaf(Abl[i * 3], Y[j * 3], fmaf(Abl[i * 3 + 1], Y[j * 3 + 1], Abl[i * 3 + 2] * Y[j * 3 + 2]));
            qr_put(qr, IX6(i, j), Alocal[IX6(i, j)] - s);
        }
        qr_put(qr, 21 + i, fmaf(Abl[i * 3], t0, fmaf(Abl[i * 3 + 1], t1, Abl[i * 3 + 2] * t2)));
    }
    qr_sync(qr);
#pragma unroll
    for (int i = 0; i < 21; ++i) A[i] = Acommon[i] + qr_get(qr, i);
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] = rb[i] - qr_get(qr, 21 + i);
    qr_sync(qr);
    // dense Cholesky of the 6x6 Schur complement (redundant on the 4 lanes, identical bits)
    float inv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float s = A[IX6(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) s -= A[IX6(j, k)] * A[IX6(j, k)];
        inv[j] = rsqrtf(fmaxf(s, 1e-12f));   // only 1/L_jj is ever used
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            float t = A[IX6(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t -= A[IX6(i, k)] * A[IX6(j, k)];
            A[IX6(i, j)] = t * inv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= A[IX6(i, k)] * b[k];
        b[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float s = b[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s -= A[IX6(k, i)] * xb[k];
        xb[i] = s * inv[i];
    }
    float x0 = t0, x1 = t1, x2 = t2;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        x0 -= Y[r * 3] * xb[r];
        x1 -= Y[r * 3 + 1] * xb[r];
        x2 -= Y[r * 3 + 2] * xb[r];
    }
    xl[0] = x0; xl[1] = x1; xl[2] = x2;
#undef LEG_SOLVE
}

// y = M x on the arrow pattern, the lane-local half: yl complete, yb_part = this lane's coupling contribution to the base
// rows (the caller adds Mbb xb and the quad sum of yb_part, inside a reduction it needs anyway)
DI void arrow_matvec_local(const float* Mll, const float* Mbl, const float* xb, const float* xl, float* yb_part, float* yl) {
    yl[0] = fmaf(Mll[0], xl[0], fmaf(Mll[1], xl[1], Mll[2] * xl[2]));
    yl[1] = fmaf(Mll[1], xl[0], fmaf(Mll[3], xl[1], Mll[4] * xl[2]));
    yl[2] = fmaf(Mll[2], xl[0], fmaf(Mll[4], xl[1], Mll[5] * xl[2]));
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        yl[0] = fmaf(Mbl[r * 3], xb[r], yl[0]);
        yl[1] = fmaf(Mbl[r * 3 + 1], xb[r], yl[1]);
        yl[2] = fmaf(Mbl[r * 3 + 2], xb[r], yl[2]);
        yb_part[r] = fmaf(Mbl[r * 3], xl[0], fmaf(Mbl[r * 3 + 1], xl[1], Mbl[r * 3 + 2] * xl[2]));
    }
}

// spatial velocity prefixes of a generalised vector: U[k] + W[k] x p = velocity of a point p on link k
DI void twist(const float* xb, const float* xl, const v3* sl, const v3* sa, v3* U, v3* W) {
    U[0] = V3(xb[0], xb[1], xb[2]);
    W[0] = V3(xb[3], xb[4], xb[5]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        U[j + 1] = fma3(xl[j], sl[j], U[j]);
        W[j + 1] = fma3(xl[j], sa[j], W[j]);
    }
}
DI float sel3(const float* a, int k) { return k == 0 ? a[0] : (k == 1 ? a[1] : a[2]); }
DI v3 sel4(const v3* A, int lev) { return lev == 0 ? A[0] : (lev == 1 ? A[1] : (lev == 2 ? A[2] : A[3])); }

// rows of the 4-sided friction pyramid for a point velocity u: n + mu t1, n - mu t1, n + mu t2, n - mu t2
// with n = +z, t1 = +y, t2 = -x of the world (rows of the base rotation in B coordinates)
DI void pyramid_rows(v3 u, v3 tx, v3 ty, v3 up, float mu, float* r) {
    float un = dot(up, u), u1 = mu * dot(ty, u), u2 = -mu * dot(tx, u);
    r[0] = un + u1; r[1] = un - u1; r[2] = un + u2; r[3] = un - u2;
}

// rows of one contact for a point velocity u: pyramidal (4 rows, see pyramid_rows) or elliptic (normal, t1 = +y, t2 = -x)
template <int CONE>
DI void contact_rows(v3 u, v3 tx, v3 ty, v3 up, float mu, float* r) {
    if (CONE) { r[0] = dot(up, u); r[1] = dot(ty, u); r[2] = -dot(tx, u); r[3] = 0.f; }
    else pyramid_rows(u, tx, ty, up, mu, r);
}

// Elliptic friction cone of one contact in the scaled coordinates U = (mu*jar_n, fri*jar_t1, fri*jar_t2):
// zone 0 = inside the cone (no force), 1 = polar cone (all three rows quadratic), 2 = cone surface
// (cost 1/2 Dm (N - mu T)^2).  mu = fri / sqrt(impratio) is the regularised friction coefficient.
struct EllZ { int zone; float cost, f0, f1, f2, N, T, U1, U2, Dm, mu; };
DI EllZ ell_eval(float j0, float j1, float j2, float fri, float mus, float Dn, float Dt) {
    EllZ z;
    z.mu = fri * mus;
    z.N = j0 * z.mu; z.U1 = j1 * fri; z.U2 = j2 * fri;
    z.T = sqrtf(z.U1 * z.U1 + z.U2 * z.U2);
    z.Dm = Dn / fmaxf(z.mu * z.mu * (1.f + z.mu * z.mu), 1e-15f);
    z.cost = 0.f; z.f0 = z.f1 = z.f2 = 0.f;
    if (z.N >= z.mu * z.T || (z.T <= 0.f && z.N >= 0.f)) z.zone = 0;
    else if (z.mu * z.N + z.T <= 0.f || (z.T <= 0.f && z.N < 0.f)) {
        z.zone = 1;
        z.cost = 0.5f * (Dn * j0 * j0 + Dt * (j1 * j1 + j2 * j2));
        z.f0 = -Dn * j0; z.f1 = -Dt * j1; z.f2 = -Dt * j2;
    } else {
        float NmT = z.N - z.mu * z.T;
        z.zone = 2;
        z.cost = 0.5f * z.Dm * NmT * NmT;
        z.f0 = -z.Dm * NmT * z.mu;
        float s = -z.f0 / z.T * fri;
        z.f1 = s * z.U1; z.f2 = s * z.U2;
    }
    return z;
}
// first and second derivative along jar + alpha*jv, evaluated at x = jar + alpha*jv; returns the zone at x
DI int ell_ls_acc(float x0, float x1, float x2, float v0, float v1, float v2, float fri, float mus, float Dn, float Dt,
                  float& e1, float& e2) {
    EllZ z = ell_eval(x0, x1, x2, fri, mus, Dn, Dt);
    if (z.zone == 1) {
        e1 += Dn * v0 * x0 + Dt * (v1 * x1 + v2 * x2);
        e2 += Dn * v0 * v0 + Dt * (v1 * v1 + v2 * v2);
    } else if (z.zone == 2) {
        float Np = v0 * z.mu, V1 = v1 * fri, V2 = v2 * fri;
        float Tp = (z.U1 * V1 + z.U2 * V2) / z.T, Tpp = (V1 * V1 + V2 * V2 - Tp * Tp) / z.T;
        float e = z.N - z.mu * z.T, ep = Np - z.mu * Tp;
        e1 += z.Dm * e * ep;
        e2 += z.Dm * (ep * ep - e * z.mu * Tpp);
    }
    return z.zone;
}

__device__ __noinline__ float2 sincos_ni(float x) {
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}

// ---------------------------------------------------------------------------------------------
// plane-vs-hull collision (mjc_PlaneConvex restated), balanced over the WARP.
//   pass 1 (every lane, its own geoms): link bounding-sphere reject, then an oriented-box cull (20 flops) per geom.
//           Survivors are laid out on a per-warp queue in shared memory (positions from one warp scan of the lanes'
//           candidate counts) as (search direction and centre height in the MESH frame, leg, geom).
//   pass 2 (any lane, any queue entry): hill-climbing support search on the polytope-edge graph (exact for a convex
//           hull; start vertex from a cube-map table of the direction; neighbour lists padded to int4 groups), then
//           up to 3 hull-graph neighbours of the support vertex become extra contacts.  Works in the mesh frame only
//           (heights and mutual distances are rigid-motion invariant), so it needs nothing of the owner's kinematics.
//   pass 3 (owner): transforms the reported vertices to the base frame and appends the contact records.
// Only 1 lane in 4 has a candidate at any time (legs in the air, culled geoms); with lane-local narrow phases the
// warp ran pass 2 at 7 of 32 lanes.  Contacts reach the owner in (geom, candidate) order whatever lane produced them,
// so results do not depend on the balancing.
// `fr` holds the lane's 4 link frames (level 0 = base): rotation (row major, link -> B) and position.
#define QG_CQ_CAP 64                                   // queue entries per warp (32 carried over + 32 pushed)
static_assert(QG_MAXGEOM_LANE <= 8, "collide_lane packs a lane's geom ids of one batch in 3 bits each");
#define QG_CQ_FLOATS (QG_CQ_CAP * 5 + 32 * 17)         // queue: float4 + meta; results: 4 float4 + count per slot
struct WarpQueue {
    float4* qd;    // [QG_CQ_CAP] (dl.x, dl.y, dl.z, zc)
    int* qmeta;    // [QG_CQ_CAP] leg | geom << 2
    float4* res;   // [32][4]     (vertex in the mesh frame, distance to the plane)
    int* rcnt;     // [32]
};
DI WarpQueue warp_queue(float* w) {
    WarpQueue q;
    q.qd = reinterpret_cast<float4*>(w);
    q.qmeta = reinterpret_cast<int*>(w + 4 * QG_CQ_CAP);
    q.res = reinterpret_cast<float4*>(w + 5 * QG_CQ_CAP);
    q.rcnt = reinterpret_cast<int*>(w + 5 * QG_CQ_CAP + 32 * 16);
    return q;
}

// pass 2a for one queue entry: support vertex of the hull along the search direction (hill climbing on the polytope-edge
// graph from a cube-map start vertex).  Returns the vertex (its .w carries the list offsets) and its height along dl.
DI float4 support_climb(const QgModelC& P, const float4* __restrict__ verts, const int4* __restrict__ cadj4, const QgGeomC& G,
                        v3 dl, float& hbest, int& nvert) {
    const float4* __restrict__ vt = verts + G.vert0;
    const int4* __restrict__ cl = cadj4 + G.cedge0;
    int best;
    {
        float ax = fabsf(dl.x), ay = fabsf(dl.y), az = fabsf(dl.z);
        int a = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
        float dm = a == 0 ? dl.x : (a == 1 ? dl.y : dl.z);
        float du = a == 0 ? dl.y : (a == 1 ? dl.z : dl.x);
        float dw = a == 0 ? dl.z : (a == 1 ? dl.x : dl.y);
        float inv = 1.f / fmaxf(fabsf(dm), 1e-20f);
        int iu = min(QG_DIRRES - 1, max(0, (int)((du * inv + 1.f) * (0.5f * QG_DIRRES))));
        int iv = min(QG_DIRRES - 1, max(0, (int)((dw * inv + 1.f) * (0.5f * QG_DIRRES))));
        best = P.dir_start[G.mesh][((2 * a + (dm > 0.f ? 1 : 0)) * QG_DIRRES + iu) * QG_DIRRES + iv];
    }
    float4 vbest = vt[best];   // .w carries the vertex' list offsets: climb graph | hull graph << 16 (int4 units)
    hbest = fmaf(dl.x, vbest.x, fmaf(dl.y, vbest.y, dl.z * vbest.z));
    int nev = 1;
#pragma unroll 1
    for (;;) {
        const int4* __restrict__ e = cl + (__float_as_int(vbest.w) & 0xffff);
        bool moved = false;
        int4 nb = __ldg(e++);
#pragma unroll 1
        for (;;) {
            const int4 nxt = __ldg(e++);   // next group in flight while this one is tested (tables are padded by one group)
            float4 v0 = vt[max(nb.x, 0)], v1 = vt[max(nb.y, 0)], v2 = vt[max(nb.z, 0)], v3_ = vt[max(nb.w, 0)];
            float h0 = fmaf(dl.x, v0.x, fmaf(dl.y, v0.y, dl.z * v0.z));
            float h1 = fmaf(dl.x, v1.x, fmaf(dl.y, v1.y, dl.z * v1.z));
            float h2 = fmaf(dl.x, v2.x, fmaf(dl.y, v2.y, dl.z * v2.z));
            float h3 = fmaf(dl.x, v3_.x, fmaf(dl.y, v3_.y, dl.z * v3_.z));
            if (nb.x >= 0 && h0 < hbest) { hbest = h0; vbest = v0; moved = true; }
            if (nb.y >= 0 && h1 < hbest) { hbest = h1; vbest = v1; moved = true; }
            if (nb.z >= 0 && h2 < hbest) { hbest = h2; vbest = v2; moved = true; }
            if (nb.w >= 0 && h3 < hbest) { hbest = h3; vbest = v3_; moved = true; }
            nev += (nb.x >= 0) + (nb.y >= 0) + (nb.z >= 0) + (nb.w >= 0);
            if (nb.w < 0) break;
            nb = nxt;
        }
        if (!moved) break;
    }
    nvert += nev;
    return vbest;
}

// The extra plane-mesh contacts: the hull-graph neighbours of the support vertex in list order (up to 4 candidates per
// geom).  A neighbour qualifies if it is within the margin and not closer than the tolerance to the candidates already
// taken.  `Taken` is that bookkeeping.
struct Taken {
    float4 prev0, prev1, prev2;
    int cnt, nout;
};
DI void taken_add(Taken& T, float4 v, float dv, float margin, float tol2, bool rule_first, float4* out) {
    if (T.cnt >= 4) return;
    float ax = v.x - T.prev0.x, ay = v.y - T.prev0.y, az = v.z - T.prev0.z;
    bool ok = !(ax * ax + ay * ay + az * az < tol2);
    if (!rule_first) {
        float bx = v.x - T.prev1.x, by = v.y - T.prev1.y, bz = v.z - T.prev1.z;
        float cx = v.x - T.prev2.x, cy = v.y - T.prev2.y, cz = v.z - T.prev2.z;
        if (T.cnt > 1 && bx * bx + by * by + bz * bz < tol2) ok = false;
        if (T.cnt > 2 && cx * cx + cy * cy + cz * cz < tol2) ok = false;
    }
    if (!ok) return;
    if (T.cnt == 1) T.prev1 = v; else if (T.cnt == 2) T.prev2 = v;
    T.cnt++;
    if (dv < margin) out[T.nout++] = make_float4(v.x, v.y, v.z, dv);
}

// pass 2b, one lane scans its own list: lists can be long (fan centres of flat faces: up to 104 neighbours) and nearly all
// entries fail the margin test, so four neighbours are fetched and tested per trip and the in-order bookkeeping runs only
// for the rare hits.
DI void scan_lane(const float4* __restrict__ vt, const int4* __restrict__ e, v3 dl, float zc, float margin, float tol2,
                  bool rule_first, Taken& T, float4* out) {
    int4 nb = __ldg(e++);
#pragma unroll 1
    for (;;) {
        const int4 nxt = __ldg(e++);       // next group in flight while this one is tested (tables are padded by one group)
        float4 v0 = vt[max(nb.x, 0)], v1 = vt[max(nb.y, 0)], v2 = vt[max(nb.z, 0)], v3_ = vt[max(nb.w, 0)];
        float d0 = zc + fmaf(dl.x, v0.x, fmaf(dl.y, v0.y, dl.z * v0.z));
        float d1 = zc + fmaf(dl.x, v1.x, fmaf(dl.y, v1.y, dl.z * v1.z));
        float d2 = zc + fmaf(dl.x, v2.x, fmaf(dl.y, v2.y, dl.z * v2.z));
        float d3 = zc + fmaf(dl.x, v3_.x, fmaf(dl.y, v3_.y, dl.z * v3_.z));
        bool c0 = nb.x >= 0 && d0 <= margin, c1 = nb.y >= 0 && d1 <= margin;
        bool c2 = nb.z >= 0 && d2 <= margin, c3 = nb.w >= 0 && d3 <= margin;
        if (c0 || c1 || c2 || c3) {
            if (c0) taken_add(T, v0, d0, margin, tol2, rule_first, out);
            if (c1) taken_add(T, v1, d1, margin, tol2, rule_first, out);
            if (c2) taken_add(T, v2, d2, margin, tol2, rule_first, out);
            if (c3) taken_add(T, v3_, d3, margin, tol2, rule_first, out);
        }
        if (T.cnt >= 4 || nb.w < 0) break;
        nb = nxt;
    }
}

// pass 2 for one queue entry (any lane, any entry); returns the number of contacts written to out[0..3].
// (A warp-cooperative version of the neighbour scan -- 32 neighbours per trip, one touching entry after the other -- was
// measured and is slower, 1.275 against 1.229 ms: in the steady state most lanes of a warp have a touching entry.)
DI int narrow_phase(const QgModelC& P, const float4* __restrict__ verts, const int4* __restrict__ adj4,
                    const int4* __restrict__ cadj4, float4 e4, int meta, float4* out, int& nvert) {
    const QgGeomC& G = P.geom[meta & 3][meta >> 2];
    const v3 dl = V3(e4.x, e4.y, e4.z);
    float hbest;
    const float4 vbest = support_climb(P, verts, cadj4, G, dl, hbest, nvert);
    const float margin = G.margin, dsup = e4.w + hbest;
    if (dsup > margin) return 0;
    Taken T;
    T.prev0 = vbest; T.prev1 = T.prev2 = make_float4(0.f, 0.f, 0.f, 0.f);
    T.cnt = 1; T.nout = 0;
    if (dsup < margin) out[T.nout++] = make_float4(vbest.x, vbest.y, vbest.z, dsup);  // rows only for dist < margin
    scan_lane(verts + G.vert0, adj4 + G.edge0 + ((unsigned)__float_as_int(vbest.w) >> 16), dl, e4.w, margin, G.tol2,
              P.rule_first != 0, T, out);
    return T.nout;
}

DI void collide_lane(const QgModelC& P, const float4* __restrict__ verts, const int4* __restrict__ adj4,
                     const int4* __restrict__ cadj4, int leg, const float* fr, v3 up, float zb, Contacts& C,
                     const WarpCounters& wc, const WarpQueue& wq, int lane) {
    __syncwarp();   // the queue aliases the quad-reduction rows: every quad of the warp is done with them
    v3 dB[4];      // "up" in each link frame
    float hk[4];   // height of each link origin above the plane
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        m3 Rk = ldm3(fr + 12 * k);
        dB[k] = tmul(Rk, up);
        hk[k] = zb + dot(up, ld3(fr + 12 * k + 9));
    }
    // pass 1, lane-local: a link whose bounding sphere clears the plane takes all its geoms with it (base, femur and
    // usually the shin of a standing robot: one compare each); the geoms of the other links go through the oriented-box
    // cull (conservative; same contacts as any cull).  Bit g of `cand` = geom g of this lane is a candidate.
    unsigned cand = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (hk[k] > P.link_reach[leg][k]) continue;
        const int g1 = P.glev[leg][k + 1];
#pragma unroll 1
        for (int g = P.glev[leg][k]; g < g1; ++g) {
            const QgGeomC& G = P.geom[leg][g];
            const float zc = hk[k] + dot(dB[k], ld3(G.pos));
            const v3 dl = tmul(ldm3(G.R), dB[k]);  // "up" in the mesh frame
            const float ext = fmaf(fabsf(dl.x), G.half[0], fmaf(fabsf(dl.y), G.half[1], fabsf(dl.z) * G.half[2]));
            if (zc - ext <= G.margin) cand |= 1u << g;
        }
    }
    // queue positions: the candidates of lane l occupy [off, off + n) of the warp's list, in geom order (warp scan)
    const int n = __popc(cand);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int off = incl - n, total = __shfl_sync(0xffffffffu, incl, 31);
    int nvert = 0;                       // vertex evaluations of this lane's narrow phases
    // rounds of 32 entries (one round unless the robots of this warp lie flat on the floor)
#pragma unroll 1
    for (int base = 0; base < total; base += 32) {
        {   // queue entry = (search direction and centre height in the MESH frame, leg, geom)
            unsigned c = cand;
            int idx = off - base;
#pragma unroll 1
            while (c) {
                const int g = __ffs(c) - 1;
                c &= c - 1;
                if (idx >= 0 && idx < 32) {
                    const QgGeomC& G = P.geom[leg][g];
                    const int lev = G.level;
                    const v3 d = sel4(dB, lev);
                    const float zc = (lev == 0 ? hk[0] : (lev == 1 ? hk[1] : (lev == 2 ? hk[2] : hk[3]))) + dot(d, ld3(G.pos));
                    const v3 dl = tmul(ldm3(G.R), d);
                    wq.qd[idx] = make_float4(dl.x, dl.y, dl.z, zc);
                    wq.qmeta[idx] = leg | (g << 2);
                }
                idx++;
            }
        }
        __syncwarp();
        // pass 2: lane i takes queue entry i (any lane, any entry: works in the mesh frame only)
        if (base + lane < total)
            wq.rcnt[lane] = narrow_phase(P, verts, adj4, cadj4, wq.qd[lane], wq.qmeta[lane], wq.res + lane * 4, nvert);
        __syncwarp();
        // pass 3 (owner): transform the reported vertices to the base frame and append the contact records, geom order
        unsigned c = cand;
        int idx = off - base;
#pragma unroll 1
        while (c) {
            const int g = __ffs(c) - 1;
            c &= c - 1;
            const int slot = idx++;
            if (slot < 0 || slot >= 32) continue;
            const int nres = wq.rcnt[slot];
            if (nres == 0) continue;
            const QgGeomC& G = P.geom[leg][g];
            const int lev = G.level;
            // static indices only: `fr` stays scalarised (registers / compiler-chosen spills) instead of a local array
            m3 Rk;
            v3 pk;
            if (lev == 0) { Rk = ldm3(fr); pk = ld3(fr + 9); }
            else if (lev == 1) { Rk = ldm3(fr + 12); pk = ld3(fr + 21); }
            else if (lev == 2) { Rk = ldm3(fr + 24); pk = ld3(fr + 33); }
            else { Rk = ldm3(fr + 36); pk = ld3(fr + 45); }
            v3 ctr = pk + mul(Rk, ld3(G.pos));
            m3 RB = matmul(Rk, ldm3(G.R));  // mesh frame -> B
#pragma unroll 1
            for (int k = 0; k < nres; ++k) {
                float4 v = wq.res[slot * 4 + k];
                if (C.n < QG_MAXCON_LANE) {
                    int ci = C.n++;
                    float dv = v.w;
                    v3 xv = ctr + mul(RB, V3(v.x, v.y, v.z));
                    v3 xc = fma3(-0.5f * dv, up, xv);
                    float r = dv - G.margin;
                    float imp = impedance(r, G.d0, G.dmax, G.width, G.mid, G.power);
                    C.geo[ci] = make_float4(xc.x, xc.y, xc.z, 1.f / fmaxf(1e-15f, (1.f - imp) / imp * G.Rfac));
                    C.par[ci] = make_float4(G.mu, G.B, G.K * imp * r, __int_as_float(lev));
                } else if (wc.count) atomicAdd(wc.row + QG_C_OVERFLOW, 1u);
            }
        }
        __syncwarp();
    }
    wc_add(wc, QG_C_NVERT, nvert);
}

// ---------------------------------------------------------------------------------------------
template <bool DEBUG, int CONE>
DI void physics_step(const QgModelC& P, const float4* __restrict__ verts, const int4* __restrict__ adj4,
                     const int4* __restrict__ cadj4, const StateRef& SR, int leg, const QuadRed& qr, const WarpQueue& wq,
                     int max_iter, int ls_iter, bool want_sensors, SensorOut& so, StepStats& st, const WarpCounters& wc, Contacts& C,
                     const QgDebugOut& dbg, int env) {
    const float h = P.timestep;
    constexpr int NR = CONE ? 3 : 4;   // rows per contact
    const float mus = P.mu_scale, impr = P.impratio;
    LaneState S;          // forward-pass copy of the state: its members die with their last use before the solver
    state_load(SR, S);
    // ---- base frame
    float qn = 1.f / sqrtf(S.qw * S.qw + S.qx * S.qx + S.qy * S.qy + S.qz * S.qz);
    float w = S.qw * qn, x = S.qx * qn, y = S.qy * qn, z = S.qz * qn;
    m3 Rb;  // world <- B
    Rb.r0 = V3(1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y));
    Rb.r1 = V3(2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x));
    Rb.r2 = V3(2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y));
    const v3 tx = Rb.r0, ty = Rb.r1, up = Rb.r2;  // world axes in B coordinates
    const v3 vB = tmul(Rb, S.vw);
    const v3 gB = tmul(Rb, ld3(P.grav));
    const float zb = S.pb.z - P.plane_z;

    // ---- leg kinematics (B coordinates) -> link frames, then collision over the lane's geoms
    float fr[48];
    fr[0] = 1.f; fr[1] = 0.f; fr[2] = 0.f; fr[3] = 0.f; fr[4] = 1.f; fr[5] = 0.f; fr[6] = 0.f; fr[7] = 0.f; fr[8] = 1.f;
    fr[9] = fr[10] = fr[11] = 0.f;
    {
        m3 Rk;
        Rk.r0 = V3(1, 0, 0); Rk.r1 = V3(0, 1, 0); Rk.r2 = V3(0, 0, 1);
        v3 pk = V3(0, 0, 0);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const QgJointC& J = P.joint[leg][k];
            pk = pk + mul(Rk, ld3(J.pos));
            m3 Rp = matmul(Rk, ldm3(J.Roff));
            float2 sc = sincos_ni(S.q[k] - J.q0);
            const float sn = sc.x, cs = sc.y;
            Rk.r0 = V3(cs * Rp.r0.x + sn * Rp.r0.y, cs * Rp.r0.y - sn * Rp.r0.x, Rp.r0.z);
            Rk.r1 = V3(cs * Rp.r1.x + sn * Rp.r1.y, cs * Rp.r1.y - sn * Rp.r1.x, Rp.r1.z);
            Rk.r2 = V3(cs * Rp.r2.x + sn * Rp.r2.y, cs * Rp.r2.y - sn * Rp.r2.x, Rp.r2.z);
            float* f = fr + 12 * (k + 1);
            f[0] = Rk.r0.x; f[1] = Rk.r0.y; f[2] = Rk.r0.z; f[3] = Rk.r1.x; f[4] = Rk.r1.y; f[5] = Rk.r1.z;
            f[6] = Rk.r2.x; f[7] = Rk.r2.y; f[8] = Rk.r2.z; f[9] = pk.x; f[10] = pk.y; f[11] = pk.z;
        }
    }
    C.n = 0;
    collide_lane(P, verts, adj4, cadj4, leg, fr, up, zb, C, wc, wq, qr.lane);
#if QG_BLOCKSYNC == 8 || QG_BLOCKSYNC == 9
    qg_pair_sync();   // experiment: only the two warps that share an SM sub-partition re-align
#elif QG_BLOCKSYNC >= 2
    __syncthreads();  // collision time varies per warp: re-align before the straight-line dynamics code
#endif

    // ---- velocity recursion and inertias along the chain
    v3 sl[3], sa[3];        // joint spatial motion about the B origin: linear p x a, angular a
    v3 ck[3];               // link CoM
    s3 Ik[3];               // link inertia about its CoM, B axes
    v3 Fk[3], Nk[3];        // RNE: inertial force and moment about the B origin
    {
        v3 omk = S.om, alk = V3(0, 0, 0), apk = -gB, pprev = V3(0, 0, 0);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const QgJointC& J = P.joint[leg][k];
            const float* f = fr + 12 * (k + 1);
            m3 Rk = ldm3(f);
            v3 pk = ld3(f + 9);
            v3 a = col2(Rk);
            // classical accelerations at zero generalised acceleration, gravity as base acceleration
            v3 r = pk - pprev;
            apk = apk + cross(alk, r) + cross(omk, cross(omk, r));
            v3 aq = S.qd[k] * a;
            alk = alk + cross(omk, aq);
            omk = omk + aq;
            pprev = pk;
            sa[k] = a;
            sl[k] = cross(pk, a);
            v3 dcm = mul(Rk, ld3(J.com));
            ck[k] = pk + dcm;
            s3 Ib;
            Ib.xx = J.I[0]; Ib.yy = J.I[1]; Ib.zz = J.I[2]; Ib.xy = J.I[3]; Ib.xz = J.I[4]; Ib.yz = J.I[5];
            Ik[k] = rot_sym(Rk, Ib);
            v3 ac = apk + cross(alk, dcm) + cross(omk, cross(omk, dcm));
            Fk[k] = J.mass * ac;
            Nk[k] = mul(Ik[k], alk) + cross(omk, mul(Ik[k], omk)) + cross(ck[k], Fk[k]);
        }
    }

    // ---- backward pass: composite inertias -> M blocks, RNE forces -> bias
    float Mll[6], Mbl[18], Mbb[21];
    int nefc = 0;   // constraint rows of the environment (quad sum)
    float bias_l[3];
    v3 fsum = V3(0, 0, 0), nsum = V3(0, 0, 0);
    float cm = 0.f;
    v3 chv = V3(0, 0, 0);
    s3 cI;
    cI.xx = cI.yy = cI.zz = cI.xy = cI.xz = cI.yz = 0.f;
#pragma unroll
    for (int k = 2; k >= 0; --k) {
        const QgJointC& J = P.joint[leg][k];
        float m = J.mass;
        v3 c = ck[k];
        float c2 = dot(c, c);
        cm += m;
        chv = fma3(m, c, chv);
        cI.xx += Ik[k].xx + m * (c2 - c.x * c.x);
        cI.yy += Ik[k].yy + m * (c2 - c.y * c.y);
        cI.zz += Ik[k].zz + m * (c2 - c.z * c.z);
        cI.xy += Ik[k].xy - m * c.x * c.y;
        cI.xz += Ik[k].xz - m * c.x * c.z;
        cI.yz += Ik[k].yz - m * c.y * c.z;
        v3 pl = fma3(cm, sl[k], cross(sa[k], chv));     // linear momentum per unit joint rate
        v3 Lo = mul(cI, sa[k]) + cross(chv, sl[k]);     // angular momentum about the B origin
        Mbl[0 * 3 + k] = pl.x; Mbl[1 * 3 + k] = pl.y; Mbl[2 * 3 + k] = pl.z;
        Mbl[3 * 3 + k] = Lo.x; Mbl[4 * 3 + k] = Lo.y; Mbl[5 * 3 + k] = Lo.z;
#pragma unroll
        for (int i = 0; i <= k; ++i) {
            float v = dot(sl[i], pl) + dot(sa[i], Lo);
            int idx = (i == 0) ? k : (i == 1 ? 2 + k : 5);  // (0,k)->0,1,2  (1,k)->3,4  (2,2)->5
            Mll[idx] = v + ((i == k) ? J.armature : 0.f);
        }
        fsum += Fk[k];
        nsum += Nk[k];
        bias_l[k] = dot(sl[k], fsum) + dot(sa[k], nsum);
    }
    // base: own body + legs (quad sums)
    {
        v3 c0 = ld3(P.base_com);
        float m0 = P.base_mass;
        v3 ac0 = -gB + cross(S.om, cross(S.om, c0));
        v3 F0 = m0 * ac0;
        s3 I0;
        I0.xx = P.base_I[0]; I0.yy = P.base_I[1]; I0.zz = P.base_I[2];
        I0.xy = P.base_I[3]; I0.xz = P.base_I[4]; I0.yz = P.base_I[5];
        v3 N0 = cross(S.om, mul(I0, S.om)) + cross(c0, F0);
        qr_put(qr, 0, fsum.x); qr_put(qr, 1, fsum.y); qr_put(qr, 2, fsum.z);
        qr_put(qr, 3, nsum.x); qr_put(qr, 4, nsum.y); qr_put(qr, 5, nsum.z);
        qr_put(qr, 6, cm); qr_put(qr, 7, chv.x); qr_put(qr, 8, chv.y); qr_put(qr, 9, chv.z);
        qr_put(qr, 10, cI.xx); qr_put(qr, 11, cI.yy); qr_put(qr, 12, cI.zz);
        qr_put(qr, 13, cI.xy); qr_put(qr, 14, cI.xz); qr_put(qr, 15, cI.yz);
        {   // the environment's constraint-row count rides along (same tests as the row setup below)
            int nl = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const QgJointC& J = P.joint[leg][k];
                nl += (J.limited && (S.q[k] - J.lo < 0.f || J.hi - S.q[k] < 0.f)) ? 1 : 0;
            }
            qr_put(qr, 16, (float)(NR * C.n + nl));
        }
        qr_sync(qr);
        fsum = V3(qr_get(qr, 0), qr_get(qr, 1), qr_get(qr, 2)) + F0;
        nsum = V3(qr_get(qr, 3), qr_get(qr, 4), qr_get(qr, 5)) + N0;
        float mt = qr_get(qr, 6) + m0;
        v3 ht = V3(qr_get(qr, 7), qr_get(qr, 8), qr_get(qr, 9)) + m0 * c0;
        float c2 = dot(c0, c0);
        s3 It;
        It.xx = qr_get(qr, 10) + I0.xx + m0 * (c2 - c0.x * c0.x);
        It.yy = qr_get(qr, 11) + I0.yy + m0 * (c2 - c0.y * c0.y);
        It.zz = qr_get(qr, 12) + I0.zz + m0 * (c2 - c0.z * c0.z);
        It.xy = qr_get(qr, 13) + I0.xy - m0 * c0.x * c0.y;
        It.xz = qr_get(qr, 14) + I0.xz - m0 * c0.x * c0.z;
        It.yz = qr_get(qr, 15) + I0.yz - m0 * c0.y * c0.z;
        nefc = (int)qr_get(qr, 16);
        qr_sync(qr);
#pragma unroll
        for (int i = 0; i < 21; ++i) Mbb[i] = 0.f;
        Mbb[IX6(0, 0)] = mt + P.base_arm[0]; Mbb[IX6(1, 1)] = mt + P.base_arm[1]; Mbb[IX6(2, 2)] = mt + P.base_arm[2];
        // angular rows x linear cols = [h]x
        Mbb[IX6(3, 1)] = -ht.z; Mbb[IX6(3, 2)] = ht.y;
        Mbb[IX6(4, 0)] = ht.z;  Mbb[IX6(4, 2)] = -ht.x;
        Mbb[IX6(5, 0)] = -ht.y; Mbb[IX6(5, 1)] = ht.x;
        Mbb[IX6(3, 3)] = It.xx + P.base_arm[3]; Mbb[IX6(4, 4)] = It.yy + P.base_arm[4]; Mbb[IX6(5, 5)] = It.zz + P.base_arm[5];
        Mbb[IX6(4, 3)] = It.xy; Mbb[IX6(5, 3)] = It.xz; Mbb[IX6(5, 4)] = It.yz;
    }

    // ---- passive + actuation -> qfrc_smooth (B form)
    float fsb[6], fsl[3];
    int unclamped = 0;   // bit k: servo k is not force-clamped (its kv term enters the implicit integrator's diagonal)
    fsb[0] = -P.base_damp[0] * vB.x - fsum.x; fsb[1] = -P.base_damp[1] * vB.y - fsum.y; fsb[2] = -P.base_damp[2] * vB.z - fsum.z;
    fsb[3] = -P.base_damp[3] * S.om.x - nsum.x; fsb[4] = -P.base_damp[4] * S.om.y - nsum.y; fsb[5] = -P.base_damp[5] * S.om.z - nsum.z;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const QgJointC& J = P.joint[leg][k];
        float tau = 0.f;
        if (J.has_act) {
            float c = S.ctrl[k];
            if (J.ctrl_limited) c = fminf(fmaxf(c, J.ctrl_lo), J.ctrl_hi);
            float ain = c;
            if (J.has_dyn) ain = S.act[k];
            float f = fmaf(J.kp, ain, J.b0) + J.b1 * (J.gear * S.q[k]) + J.b2 * (J.gear * S.qd[k]);
            bool clamped = false;
            if (J.frc_limited) {
                if (f <= J.frc_lo) { f = J.frc_lo; clamped = true; }
                else if (f >= J.frc_hi) { f = J.frc_hi; clamped = true; }
            }
            tau = J.gear * f;
            if (!clamped && P.integrator == 1) unclamped |= 1 << k;
        }
        fsl[k] = -J.damping * S.qd[k] - bias_l[k] + tau;
    }

    // ---- constraint rows: joint limits (own dofs) and pyramidal contacts (table C); jar starts as -aref
    // Active limits are rare (a joint past its range): they live in a thread-local table indexed by a run-time slot, so
    // that they cost a loop that is skipped instead of 12 registers held across the whole solve.
    // Record: (sign * (dof + 1), D, J a - aref, J v).
    float4 lim[3];
    int nlim = 0;
    v3 U[4], W[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const QgJointC& J = P.joint[leg][k];
        if (J.limited) {
            float dlo = S.q[k] - J.lo, dhi = J.hi - S.q[k];
            float dist = 0.f, sg = 0.f;
            if (dlo < 0.f) { dist = dlo; sg = 1.f; }
            else if (dhi < 0.f) { dist = dhi; sg = -1.f; }
            if (sg != 0.f) {
                float imp = impedance(dist, P.lim_d0, P.lim_dmax, P.lim_width, P.lim_mid, P.lim_power);
                lim[nlim] = make_float4(sg * (float)(k + 1), 1.f / fmaxf(1e-15f, (1.f - imp) / imp * J.invw_dof),
                                        P.lim_B * (sg * S.qd[k]) + P.lim_K * imp * dist /* = -aref */, 0.f);
                nlim++;
            }
        }
    }
    const int nc = C.n;
    {
        float qdb[6] = {vB.x, vB.y, vB.z, S.om.x, S.om.y, S.om.z};
        twist(qdb, S.qd, sl, sa, U, W);
    }
#pragma unroll 1
    for (int c = 0; c < nc; ++c) {
        const float4 g4 = C.geo[c], p4 = C.par[c];
        v3 xc = V3(g4.x, g4.y, g4.z);
        int lev = __float_as_int(p4.w);
        float rv[4], jr[4] = {0.f, 0.f, 0.f, 0.f};
        contact_rows<CONE>(sel4(U, lev) + cross(sel4(W, lev), xc), tx, ty, up, p4.x, rv);
#pragma unroll
        for (int k = 0; k < NR; ++k) jr[k] = fmaf(p4.y, rv[k], (CONE && k > 0) ? 0.f : p4.z);  // = -aref_k
        C.jar[c] = st4(jr);
    }
    wc_add(wc, QG_C_NCON, nc);
    wc_add(wc, QG_C_NEFC, NR * nc + nlim);
    if (DEBUG) { st.ncon += nc; st.nefc += NR * nc + nlim; }

    // ---- phase machine around ONE arrow solve: 0 = unconstrained acceleration (H = M, rhs = qfrc_smooth),
    //      1 = Newton direction (H = M + J^T D J, rhs = -grad), 2 = implicit integration
    //      (H = M + h*diag, rhs = qfrc_smooth + qfrc_constraint)
    float ab[6], al[3], a0b[6], a0l[3], Mab[6], Mal[3];
    float Hll[6], Hbl[18], Hc[21], rb[6], rl[3], xb[6], xl[3];
#pragma unroll
    for (int i = 0; i < 6; ++i) { Hll[i] = Mll[i]; rb[i] = fsb[i]; }
#pragma unroll
    for (int i = 0; i < 18; ++i) Hbl[i] = Mbl[i];
#pragma unroll
    for (int i = 0; i < 21; ++i) Hc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) rl[i] = fsl[i];
    int phase = 0, iter = 0, nact_last = 0, nls = 0;
    float impr_est = 0.f;
    bool done = false;
#pragma unroll 1
    for (;;) {
#if QG_BLOCKSYNC >= 5
        // experiment: warps free-run through the solver loop (warp-uniform trip count only); the block re-aligns at the
        // substep barrier
        if (!__any_sync(0xffffffffu, !done)) break;
        if (done) continue;
        arrow_solve(Hll, Hbl, Hc, Mbb, rb, rl, qr, xb, xl);
#elif QG_BLOCKSYNC
        // block-uniform trip count: the warps of a block walk the solver code together, so that one
        // instruction-cache fill serves all of them (instruction fetch is the limiter of this kernel)
        if (!__syncthreads_or(!done)) break;
#if QG_BLOCKSYNC >= 3
        if (!done) arrow_solve(Hll, Hbl, Hc, Mbb, rb, rl, qr, xb, xl);
        __syncthreads();
        if (done) continue;
#else
        if (done) continue;
        arrow_solve(Hll, Hbl, Hc, Mbb, rb, rl, qr, xb, xl);
#endif
#else
        if (done) break;
        arrow_solve(Hll, Hbl, Hc, Mbb, rb, rl, qr, xb, xl);
#endif
        if (phase == 2) { done = true; continue; }
        bool conv = false;
        if (phase == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) { a0b[i] = xb[i]; ab[i] = xb[i]; Mab[i] = fsb[i]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) { a0l[i] = xl[i]; al[i] = xl[i]; Mal[i] = fsl[i]; }
            if (nefc == 0) conv = true;
            else {
                // warm start: compare the cost at qacc_warmstart and at qacc_smooth (mj_fwdConstraint)
                v3 wlB = tmul(Rb, sb3(SR, SB_WLIN));
                float wb[6] = {wlB.x, wlB.y, wlB.z, SBW(SR, SB_WANG), SBW(SR, SB_WANG + 1), SBW(SR, SB_WANG + 2)};
                float wj[3] = {SLW(SR, SL_WJ), SLW(SR, SL_WJ + 1), SLW(SR, SL_WJ + 2)};
                float cw = 0.f, cs = 0.f;
                v3 U2[4], W2[4];
                twist(wb, wj, sl, sa, U, W);
                twist(a0b, a0l, sl, sa, U2, W2);
#pragma unroll 1
                for (int c = 0; c < nc; ++c) {
                    const float4 g4 = C.geo[c], p4 = C.par[c];
                    v3 xc = V3(g4.x, g4.y, g4.z);
                    int lev = __float_as_int(p4.w);
                    float rw[4], rs[4], jb[4], D = g4.w;
                    ld4(C.jar[c], jb);
                    contact_rows<CONE>(sel4(U, lev) + cross(sel4(W, lev), xc), tx, ty, up, p4.x, rw);
                    contact_rows<CONE>(sel4(U2, lev) + cross(sel4(W2, lev), xc), tx, ty, up, p4.x, rs);
                    if (CONE) { rw[3] = 0.f; rs[3] = 0.f; }
#pragma unroll
                    for (int k = 0; k < NR; ++k) {
                        float base = jb[k];
                        float jw = rw[k] + base, js = rs[k] + base;
                        rw[k] = jw; rs[k] = js;
                        if (!CONE) {
                            cw += (jw < 0.f) ? 0.5f * D * jw * jw : 0.f;
                            cs += (js < 0.f) ? 0.5f * D * js * js : 0.f;
                        }
                    }
                    C.jar[c] = st4(rw);   // at the warm start
                    C.jv[c] = st4(rs);    // at qacc_smooth (parked here until the cheaper point is known)
                    if (CONE) {
                        cw += ell_eval(rw[0], rw[1], rw[2], p4.x, mus, D, D * impr).cost;
                        cs += ell_eval(rs[0], rs[1], rs[2], p4.x, mus, D, D * impr).cost;
                    }
                }
#pragma unroll 1
                for (int s = 0; s < nlim; ++s) {
                    float4 r = lim[s];
                    const int k = (int)fabsf(r.x) - 1;
                    const float sg = r.x < 0.f ? -1.f : 1.f;
                    float jw = fmaf(sg, sel3(wj, k), r.z), js = fmaf(sg, sel3(a0l, k), r.z);
                    cw += (jw < 0.f) ? 0.5f * r.y * jw * jw : 0.f;
                    cs += (js < 0.f) ? 0.5f * r.y * js * js : 0.f;
                    lim[s] = make_float4(r.x, r.y, jw, js);   // .w parks the value at qacc_smooth
                }
                // M w: lane-local part now, base rows from the same reduction that sums the two costs
                float Mwb[6], Mwl[3], Mwp[6];
                arrow_matvec_local(Mll, Mbl, wb, wj, Mwp, Mwl);
                float gl = 0.f, gb = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) gl += 0.5f * (Mwl[k] - fsl[k]) * (wj[k] - a0l[k]);
                qr_put(qr, 0, cw + gl); qr_put(qr, 1, cs);
#pragma unroll
                for (int r = 0; r < 6; ++r) qr_put(qr, 2 + r, Mwp[r]);
                qr_sync(qr);
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    float t = 0.f;
#pragma unroll
                    for (int k = 0; k < 6; ++k) t = fmaf(Mbb[r >= k ? IX6(r, k) : IX6(k, r)], wb[k], t);
                    Mwb[r] = t + qr_get(qr, 2 + r);
                    gb += 0.5f * (Mwb[r] - fsb[r]) * (wb[r] - a0b[r]);
                }
                float cost_w = qr_get(qr, 0) + gb, cost_s = qr_get(qr, 1);
                qr_sync(qr);
                if (cost_w < cost_s) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) { ab[i] = wb[i]; Mab[i] = Mwb[i]; }
#pragma unroll
                    for (int i = 0; i < 3; ++i) { al[i] = wj[i]; Mal[i] = Mwl[i]; }
                } else {
#pragma unroll 1
                    for (int c = 0; c < nc; ++c) C.jar[c] = C.jv[c];
#pragma unroll 1
                    for (int s = 0; s < nlim; ++s) { float4 r = lim[s]; r.z = r.w; lim[s] = r; }
                }
            }
        } else {
            // ---- (xb, xl) is the Newton direction: exact line search on the convex piecewise quadratic
            // M v: the lane-local part now, the base rows after the line search's own reduction (their six partial sums
            // ride in the same shared-memory round trip instead of a separate one)
            float Mvb[6], Mvl[3], Mvp[6];
            arrow_matvec_local(Mll, Mbl, xb, xl, Mvp, Mvl);
            float q1l = 0.f, q2l = 0.f, q1b = 0.f, q2b = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) { q1l += xl[k] * (Mal[k] - fsl[k]); q2l += 0.5f * xl[k] * Mvl[k]; }
            twist(xb, xl, sl, sa, U, W);
            // first trial alpha = 1 (the exact minimiser when no row changes state along the step)
            float e1 = 0.f, e2 = 0.f, z1 = 0.f, z2 = 0.f;
            int flips = 0;
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
                const float4 g4 = C.geo[c], p4 = C.par[c];
                v3 xc = V3(g4.x, g4.y, g4.z);
                int lev = __float_as_int(p4.w);
                float rv[4], jb[4], D = g4.w;
                ld4(C.jar[c], jb);
                contact_rows<CONE>(sel4(U, lev) + cross(sel4(W, lev), xc), tx, ty, up, p4.x, rv);
                C.jv[c] = st4(rv);
                if (CONE) {
                    float a0 = jb[0], a1 = jb[1], a2 = jb[2];
                    int zo0 = ell_ls_acc(a0, a1, a2, rv[0], rv[1], rv[2], p4.x, mus, D, D * impr, z1, z2);
                    int zo1 = ell_ls_acc(a0 + rv[0], a1 + rv[1], a2 + rv[2], rv[0], rv[1], rv[2], p4.x, mus, D, D * impr, e1, e2);
                    flips += (zo0 != zo1 || zo0 == 2) ? 1 : 0;   // the cone surface is not quadratic: iterate
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float jv = rv[k], j0 = jb[k], j1 = j0 + jv;
                        if (j0 < 0.f) { z1 = fmaf(D * jv, j0, z1); z2 = fmaf(D * jv, jv, z2); }
                        if (j1 < 0.f) { e1 = fmaf(D * jv, j1, e1); e2 = fmaf(D * jv, jv, e2); }
                        flips += ((j0 < 0.f) != (j1 < 0.f)) ? 1 : 0;
                    }
                }
            }
#pragma unroll 1
            for (int s = 0; s < nlim; ++s) {
                float4 r = lim[s];
                const int k = (int)fabsf(r.x) - 1;
                const float jv = (r.x < 0.f ? -1.f : 1.f) * sel3(xl, k), j0 = r.z, j1 = j0 + jv;
                if (j0 < 0.f) { z1 = fmaf(r.y * jv, j0, z1); z2 = fmaf(r.y * jv, jv, z2); }
                if (j1 < 0.f) { e1 = fmaf(r.y * jv, j1, e1); e2 = fmaf(r.y * jv, jv, e2); }
                flips += ((j0 < 0.f) != (j1 < 0.f)) ? 1 : 0;
                r.w = jv;
                lim[s] = r;
            }
            qr_put(qr, 0, q1l); qr_put(qr, 1, q2l); qr_put(qr, 2, (float)flips);
            qr_put(qr, 3, z1); qr_put(qr, 4, e1); qr_put(qr, 5, e2);
#pragma unroll
            for (int r = 0; r < 6; ++r) qr_put(qr, 6 + r, Mvp[r]);
            qr_sync(qr);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < 6; ++k) t = fmaf(Mbb[r >= k ? IX6(r, k) : IX6(k, r)], xb[k], t);
                Mvb[r] = t + qr_get(qr, 6 + r);
                q1b += xb[r] * (Mab[r] - fsb[r]);
                q2b += 0.5f * xb[r] * Mvb[r];
            }
            const float q1 = qr_get(qr, 0) + q1b, q2 = qr_get(qr, 1) + q2b;
            flips = (int)qr_get(qr, 2);
            const float z1s = qr_get(qr, 3), e1s = qr_get(qr, 4), e2s = qr_get(qr, 5);
            qr_sync(qr);
            nls++;
            float alpha = 1.f;
            const float d10 = z1s + q1;                                // derivative at 0 (< 0: descent direction)
            if (flips != 0) {
                float d1 = e1s + q1 + 2.f * q2, d2 = e2s + 2.f * q2;
                float gtol = 1e-4f * fabsf(d10);
                float lo = 0.f, hi = -1.f;
                if (d10 >= 0.f) alpha = 0.f;
                else {
#pragma unroll 1
                    for (int it = 0; it < ls_iter; ++it) {
                        if (fabsf(d1) < gtol) break;
                        if (d1 < 0.f) lo = alpha; else hi = alpha;
                        float an = alpha - d1 / d2;
                        if (an <= lo || (hi > 0.f && an >= hi)) an = hi > 0.f ? 0.5f * (lo + hi) : 2.f * alpha;
                        alpha = an;
                        e1 = 0.f; e2 = 0.f;
#pragma unroll 1
                        for (int c = 0; c < nc; ++c) {
                            const float D = C.geo[c].w;
                            float jb[4], vb[4];
                            ld4(C.jar[c], jb);
                            ld4(C.jv[c], vb);
                            if (CONE) {
                                float v0 = vb[0], v1 = vb[1], v2 = vb[2];
                                ell_ls_acc(fmaf(alpha, v0, jb[0]), fmaf(alpha, v1, jb[1]), fmaf(alpha, v2, jb[2]),
                                           v0, v1, v2, C.par[c].x, mus, D, D * impr, e1, e2);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    float jv = vb[k], xx = fmaf(alpha, jv, jb[k]);
                                    if (xx < 0.f) { e1 = fmaf(D * jv, xx, e1); e2 = fmaf(D * jv, jv, e2); }
                                }
                            }
                        }
#pragma unroll 1
                        for (int s = 0; s < nlim; ++s) {
                            const float4 r = lim[s];
                            float xx = fmaf(alpha, r.w, r.z);
                            if (xx < 0.f) { e1 = fmaf(r.y * r.w, xx, e1); e2 = fmaf(r.y * r.w, r.w, e2); }
                        }
                        qr_put(qr, 0, e1); qr_put(qr, 1, e2);
                        qr_sync(qr);
                        d1 = qr_get(qr, 0) + q1 + 2.f * alpha * q2;
                        d2 = qr_get(qr, 1) + 2.f * q2;
                        qr_sync(qr);
                        nls++;
                    }
                }
            }
            if (alpha == 0.f) conv = true;
#pragma unroll
            for (int r = 0; r < 6; ++r) { ab[r] = fmaf(alpha, xb[r], ab[r]); Mab[r] = fmaf(alpha, Mvb[r], Mab[r]); }
#pragma unroll
            for (int k = 0; k < 3; ++k) { al[k] = fmaf(alpha, xl[k], al[k]); Mal[k] = fmaf(alpha, Mvl[k], Mal[k]); }
#pragma unroll 1
            for (int s = 0; s < nlim; ++s) { float4 r = lim[s]; r.z = fmaf(alpha, r.w, r.z); lim[s] = r; }
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
                const float4 a4 = C.jar[c], v4 = C.jv[c];
                C.jar[c] = make_float4(fmaf(alpha, v4.x, a4.x), fmaf(alpha, v4.y, a4.y), fmaf(alpha, v4.z, a4.z), fmaf(alpha, v4.w, a4.w));
            }
            iter++;
            if (flips == 0) conv = true;  // the quadratic model was exact along the whole step: this is the optimum
            // decrease of the cost along the step from the line-search model (-1/2 alpha d1(0) for an exact search on a
            // quadratic).  In fp32 the difference of two cost values is not resolvable once |cost| * 1e-7 > tolerance.
            impr_est = -0.5f * alpha * d10;
        }

        float gb[6], gl[3];
        if (nefc > 0) {
            // ---- constraint update at the current point: forces, J^T f, gradient  (the cost value itself is not needed:
            //      convergence uses the line-search model's decrease, see impr_est)
            v3 Fb = V3(0, 0, 0), Nb = V3(0, 0, 0);
            float tau[3] = {0.f, 0.f, 0.f};
            nact_last = 0;
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
                const float4 g4 = C.geo[c], p4 = C.par[c], j4 = C.jar[c];
                float D = g4.w, mu = p4.x;
                v3 xc = V3(g4.x, g4.y, g4.z);
                int lev = __float_as_int(p4.w);
                v3 fc;
                if (CONE) {
                    EllZ z = ell_eval(j4.x, j4.y, j4.z, mu, mus, D, D * impr);
                    nact_last += z.zone ? 3 : 0;
                    fc = fma3(z.f0, up, fma3(z.f1, ty, (-z.f2) * tx));   // rows: n = up, t1 = ty, t2 = -tx
                } else {
                    float j0 = j4.x, j1 = j4.y, j2 = j4.z, j3 = j4.w;
                    float f0 = j0 < 0.f ? -D * j0 : 0.f, f1 = j1 < 0.f ? -D * j1 : 0.f;
                    float f2 = j2 < 0.f ? -D * j2 : 0.f, f3 = j3 < 0.f ? -D * j3 : 0.f;
                    nact_last += (j0 < 0.f) + (j1 < 0.f) + (j2 < 0.f) + (j3 < 0.f);
                    // force vector in B: sum f_k w_k,  w = up +- mu*ty, up -+ mu*tx
                    fc = fma3(f0 + f1 + f2 + f3, up, fma3(mu * (f0 - f1), ty, (-mu * (f2 - f3)) * tx));
                }
                v3 nn = cross(xc, fc);
                Fb += fc;
                Nb += nn;
                tau[0] += lev >= 1 ? dot(sl[0], fc) + dot(sa[0], nn) : 0.f;
                tau[1] += lev >= 2 ? dot(sl[1], fc) + dot(sa[1], nn) : 0.f;
                tau[2] += lev >= 3 ? dot(sl[2], fc) + dot(sa[2], nn) : 0.f;
            }
#pragma unroll 1
            for (int s = 0; s < nlim; ++s) {
                const float4 r = lim[s];
                if (r.z < 0.f) {
                    const int k = (int)fabsf(r.x) - 1;
                    const float t = (r.x < 0.f ? 1.f : -1.f) * r.y * r.z;   // sign * f,  f = -D * jar
                    nact_last++;
                    tau[0] += k == 0 ? t : 0.f; tau[1] += k == 1 ? t : 0.f; tau[2] += k == 2 ? t : 0.f;
                }
            }
            qr_put(qr, 1, Fb.x); qr_put(qr, 2, Fb.y); qr_put(qr, 3, Fb.z);
            qr_put(qr, 4, Nb.x); qr_put(qr, 5, Nb.y); qr_put(qr, 6, Nb.z);
            float g2 = 0.f, g2b = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) { gl[k] = Mal[k] - fsl[k] - tau[k]; g2 += gl[k] * gl[k]; }
            qr_put(qr, 7, g2);
            qr_sync(qr);
            float fcb[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) fcb[r] = qr_get(qr, 1 + r);
            g2 = qr_get(qr, 7);
            qr_sync(qr);
#pragma unroll
            for (int r = 0; r < 6; ++r) { gb[r] = Mab[r] - fsb[r] - fcb[r]; g2b += gb[r] * gb[r]; }
            if (phase == 1 && !conv) {
                float gradient = P.scale * sqrtf(g2 + g2b);
                float improvement = P.scale * impr_est;
                conv = improvement < P.tol || gradient < P.tol || iter >= max_iter;
            }
        }

        if (conv) {
            // ---- next solve: implicit integration  (M + h*diag) qacc+ = qfrc_smooth + qfrc_constraint
#pragma unroll
            for (int i = 0; i < 6; ++i) Hll[i] = Mll[i];
#pragma unroll
            for (int k = 0; k < 3; ++k) {   // implicit-integrator diagonal: damping + servo kv (unless force-clamped)
                const QgJointC& J = P.joint[leg][k];
                float dk = J.damping - (((unclamped >> k) & 1) ? J.gear * J.gear * J.b2 : 0.f);
                Hll[k == 0 ? 0 : (k == 1 ? 3 : 5)] += h * dk;
            }
#pragma unroll
            for (int i = 0; i < 18; ++i) Hbl[i] = Mbl[i];
#pragma unroll
            for (int i = 0; i < 21; ++i) Hc[i] = 0.f;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                Hc[IX6(r, r)] = (leg == 0) ? h * P.base_damp[r] : 0.f;
                // qfrc_smooth + qfrc_constraint = M a - gradient at the solver's final point (no copy of the force kept)
                rb[r] = (nefc > 0) ? Mab[r] - gb[r] : fsb[r];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) rl[k] = (nefc > 0) ? Mal[k] - gl[k] : fsl[k];
            phase = 2;
        } else {
            // ---- next solve: Newton direction.  H = M + J^T D J over the active rows
#pragma unroll
            for (int i = 0; i < 6; ++i) Hll[i] = Mll[i];
#pragma unroll
            for (int i = 0; i < 18; ++i) Hbl[i] = Mbl[i];
#pragma unroll
            for (int i = 0; i < 21; ++i) Hc[i] = 0.f;
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
                const float4 g4 = C.geo[c], p4 = C.par[c], j4 = C.jar[c];
                float D = g4.w, mu = p4.x;
                // W = 3x3 stiffness of the contact in the basis (e0, e1, e2) = (up, ty, -tx): w00 w11 w22 w01 w02 w12
                float w00, w11, w22, w01, w02, w12;
                if (CONE) {
                    EllZ z = ell_eval(j4.x, j4.y, j4.z, mu, mus, D, D * impr);
                    if (z.zone == 0) continue;
                    if (z.zone == 1) { w00 = D; w11 = w22 = D * impr; w01 = w02 = w12 = 0.f; }
                    else {
                        float iT = 1.f / z.T, e = z.N - z.mu * z.T;
                        float a = z.Dm * z.mu * z.mu * iT * iT + z.Dm * e * z.mu * iT * iT * iT;   // coefficient of U_j U_k
                        float b = -z.Dm * e * z.mu * iT;                                       // coefficient of delta_jk
                        float sn = z.mu, stt = mu;                                             // dU/djar = diag(mu_reg, fri, fri)
                        w00 = sn * sn * z.Dm;
                        w01 = sn * stt * (-z.Dm * z.mu * z.U1 * iT);
                        w02 = sn * stt * (-z.Dm * z.mu * z.U2 * iT);
                        w11 = stt * stt * (a * z.U1 * z.U1 + b);
                        w22 = stt * stt * (a * z.U2 * z.U2 + b);
                        w12 = stt * stt * (a * z.U1 * z.U2);
                    }
                } else {
                    float a0 = j4.x < 0.f ? 1.f : 0.f, a1 = j4.y < 0.f ? 1.f : 0.f;
                    float a2 = j4.z < 0.f ? 1.f : 0.f, a3 = j4.w < 0.f ? 1.f : 0.f;
                    float na = a0 + a1 + a2 + a3;
                    if (na == 0.f) continue;
                    // rows w0 = e0 + mu e1, w1 = e0 - mu e1, w2 = e0 + mu e2, w3 = e0 - mu e2
                    w00 = D * na; w11 = D * mu * mu * (a0 + a1); w22 = D * mu * mu * (a2 + a3);
                    w01 = D * mu * (a0 - a1); w02 = D * mu * (a2 - a3); w12 = 0.f;
                }
                v3 xc = V3(g4.x, g4.y, g4.z);
                int lev = __float_as_int(p4.w);
                // Jacobian columns of the leg joints at the contact point (zero above the contact's link)
                v3 jc0 = lev >= 1 ? sl[0] + cross(sa[0], xc) : V3(0, 0, 0);
                v3 jc1 = lev >= 2 ? sl[1] + cross(sa[1], xc) : V3(0, 0, 0);
                v3 jc2 = lev >= 3 ? sl[2] + cross(sa[2], xc) : V3(0, 0, 0);
                // stiffness in B coordinates, formed once: Wb = E^T W E = sum_i e_i g_i^T with g_i = sum_j W_ij e_j
                v3 g0 = fma3(w00, up, fma3(w01, ty, (-w02) * tx));
                v3 g1 = fma3(w01, up, fma3(w11, ty, (-w12) * tx));
                v3 g2 = fma3(w02, up, fma3(w12, ty, (-w22) * tx));
                s3 Wb;
                Wb.xx = fmaf(up.x, g0.x, fmaf(ty.x, g1.x, -tx.x * g2.x));
                Wb.yy = fmaf(up.y, g0.y, fmaf(ty.y, g1.y, -tx.y * g2.y));
                Wb.zz = fmaf(up.z, g0.z, fmaf(ty.z, g1.z, -tx.z * g2.z));
                Wb.xy = fmaf(up.x, g0.y, fmaf(ty.x, g1.y, -tx.x * g2.y));
                Wb.xz = fmaf(up.x, g0.z, fmaf(ty.x, g1.z, -tx.x * g2.z));
                Wb.yz = fmaf(up.y, g0.z, fmaf(ty.y, g1.z, -tx.y * g2.z));
                const v3 Wx = V3(Wb.xx, Wb.xy, Wb.xz), Wy = V3(Wb.xy, Wb.yy, Wb.yz), Wz = V3(Wb.xz, Wb.yz, Wb.zz);
                const v3 q0 = mul(Wb, jc0), q1v = mul(Wb, jc1), q2v = mul(Wb, jc2);
                Hll[0] += dot(jc0, q0); Hll[1] += dot(jc0, q1v); Hll[2] += dot(jc0, q2v);
                Hll[3] += dot(jc1, q1v); Hll[4] += dot(jc1, q2v); Hll[5] += dot(jc2, q2v);
                v3 x0 = cross(xc, q0), x1 = cross(xc, q1v), x2 = cross(xc, q2v);
                Hbl[0] += q0.x; Hbl[1] += q1v.x; Hbl[2] += q2v.x;
                Hbl[3] += q0.y; Hbl[4] += q1v.y; Hbl[5] += q2v.y;
                Hbl[6] += q0.z; Hbl[7] += q1v.z; Hbl[8] += q2v.z;
                Hbl[9] += x0.x; Hbl[10] += x1.x; Hbl[11] += x2.x;
                Hbl[12] += x0.y; Hbl[13] += x1.y; Hbl[14] += x2.y;
                Hbl[15] += x0.z; Hbl[16] += x1.z; Hbl[17] += x2.z;
                // base block  [[W, -W X], [X W, -X W X]]
                Hc[IX6(0, 0)] += Wx.x; Hc[IX6(1, 0)] += Wy.x; Hc[IX6(1, 1)] += Wy.y;
                Hc[IX6(2, 0)] += Wz.x; Hc[IX6(2, 1)] += Wz.y; Hc[IX6(2, 2)] += Wz.z;
                v3 A0 = cross(xc, Wx), A1 = cross(xc, Wy), A2 = cross(xc, Wz);  // columns of X W
                Hc[IX6(3, 0)] += A0.x; Hc[IX6(3, 1)] += A1.x; Hc[IX6(3, 2)] += A2.x;
                Hc[IX6(4, 0)] += A0.y; Hc[IX6(4, 1)] += A1.y; Hc[IX6(4, 2)] += A2.y;
                Hc[IX6(5, 0)] += A0.z; Hc[IX6(5, 1)] += A1.z; Hc[IX6(5, 2)] += A2.z;
                // -X W X : row i = cross(xc, row_i(XW)),  row_i(XW) = (A0[i], A1[i], A2[i])
                v3 r3 = cross(xc, V3(A0.x, A1.x, A2.x)), r4 = cross(xc, V3(A0.y, A1.y, A2.y)), r5 = cross(xc, V3(A0.z, A1.z, A2.z));
                Hc[IX6(3, 3)] += r3.x; Hc[IX6(4, 3)] += r4.x; Hc[IX6(4, 4)] += r4.y;
                Hc[IX6(5, 3)] += r5.x; Hc[IX6(5, 4)] += r5.y; Hc[IX6(5, 5)] += r5.z;
            }
#pragma unroll 1
            for (int s = 0; s < nlim; ++s) {
                const float4 r = lim[s];
                if (r.z < 0.f) {
                    const int k = (int)fabsf(r.x) - 1;
                    Hll[0] += k == 0 ? r.y : 0.f; Hll[3] += k == 1 ? r.y : 0.f; Hll[5] += k == 2 ? r.y : 0.f;
                }
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) rb[r] = -gb[r];
#pragma unroll
            for (int k = 0; k < 3; ++k) rl[k] = -gl[k];
            phase = 1;
        }
    }
    wc_add(wc, QG_C_NITER, leg == 0 ? iter : 0);
    wc_add(wc, QG_C_NLS, leg == 0 ? nls : 0);
    wc_add(wc, QG_C_NACT, nact_last);
    if (DEBUG) { st.niter += (leg == 0) ? iter : 0; st.nls += nls; }
    st.last_ls = nls;

    // ---- the state again, from shared memory (nothing of it was held across the solver)
    LaneState E;
    state_load(SR, E);
    const v3 vB2 = tmul(Rb, E.vw);
    // ---- sensors of this forward pass (pre-integration state, solver qacc)
    if (want_sensors) {
        so.jq[0] = E.q[0]; so.jq[1] = E.q[1]; so.jq[2] = E.q[2];
        so.acc = V3(ab[0], ab[1], ab[2]) - tmul(Rb, ld3(P.grav));
        so.gyro = E.om;
        so.pos = E.pb;
        so.linvel = E.vw;
        so.xaxis = col0(Rb);
        so.zaxis = col2(Rb);
        so.vel = vB2;
    }
    if (DEBUG) {
        v3 aw = mul(Rb, V3(ab[0], ab[1], ab[2])), a0w = mul(Rb, V3(a0b[0], a0b[1], a0b[2]));
        if (leg == 0) {
            if (dbg.qacc) {
                float* o = dbg.qacc + env * 18;
                o[0] = aw.x; o[1] = aw.y; o[2] = aw.z; o[3] = ab[3]; o[4] = ab[4]; o[5] = ab[5];
            }
            if (dbg.qacc_smooth) {
                float* o = dbg.qacc_smooth + env * 18;
                o[0] = a0w.x; o[1] = a0w.y; o[2] = a0w.z; o[3] = a0b[3]; o[4] = a0b[4]; o[5] = a0b[5];
            }
            if (dbg.qfrc_bias) {
                float* o = dbg.qfrc_bias + env * 18;
                o[0] = fsum.x; o[1] = fsum.y; o[2] = fsum.z; o[3] = nsum.x; o[4] = nsum.y; o[5] = nsum.z;
            }
            if (dbg.M) {
                float* o = dbg.M + (size_t)env * 324;
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j <= i; ++j) { o[i * 18 + j] = Mbb[IX6(i, j)]; o[j * 18 + i] = Mbb[IX6(i, j)]; }
            }
        }
        for (int k = 0; k < 3; ++k) {
            int d = 6 + 3 * leg + k;
            if (dbg.qacc) dbg.qacc[env * 18 + d] = al[k];
            if (dbg.qacc_smooth) dbg.qacc_smooth[env * 18 + d] = a0l[k];
            if (dbg.qfrc_bias) dbg.qfrc_bias[env * 18 + d] = bias_l[k];
            if (dbg.M) {
                float* o = dbg.M + (size_t)env * 324;
                for (int r = 0; r < 6; ++r) { o[r * 18 + d] = Mbl[r * 3 + k]; o[d * 18 + r] = Mbl[r * 3 + k]; }
                for (int j = 0; j < 3; ++j) {
                    int a = k < j ? k : j, b = k < j ? j : k;
                    int idx = (a == 0) ? b : (a == 1 ? 2 + b : 5);
                    o[d * 18 + 6 + 3 * leg + j] = Mll[idx];
                }
                for (int l2 = 0; l2 < 4; ++l2)
                    if (l2 != leg)
                        for (int j = 0; j < 3; ++j) o[d * 18 + 6 + 3 * l2 + j] = 0.f;
            }
        }
    }

    // ---- semi-implicit update with qacc+ = (xb, xl); warm start for the next step = solver acceleration
    const v3 wl = mul(Rb, V3(ab[0], ab[1], ab[2]));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        SLW(SR, SL_WJ + k) = al[k];
        {   // activation dynamics: act_dot = (clamp(ctrl) - act) / tau, exact first-order filter over h
            const QgJointC& J = P.joint[leg][k];
            if (J.has_act && J.has_dyn) {
                float c = E.ctrl[k];
                if (J.ctrl_limited) c = fminf(fmaxf(c, J.ctrl_lo), J.ctrl_hi);
                SLW(SR, SL_ACT + k) = fmaf((c - E.act[k]) * J.inv_tau, J.act_fac, E.act[k]);
            }
        }
        const float qd = fmaf(h, xl[k], E.qd[k]);
        SLW(SR, SL_QD + k) = qd;
        SLW(SR, SL_Q + k) = fmaf(h, qd, E.q[k]);
    }
    const v3 vBn = fma3(h, V3(xb[0], xb[1], xb[2]), vB2);
    const v3 vwn = mul(Rb, vBn);
    const v3 omn = fma3(h, V3(xb[3], xb[4], xb[5]), E.om);
    const v3 pbn = fma3(h, vwn, E.pb);
    float wn = sqrtf(dot(omn, omn));
    float rw = 1.f, rx = 0.f, ry = 0.f, rz = 0.f;
    if (wn > 1e-15f) {
        float2 sc = sincos_ni(0.5f * h * wn);
        float s = sc.x / wn;
        rw = sc.y; rx = omn.x * s; ry = omn.y * s; rz = omn.z * s;
    }
    // the normalised quaternion of this step's forward pass, recomputed from the stored one (same arithmetic, same bits)
    const float qn2 = 1.f / sqrtf(E.qw * E.qw + E.qx * E.qx + E.qy * E.qy + E.qz * E.qz);
    const float w2 = E.qw * qn2, x2 = E.qx * qn2, y2 = E.qy * qn2, z2 = E.qz * qn2;
    const double tnew = sb_time(SR) + P.timestep_d;
    __syncwarp(qr.qm);           // every lane of the quad has read the old base state
    if (leg == 0) {              // the four lanes hold identical values: one of them writes the base words
        SBW(SR, SB_POS) = pbn.x; SBW(SR, SB_POS + 1) = pbn.y; SBW(SR, SB_POS + 2) = pbn.z;
        SBW(SR, SB_QUAT) = w2 * rw - x2 * rx - y2 * ry - z2 * rz;
        SBW(SR, SB_QUAT + 1) = w2 * rx + x2 * rw + y2 * rz - z2 * ry;
        SBW(SR, SB_QUAT + 2) = w2 * ry - x2 * rz + y2 * rw + z2 * rx;
        SBW(SR, SB_QUAT + 3) = w2 * rz + x2 * ry - y2 * rx + z2 * rw;
        SBW(SR, SB_VLIN) = vwn.x; SBW(SR, SB_VLIN + 1) = vwn.y; SBW(SR, SB_VLIN + 2) = vwn.z;
        SBW(SR, SB_VANG) = omn.x; SBW(SR, SB_VANG + 1) = omn.y; SBW(SR, SB_VANG + 2) = omn.z;
        SBW(SR, SB_WLIN) = wl.x; SBW(SR, SB_WLIN + 1) = wl.y; SBW(SR, SB_WLIN + 2) = wl.z;
        SBW(SR, SB_WANG) = ab[3]; SBW(SR, SB_WANG + 1) = ab[4]; SBW(SR, SB_WANG + 2) = ab[5];
        sb_set_time(SR, tnew);
    }
    __syncwarp(qr.qm);
}
