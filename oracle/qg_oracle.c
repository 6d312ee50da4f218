/*
 * qg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * float64 CPU restatement of the reference's hot path
 *     QuadrupedEnv.step()  =  frame_skip x mujoco.mj_step  ->  sensordata
 *     (/root/reference/src/envs/quadruped.py:153-182, physics call at :165, obs copy at :141-143)
 * for the model class of /root/reference/src/models/quadruped/{quadruped,scene}.xml.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (quadruped_gym_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the path lives in the third-party PyPI wheel `mujoco`
 * (unpinned in /root/reference/requirements.txt:1), which is neither vendored in the reference
 * nor installable here (no network), and the reference ships no tests, golden vectors or
 * fixtures for this path.  This file therefore restates MuJoCo's *published* algorithm
 * (Computation chapter: kinematics, composite rigid body, recursive Newton-Euler, soft
 * constraints with impedance/reference acceleration, pyramidal friction cones, primal Newton
 * solver, implicitfast integration) from memory, written independently of the CUDA kernels:
 * generic tree loops in world coordinates with 6-D spatial algebra about the subtree centre of
 * mass and dense nv x nv matrices -- whereas the kernels use a leg-per-lane body-frame
 * formulation with an arrow-structured factorisation.  Agreement between the two is the parity
 * check.  tests/test_mujoco_gated.py compares this file against the real mj_step wherever
 * `mujoco` imports.
 *
 * Stage order follows mj_step: forward{ kinematics, comPos, crb, factorM, collision,
 * makeConstraint, comVel, passive, referenceConstraint, rne, actuation, acceleration,
 * constraint solve, sensors } -> implicitfast integrate.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qg_oracle.h"

#define MINVAL 1e-15

/* ------------------------------------------------------------------ blob parsing */

typedef struct {
    const char* name;
    int dtype;
    uint32_t count;
    const void* data;
} section_t;

static const void* find_sec(const uint8_t* buf, size_t n, const char* name, int dtype, uint32_t* count) {
    uint32_t nsec;
    size_t off = 16;
    if (n < 16 || memcmp(buf, "QGBLOB01", 8) != 0) return NULL;
    memcpy(&nsec, buf + 8, 4);
    for (uint32_t s = 0; s < nsec; s++) {
        char nm[17];
        uint32_t code, cnt;
        if (off + 24 > n) return NULL;
        memcpy(nm, buf + off, 16);
        nm[16] = 0;
        memcpy(&code, buf + off + 16, 4);
        memcpy(&cnt, buf + off + 20, 4);
        off += 24;
        size_t nbytes = (size_t)cnt * (code == 1 ? 8 : 4);
        if (off + nbytes > n) return NULL;
        if (strcmp(nm, name) == 0) {
            if ((int)code != dtype) return NULL;
            *count = cnt;
            return buf + off;
        }
        off += nbytes + ((8 - nbytes % 8) % 8);
    }
    return NULL;
}

#define GETF(field, secname, expect)                                                   \
    do {                                                                               \
        uint32_t c_;                                                                   \
        const void* p_ = find_sec(buf, n, secname, 1, &c_);                            \
        if (!p_ || (int)c_ != (int)(expect) || (expect) > (int)(sizeof(field) / 8)) {  \
            fprintf(stderr, "qgo_model_load: bad section %s\n", secname);              \
            free(m);                                                                   \
            return -2;                                                                 \
        }                                                                              \
        memcpy(field, p_, (size_t)c_ * 8);                                             \
    } while (0)
#define GETI(field, secname, expect)                                                   \
    do {                                                                               \
        uint32_t c_;                                                                   \
        const void* p_ = find_sec(buf, n, secname, 2, &c_);                            \
        if (!p_ || (int)c_ != (int)(expect) || (expect) > (int)(sizeof(field) / 4)) {  \
            fprintf(stderr, "qgo_model_load: bad section %s\n", secname);              \
            free(m);                                                                   \
            return -2;                                                                 \
        }                                                                              \
        memcpy(field, p_, (size_t)c_ * 4);                                             \
    } while (0)

int qgo_model_load(const void* blob, size_t n, qgo_model** out) {
    const uint8_t* buf = (const uint8_t*)blob;
    qgo_model* m = (qgo_model*)calloc(1, sizeof(qgo_model));
    int sizes[8];
    uint32_t c;
    const void* p = find_sec(buf, n, "sizes", 2, &c);
    if (!m) return -1;
    if (!p || c != 8) {
        free(m);
        return -2;
    }
    memcpy(sizes, p, 32);
    m->nq = sizes[0]; m->nv = sizes[1]; m->nu = sizes[2]; m->nbody = sizes[3];
    m->njnt = sizes[4]; m->ngeom = sizes[5]; m->nmesh = sizes[6]; m->nsensordata = sizes[7];
    if (m->nq > QGO_MAXNQ || m->nv > QGO_MAXNV || m->nu > QGO_MAXNU || m->nbody > QGO_MAXNBODY ||
        m->njnt > QGO_MAXNBODY || m->ngeom > QGO_MAXNGEOM || m->nmesh > QGO_MAXNMESH ||
        m->nsensordata != QGO_NSENSORDATA) {
        free(m);
        return -3;
    }
    GETF(m->opt_f, "opt_f", 9);
    GETI(m->opt_i, "opt_i", 5);
    GETI(m->body_parent, "body_parent", m->nbody);
    GETF(m->body_pos, "body_pos", 3 * m->nbody);
    GETF(m->body_quat, "body_quat", 4 * m->nbody);
    GETF(m->body_mass, "body_mass", m->nbody);
    GETF(m->body_ipos, "body_ipos", 3 * m->nbody);
    GETF(m->body_inertia, "body_inertia", 6 * m->nbody);
    GETF(m->body_invweight0, "body_invweight0", 2 * m->nbody);
    GETI(m->jnt_type, "jnt_type", m->njnt);
    GETI(m->jnt_body, "jnt_body", m->njnt);
    GETI(m->jnt_qposadr, "jnt_qposadr", m->njnt);
    GETI(m->jnt_dofadr, "jnt_dofadr", m->njnt);
    GETF(m->jnt_axis, "jnt_axis", 3 * m->njnt);
    GETF(m->jnt_pos, "jnt_pos", 3 * m->njnt);
    GETF(m->jnt_range, "jnt_range", 2 * m->njnt);
    GETI(m->jnt_limited, "jnt_limited", m->njnt);
    GETF(m->jnt_solref, "jnt_solref", 2);
    GETF(m->jnt_solimp, "jnt_solimp", 5);
    GETF(m->qpos0, "qpos0", m->nq);
    GETF(m->dof_damping, "dof_damping", m->nv);
    GETF(m->dof_armature, "dof_armature", m->nv);
    GETF(m->dof_invweight0, "dof_invweight0", m->nv);
    GETI(m->dof_body, "dof_body", m->nv);
    GETI(m->act_dof, "act_dof", m->nu);
    GETF(m->act_gear, "act_gear", m->nu);
    GETF(m->act_gain, "act_gain", m->nu);
    GETF(m->act_bias, "act_bias", 3 * m->nu);
    GETF(m->act_tau, "act_tau", m->nu);
    GETF(m->act_ctrlrange, "act_ctrlrange", 2 * m->nu);
    GETI(m->act_ctrllimited, "act_ctrllimited", m->nu);
    GETF(m->act_frcrange, "act_frcrange", 2 * m->nu);
    GETI(m->act_frclimited, "act_frclimited", m->nu);
    GETI(m->geom_body, "geom_body", m->ngeom);
    GETF(m->geom_pos, "geom_pos", 3 * m->ngeom);
    GETF(m->geom_quat, "geom_quat", 4 * m->ngeom);
    GETI(m->geom_mesh, "geom_mesh", m->ngeom);
    GETF(m->geom_rbound, "geom_rbound", m->ngeom);
    GETF(m->geom_margin, "geom_margin", m->ngeom);
    GETF(m->geom_mu, "geom_mu", m->ngeom);
    GETF(m->geom_solref, "geom_solref", 2 * m->ngeom);
    GETF(m->geom_solimp, "geom_solimp", 5 * m->ngeom);
    GETI(m->mesh_vertadr, "mesh_vertadr", m->nmesh);
    GETI(m->mesh_vertnum, "mesh_vertnum", m->nmesh);
    GETI(m->mesh_edgeadr, "mesh_edgeadr", m->nmesh);
    m->nvert = m->mesh_vertadr[m->nmesh - 1] + m->mesh_vertnum[m->nmesh - 1];
    if (m->nvert > QGO_MAXVERT) {
        free(m);
        return -3;
    }
    GETF(m->mesh_vert, "mesh_vert", 3 * m->nvert);
    GETI(m->mesh_vert_edge, "mesh_vert_edge", m->nvert);
    p = find_sec(buf, n, "mesh_edge", 2, &c);
    if (!p || c > QGO_MAXEDGE) {
        free(m);
        return -2;
    }
    memcpy(m->mesh_edge, p, (size_t)c * 4);
    /* per-dof parent chain (dof_parentid) and per-body dof ranges */
    for (int b = 0; b < m->nbody; b++) m->body_dofadr[b] = -1, m->body_dofnum[b] = 0;
    for (int j = 0; j < m->njnt; j++) {
        int b = m->jnt_body[j], nd = m->jnt_type[j] == 0 ? 6 : 1;
        if (m->body_dofadr[b] < 0) m->body_dofadr[b] = m->jnt_dofadr[j];
        m->body_dofnum[b] += nd;
    }
    for (int d = 0; d < m->nv; d++) {
        int b = m->dof_body[d];
        if (d > m->body_dofadr[b]) {
            m->dof_parent[d] = d - 1;
        } else {
            int pb = m->body_parent[b];
            while (pb > 0 && m->body_dofnum[pb] == 0) pb = m->body_parent[pb];
            m->dof_parent[d] = pb > 0 ? m->body_dofadr[pb] + m->body_dofnum[pb] - 1 : -1;
        }
    }
    *out = m;
    return 0;
}

void qgo_model_free(qgo_model* m) { free(m); }

/* ------------------------------------------------------------------ small math */

static void quat2mat(double* R, const double* q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void mulquat(double* r, const double* a, const double* b) {
    double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                   a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                   a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                   a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
    memcpy(r, t, sizeof t);
}
static void normalize4(double* q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    for (int i = 0; i < 4; i++) q[i] /= n;
}
static void matvec3(double* r, const double* R, const double* v) {
    double t[3] = {R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
                   R[6] * v[0] + R[7] * v[1] + R[8] * v[2]};
    memcpy(r, t, sizeof t);
}
static void matmul3(double* r, const double* A, const double* B) {
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(r, t, sizeof t);
}
static void cross3(double* r, const double* a, const double* b) {
    double t[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    memcpy(r, t, sizeof t);
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* spatial algebra, [rot; lin], inertia as {Ixx,Iyy,Izz,Ixy,Ixz,Iyz, mcx,mcy,mcz, m} about the common origin */
static void mul_inert_vec(double* r, const double* I, const double* v) {
    double t[6];
    t[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2] + (I[7] * v[5] - I[8] * v[4]);
    t[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2] + (I[8] * v[3] - I[6] * v[5]);
    t[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2] + (I[6] * v[4] - I[7] * v[3]);
    t[3] = I[9] * v[3] - (I[7] * v[2] - I[8] * v[1]);
    t[4] = I[9] * v[4] - (I[8] * v[0] - I[6] * v[2]);
    t[5] = I[9] * v[5] - (I[6] * v[1] - I[7] * v[0]);
    memcpy(r, t, sizeof t);
}
static void cross_motion(double* r, const double* vel, const double* v) {
    double a[3], b[3], c[3];
    cross3(a, vel, v);
    cross3(b, vel, v + 3);
    cross3(c, vel + 3, v);
    r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
    r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(double* r, const double* vel, const double* f) {
    double a[3], b[3], c[3];
    cross3(a, vel, f);
    cross3(b, vel + 3, f + 3);
    cross3(c, vel, f + 3);
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
    r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

/* dense Cholesky of the leading n x n block (row stride ld), lower triangle in place; returns rank deficit */
static int chol_factor(double* A, int n, int ld) {
    int bad = 0;
    for (int j = 0; j < n; j++) {
        double s = A[j * ld + j];
        for (int k = 0; k < j; k++) s -= A[j * ld + k] * A[j * ld + k];
        if (s < MINVAL) { s = MINVAL; bad++; }
        s = sqrt(s);
        A[j * ld + j] = s;
        for (int i = j + 1; i < n; i++) {
            double t = A[i * ld + j];
            for (int k = 0; k < j; k++) t -= A[i * ld + k] * A[j * ld + k];
            A[i * ld + j] = t / s;
        }
    }
    return bad;
}
static void chol_solve(const double* L, int n, int ld, double* x) {
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= L[i * ld + k] * x[k];
        x[i] = s / L[i * ld + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k * ld + i] * x[k];
        x[i] = s / L[i * ld + i];
    }
}

/* ------------------------------------------------------------------ reset (mj_resetData) */

void qgo_reset(const qgo_model* m, qgo_data* d) {
    memset(d, 0, sizeof *d);
    memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
}

/* ------------------------------------------------------------------ position stage */

/* mj_kinematics: body frames from qpos (stage of mj_step, reference call site /root/reference/src/envs/quadruped.py:165;
   tree and joint conventions from /root/reference/src/models/quadruped/quadruped.xml:62-142) */
static void kinematics(const qgo_model* m, qgo_data* d) {
    double* R0 = d->xmat;
    memset(d->xpos, 0, 3 * sizeof(double));
    memset(R0, 0, 9 * sizeof(double));
    R0[0] = R0[4] = R0[8] = 1;
    for (int b = 1; b < m->nbody; b++) {
        int p = m->body_parent[b];
        double *xp = d->xpos + 3 * b, *xm = d->xmat + 9 * b;
        int j0 = -1, nj = 0;
        for (int j = 0; j < m->njnt; j++)
            if (m->jnt_body[j] == b) { if (j0 < 0) j0 = j; nj++; }
        if (nj && m->jnt_type[j0] == 0) {
            int a = m->jnt_qposadr[j0];
            double q[4] = {d->qpos[a + 3], d->qpos[a + 4], d->qpos[a + 5], d->qpos[a + 6]};
            normalize4(q);
            memcpy(xp, d->qpos + a, 3 * sizeof(double));
            quat2mat(xm, q);
            memcpy(d->xanchor + 3 * j0, xp, 3 * sizeof(double));
            continue;
        }
        double off[3], Rl[9];
        matvec3(off, d->xmat + 9 * p, m->body_pos + 3 * b);
        for (int k = 0; k < 3; k++) xp[k] = d->xpos[3 * p + k] + off[k];
        quat2mat(Rl, m->body_quat + 4 * b);
        matmul3(xm, d->xmat + 9 * p, Rl);
        for (int j = j0; j < j0 + nj; j++) {
            /* hinge: rotate by (qpos - qpos0) about the joint axis through the joint anchor */
            double th = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
            const double* ax = m->jnt_axis + 3 * j;
            double s = sin(th), c = cos(th), t = 1 - c;
            double Rj[9] = {t * ax[0] * ax[0] + c, t * ax[0] * ax[1] - s * ax[2], t * ax[0] * ax[2] + s * ax[1],
                            t * ax[0] * ax[1] + s * ax[2], t * ax[1] * ax[1] + c, t * ax[1] * ax[2] - s * ax[0],
                            t * ax[0] * ax[2] - s * ax[1], t * ax[1] * ax[2] + s * ax[0], t * ax[2] * ax[2] + c};
            double anchor[3], a2[3];
            matvec3(anchor, xm, m->jnt_pos + 3 * j);
            for (int k = 0; k < 3; k++) d->xanchor[3 * j + k] = xp[k] + anchor[k];
            matvec3(d->xaxis + 3 * j, xm, ax);
            matmul3(xm, xm, Rj);
            matvec3(a2, xm, m->jnt_pos + 3 * j);
            for (int k = 0; k < 3; k++) xp[k] = d->xanchor[3 * j + k] - a2[k];
        }
    }
    for (int b = 1; b < m->nbody; b++) {
        double c[3];
        matvec3(c, d->xmat + 9 * b, m->body_ipos + 3 * b);
        for (int k = 0; k < 3; k++) d->xipos[3 * b + k] = d->xpos[3 * b + k] + c[k];
    }
}

/* subtree CoM of the root, cinert, cdof  (mj_comPos) */
static void com_pos(const qgo_model* m, qgo_data* d) {
    double mt = 0, o[3] = {0, 0, 0};
    for (int b = 1; b < m->nbody; b++) {
        mt += m->body_mass[b];
        for (int k = 0; k < 3; k++) o[k] += m->body_mass[b] * d->xipos[3 * b + k];
    }
    for (int k = 0; k < 3; k++) d->com[k] = o[k] / mt;
    memset(d->cinert, 0, 10 * sizeof(double));
    for (int b = 1; b < m->nbody; b++) {
        const double *R = d->xmat + 9 * b, *I6 = m->body_inertia + 6 * b;
        double Ib[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]}, T[9], Rt[9], Iw[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
        matmul3(T, R, Ib);
        matmul3(Iw, T, Rt);
        double dd[3], mass = m->body_mass[b], *ci = d->cinert + 10 * b;
        for (int k = 0; k < 3; k++) dd[k] = d->xipos[3 * b + k] - d->com[k];
        double d2 = dot3(dd, dd);
        ci[0] = Iw[0] + mass * (d2 - dd[0] * dd[0]);
        ci[1] = Iw[4] + mass * (d2 - dd[1] * dd[1]);
        ci[2] = Iw[8] + mass * (d2 - dd[2] * dd[2]);
        ci[3] = Iw[1] - mass * dd[0] * dd[1];
        ci[4] = Iw[2] - mass * dd[0] * dd[2];
        ci[5] = Iw[5] - mass * dd[1] * dd[2];
        ci[6] = mass * dd[0]; ci[7] = mass * dd[1]; ci[8] = mass * dd[2];
        ci[9] = mass;
    }
    for (int j = 0; j < m->njnt; j++) {
        int b = m->jnt_body[j], dof = m->jnt_dofadr[j];
        double off[3];
        for (int k = 0; k < 3; k++) off[k] = d->com[k] - d->xanchor[3 * j + k];
        if (m->jnt_type[j] == 0) {
            for (int k = 0; k < 3; k++) {
                double* cd = d->cdof + 6 * (dof + k);
                memset(cd, 0, 6 * sizeof(double));
                cd[3 + k] = 1;
                double ax[3] = {d->xmat[9 * b + k], d->xmat[9 * b + 3 + k], d->xmat[9 * b + 6 + k]};
                double* cr = d->cdof + 6 * (dof + 3 + k);
                memcpy(cr, ax, sizeof ax);
                cross3(cr + 3, ax, off);
            }
        } else {
            double* cd = d->cdof + 6 * dof;
            memcpy(cd, d->xaxis + 3 * j, 3 * sizeof(double));
            cross3(cd + 3, d->xaxis + 3 * j, off);
        }
    }
}

/* composite rigid body -> dense M (mj_crb), then Cholesky (mj_factorM) */
static void crb(const qgo_model* m, qgo_data* d) {
    int nv = m->nv;
    double crbI[QGO_MAXNBODY * 10];
    memcpy(crbI, d->cinert, sizeof(double) * 10 * m->nbody);
    for (int b = m->nbody - 1; b > 0; b--) {
        int p = m->body_parent[b];
        if (p > 0)
            for (int k = 0; k < 10; k++) crbI[10 * p + k] += crbI[10 * b + k];
    }
    memset(d->M, 0, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) {
        double buf[6];
        mul_inert_vec(buf, crbI + 10 * m->dof_body[i], d->cdof + 6 * i);
        d->M[i * nv + i] = m->dof_armature[i];
        for (int j = i; j >= 0; j = m->dof_parent[j]) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += d->cdof[6 * j + k] * buf[k];
            d->M[i * nv + j] += s;
            d->M[j * nv + i] = d->M[i * nv + j];
        }
    }
    memcpy(d->L, d->M, sizeof(double) * nv * nv);
    chol_factor(d->L, nv, nv);
}

/* point Jacobian (translational, 3 x nv) of `point` moving with `body` (mj_jac) */
static void jac_point(const qgo_model* m, const qgo_data* d, double* jp, int body, const double* point) {
    int nv = m->nv;
    memset(jp, 0, sizeof(double) * 3 * nv);
    if (body <= 0) return;
    int dof = m->body_dofadr[body] + m->body_dofnum[body] - 1;
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = point[k] - d->com[k];
    for (; dof >= 0; dof = m->dof_parent[dof]) {
        const double* cd = d->cdof + 6 * dof;
        double t[3];
        cross3(t, cd, off);
        for (int k = 0; k < 3; k++) jp[k * nv + dof] = cd[3 + k] + t[k];
    }
}

/* mjc_PlaneConvex restated: floor plane (/root/reference/src/models/quadruped/scene.xml:21) vs the convex hull of every
   robot mesh geom (quadruped.xml:8,145-150): support vertex + up to 3 hull-graph neighbours */
static void collision(const qgo_model* m, qgo_data* d) {
    double pz = m->opt_f[7];
    int rule_first_only = m->opt_i[4];
    d->ncon = 0;
    d->nvert_tested = 0;
    for (int g = 0; g < m->ngeom; g++) {
        int b = m->geom_body[g], me = m->geom_mesh[g];
        double gc[3], Rg[9], Rw[9], margin = m->geom_margin[g];
        matvec3(gc, d->xmat + 9 * b, m->geom_pos + 3 * g);
        for (int k = 0; k < 3; k++) gc[k] += d->xpos[3 * b + k];
        /* bounding-sphere cull against the plane */
        if (gc[2] - pz - m->geom_rbound[g] > margin) continue;
        quat2mat(Rg, m->geom_quat + 4 * g);
        matmul3(Rw, d->xmat + 9 * b, Rg);
        const double* V = m->mesh_vert + 3 * m->mesh_vertadr[me];
        int nvert = m->mesh_vertnum[me], best = 0;
        double zrow[3] = {Rw[6], Rw[7], Rw[8]}, hbest = 1e300;
        for (int i = 0; i < nvert; i++) {
            double h = dot3(zrow, V + 3 * i);
            if (h < hbest) { hbest = h; best = i; }
        }
        d->nvert_tested += nvert;
        double dist = gc[2] + hbest - pz;
        if (dist > margin) continue;
        int first = d->ncon, cnt = 0;
        int cand = best;
        const int* edge = m->mesh_edge + m->mesh_edgeadr[me] + m->mesh_vert_edge[m->mesh_vertadr[me] + best];
        for (;;) {
            double vw[3], dv;
            matvec3(vw, Rw, V + 3 * cand);
            for (int k = 0; k < 3; k++) vw[k] += gc[k];
            dv = vw[2] - pz;
            int ok = (cnt == 0) || (dv <= margin);
            if (ok && cnt > 0) {
                double tol = 0.3 * m->geom_rbound[g];
                int kmax = rule_first_only ? 1 : cnt;
                for (int k = 0; k < kmax; k++) {
                    const double* pk = d->con_vert + 3 * (first + k);
                    double e[3] = {vw[0] - pk[0], vw[1] - pk[1], vw[2] - pk[2]};
                    if (sqrt(dot3(e, e)) < tol) ok = 0;
                }
            }
            if (ok && d->ncon < QGO_MAXCON) {
                int c = d->ncon++;
                memcpy(d->con_vert + 3 * c, vw, sizeof vw);
                d->con_pos[3 * c] = vw[0];
                d->con_pos[3 * c + 1] = vw[1];
                d->con_pos[3 * c + 2] = vw[2] - 0.5 * dv;
                d->con_dist[c] = dv;
                d->con_geom[c] = g;
                d->con_vertid[c] = cand;
                cnt++;
            }
            if (cnt >= 4 || *edge < 0) break;
            cand = *edge++;
        }
    }
}

/* impedance and reference acceleration parameters (mj_makeImpedance / getsolparam) */
static void impedance(const double* solref, const double* solimp, double timestep, double r,
                      double* K, double* B, double* imp) {
    double d0 = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
    double tc = solref[0], dr = solref[1];
    if (tc < 2 * timestep) tc = 2 * timestep; /* refsafe */
    *K = 1.0 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr);
    *B = 2.0 / fmax(MINVAL, dmax * tc);
    double x = fabs(r) / fmax(MINVAL, width), y;
    if (x >= 1) y = 1;
    else if (x <= 0) y = 0;
    else if (power == 1) y = x;
    else if (x <= mid) y = pow(x / mid, power) * mid; /* = x^p / mid^(p-1) */
    else y = 1 - pow((1 - x) / (1 - mid), power) * (1 - mid);
    double v = d0 + y * (dmax - d0);
    if (v < 0.0001) v = 0.0001;
    if (v > 0.9999) v = 0.9999;
    *imp = v;
}

/* mj_makeConstraint: joint-limit rows (ranges of quadruped.xml:25,30,35) then pyramidal or elliptic contact rows */
static void make_constraint(const qgo_model* m, qgo_data* d) {
    int nv = m->nv, ne = 0;
    double h = m->opt_f[0];
    for (int j = 0; j < m->njnt; j++) {
        if (m->jnt_type[j] != 3 || !m->jnt_limited[j]) continue;
        double q = d->qpos[m->jnt_qposadr[j]];
        for (int side = -1; side <= 1; side += 2) {
            double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - q);
            if (dist < 0 && ne < QGO_MAXEFC) {
                double* J = d->efc_J + (size_t)ne * nv;
                memset(J, 0, sizeof(double) * nv);
                J[m->jnt_dofadr[j]] = -side;
                d->efc_pos[ne] = dist;
                d->efc_margin[ne] = 0;
                d->efc_diagApprox[ne] = m->dof_invweight0[m->jnt_dofadr[j]];
                d->efc_type[ne] = 0;
                d->efc_id[ne] = j;
                ne++;
            }
        }
    }
    d->nlimit = ne;
    double jp[3 * QGO_MAXNV];
    for (int c = 0; c < d->ncon; c++) {
        int g = d->con_geom[c], b = m->geom_body[g];
        double mu = m->geom_mu[g] / sqrt(m->opt_f[6]);
        if (d->con_dist[c] >= m->geom_margin[g]) continue; /* includemargin */
        jac_point(m, d, jp, b, d->con_pos + 3 * c);
        /* contact frame rows: n = +z, t1 = +y, t2 = -x ; geom1 = plane (world, zero Jacobian) */
        const double* Jn = jp + 2 * nv;
        const double* Jt[2] = {jp + nv, jp};
        double sgn[2] = {1, -1};
        double tran = m->body_invweight0[2 * b];
        if (m->opt_i[1] == 1) {
            /* elliptic cone: rows normal, t1, t2; friction rows carry pos = margin = 0 (mj_instantiateContact) */
            for (int k = 0; k < 3 && ne < QGO_MAXEFC; k++) {
                double* J = d->efc_J + (size_t)ne * nv;
                for (int i = 0; i < nv; i++) J[i] = k == 0 ? Jn[i] : sgn[k - 1] * Jt[k - 1][i];
                d->efc_pos[ne] = k == 0 ? d->con_dist[c] : 0.0;
                d->efc_margin[ne] = k == 0 ? m->geom_margin[g] : 0.0;
                d->efc_diagApprox[ne] = tran;
                d->efc_type[ne] = 10 + k; /* 10: normal row of an elliptic contact */
                d->efc_id[ne] = c;
                ne++;
            }
            continue;
        }
        for (int k = 0; k < 2; k++)
            for (int s = 0; s < 2 && ne < QGO_MAXEFC; s++) {
                double* J = d->efc_J + (size_t)ne * nv;
                double w = (s == 0 ? mu : -mu) * sgn[k];
                for (int i = 0; i < nv; i++) J[i] = Jn[i] + w * Jt[k][i];
                d->efc_pos[ne] = d->con_dist[c];
                d->efc_margin[ne] = m->geom_margin[g];
                d->efc_diagApprox[ne] = tran + mu * mu * tran;
                d->efc_type[ne] = 1 + 2 * k + s; /* 1: first row of a pyramid */
                d->efc_id[ne] = c;
                ne++;
            }
    }
    d->nefc = ne;
    /* R, D, K/B/imp */
    for (int i = 0; i < ne; i++) {
        const double *sr, *si;
        if (d->efc_type[i] == 0) { sr = m->jnt_solref; si = m->jnt_solimp; }
        else { int g = d->con_geom[d->efc_id[i]]; sr = m->geom_solref + 2 * g; si = m->geom_solimp + 5 * g; }
        double K, B, imp;
        impedance(sr, si, h, d->efc_pos[i] - d->efc_margin[i], &K, &B, &imp);
        d->efc_K[i] = K; d->efc_B[i] = B; d->efc_imp[i] = imp;
        d->efc_R[i] = fmax(MINVAL, (1 - imp) / imp * d->efc_diagApprox[i]);
    }
    for (int i = 0; i < ne; i++)
        if (d->efc_type[i] == 1) { /* pyramidal: all rows share Rpy = 2 mu^2 R_first */
            int g = d->con_geom[d->efc_id[i]];
            double mu = m->geom_mu[g] / sqrt(m->opt_f[6]);
            double Rpy = 2 * mu * mu * d->efc_R[i];
            for (int k = 0; k < 4; k++) d->efc_R[i + k] = Rpy;
        }
    for (int i = 0; i < ne; i++)
        if (d->efc_type[i] == 10) { /* elliptic: friction R = R_normal / impratio (equal sliding coefficients) */
            double impratio = fmax(MINVAL, m->opt_f[6]);
            d->efc_R[i + 1] = d->efc_R[i + 2] = d->efc_R[i] / impratio;
        }
    for (int i = 0; i < ne; i++) d->efc_D[i] = 1.0 / d->efc_R[i];
}

/* elliptic cone of one contact (rows i, i+1, i+2): zone 0 = satisfied, 1 = fully quadratic, 2 = cone surface.
   Restates the three-zone cost of MuJoCo's primal solvers in the scaled coordinates U = (mu*jar_n, fri*jar_t). */
typedef struct { int zone; double cost, f[3], N, T, U1, U2, Dm, mu, fri; } ellzone;
static ellzone ell_eval(const qgo_model* m, const qgo_data* d, int i, double j0, double j1, double j2) {
    ellzone z;
    int g = d->con_geom[d->efc_id[i]];
    z.fri = m->geom_mu[g];
    z.mu = z.fri * sqrt(d->efc_R[i + 1] / d->efc_R[i]);
    z.N = j0 * z.mu; z.U1 = j1 * z.fri; z.U2 = j2 * z.fri;
    z.T = sqrt(z.U1 * z.U1 + z.U2 * z.U2);
    z.Dm = d->efc_D[i] / fmax(z.mu * z.mu * (1 + z.mu * z.mu), MINVAL);
    z.cost = 0; z.f[0] = z.f[1] = z.f[2] = 0;
    if (z.N >= z.mu * z.T || (z.T <= 0 && z.N >= 0)) z.zone = 0;
    else if (z.mu * z.N + z.T <= 0 || (z.T <= 0 && z.N < 0)) {
        z.zone = 1;
        z.cost = 0.5 * (d->efc_D[i] * j0 * j0 + d->efc_D[i + 1] * j1 * j1 + d->efc_D[i + 2] * j2 * j2);
        z.f[0] = -d->efc_D[i] * j0; z.f[1] = -d->efc_D[i + 1] * j1; z.f[2] = -d->efc_D[i + 2] * j2;
    } else {
        double NmT = z.N - z.mu * z.T;
        z.zone = 2;
        z.cost = 0.5 * z.Dm * NmT * NmT;
        z.f[0] = -z.Dm * NmT * z.mu;
        z.f[1] = -z.f[0] / z.T * z.U1 * z.fri;
        z.f[2] = -z.f[0] / z.T * z.U2 * z.fri;
    }
    return z;
}

/* ------------------------------------------------------------------ velocity stage */

/* mj_comVel: spatial velocities and cdof_dot (stage of mj_step, quadruped.py:165) */
static void com_vel(const qgo_model* m, qgo_data* d) {
    memset(d->cvel, 0, 6 * sizeof(double));
    for (int b = 1; b < m->nbody; b++) {
        double* cv = d->cvel + 6 * b;
        memcpy(cv, d->cvel + 6 * m->body_parent[b], 6 * sizeof(double));
        int d0 = m->body_dofadr[b], nd = m->body_dofnum[b];
        if (nd == 6) {
            for (int k = 0; k < 3; k++) {
                memset(d->cdof_dot + 6 * (d0 + k), 0, 6 * sizeof(double));
                for (int i = 0; i < 6; i++) cv[i] += d->cdof[6 * (d0 + k) + i] * d->qvel[d0 + k];
            }
            for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot + 6 * (d0 + k), cv, d->cdof + 6 * (d0 + k));
            for (int k = 3; k < 6; k++)
                for (int i = 0; i < 6; i++) cv[i] += d->cdof[6 * (d0 + k) + i] * d->qvel[d0 + k];
        } else {
            for (int k = 0; k < nd; k++) {
                cross_motion(d->cdof_dot + 6 * (d0 + k), cv, d->cdof + 6 * (d0 + k));
                for (int i = 0; i < 6; i++) cv[i] += d->cdof[6 * (d0 + k) + i] * d->qvel[d0 + k];
            }
        }
    }
}

/* mj_rne(flg_acc = 0): Coriolis, centrifugal and gravity forces (stage of mj_step, quadruped.py:165; gravity is the
   MuJoCo default since quadruped.xml:4 sets only the integrator) */
static void rne_bias(const qgo_model* m, qgo_data* d) {
    double cacc[QGO_MAXNBODY * 6], cfrc[QGO_MAXNBODY * 6];
    memset(cacc, 0, sizeof cacc);
    memset(cfrc, 0, sizeof cfrc);
    cacc[3] = -m->opt_f[1]; cacc[4] = -m->opt_f[2]; cacc[5] = -m->opt_f[3];
    for (int b = 1; b < m->nbody; b++) {
        double *ca = cacc + 6 * b, t[6], u[6];
        memcpy(ca, cacc + 6 * m->body_parent[b], 6 * sizeof(double));
        for (int k = 0; k < m->body_dofnum[b]; k++) {
            int dd = m->body_dofadr[b] + k;
            for (int i = 0; i < 6; i++) ca[i] += d->cdof_dot[6 * dd + i] * d->qvel[dd];
        }
        mul_inert_vec(t, d->cinert + 10 * b, ca);
        mul_inert_vec(u, d->cinert + 10 * b, d->cvel + 6 * b);
        cross_force(cfrc + 6 * b, d->cvel + 6 * b, u);
        for (int i = 0; i < 6; i++) cfrc[6 * b + i] += t[i];
    }
    for (int b = m->nbody - 1; b > 0; b--) {
        int p = m->body_parent[b];
        if (p > 0)
            for (int i = 0; i < 6; i++) cfrc[6 * p + i] += cfrc[6 * b + i];
    }
    for (int i = 0; i < m->nv; i++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += d->cdof[6 * i + k] * cfrc[6 * m->dof_body[i] + k];
        d->qfrc_bias[i] = s;
    }
}

/* mj_fwdActuation for the 12 <position> servos (/root/reference/src/models/quadruped/quadruped.xml:10-17,26-36,156-172):
   ctrl clamp, first-order activation filter, affine bias, force clamp, gear */
static void actuation(const qgo_model* m, qgo_data* d) {
    memset(d->qfrc_actuator, 0, sizeof(double) * m->nv);
    for (int i = 0; i < m->nu; i++) {
        int dof = m->act_dof[i];
        double ctrl = d->ctrl[i], gear = m->act_gear[i];
        if (m->act_ctrllimited[i]) ctrl = fmin(fmax(ctrl, m->act_ctrlrange[2 * i]), m->act_ctrlrange[2 * i + 1]);
        double act_in = ctrl;
        if (m->act_tau[i] > 0) {
            d->act_dot[i] = (ctrl - d->act[i]) / fmax(MINVAL, m->act_tau[i]);
            act_in = d->act[i];
        } else d->act_dot[i] = 0;
        /* joint transmission: length = gear*qpos, velocity = gear*qvel */
        int qadr = -1;
        for (int j = 0; j < m->njnt; j++)
            if (m->jnt_dofadr[j] == dof) qadr = m->jnt_qposadr[j];
        double len = gear * d->qpos[qadr], vel = gear * d->qvel[dof];
        double f = m->act_gain[i] * act_in + m->act_bias[3 * i] + m->act_bias[3 * i + 1] * len + m->act_bias[3 * i + 2] * vel;
        d->act_clamped[i] = 0;
        if (m->act_frclimited[i]) {
            if (f <= m->act_frcrange[2 * i]) { f = m->act_frcrange[2 * i]; d->act_clamped[i] = 1; }
            else if (f >= m->act_frcrange[2 * i + 1]) { f = m->act_frcrange[2 * i + 1]; d->act_clamped[i] = 1; }
        }
        d->actuator_force[i] = f;
        d->qfrc_actuator[dof] += gear * f;
    }
}

/* ------------------------------------------------------------------ constraint solver (primal Newton) */

static const qgo_model* g_model_for_cone; /* set by fwd_constraint: the cone helpers need geom_mu */

static double constraint_update(const qgo_data* d, const double* jar, double* force, int* active) {
    double cost = 0;
    for (int i = 0; i < d->nefc; i++) {
        if (d->efc_type[i] == 10) {
            ellzone z = ell_eval(g_model_for_cone, d, i, jar[i], jar[i + 1], jar[i + 2]);
            for (int k = 0; k < 3; k++) { force[i + k] = z.f[k]; active[i + k] = z.zone; }
            cost += z.cost;
            i += 2;
            continue;
        }
        if (jar[i] < 0) {
            force[i] = -d->efc_D[i] * jar[i];
            active[i] = 1;
            cost += 0.5 * d->efc_D[i] * jar[i] * jar[i];
        } else { force[i] = 0; active[i] = 0; }
    }
    return cost;
}

typedef struct { double alpha, cost, d1, d2; } lspoint;

static lspoint ls_eval(const qgo_data* d, const double* jar, const double* jv, const double* quadG, double a) {
    lspoint p = {a, quadG[0] + a * quadG[1] + a * a * quadG[2], quadG[1] + 2 * a * quadG[2], 2 * quadG[2]};
    for (int i = 0; i < d->nefc; i++) {
        if (d->efc_type[i] == 10) {
            double x0 = jar[i] + a * jv[i], x1 = jar[i + 1] + a * jv[i + 1], x2 = jar[i + 2] + a * jv[i + 2];
            ellzone z = ell_eval(g_model_for_cone, d, i, x0, x1, x2);
            if (z.zone == 1) {
                for (int k = 0; k < 3; k++) {
                    double xk = jar[i + k] + a * jv[i + k];
                    p.d1 += d->efc_D[i + k] * jv[i + k] * xk;
                    p.d2 += d->efc_D[i + k] * jv[i + k] * jv[i + k];
                }
            } else if (z.zone == 2) {
                double Np = jv[i] * z.mu, V1 = jv[i + 1] * z.fri, V2 = jv[i + 2] * z.fri;
                double Tp = (z.U1 * V1 + z.U2 * V2) / z.T, Tpp = (V1 * V1 + V2 * V2 - Tp * Tp) / z.T;
                double e = z.N - z.mu * z.T, ep = Np - z.mu * Tp;
                p.d1 += z.Dm * e * ep;
                p.d2 += z.Dm * (ep * ep - e * z.mu * Tpp);
            }
            p.cost += z.cost;
            i += 2;
            continue;
        }
        double x = jar[i] + a * jv[i];
        if (x < 0) {
            p.cost += 0.5 * d->efc_D[i] * x * x;
            p.d1 += d->efc_D[i] * jv[i] * x;
            p.d2 += d->efc_D[i] * jv[i] * jv[i];
        }
    }
    return p;
}

/* exact-to-tolerance 1-D minimisation of the convex piecewise-quadratic cost along `search`:
   safeguarded Newton on the derivative (bracket [lo, hi], bisection when Newton leaves it) */
static double linesearch(const qgo_model* m, qgo_data* d, const double* jar, const double* jv,
                         const double* quadG, double snorm, double scale) {
    double gtol = m->opt_f[4] * m->opt_f[5] * snorm / scale;
    double lo = 0, hi = -1, a = 0;
    int maxit = m->opt_i[3];
    lspoint p = ls_eval(d, jar, jv, quadG, 0);
    if (p.d1 >= 0) return 0;
    for (int it = 0; it < maxit; it++) {
        d->ls_evals++;
        double an = p.alpha - p.d1 / p.d2;
        if (an <= lo || (hi > 0 && an >= hi)) an = hi > 0 ? 0.5 * (lo + hi) : 2 * (lo > 0 ? lo : 1e-3);
        p = ls_eval(d, jar, jv, quadG, an);
        a = an;
        if (fabs(p.d1) < gtol) break;
        if (p.d1 < 0) lo = an; else hi = an;
    }
    return a;
}

/* mj_solNewton restated: primal Newton with exact line search (solver = MuJoCo default; quadruped.xml:4 does not override
   it; stage of mj_step, quadruped.py:165) */
static void solve_newton(const qgo_model* m, qgo_data* d) {
    int nv = m->nv, ne = d->nefc;
    double Ma[QGO_MAXNV], grad[QGO_MAXNV], search[QGO_MAXNV], Mv[QGO_MAXNV];
    static __thread double jar[QGO_MAXEFC], jv[QGO_MAXEFC], H[QGO_MAXNV * QGO_MAXNV];
    static __thread int active[QGO_MAXEFC];
    double scale = 1.0 / (m->opt_f[8] * (nv > 1 ? nv : 1)), tol = m->opt_f[4];
    double cost, gauss;
    d->solver_niter = 0;

#define MATVEC(out, A_, x_) for (int i_ = 0; i_ < nv; i_++) { double s_ = 0; for (int k_ = 0; k_ < nv; k_++) s_ += (A_)[i_ * nv + k_] * (x_)[k_]; (out)[i_] = s_; }
#define JVEC(out, x_) for (int i_ = 0; i_ < ne; i_++) { double s_ = 0; const double* J_ = d->efc_J + (size_t)i_ * nv; for (int k_ = 0; k_ < nv; k_++) s_ += J_[k_] * (x_)[k_]; (out)[i_] = s_; }

    MATVEC(Ma, d->M, d->qacc);
    JVEC(jar, d->qacc);
    for (int i = 0; i < ne; i++) jar[i] -= d->efc_aref[i];

    for (int iter = 0;; iter++) {
        /* constraint state, cost, gradient, Hessian, Newton direction */
        double oldcost = 0;
        if (iter > 0) oldcost = d->solver_cost;
        cost = constraint_update(d, jar, d->efc_force, active);
        gauss = 0;
        for (int i = 0; i < nv; i++) gauss += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
        cost += gauss;
        d->solver_cost = cost;
        for (int i = 0; i < nv; i++) {
            double s = 0;
            for (int r = 0; r < ne; r++) s += d->efc_J[(size_t)r * nv + i] * d->efc_force[r];
            d->qfrc_constraint[i] = s;
            grad[i] = Ma[i] - d->qfrc_smooth[i] - s;
        }
        if (iter > 0) {
            double gn = 0;
            for (int i = 0; i < nv; i++) gn += grad[i] * grad[i];
            double improvement = scale * (oldcost - cost), gradient = scale * sqrt(gn);
            d->solver_niter = iter;
            if (improvement < tol || gradient < tol || iter >= m->opt_i[2]) break;
        }
        memcpy(H, d->M, sizeof(double) * nv * nv);
        for (int r = 0; r < ne; r++) {
            if (d->efc_type[r] == 10) {
                /* 3x3 block W of the contact in jar coordinates, H += J_c^T W J_c */
                double W[9] = {0};
                if (active[r] == 1) { W[0] = d->efc_D[r]; W[4] = d->efc_D[r + 1]; W[8] = d->efc_D[r + 2]; }
                else if (active[r] == 2) {
                    ellzone z = ell_eval(m, d, r, jar[r], jar[r + 1], jar[r + 2]);
                    double S[3] = {z.mu, z.fri, z.fri}, U[3] = {z.N, z.U1, z.U2}, e = z.N - z.mu * z.T, Hs[9];
                    Hs[0] = z.Dm;
                    for (int j = 1; j < 3; j++) Hs[j] = Hs[3 * j] = -z.Dm * z.mu * U[j] / z.T;
                    for (int j = 1; j < 3; j++)
                        for (int k = 1; k < 3; k++)
                            Hs[3 * j + k] = z.Dm * z.mu * z.mu * U[j] * U[k] / (z.T * z.T) -
                                            z.Dm * e * z.mu * ((j == k ? 1.0 : 0.0) / z.T - U[j] * U[k] / (z.T * z.T * z.T));
                    for (int j = 0; j < 3; j++)
                        for (int k = 0; k < 3; k++) W[3 * j + k] = S[j] * Hs[3 * j + k] * S[k];
                }
                if (active[r]) {
                    const double* Jc = d->efc_J + (size_t)r * nv;
                    for (int a = 0; a < 3; a++)
                        for (int bb = 0; bb < 3; bb++) {
                            double w = W[3 * a + bb];
                            if (w == 0) continue;
                            for (int i = 0; i < nv; i++) {
                                double wi = w * Jc[a * nv + i];
                                if (wi == 0) continue;
                                for (int k = 0; k <= i; k++) H[i * nv + k] += wi * Jc[bb * nv + k];
                            }
                        }
                }
                r += 2;
                continue;
            }
            if (active[r]) {
                const double* J = d->efc_J + (size_t)r * nv;
                for (int i = 0; i < nv; i++) {
                    if (J[i] == 0) continue;
                    double w = d->efc_D[r] * J[i];
                    for (int k = 0; k <= i; k++) H[i * nv + k] += w * J[k];
                }
            }
        }
        chol_factor(H, nv, nv);
        for (int i = 0; i < nv; i++) search[i] = -grad[i];
        chol_solve(H, nv, nv, search);

        /* line search */
        double snorm = 0, quadG[3];
        for (int i = 0; i < nv; i++) snorm += search[i] * search[i];
        snorm = sqrt(snorm);
        if (snorm < MINVAL) { d->solver_niter = iter; break; }
        MATVEC(Mv, d->M, search);
        JVEC(jv, search);
        quadG[0] = gauss; quadG[1] = 0; quadG[2] = 0;
        for (int i = 0; i < nv; i++) {
            quadG[1] += search[i] * (Ma[i] - d->qfrc_smooth[i]);
            quadG[2] += 0.5 * search[i] * Mv[i];
        }
        double alpha = linesearch(m, d, jar, jv, quadG, snorm, scale);
        if (alpha == 0) { d->solver_niter = iter; break; }
        for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; }
        for (int i = 0; i < ne; i++) jar[i] += alpha * jv[i];
    }
}

/* mj_fwdConstraint: warm start selection, then the solver (stage of mj_step, quadruped.py:165) */
static void fwd_constraint(const qgo_model* m, qgo_data* d) {
    int nv = m->nv, ne = d->nefc;
    g_model_for_cone = m;
    d->ls_evals = 0;
    if (ne == 0) {
        memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
        memset(d->qfrc_constraint, 0, sizeof(double) * nv);
        d->solver_niter = 0;
        return;
    }
    /* warm start: pick the cheaper of qacc_warmstart and qacc_smooth */
    static __thread double jar[QGO_MAXEFC], frc[QGO_MAXEFC];
    static __thread int act[QGO_MAXEFC];
    double cost_w, cost_s, Ma[QGO_MAXNV];
    for (int i = 0; i < ne; i++) {
        double s = 0, t = 0;
        for (int k = 0; k < nv; k++) {
            s += d->efc_J[(size_t)i * nv + k] * d->qacc_warmstart[k];
            t += d->efc_J[(size_t)i * nv + k] * d->qacc_smooth[k];
        }
        jar[i] = s - d->efc_aref[i];
        frc[i] = t - d->efc_aref[i];
    }
    cost_w = constraint_update(d, jar, d->efc_force, act);
    for (int i = 0; i < nv; i++) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += d->M[i * nv + k] * d->qacc_warmstart[k];
        Ma[i] = s;
    }
    for (int i = 0; i < nv; i++) cost_w += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc_warmstart[i] - d->qacc_smooth[i]);
    cost_s = constraint_update(d, frc, d->efc_force, act);
    memcpy(d->qacc, cost_w < cost_s ? d->qacc_warmstart : d->qacc_smooth, sizeof(double) * nv);
    solve_newton(m, d);
}

/* ------------------------------------------------------------------ forward / step */

/* mj_sensorPos/Vel/Acc for the sensor block of /root/reference/src/models/quadruped/quadruped.xml:174-217, read by
   _get_obs at /root/reference/src/envs/quadruped.py:141-143 */
static void sensors(const qgo_model* m, qgo_data* d) {
    /* site FRAME sits at the free body's origin with identity orientation (quadruped.xml:69) */
    const double* R = d->xmat + 9;
    double* s = d->sensordata;
    for (int i = 0; i < 12; i++) s[i] = d->qpos[7 + i];
    double a[3] = {d->qacc[0] - m->opt_f[1], d->qacc[1] - m->opt_f[2], d->qacc[2] - m->opt_f[3]};
    for (int k = 0; k < 3; k++) {
        s[12 + k] = R[k] * a[0] + R[3 + k] * a[1] + R[6 + k] * a[2];                        /* accelerometer */
        s[15 + k] = d->qvel[3 + k];                                                           /* gyro */
        s[18 + k] = d->xpos[3 + k];                                                           /* framepos */
        s[21 + k] = d->qvel[k];                                                               /* framelinvel */
        s[24 + k] = R[3 * k];                                                                 /* framexaxis */
        s[27 + k] = R[3 * k + 2];                                                             /* framezaxis */
        s[30 + k] = R[k] * d->qvel[0] + R[3 + k] * d->qvel[1] + R[6 + k] * d->qvel[2];       /* velocimeter */
    }
}

void qgo_forward(const qgo_model* m, qgo_data* d) {
    int nv = m->nv;
    kinematics(m, d);
    com_pos(m, d);
    crb(m, d);
    collision(m, d);
    make_constraint(m, d);
    com_vel(m, d);
    for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
    /* reference acceleration (mj_referenceConstraint) */
    for (int i = 0; i < d->nefc; i++) {
        double v = 0;
        for (int k = 0; k < nv; k++) v += d->efc_J[(size_t)i * nv + k] * d->qvel[k];
        d->efc_vel[i] = v;
        d->efc_aref[i] = -d->efc_B[i] * v - d->efc_K[i] * d->efc_imp[i] * (d->efc_pos[i] - d->efc_margin[i]);
    }
    rne_bias(m, d);
    actuation(m, d);
    for (int i = 0; i < nv; i++) {
        d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
        d->qacc_smooth[i] = d->qfrc_smooth[i];
    }
    chol_solve(d->L, nv, nv, d->qacc_smooth);
    fwd_constraint(m, d);
    sensors(m, d);
}

static int bad_state(const double* x, int n) {
    for (int i = 0; i < n; i++)
        if (!(fabs(x[i]) < 1e10)) return 1;
    return 0;
}

/* mj_implicit (integrator="implicitfast", /root/reference/src/models/quadruped/quadruped.xml:4) followed by mj_advance */
static void integrate(const qgo_model* m, qgo_data* d) {
    int nv = m->nv;
    double h = m->opt_f[0], qacc[QGO_MAXNV], H[QGO_MAXNV * QGO_MAXNV];
    if (m->opt_i[0] == 1) {
        /* implicitfast: (M - h*dF/dv) qacc = qfrc_smooth + qfrc_constraint, dF/dv diagonal here */
        memcpy(H, d->M, sizeof(double) * nv * nv);
        for (int i = 0; i < nv; i++) H[i * nv + i] += h * m->dof_damping[i];
        for (int i = 0; i < m->nu; i++) {
            if (m->act_frclimited[i] && d->act_clamped[i]) continue;
            int dof = m->act_dof[i];
            H[dof * nv + dof] -= h * m->act_gear[i] * m->act_gear[i] * m->act_bias[3 * i + 2];
        }
        chol_factor(H, nv, nv);
        for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
        chol_solve(H, nv, nv, qacc);
    } else {
        /* semi-implicit Euler with implicit joint damping (mj_Euler) */
        memcpy(H, d->M, sizeof(double) * nv * nv);
        for (int i = 0; i < nv; i++) H[i * nv + i] += h * m->dof_damping[i];
        chol_factor(H, nv, nv);
        for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
        chol_solve(H, nv, nv, qacc);
    }
    /* mj_advance */
    for (int i = 0; i < m->nu; i++) {
        double tau = m->act_tau[i];
        if (tau > 0) d->act[i] += d->act_dot[i] * tau * (1 - exp(-h / tau));
    }
    for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
    for (int j = 0; j < m->njnt; j++) {
        int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
        if (m->jnt_type[j] == 0) {
            for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
            double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
            double n = sqrt(dot3(w, w)), qr[4] = {1, 0, 0, 0};
            if (n > MINVAL) {
                double ang = h * n, s = sin(0.5 * ang) / n;
                qr[0] = cos(0.5 * ang); qr[1] = w[0] * s; qr[2] = w[1] * s; qr[3] = w[2] * s;
            }
            normalize4(d->qpos + qa + 3);
            mulquat(d->qpos + qa + 3, d->qpos + qa + 3, qr);
        } else {
            d->qpos[qa] += h * d->qvel[da];
        }
    }
    d->time += h;
    memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
}

void qgo_step(const qgo_model* m, qgo_data* d) {
    if (bad_state(d->qpos, m->nq) || bad_state(d->qvel, m->nv)) {
        int w = d->warnings + 1;
        double ctrl[QGO_MAXNU];
        memcpy(ctrl, d->ctrl, sizeof ctrl);
        qgo_reset(m, d);
        memcpy(d->ctrl, ctrl, sizeof ctrl);
        d->warnings = w;
    }
    qgo_forward(m, d);
    if (bad_state(d->qacc, m->nv)) {
        int w = d->warnings + 1;
        double ctrl[QGO_MAXNU];
        memcpy(ctrl, d->ctrl, sizeof ctrl);
        qgo_reset(m, d);
        memcpy(d->ctrl, ctrl, sizeof ctrl);
        d->warnings = w;
        qgo_forward(m, d);
    }
    integrate(m, d);
}

/* QuadrupedEnv.step() physics part: ctrl[:] = clip(action,-1,1); frame_skip x mj_step (quadruped.py:160-165) */
void qgo_env_step(const qgo_model* m, qgo_data* d, const double* action, int frame_skip) {
    for (int i = 0; i < m->nu; i++) d->ctrl[i] = fmin(fmax(action[i], -1.0), 1.0);
    for (int s = 0; s < frame_skip; s++) qgo_step(m, d);
}

/* batched driver used by the CPU baseline: n independent envs, each stepped `n_steps` env-steps with
   actions[step][env][nu]; optional auto-reset on time >= max_time or zaxis_z < 0 */
void qgo_rollout(const qgo_model* m, qgo_data* d, int n_envs, const double* actions, int n_steps,
                 int frame_skip, double max_time, int auto_reset, double* obs_out) {
    for (int s = 0; s < n_steps; s++)
        for (int e = 0; e < n_envs; e++) {
            qgo_data* de = d + e;
            qgo_env_step(m, de, actions + ((size_t)s * n_envs + e) * m->nu, frame_skip);
            if (obs_out) memcpy(obs_out + ((size_t)s * n_envs + e) * QGO_NSENSORDATA, de->sensordata, sizeof(double) * QGO_NSENSORDATA);
            if (auto_reset && (de->time >= max_time || de->sensordata[29] < 0)) {
                qgo_reset(m, de);
                for (int i = 0; i < m->nu; i++) de->ctrl[i] = (i % 3 == 2) ? -0.5 : 0.0;
            }
        }
}

size_t qgo_sizeof_data(void) { return sizeof(qgo_data); }
size_t qgo_sizeof_model(void) { return sizeof(qgo_model); }
