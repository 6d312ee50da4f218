// qg_api.cu -- C ABI of libquadgym.so (declared in include/quadgym.h): model blob -> device tables,
// batch state planes in HBM, launches on the caller's stream.  No torch types, no CPU fallback.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/quadgym.h"
#include "qg_walk.cuh"

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return fail(QG_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_));     \
    } while (0)

struct qg_model {
    QgModelC c;
    std::vector<float> verts;     // xyz_ per hull vertex
    std::vector<int> vert_adj;    // per vertex: start of its neighbour list in int4 units, relative to the mesh' edge0 (rides in verts[].w)
    std::vector<int> adj;         // neighbour lists (local vertex ids), -1 terminated and padded to groups of 4
    std::vector<int> vert_cadj, cadj;  // same for the polytope-edge graph (hill climbing)
    int sizes[8];
    double timestep;
};

// One launch domain of the step kernel: the whole batch (device path) or one segment of the pipelined host path.
// Each has its own chunk counter, binning histogram and slice of the slot -> env permutation.
struct Segment {
    int env0 = 0, cnt = 0;
    int* d_chunk = nullptr;       // [2] chunk counter of the persistent step kernel, blocks-done counter
    int* d_bin_count = nullptr;   // [2][QG_NBINS] = counts, cursors
    bool perm_valid = false;
    cudaStream_t st = nullptr;    // host path only
    cudaEvent_t done = nullptr;   // the segment's outputs are in the host buffers
    cudaEvent_t binned = nullptr; // ... and its next-step permutation is built (everything on `st` is finished)
};

struct qg_batch {
    int n, device;
    QgModelC* d_model;
    float4* d_verts;
    int4 *d_adj4, *d_cadj4;
    float4* d_state;
    QgCounters* d_ctr;
    QgStepOpts opts;
    size_t smem;
    int num_sms, cone;
    // environment binning (slot -> env permutation refreshed after every step launch)
    int* d_perm;                 // [N] slot -> env; the full-batch launch and the host segments lay it out differently
    unsigned char* d_bin_key;    // [N] key the step kernel leaves for the next launch's binning
    bool binning;
    Segment full;                // the device path's launch domain
    std::vector<Segment> segs;   // the host path's (qg_step_host), created on first use
    int perm_layout;             // 0 = d_perm holds one permutation of the batch, 1 = one permutation per host segment
    cudaEvent_t ev_start;
    // device staging for the host-buffer path (qg_step_host)
    float *d_act, *d_obs, *d_rew, *d_terms, *d_tobs;
    unsigned char* d_term;
    // WalkingQuadrupedEnv reward stack (qg_walk_*)
    bool walk_on;
    QgWalkState walk;
    QgWalkOpts wopts;
    std::vector<void*> walk_allocs, po_allocs;
    bool po_on;
    QgPoState po;
};

extern "C" const char* qg_last_error(void) { return g_err; }
extern "C" const char* qg_version(void) { return "quadgym-b200 0.1 (sm_100a)"; }
extern "C" unsigned long long qg_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------------ blob
struct Blob {
    std::map<std::string, std::vector<double>> f;
    std::map<std::string, std::vector<int>> i;
};

static int parse_blob(const void* blob, size_t n, Blob& out) {
    const unsigned char* b = (const unsigned char*)blob;
    if (!blob || n < 16 || memcmp(b, "QGBLOB01", 8) != 0) return fail(QG_EBLOB, "bad blob magic");
    uint32_t nsec;
    memcpy(&nsec, b + 8, 4);
    size_t off = 16;
    for (uint32_t s = 0; s < nsec; ++s) {
        if (off + 24 > n) return fail(QG_EBLOB, "truncated blob (section header %u)", s);
        char name[17];
        memcpy(name, b + off, 16);
        name[16] = 0;
        uint32_t code, cnt;
        memcpy(&code, b + off + 16, 4);
        memcpy(&cnt, b + off + 20, 4);
        off += 24;
        size_t nbytes = (size_t)cnt * (code == 1 ? 8 : 4);
        if (code != 1 && code != 2) return fail(QG_EBLOB, "section %s: bad dtype %u", name, code);
        if (off + nbytes > n) return fail(QG_EBLOB, "truncated blob (section %s)", name);
        if (code == 1) {
            std::vector<double> v(cnt);
            memcpy(v.data(), b + off, nbytes);
            out.f[name] = std::move(v);
        } else {
            std::vector<int> v(cnt);
            memcpy(v.data(), b + off, nbytes);
            out.i[name] = std::move(v);
        }
        off += nbytes + ((8 - nbytes % 8) % 8);
    }
    return QG_OK;
}

static void quat2mat(const double* q, float* R) {
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    R[0] = (float)(1 - 2 * (y * y + z * z)); R[1] = (float)(2 * (x * y - w * z)); R[2] = (float)(2 * (x * z + w * y));
    R[3] = (float)(2 * (x * y + w * z)); R[4] = (float)(1 - 2 * (x * x + z * z)); R[5] = (float)(2 * (y * z - w * x));
    R[6] = (float)(2 * (x * z - w * y)); R[7] = (float)(2 * (y * z + w * x)); R[8] = (float)(1 - 2 * (x * x + y * y));
}

extern "C" int qg_model_load(const void* blob, size_t nbytes, qg_model** out) {
    if (!out) return fail(QG_EINVAL, "out is NULL");
    Blob B;
    int rc = parse_blob(blob, nbytes, B);
    if (rc) return rc;
    static const char* need_f[] = {"opt_f", "body_pos", "body_quat", "body_mass", "body_ipos", "body_inertia",
                                   "body_invweight0", "jnt_axis", "jnt_pos", "jnt_range", "jnt_solref", "jnt_solimp",
                                   "qpos0", "dof_damping", "dof_armature", "dof_invweight0", "act_gear", "act_gain",
                                   "act_bias", "act_tau", "act_ctrlrange", "act_frcrange", "geom_pos", "geom_quat",
                                   "geom_rbound", "geom_margin", "geom_mu", "geom_solref", "geom_solimp", "mesh_vert"};
    static const char* need_i[] = {"sizes", "opt_i", "body_parent", "jnt_type", "jnt_body", "jnt_qposadr", "jnt_dofadr",
                                   "jnt_limited", "dof_body", "act_dof", "act_ctrllimited", "act_frclimited",
                                   "geom_body", "geom_mesh", "mesh_vertadr", "mesh_vertnum", "mesh_edgeadr",
                                   "mesh_vert_edge", "mesh_edge"};
    for (const char* k : need_f)
        if (!B.f.count(k)) return fail(QG_EBLOB, "missing section %s", k);
    for (const char* k : need_i)
        if (!B.i.count(k)) return fail(QG_EBLOB, "missing section %s", k);
    const std::vector<int>& sz = B.i["sizes"];
    if (sz.size() != 8) return fail(QG_EBLOB, "bad sizes section");
    const int nq = sz[0], nv = sz[1], nu = sz[2], nbody = sz[3], njnt = sz[4], ngeom = sz[5], nmesh = sz[6];
    if (nq != QG_NQ || nv != QG_NV || nbody != 2 + QG_NLEG * QG_NLINK || njnt != 1 + QG_NLEG * QG_NLINK || sz[7] != QG_NSENSORDATA)
        return fail(QG_EMODEL, "unsupported sizes nq=%d nv=%d nbody=%d njnt=%d (need free base + 4 legs x 3 hinges)", nq, nv, nbody, njnt);
    auto F = [&](const char* k) -> const std::vector<double>& { return B.f[k]; };
    auto I = [&](const char* k) -> const std::vector<int>& { return B.i[k]; };
#define CHECK_LEN(vec, len, name) \
    if ((int)(vec).size() != (int)(len)) return fail(QG_EBLOB, "section %s: expected %d values, got %d", name, (int)(len), (int)(vec).size());
    CHECK_LEN(F("opt_f"), 9, "opt_f");
    CHECK_LEN(I("opt_i"), 5, "opt_i");
    CHECK_LEN(I("body_parent"), nbody, "body_parent");
    CHECK_LEN(F("body_pos"), 3 * nbody, "body_pos");
    CHECK_LEN(F("body_quat"), 4 * nbody, "body_quat");
    CHECK_LEN(F("body_mass"), nbody, "body_mass");
    CHECK_LEN(F("body_ipos"), 3 * nbody, "body_ipos");
    CHECK_LEN(F("body_inertia"), 6 * nbody, "body_inertia");
    CHECK_LEN(F("body_invweight0"), 2 * nbody, "body_invweight0");
    CHECK_LEN(I("jnt_type"), njnt, "jnt_type");
    CHECK_LEN(I("jnt_body"), njnt, "jnt_body");
    CHECK_LEN(I("jnt_dofadr"), njnt, "jnt_dofadr");
    CHECK_LEN(I("jnt_qposadr"), njnt, "jnt_qposadr");
    CHECK_LEN(F("jnt_axis"), 3 * njnt, "jnt_axis");
    CHECK_LEN(F("jnt_pos"), 3 * njnt, "jnt_pos");
    CHECK_LEN(F("jnt_range"), 2 * njnt, "jnt_range");
    CHECK_LEN(I("jnt_limited"), njnt, "jnt_limited");
    CHECK_LEN(F("qpos0"), nq, "qpos0");
    CHECK_LEN(F("dof_damping"), nv, "dof_damping");
    CHECK_LEN(F("dof_armature"), nv, "dof_armature");
    CHECK_LEN(F("dof_invweight0"), nv, "dof_invweight0");
    CHECK_LEN(I("act_dof"), nu, "act_dof");
    CHECK_LEN(F("act_gear"), nu, "act_gear");
    CHECK_LEN(F("act_gain"), nu, "act_gain");
    CHECK_LEN(F("act_bias"), 3 * nu, "act_bias");
    CHECK_LEN(F("act_tau"), nu, "act_tau");
    CHECK_LEN(F("act_ctrlrange"), 2 * nu, "act_ctrlrange");
    CHECK_LEN(F("act_frcrange"), 2 * nu, "act_frcrange");
    CHECK_LEN(I("act_ctrllimited"), nu, "act_ctrllimited");
    CHECK_LEN(I("act_frclimited"), nu, "act_frclimited");
    CHECK_LEN(I("geom_body"), ngeom, "geom_body");
    CHECK_LEN(F("geom_pos"), 3 * ngeom, "geom_pos");
    CHECK_LEN(F("geom_quat"), 4 * ngeom, "geom_quat");
    CHECK_LEN(I("geom_mesh"), ngeom, "geom_mesh");
    CHECK_LEN(F("geom_rbound"), ngeom, "geom_rbound");
    CHECK_LEN(F("geom_margin"), ngeom, "geom_margin");
    CHECK_LEN(F("geom_mu"), ngeom, "geom_mu");
    CHECK_LEN(F("geom_solref"), 2 * ngeom, "geom_solref");
    CHECK_LEN(F("geom_solimp"), 5 * ngeom, "geom_solimp");
    CHECK_LEN(I("mesh_vertadr"), nmesh, "mesh_vertadr");
    CHECK_LEN(I("mesh_vertnum"), nmesh, "mesh_vertnum");
    CHECK_LEN(I("mesh_edgeadr"), nmesh, "mesh_edgeadr");
    if (nu > QG_NU) return fail(QG_EMODEL, "too many actuators (%d)", nu);

    qg_model* m = new qg_model();
    QgModelC& c = m->c;
    memset(&c, 0, sizeof c);
    memcpy(m->sizes, sz.data(), sizeof m->sizes);
    const std::vector<double>& of = F("opt_f");
    const std::vector<int>& oi = I("opt_i");
    m->timestep = of[0];
    c.timestep = (float)of[0];
    c.timestep_d = of[0];
    c.grav[0] = (float)of[1]; c.grav[1] = (float)of[2]; c.grav[2] = (float)of[3];
    c.plane_z = (float)of[7];
    c.tol = (float)std::fmax(of[4], 1e-6);  // fp32 floor on the solver tolerance ("precision")
    c.scale = (float)(1.0 / (of[8] * (nv > 1 ? nv : 1)));
    c.integrator = oi[0];
    if (oi[1] != 0 && oi[1] != 1) { delete m; return fail(QG_EMODEL, "unknown friction cone type %d", oi[1]); }
    if (oi[1] == 0 && of[6] != 1.0) { delete m; return fail(QG_EMODEL, "impratio != 1 is supported for elliptic cones only"); }
    if (!(of[6] > 0)) { delete m; return fail(QG_EMODEL, "impratio must be positive"); }
    c.cone = oi[1];
    c.impratio = (float)of[6];
    c.mu_scale = (float)(1.0 / std::sqrt(of[6]));
    c.max_iter = oi[2] < 20 ? oi[2] : 20;
    c.ls_iter = oi[3] < 12 ? oi[3] : 12;
    c.rule_first = oi[4];

    // ---- topology: body 1 = free base, then 4 chains of 3 single-hinge bodies
    const std::vector<int>& bp = I("body_parent");
    const std::vector<int>& jt = I("jnt_type");
    const std::vector<int>& jb = I("jnt_body");
    const std::vector<int>& jd = I("jnt_dofadr");
    const std::vector<int>& jq = I("jnt_qposadr");
#define BADMODEL(...) { delete m; return fail(QG_EMODEL, __VA_ARGS__); }
    if (bp[1] != 0 || jb[0] != 1 || jt[0] != 0 || jd[0] != 0 || jq[0] != 0) BADMODEL("body 1 must be the free-floating base");
    int body_leg[64], body_level[64];
    for (int b = 0; b < nbody; ++b) body_leg[b] = -1, body_level[b] = (b == 1 ? 0 : -1);
    int nleg = 0;
    std::vector<int> body_joint(nbody, -1);
    for (int j = 0; j < njnt; ++j) {
        if (body_joint[jb[j]] >= 0) BADMODEL("body %d has more than one joint", jb[j]);
        body_joint[jb[j]] = j;
    }
    for (int b = 2; b < nbody; ++b) {
        int p = bp[b];
        if (p == 1) {
            if (nleg >= QG_NLEG) BADMODEL("more than %d legs", QG_NLEG);
            body_leg[b] = nleg++;
            body_level[b] = 1;
        } else {
            if (p < 2 || body_leg[p] < 0 || body_level[p] >= QG_NLINK) BADMODEL("body %d is not part of a 3-link leg chain", b);
            body_leg[b] = body_leg[p];
            body_level[b] = body_level[p] + 1;
            for (int b2 = 2; b2 < b; ++b2)
                if (b2 != b && bp[b2] == p) BADMODEL("leg links must form a chain (body %d has two children)", p);
        }
        int j = body_joint[b];
        if (j < 0 || jt[j] != 3) BADMODEL("body %d needs exactly one hinge joint", b);
        int l = body_leg[b], k = body_level[b] - 1;
        if (jd[j] != 6 + 3 * l + k || jq[j] != 7 + 3 * l + k) BADMODEL("joint %d: dof order must be base, then legs in order", j);
        const double* ax = &F("jnt_axis")[3 * j];
        const double* jp = &F("jnt_pos")[3 * j];
        if (std::fabs(ax[0]) > 1e-9 || std::fabs(ax[1]) > 1e-9 || std::fabs(ax[2] - 1) > 1e-9) BADMODEL("joint %d: hinge axis must be the local z axis", j);
        if (std::fabs(jp[0]) + std::fabs(jp[1]) + std::fabs(jp[2]) > 1e-12) BADMODEL("joint %d: hinge must pass through the body origin", j);
        QgJointC& J = c.joint[l][k];
        for (int i = 0; i < 3; ++i) J.pos[i] = (float)F("body_pos")[3 * b + i];
        quat2mat(&F("body_quat")[4 * b], J.Roff);
        for (int i = 0; i < 3; ++i) J.com[i] = (float)F("body_ipos")[3 * b + i];
        J.mass = (float)F("body_mass")[b];
        for (int i = 0; i < 6; ++i) J.I[i] = (float)F("body_inertia")[6 * b + i];
        J.q0 = (float)F("qpos0")[jq[j]];
        J.lo = (float)F("jnt_range")[2 * j];
        J.hi = (float)F("jnt_range")[2 * j + 1];
        J.limited = I("jnt_limited")[j];
        J.damping = (float)F("dof_damping")[jd[j]];
        J.armature = (float)F("dof_armature")[jd[j]];
        J.invw_dof = (float)F("dof_invweight0")[jd[j]];
    }
    if (nleg != QG_NLEG) BADMODEL("expected %d legs, found %d", QG_NLEG, nleg);
    for (int l = 0; l < QG_NLEG; ++l)
        for (int k = 0; k < QG_NLINK; ++k)
            if (c.joint[l][k].mass <= 0.f) BADMODEL("leg %d link %d has no mass", l, k);
    // base
    c.base_mass = (float)F("body_mass")[1];
    for (int i = 0; i < 3; ++i) c.base_com[i] = (float)F("body_ipos")[3 + i];
    for (int i = 0; i < 6; ++i) c.base_I[i] = (float)F("body_inertia")[6 + i];
    for (int i = 0; i < 6; ++i) { c.base_damp[i] = (float)F("dof_damping")[i]; c.base_arm[i] = (float)F("dof_armature")[i]; }
    if (c.base_damp[0] != c.base_damp[1] || c.base_damp[0] != c.base_damp[2] || c.base_arm[0] != c.base_arm[1] || c.base_arm[0] != c.base_arm[2])
        BADMODEL("damping/armature of the base translation must be isotropic");
    for (int i = 0; i < 19; ++i) c.qpos0[i] = (float)F("qpos0")[i];
    const double h = of[0];
    // joint-limit solver parameters
    {
        const std::vector<double>& sr = F("jnt_solref");
        const std::vector<double>& si = F("jnt_solimp");
        double tc = std::fmax(sr[0], 2 * h), dr = sr[1], dmax = si[1];
        c.lim_K = (float)(1.0 / std::fmax(1e-15, dmax * dmax * tc * tc * dr * dr));
        c.lim_B = (float)(2.0 / std::fmax(1e-15, dmax * tc));
        c.lim_d0 = (float)si[0]; c.lim_dmax = (float)si[1]; c.lim_width = (float)si[2]; c.lim_mid = (float)si[3]; c.lim_power = (float)si[4];
    }
    // actuators
    for (int a = 0; a < nu; ++a) {
        int dof = I("act_dof")[a];
        if (dof < 6 || dof >= 18) BADMODEL("actuator %d must act on a leg hinge", a);
        QgJointC& J = c.joint[(dof - 6) / 3][(dof - 6) % 3];
        if (J.has_act) BADMODEL("more than one actuator on dof %d", dof);
        J.has_act = 1;
        J.gear = (float)F("act_gear")[a];
        J.kp = (float)F("act_gain")[a];
        J.b0 = (float)F("act_bias")[3 * a]; J.b1 = (float)F("act_bias")[3 * a + 1]; J.b2 = (float)F("act_bias")[3 * a + 2];
        double tau = F("act_tau")[a];
        J.has_dyn = tau > 0;
        J.inv_tau = tau > 0 ? (float)(1.0 / tau) : 0.f;
        J.act_fac = tau > 0 ? (float)(tau * (1.0 - std::exp(-h / tau))) : 0.f;
        J.ctrl_limited = I("act_ctrllimited")[a];
        J.frc_limited = I("act_frclimited")[a];
        J.ctrl_lo = (float)F("act_ctrlrange")[2 * a]; J.ctrl_hi = (float)F("act_ctrlrange")[2 * a + 1];
        J.frc_lo = (float)F("act_frcrange")[2 * a]; J.frc_hi = (float)F("act_frcrange")[2 * a + 1];
    }
    // vertex tables
    const std::vector<double>& mv = F("mesh_vert");
    int nvert = (int)mv.size() / 3;
    if (nvert > QG_MAXVERT) BADMODEL("too many hull vertices (%d > %d)", nvert, QG_MAXVERT);
    CHECK_LEN(I("mesh_vert_edge"), nvert, "mesh_vert_edge");
    m->verts.resize(4 * (size_t)nvert);
    for (int i = 0; i < nvert; ++i) {
        m->verts[4 * i] = (float)mv[3 * i]; m->verts[4 * i + 1] = (float)mv[3 * i + 1]; m->verts[4 * i + 2] = (float)mv[3 * i + 2];
        m->verts[4 * i + 3] = 0.f;
    }
    // neighbour lists re-packed per mesh: each list is -1 terminated and padded with -1 to a multiple of 4 ints.
    // Two graphs: the full hull triangulation (extra plane-mesh contacts enumerate its neighbour order) and the
    // polytope-edge graph without diagonals of coplanar facets (support-search hill climbing).
    std::vector<int> mesh_adj0(nmesh, 0), mesh_cadj0(nmesh, 0);  // start of each mesh' lists in int4 units
    auto pack_adj = [&](const std::vector<int>& ve, const std::vector<int>& ed, const std::vector<int>& eadr,
                        std::vector<int>& vert_out, std::vector<int>& adj_out, std::vector<int>& mesh0) -> bool {
        vert_out.assign(nvert, 0);
        for (int me = 0; me < nmesh; ++me) {
            int v0 = I("mesh_vertadr")[me], vn = I("mesh_vertnum")[me], e0 = eadr[me];
            mesh0[me] = (int)adj_out.size() / 4;
            for (int v = 0; v < vn; ++v) {
                vert_out[v0 + v] = (int)adj_out.size() / 4 - mesh0[me];
                size_t i = (size_t)e0 + ve[v0 + v];
                while (i < ed.size() && ed[i] >= 0) {
                    if (ed[i] >= vn) return false;
                    adj_out.push_back(ed[i++]);
                }
                adj_out.push_back(-1);
                while (adj_out.size() % 4) adj_out.push_back(-1);
            }
        }
        return true;
    };
    if (!pack_adj(I("mesh_vert_edge"), I("mesh_edge"), I("mesh_edgeadr"), m->vert_adj, m->adj, mesh_adj0))
        BADMODEL("mesh_edge: neighbour index out of range");
    if (B.i.count("mesh_cedge") && B.i.count("mesh_vert_cedge") && B.i.count("mesh_cedgeadr")) {
        CHECK_LEN(I("mesh_vert_cedge"), nvert, "mesh_vert_cedge");
        CHECK_LEN(I("mesh_cedgeadr"), nmesh, "mesh_cedgeadr");
        if (!pack_adj(I("mesh_vert_cedge"), I("mesh_cedge"), I("mesh_cedgeadr"), m->vert_cadj, m->cadj, mesh_cadj0))
            BADMODEL("mesh_cedge: neighbour index out of range");
    } else {  // older blobs: climb on the full triangulation
        m->vert_cadj = m->vert_adj;
        m->cadj = m->adj;
        mesh_cadj0 = mesh_adj0;
    }
    // one group of terminators past the last list of each table: the scans prefetch the next int4 group
    for (int i = 0; i < 4; ++i) { m->adj.push_back(-1); m->cadj.push_back(-1); }
    // the two list offsets of a vertex ride in the .w of its float4 (climb list in the low, hull list in the high half):
    // the support search reads them with the vertex instead of through one more dependent load per hop
    for (int i = 0; i < nvert; ++i) {
        if (m->vert_cadj[i] > 0xffff || m->vert_adj[i] > 0xffff) BADMODEL("hull graph too large for 16-bit list offsets");
        unsigned w = (unsigned)m->vert_cadj[i] | ((unsigned)m->vert_adj[i] << 16);
        memcpy(&m->verts[4 * (size_t)i + 3], &w, 4);
    }
    c.nvert = nvert;
    if (nmesh > QG_MAXMESH) BADMODEL("too many meshes (%d > %d)", nmesh, QG_MAXMESH);
    // support-search start table: exhaustive argmin at the centre direction of every cube-map cell
    for (int me = 0; me < nmesh; ++me) {
        int v0 = I("mesh_vertadr")[me], vn = I("mesh_vertnum")[me];
        for (int face = 0; face < 6; ++face)
            for (int iu = 0; iu < QG_DIRRES; ++iu)
                for (int iv = 0; iv < QG_DIRRES; ++iv) {
                    int a = face >> 1, o0 = (a + 1) % 3, o1 = (a + 2) % 3;
                    double d[3];
                    d[a] = (face & 1) ? 1.0 : -1.0;
                    d[o0] = (iu + 0.5) * 2.0 / QG_DIRRES - 1.0;
                    d[o1] = (iv + 0.5) * 2.0 / QG_DIRRES - 1.0;
                    int best = 0;
                    double hb = 1e300;
                    for (int i = 0; i < vn; ++i) {
                        double hh = d[0] * mv[3 * (v0 + i)] + d[1] * mv[3 * (v0 + i) + 1] + d[2] * mv[3 * (v0 + i) + 2];
                        if (hh < hb) { hb = hh; best = i; }
                    }
                    c.dir_start[me][(face * QG_DIRRES + iu) * QG_DIRRES + iv] = (unsigned short)best;
                }
    }
    // geoms: leg geoms to their lane, base geoms to the lane with the fewest geoms so far
    std::vector<int> lane_of(ngeom), level_of(ngeom);
    int count[QG_NLEG] = {0, 0, 0, 0};
    for (int g = 0; g < ngeom; ++g) {
        int b = I("geom_body")[g];
        if (b < 1 || b >= nbody) BADMODEL("geom %d: bad body", g);
        level_of[g] = body_level[b];
        lane_of[g] = body_leg[b];
        if (b != 1) count[lane_of[g]]++;
    }
    for (int g = 0; g < ngeom; ++g)
        if (I("geom_body")[g] == 1) {
            int best = 0;
            for (int l = 1; l < QG_NLEG; ++l)
                if (count[l] < count[best]) best = l;
            lane_of[g] = best;
            count[best]++;
        }
    for (int l = 0; l < QG_NLEG; ++l) {
        if (count[l] > QG_MAXGEOM_LANE) BADMODEL("too many geoms on lane %d", l);
        int n = 0;
        for (int lev = 0; lev <= QG_NLINK; ++lev) {
            c.glev[l][lev] = n;
            for (int g = 0; g < ngeom; ++g) {
                if (lane_of[g] != l || level_of[g] != lev) continue;
                QgGeomC& G = c.geom[l][n++];
                int b = I("geom_body")[g], me = I("geom_mesh")[g];
                for (int i = 0; i < 3; ++i) G.pos[i] = (float)F("geom_pos")[3 * g + i];
                quat2mat(&F("geom_quat")[4 * g], G.R);
                int v0 = I("mesh_vertadr")[me], vn = I("mesh_vertnum")[me];
                double hx = 0, hy = 0, hz = 0;
                for (int i = v0; i < v0 + vn; ++i) {
                    hx = std::fmax(hx, std::fabs(mv[3 * i])); hy = std::fmax(hy, std::fabs(mv[3 * i + 1])); hz = std::fmax(hz, std::fabs(mv[3 * i + 2]));
                }
                G.half[0] = (float)(hx * 1.0001 + 1e-7); G.half[1] = (float)(hy * 1.0001 + 1e-7); G.half[2] = (float)(hz * 1.0001 + 1e-7);
                G.margin = (float)F("geom_margin")[g];
                double mu = F("geom_mu")[g];
                G.mu = (float)mu;
                const double* sr = &F("geom_solref")[2 * g];
                const double* si = &F("geom_solimp")[5 * g];
                double tc = std::fmax(sr[0], 2 * h), dr = sr[1], dmax = si[1];
                G.K = (float)(1.0 / std::fmax(1e-15, dmax * dmax * tc * tc * dr * dr));
                G.B = (float)(2.0 / std::fmax(1e-15, dmax * tc));
                G.d0 = (float)si[0]; G.dmax = (float)si[1]; G.width = (float)si[2]; G.mid = (float)si[3]; G.power = (float)si[4];
                double tran = F("body_invweight0")[2 * b];
                // regulariser per unit (1-imp)/imp: pyramidal rows share 2 mu^2 (1+mu^2) tran, the elliptic normal row has tran
                G.Rfac = (float)(oi[1] == 1 ? tran : 2.0 * mu * mu * (1.0 + mu * mu) * tran);
                double rb = F("geom_rbound")[g];
                G.tol2 = (float)(0.09 * rb * rb);
                G.vert0 = v0; G.nvert = vn;
                G.edge0 = mesh_adj0[me];
                G.cedge0 = mesh_cadj0[me];
                G.level = lev;
                G.mesh = me;
                // bounding sphere of this hull about the LINK origin (feeds link_reach)
                double rr = 0;
                for (int i = v0; i < v0 + vn; ++i) {
                    double x[3];
                    for (int a = 0; a < 3; ++a)
                        x[a] = G.pos[a] + G.R[3 * a] * mv[3 * i] + G.R[3 * a + 1] * mv[3 * i + 1] + G.R[3 * a + 2] * mv[3 * i + 2];
                    rr = std::fmax(rr, std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]));
                }
                c.link_reach[l][lev] = std::fmax(c.link_reach[l][lev], (float)(rr * 1.0001 + 1e-6 + G.margin));
            }
            if (c.glev[l][lev] == n) c.link_reach[l][lev] = -1e30f;   // no geom at this level: always skipped
        }
        c.glev[l][QG_NLINK + 1] = n;
        c.ngeom[l] = n;
    }
#undef BADMODEL
#undef CHECK_LEN
    *out = m;
    return QG_OK;
}

extern "C" void qg_model_destroy(qg_model* m) { delete m; }

extern "C" int qg_model_info(const qg_model* m, int* sizes, double* timestep) {
    if (!m) return fail(QG_EINVAL, "model is NULL");
    if (sizes) memcpy(sizes, m->sizes, sizeof m->sizes);
    if (timestep) *timestep = m->timestep;
    return QG_OK;
}

// ------------------------------------------------------------------------------------------ batch
static void default_opts(const qg_model* m, QgStepOpts& o) {
    memset(&o, 0, sizeof o);
    o.max_time = 10.0;
    o.flip_termination = 0;
    o.auto_reset = 0;
    o.max_iter = m->c.max_iter;
    o.ls_iter = m->c.ls_iter;
    o.random_yaw = 0;
    o.seed = 0;
    o.env_offset = 0;
    for (int i = 0; i < 12; ++i) o.reset_ctrl[i] = (i % 3 == 2) ? -0.5f : 0.f;  // quadruped.py:124
    o.n_terms = 0;
}

static int batch_create_impl(const qg_model* m, int n_envs, int device, qg_batch* b);
static int segment_alloc(Segment& sg, bool with_stream) {
    CUDA_OK(cudaMalloc(&sg.d_bin_count, sizeof(int) * 2 * QG_NBINS));
    CUDA_OK(cudaMalloc(&sg.d_chunk, 2 * sizeof(int)));
    CUDA_OK(cudaMemset(sg.d_chunk, 0, 2 * sizeof(int)));
    sg.perm_valid = false;
    if (with_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&sg.st, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&sg.done, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&sg.binned, cudaEventDisableTiming));
    }
    return QG_OK;
}
static void segment_free(Segment& sg) {
    cudaFree(sg.d_bin_count); cudaFree(sg.d_chunk);
    if (sg.st) cudaStreamDestroy(sg.st);
    if (sg.done) cudaEventDestroy(sg.done);
    if (sg.binned) cudaEventDestroy(sg.binned);
    sg = Segment();
}
static int reset_impl(qg_batch* b, const uint8_t* mask_dev, uint64_t seed, int random_yaw, long long env_offset,
                      int clear_env_state, cudaStream_t st);

extern "C" int qg_batch_create(const qg_model* m, int n_envs, int device, qg_batch** out) {
    if (!m || !out || n_envs <= 0) return fail(QG_EINVAL, "bad arguments to qg_batch_create");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(QG_ECUDA, "no CUDA device available (%s); libquadgym has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (device < 0 || device >= ndev) return fail(QG_EINVAL, "device %d out of range (%d devices)", device, ndev);
    CUDA_OK(cudaSetDevice(device));
    qg_batch* b = new qg_batch();
    int rc = batch_create_impl(m, n_envs, device, b);
    if (rc) { qg_batch_destroy(b); return rc; }   // frees whatever was allocated before the failure
    *out = b;
    return QG_OK;
}

static int batch_create_impl(const qg_model* m, int n_envs, int device, qg_batch* b) {
    b->d_model = nullptr; b->d_verts = nullptr; b->d_adj4 = b->d_cadj4 = nullptr; b->d_state = nullptr; b->d_ctr = nullptr;
    b->d_perm = nullptr; b->d_bin_key = nullptr; b->perm_layout = 0; b->ev_start = nullptr;
    b->d_terms = b->d_tobs = nullptr;
    b->walk_on = false;
    b->po_on = false;
    memset(&b->po, 0, sizeof b->po);
    memset(&b->walk, 0, sizeof b->walk);
    memset(&b->wopts, 0, sizeof b->wopts);
    b->d_act = b->d_obs = b->d_rew = nullptr;
    b->d_term = nullptr;
    b->n = n_envs;
    b->device = device;
    default_opts(m, b->opts);
    {
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        b->num_sms = prop.multiProcessorCount;
    }
    size_t nv = m->verts.size() / 4;
    CUDA_OK(cudaMalloc(&b->d_model, sizeof(QgModelC)));
    CUDA_OK(cudaMalloc(&b->d_verts, sizeof(float4) * (nv ? nv : 1)));
    CUDA_OK(cudaMalloc(&b->d_adj4, sizeof(int) * (m->adj.size() ? m->adj.size() : 4)));
    CUDA_OK(cudaMalloc(&b->d_cadj4, sizeof(int) * (m->cadj.size() ? m->cadj.size() : 4)));
    CUDA_OK(cudaMalloc(&b->d_state, sizeof(float4) * (size_t)QG_NPLANE * n_envs));
    CUDA_OK(cudaMalloc(&b->d_ctr, sizeof(QgCounters)));
    CUDA_OK(cudaMemcpy(b->d_model, &m->c, sizeof(QgModelC), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(b->d_verts, m->verts.data(), sizeof(float) * m->verts.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(b->d_adj4, m->adj.data(), sizeof(int) * m->adj.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(b->d_cadj4, m->cadj.data(), sizeof(int) * m->cadj.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemset(b->d_state, 0, sizeof(float4) * (size_t)QG_NPLANE * n_envs));
    CUDA_OK(cudaMemset(b->d_ctr, 0, sizeof(QgCounters)));
    CUDA_OK(cudaMalloc(&b->d_perm, sizeof(int) * n_envs));
    CUDA_OK(cudaMalloc(&b->d_bin_key, n_envs));
    CUDA_OK(cudaMemset(b->d_bin_key, 0, n_envs));
    b->full.env0 = 0;
    b->full.cnt = n_envs;
    { int rc0 = segment_alloc(b->full, false); if (rc0) return rc0; }
    // environment binning (slot -> env permutation refreshed every step from the last physics step's per-leg contact count
    // and line-search evaluations): measured -2 % step time at 65,536 envs, for 2 extra tiny launches per step.  On by default
    // only for batches large enough to pay for those; QG_BINNING=0/1 overrides.  Results per environment do not depend
    // on the slot (test_env_binning_does_not_change_results).
    b->binning = n_envs >= 32768;
    if (const char* ev = getenv("QG_BINNING")) b->binning = atoi(ev) != 0;   // tests / experiments
    b->smem = ((sizeof(QgModelC) + 15) & ~size_t(15)) + sizeof(float4) * nv + sizeof(float) * (QG_QR_SLOTS * 32) * (QG_BLOCK / 32) +
              sizeof(float) * (SB_NWORDS * (QG_BLOCK / 4) + SL_NWORDS * QG_BLOCK);   // + the block's resident state
    CUDA_OK(cudaFuncSetAttribute(qg_step_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
    CUDA_OK(cudaFuncSetAttribute(qg_step_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
    CUDA_OK(cudaFuncSetAttribute(qg_step_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
    CUDA_OK(cudaFuncSetAttribute(qg_step_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
    if (const char* ev = getenv("QG_CARVEOUT")) {   // tuning experiments: shared-memory carve-out in percent
        int pct = atoi(ev);
        cudaFuncSetAttribute(qg_step_kernel<false, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(qg_step_kernel<false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    b->cone = m->c.cone;
    int rc = reset_impl(b, nullptr, 0, 0, 0, 1, nullptr);
    if (rc) return rc;
    CUDA_OK(cudaDeviceSynchronize());
    return QG_OK;
}

extern "C" void qg_batch_destroy(qg_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    cudaFree(b->d_model); cudaFree(b->d_verts); cudaFree(b->d_adj4); cudaFree(b->d_cadj4);
    cudaFree(b->d_state); cudaFree(b->d_ctr); cudaFree(b->d_perm); cudaFree(b->d_bin_key);
    segment_free(b->full);
    for (Segment& sg : b->segs) segment_free(sg);
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    cudaFree(b->d_act); cudaFree(b->d_obs); cudaFree(b->d_rew); cudaFree(b->d_term); cudaFree(b->d_terms); cudaFree(b->d_tobs);
    for (void* p : b->walk_allocs) cudaFree(p);
    for (void* p : b->po_allocs) cudaFree(p);
    delete b;
}

extern "C" int qg_batch_num_envs(const qg_batch* b) { return b ? b->n : 0; }

extern "C" int qg_set_options(qg_batch* b, double max_time, int flip_termination, int auto_reset,
                              int solver_iterations, int ls_iterations) {
    if (!b) return fail(QG_EINVAL, "batch is NULL");
    b->opts.max_time = max_time;
    b->opts.flip_termination = flip_termination;
    b->opts.auto_reset = auto_reset;
    if (solver_iterations > 0) b->opts.max_iter = solver_iterations;
    if (ls_iterations > 0) b->opts.ls_iter = ls_iterations;
    return QG_OK;
}

extern "C" int qg_set_reward_table(qg_batch* b, int n_terms, const int* term_ids, const double* weights,
                                   const double* params) {
    if (!b) return fail(QG_EINVAL, "batch is NULL");
    if (n_terms < 0 || n_terms > QG_MAX_TERMS) return fail(QG_EINVAL, "n_terms %d out of range [0, %d]", n_terms, QG_MAX_TERMS);
    for (int i = 0; i < n_terms; ++i) {
        if (term_ids[i] < 0 || term_ids[i] >= QG_NUM_TERMS) return fail(QG_EINVAL, "unknown reward term id %d", term_ids[i]);
        b->opts.term_id[i] = term_ids[i];
        b->opts.term_w[i] = weights ? weights[i] : 1.0;
        b->opts.term_p[i] = params ? params[i] : 0.0;
    }
    b->opts.n_terms = n_terms;
    return QG_OK;
}

static inline int nblocks(int n_envs) { return (4 * n_envs + 255) / 256; }   // helper kernels: 256 threads, 4 per env

// step-kernel block size: QG_BLOCK (warps of a block share instruction-cache fills through the block barriers).  Small
// batches halve it while the halved grid still fits ONE block per SM (more SMs busy, every SM still runs a single
// lockstep block); two desynchronised half-size blocks on one SM are slower than one full block (measured, C4: 0.716
// against 0.633 ms), so a batch that fills more than half of the SMs keeps QG_BLOCK.
static int step_block(const qg_batch* b, int n_envs) {
    int blk = QG_BLOCK;
    static const int forced = getenv("QG_STEP_BLOCK") ? atoi(getenv("QG_STEP_BLOCK")) : 0;   // tuning experiments
    if (forced >= 32 && forced <= QG_BLOCK && forced % 32 == 0) return forced;
    while (blk > 64 && 2 * ((4 * n_envs + blk - 1) / blk) <= b->num_sms) blk >>= 1;
    return blk;
}

// clear_env_state: also zero the episode counter and the reward memory (first control cost) -- batch creation only;
// a user-level reset() keeps both, like the reference (walking_quad.py:106-115 never resets previous_ctrl_cost)
static int reset_impl(qg_batch* b, const uint8_t* mask_dev, uint64_t seed, int random_yaw, long long env_offset,
                      int clear_env_state, cudaStream_t st) {
    b->opts.seed = seed;
    b->opts.random_yaw = random_yaw;
    b->opts.env_offset = env_offset;
    qg_reset_kernel<<<nblocks(b->n), 256, 0, st>>>(b->d_model, b->d_state, b->n, mask_dev, b->opts, clear_env_state);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}

extern "C" int qg_reset(qg_batch* b, const uint8_t* mask_dev, uint64_t seed, int random_yaw, long long env_offset,
                        void* stream) {
    if (!b) return fail(QG_EINVAL, "batch is NULL");
    CUDA_OK(cudaSetDevice(b->device));
    return reset_impl(b, mask_dev, seed, random_yaw, env_offset, 0, (cudaStream_t)stream);
}

// One launch of the step kernel over the environments of `sg` (all I/O pointers are whole-batch arrays indexed by the
// absolute env id).
template <bool DEBUG>
static int launch_step(qg_batch* b, Segment& sg, const float* action, int clip, int frame_skip, float* obs, float* reward,
                       float* terms, unsigned char* terminated, float* terminal_obs, QgDebugOut dbg, cudaStream_t st) {
    const int blk = step_block(b, sg.cnt);
    auto kern = b->cone ? qg_step_kernel<DEBUG, 1> : qg_step_kernel<DEBUG, 0>;
    const int nchunks = (4 * sg.cnt + blk - 1) / blk;
    const int resident = b->num_sms * QG_MINBLOCKS * (QG_BLOCK / blk);   // register file: QG_BLOCK * QG_MINBLOCKS threads per SM
    kern<<<nchunks < resident ? nchunks : resident, blk, b->smem, st>>>(b->d_model, b->d_verts, b->d_adj4, b->d_cadj4,
                                                                     b->d_state, b->n, action, clip, frame_skip, obs, reward,
                                                                     terms, terminated, terminal_obs, b->opts, b->d_ctr, dbg,
                                                                     sg.perm_valid ? b->d_perm + sg.env0 : nullptr,
                                                                     b->binning ? b->d_bin_key : nullptr, sg.d_chunk, nchunks,
                                                                     sg.env0, sg.cnt);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}

// the two binning kernels that prepare the segment's next launch (slot -> env map: environments grouped by the solver
// effort of their last physics step)
static int launch_binning(qg_batch* b, Segment& sg, cudaStream_t st) {
    if (b->binning) {
        CUDA_OK(cudaMemsetAsync(sg.d_bin_count, 0, sizeof(int) * 2 * QG_NBINS, st));
        qg_bin_hist_kernel<<<64, 256, 0, st>>>(b->d_bin_key + sg.env0, sg.cnt, sg.d_bin_count);
        qg_bin_scatter_kernel<<<(sg.cnt + 255) / 256, 256, 0, st>>>(b->d_bin_key + sg.env0, sg.cnt, sg.d_bin_count,
                                                                   sg.d_bin_count + QG_NBINS, b->d_perm + sg.env0, sg.env0);
        g_launches += 2;
        CUDA_OK(cudaGetLastError());
        sg.perm_valid = true;
    }
    return QG_OK;
}

// d_perm is laid out either as one permutation of the batch (device path) or as one per host segment: switching paths
// drops the stale permutations (the next launch runs in env order and rebuilds them)
static void use_perm_layout(qg_batch* b, int layout) {
    if (b->perm_layout == layout) return;
    b->perm_layout = layout;
    b->full.perm_valid = false;
    for (Segment& sg : b->segs) sg.perm_valid = false;
}

extern "C" int qg_step(qg_batch* b, const float* action_dev, int frame_skip, float* obs_dev, float* reward_dev,
                       float* terms_dev, uint8_t* terminated_dev, float* terminal_obs_dev, void* stream) {
    if (!b || !action_dev || !obs_dev || !reward_dev || !terminated_dev) return fail(QG_EINVAL, "qg_step: NULL argument");
    if (frame_skip < 1) return fail(QG_EINVAL, "frame_skip must be >= 1");
    CUDA_OK(cudaSetDevice(b->device));
    QgDebugOut dbg;
    memset(&dbg, 0, sizeof dbg);
    use_perm_layout(b, 0);
    int rc = launch_step<false>(b, b->full, action_dev, 1, frame_skip, obs_dev, reward_dev, terms_dev, terminated_dev,
                                terminal_obs_dev, dbg, (cudaStream_t)stream);
    return rc ? rc : launch_binning(b, b->full, (cudaStream_t)stream);
}

// Host-buffer path, pipelined: the batch is cut into segments of contiguous environments, each with its own stream:
//   H2D(actions of segment s) -> step kernel over segment s -> D2H(outputs of segment s).
// The copies of one segment run under the kernels of the others (the persistent kernel of segment s+1 takes over the
// SMs one by one as the blocks of segment s run out of chunks, so the cut costs no wave efficiency); only the first
// segment's H2D and the last segment's D2H are exposed.  Results do not depend on the segmentation (environments are
// independent, the binning permutation stays inside a segment).
static int host_segments(qg_batch* b) {
    if (!b->segs.empty()) return QG_OK;
    int nseg = b->n >= 32768 ? 4 : (b->n >= 8192 ? 2 : 1);
    if (const char* ev = getenv("QG_HOST_SEGMENTS")) nseg = atoi(ev) > 0 ? atoi(ev) : nseg;   // tuning experiments
    if (nseg > b->n) nseg = b->n;
    // Only the first segment's H2D and the last segment's D2H are exposed, so with 4 or more segments the two outer
    // ones get half the share of an inner one.  Boundaries on multiples of 64 environments (whole chunks, 16-byte
    // aligned flag slices).
    const int shares = nseg >= 4 ? 2 * nseg - 2 : nseg;
    int e0 = 0;
    for (int s = 0; s < nseg && e0 < b->n; ++s) {
        const int w = (nseg >= 4 && s > 0 && s < nseg - 1) ? 2 : 1;
        int c = (int)(((long long)b->n * w / shares + 63) / 64 * 64);
        if (s == nseg - 1 || e0 + c > b->n) c = b->n - e0;
        Segment sg;
        sg.env0 = e0;
        sg.cnt = c;
        int rc = segment_alloc(sg, true);
        b->segs.push_back(sg);   // pushed even on failure so that qg_batch_destroy frees what was allocated
        if (rc) return rc;
        e0 += c;
    }
    CUDA_OK(cudaEventCreateWithFlags(&b->ev_start, cudaEventDisableTiming));
    return QG_OK;
}

extern "C" int qg_step_host_async(qg_batch* b, const float* action_host, int frame_skip, float* obs_host, float* reward_host,
                                  float* terms_host, uint8_t* terminated_host, float* terminal_obs_host, void* stream) {
    if (!b || !action_host || !obs_host || !reward_host || !terminated_host) return fail(QG_EINVAL, "qg_step_host: NULL argument");
    if (frame_skip < 1) return fail(QG_EINVAL, "frame_skip must be >= 1");
    CUDA_OK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = b->n;
    if (!b->d_act) {
        CUDA_OK(cudaMalloc(&b->d_act, n * 12 * sizeof(float)));
        CUDA_OK(cudaMalloc(&b->d_obs, n * 33 * sizeof(float)));
        CUDA_OK(cudaMalloc(&b->d_rew, n * sizeof(float)));
        CUDA_OK(cudaMalloc(&b->d_term, n));
    }
    const int nt = b->opts.n_terms;
    if (terms_host && nt > 0 && !b->d_terms) CUDA_OK(cudaMalloc(&b->d_terms, n * QG_MAX_TERMS * sizeof(float)));
    if (terminal_obs_host && !b->d_tobs) CUDA_OK(cudaMalloc(&b->d_tobs, n * 33 * sizeof(float)));
    int rc = host_segments(b);
    if (rc) return rc;
    use_perm_layout(b, 1);
    QgDebugOut dbg;
    memset(&dbg, 0, sizeof dbg);
    float* d_terms = (terms_host && nt > 0) ? b->d_terms : nullptr;
    float* d_tobs = terminal_obs_host ? b->d_tobs : nullptr;
    CUDA_OK(cudaEventRecord(b->ev_start, st));   // the segments start after the caller's earlier work on `stream`
    // the copies run straight from / into the caller's buffers: page-locked buffers (cudaHostAlloc, torch pin_memory)
    // are DMA-ed asynchronously, pageable ones go through the driver's staging
    for (Segment& sg : b->segs) {
        const size_t e0 = sg.env0, c = sg.cnt;
        CUDA_OK(cudaStreamWaitEvent(sg.st, b->ev_start, 0));
        CUDA_OK(cudaMemcpyAsync(b->d_act + e0 * 12, action_host + e0 * 12, c * 12 * sizeof(float), cudaMemcpyHostToDevice, sg.st));
        rc = launch_step<false>(b, sg, b->d_act, 1, frame_skip, b->d_obs, b->d_rew, d_terms, b->d_term, d_tobs, dbg, sg.st);
        if (rc) return rc;
    }
    // D2H in segment order on each segment's stream (issued after ALL launches: a copy queued before a later segment's
    // H2D would hold the copy engine's queue while its kernel runs)
    for (Segment& sg : b->segs) {
        const size_t e0 = sg.env0, c = sg.cnt;
        CUDA_OK(cudaMemcpyAsync(obs_host + e0 * 33, b->d_obs + e0 * 33, c * 33 * sizeof(float), cudaMemcpyDeviceToHost, sg.st));
        CUDA_OK(cudaMemcpyAsync(reward_host + e0, b->d_rew + e0, c * sizeof(float), cudaMemcpyDeviceToHost, sg.st));
        CUDA_OK(cudaMemcpyAsync(terminated_host + e0, b->d_term + e0, c, cudaMemcpyDeviceToHost, sg.st));
        if (d_terms) CUDA_OK(cudaMemcpyAsync(terms_host + e0 * nt, d_terms + e0 * nt, c * nt * sizeof(float), cudaMemcpyDeviceToHost, sg.st));
        if (d_tobs) CUDA_OK(cudaMemcpyAsync(terminal_obs_host + e0 * 33, d_tobs + e0 * 33, c * 33 * sizeof(float), cudaMemcpyDeviceToHost, sg.st));
        CUDA_OK(cudaEventRecord(sg.done, sg.st));
    }
    // binning for the next step AFTER the copies in stream order: the tiny kernels cannot start while the next segment's
    // persistent blocks hold every SM's registers, and the D2H must not wait for them.  Later work on the caller's stream
    // (a qg_step, a state read) is ordered after everything the segments did; qg_host_wait only waits for the copies.
    for (Segment& sg : b->segs) {
        rc = launch_binning(b, sg, sg.st);
        if (rc) return rc;
        CUDA_OK(cudaEventRecord(sg.binned, sg.st));
        CUDA_OK(cudaStreamWaitEvent(st, sg.binned, 0));
    }
    return QG_OK;
}

extern "C" int qg_host_wait(qg_batch* b, void* stream) {
    if (!b) return fail(QG_EINVAL, "batch is NULL");
    CUDA_OK(cudaSetDevice(b->device));
    (void)stream;   // outputs are complete when every segment's copies are; the caller's stream needs no host-side wait
    for (Segment& sg : b->segs) CUDA_OK(cudaEventSynchronize(sg.done));
    return QG_OK;
}

extern "C" int qg_step_host(qg_batch* b, const float* action_host, int frame_skip, float* obs_host, float* reward_host,
                            float* terms_host, uint8_t* terminated_host, float* terminal_obs_host, void* stream) {
    int rc = qg_step_host_async(b, action_host, frame_skip, obs_host, reward_host, terms_host, terminated_host,
                                terminal_obs_host, stream);
    if (rc) return rc;
    return qg_host_wait(b, stream);
}

extern "C" int qg_debug_step(qg_batch* b, const float* ctrl_dev, float* qacc_dev, float* qacc_smooth_dev,
                             float* qfrc_bias_dev, float* M_dev, int* counts_dev, float* sensordata_dev, void* stream) {
    if (!b || !ctrl_dev) return fail(QG_EINVAL, "qg_debug_step: NULL argument");
    CUDA_OK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = b->n;
    // scratch outputs the step kernel always writes
    float* scratch = nullptr;
    CUDA_OK(cudaMalloc(&scratch, n * (33 + 1) * sizeof(float) + n));
    QgDebugOut dbg;
    dbg.qacc = qacc_dev; dbg.qacc_smooth = qacc_smooth_dev; dbg.qfrc_bias = qfrc_bias_dev; dbg.M = M_dev;
    dbg.counts = counts_dev; dbg.sensordata = sensordata_dev;
    QgStepOpts saved = b->opts;
    b->opts.auto_reset = 0;
    b->opts.n_terms = 0;
    use_perm_layout(b, 0);
    int rc = launch_step<true>(b, b->full, ctrl_dev, 0, 1, scratch, scratch + n * 33, nullptr, (unsigned char*)(scratch + n * 34),
                               nullptr, dbg, st);
    if (!rc) rc = launch_binning(b, b->full, st);
    b->opts = saved;
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(QG_ECUDA, "qg_debug_step: %s", cudaGetErrorString(e));
    return QG_OK;
}

extern "C" int qg_get_state(qg_batch* b, float* qpos_dev, float* qvel_dev, float* act_dev, float* warm_dev,
                            double* time_dev, float* ctrl_dev, void* stream) {
    if (!b) return fail(QG_EINVAL, "batch is NULL");
    CUDA_OK(cudaSetDevice(b->device));
    qg_get_state_kernel<<<nblocks(b->n), 256, 0, (cudaStream_t)stream>>>(b->d_state, b->n, qpos_dev, qvel_dev, act_dev,
                                                                               warm_dev, time_dev, ctrl_dev);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}

extern "C" int qg_set_state(qg_batch* b, const float* qpos_dev, const float* qvel_dev, const float* act_dev,
                            const float* warm_dev, const double* time_dev, const float* ctrl_dev, void* stream) {
    if (!b) return fail(QG_EINVAL, "batch is NULL");
    CUDA_OK(cudaSetDevice(b->device));
    qg_set_state_kernel<<<nblocks(b->n), 256, 0, (cudaStream_t)stream>>>(b->d_state, b->n, qpos_dev, qvel_dev, act_dev,
                                                                               warm_dev, time_dev, ctrl_dev);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}

extern "C" int qg_get_counters(qg_batch* b, qg_counters* out_host, int reset, void* stream) {
    if (!b || !out_host) return fail(QG_EINVAL, "qg_get_counters: NULL argument");
    CUDA_OK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    static_assert(sizeof(qg_counters) == sizeof(QgCounters), "counter layouts differ");
    CUDA_OK(cudaMemcpyAsync(out_host, b->d_ctr, sizeof(QgCounters), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (reset) CUDA_OK(cudaMemsetAsync(b->d_ctr, 0, sizeof(QgCounters), st));
    return QG_OK;
}

extern "C" int qg_fp32_peak(int device, int iters, double* tflops_out) {
    if (!tflops_out || iters < 1) return fail(QG_EINVAL, "qg_fp32_peak: bad argument");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    float* out = nullptr;
    CUDA_OK(cudaMalloc(&out, sizeof(float) * threads * blocks));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    qg_ffma_kernel<<<blocks, threads>>>(out, iters / 4 + 1, 0.999f, 0.001f);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_OK(cudaEventRecord(e0));
        qg_ffma_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        CUDA_OK(cudaEventRecord(e1));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8 * 16 * (double)iters * threads * blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        g_launches++;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_out = best;
    return QG_OK;
}

// ------------------------------------------------------------------------------------------ walking env
template <typename T>
static int walk_alloc(std::vector<void*>& owner, T** p, size_t count) {
    CUDA_OK(cudaMalloc(p, sizeof(T) * count));
    CUDA_OK(cudaMemset(*p, 0, sizeof(T) * count));
    owner.push_back(*p);
    return QG_OK;
}

extern "C" int qg_walk_set_sample_options(qg_batch* b, const double* sample_opts, const int* sample_has) {
    if (!b) return fail(QG_EINVAL, "batch is NULL");
    QgWalkOpts& o = b->wopts;
    o.min_speed = sample_opts ? sample_opts[0] : 0.0;
    o.max_speed = sample_opts ? sample_opts[1] : 1.0;
    o.fixed_heading = sample_opts ? sample_opts[2] : 0.0;
    o.fixed_velocity_angle = sample_opts ? sample_opts[3] : 0.0;
    o.fixed_speed = sample_opts ? sample_opts[4] : 0.0;
    o.has_heading = (sample_opts && sample_has) ? sample_has[0] : 0;
    o.has_velocity_angle = (sample_opts && sample_has) ? sample_has[1] : 0;
    o.has_speed = (sample_opts && sample_has) ? sample_has[2] : 0;
    return QG_OK;
}

extern "C" int qg_walk_enable(qg_batch* b, int window, double dt, double timestep, int frame_skip, double settling_time,
                              int random_controls, const double* sample_opts, const int* sample_has) {
    if (!b || window < 1 || frame_skip < 1) return fail(QG_EINVAL, "qg_walk_enable: bad argument");
    CUDA_OK(cudaSetDevice(b->device));
    for (void* p : b->walk_allocs) cudaFree(p);
    b->walk_allocs.clear();
    QgWalkState& W = b->walk;
    const size_t n = b->n;
    W.n = b->n; W.window = window; W.dt = dt; W.timestep = timestep; W.frame_skip = frame_skip; W.ema_alpha = 0.80;
    int rc = 0;
    rc |= walk_alloc(b->walk_allocs, &W.velocity, 3 * n); rc |= walk_alloc(b->walk_allocs, &W.heading, 3 * n); rc |= walk_alloc(b->walk_allocs, &W.global_velocity, 3 * n);
    rc |= walk_alloc(b->walk_allocs, &W.ideal_position, 3 * n); rc |= walk_alloc(b->walk_allocs, &W.prev_derive, n); rc |= walk_alloc(b->walk_allocs, &W.first_ctrl_cost, n);
    rc |= walk_alloc(b->walk_allocs, &W.prev_ctrl, 12 * n); rc |= walk_alloc(b->walk_allocs, &W.signal_ring, (size_t)window * 12 * n);
    rc |= walk_alloc(b->walk_allocs, &W.cross_ring, (size_t)window * 12 * n); rc |= walk_alloc(b->walk_allocs, &W.cross_count, 12 * n);
    rc |= walk_alloc(b->walk_allocs, &W.prev_sample, 12 * n); rc |= walk_alloc(b->walk_allocs, &W.f_est, 12 * n); rc |= walk_alloc(b->walk_allocs, &W.a_est, 12 * n);
    rc |= walk_alloc(b->walk_allocs, &W.prev_sign, 12 * n); rc |= walk_alloc(b->walk_allocs, &W.buffer_index, n); rc |= walk_alloc(b->walk_allocs, &W.sample_count, n);
    rc |= walk_alloc(b->walk_allocs, &W.flags, n); rc |= walk_alloc(b->walk_allocs, &W.episode, n);
    W.blk = (int)std::ceil(std::sqrt((double)window));      // block extrema of the signal ring: 2 sqrt(window) loads per update
    W.nblk = (window + W.blk - 1) / W.blk;
    rc |= walk_alloc(b->walk_allocs, &W.blk_max, (size_t)W.nblk * 12 * n); rc |= walk_alloc(b->walk_allocs, &W.blk_min, (size_t)W.nblk * 12 * n);
    if (rc) return QG_ECUDA;
    QgWalkOpts& o = b->wopts;
    o.random_controls = random_controls;
    o.auto_reset = 1;
    o.seed = b->opts.seed;
    o.env_offset = b->opts.env_offset;
    qg_walk_set_sample_options(b, sample_opts, sample_has);
    for (int i = 0; i < 12; ++i) o.joint_centers[i] = b->opts.reset_ctrl[i];   // walking_quad.py:36-39 == quadruped.py:124
    b->opts.settling_time = settling_time;
    b->walk_on = true;
    // hard reset with the batch's current seed / env_offset (qg_walk_reset re-keys the sampler on every call); the launch
    // goes to the NULL stream, so wait for it: callers continue on their own (possibly non-blocking) streams
    int rc2 = qg_walk_reset(b, nullptr, 1, b->opts.seed, b->opts.env_offset, nullptr);
    if (rc2) return rc2;
    CUDA_OK(cudaDeviceSynchronize());
    return QG_OK;
}

extern "C" int qg_walk_reset(qg_batch* b, const uint8_t* mask_dev, int hard, uint64_t seed, long long env_offset, void* stream) {
    if (!b || !b->walk_on) return fail(QG_EINVAL, "qg_walk_reset: walking mode is not enabled");
    CUDA_OK(cudaSetDevice(b->device));
    // the command sampler is keyed on (seed, env_offset + env, episode): always taken from the call, so that shards
    // and seeds set after qg_walk_enable reach it (`hard` only adds: clear the estimator / first control cost / episode)
    b->wopts.seed = seed;
    b->wopts.env_offset = env_offset;
    qg_walk_reset_kernel<<<(b->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(b->walk, b->wopts, mask_dev, hard);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}

extern "C" int qg_walk_set_commands(qg_batch* b, const double* speed_alpha_theta_dev, const uint8_t* mask_dev, void* stream) {
    if (!b || !b->walk_on || !speed_alpha_theta_dev) return fail(QG_EINVAL, "qg_walk_set_commands: bad argument");
    CUDA_OK(cudaSetDevice(b->device));
    qg_walk_set_commands_kernel<<<(b->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(b->walk, speed_alpha_theta_dev, mask_dev);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}

extern "C" int qg_walk_get_commands(qg_batch* b, double* velocity_dev, double* heading_dev, double* global_velocity_dev,
                                    double* ideal_position_dev, double* f_est_dev, double* a_est_dev, void* stream) {
    if (!b || !b->walk_on) return fail(QG_EINVAL, "qg_walk_get_commands: walking mode is not enabled");
    CUDA_OK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = b->n;
    if (velocity_dev) CUDA_OK(cudaMemcpyAsync(velocity_dev, b->walk.velocity, 3 * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (heading_dev) CUDA_OK(cudaMemcpyAsync(heading_dev, b->walk.heading, 3 * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (global_velocity_dev) CUDA_OK(cudaMemcpyAsync(global_velocity_dev, b->walk.global_velocity, 3 * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (ideal_position_dev) CUDA_OK(cudaMemcpyAsync(ideal_position_dev, b->walk.ideal_position, 3 * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (f_est_dev) CUDA_OK(cudaMemcpyAsync(f_est_dev, b->walk.f_est, 12 * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (a_est_dev) CUDA_OK(cudaMemcpyAsync(a_est_dev, b->walk.a_est, 12 * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return QG_OK;
}

extern "C" int qg_walk_step(qg_batch* b, float* obs_dev, const float* ctrl_dev, const uint8_t* terminated_dev,
                            float* terminal_obs_dev, float* reward_dev, float* terms_dev, double* reward64_dev,
                            double* terms64_dev, int auto_reset, void* stream) {
    if (!b || !b->walk_on || !obs_dev) return fail(QG_EINVAL, "qg_walk_step: bad argument");
    CUDA_OK(cudaSetDevice(b->device));
    b->wopts.auto_reset = auto_reset;
    qg_walk_kernel<<<(b->n + QG_WALK_ENVS_PER_BLOCK - 1) / QG_WALK_ENVS_PER_BLOCK, 12 * QG_WALK_ENVS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
        b->walk, b->wopts, obs_dev, ctrl_dev, b->d_state, terminated_dev, terminal_obs_dev, reward_dev, terms_dev, reward64_dev, terms64_dev);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}

// ------------------------------------------------------------------------------------------ partially observable env
extern "C" int qg_po_enable(qg_batch* b, int obs_window, double Dt, double beta, double settling_time) {
    if (!b || !b->walk_on || obs_window < 1) return fail(QG_EINVAL, "qg_po_enable: enable the walking stack first (qg_walk_enable)");
    CUDA_OK(cudaSetDevice(b->device));
    QgPoState& P = b->po;
    P.n = b->n; P.window = obs_window; P.Dt = Dt; P.beta = beta; P.settle_half = settling_time / 2;
    for (void* p : b->po_allocs) cudaFree(p);
    b->po_allocs.clear();
    if (walk_alloc(b->po_allocs, &P.q, 4 * (size_t)b->n) || walk_alloc(b->po_allocs, &P.is_view, (size_t)b->n) ||
        walk_alloc(b->po_allocs, &P.ring, (size_t)b->n * obs_window * QG_PO_FRAME) || walk_alloc(b->po_allocs, &P.head_ctr, 2)) return QG_ECUDA;
    std::vector<double> q0(4 * (size_t)b->n, 0.0);
    for (int i = 0; i < b->n; ++i) q0[4 * (size_t)i] = 1.0;      // computed_orientation = [1, 0, 0, 0] (po_walking_quad.py:19)
    CUDA_OK(cudaMemcpy(P.q, q0.data(), q0.size() * sizeof(double), cudaMemcpyHostToDevice));
    b->po_on = true;
    return QG_OK;
}

extern "C" int qg_po_observe(qg_batch* b, const float* sensordata_dev, const uint8_t* terminated_dev, float* stacked_dev,
                             float* terminal_stacked_dev, int auto_reset, int is_reset_call, void* stream) {
    if (!b || !b->po_on || !stacked_dev || (!is_reset_call && !sensordata_dev)) return fail(QG_EINVAL, "qg_po_observe: bad argument");
    CUDA_OK(cudaSetDevice(b->device));
    static const int po_block = getenv("QG_PO_BLOCK") ? atoi(getenv("QG_PO_BLOCK")) : 256;   // tuning experiments
    qg_po_kernel<<<(b->n + QG_PO_ENVS_PER_BLOCK - 1) / QG_PO_ENVS_PER_BLOCK, po_block, 0, (cudaStream_t)stream>>>(
        b->po, b->walk, b->wopts, sensordata_dev ? sensordata_dev : stacked_dev, b->d_state, terminated_dev, stacked_dev,
        terminal_stacked_dev, auto_reset, is_reset_call);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return QG_OK;
}
