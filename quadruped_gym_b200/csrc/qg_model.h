// qg_model.h -- device-side constant tables of the compiled model (host fills, kernels read).
//
// The kernels are specialised on the TOPOLOGY of the reference robot
// (/root/reference/src/models/quadruped/quadruped.xml:62-142): one free-floating base and
// QG_NLEG = 4 legs of QG_NLINK = 3 single-hinge bodies, hinge axis = local z through the body
// origin.  Every NUMBER (masses, inertias, offsets, gains, ranges, hull vertices ...) comes from
// the model blob, so edited MJCFs of the same topology work unchanged; qg_model_load rejects
// anything else with QG_EMODEL.
#pragma once
#include <stdint.h>

#define QG_NLEG 4
#define QG_NLINK 3
#define QG_MAXGEOM_LANE 8   // geoms handled by one lane (leg): 5 on the leg + its share of the base
#define QG_MAXCON_LANE 24   // contacts one lane can hold
#define QG_MAXVERT 1024     // unique hull vertices over all meshes (float4 each in shared memory)
#define QG_MAX_TERMS_ 16
#define QG_MAXMESH 8
#ifndef QG_DIRRES
#define QG_DIRRES 8          // support-search start table: cube map, 6 faces x QG_DIRRES^2 cells per mesh
#endif
#define QG_DIRCELLS (6 * QG_DIRRES * QG_DIRRES)

// state planes: float4 S[plane * N + env]
#define QG_PL_POS 0     // base position (world)            x y z _
#define QG_PL_QUAT 1    // base quaternion                   w x y z
#define QG_PL_VLIN 2    // base linear velocity (world)      x y z _
#define QG_PL_VANG 3    // base angular velocity (body)      x y z _
#define QG_PL_WLIN 4    // qacc_warmstart[0:3] (world)
#define QG_PL_WANG 5    // qacc_warmstart[3:6]
#define QG_PL_TIME 6    // time (double in .x/.y), episode counter (int in .z)
#define QG_PL_AUX 7     // flags (.x as int: bit0 = first control cost set), first control cost (double in .y/.z)
#define QG_PL_LEG0 8    // 4 planes per leg:
                        //   +0: q0 q1 q2 qd0   +1: qd1 qd2 act0 act1   +2: act2 w0 w1 w2
                        //   +3: ctrl0 ctrl1 ctrl2 _      (data.ctrl; also previous_ctrl of control_cost)
#define QG_NPLANE (QG_PL_LEG0 + 4 * QG_NLEG)

struct QgJointC {
    float pos[3];      // body position in the parent body frame
    float Roff[9];     // body orientation in the parent body frame (row major)
    float com[3];      // body CoM in the body frame
    float mass;
    float I[6];        // inertia about the CoM, body axes: xx yy zz xy xz yz
    float q0, lo, hi;  // reference angle, range
    int limited;
    float damping, armature, invw_dof;
    // position servo acting on this joint
    int has_act, has_dyn, ctrl_limited, frc_limited;
    float gear, kp, b0, b1, b2;
    float inv_tau, act_fac;  // 1/tau, tau*(1-exp(-h/tau))
    float ctrl_lo, ctrl_hi, frc_lo, frc_hi;
};

struct QgGeomC {
    float pos[3];   // geom centre in the body frame
    float R[9];     // geom (mesh) frame -> body frame
    float half[3];  // per-axis max |coord| of the hull in the mesh frame (OBB cull)
    float margin, mu;
    float K, B;     // reference-acceleration stiffness / damping from solref
    float d0, dmax, width, mid, power;  // solimp
    float Rfac;     // 2 mu^2 (1+mu^2) * body_invweight0_trans  (pyramidal regulariser / ((1-imp)/imp))
    float tol2;     // (0.3 * rbound)^2 : minimum separation of extra plane-mesh contacts
    int vert0, nvert;  // slice of the vertex table
    int edge0;         // start of this mesh's neighbour lists in the int4-packed adjacency table (int4 units)
    int cedge0;        // same for the polytope-edge graph used by the hill climb
    int level;         // 0 = base, 1..3 = leg link
    int mesh;          // mesh id (row of dir_start)
};

struct alignas(16) QgModelC {
    float timestep, plane_z;
    double timestep_d;
    float grav[3];
    float base_mass, base_com[3], base_I[6];
    float base_damp[6], base_arm[6];
    float tol, scale;  // solver tolerance, 1/(meaninertia*nv)
    int max_iter, ls_iter, rule_first, integrator;
    int cone;                  // 0 pyramidal, 1 elliptic
    float impratio, mu_scale;  // elliptic: friction-row D = impratio * D_normal, regularised mu = fri * mu_scale (= 1/sqrt(impratio))
    float lim_K, lim_B, lim_d0, lim_dmax, lim_width, lim_mid, lim_power;
    float qpos0[19];
    QgJointC joint[QG_NLEG][QG_NLINK];
    int ngeom[QG_NLEG];
    int glev[QG_NLEG][5];  // geoms of lane l at tree level k are geom[l][glev[l][k] .. glev[l][k+1])
    QgGeomC geom[QG_NLEG][QG_MAXGEOM_LANE];
    // link_reach[l][k]: no geom of lane l at tree level k can be within its margin of the plane while the link origin is
    // higher than this above it (bounding sphere of the link's hulls about the link origin + the largest margin)
    float link_reach[QG_NLEG][4];
    int nvert;             // hull vertices over all meshes (float4 table staged in shared memory)
    int pad_;
    // hill-climb start vertex (local id) for the support search, indexed by the cube-map cell of the
    // search direction in the mesh frame
    unsigned short dir_start[QG_MAXMESH][QG_DIRCELLS];
};

struct QgStepOpts {
    double max_time;
    int flip_termination, auto_reset, max_iter, ls_iter;
    int random_yaw;
    double settling_time;   // data.time < settling_time -> action := joint centres (walking_quad.py:142-143)
    unsigned long long seed;
    long long env_offset;
    float reset_ctrl[12];
    int n_terms;
    int term_id[QG_MAX_TERMS_];
    double term_w[QG_MAX_TERMS_];
    double term_p[QG_MAX_TERMS_];
};

struct QgDebugOut {
    float *qacc, *qacc_smooth, *qfrc_bias, *M, *sensordata;
    int* counts;
};
