/* qg_oracle.h -- TEST INFRASTRUCTURE (see qg_oracle.c header).  float64 CPU restatement of
 * frame_skip x mj_step + sensor readout (/root/reference/src/envs/quadruped.py:153-167). */
#ifndef QG_ORACLE_H
#define QG_ORACLE_H
#include <stddef.h>

#define QGO_MAXNQ 32
#define QGO_MAXNV 24
#define QGO_MAXNU 16
#define QGO_MAXNBODY 16
#define QGO_MAXNGEOM 32
#define QGO_MAXNMESH 8
#define QGO_MAXVERT 2048
#define QGO_MAXEDGE 16384
#define QGO_MAXCON 100
#define QGO_MAXEFC 424
#define QGO_NSENSORDATA 33

typedef struct qgo_model {
    int nq, nv, nu, nbody, njnt, ngeom, nmesh, nsensordata, nvert;
    double opt_f[9]; /* timestep, gravity[3], tolerance, ls_tolerance, impratio, plane_z, meaninertia */
    int opt_i[5];    /* integrator, cone, iterations, ls_iterations, plane-mesh neighbour rule */
    int body_parent[QGO_MAXNBODY], body_dofadr[QGO_MAXNBODY], body_dofnum[QGO_MAXNBODY];
    double body_pos[3 * QGO_MAXNBODY], body_quat[4 * QGO_MAXNBODY], body_mass[QGO_MAXNBODY];
    double body_ipos[3 * QGO_MAXNBODY], body_inertia[6 * QGO_MAXNBODY], body_invweight0[2 * QGO_MAXNBODY];
    int jnt_type[QGO_MAXNBODY], jnt_body[QGO_MAXNBODY], jnt_qposadr[QGO_MAXNBODY], jnt_dofadr[QGO_MAXNBODY];
    int jnt_limited[QGO_MAXNBODY];
    double jnt_axis[3 * QGO_MAXNBODY], jnt_pos[3 * QGO_MAXNBODY], jnt_range[2 * QGO_MAXNBODY];
    double jnt_solref[2], jnt_solimp[5];
    double qpos0[QGO_MAXNQ];
    double dof_damping[QGO_MAXNV], dof_armature[QGO_MAXNV], dof_invweight0[QGO_MAXNV];
    int dof_body[QGO_MAXNV], dof_parent[QGO_MAXNV];
    int act_dof[QGO_MAXNU], act_ctrllimited[QGO_MAXNU], act_frclimited[QGO_MAXNU];
    double act_gear[QGO_MAXNU], act_gain[QGO_MAXNU], act_bias[3 * QGO_MAXNU], act_tau[QGO_MAXNU];
    double act_ctrlrange[2 * QGO_MAXNU], act_frcrange[2 * QGO_MAXNU];
    int geom_body[QGO_MAXNGEOM], geom_mesh[QGO_MAXNGEOM];
    double geom_pos[3 * QGO_MAXNGEOM], geom_quat[4 * QGO_MAXNGEOM], geom_rbound[QGO_MAXNGEOM];
    double geom_margin[QGO_MAXNGEOM], geom_mu[QGO_MAXNGEOM], geom_solref[2 * QGO_MAXNGEOM];
    double geom_solimp[5 * QGO_MAXNGEOM];
    int mesh_vertadr[QGO_MAXNMESH], mesh_vertnum[QGO_MAXNMESH], mesh_edgeadr[QGO_MAXNMESH];
    double mesh_vert[3 * QGO_MAXVERT];
    int mesh_vert_edge[QGO_MAXVERT], mesh_edge[QGO_MAXEDGE];
} qgo_model;

typedef struct qgo_data {
    /* state (mjData) */
    double qpos[QGO_MAXNQ], qvel[QGO_MAXNV], act[QGO_MAXNU], ctrl[QGO_MAXNU], qacc_warmstart[QGO_MAXNV];
    double time;
    /* outputs of the last forward pass */
    double qacc[QGO_MAXNV], qacc_smooth[QGO_MAXNV], sensordata[QGO_NSENSORDATA];
    double xpos[3 * QGO_MAXNBODY], xmat[9 * QGO_MAXNBODY], xipos[3 * QGO_MAXNBODY];
    double xanchor[3 * QGO_MAXNBODY], xaxis[3 * QGO_MAXNBODY], com[3];
    double cinert[10 * QGO_MAXNBODY], cdof[6 * QGO_MAXNV], cdof_dot[6 * QGO_MAXNV], cvel[6 * QGO_MAXNBODY];
    double M[QGO_MAXNV * QGO_MAXNV], L[QGO_MAXNV * QGO_MAXNV];
    double qfrc_bias[QGO_MAXNV], qfrc_passive[QGO_MAXNV], qfrc_actuator[QGO_MAXNV], qfrc_smooth[QGO_MAXNV];
    double qfrc_constraint[QGO_MAXNV], act_dot[QGO_MAXNU], actuator_force[QGO_MAXNU];
    int act_clamped[QGO_MAXNU];
    int ncon, nefc, nlimit, solver_niter, ls_evals, nvert_tested, warnings, pad_;
    double solver_cost;
    double con_pos[3 * QGO_MAXCON], con_vert[3 * QGO_MAXCON], con_dist[QGO_MAXCON];
    int con_geom[QGO_MAXCON], con_vertid[QGO_MAXCON];
    double efc_J[QGO_MAXEFC * QGO_MAXNV], efc_pos[QGO_MAXEFC], efc_margin[QGO_MAXEFC];
    double efc_diagApprox[QGO_MAXEFC], efc_K[QGO_MAXEFC], efc_B[QGO_MAXEFC], efc_imp[QGO_MAXEFC];
    double efc_R[QGO_MAXEFC], efc_D[QGO_MAXEFC], efc_vel[QGO_MAXEFC], efc_aref[QGO_MAXEFC];
    double efc_force[QGO_MAXEFC];
    int efc_type[QGO_MAXEFC], efc_id[QGO_MAXEFC];
} qgo_data;

int qgo_model_load(const void* blob, size_t n, qgo_model** out);
void qgo_model_free(qgo_model* m);
void qgo_reset(const qgo_model* m, qgo_data* d);
void qgo_forward(const qgo_model* m, qgo_data* d);
void qgo_step(const qgo_model* m, qgo_data* d);
void qgo_env_step(const qgo_model* m, qgo_data* d, const double* action, int frame_skip);
void qgo_rollout(const qgo_model* m, qgo_data* d, int n_envs, const double* actions, int n_steps,
                 int frame_skip, double max_time, int auto_reset, double* obs_out);
size_t qgo_sizeof_data(void);
size_t qgo_sizeof_model(void);
#endif
