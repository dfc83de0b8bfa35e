/*
 * ndp_nmpc.h -- C ABI of the B200-native batched NMPC engine (libndp_nmpc_b200.so).
 *
 * Drop-in boundary for the per-step hot loop of Li-Jinjie/ndp_nmpc_qd.  The reference
 * reaches its solver through acados_template's ctypes binding of a generated shared
 * library (AcadosOcpSolver, constructed at
 * ndp_nmpc/scripts/nmpc_ctl/nmpc_body_rate_ctl.py:84); the entry points below are what
 * that binding would bind instead.  Each one cites the reference call it replaces.
 *
 * Conventions
 *   - plain C, no torch / C++ types.  All `dev` pointers are DEVICE pointers borrowed from
 *     the caller (e.g. tensor.data_ptr()); the library never frees caller memory and owns
 *     its iterate / workspace.  `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - element type of every `dev` array is the engine precision chosen at ndp_create
 *     (float for NDP_F32, double for NDP_F64) unless stated otherwise.
 *   - every function returns 0 on success, <0 on API misuse (NDP_E_*), >0 = cudaError_t.
 *     ndp_last_error() gives a message.  Per-problem solver status uses acados' codes
 *     (0 success, 1 NaN, 2 max iter, 3 min step, 4 QP failure), as tested by
 *     nmpc_body_rate_ctl.py:109-110.
 *   - a handle is not re-entrant; calls on one handle are ordered on the given stream, so a
 *     ndp_get issued while a solve is in flight returns a consistent snapshot (the
 *     reference's viz thread reads solver.get(i,"x") concurrently, nmpc_node.py:233-237).
 *   - batched layout: [B][stage][component], contiguous per problem.
 */
#ifndef NDP_NMPC_H
#define NDP_NMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDP_NX 10 /* state  (p, v, qw qx qy qz)   nmpc_body_rate_ctl.py:120-130 */
#define NDP_NU 4  /* input  (wx, wy, wz, c)       nmpc_body_rate_ctl.py:139-144 */
#define NDP_NY 14 /* cost output y = [x-part; u]  nmpc_body_rate_ctl.py:168-180 */
#define NDP_NP 8  /* parameter slot: q_r(4), f(3), pad   ndp_nmpc_body_rate_ctl.py:197 */
#define NDP_N_MAX 128

enum ndp_precision { NDP_F32 = 0, NDP_F64 = 1 };

/* fields of ndp_set / ndp_get -- the strings of AcadosOcpSolver.set/get used by the
 * reference ("x","u","yref","p": nmpc_body_rate_ctl.py:89-104; get "x": nmpc_node.py:237) */
enum ndp_field {
    NDP_FIELD_X = 0,    /* [10]  iterate state, stages 0..N     */
    NDP_FIELD_U = 1,    /* [4]   iterate input, stages 0..N-1   */
    NDP_FIELD_YREF = 2, /* [14]  stages 0..N-1; [10] at stage N */
    NDP_FIELD_P = 3     /* [np]  np = 4 (q_r) or 7 (q_r, f)     */
};

enum ndp_error {
    NDP_OK = 0,
    NDP_E_ARG = -1,    /* bad argument (null pointer, unknown field, stage out of range) */
    NDP_E_CONFIG = -2, /* unsupported configuration */
    NDP_E_ALLOC = -3,
    NDP_E_STATE = -4
};

/* Mirrors the constants the reference's controller reads from params/nmpc_params.py:9-35
 * and params/fhnp_params.py:9-19 when it builds the AcadosOcp (nmpc_body_rate_ctl.py:36-80). */
typedef struct ndp_config {
    int32_t N;          /* shooting intervals (N_node)                   */
    int32_t precision;  /* enum ndp_precision                            */
    int32_t batch;      /* number of independent problems in the handle  */
    int32_t np;         /* 4 = NMPCBodyRateController, 7 = NDPNMPCBodyRateController */
    double T;           /* horizon length [s] (T_horizon); h = T / N     */
    double mass;        /* [kg]                                          */
    double gravity;     /* [m/s^2]                                       */
    double Q[NDP_NX];   /* diag of Q   (W = blkdiag(Q,R), W_e = Q)       */
    double R[NDP_NU];   /* diag of R                                     */
    double u_min[NDP_NU], u_max[NDP_NU]; /* lbu/ubu, stages 0..N-1       */
    double v_min[3], v_max[3];           /* lbx/ubx on idx 3,4,5, stages 1..N-1 */
    int32_t ipm_max_iter; /* acados qp_solver_iter_max (50)              */
    int32_t polish_max;   /* active-set refinement rounds after the IPM  */
    double ipm_tol_mu;    /* IPM complementarity target (<=0: precision default) */
    int32_t active_set_first; /* active-set rounds tried from the unconstrained step's violated bounds before
                               * the IPM (0: IPM first; default 8); both routes end at the same QP solution */
    int32_t active_set_warm;  /* 1: a solve that ends with active input bounds leaves them as the first guess of the
                               * problem's next solve (skips the unconstrained sweep; fewer sweeps on average in
                               * saturated closed loops, but a stale guess can cost the slowest problem more rounds,
                               * so the default is 0); ndp_reset clears the guess */
} ndp_config;

typedef struct ndp_handle ndp_handle;

/* Fill cfg with the reference's constants (N=20, T=2, weights, bounds, mass, gravity). */
void ndp_default_config(ndp_config* cfg);

/* AcadosOcpSolver(ocp, json_file, build) -- nmpc_body_rate_ctl.py:84.  Allocates the iterate
 * (zero-initialised like acados), reference/parameter storage and workspace on the current
 * CUDA device. */
int ndp_create(const ndp_config* cfg, ndp_handle** out);
int ndp_destroy(ndp_handle* h);

/* solver.set(stage, field, value) -- nmpc_body_rate_ctl.py:89-91,97-104.
 * stage >= 0: dev -> [B][dim] with row stride ld (elements), one stage for all problems.
 * stage == -1: dev -> [B][n_stages][dim] contiguous, all stages at once. */
int ndp_set(ndp_handle* h, int field, int stage, const void* dev, int64_t ld, void* stream);

/* solver.get(stage, field) -- nmpc_node.py:235-237 (same addressing as ndp_set). */
int ndp_get(ndp_handle* h, int field, int stage, void* dev, int64_t ld, void* stream);

/* controller.reset(xr, ur) -- nmpc_body_rate_ctl.py:86-91: iterate <- (xr[B][N+1][10], ur[B][N][4]). */
int ndp_reset(ndp_handle* h, const void* xr_dev, const void* ur_dev, void* stream);

/* The 42 solver.set calls of controller.update() in one launch --
 * nmpc_body_rate_ctl.py:95-104 / ndp_nmpc_body_rate_ctl.py:93-104:
 * yref_k = [xr_k; ur_k], p_k = [xr_k[6:10]; f_k].  f_dev may be NULL (zeros / np == 4). */
int ndp_set_reference(ndp_handle* h, const void* xr_dev, const void* ur_dev, const void* f_dev, void* stream);

/* The acados-style mirror's whole solve_for_x0 from PINNED HOST buffers in engine precision (latency path of the
 * one-problem controller, nmpc_body_rate_ctl.py:95-107): x0 [B][10], yref [B][N+1][14] (stage rows [xr; ur], terminal
 * row: first 10), p [B][N+1][8] (q_r(4) f(3) pad) -> [upload of yref / p when upload_ref != 0,] one SQP_RTI step,
 * u0 [B][4] and status [B] back in host memory.  Synchronous.  The sequence is captured once per set of buffers into a
 * CUDA graph and replayed; small batches read x0 / write u0 over PCIe from the kernel itself. */
int ndp_solve_host(ndp_handle* h, const void* x0_host, const void* yref_host, const void* p_host, int upload_ref,
                   void* u0_host, int32_t* status_host);

/* u0 = solver.solve_for_x0(x0) -- nmpc_body_rate_ctl.py:107: one SQP_RTI step for every problem
 * (x0 constraint, linearise, QP, full step).  x0_dev [B][10], u0_dev [B][4] (may be NULL). */
int ndp_solve(ndp_handle* h, const void* x0_dev, void* u0_dev, void* stream);

/* ndp_update_ex flag: f_dev is written by the kernel launched immediately before on the same stream (the downwash MLP)
 * and every other input is older than that kernel.  The solve is then launched as a programmatic dependent of it: its
 * prologue and the staging of (iterate, x0, xr, ur) run while the MLP drains, and it synchronises on the MLP
 * (griddepcontrol.wait) only before it reads f. */
#define NDP_UPDATE_F_FROM_PREVIOUS_KERNEL 1
int ndp_update_ex(ndp_handle* h, const void* x0_dev, const void* xr_dev, const void* ur_dev, const void* f_dev, void* u0_dev,
                  int flags, void* stream);

/* controller.update(x0, xr, ur[, f]) in ONE launch -- nmpc_body_rate_ctl.py:93-112: the reference
 * upload of ndp_set_reference fused into the solve (yref / p are stored as if set).  f_dev may be NULL. */
int ndp_update(ndp_handle* h, const void* x0_dev, const void* xr_dev, const void* ur_dev, const void* f_dev,
               void* u0_dev, void* stream);

/* solver.status -- nmpc_body_rate_ctl.py:109.  status_dev: int32 [B]. */
int ndp_status(ndp_handle* h, int32_t* status_dev, void* stream);

/* Per-problem solve statistics: int32 [B][4] = {Riccati factorisations, IPM iterations,
 * active-set rounds, active bounds at the solution}. */
int ndp_stats(ndp_handle* h, int32_t* stats_dev, void* stream);

/* Number of kernel launches issued on behalf of this handle so far (two per solve: the nominal SQP-RTI kernel and the
 * constrained-QP kernel that takes the problems whose unconstrained step leaves its box; the latter exits at once when
 * there are none). */
int64_t ndp_launch_count(const ndp_handle* h);

/* Measurement aid: with enable != 0 every solve records CUDA events on its stream before, between and after its two
 * kernels; ndp_last_kernel_ms waits for the last solve and returns the two device durations (the roofline figure of
 * bench.py is the nominal kernel's own time).  The event between the kernels costs the second one its programmatic
 * overlap, so leave this off outside measurements. */
int ndp_kernel_timing(ndp_handle* h, int enable);
int ndp_last_kernel_ms(ndp_handle* h, float* nominal_ms, float* constrained_ms);

const char* ndp_last_error(void);

/* ---- batched RK4 integrator with forward sensitivities (standalone entry point) ----
 * acados ERK sim with sens_forw (integrator_type="ERK", nmpc_body_rate_ctl.py:76).
 * x[M][10], u[M][4], f[M][3] (may be NULL) -> xn[M][10], AB[M][10][14] = [S_x S_u] row-major. */
int ndp_rk4_sens(int precision, int64_t M, double h, double mass, double gravity, const void* x_dev,
                 const void* u_dev, const void* f_dev, void* xn_dev, void* AB_dev, void* stream);

/* ---- downwash MLP (6-128-64-128-3, ReLU) ----
 * DownwashNN.__init__/update -- dnwash_nn_est/downwash_nn.py:10-29, net: nn_net.py:7-18. */
typedef struct ndp_mlp ndp_mlp;

/* Weights are HOST pointers, row-major [out][in] fp32 as in the torch state_dict
 * ("0.weight" [128][6], "2.weight" [64][128], "4.weight" [128][64], "6.weight" [3][128]). */
int ndp_mlp_create(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                   const float* b3, const float* W4, const float* b4, ndp_mlp** out);
int ndp_mlp_destroy(ndp_mlp* m);

/* Fused feature construction + MLP for P (ego, neighbour) pairs:
 *   f[p][k][:] = gate_p * MLP( (other[p][k] - ego[p][k])[0:6] ),  k = 0..n_nodes-1
 * ego/other: [P][n_nodes][10] in `precision`; gate_xy (may be NULL = always on): [P][2] ego
 * odometry x,y -- the force is zeroed when the neighbour's node-0 horizontal distance to it is
 * >= r_horiz (ndp_nmpc_leader_node.py:65-76, params/downwash_params.py:10).
 * out: [P][n_nodes][3] in `precision`.  accumulate != 0 adds into out (sum over neighbours).
 * path: 0 = auto (<= 512 rows: 3, >= 2048 rows: 2, else 1), 1 = CUDA-core fp32 tile kernel, 2 = tcgen05 tensor-core
 * kernel, 3 = latency kernel, one CTA per row (the reference's own 21-row call). */
int ndp_mlp_forward_pairs(ndp_mlp* m, int precision, int64_t P, int32_t n_nodes, const void* ego_dev,
                          const void* other_dev, const void* gate_xy_dev, double r_horiz, void* out_dev,
                          int accumulate, int path, void* stream);

/* Same with an explicit row stride of `other`: other_ld = 10 (full state rows, as above) or 6
 * ([P][n_nodes][6]: only the position+velocity columns, the slice downwash_nn.py:24 reads). */
int ndp_mlp_forward_pairs_ex(ndp_mlp* m, int precision, int64_t P, int32_t n_nodes, const void* ego_dev,
                             const void* other_dev, int32_t other_ld, const void* gate_xy_dev, double r_horiz,
                             void* out_dev, int accumulate, int path, void* stream);

/* Plain rows: in [M][6] fp32 -> out [M][3] fp32 (the nn.Sequential itself). */
int ndp_mlp_forward_rows(ndp_mlp* m, int64_t M, const float* in_dev, float* out_dev, int path, void* stream);

/* Swarm: all-pairs gated sum.  traj [n_all][n_nodes][6] fp32 (positions+velocities of every quad's
 * reference horizon, e.g. the all-gathered tensor), egos are rows [ego_begin, ego_begin+n_ego).
 * odom_xy [n_ego][2] (may be NULL: use the ego's own node 0).  out [n_ego][n_nodes][3] in `precision`:
 *   f_i = sum over j != i with |xy_j(node 0) - xy_i| < r_horiz of MLP(traj_j - traj_i). */
int ndp_mlp_forward_swarm(ndp_mlp* m, int precision, int64_t n_all, int64_t ego_begin, int64_t n_ego,
                          int32_t n_nodes, const float* traj_dev, const float* odom_xy_dev, double r_horiz,
                          void* out_dev, int path, void* stream);

/* Same, with the trajectory tensor split in n_parts (<= 16) equal contiguous parts of part_rows quads that may
 * live on different GPUs: part_ptrs is a HOST array of device pointers, part r holding rows
 * [r * part_rows, (r+1) * part_rows) -- the local shard and the peers' shards mapped over NVLink (symmetric
 * memory).  The kernels read neighbour horizons directly from the owning GPU: the exchange of the coupled
 * swarm is fused into the gated feature construction instead of a separate all-gather. */
int ndp_mlp_forward_swarm_parts(ndp_mlp* m, int precision, int32_t n_parts, const float* const* part_ptrs, int64_t part_rows,
                                int64_t n_all, int64_t ego_begin, int64_t n_ego, int32_t n_nodes, const float* odom_xy_dev,
                                double r_horiz, void* out_dev, int path, void* stream);

/* Upper bound (pairs) up to which the swarm entry points size their pair buffers for the worst case
 * n_ego * (n_all - 1) and never read the pair count back (default 2 Mi pairs).  Above it they start from the
 * budget, read the 4-byte count back once per call and grow on demand. */
int ndp_mlp_set_pair_budget(ndp_mlp* m, int64_t max_pairs);

/* Swarm entry points: restrict the pairwise interaction to contiguous blocks of `group` quads (independent scenarios
 * batched side by side, like the `group` of ndp_plant_create); 0 (default): every quad sees every other. */
int ndp_mlp_set_group(ndp_mlp* m, int32_t group);

int64_t ndp_mlp_launch_count(const ndp_mlp* m);

/* ---- host-buffer step pipeline ----
 * One control step for the whole batch straight from HOST memory: what the ROS node does per timer
 * tick -- DownwashNN.update(other, ego_ref) (ndp_nmpc_leader_node.py:60-76, incl. its H2D/D2H,
 * downwash_nn.py:22-28) followed by controller.update(x0, xr, ur, f) (nmpc_node.py:202-209,
 * ndp_nmpc_body_rate_ctl.py:91-112) -- as ONE asynchronous submission: H2D copy of the step record,
 * MLP + RTI kernels, D2H copy of (u0, status).  `depth` slots of pinned host memory are owned by the
 * pipeline; the caller fills a slot's input arrays in place, submits it and later waits for it.
 * Uploads, kernels and downloads of different slots overlap (three streams); solves stay ordered
 * (the iterate is warm-started from the previous solve, nmpc_body_rate_ctl.py:86-112).
 * mlp == NULL: plain NMPC (np = 4), no neighbour arrays.  r_horiz: params/downwash_params.py:10. */
typedef struct ndp_pipeline ndp_pipeline;
int ndp_pipeline_create(ndp_handle* h, ndp_mlp* mlp, double r_horiz, int depth, ndp_pipeline** out);
int ndp_pipeline_destroy(ndp_pipeline* p);
/* Pinned HOST arrays of a slot, engine precision: x0 [B][10], xr [B][N+1][10], ur [B][N][4],
 * other [B][N+1][6] (neighbour horizon, position+velocity columns), gate_xy [B][2] (ego odometry x,y) -> u0 [B][4], status int32 [B].
 * Any output pointer may be NULL; other / gate_xy are NULL without an MLP. */
int ndp_pipeline_buffers(ndp_pipeline* p, int slot, void** x0, void** xr, void** ur, void** other,
                         void** gate_xy, void** u0, int32_t** status);
int ndp_pipeline_submit(ndp_pipeline* p, int slot); /* asynchronous */
int ndp_pipeline_wait(ndp_pipeline* p, int slot);   /* blocks until the slot's u0 / status are in host memory */
int ndp_pipeline_bytes(const ndp_pipeline* p, int64_t* h2d_bytes_per_step, int64_t* d2h_bytes_per_step);
void* ndp_pipeline_stream(ndp_pipeline* p);         /* the compute stream (cudaStream_t) */

/* ---- device-resident sliding reference lists ----
 * NMPCRefPublisher (pt_pub/pt_publisher.py:57-103): `len` = long_list_size = 101 points (x[10], u[4]) per quadrotor at
 * ts_nmpc; every tick get_nmpc_pts pops the front, appends ONE new point and returns every `stride` = 5th point
 * (params/nmpc_params.py:40-43): xr = list[0::5] (N+1 nodes), ur = the same without its last entry.  The lists stay on
 * the device as rings, so a control tick uploads one point per quadrotor (56 B) instead of the horizon (1 160 B);
 * with_other adds the neighbour's list (the 6 position / velocity columns DownwashNN reads, downwash_nn.py:24).
 * reset: _gen_long_list_w_traj -- x_long [B][len][10], u_long [B][len][4], other_long [B][len][6] (host or device memory).
 * push: new_x [B][10], new_u [B][4], new_other [B][6] -> xr [B][N+1][10], ur [B][N][4], other [B][N+1][6] (device). */
typedef struct ndp_longlist ndp_longlist;
int ndp_longlist_create(int precision, int64_t B, int32_t N, int32_t stride, int32_t len, int with_other, ndp_longlist** out);
int ndp_longlist_destroy(ndp_longlist* l);
int ndp_longlist_reset(ndp_longlist* l, const void* x_long, const void* u_long, const void* other_long, void* stream);
int ndp_longlist_push(ndp_longlist* l, const void* new_x_dev, const void* new_u_dev, const void* new_other_dev, void* xr_dev,
                      void* ur_dev, void* other_dev, void* stream);
int64_t ndp_longlist_launch_count(const ndp_longlist* l);
/* Step pipeline over such lists: a slot's input record is x0 [B][10], ONE new ego point (ndp_pipeline_buffers: xr ->
 * new_x [B][10], ur -> new_u [B][4]), the neighbour's new point (other -> [B][6]) and gate_xy [B][2] -- 128 B per problem
 * instead of 1 712 B; each submitted step pushes the points into the lists before the MLP / SQP-RTI kernels. */
int ndp_pipeline_create_ll(ndp_handle* h, ndp_mlp* mlp, double r_horiz, int depth, ndp_longlist* ll, ndp_pipeline** out);

/* ---- batched dop_sim quadrotor plant (closed-loop rollouts) ----
 * dop_sim/scripts/quadrotor/mul_quadrotors.py:19-50 MulQuadrotors(num_agent, ts_sim, ts_control, float64,
 * has_downwash, has_motor_model, has_battery).forward(ts_sim, ego_states[n,35,1], body_rate_cmd[n,4,1]).
 * state_dev: double [n][35] (index map params/def_mul_states.py:10-31), updated IN PLACE like the
 * reference; cmd_dev: double [n][4] = (roll, pitch, yaw rate [rad/s], throttle 0..1).
 * group: the pairwise downwash (a_dynamics/qd_dynamics.py:161-198) couples agents of the same contiguous
 * block of `group` agents (<= 0 or n: all pairs, the reference; smaller: independent Monte-Carlo scenarios). */
typedef struct ndp_plant ndp_plant;
int ndp_plant_create(int64_t n, double ts_sim, double ts_ctl, int has_downwash, int has_motor_model, int has_battery,
                     int64_t group, ndp_plant** out);
int ndp_plant_destroy(ndp_plant* p);
int ndp_plant_reset(ndp_plant* p, void* stream); /* PID states, control / simulation clocks */
/* MulQuadrotors.forward: autopilot when the control clock says so (every call while ctl_t > ts_ctl), then dynamics */
int ndp_plant_forward(ndp_plant* p, double ts_sim, double* state_dev, const double* cmd_dev, void* stream);
/* b_autopilot/atp_rate.py:60-112 AtpRate.forward (PID rate loops pid_control.py:36-65): stores the rotor commands */
int ndp_plant_autopilot(ndp_plant* p, const double* state_dev, const double* cmd_dev, double all_sim_t, void* stream);
/* a_dynamics/qd_dynamics.py:75-98 QdDynamics.forward with the stored rotor commands */
int ndp_plant_dynamics(ndp_plant* p, double dt, double* state_dev, void* stream);
/* pt_pub/pt_publisher.py:106-122 odom_2_nmpc_x: plant state -> NMPC x0 [n][10] = (p, v, qw, qx, qy, qz) in `precision` */
int ndp_plant_nmpc_x0(int64_t n, const double* state_dev, int precision, void* x0_dev, void* stream);
/* nmpc_node.py:273-283 nmpc_u_2_att_tgt: cmd = (u0[0:3], u0[3] * mass / k_throttle)  (thrust 0 if k_throttle == 0) */
int ndp_plant_cmd_from_u0(int64_t n, int precision, const void* u0_dev, double mass, double k_throttle, double* cmd_dev,
                          void* stream);
int64_t ndp_plant_launch_count(const ndp_plant* p);

/* ---- batched NMPC reference generation ----
 * ndp_nmpc/scripts/pt_pub/: NMPCRefPublisher.get_nmpc_pts (pt_publisher.py:78-103) = piecewise polynomial
 * evaluation (base_pt_publisher.py:81-148) + diff_flatness (pt_publisher.py:188-248) + traj_full_pt_2_x_u
 * (:124-147), for B problems at once and without leaving the device.
 * create: HOST arrays holding n_traj TrajCoefficients messages (ndp_nmpc/msg/TrajCoefficients.msg):
 *   seg_off [n_traj+1] (first segment index of each trajectory), t_cum: per trajectory n_seg+1 knot times,
 *   concatenated; cx, cy, cz [total_seg][8], cyaw [total_seg][4] ascending powers of the normalised segment
 *   time; final_pt [n_traj][3].
 * horizon: xr[b][k] / ur[b][k] = reference at t0[b] + k * th_pred of trajectory traj_id[b] (NULL: 0), position
 *   shifted by offset[b][0:3] (NULL: none; formation offsets, nmpc_follower_node.py:44-74).  t0 double [B],
 *   xr [B][N+1][10], ur [B][N][4] in `precision`. */
typedef struct ndp_refgen ndp_refgen;
int ndp_refgen_create(int32_t n_traj, const int32_t* seg_off, const double* t_cum, const double* cx, const double* cy,
                      const double* cz, const double* cyaw, const double* final_pt, ndp_refgen** out);
int ndp_refgen_destroy(ndp_refgen* g);
int ndp_refgen_horizon(ndp_refgen* g, int precision, int64_t B, const int32_t* traj_id_dev, const double* t0_dev, int32_t N,
                       double th_pred, const double* offset_dev, void* xr_dev, void* ur_dev, void* stream);
int64_t ndp_refgen_launch_count(const ndp_refgen* g);

/* ---- wire formats between the nodes, batched ----
 * PredXU (ndp_nmpc/msg/PredXU.msg:1-4): Float64MultiArray[] x, Float64MultiArray[] u.  One quadrotor's payload here is
 * the flat float64 array [x_0 .. x_N (10 each) | u_0 .. u_{N-1} (4 each)], ndp_predxu_len(N) = (N+1)*10 + N*4 doubles
 * (2 320 B at N = 20); msg_dev is [B][ndp_predxu_len(N)].
 * pack: do_pub_ref (nmpc_node.py:116-133) -- the reference horizon (xr [B][N+1][10], ur [B][N][4] in `precision`).
 * unpack: FollowerNode.sub_pred_callback (nmpc_follower_node.py:57-74) -- x rows get offset[b][0:3] (float64 [B][3],
 * the alpha-filtered formation offset; NULL: none, as in ndp_nmpc_leader_node.py:69-71) added to their position. */
int64_t ndp_predxu_len(int32_t N);
int ndp_predxu_pack(int precision, int64_t B, int32_t N, const void* xr_dev, const void* ur_dev, double* msg_dev, void* stream);
int ndp_predxu_unpack(int precision, int64_t B, int32_t N, const double* msg_dev, const double* offset_dev, void* xr_dev,
                      void* ur_dev, void* stream);

/* ---- hover-throttle estimator hook, batched on the device ----
 * HoverThrottleEstimator(ts).update(vz, throttle) -- hv_throttle_est/hover_throttle_estimator.py:15-53, called at 50 Hz
 * from nmpc_node.py:251-253 with vz = odometry twist.linear.z and throttle = the last AttitudeTarget thrust.
 * est_dev: float64 [n][8] = {vz[k-1], az[k-1], f_collect, k_throttle, P00, P01, P10, P11}; init sets x = (0, 50),
 * P = I (estimator_params.py:13).  vz / throttle: float64 with element strides vz_ld / throttle_ld (e.g. the plant
 * state column 15 with stride 35 and the command column 3 with stride 4).  k_throttle_dev [n] (may be NULL) receives
 * the estimate that nmpc_u_2_att_tgt divides by. */
int ndp_hover_throttle_init(int64_t n, double* est_dev, double* k_throttle_dev, void* stream);
int ndp_hover_throttle_update(int64_t n, double ts, const double* vz_dev, int64_t vz_ld, const double* throttle_dev,
                              int64_t throttle_ld, double* est_dev, double* k_throttle_dev, void* stream);
/* nmpc_u_2_att_tgt with one k_throttle per quadrotor (float64 [n] on the device) -- nmpc_node.py:273-283 */
int ndp_plant_cmd_from_u0_dev(int64_t n, int precision, const void* u0_dev, double mass, const double* k_throttle_dev,
                              double* cmd_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NDP_NMPC_H */
