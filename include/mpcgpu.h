/*
 * mpcgpu.h -- C ABI of the B200 batched SQP backend for mpc_planner's NLP solve loop.
 *
 * This is the drop-in boundary.  Everything below is plain C: pointers, sizes and
 * PODs; no C++/torch types.  The entry points are what a CUDA branch of the reference's
 * `MPCPlanner::Solver` (mpc_planner_solver/include/mpc_planner_solver/solver_interface.h:4-12,
 * acados_solver_interface.h:93-222) binds instead of the generated acados C API
 * (Solver_acados_create_capsule / _update_params / _solve / ocp_nlp_out_get ...; call sites
 * mpc_planner_solver/src/acados_solver_interface.cpp:17,33,124-142,149-155,164-195,279-283).
 *
 * Data layouts are the reference's own (`AcadosParameters`, acados_solver_interface.h:51-91):
 *   xinit           [nx]                      initial state
 *   x0              [(nu+nx)*(N+1)]           warm start, node-major, [u_k ; x_k] per node
 *   all_parameters  [npar*N]                  stage-major parameter vector
 *   xtraj           [nx*(N+1)], utraj [nu*N]  output (`AcadosOutput`, :127-147)
 * A batch is P such problems, concatenated solve-major.
 *
 * Return convention of every function: 0 = ok, negative = error (mpcgpu_strerror explains).
 * Per-solve exit codes follow the reference (`completeOneIteration`,
 * acados_solver_interface.cpp:197-203): 1 = success, 0 = failure, 2 = max iterations,
 * 3 = min step, 4 = QP failure.
 */
#ifndef MPCGPU_H
#define MPCGPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double mpcgpu_real;     /* every real number at this boundary is an IEEE-754 binary64 */

#define MPCGPU_MAX_NZ 9          /* nu + nx upper bound (bicycle: 3 + 6)                       */
#define MPCGPU_MAX_CON_MODULES 4 /* constraint blocks stacked in one problem                   */

/* Dynamics models, solver_generator/solver_model.py */
enum mpcgpu_model {
    MPCGPU_MODEL_UNICYCLE = 0,       /* ContouringSecondOrderUnicycleModel          :193-214, nx=5 nu=2 */
    MPCGPU_MODEL_UNICYCLE_SLACK = 1, /* ContouringSecondOrderUnicycleModelWithSlack :274-298, nx=6 nu=2 */
    MPCGPU_MODEL_BICYCLE = 2         /* BicycleModel2ndOrder                        :302-351, nx=6 nu=3 */
};

/* Path-tracking objective, mpc_planner_modules/scripts */
enum mpcgpu_objective {
    MPCGPU_OBJ_CONTOURING = 0,      /* contouring.py:48-98 (MPCC)                                 */
    MPCGPU_OBJ_CURVATURE_AWARE = 1, /* curvature_aware_contouring.py:48-105 (CA-MPC cost)         */
    MPCGPU_OBJ_GENERATED = 2,       /* stage cost emitted by mpc_planner_b200/codegen.py from the module scripts (any objective
                                       stack); only per-problem libraries built with that header have such a kernel          */
    MPCGPU_OBJ_GOAL = 3             /* MPCBase(a, w) + goal_module.py:26-41 (configuration_lmpcc): hand-restated in the CPU
                                       oracle only, as the checker of the GENERATED kernel of that stack                      */
};

/* Constraint blocks, in module order (solver_definition.py:36-49) */
enum mpcgpu_constraint_kind {
    MPCGPU_CON_LINEAR = 0,    /* guidance_constraints.py:95-110   a1 x + a2 y - b <= 0            */
    MPCGPU_CON_ELLIPSOID = 1, /* ellipsoid_constraints.py:65-119  d' M d >= 1                     */
    MPCGPU_CON_GAUSSIAN = 2,  /* gaussian_constraints.py:68-117   chance constraint >= 0          */
    MPCGPU_CON_DECOMP = 3,    /* decomp_constraints.py:68-98      a.disc - (b + slack) <= 0       */
    MPCGPU_CON_SCENARIO = 4,  /* scenario_constraints.py:64-94    a.disc - (b + slack) <= 0       */
    MPCGPU_CON_GENERATED = 5  /* all rows emitted by mpc_planner_b200/codegen.py (one block, con_count = nh; rows may depend on
                                 x, y, psi and the slack state only; their bound side comes with the generated code)         */
};

/*
 * The NLP handed to the solver: what `generate_acados_solver` (solver_generator/
 * generate_acados_solver.py:68-177) configures, as data.  Parameter indices are positions in
 * the per-stage parameter vector, in the order `define_parameters` (solver_definition.py:5-16)
 * registers them (objectives first, then constraints; duplicates registered once,
 * util/parameters.py:43-44).
 */
typedef struct mpcgpu_problem_desc {
    int model;     /* enum mpcgpu_model */
    int objective; /* enum mpcgpu_objective */
    int N;         /* shooting intervals (generate_acados_solver.py:88) */
    int nx, nu, npar, nh;
    double dt;     /* integrator_step; tf = N*dt (:143) */
    int erk_steps; /* sim_method_num_steps = 3 (:150); 4 stages per step (:149) */
    int n_segments; /* contouring/num_segments */

    /* objective parameter indices (-1 = absent) */
    int p_acceleration, p_angular_velocity, p_slack, p_velocity, p_reference_velocity;
    int weigh_velocity; /* 1 if MPCBase weighs v with w_v (v - v_ref)^2 (generate_jackalsimulator_solver.py:49-52) */
    int p_contour, p_lag;
    int p_spline; /* first of n_segments x {x_a,x_b,x_c,x_d,y_a,y_b,y_c,y_d,start} (contouring.py:33-44) */

    /* shared constraint parameters */
    int p_ego_disc_radius, p_ego_disc_offset; /* one disc (n_discs = 1) */

    int n_con_modules;
    int con_kind[MPCGPU_MAX_CON_MODULES];  /* enum mpcgpu_constraint_kind */
    int con_count[MPCGPU_MAX_CON_MODULES]; /* rows of h contributed */
    int con_param[MPCGPU_MAX_CON_MODULES]; /* index of the block's first own parameter */

    /* z = [u; x] box (solver_model.py lower_bound/upper_bound; generate_acados_solver.py:100-107) */
    double lower_bound[MPCGPU_MAX_NZ], upper_bound[MPCGPU_MAX_NZ];

    /* solver options mirrored from generate_acados_solver.py:143-173 and settings.yaml:14-18 */
    int n_sqp;          /* acados.iterations (10) */
    int qp_iter_max;    /* qp_solver_iter_max (50) */
    double qp_tol;      /* qp_tol (1e-5) */
    double reg_epsilon; /* MIRROR threshold (acados default 1e-4) */
    double qp_mu0;      /* IPM initial barrier (HPIPM BALANCE: 1e1) */
    double qp_thr0;     /* IPM initial slack floor (HPIPM: 1e-1) */
    int qp_warm_start;  /* qp_solver_warm_start (generate_acados_solver.py:173: 2 = primal and dual): RTI iterations > 0
                           start the IPM from the previous QP's multipliers */
    double qp_ws_thr;   /* warm start: slack floor t = max(row value, qp_ws_thr) */
    double qp_ws_mu;    /* warm start: multiplier floor lam = max(lam_prev, qp_ws_mu / t) */
    double qp_relax;    /* every inequality row is relaxed by this amount (1e-7) so that rows pinned exactly at their
                           bound (the slack state of the slack models: zero dynamics, x0 = 0 = lower bound) keep an
                           interior; same purpose as IPOPT's bound_relax_factor */
    double res_eq_fail; /* res_eq above this demotes success (acados_solver_interface.cpp:176-181) */
    int p_goal;         /* MPCGPU_OBJ_GOAL: first of goal_weight, goal_x, goal_y (goal_module.py:19-22); else -1 */
} mpcgpu_problem_desc;

/* Per-solve result record (`AcadosInfo`, acados_solver_interface.h:96-125) */
typedef struct mpcgpu_info {
    double pobj;    /* ocp_nlp_eval_cost at the final iterate (:167-168) */
    double res_eq;  /* max dynamics defect at the final iterate */
    double res_stat_qp; /* last QP: stationarity residual (inf-norm) */
    double res_comp_qp; /* last QP: complementarity residual (inf-norm) */
    int exit_code;  /* 1 = success ... (see top) */
    int qp_status;  /* 0 ok, 2 max iter, 3 min step, 4 NaN (acados numbering) */
    int sqp_iter;   /* RTI iterations executed */
    int qp_iter;    /* interior-point iterations executed, summed over the solve */
} mpcgpu_info;

typedef struct mpcgpu_solver *mpcgpu_handle;

/* Library / device introspection (no GPU needed for the first two) */
const char *mpcgpu_version(void);
const char *mpcgpu_strerror(int code);
int mpcgpu_device_count(void);

/* 1 if a kernel specialisation for this problem was compiled into the library, else 0.
 * Needs no GPU.  Replaces "does the generated Solver exist" of the reference build. */
int mpcgpu_problem_supported(const mpcgpu_problem_desc *desc);

/* Create a batched solver for up to max_batch simultaneous solves on CUDA device `device`.
 * Replaces Solver_acados_create_capsule + Solver_acados_create_with_discretization
 * (acados_solver_interface.cpp:17,33).  Device buffers are owned by the handle. */
int mpcgpu_create(const mpcgpu_problem_desc *desc, int device, int max_batch, mpcgpu_handle *out);
int mpcgpu_destroy(mpcgpu_handle h);

/* The problem, device and capacity a handle was created with (any output may be NULL). */
int mpcgpu_describe(mpcgpu_handle h, mpcgpu_problem_desc *desc, int *device, int *max_batch);

/* Host-buffer path (what Solver::solve() uses): upload -> solve -> download, synchronous.
 *   xinit[P*nx], x0[P*(nu+nx)*(N+1)], params[P*npar*N]  -> xtraj[P*nx*(N+1)], utraj[P*nu*N], info[P]
 * Replaces initializeOneIteration + n_sqp x solveOneIteration + completeOneIteration
 * (acados_solver_interface.cpp:86-204) for P solvers at once.  best (may be NULL) receives, per
 * selection group (see mpcgpu_set_group), the index of the successful solve with the smallest pobj
 * (FindBestPlanner, mpc_planner_modules/src/guidance_constraints.cpp:416-434), or -1 if none. */
int mpcgpu_solve_host(mpcgpu_handle h, int P, const double *xinit, const double *x0,
                      const double *params, double *xtraj, double *utraj, mpcgpu_info *info,
                      int *best);

/* As mpcgpu_solve_host with the persistent multipliers of the P solver slots travelling WITH the call (needs
 * mpcgpu_set_stateful(h, 1)): state[P * mpcgpu_state_size] is uploaded next to the inputs, used, and the updated multipliers
 * come back next to the outputs -- what a reference Solver's acados memory does across `operator=` and solve()
 * (acados_solver_interface.cpp:67-77), without the two synchronous round trips of mpcgpu_set_state / mpcgpu_get_state. */
int mpcgpu_solve_host_stateful(mpcgpu_handle h, int P, const double *xinit, const double *x0, const double *params,
                               double *state, double *xtraj, double *utraj, mpcgpu_info *info, int *best);

/* Device-buffer path: same arrays already resident in HBM; asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = default stream).  The call selects the handle's device.  It uses the
 * handle's scratch buffers (transposed parameters, persistent multipliers, timing events): ONE solve call may be in
 * flight per handle -- enqueue calls of one handle on one stream, or synchronise between streams; independent
 * concurrent solves need separate handles (as every reference `Solver` instance has its own acados memory). */
int mpcgpu_solve_device(mpcgpu_handle h, int P, const double *d_xinit, const double *d_x0,
                        const double *d_params, double *d_xtraj, double *d_utraj,
                        mpcgpu_info *d_info, int *d_best, void *stream);

/* T-MPC++ control steps with DEVICE-SIDE construction of the topology halfspaces and parameter packing.
 * Replaces, for S control steps x G guesses each, the per-planner work of GuidanceConstraints::optimize before
 * solve(): `*local_solver = *_solver` (guidance_constraints.cpp:298) and LinearizedConstraints::update /
 * projectToSafety / setParameters in topology mode (linearized_constraints.cpp:49-189).
 *   xinit[S*nx], x0[S*G*(nu+nx)*(N+1)] (the guesses' warm starts = guidance trajectories),
 *   scene_params[S*N*npar]: the MAIN solver's parameters of each control step (weights, spline, obstacle
 *       predictions; of the lin_constraint slots only the rows behind the obstacles are read: static halfspaces of
 *       guided planners, linearized_constraints.cpp:111-123 -- dummies (1, 0, x0 + 100) if there are none) -- uploaded
 *       once per step, not once per guess;
 *   guided[G]: 1 = guided planner (halfspaces a = (o - p)/|o - p|, b = a.o - (1e-3 + robot_radius) from its warm
 *       start p_k, after projecting p_k out of the obstacle discs), 0 = non-guided planner (dummy halfspaces
 *       (1, 0, x0 + 100) everywhere); stage 0 always gets the dummies.
 * Needs a problem with a LINEAR block followed by an ELLIPSOID or GAUSSIAN block.  Outputs as
 * mpcgpu_solve_host with P = S*G; best[S] is the winner within each control step. */
int mpcgpu_solve_tmpc_host(mpcgpu_handle h, int S, int G, const double *xinit, const double *x0,
                           const double *scene_params, const int *guided, double robot_radius, double *xtraj,
                           double *utraj, mpcgpu_info *info, int *best);

/* T-MPC++ control steps from a COMPACT scene description: the whole per-stage parameter vector of every guess is built on
 * the device.  Replaces, for S control steps x G guesses, the host-side stage-by-stage writes of MPCBaseModule /
 * Contouring / EllipsoidConstraints / GaussianConstraints / DecompConstraints ::setParameters (mpc_base.cpp:23-35,
 * contouring.cpp:94-124, ellipsoid_constraints.cpp:34-90, gaussian_constraints.cpp:31-79, decomp_constraints.cpp:90-114,
 * 150-189) and LinearizedConstraints::update / projectToSafety / setParameters (linearized_constraints.cpp:49-189); the
 * host uploads what those functions READ (RealTimeData / ModuleData), not what they write:
 *   base_params[S*npar]      the stage-invariant part of one stage's parameter vector: weights, spline coefficients of the
 *                            segment window, ego disc radius / offset (the constraint blocks' slots are ignored)
 *   obstacles[S*n_obs*N*5]   prediction step j = 0..N-1 of obstacle o: x, y, psi, major, minor; step j feeds stage j + 1,
 *                            stage 0 gets the reference's far-away dummies
 *   obstacle_radius[S*n_obs], obstacle_flags[S*n_obs] (bit 0: Gaussian prediction, else deterministic: major = minor = 0,
 *                            chi = 1; bit 1: static obstacle: Gaussian block uses 0.001 for its axes), risk (chi =
 *                            ExponentialQuantile(0.5, 1 - risk) for the ellipsoid block, `risk` for the Gaussian block)
 *   halfspaces[S*N*n_halfspaces*3] or NULL   rows (a1, a2, b) of the DECOMP block of stage k (polyhedron k - 1), the
 *                            first zero / NaN normal ends a stage's list, dummies (1, 0, x0 + 100) behind it
 *   xinit, x0, guided, robot_radius, outputs: as mpcgpu_solve_tmpc_host.
 * Every block kind except SCENARIO (built by the external scenario_module) is supported.  Per control step the upload is
 * npar + 5 N n_obs + 2 n_obs (+ 3 N n_halfspaces) doubles instead of N * npar. */
int mpcgpu_solve_tmpc_scene_host(mpcgpu_handle h, int S, int G, const double *xinit, const double *x0, const double *base_params,
                                 int n_obs, const double *obstacles, const double *obstacle_radius, const int *obstacle_flags,
                                 double risk, int n_halfspaces, const double *halfspaces, const int *guided, double robot_radius,
                                 double *xtraj, double *utraj, mpcgpu_info *info, int *best);

/* Debug / parity: the parameter blocks the device-side packing of the last T-MPC call left in HBM, in the reference's
 * stage-major layout params[P][N][npar]. */
int mpcgpu_get_packed_params(mpcgpu_handle h, int P, double *params);

/* Persistent NLP multipliers ("stateful" solver slots).  In the reference every LocalPlanner owns a Solver whose
 * acados memory survives `*local_solver = *_solver` (operator= copies parameters only,
 * acados_solver_interface.cpp:67-77): the multipliers of its previous solve enter the first exact Hessian of the next
 * control step.  With mpcgpu_set_stateful(h, 1) solve slot p (the p-th problem of every solve call on this handle)
 * keeps [pi (N+1) x nx | multipliers of h, N x nh] (mpcgpu_state_size doubles) in HBM between calls; a QP failure
 * clears the slot, as does mpcgpu_reset_state (Solver::reset, :79-84).  get/set copy slots to/from host memory. */
int mpcgpu_set_stateful(mpcgpu_handle h, int enable);
int mpcgpu_state_size(mpcgpu_handle h);
int mpcgpu_reset_state(mpcgpu_handle h, int first, int count);
int mpcgpu_get_state(mpcgpu_handle h, int first, int count, double *state);
int mpcgpu_set_state(mpcgpu_handle h, int first, int count, const double *state);

/* Selection groups: consecutive blocks of `group` solves are the guesses of one control step;
 * `best` / `d_best` then receive ceil(P/group) entries, each the within-group index of the winner
 * (or -1).  group = 0 (default) treats the whole batch as one group. */
int mpcgpu_set_group(mpcgpu_handle h, int group);

/* Solver::solve()'s timeout (mpc_planner_solver/src/acados_solver_interface.cpp:91-117, `_params.solver_timeout`,
 * set per control step at mpc_planner/src/planner.cpp:118 and guidance_constraints.cpp:258,333): after each SQP-RTI
 * iteration a solve stops iterating once (time since the start of the CALL) + (its average iteration time) >= seconds,
 * then completes as usual (trajectories, cost, exit code; mpcgpu_info.sqp_iter tells how many iterations ran).  Time is
 * read on the device (%globaltimer): the call's clock starts in its first kernel (parameter transposition / packing, i.e.
 * after the host-to-device copies of the host path have landed), so the later waves of a batch larger than the GPU holds at
 * once do not get a fresh budget; it applies to every following solve call on the handle.
 * seconds <= 0 disables the check (default) -- note that the reference with solver_timeout = 0 stops after one iteration. */
int mpcgpu_set_timeout(mpcgpu_handle h, double seconds);

/* Pinned host staging buffers for the host path (optional; plain malloc'd memory also works). */
int mpcgpu_host_alloc(void **ptr, unsigned long long bytes);
int mpcgpu_host_free(void *ptr);

/* Kernel launches issued by this handle so far (bench.py reports them). */
long long mpcgpu_launch_count(mpcgpu_handle h);

/* Average device time (ms, CUDA events on the launching stream) of the SQP kernel over the
 * launches since the last call; resets the accumulator.  Only counted when timing is enabled. */
int mpcgpu_set_timing(mpcgpu_handle h, int enabled);
double mpcgpu_kernel_ms(mpcgpu_handle h, int *n_launches);

/* Debug / parity entry points: evaluate the stage functions and their derivatives on the device
 * for M independent (z, p[, lam, pi]) samples.  Outputs (any may be NULL):
 *   cost[M], cost_grad[M*nz], cost_hess[M*nz*nz]   dt-unscaled l(z,p)
 *   h[M*nh], h_jac[M*nh*nz], h_hess[M*nz*nz]       sum_i lam_i * hess h_i
 *   xnext[M*nx], AB[M*nx*nz] (d xnext / d [u;x]), dyn_hess[M*nz*nz] = sum_j pi_j * hess xnext_j */
int mpcgpu_eval_stage(mpcgpu_handle h, int M, const double *z, const double *p, const double *lam,
                      const double *pi, double *cost, double *cost_grad, double *cost_hess,
                      double *hval, double *h_jac, double *h_hess, double *xnext, double *AB,
                      double *dyn_hess);

/* Measure the FP64 FMA peak of the device (dependent-free FMA chains on all SMs); GFLOP/s. */
double mpcgpu_measure_fp64_gflops(int device);

#ifdef __cplusplus
}
#endif
#endif /* MPCGPU_H */
