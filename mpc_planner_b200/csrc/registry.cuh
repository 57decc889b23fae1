// registry.cuh -- the set of problem shapes compiled into the library ("generated solvers").
// The reference generates one acados solver per module stack offline
// (solver_generator/generate_solver.py:13-61); here every supported stack is one template
// instantiation of the kernels, selected at run time by matching mpcgpu_problem_desc.
#pragma once

#include "sqp_core.cuh"

namespace mpcgpu {

#ifdef MPCGPU_GENERATED_REGISTRY
// A per-problem library: `python -m mpc_planner_b200.generate --build` emits a header with the Config<> typedef(s) of
// ONE module stack and its own MPCGPU_CONFIG_LIST, and compiles this translation unit against it
// (the analogue of generate_acados_solver.py producing one solver library per stack).
#include MPCGPU_GENERATED_REGISTRY
#else
// cfg 1: MPCBase + Contouring + Ellipsoid(4)   generate_jackalsimulator_solver.py:61-66
typedef Config<MPCGPU_MODEL_UNICYCLE, MPCGPU_OBJ_CONTOURING, 20, 5, MPCGPU_CON_ELLIPSOID, 4> cfg_basic_n20_o4;
// cfg 2/5: T-MPC: Guidance(linear 8 + Ellipsoid 8)   :94-103
typedef Config<MPCGPU_MODEL_UNICYCLE, MPCGPU_OBJ_CONTOURING, 20, 5, MPCGPU_CON_LINEAR, 8, MPCGPU_CON_ELLIPSOID, 8> cfg_tmpc_n20_o8;
// T-MPC with the CC-MPC submodule (generate_jackal_solver.py:53-73)
typedef Config<MPCGPU_MODEL_UNICYCLE, MPCGPU_OBJ_CONTOURING, 20, 5, MPCGPU_CON_LINEAR, 8, MPCGPU_CON_GAUSSIAN, 8> cfg_tmpc_gauss_n20_o8;
// shipped jackalsimulator default: N=30, 12 obstacles (settings.yaml:2,34)
typedef Config<MPCGPU_MODEL_UNICYCLE, MPCGPU_OBJ_CONTOURING, 30, 5, MPCGPU_CON_LINEAR, 12, MPCGPU_CON_ELLIPSOID, 12> cfg_tmpc_n30_o12;
// cfg 3: slack model + CA cost + Guidance + Decomp   generate_rosnavigation_solver.py:91-112
typedef Config<MPCGPU_MODEL_UNICYCLE_SLACK, MPCGPU_OBJ_CURVATURE_AWARE, 30, 5, MPCGPU_CON_LINEAR, 12, MPCGPU_CON_ELLIPSOID, 12, MPCGPU_CON_DECOMP, 12> cfg_tmpc_decomp_n30_o12;
// cfg 4: SH-MPC scenario halfspaces   generate_jackalsimulator_solver.py:69-91
typedef Config<MPCGPU_MODEL_UNICYCLE_SLACK, MPCGPU_OBJ_CONTOURING, 20, 5, MPCGPU_CON_SCENARIO, 24> cfg_safe_horizon_n20;

// BASELINE configs[0] as written: BicycleModel2ndOrder (solver_model.py:302-351) + MPCC + 4 ellipsoidal obstacles, N = 20
typedef Config<MPCGPU_MODEL_BICYCLE, MPCGPU_OBJ_CONTOURING, 20, 5, MPCGPU_CON_ELLIPSOID, 4> cfg_bicycle_n20_o4;

#define MPCGPU_CONFIG_LIST(X) \
    X(cfg_basic_n20_o4)       \
    X(cfg_bicycle_n20_o4)     \
    X(cfg_tmpc_n20_o8)        \
    X(cfg_tmpc_gauss_n20_o8)  \
    X(cfg_tmpc_n30_o12)       \
    X(cfg_tmpc_decomp_n30_o12) \
    X(cfg_safe_horizon_n20)

#endif  // MPCGPU_GENERATED_REGISTRY

template <class CFG>
inline bool config_matches(const mpcgpu_problem_desc &d) {
    if (d.model != CFG::MODEL || d.objective != CFG::OBJ || d.N != CFG::N || d.n_segments != CFG::NSEG) return false;
    if (d.nx != CFG::NX || d.nu != CFG::NU || d.nh != CFG::NH || d.n_con_modules != CFG::NMOD) return false;
    if (d.erk_steps != 3) return false;
    for (int m = 0; m < CFG::NMOD; m++)
        if (d.con_kind[m] != CFG::kind(m) || d.con_count[m] != CFG::count(m)) return false;
    return true;
}

inline Layout make_layout(const mpcgpu_problem_desc &d) {
    Layout L;
    L.npar = d.npar;
    L.p_acceleration = d.p_acceleration;
    L.p_angular_velocity = d.p_angular_velocity;
    L.p_slack = d.p_slack;
    L.p_velocity = d.p_velocity;
    L.p_reference_velocity = d.p_reference_velocity;
    L.weigh_velocity = d.weigh_velocity;
    L.p_contour = d.p_contour;
    L.p_lag = d.p_lag;
    L.p_spline = d.p_spline;
    L.p_ego_disc_radius = d.p_ego_disc_radius;
    L.p_ego_disc_offset = d.p_ego_disc_offset;
    for (int m = 0; m < 3; m++) L.con_param[m] = (m < d.n_con_modules) ? d.con_param[m] : 0;
    for (int i = 0; i < MPCGPU_MAX_NZ; i++) {
        L.lower_bound[i] = d.lower_bound[i];
        L.upper_bound[i] = d.upper_bound[i];
    }
    L.dt = d.dt;
    L.n_sqp = d.n_sqp;
    L.qp_iter_max = d.qp_iter_max;
    L.qp_tol = d.qp_tol;
    L.reg_epsilon = d.reg_epsilon;
    L.qp_mu0 = d.qp_mu0;
    L.qp_thr0 = d.qp_thr0;
    L.res_eq_fail = d.res_eq_fail;
    L.qp_warm_start = d.qp_warm_start;
    L.qp_ws_thr = d.qp_ws_thr;
    L.qp_ws_mu = d.qp_ws_mu;
    L.qp_relax = d.qp_relax;
    L.timeout_s = 0.0;  // mpcgpu_set_timeout
    L.t_call = nullptr;
    return L;
}

// Dense stage-function evaluation for one sample (debug / parity entry point, see
// mpcgpu_eval_stage in include/mpcgpu.h).  Any output pointer may be null.
template <class CFG>
MPC_HD void eval_stage_sample(const Layout &L, const double *z, const double *p, const double *lam, const double *pi,
                              double *cost, double *cost_grad, double *cost_hess, double *hval, double *h_jac,
                              double *h_hess, double *xnext, double *AB, double *dyn_hess) {
    constexpr int NZ = CFG::NZ, NX = CFG::NX, NH = CFG::NH, NHP = NZ * (NZ + 1) / 2;
    ParamP pa{p};
    {
        double g[NZ], Hp[NHP];
        for (int i = 0; i < NZ; i++) g[i] = 0.0;
        for (int i = 0; i < NHP; i++) Hp[i] = 0.0;
        const double c = stage_cost<CFG>(pa, L, z, 1.0, g, Hp);
        if (cost) *cost = c;
        if (cost_grad) for (int i = 0; i < NZ; i++) cost_grad[i] = g[i];
        if (cost_hess) for (int i = 0; i < NZ; i++) for (int j = 0; j < NZ; j++) cost_hess[i * NZ + j] = Hp[trisym(i, j)];
    }
    if (NH > 0 && (hval || h_jac || h_hess)) {
        double hv[NH > 0 ? NH : 1], J[NH > 0 ? 3 * NH : 1], Hx[6] = {0, 0, 0, 0, 0, 0};
        stage_constraints<CFG, false>(pa, L, z, hv, J, lam, (lam && h_hess) ? Hx : nullptr);
        for (int j = 0; j < NH; j++) {
            if (hval) hval[j] = hv[j];
            if (h_jac) {
                for (int i = 0; i < NZ; i++) h_jac[j * NZ + i] = 0.0;
                h_jac[j * NZ + CFG::IX] = J[3 * j];
                h_jac[j * NZ + CFG::IY] = J[3 * j + 1];
                h_jac[j * NZ + CFG::IPSI] = J[3 * j + 2];
                if (CFG::ANY_SLACK_ROWS && row_has_slack<CFG>(j)) h_jac[j * NZ + (CFG::HAS_SLACK ? CFG::ISLACK : 0)] = -1.0;
            }
        }
        if (h_hess) {
            for (int i = 0; i < NZ * NZ; i++) h_hess[i] = 0.0;
            const int idx[3] = {CFG::IX, CFG::IY, CFG::IPSI};
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) h_hess[idx[a] * NZ + idx[b]] = Hx[trisym(a, b)];
        }
    }
    if (xnext || AB || dyn_hess) {
        double xn[NX], BA[NX * NZ], Hd[CFG::NHD];
        for (int e = 0; e < CFG::NHD; e++) Hd[e] = 0.0;
        model_step<CFG>(L, z, xn, BA, pi, (pi && dyn_hess) ? Hd : nullptr);
        if (xnext) for (int i = 0; i < NX; i++) xnext[i] = xn[i];
        if (AB) for (int i = 0; i < NX * NZ; i++) AB[i] = BA[i];
        if (dyn_hess) {
            for (int i = 0; i < NZ * NZ; i++) dyn_hess[i] = 0.0;
            for (int e = 0; e < CFG::NHD; e++) {
                int hi, hj;
                dyn_hess_pair<CFG>(e, hi, hj);
                dyn_hess[hi * NZ + hj] = Hd[e];
                dyn_hess[hj * NZ + hi] = Hd[e];
            }
        }
    }
}

}  // namespace mpcgpu
