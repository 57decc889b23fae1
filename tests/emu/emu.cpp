// emu.cpp -- HOST EMULATION of the CUDA kernel source (TEST INFRASTRUCTURE, not a product path).
// Compiles mpc_planner_b200/csrc/sqp_core.cuh with g++: every warp phase becomes a sequential
// loop over 32 lanes and "shared memory" is a heap buffer.  This lets the CPU test-suite exercise
// the kernel's algorithm (against the oracle) without a GPU.  The shipped library never uses it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mpcgpu { extern long g_mirror_slow, g_mirror_calls; }
#include "../../mpc_planner_b200/csrc/registry.cuh"

namespace mpcgpu { long g_mirror_slow = 0, g_mirror_calls = 0, g_mirror_general = 0; long long g_emu_clock_ns = 0, g_emu_clock_step_ns = 0; }
using namespace mpcgpu;
static double g_emu_timeout_s = 0.0;
// fake clock of the emulation: every clock read advances by `step_ns`; `timeout_s` as mpcgpu_set_timeout
extern "C" void emu_set_timeout(double timeout_s, long long step_ns) { g_emu_timeout_s = timeout_s; g_emu_clock_step_ns = step_ns; g_emu_clock_ns = 0; }
extern "C" void emu_mirror_stats(long *calls, long *slow) { *calls = g_mirror_calls; *slow = g_mirror_slow; }
extern "C" long emu_mirror_general_count() { return g_mirror_general; }
// round-robin schedule of the parallel Jacobi (csrc/sqp_core.cuh tournament_pair): pair j of round r for m indices
extern "C" void emu_tournament_pair(int m, int r, int j, int *p, int *q) { tournament_pair(m, r, j, *p, *q); }
// compare + select forms of fmax used by the row bodies (csrc/sqp_core.cuh)
extern "C" double emu_floor_at(double v, double floor) { return mpc_floor_at(v, floor); }
extern "C" double emu_raise(double acc, double v) { mpc_raise(acc, v); return acc; }
#ifndef MPCGPU_GENERATED_REGISTRY  // (the shipped stacks; a per-problem emulation has its own Config list)
// rows of h that linearisation thread `q` of 4 evaluates for a stack (kind, count) x 3: every row exactly once over q = 0..3
template <class CFG>
static void quarter_rows(int q, int *hit) {
    double z[CFG::NZ] = {0}, hv[CFG::NH > 0 ? CFG::NH : 1], J[CFG::NH > 0 ? 3 * CFG::NH : 1], w[CFG::NH > 0 ? CFG::NH : 1] = {0}, Hx[6] = {0};
    for (int j = 0; j < CFG::NH; j++) hv[j] = -12345.0;
    Layout L = {};
    L.p_ego_disc_offset = -1;
    std::vector<double> par(4096, 0.5);
    ParamP pa{par.data()};
    for (int m = 0; m < 3; m++) L.con_param[m] = 16 + 200 * m;
    L.p_ego_disc_radius = 1;
    stage_constraints<CFG, true>(pa, L, z, hv, J, w, Hx, 0, CFG::NH, q, 4);
    for (int j = 0; j < CFG::NH; j++) hit[j] += (hv[j] != -12345.0);
}
extern "C" int emu_quarter_rows(int cfg, int q, int *hit) {
    switch (cfg) {
        case 1: quarter_rows<cfg_basic_n20_o4>(q, hit); return cfg_basic_n20_o4::NH;
        case 2: quarter_rows<cfg_tmpc_n20_o8>(q, hit); return cfg_tmpc_n20_o8::NH;
        case 3: quarter_rows<cfg_tmpc_decomp_n30_o12>(q, hit); return cfg_tmpc_decomp_n30_o12::NH;
        case 4: quarter_rows<cfg_safe_horizon_n20>(q, hit); return cfg_safe_horizon_n20::NH;
        case 5: quarter_rows<cfg_tmpc_gauss_n20_o8>(q, hit); return cfg_tmpc_gauss_n20_o8::NH;
    }
    return -1;
}
#endif

template <class CFG>
static int solve_cfg(const mpcgpu_problem_desc &d, int P, const double *xinit, const double *x0, const double *params,
                     double *xtraj, double *utraj, mpcgpu_info *info, double *iterates, double *state) {
    typedef SL<CFG> S;
    Layout L = make_layout(d);
    L.timeout_s = g_emu_timeout_s;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU;
    std::vector<double> sm(S::TOTAL), pt((size_t)d.npar * pt_stride(N));
    for (int p = 0; p < P; p++) {
        const double *par = params + (size_t)p * d.npar * N;
        std::fill(pt.begin(), pt.end(), 0.0);
        for (int k = 0; k < N; k++)
            for (int i = 0; i < d.npar; i++) pt[(size_t)i * pt_stride(N) + k] = par[(size_t)k * d.npar + i];
        sqp_solve_cta<CFG>(L, sm.data(), xinit + (size_t)p * NX, x0 + (size_t)p * NZ * (N + 1), pt.data(),
                            xtraj + (size_t)p * NX * (N + 1), utraj + (size_t)p * NU * N, info + p,
                            iterates ? iterates + (size_t)p * d.n_sqp * (N + 1) * NZ : nullptr,
                            state ? state + (size_t)p * ((N + 1) * NX + N * CFG::NH) : nullptr);
    }
    return 0;
}

extern "C" {

int emu_problem_supported(const mpcgpu_problem_desc *d) {
#define X(C) \
    if (config_matches<C>(*d)) return 1;
    MPCGPU_CONFIG_LIST(X)
#undef X
    return 0;
}

int emu_solve(const mpcgpu_problem_desc *d, int P, const double *xinit, const double *x0, const double *params,
              double *xtraj, double *utraj, mpcgpu_info *info, double *iterates) {
#define X(C) \
    if (config_matches<C>(*d)) return solve_cfg<C>(*d, P, xinit, x0, params, xtraj, utraj, info, iterates, nullptr);
    MPCGPU_CONFIG_LIST(X)
#undef X
    return -2;
}

// as emu_solve with the persistent multipliers of P solve slots (P x mpcgpu_state_size doubles, in/out)
int emu_solve_stateful(const mpcgpu_problem_desc *d, int P, const double *xinit, const double *x0, const double *params,
                       double *xtraj, double *utraj, mpcgpu_info *info, double *state) {
#define X(C) \
    if (config_matches<C>(*d)) return solve_cfg<C>(*d, P, xinit, x0, params, xtraj, utraj, info, nullptr, state);
    MPCGPU_CONFIG_LIST(X)
#undef X
    return -2;
}

int emu_eval_stage(const mpcgpu_problem_desc *d, int M, const double *z, const double *p, const double *lam,
                   const double *pi, double *cost, double *cost_grad, double *cost_hess, double *hval, double *h_jac,
                   double *h_hess, double *xnext, double *AB, double *dyn_hess) {
    const Layout L = make_layout(*d);
    const int nz = d->nx + d->nu, nx = d->nx, nh = d->nh, npar = d->npar;
#define X(C)                                                                                                                \
    {                                                                                                                       \
        if (config_matches<C>(*d)) {                                                                                        \
            for (int m = 0; m < M; m++)                                                                                     \
                eval_stage_sample<C>(L, z + (size_t)m * nz, p + (size_t)m * npar, lam ? lam + (size_t)m * nh : nullptr,     \
                                     pi ? pi + (size_t)m * nx : nullptr, cost ? cost + m : nullptr,                         \
                                     cost_grad ? cost_grad + (size_t)m * nz : nullptr,                                      \
                                     cost_hess ? cost_hess + (size_t)m * nz * nz : nullptr, hval ? hval + (size_t)m * nh : nullptr, \
                                     h_jac ? h_jac + (size_t)m * nh * nz : nullptr, h_hess ? h_hess + (size_t)m * nz * nz : nullptr, \
                                     xnext ? xnext + (size_t)m * nx : nullptr, AB ? AB + (size_t)m * nx * nz : nullptr,     \
                                     dyn_hess ? dyn_hess + (size_t)m * nz * nz : nullptr);                                  \
            return 0;                                                                                                       \
        }                                                                                                                   \
    }
    MPCGPU_CONFIG_LIST(X)
#undef X
    return -2;
}
}
