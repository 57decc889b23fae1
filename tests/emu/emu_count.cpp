// emu_count.cpp -- TEST INFRASTRUCTURE (scripts/count_flops.py).  The host emulation of the kernel source compiled with
// `double` replaced by a counting wrapper: the FLOPs the kernel's ALGORITHM executes, per phase category, measured from the
// code instead of estimated.  The emulation runs the single-thread Riccati sweep (structural zeros of
// [B A] skipped) -- the tensor-core sweep of the device build does the same algebra as dense 8 x 8 x 8 products, so this is
// the algorithmic (lower) count.
#include <math.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mpcgpu.h"
#include "counting_double.h"

mpc_flop_counters g_flops = {0, 0, 0, 0, 0};
static long long g_phase[16][5];
static mpc_flop_counters g_last = {0, 0, 0, 0, 0};
static void count_tick(int cat) {
    g_phase[cat][0] += g_flops.add - g_last.add;
    g_phase[cat][1] += g_flops.mul - g_last.mul;
    g_phase[cat][2] += g_flops.div - g_last.div;
    g_phase[cat][3] += g_flops.sqrt_ - g_last.sqrt_;
    g_phase[cat][4] += g_flops.trans - g_last.trans;
    g_last = g_flops;
}
#define MPC_TICK(cat) count_tick(cat);
#define MPC_COUNT_FLOPS 1

#define double mpc_cdouble
namespace mpcgpu { extern long g_mirror_slow, g_mirror_calls; }
#include "../../mpc_planner_b200/csrc/registry.cuh"
namespace mpcgpu { long g_mirror_slow = 0, g_mirror_calls = 0, g_mirror_general = 0; long long g_emu_clock_ns = 0, g_emu_clock_step_ns = 0; }
#undef double
using namespace mpcgpu;

template <class CFG>
static int solve_cfg(const mpcgpu_problem_desc &d, int P, const double *xinit, const double *x0, const double *params, mpcgpu_info *info) {
    typedef SL<CFG> S;
    Layout L = make_layout(d);
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU;
    std::vector<mpc_cdouble> sm(S::TOTAL), pt((size_t)d.npar * pt_stride(N)), xi(NX), xw((size_t)NZ * (N + 1)), xt((size_t)NX * (N + 1)), ut((size_t)NU * N);
    for (int p = 0; p < P; p++) {
        const double *par = params + (size_t)p * d.npar * N;
        for (auto &v : pt) v = mpc_cdouble(0.0);
        for (int k = 0; k < N; k++)
            for (int i = 0; i < d.npar; i++) pt[(size_t)i * pt_stride(N) + k] = mpc_cdouble(par[(size_t)k * d.npar + i]);
        for (int i = 0; i < NX; i++) xi[i] = mpc_cdouble(xinit[(size_t)p * NX + i]);
        for (int i = 0; i < NZ * (N + 1); i++) xw[i] = mpc_cdouble(x0[(size_t)p * NZ * (N + 1) + i]);
        g_last = g_flops;
        sqp_solve_cta<CFG>(L, sm.data(), xi.data(), xw.data(), pt.data(), xt.data(), ut.data(), info + p, nullptr, nullptr);
        count_tick(TK_FINAL);
    }
    return 0;
}

extern "C" {
// phase[16][5]: add, mul, div, sqrt, transcendental per phase category (enum TK_* of csrc/sqp_core.cuh), summed over the P solves
int emu_count_solve(const mpcgpu_problem_desc *d, int P, const double *xinit, const double *x0, const double *params, mpcgpu_info *info,
                    long long *phase) {
    memset(g_phase, 0, sizeof(g_phase));
    int rc = -2;
#define X(C) \
    if (rc == -2 && config_matches<C>(*d)) rc = solve_cfg<C>(*d, P, xinit, x0, params, info);
    MPCGPU_CONFIG_LIST(X)
#undef X
    memcpy(phase, g_phase, sizeof(g_phase));
    return rc;
}
}
