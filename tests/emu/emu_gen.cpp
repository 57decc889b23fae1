// emu_gen.cpp -- TEST INFRASTRUCTURE: host compile of a GENERATED stage-function header (mpc_planner_b200/codegen.py) next to
// the hand-derived functions of csrc/stage_functions.cuh, so that the CPU suite can compare the two (and the generated
// ones against golden values from the reference's Python).  Compile with -DMPCGPU_GENERATED_STAGE="\"path/stage_<name>.cuh\"" (the
// same switch the per-problem CUDA libraries are built with).
#include <cstring>

#include "../../mpc_planner_b200/csrc/stage_functions.cuh"  // includes the generated header (MPCGPU_GENERATED_STAGE)

using namespace mpcgpu;

namespace {
struct ParamPlain {
    const double *p;
    double operator()(int i) const { return p[i]; }
};
}  // namespace

extern "C" {
int gen_dims(int *nz, int *nh, int *npar) {
    *nz = MPCGPU_GEN_NZ;
    *nh = MPCGPU_GEN_NH;
    *npar = MPCGPU_GEN_NPAR;
    return 0;
}
// M samples: z[M*nz], p[M*npar], w[M*nh] (plain weights of the constraint Hessian, may be null) ->
// cost[M], cost_grad[M*nz], cost_hess[M*nz*nz], h[M*nh], h_jac3[M*nh*3] (d/d(x,y,psi)), h_hess6[M*6] (packed lower over x,y,psi)
int gen_eval(int M, const double *z, const double *p, const double *w, double *cost, double *cost_grad, double *cost_hess, double *h,
             double *h_jac3, double *h_hess6) {
    constexpr int NZ = MPCGPU_GEN_NZ, NH = MPCGPU_GEN_NH, NPAR = MPCGPU_GEN_NPAR, NHP = NZ * (NZ + 1) / 2;
    for (int m = 0; m < M; m++) {
        ParamPlain pa{p + (size_t)m * NPAR};
        const double *zm = z + (size_t)m * NZ;
        double g[NZ], Hp[NHP];
        for (int i = 0; i < NZ; i++) g[i] = 0.0;
        for (int i = 0; i < NHP; i++) Hp[i] = 0.0;
        const double c = gen::stage_cost(pa, zm, 1.0, g, Hp);
        const double c0 = gen::stage_cost(pa, zm, 1.0, (double *)nullptr, (double *)nullptr);
        if (!(fabs(c - c0) <= 1e-13 * (1.0 + fabs(c))) && !(c != c && c0 != c0)) return -9;  // the value-only path (its own CSE) agrees to rounding
        cost[m] = c;
        for (int i = 0; i < NZ; i++) {
            cost_grad[(size_t)m * NZ + i] = g[i];
            for (int j = 0; j < NZ; j++) cost_hess[((size_t)m * NZ + i) * NZ + j] = Hp[i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i];
        }
        double Hx[6] = {0, 0, 0, 0, 0, 0};
        gen::stage_constraints<false>(pa, zm, h + (size_t)m * NH, h_jac3 + (size_t)m * NH * 3, w ? w + (size_t)m * NH : nullptr, w ? Hx : nullptr, 0, NH);
        for (int i = 0; i < 6; i++) h_hess6[(size_t)m * 6 + i] = Hx[i];
    }
    return 0;
}
}
