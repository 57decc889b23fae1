// mpcgpu.cu -- CUDA kernels and the C ABI (include/mpcgpu.h) of the B200 batched SQP backend.
//
// Launch geometry: one 128-thread CTA per solve, the whole NLP/QP of that solve in shared memory
// (SL<CFG>::TOTAL doubles: 56 KB for the T-MPC stack => 4 CTAs = 16 warps per SM).
// Kernels per batch:  transpose_params (reference stage-major parameters -> lane-major so that
// lane = stage loads coalesce)  ->  sqp_kernel  ->  select_best (FindBestPlanner argmin).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "projection.cuh"
#include "registry.cuh"

using namespace mpcgpu;

namespace {

enum {
    MPCGPU_OK = 0,
    MPCGPU_ERR_ARG = -1,
    MPCGPU_ERR_UNSUPPORTED = -2,
    MPCGPU_ERR_CUDA = -3,
    MPCGPU_ERR_NO_DEVICE = -4,
    MPCGPU_ERR_BATCH = -5,
    MPCGPU_ERR_ALLOC = -6
};

thread_local char g_last_cuda_error[256] = "";

#define CUDA_TRY(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", #expr, cudaGetErrorString(e__)); \
            return MPCGPU_ERR_CUDA;                                                                           \
        }                                                                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// params[P][N][npar] (AcadosParameters::all_parameters, stage-major) -> pt[P][npar][pt_stride(N)]
__device__ __forceinline__ void stamp_call_start(long long *t_call) {
    if (t_call && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *t_call = (long long)t;
    }
}

__global__ void transpose_params_kernel(const double *__restrict__ params, double *__restrict__ pt, int P, int N, int npar, long long *t_call) {
    stamp_call_start(t_call);
    const long long total = (long long)P * N * npar;
    const int PT = pt_stride(N);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % npar);
        const long long pk = idx / npar;
        const int k = (int)(pk % N);
        const long long p = pk / N;
        pt[(p * npar + i) * PT + k] = params[idx];
    }
}

// Resident CTAs per SM the kernel is compiled for: what the shared-memory image allows (228 KB per SM, 1 KB
// reserved per CTA), at most 4 (= 128 registers per thread).  Measured on B200 for the T-MPC stack (56 KB):
// 2 / 3 / 4 CTAs per SM -> 197k / 245k / 271k solves/s.
// T-MPC++: per (control step s, guess g) build the solve's transposed parameter block from the step's main-solver
// parameters and, for guided guesses, the topology halfspaces from the guess's warm start
// (LinearizedConstraints::update / projectToSafety / setParameters, linearized_constraints.cpp:49-189).
__global__ void pack_tmpc_kernel(const double *__restrict__ scene_params, const double *__restrict__ xinit_scene,
                                 const double *__restrict__ x0, const int *__restrict__ guided, double *__restrict__ pt,
                                 double *__restrict__ xinit_out, int S, int G, int N, int npar, int nx, int nu, int lin0, int n_lin,
                                 int obs0, int obs_stride, int n_obs, double robot_radius, long long *t_call) {
    stamp_call_start(t_call);
    const int p = blockIdx.x, s = p / G, g = p - s * G;
    if (s >= S) return;
    const double *sp = scene_params + (size_t)s * N * npar;
    const int PT = pt_stride(N);
    double *ptp = pt + (size_t)p * npar * PT;
    for (int idx = threadIdx.x; idx < N * npar; idx += blockDim.x) {
        const int k = idx / npar, i = idx - k * npar;
        ptp[i * PT + k] = sp[idx];
    }
    if (threadIdx.x < nx) xinit_out[(size_t)p * nx + threadIdx.x] = xinit_scene[(size_t)s * nx + threadIdx.x];
    __syncthreads();
    const int nz = nx + nu;
    const double x_init = xinit_scene[(size_t)s * nx];
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const bool build = (guided[g] != 0) && (k >= 1);
        double px = 0.0, py = 0.0;
        const double R = 1e-3 + robot_radius;  // topology mode: obstacle radius 1e-3 (linearized_constraints.cpp:101)
        if (build) {
            px = x0[((size_t)p * (N + 1) + k) * nz + nu];
            py = x0[((size_t)p * (N + 1) + k) * nz + nu + 1];
            // projectToSafety (linearized_constraints.cpp:130-148): at most 3 Douglas-Rachford sweeps over the obstacles,
            // anchor = obstacle 0 of this stage
            if (n_obs > 0) {
                const double ax = sp[k * npar + obs0], ay = sp[k * npar + obs0 + 1];
                for (int sweep = 0; sweep < 3; sweep++)
                    for (int o = 0; o < n_obs; o++)
                        douglas_rachford_projection(px, py, sp[k * npar + obs0 + obs_stride * o], sp[k * npar + obs0 + obs_stride * o + 1], ax, ay, R);
            }
        }
        for (int j = 0; j < n_lin; j++) {
            double a1 = 1.0, a2 = 0.0, b = x_init + 100.0;  // dummies (linearized_constraints.cpp:53,155-166,181-187)
            if (build && j < n_obs) {
                const double ox = sp[k * npar + obs0 + obs_stride * j], oy = sp[k * npar + obs0 + obs_stride * j + 1];
                double dx = ox - px, dy = oy - py, dist = sqrt(dx * dx + dy * dy);
                if (dist < 1e-9) { dx = 1.0; dy = 0.0; dist = 1.0; }  // same guard as the projection: never a NaN normal
                a1 = dx / dist;
                a2 = dy / dist;
                b = a1 * ox + a2 * oy - R;
            } else if (build) {
                // rows behind the obstacles: the static / road halfspaces that LinearizedConstraints::update copies into the
                // `_n_other_halfspaces` rows (linearized_constraints.cpp:111-123) -- taken from the control step's own
                // parameter block (callers without static halfspaces leave the dummies there)
                a1 = sp[k * npar + lin0 + 3 * j];
                a2 = sp[k * npar + lin0 + 3 * j + 1];
                b = sp[k * npar + lin0 + 3 * j + 2];
            }
            ptp[(lin0 + 3 * j) * PT + k] = a1;
            ptp[(lin0 + 3 * j + 1) * PT + k] = a2;
            ptp[(lin0 + 3 * j + 2) * PT + k] = b;
        }
    }
}

// T-MPC++ from a COMPACT scene description: per (control step s, guess g) the whole transposed parameter block is built on
// the device -- what MPCBaseModule / Contouring / EllipsoidConstraints / GaussianConstraints / DecompConstraints ::
// setParameters and LinearizedConstraints::update + setParameters write stage by stage on the host in the reference
// (mpc_base.cpp:23-35, contouring.cpp:94-124, ellipsoid_constraints.cpp:34-90, gaussian_constraints.cpp:31-79,
// decomp_constraints.cpp:90-114,150-189, linearized_constraints.cpp:49-189).  See mpcgpu_solve_tmpc_scene_host.
struct SceneBlocks {
    int n_mod, kind[3], p0[3], count[3];
};
__global__ void pack_scene_kernel(const double *__restrict__ base, const double *__restrict__ obstacles, const double *__restrict__ radius,
                                  const int *__restrict__ flags, const double *__restrict__ halfspaces, int n_hs, double risk,
                                  const double *__restrict__ xinit_scene, const double *__restrict__ x0, const int *__restrict__ guided,
                                  double *__restrict__ pt, double *__restrict__ xinit_out, int S, int G, int N, int npar, int nx, int nu,
                                  SceneBlocks B, int n_obs, double robot_radius, long long *t_call) {
    stamp_call_start(t_call);
    const int p = blockIdx.x, s = p / G, g = p - s * G;
    if (s >= S) return;
    const int PT = pt_stride(N), nz = nx + nu;
    double *ptp = pt + (size_t)p * npar * PT;
    const double *bs = base + (size_t)s * npar;
    for (int idx = threadIdx.x; idx < N * npar; idx += blockDim.x) {
        const int i = idx / N, k = idx - i * N;
        ptp[i * PT + k] = bs[i];  // stage-invariant part (weights, spline, ego disc): the same at every stage
    }
    if (threadIdx.x < nx) xinit_out[(size_t)p * nx + threadIdx.x] = xinit_scene[(size_t)s * nx + threadIdx.x];
    __syncthreads();
    const double x_init = xinit_scene[(size_t)s * nx], y_init = xinit_scene[(size_t)s * nx + 1];
    const double *ob = obstacles + (size_t)s * n_obs * N * 5;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        for (int m = 0; m < B.n_mod; m++) {
            const int kind = B.kind[m], p0 = B.p0[m], cnt = B.count[m];
            if (kind == MPCGPU_CON_ELLIPSOID) {
                for (int o = 0; o < cnt; o++) {
                    double v[7] = {x_init + 50.0, y_init + 50.0, 0.0, 0.0, 0.0, 1.0, 0.1};  // stage 0 dummies (:42-55)
                    if (k >= 1 && o < n_obs) {
                        const double *st = ob + ((size_t)o * N + (k - 1)) * 5;  // prediction step k - 1 feeds stage k (:67-70)
                        const bool gauss = (flags[(size_t)s * n_obs + o] & 1) != 0;
                        v[0] = st[0]; v[1] = st[1]; v[2] = st[2];
                        v[3] = gauss ? st[3] : 0.0;
                        v[4] = gauss ? st[4] : 0.0;
                        v[5] = gauss ? -log(risk) / 0.5 : 1.0;  // RosTools::ExponentialQuantile(0.5, 1 - risk) (:80)
                        v[6] = radius[(size_t)s * n_obs + o];
                    }
                    for (int c = 0; c < 7; c++) ptp[(p0 + 7 * o + c) * PT + k] = v[c];
                }
            } else if (kind == MPCGPU_CON_GAUSSIAN) {
                for (int o = 0; o < cnt; o++) {
                    double v[6] = {x_init + 100.0, y_init + 100.0, 0.1, 0.1, 0.05, 0.1};  // (:39-52)
                    if (k >= 1 && o < n_obs && (flags[(size_t)s * n_obs + o] & 1)) {
                        const double *st = ob + ((size_t)o * N + (k - 1)) * 5;
                        const bool is_static = (flags[(size_t)s * n_obs + o] & 2) != 0;  // static obstacles: no uncertainty (:72-73)
                        v[0] = st[0]; v[1] = st[1];
                        v[2] = is_static ? 0.001 : st[3];
                        v[3] = is_static ? 0.001 : st[4];
                        v[4] = risk;
                        v[5] = radius[(size_t)s * n_obs + o];
                    }
                    for (int c = 0; c < 6; c++) ptp[(p0 + 6 * o + c) * PT + k] = v[c];
                }
            } else if (kind == MPCGPU_CON_LINEAR) {
                const bool build = (guided[g] != 0) && (k >= 1);
                double px = 0.0, py = 0.0;
                const double R = 1e-3 + robot_radius;  // topology mode (linearized_constraints.cpp:101)
                if (build) {
                    px = x0[((size_t)p * (N + 1) + k) * nz + nu];
                    py = x0[((size_t)p * (N + 1) + k) * nz + nu + 1];
                    if (n_obs > 0) {
                        const double *a0 = ob + (size_t)(k - 1) * 5;
                        for (int sweep = 0; sweep < 3; sweep++)
                            for (int o = 0; o < n_obs; o++) {
                                const double *st = ob + ((size_t)o * N + (k - 1)) * 5;
                                douglas_rachford_projection(px, py, st[0], st[1], a0[0], a0[1], R);
                            }
                    }
                }
                for (int j = 0; j < cnt; j++) {
                    double a1 = 1.0, a2 = 0.0, b = x_init + 100.0;
                    if (build && j < n_obs) {
                        const double *st = ob + ((size_t)j * N + (k - 1)) * 5;
                        double dx = st[0] - px, dy = st[1] - py, dist = sqrt(dx * dx + dy * dy);
                        if (dist < 1e-9) { dx = 1.0; dy = 0.0; dist = 1.0; }
                        a1 = dx / dist;
                        a2 = dy / dist;
                        b = a1 * st[0] + a2 * st[1] - R;
                    }
                    ptp[(p0 + 3 * j) * PT + k] = a1;
                    ptp[(p0 + 3 * j + 1) * PT + k] = a2;
                    ptp[(p0 + 3 * j + 2) * PT + k] = b;
                }
            } else {  // DECOMP: rows from outside up to the first zero / NaN normal, dummies behind (decomp_constraints.cpp:90-114)
                bool valid = (k >= 1) && halfspaces != nullptr;
                for (int j = 0; j < cnt; j++) {
                    double a1 = 1.0, a2 = 0.0, b = x_init + 100.0;
                    if (valid && j < n_hs) {
                        const double *hs = halfspaces + (((size_t)s * N + k) * n_hs + j) * 3;
                        if (sqrt(hs[0] * hs[0] + hs[1] * hs[1]) < 1e-3 || hs[0] != hs[0]) valid = false;
                        else { a1 = hs[0]; a2 = hs[1]; b = hs[2]; }
                    } else {
                        valid = false;
                    }
                    ptp[(p0 + 3 * j) * PT + k] = a1;
                    ptp[(p0 + 3 * j + 1) * PT + k] = a2;
                    ptp[(p0 + 3 * j + 2) * PT + k] = b;
                }
            }
        }
    }
}

template <class CFG>
constexpr int sqp_min_ctas() {
#if defined(MPC_MIN_CTAS)  // development builds (scripts/residency_sweep.py): fixed register budget
    return MPC_MIN_CTAS;
#endif
    return (233472 / (int)(SL<CFG>::TOTAL * sizeof(double) + 1024)) >= 4   ? 4
           : (233472 / (int)(SL<CFG>::TOTAL * sizeof(double) + 1024)) >= 1 ? (233472 / (int)(SL<CFG>::TOTAL * sizeof(double) + 1024))
                                                                            : 1;
}
template <class CFG>
__global__ void __launch_bounds__(MPC_NT, sqp_min_ctas<CFG>()) sqp_kernel(Layout L, int P, const double *__restrict__ xinit, const double *__restrict__ x0,
                                                 const double *__restrict__ pt, double *__restrict__ xtraj,
                                                 double *__restrict__ utraj, mpcgpu_info *__restrict__ info,
                                                 unsigned int *__restrict__ sm_slots, double *__restrict__ state) {
    extern __shared__ double sm[];
    __shared__ int s_serial_tid;
    const int p = blockIdx.x;
    if (p >= P) return;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU;
    // The serial Riccati sweeps run in ONE thread per CTA.  Warp w of every CTA lives on SM sub-partition w % 4, so
    // co-resident CTAs would all serialise on sub-partition 0; a per-SM ticket spreads their serial warps round-robin.
    if (threadIdx.x == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_serial_tid = (int)((atomicAdd(&sm_slots[smid & 255], 1u) % (MPC_NT / 32)) * 32);
    }
    __syncthreads();
    sqp_solve_cta<CFG>(L, sm, xinit + (size_t)p * NX, x0 + (size_t)p * NZ * (N + 1), pt + (size_t)p * L.npar * pt_stride(N),
                        xtraj + (size_t)p * NX * (N + 1), utraj + (size_t)p * NU * N, info + p, nullptr,
                        state ? state + (size_t)p * ((N + 1) * NX + N * CFG::NH) : nullptr, s_serial_tid);
}

// FindBestPlanner (mpc_planner_modules/src/guidance_constraints.cpp:416-434): per group of `group`
// consecutive solves (the guesses of one control step), index of the successful solve with the
// smallest objective, -1 if none.  One warp per group; ties resolve to the lowest index as in the
// reference's serial loop.
__global__ void select_best_kernel(const mpcgpu_info *__restrict__ info, int P, int group, int n_groups, int *__restrict__ best) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_groups) return;
    const int base = warp * group, end = min(base + group, P);
    double bo = 1e10;  // the reference's sentinel (guidance_constraints.cpp:419)
    int bi = 0x7fffffff;
    for (int i = base + lane; i < end; i += 32) {
        const double o = info[i].pobj;
        if (info[i].exit_code == 1 && (o < bo)) {
            bo = o;
            bi = i;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, bo, off);
        const int j = __shfl_xor_sync(0xffffffffu, bi, off);
        if (o < bo || (o == bo && j < bi)) {
            bo = o;
            bi = j;
        }
    }
    if (lane == 0) best[warp] = (bi == 0x7fffffff) ? -1 : bi - base;
}

template <class CFG>
__global__ void eval_stage_kernel(Layout L, int M, const double *z, const double *p, const double *lam, const double *pi,
                                  double *cost, double *cost_grad, double *cost_hess, double *hval, double *h_jac,
                                  double *h_hess, double *xnext, double *AB, double *dyn_hess) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    constexpr int nz = CFG::NZ, nx = CFG::NX, nh = CFG::NH;
    eval_stage_sample<CFG>(L, z + (size_t)m * nz, p + (size_t)m * L.npar, lam ? lam + (size_t)m * nh : nullptr,
                           pi ? pi + (size_t)m * nx : nullptr, cost ? cost + m : nullptr,
                           cost_grad ? cost_grad + (size_t)m * nz : nullptr, cost_hess ? cost_hess + (size_t)m * nz * nz : nullptr,
                           hval ? hval + (size_t)m * nh : nullptr, h_jac ? h_jac + (size_t)m * nh * nz : nullptr,
                           h_hess ? h_hess + (size_t)m * nz * nz : nullptr, xnext ? xnext + (size_t)m * nx : nullptr,
                           AB ? AB + (size_t)m * nx * nz : nullptr, dyn_hess ? dyn_hess + (size_t)m * nz * nz : nullptr);
}

// dependent-free FP64 FMA chains: 8 independent accumulators per thread
__global__ void fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------------------------
// solver object
// ---------------------------------------------------------------------------------------------
struct KernelTable {
    const char *name;
    bool (*matches)(const mpcgpu_problem_desc &);
    int smem_bytes;
    cudaError_t (*prepare)(int smem_bytes);
    void (*launch_sqp)(const Layout &, int P, const double *, const double *, const double *, double *, double *, mpcgpu_info *,
                       unsigned int *, double *state, int smem, cudaStream_t);
    void (*launch_eval)(const Layout &, int M, const double *, const double *, const double *, const double *, double *, double *,
                        double *, double *, double *, double *, double *, double *, double *, cudaStream_t);
};

template <class CFG>
cudaError_t prepare_cfg(int smem_bytes) {
    return cudaFuncSetAttribute(sqp_kernel<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}
template <class CFG>
void launch_sqp_cfg(const Layout &L, int P, const double *xinit, const double *x0, const double *pt, double *xtraj, double *utraj,
                    mpcgpu_info *info, unsigned int *sm_slots, double *state, int smem, cudaStream_t s) {
#if defined(MPC_DEV_KNOBS)  // development builds: MPCGPU_SMEM_PAD bytes of unused dynamic shared memory lower the residency
    if (const char *e = getenv("MPCGPU_SMEM_PAD")) {
        smem += atoi(e);
        cudaFuncSetAttribute(sqp_kernel<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
#endif
    sqp_kernel<CFG><<<P, MPC_NT, smem, s>>>(L, P, xinit, x0, pt, xtraj, utraj, info, sm_slots, state);
}
template <class CFG>
void launch_eval_cfg(const Layout &L, int M, const double *z, const double *p, const double *lam, const double *pi, double *cost,
                     double *cost_grad, double *cost_hess, double *hval, double *h_jac, double *h_hess, double *xnext, double *AB,
                     double *dyn_hess, cudaStream_t s) {
    eval_stage_kernel<CFG><<<(M + 63) / 64, 64, 0, s>>>(L, M, z, p, lam, pi, cost, cost_grad, cost_hess, hval, h_jac, h_hess, xnext,
                                                       AB, dyn_hess);
}

#define X(C) {#C, &config_matches<C>, (int)(SL<C>::TOTAL * sizeof(double)), &prepare_cfg<C>, &launch_sqp_cfg<C>, &launch_eval_cfg<C>},
const KernelTable g_table[] = {MPCGPU_CONFIG_LIST(X)};
#undef X
const int g_table_size = sizeof(g_table) / sizeof(g_table[0]);

const KernelTable *find_kernel(const mpcgpu_problem_desc &d) {
    for (int i = 0; i < g_table_size; i++)
        if (g_table[i].matches(d)) return &g_table[i];
    return nullptr;
}

}  // namespace

// Range checks of everything the kernels index with: a wrong descriptor must be refused at the ABI, not read out of bounds
// on the device (parameters are fetched as __ldg(base + index * stride + stage)).
static bool desc_is_valid(const mpcgpu_problem_desc &d) {
    const int nz = d.nx + d.nu;
    if (d.N < 1 || d.N + 1 > 32 || d.nx <= 0 || d.nu <= 0 || nz > MPCGPU_MAX_NZ || d.npar <= 0 || d.nh < 0) return false;
    if (!(d.dt > 0.0) || d.n_sqp < 1 || d.qp_iter_max < 0 || !(d.qp_tol > 0.0) || !(d.reg_epsilon > 0.0)) return false;
    if (!(d.qp_mu0 > 0.0) || !(d.qp_thr0 > 0.0) || !(d.qp_relax >= 0.0) || !(d.res_eq_fail > 0.0)) return false;
    if (d.qp_warm_start != 0 && (!(d.qp_ws_thr > 0.0) || !(d.qp_ws_mu > 0.0))) return false;
    auto in_range = [&](int idx, int extent) { return idx >= 0 && extent >= 1 && idx + extent <= d.npar; };
    auto optional = [&](int idx) { return idx == -1 || in_range(idx, 1); };
    const bool gen_obj = (d.objective == MPCGPU_OBJ_GENERATED);  // emitted code: its parameter indices are compiled in
    if (!gen_obj) {
    if (!in_range(d.p_acceleration, 1) || !in_range(d.p_angular_velocity, 1)) return false;
    if (!optional(d.p_slack) || !optional(d.p_ego_disc_offset)) return false;
    const bool needs_vref = d.weigh_velocity != 0 || d.objective == MPCGPU_OBJ_CURVATURE_AWARE;
    if (needs_vref ? (!in_range(d.p_velocity, 1) || !in_range(d.p_reference_velocity, 1))
                   : (!optional(d.p_velocity) || !optional(d.p_reference_velocity))) return false;
    if (!in_range(d.p_contour, 1)) return false;
    if (d.objective == MPCGPU_OBJ_CONTOURING ? !in_range(d.p_lag, 1) : !optional(d.p_lag)) return false;
    if (d.n_segments < 1 || !in_range(d.p_spline, 9 * d.n_segments)) return false;
    }
    if (d.n_con_modules < 0 || d.n_con_modules > 3) return false;
    int rows = 0;
    bool needs_disc = false;
    for (int m = 0; m < d.n_con_modules; m++) {
        int stride = 3;
        switch (d.con_kind[m]) {
        case MPCGPU_CON_LINEAR: break;
        case MPCGPU_CON_ELLIPSOID: stride = 7; needs_disc = true; break;
        case MPCGPU_CON_GAUSSIAN: stride = 6; needs_disc = true; break;
        case MPCGPU_CON_DECOMP:
        case MPCGPU_CON_SCENARIO: break;
        case MPCGPU_CON_GENERATED: stride = 0; break;
        default: return false;
        }
        if (d.con_count[m] < 1) return false;
        if (stride > 0 && !in_range(d.con_param[m], stride * d.con_count[m])) return false;
        rows += d.con_count[m];
    }
    if (rows != d.nh) return false;
    if (needs_disc && !in_range(d.p_ego_disc_radius, 1)) return false;
    for (int i = 0; i < nz; i++) {
        if (d.lower_bound[i] != d.lower_bound[i] || d.upper_bound[i] != d.upper_bound[i]) return false;
        if (!(d.lower_bound[i] <= d.upper_bound[i])) return false;
    }
    return true;
}

struct mpcgpu_solver {
    mpcgpu_problem_desc desc;
    Layout layout;
    const KernelTable *kt;
    int device, max_batch;
    cudaStream_t stream;  // owned streams for the host path: compute, H2D, D2H
    cudaStream_t stream_in, stream_out, stream2;  // stream2: second compute stream (chunk kernels overlap their tails)
    cudaEvent_t ev_in[8], ev_cmp[8];
    // device buffers (host path)
    double *d_xinit, *d_x0, *d_params, *d_pt, *d_xtraj, *d_utraj;
    mpcgpu_info *d_info;
    int *d_best;
    unsigned int *d_sm_slots;  // per-SM tickets (see sqp_kernel)
    double *d_state;           // persistent NLP multipliers per solve slot (mpcgpu_set_stateful), else null
    long long *d_t_call;       // per chunk: %globaltimer at the first kernel of the call (solver timeout)
    size_t n_state;
    int *d_guided;             // T-MPC guided flags (lazy)
    double *d_xinit_scene;     // T-MPC per-control-step initial states (lazy)
    double *d_scene;           // compact scene tensors of the scene path (lazy): base | obstacles | radius | halfspaces
    int *d_scene_flags;
    size_t scene_doubles;
    long long launches;
    int group;  // solves per selection group (0 = the whole batch)
    int timing;
    cudaEvent_t ev0, ev1;
    double ms_acc;
    int ms_n;
    // sizes per solve
    size_t n_xinit, n_x0, n_params, n_xtraj, n_utraj;
};

extern "C" {

const char *mpcgpu_version(void) { return "mpcgpu 0.1 (sm_100a)"; }

const char *mpcgpu_strerror(int code) {
    switch (code) {
    case MPCGPU_OK: return "ok";
    case MPCGPU_ERR_ARG: return "invalid argument";
    case MPCGPU_ERR_UNSUPPORTED: return "no kernel specialisation compiled for this problem (see csrc/registry.cuh)";
    case MPCGPU_ERR_CUDA: return g_last_cuda_error[0] ? g_last_cuda_error : "CUDA error";
    case MPCGPU_ERR_NO_DEVICE: return "no CUDA device available: the CUDA backend has no CPU fallback";
    case MPCGPU_ERR_BATCH: return "batch larger than max_batch of the handle";
    case MPCGPU_ERR_ALLOC: return "allocation failed";
    }
    return "unknown error";
}

int mpcgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int mpcgpu_problem_supported(const mpcgpu_problem_desc *desc) { return (desc && find_kernel(*desc)) ? 1 : 0; }

int mpcgpu_create(const mpcgpu_problem_desc *desc, int device, int max_batch, mpcgpu_handle *out) {
    if (!desc || !out || max_batch <= 0) return MPCGPU_ERR_ARG;
    *out = nullptr;
    const KernelTable *kt = find_kernel(*desc);
    if (!kt) return MPCGPU_ERR_UNSUPPORTED;
    if (!desc_is_valid(*desc)) return MPCGPU_ERR_ARG;
    if (mpcgpu_device_count() <= device || device < 0) return MPCGPU_ERR_NO_DEVICE;
    CUDA_TRY(cudaSetDevice(device));
    mpcgpu_solver *h = new (std::nothrow) mpcgpu_solver();
    if (!h) return MPCGPU_ERR_ALLOC;
    memset(h, 0, sizeof(*h));
    h->desc = *desc;
    h->layout = make_layout(*desc);
    h->kt = kt;
    h->device = device;
    h->max_batch = max_batch;
    const int N = desc->N, nx = desc->nx, nu = desc->nu, nz = nx + nu;
    h->n_xinit = nx;
    h->n_x0 = (size_t)nz * (N + 1);
    h->n_params = (size_t)desc->npar * N;
    h->n_xtraj = (size_t)nx * (N + 1);
    h->n_utraj = (size_t)nu * N;
    h->n_state = (size_t)(N + 1) * nx + (size_t)N * desc->nh;
    cudaError_t e = kt->prepare(kt->smem_bytes);
    if (e != cudaSuccess) {
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cudaFuncSetAttribute(smem=%d): %s", kt->smem_bytes, cudaGetErrorString(e));
        delete h;
        return MPCGPU_ERR_CUDA;
    }
    const size_t B = (size_t)max_batch;
    bool ok = true;
    ok = ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->stream_in, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->stream_out, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 8 && ok; i++)
        ok = cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&h->ev_cmp[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_xinit, B * h->n_xinit * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_x0, B * h->n_x0 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_params, B * h->n_params * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_pt, B * (size_t)desc->npar * pt_stride(desc->N) * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_xtraj, B * h->n_xtraj * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_utraj, B * h->n_utraj * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_info, B * sizeof(mpcgpu_info)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_best, B * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_sm_slots, 256 * sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_t_call, 8 * sizeof(long long)) == cudaSuccess;
    if (ok) ok = cudaMemset(h->d_sm_slots, 0, 256 * sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
    if (ok) ok = cudaMemset(h->d_pt, 0, B * (size_t)desc->npar * pt_stride(desc->N) * sizeof(double)) == cudaSuccess;
    if (!ok) {
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        mpcgpu_destroy(h);
        return MPCGPU_ERR_CUDA;
    }
    *out = h;
    return MPCGPU_OK;
}

int mpcgpu_describe(mpcgpu_handle h, mpcgpu_problem_desc *desc, int *device, int *max_batch) {
    if (!h) return MPCGPU_ERR_ARG;
    if (desc) *desc = h->desc;
    if (device) *device = h->device;
    if (max_batch) *max_batch = h->max_batch;
    return MPCGPU_OK;
}

int mpcgpu_destroy(mpcgpu_handle h) {
    if (!h) return MPCGPU_ERR_ARG;
    cudaSetDevice(h->device);
    cudaFree(h->d_xinit);
    cudaFree(h->d_x0);
    cudaFree(h->d_params);
    cudaFree(h->d_pt);
    cudaFree(h->d_xtraj);
    cudaFree(h->d_utraj);
    cudaFree(h->d_info);
    cudaFree(h->d_best);
    cudaFree(h->d_sm_slots);
    cudaFree(h->d_t_call);
    cudaFree(h->d_state);
    cudaFree(h->d_guided);
    cudaFree(h->d_xinit_scene);
    cudaFree(h->d_scene);
    cudaFree(h->d_scene_flags);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream_in) cudaStreamDestroy(h->stream_in);
    if (h->stream_out) cudaStreamDestroy(h->stream_out);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    for (int i = 0; i < 8; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_cmp[i]) cudaEventDestroy(h->ev_cmp[i]);
    }
    delete h;
    return MPCGPU_OK;
}

// enqueue transpose -> sqp -> select_best for `count` solves whose scratch slot starts at `slot`
static int enqueue_solve(mpcgpu_handle h, int slot, int count, const double *d_xinit, const double *d_x0, const double *d_params,
                         double *d_xtraj, double *d_utraj, mpcgpu_info *d_info, int *d_best, cudaStream_t s, bool timed, int chunk = 0) {
    const int N = h->desc.N, npar = h->desc.npar;
    double *pt = h->d_pt + (size_t)slot * npar * pt_stride(N);
    const long long total = (long long)count * N * npar;
    int tb = (int)((total + 255) / 256);
    if (tb > 148 * 16) tb = 148 * 16;
    long long *t_call = (h->layout.timeout_s > 0.0) ? h->d_t_call + chunk : nullptr;
    transpose_params_kernel<<<tb, 256, 0, s>>>(d_params, pt, count, N, npar, t_call);
    if (timed) cudaEventRecord(h->ev0, s);
    Layout L = h->layout;
    L.t_call = t_call;
    h->kt->launch_sqp(L, count, d_xinit, d_x0, pt, d_xtraj, d_utraj, d_info, h->d_sm_slots,
                      h->d_state ? h->d_state + (size_t)slot * h->n_state : nullptr, h->kt->smem_bytes, s);
    if (timed) cudaEventRecord(h->ev1, s);
    h->launches += 2;
    if (d_best) {
        const int group = (h->group > 0) ? h->group : count, n_groups = (count + group - 1) / group;
        select_best_kernel<<<(n_groups + 3) / 4, 128, 0, s>>>(d_info, count, group, n_groups, d_best);
        h->launches += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return MPCGPU_OK;
}

int mpcgpu_solve_device(mpcgpu_handle h, int P, const double *d_xinit, const double *d_x0, const double *d_params,
                        double *d_xtraj, double *d_utraj, mpcgpu_info *d_info, int *d_best, void *stream) {
    if (!h || P <= 0 || !d_xinit || !d_x0 || !d_params || !d_xtraj || !d_utraj || !d_info) return MPCGPU_ERR_ARG;
    if (P > h->max_batch) return MPCGPU_ERR_BATCH;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = enqueue_solve(h, 0, P, d_xinit, d_x0, d_params, d_xtraj, d_utraj, d_info, d_best, s, h->timing != 0);
    if (rc != MPCGPU_OK) return rc;
    if (h->timing) {
        CUDA_TRY(cudaEventSynchronize(h->ev1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->ms_acc += ms;
        h->ms_n += 1;
    }
    return MPCGPU_OK;
}

static int solve_host_impl(mpcgpu_handle h, int P, const double *xinit, const double *x0, const double *params, double *state, double *xtraj,
                           double *utraj, mpcgpu_info *info, int *best) {
    if (!h || P <= 0 || !xinit || !x0 || !params || !xtraj || !utraj || !info) return MPCGPU_ERR_ARG;
    if (P > h->max_batch) return MPCGPU_ERR_BATCH;
    if (state && !h->d_state) return MPCGPU_ERR_ARG;  // mpcgpu_set_stateful(h, 1) first
    CUDA_TRY(cudaSetDevice(h->device));
    // Large batches are split into chunks pipelined over three streams (H2D | kernels | D2H) so that the PCIe
    // copies of chunk i+1 / i-1 overlap the solve of chunk i.  Chunks are whole selection groups.
    const int group = (h->group > 0) ? h->group : P;
    int n_chunks = (P >= 2048 && h->group > 0) ? 4 : 1;
    int chunk = (P + n_chunks - 1) / n_chunks;
    chunk = ((chunk + group - 1) / group) * group;
    n_chunks = (P + chunk - 1) / chunk;
    for (int c = 0; c < n_chunks; c++) {
        const size_t o = (size_t)c * chunk, n = (size_t)((o + chunk <= (size_t)P) ? chunk : P - o);
        CUDA_TRY(cudaMemcpyAsync(h->d_xinit + o * h->n_xinit, xinit + o * h->n_xinit, n * h->n_xinit * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(h->d_x0 + o * h->n_x0, x0 + o * h->n_x0, n * h->n_x0 * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(h->d_params + o * h->n_params, params + o * h->n_params, n * h->n_params * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        if (state) CUDA_TRY(cudaMemcpyAsync(h->d_state + o * h->n_state, state + o * h->n_state, n * h->n_state * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaEventRecord(h->ev_in[c], h->stream_in));
    }
    for (int c = 0; c < n_chunks; c++) {
        const size_t o = (size_t)c * chunk, n = (size_t)((o + chunk <= (size_t)P) ? chunk : P - o);
        cudaStream_t sc = (c & 1) ? h->stream2 : h->stream;  // alternate compute streams: the partial last wave of one
                                                              // chunk overlaps the first wave of the next
        CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_in[c], 0));
        int rc = enqueue_solve(h, (int)o, (int)n, h->d_xinit + o * h->n_xinit, h->d_x0 + o * h->n_x0, h->d_params + o * h->n_params,
                               h->d_xtraj + o * h->n_xtraj, h->d_utraj + o * h->n_utraj, h->d_info + o, h->d_best + o / group, sc, h->timing != 0 && n_chunks == 1, c);
        if (rc != MPCGPU_OK) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_cmp[c], sc));
        CUDA_TRY(cudaStreamWaitEvent(h->stream_out, h->ev_cmp[c], 0));
        CUDA_TRY(cudaMemcpyAsync(xtraj + o * h->n_xtraj, h->d_xtraj + o * h->n_xtraj, n * h->n_xtraj * sizeof(double), cudaMemcpyDeviceToHost, h->stream_out));
        CUDA_TRY(cudaMemcpyAsync(utraj + o * h->n_utraj, h->d_utraj + o * h->n_utraj, n * h->n_utraj * sizeof(double), cudaMemcpyDeviceToHost, h->stream_out));
        CUDA_TRY(cudaMemcpyAsync(info + o, h->d_info + o, n * sizeof(mpcgpu_info), cudaMemcpyDeviceToHost, h->stream_out));
        if (state) CUDA_TRY(cudaMemcpyAsync(state + o * h->n_state, h->d_state + o * h->n_state, n * h->n_state * sizeof(double), cudaMemcpyDeviceToHost, h->stream_out));
        if (best) {
            const size_t ng = (n + group - 1) / group;
            CUDA_TRY(cudaMemcpyAsync(best + o / group, h->d_best + o / group, ng * sizeof(int), cudaMemcpyDeviceToHost, h->stream_out));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream_out));
    if (h->timing && n_chunks == 1) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->ms_acc += ms;
        h->ms_n += 1;
    }
    return MPCGPU_OK;
}

int mpcgpu_solve_host(mpcgpu_handle h, int P, const double *xinit, const double *x0, const double *params, double *xtraj,
                      double *utraj, mpcgpu_info *info, int *best) {
    return solve_host_impl(h, P, xinit, x0, params, nullptr, xtraj, utraj, info, best);
}

int mpcgpu_solve_host_stateful(mpcgpu_handle h, int P, const double *xinit, const double *x0, const double *params, double *state,
                               double *xtraj, double *utraj, mpcgpu_info *info, int *best) {
    if (!state) return MPCGPU_ERR_ARG;
    return solve_host_impl(h, P, xinit, x0, params, state, xtraj, utraj, info, best);
}

int mpcgpu_host_alloc(void **ptr, unsigned long long bytes) {
    if (!ptr) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return MPCGPU_OK;
}
int mpcgpu_host_free(void *ptr) {
    CUDA_TRY(cudaFreeHost(ptr));
    return MPCGPU_OK;
}

int mpcgpu_solve_tmpc_host(mpcgpu_handle h, int S, int G, const double *xinit, const double *x0, const double *scene_params,
                           const int *guided, double robot_radius, double *xtraj, double *utraj, mpcgpu_info *info, int *best) {
    if (!h || S <= 0 || G <= 0 || !xinit || !x0 || !scene_params || !guided || !xtraj || !utraj || !info) return MPCGPU_ERR_ARG;
    const int P = S * G;
    if (P > h->max_batch) return MPCGPU_ERR_BATCH;
    const mpcgpu_problem_desc &d = h->desc;
    int m_lin = -1, m_obs = -1;
    for (int m = 0; m < d.n_con_modules; m++) {
        if (d.con_kind[m] == MPCGPU_CON_LINEAR && m_lin < 0) m_lin = m;
        if ((d.con_kind[m] == MPCGPU_CON_ELLIPSOID || d.con_kind[m] == MPCGPU_CON_GAUSSIAN) && m_obs < 0) m_obs = m;
    }
    if (m_lin < 0 || m_obs < 0) return MPCGPU_ERR_UNSUPPORTED;
    const int obs_stride = (d.con_kind[m_obs] == MPCGPU_CON_ELLIPSOID) ? 7 : 6;
    const int n_obs = d.con_count[m_lin] < d.con_count[m_obs] ? d.con_count[m_lin] : d.con_count[m_obs];
    if (G > 1024) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->d_guided) {
        CUDA_TRY(cudaMalloc(&h->d_guided, 1024 * sizeof(int)));
        CUDA_TRY(cudaMalloc(&h->d_xinit_scene, (size_t)h->max_batch * h->n_xinit * sizeof(double)));
    }
    // Chunks of whole control steps, pipelined like mpcgpu_solve_host (H2D | two compute streams | D2H).
    int n_chunks = (P >= 2048) ? 4 : 1;
    int sc_chunk = (S + n_chunks - 1) / n_chunks;
    n_chunks = (S + sc_chunk - 1) / sc_chunk;
    CUDA_TRY(cudaMemcpyAsync(h->d_guided, guided, (size_t)G * sizeof(int), cudaMemcpyHostToDevice, h->stream_in));
    for (int c = 0; c < n_chunks; c++) {
        const size_t s0 = (size_t)c * sc_chunk, ns = (s0 + sc_chunk <= (size_t)S) ? sc_chunk : S - s0, p0 = s0 * G, np = ns * G;
        // the control steps' parameter blocks are staged in d_params (S blocks <= P blocks)
        CUDA_TRY(cudaMemcpyAsync(h->d_params + s0 * h->n_params, scene_params + s0 * h->n_params, ns * h->n_params * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(h->d_x0 + p0 * h->n_x0, x0 + p0 * h->n_x0, np * h->n_x0 * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(h->d_xinit_scene + s0 * h->n_xinit, xinit + s0 * h->n_xinit, ns * h->n_xinit * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaEventRecord(h->ev_in[c], h->stream_in));
    }
    for (int c = 0; c < n_chunks; c++) {
        const size_t s0 = (size_t)c * sc_chunk, ns = (s0 + sc_chunk <= (size_t)S) ? sc_chunk : S - s0, p0 = s0 * G, np = ns * G;
        cudaStream_t sc = (c & 1) ? h->stream2 : h->stream;
        CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_in[c], 0));
        pack_tmpc_kernel<<<(unsigned)np, 128, 0, sc>>>(h->d_params + s0 * h->n_params, h->d_xinit_scene + s0 * h->n_xinit, h->d_x0 + p0 * h->n_x0, h->d_guided,
                                                       h->d_pt + p0 * (size_t)d.npar * pt_stride(d.N), h->d_xinit + p0 * h->n_xinit, (int)ns, G, d.N, d.npar, d.nx, d.nu,
                                                       d.con_param[m_lin], d.con_count[m_lin], d.con_param[m_obs], obs_stride, n_obs, robot_radius,
                                                       (h->layout.timeout_s > 0.0) ? h->d_t_call + c : nullptr);
        const bool timed = h->timing != 0 && n_chunks == 1;
        if (timed) cudaEventRecord(h->ev0, sc);
        Layout L = h->layout;
        L.t_call = (h->layout.timeout_s > 0.0) ? h->d_t_call + c : nullptr;
        h->kt->launch_sqp(L, (int)np, h->d_xinit + p0 * h->n_xinit, h->d_x0 + p0 * h->n_x0, h->d_pt + p0 * (size_t)d.npar * pt_stride(d.N),
                          h->d_xtraj + p0 * h->n_xtraj, h->d_utraj + p0 * h->n_utraj, h->d_info + p0, h->d_sm_slots,
                          h->d_state ? h->d_state + p0 * h->n_state : nullptr, h->kt->smem_bytes, sc);
        if (timed) cudaEventRecord(h->ev1, sc);
        select_best_kernel<<<(unsigned)((ns + 3) / 4), 128, 0, sc>>>(h->d_info + p0, (int)np, G, (int)ns, h->d_best + s0);
        h->launches += 3;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(h->ev_cmp[c], sc));
        CUDA_TRY(cudaStreamWaitEvent(h->stream_out, h->ev_cmp[c], 0));
        CUDA_TRY(cudaMemcpyAsync(xtraj + p0 * h->n_xtraj, h->d_xtraj + p0 * h->n_xtraj, np * h->n_xtraj * sizeof(double), cudaMemcpyDeviceToHost, h->stream_out));
        CUDA_TRY(cudaMemcpyAsync(utraj + p0 * h->n_utraj, h->d_utraj + p0 * h->n_utraj, np * h->n_utraj * sizeof(double), cudaMemcpyDeviceToHost, h->stream_out));
        CUDA_TRY(cudaMemcpyAsync(info + p0, h->d_info + p0, np * sizeof(mpcgpu_info), cudaMemcpyDeviceToHost, h->stream_out));
        if (best) CUDA_TRY(cudaMemcpyAsync(best + s0, h->d_best + s0, ns * sizeof(int), cudaMemcpyDeviceToHost, h->stream_out));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream_out));
    if (h->timing && n_chunks == 1) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->ms_acc += ms;
        h->ms_n += 1;
    }
    return MPCGPU_OK;
}

int mpcgpu_solve_tmpc_scene_host(mpcgpu_handle h, int S, int G, const double *xinit, const double *x0, const double *base_params,
                                 int n_obs, const double *obstacles, const double *obstacle_radius, const int *obstacle_flags, double risk,
                                 int n_halfspaces, const double *halfspaces, const int *guided, double robot_radius, double *xtraj,
                                 double *utraj, mpcgpu_info *info, int *best) {
    if (!h || S <= 0 || G <= 0 || G > 1024 || !xinit || !x0 || !base_params || !guided || !xtraj || !utraj || !info || n_obs < 0 || n_halfspaces < 0) return MPCGPU_ERR_ARG;
    if (n_obs > 0 && (!obstacles || !obstacle_radius || !obstacle_flags)) return MPCGPU_ERR_ARG;
    if (!(risk > 0.0 && risk < 1.0)) return MPCGPU_ERR_ARG;
    const int P = S * G;
    if (P > h->max_batch) return MPCGPU_ERR_BATCH;
    const mpcgpu_problem_desc &d = h->desc;
    SceneBlocks B;
    B.n_mod = d.n_con_modules;
    for (int m = 0; m < 3; m++) {
        B.kind[m] = (m < d.n_con_modules) ? d.con_kind[m] : 0;
        B.p0[m] = (m < d.n_con_modules) ? d.con_param[m] : 0;
        B.count[m] = (m < d.n_con_modules) ? d.con_count[m] : 0;
        if (m < d.n_con_modules && d.con_kind[m] == MPCGPU_CON_SCENARIO) return MPCGPU_ERR_UNSUPPORTED;  // built by the external scenario_module
        if (m < d.n_con_modules && (d.con_kind[m] == MPCGPU_CON_ELLIPSOID || d.con_kind[m] == MPCGPU_CON_GAUSSIAN) && n_obs > d.con_count[m]) return MPCGPU_ERR_ARG;
    }
    const int N = d.N, npar = d.npar;
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->d_guided) {
        CUDA_TRY(cudaMalloc(&h->d_guided, 1024 * sizeof(int)));
        CUDA_TRY(cudaMalloc(&h->d_xinit_scene, (size_t)h->max_batch * h->n_xinit * sizeof(double)));
    }
    // staging of the compact scene: [base S*npar | obstacles S*n_obs*N*5 | radius S*n_obs | halfspaces S*N*n_hs*3]
    const size_t n_base = (size_t)S * npar, n_ob = (size_t)S * n_obs * N * 5, n_rad = (size_t)S * n_obs;
    const size_t n_hs = halfspaces ? (size_t)S * N * n_halfspaces * 3 : 0, need = n_base + n_ob + n_rad + n_hs;
    if (need > h->scene_doubles) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream2));
        cudaFree(h->d_scene);
        cudaFree(h->d_scene_flags);
        h->d_scene = nullptr;
        h->d_scene_flags = nullptr;
        h->scene_doubles = 0;
        CUDA_TRY(cudaMalloc(&h->d_scene, need * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->d_scene_flags, ((size_t)h->max_batch * 64 + 64) * sizeof(int)));
        h->scene_doubles = need;
    }
    if ((size_t)S * n_obs > (size_t)h->max_batch * 64 + 64) return MPCGPU_ERR_ARG;
    double *d_base = h->d_scene, *d_ob = d_base + n_base, *d_rad = d_ob + n_ob, *d_hs = d_rad + n_rad;
    int n_chunks = (P >= 2048) ? 4 : 1;
    int sc_chunk = (S + n_chunks - 1) / n_chunks;
    n_chunks = (S + sc_chunk - 1) / sc_chunk;
    CUDA_TRY(cudaMemcpyAsync(h->d_guided, guided, (size_t)G * sizeof(int), cudaMemcpyHostToDevice, h->stream_in));
    for (int c = 0; c < n_chunks; c++) {
        const size_t s0 = (size_t)c * sc_chunk, ns = (s0 + sc_chunk <= (size_t)S) ? sc_chunk : S - s0, p0 = s0 * G, np = ns * G;
        CUDA_TRY(cudaMemcpyAsync(d_base + s0 * npar, base_params + s0 * npar, ns * npar * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        if (n_obs > 0) {
            CUDA_TRY(cudaMemcpyAsync(d_ob + s0 * n_obs * N * 5, obstacles + s0 * n_obs * N * 5, ns * n_obs * N * 5 * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
            CUDA_TRY(cudaMemcpyAsync(d_rad + s0 * n_obs, obstacle_radius + s0 * n_obs, ns * n_obs * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
            CUDA_TRY(cudaMemcpyAsync(h->d_scene_flags + s0 * n_obs, obstacle_flags + s0 * n_obs, ns * n_obs * sizeof(int), cudaMemcpyHostToDevice, h->stream_in));
        }
        if (n_hs) CUDA_TRY(cudaMemcpyAsync(d_hs + s0 * N * n_halfspaces * 3, halfspaces + s0 * N * n_halfspaces * 3, ns * N * n_halfspaces * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(h->d_x0 + p0 * h->n_x0, x0 + p0 * h->n_x0, np * h->n_x0 * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(h->d_xinit_scene + s0 * h->n_xinit, xinit + s0 * h->n_xinit, ns * h->n_xinit * sizeof(double), cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaEventRecord(h->ev_in[c], h->stream_in));
    }
    for (int c = 0; c < n_chunks; c++) {
        const size_t s0 = (size_t)c * sc_chunk, ns = (s0 + sc_chunk <= (size_t)S) ? sc_chunk : S - s0, p0 = s0 * G, np = ns * G;
        cudaStream_t sc = (c & 1) ? h->stream2 : h->stream;
        CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_in[c], 0));
        long long *t_call = (h->layout.timeout_s > 0.0) ? h->d_t_call + c : nullptr;
        pack_scene_kernel<<<(unsigned)np, 128, 0, sc>>>(d_base + s0 * npar, d_ob + s0 * n_obs * N * 5, d_rad + s0 * n_obs, h->d_scene_flags + s0 * n_obs,
                                                        n_hs ? d_hs + s0 * N * n_halfspaces * 3 : nullptr, n_halfspaces, risk, h->d_xinit_scene + s0 * h->n_xinit,
                                                        h->d_x0 + p0 * h->n_x0, h->d_guided, h->d_pt + p0 * (size_t)npar * pt_stride(N), h->d_xinit + p0 * h->n_xinit,
                                                        (int)ns, G, N, npar, d.nx, d.nu, B, n_obs, robot_radius, t_call);
        const bool timed = h->timing != 0 && n_chunks == 1;
        if (timed) cudaEventRecord(h->ev0, sc);
        Layout L = h->layout;
        L.t_call = t_call;
        h->kt->launch_sqp(L, (int)np, h->d_xinit + p0 * h->n_xinit, h->d_x0 + p0 * h->n_x0, h->d_pt + p0 * (size_t)npar * pt_stride(N),
                          h->d_xtraj + p0 * h->n_xtraj, h->d_utraj + p0 * h->n_utraj, h->d_info + p0, h->d_sm_slots,
                          h->d_state ? h->d_state + p0 * h->n_state : nullptr, h->kt->smem_bytes, sc);
        if (timed) cudaEventRecord(h->ev1, sc);
        select_best_kernel<<<(unsigned)((ns + 3) / 4), 128, 0, sc>>>(h->d_info + p0, (int)np, G, (int)ns, h->d_best + s0);
        h->launches += 3;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(h->ev_cmp[c], sc));
        CUDA_TRY(cudaStreamWaitEvent(h->stream_out, h->ev_cmp[c], 0));
        CUDA_TRY(cudaMemcpyAsync(xtraj + p0 * h->n_xtraj, h->d_xtraj + p0 * h->n_xtraj, np * h->n_xtraj * sizeof(double), cudaMemcpyDeviceToHost, h->stream_out));
        CUDA_TRY(cudaMemcpyAsync(utraj + p0 * h->n_utraj, h->d_utraj + p0 * h->n_utraj, np * h->n_utraj * sizeof(double), cudaMemcpyDeviceToHost, h->stream_out));
        CUDA_TRY(cudaMemcpyAsync(info + p0, h->d_info + p0, np * sizeof(mpcgpu_info), cudaMemcpyDeviceToHost, h->stream_out));
        if (best) CUDA_TRY(cudaMemcpyAsync(best + s0, h->d_best + s0, ns * sizeof(int), cudaMemcpyDeviceToHost, h->stream_out));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream_out));
    if (h->timing && n_chunks == 1) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->ms_acc += ms;
        h->ms_n += 1;
    }
    return MPCGPU_OK;
}

// Debug / parity: the parameter blocks the device-side packing of the last scene call left in HBM, back in the reference's
// stage-major layout params[P][N][npar].
int mpcgpu_get_packed_params(mpcgpu_handle h, int P, double *params) {
    if (!h || P <= 0 || P > h->max_batch || !params) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->device));
    const int N = h->desc.N, npar = h->desc.npar, PT = pt_stride(N);
    double *tmp = (double *)malloc((size_t)P * npar * PT * sizeof(double));
    if (!tmp) return MPCGPU_ERR_ALLOC;
    cudaError_t e = cudaMemcpy(tmp, h->d_pt, (size_t)P * npar * PT * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        for (int p = 0; p < P; p++)
            for (int k = 0; k < N; k++)
                for (int i = 0; i < npar; i++) params[((size_t)p * N + k) * npar + i] = tmp[((size_t)p * npar + i) * PT + k];
    free(tmp);
    CUDA_TRY(e);
    return MPCGPU_OK;
}

int mpcgpu_state_size(mpcgpu_handle h) { return h ? (int)h->n_state : 0; }

int mpcgpu_set_stateful(mpcgpu_handle h, int enable) {
    if (!h) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (enable && !h->d_state) {
        CUDA_TRY(cudaMalloc(&h->d_state, (size_t)h->max_batch * h->n_state * sizeof(double)));
        CUDA_TRY(cudaMemset(h->d_state, 0, (size_t)h->max_batch * h->n_state * sizeof(double)));
    } else if (!enable && h->d_state) {
        CUDA_TRY(cudaFree(h->d_state));
        h->d_state = nullptr;
    }
    return MPCGPU_OK;
}

static int state_range_ok(mpcgpu_handle h, int first, int count) {
    return h && h->d_state && first >= 0 && count >= 0 && first + count <= h->max_batch;
}
int mpcgpu_reset_state(mpcgpu_handle h, int first, int count) {
    if (!state_range_ok(h, first, count)) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemsetAsync(h->d_state + (size_t)first * h->n_state, 0, (size_t)count * h->n_state * sizeof(double), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MPCGPU_OK;
}
int mpcgpu_get_state(mpcgpu_handle h, int first, int count, double *state) {
    if (!state_range_ok(h, first, count) || !state) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpy(state, h->d_state + (size_t)first * h->n_state, (size_t)count * h->n_state * sizeof(double), cudaMemcpyDeviceToHost));
    return MPCGPU_OK;
}
int mpcgpu_set_state(mpcgpu_handle h, int first, int count, const double *state) {
    if (!state_range_ok(h, first, count) || !state) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpy(h->d_state + (size_t)first * h->n_state, state, (size_t)count * h->n_state * sizeof(double), cudaMemcpyHostToDevice));
    return MPCGPU_OK;
}

int mpcgpu_set_group(mpcgpu_handle h, int group) {
    if (!h || group < 0) return MPCGPU_ERR_ARG;
    h->group = group;
    return MPCGPU_OK;
}

int mpcgpu_set_timeout(mpcgpu_handle h, double seconds) {
    if (!h || seconds != seconds) return MPCGPU_ERR_ARG;
    h->layout.timeout_s = seconds > 0.0 ? seconds : 0.0;
    return MPCGPU_OK;
}

long long mpcgpu_launch_count(mpcgpu_handle h) { return h ? h->launches : 0; }

int mpcgpu_set_timing(mpcgpu_handle h, int enabled) {
    if (!h) return MPCGPU_ERR_ARG;
    h->timing = enabled;
    h->ms_acc = 0.0;
    h->ms_n = 0;
    return MPCGPU_OK;
}
double mpcgpu_kernel_ms(mpcgpu_handle h, int *n_launches) {
    if (!h) return 0.0;
    const double v = h->ms_n ? h->ms_acc / h->ms_n : 0.0;
    if (n_launches) *n_launches = h->ms_n;
    h->ms_acc = 0.0;
    h->ms_n = 0;
    return v;
}

int mpcgpu_eval_stage(mpcgpu_handle h, int M, const double *z, const double *p, const double *lam, const double *pi, double *cost,
                      double *cost_grad, double *cost_hess, double *hval, double *h_jac, double *h_hess, double *xnext, double *AB,
                      double *dyn_hess) {
    if (!h || M <= 0 || !z || !p) return MPCGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->device));
    const int nz = h->desc.nx + h->desc.nu, nx = h->desc.nx, nh = h->desc.nh, npar = h->desc.npar;
    struct Buf {
        double *d;
        double *host;
        size_t n;
    };
    const double *ins[4] = {z, p, lam, pi};
    const size_t in_n[4] = {(size_t)M * nz, (size_t)M * npar, (size_t)M * nh, (size_t)M * nx};
    double *d_in[4] = {nullptr, nullptr, nullptr, nullptr};
    Buf outs[9] = {{nullptr, cost, (size_t)M},          {nullptr, cost_grad, (size_t)M * nz}, {nullptr, cost_hess, (size_t)M * nz * nz},
                   {nullptr, hval, (size_t)M * nh},     {nullptr, h_jac, (size_t)M * nh * nz}, {nullptr, h_hess, (size_t)M * nz * nz},
                   {nullptr, xnext, (size_t)M * nx},    {nullptr, AB, (size_t)M * nx * nz},   {nullptr, dyn_hess, (size_t)M * nz * nz}};
    int rc = MPCGPU_OK;
    for (int i = 0; i < 4 && rc == MPCGPU_OK; i++)
        if (ins[i] && in_n[i]) {
            if (cudaMalloc(&d_in[i], in_n[i] * sizeof(double)) != cudaSuccess ||
                cudaMemcpy(d_in[i], ins[i], in_n[i] * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
                rc = MPCGPU_ERR_CUDA;
        }
    for (int i = 0; i < 9 && rc == MPCGPU_OK; i++)
        if (outs[i].host && outs[i].n)
            if (cudaMalloc(&outs[i].d, outs[i].n * sizeof(double)) != cudaSuccess) rc = MPCGPU_ERR_CUDA;
    if (rc == MPCGPU_OK) {
        h->kt->launch_eval(h->layout, M, d_in[0], d_in[1], d_in[2], d_in[3], outs[0].d, outs[1].d, outs[2].d, outs[3].d, outs[4].d,
                           outs[5].d, outs[6].d, outs[7].d, outs[8].d, 0);
        h->launches += 1;
        if (cudaDeviceSynchronize() != cudaSuccess) rc = MPCGPU_ERR_CUDA;
    }
    for (int i = 0; i < 9 && rc == MPCGPU_OK; i++)
        if (outs[i].d && cudaMemcpy(outs[i].host, outs[i].d, outs[i].n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = MPCGPU_ERR_CUDA;
    if (rc != MPCGPU_OK) snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "eval_stage: %s", cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 4; i++) cudaFree(d_in[i]);
    for (int i = 0; i < 9; i++) cudaFree(outs[i].d);
    return rc;
}

double mpcgpu_measure_fp64_gflops(int device) {
    if (mpcgpu_device_count() <= device || device < 0) return -1.0;
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double *d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads>>>(d_out, 1024);  // warm up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d_out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        const double gf = flops / (ms * 1e-3) * 1e-9;
        if (gf > best) best = gf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}

}  // extern "C"
