// mpcgpu_dist.cu -- multi-GPU best-trajectory reduction (include/mpcgpu_dist.h): NCCL over NVLink 5 / NVSwitch, one
// process per GPU.  Replaces, across GPUs, GuidanceConstraints::FindBestPlanner and the winner copy
// (mpc_planner_modules/src/guidance_constraints.cpp:416-434, 382-384).
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "mpcgpu_dist.h"

namespace {

enum { DIST_OK = 0, DIST_ERR_ARG = -1, DIST_ERR_CUDA = -3, DIST_ERR_NCCL = -7, DIST_ERR_SOLVER = -8 };

thread_local char g_err[256] = "";

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            snprintf(g_err, sizeof(g_err), "%s: %s", #expr, cudaGetErrorString(e__));      \
            return DIST_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)
#define NCCL_TRY(expr)                                                                     \
    do {                                                                                   \
        ncclResult_t r__ = (expr);                                                         \
        if (r__ != ncclSuccess) {                                                          \
            snprintf(g_err, sizeof(g_err), "%s: %s", #expr, ncclGetErrorString(r__));      \
            return DIST_ERR_NCCL;                                                          \
        }                                                                                  \
    } while (0)

// (objective, global index) of the local winner of each control step; objective = 1e10 sentinel when the block has no
// successful solve (guidance_constraints.cpp:419)
__global__ void local_pairs_kernel(const mpcgpu_info *__restrict__ info, const int *__restrict__ best_local, int S, int G_local, int g0,
                                   double *__restrict__ pairs) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int b = best_local[s];
    pairs[2 * s] = (b >= 0) ? info[(size_t)s * G_local + b].pobj : 1e10;
    pairs[2 * s + 1] = (b >= 0) ? (double)(g0 + b) : -1.0;
}

// One CTA per control step: every rank derives the same winner from the gathered pairs; the owner fills the exchange
// record [best | xtraj | utraj | info] (64-bit words), everybody else zeros.
__global__ void fill_exchange_kernel(const double *__restrict__ gathered, int nranks, int rank, int S, int G_local, int g0,
                                     const double *__restrict__ xtraj, const double *__restrict__ utraj, const mpcgpu_info *__restrict__ info,
                                     int n_xtraj, int n_utraj, unsigned long long *__restrict__ exch, int words) {
    const int s = blockIdx.x;
    __shared__ int s_owner, s_best;
    if (threadIdx.x == 0) {
        double bo = 1e10;  // strict '<' in rank (= global index) order: ties resolve to the lowest global index
        int owner = -1, best = -1;
        for (int r = 0; r < nranks; r++) {
            const double o = gathered[((size_t)r * S + s) * 2], idx = gathered[((size_t)r * S + s) * 2 + 1];
            if (idx >= 0.0 && o < bo) {
                bo = o;
                owner = r;
                best = (int)idx;
            }
        }
        s_owner = owner;
        s_best = best;
    }
    __syncthreads();
    unsigned long long *rec = exch + (size_t)s * words;
    const bool mine = (s_owner == rank);
    const int b = s_best - g0;  // local index on the owner
    const size_t p = (size_t)s * G_local + (mine ? b : 0);
    const int n_info = (int)(sizeof(mpcgpu_info) / 8);
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
        unsigned long long v = 0ull;
        if (mine) {
            if (i == 0) v = (unsigned long long)(long long)s_best;
            else if (i < 1 + n_xtraj) v = (unsigned long long)__double_as_longlong(xtraj[p * n_xtraj + (i - 1)]);
            else if (i < 1 + n_xtraj + n_utraj) v = (unsigned long long)__double_as_longlong(utraj[p * n_utraj + (i - 1 - n_xtraj)]);
            else v = reinterpret_cast<const unsigned long long *>(info + p)[i - 1 - n_xtraj - n_utraj];
        } else if (s_owner < 0 && i == 0 && rank == 0) {
            v = (unsigned long long)(long long)-1;  // nobody succeeded: rank 0 contributes best = -1
        }
        rec[i] = v;
    }
    (void)n_info;
}

__global__ void unpack_exchange_kernel(const unsigned long long *__restrict__ exch, int words, int S, int n_xtraj, int n_utraj,
                                       int *__restrict__ best_global, double *__restrict__ win_xtraj, double *__restrict__ win_utraj,
                                       mpcgpu_info *__restrict__ win_info) {
    const int s = blockIdx.x;
    const unsigned long long *rec = exch + (size_t)s * words;
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
        const unsigned long long v = rec[i];
        if (i == 0) best_global[s] = (int)(long long)v;
        else if (i < 1 + n_xtraj) win_xtraj[(size_t)s * n_xtraj + (i - 1)] = __longlong_as_double((long long)v);
        else if (i < 1 + n_xtraj + n_utraj) win_utraj[(size_t)s * n_utraj + (i - 1 - n_xtraj)] = __longlong_as_double((long long)v);
        else reinterpret_cast<unsigned long long *>(win_info + s)[i - 1 - n_xtraj - n_utraj] = v;
    }
}

}  // namespace

struct mpcgpu_dist {
    mpcgpu_handle solver;
    ncclComm_t comm;
    int rank, nranks, device;
    int n_xtraj, n_utraj, words, max_steps;
    double *d_pairs, *d_gathered;
    unsigned long long *d_exch;
    int *d_best_local;
};

extern "C" {

const char *mpcgpu_dist_strerror(int code) {
    switch (code) {
    case DIST_OK: return "ok";
    case DIST_ERR_ARG: return "invalid argument";
    case DIST_ERR_CUDA: return g_err[0] ? g_err : "CUDA error";
    case DIST_ERR_NCCL: return g_err[0] ? g_err : "NCCL error";
    case DIST_ERR_SOLVER: return g_err[0] ? g_err : "single-GPU solver error";
    }
    return "unknown error";
}

int mpcgpu_dist_unique_id(void *id) {
    if (!id) return DIST_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == MPCGPU_DIST_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NCCL_TRY(ncclGetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return DIST_OK;
}

int mpcgpu_dist_create(mpcgpu_handle solver, int rank, int nranks, const void *id, mpcgpu_dist_handle *out) {
    if (!solver || !id || !out || nranks < 1 || rank < 0 || rank >= nranks) return DIST_ERR_ARG;
    *out = nullptr;
    mpcgpu_problem_desc desc;
    int device = -1, max_batch = 0;
    if (mpcgpu_describe(solver, &desc, &device, &max_batch) != 0) return DIST_ERR_ARG;
    CUDA_TRY(cudaSetDevice(device));
    mpcgpu_dist *d = new (std::nothrow) mpcgpu_dist();
    if (!d) return DIST_ERR_ARG;
    memset(d, 0, sizeof(*d));
    d->solver = solver;
    d->rank = rank;
    d->nranks = nranks;
    d->device = device;
    d->n_xtraj = desc.nx * (desc.N + 1);
    d->n_utraj = desc.nu * desc.N;
    d->words = 1 + d->n_xtraj + d->n_utraj + (int)(sizeof(mpcgpu_info) / 8);
    d->max_steps = max_batch;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = ncclCommInitRank(&d->comm, nranks, u, rank);
    if (r != ncclSuccess) {
        snprintf(g_err, sizeof(g_err), "ncclCommInitRank: %s", ncclGetErrorString(r));
        delete d;
        return DIST_ERR_NCCL;
    }
    bool ok = cudaMalloc(&d->d_pairs, (size_t)d->max_steps * 2 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_gathered, (size_t)d->max_steps * 2 * nranks * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_exch, (size_t)d->max_steps * d->words * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_best_local, (size_t)d->max_steps * sizeof(int)) == cudaSuccess;
    if (!ok) {
        snprintf(g_err, sizeof(g_err), "device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        mpcgpu_dist_destroy(d);
        return DIST_ERR_CUDA;
    }
    *out = d;
    return DIST_OK;
}

int mpcgpu_dist_destroy(mpcgpu_dist_handle d) {
    if (!d) return DIST_ERR_ARG;
    cudaSetDevice(d->device);
    cudaFree(d->d_pairs);
    cudaFree(d->d_gathered);
    cudaFree(d->d_exch);
    cudaFree(d->d_best_local);
    if (d->comm) ncclCommDestroy(d->comm);
    delete d;
    return DIST_OK;
}

int mpcgpu_dist_reduce_best(mpcgpu_dist_handle d, int S, int G_local, int g0, const double *d_xtraj, const double *d_utraj,
                            const mpcgpu_info *d_info, const int *d_best_local, int *d_best_global, double *d_win_xtraj,
                            double *d_win_utraj, mpcgpu_info *d_win_info, void *stream) {
    if (!d || S <= 0 || S > d->max_steps || G_local <= 0 || g0 < 0 || !d_xtraj || !d_utraj || !d_info || !d_best_local || !d_best_global ||
        !d_win_xtraj || !d_win_utraj || !d_win_info)
        return DIST_ERR_ARG;
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t s = (cudaStream_t)stream;
    local_pairs_kernel<<<(S + 127) / 128, 128, 0, s>>>(d_info, d_best_local, S, G_local, g0, d->d_pairs);
    NCCL_TRY(ncclAllGather(d->d_pairs, d->d_gathered, (size_t)2 * S, ncclFloat64, d->comm, s));
    fill_exchange_kernel<<<S, 128, 0, s>>>(d->d_gathered, d->nranks, d->rank, S, G_local, g0, d_xtraj, d_utraj, d_info, d->n_xtraj, d->n_utraj,
                                          d->d_exch, d->words);
    NCCL_TRY(ncclAllReduce(d->d_exch, d->d_exch, (size_t)S * d->words, ncclUint64, ncclSum, d->comm, s));
    unpack_exchange_kernel<<<S, 128, 0, s>>>(d->d_exch, d->words, S, d->n_xtraj, d->n_utraj, d_best_global, d_win_xtraj, d_win_utraj, d_win_info);
    CUDA_TRY(cudaGetLastError());
    return DIST_OK;
}

int mpcgpu_dist_solve_tmpc_device(mpcgpu_dist_handle d, int S, int G_local, int g0, const double *d_xinit, const double *d_x0,
                                  const double *d_params, double *d_xtraj, double *d_utraj, mpcgpu_info *d_info,
                                  int *d_best_global, double *d_win_xtraj, double *d_win_utraj, mpcgpu_info *d_win_info,
                                  void *stream) {
    if (!d || S <= 0 || G_local <= 0 || S > d->max_steps) return DIST_ERR_ARG;
    int rc = mpcgpu_set_group(d->solver, G_local);
    if (rc == 0) rc = mpcgpu_solve_device(d->solver, S * G_local, d_xinit, d_x0, d_params, d_xtraj, d_utraj, d_info, d->d_best_local, stream);
    if (rc != 0) {
        snprintf(g_err, sizeof(g_err), "mpcgpu_solve_device: %s", mpcgpu_strerror(rc));
        return DIST_ERR_SOLVER;
    }
    return mpcgpu_dist_reduce_best(d, S, G_local, g0, d_xtraj, d_utraj, d_info, d->d_best_local, d_best_global, d_win_xtraj, d_win_utraj,
                                   d_win_info, stream);
}

}  // extern "C"
