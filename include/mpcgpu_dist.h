/*
 * mpcgpu_dist.h -- multi-GPU best-trajectory reduction of the B200 batched SQP backend (C ABI of libmpcgpu_dist.so).
 *
 * One process per GPU.  The P initial guesses of a T-MPC++ control step are split over the ranks in contiguous blocks
 * (SURVEY.md 8e); every rank solves its block with the single-GPU backend (include/mpcgpu.h) and the ranks then agree on
 * the winner with NCCL over NVLink: this replaces, across GPUs, `GuidanceConstraints::FindBestPlanner` and the copy of the
 * winner's output / info into the main solver (mpc_planner_modules/src/guidance_constraints.cpp:416-434, 382-384).
 *
 * Device-side protocol (no host synchronisation between the solve and the winner being available on every rank):
 *   1. local argmin per control step (select_best of the single-GPU path) -> (objective, global guess index) pair
 *   2. ncclAllGather of the pairs (16 B per control step and rank)
 *   3. every rank evaluates the same argmin over the gathered pairs in rank order (ties: lowest global index, as the
 *      reference's serial loop); the owner of the winner writes xtraj | utraj | info into the exchange buffer, all other
 *      ranks write zeros
 *   4. ncclAllReduce(sum) of the exchange buffer as 64-bit integers (bit patterns: x + 0 is exact, no -0 / NaN issue)
 * Failed solves never win (objective = the reference's 1e10 sentinel); if no rank has a successful solve, best = -1.
 *
 * Return convention: 0 = ok, negative = error (mpcgpu_dist_strerror).  libmpcgpu_dist.so links libnccl.so.2 and
 * libmpcgpu.so; the single-GPU library itself has no NCCL dependency.
 */
#ifndef MPCGPU_DIST_H
#define MPCGPU_DIST_H

#include "mpcgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MPCGPU_DIST_ID_BYTES 128 /* sizeof(ncclUniqueId) */

typedef struct mpcgpu_dist *mpcgpu_dist_handle;

const char *mpcgpu_dist_strerror(int code);

/* Rank 0 creates the rendezvous token (ncclGetUniqueId) and ships it to the other ranks by whatever channel the host
 * program has (bench.py: torch.distributed broadcast; a ROS-free C++ host: a file or a socket). */
int mpcgpu_dist_unique_id(void *id /* MPCGPU_DIST_ID_BYTES */);

/* Collective: ncclCommInitRank on the device of `solver`.  `solver` must outlive the returned handle. */
int mpcgpu_dist_create(mpcgpu_handle solver, int rank, int nranks, const void *id, mpcgpu_dist_handle *out);
int mpcgpu_dist_destroy(mpcgpu_dist_handle d);

/* S control steps whose guesses are split over the ranks: this rank holds guesses [g0, g0 + G_local) of each step,
 * device arrays laid out [S][G_local][...] exactly as mpcgpu_solve_device expects with P = S * G_local.
 * Asynchronous on `stream`.  Outputs, identical on every rank once the stream has drained:
 *   d_best_global[S]         global guess index of the winner of each step (-1: no successful solve on any rank)
 *   d_win_xtraj[S*nx*(N+1)], d_win_utraj[S*nu*N], d_win_info[S]   the winner's trajectory and info record
 * plus the local results in d_xtraj / d_utraj / d_info as in mpcgpu_solve_device. */
int mpcgpu_dist_solve_tmpc_device(mpcgpu_dist_handle d, int S, int G_local, int g0, const double *d_xinit, const double *d_x0,
                                  const double *d_params, double *d_xtraj, double *d_utraj, mpcgpu_info *d_info,
                                  int *d_best_global, double *d_win_xtraj, double *d_win_utraj, mpcgpu_info *d_win_info,
                                  void *stream);

/* Only the reduction (steps 1-4) on results that are already in HBM; d_best_local[S] = within-block winner per step
 * (-1 if none), as written by mpcgpu_solve_device with mpcgpu_set_group(G_local). */
int mpcgpu_dist_reduce_best(mpcgpu_dist_handle d, int S, int G_local, int g0, const double *d_xtraj, const double *d_utraj,
                            const mpcgpu_info *d_info, const int *d_best_local, int *d_best_global, double *d_win_xtraj,
                            double *d_win_utraj, mpcgpu_info *d_win_info, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MPCGPU_DIST_H */
