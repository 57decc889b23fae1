// profiling.h -- PROFILE_SCOPE / PROFILE_FUNCTION of the reference (ros_tools Instrumentor, un-vendored; call sites
// mpc_planner/src/planner.cpp:69,93,101,123 and mpc_planner_modules/src/guidance_constraints.cpp:266,282,366) as NVTX ranges:
// with -DMPC_PLANNER_NVTX the host phases of a control step show up on the Nsight Systems / ncu timeline next to the
// kernels they launch; without it the macros vanish (as the reference's do without the profiler enabled).
#pragma once

#if defined(MPC_PLANNER_NVTX)
#include <nvtx3/nvToolsExt.h>
namespace MPCPlanner {
struct NvtxScope {
    explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
    NvtxScope(const NvtxScope &) = delete;
    NvtxScope &operator=(const NvtxScope &) = delete;
};
}  // namespace MPCPlanner
#define MPC_PROFILE_CAT2(a, b) a##b
#define MPC_PROFILE_CAT(a, b) MPC_PROFILE_CAT2(a, b)
#define PROFILE_SCOPE(name) ::MPCPlanner::NvtxScope MPC_PROFILE_CAT(nvtx_scope_, __LINE__)(name)
#define PROFILE_FUNCTION() PROFILE_SCOPE(__func__)
#else
#define PROFILE_SCOPE(name)
#define PROFILE_FUNCTION()
#endif
