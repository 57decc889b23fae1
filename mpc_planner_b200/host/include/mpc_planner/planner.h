// planner.h -- `MPCPlanner::Planner` for the CUDA backend: the per-control-step orchestration of
// mpc_planner/src/planner.cpp:24-158 (+ reset :230-246, isObjectiveReached :248-254) without ROS, yaml-cpp, the
// profiling / visualisation / DataSaver calls.  Same call order as the reference:
//   warm start (initializeWarmstart | initializeWithBraking) -> setXinit -> module->update ->
//   module->setParameters(k = 0..N-1) -> loadWarmstart -> first module whose optimize() takes over
//   (GuidanceConstraints: ONE batched launch for all guesses), else solver->solve().
#ifndef MPC_PLANNER_CUDA_PLANNER_H
#define MPC_PLANNER_CUDA_PLANNER_H

#include <memory>
#include <string>
#include <vector>

#include <mpc_planner_modules/modules.h>

namespace MPCPlanner {

struct Trajectory {  // mpc_planner_types/data_types.h
    double dt{0.};
    int N{0};
    std::vector<Vec2> positions;
    Trajectory(double dt_ = 0., int N_ = 0) : dt(dt_), N(N_) {}
    void add(double x, double y) { positions.push_back({x, y}); }
};

struct PlannerOutput {  // mpc_planner/planner.h
    Trajectory trajectory;
    bool success{false};
    PlannerOutput(double dt, int N) : trajectory(dt, N) {}
    PlannerOutput() = default;
};

/** What the generated modules.h provides in the reference (generate_cpp_files.py: one emplace_back per module of
 *  the stack, in module order); here the stack follows the MPC_PLANNER_HAS_* macros of mpc_planner_generated.h. */
void initializeModules(std::vector<std::shared_ptr<ControllerModule>> &modules, std::shared_ptr<Solver> solver);

class Planner {
  public:
    Planner();
    PlannerOutput solveMPC(State &state, RealTimeData &data);
    double getSolution(int k, std::string &&var_name) const;
    void reset(State &state, RealTimeData &data, bool success = true);
    bool isObjectiveReached(const State &state, const RealTimeData &data) const;

    // inputs that reach the reference modules through ROS callbacks / external planners
    void setReferencePath(std::shared_ptr<Spline2D> path) { _path = path; }                    // onDataReceived("reference_path")
    void setGuidanceTrajectories(const std::vector<GuidanceTrajectory> &trajectories);           // GlobalGuidance::Update
    int bestPlannerIndex() const;

    bool shift_previous_solution_forward{true};  // CONFIG["shift_previous_solution_forward"] && CONFIG["enable_output"]
    int last_exit_flag{EXIT_CODE_NOT_OPTIMIZED_YET};
    std::shared_ptr<Solver> _solver;
    std::vector<std::shared_ptr<ControllerModule>> _modules;

  private:
    PlannerOutput _output;
    ModuleData _module_data;
    Trajectory _warmstart;
    std::shared_ptr<Spline2D> _path;
    bool _is_data_ready{false}, _was_reset{true};
};
}  // namespace MPCPlanner
#endif
