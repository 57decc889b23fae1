// planner.cpp -- see include/mpc_planner/planner.h.  Reference: mpc_planner/src/planner.cpp.
#include <mpc_planner/planner.h>
#include <mpc_planner_util/profiling.h>

namespace MPCPlanner {

void initializeModules(std::vector<std::shared_ptr<ControllerModule>> &modules, std::shared_ptr<Solver> solver) {
    modules.clear();
    modules.emplace_back(std::make_shared<MPCBaseModule>(solver));
    modules.emplace_back(std::make_shared<Contouring>(solver));
#if defined(MPC_PLANNER_HAS_LINEAR) && defined(MPC_PLANNER_HAS_ELLIPSOID)
    modules.emplace_back(std::make_shared<GuidanceConstraints>(solver, CONFIG().n_paths));  // T-MPC++ adds the non-guided planner
#elif defined(MPC_PLANNER_HAS_ELLIPSOID)
    modules.emplace_back(std::make_shared<EllipsoidConstraints>(solver));
#elif defined(MPC_PLANNER_HAS_GAUSSIAN)
    modules.emplace_back(std::make_shared<GaussianConstraints>(solver));
#endif
#if defined(MPC_PLANNER_HAS_DECOMP)
    modules.emplace_back(std::make_shared<DecompConstraints>(solver));
#endif
}

Planner::Planner() {  // :24-35
    _solver = std::make_shared<Solver>();
    _solver->reset();
    initializeModules(_modules, _solver);
}

void Planner::setGuidanceTrajectories(const std::vector<GuidanceTrajectory> &trajectories) {
    for (auto &module : _modules)
        if (auto *g = dynamic_cast<GuidanceConstraints *>(module.get())) g->setGuidanceTrajectories(trajectories);
}

int Planner::bestPlannerIndex() const {
    for (auto &module : _modules)
        if (auto *g = dynamic_cast<GuidanceConstraints *>(module.get())) return g->best_planner_index();
    return -1;
}

PlannerOutput Planner::solveMPC(State &state, RealTimeData &data) {  // :37-158
    PROFILE_SCOPE("Planning");  // :69
    const bool was_feasible = _output.success;
    _output = PlannerOutput(_solver->dt, _solver->N);
    _module_data = ModuleData();  // reset module data (:43); the path arrives through onDataReceived in the reference
    _module_data.path = _path;

    _is_data_ready = (_path != nullptr);
    std::string missing_data;
    for (auto &module : _modules) _is_data_ready = _is_data_ready & module->isDataReady(data, missing_data);
    if (!_is_data_ready) {
        _output.success = false;
        return _output;
    }
    _was_reset = false;

    // initial guess (:78-86)
    if (was_feasible)
        _solver->initializeWarmstart(state, shift_previous_solution_forward);
    else
        _solver->initializeWithBraking(state);
    _solver->setXinit(state);  // :88

    {
        PROFILE_SCOPE("Update");  // :93
        for (auto &module : _modules) module->update(state, data, _module_data);  // :95-96
    }
    _solver->setXinit(state);  // Contouring::update writes the spline state (contouring.cpp:37,42)
    {
        PROFILE_SCOPE("SetParameters");  // :101
        for (int k = 0; k < _solver->N; k++)  // :102-108
            for (auto &module : _modules) module->setParameters(data, _module_data, k);
    }

    _warmstart = Trajectory();
    for (int k = 0; k < _solver->N; k++) _warmstart.add(_solver->getEgoPrediction(k, "x"), _solver->getEgoPrediction(k, "y"));
    _solver->loadWarmstart();  // :115
    _solver->_params.solver_timeout = remainingPlanningTime(data, 1. / CONFIG().control_frequency, 0.006);  // :117-118

    int exit_flag = EXIT_CODE_NOT_OPTIMIZED_YET;  // :125-133
    {
        PROFILE_SCOPE("Optimization");  // :123
        for (auto &module : _modules) {
            exit_flag = module->optimize(state, data, _module_data);
            if (exit_flag != EXIT_CODE_NOT_OPTIMIZED_YET) break;
        }
        if (exit_flag == EXIT_CODE_NOT_OPTIMIZED_YET) exit_flag = _solver->solve();
    }
    last_exit_flag = exit_flag;

    if (exit_flag != 1) {  // :140-146
        _output.success = false;
        return _output;
    }
    _output.success = true;
    for (int k = 1; k < _solver->N; k++) _output.trajectory.add(_solver->getOutput(k, "x"), _solver->getOutput(k, "y"));
    return _output;
}

double Planner::getSolution(int k, std::string &&var_name) const { return _solver->getOutput(k, var_name); }

void Planner::reset(State &state, RealTimeData &data, bool) {  // :230-246
    _solver->reset();
    for (auto &module : _modules) module->reset();
    state = State();
    data = RealTimeData();
    _output = PlannerOutput();
    _was_reset = true;
}

bool Planner::isObjectiveReached(const State &state, const RealTimeData &data) const {  // :248-254
    bool objective_reached = true;
    for (auto &module : _modules) objective_reached = objective_reached && module->isObjectiveReached(state, data);
    return objective_reached;
}
}  // namespace MPCPlanner
