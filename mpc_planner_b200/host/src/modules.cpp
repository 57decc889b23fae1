// modules.cpp -- see mpc_planner_modules/modules.h.  Each function cites the reference code it mirrors.
#include <mpc_planner_modules/modules.h>

#include <algorithm>

#include <mpc_planner_util/profiling.h>

#include "../../csrc/projection.cuh"
#include <cmath>

namespace MPCPlanner {

Config &CONFIG() {
    static Config c = [] {
        Config d;
        d.weights = {{"acceleration", 0.34}, {"angular_velocity", 0.85}, {"velocity", 0.55}, {"reference_velocity", 2.0},
                     {"contour", 0.05}, {"lag", 0.75}, {"slack", 10000.}, {"terminal_angle", 100.0}, {"terminal_contouring", 10.0}};
        d.max_obstacles = 8;
        return d;
    }();
    return c;
}

Vec2 Disc::getPosition(const Vec2 &robot_position, double angle) const {  // data_types.cpp: pos + R(angle) [offset, 0]
    return {robot_position[0] + offset * std::cos(angle), robot_position[1] + offset * std::sin(angle)};
}

void propagatePredictionUncertainty(Prediction &prediction, double dt, int N) {  // data_preparation.cpp:170-186
    if (prediction.type != PredictionType::GAUSSIAN) return;
    double major = 0., minor = 0.;
    for (int k = 0; k < N && k < (int)prediction.modes[0].size(); k++) {
        major = std::sqrt(major * major + std::pow(prediction.modes[0][k].major_radius * dt, 2.));
        minor = std::sqrt(minor * minor + std::pow(prediction.modes[0][k].minor_radius * dt, 2.));
        prediction.modes[0][k].major_radius = major;
        prediction.modes[0][k].minor_radius = minor;
    }
}

Vec2 Spline2D::getPoint(double s) const {
    int i = 0;
    while (i + 1 < (int)segments.size() && s >= segments[i + 1].start) i++;
    const PathSegment &g = segments[i];
    const double t = s - g.start;
    return {((g.ax * t + g.bx) * t + g.cx) * t + g.dx, ((g.ay * t + g.by) * t + g.cy) * t + g.dy};
}
void Spline2D::findClosestPoint(const Vec2 &pos, int &segment_out, double &s_out) const {
    double best = 1e300;
    const double s_end = segments.empty() ? 0. : (length > 0. ? length : segments.back().start + 5.0);
    for (double s = segments.empty() ? 0. : segments[0].start; s <= s_end; s += 0.01) {
        const Vec2 p = getPoint(s);
        const double d = (p[0] - pos[0]) * (p[0] - pos[0]) + (p[1] - pos[1]) * (p[1] - pos[1]);
        if (d < best) {
            best = d;
            s_out = s;
        }
    }
    segment_out = 0;
    while (segment_out + 1 < (int)segments.size() && s_out >= segments[segment_out + 1].start) segment_out++;
}

// ---- MPCBaseModule (mpc_base.cpp:12-35) ----
MPCBaseModule::MPCBaseModule(std::shared_ptr<Solver> solver) : ControllerModule(ModuleType::OBJECTIVE, solver, "mpc_base") {
    _weight_names = WEIGHT_PARAMS;
}
void MPCBaseModule::setParameters(const RealTimeData &, const ModuleData &, int k) {
    for (auto &weight : _weight_names) _solver->setParameter(k, weight, CONFIG().weights.at(weight));
}

// ---- Contouring (contouring.cpp:12-124) ----
Contouring::Contouring(std::shared_ptr<Solver> solver) : ControllerModule(ModuleType::OBJECTIVE, solver, "contouring") {
    _n_segments = CONFIG().num_segments;
}
void Contouring::update(State &state, const RealTimeData &, ModuleData &module_data) {  // :28-48
    _spline = module_data.path;
    double closest_s = 0.;
    _spline->findClosestPoint(state.getPos(), _closest_segment, closest_s);
    if (module_data.current_path_segment == -1) module_data.current_path_segment = _closest_segment;
    state.set("spline", closest_s);  // the solver's spline state starts at the closest point
    _closest_segment = std::min(_closest_segment, std::max(0, _spline->numSegments() - _n_segments));
}
bool Contouring::isObjectiveReached(const State &state, const RealTimeData &) {  // :167-179
    if (!_spline) return false;
    const Vec2 end = _spline->getPoint(_spline->length), pos = state.getPos();
    return std::hypot(pos[0] - end[0], pos[1] - end[1]) < 1.0;
}
void Contouring::setParameters(const RealTimeData &, const ModuleData &, int k) {  // :50-92
    setSolverParameterContour(k, _solver->_params, CONFIG().weights.at("contour"));
    setSolverParameterLag(k, _solver->_params, CONFIG().weights.at("lag"));
    setSolverParameterTerminalAngle(k, _solver->_params, CONFIG().weights.at("terminal_angle"));
    setSolverParameterTerminalContouring(k, _solver->_params, CONFIG().weights.at("terminal_contouring"));
    setSplineParameters(k);
}
void Contouring::setSplineParameters(int k) {  // :94-124
    for (int i = 0; i < _n_segments; i++) {
        const PathSegment &g = _spline->segments[_closest_segment + i];
        setSolverParameterSplineXA(k, _solver->_params, g.ax, i);
        setSolverParameterSplineXB(k, _solver->_params, g.bx, i);
        setSolverParameterSplineXC(k, _solver->_params, g.cx, i);
        setSolverParameterSplineXD(k, _solver->_params, g.dx, i);
        setSolverParameterSplineYA(k, _solver->_params, g.ay, i);
        setSolverParameterSplineYB(k, _solver->_params, g.by, i);
        setSolverParameterSplineYC(k, _solver->_params, g.cy, i);
        setSolverParameterSplineYD(k, _solver->_params, g.dy, i);
        setSolverParameterSplineStart(k, _solver->_params, g.start, i);
    }
}

#if defined(MPC_PLANNER_HAS_ELLIPSOID)
// ---- EllipsoidConstraints (ellipsoid_constraints.cpp:12-90) ----
EllipsoidConstraints::EllipsoidConstraints(std::shared_ptr<Solver> solver)
    : ControllerModule(ModuleType::CONSTRAINT, solver, "ellipsoid_constraints") {
    _n_discs = CONFIG().n_discs;
    _robot_radius = CONFIG().robot_radius;
    _risk = CONFIG().risk;
}
void EllipsoidConstraints::update(State &state, const RealTimeData &, ModuleData &) {  // :24-32
    _dummy_x = state.get("x") + 50;
    _dummy_y = state.get("y") + 50;
}
void EllipsoidConstraints::setParameters(const RealTimeData &data, const ModuleData &, int k) {  // :34-90
    setSolverParameterEgoDiscRadius(k, _solver->_params, _robot_radius);
    for (int d = 0; d < _n_discs; d++) setSolverParameterEgoDiscOffset(k, _solver->_params, data.robot_area[d].offset, d);
    if (k == 0) {  // dummies
        for (size_t i = 0; i < data.dynamic_obstacles.size(); i++) {
            setSolverParameterEllipsoidObstX(0, _solver->_params, _dummy_x, i);
            setSolverParameterEllipsoidObstY(0, _solver->_params, _dummy_y, i);
            setSolverParameterEllipsoidObstPsi(0, _solver->_params, 0., i);
            setSolverParameterEllipsoidObstR(0, _solver->_params, 0.1, i);
            setSolverParameterEllipsoidObstMajor(0, _solver->_params, 0., i);
            setSolverParameterEllipsoidObstMinor(0, _solver->_params, 0., i);
            setSolverParameterEllipsoidObstChi(0, _solver->_params, 1., i);
        }
        return;
    }
    for (size_t i = 0; i < data.dynamic_obstacles.size(); i++) {
        const auto &obstacle = data.dynamic_obstacles[i];
        const auto &mode = obstacle.prediction.modes[0];
        // the first prediction step is index 1 of the optimisation problem: k-1 maps to the predictions of stage k
        setSolverParameterEllipsoidObstX(k, _solver->_params, mode[k - 1].position[0], i);
        setSolverParameterEllipsoidObstY(k, _solver->_params, mode[k - 1].position[1], i);
        setSolverParameterEllipsoidObstPsi(k, _solver->_params, mode[k - 1].angle, i);
        setSolverParameterEllipsoidObstR(k, _solver->_params, obstacle.radius, i);
        if (obstacle.prediction.type == PredictionType::DETERMINISTIC) {
            setSolverParameterEllipsoidObstMajor(k, _solver->_params, 0., i);
            setSolverParameterEllipsoidObstMinor(k, _solver->_params, 0., i);
            setSolverParameterEllipsoidObstChi(k, _solver->_params, 1., i);
        } else if (obstacle.prediction.type == PredictionType::GAUSSIAN) {
            const double chi = -std::log(_risk) / 0.5;  // RosTools::ExponentialQuantile(0.5, 1 - risk)
            setSolverParameterEllipsoidObstMajor(k, _solver->_params, mode[k - 1].major_radius, i);
            setSolverParameterEllipsoidObstMinor(k, _solver->_params, mode[k - 1].minor_radius, i);
            setSolverParameterEllipsoidObstChi(k, _solver->_params, chi, i);
        }
    }
}

#endif  // ELLIPSOID
#ifdef MPC_PLANNER_HAS_GAUSSIAN
// ---- GaussianConstraints (gaussian_constraints.cpp:14-79) ----
GaussianConstraints::GaussianConstraints(std::shared_ptr<Solver> solver)
    : ControllerModule(ModuleType::CONSTRAINT, solver, "gaussian_constraints") {}
void GaussianConstraints::update(State &state, const RealTimeData &, ModuleData &) {
    _dummy_x = state.get("x") + 100.;
    _dummy_y = state.get("y") + 100.;
}
void GaussianConstraints::setParameters(const RealTimeData &data, const ModuleData &, int k) {
    setSolverParameterEgoDiscRadius(k, _solver->_params, CONFIG().robot_radius);
    for (int d = 0; d < CONFIG().n_discs; d++) setSolverParameterEgoDiscOffset(k, _solver->_params, data.robot_area[d].offset, d);
    if (k == 0) {
        for (size_t i = 0; i < data.dynamic_obstacles.size(); i++) {
            setSolverParameterGaussianObstX(k, _solver->_params, _dummy_x, i);
            setSolverParameterGaussianObstY(k, _solver->_params, _dummy_y, i);
            setSolverParameterGaussianObstMajor(k, _solver->_params, 0.1, i);
            setSolverParameterGaussianObstMinor(k, _solver->_params, 0.1, i);
            setSolverParameterGaussianObstRisk(k, _solver->_params, 0.05, i);
            setSolverParameterGaussianObstR(k, _solver->_params, 0.1, i);
        }
        return;
    }
    for (size_t i = 0; i < data.dynamic_obstacles.size(); i++) {
        const auto &obstacle = data.dynamic_obstacles[i];
        if (obstacle.prediction.type != PredictionType::GAUSSIAN) continue;
        const auto &step = obstacle.prediction.modes[0][k - 1];
        setSolverParameterGaussianObstX(k, _solver->_params, step.position[0], i);
        setSolverParameterGaussianObstY(k, _solver->_params, step.position[1], i);
        const bool dynamic = obstacle.type == ObstacleType::DYNAMIC;  // static obstacles have no uncertainty (:72-73)
        setSolverParameterGaussianObstMajor(k, _solver->_params, dynamic ? step.major_radius : 0.001, i);
        setSolverParameterGaussianObstMinor(k, _solver->_params, dynamic ? step.minor_radius : 0.001, i);
        setSolverParameterGaussianObstRisk(k, _solver->_params, CONFIG().risk, i);
        setSolverParameterGaussianObstR(k, _solver->_params, CONFIG().obstacle_radius, i);
    }
}
#endif

#if defined(MPC_PLANNER_HAS_LINEAR)
// ---- LinearizedConstraints (linearized_constraints.cpp:12-189) ----
LinearizedConstraints::LinearizedConstraints(std::shared_ptr<Solver> solver)
    : ControllerModule(ModuleType::CONSTRAINT, solver, "linearized_constraints") {
    _n_discs = CONFIG().n_discs;
    _n_other_halfspaces = CONFIG().add_halfspaces;
    _max_obstacles = CONFIG().max_obstacles;
    const int n_constraints = _max_obstacles + _n_other_halfspaces;
    _a1.assign(_n_discs, std::vector<std::vector<double>>(_solver->N, std::vector<double>(n_constraints, 0.)));
    _a2 = _b = _a1;
}
void LinearizedConstraints::setTopologyConstraints() {  // :43-47
    _n_discs = 1;
    _use_guidance = true;
}
void LinearizedConstraints::update(State &state, const RealTimeData &data, ModuleData &module_data) {  // :49-128
    _dummy_b = state.get("x") + 100.;
    std::vector<DynamicObstacle> copied_obstacles = data.dynamic_obstacles;  // thread safe
    _num_obstacles = (int)copied_obstacles.size();
    for (int k = 1; k < _solver->N; k++) {
        for (int d = 0; d < _n_discs; d++) {
            Vec2 pos = _solver->getEgoPredictionPosition(k);  // k = 0 is the initial state
            if (!_use_guidance) pos = data.robot_area[d].getPosition(pos, _solver->getEgoPrediction(k, "psi"));
            projectToSafety(copied_obstacles, k, pos);  // ensure that the vehicle position is collision-free
            for (size_t obs_id = 0; obs_id < copied_obstacles.size(); obs_id++) {
                const auto &o = copied_obstacles[obs_id];
                const Vec2 &op = o.prediction.modes[0][k - 1].position;
                const double diff_x = op[0] - pos[0], diff_y = op[1] - pos[1];
                const double dist = std::sqrt(diff_x * diff_x + diff_y * diff_y);
                _a1[d][k][obs_id] = diff_x / dist;  // normalised normal vector
                _a2[d][k][obs_id] = diff_y / dist;
                const double radius = _use_guidance ? 1e-3 : o.radius;
                _b[d][k][obs_id] = _a1[d][k][obs_id] * op[0] + _a2[d][k][obs_id] * op[1] - (radius + CONFIG().robot_radius);
            }
            if (!module_data.static_obstacles.empty()) {
                const int num_halfspaces = std::min((int)module_data.static_obstacles[k].size(), _n_other_halfspaces);
                for (int h = 0; h < num_halfspaces; h++) {
                    const int obs_id = (int)copied_obstacles.size() + h;
                    _a1[d][k][obs_id] = module_data.static_obstacles[k][h].A[0];
                    _a2[d][k][obs_id] = module_data.static_obstacles[k][h].A[1];
                    _b[d][k][obs_id] = module_data.static_obstacles[k][h].b;
                }
            }
        }
    }
}
void LinearizedConstraints::projectToSafety(const std::vector<DynamicObstacle> &copied_obstacles, int k, Vec2 &pos) {  // :130-148
    if (copied_obstacles.empty()) return;
    // RosTools::DouglasRachford::douglasRachfordProjection (ros_tools, un-vendored): restated in csrc/projection.cuh, the
    // one definition host and device share.  At most 3 sweeps, anchor = obstacle 0 of the stage (:136-146).
    const Vec2 &anchor = copied_obstacles[0].prediction.modes[0][k - 1].position;
    for (int iterate = 0; iterate < 3; iterate++)
        for (const auto &obstacle : copied_obstacles) {
            const double r = (_use_guidance ? 1e-3 : obstacle.radius) + CONFIG().robot_radius;
            const Vec2 &op = obstacle.prediction.modes[0][k - 1].position;
            mpcgpu::douglas_rachford_projection(pos[0], pos[1], op[0], op[1], anchor[0], anchor[1], r);
        }
}
void LinearizedConstraints::setParameters(const RealTimeData &data, const ModuleData &, int k) {  // :150-189
    int constraint_counter = 0;
    if (k == 0) {
        for (int i = 0; i < _max_obstacles + _n_other_halfspaces; i++) {
            setSolverParameterLinConstraintA1(0, _solver->_params, _dummy_a1, constraint_counter);
            setSolverParameterLinConstraintA2(0, _solver->_params, _dummy_a2, constraint_counter);
            setSolverParameterLinConstraintB(0, _solver->_params, _dummy_b, constraint_counter);
            constraint_counter++;
        }
        return;
    }
    for (int d = 0; d < _n_discs; d++) {
        for (size_t i = 0; i < data.dynamic_obstacles.size() + _n_other_halfspaces; i++) {
            setSolverParameterLinConstraintA1(k, _solver->_params, _a1[d][k][i], constraint_counter);
            setSolverParameterLinConstraintA2(k, _solver->_params, _a2[d][k][i], constraint_counter);
            setSolverParameterLinConstraintB(k, _solver->_params, _b[d][k][i], constraint_counter);
            constraint_counter++;
        }
        for (int i = (int)data.dynamic_obstacles.size() + _n_other_halfspaces; i < _max_obstacles + _n_other_halfspaces; i++) {
            setSolverParameterLinConstraintA1(k, _solver->_params, _dummy_a1, constraint_counter);
            setSolverParameterLinConstraintA2(k, _solver->_params, _dummy_a2, constraint_counter);
            setSolverParameterLinConstraintB(k, _solver->_params, _dummy_b, constraint_counter);
            constraint_counter++;
        }
    }
}

#endif  // LINEAR
#if defined(MPC_PLANNER_HAS_DECOMP)
// ---- DecompConstraints (decomp_constraints.cpp:90-114, 150-189): polyhedra come in through ModuleData ----
DecompConstraints::DecompConstraints(std::shared_ptr<Solver> solver) : ControllerModule(ModuleType::CONSTRAINT, solver, "decomp_constraints") {
    _max_constraints = CONFIG().decomp_max_constraints;
    _a1.assign(_solver->N, std::vector<double>(_max_constraints, 0.));
    _a2 = _b = _a1;
}
void DecompConstraints::update(State &state, const RealTimeData &, ModuleData &module_data) {
    _dummy_b = state.get("x") + 100.;
    for (int k = 0; k < _solver->N - 1 && k < (int)module_data.decomp_polyhedra.size(); k++) {  // polyhedron k -> stage k + 1 (:90-114)
        const auto &poly = module_data.decomp_polyhedra[k];
        int n_valid = 0;  // rows up to the first zero / NaN normal are used (:97-100), dummies after it
        while (n_valid < (int)poly.size() && n_valid < _max_constraints) {
            const double a0 = poly[n_valid].A[0], a1 = poly[n_valid].A[1];
            if (std::sqrt(a0 * a0 + a1 * a1) < 1e-3 || a0 != a0) break;
            n_valid++;
        }
        for (int i = 0; i < _max_constraints; i++) {
            const bool have = i < n_valid;
            _a1[k + 1][i] = have ? poly[i].A[0] : _dummy_a1;
            _a2[k + 1][i] = have ? poly[i].A[1] : _dummy_a2;
            _b[k + 1][i] = have ? poly[i].b : _dummy_b;
        }
    }
}
void DecompConstraints::setParameters(const RealTimeData &data, const ModuleData &, int k) {
    setSolverParameterEgoDiscOffset(k, _solver->_params, data.robot_area[0].offset, 0);
    for (int i = 0; i < _max_constraints; i++) {
        setSolverParameterDecompA1(k, _solver->_params, k == 0 ? _dummy_a1 : _a1[k][i], i);
        setSolverParameterDecompA2(k, _solver->_params, k == 0 ? _dummy_a2 : _a2[k][i], i);
        setSolverParameterDecompB(k, _solver->_params, k == 0 ? _dummy_b : _b[k][i], i);
    }
}
#endif  // DECOMP
double remainingPlanningTime(const RealTimeData &data, double planning_time, double margin) {
    if (data.planning_start_time == std::chrono::system_clock::time_point{}) return 0.0;
    const std::chrono::duration<double> used_time = std::chrono::system_clock::now() - data.planning_start_time;
    const double t = planning_time - used_time.count() - margin;
    return t > 0.0 ? t : 1e-9;
}

#if defined(MPC_PLANNER_HAS_LINEAR) && defined(MPC_PLANNER_HAS_ELLIPSOID)
// ---- GuidanceConstraints (guidance_constraints.cpp:17-57, 264-434) ----
template <class SC>
GuidanceConstraintsT<SC>::LocalPlanner::LocalPlanner(int _id, bool _is_original_planner) : id(_id), is_original_planner(_is_original_planner) {
    local_solver = std::make_shared<Solver>(_id + 1);  // 1..n_solvers (0 is the regular solver)
    guidance_constraints = std::make_unique<LinearizedConstraints>(local_solver);
    safety_constraints = std::make_unique<SC>(local_solver);
    guidance_constraints->setTopologyConstraints();
}
template <class SC>
GuidanceConstraintsT<SC>::GuidanceConstraintsT(std::shared_ptr<Solver> solver, int n_paths)
    : ControllerModule(ModuleType::CONSTRAINT, solver, "guidance_constraints") {
    for (int i = 0; i < n_paths; i++) planners_.emplace_back(i);
    if (CONFIG().use_tmpcpp) planners_.emplace_back(n_paths, true);  // the non-guided planner, in parallel
}
template <class SC>
void GuidanceConstraintsT<SC>::initializeSolverWithGuidance(LocalPlanner &planner) {  // :390-414
    auto &solver = planner.local_solver;
    const GuidanceTrajectory &g = _guidance[planner.id];
    for (int k = 1; k < solver->N; k++) {  // the 0th velocity is the current velocity
        solver->setEgoPrediction(k, "x", g.positions[k][0]);
        solver->setEgoPrediction(k, "y", g.positions[k][1]);
        solver->setEgoPrediction(k, "psi", std::atan2(g.velocities[k][1], g.velocities[k][0]));
        solver->setEgoPrediction(k, "v", std::sqrt(g.velocities[k][0] * g.velocities[k][0] + g.velocities[k][1] * g.velocities[k][1]));
    }
}
template <class SC>
int GuidanceConstraintsT<SC>::FindBestPlanner() {  // :416-434
    double best_solution = 1e10;
    int best_index = -1;
    for (size_t i = 0; i < planners_.size(); i++) {
        auto &planner = planners_[i];
        if (planner.disabled) continue;
        if (planner.result.success && planner.result.objective < best_solution) {
            best_solution = planner.result.objective;
            best_index = (int)i;
        }
    }
    return best_index;
}
template <class SC>
void GuidanceConstraintsT<SC>::mapGuidanceTrajectoriesToPlanners() {  // :192-250
    // keep a guidance trajectory on the planner that optimised its homotopy class last time, so that this planner's own
    // previous solution (and multipliers) remain a valid initialisation
    std::vector<int> remaining_trajectories;
    for (auto &planner : planners_) planner.taken = planner.existing_guidance = false;
    _map_homotopy_class_to_planner.clear();
    for (int i = 0; i < (int)_guidance.size(); i++) {
        const int homotopy_class = _guidance[i].topology_class;
        bool planner_found = false;
        for (size_t p = 0; p < planners_.size(); p++) {
            if (planners_[p].result.guidance_ID == homotopy_class && !planners_[p].taken) {
                _map_homotopy_class_to_planner[i] = (int)p;
                planners_[p].taken = planners_[p].existing_guidance = true;
                planner_found = true;
                break;
            }
        }
        if (!planner_found) remaining_trajectories.push_back(i);
    }
    for (int i : remaining_trajectories)  // as written in the reference (no break): the last free planner wins, all get marked
        for (size_t p = 0; p < planners_.size(); p++)
            if (!planners_[p].taken) {
                _map_homotopy_class_to_planner[i] = (int)p;
                planners_[p].taken = true;
                planners_[p].existing_guidance = false;
            }
}

template <class SC>
std::vector<typename GuidanceConstraintsT<SC>::LocalPlanner *> GuidanceConstraintsT<SC>::prepare(State &state, const RealTimeData &data,
                                                                                                 ModuleData &module_data) {
    std::vector<LocalPlanner *> active;
    mapGuidanceTrajectoriesToPlanners();  // :108 (uses the previous control step's results, before they are reset below)
    // what the reference does inside `#pragma omp parallel for` per planner up to solver->solve() (:279-337)
    for (auto &planner : planners_) {
        planner.result = SolverResult();
        planner.disabled = false;
        if (planner.id >= (int)_guidance.size() && !planner.is_original_planner) {  // only the solvers that are needed
            planner.disabled = true;
            continue;
        }
        auto &solver = planner.local_solver;
        *solver = *_solver;  // copy the main solver (parameters only)
        if (planner.is_original_planner || !CONFIG().enable_constraints) {
            planner.guidance_constraints->update(state, empty_data_, module_data);
            planner.safety_constraints->update(state, data, module_data);
        } else {
            if (CONFIG().warmstart_with_mpc_solution && planner.existing_guidance)  // :310-313
                planner.local_solver->initializeWarmstart(state, CONFIG().shift_previous_solution_forward);
            else
                initializeSolverWithGuidance(planner);
            planner.guidance_constraints->update(state, data, module_data);  // linearisation of the topology constraints
            planner.safety_constraints->update(state, data, module_data);    // collision avoidance constraints
        }
        for (int k = 0; k < _solver->N; k++) {
            planner.guidance_constraints->setParameters(planner.is_original_planner ? empty_data_ : data, module_data, k);
            planner.safety_constraints->setParameters(data, module_data, k);
        }
        // timeout: planning time - used time - time necessary afterwards (:331-333)
        planner.local_solver->_params.solver_timeout = remainingPlanningTime(data, 1. / CONFIG().control_frequency, 0.006);
        planner.local_solver->loadWarmstart();
        active.push_back(&planner);
    }
    return active;
}

template <class SC>
int GuidanceConstraintsT<SC>::optimize(State &state, const RealTimeData &data, ModuleData &module_data) {  // :264-388
    PROFILE_FUNCTION();  // :266
    if (!CONFIG().use_tmpcpp && _guidance.empty()) return 0;
    std::vector<LocalPlanner *> active = prepare(state, data, module_data);
    std::vector<Solver *> batch;
    for (LocalPlanner *p : active) batch.push_back(p->local_solver.get());
    // ... and the P solves as ONE batched launch
    std::vector<int> exit_codes;
    {
        PROFILE_SCOPE("Guidance Constraints: Parallel Optimization");  // :282 (here: one batched launch)
        Solver::solveBatch(batch, exit_codes);
    }
    PROFILE_SCOPE("Decision");  // :366
    for (size_t i = 0; i < active.size(); i++) {
        LocalPlanner &planner = *active[i];
        planner.result.exit_code = exit_codes[i];
        planner.result.success = planner.result.exit_code == 1;
        planner.result.objective = planner.local_solver->_info.pobj;
        if (planner.is_original_planner) {
            planner.result.guidance_ID = 2 * (int)planners_.size();
        } else {
            planner.result.guidance_ID = _guidance[planner.id].topology_class;
            if (_guidance[planner.id].previously_selected) planner.result.objective *= CONFIG().selection_weight_consistency;
        }
    }
    best_planner_index_ = FindBestPlanner();
    if (best_planner_index_ == -1) return planners_[0].result.exit_code;
    auto &best_planner = planners_[best_planner_index_];
    _solver->_output = best_planner.local_solver->_output;  // load the solution into the main solver (:382-384)
    _solver->_info = best_planner.local_solver->_info;
    _solver->_params = best_planner.local_solver->_params;
    return best_planner.result.exit_code;
}
template class GuidanceConstraintsT<EllipsoidConstraints>;
#endif  // LINEAR && ELLIPSOID

#if defined(MPC_PLANNER_HAS_SCENARIO)
// ---- HalfspaceScenarioModule: stand-in for the external ScenarioModule::ScenarioModule ----
void HalfspaceScenarioModule::initialize(std::shared_ptr<Solver> solver) {
    _solver = solver;
    _n_constraints = 0;
    while (_solver->hasParameter("disc_0_scenario_constraint_" + std::to_string(_n_constraints) + "_a1")) _n_constraints++;
}
void HalfspaceScenarioModule::update(const RealTimeData &, const ModuleData &) {}
void HalfspaceScenarioModule::setParameters(const RealTimeData &data, int k) {
    // dummies (1, 0, x0 + 100), as the linear modules of the reference use for k = 0 and unused rows
    // (linearized_constraints.cpp:155-166): always satisfied, never active
    const double dummy_b = _solver->_params.xinit[0] + 100.0;
    if (_solver->hasParameter("ego_disc_0_offset")) _solver->setParameter(k, "ego_disc_0_offset", data.robot_area.empty() ? 0.0 : data.robot_area[0].offset);
    for (int i = 0; i < _n_constraints; i++) {
        const std::string base = "disc_0_scenario_constraint_" + std::to_string(i);
        const bool have = k >= 1 && k < (int)_halfspaces.size() && i < (int)_halfspaces[k].size();
        _solver->setParameter(k, base + "_a1", have ? _halfspaces[k][i].A[0] : 1.0);
        _solver->setParameter(k, base + "_a2", have ? _halfspaces[k][i].A[1] : 0.0);
        _solver->setParameter(k, base + "_b", have ? _halfspaces[k][i].b : dummy_b);
    }
}

// ---- ScenarioConstraints (scenario_constraints.cpp:17-108) ----
template <class SM>
ScenarioConstraintsT<SM>::ScenarioSolver::ScenarioSolver(int id) {  // :17-21
    solver = std::make_shared<Solver>(id);
    scenario_module.initialize(solver);
}
template <class SM>
ScenarioConstraintsT<SM>::ScenarioConstraintsT(std::shared_ptr<Solver> solver, int parallel_solvers)
    : ControllerModule(ModuleType::CONSTRAINT, solver, "scenario_constraints") {  // :23-36
    for (int i = 0; i < parallel_solvers; i++) _scenario_solvers.emplace_back(std::make_unique<ScenarioSolver>(i));
}
template <class SM>
void ScenarioConstraintsT<SM>::update(State &, const RealTimeData &data, ModuleData &module_data) {  // :38-49
    for (auto &solver : _scenario_solvers) {
        *solver->solver = *_solver;  // copy the main solver, including its initial guess
        solver->scenario_module.update(data, module_data);
    }
}
template <class SM>
int ScenarioConstraintsT<SM>::optimize(State &, const RealTimeData &data, ModuleData &) {  // :58-108
    std::vector<Solver *> batch;
    for (auto &solver : _scenario_solvers) {
        // planning timeout (:69-71; margin 0.008 here, 0.006 in the other two places)
        const double timeout = remainingPlanningTime(data, 1. / CONFIG().control_frequency, 0.008);
        *solver->solver = *_solver;  // copy solver parameters and initial guess (:74)
        solver->solver->_params.solver_timeout = timeout;
        for (int k = 0; k < _solver->N; k++) solver->scenario_module.setParameters(data, k);  // :77-80
        solver->solver->loadWarmstart();                                                     // :82
        batch.push_back(solver->solver.get());
    }
    // `solver->exit_code = solver->scenario_module.optimize(data)` (:84) for all solvers: ONE batched launch, then the
    // module's post-solve part
    std::vector<int> exit_codes;
    Solver::solveBatch(batch, exit_codes);
    for (size_t i = 0; i < _scenario_solvers.size(); i++) _scenario_solvers[i]->exit_code = _scenario_solvers[i]->scenario_module.completeOptimize(exit_codes[i]);
    // retrieve the lowest-cost solution (:88-99)
    double lowest_cost = 1e9;
    _best_solver = -1;
    for (size_t i = 0; i < _scenario_solvers.size(); i++) {
        auto &solver = _scenario_solvers[i];
        if (solver->exit_code == 1 && solver->solver->_info.pobj < lowest_cost) {
            lowest_cost = solver->solver->_info.pobj;
            _best_solver = (int)i;
        }
    }
    if (_best_solver < 0) return _scenario_solvers.front()->exit_code;  // no feasible solution (:100-101)
    auto &best = _scenario_solvers[_best_solver];
    _solver->_output = best->solver->_output;  // load the solution into the main solver (:103-105)
    _solver->_info = best->solver->_info;
    _solver->_params = best->solver->_params;
    return best->exit_code;
}
template class ScenarioConstraintsT<HalfspaceScenarioModule>;
#endif  // SCENARIO
}  // namespace MPCPlanner
