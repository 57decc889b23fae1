// modules.h -- online (C++) half of the reference's modules for the CUDA backend: the functions that
// pack `Solver::_params.all_parameters` before a solve, and the batched T-MPC++ loop.
//
// Mirrors (same names, argument meaning and packing semantics; Eigen / yaml-cpp / ros_tools free):
//   types              mpc_planner_types/include/mpc_planner_types/{data_types,realtime_data,module_data}.h
//   ControllerModule   mpc_planner_modules/include/mpc_planner_modules/controller_module.h:35-137
//   MPCBaseModule      src/mpc_base.cpp:23-35
//   Contouring         src/contouring.cpp:28-124     (spline container instead of RosTools::Spline2D)
//   EllipsoidConstraints  src/ellipsoid_constraints.cpp:24-90
//   GaussianConstraints   src/gaussian_constraints.cpp:23-79
//   LinearizedConstraints src/linearized_constraints.cpp:43-189
//   DecompConstraints::setParameters  src/decomp_constraints.cpp:150-189 (polyhedra are an input: DecompUtil is external)
//   GuidanceConstraints   src/guidance_constraints.cpp:17-57, 264-434 (guidance trajectories are an input:
//                         guidance_planner is external); the OpenMP loop over planners becomes ONE batched launch.
#ifndef MPC_PLANNER_CUDA_MODULES_H
#define MPC_PLANNER_CUDA_MODULES_H

#include <array>
#include <chrono>
#include <map>
#include <unordered_map>
#include <memory>
#include <string>
#include <vector>

#include "mpc_planner_generated.h"
#include <mpc_planner_solver/solver_interface.h>

#define EXIT_CODE_NOT_OPTIMIZED_YET -999

namespace MPCPlanner {

typedef std::array<double, 2> Vec2;

// ---- mpc_planner_types ----
struct Disc {
    double offset, radius;
    Disc(double offset_, double radius_) : offset(offset_), radius(radius_) {}
    Vec2 getPosition(const Vec2 &robot_position, double angle) const;
};
struct Halfspace {
    Vec2 A;
    double b;
};
typedef std::vector<Halfspace> StaticObstacle;  // for all k, a halfspace
enum class PredictionType { DETERMINISTIC = 0, GAUSSIAN, NONGAUSSIAN, NONE };
struct PredictionStep {
    Vec2 position;
    double angle, major_radius, minor_radius;
};
typedef std::vector<PredictionStep> Mode;
struct Prediction {
    PredictionType type{PredictionType::NONE};
    std::vector<Mode> modes;
    std::vector<double> probabilities;
};
enum class ObstacleType { STATIC = 0, DYNAMIC };
struct DynamicObstacle {
    int index;
    Vec2 position;
    double angle, radius;
    ObstacleType type{ObstacleType::DYNAMIC};
    Prediction prediction;
};
struct RealTimeData {
    std::vector<Disc> robot_area;
    std::vector<DynamicObstacle> dynamic_obstacles;
    // set by the caller at the start of a control step (data_types.h; the ROS wrappers do it in their loop callback); the
    // solver timeouts are derived from it (planner.cpp:117-118, guidance_constraints.cpp:331-333).  Left at its default
    // (epoch) the timeouts are not touched.
    std::chrono::system_clock::time_point planning_start_time{};
};
// remaining time budget of this control step minus `margin`, as planner.cpp:117-118 computes it; a used-up budget becomes a
// tiny positive one (the reference then stops after its first RTI iteration), an unset start time gives 0 (= unbounded)
double remainingPlanningTime(const RealTimeData &data, double planning_time, double margin);
// the reference path as the contouring module consumes it: cubic segments in arc length
// (what RosTools::Spline2D::getParameters / getSegmentStart return, contouring.cpp:100-123)
struct PathSegment {
    double ax, bx, cx, dx, ay, by, cy, dy, start;
};
struct Spline2D {
    std::vector<PathSegment> segments;
    double length{0.};
    Vec2 getPoint(double s) const;
    void findClosestPoint(const Vec2 &pos, int &segment_out, double &s_out) const;
    int numSegments() const { return (int)segments.size(); }
};
struct ModuleData {
    std::vector<StaticObstacle> static_obstacles;
    std::shared_ptr<Spline2D> path;
    std::vector<std::vector<Halfspace>> decomp_polyhedra;  // per stage k: halfspaces of the free-space polyhedron
    int current_path_segment{-1};
};
// CONFIG stand-in: the handful of settings.yaml keys the modules read
struct Config {
    std::map<std::string, double> weights;  // settings.yaml:75-88
    double robot_radius{0.325}, obstacle_radius{0.4}, risk{0.05};
    int n_discs{1}, max_obstacles{8}, num_segments{5}, add_halfspaces{0}, decomp_max_constraints{12};
    bool use_tmpcpp{true}, enable_constraints{true};
    bool warmstart_with_mpc_solution{false};   // t-mpc/warmstart_with_mpc_solution (settings.yaml:68)
    bool shift_previous_solution_forward{true};
    double control_frequency{20.};             // settings.yaml:7
    double selection_weight_consistency{1.0};
    int n_paths{4};  // guidance_planner n_paths: guided planners (T-MPC++ adds the non-guided one)
};
Config &CONFIG();

void propagatePredictionUncertainty(Prediction &prediction, double dt, int N);  // mpc_planner/src/data_preparation.cpp:170-186

// ---- ControllerModule ----
enum class ModuleType { OBJECTIVE = 0, CONSTRAINT, UNDEFINED };
class ControllerModule {
  public:
    ControllerModule(ModuleType module_type, std::shared_ptr<Solver> solver, const std::string &module_name)
        : type(module_type), _solver(solver), _name(module_name) {}
    virtual ~ControllerModule() {}
    virtual void update(State &, const RealTimeData &, ModuleData &) {}
    virtual void setParameters(const RealTimeData &, const ModuleData &, int) {}
    virtual bool isDataReady(const RealTimeData &, std::string &) { return true; }
    virtual int optimize(State &, const RealTimeData &, ModuleData &) { return EXIT_CODE_NOT_OPTIMIZED_YET; }
    virtual void reset() {}
    virtual bool isObjectiveReached(const State &, const RealTimeData &) { return true; }
    ModuleType type;

  protected:
    std::shared_ptr<Solver> _solver;
    std::string _name;
};

class MPCBaseModule : public ControllerModule {
  public:
    explicit MPCBaseModule(std::shared_ptr<Solver> solver);
    void setParameters(const RealTimeData &data, const ModuleData &module_data, int k) override;

  private:
    std::vector<std::string> _weight_names;
};

class Contouring : public ControllerModule {
  public:
    explicit Contouring(std::shared_ptr<Solver> solver);
    void update(State &state, const RealTimeData &data, ModuleData &module_data) override;
    void setParameters(const RealTimeData &data, const ModuleData &module_data, int k) override;

  private:
    bool isObjectiveReached(const State &state, const RealTimeData &data) override;  // contouring.cpp:167-179
    void setSplineParameters(int k);
    std::shared_ptr<Spline2D> _spline;
    int _closest_segment{0}, _n_segments;
};

class EllipsoidConstraints : public ControllerModule {
  public:
    explicit EllipsoidConstraints(std::shared_ptr<Solver> solver);
    void update(State &state, const RealTimeData &data, ModuleData &module_data) override;
    void setParameters(const RealTimeData &data, const ModuleData &module_data, int k) override;

  private:
    double _dummy_x{0.}, _dummy_y{0.}, _robot_radius, _risk;
    int _n_discs;
};

class GaussianConstraints : public ControllerModule {
  public:
    explicit GaussianConstraints(std::shared_ptr<Solver> solver);
    void update(State &state, const RealTimeData &data, ModuleData &module_data) override;
    void setParameters(const RealTimeData &data, const ModuleData &module_data, int k) override;

  private:
    double _dummy_x{0.}, _dummy_y{0.};
};

class LinearizedConstraints : public ControllerModule {
  public:
    explicit LinearizedConstraints(std::shared_ptr<Solver> solver);
    void setTopologyConstraints();
    void update(State &state, const RealTimeData &data, ModuleData &module_data) override;
    void setParameters(const RealTimeData &data, const ModuleData &module_data, int k) override;

  private:
    void projectToSafety(const std::vector<DynamicObstacle> &copied_obstacles, int k, Vec2 &pos);
    std::vector<std::vector<std::vector<double>>> _a1, _a2, _b;  // [disc][k][constraint]
    int _n_discs, _n_other_halfspaces, _max_obstacles, _num_obstacles{0};
    bool _use_guidance{false};
    double _dummy_a1{1.}, _dummy_a2{0.}, _dummy_b{0.};
};

class DecompConstraints : public ControllerModule {
  public:
    explicit DecompConstraints(std::shared_ptr<Solver> solver);
    void update(State &state, const RealTimeData &data, ModuleData &module_data) override;
    void setParameters(const RealTimeData &data, const ModuleData &module_data, int k) override;

  private:
    std::vector<std::vector<double>> _a1, _a2, _b;  // [k][constraint]
    int _max_constraints;
    double _dummy_a1{1.}, _dummy_a2{0.}, _dummy_b{0.};
};

// One guidance trajectory as the guidance planner hands it over (positions and velocities at t = k dt)
struct GuidanceTrajectory {
    std::vector<Vec2> positions, velocities;  // index k = 0..N
    int topology_class{0};
    bool previously_selected{false};
};

template <class SafetyConstraints>
class GuidanceConstraintsT : public ControllerModule {
  public:
    struct SolverResult {
        int exit_code{0};
        double objective{1e10};
        bool success{false};
        int guidance_ID{0};
    };
    struct LocalPlanner {
        int id;
        bool is_original_planner, disabled{true};
        bool taken{false}, existing_guidance{false};  // guidance_constraints.h:97-98
        std::shared_ptr<Solver> local_solver;
        std::unique_ptr<LinearizedConstraints> guidance_constraints;
        std::unique_ptr<SafetyConstraints> safety_constraints;
        SolverResult result;
        LocalPlanner(int _id, bool _is_original_planner = false);
    };
    GuidanceConstraintsT(std::shared_ptr<Solver> solver, int n_paths);
    /** the guidance planner's output for this control step (replaces GlobalGuidance::Update) */
    void setGuidanceTrajectories(const std::vector<GuidanceTrajectory> &trajectories) { _guidance = trajectories; }
    /** everything the reference does per planner before `solver->solve()` (copy of the main solver's parameters,
     *  guidance warm start, constraint construction, parameter packing, loadWarmstart); returns the active planners */
    std::vector<LocalPlanner *> prepare(State &state, const RealTimeData &data, ModuleData &module_data);
    int optimize(State &state, const RealTimeData &data, ModuleData &module_data) override;
    int best_planner_index() const { return best_planner_index_; }
    const std::unordered_map<int, int> &homotopyClassToPlanner() const { return _map_homotopy_class_to_planner; }
    std::vector<LocalPlanner> planners_;

  private:
    void initializeSolverWithGuidance(LocalPlanner &planner);
    void mapGuidanceTrajectoriesToPlanners();  // :192-250
    std::unordered_map<int, int> _map_homotopy_class_to_planner;
    int FindBestPlanner();
    std::vector<GuidanceTrajectory> _guidance;
    RealTimeData empty_data_;
    int best_planner_index_{-1};
};
typedef GuidanceConstraintsT<EllipsoidConstraints> GuidanceConstraints;

// ---- Safe Horizon MPC (mpc_planner_modules/src/scenario_constraints.cpp:17-108) ----
// The reference's ScenarioConstraints runs `parallel_solvers` (4) local solvers, each with its own sample set, through
// ScenarioModule::ScenarioModule of the un-vendored oscardegroot/scenario_module (sampling, halfspace construction, the
// SH-MPC support/removal loop).  ScenarioConstraintsT keeps the reference's control flow around that module and turns the
// `#pragma omp parallel for` over the solvers into ONE batched launch; what it needs from the module is the surface the
// reference uses (`initialize`, `update`, `setParameters(data, k)`) plus `completeOptimize(exit_code)`, the part of the
// module's `optimize` that follows its solve.
// HalfspaceScenarioModule is the stand-in used here: the scenario halfspaces of every stage are handed in from outside
// (as DecompConstraints takes its polyhedra) and written to `disc_0_scenario_constraint_<i>_{a1,a2,b}`
// (scenario_constraints.py:26,64-94: 24 rows, always with the slack state).
class HalfspaceScenarioModule {
  public:
    void initialize(std::shared_ptr<Solver> solver);
    /** per stage k = 0..N-1 the halfspaces a1 x + a2 y <= b of this solver's sample set (more than the solver's rows: cut) */
    void setHalfspaces(const std::vector<std::vector<Halfspace>> &per_stage) { _halfspaces = per_stage; }
    void update(const RealTimeData &data, const ModuleData &module_data);
    void setParameters(const RealTimeData &data, int k);
    int completeOptimize(int exit_code) { return exit_code; }

  private:
    std::shared_ptr<Solver> _solver;
    std::vector<std::vector<Halfspace>> _halfspaces;
    int _n_constraints{0};
};

template <class ScenarioModule>
class ScenarioConstraintsT : public ControllerModule {
  public:
    struct ScenarioSolver {
        std::shared_ptr<Solver> solver;
        ScenarioModule scenario_module;
        int exit_code{0};
        explicit ScenarioSolver(int id);
    };
    ScenarioConstraintsT(std::shared_ptr<Solver> solver, int parallel_solvers = 4);  // scenario_constraints/parallel_solvers
    void update(State &state, const RealTimeData &data, ModuleData &module_data) override;
    void setParameters(const RealTimeData &, const ModuleData &, int) override {}  // :51-56 (the local solvers carry the parameters)
    int optimize(State &state, const RealTimeData &data, ModuleData &module_data) override;
    int best_solver_index() const { return _best_solver; }
    std::vector<std::unique_ptr<ScenarioSolver>> _scenario_solvers;

  private:
    int _best_solver{-1};
};
typedef ScenarioConstraintsT<HalfspaceScenarioModule> ScenarioConstraints;
}  // namespace MPCPlanner
#endif
