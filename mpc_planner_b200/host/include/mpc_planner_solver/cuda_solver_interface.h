// cuda_solver_interface.h -- `MPCPlanner::Solver` backed by the B200 batched SQP library.
//
// Drop-in for mpc_planner_solver/include/mpc_planner_solver/acados_solver_interface.h:93-222: same
// public members and semantics (parameters/warm start/output plumbing of
// src/acados_solver_interface.cpp:207-389), with `solve()` forwarding to the C ABI
// (include/mpcgpu.h) instead of the generated acados solver.  Selected the way the reference
// selects acados vs FORCES (solver_interface.h:4-12): the generated header defines
// MPC_PLANNER_CUDA_SOLVER.  No yaml-cpp / Eigen / ros_tools: the maps come from the generated header.
#ifndef MPC_PLANNER_CUDA_SOLVER_INTERFACE_H
#define MPC_PLANNER_CUDA_SOLVER_INTERFACE_H

#include <array>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "mpcgpu.h"
#include "mpc_planner_solver/state.h"

namespace MPCPlanner {

// AcadosParameters (acados_solver_interface.h:51-91)
struct CudaParameters {
    std::vector<double> xinit;           // [nx]
    std::vector<double> x0;              // [(nu+nx)*(N+1)]  warm start: [u0, x0 | u1 x1 | ... | uN xN]
    std::vector<double> all_parameters;  // [npar*N]
    double solver_timeout{0.};
    double *getU0() { return x0.data(); }
    CudaParameters();
};

class Solver {
  public:
    struct Info {  // AcadosInfo (:96-125)
        double min_time{1e12}, kkt_norm_inf{0.}, elapsed_time{0.};
        int sqp_iter{0};
        double nlp_res{0.}, solvetime{0.};
        int qp_status{0};
        double pobj{0.};
    };
    struct Output {  // AcadosOutput (:127-147)
        std::vector<double> xtraj, utraj;
        Output();
    };

    int _solver_id;
    CudaParameters _params;
    Info _info;
    Output _output;

    int N;
    unsigned int nu, nx, nvar, npar;
    double dt;
    const std::map<std::string, int> &_parameter_map;
    const std::map<std::string, ModelMapEntry> &_model_map;
    int _num_iterations;

    explicit Solver(int solver_id = 0);
    ~Solver();
    Solver(const Solver &) = delete;
    /** Copy data from another solver: parameters only (acados_solver_interface.cpp:67-77). */
    Solver &operator=(const Solver &rhs);

    void reset();
    int solve();
    // one-iteration-at-a-time interface (:121-204); the CUDA backend runs all RTI iterations in one
    // launch, so solveOneIteration() performs the whole solve on its first call
    void initializeOneIteration();
    int solveOneIteration();
    int completeOneIteration();

    bool hasParameter(const std::string &parameter) const;
    void setParameter(int k, const std::string &parameter, double value);
    double getParameter(int k, const std::string &parameter) const;

    void setXinit(const std::string &state_name, double value);
    void setXinit(const State &state);

    void setEgoPrediction(unsigned int k, const std::string &var_name, double value);
    double getEgoPrediction(unsigned int k, const std::string &var_name) const;
    void setEgoPredictionPosition(unsigned int k, const std::array<double, 2> &value);
    std::array<double, 2> getEgoPredictionPosition(unsigned int k) const;

    void loadWarmstart();
    void initializeWarmstart(const State &state, bool shift_previous_solution_forward);
    void initializeWithState(const State &initial_state);
    void initializeWithBraking(const State &initial_state);

    double getOutput(int k, const std::string &state_name) const;
    std::string explainExitFlag(int exitflag) const;
    void printIfBoundLimited() const;

    /** Batch entry point used by the GPU replacement of GuidanceConstraints::optimize
     *  (mpc_planner_modules/src/guidance_constraints.cpp:264-388): solves all `solvers` in ONE launch
     *  and returns the index of the best successful one (FindBestPlanner :416-434), -1 if none.
     *  exit_codes receives each solver's exit code. */
    static int solveBatch(const std::vector<Solver *> &solvers, std::vector<int> &exit_codes);

    /** The NLP multipliers of this solver's previous solve ([pi | multipliers of h], mpcgpu_state_size doubles).
     *  They belong to the Solver object like acados' nlp_out does: kept by operator= (:67-77), cleared by reset()
     *  (:79-84), fed to the first exact Hessian of the next solve.  Set persistent_multipliers = false for a solver
     *  that starts every solve from zero multipliers. */
    static bool persistent_multipliers;
    const std::vector<double> &nlpMultipliers() const { return _nlp_multipliers; }

  private:
    std::vector<double> _nlp_multipliers;
    mpcgpu_problem_desc _desc;
    mpcgpu_handle _handle{nullptr};  // created lazily: constructing a Solver needs no GPU
    std::vector<double> _loaded_x0;  // snapshot taken by loadWarmstart()
    int _exit_code_one_iter{-1};
    bool _iteration_done{false};
    mpcgpu_handle handle(int max_batch);
    void storeResult(const double *xtraj, const double *utraj, const mpcgpu_info &info);
    int _handle_batch{0};
};
}  // namespace MPCPlanner
#endif
