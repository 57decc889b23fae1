// cuda_solver_interface.cpp -- see cuda_solver_interface.h.  Each function cites the reference
// member it mirrors (mpc_planner_solver/src/acados_solver_interface.cpp).
#include "mpc_planner_generated.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace MPCPlanner {

// ---- State (src/state.cpp:7-43) ----
State::State() { initialize(); }
void State::initialize() {
    _state.assign(SOLVER_NX, 0.0);
    _nu = SOLVER_NU;
}
double State::get(const std::string &var_name) const { return _state[generatedModelMap().at(var_name).index - _nu]; }
void State::set(const std::string &var_name, double value) { _state[generatedModelMap().at(var_name).index - _nu] = value; }
void State::print() const {
    for (const auto &kv : generatedModelMap())
        if (kv.second.type == 'x') std::printf("%s: %f\n", kv.first.c_str(), get(kv.first));
}

// ---- parameters / output PODs ----
CudaParameters::CudaParameters()
    : xinit(SOLVER_NX, 0.0), x0((SOLVER_NU + SOLVER_NX) * (SOLVER_N + 1), 0.0), all_parameters(SOLVER_NP * SOLVER_N, 0.0) {}
Solver::Output::Output() : xtraj(SOLVER_NX * (SOLVER_N + 1), 0.0), utraj(SOLVER_NU * SOLVER_N, 0.0) {}

// ---- Solver ----
Solver::Solver(int solver_id)  // :9-49
    : _solver_id(solver_id), _parameter_map(generatedParameterMap()), _model_map(generatedModelMap()) {
    _desc = generatedProblemDesc();
    N = SOLVER_N;
    nu = SOLVER_NU;
    nx = SOLVER_NX;
    nvar = SOLVER_NU + SOLVER_NX;
    npar = SOLVER_NP;
    dt = SOLVER_DT;
    _num_iterations = SOLVER_ITERATIONS;
    if (!mpcgpu_problem_supported(&_desc)) {  // the reference exits when the generated solver cannot be created (:35-39)
        std::printf("mpcgpu: no kernel specialisation for this problem. Exiting.\n");
        std::exit(1);
    }
    reset();
}

bool Solver::persistent_multipliers = true;

Solver::~Solver() {
    if (_handle) mpcgpu_destroy(_handle);
}

Solver &Solver::operator=(const Solver &rhs) {  // :67-77: parameters only; output / solver memory are kept
    _params = rhs._params;
    return *this;
}

void Solver::reset() {  // :79-84
    _params = CudaParameters();
    _info = Info();
    _output = Output();
    _loaded_x0 = _params.x0;
    _nlp_multipliers.assign((size_t)(_desc.N + 1) * _desc.nx + (size_t)_desc.N * _desc.nh, 0.0);
}

mpcgpu_handle Solver::handle(int max_batch) {
    if (_handle && _handle_batch >= max_batch) return _handle;
    if (_handle) mpcgpu_destroy(_handle);
    _handle = nullptr;
    int rc = mpcgpu_create(&_desc, 0, max_batch, &_handle);
    if (rc != 0) throw std::runtime_error(std::string("mpcgpu_create: ") + mpcgpu_strerror(rc));
    _handle_batch = max_batch;
    rc = mpcgpu_set_stateful(_handle, 1);
    if (rc != 0) throw std::runtime_error(std::string("mpcgpu_set_stateful: ") + mpcgpu_strerror(rc));
    return _handle;
}

void Solver::storeResult(const double *xtraj, const double *utraj, const mpcgpu_info &info) {
    std::copy(xtraj, xtraj + _output.xtraj.size(), _output.xtraj.begin());
    std::copy(utraj, utraj + _output.utraj.size(), _output.utraj.begin());
    _info.pobj = info.pobj;
    _info.qp_status = info.qp_status;
    _info.sqp_iter = info.sqp_iter;
    _info.nlp_res = info.res_eq;
    _exit_code_one_iter = info.exit_code;
}

int Solver::solve() {  // :86-119 -> one launch running all RTI iterations (the iteration loop lives on the device)
    initializeOneIteration();
    solveOneIteration();
    return completeOneIteration();
}

void Solver::initializeOneIteration() {  // :121-143
    _info = Info();
    _iteration_done = false;
}

namespace {
// Pinned staging of one batched launch, owned by the lead solver's handle and reused from control step to control step
// (round 1 heap-allocated five vectors and made two synchronous state round trips per step).
struct BatchStaging {
    void *base = nullptr;
    size_t bytes = 0;
    double *xinit = nullptr, *x0 = nullptr, *par = nullptr, *mult = nullptr, *xtraj = nullptr, *utraj = nullptr;
    mpcgpu_info *info = nullptr;
    void ensure(size_t P, size_t nxi, size_t nx0, size_t np, size_t ns, size_t nxt, size_t nut) {
        const size_t doubles = P * (nxi + nx0 + np + ns + nxt + nut);
        const size_t need = doubles * sizeof(double) + P * sizeof(mpcgpu_info);
        if (need > bytes) {
            if (base) mpcgpu_host_free(base);
            base = nullptr;
            if (mpcgpu_host_alloc(&base, need) != 0) throw std::runtime_error("mpcgpu_host_alloc failed");
            bytes = need;
        }
        double *p = static_cast<double *>(base);
        xinit = p; p += P * nxi;
        x0 = p; p += P * nx0;
        par = p; p += P * np;
        mult = p; p += P * ns;
        xtraj = p; p += P * nxt;
        utraj = p; p += P * nut;
        info = reinterpret_cast<mpcgpu_info *>(p);
    }
    ~BatchStaging() {
        if (base) mpcgpu_host_free(base);
    }
};
BatchStaging &staging() {
    static thread_local BatchStaging s;  // a Solver is not re-entrant (SURVEY 8b "Threading"); one staging area per host thread
    return s;
}
}  // namespace

int Solver::solveOneIteration() {  // :145-160
    if (_iteration_done) return _exit_code_one_iter == 1 ? 0 : _exit_code_one_iter;
    std::vector<Solver *> one{this};
    std::vector<int> codes;
    solveBatch(one, codes);
    return _exit_code_one_iter == 1 ? 0 : _exit_code_one_iter;
}

int Solver::completeOneIteration() { return _exit_code_one_iter; }  // :162-204 (exit-code mapping happens on the device)

int Solver::solveBatch(const std::vector<Solver *> &solvers, std::vector<int> &exit_codes) {
    const int P = (int)solvers.size();
    exit_codes.assign(P, 0);
    if (P == 0) return -1;
    Solver *lead = solvers[0];
    const size_t nxi = lead->_params.xinit.size(), nx0 = lead->_params.x0.size(), np = lead->_params.all_parameters.size();
    const size_t nxt = lead->_output.xtraj.size(), nut = lead->_output.utraj.size(), ns = lead->_nlp_multipliers.size();
    mpcgpu_handle h = lead->handle(P);
    BatchStaging &st = staging();
    st.ensure(P, nxi, nx0, np, ns, nxt, nut);
    for (int i = 0; i < P; i++) {
        std::copy(solvers[i]->_params.xinit.begin(), solvers[i]->_params.xinit.end(), st.xinit + i * nxi);
        std::copy(solvers[i]->_loaded_x0.begin(), solvers[i]->_loaded_x0.end(), st.x0 + i * nx0);
        std::copy(solvers[i]->_params.all_parameters.begin(), solvers[i]->_params.all_parameters.end(), st.par + i * np);
        // each Solver owns its multipliers (slot i of the launch <- solvers[i]'s, and back): they travel with the call
        if (persistent_multipliers) std::copy(solvers[i]->_nlp_multipliers.begin(), solvers[i]->_nlp_multipliers.end(), st.mult + i * ns);
        else std::fill(st.mult + i * ns, st.mult + (i + 1) * ns, 0.0);
    }
    int best = -1;
    int rc = 0;
    {   // guidance_constraints.cpp:333 gives every local planner the same budget; one launch: the smallest, unbounded if any is unset
        double tmo = solvers[0]->_params.solver_timeout;
        for (const Solver *sv : solvers) tmo = (sv->_params.solver_timeout <= 0.0 || tmo <= 0.0) ? 0.0 : std::min(tmo, sv->_params.solver_timeout);
        rc = mpcgpu_set_timeout(h, tmo);
    }
    if (rc == 0) rc = mpcgpu_solve_host_stateful(h, P, st.xinit, st.x0, st.par, st.mult, st.xtraj, st.utraj, st.info, &best);
    if (rc != 0) throw std::runtime_error(std::string("mpcgpu_solve_host_stateful: ") + mpcgpu_strerror(rc));
    for (int i = 0; i < P; i++) {
        std::copy(st.mult + i * ns, st.mult + (i + 1) * ns, solvers[i]->_nlp_multipliers.begin());
        solvers[i]->_info = Info();
        solvers[i]->storeResult(st.xtraj + i * nxt, st.utraj + i * nut, st.info[i]);
        solvers[i]->_iteration_done = true;
        exit_codes[i] = st.info[i].exit_code;
    }
    return best;
}

// PARAMETERS (:207-225)
bool Solver::hasParameter(const std::string &parameter) const { return _parameter_map.count(parameter) > 0; }
void Solver::setParameter(int k, const std::string &parameter, double value) {
    _params.all_parameters[k * npar + _parameter_map.at(parameter)] = value;
}
double Solver::getParameter(int k, const std::string &parameter) const {
    return _params.all_parameters[k * npar + _parameter_map.at(parameter)];
}

// XINIT (:229-246)
void Solver::setXinit(const std::string &state_name, double value) { _params.xinit[_model_map.at(state_name).index - nu] = value; }
void Solver::setXinit(const State &state) {
    for (const auto &kv : _model_map)
        if (kv.second.type == 'x') setXinit(kv.first, state.get(kv.first));
}

// WARMSTART (:250-376)
void Solver::setEgoPrediction(unsigned int k, const std::string &var_name, double value) {
    _params.x0[k * nvar + _model_map.at(var_name).index] = value;
}
double Solver::getEgoPrediction(unsigned int k, const std::string &var_name) const {
    return _params.x0[k * nvar + _model_map.at(var_name).index];
}
void Solver::setEgoPredictionPosition(unsigned int k, const std::array<double, 2> &value) {
    setEgoPrediction(k, "x", value[0]);
    setEgoPrediction(k, "y", value[1]);
}
std::array<double, 2> Solver::getEgoPredictionPosition(unsigned int k) const {
    return {getEgoPrediction(k, "x"), getEgoPrediction(k, "y")};
}

void Solver::loadWarmstart() { _loaded_x0 = _params.x0; }  // :274-284: x (k=0..N) and u (k<N) into the solver

void Solver::initializeWithState(const State &initial_state) {  // :286-301
    for (int k = 0; k <= N; k++)
        for (const auto &kv : _model_map)
            setEgoPrediction(k, kv.first, kv.second.type == 'x' ? initial_state.get(kv.first) : 0.0);
}

void Solver::initializeWithBraking(const State &initial_state) {  // :303-342
    initializeWithState(initial_state);
    const double deceleration = std::abs(SOLVER_DECELERATION_AT_INFEASIBLE);
    double x = initial_state.get("x"), y = initial_state.get("y"), psi = initial_state.get("psi"), v = initial_state.get("v");
    double spline = initial_state.get("spline"), a = -deceleration;
    setEgoPrediction(0, "x", x);
    setEgoPrediction(0, "y", y);
    setEgoPrediction(0, "psi", psi);
    setEgoPrediction(0, "v", v);
    setEgoPrediction(0, "spline", spline);
    setEgoPrediction(0, "a", a);
    setEgoPrediction(0, "w", 0);
    for (int k = 1; k <= N; k++) {
        x += v * dt * std::cos(psi);
        y += v * dt * std::sin(psi);
        spline += v * dt;
        v += a * dt;
        v = std::max(v, 0.);
        setEgoPrediction(k, "x", x);
        setEgoPrediction(k, "y", y);
        setEgoPrediction(k, "psi", psi);
        setEgoPrediction(k, "v", v);
        setEgoPrediction(k, "spline", spline);
        setEgoPrediction(k, "a", a);
        setEgoPrediction(k, "w", 0);
    }
}

void Solver::initializeWarmstart(const State &initial_state, bool shift_previous_solution_forward) {  // :344-376
    if (shift_previous_solution_forward) {
        for (int k = 0; k <= N; k++)
            for (const auto &kv : _model_map) {
                const std::string &name = kv.first;
                if (k == 0) setEgoPrediction(0, name, kv.second.type == 'x' ? initial_state.get(name) : getOutput(1, name));
                else if (k == N - 1) setEgoPrediction(N - 1, name, getOutput(N - 1, name));
                else if (k == N) setEgoPrediction(N, name, getOutput(N - 1, name));
                else setEgoPrediction(k, name, getOutput(k + 1, name));
            }
    } else {
        for (int k = 0; k < N; k++)
            for (const auto &kv : _model_map) setEgoPrediction(k, kv.first, getOutput(k, kv.first));
    }
}

// OUTPUT (:379-389)
double Solver::getOutput(int k, const std::string &state_name) const {
    const ModelMapEntry &e = _model_map.at(state_name);
    if (e.type == 'x') return _output.xtraj[k * nx + e.index - nu];
    return _output.utraj[std::min(k, N - 1) * nu + e.index];
}

std::string Solver::explainExitFlag(int exitflag) const {  // :391-424
    switch (exitflag) {
    case 1: return "Success";
    case 0: return "Failure (no more information)";
    case 2: return "Failure (maximum number of iterations reached)";
    case 3: return "Failure (minimum step size reached)";
    case 4: break;
    default: return "Unknown exit code";
    }
    switch (_info.qp_status) {
    case 1: return "QP Failure: No more information on QP failure";
    case 2: return "QP Failure: Max Iterations";
    case 3: return "QP Failure: Minimal Step Reached";
    case 4: return "QP Failure: NAN in solution";
    case 5: return "QP Failure: Inconsistent Equality Constraints";
    default: return "QP Failure: UNKNOWN";
    }
}

void Solver::printIfBoundLimited() const {  // :426-447
    for (int k = 0; k < N; k++)
        for (const auto &kv : _model_map) {
            if (k == 0 && kv.second.type == 'x') continue;
            const double v = getOutput(k, kv.first);
            if (std::abs(v - kv.second.lower) < 1e-2) std::printf("%s limited by lower bound\n", kv.first.c_str());
            if (std::abs(v - kv.second.upper) < 1e-2) std::printf("%s limited by upper bound\n", kv.first.c_str());
        }
}
}  // namespace MPCPlanner
