// test_solver.cpp -- the assertions of the reference's mpc_planner_solver/test/test_solver.cpp:52-134
// (State set/get, Solver dims, parameter set/get, setXinit sums, ego-prediction indexing, warm-start
// copy semantics, params-only operator=) against the CUDA-backed Solver; plus, in "gpu" mode, a real
// solve() and a solveBatch() whose inputs/outputs are exchanged with the Python tests through
// binary files.   usage: test_solver cpu | gpu <in.bin> <out.bin>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mpc_planner_generated.h"
#include <mpc_planner_solver/solver_interface.h>

using namespace MPCPlanner;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                    \
        }                                                                    \
    } while (0)

static void test_state() {  // test_solver.cpp:52-65
    State state;
    state.set("x", 1.5);
    state.set("y", 3.5);
    CHECK(state.get("y") == 3.5);
    CHECK(state.get("x") == 1.5);
    CHECK(state.getPos()[0] == 1.5);
    CHECK(state.getPos()[1] == 3.5);
}

static void test_solver_semantics() {  // test_solver.cpp:67-134
    Solver solver;
    CHECK(solver.nu + solver.nx == solver.nvar);
    CHECK(solver.npar > 0);
    CHECK(solver.dt > 0.);
    for (int k = 0; k < solver.N; k++) solver.setParameter(k, "reference_velocity", 1.);
    for (int k = 0; k < solver.N; k++) CHECK(solver.getParameter(k, "reference_velocity") == 1.);
    CHECK(solver.hasParameter("lag") && !solver.hasParameter("no_such_parameter"));

    solver.setXinit("x", 5.4);
    solver.setXinit("y", 1.4);
    double sum_init = 0.;
    for (unsigned int i = 0; i < solver.nx; i++) sum_init += solver._params.xinit[i];
    CHECK(std::abs(sum_init - 6.8) < 1e-5);

    State state;
    state.set("x", 4.4);
    state.set("y", 1.2);
    solver.setXinit(state);
    sum_init = 0.;
    for (unsigned int i = 0; i < solver.nx; i++) sum_init += solver._params.xinit[i];
    CHECK(std::abs(sum_init - 5.6) < 1e-5);

    state = State();
    solver.setXinit(state);

    for (int k = 0; k < solver.N; k++) {
        solver.setEgoPrediction(k, "x", k * 1.);
        solver.setEgoPrediction(k, "y", 0.);
    }
    for (int k = 0; k < solver.N; k++) {
        CHECK(solver.getEgoPrediction(k, "x") == k * 1);
        CHECK(solver.getEgoPrediction(k, "y") == 0);
    }
    // [u;x] node layout of x0 (acados_solver_interface.h:54): x of node k at k*nvar + nu + 0
    CHECK(solver._params.x0[3 * solver.nvar + solver.nu] == 3.0);

    // put a fake plan into the output as the solver would, then warm start from it (non-shifted variant
    // keeps node k at node k; acados_solver_interface.cpp:366-375)
    const int x_index = solver._model_map.at("x").index - solver.nu;
    solver._output.xtraj[0 * solver.nx + x_index] = 0.0;
    solver._output.xtraj[1 * solver.nx + x_index] = 1.0;
    solver._output.xtraj[2 * solver.nx + x_index] = 2.0;
    state.set("x", 0.8);
    solver.initializeWarmstart(state, false);
    CHECK(solver.getEgoPrediction(1, "x") == 1.0);
    CHECK(solver.getEgoPrediction(2, "x") == 2.0);
    solver.initializeWarmstart(state, true);  // shifted: x_{k+1} initialises x_k, node 0 is the state
    CHECK(solver.getEgoPrediction(0, "x") == 0.8);
    CHECK(solver.getEgoPrediction(1, "x") == 2.0);
    solver.loadWarmstart();

    Solver solver2(1);
    solver2.setParameter(0, "reference_velocity", 3.);
    CHECK(solver._solver_id != solver2._solver_id);
    solver2._output.xtraj[0] = 42.0;
    solver2 = solver;  // parameters only
    CHECK(solver2.getParameter(0, "reference_velocity") == 1.);
    CHECK(solver2._output.xtraj[0] == 42.0);

    // generated bundle setters (generate_cpp_files.py:235-254)
    setSolverParameterSplineXA(2, solver._params, 7.5, 1);
    CHECK(solver.getParameter(2, "spline_x1_a") == 7.5);
    setSolverParameterContour(3, solver._params, 0.25);
    CHECK(solver.getParameter(3, "contour") == 0.25);

    // braking fallback (acados_solver_interface.cpp:303-342)
    State s2;
    s2.set("v", 1.0);
    solver.initializeWithBraking(s2);
    CHECK(std::abs(solver.getEgoPrediction(1, "x") - 1.0 * solver.dt) < 1e-12);
    CHECK(solver.getEgoPrediction(solver.N, "v") == 0.0);
    CHECK(solver.explainExitFlag(1) == "Success");
}

static std::vector<double> read_doubles(FILE *f, size_t n) {
    std::vector<double> v(n);
    if (std::fread(v.data(), sizeof(double), n, f) != n) {
        std::printf("short read\n");
        std::exit(2);
    }
    return v;
}

static void test_gpu(const char *in_path, const char *out_path) {
    FILE *f = std::fopen(in_path, "rb");
    CHECK(f != nullptr);
    int P = 0;
    CHECK(std::fread(&P, sizeof(int), 1, f) == 1);
    std::vector<Solver *> solvers;
    for (int i = 0; i < P; i++) {
        Solver *s = new Solver(i);
        s->_params.xinit = read_doubles(f, s->_params.xinit.size());
        s->_params.x0 = read_doubles(f, s->_params.x0.size());
        s->_params.all_parameters = read_doubles(f, s->_params.all_parameters.size());
        s->loadWarmstart();
        solvers.push_back(s);
    }
    std::fclose(f);
    // one Solver::solve() (P = 1 path) ...
    Solver single(100);
    single = *solvers[0];
    single.loadWarmstart();
    int exit_single = single.solve();
    // ... and the batch path
    std::vector<int> exit_codes;
    int best = Solver::solveBatch(solvers, exit_codes);
    CHECK(exit_single == exit_codes[0]);
    for (size_t i = 0; i < single._output.xtraj.size(); i++) CHECK(single._output.xtraj[i] == solvers[0]->_output.xtraj[i]);
    CHECK(single._info.pobj == solvers[0]->_info.pobj);
    FILE *o = std::fopen(out_path, "wb");
    CHECK(o != nullptr);
    std::fwrite(&best, sizeof(int), 1, o);
    for (int i = 0; i < P; i++) {
        std::fwrite(&exit_codes[i], sizeof(int), 1, o);
        std::fwrite(&solvers[i]->_info.pobj, sizeof(double), 1, o);
        std::fwrite(solvers[i]->_output.xtraj.data(), sizeof(double), solvers[i]->_output.xtraj.size(), o);
        std::fwrite(solvers[i]->_output.utraj.data(), sizeof(double), solvers[i]->_output.utraj.size(), o);
        // getOutput(k, name) indexing (acados_solver_interface.cpp:379-389)
        CHECK(solvers[i]->getOutput(1, "x") == solvers[i]->_output.xtraj[1 * solvers[i]->nx + 0]);
        CHECK(solvers[i]->getOutput(0, "w") == solvers[i]->_output.utraj[1]);
    }
    // second control step on `single`: parameters re-copied from solver 0 (operator= keeps single's solver memory),
    // warm start = its own previous solution; the multipliers of the first solve enter the first exact Hessian.
    CHECK(!single.nlpMultipliers().empty());
    bool any_mult = false;
    for (double v : single.nlpMultipliers()) any_mult = any_mult || v != 0.0;
    CHECK(any_mult);
    single = *solvers[0];
    for (int k = 0; k < single.N; k++) {
        single.setEgoPrediction(k, "a", single.getOutput(k, "a"));
        single.setEgoPrediction(k, "w", single.getOutput(k, "w"));
    }
    for (int k = 0; k <= single.N; k++)
        for (const char *name : {"x", "y", "psi", "v", "spline"}) single.setEgoPrediction(k, name, single.getOutput(k, name));
    single.loadWarmstart();
    int exit_second = single.solve();
    std::fwrite(&exit_second, sizeof(int), 1, o);
    std::fwrite(&single._info.pobj, sizeof(double), 1, o);
    std::fwrite(single._output.xtraj.data(), sizeof(double), single._output.xtraj.size(), o);
    // _params.solver_timeout (acados_solver_interface.cpp:91-117): a budget no iteration fits in -> one RTI iteration
    single = *solvers[0];
    single._params.solver_timeout = 1e-9;
    single.loadWarmstart();
    (void)single.solve();
    CHECK(single._info.sqp_iter == 1);
    single._params.solver_timeout = 10.0;
    single.loadWarmstart();
    (void)single.solve();
    CHECK(single._info.sqp_iter > 1);
    single.reset();
    for (double v : single.nlpMultipliers()) CHECK(v == 0.0);
    std::fclose(o);
    for (Solver *s : solvers) delete s;
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    test_state();
    test_solver_semantics();
    if (mode == "gpu") {
        CHECK(argc == 4);
        test_gpu(argv[2], argv[3]);
    }
    std::printf("test_solver %s: all checks passed\n", mode.c_str());
    return 0;
}
