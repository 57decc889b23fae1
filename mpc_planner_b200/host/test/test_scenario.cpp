// test_scenario.cpp -- Safe Horizon MPC control flow (mpc_planner_modules/src/scenario_constraints.cpp:38-108) with the
// CUDA-backed Solver: the `parallel_solvers` local solvers (each with its own scenario halfspaces) are solved as ONE batch.
//   usage: test_scenario cpu | gpu <in.bin> <out.bin>
//   in.bin : int P | xinit[NX] x0[(N+1) NZ] params[N NP] of the main solver | P x (N x n_rows x {a1, a2, b}) halfspaces
//   out.bin: int best | int exit | per solver {int exit, double pobj, xtraj, utraj} | main solver xtraj, pobj
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <mpc_planner_modules/modules.h>

using namespace MPCPlanner;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__); \
            std::exit(1);                                                    \
        }                                                                    \
    } while (0)

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    auto solver = std::make_shared<Solver>(0);
    ScenarioConstraints scenario(solver, 4);
    CHECK(scenario._scenario_solvers.size() == 4);
    RealTimeData data;
    data.robot_area.emplace_back(0.0, CONFIG().robot_radius);
    ModuleData module_data;
    State state;
    {   // parameter packing of the stand-in module: dummies at k = 0 and for missing rows, halfspaces elsewhere
        auto &ss = *scenario._scenario_solvers[1];
        std::vector<std::vector<Halfspace>> hs(solver->N);
        hs[3].push_back(Halfspace{{0.6, 0.8}, 2.5});
        ss.scenario_module.setHalfspaces(hs);
        ss.solver->_params.xinit[0] = 1.0;
        for (int k = 0; k < solver->N; k++) ss.scenario_module.setParameters(data, k);
        CHECK(ss.solver->getParameter(3, "disc_0_scenario_constraint_0_a1") == 0.6);
        CHECK(ss.solver->getParameter(3, "disc_0_scenario_constraint_0_a2") == 0.8);
        CHECK(ss.solver->getParameter(3, "disc_0_scenario_constraint_0_b") == 2.5);
        CHECK(ss.solver->getParameter(3, "disc_0_scenario_constraint_1_a1") == 1.0);
        CHECK(ss.solver->getParameter(3, "disc_0_scenario_constraint_1_b") == 101.0);
        CHECK(ss.solver->getParameter(0, "disc_0_scenario_constraint_0_b") == 101.0);
        CHECK(ss.solver->getParameter(2, "disc_0_scenario_constraint_23_a2") == 0.0);
    }
    if (mode == "gpu") {
        CHECK(argc == 4);
        FILE *f = std::fopen(argv[2], "rb");
        CHECK(f != nullptr);
        int P = 0;
        CHECK(std::fread(&P, sizeof(int), 1, f) == 1 && P == 4);
        CHECK(std::fread(solver->_params.xinit.data(), sizeof(double), solver->_params.xinit.size(), f) == solver->_params.xinit.size());
        CHECK(std::fread(solver->_params.x0.data(), sizeof(double), solver->_params.x0.size(), f) == solver->_params.x0.size());
        CHECK(std::fread(solver->_params.all_parameters.data(), sizeof(double), solver->_params.all_parameters.size(), f) ==
              solver->_params.all_parameters.size());
        int n_rows = 0;
        CHECK(std::fread(&n_rows, sizeof(int), 1, f) == 1);
        for (int p = 0; p < P; p++) {
            std::vector<std::vector<Halfspace>> hs(solver->N);
            for (int k = 0; k < solver->N; k++)
                for (int i = 0; i < n_rows; i++) {
                    double v[3];
                    CHECK(std::fread(v, sizeof(double), 3, f) == 3);
                    hs[k].push_back(Halfspace{{v[0], v[1]}, v[2]});
                }
            scenario._scenario_solvers[p]->scenario_module.setHalfspaces(hs);
        }
        std::fclose(f);
        data.planning_start_time = std::chrono::system_clock::now();
        scenario.update(state, data, module_data);                    // planner.cpp:95-96
        for (int k = 0; k < solver->N; k++) scenario.setParameters(data, module_data, k);
        solver->loadWarmstart();                                      // :115
        const int exit_code = scenario.optimize(state, data, module_data);
        FILE *o = std::fopen(argv[3], "wb");
        CHECK(o != nullptr);
        const int best = scenario.best_solver_index();
        std::fwrite(&best, sizeof(int), 1, o);
        std::fwrite(&exit_code, sizeof(int), 1, o);
        for (auto &ss : scenario._scenario_solvers) {
            CHECK(ss->solver->_params.solver_timeout > 0.0);  // :69-71
            std::fwrite(&ss->exit_code, sizeof(int), 1, o);
            std::fwrite(&ss->solver->_info.pobj, sizeof(double), 1, o);
            std::fwrite(ss->solver->_output.xtraj.data(), sizeof(double), ss->solver->_output.xtraj.size(), o);
            std::fwrite(ss->solver->_output.utraj.data(), sizeof(double), ss->solver->_output.utraj.size(), o);
        }
        std::fwrite(solver->_output.xtraj.data(), sizeof(double), solver->_output.xtraj.size(), o);
        std::fwrite(&solver->_info.pobj, sizeof(double), 1, o);
        std::fclose(o);
    }
    std::printf("test_scenario %s: all checks passed\n", mode.c_str());
    return 0;
}
