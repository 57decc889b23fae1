// test_modules.cpp -- one control step of the reference's Planner loop (mpc_planner/src/planner.cpp:37-158) with
// the CUDA-backed Solver and the C++ module layer: warm start -> module.update -> module.setParameters(k) ->
// GuidanceConstraints::optimize.   The scene comes from a binary file written by tests/test_host_cpp.py; every
// planner's packed inputs (and, in "gpu" mode, results) are written back for cross-checking against the oracle.
//   usage: test_modules pack|gpu <scene.bin> <out.bin>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <mpc_planner_modules/modules.h>

using namespace MPCPlanner;

static double rd(FILE *f) {
    double v;
    if (std::fread(&v, sizeof(double), 1, f) != 1) std::exit(2);
    return v;
}
static int ri(FILE *f) { return (int)rd(f); }

int main(int argc, char **argv) {
    if (argc != 4) return 3;
    const std::string mode = argv[1];
    FILE *f = std::fopen(argv[2], "rb");
    if (!f) return 4;
    // ---- scene ----
    State state;
    for (const char *n : {"x", "y", "psi", "v", "spline"}) state.set(n, rd(f));
    auto path = std::make_shared<Spline2D>();
    const int n_seg = ri(f);
    for (int i = 0; i < n_seg; i++) {
        PathSegment g;
        g.ax = rd(f); g.bx = rd(f); g.cx = rd(f); g.dx = rd(f);
        g.ay = rd(f); g.by = rd(f); g.cy = rd(f); g.dy = rd(f);
        g.start = rd(f);
        path->segments.push_back(g);
    }
    path->length = rd(f);
    RealTimeData data;
    data.robot_area.emplace_back(0.0, CONFIG().robot_radius);
    const int n_obs = ri(f), n_steps = ri(f);
    for (int o = 0; o < n_obs; o++) {
        DynamicObstacle obs;
        obs.index = o;
        obs.angle = 0.;
        obs.radius = CONFIG().obstacle_radius;
        obs.prediction.type = PredictionType::DETERMINISTIC;
        obs.prediction.modes.emplace_back();
        for (int k = 0; k < n_steps; k++) {
            PredictionStep st;
            st.position = {rd(f), rd(f)};
            st.angle = 0.;
            st.major_radius = st.minor_radius = 0.;
            obs.prediction.modes[0].push_back(st);
        }
        obs.position = obs.prediction.modes[0][0].position;
        data.dynamic_obstacles.push_back(obs);
    }
    const int n_guid = ri(f), n_pts = ri(f);
    std::vector<GuidanceTrajectory> guidance(n_guid);
    for (int g = 0; g < n_guid; g++) {
        guidance[g].topology_class = g;
        guidance[g].previously_selected = (ri(f) != 0);
        for (int k = 0; k < n_pts; k++) guidance[g].positions.push_back({rd(f), rd(f)});
        for (int k = 0; k < n_pts; k++) guidance[g].velocities.push_back({rd(f), rd(f)});
    }
    CONFIG().selection_weight_consistency = rd(f);
    std::fclose(f);

    // ---- Planner::solveMPC (planner.cpp:37-158) ----
    auto solver = std::make_shared<Solver>(0);
    ModuleData module_data;
    module_data.path = path;
    MPCBaseModule base(solver);
    Contouring contouring(solver);
    GuidanceConstraints guidance_module(solver, n_guid);
    guidance_module.setGuidanceTrajectories(guidance);
    solver->initializeWithBraking(state);                       // :83-86 (no previous solution)
    solver->setXinit(state);                                    // :88
    base.update(state, data, module_data);                      // :95-96
    contouring.update(state, data, module_data);
    solver->setXinit(state);                                    // spline state set by Contouring::update
    for (int k = 0; k < solver->N; k++) {                       // :102-108
        base.setParameters(data, module_data, k);
        contouring.setParameters(data, module_data, k);
    }
    solver->loadWarmstart();                                    // :115
    int exit_code = EXIT_CODE_NOT_OPTIMIZED_YET;
    std::vector<GuidanceConstraints::LocalPlanner *> active;
    if (mode == "gpu") {
        exit_code = guidance_module.optimize(state, data, module_data);  // :126-131
        for (auto &p : guidance_module.planners_)
            if (!p.disabled) active.push_back(&p);
    } else {
        active = guidance_module.prepare(state, data, module_data);
    }
    FILE *o = std::fopen(argv[3], "wb");
    if (!o) return 5;
    auto wd = [&](double v) { std::fwrite(&v, sizeof(double), 1, o); };
    wd((double)active.size());
    wd((double)exit_code);
    wd((double)guidance_module.best_planner_index());
    for (auto *p : active) {
        Solver &s = *p->local_solver;
        wd((double)p->id);
        wd((double)p->is_original_planner);
        std::fwrite(s._params.xinit.data(), sizeof(double), s._params.xinit.size(), o);
        std::fwrite(s._params.x0.data(), sizeof(double), s._params.x0.size(), o);
        std::fwrite(s._params.all_parameters.data(), sizeof(double), s._params.all_parameters.size(), o);
        wd((double)p->result.exit_code);
        wd(p->result.objective);
        std::fwrite(s._output.xtraj.data(), sizeof(double), s._output.xtraj.size(), o);
        std::fwrite(s._output.utraj.data(), sizeof(double), s._output.utraj.size(), o);
    }
    // the main solver after the decision (guidance_constraints.cpp:382-384)
    std::fwrite(solver->_output.xtraj.data(), sizeof(double), solver->_output.xtraj.size(), o);
    wd(solver->_info.pobj);
    std::fclose(o);
    if (mode == "pack") {
        // ---- planner <-> homotopy class mapping and timeouts of the NEXT control step (guidance_constraints.cpp:192-250,
        // 310-313, 331-333): pretend the previous step's planners 0..3 optimised classes 7, 5, 9, 5 ----
        auto check = [](bool ok, int line) {
            if (!ok) {
                std::fprintf(stderr, "check failed at line %d\n", line);
                std::exit(1);
            }
        };
        // pretend the previous step's planners 0..2 optimised classes 7, 5, 5
        const int prev[3] = {7, 5, 5};
        for (int p = 0; p < 3 && p < (int)guidance_module.planners_.size(); p++) guidance_module.planners_[p].result.guidance_ID = prev[p];
        if (n_guid >= 2 && guidance_module.planners_.size() >= 3) {
            guidance[0].topology_class = 5;   // -> planner 1 (the first planner that solved class 5): an "existing" guidance
            guidance[1].topology_class = 11;  // new class: the free planners are handed out as the reference's loop does
            guidance_module.setGuidanceTrajectories(guidance);
            check(data.planning_start_time == std::chrono::system_clock::time_point{}, __LINE__);
            guidance_module.prepare(state, data, module_data);
            const auto &map = guidance_module.homotopyClassToPlanner();
            check(map.at(0) == 1 && guidance_module.planners_[1].existing_guidance, __LINE__);
            check(map.count(1) == 1 && map.at(1) == (int)guidance_module.planners_.size() - 1, __LINE__);  // no `break` in :232-241
            check(!guidance_module.planners_[0].existing_guidance && !guidance_module.planners_[map.at(1)].existing_guidance, __LINE__);
            check(guidance_module.planners_[0].taken, __LINE__);
            for (auto &p : guidance_module.planners_)
                if (!p.disabled) check(p.local_solver->_params.solver_timeout == 0.0, __LINE__);  // start time unset: untouched budget
            data.planning_start_time = std::chrono::system_clock::now();
            guidance_module.prepare(state, data, module_data);
            for (auto &p : guidance_module.planners_)
                if (!p.disabled) check(p.local_solver->_params.solver_timeout > 0.0 && p.local_solver->_params.solver_timeout < 1. / CONFIG().control_frequency, __LINE__);
            data.planning_start_time = std::chrono::system_clock::now() - std::chrono::seconds(1);  // budget used up
            guidance_module.prepare(state, data, module_data);
            for (auto &p : guidance_module.planners_)
                if (!p.disabled) check(p.local_solver->_params.solver_timeout == 1e-9, __LINE__);  // -> one RTI iteration, as in the reference
            std::printf("homotopy mapping and timeout checks passed\n");
        }
    }
    std::printf("test_modules %s: %zu planners, exit %d, best %d\n", mode.c_str(), active.size(), exit_code, guidance_module.best_planner_index());
    return 0;
}
