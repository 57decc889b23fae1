// test_packing.cpp -- parameter packing of the constraint modules that the T-MPC test scene does not exercise (CPU only):
//   GaussianConstraints::setParameters (mpc_planner_modules/src/gaussian_constraints.cpp:31-79)   [stack with the CC-MPC module]
//   DecompConstraints::update / setParameters (decomp_constraints.cpp:86-114, 150-189)            [cfg 3]
// Built once per module stack (the generated header says which modules exist).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include <mpc_planner_modules/modules.h>

using namespace MPCPlanner;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__); \
            std::exit(1);                                                    \
        }                                                                    \
    } while (0)

int main() {
    auto solver = std::make_shared<Solver>(0);
    State state;
    state.set("x", 2.0);
    state.set("y", -1.0);
    RealTimeData data;
    data.robot_area.emplace_back(0.0, CONFIG().robot_radius);
    ModuleData module_data;
    const int N = solver->N;
    int checked = 0;
#if defined(MPC_PLANNER_HAS_GAUSSIAN)
    {
        for (int o = 0; o < CONFIG().max_obstacles; o++) {
            DynamicObstacle obs;
            obs.index = o;
            obs.radius = CONFIG().obstacle_radius;
            obs.type = (o == 1) ? ObstacleType::STATIC : ObstacleType::DYNAMIC;
            obs.prediction.type = PredictionType::GAUSSIAN;
            obs.prediction.modes.emplace_back();
            for (int k = 0; k < N; k++) {
                PredictionStep st;
                st.position = {10.0 + o + 0.1 * k, -3.0 + 0.2 * k};
                st.angle = 0.;
                st.major_radius = 0.3 + 0.01 * k;
                st.minor_radius = 0.2 + 0.01 * k;
                obs.prediction.modes[0].push_back(st);
            }
            obs.position = obs.prediction.modes[0][0].position;
            data.dynamic_obstacles.push_back(obs);
        }
        GaussianConstraints gaussian(solver);
        gaussian.update(state, data, module_data);
        for (int k = 0; k < N; k++) gaussian.setParameters(data, module_data, k);
        // k = 0: dummies far away (:39-52); k >= 1: prediction step k - 1 (:62-63), static obstacles without uncertainty (:70-74)
        CHECK(solver->getParameter(0, "gaussian_obst_0_x") == 102.0 && solver->getParameter(0, "gaussian_obst_0_y") == 99.0);
        CHECK(solver->getParameter(0, "gaussian_obst_3_major") == 0.1 && solver->getParameter(0, "gaussian_obst_3_risk") == 0.05);
        CHECK(solver->getParameter(0, "gaussian_obst_3_r") == 0.1);
        CHECK(solver->getParameter(4, "gaussian_obst_2_x") == 10.0 + 2 + 0.1 * 3);
        CHECK(solver->getParameter(4, "gaussian_obst_2_y") == -3.0 + 0.2 * 3);
        CHECK(solver->getParameter(4, "gaussian_obst_2_major") == 0.3 + 0.01 * 3 && solver->getParameter(4, "gaussian_obst_2_minor") == 0.2 + 0.01 * 3);
        CHECK(solver->getParameter(4, "gaussian_obst_1_major") == 0.001 && solver->getParameter(4, "gaussian_obst_1_minor") == 0.001);
        CHECK(solver->getParameter(4, "gaussian_obst_2_risk") == CONFIG().risk);
        CHECK(solver->getParameter(4, "gaussian_obst_2_r") == CONFIG().obstacle_radius);
        CHECK(solver->getParameter(7, "ego_disc_radius") == CONFIG().robot_radius && solver->getParameter(7, "ego_disc_0_offset") == 0.0);
        checked++;
    }
#endif
#if defined(MPC_PLANNER_HAS_DECOMP)
    {
        const int M = CONFIG().decomp_max_constraints;
        module_data.decomp_polyhedra.assign(N - 1, {});
        for (int k = 0; k < N - 1; k++) {
            for (int i = 0; i < 5; i++) module_data.decomp_polyhedra[k].push_back(Halfspace{{std::cos(1.0 * i), std::sin(1.0 * i)}, 3.0 + k + 0.1 * i});
        }
        module_data.decomp_polyhedra[2][3].A = {0.0, 0.0};  // a zero normal ends the polyhedron's rows (:97-100)
        module_data.decomp_polyhedra[4][1].A = {std::numeric_limits<double>::quiet_NaN(), 1.0};
        DecompConstraints decomp(solver);
        decomp.update(state, data, module_data);
        for (int k = 0; k < N; k++) decomp.setParameters(data, module_data, k);
        const double dummy_b = 2.0 + 100.0;
        // k = 0: dummies (:153-168); polyhedron k -> stage k + 1 (:103-105); dummies (1, 0, x + 100) after the valid rows (:108-113)
        CHECK(solver->getParameter(0, "disc_0_decomp_0_a1") == 1.0 && solver->getParameter(0, "disc_0_decomp_0_a2") == 0.0);
        CHECK(solver->getParameter(0, "disc_0_decomp_0_b") == dummy_b);
        CHECK(solver->getParameter(1, "disc_0_decomp_2_a1") == std::cos(2.0) && solver->getParameter(1, "disc_0_decomp_2_a2") == std::sin(2.0));
        CHECK(solver->getParameter(1, "disc_0_decomp_2_b") == 3.0 + 0 + 0.2);
        CHECK(solver->getParameter(1, "disc_0_decomp_5_a1") == 1.0 && solver->getParameter(1, "disc_0_decomp_5_b") == dummy_b);
        CHECK(solver->getParameter(1, "disc_0_decomp_" + std::to_string(M - 1) + "_b") == dummy_b);
        CHECK(solver->getParameter(3, "disc_0_decomp_2_b") == 3.0 + 2 + 0.2);                                            // rows before the zero normal
        CHECK(solver->getParameter(3, "disc_0_decomp_3_a1") == 1.0 && solver->getParameter(3, "disc_0_decomp_4_b") == dummy_b);  // and dummies from it on
        CHECK(solver->getParameter(5, "disc_0_decomp_0_b") == 3.0 + 4 && solver->getParameter(5, "disc_0_decomp_1_b") == dummy_b);  // NaN row
        CHECK(solver->getParameter(N - 1, "disc_0_decomp_0_b") == 3.0 + (N - 2));
        checked++;
    }
#endif
    CHECK(checked >= 1);
    std::printf("test_packing: %d module(s) checked, all checks passed\n", checked);
    return 0;
}
