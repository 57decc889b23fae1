// test_planner.cpp -- closed loop through MPCPlanner::Planner::solveMPC (mpc_planner/src/planner.cpp:37-158) on the
// CUDA backend, without ROS: a unicycle follows a reference path among constant-velocity obstacles; every control
// step = Planner::solveMPC (T-MPC++: all guesses in ONE launch) + apply the first input for dt (or brake when the
// MPC fails, as the reference's control node does).  The guidance planner is external to the reference tree; its
// output is synthesised here (pass each relevant obstacle on the left / right).
//   usage: test_planner cpu                     (construction + data-readiness semantics, no GPU)
//          test_planner gpu <scene.bin> <out.bin>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <mpc_planner/planner.h>

using namespace MPCPlanner;

#define CHECK(c)                                                               \
    do {                                                                       \
        if (!(c)) {                                                            \
            std::printf("CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #c); \
            std::exit(2);                                                      \
        }                                                                      \
    } while (0)

static double rd(FILE *f) {
    double v;
    CHECK(std::fread(&v, sizeof(double), 1, f) == 1);
    return v;
}
static int ri(FILE *f) { return (int)rd(f); }

static void path_frame(const Spline2D &path, double s, Vec2 &p, Vec2 &tangent) {
    p = path.getPoint(s);
    const Vec2 q = path.getPoint(s + 1e-3);
    const double n = std::hypot(q[0] - p[0], q[1] - p[1]);
    tangent = {(q[0] - p[0]) / n, (q[1] - p[1]) / n};
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    {   // no data yet: solveMPC reports failure and does not touch the GPU (planner.cpp:46-58)
        Planner planner;
        State state;
        RealTimeData data;
        PlannerOutput out = planner.solveMPC(state, data);
        CHECK(!out.success);
        CHECK(planner._modules.size() >= 3);
        CHECK(!planner.isObjectiveReached(state, data));
    }
    if (mode != "gpu") {
        std::printf("test_planner cpu: all checks passed\n");
        return 0;
    }
    CHECK(argc == 4);
    FILE *f = std::fopen(argv[2], "rb");
    CHECK(f != nullptr);
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
    const int n_obs = ri(f);
    std::vector<Vec2> obs_pos(n_obs), obs_vel(n_obs);
    for (int o = 0; o < n_obs; o++) {
        obs_pos[o] = {rd(f), rd(f)};
        obs_vel[o] = {rd(f), rd(f)};
    }
    const int n_steps = ri(f);
    const double v_ref = rd(f);
    std::fclose(f);

    Planner planner;
    planner.setReferencePath(path);
    const int N = planner._solver->N;
    const double dt = planner._solver->dt;
    FILE *o = std::fopen(argv[3], "wb");
    CHECK(o != nullptr);
    auto wd = [&](double v) { std::fwrite(&v, sizeof(double), 1, o); };
    wd((double)n_steps);
    int previous_class = -1;
    for (int step = 0; step < n_steps; step++) {
        RealTimeData data;
        data.robot_area.emplace_back(0.0, CONFIG().robot_radius);
        for (int ob = 0; ob < CONFIG().max_obstacles; ob++) {  // ensureObstacleSize pads with far-away dummies
            DynamicObstacle obs;
            obs.index = ob;
            obs.angle = 0.;
            obs.radius = CONFIG().obstacle_radius;
            obs.prediction.type = PredictionType::DETERMINISTIC;
            obs.prediction.modes.emplace_back();
            for (int k = 0; k < N; k++) {
                PredictionStep st;
                if (ob < n_obs) st.position = {obs_pos[ob][0] + (k + 1) * dt * obs_vel[ob][0], obs_pos[ob][1] + (k + 1) * dt * obs_vel[ob][1]};
                else st.position = {state.get("x") + 100.0 + ob, state.get("y") + 100.0};
                st.angle = 0.;
                st.major_radius = st.minor_radius = 0.;
                obs.prediction.modes[0].push_back(st);
            }
            obs.position = (ob < n_obs) ? obs_pos[ob] : obs.prediction.modes[0][0].position;
            data.dynamic_obstacles.push_back(obs);
        }
        // guidance stand-in: follow the path at v_ref with a lateral bump of either sign around the obstacle that
        // comes closest to the nominal motion
        int seg = 0;
        double s0 = 0.;
        path->findClosestPoint(state.getPos(), seg, s0);
        double best_d = 1e9, bump_t = 0., bump_rel = 0.;
        for (int k = 1; k <= N; k++) {
            Vec2 p, t;
            path_frame(*path, s0 + v_ref * k * dt, p, t);
            for (int ob = 0; ob < n_obs; ob++) {
                const Vec2 q = {obs_pos[ob][0] + k * dt * obs_vel[ob][0], obs_pos[ob][1] + k * dt * obs_vel[ob][1]};
                const double d = std::hypot(q[0] - p[0], q[1] - p[1]);
                if (d < best_d) {
                    best_d = d;
                    bump_t = k * dt;
                    bump_rel = -(q[0] - p[0]) * t[1] + (q[1] - p[1]) * t[0];  // lateral position of the obstacle
                }
            }
        }
        std::vector<GuidanceTrajectory> guidance;
        for (int g = 0; g < 2; g++) {
            GuidanceTrajectory traj;
            traj.topology_class = g;
            traj.previously_selected = (g == previous_class);
            const double side = (g == 0) ? 1.0 : -1.0;
            const double amp = (best_d < 2.5) ? bump_rel + side * 1.4 : side * 0.3;
            for (int k = 0; k <= N; k++) {
                Vec2 p, t;
                path_frame(*path, s0 + v_ref * k * dt, p, t);
                const double w = amp * std::exp(-0.5 * std::pow((k * dt - bump_t) / 1.0, 2)) * std::min(1.0, k * dt / std::max(bump_t, dt));
                traj.positions.push_back({p[0] - w * t[1], p[1] + w * t[0]});
            }
            for (int k = 0; k <= N; k++) {
                const int a = std::max(k - 1, 0), b = std::min(k + 1, N);
                traj.velocities.push_back({(traj.positions[b][0] - traj.positions[a][0]) / ((b - a) * dt),
                                           (traj.positions[b][1] - traj.positions[a][1]) / ((b - a) * dt)});
            }
            guidance.push_back(traj);
        }
        planner.setGuidanceTrajectories(guidance);

        data.planning_start_time = std::chrono::system_clock::now();  // what the ROS wrapper's loop callback does
        PlannerOutput out = planner.solveMPC(state, data);
        double a = -3.0, w = 0.0;  // braking command when the MPC fails
        if (out.success) {
            a = planner.getSolution(0, "a");
            w = planner.getSolution(0, "w");
            const int best = planner.bestPlannerIndex();
            previous_class = (best >= 0 && best < 2) ? best : -1;
        }
        // record the state BEFORE applying the input
        double clearance = 1e9;
        for (int ob = 0; ob < n_obs; ob++) clearance = std::min(clearance, std::hypot(obs_pos[ob][0] - state.get("x"), obs_pos[ob][1] - state.get("y")));
        wd((double)out.success);
        wd((double)planner.last_exit_flag);
        wd((double)planner.bestPlannerIndex());
        wd(state.get("x")); wd(state.get("y")); wd(state.get("psi")); wd(state.get("v")); wd(state.get("spline"));
        wd(a); wd(w); wd(clearance);
        // plant: second-order unicycle, RK4 with 10 substeps (solver_model.py:193-214)
        double x = state.get("x"), y = state.get("y"), psi = state.get("psi"), v = state.get("v");
        if (v + a * dt < 0.0) a = -v / dt;  // the braking command does not reverse
        const int sub = 10;
        const double h = dt / sub;
        for (int i = 0; i < sub; i++) {
            auto fx = [&](double ps, double vv) { return vv * std::cos(ps); };
            auto fy = [&](double ps, double vv) { return vv * std::sin(ps); };
            const double k1x = fx(psi, v), k1y = fy(psi, v);
            const double k2x = fx(psi + 0.5 * h * w, v + 0.5 * h * a), k2y = fy(psi + 0.5 * h * w, v + 0.5 * h * a);
            const double k4x = fx(psi + h * w, v + h * a), k4y = fy(psi + h * w, v + h * a);
            x += h / 6.0 * (k1x + 4.0 * k2x + k4x);
            y += h / 6.0 * (k1y + 4.0 * k2y + k4y);
            psi += h * w;
            v += h * a;
        }
        state.set("x", x); state.set("y", y); state.set("psi", psi); state.set("v", v);
        for (int ob = 0; ob < n_obs; ob++) {
            obs_pos[ob][0] += dt * obs_vel[ob][0];
            obs_pos[ob][1] += dt * obs_vel[ob][1];
        }
    }
    std::fclose(o);
    std::printf("test_planner gpu: %d control steps done\n", n_steps);
    return 0;
}
