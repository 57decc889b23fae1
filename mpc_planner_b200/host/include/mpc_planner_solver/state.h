// state.h -- name-keyed state vector; mirrors mpc_planner_solver/include/mpc_planner_solver/state.h:13-30
// and src/state.cpp:7-43 (model_map lookups; no Eigen / yaml-cpp).
#ifndef MPC_PLANNER_CUDA_STATE_H
#define MPC_PLANNER_CUDA_STATE_H

#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace MPCPlanner {
struct ModelMapEntry;
const std::map<std::string, ModelMapEntry> &generatedModelMap();

class State {
  public:
    State();
    void initialize();
    double get(const std::string &var_name) const;
    std::array<double, 2> getPos() const { return {get("x"), get("y")}; }
    void set(const std::string &var_name, double value);
    void print() const;

  private:
    std::vector<double> _state;
    int _nu;
};
}  // namespace MPCPlanner
#endif
