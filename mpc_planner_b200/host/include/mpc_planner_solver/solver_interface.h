// solver_interface.h -- third branch of the reference's compile-time solver switch
// (mpc_planner_solver/include/mpc_planner_solver/solver_interface.h:4-12).
#ifndef MPC_PLANNER_SOLVER_INTERFACE_H
#define MPC_PLANNER_SOLVER_INTERFACE_H
#if defined(MPC_PLANNER_CUDA_SOLVER)
#include <mpc_planner_solver/cuda_solver_interface.h>
#elif defined(ACADOS_SOLVER)
#include <mpc_planner_solver/acados_solver_interface.h>
#else
#include <mpc_planner_solver/forces_solver_interface.h>
#endif
#endif
