// projection.cuh -- Douglas-Rachford projection of a warm-start position out of the obstacle discs
// (LinearizedConstraints::projectToSafety, mpc_planner_modules/src/linearized_constraints.cpp:130-148).
//
// The reference calls RosTools::DouglasRachford::douglasRachfordProjection(pos, obstacle, anchor, r, pos), whose source
// (ros_tools) is NOT under /root/reference.  What is restated here is the published algorithm the name denotes, with the
// argument roles of the call site: Douglas-Rachford splitting for the two sets
//     A = complement of the anchor's disc (anchor = obstacle 0 of the same stage),   B = complement of the obstacle's disc,
//     p <- ( p + R_B(R_A(p)) ) / 2,     R_X = 2 P_X - I   (reflection through the projection onto X),
// where the projection of a point inside a disc (|p - c| < r) is the boundary point in the direction of the start pose
// seen from the centre, c + r (s - c) / |s - c|, and the identity outside.  A fixed point of the iteration lies in A and B;
// the caller runs at most 3 sweeps over the obstacles (:136).  Un-pinnable against ros_tools:
// for a point inside ONE disc the result is the radial push-out midpoint p/2 + (boundary point)/2 ... i.e. the iteration
// moves it towards the boundary geometrically; outside every disc it is the identity, which is the only case the shipped
// scenes exercise.  Host (modules.cpp) and device (pack kernels) share this one definition.
#pragma once

#if defined(__CUDACC__)
#define MPC_PROJ_HD __host__ __device__ __forceinline__
#else
#define MPC_PROJ_HD inline
#endif

#include <math.h>

namespace mpcgpu {

MPC_PROJ_HD void dr_project(double px, double py, double cx, double cy, double r, double sx, double sy, double &ox, double &oy) {
    const double dx = px - cx, dy = py - cy;
    if (sqrt(dx * dx + dy * dy) < r) {
        double ux = sx - cx, uy = sy - cy, n = sqrt(ux * ux + uy * uy);
        if (n < 1e-9) { ux = 1.0; uy = 0.0; n = 1.0; }  // start pose on the centre: any direction leaves the disc
        ox = cx + ux / n * r;
        oy = cy + uy / n * r;
    } else {
        ox = px;
        oy = py;
    }
}

MPC_PROJ_HD void dr_reflect(double px, double py, double cx, double cy, double r, double sx, double sy, double &ox, double &oy) {
    double qx, qy;
    dr_project(px, py, cx, cy, r, sx, sy, qx, qy);
    ox = 2.0 * qx - px;
    oy = 2.0 * qy - py;
}

// pos <- (pos + R_delta(R_anchor(pos))) / 2, start pose = pos (as the reference's call passes it)
MPC_PROJ_HD void douglas_rachford_projection(double &px, double &py, double delta_x, double delta_y, double anchor_x, double anchor_y, double r) {
    double ax, ay, bx, by;
    dr_reflect(px, py, anchor_x, anchor_y, r, px, py, ax, ay);
    dr_reflect(ax, ay, delta_x, delta_y, r, px, py, bx, by);
    px = 0.5 * (px + bx);
    py = 0.5 * (py + by);
}

}  // namespace mpcgpu
