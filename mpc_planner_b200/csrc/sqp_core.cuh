// sqp_core.cuh -- one full `Solver::solve()` (mpc_planner_solver/src/acados_solver_interface.cpp:86-204)
// executed by ONE WARP with the whole problem resident in shared memory.
//
// Replaces, for the hot path, what the reference delegates to acados (SQP-RTI), HPIPM (IPM QP) and
// BLASFEO: per RTI iteration  linearise all stages (lane = stage)  ->  MIRROR-regularise  ->
// Mehrotra predictor-corrector interior point on the stage-structured QP with a Riccati
// factorisation (warp-cooperative, lanes = matrix entries)  ->  full step.
//
// Structure: the routine is a sequence of warp "phases".  Inside a phase every lane works on
// data written by earlier phases and writes a disjoint part of shared memory; phases are separated
// by __syncwarp().  No register state crosses a phase boundary, so the same source compiles for
// the host as a sequential 32-lane emulation (tests/emu), used to debug without a GPU.
#pragma once

#include "stage_functions.cuh"

#if defined(__CUDACC__)
#define MPC_DEV __device__ __forceinline__
#define MPC_LANES \
    {             \
        const int lane = threadIdx.x & 31;
#define MPC_LANES_END \
    }                 \
    __syncwarp();
#else
#define MPC_DEV inline
#define MPC_LANES for (int lane = 0; lane < 32; ++lane) {
#define MPC_LANES_END }
#endif

namespace mpcgpu {

// parameter accessor over the per-solve transposed parameter block pt[npar][32] (lane = stage)
struct ParamT {
    const double *base;
    int k;
    MPC_HD double operator()(int i) const {
#if defined(__CUDA_ARCH__)
        return __ldg(base + i * 32 + k);
#else
        return base[i * 32 + k];
#endif
    }
};
// accessor over a plain per-stage parameter vector (debug / parity entry point)
struct ParamP {
    const double *p;
    MPC_HD double operator()(int i) const { return p[i]; }
};

// ---------------------------------------------------------------------------------------------
// Shared-memory image of one stage (offsets in doubles; odd stride => lane = stage accesses are
// bank-conflict free for 8-byte words)
// ---------------------------------------------------------------------------------------------
template <class CFG>
struct SL {
    static constexpr int NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NH = CFG::NH, NROW = CFG::NROW;
    static constexpr int NHP = NZ * (NZ + 1) / 2, NPP = NX * (NX + 1) / 2, NUP = NU * (NU + 1) / 2;
    static constexpr int Z = 0;                 // NLP iterate [u;x]
    static constexpr int PI = Z + NZ;           // NLP multiplier of the equation defining x_k
    static constexpr int LAM = PI + NX;         // inequality multipliers (IPM variable = NLP multiplier after the QP)
    static constexpr int T = LAM + NROW;        // IPM slacks
    static constexpr int H = T + NROW;          // regularised Hessian block, packed lower
    static constexpr int GR = H + NHP;          // QP gradient
    static constexpr int BA = GR + NZ;          // [B A], row-major NX x NZ
    static constexpr int C = BA + NX * NZ;      // Phi(x,u) - x_next
    static constexpr int JR = C + NX;           // sign-folded rows of dh/d(x,y,psi), NH x 3
    static constexpr int RH = JR + 3 * NH;      // sign-folded h residual s (h - bound)
    static constexpr int CORR = RH + NH;        // Mehrotra second-order term dt_aff * dlam_aff
    static constexpr int GD = CORR + NROW;      // barrier Hessian: diagonal part
    static constexpr int GB = GD + NZ;          //   (x,y,psi) block, packed lower 6
    static constexpr int GS = GB + 6;           //   slack cross terms (x,y,psi,slack)
    static constexpr int GT = GS + 4;           // barrier-modified gradient
    static constexpr int P = GT + NZ;           // Riccati cost-to-go matrix, packed lower
    static constexpr int PC = P + NPP;          // P_k c_{k-1}
    static constexpr int PV = PC + NX;          // Riccati cost-to-go vector
    static constexpr int MV = PV + NX;          // m = gt + BA'(P c + p) of this stage
    static constexpr int LXU = MV + NZ;         // L_xu = M_xu L_uu^-T, NX x NU
    static constexpr int LUU = LXU + NU * NX;   // Cholesky factor of M_uu as {1/l00, l10, 1/l11}
    static constexpr int DZ = LUU + NUP;        // QP primal iterate
    static constexpr int DPI = DZ + NZ;         // QP equality multiplier iterate
    static constexpr int DZN = DPI + NX;        // Newton target (primal)
    static constexpr int PIN = DZN + NZ;        // Newton target (equality multiplier)
    static constexpr int RAW = PIN + NX;
    static constexpr int STRIDE = RAW | 1;      // odd
    // per-warp scratch after the (N+1) stages
    static constexpr int W_M = 0;               // NHP: M matrix of the stage being factorised
    static constexpr int W_RED = W_M + NHP;     // 6 x 32 reduction staging
    static constexpr int W_SC = W_RED + 6 * 32; // scalars
    static constexpr int W_DX0 = W_SC + 16;     // NX
    static constexpr int W_SIZE = W_DX0 + NX;
    static constexpr int TOTAL = (CFG::N + 1) * STRIDE + W_SIZE;  // doubles per warp
};
// scalar slots in W_SC
enum { SC_MU = 0, SC_SIGMU, SC_ALPHA, SC_RS, SC_RE, SC_RI, SC_RC, SC_FLAG, SC_POBJ, SC_RESEQ };

template <class CFG>
MPC_HD bool box_active(const Layout &L, int k, int i, bool upper) {
    if (k == 0 && i >= CFG::NU) return false;  // x0 is fixed: no state box at node 0
    if (k >= CFG::N) return false;             // no bounds at the terminal node
    return upper ? (L.upper_bound[i] < 1e10) : (L.lower_bound[i] > -1e10);
}
template <class CFG>
MPC_HD constexpr bool row_has_slack(int j) {
    return CFG::kind_uses_slack(j < CFG::CNT0 ? CFG::KIND0 : (j < CFG::CNT0 + CFG::CNT1 ? CFG::KIND1 : CFG::KIND2));
}

MPC_HD double mpc_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// NaN-propagating running max
MPC_HD void upd_max(double &acc, double v) {
    if (acc != acc) return;
    if (v != v || v > acc) acc = v;
}

// ---------------------------------------------------------------------------------------------
// Linearise stage k (one lane).  Writes H (MIRROR-regularised), GR, BA, C, JR, RH of stage k.
// ---------------------------------------------------------------------------------------------
template <class CFG, class PA>
MPC_HD void linearize_stage(const PA &pa, const Layout &L, double *st, const double *st_next, int k) {
    typedef SL<CFG> S;
    constexpr int NZ = CFG::NZ, NX = CFG::NX, NH = CFG::NH;
    double z[NZ], Hp[S::NHP], g[NZ];
#pragma unroll
    for (int i = 0; i < NZ; i++) {
        z[i] = st[S::Z + i];
        g[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < S::NHP; i++) Hp[i] = 0.0;
    stage_cost<CFG>(pa, L, z, L.dt, g, Hp);
    {
        double xn[NX], pin[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) pin[i] = st_next[S::PI + i];
        unicycle_step<CFG>(L, z, xn, st + S::BA, pin, Hp);
        const int xoff = S::Z + CFG::NU;
#pragma unroll
        for (int i = 0; i < NX; i++) st[S::C + i] = xn[i] - st_next[xoff + i];
    }
    if (NH > 0 && k >= 1) {
        double Hx[6] = {0, 0, 0, 0, 0, 0};
        // raw h and dh/d(x,y,psi) go straight to shared memory, multipliers are read from it
        stage_constraints<CFG, true>(pa, L, z, st + S::RH, st + S::JR, st + S::LAM + 2 * NZ, Hx);
#pragma unroll
        for (int j = 0; j < NH; j++) {  // fold the bound side: row value s (h - bound) >= 0, row gradient s dh
            const int kind = row_kind<CFG>(j);
            const double sgn = row_sign<CFG>(kind);
            st[S::RH + j] = sgn * (st[S::RH + j] - row_bound<CFG>(kind));
            st[S::JR + 3 * j] *= sgn;
            st[S::JR + 3 * j + 1] *= sgn;
            st[S::JR + 3 * j + 2] *= sgn;
        }
        const int idx[3] = {CFG::IX, CFG::IY, CFG::IPSI};
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b <= a; b++) Hp[trisym(idx[a], idx[b])] += Hx[tri(a, b)];
    }
    mirror_block<NZ>(Hp, L.reg_epsilon);
#pragma unroll
    for (int i = 0; i < S::NHP; i++) st[S::H + i] = Hp[i];
#pragma unroll
    for (int i = 0; i < NZ; i++) st[S::GR + i] = g[i];
}

// value of inequality row r of stage k at primal point v (length NZ, node layout): r + Jr v
template <class CFG>
MPC_HD double hrow_value(const double *st, int j, const double *v) {
    typedef SL<CFG> S;
    double cv = st[S::RH + j] + st[S::JR + 3 * j] * v[CFG::IX] + st[S::JR + 3 * j + 1] * v[CFG::IY] + st[S::JR + 3 * j + 2] * v[CFG::IPSI];
    if (CFG::ANY_SLACK_ROWS && row_has_slack<CFG>(j)) cv += v[CFG::HAS_SLACK ? CFG::ISLACK : 0];
    return cv;
}

// reduce 32 staged values by lane 0 (tree order)
MPC_HD double red_max32(const double *r) {
    double m = 0.0;
    for (int i = 0; i < 32; i++) upd_max(m, r[i]);
    return m;
}
MPC_HD double red_min32(const double *r) {
    double m = r[0];
    for (int i = 1; i < 32; i++) m = (r[i] < m) ? r[i] : m;
    return m;
}
MPC_HD double red_sum32(const double *r) {
    double s = 0.0;
    for (int i = 0; i < 32; i++) s += r[i];
    return s;
}

// ---------------------------------------------------------------------------------------------
// Riccati sweeps (warp cooperative)
// ---------------------------------------------------------------------------------------------
// Backward factorisation: for k = N-1..0   M = H + Gamma-terms + [B A]' P+ [B A];
//   Luu = chol(Muu);  Lxu = Mxu Luu^-T;  P = Mxx - Lxu Lxu';  PC(k) = P_k c_{k-1}.
template <class CFG>
MPC_DEV void riccati_factor(double *sm) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU;
    double *W = sm + (N + 1) * S::STRIDE;
    // terminal: P_N = eps I is set once per solve; PC(N) = P_N c_{N-1}
    MPC_LANES
    if (lane < NX) {
        const double *stN = sm + N * S::STRIDE, *stp = sm + (N - 1) * S::STRIDE;
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < NX; m++) acc += stN[S::P + trisym(lane, m)] * stp[S::C + m];
        sm[N * S::STRIDE + S::PC + lane] = acc;
    }
    MPC_LANES_END
    for (int k = N - 1; k >= 0; k--) {
        double *st = sm + k * S::STRIDE;
        const double *stn = sm + (k + 1) * S::STRIDE;
        // R1: M entries (packed lower)
        MPC_LANES
        for (int e = lane; e < S::NHP; e += 32) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= e) i++;
            const int j = e - i * (i + 1) / 2;  // i >= j
            double acc = st[S::H + e];
            if (i == j) acc += st[S::GD + i];
            // barrier block over (x,y,psi) and slack cross terms
            {
                const int ai = (i == CFG::IX) ? 0 : (i == CFG::IY) ? 1 : (i == CFG::IPSI) ? 2 : -1;
                const int aj = (j == CFG::IX) ? 0 : (j == CFG::IY) ? 1 : (j == CFG::IPSI) ? 2 : -1;
                if (ai >= 0 && aj >= 0) acc += st[S::GB + trisym(ai, aj)];
                if (CFG::ANY_SLACK_ROWS) {
                    if (i == CFG::ISLACK && aj >= 0) acc += st[S::GS + aj];
                    if (i == CFG::ISLACK && j == CFG::ISLACK) acc += st[S::GS + 3];
                }
            }
#pragma unroll
            for (int l = 0; l < NX; l++) {
                double wl = 0.0;  // (P+ BA)[l][j]
#pragma unroll
                for (int m = 0; m < NX; m++) wl += stn[S::P + trisym(l, m)] * st[S::BA + m * NZ + j];
                acc += st[S::BA + l * NZ + i] * wl;
            }
            W[S::W_M + e] = acc;
        }
        MPC_LANES_END
        // R2: Cholesky of M_uu (every lane, redundantly), L_xu = M_xu L_uu^-T, P = M_xx - L_xu L_xu'
        MPC_LANES
        {
            static_assert(NU == 2, "input block is factorised in closed form for nu = 2");
            const double m00 = W[S::W_M + tri(0, 0)], m10 = W[S::W_M + tri(1, 0)], m11 = W[S::W_M + tri(1, 1)];
            const double il00 = mpc_rsqrt(m00);
            const double l10 = m10 * il00;
            const double il11 = mpc_rsqrt(m11 - l10 * l10);
            // lanes [0, NPP): P entries ; lanes [NPP, NPP + NX): rows of L_xu ; next lane: L_uu
            for (int e = lane; e < S::NPP + NX + 1; e += 32) {
                if (e < S::NPP) {
                    int i = 0;
                    while ((i + 1) * (i + 2) / 2 <= e) i++;
                    const int j = e - i * (i + 1) / 2;
                    const double li0 = W[S::W_M + tri(NU + i, 0)] * il00, li1 = (W[S::W_M + tri(NU + i, 1)] - li0 * l10) * il11;
                    const double lj0 = W[S::W_M + tri(NU + j, 0)] * il00, lj1 = (W[S::W_M + tri(NU + j, 1)] - lj0 * l10) * il11;
                    st[S::P + e] = W[S::W_M + tri(NU + i, NU + j)] - li0 * lj0 - li1 * lj1;
                } else if (e < S::NPP + NX) {
                    const int i = e - S::NPP;
                    const double li0 = W[S::W_M + tri(NU + i, 0)] * il00, li1 = (W[S::W_M + tri(NU + i, 1)] - li0 * l10) * il11;
                    st[S::LXU + i * NU] = li0;
                    st[S::LXU + i * NU + 1] = li1;
                } else {
                    st[S::LUU] = il00;
                    st[S::LUU + 1] = l10;
                    st[S::LUU + 2] = il11;
                }
            }
        }
        MPC_LANES_END
        if (k >= 1) {
            MPC_LANES
            if (lane < NX) {
                const double *stp = sm + (k - 1) * S::STRIDE;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < NX; m++) acc += st[S::P + trisym(lane, m)] * stp[S::C + m];
                st[S::PC + lane] = acc;
            }
            MPC_LANES_END
        }
    }
}

// Backward vector sweep + forward sweep for the gradient GT:  writes DZN (all nodes) and PIN.
template <class CFG>
MPC_DEV void riccati_solve(double *sm) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU;
    double *W = sm + (N + 1) * S::STRIDE;
    // terminal: pv_N = gt_N (x part lives at [NU, NZ) of the node layout)
    MPC_LANES
    if (lane < NX) sm[N * S::STRIDE + S::PV + lane] = sm[N * S::STRIDE + S::GT + NU + lane];
    MPC_LANES_END
    for (int k = N - 1; k >= 0; k--) {
        double *st = sm + k * S::STRIDE;
        double *stn = sm + (k + 1) * S::STRIDE;
        // stage k+1's pv was produced by the previous phase (or the terminal init)
        MPC_LANES
        if (lane < NZ) {
            double acc = st[S::GT + lane];
#pragma unroll
            for (int l = 0; l < NX; l++) acc += st[S::BA + l * NZ + lane] * (stn[S::PV + l] + stn[S::PC + l]);
            st[S::MV + lane] = acc;
        }
        MPC_LANES_END
        MPC_LANES
        if (lane < NX) {  // pv = m_x - L_xu y,  y = L_uu^-1 m_u
            const double y0 = st[S::MV] * st[S::LUU];
            const double y1 = (st[S::MV + 1] - st[S::LUU + 1] * y0) * st[S::LUU + 2];
            st[S::PV + lane] = st[S::MV + NU + lane] - st[S::LXU + lane * NU] * y0 - st[S::LXU + lane * NU + 1] * y1;
        }
        MPC_LANES_END
    }
    // forward
    MPC_LANES
    if (lane < NX) sm[S::DZN + NU + lane] = W[S::W_DX0 + lane];
    MPC_LANES_END
    for (int k = 0; k < N; k++) {
        double *st = sm + k * S::STRIDE;
        double *stn = sm + (k + 1) * S::STRIDE;
        MPC_LANES
        if (lane < NU) {  // du = -L_uu^-T (L_uu^-1 m_u + L_xu' dx)
            const double il00 = st[S::LUU], l10 = st[S::LUU + 1], il11 = st[S::LUU + 2];
            double w0 = st[S::MV] * il00;
            double w1 = (st[S::MV + 1] - l10 * w0) * il11;
#pragma unroll
            for (int l = 0; l < NX; l++) {
                w0 += st[S::LXU + l * NU] * st[S::DZN + NU + l];
                w1 += st[S::LXU + l * NU + 1] * st[S::DZN + NU + l];
            }
            const double k1 = w1 * il11;
            const double k0 = (w0 - l10 * k1) * il00;
            st[S::DZN + lane] = (lane == 0) ? -k0 : -k1;
        }
        MPC_LANES_END
        MPC_LANES
        if (lane < NX) {
            double acc = st[S::C + lane];
#pragma unroll
            for (int j = 0; j < NZ; j++) acc += st[S::BA + lane * NZ + j] * st[S::DZN + j];
            stn[S::DZN + NU + lane] = acc;
        }
        MPC_LANES_END
        MPC_LANES
        if (lane < NX) {  // pi+ = P+ dx+ + pv+
            double acc = stn[S::PV + lane];
#pragma unroll
            for (int m = 0; m < NX; m++) acc += stn[S::P + trisym(lane, m)] * stn[S::DZN + NU + m];
            stn[S::PIN + lane] = acc;
        }
        MPC_LANES_END
    }
}

// ---------------------------------------------------------------------------------------------
// Interior-point QP solve.  Returns the QP status (acados numbering: 0 ok, 2 max iter,
// 3 min step, 4 NaN) -- uniform across the warp.  Iteration count is added to *qp_iter_total.
// ---------------------------------------------------------------------------------------------
template <class CFG>
MPC_DEV int ipm_solve(const Layout &L, double *sm, int *iters_out) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NH = CFG::NH;
    double *W = sm + (N + 1) * S::STRIDE;
    const double T_MIN = 1e-10, LAM_MIN = 1e-16, ALPHA_MIN = 1e-12;
    // number of inequality rows in the QP (uniform)
    int nc = 0;
    for (int i = 0; i < NZ; i++) {
        nc += (L.lower_bound[i] > -1e10) ? ((i < NU) ? N : N - 1) : 0;
        nc += (L.upper_bound[i] < 1e10) ? ((i < NU) ? N : N - 1) : 0;
    }
    nc += NH * (N - 1);
    // cold start (HPIPM-style): dz = 0, pi = 0, t = max(r, thr0), lam = mu0 / t
    MPC_LANES
    if (lane <= N) {
        double *st = sm + lane * S::STRIDE;
        const int k = lane;
#pragma unroll
        for (int i = 0; i < NZ; i++) st[S::DZ + i] = (k == 0 && i >= NU) ? W[S::W_DX0 + i - NU] : 0.0;
#pragma unroll
        for (int i = 0; i < NX; i++) st[S::DPI + i] = 0.0;
#pragma unroll
        for (int i = 0; i < NZ; i++) {
            const double zi = st[S::Z + i];
            {
                const bool act = box_active<CFG>(L, k, i, false);
                const double tt = fmax(zi - L.lower_bound[i], L.qp_thr0);
                st[S::T + i] = act ? tt : 1.0;
                st[S::LAM + i] = act ? L.qp_mu0 / tt : 0.0;
                st[S::CORR + i] = 0.0;
            }
            {
                const bool act = box_active<CFG>(L, k, i, true);
                const double tt = fmax(L.upper_bound[i] - zi, L.qp_thr0);
                st[S::T + NZ + i] = act ? tt : 1.0;
                st[S::LAM + NZ + i] = act ? L.qp_mu0 / tt : 0.0;
                st[S::CORR + NZ + i] = 0.0;
            }
        }
        const bool hact = (k >= 1 && k < N);
#pragma unroll
        for (int j = 0; j < NH; j++) {
            const double tt = fmax(hact ? st[S::RH + j] : 1.0, L.qp_thr0);
            st[S::T + 2 * NZ + j] = hact ? tt : 1.0;
            st[S::LAM + 2 * NZ + j] = hact ? L.qp_mu0 / tt : 0.0;
            st[S::CORR + 2 * NZ + j] = 0.0;
        }
    }
    MPC_LANES_END

    int status = 0, iter = 0;
    for (;; iter++) {
        // ---- residuals of the current iterate (lane = node) ----
        MPC_LANES
        {
            double rs = 0.0, re = 0.0, ri = 0.0, rc = 0.0, musum = 0.0;
            if (lane <= N) {
                const int k = lane;
                const double *st = sm + k * S::STRIDE;
                double dz[NZ], rg[NZ];
#pragma unroll
                for (int i = 0; i < NZ; i++) dz[i] = st[S::DZ + i];
                if (k < N) {
#pragma unroll
                    for (int i = 0; i < NZ; i++) {
                        double acc = st[S::GR + i];
#pragma unroll
                        for (int j = 0; j < NZ; j++) acc += st[S::H + trisym(i, j)] * dz[j];
                        rg[i] = acc;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NZ; i++) rg[i] = (i >= NU) ? L.reg_epsilon * dz[i] : 0.0;  // H_N = eps I, g_N = 0
                }
#pragma unroll
                for (int i = 0; i < NZ; i++) {
                    if (box_active<CFG>(L, k, i, false)) {
                        const double l = st[S::LAM + i], tt = st[S::T + i];
                        const double cv = st[S::Z + i] - L.lower_bound[i] + dz[i];
                        rg[i] -= l;
                        upd_max(ri, fabs(cv - tt));
                        upd_max(rc, fabs(l * tt));
                        musum += l * tt;
                    }
                    if (box_active<CFG>(L, k, i, true)) {
                        const double l = st[S::LAM + NZ + i], tt = st[S::T + NZ + i];
                        const double cv = L.upper_bound[i] - st[S::Z + i] - dz[i];
                        rg[i] += l;
                        upd_max(ri, fabs(cv - tt));
                        upd_max(rc, fabs(l * tt));
                        musum += l * tt;
                    }
                }
                if (NH > 0 && k >= 1 && k < N) {
#pragma unroll
                    for (int j = 0; j < NH; j++) {
                        const double l = st[S::LAM + 2 * NZ + j], tt = st[S::T + 2 * NZ + j];
                        const double cv = hrow_value<CFG>(st, j, dz);
                        rg[CFG::IX] -= st[S::JR + 3 * j] * l;
                        rg[CFG::IY] -= st[S::JR + 3 * j + 1] * l;
                        rg[CFG::IPSI] -= st[S::JR + 3 * j + 2] * l;
                        if (CFG::ANY_SLACK_ROWS && row_has_slack<CFG>(j)) rg[CFG::HAS_SLACK ? CFG::ISLACK : 0] -= l;
                        upd_max(ri, fabs(cv - tt));
                        upd_max(rc, fabs(l * tt));
                        musum += l * tt;
                    }
                }
                if (k < N) {
                    const double *stn = sm + (k + 1) * S::STRIDE;
#pragma unroll
                    for (int i = 0; i < NZ; i++)
#pragma unroll
                        for (int l = 0; l < NX; l++) rg[i] += st[S::BA + l * NZ + i] * stn[S::DPI + l];
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        double acc = st[S::C + i] - stn[S::DZ + NU + i];
#pragma unroll
                        for (int j = 0; j < NZ; j++) acc += st[S::BA + i * NZ + j] * dz[j];
                        upd_max(re, fabs(acc));
                    }
                }
                if (k >= 1) {
#pragma unroll
                    for (int i = 0; i < NX; i++) rg[NU + i] -= st[S::DPI + i];
                }
#pragma unroll
                for (int i = 0; i < NZ; i++) {
                    if (k == 0 && i >= NU) continue;  // x0 eliminated
                    if (k == N && i < NU) continue;
                    upd_max(rs, fabs(rg[i]));
                }
            }
            W[S::W_RED + 0 * 32 + lane] = rs;
            W[S::W_RED + 1 * 32 + lane] = re;
            W[S::W_RED + 2 * 32 + lane] = ri;
            W[S::W_RED + 3 * 32 + lane] = rc;
            W[S::W_RED + 4 * 32 + lane] = musum;
        }
        MPC_LANES_END
        MPC_LANES
        if (lane < 5) {
            const double v = (lane == 4) ? red_sum32(W + S::W_RED + 4 * 32) : red_max32(W + S::W_RED + lane * 32);
            if (lane == 4) W[S::W_SC + SC_MU] = (nc > 0) ? v / nc : 0.0;
            else W[S::W_SC + SC_RS + lane] = v;
        }
        MPC_LANES_END
        {
            const double rs = W[S::W_SC + SC_RS], re = W[S::W_SC + SC_RE], ri = W[S::W_SC + SC_RI], rc = W[S::W_SC + SC_RC];
#if !defined(__CUDACC__) && defined(MPC_EMU_TRACE)
            if (getenv("EMU_DEBUG")) fprintf(stderr, "  ipm %2d: stat %.3e eq %.3e ineq %.3e comp %.3e mu %.3e\n", iter, rs, re, ri, rc, W[S::W_SC + SC_MU]);
#endif
            if (!(rs == rs) || !(re == re) || !(ri == ri) || !(rc == rc)) { status = 4; break; }
            if (rs < L.qp_tol && re < L.qp_tol && ri < L.qp_tol && rc < L.qp_tol) { status = 0; break; }
            if (iter >= L.qp_iter_max) { status = 2; break; }
        }
        const double mu = W[S::W_SC + SC_MU];

        double alpha = 1.0;
        for (int pass = 0; pass < 2; pass++) {
            const double sigma_mu = (pass == 0) ? 0.0 : W[S::W_SC + SC_SIGMU];
            // ---- barrier terms: Gamma = lam/t ; GD/GB/GS (pass 0) and modified gradient GT ----
            MPC_LANES
            if (lane <= N) {
                const int k = lane;
                double *st = sm + k * S::STRIDE;
                double gt[NZ], gd[NZ], gb[6] = {0, 0, 0, 0, 0, 0}, gs[4] = {0, 0, 0, 0};
#pragma unroll
                for (int i = 0; i < NZ; i++) {
                    gt[i] = (k < N) ? st[S::GR + i] : 0.0;
                    gd[i] = 0.0;
                }
#pragma unroll
                for (int i = 0; i < NZ; i++) {
                    if (box_active<CFG>(L, k, i, false)) {
                        const double l = st[S::LAM + i], tt = st[S::T + i], gam = l / tt;
                        const double r = st[S::Z + i] - L.lower_bound[i];
                        gd[i] += gam;
                        gt[i] -= l + (sigma_mu - st[S::CORR + i]) / tt - gam * r;
                    }
                    if (box_active<CFG>(L, k, i, true)) {
                        const double l = st[S::LAM + NZ + i], tt = st[S::T + NZ + i], gam = l / tt;
                        const double r = L.upper_bound[i] - st[S::Z + i];
                        gd[i] += gam;
                        gt[i] += l + (sigma_mu - st[S::CORR + NZ + i]) / tt - gam * r;
                    }
                }
                if (NH > 0 && k >= 1 && k < N) {
#pragma unroll
                    for (int j = 0; j < NH; j++) {
                        const double l = st[S::LAM + 2 * NZ + j], tt = st[S::T + 2 * NZ + j], gam = l / tt;
                        const double coef = l + (sigma_mu - st[S::CORR + 2 * NZ + j]) / tt - gam * st[S::RH + j];
                        const double j0 = st[S::JR + 3 * j], j1 = st[S::JR + 3 * j + 1], j2 = st[S::JR + 3 * j + 2];
                        gt[CFG::IX] -= j0 * coef;
                        gt[CFG::IY] -= j1 * coef;
                        gt[CFG::IPSI] -= j2 * coef;
                        gb[0] += gam * j0 * j0;
                        gb[1] += gam * j1 * j0;
                        gb[2] += gam * j1 * j1;
                        gb[3] += gam * j2 * j0;
                        gb[4] += gam * j2 * j1;
                        gb[5] += gam * j2 * j2;
                        if (CFG::ANY_SLACK_ROWS && row_has_slack<CFG>(j)) {
                            gt[CFG::HAS_SLACK ? CFG::ISLACK : 0] -= coef;
                            gs[0] += gam * j0;
                            gs[1] += gam * j1;
                            gs[2] += gam * j2;
                            gs[3] += gam;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NZ; i++) st[S::GT + i] = gt[i];
                if (pass == 0) {
#pragma unroll
                    for (int i = 0; i < NZ; i++) st[S::GD + i] = gd[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) st[S::GB + i] = gb[i];
#pragma unroll
                    for (int i = 0; i < 4; i++) st[S::GS + i] = gs[i];
                }
            }
            MPC_LANES_END
            if (pass == 0) riccati_factor<CFG>(sm);
            riccati_solve<CFG>(sm);
            // ---- step length to the boundary ----
            MPC_LANES
            {
                double amax = 1.0;
                if (lane < N) {
                    const int k = lane;
                    const double *st = sm + k * S::STRIDE;
                    double dzn[NZ];
#pragma unroll
                    for (int i = 0; i < NZ; i++) dzn[i] = st[S::DZN + i];
#pragma unroll
                    for (int r = 0; r < CFG::NROW; r++) {
                        double cv;
                        bool act;
                        if (r < NZ) {
                            act = box_active<CFG>(L, k, r, false);
                            cv = st[S::Z + r] - L.lower_bound[r] + dzn[r];
                        } else if (r < 2 * NZ) {
                            act = box_active<CFG>(L, k, r - NZ, true);
                            cv = L.upper_bound[r - NZ] - st[S::Z + r - NZ] - dzn[r - NZ];
                        } else {
                            act = (k >= 1);
                            cv = act ? hrow_value<CFG>(st, r - 2 * NZ, dzn) : 0.0;
                        }
                        if (!act) continue;
                        const double l = st[S::LAM + r], tt = st[S::T + r];
                        const double dt = cv - tt;
                        const double dl = (sigma_mu - st[S::CORR + r]) / tt - (l / tt) * cv;
                        if (dt < 0.0) amax = fmin(amax, -tt / dt);
                        if (dl < 0.0) amax = fmin(amax, -l / dl);
                    }
                }
                W[S::W_RED + lane] = amax;
            }
            MPC_LANES_END
            MPC_LANES
            if (lane == 0) W[S::W_SC + SC_ALPHA] = red_min32(W + S::W_RED);
            MPC_LANES_END
            alpha = W[S::W_SC + SC_ALPHA];
            if (pass == 0) {
                if (nc == 0) break;
                // ---- centering parameter and second-order correction ----
                MPC_LANES
                {
                    double msum = 0.0;
                    if (lane < N) {
                        const int k = lane;
                        double *st = sm + k * S::STRIDE;
                        double dzn[NZ];
#pragma unroll
                        for (int i = 0; i < NZ; i++) dzn[i] = st[S::DZN + i];
#pragma unroll
                        for (int r = 0; r < CFG::NROW; r++) {
                            double cv;
                            bool act;
                            if (r < NZ) {
                                act = box_active<CFG>(L, k, r, false);
                                cv = st[S::Z + r] - L.lower_bound[r] + dzn[r];
                            } else if (r < 2 * NZ) {
                                act = box_active<CFG>(L, k, r - NZ, true);
                                cv = L.upper_bound[r - NZ] - st[S::Z + r - NZ] - dzn[r - NZ];
                            } else {
                                act = (k >= 1);
                                cv = act ? hrow_value<CFG>(st, r - 2 * NZ, dzn) : 0.0;
                            }
                            if (!act) continue;
                            const double l = st[S::LAM + r], tt = st[S::T + r];
                            const double dt = cv - tt;
                            const double dl = -(l / tt) * cv;  // sigma_mu = corr = 0 in the affine pass
                            msum += (l + alpha * dl) * (tt + alpha * dt);
                            st[S::CORR + r] = dt * dl;
                        }
                    }
                    W[S::W_RED + lane] = msum;
                }
                MPC_LANES_END
                MPC_LANES
                if (lane == 0) {
                    const double mu_aff = red_sum32(W + S::W_RED) / nc;
                    const double ratio = (mu > 0.0) ? mu_aff / mu : 0.0;
                    // Mehrotra centering, floored at a fraction of the complementarity tolerance: driving mu
                    // far below qp_tol only makes lam/t huge and buries the stationarity residual in round-off
                    W[S::W_SC + SC_SIGMU] = fmax(ratio * ratio * ratio * mu, 0.1 * L.qp_tol);
                }
                MPC_LANES_END
            }
        }
        const double sigma_mu = (nc > 0) ? W[S::W_SC + SC_SIGMU] : 0.0;
        alpha *= 0.995;  // fraction to the boundary; never a full step
        if (!(sigma_mu == sigma_mu) || !(alpha == alpha)) { status = 4; break; }
        if (alpha < ALPHA_MIN) { status = 3; break; }
        // ---- update ----
        MPC_LANES
        if (lane <= N) {
            const int k = lane;
            double *st = sm + k * S::STRIDE;
            double dzn[NZ];
#pragma unroll
            for (int i = 0; i < NZ; i++) dzn[i] = st[S::DZN + i];
            if (k < N) {
#pragma unroll
                for (int r = 0; r < CFG::NROW; r++) {
                    double cv;
                    bool act;
                    if (r < NZ) {
                        act = box_active<CFG>(L, k, r, false);
                        cv = st[S::Z + r] - L.lower_bound[r] + dzn[r];
                    } else if (r < 2 * NZ) {
                        act = box_active<CFG>(L, k, r - NZ, true);
                        cv = L.upper_bound[r - NZ] - st[S::Z + r - NZ] - dzn[r - NZ];
                    } else {
                        act = (k >= 1);
                        cv = act ? hrow_value<CFG>(st, r - 2 * NZ, dzn) : 0.0;
                    }
                    if (act) {
                        const double l = st[S::LAM + r], tt = st[S::T + r];
                        const double dt = cv - tt;
                        const double dl = (sigma_mu - st[S::CORR + r]) / tt - (l / tt) * cv;
                        st[S::T + r] = fmax(tt + alpha * dt, T_MIN);
                        st[S::LAM + r] = fmax(l + alpha * dl, LAM_MIN);
                    }
                    st[S::CORR + r] = 0.0;
                }
            }
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                if (k == 0 && i >= NU) continue;
                if (k == N && i < NU) continue;
                st[S::DZ + i] += alpha * (dzn[i] - st[S::DZ + i]);
            }
            if (k >= 1) {
#pragma unroll
                for (int i = 0; i < NX; i++) st[S::DPI + i] += alpha * (st[S::PIN + i] - st[S::DPI + i]);
            }
        }
        MPC_LANES_END
    }
    *iters_out = iter;
    return status;
}

// ---------------------------------------------------------------------------------------------
// The whole solve.  sm: SL<CFG>::TOTAL doubles of warp-private shared memory.
//   xinit[NX], x0[(N+1)*NZ] (reference layout), pt = parameters transposed to [npar][32]
//   xtraj[(N+1)*NX], utraj[N*NU], info; iterates (may be null): [n_sqp][N+1][NZ]
// ---------------------------------------------------------------------------------------------
template <class CFG>
MPC_DEV void sqp_solve_warp(const Layout &L, double *sm, const double *xinit, const double *x0, const double *pt,
                            double *xtraj, double *utraj, mpcgpu_info *info, double *iterates) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU;
    double *W = sm + (N + 1) * S::STRIDE;
    // loadWarmstart (acados_solver_interface.cpp:274-284): x (k = 0..N) and u (k < N) from x0;
    // fresh solver: zero multipliers
    MPC_LANES
    if (lane <= N) {
        double *st = sm + lane * S::STRIDE;
        for (int i = 0; i < S::STRIDE; i++) st[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NZ; i++) st[S::Z + i] = (lane == N && i < NU) ? 0.0 : x0[lane * NZ + i];
        if (lane == N) {
#pragma unroll
            for (int i = 0; i < NX; i++) st[S::P + tri(i, i)] = L.reg_epsilon;  // MIRROR of the empty terminal block
        }
    }
    MPC_LANES_END

    int qp_status = 0, sqp_iter = 0, qp_iter_total = 0;
    for (int it = 0; it < L.n_sqp; it++) {
        // preparation: linearise every stage at the current iterate
        MPC_LANES
        if (lane < N) {
            ParamT pa{pt, lane};
            linearize_stage<CFG>(pa, L, sm + lane * S::STRIDE, sm + (lane + 1) * S::STRIDE, lane);
        }
        if (lane >= 32 - NX) {
            const int i = lane - (32 - NX);
            W[S::W_DX0 + i] = xinit[i] - sm[S::Z + NU + i];
        }
        MPC_LANES_END
        // feedback: QP
        int iters = 0;
        qp_status = ipm_solve<CFG>(L, sm, &iters);
        qp_iter_total += iters;
        sqp_iter = it + 1;
        if (qp_status == 3 || qp_status == 4) break;  // QP failure: no step (acados sqp_rti)
        // full step; multipliers <- QP multipliers (LAM already holds them)
        MPC_LANES
        if (lane <= N) {
            double *st = sm + lane * S::STRIDE;
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                if (lane == N && i < NU) continue;
                st[S::Z + i] += st[S::DZ + i];
            }
            if (lane >= 1) {
#pragma unroll
                for (int i = 0; i < NX; i++) st[S::PI + i] = st[S::DPI + i];
            }
            if (iterates) {
#pragma unroll
                for (int i = 0; i < NZ; i++) iterates[(it * (N + 1) + lane) * NZ + i] = (lane < N) ? st[S::Z + i] : ((i < NX) ? st[S::Z + NU + i] : 0.0);
            }
        }
        MPC_LANES_END
        if (qp_status != 0) break;  // Q4: the reference loop stops on any non-zero qp_status
    }

    // completeOneIteration (acados_solver_interface.cpp:162-204): cost, trajectories, res_eq
    MPC_LANES
    {
        double c = 0.0, defect = 0.0;
        if (lane < N) {
            ParamT pa{pt, lane};
            const double *st = sm + lane * S::STRIDE, *stn = sm + (lane + 1) * S::STRIDE;
            double z[NZ], xn[NX];
#pragma unroll
            for (int i = 0; i < NZ; i++) z[i] = st[S::Z + i];
            c = L.dt * stage_cost<CFG>(pa, L, z, 0.0, nullptr, nullptr);
            unicycle_step<CFG>(L, z, xn, nullptr, nullptr, nullptr);
#pragma unroll
            for (int i = 0; i < NX; i++) upd_max(defect, fabs(xn[i] - stn[S::Z + NU + i]));
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < NX; i++) upd_max(defect, fabs(z[NU + i] - xinit[i]));
            }
#pragma unroll
            for (int i = 0; i < NU; i++) utraj[lane * NU + i] = z[i];
        }
        if (lane <= N) {
#pragma unroll
            for (int i = 0; i < NX; i++) xtraj[lane * NX + i] = sm[lane * S::STRIDE + S::Z + NU + i];
        }
        W[S::W_RED + lane] = c;
        W[S::W_RED + 32 + lane] = defect;
    }
    MPC_LANES_END
    MPC_LANES
    if (lane == 0) {
        const double pobj = red_sum32(W + S::W_RED), res_eq = red_max32(W + S::W_RED + 32);
        int exit_code = 1;
        if (qp_status == 3 || qp_status == 4) exit_code = 4;  // ACADOS_QP_FAILURE
        if (!(res_eq <= L.res_eq_fail) && exit_code == 1) exit_code = 4;  // acados_solver_interface.cpp:176-181
        info->pobj = pobj;
        info->res_eq = res_eq;
        info->res_stat_qp = W[S::W_SC + SC_RS];
        info->res_comp_qp = W[S::W_SC + SC_RC];
        info->exit_code = exit_code;
        info->qp_status = qp_status;
        info->sqp_iter = sqp_iter;
        info->qp_iter = qp_iter_total;
    }
    MPC_LANES_END
}

}  // namespace mpcgpu
