// sqp_core.cuh -- one full `Solver::solve()` (mpc_planner_solver/src/acados_solver_interface.cpp:86-204)
// executed by ONE CTA of NT threads with the whole problem resident in shared memory.
//
// Replaces, for the hot path, what the reference delegates to acados (SQP-RTI), HPIPM (IPM QP) and
// BLASFEO: per RTI iteration  linearise all stages (threads = stage x function part)  ->
// MIRROR-regularise  ->  Mehrotra predictor-corrector interior point on the stage-structured QP
// (threads = stage x inequality row) with a Riccati factorisation (device: one warp on the FP64 tensor-core path,
// riccati_factor_mma; host emulation: one thread) and single-thread vector sweeps through the stored feedback gain
// (DESIGN.md 4 lists the variants that were built and measured)  ->  full step.
//
// Structure: the routine is a sequence of CTA "phases".  Inside a phase every thread works on data
// written by earlier phases and writes a disjoint part of shared memory; phases are separated by
// __syncthreads().  No register state crosses a phase boundary (exceptions: thread 0's timeout clocks and CTA-uniform
// scalars), so the same source compiles for the host as a sequential NT-thread emulation (tests/emu), used to debug
// without a GPU.
#pragma once

#include "stage_functions.cuh"

#if defined(__CUDACC__)
#define MPC_DEV __device__ __forceinline__
#define MPC_PHASE \
    {             \
        const int tid = threadIdx.x;
#define MPC_PHASE_END \
    }                 \
    __syncthreads();
// phase end that also ORs a per-thread flag over the CTA (the emulation's flag is one variable shared by the tid loop)
#define MPC_PHASE_END_OR(flag) \
    }                          \
    flag = __syncthreads_or(flag);
#else
#define MPC_DEV inline
#define MPC_PHASE for (int tid = 0; tid < MPC_NT; ++tid) {
#define MPC_PHASE_END }
#define MPC_PHASE_END_OR(flag) }
#endif

// Development aid (-DMPC_PROFILE, scripts/profile_phases.py): thread 0 accumulates clock64() per phase
// category in shared memory and the kernel writes the totals over the first doubles of utraj.
#if defined(MPC_TICK)
// supplied by the including translation unit (tests/emu/emu_count.cpp: FLOP counters per phase category)
#elif defined(MPC_PROFILE) && defined(__CUDA_ARCH__)
#define MPC_TICK(cat)                                                  \
    if (threadIdx.x == 0) {                                            \
        const long long now__ = clock64();                             \
        mpc_prof[cat] += (double)(now__ - (long long)mpc_prof[31]);    \
        mpc_prof[31] = (double)now__;                                  \
    }
#else
#define MPC_TICK(cat)
#endif

namespace mpcgpu {

#if defined(MPC_NT_OVERRIDE)
constexpr int MPC_NT = MPC_NT_OVERRIDE;
#else
constexpr int MPC_NT = 128;  // threads per solve
#endif
// partials per reduced quantity: one per warp on the device (threads pre-reduce with shuffles),
// one per thread in the host emulation
#if defined(__CUDACC__)
constexpr int MPC_RED_SLOTS = MPC_NT / 32;
#else
constexpr int MPC_RED_SLOTS = MPC_NT;
#endif
enum { TK_LIN = 0, TK_ASM, TK_INIT, TK_PA, TK_PB, TK_ACC0, TK_RED, TK_FACTOR, TK_SOLVE, TK_PD, TK_ACC1, TK_PF, TK_STEP, TK_FINAL };

// parameter accessor over the per-solve transposed parameter block pt[npar][PT] (column = stage, PT = pt_stride(N))
template <int PT>
struct ParamT {
    const double *base;
    int k;
    MPC_HD double operator()(int i) const {
#if defined(__CUDA_ARCH__)
        return __ldg(base + i * PT + k);
#else
        return base[i * PT + k];
#endif
    }
};
// accessor over a plain per-stage parameter vector (debug / parity entry point)
struct ParamP {
    const double *p;
    MPC_HD double operator()(int i) const { return p[i]; }
};



// NaN-propagating running max of values that are >= +0 or NaN (every caller passes fabs(.) and starts from 0): on the
// device an unsigned integer max of the bit patterns (monotone for such values, NaNs sort above +inf)
MPC_HD void upd_max(double &acc, double v) {
#if defined(__CUDA_ARCH__)
    const unsigned long long a = (unsigned long long)__double_as_longlong(acc), b = (unsigned long long)__double_as_longlong(v);
    acc = __longlong_as_double((long long)(b > a ? b : a));
#else
    if (acc != acc) return;
    if (v != v || v > acc) acc = v;
#endif
}

// Nanosecond clock for the solver timeout: %globaltimer on the device.  The host emulation uses a deterministic fake clock
// (tests/emu: every read advances it by emu_clock_step_ns) so that the CPU suite can pin the iteration-count logic.
#if !defined(__CUDACC__)
extern long long g_emu_clock_ns, g_emu_clock_step_ns;
extern long g_mirror_general;  // stage blocks that went through the CTA-parallel Jacobi (MPC_EMU_TRACE builds)
#endif
MPC_DEV long long mpc_clock_ns() {
#if defined(__CUDA_ARCH__)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (long long)t;
#elif !defined(__CUDACC__)
    g_emu_clock_ns += g_emu_clock_step_ns;
    return g_emu_clock_ns;
#else
    return 0;
#endif
}

// ---------------------------------------------------------------------------------------------
// Shared-memory image of one stage (offsets in doubles; odd stride => "thread = stage" accesses
// are bank-conflict free for 8-byte words)
// ---------------------------------------------------------------------------------------------
template <class CFG>
struct SL {
    static constexpr int NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NH = CFG::NH, NROW = CFG::NROW;
    static constexpr int NHP = NZ * (NZ + 1) / 2, NPP = NX * (NX + 1) / 2;
    static constexpr int NE = CFG::NE;          // columns of A that differ from the identity: (psi, v) / (psi, v, delta)
    static constexpr int NLUU = NU * (NU + 1) / 2;  // Cholesky factor of M_uu: packed lower, RECIPROCAL diagonal
    static constexpr int Z = 0;                 // NLP iterate [u;x]
    static constexpr int LAM = Z + NZ;          // inequality multipliers (IPM variable = NLP multiplier after the QP)
    static constexpr int T = LAM + NROW;        // IPM slacks
    static constexpr int TINV = T + NROW;       // 1/t
    static constexpr int CORR = TINV + NROW;    // Mehrotra second-order term dt_aff * dlam_aff
    static constexpr int H = CORR + NROW;       // regularised Hessian block, packed lower
    static constexpr int GR = H + NHP;          // QP gradient
    static constexpr int BB = GR + NZ;          // B, row-major NX x NU
    static constexpr int BE = BB + NX * NU;     // E = A - I, its non-zero columns (psi, v), row-major NX x NE
    static constexpr int C = BE + NX * NE;      // Phi(x,u) - x_next
    static constexpr int JR = C + NX;           // sign-folded rows of dh/d(x,y,psi), NH x 3
    static constexpr int RH = JR + 3 * NH;      // sign-folded h residual s (h - bound) + relaxation
    // ---- IPM / Riccati work area (also used as scratch by the linearisation) ----
    static constexpr int WK = RH + NH;
    static constexpr int GD = WK;               // barrier Hessian: diagonal part
    static constexpr int GB = GD + NZ;          //   (x,y,psi) block, packed lower 6
    static constexpr int GS = GB + 6;           //   slack cross terms (x,y,psi,slack); only for slack rows
    static constexpr int GT = GS + (CFG::ANY_SLACK_ROWS ? 4 : 0);  // barrier-modified gradient
    static constexpr int P = GT + NZ;           // Riccati cost-to-go matrix, packed lower
    static constexpr int PC = P + NPP;          // P_k c_{k-1}
    static constexpr int PV = PC + NX;          // Riccati cost-to-go vector
    static constexpr int YV = PV + NX;          // L_uu^-1 m_u
    static constexpr int LXU = YV + NU;         // L_xu = M_xu L_uu^-T, NX x NU
    static constexpr int LUU = LXU + NU * NX;   // Cholesky factor of M_uu as {1/l00, l10, 1/l11 (, l20, l21, 1/l22)}
    static constexpr int DZ = LUU + NLUU;       // QP primal iterate
    static constexpr int DPI = DZ + NZ;         // equality multipliers: QP iterate during the IPM, NLP multiplier pi_k
                                                // of the equation defining x_k between QPs (read by the linearisation)
    static constexpr int PI = DPI;
    static constexpr int DZN = DPI + NX;        // Newton target (primal); the multiplier target P dx + pv is formed on the fly
    // MIRROR work space of the CTA-parallel Jacobi: the working copy of the block lives in the T / TINV / CORR rows (dead
    // between QPs) when they hold NZ x NZ doubles, else in an extra area behind the stage image (bicycle with few rows)
    static constexpr int DZS = DZN + NZ;        // the QP iterate after the step of the current IPM iteration: written by the row phase that
                                                // takes the step, read by the accumulations, which copy it back into DZ (see rows_phase)
    static constexpr int ZS = DZS + NZ;         // one double that is 0.0 for the whole solve (set at load time, never written): the
                                                // tensor-core sweep points its structurally absent loads here instead of predicating them
    static constexpr int MIR_EXTRA = (3 * NROW >= NZ * NZ) ? 0 : NZ * NZ;
    static constexpr int MW = MIR_EXTRA ? ZS + 1 : T;
    static constexpr int RAW = ZS + 1 + MIR_EXTRA;
    static constexpr int STRIDE = RAW | 1;      // odd
    // linearisation scratch inside the work area (dead while linearising): partial Hessians.  It must not touch
    // DPI (= pi, read by the dynamics part of the NEXT stage's... of this stage's predecessor)
    static constexpr int LS_HC = WK;            // cost Hessian (packed lower NZ)   -- written by the cost part
    static constexpr int LS_GC = LS_HC + NHP;   // cost gradient NZ
    static constexpr int LS_HD = LS_GC + NZ;    // dynamics Hessian entries (CFG::NHD)
    static constexpr int LS_HX = LS_HD + CFG::NHD;  // constraint Hessian parts: 4 x 6
    static constexpr int LS_END = LS_HX + 24;
    static_assert(LS_END <= DPI, "linearisation scratch must stay below the multipliers pi (stages k < N only)");
    // per-CTA scratch.  On the device it lives inside the terminal stage's unused rows area ([LAM, WK): the terminal
    // node has no inequality rows, Hessian, dynamics or constraints); the host emulation needs one reduction slot per
    // thread and keeps it after the stages.
    static constexpr int W_RED = 0;             // reduction staging: 6 quantities x RED_SLOTS partials
    static constexpr int W_DX0 = W_RED + 6 * MPC_RED_SLOTS;  // NX
    static constexpr int W_LB = W_DX0 + NX;     // NZ lower bounds (copy of Layout, for dynamic indexing)
    static constexpr int W_UB = W_LB + NZ;      // NZ upper bounds
    static constexpr int W_STID = W_UB + NZ;    // id of the thread that runs the serial Riccati sweeps
    static constexpr int W_TMO = W_STID + 1;    // timeout flag of the current SQP iteration
    static constexpr int W_MBUF = W_TMO + 1;    // lane-cooperative factorisation: double-buffered staging tile (packed lower NZ x NZ)
    static constexpr int W_PROF = W_MBUF + 2 * NHP;  // 32 doubles (profiling builds)
#if defined(MPC_PROFILE)
    static constexpr int W_SIZE = W_PROF + 32;
#else
    static constexpr int W_SIZE = W_PROF;
#endif
#if defined(__CUDACC__)
    static constexpr int W_BASE = CFG::N * STRIDE + LAM;
    static_assert(W_SIZE <= WK - LAM, "CTA scratch must fit in the terminal stage's unused area");
    static constexpr int TOTAL = (CFG::N + 1) * STRIDE;  // doubles per CTA
#else
    static constexpr int W_BASE = (CFG::N + 1) * STRIDE;
    static constexpr int TOTAL = (CFG::N + 1) * STRIDE + W_SIZE;
#endif
};

template <class CFG>
MPC_HD bool box_active(const Layout &L, int k, int i, bool upper) {
    if (k == 0 && i >= CFG::NU) return false;  // x0 is fixed: no state box at node 0
    if (k >= CFG::N) return false;             // no bounds at the terminal node
    return upper ? (L.upper_bound[i] < 1e10) : (L.lower_bound[i] > -1e10);
}
template <class CFG>
MPC_HD bool row_has_slack(int j) {
    return CFG::kind_uses_slack(j < CFG::CNT0 ? CFG::KIND0 : (j < CFG::CNT0 + CFG::CNT1 ? CFG::KIND1 : CFG::KIND2));
}
template <class CFG>
MPC_HD bool row_active(const Layout &L, int k, int r) {
    if (r < CFG::NZ) return box_active<CFG>(L, k, r, false);
    if (r < 2 * CFG::NZ) return box_active<CFG>(L, k, r - CFG::NZ, true);
    return k >= 1 && k < CFG::N;
}
// value of inequality row r of stage k at QP point v (node layout, length NZ): r + Jr v  (>= 0 feasible)
template <class CFG>
MPC_HD double row_value(const Layout &L, const double *st, int r, const double *v) {
    typedef SL<CFG> S;
    constexpr int NZ = CFG::NZ;
    if (r < NZ) return st[S::Z + r] - L.lower_bound[r] + v[r];
    if (r < 2 * NZ) return L.upper_bound[r - NZ] - st[S::Z + r - NZ] - v[r - NZ];
    const int j = r - 2 * NZ;
    double cv = st[S::RH + j] + st[S::JR + 3 * j] * v[CFG::IX] + st[S::JR + 3 * j + 1] * v[CFG::IY] + st[S::JR + 3 * j + 2] * v[CFG::IPSI];
    if (CFG::ANY_SLACK_ROWS && row_has_slack<CFG>(j)) cv += v[CFG::HAS_SLACK ? CFG::ISLACK : 0];
    return cv;
}

// ---------------------------------------------------------------------------------------------
// CTA reductions.  kinds: 0 max (NaN propagating), 1 sum.  `red_put` is called by EVERY thread
// inside the producing phase (device: warp-shuffle pre-reduction, lane 0 stores one partial per
// warp; emulation: one partial per thread); after the phase barrier `red_get` combines the
// partials (redundantly in every thread: no second barrier).
// ---------------------------------------------------------------------------------------------
MPC_HD double red_combine(int kind, double a, double b) {
    if (kind == 1) return a + b;
    if (a != a) return a;
    return (b != b || b > a) ? b : a;
}

// kind 0 is only used on values that are >= +0 or NaN (maxima of fabs / ratio tests started at 0): for those the IEEE
// bit pattern, read as an unsigned integer, is monotone and every NaN sorts above +inf, so the NaN-propagating max is an
// integer max -- two REDUX.MAX per warp instead of five double shuffles + compare/select chains (and far less code:
// the reductions were 11 KB of the kernel's instruction footprint).
template <class CFG>
MPC_DEV void red_put(double *W, int tid, int q, int kind, double v) {
    typedef SL<CFG> S;
#if defined(__CUDA_ARCH__)
    if (kind == 0) {
        const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu, lo = (unsigned)__double2loint(v);  // (-0 -> +0)
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, (hi == mh) ? lo : 0u);
        v = __hiloint2double((int)mh, (int)ml);
    } else {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    }
    if ((tid & 31) == 0) W[S::W_RED + q * MPC_RED_SLOTS + (tid >> 5)] = v;
#else
    W[S::W_RED + q * MPC_RED_SLOTS + tid] = v;
#endif
}

template <class CFG>
MPC_DEV double red_get(const double *W, int q, int kind) {
    typedef SL<CFG> S;
#if defined(__CUDA_ARCH__)
    if (kind == 0) {
        unsigned long long m = (unsigned long long)__double_as_longlong(W[S::W_RED + q * MPC_RED_SLOTS]);
#pragma unroll
        for (int i = 1; i < MPC_RED_SLOTS; i++) {
            const unsigned long long o = (unsigned long long)__double_as_longlong(W[S::W_RED + q * MPC_RED_SLOTS + i]);
            m = o > m ? o : m;
        }
        return __longlong_as_double((long long)m);
    }
#endif
    double v = W[S::W_RED + q * MPC_RED_SLOTS];
#pragma unroll
    for (int i = 1; i < MPC_RED_SLOTS; i++) v = red_combine(kind, v, W[S::W_RED + q * MPC_RED_SLOTS + i]);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Linearisation, one (stage, part) per thread.  Parts: 0 cost, 1 dynamics, 2..5 constraint rows
// (quarters of the h vector).  Results land in the stage's BA/C/JR/RH plus partial Hessians in the
// linearisation scratch; `assemble_stage` then sums them and applies MIRROR.
// ---------------------------------------------------------------------------------------------
template <class CFG, class PA>
MPC_HD void linearize_part(const PA &pa, const Layout &L, double *st, const double *st_next, int k, int part) {
    typedef SL<CFG> S;
    constexpr int NZ = CFG::NZ, NX = CFG::NX, NH = CFG::NH;
    double z[NZ];
#pragma unroll
    for (int i = 0; i < NZ; i++) z[i] = st[S::Z + i];
#if defined(MPC_LIN_SKIP)  // development (scripts/profile_phases.py): leave parts out to time the others (results are garbage)
    if ((MPC_LIN_SKIP >> (part < 2 ? part : 2)) & 1) return;
#endif
    if (part == 0) {
        double Hp[S::NHP], g[NZ];
#pragma unroll
        for (int i = 0; i < NZ; i++) g[i] = 0.0;
#pragma unroll
        for (int i = 0; i < S::NHP; i++) Hp[i] = 0.0;
        stage_cost<CFG>(pa, L, z, L.dt, g, Hp);
#pragma unroll
        for (int i = 0; i < S::NHP; i++) st[S::LS_HC + i] = Hp[i];
#pragma unroll
        for (int i = 0; i < NZ; i++) st[S::LS_GC + i] = g[i];
    } else if (part == 1) {
        double xn[NX], pin[NX], Hd[CFG::NHD];
#pragma unroll
        for (int i = 0; i < NX; i++) pin[i] = st_next[S::PI + i];
        model_step<CFG>(L, z, xn, nullptr, pin, Hd, st + S::BB, st + S::BE);
#pragma unroll
        for (int i = 0; i < CFG::NHD; i++) st[S::LS_HD + i] = Hd[i];
#pragma unroll
        for (int i = 0; i < NX; i++) st[S::C + i] = xn[i] - st_next[S::Z + CFG::NU + i];
    } else {
        const int q = part - 2;
        double Hx[6] = {0, 0, 0, 0, 0, 0};
        if (NH > 0 && k >= 1) {
            // a quarter of the rows of h; raw h and dh/d(x,y,psi) go straight to shared memory.  Hand-written stacks: quarter q of
            // EVERY constraint block (each thread gets its share of the expensive ellipsoid / Gaussian rows); emitted stage
            // functions (codegen.py) take a contiguous row range
            constexpr bool GEN = (CFG::KIND0 == MPCGPU_CON_GENERATED);
            if (GEN) stage_constraints<CFG, true>(pa, L, z, st + S::RH, st + S::JR, st + S::LAM + 2 * NZ, Hx, (NH * q) / 4, (NH * (q + 1)) / 4);
            else stage_constraints<CFG, true>(pa, L, z, st + S::RH, st + S::JR, st + S::LAM + 2 * NZ, Hx, 0, NH, q, 4);
#pragma unroll
            for (int m = 0; m < (GEN ? 1 : 3); m++) {
                const int n = GEN ? NH : CFG::count(m), row0 = GEN ? 0 : CFG::offset(m);
                for (int j = row0 + (n * q) / 4; j < row0 + (n * (q + 1)) / 4; j++) {  // fold the bound side: row value s (h - bound) >= 0, gradient s dh
                    const double sgn = row_sign_of<CFG>(j);
                    st[S::RH + j] = sgn * (st[S::RH + j] - row_bound_of<CFG>(j)) + L.qp_relax;
                    st[S::JR + 3 * j] *= sgn;
                    st[S::JR + 3 * j + 1] *= sgn;
                    st[S::JR + 3 * j + 2] *= sgn;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) st[S::LS_HX + 6 * q + i] = Hx[i];
    }
}

template <class CFG>
MPC_HD void assemble_stage(const Layout &L, double *st) {
    typedef SL<CFG> S;
    constexpr int NZ = CFG::NZ;
    double Hp[S::NHP];
#pragma unroll
    for (int i = 0; i < S::NHP; i++) Hp[i] = st[S::LS_HC + i];
    // dynamics: pi-weighted second derivatives over the variables Phi is nonlinear in, order as written by model_step
#pragma unroll
    for (int e = 0; e < CFG::NHD; e++) {
        int hi, hj;
        dyn_hess_pair<CFG>(e, hi, hj);
        Hp[trisym(hi, hj)] += st[S::LS_HD + e];
    }
    {
        const int idx[3] = {CFG::IX, CFG::IY, CFG::IPSI};
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b <= a; b++)
                Hp[trisym(idx[a], idx[b])] += st[S::LS_HX + tri(a, b)] + st[S::LS_HX + 6 + tri(a, b)] + st[S::LS_HX + 12 + tri(a, b)] + st[S::LS_HX + 18 + tri(a, b)];
    }
    double g[NZ];
#pragma unroll
    for (int i = 0; i < NZ; i++) g[i] = st[S::LS_GC + i];
#pragma unroll
    for (int i = 0; i < S::NHP; i++) st[S::H + i] = Hp[i];
#pragma unroll
    for (int i = 0; i < NZ; i++) st[S::GR + i] = g[i];
    {   // block structure flag for mirror_stage (DZN is dead between QPs and outside the linearisation scratch)
        const int ia[4] = {CFG::IA, CFG::IW, CFG::IPSI, CFG::IV};
        const int ib[4] = {CFG::IX, CFG::IY, CFG::IS, CFG::HAS_SLACK ? CFG::ISLACK : 0};
        bool decoupled = (CFG::OBJ == MPCGPU_OBJ_CONTOURING) && !CFG::BICYCLE;
        if constexpr (!CFG::BICYCLE) {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < NZ - 4; b++) decoupled = decoupled && (Hp[trisym(ia[a], ib[b])] == 0.0);
        }
        static_assert(S::DZN >= S::LS_END, "flag must not overlap the linearisation scratch");
        st[S::DZN] = decoupled ? 1.0 : 0.0;
    }
}

// MIRROR of the assembled stage Hessian st[H] (in place), thread = (stage, part).  For the MPCC objective with a centred
// disc the block is exactly block diagonal over {a, w, psi, v} and {x, y, spline (, slack)} (checked at run time): the two
// blocks are then regularised by two different threads (part 0 / part 1) -- the eigen-decomposition of a block-diagonal
// matrix is the union of its blocks' -- else part 0 runs the general path on the whole block.
template <class CFG>
MPC_HD void mirror_stage(const Layout &L, double *st, int part) {
    typedef SL<CFG> S;
    constexpr int NZ = CFG::NZ;
    // Jacobi work space of the general path: M in the (dead between QPs) T/TINV/CORR rows area, V in the IPM work area
    static_assert(S::DZN - S::WK >= NZ * NZ, "MIRROR work space");
    constexpr bool BLOCKS = (CFG::OBJ == MPCGPU_OBJ_CONTOURING) && !CFG::BICYCLE;
    constexpr int NA = 4, NB = NZ - NA;
    const int ia[4] = {CFG::IA, CFG::IW, CFG::IPSI, CFG::IV};
    const int ib[4] = {CFG::IX, CFG::IY, CFG::IS, CFG::HAS_SLACK ? CFG::ISLACK : 0};
    double *Hp = st + S::H;
    if constexpr (BLOCKS) {
        if (st[S::DZN] == 1.0) {  // written by assemble_stage: exactly block diagonal (part 0 may set 2 concurrently)
            if (part == 0) mirror_subblock<NA>(Hp, L.reg_epsilon, ia);
            else mirror_subblock<NB>(Hp, L.reg_epsilon, ib);
            return;
        }
    }
    // general path: PD test here (MIRROR leaves a block with all eigenvalues above eps unchanged); an indefinite block is
    // flagged (DZN = 2) for the CTA-parallel Jacobi of mirror_general_cta
    if (part == 0 && !packed_block_is_pd<NZ>(Hp, L.reg_epsilon)) st[S::DZN] = 2.0;
}

// ---------------------------------------------------------------------------------------------
// MIRROR of the stage blocks that are neither block diagonal nor positive definite (CA objective, offset discs):
// cyclic Jacobi parallelised the classical way -- a round-robin tournament orders the n(n-1)/2 index pairs into m - 1
// rounds (m = n rounded up to even) of m/2 DISJOINT pairs; thread (stage, j) applies pair j of the round.  Disjoint pairs
// commute: all column updates (A <- A J, V <- V J) of a round, barrier, all row updates (A <- J' A), barrier.  The three
// entries a rotation is computed from are only ever changed by that pair itself, so no barrier is needed before them.
// Serial (one thread per stage, 28 rotations per sweep of an 8 x 8 block) this was 32 % of a cfg 3 solve.
//   work space per stage: Mw = st[MW..] (NZ x NZ: the T / TINV / CORR rows, dead between QPs, or an extra area), Vw = st[WK..]
//   (NZ x NZ), (c, s) of the round behind Vw.
// ---------------------------------------------------------------------------------------------
MPC_HD void tournament_pair(int m, int r, int j, int &p, int &q) {
    if (j == 0) {
        p = r;
        q = m - 1;
    } else {
        const int a = (r + j) % (m - 1), b = (r - j + (m - 1)) % (m - 1);
        p = a < b ? a : b;
        q = a < b ? b : a;
    }
}

template <class CFG>
MPC_DEV void mirror_general_cta(const Layout &L, double *sm) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NT = MPC_NT;
    constexpr int M = (NZ + 1) & ~1, G = M / 2;  // rounds use M indices (index NZ is a dummy when NZ is odd); G pairs per round
    constexpr int CS = S::WK + NZ * NZ;          // (c, s) of the current round: 2 G doubles
    static_assert(S::DZN - S::WK >= NZ * NZ + 2 * G, "MIRROR work space");
    const double eps = L.reg_epsilon;
    // copy the packed block into the work space, V = I (rows split over the G threads of a stage)
    MPC_PHASE
    for (int item = tid; item < N * G; item += NT) {  // item = (stage, pair of the round)
        double *st = sm + (item / G) * S::STRIDE;
#if !defined(__CUDACC__) && defined(MPC_EMU_TRACE)
        if (st[S::DZN] == 2.0 && item % G == 0) g_mirror_general++;
#endif
        if (st[S::DZN] == 2.0)
            for (int i = item % G; i < NZ; i += G)
                for (int j = 0; j < NZ; j++) {
                    st[S::MW + i * NZ + j] = st[S::H + trisym(i, j)];
                    st[S::WK + i * NZ + j] = (i == j) ? 1.0 : 0.0;
                }
    }
    MPC_PHASE_END
    for (int sweep = 0; sweep < 30; sweep++) {
        int active = 0;
        MPC_PHASE
        for (int item = tid; item < N * G; item += NT) {  // convergence of each flagged stage (same rule as the register Jacobi)
            if (item % G != 0) continue;
            double *st = sm + (item / G) * S::STRIDE;
            if (st[S::DZN] == 2.0) {
                const double *Mw = st + S::MW;
                double offd = 0.0, diag = 0.0;
                for (int i = 0; i < NZ; i++) {
                    diag += Mw[i * NZ + i] * Mw[i * NZ + i];
                    for (int j = i + 1; j < NZ; j++) offd += Mw[i * NZ + j] * Mw[i * NZ + j];
                }
                if (offd <= 1e-30 * (diag + offd) || offd == 0.0) st[S::DZN] = 3.0;  // converged: rebuild below
                else active |= 1;
            }
        }
        MPC_PHASE_END_OR(active)
        if (!active) break;
        for (int r = 0; r < M - 1; r++) {
            MPC_PHASE
            for (int item = tid; item < N * G; item += NT) {  // rotation of the own pair; column updates of A and V
                double *st = sm + (item / G) * S::STRIDE;
                const int j = item % G;
                int p, q;
                tournament_pair(M, r, j, p, q);
                double c = 1.0, sn = 0.0;
                if (st[S::DZN] == 2.0 && q < NZ) {
                    double *Mw = st + S::MW, *Vw = st + S::WK;
                    const double apq = Mw[p * NZ + q], app = Mw[p * NZ + p], aqq = Mw[q * NZ + q];
                    if (!(fabs(apq) <= 1e-19 * (fabs(app) + fabs(aqq)))) {  // rotations that cannot change the diagonal are skipped
                        jacobi_rotation(app, aqq, apq, c, sn);
                        for (int k = 0; k < NZ; k++) {
                            const double mkp = Mw[k * NZ + p], mkq = Mw[k * NZ + q];
                            Mw[k * NZ + p] = c * mkp - sn * mkq;
                            Mw[k * NZ + q] = sn * mkp + c * mkq;
                            const double vkp = Vw[k * NZ + p], vkq = Vw[k * NZ + q];
                            Vw[k * NZ + p] = c * vkp - sn * vkq;
                            Vw[k * NZ + q] = sn * vkp + c * vkq;
                        }
                    }
                }
                st[CS + 2 * j] = c;
                st[CS + 2 * j + 1] = sn;
            }
            MPC_PHASE_END
            MPC_PHASE
            for (int item = tid; item < N * G; item += NT) {  // row updates
                double *st = sm + (item / G) * S::STRIDE;
                const int j = item % G;
                int p, q;
                tournament_pair(M, r, j, p, q);
                const double c = st[CS + 2 * j], sn = st[CS + 2 * j + 1];
                if (st[S::DZN] == 2.0 && q < NZ && sn != 0.0) {
                    double *Mw = st + S::MW;
                    for (int k = 0; k < NZ; k++) {
                        const double mpk = Mw[p * NZ + k], mqk = Mw[q * NZ + k];
                        Mw[p * NZ + k] = c * mpk - sn * mqk;
                        Mw[q * NZ + k] = sn * mpk + c * mqk;
                    }
                }
            }
            MPC_PHASE_END
        }
    }
    // rebuild V f(D) V' into the packed block (rows split over the G threads): eigenvalues e -> eps if |e| <= eps, else |e|
    MPC_PHASE
    for (int item = tid; item < N * G; item += NT) {
        double *st = sm + (item / G) * S::STRIDE;
        if (st[S::DZN] >= 2.0) {
            const double *Mw = st + S::MW, *Vw = st + S::WK;
            double e[NZ];
            for (int k = 0; k < NZ; k++) {
                double l = Mw[k * NZ + k];
                if (l >= -eps && l <= eps) l = eps;
                else if (l < 0.0) l = -l;
                e[k] = l;
            }
            for (int i = item % G; i < NZ; i += G)
                for (int j = 0; j <= i; j++) {
                    double acc = 0.0;
                    for (int k = 0; k < NZ; k++) acc += Vw[i * NZ + k] * e[k] * Vw[j * NZ + k];
                    st[S::H + tri(i, j)] = acc;
                }
        }
    }
    MPC_PHASE_END
}


// ---------------------------------------------------------------------------------------------
// Riccati sweeps.  The recursion is a serial chain over the horizon whose per-stage blocks are
// tiny (nx = 5..6, nu = 2); measured on B200, a cooperative version (threads = matrix entries,
// __syncthreads between the steps of a stage) spends most of its time in barriers and index
// arithmetic.  Here ONE thread runs the whole sweep out of registers with fully unrolled,
// constant-offset shared-memory accesses, exploiting the structure of the unicycle family:
// A = I + E with E non-zero only in the (psi, v) columns.
// ---------------------------------------------------------------------------------------------
template <class CFG>
struct RicIdx {
    static constexpr int NE = CFG::NE;  // columns of A that differ from the identity: psi, v (, delta)
    MPC_HD static constexpr int ecol(int c) { return 2 + c; }  // x-index of psi, v, delta
    MPC_HD static constexpr int eidx(int xi) { return (xi >= 2 && xi < 2 + NE) ? xi - 2 : -1; }
    // structural non-zeros of the discrete dynamics of the unicycle family (stage_functions.cuh,
    // unicycle_step): rows x=0, y=1, psi=2, v=3, spline=4 (, slack=5: identity row)
    //   E = A - I :  column psi: rows x, y ;  column v: rows x, y, spline
    //   B         :  column a: rows x, y, v, spline ;  column w: rows x, y, psi
    // (the bicycle takes the dense sweeps below: every entry counts)
    MPC_HD static constexpr bool enz(int l, int c) { return CFG::BICYCLE ? true : (c == 0 ? (l == 0 || l == 1) : (l == 0 || l == 1 || l == 4)); }
    MPC_HD static constexpr bool bnz(int l, int u) { return CFG::BICYCLE ? true : (u == 0 ? (l == 0 || l == 1 || l == 3 || l == 4) : (l == 0 || l == 1 || l == 2)); }
};

// Backward factorisation: for k = N-1..0   M = H + barrier terms + [B A]' P+ [B A];
//   Luu = chol(Muu);  Lxu = Mxu Luu^-T;  P = Mxx - Lxu Lxu';  PC(k+1) = P_{k+1} c_k.
// The backward VECTOR sweep for the gradient currently in GT (the predictor right-hand side) rides along: its ~45
// FP64 operations per stage are independent of the matrix recursion and fill that chain's latency bubbles, which is
// cheaper than a separate pass over the horizon (riccati_solve<CFG, false> then only runs the forward sweep).
template <class CFG>
MPC_DEV void riccati_factor(double *sm) {
    typedef SL<CFG> S;
    typedef RicIdx<CFG> R;
    constexpr int N = CFG::N, NX = CFG::NX, NU = CFG::NU, NE = R::NE;
    static_assert(NU == 2, "input block is factorised in closed form for nu = 2");
    const int serial_tid = (int)sm[S::W_BASE + S::W_STID];
    MPC_PHASE
    if (tid == serial_tid) {
        double Pn[S::NPP], pv[NX];
#pragma unroll
        for (int e = 0; e < S::NPP; e++) Pn[e] = sm[N * S::STRIDE + S::P + e];
#pragma unroll
        for (int i = 0; i < NX; i++) {
            pv[i] = sm[N * S::STRIDE + S::GT + NU + i];  // terminal: pv_N = gt_N
            sm[N * S::STRIDE + S::PV + i] = pv[i];
        }
        for (int k = N - 1; k >= 0; k--) {
            double *st = sm + k * S::STRIDE;
            double *stn = st + S::STRIDE;
            double B[NX][NU], E[NX][NE], PB[NX][NU], PE[NX][NE], qv[NX];
#pragma unroll
            for (int l = 0; l < NX; l++) {
#pragma unroll
                for (int u = 0; u < NU; u++) B[l][u] = st[S::BB + l * NU + u];
#pragma unroll
                for (int c = 0; c < NE; c++) E[l][c] = st[S::BE + l * NE + c];
            }
#pragma unroll
            for (int l = 0; l < NX; l++) {
                double pc = 0.0;
#pragma unroll
                for (int u = 0; u < NU; u++) PB[l][u] = 0.0;
#pragma unroll
                for (int c = 0; c < NE; c++) PE[l][c] = 0.0;
#pragma unroll
                for (int m = 0; m < NX; m++) {
                    const double plm = Pn[trisym(l, m)];
#pragma unroll
                    for (int u = 0; u < NU; u++)
                        if (R::bnz(m, u)) PB[l][u] += plm * B[m][u];
#pragma unroll
                    for (int c = 0; c < NE; c++)
                        if (R::enz(m, c)) PE[l][c] += plm * E[m][c];
                    pc += plm * st[S::C + m];
                }
                stn[S::PC + l] = pc;
                qv[l] = pv[l] + pc;  // vector sweep: q = pv+ + P+ c
            }
            // vector sweep: m = gt + [B A]' q
            double mv0 = st[S::GT + 0], mv1 = st[S::GT + 1];
#pragma unroll
            for (int l = 0; l < NX; l++) {
                if (R::bnz(l, 0)) mv0 += B[l][0] * qv[l];
                if (R::bnz(l, 1)) mv1 += B[l][1] * qv[l];
            }
            // M_uu
            double m00 = st[S::H + tri(0, 0)] + st[S::GD + 0], m10 = st[S::H + tri(1, 0)], m11 = st[S::H + tri(1, 1)] + st[S::GD + 1];
#pragma unroll
            for (int l = 0; l < NX; l++) {
                if (R::bnz(l, 0)) m00 += B[l][0] * PB[l][0];
                if (R::bnz(l, 1)) m10 += B[l][1] * PB[l][0];
                if (R::bnz(l, 1)) m11 += B[l][1] * PB[l][1];
            }
            // the two reciprocal square roots are independent: 1/l11 = rsqrt(m11 - m10^2/m00) = sqrt(m00) rsqrt(m00 m11 - m10^2)
            // (one rsqrt latency, ~75 cycles, off the per-stage critical path)
            const double il00 = mpc_rsqrt(m00);
            const double rdet = mpc_rsqrt(m11 * m00 - m10 * m10);
            const double l10 = m10 * il00;
            const double il11 = rdet * (m00 * il00);
            st[S::LUU] = il00;
            st[S::LUU + 1] = l10;
            st[S::LUU + 2] = il11;
            const double yv0 = mv0 * il00;  // vector sweep: y = L_uu^-1 m_u
            const double yv1 = (mv1 - l10 * yv0) * il11;
            st[S::YV] = yv0;
            st[S::YV + 1] = yv1;
            // M_xu = H_xu + A' P B = H_xu + PB + E' PB (rows psi, v only);  L_xu = M_xu L_uu^-T
            double Lx[NX][NU];
#pragma unroll
            for (int i = 0; i < NX; i++) {
                double mx0 = st[S::H + tri(NU + i, 0)] + PB[i][0], mx1 = st[S::H + tri(NU + i, 1)] + PB[i][1];
                if (R::eidx(i) >= 0) {
#pragma unroll
                    for (int l = 0; l < NX; l++) {
                        if (!R::enz(l, R::eidx(i) >= 0 ? R::eidx(i) : 0)) continue;
                        mx0 += E[l][R::eidx(i) >= 0 ? R::eidx(i) : 0] * PB[l][0];
                        mx1 += E[l][R::eidx(i) >= 0 ? R::eidx(i) : 0] * PB[l][1];
                    }
                }
                Lx[i][0] = mx0 * il00;
                Lx[i][1] = (mx1 - Lx[i][0] * l10) * il11;
                {   // stored for the vector sweeps: the feedback gain, transposed, Kt = L_xu L_uu^-1 (riccati_solve)
                    const double kt1 = Lx[i][1] * il11;
                    st[S::LXU + i * NU] = (Lx[i][0] - kt1 * l10) * il00;
                    st[S::LXU + i * NU + 1] = kt1;
                }
                // vector sweep: pv_i = m_x,i - L_xu,i y,  m_x = gt_x + A' q,  A = I + E
                double mx = st[S::GT + NU + i] + qv[i];
                if (R::eidx(i) >= 0) {
#pragma unroll
                    for (int l = 0; l < NX; l++)
                        if (R::enz(l, R::eidx(i) >= 0 ? R::eidx(i) : 0)) mx += E[l][R::eidx(i) >= 0 ? R::eidx(i) : 0] * qv[l];
                }
                pv[i] = mx - Lx[i][0] * yv0 - Lx[i][1] * yv1;
                st[S::PV + i] = pv[i];
            }
            // M_xx = H_xx + barrier + A' P A,  A' P A = P + PE Sel + Sel' PE' + Sel' (E' PE) Sel;  P = M_xx - L_xu L_xu'
            double Pk[S::NPP];
#pragma unroll
            for (int i = 0; i < NX; i++) {
#pragma unroll
                for (int j = 0; j <= i; j++) {
                    double acc = st[S::H + tri(NU + i, NU + j)] + Pn[tri(i, j)];
                    if (i == j) acc += st[S::GD + NU + i];
                    {
                        const int ai = (NU + i == CFG::IX) ? 0 : (NU + i == CFG::IY) ? 1 : (NU + i == CFG::IPSI) ? 2 : -1;
                        const int aj = (NU + j == CFG::IX) ? 0 : (NU + j == CFG::IY) ? 1 : (NU + j == CFG::IPSI) ? 2 : -1;
                        if (ai >= 0 && aj >= 0) acc += st[S::GB + trisym(ai >= 0 ? ai : 0, aj >= 0 ? aj : 0)];
                        if (CFG::ANY_SLACK_ROWS) {
                            if (NU + i == CFG::ISLACK && aj >= 0) acc += st[S::GS + (aj >= 0 ? aj : 0)];
                            if (NU + i == CFG::ISLACK && NU + j == CFG::ISLACK) acc += st[S::GS + 3];
                        }
                    }
                    if (R::eidx(j) >= 0) acc += PE[i][R::eidx(j) >= 0 ? R::eidx(j) : 0];
                    if (R::eidx(i) >= 0) acc += PE[j][R::eidx(i) >= 0 ? R::eidx(i) : 0];
                    if (R::eidx(i) >= 0 && R::eidx(j) >= 0) {
#pragma unroll
                        for (int l = 0; l < NX; l++)
                            if (R::enz(l, R::eidx(i) >= 0 ? R::eidx(i) : 0)) acc += E[l][R::eidx(i) >= 0 ? R::eidx(i) : 0] * PE[l][R::eidx(j) >= 0 ? R::eidx(j) : 0];
                    }
                    acc -= Lx[i][0] * Lx[j][0] + Lx[i][1] * Lx[j][1];
                    Pk[tri(i, j)] = acc;
                    st[S::P + tri(i, j)] = acc;
                }
            }
#pragma unroll
            for (int e = 0; e < S::NPP; e++) Pn[e] = Pk[e];
        }
    }
    MPC_PHASE_END
}

// ---------------------------------------------------------------------------------------------
// Lane-cooperative backward factorisation (one warp, lane = column of the stage KKT block).
//
// A lone thread issues one FP64 instruction per two cycles with 1 of 32 lanes useful: the serial
// factorisation above is bound by the FP64 pipe's ISSUE rate, not by its dependency chain.  Here lane j < NZ owns
// column j of  M = H + G + [B A]' P+ [B A]  and lane NZ carries the vector recursion of the affine right-hand side
// (m = gt + [B A]'(pv+ + P+ c)), which has exactly the algebra of one more column -- so the backward
// sweep of the predictor solve rides along for free.  Per stage:
//   step A (no communication): column of W = P+ [B A | c] (P+ replicated in every lane's registers), column of M;
//          the columns go to a double-buffered staging tile (packed lower NZ x NZ) in shared memory;     __syncwarp
//   step B (replicated): every lane reads the lower triangle, forms L_uu, L_xu and the Schur complement P
//          (so P is again replicated: ONE exchange per stage), plus the Schur update of its own column
//          (P column -> shared memory; for the vector lane: y and pv).
// The host emulation runs the lanes of each step one after the other.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
#define MPC_LANES_BEGIN {
#define MPC_LANES_END \
    }                 \
    __syncwarp();
#define MPC_LANES_END_NOSYNC }
#define MPC_LV(x) x
#define MPC_LANE_DECL(T, name, n) T name[n]
#else
#define MPC_LANES_BEGIN for (int lane = 0; lane < NL; ++lane) {
#define MPC_LANES_END }
#define MPC_LANES_END_NOSYNC }
#define MPC_LV(x) x[lane]
#define MPC_LANE_DECL(T, name, n) T name[32][n]
#endif

// Hides how a per-lane table entry was derived from the lane id: the compiler then keeps the entry in a register and
// predicates on it instead of re-deriving it through (divergent) branches in the stage loop.
MPC_DEV void mpc_opaque(int &v) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(v));
#else
    (void)v;
#endif
}

MPC_DEV void mpc_opaque(double *&v) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(v));
#else
    (void)v;
#endif
}

template <class CFG>
MPC_DEV void riccati_factor_lanes(double *sm) {
    typedef SL<CFG> S;
    typedef RicIdx<CFG> R;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NE = R::NE;
    constexpr int NL = NZ + 1;  // active lanes: NZ matrix columns + the vector lane
    static_assert(NU == 2 && NL <= 32, "input block is factorised in closed form for nu = 2");
    double *W = sm + S::W_BASE;
    const int serial_tid = (int)W[S::W_STID];
    MPC_PHASE
#if defined(__CUDACC__)
    if ((tid & ~31) == serial_tid) {
        const int lane_raw = tid & 31;
        const int lane = lane_raw < NL ? lane_raw : NL - 1;  // surplus lanes shadow the vector lane (no stores)
        const bool live = lane_raw < NL;
#else
    if (tid == serial_tid) {
        const bool live = true;
#endif
        MPC_LANE_DECL(double, Pn, S::NPP);   // P+ (replicated)
        MPC_LANE_DECL(double, add, NX);      // vector lane: pv+; matrix lanes: 0
        MPC_LANE_DECL(double, Mc, NZ);       // own column of M (vector lane: m)
        MPC_LANE_DECL(double, Wc, NX);       // own column of W
        MPC_LANE_DECL(int, hoff, NZ);        // offsets of the own column of H (vector lane: GT)
        MPC_LANE_DECL(int, cinfo, 4);        // column of [B A | c]: offset, stride, present, identity row
        MPC_LANE_DECL(int, goff, 4);         // barrier block entries of the own column: rows x, y, psi; slack row
        MPC_LANE_DECL(int, gflag, 2);        // own column in the (x, y, psi) block; slack row entry present
        MPC_LANE_DECL(int, sto, 4);          // where the own (o0, o1) pair and the own Schur column go; row mask of the latter; flags
        MPC_LANES_BEGIN
        const bool vec = (lane == NZ);
#pragma unroll
        for (int e = 0; e < S::NPP; e++) MPC_LV(Pn)[e] = sm[N * S::STRIDE + S::P + e];
#pragma unroll
        for (int i = 0; i < NX; i++) {
            const double g = sm[N * S::STRIDE + S::GT + NU + i];  // terminal: pv_N = gt_N
            MPC_LV(add)[i] = vec ? g : 0.0;
            if (vec && live) sm[N * S::STRIDE + S::PV + i] = g;
        }
#pragma unroll
        for (int i = 0; i < NZ; i++) MPC_LV(hoff)[i] = vec ? (S::GT + i) : (S::H + trisym(i, lane < NZ ? lane : 0));
        {
            int ofs = 0, str = 0, has = 0, idr = -1;
            if (vec) {
                ofs = S::C; str = 1; has = 1;
            } else if (lane < NU) {
                ofs = S::BB + lane; str = NU; has = 1;
            } else {
                idr = lane - NU;
                const int e = R::eidx(idr);
                if (e >= 0) { ofs = S::BE + e; str = NE; has = 1; }
            }
            MPC_LV(cinfo)[0] = ofs; MPC_LV(cinfo)[1] = str; MPC_LV(cinfo)[2] = has; MPC_LV(cinfo)[3] = idr;
            const int aj = (lane == CFG::IX) ? 0 : (lane == CFG::IY) ? 1 : (lane == CFG::IPSI) ? 2 : -1;
            const int ajc = aj >= 0 ? aj : 0;
#pragma unroll
            for (int a = 0; a < 3; a++) MPC_LV(goff)[a] = S::GB + trisym(a, ajc);
            MPC_LV(gflag)[0] = (aj >= 0);
            // slack row of the own column: GS[aj] under an (x, y, psi) column, GS[3] under the slack column
            MPC_LV(goff)[3] = S::GS + ((CFG::ANY_SLACK_ROWS && lane == CFG::ISLACK) ? 3 : ajc);
            MPC_LV(gflag)[1] = CFG::ANY_SLACK_ROWS && (aj >= 0 || lane == CFG::ISLACK);
            MPC_LV(sto)[0] = vec ? S::YV : S::LXU + (lane >= NU ? lane - NU : 0) * NU;
            MPC_LV(sto)[1] = vec ? S::PV : S::P + (lane >= NU ? lane - NU : 0);
            {   // rows of the own Schur column that are stored: vector lane: all of pv; x lane: the lower triangle of P
                int mask = 0;
#pragma unroll
                for (int i = 0; i < NX; i++) mask |= ((int)live & ((int)vec | ((int)(lane >= NU) & (int)(i >= lane - NU)))) << i;
                MPC_LV(sto)[2] = mask;
                // flags: 1 vector lane, 2 store (o0, o1), 4 store L_uu, 8 live vector lane, 16.. rows of the staging tile
                int fl = (int)vec | (((int)live & (int)(lane >= NU)) << 1) | (((int)live & (int)(lane == 0)) << 2) | (((int)live & (int)vec) << 3);
#pragma unroll
                for (int i = 0; i < NZ; i++) fl |= ((int)live & (int)(lane < NZ) & (int)(i >= lane)) << (4 + i);
                MPC_LV(sto)[3] = fl;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                mpc_opaque(MPC_LV(sto)[i]);
                mpc_opaque(MPC_LV(cinfo)[i]);
                mpc_opaque(MPC_LV(goff)[i]);
            }
            mpc_opaque(MPC_LV(gflag)[0]);
            mpc_opaque(MPC_LV(gflag)[1]);
#pragma unroll
            for (int i = 0; i < NZ; i++) mpc_opaque(MPC_LV(hoff)[i]);
        }
        MPC_LANES_END_NOSYNC
        for (int k = N - 1; k >= 0; k--) {
            double *st = sm + k * S::STRIDE;
            double *stn = st + S::STRIDE;
            double *buf = W + S::W_MBUF + (k & 1) * S::NHP;
            // ---- step A: own column of W and of M ----
            MPC_LANES_BEGIN
            const int fl = MPC_LV(sto)[3];
            double col[NX];
#pragma unroll
            for (int l = 0; l < NX; l++) {
                double v = MPC_LV(cinfo)[2] ? st[MPC_LV(cinfo)[0] + MPC_LV(cinfo)[1] * l] : 0.0;
                if (l == MPC_LV(cinfo)[3]) v += 1.0;
                col[l] = v;
            }
            // short dependency chains throughout: the FP64 pipe's result latency, not its issue rate, bounds this loop
#pragma unroll
            for (int l = 0; l < NX; l++) {
                double t0 = MPC_LV(Pn)[trisym(l, 0)] * col[0];
                double t1 = MPC_LV(Pn)[trisym(l, NX - 1)] * col[NX - 1];
#pragma unroll
                for (int m = 1; m < NX - 1; m++) {
                    if (m & 1) t0 += MPC_LV(Pn)[trisym(l, m)] * col[m];
                    else t1 += MPC_LV(Pn)[trisym(l, m)] * col[m];
                }
                const double acc = t0 + t1;
                if (fl & 8) stn[S::PC + l] = acc;  // P+ c (vector lane)
                MPC_LV(Wc)[l] = acc + MPC_LV(add)[l];
            }
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                double acc = st[MPC_LV(hoff)[i]];  // H / gradient part (+ barrier terms below)
                double dot = 0.0;                  // [B A]' W part
                if (i < NU) {
                    bool firstt = true;
#pragma unroll
                    for (int l = 0; l < NX; l++)
                        if (R::bnz(l, i < NU ? i : 0)) {
                            const double b = st[S::BB + l * NU + (i < NU ? i : 0)];
                            dot = firstt ? b * MPC_LV(Wc)[l] : dot + b * MPC_LV(Wc)[l];
                            firstt = false;
                        }
                } else {
                    dot = MPC_LV(Wc)[i >= NU ? i - NU : 0];
                    const int e = R::eidx(i >= NU ? i - NU : 0);
                    if (e >= 0) {
#pragma unroll
                        for (int l = 0; l < NX; l++)
                            if (R::enz(l, e >= 0 ? e : 0)) dot += st[S::BE + l * NE + (e >= 0 ? e : 0)] * MPC_LV(Wc)[l];
                    }
                }
                {   // barrier Hessian terms of entry (i, lane): unconditional loads + selects (a lane-dependent branch
                    // costs a divergence/reconvergence round trip per use)
                    const double gd = st[S::GD + i];
                    acc += (i == lane) ? gd : 0.0;
                    const int ai = (i == CFG::IX) ? 0 : (i == CFG::IY) ? 1 : (i == CFG::IPSI) ? 2 : -1;
                    if (ai >= 0) {
                        const double gb = st[MPC_LV(goff)[ai >= 0 ? ai : 0]];
                        acc += MPC_LV(gflag)[0] ? gb : 0.0;
                    }
                    if (CFG::ANY_SLACK_ROWS) {
                        if (ai >= 0) {  // (xyz row, slack column)
                            const double gs = st[S::GS + (ai >= 0 ? ai : 0)];
                            acc += (lane == CFG::ISLACK) ? gs : 0.0;
                        }
                        if (i == CFG::ISLACK) {  // (slack row, xyz or slack column)
                            const double gs = st[MPC_LV(goff)[3]];
                            acc += MPC_LV(gflag)[1] ? gs : 0.0;
                        }
                    }
                    acc += dot;
                    if ((fl >> (4 + i)) & 1) buf[tri(i, 0) + lane] = acc;  // lower triangle, packed
                }
                MPC_LV(Mc)[i] = acc;
            }
            MPC_LANES_END
            // ---- step B: replicated Schur complement + own column ----
            MPC_LANES_BEGIN
            const int fl = MPC_LV(sto)[3];
            const bool vec = fl & 1;
            // L_uu = chol(M_uu) with the two reciprocal square roots independent of each other:
            //   1/l11 = rsqrt(m11 - m10^2/m00) = sqrt(m00) rsqrt(m00 m11 - m10^2)
            const double m00 = buf[tri(0, 0)], m10 = buf[tri(1, 0)], m11 = buf[tri(1, 1)];
            const double il00 = mpc_rsqrt(m00);
            const double rdet = mpc_rsqrt(m11 * m00 - m10 * m10);
            const double l10 = m10 * il00;
            const double il11 = rdet * (m00 * il00);
            const double cx = l10 * il00 * il11;
            double Lx[NX][NU];
#pragma unroll
            for (int i = 0; i < NX; i++) {  // L_xu = M_xu L_uu^-T
                const double b0 = buf[tri(NU + i, 0)], b1 = buf[tri(NU + i, 1)];
                Lx[i][0] = b0 * il00;
                Lx[i][1] = b1 * il11 - b0 * cx;
            }
#pragma unroll
            for (int i = 0; i < NX; i++)
#pragma unroll
                for (int j = 0; j <= i; j++)
                    MPC_LV(Pn)[tri(i, j)] = (buf[tri(NU + i, NU + j)] - Lx[i][1] * Lx[j][1]) - Lx[i][0] * Lx[j][0];
            // own column: (o0, o1) = L_uu^-1 [M_0j; M_1j];  x rows minus L_xu (o0, o1)'
            const double o0 = MPC_LV(Mc)[0] * il00;
            const double o1 = MPC_LV(Mc)[1] * il11 - MPC_LV(Mc)[0] * cx;
            {   // stores under simple predicates (no lane-dependent branches)
                if (fl & 2) {  // vector lane: y;  x lanes: row of Kt = L_xu L_uu^-1 (what riccati_solve reads)
                    const double s1 = vec ? o1 : o1 * il11;
                    st[MPC_LV(sto)[0]] = vec ? o0 : (o0 - s1 * l10) * il00;
                    st[MPC_LV(sto)[0] + 1] = s1;
                }
                if (fl & 4) {
                    st[S::LUU] = il00;
                    st[S::LUU + 1] = l10;
                    st[S::LUU + 2] = il11;
                }
            }
#pragma unroll
            for (int i = 0; i < NX; i++) {
                const double v = (MPC_LV(Mc)[NU + i] - Lx[i][1] * o1) - Lx[i][0] * o0;
                MPC_LV(add)[i] = vec ? v : 0.0;
                // vector lane: pv_i;  x lane: P(i, lane - NU) of the lower triangle
                double *dst = st + MPC_LV(sto)[1] + (vec ? i : tri(i, 0));
                mpc_opaque(dst);  // address formed unconditionally: the store below is one predicated instruction
                if ((MPC_LV(sto)[2] >> i) & 1) *dst = v;
            }
            MPC_LANES_END_NOSYNC
        }
    }
    MPC_PHASE_END
}

// ---------------------------------------------------------------------------------------------
// Backward factorisation on the FP64 tensor-core path (DMMA.8x8x4): one warp, all 32 lanes.
//
// The single-thread sweep above issues ~300 FP64 + ~210 other instructions per stage with 1 of 32 lanes useful and
// needs ~1150 cycles per stage (1750 with four solves per SM).  Here the three matrix products of a stage are five
// `mma.sync.m8n8k4.f64` (measured on B200: 26 cycles dependent, 18 cycles issue-to-issue; scripts/micro/dmma_micro.cu):
//   index n in [0, NZ]: row of the augmented block (n < NZ: variable of z = [u; x]; n = NZ: the "vector row" that carries
//   the affine recursion, which has the algebra of one more row);  a, r, k in [0, NZ): variables of z.
//     Pt[n][a]   cost-to-go of the next stage:  x rows / columns = P+,  vector row = pv+;   u rows / columns: unused
//     GEMM 1     Wt[n][r] = sum_k At[n][k] Pt[r][k]      At[n][k] = [B A]_{k,n} (x rows k),  At[NZ][k] = c_k
//                         = (P+ [B A | c])'              the vector row then receives + pv+  (q = pv+ + P+ c)
//     GEMM 2     Mt[n][a] = sum_r Wt[n][r] [B A]_{r,a} + (H + G | gt)[n][a]         = M (symmetric) | m
//     Cholesky   of M_uu in closed form (as in the serial sweep), L[n][0..1] = rows of L_xu (x rows) | y (vector row)
//     GEMM 3     Pt[n][a] = Mt[n][a] - sum_{k<2} L[n][k] L[a][k]                    = P | pv
// Fragment layouts of m8n8k4 (g = lane / 4, q = lane % 4):  A: (row g, k q);  B: (k q, column g);  C/D: (row g,
// columns 2q, 2q + 1).  A product's contraction index may be enumerated in any order, so the two C registers of a lane
// serve directly as the A fragments (and, Pt being symmetric in its x block, as the B fragments) of the next product
// with the contraction split into "even" (k = 2q) and "odd" (k = 2q + 1) halves: NO shuffle or shared-memory round trip
// between the products.  Per stage the chain is 2 + 2 + 1 dependent DMMA, one broadcast of M_uu, the two reciprocal
// square roots and two quad-local shuffles: ~330 cycles.  Everything else of a stage (its [B A], c, H + G, gt) is
// loop-carried-independent shared-memory traffic.  NZ <= 8 (one column tile); the vector row makes a second row tile
// for the slack models (NZ = 8).  Rows / columns of Pt that belong to u hold finite don't-care values: every product
// that contracts over them multiplies by structural zeros of [B A].
// Device only: the host emulation (tests/emu) runs the single-thread sweep, to which this one is algebraically identical.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
MPC_DEV void mpc_dmma(double &d0, double &d1, const double a, const double b, const double c0, const double c1) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1)
        : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// shared-memory accesses of the tensor-core sweep by 32-bit shared-window address: the window base of the CTA is read
// once (the compiler otherwise re-derives it from SR_CgaCtaId inside the stage loop when registers are tight)
MPC_DEV double mpc_lds(unsigned addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
MPC_DEV double mpc_lds_if(bool p, unsigned addr) {  // p ? load : 0.0  (one predicated LDS)
    double v = 0.0;
    asm("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q ld.shared.f64 %0, [%1];\n}" : "+d"(v) : "r"(addr), "r"((int)p));
    return v;
}
MPC_DEV double mpc_sel(bool p, double a, double b) {  // p ? a : b as one select (no divergent branch on lane-dependent p)
    double r;
    asm("{\n .reg .pred q;\n setp.ne.s32 q, %3, 0;\n selp.f64 %0, %1, %2, q;\n}" : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
    return r;
}
MPC_DEV void mpc_sts_if(bool p, unsigned addr, double v) {
    asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.shared.f64 [%0], %1;\n}" ::"r"(addr), "d"(v), "r"((int)p) : "memory");
}

template <class CFG>
MPC_DEV void riccati_factor_mma(double *sm) {
    typedef SL<CFG> S;
    typedef RicIdx<CFG> R;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NE = R::NE;
    constexpr int RT = (NZ + 1 + 7) / 8;              // row tiles: NZ variables + the vector row
    constexpr int VT = NZ / 8, VG = NZ % 8;           // tile / row group of the vector row
    static_assert(NU == 2 && NZ <= 8, "one column tile; input block factorised in closed form for nu = 2");
    const int serial_tid = (int)sm[S::W_BASE + S::W_STID];
    if (((int)threadIdx.x & ~31) == serial_tid) {
        const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
        const unsigned FULL = 0xffffffffu;
        unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
        asm volatile("" : "+r"(sbase));
        // ---- per-lane tables (loop invariant): byte offsets inside a stage image.  A structurally absent entry points at the
        // stage's zero slot (S::ZS), so every load of the stage loop is unconditional; stores keep a flag ----
        // slot s <-> column / contraction index a_s = 2q + s;  its x row l_s = a_s - NU
        int f_off[2], c_off[2], h_off[RT][2], gb_off[RT][2], gs_off[RT][2], gd_off[RT][2], p_off[RT][2], lx_off[RT];
        int flags = 0;          // bit 2/3: c_has[s] (store of P+ c)
        int fl[RT][2], fx[RT];  // per (row tile, slot): bit 4: P / pv entry is stored;  per row tile: own L entry is stored
        double f_id[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int a = 2 * q + s, l = a - NU;
            const bool lv = (l >= 0 && l < NX);
            // [B A]_{l, g}: B (g < NU), else A = I + E with E stored for its non-zero columns (psi, v)
            f_off[s] = S::ZS; f_id[s] = 0.0;
            if (lv && g < NZ) {
                if (g < NU) f_off[s] = S::BB + l * NU + g;
                else {
                    const int j = g - NU, e = (j == 2) ? 0 : ((j == 3) ? 1 : -1);
                    f_id[s] = (l == j) ? 1.0 : 0.0;
                    if (e >= 0) f_off[s] = S::BE + l * NE + e;
                }
            }
            if (lv) flags |= 4 << s;
            c_off[s] = lv ? S::C + l : S::ZS;
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
                const int n = 8 * rt + g;
                int &F = fl[rt][s];
                F = 0;
                h_off[rt][s] = gb_off[rt][s] = gs_off[rt][s] = gd_off[rt][s] = S::ZS;
                p_off[rt][s] = 0;
                if (a < NZ) {
                    if (n < NZ) {
                        h_off[rt][s] = S::H + trisym(n, a);
                        if (n == a) gd_off[rt][s] = S::GD + a;
                        const int an = (n == CFG::IX) ? 0 : (n == CFG::IY) ? 1 : (n == CFG::IPSI) ? 2 : -1;
                        const int aa = (a == CFG::IX) ? 0 : (a == CFG::IY) ? 1 : (a == CFG::IPSI) ? 2 : -1;
                        if (an >= 0 && aa >= 0) gb_off[rt][s] = S::GB + trisym(an, aa);
                        if (CFG::ANY_SLACK_ROWS) {
                            if (n == CFG::ISLACK && aa >= 0) gs_off[rt][s] = S::GS + aa;
                            if (a == CFG::ISLACK && an >= 0) gs_off[rt][s] = S::GS + an;
                            if (n == CFG::ISLACK && a == CFG::ISLACK) gs_off[rt][s] = S::GS + 3;
                        }
                        if (n >= NU && l >= 0 && l <= n - NU) { F |= 16; p_off[rt][s] = S::P + tri(n - NU, l); }  // lower triangle of P
                    } else if (n == NZ) {
                        h_off[rt][s] = S::GT + a;
                        if (lv) { F |= 16; p_off[rt][s] = S::PV + l; }
                    }
                }
            }
        }
#pragma unroll
        for (int rt = 0; rt < RT; rt++) {
            const int n = 8 * rt + g;
            lx_off[rt] = 0;
            fx[rt] = 0;
            if (q < NU) {
                if (n >= NU && n < NZ) { fx[rt] = 1; lx_off[rt] = S::LXU + (n - NU) * NU + q; }
                if (n == NZ) { fx[rt] = 1; lx_off[rt] = S::YV + q; }
            }
        }
        // keep the tables in registers (opaque to the optimiser: it would otherwise re-derive them from the lane id
        // through select chains inside the stage loop)
#pragma unroll
        for (int s = 0; s < 2; s++) {
            f_off[s] *= 8; c_off[s] *= 8;
            mpc_opaque(f_off[s]); mpc_opaque(c_off[s]);
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
                h_off[rt][s] *= 8; gd_off[rt][s] *= 8; gb_off[rt][s] *= 8; gs_off[rt][s] *= 8; p_off[rt][s] *= 8;
                mpc_opaque(h_off[rt][s]); mpc_opaque(gd_off[rt][s]); mpc_opaque(gb_off[rt][s]); mpc_opaque(p_off[rt][s]);
                if (CFG::ANY_SLACK_ROWS) mpc_opaque(gs_off[rt][s]);
                mpc_opaque(fl[rt][s]);
            }
        }
#pragma unroll
        for (int rt = 0; rt < RT; rt++) { lx_off[rt] *= 8; mpc_opaque(lx_off[rt]); mpc_opaque(fx[rt]); }
        mpc_opaque(flags);
        const bool is_vec = (g == VG), lane0 = (lane == 0);
        // ---- terminal node: P_N (eps I after MIRROR), pv_N = gt_N ----
        double Pt[RT][2];
        {
            const double *stN = sm + N * S::STRIDE;
#pragma unroll
            for (int rt = 0; rt < RT; rt++)
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const int n = 8 * rt + g, a = 2 * q + s, l = a - NU;
                    double v = 0.0;
                    if (l >= 0 && l < NX) {
                        if (n >= NU && n < NZ) v = stN[S::P + trisym(n - NU, l)];
                        if (n == NZ) {
                            v = stN[S::GT + a];
                            sm[N * S::STRIDE + S::PV + l] = v;
                        }
                    }
                    Pt[rt][s] = v;
                }
        }
        // stage constants of stage k (independent of the recursion): loaded one stage ahead, all loads issued before
        // any of them is consumed
        double fv[2], cv[2], ci[RT][2];
        double raw[RT][2][4];
        auto load_stage = [&](const unsigned st) {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                fv[s] = mpc_lds(st + f_off[s]);
                cv[s] = mpc_lds(st + c_off[s]);
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    raw[rt][s][0] = mpc_lds(st + h_off[rt][s]);
                    raw[rt][s][1] = mpc_lds(st + gd_off[rt][s]);
                    raw[rt][s][2] = mpc_lds(st + gb_off[rt][s]);
                    raw[rt][s][3] = CFG::ANY_SLACK_ROWS ? mpc_lds(st + gs_off[rt][s]) : 0.0;
                }
            }
        };
        auto finish_stage = [&]() {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                fv[s] += f_id[s];
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    double v = raw[rt][s][0] + raw[rt][s][1];  // same association as the serial sweep: (H + GD) + GB (+ GS)
                    v += raw[rt][s][2];
                    if (CFG::ANY_SLACK_ROWS) v += raw[rt][s][3];
                    ci[rt][s] = v;
                }
            }
        };
        load_stage(sbase + (N - 1) * S::STRIDE * 8);
        finish_stage();
#pragma unroll 1
        for (int k = N - 1; k >= 0; k--) {
            const unsigned st = sbase + k * (S::STRIDE * 8), stn = st + S::STRIDE * 8;
            // GEMM 1: Wt = At Pt' (+ pv+ on the vector row)
            double w[RT][2];
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
                const int n = 8 * rt + g;
                const double a0 = (n < NZ) ? fv[0] : ((n == NZ) ? cv[0] : 0.0);
                const double a1 = (n < NZ) ? fv[1] : ((n == NZ) ? cv[1] : 0.0);
                mpc_dmma(w[rt][0], w[rt][1], a0, Pt[0][0], 0.0, 0.0);
                mpc_dmma(w[rt][0], w[rt][1], a1, Pt[0][1], w[rt][0], w[rt][1]);
            }
            const double fb0 = fv[0], fb1 = fv[1];
            double cc[RT][2];
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
                cc[rt][0] = ci[rt][0];
                cc[rt][1] = ci[rt][1];
            }
            load_stage(st - (k > 0 ? S::STRIDE * 8 : 0));  // next stage's constants: in flight under the products
            // vector row: P+ c goes to shared memory (corrector sweep), then q = pv+ + P+ c
#pragma unroll
            for (int s = 0; s < 2; s++) {
                mpc_sts_if(is_vec && ((flags >> (2 + s)) & 1), stn + (S::PC - NU) * 8 + 16 * q + 8 * s, w[VT][s]);
                w[VT][s] += mpc_sel(is_vec, Pt[VT][s], 0.0);
            }
            // GEMM 2: Mt = Wt [B A] + (H + G | gt)
            double m[RT][2];
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
                mpc_dmma(m[rt][0], m[rt][1], w[rt][0], fb0, cc[rt][0], cc[rt][1]);
                mpc_dmma(m[rt][0], m[rt][1], w[rt][1], fb1, m[rt][0], m[rt][1]);
            }
            finish_stage();
            // L_uu = chol(M_uu): M_00 lane 0 slot 0; M_10, M_11 lane 4 slots 0, 1
            const double m00 = __shfl_sync(FULL, m[0][0], 0), m10 = __shfl_sync(FULL, m[0][0], 4), m11 = __shfl_sync(FULL, m[0][1], 4);
            double t0[RT], t1[RT];
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {  // columns 0, 1 of the own row: slots of the quad's first lane
                t0[rt] = __shfl_sync(FULL, m[rt][0], lane & ~3);
                t1[rt] = __shfl_sync(FULL, m[rt][1], lane & ~3);
            }
            double il00, rdet;
            mpc_rsqrt_pair(m00, m11 * m00 - m10 * m10, il00, rdet);
            const double l10 = m10 * il00;
            const double il11 = rdet * (m00 * il00);
            mpc_sts_if(lane0, st + S::LUU * 8, il00);
            mpc_sts_if(lane0, st + S::LUU * 8 + 8, l10);
            mpc_sts_if(lane0, st + S::LUU * 8 + 16, il11);
            // rows of L_xu (x rows), y (vector row): lane (g, q) keeps column q
            double Lq[RT];
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
                const double L0 = t0[rt] * il00;
                const double L1 = (t1[rt] - L0 * l10) * il11;
                Lq[rt] = mpc_sel(q == 1, L1, mpc_sel(q == 0, L0, 0.0));
                // what the vector sweeps read (riccati_solve): x rows  Kt = L_xu L_uu^-1  (the feedback gain, transposed:
                // du = f - Kt' dx), vector row  y = L_uu^-1 m_u
                const double K1 = L1 * il11;
                const double K0 = (L0 - K1 * l10) * il00;
                const bool yrow = (rt == VT) && is_vec;
                mpc_sts_if(fx[rt] != 0, st + lx_off[rt], mpc_sel(yrow, Lq[rt], mpc_sel(q == 1, K1, K0)));
            }
            // GEMM 3: Pt = Mt - L L'
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
                mpc_dmma(Pt[rt][0], Pt[rt][1], -Lq[rt], Lq[0], m[rt][0], m[rt][1]);
#pragma unroll
                for (int s = 0; s < 2; s++) mpc_sts_if(fl[rt][s] & 16, st + p_off[rt][s], Pt[rt][s]);
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Vector sweeps on the FP64 tensor-core path (one warp; device only, pairs with riccati_factor_mma) -- a MEASURED
// ALTERNATIVE kept under -DMPC_SOLVE_MMA, not the shipped path: parity with the oracle 5.5e-11, but 15 % slower than the
// single-thread sweeps (bench batch 7.57 vs 6.58 ms, one control step 0.799 vs 0.680 ms).
//
// A lone thread runs the shipped sweeps (riccati_solve): 19 % of the kernel's warp instructions at 1 of 32 lanes, a quarter
// of a solve's time under load, ~170 cycles per stage and sweep.  With the feedback gain stored as Kt = L_xu L_uu^-1
// (du = f - Kt' dx with f = -L_uu^-T y) both sweeps are ONE vector-matrix product per stage with the closed-loop
// matrix Acl = A - B Kt':
//   backward   [pv_k | m_u,k]   = [gt_x | 0] + [q_k+1 | gt_u,k] Mb_k ,   q_k = pv_k + P_k c_k-1 ,   y_k = L_uu^-1 m_u,k
//              Mb = [ Acl   B ]      (pv = gt_x - Kt gt_u + Acl' q:  the recursion  m = gt + [B A]' q,  y = L_uu^-1 m_u,
//                   [ -Kt'  I ]       pv = m_x - L_xu y  of riccati_solve with y eliminated;  m_u = gt_u + B' q)
//   forward    [dx_k+1 | du_k]  = [c_k | 0] + [dx_k | f_k] Mf_k
//              Mf = [ Acl'  -Kt ]
//                   [ B'     I  ]
// The row vector is row 0 of an m8n8k4 tile (columns: NX states, then the NU inputs); with the contraction enumerated
// "even / odd" the two C registers of a lane are the A fragments of the next product (see riccati_factor_mma): the chain
// per stage is two dependent DMMA (+ one add in the backward sweep).  Every lane forms its two matrix entries
// (2q + s, g) of the stage from [B A], Kt through per-lane offset tables (absent operands read the stage's zero slot), one
// stage ahead of the chain.  Same algebra as the serial sweeps up to rounding.  Why it loses: the chain is short (two DMMA),
// but a warp issues in order and forming the entries on the fly costs ~17 LDS + 6 FMA per stage on top of the products --
// ~50 warp instructions per stage and sweep against the lone thread's ~70 cheap ones, ~310 cycles against ~170.  Storing the
// closed-loop matrix (25 doubles per stage) would fix that; the stage image has 2 doubles left at four solves per SM.
// ---------------------------------------------------------------------------------------------
MPC_DEV double mpc_lds_v(unsigned addr) {  // load of data written earlier in the same function by another lane
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

template <class CFG>
MPC_DEV void riccati_solve_mma(double *sm, bool BACKWARD) {
    typedef SL<CFG> S;
    typedef RicIdx<CFG> R;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NE = R::NE;
    static_assert(NU == 2 && NZ <= 8, "one tile: NX state columns + NU input columns");
    constexpr int SB = S::STRIDE * 8;  // bytes per stage image
    const int serial_tid = (int)sm[S::W_BASE + S::W_STID];
    if (((int)threadIdx.x & ~31) == serial_tid) {
        const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
        const unsigned FULL = 0xffffffffu;
        unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
        asm volatile("" : "+r"(sbase));
        // row-vector bookkeeping of the own two slots (columns 2q, 2q + 1 of row g; only row 0 carries data)
        int xcol[2], ins[2];  // byte offset of the column inside an NX-vector (-1: not a state column of row 0); 1 / 2: input column 0 / 1
#pragma unroll
        for (int s2 = 0; s2 < 2; s2++) {
            const int col = 2 * q + s2;
            xcol[s2] = (g == 0 && col < NX) ? 8 * col : -1;
            ins[s2] = (g == 0 && col >= NX && col < NZ) ? 1 + (col - NX) : 0;
            mpc_opaque(xcol[s2]);
            mpc_opaque(ins[s2]);
        }
        // matrix entry (i = 2q + s, n = g) of the stage:  id + sgn * ld(eo) - ld(b0) ld(k0) - ld(b1) ld(k1)
        auto tables = [&](const bool fwd, int (&eo)[2], int (&b0)[2], int (&k0)[2], int (&b1)[2], int (&k1)[2], double (&idv)[2], double (&sg)[2]) {
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                const int i = 2 * q + s2, n = g;
                eo[s2] = b0[s2] = k0[s2] = b1[s2] = k1[s2] = S::ZS;
                idv[s2] = 0.0;
                sg[s2] = 1.0;
                if (i < NX && n < NX) {
                    const int r = fwd ? n : i, c = fwd ? i : n;  // Acl[r][c]; the forward sweep multiplies by Acl'
                    idv[s2] = (r == c) ? 1.0 : 0.0;
                    const int e = R::eidx(c);
                    if (e >= 0) eo[s2] = S::BE + r * NE + e;
                    b0[s2] = S::BB + r * NU;
                    b1[s2] = S::BB + r * NU + 1;
                    k0[s2] = S::LXU + c * NU;
                    k1[s2] = S::LXU + c * NU + 1;
                } else if (i >= NX && i < NZ && n < NX) {
                    if (fwd) eo[s2] = S::BB + n * NU + (i - NX);
                    else { eo[s2] = S::LXU + n * NU + (i - NX); sg[s2] = -1.0; }
                } else if (i < NX && n >= NX && n < NZ) {
                    if (fwd) { eo[s2] = S::LXU + i * NU + (n - NX); sg[s2] = -1.0; }
                    else eo[s2] = S::BB + i * NU + (n - NX);
                } else if (i >= NX && i < NZ && n >= NX && n < NZ) {
                    idv[s2] = (i == n) ? 1.0 : 0.0;
                }
                eo[s2] *= 8; b0[s2] *= 8; k0[s2] *= 8; b1[s2] *= 8; k1[s2] *= 8;
                mpc_opaque(eo[s2]); mpc_opaque(b0[s2]); mpc_opaque(k0[s2]); mpc_opaque(b1[s2]); mpc_opaque(k1[s2]);
            }
        };
        int eo[2], b0[2], k0[2], b1[2], k1[2];
        double idv[2], sg[2];
        double x[2];
        if (BACKWARD) {
            tables(false, eo, b0, k0, b1, k1, idv, sg);
            // q_N = gt_N + P_N c_N-1 (pv_N = gt_N), input columns: gt_u of stage N - 1
            {
                const unsigned stN = sbase + N * SB, stp = stN - SB;
                const double u0 = mpc_lds(stp + S::GT * 8), u1 = mpc_lds(stp + S::GT * 8 + 8);
#pragma unroll
                for (int s2 = 0; s2 < 2; s2++) {
                    const bool xc = xcol[s2] >= 0;
                    const double gtn = mpc_lds(xc ? stN + (S::GT + NU) * 8 + xcol[s2] : stN + S::ZS * 8);
                    const double pc = mpc_lds(xc ? stN + S::PC * 8 + xcol[s2] : stN + S::ZS * 8);
                    mpc_sts_if(xc, stN + S::PV * 8 + (xc ? xcol[s2] : 0), gtn);
                    x[s2] = mpc_sel(ins[s2] == 1, u0, mpc_sel(ins[s2] == 2, u1, gtn + pc));
                }
            }
            double m[2], cf[2], pcv[2], lu[3], gu[2];
            auto load_b = [&](const unsigned st, const unsigned stp) {
#pragma unroll
                for (int s2 = 0; s2 < 2; s2++) {
                    double e = fma(sg[s2], mpc_lds(st + eo[s2]), idv[s2]);
                    e = fma(-mpc_lds(st + b0[s2]), mpc_lds(st + k0[s2]), e);
                    e = fma(-mpc_lds(st + b1[s2]), mpc_lds(st + k1[s2]), e);
                    m[s2] = e;
                    const bool xc = xcol[s2] >= 0;
                    cf[s2] = mpc_lds(xc ? st + (S::GT + NU) * 8 + xcol[s2] : st + S::ZS * 8);
                    pcv[s2] = mpc_lds(xc ? st + S::PC * 8 + xcol[s2] : st + S::ZS * 8);
                }
                lu[0] = mpc_lds(st + S::LUU * 8);
                lu[1] = mpc_lds(st + S::LUU * 8 + 8);
                lu[2] = mpc_lds(st + S::LUU * 8 + 16);
                gu[0] = mpc_lds(stp + S::GT * 8);  // gt_u of the NEXT stage of the sweep (k - 1)
                gu[1] = mpc_lds(stp + S::GT * 8 + 8);
            };
            load_b(sbase + (N - 1) * SB, sbase + (N > 1 ? N - 2 : 0) * SB);
#pragma unroll 1
            for (int k = N - 1; k >= 0; k--) {
                const unsigned st = sbase + k * SB;
                const double m0 = m[0], m1 = m[1], c0 = cf[0], c1 = cf[1], p0 = pcv[0], p1 = pcv[1];
                const double il00 = lu[0], l10 = lu[1], il11 = lu[2], u0 = gu[0], u1 = gu[1];
                double d0, d1;
                mpc_dmma(d0, d1, x[0], m0, c0, c1);
                mpc_dmma(d0, d1, x[1], m1, d0, d1);
                if (k > 0) load_b(st - SB, st - (k > 1 ? 2 : 1) * SB);  // next stage's entries: in flight under the products
                mpc_sts_if(xcol[0] >= 0, st + S::PV * 8 + (xcol[0] >= 0 ? xcol[0] : 0), d0);
                mpc_sts_if(xcol[1] >= 0, st + S::PV * 8 + (xcol[1] >= 0 ? xcol[1] : 0), d1);
                // y = L_uu^-1 m_u: m_u sits in columns NX, NX + 1 of row 0
                const double mu0 = __shfl_sync(FULL, (NX & 1) ? d1 : d0, NX >> 1);
                const double mu1 = __shfl_sync(FULL, ((NX + 1) & 1) ? d1 : d0, (NX + 1) >> 1);
                const double y0 = mu0 * il00;
                const double y1 = (mu1 - l10 * y0) * il11;
                mpc_sts_if(lane == 0, st + S::YV * 8, y0);
                mpc_sts_if(lane == 0, st + S::YV * 8 + 8, y1);
                // next row vector: q_k = pv_k + P_k c_k-1 in the state columns, gt_u of stage k - 1 in the input columns
                x[0] = mpc_sel(ins[0] == 1, u0, mpc_sel(ins[0] == 2, u1, d0 + p0));
                x[1] = mpc_sel(ins[1] == 1, u0, mpc_sel(ins[1] == 2, u1, d1 + p1));
            }
            __syncwarp();
        }
        // ---- forward: [dx_k+1 | du_k] = [c_k | 0] + [dx_k | f_k] Mf_k,  f = -L_uu^-T y ----
        tables(true, eo, b0, k0, b1, k1, idv, sg);
        double m[2], cf[2], fn[2];
        auto load_f = [&](const unsigned st) {
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                double e = fma(sg[s2], mpc_lds(st + eo[s2]), idv[s2]);
                e = fma(-mpc_lds(st + b0[s2]), mpc_lds(st + k0[s2]), e);
                e = fma(-mpc_lds(st + b1[s2]), mpc_lds(st + k1[s2]), e);
                m[s2] = e;
                cf[s2] = mpc_lds(xcol[s2] >= 0 ? st + S::C * 8 + xcol[s2] : st + S::ZS * 8);
            }
        };
        auto load_fvec = [&](const unsigned st) {  // f of stage st (y may have been written by the backward sweep above)
            const double y0 = mpc_lds_v(st + S::YV * 8), y1 = mpc_lds_v(st + S::YV * 8 + 8);
            const double il00 = mpc_lds(st + S::LUU * 8), l10 = mpc_lds(st + S::LUU * 8 + 8), il11 = mpc_lds(st + S::LUU * 8 + 16);
            fn[1] = -(y1 * il11);
            fn[0] = -((y0 + l10 * fn[1]) * il00);
        };
        {
            const unsigned W = sbase + S::W_BASE * 8;
            load_fvec(sbase);
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                const bool xc = xcol[s2] >= 0;
                const double v = mpc_lds(xc ? W + S::W_DX0 * 8 + xcol[s2] : sbase + S::ZS * 8);
                mpc_sts_if(xc, sbase + (S::DZN + NU) * 8 + (xc ? xcol[s2] : 0), v);
                x[s2] = mpc_sel(ins[s2] == 1, fn[0], mpc_sel(ins[s2] == 2, fn[1], v));
            }
            load_f(sbase);
            if (N > 1) load_fvec(sbase + SB);
        }
#pragma unroll 1
        for (int k = 0; k < N; k++) {
            const unsigned st = sbase + k * SB, stn = st + SB;
            const double m0 = m[0], m1 = m[1], c0 = cf[0], c1 = cf[1], f0 = fn[0], f1 = fn[1];
            double d0, d1;
            mpc_dmma(d0, d1, x[0], m0, c0, c1);
            mpc_dmma(d0, d1, x[1], m1, d0, d1);
            if (k + 1 < N) {
                load_f(stn);
                if (k + 2 < N) load_fvec(stn + SB);
            }
            // dx_k+1 -> DZN of the next stage, du_k -> DZN of this stage
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                const double d = s2 ? d1 : d0;
                mpc_sts_if(xcol[s2] >= 0, stn + (S::DZN + NU) * 8 + (xcol[s2] >= 0 ? xcol[s2] : 0), d);
                mpc_sts_if(ins[s2] != 0, st + S::DZN * 8 + (ins[s2] == 2 ? 8 : 0), d);
            }
            x[0] = mpc_sel(ins[0] == 1, f0, mpc_sel(ins[0] == 2, f1, d0));
            x[1] = mpc_sel(ins[1] == 1, f0, mpc_sel(ins[1] == 2, f1, d1));
        }
    }
    __syncthreads();
}
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// Dense Riccati sweeps for models outside the unicycle family (bicycle: nx = 6, nu = 3, A = I + E with three columns):
// the same recursion as riccati_factor / riccati_solve with every entry of [B A] counted, a general nu x nu Cholesky and
// run-time loops; one thread (BASELINE configs[0] is a P = 1 problem: latency of the path matters little next to T-MPC).
// ---------------------------------------------------------------------------------------------
template <class CFG>
MPC_DEV void riccati_factor_dense(double *sm) {
    typedef SL<CFG> S;
    typedef RicIdx<CFG> R;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NE = R::NE;
    static_assert(!CFG::ANY_SLACK_ROWS || !CFG::BICYCLE, "slack rows with the bicycle's slack input: not built");
    const int serial_tid = (int)sm[S::W_BASE + S::W_STID];
    MPC_PHASE
    if (tid == serial_tid) {
        double Pn[NX][NX], pv[NX];
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NX; j++) Pn[i][j] = sm[N * S::STRIDE + S::P + trisym(i, j)];
        for (int i = 0; i < NX; i++) {
            pv[i] = sm[N * S::STRIDE + S::GT + NU + i];  // terminal: pv_N = gt_N
            sm[N * S::STRIDE + S::PV + i] = pv[i];
        }
        for (int k = N - 1; k >= 0; k--) {
            double *st = sm + k * S::STRIDE;
            double *stn = st + S::STRIDE;
            double F[NX][NZ], W[NX][NZ], q[NX], M[NZ][NZ], m[NZ];
            for (int l = 0; l < NX; l++) {
                for (int u = 0; u < NU; u++) F[l][u] = st[S::BB + l * NU + u];
                for (int j = 0; j < NX; j++) F[l][NU + j] = ((l == j) ? 1.0 : 0.0) + (R::eidx(j) >= 0 ? st[S::BE + l * NE + (R::eidx(j) >= 0 ? R::eidx(j) : 0)] : 0.0);
            }
            for (int l = 0; l < NX; l++) {
                double pc = 0.0;
                for (int j = 0; j < NX; j++) pc += Pn[l][j] * st[S::C + j];
                stn[S::PC + l] = pc;
                q[l] = pv[l] + pc;
                for (int a = 0; a < NZ; a++) {
                    double acc = 0.0;
                    for (int j = 0; j < NX; j++) acc += Pn[l][j] * F[j][a];
                    W[l][a] = acc;
                }
            }
            for (int a = 0; a < NZ; a++) {
                for (int b = 0; b <= a; b++) {
                    double acc = st[S::H + tri(a, b)];
                    if (a == b) acc += st[S::GD + a];
                    const int aa = (a == CFG::IX) ? 0 : (a == CFG::IY) ? 1 : (a == CFG::IPSI) ? 2 : -1;
                    const int ab = (b == CFG::IX) ? 0 : (b == CFG::IY) ? 1 : (b == CFG::IPSI) ? 2 : -1;
                    if (aa >= 0 && ab >= 0) acc += st[S::GB + trisym(aa, ab)];
                    for (int l = 0; l < NX; l++) acc += F[l][a] * W[l][b];
                    M[a][b] = acc;
                    M[b][a] = acc;
                }
                double mv = st[S::GT + a];
                for (int l = 0; l < NX; l++) mv += F[l][a] * q[l];
                m[a] = mv;
            }
            // L_uu = chol(M_uu): packed lower with reciprocal diagonal
            double Lu[NU][NU];
            for (int i = 0; i < NU; i++)
                for (int j = 0; j <= i; j++) {
                    double acc = M[i][j];
                    for (int c = 0; c < j; c++) acc -= Lu[i][c] * Lu[j][c];
                    Lu[i][j] = (i == j) ? mpc_rsqrt(acc) : acc * Lu[j][j];  // diagonal entries hold 1 / l_ii
                    st[S::LUU + tri(i, j)] = Lu[i][j];
                }
            double y[NU];
            for (int i = 0; i < NU; i++) {
                double acc = m[i];
                for (int c = 0; c < i; c++) acc -= Lu[i][c] * y[c];
                y[i] = acc * Lu[i][i];
                st[S::YV + i] = y[i];
            }
            double Lx[NX][NU];
            for (int i = 0; i < NX; i++) {
                for (int j = 0; j < NU; j++) {
                    double acc = M[NU + i][j];
                    for (int c = 0; c < j; c++) acc -= Lx[i][c] * Lu[j][c];
                    Lx[i][j] = acc * Lu[j][j];
                    st[S::LXU + i * NU + j] = Lx[i][j];
                }
                double p = m[NU + i];
                for (int c = 0; c < NU; c++) p -= Lx[i][c] * y[c];
                pv[i] = p;
                st[S::PV + i] = p;
            }
            for (int i = 0; i < NX; i++)
                for (int j = 0; j <= i; j++) {
                    double acc = M[NU + i][NU + j];
                    for (int c = 0; c < NU; c++) acc -= Lx[i][c] * Lx[j][c];
                    Pn[i][j] = acc;
                    Pn[j][i] = acc;
                    st[S::P + tri(i, j)] = acc;
                }
        }
    }
    MPC_PHASE_END
}

template <class CFG>
MPC_DEV void riccati_solve_dense(double *sm, bool BACKWARD) {
    typedef SL<CFG> S;
    typedef RicIdx<CFG> R;
    constexpr int N = CFG::N, NX = CFG::NX, NU = CFG::NU, NE = R::NE;
    double *W = sm + S::W_BASE;
    const int serial_tid = (int)W[S::W_STID];
    MPC_PHASE
    if (tid == serial_tid) {
        if (BACKWARD) {
            double pv[NX];
            for (int i = 0; i < NX; i++) {
                pv[i] = sm[N * S::STRIDE + S::GT + NU + i];
                sm[N * S::STRIDE + S::PV + i] = pv[i];
            }
            for (int k = N - 1; k >= 0; k--) {
                double *st = sm + k * S::STRIDE;
                const double *stn = st + S::STRIDE;
                double q[NX], y[NU];
                for (int l = 0; l < NX; l++) q[l] = pv[l] + stn[S::PC + l];
                for (int i = 0; i < NU; i++) {
                    double acc = st[S::GT + i];
                    for (int l = 0; l < NX; l++) acc += st[S::BB + l * NU + i] * q[l];
                    for (int c = 0; c < i; c++) acc -= st[S::LUU + tri(i, c)] * y[c];
                    y[i] = acc * st[S::LUU + tri(i, i)];
                    st[S::YV + i] = y[i];
                }
                for (int i = 0; i < NX; i++) {
                    double mx = st[S::GT + NU + i] + q[i];
                    if (R::eidx(i) >= 0)
                        for (int l = 0; l < NX; l++) mx += st[S::BE + l * NE + (R::eidx(i) >= 0 ? R::eidx(i) : 0)] * q[l];
                    for (int c = 0; c < NU; c++) mx -= st[S::LXU + i * NU + c] * y[c];
                    pv[i] = mx;
                    st[S::PV + i] = mx;
                }
            }
        }
        double dx[NX];
        for (int i = 0; i < NX; i++) {
            dx[i] = W[S::W_DX0 + i];
            sm[S::DZN + NU + i] = dx[i];
        }
        for (int k = 0; k < N; k++) {
            double *st = sm + k * S::STRIDE;
            double *stn = st + S::STRIDE;
            double w[NU], du[NU], dxn[NX];
            for (int c = 0; c < NU; c++) {
                double acc = st[S::YV + c];
                for (int l = 0; l < NX; l++) acc += st[S::LXU + l * NU + c] * dx[l];
                w[c] = acc;
            }
            for (int c = NU - 1; c >= 0; c--) {  // du = -L_uu^-T w
                double acc = w[c];
                for (int r = c + 1; r < NU; r++) acc -= st[S::LUU + tri(r, c)] * du[r];
                du[c] = acc * st[S::LUU + tri(c, c)];
            }
            for (int c = 0; c < NU; c++) {
                du[c] = du[c];
                st[S::DZN + c] = -du[c];
            }
            for (int i = 0; i < NX; i++) {
                double acc = st[S::C + i] + dx[i];
                for (int c = 0; c < NU; c++) acc -= st[S::BB + i * NU + c] * du[c];
                for (int c = 0; c < NE; c++) acc += st[S::BE + i * NE + c] * dx[R::ecol(c)];
                dxn[i] = acc;
                stn[S::DZN + NU + i] = acc;
            }
            for (int i = 0; i < NX; i++) dx[i] = dxn[i];
        }
    }
    MPC_PHASE_END
}

// (Measured alternative: the same sweeps with every run of consecutive doubles moved 16 bytes at a time -- LDS.128 / STS.128
// where the odd stage stride leaves the address aligned, parity-specialised (odd, even) stage pairs: a third fewer loads, no
// gain (bench batch 6.11 -> 6.16 ms, one control step 0.640 -> 0.648 ms): the lone thread waits on its FP64 chain, not on the
// load instructions.)
// Backward vector sweep + forward sweep for the gradient GT:  writes DZN (all nodes) and PV.
template <class CFG>
MPC_DEV void riccati_solve(double *sm, bool BACKWARD) {
    typedef SL<CFG> S;
    typedef RicIdx<CFG> R;
    constexpr int N = CFG::N, NX = CFG::NX, NU = CFG::NU, NE = R::NE;
    double *W = sm + S::W_BASE;
    const int serial_tid = (int)W[S::W_STID];
    MPC_PHASE
    if (tid == serial_tid) {
        // backward: m = gt + BA'(pv+ + P+ c);  pv = m_x - Kt m_u  (Kt = L_xu L_uu^-1, stored by the factorisation);  y = Luu^-1 m_u
        // (BACKWARD = false: riccati_factor_lanes already ran this sweep for the current gradient)
        if (BACKWARD) {
        double pv[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) {
            pv[i] = sm[N * S::STRIDE + S::GT + NU + i];  // terminal: pv_N = gt_N
            sm[N * S::STRIDE + S::PV + i] = pv[i];
        }
#pragma unroll 2
        for (int k = N - 1; k >= 0; k--) {
            double *st = sm + k * S::STRIDE;
            const double *stn = st + S::STRIDE;
            double q[NX];
#pragma unroll
            for (int l = 0; l < NX; l++) q[l] = pv[l] + stn[S::PC + l];
            double m0 = st[S::GT + 0], m1 = st[S::GT + 1];
#pragma unroll
            for (int l = 0; l < NX; l++) {
                if (R::bnz(l, 0)) m0 += st[S::BB + l * NU + 0] * q[l];
                if (R::bnz(l, 1)) m1 += st[S::BB + l * NU + 1] * q[l];
            }
            // y = L_uu^-1 m_u is only stored (the forward sweep starts from it): the recursion itself continues from m_u
            // through the gain, pv = m_x - Kt m_u (= m_x - L_xu y), three links of the dependent chain shorter
            const double y0 = m0 * st[S::LUU];
            const double y1 = (m1 - st[S::LUU + 1] * y0) * st[S::LUU + 2];
            st[S::YV] = y0;
            st[S::YV + 1] = y1;
#pragma unroll
            for (int i = 0; i < NX; i++) {
                double mx = st[S::GT + NU + i] + q[i];  // A = I + E
                if (R::eidx(i) >= 0) {
#pragma unroll
                    for (int l = 0; l < NX; l++)
                        if (R::enz(l, R::eidx(i) >= 0 ? R::eidx(i) : 0)) mx += st[S::BE + l * NE + (R::eidx(i) >= 0 ? R::eidx(i) : 0)] * q[l];
                }
                pv[i] = mx - st[S::LXU + i * NU] * m0 - st[S::LXU + i * NU + 1] * m1;
                st[S::PV + i] = pv[i];
            }
        }
        }
        // forward: du = -Luu^-T y - Kt' dx;  dx+ = c + B du + A dx
        double dx[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) {
            dx[i] = W[S::W_DX0 + i];
            sm[S::DZN + NU + i] = dx[i];
        }
#pragma unroll 2
        for (int k = 0; k < N; k++) {
            double *st = sm + k * S::STRIDE;
            double *stn = st + S::STRIDE;
            // du = f - Kt' dx with f = -L_uu^-T y (independent of dx: off the chain);  dx+ = (c + A dx) + B du, the bracket in
            // parallel with du
            const double il00 = st[S::LUU], l10 = st[S::LUU + 1], il11 = st[S::LUU + 2];
            const double f1 = -(st[S::YV + 1] * il11);
            const double f0 = -((st[S::YV] + l10 * f1) * il00);
            double du0 = f0, du1 = f1;
#pragma unroll
            for (int l = 0; l < NX; l++) {
                du0 -= st[S::LXU + l * NU] * dx[l];
                du1 -= st[S::LXU + l * NU + 1] * dx[l];
            }
            st[S::DZN] = du0;
            st[S::DZN + 1] = du1;
            double dxn[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) {
                double acc = st[S::C + i] + dx[i];
#pragma unroll
                for (int c = 0; c < NE; c++)
                    if (R::enz(i, c)) acc += st[S::BE + i * NE + c] * dx[R::ecol(c)];
                if (R::bnz(i, 0)) acc += st[S::BB + i * NU] * du0;
                if (R::bnz(i, 1)) acc += st[S::BB + i * NU + 1] * du1;
                dxn[i] = acc;
                stn[S::DZN + NU + i] = acc;
            }
#pragma unroll
            for (int i = 0; i < NX; i++) dx[i] = dxn[i];
        }
    }
    MPC_PHASE_END
    // the equality-multiplier target pi_k = P_k dx_k + pv_k is formed where it is consumed (the step phase PA of ipm_solve)
}

// ---------------------------------------------------------------------------------------------
// Per-stage accumulations over the h rows (thread = stage x variable; the x, y, psi threads
// carry the sums over the h rows).
//   MODE 0: after an update: stationarity / equality residuals of the current iterate, barrier
//           Hessian terms GD/GB/GS and the affine gradient GT (sigma mu = 0, corr = 0)
//   MODE 1: corrector gradient: GT -= Jr' ((sigma mu - corr) / t)
// ---------------------------------------------------------------------------------------------
template <class CFG, int MODE>
MPC_DEV void stage_accumulate(const Layout &L, double *sm, double sigma_mu) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NH = CFG::NH, NT = MPC_NT;
    double *W = sm + S::W_BASE;
    MPC_PHASE
    {
        double rs = 0.0, re = 0.0;
        for (int it = tid; it < (N + 1) * NZ; it += NT) {
            // item = (variable slot v, stage k), stage fastest; the heavy variables (x, y, psi) take slots 0..2 so that they
            // spread over distinct threads, the remaining variables follow in their own order
            const int v = it / (N + 1), k = it - v * (N + 1);
            const int i = (v < 3) ? CFG::IX + v : ((v - 3 < CFG::IX) ? v - 3 : v);
            if (k == N && i < NU) continue;
            double *st = sm + k * S::STRIDE;
            const int a = i - CFG::IX;  // 0,1,2 for x,y,psi
            const bool heavy = (a >= 0 && a < 3);
            double gt = (MODE == 0) ? ((k < N) ? st[S::GR + i] : 0.0) : st[S::GT + i];
            double rg = 0.0, gd = 0.0;
            if (MODE == 0) {
                // (H dz + g)_i + (BA' dpi+)_i - dpi_i
                if (k < N) {
                    rg = st[S::GR + i];
#pragma unroll
                    for (int j = 0; j < NZ; j++) rg += st[S::H + trisym(i, j)] * st[S::DZS + j];
                    const double *stn = sm + (k + 1) * S::STRIDE;
                    // (BA' dpi+)_i with A = I + E
                    if (i < NU) {
#pragma unroll
                        for (int l = 0; l < NX; l++) rg += st[S::BB + l * NU + i] * stn[S::DPI + l];
                    } else {
                        rg += stn[S::DPI + i - NU];
                        const int ce = RicIdx<CFG>::eidx(i - NU);
                        if (ce >= 0) {
#pragma unroll
                            for (int l = 0; l < NX; l++) rg += st[S::BE + l * S::NE + ce] * stn[S::DPI + l];
                        }
                    }
                } else {
                    rg = L.reg_epsilon * st[S::DZS + i];  // H_N = eps I, g_N = 0
                }
                if (k >= 1 && i >= NU) rg -= st[S::DPI + i - NU];
            }
            // box rows of variable i
            if (box_active<CFG>(L, k, i, false)) {
                const double l = st[S::LAM + i], ti = st[S::TINV + i];
                if (MODE == 0) {
                    const double gam = l * ti;
                    rg -= l;
                    gd += gam;
                    gt -= l - gam * (st[S::Z + i] - L.lower_bound[i] + L.qp_relax);
                } else {
                    gt -= (sigma_mu - st[S::CORR + i]) * ti;
                }
            }
            if (box_active<CFG>(L, k, i, true)) {
                const double l = st[S::LAM + NZ + i], ti = st[S::TINV + NZ + i];
                if (MODE == 0) {
                    const double gam = l * ti;
                    rg += l;
                    gd += gam;
                    gt += l - gam * (L.upper_bound[i] - st[S::Z + i] + L.qp_relax);
                } else {
                    gt += (sigma_mu - st[S::CORR + NZ + i]) * ti;
                }
            }
            // h rows
            if (NH > 0 && k >= 1 && k < N) {
                if (heavy) {
                    double gb0 = 0.0, gb1 = 0.0, gb2 = 0.0, gs = 0.0;
#pragma unroll 4
                    for (int j = 0; j < NH; j++) {
                        const double l = st[S::LAM + 2 * NZ + j], ti = st[S::TINV + 2 * NZ + j];
                        const double ja = st[S::JR + 3 * j + a];
                        if (MODE == 0) {
                            const double gam = l * ti;
                            rg -= ja * l;
                            gt -= ja * (l - gam * st[S::RH + j]);
                            const double gja = gam * ja;
                            // row a of the packed (x,y,psi) block: entries (a,0..a)
                            gb0 += gja * st[S::JR + 3 * j];
                            if (a >= 1) gb1 += gja * st[S::JR + 3 * j + 1];
                            if (a >= 2) gb2 += gja * ja;
                            if (CFG::ANY_SLACK_ROWS && row_has_slack<CFG>(j)) gs += gja;
                        } else {
                            gt -= ja * (sigma_mu - st[S::CORR + 2 * NZ + j]) * ti;
                        }
                    }
                    if (MODE == 0) {
                        st[S::GB + tri(a, 0)] = gb0;
                        if (a >= 1) st[S::GB + tri(a, 1)] = gb1;
                        if (a >= 2) st[S::GB + tri(a, 2)] = gb2;
                        if (CFG::ANY_SLACK_ROWS) st[S::GS + a] = gs;
                    }
                } else if (CFG::ANY_SLACK_ROWS && i == CFG::ISLACK) {
                    double gss = 0.0;
                    for (int j = 0; j < NH; j++) {
                        if (!row_has_slack<CFG>(j)) continue;
                        const double l = st[S::LAM + 2 * NZ + j], ti = st[S::TINV + 2 * NZ + j];
                        if (MODE == 0) {
                            const double gam = l * ti;
                            rg -= l;
                            gt -= l - gam * st[S::RH + j];
                            gss += gam;
                        } else {
                            gt -= (sigma_mu - st[S::CORR + 2 * NZ + j]) * ti;
                        }
                    }
                    if (MODE == 0) st[S::GS + 3] = gss;
                }
            } else if (MODE == 0 && heavy) {
                st[S::GB + tri(a, 0)] = 0.0;
                if (a >= 1) st[S::GB + tri(a, 1)] = 0.0;
                if (a >= 2) st[S::GB + tri(a, 2)] = 0.0;
                if (CFG::ANY_SLACK_ROWS) st[S::GS + a] = 0.0;
            } else if (MODE == 0 && CFG::ANY_SLACK_ROWS && i == CFG::ISLACK) {
                st[S::GS + 3] = 0.0;
            }
            st[S::GT + i] = gt;
            if (MODE == 0) {
                st[S::DZ + i] = st[S::DZS + i];  // the stepped iterate becomes the iterate (every read of this phase goes to DZS)
                st[S::GD + i] = gd;
                if (!(k == 0 && i >= NU)) upd_max(rs, fabs(rg));  // x0 is eliminated
                if (k < N && i >= NU) {                            // equality residual of row i - NU
                    const int r = i - NU;
                    const double *stn = sm + (k + 1) * S::STRIDE;
                    double acc = st[S::C + r] - stn[S::DZS + NU + r];
#pragma unroll
                    for (int u = 0; u < NU; u++) acc += st[S::BB + r * NU + u] * st[S::DZS + u];
                    acc += st[S::DZS + NU + r] + st[S::BE + r * S::NE] * st[S::DZS + NU + 2] + st[S::BE + r * S::NE + 1] * st[S::DZS + NU + 3];
                    if (S::NE > 2) acc += st[S::BE + r * S::NE + (S::NE > 2 ? 2 : 0)] * st[S::DZS + NU + 4];
                    upd_max(re, fabs(acc));
                }
            }
        }
        if (MODE == 0) {
            red_put<CFG>(W, tid, 0, 0, rs);
            red_put<CFG>(W, tid, 1, 0, re);
        }
    }
    MPC_PHASE_END
}

// ---------------------------------------------------------------------------------------------
// Row phases of the IPM: thread = (stage, variable) for the two box rows of a variable and
// thread = (stage, h row) for the general rows -- class-specific straight-line code, no divergence
// between row kinds inside a warp.
//   RP_UPDATE : (init | step) of t, lam; 1/t; inequality / complementarity residuals, sum lam t
//   RP_AFFINE : ratio test of the affine step, centering sums S1, S2, second-order term
//   RP_STEP   : ratio test of the combined step
// ---------------------------------------------------------------------------------------------
enum { RP_UPDATE = 0, RP_AFFINE = 1, RP_STEP = 2 };

struct RowCtx {
    // inputs
    bool first, warm;
    double alpha, sigma_mu, thr0, mu0, ws_thr, ws_mu;
    // accumulators
    double a0, a1, a2;
};

// fmax() with the second operand known not to be NaN, as one compare + select (fmax proper carries NaN handling for both
// operands: it was half of the row bodies' instructions).  Same value as fmax(v, floor) for every v, NaN included.
MPC_HD double mpc_floor_at(double v, double floor) { return (v > floor) ? v : floor; }
// acc = fmax(acc, v) for a running maximum that is never NaN itself (starts at 0, only ever takes a compared-greater value):
// a NaN candidate is dropped, exactly as fmax drops it.
MPC_HD void mpc_raise(double &acc, double v) {
    if (v > acc) acc = v;
}

template <class CFG, int MODE>
MPC_HD void row_body(double *st, int r, double cv, double cvn, RowCtx &c) {
    typedef SL<CFG> S;
    const double T_MIN = 1e-10, LAM_MIN = 1e-16;
    if (MODE == RP_UPDATE) {
        double tt, l;
        if (c.first) {
            if (c.warm) {  // qp_solver_warm_start = 2: keep the previous QP's multipliers, re-centre the slacks
                tt = mpc_floor_at(cv, c.ws_thr);
                l = mpc_floor_at(st[S::LAM + r], c.ws_mu * mpc_rcp(tt));  // the floor is finite: tt >= ws_thr > 0
            } else {
                tt = mpc_floor_at(cv, c.thr0);
                l = c.mu0 * mpc_rcp(tt);
            }
        } else {
            const double t0 = st[S::T + r], l0 = st[S::LAM + r], ti = st[S::TINV + r];
            const double dt = cvn - t0;
            const double dl = (c.sigma_mu - st[S::CORR + r]) * ti - l0 * ti * cvn;
            tt = mpc_floor_at(t0 + c.alpha * dt, T_MIN);
            l = mpc_floor_at(l0 + c.alpha * dl, LAM_MIN);
        }
        st[S::T + r] = tt;
        st[S::LAM + r] = l;
        st[S::TINV + r] = mpc_rcp(tt);
        st[S::CORR + r] = 0.0;
        upd_max(c.a0, fabs(cv - tt));  // inequality residual
        upd_max(c.a1, fabs(l * tt));   // complementarity residual
        c.a2 += l * tt;
    } else if (MODE == RP_AFFINE) {
        const double tt = st[S::T + r], l = st[S::LAM + r], ti = st[S::TINV + r];
        const double dt = cvn - tt;
        const double dl = -l * ti * cvn;                 // sigma mu = corr = 0
        mpc_raise(c.a0, -dt * ti);                       // ratio test: max of -dt/t and -dl/lam = cvn/t
        mpc_raise(c.a0, ti * cvn);
        c.a1 += l * dt + tt * dl;
        c.a2 += dt * dl;
        st[S::CORR + r] = dt * dl;
    } else {
        const double tt = st[S::T + r], l = st[S::LAM + r], ti = st[S::TINV + r];
        const double dt = cvn - tt;
        const double dl = (c.sigma_mu - st[S::CORR + r]) * ti - l * ti * cvn;
        mpc_raise(c.a0, -dt * ti);
        mpc_raise(c.a0, -dl * mpc_rcp(l));
    }
}

template <class CFG, int MODE>
MPC_DEV void rows_phase(const Layout &L, double *sm, RowCtx ctx) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NU = CFG::NU, NH = CFG::NH, NT = MPC_NT;
    double *W = sm + S::W_BASE;
    MPC_PHASE
    {
        RowCtx c = ctx;
        c.a0 = c.a1 = c.a2 = 0.0;
        // RP_UPDATE also TAKES the primal / equality-multiplier step of the previous iteration (it used to be a phase of its own):
        //   dz <- dz + alpha (dzn - dz),   dpi <- dpi + alpha (P dxn + pv - dpi)
        // The stepped dz goes to the shadow array DZS -- other threads of this phase still read the un-stepped DZ of the same
        // stage (general rows: x, y, psi) and step it on the fly -- and the accumulation phase that follows reads DZS and
        // copies it back.  Same expressions as the separate phase had (iterates agree to rounding: the compiler contracts them differently).
        const bool step = (MODE == RP_UPDATE) && !c.first;
        constexpr int NK = (MODE == RP_UPDATE) ? N + 1 : N;  // the update also visits the terminal node (no rows there)
        // boxes: both rows of variable i of stage k (x0 is fixed: no state box at node 0)
        for (int it = tid; it < NK * NZ; it += NT) {
            // stage fastest: neighbouring lanes sit an odd number of doubles apart (the stage stride) -- variable fastest put
            // lanes (k, i + 1) and (k + 1, i) on the same bank pair (46 % excess wavefronts in the row bodies)
            const int i = it / NK, k = it - i * NK;
            double *st = sm + k * S::STRIDE;
            double v = st[S::DZ + i];
            const double vn = st[S::DZN + i];
            if (MODE == RP_UPDATE) {
                if (step && !((k == 0 && i >= NU) || (k == N && i < NU))) v += c.alpha * (vn - v);
                st[S::DZS + i] = v;
                if (step && k >= 1 && i < CFG::NX) {
                    // Newton target of the equality multiplier, pi_k = P_k dx_k + pv_k (dx = the target's state part)
                    double pin = st[S::PV + i];
#pragma unroll
                    for (int m = 0; m < CFG::NX; m++) pin += st[S::P + trisym(i, m)] * st[S::DZN + NU + m];
                    st[S::DPI + i] += c.alpha * (pin - st[S::DPI + i]);
                }
                if (k == N) continue;
            }
            if (k == 0 && i >= NU) continue;
            const double lb = W[S::W_LB + i], ub = W[S::W_UB + i];
            const double z = st[S::Z + i];
            // rows are relaxed by qp_relax (see mpcgpu.h): an interior exists even for rows pinned at their bound.
            // One copy of the row body for both sides (run-time loop): lower z - lb + v, upper ub - z - v = -(z - ub + v)
#pragma unroll 1
            for (int side = 0; side < 2; side++) {
                const double bnd = side ? ub : lb, sgn = side ? -1.0 : 1.0;
                const bool active = side ? (ub < 1e10) : (lb > -1e10);
                if (active) row_body<CFG, MODE>(st, side * NZ + i, sgn * (z - bnd + v) + L.qp_relax, sgn * (z - bnd + vn) + L.qp_relax, c);
            }
        }
        // general rows of stages 1..N-1
        for (int it = tid; it < (N - 1) * NH; it += NT) {
            const int k = 1 + it / (NH > 0 ? NH : 1), j = it - (k - 1) * NH;
            double *st = sm + k * S::STRIDE;
            const double j0 = st[S::JR + 3 * j], j1 = st[S::JR + 3 * j + 1], j2 = st[S::JR + 3 * j + 2], rh = st[S::RH + j];
            const double nx = st[S::DZN + CFG::IX], ny = st[S::DZN + CFG::IY], np = st[S::DZN + CFG::IPSI];
            double cv = 0.0;
            if (MODE == RP_UPDATE) {  // the other modes only use the row value at the Newton target
                double dx = st[S::DZ + CFG::IX], dy = st[S::DZ + CFG::IY], dp = st[S::DZ + CFG::IPSI];
                if (step) {
                    dx += c.alpha * (nx - dx);
                    dy += c.alpha * (ny - dy);
                    dp += c.alpha * (np - dp);
                }
                cv = rh + j0 * dx + j1 * dy + j2 * dp;
            }
            double cvn = rh + j0 * nx + j1 * ny + j2 * np;
            if (CFG::ANY_SLACK_ROWS && row_has_slack<CFG>(j)) {
                const double ns = st[S::DZN + (CFG::HAS_SLACK ? CFG::ISLACK : 0)];
                if (MODE == RP_UPDATE) {
                    double ds = st[S::DZ + (CFG::HAS_SLACK ? CFG::ISLACK : 0)];
                    if (step) ds += c.alpha * (ns - ds);
                    cv += ds;
                }
                cvn += ns;
            }
            row_body<CFG, MODE>(st, 2 * NZ + j, cv, cvn, c);
        }
        if (MODE == RP_UPDATE) {
            red_put<CFG>(W, tid, 2, 0, c.a0);
            red_put<CFG>(W, tid, 3, 0, c.a1);
            red_put<CFG>(W, tid, 4, 1, c.a2);
        } else {
            red_put<CFG>(W, tid, 0, 0, c.a0);
            if (MODE == RP_AFFINE) {
                red_put<CFG>(W, tid, 1, 1, c.a1);
                red_put<CFG>(W, tid, 2, 1, c.a2);
            }
        }
    }
    MPC_PHASE_END
}

// ---------------------------------------------------------------------------------------------
// Interior-point QP solve.  Returns the QP status (acados numbering: 0 ok, 2 max iter,
// 3 min step, 4 NaN) -- uniform across the CTA.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__) && !defined(MPC_RICCATI_SERIAL) && !defined(MPC_RICCATI_LANES) && defined(MPC_SOLVE_MMA)
#define MPC_RICCATI_SOLVE riccati_solve_mma  // measured alternative: pairs with riccati_factor_mma storing Kt in the L_xu slot
#else
#define MPC_RICCATI_SOLVE riccati_solve
#endif

template <class CFG>
MPC_DEV int ipm_solve(const Layout &L, double *sm, int *iters_out, bool warm, double *res_out) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NH = CFG::NH, NT = MPC_NT;
    double *W = sm + S::W_BASE;
    double *mpc_prof = W + S::W_PROF;
    (void)mpc_prof;
    const double ALPHA_MIN = 1e-12;
    // number of inequality rows in the QP (uniform)
    int nc = 0;
    for (int i = 0; i < NZ; i++) {
        nc += (L.lower_bound[i] > -1e10) ? ((i < NU) ? N : N - 1) : 0;
        nc += (L.upper_bound[i] < 1e10) ? ((i < NU) ? N : N - 1) : 0;
    }
    nc += NH * (N - 1);
    const double inv_nc = (nc > 0) ? 1.0 / nc : 0.0;

    // start: dz = 0, pi = 0; slacks / multipliers are initialised in the first row phase (cold: HPIPM-style
    // t = max(r, thr0), lam = mu0 / t; warm: see there)
    MPC_PHASE
    for (int it = tid; it < (N + 1) * NZ; it += NT) {
        const int i = it / (N + 1), k = it - i * (N + 1);  // stage fastest: bank-conflict free
        double *st = sm + k * S::STRIDE;
        st[S::DZ + i] = (k == 0 && i >= NU) ? W[S::W_DX0 + i - NU] : 0.0;
        st[S::DZN + i] = st[S::DZ + i];
        if (i < NX) st[S::DPI + i] = 0.0;
    }
    MPC_PHASE_END

    MPC_TICK(TK_INIT)
    int status = 0, iter = 0;
    double alpha = 0.0, sigma_mu = 0.0;
    for (;; iter++) {
        // (PA, the primal / equality-multiplier step of the previous iteration, is taken inside the row phase below)
        // ---- PA + PB: step of dz, dpi; rows: (init | step) of t, lam; 1/t; residual pieces ----
        {
            RowCtx c;
            c.first = (iter == 0);
            c.warm = warm;
            c.alpha = alpha;
            c.sigma_mu = sigma_mu;
            c.thr0 = L.qp_thr0;
            c.mu0 = L.qp_mu0;
            c.ws_thr = L.qp_ws_thr;
            c.ws_mu = L.qp_ws_mu;
            rows_phase<CFG, RP_UPDATE>(L, sm, c);
        }
        MPC_TICK(TK_PB)
        // ---- PC: stationarity / equality residuals, barrier Hessian terms, affine gradient ----
        stage_accumulate<CFG, 0>(L, sm, 0.0);
        MPC_TICK(TK_ACC0)
        const double rs = red_get<CFG>(W, 0, 0), re = red_get<CFG>(W, 1, 0), ri = red_get<CFG>(W, 2, 0), rc = red_get<CFG>(W, 3, 0);
        const double mu = red_get<CFG>(W, 4, 1) * inv_nc;
        res_out[0] = rs;
        res_out[1] = rc;
        MPC_TICK(TK_RED)
#if !defined(__CUDACC__) && defined(MPC_EMU_TRACE)
        if (getenv("EMU_DEBUG")) fprintf(stderr, "  ipm %2d: stat %.3e eq %.3e ineq %.3e comp %.3e mu %.3e\n", iter, rs, re, ri, rc, mu);
#endif
        if (!(rs == rs) || !(re == re) || !(ri == ri) || !(rc == rc)) { status = 4; break; }
        if (rs < L.qp_tol && re < L.qp_tol && ri < L.qp_tol && rc < L.qp_tol) { status = 0; break; }
        if (iter >= L.qp_iter_max) { status = 2; break; }

        if constexpr (CFG::BICYCLE) {
            riccati_factor_dense<CFG>(sm);  // dense sweeps (nu = 3, three non-identity columns of A)
        } else {
#if defined(MPC_RICCATI_LANES)
            riccati_factor_lanes<CFG>(sm);  // measured alternative: same throughput, 5 % more solo latency
#elif defined(__CUDACC__) && !defined(MPC_RICCATI_SERIAL)
            riccati_factor_mma<CFG>(sm);    // FP64 tensor-core products, one warp (+ backward sweep of the predictor rhs)
#else
            riccati_factor<CFG>(sm);        // + backward sweep of the predictor right-hand side
#endif
        }
        MPC_TICK(TK_FACTOR)
        if constexpr (CFG::BICYCLE) riccati_solve_dense<CFG>(sm, false);
        else MPC_RICCATI_SOLVE<CFG>(sm, false);
        MPC_TICK(TK_SOLVE)
        sigma_mu = 0.0;
        if (nc > 0) {
            // ---- PD: affine step: ratio test, centering sums, second-order term ----
            {
                RowCtx c;
                c.first = false;
                c.warm = false;
                c.alpha = 0.0;
                c.sigma_mu = 0.0;
                rows_phase<CFG, RP_AFFINE>(L, sm, c);
            }
            MPC_TICK(TK_PD)
            {
                const double a_aff = 1.0 / fmax(1.0, red_get<CFG>(W, 0, 0));
                // mu_aff = sum (lam + a dl)(t + a dt) / nc = mu + a S1/nc + a^2 S2/nc
                const double mu_aff = mu + a_aff * (red_get<CFG>(W, 1, 1) + a_aff * red_get<CFG>(W, 2, 1)) * inv_nc;
                const double ratio = (mu > 0.0) ? mu_aff / mu : 0.0;
                // Mehrotra centering, floored at a fraction of the complementarity tolerance: driving mu far
                // below qp_tol only makes lam/t huge and buries the stationarity residual in round-off
                sigma_mu = fmax(ratio * ratio * ratio * mu, 0.1 * L.qp_tol);
            }
            // ---- PE: corrector gradient and solve ----
            stage_accumulate<CFG, 1>(L, sm, sigma_mu);
            MPC_TICK(TK_ACC1)
            if constexpr (CFG::BICYCLE) riccati_solve_dense<CFG>(sm, true);
            else MPC_RICCATI_SOLVE<CFG>(sm, true);
            MPC_TICK(TK_SOLVE)
        }
        // ---- PF: step length to the boundary ----
        {
            RowCtx c;
            c.first = false;
            c.warm = false;
            c.alpha = 0.0;
            c.sigma_mu = sigma_mu;
            rows_phase<CFG, RP_STEP>(L, sm, c);
        }
        MPC_TICK(TK_PF)
        {   // fraction-to-the-boundary rule: alpha = min(1, 0.995 alpha_max), alpha_max = 1 / max_rows(-d/val)
            const double inv_amax = red_get<CFG>(W, 0, 0);
            alpha = (inv_amax > 0.995) ? 0.995 / inv_amax : 1.0;
        }
        if (!(sigma_mu == sigma_mu) || !(alpha == alpha)) { status = 4; break; }
        if (alpha < ALPHA_MIN) { status = 3; break; }
    }
    *iters_out = iter;
    return status;
}

// ---------------------------------------------------------------------------------------------
// The whole solve.  sm: SL<CFG>::TOTAL doubles of CTA-private shared memory.
//   xinit[NX], x0[(N+1)*NZ] (reference layout), pt = parameters transposed to [npar][pt_stride(N)]
//   xtraj[(N+1)*NX], utraj[N*NU], info; iterates (may be null): [n_sqp][N+1][NZ]
// ---------------------------------------------------------------------------------------------
template <class CFG>
MPC_DEV void sqp_solve_cta(const Layout &L, double *sm, const double *xinit, const double *x0, const double *pt,
                           double *xtraj, double *utraj, mpcgpu_info *info, double *iterates, double *state = nullptr,
                           int serial_tid = 0) {
    typedef SL<CFG> S;
    constexpr int N = CFG::N, NZ = CFG::NZ, NX = CFG::NX, NU = CFG::NU, NH = CFG::NH, NT = MPC_NT;
    static_assert(N + 1 <= 32, "one stage per lane");
    double *W = sm + S::W_BASE;
    double *mpc_prof = W + S::W_PROF;
    (void)mpc_prof;
    // loadWarmstart (acados_solver_interface.cpp:274-284): x (k = 0..N) and u (k < N) from x0;
    // fresh solver: zero multipliers.  With `state` ([pi (N+1) x NX | h multipliers N x NH], mpcgpu_set_stateful) the
    // NLP multipliers of this solver slot's previous solve enter the first exact Hessian, as they do in a reference
    // local solver whose acados memory survives `operator=` (acados_solver_interface.cpp:67-77).
    MPC_PHASE
    for (int it = tid; it < (N + 1) * S::STRIDE; it += NT) {
        const int k = it / S::STRIDE, o = it - k * S::STRIDE;
        double v = 0.0;
        if (o < NZ) v = (k == N && o < NU) ? 0.0 : x0[k * NZ + o];
        if (k == N && o >= S::P && o < S::P + S::NPP) {  // MIRROR of the empty terminal block: eps I
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= o - S::P) i++;
            if (o - S::P == tri(i, i)) v = L.reg_epsilon;
        }
        if (state) {
            if (o >= S::PI && o < S::PI + NX) v = state[k * NX + (o - S::PI)];
            if (NH > 0 && k < N && o >= S::LAM + 2 * NZ && o < S::LAM + 2 * NZ + NH) v = state[(N + 1) * NX + k * NH + (o - S::LAM - 2 * NZ)];
        }
        sm[it] = v;
    }
    MPC_PHASE_END
    MPC_PHASE
    if (tid < NZ) {
        W[S::W_LB + tid] = L.lower_bound[tid];
        W[S::W_UB + tid] = L.upper_bound[tid];
    }
    if (tid == NZ) W[S::W_STID] = (double)serial_tid;
    if (tid == NZ + 1) W[S::W_TMO] = 0.0;
    MPC_PHASE_END
    // Solver::solve()'s timeout (acados_solver_interface.cpp:91-117): stop iterating once the time since the start of the
    // solve plus the average iteration time reaches the budget.  Only thread 0 keeps the clocks; the decision travels
    // through shared memory with the barrier that ends the step phase.
    long long t_solve0 = 0, t_iter0 = 0, t_iters = 0;
    if (L.timeout_s > 0.0) {
        t_iter0 = mpc_clock_ns();
        // the budget runs from the start of the CALL (first kernel of the launch sequence) when the host layer provides
        // that timestamp: solves of a later wave of a large batch do not get a fresh budget
        t_solve0 = L.t_call ? *L.t_call : t_iter0;
    }
#if defined(MPC_PROFILE) && defined(__CUDA_ARCH__)
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; i++) mpc_prof[i] = 0.0;
        mpc_prof[31] = (double)clock64();
    }
    __syncthreads();
#endif

    int qp_status = 0, sqp_iter = 0, qp_iter_total = 0;
    double qp_res[2] = {0.0, 0.0};  // last QP: stationarity / complementarity residuals
    for (int sq = 0; sq < L.n_sqp; sq++) {
        // preparation: linearise every stage at the current iterate
        MPC_PHASE
        // (stage, part) items packed over consecutive threads: warp 0 holds the cost lanes and 12 of the dynamics lanes, whose
        // divergent paths run one after the other (cost 3.4k + dynamics 2.2k = 5.7k cycles per linearisation, cfg 2; the
        // constraint quarters hide behind them).  Measured alternatives that keep cost and dynamics in different warps -- one
        // part per warp, or the slot order cost / constraints / dynamics / constraints -- both LOSE: the two long paths running
        // concurrently take 9-10k cycles (6.13 -> 6.28 .. 6.42 ms on the bench batch); the phase is bound by something the two
        // streams share (instruction fetch of ~20 KB of once-per-linearisation code is the suspect), not by the lanes.
        for (int it = tid; it < 6 * N; it += NT) {
            const int part = it / N, k = it - part * N;  // same part in neighbouring threads
            ParamT<pt_stride(N)> pa{pt, k};
            linearize_part<CFG>(pa, L, sm + k * S::STRIDE, sm + (k + 1) * S::STRIDE, k, part);
        }
        if (tid >= NT - NX) {
            const int i = tid - (NT - NX);
            W[S::W_DX0 + i] = xinit[i] - sm[S::Z + NU + i];
        }
        MPC_PHASE_END
        MPC_TICK(TK_LIN)
        MPC_PHASE
        {   // stages spread over the warps (lanes 0..ceil(N/4)-1 of each)
            const int k = (tid & 31) * (NT / 32) + (tid >> 5);
            if ((tid & 31) < (N + NT / 32 - 1) / (NT / 32) && k < N) assemble_stage<CFG>(L, sm + k * S::STRIDE);
        }
        MPC_PHASE_END
        int need_general = 0;
        MPC_PHASE
        {   // MIRROR: (stage, part) items spread over the warps: the Jacobi sweeps diverge per item
            constexpr int NI = 2 * N, NW = NT / 32;
            const int it = (tid & 31) * NW + (tid >> 5);
            if ((tid & 31) < (NI + NW - 1) / NW && it < NI) {
                mirror_stage<CFG>(L, sm + (it >> 1) * S::STRIDE, it & 1);
                // only the thread that may have written the flag reads it back (the part-1 thread of the same stage would
                // race with that write: compute-sanitizer racecheck, profiles/r02_sanitizer.md)
                if ((it & 1) == 0 && sm[(it >> 1) * S::STRIDE + S::DZN] == 2.0) need_general |= 1;
            }
        }
        MPC_PHASE_END_OR(need_general)
        if (need_general) mirror_general_cta<CFG>(L, sm);
        MPC_TICK(TK_ASM)
        // feedback: QP
        int iters = 0;
        qp_status = ipm_solve<CFG>(L, sm, &iters, sq > 0 && L.qp_warm_start != 0, qp_res);
        qp_iter_total += iters;
        sqp_iter = sq + 1;
        if (qp_status == 3 || qp_status == 4) break;  // QP failure: no step (acados sqp_rti)
        // full step; multipliers <- QP multipliers (LAM already holds them)
        MPC_PHASE
        for (int it = tid; it < (N + 1) * NZ; it += NT) {
            const int i = it / (N + 1), k = it - i * (N + 1);  // stage fastest: bank-conflict free
            double *st = sm + k * S::STRIDE;
            if (!(k == N && i < NU)) st[S::Z + i] += st[S::DZ + i];
            if (iterates) {
                const double zi = st[S::Z + i];
                if (k < N) iterates[(sq * (N + 1) + k) * NZ + i] = zi;
                else iterates[(sq * (N + 1) + k) * NZ + ((i >= NU) ? i - NU : NX + i)] = (i >= NU) ? zi : 0.0;
            }
        }
        if (tid == 0 && L.timeout_s > 0.0) {
            const long long now = mpc_clock_ns();
            t_iters += now - t_iter0;
            t_iter0 = now;
            const double avg = (double)t_iters / (double)(sq + 1), elapsed = (double)(now - t_solve0);
            W[S::W_TMO] = (elapsed + avg >= L.timeout_s * 1e9) ? 1.0 : 0.0;
        }
        MPC_PHASE_END
        MPC_TICK(TK_STEP)
        if (qp_status != 0) break;  // Q4: the reference loop stops on any non-zero qp_status
        if (W[S::W_TMO] != 0.0) break;  // out of time (:112-116)
    }

    // A QP failure leaves the loop straight after the residual check of ipm_solve: every thread has read the reduction
    // slots there, and the phase below writes them again -- one barrier in between (racecheck, profiles/r02_sanitizer.md)
#if defined(__CUDACC__)
    __syncthreads();
#endif
    // completeOneIteration (acados_solver_interface.cpp:162-204): cost, trajectories, res_eq
    MPC_PHASE
    {
        double c = 0.0, defect = 0.0;
        if (tid < N) {
            ParamT<pt_stride(N)> pa{pt, tid};
            const double *st = sm + tid * S::STRIDE, *stn = sm + (tid + 1) * S::STRIDE;
            double z[NZ], xn[NX];
#pragma unroll
            for (int i = 0; i < NZ; i++) z[i] = st[S::Z + i];
            c = L.dt * stage_cost<CFG>(pa, L, z, 0.0, nullptr, nullptr);
            model_step<CFG>(L, z, xn, nullptr, nullptr, nullptr);
#pragma unroll
            for (int i = 0; i < NX; i++) upd_max(defect, fabs(xn[i] - stn[S::Z + NU + i]));
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < NX; i++) upd_max(defect, fabs(z[NU + i] - xinit[i]));
            }
        }
        for (int it = tid; it < (N + 1) * NX; it += NT) xtraj[it] = sm[(it / NX) * S::STRIDE + S::Z + NU + it % NX];
        for (int it = tid; it < N * NU; it += NT) utraj[it] = sm[(it / NU) * S::STRIDE + S::Z + it % NU];
        red_put<CFG>(W, tid, 0, 1, c);
        red_put<CFG>(W, tid, 1, 0, defect);
    }
    MPC_PHASE_END
    MPC_PHASE
    {   // every thread combines the partials (no second barrier) and derives the exit code
        const double pobj = red_get<CFG>(W, 0, 1), res_eq = red_get<CFG>(W, 1, 0);
        int exit_code = 1;
        if (qp_status == 3 || qp_status == 4) exit_code = 4;  // ACADOS_QP_FAILURE
        if (!(res_eq <= L.res_eq_fail) && exit_code == 1) exit_code = 4;  // acados_solver_interface.cpp:176-181
        if (!(pobj == pobj) || isinf(pobj)) exit_code = 4;               // a non-finite objective is never a success
        if (tid == 0) {
            info->pobj = mpc_real(pobj);
            info->res_eq = mpc_real(res_eq);
            info->res_stat_qp = mpc_real(qp_res[0]);
            info->res_comp_qp = mpc_real(qp_res[1]);
            info->exit_code = exit_code;
            info->qp_status = qp_status;
            info->sqp_iter = sqp_iter;
            info->qp_iter = qp_iter_total;
        }
        if (state) {
            // Persist the NLP multipliers of this solver slot.  ANY non-success exit clears them: the reference calls
            // Solver_acados_reset(capsule, 1) + ocp_nlp_solver_reset_qp_memory whenever the exit code is not
            // ACADOS_SUCCESS -- including the res_eq > 1e-2 demotion (acados_solver_interface.cpp:176-191).
            // Non-finite values are never persisted.
            const bool keep = (exit_code == 1);
            for (int it = tid; it < (N + 1) * NX; it += NT) {
                const double v = sm[(it / NX) * S::STRIDE + S::PI + it % NX];
                state[it] = (keep && isfinite(v)) ? v : 0.0;
            }
            if (NH > 0)
                for (int it = tid; it < N * NH; it += NT) {
                    const double v = sm[(it / NH) * S::STRIDE + S::LAM + 2 * NZ + it % NH];
                    state[(N + 1) * NX + it] = (keep && isfinite(v)) ? v : 0.0;
                }
        }
    }
    MPC_PHASE_END
#if defined(MPC_PROFILE) && defined(__CUDA_ARCH__)
    MPC_TICK(TK_FINAL)
    if (threadIdx.x == 0)
        for (int i = 0; i < 16; i++) utraj[i] = mpc_prof[i];
#endif
}

}  // namespace mpcgpu
