// counting_double.h -- TEST INFRASTRUCTURE.  A drop-in for `double` that counts the floating-point operations the kernel
// source executes (scripts/count_flops.py: `#define double mpc_cdouble` in front of csrc/*.cuh in the host emulation).
// Convention: add / sub / mul / div / sqrt = 1 FLOP each (an FMA = 2), transcendental functions counted separately,
// comparisons / max / fabs / negation = 0.
#pragma once
#include <cmath>

struct mpc_flop_counters {
    long long add, mul, div, sqrt_, trans;
};
extern mpc_flop_counters g_flops;

struct mpc_cdouble {
    double v;
    constexpr mpc_cdouble() : v(0.0) {}
    constexpr mpc_cdouble(double x) : v(x) {}
    constexpr mpc_cdouble(int x) : v(x) {}
    constexpr mpc_cdouble(long x) : v((double)x) {}
    constexpr mpc_cdouble(long long x) : v((double)x) {}
    constexpr mpc_cdouble(unsigned x) : v(x) {}
    explicit constexpr operator double() const { return v; }
    explicit constexpr operator int() const { return (int)v; }
    explicit constexpr operator long long() const { return (long long)v; }
    explicit constexpr operator bool() const { return v != 0.0; }
    mpc_cdouble &operator+=(mpc_cdouble o) { g_flops.add++; v += o.v; return *this; }
    mpc_cdouble &operator-=(mpc_cdouble o) { g_flops.add++; v -= o.v; return *this; }
    mpc_cdouble &operator*=(mpc_cdouble o) { g_flops.mul++; v *= o.v; return *this; }
    mpc_cdouble &operator/=(mpc_cdouble o) { g_flops.div++; v /= o.v; return *this; }
    mpc_cdouble operator-() const { return mpc_cdouble(-v); }
    mpc_cdouble operator+() const { return *this; }
};
#define MPC_CD_BIN(op, ctr)                                                                                   \
    inline mpc_cdouble operator op(mpc_cdouble a, mpc_cdouble b) { g_flops.ctr++; return mpc_cdouble(a.v op b.v); } \
    inline mpc_cdouble operator op(mpc_cdouble a, double b) { g_flops.ctr++; return mpc_cdouble(a.v op b); }       \
    inline mpc_cdouble operator op(double a, mpc_cdouble b) { g_flops.ctr++; return mpc_cdouble(a op b.v); }       \
    inline mpc_cdouble operator op(mpc_cdouble a, int b) { g_flops.ctr++; return mpc_cdouble(a.v op b); }          \
    inline mpc_cdouble operator op(int a, mpc_cdouble b) { g_flops.ctr++; return mpc_cdouble(a op b.v); }
MPC_CD_BIN(+, add)
MPC_CD_BIN(-, add)
MPC_CD_BIN(*, mul)
MPC_CD_BIN(/, div)
#undef MPC_CD_BIN
#define MPC_CD_CMP(op)                                                                   \
    inline bool operator op(mpc_cdouble a, mpc_cdouble b) { return a.v op b.v; }         \
    inline bool operator op(mpc_cdouble a, double b) { return a.v op b; }                \
    inline bool operator op(double a, mpc_cdouble b) { return a op b.v; }                \
    inline bool operator op(mpc_cdouble a, int b) { return a.v op b; }                   \
    inline bool operator op(int a, mpc_cdouble b) { return a op b.v; }
MPC_CD_CMP(<)
MPC_CD_CMP(>)
MPC_CD_CMP(<=)
MPC_CD_CMP(>=)
MPC_CD_CMP(==)
MPC_CD_CMP(!=)
#undef MPC_CD_CMP
inline mpc_cdouble sqrt(mpc_cdouble a) { g_flops.sqrt_++; return mpc_cdouble(std::sqrt(a.v)); }
inline mpc_cdouble fabs(mpc_cdouble a) { return mpc_cdouble(std::fabs(a.v)); }
inline mpc_cdouble fmax(mpc_cdouble a, mpc_cdouble b) { return mpc_cdouble(std::fmax(a.v, b.v)); }
inline mpc_cdouble fmin(mpc_cdouble a, mpc_cdouble b) { return mpc_cdouble(std::fmin(a.v, b.v)); }
inline mpc_cdouble fmax(mpc_cdouble a, double b) { return mpc_cdouble(std::fmax(a.v, b)); }
inline mpc_cdouble fmax(double a, mpc_cdouble b) { return mpc_cdouble(std::fmax(a, b.v)); }
inline mpc_cdouble fmin(mpc_cdouble a, double b) { return mpc_cdouble(std::fmin(a.v, b)); }
inline mpc_cdouble fmin(double a, mpc_cdouble b) { return mpc_cdouble(std::fmin(a, b.v)); }
inline mpc_cdouble exp(mpc_cdouble a) { g_flops.trans++; return mpc_cdouble(std::exp(a.v)); }
inline mpc_cdouble log(mpc_cdouble a) { g_flops.trans++; return mpc_cdouble(std::log(a.v)); }
inline mpc_cdouble sin(mpc_cdouble a) { g_flops.trans++; return mpc_cdouble(std::sin(a.v)); }
inline mpc_cdouble cos(mpc_cdouble a) { g_flops.trans++; return mpc_cdouble(std::cos(a.v)); }
inline mpc_cdouble tan(mpc_cdouble a) { g_flops.trans++; return mpc_cdouble(std::tan(a.v)); }
inline mpc_cdouble atan(mpc_cdouble a) { g_flops.trans++; return mpc_cdouble(std::atan(a.v)); }
inline mpc_cdouble erf(mpc_cdouble a) { g_flops.trans++; return mpc_cdouble(std::erf(a.v)); }
inline mpc_cdouble fmod(mpc_cdouble a, mpc_cdouble b) { g_flops.div++; return mpc_cdouble(std::fmod(a.v, b.v)); }
inline void sincos(mpc_cdouble a, mpc_cdouble *s, mpc_cdouble *c) { g_flops.trans += 2; s->v = std::sin(a.v); c->v = std::cos(a.v); }
inline bool isfinite(mpc_cdouble a) { return std::isfinite(a.v); }
inline bool isinf(mpc_cdouble a) { return std::isinf(a.v); }
inline bool isnan(mpc_cdouble a) { return std::isnan(a.v); }
