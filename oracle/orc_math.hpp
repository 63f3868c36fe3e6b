// oracle/orc_math.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle). Not part of the shipped product path.
//
// Small templated vector / quaternion / dual-number helpers for the CPU restatement of the
// ConstrainedControl.jl hot path.  Scalar type T is `double` for the Newton solve and `Dual`
// (forward-mode AD) for `linearsystem`.
//
// Quaternion convention: scalar first, q = (s, v1, v2, v3) -- reference `lqr.jl:101-102`
// (`qerr = qd \ q`, `imag(qerr)`), SURVEY.md A.1.
#pragma once
#include <cmath>
#include <cstddef>

namespace orc {

// ---------------------------------------------------------------- forward-mode dual number
struct Dual {
    double v, d;
    Dual() : v(0), d(0) {}
    Dual(double v_) : v(v_), d(0) {}
    Dual(double v_, double d_) : v(v_), d(d_) {}
};
inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
inline Dual operator-(Dual a) { return {-a.v, -a.d}; }
inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
inline Dual operator/(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
inline Dual& operator+=(Dual& a, Dual b) { a = a + b; return a; }
inline Dual& operator-=(Dual& a, Dual b) { a = a - b; return a; }
inline Dual sqrt(Dual a) { double r = std::sqrt(a.v); return {r, a.d / (2.0 * r)}; }
inline double sqrt(double a) { return std::sqrt(a); }
inline double val(double a) { return a; }
inline double val(Dual a) { return a.v; }

// ---------------------------------------------------------------- counting scalar
// A double that counts the floating-point operations performed on it (per thread).  Instantiating the templated Newton solve with it
// (oracle.cpp: orc_count_flops) gives the EXACT operation count of the literal algorithm -- residual, analytic KKT blocks, mixed-node block
// LDU in dfs_list order, back-substitution, line search -- for one mechanism at one state: the "algorithmic flops" of SURVEY.md 8(d).
// Convention: add/sub/mul/div/sqrt/compare of two Count values = 1 flop each (an FMA-capable machine executes a mul+add pair as one
// instruction but it is 2 flops); negation and copies are free; products with an exact 0.0 or 1.0 constant still count (the literal code
// performs them).
struct FlopTally { long long add = 0, mul = 0, div = 0, sqrt_ = 0, cmp = 0; long long total() const { return add + mul + div + sqrt_; } };
inline FlopTally& flop_tally() { static thread_local FlopTally t; return t; }
struct Count {
    double v;
    Count() : v(0) {}
    Count(double v_) : v(v_) {}
};
inline Count operator+(Count a, Count b) { flop_tally().add++; return {a.v + b.v}; }
inline Count operator-(Count a, Count b) { flop_tally().add++; return {a.v - b.v}; }
inline Count operator-(Count a) { return {-a.v}; }
inline Count operator*(Count a, Count b) { flop_tally().mul++; return {a.v * b.v}; }
inline Count operator/(Count a, Count b) { flop_tally().div++; return {a.v / b.v}; }
inline Count& operator+=(Count& a, Count b) { a = a + b; return a; }
inline Count& operator-=(Count& a, Count b) { a = a - b; return a; }
inline Count& operator*=(Count& a, Count b) { a = a * b; return a; }
inline bool operator==(Count a, Count b) { return a.v == b.v; }
inline bool operator>(Count a, Count b) { flop_tally().cmp++; return a.v > b.v; }
inline bool operator<(Count a, Count b) { flop_tally().cmp++; return a.v < b.v; }
inline Count sqrt(Count a) { flop_tally().sqrt_++; return {std::sqrt(a.v)}; }
inline double val(Count a) { return a.v; }

// ---------------------------------------------------------------- 3-vectors / quaternions
template <class T> struct V3 { T x, y, z; };
template <class T> struct Q4 { T s, x, y, z; };

template <class T> inline V3<T> mk3(T x, T y, T z) { return {x, y, z}; }
template <class T> inline V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline V3<T> operator-(V3<T> a) { return {-a.x, -a.y, -a.z}; }
template <class T> inline V3<T> operator*(T s, V3<T> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> inline T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline V3<T> cross(V3<T> a, V3<T> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T> inline T comp(V3<T> a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

template <class T> inline Q4<T> qmul(Q4<T> a, Q4<T> b) {
    return {a.s * b.s - a.x * b.x - a.y * b.y - a.z * b.z,
            a.s * b.x + b.s * a.x + (a.y * b.z - a.z * b.y),
            a.s * b.y + b.s * a.y + (a.z * b.x - a.x * b.z),
            a.s * b.z + b.s * a.z + (a.x * b.y - a.y * b.x)};
}
template <class T> inline Q4<T> qconj(Q4<T> a) { return {a.s, -a.x, -a.y, -a.z}; }
template <class T> inline V3<T> qvec(Q4<T> a) { return {a.x, a.y, a.z}; }
// R(q) p
template <class T> inline V3<T> qrot(Q4<T> q, V3<T> p) {
    V3<T> v = qvec(q);
    V3<T> t = cross(v, p);
    t = T(2.0) * t;
    return p + q.s * t + cross(v, t);
}
// R(q)^T p
template <class T> inline V3<T> qrotT(Q4<T> q, V3<T> p) { return qrot(qconj(q), p); }

template <class T> inline V3<T> lift3(const double* p) { return {T(p[0]), T(p[1]), T(p[2])}; }
template <class T> inline Q4<T> lift4(const double* p) { return {T(p[0]), T(p[1]), T(p[2]), T(p[3])}; }

}  // namespace orc
