// csrc/ccb_dev_dual.cuh -- scalar-generic (double / forward-mode dual) device helpers for the linearisation kernel K2.
// The second-order terms of linearsystem (constraint Hessians contracted with λ_d, input wrenches rotating with the
// bodies) are obtained by pushing one tangent direction through the analytic first-order rows.
#pragma once
#include <cuda_runtime.h>

namespace ccb {

struct Dual {
    double v, d;
    __device__ __forceinline__ Dual() : v(0.0), d(0.0) {}
    __device__ __forceinline__ Dual(double v_) : v(v_), d(0.0) {}
    __device__ __forceinline__ Dual(double v_, double d_) : v(v_), d(d_) {}
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, fma(a.d, b.v, a.v * b.d)}; }

template <class T> struct TV3 { T x, y, z; };
template <class T> struct TQ4 { T s, x, y, z; };

template <class T> __device__ __forceinline__ TV3<T> operator+(TV3<T> a, TV3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> __device__ __forceinline__ TV3<T> operator-(TV3<T> a, TV3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> __device__ __forceinline__ TV3<T> operator-(TV3<T> a) { return {-a.x, -a.y, -a.z}; }
template <class T> __device__ __forceinline__ TV3<T> scale(T s, TV3<T> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> __device__ __forceinline__ T tdot(TV3<T> a, TV3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> __device__ __forceinline__ TV3<T> tcross(TV3<T> a, TV3<T> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T> __device__ __forceinline__ TQ4<T> tqmul(TQ4<T> a, TQ4<T> b) {
    return {a.s * b.s - a.x * b.x - a.y * b.y - a.z * b.z, a.s * b.x + b.s * a.x + (a.y * b.z - a.z * b.y),
            a.s * b.y + b.s * a.y + (a.z * b.x - a.x * b.z), a.s * b.z + b.s * a.z + (a.x * b.y - a.y * b.x)};
}
template <class T> __device__ __forceinline__ TQ4<T> tqconj(TQ4<T> a) { return {a.s, -a.x, -a.y, -a.z}; }
template <class T> __device__ __forceinline__ TV3<T> tqvec(TQ4<T> a) { return {a.x, a.y, a.z}; }
template <class T> __device__ __forceinline__ TV3<T> tqrot(TQ4<T> q, TV3<T> p) {
    TV3<T> v = tqvec(q);
    TV3<T> t = scale(T(2.0), tcross(v, p));
    return p + scale(q.s, t) + tcross(v, t);
}
template <class T> __device__ __forceinline__ TV3<T> tqrotT(TQ4<T> q, TV3<T> p) { return tqrot(tqconj(q), p); }
template <class T> __device__ __forceinline__ TV3<T> lift(const double* p) { return {T(__ldg(p)), T(__ldg(p + 1)), T(__ldg(p + 2))}; }
template <class T> __device__ __forceinline__ TQ4<T> liftq(const double* p) { return {T(__ldg(p)), T(__ldg(p + 1)), T(__ldg(p + 2)), T(__ldg(p + 3))}; }

// tangent Jacobian rows of one scalar constraint row (same maths as row_eval in ccb_dev_math.cuh), generic scalar
template <class T>
__device__ __forceinline__ void row_jac_t(int kind, TV3<T> pr, TV3<T> p1, TV3<T> p2, TQ4<T> qoff, TV3<T> xa, TQ4<T> qa, TV3<T> xb, TQ4<T> qb,
                                          TV3<T>& Gax, TV3<T>& Gap, TV3<T>& Gbx, TV3<T>& Gbp) {
    if (kind == 0) {
        TV3<T> w = xb + tqrot(qb, p2) - xa;
        TV3<T> c = tqrotT(qa, w);
        TV3<T> n = tqrot(qa, pr);
        TV3<T> m = tqrotT(qb, n);
        Gbx = n;
        Gbp = scale(T(-2.0), tcross(m, p2));
        Gax = -n;
        Gap = scale(T(2.0), tcross(pr, c));
    } else {
        TQ4<T> qab = tqmul(tqconj(qa), qb);
        TQ4<T> qe = tqmul(qab, tqconj(qoff));
        TQ4<T> e = {T(0.0), pr.x, pr.y, pr.z};
        Gbx = {T(0.0), T(0.0), T(0.0)};
        Gax = Gbx;
        Gbp = tqvec(tqmul(tqmul(tqconj(qab), e), qoff));
        Gap = -tqvec(tqmul(e, tqconj(qe)));
    }
}

// body wrench of a unit-free joint input u (applyFτ!): returns (F world, tau body) for `side` (0 parent, 1 child)
template <class T>
__device__ __forceinline__ void input_wrench_t(int act_kind, T u, TV3<T> axis, TV3<T> p1, TV3<T> p2, TQ4<T> qa, TQ4<T> qb, int side, TV3<T>& F, TV3<T>& tau) {
    TV3<T> ax = scale(u, axis);
    TV3<T> zero = {T(0.0), T(0.0), T(0.0)};
    if (act_kind == 1) {
        TV3<T> Fb = tqrot(qa, ax);
        if (side == 0) { F = -Fb; tau = -tcross(p1, ax); }
        else { F = Fb; tau = tcross(p2, tqrotT(qb, Fb)); }
    } else {
        F = zero;
        if (side == 0) tau = -ax;
        else tau = tqrotT(qb, tqrot(qa, ax));
    }
}

}  // namespace ccb
