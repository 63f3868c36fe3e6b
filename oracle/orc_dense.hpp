// oracle/orc_dense.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle).
// Plain dense row-major helpers standing in for Julia's LinearAlgebra (`*`, `\`, `/`, `norm`) at the
// call sites src/control/lqr.jl:151-172 and src/control/lqr_tracking.jl:89-110.
#pragma once
#include <cmath>
#include <vector>
#include "orc_math.hpp"

namespace orc {

typedef std::vector<double> Mat;  // row-major, dimensions carried by the caller

// The dense helpers are templated on the scalar so that the SAME restatement of dlqr can be run in Float64 (the reference's
// arithmetic: every parity test uses this instantiation), in x87 extended precision (long double, eps 1.1e-19) and in IEEE
// binary128 (__float128, eps 1.9e-34).  The wider instantiations exist for ONE purpose: to measure how much of a gain
// difference between two Float64 implementations is rounding amplified by the recursion itself (tests/test_oracle_quad.py).
inline double t_abs(double x) { return std::fabs(x); }
inline long double t_abs(long double x) { return std::fabs(x); }
inline double t_sqrt(double x) { return std::sqrt(x); }
inline long double t_sqrt(long double x) { return std::sqrt(x); }
inline Count t_abs(Count x) { return {std::fabs(x.v)}; }
inline Count t_sqrt(Count x) { return sqrt(x); }
#ifdef ORC_WITH_QUAD
inline __float128 t_abs(__float128 x) { return x < 0 ? -x : x; }
inline __float128 t_sqrt(__float128 x) {       // Newton on the double estimate: three steps reach binary128 accuracy
    if (x <= 0) return 0;
    __float128 y = (__float128)std::sqrt((double)x);
    for (int i = 0; i < 3; i++) y = (y + x / y) / 2;
    return y;
}
#endif

// C[m x n] = A[m x k] * B[k x n]
template <class T> inline void mm(const T* A, const T* B, T* C, int m, int k, int n) {
    for (int i = 0; i < m; i++) {
        T* c = C + (size_t)i * n;
        for (int j = 0; j < n; j++) c[j] = T(0.0);
        for (int p = 0; p < k; p++) {
            T a = A[(size_t)i * k + p];
            if (a == T(0.0)) continue;
            const T* b = B + (size_t)p * n;
            for (int j = 0; j < n; j++) c[j] += a * b[j];
        }
    }
}
// C[m x n] = A^T * B with A[k x m], B[k x n]
template <class T> inline void mtm(const T* A, const T* B, T* C, int m, int k, int n) {
    for (int i = 0; i < m * n; i++) C[i] = T(0.0);
    for (int p = 0; p < k; p++)
        for (int i = 0; i < m; i++) {
            T a = A[(size_t)p * m + i];
            if (a == T(0.0)) continue;
            const T* b = B + (size_t)p * n;
            T* c = C + (size_t)i * n;
            for (int j = 0; j < n; j++) c[j] += a * b[j];
        }
}

// In-place LU with partial pivoting (LAPACK getrf semantics); returns 0 on success, k+1 on exact zero pivot.
template <class T> inline int lu_factor(T* A, int n, int* piv) {
    for (int k = 0; k < n; k++) {
        int p = k;
        T best = t_abs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) {
            T v = t_abs(A[(size_t)i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (best == T(0.0)) return k + 1;
        if (p != k)
            for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
        T inv = T(1.0) / A[(size_t)k * n + k];
        for (int i = k + 1; i < n; i++) {
            T l = A[(size_t)i * n + k] * inv;
            A[(size_t)i * n + k] = l;
            if (l == T(0.0)) continue;
            for (int j = k + 1; j < n; j++) A[(size_t)i * n + j] -= l * A[(size_t)k * n + j];
        }
    }
    return 0;
}
// Solve A X = B for X (B is n x nrhs, overwritten)
template <class T> inline void lu_solve(const T* LU, const int* piv, int n, T* B, int nrhs) {
    for (int k = 0; k < n; k++)
        if (piv[k] != k)
            for (int j = 0; j < nrhs; j++) std::swap(B[(size_t)k * nrhs + j], B[(size_t)piv[k] * nrhs + j]);
    for (int k = 0; k < n; k++)
        for (int i = k + 1; i < n; i++) {
            T l = LU[(size_t)i * n + k];
            if (l == T(0.0)) continue;
            for (int j = 0; j < nrhs; j++) B[(size_t)i * nrhs + j] -= l * B[(size_t)k * nrhs + j];
        }
    for (int k = n - 1; k >= 0; k--) {
        T inv = T(1.0) / LU[(size_t)k * n + k];
        for (int j = 0; j < nrhs; j++) B[(size_t)k * nrhs + j] *= inv;
        for (int i = 0; i < k; i++) {
            T u = LU[(size_t)i * n + k];
            if (u == T(0.0)) continue;
            for (int j = 0; j < nrhs; j++) B[(size_t)i * nrhs + j] -= u * B[(size_t)k * nrhs + j];
        }
    }
}

// small dense inverse by Gauss-Jordan with partial pivoting (n <= 16); returns false on zero pivot
template <class T> inline bool small_inv(const T* A, T* Ai, int n) {
    T W[16 * 32];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) { W[i * 2 * n + j] = A[i * n + j]; W[i * 2 * n + n + j] = (i == j) ? T(1.0) : T(0.0); }
    for (int k = 0; k < n; k++) {
        int p = k;
        T best = t_abs(W[k * 2 * n + k]);
        for (int i = k + 1; i < n; i++)
            if (t_abs(W[i * 2 * n + k]) > best) { best = t_abs(W[i * 2 * n + k]); p = i; }
        if (best == T(0.0)) return false;
        if (p != k) for (int j = 0; j < 2 * n; j++) std::swap(W[k * 2 * n + j], W[p * 2 * n + j]);
        T inv = T(1.0) / W[k * 2 * n + k];
        for (int j = 0; j < 2 * n; j++) W[k * 2 * n + j] *= inv;
        for (int i = 0; i < n; i++) {
            if (i == k) continue;
            T l = W[i * 2 * n + k];
            if (l == T(0.0)) continue;
            for (int j = 0; j < 2 * n; j++) W[i * 2 * n + j] -= l * W[k * 2 * n + j];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Ai[i * n + j] = W[i * 2 * n + n + j];
    return true;
}

}  // namespace orc
