// oracle/orc_dense.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle).
// Plain dense row-major helpers standing in for Julia's LinearAlgebra (`*`, `\`, `/`, `norm`) at the
// call sites src/control/lqr.jl:151-172 and src/control/lqr_tracking.jl:89-110.
#pragma once
#include <cmath>
#include <vector>

namespace orc {

typedef std::vector<double> Mat;  // row-major, dimensions carried by the caller

// C[m x n] = A[m x k] * B[k x n]
inline void mm(const double* A, const double* B, double* C, int m, int k, int n) {
    for (int i = 0; i < m; i++) {
        double* c = C + (size_t)i * n;
        for (int j = 0; j < n; j++) c[j] = 0.0;
        for (int p = 0; p < k; p++) {
            double a = A[(size_t)i * k + p];
            if (a == 0.0) continue;
            const double* b = B + (size_t)p * n;
            for (int j = 0; j < n; j++) c[j] += a * b[j];
        }
    }
}
// C[m x n] = A^T * B with A[k x m], B[k x n]
inline void mtm(const double* A, const double* B, double* C, int m, int k, int n) {
    for (int i = 0; i < m * n; i++) C[i] = 0.0;
    for (int p = 0; p < k; p++)
        for (int i = 0; i < m; i++) {
            double a = A[(size_t)p * m + i];
            if (a == 0.0) continue;
            const double* b = B + (size_t)p * n;
            double* c = C + (size_t)i * n;
            for (int j = 0; j < n; j++) c[j] += a * b[j];
        }
}

// In-place LU with partial pivoting (LAPACK getrf semantics); returns 0 on success, k+1 on exact zero pivot.
inline int lu_factor(double* A, int n, int* piv) {
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = std::fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) {
            double v = std::fabs(A[(size_t)i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (best == 0.0) return k + 1;
        if (p != k)
            for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
        double inv = 1.0 / A[(size_t)k * n + k];
        for (int i = k + 1; i < n; i++) {
            double l = A[(size_t)i * n + k] * inv;
            A[(size_t)i * n + k] = l;
            if (l == 0.0) continue;
            for (int j = k + 1; j < n; j++) A[(size_t)i * n + j] -= l * A[(size_t)k * n + j];
        }
    }
    return 0;
}
// Solve A X = B for X (B is n x nrhs, overwritten)
inline void lu_solve(const double* LU, const int* piv, int n, double* B, int nrhs) {
    for (int k = 0; k < n; k++)
        if (piv[k] != k)
            for (int j = 0; j < nrhs; j++) std::swap(B[(size_t)k * nrhs + j], B[(size_t)piv[k] * nrhs + j]);
    for (int k = 0; k < n; k++)
        for (int i = k + 1; i < n; i++) {
            double l = LU[(size_t)i * n + k];
            if (l == 0.0) continue;
            for (int j = 0; j < nrhs; j++) B[(size_t)i * nrhs + j] -= l * B[(size_t)k * nrhs + j];
        }
    for (int k = n - 1; k >= 0; k--) {
        double inv = 1.0 / LU[(size_t)k * n + k];
        for (int j = 0; j < nrhs; j++) B[(size_t)k * nrhs + j] *= inv;
        for (int i = 0; i < k; i++) {
            double u = LU[(size_t)i * n + k];
            if (u == 0.0) continue;
            for (int j = 0; j < nrhs; j++) B[(size_t)i * nrhs + j] -= u * B[(size_t)k * nrhs + j];
        }
    }
}

// small dense inverse by Gauss-Jordan with partial pivoting (n <= 16); returns false on zero pivot
inline bool small_inv(const double* A, double* Ai, int n) {
    double W[16 * 32];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) { W[i * 2 * n + j] = A[i * n + j]; W[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0; }
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = std::fabs(W[k * 2 * n + k]);
        for (int i = k + 1; i < n; i++)
            if (std::fabs(W[i * 2 * n + k]) > best) { best = std::fabs(W[i * 2 * n + k]); p = i; }
        if (best == 0.0) return false;
        if (p != k) for (int j = 0; j < 2 * n; j++) std::swap(W[k * 2 * n + j], W[p * 2 * n + j]);
        double inv = 1.0 / W[k * 2 * n + k];
        for (int j = 0; j < 2 * n; j++) W[k * 2 * n + j] *= inv;
        for (int i = 0; i < n; i++) {
            if (i == k) continue;
            double l = W[i * 2 * n + k];
            if (l == 0.0) continue;
            for (int j = 0; j < 2 * n; j++) W[i * 2 * n + j] -= l * W[k * 2 * n + j];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Ai[i * n + j] = W[i * 2 * n + n + j];
    return true;
}

}  // namespace orc
