// csrc/ccb_are.cu -- K5: batched algebraic Riccati equations (reference src/util/util.jl:1-59: dare, care, lqr, dlqr).
//
// The reference forms the (generalised) Schur decomposition of the 2n x 2n symplectic pencil / Hamiltonian matrix and reads the
// stabilising solution off the ordered basis (util.jl:16-19, :28-36).  An ordered QZ is a poor fit for thousands of small
// independent problems on a GPU (data-dependent bulge chasing, one eigenvalue swap at a time), so the device computes the SAME
// matrix -- the unique symmetric stabilising solution -- with the structure-preserving doubling algorithm:
//     DARE  X = A'X (I + G X)^-1 A + H,   G = B R^-1 B',  H = Q:
//         W = I + G_k H_k ;  A_{k+1} = A_k W^-1 A_k ;  G_{k+1} = G_k + A_k W^-1 G_k A_k' ;  H_{k+1} = H_k + A_k' H_k W^-1 A_k ;  H_k -> X
//     (quadratic convergence: 5-20 iterations of n x n products and one pivoted LU each), and
//     CARE  A'X + X A - X G X + H = 0  through the Cayley transform with shift gamma (A_g = A - gamma I, V = A_g + G A_g^-T H):
//         A_0 = I + 2 gamma V^-1 ,  G_0 = 2 gamma V^-1 (A_g^-1 G)' ,  H_0 = 2 gamma V^-T (A_g^-T H)' ,  then the same doubling.
// One CTA per problem; the 8 n x n work matrices live in shared memory when they fit (n <= 56), else in a global workspace.
// Optionally the gain is formed in the same kernel: dlqr  K = (R + B'X B)^-1 B'X A (util.jl:45-49), lqr  K = R^-1 B'X (util.jl:39-43).
#include "ccb_cta_linalg.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

using namespace cta;

__device__ __forceinline__ double cta_sum(double v, double* red) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < RT / 32; i++) t += red[i];      // the same order in every thread
    return t;
}

// kind 0: DARE, 1: CARE.  A [S][n][n], B [S][n][m], Q [S][n][n], R [S][m][m] row-major.  X [S][n][n] and K [S][m][n] may be null.
__global__ void __launch_bounds__(RT) are_sda_kernel(int kind, int n, int m, const double* __restrict__ Ain, const double* __restrict__ Bin,
                                                     const double* __restrict__ Qin, const double* __restrict__ Rin, double tol, int maxit,
                                                     double* __restrict__ Xout, double* __restrict__ Kout, int* __restrict__ iters,
                                                     int* __restrict__ status, double* __restrict__ ws, int use_smem) {
    __shared__ int piv;
    __shared__ double red[32];
    extern __shared__ double are_sm[];
    const size_t p = blockIdx.x, nn = (size_t)n * n;
    const int tid = threadIdx.x;
    const double* A0 = Ain + p * nn;
    const double* B0 = Bin + p * (size_t)n * m;
    const double* Q0 = Qin + p * nn;
    const double* R0 = Rin + p * (size_t)m * m;
    double* base = use_smem ? are_sm : ws + p * 8 * nn;
    double *Ak = base, *Gk = base + nn, *Hk = base + 2 * nn, *W = base + 3 * nn, *X12 = base + 4 * nn /* n x 2n */, *T1 = base + 6 * nn, *T2 = base + 7 * nn;

    // ---- G = B R^-1 B'  (util.jl:6, :23)
    for (int e = tid; e < m * m; e += RT) T2[e] = R0[e];
    for (int e = tid; e < m * n; e += RT) { const int i = e / n, j = e % n; T1[e] = B0[(size_t)j * m + i]; }   // B'
    __syncthreads();
    cta_lu_solve(m, T2, m, T1, n, n, &piv, red);                      // T1 = R^-1 B'
    cta_gemm<false>(n, n, m, B0, m, T1, n, Gk, n, nullptr, 0, 1.0);
    for (size_t e = tid; e < nn; e += RT) { Ak[e] = A0[e]; Hk[e] = Q0[e]; }
    __syncthreads();

    if (kind == 1) {
        // ---- Cayley transform of the Hamiltonian pencil
        double rs = 0.0;
        for (int i = tid; i < n; i += RT) { double s = 0.0; for (int j = 0; j < n; j++) s += fabs(A0[(size_t)i * n + j]); rs = fmax(rs, s); }
        // max over the CTA through the sum helper's buffer
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rs = fmax(rs, __shfl_xor_sync(0xffffffffu, rs, o));
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = rs;
        __syncthreads();
        double gamma = 1.0;
        for (int i = 0; i < RT / 32; i++) gamma = fmax(gamma, red[i]);
        __syncthreads();
        double* Ag = X12;            // A - gamma I            (kept)
        double* Z = X12 + nn;        // A_g^-T H               (kept)
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); Ag[e] = A0[e] - (i == j ? gamma : 0.0); }
        __syncthreads();
        // T1 = A_g^-1 G
        for (size_t e = tid; e < nn; e += RT) { W[e] = Ag[e]; T1[e] = Gk[e]; }
        __syncthreads();
        cta_lu_solve(n, W, n, T1, n, n, &piv, red);
        // Z = A_g^-T H
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); W[e] = Ag[(size_t)j * n + i]; Z[e] = Hk[e]; }
        __syncthreads();
        cta_lu_solve(n, W, n, Z, n, n, &piv, red);
        // V = A_g + G Z  -> W
        cta_gemm<false>(n, n, n, Gk, n, Z, n, W, n, Ag, n, 1.0);
        __syncthreads();
        for (size_t e = tid; e < nn; e += RT) T2[e] = W[e];           // keep V for the transposed solve
        __syncthreads();
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); Ak[e] = (i == j) ? 1.0 : 0.0; }
        __syncthreads();
        cta_lu_solve(n, W, n, Ak, n, n, &piv, red);                   // Ak = V^-1
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); W[e] = T1[(size_t)j * n + i]; }   // (A_g^-1 G)'
        __syncthreads();
        cta_gemm<false>(n, n, n, Ak, n, W, n, Gk, n, nullptr, 0, 2.0 * gamma);      // G_0 = 2 gamma V^-1 (A_g^-1 G)'
        __syncthreads();
        // H_0 = 2 gamma V^-T Z'
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); W[e] = T2[(size_t)j * n + i]; Hk[e] = Z[(size_t)j * n + i]; }
        __syncthreads();
        cta_lu_solve(n, W, n, Hk, n, n, &piv, red);
        for (size_t e = tid; e < nn; e += RT) {
            const int i = (int)(e / n), j = (int)(e % n);
            Hk[e] *= 2.0 * gamma;
            Ak[e] = fma(2.0 * gamma, Ak[e], (i == j) ? 1.0 : 0.0);     // A_0 = I + 2 gamma V^-1
        }
        __syncthreads();
    }

    // ---- doubling
    int it = 0, st = 1;                                                 // status 1: not converged within maxit
    for (; it < maxit;) {
        it++;
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); W[e] = (i == j) ? 1.0 : 0.0; }
        __syncthreads();
        cta_gemm<false>(n, n, n, Gk, n, Hk, n, W, n, W, n, 1.0);       // W = I + G H
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); X12[(size_t)i * 2 * n + j] = Ak[e]; X12[(size_t)i * 2 * n + n + j] = Gk[e]; }
        __syncthreads();
        cta_lu_solve(n, W, n, X12, 2 * n, 2 * n, &piv, red);           // X1 = W^-1 A | X2 = W^-1 G
        cta_gemm<false>(n, n, n, Hk, n, X12, 2 * n, T1, n, nullptr, 0, 1.0);         // T1 = H X1
        __syncthreads();
        cta_gemm<true>(n, n, n, Ak, n, T1, n, T2, n, nullptr, 0, 1.0);                // T2 = A' H X1
        __syncthreads();
        double dn = 0.0, hn = 0.0;
        for (size_t e = tid; e < nn; e += RT) { const double d = T2[e], h = Hk[e] + d; Hk[e] = h; dn = fma(d, d, dn); hn = fma(h, h, hn); }
        dn = cta_sum(dn, red);
        hn = cta_sum(hn, red);
        cta_gemm<false>(n, n, n, Ak, n, X12 + n, 2 * n, T1, n, nullptr, 0, 1.0);     // T1 = A X2
        for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); T2[e] = Ak[(size_t)j * n + i]; }     // A'
        __syncthreads();
        cta_gemm<false>(n, n, n, T1, n, T2, n, Gk, n, Gk, n, 1.0);                    // G += A X2 A'
        __syncthreads();
        cta_gemm<false>(n, n, n, Ak, n, X12, 2 * n, T1, n, nullptr, 0, 1.0);          // T1 = A X1
        __syncthreads();
        for (size_t e = tid; e < nn; e += RT) Ak[e] = T1[e];
        __syncthreads();
        if (!(dn == dn) || !(hn == hn) || isinf(hn)) { st = 2; break; }                // breakdown (singular W: no stabilising solution)
        if (dn <= tol * tol * hn) { st = 0; break; }
    }
    // ---- X = (H + H') / 2
    for (size_t e = tid; e < nn; e += RT) { const int i = (int)(e / n), j = (int)(e % n); W[e] = 0.5 * (Hk[e] + Hk[(size_t)j * n + i]); }
    __syncthreads();
    if (Xout)
        for (size_t e = tid; e < nn; e += RT) Xout[p * nn + e] = W[e];
    if (Kout) {
        double* BtX = T1;                 // m x n
        double* Sm = T2;                  // m x m
        cta_gemm<true>(m, n, n, B0, m, W, n, BtX, n, nullptr, 0, 1.0);               // B'X
        __syncthreads();
        if (kind == 0) {
            cta_gemm<false>(m, m, n, BtX, n, B0, m, Sm, m, R0, m, 1.0);              // R + B'X B
            cta_gemm<false>(m, n, n, BtX, n, A0, n, X12, n, nullptr, 0, 1.0);        // B'X A
            __syncthreads();
            cta_lu_solve(m, Sm, m, X12, n, n, &piv, red);
            for (int e = tid; e < m * n; e += RT) Kout[p * (size_t)m * n + e] = X12[e];
        } else {
            for (int e = tid; e < m * m; e += RT) Sm[e] = R0[e];
            __syncthreads();
            cta_lu_solve(m, Sm, m, BtX, n, n, &piv, red);
            for (int e = tid; e < m * n; e += RT) Kout[p * (size_t)m * n + e] = BtX[e];
        }
    }
    if (tid == 0) { if (iters) iters[p] = it; if (status) status[p] = st; }
}

}  // namespace

size_t are_ws_doubles(int n) { return 8 * (size_t)n * n; }

int are_sda_device(int kind, int S, int n, int m, const double* A, const double* B, const double* Q, const double* R, double tol, int maxit, double* X,
                   double* K, int* iters, int* status, double* ws, cudaStream_t stream) {
    const size_t smem = are_ws_doubles(n) * sizeof(double);
    const int use_smem = smem <= 200 * 1024;
    if (use_smem) CCB_CUDA(cudaFuncSetAttribute(are_sda_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    are_sda_kernel<<<S, RT, use_smem ? smem : 0, stream>>>(kind, n, m, A, B, Q, R, tol, maxit, X, K, iters, status, ws, use_smem);
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
