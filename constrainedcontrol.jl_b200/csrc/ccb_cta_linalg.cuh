// csrc/ccb_cta_linalg.cuh -- CTA-cooperative dense Float64 helpers shared by the Riccati kernels (K3, projection) and the ARE solver:
// register-tiled FMA GEMM and LU with partial pivoting on row-major matrices in global or shared memory.  All threads of a
// CTA of RT threads must call them together.
#pragma once
#include "ccb_dev_math.cuh"

namespace ccb {
namespace cta {

constexpr int RT = 256;   // threads per CTA

// C[M x N] = beta_src + sign * op(A) B   with op(A) = A (K x ... row-major [M][K]) or A^T (A stored [K][M]); all row-major.
// `src` (may be null) is added: C = src + sign * op(A) B.  Every thread of the CTA takes TM x TN register tiles.
template <int TM, int TN, bool TA>
__device__ __forceinline__ void cta_gemm_t(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                           double* __restrict__ C, int ldc, const double* __restrict__ src, int lds, double sign) {
    const int tm = (M + TM - 1) / TM, tn = (N + TN - 1) / TN;
    for (int t = threadIdx.x; t < tm * tn; t += RT) {
        const int i0 = (t / tn) * TM, j0 = (t % tn) * TN;
        double acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < TN; j++) acc[i][j] = 0.0;
        for (int k = 0; k < K; k++) {
            double a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = (i0 + i < M) ? (TA ? A[(size_t)k * lda + i0 + i] : A[(size_t)(i0 + i) * lda + k]) : 0.0;
            // the left operands of the projection (G, Bλ) are block sparse with exact zeros (a row of G touches 2 of the nb bodies):
            // skipping a zero multiplicand skips the loads of B as well and adds exactly nothing
            bool any = false;
#pragma unroll
            for (int i = 0; i < TM; i++) any |= (a[i] != 0.0);
            if (!any) continue;
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = (j0 + j < N) ? B[(size_t)k * ldb + j0 + j] : 0.0;
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < TN; j++)
                if (i0 + i < M && j0 + j < N) {
                    const double s = src ? src[(size_t)(i0 + i) * lds + j0 + j] : 0.0;
                    C[(size_t)(i0 + i) * ldc + j0 + j] = fma(sign, acc[i][j], s);
                }
    }
}
template <bool TA>
__device__ __forceinline__ void cta_gemm(int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                                         const double* src, int lds, double sign) {
    if ((long long)M * N >= 16LL * RT) cta_gemm_t<4, 4, TA>(M, N, K, A, lda, B, ldb, C, ldc, src, lds, sign);
    else if ((long long)M * N >= 4LL * RT) cta_gemm_t<2, 2, TA>(M, N, K, A, lda, B, ldb, C, ldc, src, lds, sign);
    else cta_gemm_t<1, 1, TA>(M, N, K, A, lda, B, ldb, C, ldc, src, lds, sign);
}

// in-place LU with partial pivoting of an m x m row-major matrix in global/shared memory + the same row swaps and
// eliminations applied to an m x nr right-hand side block; then back substitution.  All threads cooperate.
__device__ inline void cta_lu_solve(int m, double* Mx, int ldm, double* Bx, int ldb, int nr, int* piv_sh, double* red_sh) {
    const int tid = threadIdx.x;
    for (int k = 0; k < m; k++) {
        if (tid < 32) {   // pivot search by one warp
            double best = -1.0;
            int bi = k;
            for (int i = k + tid; i < m; i += 32) { double v = fabs(Mx[(size_t)i * ldm + k]); if (v > best) { best = v; bi = i; } }
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) *piv_sh = bi;
        }
        __syncthreads();
        const int p = *piv_sh;
        if (p != k) {
            for (int j = tid; j < m; j += RT) { double t = Mx[(size_t)k * ldm + j]; Mx[(size_t)k * ldm + j] = Mx[(size_t)p * ldm + j]; Mx[(size_t)p * ldm + j] = t; }
            for (int j = tid; j < nr; j += RT) { double t = Bx[(size_t)k * ldb + j]; Bx[(size_t)k * ldb + j] = Bx[(size_t)p * ldb + j]; Bx[(size_t)p * ldb + j] = t; }
        }
        __syncthreads();
        const double inv = 1.0 / Mx[(size_t)k * ldm + k];
        // elimination: a thread owns one column of [M | B] (no integer divisions in the loop), RT / 128 row groups
        const int wid = (m - k - 1) + nr;
        const int jj0 = tid & 127, ig = tid >> 7, ng = RT >> 7;
        for (int jj = jj0; jj < wid; jj += 128) {
            double* col;
            size_t ld;
            if (jj < m - k - 1) { col = Mx + k + 1 + jj; ld = ldm; }
            else { col = Bx + (jj - (m - k - 1)); ld = ldb; }
            const double pk = col[(size_t)k * ld];
            for (int i = k + 1 + ig; i < m; i += ng) {
                const double l = Mx[(size_t)i * ldm + k] * inv;
                col[(size_t)i * ld] = fma(-l, pk, col[(size_t)i * ld]);
            }
        }
        __syncthreads();
    }
    // back substitution: column-parallel
    for (int j = tid; j < nr; j += RT) {
        for (int i = m - 1; i >= 0; i--) {
            double x = Bx[(size_t)i * ldb + j];
            for (int c = i + 1; c < m; c++) x = fma(-Mx[(size_t)i * ldm + c], Bx[(size_t)c * ldb + j], x);
            Bx[(size_t)i * ldb + j] = x / Mx[(size_t)i * ldm + i];
        }
    }
    __syncthreads();
}

}  // namespace cta
}  // namespace ccb
