// csrc/ccb_riccati.cu -- K3: batched constrained finite-horizon Riccati recursion
// (reference src/control/lqr.jl:141-184 `dlqr(A,Bu,Bλ,G,Q,R,N)` and the identical loop body of
// src/control/lqr_tracking.jl:87-115).
//
// The reference solves, per step k,  M K = b  with  M = [R + D'P Bu, D'P Bλ; G Bu, G Bλ],  b = [D'P; G] A,
// D = Bu - Bλ (G Bλ)^-1 G Bu  (lqr.jl:151-160).  Eliminating the (loop-invariant) second block row gives the same gains
// from a standard Riccati step on the constraint-projected pair
//        Ã = A - Bλ (G Bλ)^-1 G A ,      D   (so that  G Ã = 0, G D = 0):
//        Ku = (R + D'P D)^-1 D'P Ã ,   Abar = A - Bu Ku - Bλ Kλ = Ã - D Ku ,   P+ = Q + Ku'R Ku + Abar'P Abar   (lqr.jl:169-170)
// so the device does  (1) `project`: Ã, D once per linearisation (LU with partial pivoting of G Bλ), and
// (2) `sweep`: k = N-1 .. 1 with the early break ‖P - P+‖_F < 1e-5 (lqr.jl:172) and the back-fill (lqr.jl:179-181).
// One CTA per system; matrices live in a global workspace (L1/L2 resident for the small mechanisms).
#include <cstdlib>
#include "ccb_cta_linalg.cuh"
#include "ccb_dev_math.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

using namespace cta;

// (1) projection: Atil = A - Bl Y_A, D = Bu - Bl Y_u with (G Bl) [Y_A | Y_u] = [G A | G Bu]
// workspace per problem: GB [ml][ml], Y [ml][n+nu]
__global__ void __launch_bounds__(RT) riccati_project_kernel(int n, int nu, int ml, const double* __restrict__ A, const double* __restrict__ Bu,
                                                             const double* __restrict__ Bl, const double* __restrict__ G, double* __restrict__ At,
                                                             double* __restrict__ D, double* __restrict__ ws, int use_smem) {
    __shared__ int piv;
    __shared__ double red[32];
    const size_t p = blockIdx.x;
    A += p * n * n; Bu += p * n * nu; Bl += p * n * ml; G += p * ml * n; At += p * n * n; D += p * n * nu;
    if (ml == 0) {
        for (int t = threadIdx.x; t < n * n; t += RT) At[t] = A[t];
        for (int t = threadIdx.x; t < n * nu; t += RT) D[t] = Bu[t];
        return;
    }
    // G Bλ and its right-hand sides live in shared memory when they fit (the pivoted LU is a chain of ml barriers: at global-memory
    // latency it dominated the whole projection), else in the global workspace (32-pendulum: ml = 165)
    extern __shared__ double proj_sm[];
    double* GB = use_smem ? proj_sm : ws + p * ((size_t)ml * ml + (size_t)ml * (n + nu));
    double* Y = GB + (size_t)ml * ml;
    cta_gemm<false>(ml, ml, n, G, n, Bl, ml, GB, ml, nullptr, 0, 1.0);
    cta_gemm<false>(ml, n, n, G, n, A, n, Y, n + nu, nullptr, 0, 1.0);
    cta_gemm<false>(ml, nu, n, G, n, Bu, nu, Y + n, n + nu, nullptr, 0, 1.0);
    __syncthreads();
    cta_lu_solve(ml, GB, ml, Y, n + nu, n + nu, &piv, red);
    cta_gemm<false>(n, n, ml, Bl, ml, Y, n + nu, At, n, A, n, -1.0);
    cta_gemm<false>(n, nu, ml, Bl, ml, Y + n, n + nu, D, nu, Bu, nu, -1.0);
}

// (2) the sweep.  Per-system workspace: P, P1, Abar, W  (4 n^2) + DtP, rhs (2 nu n) + Sc (nu^2)
// At/D: [S][T][..] with T = 1 (time invariant, stride_t = 0) or N-1 (tracking; entry k-1 is the linearisation at step k)
__global__ void __launch_bounds__(RT) riccati_sweep_kernel(int n, int nu, int N, const double* __restrict__ At, const double* __restrict__ D,
                                                           long long stride_t_A, long long stride_t_D, long long stride_s_A, long long stride_s_D,
                                                           const double* __restrict__ Q, long long Q_bs, const double* __restrict__ R, long long R_bs,
                                                           double* __restrict__ Ku, int* __restrict__ kexit, int* __restrict__ iters,
                                                           double* __restrict__ ws) {
    __shared__ double red[RT / 32];
    __shared__ double Sc[8 * 8], Rs[8 * 8];
    __shared__ int brk;
    const size_t s = blockIdx.x;
    const int tid = threadIdx.x;
    const size_t nn = (size_t)n * n;
    double* P = ws + s * (4 * nn + 2 * (size_t)nu * n);
    double* P1 = P + nn;
    double* Abar = P1 + nn;
    double* W = Abar + nn;
    double* DtP = W + nn;
    double* Kk = DtP + (size_t)nu * n;
    const double* Qs = Q + s * Q_bs;
    const double* Rg = R + s * R_bs;
    double* Kout = Ku + s * (size_t)(N - 1) * nu * n;
    for (size_t t = tid; t < nn; t += RT) P[t] = Qs[t];            // lqr.jl:147  Pk = Q
    for (int t = tid; t < nu * nu; t += RT) Rs[t] = Rg[t];
    __syncthreads();
    int k, it = 0;
    for (k = N - 1; k >= 1; k--) {                                  // lqr.jl:150
        it++;
        const double* Ak = At + s * stride_s_A + (size_t)(k - 1) * stride_t_A;
        const double* Dk = D + s * stride_s_D + (size_t)(k - 1) * stride_t_D;
        cta_gemm<true>(nu, n, n, Dk, nu, P, n, DtP, n, nullptr, 0, 1.0);          // D'P
        __syncthreads();
        // Sc = R + D'P D ;  rhs = D'P Ã
        for (int t = tid; t < nu * nu; t += RT) {
            const int i = t / nu, j = t % nu;
            double acc = 0.0;
            for (int r = 0; r < n; r++) acc = fma(DtP[(size_t)i * n + r], Dk[(size_t)r * nu + j], acc);
            Sc[t] = Rs[t] + acc;
        }
        cta_gemm<false>(nu, n, n, DtP, n, Ak, n, Kk, n, nullptr, 0, 1.0);
        __syncthreads();
        // Ku = Sc^-1 rhs : LU with partial pivoting of the tiny nu x nu block by one thread, columns in parallel
        if (tid == 0) {
            for (int c = 0; c < nu; c++) {
                int p = c;
                for (int i = c + 1; i < nu; i++) if (fabs(Sc[i * nu + c]) > fabs(Sc[p * nu + c])) p = i;
                red[c] = (double)p;
                if (p != c) for (int j = 0; j < nu; j++) { double t = Sc[c * nu + j]; Sc[c * nu + j] = Sc[p * nu + j]; Sc[p * nu + j] = t; }
                const double inv = 1.0 / Sc[c * nu + c];
                for (int i = c + 1; i < nu; i++) {
                    const double l = Sc[i * nu + c] * inv;
                    Sc[i * nu + c] = l;
                    for (int j = c + 1; j < nu; j++) Sc[i * nu + j] = fma(-l, Sc[c * nu + j], Sc[i * nu + j]);
                }
            }
        }
        __syncthreads();
        for (int c = tid; c < n; c += RT) {
            double x[8];
            for (int i = 0; i < nu; i++) x[i] = Kk[(size_t)i * n + c];
            for (int i = 0; i < nu; i++) { const int p = (int)red[i]; if (p != i) { double t = x[i]; x[i] = x[p]; x[p] = t; } }
            for (int i = 0; i < nu; i++) for (int j = 0; j < i; j++) x[i] = fma(-Sc[i * nu + j], x[j], x[i]);
            for (int i = nu - 1; i >= 0; i--) { for (int j = i + 1; j < nu; j++) x[i] = fma(-Sc[i * nu + j], x[j], x[i]); x[i] /= Sc[i * nu + i]; }
            for (int i = 0; i < nu; i++) { Kk[(size_t)i * n + c] = x[i]; Kout[(size_t)(k - 1) * nu * n + (size_t)i * n + c] = x[i]; }   // lqr.jl:162-164
        }
        __syncthreads();
        cta_gemm<false>(n, n, nu, Dk, nu, Kk, n, Abar, n, Ak, n, -1.0);          // Abar = Ã - D Ku         lqr.jl:169
        __syncthreads();
        cta_gemm<false>(n, n, n, P, n, Abar, n, W, n, nullptr, 0, 1.0);          // W = P Abar
        // RK = R Ku stored over DtP
        for (int t = tid; t < nu * n; t += RT) {
            const int i = t / n, c = t % n;
            double acc = 0.0;
            for (int j = 0; j < nu; j++) acc = fma(Rs[i * nu + j], Kk[(size_t)j * n + c], acc);
            DtP[t] = acc;
        }
        __syncthreads();
        cta_gemm<true>(n, n, n, Abar, n, W, n, P1, n, Qs, n, 1.0);               // Q + Abar' P Abar
        __syncthreads();
        double nd = 0.0;
        for (size_t t = tid; t < nn; t += RT) {                                  // + Ku' R Ku ; ‖P - P+‖_F
            const int i = (int)(t / n), j = (int)(t % n);
            double acc = P1[t];
            for (int q = 0; q < nu; q++) acc = fma(Kk[(size_t)q * n + i], DtP[(size_t)q * n + j], acc);
            P1[t] = acc;
            const double d = P[t] - acc;
            nd = fma(d, d, nd);
        }
        for (int o = 16; o > 0; o >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, o);
        if ((tid & 31) == 0) red[tid >> 5] = nd;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < RT / 32; w++) tot += red[w];
            brk = sqrt(tot) < 1e-5;                                              // lqr.jl:172
        }
        __syncthreads();
        if (brk) break;
        double* tmp = P; P = P1; P1 = tmp;                                       // lqr.jl:176
        __syncthreads();
    }
    if (k < 1) k = 1;                                                             // a completed loop leaves k == 1
    for (int k2 = k - 1; k2 >= 1; k2--) {                                         // lqr.jl:179-181 back-fill
        for (int t = tid; t < nu * n; t += RT) Kout[(size_t)(k2 - 1) * nu * n + t] = Kout[(size_t)k2 * nu * n + t];
        __syncthreads();
    }
    if (tid == 0) { kexit[s] = k; iters[s] = it; }
}

}  // namespace

bool riccati_dmma_eligible(int n, int nu);
int riccati_sweep_dmma_device(int S, int n, int nu, int N, const double* At, const double* D, long long stride_t_A, long long stride_t_D,
                              long long stride_s_A, long long stride_s_D, const double* Q, long long Q_bs, const double* R, long long R_bs, double* Ku,
                              int* kexit, int* iters, double* ws, long long ws_stride, cudaStream_t stream, int k_hi, int k_lo, int resume);

bool riccati_big_eligible(int n, int nu);
int riccati_sweep_big_device(int S, int n, int nu, int N, const double* At, const double* D, long long stride_t_A, long long stride_t_D,
                             long long stride_s_A, long long stride_s_D, const double* Q, long long Q_bs, const double* R, long long R_bs, double* Ku,
                             int* kexit, int* iters, double* ws, long long ws_stride, cudaStream_t stream);

size_t riccati_project_ws_doubles(int n, int nu, int ml) { return (size_t)ml * ml + (size_t)ml * (n + nu); }
size_t riccati_sweep_ws_doubles(int n, int nu) { return 4 * (size_t)n * n + 2 * (size_t)nu * n; }

int riccati_project_device(int P, int n, int nu, int ml, const double* A, const double* Bu, const double* Bl, const double* G, double* At, double* D,
                           double* ws, cudaStream_t stream) {
    const size_t smem = ((size_t)ml * ml + (size_t)ml * (n + nu)) * sizeof(double);
    const int use_smem = smem <= 200 * 1024 ? 1 : 0;
    if (use_smem) CCB_CUDA(cudaFuncSetAttribute(riccati_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    riccati_project_kernel<<<P, RT, use_smem ? smem : 0, stream>>>(n, nu, ml, A, Bu, Bl, G, At, D, ws, use_smem);
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

int riccati_sweep_device(int S, int n, int nu, int N, const double* At, const double* D, long long stride_t_A, long long stride_t_D,
                         long long stride_s_A, long long stride_s_D, const double* Q, long long Q_bs, const double* R, long long R_bs, double* Ku,
                         int* kexit, int* iters, double* ws, cudaStream_t stream, int k_hi, int k_lo, int resume) {
    // k_hi > 0: run only k = k_hi .. k_lo of the sweep (resume != 0: continuing an earlier launch) -- K4 only
    if (nu > 8) { set_error("riccati: more than 8 controlled eqcs is not supported yet"); return CCB_E_UNSUPPORTED; }
    if (k_hi > 0 && !(riccati_dmma_eligible(n, nu) && !getenv("CCB_NO_DMMA"))) { set_error("riccati: partial sweeps need the K4 kernel"); return CCB_E_UNSUPPORTED; }
    if (riccati_dmma_eligible(n, nu) && !getenv("CCB_NO_DMMA"))      // K4: FP64 tensor-core sweep, P / Abar resident in shared memory
        return riccati_sweep_dmma_device(S, n, nu, N, At, D, stride_t_A, stride_t_D, stride_s_A, stride_s_D, Q, Q_bs, R, R_bs, Ku, kexit, iters, ws,
                                         (long long)riccati_sweep_ws_doubles(n, nu), stream, k_hi, k_lo, resume);
    if (n > 108 && riccati_big_eligible(n, nu) && !getenv("CCB_NO_DMMA"))   // K4b: blocked FP64 tensor-core GEMMs on the global workspace
        return riccati_sweep_big_device(S, n, nu, N, At, D, stride_t_A, stride_t_D, stride_s_A, stride_s_D, Q, Q_bs, R, R_bs, Ku, kexit, iters, ws,
                                        (long long)riccati_sweep_ws_doubles(n, nu), stream);
    riccati_sweep_kernel<<<S, RT, 0, stream>>>(n, nu, N, At, D, stride_t_A, stride_t_D, stride_s_A, stride_s_D, Q, Q_bs, R, R_bs, Ku, kexit, iters, ws);
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
