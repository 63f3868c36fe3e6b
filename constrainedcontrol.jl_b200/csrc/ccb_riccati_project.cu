// csrc/ccb_riccati_project.cu -- constraint projection of a linearisation, structure-aware:  Ã = A - Bλ (G Bλ)^-1 G A,  D = Bu - Bλ (G Bλ)^-1 G Bu
// (reference src/control/lqr.jl:151 `D = Bu - Bλ/(G*Bλ)*G*Bu` and the second block row of `M \ b`, lqr.jl:152-160; identical text in
// src/control/lqr_tracking.jl:89-99, where it runs once per trajectory knot: 64 x 1999 = 128 k projections for the Sawyer TrackingLQR config).
//
// The dense kernel (riccati_project_kernel, ccb_riccati.cu) treats G [mλ x n] and Bλ [n x mλ] as dense and runs a pivoted LU with three CTA
// barriers per pivot: 225 k warp instructions per Sawyer problem, 61 of the 90 ms of the TrackingLQR synthesis (r1 ncu: FP64 pipe 10 %,
// stalls long_scoreboard 4.0 / barrier 2.4 per issue).  What the matrices ARE: a constraint row belongs to one joint and touches the
// 12-column blocks of that joint's two bodies only, and a multiplier acts on the same two bodies only, so
//     G  = rows of 2 x 12 numbers,   Bλ = columns of 2 x 12 numbers,   (G Bλ)[r][c] != 0 only where the joints of r and c share a body.
// One CTA of 128 threads per linearisation, everything but A / Ã in shared memory (50 KB for the Sawyer arm, 4 CTAs per SM):
//     phase 0   compact copies Gc [mλ][24], Bc [mλ][24] (the two body blocks of every row / column)
//     phase 1   warp 0: M = G Bλ from the compact blocks, then LU with partial pivoting inside the warp (no CTA barrier)
//               warps 1-3, at the same time: Y = G [A | Bu]   (24 rows of A per joint, each loaded once per joint)
//     phase 2   one thread per column of Y: row swaps, forward and backward substitution with the column in registers
//     phase 3   Ã = A - Bλ Y, D = Bu - Bλ Y_u: a row of body b only sees the multipliers of the joints incident to b
// ~20 k warp instructions per Sawyer problem.  Same LU (partial pivoting, same update formula) as the dense kernel; used whenever the
// mechanism is known (LQR / TrackingLQR constructors); ccb_dlqr with caller-supplied matrices keeps the dense kernel.
#include "ccb_internal.h"

namespace ccb {

namespace {

constexpr int PT = 128;   // threads per CTA

template <int MLMAX>
__global__ void __launch_bounds__(PT, MLMAX > 40 ? 2 : 4) riccati_project_mech_kernel(const DevMech Mv, int n, int nu, int ml, int ldm, int ldy, const double* __restrict__ A,
                                                                   const double* __restrict__ Bu, const double* __restrict__ Bl,
                                                                   const double* __restrict__ G, double* __restrict__ At, double* __restrict__ D) {
    extern __shared__ __align__(16) double sm[];
    double* Ms = sm;                              // [ml][ldm]   G Bλ, then its LU factors (unit lower | upper)
    double* Y = Ms + (size_t)ml * ldm;            // [ml][ldy]   G [A | Bu], then (G Bλ)^-1 G [A | Bu]
    double* Gc = Y + (size_t)ml * ldy;            // [ml][24]    row r of G on the parent block | child block of its joint
    double* Bc = Gc + (size_t)ml * 24;            // [ml][24]    column r of Bλ on the same two blocks
    int* rb = reinterpret_cast<int*>(Bc + (size_t)ml * 24);   // [ml][2] parent / child body of the row's joint (-1: origin)
    int* piv = rb + 2 * ml;                       // [ml]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t p = blockIdx.x;
    A += p * n * n; Bu += p * n * nu; Bl += p * n * ml; G += p * ml * n; At += p * n * n; D += p * n * nu;
    const int W = n + nu;

    for (int r = tid; r < ml; r += PT) {
        const int jt = __ldg(Mv.row_joint + r);
        rb[2 * r] = __ldg(Mv.j_parent + jt);
        rb[2 * r + 1] = __ldg(Mv.j_child + jt);
    }
    __syncthreads();
    for (int e = tid; e < ml * 24; e += PT) {
        const int r = e / 24, q = e - 24 * r, a = q >= 12 ? 1 : 0, k = q - 12 * a, b = rb[2 * r + a];
        Gc[e] = b >= 0 ? G[(size_t)r * n + 12 * b + k] : 0.0;
        Bc[e] = b >= 0 ? Bl[(size_t)(12 * b + k) * ml + r] : 0.0;
    }
    __syncthreads();

    if (warp == 0) {
        // ---- M = G Bλ from the compact blocks
        for (int e = lane; e < ml * ml; e += 32) {
            const int r = e / ml, c = e - ml * r;
            double acc = 0.0;
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int ba = rb[2 * r + a];
                if (ba < 0) continue;
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    if (rb[2 * c + b] != ba) continue;
                    const double* g = Gc + r * 24 + 12 * a;
                    const double* bl = Bc + c * 24 + 12 * b;
#pragma unroll
                    for (int k = 0; k < 12; k++) acc = fma(g[k], bl[k], acc);
                }
            }
            Ms[(size_t)r * ldm + c] = acc;
        }
        __syncwarp();
        // ---- LU with partial pivoting (first row of largest magnitude), in place
        for (int k = 0; k < ml; k++) {
            double best = -1.0;
            int bi = k;
            for (int i = k + lane; i < ml; i += 32) { const double v = fabs(Ms[(size_t)i * ldm + k]); if (v > best) { best = v; bi = i; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) piv[k] = bi;
            if (bi != k)
                for (int c = lane; c < ml; c += 32) { const double t = Ms[(size_t)k * ldm + c]; Ms[(size_t)k * ldm + c] = Ms[(size_t)bi * ldm + c]; Ms[(size_t)bi * ldm + c] = t; }
            __syncwarp();
            const double inv = 1.0 / Ms[(size_t)k * ldm + k];
            const double* rk = Ms + (size_t)k * ldm;
            for (int i = k + 1 + lane; i < ml; i += 32) {
                double* ri = Ms + (size_t)i * ldm;
                const double l = ri[k] * inv;
                ri[k] = l;
                for (int c = k + 1; c < ml; c++) ri[c] = fma(-l, rk[c], ri[c]);
            }
            __syncwarp();
        }
    } else {
        // ---- Y = G [A | Bu]: one item = (joint, column); the 24 rows of [A | Bu] the joint sees are loaded once
        for (int idx = tid - 32; idx < Mv.nj * W; idx += PT - 32) {
            const int jt = idx / W, j = idx - W * jt;
            const int r0 = __ldg(Mv.j_row0 + jt), dim = __ldg(Mv.j_dim + jt);
            double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int b = a == 0 ? __ldg(Mv.j_parent + jt) : __ldg(Mv.j_child + jt);
                if (b < 0) continue;
                const double* src = j < n ? A + (size_t)(12 * b) * n + j : Bu + (size_t)(12 * b) * nu + (j - n);
                const int ld = j < n ? n : nu;
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    const double v = src[(size_t)k * ld];
#pragma unroll
                    for (int rr = 0; rr < 6; rr++)
                        if (rr < dim) acc[rr] = fma(Gc[(r0 + rr) * 24 + 12 * a + k], v, acc[rr]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < 6; rr++)
                if (rr < dim) Y[(size_t)(r0 + rr) * ldy + j] = acc[rr];
        }
    }
    __syncthreads();

    // ---- (G Bλ) X = Y: one thread per right-hand side, the column lives in registers
    for (int j = tid; j < W; j += PT) {
        for (int k = 0; k < ml; k++) {
            const int pk = piv[k];
            if (pk != k) { const double t = Y[(size_t)k * ldy + j]; Y[(size_t)k * ldy + j] = Y[(size_t)pk * ldy + j]; Y[(size_t)pk * ldy + j] = t; }
        }
        double y[MLMAX];
#pragma unroll
        for (int i = 0; i < MLMAX; i++) y[i] = i < ml ? Y[(size_t)i * ldy + j] : 0.0;
        // constant trip counts everywhere so that the loops unroll completely and y[] stays in registers (a triangular inner bound left
        // y[] in local memory for MLMAX = 40); rows / columns beyond ml multiply exact zeros of y and are skipped by the uniform guards
#pragma unroll
        for (int i = 1; i < MLMAX; i++) {
            if (i < ml) {
                const double* ri = Ms + (size_t)i * ldm;
                double acc = y[i];
#pragma unroll
                for (int c = 0; c < MLMAX; c++)
                    if (c < i) acc = fma(-ri[c], y[c], acc);
                y[i] = acc;
            }
        }
#pragma unroll
        for (int i = MLMAX - 1; i >= 0; i--) {
            if (i < ml) {
                const double* ri = Ms + (size_t)i * ldm;
                double acc = y[i];
#pragma unroll
                for (int c = 0; c < MLMAX; c++)
                    if (c > i && c < ml) acc = fma(-ri[c], y[c], acc);
                y[i] = acc / ri[i];
            }
        }
#pragma unroll
        for (int i = 0; i < MLMAX; i++)
            if (i < ml) Y[(size_t)i * ldy + j] = y[i];
    }
    __syncthreads();

    // ---- Ã = A - Bλ X_A,  D = Bu - Bλ X_u: row i of body b sees the multipliers of the joints incident to b
    for (int idx = tid; idx < n * W; idx += PT) {
        const int i = idx / W, j = idx - W * i;
        const int b = i / 12, kk = i - 12 * b;
        double v = j < n ? A[(size_t)i * n + j] : Bu[(size_t)i * nu + (j - n)];
        const int q1 = __ldg(Mv.bj_ptr + b + 1);
        for (int q = __ldg(Mv.bj_ptr + b); q < q1; q++) {
            const int jt = __ldg(Mv.bj_joint + q), side = __ldg(Mv.bj_side + q);
            const int r0 = __ldg(Mv.j_row0 + jt), dim = __ldg(Mv.j_dim + jt);
            for (int rr = 0; rr < dim; rr++) v = fma(-Bc[(r0 + rr) * 24 + 12 * side + kk], Y[(size_t)(r0 + rr) * ldy + j], v);
        }
        if (j < n) At[(size_t)i * n + j] = v;
        else D[(size_t)i * nu + (j - n)] = v;
    }
}

size_t project_mech_smem(int n, int nu, int ml, int& ldm, int& ldy) {
    ldm = ml | 1;
    ldy = (n + nu) | 1;
    return ((size_t)ml * ldm + (size_t)ml * ldy + (size_t)ml * 48) * sizeof(double) + (size_t)ml * 3 * sizeof(int);
}

}  // namespace

// 0 = launched; 1 = not eligible (caller falls back to the dense kernel); < 0 error
int riccati_project_mech_device(const DevMech& M, int P, int nu, const double* A, const double* Bu, const double* Bl, const double* G, double* At,
                                double* D, cudaStream_t stream) {
    const int n = 12 * M.nb, ml = M.ml;
    if (ml < 1 || ml > 64 || getenv("CCB_PROJECT_DENSE")) return 1;
    int ldm, ldy;
    const size_t smem = project_mech_smem(n, nu, ml, ldm, ldy);
    if (smem > 100 * 1024) return 1;
#define CCB_PROJ_LAUNCH(MLMAX_)                                                                                                          \
    do {                                                                                                                                 \
        CCB_CUDA(cudaFuncSetAttribute(riccati_project_mech_kernel<MLMAX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        riccati_project_mech_kernel<MLMAX_><<<P, PT, smem, stream>>>(M, n, nu, ml, ldm, ldy, A, Bu, Bl, G, At, D);                       \
    } while (0)
    if (ml <= 24) CCB_PROJ_LAUNCH(24);
    else if (ml <= 40) CCB_PROJ_LAUNCH(40);
    else CCB_PROJ_LAUNCH(64);
#undef CCB_PROJ_LAUNCH
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ccb
