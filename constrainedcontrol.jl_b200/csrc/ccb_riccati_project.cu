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

// MLMAX: compile-time bound of mλ (the substitution is straight-line code over MLMAX rows with the right-hand side in registers; the LU
// factors are padded with an identity to MLMAX x MLMAX so that no guard is needed); MINB: CTAs per SM the register budget is set for.
template <int MLMAX, int MINB>
__global__ void __launch_bounds__(PT, MINB) riccati_project_mech_kernel(const DevMech Mv, int n, int nu, int ml, int ldm, int ldy, const double* __restrict__ A,
                                                                   const double* __restrict__ Bu, const double* __restrict__ Bl,
                                                                   const double* __restrict__ G, double* __restrict__ At, double* __restrict__ D) {
    extern __shared__ __align__(16) double sm[];
    double* Ms = sm;                              // [MLMAX][ldm] G Bλ, then its LU factors (unit lower | upper), identity beyond mλ
    double* Y = Ms + (size_t)MLMAX * ldm;         // [ml][ldy]   G [A | Bu], then (G Bλ)^-1 G [A | Bu]
    double* Gc = Y + (size_t)ml * ldy;            // [ml][24]    row r of G on the parent block | child block of its joint
    Gc += (Gc - sm) & 1;                          // 16-byte aligned: phase 3 reads the Bλ blocks as double2
    double* Bc = Gc + (size_t)ml * 24;            // [ml][24]    column r of Bλ on the same two blocks
    int* rb = reinterpret_cast<int*>(Bc + (size_t)ml * 24);   // [ml][2] parent / child body of the row's joint (-1: origin)
    int* piv = rb + 2 * ml;                       // [ml]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t p = blockIdx.x;
    A += p * n * n; Bu += p * n * nu; Bl += p * n * ml; G += p * ml * n; At += p * n * n; D += p * n * nu;
    const int W = n + nu;

    for (int e = tid; e < MLMAX * ldm; e += PT) { const int i = e / ldm, c = e - ldm * i; Ms[e] = (i == c && i >= ml) ? 1.0 : 0.0; }
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

    // ---- M = G Bλ from the compact blocks: all threads (it used to be warp 0's job: a third of the serial chain the other warps wait for)
    for (int e = tid; e < ml * ml; e += PT) {
        const int r = e / ml, c = e - ml * r;
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const int ba = rb[2 * r + a];
            if (ba < 0) continue;
#pragma unroll
            for (int b = 0; b < 2; b++) {
                if (rb[2 * c + b] != ba) continue;
                const double2* g = reinterpret_cast<const double2*>(Gc + r * 24 + 12 * a);
                const double2* bl = reinterpret_cast<const double2*>(Bc + c * 24 + 12 * b);
#pragma unroll
                for (int k = 0; k < 6; k++) { const double2 g2 = g[k], b2 = bl[k]; acc = fma(g2.x, b2.x, acc); acc = fma(g2.y, b2.y, acc); }
            }
        }
        Ms[(size_t)r * ldm + c] = acc;
    }
    __syncthreads();

    if (warp == 0) {
        // ---- LU with partial pivoting (first row of largest magnitude), in place
        for (int k = 0; k < ml; k++) {
            // partial pivoting: the first row of largest magnitude.  |v| >= 0 orders like its bit pattern, so the warp-wide maximum is two
            // 32-bit REDUX maxima (high, low word) and a REDUX minimum over the row indices instead of a five-round shuffle butterfly (10 dependent shuffles per pivot)
            unsigned long long key = 0ull;
            int bi = k;
            for (int i = k + lane; i < ml; i += 32) {
                const unsigned long long v = (unsigned long long)__double_as_longlong(fabs(Ms[(size_t)i * ldm + k]));
                if (i == k + lane || v > key) { key = v; bi = i; }
            }
            const bool has = k + lane < ml;
            const unsigned hi = has ? (unsigned)(key >> 32) : 0u;
            const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (has && hi == mhi) ? (unsigned)key : 0u;
            const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
            bi = (int)__reduce_min_sync(0xffffffffu, (has && hi == mhi && lo == mlo) ? (unsigned)bi : 0xffffffffu);     // first such row
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
#pragma unroll 4
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
        // straight-line code: the factors are padded with an identity, rows / columns beyond mλ contribute exact zeros
#pragma unroll
        for (int i = 1; i < MLMAX; i++) {
            const double* ri = Ms + i * ldm;
            double acc = y[i];
#pragma unroll
            for (int c = 0; c < MLMAX; c++)
                if (c < i) acc = fma(-ri[c], y[c], acc);
            y[i] = acc;
        }
#pragma unroll
        for (int i = MLMAX - 1; i >= 0; i--) {
            const double* ri = Ms + i * ldm;
            double acc = y[i];
#pragma unroll
            for (int c = 0; c < MLMAX; c++)
                if (c > i) acc = fma(-ri[c], y[c], acc);
            y[i] = acc / ri[i];
        }
#pragma unroll
        for (int i = 0; i < MLMAX; i++)
            if (i < ml) Y[(size_t)i * ldy + j] = y[i];
    }
    __syncthreads();

    // ---- Ã = A - Bλ X_A,  D = Bu - Bλ X_u: one item = (body b, column j), the body's 12 rows in registers; they only see the multipliers of
    // the joints incident to b (a solved multiplier row is read once per item, the Bλ block entries are broadcast loads)
    for (int idx = tid; idx < Mv.nb * W; idx += PT) {
        const int b = idx / W, j = idx - W * b;
        const bool isA = j < n;
        const double* src = isA ? A + (size_t)(12 * b) * n + j : Bu + (size_t)(12 * b) * nu + (j - n);
        double* dst = isA ? At + (size_t)(12 * b) * n + j : D + (size_t)(12 * b) * nu + (j - n);
        const int ld = isA ? n : nu;
        double v[12];
#pragma unroll
        for (int kk = 0; kk < 12; kk++) v[kk] = src[(size_t)kk * ld];
        const int q1 = __ldg(Mv.bj_ptr + b + 1);
        for (int q = __ldg(Mv.bj_ptr + b); q < q1; q++) {
            const int jt = __ldg(Mv.bj_joint + q), side = __ldg(Mv.bj_side + q);
            const int r0 = __ldg(Mv.j_row0 + jt), dim = __ldg(Mv.j_dim + jt);
#pragma unroll
            for (int rr = 0; rr < 6; rr++) {
                if (rr < dim) {
                    const double yr = Y[(size_t)(r0 + rr) * ldy + j];
                    const double2* bc = reinterpret_cast<const double2*>(Bc + (r0 + rr) * 24 + 12 * side);
#pragma unroll
                    for (int k2 = 0; k2 < 6; k2++) {
                        const double2 c2 = bc[k2];
                        v[2 * k2] = fma(-c2.x, yr, v[2 * k2]);
                        v[2 * k2 + 1] = fma(-c2.y, yr, v[2 * k2 + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < 12; kk++) dst[(size_t)kk * ld] = v[kk];
    }
}

int project_mlmax(int ml) { return ml <= 12 ? 12 : ml <= 24 ? 24 : ml <= 36 ? 36 : ml <= 48 ? 48 : 64; }
size_t project_mech_smem(int n, int nu, int ml, int& ldm, int& ldy) {
    const int mlmax = project_mlmax(ml);
    ldm = mlmax | 1;
    ldy = (n + nu) | 1;
    // Ms [mlmax][ldm] | Y [ml][ldy] (+1 double so that Gc / Bc start 16-byte aligned) | Gc, Bc [ml][24] | rb, piv
    size_t d = (size_t)mlmax * ldm + (size_t)ml * ldy;
    d += d & 1;
    return (d + (size_t)ml * 48) * sizeof(double) + (size_t)ml * 3 * sizeof(int);
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
#define CCB_PROJ_LAUNCH(MLMAX_, MINB_)                                                                                                          \
    do {                                                                                                                                        \
        CCB_CUDA(cudaFuncSetAttribute(riccati_project_mech_kernel<MLMAX_, MINB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        riccati_project_mech_kernel<MLMAX_, MINB_><<<P, PT, smem, stream>>>(M, n, nu, ml, ldm, ldy, A, Bu, Bl, G, At, D);                       \
    } while (0)
    switch (project_mlmax(ml)) {
        case 12: CCB_PROJ_LAUNCH(12, 6); break;
        case 24: CCB_PROJ_LAUNCH(24, 5); break;
        case 36: CCB_PROJ_LAUNCH(36, 4); break;
        case 48: CCB_PROJ_LAUNCH(48, 3); break;
        default: CCB_PROJ_LAUNCH(64, 2); break;
    }
#undef CCB_PROJ_LAUNCH
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ccb
