// csrc/ccb_riccati_project.cu -- constraint projection of a linearisation, structure-aware:  Ã = A - Bλ (G Bλ)^-1 G A,  D = Bu - Bλ (G Bλ)^-1 G Bu
// (reference src/control/lqr.jl:151 `D = Bu - Bλ/(G*Bλ)*G*Bu` and the second block row of `M \ b`, lqr.jl:152-160; identical text in
// src/control/lqr_tracking.jl:89-99, where it runs once per trajectory knot: 64 x 1999 = 128 k projections for the Sawyer TrackingLQR config).
//
// The dense kernel (riccati_project_kernel, ccb_riccati.cu) treats G [mλ x n] and Bλ [n x mλ] as dense and runs a pivoted LU with three CTA
// barriers per pivot: 225 k warp instructions per Sawyer problem, 61 of the 90 ms of the TrackingLQR synthesis (r1 ncu: FP64 pipe 10 %,
// stalls long_scoreboard 4.0 / barrier 2.4 per issue).  What the matrices ARE: a constraint row belongs to one joint and touches the
// 12-column blocks of that joint's two bodies only, and a multiplier acts on the same two bodies only, so
//     G  = rows of 2 x 12 numbers,   Bλ = columns of 2 x 12 numbers,   (G Bλ)[r][c] != 0 only where the joints of r and c share a body.
// Two kernels (r2 history: ONE CTA-per-problem kernel with warp 0 factorising while warps 1-3 formed G [A | Bu] left three warps at the
// barrier for most of the LU's dependent chain -- ncu: barrier 6.4 stall cycles per issue, 33 ms for the 128 k problems; profiles/proj_r2_v*):
//   project_lu_kernel      ONE WARP per problem, 4 problems per CTA: M = G Bλ straight from the two body blocks, LU with partial pivoting
//                          inside the warp (REDUX pivot search, no CTA barrier), factors + pivots to the workspace.  20 LU warps per SM hide
//                          each other's dependent chains.
//   project_apply_kernel   one CTA of 128 threads per problem, every warp busy in every phase:
//       phase 0   compact copies Gc [mλ][24], Bc [mλ][24] (the two body blocks of every row / column), the LU factors (padded to MLMAX)
//       phase 1   Y = G [A | Bu]   (24 rows of A per joint, each loaded once per joint)
//       phase 2   one thread per column of Y: row swaps, forward and backward substitution, the column in registers, straight-line code
//       phase 3   Ã = A - Bλ Y, D = Bu - Bλ Y_u: the 12 rows of a body only see the multipliers of the joints incident to it
// Same LU (partial pivoting, first row of largest magnitude, same update formula) as the dense kernel; used whenever the mechanism is known
// (LQR / TrackingLQR constructors); ccb_dlqr with caller-supplied matrices keeps the dense kernel.
#include "ccb_internal.h"

namespace ccb {

namespace {

constexpr int PT = 128;   // threads per CTA

// ---------------------------------------------------------------------------------------------------------------- kernel A: M = G Bλ and its LU
// workspace of problem p: LU factors [ml][ml] row-major (unit lower | upper), then the ml pivot rows as int32
template <int MLMAX>
__global__ void __launch_bounds__(PT) project_lu_kernel(const DevMech Mv, int P, int n, int ml, int ldm, const double* __restrict__ Bl, const double* __restrict__ G,
                                                        double* __restrict__ ws, long long ws_stride) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t per_warp = (size_t)MLMAX * ldm + (size_t)((3 * ml + 1) / 2 + 1);      // Ms | rb [ml][2], piv [ml] (ints)
    double* Ms = sm + warp * per_warp;
    int* rb = reinterpret_cast<int*>(Ms + (size_t)MLMAX * ldm);
    int* piv = rb + 2 * ml;
    const long long p = (long long)blockIdx.x * (PT / 32) + warp;
    if (p >= P) return;                                     // warps are independent: no CTA barrier in this kernel
    Bl += (size_t)p * n * ml; G += (size_t)p * ml * n;
    for (int r = lane; r < ml; r += 32) {
        const int jt = __ldg(Mv.row_joint + r);
        rb[2 * r] = __ldg(Mv.j_parent + jt);
        rb[2 * r + 1] = __ldg(Mv.j_child + jt);
    }
    __syncwarp();
    // consecutive lanes take consecutive columns c of one row r: the G block is a broadcast load, the Bλ block rows are coalesced
    for (int e = lane; e < ml * ml; e += 32) {
        const int r = e / ml, c = e - ml * r;
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const int ba = rb[2 * r + a];
            if (ba < 0) continue;
            if (rb[2 * c] != ba && rb[2 * c + 1] != ba) continue;
            const double* g = G + (size_t)r * n + 12 * ba;
            const double* bl = Bl + (size_t)(12 * ba) * ml + c;
#pragma unroll
            for (int k = 0; k < 12; k++) acc = fma(g[k], bl[(size_t)k * ml], acc);
        }
        Ms[(size_t)r * ldm + c] = acc;
    }
    __syncwarp();
    {
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
    }
    double* out = ws + (size_t)p * ws_stride;
    for (int e = lane; e < ml * ml; e += 32) { const int r = e / ml, c = e - ml * r; out[e] = Ms[(size_t)r * ldm + c]; }
    int* pout = reinterpret_cast<int*>(out + (size_t)ml * ml);
    for (int r = lane; r < ml; r += 32) pout[r] = piv[r];
}

// ---------------------------------------------------------------------------------------------------------------- kernel B: apply
// MLMAX: compile-time bound of mλ (the substitution is straight-line code over MLMAX rows with the right-hand side in registers; the LU
// factors are padded with an identity to MLMAX x MLMAX so that no guard is needed); MINB: CTAs per SM the register budget is set for.
template <int MLMAX, int MINB>
__global__ void __launch_bounds__(PT, MINB) project_apply_kernel(const DevMech Mv, int n, int nu, int ml, int ldm, int ldy, const double* __restrict__ A,
                                                                 const double* __restrict__ Bu, const double* __restrict__ Bl, const double* __restrict__ G,
                                                                 double* __restrict__ At, double* __restrict__ D, const double* __restrict__ ws, long long ws_stride) {
    extern __shared__ __align__(16) double sm[];
    double* Mc = sm;                              // [MLMAX][MLMAX] LU factors of G Bλ, COLUMN-major (Mc[c MLMAX + i] = LU[i][c]), identity beyond mλ
    double* dinv = Mc + MLMAX * MLMAX;            // [MLMAX] reciprocals of the pivots U[i][i]
    double* Y = dinv + MLMAX;                     // [ml][ldy]   G [A | Bu], then (G Bλ)^-1 G [A | Bu]
    (void)ldm;
    double* Gc = Y + (size_t)ml * ldy;            // [ml][24]    row r of G on the parent block | child block of its joint
    Gc += (Gc - sm) & 1;                          // 16-byte aligned: phase 3 reads the Bλ blocks as double2
    double* Bc = Gc + (size_t)ml * 24;            // [ml][24]    column r of Bλ on the same two blocks
    int* rb = reinterpret_cast<int*>(Bc + (size_t)ml * 24);   // [ml][2] parent / child body of the row's joint (-1: origin)
    int* piv = rb + 2 * ml;                       // [ml]
    const int tid = threadIdx.x;
    const size_t p = blockIdx.x;
    A += p * n * n; Bu += p * n * nu; Bl += p * n * ml; G += p * ml * n; At += p * n * n; D += p * n * nu;
    ws += p * ws_stride;
    const int W = n + nu;

    for (int e = tid; e < MLMAX * MLMAX; e += PT) {
        const int c = e / MLMAX, i = e - MLMAX * c;
        const double v = (i < ml && c < ml) ? ws[(size_t)i * ml + c] : ((i == c) ? 1.0 : 0.0);
        Mc[e] = v;
        if (i == c) dinv[i] = 1.0 / v;
    }
    for (int r = tid; r < ml; r += PT) {
        const int jt = __ldg(Mv.row_joint + r);
        rb[2 * r] = __ldg(Mv.j_parent + jt);
        rb[2 * r + 1] = __ldg(Mv.j_child + jt);
        piv[r] = reinterpret_cast<const int*>(ws + (size_t)ml * ml)[r];
    }
    __syncthreads();
    for (int e = tid; e < ml * 24; e += PT) {
        const int r = e / 24, q = e - 24 * r, a = q >= 12 ? 1 : 0, k = q - 12 * a, b = rb[2 * r + a];
        Gc[e] = b >= 0 ? G[(size_t)r * n + 12 * b + k] : 0.0;
        Bc[e] = b >= 0 ? Bl[(size_t)(12 * b + k) * ml + r] : 0.0;
    }
    __syncthreads();
    // ---- Y = G [A | Bu]: one item = (joint, column); the 24 rows of [A | Bu] the joint sees are loaded once
    for (int idx = tid; idx < Mv.nj * W; idx += PT) {
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
            for (int k2 = 0; k2 < 6; k2++) {
                const double v0 = src[(size_t)(2 * k2) * ld], v1 = src[(size_t)(2 * k2 + 1) * ld];
#pragma unroll
                for (int rr = 0; rr < 6; rr++)
                    if (rr < dim) {
                        const double2 g2 = *reinterpret_cast<const double2*>(Gc + (r0 + rr) * 24 + 12 * a + 2 * k2);
                        acc[rr] = fma(g2.x, v0, acc[rr]);
                        acc[rr] = fma(g2.y, v1, acc[rr]);
                    }
            }
        }
#pragma unroll
        for (int rr = 0; rr < 6; rr++)
            if (rr < dim) Y[(size_t)(r0 + rr) * ldy + j] = acc[rr];
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
        // straight-line, COLUMN-oriented substitution: eliminating unknown c updates every remaining row independently (instruction-level
        // parallelism MLMAX - c instead of one dependent accumulation chain per row), the factor column is read as 16-byte pairs (broadcast
        // loads), the pivots enter as precomputed reciprocals.  The factors are padded with an identity: rows beyond mλ stay exactly zero.
#pragma unroll
        for (int c = 0; c < MLMAX - 1; c++) {                      // L y' = y  (unit lower)
            const double yc = y[c];
            const double2* col = reinterpret_cast<const double2*>(Mc + c * MLMAX);
#pragma unroll
            for (int i2 = 0; i2 < MLMAX / 2; i2++) {
                if (2 * i2 + 1 > c) {
                    const double2 m = col[i2];
                    if (2 * i2 > c) y[2 * i2] = fma(-m.x, yc, y[2 * i2]);
                    y[2 * i2 + 1] = fma(-m.y, yc, y[2 * i2 + 1]);
                }
            }
        }
#pragma unroll
        for (int c = MLMAX - 1; c >= 0; c--) {                     // U x = y'
            const double xc = y[c] * dinv[c];
            y[c] = xc;
            const double2* col = reinterpret_cast<const double2*>(Mc + c * MLMAX);
#pragma unroll
            for (int i2 = 0; i2 < MLMAX / 2; i2++) {
                if (2 * i2 < c) {
                    const double2 m = col[i2];
                    y[2 * i2] = fma(-m.x, xc, y[2 * i2]);
                    if (2 * i2 + 1 < c) y[2 * i2 + 1] = fma(-m.y, xc, y[2 * i2 + 1]);
                }
            }
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
    // apply kernel: Mc [mlmax][mlmax] | dinv [mlmax] | Y [ml][ldy] (+1 double so that Gc / Bc start 16-byte aligned) | Gc, Bc [ml][24] | rb, piv
    size_t d = (size_t)mlmax * mlmax + mlmax + (size_t)ml * ldy;
    d += d & 1;
    return (d + (size_t)ml * 48) * sizeof(double) + (size_t)ml * 3 * sizeof(int);
}

}  // namespace

// 0 = launched; 1 = not eligible (caller falls back to the dense kernel); < 0 error.  ws: P * riccati_project_ws_doubles(n, nu, ml) doubles.
int riccati_project_mech_device(const DevMech& M, int P, int nu, const double* A, const double* Bu, const double* Bl, const double* G, double* At,
                                double* D, double* ws, long long ws_stride, cudaStream_t stream) {
    const int n = 12 * M.nb, ml = M.ml;
    if (ml < 1 || ml > 64 || getenv("CCB_PROJECT_DENSE")) return 1;
    if (ws_stride < (long long)ml * ml + (ml + 1) / 2) return 1;
    int ldm, ldy;
    const size_t smem = project_mech_smem(n, nu, ml, ldm, ldy);
    if (smem > 100 * 1024) return 1;
    const int mlmax = project_mlmax(ml);
    const size_t smem_lu = (PT / 32) * ((size_t)mlmax * ldm + (size_t)((3 * ml + 1) / 2 + 1)) * sizeof(double);
    const int grid_lu = (P + PT / 32 - 1) / (PT / 32);
#define CCB_PROJ_LAUNCH(MLMAX_, MINB_)                                                                                                     \
    do {                                                                                                                                   \
        CCB_CUDA(cudaFuncSetAttribute(project_lu_kernel<MLMAX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lu));              \
        project_lu_kernel<MLMAX_><<<grid_lu, PT, smem_lu, stream>>>(M, P, n, ml, ldm, Bl, G, ws, ws_stride);                               \
        CCB_CUDA(cudaFuncSetAttribute(project_apply_kernel<MLMAX_, MINB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        project_apply_kernel<MLMAX_, MINB_><<<P, PT, smem, stream>>>(M, n, nu, ml, ldm, ldy, A, Bu, Bl, G, At, D, ws, ws_stride);          \
    } while (0)
    switch (mlmax) {
        case 12: CCB_PROJ_LAUNCH(12, 6); break;
        case 24: CCB_PROJ_LAUNCH(24, 5); break;
        case 36: CCB_PROJ_LAUNCH(36, 4); break;
        case 48: CCB_PROJ_LAUNCH(48, 3); break;
        default: CCB_PROJ_LAUNCH(64, 2); break;
    }
#undef CCB_PROJ_LAUNCH
    g_launches += 2;
    CCB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ccb
