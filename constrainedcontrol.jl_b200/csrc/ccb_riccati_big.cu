// csrc/ccb_riccati_big.cu -- K4b: the constrained Riccati sweep (reference src/control/lqr.jl:141-184, lqr_tracking.jl:87-115) for
// state dimensions whose P and Abar do not fit in one SM's shared memory (n > 108: 16- and 32-link pendulums, n = 204 / 396).
//
// Same projected recursion as K3 / K4 (ccb_riccati.cu header):
//     Ku = (R + D'P D)^-1 D'P Ã ,  Abar = Ã - D Ku ,  W = P Abar ,  P+ = Q + Abar' W + Ku' R Ku .
// One CTA of 16 warps per system; P, P+, Abar, W live in a global workspace (L2 / HBM) and the two n^3 products are blocked
// FP64 tensor-core GEMMs: the output is cut into square blocks of tb x tb 8x8 tiles (tb <= 16, chosen so that the blocks
// cover n with the least padding: 4 x 104 for n = 396, 2 x 104 for n = 204), warp w owns the 8-row strip w of the block, and
// 16-deep k-panels of both operands are staged through shared memory (cp.async, four stages in flight, one barrier per panel).  Both products are of the form A' B with row-major A, B:  W = P'Abar because P is symmetric,  Abar'W as is;
// so one panel layout serves both and every fragment load is conflict free (row length 132 = 4 mod 16).
// P+ is symmetric: only blocks on or above the diagonal are formed and mirrored, and the diagonal blocks are averaged with
// their transposes so that P stays exactly symmetric (for why that matters).
#include <algorithm>
#include <cstdlib>
#include "ccb_dev_math.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

constexpr int BT = 16;           // tiles per block edge (max) = warps per CTA (16 warps: 128 registers per thread)
constexpr int NTB = BT * 32;     // threads per CTA
constexpr int BK = 16;           // k-panel depth
constexpr int PLD = 132;         // panel row length in shared memory: = 4 (mod 16) so that a half-warp's fragment load touches 16 distinct banks; >= 8 BT
constexpr int PANEL = BK * PLD;  // doubles per operand panel
constexpr int NSLOT = (2 * PANEL / 2 + NTB - 1) / NTB;   // 16-byte pieces per thread of one panel pair
constexpr int NST = 4;           // panel pairs in flight (cp.async stages)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// acc (+)= A' B restricted to rows i0 .. i0 + 8 tbm of A' (columns of A) and columns j0 .. j0 + 8 tbn of B.
// A, B: [K][ld] row-major in global memory, valid columns < ncol.  sm: 2 NST PANEL doubles.  All NTB threads call it.
// On exit thread (warp w < tbm, g, t4) holds in acc[j] the elements (i0 + 8 w + g, j0 + 8 j + 2 t4 + {0, 1}).
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int src_bytes) {      // src_bytes 0: zero fill
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int TBN>
__device__ __forceinline__ void block_gemm_tn_t(int K, int ncol, const double* A, int lda, int i0, int tbm, const double* B, int ldb, int j0,
                                                double* sm, double (&acc)[BT][2]) {
    constexpr int tbn = TBN;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int j = 0; j < BT; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
    const int np = (K + BK - 1) / BK;
    // a panel pair is PANEL 16-byte pieces [op][row][column pair]; thread t copies pieces t, t + NTB, ... straight into shared memory
    // (cp.async, zero fill outside the matrix): NST stages in flight hide the HBM latency behind NST - 1 panels of DMMAs
    auto issue = [&](int p) {
        if (p < np) {
            double* dst = sm + (size_t)(p % NST) * 2 * PANEL;
#pragma unroll
            for (int q = 0; q < NSLOT; q++) {
                const int e = tid + q * NTB;
                if (e < PANEL) {
                    const int op = e / (BK * (PLD / 2)), rem = e % (BK * (PLD / 2)), row = rem / (PLD / 2), col = 2 * (rem % (PLD / 2));
                    const int kk = p * BK + row;
                    const int x = (op ? j0 : i0) + col;
                    const bool ok = kk < K && x < ncol && col < 8 * (op ? tbn : tbm);
                    const double* src = ok ? (op ? B : A) + (size_t)kk * (op ? ldb : lda) + x : A;
                    cp_async16(dst + 2 * e, src, ok ? 16 : 0);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int p = 0; p < NST - 1; p++) issue(p);
    for (int p = 0; p < np; p++) {
        cp_async_wait<NST - 2>();          // my pieces of panel p have landed
        __syncthreads();                   // ... and everybody else's; stage (p - 1) % NST is free again
        issue(p + NST - 1);
        if (w < tbm) {
            const double* As = sm + (size_t)(p % NST) * 2 * PANEL + 8 * w + g;
            const double* Bs = sm + (size_t)(p % NST) * 2 * PANEL + PANEL + g;
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                const double a = As[(kk + t4) * PLD];
                const double* br = Bs + (kk + t4) * PLD;
                // fragment loads in groups of (up to) 8 ahead of their DMMAs: the load latency overlaps instead of serialising
#pragma unroll
                for (int j0g = 0; j0g < TBN; j0g += 8) {
                    double b[8];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j0g + j < TBN) b[j] = br[8 * (j0g + j)];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j0g + j < TBN) dmma(acc[j0g + j][0], acc[j0g + j][1], a, b[j]);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();                       // the stages may be refilled by the next block
}

// workspace per system: P, P1, Abar, W (4 n^2).  Dynamic shared memory: 4 PANEL + 3 nu n doubles.
template <int TB>
__global__ void __launch_bounds__(NTB, 1) riccati_sweep_big_kernel(int n, int nu, int N, const double* __restrict__ At, const double* __restrict__ D,
                                                                    long long stride_t_A, long long stride_t_D, long long stride_s_A, long long stride_s_D,
                                                                    const double* __restrict__ Q, long long Q_bs, const double* __restrict__ R, long long R_bs,
                                                                    double* __restrict__ Ku, int* __restrict__ kexit, int* __restrict__ iters, double* ws,
                                                                    long long ws_stride) {
    extern __shared__ double big_sm[];
    __shared__ double red[BT];
    __shared__ double Sc[64], Rs[64], piv[8];
    __shared__ int brk;
    const size_t s = blockIdx.x, nn = (size_t)n * n;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    double* panels = big_sm;
    double* DtP = big_sm + 2 * NST * PANEL;            // [nu][n]  (later R Ku)
    double* Kk = DtP + (size_t)nu * n;           // [nu][n]
    double* RK = Kk + (size_t)nu * n;            // [nu][n]
    double* P = ws + s * ws_stride;
    double* P1 = P + nn;
    double* Abar = P1 + nn;
    double* W = Abar + nn;
    const double* Qs = Q + s * Q_bs;
    const double* Rg = R + s * R_bs;
    double* Kout = Ku + s * (size_t)(N - 1) * nu * n;
    constexpr int tb = TB;                                                                  // blocks of TB x TB tiles (the last one zero padded)
    const int T = (n + 7) / 8, nblk = (T + tb - 1) / tb;

    for (size_t t = tid; t < nn; t += NTB) P[t] = Qs[t];             // lqr.jl:147  Pk = Q
    for (int t = tid; t < nu * nu; t += NTB) Rs[t] = Rg[t];
    __syncthreads();

    int k, it = 0;
    for (k = N - 1; k >= 1; k--) {                                    // lqr.jl:150
        it++;
        const double* Ak = At + s * stride_s_A + (size_t)(k - 1) * stride_t_A;
        const double* Dk = D + s * stride_s_D + (size_t)(k - 1) * stride_t_D;
        // ---- D'P (nu x n): a thread per column, coalesced rows of P
        for (int j = tid; j < n; j += NTB) {
            double a[8];
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] = 0.0;
#pragma unroll 8
            for (int r = 0; r < n; r++) {
                const double p = P[(size_t)r * n + j];
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (q < nu) a[q] = fma(Dk[(size_t)r * nu + q], p, a[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (q < nu) DtP[(size_t)q * n + j] = a[q];
        }
        __syncthreads();
        // ---- Sc = R + D'P D ;  rhs = D'P Ã
        for (int t = tid; t < nu * nu; t += NTB) {
            const int i = t / nu, j = t % nu;
            double acc = 0.0;
            for (int r = 0; r < n; r++) acc = fma(DtP[(size_t)i * n + r], Dk[(size_t)r * nu + j], acc);
            Sc[t] = Rs[t] + acc;
        }
        for (int j = tid; j < n; j += NTB) {
            double a[8];
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] = 0.0;
#pragma unroll 8
            for (int r = 0; r < n; r++) {
                const double av = Ak[(size_t)r * n + j];
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (q < nu) a[q] = fma(DtP[(size_t)q * n + r], av, a[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (q < nu) Kk[(size_t)q * n + j] = a[q];
        }
        __syncthreads();
        // ---- Ku = Sc^-1 rhs : LU with partial pivoting of the tiny nu x nu block by one thread, columns in parallel
        if (tid == 0) {
            for (int c = 0; c < nu; c++) {
                int p = c;
                for (int i = c + 1; i < nu; i++) if (fabs(Sc[i * nu + c]) > fabs(Sc[p * nu + c])) p = i;
                piv[c] = (double)p;
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
        for (int c = tid; c < n; c += NTB) {
            double x[8];
            for (int i = 0; i < nu; i++) x[i] = Kk[(size_t)i * n + c];
            for (int i = 0; i < nu; i++) { const int p = (int)piv[i]; if (p != i) { double t = x[i]; x[i] = x[p]; x[p] = t; } }
            for (int i = 0; i < nu; i++) for (int j = 0; j < i; j++) x[i] = fma(-Sc[i * nu + j], x[j], x[i]);
            for (int i = nu - 1; i >= 0; i--) { for (int j = i + 1; j < nu; j++) x[i] = fma(-Sc[i * nu + j], x[j], x[i]); x[i] /= Sc[i * nu + i]; }
            for (int i = 0; i < nu; i++) { Kk[(size_t)i * n + c] = x[i]; Kout[(size_t)(k - 1) * nu * n + (size_t)i * n + c] = x[i]; }   // lqr.jl:162-164
            for (int i = 0; i < nu; i++) {
                double acc = 0.0;
                for (int j = 0; j < nu; j++) acc = fma(Rs[i * nu + j], x[j], acc);
                RK[(size_t)i * n + c] = acc;
            }
        }
        __syncthreads();
        // ---- Abar = Ã - D Ku                                                                        lqr.jl:169
        for (int i = w; i < n; i += BT) {                             // a row per warp: coalesced, no integer divisions
            double dq[8];
#pragma unroll
            for (int q = 0; q < 8; q++) dq[q] = q < nu ? -Dk[(size_t)i * nu + q] : 0.0;
#pragma unroll 4
            for (int j = lane; j < n; j += 32) {
                double acc = Ak[(size_t)i * n + j];
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (q < nu) acc = fma(dq[q], Kk[(size_t)q * n + j], acc);
                Abar[(size_t)i * n + j] = acc;
            }
        }
        __syncthreads();
        // ---- W = P Abar = P' Abar
        for (int bi = 0; bi < nblk; bi++)
            for (int bj = 0; bj < nblk; bj++) {
                const int ti0 = bi * tb, tj0 = bj * tb, tbm = min(tb, T - ti0);
                double acc[BT][2];
                block_gemm_tn_t<TB>(n, n, P, n, 8 * ti0, tbm, Abar, n, 8 * tj0, panels, acc);
                const int i = 8 * (ti0 + w) + g;
                if (w < tbm && i < n) {
#pragma unroll
                    for (int j = 0; j < TB; j++) {
                        const int c = 8 * (tj0 + j) + 2 * t4;
                        if (c < n) *reinterpret_cast<double2*>(W + (size_t)i * n + c) = make_double2(acc[j][0], acc[j][1]);
                    }
                }
            }
        __syncthreads();
        // ---- P+ = Q + Abar' W + Ku' R Ku : blocks on and above the diagonal, mirrored                lqr.jl:170
        double nd = 0.0;
        for (int bi = 0; bi < nblk; bi++)
            for (int bj = bi; bj < nblk; bj++) {
                const int ti0 = bi * tb, tj0 = bj * tb, tbm = min(tb, T - ti0);
                double acc[BT][2];
                block_gemm_tn_t<TB>(n, n, Abar, n, 8 * ti0, tbm, W, n, 8 * tj0, panels, acc);
                const int i = 8 * (ti0 + w) + g;
                if (w < tbm) {                                   // warp-uniform: the shuffles below need the whole warp
                    const bool iok = i < n;
                    double ki[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) ki[q] = (q < nu && iok) ? Kk[(size_t)q * n + i] : 0.0;
#pragma unroll
                    for (int j = 0; j < TB; j++) {
                        const int c = 8 * (tj0 + j) + 2 * t4;
                        const bool ok = iok && c < n;
                        double v0 = acc[j][0], v1 = acc[j][1];
                        if (ok) {
                            const double2 qv = *reinterpret_cast<const double2*>(Qs + (size_t)i * n + c);
                            v0 += qv.x; v1 += qv.y;
#pragma unroll
                            for (int q = 0; q < 8; q++)
                                if (q < nu) { v0 = fma(ki[q], RK[(size_t)q * n + c], v0); v1 = fma(ki[q], RK[(size_t)q * n + c + 1], v1); }
                        }
                        if (bj == bi && j == w) {      // diagonal tile: average with its own transpose (exact symmetry)
                            const int src0 = (2 * t4) * 4 + (g >> 1), src1 = (2 * t4 + 1) * 4 + (g >> 1);
                            const double u00 = __shfl_sync(0xffffffffu, v0, src0), u01 = __shfl_sync(0xffffffffu, v1, src0);
                            const double u10 = __shfl_sync(0xffffffffu, v0, src1), u11 = __shfl_sync(0xffffffffu, v1, src1);
                            v0 = 0.5 * (v0 + ((g & 1) ? u01 : u00));
                            v1 = 0.5 * (v1 + ((g & 1) ? u11 : u10));
                        }
                        if (ok && (bj > bi || j >= w)) {
                            *reinterpret_cast<double2*>(P1 + (size_t)i * n + c) = make_double2(v0, v1);
                            const double2 po = *reinterpret_cast<const double2*>(P + (size_t)i * n + c);       // ‖P - P+‖_F^2: a mirrored tile counts twice
                            const double d0 = po.x - v0, d1 = po.y - v1, wgt = (bj > bi || j > w) ? 2.0 : 1.0;
                            nd = fma(wgt * d0, d0, nd); nd = fma(wgt * d1, d1, nd);
                        }
                        if (ok && (bj > bi || j > w)) { P1[(size_t)c * n + i] = v0; P1[(size_t)(c + 1) * n + i] = v1; }
                    }
                }
            }
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, o);
        if (lane == 0) red[w] = nd;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int q = 0; q < BT; q++) tot += red[q];
            brk = sqrt(tot) < 1e-5;                                                                  // lqr.jl:172
        }
        __syncthreads();
        if (brk) break;
        double* tmp = P; P = P1; P1 = tmp;                                                           // lqr.jl:176
    }
    if (k < 1) k = 1;                                                                                // a completed loop leaves k == 1
    for (int k2 = k - 1; k2 >= 1; k2--) {                                                            // lqr.jl:179-181 back-fill
        for (int t = tid; t < nu * n; t += NTB) Kout[(size_t)(k2 - 1) * nu * n + t] = Kout[(size_t)k2 * nu * n + t];
        __syncthreads();
    }
    if (tid == 0) { kexit[s] = k; iters[s] = it; }
}

}  // namespace

// n even (it is 12 nb), nu <= 8; the nu x n operands must fit next to the panels
bool riccati_big_eligible(int n, int nu) {
    return n % 2 == 0 && nu >= 1 && nu <= 8 && (2 * (size_t)NST * PANEL + 3 * (size_t)nu * n) * sizeof(double) <= 225 * 1024;
}

int riccati_sweep_big_device(int S, int n, int nu, int N, const double* At, const double* D, long long stride_t_A, long long stride_t_D,
                             long long stride_s_A, long long stride_s_D, const double* Q, long long Q_bs, const double* R, long long R_bs, double* Ku,
                             int* kexit, int* iters, double* ws, long long ws_stride, cudaStream_t stream) {
    const size_t smem = (2 * (size_t)NST * PANEL + 3 * (size_t)nu * n) * sizeof(double);
    const int T = (n + 7) / 8, nblk = (T + BT - 1) / BT, tb = std::max(8, (T + nblk - 1) / nblk);     // least padding with blocks of at most BT tiles
    switch (tb) {
#define CCB_BIG(TB_)                                                                                                                              \
    case TB_:                                                                                                                                     \
        CCB_CUDA(cudaFuncSetAttribute(riccati_sweep_big_kernel<TB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        riccati_sweep_big_kernel<TB_><<<S, NTB, smem, stream>>>(n, nu, N, At, D, stride_t_A, stride_t_D, stride_s_A, stride_s_D, Q, Q_bs, R, R_bs, \
                                                                Ku, kexit, iters, ws, ws_stride);                                                 \
        break;
        CCB_BIG(8) CCB_BIG(9) CCB_BIG(10) CCB_BIG(11) CCB_BIG(12) CCB_BIG(13) CCB_BIG(14) CCB_BIG(15) CCB_BIG(16)
#undef CCB_BIG
        default: set_error("riccati (blocked): unsupported block shape"); return CCB_E_UNSUPPORTED;
    }
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
