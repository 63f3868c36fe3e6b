// csrc/ccb_riccati_dmma.cu -- K4: the constrained Riccati sweep with FP64 tensor-core tiles (mma.sync.m8n8k4.f64 = SASS DMMA).
//
// Same recursion as K3 (ccb_riccati.cu; reference src/control/lqr.jl:141-184 and src/control/lqr_tracking.jl:87-115) on the
// constraint-projected pair (Ã, D):
//     Ku = (R + D'P D)^-1 D'P Ã ,   Abar = Ã - D Ku ,   P+ = Q + Ku'R Ku + Abar'P Abar ,   break if ‖P - P+‖_F < 1e-5
// One CTA per system, T = ceil(n/8) warps, warp w owns the 8-row strip w of every n x n product.  P (later W = P Abar) and
// Abar stay in shared memory for the whole sweep (2 n^2 doubles: 113 KB for the Sawyer arm, 187 KB for the 8-pendulum), the
// skinny nu-row operands (D'P, Ku, R Ku, D) next to them; Ã, D, Q are read from global memory (L2) in fragment layout and a
// copy of P lives in global memory for the ‖P - P+‖ test.  nu is padded to 8 (one tile row).  Leading dimensions are chosen
// so that every fragment load (8 rows x 4 doubles, or 4 rows x 8 doubles) is bank-conflict free: ld mod 16 in {4, 12}.
//
// Fragment layout of mma.sync.aligned.m8n8k4.row.col.f64 (g = lane >> 2, t = lane & 3):
//     a = A[g][t]      b = B[t][g]      c0,c1 = C[g][2t], C[g][2t+1]
#include "ccb_internal.h"

namespace ccb {

namespace {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__host__ __device__ inline int dmma_ld(int n) { return (n % 16 == 4 || n % 16 == 12) ? n : n + 4; }

// NU1: scalar input (nu == 1): the nu x nu solve is a division and the generic elimination is compiled out (it costs ~128 registers).
// STAGE: 3 = Ã in a third shared-memory matrix and the previous P of the tiles a warp forms in registers (72 <= n <= 84, time invariant);
// (time-invariant problems) 2 = Ã, Q and the previous P also live in shared memory, so an iteration touches global memory only
// to store the gain row (small n); 1 = the previous P is kept in a third shared-memory matrix instead of a global round trip (72 <= n <= 84:
// three n x n matrices fit in 227 KB); 0 = Ã, Q stream from L2 and the previous P goes through the global copy.
#ifndef K4_MINB_NUM
#define K4_MINB_NUM 24   // resident warps per SM the small-problem variants are compiled for (registers = 64 K / (32 * this))
#endif
template <int T, bool NU1, int STAGE>
__global__ void __launch_bounds__(32 * T, (NU1 && T <= 8) ? (K4_MINB_NUM / T > 0 ? K4_MINB_NUM / T : 1) : 1) riccati_sweep_dmma_kernel(int n, int nu, int N, const double* __restrict__ At, const double* __restrict__ D,
                                                                     long long stride_t_A, long long stride_t_D, long long stride_s_A,
                                                                     long long stride_s_D, const double* __restrict__ Q, long long Q_bs,
                                                                     const double* __restrict__ R, long long R_bs, double* __restrict__ Ku,
                                                                     int* __restrict__ kexit, int* __restrict__ iters, double* __restrict__ ws,
                                                                     long long ws_stride, int k_hi, int k_lo, int resume, const PeerTables peers) {
    // k_hi .. k_lo: the part of the sweep this launch runs (k_hi = N - 1, k_lo = 1: all of it).  resume != 0: continue a sweep whose
    // earlier part ran in a previous launch -- P comes from the global copy, kexit[s] != 0 means the system has already left the loop.
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = 32 * T;
    constexpr bool A_SM = STAGE >= 2, Q_SM = STAGE == 2, POLD_REG = STAGE >= 2;   // what a staged variant keeps on chip
    const int LD = dmma_ld(n);
    double* Ps = sm;                          // [n][LD]  P, then W = P Abar
    double* As = Ps + (size_t)n * LD;         // [n][LD]  Abar
    constexpr int NR = NU1 ? 1 : 8;           // rows of the skinny operands (nu padded to one 8-row tile unless the input is scalar)
    double* DtP = As + (size_t)n * LD;        // [NR][LD]  D'P
    double* Kk = DtP + NR * LD;               // [NR][LD]  rhs, then Ku (rows >= nu are zero)
    double* RK = Kk + NR * LD;                // [NR][LD]  R Ku
    double* Dp = RK + NR * LD;                // [n][8]   D padded to 8 columns
    double* Sc = Dp + (size_t)n * 8;          // [8][8]   R + D'P D
    double* Rs = Sc + 64;                     // [8][8]   R padded with an identity
    double* red = Rs + 64;                    // [T] partial sums (padded to 16)
    double* Asm = red + 16;                   // STAGE >= 2: [n][LD] Ã ; STAGE == 1: [n][LD] previous P
    double* Qsm = Asm + (STAGE ? (size_t)n * LD : 0);   // STAGE == 2: [n][LD] Q
    double* Pold = Qsm + (STAGE == 2 ? (size_t)n * LD : 0);  // (unused since the previous P of a staged problem is kept in registers)
    __shared__ int brk;

    const size_t s = blockIdx.x;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const size_t nn = (size_t)n * n;
    double* Pg = ws + s * ws_stride;          // global copy of P
    const double* Qs = Q + s * Q_bs;
    const double* Rg = R + s * R_bs;
    double* Kout = Ku + s * (size_t)(N - 1) * nu * n;
    const size_t kbase = (size_t)(peers.s_offset + (long long)s) * (size_t)(N - 1) * nu * n;      // my system's slot in the peer tables
    // Peer tables (multi-GPU): the gain rows go to the local table as they are produced and are pushed to every peer table in
    // contiguous runs of kPeerFlush rows (rows of consecutive steps are adjacent in [s][k][nu][n]): a few hundred KB of streaming 16-byte
    // stores per peer every kPeerFlush iterations, hidden under the recursion.  (Row-by-row remote stores were 20x slower on 8 GPUs:
    // 128 CTAs x 8 tables of ~10 GB touch more 2 MB pages than the TLB holds.)
    constexpr int kPeerFlush = 64;
    int flush_hi = k_hi;                                      // rows k-1 for k in (.., flush_hi] are not pushed yet
    // The CTAs of a launch iterate in step, and so do the launches of the other GPUs: with one flush phase and one peer order for everybody the
    // traffic would come in bursts of gridDim.x runs aimed at ONE table at a time.  Each CTA therefore starts with a first run of its own length
    // (its flushes then fall on different iterations than its neighbours') and walks the tables from its own starting peer.
    int flush_len = (peers.n > 0 && peers.spread) ? 1 + (int)((s * 29) % kPeerFlush) : kPeerFlush;
    const int peer0 = (peers.n > 0 && peers.spread) ? (int)((s + (size_t)(peers.s_offset / gridDim.x) + 1) % peers.n) : 0;
    auto peer_flush = [&](int k_from) {                       // push rows (k_from - 1) .. (flush_hi - 1); all threads, after a barrier
        if (peers.n == 0 || flush_hi < k_from) return;
        const size_t off = (size_t)(k_from - 1) * nu * n, cnt2 = (size_t)(flush_hi - k_from + 1) * nu * n / 2;
        const double2* src = reinterpret_cast<const double2*>(Kout + off);
        for (int pi = 0; pi < peers.n; pi++) {
            const int pt = (pi + peer0) % peers.n;
            double2* dst = reinterpret_cast<double2*>(peers.t[pt] + kbase + off);
            for (size_t e = tid; e < cnt2; e += NT) dst[e] = src[e];
        }
        flush_hi = k_from - 1;
    };
    const int i0 = 8 * w;                                     // my strip
    const int ra = min(i0 + g, n - 1);                        // clamped row of my A fragments / C fragments
    const bool row_ok = i0 + g < n;
    int cbj[T];                                               // clamped column of my B fragment in tile column j
    bool ok0[T], ok1[T];                                      // which of my two C elements of tile column j exist
#pragma unroll
    for (int j = 0; j < T; j++) { cbj[j] = min(8 * j + g, n - 1); ok0[j] = row_ok && 8 * j + 2 * t4 < n; ok1[j] = row_ok && 8 * j + 2 * t4 + 1 < n; }
    // P+ is symmetric: strip w only forms the tiles (w, (w + d) mod T), d = 0 .. T/2 (every unordered pair of strips once) and mirrors them
    constexpr int H = T / 2 + 1;
    int jd[H], cbd[H];
    bool use_d[H], okd0[H], okd1[H];
#pragma unroll
    for (int d = 0; d < H; d++) {
        jd[d] = (w + d >= T) ? w + d - T : w + d;
        use_d[d] = (T % 2 == 1) || d < T / 2 || w < T / 2;        // even T: the distance-T/2 pairs would be formed twice
        cbd[d] = min(8 * jd[d] + g, n - 1);
        okd0[d] = use_d[d] && row_ok && 8 * jd[d] + 2 * t4 < n;
        okd1[d] = use_d[d] && row_ok && 8 * jd[d] + 2 * t4 + 1 < n;
    }
    double* ps_c = Ps + (size_t)ra * LD + 2 * t4;             // my C elements of tile column j are at [8 j], [8 j + 1]
    double* as_c = As + (size_t)ra * LD + 2 * t4;
    double* pg_c = Pg + (size_t)ra * n + 2 * t4;
    const double* q_c = Q_SM ? Qsm + (size_t)ra * LD + 2 * t4 : Qs + (size_t)ra * n + 2 * t4;
    // previous P for the ‖P - P+‖ test: registers (STAGE 2), shared memory (STAGE 1: 72 <= n <= 84) or the global copy (STAGE 0)
    double* po_c = STAGE == 1 ? Asm + (size_t)ra * LD + 2 * t4 : pg_c;
    (void)Pold;

    if (resume && kexit[s] != 0) return;                     // broke out of the loop in an earlier launch (block-uniform)
    if (resume) {
        for (size_t e = tid; e < nn; e += NT) { const int i = (int)(e / n), j = (int)(e % n); Ps[(size_t)i * LD + j] = Pg[e]; }
    } else {
        for (size_t e = tid; e < nn; e += NT) { const int i = (int)(e / n), j = (int)(e % n); const double v = Qs[e]; Ps[(size_t)i * LD + j] = v; Pg[e] = v; }   // lqr.jl:147
    }
    if (A_SM) {
        const double* A0 = At + s * stride_s_A;
        for (size_t e = tid; e < nn; e += NT) { const int i = (int)(e / n), j = (int)(e % n); Asm[(size_t)i * LD + j] = A0[e]; if (Q_SM) Qsm[(size_t)i * LD + j] = Qs[e]; }
    }
    for (int e = tid; e < 64; e += NT) { const int i = e >> 3, j = e & 7; Rs[e] = (i < nu && j < nu) ? Rg[i * nu + j] : (i == j ? 1.0 : 0.0); }
    for (int e = tid; e < NR * LD; e += NT) { DtP[e] = 0.0; Kk[e] = 0.0; RK[e] = 0.0; }
    __syncthreads();

    int k, it = 0;
    for (k = k_hi; k >= k_lo; k--) {                          // lqr.jl:150
        it++;
        const double* Ak = At + s * stride_s_A + (size_t)(k - 1) * stride_t_A;
        const double* Dk = D + s * stride_s_D + (size_t)(k - 1) * stride_t_D;
        if (stride_t_A != 0 && k > 1) {   // time-varying (TrackingLQR): pull the next step's Ã, D into L2 while this step computes
            const char* pa = reinterpret_cast<const char*>(Ak - stride_t_A);
            for (size_t off = (size_t)tid * 128; off < nn * sizeof(double); off += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + off));
            const char* pd = reinterpret_cast<const char*>(Dk - stride_t_D);
            for (size_t off = (size_t)tid * 128; off < (size_t)n * nu * sizeof(double); off += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pd + off));
        }
        if (A_SM && stride_t_A != 0) {
            // time-varying Ã (TrackingLQR) on the staged variant: this step's Ã_k streams into the third shared-memory matrix with cp.async
            // while phase A runs, instead of being read fragment by fragment from L2 in phases B and D (r2 ncu of the time-varying sweep:
            // long_scoreboard 2.0 stall cycles per issue).  Its last readers (phase D of the previous step) are behind a barrier.
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(Asm);
            for (int e = tid; e < (n * n) / 2; e += NT) {
                const int i = (2 * e) / n, j = 2 * e - i * n;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (unsigned)((i * LD + j) * sizeof(double))), "l"(Ak + (size_t)i * n + j));
            }
            asm volatile("cp.async.commit_group;");
        }
        if (it == 1 || stride_t_D != 0) {
            for (int e = tid; e < n * 8; e += NT) { const int i = e >> 3, j = e & 7; Dp[e] = j < nu ? Dk[(size_t)i * nu + j] : 0.0; }
            __syncthreads();
        }
        if (NU1) {
            // scalar input: D'P, D'P Ã and R + D'P D are vector-matrix products -- plain FMAs (a padded 8-row DMMA tile would do 8x the work)
            for (int j = tid; j < n; j += NT) {                       // (A) DtP[j] = sum_k D[k] P[k][j]
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                const double* pc = Ps + j;
                for (int k0 = 0; k0 < n; k0 += 4) {
                    a0 = fma(Dp[8 * k0], pc[(size_t)k0 * LD], a0); a1 = fma(Dp[8 * (k0 + 1)], pc[(size_t)(k0 + 1) * LD], a1);
                    a2 = fma(Dp[8 * (k0 + 2)], pc[(size_t)(k0 + 2) * LD], a2); a3 = fma(Dp[8 * (k0 + 3)], pc[(size_t)(k0 + 3) * LD], a3);
                }
                DtP[j] = (a0 + a1) + (a2 + a3);
            }
            if (A_SM && stride_t_A != 0) asm volatile("cp.async.wait_all;");
            __syncthreads();
            const size_t lda = A_SM ? (size_t)LD : (size_t)n;
            const double* Ab = A_SM ? Asm : Ak;
            for (int j = tid; j < n; j += NT) {                       // (B) rhs[j] = sum_k DtP[k] Ã[k][j]
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                const double* ac = Ab + j;
                for (int k0 = 0; k0 < n; k0 += 4) {
                    a0 = fma(DtP[k0], ac[(size_t)k0 * lda], a0); a1 = fma(DtP[k0 + 1], ac[(size_t)(k0 + 1) * lda], a1);
                    a2 = fma(DtP[k0 + 2], ac[(size_t)(k0 + 2) * lda], a2); a3 = fma(DtP[k0 + 3], ac[(size_t)(k0 + 3) * lda], a3);
                }
                Kk[j] = (a0 + a1) + (a2 + a3);
            }
            if (w == 0) {                                             // Sc = R + sum_k DtP[k] D[k]  (one warp, fixed-order reduction)
                double a0 = 0.0;
                for (int k0 = lane; k0 < n; k0 += 32) a0 = fma(DtP[k0], Dp[8 * k0], a0);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                if (lane == 0) Sc[0] = Rs[0] + a0;
            }
            __syncthreads();
        } else {
        // ---- (A) D'P: tile column w.  a = D^T fragment, b = P fragment
            {
                const int j0 = 8 * w, cb = min(j0 + g, n - 1);
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                for (int k0 = 0; k0 < n; k0 += 8) {
                    dmma(c0, c1, Dp[(k0 + t4) * 8 + g], Ps[(size_t)(k0 + t4) * LD + cb]);
                    if (k0 + 4 < n) dmma(e0, e1, Dp[(k0 + 4 + t4) * 8 + g], Ps[(size_t)(k0 + 4 + t4) * LD + cb]);
                }
                const int col = j0 + 2 * t4;
                if (col < n) DtP[g * LD + col] = c0 + e0;
                if (col + 1 < n) DtP[g * LD + col + 1] = c1 + e1;
            }
            if (A_SM && stride_t_A != 0) asm volatile("cp.async.wait_all;");
            __syncthreads();
            // ---- (B) rhs = D'P Ã (tile column w; Ã from global) and Sc = R + D'P D (warp 0)
            {
                const int j0 = 8 * w, cb = min(j0 + g, n - 1);
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                const size_t lda = A_SM ? (size_t)LD : (size_t)n;
                const double* ag = (A_SM ? Asm : Ak) + (size_t)t4 * lda + cb;
                const double* dr = DtP + g * LD + t4;
                int k0 = 0;
                for (; k0 + 16 <= n; k0 += 16) {                              // four independent loads in flight
                    const double b0 = ag[0], b1 = ag[4 * lda], b2 = ag[8 * lda], b3 = ag[12 * lda];
                    dmma(c0, c1, dr[k0], b0); dmma(e0, e1, dr[k0 + 4], b1); dmma(c0, c1, dr[k0 + 8], b2); dmma(e0, e1, dr[k0 + 12], b3);
                    ag += 16 * lda;
                }
                for (; k0 < n; k0 += 4) { dmma(c0, c1, dr[k0], ag[0]); ag += 4 * lda; }
                const int col = j0 + 2 * t4;
                if (col < n) Kk[g * LD + col] = c0 + e0;
                if (col + 1 < n) Kk[g * LD + col + 1] = c1 + e1;
                if (w == 0) {
                    double s0 = Rs[g * 8 + 2 * t4], s1 = Rs[g * 8 + 2 * t4 + 1];
                    for (int k0 = 0; k0 < n; k0 += 4) dmma(s0, s1, DtP[g * LD + k0 + t4], Dp[(k0 + t4) * 8 + g]);
                    Sc[g * 8 + 2 * t4] = s0; Sc[g * 8 + 2 * t4 + 1] = s1;
                }
            }
            __syncthreads();
        }
        // ---- (C) Ku = Sc^-1 rhs
        if (NU1) {                                                       // scalar input: Ku = rhs / Sc
            const double inv = 1.0 / Sc[0], r00 = Rs[0];
            for (int c = tid; c < n; c += NT) {
                const double x0 = Kk[c] * inv;
                Kk[c] = x0;
                Kout[(size_t)(k - 1) * n + c] = x0;                          // lqr.jl:162-164
                RK[c] = r00 * x0;
            }
        } else {
            // Sc = R + D'P D is symmetric positive definite: elimination without pivoting, redundantly per thread in registers
            for (int c = tid; c < n; c += NT) {
                double Mx[8][8], x[8], invd[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    x[i] = Kk[i * LD + c];
#pragma unroll
                    for (int j = 0; j < 8; j++) Mx[i][j] = Sc[i * 8 + j];
                }
#pragma unroll
                for (int cc = 0; cc < 8; cc++) {
                    invd[cc] = 1.0 / Mx[cc][cc];
#pragma unroll
                    for (int i = cc + 1; i < 8; i++) {
                        const double l = Mx[i][cc] * invd[cc];
#pragma unroll
                        for (int j = cc + 1; j < 8; j++) Mx[i][j] = fma(-l, Mx[cc][j], Mx[i][j]);
                        x[i] = fma(-l, x[cc], x[i]);
                    }
                }
#pragma unroll
                for (int i = 7; i >= 0; i--) {
#pragma unroll
                    for (int j = i + 1; j < 8; j++) x[i] = fma(-Mx[i][j], x[j], x[i]);
                    x[i] *= invd[i];
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    Kk[i * LD + c] = x[i];
                    if (i < nu) Kout[(size_t)(k - 1) * nu * n + (size_t)i * n + c] = x[i];      // lqr.jl:162-164
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; j++) acc = fma(Rs[i * 8 + j], x[j], acc);
                    RK[i * LD + c] = i < nu ? acc : 0.0;
                }
            }
        }
        __syncthreads();
        // ---- (D) Abar = Ã - D Ku  (strip w; rank-8 update of Ã read from global in C layout)          lqr.jl:169
        double acc[T][2];
        {
            const double* a_c = A_SM ? Asm + (size_t)ra * LD + 2 * t4 : Ak + (size_t)ra * n + 2 * t4;
#pragma unroll
            for (int j = 0; j < T; j++) { acc[j][0] = ok0[j] ? a_c[8 * j] : 0.0; acc[j][1] = ok1[j] ? a_c[8 * j + 1] : 0.0; }
            if (NU1) {
                const double dneg = -Dp[ra * 8];
                const double* kc = Kk + 2 * t4;
#pragma unroll
                for (int j = 0; j < T; j++) {
                    if (ok0[j]) as_c[8 * j] = fma(dneg, kc[8 * j], acc[j][0]);
                    if (ok1[j]) as_c[8 * j + 1] = fma(dneg, kc[8 * j + 1], acc[j][1]);
                }
            } else {
                const double a0 = -Dp[ra * 8 + t4], a1 = -Dp[ra * 8 + 4 + t4];
                const double *k0row = Kk + t4 * LD, *k1row = Kk + (4 + t4) * LD;
#pragma unroll
                for (int j = 0; j < T; j++) {
                    dmma(acc[j][0], acc[j][1], a0, k0row[cbj[j]]);
                    dmma(acc[j][0], acc[j][1], a1, k1row[cbj[j]]);
                    if (ok0[j]) as_c[8 * j] = acc[j][0];
                    if (ok1[j]) as_c[8 * j + 1] = acc[j][1];
                }
            }
        }
        __syncthreads();
        // ---- (E) W = P Abar (strip w), kept in registers until every warp has finished reading P
#pragma unroll
        for (int j = 0; j < T; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
        {
            const double* arow = Ps + (size_t)ra * LD + t4;
            const double* brow = As + (size_t)t4 * LD;
            for (int k0 = 0; k0 < n; k0 += 4) {
                const double a = arow[k0];
#pragma unroll
                for (int j = 0; j < T; j++) dmma(acc[j][0], acc[j][1], a, brow[cbj[j]]);
                brow += 4 * LD;
            }
        }
        __syncthreads();
        double pold[POLD_REG ? H : 1][2];
        if constexpr (POLD_REG) {                                      // small staged problems: previous P of my tiles stays in registers
#pragma unroll
            for (int d = 0; d < H; d++) { pold[d][0] = ps_c[8 * jd[d]]; pold[d][1] = ps_c[8 * jd[d] + 1]; }
        }
#pragma unroll
        for (int j = 0; j < T; j++) {                                    // old P -> shared / global copy (for the ‖P - P+‖ test), W -> shared
            if constexpr (POLD_REG) {
                if (ok0[j]) ps_c[8 * j] = acc[j][0];
                if (ok1[j]) ps_c[8 * j + 1] = acc[j][1];
            } else {
                if (ok0[j]) { po_c[8 * j] = ps_c[8 * j]; ps_c[8 * j] = acc[j][0]; }
                if (ok1[j]) { po_c[8 * j + 1] = ps_c[8 * j + 1]; ps_c[8 * j + 1] = acc[j][1]; }
            }
        }
        __syncthreads();
        // ---- (F) P+ = Q + Abar' W + Ku' (R Ku): tiles (w, jd[d]) only, then mirrored                    lqr.jl:170
        double accF[H][2];
#pragma unroll
        for (int d = 0; d < H; d++) { accF[d][0] = okd0[d] ? q_c[8 * jd[d]] : 0.0; accF[d][1] = okd1[d] ? q_c[8 * jd[d] + 1] : 0.0; }
        {
            const double* arow = As + (size_t)t4 * LD + ra;              // Abar^T fragment
            const double* brow = Ps + (size_t)t4 * LD;                   // W
            for (int k0 = 0; k0 < n; k0 += 4) {
                const double a = *arow;
#pragma unroll
                for (int d = 0; d < H; d++)
                    if (use_d[d]) dmma(accF[d][0], accF[d][1], a, brow[cbd[d]]);
                arow += 4 * LD; brow += 4 * LD;
            }
        }
        if (NU1) {                                                       // + Ku' (R Ku): rank one
            const double ki = Kk[ra];
            const double* rc = RK + 2 * t4;
#pragma unroll
            for (int d = 0; d < H; d++) {
                if (okd0[d]) accF[d][0] = fma(ki, rc[8 * jd[d]], accF[d][0]);
                if (okd1[d]) accF[d][1] = fma(ki, rc[8 * jd[d] + 1], accF[d][1]);
            }
        } else {
#pragma unroll
            for (int k0 = 0; k0 < 8; k0 += 4) {
                const double a = Kk[(k0 + t4) * LD + ra];                // Ku^T fragment
                const double* brow = RK + (k0 + t4) * LD;
#pragma unroll
                for (int d = 0; d < H; d++)
                    if (use_d[d]) dmma(accF[d][0], accF[d][1], a, brow[cbd[d]]);
            }
        }
        {   // the diagonal tile is averaged with its own transpose so that P stays EXACTLY symmetric: rounding-level asymmetry confined to
            // the diagonal tiles is amplified by this recursion on stiff problems (8-pendulum, N = 300: gains off by O(1))
            const int src0 = (2 * t4) * 4 + (g >> 1), src1 = (2 * t4 + 1) * 4 + (g >> 1);
            const double u00 = __shfl_sync(0xffffffffu, accF[0][0], src0), u01 = __shfl_sync(0xffffffffu, accF[0][1], src0);
            const double u10 = __shfl_sync(0xffffffffu, accF[0][0], src1), u11 = __shfl_sync(0xffffffffu, accF[0][1], src1);
            accF[0][0] = 0.5 * (accF[0][0] + ((g & 1) ? u01 : u00));
            accF[0][1] = 0.5 * (accF[0][1] + ((g & 1) ? u11 : u10));
        }
        double nd = 0.0;                                                 // ‖P - P+‖_F^2: off-diagonal tiles stand for their mirror image too
#pragma unroll
        for (int d = 0; d < H; d++) {
            double p0, p1;
            if constexpr (POLD_REG) { p0 = pold[d][0]; p1 = pold[d][1]; }
            else { p0 = okd0[d] ? po_c[8 * jd[d]] : 0.0; p1 = okd1[d] ? po_c[8 * jd[d] + 1] : 0.0; }
            const double d0 = okd0[d] ? p0 - accF[d][0] : 0.0, d1 = okd1[d] ? p1 - accF[d][1] : 0.0;
            const double wgt = d == 0 ? 1.0 : 2.0;
            nd = fma(wgt * d0, d0, nd); nd = fma(wgt * d1, d1, nd);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, o);
        if (lane == 0) red[w] = nd;
        __syncthreads();                                                 // every warp is done reading W
        if (tid == 0) {
            double tot = 0.0;
            for (int q = 0; q < T; q++) tot += red[q];
            brk = sqrt(tot) < 1e-5;                                      // lqr.jl:172
        }
#pragma unroll
        for (int d = 0; d < H; d++) {                                    // lqr.jl:176  Pk = Pkp1: my tiles and their transposes
            const int c0 = 8 * jd[d] + 2 * t4;
            if (okd0[d]) { ps_c[8 * jd[d]] = accF[d][0]; if (d > 0) Ps[(size_t)c0 * LD + ra] = accF[d][0]; }
            if (okd1[d]) { ps_c[8 * jd[d] + 1] = accF[d][1]; if (d > 0) Ps[(size_t)(c0 + 1) * LD + ra] = accF[d][1]; }
        }
        __syncthreads();
        if (brk) break;
        if (flush_hi - k + 1 >= flush_len) { peer_flush(k); flush_len = kPeerFlush; }
    }
    if (k < k_lo && k_lo > 1) {
        peer_flush(k_lo);                                          // this launch's part is done, the sweep is not: park P for the next launch
        for (size_t e = tid; e < nn; e += NT) { const int i = (int)(e / n), j = (int)(e % n); Pg[e] = Ps[(size_t)i * LD + j]; }
        if (tid == 0) { kexit[s] = 0; iters[s] = (resume ? iters[s] : 0) + it; }
        return;
    }
    if (k < 1) k = 1;                                                    // a completed loop leaves k == 1
    if (peers.n > 0 && peers.defer_backfill) {
        // multi-GPU exchange with deferred back-fill: only the rows the recursion produced cross NVLink; every GPU replicates row k - 1 below the
        // break index in its OWN table after the exchange barrier (ccb_gains_backfill) -- for the Sawyer arm at N = 2000 that is a fifth of the
        // table, and it used to leave all the CTAs of all the GPUs in one burst at the end of the sweep, with nothing left to hide it under
        peer_flush(k);
        if (tid == 0) { kexit[s] = k; iters[s] = (resume ? iters[s] : 0) + it; }
        return;
    }
    for (int k2 = k - 1; k2 >= 1; k2--) {                                // lqr.jl:179-181 back-fill
        for (int e = tid; e < nu * n; e += NT) Kout[(size_t)(k2 - 1) * nu * n + e] = Kout[(size_t)k2 * nu * n + e];
        __syncthreads();
    }
    peer_flush(1);                                                       // whatever is left, the back-filled rows included
    if (tid == 0) { kexit[s] = k; iters[s] = (resume ? iters[s] : 0) + it; }
}

size_t dmma_smem_bytes(int n, int T, int stage, bool nu1 = false) {
    const int LD = dmma_ld(n);
    return ((size_t)(stage == 2 ? 4 : stage ? 3 : 2) * n * LD + 3 * (nu1 ? 1 : 8) * LD + (size_t)n * 8 + 128 + 16 + 8) * sizeof(double);
}

template <int T, bool NU1, int STAGE>
int launch_dmma(int S, int n, int nu, int N, const double* At, const double* D, long long stride_t_A, long long stride_t_D, long long stride_s_A,
                long long stride_s_D, const double* Q, long long Q_bs, const double* R, long long R_bs, double* Ku, int* kexit, int* iters, double* ws,
                long long ws_stride, cudaStream_t stream, int k_hi, int k_lo, int resume) {
    const size_t smem = dmma_smem_bytes(n, T, STAGE, NU1);
    CCB_CUDA(cudaFuncSetAttribute(riccati_sweep_dmma_kernel<T, NU1, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    riccati_sweep_dmma_kernel<T, NU1, STAGE><<<S, 32 * T, smem, stream>>>(n, nu, N, At, D, stride_t_A, stride_t_D, stride_s_A, stride_s_D, Q, Q_bs, R, R_bs, Ku, kexit,
                                                             iters, ws, ws_stride, k_hi, k_lo, resume, g_peer_tables);
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace

// n must be a multiple of 4 (it is 12 nb), nu <= 8, and 2 n^2 doubles must fit in shared memory (n <= 108)
bool riccati_dmma_eligible(int n, int nu) {
    if (n % 4 != 0 || nu > 8 || n < 8) return false;
    const int T = (n + 7) / 8;
    return T <= 14 && dmma_smem_bytes(n, T, 0) <= 227 * 1024;
}

int riccati_sweep_dmma_device(int S, int n, int nu, int N, const double* At, const double* D, long long stride_t_A, long long stride_t_D,
                              long long stride_s_A, long long stride_s_D, const double* Q, long long Q_bs, const double* R, long long R_bs, double* Ku,
                              int* kexit, int* iters, double* ws, long long ws_stride, cudaStream_t stream, int k_hi, int k_lo, int resume) {
    const int T = (n + 7) / 8;
    if (k_hi <= 0) { k_hi = N - 1; k_lo = 1; resume = 0; }
    const bool stage = stride_t_A == 0 && n <= 36;       // small time-invariant problems: Ã, Q, previous P in shared memory as well
    const bool stage_a = !stage && n >= 72 && dmma_smem_bytes(n, T, 1) <= 227 * 1024;
    const bool stage_b = stage_a && (stride_t_A == 0 || !getenv("CCB_K4_TV_STAGE1"));     // ... and, when Ã does not change, Ã in that third matrix with the previous P of my tiles in registers   // 72 <= n <= 84: previous P in shared memory (for smaller n the extra shared memory costs more residency than it saves)
#define CCB_DMMA_ARGS S, n, nu, N, At, D, stride_t_A, stride_t_D, stride_s_A, stride_s_D, Q, Q_bs, R, R_bs, Ku, kexit, iters, ws, ws_stride, stream, k_hi, k_lo, resume
#define CCB_DMMA_CASE(TT)                                                                      \
    case TT:                                                                                   \
        if constexpr (TT <= 5) { if (stage) return nu == 1 ? launch_dmma<TT, true, 2>(CCB_DMMA_ARGS) : launch_dmma<TT, false, 2>(CCB_DMMA_ARGS); } \
        if constexpr (TT >= 9 && TT <= 11) { if (stage_b) return nu == 1 ? launch_dmma<TT, true, 3>(CCB_DMMA_ARGS) : launch_dmma<TT, false, 3>(CCB_DMMA_ARGS); } \
        if constexpr (TT >= 9 && TT <= 11) { if (stage_a) return nu == 1 ? launch_dmma<TT, true, 1>(CCB_DMMA_ARGS) : launch_dmma<TT, false, 1>(CCB_DMMA_ARGS); } \
        return nu == 1 ? launch_dmma<TT, true, 0>(CCB_DMMA_ARGS) : launch_dmma<TT, false, 0>(CCB_DMMA_ARGS);
    switch (T) {
        CCB_DMMA_CASE(1) CCB_DMMA_CASE(2) CCB_DMMA_CASE(3) CCB_DMMA_CASE(4) CCB_DMMA_CASE(5) CCB_DMMA_CASE(6) CCB_DMMA_CASE(7)
        CCB_DMMA_CASE(8) CCB_DMMA_CASE(9) CCB_DMMA_CASE(10) CCB_DMMA_CASE(11) CCB_DMMA_CASE(12) CCB_DMMA_CASE(13) CCB_DMMA_CASE(14)
    }
#undef CCB_DMMA_CASE
#undef CCB_DMMA_ARGS
    set_error("riccati_dmma: unsupported size");
    return CCB_E_UNSUPPORTED;
}

}  // namespace ccb
