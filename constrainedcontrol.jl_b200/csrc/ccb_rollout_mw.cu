// csrc/ccb_rollout_mw.cu -- K1m: the row-per-lane rollout kernel for LONG serial chains (Sawyer arm: 7 joints = 35 rows,
// 8-pendulum: 45 rows, 32-pendulum: 165 rows), one CTA of ceil(5 nj / 32) warps per system.
//
// Same decomposition as K1r (ccb_rollout_rows.cu): one lane per scalar constraint row, 5 lanes per joint in chain order,
// the joint's child body redundantly in its 5 lanes, the block-tridiagonal constraint Schur complement S = Gv D^-1 G^T with
// one row per lane.  A system no longer fits a warp, so the neighbour exchanges that K1r does with warp shuffles go through
// small shared-memory mailboxes separated by CTA barriers:
//   * poses / residuals / 3x3 blocks of the parent body, multipliers of the own and the child joint, force sums;
//   * the pivot row of the banded elimination (double buffered: ONE barrier per pivot).  Pivots run from the leaf row to
//     the root row (ldu_factorization! role); rows of the pivot's own block are eliminated above AND below the pivot
//     (Gauss-Jordan inside the 5x5 block), so the substitution root -> leaf (ldu_backsubstitution! role) needs one mailbox
//     round per JOINT instead of one per row.
// Semantics identical to K1 / K1c / K1r: simulate! loop of ConstrainedDynamics.jl with the control
// laws of reference src/control/lqr.jl:89-139, lqr_tracking.jl:46-71, pid.jl:43-88 fused.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include "ccb_dev_math.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

constexpr int kGS = 62;   // doubles per joint in the knot-row table: 5 rows x 12 (+2)
constexpr int kCS = 24;   // constants per joint: p1 3, p2 3, qoff 4, axis 3, J 9, mass 1 (+1)


__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// NJT: number of joints as a compile-time constant (0 = run time): the shared-memory mailbox offsets fold into immediates
// MAXR: register cap of the variant.  The kernel is latency bound, so what matters is how the batch quantises into waves of resident
// CTAs: 255 registers (no spills, 4 CTAs of 2 warps per SM) when the batch fits in one such wave or needs many, fewer registers (more
// resident CTAs, some spills) when that turns two waves into one or lifts a long chain above ~8 warps per SM (launch_rollout_mw).
template <int NJT, int MAXR>
__global__ void __maxnreg__(MAXR) rollout_mw_kernel(const DevMech M, const DevCtrl C, const SimArgs A, const int* __restrict__ depth) {
    extern __shared__ __align__(16) double smem[];
    const int nj = NJT ? NJT : M.nj;
    double* gk = smem;                          // [nj][kGS]  knot rows: row t at t*12: GkA[6] | GkB[6]
    double* cst = gk + nj * kGS;                // [nj][kCS]
    double* xpose = cst + nj * kCS;             // [nj][8]    body pose mailbox (knot pose, then x3 q3)
    double* xd = xpose + nj * 8;                // [nj][6]    body residual d
    double* xM = xd + nj * 6;                   // [nj][9]    Tw D^-1 of the body
    double* xl = xM + nj * 9;                   // [5 nj + 5] multipliers / substitution results per row (+5: reads past the leaf)
    double* xf = xl + 5 * nj + 5;               // [nj][6]    force sums
    double* pivb = xf + nj * 6;                 // [2][2][12] pivot rows of the two elimination streams, double buffered
    double* red = pivb + 48;                    // [5 nj]     per-row partials
    double* red2 = red + 5 * nj;                // [nj]       per-joint partials
    double* xu = red2 + nj;                     // [nj + 1]   joint inputs
    double* ured = xu + nj + 1;                 // [8][5 nj]  control-law partials (up to 8 inputs per round)
    double* ured2 = ured + 40 * nj;             // [8][nj]
    int* jorder = reinterpret_cast<int*>(ured2 + 8 * nj);   // [nj] joint at chain position

    const int q = threadIdx.x;
    const bool lane_ok = q < 5 * nj;
    const int pos = lane_ok ? q / 5 : 0, r = lane_ok ? q - 5 * pos : q % 5;
    const bool has_p = pos > 0, has_c = pos < nj - 1;
    const double dt = M.dt, idt = 1.0 / M.dt, c4 = 4.0 / (M.dt * M.dt);
    const int nst = 13 * M.nb;

    // ---- chain order and joint constants -> shared memory
    for (int j = q; j < nj; j += blockDim.x) jorder[__ldg(depth + j)] = j;
    __syncthreads();
    for (int i = q; i < nj * kCS; i += blockDim.x) {
        const int p = i / kCS, k = i - p * kCS, j = jorder[p], bb = __ldg(M.j_child + j);
        double v = 0.0;
        if (k < 3) v = __ldg(M.j_p1 + 3 * j + k);
        else if (k < 6) v = __ldg(M.j_p2 + 3 * j + k - 3);
        else if (k < 10) v = __ldg(M.j_qoff + 4 * j + k - 6);
        else if (k < 13) v = __ldg(M.j_axis + 3 * j + k - 10);
        else if (k < 22) v = __ldg(M.inertia + 9 * bb + k - 13);
        else if (k == 22) v = __ldg(M.mass + bb);
        cst[i] = v;
    }
    for (int i = q; i < 5 * nj + 5; i += blockDim.x) xl[i] = 0.0;
    __syncthreads();
    const int jt = jorder[pos];
    const int cjt = has_c ? jorder[pos + 1] : -1;
    const int b = __ldg(M.j_child + jt);
    const int dim = __ldg(M.j_dim + jt), r0 = __ldg(M.j_row0 + jt);
    const bool rowpad = !lane_ok || r >= dim;
    const int rr = r < dim ? r : 0;
    const int kind = __ldg(M.row_kind + r0 + rr);
    const V3 pr = ldg3(M.row_proj + 3 * (r0 + rr));
    const int act_kind = __ldg(M.eqc_act_kind + jt);
    const int cact_kind = has_c ? __ldg(M.eqc_act_kind + cjt) : 0;
    double* gk_own = gk + pos * kGS;
    const double* gk_par = gk + (has_p ? pos - 1 : pos) * kGS;
    const double* gk_chi = gk + (has_c ? pos + 1 : pos) * kGS;
    const double* cj = cst + pos * kCS;
    const double* ccj = cst + (has_c ? pos + 1 : pos) * kCS;
    const double mass = cj[22], kdt = dt / mass;
    const double pkdt = has_p ? dt / cst[(pos - 1) * kCS + 22] : 0.0;
    const int ppos = has_p ? pos - 1 : pos;

    // ---- work loop.  Plain launch (A.persist == nullptr): one pass, the CTA owns system blockIdx.x for all A.steps steps.
    // Persistent time-sliced launch (launch_rollout_mw): W = gridDim.x resident CTAs for B = q W + e systems.  The rollouts are cut into
    // A.nchunks chunks of steps with a STATIC schedule: in round c worker w simulates chunk c of its own systems w, w + W, .. (nobody else ever
    // touches them) and then, when its turn comes, chunk c of one of the e extra systems -- extra system q W + j goes to worker (c e + j) mod W, so
    // the extra work rotates over the workers and the chunk an extra item continues was stored a whole round earlier by another CTA (flag per
    // system, release / acquire through __threadfence; state reloaded with ld.cg since L1 is not coherent).  This lets a batch that does not fit
    // one wave of the spill-free 255-register variant run on it anyway: 1024 Sawyer arms on 592 CTAs.  Bit-identical to the plain launch.
    const bool persist = A.persist != nullptr;
    const int nworkers = gridDim.x, qown = persist ? A.B / nworkers : 1, extra = persist ? A.B - qown * nworkers : 0;
    double z[13], s[6], lam;
    for (int slot = 0;; slot++) {
    int sys = blockIdx.x, chunk = 0;
    if (persist) {
        chunk = slot / (qown + 1);
        if (chunk >= A.nchunks) break;
        const int i = slot - chunk * (qown + 1);
        if (i < qown) {
            sys = (int)blockIdx.x + i * nworkers;
        } else {
            int j = ((int)blockIdx.x - (int)(((long long)chunk * extra) % nworkers)) % nworkers;
            if (j < 0) j += nworkers;
            if (j >= extra) continue;                 // CTA-uniform
            sys = qown * nworkers + j;
            if (chunk > 0) {
                if (q == 0) {
                    const volatile int* flag = A.persist + 1 + sys;
                    while (*flag < chunk) __nanosleep(2000);
                }
                __syncthreads();
                __threadfence();
            }
        }
    } else if (slot > 0) {
        break;
    }
    const int step_lo = persist ? chunk * A.chunk_steps : 0;
    const int step_hi = persist ? min(A.steps, step_lo + A.chunk_steps) : A.steps;
    if (persist) {
        for (int i = q; i < 5 * nj + 5; i += blockDim.x) xl[i] = 0.0;
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 13; i++) z[i] = persist ? __ldcg(A.state + (size_t)sys * nst + 13 * b + i) : A.state[(size_t)sys * nst + 13 * b + i];
    lam = (lane_ok && r < dim) ? (persist ? __ldcg(A.lambda + (size_t)sys * M.ml + r0 + r) : A.lambda[(size_t)sys * M.ml + r0 + r]) : 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) s[i] = 0.0;

    // sum over the rows of the system in one fixed order (every thread gets the same bits): rows -> joints -> system
    auto sys_sum = [&](double v) {
        if (lane_ok) red[q] = v;
        __syncthreads();
        if (lane_ok && r == 0) red2[pos] = ((red[q] + red[q + 1]) + (red[q + 2] + red[q + 3])) + red[q + 4];
        __syncthreads();
        double acc = 0.0;
        for (int p = 0; p < nj; p++) acc += red2[p];
        return acc;
    };

    int total_iters = 0;
    bool failed = false;

    for (int step = step_lo; step < step_hi; step++) {
        const int k = A.k0 + step;
        if (lane_ok && A.storage) {
            double* out = A.storage + ((size_t)step * A.B + sys) * nst + 13 * b;
#pragma unroll
            for (int i = 0; i < 13; i++)
                if (i % 5 == r) out[i] = z[i];
        }
        // ---------------- control!(k) part 1 (PID / open-loop table) needs the parent's continuous pose
        double u = 0.0;
        if (C.kind == CCB_CTRL_PID || C.kind == CCB_CTRL_TABLE) {
            if (lane_ok && r == 0) {
#pragma unroll
                for (int i = 0; i < 7; i++) xpose[pos * 8 + i] = z[i];
            }
            __syncthreads();
            if (lane_ok && r == 0) {
                for (int i = 0; i < C.nu; i++) {
                    if (__ldg(C.eqc + i) != jt) continue;
                    if (C.kind == CCB_CTRL_TABLE) {
                        if (k >= 1 && k <= C.N) u += __ldg(C.U + (size_t)sys * C.U_bs + (size_t)(k - 1) * C.nu + i);
                        continue;
                    }
                    V3 xa = {0.0, 0.0, 0.0};
                    Q4 qa = {1.0, 0.0, 0.0, 0.0};
                    if (has_p) { xa = ld3(xpose + ppos * 8); qa = ld4(xpose + ppos * 8 + 3); }
                    const V3 xb = v3(z[0], z[1], z[2]);
                    const Q4 qb = {z[3], z[4], z[5], z[6]};
                    const V3 axis = ld3(cj + 10);
                    double coord;
                    if (act_kind == 1) {
                        coord = dot(axis, qrotT(qa, xb + qrot(qb, ld3(cj + 3)) - xa) - ld3(cj));
                    } else {
                        Q4 qe = qmul(qmul(qconj(qa), qb), qconj(ld4(cj + 6)));
                        double nv = sqrt(qe.x * qe.x + qe.y * qe.y + qe.z * qe.z);
                        coord = (nv == 0.0) ? 0.0 : 2.0 * atan2(nv, qe.s) * (axis.x * qe.x + axis.y * qe.y + axis.z * qe.z) / nv;
                    }
                    double err = __ldg(C.goal + (size_t)sys * C.goal_bs + i) - coord;
                    if (act_kind == 2) {
                        const double pi = 3.14159265358979323846;
                        if (err > pi) err -= 2.0 * pi;
                        else if (err < -pi) err += 2.0 * pi;
                    }
                    double* integ = C.pid_int + (size_t)sys * C.nu + i;
                    double* last = C.pid_last + (size_t)sys * C.nu + i;
                    if (k == 1) *last = err;
                    const double in = *integ + err * dt;
                    *integ = in;
                    const double der = (err - *last) / dt;
                    u += __ldg(C.P + i) * err + __ldg(C.I + i) * in + __ldg(C.D + i) * der;
                    *last = err;
                }
            }
            __syncthreads();
        }
        // ---------------- knot
        if (!A.knot_input) {
            const V3 v = v3(z[7], z[8], z[9]), w = v3(z[10], z[11], z[12]);
            z[0] = fma(dt, v.x, z[0]); z[1] = fma(dt, v.y, z[1]); z[2] = fma(dt, v.z, z[2]);
            const double sq = sqrt(c4 - dot(w, w));
            const Q4 q2 = qmul(Q4{z[3], z[4], z[5], z[6]}, Q4{0.5 * dt * sq, 0.5 * dt * w.x, 0.5 * dt * w.y, 0.5 * dt * w.z});
            z[3] = q2.s; z[4] = q2.x; z[5] = q2.y; z[6] = q2.z;
        }
#pragma unroll
        for (int i = 0; i < 6; i++) s[i] = z[7 + i];
        if (lane_ok && r == 0) {
#pragma unroll
            for (int i = 0; i < 7; i++) xpose[pos * 8 + i] = z[i];
        }
        __syncthreads();
        V3 xa2 = {0.0, 0.0, 0.0};
        Q4 qa2 = {1.0, 0.0, 0.0, 0.0};
        if (has_p) { xa2 = ld3(xpose + ppos * 8); qa2 = ld4(xpose + ppos * 8 + 3); }
        // ---------------- control!(k) part 2: LQR-type laws on the knot state (lqr.jl:98-108)
        {
            const int n = 12 * M.nb;
            bool lqr = false;
            const int csys = C.sys_div > 1 ? sys / C.sys_div : sys;   // several systems may share one table (ccb_ctrl_set_replication)
            const double *K = nullptr, *zd = nullptr, *Fd = nullptr;
            if (C.kind == CCB_CTRL_LQR_INF) {
                lqr = true;
                K = C.K + (size_t)csys * C.K_bs; zd = C.zd + (size_t)csys * C.zd_bs; Fd = C.Fd + (size_t)csys * C.Fd_bs;
            } else if (C.kind == CCB_CTRL_LQR_FIN) {
                lqr = k < C.N;
                K = C.K + (size_t)csys * C.K_bs + (size_t)(k - 1) * C.nu * n; zd = C.zd + (size_t)csys * C.zd_bs; Fd = C.Fd + (size_t)csys * C.Fd_bs;
            } else if (C.kind == CCB_CTRL_TRACKING) {
                lqr = k < C.N;
                K = C.K + (size_t)csys * C.K_bs + (size_t)(k - 1) * C.nu * n; zd = C.zd + (size_t)csys * C.zd_bs + (size_t)(k - 1) * 13 * M.nb;
                Fd = C.Fd + (size_t)csys * C.Fd_bs + (size_t)(k - 1) * C.nu;
            }
            if (lqr) {
                V3 dz = {0.0, 0.0, 0.0};
                if (lane_ok && r < 4) {
                    const double* d = zd + 13 * b;
                    if (r == 0) dz = v3(z[0] - __ldg(d), z[1] - __ldg(d + 1), z[2] - __ldg(d + 2));
                    else if (r == 1) dz = v3(z[7] - __ldg(d + 7), z[8] - __ldg(d + 8), z[9] - __ldg(d + 9));
                    else if (r == 2) dz = qvec(qmul(qconj(ldg4(d + 3)), Q4{z[3], z[4], z[5], z[6]}));
                    else dz = v3(z[10] - __ldg(d + 10), z[11] - __ldg(d + 11), z[12] - __ldg(d + 12));
                }
                // all inputs at once: the gain rows are fetched with independent loads and reduced with ONE pair of barriers
                // (rows -> joints -> system, fixed order), up to 8 inputs per round
                for (int i0 = 0; i0 < C.nu; i0 += 8) {
                    const int ni = min(8, C.nu - i0);
                    double part[8];
#pragma unroll
                    for (int ii = 0; ii < 8; ii++) {
                        part[ii] = 0.0;
                        if (ii < ni && lane_ok && r < 4) {
                            const double* Kr = K + (size_t)(i0 + ii) * n + 12 * b + 3 * r;
                            part[ii] = fma(__ldg(Kr), dz.x, fma(__ldg(Kr + 1), dz.y, __ldg(Kr + 2) * dz.z));
                        }
                    }
                    if (lane_ok) {
#pragma unroll
                        for (int ii = 0; ii < 8; ii++) if (ii < ni) ured[ii * 5 * nj + q] = part[ii];
                    }
                    __syncthreads();
                    if (lane_ok && r == 0) {
                        for (int ii = 0; ii < ni; ii++) {
                            const double* rp = ured + ii * 5 * nj + q;
                            ured2[ii * nj + pos] = ((rp[0] + rp[1]) + (rp[2] + rp[3])) + rp[4];
                        }
                    }
                    __syncthreads();
                    if (lane_ok && r == 0) {
                        for (int ii = 0; ii < ni; ii++) {
                            if (__ldg(C.eqc + i0 + ii) != jt) continue;
                            double tot = 0.0;
                            for (int p = 0; p < nj; p++) tot += ured2[ii * nj + p];
                            u += __ldg(Fd + i0 + ii) - tot;
                        }
                    }
                    __syncthreads();
                }
            }
        }
        if (lane_ok && r == 0) xu[pos] = u;
        if (q == 0) xu[nj] = 0.0;
        __syncthreads();
        u = xu[pos];
        const double uc = has_c ? xu[pos + 1] : 0.0;
        // ---------------- applyFτ!
        V3 Fext = {0.0, 0.0, 0.0}, Text = {0.0, 0.0, 0.0};
        if (C.kind != CCB_CTRL_NONE) {
            const Q4 qb = {z[3], z[4], z[5], z[6]};
            if (act_kind != 0 && u != 0.0) {
                const V3 ax = u * ld3(cj + 10);
                if (act_kind == 1) { const V3 Fb = qrot(qa2, ax); Fext = Fext + Fb; Text = Text + cross(ld3(cj + 3), qrotT(qb, Fb)); }
                else Text = Text + qrotT(qb, qrot(qa2, ax));
            }
            if (has_c && uc != 0.0) {
                const V3 ax = uc * ld3(ccj + 10);
                if (cact_kind == 1) { Fext = Fext - qrot(qb, ax); Text = Text - cross(ld3(ccj), ax); }
                else if (cact_kind == 2) Text = Text - ax;
            }
        }
        double c0[6];
        {
            const V3 v1 = v3(z[7], z[8], z[9]), w1 = v3(z[10], z[11], z[12]);
            const double sq1 = sqrt(c4 - dot(w1, w1));
            const V3 Jw1 = mat3_mul(cj + 13, w1);
            const V3 cR = cross(w1, Jw1) - sq1 * Jw1 - 2.0 * Text;
            c0[0] = -mass * v1.x * idt - Fext.x; c0[1] = -mass * v1.y * idt - Fext.y; c0[2] = -mass * (v1.z * idt + M.g) - Fext.z;
            c0[3] = cR.x; c0[4] = cR.y; c0[5] = cR.z;
        }
        // ---------------- my Jacobian row at the knot
        {
            double g;
            V3 Gax, Gap, Gbx, Gbp;
            row_eval(kind, pr, ld3(cj), ld3(cj + 3), ld4(cj + 6), xa2, qa2, v3(z[0], z[1], z[2]), Q4{z[3], z[4], z[5], z[6]}, true, g, Gax, Gap, Gbx, Gbp);
            if (!has_p) { Gax = v3(0.0, 0.0, 0.0); Gap = Gax; }
            if (rowpad) { Gax = v3(0.0, 0.0, 0.0); Gap = Gax; Gbx = Gax; Gbp = Gax; }
            if (lane_ok) {
                double* o = gk_own + 12 * r;
                st3(o, Gax); st3(o + 3, Gap); st3(o + 6, Gbx); st3(o + 9, Gbp);
            }
            __syncthreads();
        }
        double FB[5], FA[5];
#pragma unroll
        for (int t = 0; t < 5; t++) {
            FB[t] = gk_own[12 * t + 6 + r];
            FA[t] = has_c ? gk_chi[12 * t + r] : 0.0;
        }

        double p3[7], Dinv[9], Mw[9], d[6];
        double rowP[5], rowO[5], rowC[5], rhs;
        // force sum: xl must hold the multipliers (or steps) of every row; returns the 6-vector acting on my body
        auto force_sum = [&](double* out6) {
            double f = 0.0, f5 = 0.0;
            const double* lo = xl + 5 * pos;
            const double* lc = xl + 5 * pos + 5;         // child rows (zeros past the leaf; FA = 0 there anyway)
#pragma unroll
            for (int t = 0; t < 5; t++) {
                f = fma(FB[t], lo[t], fma(FA[t], lc[t], f));
                f5 = fma(gk_own[12 * t + 11], lo[t], f5);
                if (has_c) f5 = fma(gk_chi[12 * t + 5], lc[t], f5);
            }
            if (lane_ok) { xf[pos * 6 + r] = f; if (r == 0) xf[pos * 6 + 5] = f5; }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 6; c++) out6[c] = xf[pos * 6 + c];
        };
        auto evaluate = [&](bool jac) -> double {
            const V3 v2 = v3(s[0], s[1], s[2]), w2 = v3(s[3], s[4], s[5]);
            const double sq2 = sqrt(c4 - dot(w2, w2));
            {
                p3[0] = fma(dt, v2.x, z[0]); p3[1] = fma(dt, v2.y, z[1]); p3[2] = fma(dt, v2.z, z[2]);
                const Q4 q3 = qmul(Q4{z[3], z[4], z[5], z[6]}, Q4{0.5 * dt * sq2, 0.5 * dt * w2.x, 0.5 * dt * w2.y, 0.5 * dt * w2.z});
                p3[3] = q3.s; p3[4] = q3.x; p3[5] = q3.y; p3[6] = q3.z;
            }
            if (lane_ok) {
                xl[q] = lam;
                if (r == 0) {
#pragma unroll
                    for (int i = 0; i < 7; i++) xpose[pos * 8 + i] = p3[i];
                }
            }
            __syncthreads();
            double J[9];
#pragma unroll
            for (int i = 0; i < 9; i++) J[i] = cj[13 + i];
            const V3 Jw2 = mat3_mul(J, w2);
            const V3 dR = sq2 * Jw2 + cross(w2, Jw2);
            double fs[6];
            force_sum(fs);
            d[0] = fma(mass * idt, v2.x, c0[0]) - fs[0]; d[1] = fma(mass * idt, v2.y, c0[1]) - fs[1]; d[2] = fma(mass * idt, v2.z, c0[2]) - fs[2];
            d[3] = dR.x + c0[3] - fs[3]; d[4] = dR.y + c0[4] - fs[4]; d[5] = dR.z + c0[5] - fs[5];
            double ss = 0.0;
            if (lane_ok && r == 0) {
#pragma unroll
                for (int c = 0; c < 6; c++) ss = fma(d[c], d[c], ss);
            }
            if (jac) {
                double Dr[9], Tw[9];
                const double w[3] = {w2.x, w2.y, w2.z}, jw[3] = {Jw2.x, Jw2.y, Jw2.z};
                const double isq = 1.0 / sq2;
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int qq = 0; qq < 3; qq++) {
                        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
                        const double wxJ = w[i1] * J[i2 * 3 + qq] - w[i2] * J[i1 * 3 + qq];
                        const double jwx = (qq == i1) ? -jw[i2] : ((qq == i2) ? jw[i1] : 0.0);
                        Dr[i * 3 + qq] = sq2 * J[i * 3 + qq] + wxJ - jwx - jw[i] * w[qq] * isq;
                    }
                inv3(Dr, Dinv);
                const double c = 0.25 * dt * dt;
                Tw[0] = c * (w[0] * w[0] * isq + sq2); Tw[1] = c * (w[0] * w[1] * isq + w[2]); Tw[2] = c * (w[0] * w[2] * isq - w[1]);
                Tw[3] = c * (w[1] * w[0] * isq - w[2]); Tw[4] = c * (w[1] * w[1] * isq + sq2); Tw[5] = c * (w[1] * w[2] * isq + w[0]);
                Tw[6] = c * (w[2] * w[0] * isq + w[1]); Tw[7] = c * (w[2] * w[1] * isq - w[0]); Tw[8] = c * (w[2] * w[2] * isq + sq2);
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int qq = 0; qq < 3; qq++) Mw[i * 3 + qq] = fma(Tw[i * 3], Dinv[qq], fma(Tw[i * 3 + 1], Dinv[3 + qq], Tw[i * 3 + 2] * Dinv[6 + qq]));
                if (lane_ok && r == 0) {
#pragma unroll
                    for (int i = 0; i < 6; i++) xd[pos * 6 + i] = d[i];
#pragma unroll
                    for (int i = 0; i < 9; i++) xM[pos * 9 + i] = Mw[i];
                }
                __syncthreads();
            }
            V3 xa3 = {0.0, 0.0, 0.0};
            Q4 qa3 = {1.0, 0.0, 0.0, 0.0};
            if (has_p) { xa3 = ld3(xpose + ppos * 8); qa3 = ld4(xpose + ppos * 8 + 3); }
            double g;
            V3 Gax, Gap, Gbx, Gbp;
            row_eval(kind, pr, ld3(cj), ld3(cj + 3), ld4(cj + 6), xa3, qa3, v3(p3[0], p3[1], p3[2]), Q4{p3[3], p3[4], p3[5], p3[6]}, jac, g, Gax, Gap, Gbx, Gbp);
            if (rowpad) g = 0.0;
            ss = fma(g, g, ss);
            rhs = g;
            if (jac) {
                const double* pd = xd + ppos * 6;
                const double* pM = xM + ppos * 9;
                const double cb = dt * kdt, ca = dt * pkdt;
                double wB[6], wA[6];
                { const V3 t = vec_mat3(Gbp, Mw); wB[0] = cb * Gbx.x; wB[1] = cb * Gbx.y; wB[2] = cb * Gbx.z; wB[3] = t.x; wB[4] = t.y; wB[5] = t.z; }
                { const V3 t = vec_mat3(Gap, pM); wA[0] = ca * Gax.x; wA[1] = ca * Gax.y; wA[2] = ca * Gax.z; wA[3] = t.x; wA[4] = t.y; wA[5] = t.z; }
                if (!has_p) {
#pragma unroll
                    for (int c = 0; c < 6; c++) wA[c] = 0.0;
                }
                double acc = g;
#pragma unroll
                for (int c = 0; c < 6; c++) acc = fma(-wB[c], d[c], acc);
#pragma unroll
                for (int c = 0; c < 6; c++) acc = fma(-wA[c], pd[c], acc);
                rhs = acc;
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    const double2* go = reinterpret_cast<const double2*>(gk_own + 12 * t);
                    const double2* gpp = reinterpret_cast<const double2*>(gk_par + 12 * t + 6);
                    const double2* gcc = reinterpret_cast<const double2*>(gk_chi + 12 * t);
                    double so = 0.0, sp = 0.0, sc = 0.0;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double2 a2 = go[c], b2 = go[3 + c], p2 = gpp[c], c2 = gcc[c];
                        so = fma(wA[2 * c], a2.x, fma(wA[2 * c + 1], a2.y, so));
                        so = fma(wB[2 * c], b2.x, fma(wB[2 * c + 1], b2.y, so));
                        sp = fma(wA[2 * c], p2.x, fma(wA[2 * c + 1], p2.y, sp));
                        sc = fma(wB[2 * c], c2.x, fma(wB[2 * c + 1], c2.y, sc));
                    }
                    rowO[t] = so;
                    rowP[t] = has_p ? sp : 0.0;
                    rowC[t] = has_c ? sc : 0.0;
                }
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    if (t == r) rowO[t] += M.reg;
                    if (rowpad) { rowO[t] = (t == r) ? 1.0 : 0.0; rowP[t] = 0.0; rowC[t] = 0.0; }
                }
                if (rowpad) rhs = 0.0;
            }
            return sqrt(sys_sum(ss));
        };

        double dl = 0.0, ds[6];
        // S dλ = rhs, one row per lane, block tridiagonal along the chain.  Two-sided ("twisted") elimination: the blocks below the
        // middle block mid = nj / 2 are eliminated from the leaf inwards and, in the same rounds, the blocks above it from the root
        // inwards (the two pivot streams touch disjoint rows except for the middle block, which takes both updates); then the middle
        // block; then the substitution runs outwards from the middle in both directions at once.  5 (max(mid, nj-1-mid) + 1) pivot
        // rounds instead of 5 nj.  Inside a block the elimination is Gauss-Jordan, so a block's rows end up diagonal in their own block.
        // Lanes above the middle swap their parent / child couplings, so that for every lane rowP is the coupling TOWARDS the middle
        // (still to be eliminated against) and rowC the coupling AWAY from it, and one update code serves both streams.
        auto solve = [&]() {
            double myinv = 1.0;
            int par = 0;
            const int mid = nj >> 1, nst_blocks = max(mid, nj - 1 - mid);
            const bool top = lane_ok && pos < mid, ismid = lane_ok && pos == mid;
            if (top) {
#pragma unroll
                for (int c = 0; c < 5; c++) { const double t = rowP[c]; rowP[c] = rowC[c]; rowC[c] = t; }
            }
            const int wlo = (q >> 5) << 5, whi = wlo + 31;     // rows held by my warp
            const bool warp_has_mid = !(wlo > 5 * mid + 4 || whi < 5 * mid);
            for (int sb = 0; sb < nst_blocks; sb++) {
                const int ppb = nj - 1 - sb, ppt = sb;           // pivot blocks of the two streams
                const bool vb = ppb > mid, vt = ppt < mid;
                const int ppv = top ? ppt : ppb;                 // the pivot block that can concern me, and its neighbour towards the middle
                const int nbv = top ? ppt + 1 : ppb - 1;
                const bool vv = top ? vt : vb;
#pragma unroll
                for (int rk = 4; rk >= 0; rk--) {
                    double* bufB = pivb + 24 * par;
                    double* bufT = bufB + 12;
                    par ^= 1;
                    if (lane_ok && vv && pos == ppv && r == rk) {       // publish my row: own block, coupling towards the middle, rhs, pivot
                        double* buf = top ? bufT : bufB;
#pragma unroll
                        for (int c = 0; c < 5; c++) { buf[c] = rowO[c]; buf[5 + c] = rowP[c]; }
                        buf[10] = rhs; buf[11] = rowO[rk];
                    }
                    __syncthreads();
                    // only the warps holding rows of a pivot block or of its inner neighbour have anything to update
                    const bool hitB = vb && !(wlo > 5 * ppb + 4 || whi < 5 * (ppb - 1));
                    const bool hitT = vt && !(wlo > 5 * (ppt + 1) + 4 || whi < 5 * ppt);
                    if (!hitB && !hitT) continue;
                    {
                        const double* buf = top ? bufT : bufB;
                        const double inv = rcp_nr(vv ? buf[11] : 1.0);
                        if (lane_ok && vv && pos == ppv && r == rk) myinv = inv;
                        const bool same = lane_ok && vv && pos == ppv && r != rk, nbr = lane_ok && vv && pos == nbv;
                        const double fS = same ? rowO[rk] * inv : 0.0;
                        const double fN = nbr ? rowC[rk] * inv : 0.0;
#pragma unroll
                        for (int c = 0; c < rk; c++) {
                            const double pv = buf[c];
                            rowO[c] = fma(-fS, pv, rowO[c]);
                            rowC[c] = fma(-fN, pv, rowC[c]);
                        }
#pragma unroll
                        for (int c = 0; c < 5; c++) {
                            const double pv = buf[5 + c];
                            rowP[c] = fma(-fS, pv, rowP[c]);
                            rowO[c] = fma(-fN, pv, rowO[c]);
                        }
                        rhs = fma(-(fS + fN), buf[10], rhs);
                    }
                    // the middle block is the inner neighbour of the root-side stream as well: its coupling to block mid - 1 is rowP
                    if (warp_has_mid && vt && ppt + 1 == mid) {
                        const double inv = rcp_nr(bufT[11]);
                        const double fN = ismid ? rowP[rk] * inv : 0.0;
#pragma unroll
                        for (int c = 0; c < rk; c++) rowP[c] = fma(-fN, bufT[c], rowP[c]);
#pragma unroll
                        for (int c = 0; c < 5; c++) rowO[c] = fma(-fN, bufT[5 + c], rowO[c]);
                        rhs = fma(-fN, bufT[10], rhs);
                    }
                }
            }
            // the middle block (nothing outside it is updated: its neighbours are already eliminated)
#pragma unroll
            for (int rk = 4; rk >= 0; rk--) {
                double* buf = pivb + 24 * par;
                par ^= 1;
                if (ismid && r == rk) {
#pragma unroll
                    for (int c = 0; c < 5; c++) buf[c] = rowO[c];
                    buf[10] = rhs; buf[11] = rowO[rk];
                }
                __syncthreads();
                if (!warp_has_mid) continue;
                const double inv = rcp_nr(buf[11]);
                if (ismid && r == rk) myinv = inv;
                const double fS = (ismid && r != rk) ? rowO[rk] * inv : 0.0;
#pragma unroll
                for (int c = 0; c < rk; c++) rowO[c] = fma(-fS, buf[c], rowO[c]);
                rhs = fma(-fS, buf[10], rhs);
            }
            // substitution outwards from the middle, one block of each side per round: x = (rhs - (coupling towards the middle) x_inner) / pivot
            for (int j = 0; j <= nst_blocks; j++) {
                if (lane_ok && ((j == 0 && ismid) || (j > 0 && (pos == mid - j || pos == mid + j)))) {
                    double x = rhs;
                    if (j > 0) {
                        const double* xp = xl + 5 * (top ? pos + 1 : pos - 1);
#pragma unroll
                        for (int c = 0; c < 5; c++) x = fma(-rowP[c], xp[c], x);
                    }
                    x *= myinv;
                    rhs = x;
                    xl[q] = x;
                }
                __syncthreads();
            }
            dl = rhs;
            double fs[6];
            force_sum(fs);
            const V3 dw = mat3_mul(Dinv, v3(d[3] + fs[3], d[4] + fs[4], d[5] + fs[5]));
            ds[0] = kdt * (d[0] + fs[0]); ds[1] = kdt * (d[1] + fs[1]); ds[2] = kdt * (d[2] + fs[2]);
            ds[3] = dw.x; ds[4] = dw.y; ds[5] = dw.z;
        };

        // newton! / lineSearch! around a single evaluation site (as in K1r): the evaluation at an accepted line-search point is the
        // evaluation the next Newton iteration starts from, so it also forms the Jacobian rows; the iterates are unchanged
        bool done = false, first = true;
        double normf0 = 0.0, alpha = 1.0, full = 0.0, lam0 = lam;
        double s0[6];
#pragma unroll
        for (int i = 0; i < 6; i++) s0[i] = s[i];
        int n = 0, ls = 0;
        for (;;) {
            const double nfe = evaluate(true);
            if (first) {
                normf0 = nfe;
                first = false;
            } else {
                if (nfe >= normf0 && ls < A.line_iter) {          // lineSearch!: not decreased -> halve the step and try again
                    alpha *= 0.5; ls++;
#pragma unroll
                    for (int i = 0; i < 6; i++) s[i] = fma(-alpha, ds[i], s0[i]);
                    lam = fma(-alpha, dl, lam0);
                    continue;
                }
                if (alpha * full < A.eps && nfe < A.eps) done = true;
                normf0 = nfe;
                if (done || n >= A.newton_iter) break;
            }
            if (!(nfe == nfe)) break;      // NaN residual (accepted by lineSearch! like the reference: NaN >= x is false): nothing can recover, step failed
            solve();
            n++;
            double sd = lane_ok ? dl * dl : 0.0;
            if (lane_ok && r == 0) {
#pragma unroll
                for (int i = 0; i < 6; i++) sd = fma(ds[i], ds[i], sd);
            }
            full = sqrt(sys_sum(sd));
#pragma unroll
            for (int i = 0; i < 6; i++) s0[i] = s[i];
            lam0 = lam;
            alpha = 1.0; ls = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) s[i] = s0[i] - ds[i];
            lam = lam0 - dl;
            // a step below eps taken from a residual below eps ends the solve whatever scale the line search would pick
            if (full < A.eps && normf0 < A.eps) { done = true; break; }
        }
        const int it_used = n;
        total_iters += it_used;
        if (!done) failed = true;
        if (!A.knot_input) {
#pragma unroll
            for (int i = 0; i < 6; i++) z[7 + i] = s[i];
        }
    }
    if (lane_ok && A.knot_input) {
        if (r == 0) {
#pragma unroll
            for (int i = 0; i < 6; i++) A.s_out[(size_t)sys * 6 * M.nb + 6 * b + i] = s[i];
        }
        if (r < dim) A.lambda[(size_t)sys * M.ml + r0 + r] = lam;
    } else if (lane_ok) {
#pragma unroll
        for (int i = 0; i < 13; i++)
            if (i % 5 == r) A.state[(size_t)sys * nst + 13 * b + i] = z[i];
        if (r < dim) A.lambda[(size_t)sys * M.ml + r0 + r] = lam;
    }
    if (A.status && q == 0) write_status(A, sys, total_iters, failed, persist && chunk > 0);
    if (persist) {
        const bool shared = sys >= qown * nworkers;      // an extra system moves on to another CTA: release my stores, then the flag
        if (shared) __threadfence();
        __syncthreads();                                  // (own systems: the next chunk reloads what the other threads of this CTA stored)
        if (shared && q == 0) atomicExch(A.persist + 1 + sys, chunk + 1);
    }
    }   // work loop
}

}  // namespace

// Eligibility: a single serial chain (every joint at a distinct depth) with joints of dimension <= 5 and at most 1024 rows
bool mw_kernel_eligible(const DevMech& M) {
    return M.is_chain && M.nj >= 1 && M.chain_max_dim <= 5 && M.chain_max_depth == M.nj - 1 && 5 * M.nj <= 1024;
}

int launch_rollout_mw(const DevMech& M, const DevCtrl& C, const SimArgs& a_in, cudaStream_t stream) {
    SimArgs a = a_in;
    const int nj = M.nj;
    const int threads = ((5 * nj + 31) / 32) * 32;
    const size_t doubles = (size_t)nj * kGS + (size_t)nj * kCS + (size_t)nj * 8 + (size_t)nj * 6 + (size_t)nj * 9 + (5 * nj + 5) + (size_t)nj * 6 + 48 +
                           5 * nj + nj + (nj + 1) + 40 * nj + 8 * nj;
    const size_t smem = doubles * sizeof(double) + (size_t)nj * sizeof(int) + 16;
    if (smem > 227 * 1024) { set_error("mechanism too large for the multi-warp rows kernel"); return CCB_E_UNSUPPORTED; }
    int dev = 0, sms = 148;
    CCB_CUDA(cudaGetDevice(&dev));
    CCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static const int caps[3] = {255, 168, 128};
    int occ[3] = {0, 0, 0};
#define CCB_MW_OCC(NJT)                                                                                                               \
    do {                                                                                                                             \
        CCB_CUDA(cudaFuncSetAttribute(rollout_mw_kernel<NJT, 255>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));        \
        CCB_CUDA(cudaFuncSetAttribute(rollout_mw_kernel<NJT, 168>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));         \
        CCB_CUDA(cudaFuncSetAttribute(rollout_mw_kernel<NJT, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));         \
        CCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], rollout_mw_kernel<NJT, 255>, threads, smem));                \
        CCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], rollout_mw_kernel<NJT, 168>, threads, smem));                \
        CCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], rollout_mw_kernel<NJT, 128>, threads, smem));                \
    } while (0)
#define CCB_MW_LAUNCH(NJT)                                                                                                            \
    do {                                                                                                                             \
        if (pick == 0) rollout_mw_kernel<NJT, 255><<<grid, threads, smem, stream>>>(M, C, a, M.chain_depth);                         \
        else if (pick == 1) rollout_mw_kernel<NJT, 168><<<grid, threads, smem, stream>>>(M, C, a, M.chain_depth);                    \
        else rollout_mw_kernel<NJT, 128><<<grid, threads, smem, stream>>>(M, C, a, M.chain_depth);                                   \
    } while (0)
    // the attribute / occupancy queries are per (device, chain length, shared-memory size): remembered between launches
    // (the shared-memory opt-in is raised to the device maximum once per variant, so a remembered entry never needs it again)
    struct OccKey { int dev, nj; size_t smem; int occ[3]; };
    static OccKey cache[8];
    static int ncache = 0;
    static std::mutex cache_mu;
    std::lock_guard<std::mutex> lock(cache_mu);
    int hit = -1;
    for (int i = 0; i < ncache; i++)
        if (cache[i].dev == dev && cache[i].nj == nj && cache[i].smem == smem) hit = i;
    if (hit >= 0) {
        for (int i = 0; i < 3; i++) occ[i] = cache[hit].occ[i];
    } else {
        switch (nj) {          // the BASELINE chain lengths get their own instantiation: Sawyer arm, 8-pendulum, 32-pendulum
            case 7: CCB_MW_OCC(7); break;
            case 9: CCB_MW_OCC(9); break;
            case 33: CCB_MW_OCC(33); break;
            default: CCB_MW_OCC(0); break;
        }
        OccKey& k = cache[ncache < 8 ? ncache++ : 7];
        k.dev = dev; k.nj = nj; k.smem = smem;
        for (int i = 0; i < 3; i++) k.occ[i] = occ[i];
    }
    // which register cap: one wave at 255 if the batch fits; else the largest cap that still makes ONE wave; else (many waves) the largest
    // cap that keeps at least 8 warps per SM resident
    int pick = -1;
    const int wpc = threads / 32;
    if ((long long)a.B <= (long long)sms * occ[0]) pick = 0;
    for (int i = 1; i < 3 && pick < 0; i++)
        if ((long long)a.B <= (long long)sms * occ[i]) pick = i;
    for (int i = 0; i < 3 && pick < 0; i++)
        if (occ[i] * wpc >= 8) pick = i;
    if (pick < 0) pick = 2;
    if (const char* e = getenv("CCB_K1M_REGS")) {      // experiments: force a variant
        const int v = atoi(e);
        for (int i = 0; i < 3; i++) if (caps[i] == v) pick = i;
    }
    // Persistent time-sliced launch (work loop in the kernel): a batch that does not fit one wave of the spill-free 255-register variant runs on
    // exactly that wave anyway, W = sms * occ[0] CTAs simulating B systems in chunks of steps.  Measured on 2000 Sawyer steps
    // (tools/k1m_probe.py): the 255-register variant carries 592 systems in 38.0 ms (15.6 systems/ms), the 128-register variant that a plain
    // launch needs to get 1024 systems into one wave 1024 in 73.4 ms (13.9/ms, spills), 168 registers 592 in 48.8 ms.  Two-warp CTAs only
    // (Sawyer arm, 8-pendulum): longer chains want more resident warps than 255 registers allow.  Not for PID (controller state on the device).
    int grid = a.B;
    const bool persist = a.persist && !a.knot_input && a.steps >= 256 && wpc <= 2 && occ[0] > 0 && (long long)a.B > (long long)sms * occ[0] &&
                         C.kind != CCB_CTRL_PID && !getenv("CCB_K1M_NO_PERSIST") && !getenv("CCB_K1M_REGS");
    if (persist) {
        pick = 0;
        grid = sms * occ[0];
        const int want = std::max(1, std::min(32, a.steps / 32));
        a.chunk_steps = (a.steps + want - 1) / want;
        a.nchunks = (a.steps + a.chunk_steps - 1) / a.chunk_steps;
        CCB_CUDA(cudaMemsetAsync(a.persist, 0, (size_t)(a.B + 2) * sizeof(int), stream));
    } else {
        a.persist = nullptr; a.nchunks = 1; a.chunk_steps = a.steps;
    }
    switch (nj) {
        case 7: CCB_MW_LAUNCH(7); break;
        case 9: CCB_MW_LAUNCH(9); break;
        case 33: CCB_MW_LAUNCH(33); break;
        default: CCB_MW_LAUNCH(0); break;
    }
#undef CCB_MW_LAUNCH
#undef CCB_MW_OCC
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
