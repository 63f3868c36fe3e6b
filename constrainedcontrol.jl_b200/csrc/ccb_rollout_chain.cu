// csrc/ccb_rollout_chain.cu -- K1c: fused closed-loop rollout for serial-chain mechanisms (every BASELINE config but
// the closed-chain deltabot: pendulum, cartpole, n-pendulum, acrobot, Sawyer arm, ...).
//
// Mapping: ONE LANE owns ONE JOINT AND ITS CHILD BODY; a system of nj joints occupies nj consecutive lanes and a warp
// holds floor(32/nj) systems, so every lane of the warp runs the same scalar code on a different (joint, body) pair
// (the general kernel K1 in ccb_rollout.cu needs 8..32 lanes per system with ~1/3 of them busy).  Neighbouring lanes
// exchange the few vectors that couple them (parent pose, parent residual, Schur update, λ step) with warp shuffles;
// the per-lane blocks that do not fit in registers (Jacobian rows at the knot, the 5x5/6x6 rows of the Schur complement)
// live in shared memory, lane-interleaved so that every access is conflict free.
//
// Same algorithm and the same semantics as K1 (simulate! loop of ConstrainedDynamics.jl, SURVEY.md 3.2, with the control
// laws of reference src/control/lqr.jl:89-139, lqr_tracking.jl:46-71, pid.jl:43-88 fused): bodies are eliminated first,
// the constraint Schur complement S = Gv D^-1 G^T is swept leaf -> root along the chain (ldu_factorization! role) with the
// rhs carried along, then root -> leaf back substitution (ldu_backsubstitution! role).
#include <algorithm>
#include "ccb_dev_math.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

// lane-interleaved shared memory: element i of this lane's private array
#define LSM(base, i) (base)[(i) * 32 + wl]

template <int MD> struct ChainSmem {
    // offsets in "doubles per lane"
    static constexpr int GKA = 0;                  // [MD][6] own rows at the knot, parent side
    static constexpr int GKB = GKA + MD * 6;       // [MD][6] own rows at the knot, child side
    static constexpr int PGB = GKB + MD * 6;       // [MD][6] parent's joint rows, child side (act on my parent body)
    static constexpr int CGA = PGB + MD * 6;       // [MD][6] child's joint rows, parent side (act on my body)
    static constexpr int SJJ = CGA + MD * 6;       // [MD][MD]
    static constexpr int SJP = SJJ + MD * MD;      // [MD][MD] coupling to the parent joint; after the solve: U
    static constexpr int SJC = SJP + MD * MD;      // [MD][MD] coupling to the child joint
    static constexpr int S0 = SJC + MD * MD;       // [6] iterate before the step
    static constexpr int L0 = S0 + 6;              // [MD]
    static constexpr int DS = L0 + MD;             // [6] body step
    static constexpr int DL = DS + 6;              // [MD] lambda step
    static constexpr int TOTAL = DL + MD;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <int MD>
__global__ void __launch_bounds__(32) rollout_chain_kernel(const DevMech M, const DevCtrl C, const SimArgs A, const int* __restrict__ chain_parent,
                                                           const int* __restrict__ chain_child, const int* __restrict__ chain_depth, int max_depth, int spw) {
    extern __shared__ double smem[];
    using L = ChainSmem<MD>;
    const int wl = threadIdx.x;                       // lane in warp
    const int G = M.nj;                               // lanes per system
    // spw = systems per warp (<= 32/G): small batches use partly filled warps so that more SM sub-partitions get a warp
    const int sw = wl / G;                            // system slot in the warp
    const int j = wl - sw * G;                        // my joint
    const int sys = blockIdx.x * spw + sw;
    const bool valid = (sw < spw) && (sys < A.B);
    const int base = sw * G;                          // first lane of my system
    const int jj = valid ? j : 0;
    const int b = __ldg(M.j_child + jj);              // my body
    const int pj = __ldg(chain_parent + jj), cj = __ldg(chain_child + jj);   // parent / child joint or -1
    const int pl = (valid && pj >= 0) ? base + pj : wl, cl = (valid && cj >= 0) ? base + cj : wl;   // their lanes (self when absent)
    const int depth = __ldg(chain_depth + jj);
    const int dim = __ldg(M.j_dim + jj), r0 = __ldg(M.j_row0 + jj);
    const int pdim = pj >= 0 ? __ldg(M.j_dim + pj) : 0, cdim = cj >= 0 ? __ldg(M.j_dim + cj) : 0;
    const bool has_p = valid && pj >= 0, has_c = valid && cj >= 0;
    const double dt = M.dt, idt = 1.0 / M.dt, c4 = 4.0 / (M.dt * M.dt);
    const double mass = __ldg(M.mass + b), kdt = dt / mass;
    const int pb = pj >= 0 ? __ldg(M.j_child + pj) : -1;
    const double pkdt = pb >= 0 ? dt / __ldg(M.mass + pb) : 0.0;
    double J[9];
#pragma unroll
    for (int i = 0; i < 9; i++) J[i] = __ldg(M.inertia + 9 * b + i);
    const V3 p1 = ldg3(M.j_p1 + 3 * jj), p2 = ldg3(M.j_p2 + 3 * jj), axis = ldg3(M.j_axis + 3 * jj);
    const Q4 qoff = ldg4(M.j_qoff + 4 * jj);
    int kindmask = 0;
    for (int r = 0; r < dim; r++) kindmask |= __ldg(M.row_kind + r0 + r) << r;
    // chain mechanisms: eqc index == joint index (checked by chain_tables)
    const int act_kind = __ldg(M.eqc_act_kind + jj);
    const int nst = 13 * M.nb;

    // ---- state in registers
    double z[13], s[6], lam[MD];
    if (valid) {
#pragma unroll
        for (int i = 0; i < 13; i++) z[i] = A.state[(size_t)sys * nst + 13 * b + i];
#pragma unroll
        for (int r = 0; r < MD; r++) lam[r] = r < dim ? A.lambda[(size_t)sys * M.ml + r0 + r] : 0.0;
    } else {
#pragma unroll
        for (int i = 0; i < 13; i++) z[i] = i == 3 ? 1.0 : 0.0;
#pragma unroll
        for (int r = 0; r < MD; r++) lam[r] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) s[i] = 0.0;

    // sum over the lanes of my system
    auto sys_sum = [&](double v) {
        double acc = 0.0;
        for (int k = 0; k < G; k++) acc += shfl_d(v, base + k);
        return acc;
    };

    int total_iters = 0;
    bool failed = false;

    for (int step = 0; step < A.steps; step++) {
        const int k = A.k0 + step;
        if (valid && A.storage) {
            double* out = A.storage + ((size_t)step * A.B + sys) * nst + 13 * b;
#pragma unroll
            for (int i = 0; i < 13; i++) out[i] = z[i];
        }
        // ---------------- control!(k) part 1: laws that read the continuous state (PID), open-loop tables
        double u = 0.0;   // input of MY eqc
        if (C.kind == CCB_CTRL_PID || C.kind == CCB_CTRL_TABLE) {
            // parent's continuous pose (PID minimal coordinates)
            double pz[7];
#pragma unroll
            for (int i = 0; i < 7; i++) pz[i] = shfl_d(z[i], pl);
            if (valid) {
                for (int i = 0; i < C.nu; i++) {
                    if (__ldg(C.eqc + i) != jj) continue;
                    if (C.kind == CCB_CTRL_TABLE) {
                        if (k >= 1 && k <= C.N) u += __ldg(C.U + (size_t)sys * C.U_bs + (size_t)(k - 1) * C.nu + i);
                        continue;
                    }
                    V3 xa = {0.0, 0.0, 0.0};
                    Q4 qa = {1.0, 0.0, 0.0, 0.0};
                    if (has_p) { xa = v3(pz[0], pz[1], pz[2]); qa = Q4{pz[3], pz[4], pz[5], pz[6]}; }
                    V3 xb = v3(z[0], z[1], z[2]);
                    Q4 qb = {z[3], z[4], z[5], z[6]};
                    double coord;
                    if (act_kind == 1) {
                        coord = dot(axis, qrotT(qa, xb + qrot(qb, p2) - xa) - p1);
                    } else {
                        Q4 qe = qmul(qmul(qconj(qa), qb), qconj(qoff));
                        double nv = sqrt(qe.x * qe.x + qe.y * qe.y + qe.z * qe.z);
                        coord = (nv == 0.0) ? 0.0 : 2.0 * atan2(nv, qe.s) * (axis.x * qe.x + axis.y * qe.y + axis.z * qe.z) / nv;
                    }
                    double err = __ldg(C.goal + (size_t)sys * C.goal_bs + i) - coord;
                    if (act_kind == 2) {           // pid.jl:46-53 (Translational3 + Rotational2)
                        const double pi = 3.14159265358979323846;
                        if (err > pi) err -= 2.0 * pi;
                        else if (err < -pi) err += 2.0 * pi;
                    }
                    double* integ = C.pid_int + (size_t)sys * C.nu + i;
                    double* last = C.pid_last + (size_t)sys * C.nu + i;
                    if (k == 1) *last = err;                  // pid.jl:73
                    const double in = *integ + err * dt;      // pid.jl:76
                    *integ = in;
                    const double der = (err - *last) / dt;    // pid.jl:77
                    u += __ldg(C.P + i) * err + __ldg(C.I + i) * in + __ldg(C.D + i) * der;
                    *last = err;
                }
            }
        }
        // ---------------- knot: x2 = xc + vc dt, q2 = qc (x) w-bar(wc); Newton starts from (v1, w1)
        if (!A.knot_input) {
            V3 v = v3(z[7], z[8], z[9]), w = v3(z[10], z[11], z[12]);
            z[0] = fma(dt, v.x, z[0]); z[1] = fma(dt, v.y, z[1]); z[2] = fma(dt, v.z, z[2]);
            const double sq = sqrt(c4 - dot(w, w));
            Q4 q2 = qmul(Q4{z[3], z[4], z[5], z[6]}, Q4{0.5 * dt * sq, 0.5 * dt * w.x, 0.5 * dt * w.y, 0.5 * dt * w.z});
            z[3] = q2.s; z[4] = q2.x; z[5] = q2.y; z[6] = q2.z;
        }
#pragma unroll
        for (int i = 0; i < 6; i++) s[i] = z[7 + i];
        // parent's knot pose (fixed for the step)
        double pz2[7];
#pragma unroll
        for (int i = 0; i < 7; i++) pz2[i] = shfl_d(z[i], pl);
        const V3 xa2 = has_p ? v3(pz2[0], pz2[1], pz2[2]) : v3(0.0, 0.0, 0.0);
        const Q4 qa2 = has_p ? Q4{pz2[3], pz2[4], pz2[5], pz2[6]} : Q4{1.0, 0.0, 0.0, 0.0};
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
                double dz[12];
                if (valid) {
                    const double* d = zd + 13 * b;
                    Q4 qe = qmul(qconj(ldg4(d + 3)), Q4{z[3], z[4], z[5], z[6]});
#pragma unroll
                    for (int i = 0; i < 3; i++) { dz[i] = z[i] - __ldg(d + i); dz[3 + i] = z[7 + i] - __ldg(d + 7 + i); dz[9 + i] = z[10 + i] - __ldg(d + 10 + i); }
                    dz[6] = qe.x; dz[7] = qe.y; dz[8] = qe.z;
                }
                for (int i = 0; i < C.nu; i++) {
                    double part = 0.0;
                    if (valid) {
                        const double* Kr = K + (size_t)i * n + 12 * b;
#pragma unroll
                        for (int c = 0; c < 12; c++) part = fma(__ldg(Kr + c), dz[c], part);
                    }
                    const double tot = sys_sum(part);
                    if (valid && __ldg(C.eqc + i) == jj) u += __ldg(Fd + i) - tot;
                }
            }
        }
        // ---------------- applyFτ!: my body gets the child side of my joint's input and the parent side of my child's
        V3 Fext = {0.0, 0.0, 0.0}, Text = {0.0, 0.0, 0.0};
        {
            const double uc = shfl_d(u, cl);            // child's input
            double cq[4];                               // child's knot orientation
#pragma unroll
            for (int i = 0; i < 4; i++) cq[i] = shfl_d(z[3 + i], cl);
            if (valid && C.kind != CCB_CTRL_NONE) {
                const Q4 qb = {z[3], z[4], z[5], z[6]};
                if (act_kind != 0 && u != 0.0) {
                    V3 ax = u * axis;
                    if (act_kind == 1) { V3 Fb = qrot(qa2, ax); Fext = Fext + Fb; Text = Text + cross(p2, qrotT(qb, Fb)); }
                    else Text = Text + qrotT(qb, qrot(qa2, ax));
                }
                if (has_c && uc != 0.0) {
                    const int ck = __ldg(M.eqc_act_kind + cj);
                    V3 ax = uc * ldg3(M.j_axis + 3 * cj);
                    (void)cq;
                    if (ck == 1) { Fext = Fext - qrot(qb, ax); Text = Text - cross(ldg3(M.j_p1 + 3 * cj), ax); }
                    else if (ck == 2) Text = Text - ax;
                }
            }
        }
        // ---------------- Jacobian rows at the knot (they carry the constraint forces; fixed for the step)
        if (valid) {
            const V3 xb = v3(z[0], z[1], z[2]);
            const Q4 qb = {z[3], z[4], z[5], z[6]};
#pragma unroll
            for (int r = 0; r < MD; r++) {
                if (r < dim) {
                    double g;
                    V3 Gax, Gap, Gbx, Gbp;
                    row_eval((kindmask >> r) & 1, ldg3(M.row_proj + 3 * (r0 + r)), p1, p2, qoff, xa2, qa2, xb, qb, true, g, Gax, Gap, Gbx, Gbp);
                    if (!has_p) { Gax = v3(0.0, 0.0, 0.0); Gap = Gax; }
                    LSM(smem, L::GKA + r * 6 + 0) = Gax.x; LSM(smem, L::GKA + r * 6 + 1) = Gax.y; LSM(smem, L::GKA + r * 6 + 2) = Gax.z;
                    LSM(smem, L::GKA + r * 6 + 3) = Gap.x; LSM(smem, L::GKA + r * 6 + 4) = Gap.y; LSM(smem, L::GKA + r * 6 + 5) = Gap.z;
                    LSM(smem, L::GKB + r * 6 + 0) = Gbx.x; LSM(smem, L::GKB + r * 6 + 1) = Gbx.y; LSM(smem, L::GKB + r * 6 + 2) = Gbx.z;
                    LSM(smem, L::GKB + r * 6 + 3) = Gbp.x; LSM(smem, L::GKB + r * 6 + 4) = Gbp.y; LSM(smem, L::GKB + r * 6 + 5) = Gbp.z;
                } else {
#pragma unroll
                    for (int c = 0; c < 6; c++) { LSM(smem, L::GKA + r * 6 + c) = 0.0; LSM(smem, L::GKB + r * 6 + c) = 0.0; }
                }
            }
        }
        __syncwarp();
        // neighbours' rows: parent's child-side rows act on my parent body, child's parent-side rows act on my body
        for (int i = 0; i < MD * 6; i++) {
            const double vb = smem[(L::GKB + i) * 32 + pl];   // direct read of the neighbour lane's column (same warp)
            const double va = smem[(L::GKA + i) * 32 + cl];
            LSM(smem, L::PGB + i) = has_p ? vb : 0.0;
            LSM(smem, L::CGA + i) = has_c ? va : 0.0;
        }
        __syncwarp();

        // ---------------- newton! with lineSearch!
        // evaluates the residual at (s, lam); with JAC also builds rhs and the S blocks.  Returns the system's |f|.
        double p3[7], Dinv[9], Mw[9], d[6], rhs[MD];
        auto evaluate = [&](bool jac, bool act) -> double {
            double ss = 0.0;
            // --- my body
            const V3 v1 = v3(z[7], z[8], z[9]), w1 = v3(z[10], z[11], z[12]), v2 = v3(s[0], s[1], s[2]), w2 = v3(s[3], s[4], s[5]);
            const double sq2 = sqrt(c4 - dot(w2, w2));
            {
                p3[0] = fma(dt, v2.x, z[0]); p3[1] = fma(dt, v2.y, z[1]); p3[2] = fma(dt, v2.z, z[2]);
                Q4 q3 = qmul(Q4{z[3], z[4], z[5], z[6]}, Q4{0.5 * dt * sq2, 0.5 * dt * w2.x, 0.5 * dt * w2.y, 0.5 * dt * w2.z});
                p3[3] = q3.s; p3[4] = q3.x; p3[5] = q3.y; p3[6] = q3.z;
            }
            const double sq1 = sqrt(c4 - dot(w1, w1));
            const V3 Jw1 = mat3_mul(J, w1), Jw2 = mat3_mul(J, w2);
            const V3 dR = sq2 * Jw2 + cross(w2, Jw2) - sq1 * Jw1 + cross(w1, Jw1) - 2.0 * Text;
            d[0] = mass * ((v2.x - v1.x) * idt) - Fext.x; d[1] = mass * ((v2.y - v1.y) * idt) - Fext.y; d[2] = mass * ((v2.z - v1.z) * idt - M.g) - Fext.z;
            d[3] = dR.x; d[4] = dR.y; d[5] = dR.z;
            // constraint forces: my joint (child side) and my child's joint (parent side)
            double tA[6] = {0, 0, 0, 0, 0, 0};   // force of my joint on my PARENT body = GkA^T lam
#pragma unroll
            for (int r = 0; r < MD; r++) {
                const double l = lam[r];
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    d[c] = fma(-LSM(smem, L::GKB + r * 6 + c), l, d[c]);
                    tA[c] = fma(LSM(smem, L::GKA + r * 6 + c), l, tA[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 6; c++) { const double t = shfl_d(tA[c], cl); if (has_c) d[c] -= t; }
            if (act) {
#pragma unroll
                for (int c = 0; c < 6; c++) ss = fma(d[c], d[c], ss);
            }
            double Tw[9];
            if (jac) {
                double Dr[9];
                const double w[3] = {w2.x, w2.y, w2.z}, jw[3] = {Jw2.x, Jw2.y, Jw2.z};
                const double isq = 1.0 / sq2;
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
                        const double wxJ = w[i1] * J[i2 * 3 + q] - w[i2] * J[i1 * 3 + q];
                        const double jwx = (q == i1) ? -jw[i2] : ((q == i2) ? jw[i1] : 0.0);
                        Dr[i * 3 + q] = sq2 * J[i * 3 + q] + wxJ - jwx - jw[i] * w[q] * isq;
                    }
                inv3(Dr, Dinv);
                const double c = 0.25 * dt * dt;
                Tw[0] = c * (w[0] * w[0] * isq + sq2); Tw[1] = c * (w[0] * w[1] * isq + w[2]); Tw[2] = c * (w[0] * w[2] * isq - w[1]);
                Tw[3] = c * (w[1] * w[0] * isq - w[2]); Tw[4] = c * (w[1] * w[1] * isq + sq2); Tw[5] = c * (w[1] * w[2] * isq + w[0]);
                Tw[6] = c * (w[2] * w[0] * isq + w[1]); Tw[7] = c * (w[2] * w[1] * isq - w[0]); Tw[8] = c * (w[2] * w[2] * isq + sq2);
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int q = 0; q < 3; q++) Mw[i * 3 + q] = fma(Tw[i * 3], Dinv[q], fma(Tw[i * 3 + 1], Dinv[3 + q], Tw[i * 3 + 2] * Dinv[6 + q]));
            }
            // --- parent's quantities
            double pp3[7], pd[6], pM[9];
#pragma unroll
            for (int i = 0; i < 7; i++) pp3[i] = shfl_d(p3[i], pl);
            if (jac) {
#pragma unroll
                for (int i = 0; i < 6; i++) pd[i] = shfl_d(d[i], pl);
#pragma unroll
                for (int i = 0; i < 9; i++) pM[i] = shfl_d(Mw[i], pl);
            }
            const V3 xa3 = has_p ? v3(pp3[0], pp3[1], pp3[2]) : v3(0.0, 0.0, 0.0);
            const Q4 qa3 = has_p ? Q4{pp3[3], pp3[4], pp3[5], pp3[6]} : Q4{1.0, 0.0, 0.0, 0.0};
            const V3 xb3 = v3(p3[0], p3[1], p3[2]);
            const Q4 qb3 = {p3[3], p3[4], p3[5], p3[6]};
            // --- my joint's rows
#pragma unroll
            for (int r = 0; r < MD; r++) {
                rhs[r] = 0.0;
                if (r < dim) {
                    double g;
                    V3 Gax, Gap, Gbx, Gbp;
                    row_eval((kindmask >> r) & 1, ldg3(M.row_proj + 3 * (r0 + r)), p1, p2, qoff, xa3, qa3, xb3, qb3, jac, g, Gax, Gap, Gbx, Gbp);
                    if (act) ss = fma(g, g, ss);
                    rhs[r] = g;
                    if (jac) {
                        const double cb = dt * kdt, ca = dt * pkdt;
                        double wB[6] = {cb * Gbx.x, cb * Gbx.y, cb * Gbx.z, 0, 0, 0}, wA[6] = {0, 0, 0, 0, 0, 0};
                        { V3 t = vec_mat3(Gbp, Mw); wB[3] = t.x; wB[4] = t.y; wB[5] = t.z; }
                        double acc = g;
#pragma unroll
                        for (int c = 0; c < 6; c++) acc = fma(-wB[c], d[c], acc);
                        if (has_p) {
                            wA[0] = ca * Gax.x; wA[1] = ca * Gax.y; wA[2] = ca * Gax.z;
                            V3 t = vec_mat3(Gap, pM); wA[3] = t.x; wA[4] = t.y; wA[5] = t.z;
#pragma unroll
                            for (int c = 0; c < 6; c++) acc = fma(-wA[c], pd[c], acc);
                        }
                        rhs[r] = acc;
#pragma unroll
                        for (int t = 0; t < MD; t++) {
                            double sjj = 0.0, sjp = 0.0, sjc = 0.0;
#pragma unroll
                            for (int c = 0; c < 6; c++) {
                                sjj = fma(wA[c], LSM(smem, L::GKA + t * 6 + c), fma(wB[c], LSM(smem, L::GKB + t * 6 + c), sjj));
                                sjp = fma(wA[c], LSM(smem, L::PGB + t * 6 + c), sjp);
                                sjc = fma(wB[c], LSM(smem, L::CGA + t * 6 + c), sjc);
                            }
                            if (t == r) sjj += M.reg;
                            LSM(smem, L::SJJ + r * MD + t) = sjj;
                            LSM(smem, L::SJP + r * MD + t) = sjp;
                            LSM(smem, L::SJC + r * MD + t) = sjc;
                        }
                    }
                } else if (jac) {   // padding rows: identity on the diagonal, no coupling
#pragma unroll
                    for (int t = 0; t < MD; t++) {
                        LSM(smem, L::SJJ + r * MD + t) = (t == r) ? 1.0 : 0.0;
                        LSM(smem, L::SJP + r * MD + t) = 0.0;
                        LSM(smem, L::SJC + r * MD + t) = 0.0;
                    }
                }
            }
            return sqrt(sys_sum(ss));
        };

        // leaf -> root sweep + root -> leaf back substitution + body steps; on exit DL/DS hold the Newton step
        auto solve = [&](bool act) {
            __syncwarp();
            for (int lev = max_depth; lev >= 0; lev--) {
                const bool mine = act && depth == lev;
                // (1) absorb my child's eliminated block:  S_jj -= S_jc X_c[:, :dim] ; rhs -= S_jc y_c   (X_c sits in the child's SJP / rhs)
                if (lev < max_depth) {
#pragma unroll
                    for (int kk = 0; kk < MD; kk++) {
                        const double yk = shfl_d(rhs[kk], cl);
                        double xrow[MD];
#pragma unroll
                        for (int t = 0; t < MD; t++) xrow[t] = smem[(L::SJP + kk * MD + t) * 32 + cl];
                        if (mine && has_c) {
#pragma unroll
                            for (int r = 0; r < MD; r++) {
                                const double f = LSM(smem, L::SJC + r * MD + kk);
                                rhs[r] = fma(-f, yk, rhs[r]);
#pragma unroll
                                for (int t = 0; t < MD; t++) LSM(smem, L::SJJ + r * MD + t) = fma(-f, xrow[t], LSM(smem, L::SJJ + r * MD + t));
                            }
                        }
                    }
                }
                // (2) factorise my block (LU without pivoting) and solve [S_jp | rhs]
                if (mine) {
                    double Am[MD][MD], invd[MD];
#pragma unroll
                    for (int r = 0; r < MD; r++)
#pragma unroll
                        for (int t = 0; t < MD; t++) Am[r][t] = LSM(smem, L::SJJ + r * MD + t);
#pragma unroll
                    for (int kk = 0; kk < MD; kk++) {
                        invd[kk] = 1.0 / Am[kk][kk];
#pragma unroll
                        for (int r = kk + 1; r < MD; r++) {
                            Am[r][kk] *= invd[kk];
#pragma unroll
                            for (int t = kk + 1; t < MD; t++) Am[r][t] = fma(-Am[r][kk], Am[kk][t], Am[r][t]);
                        }
                    }
#pragma unroll
                    for (int col = 0; col <= MD; col++) {
                        if (col < MD && !has_p) continue;
                        double x[MD];
#pragma unroll
                        for (int r = 0; r < MD; r++) x[r] = col < MD ? LSM(smem, L::SJP + r * MD + col) : rhs[r];
#pragma unroll
                        for (int kk = 0; kk < MD; kk++)
#pragma unroll
                            for (int r = kk + 1; r < MD; r++) x[r] = fma(-Am[r][kk], x[kk], x[r]);
#pragma unroll
                        for (int kk = MD - 1; kk >= 0; kk--) {
#pragma unroll
                            for (int t = kk + 1; t < MD; t++) x[kk] = fma(-Am[kk][t], x[t], x[kk]);
                            x[kk] *= invd[kk];
                        }
#pragma unroll
                        for (int r = 0; r < MD; r++) {
                            if (col < MD) LSM(smem, L::SJP + r * MD + col) = x[r];
                            else rhs[r] = x[r];
                        }
                    }
                }
                __syncwarp();
            }
            // back substitution root -> leaf: dl = y - U dl_parent
            for (int lev = 0; lev <= max_depth; lev++) {
                double pdl[MD];
#pragma unroll
                for (int t = 0; t < MD; t++) pdl[t] = shfl_d(rhs[t], pl);
                if (act && depth == lev && has_p) {
#pragma unroll
                    for (int r = 0; r < MD; r++) {
                        double x = rhs[r];
#pragma unroll
                        for (int t = 0; t < MD; t++) x = fma(-LSM(smem, L::SJP + r * MD + t), pdl[t], x);
                        rhs[r] = x;
                    }
                }
            }
            // body step ds = D^-1 (d + GkB^T dl + [child] GkA_child^T dl_child)
            double t6[6], tA[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < 6; c++) t6[c] = d[c];
#pragma unroll
            for (int r = 0; r < MD; r++) {
                const double l = rhs[r];
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    t6[c] = fma(LSM(smem, L::GKB + r * 6 + c), l, t6[c]);
                    tA[c] = fma(LSM(smem, L::GKA + r * 6 + c), l, tA[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 6; c++) { const double t = shfl_d(tA[c], cl); if (has_c) t6[c] += t; }
            if (act) {
                const V3 dw = mat3_mul(Dinv, v3(t6[3], t6[4], t6[5]));
                LSM(smem, L::DS + 0) = kdt * t6[0]; LSM(smem, L::DS + 1) = kdt * t6[1]; LSM(smem, L::DS + 2) = kdt * t6[2];
                LSM(smem, L::DS + 3) = dw.x; LSM(smem, L::DS + 4) = dw.y; LSM(smem, L::DS + 5) = dw.z;
#pragma unroll
                for (int r = 0; r < MD; r++) LSM(smem, L::DL + r) = rhs[r];
            }
        };

        bool done = !valid;
        double normf0 = 0.0;
        int it_used = 0;
        for (int n = 1; n <= A.newton_iter; n++) {
            if (__all_sync(0xffffffffu, done)) break;
            const bool act = !done;
            const double nf = evaluate(true, act);
            if (n == 1) normf0 = nf;
            solve(act);
            double sd = 0.0;
            if (act) {
#pragma unroll
                for (int i = 0; i < 6; i++) { const double v = LSM(smem, L::DS + i); sd = fma(v, v, sd); LSM(smem, L::S0 + i) = s[i]; }
#pragma unroll
                for (int r = 0; r < MD; r++) { const double v = LSM(smem, L::DL + r); sd = fma(v, v, sd); LSM(smem, L::L0 + r) = lam[r]; }
            }
            const double full = sqrt(sys_sum(sd));
            const bool tiny = act && full < A.eps && normf0 < A.eps;
            double alpha = 1.0, normf1 = normf0;
            bool ls_done = !act || tiny;
            if (tiny) {
#pragma unroll
                for (int i = 0; i < 6; i++) s[i] = LSM(smem, L::S0 + i) - LSM(smem, L::DS + i);
#pragma unroll
                for (int r = 0; r < MD; r++) lam[r] = LSM(smem, L::L0 + r) - LSM(smem, L::DL + r);
            }
            for (int ls = 0; ls <= A.line_iter; ls++) {
                if (__all_sync(0xffffffffu, ls_done)) break;
                const bool la = !ls_done;
                if (la) {
#pragma unroll
                    for (int i = 0; i < 6; i++) s[i] = fma(-alpha, LSM(smem, L::DS + i), LSM(smem, L::S0 + i));
#pragma unroll
                    for (int r = 0; r < MD; r++) lam[r] = fma(-alpha, LSM(smem, L::DL + r), LSM(smem, L::L0 + r));
                }
                const double nf1 = evaluate(false, la);
                if (la) {
                    normf1 = nf1;
                    if (nf1 >= normf0 && ls < A.line_iter) alpha *= 0.5;
                    else ls_done = true;
                }
            }
            if (act) {
                it_used = n;
                const double nd = alpha * full;
                if (tiny || (nd < A.eps && normf1 < A.eps)) done = true;
                if (!(normf1 == normf1)) { done = true; failed = true; }   // NaN residual: no iterate can recover, the step has failed
                normf0 = normf1;
            }
        }
        if (valid) {
            total_iters += it_used;
            if (!done) failed = true;
        }
        if (!A.knot_input) {
#pragma unroll
            for (int i = 0; i < 6; i++) z[7 + i] = s[i];      // updatestate!: (xc,qc,vc,wc) <- (x2,q2,v2,w2)
        }
    }
    // a failure anywhere in the system marks the system
    const double anyfail = sys_sum(failed ? 1.0 : 0.0);
    if (valid && A.knot_input) {
#pragma unroll
        for (int i = 0; i < 6; i++) A.s_out[(size_t)sys * 6 * M.nb + 6 * b + i] = s[i];
        for (int r = 0; r < dim; r++) A.lambda[(size_t)sys * M.ml + r0 + r] = lam[r];
    } else if (valid) {
#pragma unroll
        for (int i = 0; i < 13; i++) A.state[(size_t)sys * nst + 13 * b + i] = z[i];
        for (int r = 0; r < dim; r++) A.lambda[(size_t)sys * M.ml + r0 + r] = lam[r];
    }
    if (valid && A.status && j == 0) write_status(A, sys, total_iters, anyfail > 0.0);
}

}  // namespace

// Eligibility + tables for the chain kernel: every eqc holds one joint, every body is the child of exactly one joint and
// the parent of at most one, eqc e's joint is joint e, at most 32 joints.
bool chain_tables(const HostMech& H, std::vector<int>& parent, std::vector<int>& child, std::vector<int>& depth, int& max_depth, int& max_dim) {
    if (H.nj != H.ne || H.nj != H.nb || H.nj < 1) return false;   // (K1c itself needs nj <= 32: checked at launch)
    std::vector<int> joint_of_body(H.nb, -1), kids(H.nb, 0);
    max_dim = 0;
    for (int j = 0; j < H.nj; j++) {
        if (H.eqc_ptr[j] != j) return false;
        if (joint_of_body[H.j_child[j]] >= 0) return false;
        joint_of_body[H.j_child[j]] = j;
        if (H.j_parent[j] >= 0 && ++kids[H.j_parent[j]] > 1) return false;
        max_dim = std::max(max_dim, H.j_dim[j]);
    }
    parent.assign(H.nj, -1); child.assign(H.nj, -1); depth.assign(H.nj, 0);
    for (int j = 0; j < H.nj; j++) {
        if (H.j_parent[j] >= 0) {
            int p = joint_of_body[H.j_parent[j]];
            if (p < 0) return false;
            parent[j] = p; child[p] = j;
        }
    }
    max_depth = 0;
    for (int j = 0; j < H.nj; j++) {
        int d = 0, q = j, guard = 0;
        while (parent[q] >= 0 && guard++ <= H.nj) { q = parent[q]; d++; }
        if (guard > H.nj) return false;
        depth[j] = d; max_depth = std::max(max_depth, d);
    }
    return true;
}

int launch_rollout_chain(const DevMech& M, const DevCtrl& C, const SimArgs& a, const int* d_parent, const int* d_child, const int* d_depth,
                         int max_depth, int max_dim, cudaStream_t stream) {
    // Systems per warp: the kernel is latency bound per warp (one warp issues ~0.2 instr/cycle), so a small batch is spread
    // over partly filled warps -- but never over more CTAs than are resident at once (shared memory bounds that), because a
    // second wave doubles the time.
    int dev = 0, sms = 148, resident = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem_cta = (size_t)(max_dim <= 5 ? ChainSmem<5>::TOTAL : ChainSmem<6>::TOTAL) * 32 * sizeof(double);
    resident = std::max(1, (int)((227 * 1024) / (smem_cta + 1024)));
    const int spw_max = 32 / M.nj;
    const long long slots = (long long)sms * resident;
    int spw = a.spw > 0 ? a.spw : (int)((a.B + slots - 1) / slots);
    spw = std::max(1, std::min(spw, spw_max));
    const int grid = (a.B + spw - 1) / spw;
    if (max_dim <= 5) {
        const size_t smem = (size_t)ChainSmem<5>::TOTAL * 32 * sizeof(double);
        CCB_CUDA(cudaFuncSetAttribute(rollout_chain_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rollout_chain_kernel<5><<<grid, 32, smem, stream>>>(M, C, a, d_parent, d_child, d_depth, max_depth, spw);
    } else {
        const size_t smem = (size_t)ChainSmem<6>::TOTAL * 32 * sizeof(double);
        CCB_CUDA(cudaFuncSetAttribute(rollout_chain_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rollout_chain_kernel<6><<<grid, 32, smem, stream>>>(M, C, a, d_parent, d_child, d_depth, max_depth, spw);
    }
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
