// csrc/ccb_rollout_rows.cu -- K1r: fused closed-loop rollout for short serial chains (pendulum, cartpole, acrobot,
// double pendulum, triple cartpole: up to 6 joints of dimension <= 5), the headline kernel of bench.py.
//
// Mapping: ONE LANE owns ONE SCALAR CONSTRAINT ROW (one lambda component).  A joint is a group of 5 consecutive lanes, a
// system of NJ joints is 5*NJ consecutive lanes in chain order (root first) and a warp holds floor(32 / (5 NJ)) systems.
//   * the 5 lanes of a group keep the joint's child body (pose, velocities, wrench, 3x3 blocks) redundantly in registers,
//     so body quantities need no exchange inside the group;
//   * a lane evaluates its own constraint row, its row of Gv D^-1 and its row of the constraint Schur complement
//     S = Gv D^-1 G^T, which is block tridiagonal for a chain: [parent block | own block | child block | rhs] = 16 registers;
//   * S is factorised with ONE ROW PER LANE: banded elimination from the leaf row to the root row (ldu_factorization!
//     role, no fill outside the band), then substitution root -> leaf (ldu_backsubstitution! role); a pivot costs one
//     broadcast of the pivot row (warp shuffles) and one FMA per entry in the lanes of the two blocks it touches;
//   * the only shared memory is the table of Jacobian rows at the knot (fixed during a step, read with broadcast loads when
//     the rows of S are formed) and the joint constants.
// Compared with K1c (one lane per joint, ccb_rollout_chain.cu) the dependent instruction chain of a step is ~5x shorter,
// which is what bounds a 4096-system batch on 148 SMs; compared with K1 (ccb_rollout.cu) nothing is interpreted from
// index tables.
//
// Same algorithm and semantics as K1 / K1c: simulate! loop of ConstrainedDynamics.jl with the
// control laws of reference src/control/lqr.jl:89-139, lqr_tracking.jl:46-71, pid.jl:43-88 fused.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "ccb_dev_math.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

constexpr int kGS = 70;   // doubles per joint group in the knot-row table: 5 rows x 12, padded so that (a) the double2 row loads of the 6 groups of a
                          // warp hit distinct 16-byte bank groups (35 g mod 8) and (b) the 5-lane column reads of the force sums barely overlap (6 g mod 16)
constexpr int kCS = 24;   // doubles of constants per joint: p1 3, p2 3, qoff 4, axis 3, J 9, mass 1 (+1)

struct RowsOrder { int joint[6]; };

// 168 registers: twelve 1-warp CTAs per SM, so the 1366 CTAs of the 4096-system cartpole batch are resident in ONE wave
// (measured on the 4096x2000 cartpole batch: 128 -> 42.5 ms, 144 -> 40.3, 168 -> 37.9, 184 -> 48.7, 200 -> 48.1, 255 -> 43.3)
#ifndef K1R_MAXREG
#define K1R_MAXREG 168
#endif
// the persistent variant runs at most two warps per scheduler: all 255 registers fit (8 x 32 x 256 = 64 K)
#ifndef K1R_PERSIST_MAXREG
#define K1R_PERSIST_MAXREG 255
#endif

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// 1/x for the LU pivots: hardware approximation refined by two Newton steps (relative error ~1e-16, no slow path)
__device__ __forceinline__ double rcp_fast(double x) {
#ifdef K1R_EXACT_DIV
    return 1.0 / x;
#else
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#endif
}

// GEN: false = the LQR-family instantiation (no controller, LQR Inf / finite, TrackingLQR): the PID and open-loop-table code (joint coordinate,
// atan2, integrator state) is compiled out, which shortens the hot loop's SASS and frees registers; true = every law.
// PERSIST: the time-sliced persistent variant (work loop below); a separate instantiation so that the plain kernel keeps its register allocation.
template <int NJ, bool GEN, bool PERSIST>
__global__ void __maxnreg__(PERSIST ? K1R_PERSIST_MAXREG : K1R_MAXREG) rollout_rows_kernel(const DevMech M, const DevCtrl C, const SimArgs A, const RowsOrder ord) {
    extern __shared__ __align__(16) double smem[];
    constexpr int G = 5 * NJ, SPW = 32 / G;
    double* gk = smem;                      // [SPW][NJ][kGS]: row t of a joint at t*12: GkA[6] | GkB[6]
    double* cst = smem + SPW * NJ * kGS;    // [NJ][kCS]
    double* stash = cst + NJ * kCS;         // [SPW][NJ][17]: body residual d (6) and inverse rotational block Dinv (9) between evaluation and solve
    const int wl = threadIdx.x;
    const int sw = wl / G;
    const int q = wl - sw * G;
    const int pos = q / 5, r = q - 5 * pos;         // joint position in the chain (0 = root), row in the joint
    const bool lane_ok = sw < SPW;
    const int base = lane_ok ? sw * G : 0;           // first lane of my system (idle lanes shadow system slot 0, read only)
    const int g0 = base + 5 * pos;                   // first lane of my joint group
    const bool has_p = pos > 0, has_c = pos < NJ - 1;
    const int gp = has_p ? g0 - 5 : g0, gc = has_c ? g0 + 5 : g0;   // parent / child group (self when absent)
    const int plane = has_p ? g0 - 5 + r : (lane_ok ? wl : g0 + r); // lane of the parent group holding the parent body
    const int jt = ord.joint[pos];
    const int cjt = has_c ? ord.joint[pos + 1] : -1;
    const int b = __ldg(M.j_child + jt);
    const int dim = __ldg(M.j_dim + jt), r0 = __ldg(M.j_row0 + jt);
    const int rr = r < dim ? r : 0;
    const int kind = __ldg(M.row_kind + r0 + rr);
    const V3 pr = ldg3(M.row_proj + 3 * (r0 + rr));
    const int act_kind = __ldg(M.eqc_act_kind + jt);       // chain mechanisms: eqc index == joint index (chain_tables)
    const int cact_kind = has_c ? __ldg(M.eqc_act_kind + cjt) : 0;
    const double dt = M.dt, idt = 1.0 / M.dt, c4 = 4.0 / (M.dt * M.dt);
    const int nst = 13 * M.nb;
    const int swz = lane_ok ? sw : 0;
    double* gk_own = gk + (swz * NJ + pos) * kGS;
    const double* gk_par = gk + (swz * NJ + (has_p ? pos - 1 : pos)) * kGS;
    const double* gk_chi = gk + (swz * NJ + (has_c ? pos + 1 : pos)) * kGS;
    const double* cj = cst + pos * kCS;
    const double* ccj = cst + (has_c ? pos + 1 : pos) * kCS;
    double* st_own = stash + (swz * NJ + pos) * 17;     // stride 17: the groups of a warp land in distinct banks

    // ---- joint constants -> shared memory (one copy per joint position, shared by the systems of the warp)
    for (int i = wl; i < NJ * kCS; i += 32) {
        const int p = i / kCS, k = i - p * kCS, j = ord.joint[p], bb = __ldg(M.j_child + j);
        double v = 0.0;
        if (k < 3) v = __ldg(M.j_p1 + 3 * j + k);
        else if (k < 6) v = __ldg(M.j_p2 + 3 * j + k - 3);
        else if (k < 10) v = __ldg(M.j_qoff + 4 * j + k - 6);
        else if (k < 13) v = __ldg(M.j_axis + 3 * j + k - 10);
        else if (k < 22) v = __ldg(M.inertia + 9 * bb + k - 13);
        else if (k == 22) v = __ldg(M.mass + bb);
        cst[i] = v;
    }
    __syncwarp();
    const double mass = cj[22], kdt = dt / mass;
    const double pkdt = has_p ? dt / cst[(pos - 1) * kCS + 22] : 0.0;

    // ---- work loop.  Plain launch: ONE pass, the warp owns the systems blockIdx.x * SPW .. for all A.steps steps.
    // Persistent launch (PERSIST, launch_rollout_rows): the grid is exactly two warps per SM sub-partition (W workers) for a batch of
    // ntrip = W + extra warp-loads ("triples" of systems), 0 < extra < 0.3 W.  A plain launch of such a batch runs at the speed of the schedulers
    // that got three warps while the others idle a third of the time (4096 cartpoles = 1366 warps = 2.3 per scheduler).  Here the rollout is cut
    // into nchunks chunks of steps and the schedule is STATIC: worker w simulates its own triple w chunk after chunk (no dependency on anybody)
    // and, after its chunk c, one chunk of an EXTRA triple when its turn comes -- extra triple W + j, chunk c goes to worker (c extra + j) mod W,
    // so the extra work rotates over the workers (everybody ends up with nchunks + nchunks extra / W +- 1 items) and the chunk an extra item
    // continues was finished one whole round earlier by another worker: the flag it checks (release / acquire through __threadfence) is
    // practically always set.  (A first version handed (chunk, triple) items out from a global counter: workers picked up chunks whose
    // predecessor was still running elsewhere and spun, the spinning warps stole issue slots from their scheduler mates, the convoy collapsed
    // to 170 ms instead of 31 ms -- tools/persist_probe.py.)
    const int ntrip = (A.B + SPW - 1) / SPW;
    const int nworkers = gridDim.x, extra = PERSIST ? ntrip - nworkers : 0;
    for (int slot = 0;; slot++) {
    int trip = blockIdx.x, chunk = 0;
    if (PERSIST) {
        chunk = slot >> 1;
        if (chunk >= A.nchunks) break;
        if (slot & 1) {
            int j = ((int)blockIdx.x - chunk * extra) % nworkers;
            if (j < 0) j += nworkers;
            if (j >= extra) continue;
            trip = nworkers + j;
            if (chunk > 0) {
                // Every lane polls the same word (one broadcast load) and the exit is a warp VOTE, so the wait is warp-uniform by construction.
                // A lane-0-only spin (`if (wl == 0) while (*flag < chunk) ...; __syncwarp()`) made ptxas give up on proving the warp converged
                // for the rest of the kernel: every later shfl.sync got a BRA.DIV-guarded slow path -- 804 instead of 402 SHFL, 1934 instead of
                // 312 MOV, 7888 instead of 4496 instructions in the cartpole instantiation -- and the persistent launch ran 7 % slower per warp
                // than the plain kernel at the same register count (3552 systems: 24.7 against 23.1 ms).
                const volatile int* flag = A.persist + 1 + trip;
                while (!__all_sync(0xffffffffu, *flag >= chunk)) __nanosleep(2000);
                __threadfence();
            }
        }
    } else if (slot > 0) {
        break;
    }
    const int step_lo = PERSIST ? chunk * A.chunk_steps : 0;
    const int step_hi = PERSIST ? min(A.steps, step_lo + A.chunk_steps) : A.steps;
    const int sys = trip * SPW + sw;
    const bool valid = lane_ok && sys < A.B;
    const bool rowpad = !valid || r >= dim;          // padding row: identity in S, zero everywhere else

    // ---- state in registers: the group's body (redundant in its 5 lanes) and my multiplier
    double z[13], lam;      // z[7..12]: (v1, w1) until the step's constants are formed, then the accepted Newton iterate (v2, w2)
    if (valid) {
        // (__ldcg: in a persistent launch the previous chunk of these systems was written by another SM; L1 is not coherent)
#pragma unroll
        for (int i = 0; i < 13; i++) z[i] = PERSIST ? __ldcg(A.state + (size_t)sys * nst + 13 * b + i) : A.state[(size_t)sys * nst + 13 * b + i];
        lam = r < dim ? (PERSIST ? __ldcg(A.lambda + (size_t)sys * M.ml + r0 + r) : A.lambda[(size_t)sys * M.ml + r0 + r]) : 0.0;
    } else {
#pragma unroll
        for (int i = 0; i < 13; i++) z[i] = i == 3 ? 1.0 : 0.0;
        lam = 0.0;
    }

    // sum over the lanes of my system, evaluated in one fixed order so that every lane holds the same bits
    auto sys_sum = [&](double v) {
        const double v0 = shfl_d(v, g0), v1 = shfl_d(v, g0 + 1), v2 = shfl_d(v, g0 + 2), v3_ = shfl_d(v, g0 + 3), v4 = shfl_d(v, g0 + 4);
        const double gs = ((v0 + v1) + (v2 + v3_)) + v4;
        double acc = 0.0;
#pragma unroll
        for (int p = 0; p < NJ; p++) acc += shfl_d(gs, base + 5 * p);
        return acc;
    };

    int total_iters = 0;
    bool failed = false;

    for (int step = step_lo; step < step_hi; step++) {
        const int k = A.k0 + step;
        if (valid && A.storage) {   // saveToStorage!: (xc,qc,vc,wc) at the start of step k; the 5 lanes interleave the 13 words
            // (Measured r2: staging the warp's contiguous 13 nb SPW doubles through the stash and streaming them out with consecutive lanes on
            // consecutive doubles is SLOWER -- 34.7 ms against 32.9 ms with recording, and 32.6 against 31.4 ms without, because it shifts the
            // register allocation at the 168 cap; the stores below already fill whole sectors, the five lanes of a joint cover 13 adjacent
            // words, and at 1.7 GB per 33 ms the write stream is 0.8 % of the HBM roof.)
            double* out = A.storage + ((size_t)step * A.B + sys) * nst + 13 * b;
#pragma unroll
            for (int i = 0; i < 13; i++)
                if (i % 5 == r) out[i] = z[i];
        }
        // ---------------- control!(k) part 1: laws that read the continuous state (PID), open-loop tables
        double u = 0.0;
        if (GEN && (C.kind == CCB_CTRL_PID || C.kind == CCB_CTRL_TABLE)) {
            double pz[7];
#pragma unroll
            for (int i = 0; i < 7; i++) pz[i] = shfl_d(z[i], plane);
            if (valid && r == 0) {
                for (int i = 0; i < C.nu; i++) {
                    if (__ldg(C.eqc + i) != jt) continue;
                    if (C.kind == CCB_CTRL_TABLE) {
                        if (k >= 1 && k <= C.N) u += __ldg(C.U + (size_t)sys * C.U_bs + (size_t)(k - 1) * C.nu + i);
                        continue;
                    }
                    V3 xa = {0.0, 0.0, 0.0};
                    Q4 qa = {1.0, 0.0, 0.0, 0.0};
                    if (has_p) { xa = v3(pz[0], pz[1], pz[2]); qa = Q4{pz[3], pz[4], pz[5], pz[6]}; }
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
            const V3 v = v3(z[7], z[8], z[9]), w = v3(z[10], z[11], z[12]);
            z[0] = fma(dt, v.x, z[0]); z[1] = fma(dt, v.y, z[1]); z[2] = fma(dt, v.z, z[2]);
            const double sq = sqrt(c4 - dot(w, w));
            const Q4 q2 = qmul(Q4{z[3], z[4], z[5], z[6]}, Q4{0.5 * dt * sq, 0.5 * dt * w.x, 0.5 * dt * w.y, 0.5 * dt * w.z});
            z[3] = q2.s; z[4] = q2.x; z[5] = q2.y; z[6] = q2.z;
        }
        // parent's knot pose (fixed for the step)
        V3 xa2 = {0.0, 0.0, 0.0};
        Q4 qa2 = {1.0, 0.0, 0.0, 0.0};
        {
            double t[7];
#pragma unroll
            for (int i = 0; i < 7; i++) t[i] = shfl_d(z[i], plane);
            if (has_p) { xa2 = v3(t[0], t[1], t[2]); qa2 = Q4{t[3], t[4], t[5], t[6]}; }
        }
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
                // lane r < 4 owns one 3-vector of the body's Δz = [x2 - xd; v1 - vd; imag(qd \ q2); w1 - wd]
                V3 dz = {0.0, 0.0, 0.0};
                if (valid && r < 4) {
                    const double* d = zd + 13 * b;
                    if (r == 0) dz = v3(z[0] - __ldg(d), z[1] - __ldg(d + 1), z[2] - __ldg(d + 2));
                    else if (r == 1) dz = v3(z[7] - __ldg(d + 7), z[8] - __ldg(d + 8), z[9] - __ldg(d + 9));
                    else if (r == 2) dz = qvec(qmul(qconj(ldg4(d + 3)), Q4{z[3], z[4], z[5], z[6]}));
                    else dz = v3(z[10] - __ldg(d + 10), z[11] - __ldg(d + 11), z[12] - __ldg(d + 12));
                }
                for (int i = 0; i < C.nu; i++) {
                    double part = 0.0;
                    if (valid && r < 4) {
                        const double* Kr = K + (size_t)i * n + 12 * b + 3 * r;
                        part = fma(__ldg(Kr), dz.x, fma(__ldg(Kr + 1), dz.y, __ldg(Kr + 2) * dz.z));
                    }
                    const double tot = sys_sum(part);
                    if (valid && r == 0 && __ldg(C.eqc + i) == jt) u += __ldg(Fd + i) - tot;
                }
            }
        }
        u = shfl_d(u, g0);                       // the joint's input, known to its 5 lanes
        double uc = shfl_d(u, gc);               // the child joint's input
        if (!has_c) uc = 0.0;
        // ---------------- applyFτ!: my body gets the child side of my joint's input and the parent side of my child's
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
        // λ-free part of the body residual that does not change during the Newton iteration
        double c0[6];
        {
            const V3 v1 = v3(z[7], z[8], z[9]), w1 = v3(z[10], z[11], z[12]);
            const double sq1 = sqrt(c4 - dot(w1, w1));
            const V3 Jw1 = mat3_mul(cj + 13, w1);
            const V3 cR = cross(w1, Jw1) - sq1 * Jw1 - 2.0 * Text;
            c0[0] = -mass * v1.x * idt - Fext.x; c0[1] = -mass * v1.y * idt - Fext.y; c0[2] = -mass * (v1.z * idt + M.g) - Fext.z;
            c0[3] = cR.x; c0[4] = cR.y; c0[5] = cR.z;
        }
        // ---------------- my Jacobian row at the knot (it carries the constraint force; fixed for the step)
        {
            double g;
            V3 Gax, Gap, Gbx, Gbp;
            row_eval(kind, pr, ld3(cj), ld3(cj + 3), ld4(cj + 6), xa2, qa2, v3(z[0], z[1], z[2]), Q4{z[3], z[4], z[5], z[6]}, true, g, Gax, Gap, Gbx, Gbp);
            if (!has_p) { Gax = v3(0.0, 0.0, 0.0); Gap = Gax; }
            if (rowpad) { Gax = v3(0.0, 0.0, 0.0); Gap = Gax; Gbx = Gax; Gbp = Gax; }
            __syncwarp();
            if (lane_ok) {
                double* o = gk_own + 12 * r;
                st3(o, Gax); st3(o + 3, Gap); st3(o + 6, Gbx); st3(o + 9, Gbp);
            }
            __syncwarp();
        }
        // transposed copies for the force sums: lane r owns component r of G_own,B^T λ_own + G_child,A^T λ_child

        // ---------------- newton! with lineSearch!
        double rowP[5], rowO[5], rowC[5], rhs;     // my row of S: parent block | own block | child block | rhs
        // G^T-type sum over the rows acting on my body: own joint (child side) + child joint (parent side)
        auto force_sum = [&](double mine, double* out6) {
            double f = 0.0, f5 = 0.0;
#pragma unroll
            for (int t = 0; t < 5; t++) {
                const double lo = shfl_d(mine, g0 + t), lc = shfl_d(mine, gc + t);
                f = fma(gk_own[12 * t + 6 + r], lo, f);                 // component r (lanes 0..4 of the group)
                if (has_c) f = fma(gk_chi[12 * t + r], lc, f);
                f5 = fma(gk_own[12 * t + 11], lo, f5);                 // sixth component: every lane, table read (broadcast)
                if (has_c) f5 = fma(gk_chi[12 * t + 5], lc, f5);
            }
#pragma unroll
            for (int c = 0; c < 5; c++) out6[c] = shfl_d(f, g0 + c);
            out6[5] = f5;
        };
        double dl = 0.0, ds[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // pending Newton step
        double aeff = 0.0;                                        // line-search scale of the pending step (0: evaluate the accepted iterate)
        auto evaluate = [&](bool jac) -> double {
            const V3 v2 = v3(fma(-aeff, ds[0], z[7]), fma(-aeff, ds[1], z[8]), fma(-aeff, ds[2], z[9]));
            const V3 w2 = v3(fma(-aeff, ds[3], z[10]), fma(-aeff, ds[4], z[11]), fma(-aeff, ds[5], z[12]));
            const double lamt = fma(-aeff, dl, lam);
            double p3[7], Dinv[9], Mw[9], d[6];
            const double sq2 = sqrt(c4 - dot(w2, w2));
            {
                p3[0] = fma(dt, v2.x, z[0]); p3[1] = fma(dt, v2.y, z[1]); p3[2] = fma(dt, v2.z, z[2]);
                const Q4 q3 = qmul(Q4{z[3], z[4], z[5], z[6]}, Q4{0.5 * dt * sq2, 0.5 * dt * w2.x, 0.5 * dt * w2.y, 0.5 * dt * w2.z});
                p3[3] = q3.s; p3[4] = q3.x; p3[5] = q3.y; p3[6] = q3.z;
            }
            double J[9];
#pragma unroll
            for (int i = 0; i < 9; i++) J[i] = cj[13 + i];
            const V3 Jw2 = mat3_mul(J, w2);
            const V3 dR = sq2 * Jw2 + cross(w2, Jw2);
            double fs[6];
            force_sum(lamt, fs);
            d[0] = fma(mass * idt, v2.x, c0[0]) - fs[0]; d[1] = fma(mass * idt, v2.y, c0[1]) - fs[1]; d[2] = fma(mass * idt, v2.z, c0[2]) - fs[2];
            d[3] = dR.x + c0[3] - fs[3]; d[4] = dR.y + c0[4] - fs[4]; d[5] = dR.z + c0[5] - fs[5];
            double ss = 0.0;
            if (valid && r == 0) {
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
            }
            // parent body's pose (and, for the Jacobian, its residual and Tw D^-1) from the parent group
            V3 xa3 = {0.0, 0.0, 0.0};
            Q4 qa3 = {1.0, 0.0, 0.0, 0.0};
            {
                double t[7];
#pragma unroll
                for (int i = 0; i < 7; i++) t[i] = shfl_d(p3[i], plane);
                if (has_p) { xa3 = v3(t[0], t[1], t[2]); qa3 = Q4{t[3], t[4], t[5], t[6]}; }
            }
            double g;
            V3 Gax, Gap, Gbx, Gbp;
            {
                const V3 p1 = ld3(cj), p2 = ld3(cj + 3);
                const Q4 qoff = ld4(cj + 6), qb3 = Q4{p3[3], p3[4], p3[5], p3[6]};
                const JointFrame jf = joint_frame(p2, qoff, xa3, qa3, v3(p3[0], p3[1], p3[2]), qb3);      // shared by the rows of the joint: no divergence
                row_finish(kind, pr, p1, p2, qoff, qa3, qb3, jf, jac, g, Gax, Gap, Gbx, Gbp);
            }
            if (rowpad) g = 0.0;
            ss = fma(g, g, ss);
            rhs = g;
            if (jac) {
                double pd[6], pM[9];
#pragma unroll
                for (int i = 0; i < 6; i++) pd[i] = shfl_d(d[i], plane);
#pragma unroll
                for (int i = 0; i < 9; i++) pM[i] = shfl_d(Mw[i], plane);
                // my row of Gv D^-1:  Gv = [G_x dt, G_phi Tw],  D = blkdiag(m/dt I, Drot)
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
                // my row of S against the rows of my joint, of the parent joint (through my parent body) and of the child joint
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    const double2* go = reinterpret_cast<const double2*>(gk_own + 12 * t);
                    double so = 0.0;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double2 a2 = go[c], b2 = go[3 + c];
                        so = fma(wA[2 * c], a2.x, fma(wA[2 * c + 1], a2.y, so));
                        so = fma(wB[2 * c], b2.x, fma(wB[2 * c + 1], b2.y, so));
                    }
                    rowO[t] = so;
                    if (NJ > 1) {
                        const double2* gpp = reinterpret_cast<const double2*>(gk_par + 12 * t + 6);
                        const double2* gcc = reinterpret_cast<const double2*>(gk_chi + 12 * t);
                        double sp = 0.0, sc = 0.0;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const double2 a2 = gpp[c], b2 = gcc[c];
                            sp = fma(wA[2 * c], a2.x, fma(wA[2 * c + 1], a2.y, sp));
                            sc = fma(wB[2 * c], b2.x, fma(wB[2 * c + 1], b2.y, sc));
                        }
                        rowP[t] = has_p ? sp : 0.0;
                        rowC[t] = has_c ? sc : 0.0;
                    } else {
                        rowP[t] = 0.0; rowC[t] = 0.0;
                    }
                }
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    if (t == r) rowO[t] += M.reg;
                    if (rowpad) { rowO[t] = (t == r) ? 1.0 : 0.0; rowP[t] = 0.0; rowC[t] = 0.0; }
                }
                if (rowpad) rhs = 0.0;
                // d and Dinv are needed again for the body step after the solve: park them in shared memory (3 words per lane)
                if (lane_ok) {
#pragma unroll
                    for (int e = 0; e < 15; e++)
                        if (e % 5 == r) st_own[e] = e < 6 ? d[e] : Dinv[e - 6];
                }
            }
            return sqrt(sys_sum(ss));
        };

        // S dλ = rhs with one row per lane.  Elimination runs from the last row of the leaf joint to the first row of the
        // root joint; a pivot row only has entries to its left (parent block + own block) and it updates the other rows of its own
        // block (Gauss-Jordan: the block ends up diagonal) and the rows of the parent block: no fill.  On exit `rhs` holds dλ.
        auto solve = [&](double& xdl, double* xds) {
            __syncwarp();
            double myinv = 1.0;
#pragma unroll
            for (int pp = NJ - 1; pp >= 0; pp--) {
#pragma unroll
                for (int rk = 4; rk >= 0; rk--) {
                    const int src = base + 5 * pp + rk;
                    double piv = shfl_d(rowO[rk], src);
                    if (!lane_ok) piv = 1.0;
                    const double inv = rcp_fast(piv);
                    if (pos == pp && r == rk) myinv = inv;
                    const bool same = (pos == pp) && (r != rk), par = (pos == pp - 1);
                    const double fS = same ? rowO[rk] * inv : 0.0;       // factor when I am another row of the pivot's block (Gauss-Jordan inside the block)
                    const double fP = par ? rowC[rk] * inv : 0.0;        // factor when I am a row of the pivot's parent block
#pragma unroll
                    for (int c = 0; c < rk; c++) {
                        const double pv = shfl_d(rowO[c], src);
                        rowO[c] = fma(-fS, pv, rowO[c]);
                        if (pp > 0) rowC[c] = fma(-fP, pv, rowC[c]);
                    }
                    if (pp > 0) {
#pragma unroll
                        for (int c = 0; c < 5; c++) {
                            const double pv = shfl_d(rowP[c], src);
                            rowP[c] = fma(-fS, pv, rowP[c]);
                            rowO[c] = fma(-fP, pv, rowO[c]);
                        }
                    }
                    const double pr_ = shfl_d(rhs, src);
                    rhs = fma(-(fS + fP), pr_, rhs);
                }
            }
            // substitution, root joint first: the blocks are diagonal after the Gauss-Jordan sweeps, so a joint's five multipliers come
            // out together: x = (rhs - (coupling to the parent block) x_parent) / pivot
#pragma unroll
            for (int pp = 0; pp < NJ; pp++) {
                double xs = rhs;
                if (pp > 0) {
#pragma unroll
                    for (int c = 0; c < 5; c++) {
                        const double xpar = shfl_d(rhs, base + 5 * (pp - 1) + c);
                        xs = fma(-rowP[c], xpar, xs);
                    }
                }
                if (pos == pp) rhs = xs * myinv;
            }
            xdl = rhs;
            // body step ds = D^-1 (d + G_own,B^T dλ_own + G_child,A^T dλ_child)
            double fs[6];
            force_sum(xdl, fs);
            const V3 dw = mat3_mul(st_own + 6, v3(st_own[3] + fs[3], st_own[4] + fs[4], st_own[5] + fs[5]));
            xds[0] = kdt * (st_own[0] + fs[0]); xds[1] = kdt * (st_own[1] + fs[1]); xds[2] = kdt * (st_own[2] + fs[2]);
            xds[3] = dw.x; xds[4] = dw.y; xds[5] = dw.z;
        };

        // newton! / lineSearch! as one loop around a single evaluation site.  The residual at an accepted line-search point is the
        // residual the next Newton iteration starts from, so every evaluation also forms the Jacobian rows: an iteration costs one
        // evaluation + one solve (instead of evaluation + solve + line-search evaluation), the iterates are unchanged.
        bool done = !valid, pending = false, first = true, out_of_iters = false;
        int n = 0, ls = 0;
        double alpha = 1.0, full = 0.0, normf0 = 0.0;
        for (;;) {
            aeff = pending ? alpha : 0.0;
            const double nfe = evaluate(true);
            bool want_solve = false;
            if (first) {
                normf0 = nfe;
                want_solve = !done;
            } else if (pending) {
                if (nfe >= normf0 && ls < A.line_iter) {          // lineSearch!: not decreased -> halve the step and try again
                    alpha *= 0.5; ls++;
                } else {                                          // accepted (or out of halvings): s = s_prev - alpha * step
                    pending = false;
#pragma unroll
                    for (int i = 0; i < 6; i++) z[7 + i] = fma(-alpha, ds[i], z[7 + i]);
                    lam = fma(-alpha, dl, lam);
                    if (alpha * full < A.eps && nfe < A.eps) done = true;
                    normf0 = nfe;
                    if (!done && n >= A.newton_iter) { done = true; out_of_iters = true; }
                    want_solve = !done;
                }
            }
            first = false;
            // NaN residual (e.g. |w| beyond 2/dt): lineSearch! has accepted it like the reference does (NaN >= x is false), no later iterate can
            // recover, so the step is abandoned as failed instead of burning newtonIter x lineIter evaluations
            if (!(nfe == nfe) && !done && !pending) { done = true; out_of_iters = true; want_solve = false; }
            if (!__any_sync(0xffffffffu, want_solve)) {
                if (!__any_sync(0xffffffffu, pending)) break;
                continue;
            }
            double dl_new, ds_new[6];
            solve(dl_new, ds_new);
            double sd = dl_new * dl_new;
            if (r == 0) {
#pragma unroll
                for (int i = 0; i < 6; i++) sd = fma(ds_new[i], ds_new[i], sd);
            }
            if (!valid) sd = 0.0;
            const double fullv = sqrt(sys_sum(sd));
            if (want_solve) {
                n++;
                full = fullv; dl = dl_new;
#pragma unroll
                for (int i = 0; i < 6; i++) ds[i] = ds_new[i];
                alpha = 1.0; ls = 0;
                // a step below eps taken from a residual below eps ends the solve whatever scale the line search would pick
                if (full < A.eps && normf0 < A.eps) {
#pragma unroll
                    for (int i = 0; i < 6; i++) z[7 + i] -= ds[i];
                    lam -= dl;
                    done = true;
                } else {
                    pending = true;
                }
            }
            if (!__any_sync(0xffffffffu, pending)) break;
        }
        const int it_used = n;
        if (out_of_iters) done = false;
        if (valid) {
            total_iters += it_used;
            if (!done) failed = true;
        }
        // updatestate!: (xc,qc,vc,wc) <- (x2,q2,v2,w2): z[0..6] already holds the knot, z[7..12] the converged (v2, w2)
    }
    if (valid && A.knot_input) {
        if (r == 0) {
#pragma unroll
            for (int i = 0; i < 6; i++) A.s_out[(size_t)sys * 6 * M.nb + 6 * b + i] = z[7 + i];
        }
        if (r < dim) A.lambda[(size_t)sys * M.ml + r0 + r] = lam;
    } else if (valid) {
#pragma unroll
        for (int i = 0; i < 13; i++)
            if (i % 5 == r) A.state[(size_t)sys * nst + 13 * b + i] = z[i];
        if (r < dim) A.lambda[(size_t)sys * M.ml + r0 + r] = lam;
    }
    if (valid && A.status && q == 0) write_status(A, sys, total_iters, failed, PERSIST && chunk > 0);
    if (PERSIST) {
        if (trip >= nworkers) __threadfence();      // an extra triple moves on to another worker: release my stores, then the flag
        __syncwarp();                               // (own triple: the next chunk reloads what the other lanes of this warp stored)
        if (trip >= nworkers && wl == 0) atomicExch(A.persist + 1 + trip, chunk + 1);
    }
    }   // work loop
}

}  // namespace

// Eligibility (M.rows_ok, set in get_dev_mech): a single serial chain (every joint at a distinct depth) of at most 6 joints,
// each of dimension <= 5.
int launch_rollout_rows(const DevMech& M, const DevCtrl& C, const SimArgs& a_in, cudaStream_t stream) {
    RowsOrder ord;
    for (int j = 0; j < 6; j++) ord.joint[j] = M.rows_order[j];
    const int G = 5 * M.nj, spw = 32 / G;
    int grid = (a_in.B + spw - 1) / spw;
    size_t smem = (size_t)(spw * M.nj * kGS + M.nj * kCS + spw * M.nj * 17) * sizeof(double);
    SimArgs a = a_in;
    // Two launches of the 255-register (spill-free) instantiation, which holds at most two one-warp CTAs per scheduler (slots2 = 8 per SM):
    //  * whole: the batch needs no more than slots2 warps anyway -- one chunk, no extra triples, i.e. the plain kernel's code without its
    //    168-register cap and its 300 bytes of spills.  2000 cartpole steps (tools/persist_probe4.py): 512 systems 17.7 against 19.6 ms, 2048:
    //    22.0 / 23.9, 3552: 22.4 / 24.3; pendulum 4096 x 1000: 9.3 / 10.0; triple cartpole 1184 x 1000: 20.4 / 22.5.
    //  * time-sliced: between 1x and 1.35x slots2 warps a plain launch runs at the three-warps-per-scheduler time T(3) with a good part of the
    //    third-warp slots empty; slots2 persistent workers cut the rollout into chunks of steps and share the extra triples on a static rotating
    //    schedule (work loop in the kernel): about T(2) x warps / slots2.  4096 cartpoles (1.15 slots2) 26.3 against 31.7 ms plain, 4400: 28.3 /
    //    31.9, 4800 (1.35): 30.7 / 31.7, 5328 (1.5): 34.0 / 32.6 -- hence the window.  Not with PID: its integrator state sits in global memory
    //    and an extra triple moves from SM to SM between chunks (L1 is not coherent).
    // Both are bit-identical to the plain launch (tests/test_gpu_edges.py).  CCB_K1R_WINDOW (percent), CCB_K1R_WHOLE (0 / 1), CCB_K1R_NO_PERSIST,
    // CCB_K1R_FORCE_PERSIST, CCB_K1R_CHUNKS: A/B knobs of the probes under tools/.
    static int nsm = 0;
    if (nsm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
    const int slots2 = 8 * nsm;
    int window = 135, whole = 1;
    if (const char* e = getenv("CCB_K1R_WINDOW")) window = atoi(e);
    if (const char* e = getenv("CCB_K1R_WHOLE")) whole = atoi(e);
    const bool eligible = a.persist && !a.knot_input && a.steps >= 256 && !getenv("CCB_K1R_NO_PERSIST");
    const bool sliced = grid > slots2 && C.kind != CCB_CTRL_PID && ((long long)grid * 100 <= (long long)slots2 * window || getenv("CCB_K1R_FORCE_PERSIST"));
    const bool fits = grid <= slots2 && (whole || getenv("CCB_K1R_FORCE_PERSIST"));
    const bool persist = eligible && (sliced || fits);
    if (persist) {
        if (sliced) {
            const int want = std::max(1, std::min(32, a.steps / 32));      // 32 chunks of >= 32 steps: every worker gets 32 + 32 extra / slots +- 1 items
            a.chunk_steps = (a.steps + want - 1) / want;
            if (const char* e = getenv("CCB_K1R_CHUNKS")) { const int nc = std::max(1, atoi(e)); a.chunk_steps = (a.steps + nc - 1) / nc; }
            a.nchunks = (a.steps + a.chunk_steps - 1) / a.chunk_steps;
            grid = slots2;
        } else {
            a.chunk_steps = a.steps; a.nchunks = 1;                        // workers = triples, no extra items
        }
        if (sliced) CCB_CUDA(cudaMemsetAsync(a.persist, 0, (size_t)(a.B + 2) * sizeof(int), stream));      // the flags of the extra triples (ccb_api.cu: B + 2 ints)
#ifndef K1R_NO_SMEM_CAP
        if (K1R_PERSIST_MAXREG < 224) smem = std::max(smem, (size_t)26 * 1024);     // with fewer registers shared memory caps the residency at 8 per SM
#endif
        if (getenv("CCB_K1R_DEBUG")) {
            int occ = -1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rollout_rows_kernel<2, true, true>, 32, smem);
            fprintf(stderr, "K1r persistent: grid %d, nchunks %d x %d steps, smem %zu, occupancy(NJ=2) %d CTAs/SM\n", grid, a.nchunks, a.chunk_steps, smem, occ);
        }
    } else {
        a.persist = nullptr; a.nchunks = 1; a.chunk_steps = a.steps;
    }
    // Measured (r2, 4096 x 2000 cartpole bench, tools/ab_k1r.sh): the LQR-only instantiation is 4 % shorter in SASS and NOT faster --
    // 32.0 ms against 31.4 ms for the one-for-all instantiation (31.6 ms with 160 registers): ptxas's register allocation at the 168 cap
    // moves more than the removed PID / table code saves.  The split stays available for A/B runs (-DK1R_SPLIT).
#ifdef K1R_SPLIT
    const bool gen = C.kind == CCB_CTRL_PID || C.kind == CCB_CTRL_TABLE;
#else
    const bool gen = true;
#endif
#define K1R_LAUNCH(NJ_)                                                                                   \
    case NJ_:                                                                                             \
        if (persist) rollout_rows_kernel<NJ_, true, true><<<grid, 32, smem, stream>>>(M, C, a, ord);      \
        else if (gen) rollout_rows_kernel<NJ_, true, false><<<grid, 32, smem, stream>>>(M, C, a, ord);    \
        else rollout_rows_kernel<NJ_, false, false><<<grid, 32, smem, stream>>>(M, C, a, ord);            \
        break;
    switch (M.nj) {
        K1R_LAUNCH(1) K1R_LAUNCH(2) K1R_LAUNCH(3) K1R_LAUNCH(4) K1R_LAUNCH(5)
        default:
            if (persist) rollout_rows_kernel<6, true, true><<<grid, 32, smem, stream>>>(M, C, a, ord);
            else if (gen) rollout_rows_kernel<6, true, false><<<grid, 32, smem, stream>>>(M, C, a, ord);
            else rollout_rows_kernel<6, false, false><<<grid, 32, smem, stream>>>(M, C, a, ord);
            break;
    }
#undef K1R_LAUNCH
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
