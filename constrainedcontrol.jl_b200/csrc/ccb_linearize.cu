// csrc/ccb_linearize.cu -- K2: batched `linearsystem` (reference call sites src/control/lqr.jl:63,
// src/control/lqr_tracking.jl:88; maths per SURVEY.md A.6).
//
// For every setpoint z_d = (x2,q2,v1,w1) with inputs Fτd the setpoint multipliers λ_d and velocities (v2,w2) come from
// one Newton solve (the rollout kernel in knot mode); this kernel then differentiates the implicit one-step map with
// u and λ as independent inputs.  With λ held fixed the step decouples per body, so A is block sparse:
//   δs_i = -D_i^-1 [ Σ_j ∂d_i/∂z_j δz_j + ∂d_i/∂u δu + ∂d_i/∂λ δλ ],  δx3 = δx2 + Δt δv2,  δφ3 = R(Ω)^T δφ2 + T_w δw2
// One CTA per setpoint; a thread owns one column of one 12x12 body block (or one column of Bu / Bλ, or one row of G).
// The second-order terms (constraint Hessians contracted with λ_d, input wrenches rotating with the bodies) are obtained
// with forward-mode dual numbers pushed through the analytic Jacobian rows.
#include <cstring>
#include "ccb_dev_dual.cuh"
#include "ccb_dev_math.cuh"
#include "ccb_internal.h"

namespace ccb {

namespace {

constexpr int LIN_THREADS = 128;
constexpr int BODY_SM = 43;   // x2 q2 (7) | x3 q3 (7) | Dinv (9) | Tw (9) | R(Omega)^T (9) | dt/m | 1 pad

template <class T> __device__ __forceinline__ void load_pose(const double* p, TV3<T>& x, TQ4<T>& q) {
    x = {T(p[0]), T(p[1]), T(p[2])};
    q = {T(p[3]), T(p[4]), T(p[5]), T(p[6])};
}

__global__ void __launch_bounds__(LIN_THREADS) linearize_kernel(const DevMech M, const double* __restrict__ zd, const double* __restrict__ ssol,
                                                                const double* __restrict__ lam, const double* __restrict__ Fd,
                                                                const int* __restrict__ eqc, int nu, double* __restrict__ A,
                                                                double* __restrict__ Bu, double* __restrict__ Bl, double* __restrict__ G) {
    extern __shared__ double sm[];
    const int sp = blockIdx.x, tid = threadIdx.x;
    const int nb = M.nb, ml = M.ml, n = 12 * nb;
    double* bod = sm;                      // [nb][BODY_SM]
    double* ue = sm + nb * BODY_SM;        // [ne] joint input at the setpoint
    const double* z = zd + (size_t)sp * 13 * nb;
    const double* s = ssol + (size_t)sp * 6 * nb;
    const double* l = lam + (size_t)sp * ml;
    const double dt = M.dt, c4 = 4.0 / (M.dt * M.dt);

    for (int e = tid; e < M.ne; e += LIN_THREADS) ue[e] = 0.0;
    __syncthreads();
    if (tid == 0) for (int k = 0; k < nu; k++) ue[eqc[k]] += Fd[(size_t)sp * nu + k];
    for (int b = tid; b < nb; b += LIN_THREADS) {
        double* o = bod + b * BODY_SM;
        const double* zb = z + 13 * b;
        V3 w2 = ld3(s + 6 * b + 3), v2 = ld3(s + 6 * b);
        for (int i = 0; i < 7; i++) o[i] = zb[i];
        const double sq2 = sqrt(c4 - dot(w2, w2)), isq = 1.0 / sq2;
        Q4 om = {0.5 * dt * sq2, 0.5 * dt * w2.x, 0.5 * dt * w2.y, 0.5 * dt * w2.z};
        st3(o + 7, ld3(zb) + dt * v2);
        st4(o + 10, qmul(ld4(zb + 3), om));
        double J[9];
        for (int i = 0; i < 9; i++) J[i] = __ldg(M.inertia + 9 * b + i);
        V3 Jw = mat3_mul(J, w2);
        const double w[3] = {w2.x, w2.y, w2.z}, jw[3] = {Jw.x, Jw.y, Jw.z};
        double Dr[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
                double wxJ = w[i1] * J[i2 * 3 + j] - w[i2] * J[i1 * 3 + j];
                double jwx = (j == i1) ? -jw[i2] : ((j == i2) ? jw[i1] : 0.0);
                Dr[i * 3 + j] = sq2 * J[i * 3 + j] + wxJ - jwx - jw[i] * w[j] * isq;
            }
        inv3(Dr, o + 14);
        const double c = 0.25 * dt * dt;
        double* T = o + 23;
        T[0] = c * (w[0] * w[0] * isq + sq2); T[1] = c * (w[0] * w[1] * isq + w[2]); T[2] = c * (w[0] * w[2] * isq - w[1]);
        T[3] = c * (w[1] * w[0] * isq - w[2]); T[4] = c * (w[1] * w[1] * isq + sq2); T[5] = c * (w[1] * w[2] * isq + w[0]);
        T[6] = c * (w[2] * w[0] * isq + w[1]); T[7] = c * (w[2] * w[1] * isq - w[0]); T[8] = c * (w[2] * w[2] * isq + sq2);
        for (int j = 0; j < 3; j++) {      // column j of R(Omega)^T
            V3 cj = qrotT(om, v3(j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0));
            o[32 + 0 * 3 + j] = cj.x; o[32 + 1 * 3 + j] = cj.y; o[32 + 2 * 3 + j] = cj.z;
        }
        o[41] = dt / __ldg(M.mass + b);
    }
    __syncthreads();

    // ---- (a) columns of A
    const int npairs = __ldg(M.cb_ptr + nb);
    for (int it = tid; it < npairs * 12; it += LIN_THREADS) {
        const int pidx = it / 12, c = it - pidx * 12;
        int i = 0;
        while (__ldg(M.cb_ptr + i + 1) <= pidx) i++;
        const int j = __ldg(M.cb_body + pidx);
        const double* bi = bod + i * BODY_SM;
        double E[6] = {0, 0, 0, 0, 0, 0};
        if (c >= 3 && c < 6) {
            if (i != j) continue;
            E[c - 3] = -__ldg(M.mass + i) / dt;                         // ∂d_T/∂v1
        } else if (c >= 9) {
            if (i != j) continue;                                         // ∂d_R/∂w1 = -∂[(sq1 I - [w1]x) J w1]/∂w1
            const int k = c - 9;
            double J[9];
            for (int q = 0; q < 9; q++) J[q] = __ldg(M.inertia + 9 * i + q);
            V3 w1 = ld3(z + 13 * i + 10);
            const double sq1 = sqrt(c4 - dot(w1, w1));
            V3 Jw = mat3_mul(J, w1);
            V3 ek = v3(k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0);
            V3 Jek = v3(J[k], J[3 + k], J[6 + k]);
            const double wk = k == 0 ? w1.x : (k == 1 ? w1.y : w1.z);
            V3 df = sq1 * Jek - (wk / sq1) * Jw - cross(ek, Jw) - cross(w1, Jek);
            E[3] = -df.x; E[4] = -df.y; E[5] = -df.z;
        } else {
            const int dir = c < 3 ? c : c - 3;   // 0..2 position, 3..5 orientation of body j
            for (int e = __ldg(M.bj_ptr + i); e < __ldg(M.bj_ptr + i + 1); e++) {
                const int jn = __ldg(M.bj_joint + e), side = __ldg(M.bj_side + e);
                const int a = __ldg(M.j_parent + jn), b = __ldg(M.j_child + jn);
                if (j != a && j != b) continue;
                TV3<Dual> xa = {Dual(0.0), Dual(0.0), Dual(0.0)}, xb;
                TQ4<Dual> qa = {Dual(1.0), Dual(0.0), Dual(0.0), Dual(0.0)}, qb;
                if (a >= 0) load_pose<Dual>(bod + a * BODY_SM, xa, qa);
                load_pose<Dual>(bod + b * BODY_SM, xb, qb);
                TV3<Dual>& xp = (j == a) ? xa : xb;
                TQ4<Dual>& qp = (j == a) ? qa : qb;
                if (dir < 3) {
                    (dir == 0 ? xp.x : (dir == 1 ? xp.y : xp.z)).d = 1.0;
                } else {   // q <- q (x) [1; eps e]  =>  dq = q (x) [0; e]
                    Q4 q0 = {qp.s.v, qp.x.v, qp.y.v, qp.z.v};
                    Q4 dq = qmul(q0, Q4{0.0, dir == 3 ? 1.0 : 0.0, dir == 4 ? 1.0 : 0.0, dir == 5 ? 1.0 : 0.0});
                    qp.s.d = dq.s; qp.x.d = dq.x; qp.y.d = dq.y; qp.z.d = dq.z;
                }
                const int r0 = __ldg(M.j_row0 + jn), dj = __ldg(M.j_dim + jn);
                TV3<Dual> p1 = lift<Dual>(M.j_p1 + 3 * jn), p2 = lift<Dual>(M.j_p2 + 3 * jn);
                TQ4<Dual> qoff = liftq<Dual>(M.j_qoff + 4 * jn);
                for (int r = r0; r < r0 + dj; r++) {   // -∂(G^T λ)/∂dir : constraint Hessian contracted with λ_d
                    TV3<Dual> Gax, Gap, Gbx, Gbp;
                    row_jac_t<Dual>(__ldg(M.row_kind + r), lift<Dual>(M.row_proj + 3 * r), p1, p2, qoff, xa, qa, xb, qb, Gax, Gap, Gbx, Gbp);
                    const TV3<Dual>& gx = side == 0 ? Gax : Gbx;
                    const TV3<Dual>& gp = side == 0 ? Gap : Gbp;
                    const double lr = l[r];
                    E[0] -= lr * gx.x.d; E[1] -= lr * gx.y.d; E[2] -= lr * gx.z.d;
                    E[3] -= lr * gp.x.d; E[4] -= lr * gp.y.d; E[5] -= lr * gp.z.d;
                }
                // input wrench rotating with the bodies ("linearforcemapping")
                for (int q = 0; q < M.ne; q++) {
                    if (__ldg(M.eqc_act_joint + q) != jn) continue;
                    const double u = ue[q];
                    if (u == 0.0) continue;
                    TV3<Dual> F, tau;
                    input_wrench_t<Dual>(__ldg(M.eqc_act_kind + q), Dual(u), lift<Dual>(M.j_axis + 3 * jn), p1, p2, qa, qb, side, F, tau);
                    E[0] -= F.x.d; E[1] -= F.y.d; E[2] -= F.z.d;
                    E[3] -= 2.0 * tau.x.d; E[4] -= 2.0 * tau.y.d; E[5] -= 2.0 * tau.z.d;
                }
            }
        }
        const double kdt = bi[41];
        V3 Nv = v3(-kdt * E[0], -kdt * E[1], -kdt * E[2]);
        V3 Nw = -1.0 * mat3_mul(bi + 14, v3(E[3], E[4], E[5]));
        V3 dx3 = dt * Nv, dp3 = mat3_mul(bi + 23, Nw);
        if (i == j && c < 3) { if (c == 0) dx3.x += 1.0; else if (c == 1) dx3.y += 1.0; else dx3.z += 1.0; }
        if (i == j && c >= 6 && c < 9) { const int k = c - 6; dp3.x += bi[32 + k]; dp3.y += bi[35 + k]; dp3.z += bi[38 + k]; }
        double* Ac = A + (size_t)sp * n * n + (size_t)(12 * i) * n + 12 * j + c;
        Ac[0 * n] = dx3.x; Ac[1 * n] = dx3.y; Ac[2 * n] = dx3.z;
        Ac[3 * n] = Nv.x; Ac[4 * n] = Nv.y; Ac[5 * n] = Nv.z;
        Ac[6 * n] = dp3.x; Ac[7 * n] = dp3.y; Ac[8 * n] = dp3.z;
        Ac[9 * n] = Nw.x; Ac[10 * n] = Nw.y; Ac[11 * n] = Nw.z;
    }

    // ---- (b) columns of Bu, (c) columns of Bλ: same pattern with E = -∂(F,2τ)/∂u and E = -G_r^T
    for (int it = tid; it < 2 * nu + 2 * ml; it += LIN_THREADS) {
        double E[6];
        int i, col;
        double* Bout;
        int ld;
        if (it < 2 * nu) {
            const int k = it >> 1, side = it & 1;
            const int q = eqc[k], jn = __ldg(M.eqc_act_joint + q);
            const int a = __ldg(M.j_parent + jn), b = __ldg(M.j_child + jn);
            i = side == 0 ? a : b;
            if (i < 0) continue;
            TV3<double> F, tau;
            TQ4<double> qa = {1.0, 0.0, 0.0, 0.0}, qb;
            TV3<double> tmp;
            if (a >= 0) load_pose<double>(bod + a * BODY_SM, tmp, qa);
            load_pose<double>(bod + b * BODY_SM, tmp, qb);
            input_wrench_t<double>(__ldg(M.eqc_act_kind + q), 1.0, lift<double>(M.j_axis + 3 * jn), lift<double>(M.j_p1 + 3 * jn),
                                   lift<double>(M.j_p2 + 3 * jn), qa, qb, side, F, tau);
            E[0] = -F.x; E[1] = -F.y; E[2] = -F.z; E[3] = -2.0 * tau.x; E[4] = -2.0 * tau.y; E[5] = -2.0 * tau.z;
            Bout = Bu + (size_t)sp * n * nu; ld = nu; col = k;
        } else {
            const int r = (it - 2 * nu) >> 1, side = (it - 2 * nu) & 1;
            const int jn = __ldg(M.row_joint + r);
            const int a = __ldg(M.j_parent + jn), b = __ldg(M.j_child + jn);
            i = side == 0 ? a : b;
            if (i < 0) continue;
            TV3<double> xa = {0.0, 0.0, 0.0}, xb;
            TQ4<double> qa = {1.0, 0.0, 0.0, 0.0}, qb;
            if (a >= 0) load_pose<double>(bod + a * BODY_SM, xa, qa);
            load_pose<double>(bod + b * BODY_SM, xb, qb);
            TV3<double> Gax, Gap, Gbx, Gbp;
            row_jac_t<double>(__ldg(M.row_kind + r), lift<double>(M.row_proj + 3 * r), lift<double>(M.j_p1 + 3 * jn), lift<double>(M.j_p2 + 3 * jn),
                              liftq<double>(M.j_qoff + 4 * jn), xa, qa, xb, qb, Gax, Gap, Gbx, Gbp);
            const TV3<double>& gx = side == 0 ? Gax : Gbx;
            const TV3<double>& gp = side == 0 ? Gap : Gbp;
            E[0] = -gx.x; E[1] = -gx.y; E[2] = -gx.z; E[3] = -gp.x; E[4] = -gp.y; E[5] = -gp.z;
            Bout = Bl + (size_t)sp * n * ml; ld = ml; col = r;
        }
        const double* bi = bod + i * BODY_SM;
        const double kdt = bi[41];
        V3 Nv = v3(-kdt * E[0], -kdt * E[1], -kdt * E[2]);
        V3 Nw = -1.0 * mat3_mul(bi + 14, v3(E[3], E[4], E[5]));
        V3 dx3 = dt * Nv, dp3 = mat3_mul(bi + 23, Nw);
        double* o = Bout + (size_t)(12 * i) * ld + col;
        o[0 * ld] = dx3.x; o[1 * ld] = dx3.y; o[2 * ld] = dx3.z;
        o[3 * ld] = Nv.x; o[4 * ld] = Nv.y; o[5 * ld] = Nv.z;
        o[6 * ld] = dp3.x; o[7 * ld] = dp3.y; o[8 * ld] = dp3.z;
        o[9 * ld] = Nw.x; o[10 * ld] = Nw.y; o[11 * ld] = Nw.z;
    }

    // ---- (d) rows of G = ∂g(x3,q3)/∂z+  (non-zero only in the δx, δφ columns)
    for (int r = tid; r < ml; r += LIN_THREADS) {
        const int jn = __ldg(M.row_joint + r);
        const int a = __ldg(M.j_parent + jn), b = __ldg(M.j_child + jn);
        TV3<double> xa = {0.0, 0.0, 0.0}, xb;
        TQ4<double> qa = {1.0, 0.0, 0.0, 0.0}, qb;
        if (a >= 0) load_pose<double>(bod + a * BODY_SM + 7, xa, qa);
        load_pose<double>(bod + b * BODY_SM + 7, xb, qb);
        TV3<double> Gax, Gap, Gbx, Gbp;
        row_jac_t<double>(__ldg(M.row_kind + r), lift<double>(M.row_proj + 3 * r), lift<double>(M.j_p1 + 3 * jn), lift<double>(M.j_p2 + 3 * jn),
                          liftq<double>(M.j_qoff + 4 * jn), xa, qa, xb, qb, Gax, Gap, Gbx, Gbp);
        double* g = G + (size_t)sp * ml * n + (size_t)r * n;
        g[12 * b + 0] = Gbx.x; g[12 * b + 1] = Gbx.y; g[12 * b + 2] = Gbx.z;
        g[12 * b + 6] = Gbp.x; g[12 * b + 7] = Gbp.y; g[12 * b + 8] = Gbp.z;
        if (a >= 0) {
            g[12 * a + 0] = Gax.x; g[12 * a + 1] = Gax.y; g[12 * a + 2] = Gax.z;
            g[12 * a + 6] = Gap.x; g[12 * a + 7] = Gap.y; g[12 * a + 8] = Gap.z;
        }
    }
}

}  // namespace

// Device-resident linearisation of S setpoints.  All pointers are device pointers; A, Bu, Bl, G are overwritten.
// d_lam [S][ml] is used as the warm start of the setpoint Newton solve and returns λ_d; d_s [S][nb][6] returns (v2,w2).
int linearize_device(HostMech* H, int device, int S, double* d_zd, const double* d_Fd, const int* d_eqc, int nu, const ccb_sim_opts* opts,
                     double* d_A, double* d_Bu, double* d_Bl, double* d_G, double* d_lam, double* d_s, int* d_status, cudaStream_t stream) {
    const DevMech* view;
    int rc = get_dev_mech(H, device, &view);
    if (rc != CCB_OK) return rc;
    const int nb = H->nb, ml = H->ml, n = 12 * nb;
    // setpoint Newton solve: z_d is the knot, inputs Fτd applied as a one-row open-loop table
    DevCtrl C;
    std::memset(&C, 0, sizeof(C));
    C.kind = nu > 0 ? CCB_CTRL_TABLE : CCB_CTRL_NONE; C.nu = nu; C.N = 1; C.eqc = d_eqc; C.U = d_Fd; C.U_bs = nu;
    SimArgs a;
    a.status_accum = 0; a.persist = nullptr; a.nchunks = 1; a.chunk_steps = 1;
    a.state = d_zd; a.lambda = d_lam; a.storage = nullptr; a.status = d_status; a.B = S; a.k0 = 1; a.steps = 1;
    a.eps = opts && opts->eps > 0 ? opts->eps : 1e-10;
    a.newton_iter = opts && opts->newton_iter > 0 ? opts->newton_iter : 100;
    a.line_iter = opts && opts->line_iter > 0 ? opts->line_iter : 10;
    a.knot_input = 1; a.s_out = d_s; a.spw = 0;
    rc = launch_rollout(*view, C, a, opts ? opts->lanes : 0, H->lanes, stream);
    if (rc != CCB_OK) return rc;
    CCB_CUDA(cudaMemsetAsync(d_A, 0, (size_t)S * n * n * sizeof(double), stream));
    if (nu > 0) CCB_CUDA(cudaMemsetAsync(d_Bu, 0, (size_t)S * n * nu * sizeof(double), stream));
    if (ml > 0) {
        CCB_CUDA(cudaMemsetAsync(d_Bl, 0, (size_t)S * n * ml * sizeof(double), stream));
        CCB_CUDA(cudaMemsetAsync(d_G, 0, (size_t)S * ml * n * sizeof(double), stream));
    }
    const size_t smem = ((size_t)nb * BODY_SM + (H->ne > 0 ? H->ne : 1)) * sizeof(double);
    linearize_kernel<<<S, LIN_THREADS, smem, stream>>>(*view, d_zd, d_s, d_lam, d_Fd, d_eqc, nu, d_A, d_Bu, d_Bl, d_G);
    g_launches++;
    CCB_CUDA(cudaGetLastError());
    return CCB_OK;
}

}  // namespace ccb
