// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU (Float64) restatement of the ConstrainedControl.jl hot path, used ONLY as the checker by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The shipped
// product path (constrainedcontrol.jl_b200/csrc, libccb.so) never links, loads or calls this file.
//
// PARITY UNPINNED at the numeric (1e-9) level: the reference's own tests are `@test true` smoke runs
// (test/examples/*.jl) and no Julia toolchain exists here, so there are no reference OUTPUTS to compare with.
// The restatement is pinned by known-answer / invariant tests (tests/test_oracle_*.py) and,
// on the one numeric fixture the reference owns -- the open-loop input sequence U of
// examples/trackingLQR_triple_cartpole.jl:1 -- by its known outcome: played through this restatement, U swings the
// triple pendulum on the cart up to within 0.6 degrees of upright, which a model 1 % off in mass, force scale or
// time step misses by 0.1 - 0.2 rad (tests/test_oracle_reference_fixture.py).
//
// What is restated, and from where:
//   dlqr (constrained finite-horizon Riccati)   src/control/lqr.jl:141-184
//   tracking dlqr (relinearise every k)         src/control/lqr_tracking.jl:73-122
//   control_lqr! finite / infinite              src/control/lqr.jl:89-114, :116-139
//   control_trackinglqr!                        src/control/lqr_tracking.jl:46-71
//   control_pid!, error_pid, stateError_pid     src/control/pid.jl:43-88
//   simulate!/newton!/lineSearch!/ldu_*/linearsystem : un-vendored ConstrainedDynamics.jl ^0.9.1
//       (Project.toml:7,12) + GraphBasedSystems.jl; algorithm per README.md:14,16 and SURVEY.md App. A.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>
#include "orc_dense.hpp"
#include "orc_mech.hpp"

using namespace orc;

namespace {

// ---------------------------------------------------------------------------------------------
// graph + symbolic analysis (tie-breaking restatement-defined: ascending ids, post-order)
// ---------------------------------------------------------------------------------------------
void dfs_visit(int v, int par, const std::vector<std::vector<int>>& adj, std::vector<int>& state, Mech& M) {
    state[v] = 1;
    M.parents[v] = par + 1;  // 1-based id, 0 = root
    for (int w : adj[v]) {
        if (w == par) continue;
        if (state[w] == 0) dfs_visit(w, v, adj, state, M);
        else if (state[w] == 1) { M.loop_edges.push_back(v + 1); M.loop_edges.push_back(w + 1); }  // back edge
    }
    state[v] = 2;
    M.dfs_list.push_back(v + 1);
}

void build_graph(Mech& M) {
    const int N = M.ne + M.nb;
    M.nnodes = N;
    M.eqc_id.resize(M.ne); M.body_id.resize(M.nb);
    for (int e = 0; e < M.ne; e++) M.eqc_id[e] = e + 1;           // eqcs first (SURVEY A.1 decision)
    for (int b = 0; b < M.nb; b++) M.body_id[b] = M.ne + b + 1;
    std::vector<std::vector<int>> adj(N);
    auto add_edge = [&](int a, int b) {
        if (std::find(adj[a].begin(), adj[a].end(), b) == adj[a].end()) { adj[a].push_back(b); adj[b].push_back(a); }
    };
    for (int e = 0; e < M.ne; e++) {
        if (M.eqc_parent[e] >= 0) add_edge(e, M.ne + M.eqc_parent[e]);
        for (int j = M.eqc_ptr[e]; j < M.eqc_ptr[e + 1]; j++) add_edge(e, M.ne + M.joints[j].child);
    }
    for (auto& a : adj) std::sort(a.begin(), a.end());
    M.parents.assign(N, 0);
    M.dfs_list.clear(); M.loop_edges.clear();
    std::vector<int> state(N, 0);
    // The origin is the root of the mechanism graph: the depth-first search starts at the equality constraints attached to it (ascending
    // ids), whatever their position in the list -- an origin joint with a single body must not end up as a LEAF of the elimination tree
    // (its diagonal block is zero until its body has been eliminated).  Components without an origin joint follow in ascending id order.
    for (int v = 0; v < M.ne; v++) if (state[v] == 0 && M.eqc_parent[v] < 0) dfs_visit(v, -1, adj, state, M);
    for (int v = 0; v < N; v++) if (state[v] == 0) dfs_visit(v, -1, adj, state, M);

    // block dims / offsets (dense KKT layout: bodies first, then lambda in eqc order)
    M.node_dim.resize(N); M.node_off.resize(N);
    for (int e = 0; e < M.ne; e++) { M.node_dim[e] = M.eqc_dim[e]; M.node_off[e] = 6 * M.nb + M.eqc_row0[e]; }
    for (int b = 0; b < M.nb; b++) { M.node_dim[M.ne + b] = 6; M.node_off[M.ne + b] = 6 * b; }
    // symbolic elimination in dfs order
    std::vector<int> pos(N);
    for (int i = 0; i < N; i++) pos[M.dfs_list[i] - 1] = i;
    std::vector<char> nz((size_t)N * N, 0);
    for (int v = 0; v < N; v++) { nz[(size_t)v * N + v] = 1; for (int w : adj[v]) nz[(size_t)v * N + w] = 1; }
    for (int i = 0; i < N; i++) {
        int c = M.dfs_list[i] - 1;
        std::vector<int> later;
        for (int p = 0; p < N; p++) if (pos[p] > i && nz[(size_t)p * N + c]) later.push_back(p);
        for (int p : later) for (int q : later) nz[(size_t)p * N + q] = 1;
    }
    M.blk.assign((size_t)N * N, -1);
    M.pool = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            if (nz[(size_t)i * N + j]) { M.blk[(size_t)i * N + j] = M.pool; M.pool += M.node_dim[i] * M.node_dim[j]; }
}

Mech* build_mech(int nb, int ne, int nj, double dt, double g, const double* mass, const double* inertia, const int* eqc_parent,
                 const int* eqc_ptr, const int* jchild, const int* jnT, const int* jnR, const double* jaxis, const double* jp1,
                 const double* jp2, const double* jqoff, double reg) {
    Mech* M = new Mech();
    M->nb = nb; M->ne = ne; M->nj = nj; M->dt = dt; M->g = g; M->reg = reg;
    M->m.assign(mass, mass + nb);
    M->J.assign(inertia, inertia + 9 * nb);
    M->eqc_parent.assign(eqc_parent, eqc_parent + ne);
    M->eqc_ptr.assign(eqc_ptr, eqc_ptr + ne + 1);
    M->joints.resize(nj);
    M->eqc_dim.assign(ne, 0); M->eqc_row0.assign(ne, 0);
    M->act_joint.assign(ne, -1); M->act_kind.assign(ne, 0);
    int row = 0;
    for (int e = 0; e < ne; e++) {
        M->eqc_row0[e] = row;
        std::vector<int> kids;
        for (int j = eqc_ptr[e]; j < eqc_ptr[e + 1]; j++) {
            Joint& J = M->joints[j];
            J.eqc = e; J.parent = eqc_parent[e]; J.child = jchild[j]; J.nT = jnT[j]; J.nR = jnR[j];
            J.dim = J.nT + J.nR; J.row0 = row; row += J.dim;
            double n = std::sqrt(jaxis[3 * j] * jaxis[3 * j] + jaxis[3 * j + 1] * jaxis[3 * j + 1] + jaxis[3 * j + 2] * jaxis[3 * j + 2]);
            for (int i = 0; i < 3; i++) { J.axis[i] = n > 0 ? jaxis[3 * j + i] / n : (i == 0 ? 1.0 : 0.0); J.p1[i] = jp1[3 * j + i]; J.p2[i] = jp2[3 * j + i]; }
            for (int i = 0; i < 4; i++) J.qoff[i] = jqoff[4 * j + i];
            set_projector(J.nT, J.axis, J.PT);
            set_projector(J.nR, J.axis, J.PR);
            if (std::find(kids.begin(), kids.end(), J.child) == kids.end()) kids.push_back(J.child);
        }
        M->eqc_dim[e] = row - M->eqc_row0[e];
        // pid.jl:31-37 : only 6*|unique(childids)| - length(eqc) == 1 is a 1-DoF (actuatable) constraint
        if (6 * (int)kids.size() - M->eqc_dim[e] == 1 && eqc_ptr[e + 1] - eqc_ptr[e] == 1) {
            int j = eqc_ptr[e];
            M->act_joint[e] = j;
            M->act_kind[e] = (M->joints[j].nT == 2) ? 1 : 2;
        }
    }
    M->ml = row;
    build_graph(*M);
    return M;
}

// ---------------------------------------------------------------------------------------------
// Newton workspace + KKT assembly (analytic blocks)
// ---------------------------------------------------------------------------------------------
template <class T> struct WorkT {
    int n;                         // 6 nb + ml
    std::vector<T> z, s, lam, s0, lam0, f, delta, u;
    std::vector<T> Md;             // dense KKT
    std::vector<int> piv;
    std::vector<T> pool, dinv, tmp;        // sparse blocks, diagonal inverses
    std::vector<int> dinv_off;
    std::vector<std::vector<int>> later;   // per dfs position: later neighbours (node index)
    explicit WorkT(const Mech& M) {
        n = 6 * M.nb + M.ml;
        z.resize(13 * M.nb); s.resize(6 * M.nb); lam.assign(M.ml, T(0.0)); s0 = s; lam0 = lam;
        f.resize(n); delta.resize(n); u.assign(M.ne, T(0.0));
        Md.resize((size_t)n * n); piv.resize(n);
        pool.resize(M.pool); tmp.resize(16 * 16);
        const int N = M.nnodes;
        dinv_off.resize(N);
        int o = 0;
        for (int v = 0; v < N; v++) { dinv_off[v] = o; o += M.node_dim[v] * M.node_dim[v]; }
        dinv.resize(o);
        std::vector<int> pos(N);
        for (int i = 0; i < N; i++) pos[M.dfs_list[i] - 1] = i;
        later.resize(N);
        for (int i = 0; i < N; i++) {
            int c = M.dfs_list[i] - 1;
            for (int p = 0; p < N; p++) if (pos[p] > i && M.blk[(size_t)p * N + c] >= 0) later[i].push_back(p);
        }
    }
};

typedef WorkT<double> Work;       // the reference's arithmetic; WorkT<Count> (oracle/flopcount: orc_count_flops) only COUNTS operations

template <class T> inline T norm2(const T* a, int n) { T s = T(0.0); for (int i = 0; i < n; i++) s += a[i] * a[i]; return orc::sqrt(s); }

// per-joint analytic blocks: G at (x2,q2) and Gv = d g(x3,q3)/d(v2,w2) for parent / child
template <class T>
void joint_blocks(const Mech& M, const Joint& J, const T* z, const T* s, T* Ga, T* Gb, T* Gva, T* Gvb) {
    Pose<T> a2 = J.parent >= 0 ? body_pose2(z, J.parent) : origin_pose<T>();
    Pose<T> b2 = body_pose2(z, J.child);
    joint_G(J, a2, b2, Ga, Gb);
    Pose<T> a3 = origin_pose<T>(), b3;
    T Ta[9], Tb[9];
    if (J.parent >= 0) {
        const T* sa = s + 6 * J.parent;
        next_pose<T>(a2, {sa[0], sa[1], sa[2]}, {sa[3], sa[4], sa[5]}, M.dt, a3);
        T_omega(sa + 3, M.dt, Ta);
    }
    const T* sb = s + 6 * J.child;
    next_pose<T>(b2, {sb[0], sb[1], sb[2]}, {sb[3], sb[4], sb[5]}, M.dt, b3);
    T_omega(sb + 3, M.dt, Tb);
    T G3a[36], G3b[36];
    joint_G(J, a3, b3, G3a, G3b);
    for (int r = 0; r < J.dim; r++) {
        for (int c = 0; c < 3; c++) {
            Gvb[r * 6 + c] = G3b[r * 6 + c] * T(M.dt);
            Gvb[r * 6 + 3 + c] = G3b[r * 6 + 3] * Tb[0 * 3 + c] + G3b[r * 6 + 4] * Tb[1 * 3 + c] + G3b[r * 6 + 5] * Tb[2 * 3 + c];
            if (J.parent >= 0) {
                Gva[r * 6 + c] = G3a[r * 6 + c] * T(M.dt);
                Gva[r * 6 + 3 + c] = G3a[r * 6 + 3] * Ta[0 * 3 + c] + G3a[r * 6 + 4] * Ta[1 * 3 + c] + G3a[r * 6 + 5] * Ta[2 * 3 + c];
            }
        }
    }
}

void assemble_dense(const Mech& M, Work& W) {
    const int n = W.n, nb = M.nb;
    std::fill(W.Md.begin(), W.Md.end(), 0.0);
    double D[36], Ga[36], Gb[36], Gva[36], Gvb[36];
    for (int b = 0; b < nb; b++) {
        body_D(M.m[b], &M.J[9 * b], M.dt, &W.s[6 * b + 3], D);
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) W.Md[(size_t)(6 * b + i) * n + 6 * b + j] = D[i * 6 + j];
    }
    for (const Joint& J : M.joints) {
        joint_blocks(M, J, W.z.data(), W.s.data(), Ga, Gb, Gva, Gvb);
        for (int r = 0; r < J.dim; r++) {
            int row = 6 * nb + J.row0 + r;
            for (int c = 0; c < 6; c++) {
                W.Md[(size_t)(6 * J.child + c) * n + row] = -Gb[r * 6 + c];
                W.Md[(size_t)row * n + 6 * J.child + c] = Gvb[r * 6 + c];
                if (J.parent >= 0) {
                    W.Md[(size_t)(6 * J.parent + c) * n + row] = -Ga[r * 6 + c];
                    W.Md[(size_t)row * n + 6 * J.parent + c] = Gva[r * 6 + c];
                }
            }
            W.Md[(size_t)row * n + row] = M.reg;
        }
    }
}

// mixed-node block LDU in dfs_list order: ldu_factorization! / ldu_backsubstitution!
template <class T>
bool solve_sparse(const Mech& M, WorkT<T>& W) {
    const int N = M.nnodes, nb = M.nb, ne = M.ne;
    std::fill(W.pool.begin(), W.pool.end(), T(0.0));
    auto B = [&](int i, int j) -> T* { return W.pool.data() + M.blk[(size_t)i * N + j]; };
    T D[36], Ga[36], Gb[36], Gva[36], Gvb[36];
    for (int b = 0; b < nb; b++) {
        body_D(M.m[b], &M.J[9 * b], M.dt, &W.s[6 * b + 3], D);
        std::copy(D, D + 36, B(ne + b, ne + b));
    }
    for (int e = 0; e < ne; e++) { T* d = B(e, e); int c = M.eqc_dim[e]; for (int i = 0; i < c; i++) d[i * c + i] = T(M.reg); }
    for (const Joint& J : M.joints) {
        joint_blocks(M, J, W.z.data(), W.s.data(), Ga, Gb, Gva, Gvb);
        const int e = J.eqc, c = M.eqc_dim[e], r0 = J.row0 - M.eqc_row0[e];
        for (int r = 0; r < J.dim; r++)
            for (int k = 0; k < 6; k++) {
                B(ne + J.child, e)[k * c + r0 + r] += -Gb[r * 6 + k];
                B(e, ne + J.child)[(r0 + r) * 6 + k] += Gvb[r * 6 + k];
                if (J.parent >= 0) {
                    B(ne + J.parent, e)[k * c + r0 + r] += -Ga[r * 6 + k];
                    B(e, ne + J.parent)[(r0 + r) * 6 + k] += Gva[r * 6 + k];
                }
            }
    }
    // rhs
    std::vector<T>& x = W.delta;
    x = W.f;
    // factorisation
    for (int i = 0; i < N; i++) {
        const int c = M.dfs_list[i] - 1, dc = M.node_dim[c];
        T* Dinv = W.dinv.data() + W.dinv_off[c];
        if (!small_inv(B(c, c), Dinv, dc)) { if (getenv("ORC_DEBUG")) fprintf(stderr, "solve_sparse: singular pivot block at dfs position %d (node id %d, dim %d)\n", i, c + 1, dc); return false; }
        const std::vector<int>& lat = W.later[i];
        for (int p : lat) {
            const int dp = M.node_dim[p];
            T* Mpc = B(p, c);
            T* L = W.tmp.data();
            mm(Mpc, Dinv, L, dp, dc, dc);                        // L_pc = M_pc D_c^-1
            for (int q : lat) {
                const int dq = M.node_dim[q];
                const T* Mcq = B(c, q);
                T* Mpq = B(p, q);
                for (int a = 0; a < dp; a++)
                    for (int k = 0; k < dc; k++) {
                        T l = L[a * dc + k];
                        if (l == T(0.0)) continue;
                        for (int b2 = 0; b2 < dq; b2++) Mpq[a * dq + b2] -= l * Mcq[k * dq + b2];
                    }
            }
            std::copy(L, L + dp * dc, Mpc);
        }
        for (int q : lat) {                                      // U_cq = D_c^-1 M_cq
            const int dq = M.node_dim[q];
            T* Mcq = B(c, q);
            T* U = W.tmp.data();
            mm(Dinv, Mcq, U, dc, dc, dq);
            std::copy(U, U + dc * dq, Mcq);
        }
    }
    // forward substitution
    for (int i = 0; i < N; i++) {
        const int c = M.dfs_list[i] - 1, dc = M.node_dim[c];
        const T* xc = x.data() + M.node_off[c];
        for (int p : W.later[i]) {
            const int dp = M.node_dim[p];
            const T* L = B(p, c);
            T* xp = x.data() + M.node_off[p];
            for (int a = 0; a < dp; a++) { T s = T(0.0); for (int k = 0; k < dc; k++) s += L[a * dc + k] * xc[k]; xp[a] -= s; }
        }
    }
    // diagonal + backward substitution
    for (int i = N - 1; i >= 0; i--) {
        const int c = M.dfs_list[i] - 1, dc = M.node_dim[c];
        T* xc = x.data() + M.node_off[c];
        const T* Dinv = W.dinv.data() + W.dinv_off[c];
        T t[16];
        for (int a = 0; a < dc; a++) { T s = T(0.0); for (int k = 0; k < dc; k++) s += Dinv[a * dc + k] * xc[k]; t[a] = s; }
        for (int q : W.later[i]) {
            const int dq = M.node_dim[q];
            const T* U = B(c, q);
            const T* xq = x.data() + M.node_off[q];
            for (int a = 0; a < dc; a++) { T s = T(0.0); for (int k = 0; k < dq; k++) s += U[a * dq + k] * xq[k]; t[a] -= s; }
        }
        for (int a = 0; a < dc; a++) xc[a] = t[a];
    }
    return true;
}

bool solve_dense(const Mech& M, Work& W) {
    assemble_dense(M, W);
    if (lu_factor(W.Md.data(), W.n, W.piv.data()) != 0) return false;
    W.delta = W.f;
    lu_solve(W.Md.data(), W.piv.data(), W.n, W.delta.data(), 1);
    return true;
}

template <class T> inline void eval_f(const Mech& M, WorkT<T>& W) { residual<T>(M, W.z.data(), W.s.data(), W.lam.data(), W.u.data(), W.f.data()); }
inline bool solve_any(const Mech& M, WorkT<double>& W, int solver);
template <class T> inline bool solve_any(const Mech& M, WorkT<T>& W, int) { return solve_sparse<T>(M, W); }

// newton! + lineSearch!: eps = 1e-10, newtonIter = 100, lineIter = 10 by default.
// Returns the number of Newton iterations used, negative if not converged / singular.
// Instrumentation for tests/test_gpu_fullsize.py (never changes the iterates): with a band tau > 1 set, a Newton iteration counts as MARGINAL
// when its termination decision `nd < eps && normf1 < eps` would come out differently with nd and normf1 scaled by 1/tau or by tau.  After the
// first (nearly exact) Newton step of a mildly nonlinear time step the next step is rounding noise amplified by the KKT inverse, ~1e-11..1e-9 --
// right at the default eps = 1e-10 -- so its termination decision, and with it the ITERATION COUNT, is decided by rounding and differs between
// any two Float64 implementations (already between this file's dense and block-sparse solvers: tests/test_oracle_control.py).
static double g_marginal_band = 0.0;
static thread_local long g_marginal_count = 0;

template <class T>
int newton(const Mech& M, WorkT<T>& W, double eps, int newton_iter, int line_iter, int solver) {
    const int ns = 6 * M.nb, ml = M.ml;
    eval_f(M, W);
    T normf0 = norm2(W.f.data(), W.n);
    for (int n = 1; n <= newton_iter; n++) {
        eval_f(M, W);   // setentries!: residual at the current iterate
        bool ok = solve_any(M, W, solver);
        if (!ok) return -1000 - n;
        W.s0 = W.s; W.lam0 = W.lam;
        T normf1 = normf0;
        int scale = 0;
        for (int it = 0; it < line_iter + 1; it++) {
            T a = T(1.0 / (double)(1 << scale));
            for (int i = 0; i < ns; i++) W.s[i] = W.s0[i] - a * W.delta[i];
            for (int i = 0; i < ml; i++) W.lam[i] = W.lam0[i] - a * W.delta[ns + i];
            eval_f(M, W);
            normf1 = norm2(W.f.data(), W.n);
            if (val(normf1) >= val(normf0)) scale++;
            else break;
        }
        T nd = T(0.0);
        for (int i = 0; i < ns; i++) { T d = W.s[i] - W.s0[i]; nd += d * d; }
        for (int i = 0; i < ml; i++) { T d = W.lam[i] - W.lam0[i]; nd += d * d; }
        nd = orc::sqrt(nd);
        if (g_marginal_band > 1.0) {
            const double t = g_marginal_band, ndv = val(nd), nfv = val(normf1);
            if ((ndv / t < eps && nfv / t < eps) != (ndv * t < eps && nfv * t < eps)) g_marginal_count++;
        }
        if (val(nd) < eps && val(normf1) < eps) return n;
        normf0 = normf1;
    }
    return -newton_iter;
}

inline bool solve_any(const Mech& M, WorkT<double>& W, int solver) { return solver == 0 ? solve_dense(M, W) : solve_sparse<double>(M, W); }

// ---------------------------------------------------------------------------------------------
// controllers
// ---------------------------------------------------------------------------------------------
struct orc_ctrl {
    int32_t kind;        // 0 none, 1 open-loop table, 2 LQR Inf, 3 LQR finite, 4 TrackingLQR, 5 PID
    int32_t nu;          // number of controlled eqcs
    int32_t N;           // horizon in steps (finite / tracking); K has N-1 rows
    int32_t pad;
    const int32_t* eqc;  // [nu] eqc indices (0-based)
    const double* K;     // [rows][nu][12 nb]   rows = 1 (Inf) or N-1
    int64_t K_bstride;   // per-system stride in doubles, 0 = shared
    const double* zd;    // LQR: [nb][13] ; tracking: [N][nb][13]   (x,q,v,w) desired
    int64_t zd_bstride;
    const double* Fd;    // LQR: [nu] ; tracking: [N][nu]
    int64_t Fd_bstride;
    const double* U;     // table: [steps][nu]
    int64_t U_bstride;
    const double* P; const double* I; const double* D;  // [nu]
    const double* goal;  // [nu]
    int64_t goal_bstride;
    double* pid_int;     // [B][nu] in/out  (pid.jl:76 -- never reset)
    double* pid_last;    // [B][nu] in/out
};

// Delta z (lqr.jl:90-104): per body [x2 - xd; v1 - vd; imag(qd \ q2); w1 - wd]
void delta_z(const Mech& M, const double* z, const double* zd, double* dz) {
    for (int b = 0; b < M.nb; b++) {
        const double* p = z + 13 * b;
        const double* d = zd + 13 * b;
        for (int i = 0; i < 3; i++) {
            dz[12 * b + i] = p[i] - d[i];
            dz[12 * b + 3 + i] = p[7 + i] - d[7 + i];
            dz[12 * b + 9 + i] = p[10 + i] - d[10 + i];
        }
        Q4<double> qe = qmul(qconj(Q4<double>{d[3], d[4], d[5], d[6]}), Q4<double>{p[3], p[4], p[5], p[6]});
        dz[12 * b + 6] = qe.x; dz[12 * b + 7] = qe.y; dz[12 * b + 8] = qe.z;
    }
}

// minimalCoordinates(mechanism, eqc)[1] for a 1-DoF eqc, on the continuous state (xc, qc)  
double minimal_coordinate(const Mech& M, int e, const double* state) {
    const Joint& J = M.joints[M.act_joint[e]];
    Pose<double> a = J.parent >= 0 ? body_pose2(state, J.parent) : origin_pose<double>();
    Pose<double> b = body_pose2(state, J.child);
    if (M.act_kind[e] == 1) {
        V3<double> w = b.x + qrot(b.q, lift3<double>(J.p2)) - a.x;
        V3<double> gT = qrotT(a.q, w) - lift3<double>(J.p1);
        return J.axis[0] * gT.x + J.axis[1] * gT.y + J.axis[2] * gT.z;
    }
    Q4<double> qe = qmul(qmul(qconj(a.q), b.q), qconj(lift4<double>(J.qoff)));
    double nv = std::sqrt(qe.x * qe.x + qe.y * qe.y + qe.z * qe.z);
    if (nv == 0.0) return 0.0;
    double ang = 2.0 * std::atan2(nv, qe.s);
    return ang * (J.axis[0] * qe.x + J.axis[1] * qe.y + J.axis[2] * qe.z) / nv;
}

// one control! call at (1-based) step k for system `sys`; fills W.u
void control(const Mech& M, Work& W, const orc_ctrl* C, int sys, int k, const double* state) {
    std::fill(W.u.begin(), W.u.end(), 0.0);
    if (!C || C->kind == 0) return;
    const int n = 12 * M.nb, nu = C->nu;
    if (C->kind == 1) {   // open loop (examples/trackingLQR_triple_cartpole.jl:45-47)
        const double* U = C->U + (size_t)sys * C->U_bstride + (size_t)(k - 1) * nu;
        for (int i = 0; i < nu; i++) W.u[C->eqc[i]] += U[i];
        return;
    }
    if (C->kind == 5) {   // pid.jl:69-88
        double* integ = C->pid_int + (size_t)sys * nu;
        double* last = C->pid_last + (size_t)sys * nu;
        const double* goal = C->goal + (size_t)sys * C->goal_bstride;
        for (int i = 0; i < nu; i++) {
            const int e = C->eqc[i];
            const Joint& J = M.joints[M.act_joint[e]];
            double err = goal[i] - minimal_coordinate(M, e, state);
            if (J.nT == 3 && J.nR == 2) {   // pid.jl:44-53 angle correction for Translational3 + Rotational2
                if (err > M_PI) err -= 2 * M_PI;
                else if (err < -M_PI) err += 2 * M_PI;
            }
            if (k == 1) last[i] = err;       // pid.jl:73
            integ[i] += err * M.dt;          // pid.jl:76
            double der = (err - last[i]) / M.dt;
            W.u[e] += C->P[i] * err + C->I[i] * integ[i] + C->D[i] * der;
            last[i] = err;
        }
        return;
    }
    std::vector<double> dz(n);
    const double* zd; const double* Fd; const double* K;
    if (C->kind == 2) {          // lqr.jl:116-139
        zd = C->zd + (size_t)sys * C->zd_bstride; Fd = C->Fd + (size_t)sys * C->Fd_bstride;
        K = C->K + (size_t)sys * C->K_bstride;
    } else if (C->kind == 3) {   // lqr.jl:89-114
        zd = C->zd + (size_t)sys * C->zd_bstride; Fd = C->Fd + (size_t)sys * C->Fd_bstride;
        if (!(k < C->N)) return;
        K = C->K + (size_t)sys * C->K_bstride + (size_t)(k - 1) * nu * n;
    } else {                     // lqr_tracking.jl:46-71
        if (!(k < C->N)) return;                  // lqr_tracking.jl:63
        zd = C->zd + (size_t)sys * C->zd_bstride + (size_t)(k - 1) * 13 * M.nb;
        Fd = C->Fd + (size_t)sys * C->Fd_bstride + (size_t)(k - 1) * nu;
        K = C->K + (size_t)sys * C->K_bstride + (size_t)(k - 1) * nu * n;
    }
    delta_z(M, W.z.data(), zd, dz.data());
    for (int i = 0; i < nu; i++) {
        double acc = 0;
        for (int j = 0; j < n; j++) acc += K[(size_t)i * n + j] * dz[j];
        W.u[C->eqc[i]] += Fd[i] - acc;
    }
}

// discretizestate!: knot (x2,q2) from the continuous state; z = (x2, q2, v1, w1)
void knot_from_state(const Mech& M, const double* st, double* z) {
    for (int b = 0; b < M.nb; b++) {
        const double* p = st + 13 * b;
        Pose<double> c = body_pose2(st, b), k2;
        next_pose<double>(c, {p[7], p[8], p[9]}, {p[10], p[11], p[12]}, M.dt, k2);
        double* o = z + 13 * b;
        o[0] = k2.x.x; o[1] = k2.x.y; o[2] = k2.x.z; o[3] = k2.q.s; o[4] = k2.q.x; o[5] = k2.q.y; o[6] = k2.q.z;
        for (int i = 7; i < 13; i++) o[i] = p[i];
    }
}

// one simulate! iteration for one system: control!, newton!, updatestate!
int sim_step(const Mech& M, Work& W, const orc_ctrl* C, int sys, int k, double* st, double* lam, double eps, int nit, int lit, int solver) {
    knot_from_state(M, st, W.z.data());
    control(M, W, C, sys, k, st);
    for (int b = 0; b < M.nb; b++) for (int i = 0; i < 6; i++) W.s[6 * b + i] = st[13 * b + 7 + i];
    for (int i = 0; i < M.ml; i++) W.lam[i] = lam[i];
    int it = newton(M, W, eps, nit, lit, solver);
    for (int b = 0; b < M.nb; b++) {
        for (int i = 0; i < 7; i++) st[13 * b + i] = W.z[13 * b + i];
        for (int i = 0; i < 6; i++) st[13 * b + 7 + i] = W.s[6 * b + i];
    }
    for (int i = 0; i < M.ml; i++) lam[i] = W.lam[i];
    return it;
}

// ---------------------------------------------------------------------------------------------
// linearsystem by forward-mode AD of the implicit one-step map
// ---------------------------------------------------------------------------------------------
int linearize(const Mech& M, Work& W, const double* zd, const double* Fd, const int* eqc_idx, int nu, double* lam_io, double eps, int nit,
              int lit, double* A, double* Bu, double* Bl, double* G) {
    const int nb = M.nb, n = 12 * nb, ml = M.ml;
    std::copy(zd, zd + 13 * nb, W.z.begin());
    std::fill(W.u.begin(), W.u.end(), 0.0);
    for (int i = 0; i < nu; i++) W.u[eqc_idx[i]] += Fd[i];
    for (int b = 0; b < nb; b++) for (int i = 0; i < 6; i++) W.s[6 * b + i] = zd[13 * b + 7 + i];
    for (int i = 0; i < ml; i++) W.lam[i] = lam_io ? lam_io[i] : 0.0;
    int it = newton(M, W, eps, nit, lit, 0);
    if (lam_io) for (int i = 0; i < ml; i++) lam_io[i] = W.lam[i];

    std::vector<Dual> zD(13 * nb), sD(6 * nb), lD(ml), uD(M.ne), fD(W.n);
    std::vector<double> Dinv(36 * nb), Dm(36);
    for (int b = 0; b < nb; b++) {
        body_D(M.m[b], &M.J[9 * b], M.dt, &W.s[6 * b + 3], Dm.data());
        small_inv(Dm.data(), &Dinv[36 * b], 6);
    }
    const int ndir = n + nu + ml;
    std::vector<double> col(n);
    for (int dir = 0; dir < ndir; dir++) {
        for (int i = 0; i < 13 * nb; i++) zD[i] = Dual(W.z[i]);
        for (int i = 0; i < 6 * nb; i++) sD[i] = Dual(W.s[i]);
        for (int i = 0; i < ml; i++) lD[i] = Dual(W.lam[i]);
        for (int i = 0; i < M.ne; i++) uD[i] = Dual(W.u[i]);
        if (dir < n) {
            const int b = dir / 12, c = dir % 12;
            Dual* p = &zD[13 * b];
            if (c < 3) p[c].d = 1.0;                 // dx
            else if (c < 6) p[7 + (c - 3)].d = 1.0;  // dv1
            else if (c < 9) {                         // dphi : dq = q (x) [0; e]
                Q4<double> q = {W.z[13 * b + 3], W.z[13 * b + 4], W.z[13 * b + 5], W.z[13 * b + 6]};
                Q4<double> e = {0, c == 6 ? 1.0 : 0.0, c == 7 ? 1.0 : 0.0, c == 8 ? 1.0 : 0.0};
                Q4<double> dq = qmul(q, e);
                p[3].d = dq.s; p[4].d = dq.x; p[5].d = dq.y; p[6].d = dq.z;
            } else p[10 + (c - 9)].d = 1.0;          // dw1
        } else if (dir < n + nu) uD[eqc_idx[dir - n]].d = 1.0;
        else lD[dir - n - nu].d = 1.0;
        residual<Dual>(M, zD.data(), sD.data(), lD.data(), uD.data(), fD.data());
        // ds_b = -D_b^-1 dd_b   (with lambda held as an independent input the step decouples per body)
        for (int b = 0; b < nb; b++)
            for (int i = 0; i < 6; i++) {
                double acc = 0;
                for (int k = 0; k < 6; k++) acc += Dinv[36 * b + i * 6 + k] * fD[6 * b + k].d;
                sD[6 * b + i].d = -acc;
            }
        for (int b = 0; b < nb; b++) {
            const Dual* p = &zD[13 * b];
            const Dual* sb = &sD[6 * b];
            Pose<Dual> p2 = {{p[0], p[1], p[2]}, {p[3], p[4], p[5], p[6]}}, p3;
            next_pose<Dual>(p2, {sb[0], sb[1], sb[2]}, {sb[3], sb[4], sb[5]}, M.dt, p3);
            Q4<double> q3 = {p3.q.s.v, p3.q.x.v, p3.q.y.v, p3.q.z.v}, dq3 = {p3.q.s.d, p3.q.x.d, p3.q.y.d, p3.q.z.d};
            Q4<double> ph = qmul(qconj(q3), dq3);   // dphi3 = V L(q3)^T dq3
            double* o = &col[12 * b];
            o[0] = p3.x.x.d; o[1] = p3.x.y.d; o[2] = p3.x.z.d;
            o[3] = sb[0].d; o[4] = sb[1].d; o[5] = sb[2].d;
            o[6] = ph.x; o[7] = ph.y; o[8] = ph.z;
            o[9] = sb[3].d; o[10] = sb[4].d; o[11] = sb[5].d;
        }
        for (int i = 0; i < n; i++) {
            if (dir < n) A[(size_t)i * n + dir] = col[i];
            else if (dir < n + nu) Bu[(size_t)i * nu + (dir - n)] = col[i];
            else Bl[(size_t)i * ml + (dir - n - nu)] = col[i];
        }
    }
    // G = d g(x3,q3) / d z+   (non-zero only in the dx, dphi columns)
    std::fill(G, G + (size_t)ml * n, 0.0);
    double Ga[36], Gb[36];
    for (const Joint& J : M.joints) {
        Pose<double> a3 = origin_pose<double>(), b3;
        if (J.parent >= 0) {
            const double* sa = &W.s[6 * J.parent];
            next_pose<double>(body_pose2(W.z.data(), J.parent), {sa[0], sa[1], sa[2]}, {sa[3], sa[4], sa[5]}, M.dt, a3);
        }
        const double* sb = &W.s[6 * J.child];
        next_pose<double>(body_pose2(W.z.data(), J.child), {sb[0], sb[1], sb[2]}, {sb[3], sb[4], sb[5]}, M.dt, b3);
        joint_G(J, a3, b3, Ga, Gb);
        for (int r = 0; r < J.dim; r++)
            for (int c = 0; c < 3; c++) {
                G[(size_t)(J.row0 + r) * n + 12 * J.child + c] = Gb[r * 6 + c];
                G[(size_t)(J.row0 + r) * n + 12 * J.child + 6 + c] = Gb[r * 6 + 3 + c];
                if (J.parent >= 0) {
                    G[(size_t)(J.row0 + r) * n + 12 * J.parent + c] = Ga[r * 6 + c];
                    G[(size_t)(J.row0 + r) * n + 12 * J.parent + 6 + c] = Ga[r * 6 + 3 + c];
                }
            }
    }
    return it;
}

// ---------------------------------------------------------------------------------------------
// constrained Riccati step shared by lqr.jl:150-177 and lqr_tracking.jl:87-115 (identical text there)
// Returns true when the early-break test `norm(Pk - Pkp1) < 1e-5` fires.  Writes Ku row block [mu][n].
// ---------------------------------------------------------------------------------------------
template <class T> struct RicWorkT {
    std::vector<T> GBl, GBlT, X, XG, D, DtP, M, b, Abar, t1, t2, P1, GBu, PA;
    std::vector<int> piv;
};
typedef RicWorkT<double> RicWork;

template <class T>
bool riccati_step(const T* A, const T* Bu, const T* Bl, const T* G, const T* Q, const T* R, int n, int mu,
                  int ml, std::vector<T>& P, T* Kuk_out, RicWorkT<T>& w, T* normdiff) {
    typedef std::vector<T> Mat;
    const int q = mu + ml;
    w.GBl.resize((size_t)ml * ml); w.GBlT.resize((size_t)ml * ml); w.X.resize((size_t)n * ml); w.XG.resize((size_t)n * n);
    w.D.resize((size_t)n * mu); w.DtP.resize((size_t)mu * n); w.M.resize((size_t)q * q); w.b.resize((size_t)q * n);
    w.Abar.resize((size_t)n * n); w.t1.resize((size_t)n * std::max(n, q)); w.t2.resize((size_t)n * std::max(n, q)); w.P1.resize((size_t)n * n);
    w.GBu.resize((size_t)ml * mu); w.piv.resize(std::max(q, ml) + 1);
    // D = Bu - Bl/(G*Bl)*G*Bu                                   lqr.jl:151   (left-assoc: ((Bl/(G Bl)) G) Bu)
    mm(G, Bl, w.GBl.data(), ml, n, ml);
    if (ml > 0) {
        // X = Bl / (G Bl)  <=>  (G Bl)^T X^T = Bl^T
        for (int i = 0; i < ml; i++) for (int j = 0; j < ml; j++) w.GBlT[(size_t)i * ml + j] = w.GBl[(size_t)j * ml + i];
        Mat XT((size_t)ml * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < ml; j++) XT[(size_t)j * n + i] = Bl[(size_t)i * ml + j];
        lu_factor(w.GBlT.data(), ml, w.piv.data());
        lu_solve(w.GBlT.data(), w.piv.data(), ml, XT.data(), n);
        for (int i = 0; i < n; i++) for (int j = 0; j < ml; j++) w.X[(size_t)i * ml + j] = XT[(size_t)j * n + i];
    }
    mm(w.X.data(), G, w.XG.data(), n, ml, n);
    mm(w.XG.data(), Bu, w.D.data(), n, n, mu);
    for (int i = 0; i < n * mu; i++) w.D[i] = Bu[i] - w.D[i];
    // M = [R + D'P Bu, D'P Bl; G Bu, G Bl]                      lqr.jl:152-157
    mtm(w.D.data(), P.data(), w.DtP.data(), mu, n, n);
    Mat M11((size_t)mu * mu), M12((size_t)mu * ml);
    mm(w.DtP.data(), Bu, M11.data(), mu, n, mu);
    mm(w.DtP.data(), Bl, M12.data(), mu, n, ml);
    mm(G, Bu, w.GBu.data(), ml, n, mu);
    for (int i = 0; i < mu; i++) {
        for (int j = 0; j < mu; j++) w.M[(size_t)i * q + j] = R[(size_t)i * mu + j] + M11[(size_t)i * mu + j];
        for (int j = 0; j < ml; j++) w.M[(size_t)i * q + mu + j] = M12[(size_t)i * ml + j];
    }
    for (int i = 0; i < ml; i++) {
        for (int j = 0; j < mu; j++) w.M[(size_t)(mu + i) * q + j] = w.GBu[(size_t)i * mu + j];
        for (int j = 0; j < ml; j++) w.M[(size_t)(mu + i) * q + mu + j] = w.GBl[(size_t)i * ml + j];
    }
    // b = [D'P; G] A                                            lqr.jl:158
    Mat top((size_t)q * n);
    for (int i = 0; i < mu; i++) for (int j = 0; j < n; j++) top[(size_t)i * n + j] = w.DtP[(size_t)i * n + j];
    for (int i = 0; i < ml; i++) for (int j = 0; j < n; j++) top[(size_t)(mu + i) * n + j] = G[(size_t)i * n + j];
    mm(top.data(), A, w.b.data(), q, n, n);
    // Kk = M \ b                                                lqr.jl:160
    lu_factor(w.M.data(), q, w.piv.data());
    lu_solve(w.M.data(), w.piv.data(), q, w.b.data(), n);
    const T* Kk = w.b.data();
    for (int i = 0; i < mu * n; i++) Kuk_out[i] = Kk[i];        // lqr.jl:162-164
    // Abar = A - Bu Kuk - Bl Klk                                lqr.jl:169
    mm(Bu, Kk, w.t1.data(), n, mu, n);
    mm(Bl, Kk + (size_t)mu * n, w.t2.data(), n, ml, n);
    for (int i = 0; i < n * n; i++) w.Abar[i] = A[i] - w.t1[i] - w.t2[i];
    // Pkp1 = Q + Kuk' R Kuk + Abar' Pk Abar                     lqr.jl:170
    Mat RK((size_t)mu * n), KRK((size_t)n * n);
    mm(R, Kk, RK.data(), mu, mu, n);
    mtm(Kk, RK.data(), KRK.data(), n, mu, n);
    mtm(w.Abar.data(), P.data(), w.t1.data(), n, n, n);         // Abar' P
    mm(w.t1.data(), w.Abar.data(), w.t2.data(), n, n, n);
    T nd = 0;
    for (int i = 0; i < n * n; i++) {
        w.P1[i] = Q[i] + KRK[i] + w.t2[i];
        T d = P[i] - w.P1[i];
        nd += d * d;
    }
    nd = t_sqrt(nd);
    *normdiff = nd;
    if (nd < T(1e-5)) return true;                                  // lqr.jl:172 (Pk is NOT updated on break)
    P = w.P1;
    return false;
}

}  // namespace

// The same loop (lqr.jl:141-184) with the recursion carried in a wider scalar: prec 0 = Float64 (identical to orc_dlqr), 1 = x87 extended
// (long double), 2 = IEEE binary128 (__float128).  Inputs and outputs are Float64 (the linearisation is the reference's Float64 data; gains are
// rounded on output).  normdiff_trace [N-1] (optional) receives ||Pk - Pkp1||_F of every executed iteration -- the quantity the early break
// of lqr.jl:172 compares with 1e-5.  Pmax_out (optional): max |P_ij| at exit.  Purpose: separating "two Float64 implementations differ" from
// "the recursion amplifies rounding" (tests/test_oracle_quad.py, tools/quad_gains.py); never a parity reference by itself.
template <class T>
int dlqr_mp(const double* A_, const double* Bu_, const double* Bl_, const double* G_, const double* Q_, const double* R_, int n, int mu, int ml,
                   int N, double* Ku, int* iters_out, double* normdiff_trace, double* Pmax_out) {
    auto up = [](const double* p, size_t cnt) { std::vector<T> v(cnt); for (size_t i = 0; i < cnt; i++) v[i] = (T)p[i]; return v; };
    std::vector<T> A = up(A_, (size_t)n * n), Bu = up(Bu_, (size_t)n * mu), Bl = up(Bl_, (size_t)n * ml), G = up(G_, (size_t)ml * n),
                   Q = up(Q_, (size_t)n * n), R = up(R_, (size_t)mu * mu), P = Q, Kk((size_t)mu * n);
    RicWorkT<T> w;
    int k = 0, iters = 0;
    T nd;
    for (k = N - 1; k >= 1; k--) {
        bool brk = riccati_step<T>(A.data(), Bu.data(), Bl.data(), G.data(), Q.data(), R.data(), n, mu, ml, P, Kk.data(), w, &nd);
        for (int i = 0; i < mu * n; i++) Ku[(size_t)(k - 1) * mu * n + i] = (double)Kk[i];
        if (normdiff_trace) normdiff_trace[iters] = (double)nd;
        iters++;
        if (brk) break;
    }
    if (k < 1) k = 1;
    for (int k2 = k - 1; k2 >= 1; k2--) std::memcpy(Ku + (size_t)(k2 - 1) * mu * n, Ku + (size_t)k2 * mu * n, sizeof(double) * mu * n);
    if (iters_out) *iters_out = iters;
    if (Pmax_out) { T m = 0; for (auto& v : P) if (t_abs(v) > m) m = t_abs(v); *Pmax_out = (double)m; }
    return k;
}

// tracking dlqr (lqr_tracking.jl:73-122) with the recursion in a wider scalar; every linearisation is the Float64 one (the data of the problem)
template <class T>
int tracking_dlqr_mp(Mech& M, const double* traj, const double* Ftau, const int* eqc_idx, int nu, const double* Q_, const double* R_, int N,
                     double eps, int newton_iter, int line_iter, double* Ku, int* iters_out) {
    const int n = 12 * M.nb, ml = M.ml;
    Work W(M);
    Mat A((size_t)n * n), Bu((size_t)n * nu), Bl((size_t)n * ml), G((size_t)ml * n);
    auto up = [](const double* p, size_t cnt) { std::vector<T> v(cnt); for (size_t i = 0; i < cnt; i++) v[i] = (T)p[i]; return v; };
    std::vector<T> Q = up(Q_, (size_t)n * n), R = up(R_, (size_t)nu * nu), P = Q, Kk((size_t)nu * n);
    std::vector<double> lam(ml, 0.0);
    RicWorkT<T> w;
    int k = 0, iters = 0;
    T nd;
    for (k = N - 1; k >= 1; k--) {
        iters++;
        linearize(M, W, traj + (size_t)(k - 1) * 13 * M.nb, Ftau + (size_t)(k - 1) * nu, eqc_idx, nu, lam.data(), eps, newton_iter, line_iter,
                  A.data(), Bu.data(), Bl.data(), G.data());   // lqr_tracking.jl:88
        std::vector<T> At = up(A.data(), A.size()), But = up(Bu.data(), Bu.size()), Blt = up(Bl.data(), Bl.size()), Gt = up(G.data(), G.size());
        bool brk = riccati_step<T>(At.data(), But.data(), Blt.data(), Gt.data(), Q.data(), R.data(), n, nu, ml, P, Kk.data(), w, &nd);
        for (int i = 0; i < nu * n; i++) Ku[(size_t)(k - 1) * nu * n + i] = (double)Kk[i];
        if (brk) break;
    }
    if (k < 1) k = 1;
    for (int k2 = k - 1; k2 >= 1; k2--) std::memcpy(Ku + (size_t)(k2 - 1) * nu * n, Ku + (size_t)k2 * nu * n, sizeof(double) * nu * n);
    if (iters_out) *iters_out = iters;
    return k;
}

// =============================================================================================
// C interface (ctypes)
// =============================================================================================
extern "C" {

void* orc_mech_create(int nb, int ne, int nj, double dt, double g, const double* mass, const double* inertia, const int* eqc_parent,
                      const int* eqc_ptr, const int* jchild, const int* jnT, const int* jnR, const double* jaxis, const double* jp1,
                      const double* jp2, const double* jqoff, double reg) {
    return build_mech(nb, ne, nj, dt, g, mass, inertia, eqc_parent, eqc_ptr, jchild, jnT, jnR, jaxis, jp1, jp2, jqoff, reg);
}
void orc_mech_destroy(void* h) { delete (Mech*)h; }

// out: nb, ne, nj, ml, nnodes, nloop_edges
void orc_mech_dims(void* h, int* out) {
    Mech* M = (Mech*)h;
    out[0] = M->nb; out[1] = M->ne; out[2] = M->nj; out[3] = M->ml; out[4] = M->nnodes; out[5] = (int)M->loop_edges.size() / 2;
}
void orc_mech_graph(void* h, int* eqc_id, int* body_id, int* dfs_list, int* parents, int* loop_edges, int* eqc_dim, int* eqc_row0,
                    int* act_kind) {
    Mech* M = (Mech*)h;
    std::copy(M->eqc_id.begin(), M->eqc_id.end(), eqc_id);
    std::copy(M->body_id.begin(), M->body_id.end(), body_id);
    std::copy(M->dfs_list.begin(), M->dfs_list.end(), dfs_list);
    std::copy(M->parents.begin(), M->parents.end(), parents);
    std::copy(M->loop_edges.begin(), M->loop_edges.end(), loop_edges);
    std::copy(M->eqc_dim.begin(), M->eqc_dim.end(), eqc_dim);
    std::copy(M->eqc_row0.begin(), M->eqc_row0.end(), eqc_row0);
    std::copy(M->act_kind.begin(), M->act_kind.end(), act_kind);
}

// residual f(z, s, lam, u) with z = (x2,q2,v1,w1) per body
void orc_residual(void* h, const double* z, const double* s, const double* lam, const double* u, double* f) {
    residual<double>(*(Mech*)h, z, s, lam, u, f);
}
// analytic dense KKT matrix (row-major (6nb+ml)^2) and residual
void orc_kkt(void* h, const double* z, const double* s, const double* lam, const double* u, double* Mout, double* f) {
    Mech& M = *(Mech*)h;
    Work W(M);
    std::copy(z, z + 13 * M.nb, W.z.begin()); std::copy(s, s + 6 * M.nb, W.s.begin());
    std::copy(lam, lam + M.ml, W.lam.begin()); std::copy(u, u + M.ne, W.u.begin());
    eval_f(M, W);
    assemble_dense(M, W);
    std::copy(W.Md.begin(), W.Md.end(), Mout);
    std::copy(W.f.begin(), W.f.end(), f);
}
// the same matrix by forward-mode AD of the residual (no regularisation term)
void orc_kkt_ad(void* h, const double* z, const double* s, const double* lam, const double* u, double* Mout) {
    Mech& M = *(Mech*)h;
    const int n = 6 * M.nb + M.ml;
    std::vector<Dual> zD(13 * M.nb), sD(6 * M.nb), lD(M.ml), uD(M.ne), fD(n);
    for (int dir = 0; dir < n; dir++) {
        for (int i = 0; i < 13 * M.nb; i++) zD[i] = Dual(z[i]);
        for (int i = 0; i < 6 * M.nb; i++) sD[i] = Dual(s[i]);
        for (int i = 0; i < M.ml; i++) lD[i] = Dual(lam[i]);
        for (int i = 0; i < M.ne; i++) uD[i] = Dual(u[i]);
        if (dir < 6 * M.nb) sD[dir].d = 1.0; else lD[dir - 6 * M.nb].d = 1.0;
        residual<Dual>(M, zD.data(), sD.data(), lD.data(), uD.data(), fD.data());
        for (int i = 0; i < n; i++) Mout[(size_t)i * n + dir] = fD[i].d;
    }
}

// solve the KKT system at (z,s,lam,u) with both solvers -> delta_dense, delta_sparse (each 6nb+ml)
int orc_solve_both(void* h, const double* z, const double* s, const double* lam, const double* u, double* dd, double* ds) {
    Mech& M = *(Mech*)h;
    Work W(M);
    std::copy(z, z + 13 * M.nb, W.z.begin()); std::copy(s, s + 6 * M.nb, W.s.begin());
    std::copy(lam, lam + M.ml, W.lam.begin()); std::copy(u, u + M.ne, W.u.begin());
    eval_f(M, W);
    bool a = solve_dense(M, W);
    std::copy(W.delta.begin(), W.delta.end(), dd);
    bool b = solve_sparse(M, W);
    std::copy(W.delta.begin(), W.delta.end(), ds);
    return (a ? 0 : 1) + (b ? 0 : 2);
}

static int* g_marginal_out = nullptr;
// marginal_out [B] receives, per system, the number of Newton iterations of the NEXT orc_simulate call whose termination decision is
// marginal within the band (see newton); band <= 1 or NULL switches the instrumentation off again
void orc_set_marginal(double band, int* marginal_out) { g_marginal_band = band; g_marginal_out = marginal_out; }

// simulate!: states [B][nb][13] (xc,qc,vc,wc) in/out, lambdas [B][ml] in/out,
// storage [steps][B][nb][13] or NULL, iters_out [B] = total Newton iterations (negated if any step failed)
void orc_simulate(void* h, int B, double* states, double* lambdas, int k0, int steps, const orc_ctrl* C, double eps, int newton_iter,
                  int line_iter, int solver, double* storage, int* iters_out, int nthreads) {
    Mech& M = *(Mech*)h;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
    auto run = [&](int lo, int hi) {
        Work W(M);
        const int sn = 13 * M.nb;
        for (int sys = lo; sys < hi; sys++) {
            double* st = states + (size_t)sys * sn;
            double* lam = lambdas + (size_t)sys * M.ml;
            int tot = 0; bool bad = false;
            g_marginal_count = 0;
            for (int k = 0; k < steps; k++) {
                if (storage) std::memcpy(storage + ((size_t)k * B + sys) * sn, st, sizeof(double) * sn);
                int it = sim_step(M, W, C, sys, k0 + k, st, lam, eps, newton_iter, line_iter, solver);
                if (it < 0) { bad = true; tot += -it > 1000 ? 0 : -it; } else tot += it;
            }
            if (iters_out) iters_out[sys] = bad ? -tot : tot;
            if (g_marginal_out) g_marginal_out[sys] = (int)g_marginal_count;
        }
    };
    if (nthreads == 1) { run(0, B); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        int lo = (int)((long long)B * t / nthreads), hi = (int)((long long)B * (t + 1) / nthreads);
        th.emplace_back(run, lo, hi);
    }
    for (auto& t : th) t.join();
}

// u that control! would apply at step k (for unit tests of the control laws); state [nb][13]
void orc_control(void* h, const orc_ctrl* C, int sys, int k, const double* state, double* u_out) {
    Mech& M = *(Mech*)h;
    Work W(M);
    knot_from_state(M, state, W.z.data());
    control(M, W, C, sys, k, state);
    std::copy(W.u.begin(), W.u.end(), u_out);
}

// linearsystem(mechanism, xd, vd, qd, wd, Ftaud, bodyids, eqcids) -> A, Bu, Bl, G    (lqr.jl:63)
// zd [nb][13] = (xd,qd,vd,wd); lam_io [ml] in: warm start, out: setpoint multipliers (may be NULL)
int orc_linearize(void* h, const double* zd, const double* Fd, const int* eqc_idx, int nu, double* lam_io, double eps, int newton_iter,
                  int line_iter, double* A, double* Bu, double* Bl, double* G) {
    Mech& M = *(Mech*)h;
    Work W(M);
    return linearize(M, W, zd, Fd, eqc_idx, nu, lam_io, eps, newton_iter, line_iter, A, Bu, Bl, G);
}

// dlqr(A,Bu,Bl,G,Q,R,N) lqr.jl:141-184.  Ku [N-1][mu][n]; returns the loop variable k at exit
// (k after a completed loop is 1; C.4 of SURVEY).  iters_out = executed iterations.
int orc_dlqr(const double* A, const double* Bu, const double* Bl, const double* G, const double* Q, const double* R, int n, int mu, int ml,
             int N, double* Ku, int* iters_out, double* P_out) {
    Mat P(Q, Q + (size_t)n * n);                                  // lqr.jl:147  Pk = Q
    RicWork w;
    int k = 0, iters = 0;
    double nd;
    for (k = N - 1; k >= 1; k--) {                               // lqr.jl:150
        iters++;
        bool brk = riccati_step(A, Bu, Bl, G, Q, R, n, mu, ml, P, Ku + (size_t)(k - 1) * mu * n, w, &nd);
        if (brk) break;
    }
    if (k < 1) k = 1;                                            // completed loop leaves k == 1
    for (int k2 = k - 1; k2 >= 1; k2--)                          // lqr.jl:179-181 back-fill
        std::memcpy(Ku + (size_t)(k2 - 1) * mu * n, Ku + (size_t)k2 * mu * n, sizeof(double) * mu * n);
    if (iters_out) *iters_out = iters;
    if (P_out) std::copy(P.begin(), P.end(), P_out);
    return k;
}

int orc_dlqr_mp(int prec, const double* A, const double* Bu, const double* Bl, const double* G, const double* Q, const double* R, int n, int mu,
                int ml, int N, double* Ku, int* iters_out, double* normdiff_trace, double* Pmax_out) {
    if (prec == 0) return dlqr_mp<double>(A, Bu, Bl, G, Q, R, n, mu, ml, N, Ku, iters_out, normdiff_trace, Pmax_out);
    if (prec == 1) return dlqr_mp<long double>(A, Bu, Bl, G, Q, R, n, mu, ml, N, Ku, iters_out, normdiff_trace, Pmax_out);
#ifdef ORC_WITH_QUAD
    if (prec == 2) return dlqr_mp<__float128>(A, Bu, Bl, G, Q, R, n, mu, ml, N, Ku, iters_out, normdiff_trace, Pmax_out);
#endif
    return -1;
}

// dlqr(mechanism, xd, vd, qd, wd, Ftaud, eqcids, Q, R, N) lqr_tracking.jl:73-122
// traj [N][nb][13], Ftau [N][nu]; Ku [N-1][nu][n]
int orc_tracking_dlqr(void* h, const double* traj, const double* Ftau, const int* eqc_idx, int nu, const double* Q, const double* R, int N,
                      double eps, int newton_iter, int line_iter, double* Ku, int* iters_out) {
    Mech& M = *(Mech*)h;
    const int n = 12 * M.nb, ml = M.ml;
    Work W(M);
    Mat A((size_t)n * n), Bu((size_t)n * nu), Bl((size_t)n * ml), G((size_t)ml * n), P(Q, Q + (size_t)n * n);
    std::vector<double> lam(ml, 0.0);
    RicWork w;
    int k = 0, iters = 0;
    double nd;
    for (k = N - 1; k >= 1; k--) {
        iters++;
        linearize(M, W, traj + (size_t)(k - 1) * 13 * M.nb, Ftau + (size_t)(k - 1) * nu, eqc_idx, nu, lam.data(), eps, newton_iter, line_iter,
                  A.data(), Bu.data(), Bl.data(), G.data());   // lqr_tracking.jl:88
        bool brk = riccati_step(A.data(), Bu.data(), Bl.data(), G.data(), Q, R, n, nu, ml, P, Ku + (size_t)(k - 1) * nu * n, w, &nd);
        if (brk) break;
    }
    if (k < 1) k = 1;
    for (int k2 = k - 1; k2 >= 1; k2--) std::memcpy(Ku + (size_t)(k2 - 1) * nu * n, Ku + (size_t)k2 * nu * n, sizeof(double) * nu * n);
    if (iters_out) *iters_out = iters;
    return k;
}

int orc_tracking_dlqr_mp(int prec, void* h, const double* traj, const double* Ftau, const int* eqc_idx, int nu, const double* Q, const double* R,
                         int N, double eps, int newton_iter, int line_iter, double* Ku, int* iters_out) {
    Mech& M = *(Mech*)h;
    if (prec == 0) return tracking_dlqr_mp<double>(M, traj, Ftau, eqc_idx, nu, Q, R, N, eps, newton_iter, line_iter, Ku, iters_out);
    if (prec == 1) return tracking_dlqr_mp<long double>(M, traj, Ftau, eqc_idx, nu, Q, R, N, eps, newton_iter, line_iter, Ku, iters_out);
#ifdef ORC_WITH_QUAD
    if (prec == 2) return tracking_dlqr_mp<__float128>(M, traj, Ftau, eqc_idx, nu, Q, R, N, eps, newton_iter, line_iter, Ku, iters_out);
#endif
    return -1;
}

// EXACT floating-point operation count of ONE newton! solve of the literal algorithm ("exact count from the oracle instrumented
// with a counting scalar type"): the templated Newton solve (residual, analytic KKT blocks, mixed-node block LDU in dfs_list order,
// back-substitution, line search) instantiated with orc::Count at the knot `state` [nb][13] with warm start `lam` [ml] and inputs u [ne].
// out[0..4] = adds, muls, divs, sqrts, compares over the whole solve; out[5] = Newton iterations it took.  Dividing by the iterations gives
// flops per Newton iteration (each iteration = 1 factorisation + solve + >= 1 residual evaluation of the line search; the one initial residual
// evaluation is included in the total).
int orc_count_flops(void* h, const double* state, const double* lam, const double* u, double eps, int newton_iter, int line_iter, long long* out) {
    Mech& M = *(Mech*)h;
    WorkT<Count> W(M);
    std::vector<double> z(13 * M.nb);
    knot_from_state(M, state, z.data());
    for (int i = 0; i < 13 * M.nb; i++) W.z[i] = Count(z[i]);
    for (int b = 0; b < M.nb; b++) for (int i = 0; i < 6; i++) W.s[6 * b + i] = Count(state[13 * b + 7 + i]);
    for (int i = 0; i < M.ml; i++) W.lam[i] = Count(lam ? lam[i] : 0.0);
    for (int e = 0; e < M.ne; e++) W.u[e] = Count(u ? u[e] : 0.0);
    flop_tally() = FlopTally();
    int it = newton<Count>(M, W, eps, newton_iter, line_iter, 1);
    const FlopTally& t = flop_tally();
    out[0] = t.add; out[1] = t.mul; out[2] = t.div; out[3] = t.sqrt_; out[4] = t.cmp; out[5] = it;
    return it;
}

// The same count over a closed-loop ROLLOUT of one system: simulate! with the control law evaluated in Float64 (its 2 nu n flops are not
// counted) and every newton! solve run in the counting scalar, which performs the identical operations in the identical order, so the
// trajectory is bit-identical to orc_simulate's.  state [nb][13] and lam [ml] are updated in place; out[0..4] = adds, muls, divs, sqrts,
// compares summed over the steps, out[5] = total Newton iterations.  flops per sim step = (out[0] + .. + out[3]) / steps.
int orc_count_flops_rollout(void* h, double* state, double* lam, const orc_ctrl* C, int sys, int k0, int steps, double eps, int newton_iter,
                            int line_iter, long long* out) {
    Mech& M = *(Mech*)h;
    Work Wd(M);
    WorkT<Count> W(M);
    FlopTally sum;
    long long iters = 0;
    for (int k = 0; k < steps; k++) {
        knot_from_state(M, state, Wd.z.data());
        control(M, Wd, C, sys, k0 + k, state);
        for (int i = 0; i < 13 * M.nb; i++) W.z[i] = Count(Wd.z[i]);
        for (int b = 0; b < M.nb; b++) for (int i = 0; i < 6; i++) W.s[6 * b + i] = Count(state[13 * b + 7 + i]);
        for (int i = 0; i < M.ml; i++) W.lam[i] = Count(lam[i]);
        for (int e = 0; e < M.ne; e++) W.u[e] = Count(Wd.u[e]);
        flop_tally() = FlopTally();
        int it = newton<Count>(M, W, eps, newton_iter, line_iter, 1);
        const FlopTally& t = flop_tally();
        sum.add += t.add; sum.mul += t.mul; sum.div += t.div; sum.sqrt_ += t.sqrt_; sum.cmp += t.cmp;
        if (it < 0) return it;
        iters += it;
        for (int b = 0; b < M.nb; b++) {
            for (int i = 0; i < 7; i++) state[13 * b + i] = Wd.z[13 * b + i];
            for (int i = 0; i < 6; i++) state[13 * b + 7 + i] = W.s[6 * b + i].v;
        }
        for (int i = 0; i < M.ml; i++) lam[i] = W.lam[i].v;
    }
    out[0] = sum.add; out[1] = sum.mul; out[2] = sum.div; out[3] = sum.sqrt_; out[4] = sum.cmp; out[5] = iters;
    return 0;
}

// batched LQR synthesis for the CPU baseline: B independent (A,Bu,Bl,G,Q,R) problems over threads
void orc_dlqr_batch(int B, const double* A, const double* Bu, const double* Bl, const double* G, const double* Q, const double* R, int n,
                    int mu, int ml, int N, double* Ku, int* kexit, int* iters, int nthreads) {
    auto run = [&](int lo, int hi) {
        for (int s = lo; s < hi; s++)
            kexit[s] = orc_dlqr(A + (size_t)s * n * n, Bu + (size_t)s * n * mu, Bl + (size_t)s * n * ml, G + (size_t)s * ml * n,
                                Q + (size_t)s * n * n, R + (size_t)s * mu * mu, n, mu, ml, N, Ku + (size_t)s * (N - 1) * mu * n, iters + s, nullptr);
    };
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(run, (int)((long long)B * t / nthreads), (int)((long long)B * (t + 1) / nthreads));
    for (auto& t : th) t.join();
}

}  // extern "C"
