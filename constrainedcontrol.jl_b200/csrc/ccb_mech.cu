// csrc/ccb_mech.cu -- host side of the mechanism handle: validation, node ids, spanning-tree order of the
// mechanism graph (the GraphBasedSystems `dfs_list` / `parents` role), the symbolic plan of
// the device Schur sweep, and the upload of the flattened tables.
//
// Replaces, for the hot path, what `Mechanism(origin, bodies, eqcs; g, Δt)` sets up upstream
// (reference examples/lqr_cartpole.jl:32; ids via getid, reference src/control/lqr_tracking.jl:74).
#include <algorithm>
#include <cmath>
#include <cstring>
#include "ccb_internal.h"

namespace ccb {

static thread_local std::string g_err;
int64_t g_launches = 0;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }
int cuda_fail(cudaError_t e, const char* what) {
    set_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
    return CCB_E_CUDA;
}

// rows of the projector for a joint part constraining n of 3 directions about `axis` (unit):
// n=3 identity, n=2 two unit vectors orthogonal to the axis, n=1 the axis itself.
static void projector_rows(int n, const double* ax, double out[3][3]) {
    std::memset(out, 0, sizeof(double) * 9);
    if (n == 3) { out[0][0] = out[1][1] = out[2][2] = 1.0; return; }
    if (n == 1) { for (int i = 0; i < 3; i++) out[0][i] = ax[i]; return; }
    if (n == 2) {
        // helper direction least aligned with the axis, Gram-Schmidt, then axis x v1
        double h[3] = {0, 0, 0};
        h[std::fabs(ax[0]) < 0.9 ? 0 : 1] = 1.0;
        double dp = h[0] * ax[0] + h[1] * ax[1] + h[2] * ax[2];
        double v1[3], nrm = 0;
        for (int i = 0; i < 3; i++) { v1[i] = h[i] - dp * ax[i]; nrm += v1[i] * v1[i]; }
        nrm = std::sqrt(nrm);
        for (int i = 0; i < 3; i++) v1[i] /= nrm;
        out[0][0] = v1[0]; out[0][1] = v1[1]; out[0][2] = v1[2];
        out[1][0] = ax[1] * v1[2] - ax[2] * v1[1];
        out[1][1] = ax[2] * v1[0] - ax[0] * v1[2];
        out[1][2] = ax[0] * v1[1] - ax[1] * v1[0];
    }
}

int build_host_mech(const ccb_mech_desc* d, HostMech* M) {
    if (!d || d->nb < 1 || d->ne < 0 || d->nj < d->ne) { set_error("ccb_mech_create: bad sizes"); return CCB_E_INVALID; }
    if (!(d->dt > 0)) { set_error("ccb_mech_create: dt must be positive"); return CCB_E_INVALID; }
    M->nb = d->nb; M->ne = d->ne; M->nj = d->nj; M->dt = d->dt; M->g = d->g; M->reg = d->reg; M->id_scheme = d->id_scheme;
    M->mass.assign(d->mass, d->mass + d->nb);
    M->inertia.assign(d->inertia, d->inertia + 9 * d->nb);
    M->eqc_parent.assign(d->eqc_parent, d->eqc_parent + d->ne);
    M->eqc_ptr.assign(d->eqc_joint_ptr, d->eqc_joint_ptr + d->ne + 1);
    if (M->eqc_ptr[0] != 0 || M->eqc_ptr[d->ne] != d->nj) { set_error("ccb_mech_create: eqc_joint_ptr does not cover the joints"); return CCB_E_INVALID; }
    const int nj = d->nj, nb = d->nb, ne = d->ne;
    M->j_eqc.resize(nj); M->j_parent.resize(nj); M->j_child.resize(nj); M->j_nT.resize(nj); M->j_nR.resize(nj);
    M->j_dim.resize(nj); M->j_row0.resize(nj);
    M->j_axis.resize(3 * nj); M->j_p1.assign(d->joint_p1, d->joint_p1 + 3 * nj); M->j_p2.assign(d->joint_p2, d->joint_p2 + 3 * nj);
    M->j_qoff.assign(d->joint_qoff, d->joint_qoff + 4 * nj);
    M->eqc_dim.assign(ne, 0); M->eqc_row0.assign(ne, 0); M->act_joint.assign(ne, -1); M->act_kind.assign(ne, 0);
    int row = 0;
    for (int e = 0; e < ne; e++) {
        if (M->eqc_parent[e] < -1 || M->eqc_parent[e] >= nb) { set_error("ccb_mech_create: eqc parent out of range"); return CCB_E_INVALID; }
        M->eqc_row0[e] = row;
        std::vector<int> kids;
        for (int j = M->eqc_ptr[e]; j < M->eqc_ptr[e + 1]; j++) {
            int c = d->joint_child[j], nT = d->joint_nT[j], nR = d->joint_nR[j];
            if (c < 0 || c >= nb || nT < 0 || nT > 3 || nR < 0 || nR > 3 || nT + nR == 0) { set_error("ccb_mech_create: bad joint"); return CCB_E_INVALID; }
            if (c == M->eqc_parent[e]) { set_error("ccb_mech_create: joint connects a body to itself"); return CCB_E_INVALID; }
            M->j_eqc[j] = e; M->j_parent[j] = M->eqc_parent[e]; M->j_child[j] = c; M->j_nT[j] = nT; M->j_nR[j] = nR;
            M->j_dim[j] = nT + nR; M->j_row0[j] = row;
            const double* a = d->joint_axis + 3 * j;
            double nrm = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            double ax[3] = {1, 0, 0};
            if (nrm > 0) { ax[0] = a[0] / nrm; ax[1] = a[1] / nrm; ax[2] = a[2] / nrm; }
            for (int i = 0; i < 3; i++) M->j_axis[3 * j + i] = ax[i];
            double PT[3][3], PR[3][3];
            projector_rows(nT, ax, PT);
            projector_rows(nR, ax, PR);
            for (int r = 0; r < nT; r++) { M->row_joint.push_back(j); M->row_kind.push_back(0); for (int i = 0; i < 3; i++) M->row_proj.push_back(PT[r][i]); }
            for (int r = 0; r < nR; r++) { M->row_joint.push_back(j); M->row_kind.push_back(1); for (int i = 0; i < 3; i++) M->row_proj.push_back(PR[r][i]); }
            row += nT + nR;
            if (std::find(kids.begin(), kids.end(), c) == kids.end()) kids.push_back(c);
        }
        M->eqc_dim[e] = row - M->eqc_row0[e];
        // reference src/control/pid.jl:31-37: a controllable eqc has 6*|unique(childids)| - length(eqc) == 1
        if (M->eqc_ptr[e + 1] - M->eqc_ptr[e] == 1 && 6 * (int)kids.size() - M->eqc_dim[e] == 1) {
            int j = M->eqc_ptr[e];
            M->act_joint[e] = j;
            M->act_kind[e] = (M->j_nT[j] == 2) ? 1 : 2;
        }
    }
    M->ml = row;

    // ---- mixed-node graph: ids, iterative depth-first search (ascending ids), post-order = dfs_list
    const int N = ne + nb;
    M->nnodes = N;
    M->eqc_id.resize(ne); M->body_id.resize(nb);
    for (int e = 0; e < ne; e++) M->eqc_id[e] = (M->id_scheme == 0) ? e + 1 : nb + e + 1;
    for (int b = 0; b < nb; b++) M->body_id[b] = (M->id_scheme == 0) ? ne + b + 1 : b + 1;
    std::vector<std::vector<int>> nbrs(N + 1);   // 1-based ids
    auto link = [&](int u, int v) {
        if (std::find(nbrs[u].begin(), nbrs[u].end(), v) != nbrs[u].end()) return;
        nbrs[u].push_back(v); nbrs[v].push_back(u);
    };
    for (int j = 0; j < nj; j++) {
        int e = M->eqc_id[M->j_eqc[j]];
        if (M->j_parent[j] >= 0) link(e, M->body_id[M->j_parent[j]]);
        link(e, M->body_id[M->j_child[j]]);
    }
    for (int v = 1; v <= N; v++) std::sort(nbrs[v].begin(), nbrs[v].end());
    M->parents.assign(N, 0);
    std::vector<int> colour(N + 1, 0), cursor(N + 1, 0), stack;
    // The origin is the root of the mechanism graph: the search starts at the equality constraints attached to it (ascending ids), whatever
    // their position in the list; components without an origin joint follow in ascending id order (same rule as the oracle: bit-exact test).
    std::vector<int> roots;
    {
        std::vector<char> at_origin(ne, 0);
        for (int j = 0; j < nj; j++) if (M->j_parent[j] < 0) at_origin[M->j_eqc[j]] = 1;
        std::vector<int> ids;
        for (int e = 0; e < ne; e++) if (at_origin[e]) ids.push_back(M->eqc_id[e]);
        std::sort(ids.begin(), ids.end());
        roots = ids;
        for (int v = 1; v <= N; v++) roots.push_back(v);
    }
    for (int root : roots) {
        if (colour[root]) continue;
        stack.push_back(root); colour[root] = 1;
        while (!stack.empty()) {
            int v = stack.back();
            if (cursor[v] < (int)nbrs[v].size()) {
                int w = nbrs[v][cursor[v]++];
                if (w == M->parents[v - 1]) continue;
                if (colour[w] == 0) { colour[w] = 1; M->parents[w - 1] = v; stack.push_back(w); }
                else if (colour[w] == 1) { M->loop_edges.push_back(v); M->loop_edges.push_back(w); }
            } else {
                colour[v] = 2; M->dfs_list.push_back(v); stack.pop_back();
            }
        }
    }

    // ---- joint-level plan for the device: eliminate bodies first (block diagonal), then sweep the joints
    // in the order their eqcs appear in dfs_list (leaf first); symbolic fill for closed chains.
    std::vector<int> id_to_eqc(N + 1, -1);
    for (int e = 0; e < ne; e++) id_to_eqc[M->eqc_id[e]] = e;
    for (int id : M->dfs_list) {
        int e = id_to_eqc[id];
        if (e < 0) continue;
        for (int j = M->eqc_ptr[e]; j < M->eqc_ptr[e + 1]; j++) M->elim.push_back(j);
    }
    std::vector<int> pos(nj);
    for (int i = 0; i < nj; i++) pos[M->elim[i]] = i;
    auto pairs_of = [&](int a, int b) {   // shared-body mask between joints a and b
        int sa[2] = {M->j_parent[a], M->j_child[a]}, sb[2] = {M->j_parent[b], M->j_child[b]};
        int mask = 0;
        for (int x = 0; x < 2; x++) for (int y = 0; y < 2; y++) if (sa[x] >= 0 && sa[x] == sb[y]) mask |= 1 << (2 * x + y);
        return mask;
    };
    std::vector<char> nz((size_t)nj * nj, 0);
    for (int a = 0; a < nj; a++) for (int b = 0; b < nj; b++) nz[(size_t)a * nj + b] = (a == b) || pairs_of(a, b) != 0;
    for (int i = 0; i < nj; i++) {
        int c = M->elim[i];
        std::vector<int> lat;
        for (int p = 0; p < nj; p++) if (pos[p] > i && nz[(size_t)p * nj + c]) lat.push_back(p);
        for (int p : lat) for (int q : lat) nz[(size_t)p * nj + q] = 1;
    }
    // row storage: self block first, then every other coupled joint in ascending index
    M->j_soff.resize(nj); M->j_ncol.resize(nj); M->j_self.assign(nj, 0);
    std::vector<std::vector<int>> cpos(nj, std::vector<int>(nj, -1));
    int soff = 0;
    for (int j = 0; j < nj; j++) {
        int c = 0;
        cpos[j][j] = 0; c += M->j_dim[j];
        for (int q = 0; q < nj; q++) if (q != j && nz[(size_t)j * nj + q]) { cpos[j][q] = c; c += M->j_dim[q]; }
        M->j_ncol[j] = c; M->j_soff[j] = soff; soff += c * M->j_dim[j];
    }
    M->S_doubles = soff;
    M->nbr_ptr.assign(nj + 1, 0);
    for (int j = 0; j < nj; j++) {
        for (int q = 0; q < nj; q++) {
            int mask = pairs_of(j, q);
            if (q == j || mask) { M->nbr_joint.push_back(q); M->nbr_cpos.push_back(cpos[j][q]); M->nbr_pairs.push_back(mask); }
        }
        M->nbr_ptr[j + 1] = (int)M->nbr_joint.size();
    }
    M->acol_ptr.assign(nj + 1, 0); M->lat_ptr.assign(nj + 1, 0);
    for (int c = 0; c < nj; c++) {
        std::vector<int> lat;
        for (int p = 0; p < nj; p++) if (p != c && pos[p] > pos[c] && nz[(size_t)p * nj + c]) lat.push_back(p);
        std::sort(lat.begin(), lat.end(), [&](int a, int b) { return pos[a] < pos[b]; });
        std::vector<std::pair<int, int>> act;   // (owner joint q, offset t) per active scalar column
        for (int q : lat) for (int t = 0; t < M->j_dim[q]; t++) { M->acol.push_back(cpos[c][q] + t); M->acol_row.push_back(M->j_row0[q] + t); act.push_back({q, t}); }
        M->acol_ptr[c + 1] = (int)M->acol.size();
        for (int p : lat) {
            M->lat_p.push_back(p); M->lat_pc0.push_back(cpos[p][c]); M->lat_map.push_back((int)M->map.size());
            for (auto& qt : act) M->map.push_back(cpos[p][qt.first] + qt.second);
        }
        M->lat_ptr[c + 1] = (int)M->lat_p.size();
    }
    // rows incident to each body
    M->binc_ptr.assign(nb + 1, 0);
    for (int b = 0; b < nb; b++) {
        for (int r = 0; r < M->ml; r++) {
            int j = M->row_joint[r];
            if (M->j_parent[j] == b) { M->binc_row.push_back(r); M->binc_side.push_back(0); }
            if (M->j_child[j] == b) { M->binc_row.push_back(r); M->binc_side.push_back(1); }
        }
        M->binc_ptr[b + 1] = (int)M->binc_row.size();
    }
    // joints incident to each body and the bodies it is coupled to through them
    M->bj_ptr.assign(nb + 1, 0); M->cb_ptr.assign(nb + 1, 0);
    for (int b = 0; b < nb; b++) {
        std::vector<int> cb = {b};
        for (int j = 0; j < nj; j++) {
            int side = M->j_parent[j] == b ? 0 : (M->j_child[j] == b ? 1 : -1);
            if (side < 0) continue;
            M->bj_joint.push_back(j); M->bj_side.push_back(side);
            int other = side == 0 ? M->j_child[j] : M->j_parent[j];
            if (other >= 0 && std::find(cb.begin(), cb.end(), other) == cb.end()) cb.push_back(other);
        }
        M->bj_ptr[b + 1] = (int)M->bj_joint.size();
        std::sort(cb.begin() + 1, cb.end());
        M->cb_body.insert(M->cb_body.end(), cb.begin(), cb.end());
        M->cb_ptr[b + 1] = (int)M->cb_body.size();
    }
    M->is_chain = chain_tables(*M, M->chain_parent, M->chain_child, M->chain_depth, M->chain_max_depth, M->chain_max_dim);
    // default lanes per system: enough for one pass over the constraint rows, at most a warp
    int want = std::max(M->ml, 2 * nb);
    M->lanes = want <= 12 ? 8 : want <= 24 ? 16 : 32;   // measured: fewer lanes per system win (issue-bound)
    return CCB_OK;
}

int get_dev_mech(HostMech* M, int device, const DevMech** view) {
    auto it = M->dev.find(device);
    if (it != M->dev.end()) { *view = &it->second.view; return CCB_OK; }
    CCB_CUDA(cudaSetDevice(device));
    std::vector<int> ib;
    std::vector<double> db;
    auto addi = [&](const std::vector<int>& v) { size_t o = ib.size(); ib.insert(ib.end(), v.begin(), v.end()); if (v.empty()) ib.push_back(0); return o; };
    auto addd = [&](const std::vector<double>& v) { size_t o = db.size(); db.insert(db.end(), v.begin(), v.end()); if (v.empty()) db.push_back(0); return o; };
    size_t o_binc_ptr = addi(M->binc_ptr), o_binc_row = addi(M->binc_row), o_binc_side = addi(M->binc_side), o_row_joint = addi(M->row_joint),
           o_row_kind = addi(M->row_kind), o_jp = addi(M->j_parent), o_jc = addi(M->j_child), o_jr0 = addi(M->j_row0), o_jd = addi(M->j_dim),
           o_soff = addi(M->j_soff), o_ncol = addi(M->j_ncol), o_self = addi(M->j_self), o_nbp = addi(M->nbr_ptr), o_nbj = addi(M->nbr_joint),
           o_nbc = addi(M->nbr_cpos), o_nbm = addi(M->nbr_pairs), o_elim = addi(M->elim), o_acp = addi(M->acol_ptr), o_ac = addi(M->acol), o_acr = addi(M->acol_row),
           o_lp = addi(M->lat_ptr), o_lpp = addi(M->lat_p), o_lpc = addi(M->lat_pc0), o_lm = addi(M->lat_map), o_map = addi(M->map),
           o_aj = addi(M->act_joint), o_ak = addi(M->act_kind), o_bjp = addi(M->bj_ptr), o_bjj = addi(M->bj_joint), o_bjs = addi(M->bj_side),
           o_cbp = addi(M->cb_ptr), o_cbb = addi(M->cb_body), o_chp = addi(M->chain_parent), o_chc = addi(M->chain_child),
           o_chd = addi(M->chain_depth);
    size_t o_mass = addd(M->mass), o_in = addd(M->inertia), o_rp = addd(M->row_proj), o_p1 = addd(M->j_p1), o_p2 = addd(M->j_p2),
           o_qo = addd(M->j_qoff), o_ax = addd(M->j_axis);
    HostMech::Upload up;
    CCB_CUDA(cudaMalloc(&up.ibuf, ib.size() * sizeof(int)));
    CCB_CUDA(cudaMalloc(&up.dbuf, db.size() * sizeof(double)));
    CCB_CUDA(cudaMemcpy(up.ibuf, ib.data(), ib.size() * sizeof(int), cudaMemcpyHostToDevice));
    CCB_CUDA(cudaMemcpy(up.dbuf, db.data(), db.size() * sizeof(double), cudaMemcpyHostToDevice));
    const int* I = (const int*)up.ibuf;
    const double* D = (const double*)up.dbuf;
    DevMech& v = up.view;
    v.nb = M->nb; v.ne = M->ne; v.nj = M->nj; v.ml = M->ml; v.dt = M->dt; v.g = M->g; v.reg = M->reg; v.nu_all = 0;
    v.mass = D + o_mass; v.inertia = D + o_in; v.row_proj = D + o_rp; v.j_p1 = D + o_p1; v.j_p2 = D + o_p2; v.j_qoff = D + o_qo; v.j_axis = D + o_ax;
    v.binc_ptr = I + o_binc_ptr; v.binc_row = I + o_binc_row; v.binc_side = I + o_binc_side; v.row_joint = I + o_row_joint; v.row_kind = I + o_row_kind;
    v.j_parent = I + o_jp; v.j_child = I + o_jc; v.j_row0 = I + o_jr0; v.j_dim = I + o_jd; v.j_soff = I + o_soff; v.j_ncol = I + o_ncol; v.j_self = I + o_self;
    v.nbr_ptr = I + o_nbp; v.nbr_joint = I + o_nbj; v.nbr_cpos = I + o_nbc; v.nbr_pairs = I + o_nbm; v.elim = I + o_elim; v.acol_ptr = I + o_acp;
    v.acol = I + o_ac; v.acol_row = I + o_acr; v.lat_ptr = I + o_lp; v.lat_p = I + o_lpp; v.lat_pc0 = I + o_lpc; v.lat_map = I + o_lm; v.map = I + o_map;
    v.eqc_act_joint = I + o_aj; v.eqc_act_kind = I + o_ak;
    v.bj_ptr = I + o_bjp; v.bj_joint = I + o_bjj; v.bj_side = I + o_bjs; v.cb_ptr = I + o_cbp; v.cb_body = I + o_cbb;
    v.is_chain = M->is_chain ? 1 : 0; v.chain_max_depth = M->chain_max_depth; v.chain_max_dim = M->chain_max_dim;
    v.chain_parent = I + o_chp; v.chain_child = I + o_chc; v.chain_depth = I + o_chd;
    v.rows_ok = (M->is_chain && M->nj >= 1 && M->nj <= 6 && M->chain_max_dim <= 5 && M->chain_max_depth == M->nj - 1) ? 1 : 0;
    for (int j = 0; j < 6; j++) v.rows_order[j] = 0;
    if (v.rows_ok) for (int j = 0; j < M->nj; j++) v.rows_order[M->chain_depth[j]] = j;
    // shared-memory layout of one system
    int o = 0;
    const int nb = M->nb, ml = M->ml;
    v.o_z = o; o += 13 * nb;      // x2 q2 v1 w1
    v.o_s = o; o += 6 * nb;       // v2 w2 (current iterate)
    v.o_s0 = o; o += 6 * nb;      // iterate before the line search / Newton step
    v.o_p3 = o; o += 7 * nb;      // x3 q3
    v.o_dinv = o; o += 9 * nb;    // inverse of the rotational 3x3 of D_b
    v.o_tw = o; o += 9 * nb;      // T_w(w2)
    v.o_d = o; o += 6 * nb;       // body residual d_b, later the body step
    v.o_fext = o; o += 6 * nb;    // input wrench (F world, tau body)
    v.o_lam = o; o += ml;
    v.o_lam0 = o; o += ml;
    v.o_rhs = o; o += ml;         // g, then Schur rhs, then the lambda step
    v.o_gk = o; o += 12 * ml;     // G rows at the knot (x2,q2): parent 6 | child 6
    v.o_S = o; o += M->S_doubles;
    v.o_u = o; o += M->ne > 0 ? M->ne : 1;
    v.o_dz = o; o += 12 * nb;
    v.o_ds = o; o += 6 * nb;      // Newton step of the bodies
    v.o_dl = o; o += ml;          // Newton step of the multipliers
    v.smem_doubles = o;
    v.S_doubles = M->S_doubles;
    M->dev[device] = up;
    *view = &M->dev[device].view;
    return CCB_OK;
}

void free_dev_mech(HostMech* M) {
    for (auto& kv : M->dev) {
        cudaSetDevice(kv.first);
        cudaFree(kv.second.ibuf);
        cudaFree(kv.second.dbuf);
    }
    M->dev.clear();
}

}  // namespace ccb
