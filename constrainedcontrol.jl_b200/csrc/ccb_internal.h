// csrc/ccb_internal.h -- internal structures of libccb.so (host side of the B200 hot path)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include "../../include/ccb.h"

namespace ccb {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define CCB_CUDA(call)                                              \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ::ccb::cuda_fail(_e, #call);  \
    } while (0)

extern int64_t g_launches;
int finish_uploads();   // ccb_api.cu: legacy-stream uploads of a create function have landed (batch streams are non-blocking)

// ---------------------------------------------------------------------------------------------
// Flattened, device-resident description of one mechanism + the symbolic plan of its Schur sweep.
// All pointers are device pointers into ONE allocation (ints) + ONE allocation (doubles).
// ---------------------------------------------------------------------------------------------
struct DevMech {
    int nb, ne, nj, ml, nu_all;
    double dt, g, reg;
    // bodies
    const double* mass;        // [nb]
    const double* inertia;     // [nb][9]
    const int* binc_ptr;       // [nb+1]  rows incident to body b ...
    const int* binc_row;       //         ... row index
    const int* binc_side;      //         ... 0: b is the parent of that row's joint, 1: the child
    // rows (one per scalar constraint = one lambda component)
    const int* row_joint;      // [ml]
    const int* row_kind;       // [ml] 0 translational, 1 rotational
    const double* row_proj;    // [ml][3]
    // joints
    const int* j_parent;       // [nj] body or -1
    const int* j_child;        // [nj]
    const int* j_row0;         // [nj]
    const int* j_dim;          // [nj]
    const double* j_p1;        // [nj][3]
    const double* j_p2;        // [nj][3]
    const double* j_qoff;      // [nj][4]
    const double* j_axis;      // [nj][3] normalised
    // Schur complement storage: row i of joint j lives at S + j_soff[j] + i * j_ncol[j]
    const int* j_soff;         // [nj]
    const int* j_ncol;         // [nj]
    const int* j_self;         // [nj] column position of the joint's own block
    // assembly plan: structural neighbours (sharing a body), self included
    const int* nbr_ptr;        // [nj+1]
    const int* nbr_joint;
    const int* nbr_cpos;       // column position of the neighbour's block in this joint's rows
    const int* nbr_pairs;      // bit (2*side_self + side_nbr) set if those two sides are the same body
    // elimination plan
    const int* elim;           // [nj] joint order (leaf first)
    const int* acol_ptr;       // [nj+1] active (= later) scalar columns of joint c in its own row storage
    const int* acol;
    const int* acol_row;       //   global lambda row that owns that column
    const int* lat_ptr;        // [nj+1] later neighbours of joint c
    const int* lat_p;          //   the later joint p
    const int* lat_pc0;        //   column position of c's block inside p's rows
    const int* lat_map;        //   offset into map[]: for each active column i of c, its position inside p's rows
    const int* map;
    // body -> incident joints, and body -> coupled bodies (self first): block-sparsity of the linearisation
    const int* bj_ptr;         // [nb+1]
    const int* bj_joint;
    const int* bj_side;
    const int* cb_ptr;         // [nb+1]
    const int* cb_body;
    // serial-chain fast path (ccb_rollout_chain.cu): parent / child joint and depth of every joint
    int is_chain, chain_max_depth, chain_max_dim;
    int rows_ok, rows_order[6];   // row-per-lane fast path (ccb_rollout_rows.cu): joint at every chain position
    const int* chain_parent;
    const int* chain_child;
    const int* chain_depth;
    // actuation (1-DoF eqcs)
    const int* eqc_act_joint;  // [ne] or -1
    const int* eqc_act_kind;   // [ne] 0 none, 1 force, 2 torque
    // shared-memory layout of one system (offsets in doubles)
    int o_z, o_s, o_s0, o_p3, o_dinv, o_tw, o_d, o_fext, o_lam, o_lam0, o_rhs, o_gk, o_S, o_u, o_dz, o_ds, o_dl, smem_doubles;
    int S_doubles;
};

struct HostMech {
    // description (host copies)
    int nb = 0, ne = 0, nj = 0, ml = 0, id_scheme = 0;
    double dt = 0.01, g = -9.81, reg = 0.0;
    std::vector<double> mass, inertia, j_axis, j_p1, j_p2, j_qoff, row_proj;
    std::vector<int> eqc_parent, eqc_ptr, eqc_dim, eqc_row0, act_joint, act_kind;
    std::vector<int> j_eqc, j_parent, j_child, j_nT, j_nR, j_dim, j_row0, row_joint, row_kind;
    // mixed-node graph (bit-exact contract)
    int nnodes = 0;
    std::vector<int> eqc_id, body_id, dfs_list, parents, loop_edges;
    // joint-level plan
    std::vector<int> elim, j_soff, j_ncol, j_self, nbr_ptr, nbr_joint, nbr_cpos, nbr_pairs, acol_ptr, acol, acol_row, lat_ptr, lat_p, lat_pc0, lat_map,
        map, binc_ptr, binc_row, binc_side, bj_ptr, bj_joint, bj_side, cb_ptr, cb_body;
    int S_doubles = 0;
    bool is_chain = false;
    int chain_max_depth = 0, chain_max_dim = 0;
    std::vector<int> chain_parent, chain_child, chain_depth;
    int lanes = 8;   // default lanes per system
    // per-device uploads
    struct Upload { DevMech view; void* ibuf = nullptr; void* dbuf = nullptr; };
    std::map<int, Upload> dev;
};

int build_host_mech(const ccb_mech_desc* d, HostMech* out);
int get_dev_mech(HostMech* m, int device, const DevMech** view);
void free_dev_mech(HostMech* m);

struct DevCtrl {
    int kind, nu, N;
    const int* eqc;
    const double* K; long long K_bs;
    const double* zd; long long zd_bs;
    const double* Fd; long long Fd_bs;
    const double* U; long long U_bs;
    const double *P, *I, *D;
    const double* goal; long long goal_bs;
    double* pid_int;
    double* pid_last;
    int sys_div;        // LQR-type tables: system s uses table s / sys_div (0 or 1: one table per system or one shared table)
};

struct SimArgs {
    double* state;      // [B][nb][13]
    double* lambda;     // [B][ml]
    double* storage;    // [steps][B][nb][13] or null
    int* status;        // [B] or null
    int B, k0, steps;
    double eps;
    int newton_iter, line_iter;
    int knot_input;     // 1: `state` already holds the knot (x2,q2,v1,w1) of a setpoint; one Newton solve, state untouched (linearsystem)
    double* s_out;      // [B][nb][6] solution (v2,w2) when knot_input
    int spw;            // chain kernel: systems per warp (0 = automatic)
    int status_accum;   // 1: this launch continues a rollout split over several launches: fold its counts into what `status` already holds
    // persistent time-sliced launch of K1r (ccb_rollout_rows.cu): persist[0] = work counter, persist[1 + t] = chunks of triple t done (zeroed by
    // the launcher); nullptr: plain launch
    int* persist;
    int nchunks, chunk_steps;
};

// status[s] = total Newton iterations of the rollout, negated if any step failed (0: failed before the first iteration); a continuing
// launch (status_accum) adds to the earlier launches' count and keeps their failure flag
__device__ __forceinline__ void write_status(const SimArgs& A, int sys, int total, bool failed, bool later_chunk = false) {
    if (A.status_accum || later_chunk) {
        const int o = __ldcg(A.status + sys);
        if (o <= 0) failed = true;
        total += o < 0 ? -o : o;
    }
    A.status[sys] = failed ? -total : total;
}

// lanes: 0 = automatic (row-per-lane kernel K1r for short serial chains, else the chain kernel K1c for serial chains, else the
// multi-warp rows kernel K1m for longer serial chains, else the general kernel K1 with default_lanes), 5 = force K1r, 6 = force K1m,
// 1 = force K1c, 8/16/32 = K1 with that many lanes per system
int launch_rollout(const DevMech& M, const DevCtrl& C, const SimArgs& a, int lanes, int default_lanes, cudaStream_t stream);
bool chain_tables(const HostMech& H, std::vector<int>& parent, std::vector<int>& child, std::vector<int>& depth, int& max_depth, int& max_dim);
int launch_rollout_chain(const DevMech& M, const DevCtrl& C, const SimArgs& a, const int* d_parent, const int* d_child, const int* d_depth,
                         int max_depth, int max_dim, cudaStream_t stream);
int rollout_smem_bytes(const DevMech& M, int lanes);
int launch_rollout_rows(const DevMech& M, const DevCtrl& C, const SimArgs& a, cudaStream_t stream);
bool mw_kernel_eligible(const DevMech& M);
// Extra destinations of the gain rows a Riccati sweep produces: device tables laid out [S_total][N-1][nu][n] (this rank's own and
// peer GPUs' tables mapped through CUDA IPC); system s of the launch goes to index s_offset + s of every table.  The all-gather of
// SURVEY 8(e) fused into the sweep: the rows leave over NVLink while the recursion runs.  Set for the calling thread by
// ccb_lqr_synthesize_peers (ccb_synth.cu), read by the K4 launcher; n == 0: none.
struct PeerTables { double* t[8]; int n; long long s_offset; int spread; int defer_backfill; };
// spread: stagger the CTAs' flush phases and rotate their peer order.  defer_backfill: push only the rows the recursion produced (k >= kexit); the
// rows below the break index (copies of row kexit - 1, lqr.jl:179-181) are written by every GPU into its own table afterwards (ccb_gains_backfill)
extern thread_local PeerTables g_peer_tables;
// K5 (ccb_are.cu): kind 0 DARE, 1 CARE; device pointers; ws >= S * are_ws_doubles(n) doubles
size_t are_ws_doubles(int n);
int are_sda_device(int kind, int S, int n, int m, const double* A, const double* B, const double* Q, const double* R, double tol, int maxit, double* X,
                   double* K, int* iters, int* status, double* ws, cudaStream_t stream);
int launch_rollout_mw(const DevMech& M, const DevCtrl& C, const SimArgs& a, cudaStream_t stream);

}  // namespace ccb

struct ccb_mech { ccb::HostMech h; };
struct ccb_batch {
    ccb_mech* mech;
    int B, device;
    double* state = nullptr;
    double* lambda = nullptr;
    cudaStream_t stream = nullptr;
    double* storage_dev = nullptr; size_t storage_cap = 0;
    int* status_dev = nullptr;
    // record = true pipeline (ccb_simulate): device double buffer of trajectory chunks, drained by a copy stream while the next chunk computes
    int* persist_buf = nullptr; size_t persist_cap = 0;      // work counter + per-triple chunk flags of K1r's persistent launches
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_kernel[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
};
struct ccb_ctrl {
    ccb::DevCtrl v;
    int device, B;
    std::vector<void*> allocs;
};
