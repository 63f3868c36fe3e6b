/* include/ccb.h -- C-ABI of libccb.so: B200-native batched closed-loop rollouts and LQR gain synthesis
 * behind the ConstrainedControl.jl controller API.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own (it is 472 lines of
 * pure Julia calling ConstrainedDynamics.jl in-process), so each entry point below names the Julia call it
 * replaces for a *batch* of mechanism states; the Julia shim that `ccall`s them is shown in INTEGRATION.md
 * and shipped as constrainedcontrol.jl_b200/julia/ConstrainedControlB200.jl.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns 0 on success, <0 on error (CCB_E_*), never
 *     throws; the message of the last error on the calling thread is `ccb_last_error()`.
 *   - opaque handles are owned by the library; free with the matching *_destroy (Julia: `finalizer`).
 *   - host buffers passed in are borrowed for the duration of the call; outputs go to caller-allocated
 *     buffers.  All matrices are row-major, indices 0-based, node ids 1-based (as in Julia).
 *   - body state layout is 13 doubles per body: xc[3], qc[4] (scalar first), vc[3], wc[3]
 *     (`body.state.{xc,qc,vc,ωc}`, reference examples/lqr_sawyer.jl:16-17); a system is [nb][13].
 *   - tangent (gain / Δz) ordering per body is 12 = [δx, δv, δφ, δω] (reference src/control/lqr.jl:92-95).
 *   - there is NO CPU fallback: without a CUDA device every compute call returns CCB_E_CUDA.
 */
#ifndef CCB_H
#define CCB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCB_OK 0
#define CCB_E_INVALID -1  /* bad argument (the reference's @assert, lqr.jl:59-60, pid.jl:20,36) */
#define CCB_E_CUDA -2     /* CUDA runtime error or no device */
#define CCB_E_UNSUPPORTED -3
#define CCB_E_NCCL -4

const char* ccb_last_error(void);
int ccb_version(void);
/* number of visible CUDA devices (0 without a GPU; never an error) */
int ccb_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Mechanism description: replaces `Mechanism(origin, bodies, eqcs; g, Δt)` (reference
 * examples/lqr_cartpole.jl:32) for the purposes of the hot path.  An equality constraint (eqc) holds one
 * or more joints that share the parent body (examples/lqr_deltabot.jl:25); a joint constrains nT
 * translational and nR rotational directions (Revolute = 3+2, Prismatic = 2+3, Fixed = 3+3,
 * FixedOrientation = 0+3).
 * ------------------------------------------------------------------------------------------------ */
typedef struct ccb_mech_desc {
    int32_t nb;                   /* bodies */
    int32_t ne;                   /* equality constraints */
    int32_t nj;                   /* joints (>= ne) */
    int32_t id_scheme;            /* 0: eqc ids 1..ne, body ids ne+1..ne+nb ; 1: bodies first */
    double dt;                    /* Δt */
    double g;                     /* gravity along z (e.g. -9.81) */
    double reg;                   /* diagonal regularisation of the constraint Schur complement (0 = none) */
    const double* mass;           /* [nb] */
    const double* inertia;        /* [nb][9] row-major, body frame */
    const int32_t* eqc_parent;    /* [ne] body index, -1 = origin */
    const int32_t* eqc_joint_ptr; /* [ne+1] CSR: joints of eqc e are eqc_joint_ptr[e] .. eqc_joint_ptr[e+1]-1 */
    const int32_t* joint_child;   /* [nj] body index */
    const int32_t* joint_nT;      /* [nj] 0..3 */
    const int32_t* joint_nR;      /* [nj] 0..3 */
    const double* joint_axis;     /* [nj][3] free axis (parent frame) */
    const double* joint_p1;       /* [nj][3] vertex in the parent frame */
    const double* joint_p2;       /* [nj][3] vertex in the child frame */
    const double* joint_qoff;     /* [nj][4] orientation offset */
} ccb_mech_desc;

typedef struct ccb_mech ccb_mech;

int ccb_mech_create(const ccb_mech_desc* desc, ccb_mech** out);
int ccb_mech_destroy(ccb_mech* m);
/* out[0..7] = nb, ne, nj, ml (Σ joint dims), nnodes, nloops (non-tree edges), n (=12 nb), lanes per system */
int ccb_mech_dims(const ccb_mech* m, int32_t out[8]);
/* Integer graph data (the bit-exact part of the parity contract):
 *   eqc_id[ne], body_id[nb]      node ids
 *   dfs_list[nnodes]             elimination order, children before parents (GraphBasedSystems `dfs_list`)
 *   parents[nnodes]              spanning-tree parent id of node id-1, 0 for a root
 *   loop_edges[2*nloops]         non-tree edges (id pairs)
 *   eqc_dim[ne], eqc_row0[ne]    rows of λ per eqc and their offset
 *   act_kind[ne]                 0: not a 1-DoF eqc, 1: force input along the axis, 2: torque input
 *   joint_elim[nj]               joint elimination order used by the device Schur sweep
 * Any pointer may be NULL. */
int ccb_mech_graph(const ccb_mech* m, int32_t* eqc_id, int32_t* body_id, int32_t* dfs_list, int32_t* parents, int32_t* loop_edges,
                   int32_t* eqc_dim, int32_t* eqc_row0, int32_t* act_kind, int32_t* joint_elim);

/* ------------------------------------------------------------------------------------------------
 * Batch of B independent copies of one mechanism, resident on one CUDA device.
 * Replaces "an array of Mechanism objects" on the Julia side.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ccb_batch ccb_batch;

int ccb_batch_create(ccb_mech* m, int32_t B, int32_t device, ccb_batch** out);
int ccb_batch_destroy(ccb_batch* b);
/* state [B][nb][13]; lambda [B][ml] or NULL (= zeros: fresh mechanisms have λ = 0).  Host -> device. */
int ccb_batch_set_state(ccb_batch* b, const double* state, const double* lambda);
/* device -> host; either pointer may be NULL */
int ccb_batch_get_state(ccb_batch* b, double* state, double* lambda);
/* raw device pointers for callers that keep data resident (torch tensors, CuArrays): state [B][nb][13], lambda [B][ml] */
int ccb_batch_device_ptrs(ccb_batch* b, void** state_dev, void** lambda_dev);

/* ------------------------------------------------------------------------------------------------
 * Controllers.  Arbitrary Julia closures cannot run on the device, so `control!` is an enumerated law
 * with parameter tables:
 *   CCB_CTRL_NONE      uncontrolled
 *   CCB_CTRL_TABLE     open loop u[k]                 examples/trackingLQR_triple_cartpole.jl:45-47
 *   CCB_CTRL_LQR_INF   control_lqr!(…, ::LQR{T,Inf})  src/control/lqr.jl:116-139  (K[1] at every k)
 *   CCB_CTRL_LQR_FIN   control_lqr!(…, ::LQR{T,N})    src/control/lqr.jl:89-114   (K[k] while k < N)
 *   CCB_CTRL_TRACKING  control_trackinglqr!           src/control/lqr_tracking.jl:46-71
 *   CCB_CTRL_PID       control_pid!                   src/control/pid.jl:69-88
 * A *_bstride of 0 shares one table across the batch; otherwise it is the per-system stride in doubles.
 * ------------------------------------------------------------------------------------------------ */
enum { CCB_CTRL_NONE = 0, CCB_CTRL_TABLE = 1, CCB_CTRL_LQR_INF = 2, CCB_CTRL_LQR_FIN = 3, CCB_CTRL_TRACKING = 4, CCB_CTRL_PID = 5 };

typedef struct ccb_ctrl_desc {
    int32_t kind;
    int32_t nu;            /* controlled eqcs */
    int32_t N;             /* horizon in steps (LQR_FIN / TRACKING): gains have N-1 rows; TABLE: number of rows of U */
    int32_t B;             /* batch size the per-system tables are laid out for */
    const int32_t* eqc;    /* [nu] eqc indices (0-based, position in mechanism.eqconstraints) */
    const double* K;       /* [rows][nu][12 nb], rows = 1 (INF) or N-1 */
    int64_t K_bstride;
    const double* zd;      /* LQR: [nb][13] desired (xd,qd,vd,ωd) ; TRACKING: [N][nb][13] */
    int64_t zd_bstride;
    const double* Fd;      /* LQR: [nu] (Fτd) ; TRACKING: [N][nu] */
    int64_t Fd_bstride;
    const double* U;       /* TABLE: [N][nu] */
    int64_t U_bstride;
    const double* P;       /* PID gains [nu] */
    const double* I;
    const double* D;
    const double* goal;    /* [nu] */
    int64_t goal_bstride;
} ccb_ctrl_desc;

typedef struct ccb_ctrl ccb_ctrl;

/* copies every table to `device`; host pointers are not retained */
int ccb_ctrl_create(const ccb_mech* m, const ccb_ctrl_desc* desc, int32_t device, ccb_ctrl** out);
int ccb_ctrl_destroy(ccb_ctrl* c);
/* LQR / TrackingLQR controllers laid out for S tables: let `per_table` consecutive systems of a batch of S * per_table systems share
 * table s / per_table (e.g. 16 perturbed initial states per reference trajectory) instead of replicating it */
int ccb_ctrl_set_replication(ccb_ctrl* c, int32_t per_table);
/* PID controller memory (`integratederrors`, `lasterrors`; pid.jl:9-10) [B][nu]; persists across simulate calls (pid.jl:76) */
int ccb_ctrl_pid_state(ccb_ctrl* c, double* integrated, double* last, int32_t set /* 0: read, 1: write */);

/* ------------------------------------------------------------------------------------------------
 * simulate!(mechanism, steps, controller!; record)  (reference examples/lqr_pendulum.jl:43 and the
 * ConstrainedDynamics.jl time loop) for every system of the batch:
 *   for k = k0 .. k0+steps-1 : record, control!(k), applyFτ!, newton!(ε, newtonIter, lineIter), updatestate!
 * storage (host, may be NULL): [steps][B][nb][13] = (xc,qc,vc,ωc) at the start of each step (Storage{T,N}).
 * status (host, may be NULL): [B] total Newton iterations, negated if any step failed to converge.  A step whose residual has become NaN
 * (e.g. |ω| beyond 2/Δt) is abandoned at once -- it is flagged like a step that ran into newtonIter, its state is NaN as in the reference.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ccb_sim_opts {
    double eps;           /* newton! ε, default 1e-10 */
    int32_t newton_iter;  /* default 100 */
    int32_t line_iter;    /* default 10 */
    int32_t lanes;        /* kernel choice: 0 = automatic; 5 = K1r (one lane per constraint row, <= 6 joints); 6 = K1m (same, one CTA per
                             system, long chains); 1 = K1c (one lane per joint); 8/16/32 = general kernel K1 with that many lanes */
    int32_t reserved;     /* chain kernel: systems per warp, 0 = automatic (tuning knob) */
} ccb_sim_opts;

int ccb_simulate(ccb_batch* b, ccb_ctrl* ctrl /* may be NULL */, int32_t k0, int32_t steps, const ccb_sim_opts* opts /* may be NULL */,
                 double* storage, int32_t* status);
/* same, everything stays on the device: storage_dev / status_dev are device pointers (or NULL); asynchronous on the
 * batch's stream -- call ccb_batch_sync() before reading */
int ccb_simulate_device(ccb_batch* b, ccb_ctrl* ctrl, int32_t k0, int32_t steps, const ccb_sim_opts* opts, double* storage_dev,
                        int32_t* status_dev);
int ccb_batch_sync(ccb_batch* b);
/* the CUDA stream (cudaStream_t) the batch launches on -- time kernels with events recorded on THIS stream */
int ccb_batch_stream(ccb_batch* b, void** stream);
/* number of kernels launched by this library on the calling process so far (bench.py's gpu_launches) */
int64_t ccb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * linearsystem(mechanism, xd, vd, qd, ωd, Fτd, bodyids, eqcids) -> A, Bu, Bλ, G   (reference
 * src/control/lqr.jl:63, src/control/lqr_tracking.jl:88) for S setpoints at once.
 *   zd [S][nb][13], Fd [S][nu], eqc [nu]
 *   A [S][n][n], Bu [S][n][nu], Bl [S][n][ml], G [S][ml][n]  (n = 12 nb), lambda_d [S][ml]; any output may be NULL
 * ------------------------------------------------------------------------------------------------ */
int ccb_linearize(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu,
                  const ccb_sim_opts* opts, double* A, double* Bu, double* Bl, double* G, double* lambda_d);

/* ------------------------------------------------------------------------------------------------
 * dlqr(A, Bu, Bλ, G, Q, R, N)  (reference src/control/lqr.jl:141-184) for S independent problems.
 *   Q [S][n][n] and R [S][nu][nu] are the Δt-scaled block-diagonal weights of lqr.jl:18-19 (the caller scales);
 *   Q_bstride/R_bstride 0 shares them.  Ku [S][N-1][nu][n]; kexit [S] = value of the loop variable k at exit
 *   (1 if the loop completed; lqr.jl:149-150,179); iters [S] executed iterations.
 * ------------------------------------------------------------------------------------------------ */
int ccb_dlqr(int32_t device, int32_t S, int32_t n, int32_t nu, int32_t ml, int32_t N, const double* A, const double* Bu, const double* Bl,
             const double* G, const double* Q, int64_t Q_bstride, const double* R, int64_t R_bstride, double* Ku, int32_t* kexit,
             int32_t* iters);

/* LQR(mechanism, bodyids, eqcids, Q, R, horizon; xd, vd, qd, ωd, Fτd) gain synthesis (lqr.jl:49-66 + :17-47) =
 * ccb_linearize + ccb_dlqr without leaving the device.  N = Ntemp of lqr.jl:21-27. */
int ccb_lqr_synthesize(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu,
                       const double* Q, int64_t Q_bstride, const double* R, int64_t R_bstride, int32_t N, const ccb_sim_opts* opts,
                       double* Ku, int32_t* kexit, int32_t* iters);

/* The same constructor with the gain table kept on the device: synthesis + ccb_ctrl_create in one call, nothing but the
 * iteration counters comes back to the host.  infinite != 0: N is Ntemp = ceil(10/Δt), only Ku[1] is kept (lqr.jl:26,42) and
 * converged[s] (may be NULL) = (Ku[1] == Ku[2]) (lqr.jl:41); infinite == 0: control_lqr! for k < N (lqr.jl:106).
 * The controller is laid out for S systems (S == 1: shared by every system of any batch). */
int ccb_lqr_create(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu, const double* Q,
                   int64_t Q_bstride, const double* R, int64_t R_bstride, int32_t N, int32_t infinite, const ccb_sim_opts* opts, ccb_ctrl** out,
                   int32_t* kexit, int32_t* iters, int32_t* converged);
/* copy the gain table of an LQR / TrackingLQR controller to the host: [B or 1][rows][nu][n]; count = number of doubles expected */
int ccb_ctrl_get_gains(ccb_ctrl* c, int32_t n, double* K, int64_t count);

/* dlqr(mechanism, xd, vd, qd, ωd, Fτd, eqcids, Q, R, N)  (reference src/control/lqr_tracking.jl:73-122) for S
 * reference trajectories: traj [S][N][nb][13], Ftau [S][N][nu]; Ku [S][N-1][nu][n]. */
int ccb_tracking_synthesize(ccb_mech* m, int32_t device, int32_t S, int32_t N, const double* traj, const double* Ftau, const int32_t* eqc,
                            int32_t nu, const double* Q, const double* R, const ccb_sim_opts* opts, double* Ku, int32_t* kexit,
                            int32_t* iters);

/* TrackingLQR(mechanism, storage, Fτ, eqcids, Q, R) (lqr_tracking.jl:17-43) with the gain tables, the desired trajectories and the
 * feed-forward kept on the device: synthesis + ccb_ctrl_create(CCB_CTRL_TRACKING) in one call.  The controller is laid out for S
 * systems (S == 1: shared by every system of any batch); read the gains back with ccb_ctrl_get_gains. */
int ccb_tracking_create(ccb_mech* m, int32_t device, int32_t S, int32_t N, const double* traj, const double* Ftau, const int32_t* eqc, int32_t nu,
                        const double* Q, const double* R, const ccb_sim_opts* opts, ccb_ctrl** out, int32_t* kexit, int32_t* iters);

/* ------------------------------------------------------------------------------------------------
 * Algebraic Riccati utilities of the reference (src/util/util.jl) for S independent problems, row-major:
 *   A [S][n][n], B [S][n][m], Q [S][n][n], R [S][m][m]  (1 <= m <= n; Q, R symmetric, R positive definite).
 *   kind CCB_ARE_DISCRETE:    X = dare(A,B,Q,R)  (util.jl:1-20),   K = dlqr(A,B,Q,R) = (R + B'XB) \ B'XA  (util.jl:45-49)
 *   kind CCB_ARE_CONTINUOUS:  X = care(A,B,Q,R)  (util.jl:22-37),  K = lqr(A,B,Q,R)  = R \ B'X           (util.jl:39-43)
 * X [S][n][n] and K [S][m][n]: either may be NULL.  dlqr(A,B,Q,R,Δt) (util.jl:51-59) is the discrete call on
 * (A Δt + I, B Δt, Q Δt, R Δt): the caller scales.  iters [S] (may be NULL) = doubling iterations; status [S] (may be
 * NULL): 0 converged, 1 not converged (no stabilising solution within 64 doublings), 2 numerical breakdown.
 * The device computes the stabilising solution with the structure-preserving doubling algorithm instead of the ordered
 * (generalised) Schur form of util.jl:16-17,28-29: same matrix, no data-dependent eigenvalue reordering.
 * ------------------------------------------------------------------------------------------------ */
#define CCB_ARE_DISCRETE 0
#define CCB_ARE_CONTINUOUS 1
int ccb_are_solve(int32_t device, int32_t kind, int32_t S, int32_t n, int32_t m, const double* A, const double* B, const double* Q,
                  const double* R, double* X, double* K, int32_t* iters, int32_t* status);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: LQR syntheses sharded over the GPUs of a node, every GPU needing every gain table.  Instead of an
 * all-gather after the fact, the Riccati sweep stores each gain row into every GPU's table while it runs (plain stores into peer
 * memory over NVLink): a table is allocated on each rank (ccb_gains_alloc, which also returns its CUDA IPC handle), the 64-byte
 * handles are exchanged by the host (torch.distributed / MPI / a Julia Channel), opened on the other ranks (ccb_gains_open), and
 * ccb_lqr_synthesize_peers = ccb_lqr_synthesize with the gains of its S systems written at system index s_offset + s of every
 * table in `tables` (1..8; the rank's own and the opened peers').  Tables are [S_total][N-1][nu][n].  The caller synchronises the
 * ranks (a barrier) before anybody reads a table.  Shapes the tensor-core sweep K4 does not take are synthesised locally and copied
 * to the tables.  ccb_tracking_synthesize_peers is the same for ccb_tracking_synthesize (the SURVEY 8(e) case: 64 TrackingLQR tables of
 * 9.4 MB each that all 8 GPUs need).  ccb_lqr_ctrl_from_table builds a finite-horizon LQR controller (lqr.jl:89-114) on such a table (borrowed).
 * ------------------------------------------------------------------------------------------------ */
int ccb_gains_alloc(int32_t device, int64_t count, double** dptr, unsigned char ipc_handle[64]);
int ccb_gains_open(int32_t device, const unsigned char ipc_handle[64], double** dptr);
int ccb_gains_close(double* dptr);
int ccb_gains_free(double* dptr);
int ccb_gains_read(int32_t device, const double* dptr, int64_t offset, double* host, int64_t count);
/* Deferred back-fill of the exchange (optional).  dlqr fills the rows below its break index with copies of the last row it computed
 * (reference src/control/lqr.jl:179-181: `for k2 = k-1:-1:1  Ku[k2] = Ku[k2+1]`); for the Sawyer arm at N = 2000 that is a fifth of every
 * table, and there is no reason to send copies over NVLink.  After ccb_gains_defer_backfill(1) the ccb_lqr_synthesize_peers calls of
 * this thread push only the rows the recursion produced (k >= kexit) when the sweep itself does the stores; once the exchange barrier has
 * passed and the break indices of ALL systems are known on every rank, each rank completes its own table with ccb_gains_backfill
 * (kexit: host, [S], values 1 .. N-1 as returned by the synthesis; idempotent on a table that is already complete). */
int ccb_gains_defer_backfill(int32_t on);
int ccb_gains_backfill(int32_t device, double* table, int64_t S, int32_t N, int32_t nu, int32_t n, const int32_t* kexit);
int ccb_lqr_synthesize_peers(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu,
                             const double* Q, int64_t Q_bstride, const double* R, int64_t R_bstride, int32_t N, const ccb_sim_opts* opts,
                             int32_t ntables, double* const* tables, int64_t s_offset, int32_t* kexit, int32_t* iters);
int ccb_tracking_synthesize_peers(ccb_mech* m, int32_t device, int32_t S, int32_t N, const double* traj, const double* Ftau, const int32_t* eqc,
                                  int32_t nu, const double* Q, const double* R, const ccb_sim_opts* opts, int32_t ntables, double* const* tables,
                                  int64_t s_offset, int32_t* kexit, int32_t* iters);
int ccb_lqr_ctrl_from_table(ccb_mech* m, int32_t device, int32_t S, int32_t N, int32_t nu, const int32_t* eqc, const double* K_dev,
                            const double* zd, const double* Fd, ccb_ctrl** out);
/* the same for control_trackinglqr! (lqr_tracking.jl:46-71): traj [S][N][nb][13], Ftau [S][N][nu] are copied, the gain table is borrowed */
int ccb_tracking_ctrl_from_table(ccb_mech* m, int32_t device, int32_t S, int32_t N, int32_t nu, const int32_t* eqc, const double* K_dev,
                                 const double* traj, const double* Ftau, ccb_ctrl** out);

/* The synthesis calls keep their device scratch (up to a quarter of the device memory; CCB_ARENA_MAX_GB overrides) between calls so
 * that repeated syntheses do not pay cudaMalloc / cudaFree again; this gives it back.  No reference counterpart (Julia's GC). */
int ccb_scratch_release(void);

/* Device time (ms, CUDA events on the launching stream) of the phases of the last ccb_dlqr / ccb_lqr_synthesize /
 * ccb_tracking_synthesize call on the calling thread: out[0] linearisation (K1 knot solve + K2), out[1] constraint projection,
 * out[2] Riccati sweep (K3 / K4), out[3] elapsed device time of the whole call when the phases overlap (the TrackingLQR
 * pipeline: out[0..2] are then per-phase sums of overlapping work), else 0.  Measurement aid for bench.py (no reference counterpart). */
int ccb_synth_phase_ms(double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* CCB_H */
