// csrc/ccb_api.cu -- the extern "C" surface declared in include/ccb.h (handles, copies, launches).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include "ccb_internal.h"

namespace ccb {
const char* last_error_cstr();
}
using namespace ccb;

namespace ccb {
// Uploads and memsets of the create functions run on the legacy default stream; batch streams are non-blocking and therefore NOT ordered
// against it (and a cudaMemcpy from pageable memory may return before its DMA has landed).  Every create function ends with this, so a
// rollout launched on a batch stream right after it sees the finished tables.
int finish_uploads() {
    CCB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    return CCB_OK;
}
}  // namespace ccb

extern "C" {

const char* ccb_last_error(void) { return last_error_cstr(); }
int ccb_version(void) { return 100; }
int ccb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int64_t ccb_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------------------ mechanism
int ccb_mech_create(const ccb_mech_desc* desc, ccb_mech** out) {
    if (!out) { set_error("ccb_mech_create: out is NULL"); return CCB_E_INVALID; }
    ccb_mech* m = new (std::nothrow) ccb_mech();
    if (!m) { set_error("out of memory"); return CCB_E_INVALID; }
    int rc = build_host_mech(desc, &m->h);
    if (rc != CCB_OK) { delete m; return rc; }
    *out = m;
    return CCB_OK;
}
int ccb_mech_destroy(ccb_mech* m) {
    if (!m) return CCB_OK;
    free_dev_mech(&m->h);
    delete m;
    return CCB_OK;
}
int ccb_mech_dims(const ccb_mech* m, int32_t out[8]) {
    if (!m || !out) { set_error("ccb_mech_dims: NULL"); return CCB_E_INVALID; }
    out[0] = m->h.nb; out[1] = m->h.ne; out[2] = m->h.nj; out[3] = m->h.ml; out[4] = m->h.nnodes; out[5] = (int)m->h.loop_edges.size() / 2;
    out[6] = 12 * m->h.nb; out[7] = m->h.lanes;
    return CCB_OK;
}
static void copy_out(const std::vector<int>& v, int32_t* o) { if (o && !v.empty()) std::memcpy(o, v.data(), v.size() * sizeof(int)); }
int ccb_mech_graph(const ccb_mech* m, int32_t* eqc_id, int32_t* body_id, int32_t* dfs_list, int32_t* parents, int32_t* loop_edges,
                   int32_t* eqc_dim, int32_t* eqc_row0, int32_t* act_kind, int32_t* joint_elim) {
    if (!m) { set_error("ccb_mech_graph: NULL"); return CCB_E_INVALID; }
    copy_out(m->h.eqc_id, eqc_id); copy_out(m->h.body_id, body_id); copy_out(m->h.dfs_list, dfs_list); copy_out(m->h.parents, parents);
    copy_out(m->h.loop_edges, loop_edges); copy_out(m->h.eqc_dim, eqc_dim); copy_out(m->h.eqc_row0, eqc_row0); copy_out(m->h.act_kind, act_kind);
    copy_out(m->h.elim, joint_elim);
    return CCB_OK;
}

// ------------------------------------------------------------------------------------------------ batch
int ccb_batch_create(ccb_mech* m, int32_t B, int32_t device, ccb_batch** out) {
    if (!m || !out || B < 1) { set_error("ccb_batch_create: bad arguments"); return CCB_E_INVALID; }
    if (ccb_device_count() <= device || device < 0) { set_error("ccb_batch_create: no such CUDA device (this library has no CPU fallback)"); return CCB_E_CUDA; }
    const DevMech* view;
    int rc = get_dev_mech(&m->h, device, &view);
    if (rc != CCB_OK) return rc;
    CCB_CUDA(cudaSetDevice(device));
    ccb_batch* b = new ccb_batch();
    b->mech = m; b->B = B; b->device = device;
    const size_t nl = (size_t)B * (m->h.ml > 0 ? m->h.ml : 1) * sizeof(double);
    cudaError_t e = cudaMalloc(&b->state, (size_t)B * m->h.nb * 13 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->lambda, nl);
    if (e == cudaSuccess) e = cudaMalloc(&b->status_dev, (size_t)B * sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->lambda, 0, nl, b->stream);     // on the consuming stream
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    if (e != cudaSuccess) { ccb_batch_destroy(b); return cuda_fail(e, "ccb_batch_create"); }
    *out = b;
    return CCB_OK;
}
int ccb_batch_destroy(ccb_batch* b) {
    if (!b) return CCB_OK;
    cudaSetDevice(b->device);
    cudaFree(b->state); cudaFree(b->lambda); cudaFree(b->status_dev); cudaFree(b->storage_dev); cudaFree(b->persist_buf);
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    for (int i = 0; i < 2; i++) { if (b->ev_kernel[i]) cudaEventDestroy(b->ev_kernel[i]); if (b->ev_copy[i]) cudaEventDestroy(b->ev_copy[i]); }
    delete b;
    return CCB_OK;
}
int ccb_batch_set_state(ccb_batch* b, const double* state, const double* lambda) {
    if (!b || !state) { set_error("ccb_batch_set_state: NULL"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(b->device));
    const size_t ns = (size_t)b->B * b->mech->h.nb * 13 * sizeof(double), nl = (size_t)b->B * b->mech->h.ml * sizeof(double);
    CCB_CUDA(cudaMemcpyAsync(b->state, state, ns, cudaMemcpyHostToDevice, b->stream));
    if (nl) {
        if (lambda) CCB_CUDA(cudaMemcpyAsync(b->lambda, lambda, nl, cudaMemcpyHostToDevice, b->stream));
        else CCB_CUDA(cudaMemsetAsync(b->lambda, 0, nl, b->stream));
    }
    CCB_CUDA(cudaStreamSynchronize(b->stream));
    return CCB_OK;
}
int ccb_batch_get_state(ccb_batch* b, double* state, double* lambda) {
    if (!b) { set_error("ccb_batch_get_state: NULL"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(b->device));
    const size_t ns = (size_t)b->B * b->mech->h.nb * 13 * sizeof(double), nl = (size_t)b->B * b->mech->h.ml * sizeof(double);
    if (state) CCB_CUDA(cudaMemcpyAsync(state, b->state, ns, cudaMemcpyDeviceToHost, b->stream));
    if (lambda && nl) CCB_CUDA(cudaMemcpyAsync(lambda, b->lambda, nl, cudaMemcpyDeviceToHost, b->stream));
    CCB_CUDA(cudaStreamSynchronize(b->stream));
    return CCB_OK;
}
int ccb_batch_device_ptrs(ccb_batch* b, void** state_dev, void** lambda_dev) {
    if (!b) { set_error("ccb_batch_device_ptrs: NULL"); return CCB_E_INVALID; }
    if (state_dev) *state_dev = b->state;
    if (lambda_dev) *lambda_dev = b->lambda;
    return CCB_OK;
}
int ccb_batch_sync(ccb_batch* b) {
    if (!b) { set_error("ccb_batch_sync: NULL"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(b->device));
    CCB_CUDA(cudaStreamSynchronize(b->stream));
    return CCB_OK;
}
int ccb_batch_stream(ccb_batch* b, void** stream) {
    if (!b || !stream) { set_error("ccb_batch_stream: NULL"); return CCB_E_INVALID; }
    *stream = (void*)b->stream;
    return CCB_OK;
}

// ------------------------------------------------------------------------------------------------ controllers
static int upload(ccb_ctrl* c, const void* host, size_t bytes, const void** dev) {
    *dev = nullptr;
    if (!host || bytes == 0) return CCB_OK;
    void* p = nullptr;
    CCB_CUDA(cudaMalloc(&p, bytes));
    c->allocs.push_back(p);
    CCB_CUDA(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
    *dev = p;
    return CCB_OK;
}

// integrator / last-error state of a PID controller for `B` systems (pid.jl:5-6), zeroed; replaces an earlier allocation
static int pid_state_alloc(ccb_ctrl* c, int B) {
    const size_t bytes = (size_t)B * c->v.nu * sizeof(double);
    for (double** slot : {&c->v.pid_int, &c->v.pid_last}) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(PID state)");
        c->allocs.push_back(q);
        *slot = (double*)q;
        if ((e = cudaMemset(q, 0, bytes)) != cudaSuccess) return cuda_fail(e, "cudaMemset(PID state)");
    }
    return CCB_OK;
}

int ccb_ctrl_create(const ccb_mech* m, const ccb_ctrl_desc* d, int32_t device, ccb_ctrl** out) {
    if (!m || !d || !out) { set_error("ccb_ctrl_create: NULL"); return CCB_E_INVALID; }
    if (ccb_device_count() <= device || device < 0) { set_error("ccb_ctrl_create: no such CUDA device (this library has no CPU fallback)"); return CCB_E_CUDA; }
    const HostMech& H = m->h;
    const int nu = d->nu, n = 12 * H.nb, B = d->B > 0 ? d->B : 1;
    if (d->kind < CCB_CTRL_NONE || d->kind > CCB_CTRL_PID) { set_error("ccb_ctrl_create: unknown kind"); return CCB_E_INVALID; }
    if (d->kind != CCB_CTRL_NONE) {
        if (nu < 1 || !d->eqc) { set_error("ccb_ctrl_create: no controlled eqcs"); return CCB_E_INVALID; }
        for (int i = 0; i < nu; i++) {
            int e = d->eqc[i];
            // "Only 1 DOF joints are supported" (reference src/control/pid.jl:20,36; lqr.jl K is 1 x NK per eqc)
            if (e < 0 || e >= H.ne || H.act_kind[e] == 0) { set_error("ccb_ctrl_create: Only 1 DOF joints are supported"); return CCB_E_INVALID; }
        }
    }
    if ((d->kind == CCB_CTRL_LQR_FIN || d->kind == CCB_CTRL_TRACKING) && d->N < 2) { set_error("ccb_ctrl_create: horizon N must be >= 2"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(device));
    ccb_ctrl* c = new ccb_ctrl();
    c->device = device; c->B = B;
    std::memset(&c->v, 0, sizeof(c->v));
    c->v.kind = d->kind; c->v.nu = nu; c->v.N = d->N;
    auto mult = [&](int64_t bs) { return bs == 0 ? (size_t)1 : (size_t)B; };
    int rc = CCB_OK;
    const void* p;
    auto fail = [&](int r) { ccb_ctrl_destroy(c); return r; };
    if (d->kind != CCB_CTRL_NONE) { if ((rc = upload(c, d->eqc, nu * sizeof(int), &p))) return fail(rc); c->v.eqc = (const int*)p; }
    if (d->kind == CCB_CTRL_LQR_INF || d->kind == CCB_CTRL_LQR_FIN || d->kind == CCB_CTRL_TRACKING) {
        if (!d->K || !d->zd || !d->Fd) { set_error("ccb_ctrl_create: K, zd and Fd are required"); return fail(CCB_E_INVALID); }
        const size_t rows = d->kind == CCB_CTRL_LQR_INF ? 1 : (size_t)d->N - 1;
        const size_t nz = d->kind == CCB_CTRL_TRACKING ? (size_t)d->N : 1;
        if ((rc = upload(c, d->K, mult(d->K_bstride) * rows * nu * n * sizeof(double), &p))) return fail(rc);
        c->v.K = (const double*)p; c->v.K_bs = d->K_bstride;
        if ((rc = upload(c, d->zd, mult(d->zd_bstride) * nz * H.nb * 13 * sizeof(double), &p))) return fail(rc);
        c->v.zd = (const double*)p; c->v.zd_bs = d->zd_bstride;
        if ((rc = upload(c, d->Fd, mult(d->Fd_bstride) * nz * nu * sizeof(double), &p))) return fail(rc);
        c->v.Fd = (const double*)p; c->v.Fd_bs = d->Fd_bstride;
    } else if (d->kind == CCB_CTRL_TABLE) {
        if (!d->U || d->N < 1) { set_error("ccb_ctrl_create: U[N][nu] is required"); return fail(CCB_E_INVALID); }
        if ((rc = upload(c, d->U, mult(d->U_bstride) * (size_t)d->N * nu * sizeof(double), &p))) return fail(rc);
        c->v.U = (const double*)p; c->v.U_bs = d->U_bstride;
    } else if (d->kind == CCB_CTRL_PID) {
        if (!d->P || !d->I || !d->D || !d->goal) { set_error("ccb_ctrl_create: P, I, D, goal are required"); return fail(CCB_E_INVALID); }
        if ((rc = upload(c, d->P, nu * sizeof(double), &p))) return fail(rc); c->v.P = (const double*)p;
        if ((rc = upload(c, d->I, nu * sizeof(double), &p))) return fail(rc); c->v.I = (const double*)p;
        if ((rc = upload(c, d->D, nu * sizeof(double), &p))) return fail(rc); c->v.D = (const double*)p;
        if ((rc = upload(c, d->goal, mult(d->goal_bstride) * nu * sizeof(double), &p))) return fail(rc);
        c->v.goal = (const double*)p; c->v.goal_bs = d->goal_bstride;
        if ((rc = pid_state_alloc(c, B))) return fail(rc);
    }
    if ((rc = finish_uploads())) return fail(rc);
    *out = c;
    return CCB_OK;
}
int ccb_ctrl_destroy(ccb_ctrl* c) {
    if (!c) return CCB_OK;
    cudaSetDevice(c->device);
    for (void* p : c->allocs) cudaFree(p);
    delete c;
    return CCB_OK;
}
int ccb_ctrl_set_replication(ccb_ctrl* c, int32_t per_table) {
    if (!c || per_table < 1) { set_error("ccb_ctrl_set_replication: bad arguments"); return CCB_E_INVALID; }
    if (c->v.kind != CCB_CTRL_LQR_INF && c->v.kind != CCB_CTRL_LQR_FIN && c->v.kind != CCB_CTRL_TRACKING) {
        set_error("ccb_ctrl_set_replication: only LQR / TrackingLQR controllers carry per-system tables"); return CCB_E_INVALID;
    }
    c->v.sys_div = per_table;
    return CCB_OK;
}
int ccb_ctrl_pid_state(ccb_ctrl* c, double* integrated, double* last, int32_t set) {
    if (!c || c->v.kind != CCB_CTRL_PID) { set_error("ccb_ctrl_pid_state: not a PID controller"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->B * c->v.nu * sizeof(double);
    if (set) {
        if (integrated) CCB_CUDA(cudaMemcpy(c->v.pid_int, integrated, bytes, cudaMemcpyHostToDevice));
        if (last) CCB_CUDA(cudaMemcpy(c->v.pid_last, last, bytes, cudaMemcpyHostToDevice));
        return finish_uploads();
    } else {
        if (integrated) CCB_CUDA(cudaMemcpy(integrated, c->v.pid_int, bytes, cudaMemcpyDeviceToHost));
        if (last) CCB_CUDA(cudaMemcpy(last, c->v.pid_last, bytes, cudaMemcpyDeviceToHost));
    }
    return CCB_OK;
}

// ------------------------------------------------------------------------------------------------ simulate!
static int simulate_launch(ccb_batch* b, ccb_ctrl* ctrl, int32_t k0, int32_t steps, const ccb_sim_opts* opts, double* storage_dev, int32_t* status_dev,
                           int status_accum) {
    if (!b || steps < 0) { set_error("ccb_simulate: bad arguments"); return CCB_E_INVALID; }
    if (ctrl && ctrl->device != b->device) { set_error("ccb_simulate: controller lives on another device"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(b->device));
    if (ctrl && (long long)ctrl->B * (ctrl->v.sys_div > 1 ? ctrl->v.sys_div : 1) != b->B) {
        // a controller laid out for another batch size is only usable when NOTHING in it is indexed by the system: every per-system
        // stride zero (one shared table / goal).  PID additionally carries per-system integrator state [B][nu] (pid.jl:5-6): a shared-goal
        // PID built for one system is re-sized to this batch (state reset), anything else is refused.
        const DevCtrl& v = ctrl->v;
        const bool shared = v.K_bs == 0 && v.zd_bs == 0 && v.Fd_bs == 0 && v.U_bs == 0 && v.goal_bs == 0 && v.sys_div <= 1;
        if (!shared || ctrl->B != 1) { set_error("ccb_simulate: controller tables were laid out for another batch size"); return CCB_E_INVALID; }
        if (v.kind == CCB_CTRL_PID) {
            int rc0 = pid_state_alloc(ctrl, b->B);
            if (rc0 != CCB_OK) return rc0;
            ctrl->B = b->B;
            if ((rc0 = finish_uploads()) != CCB_OK) return rc0;
        }
    }
    const DevMech* view;
    int rc = get_dev_mech(&b->mech->h, b->device, &view);
    if (rc != CCB_OK) return rc;
    DevCtrl C;
    std::memset(&C, 0, sizeof(C));
    if (ctrl) C = ctrl->v;
    SimArgs a;
    a.state = b->state; a.lambda = b->lambda; a.storage = storage_dev; a.status = status_dev;
    a.B = b->B; a.k0 = k0; a.steps = steps; a.knot_input = 0; a.s_out = nullptr; a.spw = opts ? opts->reserved : 0;
    a.status_accum = status_accum;
    a.persist = nullptr; a.nchunks = 1; a.chunk_steps = steps;
    if (steps >= 256) {       // candidate for the persistent time-sliced launches of K1r / K1m (the launchers decide)
        const size_t need = (size_t)(b->B + 2) * sizeof(int);
        if (b->persist_cap < need) {
            cudaFree(b->persist_buf); b->persist_buf = nullptr; b->persist_cap = 0;
            if (cudaMalloc(&b->persist_buf, need) == cudaSuccess) b->persist_cap = need; else cudaGetLastError();
        }
        a.persist = b->persist_buf;
    }
    a.eps = opts && opts->eps > 0 ? opts->eps : 1e-10;
    a.newton_iter = opts && opts->newton_iter > 0 ? opts->newton_iter : 100;
    a.line_iter = opts && opts->line_iter > 0 ? opts->line_iter : 10;
    return launch_rollout(*view, C, a, opts ? opts->lanes : 0, b->mech->h.lanes, b->stream);
}

int ccb_simulate_device(ccb_batch* b, ccb_ctrl* ctrl, int32_t k0, int32_t steps, const ccb_sim_opts* opts, double* storage_dev,
                        int32_t* status_dev) {
    return simulate_launch(b, ctrl, k0, steps, opts, storage_dev, status_dev, 0);
}

// record = true with a HOST trajectory buffer [steps][B][nb][13] (104 nb bytes per system-step: 1.7 GB for 4096 cartpoles x 2000 steps, as
// long on PCIe as the rollout itself).  The rollout is cut into chunks of steps (the kernels resume bit-exactly from the stored state,
// multipliers, PID integrators and step index k0): chunk c computes into one half of a device double buffer while the copy stream drains
// chunk c-1 from the other half -- one cudaMemcpyAsync per chunk, ordered by events.  The copy of a chunk is issued AFTER the next chunk's
// launch, so the overlap also holds for pageable host memory (where the copy call blocks the host).  Device memory: 2 chunks, not the
// whole trajectory.
static size_t record_chunk_bytes() {
    if (const char* e = getenv("CCB_RECORD_CHUNK_MB")) return (size_t)std::max(1, atoi(e)) << 20;
    return (size_t)64 << 20;
}
static int simulate_record_pipelined(ccb_batch* b, ccb_ctrl* ctrl, int32_t k0, int32_t steps, const ccb_sim_opts* opts, double* storage, size_t per_step) {
    const size_t step_bytes = per_step * sizeof(double);
    const size_t target = record_chunk_bytes();
    int chunk = (int)std::min<size_t>((size_t)steps, std::max<size_t>(8, target / step_bytes));
    const int nchunks = (steps + chunk - 1) / chunk;
    const size_t need = 2 * (size_t)chunk * step_bytes;
    if (b->storage_cap < need) {
        cudaFree(b->storage_dev); b->storage_dev = nullptr; b->storage_cap = 0;
        cudaError_t e = cudaMalloc(&b->storage_dev, need);
        if (e != cudaSuccess) { b->storage_dev = nullptr; return cuda_fail(e, "cudaMalloc(storage chunks)"); }
        b->storage_cap = need;
    }
    if (!b->copy_stream) {
        CCB_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CCB_CUDA(cudaEventCreateWithFlags(&b->ev_kernel[i], cudaEventDisableTiming));
            CCB_CUDA(cudaEventCreateWithFlags(&b->ev_copy[i], cudaEventDisableTiming));
        }
    }
    auto drain = [&](int c) -> int {            // D2H of chunk c behind its kernel
        const int buf = c & 1, s0 = c * chunk, sc = std::min(chunk, steps - s0);
        CCB_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_kernel[buf], 0));
        CCB_CUDA(cudaMemcpyAsync(storage + (size_t)s0 * per_step, b->storage_dev + (size_t)buf * chunk * per_step, (size_t)sc * step_bytes, cudaMemcpyDeviceToHost,
                                 b->copy_stream));
        CCB_CUDA(cudaEventRecord(b->ev_copy[buf], b->copy_stream));
        return CCB_OK;
    };
    int rc = CCB_OK;
    for (int c = 0; c < nchunks && rc == CCB_OK; c++) {
        const int buf = c & 1, s0 = c * chunk, sc = std::min(chunk, steps - s0);
        if (c >= 2) CCB_CUDA(cudaStreamWaitEvent(b->stream, b->ev_copy[buf], 0));      // the half is free again
        rc = simulate_launch(b, ctrl, k0 + s0, sc, opts, b->storage_dev + (size_t)buf * chunk * per_step, b->status_dev, c > 0);
        if (rc != CCB_OK) break;
        CCB_CUDA(cudaEventRecord(b->ev_kernel[buf], b->stream));
        if (c >= 1) rc = drain(c - 1);
    }
    if (rc == CCB_OK) rc = drain(nchunks - 1);
    cudaError_t e1 = cudaStreamSynchronize(b->copy_stream), e2 = cudaStreamSynchronize(b->stream);
    if (rc != CCB_OK) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "record pipeline (copy stream)");
    if (e2 != cudaSuccess) return cuda_fail(e2, "record pipeline (compute stream)");
    return CCB_OK;
}

int ccb_simulate(ccb_batch* b, ccb_ctrl* ctrl, int32_t k0, int32_t steps, const ccb_sim_opts* opts, double* storage, int32_t* status) {
    if (!b) { set_error("ccb_simulate: NULL batch"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(b->device));
    const size_t per_step = (size_t)b->B * b->mech->h.nb * 13;
    double* sdev = nullptr;
    if (storage && steps > 16 && per_step * steps * sizeof(double) >= 2 * record_chunk_bytes() && !getenv("CCB_RECORD_NO_PIPELINE")) {
        int rc = simulate_record_pipelined(b, ctrl, k0, steps, opts, storage, per_step);
        if (rc != CCB_OK) return rc;
        if (status) CCB_CUDA(cudaMemcpyAsync(status, b->status_dev, (size_t)b->B * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        CCB_CUDA(cudaStreamSynchronize(b->stream));
        return CCB_OK;
    }
    if (storage && steps > 0) {
        const size_t need = per_step * steps * sizeof(double);
        if (b->storage_cap < need) {
            cudaFree(b->storage_dev); b->storage_dev = nullptr; b->storage_cap = 0;
            cudaError_t e = cudaMalloc(&b->storage_dev, need);
            if (e != cudaSuccess) { b->storage_dev = nullptr; return cuda_fail(e, "cudaMalloc(storage)"); }
            b->storage_cap = need;
        }
        sdev = b->storage_dev;
    }
    int rc = ccb_simulate_device(b, ctrl, k0, steps, opts, sdev, b->status_dev);
    if (rc != CCB_OK) return rc;
    if (sdev) CCB_CUDA(cudaMemcpyAsync(storage, sdev, per_step * steps * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (status) CCB_CUDA(cudaMemcpyAsync(status, b->status_dev, (size_t)b->B * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CCB_CUDA(cudaStreamSynchronize(b->stream));
    return CCB_OK;
}

}  // extern "C"
