// csrc/ccb_synth.cu -- C-ABI entry points for linearsystem / dlqr / LQR and TrackingLQR gain synthesis (K2 + K3).
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "ccb_internal.h"

namespace ccb {
int linearize_device(HostMech* H, int device, int S, double* d_zd, const double* d_Fd, const int* d_eqc, int nu, const ccb_sim_opts* opts,
                     double* d_A, double* d_Bu, double* d_Bl, double* d_G, double* d_lam, double* d_s, int* d_status, cudaStream_t stream);
size_t riccati_project_ws_doubles(int n, int nu, int ml);
size_t riccati_sweep_ws_doubles(int n, int nu);
int riccati_project_device(int P, int n, int nu, int ml, const double* A, const double* Bu, const double* Bl, const double* G, double* At, double* D,
                           double* ws, cudaStream_t stream);
int riccati_project_mech_device(const DevMech& M, int P, int nu, const double* A, const double* Bu, const double* Bl, const double* G, double* At,
                                double* D, double* ws, long long ws_stride, cudaStream_t stream);
// projection of P linearisations of mechanism H: the structure-aware kernel (ccb_riccati_project.cu) when it applies, else the dense one
static int project_linearisations(HostMech* H, int device, int P, int nu, const double* A, const double* Bu, const double* Bl, const double* G, double* At,
                                  double* D, double* ws, cudaStream_t stream) {
    const DevMech* view;
    int rc = get_dev_mech(H, device, &view);
    if (rc != CCB_OK) return rc;
    rc = riccati_project_mech_device(*view, P, nu, A, Bu, Bl, G, At, D, ws, (long long)riccati_project_ws_doubles(12 * H->nb, nu, H->ml), stream);
    if (rc == 1) rc = riccati_project_device(P, 12 * H->nb, nu, H->ml, A, Bu, Bl, G, At, D, ws, stream);
    return rc;
}

int riccati_sweep_device(int S, int n, int nu, int N, const double* At, const double* D, long long stride_t_A, long long stride_t_D,
                         long long stride_s_A, long long stride_s_D, const double* Q, long long Q_bs, const double* R, long long R_bs, double* Ku,
                         int* kexit, int* iters, double* ws, cudaStream_t stream, int k_hi = 0, int k_lo = 0, int resume = 0);
bool riccati_dmma_eligible(int n, int nu);
}
using namespace ccb;
namespace ccb { thread_local PeerTables g_peer_tables = {{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, 0, 0, 1, 0}; }
static thread_local int g_defer_backfill = 0;     // ccb_gains_defer_backfill

namespace {

// Device scratch for one API call.  Requests are carved from a per-device arena that persists across calls (cudaMalloc / cudaFree
// synchronise the device and cost more than a cartpole synthesis batch, and ~0.1 s for the tens of GB of a TrackingLQR batch); the
// arena grows after a call that did not fit, up to a quarter of the device memory (CCB_ARENA_MAX_GB overrides), requests beyond
// that fall back to cudaMalloc and are released when the call returns.  ccb_scratch_release() gives the arenas back.  Not thread
// safe across calls on one device, like the handles themselves.
struct Arena { char* base = nullptr; size_t cap = 0, off = 0, want = 0; };
static std::map<int, Arena> g_arenas;
static size_t arena_max() {
    static size_t v = 0;
    if (v) return v;
    if (const char* e = getenv("CCB_ARENA_MAX_GB")) { v = (size_t)std::max(1.0, atof(e)) << 30; return v; }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = (size_t)16 << 30; }
    v = std::max<size_t>((size_t)4 << 30, total_b / 4);
    return v;
}

struct Scratch {
    std::vector<void*> ptrs;
    Arena* ar = nullptr;
    size_t start = 0, asked = 0;
    Scratch() {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) { ar = &g_arenas[dev]; start = ar->off; }
    }
    ~Scratch() {
        for (void* p : ptrs) cudaFree(p);
        if (!ar) return;
        ar->off = start;
        ar->want = std::max(ar->want, asked <= arena_max() ? asked : ar->want);
        if (start == 0 && ar->want > ar->cap) {      // grow for the next call
            cudaFree(ar->base);
            ar->base = nullptr; ar->cap = 0;
            void* p = nullptr;
            const size_t cap = ar->want + ar->want / 4;
            if (cudaMalloc(&p, cap) == cudaSuccess) { ar->base = (char*)p; ar->cap = cap; } else cudaGetLastError();
        }
    }
    template <class T> int get(T** out, size_t count) {
        const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
        asked += bytes;
        if (ar && ar->base && ar->off + bytes <= ar->cap) {
            *out = (T*)(ar->base + ar->off);
            ar->off += bytes;
            return CCB_OK;
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
        ptrs.push_back(p);
        *out = (T*)p;
        return CCB_OK;
    }
    template <class T> int put(T** out, const T* host, size_t count) {
        int rc = get(out, count);
        if (rc != CCB_OK) return rc;
        if (count) CCB_CUDA(cudaMemcpy(*out, host, count * sizeof(T), cudaMemcpyHostToDevice));
        return CCB_OK;
    }
};
#define CCB_TRY(x) do { int _rc = (x); if (_rc != CCB_OK) return _rc; } while (0)

// device times of the phases of the last synthesis call on this thread (CUDA events on the launching stream)
thread_local double g_phase_ms[4] = {0, 0, 0, 0};   // linearize (K1 knot solve + K2), project, sweep (K3/K4), elapsed total when the phases overlap (else 0)
struct PhaseTimer {
    cudaEvent_t ev[4];
    PhaseTimer() { for (auto& e : ev) cudaEventCreate(&e); }
    ~PhaseTimer() { for (auto& e : ev) cudaEventDestroy(e); }
    void mark(int i) { cudaEventRecord(ev[i], 0); }
    void finish(bool accumulate) {   // call after a device synchronisation
        for (int i = 0; i < 3; i++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
            g_phase_ms[i] = (accumulate ? g_phase_ms[i] : 0.0) + ms;
        }
        if (!accumulate) g_phase_ms[3] = 0.0;       // only the overlapped tracking pipeline reports a separate total
    }
};

// every linearisation rests on one Newton solve of its knot (lambda_d, v2, w2): a knot whose solve did not converge would silently turn
// into garbage A, B, G and gains, so the synthesis entry points fail instead (status <= 0 marks a failed solve, ccb_simulate's convention)
__global__ void count_failed_kernel(const int* __restrict__ status, int count, int* __restrict__ out) {
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) bad += status[i] <= 0;
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(out, bad);
}
int check_knot_status(const int* d_status, int count, int* d_counter, const char* who) {
    if (cudaMemset(d_counter, 0, sizeof(int)) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaMemset");
    count_failed_kernel<<<std::min(148, (count + 255) / 256), 256>>>(d_status, count, d_counter);
    g_launches++;
    int bad = 0;
    cudaError_t e = cudaMemcpy(&bad, d_counter, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(knot status)");
    if (bad) {
        set_error(std::string(who) + ": linearsystem: the Newton solve of " + std::to_string(bad) + " of " + std::to_string(count) +
                  " setpoints / trajectory knots did not converge (not an equilibrium within newton_iter iterations?)");
        return CCB_E_INVALID;
    }
    return CCB_OK;
}

int check_device(int device, const char* who) {
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || device < 0 || device >= nd) {
        cudaGetLastError();
        set_error(std::string(who) + ": no such CUDA device (this library has no CPU fallback)");
        return CCB_E_CUDA;
    }
    CCB_CUDA(cudaSetDevice(device));
    return CCB_OK;
}

int check_eqcs(const HostMech& H, const int32_t* eqc, int nu, const char* who) {
    for (int i = 0; i < nu; i++)
        if (eqc[i] < 0 || eqc[i] >= H.ne || H.act_kind[eqc[i]] == 0) { set_error(std::string(who) + ": Only 1 DOF joints are supported"); return CCB_E_INVALID; }
    return CCB_OK;
}

}  // namespace

extern "C" {

int ccb_linearize(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu,
                  const ccb_sim_opts* opts, double* A, double* Bu, double* Bl, double* G, double* lambda_d) {
    if (!m || S < 1 || !zd || nu < 0 || (nu > 0 && (!Fd || !eqc))) { set_error("ccb_linearize: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_linearize"));
    CCB_TRY(check_device(device, "ccb_linearize"));
    const int nb = m->h.nb, ml = m->h.ml, n = 12 * nb;
    Scratch sc;
    double *d_zd, *d_Fd, *d_A, *d_Bu, *d_Bl, *d_G, *d_lam, *d_s;
    int *d_eqc, *d_status;
    CCB_TRY(sc.put(&d_zd, zd, (size_t)S * nb * 13));
    CCB_TRY(sc.put(&d_Fd, Fd, (size_t)S * nu));
    CCB_TRY(sc.put(&d_eqc, (const int*)eqc, (size_t)nu));
    CCB_TRY(sc.get(&d_A, (size_t)S * n * n)); CCB_TRY(sc.get(&d_Bu, (size_t)S * n * nu)); CCB_TRY(sc.get(&d_Bl, (size_t)S * n * ml));
    CCB_TRY(sc.get(&d_G, (size_t)S * ml * n)); CCB_TRY(sc.get(&d_lam, (size_t)S * ml)); CCB_TRY(sc.get(&d_s, (size_t)S * nb * 6));
    CCB_TRY(sc.get(&d_status, (size_t)S)); int* d_cnt; CCB_TRY(sc.get(&d_cnt, (size_t)1));
    CCB_CUDA(cudaMemset(d_lam, 0, std::max<size_t>((size_t)S * ml, 1) * sizeof(double)));
    CCB_TRY(linearize_device(&m->h, device, S, d_zd, d_Fd, d_eqc, nu, opts, d_A, d_Bu, d_Bl, d_G, d_lam, d_s, d_status, 0));
    CCB_CUDA(cudaDeviceSynchronize());
    CCB_TRY(check_knot_status(d_status, S, d_cnt, "ccb_linearize"));
    if (A) CCB_CUDA(cudaMemcpy(A, d_A, (size_t)S * n * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (Bu && nu) CCB_CUDA(cudaMemcpy(Bu, d_Bu, (size_t)S * n * nu * sizeof(double), cudaMemcpyDeviceToHost));
    if (Bl && ml) CCB_CUDA(cudaMemcpy(Bl, d_Bl, (size_t)S * n * ml * sizeof(double), cudaMemcpyDeviceToHost));
    if (G && ml) CCB_CUDA(cudaMemcpy(G, d_G, (size_t)S * ml * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (lambda_d && ml) CCB_CUDA(cudaMemcpy(lambda_d, d_lam, (size_t)S * ml * sizeof(double), cudaMemcpyDeviceToHost));
    return CCB_OK;
}

int ccb_dlqr(int32_t device, int32_t S, int32_t n, int32_t nu, int32_t ml, int32_t N, const double* A, const double* Bu, const double* Bl,
             const double* G, const double* Q, int64_t Q_bs, const double* R, int64_t R_bs, double* Ku, int32_t* kexit, int32_t* iters) {
    if (S < 1 || n < 1 || nu < 1 || ml < 0 || N < 2 || !A || !Bu || !Q || !R || !Ku) { set_error("ccb_dlqr: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_device(device, "ccb_dlqr"));
    Scratch sc;
    double *d_A, *d_Bu, *d_Bl, *d_G, *d_Q, *d_R, *d_At, *d_D, *d_Ku, *d_ws1, *d_ws2;
    int *d_kexit, *d_iters;
    CCB_TRY(sc.put(&d_A, A, (size_t)S * n * n)); CCB_TRY(sc.put(&d_Bu, Bu, (size_t)S * n * nu));
    CCB_TRY(sc.put(&d_Bl, Bl, (size_t)S * n * ml)); CCB_TRY(sc.put(&d_G, G, (size_t)S * ml * n));
    CCB_TRY(sc.put(&d_Q, Q, (Q_bs ? (size_t)S : 1) * n * n)); CCB_TRY(sc.put(&d_R, R, (R_bs ? (size_t)S : 1) * nu * nu));
    CCB_TRY(sc.get(&d_At, (size_t)S * n * n)); CCB_TRY(sc.get(&d_D, (size_t)S * n * nu));
    CCB_TRY(sc.get(&d_Ku, (size_t)S * (N - 1) * nu * n));
    CCB_TRY(sc.get(&d_ws1, (size_t)S * riccati_project_ws_doubles(n, nu, ml)));
    CCB_TRY(sc.get(&d_ws2, (size_t)S * riccati_sweep_ws_doubles(n, nu)));
    CCB_TRY(sc.get(&d_kexit, (size_t)S)); CCB_TRY(sc.get(&d_iters, (size_t)S));
    PhaseTimer pt;
    pt.mark(0); pt.mark(1);
    CCB_TRY(riccati_project_device(S, n, nu, ml, d_A, d_Bu, d_Bl, d_G, d_At, d_D, d_ws1, 0));
    pt.mark(2);
    CCB_TRY(riccati_sweep_device(S, n, nu, N, d_At, d_D, 0, 0, (long long)n * n, (long long)n * nu, d_Q, Q_bs, d_R, R_bs, d_Ku, d_kexit, d_iters, d_ws2, 0));
    pt.mark(3);
    CCB_CUDA(cudaDeviceSynchronize());
    pt.finish(false);
    CCB_CUDA(cudaMemcpy(Ku, d_Ku, (size_t)S * (N - 1) * nu * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (kexit) CCB_CUDA(cudaMemcpy(kexit, d_kexit, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost));
    if (iters) CCB_CUDA(cudaMemcpy(iters, d_iters, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost));
    return CCB_OK;
}

// dare / care / dlqr / lqr of src/util/util.jl:1-49 for S independent problems (K5, ccb_are.cu)
int ccb_are_solve(int32_t device, int32_t kind, int32_t S, int32_t n, int32_t m, const double* A, const double* B, const double* Q, const double* R,
                  double* X, double* K, int32_t* iters, int32_t* status) {
    if ((kind != CCB_ARE_DISCRETE && kind != CCB_ARE_CONTINUOUS) || S < 1 || n < 1 || m < 1 || m > n || !A || !B || !Q || !R || (!X && !K)) {
        set_error("ccb_are_solve: bad arguments (kind, S >= 1, 1 <= m <= n, A, B, Q, R and at least one of X, K)");
        return CCB_E_INVALID;
    }
    CCB_TRY(check_device(device, "ccb_are_solve"));
    Scratch sc;
    double *d_A, *d_B, *d_Q, *d_R, *d_X = nullptr, *d_K = nullptr, *d_ws;
    int *d_it, *d_st;
    const size_t nn = (size_t)n * n;
    CCB_TRY(sc.put(&d_A, A, S * nn)); CCB_TRY(sc.put(&d_B, B, (size_t)S * n * m)); CCB_TRY(sc.put(&d_Q, Q, S * nn)); CCB_TRY(sc.put(&d_R, R, (size_t)S * m * m));
    if (X) CCB_TRY(sc.get(&d_X, S * nn));
    if (K) CCB_TRY(sc.get(&d_K, (size_t)S * m * n));
    CCB_TRY(sc.get(&d_it, (size_t)S)); CCB_TRY(sc.get(&d_st, (size_t)S));
    CCB_TRY(sc.get(&d_ws, are_ws_doubles(n) * sizeof(double) <= 200 * 1024 ? (size_t)1 : (size_t)S * are_ws_doubles(n)));
    PhaseTimer pt;
    pt.mark(0); pt.mark(1); pt.mark(2);
    CCB_TRY(are_sda_device(kind, S, n, m, d_A, d_B, d_Q, d_R, 1e-15, 64, d_X, d_K, d_it, d_st, d_ws, 0));
    pt.mark(3);
    CCB_CUDA(cudaDeviceSynchronize());
    pt.finish(false);
    if (X) CCB_CUDA(cudaMemcpy(X, d_X, S * nn * sizeof(double), cudaMemcpyDeviceToHost));
    if (K) CCB_CUDA(cudaMemcpy(K, d_K, (size_t)S * m * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (iters) CCB_CUDA(cudaMemcpy(iters, d_it, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost));
    if (status) CCB_CUDA(cudaMemcpy(status, d_st, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost));
    return CCB_OK;
}

int ccb_scratch_release(void) {
    for (auto& kv : g_arenas) {
        if (kv.second.off != 0) { set_error("ccb_scratch_release: a call is in flight"); return CCB_E_INVALID; }
        if (kv.second.base) { cudaSetDevice(kv.first); cudaFree(kv.second.base); }
        kv.second = Arena();
    }
    cudaGetLastError();
    return CCB_OK;
}

int ccb_synth_phase_ms(double out[4]) {
    for (int i = 0; i < 4; i++) out[i] = g_phase_ms[i];
    return CCB_OK;
}

// LQR synthesis on the device: linearsystem (K1 knot solve + K2) -> projection -> Riccati sweep (K3/K4).  The gain table
// d_Ku [S][N-1][nu][n] is allocated here and handed to the caller (cudaFree).
// pipe_n > 0 (ccb_lqr_synthesize_peers, shards of more than one wave of sweep CTAs): the sweep runs in groups of one wave and the finished
// tables of a group are copied to the pipe_n tables (own + peers, slot pipe_off + s) by the copy engines on a second stream WHILE the next
// group sweeps.
static int lqr_synth_device(ccb_mech* m, int device, int S, const double* zd, const double* Fd, const int32_t* eqc, int nu, const double* Q,
                            int64_t Q_bs, const double* R, int64_t R_bs, int N, const ccb_sim_opts* opts, double** d_Ku_out, int32_t* kexit,
                            int32_t* iters, int pipe_n = 0, double* const* pipe_tables = nullptr, long long pipe_off = 0) {
    const int nb = m->h.nb, ml = m->h.ml, n = 12 * nb;
    Scratch sc;
    double *d_zd, *d_Fd, *d_A, *d_Bu, *d_Bl, *d_G, *d_lam, *d_s, *d_Q, *d_R, *d_At, *d_D, *d_ws1, *d_ws2;
    int *d_eqc, *d_status, *d_kexit, *d_iters;
    CCB_TRY(sc.put(&d_zd, zd, (size_t)S * nb * 13)); CCB_TRY(sc.put(&d_Fd, Fd, (size_t)S * nu)); CCB_TRY(sc.put(&d_eqc, (const int*)eqc, (size_t)nu));
    CCB_TRY(sc.put(&d_Q, Q, (Q_bs ? (size_t)S : 1) * n * n)); CCB_TRY(sc.put(&d_R, R, (R_bs ? (size_t)S : 1) * nu * nu));
    CCB_TRY(sc.get(&d_A, (size_t)S * n * n)); CCB_TRY(sc.get(&d_Bu, (size_t)S * n * nu)); CCB_TRY(sc.get(&d_Bl, (size_t)S * n * ml));
    CCB_TRY(sc.get(&d_G, (size_t)S * ml * n)); CCB_TRY(sc.get(&d_lam, (size_t)S * ml)); CCB_TRY(sc.get(&d_s, (size_t)S * nb * 6));
    CCB_TRY(sc.get(&d_status, (size_t)S)); CCB_TRY(sc.get(&d_At, (size_t)S * n * n)); CCB_TRY(sc.get(&d_D, (size_t)S * n * nu));
    CCB_TRY(sc.get(&d_ws1, (size_t)S * riccati_project_ws_doubles(n, nu, ml)));
    CCB_TRY(sc.get(&d_ws2, (size_t)S * riccati_sweep_ws_doubles(n, nu)));
    CCB_TRY(sc.get(&d_kexit, (size_t)S)); CCB_TRY(sc.get(&d_iters, (size_t)S));
    int* d_cnt; CCB_TRY(sc.get(&d_cnt, (size_t)1));
    double* d_Ku = nullptr;
    CCB_CUDA(cudaMalloc(&d_Ku, (size_t)S * (N - 1) * nu * n * sizeof(double)));
    auto bail = [&](int rc) { cudaFree(d_Ku); return rc; };
    if (cudaMemset(d_lam, 0, (size_t)S * ml * sizeof(double)) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaMemset"));
    PhaseTimer pt;
    int rc;
    pt.mark(0);
    if ((rc = linearize_device(&m->h, device, S, d_zd, d_Fd, d_eqc, nu, opts, d_A, d_Bu, d_Bl, d_G, d_lam, d_s, d_status, 0)) != CCB_OK) return bail(rc);
    pt.mark(1);
    if ((rc = project_linearisations(&m->h, device, S, nu, d_A, d_Bu, d_Bl, d_G, d_At, d_D, d_ws1, 0)) != CCB_OK) return bail(rc);
    pt.mark(2);
    if (pipe_n > 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const size_t per = (size_t)(N - 1) * nu * n;
        cudaStream_t ss = nullptr, sc2 = nullptr;
        cudaEvent_t ev = nullptr;
        auto done = [&](int r) { if (ev) cudaEventDestroy(ev); if (ss) cudaStreamDestroy(ss); if (sc2) cudaStreamDestroy(sc2); return r; };
        if (cudaDeviceSynchronize() != cudaSuccess || cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&sc2, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
            return bail(done(cuda_fail(cudaGetLastError(), "peer pipeline set-up")));
        const size_t wsd = riccati_sweep_ws_doubles(n, nu);
        for (int g0 = 0; g0 < S; g0 += sms) {
            const int gs = std::min(sms, S - g0);
            rc = riccati_sweep_device(gs, n, nu, N, d_At + (size_t)g0 * n * n, d_D + (size_t)g0 * n * nu, 0, 0, (long long)n * n, (long long)n * nu,
                                      d_Q + (size_t)g0 * Q_bs, Q_bs, d_R + (size_t)g0 * R_bs, R_bs, d_Ku + (size_t)g0 * per, d_kexit + g0, d_iters + g0,
                                      d_ws2 + (size_t)g0 * wsd, ss);
            if (rc != CCB_OK) { cudaDeviceSynchronize(); return bail(done(rc)); }
            cudaEventRecord(ev, ss);
            cudaStreamWaitEvent(sc2, ev, 0);
            for (int i = 0; i < pipe_n; i++)
                if (cudaMemcpyAsync(pipe_tables[i] + (size_t)(pipe_off + g0) * per, d_Ku + (size_t)g0 * per, (size_t)gs * per * sizeof(double), cudaMemcpyDefault,
                                    sc2) != cudaSuccess) { cudaDeviceSynchronize(); return bail(done(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(peer table)"))); }
        }
        const cudaError_t e1 = cudaStreamSynchronize(ss), e2 = cudaStreamSynchronize(sc2);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return bail(done(cuda_fail(e1 != cudaSuccess ? e1 : e2, "peer pipeline")));
        done(0);
    } else if ((rc = riccati_sweep_device(S, n, nu, N, d_At, d_D, 0, 0, (long long)n * n, (long long)n * nu, d_Q, Q_bs, d_R, R_bs, d_Ku, d_kexit, d_iters, d_ws2,
                                          0)) != CCB_OK) return bail(rc);
    pt.mark(3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaDeviceSynchronize"));
    pt.finish(false);
    if ((rc = check_knot_status(d_status, S, d_cnt, "LQR synthesis")) != CCB_OK) return bail(rc);
    if (kexit && cudaMemcpy(kexit, d_kexit, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaMemcpy"));
    if (iters && cudaMemcpy(iters, d_iters, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaMemcpy"));
    *d_Ku_out = d_Ku;
    return CCB_OK;
}

int ccb_lqr_synthesize(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu,
                       const double* Q, int64_t Q_bs, const double* R, int64_t R_bs, int32_t N, const ccb_sim_opts* opts, double* Ku,
                       int32_t* kexit, int32_t* iters) {
    if (!m || S < 1 || !zd || nu < 1 || !Fd || !eqc || !Q || !R || N < 2 || !Ku) { set_error("ccb_lqr_synthesize: bad arguments"); return CCB_E_INVALID; }
    if (m->h.ml == 0) { set_error("ccb_lqr_synthesize: mechanism without joints (reference lqr.jl:30-37 dead branch) is out of scope"); return CCB_E_UNSUPPORTED; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_lqr_synthesize"));
    CCB_TRY(check_device(device, "ccb_lqr_synthesize"));
    double* d_Ku = nullptr;
    CCB_TRY(lqr_synth_device(m, device, S, zd, Fd, eqc, nu, Q, Q_bs, R, R_bs, N, opts, &d_Ku, kexit, iters));
    cudaError_t e = cudaMemcpy(Ku, d_Ku, (size_t)S * (N - 1) * nu * 12 * m->h.nb * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_Ku);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(Ku)");
    return CCB_OK;
}

// rows r of a [S][rows_all][w] table equal row r+1 ?  (lqr.jl:41 `Ku[1] != Ku[2]`)
__global__ void rows_equal_kernel(const double* K, long long sys_stride, int w, int S, int* out) {
    const int s = blockIdx.x;
    __shared__ int differ;
    if (threadIdx.x == 0) differ = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < w; i += blockDim.x)
        if (K[(size_t)s * sys_stride + i] != K[(size_t)s * sys_stride + w + i]) differ = 1;
    __syncthreads();
    if (threadIdx.x == 0) out[s] = differ ? 0 : 1;
}

int ccb_lqr_create(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu, const double* Q,
                   int64_t Q_bs, const double* R, int64_t R_bs, int32_t N, int32_t infinite, const ccb_sim_opts* opts, ccb_ctrl** out,
                   int32_t* kexit, int32_t* iters, int32_t* converged) {
    if (!m || S < 1 || !zd || nu < 1 || !Fd || !eqc || !Q || !R || N < 2 || !out) { set_error("ccb_lqr_create: bad arguments"); return CCB_E_INVALID; }
    if (m->h.ml == 0) { set_error("ccb_lqr_create: mechanism without joints (reference lqr.jl:30-37 dead branch) is out of scope"); return CCB_E_UNSUPPORTED; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_lqr_create"));
    CCB_TRY(check_device(device, "ccb_lqr_create"));
    const int nb = m->h.nb, n = 12 * nb;
    double* d_Ku = nullptr;
    CCB_TRY(lqr_synth_device(m, device, S, zd, Fd, eqc, nu, Q, Q_bs, R, R_bs, N, opts, &d_Ku, kexit, iters));
    ccb_ctrl* c = new ccb_ctrl();
    c->device = device; c->B = S;
    std::memset(&c->v, 0, sizeof(c->v));
    c->v.nu = nu;
    auto fail = [&](int rc) { cudaFree(d_Ku); ccb_ctrl_destroy(c); return rc; };
    const size_t row = (size_t)nu * n, rows_all = (size_t)N - 1;
    if (converged) {
        if (N > 2) {
            int* d_c = nullptr;
            if (cudaMalloc(&d_c, S * sizeof(int)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
            rows_equal_kernel<<<S, 128>>>(d_Ku, (long long)(rows_all * row), (int)row, S, d_c);
            g_launches++;
            cudaError_t e = cudaMemcpy(converged, d_c, S * sizeof(int), cudaMemcpyDeviceToHost);
            cudaFree(d_c);
            if (e != cudaSuccess) return fail(cuda_fail(e, "rows_equal"));
        } else {
            for (int s = 0; s < S; s++) converged[s] = 1;
        }
    }
    if (infinite) {   // lqr.jl:42: keep only Ku[1]
        double* d_K1 = nullptr;
        if (cudaMalloc(&d_K1, (size_t)S * row * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
        cudaError_t e = cudaMemcpy2D(d_K1, row * sizeof(double), d_Ku, rows_all * row * sizeof(double), row * sizeof(double), S, cudaMemcpyDeviceToDevice);
        cudaFree(d_Ku);
        d_Ku = d_K1;
        if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMemcpy2D"));
        c->v.kind = CCB_CTRL_LQR_INF; c->v.N = 0; c->v.K_bs = S > 1 ? (long long)row : 0;
    } else {
        c->v.kind = CCB_CTRL_LQR_FIN; c->v.N = N; c->v.K_bs = S > 1 ? (long long)(rows_all * row) : 0;
    }
    c->allocs.push_back(d_Ku);
    c->v.K = d_Ku;
    d_Ku = nullptr;
    void *p_eqc = nullptr, *p_zd = nullptr, *p_Fd = nullptr;
    if (cudaMalloc(&p_eqc, nu * sizeof(int)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_eqc);
    if (cudaMalloc(&p_zd, (size_t)S * nb * 13 * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_zd);
    if (cudaMalloc(&p_Fd, (size_t)S * nu * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_Fd);
    if (cudaMemcpy(p_eqc, eqc, nu * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p_zd, zd, (size_t)S * nb * 13 * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p_Fd, Fd, (size_t)S * nu * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaMemcpy"));
    c->v.eqc = (const int*)p_eqc;
    c->v.zd = (const double*)p_zd; c->v.zd_bs = S > 1 ? (long long)nb * 13 : 0;
    c->v.Fd = (const double*)p_Fd; c->v.Fd_bs = S > 1 ? (long long)nu : 0;
    { int rcu = finish_uploads(); if (rcu != CCB_OK) return fail(rcu); }
    *out = c;
    return CCB_OK;
}

// ---- gain tables that peer GPUs of the node can write into (CUDA IPC) and the synthesis that fills them: SURVEY 8(e)'s all-gather of
// gain tables fused into the Riccati sweep (the rows leave over NVLink as plain stores while the recursion runs)
int ccb_gains_alloc(int32_t device, int64_t count, double** dptr, unsigned char ipc_handle[64]) {
    if (count < 1 || !dptr) { set_error("ccb_gains_alloc: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_device(device, "ccb_gains_alloc"));
    void* p = nullptr;
    CCB_CUDA(cudaMalloc(&p, (size_t)count * sizeof(double)));
    if (ipc_handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
        std::memcpy(ipc_handle, &h, 64);
    }
    *dptr = (double*)p;
    return CCB_OK;
}

int ccb_gains_open(int32_t device, const unsigned char ipc_handle[64], double** dptr) {
    if (!ipc_handle || !dptr) { set_error("ccb_gains_open: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_device(device, "ccb_gains_open"));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, 64);
    void* p = nullptr;
    CCB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr = (double*)p;
    return CCB_OK;
}

int ccb_gains_close(double* dptr) {
    if (dptr) CCB_CUDA(cudaIpcCloseMemHandle(dptr));
    return CCB_OK;
}

int ccb_gains_free(double* dptr) {
    if (dptr) CCB_CUDA(cudaFree(dptr));
    return CCB_OK;
}

namespace {
// rows 0 .. kexit[s] - 2 of system s <- row kexit[s] - 1 (lqr.jl:179-181), 16-byte stores; grid (S, row chunks)
__global__ void gains_backfill_kernel(double* __restrict__ table, const int* __restrict__ kexit, long long per, int row) {
    const int s = blockIdx.x, k = kexit[s];
    if (k <= 1) return;
    double* base = table + (size_t)s * per;
    const double2* src = reinterpret_cast<const double2*>(base + (size_t)(k - 1) * row);
    const int row2 = row / 2;
    for (int r = blockIdx.y; r < k - 1; r += gridDim.y) {
        double2* dst = reinterpret_cast<double2*>(base + (size_t)r * row);
        for (int e = threadIdx.x; e < row2; e += blockDim.x) dst[e] = src[e];
    }
}
}  // namespace

int ccb_gains_defer_backfill(int32_t on) {
    g_defer_backfill = on ? 1 : 0;
    return CCB_OK;
}

int ccb_gains_backfill(int32_t device, double* table, int64_t S, int32_t N, int32_t nu, int32_t n, const int32_t* kexit) {
    if (!table || !kexit || S < 0 || N < 2 || nu < 1 || n < 1) { set_error("ccb_gains_backfill: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_device(device, "ccb_gains_backfill"));
    if (S == 0) return CCB_OK;
    for (int64_t s = 0; s < S; s++)
        if (kexit[s] < 1 || kexit[s] > N - 1) { set_error("ccb_gains_backfill: kexit[" + std::to_string(s) + "] = " + std::to_string(kexit[s]) + " outside 1 .. N-1"); return CCB_E_INVALID; }
    const int row = nu * n;
    int* d_k = nullptr;
    CCB_CUDA(cudaMalloc(&d_k, (size_t)S * sizeof(int)));
    cudaError_t e = cudaMemcpy(d_k, kexit, (size_t)S * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        if (row % 2 == 0) {
            gains_backfill_kernel<<<dim3((unsigned)S, 16), 256>>>(table, d_k, (long long)(N - 1) * row, row);
            g_launches++;
            e = cudaGetLastError();
        } else {                     // odd row length (nu n odd): rows are not 16-byte aligned, copy them one by one
            const size_t per = (size_t)(N - 1) * row;
            for (int64_t s = 0; s < S && e == cudaSuccess; s++)
                for (int r = 0; r < kexit[s] - 1 && e == cudaSuccess; r++)
                    e = cudaMemcpyAsync(table + s * per + (size_t)r * row, table + s * per + (size_t)(kexit[s] - 1) * row, (size_t)row * sizeof(double), cudaMemcpyDeviceToDevice, 0);
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    cudaFree(d_k);
    if (e != cudaSuccess) return cuda_fail(e, "ccb_gains_backfill");
    return CCB_OK;
}

int ccb_gains_read(int32_t device, const double* dptr, int64_t offset, double* host, int64_t count) {
    if (!dptr || !host || count < 0 || offset < 0) { set_error("ccb_gains_read: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_device(device, "ccb_gains_read"));
    CCB_CUDA(cudaMemcpy(host, dptr + offset, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost));
    return CCB_OK;
}

int ccb_lqr_synthesize_peers(ccb_mech* m, int32_t device, int32_t S, const double* zd, const double* Fd, const int32_t* eqc, int32_t nu,
                             const double* Q, int64_t Q_bs, const double* R, int64_t R_bs, int32_t N, const ccb_sim_opts* opts, int32_t ntables,
                             double* const* tables, int64_t s_offset, int32_t* kexit, int32_t* iters) {
    if (!m || S < 1 || !zd || nu < 1 || !Fd || !eqc || !Q || !R || N < 2 || ntables < 1 || ntables > 8 || !tables || s_offset < 0) {
        set_error("ccb_lqr_synthesize_peers: bad arguments (1..8 tables)");
        return CCB_E_INVALID;
    }
    if (m->h.ml == 0) { set_error("ccb_lqr_synthesize_peers: mechanism without joints is out of scope"); return CCB_E_UNSUPPORTED; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_lqr_synthesize_peers"));
    CCB_TRY(check_device(device, "ccb_lqr_synthesize_peers"));
    const int n = 12 * m->h.nb;
    const size_t per = (size_t)(N - 1) * nu * n;
    // K4 stores the rows into the tables itself (in-kernel peer stores under the recursion) when the shard is ONE wave of sweep CTAs -- there is
    // nothing else to overlap the transfer with.  Larger shards sweep in groups of one wave and the copy engines move a finished group while the
    // next one sweeps (measured r2 on 4 B200s, 256 Sawyer syntheses per rank: in-kernel stores 139 ms -- the remote stores of two waves of CTAs hold
    // the sweeps up -- against 77 ms for synthesise + ncclAllGather; 2 GPUs 124 vs 135 ms; 8 GPUs, one wave, 43 vs 48 ms).
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const bool dmma = riccati_dmma_eligible(n, nu) && !getenv("CCB_NO_DMMA");
    const bool fused = dmma && (S <= sms || getenv("CCB_PEER_INKERNEL"));
    const bool piped = dmma && !fused;
    if (fused) {
        g_peer_tables.n = ntables;
        g_peer_tables.s_offset = s_offset;
        g_peer_tables.spread = getenv("CCB_PEER_NO_SPREAD") ? 0 : 1;
        g_peer_tables.defer_backfill = g_defer_backfill;
        for (int i = 0; i < ntables; i++) g_peer_tables.t[i] = tables[i];
    }
    double* d_Ku = nullptr;
    int rc = lqr_synth_device(m, device, S, zd, Fd, eqc, nu, Q, Q_bs, R, R_bs, N, opts, &d_Ku, kexit, iters, piped ? ntables : 0, tables, s_offset);
    g_peer_tables.n = 0;
    g_peer_tables.defer_backfill = 0;
    if (rc != CCB_OK) return rc;
    if (!fused && !piped) {            // other kernels: the finished shard is copied to every table (peer copies over NVLink)
        for (int i = 0; i < ntables && rc == CCB_OK; i++)
            if (cudaMemcpy(tables[i] + (size_t)s_offset * per, d_Ku, (size_t)S * per * sizeof(double), cudaMemcpyDefault) != cudaSuccess)
                rc = cuda_fail(cudaGetLastError(), "cudaMemcpy(peer table)");
    }
    cudaFree(d_Ku);
    return rc;
}

// finite-horizon LQR controller over a device-resident gain table [S][N-1][nu][n] (borrowed: the caller keeps it alive and frees it)
int ccb_lqr_ctrl_from_table(ccb_mech* m, int32_t device, int32_t S, int32_t N, int32_t nu, const int32_t* eqc, const double* K_dev, const double* zd,
                            const double* Fd, ccb_ctrl** out) {
    if (!m || S < 1 || N < 2 || nu < 1 || !eqc || !K_dev || !zd || !Fd || !out) { set_error("ccb_lqr_ctrl_from_table: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_lqr_ctrl_from_table"));
    CCB_TRY(check_device(device, "ccb_lqr_ctrl_from_table"));
    const int nb = m->h.nb, n = 12 * nb;
    ccb_ctrl* c = new ccb_ctrl();
    c->device = device; c->B = S;
    std::memset(&c->v, 0, sizeof(c->v));
    c->v.nu = nu; c->v.kind = CCB_CTRL_LQR_FIN; c->v.N = N;
    c->v.K = K_dev; c->v.K_bs = S > 1 ? (long long)(N - 1) * nu * n : 0;
    auto fail = [&](int rc) { ccb_ctrl_destroy(c); return rc; };
    void *p_eqc = nullptr, *p_zd = nullptr, *p_Fd = nullptr;
    if (cudaMalloc(&p_eqc, nu * sizeof(int)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_eqc);
    if (cudaMalloc(&p_zd, (size_t)S * nb * 13 * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_zd);
    if (cudaMalloc(&p_Fd, (size_t)S * nu * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_Fd);
    if (cudaMemcpy(p_eqc, eqc, nu * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p_zd, zd, (size_t)S * nb * 13 * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p_Fd, Fd, (size_t)S * nu * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaMemcpy"));
    c->v.eqc = (const int*)p_eqc;
    c->v.zd = (const double*)p_zd; c->v.zd_bs = S > 1 ? (long long)nb * 13 : 0;
    c->v.Fd = (const double*)p_Fd; c->v.Fd_bs = S > 1 ? (long long)nu : 0;
    { int rcu = finish_uploads(); if (rcu != CCB_OK) return fail(rcu); }
    *out = c;
    return CCB_OK;
}

// TrackingLQR controller (control_trackinglqr!, lqr_tracking.jl:46-71) over a device-resident gain table [S][N-1][nu][n] (borrowed), with the
// desired trajectories [S][N][nb][13] and feed-forward inputs [S][N][nu] copied from the host
int ccb_tracking_ctrl_from_table(ccb_mech* m, int32_t device, int32_t S, int32_t N, int32_t nu, const int32_t* eqc, const double* K_dev,
                                 const double* traj, const double* Ftau, ccb_ctrl** out) {
    if (!m || S < 1 || N < 2 || nu < 1 || !eqc || !K_dev || !traj || !Ftau || !out) { set_error("ccb_tracking_ctrl_from_table: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_tracking_ctrl_from_table"));
    CCB_TRY(check_device(device, "ccb_tracking_ctrl_from_table"));
    const int nb = m->h.nb, n = 12 * nb;
    ccb_ctrl* c = new ccb_ctrl();
    c->device = device; c->B = S;
    std::memset(&c->v, 0, sizeof(c->v));
    c->v.nu = nu; c->v.kind = CCB_CTRL_TRACKING; c->v.N = N;
    c->v.K = K_dev; c->v.K_bs = S > 1 ? (long long)(N - 1) * nu * n : 0;
    auto fail = [&](int rc) { ccb_ctrl_destroy(c); return rc; };
    void *p_eqc = nullptr, *p_zd = nullptr, *p_Fd = nullptr;
    if (cudaMalloc(&p_eqc, nu * sizeof(int)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_eqc);
    if (cudaMalloc(&p_zd, (size_t)S * N * nb * 13 * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_zd);
    if (cudaMalloc(&p_Fd, (size_t)S * N * nu * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc"));
    c->allocs.push_back(p_Fd);
    if (cudaMemcpy(p_eqc, eqc, nu * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p_zd, traj, (size_t)S * N * nb * 13 * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p_Fd, Ftau, (size_t)S * N * nu * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaMemcpy"));
    c->v.eqc = (const int*)p_eqc;
    c->v.zd = (const double*)p_zd; c->v.zd_bs = S > 1 ? (long long)N * nb * 13 : 0;
    c->v.Fd = (const double*)p_Fd; c->v.Fd_bs = S > 1 ? (long long)N * nu : 0;
    { int rcu = finish_uploads(); if (rcu != CCB_OK) return fail(rcu); }
    *out = c;
    return CCB_OK;
}

int ccb_ctrl_get_gains(ccb_ctrl* c, int32_t n, double* K, int64_t count) {
    if (!c || !K || !c->v.K) { set_error("ccb_ctrl_get_gains: controller has no gain table"); return CCB_E_INVALID; }
    const int64_t rows = c->v.kind == CCB_CTRL_LQR_INF ? 1 : (int64_t)c->v.N - 1;
    const int64_t have = (c->v.K_bs ? (int64_t)c->B : 1) * rows * c->v.nu * n;
    if (count != have) { set_error("ccb_ctrl_get_gains: the table holds " + std::to_string(have) + " doubles"); return CCB_E_INVALID; }
    CCB_CUDA(cudaSetDevice(c->device));
    CCB_CUDA(cudaMemcpy(K, c->v.K, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost));
    return CCB_OK;
}

// TrackingLQR synthesis on the device for S trajectories.  Ku goes either to the host (Ku_host) or stays in d_Ku_all
// ([S][N-1][nu][n], allocated by the caller); traj / Ftau are host arrays; d_traj_keep / d_F_keep (optional, device, whole
// batch) receive copies of the desired trajectory and feed-forward for the controller.
// [s][k][w] -> [k][s][w]: the step-major layout of the tracking pipeline
__global__ void step_major_kernel(const double* __restrict__ in, double* __restrict__ out, int cs, int N, int W) {
    const size_t total = (size_t)cs * N * W;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(e % W);
        const size_t p = e / W;
        const int sI = (int)(p % cs), k = (int)(p / cs);
        out[e] = in[((size_t)sI * N + k) * W + w];
    }
}

static int tracking_synth_device(ccb_mech* m, int device, int S, int N, const double* traj, const double* Ftau, const int32_t* eqc, int nu,
                                 const double* Q, const double* R, const ccb_sim_opts* opts, double* Ku_host, double* d_Ku_all,
                                 double* d_traj_keep, double* d_F_keep, int32_t* kexit, int32_t* iters, int peer_copy_n = 0,
                                 double* const* peer_copy_t = nullptr) {
    const int nb = m->h.nb, ml = m->h.ml, n = 12 * nb;
    // every (trajectory, k) linearisation is independent (lqr_tracking.jl:88); only Pk carries the recursion.
    // Chunk over trajectories so that the per-chunk matrices use at most half of the free HBM (180 GB per B200), in equal chunks.
    const size_t per_traj = (size_t)N * ((size_t)2 * n * n + 2 * (size_t)n * nu + 2 * (size_t)n * ml + 2 * (size_t)ml * (ml + n + nu)) * sizeof(double);
    size_t free_b = 0, total_b = 0;
    CCB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, (free_b / 2) / std::max<size_t>(per_traj, 1)));
    const int passes = (S + chunk - 1) / chunk;
    chunk = (S + passes - 1) / passes;
    Scratch sc;
    double *d_Q, *d_R;
    int* d_eqc;
    CCB_TRY(sc.put(&d_Q, Q, (size_t)n * n)); CCB_TRY(sc.put(&d_R, R, (size_t)nu * nu)); CCB_TRY(sc.put(&d_eqc, (const int*)eqc, (size_t)nu));
    double *d_zd, *d_Fd, *d_A, *d_Bu, *d_Bl, *d_G, *d_lam, *d_s, *d_At, *d_D, *d_Ku = nullptr, *d_ws1, *d_ws2;
    int *d_status, *d_kexit, *d_iters;
    const size_t P = (size_t)chunk * N;
    CCB_TRY(sc.get(&d_zd, P * nb * 13)); CCB_TRY(sc.get(&d_Fd, P * nu));
    CCB_TRY(sc.get(&d_A, P * n * n)); CCB_TRY(sc.get(&d_Bu, P * n * nu)); CCB_TRY(sc.get(&d_Bl, P * n * ml)); CCB_TRY(sc.get(&d_G, P * ml * n));
    CCB_TRY(sc.get(&d_lam, P * ml)); CCB_TRY(sc.get(&d_s, P * nb * 6)); CCB_TRY(sc.get(&d_status, P));
    CCB_TRY(sc.get(&d_At, P * n * n)); CCB_TRY(sc.get(&d_D, P * n * nu));
    if (!d_Ku_all) CCB_TRY(sc.get(&d_Ku, (size_t)chunk * (N - 1) * nu * n));
    CCB_TRY(sc.get(&d_ws1, P * riccati_project_ws_doubles(n, nu, ml)));
    CCB_TRY(sc.get(&d_ws2, (size_t)chunk * riccati_sweep_ws_doubles(n, nu)));
    CCB_TRY(sc.get(&d_kexit, (size_t)chunk)); CCB_TRY(sc.get(&d_iters, (size_t)chunk));
    int* d_cnt; CCB_TRY(sc.get(&d_cnt, (size_t)1));
    // k-pipeline (K4 sizes): the sweep consumes the linearisations from k = N-1 downwards and needs ~1 CTA per trajectory, so the
    // linearisation + projection of the next block of steps runs on a second stream while the sweep works through the previous one
    // (stream-ordered through events; the sweep is launched once per block and carries P through the global workspace).
    // Within a pass the per-(trajectory, step) arrays are laid out step-major, p = k * cs + s, so that a block of steps is contiguous.
    const char* env_pipe = getenv("CCB_TRACK_PIPELINE");
    const int nblk = std::min(16, (N - 1) / 128);
    const bool pipeline = riccati_dmma_eligible(n, nu) && !getenv("CCB_NO_DMMA") && nblk >= 2 && !(env_pipe && env_pipe[0] == '0');
    double *d_zs = nullptr, *d_fs = nullptr;      // trajectory-major staging of the inputs when the caller keeps no device copy
    if (pipeline && !d_traj_keep) CCB_TRY(sc.get(&d_zs, P * nb * 13));
    if (pipeline && !d_F_keep) CCB_TRY(sc.get(&d_fs, P * nu));
    PhaseTimer pt;
    const long long peer_off0 = g_peer_tables.s_offset;           // ccb_tracking_synthesize_peers: slot of trajectory 0 in the peer tables
    struct PeerOffsetRestore { long long v; ~PeerOffsetRestore() { g_peer_tables.s_offset = v; } } peer_restore{peer_off0};
    for (int s0 = 0; s0 < S; s0 += chunk) {
        const int cs = std::min(chunk, S - s0);
        const int Pc = cs * N;
        g_peer_tables.s_offset = peer_off0 + s0;
        double* Kdst = d_Ku_all ? d_Ku_all + (size_t)s0 * (N - 1) * nu * n : d_Ku;
        const double* h_traj = traj + (size_t)s0 * N * nb * 13;
        const double* h_F = Ftau + (size_t)s0 * N * nu;
        if (d_traj_keep) CCB_CUDA(cudaMemcpy(d_traj_keep + (size_t)s0 * N * nb * 13, h_traj, (size_t)Pc * nb * 13 * sizeof(double), cudaMemcpyHostToDevice));
        if (d_F_keep) CCB_CUDA(cudaMemcpy(d_F_keep + (size_t)s0 * N * nu, h_F, (size_t)Pc * nu * sizeof(double), cudaMemcpyHostToDevice));
        CCB_CUDA(cudaMemset(d_lam, 0, std::max<size_t>((size_t)Pc * ml, 1) * sizeof(double)));
        if (!pipeline) {
            CCB_CUDA(cudaMemcpy(d_zd, h_traj, (size_t)Pc * nb * 13 * sizeof(double), cudaMemcpyHostToDevice));
            CCB_CUDA(cudaMemcpy(d_Fd, h_F, (size_t)Pc * nu * sizeof(double), cudaMemcpyHostToDevice));
            pt.mark(0);
            CCB_TRY(linearize_device(&m->h, device, Pc, d_zd, d_Fd, d_eqc, nu, opts, d_A, d_Bu, d_Bl, d_G, d_lam, d_s, d_status, 0));
            pt.mark(1);
            CCB_TRY(project_linearisations(&m->h, device, Pc, nu, d_A, d_Bu, d_Bl, d_G, d_At, d_D, d_ws1, 0));
            pt.mark(2);
            CCB_TRY(riccati_sweep_device(cs, n, nu, N, d_At, d_D, (long long)n * n, (long long)n * nu, (long long)N * n * n, (long long)N * n * nu, d_Q, 0, d_R,
                                         0, Kdst, d_kexit, d_iters, d_ws2, 0));
            pt.mark(3);
            CCB_CUDA(cudaDeviceSynchronize());
            pt.finish(s0 > 0);
            if (s0 == 0) g_phase_ms[3] = 0.0;
            CCB_TRY(check_knot_status(d_status, Pc, d_cnt, "TrackingLQR synthesis"));
        } else {
            // step-major copies: row k of trajectory s goes to p = k * cs + s (contiguous upload, transposed on the device)
            const double* src_z = d_traj_keep ? d_traj_keep + (size_t)s0 * N * nb * 13 : nullptr;
            const double* src_f = d_F_keep ? d_F_keep + (size_t)s0 * N * nu : nullptr;
            if (!src_z) { CCB_CUDA(cudaMemcpy(d_zs, h_traj, (size_t)Pc * nb * 13 * sizeof(double), cudaMemcpyHostToDevice)); src_z = d_zs; }
            if (!src_f) { CCB_CUDA(cudaMemcpy(d_fs, h_F, (size_t)Pc * nu * sizeof(double), cudaMemcpyHostToDevice)); src_f = d_fs; }
            step_major_kernel<<<1184, 256>>>(src_z, d_zd, cs, N, nb * 13);
            step_major_kernel<<<296, 256>>>(src_f, d_Fd, cs, N, nu);
            g_launches += 2;
            CCB_CUDA(cudaGetLastError());
            CCB_CUDA(cudaDeviceSynchronize());
            cudaStream_t sa = nullptr, sb = nullptr;
            std::vector<cudaEvent_t> ev;       // per block: lin start, lin end / proj start, proj end, sweep start, sweep end
            auto cleanup = [&]() {
                for (auto e : ev) cudaEventDestroy(e);
                if (sa) cudaStreamDestroy(sa);
                if (sb) cudaStreamDestroy(sb);
            };
            auto bail = [&](int rc) { cudaDeviceSynchronize(); cleanup(); return rc; };
            if (cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking) != cudaSuccess)
                return bail(cuda_fail(cudaGetLastError(), "cudaStreamCreate"));
            ev.resize((size_t)5 * nblk);
            for (auto& e : ev) if (cudaEventCreate(&e) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaEventCreate"));
            const int E = N - 1;                         // entries 0 .. N-2 are used (entry k-1 is the linearisation at step k)
            int rc = CCB_OK;
            for (int c = nblk - 1; c >= 0 && rc == CCB_OK; c--) {
                const int e_lo = (int)((long long)E * c / nblk), e_hi = (int)((long long)E * (c + 1) / nblk);     // [e_lo, e_hi)
                const size_t off = (size_t)e_lo * cs;
                const int Pb = (e_hi - e_lo) * cs;
                cudaEvent_t* e5 = &ev[(size_t)5 * c];
                cudaEventRecord(e5[0], sa);
                rc = linearize_device(&m->h, device, Pb, d_zd + off * nb * 13, d_Fd + off * nu, d_eqc, nu, opts, d_A + off * n * n, d_Bu + off * n * nu,
                                      d_Bl + off * n * ml, d_G + off * ml * n, d_lam + off * ml, d_s + off * nb * 6, d_status + off, sa);
                if (rc != CCB_OK) break;
                cudaEventRecord(e5[1], sa);
                rc = project_linearisations(&m->h, device, Pb, nu, d_A + off * n * n, d_Bu + off * n * nu, d_Bl + off * n * ml, d_G + off * ml * n,
                                            d_At + off * n * n, d_D + off * n * nu, d_ws1 + off * riccati_project_ws_doubles(n, nu, ml), sa);
                if (rc != CCB_OK) break;
                cudaEventRecord(e5[2], sa);
                cudaStreamWaitEvent(sb, e5[2], 0);
                cudaEventRecord(e5[3], sb);
                rc = riccati_sweep_device(cs, n, nu, N, d_At, d_D, (long long)cs * n * n, (long long)cs * n * nu, (long long)n * n, (long long)n * nu, d_Q, 0, d_R, 0,
                                          Kdst, d_kexit, d_iters, d_ws2, sb, e_hi, e_lo + 1, c != nblk - 1);
                cudaEventRecord(e5[4], sb);
            }
            if (rc != CCB_OK) return bail(rc);
            if (cudaDeviceSynchronize() != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "tracking pipeline"));
            double ph[3] = {0, 0, 0};
            for (int c = 0; c < nblk; c++) {
                float ms;
                if (cudaEventElapsedTime(&ms, ev[5 * c], ev[5 * c + 1]) == cudaSuccess) ph[0] += ms;
                if (cudaEventElapsedTime(&ms, ev[5 * c + 1], ev[5 * c + 2]) == cudaSuccess) ph[1] += ms;
                if (cudaEventElapsedTime(&ms, ev[5 * c + 3], ev[5 * c + 4]) == cudaSuccess) ph[2] += ms;
            }
            float total = 0.f;
            cudaEventElapsedTime(&total, ev[(size_t)5 * (nblk - 1)], ev[4]);      // first linearisation start .. last sweep end
            for (int i = 0; i < 3; i++) g_phase_ms[i] = (s0 > 0 ? g_phase_ms[i] : 0.0) + ph[i];
            g_phase_ms[3] = (s0 > 0 ? g_phase_ms[3] : 0.0) + total;
            cudaGetLastError();
            cleanup();
            CCB_TRY(check_knot_status(d_status, E * cs, d_cnt, "TrackingLQR synthesis"));      // step-major: entries 0 .. N-2 of every trajectory
        }
        if (peer_copy_n > 0)                                       // kernels that do not store into the peer tables themselves: copy the finished pass
            for (int i = 0; i < peer_copy_n; i++)
                CCB_CUDA(cudaMemcpy(peer_copy_t[i] + (size_t)(peer_off0 + s0) * (N - 1) * nu * n, Kdst, (size_t)cs * (N - 1) * nu * n * sizeof(double), cudaMemcpyDefault));
        if (Ku_host) CCB_CUDA(cudaMemcpy(Ku_host + (size_t)s0 * (N - 1) * nu * n, Kdst, (size_t)cs * (N - 1) * nu * n * sizeof(double), cudaMemcpyDeviceToHost));
        if (kexit) CCB_CUDA(cudaMemcpy(kexit + s0, d_kexit, (size_t)cs * sizeof(int), cudaMemcpyDeviceToHost));
        if (iters) CCB_CUDA(cudaMemcpy(iters + s0, d_iters, (size_t)cs * sizeof(int), cudaMemcpyDeviceToHost));
    }
    return CCB_OK;
}

int ccb_tracking_synthesize(ccb_mech* m, int32_t device, int32_t S, int32_t N, const double* traj, const double* Ftau, const int32_t* eqc,
                            int32_t nu, const double* Q, const double* R, const ccb_sim_opts* opts, double* Ku, int32_t* kexit, int32_t* iters) {
    if (!m || S < 1 || N < 2 || !traj || !Ftau || nu < 1 || !eqc || !Q || !R || !Ku) { set_error("ccb_tracking_synthesize: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_tracking_synthesize"));
    CCB_TRY(check_device(device, "ccb_tracking_synthesize"));
    return tracking_synth_device(m, device, S, N, traj, Ftau, eqc, nu, Q, R, opts, Ku, nullptr, nullptr, nullptr, kexit, iters);
}

// ccb_tracking_synthesize with the gain tables of the S trajectories stored at index s_offset + s of every table (see ccb_lqr_synthesize_peers)
int ccb_tracking_synthesize_peers(ccb_mech* m, int32_t device, int32_t S, int32_t N, const double* traj, const double* Ftau, const int32_t* eqc,
                                  int32_t nu, const double* Q, const double* R, const ccb_sim_opts* opts, int32_t ntables, double* const* tables,
                                  int64_t s_offset, int32_t* kexit, int32_t* iters) {
    if (!m || S < 1 || N < 2 || !traj || !Ftau || nu < 1 || !eqc || !Q || !R || ntables < 1 || ntables > 8 || !tables || s_offset < 0) {
        set_error("ccb_tracking_synthesize_peers: bad arguments (1..8 tables)");
        return CCB_E_INVALID;
    }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_tracking_synthesize_peers"));
    CCB_TRY(check_device(device, "ccb_tracking_synthesize_peers"));
    const bool fused = riccati_dmma_eligible(12 * m->h.nb, nu) && !getenv("CCB_NO_DMMA");
    g_peer_tables.n = fused ? ntables : 0;
    g_peer_tables.s_offset = s_offset;
    g_peer_tables.spread = getenv("CCB_PEER_NO_SPREAD") ? 0 : 1;
    for (int i = 0; i < ntables; i++) g_peer_tables.t[i] = tables[i];
    const int rc = tracking_synth_device(m, device, S, N, traj, Ftau, eqc, nu, Q, R, opts, nullptr, nullptr, nullptr, nullptr, kexit, iters,
                                         fused ? 0 : ntables, tables);
    g_peer_tables.n = 0; g_peer_tables.s_offset = 0;
    return rc;
}

int ccb_tracking_create(ccb_mech* m, int32_t device, int32_t S, int32_t N, const double* traj, const double* Ftau, const int32_t* eqc, int32_t nu,
                        const double* Q, const double* R, const ccb_sim_opts* opts, ccb_ctrl** out, int32_t* kexit, int32_t* iters) {
    if (!m || S < 1 || N < 2 || !traj || !Ftau || nu < 1 || !eqc || !Q || !R || !out) { set_error("ccb_tracking_create: bad arguments"); return CCB_E_INVALID; }
    CCB_TRY(check_eqcs(m->h, eqc, nu, "ccb_tracking_create"));
    CCB_TRY(check_device(device, "ccb_tracking_create"));
    const int nb = m->h.nb, n = 12 * nb;
    ccb_ctrl* c = new ccb_ctrl();
    c->device = device; c->B = S;
    std::memset(&c->v, 0, sizeof(c->v));
    c->v.kind = CCB_CTRL_TRACKING; c->v.nu = nu; c->v.N = N;
    auto fail = [&](int rc) { ccb_ctrl_destroy(c); return rc; };
    void *pK = nullptr, *pz = nullptr, *pF = nullptr, *pe = nullptr;
    if (cudaMalloc(&pK, (size_t)S * (N - 1) * nu * n * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc(K)"));
    c->allocs.push_back(pK);
    if (cudaMalloc(&pz, (size_t)S * N * nb * 13 * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc(zd)"));
    c->allocs.push_back(pz);
    if (cudaMalloc(&pF, (size_t)S * N * nu * sizeof(double)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc(Fd)"));
    c->allocs.push_back(pF);
    if (cudaMalloc(&pe, nu * sizeof(int)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc(eqc)"));
    c->allocs.push_back(pe);
    if (cudaMemcpy(pe, eqc, nu * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMemcpy(eqc)"));
    int rc = tracking_synth_device(m, device, S, N, traj, Ftau, eqc, nu, Q, R, opts, nullptr, (double*)pK, (double*)pz, (double*)pF, kexit, iters);
    if (rc != CCB_OK) return fail(rc);
    c->v.eqc = (const int*)pe;
    c->v.K = (const double*)pK; c->v.K_bs = S > 1 ? (long long)(N - 1) * nu * n : 0;
    c->v.zd = (const double*)pz; c->v.zd_bs = S > 1 ? (long long)N * nb * 13 : 0;
    c->v.Fd = (const double*)pF; c->v.Fd_bs = S > 1 ? (long long)N * nu : 0;
    *out = c;
    return CCB_OK;
}

}  // extern "C"
