// =============================================================================
// ttc_engine.cu — host side of the B200-native TT-cross sweep + the C-ABI of
// include/ttcross_b200.h.
//
// Control flow follows dtt_dmrgg (reference lib/dmrgg.f90:11-1050); every bond
// visit is a short chain of kernels whose control flow (rook-loop termination,
// accept test) lives on the device.  The host's only per-visit duties are the
// lottery (rnd.f90:105-144; kept on the host because its sequentially accumulated
// cumulative weights define the reference semantics) and mirroring the
// pivot tape.  There is no CPU fallback: without a CUDA device every entry point
// that computes returns TTC_ERR_CUDA.
// =============================================================================
#include "../../include/ttcross_b200.h"
#include "ttc_device.cuh"
#include "ttc_visit.cuh"
#include "ttc_sweep.cuh"
#include "ttc_superblock.cuh"
#include "ttc_qr.cuh"
#include "ttc_post.cuh"
#include "ttc_nccl.hpp"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace ttc;

namespace {

typedef unsigned long long u64;

// rnd.f90:128-144
inline int find_d(int n, const double* x, double y) {
    if (n == 0) return 0;
    if (y < x[0]) return 0;
    if (x[n - 1] <= y) return n;
    int s = 1, t = n, pos = (t + s) / 2;
    while (t - s > 1) {
        if (y < x[pos - 1]) t = pos; else s = pos;
        pos = (s + t) / 2;
    }
    return pos;
}

// tt.f90:1228-1245
double erank(int d, const std::vector<int>& n /*1..d*/, const std::vector<int>& r /*0..d*/) {
    if (d <= 0) return -1.0;
    if (d == 1) return 0.0;
    double rk = 0.0;
    for (int i = 1; i <= d; ++i) rk = rk + (double)(r[i - 1] * n[i] * r[i]);
    if (rk == 0.0) return rk;
    int b = r[0] * n[1] + n[d] * r[d];
    if (d == 2) return rk / b;
    int a = 0;
    for (int i = 2; i <= d - 1; ++i) a += n[i];
    return (std::sqrt((double)b * b + 4.0 * a * rk) - b) / (2.0 * a);
}

// Fortran 'e' edit descriptor
std::string fmt_e(double v, int w, int dgt) {
    char buf[128];
    std::string s;
    if (v == 0.0) s = "0." + std::string(dgt, '0') + "E+00";
    else if (std::isnan(v)) s = "NaN";
    else if (std::isinf(v)) s = v > 0 ? "Infinity" : "-Infinity";
    else {
        std::snprintf(buf, sizeof buf, "%.*e", dgt - 1, std::fabs(v));
        std::string t = buf;
        size_t epos = t.find('e');
        int ex = std::atoi(t.c_str() + epos + 1) + 1;
        std::string digits;
        for (size_t i = 0; i < epos; ++i) if (t[i] != '.') digits += t[i];
        s = std::string(v < 0 ? "-" : "") + "0." + digits;
        char eb[16];
        std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
        s += eb;
        if ((int)s.size() > w) { size_t z = (s[0] == '-') ? 1 : 0; if (s[z] == '0') s.erase(z, 1); }
    }
    if ((int)s.size() > w) s = std::string(w, '*');
    if ((int)s.size() < w) s = std::string(w - s.size(), ' ') + s;
    return s;
}

struct PivRec { int it, vrank, bond, ii, jj, kk, qq, upd; double pivot; };

enum KClass { KC_VISITS = 0, KC_LOT, KC_FIBER, KC_REDUCE, KC_SUPERBLOCK, KC_ACCEPT, KC_UPDATE, KC_NBR, KC_EXCHANGE, KC_QUAD, KC_INIT, KC_FINAL, KC_MISC, KC_COUNT };
const char* kclass_names[KC_COUNT] = {"bond_visits_cluster", "lottery_eval", "fiber_eval_residual", "argmax_reduce", "superblock", "accept", "rank1_update",
                                      "neighbour_factors", "exchange", "quadrature", "init", "finalise", "misc"};

}  // namespace

struct ttc_handle {
    // problem
    int kind = TTC_ISING, d = 0, ising_id = 1;
    std::vector<int> n;                 // 1..d (n[0] = n[d+1] = 1)
    std::vector<double> par, aux, quad; // quad concatenated, empty = absent
    bool has_tru = false; double tru = 0;
    bool par_dirty = false;
    int P = 1; std::vector<int> own; bool own_given = false;
    u64 seed = 1; ttc_uniform_cb ucb = nullptr; void* ucb_ctx = nullptr;
    int verbose = 0, device = 0, profile = 0;
    int alloc_device = -1;              // the device the current blocks / streams / windows live on (device may be changed between runs)
    std::string err;

    // run parameters / state
    bool ran = false;
    int Rmax = 0, nmax = 0, nlotmax = 0, piv = 3;
    std::vector<int> rk_h, rks_h;
    std::vector<std::vector<std::array<int, 4>>> vip_h;
    std::vector<u64> rng_k;
    DevPlan plan;
    std::vector<void*> allocs; std::vector<size_t> alloc_bytes;
    std::vector<std::pair<void*, size_t>> host_blocks;
    double* stage_h = nullptr; size_t stage_cap = 0;   // pinned staging for result copy-out
    cudaStream_t stream = nullptr;
    cudaStream_t stream_q = nullptr;                   // overlapped per-sweep quadrature (lower priority than `stream`)
    std::vector<cudaEvent_t> ev_fork, ev_join;         // main -> quadrature stream / back, one pair per sweep of a graph
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evs0 = nullptr, evs1 = nullptr;        // around the persistent sweep kernel
    double sweep_ms = 0; int persistent_used = 0;
    char* log_h = nullptr; size_t log_cap = 0;       // pinned mirror of ctrl | slog | rklog | vlog | rk (one synchronisation for all logs)
    bool quad_cached = false; double quad_value = 0;   // dtt_quad of the finalised train, taken at the end of ttc_dmrgg (one process)
    int* lot_h = nullptr; VisitOut* out_h = nullptr; SweepOut* sweep_h = nullptr;   // pinned
    double* pack_d = nullptr; size_t pack_cap = 0;
    // ttc_bind_cores: a caller-owned host buffer every ttc_dmrgg fills with this process's cores before it returns
    double* bind_out = nullptr; long long bind_cap = 0; bool bind_pinned = false, bind_filled = false; size_t bind_count = 0;
    void* flush_d = nullptr; size_t flush_cap = 0;
    double* initb = nullptr;
    int* ready_h = nullptr;              // pinned mirror of Ctrl::ready
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    std::vector<long long> graph_sig;
    long long graph_nodes[2] = {0, 0};
    long long graph_kc[2][16] = {{0}};
    long long setup_serial = 0;
    int no_graph = 0;
    bool lua_fused_ok = false;           // k_lua_fused replaces k_lua_r_w + k_lua_l_w (+ k_pack_all) when its slabs fit in shared memory
    bool use_wave = true;                // warp-wavefront / shared-memory support kernels (needs Rmax <= 32*MAXRPL)
    size_t sm_contract = 0, sm_lua = 0, sm_mat3 = 0, sm_ext = 0, sm_lot = 0, sm_fiber = 0, sm_sb = 0;
    size_t sm_xf = 0; bool xf_ok = false;             // k_exchange_fused: aux | 2 d | two staged LU tables
    int force_sync = 0, force_host_lottery = 0, force_simple = 0, force_split = 0;
    int exp_mode = 0;                   // 1: deterministic exp (include/ttc_detexp.h)
    int converged = 0;                  // the last run ended on the accuracy criterion
    size_t sm_qinc = 0; int qinc_stage = 0;
    size_t sm_sbt = 0, sm_sbm = 0; bool sbt_ok = false, sbm_ok = false;      // tiled superblock kernel (ttc_superblock.cuh)
    int cluster_size = 16, cluster_threads = 256;   // measured best on B200 (16 x 256 beats the portable 8 x 512 by 7 %)
    size_t sm_visit = 0, sm_sweep = 0; bool cluster_ok = false;
    int sweep_threads = 256, sweep_cluster = 16;   // geometry of the persistent kernel (chosen in setup_device; TTC_SWEEP_THREADS / TTC_SWEEP_CLUSTER)
    bool persist_ok = false;                       // the persistent sweep kernel (ttc_sweep.cuh) fits the device in one cooperative wave
    double* chainS = nullptr;                      // [maxsweeps][P + 1][Rmax^2] chain products of the per-sweep quadrature after the loop
    double* init_scal = nullptr; int* init_ind0 = nullptr;   // device-side initial cross (k_init_pick / k_init_state)
    int nsm = 148;
    // core blocks over processes (one per GPU): NCCL communicator of ttc_comm_init, this process's rank
    NcclComm comm = nullptr; int nproc = 1, prank = 0;
    int timeline = 0; unsigned long long* tlog_d = nullptr; int* tlog_n_d = nullptr;
    size_t mb1_bytes = 0, mb2_count = 0, nbl_send = 0, nbl_recv = 0;   // message sizes per process / neighbour
    // peer-memory transport (CUDA IPC windows); p2p = false -> NCCL send/recv/all-gather
    bool p2p = false; char* win = nullptr; size_t win_bytes = 0; std::vector<char*> peer_ptrs; char** peer_win_d = nullptr;
    unsigned long long mp_run = 0;

    std::vector<int> setup_sig;
    std::vector<double> quad_or_ones() const {
        if (!quad.empty()) return quad;
        size_t tot = 0; for (int p = 1; p <= d; ++p) tot += n[p];
        return std::vector<double>(tot, 1.0);
    }

    // results
    i64 neval = 0; int nsweeps = 0; double seconds = 0, device_ms = 0;
    std::vector<double> s_val, s_neval, s_amax, s_pivotmax, s_erank, s_time;
    std::vector<PivRec> pivlog;
    std::string text;
    long long launches = 0;
    long long kc_launch[KC_COUNT] = {0};
    double kc_ms[KC_COUNT] = {0};
};

namespace {

int run_dmrgg(ttc_handle* h, int maxrank, double accuracy, int pivoting);
std::string g_create_err;

#define NCCL_TRY(h, call)                                                                                   \
    do {                                                                                                    \
        int e_ = (call);                                                                                    \
        if (e_ != 0) {                                                                                      \
            (h)->err = std::string("NCCL error: ") + nccl_api().GetErrorString(e_) + " at " #call;          \
            return TTC_ERR_COMM;                                                                            \
        }                                                                                                   \
    } while (0)
#define CUDA_TRY(h, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            (h)->err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;                 \
            return TTC_ERR_CUDA;                                                                            \
        }                                                                                                   \
    } while (0)

// ----------------------------------------------------------------------------
// process-level caching of device and pinned-host blocks: a handle created for the same problem shape as a
// destroyed one gets its blocks back without cudaMalloc / cudaMallocHost (both cost far more than a sweep)
// ----------------------------------------------------------------------------
struct BlockPool {
    std::multimap<std::pair<int, size_t>, void*> free_dev, free_host;
    size_t cached_dev = 0, cached_host = 0;
    static size_t round(size_t b) { return (b + 511) & ~(size_t)511; }
    cudaError_t get_dev(int dev, size_t bytes, void** p) {
        bytes = round(bytes);
        auto it = free_dev.find({dev, bytes});
        if (it != free_dev.end()) { *p = it->second; free_dev.erase(it); cached_dev -= bytes; return cudaSuccess; }
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) { trim(); e = cudaMalloc(p, bytes); }
        return e;
    }
    void put_dev(int dev, size_t bytes, void* p) {
        bytes = round(bytes);
        if (cached_dev + bytes > ((size_t)4 << 30)) { cudaFree(p); return; }
        free_dev.insert({{dev, bytes}, p}); cached_dev += bytes;
    }
    cudaError_t get_host(size_t bytes, void** p) {
        bytes = round(bytes);
        auto it = free_host.find({0, bytes});
        if (it != free_host.end()) { *p = it->second; free_host.erase(it); cached_host -= bytes; return cudaSuccess; }
        return cudaMallocHost(p, bytes);
    }
    void put_host(size_t bytes, void* p) {
        bytes = round(bytes);
        if (cached_host + bytes > ((size_t)256 << 20)) { cudaFreeHost(p); return; }
        free_host.insert({{0, bytes}, p}); cached_host += bytes;
    }
    void trim() {
        for (auto& kv : free_dev) cudaFree(kv.second);
        free_dev.clear(); cached_dev = 0;
    }
};
BlockPool g_pool;
std::mutex g_pool_mu;

template <class T>
int dev_alloc(ttc_handle* h, T** p, size_t count, bool zero = true) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    { std::lock_guard<std::mutex> lk(g_pool_mu); CUDA_TRY(h, g_pool.get_dev(h->device, bytes, &q)); }
    h->alloc_bytes.push_back(bytes);
    if (zero) CUDA_TRY(h, cudaMemsetAsync(q, 0, bytes, h->stream));
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}
template <class T>
int dev_upload(ttc_handle* h, T** p, const std::vector<T>& v) {
    int st = dev_alloc(h, p, v.size(), false);
    if (st) return st;
    if (!v.empty()) CUDA_TRY(h, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

// all ranks reach this point (tiny NCCL all-gather + stream synchronisation)
int comm_barrier(ttc_handle* h) {
    if (!h->comm || !h->stream) return 0;
    int* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(int) * (h->nproc + 1)) != cudaSuccess) return TTC_ERR_CUDA;
    int e = nccl_api().AllGather(d + h->nproc, d, 1, NCCL_INT8, h->comm, h->stream);
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    return (e || ce != cudaSuccess) ? TTC_ERR_COMM : 0;
}
// unmap the peers' windows, wait until every rank has done so, free our own
void window_teardown(ttc_handle* h) {
    if (!h->win) return;
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int g = 0; g < (int)h->peer_ptrs.size(); ++g)
        if (g != h->prank && h->peer_ptrs[g]) cudaIpcCloseMemHandle(h->peer_ptrs[g]);
    h->peer_ptrs.clear();
    comm_barrier(h);
    cudaFree(h->win); h->win = nullptr; h->win_bytes = 0;
    if (h->peer_win_d) { cudaFree(h->peer_win_d); h->peer_win_d = nullptr; }
    h->p2p = false;
}
// one window per rank: mb1_recv | mb2_recv | nb_recv_l | nb_recv_r | flags; IPC handles travel by NCCL all-gather
int window_setup(ttc_handle* h, size_t w1, size_t w2, size_t slab, size_t rowinv, size_t vlog_bytes, size_t rklog_bytes, size_t chs_bytes) {
    DevPlan& D = h->plan;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o1 = 0, o2 = o1 + al(w1 * h->nproc * 8), o3 = o2 + al(w2 * h->nproc * 8), o4 = o3 + al(rowinv * 8),
                 o5 = o4 + al(slab * 8), o6 = o5 + al(3 * (size_t)h->nproc * 8);
    // persistent sweep kernel (ttc_sweep.cuh): mailboxes | 2 slabs from the left | 2 slabs from the right | visit log | rank log | chain products
    const size_t sl_b = al(((size_t)h->nmax * h->Rmax + h->d + 2 * (size_t)h->Rmax + 2) * 8), sr_b = al(((size_t)h->Rmax * h->nmax + h->d) * 8);
    const size_t o7 = o6 + al((size_t)h->P * sizeof(SweepMail)), o8 = o7 + 2 * sl_b, o9 = o8 + 2 * sr_b, o10 = o9 + al(vlog_bytes),
                 o11 = o10 + al(rklog_bytes), tot = o11 + al(chs_bytes);
    CUDA_TRY(h, cudaMalloc((void**)&h->win, tot));
    h->win_bytes = tot;
    CUDA_TRY(h, cudaMemsetAsync(h->win, 0, tot, h->stream));
    D.win_mb1 = (long long)o1; D.win_mb2 = (long long)o2; D.win_nbl = (long long)o3; D.win_nbr = (long long)o4; D.win_flg = (long long)o5;
    D.mb1_recv = (unsigned long long*)(h->win + o1); D.mb2_recv = (double*)(h->win + o2);
    D.nb_recv_l = (double*)(h->win + o3); D.nb_recv_r = (double*)(h->win + o4);
    D.win_flags = (unsigned long long*)(h->win + o5);
    D.win_mail = (long long)o6; D.win_sl = (long long)o7; D.win_sr = (long long)o8; D.slab_l_bytes = (long long)sl_b; D.slab_r_bytes = (long long)sr_b;
    D.win_vlog = (long long)o9; D.win_rklog = (long long)o10; D.win_chs = (long long)o11;
    D.mail = (SweepMail*)(h->win + o6);
    D.vlog = (VisitOut*)(h->win + o9); D.rklog = (int*)(h->win + o10); D.chainS = (double*)(h->win + o11);
    cudaIpcMemHandle_t mine;
    CUDA_TRY(h, cudaIpcGetMemHandle(&mine, h->win));
    char* dh = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&dh, sizeof(mine) * h->nproc));
    CUDA_TRY(h, cudaMemcpyAsync(dh + sizeof(mine) * h->prank, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
    int e = nccl_api().AllGather(dh + sizeof(mine) * h->prank, dh, sizeof(mine), NCCL_INT8, h->comm, h->stream);
    std::vector<cudaIpcMemHandle_t> all(h->nproc);
    cudaError_t ce = cudaMemcpyAsync(all.data(), dh, sizeof(mine) * h->nproc, cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    cudaFree(dh);
    if (e) { h->err = "window_setup: NCCL all-gather of the IPC handles failed"; return TTC_ERR_COMM; }
    CUDA_TRY(h, ce);
    h->peer_ptrs.assign(h->nproc, nullptr);
    for (int g = 0; g < h->nproc; ++g) {
        if (g == h->prank) { h->peer_ptrs[g] = h->win; continue; }
        void* q = nullptr;
        cudaError_t oe = cudaIpcOpenMemHandle(&q, all[g], cudaIpcMemLazyEnablePeerAccess);
        if (oe != cudaSuccess) { (void)cudaGetLastError(); h->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(oe); return TTC_ERR_COMM; }
        h->peer_ptrs[g] = (char*)q;
    }
    CUDA_TRY(h, cudaMalloc((void**)&h->peer_win_d, sizeof(char*) * h->nproc));
    CUDA_TRY(h, cudaMemcpyAsync(h->peer_win_d, h->peer_ptrs.data(), sizeof(char*) * h->nproc, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    D.peer_win = h->peer_win_d;
    h->p2p = true;
    return 0;
}

void free_device(ttc_handle* h) {
    h->setup_sig.clear();
    // everything below belongs to the device the state was built on, which ttc_set_device may since have left behind
    const int adev = h->alloc_device >= 0 ? h->alloc_device : h->device;
    if (h->alloc_device >= 0) cudaSetDevice(adev);
    if (h->stream) cudaStreamSynchronize(h->stream);
    window_teardown(h);
    for (int g = 0; g < 2; ++g) if (h->gexec[g]) { cudaGraphExecDestroy(h->gexec[g]); h->gexec[g] = nullptr; }
    h->graph_sig.clear();
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < h->allocs.size(); ++i) g_pool.put_dev(adev, h->alloc_bytes[i], h->allocs[i]);
        for (auto& hb : h->host_blocks) g_pool.put_host(hb.second, hb.first);
        if (h->pack_d) { g_pool.put_dev(adev, h->pack_cap * sizeof(double), h->pack_d); }
    }
    h->allocs.clear(); h->alloc_bytes.clear(); h->host_blocks.clear();
    h->lot_h = nullptr; h->out_h = nullptr; h->sweep_h = nullptr; h->ready_h = nullptr; h->stage_h = nullptr; h->stage_cap = 0;
    h->log_h = nullptr; h->log_cap = 0;
    h->pack_d = nullptr; h->pack_cap = 0;
    if (h->ev0) { cudaEventDestroy(h->ev0); h->ev0 = nullptr; }
    if (h->ev1) { cudaEventDestroy(h->ev1); h->ev1 = nullptr; }
    if (h->evs0) { cudaEventDestroy(h->evs0); h->evs0 = nullptr; }
    if (h->evs1) { cudaEventDestroy(h->evs1); h->evs1 = nullptr; }
    for (cudaEvent_t e : h->ev_fork) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_join) cudaEventDestroy(e);
    h->ev_fork.clear(); h->ev_join.clear();
    if (h->stream_q) { cudaStreamDestroy(h->stream_q); h->stream_q = nullptr; }
    if (h->stream) { cudaStreamDestroy(h->stream); h->stream = nullptr; }
    h->alloc_device = -1;
}

// kernel launch with accounting; in profile mode each launch is bracketed by events (serialising, diagnostic only)
struct Launcher {
    ttc_handle* h;
    cudaEvent_t a = nullptr, b = nullptr;
    const bool trace = std::getenv("TTC_TRACE") != nullptr;
    explicit Launcher(ttc_handle* hh) : h(hh) {
        if (h->profile) { cudaEventCreate(&a); cudaEventCreate(&b); }
    }
    ~Launcher() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    template <class F>
    void operator()(KClass kc, F&& f) {
        if (h->profile) cudaEventRecord(a, h->stream);
        f();
        h->launches += 1;
        h->kc_launch[kc] += 1;
        if (trace) {        // TTC_TRACE: name the launch a configuration error belongs to (the error itself stays pending for the caller)
            const cudaError_t e = cudaPeekAtLastError();
            if (e != cudaSuccess) std::fprintf(stderr, "[ttc trace] launch %lld (class %d): %s\n", (long long)h->launches, (int)kc, cudaGetErrorString(e));
        }
        if (h->profile) {
            cudaEventRecord(b, h->stream);
            cudaEventSynchronize(b);
            float ms = 0; cudaEventElapsedTime(&ms, a, b);
            h->kc_ms[kc] += ms;
        }
    }
};

inline int cdiv(i64 a, i64 b) { return (int)((a + b - 1) / b); }

// Opt-in to large dynamic shared memory, MONOTONIC per (device, kernel): the attribute is a per-function cap shared by every handle
// of the process, so a handle with small ranks must never lower what a handle with large ranks has asked for (its next launch
// would fail with "invalid argument").  The cap does not influence occupancy -- that follows the size passed at launch.
template <class F>
cudaError_t optin_smem(F fn, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> cur;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    size_t& c = cur[{dev, reinterpret_cast<const void*>(fn)}];
    if (bytes <= c) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) c = bytes;
    return e;
}


// dispatch on the integrand family
#define KIND_SWITCH(kind, ...)                                                   \
    switch (kind) {                                                              \
        case TTC_ISING:   { constexpr int K = KIND_ISING;   __VA_ARGS__; } break; \
        case TTC_STDNORM: { constexpr int K = KIND_STDNORM; __VA_ARGS__; } break; \
        case TTC_COSCOEF: { constexpr int K = KIND_COSCOEF; __VA_ARGS__; } break; \
        default:          { constexpr int K = KIND_MVN;     __VA_ARGS__; } break; \
    }

// the bond-visit kernels (k_visits, k_sweeps) have a dedicated instance for Ising C (streaming evaluation, ttc_visit.cuh)
#define VISIT_KIND_SWITCH(h, ...)                                                                                     \
    if ((h)->kind == TTC_ISING && (h)->ising_id == 1) { constexpr int K = KIND_ISINGC; __VA_ARGS__; }                 \
    else KIND_SWITCH((h)->kind, __VA_ARGS__)

static int QINC_THREADS = std::getenv("TTC_QINC_THREADS") ? std::atoi(std::getenv("TTC_QINC_THREADS")) : 1024;   // k_quad_inc: staging is latency-bound, more loads in flight
int threads_for(const ttc_handle* h) { return h->kind == TTC_MVN ? 64 : 256; }
size_t aux_smem(const ttc_handle* h) { return (size_t)h->plan.auxsm * sizeof(double); }

int ensure_pack(ttc_handle* h, size_t cnt, bool stage = true) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (cnt > h->pack_cap) {
        const int adev = h->alloc_device >= 0 ? h->alloc_device : h->device;
        if (h->pack_d) g_pool.put_dev(adev, h->pack_cap * sizeof(double), h->pack_d);
        h->pack_d = nullptr; h->pack_cap = 0;
        void* q = nullptr;
        CUDA_TRY(h, g_pool.get_dev(adev, cnt * sizeof(double), &q));
        h->pack_d = (double*)q; h->pack_cap = cnt;
    }
    if (stage && cnt > h->stage_cap) {
        void* q = nullptr;
        CUDA_TRY(h, g_pool.get_host(cnt * sizeof(double), &q));
        h->host_blocks.push_back({q, cnt * sizeof(double)});
        h->stage_h = (double*)q; h->stage_cap = cnt;
    }
    return 0;
}

int check_device(ttc_handle* h) {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt <= 0) {
        h->err = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                 "); ttcross_b200 has no CPU fallback";
        return TTC_ERR_CUDA;
    }
    if (h->device < 0 || h->device >= cnt) { h->err = "bad CUDA device index"; return TTC_ERR_ARG; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    return 0;
}

// ----------------------------------------------------------------------------
// device state construction
// ----------------------------------------------------------------------------
int setup_device(ttc_handle* h, int maxrank) {
    // device buffers are kept between runs of the same shape (nothing needs re-zeroing: every slice is written
    // before it is read), so a repeated ttc_dmrgg does not pay cudaMalloc again
    std::vector<int> sig = {h->d, h->P, maxrank > 0 ? maxrank : 64, h->kind, h->device, h->nproc, h->prank};
    sig.insert(sig.end(), h->own.begin(), h->own.end());
    if (h->stream && sig == h->setup_sig) {
        int st0 = check_device(h);
        if (st0) return st0;
        h->plan.piv = h->piv;
        if (h->par_dirty) {
            CUDA_TRY(h, cudaMemcpyAsync(const_cast<double*>(h->plan.par), h->par.data(), h->par.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            if (h->plan.parT) {             // the padded copy the TMA staging reads
                const int NT = h->plan.NT;
                CUDA_TRY(h, cudaMemcpyAsync(const_cast<double*>(h->plan.parT), h->par.data(), std::min<size_t>(h->nmax, h->par.size()) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
                if (h->kind == TTC_ISING)
                    CUDA_TRY(h, cudaMemcpyAsync(const_cast<double*>(h->plan.parT) + NT, h->par.data() + h->n[1], (size_t)h->nmax * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            }
            h->par_dirty = false;
        }
        CUDA_TRY(h, cudaMemcpyAsync(const_cast<double*>(h->plan.quadw), h->quad_or_ones().data(), h->quad_or_ones().size() * sizeof(double),
                                    cudaMemcpyHostToDevice, h->stream));
        return 0;
    }
    free_device(h);
    h->setup_sig = sig;
    h->par_dirty = false;
    h->setup_serial += 1;
    int st = check_device(h);
    if (st) return st;
    h->alloc_device = h->device;
    CUDA_TRY(h, cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, h->device));   // (cudaGetDeviceProperties costs milliseconds)
    {
        // the sweep stream outranks the quadrature stream: when both have CTAs to place, the bond-visit clusters go first
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(h, cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(h, cudaStreamCreateWithPriority(&h->stream_q, cudaStreamNonBlocking, prio_lo));
    }
    CUDA_TRY(h, cudaEventCreate(&h->ev0));
    CUDA_TRY(h, cudaEventCreate(&h->ev1));
    CUDA_TRY(h, cudaEventCreate(&h->evs0));
    CUDA_TRY(h, cudaEventCreate(&h->evs1));

    const int d = h->d, P = h->P;
    h->Rmax = maxrank > 0 ? maxrank : 64;
    h->nmax = *std::max_element(h->n.begin() + 1, h->n.begin() + d + 1);
    h->nlotmax = 2 * h->Rmax + 2 * h->nmax;
    const int Rmax = h->Rmax;
    DevPlan& D = h->plan;
    std::memset(&D, 0, sizeof D);
    D.d = d; D.P = P; D.Rmax = Rmax; D.nmax = h->nmax; D.piv = h->piv; D.kind = h->kind; D.ising_id = h->ising_id;
    D.nlotmax = h->nlotmax;
    D.nproc = h->nproc; D.prank = h->prank;
    D.v0 = proc_v0(P, h->nproc, h->prank); D.nv = proc_v0(P, h->nproc, h->prank + 1) - D.v0;
    D.vper = 0;
    for (int g = 0; g < h->nproc; ++g) D.vper = std::max(D.vper, proc_v0(P, h->nproc, g + 1) - proc_v0(P, h->nproc, g));
    D.c_lo = h->own[D.v0];
    D.c_hi = (D.v0 + D.nv == P) ? d : h->own[D.v0 + D.nv] - 1;
    D.auxsm = 0;
    if (h->kind == TTC_MVN) {
        size_t need = (size_t)d * d * sizeof(double) + (size_t)Rmax * sizeof(double);
        // the d x d matrix goes to shared memory only while it is small: at d = 64 (32 KB) it would halve the CTAs per SM,
        // and every thread reads the same element at the same time anyway (one broadcast L1 hit)
        if (need <= 100 * 1024 && (size_t)d * d * sizeof(double) <= 8 * 1024 && !(std::getenv("TTC_MVN_AUX_GLOBAL"))) D.auxsm = d * d;
    }
    // shared-memory staging of node/weight values (left + right tables of a bond visit hold d-2 positions in total)
    {
        // (sized for the TMA layout of the persistent kernel: vectors padded to an even length NT, table rows to an even RT)
        const size_t stage_d = 4 * (size_t)((h->nmax + 1) & ~1) + 2 * (size_t)std::max(0, d - 2) * ((Rmax + 1) & ~1);
        const size_t need = ((size_t)D.auxsm + Rmax + stage_d) * sizeof(double) + (size_t)(3 * Rmax + 8) * sizeof(int);
        D.stage = (need <= 200 * 1024 && !h->force_simple) ? 1 : 0;
        D.stage_max = D.stage ? (int)stage_d : 0;
        h->sm_lot = ((size_t)D.auxsm + D.stage_max) * sizeof(double) + (size_t)(3 * Rmax + 8) * sizeof(int);
        h->sm_fiber = ((size_t)D.auxsm + Rmax + D.stage_max) * sizeof(double);
        h->sm_sb = ((size_t)D.auxsm + D.stage_max) * sizeof(double);
    }

    std::vector<i64> offL(d + 2, 0), offR(d + 2, 0), coreOff(d + 2, 0), quadOff(d + 2, 0);
    i64 accL = 0, accR = 0, accC = 0, accQ = 0;
    for (int p = 0; p <= d; ++p) { offL[p] = accL; accL += (i64)Rmax * p; }
    offL[d + 1] = accL;
    for (int p = 0; p <= d + 1; ++p) { offR[p] = accR; if (p <= d) accR += (i64)Rmax * std::max(0, d - p); }
    for (int p = 1; p <= d; ++p) { coreOff[p] = accC; accC += (i64)Rmax * h->n[p] * Rmax; quadOff[p] = accQ; accQ += h->n[p]; }

    int *dn, *down, *drk, *drks, *dvip, *dL, *dR, *dlot;
    i64 *doffL, *doffR, *dcoreOff, *dquadOff;
    double *dpar, *daux, *darg, *dcol, *drow, *dinv, *da1, *db1, *da2, *db2, *dlraw, *dlres, *dquad, *dttqq, *dch, *dch2;
    Partial* dpart; VState* dst; VisitOut* dout; SweepOut* dsw;
    std::vector<int> nv(h->n.begin(), h->n.end());
#define TRY(x) do { int s_ = (x); if (s_) return s_; } while (0)
    TRY(dev_upload(h, &dn, nv));
    TRY(dev_upload(h, &down, h->own));
    TRY(dev_upload(h, &dpar, h->par));
    std::vector<double> auxv = h->aux; if (auxv.empty()) auxv.push_back(0.0);
    TRY(dev_upload(h, &daux, auxv));
    TRY(dev_upload(h, &doffL, offL));
    TRY(dev_upload(h, &doffR, offR));
    TRY(dev_upload(h, &dcoreOff, coreOff));
    TRY(dev_upload(h, &dquadOff, quadOff));
    std::vector<double> qv = h->quad_or_ones();
    TRY(dev_upload(h, &dquad, qv));
    TRY(dev_alloc(h, &dL, (size_t)accL + 1, false));
    TRY(dev_alloc(h, &dR, (size_t)accR + 1, false));
    TRY(dev_alloc(h, &dvip, (size_t)(d + 1) * Rmax * 4));
    TRY(dev_alloc(h, &drk, (size_t)d + 2));
    TRY(dev_alloc(h, &drks, (size_t)d + 2));
    TRY(dev_alloc(h, &darg, (size_t)accC, false));
    TRY(dev_alloc(h, &dcol, (size_t)accC, false));
    TRY(dev_alloc(h, &drow, (size_t)accC, false));
    TRY(dev_alloc(h, &dinv, (size_t)(d + 1) * Rmax * Rmax, false));
    const size_t fsz = (size_t)P * Rmax * h->nmax;
    TRY(dev_alloc(h, &da1, fsz, false)); TRY(dev_alloc(h, &db1, fsz, false)); TRY(dev_alloc(h, &da2, fsz, false)); TRY(dev_alloc(h, &db2, fsz, false));
    TRY(dev_alloc(h, &dlot, (size_t)P * 4 * h->nlotmax));
    TRY(dev_alloc(h, &dlraw, (size_t)P * h->nlotmax)); TRY(dev_alloc(h, &dlres, (size_t)P * h->nlotmax));
    TRY(dev_alloc(h, &dpart, (size_t)P * 2 * GMAX));
    TRY(dev_alloc(h, &dst, (size_t)P)); TRY(dev_alloc(h, &dout, (size_t)P)); TRY(dev_alloc(h, &dsw, 1));
    TRY(dev_alloc(h, &dttqq, (size_t)(d + 1) * Rmax * Rmax));
    { double* dttqy; int* dqext; TRY(dev_alloc(h, &dttqy, (size_t)(d + 1) * Rmax * Rmax)); TRY(dev_alloc(h, &dqext, (size_t)2 * (d + 2))); h->plan.ttqy = dttqy; h->plan.qext = dqext; }
    TRY(dev_alloc(h, &dch, (size_t)(P + 1) * Rmax * Rmax)); TRY(dev_alloc(h, &dch2, (size_t)(P + 1) * Rmax * Rmax));
#undef TRY
    D.n = dn; D.own = down; D.par = dpar; D.aux = daux; D.Lidx = dL; D.Ridx = dR; D.offL = doffL; D.offR = doffR;
    D.vip = dvip; D.rk = drk; D.rks = drks; D.arg = darg; D.col = dcol; D.rowT = drow; D.coreOff = dcoreOff; D.inv = dinv;
    D.acol1 = da1; D.bcol1 = db1; D.arow1 = da2; D.brow1 = db2; D.lot = dlot; D.lraw = dlraw; D.lres = dlres;
    D.part = dpart; D.st = dst; D.out = dout; D.quadw = dquad; D.quadOff = dquadOff; D.ttqq = dttqq; D.chain = dch;
    D.chain2 = dch2; D.sweep_out = dsw;

    {
        const int nn0 = *std::min_element(h->n.begin() + 1, h->n.begin() + d + 1);
        int s0 = dev_alloc(h, &h->initb, (size_t)nn0 * std::max(8, P));
        if (s0) return s0;
    }
    {
        // one pinned block: lot | out | sweep | ready
        size_t b_lot = ((size_t)P * 4 * h->nlotmax * sizeof(int) + 63) & ~(size_t)63;
        size_t b_out = ((size_t)P * sizeof(VisitOut) + 63) & ~(size_t)63;
        size_t b_sw = (sizeof(SweepOut) + 63) & ~(size_t)63;
        size_t tot = b_lot + b_out + b_sw + 64;
        void* hb = nullptr;
        { std::lock_guard<std::mutex> lk(g_pool_mu); CUDA_TRY(h, g_pool.get_host(tot, &hb)); }
        h->host_blocks.push_back({hb, tot});
        char* c = (char*)hb;
        h->lot_h = (int*)c; h->out_h = (VisitOut*)(c + b_lot); h->sweep_h = (SweepOut*)(c + b_lot + b_out);
        h->ready_h = (int*)(c + b_lot + b_out + b_sw);
    }
    {
        int maxnb0 = 0;
        for (int v = 0; v < P; ++v) maxnb0 = std::max(maxnb0, h->own[v + 1] - h->own[v]);
        D.maxnb = maxnb0; D.maxsweeps = Rmax;
        Ctrl* dctrl; VisitOut* dvlog; SweepOut* dslog; int* drklog;
        int s1 = dev_alloc(h, &dctrl, 1); if (s1) return s1;
        s1 = dev_alloc(h, &dvlog, (size_t)Rmax * maxnb0 * P); if (s1) return s1;
        s1 = dev_alloc(h, &dslog, (size_t)Rmax + 1); if (s1) return s1;
        s1 = dev_alloc(h, &drklog, (size_t)(Rmax + 1) * (d + 1)); if (s1) return s1;
        D.ctrl = dctrl; D.vlog = dvlog; D.slog = dslog; D.rklog = drklog;
        unsigned int* dtick; s1 = dev_alloc(h, &dtick, (size_t)P + 1); if (s1) return s1;
        D.tickets = dtick;
        unsigned int* dbt; int* dqs;
        s1 = dev_alloc(h, &dbt, 2); if (s1) return s1;
        s1 = dev_alloc(h, &dqs, (size_t)2 * (d + 2)); if (s1) return s1;
        D.btick = dbt; D.qsnap = dqs;
        if (h->nproc > 1) {
            const size_t w1 = (size_t)mb1_slot_words(maxnb0) * D.vper, w2 = (size_t)mb2_slot_doubles(Rmax) * D.vper;
            const size_t slab = (size_t)Rmax * h->nmax, rowinv = slab + (size_t)Rmax * Rmax;
            unsigned long long *s1, *r1; double *s2, *r2, *sl, *rl, *sr, *rr;
            s1 = nullptr; r1 = nullptr; s2 = r2 = sl = rl = sr = rr = nullptr;
            if (dev_alloc(h, &s1, w1) || dev_alloc(h, &s2, w2) || dev_alloc(h, &sl, slab) || dev_alloc(h, &sr, rowinv)) return TTC_ERR_CUDA;
            D.mb1_send = s1; D.mb2_send = s2; D.nb_send_l = sl; D.nb_send_r = sr;
            h->mb1_bytes = w1 * 8; h->mb2_count = w2; h->nbl_send = slab; h->nbl_recv = rowinv;
            D.peer_win = nullptr; D.win_flags = nullptr;
            VisitOut* vlog_plain = D.vlog; int* rklog_plain = D.rklog;
            int pe = std::getenv("TTC_MP_NCCL") ? TTC_ERR_COMM
                                                : window_setup(h, w1, w2, slab, rowinv, (size_t)Rmax * maxnb0 * P * sizeof(VisitOut), (size_t)(Rmax + 1) * (d + 1) * sizeof(int),
                                                               (size_t)Rmax * (P + 1) * Rmax * Rmax * sizeof(double));   // peer-memory windows (CUDA IPC)...
            if (pe) { D.vlog = vlog_plain; D.rklog = rklog_plain; D.chainS = nullptr; D.mail = nullptr; }
            if (pe) {                                                                                    // ...or NCCL receive buffers
                if (h->win) { cudaFree(h->win); h->win = nullptr; }
                h->p2p = false; D.peer_win = nullptr; D.win_flags = nullptr;
                if (!std::getenv("TTC_MP_NCCL") && std::getenv("TTC_MP_P2P_REQUIRED")) return pe;
                if (dev_alloc(h, &r1, w1 * h->nproc) || dev_alloc(h, &r2, w2 * h->nproc) || dev_alloc(h, &rl, rowinv) || dev_alloc(h, &rr, slab)) return TTC_ERR_CUDA;
                D.mb1_recv = r1; D.mb2_recv = r2; D.nb_recv_l = rl; D.nb_recv_r = rr;
            }
        }
    }

    {
        const size_t R = (size_t)Rmax;
        h->use_wave = (Rmax <= 32 * MAXRPL) && !h->force_simple;
        h->sm_contract = (size_t)h->nmax * sizeof(double);      // k_quad_contract_sm: the weights of one core
        h->sm_lua = (3 * R * R + R) * sizeof(double);
        h->sm_mat3 = 3 * R * R * sizeof(double);
        h->sm_ext = (R * R + R) * sizeof(double);
        if (h->sm_lua > 200 * 1024) h->use_wave = false;
        {   // k_quad_inc: two packed-LU tables + a staging area for the new row and column (as much as fits in ~96 KB)
            const size_t fixed = (4 * R * R + 3 * R) * sizeof(double);
            size_t stage = std::min<size_t>((size_t)(2 * R + 1) * (h->nmax + 1) + 64, fixed < 136 * 1024 ? (200 * 1024 - fixed) / sizeof(double) - 16 : (64 * 1024) / sizeof(double));
            if (fixed + stage * sizeof(double) > 200 * 1024) h->use_wave = false;
            h->qinc_stage = (int)stage; h->sm_qinc = fixed + stage * sizeof(double);
            if (h->use_wave) optin_smem(k_quad_inc, (int)h->sm_qinc);
        }
        if (h->use_wave) {
            if (h->sm_contract > 32 * 1024) optin_smem(k_quad_contract_sm, (int)h->sm_contract);
            optin_smem(k_quad_lua_sm, (int)h->sm_lua);
            optin_smem(k_quad_chain_sm, (int)h->sm_mat3);
            optin_smem(k_quad_tree_sm, (int)h->sm_mat3);
            optin_smem(k_update_nbr_w, (int)h->sm_ext);
            optin_smem(k_exchange_extend_w, (int)h->sm_ext);
            optin_smem(k_lua_r_w, (int)h->sm_ext);
            optin_smem(k_lua_l_w, (int)h->sm_ext);
        }
        // fused finalisation (k_lua_fused): both packed LUs + one slab per warp in shared memory
        h->lua_fused_ok = !h->force_simple && lua_fused_smem(Rmax) <= 200 * 1024 && !std::getenv("TTC_NO_LUA_FUSED");
        if (h->lua_fused_ok) optin_smem(k_lua_fused, (int)lua_fused_smem(Rmax));
    }
    // tiled superblock kernel (ttc_superblock.cuh): column-factor slab + per-tile tables in shared memory
    {
        h->sm_sbt = ((size_t)D.auxsm + D.stage_max + sb_tile_doubles(Rmax, d)) * sizeof(double);
        h->sm_sbm = h->sm_sbt + sb_mma_doubles() * sizeof(double);      // DMMA fast mode: + the parked tile
        h->sbt_ok = D.stage && h->sm_sbt <= 200 * 1024 && !h->force_simple && !std::getenv("TTC_NO_TILED_SUPERBLOCK");
        if (h->sbt_ok) {
            cudaError_t ce = cudaSuccess;
            const int bt = (int)h->sm_sbt;
            KIND_SWITCH(h->kind,
                ce = optin_smem(k_superblock_t<K, 0, 0>, bt);
                if (ce == cudaSuccess) ce = optin_smem(k_superblock_t<K, 1, 0>, bt);
                if (ce == cudaSuccess) ce = optin_smem(k_superblock_t<K, 0, 1>, bt);
                h->sbm_ok = ce == cudaSuccess && h->sm_sbm <= 220 * 1024 && Rmax % 4 == 0 &&
                            optin_smem(k_superblock_t<K, 0, 2>, (int)h->sm_sbm) == cudaSuccess;
                (void)cudaGetLastError();
            );
            if (ce != cudaSuccess) { (void)cudaGetLastError(); h->sbt_ok = false; }
        }
    }
    // cluster kernel of the bond visits (ttc_visit.cuh): needs the staged evaluation inputs and the wavefront updates
    {
        // measured on B200: clusters of 16 x 256 threads while all of them fit the chip at once (config B: 8 partitions, 3.5 ms
        // against 3.8 ms with 8 x 512); with many partitions the portable 8-CTA clusters win (config E, 63 partitions: 59 ms
        // against 100 ms with 16)
        if (D.nv * 16 <= h->nsm) { h->cluster_size = 16; h->cluster_threads = 256; }
        else { h->cluster_size = 8; h->cluster_threads = 256; }
        if (const char* e = std::getenv("TTC_CLUSTER_SIZE")) h->cluster_size = std::atoi(e);
        if (const char* e = std::getenv("TTC_CLUSTER_THREADS")) h->cluster_threads = std::atoi(e);
        const size_t RE = std::max(Rmax, 32);       // (the exchange of the persistent kernel stages a 32 x 32 packed block)
        D.auxsm_p = D.auxsm;
        if (h->kind == TTC_MVN && (size_t)d * d * sizeof(double) <= 64 * 1024) D.auxsm_p = (d * d + 1) & ~1;
        h->sm_visit = ((size_t)D.auxsm + 5 * (size_t)Rmax + RE * RE + RE + D.stage_max + 2) * sizeof(double) + (size_t)(4 * Rmax + 8) * sizeof(int);
        h->sm_sweep = h->sm_visit + (size_t)(std::max(D.auxsm, D.auxsm_p) - D.auxsm) * sizeof(double);
        if (h->sm_sweep > 200 * 1024) { D.auxsm_p = D.auxsm; h->sm_sweep = h->sm_visit; }
        h->cluster_ok = D.stage && h->use_wave && h->sm_visit <= 200 * 1024 && h->cluster_size >= 1 && h->cluster_size <= MAXCS && h->cluster_threads >= 32 &&
                        h->cluster_threads <= VISIT_MAXTHREADS && h->cluster_threads % 32 == 0 &&
                        !std::getenv("TTC_NO_CLUSTER");
        if (h->cluster_ok) {
            cudaError_t ce = cudaSuccess;
            VISIT_KIND_SWITCH(h,
                ce = optin_smem(k_visits<K>, (int)h->sm_visit);
                if (ce == cudaSuccess && h->cluster_size > 8) ce = cudaFuncSetAttribute(k_visits<K>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            );
            if (ce != cudaSuccess) { (void)cudaGetLastError(); h->cluster_ok = false; }
            // clusters of 16 are a non-portable size: ask the driver whether one fits, else fall back to the portable 8 x 512
            auto fits = [&](int cs, int tb) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs, 1, 1); cfg.blockDim = dim3(tb, 1, 1); cfg.dynamicSmemBytes = h->sm_visit;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int ncl = 0; cudaError_t e2 = cudaSuccess;
                VISIT_KIND_SWITCH(h, e2 = cudaOccupancyMaxActiveClusters(&ncl, k_visits<K>, &cfg));
                if (e2 != cudaSuccess) { (void)cudaGetLastError(); return false; }
                return ncl >= 1;
            };
            if (h->cluster_ok && !fits(h->cluster_size, h->cluster_threads)) {
                h->cluster_size = 8; h->cluster_threads = 256;
                if (!fits(8, 256)) h->cluster_ok = false;
            }
        }
    }
    // persistent sweep kernel: same cluster shape and shared memory; every cluster must be resident at once (cooperative launch)
    {
        if (!h->p2p) {                 // one process (or the NCCL transport, which keeps the per-sweep schedule): plain allocations
            SweepMail* dmail = nullptr;
            { int s1 = dev_alloc(h, &dmail, (size_t)P); if (s1) return s1; }
            D.mail = dmail; D.win_mail = 0; D.chainS = nullptr;
        }
        h->persist_ok = false;
        // geometry of the persistent kernel: every cluster of the process must be resident at once.  The kernel runs at one
        // CTA of 256 threads per SM (255 registers: the factor prefetch of the fiber loop), and this B200 has seven GPCs of
        // ~20 SMs and one of 8 (measured with cudaOccupancyMaxActiveClusters: 7 clusters of 12..16 CTAs at one CTA per SM, 15 of
        // 8), so the candidates are tried from the most threads per partition downwards and the first that fits is taken:
        // 16 x 256 (up to 7 partitions), 16 x 128 (two CTAs per SM where needed: 14), 8 x 256 (15), 8 x 128, 4 x 128 ...
        h->sweep_threads = 0; h->sweep_cluster = 0;
        if (h->cluster_ok && (h->nproc == 1 || h->p2p) && !std::getenv("TTC_NO_PERSISTENT")) {
            const int maxt = h->kind == TTC_MVN ? SWEEP_MAXTHREADS_MVN : SWEEP_MAXTHREADS;
            std::vector<std::pair<int, int>> cand = {{16, 256}, {16, 128}, {8, 256}, {8, 128}, {4, 128}, {2, 128}, {1, 128}};
            if (h->kind == TTC_MVN) cand = {{16, 192}, {8, 192}, {4, 192}, {4, 128}, {2, 192}, {1, 192}};
            if (std::getenv("TTC_SWEEP_THREADS") || std::getenv("TTC_SWEEP_CLUSTER")) {
                const int cs = std::getenv("TTC_SWEEP_CLUSTER") ? std::atoi(std::getenv("TTC_SWEEP_CLUSTER")) : 16;
                const int th = std::getenv("TTC_SWEEP_THREADS") ? std::atoi(std::getenv("TTC_SWEEP_THREADS")) : 256;
                cand = {{std::max(1, std::min(MAXCS, cs)), std::max(32, std::min(maxt, th / 32 * 32))}};
            }
            cudaError_t ce = cudaSuccess;
            VISIT_KIND_SWITCH(h,
                ce = optin_smem(k_sweeps<K>, (int)h->sm_sweep);
                if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_sweeps<K>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            );
            for (size_t ci = 0; ci < cand.size() && ce == cudaSuccess && !h->persist_ok; ++ci) {
                const int cs = cand[ci].first, th = std::min(cand[ci].second, maxt);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs, D.nv, 1); cfg.blockDim = dim3(th, 1, 1); cfg.dynamicSmemBytes = h->sm_sweep;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int ncl = 0;
                cudaError_t c2 = cudaSuccess;
                VISIT_KIND_SWITCH(h, c2 = cudaOccupancyMaxActiveClusters(&ncl, k_sweeps<K>, &cfg));
                if (c2 != cudaSuccess) { (void)cudaGetLastError(); ncl = 0; }
                h->persist_ok = ncl >= D.nv && P <= 64;
                if (h->persist_ok) { h->sweep_cluster = cs; h->sweep_threads = th; }
                if (std::getenv("TTC_TRACE")) std::fprintf(stderr, "[ttc trace] persistent kernel: %d clusters of %d x %d threads resident at once (need %d), smem %zu B: %s\n", ncl, cs, th, D.nv, h->sm_sweep, h->persist_ok ? "on" : "off");
            }
            if (ce != cudaSuccess) (void)cudaGetLastError();
        }
        if (h->persist_ok) {
            if (!h->p2p) {
                double* dcs = nullptr;
                int s1 = dev_alloc(h, &dcs, (size_t)Rmax * (P + 1) * Rmax * Rmax, false); if (s1) return s1;
                D.chainS = dcs;
            }
            h->chainS = D.chainS;
            { double* sc = nullptr; int* i0 = nullptr; if (dev_alloc(h, &sc, 4) || dev_alloc(h, &i0, (size_t)d + 4)) return TTC_ERR_CUDA; h->init_scal = sc; h->init_ind0 = i0; }
            // value tables + padded node / weight vectors for the TMA staging (ttc_visit.cuh)
            D.RT = (Rmax + 1) & ~1; D.NT = (h->nmax + 1) & ~1;
            const size_t tl = (size_t)(accL / Rmax) * D.RT + 2, tr = (size_t)(accR / Rmax) * D.RT + 2;
            double *xl = nullptr, *wl = nullptr, *xr = nullptr, *wr = nullptr, *pt = nullptr;
            if (dev_alloc(h, &xl, tl) || dev_alloc(h, &wl, tl) || dev_alloc(h, &xr, tr) || dev_alloc(h, &wr, tr)) return TTC_ERR_CUDA;
            std::vector<double> ptv(2 * (size_t)D.NT, 0.0);
            for (int x = 0; x < h->nmax && x < (int)h->par.size(); ++x) ptv[x] = h->par[x];
            if (h->kind == TTC_ISING) for (int x = 0; x < h->nmax; ++x) ptv[D.NT + x] = h->par[h->n[1] + x];
            if (dev_upload(h, &pt, ptv)) return TTC_ERR_CUDA;
            D.XLg = xl; D.WLg = wl; D.XRg = xr; D.WRg = wr; D.parT = pt;
            optin_smem(k_quad_lua_all, (int)h->sm_lua);
            optin_smem(k_quad_chain_all, (int)h->sm_mat3);
            optin_smem(k_quad_tree_all, (int)h->sm_mat3);
        }
    }
    // opt in to more than 48 KB of dynamic shared memory where the staging areas need it
    {
        const int bl = (int)h->sm_lot, bf = (int)h->sm_fiber, bs = (int)h->sm_sb, ba = (int)aux_smem(h);
        KIND_SWITCH(h->kind,
            if (bl > 32 * 1024) optin_smem(k_lot<K>, bl);
            if (bf > 32 * 1024) { optin_smem(k_fiber<K, 0>, bf);
                                  optin_smem(k_fiber<K, 1>, bf); }
            if (bs > 32 * 1024) { optin_smem(k_superblock<K, 0>, bs);
                                  optin_smem(k_superblock<K, 1>, bs); }
            if (ba + 16 * h->d > 32 * 1024) optin_smem(k_exchange_corner<K>, ba + 16 * h->d);
            {
                h->sm_xf = (size_t)ba + ((size_t)2 * h->d + (size_t)2 * h->Rmax * h->Rmax + h->Rmax) * sizeof(double);
                h->xf_ok = h->sm_xf <= (size_t)200 * 1024 &&
                           optin_smem(k_exchange_fused<K>, (int)h->sm_xf) == cudaSuccess;
            }
            if (ba > 32 * 1024) {
                                  optin_smem(k_init_search<K>, ba);
                                  optin_smem(k_init_cross<K>, ba); }
        );
    }
    return 0;
}

// ----------------------------------------------------------------------------
// host lottery for one bond visit of one virtual rank (dmrgg.f90:425-452, rnd.f90:105-126)
// ----------------------------------------------------------------------------
void host_lottery(ttc_handle* h, int v, int p, int r0, int r1, int n1, int n2, int r2, int* lot /*[4][nlotmax]*/,
                  std::vector<double>& pcol, std::vector<double>& prow, std::vector<double>& u) {
    const int nlot = r0 + n1 + n2 + r2;
    const int m = r0 * n1, n = n2 * r2;
    pcol.assign(m + 1, 1.0); prow.assign(n + 1, 1.0);    // slot x+1 temporarily holds weight x
    for (int s = 0; s < r1; ++s) {
        const auto& t = h->vip_h[p][s];
        pcol[(t[0] - 1) + (size_t)r0 * (t[1] - 1) + 1] = 0.0;
        prow[(t[2] - 1) + (size_t)n2 * (t[3] - 1) + 1] = 0.0;
    }
    double scol = 0.0, srow = 0.0;
    for (int i = 1; i <= m; ++i) scol = scol + pcol[i];
    for (int j = 1; j <= n; ++j) srow = srow + prow[j];
    pcol[0] = 0.0; for (int i = 1; i <= m; ++i) pcol[i] = pcol[i - 1] + pcol[i] / scol;
    prow[0] = 0.0; for (int j = 1; j <= n; ++j) prow[j] = prow[j - 1] + prow[j] / srow;
    u.resize(2 * (size_t)nlot);
    if (h->ucb) h->ucb(h->ucb_ctx, v, 2 * nlot, u.data());
    else for (int x = 0; x < 2 * nlot; ++x) u[x] = stream_uniform(h->seed, v, h->rng_k[v] + x);
    h->rng_k[v] += 2 * (u64)nlot;
    const int L = h->nlotmax;
    for (int x = 0; x < nlot; ++x) {
        int c = find_d(m + 1, pcol.data(), u[x]);        if (c > m) c = m;
        int w = find_d(n + 1, prow.data(), u[nlot + x]); if (w > n) w = n;
        lot[x] = (c - 1) % r0 + 1;
        lot[L + x] = (c - 1) / r0 + 1;
        lot[2 * L + x] = (w - 1) % n2 + 1;
        lot[3 * L + x] = (w - 1) / n2 + 1;
    }
}

void host_dims(const ttc_handle* h, int v, int dir, int pp, int& active, int& p, int& r0, int& r1, int& r2) {
    int lo = h->own[v], hi = h->own[v + 1];
    active = (pp <= hi - lo);
    p = (dir == 1) ? lo + pp - 1 : hi - pp;
    if (!active) p = lo;
    r0 = (p - 1 >= lo) ? h->rk_h[p - 1] : h->rks_h[p - 1];
    r1 = h->rk_h[p];
    r2 = (p + 1 <= hi - 1) ? h->rk_h[p + 1] : h->rks_h[p + 1];
}

// ----------------------------------------------------------------------------
// exchange between processes (one per GPU) after a sweep.  Phase 1 replaces the tape chains, the MAX allreduce inputs
// and the LEFT/RIGHT block share of the reference (dmrgg.f90:763-958, dmrggmp.f90:572-629) plus dtt_lua's inv hand-off
// (:1209-1246): ONE NCCL group = an all-gather of the per-virtual-rank mailboxes + a send/recv pair per neighbour.
// Phase 2 (after the corner evaluations and the per-rank quadrature chains) all-gathers chain products, amax, neval.
// Message sizes are capacity-sized (the ranks live on the device; the host never waits for them).
// ----------------------------------------------------------------------------
int mp_phase1(ttc_handle* h, Launcher& L) {
    const DevPlan& D = h->plan;
    cudaStream_t s = h->stream;
    NcclApi& N = nccl_api();
    L(KC_EXCHANGE, [&] { k_mp_pack1<<<D.nv + 2, 256, 0, s>>>(D); });
    if (h->p2p) {     // stores into the peers' windows + flags; the unpack kernels wait on the flags
        L(KC_EXCHANGE, [&] { k_mp_unpack1<<<D.P, 128, 0, s>>>(D); });
        L(KC_EXCHANGE, [&] { k_mp_unpack1b<<<dim3(8, 2), 256, 0, s>>>(D); });
        return 0;
    }
    NCCL_TRY(h, N.GroupStart());
    int e = N.AllGather(D.mb1_send, D.mb1_recv, h->mb1_bytes, NCCL_INT8, h->comm, s);
    if (!e && h->prank > 0) {
        e = N.Send(D.nb_send_l, h->nbl_send, NCCL_FLOAT64, h->prank - 1, h->comm, s);
        if (!e) e = N.Recv(D.nb_recv_l, h->nbl_recv, NCCL_FLOAT64, h->prank - 1, h->comm, s);
    }
    if (!e && h->prank < h->nproc - 1) {
        e = N.Send(D.nb_send_r, h->nbl_recv, NCCL_FLOAT64, h->prank + 1, h->comm, s);
        if (!e) e = N.Recv(D.nb_recv_r, h->nbl_send, NCCL_FLOAT64, h->prank + 1, h->comm, s);
    }
    int e2 = N.GroupEnd();
    NCCL_TRY(h, e);
    NCCL_TRY(h, e2);
    L(KC_EXCHANGE, [&] { k_mp_unpack1<<<D.P, 128, 0, s>>>(D); });
    L(KC_EXCHANGE, [&] { k_mp_unpack1b<<<dim3(8, 2), 256, 0, s>>>(D); });
    return 0;
}
int mp_phase2(ttc_handle* h, Launcher& L, int final) {
    const DevPlan& D = h->plan;
    cudaStream_t s = h->stream;
    NcclApi& N = nccl_api();
    L(KC_EXCHANGE, [&] { k_mp_pack2<<<D.nv, 256, 0, s>>>(D, final); });
    if (h->p2p) {
        L(KC_EXCHANGE, [&] { k_mp_unpack2<<<D.P, 256, 0, s>>>(D, final); });
        return 0;
    }
    NCCL_TRY(h, N.AllGather(D.mb2_send, D.mb2_recv, h->mb2_count, NCCL_FLOAT64, h->comm, s));
    L(KC_EXCHANGE, [&] { k_mp_unpack2<<<D.P, 256, 0, s>>>(D, final); });
    return 0;
}

// quadrature of the current cores -> sweep_out->val.  with_lua: the per-sweep path of dmrgg.f90:975-993.
// Several processes: each contracts its own cores and chains its own virtual ranks; the chain products are
// all-gathered (phase 2) and every process runs the reference's binary tree (dmrgg.f90:1355-1405) on all of them.
// log_maxrank > 0: the last kernel also runs the end-of-sweep bookkeeping (k_sweep_log's body); returns 1 in *logged then.
// log_maxrank < 0: overlapped mode (second stream `qs`, rank snapshot, value recorded by quad_record; see k_sweep_log).
int launch_quad(ttc_handle* h, Launcher& L, bool with_lua, bool use_weights, int final = 0, int log_maxrank = 0, bool* logged = nullptr,
                cudaStream_t qs = nullptr) {
    if (logged) *logged = false;
    const DevPlan& D = h->plan;
    cudaStream_t s = qs ? qs : h->stream;
    const int ovl = log_maxrank < 0 ? 1 : 0;
    const int R = h->Rmax;
    const int ncore = D.c_hi - D.c_lo + 1;
    if (!h->use_wave) {
        L(KC_QUAD, [&] { k_quad_contract<<<dim3(cdiv((i64)R * R, 256), ncore), 256, 0, s>>>(D, use_weights ? 1 : 0); });
        if (with_lua) L(KC_QUAD, [&] { k_quad_lua<<<ncore, 128, 0, s>>>(D); });
        L(KC_QUAD, [&] { k_quad_chain<<<D.nv, 256, 0, s>>>(D); });
        if (h->nproc > 1) { int e = mp_phase2(h, L, final); if (e) return e; }
        L(KC_QUAD, [&] { k_quad_tree<<<1, 256, 0, s>>>(D); });
        return 0;
    }
    if (with_lua && !h->force_split) {
        // per-sweep path: only the new row / column of every contracted core (k_quad_inc)
        L(KC_QUAD, [&] { k_quad_inc<<<ncore, QINC_THREADS, h->sm_qinc, s>>>(D, use_weights ? 1 : 0, h->qinc_stage, ovl); });
    } else {
        L(KC_QUAD, [&] { k_quad_contract_sm<<<dim3(cdiv(R, QC_WARPS), ncore), 32 * QC_WARPS, h->sm_contract, s>>>(D, use_weights ? 1 : 0); });
        if (with_lua) L(KC_QUAD, [&] { k_quad_lua_sm<<<ncore, 512, h->sm_lua, s>>>(D); });
    }
    L(KC_QUAD, [&] { k_quad_chain_sm<<<D.nv, 512, h->sm_mat3, s>>>(D, log_maxrank); });
    if (logged && log_maxrank > 0) *logged = true;
    if (h->nproc > 1) { int e = mp_phase2(h, L, final); if (e) return e; }
    for (int q = 1; q < h->P; q *= 2) {
        const int last = (2 * q >= h->P) ? 1 : 0;
        L(KC_QUAD, [&] { k_quad_tree_sm<<<cdiv(h->P, 2 * q), 512, h->sm_mat3, s>>>(D, q, last, log_maxrank); });
    }
    return 0;
}

// ----------------------------------------------------------------------------
// the sweep
// ----------------------------------------------------------------------------
struct Trace {
    bool on; std::chrono::steady_clock::time_point t0; const char* what;
    Trace(const char* w) : on(std::getenv("TTC_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
    void lap(const char* tag) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[ttc trace] %s/%s: %.3f ms\n", what, tag, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

int run_dmrgg(ttc_handle* h, int maxrank, double accuracy, int pivoting) {
    Trace tr("dmrgg");
    h->ran = false;
    h->pivlog.clear(); h->text.clear();
    h->s_val.clear(); h->s_neval.clear(); h->s_amax.clear(); h->s_pivotmax.clear(); h->s_erank.clear(); h->s_time.clear();
    std::fill(h->kc_launch, h->kc_launch + KC_COUNT, 0); std::fill(h->kc_ms, h->kc_ms + KC_COUNT, 0.0);
    auto tstart = std::chrono::steady_clock::now();
    auto timef = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - tstart).count(); };

    const int d = h->d, m = d, l = 1;
    if (h->P >= m) { h->err = "nproc exceeds or equal dimension, cannot proceed"; h->text += h->err + "\n"; return TTC_ERR_NPROC; }
    if (pivoting < -1) { h->err = "dtt_dmrgg: unknown pivoting"; return TTC_ERR_PIVOTING; }
    h->piv = pivoting;
    if (!h->own_given) { h->own.assign(h->P + 1, 0); ttc_share(l, m - 1, h->P, h->own.data()); }
    for (int v = 0; v < h->P; ++v)
        if (h->own[v + 1] < h->own[v] || h->own[0] != 1 || h->own[h->P] != m) { h->err = "bad partition"; return TTC_ERR_ARG; }
    if (h->nproc > h->P) { h->err = "more processes than partitions: call ttc_set_partition with nparts >= the communicator size"; return TTC_ERR_ARG; }
    if (h->nproc > 1 && h->ucb) { h->err = "a uniform callback needs the host lottery, which the multi-process sweep does not support"; return TTC_ERR_ARG; }
    int st = setup_device(h, maxrank);
    if (st) return st;
    tr.lap("setup_device");
    const int P = h->P, Rmax = h->Rmax;
    DevPlan& D = h->plan;
    cudaStream_t s = h->stream;
    Launcher L(h);
    const int TB = threads_for(h);
    const size_t smA = aux_smem(h);
    const size_t smF = h->sm_fiber, smL = h->sm_lot, smS = h->sm_sb;
    const double eps = 2.220446049250313e-16;
    const double small_element = 10 * eps, small_pivot = 1.e-5;
    const bool has_quad = !h->quad.empty();
    char line[512];

    D.piv = h->piv;
    D.exp_mode = h->exp_mode;
    if (h->timeline) {
        if (!h->tlog_d) { CUDA_TRY(h, cudaMalloc((void**)&h->tlog_d, 3 * 65536 * sizeof(unsigned long long))); CUDA_TRY(h, cudaMalloc((void**)&h->tlog_n_d, sizeof(int))); }
        CUDA_TRY(h, cudaMemsetAsync(h->tlog_n_d, 0, sizeof(int), s));
        D.tlog = h->tlog_d; D.tlog_n = h->tlog_n_d; D.tlog_cap = 65536;
    } else { D.tlog = nullptr; D.tlog_n = nullptr; D.tlog_cap = 0; }
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    h->mp_run += 1;
    L(KC_MISC, [&] { k_run_begin<<<1, 32, 0, s>>>(D, h->seed, accuracy >= 0 ? 1 : 0, accuracy, h->mp_run); });

    // ---- initial cross search (dmrgg.f90:150-217)
    const int snum = std::max(8, P);
    const int nn = *std::min_element(h->n.begin() + 1, h->n.begin() + d + 1);
    std::vector<int> shifts(P + 1);
    for (int p = 0; p < P; ++p) shifts[p] = (int)((double)snum * (double)p / P);
    shifts[P] = snum;
    double* db = h->initb;
    KIND_SWITCH(h->kind, L(KC_INIT, [&] { k_init_search<K><<<cdiv((i64)nn * snum, TB), TB, smA, s>>>(D, nn, snum, db); }));
    // the whole sweep loop as one persistent cooperative kernel (ttc_sweep.cuh) whenever the cluster kernel applies and every
    // cluster of this process is resident at once; TTC_NO_PERSISTENT=1 keeps the per-sweep schedule (graphs of k_visits + ...)
    const bool persistent = h->persist_ok && h->cluster_ok && (h->nproc == 1 || h->p2p) && !h->ucb && !h->verbose && !h->force_sync && !h->force_host_lottery &&
                            h->piv >= 0 && !h->force_split && !h->profile && h->use_wave;
    double val = 0, val_prev = 0, t_init = 0;
    i64 nevalall = 0;
    auto push_series = [&](double v_, i64 ne, double am, double pm, double er, double t) {
        h->s_val.push_back(v_); h->s_neval.push_back((double)ne); h->s_amax.push_back(am); h->s_pivotmax.push_back(pm);
        h->s_erank.push_back(er); h->s_time.push_back(t);
    };
    if (persistent) {
        // ---- initial cross entirely on the device (k_init_pick / k_init_state, ttc_sweep.cuh): no host round trip before the sweeps
        L(KC_INIT, [&] { k_init_pick<<<1, 1024, 0, s>>>(D, nn, snum, db, h->init_scal, h->init_ind0); });
        KIND_SWITCH(h->kind, L(KC_INIT, [&] { k_init_cross<K><<<dim3(cdiv(h->nmax, TB), d), TB, smA, s>>>(D); }));
        L(KC_INIT, [&] { k_init_factors<<<dim3(cdiv(h->nmax, 256), d), 256, 0, s>>>(D); });
        {
            const size_t smf = (size_t)d * h->nmax * sizeof(double);
            const int stage_w = (has_quad && 2 * smf <= 200 * 1024) ? 1 : 0;       // quadrature weights staged beside the fibers
            const size_t smi = stage_w ? 2 * smf : smf;
            // the 48 KB default covers static + dynamic shared memory together (the kernel holds ~9 KB of static tables)
            if (smi > 32 * 1024) CUDA_TRY(h, optin_smem(k_init_state, (int)smi));
            L(KC_INIT, [&] { k_init_state<<<1, 1024, smi, s>>>(D, nn, snum, h->init_scal, h->init_ind0, has_quad ? 1 : 0, stage_w); });
        }
        h->rk_h.assign(d + 2, 1); h->rks_h.assign(d + 2, 1);
        h->rng_k.assign(P, 0);
        t_init = timef();
    } else {
    std::vector<double> b((size_t)nn * snum);
    CUDA_TRY(h, cudaMemcpyAsync(b.data(), db, b.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    int gilot = 1; double gmax = std::fabs(b[0]);
    for (int x = 2; x <= nn * snum; ++x) if (std::fabs(b[x - 1]) > gmax) { gmax = std::fabs(b[x - 1]); gilot = x; }
    std::vector<double> amax_v(P, gmax);
    std::vector<i64> neval_v(P, 0);
    if (P == 1) amax_v[0] = gmax;
    else {
        // each rank's own local maximum is overwritten by the MAXLOC result (dmrgg.f90:193-203)
    }
    for (int v = 0; v < P; ++v) neval_v[v] = (i64)nn * (shifts[v + 1] - shifts[v]);
    std::vector<int> ind0(d + 2, 1);
    {
        int sft = (gilot - 1) / nn, k = (gilot - 1) % nn + 1;
        for (int p = 1; p <= d; ++p) ind0[p] = (k - 1 + sft * (p - 1)) % h->n[p] + 1;
    }
    // pivot 1 of every bond: vip, flat index tables, ranks
    {
        std::vector<int> vip((size_t)(d + 1) * Rmax * 4, 0), rk(d + 2, 1);
        for (int p = 0; p <= d; ++p) {
            int* t = &vip[(size_t)p * Rmax * 4];
            t[0] = 1; t[3] = 1;
            t[1] = (p >= 1 && p <= d - 1) ? ind0[p] : 1;
            t[2] = (p >= 1 && p <= d - 1) ? ind0[p + 1] : 1;
        }
        CUDA_TRY(h, cudaMemcpyAsync(D.vip, vip.data(), vip.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaMemcpyAsync(D.rk, rk.data(), rk.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaMemcpyAsync(D.rks, rk.data(), rk.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        i64 accL = (i64)Rmax * d * (d + 1) / 2, accR = accL;
        std::vector<int> Lt((size_t)accL + 1, 0), Rt((size_t)accR + 1, 0);
        i64 oL = 0, oR = 0;
        for (int p = 0; p <= d; ++p) {
            for (int pos = 0; pos < p; ++pos) Lt[(size_t)(oL + (i64)pos * Rmax)] = ind0[pos + 1];
            oL += (i64)Rmax * p;
            for (int pos = 0; pos < d - p; ++pos) Rt[(size_t)(oR + (i64)pos * Rmax)] = ind0[p + pos + 1];
            oR += (i64)Rmax * (d - p);
        }
        // sizes: sum_{p=0..d} Rmax*p == Rmax*d*(d+1)/2 for both tables
        CUDA_TRY(h, cudaMemcpyAsync(D.Lidx, Lt.data(), (size_t)accL * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaMemcpyAsync(D.Ridx, Rt.data(), (size_t)accR * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaStreamSynchronize(s));
    }
    h->rk_h.assign(d + 2, 1); h->rks_h.assign(d + 2, 1);
    h->vip_h.assign(d + 1, {});
    for (int p = 0; p <= d; ++p)
        h->vip_h[p].push_back({1, (p >= 1 && p <= d - 1) ? ind0[p] : 1, (p >= 1 && p <= d - 1) ? ind0[p + 1] : 1, 1});
    h->rng_k.assign(P, 0);

    tr.lap("init_search+tables");
    // ---- initial cross fibers and factors (dmrgg.f90:220-248)
    KIND_SWITCH(h->kind, L(KC_INIT, [&] { k_init_cross<K><<<dim3(cdiv(h->nmax, TB), d), TB, smA, s>>>(D); }));
    L(KC_INIT, [&] { k_init_factors<<<dim3(cdiv(h->nmax, 256), d), 256, 0, s>>>(D); });
    if (has_quad && h->use_wave && !h->force_split && !persistent)      // contracted cores of the rank-1 train: extents (1,1) for the incremental quadrature
        L(KC_INIT, [&] { k_quad_inc<<<D.c_hi - D.c_lo + 1, QINC_THREADS, h->sm_qinc, s>>>(D, 1, h->qinc_stage, 0); });
    // fibers back to the host for the scalar bookkeeping of the '0::' line
    std::vector<std::vector<double>> fib(d + 1);
    for (int p = 1; p <= d; ++p) fib[p].resize(h->n[p]);
    {
        // one strided copy per core
        i64 off = 0;
        for (int p = 1; p <= d; ++p) {
            CUDA_TRY(h, cudaMemcpy2DAsync(fib[p].data(), sizeof(double), D.arg + off, (size_t)Rmax * sizeof(double), sizeof(double),
                                          h->n[p], cudaMemcpyDeviceToHost, s));
            off += (i64)Rmax * h->n[p] * Rmax;
        }
        CUDA_TRY(h, cudaStreamSynchronize(s));
    }
    for (int v = 0; v < P; ++v) {
        for (int p = h->own[v]; p <= h->own[v + 1]; ++p) {
            neval_v[v] += h->n[p];
            for (int j = 0; j < h->n[p]; ++j) amax_v[v] = std::max(amax_v[v], std::fabs(fib[p][j]));
        }
    }
    {
        std::vector<VState> sv(P);
        std::memset(sv.data(), 0, sizeof(VState) * P);
        for (int v = 0; v < P; ++v) {
            sv[v].amax = amax_v[v]; sv[v].pivotmax = -1; sv[v].pivotmin = -1; sv[v].pivotmax_prev = amax_v[v]; sv[v].neval = neval_v[v];
        }
        CUDA_TRY(h, cudaMemcpyAsync(D.st, sv.data(), sizeof(VState) * P, cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaStreamSynchronize(s));
    }
    if (has_quad) {
        std::vector<double> part(P);
        std::vector<i64> qoff(d + 2, 0);
        for (int p = 1; p <= d; ++p) qoff[p + 1] = qoff[p] + h->n[p];
        auto ddot = [&](int p) { double t = 0.0; for (int j = 0; j < h->n[p]; ++j) t = t + fib[p][j] * h->quad[(size_t)qoff[p] + j]; return t; };
        for (int v = 0; v < P; ++v) {
            double x = 1.0;
            for (int p = h->own[v]; p <= h->own[v + 1] - 1; ++p) x = x * ddot(p) / fib[p][ind0[p] - 1];
            if (v == P - 1) x = x * ddot(m);
            part[v] = x;
        }
        val = part[0];
        for (int v = 1; v < P; ++v) val = val * part[v];
        val_prev = val;
    }
    for (int v = 0; v < P; ++v) nevalall += neval_v[v];
    {
        double t2 = timef();
        double er = erank(d, h->n, h->rk_h);
        std::snprintf(line, sizeof line, "%3d%2s rank%5.1f time: %s n_evals: %10lld", 0, "::", er, fmt_e(t2, 9, 3).c_str(), nevalall);
        std::string str = line;
        if (has_quad) str += " val " + fmt_e(val, 20, 14);
        h->text += str + "\n";
        if (h->verbose) { std::puts(str.c_str()); std::fflush(stdout); }
        push_series(val, nevalall, amax_v[0], -1.0, er, t2);
    }

    }
    // ---- main loop (dmrgg.f90:309-1020)
    // Two host modes.  ASYNC (default): the lottery runs on the device, so every sweep is enqueued without waiting;
    // the exit test lives in k_sweep_log and later sweeps turn into no-ops once it fires.  The host polls a
    // pinned mirror of the ready flag only to stop enqueuing early.  SYNC (uniform callback, or verbose progress
    // lines): one stream synchronisation per bond visit so that the host can draw the lottery / print.
    const bool multi = h->nproc > 1;      // several processes: always asynchronous with the device lottery, every process
                                          // enqueues the same number of sweeps (their NCCL calls must pair up)
    const bool sync_mode = !multi && ((h->ucb != nullptr) || h->verbose || h->force_sync);
    const bool dev_lot = multi || ((h->ucb == nullptr) && !h->force_host_lottery);
    const int NV = D.nv;
    D.dev_lottery = dev_lot ? 1 : 0;
    int maxnb = 0;
    for (int v = 0; v < P; ++v) maxnb = std::max(maxnb, h->own[v + 1] - h->own[v]);
    const int last_sweep = (maxrank > 0) ? maxrank - 1 : Rmax - 1;
    std::vector<double> pcol, prow, ubuf;
    int it = 0, cur_it = 0;

    auto visit_log_index = [&](int it_, int pp_, int v_) { return ((size_t)(it_ - 1) * maxnb + (pp_ - 1)) * P + v_; };

    auto enqueue_visit = [&](int dir, int pp, int rb) -> int {
        // rb: upper bound of every rank during this sweep (ranks grow by at most one per sweep)
        int maxcol = rb * h->nmax, maxrow = rb * h->nmax, maxlot = 2 * rb + 2 * h->nmax;
        i64 maxsb = (i64)maxcol * maxrow;
        if (sync_mode) {
            maxcol = maxrow = maxlot = 1; maxsb = 1;
            for (int v = 0; v < P; ++v) {
                int active, p, r0, r1, r2;
                host_dims(h, v, dir, pp, active, p, r0, r1, r2);
                if (!active) continue;
                const int n1 = h->n[p], n2 = h->n[p + 1];
                maxcol = std::max(maxcol, r0 * n1); maxrow = std::max(maxrow, n2 * r2);
                maxlot = std::max(maxlot, r0 + n1 + n2 + r2);
                maxsb = std::max(maxsb, (i64)r0 * n1 * n2 * r2);
                if (h->piv >= 0 && !dev_lot)
                    host_lottery(h, v, p, r0, r1, n1, n2, r2, h->lot_h + (size_t)v * 4 * h->nlotmax, pcol, prow, ubuf);
            }
        }
        const int Gc = std::min(GMAX, cdiv(maxcol, TB)), Gr = std::min(GMAX, cdiv(maxrow, TB));
        if (h->piv == -1) {
            if (h->sbt_ok) {
                const int nrb = cdiv(maxcol, SB_TM);
                const int nsp = std::max(1, std::min({(SB_MINB * h->nsm) / std::max(1, nrb * NV), cdiv(maxrow, SB_TN), GMAX / nrb}));   // whole waves: SB_MINB CTAs per SM
                KIND_SWITCH(h->kind, L(KC_SUPERBLOCK, [&] { k_superblock_t<K, 0, 0><<<dim3(nrb, nsp, NV), SB_TM, h->sm_sbt, s>>>(D, dir, pp, 0, 0, nullptr, nullptr); }));
            } else {
                const int Gs = (int)std::min<i64>(GMAX, std::max<i64>(1, (maxsb + TB - 1) / TB));
                KIND_SWITCH(h->kind, L(KC_SUPERBLOCK, [&] { k_superblock<K, 0><<<dim3(Gs, NV), TB, smS, s>>>(D, dir, pp, 0, 0, nullptr, nullptr); }));
            }
            KIND_SWITCH(h->kind, L(KC_FIBER, [&] { k_fiber<K, 0><<<dim3(Gc, NV), TB, smF, s>>>(D, dir, pp, 2); }));
            KIND_SWITCH(h->kind, L(KC_FIBER, [&] { k_fiber<K, 1><<<dim3(Gr, NV), TB, smF, s>>>(D, dir, pp, 2); }));
        } else {
            if (!dev_lot) CUDA_TRY(h, cudaMemcpyAsync(D.lot, h->lot_h, (size_t)P * 4 * h->nlotmax * sizeof(int), cudaMemcpyHostToDevice, s));
            const int Gl = std::min(GMAX, cdiv(maxlot, TB));
            KIND_SWITCH(h->kind, L(KC_LOT, [&] { k_lot<K><<<dim3(Gl, NV), TB, smL, s>>>(D, dir, pp); }));
            if (h->piv == 0) {
                KIND_SWITCH(h->kind, L(KC_FIBER, [&] { k_fiber<K, 0><<<dim3(Gc, NV), TB, smF, s>>>(D, dir, pp, 1); }));
                KIND_SWITCH(h->kind, L(KC_FIBER, [&] { k_fiber<K, 1><<<dim3(Gr, NV), TB, smF, s>>>(D, dir, pp, 1); }));
            } else {
                // rook loop (dmrgg.f90:515-582): at most 2*piv fibers, alternating, starting with the row in '<<' sweeps
                int isrow = (dir == 2) ? 1 : 0;
                for (int c = 0; c < 2 * h->piv; ++c) {
                    if (!isrow) { KIND_SWITCH(h->kind, L(KC_FIBER, [&] { k_fiber<K, 0><<<dim3(Gc, NV), TB, smF, s>>>(D, dir, pp, 0); })); }
                    else        { KIND_SWITCH(h->kind, L(KC_FIBER, [&] { k_fiber<K, 1><<<dim3(Gr, NV), TB, smF, s>>>(D, dir, pp, 0); })); }
                    isrow ^= 1;
                }
            }
        }
        L(KC_ACCEPT, [&] { k_accept<<<dim3(1, NV), 128, 0, s>>>(D, dir, pp, small_element, small_pivot); });
        // neighbour factors first (they read the old rank), then the rank-1 append whose last CTA bumps r(p)
        if (maxnb > 1) {
            if (h->use_wave) L(KC_NBR, [&] { k_update_nbr_w<<<dim3(cdiv(h->nmax, 8), NV, 2), 256, h->sm_ext, s>>>(D, dir, pp); });
            else L(KC_NBR, [&] { k_update_nbr<<<dim3(cdiv(2 * h->nmax, 64), NV), 64, 0, s>>>(D, dir, pp); });
        }
        L(KC_UPDATE, [&] { k_update_main<<<dim3(std::min(GMAX, cdiv(maxcol + maxrow, 256)), NV), 256, 0, s>>>(D, dir, pp); });
        if (sync_mode) {
            VisitOut* src = D.vlog + visit_log_index(cur_it, pp, 0);
            CUDA_TRY(h, cudaMemcpyAsync(h->out_h, src, (size_t)P * sizeof(VisitOut), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(h, cudaStreamSynchronize(s));
            for (int v = 0; v < P; ++v) {
                const VisitOut& O = h->out_h[v];
                if (!O.active) continue;
                if (O.upd) { h->vip_h[O.bond].push_back({O.ii, O.jj, O.kk, O.qq}); h->rk_h[O.bond] += 1; }
            }
        }
        return 0;
    };
    // one cluster per virtual rank runs the whole visit list of the sweep (ttc_visit.cuh) when the lottery is on the device
    const bool use_cluster = h->cluster_ok && !sync_mode && dev_lot && h->piv >= 0 && !h->force_split;
    // per-sweep quadrature beside the next sweep's bond visits (second stream); TTC_NO_QUAD_OVERLAP=1 keeps it in line
    // Only when the bond-visit clusters of all partitions fit on the device at once and leave SMs over: with several waves
    // of clusters (config E: 63 partitions) the quadrature CTAs get in the way of cluster placement and the sweep slows down.
    bool overlap = has_quad && !multi && !sync_mode && dev_lot && h->use_wave && !h->force_split && !h->profile &&
                   (!use_cluster || NV * h->cluster_size < h->nsm);
    if (const char* e = std::getenv("TTC_QUAD_OVERLAP")) overlap = overlap && std::atoi(e) != 0;      // 0: keep the quadrature in line
    bool q_pending = false; cudaEvent_t q_last = nullptr; int q_slot = 0;
    cudaEvent_t q_prev = nullptr;           // fused sweep: the join event of the quadrature before the last one
    auto quad_rejoin = [&]() -> int {       // the sweep stream waits for the last quadrature (end of a graph / of the run)
        if (q_pending) CUDA_TRY(h, cudaStreamWaitEvent(s, q_last, 0));
        q_pending = false; q_prev = nullptr;
        return 0;
    };
    // fused sweep (single process): the close of the sweep runs in the sweep's last kernel -- k_exchange_fused (corner +
    // extensions + close in one launch), or k_visits itself with one partition; the quadrature then always runs in its
    // overlapped form (own sweep counter, rank snapshot), on the second stream when `overlap`, else in line.
    // The fused exchange spends one warp per mode index: it pays when its grid is about one wave (config E with 62
    // boundaries and a 12k-flop integrand is better off with the separate corner kernel).  TTC_FUSED_SWEEP=0 disables.
    bool fused = use_cluster && !multi && h->use_wave && (P == 1 || (h->xf_ok && cdiv(2 * h->nmax, 8) * (P - 1) <= 4 * h->nsm));
    if (const char* e = std::getenv("TTC_FUSED_SWEEP")) fused = fused && std::atoi(e) != 0;
    const int eff_maxrank = maxrank > 0 ? maxrank : Rmax;        // no maxrank: the rank capacity ends the run
    auto enqueue_sweep = [&](int dir, int rb) -> int {
        if (sync_mode) h->rks_h = h->rk_h;
        if (fused && overlap && q_prev) {
            // the close inside this sweep's kernel overwrites the rank snapshot buffer the quadrature of two sweeps ago read
            CUDA_TRY(h, cudaStreamWaitEvent(s, q_prev, 0));
            q_prev = nullptr;
        }
        if (use_cluster) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(h->cluster_size, NV, 1);
            cfg.blockDim = dim3(h->cluster_threads, 1, 1);
            cfg.dynamicSmemBytes = h->sm_visit;
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = h->cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t ce = cudaSuccess;
            VISIT_KIND_SWITCH(h, L(KC_VISITS, [&] { ce = cudaLaunchKernelEx(&cfg, k_visits<K>, D, dir, small_element, small_pivot, multi ? 0 : 1, (fused && P == 1) ? eff_maxrank : 0); }));
            CUDA_TRY(h, ce);
        } else {
            for (int pp = 1; pp <= maxnb; ++pp) { int e = enqueue_visit(dir, pp, rb); if (e) return e; }
        }
        if (fused) {
            if (P > 1) {
                const int nbnd = boundary_count(D);
                KIND_SWITCH(h->kind, L(KC_EXCHANGE, [&] { k_exchange_fused<K><<<dim3(cdiv(2 * h->nmax, 8), nbnd), 256, h->sm_xf, s>>>(D, eff_maxrank); }));
            }
            if (!has_quad) return 0;
            if (!overlap) return launch_quad(h, L, true, true, 0, -1, nullptr, nullptr);
            const size_t slot = (size_t)q_slot++ % h->ev_fork.size();
            CUDA_TRY(h, cudaEventRecord(h->ev_fork[slot], s));
            CUDA_TRY(h, cudaStreamWaitEvent(h->stream_q, h->ev_fork[slot], 0));
            { int e = launch_quad(h, L, true, true, 0, -1, nullptr, h->stream_q); if (e) return e; }
            CUDA_TRY(h, cudaEventRecord(h->ev_join[slot], h->stream_q));
            q_prev = q_pending ? q_last : nullptr;
            q_last = h->ev_join[slot]; q_pending = true;
            return 0;
        }
        if (P > 1) {
            if (multi) { int e = mp_phase1(h, L); if (e) return e; }
            const int nbnd = boundary_count(D);
            if (!multi && !use_cluster) L(KC_EXCHANGE, [&] { k_allreduce<<<1, 32, 0, s>>>(D); });   // (several processes: folded into k_mp_unpack1b)
            KIND_SWITCH(h->kind, L(KC_EXCHANGE, [&] { k_exchange_corner<K><<<dim3(1, nbnd), TB, smA + 2 * (size_t)d * sizeof(double), s>>>(D); }));
            if (h->use_wave) L(KC_EXCHANGE, [&] { k_exchange_extend_w<<<dim3(cdiv(h->nmax, 8), nbnd, 2), 256, h->sm_ext, s>>>(D); });
            else L(KC_EXCHANGE, [&] { k_exchange_extend<<<dim3(cdiv(2 * h->nmax, 64), nbnd), 64, 0, s>>>(D); });
        }
        if (overlap) {
            // close the sweep now (the exit test needs no quadrature value) and let its quadrature run on the second stream
            // beside the next sweep's bond visits; the close overwrites the rank snapshot the previous quadrature reads
            const size_t slot = (size_t)q_slot++ % h->ev_fork.size();
            if (q_pending) CUDA_TRY(h, cudaStreamWaitEvent(s, q_last, 0));
            L(KC_MISC, [&] { k_sweep_log<<<1, 128, 0, s>>>(D, eff_maxrank, 0); });
            CUDA_TRY(h, cudaEventRecord(h->ev_fork[slot], s));
            CUDA_TRY(h, cudaStreamWaitEvent(h->stream_q, h->ev_fork[slot], 0));
            { int e = launch_quad(h, L, true, true, 0, -1, nullptr, h->stream_q); if (e) return e; }
            CUDA_TRY(h, cudaEventRecord(h->ev_join[slot], h->stream_q));
            q_last = h->ev_join[slot]; q_pending = true;
            return 0;
        }
        bool logged = false;
        if (has_quad) { int e = launch_quad(h, L, true, true, 0, h->force_split ? 0 : eff_maxrank, &logged); if (e) return e; }
        else if (multi) { int e = mp_phase2(h, L, 0); if (e) return e; }
        if (!logged) L(KC_MISC, [&] { k_sweep_log<<<1, 128, 0, s>>>(D, eff_maxrank, 1); });
        return 0;
    };

    tr.lap("init_cross");
    *h->ready_h = 0;
    // In asynchronous mode a sweep is a fixed kernel sequence (sweep number, seed and thresholds live in device memory),
    // so it is captured once per direction into a CUDA graph and replayed: one graph launch per sweep instead of ~25
    // kernel launches.  Grids are sized for the rank capacity; surplus CTAs exit at once.
    // sweeps per graph: even (every graph starts with a '>>' sweep), at most 16, and chosen so that the expected number of
    // sweeps (maxrank - 1 when the rank bound ends the run) leaves as few no-op sweeps as possible at the end.  The overlapped
    // quadrature rejoins at the end of a graph, so longer graphs also expose fewer quadratures.
    int graph_sweeps;
    {
        const int S = std::max(2, (maxrank > 0 ? maxrank : Rmax) - 1);
        const int ng = cdiv(S, 16);
        graph_sweeps = 2 * cdiv(cdiv(S, ng), 2);
    }
    if (const char* e = std::getenv("TTC_GRAPH_SWEEPS")) graph_sweeps = std::max(2, 2 * (std::atoi(e) / 2));
    if (overlap)
        while ((int)h->ev_fork.size() < std::max(graph_sweeps, 2)) {
            cudaEvent_t a = nullptr, b = nullptr;
            CUDA_TRY(h, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            CUDA_TRY(h, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            h->ev_fork.push_back(a); h->ev_join.push_back(b);
        }
    bool persist_quad_pending = false;
    const bool use_graph = !persistent && !sync_mode && !h->profile && !h->no_graph && (!multi || h->p2p || std::getenv("TTC_MP_GRAPH") != nullptr);
    if (persistent) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(h->sweep_cluster, NV, 1);
        cfg.blockDim = dim3(h->sweep_threads, 1, 1);
        cfg.dynamicSmemBytes = h->sm_sweep;
        cfg.stream = s;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = h->sweep_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        cudaError_t ce = cudaSuccess;
        CUDA_TRY(h, cudaEventRecord(h->evs0, s));
        VISIT_KIND_SWITCH(h, L(KC_VISITS, [&] { ce = cudaLaunchKernelEx(&cfg, k_sweeps<K>, D, last_sweep, eff_maxrank, small_element, small_pivot); }));
        if (ce != cudaSuccess && h->nproc == 1) {
            // the driver refused the cooperative launch (e.g. another context holds SMs): not an error of the problem -- run the
            // per-sweep schedule instead (the device state built so far is reused; the initial cross is redone on that path)
            (void)cudaGetLastError();
            if (std::getenv("TTC_TRACE")) std::fprintf(stderr, "[ttc trace] persistent kernel: cooperative launch refused (%s), per-sweep schedule instead\n", cudaGetErrorString(ce));
            h->persist_ok = false;
            return run_dmrgg(h, maxrank, accuracy, pivoting);
        }
        CUDA_TRY(h, ce);
        CUDA_TRY(h, cudaEventRecord(h->evs1, s));
        if (has_quad) {
            // per-sweep quadrature values of ALL sweeps at once (they feed the printed lines only, dmrgg.f90:975-1008)
            const int R = h->Rmax, ncore = D.c_hi - D.c_lo + 1;
            // The contraction reads the raw cores, which the finalisation (dtt_lua below) rewrites in place: it stays on the sweep
            // stream; the rest works on the contracted copies only and runs on the second stream beside the finalisation.
            L(KC_QUAD, [&] { k_quad_contract_sm<<<dim3(cdiv(R, QC_WARPS), ncore), 32 * QC_WARPS, h->sm_contract, s>>>(D, 1); });
            if (h->ev_fork.empty()) {
                cudaEvent_t a = nullptr, b = nullptr;
                CUDA_TRY(h, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
                CUDA_TRY(h, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
                h->ev_fork.push_back(a); h->ev_join.push_back(b);
            }
            cudaStream_t qs = h->stream_q;
            CUDA_TRY(h, cudaEventRecord(h->ev_fork[0], s));
            CUDA_TRY(h, cudaStreamWaitEvent(qs, h->ev_fork[0], 0));
            L(KC_QUAD, [&] { k_quad_lua_all<<<ncore, 512, h->sm_lua, qs>>>(D); });
            L(KC_QUAD, [&] { k_quad_chain_all<<<dim3(NV, last_sweep), 512, h->sm_mat3, qs>>>(D, h->chainS); });
            L(KC_QUAD, [&] { k_quad_tree_all<<<last_sweep, 512, h->sm_mat3, qs>>>(D, h->chainS); });
            CUDA_TRY(h, cudaEventRecord(h->ev_join[0], qs));
            persist_quad_pending = true;
        }
    }
    if (use_graph) {
        std::vector<long long> gsig = {(long long)h->piv, (long long)has_quad, (long long)maxrank, (long long)h->use_wave, (long long)dev_lot,
                                       (long long)h->setup_serial, (long long)h->timeline, (long long)use_cluster,
                                       (long long)h->cluster_size, (long long)h->cluster_threads};
        gsig.push_back(graph_sweeps);
        gsig.push_back((long long)overlap);
        gsig.push_back((long long)fused);
        gsig.push_back((long long)h->exp_mode);
        if (gsig != h->graph_sig) {
            // ONE graph holds `graph_sweeps` consecutive sweeps ('>>', '<<', '>>', ...): fewer graph boundaries on the device.
            // Sweeps past the exit condition are no-ops (the ready flag is tested by every kernel).
            for (int gdir = 0; gdir < 2; ++gdir) if (h->gexec[gdir]) { cudaGraphExecDestroy(h->gexec[gdir]); h->gexec[gdir] = nullptr; }
            cudaGraph_t g = nullptr;
            CUDA_TRY(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            long long l0 = h->launches;
            long long kc0[KC_COUNT];
            std::copy(h->kc_launch, h->kc_launch + KC_COUNT, kc0);
            int e = 0;
            for (int q = 0; q < graph_sweeps && !e; ++q) e = enqueue_sweep(q % 2 == 0 ? 1 : 2, Rmax);
            if (!e) e = quad_rejoin();
            h->graph_nodes[0] = h->launches - l0;
            h->launches = l0;
            for (int c = 0; c < KC_COUNT; ++c) { h->graph_kc[0][c] = h->kc_launch[c] - kc0[c]; h->kc_launch[c] = kc0[c]; }
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (e) return e;
            CUDA_TRY(h, ce);
            CUDA_TRY(h, cudaGraphInstantiate(&h->gexec[0], g, 0));
            cudaGraphDestroy(g);
            h->graph_sig = gsig;
        }
    }
    tr.lap("graph_capture");
    for (it = 1; !persistent && it <= last_sweep; ++it) {
        if (!multi && *(volatile int*)h->ready_h) break;   // device already reached its exit condition
        int e = 0;
        if (use_graph) {
            CUDA_TRY(h, cudaGraphLaunch(h->gexec[0], s));  // sweeps it .. it + graph_sweeps - 1
            h->launches += h->graph_nodes[0];
            for (int c = 0; c < KC_COUNT; ++c) h->kc_launch[c] += h->graph_kc[0][c];
            it += graph_sweeps - 1;
        } else {
            cur_it = it;
            e = enqueue_sweep(2 - it % 2, std::min(it + 1, Rmax));
        }
        if (e) return e;
        CUDA_TRY(h, cudaMemcpyAsync(h->ready_h, &D.ctrl->ready, sizeof(int), cudaMemcpyDeviceToHost, s));
        if (sync_mode) {
            CUDA_TRY(h, cudaStreamSynchronize(s));
            CUDA_TRY(h, cudaGetLastError());
            if (h->verbose) {
                SweepOut so; std::vector<int> rkl(d + 1);
                CUDA_TRY(h, cudaMemcpy(&so, D.slog + it, sizeof so, cudaMemcpyDeviceToHost));
                if (so.valid) {
                    double er = erank(d, h->n, h->rk_h);
                    std::snprintf(line, sizeof line, "%3d%2s rank%5.1f time: %s n_evals: %10lld", it, (2 - it % 2) == 1 ? ">>" : "<<", er,
                                  fmt_e(timef(), 9, 3).c_str(), so.neval);
                    std::string str = line;
                    if (has_quad) str += (h->has_tru ? " err " + fmt_e(std::fabs(1.0 - so.val / h->tru), 8, 3) : " cnv " + fmt_e(std::fabs(1.0 - so.val / val_prev), 8, 3)) + " val " + fmt_e(so.val, 20, 14);
                    val_prev = so.val;
                    std::puts(str.c_str()); std::fflush(stdout);
                }
            }
        }
    }

    { int e = quad_rejoin(); if (e) return e; }      // (eager mode; a graph rejoins at its end) the finalisation rewrites the cores
    // ---- finalise (dmrgg.f90:1028-1029); not gated by the ready flag.  Each process finalises the cores it owns.
    const int ncore_own = D.c_hi - D.c_lo + 1;
    // ttc_bind_cores: the packed copy of this process's cores is formed on the device right here (by the fused finalisation
    // kernel, else by k_pack_all below), with offsets from the device-side ranks; its transfer into the caller's buffer starts
    // as soon as the logs (which carry the ranks) have arrived, and runs beside the host's log processing below
    h->bind_filled = false;
    const bool bound = h->bind_out != nullptr;
    if (bound) {
        size_t capcnt = 0;
        for (int k = D.c_lo; k <= D.c_hi; ++k) capcnt += (size_t)std::min<i64>(Rmax, k == 1 ? 1 : Rmax) * h->n[k] * std::min<i64>(Rmax, k == d ? 1 : Rmax);
        { int st = ensure_pack(h, capcnt, !h->bind_pinned); if (st) return st; }
    }
    // the logs travel with the same synchronisation: ctrl | slog | rklog | vlog | rk into one pinned block (capacity-sized)
    const size_t lb_ctrl = 0, lb_slog = 256, lb_rklog = lb_slog + (((size_t)(Rmax + 1) * sizeof(SweepOut) + 63) & ~(size_t)63),
                 lb_vlog = lb_rklog + (((size_t)(Rmax + 1) * (d + 1) * sizeof(int) + 63) & ~(size_t)63),
                 lb_rk = lb_vlog + (((size_t)Rmax * maxnb * P * sizeof(VisitOut) + 63) & ~(size_t)63),
                 lb_tot = lb_rk + (((size_t)(d + 2) * sizeof(int) + 63) & ~(size_t)63);
    if (lb_tot > h->log_cap) {
        void* q = nullptr;
        { std::lock_guard<std::mutex> lk(g_pool_mu); CUDA_TRY(h, g_pool.get_host(lb_tot, &q)); }
        h->host_blocks.push_back({q, lb_tot});
        h->log_h = (char*)q; h->log_cap = lb_tot;
    }
    const bool lua_fused = h->lua_fused_ok;
    if (lua_fused) {
        L(KC_FINAL, [&] { k_lua_fused<<<dim3(cdiv(h->nmax, LF_WARPS), ncore_own), 32 * LF_WARPS, lua_fused_smem(Rmax), s>>>(D, bound ? h->pack_d : nullptr); });
    } else if (h->use_wave) {
        L(KC_FINAL, [&] { k_lua_r_w<<<dim3(std::min(512, cdiv((i64)h->nmax * Rmax, 8)), ncore_own), 256, h->sm_ext, s>>>(D); });
        L(KC_FINAL, [&] { k_lua_l_w<<<dim3(std::min(512, cdiv((i64)h->nmax * Rmax, 8)), ncore_own), 256, h->sm_ext, s>>>(D); });
    } else {
        L(KC_FINAL, [&] { k_lua_r<<<dim3(cdiv((i64)h->nmax * Rmax, 128), ncore_own), 128, 0, s>>>(D); });
        L(KC_FINAL, [&] { k_lua_l<<<dim3(cdiv((i64)h->nmax * Rmax, 128), ncore_own), 128, 0, s>>>(D); });
    }
    if (persist_quad_pending) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_join[0], 0));      // the per-sweep values are part of the run
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    // Every reference driver follows dtt_dmrgg with dtt_quad of the finalised train (test_crs_ising.f90:158): with one process
    // its few small kernels are enqueued right here (after the timed region of ttc_device_ms) and ttc_quad returns the value
    // without another launch + synchronisation round.
    h->quad_cached = false;
    if (h->nproc == 1) {
        int stq = launch_quad(h, L, false, !h->quad.empty(), 1);
        if (stq) return stq;
        CUDA_TRY(h, cudaMemcpyAsync(h->sweep_h, h->plan.sweep_out, sizeof(SweepOut), cudaMemcpyDeviceToHost, s));
    }
    if (bound && !lua_fused)
        L(KC_FINAL, [&] { k_pack_all<<<dim3(std::min(256, cdiv((i64)Rmax * h->nmax * Rmax, 256)), ncore_own), 256, 0, s>>>(D, h->pack_d); });
    CUDA_TRY(h, cudaMemcpyAsync(h->log_h + lb_ctrl, D.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->log_h + lb_slog, D.slog, (size_t)(Rmax + 1) * sizeof(SweepOut), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->log_h + lb_rklog, D.rklog, (size_t)(Rmax + 1) * (d + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->log_h + lb_vlog, D.vlog, (size_t)Rmax * maxnb * P * sizeof(VisitOut), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->log_h + lb_rk, D.rk, (size_t)(d + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    CUDA_TRY(h, cudaGetLastError());
    if (h->nproc == 1) { h->quad_value = h->sweep_h->val; h->quad_cached = true; }
    // bound cores: the transfer of the packed copy into the caller's buffer starts as soon as the logs (which carry the ranks)
    // have arrived and runs beside the host's log processing below.  (Measured: starting it right behind k_lua_fused on the
    // second stream, after an early read of the ranks, is 0.07 ms SLOWER end to end at config B -- a second host wake-up and
    // the log transfers queueing behind 13 MB cost more than the closing quadrature it would hide.)
    size_t bind_tot = 0;
    bool bind_small = false;
    if (bound) {
        const int* rkd = (const int*)(h->log_h + lb_rk);
        for (int k = D.c_lo; k <= D.c_hi; ++k) bind_tot += (size_t)rkd[k - 1] * h->n[k] * rkd[k];
        bind_small = (long long)bind_tot > h->bind_cap;
        if (!bind_small && bind_tot > 0)
            CUDA_TRY(h, cudaMemcpyAsync(h->bind_pinned ? h->bind_out : h->stage_h, h->pack_d, bind_tot * sizeof(double), cudaMemcpyDeviceToHost, s));
    }

    tr.lap("sweeps+finalise");
    // ---- read the logs back and rebuild the reference's report
    Ctrl ctrl;
    std::memcpy(&ctrl, h->log_h + lb_ctrl, sizeof ctrl);
    if (ctrl.error == 3) { h->err = "multi-GPU exchange timed out waiting for a peer rank"; return TTC_ERR_COMM; }
    if (ctrl.error == 2) { h->err = "internal: the incremental quadrature saw a rank grow by more than one in a sweep"; return TTC_ERR_STATE; }
    if (ctrl.error) { h->err = "rank capacity exceeded (pass maxrank)"; return TTC_ERR_RANK; }
    it = ctrl.nsweeps;
    h->converged = (ctrl.has_accuracy && ctrl.strike >= 3) ? 1 : 0;
    {
        std::vector<SweepOut> slog(it + 1);
        std::vector<int> rklog((size_t)(it + 1) * (d + 1));
        std::vector<VisitOut> vlog((size_t)std::max(it, 1) * maxnb * P);
        if (it > 0) {
            std::memcpy(slog.data(), h->log_h + lb_slog, slog.size() * sizeof(SweepOut));
            std::memcpy(rklog.data(), h->log_h + lb_rklog, rklog.size() * sizeof(int));
            std::memcpy(vlog.data(), h->log_h + lb_vlog, (size_t)it * maxnb * P * sizeof(VisitOut));
        }
        if (persistent) {
            // the '0::' line of the initial cross (dmrgg.f90:250-300) from the record k_init_state left in slog[0]
            std::vector<int> rk1(d + 2, 1);
            const double er0 = erank(d, h->n, rk1);
            SweepOut s0;
            std::memcpy(&s0, h->log_h + lb_slog, sizeof s0);
            std::snprintf(line, sizeof line, "%3d%2s rank%5.1f time: %s n_evals: %10lld", 0, "::", er0, fmt_e(t_init, 9, 3).c_str(), (long long)s0.neval);
            std::string str0 = line;
            if (has_quad) str0 += " val " + fmt_e(s0.val, 20, 14);
            h->text += str0 + "\n";
            push_series(has_quad ? s0.val : 0.0, s0.neval, s0.amax, -1.0, er0, t_init);
            nevalall = s0.neval;
        }
        double vprev = h->s_val.empty() ? 0.0 : h->s_val[0];
        for (int sw = 1; sw <= it; ++sw) {
            const SweepOut& SO = slog[sw];
            std::vector<int> rkv(rklog.begin() + (size_t)sw * (d + 1), rklog.begin() + (size_t)(sw + 1) * (d + 1));
            rkv.push_back(1);
            // tape order of the reference: rank by rank, each rank's bonds in visit order
            for (int v = 0; v < P; ++v)
                for (int pp = 1; pp <= maxnb; ++pp) {
                    const VisitOut& O = vlog[visit_log_index(sw, pp, v)];
                    if (O.active) h->pivlog.push_back({sw, v, O.bond, O.ii, O.jj, O.kk, O.qq, O.upd, O.pivot});
                }
            double t2 = 1e-9 * (double)SO.t_ns;
            double er = erank(d, h->n, rkv);
            std::snprintf(line, sizeof line, "%3d%2s rank%5.1f time: %s n_evals: %10lld", sw, (2 - sw % 2) == 1 ? ">>" : "<<", er,
                          fmt_e(t2, 9, 3).c_str(), SO.neval);
            std::string str = line;
            if (has_quad) {
                if (h->has_tru) str += " err " + fmt_e(std::fabs(1.0 - SO.val / h->tru), 8, 3) + " val " + fmt_e(SO.val, 20, 14);
                else str += " cnv " + fmt_e(std::fabs(1.0 - SO.val / vprev), 8, 3) + " val " + fmt_e(SO.val, 20, 14);
                vprev = SO.val;
            }
            h->text += str + "\n";
            push_series(has_quad ? SO.val : 0.0, SO.neval, SO.amax, SO.pivotmax, er, t2);
            nevalall = SO.neval;
        }
        std::vector<int> rkf(d + 2, 1);
        std::memcpy(rkf.data(), h->log_h + lb_rk, (size_t)(d + 1) * sizeof(int));
        h->rk_h = rkf;
    }
    tr.lap("logs");
    h->nsweeps = it;
    float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->device_ms = ms;
    h->persistent_used = persistent ? 1 : 0;
    h->sweep_ms = 0;
    if (persistent) { float sm = 0; cudaEventElapsedTime(&sm, h->evs0, h->evs1); h->sweep_ms = sm; }
    h->neval = nevalall;
    if (bound && !bind_small) {
        CUDA_TRY(h, cudaStreamSynchronize(s));
        if (!h->bind_pinned) std::memcpy(h->bind_out, h->stage_h, bind_tot * sizeof(double));
        h->bind_filled = true; h->bind_count = bind_tot;
        tr.lap("bound cores");
    }
    h->seconds = timef();
    h->ran = true;
    if (h->verbose && multi) {            // several processes run asynchronously: the sweep lines are printed once the run is over
        std::fputs(h->text.c_str(), stdout);
        std::fflush(stdout);
    }
    if (bind_small) { h->err = "ttc_bind_cores: the bound buffer is too small for the cores of this run (ttc_cores still returns them)"; return TTC_ERR_ARG; }
    return TTC_OK;
}

}  // namespace

// =============================================================================
// C-ABI
// =============================================================================
extern "C" {

int ttc_version(void) { return 100; }

int ttc_create(ttc_handle** out, int kind, int d, const int* n, const double* par, long npar, const double* aux, long naux) {
    if (!out) { g_create_err = "ttc_create: out is NULL"; return TTC_ERR_ARG; }
    *out = nullptr;
    if (kind != TTC_ISING && kind != TTC_STDNORM && kind != TTC_MVN && kind != TTC_COSCOEF) { g_create_err = "ttc_create: unknown integrand kind"; return TTC_ERR_ARG; }
    std::vector<double> cospar;
    if (kind == TTC_COSCOEF) {      // calc_coefficient reads ind(j) - 1, not par: the "node" of mode index k is k - 1
        if (d < 2 || !n) { g_create_err = "ttc_create: need d >= 2 and n"; return TTC_ERR_ARG; }
        if (d > COS_MAXD) { g_create_err = "ttc_create: the COS coefficient sums 2^(d-1) sign vectors; d is limited to 24"; return TTC_ERR_ARG; }
        if (!aux || naux < (long)d + (long)d * d + 2) { g_create_err = "ttc_create: COS aux must hold mu(d) | sigma(d,d) | lower | upper"; return TTC_ERR_ARG; }
        int nm = 1; for (int i = 0; i < d; ++i) nm = std::max(nm, n[i]);
        cospar.resize(nm); for (int i = 0; i < nm; ++i) cospar[i] = (double)i;
        par = cospar.data(); npar = nm;
    }
    if (d < 2 || !n || !par || npar <= 0) { g_create_err = "ttc_create: need d >= 2, n and par"; return TTC_ERR_ARG; }
    for (int i = 0; i < d; ++i) if (n[i] < 1) { g_create_err = "ttc_create: mode sizes must be positive"; return TTC_ERR_ARG; }
    if (kind == TTC_ISING) {
        for (int i = 1; i < d; ++i) if (n[i] != n[0]) { g_create_err = "ttc_create: the Ising integrand needs equal mode sizes"; return TTC_ERR_ARG; }
        if (npar < 2L * n[0] + 1) { g_create_err = "ttc_create: Ising par must hold nodes(n) | weights(n) | id"; return TTC_ERR_ARG; }
        int id = (int)par[2 * n[0]];
        if (id < 1 || id > 3) { g_create_err = "unknown id"; return TTC_ERR_ARG; }
    } else {
        int nm = *std::max_element(n, n + d);
        if (npar < nm) { g_create_err = "ttc_create: par must hold the nodes"; return TTC_ERR_ARG; }
    }
    if (kind == TTC_MVN && (!aux || naux < (long)d + (long)d * d + 1)) { g_create_err = "ttc_create: MVN aux must hold mu(d) | inv_cov(d,d) | denom"; return TTC_ERR_ARG; }
    ttc_handle* h = new ttc_handle();
    h->kind = kind; h->d = d;
    h->n.assign(d + 2, 1);
    for (int i = 0; i < d; ++i) h->n[i + 1] = n[i];
    h->par.assign(par, par + npar);
    if (aux && naux > 0) h->aux.assign(aux, aux + naux);
    if (kind == TTC_ISING) h->ising_id = (int)par[2 * n[0]];
    h->P = 1; h->own = {1, d};
    std::memset(&h->plan, 0, sizeof h->plan);
    *out = h;
    return TTC_OK;
}

void ttc_destroy(ttc_handle* h) {
    if (!h) return;
    if (h->bind_out && h->bind_pinned) { cudaHostUnregister(h->bind_out); cudaGetLastError(); }
    if (h->stream || !h->allocs.empty()) free_device(h);      // (switches to the device the blocks were allocated on)
    if (h->win) { cudaSetDevice(h->device); window_teardown(h); }
    if (h->comm) { cudaSetDevice(h->device); nccl_api().CommDestroy(h->comm); h->comm = nullptr; }
    if (h->flush_d) cudaFree(h->flush_d);
    if (h->tlog_d) { cudaFree(h->tlog_d); cudaFree(h->tlog_n_d); }
    delete h;
}

const char* ttc_last_error(const ttc_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int ttc_set_device(ttc_handle* h, int dev) {
    if (!h) return TTC_ERR_ARG;
    if (dev == h->device) return TTC_OK;
    if (h->comm) { h->err = "ttc_set_device: the communicator of ttc_comm_init is bound to the current device"; return TTC_ERR_STATE; }
    // device state of an earlier run lives on the old device: release it there (free_device switches to the device the
    // blocks were allocated on), so that nothing of GPU A is ever handed to a kernel on GPU B
    if (h->stream || !h->allocs.empty()) { free_device(h); h->ran = false; }
    if (h->flush_d) { cudaFree(h->flush_d); h->flush_d = nullptr; h->flush_cap = 0; }
    if (h->tlog_d) { cudaFree(h->tlog_d); cudaFree(h->tlog_n_d); h->tlog_d = nullptr; h->tlog_n_d = nullptr; }
    h->device = dev;
    return TTC_OK;
}
int ttc_set_partition(ttc_handle* h, int nparts, const int* own) {
    if (!h || nparts < 1) return TTC_ERR_ARG;
    h->P = nparts;
    if (own) { h->own.assign(own, own + nparts + 1); h->own_given = true; }
    else { h->own_given = false; h->own.assign(nparts + 1, 0); ttc_share(1, h->d - 1, nparts, h->own.data()); }
    return TTC_OK;
}
int ttc_set_quad(ttc_handle* h, const double* quad) {
    if (h) h->quad_cached = false;
    if (!h) return TTC_ERR_ARG;
    h->quad.clear();
    if (quad) { size_t tot = 0; for (int p = 1; p <= h->d; ++p) tot += h->n[p]; h->quad.assign(quad, quad + tot); }
    return TTC_OK;
}
int ttc_set_par(ttc_handle* h, const double* par, long npar) {
    if (!h || !par) return TTC_ERR_ARG;
    if ((size_t)npar != h->par.size()) { h->err = "ttc_set_par: the parameter blob must keep its length"; return TTC_ERR_ARG; }
    if (h->kind == TTC_ISING && (int)par[2 * h->n[1]] != h->ising_id) { h->err = "ttc_set_par: the integrand id must not change"; return TTC_ERR_ARG; }
    h->par.assign(par, par + npar);
    h->par_dirty = true;
    return TTC_OK;
}
int ttc_set_tru(ttc_handle* h, int present, double tru) { if (!h) return TTC_ERR_ARG; h->has_tru = present != 0; h->tru = tru; return TTC_OK; }
int ttc_set_seed(ttc_handle* h, unsigned long long seed) { if (!h) return TTC_ERR_ARG; h->seed = seed; return TTC_OK; }
int ttc_set_uniform_callback(ttc_handle* h, ttc_uniform_cb cb, void* ctx) { if (!h) return TTC_ERR_ARG; h->ucb = cb; h->ucb_ctx = ctx; return TTC_OK; }
int ttc_set_verbose(ttc_handle* h, int v) { if (!h) return TTC_ERR_ARG; h->verbose = v; return TTC_OK; }
int ttc_set_exp_mode(ttc_handle* h, int mode) { if (!h || mode < 0 || mode > 1) return TTC_ERR_ARG; h->exp_mode = mode; return TTC_OK; }
int ttc_converged(const ttc_handle* h) { return h ? h->converged : 0; }
int ttc_set_lottery_mode(ttc_handle* h, int mode) {
    if (!h || mode < 0 || mode > 4) return TTC_ERR_ARG;
    h->force_host_lottery = (mode == 1); h->force_sync = (mode == 1 || mode == 2);
    h->force_simple = (mode == 3);      // 3: the plain (non-wavefront) support kernels, asynchronous
    h->force_split = (mode == 4);       // 4: one kernel per step of a bond visit instead of the cluster kernel
    h->setup_sig.clear();
    return TTC_OK;
}
int ttc_set_timeline(ttc_handle* h, int on) { if (!h) return TTC_ERR_ARG; h->timeline = on; return TTC_OK; }
long ttc_timeline(const ttc_handle* h, long cap, int* ids, unsigned long long* t_ns, const char** names, int names_cap) {
    if (!h || !h->tlog_d) return 0;
    int n = 0;
    cudaMemcpy(&n, h->tlog_n_d, sizeof n, cudaMemcpyDeviceToHost);
    n = std::min(n, 65536);
    std::vector<unsigned long long> buf(3 * (size_t)n);
    if (n) cudaMemcpy(buf.data(), h->tlog_d, buf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    for (long i = 0; i < std::min<long>(n, cap); ++i) {
        if (ids) ids[i] = (int)buf[3 * i];
        if (t_ns) t_ns[i] = buf[3 * i + 1];
        if (t_ns && cap >= 2 * (long)n) t_ns[n + i] = buf[3 * i + 2];      // SM cycle counters after the timestamps
    }
    const int nn = (int)(sizeof(tl_names) / sizeof(tl_names[0]));
    for (int i = 0; names && i < std::min(nn, names_cap); ++i) names[i] = tl_names[i];
    return n;
}
int ttc_set_profile(ttc_handle* h, int on) { if (!h) return TTC_ERR_ARG; h->profile = on; return TTC_OK; }

int ttc_dmrgg(ttc_handle* h, int maxrank, double accuracy, int pivoting) {
    if (!h) return TTC_ERR_ARG;
    h->err.clear();
    return run_dmrgg(h, maxrank, accuracy, pivoting);
}

int ttc_ranks(const ttc_handle* h, int* r) {
    if (!h || !r) return TTC_ERR_ARG;
    if (!h->ran) return TTC_ERR_STATE;
    for (int p = 0; p <= h->d; ++p) r[p] = h->rk_h[p];
    return TTC_OK;
}
int ttc_core(ttc_handle* h, int k, double* out) {
    if (!h || !out || k < 1 || k > h->d) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_core before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (k < h->plan.c_lo || k > h->plan.c_hi) {   // like the reference: on return each rank holds only its own cores
        h->err = "ttc_core: core " + std::to_string(k) + " is held by another rank (see ttc_core_range)";
        return TTC_ERR_STATE;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    size_t cnt = (size_t)h->rk_h[k - 1] * h->n[k] * h->rk_h[k];
    { int st = ensure_pack(h, cnt); if (st) return st; }
    k_pack_core<<<std::min(1024, cdiv((i64)cnt, 256)), 256, 0, h->stream>>>(h->plan, k, h->pack_d);
    h->launches += 1;
    CUDA_TRY(h, cudaMemcpyAsync(h->stage_h, h->pack_d, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    std::memcpy(out, h->stage_h, cnt * sizeof(double));
    return TTC_OK;
}
// all cores at once, concatenated in core order (one gather pass, one pinned device-to-host copy)
int ttc_cores(ttc_handle* h, double* out, long long cap) {
    if (!h || !out) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_cores before ttc_dmrgg"; return TTC_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int klo = h->plan.c_lo, khi = h->plan.c_hi;      // several processes: this rank's own cores only
    size_t tot = 0;
    for (int k = klo; k <= khi; ++k) tot += (size_t)h->rk_h[k - 1] * h->n[k] * h->rk_h[k];
    if ((long long)tot > cap) { h->err = "ttc_cores: output buffer too small"; return TTC_ERR_ARG; }
    if (h->bind_filled && out == h->bind_out && tot == h->bind_count) return TTC_OK;     // ttc_dmrgg already delivered them (ttc_bind_cores)
    { int st = ensure_pack(h, tot); if (st) return st; }
    size_t off = 0;
    for (int k = klo; k <= khi; ++k) {
        size_t cnt = (size_t)h->rk_h[k - 1] * h->n[k] * h->rk_h[k];
        k_pack_core<<<std::min(1024, cdiv((i64)cnt, 256)), 256, 0, h->stream>>>(h->plan, k, h->pack_d + off);
        h->launches += 1;
        off += cnt;
    }
    // device -> pinned staging in four slices; each slice is copied to the caller's (pageable) buffer by its own host
    // thread as soon as its DMA has finished, so DMA and host copies overlap and the host copy runs at memory bandwidth
    const size_t bytes = tot * sizeof(double);
    const int nsl = bytes >= ((size_t)2 << 20) ? 4 : 1;
    cudaEvent_t evs[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t b0[5];
    for (int q = 0; q <= nsl; ++q) b0[q] = (bytes * q / nsl) & ~(size_t)63;
    b0[nsl] = bytes;
    for (int q = 0; q < nsl; ++q) {
        CUDA_TRY(h, cudaMemcpyAsync((char*)h->stage_h + b0[q], (const char*)h->pack_d + b0[q], b0[q + 1] - b0[q], cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventCreateWithFlags(&evs[q], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventRecord(evs[q], h->stream));
    }
    {
        std::vector<std::thread> th;
        char* dst = (char*)out; const char* src = (const char*)h->stage_h;
        for (int q = 1; q < nsl; ++q)
            th.emplace_back([=]() { cudaEventSynchronize(evs[q]); std::memcpy(dst + b0[q], src + b0[q], b0[q + 1] - b0[q]); });
        cudaEventSynchronize(evs[0]);
        std::memcpy(dst, src, b0[1] - b0[0]);
        for (auto& t : th) t.join();
    }
    for (int q = 0; q < nsl; ++q) cudaEventDestroy(evs[q]);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return TTC_OK;
}
int ttc_bind_cores(ttc_handle* h, double* out, long long cap) {
    if (!h || (out && cap <= 0)) return TTC_ERR_ARG;
    if (h->bind_out && h->bind_pinned) { cudaHostUnregister(h->bind_out); cudaGetLastError(); }
    h->bind_out = nullptr; h->bind_cap = 0; h->bind_pinned = false; h->bind_filled = false; h->bind_count = 0;
    if (!out) return TTC_OK;
    { int st = check_device(h); if (st) return st; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    // page-lock the caller's pages so the device writes into them directly; a buffer that cannot be registered (already
    // registered by the caller, exotic mapping) goes through the pinned staging block and one host copy instead
    if (cudaHostRegister(out, (size_t)cap * sizeof(double), cudaHostRegisterPortable) == cudaSuccess) h->bind_pinned = true;
    else cudaGetLastError();
    h->bind_out = out; h->bind_cap = cap;
    return TTC_OK;
}
long long ttc_neval(const ttc_handle* h) { return h ? h->neval : 0; }
int ttc_nsweeps(const ttc_handle* h) { return h ? h->nsweeps : 0; }
double ttc_seconds(const ttc_handle* h) { return h ? h->seconds : 0; }
int ttc_sweep_series(const ttc_handle* h, int which, double* out) {
    if (!h || !out) return TTC_ERR_ARG;
    const std::vector<double>* v = nullptr;
    switch (which) {
        case 0: v = &h->s_val; break; case 1: v = &h->s_neval; break; case 2: v = &h->s_amax; break;
        case 3: v = &h->s_pivotmax; break; case 4: v = &h->s_erank; break; case 5: v = &h->s_time; break;
        default: return TTC_ERR_ARG;
    }
    std::copy(v->begin(), v->end(), out);
    return TTC_OK;
}
long ttc_pivlog_count(const ttc_handle* h) { return h ? (long)h->pivlog.size() : 0; }
int ttc_pivlog(const ttc_handle* h, int* ints, double* vals) {
    if (!h || !ints || !vals) return TTC_ERR_ARG;
    for (size_t i = 0; i < h->pivlog.size(); ++i) {
        const PivRec& r = h->pivlog[i];
        int* o = ints + 8 * i;
        o[0] = r.it; o[1] = r.vrank; o[2] = r.bond; o[3] = r.ii; o[4] = r.jj; o[5] = r.kk; o[6] = r.qq; o[7] = r.upd;
        vals[i] = r.pivot;
    }
    return TTC_OK;
}
long ttc_text(const ttc_handle* h, char* buf, long cap) {
    if (!h) return 0;
    if (buf && cap > 0) { long c = std::min<long>(cap - 1, (long)h->text.size()); std::memcpy(buf, h->text.data(), c); buf[c] = 0; }
    return (long)h->text.size();
}

int ttc_quad(ttc_handle* h, double* val) {
    if (!h || !val) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_quad before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (h->quad_cached) { *val = h->quad_value; return TTC_OK; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    Launcher L(h);
    // peer windows: the last sweep's phase-2 slots of a slower rank must have been consumed before anybody overwrites them
    // with the final chain products (inside the sweeps phase 1 plays that role); ttc_quad is rare, a barrier is cheap
    if (h->nproc > 1 && h->p2p) { int st = comm_barrier(h); if (st) { h->err = "ttc_quad: barrier failed"; return st; } }
    { int st = launch_quad(h, L, false, !h->quad.empty(), 1); if (st) return st; }   // collective when there are several processes
    CUDA_TRY(h, cudaMemcpyAsync(h->sweep_h, h->plan.sweep_out, sizeof(SweepOut), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *val = h->sweep_h->val;
    return TTC_OK;
}

// lib/quad.f90:97-131
void ttc_lgwt(int n, double* x, double* w) {
    const double tpi = 6.283185307179586476925286766559005768394338798750211641949889184615632812572417997256069650684234;
    const double small = 5 * 2.220446049250313e-16;
    const int m = (n + 1) / 2;
    for (int i = 1; i <= m; ++i) {
        double z = std::cos((tpi * (4 * i - 1)) / (8 * n + 4));
        double p1, p2, p3, pp, z1;
        do {
            p1 = 1.0; p2 = 0.0;
            for (int j = 1; j <= n; ++j) { p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
            pp = n * (z * p1 - p2) / (z * z - 1);
            z1 = z;
            z = z1 - p1 / pp;
        } while (std::fabs(z - z1) > small);
        x[i - 1] = -z; x[n - i] = z;
        w[i - 1] = 2.0 / ((1 - z * z) * pp * pp);
        w[n - i] = w[i - 1];
    }
}
// lib/default.f90:80-97
void ttc_share(int first, int last, int nproc, int* own) {
    own[0] = first;
    for (int p = 1; p <= nproc - 1; ++p) own[p] = first + (int)((double)(last - first + 1) * (double)p / nproc);
    own[nproc] = last + 1;
}
double ttc_stream_uniform(unsigned long long seed, int vrank, unsigned long long k) { return stream_uniform(seed, vrank, k); }

int ttc_fp64_peak(int device, int fma, double* tflops) {
    if (!tflops) return TTC_ERR_ARG;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return TTC_ERR_CUDA;
    cudaSetDevice(device);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    double* out = nullptr;
    if (cudaMalloc((void**)&out, (size_t)nsm * 2 * 512 * sizeof(double)) != cudaSuccess) return TTC_ERR_CUDA;
    const int iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        if (fma) k_fp64_peak<1><<<nsm * 2, 512>>>(out, iters, 0.5); else k_fp64_peak<0><<<nsm * 2, 512>>>(out, iters, 0.5);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return TTC_ERR_CUDA;
    const double flops = (double)nsm * 2 * 512 * (double)iters * 64 * 2;
    *tflops = flops / (best * 1e-3) / 1e12;
    return TTC_OK;
}

// cooperative launch of k_qr_panel on device buffers (m >= n); scratch: part[G], partw[2*G*n], head[n+2]
struct QrScratch { double* part = nullptr; double* pw = nullptr; double* head = nullptr; size_t cap_g = 0, cap_n = 0;
    double* tsqr = nullptr; size_t cap_t = 0;                       // TSQR: sign vector | per level R factors (| Q blocks)
    void release() { cudaFree(part); cudaFree(pw); cudaFree(head); cudaFree(tsqr); part = pw = head = tsqr = nullptr; cap_g = cap_n = cap_t = 0; } };
int qr_geometry(int nsm, int m, int n, int& G, int& rpb, size_t& smem) {
    const int rows_target = std::getenv("TTC_QR_ROWS") ? std::max(16, std::atoi(std::getenv("TTC_QR_ROWS"))) : 128;
    G = std::max(1, std::min(nsm, (m + rows_target - 1) / rows_target));
    while (G < nsm && ((size_t)((m + G - 1) / G) * n + n) * sizeof(double) > 200 * 1024) ++G;
    rpb = (m + G - 1) / G;
    G = (m + rpb - 1) / rpb;
    smem = ((size_t)rpb * n + n) * sizeof(double);
    return smem <= 200 * 1024 ? 0 : TTC_ERR_ARG;
}
// TSQR (ttc_qr.cuh): level-1 blocks of ~256 rows, a tree of stacked R factors with fan-in 256/n, the chain of Q slices applied to
// the first-level blocks, LAPACK's signs by Householder reconstruction.  Applies to m >= 8n, n <= 64; TTC_NO_TSQR=1 keeps k_qr_panel.
static bool tsqr_applicable(int m, int n) { return n >= 1 && n <= 64 && m >= 8 * n && !std::getenv("TTC_NO_TSQR"); }
cudaError_t tsqr_launch(cudaStream_t s, const double* da, int m, int n, double* dq, double* dr, QrScratch& sc) {
    // blocks of <= 256 rows (8 per lane, ttc_qr.cuh): first-level blocks of ~256 rows, upper nodes stack F = 256 / n (>= 4) R factors
    const int F = std::max(4, 256 / n);
    const int G1 = (m + 255) / 256;
    std::vector<int> G = {G1};
    while (G.back() > 1) G.push_back((G.back() + F - 1) / F);
    const int L = (int)G.size();                                  // levels 1 .. L (level L has one node)
    if (L > 8) return cudaErrorInvalidValue;
    size_t need = (size_t)n + 8;
    for (int l = 0; l < L; ++l) need += (size_t)G[l] * n * n + (size_t)G[l] * n + (l >= 1 ? (size_t)G[l] * F * n * n : 0);
    if (need > sc.cap_t) {
        if (sc.tsqr) cudaFree(sc.tsqr);
        sc.tsqr = nullptr; sc.cap_t = 0;
        cudaError_t e = cudaMalloc((void**)&sc.tsqr, need * sizeof(double));
        if (e != cudaSuccess) return e;
        sc.cap_t = need;
    }
    std::vector<double*> Rl(L), Ql(L, nullptr), Tl(L);
    double* p = sc.tsqr;
    double* sgn = p; p += n + 8;
    for (int l = 0; l < L; ++l) {
        Rl[l] = p; p += (size_t)G[l] * n * n;
        Tl[l] = p; p += (size_t)G[l] * n;
        if (l >= 1) { Ql[l] = p; p += (size_t)G[l] * F * n * n; }
    }
    const size_t sm_f = ((size_t)n * TQ_LDV + 2 * n) * sizeof(double);                  // V (n reflectors, ld 256) | tau | one mbarrier per reflector
    const size_t sm_l = ((size_t)n * TQ_LDV + n + (size_t)2 * n * n) * sizeof(double);  // V / Q1 | tau | M | T
    TsqrLevels LV; LV.count = L - 1; LV.first[0] = 0;
    for (int l = 1; l < L; ++l) { LV.q[l - 1] = Ql[l]; LV.tau[l - 1] = Tl[l]; LV.nsrc[l - 1] = G[l - 1]; LV.first[l] = LV.first[l - 1] + G[l]; }
    // the chain of factor launches (only R feeds the next level), then the explicit Q of all upper nodes in one launch, then the
    // first-level blocks: dorg2r + chain of slices + block product in one kernel
    auto run = [&](auto kfac, auto kformq, auto kleaf) -> cudaError_t {
        cudaError_t e2 = optin_smem(kfac, (int)sm_f);
        if (e2 == cudaSuccess) e2 = optin_smem(kformq, (int)sm_f);
        if (e2 == cudaSuccess) e2 = optin_smem(kleaf, (int)sm_l);
        if (e2 != cudaSuccess) return e2;
        kfac<<<G1, TQ_THREADS, sm_f, s>>>(da, 1, m, n, m, G1, 0, F, Rl[0], dq, m, Tl[0]);
        for (int l = 1; l < L; ++l) kfac<<<G[l], TQ_THREADS, sm_f, s>>>(Rl[l - 1], l + 1, m, n, m, G1, G[l - 1], F, Rl[l], Ql[l], m, Tl[l]);
        if (L > 1) kformq<<<LV.first[L - 1], TQ_THREADS, sm_f, s>>>(n, F, LV);
        kleaf<<<G1, TQ_THREADS, sm_l, s>>>(dq, m, n, m, G1, F, Tl[0], LV);
        return cudaSuccess;
    };
    cudaError_t e;
    switch ((n + TQ_WARPS - 1) / TQ_WARPS) {
        case 1: e = run(k_tsqr_factor<1>, k_tsqr_formq<1>, k_tsqr_leaf<1>); break;
        case 2: e = run(k_tsqr_factor<2>, k_tsqr_formq<2>, k_tsqr_leaf<2>); break;
        case 3: e = run(k_tsqr_factor<3>, k_tsqr_formq<3>, k_tsqr_leaf<3>); break;
        default: e = run(k_tsqr_factor<4>, k_tsqr_formq<4>, k_tsqr_leaf<4>); break;
    }
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(dr, Rl[L - 1], (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
    k_tsqr_sign<<<1, TSQR_THREADS, ((size_t)n * n + n) * sizeof(double), s>>>(dq, n, m, dr, sgn);
    k_tsqr_scale<<<std::min(2048, (int)(((long long)m * n + 255) / 256)), 256, 0, s>>>(dq, m, n, m, sgn);
    return cudaGetLastError();
}
cudaError_t qr_launch(cudaStream_t s, int nsm, const double* da, int m, int n, double* dq, double* dr, QrScratch& sc) {
    if (tsqr_applicable(m, n)) return tsqr_launch(s, da, m, n, dq, dr, sc);
    int G, rpb; size_t smem;
    if (qr_geometry(nsm, m, n, G, rpb, smem)) return cudaErrorInvalidValue;
    if ((size_t)G > sc.cap_g || (size_t)n > sc.cap_n) {
        sc.release();
        cudaError_t e = cudaMalloc((void**)&sc.part, (size_t)nsm * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void**)&sc.pw, (size_t)2 * nsm * std::max(n, 128) * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void**)&sc.head, (size_t)(std::max(n, 128) + 2) * sizeof(double));
        if (e != cudaSuccess) return e;
        sc.cap_g = nsm; sc.cap_n = std::max(n, 128);
    }
    cudaError_t e = optin_smem(k_qr_panel, (int)smem);
    if (e != cudaSuccess) return e;
    int lda = m;
    void* args[] = {(void*)&da, (void*)&m, (void*)&n, (void*)&lda, (void*)&dq, (void*)&dr, (void*)&sc.part, (void*)&sc.pw, (void*)&sc.head, (void*)&rpb};
    return cudaLaunchCooperativeKernel((void*)k_qr_panel, dim3(G), dim3(QR_THREADS), args, smem, s);
}

// dtt_ort (lib/tt.f90:130-198): orthogonalise the train of the last ttc_dmrgg from the left, in place on the device.
// Afterwards ttc_core / ttc_cores / ttc_quad see the orthogonalised train (ranks are unchanged: a cross result has
// r(k) <= r(k-1) n(k), the only case handled).  Single process only.
int ttc_ort(ttc_handle* h) {
    if (!h) return TTC_ERR_ARG;
    h->quad_cached = false; h->bind_filled = false;
    if (!h->ran) { h->err = "ttc_ort before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (h->nproc > 1) { h->err = "ttc_ort: not collective yet (run it on a single-process handle)"; return TTC_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
    if (!coop) { h->err = "ttc_ort: device lacks cooperative launch"; return TTC_ERR_CUDA; }
    const int d = h->d, R = h->Rmax;
    cudaStream_t s = h->stream;
    const DevPlan& D = h->plan;
    size_t maxel = 0;
    for (int k = 1; k <= d; ++k) {
        maxel = std::max(maxel, (size_t)h->rk_h[k - 1] * h->n[k] * h->rk_h[k]);
        if (k < d && (long long)h->rk_h[k - 1] * h->n[k] < h->rk_h[k]) { h->err = "ttc_ort: an unfolding with fewer rows than columns is not supported"; return TTC_ERR_ARG; }
    }
    double *ua = nullptr, *uq = nullptr, *ur = nullptr, *acc = nullptr;
    QrScratch sc;
    auto cleanup = [&]() { cudaFree(ua); cudaFree(uq); cudaFree(ur); cudaFree(acc); sc.release(); };
    cudaError_t e = cudaMalloc((void**)&ua, maxel * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&uq, maxel * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ur, (size_t)R * R * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&acc, 4 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(acc, 0, 4 * sizeof(double), s);
    std::vector<i64> coreOff(d + 2, 0);
    { i64 a = 0; for (int p = 1; p <= d; ++p) { coreOff[p] = a; a += (i64)R * h->n[p] * R; } }
    for (int k = 1; k < d && e == cudaSuccess; ++k) {
        const int r0 = h->rk_h[k - 1], nk = h->n[k], r1 = h->rk_h[k];
        const int mm = r0 * nk, nn = r1;
        const i64 kk = (i64)h->n[k + 1] * h->rk_h[k + 1];
        k_pack_core<<<std::min(1024, cdiv((i64)mm * nn, 256)), 256, 0, s>>>(D, k, ua);                 // unfolding mm x nn
        e = qr_launch(s, h->nsm, ua, mm, nn, uq, ur, sc);
        if (e != cudaSuccess) break;
        k_ort_rnorm<<<1, 256, 0, s>>>(ur, nn, acc);
        k_ort_store_q<<<std::min(1024, cdiv((i64)mm * nn, 256)), 256, 0, s>>>(uq, D.arg + coreOff[k], r0, nk, r1, R);
        k_pack_core<<<std::min(1024, cdiv((i64)nn * kk, 256)), 256, 0, s>>>(D, k + 1, ua);             // nn x kk copy of core k+1
        k_ort_apply_r<<<std::min(2048, cdiv((i64)nn * kk, 256)), 256, 0, s>>>(ur, ua, D.arg + coreOff[k + 1], nn, kk, R);
        h->launches += 6;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        const i64 cols = (i64)h->n[d] * h->rk_h[d];
        k_ort_lastnorm<<<1, 256, 0, s>>>(D.arg + coreOff[d], h->rk_h[d - 1], cols, R, acc);
        for (int k = 1; k <= d; ++k) {
            const i64 ck = (i64)h->n[k] * h->rk_h[k];
            k_ort_scale<<<std::min(1024, cdiv((i64)h->rk_h[k - 1] * ck, 256)), 256, 0, s>>>(D.arg + coreOff[k], h->rk_h[k - 1], ck, R, acc, d, k == d ? 1 : 0);
        }
        h->launches += d + 1;
        e = cudaStreamSynchronize(s);
    }
    cleanup();
    if (e != cudaSuccess) { h->err = std::string("ttc_ort: CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    return TTC_OK;
}

// chop (lib/mat.f90:433-455): number of singular values kept for relative accuracy tol and rank cap rmax
static int chop_ref(const std::vector<double>& sv, bool has_tol, double tol, bool has_rmax, int rmax) {
    int r = (int)sv.size(); double er2 = 0.0;
    if (has_rmax && rmax < r) { for (int i = rmax; i < r; ++i) er2 += sv[i] * sv[i]; r = rmax; }
    if (has_tol) {
        double ss = 0.0; for (double x : sv) ss += x * x;
        const double nrm = std::sqrt(ss), bound = tol * tol * nrm * nrm;
        double er = er2 + sv[r - 1] * sv[r - 1];
        while (er < bound && r > 1) { er2 = er; r = r - 1; er = er + sv[r - 1] * sv[r - 1]; }
    }
    return r;
}
// dtt_svd (lib/tt.f90:307-368): TT rounding of the train of the last ttc_dmrgg on the device: ttc_ort, then right to left an SVD
// of every unfolding (QR of the transpose + one-sided Jacobi on the small factor), truncated by mat.f90's chop(tol, rmax);
// U S goes into the core on the left, norms are equalised like in dtt_ort.  tol < 0 / rmax <= 0: absent.  Ranks shrink:
// ttc_ranks / ttc_core / ttc_quad see the rounded train.  Single process only; needs r <= 64.
int ttc_svd(ttc_handle* h, double tol, int rmax) {
    if (!h) return TTC_ERR_ARG;
    h->quad_cached = false; h->bind_filled = false;
    if (!h->ran) { h->err = "ttc_svd before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (h->nproc > 1) { h->err = "ttc_svd: not collective yet (run it on a single-process handle)"; return TTC_ERR_STATE; }
    for (int k = 0; k <= h->d; ++k) if (h->rk_h[k] > 64) { h->err = "ttc_svd: ranks above 64 are not supported"; return TTC_ERR_ARG; }
    int st = ttc_ort(h);
    if (st) return st;
    const int d = h->d, R = h->Rmax;
    cudaStream_t s = h->stream;
    const DevPlan& D = h->plan;
    size_t maxel = 0;
    for (int k = 1; k <= d; ++k) maxel = std::max(maxel, (size_t)h->rk_h[k - 1] * h->n[k] * h->rk_h[k]);
    double *T = nullptr, *Q = nullptr, *Rm = nullptr, *U = nullptr, *W = nullptr, *sv = nullptr, *tmp = nullptr, *acc = nullptr;
    QrScratch sc;
    auto cleanup = [&]() { cudaFree(T); cudaFree(Q); cudaFree(Rm); cudaFree(U); cudaFree(W); cudaFree(sv); cudaFree(tmp); cudaFree(acc); sc.release(); };
    cudaError_t e = cudaMalloc((void**)&T, maxel * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&Q, maxel * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&tmp, maxel * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&Rm, (size_t)R * R * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&U, (size_t)R * R * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&W, (size_t)R * R * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&sv, 2 * (size_t)R * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&acc, 4 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(acc, 0, 4 * sizeof(double), s);
    std::vector<i64> coreOff(d + 2, 0);
    { i64 a = 0; for (int p = 1; p <= d; ++p) { coreOff[p] = a; a += (i64)R * h->n[p] * R; } }
    double lognrm = 0.0;
    for (int k = d; k >= 2 && e == cudaSuccess; --k) {
        const int mm = h->rk_h[k - 1];
        const i64 nn = (i64)h->n[k] * h->rk_h[k], kk0 = (i64)h->rk_h[k - 2] * h->n[k - 1];
        if (nn < mm) { cleanup(); h->err = "ttc_svd: an unfolding with more rows than columns is not supported"; return TTC_ERR_ARG; }
        k_svd_pack_t<<<std::min(2048, cdiv((i64)mm * nn, 256)), 256, 0, s>>>(D.arg + coreOff[k], mm, nn, R, T);
        e = qr_launch(s, h->nsm, T, (int)nn, mm, Q, Rm, sc);
        if (e != cudaSuccess) break;
        const size_t smem = (2 * (size_t)mm * mm + 2 * (size_t)mm) * sizeof(double) + (2 * (size_t)mm + 4) * sizeof(int);
        if (smem > 32 * 1024) optin_smem(k_svd_small, (int)smem);
        k_svd_small<<<1, 32 * std::max(1, std::min(32, (mm + 1) / 2)), smem, s>>>(Rm, mm, U, sv, W);
        std::vector<double> svh(mm);
        e = cudaMemcpyAsync(svh.data(), sv, mm * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) break;
        const int rr = chop_ref(svh, tol >= 0, tol, rmax > 0, rmax);
        double ss = 0.0; for (int j = 0; j < rr; ++j) ss += svh[j] * svh[j];
        const double nrm = std::sqrt(ss);                              // dnrm2(rr, s, 1) over the kept values (tt.f90:333)
        std::vector<double> sn(svh.begin(), svh.begin() + rr);
        if (nrm != 0.0) { const double sc1 = 1.0 / nrm; for (double& x : sn) x = sc1 * x; lognrm = lognrm + std::log(nrm); }
        e = cudaMemcpyAsync(sv + R, sn.data(), rr * sizeof(double), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) break;
        k_svd_apply_left<<<std::min(2048, cdiv(kk0 * rr, 256)), 256, 0, s>>>(D.arg + coreOff[k - 1], kk0, h->n[k - 1], R, U, sv + R, mm, rr, tmp);
        k_svd_store_left<<<std::min(2048, cdiv(kk0 * rr, 256)), 256, 0, s>>>(tmp, D.arg + coreOff[k - 1], kk0, h->n[k - 1], R, rr);
        k_svd_apply_right<<<std::min(2048, cdiv(nn * rr, 256)), 256, 0, s>>>(Q, nn, W, mm, rr, D.arg + coreOff[k], R);
        h->rk_h[k - 1] = rr;
        e = cudaMemcpyAsync(D.rk + (k - 1), &h->rk_h[k - 1], sizeof(int), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);            // (sn and rk_h are host temporaries of this iteration)
        h->launches += 6;
    }
    if (e == cudaSuccess) {
        // first core normalised, then every core scaled by exp(lognrm / d) (tt.f90:354-364)
        e = cudaMemcpyAsync(acc, &lognrm, sizeof(double), cudaMemcpyHostToDevice, s);
        const i64 c1 = (i64)h->n[1] * h->rk_h[1];
        k_ort_lastnorm<<<1, 256, 0, s>>>(D.arg + coreOff[1], h->rk_h[0], c1, R, acc);
        for (int k = 1; k <= d; ++k) {
            const i64 ck = (i64)h->n[k] * h->rk_h[k];
            k_ort_scale<<<std::min(1024, cdiv((i64)h->rk_h[k - 1] * ck, 256)), 256, 0, s>>>(D.arg + coreOff[k], h->rk_h[k - 1], ck, R, acc, d, k == 1 ? 1 : 0);
        }
        h->launches += d + 1;
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    cleanup();
    if (e != cudaSuccess) { h->err = std::string("ttc_svd: CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    return TTC_OK;
}

// ort0_d (lib/ort.f90:17-81): thin QR of an m x n block, host buffers in and out (column-major, leading dimension m).
// q: m x n, r: n x n.  m < n follows the reference's early-return branch (:32-46).  ms: kernel time per run (CUDA events).
int ttc_qr_thin(int device, int m, int n, const double* a, double* q, double* r, int reps, double* ms) {
    if (!a || !q || !r || m < 1 || n < 1 || reps < 1) return TTC_ERR_ARG;
    if (ms) *ms = 0.0;
    if (m < n) {
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) r[i + (size_t)n * j] = (i < m) ? a[i + (size_t)m * j] : 0.0;
        for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) q[i + (size_t)m * j] = (i == j) ? 1.0 : 0.0;
        return TTC_OK;
    }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) { g_create_err = "ttc_qr_thin: no CUDA device (there is no CPU fallback)"; return TTC_ERR_CUDA; }
    cudaSetDevice(device);
    int nsm = 0, coop = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (!coop) { g_create_err = "ttc_qr_thin: device lacks cooperative launch"; return TTC_ERR_CUDA; }
    int G, rpb; size_t smem;
    if (qr_geometry(nsm, m, n, G, rpb, smem)) { g_create_err = "ttc_qr_thin: the block does not fit the shared memory of one cooperative launch"; return TTC_ERR_ARG; }
    double *da = nullptr, *dq = nullptr, *dr = nullptr;
    QrScratch sc;
    auto cleanup = [&]() { cudaFree(da); cudaFree(dq); cudaFree(dr); sc.release(); };
    cudaError_t e = cudaMalloc((void**)&da, (size_t)m * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dq, (size_t)m * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dr, (size_t)n * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(da, a, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    for (int rep = 0; rep <= reps && e == cudaSuccess; ++rep) {      // run 0 is the warm-up
        if (rep == 1) cudaEventRecord(ev0);
        e = qr_launch(nullptr, nsm, da, m, n, dq, dr, sc);
    }
    cudaEventRecord(ev1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(q, dq, (size_t)m * n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(r, dr, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost);
    float t = 0; cudaEventElapsedTime(&t, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    cleanup();
    if (e != cudaSuccess) { g_create_err = std::string("ttc_qr_thin: CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    if (ms) *ms = t / reps;
    return TTC_OK;
}

// ztt_quad (lib/dmrgg.f90:1418-1523) for `nsets` complex rank-1 weight tensors at once (test_crs_chf.f90:153-168 loops over
// 32 of them).  wre / wim: [nsets][n(1)+...+n(d)]; out_re / out_im: [nsets].  Host buffers; single process only.
int ttc_quad_complex(ttc_handle* h, int nsets, const double* wre, const double* wim, double* out_re, double* out_im) {
    if (!h || nsets < 1 || !wre || !wim || !out_re || !out_im) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_quad_complex before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (h->nproc > 1) { h->err = "ttc_quad_complex: not collective yet (run it on a single-process handle)"; return TTC_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    size_t wlen = 0;
    for (int p = 1; p <= h->d; ++p) wlen += h->n[p];
    double *dw = nullptr, *dout = nullptr;
    const size_t wtot = (size_t)nsets * wlen;
    cudaError_t e = cudaMalloc((void**)&dw, 2 * wtot * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, 2 * (size_t)nsets * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dw, wre, wtot * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dw + wtot, wim, wtot * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const size_t smem = (4 * (size_t)h->Rmax + 2 * (size_t)h->nmax) * sizeof(double);
        k_zquad<<<nsets, 256, smem, h->stream>>>(h->plan, nsets, dw, dw + wtot, dout, dout + nsets, (long long)wlen);
        h->launches += 1;
        e = cudaGetLastError();
    }
    std::vector<double> out(2 * (size_t)nsets);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out.data(), dout, out.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dw); cudaFree(dout);
    if (e != cudaSuccess) { h->err = std::string("ttc_quad_complex: CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    for (int q = 0; q < nsets; ++q) { out_re[q] = out[q]; out_im[q] = out[nsets + q]; }
    return TTC_OK;
}

// dtt_ijk (lib/tt.f90:630-652) for `count` multi-indices at once: ind [count][d], 1-based; values [count].  Host buffers.
int ttc_values(ttc_handle* h, long long count, const int* ind, double* values) {
    if (!h || count < 1 || !ind || !values) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_values before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (h->nproc > 1) { h->err = "ttc_values: not collective yet (run it on a single-process handle)"; return TTC_ERR_STATE; }
    for (long long x = 0; x < count * h->d; ++x)
        if (ind[x] < 1 || ind[x] > h->n[1 + x % h->d]) { h->err = "ttc_values: index out of range"; return TTC_ERR_ARG; }   // dtt_ijk returns -3
    CUDA_TRY(h, cudaSetDevice(h->device));
    int* di = nullptr; double* dv = nullptr;
    cudaError_t e = cudaMalloc((void**)&di, (size_t)count * h->d * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dv, (size_t)count * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(di, ind, (size_t)count * h->d * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const int nw = 8;
        const size_t smem = (size_t)nw * 2 * h->Rmax * sizeof(double);
        const int grid = (int)std::min<long long>((count + nw - 1) / nw, (long long)h->nsm * 8);
        k_tt_values<<<grid, 32 * nw, smem, h->stream>>>(h->plan, count, di, dv);
        h->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, dv, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(di); cudaFree(dv);
    if (e != cudaSuccess) { h->err = std::string("ttc_values: CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    return TTC_OK;
}
// dtt_accchk (lib/dmrgg.f90:1081-1166): Monte-Carlo accuracy check of the train against the integrand at nlot random entries.
// out[0..3] = einf, efro, ainf, afro; pivot (d ints, may be NULL) = the multi-index of the largest error.
int ttc_accchk(ttc_handle* h, long long nlot, unsigned long long seed, double* out, int* pivot) {
    if (!h || nlot < 1 || !out) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_accchk before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (h->nproc > 1) { h->err = "ttc_accchk: not collective yet (run it on a single-process handle)"; return TTC_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int nw = 8;
    const int grid = (int)std::min<long long>((nlot + nw - 1) / nw, (long long)h->nsm * 4);
    double* dpart = nullptr; long long* darg = nullptr;
    cudaError_t e = cudaMalloc((void**)&dpart, (size_t)grid * 4 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&darg, (size_t)grid * sizeof(long long));
    const size_t smem = aux_smem(h) + (size_t)nw * (2 * h->Rmax + h->d) * sizeof(double);
    if (e == cudaSuccess) {
        KIND_SWITCH(h->kind,
            if (smem > 32 * 1024) optin_smem(k_accchk<K>, (int)smem);
            k_accchk<K><<<grid, 32 * nw, smem, h->stream>>>(h->plan, nlot, seed, dpart, darg));
        h->launches += 1;
        e = cudaGetLastError();
    }
    std::vector<double> part((size_t)grid * 4); std::vector<long long> argp(grid);
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), dpart, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(argp.data(), darg, argp.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dpart); cudaFree(darg);
    if (e != cudaSuccess) { h->err = std::string("ttc_accchk: CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    double einf = -1.0, efro = 0.0, ainf = 0.0, afro = 0.0; long long ex = 0;
    for (int g = 0; g < grid; ++g) {
        if (einf < part[4 * g] || (einf == part[4 * g] && argp[g] < ex)) { einf = part[4 * g]; ex = argp[g]; }
        efro += part[4 * g + 1]; ainf = std::max(ainf, part[4 * g + 2]); afro += part[4 * g + 3];
    }
    out[0] = std::max(einf, 0.0); out[1] = std::sqrt(efro); out[2] = ainf; out[3] = std::sqrt(afro);
    if (pivot)
        for (int p = 0; p < h->d; ++p) {
            const double u = stream_uniform(seed, 0x7fffffff, (unsigned long long)ex * h->d + p);
            int v = (int)(u * h->n[p + 1]) + 1; if (v > h->n[p + 1]) v = h->n[p + 1];
            pivot[p] = v;
        }
    return TTC_OK;
}

// ---- TT files in the reference's stream format (lib/ttio.f90:10-17 header, :29-108 dtt_write, :196-296 dtt_read) ----
// layout (little-endian, no record markers): 'TT      ' | ver(2) = 1,0 | inf(4) = tt_size,0,0,0 | comment(64) | i(8) with
// i(1) = l, i(2) = m  [128 bytes]  | l, m | n(l:m) | r(l-1:m)  [int32]  | all cores concatenated [float64]
namespace {
struct TtHead { char txt[8]; int32_t ver[2]; int32_t inf[4]; char comment[64]; int32_t i[8]; };
static_assert(sizeof(TtHead) == 128, "tthead of ttio.f90 is 128 bytes");
}
int ttc_tt_write(const char* path, int l, int m, const int* n, const int* r, const double* cores) {
    if (!path || !n || !r || !cores || m < l || l < 1) { g_create_err = "ttc_tt_write: bad arguments"; return TTC_ERR_ARG; }
    const int d = m - l + 1;
    size_t tot = 0;
    for (int k = 0; k < d; ++k) tot += (size_t)r[k] * n[k] * r[k + 1];
    FILE* f = std::fopen(path, "wb");
    if (!f) { g_create_err = std::string("dtt_write: error opening file: ") + path; return TTC_ERR_ARG; }
    TtHead hd;
    std::memcpy(hd.txt, "TT      ", 8);
    hd.ver[0] = 1; hd.ver[1] = 0;
    hd.inf[0] = 2048; hd.inf[1] = hd.inf[2] = hd.inf[3] = 0;      // tt_size (lib/tt.f90:16)
    std::memset(hd.comment, ' ', sizeof hd.comment);
    std::memset(hd.i, 0, sizeof hd.i);
    hd.i[0] = l; hd.i[1] = m;
    std::vector<int32_t> ints;
    ints.push_back(l); ints.push_back(m);
    for (int k = 0; k < d; ++k) ints.push_back(n[k]);
    for (int k = 0; k <= d; ++k) ints.push_back(r[k]);
    bool ok = std::fwrite(&hd, sizeof hd, 1, f) == 1 && std::fwrite(ints.data(), sizeof(int32_t), ints.size(), f) == ints.size() &&
              std::fwrite(cores, sizeof(double), tot, f) == tot;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { g_create_err = "dtt_write: error writing file"; return TTC_ERR_ARG; }
    return TTC_OK;
}
// header of a TT file: l, m, then n(l:m) and r(l-1:m) if the caller's arrays hold cap / cap + 1 entries; *ncore = doubles of the cores
int ttc_tt_read_header(const char* path, int* l, int* m, int* n, int* r, int cap, long long* ncore) {
    if (!path || !l || !m) { g_create_err = "ttc_tt_read_header: bad arguments"; return TTC_ERR_ARG; }
    FILE* f = std::fopen(path, "rb");
    if (!f) { g_create_err = std::string("dtt_read: file not exist: ") + path; return TTC_ERR_ARG; }
    TtHead hd; int32_t lm[2];
    bool ok = std::fread(&hd, sizeof hd, 1, f) == 1;
    if (ok && !(hd.txt[0] == 'T' && hd.txt[1] == 'T')) { std::fclose(f); g_create_err = "dtt_read: not TT header in file"; return TTC_ERR_ARG; }
    if (ok && hd.ver[0] != 1) { std::fclose(f); g_create_err = "dtt_read: not correct version of TT file"; return TTC_ERR_ARG; }
    ok = ok && std::fread(lm, sizeof(int32_t), 2, f) == 2;
    if (!ok || lm[1] < lm[0] || lm[0] < 1 || lm[1] - lm[0] + 1 > 2048) { std::fclose(f); g_create_err = "dtt_read: error reading header / lm"; return TTC_ERR_ARG; }
    *l = lm[0]; *m = lm[1];
    const int d = lm[1] - lm[0] + 1;
    std::vector<int32_t> ints(2 * (size_t)d + 1);
    ok = std::fread(ints.data(), sizeof(int32_t), ints.size(), f) == ints.size();
    std::fclose(f);
    if (!ok) { g_create_err = "dtt_read: error reading nr"; return TTC_ERR_ARG; }
    long long tot = 0;
    for (int k = 0; k < d; ++k) tot += (long long)ints[d + k] * ints[k] * ints[d + k + 1];
    if (ncore) *ncore = tot;
    if (n && r) {
        if (cap < d) { g_create_err = "ttc_tt_read_header: arrays too small"; return TTC_ERR_ARG; }
        for (int k = 0; k < d; ++k) n[k] = ints[k];
        for (int k = 0; k <= d; ++k) r[k] = ints[d + k];
    }
    return TTC_OK;
}
int ttc_tt_read_cores(const char* path, double* cores, long long cap) {
    int l, m; long long tot = 0;
    int st = ttc_tt_read_header(path, &l, &m, nullptr, nullptr, 0, &tot);
    if (st) return st;
    if (!cores || cap < tot) { g_create_err = "ttc_tt_read_cores: output buffer too small"; return TTC_ERR_ARG; }
    FILE* f = std::fopen(path, "rb");
    if (!f) { g_create_err = "dtt_read: error opening file"; return TTC_ERR_ARG; }
    const long off = (long)sizeof(TtHead) + 4L * (2 + 2 * (m - l + 1) + 1);
    bool ok = std::fseek(f, off, SEEK_SET) == 0 && std::fread(cores, sizeof(double), (size_t)tot, f) == (size_t)tot;
    std::fclose(f);
    if (!ok) { g_create_err = "dtt_read: error reading cores"; return TTC_ERR_ARG; }
    return TTC_OK;
}
// dtt_write(arg, fnam) for the train held by the handle (the cores this rank holds)
int ttc_write(ttc_handle* h, const char* path) {
    if (!h || !path) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_write before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (h->nproc > 1) { h->err = "ttc_write: gather the cores first (each rank holds its own block)"; return TTC_ERR_STATE; }
    size_t tot = 0;
    for (int k = 1; k <= h->d; ++k) tot += (size_t)h->rk_h[k - 1] * h->n[k] * h->rk_h[k];
    std::vector<double> buf(tot);
    int st = ttc_cores(h, buf.data(), (long long)tot);
    if (st) return st;
    st = ttc_tt_write(path, 1, h->d, h->n.data() + 1, h->rk_h.data(), buf.data());
    if (st) h->err = g_create_err;
    return st;
}

int ttc_l2_flush(ttc_handle* h, long long bytes) {
    if (!h || bytes <= 0) return TTC_ERR_ARG;
    int st = check_device(h);
    if (st) return st;
    if ((size_t)bytes > h->flush_cap) {
        if (h->flush_d) cudaFree(h->flush_d);
        h->flush_d = nullptr; h->flush_cap = 0;
        CUDA_TRY(h, cudaMalloc(&h->flush_d, (size_t)bytes));
        h->flush_cap = (size_t)bytes;
    }
    CUDA_TRY(h, cudaMemset(h->flush_d, 1, (size_t)bytes));
    CUDA_TRY(h, cudaDeviceSynchronize());
    return TTC_OK;
}

// host execution of the closed-form lottery used by the device (unit-test hook: compared against the literal rnd.f90 loop)
int ttc_lottery_closed_form(int m, const int* zeros_sorted_distinct, int nz, const double* u, int count, int* cells) {
    if (m < 1 || nz < 0 || nz >= m || !u || !cells) return TTC_ERR_ARG;
    std::vector<LotSeg> seg(MAXSEG);
    int ns = build_segments(m - nz, seg.data());
    if (ns >= MAXSEG) return TTC_ERR_ARG;
    for (int x = 0; x < count; ++x) cells[x] = lot_draw(seg.data(), ns, m - nz, m, zeros_sorted_distinct, nz, u[x]);
    return TTC_OK;
}

// the table-free draw of the cluster kernel (lot_draw_fast), executed on the host (test hook)
int ttc_lottery_fast(int m, const int* zeros_sorted_distinct, int nz, const double* u, int count, int* cells) {
    if (m < 1 || nz < 0 || nz >= m || !u || !cells) return TTC_ERR_ARG;
    for (int x = 0; x < count; ++x) cells[x] = lot_draw_fast(m - nz, m, zeros_sorted_distinct, nz, u[x]);
    return TTC_OK;
}

long long ttc_launch_count(const ttc_handle* h) { return h ? h->launches : 0; }
double ttc_device_ms(const ttc_handle* h) { return h ? h->device_ms : 0; }
double ttc_sweep_kernel_ms(const ttc_handle* h) { return h ? h->sweep_ms : 0; }
int ttc_sweep_geometry(const ttc_handle* h, int* cluster, int* threads) {
    if (!h) return 0;
    if (cluster) *cluster = h->sweep_cluster;
    if (threads) *threads = h->sweep_threads;
    return h->persistent_used;
}
int ttc_profile(const ttc_handle* h, int cap, const char** names, long long* launches, double* ms) {
    if (!h) return 0;
    int c = std::min(cap, (int)KC_COUNT);
    for (int i = 0; i < c; ++i) { if (names) names[i] = kclass_names[i]; if (launches) launches[i] = h->kc_launch[i]; if (ms) ms[i] = h->kc_ms[i]; }
    return KC_COUNT;
}

int ttc_superblock_probe_ex(ttc_handle* h, int bond, int store, int reps, int variant, long long* out_idx, double* out_val, double* ms, long long* count);
int ttc_superblock_probe(ttc_handle* h, int bond, int store, int reps, long long* out_idx, double* out_val, double* ms, long long* count) {
    return ttc_superblock_probe_ex(h, bond, store, reps, 0, out_idx, out_val, ms, count);
}
// variant 0: tiled kernel, reference arithmetic (what ttc_dmrgg runs); 1: the plain one-thread-per-element kernel;
// 2: tiled kernel with the residual update contracted into DFMA (not bit-exact; FP64 ceiling measurement only)
int ttc_superblock_probe_ex(ttc_handle* h, int bond, int store, int reps, int variant, long long* out_idx, double* out_val, double* ms, long long* count) {
    if (!h || bond < 1 || bond > h->d - 1 || reps < 1 || variant < 0 || variant > 3) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_superblock_probe before ttc_dmrgg"; return TTC_ERR_STATE; }
    if (bond < h->own[h->plan.v0] || bond >= h->own[h->plan.v0 + h->plan.nv]) { h->err = "ttc_superblock_probe: bond belongs to another rank"; return TTC_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const DevPlan& D = h->plan;
    cudaStream_t s = h->stream;
    const i64 tot = (i64)h->rk_h[bond - 1] * h->n[bond] * h->n[bond + 1] * h->rk_h[bond + 1];
    if (count) *count = tot;
    double* a_out = nullptr;
    if (store) CUDA_TRY(h, cudaMalloc((void**)&a_out, (size_t)tot * sizeof(double)));
    Partial* pout = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&pout, 2 * sizeof(Partial)));
    const int TB = threads_for(h);
    // persistent-style grid: a few CTAs per SM, grid-stride over the superblock
    const int Gs = (int)std::min<i64>(GMAX, std::max<i64>(1, std::min<i64>((tot + TB - 1) / TB, (i64)h->nsm * (h->kind == TTC_MVN ? 8 : 6))));
    const size_t smA = aux_smem(h);
    if (variant != 1 && !h->sbt_ok) { cudaFree(pout); if (a_out) cudaFree(a_out); h->err = "tiled superblock kernel unavailable for this shape"; return TTC_ERR_STATE; }
    if (variant == 3 && !h->sbm_ok) { cudaFree(pout); if (a_out) cudaFree(a_out); h->err = "the DMMA superblock variant needs maxrank % 4 == 0 and its tile in shared memory"; return TTC_ERR_STATE; }
    if (variant >= 2 && store) { cudaFree(pout); if (a_out) cudaFree(a_out); h->err = "the DFMA variant has no stored form"; return TTC_ERR_ARG; }
    const int m1 = h->rk_h[bond - 1] * h->n[bond], nc = h->n[bond + 1] * h->rk_h[bond + 1];
    const int nrb = cdiv(m1, SB_TM);
    const int nsp = std::max(1, std::min({(SB_MINB * h->nsm) / std::max(1, nrb), cdiv(nc, SB_TN), GMAX / nrb}));   // whole waves: SB_MINB CTAs per SM
    auto launch = [&]() {
        if (variant == 1) {
            if (store) { KIND_SWITCH(h->kind, k_superblock<K, 1><<<Gs, TB, h->sm_sb, s>>>(D, 1, 1, bond, 0, a_out, pout)); }
            else       { KIND_SWITCH(h->kind, k_superblock<K, 0><<<Gs, TB, h->sm_sb, s>>>(D, 1, 1, bond, 0, nullptr, pout)); }
        } else if (variant == 3) {
            const int nsp1 = std::max(1, std::min({h->nsm / std::max(1, nrb), cdiv(nc, SB_TN), GMAX / nrb}));      // one CTA per SM (the parked tile)
            KIND_SWITCH(h->kind, k_superblock_t<K, 0, 2><<<dim3(nrb, nsp1, 1), SB_TM, h->sm_sbm, s>>>(D, 1, 1, bond, 0, nullptr, pout));
        } else if (variant == 2) {
            KIND_SWITCH(h->kind, k_superblock_t<K, 0, 1><<<dim3(nrb, nsp, 1), SB_TM, h->sm_sbt, s>>>(D, 1, 1, bond, 0, nullptr, pout));
        } else {
            if (store) { KIND_SWITCH(h->kind, k_superblock_t<K, 1, 0><<<dim3(nrb, nsp, 1), SB_TM, h->sm_sbt, s>>>(D, 1, 1, bond, 0, a_out, pout)); }
            else       { KIND_SWITCH(h->kind, k_superblock_t<K, 0, 0><<<dim3(nrb, nsp, 1), SB_TM, h->sm_sbt, s>>>(D, 1, 1, bond, 0, nullptr, pout)); }
        }
        h->launches += 1;
    };
    launch();   // warm-up
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, s);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(b, s);
    Partial hp[2];
    cudaError_t e = cudaMemcpyAsync(hp, pout, sizeof hp, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    float t = 0; cudaEventElapsedTime(&t, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(pout); if (a_out) cudaFree(a_out);
    if (e != cudaSuccess) { h->err = std::string("CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    if (ms) *ms = t / reps;
    if (out_idx) { out_idx[0] = hp[0].idx; out_idx[1] = hp[1].idx; }
    if (out_val) { out_val[0] = hp[0].val; out_val[1] = hp[1].val; }
    return TTC_OK;
}

int ttc_fiber_probe(ttc_handle* h, int bond, int isrow, int ii, int jj, int kk, int qq, double* fiber, double* resid, int reps, double* ms) {
    if (!h || bond < 1 || bond > h->d - 1 || reps < 1) return TTC_ERR_ARG;
    if (!h->ran) { h->err = "ttc_fiber_probe before ttc_dmrgg"; return TTC_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const DevPlan& D = h->plan;
    cudaStream_t s = h->stream;
    int v = 0;
    while (v < h->P - 1 && bond >= h->own[v + 1]) ++v;
    if (v < h->plan.v0 || v >= h->plan.v0 + h->plan.nv) { h->err = "ttc_fiber_probe: bond belongs to another rank"; return TTC_ERR_STATE; }
    const int pp = bond - h->own[v] + 1;
    CUDA_TRY(h, cudaMemcpyAsync(D.rks, D.rk, (size_t)(h->d + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s));   // rks := rk
    VState S;
    CUDA_TRY(h, cudaMemcpyAsync(&S, D.st + v, sizeof S, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    VState S0 = S;
    S.ii = ii; S.jj = jj; S.kk = kk; S.qq = qq; S.done = 0;
    CUDA_TRY(h, cudaMemcpyAsync(D.st + v, &S, sizeof S, cudaMemcpyHostToDevice, s));
    const int TB = threads_for(h);
    const int cnt = isrow ? h->n[bond + 1] * h->rk_h[bond + 1] : h->rk_h[bond - 1] * h->n[bond];
    const int G = std::min(GMAX, cdiv(cnt, TB));
    const size_t smF = h->sm_fiber;
    // only virtual rank v must run: launch a 1-wide grid in y and shift the plan so blockIdx.y = 0 maps to v
    DevPlan Dv = D;
    Dv.v0 = 0; Dv.nv = 1; Dv.own = D.own + v; Dv.P = 1; Dv.st = D.st + v; Dv.part = D.part + (size_t)v * 2 * GMAX; Dv.tickets = D.tickets + v;
    Dv.acol1 = D.acol1 + (size_t)v * h->Rmax * h->nmax; Dv.bcol1 = D.bcol1 + (size_t)v * h->Rmax * h->nmax;
    Dv.arow1 = D.arow1 + (size_t)v * h->Rmax * h->nmax; Dv.brow1 = D.brow1 + (size_t)v * h->Rmax * h->nmax;
    auto launch = [&]() {
        if (isrow) { KIND_SWITCH(h->kind, k_fiber<K, 1><<<dim3(G, 1), TB, smF, s>>>(Dv, 1, pp, 1)); }
        else       { KIND_SWITCH(h->kind, k_fiber<K, 0><<<dim3(G, 1), TB, smF, s>>>(Dv, 1, pp, 1)); }
        h->launches += 1;
    };
    launch();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, s);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(b, s);
    cudaError_t e = cudaSuccess;
    if (fiber) e = cudaMemcpyAsync(fiber, isrow ? Dv.arow1 : Dv.acol1, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && resid) e = cudaMemcpyAsync(resid, isrow ? Dv.brow1 : Dv.bcol1, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(D.st + v, &S0, sizeof S0, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    float t = 0; cudaEventElapsedTime(&t, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (e != cudaSuccess) { h->err = std::string("CUDA error: ") + cudaGetErrorString(e); return TTC_ERR_CUDA; }
    if (ms) *ms = t / reps;
    return TTC_OK;
}

int ttc_core_range(const ttc_handle* h, int* first, int* last) {
    if (!h || !first || !last) return TTC_ERR_ARG;
    if (!h->ran) return TTC_ERR_STATE;
    *first = h->plan.c_lo; *last = h->plan.c_hi;
    return TTC_OK;
}

// ---- multi-GPU: one process per GPU (NCCL bound at run time, ttc_nccl.hpp)
int ttc_comm_unique_id(void* id128) {
    if (!id128) return TTC_ERR_ARG;
    NcclApi& N = nccl_api();
    if (!N.load()) { g_create_err = N.err; return TTC_ERR_COMM; }
    NcclUniqueId id;
    int e = N.GetUniqueId(&id);
    if (e != 0) { g_create_err = std::string("ncclGetUniqueId: ") + N.GetErrorString(e); return TTC_ERR_COMM; }
    std::memcpy(id128, &id, sizeof id);
    return TTC_OK;
}
int ttc_comm_init(ttc_handle* h, int nranks, int rank, const void* id128) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { if (h) h->err = "ttc_comm_init: bad arguments"; return TTC_ERR_ARG; }
    if (h->comm) { h->err = "ttc_comm_init: communicator already initialised"; return TTC_ERR_STATE; }
    int st = check_device(h);          // no CUDA device -> TTC_ERR_CUDA: there is no CPU fallback
    if (st) return st;
    NcclApi& N = nccl_api();
    if (!N.load()) { h->err = N.err; return TTC_ERR_COMM; }
    NcclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    NCCL_TRY(h, N.CommInitRank(&h->comm, nranks, id, rank));
    h->nproc = nranks; h->prank = rank;
    h->setup_sig.clear();
    return TTC_OK;
}
int ttc_comm_rank(const ttc_handle* h, int* nranks, int* rank) {
    if (!h) return TTC_ERR_ARG;
    if (nranks) *nranks = h->nproc;
    if (rank) *rank = h->prank;
    return TTC_OK;
}

}  // extern "C"
