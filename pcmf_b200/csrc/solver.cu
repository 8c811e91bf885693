// libpcmf_b200: host-side engine + C ABI (include/pcmf_b200.h).
//
// Mirrors, on the host, the reference's wrapper (wrapper_gap_factor.h:164-393, wrapper_sparse_gap_factor.h:176-435:
// argument checks, init dispatch, optimize, order, output) and iteration driver (algorithm.h:158-210, :412-450);
// everything per-iteration runs in the sm_100a kernels of kernels.cuh.  There is no CPU compute path.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <unordered_map>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "../../include/pcmf_b200_dev.h"
#define PCMF_COMMON_KERNELS 1
#include "launch.cuh"

using namespace pcmf;

namespace {

struct Err : std::runtime_error {
    int code;
    Err(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Err(code, buf);
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) fail(PCMF_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                                    cudaGetErrorString(e_));                                             \
    } while (0)

// PCMF_TIMING=1 prints the wall-clock of the set-up stages (synchronising the stream at each mark)
struct StageTimer {
    cudaStream_t st; bool on; double t;
    explicit StageTimer(cudaStream_t s) : st(s), on(getenv("PCMF_TIMING") != nullptr), t(0) { if (on) { cudaStreamSynchronize(st); t = mark_now(); } }
    static double mark_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const double n = mark_now();
        fprintf(stderr, "[pcmf timing] %-28s %8.1f ms\n", what, (n - t) * 1e3);
        t = n;
    }
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Device memory goes through a process-wide cache of released blocks: what one fit releases is handed to the next one
// (an R session calls pCMF() many times) instead of going back to the driver.  Measured at C4: after a cudaFree of the
// ~60 GB of a finished fit, the cudaMalloc calls of the next pcmf_create were 3-20x slower than in a fresh process
// (create 2.0 s against 0.5 s), and the stream-ordered pool (cudaMallocAsync) maps fresh memory at ~32 ms per GB.  A block
// is reused for a request between 80 % and 100 % of its size; pcmf_release_cached_memory() returns everything.
struct BlockCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_blocks;     // (device, bytes) -> pointer
    std::unordered_map<void *, std::pair<int, size_t>> live;        // pointer -> (device, bytes)
    double miss_seconds = 0; size_t misses = 0, miss_bytes = 0;     // cudaMalloc calls that the cache could not serve (PCMF_TIMING)
    void *get(size_t bytes) {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_blocks.lower_bound({dev, bytes});
            if (it != free_blocks.end() && it->first.first == dev && it->first.second <= bytes + bytes / 4 + 4096) {
                void *p = it->second;
                live[p] = it->first;
                free_blocks.erase(it);
                return p;
            }
        }
        void *p = nullptr;
        const double t0 = now_s();
        cudaError_t e = cudaMalloc(&p, bytes);
        { std::lock_guard<std::mutex> lk(mu); miss_seconds += now_s() - t0; ++misses; miss_bytes += bytes; }
        if (now_s() - t0 > 5e-3 && getenv("PCMF_TIMING"))
            fprintf(stderr, "[pcmf timing] cudaMalloc of %.1f MB took %.1f ms (%zu cached blocks)\n", bytes / 1e6, (now_s() - t0) * 1e3, free_blocks.size());
        if (e != cudaSuccess) {           // out of memory: give the cached blocks back and try once more
            cudaGetLastError();
            release_all();
            CK(cudaMalloc(&p, bytes));
        }
        std::lock_guard<std::mutex> lk(mu);
        live[p] = {dev, bytes};
        return p;
    }
    bool has_block(size_t bytes) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        std::lock_guard<std::mutex> lk(mu);
        auto it = free_blocks.lower_bound({dev, bytes});
        return it != free_blocks.end() && it->first.first == dev && it->first.second <= bytes + bytes / 4 + 4096;
    }
    void put(void *p) {
        if (!p) return;
        std::pair<int, size_t> key{-1, 0};
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = live.find(p);
            if (it != live.end()) key = it->second;
        }
        // like cudaFree: the block is idle on every stream of ITS device before anyone can reuse it
        int cur = 0;
        cudaGetDevice(&cur);
        if (key.first >= 0 && key.first != cur) cudaSetDevice(key.first);
        cudaDeviceSynchronize();
        if (key.first < 0) cudaFree(p);
        if (key.first >= 0 && key.first != cur) cudaSetDevice(cur);
        if (key.first < 0) return;
        std::lock_guard<std::mutex> lk(mu);
        free_blocks.insert({key, p});
        live.erase(p);
    }
    void release_all() {
        std::lock_guard<std::mutex> lk(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto &kv : free_blocks) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
        free_blocks.clear();
        cudaSetDevice(cur);
    }
};
BlockCache &block_cache() { static BlockCache c; return c; }

// Device -> user memory through two pinned staging chunks: the DMA of chunk c+1 overlaps the host copy of chunk c, and
// the host copy (which pays the first-touch page faults of the caller's freshly allocated result arrays) runs on several
// threads.  Measured need: at C4 the result is ~1 GB and a plain cudaMemcpy into pageable memory ran at ~4 GB/s.
struct HostStager {
    static constexpr size_t CH = (size_t)32 << 20;
    // staging buffers and events belong to the device that was current when they were created: one set per device
    // (a multi-device process -- pcmf_options.ngpus, or an R session that changes options(pCMF.device) -- uses several)
    struct PerDev { std::mutex mu; void *buf[2] = {nullptr, nullptr}; cudaEvent_t ev[2]; };
    std::mutex map_mu;
    std::map<int, PerDev> per;
    int nthreads() const { return (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency())); }
    static void par_for(size_t bytes, int T, const std::function<void(size_t, size_t)> &f) {
        size_t per = ((bytes + T - 1) / T + 4095) & ~(size_t)4095;
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) {
            size_t lo = per * t, hi = std::min(bytes, lo + per);
            if (lo < hi) th.emplace_back(f, lo, hi);
        }
        f(0, std::min(bytes, per));
        for (auto &x : th) x.join();
    }
    void d2h(void *dst, const void *src, size_t bytes, cudaStream_t st) {
        if (bytes < ((size_t)4 << 20)) {
            CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            return;
        }
        int dev = 0;
        CK(cudaGetDevice(&dev));
        PerDev *pd;
        { std::lock_guard<std::mutex> lk(map_mu); pd = &per[dev]; }
        std::lock_guard<std::mutex> lk(pd->mu);
        void **buf = pd->buf; cudaEvent_t *ev = pd->ev;
        if (!buf[0]) {
            for (int i = 0; i < 2; ++i) { CK(cudaHostAlloc(&buf[i], CH, cudaHostAllocPortable)); CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); }
        }
        const int T = nthreads();
        const size_t nch = (bytes + CH - 1) / CH;
        auto drain = [&](size_t c) {
            CK(cudaEventSynchronize(ev[c & 1]));
            const size_t off = c * CH, len = std::min(CH, bytes - off);
            char *d = static_cast<char *>(dst) + off; const char *b = static_cast<const char *>(buf[c & 1]);
            par_for(len, T, [=](size_t lo, size_t hi) { std::memcpy(d + lo, b + lo, hi - lo); });
        };
        for (size_t c = 0; c < nch; ++c) {
            const size_t off = c * CH, len = std::min(CH, bytes - off);
            CK(cudaMemcpyAsync(buf[c & 1], static_cast<const char *>(src) + off, len, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ev[c & 1], st));
            if (c > 0) drain(c - 1);
        }
        drain(nch - 1);
    }
    void release() {
        std::lock_guard<std::mutex> lk(map_mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto &kv : per) {
            std::lock_guard<std::mutex> l2(kv.second.mu);
            cudaSetDevice(kv.first);
            for (int i = 0; i < 2; ++i) if (kv.second.buf[i]) { cudaFreeHost(kv.second.buf[i]); cudaEventDestroy(kv.second.ev[i]); kv.second.buf[i] = nullptr; }
        }
        cudaSetDevice(cur);
    }
};
HostStager &host_stager() { static HostStager h; return h; }

template <typename T>
T *dmalloc(size_t count) {
    return static_cast<T *>(block_cache().get(std::max<size_t>(count, 1) * sizeof(T)));
}
void dfree(const void *p) { block_cache().put(const_cast<void *>(p)); }

enum Phase : int { PH_GENE_E = 0, PH_ROWS, PH_ZI_B, PH_AR1, PH_MSTEP_V, PH_GENE_S, PH_SPARSE, PH_ZI_A, PH_HYPER, PH_COUNT };
const char *kPhaseNames[PH_COUNT] = {"estep_genes", "estep_cells_mstep_u", "zi_pass_genes", "allreduce", "mstep_v",
                                     "sparse_gene_pass", "sparse_update", "zi_pass_cells", "hyper_scalars"};

}  // namespace

struct pcmf_solver {
    pcmf_options opt{};
    int K = 0, KP = 0, p = 0;
    int64_t n = 0, n_global = 0, row_offset = 0, nnz = 0;
    bool zi = false, sparse = false, f32 = true;
    int device = 0;
    cudaStream_t stream = nullptr;
    // prob_D of the last cell pass, 32-bit fixed point in 8 x 8 tiles (zi_mma.cuh): the next gene pass streams it back
    // instead of recomputing it.  Only when it fits beside everything else; dropped by anything that redefines prob_D.
    uint32_t *pstore = nullptr; bool pstore_valid = false;
    int64_t *blkptr = nullptr; int nblk = 1;      // cell blocks of the gene passes (ColArgs::blkptr)
    bool zi_at_init = false;            // prob_D still is (X > 0) and the packed records of pass A do not exist yet
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_vals = nullptr;   // value upload of a host CSR, overlapped with the CSC sort
    bool own_stream = false;
    const Launchers *L = nullptr;
    int sms = 148;

    // X
    const int64_t *rowptr = nullptr; const int32_t *colidx = nullptr; const void *val_csr = nullptr;
    bool own_csr = false;
    int64_t *colptr = nullptr; int32_t *rowidx = nullptr; void *val_csc = nullptr;
    uint32_t *bitT = nullptr, *bitB = nullptr;
    double *zpackA = nullptr, *zpackB = nullptr;   // fragment-ordered scratch of the DMMA ZI passes
    bool fp32_mode = false;                        // pcmf_options.precision = 1: dense ZI passes on tcgen05 (zi_tc.cuh), FP32 gathers
    unsigned char *tcpackA = nullptr, *tcpackB = nullptr;   // their tile records (genes / cells)
    // explicit dense prob_D of a warm start: held until the first zero-inflation update (init_zi_param(prob_D, prior_D))
    double *Pexp = nullptr, *valw_csr = nullptr, *valw_csc = nullptr;
    bool explicitP = false, freq_ones = false;
    // conv_mode = 2 (RV coefficient, gap_factor_model.cpp:265-269): E[U], E[V] of the previous iterate are kept
    bool keep_old = false;
    double *EVold = nullptr, *rvbuf = nullptr;
    double rv_crit = 0;
    double rv_criterion();
    // 2-D tiled copy of X for the non-zero passes (xtile.cuh); when it exists the CSR / CSC passes only serve the iteration
    // that starts from an explicit prob_D, the monitors and the factor ordering
    // zero part of the log-likelihood monitor evaluated by the cell-side ZI pass on the way (k_zi_cells_mma<.., WITH_LL>):
    // requested by pcmf_optimize when monitor = 1; valid for the state of the step that produced it
    bool want_ll_zero = false; uint64_t state_ver = 0, ll_zero_ver = ~0ull; double ll_zero = 0;
    bool tiled = false;
    XTile xt{};
    void *xt_buf[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t *qmap = nullptr;                      // CSC entry -> position of its q in the tiled copy
    double2 *eupairs = nullptr;                    // (eu, eu o ElogU) rows for the B2 gene pass on the tiles (sparse models)
    void tile_gene_pass(int mode, ColArgs &ca, bool v_f32_csc);
    void build_tiles(const int64_t *perm_out);
    uint32_t *csr2csc = nullptr;                   // nnz: CSC position of every CSR entry (stored-q cell pass)
    double *qcsc = nullptr;                        // nnz: q_ij = x_ij / T_ij left by the gene passes, CSC order
    double *Lg = nullptr, *colsumX = nullptr, *colnnz = nullptr, *rowsumX = nullptr;
    double lgx_const = 0;
    double *xlogx = nullptr;            // per gene: sum_i x log x - x
    double ll_XX = 0, ll_Xmean = 0;     // poisson_loglike(X, X), poisson_loglike_vec(X, colmean(X)): constants of the deviance monitors
    double *mon = nullptr;              // monitor / ordering scratch: 8 + 64 + 64 doubles
    int32_t *d_order = nullptr;         // K
    double *old0[4] = {nullptr, nullptr, nullptr, nullptr};   // multi-init: previous run's last a1, a2, b1, b2
    int first_mode = 1;                 // value of first_iter for the first iteration (1: old = Ones, 2: old = old0)
    std::mt19937 rng;                   // one stream for all random initialisations (wrapper_gap_factor.h:277-284)
    // raw draws of the first random initialisation, produced by a host thread while the GPU preprocesses X
    struct DrawAhead { std::thread th; std::vector<uint32_t> u, v; unsigned long long start = 0; bool used = true; } ahead;
    void start_draw_ahead(const pcmf_init *in, int64_t n_rows, int64_t n_glob, int64_t row_off, int32_t genes);
    unsigned long long rng_consumed = 0;

    // U side (n x KP)
    double *a1 = nullptr, *a2 = nullptr, *EU[2] = {nullptr, nullptr}, *ElogU = nullptr, *eu = nullptr, *PV = nullptr;
    int cur = 0;   // EU[cur] is the current E[U]
    // V side (p x KP)
    double *b1 = nullptr, *b2 = nullptr, *EV = nullptr, *ElogV = nullptr, *ev = nullptr, *evw = nullptr, *EVS = nullptr,
           *wS = nullptr, *prob_S = nullptr, *Sd = nullptr;
    int32_t *S = nullptr, *flag = nullptr;
    double *prior_S = nullptr, *prior_D = nullptr, *cz = nullptr, *colw = nullptr;
    double *hyp = nullptr;   // alpha1, alpha2, beta1, beta2, colsum(EV o prob_S) kept for the next a2 update: 5 x 64
    // hyper-parameter matrices whose rows differ (only reachable from user-supplied alpha / beta, gap_factor_model.cpp:126-132):
    // alpha1, alpha2 (n x KP), beta1, beta2 (p x KP) and the two Gamma ELBO terms k_hyper_rows accumulates
    double *hrow[4] = {nullptr, nullptr, nullptr, nullptr}, *rowsg = nullptr;
    bool rows_mode = false;
    void hyper_update(bool update, const double *EU_now);
    double *csv = nullptr;   // cs_ev, cs_elogv, cs_evs: 3 x 64
    // reduction buffers
    double *ar1 = nullptr; size_t ar1_B1 = 0, ar1_tmp = 0, ar1_cseu = 0, ar1_cselog = 0, ar1_usc = 0, ar1_core = 0, ar1_B2 = 0,
                                  ar1_xlogT = 0, ar1_total = 0;
    double *ar2 = nullptr; size_t ar2_B1 = 0, ar2_B2 = 0, ar2_xlogT = 0, ar2_total = 0;
    double *ar3 = nullptr; size_t ar3_csP = 0, ar3_zsc = 0, ar3_total = 0;
    double *scal = nullptr;
    long long *guard[2] = {nullptr, nullptr};
    int gcur = 0;
    double2 *rowext = nullptr, *colext = nullptr;
    uint8_t *unchanged = nullptr;                   // per gene: mask row unchanged by the last sparse update
    double *ar2keep = nullptr;                      // this rank's sparse-pass sums before the all-reduce (world > 1)
    bool sparse_sums_valid = false;                 // ar2 / ar2keep hold the sums of the current statistics
    uint8_t *allmasked = nullptr;                   // per gene: every factor masked (sparse models)
    unsigned long long *dbg = nullptr;              // [0] cells kernel, [1] gene kernel: entries that took the exact formula   // per-cell / per-gene (min_k, max_k) of E[log U] / E[log V]
    double *h_scal = nullptr;   // pinned

    // iteration state (algorithm.h:50-64)
    int iter = 0;
    bool first_iter = true, converged = false;
    pcmf_init init_copy{};              // the caller's warm-start pointers (multi-init re-initialises from them)
    int nb_iter_stab = 0;
    double last_nodata = 0;        // ELBO without its data term, state after the last update
    bool have_pending = false;     // the ELBO of iteration (iter-1) still waits for its data term
    std::vector<double> elbo_hist; // per iteration (monitor)
    std::vector<int> factor_order;

    // instrumentation
    int64_t launches = 0;
    int64_t order_exact_rounds = 0;     // rounds of the greedy factor ordering that were repeated with FP64 logarithms
    double ph_ms[PH_COUNT] = {0};
    int64_t ph_launch[PH_COUNT] = {0};
    cudaEvent_t ev_a[PH_COUNT] = {nullptr}, ev_b[PH_COUNT] = {nullptr};
    bool ph_used[PH_COUNT] = {false};

    ~pcmf_solver();
    void setup_problem(const pcmf_problem &pr);
    void build_csc();
    void alloc_state();
    void init_state(const pcmf_init *in, bool reinit = false);
    void zero_gene_sums(double *B1, double *B2, double *xlogT, int mode);
    void run_init_stats(bool ones_stats = false, bool update_hyper = true);
    void allreduce(double *ptr, size_t count);
    void step(bool want_data);
    double data_term_now();
    struct Mon { double loglik, deviance, exp_dev; };
    Mon monitors_now(bool want_loglik);
    void order_factor_now();
    void multi_init(pcmf_result *res);
    void perturb_now();
    struct Saved { std::vector<std::pair<void *, size_t>> arrays; std::vector<void *> copies; int cur, gcur; bool first_iter, sparse_sums_valid, have_pending;
                   double last_nodata; std::vector<double> h; int iter, first_mode; };
    std::vector<std::pair<void *, size_t>> state_arrays();
    void save_state(Saved &sv);
    void restore_state(const Saved &sv);
    void begin(Phase ph) { CK(cudaEventRecord(ev_a[ph], stream)); }
    void end(Phase ph, int nl) { CK(cudaEventRecord(ev_b[ph], stream)); ph_used[ph] = true; ph_launch[ph] += nl; launches += nl; }
    void collect_phase_times() {
        for (int i = 0; i < PH_COUNT; ++i)
            if (ph_used[i]) { float ms = 0; CK(cudaEventElapsedTime(&ms, ev_a[i], ev_b[i])); ph_ms[i] += ms; ph_used[i] = false; }
    }
    void upload_colmajor(const double *host, int64_t rows, double *dst, double padval);
    void download_colmajor(const double *src, int64_t rows, double *host);
    void read_scal() {
        CK(cudaMemcpyAsync(h_scal, scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    int grid_for(int64_t work, int block) const {
        int64_t g = (work + block - 1) / block;
        return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)sms * 16));
    }
};

pcmf_solver::~pcmf_solver() {
    if (ahead.th.joinable()) ahead.th.join();
    cudaSetDevice(device);
    if (own_csr) { dfree((void *)rowptr); dfree((void *)colidx); dfree((void *)val_csr); }
    void *ptrs[] = {colptr, rowidx, val_csc, bitT, bitB, Lg, colsumX, colnnz, rowsumX, a1, a2, EU[0], EU[1], ElogU, eu, PV,
                    b1, b2, EV, ElogV, ev, evw, EVS, wS, prob_S, Sd, S, flag, prior_S, prior_D, cz, colw, hyp, csv, ar1, ar2,
                    ar3, scal, guard[0], guard[1], rowext, colext, allmasked, dbg, unchanged, ar2keep, zpackA, zpackB, EVold, rvbuf, Pexp, valw_csr, valw_csc, csr2csc, qcsc, xlogx, mon, d_order, old0[0], old0[1], old0[2], old0[3], pstore, blkptr, hrow[0], hrow[1], hrow[2], hrow[3], rowsg, tcpackA, tcpackB};
    for (void *q : ptrs) if (q) dfree(q);
    for (void *q : xt_buf) if (q) dfree(q);
    if (h_scal) cudaFreeHost(h_scal);
    for (int i = 0; i < PH_COUNT; ++i) { if (ev_a[i]) cudaEventDestroy(ev_a[i]); if (ev_b[i]) cudaEventDestroy(ev_b[i]); }
    if (ev_vals) cudaEventDestroy(ev_vals);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (own_stream && stream) cudaStreamDestroy(stream);
}

void pcmf_solver::allreduce(double *ptr, size_t count) {
    if (opt.world <= 1 || count == 0) return;
    if (!opt.allreduce) fail(PCMF_ECOMM, "world > 1 but no allreduce callback was supplied");
    const int rc = opt.allreduce(opt.allreduce_ctx, ptr, count, (void *)stream);
    if (rc != 0) fail(PCMF_ECOMM, "allreduce callback failed with code %d", rc);
}

// ---- preprocessing kernels local to this TU ------------------------------------------------------------------
namespace {
// dense column-major block (n x nb) -> per-cell non-zero counts; flag is raised when a value is not exact in FP32
template <typename T>
__global__ void k_dense_count(int64_t n, int64_t nb, const T *X, int64_t *cnt, int *flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = 0; bool bad = false;
        for (int64_t j = 0; j < nb; ++j) {
            const double x = (double)X[i + j * n];
            if (x != 0.0) { ++c; bad |= ((double)(float)x != x); }
        }
        cnt[i] += c;
        if (bad) atomicOr(flag, 1);
    }
}
// second sweep: write (gene, value) of every non-zero at the cell's running position (genes ascending within a cell)
template <typename T>
__global__ void k_dense_fill(int64_t n, int64_t nb, int32_t j0, const T *X, int64_t *pos, int32_t *colidx, void *val, int f32) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t e = pos[i];
        for (int64_t j = 0; j < nb; ++j) {
            const double x = (double)X[i + j * n];
            if (x != 0.0) {
                colidx[e] = j0 + (int32_t)j;
                if (f32) static_cast<float *>(val)[e] = (float)x; else static_cast<double *>(val)[e] = x;
                ++e;
            }
        }
        pos[i] = e;
    }
}
// CSR input is taken as it is (no copy for a device-resident matrix): check what every later kernel relies on.
// flag bit 1: row pointers not ascending / not starting at 0, 2: gene index outside [0, p), 4: an explicit zero (it would
// count as X != 0 in the zero-inflation occupancy, zi_gap_factor_model.cpp:104, and in the per-gene non-zero counts)
template <typename VT>
__global__ void k_validate_csr(int64_t n, int32_t p, int64_t nnz, const int64_t *rowptr, const int32_t *colidx, const VT *val, int *flag) {
    int bad = 0;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = gid; i < n; i += gsz) {
        const int64_t a = rowptr[i], b = rowptr[i + 1];
        if (a > b || a < 0 || b > nnz || (i == 0 && a != 0)) bad |= 1;
    }
    for (int64_t e = gid; e < nnz; e += gsz) {
        const int32_t j = colidx[e];
        if (j < 0 || j >= p) bad |= 2;
        if (val && val[e] == (VT)0) bad |= 4;
    }
    if (bad) atomicOr(flag, bad);
}
__global__ void k_expand_rows(int64_t n, const int64_t *rowptr, int32_t *rowids, int64_t *perm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nw)
        for (int64_t e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) { rowids[e] = (int32_t)i; perm[e] = e; }
}
__global__ void k_gather_rowidx(int64_t nnz, const int64_t *perm, const int32_t *rowids, int32_t *rowidx) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (int64_t)gridDim.x * blockDim.x) rowidx[t] = rowids[perm[t]];
}
template <typename VT>
__global__ void k_gather_vals(int64_t nnz, const int64_t *perm, const VT *val, VT *val_csc) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (int64_t)gridDim.x * blockDim.x) val_csc[t] = val[perm[t]];
}
// dgCMatrix input: sort keys (cell index; explicit zeros get the key n and fall off the end), gene index per entry
__global__ void k_csc_keys(int32_t p, int64_t n, const int32_t *colptr, const int32_t *rowidx, const double *x, int32_t *keys,
                           int32_t *colids, int64_t *perm, int *flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false, oob = false;
    for (int64_t j = warp; j < p; j += nw)
        for (int64_t e = colptr[j] + lane; e < colptr[j + 1]; e += 32) {
            const double v = x[e];
            const int32_t r = rowidx[e];
            oob |= (r < 0 || r >= n);
            keys[e] = (v != 0.0 && r >= 0 && r < n) ? r : (int32_t)n;
            colids[e] = (int32_t)j; perm[e] = e;
            bad |= ((double)(float)v != v);
        }
    if (bad) atomicOr(flag, 1);
    if (oob) atomicOr(flag, 2);
}
template <typename VT>
__global__ void k_gather_csr_from_csc(int64_t nnz, const int64_t *perm, const int32_t *colids, const double *x, int32_t *colidx, VT *val) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = perm[t];
        colidx[t] = colids[e];
        val[t] = (VT)x[e];
    }
}
// cell-block boundaries inside every gene's CSC segment (rows ascend within a gene): blkptr[b * p + j] = first position of
// gene j whose cell index is >= b * rows_blk; b = nblk gives colptr[j + 1]
__global__ void k_block_ptr(int32_t p, int nblk, int64_t rows_blk, const int64_t *colptr, const int32_t *rowidx, int64_t *blkptr) {
    const int64_t total = (int64_t)(nblk + 1) * p;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / p); const int j = (int)(t - (int64_t)b * p);
        int64_t lo = colptr[j], hi = colptr[j + 1];
        if (b == nblk) lo = hi;
        const int64_t bound = (int64_t)b * rows_blk;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowidx[mid] < bound) lo = mid + 1; else hi = mid; }
        blkptr[t] = lo;
    }
}
__global__ void k_invert_perm(int64_t nnz, const int64_t *perm, uint32_t *inv) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (int64_t)gridDim.x * blockDim.x)
        inv[perm[t]] = (uint32_t)t;
}
__global__ void k_colptr_from_sorted(int64_t nnz, int32_t p, const int32_t *keys, int64_t *colptr) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= p; j += gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nnz;   // first position with key >= j
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < j) lo = mid + 1; else hi = mid; }
        colptr[j] = lo;
    }
}
__global__ void k_init_guard(long long *g) {
    if (threadIdx.x == 0) {
        g[G_MINU] = dbl_ordered(1e300); g[G_MAXU] = dbl_ordered(-1e300);
        g[G_MINV] = dbl_ordered(1e300); g[G_MAXV] = dbl_ordered(-1e300);
    }
}
__global__ void k_refresh_weights(int32_t p, int K, int KP, int sparse, const double *prob_S, const double *colw, const double *ev,
                                  double *wS, double *evw) {
    const int64_t total = (int64_t)p * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % KP); const int j = (int)(o / KP);
        if (k >= K) continue;
        const double w = (sparse ? prob_S[o] : 1.0) * colw[j];
        wS[o] = w; evw[o] = ev[o] * w;
    }
}
// random_init_model_param (gap_factor_model.cpp:159-168): shape-1 Gamma with rate sqrt(K / mean) by inversion of
// u = (r + 0.5) / 2^32, mean = sums[row] / denom; pads are 1
__global__ void k_init_gamma_rows(int64_t rows, int K, int KP, const uint32_t *raw, const double *sums, double denom, double *dst) {
    const int64_t total = rows * KP;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(o % KP);
        const int64_t r = o / KP;
        if (k >= K) { dst[o] = 1.0; continue; }
        const double rate = sqrt((double)K / (sums[r] / denom));
        dst[o] = -log(((double)raw[r * K + k] + 0.5) / 4294967296.0) / rate;
    }
}
// perturb_param (gap_factor_model.cpp:177-192): shape += Uniform(-level, level), floored at 1e-6; raw draws row by row
__global__ void k_perturb_rows(int64_t rows, int K, int KP, const uint32_t *raw, double level, double *dst) {
    const int64_t total = rows * K;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / K; const int k = (int)(e - r * K);
        const double u = ((double)raw[e] + 0.5) / 4294967296.0;
        const double v = dst[r * KP + k] + (-level + 2.0 * level * u);
        dst[r * KP + k] = v < 1e-6 ? 1e-6 : v;
    }
}
// init_zi_param (zi_gap_factor_model.cpp:98-105) sets prob_D = (X > 0), so the dense cell pass collapses onto the
// non-zeros: PV_i = sum_{j in nz(i)} EVS_j, sum prob_D o UVt = sum_i EU_i . PV_i, the Bernoulli self term is 0 and
// colsum(prob_D)_j = nnz_j (k_zi_init_colsum).  One warp per cell, lane k owns factor k, four gathers in flight.
__global__ void __launch_bounds__(256) k_zi_init_cells(int64_t n, int KP, const int64_t *rowptr, const int32_t *colidx, const double *EU,
                                                       const double *EVS, double *PV, double *zsc) {
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double spl = 0;
    for (int64_t i = w0; i < n; i += nw) {
        const int64_t e0 = rowptr[i], e1 = rowptr[i + 1];
        for (int k = lane; k < KP; k += 32) {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            int64_t e = e0;
            for (; e + 4 <= e1; e += 4) {
                const int32_t j0 = colidx[e], j1 = colidx[e + 1], j2 = colidx[e + 2], j3 = colidx[e + 3];
                a0 += EVS[(size_t)j0 * KP + k]; a1 += EVS[(size_t)j1 * KP + k]; a2 += EVS[(size_t)j2 * KP + k]; a3 += EVS[(size_t)j3 * KP + k];
            }
            for (; e < e1; ++e) a0 += EVS[(size_t)colidx[e] * KP + k];
            const double pv = (a0 + a1) + (a2 + a3);
            PV[(size_t)i * KP + k] = pv;
            spl += EU[(size_t)i * KP + k] * pv;
        }
    }
    spl = block_sum(spl, sh);
    if (threadIdx.x == 0) atomicAdd(zsc, spl);
}
// Gene side of the same shortcut: while prob_D still is (X > 0), prob_D^T EU_new is the sum of the EU_new rows of the gene's
// non-zero cells (zi...cpp:341 with the prob_D of :98-105) -- one gather pass instead of the dense pass B of the first iteration.
__global__ void __launch_bounds__(256) k_zi_init_genes(int32_t p, int K, int KP, const int64_t *colptr, const int32_t *rowidx, const double *EU,
                                                       double *tmpMat) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = w0; j < p; j += nw) {
        const int64_t e0 = colptr[j], e1 = colptr[j + 1];
        for (int k = lane; k < K; k += 32) {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            int64_t e = e0;
            for (; e + 4 <= e1; e += 4) {
                const int64_t i0 = rowidx[e], i1 = rowidx[e + 1], i2 = rowidx[e + 2], i3 = rowidx[e + 3];
                a0 += EU[i0 * KP + k]; a1 += EU[i1 * KP + k]; a2 += EU[i2 * KP + k]; a3 += EU[i3 * KP + k];
            }
            for (; e < e1; ++e) a0 += EU[(int64_t)rowidx[e] * KP + k];
            tmpMat[j * KP + k] += (a0 + a1) + (a2 + a3);
        }
    }
}
__global__ void k_zi_init_colsum(int32_t p, const int64_t *colptr, double *colsumP) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x) colsumP[j] = (double)(colptr[j + 1] - colptr[j]);
}
__global__ void k_vec_scale(double *dst, const double *src, int64_t n, double s) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) dst[o] = src[o] * s;
}
}  // namespace

void pcmf_solver::setup_problem(const pcmf_problem &pr) {
    n = pr.n_local; n_global = pr.n_global > 0 ? pr.n_global : pr.n_local; row_offset = pr.row_offset; p = pr.p;
    if (n <= 0 || p <= 0) fail(PCMF_EINVAL, "empty count matrix");
    const int nrep = (pr.dense_f64 != nullptr) + (pr.dense_i32 != nullptr) + (pr.csr_rowptr != nullptr) + (pr.csc_colptr != nullptr);
    if (nrep != 1) fail(PCMF_EINVAL, "exactly one of dense_f64, dense_i32, csr_*, csc_* must be given");
    if (pr.csc_colptr) {
        // Matrix::dgCMatrix (slots p, i, x): compressed columns, int32 indices, double values, host memory.  Turned into
        // the cell-major CSR on the device by one stable radix sort on the cell index (genes stay ascending inside a cell);
        // explicit zeros are dropped on the way (they must not count as X != 0, zi_gap_factor_model.cpp:104).
        if (!pr.csc_rowidx || !pr.csc_val_f64) fail(PCMF_EINVAL, "incomplete CSC input");
        if (n >= 0x7fffffffLL) fail(PCMF_EINVAL, "CSC input: too many rows for int32 row indices");
        const int64_t nz_all = pr.csc_colptr[p];
        if (pr.csc_colptr[0] != 0 || nz_all < 0) fail(PCMF_EINVAL, "CSC input: bad column pointers");
        const size_t cnt = (size_t)std::max<int64_t>(nz_all, 1);
        int32_t *cp = dmalloc<int32_t>(p + 1), *ri = dmalloc<int32_t>(cnt); double *xv = dmalloc<double>(cnt);
        CK(cudaMemcpyAsync(cp, pr.csc_colptr, sizeof(int32_t) * (p + 1), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(ri, pr.csc_rowidx, sizeof(int32_t) * nz_all, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(xv, pr.csc_val_f64, sizeof(double) * nz_all, cudaMemcpyHostToDevice, stream));
        int end_bit = 1; while ((1ll << end_bit) <= n) ++end_bit;      // keys go up to n inclusive
        size_t tmp_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (const int64_t *)nullptr,
                                           (int64_t *)nullptr, nz_all, 0, end_bit, stream));
        int32_t *keys = dmalloc<int32_t>(cnt), *keys_out = dmalloc<int32_t>(cnt), *colids = dmalloc<int32_t>(cnt);
        int64_t *perm_in = dmalloc<int64_t>(cnt), *perm_out = dmalloc<int64_t>(cnt);
        void *tmp = dmalloc<unsigned char>(tmp_bytes);
        int *d_flag = dmalloc<int>(1);
        CK(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
        k_csc_keys<<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, n, cp, ri, xv, keys, colids, perm_in, d_flag);
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, perm_in, perm_out, nz_all, 0, end_bit, stream));
        int64_t *drp = dmalloc<int64_t>(n + 1);
        k_colptr_from_sorted<<<grid_for(n + 1, 256), 256, 0, stream>>>(nz_all, (int32_t)n, keys_out, drp);
        int h_flag = 0;
        CK(cudaMemcpyAsync(&nnz, drp + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (h_flag & 2) {
            for (void *q : {(void *)cp, (void *)ri, (void *)xv, (void *)keys, (void *)keys_out, (void *)colids, (void *)perm_in,
                            (void *)perm_out, tmp, (void *)d_flag, (void *)drp}) dfree(q);
            fail(PCMF_EINVAL, "CSC input: row index out of range");
        }
        f32 = (h_flag & 1) == 0;
        int32_t *dci = dmalloc<int32_t>((size_t)std::max<int64_t>(nnz, 1));
        void *dv = f32 ? (void *)dmalloc<float>((size_t)std::max<int64_t>(nnz, 1)) : (void *)dmalloc<double>((size_t)std::max<int64_t>(nnz, 1));
        if (f32) k_gather_csr_from_csc<float><<<grid_for(nnz, 256), 256, 0, stream>>>(nnz, perm_out, colids, xv, dci, (float *)dv);
        else k_gather_csr_from_csc<double><<<grid_for(nnz, 256), 256, 0, stream>>>(nnz, perm_out, colids, xv, dci, (double *)dv);
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        dfree(cp); dfree(ri); dfree(xv); dfree(keys); dfree(keys_out); dfree(colids); dfree(perm_in); dfree(perm_out); dfree(tmp); dfree(d_flag);
        rowptr = drp; colidx = dci; val_csr = dv; own_csr = true;
    } else if (pr.csr_rowptr) {
        if (!pr.csr_colidx || (!pr.csr_val_f32 && !pr.csr_val_f64)) fail(PCMF_EINVAL, "incomplete CSR input");
        f32 = pr.csr_val_f32 != nullptr;
        if (pr.csr_on_device) {
            rowptr = pr.csr_rowptr; colidx = pr.csr_colidx;
            val_csr = f32 ? (const void *)pr.csr_val_f32 : (const void *)pr.csr_val_f64;
            CK(cudaMemcpy(&nnz, rowptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost));
            own_csr = false;
        } else {
            nnz = pr.csr_rowptr[n];
            int64_t *rp = dmalloc<int64_t>(n + 1); int32_t *ci = dmalloc<int32_t>(nnz);
            void *vv = f32 ? (void *)dmalloc<float>(nnz) : (void *)dmalloc<double>(nnz);
            CK(cudaMemcpyAsync(rp, pr.csr_rowptr, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(ci, pr.csr_colidx, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, stream));
            // the values travel on a second stream: the radix sort of build_csc only needs the structure and runs under
            // their copy (pinned host arrays; a pageable source makes the copy synchronous and nothing is lost)
            cudaStream_t s2 = nullptr;
            CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ev_vals, cudaEventDisableTiming));
            CK(cudaMemcpyAsync(vv, f32 ? (const void *)pr.csr_val_f32 : (const void *)pr.csr_val_f64,
                               (f32 ? sizeof(float) : sizeof(double)) * nnz, cudaMemcpyHostToDevice, s2));
            CK(cudaEventRecord(ev_vals, s2));
            copy_stream = s2;
            rowptr = rp; colidx = ci; val_csr = vv; own_csr = true;
        }
        {
            int *d_flag = dmalloc<int>(1);
            CK(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
            if (nnz < 0) fail(PCMF_EINVAL, "CSR input: negative number of non-zeros");
            // (the values of a host CSR are still travelling on their own stream: they are checked in build_csc)
            const void *vchk = copy_stream ? nullptr : val_csr;
            if (f32) k_validate_csr<float><<<grid_for(std::max<int64_t>(n, nnz), 256), 256, 0, stream>>>(n, p, nnz, rowptr, colidx, (const float *)vchk, d_flag);
            else k_validate_csr<double><<<grid_for(std::max<int64_t>(n, nnz), 256), 256, 0, stream>>>(n, p, nnz, rowptr, colidx, (const double *)vchk, d_flag);
            int h_flag = 0;
            CK(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            dfree(d_flag);
            if (h_flag & 1) fail(PCMF_EINVAL, "CSR input: row pointers must start at 0 and ascend");
            if (h_flag & 2) fail(PCMF_EINVAL, "CSR input: gene index out of range");
            if (h_flag & 4) fail(PCMF_EINVAL, "CSR input: explicit zeros are not allowed (drop them: they would count as X != 0)");
        }
    } else {
        // dense R matrix (wrapper_gap_factor.h:204-211), INTSXP or REALSXP, column-major: uploaded as it is and turned into
        // CSR on the device (count per cell, scan, fill in ascending gene order); column blocks bound the staging buffer
        const int64_t colblk = std::max<int64_t>(1, std::min<int64_t>(p, (int64_t)(2ull << 30) / std::max<int64_t>(1, n * 8)));
        const size_t esz = pr.dense_f64 ? 8 : 4;
        unsigned char *stage = dmalloc<unsigned char>((size_t)n * colblk * esz);
        int64_t *cnt = dmalloc<int64_t>(n + 1);
        int *d_flag = dmalloc<int>(1);
        CK(cudaMemsetAsync(cnt, 0, sizeof(int64_t) * (n + 1), stream));
        CK(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
        auto upload = [&](int64_t j0, int64_t nb) {
            const unsigned char *src = pr.dense_f64 ? (const unsigned char *)(pr.dense_f64 + j0 * n) : (const unsigned char *)(pr.dense_i32 + j0 * n);
            CK(cudaMemcpyAsync(stage, src, (size_t)n * nb * esz, cudaMemcpyHostToDevice, stream));
        };
        for (int64_t j0 = 0; j0 < p; j0 += colblk) {          // pass 1: non-zeros per cell, FP32-exactness of the values
            const int64_t nb = std::min(colblk, (int64_t)p - j0);
            upload(j0, nb);
            if (pr.dense_f64) k_dense_count<double><<<grid_for(n, 256), 256, 0, stream>>>(n, nb, (const double *)stage, cnt + 1, d_flag);
            else k_dense_count<int32_t><<<grid_for(n, 256), 256, 0, stream>>>(n, nb, (const int32_t *)stage, cnt + 1, d_flag);
        }
        int64_t *drp = dmalloc<int64_t>(n + 1);
        {
            size_t tb = 0;
            CK(cub::DeviceScan::InclusiveSum(nullptr, tb, cnt, drp, (int)(n + 1), stream));
            void *tmp = dmalloc<unsigned char>(tb);
            CK(cub::DeviceScan::InclusiveSum(tmp, tb, cnt, drp, (int)(n + 1), stream));
            CK(cudaStreamSynchronize(stream));
            dfree(tmp);
        }
        int h_flag = 0;
        CK(cudaMemcpy(&nnz, drp + n, sizeof(int64_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
        f32 = h_flag == 0;
        int32_t *dci = dmalloc<int32_t>(nnz);
        void *dv = f32 ? (void *)dmalloc<float>(nnz) : (void *)dmalloc<double>(nnz);
        CK(cudaMemcpyAsync(cnt, drp, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, stream));   // running write position per cell
        for (int64_t j0 = 0; j0 < p; j0 += colblk) {          // pass 2: fill
            const int64_t nb = std::min(colblk, (int64_t)p - j0);
            if (colblk < p) upload(j0, nb);
            if (pr.dense_f64) k_dense_fill<double><<<grid_for(n, 256), 256, 0, stream>>>(n, nb, (int32_t)j0, (const double *)stage, cnt, dci, dv, f32 ? 1 : 0);
            else k_dense_fill<int32_t><<<grid_for(n, 256), 256, 0, stream>>>(n, nb, (int32_t)j0, (const int32_t *)stage, cnt, dci, dv, f32 ? 1 : 0);
        }
        CK(cudaStreamSynchronize(stream));
        dfree(stage); dfree(cnt); dfree(d_flag);
        rowptr = drp; colidx = dci; val_csr = dv; own_csr = true;
    }
    build_csc();
}

void pcmf_solver::build_csc() {
    StageTimer tm(stream);
    // CSR -> CSC by a stable radix sort on the gene index (rows stay ascending inside a gene -> deterministic sums)
    // one scratch block for all the temporaries of the sort (one cudaMalloc / cudaFree instead of ten)
    int end_bit = 1; while ((1ll << end_bit) < (int64_t)p) ++end_bit;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (const int64_t *)nullptr,
                                       (int64_t *)nullptr, nnz, 0, end_bit, stream));
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t o_rowids = 0, o_pin = o_rowids + up(4 * (size_t)nnz), o_pout = o_pin + up(8 * (size_t)nnz),
                 o_keys = o_pout + up(8 * (size_t)nnz), o_tmp = o_keys + up(4 * (size_t)nnz), scratch_bytes = o_tmp + up(tmp_bytes);
    unsigned char *scratch = dmalloc<unsigned char>(scratch_bytes);
    tm.mark("csc: scratch allocation");
    int32_t *rowids = reinterpret_cast<int32_t *>(scratch + o_rowids);
    int64_t *perm_in = reinterpret_cast<int64_t *>(scratch + o_pin), *perm_out = reinterpret_cast<int64_t *>(scratch + o_pout);
    int32_t *keys_out = reinterpret_cast<int32_t *>(scratch + o_keys);
    void *tmp = scratch + o_tmp;
    k_expand_rows<<<grid_for(n * 32, 256), 256, 0, stream>>>(n, rowptr, rowids, perm_in);
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, colidx, keys_out, perm_in, perm_out, nnz, 0, end_bit, stream));
    tm.mark("csc: expand + radix sort");
    colptr = dmalloc<int64_t>(p + 1);
    rowidx = dmalloc<int32_t>(nnz);
    tm.mark("csc: rowidx allocation");
    // everything that only needs the STRUCTURE first -- it runs under the upload of the values (second stream)
    k_colptr_from_sorted<<<grid_for(p + 1, 256), 256, 0, stream>>>(nnz, p, keys_out, colptr);
    k_gather_rowidx<<<grid_for(nnz, 256), 256, 0, stream>>>(nnz, perm_out, rowids, rowidx);
    // The non-zero passes run on a 2-D tiled copy of X (xtile.cuh) when the cell pass has enough CTAs to fill the machine
    // (one per 256 cells) -- or, with fewer cells, when the matrix is dense enough that a tile's segments are long (C3, 50k x 5k
    // at 30 %: spike-and-slab pass 3.47 -> 2.05 ms, cell pass 1.71 -> 1.37 ms) -- and the factor block fits shared memory twice
    // over; kernel_variant bit 2048 forces the tiles (small test problems), bit 4096 keeps the CSR / CSC passes.
    tiled = f32 && KP >= XT_MINKP && KP <= XT_MAXKP && L->tile_cells && nnz > 0 && nnz < 0xffffffffLL && !(opt.dev_flags & 4096) &&
            (n >= 65536 || (n >= 16384 && nnz >= 50000000) || (opt.dev_flags & 2048));
    // ZI / sparse models: the gene passes leave q_ij in CSC order and the cell pass reads it back through this map
    if (!tiled && (zi || sparse) && nnz < 0xffffffffLL && !(opt.dev_flags & 16)) {
        csr2csc = dmalloc<uint32_t>(nnz);
        k_invert_perm<<<grid_for(nnz, 256), 256, 0, stream>>>(nnz, perm_out, csr2csc);
        qcsc = dmalloc<double>(nnz);
    }
    if (zi) {
        const size_t wT = (size_t)((p + 31) / 32) * n, wB = (size_t)((n + 31) / 32) * p;
        bitT = dmalloc<uint32_t>(wT); bitB = dmalloc<uint32_t>(wB);
        CK(cudaMemsetAsync(bitT, 0, wT * 4, stream)); CK(cudaMemsetAsync(bitB, 0, wB * 4, stream));
        k_build_bitmaps<<<grid_for(n * 32, 256), 256, 0, stream>>>(n, p, rowptr, colidx, bitT, bitB);
    }
    // gene passes walk the cells in blocks whose slice of the cell table (eu and ElogU rows, 16 KP bytes per cell) fits L2
    // with room to spare: 28 MB measured best at C4 (10 / 14 / 20 / 24 / 30 / 40 / 80 MB: 46.8 / 44.6 / 43.1 / 42.4 / 42.1 / 47.1 /
    // 50.6 ms for the spike-and-slab pass, 50.8 unblocked); PCMF_BLK_MB overrides, kernel_variant bit 512: one block
    {
        // (kernel_variant bit 1024: 64-cell blocks, so that small test problems walk several of them)
        const int64_t rows_blk = (opt.dev_flags & 1024) ? 64 : std::max<int64_t>(4096, ((int64_t)(getenv("PCMF_BLK_MB") ? atoi(getenv("PCMF_BLK_MB")) : 28) << 20) / (16 * KP));
        nblk = (opt.dev_flags & 512) ? 1 : (int)std::min<int64_t>(256, (n + rows_blk - 1) / rows_blk);
        if (nblk > 1) {
            const int64_t rb = ((n + nblk - 1) / nblk + 31) / 32 * 32;
            blkptr = dmalloc<int64_t>((size_t)(nblk + 1) * p);
            k_block_ptr<<<grid_for((int64_t)(nblk + 1) * p, 256), 256, 0, stream>>>(p, nblk, rb, colptr, rowidx, blkptr);
        }
    }
    Lg = dmalloc<double>(p); colsumX = dmalloc<double>(p); colnnz = dmalloc<double>(p); rowsumX = dmalloc<double>(n);
    xlogx = dmalloc<double>(p);
    val_csc = f32 ? (void *)dmalloc<float>(nnz) : (void *)dmalloc<double>(nnz);
    if (ev_vals) {                      // the values of a host CSR were uploaded on their own stream
        CK(cudaStreamSynchronize(stream));
        tm.mark("csc: rowidx, position map, occupancy bitmaps");
        CK(cudaStreamWaitEvent(stream, ev_vals, 0));
        CK(cudaStreamSynchronize(copy_stream));
        cudaEventDestroy(ev_vals); cudaStreamDestroy(copy_stream);
        ev_vals = nullptr; copy_stream = nullptr;
        tm.mark("csc: wait for the value upload");
    }
    if (f32) {
        k_gather_vals<float><<<grid_for(nnz, 256), 256, 0, stream>>>(nnz, perm_out, (const float *)val_csr, (float *)val_csc);
        k_gene_constants<float><<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, colptr, val_csc, Lg, colsumX, colnnz, xlogx);
        k_row_sums<float><<<grid_for(n * 32, 256), 256, 0, stream>>>(n, rowptr, val_csr, rowsumX);
    } else {
        k_gather_vals<double><<<grid_for(nnz, 256), 256, 0, stream>>>(nnz, perm_out, (const double *)val_csr, (double *)val_csc);
        k_gene_constants<double><<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, colptr, val_csc, Lg, colsumX, colnnz, xlogx);
        k_row_sums<double><<<grid_for(n * 32, 256), 256, 0, stream>>>(n, rowptr, val_csr, rowsumX);
    }
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    tm.mark("csc: value gather, gene / cell constants");
    if (tiled) { build_tiles(perm_out); tm.mark("tiled copy of X"); }
    dfree(scratch);
    // gene constants are sums over cells -> all-reduce across the row shards
    allreduce(Lg, p); allreduce(colsumX, p); allreduce(colnnz, p); allreduce(xlogx, p);
    std::vector<double> h(p), hx(p), hs(p);
    CK(cudaMemcpyAsync(h.data(), Lg, sizeof(double) * p, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(hx.data(), xlogx, sizeof(double) * p, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(hs.data(), colsumX, sizeof(double) * p, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    lgx_const = 0; for (double x : h) lgx_const += x;
    {   // explicit zeros among the stored entries (values of a host CSR arrive after the structure was checked)
        std::vector<double> hn(p);
        CK(cudaMemcpy(hn.data(), colnnz, sizeof(double) * p, cudaMemcpyDeviceToHost));
        double tot = 0; for (double x : hn) tot += x;
        if (opt.world <= 1 && tot != (double)nnz)
            fail(PCMF_EINVAL, "CSR input: explicit zeros are not allowed (drop them: they would count as X != 0)");
    }
    // constants of deviance / exp_deviance (gap...cpp:329-341): poisson_loglike(X, X) and poisson_loglike_vec(X, colmean X)
    ll_XX = -lgx_const; ll_Xmean = -lgx_const;
    for (int j = 0; j < p; ++j) {
        ll_XX += hx[j];
        const double m = hs[j] / (double)n_global;
        ll_Xmean += hs[j] * (m > 0 ? std::log(m) : 0.0) - (double)n_global * m;
    }
    CK(cudaGetLastError());
}

// The tiled copy of X (xtile.cuh) from the CSR and CSC copies: no further sort -- inside a cell's CSR segment the genes
// ascend, so the entries of a (cell, gene block) pair are a run whose length and start give the cell-order position, and
// likewise for a gene's CSC segment; the CSR -> CSC map links the two orders.
void pcmf_solver::build_tiles(const int64_t *perm_out) {
    StageTimer tm(stream);
    const int nI = (int)((n + XT_TI - 1) / XT_TI), nJ = (p + XT_TJ - 1) / XT_TJ;
    const int64_t nT = (int64_t)nI * nJ;
    const bool gene_side = L->tile_genes != nullptr;       // (else the gene passes stay on CSC and only the cell order is built)
    int64_t *tptr = dmalloc<int64_t>(nT + 1);
    uint2 *grec = gene_side ? dmalloc<uint2>(nnz) : nullptr;
    uint16_t *gofs = gene_side ? dmalloc<uint16_t>((size_t)nT * (XT_TJ + 1)) : nullptr;
    uint8_t *gperm = gene_side ? dmalloc<uint8_t>((size_t)nT * XT_TJ) : nullptr;
    uint8_t *ccol = dmalloc<uint8_t>(nnz);
    double *tq = dmalloc<double>(nnz);
    uint16_t *cofs = dmalloc<uint16_t>((size_t)nT * (XT_TI + 1));
    uint8_t *cperm = dmalloc<uint8_t>((size_t)nT * XT_TI);
    qmap = dmalloc<uint32_t>(nnz);
    if (sparse && L->tile_genes) eupairs = dmalloc<double2>((size_t)n * KP);
    void *bufs[10] = {tptr, grec, gofs, gperm, ccol, tq, cofs, cperm, qmap, eupairs};
    for (int i = 0; i < 10; ++i) xt_buf[i] = bufs[i];
    uint32_t *runofs = dmalloc<uint32_t>((size_t)nJ * n);
    int64_t *tcnt = dmalloc<int64_t>(nT + 1);
    if (gene_side) CK(cudaMemsetAsync(gofs, 0, (size_t)nT * (XT_TJ + 1) * 2, stream));
    CK(cudaMemsetAsync(cofs, 0, (size_t)nT * (XT_TI + 1) * 2, stream));
    CK(cudaMemsetAsync(tcnt, 0, (nT + 1) * 8, stream));
    tm.mark("tiles: allocation");
    k_xt_count<true><<<grid_for(n * 32, 256), 256, 0, stream>>>(n, rowptr, colidx, nJ, cofs, runofs);
    if (gene_side) k_xt_count<false><<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, colptr, rowidx, nJ, gofs, nullptr);
    k_xt_scan<XT_TI><<<grid_for(nT * 32, 128), 128, 0, stream>>>(nT, cofs, tcnt);
    if (gene_side) k_xt_scan<XT_TJ><<<grid_for(nT * 32, 128), 128, 0, stream>>>(nT, gofs, nullptr);
    {
        size_t tb = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, tcnt, tptr, nT + 1, stream));
        void *tmp = dmalloc<unsigned char>(tb);
        CK(cub::DeviceScan::ExclusiveSum(tmp, tb, tcnt, tptr, nT + 1, stream));
        CK(cudaStreamSynchronize(stream));
        dfree(tmp);
    }
    tm.mark("tiles: counts, offsets");
    k_xt_perm<XT_TI><<<grid_for(nT * 32, 128), 128, 0, stream>>>(nT, cofs, cperm);
    if (gene_side) k_xt_perm<XT_TJ><<<grid_for(nT * 32, 128), 128, 0, stream>>>(nT, gofs, gperm);
    tm.mark("tiles: segment ranking");
    k_xt_scatter_rows<<<(unsigned)std::min<int64_t>(n, (int64_t)sms * 64), 128, 0, stream>>>(n, rowptr, colidx, nJ, tptr, cofs, runofs, ccol);
    tm.mark("tiles: cell-order scatter");
    k_xt_scatter_cols<<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, n, colptr, rowidx, (const float *)val_csc, perm_out, rowptr, nJ, tptr,
                                                                           gofs, cofs, runofs, grec, qmap);
    tm.mark("tiles: gene-order scatter");
    CK(cudaMemsetAsync(tq, 0, (size_t)nnz * 8, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    dfree(runofs); dfree(tcnt);
    xt.nI = nI; xt.nJ = nJ; xt.tptr = tptr; xt.grec = grec; xt.gofs = gofs; xt.gperm = gperm; xt.ccol = ccol; xt.q = tq; xt.cofs = cofs; xt.cperm = cperm;
}

void pcmf_solver::alloc_state() {
    const size_t nK = (size_t)n * KP, pK = (size_t)p * KP;
    a1 = dmalloc<double>(nK); a2 = dmalloc<double>(nK); EU[0] = dmalloc<double>(nK); ElogU = dmalloc<double>(nK); eu = dmalloc<double>(nK);
    keep_old = opt.conv_mode == 2;
    if (keep_old) { EVold = dmalloc<double>((size_t)p * KP); rvbuf = dmalloc<double>(6 * 64 * 64); }
    if (keep_old && !zi) { EU[1] = dmalloc<double>(nK); CK(cudaMemsetAsync(EU[1], 0, nK * 8, stream)); }
    if (zi) {
        EU[1] = dmalloc<double>(nK); PV = dmalloc<double>(nK); CK(cudaMemsetAsync(EU[1], 0, nK * 8, stream));
        zpackA = dmalloc<double>(L->zi_pack_doubles(p, 0));
        if (fp32_mode) { tcpackA = dmalloc<unsigned char>(L->tc_pack_bytes(p)); tcpackB = dmalloc<unsigned char>(L->tc_pack_bytes(n)); }
        else zpackB = dmalloc<double>(L->zi_pack_doubles(n, 1));
    }
    b1 = dmalloc<double>(pK); b2 = dmalloc<double>(pK); EV = dmalloc<double>(pK); ElogV = dmalloc<double>(pK);
    ev = dmalloc<double>(pK); evw = dmalloc<double>(pK); EVS = dmalloc<double>(pK); wS = dmalloc<double>(pK);
    prob_S = dmalloc<double>(pK); Sd = dmalloc<double>(pK); S = dmalloc<int32_t>(pK); flag = dmalloc<int32_t>(p);
    prior_S = dmalloc<double>(p); prior_D = dmalloc<double>(p); cz = dmalloc<double>(p); colw = dmalloc<double>(p);
    hyp = dmalloc<double>(5 * 64); csv = dmalloc<double>(3 * 64);
    CK(cudaMemsetAsync(S, 0, pK * 4, stream)); CK(cudaMemsetAsync(flag, 0, p * 4, stream));
    k_fill<<<grid_for(pK, 256), 256, 0, stream>>>(prob_S, pK, 1.0);
    k_fill<<<grid_for(pK, 256), 256, 0, stream>>>(Sd, pK, 1.0);
    k_fill<<<grid_for(p, 256), 256, 0, stream>>>(colw, p, 1.0);
    k_fill<<<grid_for(p, 256), 256, 0, stream>>>(prior_S, p, 1.0);
    k_fill<<<grid_for(p, 256), 256, 0, stream>>>(prior_D, p, 1.0);
    // reduction buffers: [B1 | tmpMat | cs_eu | cs_elog | usc(4)] is the per-iteration core, [B2 | xlogT] only when the
    // ELBO data term is wanted
    ar1_B1 = 0; ar1_tmp = ar1_B1 + pK; ar1_cseu = ar1_tmp + (zi ? pK : 0); ar1_cselog = ar1_cseu + 64; ar1_usc = ar1_cselog + 64;
    ar1_core = ar1_usc + 4; ar1_B2 = ar1_core; ar1_xlogT = ar1_B2 + pK; ar1_total = ar1_xlogT + p;
    ar1 = dmalloc<double>(ar1_total); CK(cudaMemsetAsync(ar1, 0, ar1_total * 8, stream));
    ar2_B1 = 0; ar2_B2 = pK; ar2_xlogT = 2 * pK; ar2_total = 2 * pK + p;
    ar2 = dmalloc<double>(ar2_total); CK(cudaMemsetAsync(ar2, 0, ar2_total * 8, stream));
    ar3_csP = 0; ar3_zsc = p; ar3_total = p + 3;
    ar3 = dmalloc<double>(ar3_total); CK(cudaMemsetAsync(ar3, 0, ar3_total * 8, stream));
    scal = dmalloc<double>(S_COUNT); CK(cudaMemsetAsync(scal, 0, S_COUNT * 8, stream));
    guard[0] = dmalloc<long long>(G_COUNT); guard[1] = dmalloc<long long>(G_COUNT);
    rowext = dmalloc<double2>(n); colext = dmalloc<double2>(p);
    allmasked = dmalloc<uint8_t>(p); CK(cudaMemsetAsync(allmasked, 0, p, stream));
    unchanged = dmalloc<uint8_t>(p); CK(cudaMemsetAsync(unchanged, 0, p, stream));
    if (opt.world > 1) ar2keep = dmalloc<double>(2 * (size_t)p * KP);
    mon = dmalloc<double>(8 + 64 * 3); d_order = dmalloc<int32_t>(64);
    dbg = dmalloc<unsigned long long>(4); CK(cudaMemsetAsync(dbg, 0, 32, stream));
    k_init_guard<<<1, 32, 0, stream>>>(guard[0]); k_init_guard<<<1, 32, 0, stream>>>(guard[1]);
    // kernel_variant bit 64 forces the stored-prob_D path (small test problems), bit 128 disables it
    if (zi && !fp32_mode && !(opt.dev_flags & (4 | 128)) && ((double)n * (double)p >= 5e7 || (opt.dev_flags & 64)) &&
        (double)((n + 127) / 128 * 16) * (double)((p + 7) / 8) < 4.0e9) {
        // cell groups padded to whole CTAs of the cell pass (128 cells): it stores every tile it computes
        const size_t need = (size_t)((n + 127) / 128 * 16) * (size_t)((p + 7) / 8) * 64 * sizeof(uint32_t);
        bool fits = block_cache().has_block(need);
        for (int attempt = 0; attempt < 2 && !fits; ++attempt) {
            size_t fr = 0, tot = 0;
            CK(cudaMemGetInfo(&fr, &tot));
            fits = fr > need + std::max((size_t)8 << 30, tot / 12);     // room for outputs, monitors, saved multi-init states
            if (!fits && attempt == 0) block_cache().release_all();       // the preprocessing scratch may still sit in the cache
        }
        if (fits) pstore = dmalloc<uint32_t>(need / sizeof(uint32_t));
    }
    CK(cudaMallocHost(&h_scal, sizeof(double) * S_COUNT));
    for (int i = 0; i < PH_COUNT; ++i) { CK(cudaEventCreate(&ev_a[i])); CK(cudaEventCreate(&ev_b[i])); }
    factor_order.resize(K);
    for (int k = 0; k < K; ++k) factor_order[k] = k;
}

void pcmf_solver::upload_colmajor(const double *host, int64_t rows, double *dst, double padval) {
    double *tmp = dmalloc<double>((size_t)rows * K);
    CK(cudaMemcpyAsync(tmp, host, sizeof(double) * rows * K, cudaMemcpyHostToDevice, stream));
    k_colmajor_to_rowpad<<<grid_for(rows * KP, 256), 256, 0, stream>>>(rows, K, KP, tmp, dst, padval);
    CK(cudaStreamSynchronize(stream));
    dfree(tmp);
}
void pcmf_solver::download_colmajor(const double *src, int64_t rows, double *host) {
    double *tmp = dmalloc<double>((size_t)rows * K);
    k_rowpad_to_colmajor<<<grid_for(rows * K, 256), 256, 0, stream>>>(rows, K, KP, src, tmp);
    host_stager().d2h(host, tmp, sizeof(double) * rows * K, stream);
    dfree(tmp);
}

// The seeded random initialisation walks ONE MT19937 stream over (n_global + p) K draws on the host (~0.1 s at C4).  Nothing
// it needs depends on the GPU, so the walk starts on a host thread as soon as the arguments are known and is collected
// by init_state after the preprocessing of X; `start` is the stream position the draws assume (checked there).
void pcmf_solver::start_draw_ahead(const pcmf_init *in, int64_t n_rows, int64_t n_glob, int64_t row_off, int32_t genes) {
    static const pcmf_init none{};
    if (!in) in = &none;
    const bool random_ab = !in->U && !in->a1 && !in->alpha1;
    if (!random_ab) return;
    const bool random_S = sparse && !(in->prob_S && in->prior_S);
    ahead.start = random_S ? (unsigned long long)genes * K : 0ull;     // init_sparse_param draws p K uniforms first
    ahead.used = false;
    const std::mt19937 g0 = rng;
    const int Kc = K;
    const unsigned long long start = ahead.start;
    ahead.th = std::thread([this, g0, Kc, start, n_rows, n_glob, row_off, genes] {
        std::mt19937 g = g0;
        g.discard(start + (unsigned long long)row_off * Kc);
        ahead.u.resize((size_t)n_rows * Kc);
        for (size_t e = 0; e < ahead.u.size(); ++e) ahead.u[e] = (uint32_t)g();
        g.discard((unsigned long long)(n_glob - row_off - n_rows) * Kc);
        ahead.v.resize((size_t)genes * Kc);
        for (size_t e = 0; e < ahead.v.size(); ++e) ahead.v[e] = (uint32_t)g();
    });
}

void pcmf_solver::init_state(const pcmf_init *in, bool reinit) {
    const pcmf_init none{};
    if (!in) in = &none;
    // ---- argument consistency, with the reference's messages (wrapper_gap_factor.h:215-258)
    if ((in->U != nullptr) != (in->V != nullptr))
        fail(PCMF_EINVAL, "You should supply initial values for both U and V or for neither of them");
    const int nab = (in->a1 != nullptr) + (in->a2 != nullptr) + (in->b1 != nullptr) + (in->b2 != nullptr);
    if (nab != 0 && nab != 4) fail(PCMF_EINVAL, "You should supply initial values for all a1, a2, b1, b2 or for none of them");
    const int nhy = (in->alpha1 != nullptr) + (in->alpha2 != nullptr) + (in->beta1 != nullptr) + (in->beta2 != nullptr);
    if (nhy != 0 && nhy != 4)
        fail(PCMF_EINVAL, "You should supply initial values for all alpha1, alpha2, beta1, beta2 or for none of them");

    // hyper-parameters: Ones (gap_factor_model.cpp:56-59) unless given
    if ((in->prob_D != nullptr) != (in->prior_D != nullptr))
        fail(PCMF_EINVAL, "You should supply initial values for both prob_D and prior_D or for neither of them");
    if (zi && in->prob_D) {
        // init_zi_param(prob_D, prior_D) (zi...cpp:110-114): the given prior is overwritten by colmean(prob_D) in the
        // update_hyper_param that every initialisation ends with (zi...cpp:146-151, :370-372); freq_D keeps its
        // constructor value Ones (zi...cpp:57)
        if (Pexp) { dfree(Pexp); dfree(valw_csr); dfree(valw_csc); }
        Pexp = dmalloc<double>((size_t)n * p); valw_csr = dmalloc<double>(nnz); valw_csc = dmalloc<double>(nnz);
        CK(cudaMemcpyAsync(Pexp, in->prob_D, sizeof(double) * (size_t)n * p, cudaMemcpyHostToDevice, stream));
        explicitP = true; freq_ones = true;
    } else {
        explicitP = false;
    }
    // The reference's else-if chain (wrapper_gap_factor.h:352-375, wrapper_sparse_gap_factor.h:399-417): U, V > a1..b2 >
    // alpha / beta > random.  alpha / beta are therefore only looked at when nothing else is given, and then
    // init_hyper_param (gap_factor_model.cpp:126-132) is ALL that runs: the variational parameters and the statistics keep
    // the constructors' Ones (gap...cpp:67-70, model.cpp:220-223), nothing is drawn, no update_hyper_param.
    const bool hyper_only = nhy == 4 && !in->U && nab != 4;
    // a later multi-init run starts from the hyper-parameters the previous run ended with (the reference re-uses one
    // model object, wrapper_gap_factor.h:299-301: update_prior_gamma_param then reads log(alpha2) of that run)
    if (!reinit || hyper_only) {
        std::vector<double> h(4 * 64, 1.0);
        rows_mode = false;
        if (hyper_only) {
            const double *src[4] = {in->alpha1, in->alpha2, in->beta1, in->beta2};
            const int64_t rows[4] = {n, n, p, p};
            if (in->hyper_rows)       // n x K / p x K column-major as the reference takes them: do the rows differ?
                for (int v = 0; v < 4 && !rows_mode; ++v)
                    for (int k = 0; k < K && !rows_mode; ++k)
                        for (int64_t r = 1; r < rows[v]; ++r)
                            if (src[v][r + (int64_t)k * rows[v]] != src[v][(int64_t)k * rows[v]]) { rows_mode = true; break; }
            for (int v = 0; v < 4; ++v)
                for (int k = 0; k < K; ++k) h[64 * v + k] = in->hyper_rows ? src[v][(int64_t)k * rows[v]] : src[v][k];
            if (rows_mode) {
                for (int v = 0; v < 4; ++v) {
                    if (!hrow[v]) hrow[v] = dmalloc<double>((size_t)rows[v] * KP);
                    upload_colmajor(src[v], rows[v], hrow[v], 1.0);
                }
                if (!rowsg) rowsg = dmalloc<double>(2);
            }
        }
        CK(cudaMemcpyAsync(hyp, h.data(), sizeof(double) * 4 * 64, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
    }

    // sparse parameters first (wrapper_sparse_gap_factor.h:390-394)
    // every draw comes from ONE stream, positioned by the number of draws consumed so far, so that a cell-sharded run
    // and successive multi-init runs see the same numbers as a single process would
    auto unif = [](std::mt19937 &g) { return ((double)g() + 0.5) / 4294967296.0; };
    std::mt19937 base = rng; base.discard(rng_consumed);
    if (sparse) {
        if (in->prob_S && in->prior_S) {
            upload_colmajor(in->prob_S, p, prob_S, 0.0);
            CK(cudaMemcpy(prior_S, in->prior_S, sizeof(double) * p, cudaMemcpyHostToDevice));
        } else {
            // init_sparse_param(sel_bound, rng): uniform prob_S (sparse_gap_factor_model.cpp:71-88)
            std::vector<double> ps((size_t)p * K), pr(p);
            std::mt19937 gs = base;
            for (int j = 0; j < p; ++j) {
                double s = 0;
                for (int k = 0; k < K; ++k) { const double u = unif(gs); ps[j + (size_t)k * p] = u; s += u; }
                pr[j] = s / K;
            }
            base.discard((unsigned long long)p * K); rng_consumed += (unsigned long long)p * K;
            upload_colmajor(ps.data(), p, prob_S, 0.0);
            CK(cudaMemcpy(prior_S, pr.data(), sizeof(double) * p, cudaMemcpyHostToDevice));
        }
    }
    // variational parameters (wrapper_gap_factor.h:352-375)
    if (hyper_only) {
        for (double *A : {a1, a2}) k_fill<<<grid_for(n * KP, 256), 256, 0, stream>>>(A, n * KP, 1.0);
        for (double *A : {b1, b2}) k_fill<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(A, (int64_t)p * KP, 1.0);
        if (ahead.th.joinable()) ahead.th.join();
        ahead.used = true;
        run_init_stats(true, false);
        return;
    }
    if (in->U) {
        upload_colmajor(in->U, n, a1, 1.0); upload_colmajor(in->V, p, b1, 1.0);
        k_fill<<<grid_for(n * KP, 256), 256, 0, stream>>>(a2, n * KP, 1.0);
        k_fill<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(b2, (int64_t)p * KP, 1.0);
    } else if (nab == 4) {
        upload_colmajor(in->a1, n, a1, 1.0); upload_colmajor(in->a2, n, a2, 1.0);
        upload_colmajor(in->b1, p, b1, 1.0); upload_colmajor(in->b2, p, b2, 1.0);
    } else {
        // random_init_model_param (gap_factor_model.cpp:147-174): a1(i,.) ~ Gamma(1, rate sqrt(K / rowmean_i)),
        // b1(j,.) ~ Gamma(1, rate sqrt(K / colmean_j)); one MT19937 stream, cells first (in global order) then genes.
        // The raw 32-bit draws come off the single host stream in order; the inversion -log(u) / rate runs on the device,
        // straight into the padded row-major layout (row / column sums of X are already there).
        auto invert = [&](const std::vector<uint32_t> &raw, int64_t rows, const double *sums, double denom, double *dst) {
            uint32_t *d_raw = dmalloc<uint32_t>(raw.size());
            CK(cudaMemcpyAsync(d_raw, raw.data(), raw.size() * 4, cudaMemcpyHostToDevice, stream));
            k_init_gamma_rows<<<grid_for(rows * KP, 256), 256, 0, stream>>>(rows, K, KP, d_raw, sums, denom, dst);
            CK(cudaStreamSynchronize(stream));
            dfree(d_raw);
        };
        if (ahead.th.joinable()) ahead.th.join();
        if (!ahead.used && ahead.start == rng_consumed) {
            // the draws were produced while the GPU was busy with the preprocessing of X (pcmf_create)
            invert(ahead.u, n, rowsumX, (double)p, a1);
            invert(ahead.v, p, colsumX, (double)n_global, b1);
        } else {
            auto draw = [&](std::mt19937 g, int64_t rows, const double *sums, double denom, double *dst) {
                std::vector<uint32_t> raw((size_t)rows * K);
                for (size_t e = 0; e < raw.size(); ++e) raw[e] = (uint32_t)g();
                invert(raw, rows, sums, denom, dst);
            };
            std::mt19937 gu = base; gu.discard((unsigned long long)row_offset * K);
            draw(gu, n, rowsumX, (double)p, a1);
            std::mt19937 gv = base; gv.discard((unsigned long long)n_global * K);
            draw(gv, p, colsumX, (double)n_global, b1);
        }
        ahead.used = true; ahead.u = std::vector<uint32_t>(); ahead.v = std::vector<uint32_t>();
        rng_consumed += (unsigned long long)(n_global + p) * K;
        k_fill<<<grid_for(n * KP, 256), 256, 0, stream>>>(a2, n * KP, 1.0);
        k_fill<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(b2, (int64_t)p * KP, 1.0);
    }
    run_init_stats();
}

// the blocked gene pass adds its per-block sums atomically: start from zero
void pcmf_solver::zero_gene_sums(double *B1, double *B2, double *xlogT, int mode) {
    if (nblk <= 1 && !tiled) return;
    const size_t pK = (size_t)p * KP;
    CK(cudaMemsetAsync(B1, 0, pK * 8, stream));
    if (mode == GENE_B1_B2) CK(cudaMemsetAsync(B2, 0, pK * 8, stream));
    if (mode == GENE_B1_LOGT) CK(cudaMemsetAsync(xlogT, 0, (size_t)p * 8, stream));
}

// U_stats, V_stats, update_poisson_param, update_hyper_param after any initialisation (gap_factor_model.cpp:119-122)
void pcmf_solver::run_init_stats(bool ones_stats, bool update_hyper) {
    pstore_valid = false;             // prob_D is redefined by the initialisation
    const size_t pK = (size_t)p * KP;
    double *beta1 = hyp + 128, *beta2 = hyp + 192;
    CK(cudaMemsetAsync(scal, 0, S_COUNT * 8, stream));
    CK(cudaMemsetAsync(csv, 0, 3 * 64 * 8, stream));
    CK(cudaMemsetAsync(ar1 + ar1_cseu, 0, (64 + 64 + 4) * 8, stream));
    CK(cudaMemsetAsync(ar3, 0, ar3_total * 8, stream));
    k_init_guard<<<1, 32, 0, stream>>>(guard[gcur]);
    UInitArgs ua{n, K, KP, a1, a2, EU[cur], ElogU, eu, guard[gcur], ar1 + ar1_cseu, ar1 + ar1_cselog, ar1 + ar1_usc, ones_stats ? 1 : 0};
    k_stats_u<<<grid_for(n * KP, 256), 256, 0, stream>>>(ua);
    VArgs va{};
    va.p = p; va.K = K; va.KP = KP; va.sparse = sparse; va.zi = zi; va.first_iter = 1; va.want_data = 0; va.stats_only = 1;
    va.ones = ones_stats ? 1 : 0;
    va.beta1 = beta1; va.beta2 = beta2; va.b1 = b1; va.b2 = b2; va.EV = EV; va.ElogV = ElogV; va.prob_S = prob_S; va.Sd = Sd;
    va.colw = colw; va.ev = ev; va.evw = evw; va.EVS = EVS; va.wS = wS; va.guard_next = guard[gcur];
    va.cs_ev = csv; va.cs_elogv = csv + 64; va.cs_evs = csv + 128; va.scal = scal;
    k_mstep_genes<<<grid_for(pK, 256), 256, 0, stream>>>(va);
    if (sparse) {
        CK(cudaMemsetAsync(csv + 128, 0, 64 * 8, stream));
        SparseArgs sa{};
        sa.p = p; sa.K = K; sa.KP = KP; sa.zi = zi; sa.update = 0; sa.sel_bound = opt.sel_bound; sa.EV = EV; sa.ElogV = ElogV;
        sa.colw = colw; sa.prob_S = prob_S; sa.Sd = Sd; sa.prior_S = prior_S; sa.S = S; sa.ev = ev; sa.evw = evw; sa.EVS = EVS;
        sa.wS = wS; sa.cs_evs = csv + 128; sa.allmasked = allmasked; sa.unchanged = unchanged; sa.scal = scal;
        k_sparse_genes<<<grid_for(p, 128), 128, 0, stream>>>(sa);
    }
    k_row_extremes<<<grid_for(n, 256), 256, 0, stream>>>(n, K, KP, ElogU, rowext);
    k_row_extremes<<<grid_for(p, 256), 256, 0, stream>>>(p, K, KP, ElogV, colext);
    launches += 4 + (sparse ? 1 : 0);
    zi_at_init = false;
    if (zi && explicitP) {
        // explicit prob_D: no gene shortcut applies (flag 0, weight 1), the weights of the non-zeros go into the values
        CK(cudaMemsetAsync(flag, 0, p * 4, stream));
        k_fill<<<grid_for(p, 256), 256, 0, stream>>>(colw, p, 1.0);
        k_fill<<<grid_for(p, 256), 256, 0, stream>>>(cz, p, 0.0);
        k_refresh_weights<<<grid_for(pK, 256), 256, 0, stream>>>(p, K, KP, sparse, prob_S, colw, ev, wS, evw);
        if (f32) {
            k_weight_vals_csr<float><<<grid_for(n * 32, 256), 256, 0, stream>>>(n, rowptr, colidx, val_csr, Pexp, valw_csr, scal + S_LGX);
            k_weight_vals_csc<float><<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, n, colptr, rowidx, val_csc, Pexp, valw_csc);
        } else {
            k_weight_vals_csr<double><<<grid_for(n * 32, 256), 256, 0, stream>>>(n, rowptr, colidx, val_csr, Pexp, valw_csr, scal + S_LGX);
            k_weight_vals_csc<double><<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, n, colptr, rowidx, val_csc, Pexp, valw_csc);
        }
        k_dense_zi_cells<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, p, K, KP, Pexp, EU[cur], EVS, PV, ar3 + ar3_csP, ar3 + ar3_zsc);
        allreduce(scal + S_LGX, 1);
        launches += 6;
    } else if (zi) {
        // init_zi_param (zi_gap_factor_model.cpp:98-105): prob_D = (X > 0); realised as pass A with c_j = -inf
        ZiPrepArgs zp{p, 1, prior_D, Lg, cz, colw, flag, scal};
        k_zi_prepare<<<grid_for(p, 256), 256, 0, stream>>>(zp);
        if (opt.dev_flags & 32) {                 // the dense pass with c_j = -inf (kept to cross-check the shortcut below)
            ZiAArgs za{n, p, K, EU[cur], EVS, cz, flag, bitT, PV, ar3 + ar3_csP, ar3 + ar3_zsc, zpackA, nullptr};
            launches += 1 + L->zi_cells(za, opt.dev_flags, stream);
        } else {
            k_zi_init_cells<<<grid_for(n * 32, 256), 256, 0, stream>>>(n, KP, rowptr, colidx, EU[cur], EVS, PV, ar3 + ar3_zsc);
            k_zi_init_colsum<<<grid_for(p, 256), 256, 0, stream>>>(p, colptr, ar3 + ar3_csP);
            launches += 3;
            zi_at_init = true;                         // no packed gene records yet; prob_D is exactly (X > 0)
        }
    }
    allreduce(ar1 + ar1_cseu, 64 + 64 + 4);
    if (zi) allreduce(ar3, ar3_total);
    hyper_update(update_hyper, EU[cur]);
    k_vec_scale<<<1, 64, 0, stream>>>(hyp + 256, csv + 128, 64, 1.0);
    launches += 2;
    read_scal();
    CK(cudaGetLastError());
    last_nodata = h_scal[S_ELBO_NODATA];
    first_iter = true; have_pending = false; sparse_sums_valid = false;
    ++state_ver;
}

// update_hyper_param (gap...cpp:541-548, zi...cpp:370-372) + the scalars of the iteration (k_hyper); when the rows of the
// hyper-parameter matrices differ the Gamma part runs element-wise first (k_hyper_rows)
void pcmf_solver::hyper_update(bool update, const double *EU_now) {
    double *alpha1 = hyp, *alpha2 = hyp + 64, *beta1 = hyp + 128, *beta2 = hyp + 192;
    HyperArgs ha{n_global, p, K, zi, sparse, alpha1, alpha2, beta1, beta2, ar1 + ar1_cseu, ar1 + ar1_cselog, csv, csv + 64, csv + 128,
                 ar3 + ar3_csP, prior_D, ar1 + ar1_usc, ar3 + ar3_zsc, lgx_const, scal};
    ha.update = update ? 1 : 0;
    if (rows_mode) {
        CK(cudaMemsetAsync(rowsg, 0, 2 * sizeof(double), stream));
        k_hyper_rows<<<grid_for(n * KP, 256), 256, 0, stream>>>(n, K, KP, (double)n_global, ha.update, hrow[0], hrow[1], ar1 + ar1_cseu,
                                                                 ar1 + ar1_cselog, EU_now, ElogU, rowsg);
        k_hyper_rows<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(p, K, KP, (double)p, ha.update, hrow[2], hrow[3], csv, csv + 64, EV, ElogV,
                                                                          rowsg + 1);
        allreduce(rowsg, 1);
        launches += 2;
        ha.rows_g = rowsg;
    }
    k_hyper<<<1, 256, 0, stream>>>(ha);
}

// One VEM iteration: model.update_param() (model.cpp:243-246) + the quantities check_convergence needs.
void pcmf_solver::step(bool want_data) {
    const size_t pK = (size_t)p * KP;
    double *alpha1 = hyp, *alpha2 = hyp + 64, *beta1 = hyp + 128, *beta2 = hyp + 192;
    long long *gc = guard[gcur], *gn = guard[1 - gcur];
    const int nxt = (zi || keep_old) ? 1 - cur : cur;
    CK(cudaMemsetAsync(scal, 0, S_COUNT * 8, stream));
    CK(cudaMemsetAsync(csv, 0, 3 * 64 * 8, stream));
    CK(cudaMemsetAsync(ar1 + ar1_tmp, 0, (ar1_core - ar1_tmp) * 8, stream));   // tmpMat, cs_eu, cs_elog, usc
    if (zi) CK(cudaMemsetAsync(ar3, 0, ar3_total * 8, stream));
    k_init_guard<<<1, 32, 0, stream>>>(gn);

    // first iteration after a warm start from an explicit prob_D: the non-zeros carry prob_D_ij x_ij (FP64 values)
    const void *v_csr = explicitP ? (const void *)valw_csr : val_csr, *v_csc = explicitP ? (const void *)valw_csc : val_csc;
    const bool v_f32 = explicitP ? false : f32;
    // 1. E-step, gene side (old statistics): B1 [+ B2 | xlogT for the ELBO data term of the previous iterate]
    begin(PH_GENE_E);
    {
        const bool reuse = sparse && sparse_sums_valid;
        const double *src = (opt.world > 1) ? ar2keep : ar2;
        ColArgs ca{p, K, colptr, rowidx, v_csc, eu, ElogU, ev, ElogV, Sd, colext, allmasked, reuse ? unchanged : nullptr, src + ar2_B1,
                   src + ar2_B2, dbg, gc, ar1 + ar1_B1, ar1 + ar1_B2, ar1 + ar1_xlogT, qcsc, blkptr, nblk};
        const int mode = !want_data ? GENE_B1 : (sparse ? GENE_B1_B2 : GENE_B1_LOGT);
        zero_gene_sums(ar1 + ar1_B1, ar1 + ar1_B2, ar1 + ar1_xlogT, mode);
        if (tiled && !explicitP) tile_gene_pass(mode, ca, v_f32);
        else L->gene_pass(v_f32, mode, ca, std::min<int>(p, sms * 32), stream);
    }
    end(PH_GENE_E, 1);
    // 2. E-step, cell side + U-side M-step (in place)
    begin(PH_ROWS);
    {
        RowArgs ra{};
        ra.n = n; ra.K = K; ra.rowptr = rowptr; ra.colidx = colidx; ra.val = v_csr; ra.ev = ev; ra.evw = evw; ra.ElogV = ElogV;
        ra.Sd = Sd; ra.wS = wS; ra.allmasked = allmasked; ra.dbg = dbg; ra.a1 = a1; ra.a2 = a2; ra.EU_new = EU[nxt]; ra.ElogU = ElogU; ra.eu = eu; ra.rowext = rowext; ra.alpha1 = alpha1;
        ra.alpha2 = alpha2; ra.rate_vec = csv + 128; ra.PV = zi ? PV : nullptr; ra.first_iter = first_iter ? first_mode : 0; ra.a1_old0 = old0[0]; ra.a2_old0 = old0[1]; ra.guard_cur = gc;
        ra.guard_next = gn; ra.csr2csc = csr2csc; ra.qcsc = qcsc; ra.cs_eu = ar1 + ar1_cseu; ra.cs_elog = ar1 + ar1_cselog; ra.usc = ar1 + ar1_usc;
        // colsum(EV o prob_S) of the OLD state: csv+128 is re-accumulated during this iteration, hyp+256 keeps the copy
        ra.rate_vec = hyp + 256;
        if (rows_mode) { ra.alpha1r = hrow[0]; ra.alpha2r = hrow[1]; }
        const int block = 256;
        const int64_t warps = std::min<int64_t>(n, (int64_t)sms * 8 * 8);
        if (tiled && !explicitP) L->tile_cells(TileCellArgs{ra, xt, p}, stream);
        else L->estep_rows(v_f32, ra, (int)((warps + 7) / 8), block, stream);
    }
    end(PH_ROWS, 1);
    // 3. ZI: prob_D^T EU_new with prob_D recomputed from the triple saved when it was defined
    if (zi) {
        begin(PH_ZI_B);
        ZiBArgs zb{};
        zb.n = n; zb.p = p; zb.K = K; zb.EUz = EU[cur]; zb.EUn = EU[nxt]; zb.EVSz = EVS; zb.cz = cz; zb.flag = flag; zb.bitB = bitB;
        zb.tmpMat = ar1 + ar1_tmp; zb.pack = zpackB; zb.pstore = pstore_valid ? pstore : nullptr;
        const int gb = L->zi_genes_block(opt.dev_flags);
        const int gx = (p + gb - 1) / gb;
        // enough (gene block, cell chunk) CTAs for >= 16 full waves at 2 resident CTAs per SM: a 2.1-wave grid measured 29% slower
        int chunks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 1023) / 1024, (int64_t)(sms * 2 * 16 + gx - 1) / gx));
        int64_t rpc = (n + chunks - 1) / chunks; rpc = (rpc + 63) / 64 * 64;
        chunks = (int)((n + rpc - 1) / rpc);
        zb.rows_per_chunk = rpc;
        if (explicitP) {
            k_dense_zi_genes<<<(p + 127) / 128, 128, 0, stream>>>(n, p, K, KP, Pexp, EU[nxt], ar1 + ar1_tmp);
            end(PH_ZI_B, 1);
        } else if (zi_at_init) {           // prob_D = (X > 0): only the non-zeros contribute
            k_zi_init_genes<<<grid_for((int64_t)p * 32, 256), 256, 0, stream>>>(p, K, KP, colptr, rowidx, EU[nxt], ar1 + ar1_tmp);
            end(PH_ZI_B, 1);
        } else if (fp32_mode) {
            ZiTcArgs ta{};
            ta.n = n; ta.p = p; ta.K = K; ta.KP = KP; ta.A1 = EVS; ta.rows = p; ta.streamed = n; ta.bits = bitB; ta.cz = cz; ta.flag = flag;
            ta.tmpMat = ar1 + ar1_tmp;
            end(PH_ZI_B, L->zi_genes_tc(ta, EU[cur], EU[nxt], tcpackB, sms, stream));
        } else {
            end(PH_ZI_B, L->zi_genes(zb, chunks, opt.dev_flags, stream));
        }
    }
    // 4. all-reduce of the gene-side sums + U partials over the row shards
    begin(PH_AR1);
    allreduce(ar1, want_data ? ar1_total : ar1_core);
    end(PH_AR1, 0);
    // 5. V-side M-step
    begin(PH_MSTEP_V);
    if (keep_old) CK(cudaMemcpyAsync(EVold, EV, sizeof(double) * pK, cudaMemcpyDeviceToDevice, stream));
    {
        VArgs va{};
        va.p = p; va.K = K; va.KP = KP; va.sparse = sparse; va.zi = zi; va.first_iter = first_iter ? first_mode : 0; va.b1_old0 = old0[2]; va.b2_old0 = old0[3]; va.want_data = want_data;
        va.stats_only = 0; va.B1 = ar1 + ar1_B1; va.B2 = ar1 + ar1_B2; va.xlogT = ar1 + ar1_xlogT; va.tmpMat = ar1 + ar1_tmp;
        va.cs_eu = ar1 + ar1_cseu; va.beta1 = beta1; va.beta2 = beta2; va.b1 = b1; va.b2 = b2; va.EV = EV; va.ElogV = ElogV;
        va.prob_S = prob_S; va.Sd = Sd; va.colw = colw; va.ev = ev; va.evw = evw; va.EVS = EVS; va.wS = wS; va.guard_next = gn;
        va.cs_ev = csv; va.cs_elogv = csv + 64; va.cs_evs = csv + 128; va.scal = scal;
        if (rows_mode) { va.beta1r = hrow[2]; va.beta2r = hrow[3]; }
        k_mstep_genes<<<grid_for(pK, 256), 256, 0, stream>>>(va);
        k_row_extremes<<<grid_for(p, 256), 256, 0, stream>>>(p, K, KP, ElogV, colext);
    }
    end(PH_MSTEP_V, 2);
    // 6. spike-and-slab probabilities: second non-zero pass with the new statistics and the old mask
    if (sparse) {
        begin(PH_GENE_S);
        ColArgs ca{p, K, colptr, rowidx, v_csc, eu, ElogU, ev, ElogV, Sd, colext, allmasked, nullptr, nullptr, nullptr, dbg, gn, ar2 + ar2_B1,
                   ar2 + ar2_B2, ar2 + ar2_xlogT, qcsc, blkptr, nblk};
        zero_gene_sums(ar2 + ar2_B1, ar2 + ar2_B2, ar2 + ar2_xlogT, GENE_B1_B2);
        if (tiled && !explicitP) tile_gene_pass(GENE_B1_B2, ca, v_f32);
        else L->gene_pass(v_f32, GENE_B1_B2, ca, std::min<int>(p, sms * 32), stream);
        if (opt.world > 1) CK(cudaMemcpyAsync(ar2keep, ar2, 2 * pK * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        allreduce(ar2, 2 * pK);
        sparse_sums_valid = true;
        end(PH_GENE_S, 1);
        begin(PH_SPARSE);
        CK(cudaMemsetAsync(csv + 128, 0, 64 * 8, stream));
        SparseArgs sa{};
        sa.p = p; sa.K = K; sa.KP = KP; sa.zi = zi; sa.update = 1; sa.sel_bound = opt.sel_bound; sa.B1 = ar2 + ar2_B1; sa.B2 = ar2 + ar2_B2;
        sa.tmpMat = ar1 + ar1_tmp; sa.cs_eu = ar1 + ar1_cseu; sa.EV = EV; sa.ElogV = ElogV; sa.colw = colw; sa.prob_S = prob_S; sa.Sd = Sd;
        sa.prior_S = prior_S; sa.S = S; sa.ev = ev; sa.evw = evw; sa.EVS = EVS; sa.wS = wS; sa.cs_evs = csv + 128; sa.allmasked = allmasked; sa.unchanged = unchanged; sa.scal = scal;
        k_sparse_genes<<<grid_for(p, 128), 128, 0, stream>>>(sa);
        end(PH_SPARSE, 1);
    }
    // 7. ZI probabilities: define the new prob_D (cz/flag/colw), its row products, column sums and ELBO terms
    if (zi) {
        begin(PH_ZI_A);
        ZiPrepArgs zp{p, 0, prior_D, Lg, cz, colw, flag, scal};
        k_zi_prepare<<<grid_for(p, 256), 256, 0, stream>>>(zp);
        k_refresh_weights<<<grid_for(pK, 256), 256, 0, stream>>>(p, K, KP, sparse, prob_S, colw, ev, wS, evw);
        ZiAArgs za{n, p, K, EU[nxt], EVS, cz, flag, bitT, PV, ar3 + ar3_csP, ar3 + ar3_zsc, zpackA, pstore};
        int nl;
        if (fp32_mode) {
            ZiTcArgs ta{};
            ta.n = n; ta.p = p; ta.K = K; ta.KP = KP; ta.A1 = EU[nxt]; ta.rows = n; ta.streamed = p; ta.bits = bitT; ta.EU = EU[nxt];
            ta.PV = PV; ta.colsumP = ar3 + ar3_csP; ta.zsc = ar3 + ar3_zsc; ta.cz = cz; ta.flag = flag;
            nl = L->zi_cells_tc(ta, EVS, EVS, tcpackA, sms, stream);
        } else {
            za.want_ll = (want_ll_zero && !(opt.dev_flags & 4)) ? 1 : 0;
            nl = L->zi_cells(za, opt.dev_flags, stream);
        }
        zi_at_init = false;
        pstore_valid = pstore != nullptr;
        allreduce(ar3, ar3_total);
        end(PH_ZI_A, 2 + nl);
    }
    // 8. hyper-parameters + scalars
    begin(PH_HYPER);
    k_vec_scale<<<1, 64, 0, stream>>>(hyp + 256, csv + 128, 64, 1.0);   // keep colsum(EV o prob_S) for the next a2 update
    hyper_update(true, EU[nxt]);
    end(PH_HYPER, 2);
    if (keep_old) {
        if (first_iter) {
            // the old iterate is a1_old / a2_old: Ones / Ones, or the previous multi-init run's last iterate
            if (first_mode == 2) k_ratio<<<grid_for(n * KP, 256), 256, 0, stream>>>(n, K, KP, old0[0], old0[1], EU[cur]);
            else k_ratio<<<grid_for(n * KP, 256), 256, 0, stream>>>(n, K, KP, a1, a1, EU[cur]);
            if (first_mode == 2) k_ratio<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(p, K, KP, old0[2], old0[3], EVold);
            else k_ratio<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(p, K, KP, b1, b1, EVold);
        }
        CK(cudaMemsetAsync(rvbuf, 0, sizeof(double) * 6 * 64 * 64, stream));
        dim3 blk(32, 8);
        dim3 gu((unsigned)std::min<int64_t>((n + 31) / 32, sms * 4), (K * K + 7) / 8), gv((unsigned)std::min<int64_t>((p + 31) / 32, sms * 4), (K * K + 7) / 8);
        k_rv_gram<<<gu, blk, 0, stream>>>(n, K, KP, EU[nxt], EU[cur], rvbuf);
        k_rv_gram<<<gv, blk, 0, stream>>>(p, K, KP, EV, EVold, rvbuf + 3 * 64 * 64);
        allreduce(rvbuf, 3 * (size_t)K * K);
        launches += 2;
    }
    read_scal();
    ++state_ver;
    if (zi && want_ll_zero && !fp32_mode && !(opt.dev_flags & 4)) { ll_zero = h_scal[S_LL_ZERO]; ll_zero_ver = state_ver; }
    if (keep_old) rv_crit = rv_criterion();
    collect_phase_times();
    cur = nxt; gcur = 1 - gcur; first_iter = false;
    if (explicitP) {   // the zero-inflation update of this iteration replaced the explicit matrix by the closed form
        explicitP = false;
        sparse_sums_valid = false;   // the cached gene sums and the stored q carried the explicit weights
        dfree(Pexp); dfree(valw_csr); dfree(valw_csc);
        Pexp = valw_csr = valw_csc = nullptr;
    }
}

// Gene pass when X has its tiled copy.  A full pass streams the tiles; a pass in which most genes copy the sums of the last
// spike-and-slab pass (ColArgs::reuse) walks the CSC segments of the genes that changed instead -- its cost is proportional
// to their number, a tiled CTA would stream every cell block for a single gene -- and leaves their q where the cell pass
// reads it (ColArgs::qmap).  So do the factor widths below 16, where the tiled gene kernel does not pay (kernels_kp.cu).
void pcmf_solver::tile_gene_pass(int mode, ColArgs &ca, bool v_f32_csc) {
    if (ca.reuse || !L->tile_genes) {
        ca.qout = xt.q; ca.qmap = qmap;
        L->gene_pass(v_f32_csc, mode, ca, std::min<int>(p, sms * 32), stream);
        return;
    }
    if (mode == GENE_B1_B2) { k_eu_pairs<<<grid_for(n * KP, 256), 256, 0, stream>>>(n * KP, eu, ElogU, eupairs); ++launches; }
    L->tile_genes(mode, TileGeneArgs{ca, xt, n, 0, eupairs}, sms, stream);
}

// ELBO data term of the current state (gap...cpp:301-313, sparse...cpp:194-212): one gene pass, nothing updated.
double pcmf_solver::data_term_now() {
    const size_t pK = (size_t)p * KP;
    // Sparse models: for genes whose mask row did not change in the last spike-and-slab update the sums of that
    // update's gene pass (same statistics, same mask) are exactly the ones needed -- copy them, recompute the others.
    // The result goes to the B1 / B2 / xlogT slots of the first reduction buffer (free between iterations), so the cached
    // sums stay valid for the next E-step.
    const bool reuse = sparse && sparse_sums_valid;
    const double *src = (opt.world > 1) ? ar2keep : ar2;
    ColArgs ca{p, K, colptr, rowidx, explicitP ? (const void *)valw_csc : val_csc, eu, ElogU, ev, ElogV, Sd, colext, allmasked,
               reuse ? unchanged : nullptr, reuse ? src + ar2_B1 : nullptr, reuse ? src + ar2_B2 : nullptr, dbg, guard[gcur],
               ar1 + ar1_B1, ar1 + ar1_B2, ar1 + ar1_xlogT, qcsc, blkptr, nblk};
    zero_gene_sums(ar1 + ar1_B1, ar1 + ar1_B2, ar1 + ar1_xlogT, sparse ? GENE_B1_B2 : GENE_B1_LOGT);
    if (tiled && !explicitP) tile_gene_pass(sparse ? GENE_B1_B2 : GENE_B1_LOGT, ca, f32);
    else L->gene_pass(explicitP ? false : f32, sparse ? GENE_B1_B2 : GENE_B1_LOGT, ca, std::min<int>(p, sms * 32), stream);
    allreduce(ar1 + ar1_B1, pK);
    allreduce(ar1 + ar1_B2, pK + p);
    CK(cudaMemsetAsync(scal + S_DATA, 0, 8, stream));
    DataArgs da{p, K, KP, sparse, zi, ar1 + ar1_B1, ar1 + ar1_B2, ar1 + ar1_xlogT, ElogV, prob_S, colw, scal};
    k_data_term<<<grid_for(pK, 256), 256, 0, stream>>>(da);
    launches += 2;
    read_scal();
    return h_scal[S_DATA];
}

// Monitors of the current state.  rate = UVt (gap, sparse) or prob_D o UVt (ZI variants):
//   deviance      -2 (pl(X, rate) - pl(X, X))                 gap...cpp:329-333, zi...:218-229, sparse...:231-235
//   exp_deviance  (pl(X, rate) - pl_vec(X, colmean)) / (pl(X, X) - pl_vec(X, colmean))   gap...cpp:336-341 and variants
//   loglikelihood gap...cpp:278-284, zi...:160-167, sparse...:151-157, zi_sparse...:171-178
// pl(X, rate) = sum_nnz x clog(rate) - sum_ij rate - sum lgamma(x + 1): one non-zero pass + the total rate, which is
// the closed form of gap...cpp:298 or, for ZI, sum prob_D o UVt of the last dense pass.
pcmf_solver::Mon pcmf_solver::monitors_now(bool want_loglik) {
    CK(cudaMemsetAsync(mon, 0, 8 * sizeof(double), stream));
    if (explicitP) {
        // prob_D still is the dense matrix of a warm start (it lives until the first zero-inflation update): one plain dense
        // kernel evaluates what the closed-form passes below cannot (zi_gap_factor_model.cpp:160-167 on init_zi_param(prob_D, prior_D))
        if (f32) k_dense_zi_monitor<float><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, p, K, KP, Pexp, EU[cur], EVS, rowptr, colidx, val_csr, bitT, mon);
        else k_dense_zi_monitor<double><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, p, K, KP, Pexp, EU[cur], EVS, rowptr, colidx, val_csr, bitT, mon);
    } else {
        LamArgs la{p, K, colptr, rowidx, val_csc, EU[cur], EVS, zi ? flag : nullptr, mon};
        L->gene_loglam(f32, la, std::min<int>(p, sms * 32), stream);
    }
    launches += 1;
    if (want_loglik) {
        k_gamma_loglike<<<grid_for(n * KP, 256), 256, 0, stream>>>(n, K, KP, EU[cur], a1, a2, mon + 4);
        k_gamma_loglike<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(p, K, KP, EV, b1, b2, mon + 6);
        launches += 2;
        if (zi && !zi_at_init && !explicitP && ll_zero_ver != state_ver) {   // at the initial state prob_D = (X > 0): the zero part is log(1) = 0
            ZiAArgs za{n, p, K, EU[cur], EVS, cz, flag, bitT, PV, ar3 + ar3_csP, ar3 + ar3_zsc, zpackA, nullptr};
            L->zi_loglik(za, mon + 5, stream);
            launches += 1;
        }
    }
    allreduce(mon, 6);    // [0..3] non-zero sums, [4] Gamma term of U, [5] zero part: sums over the local cells; [6] is replicated
    if (explicitP) allreduce(mon + 7, 1);
    double h[8];
    CK(cudaMemcpyAsync(h, mon, sizeof h, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const double sum_rate = zi ? h_scal[S_SUM_PUVT] : h_scal[S_SUMUVT_CLOSED];
    const double r1 = h[0] - sum_rate - lgx_const;
    Mon m;
    m.deviance = -2.0 * (r1 - ll_XX);
    m.exp_dev = (r1 - ll_Xmean) / (ll_XX - ll_Xmean);
    m.loglik = NAN;
    if (want_loglik) {
        double ll;
        if (!zi) ll = h[1] - h_scal[S_SUMUVT_CLOSED] - lgx_const;
        else {
            // zi_poisson_loglike (likelihood.h:197-212): non-zeros log(prob_D) + pl1(x, lambda), zeros log(1 - P + P e^-lambda)
            // (the zero part comes from the cell-side ZI pass of the step when that evaluated it on the way: already all-reduced)
            if (!zi_at_init && !explicitP && ll_zero_ver == state_ver) h[5] = ll_zero;
            ll = h[1] - h[2] - lgx_const + h[5] + h[7];   // h[7]: sum log(prob_D) over the non-zeros (0 unless prob_D is explicit)
            if (h[3] > 0) ll = -INFINITY;    // log(prob_D) with prob_D = 0 on a non-zero entry (prior_D == 0 shortcut)
            ll += h_scal[S_BERND_SELF];
        }
        ll += h[4] + h[6];
        if (sparse) ll += h_scal[S_BERNS_SELF];
        m.loglik = ll;
    }
    return m;
}

// variational_em_algo::order_factor (algorithm_variational_EM.h:189-192): greedy forward selection on the partial deviance
// (model.cpp:137-187), then reorder_factor (gap...cpp:344-366, sparse...cpp:250-275).  One non-zero pass per round
// evaluates every remaining candidate; the total-rate part of each candidate is a sum of per-factor totals.
void pcmf_solver::order_factor_now() {
    // per-factor totals t_k = sum_ij [prob_D_ij] EU_ik EVS_jk
    std::vector<double> tk(64, 0.0), hsum(64);
    CK(cudaMemsetAsync(mon + 8, 0, 128 * sizeof(double), stream));
    if (zi) {
        k_dot_cols<<<grid_for(n * KP, 256), 256, 0, stream>>>(n, K, KP, EU[cur], PV, mon + 8);   // sum_i EU_ik (prob_D EVS)_ik
        allreduce(mon + 8, K);
        CK(cudaMemcpyAsync(tk.data(), mon + 8, sizeof(double) * K, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    } else {
        // closed form: colsum(EU)_k colsum(EVS)_k, both maintained by the iteration (all-reduced) -- gap...cpp:298
        CK(cudaMemcpyAsync(tk.data(), ar1 + ar1_cseu, sizeof(double) * K, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(hsum.data(), csv + 128, sizeof(double) * K, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (int k = 0; k < K; ++k) tk[k] *= hsum[k];
    }
    launches += zi ? 1 : 0;
    std::vector<int> left, chosen;
    for (int k = 0; k < K; ++k) left.push_back(k);
    std::vector<double> mask(64, 0.0), out(128);
    double base_total = 0;
    // scratch: mon + 8: scores (64) | mon + 72: error bounds of the FP32 pass (64) | mon + 136: chosen mask (64)
    const bool try_fast = !(opt.dev_flags & 16384);
    float *EUf = nullptr, *EVSf = nullptr;      // FP32 copies for the fast pass: half the bytes of the per-round gather of E[U] rows
    if (try_fast && K > 1) {
        EUf = dmalloc<float>((size_t)n * KP); EVSf = dmalloc<float>((size_t)p * KP);
        k_to_float<<<grid_for(n * KP, 256), 256, 0, stream>>>(n * KP, EU[cur], EUf);
        k_to_float<<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>((int64_t)p * KP, EVS, EVSf);
        launches += 2;
    }
    for (int k = 0; k < K; ++k) {
        CK(cudaMemcpyAsync(mon + 136, mask.data(), 64 * sizeof(double), cudaMemcpyHostToDevice, stream));
        double val_min = -1; int ind_min = 0; size_t ind_left = 0;
        // fast = 1: logarithms in FP32 with a rigorous bound; the winner must clear every other candidate by both bounds, or the
        // round is repeated exactly (fast = 0) -- a single candidate needs no pass at all to be chosen, but keeps the exact path
        if (left.size() == 1 && try_fast) {       // nothing to compare the last factor with (model.cpp:161-165 picks it whatever its value)
            ind_min = left[0]; ind_left = 0;
            chosen.push_back(ind_min);
            left.erase(left.begin());
            mask[ind_min] = 1.0; base_total += tk[ind_min];
            continue;
        }
        for (int fast = try_fast ? 1 : 0; fast >= 0; --fast) {
            CK(cudaMemsetAsync(mon + 8, 0, 128 * sizeof(double), stream));
            OrderArgs oa{p, K, colptr, rowidx, val_csc, EU[cur], EVS, zi ? flag : nullptr, mon + 136, mon + 8, fast ? mon + 72 : nullptr, EUf, EVSf};
            L->gene_order(f32, oa, std::min<int>(p, sms * 32), stream);
            launches += 1;
            allreduce(mon + 8, 64 + K);
            CK(cudaMemcpyAsync(out.data(), mon + 8, sizeof(double) * (64 + K), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            std::vector<double> res(left.size());
            for (size_t ind = 0; ind < left.size(); ++ind) {
                const int c = left[ind];
                const double r1 = out[c] - (base_total + tk[c]) - lgx_const;
                res[ind] = -2.0 * (r1 - ll_XX);
                if (ind == 0) { val_min = res[ind]; ind_min = c; ind_left = ind; }          // model.cpp:161-165
                if (res[ind] < val_min) { val_min = res[ind]; ind_min = c; ind_left = ind; }     // model.cpp:168-172
            }
            if (!fast) {
                // The sums come out of floating-point atomics, so two candidates the reference sees as an exact tie (identical
                // factors: model.cpp:161-172 keeps the FIRST minimum) differ here by rounding noise of the order 1e-14 of the
                // terms; everything within 1e-12 of the minimum counts as tied and the first of them in the reference's order wins
                for (size_t ind = 0; ind < left.size(); ++ind) {
                    const int c = left[ind];
                    const double scale = std::fabs(out[c]) + std::fabs(base_total + tk[c]) + std::fabs(lgx_const) + std::fabs(ll_XX);
                    if (res[ind] - val_min <= 2e-12 * scale) { ind_min = c; ind_left = ind; break; }
                }
                break;
            }
            // |out_c - exact| <= 6.6e-7 bound_c (k_gene_order<FAST>), res = -2 out + ...: clear margin = 2 x 2 x 6.6e-7 (bound_b + bound_c)
            bool clear = std::isfinite(val_min);
            for (size_t ind = 0; ind < left.size() && clear; ++ind)
                if (ind != ind_left) clear = res[ind] - val_min > 3e-6 * (out[64 + left[ind]] + out[64 + ind_min]);
            if (clear) break;
            order_exact_rounds += 1;
        }
        chosen.push_back(ind_min);
        left.erase(left.begin() + ind_left);
        mask[ind_min] = 1.0; base_total += tk[ind_min];
    }
    if (EUf) { CK(cudaStreamSynchronize(stream)); dfree(EUf); dfree(EVSf); }
    for (int k = 0; k < K; ++k) factor_order[k] = chosen[k];
    // reorder_factor: column c of every K-indexed matrix becomes old column factor_order[c]
    std::vector<int32_t> ord(64, 0);
    for (int k = 0; k < K; ++k) ord[k] = chosen[k];
    CK(cudaMemcpyAsync(d_order, ord.data(), sizeof(int32_t) * 64, cudaMemcpyHostToDevice, stream));
    const int64_t big = std::max<int64_t>(n, p);
    double *tmp = dmalloc<double>((size_t)big * KP);
    auto perm_d = [&](double *A, int64_t rows, int kp) {
        if (!A) return;
        k_permute_cols<double><<<grid_for(rows * kp, 256), 256, 0, stream>>>(rows, K, kp, d_order, A, tmp);
        CK(cudaMemcpyAsync(A, tmp, sizeof(double) * rows * kp, cudaMemcpyDeviceToDevice, stream));
    };
    for (double *A : {a1, a2, EU[0], EU[1], ElogU, eu, PV}) perm_d(A, n, KP);
    for (double *A : {b1, b2, EV, ElogV, ev, evw, EVS, wS, prob_S, Sd}) perm_d(A, p, KP);
    perm_d(hrow[0], n, KP); perm_d(hrow[1], n, KP); perm_d(hrow[2], p, KP); perm_d(hrow[3], p, KP);
    for (int v = 0; v < 5; ++v) perm_d(hyp + 64 * v, 1, 64);
    for (int v = 0; v < 3; ++v) perm_d(csv + 64 * v, 1, 64);
    perm_d(ar1 + ar1_cseu, 1, 64); perm_d(ar1 + ar1_cselog, 1, 64);
    k_permute_cols<int32_t><<<grid_for((int64_t)p * KP, 256), 256, 0, stream>>>(p, K, KP, d_order, S, reinterpret_cast<int32_t *>(tmp));
    CK(cudaMemcpyAsync(S, tmp, sizeof(int32_t) * (size_t)p * KP, cudaMemcpyDeviceToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    dfree(tmp);
    sparse_sums_valid = false;   // the cached gene-pass sums are in the old column order
    CK(cudaGetLastError());
}

// ---- multi-init (wrapper_gap_factor.h:296-350, wrapper_sparse_gap_factor.h:328-387) ---------------------------
std::vector<std::pair<void *, size_t>> pcmf_solver::state_arrays() {
    const size_t nK = (size_t)n * KP * 8, pK = (size_t)p * KP * 8;
    std::vector<std::pair<void *, size_t>> v;
    for (double *A : {a1, a2, EU[0], EU[1], ElogU, eu, PV}) if (A) v.push_back({A, nK});
    for (double *A : {b1, b2, EV, ElogV, ev, evw, EVS, wS, prob_S, Sd}) if (A) v.push_back({A, pK});
    v.push_back({S, (size_t)p * KP * 4}); v.push_back({flag, (size_t)p * 4});
    for (double *A : {prior_S, prior_D, cz, colw}) v.push_back({A, (size_t)p * 8});
    v.push_back({hyp, 5 * 64 * 8}); v.push_back({csv, 3 * 64 * 8});
    if (rows_mode) { v.push_back({hrow[0], nK}); v.push_back({hrow[1], nK}); v.push_back({hrow[2], pK}); v.push_back({hrow[3], pK}); }
    v.push_back({ar1, ar1_total * 8}); v.push_back({ar2, ar2_total * 8}); v.push_back({ar3, ar3_total * 8});
    if (ar2keep) v.push_back({ar2keep, 2 * pK});
    v.push_back({scal, S_COUNT * 8}); v.push_back({guard[0], G_COUNT * 8}); v.push_back({guard[1], G_COUNT * 8});
    v.push_back({rowext, (size_t)n * 16}); v.push_back({colext, (size_t)p * 16});
    v.push_back({allmasked, (size_t)p}); v.push_back({unchanged, (size_t)p});
    return v;
}
void pcmf_solver::save_state(Saved &sv) {
    if (sv.copies.empty()) {
        sv.arrays = state_arrays();
        for (auto &a : sv.arrays) sv.copies.push_back(dmalloc<unsigned char>(a.second));
    }
    for (size_t i = 0; i < sv.arrays.size(); ++i)
        CK(cudaMemcpyAsync(sv.copies[i], sv.arrays[i].first, sv.arrays[i].second, cudaMemcpyDeviceToDevice, stream));
    sv.cur = cur; sv.gcur = gcur; sv.first_iter = first_iter; sv.sparse_sums_valid = sparse_sums_valid; sv.have_pending = have_pending;
    sv.last_nodata = last_nodata; sv.h.assign(h_scal, h_scal + S_COUNT); sv.iter = iter; sv.first_mode = first_mode;
    CK(cudaStreamSynchronize(stream));
}
void pcmf_solver::restore_state(const Saved &sv) {
    for (size_t i = 0; i < sv.arrays.size(); ++i)
        CK(cudaMemcpyAsync(sv.arrays[i].first, sv.copies[i], sv.arrays[i].second, cudaMemcpyDeviceToDevice, stream));
    cur = sv.cur; gcur = sv.gcur; first_iter = sv.first_iter; sparse_sums_valid = sv.sparse_sums_valid; have_pending = sv.have_pending;
    last_nodata = sv.last_nodata; std::copy(sv.h.begin(), sv.h.end(), h_scal); iter = sv.iter; first_mode = sv.first_mode;
    ++state_ver;
    sparse_sums_valid = false;   // the stored q of the non-zeros belong to the last run, not to the restored one
    pstore_valid = false;        // and so does the stored prob_D
    CK(cudaStreamSynchronize(stream));
}

// Runs `ninit` initialisations for `iter_init` iterations each and keeps the one with the SMALLEST loss (sic:
// `tmp_loss < best_loss` although the loss is the ELBO, wrapper_gap_factor.h:342), leaving the solver in the kept
// run's state at iteration iter_init so that pcmf_optimize continues from there (my_model = tmp_model, :339-345).
// Reference details kept: the per-iteration vectors of the kept run stay in `res` for iterations < iter_init; the
// first iteration of runs after the first measures its gap against the previous run's last iterate (the `old` copies
// are never reset); with explicit U, V every run perturbs them by Uniform(+-max(U, V)/K) (perturb_now); with explicit
// a1..b2 the reference calls the same perturbation with U and V unset (undefined behaviour there) -- here those values
// are re-used unperturbed.
void pcmf_solver::multi_init(pcmf_result *res) {
    const pcmf_options saved_opt = opt;
    opt.iter_max = saved_opt.iter_init; opt.iter_min = saved_opt.iter_init; opt.epsilon = 1e-6;   // set_iter(iter_init, iter_init, 1e-6)
    opt.progress = nullptr;                                                                     // set_verbosity(false)
    const size_t nK = (size_t)n * KP * 8, pK = (size_t)p * KP * 8;
    Saved best;
    std::vector<std::vector<double>> best_vec(6);
    double *vecs[6] = {res->conv_crit, res->norm_gap, res->abs_gap, res->optim_criterion, res->loglikelihood, res->deviance};
    double best_loss = 0;
    for (int run = 0; run < saved_opt.ninit; ++run) {
        if (run > 0) {
            // the reference's m_*old members still hold the previous run's last iterate
            if (!old0[0]) { old0[0] = dmalloc<double>(nK / 8); old0[1] = dmalloc<double>(nK / 8); old0[2] = dmalloc<double>(pK / 8); old0[3] = dmalloc<double>(pK / 8); }
            CK(cudaMemcpyAsync(old0[0], a1, nK, cudaMemcpyDeviceToDevice, stream)); CK(cudaMemcpyAsync(old0[1], a2, nK, cudaMemcpyDeviceToDevice, stream));
            CK(cudaMemcpyAsync(old0[2], b1, pK, cudaMemcpyDeviceToDevice, stream)); CK(cudaMemcpyAsync(old0[3], b2, pK, cudaMemcpyDeviceToDevice, stream));
            first_mode = 2;
            init_state(&init_copy, true);   // init_sparse_param / init_zi_param / init_random with the continuing stream
        }
        if (init_copy.U) perturb_now();  // every run perturbs user-supplied factors (wrapper_gap_factor.h:312-314)
        iter = 0;                        // set_current_iter(0)
        char err[512] = {0};
        const int rc = pcmf_optimize(this, res, err, sizeof err);
        if (rc != PCMF_OK) fail(rc, "%s", err);
        const double loss = res->loss;
        if (run == 0 || loss < best_loss) {
            best_loss = loss;
            save_state(best);
            for (int v = 0; v < 6; ++v) if (vecs[v]) best_vec[v].assign(vecs[v], vecs[v] + saved_opt.iter_init);
        }
    }
    restore_state(best);
    for (int v = 0; v < 6; ++v) if (vecs[v]) std::copy(best_vec[v].begin(), best_vec[v].end(), vecs[v]);
    for (void *q : best.copies) dfree(q);
    opt = saved_opt;                     // set_iter(iter_min, iter_max, epsilon); set_verbosity(verbose)
    first_mode = 1;
}

// tmp_model.perturb_param(rng, max(U.maxCoeff(), V.maxCoeff()) / K) (wrapper_gap_factor.h:314, gap_factor_model.cpp:177-192)
void pcmf_solver::perturb_now() {
    if (opt.world > 1) fail(PCMF_EUNSUPPORTED, "multi-init from user-supplied U, V is not available on cell-sharded runs");
    double mx = -1e300;
    for (size_t e = 0; e < (size_t)n * K; ++e) mx = std::max(mx, init_copy.U[e]);
    for (size_t e = 0; e < (size_t)p * K; ++e) mx = std::max(mx, init_copy.V[e]);
    const double level = mx / K;
    std::mt19937 g = rng; g.discard(rng_consumed);
    auto add = [&](int64_t rows, double *dst) {
        std::vector<uint32_t> raw((size_t)rows * K);
        for (size_t e = 0; e < raw.size(); ++e) raw[e] = (uint32_t)g();
        uint32_t *d_raw = dmalloc<uint32_t>(raw.size());
        CK(cudaMemcpyAsync(d_raw, raw.data(), raw.size() * 4, cudaMemcpyHostToDevice, stream));
        k_perturb_rows<<<grid_for(rows * K, 256), 256, 0, stream>>>(rows, K, KP, d_raw, level, dst);
        CK(cudaStreamSynchronize(stream));
        dfree(d_raw);
    };
    add(n, a1);
    add(p, b1);
    rng_consumed += (unsigned long long)(n + p) * K;
    run_init_stats();
}

// gap_factor_model::custom_conv_criterion (gap_factor_model.cpp:265-269): 1 - min(RV(EU, EU_old), RV(EV, EV_old))
double pcmf_solver::rv_criterion() {
    std::vector<double> h(6 * 64 * 64);
    CK(cudaMemcpyAsync(h.data(), rvbuf, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    auto rv = [&](const double *g) {
        double num = 0, da = 0, db = 0;
        for (int e = 0; e < K * K; ++e) { num += g[e] * g[e]; da += g[K * K + e] * g[K * K + e]; db += g[2 * K * K + e] * g[2 * K * K + e]; }
        return num / std::sqrt(da * db);
    };
    return 1.0 - std::min(rv(h.data()), rv(h.data() + 3 * 64 * 64));
}

// ================================================================================================================
// C ABI
// ================================================================================================================
namespace {
int guarded(char *errbuf, size_t errlen, const std::function<void()> &fn) {
    try {
        fn();
        return PCMF_OK;
    } catch (const Err &e) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "%s", e.what());
        return e.code;
    } catch (const std::exception &e) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "%s", e.what());
        return PCMF_EINVAL;
    }
}

void check_options(const pcmf_problem *prob, const pcmf_options *opt) {
    if (!prob || !opt) fail(PCMF_EINVAL, "null problem/options");
    const int64_t nglob = prob->n_global > 0 ? prob->n_global : prob->n_local;
    // wrapper_gap_factor.h:210-213
    if (opt->K > std::min<int64_t>(nglob, prob->p))
        fail(PCMF_EINVAL, "Input paramter K should be lower than min(n,p) where n and p are the number of rows and columns in input matrix X");
    if (opt->K < 1) fail(PCMF_EINVAL, "K must be >= 1");
    if (padded_K(opt->K) < 0) fail(PCMF_EUNSUPPORTED, "K = %d exceeds the largest compiled factor width (64)", opt->K);
    // wrapper_sparse_gap_factor.h:236
    if (opt->sparsity && (opt->sel_bound < 0 || opt->sel_bound > 1))
        fail(PCMF_EINVAL, "Input parameter sel_bound should be a real value between 0 and 1");
    // multi-init writes the kept run's first iter_init per-iteration entries into the caller's iter_max-long vectors
    // (wrapper_gap_factor.h:296-350: the reference indexes past the end of its own vectors in that case)
    if (opt->ninit > 1 && opt->iter_init > opt->iter_max)
        fail(PCMF_EINVAL, "with ninit > 1, iter_init (%d) must not exceed iter_max (%d)", opt->iter_init, opt->iter_max);
    if (opt->precision < 0 || opt->precision > 1) fail(PCMF_EINVAL, "precision must be 0 (FP64) or 1 (FP32)");
    // algorithm.h:427
    if (opt->conv_mode < 0 || opt->conv_mode > 2) fail(PCMF_EINVAL, "`conv_mode` parameter should take value in {0,1,2}");
}
// argument consistency of the warm-start values, with the reference's messages (wrapper_gap_factor.h:215-258); host only
void check_init(const pcmf_init *in) {
    if (!in) return;
    if ((in->U != nullptr) != (in->V != nullptr))
        fail(PCMF_EINVAL, "You should supply initial values for both U and V or for neither of them");
    const int nab = (in->a1 != nullptr) + (in->a2 != nullptr) + (in->b1 != nullptr) + (in->b2 != nullptr);
    if (nab != 0 && nab != 4) fail(PCMF_EINVAL, "You should supply initial values for all a1, a2, b1, b2 or for none of them");
    const int nhy = (in->alpha1 != nullptr) + (in->alpha2 != nullptr) + (in->beta1 != nullptr) + (in->beta2 != nullptr);
    if (nhy != 0 && nhy != 4)
        fail(PCMF_EINVAL, "You should supply initial values for all alpha1, alpha2, beta1, beta2 or for none of them");
    if ((in->prob_D != nullptr) != (in->prior_D != nullptr))
        fail(PCMF_EINVAL, "You should supply initial values for both prob_D and prior_D or for neither of them");
}
}  // namespace

extern "C" int pcmf_fit_multi(const pcmf_problem *prob, const pcmf_options *opt, const pcmf_init *init, pcmf_result *res,
                              char *errbuf, size_t errlen);     // multi.cu

extern "C" {

const char *pcmf_version(void) { return "pcmf_b200 0.2.0 (sm_100a; fp64 mode, fp32 mode on tcgen05)"; }

int pcmf_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}

int pcmf_create(const pcmf_problem *prob, const pcmf_options *opt, const pcmf_init *init, pcmf_solver **out, char *errbuf,
                size_t errlen) {
    pcmf_solver *s = nullptr;
    int rc = guarded(errbuf, errlen, [&] {
        if (!out) fail(PCMF_EINVAL, "null output handle");
        check_options(prob, opt);
        check_init(init);
        if (opt->ngpus > 1) fail(PCMF_EUNSUPPORTED, "ngpus > 1 is driven by pcmf_fit (one call, whole matrices); the step-wise API is one device per solver");
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0)
            fail(PCMF_ECUDA, "no CUDA device available: libpcmf_b200 has no CPU fallback");
        s = new pcmf_solver();
        s->opt = *opt;
        if (s->opt.world < 1) s->opt.world = 1;
        if (s->opt.progress_every <= 0) s->opt.progress_every = 10;
        s->device = opt->device;
        CK(cudaSetDevice(s->device));
        cudaDeviceProp dp;
        CK(cudaGetDeviceProperties(&dp, s->device));
        s->sms = dp.multiProcessorCount;
        if (opt->stream) { s->stream = (cudaStream_t)opt->stream; s->own_stream = false; }
        else { CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->own_stream = true; }
        s->K = opt->K; s->KP = padded_K(opt->K); s->zi = opt->zero_inflation != 0; s->sparse = opt->sparsity != 0;
        s->fp32_mode = opt->precision == 1;
        s->L = get_launchers(s->KP);
        if (!s->L) fail(PCMF_EUNSUPPORTED, "no kernels compiled for padded K = %d", s->KP);
        StageTimer tm(s->stream);
        s->rng.seed(s->opt.has_seed ? s->opt.seed : (uint32_t)std::chrono::system_clock::now().time_since_epoch().count());
        if (prob && prob->n_local > 0 && prob->p > 0)
            s->start_draw_ahead(init, prob->n_local, prob->n_global > 0 ? prob->n_global : prob->n_local, prob->row_offset, prob->p);
        s->setup_problem(*prob);
        tm.mark("setup_problem (total)");
        s->alloc_state();
        tm.mark("alloc_state");
        if (init) s->init_copy = *init;
        s->init_state(init);
        tm.mark("init_state");
        if (getenv("PCMF_TIMING"))
            fprintf(stderr, "[pcmf timing] device allocations not served by the block cache so far: %zu (%.1f MB, %.1f ms in cudaMalloc)\n",
                    block_cache().misses, block_cache().miss_bytes / 1e6, block_cache().miss_seconds * 1e3);
    });
    if (rc != PCMF_OK) { delete s; s = nullptr; }
    if (out) *out = s;
    return rc;
}

void pcmf_destroy(pcmf_solver *s) {
    const double t0 = now_s();
    delete s;
    if (getenv("PCMF_TIMING")) fprintf(stderr, "[pcmf timing] %-28s %8.1f ms\n", "destroy", (now_s() - t0) * 1e3);
}

int pcmf_iterate(pcmf_solver *s, int niter, double *conv_crit_out, double *elbo_out, char *errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s) fail(PCMF_EINVAL, "null solver");
        CK(cudaSetDevice(s->device));
        for (int it = 0; it < niter; ++it) {
            // the data term of the ELBO of iterate t is a by-product of the E-step of iterate t+1 (same statistics)
            const bool want = elbo_out && it > 0;
            s->step(want);
            if (want) elbo_out[it - 1] = s->last_nodata + s->h_scal[S_DATA];
            const double crit = s->opt.conv_mode == 0 ? s->h_scal[S_ABS_GAP] : (s->opt.conv_mode == 1 ? s->h_scal[S_NORM_GAP] : s->rv_crit);
            if (conv_crit_out) conv_crit_out[it] = crit;
            s->last_nodata = s->h_scal[S_ELBO_NODATA];
            s->iter++;
        }
        if (elbo_out && niter > 0) elbo_out[niter - 1] = s->last_nodata + s->data_term_now();
    });
}

int pcmf_optimize(pcmf_solver *s, pcmf_result *res, char *errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s || !res) fail(PCMF_EINVAL, "null solver/result");
        CK(cudaSetDevice(s->device));
        const pcmf_options &o = s->opt;
        const double t0 = now_s();
        s->converged = false;
        s->nb_iter_stab = 0;     // local to optimize() in the reference (algorithm.h:161)
        s->want_ll_zero = o.monitor == 1 && res->loglikelihood != nullptr;
        // algorithm.h:158-210
        int pending = -1;   // iteration whose ELBO still lacks its data term
        while (s->iter < o.iter_max && !s->converged) {
            if (o.progress && (s->iter % o.progress_every == 0)) {
                const double c = s->iter > 0 && res->conv_crit ? res->conv_crit[s->iter - 1] : NAN;
                if (o.progress(o.progress_ctx, s->iter, c) != 0) fail(PCMF_EINTERRUPT, "interrupted at iteration %d", s->iter);
            }
            const bool want = o.monitor && pending >= 0;
            s->step(want);
            const int it = s->iter;
            if (want && res->optim_criterion) res->optim_criterion[pending] = s->last_nodata + s->h_scal[S_DATA];
            s->last_nodata = s->h_scal[S_ELBO_NODATA];
            pending = it;
            if (o.monitor == 1 && (res->loglikelihood || res->deviance)) {   // algorithm.h:391-397
                const pcmf_solver::Mon m = s->monitors_now(res->loglikelihood != nullptr);
                if (res->loglikelihood) res->loglikelihood[it] = m.loglik;
                if (res->deviance) res->deviance[it] = m.deviance;
            }
            const double ag = s->h_scal[S_ABS_GAP], ng = s->h_scal[S_NORM_GAP];
            const double crit = o.conv_mode == 0 ? ag : (o.conv_mode == 1 ? ng : s->rv_crit);   // algorithm.h:417-425
            if (res->conv_crit) res->conv_crit[it] = crit;
            if (res->abs_gap) res->abs_gap[it] = ag;
            if (res->norm_gap) res->norm_gap[it] = ng;
            if (std::fabs(crit) < o.epsilon) s->nb_iter_stab++; else s->nb_iter_stab = 0;         // algorithm.h:440-444
            if (s->nb_iter_stab > o.additional_iter && it > o.iter_min) s->converged = true;     // algorithm.h:447
            s->iter++;
        }
        // optim_criterion() after the loop (algorithm.h:209)
        const double data = s->data_term_now();
        res->loss = s->last_nodata + data;
        if (o.monitor && pending >= 0 && res->optim_criterion) res->optim_criterion[pending] = res->loss;
        res->nb_iter = s->iter;
        res->converged = s->converged ? 1 : 0;
        res->seconds_iter = now_s() - t0;
    });
}

int pcmf_elbo(pcmf_solver *s, double *elbo_out, double *terms, char *errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s) fail(PCMF_EINVAL, "null solver");
        CK(cudaSetDevice(s->device));
        // the no-data part was produced by the last k_hyper launch; h_scal still mirrors it unless data_term_now overwrote S_DATA only
        const double data = s->data_term_now();
        const double *h = s->h_scal;
        if (elbo_out) *elbo_out = s->last_nodata + data;
        if (terms) {
            terms[0] = h[S_GAMMAU]; terms[1] = h[S_ENTU]; terms[2] = h[S_GAMMAV]; terms[3] = h[S_ENTV];
            terms[4] = h[S_BERNS_PRIOR]; terms[5] = h[S_BERNS_SELF]; terms[6] = h[S_BERND_PRIOR]; terms[7] = h[S_BERND_SELF];
            terms[8] = s->zi ? h[S_SUM_PUVT] : h[S_SUMUVT_CLOSED]; terms[9] = data; terms[10] = h[S_LGX];
        }
    });
}

int pcmf_set_variational(pcmf_solver *s, const double *a1, const double *a2, const double *b1, const double *b2, char *errbuf,
                         size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s || !a1 || !a2 || !b1 || !b2) fail(PCMF_EINVAL, "null argument");
        CK(cudaSetDevice(s->device));
        s->upload_colmajor(a1, s->n, s->a1, 1.0); s->upload_colmajor(a2, s->n, s->a2, 1.0);
        s->upload_colmajor(b1, s->p, s->b1, 1.0); s->upload_colmajor(b2, s->p, s->b2, 1.0);
        s->run_init_stats();
    });
}

int pcmf_order_factor(pcmf_solver *s, char *errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s) fail(PCMF_EINVAL, "null solver");
        CK(cudaSetDevice(s->device));
        s->order_factor_now();
    });
}

int pcmf_monitors(pcmf_solver *s, double *loglikelihood, double *deviance, double *exp_deviance, char *errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s) fail(PCMF_EINVAL, "null solver");
        CK(cudaSetDevice(s->device));
        const pcmf_solver::Mon m = s->monitors_now(loglikelihood != nullptr);
        if (loglikelihood) *loglikelihood = m.loglik;
        if (deviance) *deviance = m.deviance;
        if (exp_deviance) *exp_deviance = m.exp_dev;
    });
}

int pcmf_multi_init(pcmf_solver *s, pcmf_result *res, char *errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s || !res) fail(PCMF_EINVAL, "null solver/result");
        CK(cudaSetDevice(s->device));
        s->multi_init(res);
    });
}

int pcmf_get_output(pcmf_solver *s, pcmf_result *res, char *errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        if (!s || !res) fail(PCMF_EINVAL, "null solver/result");
        CK(cudaSetDevice(s->device));
        const int K = s->K; const int64_t n = s->n; const int p = s->p;
        if (res->EU) s->download_colmajor(s->EU[s->cur], n, res->EU);
        if (res->ElogU) s->download_colmajor(s->ElogU, n, res->ElogU);
        if (res->a1) s->download_colmajor(s->a1, n, res->a1);
        if (res->a2) s->download_colmajor(s->a2, n, res->a2);
        if (res->EV) s->download_colmajor(s->EV, p, res->EV);
        if (res->ElogV) s->download_colmajor(s->ElogV, p, res->ElogV);
        if (res->b1) s->download_colmajor(s->b1, p, res->b1);
        if (res->b2) s->download_colmajor(s->b2, p, res->b2);
        std::vector<double> h(4 * 64);
        CK(cudaMemcpy(h.data(), s->hyp, sizeof(double) * 4 * 64, cudaMemcpyDeviceToHost));
        auto rep = [&](double *dst, int64_t rows, const double *vec) {   // hyper_params are returned replicated (gap...cpp:379-382)
            if (!dst) return;
            HostStager::par_for(sizeof(double) * (size_t)rows * K, host_stager().nthreads(), [=](size_t lo, size_t hi) {
                size_t e = lo / sizeof(double); const size_t e1 = hi / sizeof(double);
                while (e < e1) {                       // column k of the n x K result holds vec[k]
                    const size_t k = e / (size_t)rows, stop = std::min(e1, (k + 1) * (size_t)rows);
                    std::fill(dst + e, dst + stop, vec[k]);
                    e = stop;
                }
            });
        };
        if (s->rows_mode) {
            if (res->alpha1) s->download_colmajor(s->hrow[0], n, res->alpha1);
            if (res->alpha2) s->download_colmajor(s->hrow[1], n, res->alpha2);
            if (res->beta1) s->download_colmajor(s->hrow[2], p, res->beta1);
            if (res->beta2) s->download_colmajor(s->hrow[3], p, res->beta2);
        } else {
            rep(res->alpha1, n, h.data()); rep(res->alpha2, n, h.data() + 64); rep(res->beta1, p, h.data() + 128); rep(res->beta2, p, h.data() + 192);
        }
        if (s->sparse) {
            if (res->prob_S) s->download_colmajor(s->prob_S, p, res->prob_S);
            if (res->prior_prob_S) CK(cudaMemcpy(res->prior_prob_S, s->prior_S, sizeof(double) * p, cudaMemcpyDeviceToHost));
            if (res->S) {
                int32_t *tmp = dmalloc<int32_t>((size_t)p * K);
                k_rowpad_to_colmajor_i32<<<s->grid_for((int64_t)p * K, 256), 256, 0, s->stream>>>(p, K, s->KP, s->S, tmp);
                CK(cudaMemcpyAsync(res->S, tmp, sizeof(int32_t) * p * K, cudaMemcpyDeviceToHost, s->stream));
                CK(cudaStreamSynchronize(s->stream));
                dfree(tmp);
            }
        }
        if (s->zi) {
            if (res->prior_prob_D) CK(cudaMemcpy(res->prior_prob_D, s->prior_D, sizeof(double) * p, cudaMemcpyDeviceToHost));
            if (res->freq_D) {
                double *tmp = dmalloc<double>(p);
                if (s->freq_ones) k_fill<<<s->grid_for(p, 256), 256, 0, s->stream>>>(tmp, p, 1.0);   // never set when prob_D is given (zi...cpp:57, :110-114)
                else k_vec_scale<<<s->grid_for(p, 256), 256, 0, s->stream>>>(tmp, s->colnnz, p, 1.0 / (double)s->n_global);
                CK(cudaMemcpyAsync(res->freq_D, tmp, sizeof(double) * p, cudaMemcpyDeviceToHost, s->stream));
                CK(cudaStreamSynchronize(s->stream));
                dfree(tmp);
            }
            if (res->prob_D && s->explicitP) {
                host_stager().d2h(res->prob_D, s->Pexp, sizeof(double) * n * p, s->stream);
            } else if (res->prob_D) {
                // prob_D is defined by the triple saved at the last ZI update: EU[cur] (the EU of that update), EVS, cz/flag
                double *tmp = dmalloc<double>((size_t)n * p);
                ProbDArgs pa{n, p, K, s->KP, s->EU[s->cur], s->EVS, s->cz, s->flag, s->bitT, tmp};
                k_prob_d_dense<<<s->grid_for(n * (int64_t)p, 256), 256, 0, s->stream>>>(pa);
                host_stager().d2h(res->prob_D, tmp, sizeof(double) * n * p, s->stream);
                dfree(tmp);
            }
        }
        if (res->factor_order) for (int k = 0; k < K; ++k) res->factor_order[k] = s->factor_order[k];
        res->nb_iter = s->iter;
        res->converged = s->converged ? 1 : 0;
        CK(cudaGetLastError());
    });
}

int pcmf_phase_times(pcmf_solver *s, const char **names, double *ms, int64_t *launches, int cap) {
    if (!s) return 0;
    int c = 0;
    for (int i = 0; i < PH_COUNT && c < cap; ++i) {
        if (names) names[c] = kPhaseNames[i];
        if (ms) ms[c] = s->ph_ms[i];
        if (launches) launches[c] = s->ph_launch[i];
        ++c;
    }
    // diagnostics: number of non-zero entries that had to be recomputed with the exact (guarded) multinomial formula
    unsigned long long h[4] = {0, 0, 0, 0};
    cudaSetDevice(s->device);
    cudaMemcpy(h, s->dbg, 32, cudaMemcpyDeviceToHost);
    static const char *dn[2] = {"exact_fallback_entries_cells", "exact_fallback_entries_genes"};
    for (int i = 0; i < 2 && c < cap; ++i) {
        if (names) names[c] = dn[i];
        if (ms) ms[c] = 0.0;
        if (launches) launches[c] = (int64_t)h[i];
        ++c;
    }
    // bytes of prob_D kept between the cell pass and the next gene pass (0: the gene pass recomputes it)
    if (c < cap) {
        if (names) names[c] = "stored_prob_D_bytes";
        if (ms) ms[c] = 0.0;
        if (launches) launches[c] = s->pstore ? (int64_t)((s->n + 127) / 128 * 16) * (int64_t)((s->p + 7) / 8) * 256 : 0;
        ++c;
    }
    // rounds of the factor ordering the FP32 pass could not decide (repeated exactly)
    if (c < cap) {
        if (names) names[c] = "order_factor_exact_rounds";
        if (ms) ms[c] = 0.0;
        if (launches) launches[c] = s->order_exact_rounds;
        ++c;
    }
    // cell blocks the gene passes walk (1: whole genes in one go)
    if (c < cap) {
        if (names) names[c] = "gene_pass_cell_blocks";
        if (ms) ms[c] = 0.0;
        if (launches) launches[c] = s->nblk;
        ++c;
    }
    // 1: the non-zero passes run on the 2-D tiled copy of X (xtile.cuh), 0: on CSR / CSC
    if (c < cap) {
        if (names) names[c] = "tiled_nonzero_passes";
        if (ms) ms[c] = 0.0;
        if (launches) launches[c] = s->tiled ? 1 : 0;
        ++c;
    }
    return c;
}

int64_t pcmf_launch_count(pcmf_solver *s) { return s ? s->launches : 0; }

int pcmf_fit(const pcmf_problem *prob, const pcmf_options *opt, const pcmf_init *init, pcmf_result *res, char *errbuf, size_t errlen) {
    if (!res) { if (errbuf && errlen) snprintf(errbuf, errlen, "null result"); return PCMF_EINVAL; }
    if (prob && opt && opt->ngpus > 1) {              // the library shards the cells over the devices itself (multi.cu)
        int rc0 = guarded(errbuf, errlen, [&] { check_options(prob, opt); check_init(init); });
        return rc0 != PCMF_OK ? rc0 : pcmf_fit_multi(prob, opt, init, res, errbuf, errlen);
    }
    const double t0 = now_s();
    pcmf_solver *s = nullptr;
    int rc = pcmf_create(prob, opt, init, &s, errbuf, errlen);
    if (rc != PCMF_OK) return rc;
    res->seconds_setup = now_s() - t0;
    if (opt->ninit > 1) rc = pcmf_multi_init(s, res, errbuf, errlen);                                 // wrapper_gap_factor.h:296-350
    if (rc == PCMF_OK) rc = pcmf_optimize(s, res, errbuf, errlen);                                      // :379
    if (rc == PCMF_OK && opt->reorder_factor) rc = pcmf_order_factor(s, errbuf, errlen);               // :382-385
    if (rc == PCMF_OK) rc = pcmf_monitors(s, nullptr, nullptr, &res->exp_dev, errbuf, errlen);          // algorithm.h:367
    if (rc == PCMF_OK) rc = pcmf_get_output(s, res, errbuf, errlen);                                    // :389-392
    pcmf_destroy(s);
    res->seconds_total = now_s() - t0;
    return rc;
}

int pcmf_release_cached_memory(void) { block_cache().release_all(); host_stager().release(); return PCMF_OK; }

int pcmf_device_free(int device, void *ptr) {
    if (cudaSetDevice(device) != cudaSuccess) return PCMF_ECUDA;
    return cudaFree(ptr) == cudaSuccess ? PCMF_OK : PCMF_ECUDA;
}
int pcmf_device_to_host(int device, void *dst_host, const void *src_dev, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return PCMF_ECUDA;
    return cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? PCMF_OK : PCMF_ECUDA;
}

}  // extern "C"
