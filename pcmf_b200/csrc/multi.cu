// pcmf_options.ngpus > 1: the library shards the cells over the devices of the box ITSELF.
//
// The reference picks its parallelism from the call -- `ncores` -> omp_set_num_threads (wrapper_gap_factor.h:271-274), one
// OpenMP reduction over thread-private accumulators per E-step (gap_factor_model.cpp:474-478).  The GPU counterpart has to be
// reachable the same way: one pcmf_fit call with whole matrices in and out, no launcher, no callback to write.  This file is
// that driver, written on top of the public C ABI only: one host thread per device (the calling thread is rank 0, so the
// progress / interrupt hook still runs on the caller's thread), contiguous nnz-balanced row ranges, every rank runs the
// ordinary single-device pcmf_fit on its rows with rank / world set and an all-reduce callback that is ncclAllReduce on a
// communicator created here (ncclCommInitAll).  NCCL is loaded at run time (dlopen of libnccl.so.2 -- the system library, or
// the copy a host process such as PyTorch already mapped), so single-GPU users need no NCCL at all.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pcmf_b200.h"

namespace {

struct Nccl {
    typedef void *comm_t;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *handle = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);     // already mapped by the host process?
            if (!handle) handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle && getenv("PCMF_NCCL_LIB")) handle = dlopen(getenv("PCMF_NCCL_LIB"), RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { error = std::string("cannot load libnccl.so.2 (set PCMF_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return false; }
        CommInitAll = (int (*)(comm_t *, int, const int *))dlsym(handle, "ncclCommInitAll");
        AllReduce = (int (*)(const void *, void *, size_t, int, int, comm_t, cudaStream_t))dlsym(handle, "ncclAllReduce");
        CommDestroy = (int (*)(comm_t))dlsym(handle, "ncclCommDestroy");
        GetErrorString = (const char *(*)(int))dlsym(handle, "ncclGetErrorString");
        if (!CommInitAll || !AllReduce || !CommDestroy) { error = "libnccl.so.2 lacks ncclCommInitAll / ncclAllReduce / ncclCommDestroy"; return false; }
        return true;
    }
};
Nccl &nccl() { static Nccl n; return n; }
constexpr int kNcclDouble = 8, kNcclSum = 0;   // ncclFloat64, ncclSum (nccl.h enums, stable since NCCL 2.0)

// every rank reaches the progress hook at the same iterations (the stopping rule is a function of all-reduced scalars):
// rank 0 asks the caller's hook, everybody leaves with its answer
struct Shared {
    int world = 1;
    pcmf_progress_fn user = nullptr; void *user_ctx = nullptr;
    std::mutex mu; std::condition_variable cv;
    int arrived = 0; long generation = 0; int decision = 0;
    std::atomic<int> failed{0};
};
struct RankCtx { Shared *sh; int rank; Nccl::comm_t comm; };

int progress_all(void *ctx, int iter, double crit) {
    RankCtx *rc = static_cast<RankCtx *>(ctx);
    Shared *sh = rc->sh;
    int mine = 0;
    if (rc->rank == 0 && sh->user) mine = sh->user(sh->user_ctx, iter, crit);
    std::unique_lock<std::mutex> lk(sh->mu);
    if (rc->rank == 0) sh->decision = mine;
    const long gen = sh->generation;
    if (++sh->arrived == sh->world) { sh->arrived = 0; ++sh->generation; sh->cv.notify_all(); }
    else sh->cv.wait(lk, [&] { return sh->generation != gen; });
    return sh->decision;
}
int allreduce_nccl(void *ctx, void *devptr, size_t count, void *stream) {
    RankCtx *rc = static_cast<RankCtx *>(ctx);
    return nccl().AllReduce(devptr, devptr, count, kNcclDouble, kNcclSum, rc->comm, (cudaStream_t)stream);
}

// rows [r0, r1) of a column-major rows x cols matrix
template <typename T>
std::vector<T> slice_rows(const T *src, int64_t rows, int64_t cols, int64_t r0, int64_t r1) {
    std::vector<T> out((size_t)(r1 - r0) * cols);
    for (int64_t c = 0; c < cols; ++c) std::memcpy(out.data() + (size_t)c * (r1 - r0), src + (size_t)c * rows + r0, sizeof(T) * (size_t)(r1 - r0));
    return out;
}
void scatter_rows(const std::vector<double> &src, double *dst, int64_t rows, int64_t cols, int64_t r0, int64_t r1) {
    if (!dst) return;
    for (int64_t c = 0; c < cols; ++c) std::memcpy(dst + (size_t)c * rows + r0, src.data() + (size_t)c * (r1 - r0), sizeof(double) * (size_t)(r1 - r0));
}

void set_err(char *errbuf, size_t errlen, const std::string &m) { if (errbuf && errlen) snprintf(errbuf, errlen, "%s", m.c_str()); }

}  // namespace

extern "C" int pcmf_fit_multi(const pcmf_problem *prob, const pcmf_options *opt, const pcmf_init *init, pcmf_result *res,
                              char *errbuf, size_t errlen) {
    const int W = opt->ngpus;
    const int64_t n = prob->n_local;
    const int p = prob->p, K = opt->K;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err(errbuf, errlen, "no CUDA device available: libpcmf_b200 has no CPU fallback"); return PCMF_ECUDA; }
    if (W > ndev) { set_err(errbuf, errlen, "ngpus = " + std::to_string(W) + " but only " + std::to_string(ndev) + " CUDA device(s) are visible"); return PCMF_EINVAL; }
    if (opt->world > 1 || opt->allreduce) { set_err(errbuf, errlen, "ngpus > 1 and an external launcher (world / allreduce) exclude each other"); return PCMF_EINVAL; }
    if (opt->stream) { set_err(errbuf, errlen, "ngpus > 1: every device runs on a private stream, options.stream must be NULL"); return PCMF_EINVAL; }
    if (prob->csr_rowptr && prob->csr_on_device) { set_err(errbuf, errlen, "ngpus > 1 needs the count matrix in host memory (a device-resident CSR belongs to one device)"); return PCMF_EUNSUPPORTED; }
    if (prob->n_global > 0 && prob->n_global != n) { set_err(errbuf, errlen, "ngpus > 1 takes the whole matrix (n_global = n_local)"); return PCMF_EINVAL; }
    if (n < W) { set_err(errbuf, errlen, "fewer cells than devices"); return PCMF_EINVAL; }
    if (!nccl().load()) { set_err(errbuf, errlen, nccl().error); return PCMF_ECOMM; }

    // contiguous row ranges, balanced by non-zeros when the structure is at hand (host CSR), by rows otherwise
    std::vector<int64_t> cut(W + 1, 0);
    cut[W] = n;
    if (prob->csr_rowptr) {
        const int64_t nnz = prob->csr_rowptr[n];
        int64_t i = 0;
        for (int r = 1; r < W; ++r) {
            const int64_t target = nnz / W * r;
            while (i < n && prob->csr_rowptr[i] < target) ++i;
            cut[r] = std::min<int64_t>(std::max<int64_t>(i, cut[r - 1] + 1), n - (W - r));
        }
    } else {
        for (int r = 1; r < W; ++r) cut[r] = n / W * r + std::min<int64_t>(r, n % W);
    }

    std::vector<int> devs(W);
    for (int r = 0; r < W; ++r) devs[r] = r;
    std::vector<Nccl::comm_t> comms(W, nullptr);
    {
        const int rc = nccl().CommInitAll(comms.data(), W, devs.data());
        if (rc != 0) { set_err(errbuf, errlen, std::string("ncclCommInitAll failed: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "")); return PCMF_ECOMM; }
    }
    Shared sh;
    sh.world = W; sh.user = opt->progress; sh.user_ctx = opt->progress_ctx;
    std::vector<RankCtx> ctx(W);
    std::vector<int> rcs(W, PCMF_OK);
    std::vector<std::string> errs(W);
    std::vector<pcmf_result> rres(W);

    auto run_rank = [&](int r) {
        const int64_t r0 = cut[r], r1 = cut[r + 1], nl = r1 - r0;
        ctx[r] = RankCtx{&sh, r, comms[r]};
        // ---- this rank's rows of the input
        pcmf_problem pr = *prob;
        pr.n_local = nl; pr.n_global = n; pr.row_offset = r0;
        std::vector<double> xd; std::vector<int32_t> xi; std::vector<int64_t> rp;
        std::vector<int32_t> ccp, cri; std::vector<double> cxv;
        if (prob->dense_f64) { xd = slice_rows(prob->dense_f64, n, p, r0, r1); pr.dense_f64 = xd.data(); }
        else if (prob->dense_i32) { xi = slice_rows(prob->dense_i32, n, p, r0, r1); pr.dense_i32 = xi.data(); }
        else if (prob->csr_rowptr) {
            rp.resize(nl + 1);
            const int64_t base = prob->csr_rowptr[r0];
            for (int64_t i = 0; i <= nl; ++i) rp[i] = prob->csr_rowptr[r0 + i] - base;
            pr.csr_rowptr = rp.data(); pr.csr_colidx = prob->csr_colidx + base;
            if (prob->csr_val_f32) pr.csr_val_f32 = prob->csr_val_f32 + base;
            if (prob->csr_val_f64) pr.csr_val_f64 = prob->csr_val_f64 + base;
        } else if (prob->csc_colptr) {                       // dgCMatrix: the entries of this rank's cells, columns kept
            ccp.assign(p + 1, 0);
            for (int j = 0; j < p; ++j) {
                int32_t c = 0;
                for (int32_t e = prob->csc_colptr[j]; e < prob->csc_colptr[j + 1]; ++e) c += (prob->csc_rowidx[e] >= r0 && prob->csc_rowidx[e] < r1);
                ccp[j + 1] = ccp[j] + c;
            }
            cri.resize((size_t)std::max(ccp[p], 1)); cxv.resize((size_t)std::max(ccp[p], 1));
            for (int j = 0; j < p; ++j) {
                int32_t o = ccp[j];
                for (int32_t e = prob->csc_colptr[j]; e < prob->csc_colptr[j + 1]; ++e)
                    if (prob->csc_rowidx[e] >= r0 && prob->csc_rowidx[e] < r1) { cri[o] = (int32_t)(prob->csc_rowidx[e] - r0); cxv[o] = prob->csc_val_f64[e]; ++o; }
            }
            pr.csc_colptr = ccp.data(); pr.csc_rowidx = cri.data(); pr.csc_val_f64 = cxv.data();
        }
        // ---- options: this rank, private stream, NCCL all-reduce
        pcmf_options o = *opt;
        o.ngpus = 0; o.device = devs[r]; o.stream = nullptr; o.rank = r; o.world = W;
        o.allreduce = allreduce_nccl; o.allreduce_ctx = &ctx[r];
        o.progress = opt->progress ? progress_all : nullptr; o.progress_ctx = &ctx[r];
        // ---- warm-start values: cell-side matrices by rows, gene-side as they are
        pcmf_init in{};
        std::vector<double> sU, sa1, sa2, sal1, sal2, sPD;
        if (init) {
            in = *init;
            if (init->U) { sU = slice_rows(init->U, n, K, r0, r1); in.U = sU.data(); }
            if (init->a1) { sa1 = slice_rows(init->a1, n, K, r0, r1); in.a1 = sa1.data(); }
            if (init->a2) { sa2 = slice_rows(init->a2, n, K, r0, r1); in.a2 = sa2.data(); }
            if (init->hyper_rows && init->alpha1) { sal1 = slice_rows(init->alpha1, n, K, r0, r1); in.alpha1 = sal1.data(); }
            if (init->hyper_rows && init->alpha2) { sal2 = slice_rows(init->alpha2, n, K, r0, r1); in.alpha2 = sal2.data(); }
            if (init->prob_D) { sPD = slice_rows(init->prob_D, n, p, r0, r1); in.prob_D = sPD.data(); }
        }
        // ---- outputs: gene-side and per-iteration vectors straight into the caller's buffers on rank 0, private elsewhere
        // (every rank must record the monitors: they contain all-reduces); cell-side into row blocks scattered afterwards
        pcmf_result &rr = rres[r];
        rr = *res;
        std::vector<double> oEU, oElogU, oa1, oa2, oal1, oal2, oPD;
        std::vector<std::vector<double>> vec(6);
        auto blk = [&](double *want, std::vector<double> &buf, int64_t cols) -> double * {
            if (!want) return nullptr;
            buf.assign((size_t)nl * cols, 0.0);
            return buf.data();
        };
        rr.EU = blk(res->EU, oEU, K); rr.ElogU = blk(res->ElogU, oElogU, K); rr.a1 = blk(res->a1, oa1, K); rr.a2 = blk(res->a2, oa2, K);
        rr.alpha1 = blk(res->alpha1, oal1, K); rr.alpha2 = blk(res->alpha2, oal2, K); rr.prob_D = blk(res->prob_D, oPD, p);
        if (r != 0) {
            rr.EV = rr.ElogV = rr.b1 = rr.b2 = rr.beta1 = rr.beta2 = rr.prob_S = rr.prior_prob_S = rr.freq_D = rr.prior_prob_D = nullptr;
            rr.S = nullptr; rr.factor_order = nullptr;
            double **vp[6] = {&rr.conv_crit, &rr.norm_gap, &rr.abs_gap, &rr.optim_criterion, &rr.loglikelihood, &rr.deviance};
            for (int v = 0; v < 6; ++v)
                if (*vp[v]) { vec[v].assign((size_t)std::max(opt->iter_max, 1), 0.0); *vp[v] = vec[v].data(); }
        }
        char err[512] = {0};
        rcs[r] = pcmf_fit(&pr, &o, init ? &in : nullptr, &rr, err, sizeof err);
        errs[r] = err;
        if (rcs[r] != PCMF_OK) { sh.failed.store(1); return; }
        scatter_rows(oEU, res->EU, n, K, r0, r1); scatter_rows(oElogU, res->ElogU, n, K, r0, r1);
        scatter_rows(oa1, res->a1, n, K, r0, r1); scatter_rows(oa2, res->a2, n, K, r0, r1);
        scatter_rows(oal1, res->alpha1, n, K, r0, r1); scatter_rows(oal2, res->alpha2, n, K, r0, r1);
        scatter_rows(oPD, res->prob_D, n, p, r0, r1);
    };
    std::vector<std::thread> th;
    for (int r = 1; r < W; ++r) th.emplace_back(run_rank, r);
    run_rank(0);
    for (auto &t : th) t.join();
    for (int r = 0; r < W; ++r) { cudaSetDevice(devs[r]); nccl().CommDestroy(comms[r]); }
    cudaSetDevice(opt->device);
    for (int r = 0; r < W; ++r)
        if (rcs[r] != PCMF_OK) { set_err(errbuf, errlen, "device " + std::to_string(r) + ": " + errs[r]); return rcs[r]; }
    res->nb_iter = rres[0].nb_iter; res->converged = rres[0].converged; res->loss = rres[0].loss; res->exp_dev = rres[0].exp_dev;
    res->seconds_setup = rres[0].seconds_setup; res->seconds_iter = rres[0].seconds_iter; res->seconds_total = rres[0].seconds_total;
    return PCMF_OK;
}
