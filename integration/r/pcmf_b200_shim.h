// NOT compiled in this repository (no R / Rcpp in the build image): the shim a pCMF maintainer adds, see INTEGRATION.md section 2.
#pragma once
#include <Rcpp.h>
#include "pcmf_b200.h"

namespace pcmf_shim {

inline const double *opt_mat(const Rcpp::Nullable<Rcpp::NumericMatrix> &m) {
    return m.isNotNull() ? REAL(Rcpp::NumericMatrix(m)) : nullptr;
}
typedef Rcpp::Nullable<Rcpp::NumericMatrix> OptMat;
typedef Rcpp::Nullable<Rcpp::NumericVector> OptVec;
inline const double *opt_vec(const OptVec &v) { return v.isNotNull() ? REAL(Rcpp::NumericVector(v)) : nullptr; }
// the warm-start arguments of run_* as the C ABI takes them.  Hyper-parameters travel as the reference's n x K / p x K
// matrices (hyper_rows = 1, gap_factor_model.cpp:126-132): the library collapses identical rows and keeps rows that differ.
inline pcmf_init make_init(const OptMat &U, const OptMat &V, const OptMat &a1, const OptMat &a2, const OptMat &b1, const OptMat &b2,
                           const OptMat &alpha1, const OptMat &alpha2, const OptMat &beta1, const OptMat &beta2,
                           const OptMat &prob_S = R_NilValue, const OptVec &prior_S = R_NilValue,
                           const OptMat &prob_D = R_NilValue, const OptVec &prior_D = R_NilValue) {
    pcmf_init init{};
    init.U = opt_mat(U);   init.V = opt_mat(V);
    init.a1 = opt_mat(a1); init.a2 = opt_mat(a2); init.b1 = opt_mat(b1); init.b2 = opt_mat(b2);
    init.alpha1 = opt_mat(alpha1); init.alpha2 = opt_mat(alpha2); init.beta1 = opt_mat(beta1); init.beta2 = opt_mat(beta2);
    init.hyper_rows = 1;
    init.prob_S = opt_mat(prob_S); init.prior_S = opt_vec(prior_S);
    init.prob_D = opt_mat(prob_D); init.prior_D = opt_vec(prior_D);
    return init;
}
inline void check_init_mode(const Rcpp::CharacterVector &init_mode) {
    const std::string m = Rcpp::as<std::string>(init_mode[0]);
    if (m == "nmf") Rcpp::stop("FixMe not implemented yet");                                      // wrapper_gap_factor.h:370
    if (m != "random") Rcpp::stop("`init_mode` input parameter should be 'random' or 'nmf'");     // wrapper_gap_factor.h:372
}
static int progress_cb(void *, int iter, double crit) {
    Rcpp::Rcout << "iter " << iter << "  criterion " << crit << std::endl;   // algorithm.h:170-173
    try { Rcpp::checkUserInterrupt(); } catch (...) { return 1; }            // algorithm.h:174
    return 0;
}

inline Rcpp::List fit(SEXP X, int K, bool zi, bool sparse, double sel_bound, bool verbose, bool monitor, int iter_max,
                      int iter_min, double epsilon, int additional_iter, int conv_mode, int ninit, int iter_init,
                      bool reorder_factor, Rcpp::Nullable<uint32_t> seed, const pcmf_init &init) {
    const bool is_dgC = Rf_isS4(X) && Rf_inherits(X, "dgCMatrix");          // extension: Matrix::dgCMatrix, no dense copy
    Rcpp::IntegerVector dim = is_dgC ? Rcpp::IntegerVector(Rcpp::S4(X).slot("Dim")) : Rcpp::IntegerVector(Rf_getAttrib(X, R_DimSymbol));
    const int64_t n = dim[0]; const int p = dim[1];
    pcmf_problem prob{};  prob.n_local = prob.n_global = n;  prob.p = p;
    if (is_dgC) {                                                             // slots @p, @i, @x as they are
        Rcpp::S4 Xs(X);
        prob.csc_colptr = INTEGER(Xs.slot("p")); prob.csc_rowidx = INTEGER(Xs.slot("i")); prob.csc_val_f64 = REAL(Xs.slot("x"));
    }
    else if (TYPEOF(X) == INTSXP) prob.dense_i32 = INTEGER(X);               // wrapper_gap_factor.h:204-207
    else if (TYPEOF(X) == REALSXP) prob.dense_f64 = REAL(X);                 // wrapper_gap_factor.h:208-209
    else Rcpp::stop("X parameter should be an integer or real valued matrix");   // wrapper_gap_factor.h:210

    pcmf_options o{};
    o.K = K; o.zero_inflation = zi; o.sparsity = sparse; o.sel_bound = sel_bound;
    o.iter_max = iter_max; o.iter_min = iter_min; o.epsilon = epsilon; o.additional_iter = additional_iter;
    o.conv_mode = conv_mode; o.monitor = monitor; o.reorder_factor = reorder_factor;
    o.ninit = ninit; o.iter_init = iter_init; o.verbose = verbose;
    o.has_seed = seed.isNotNull(); if (o.has_seed) o.seed = Rcpp::as<uint32_t>(seed);
    Rcpp::Function getOption("getOption");
    o.device = Rcpp::as<int>(getOption("pCMF.device", 0));
    // the GPU counterpart of `ncores` (wrapper_gap_factor.h:271-274): options(pCMF.ngpus = 8) shards the cells over 8 devices
    // inside the library (NCCL communicator and per-device threads are its own); options(pCMF.precision = 1) selects FP32 mode
    o.ngpus = Rcpp::as<int>(getOption("pCMF.ngpus", 1));
    o.precision = Rcpp::as<int>(getOption("pCMF.precision", 0));
    o.world = 1;
    if (verbose) { o.progress = progress_cb; o.progress_every = 10; }

    // outputs: R allocates, the library fills (column-major FP64 on both sides)
    Rcpp::NumericMatrix EU(n, K), ElogU(n, K), a1(n, K), a2(n, K), al1(n, K), al2(n, K);
    Rcpp::NumericMatrix EV(p, K), ElogV(p, K), b1(p, K), b2(p, K), be1(p, K), be2(p, K);
    Rcpp::NumericVector conv(iter_max), ngap(iter_max), agap(iter_max), elbo(iter_max), ll(iter_max), dev(iter_max);
    Rcpp::NumericMatrix probS(sparse ? p : 0, sparse ? K : 0);  Rcpp::NumericVector priorS(sparse ? p : 0);
    Rcpp::IntegerMatrix S(sparse ? p : 0, sparse ? K : 0);
    Rcpp::NumericMatrix probD(zi ? n : 0, zi ? p : 0);  Rcpp::NumericVector freqD(zi ? p : 0), priorD(zi ? p : 0);
    pcmf_result r{};
    r.EU = REAL(EU); r.ElogU = REAL(ElogU); r.a1 = REAL(a1); r.a2 = REAL(a2); r.alpha1 = REAL(al1); r.alpha2 = REAL(al2);
    r.EV = REAL(EV); r.ElogV = REAL(ElogV); r.b1 = REAL(b1); r.b2 = REAL(b2); r.beta1 = REAL(be1); r.beta2 = REAL(be2);
    r.conv_crit = REAL(conv); r.norm_gap = REAL(ngap); r.abs_gap = REAL(agap);
    r.optim_criterion = REAL(elbo); r.loglikelihood = REAL(ll); r.deviance = REAL(dev);
    if (sparse) { r.prob_S = REAL(probS); r.prior_prob_S = REAL(priorS); r.S = INTEGER(S); }
    if (zi) { r.prob_D = REAL(probD); r.freq_D = REAL(freqD); r.prior_prob_D = REAL(priorD); }

    char err[512] = {0};
    const int rc = pcmf_fit(&prob, &o, &init, &r, err, sizeof err);
    if (rc != PCMF_OK) Rcpp::stop(err);                                      // same texts as the reference's Rcpp::stop

    using Rcpp::Named; using Rcpp::List;
    const int it = r.nb_iter;
    auto head = [&](Rcpp::NumericVector v) { return Rcpp::NumericVector(v.begin(), v.begin() + it); };
    List out = List::create(                                                  // gap_factor_model.cpp:369-393
        Named("factor") = List::create(Named("U") = EU, Named("V") = EV),
        Named("variational_params") = List::create(Named("a1") = a1, Named("a2") = a2, Named("b1") = b1, Named("b2") = b2),
        Named("hyper_params") = List::create(Named("alpha1") = al1, Named("alpha2") = al2, Named("beta1") = be1, Named("beta2") = be2),
        Named("stats") = List::create(Named("EU") = EU, Named("EV") = EV, Named("ElogU") = ElogU, Named("ElogV") = ElogV));
    if (sparse) out.push_back(List::create(Named("prob_S") = probS, Named("prior_prob_S") = priorS, Named("S") = S),
                              "sparse_param");                               // sparse_gap_factor_model.cpp:278-286
    if (zi) out.push_back(List::create(Named("prob_D") = probD, Named("freq_D") = freqD, Named("prior_prob_D") = priorD),
                          "ZI_param");                                       // zi_gap_factor_model.cpp:248-255
    out.push_back(List::create(Named("converged") = (bool)r.converged, Named("nb_iter") = it,
                               Named("conv_crit") = head(conv), Named("conv_mode") = conv_mode), "convergence");
    out.push_back(r.loss, "loss");                                           // algorithm.h:365
    out.push_back(r.exp_dev, "exp_dev");                                     // algorithm.h:367
    if (monitor) out.push_back(List::create(Named("norm_gap") = head(ngap), Named("abs_gap") = head(agap),
                                            Named("loglikelihood") = head(ll), Named("deviance") = head(dev),
                                            Named("optim_criterion") = head(elbo)), "monitor");   // algorithm.h:369-376
    out.push_back(seed, "seed");                                             // wrapper_gap_factor.h:391
    return out;
}
}  // namespace pcmf_shim
