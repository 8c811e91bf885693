// NOT compiled in this repository (no R / Rcpp in the build image).  Replacement body of the reference's exported
// function (pkg/src/run_zi_sparse_gap_factor.cpp:292-327): same signature, so RcppExports.cpp / RcppExports.R stay as generated.
#include <Rcpp.h>
#include "pcmf_b200_shim.h"


// [[Rcpp::export]]
SEXP run_zi_sparse_gap_factor(SEXP X, int K, double sel_bound = 0.5,
                              Rcpp::Nullable<Rcpp::NumericMatrix> U = R_NilValue, Rcpp::Nullable<Rcpp::NumericMatrix> V = R_NilValue,
                              bool verbose = true, bool monitor = true, int iter_max = 1000, int iter_min = 500,
                              Rcpp::CharacterVector init_mode = "random", double epsilon = 1e-2, int additional_iter = 10,
                              int conv_mode = 1, int ninit = 1, int iter_init = 100, int ncores = 1, bool reorder_factor = true,
                              Rcpp::Nullable<uint32_t> seed = R_NilValue,
                              Rcpp::Nullable<Rcpp::NumericMatrix> a1 = R_NilValue, Rcpp::Nullable<Rcpp::NumericMatrix> a2 = R_NilValue,
                              Rcpp::Nullable<Rcpp::NumericMatrix> b1 = R_NilValue, Rcpp::Nullable<Rcpp::NumericMatrix> b2 = R_NilValue,
                              Rcpp::Nullable<Rcpp::NumericMatrix> alpha1 = R_NilValue, Rcpp::Nullable<Rcpp::NumericMatrix> alpha2 = R_NilValue,
                              Rcpp::Nullable<Rcpp::NumericMatrix> beta1 = R_NilValue, Rcpp::Nullable<Rcpp::NumericMatrix> beta2 = R_NilValue,
                              Rcpp::Nullable<Rcpp::NumericMatrix> prob_S = R_NilValue, Rcpp::Nullable<Rcpp::NumericVector> prior_S = R_NilValue,
                              Rcpp::Nullable<Rcpp::NumericMatrix> prob_D = R_NilValue, Rcpp::Nullable<Rcpp::NumericVector> prior_D = R_NilValue) {
    (void)ncores;   // omp_set_num_threads in the reference (wrapper_gap_factor.h:271-274); there are no host loops to thread
    if (Rcpp::as<std::string>(init_mode[0]) != "random") Rcpp::stop("`init_mode` input parameter should be 'random' or 'nmf'");
    std::vector<double> h1, h2, h3, h4;
    pcmf_init init{};
    init.U = pcmf_shim::opt_mat(U);   init.V = pcmf_shim::opt_mat(V);
    init.a1 = pcmf_shim::opt_mat(a1); init.a2 = pcmf_shim::opt_mat(a2);
    init.b1 = pcmf_shim::opt_mat(b1); init.b2 = pcmf_shim::opt_mat(b2);
    if (alpha1.isNotNull() && alpha2.isNotNull() && beta1.isNotNull() && beta2.isNotNull()) {
        h1 = pcmf_shim::first_row(Rcpp::NumericMatrix(alpha1)); h2 = pcmf_shim::first_row(Rcpp::NumericMatrix(alpha2));
        h3 = pcmf_shim::first_row(Rcpp::NumericMatrix(beta1));  h4 = pcmf_shim::first_row(Rcpp::NumericMatrix(beta2));
        init.alpha1 = h1.data(); init.alpha2 = h2.data(); init.beta1 = h3.data(); init.beta2 = h4.data();
    }
    init.prob_S = pcmf_shim::opt_mat(prob_S);
    init.prior_S = prior_S.isNotNull() ? REAL(Rcpp::NumericVector(prior_S)) : nullptr;
    init.prob_D = pcmf_shim::opt_mat(prob_D);
    init.prior_D = prior_D.isNotNull() ? REAL(Rcpp::NumericVector(prior_D)) : nullptr;
    Rcpp::List output = pcmf_shim::fit(X, K, /*zi=*/true, /*sparse=*/true, sel_bound, verbose, monitor, iter_max, iter_min,
                                       epsilon, additional_iter, conv_mode, ninit, iter_init, reorder_factor, seed, init);
    output.attr("class") = "spCMF";                                          // run_zi_sparse_gap_factor.cpp:325
    return output;
}
