// NOT compiled in this repository (no R / Rcpp in the build image).  Replacement body of the reference's exported
// function (pkg/src/run_sparse_gap_factor.cpp:279-312): same signature, so RcppExports.cpp / RcppExports.R stay as generated.
#include <Rcpp.h>
#include "pcmf_b200_shim.h"

using pcmf_shim::OptMat;
using pcmf_shim::OptVec;

// [[Rcpp::export]]
SEXP run_sparse_gap_factor(SEXP X, int K, double sel_bound = 0.5,
                           OptMat U = R_NilValue, OptMat V = R_NilValue, bool verbose = true, bool monitor = true,
                           int iter_max = 1000, int iter_min = 500, Rcpp::CharacterVector init_mode = "random", double epsilon = 1e-2,
                           int additional_iter = 10, int conv_mode = 1, int ninit = 1, int iter_init = 100, int ncores = 1,
                           bool reorder_factor = true, Rcpp::Nullable<uint32_t> seed = R_NilValue,
                           OptMat a1 = R_NilValue, OptMat a2 = R_NilValue, OptMat b1 = R_NilValue, OptMat b2 = R_NilValue,
                           OptMat alpha1 = R_NilValue, OptMat alpha2 = R_NilValue, OptMat beta1 = R_NilValue, OptMat beta2 = R_NilValue,
                           OptMat prob_S = R_NilValue, OptVec prior_S = R_NilValue) {
    (void)ncores;   // omp_set_num_threads in the reference (wrapper_gap_factor.h:271-274); see options(pCMF.ngpus) in the shim
    pcmf_shim::check_init_mode(init_mode);
    const pcmf_init init = pcmf_shim::make_init(U, V, a1, a2, b1, b2, alpha1, alpha2, beta1, beta2, prob_S, prior_S);
    Rcpp::List output = pcmf_shim::fit(X, K, /*zi=*/false, /*sparse=*/true, sel_bound, verbose, monitor, iter_max, iter_min,
                                       epsilon, additional_iter, conv_mode, ninit, iter_init, reorder_factor, seed, init);
    output.attr("class") = "spCMF";
    return output;
}
