// libcluster.h -- drop-in for libcluster's include/libcluster.h: the learnXXX
// entry points of the flat / grouped mixture family with the reference's
// signatures (include/libcluster.h:177-186, :218-227, :262-271, :356-366,
// :409-419, :462-472), implemented over the B200 engine's C ABI.
// Header-only; needs Eigen 3 and liblcb200.so.  learnBEMM/EGMC/SCM/MCM are out
// of scope of this engine.
#ifndef LCB200_LIBCLUSTER_H
#define LCB200_LIBCLUSTER_H

#include <Eigen/Dense>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "distributions.h"
#include "libcluster_b200.h"

namespace libcluster {

const double PRIORVAL = 1.0;
const unsigned int TRUNC = 100;
const unsigned int SPLITITER = 15;
const double CONVERGE = 1e-5f;
const double FENGYDEL = CONVERGE / 10;
const double ZEROCUTOFF = 0.1f;

typedef std::vector<Eigen::MatrixXd> vMatrixXd;
typedef std::vector<std::vector<Eigen::MatrixXd> > vvMatrixXd;

// Device arithmetic and device ordinal of the fits started through this header.  The reference API is fp64
// throughout; the engine's measured path is LCB_F32 (1e-5 on qZ and F).  Compile with
// -DLCB200_SHIM_PRECISION=LCB_F64 (or run with LCB_SHIM_PRECISION=f64 / LCB_SHIM_DEVICE=<n> in the environment)
// for the fp64 kernels or another GPU.
#ifndef LCB200_SHIM_PRECISION
#define LCB200_SHIM_PRECISION LCB_F32
#endif
#ifndef LCB200_SHIM_DEVICE
#define LCB200_SHIM_DEVICE 0
#endif

namespace detail {
struct EngineGuard {
  lcb_engine* e;
  EngineGuard() : e(nullptr) {
    int prec = LCB200_SHIM_PRECISION, dev = LCB200_SHIM_DEVICE;
    if (const char* p = std::getenv("LCB_SHIM_PRECISION")) prec = (!std::strcmp(p, "f64") || !std::strcmp(p, "F64")) ? LCB_F64 : LCB_F32;
    if (const char* d = std::getenv("LCB_SHIM_DEVICE")) dev = std::atoi(d);
    distributions::detail::raise(lcb_create(&e, dev, prec));
  }
  ~EngineGuard() { lcb_destroy(e); }
};

// The weight prior the fit runs with.  cluster<W,C>() keeps the prior of every weight object the caller passes and
// appends default-constructed ones up to J (weights.resize(J, W()), src/cluster.cpp:192; the single-group wrappers
// pass their argument, :653,684,715).  The engine takes one prior for all groups: the callers' objects must agree
// (an empty vector means the default prior).
template <class W>
double common_weight_prior(const std::vector<W>& weights, int J) {
  if (weights.empty()) return -1.0;
  const double p0 = weights[0].prior();
  for (size_t j = 1; j < weights.size() && (int)j < J; ++j)
    if (weights[j].prior() != p0)
      throw std::invalid_argument("the device engine takes one weight prior for all groups");
  if ((int)weights.size() < J && p0 != W().prior())
    throw std::invalid_argument("the device engine takes one weight prior for all groups");
  return p0;
}

// Runs cluster<W,C>() on the device and rebuilds the caller's objects from the
// engine's statistics: qZ, weights (one per group) and clusters.
template <class W, class C>
double fit(int model, const vMatrixXd& X, vMatrixXd& qZ, std::vector<W>& weights, std::vector<C>& clusters,
           const double clusterprior, const int maxclusters, const bool sparse, const bool verbose,
           const unsigned int nthreads) {
  using distributions::detail::raise;
  if (nthreads < 1) throw std::invalid_argument("Must specify at least one thread for execution!");
  const int J = (int)X.size();
  if (J < 1) throw std::invalid_argument("no observations");
  const int D = (int)X[0].cols();
  std::vector<const double*> ptr(J);
  std::vector<int64_t> Nj(J), ld(J);
  for (int j = 0; j < J; ++j) {
    if (X[j].cols() != D) throw std::invalid_argument("X dimensions are inconsistent between groups!");
    ptr[j] = X[j].data();
    Nj[j] = X[j].rows();
    ld[j] = distributions::detail::ld_of(X[j]);
  }
  const double wprior = common_weight_prior(weights, J);
  EngineGuard g;
  const int layout = Eigen::MatrixXd::IsRowMajor ? LCB_ROW_MAJOR : LCB_COL_MAJOR;
  raise(lcb_set_data(g.e, J, ptr.data(), Nj.data(), D, ld.data(), layout));
  double F = 0;
  int K = 0;
  raise(lcb_learn(g.e, model, clusterprior, wprior, maxclusters, sparse, verbose, nthreads, &F, &K));
  qZ.resize(J);
  weights.resize(J, W());
  for (int j = 0; j < J; ++j) {
    qZ[j].resize(Nj[j], K);
    raise(lcb_get_qz(g.e, j, qZ[j].data(), Eigen::MatrixXd::IsRowMajor ? K : (Nj[j] > 0 ? Nj[j] : 1), layout));
    Eigen::ArrayXd Nk(K);
    raise(lcb_get_group_weights(g.e, j, Nk.data(), nullptr, nullptr));
    weights[j].update(Nk);
  }
  clusters.clear();
  const size_t S = (model == LCB_DGMM || model == LCB_DGMC) ? (size_t)D : (size_t)D * D;
  std::vector<double> xs(D), xxs(S);
  for (int k = 0; k < K; ++k) {
    double Ns = 0;
    raise(lcb_get_cluster(g.e, k, &Ns, xs.data(), xxs.data(), nullptr, nullptr, nullptr, nullptr));
    C c(clusterprior, (unsigned)D);
    c.load_stats(Ns, xs.data(), xxs.data());
    c.update();
    clusters.push_back(c);
  }
  return F;
}

template <class W, class C>
double fit1(int model, const Eigen::MatrixXd& X, Eigen::MatrixXd& qZ, W& weights, std::vector<C>& clusters,
            const double clusterprior, const int maxclusters, const bool verbose, const unsigned int nthreads) {
  vMatrixXd vX(1, X), vqZ;
  std::vector<W> vw(1, weights);
  const double F = fit<W, C>(model, vX, vqZ, vw, clusters, clusterprior, maxclusters, false, verbose, nthreads);
  qZ = vqZ[0];
  weights = vw[0];
  return F;
}
}  // namespace detail

inline double learnVDP(const Eigen::MatrixXd& X, Eigen::MatrixXd& qZ, distributions::StickBreak& weights,
                       std::vector<distributions::GaussWish>& clusters, const double clusterprior = PRIORVAL,
                       const int maxclusters = -1, const bool verbose = false, const unsigned int nthreads = 1) {
  return detail::fit1(LCB_VDP, X, qZ, weights, clusters, clusterprior, maxclusters, verbose, nthreads);
}
inline double learnBGMM(const Eigen::MatrixXd& X, Eigen::MatrixXd& qZ, distributions::Dirichlet& weights,
                        std::vector<distributions::GaussWish>& clusters, const double clusterprior = PRIORVAL,
                        const int maxclusters = -1, const bool verbose = false, const unsigned int nthreads = 1) {
  return detail::fit1(LCB_BGMM, X, qZ, weights, clusters, clusterprior, maxclusters, verbose, nthreads);
}
inline double learnDGMM(const Eigen::MatrixXd& X, Eigen::MatrixXd& qZ, distributions::Dirichlet& weights,
                        std::vector<distributions::NormGamma>& clusters, const double clusterprior = PRIORVAL,
                        const int maxclusters = -1, const bool verbose = false, const unsigned int nthreads = 1) {
  return detail::fit1(LCB_DGMM, X, qZ, weights, clusters, clusterprior, maxclusters, verbose, nthreads);
}
inline double learnGMC(const vMatrixXd& X, vMatrixXd& qZ, std::vector<distributions::GDirichlet>& weights,
                       std::vector<distributions::GaussWish>& clusters, const double clusterprior = PRIORVAL,
                       const int maxclusters = -1, const bool sparse = false, const bool verbose = false,
                       const unsigned int nthreads = 1) {
  return detail::fit(LCB_GMC, X, qZ, weights, clusters, clusterprior, maxclusters, sparse, verbose, nthreads);
}
inline double learnSGMC(const vMatrixXd& X, vMatrixXd& qZ, std::vector<distributions::Dirichlet>& weights,
                        std::vector<distributions::GaussWish>& clusters, const double clusterprior = PRIORVAL,
                        const int maxclusters = -1, const bool sparse = false, const bool verbose = false,
                        const unsigned int nthreads = 1) {
  return detail::fit(LCB_SGMC, X, qZ, weights, clusters, clusterprior, maxclusters, sparse, verbose, nthreads);
}
inline double learnDGMC(const vMatrixXd& X, vMatrixXd& qZ, std::vector<distributions::GDirichlet>& weights,
                        std::vector<distributions::NormGamma>& clusters, const double clusterprior = PRIORVAL,
                        const int maxclusters = -1, const bool sparse = false, const bool verbose = false,
                        const unsigned int nthreads = 1) {
  return detail::fit(LCB_DGMC, X, qZ, weights, clusters, clusterprior, maxclusters, sparse, verbose, nthreads);
}

}  // namespace libcluster
#endif
