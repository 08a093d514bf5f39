// ref_driver.cpp -- C entry points around the REFERENCE's own code for the parity tests
// and the CPU baseline.  TEST INFRASTRUCTURE ONLY.
//
// This translation unit #includes /root/reference/src/cluster.cpp where it lies (nothing
// is copied into the repository) so that, besides the public learnXXX functions, the
// file-local template vbem<W,C>() (src/cluster.cpp:177-239) can be called directly.
// It is compiled together with the reference's distributions.cpp / probutils.cpp /
// comutils.cpp against the Eigen/Boost stand-ins in oracle/refshim (see the Makefile).
#include "cluster.cpp"  // found through -I/root/reference/src

#include <cstdint>
#include <cstring>
#include <string>

namespace {
struct RefResult {
  int J = 0, K = 0, D = 0, diag = 0;
  double F = 0;
  std::vector<MatrixXd> qZ;
  std::vector<std::vector<double> > elogw, nk;
  std::vector<double> wfen, cfen, cN;
  std::vector<std::vector<double> > means, covs;
};
std::string g_err;

vMatrixXd to_groups(int J, const double* Xcat, const int64_t* Nj, int D) {
  vMatrixXd X(J);
  int64_t off = 0;
  for (int j = 0; j < J; ++j) {
    X[j] = MatrixXd(Nj[j], D);
    for (int64_t n = 0; n < Nj[j]; ++n)
      for (int d = 0; d < D; ++d) X[j](n, d) = Xcat[(off + n) * D + d];
    off += Nj[j];
  }
  return X;
}

template <class W, class C>
void harvest(RefResult* r, const vMatrixXd& qZ, const std::vector<W>& weights, const std::vector<C>& clusters, int D) {
  r->J = (int)qZ.size();
  r->K = (int)clusters.size();
  r->D = D;
  r->qZ = qZ;
  for (size_t j = 0; j < weights.size(); ++j) {
    const ArrayXd e = weights[j].Elogweight(), n = weights[j].getNk();
    r->elogw.push_back(std::vector<double>(e.data(), e.data() + e.size()));
    r->nk.push_back(std::vector<double>(n.data(), n.data() + n.size()));
    r->wfen.push_back(weights[j].fenergy());
  }
  for (size_t k = 0; k < clusters.size(); ++k) {
    const RowVectorXd m = clusters[k].getmean();
    r->means.push_back(std::vector<double>(m.data(), m.data() + m.size()));
    const MatrixXd c = MatrixXd(clusters[k].getcov());
    std::vector<double> cv((size_t)c.size());
    if (c.rows() > 1 && c.cols() > 1) {
      for (Index i = 0; i < c.rows(); ++i) for (Index j2 = 0; j2 < c.cols(); ++j2) cv[(size_t)(i * c.cols() + j2)] = c(i, j2);
    } else {
      for (Index i = 0; i < c.size(); ++i) cv[(size_t)i] = c(i);
    }
    r->covs.push_back(cv);
    r->cfen.push_back(clusters[k].fenergy());
    r->cN.push_back(clusters[k].getN());
  }
}

template <class W, class C>
double run_vbem(RefResult* r, const vMatrixXd& X, const double* q0, int K, double prior, int maxit, bool sparse) {
  const int J = (int)X.size();
  vMatrixXd qZ(J);
  int64_t off = 0;
  for (int j = 0; j < J; ++j) {
    qZ[j] = MatrixXd(X[j].rows(), K);
    for (Index n = 0; n < X[j].rows(); ++n)
      for (int k = 0; k < K; ++k) qZ[j](n, k) = q0[(off + n) * K + k];
    off += X[j].rows();
  }
  std::vector<W> weights;
  std::vector<C> clusters;
  const double F = vbem<W, C>(X, qZ, weights, clusters, prior, maxit, sparse, false);
  harvest(r, qZ, weights, clusters, (int)X[0].cols());
  return F;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }
void ref_free(void* h) { delete (RefResult*)h; }

// learnXXX of the reference (include/libcluster.h); model ids as in libcluster_b200.h
int ref_learn(int model, int J, const double* Xcat, const int64_t* Nj, int D, double prior, int maxclusters, int sparse,
              unsigned nthreads, void** out) {
  RefResult* r = new RefResult();
  *out = r;
  try {
    vMatrixXd X = to_groups(J, Xcat, Nj, D);
    vMatrixXd qZ;
    if (model == 0) {
      StickBreak w; std::vector<GaussWish> c; MatrixXd q;
      r->F = learnVDP(X[0], q, w, c, prior, maxclusters, false, nthreads);
      harvest(r, vMatrixXd(1, q), std::vector<StickBreak>(1, w), c, D);
    } else if (model == 1) {
      Dirichlet w; std::vector<GaussWish> c; MatrixXd q;
      r->F = learnBGMM(X[0], q, w, c, prior, maxclusters, false, nthreads);
      harvest(r, vMatrixXd(1, q), std::vector<Dirichlet>(1, w), c, D);
    } else if (model == 2) {
      Dirichlet w; std::vector<NormGamma> c; MatrixXd q;
      r->F = learnDGMM(X[0], q, w, c, prior, maxclusters, false, nthreads);
      harvest(r, vMatrixXd(1, q), std::vector<Dirichlet>(1, w), c, D);
    } else if (model == 3) {
      std::vector<GDirichlet> w; std::vector<GaussWish> c;
      r->F = learnGMC(X, qZ, w, c, prior, maxclusters, sparse != 0, false, nthreads);
      harvest(r, qZ, w, c, D);
    } else if (model == 4) {
      std::vector<Dirichlet> w; std::vector<GaussWish> c;
      r->F = learnSGMC(X, qZ, w, c, prior, maxclusters, sparse != 0, false, nthreads);
      harvest(r, qZ, w, c, D);
    } else {
      std::vector<GDirichlet> w; std::vector<NormGamma> c;
      r->F = learnDGMC(X, qZ, w, c, prior, maxclusters, sparse != 0, false, nthreads);
      harvest(r, qZ, w, c, D);
    }
  } catch (const std::invalid_argument& e) { g_err = e.what(); return 1;
  } catch (const std::runtime_error& e) { g_err = e.what(); return 2;
  } catch (const std::exception& e) { g_err = e.what(); return 3;
  } catch (...) { g_err = "unknown"; return 3; }
  return 0;
}

// The single-group entry points with a caller-constructed weight object (Dirichlet(alpha) / StickBreak(concentration)):
// the fit keeps that prior (src/cluster.cpp:653,684) while split refinements use default-constructed weights (:460-461)
int ref_learn_wprior(int model, const double* Xcat, int64_t N, int D, double prior, double wprior, int maxclusters,
                     unsigned nthreads, void** out) {
  RefResult* r = new RefResult();
  *out = r;
  try {
    const int64_t Nj[1] = {N};
    vMatrixXd X = to_groups(1, Xcat, Nj, D);
    MatrixXd q;
    if (model == 0) {
      StickBreak w(wprior); std::vector<GaussWish> c;
      r->F = learnVDP(X[0], q, w, c, prior, maxclusters, false, nthreads);
      harvest(r, vMatrixXd(1, q), std::vector<StickBreak>(1, w), c, D);
    } else if (model == 1) {
      Dirichlet w(wprior); std::vector<GaussWish> c;
      r->F = learnBGMM(X[0], q, w, c, prior, maxclusters, false, nthreads);
      harvest(r, vMatrixXd(1, q), std::vector<Dirichlet>(1, w), c, D);
    } else if (model == 2) {
      Dirichlet w(wprior); std::vector<NormGamma> c;
      r->F = learnDGMM(X[0], q, w, c, prior, maxclusters, false, nthreads);
      harvest(r, vMatrixXd(1, q), std::vector<Dirichlet>(1, w), c, D);
    } else {
      g_err = "ref_learn_wprior: single-group models only";
      return 1;
    }
  } catch (const std::invalid_argument& e) { g_err = e.what(); return 1;
  } catch (const std::runtime_error& e) { g_err = e.what(); return 2;
  } catch (const std::exception& e) { g_err = e.what(); return 3;
  } catch (...) { g_err = "unknown"; return 3; }
  return 0;
}

// vbem<W,C>() (src/cluster.cpp:177) from caller-supplied responsibilities q0 [N x K] row-major
int ref_vbem(int model, int J, const double* Xcat, const int64_t* Nj, int D, const double* q0, int K, double prior,
             int maxit, int sparse, int nthreads, void** out) {
  RefResult* r = new RefResult();
  *out = r;
  try {
    omp_set_num_threads(nthreads > 0 ? nthreads : 1);
    vMatrixXd X = to_groups(J, Xcat, Nj, D);
    switch (model) {
      case 0: r->F = run_vbem<StickBreak, GaussWish>(r, X, q0, K, prior, maxit, sparse != 0); break;
      case 1: r->F = run_vbem<Dirichlet, GaussWish>(r, X, q0, K, prior, maxit, sparse != 0); break;
      case 2: r->F = run_vbem<Dirichlet, NormGamma>(r, X, q0, K, prior, maxit, sparse != 0); break;
      case 3: r->F = run_vbem<GDirichlet, GaussWish>(r, X, q0, K, prior, maxit, sparse != 0); break;
      case 4: r->F = run_vbem<Dirichlet, GaussWish>(r, X, q0, K, prior, maxit, sparse != 0); break;
      default: r->F = run_vbem<GDirichlet, NormGamma>(r, X, q0, K, prior, maxit, sparse != 0); break;
    }
  } catch (const std::invalid_argument& e) { g_err = e.what(); return 1;
  } catch (const std::runtime_error& e) { g_err = e.what(); return 2;
  } catch (const std::exception& e) { g_err = e.what(); return 3;
  } catch (...) { g_err = "unknown"; return 3; }
  return 0;
}

double ref_F(void* h) { return ((RefResult*)h)->F; }
int ref_K(void* h) { return ((RefResult*)h)->K; }
int ref_J(void* h) { return ((RefResult*)h)->J; }
// qZ of group j as row-major [Nj x K]
void ref_get_qZ(void* h, int j, double* out) {
  const MatrixXd& q = ((RefResult*)h)->qZ[(size_t)j];
  for (Index n = 0; n < q.rows(); ++n) for (Index k = 0; k < q.cols(); ++k) out[n * q.cols() + k] = q(n, k);
}
int64_t ref_qrows(void* h, int j) { return ((RefResult*)h)->qZ[(size_t)j].rows(); }
int ref_qcols(void* h, int j) { return (int)((RefResult*)h)->qZ[(size_t)j].cols(); }
void ref_get_weights(void* h, int j, double* elogw, double* nk, double* fen) {
  RefResult* r = (RefResult*)h;
  std::memcpy(elogw, r->elogw[(size_t)j].data(), sizeof(double) * r->elogw[(size_t)j].size());
  std::memcpy(nk, r->nk[(size_t)j].data(), sizeof(double) * r->nk[(size_t)j].size());
  *fen = r->wfen[(size_t)j];
}
int ref_weights_K(void* h, int j) { return (int)((RefResult*)h)->elogw[(size_t)j].size(); }
void ref_get_cluster(void* h, int k, double* mean, double* cov, double* N, double* fen) {
  RefResult* r = (RefResult*)h;
  std::memcpy(mean, r->means[(size_t)k].data(), sizeof(double) * r->means[(size_t)k].size());
  std::memcpy(cov, r->covs[(size_t)k].data(), sizeof(double) * r->covs[(size_t)k].size());
  *N = r->cN[(size_t)k];
  *fen = r->cfen[(size_t)k];
}
int ref_cov_len(void* h, int k) { return (int)((RefResult*)h)->covs[(size_t)k].size(); }

}  // extern "C"
