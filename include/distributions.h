// distributions.h -- drop-in for libcluster's include/distributions.h on top of
// the B200 engine's C ABI (libcluster_b200.h).  Same namespace, class names,
// constructors and member functions as the reference (include/distributions.h:
// WeightDist :60-97, StickBreak :103, GDirichlet :147, Dirichlet :163,
// ClusterDist :200-273, GaussWish :279-337, NormGamma :343-400); the state lives
// in the engine's host-side posterior objects, and the O(N) members (addobs,
// Eloglike, splitobs) stream the caller's Eigen matrices through the GPU.
// Needs Eigen 3 (Dense) and linking against liblcb200.so.
#ifndef LCB200_DISTRIBUTIONS_H
#define LCB200_DISTRIBUTIONS_H

#include <Eigen/Dense>
#include <stdexcept>
#include <string>
#include <vector>

#include "libcluster_b200.h"

namespace distributions {

const double BETAPRIOR = 1.0, NUPRIOR = 1.0, ALPHA1PRIOR = 1.0, ALPHA2PRIOR = 1.0, APRIOR = 1.0;
typedef Eigen::Array<bool, Eigen::Dynamic, 1> ArrayXb;

namespace detail {
// status code of the C ABI -> the exception the reference throws (libcluster.h:171-175)
inline void raise(int rc) {
  if (rc == LCB_OK) return;
  const std::string msg = lcb_last_error();
  if (rc == LCB_EINVAL) throw std::invalid_argument(msg);
  if (rc == LCB_EDOMAIN) throw std::domain_error(msg);
  throw std::runtime_error(msg);
}
// one process-wide engine for the operator-level calls
inline lcb_engine* engine() {
  static lcb_engine* e = nullptr;
  if (!e) raise(lcb_create(&e, 0, LCB_F32));
  return e;
}
inline int layout_of(const Eigen::MatrixXd&) { return Eigen::MatrixXd::IsRowMajor ? LCB_ROW_MAJOR : LCB_COL_MAJOR; }
inline int64_t ld_of(const Eigen::MatrixXd& X) { return Eigen::MatrixXd::IsRowMajor ? X.cols() : X.rows(); }
}  // namespace detail

class WeightDist {
 public:
  void update(const Eigen::ArrayXd& Nk) {
    detail::raise(lcb_weights_update(h_, Nk.data(), (int)Nk.size()));
    sync();
  }
  const Eigen::ArrayXd& Elogweight() const { return Elogpi_; }
  const Eigen::ArrayXd& getNk() const { return Nk_; }
  double fenergy() const { return lcb_weights_fenergy(h_); }
  // The prior this object was constructed with (Dirichlet's alpha / StickBreak's concentration; the default
  // constructors use APRIOR / ALPHA2PRIOR = 1).  Not in the reference's interface: its fits read the private prior
  // fields of the objects they are handed (src/cluster.cpp:653,684), this header hands the value to the engine.
  double prior() const { return prior_ > 0 ? prior_ : 1.0; }
  virtual ~WeightDist() { lcb_weights_destroy(h_); }
  WeightDist(const WeightDist& o) : kind_(o.kind_), prior_(o.prior_) { create(); if (o.Nk_.size() > 0 && o.updated_) update(o.Nk_); }
  WeightDist& operator=(const WeightDist& o) {
    if (this != &o) { lcb_weights_destroy(h_); kind_ = o.kind_; prior_ = o.prior_; create(); if (o.updated_) update(o.Nk_); }
    return *this;
  }

 protected:
  WeightDist(int kind, double prior) : kind_(kind), prior_(prior) { create(); }
  void create() {
    h_ = nullptr;
    updated_ = false;
    detail::raise(lcb_weights_create(&h_, kind_, prior_));
    sync();
    updated_ = false;
  }
  void sync() {
    const int K = lcb_weights_size(h_);
    Nk_.resize(K);
    Elogpi_.resize(K);
    lcb_weights_getnk(h_, Nk_.data());
    lcb_weights_elogweight(h_, Elogpi_.data());
    updated_ = true;
  }
  lcb_weights* h_;
  int kind_;
  double prior_;
  bool updated_;
  Eigen::ArrayXd Nk_, Elogpi_;
};

class StickBreak : public WeightDist {
 public:
  StickBreak() : WeightDist(LCB_W_STICKBREAK, -1.0) {}
  explicit StickBreak(const double concentration) : WeightDist(LCB_W_STICKBREAK, check(concentration)) {}

 protected:
  StickBreak(int kind) : WeightDist(kind, -1.0) {}
  static double check(double c) {
    if (c <= 0) throw std::invalid_argument("Concentration parameter has to be > 0!");
    return c;
  }
};

class GDirichlet : public StickBreak {
 public:
  GDirichlet() : StickBreak(LCB_W_GDIRICHLET) {}
};

class Dirichlet : public WeightDist {
 public:
  Dirichlet() : WeightDist(LCB_W_DIRICHLET, -1.0) {}
  explicit Dirichlet(const double alpha) : WeightDist(LCB_W_DIRICHLET, check(alpha)) {}

 private:
  static double check(double a) {
    if (a <= 0) throw std::invalid_argument("Alpha prior must be > 0!");
    return a;
  }
};

class ClusterDist {
 public:
  void addobs(const Eigen::VectorXd& qZk, const Eigen::MatrixXd& X) {
    if (X.cols() != (Eigen::Index)D) throw std::invalid_argument("Mismatched dims. of cluster params and obs.!");
    if (qZk.rows() != X.rows()) throw std::invalid_argument("qZk and X ar not the same length!");
    detail::raise(lcb_cluster_addobs(detail::engine(), h_, qZk.data(), X.data(), X.rows(), detail::ld_of(X),
                                     detail::layout_of(X)));
  }
  void update() { detail::raise(lcb_cluster_update(h_)); N = lcb_cluster_getn(h_); }
  void clearobs() { lcb_cluster_clearobs(h_); }
  Eigen::VectorXd Eloglike(const Eigen::MatrixXd& X) const {
    Eigen::VectorXd out(X.rows());
    detail::raise(lcb_cluster_eloglike(detail::engine(), h_, X.data(), X.rows(), detail::ld_of(X), detail::layout_of(X),
                                       out.data()));
    return out;
  }
  double fenergy() const { return lcb_cluster_fenergy(h_); }
  ArrayXb splitobs(const Eigen::MatrixXd& X) const {
    std::vector<uint8_t> f((size_t)X.rows());
    detail::raise(lcb_cluster_splitobs(detail::engine(), h_, X.data(), X.rows(), detail::ld_of(X), detail::layout_of(X),
                                       f.data()));
    ArrayXb out(X.rows());
    for (Eigen::Index n = 0; n < X.rows(); ++n) out(n) = f[(size_t)n] != 0;
    return out;
  }
  double getN() const { return N; }
  double getprior() const { return prior; }
  virtual ~ClusterDist() { lcb_cluster_destroy(h_); }
  ClusterDist(const ClusterDist& o) : D(o.D), prior(o.prior), N(o.N), kind_(o.kind_) { clone(o); }
  ClusterDist& operator=(const ClusterDist& o) {
    if (this != &o) { lcb_cluster_destroy(h_); D = o.D; prior = o.prior; N = o.N; kind_ = o.kind_; clone(o); }
    return *this;
  }
  // Not in the reference: load sufficient statistics computed by the engine
  // (GaussWish/NormGamma keep theirs private, distributions.h:315-336).
  void load_stats(double N_s, const double* x_s, const double* xx_s) {
    detail::raise(lcb_cluster_set_stats(h_, N_s, x_s, xx_s));
  }

 protected:
  ClusterDist(int kind, const double prior_, const unsigned int D_) : D(D_), prior(prior_), N(0), kind_(kind) {
    h_ = nullptr;
    detail::raise(lcb_cluster_create(&h_, kind_, prior, (int)D));
  }
  void clone(const ClusterDist& o) {
    h_ = nullptr;
    detail::raise(lcb_cluster_create(&h_, kind_, prior, (int)D));
    const size_t S = kind_ == LCB_C_GAUSSWISH ? (size_t)D * D : D;
    std::vector<double> xs(D), xxs(S);
    double Ns = 0;
    lcb_cluster_get_stats(o.h_, &Ns, xs.data(), xxs.data());
    lcb_cluster_set_stats(h_, Ns, xs.data(), xxs.data());
    if (o.N > 0) lcb_cluster_update(h_);
  }
  lcb_cluster* h_;
  unsigned int D;
  double prior;
  double N;
  int kind_;
};

class GaussWish : public ClusterDist {
 public:
  GaussWish(const double clustwidth, const unsigned int D_) : ClusterDist(LCB_C_GAUSSWISH, clustwidth, D_) {}
  Eigen::RowVectorXd getmean() const {
    Eigen::RowVectorXd m(D);
    lcb_cluster_getmean(h_, m.data());
    return m;
  }
  Eigen::MatrixXd getcov() const {
    std::vector<double> c((size_t)D * D);  // row-major from the C ABI
    lcb_cluster_getcov(h_, c.data());
    Eigen::MatrixXd out(D, D);
    for (unsigned i = 0; i < D; ++i)
      for (unsigned j = 0; j < D; ++j) out(i, j) = c[(size_t)i * D + j];
    return out;
  }
};

class NormGamma : public ClusterDist {
 public:
  NormGamma(const double clustwidth, const unsigned int D_) : ClusterDist(LCB_C_NORMGAMMA, clustwidth, D_) {}
  Eigen::RowVectorXd getmean() const {
    Eigen::RowVectorXd m(D);
    lcb_cluster_getmean(h_, m.data());
    return m;
  }
  Eigen::RowVectorXd getcov() const {
    Eigen::RowVectorXd c(D);
    lcb_cluster_getcov(h_, c.data());
    return c;
  }
};

}  // namespace distributions
#endif
