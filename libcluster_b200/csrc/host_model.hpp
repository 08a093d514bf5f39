// host_model.hpp -- fp64 host side of the VB iteration: weight and cluster
// posteriors (the K-length and K x D^3 parts of the M-step), free-energy terms,
// and the packing of posterior parameters into the device layout the CUDA
// kernels consume.  Everything O(N) lives in kernels.cu.
//
// Mirrors the operator surface of the reference's include/distributions.h
// (WeightDist :60-97, ClusterDist :200-273) but is written from the math, on
// plain std::vector storage (no Eigen, no Boost).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace lcb {

// Exceptions carrying the C-ABI status they map to (see libcluster_b200.h).
struct Error {
  int status;
  std::string what;
};
[[noreturn]] void throw_invalid(const std::string& m);
[[noreturn]] void throw_runtime(const std::string& m);
[[noreturn]] void throw_domain(const std::string& m);

// include/libcluster.h:122-127 (float literals widened exactly as C++ does)
constexpr double kConverge = (double)1e-5f;
constexpr double kFengyDel = (double)1e-5f / 10;
constexpr double kZeroCutoff = (double)0.1f;
constexpr int kSplitIter = 15;

double digamma(double x);

enum WeightKind { kDirichlet = 0, kStickBreak = 1, kGDirichlet = 2 };
enum ClusterKind { kGaussWish = 0, kNormGamma = 1 };
void model_kinds(int model, int* wkind, int* ckind);

// -------------------------------------------------------------------------
// WeightDist family.  Dirichlet: distributions.cpp:222-266; StickBreak
// :83-179; GDirichlet :186-215.
class WeightPost {
 public:
  WeightPost(int kind, double prior);  // prior <= 0: default-constructed object
  void update(const double* Nk, int K);
  double fenergy() const;
  int size() const { return (int)Nk_.size(); }
  const std::vector<double>& Elogweight() const { return Elogpi_; }
  const std::vector<double>& getNk() const { return Nk_; }
  int kind() const { return kind_; }
  double prior1() const { return a1p_; }
  double prior2() const { return a2p_; }
  double prior_fenergy() const { return Fp_; }

 private:
  int kind_;
  double a1p_, a2p_, Fp_;
  std::vector<double> Nk_, a1_, a2_, Elogv_, Elognv_, Elogpi_;
  std::vector<int> ord_;
};

// -------------------------------------------------------------------------
// ClusterDist family.  GaussWish: distributions.cpp:273-399; NormGamma
// :406-517.  Sufficient statistics are loaded (set_stats / add_stats) from the
// device reduction instead of accumulated by an O(N) addobs loop.
class ClusterPost {
 public:
  ClusterPost(int kind, double clustwidth, int D);
  void clearobs();
  // raw statistics sum q, sum q x, sum q x x^T (row-major DxD, or D for diag)
  void set_stats(double N_s, const double* x_s, const double* xx_s);
  void add_stats(double N_s, const double* x_s, const double* xx_s);
  // statistics accumulated about a centre c: n = sum q, s = sum q (x - c),
  // S = sum q (x - c)(x - c)^T (or the diagonal of it).  Adds the equivalent
  // raw statistics.
  void add_centred_stats(double n, const double* s, const double* S, const double* c);
  void update();  // throws Error (domain / invalid) like the reference
  // The two halves of update() for a GaussWish, so that ranks can share the O(D^3) half: update_params() is
  // everything up to and including iW; factor() the Cholesky of iW and its inverse (throws if iW is not positive
  // definite); export_factor / import_factor move {logdW, L^-1} (1 + D*D doubles) between ranks.
  void update_params();
  void factor();
  void export_factor(double* out) const;
  void import_factor(const double* in);
  double fenergy() const;
  double getN() const { return N_; }
  double getprior() const { return prior_; }
  int dim() const { return D_; }
  int kind() const { return kind_; }
  const std::vector<double>& mean() const { return m_; }
  std::vector<double> cov() const;  // iW / nu ; NormGamma: L * nu (sic, distributions.h:375)
  double N_s() const { return N_s_; }
  const std::vector<double>& x_s() const { return x_s_; }
  const std::vector<double>& xx_s() const { return xx_s_; }
  const std::vector<double>& iW() const { return iW_; }
  double nu() const { return nu_; }
  double beta() const { return beta_; }
  double logdW() const { return logdW_; }
  double prior_fenergy() const { return F_p_; }

  // E[log p(x)] = cconst - 0.5 * || R (x - m) ||^2   with R lower-triangular
  // (GaussWish: sqrt(nu) L^-1, iW = L L^T) or diagonal (NormGamma:
  // sqrt(nu / L_d)).  Row-major DxD (or D) doubles.
  double cconst() const;
  void whitener(std::vector<double>& R) const;
  // Direction used by splitobs: principal eigenvector of iW by the power
  // method (probutils.cpp:153-186) or the one-hot of argmax L.
  void split_direction(std::vector<double>& v) const;

 private:
  int kind_, D_;
  double prior_, N_;
  double nu_p_, beta_p_, logdW_p_, F_p_;
  std::vector<double> m_p_, iW_p_;
  double nu_, beta_, logdW_;
  std::vector<double> m_, iW_;
  double N_s_;
  std::vector<double> x_s_, xx_s_;
  std::vector<double> Linv_;  // inverse of the lower Cholesky factor of iW
};

// dense helpers (row-major)
bool cholesky_lower(std::vector<double>& A, int D);
void invert_lower(const std::vector<double>& L, int D, std::vector<double>& Linv);

// Packed statistics of one iteration: [Njk (J*K) | K * (1 + D + S)] doubles,
// S = D*D (GaussWish) or D (NormGamma).
inline int64_t stat_block(int ckind, int D) { return 1 + D + (ckind == kGaussWish ? (int64_t)D * D : D); }
inline int64_t packed_len(int ckind, int J, int K, int D) { return (int64_t)J * K + K * stat_block(ckind, D); }

}  // namespace lcb
