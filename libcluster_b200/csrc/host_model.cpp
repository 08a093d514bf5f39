// host_model.cpp -- see host_model.hpp.
#include "host_model.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>

namespace lcb {

void throw_invalid(const std::string& m) { throw Error{1, m}; }
void throw_runtime(const std::string& m) { throw Error{2, m}; }
void throw_domain(const std::string& m) { throw Error{3, m}; }

static const double kPi = 3.141592653589793238462643383279502884;
// include/distributions.h:39-43
static const double kBetaPrior = 1.0, kNuPrior = 1.0, kAlpha1Prior = 1.0, kAlpha2Prior = 1.0;
// src/probutils.cpp:39-40
static const double kEigConThresh = (double)1.0e-8f;
static const int kEigMaxIter = 100;

// psi(x), x > 0: upward recurrence to x >= 8 then the Stirling series
// (stands in for boost::math::digamma at distributions.cpp:160-162,255,360).
double digamma(double x) {
  double acc = 0.0;
  for (; x < 8.0; x += 1.0) acc += 1.0 / x;
  const double i2 = 1.0 / (x * x);
  const double series =
      i2 * (1.0 / 12 - i2 * (1.0 / 120 - i2 * (1.0 / 252 - i2 * (1.0 / 240 - i2 * (1.0 / 132 -
      i2 * (691.0 / 32760 - i2 * (1.0 / 12 - i2 * (3617.0 / 8160))))))));
  return std::log(x) - 0.5 / x - series - acc;
}

void model_kinds(int model, int* wkind, int* ckind) {
  switch (model) {
    case 0: *wkind = kStickBreak; *ckind = kGaussWish; break;   // learnVDP  cluster.cpp:636
    case 1: *wkind = kDirichlet; *ckind = kGaussWish; break;    // learnBGMM :667
    case 2: *wkind = kDirichlet; *ckind = kNormGamma; break;    // learnDGMM :698
    case 3: *wkind = kGDirichlet; *ckind = kGaussWish; break;   // learnGMC  :763
    case 4: *wkind = kDirichlet; *ckind = kGaussWish; break;    // learnSGMC :787
    case 5: *wkind = kGDirichlet; *ckind = kNormGamma; break;   // learnDGMC :810
    default: throw_invalid("unknown model id");
  }
}

// ---------------------------------------------------------------- weights --
WeightPost::WeightPost(int kind, double prior) : kind_(kind) {
  if (kind < 0 || kind > 2) throw_invalid("unknown weight distribution kind");
  // prior < 0 (or NaN) is the sentinel of the default constructors; an explicit 0 is the reference's error
  // (distributions.cpp:107-108, :234-235)
  if (prior == 0) throw_invalid(kind == kDirichlet ? "Alpha prior must be > 0!" : "Concentration parameter has to be > 0!");
  a1p_ = prior > 0 ? prior : kAlpha1Prior;
  a2p_ = kAlpha2Prior;
  Fp_ = std::lgamma(a1p_) + std::lgamma(a2p_) - std::lgamma(a1p_ + a2p_);
  Nk_.assign(1, 0.0);
  a1_.assign(1, a1p_);
  a2_.assign(1, a2p_);
  Elogv_.assign(1, 0.0);
  Elognv_.assign(1, 0.0);
  Elogpi_.assign(1, 0.0);
  ord_.assign(1, 0);
}

void WeightPost::update(const double* Nk, int K) {
  Nk_.assign(Nk, Nk + K);
  a1_.resize(K);
  a2_.resize(K);
  Elogv_.resize(K);
  Elognv_.resize(K);
  Elogpi_.resize(K);
  ord_.resize(K);
  double total = 0;
  for (int k = 0; k < K; ++k) {
    a1_[k] = a1p_ + Nk[k];
    total += Nk[k];
  }
  if (kind_ == kDirichlet) {
    double asum = 0;
    for (int k = 0; k < K; ++k) asum += a1_[k];
    const double psum = digamma(asum);
    for (int k = 0; k < K; ++k) Elogpi_[k] = digamma(a1_[k]) - psum;
    return;
  }
  // size-ordered stick breaking; std::sort on (id, count) pairs with the same
  // "greater count first" predicate as the reference so that ties fall the
  // same way under the same standard library.
  std::vector<std::pair<int, double>> ov(K);
  for (int k = 0; k < K; ++k) ov[k] = std::make_pair(k, Nk[k]);
  std::sort(ov.begin(), ov.end(),
            [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.second > b.second; });
  double seen = 0, left = 0;
  for (int r = 0; r < K; ++r) {
    const int k = ov[r].first;
    ord_[r] = k;
    seen += Nk[k];
    a2_[k] = a2p_ + (total - seen);
    const double ps = digamma(a1_[k] + a2_[k]);
    Elogv_[k] = digamma(a1_[k]) - ps;
    Elognv_[k] = digamma(a2_[k]) - ps;
    Elogpi_[k] = Elogv_[k] + left;
    left += Elognv_[k];
  }
  if (kind_ == kGDirichlet) {  // last (smallest) stick takes what is left
    const int s = ord_[K - 1];
    Elogpi_[s] -= Elogv_[s];
    Elogv_[s] = 0;
    Elognv_[s] = 0;
  }
}

double WeightPost::fenergy() const {
  const int K = (int)a1_.size();
  if (kind_ == kDirichlet) {
    double asum = 0, esum = 0, t = 0;
    for (int k = 0; k < K; ++k) {
      asum += a1_[k];
      esum += Elogpi_[k];
      t += (a1_[k] - 1) * Elogpi_[k] - std::lgamma(a1_[k]);
    }
    return std::lgamma(asum) - (a1p_ - 1) * esum + t - std::lgamma(K * a1p_) + K * std::lgamma(a1p_);
  }
  auto term = [&](int k) {
    return std::lgamma(a1_[k] + a2_[k]) - std::lgamma(a1_[k]) - std::lgamma(a2_[k]) +
           (a1_[k] - a1p_) * Elogv_[k] + (a2_[k] - a2p_) * Elognv_[k];
  };
  double s = 0;
  if (kind_ == kStickBreak) {
    for (int k = 0; k < K; ++k) s += term(k);
    return K * Fp_ + s;
  }
  const int Ko = (int)ord_.size();
  for (int r = 0; r + 1 < Ko; ++r) s += term(ord_[r]);
  return (Ko - 1) * Fp_ + s;
}

// ------------------------------------------------------------ dense helpers --
// The two O(D^3) routines of the host M step.  Their inner loops run along rows; each exists in a baseline build and
// in an AVX2+FMA build chosen at run time (the library itself is compiled for generic x86-64).
#define LCB_CHOLESKY_BODY                                                        \
  for (int j = 0; j < D; ++j) {                                                  \
    double* rj = A + (size_t)j * D;                                              \
    double d = rj[j];                                                            \
    for (int p = 0; p < j; ++p) d -= rj[p] * rj[p];                              \
    if (!(d > 0.0)) return false;                                                \
    d = std::sqrt(d);                                                            \
    rj[j] = d;                                                                   \
    const double invd = 1.0 / d;                                                 \
    int i = j + 1;                                                               \
    /* four rows at a time share the loads of row j */                           \
    for (; i + 4 <= D; i += 4) {                                                 \
      double* r0 = A + (size_t)i * D;                                            \
      double* r1 = r0 + D;                                                       \
      double* r2 = r1 + D;                                                       \
      double* r3 = r2 + D;                                                       \
      double t0 = 0, t1 = 0, t2 = 0, t3 = 0;                                     \
      for (int p = 0; p < j; ++p) {                                              \
        const double x = rj[p];                                                  \
        t0 += r0[p] * x;                                                         \
        t1 += r1[p] * x;                                                         \
        t2 += r2[p] * x;                                                         \
        t3 += r3[p] * x;                                                         \
      }                                                                          \
      r0[j] = (r0[j] - t0) * invd;                                               \
      r1[j] = (r1[j] - t1) * invd;                                               \
      r2[j] = (r2[j] - t2) * invd;                                               \
      r3[j] = (r3[j] - t3) * invd;                                               \
    }                                                                            \
    for (; i < D; ++i) {                                                         \
      double* ri = A + (size_t)i * D;                                            \
      double t = 0;                                                              \
      for (int p = 0; p < j; ++p) t += ri[p] * rj[p];                            \
      ri[j] = (ri[j] - t) * invd;                                                \
    }                                                                            \
    for (int c = j + 1; c < D; ++c) rj[c] = 0.0;                                 \
  }                                                                              \
  return true;

#define LCB_INVERT_BODY                                                          \
  for (int i = 0; i < D; ++i) {                                                  \
    double* ri = Li + (size_t)i * D;                                             \
    const double* li = L + (size_t)i * D;                                        \
    int p = 0;                                                                   \
    /* four rows of L^-1 at a time share the update of row i */                  \
    for (; p + 4 <= i; p += 4) {                                                 \
      const double f0 = li[p], f1 = li[p + 1], f2 = li[p + 2], f3 = li[p + 3];   \
      const double* q0 = Li + (size_t)p * D;                                     \
      const double* q1 = q0 + D;                                                 \
      const double* q2 = q1 + D;                                                 \
      const double* q3 = q2 + D;                                                 \
      for (int c = 0; c <= p; ++c) ri[c] -= f0 * q0[c] + f1 * q1[c] + f2 * q2[c] + f3 * q3[c]; \
      ri[p + 1] -= f1 * q1[p + 1] + f2 * q2[p + 1] + f3 * q3[p + 1];             \
      ri[p + 2] -= f2 * q2[p + 2] + f3 * q3[p + 2];                              \
      ri[p + 3] -= f3 * q3[p + 3];                                               \
    }                                                                            \
    for (; p < i; ++p) {                                                         \
      const double f = li[p];                                                    \
      const double* rp = Li + (size_t)p * D;                                     \
      for (int c = 0; c <= p; ++c) ri[c] -= f * rp[c];                           \
    }                                                                            \
    const double inv = 1.0 / li[i];                                              \
    for (int c = 0; c < i; ++c) ri[c] *= inv;                                    \
    ri[i] = inv;                                                                 \
  }

static bool cholesky_generic(double* A, int D) { LCB_CHOLESKY_BODY }
static void invert_generic(const double* L, int D, double* Li) { LCB_INVERT_BODY }
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2,fma"))) static bool cholesky_avx2(double* A, int D) { LCB_CHOLESKY_BODY }
__attribute__((target("avx2,fma"))) static void invert_avx2(const double* L, int D, double* Li) { LCB_INVERT_BODY }
static bool have_avx2() {
  static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  return ok;
}
#else
static bool have_avx2() { return false; }
static bool cholesky_avx2(double* A, int D) { return cholesky_generic(A, D); }
static void invert_avx2(const double* L, int D, double* Li) { invert_generic(L, D, Li); }
#endif

bool cholesky_lower(std::vector<double>& A, int D) {
  return have_avx2() ? cholesky_avx2(A.data(), D) : cholesky_generic(A.data(), D);
}

void invert_lower(const std::vector<double>& L, int D, std::vector<double>& Li) {
  // row i of L^-1 = (e_i - sum_{p<i} L[i][p] * row p of L^-1) / L[i][i]
  Li.assign((size_t)D * D, 0.0);
  if (have_avx2()) invert_avx2(L.data(), D, Li.data());
  else invert_generic(L.data(), D, Li.data());
}

// --------------------------------------------------------------- clusters --
ClusterPost::ClusterPost(int kind, double clustwidth, int D) : kind_(kind), D_(D), prior_(clustwidth), N_(0) {
  if (kind != kGaussWish && kind != kNormGamma) throw_invalid("unknown cluster distribution kind");
  if (D < 1) throw_invalid("cluster dimensionality must be >= 1");
  if (!(clustwidth > 0)) throw_invalid("clustwidth must be > 0!");
  beta_p_ = kBetaPrior;
  m_p_.assign(D, 0.0);
  if (kind_ == kGaussWish) {
    nu_p_ = D;
    iW_p_.assign((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i) iW_p_[(size_t)i * D + i] = nu_p_ * prior_;
    logdW_p_ = -D * std::log(nu_p_ * prior_);
    F_p_ = 0;
    for (int l = 1; l <= D; ++l) F_p_ += std::lgamma((nu_p_ + 1 - l) / 2);
  } else {
    nu_p_ = kNuPrior;
    iW_p_.assign(D, nu_p_ * prior_);
    logdW_p_ = D * std::log(nu_p_ * prior_);  // holds log L_p summed
    F_p_ = 0;
  }
  x_s_.assign(D, 0.0);
  xx_s_.assign(iW_p_.size(), 0.0);
  clearobs();
}

void ClusterPost::clearobs() {
  nu_ = nu_p_;
  beta_ = beta_p_;
  m_ = m_p_;
  iW_ = iW_p_;
  logdW_ = logdW_p_;
  N_s_ = 0;
  std::fill(x_s_.begin(), x_s_.end(), 0.0);
  std::fill(xx_s_.begin(), xx_s_.end(), 0.0);
  Linv_.clear();
}

void ClusterPost::set_stats(double N_s, const double* x_s, const double* xx_s) {
  N_s_ = N_s;
  std::copy(x_s, x_s + D_, x_s_.begin());
  std::copy(xx_s, xx_s + xx_s_.size(), xx_s_.begin());
}

void ClusterPost::add_stats(double N_s, const double* x_s, const double* xx_s) {
  N_s_ += N_s;
  for (int i = 0; i < D_; ++i) x_s_[i] += x_s[i];
  for (size_t i = 0; i < xx_s_.size(); ++i) xx_s_[i] += xx_s[i];
}

void ClusterPost::add_centred_stats(double n, const double* s, const double* S, const double* c) {
  const int D = D_;
  N_s_ += n;
  for (int i = 0; i < D; ++i) x_s_[i] += s[i] + n * c[i];
  if (kind_ == kGaussWish) {
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j)
        xx_s_[(size_t)i * D + j] += S[(size_t)i * D + j] + c[i] * s[j] + s[i] * c[j] + n * c[i] * c[j];
  } else {
    for (int i = 0; i < D; ++i) xx_s_[i] += S[i] + 2 * c[i] * s[i] + n * c[i] * c[i];
  }
}

void ClusterPost::factor() {
  std::vector<double> L(iW_);
  if (!cholesky_lower(L, D_)) throw_domain("Matrix A is not positive definite.");
  double ld = 0;
  for (int i = 0; i < D_; ++i) ld += 2.0 * std::log(L[(size_t)i * D_ + i]);
  logdW_ = -ld;
  invert_lower(L, D_, Linv_);
}

void ClusterPost::update() {
  update_params();
  if (kind_ == kGaussWish) factor();
}

void ClusterPost::export_factor(double* out) const {
  out[0] = logdW_;
  std::memcpy(out + 1, Linv_.data(), sizeof(double) * (size_t)D_ * D_);
}

void ClusterPost::import_factor(const double* in) {
  logdW_ = in[0];
  Linv_.assign(in + 1, in + 1 + (size_t)D_ * D_);
}

void ClusterPost::update_params() {
  const int D = D_;
  std::vector<double> xbar(D, 0.0);
  if (N_s_ > 0)
    for (int i = 0; i < D; ++i) xbar[i] = x_s_[i] / N_s_;
  N_ = N_s_;
  beta_ = beta_p_ + N_;
  for (int i = 0; i < D; ++i) m_[i] = (beta_p_ * m_p_[i] + x_s_[i]) / beta_;
  if (kind_ == kGaussWish) {
    nu_ = nu_p_ + N_;
    const double w = beta_p_ * N_ / beta_;
    for (int i = 0; i < D; ++i) {
      const double di = xbar[i] - m_p_[i];
      for (int j = 0; j < D; ++j)
        iW_[(size_t)i * D + j] = iW_p_[(size_t)i * D + j] + (xx_s_[(size_t)i * D + j] - xbar[i] * x_s_[j]) +
                                 w * di * (xbar[j] - m_p_[j]);
    }
  } else {
    nu_ = nu_p_ + N_ / 2;
    bool bad = false;
    double ll = 0;
    for (int i = 0; i < D; ++i) {
      double Sk = 0;
      if (N_s_ > 0) Sk = xx_s_[i] - x_s_[i] * x_s_[i] / N_s_;
      const double di = xbar[i] - m_p_[i];
      iW_[i] = iW_p_[i] + Sk / 2 + (beta_p_ * N_ / (2 * beta_)) * di * di;
      if (iW_[i] <= 0) bad = true;
      ll += std::log(iW_[i]);
    }
    if (bad) throw_invalid("Calc log(L): Variance is zero or less!");
    logdW_ = ll;
  }
}

std::vector<double> ClusterPost::cov() const {
  std::vector<double> c(iW_);
  if (kind_ == kGaussWish)
    for (auto& v : c) v /= nu_;
  else
    for (auto& v : c) v *= nu_;
  return c;
}

double ClusterPost::cconst() const {
  const int D = D_;
  if (kind_ == kGaussWish) {
    double sumpsi = 0;
    for (int l = 1; l <= D; ++l) sumpsi += digamma((nu_ + 1 - l) / 2);
    return 0.5 * (sumpsi + logdW_ - D * (1 / beta_ + std::log(kPi)));
  }
  return 0.5 * (D * (digamma(nu_) - std::log(2 * kPi) - 1 / beta_) - logdW_);
}

void ClusterPost::whitener(std::vector<double>& R) const {
  const int D = D_;
  if (kind_ == kGaussWish) {
    R.assign((size_t)D * D, 0.0);
    const double s = std::sqrt(nu_);
    if (Linv_.empty()) {  // cleared state: iW is the diagonal prior
      for (int i = 0; i < D; ++i) R[(size_t)i * D + i] = s / std::sqrt(iW_[(size_t)i * D + i]);
    } else {
      for (size_t i = 0; i < R.size(); ++i) R[i] = s * Linv_[i];
    }
  } else {
    R.resize(D);
    for (int i = 0; i < D; ++i) R[i] = std::sqrt(nu_ / iW_[i]);
  }
}

double ClusterPost::fenergy() const {
  const int D = D_;
  if (kind_ == kGaussWish) {
    double sumpsi = 0, slg = 0;
    for (int l = 1; l <= D; ++l) {
      sumpsi += digamma((nu_ + 1 - l) / 2);
      slg += std::lgamma((nu_ + 1 - l) / 2);
    }
    // trace(iW^-1 iW_p) and (m - m_p)^T iW^-1 (m - m_p) through L^-1
    double tr = 0, mh = 0;
    if (Linv_.empty()) {
      for (int i = 0; i < D; ++i) {
        tr += iW_p_[(size_t)i * D + i] / iW_[(size_t)i * D + i];
        const double d = m_[i] - m_p_[i];
        mh += d * d / iW_[(size_t)i * D + i];
      }
    } else {
      for (int i = 0; i < D; ++i) {
        double z = 0;
        const double* li = &Linv_[(size_t)i * D];
        for (int p = 0; p <= i; ++p) {
          tr += li[p] * li[p] * iW_p_[(size_t)p * D + p];  // iW_p is diagonal
          z += li[p] * (m_[p] - m_p_[p]);
        }
        mh += z * z;
      }
    }
    return F_p_ +
           (D * (beta_p_ / beta_ - 1 - nu_ - std::log(beta_p_ / beta_)) + nu_ * (tr + beta_p_ * mh) +
            nu_p_ * (logdW_p_ - logdW_) + N_ * sumpsi) / 2 -
           slg;
  }
  double a = 0, b = 0;
  for (int i = 0; i < D; ++i) {
    const double d = m_[i] - m_p_[i];
    a += d * d / iW_[i];
    b += iW_p_[i] / iW_[i];
  }
  const unsigned Du = (unsigned)D;  // the reference's D/2 is unsigned integer division
  return Du * (std::lgamma(nu_p_) - std::lgamma(nu_) + N_ * digamma(nu_) / 2 - nu_) +
         (Du / 2) * (std::log(beta_) - std::log(beta_p_) - 1 + beta_p_ / beta_) + beta_p_ * nu_ / 2 * a +
         nu_p_ * (logdW_ - logdW_p_) + nu_ * b;
}

void ClusterPost::split_direction(std::vector<double>& v) const {
  const int D = D_;
  v.assign(D, 0.0);
  if (kind_ == kNormGamma) {
    int e = 0;
    for (int i = 1; i < D; ++i)
      if (iW_[i] > iW_[e]) e = i;
    v[e] = 1.0;
    return;
  }
  if (D == 1) {
    v[0] = 1.0;
    return;
  }
  std::vector<double> w(D), prev(D);
  double nrm = 0;
  for (int i = 0; i < D; ++i) {
    w[i] = (i == D - 1) ? 1.0 : -1.0 + i * (2.0 / (D - 1));
    nrm += w[i] * w[i];
  }
  nrm = std::sqrt(nrm);
  for (int i = 0; i < D; ++i) v[i] = w[i] / nrm;
  double dist = INFINITY;
  for (int it = 0; dist > kEigConThresh && it < kEigMaxIter; ++it) {
    prev = v;
    nrm = 0;
    for (int i = 0; i < D; ++i) {
      double s = 0;
      const double* row = &iW_[(size_t)i * D];
      for (int j = 0; j < D; ++j) s += row[j] * prev[j];
      w[i] = s;
      nrm += s * s;
    }
    nrm = std::sqrt(nrm);
    dist = 0;
    for (int i = 0; i < D; ++i) {
      v[i] = w[i] / nrm;
      dist += (v[i] - prev[i]) * (v[i] - prev[i]);
    }
    dist = std::sqrt(dist);
  }
}

}  // namespace lcb
