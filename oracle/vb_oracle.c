/*
 * vb_oracle.c -- CPU fp64 restatement of libcluster's variational E/M loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * engine in libcluster_b200/.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  It is never
 * linked into, imported by, or used as a fallback for the product library.
 *
 * PARITY PINNING: the reference (dsteinberg/libcluster @ c877625) ships no golden
 * vectors and no numeric assertions.  This restatement
 * is pinned against the REFERENCE'S OWN SOURCES instead: `make -C oracle ref`
 * compiles /root/reference/src/{cluster,distributions,probutils,comutils}.cpp
 * where they lie into oracle/_ref/libcluster_ref.so, against minimal stand-ins
 * for the absent third-party headers (oracle/refshim: Eigen/Dense,
 * boost/math/special_functions.hpp).  tests/test_reference_pin.py requires this
 * file to reproduce that library (F, K, qZ, posteriors; 1e-12) on the reference's
 * test fixture for all six learnXXX, on direct vbem<W,C>() calls and on split /
 * sparse runs; oracle/np_oracle.py is an independent third opinion.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Storage here is plain row-major double arrays; no Eigen.
 * Third-party arithmetic restated: Eigen LDLT -> unpivoted Cholesky (same
 * factor up to rounding for SPD input); boost::math::digamma -> recurrence +
 * asymptotic series; boost::math::lgamma / std::lgamma -> libm lgamma.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants: include/libcluster.h:122-127, include/distributions.h:39-43,
 *      src/probutils.cpp:39-40 (float literals widened to double) ---------- */
static const int SPLITITER = 15;
static const double CONVERGE = (double)1e-5f;
static const double FENGYDEL = (double)1e-5f / 10;
static const double ZEROCUTOFF = (double)0.1f;
static const double BETAPRIOR = 1.0, NUPRIOR = 1.0, ALPHA1PRIOR = 1.0,
                    ALPHA2PRIOR = 1.0;
static const double EIGCONTHRESH = (double)1.0e-8f;
static const int MAXITER = 100;
#define PI 3.14159265358979323846264338327950288

enum { ORC_OK = 0, ORC_INVALID = 1, ORC_RUNTIME = 2, ORC_DOMAIN = 3 };
enum { W_DIRICHLET = 0, W_STICKBREAK = 1, W_GDIRICHLET = 2 };
enum { C_GAUSSWISH = 0, C_NORMGAMMA = 1 };
/* model ids shared with include/libcluster_b200.h */
enum { M_VDP = 0, M_BGMM = 1, M_DGMM = 2, M_GMC = 3, M_SGMC = 4, M_DGMC = 5 };

static char g_err[256];
const char *orc_last_error(void) { return g_err; }
static int fail(int code, const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}

/* ------------------------------------------------------------------------ */
/* special functions and small dense linear algebra                          */
/* ------------------------------------------------------------------------ */

/* boost::math::digamma stand-in (call sites distributions.cpp:160-162,255,360,
 * 391,490,513; probutils.cpp:213).  x > 0 on every path used here. */
double orc_digamma(double x) {
  double r = 0.0;
  while (x < 10.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  double f = 1.0 / (x * x);
  /* psi(x) ~ ln x - 1/2x - sum B_2n / (2n x^2n) */
  double t = f * (-1.0 / 12 +
             f * (1.0 / 120 +
             f * (-1.0 / 252 +
             f * (1.0 / 240 +
             f * (-1.0 / 132 +
             f * (691.0 / 32760 +
             f * (-1.0 / 12)))))));
  return r + log(x) - 0.5 / x + t;
}

/* In-place lower Cholesky A = L L^T of a row-major DxD matrix; returns 0 if a
 * pivot is <= 0 (reference: LDLT vectorD() <= 0 checks, probutils.cpp:131,198).
 */
static int chol_lower(double *A, int D) {
  for (int j = 0; j < D; ++j) {
    double s = A[j * D + j];
    for (int p = 0; p < j; ++p) s -= A[j * D + p] * A[j * D + p];
    if (!(s > 0.0)) return 0;
    double d = sqrt(s);
    A[j * D + j] = d;
    for (int i = j + 1; i < D; ++i) {
      double t = A[i * D + j];
      for (int p = 0; p < j; ++p) t -= A[i * D + p] * A[j * D + p];
      A[i * D + j] = t / d;
    }
    for (int i = 0; i < j; ++i) A[i * D + j] = 0.0;
  }
  return 1;
}

/* probutils::logdet, src/probutils.cpp:189-202 */
static int logdet(const double *A, int D, double *out) {
  double *L = (double *)malloc(sizeof(double) * D * D);
  memcpy(L, A, sizeof(double) * D * D);
  int ok = chol_lower(L, D);
  double s = 0;
  if (ok)
    for (int i = 0; i < D; ++i) s += 2.0 * log(L[i * D + i]);
  free(L);
  *out = s;
  return ok;
}

/* solve L z = b in place (forward substitution) */
static void fwd_solve(const double *L, int D, double *b) {
  for (int i = 0; i < D; ++i) {
    double t = b[i];
    for (int p = 0; p < i; ++p) t -= L[i * D + p] * b[p];
    b[i] = t / L[i * D + i];
  }
}

/* probutils::mahaldist, src/probutils.cpp:113-138: (x-mu)^T A^-1 (x-mu) for
 * every row of X.  Returns 0 if A is not positive definite. */
static int mahaldist(const double *X, int64_t N, int D, const double *mu,
                     const double *A, double *out) {
  double *L = (double *)malloc(sizeof(double) * D * D);
  double *z = (double *)malloc(sizeof(double) * D);
  memcpy(L, A, sizeof(double) * D * D);
  if (!chol_lower(L, D)) {
    free(L);
    free(z);
    return 0;
  }
  for (int64_t n = 0; n < N; ++n) {
    for (int d = 0; d < D; ++d) z[d] = X[n * D + d] - mu[d];
    fwd_solve(L, D, z);
    double s = 0;
    for (int d = 0; d < D; ++d) s += z[d] * z[d];
    out[n] = s;
  }
  free(L);
  free(z);
  return 1;
}

/* probutils::eigpower, src/probutils.cpp:153-186 */
static double eigpower(const double *A, int D, double *eigvec) {
  if (D == 1) {
    eigvec[0] = 1.0;
    return A[0];
  }
  double *v = (double *)malloc(sizeof(double) * D);
  double *o = (double *)malloc(sizeof(double) * D);
  for (int i = 0; i < D; ++i) v[i] = -1.0 + i * (2.0 / (D - 1));
  v[D - 1] = 1.0;
  double eigval = 0;
  for (int i = 0; i < D; ++i) eigval += v[i] * v[i];
  eigval = sqrt(eigval);
  double vdist = INFINITY;
  for (int i = 0; i < D; ++i) eigvec[i] = v[i] / eigval;
  for (int it = 0; vdist > EIGCONTHRESH && it < MAXITER; ++it) {
    memcpy(o, eigvec, sizeof(double) * D);
    for (int i = 0; i < D; ++i) {
      double s = 0;
      for (int j = 0; j < D; ++j) s += A[i * D + j] * o[j];
      v[i] = s;
    }
    eigval = 0;
    for (int i = 0; i < D; ++i) eigval += v[i] * v[i];
    eigval = sqrt(eigval);
    vdist = 0;
    for (int i = 0; i < D; ++i) {
      eigvec[i] = v[i] / eigval;
      vdist += (eigvec[i] - o[i]) * (eigvec[i] - o[i]);
    }
    vdist = sqrt(vdist);
  }
  free(v);
  free(o);
  return eigval;
}

/* ------------------------------------------------------------------------ */
/* weight posteriors: src/distributions.cpp:83-266                           */
/* ------------------------------------------------------------------------ */
typedef struct {
  int kind;
  double a1p, a2p, Fp; /* Dirichlet: a1p = alpha_p */
  int K;
  double *Nk, *a1, *a2, *Elogv, *Elognv, *Elogpi;
  int *ord; /* cluster ids sorted by Nk descending (ordvec[].first) */
} Weight;

static void w_alloc(Weight *w, int K) {
  w->K = K;
  w->Nk = (double *)realloc(w->Nk, sizeof(double) * K);
  w->a1 = (double *)realloc(w->a1, sizeof(double) * K);
  w->a2 = (double *)realloc(w->a2, sizeof(double) * K);
  w->Elogv = (double *)realloc(w->Elogv, sizeof(double) * K);
  w->Elognv = (double *)realloc(w->Elognv, sizeof(double) * K);
  w->Elogpi = (double *)realloc(w->Elogpi, sizeof(double) * K);
  w->ord = (int *)realloc(w->ord, sizeof(int) * K);
}

/* ctors: distributions.cpp:83-121 (StickBreak/GDirichlet), :222-239 (Dirichlet).
 * prior <= 0 selects the default-constructed object. */
static int w_init(Weight *w, int kind, double prior) {
  memset(w, 0, sizeof *w);
  w->kind = kind;
  w->a1p = prior > 0 ? prior : ALPHA1PRIOR;
  w->a2p = ALPHA2PRIOR;
  w_alloc(w, 1);
  w->Nk[0] = 0; /* WeightDist(): Nk = Zero(1), distributions.h:95 */
  w->a1[0] = w->a1p;
  w->a2[0] = w->a2p;
  w->Elogv[0] = w->Elognv[0] = w->Elogpi[0] = 0;
  w->ord[0] = 0;
  /* priorfcalc, distributions.cpp:116-121 */
  w->Fp = lgamma(w->a1p) + lgamma(w->a2p) - lgamma(w->a1p + w->a2p);
  return ORC_OK;
}
static void w_free(Weight *w) {
  free(w->Nk); free(w->a1); free(w->a2); free(w->Elogv); free(w->Elognv);
  free(w->Elogpi); free(w->ord);
  memset(w, 0, sizeof *w);
}

/* Descending order of Nk.  The reference uses std::sort (distributions.cpp:146)
 * whose tie order is implementation-defined; libstdc++ uses insertion sort for
 * <= 16 elements, which is stable, so a stable sort reproduces it there. */
static void order_desc(const double *Nk, int K, int *ord) {
  for (int k = 0; k < K; ++k) ord[k] = k;
  for (int i = 1; i < K; ++i) {
    int c = ord[i], j = i - 1;
    while (j >= 0 && Nk[ord[j]] < Nk[c]) {
      ord[j + 1] = ord[j];
      --j;
    }
    ord[j + 1] = c;
  }
}

static void w_update(Weight *w, const double *Nk, int K) {
  w_alloc(w, K);
  memcpy(w->Nk, Nk, sizeof(double) * K);
  if (w->kind == W_DIRICHLET) { /* distributions.cpp:242-256 */
    double s = 0;
    for (int k = 0; k < K; ++k) {
      w->a1[k] = w->a1p + Nk[k];
      s += w->a1[k];
    }
    double ps = orc_digamma(s);
    for (int k = 0; k < K; ++k) w->Elogpi[k] = orc_digamma(w->a1[k]) - ps;
    return;
  }
  /* StickBreak::update, distributions.cpp:124-168 */
  double N = 0;
  for (int k = 0; k < K; ++k) {
    w->a1[k] = w->a1p + Nk[k];
    N += Nk[k];
  }
  order_desc(Nk, K, w->ord);
  double cumNk = 0, cumE = 0;
  for (int idx = 0; idx < K; ++idx) {
    int k = w->ord[idx];
    cumNk += Nk[k];
    w->a2[k] = w->a2p + (N - cumNk);
    double psisum = orc_digamma(w->a1[k] + w->a2[k]);
    w->Elogv[k] = orc_digamma(w->a1[k]) - psisum;
    w->Elognv[k] = orc_digamma(w->a2[k]) - psisum;
    w->Elogpi[k] = w->Elogv[k] + cumE;
    cumE += w->Elognv[k];
  }
  if (w->kind == W_GDIRICHLET) { /* distributions.cpp:186-196 */
    int sk = w->ord[K - 1];
    w->Elogpi[sk] = w->Elogpi[sk] - w->Elogv[sk];
    w->Elogv[sk] = 0;
    w->Elognv[sk] = 0;
  }
}

static double w_fenergy(const Weight *w) {
  int K = w->K;
  if (w->kind == W_DIRICHLET) { /* distributions.cpp:259-266 */
    double sa = 0, se = 0, t = 0;
    for (int k = 0; k < K; ++k) {
      sa += w->a1[k];
      se += w->Elogpi[k];
      t += (w->a1[k] - 1) * w->Elogpi[k] - lgamma(w->a1[k]);
    }
    return lgamma(sa) - (w->a1p - 1) * se + t - lgamma(K * w->a1p) +
           K * lgamma(w->a1p);
  }
  if (w->kind == W_STICKBREAK) { /* distributions.cpp:171-179 */
    double s = 0;
    for (int k = 0; k < K; ++k)
      s += lgamma(w->a1[k] + w->a2[k]) - lgamma(w->a1[k]) - lgamma(w->a2[k]) +
           (w->a1[k] - w->a1p) * w->Elogv[k] + (w->a2[k] - w->a2p) * w->Elognv[k];
    return K * w->Fp + s;
  }
  /* GDirichlet::fenergy, distributions.cpp:199-215 */
  double Fpi = 0;
  for (int idx = 0; idx < K - 1; ++idx) {
    int k = w->ord[idx];
    Fpi += lgamma(w->a1[k] + w->a2[k]) - lgamma(w->a1[k]) - lgamma(w->a2[k]) +
           (w->a1[k] - w->a1p) * w->Elogv[k] + (w->a2[k] - w->a2p) * w->Elognv[k];
  }
  return (K - 1) * w->Fp + Fpi;
}

/* ------------------------------------------------------------------------ */
/* cluster posteriors: src/distributions.cpp:273-517                         */
/* ------------------------------------------------------------------------ */
typedef struct {
  int kind, D;
  double prior, N;
  double nu_p, beta_p, logdW_p, F_p;
  double *m_p, *iW_p; /* iW_p: DxD (GaussWish) or D (NormGamma L_p) */
  double nu, beta, logdW;
  double *m, *iW;
  double N_s, *x_s, *xx_s;
} Cluster;

static int csz(const Cluster *c) { return c->kind == C_GAUSSWISH ? c->D * c->D : c->D; }

/* clearobs: distributions.cpp:340-353, :467-480 */
static void c_clearobs(Cluster *c) {
  int D = c->D, S = csz(c);
  c->nu = c->nu_p;
  c->beta = c->beta_p;
  memcpy(c->m, c->m_p, sizeof(double) * D);
  memcpy(c->iW, c->iW_p, sizeof(double) * S);
  c->logdW = c->logdW_p;
  c->N_s = 0;
  memset(c->x_s, 0, sizeof(double) * D);
  memset(c->xx_s, 0, sizeof(double) * S);
}

/* ctors: distributions.cpp:273-298 (GaussWish), :406-423 (NormGamma) */
static int c_init(Cluster *c, int kind, double prior, int D) {
  memset(c, 0, sizeof *c);
  if (prior <= 0) return fail(ORC_INVALID, "clustwidth must be > 0!");
  c->kind = kind;
  c->D = D;
  c->prior = prior;
  c->N = 0;
  int S = csz(c);
  c->m_p = (double *)calloc(D, sizeof(double));
  c->m = (double *)calloc(D, sizeof(double));
  c->x_s = (double *)calloc(D, sizeof(double));
  c->iW_p = (double *)calloc(S, sizeof(double));
  c->iW = (double *)calloc(S, sizeof(double));
  c->xx_s = (double *)calloc(S, sizeof(double));
  c->beta_p = BETAPRIOR;
  if (kind == C_GAUSSWISH) {
    c->nu_p = D;
    for (int i = 0; i < D; ++i) c->iW_p[i * D + i] = c->nu_p * prior;
    double ld;
    if (!logdet(c->iW_p, D, &ld)) return fail(ORC_DOMAIN, "Matrix A is not positive definite.");
    c->logdW_p = -ld;
    c->F_p = 0;
    for (int l = 1; l <= D; ++l) c->F_p += lgamma((c->nu_p + 1 - l) / 2);
  } else {
    c->nu_p = NUPRIOR;
    c->logdW_p = 0; /* holds logL_p */
    for (int i = 0; i < D; ++i) {
      c->iW_p[i] = c->nu_p * prior;
      c->logdW_p += log(c->iW_p[i]);
    }
    c->F_p = 0;
  }
  c_clearobs(c);
  return ORC_OK;
}
static void c_free(Cluster *c) {
  free(c->m_p); free(c->m); free(c->x_s); free(c->iW_p); free(c->iW); free(c->xx_s);
  memset(c, 0, sizeof *c);
}

/* addobs: distributions.cpp:301-313 (GaussWish), :426-438 (NormGamma) */
static void c_addobs(Cluster *c, const double *qk, int64_t qstride,
                     const double *X, int64_t N) {
  int D = c->D;
  for (int64_t n = 0; n < N; ++n) {
    double q = qk[n * qstride];
    const double *x = X + n * D;
    c->N_s += q;
    if (c->kind == C_GAUSSWISH) {
      for (int i = 0; i < D; ++i) {
        double qx = q * x[i];
        c->x_s[i] += qx;
        double *row = c->xx_s + (size_t)i * D;
        for (int j = 0; j < D; ++j) row[j] += qx * x[j];
      }
    } else {
      for (int i = 0; i < D; ++i) {
        double qx = q * x[i];
        c->x_s[i] += qx;
        c->xx_s[i] += qx * x[i];
      }
    }
  }
}

/* update: distributions.cpp:316-337 (GaussWish), :441-464 (NormGamma) */
static int c_update(Cluster *c) {
  int D = c->D;
  double *xk = (double *)calloc(D, sizeof(double));
  if (c->N_s > 0)
    for (int i = 0; i < D; ++i) xk[i] = c->x_s[i] / c->N_s;
  c->N = c->N_s;
  if (c->kind == C_GAUSSWISH) {
    c->nu = c->nu_p + c->N;
    c->beta = c->beta_p + c->N;
    double f = c->beta_p * c->N / c->beta;
    for (int i = 0; i < D; ++i) c->m[i] = (c->beta_p * c->m_p[i] + c->x_s[i]) / c->beta;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) {
        double Sk = c->xx_s[i * D + j] - xk[i] * c->x_s[j];
        c->iW[i * D + j] =
            c->iW_p[i * D + j] + Sk + f * (xk[i] - c->m_p[i]) * (xk[j] - c->m_p[j]);
      }
    double ld;
    int ok = logdet(c->iW, D, &ld);
    free(xk);
    if (!ok) return fail(ORC_DOMAIN, "Matrix A is not positive definite.");
    c->logdW = -ld;
  } else {
    c->beta = c->beta_p + c->N;
    c->nu = c->nu_p + c->N / 2;
    int bad = 0;
    double ll = 0;
    for (int i = 0; i < D; ++i) {
      double Sk = 0;
      if (c->N_s > 0) Sk = c->xx_s[i] - c->x_s[i] * c->x_s[i] / c->N_s;
      c->m[i] = (c->beta_p * c->m_p[i] + c->x_s[i]) / c->beta;
      double d = xk[i] - c->m_p[i];
      c->iW[i] = c->iW_p[i] + Sk / 2 + (c->beta_p * c->N / (2 * c->beta)) * d * d;
      if (c->iW[i] <= 0) bad = 1;
      ll += log(c->iW[i]);
    }
    free(xk);
    if (bad) return fail(ORC_INVALID, "Calc log(L): Variance is zero or less!");
    c->logdW = ll;
  }
  return ORC_OK;
}

/* Eloglike: distributions.cpp:356-370 (GaussWish), :483-492 (NormGamma) */
static int c_eloglike(const Cluster *c, const double *X, int64_t N, double *out) {
  int D = c->D;
  if (c->kind == C_GAUSSWISH) {
    double sumpsi = 0;
    for (int l = 1; l <= D; ++l) sumpsi += orc_digamma((c->nu + 1 - l) / 2);
    if (!mahaldist(X, N, D, c->m, c->iW, out))
      return fail(ORC_INVALID, "Matrix A is not positive definite");
    double base = sumpsi + c->logdW - D * (1 / c->beta + log(PI));
    for (int64_t n = 0; n < N; ++n) out[n] = 0.5 * (base - c->nu * out[n]);
  } else {
    double base = D * (orc_digamma(c->nu) - log(2 * PI) - 1 / c->beta) - c->logdW;
    for (int64_t n = 0; n < N; ++n) {
      double s = 0;
      for (int d = 0; d < D; ++d) {
        double t = X[n * D + d] - c->m[d];
        s += t * t * (1.0 / c->iW[d]);
      }
      out[n] = 0.5 * (base - c->nu * s);
    }
  }
  return ORC_OK;
}

/* fenergy: distributions.cpp:388-399 (GaussWish), :508-517 (NormGamma) */
static double c_fenergy(const Cluster *c) {
  int D = c->D;
  if (c->kind == C_GAUSSWISH) {
    double sumpsi = 0, slg = 0;
    for (int l = 1; l <= D; ++l) {
      sumpsi += orc_digamma((c->nu + 1 - l) / 2);
      slg += lgamma((c->nu + 1 - l) / 2);
    }
    /* trace(iW^-1 iW_p) via Cholesky: column-by-column solves */
    double *L = (double *)malloc(sizeof(double) * D * D);
    double *z = (double *)malloc(sizeof(double) * D);
    memcpy(L, c->iW, sizeof(double) * D * D);
    chol_lower(L, D);
    double tr = 0;
    for (int j = 0; j < D; ++j) {
      for (int i = 0; i < D; ++i) z[i] = c->iW_p[i * D + j];
      fwd_solve(L, D, z);
      /* back substitution L^T y = z, need only y[j] ... do full for clarity */
      for (int i = D - 1; i >= 0; --i) {
        double t = z[i];
        for (int p = i + 1; p < D; ++p) t -= L[p * D + i] * z[p];
        z[i] = t / L[i * D + i];
      }
      tr += z[j];
    }
    free(L);
    free(z);
    double mh;
    mahaldist(c->m, 1, D, c->m_p, c->iW, &mh);
    return c->F_p +
           (D * (c->beta_p / c->beta - 1 - c->nu - log(c->beta_p / c->beta)) +
            c->nu * (tr + c->beta_p * mh) + c->nu_p * (c->logdW_p - c->logdW) +
            c->N * sumpsi) / 2 - slg;
  }
  double a = 0, b = 0;
  for (int d = 0; d < D; ++d) {
    double iL = 1.0 / c->iW[d];
    double t = c->m[d] - c->m_p[d];
    a += t * t * iL;
    b += c->iW_p[d] * iL;
  }
  unsigned int Du = (unsigned int)D; /* distributions.cpp:514: D/2 is integer */
  return Du * (lgamma(c->nu_p) - lgamma(c->nu) + c->N * orc_digamma(c->nu) / 2 - c->nu) +
         (Du / 2) * (log(c->beta) - log(c->beta_p) - 1 + c->beta_p / c->beta) +
         c->beta_p * c->nu / 2 * a + c->nu_p * (c->logdW - c->logdW_p) + c->nu * b;
}

/* splitobs: distributions.cpp:373-385 (GaussWish), :495-505 (NormGamma) */
static void c_splitobs(const Cluster *c, const double *X, int64_t N, uint8_t *out) {
  int D = c->D;
  if (c->kind == C_GAUSSWISH) {
    double *v = (double *)malloc(sizeof(double) * D);
    eigpower(c->iW, D, v);
    for (int64_t n = 0; n < N; ++n) {
      double s = 0;
      for (int d = 0; d < D; ++d) s += (X[n * D + d] - c->m[d]) * v[d];
      out[n] = s >= 0;
    }
    free(v);
  } else {
    int e = 0;
    for (int d = 1; d < D; ++d)
      if (c->iW[d] > c->iW[e]) e = d;
    for (int64_t n = 0; n < N; ++n) out[n] = (X[n * D + e] - c->m[e]) >= 0;
  }
}


/* ------------------------------------------------------------------------ */
/* model state: groups of observations, responsibilities, posteriors          */
/* ------------------------------------------------------------------------ */
typedef struct {
  int J, D;
  int64_t *Nj;   /* rows per group */
  double **X;    /* J row-major [Nj x D] blocks (borrowed or owned) */
  int ownX;
} Data;

typedef struct {
  int J, K;
  int64_t *Nj;
  double **q; /* J row-major [Nj x K] */
} Resp;

static void resp_free(Resp *r) {
  if (r->q)
    for (int j = 0; j < r->J; ++j) free(r->q[j]);
  free(r->q);
  free(r->Nj);
  memset(r, 0, sizeof *r);
}
static void resp_alloc(Resp *r, int J, const int64_t *Nj, int K) {
  r->J = J;
  r->K = K;
  r->Nj = (int64_t *)malloc(sizeof(int64_t) * J);
  r->q = (double **)malloc(sizeof(double *) * J);
  for (int j = 0; j < J; ++j) {
    r->Nj[j] = Nj[j];
    r->q[j] = (double *)calloc((size_t)(Nj[j] * K + 1), sizeof(double));
  }
}

typedef struct { Weight *w; int n; } WVec;
typedef struct { Cluster *c; int n; } CVec;

static void wvec_free(WVec *v) { for (int i = 0; i < v->n; ++i) w_free(&v->w[i]); free(v->w); v->w = NULL; v->n = 0; }
static void cvec_free(CVec *v) { for (int i = 0; i < v->n; ++i) c_free(&v->c[i]); free(v->c); v->c = NULL; v->n = 0; }

/* vector::resize(n, proto): append copies of proto / truncate */
static void wvec_resize(WVec *v, int n, int kind) {
  for (int i = n; i < v->n; ++i) w_free(&v->w[i]);
  v->w = (Weight *)realloc(v->w, sizeof(Weight) * (n > 0 ? n : 1));
  for (int i = v->n; i < n; ++i) w_init(&v->w[i], kind, -1.0); /* W() default */
  v->n = n;
}
static int cvec_resize(CVec *v, int n, int kind, double prior, int D) {
  for (int i = n; i < v->n; ++i) c_free(&v->c[i]);
  v->c = (Cluster *)realloc(v->c, sizeof(Cluster) * (n > 0 ? n : 1));
  for (int i = v->n; i < n; ++i) {
    int rc = c_init(&v->c[i], kind, prior, D);
    if (rc) { v->n = i; return rc; }
  }
  v->n = n;
  return ORC_OK;
}

/* trace of every vbem iteration, for decision-level parity tests */
typedef struct { double *F; int *K; int n, cap; } Trace;
static void trace_push(Trace *t, double F, int K) {
  if (!t) return;
  if (t->n == t->cap) {
    t->cap = t->cap ? 2 * t->cap : 64;
    t->F = (double *)realloc(t->F, sizeof(double) * t->cap);
    t->K = (int *)realloc(t->K, sizeof(int) * t->cap);
  }
  t->F[t->n] = F;
  t->K[t->n] = K;
  t->n++;
}

/* ------------------------------------------------------------------------ */
/* src/cluster.cpp:53-82  updateSS                                            */
/* ------------------------------------------------------------------------ */
static void updateSS(const double *Xj, int64_t Nj, const double *qj, int K,
                     CVec *cl, int sparse, double *Njk) {
  for (int k = 0; k < K; ++k) {
    double s = 0;
    for (int64_t n = 0; n < Nj; ++n) s += qj[n * K + k];
    Njk[k] = s;
  }
  for (int k = 0; k < K; ++k) {
    int full;
    if (!sparse) full = (K > 1) ? 1 : (k == 0);
    else full = Njk[k] >= ZEROCUTOFF;
    if (full) c_addobs(&cl->c[k], qj + k, K, Xj, Nj);
  }
}

/* src/cluster.cpp:91-138  vbexpectation (+ probutils::logsumexp :141-150) */
static int vbexpectation(const double *Xj, int64_t Nj, const Weight *w,
                         const CVec *cl, double *qj, int sparse, double *Fz) {
  int K = cl->n;
  int *full = (int *)malloc(sizeof(int) * K);
  int nful = 0;
  for (int k = 0; k < K; ++k) {
    if (!sparse) full[k] = (K > 1) ? 1 : (k == 0);
    else full[k] = (k < w->K) && (w->Nk[k] >= ZEROCUTOFF);
    nful += full[k];
  }
  double *lq = (double *)malloc(sizeof(double) * (size_t)(Nj * (nful > 0 ? nful : 1) + 1));
  double *e = (double *)malloc(sizeof(double) * (size_t)(Nj + 1));
  int col = 0, rc = ORC_OK;
  for (int k = 0; k < K && !rc; ++k) {
    if (!full[k]) continue;
    rc = c_eloglike(&cl->c[k], Xj, Nj, e);
    for (int64_t n = 0; n < Nj; ++n) lq[n * nful + col] = w->Elogpi[k] + e[n];
    ++col;
  }
  double sum = 0;
  if (!rc) {
    for (int64_t n = 0; n < Nj; ++n) {
      double mx = -INFINITY, se = 0;
      for (int c = 0; c < nful; ++c) mx = fmax(mx, lq[n * nful + c]);
      for (int c = 0; c < nful; ++c) se += exp(lq[n * nful + c] - mx);
      double lz = log(se) + mx;
      sum += lz;
      col = 0;
      for (int k = 0; k < K; ++k) {
        if (full[k]) qj[n * K + k] = exp(lq[n * nful + col++] - lz);
        else qj[n * K + k] = 0;
      }
    }
  }
  free(full);
  free(lq);
  free(e);
  *Fz = -sum;
  return rc;
}

/* src/cluster.cpp:177-239  vbem (with :145-165 fenergy inlined) */
static int vbem(const Data *X, Resp *q, WVec *wts, CVec *cl, int wkind, int ckind,
                double prior, int maxit, int sparse, double *Fout, Trace *tr) {
  int J = X->J, K = q->K, D = X->D;
  wvec_resize(wts, J, wkind);
  int rc = cvec_resize(cl, K, ckind, prior, D);
  if (rc) return rc;
  double F = DBL_MAX, Fold;
  int i = 0;
  double *Njk = (double *)malloc(sizeof(double) * K);
  do {
    Fold = F;
    for (int k = 0; k < K; ++k) c_clearobs(&cl->c[k]);
    for (int j = 0; j < J; ++j) {
      updateSS(X->X[j], X->Nj[j], q->q[j], K, cl, sparse, Njk);
      w_update(&wts->w[j], Njk, K);
    }
    for (int k = 0; k < K; ++k) {
      rc = c_update(&cl->c[k]);
      if (rc) { free(Njk); return rc; }
    }
    double Fz = 0;
    for (int j = 0; j < J; ++j) {
      double f;
      rc = vbexpectation(X->X[j], X->Nj[j], &wts->w[j], cl, q->q[j], sparse, &f);
      if (rc) { free(Njk); return rc; }
      Fz += f;
    }
    double Fw = 0, Fc = 0;
    for (int j = 0; j < J; ++j) Fw += w_fenergy(&wts->w[j]);
    for (int k = 0; k < K; ++k) Fc += c_fenergy(&cl->c[k]);
    F = Fc + Fw + Fz;
    trace_push(tr, F, K);
    if ((F - Fold) / fabs(Fold) > FENGYDEL) {
      free(Njk);
      *Fout = F;
      return fail(ORC_RUNTIME, "Free energy increase!");
    }
  } while ((fabs((Fold - F) / Fold) > CONVERGE) && ((i++ < maxit) || (maxit < 0)));
  free(Njk);
  *Fout = F;
  return ORC_OK;
}

/* src/cluster.cpp:505-552  prune_clusters */
static int prune_clusters(Resp *q, WVec *wts, CVec *cl) {
  int K = cl->n, J = q->J;
  int *keep = (int *)malloc(sizeof(int) * K);
  int newK = 0;
  for (int k = 0; k < K; ++k)
    if (!(cl->c[k].N < ZEROCUTOFF)) keep[newK++] = k;
  if (newK == K) { free(keep); return 0; }
  /* erase empties */
  Cluster *nc = (Cluster *)malloc(sizeof(Cluster) * (newK > 0 ? newK : 1));
  int t = 0;
  for (int k = 0; k < K; ++k) {
    if (t < newK && keep[t] == k) nc[t++] = cl->c[k];
    else c_free(&cl->c[k]);
  }
  free(cl->c);
  cl->c = nc;
  cl->n = newK;
  double *Nk = (double *)malloc(sizeof(double) * (newK > 0 ? newK : 1));
  for (int j = 0; j < J; ++j) {
    int64_t Nj = q->Nj[j];
    double *nq = (double *)calloc((size_t)(Nj * newK + 1), sizeof(double));
    for (int k = 0; k < newK; ++k) {
      double s = 0;
      for (int64_t n = 0; n < Nj; ++n) {
        nq[n * newK + k] = q->q[j][n * K + keep[k]];
        s += nq[n * newK + k];
      }
      Nk[k] = s;
    }
    free(q->q[j]);
    q->q[j] = nq;
    w_update(&wts->w[j], Nk, newK);
  }
  q->K = newK;
  free(Nk);
  free(keep);
  return 1;
}

typedef struct { int k, tally; double Fk; } GreedOrder;
/* src/comutils.h:60-68 greedcomp; insertion sort == libstdc++ std::sort for
 * n <= 16 (tie order beyond that is implementation-defined in the reference) */
static int greedcomp(const GreedOrder *i, const GreedOrder *j) {
  if (i->tally == j->tally) return i->Fk > j->Fk;
  return i->tally < j->tally;
}

/* src/cluster.cpp:367-495  split_gr */
static int split_gr(const Data *X, const WVec *wts, const CVec *cl, Resp *q,
                    int **tally, int *ntally, double F, int maxclusters, int sparse,
                    int wkind, int ckind, int *issplit, Trace *tr) {
  int J = X->J, K = cl->n, D = X->D;
  *issplit = 0;
  if (K >= maxclusters && maxclusters >= 0) return ORC_OK;
  if (*ntally < K) {
    *tally = (int *)realloc(*tally, sizeof(int) * K);
    for (int k = *ntally; k < K; ++k) (*tally)[k] = 0;
    *ntally = K;
  } else if (*ntally > K) {
    *ntally = K; /* vector::resize truncates */
  }
  GreedOrder *ord = (GreedOrder *)malloc(sizeof(GreedOrder) * K);
  for (int k = 0; k < K; ++k) {
    ord[k].k = k;
    ord[k].tally = (*tally)[k];
    ord[k].Fk = c_fenergy(&cl->c[k]);
  }
  int rc = ORC_OK;
  for (int j = 0; j < J && !rc; ++j) {
    int64_t Nj = X->Nj[j];
    double *e = (double *)malloc(sizeof(double) * (size_t)(Nj + 1));
    for (int k = 0; k < K && !rc; ++k) {
      rc = c_eloglike(&cl->c[k], X->X[j], Nj, e);
      double LL = 0;
      for (int64_t n = 0; n < Nj; ++n)
        LL += q->q[j][n * K + k] * (wts->w[j].Elogpi[k] + e[n]);
      ord[k].Fk -= LL;
    }
    free(e);
  }
  if (rc) { free(ord); return rc; }
  for (int i = 1; i < K; ++i) { /* sort(ord, greedcomp), cluster.cpp:418 */
    GreedOrder c = ord[i];
    int p = i - 1;
    while (p >= 0 && greedcomp(&c, &ord[p])) { ord[p + 1] = ord[p]; --p; }
    ord[p + 1] = c;
  }

  for (int oi = 0; oi < K; ++oi) {
    int k = ord[oi].k;
    ++(*tally)[k];
    if (cl->c[k].N < 4) continue;

    /* partobs (comutils.cpp:56-72) + splitobs + qZref, cluster.cpp:438-453 */
    Data Xk;
    Xk.J = J; Xk.D = D; Xk.ownX = 1;
    Xk.Nj = (int64_t *)calloc(J, sizeof(int64_t));
    Xk.X = (double **)calloc(J, sizeof(double *));
    int64_t **mapidx = (int64_t **)calloc(J, sizeof(int64_t *));
    Resp qref;
    int64_t scount = 0, Mtot = 0;
    for (int j = 0; j < J; ++j) {
      int64_t Nj = X->Nj[j], M = 0;
      for (int64_t n = 0; n < Nj; ++n) M += q->q[j][n * K + k] > 0.5;
      Xk.Nj[j] = M;
      Xk.X[j] = (double *)malloc(sizeof(double) * (size_t)(M * D + 1));
      mapidx[j] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(M + 1));
      int64_t m = 0;
      for (int64_t n = 0; n < Nj; ++n)
        if (q->q[j][n * K + k] > 0.5) {
          memcpy(Xk.X[j] + m * D, X->X[j] + n * D, sizeof(double) * D);
          mapidx[j][m++] = n;
        }
      Mtot += M;
    }
    resp_alloc(&qref, J, Xk.Nj, 2);
    for (int j = 0; j < J; ++j) {
      int64_t M = Xk.Nj[j];
      uint8_t *sp = (uint8_t *)malloc((size_t)M + 1);
      c_splitobs(&cl->c[k], Xk.X[j], M, sp);
      for (int64_t m = 0; m < M; ++m) {
        qref.q[j][m * 2 + 0] = sp[m] ? 1.0 : 0.0;
        qref.q[j][m * 2 + 1] = sp[m] ? 0.0 : 1.0;
        scount += sp[m];
      }
      free(sp);
    }

    int skip = (scount < 2) || (scount > (Mtot - 2));
    WVec wspl = {0, 0};
    CVec cspl = {0, 0};
    Resp qaug;
    memset(&qaug, 0, sizeof qaug);
    double Fsplit = 0;
    if (!skip) {
      double Fs;
      rc = vbem(&Xk, &qref, &wspl, &cspl, wkind, ckind, cl->c[0].prior, SPLITITER,
                sparse, &Fs, tr);
      if (!rc) {
        for (int c = 0; c < cspl.n; ++c) /* anyempty, comutils.h:114-123 */
          if (cspl.c[c].N <= 1) skip = 1;
      }
    }
    if (!skip && !rc) {
      /* auglabels, comutils.cpp:75-104 */
      resp_alloc(&qaug, J, q->Nj, K + 1);
      for (int j = 0; j < J; ++j) {
        int64_t Nj = q->Nj[j];
        for (int64_t n = 0; n < Nj; ++n)
          memcpy(qaug.q[j] + n * (K + 1), q->q[j] + n * K, sizeof(double) * K);
        for (int64_t m = 0; m < Xk.Nj[j]; ++m)
          if (qref.q[j][m * 2 + 1] > 0.5) {
            int64_t n = mapidx[j][m];
            qaug.q[j][n * (K + 1) + K] = q->q[j][n * K + k];
            qaug.q[j][n * (K + 1) + k] = 0;
          }
      }
      rc = vbem(X, &qaug, &wspl, &cspl, wkind, ckind, cl->c[0].prior, 1, sparse,
                &Fsplit, tr);
      if (!rc)
        for (int c = 0; c < cspl.n; ++c)
          if (cspl.c[c].N <= 1) skip = 1;
    }
    int accept = 0;
    if (!skip && !rc)
      accept = (Fsplit < F) && (fabs((F - Fsplit) / F) > CONVERGE);
    if (accept) {
      resp_free(q);
      *q = qaug;
      memset(&qaug, 0, sizeof qaug);
      (*tally)[k] = 0;
      *issplit = 1;
    }
    /* cleanup */
    for (int j = 0; j < J; ++j) { free(Xk.X[j]); free(mapidx[j]); }
    free(Xk.X); free(Xk.Nj); free(mapidx);
    resp_free(&qref);
    if (qaug.q) resp_free(&qaug);
    wvec_free(&wspl);
    cvec_free(&cspl);
    if (rc || accept) break;
  }
  free(ord);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* public handle API (ctypes)                                                */
/* ------------------------------------------------------------------------ */
typedef struct {
  Data X;
  Resp q;
  WVec w;
  CVec c;
  int wkind, ckind;
  double F;
  Trace tr;
} OrcModel;

static void model_kinds(int model, int *wk, int *ck) {
  switch (model) {
    case M_VDP: *wk = W_STICKBREAK; *ck = C_GAUSSWISH; break;
    case M_BGMM: *wk = W_DIRICHLET; *ck = C_GAUSSWISH; break;
    case M_DGMM: *wk = W_DIRICHLET; *ck = C_NORMGAMMA; break;
    case M_GMC: *wk = W_GDIRICHLET; *ck = C_GAUSSWISH; break;
    case M_SGMC: *wk = W_DIRICHLET; *ck = C_GAUSSWISH; break;
    default: *wk = W_GDIRICHLET; *ck = C_NORMGAMMA; break;
  }
}

/* Xcat: all groups concatenated, row-major [sum Nj x D]. */
OrcModel *orc_model_create(int model, int J, const double *Xcat, const int64_t *Nj, int D) {
  OrcModel *m = (OrcModel *)calloc(1, sizeof(OrcModel));
  model_kinds(model, &m->wkind, &m->ckind);
  m->X.J = J; m->X.D = D; m->X.ownX = 1;
  m->X.Nj = (int64_t *)malloc(sizeof(int64_t) * J);
  m->X.X = (double **)malloc(sizeof(double *) * J);
  int64_t off = 0;
  for (int j = 0; j < J; ++j) {
    m->X.Nj[j] = Nj[j];
    m->X.X[j] = (double *)malloc(sizeof(double) * (size_t)(Nj[j] * D + 1));
    memcpy(m->X.X[j], Xcat + off * D, sizeof(double) * Nj[j] * D);
    off += Nj[j];
  }
  return m;
}

void orc_model_destroy(OrcModel *m) {
  if (!m) return;
  for (int j = 0; j < m->X.J; ++j) free(m->X.X[j]);
  free(m->X.X); free(m->X.Nj);
  resp_free(&m->q);
  wvec_free(&m->w);
  cvec_free(&m->c);
  free(m->tr.F); free(m->tr.K);
  free(m);
}

/* src/cluster.cpp:564-629 cluster<W,C>() behind the learnXXX wrappers :636-831.
 * weight_prior <= 0 means a default-constructed weight object. */
int orc_learn(OrcModel *m, double prior, double weight_prior, int maxclusters,
              int sparse, unsigned nthreads) {
  if (nthreads < 1) return fail(ORC_INVALID, "Must specify at least one thread for execution!");
  int J = m->X.J;
  resp_free(&m->q);
  resp_alloc(&m->q, J, m->X.Nj, 1);
  for (int j = 0; j < J; ++j)
    for (int64_t n = 0; n < m->X.Nj[j]; ++n) m->q.q[j][n] = 1.0;
  /* caller-supplied weight objects keep their prior (cluster.cpp:653,684) */
  wvec_free(&m->w);
  cvec_free(&m->c);
  m->w.w = (Weight *)malloc(sizeof(Weight) * J);
  m->w.n = J;
  for (int j = 0; j < J; ++j) w_init(&m->w.w[j], m->wkind, weight_prior);
  m->tr.n = 0;
  int *tally = NULL, ntally = 0, issplit = 1, rc = ORC_OK;
  double F = 0;
  while (issplit && !rc) {
    rc = vbem(&m->X, &m->q, &m->w, &m->c, m->wkind, m->ckind, prior, -1, sparse, &F, &m->tr);
    if (rc) break;
    prune_clusters(&m->q, &m->w, &m->c);
    rc = split_gr(&m->X, &m->w, &m->c, &m->q, &tally, &ntally, F, maxclusters, sparse,
                  m->wkind, m->ckind, &issplit, &m->tr);
  }
  free(tally);
  m->F = F;
  return rc;
}

/* One call of vbem() from caller-supplied responsibilities q0 (row-major
 * [N x K] over the concatenated rows); maxit as in cluster.cpp:183. */
int orc_vbem(OrcModel *m, const double *q0, int K, double prior, double weight_prior,
             int maxit, int sparse) {
  int J = m->X.J;
  resp_free(&m->q);
  resp_alloc(&m->q, J, m->X.Nj, K);
  int64_t off = 0;
  for (int j = 0; j < J; ++j) {
    memcpy(m->q.q[j], q0 + off * K, sizeof(double) * m->X.Nj[j] * K);
    off += m->X.Nj[j];
  }
  wvec_free(&m->w);
  cvec_free(&m->c);
  m->w.w = (Weight *)malloc(sizeof(Weight) * J);
  m->w.n = J;
  for (int j = 0; j < J; ++j) w_init(&m->w.w[j], m->wkind, weight_prior);
  m->tr.n = 0;
  return vbem(&m->X, &m->q, &m->w, &m->c, m->wkind, m->ckind, prior, maxit, sparse, &m->F, &m->tr);
}

double orc_F(const OrcModel *m) { return m->F; }
int orc_K(const OrcModel *m) { return m->c.n; }
int orc_trace_len(const OrcModel *m) { return m->tr.n; }
void orc_trace(const OrcModel *m, double *F, int *K) {
  memcpy(F, m->tr.F, sizeof(double) * m->tr.n);
  memcpy(K, m->tr.K, sizeof(int) * m->tr.n);
}
/* qZ of all groups concatenated, row-major [N x K] */
void orc_get_qZ(const OrcModel *m, double *out) {
  int64_t off = 0;
  int K = m->q.K;
  for (int j = 0; j < m->q.J; ++j) {
    memcpy(out + off * K, m->q.q[j], sizeof(double) * m->q.Nj[j] * K);
    off += m->q.Nj[j];
  }
}
int orc_qK(const OrcModel *m) { return m->q.K; }
void orc_get_weights(const OrcModel *m, int j, double *Elogpi, double *Nk) {
  const Weight *w = &m->w.w[j];
  if (Elogpi) memcpy(Elogpi, w->Elogpi, sizeof(double) * w->K);
  if (Nk) memcpy(Nk, w->Nk, sizeof(double) * w->K);
}
int orc_weights_K(const OrcModel *m, int j) { return m->w.w[j].K; }
double orc_weights_fenergy(const OrcModel *m, int j) { return w_fenergy(&m->w.w[j]); }
/* posterior of cluster k: mean [D], iW (DxD or D), scalars {N, nu, beta, logdW};
 * sufficient statistics {N_s, x_s [D], xx_s [DxD or D]} */
void orc_get_cluster(const OrcModel *m, int k, double *scal4, double *mean, double *iW,
                     double *N_s, double *x_s, double *xx_s) {
  const Cluster *c = &m->c.c[k];
  if (scal4) { scal4[0] = c->N; scal4[1] = c->nu; scal4[2] = c->beta; scal4[3] = c->logdW; }
  if (mean) memcpy(mean, c->m, sizeof(double) * c->D);
  if (iW) memcpy(iW, c->iW, sizeof(double) * csz(c));
  if (N_s) *N_s = c->N_s;
  if (x_s) memcpy(x_s, c->x_s, sizeof(double) * c->D);
  if (xx_s) memcpy(xx_s, c->xx_s, sizeof(double) * csz(c));
}
double orc_cluster_fenergy(const OrcModel *m, int k) { return c_fenergy(&m->c.c[k]); }

/* ---- operator-level handles (distributions.h surface) ------------------- */
Weight *orc_weight_new(int kind, double prior) {
  Weight *w = (Weight *)malloc(sizeof(Weight));
  w_init(w, kind, prior);
  return w;
}
void orc_weight_del(Weight *w) { if (w) { w_free(w); free(w); } }
void orc_weight_update(Weight *w, const double *Nk, int K) { w_update(w, Nk, K); }
int orc_weight_K(const Weight *w) { return w->K; }
void orc_weight_elogweight(const Weight *w, double *out) { memcpy(out, w->Elogpi, sizeof(double) * w->K); }
void orc_weight_getNk(const Weight *w, double *out) { memcpy(out, w->Nk, sizeof(double) * w->K); }
double orc_weight_fenergy(const Weight *w) { return w_fenergy(w); }

Cluster *orc_cluster_new(int kind, double prior, int D) {
  Cluster *c = (Cluster *)malloc(sizeof(Cluster));
  if (c_init(c, kind, prior, D)) { free(c); return NULL; }
  return c;
}
void orc_cluster_del(Cluster *c) { if (c) { c_free(c); free(c); } }
void orc_cluster_addobs(Cluster *c, const double *qk, const double *X, int64_t N) { c_addobs(c, qk, 1, X, N); }
int orc_cluster_update(Cluster *c) { return c_update(c); }
void orc_cluster_clearobs(Cluster *c) { c_clearobs(c); }
int orc_cluster_eloglike(const Cluster *c, const double *X, int64_t N, double *out) { return c_eloglike(c, X, N, out); }
double orc_cluster_fenergy1(const Cluster *c) { return c_fenergy(c); }
void orc_cluster_splitobs(const Cluster *c, const double *X, int64_t N, uint8_t *out) { c_splitobs(c, X, N, out); }
double orc_cluster_getN(const Cluster *c) { return c->N; }
void orc_cluster_get(const Cluster *c, double *scal4, double *mean, double *iW,
                     double *N_s, double *x_s, double *xx_s) {
  if (scal4) { scal4[0] = c->N; scal4[1] = c->nu; scal4[2] = c->beta; scal4[3] = c->logdW; }
  if (mean) memcpy(mean, c->m, sizeof(double) * c->D);
  if (iW) memcpy(iW, c->iW, sizeof(double) * csz(c));
  if (N_s) *N_s = c->N_s;
  if (x_s) memcpy(x_s, c->x_s, sizeof(double) * c->D);
  if (xx_s) memcpy(xx_s, c->xx_s, sizeof(double) * csz(c));
}

/* ---- timed body for the CPU baseline: one vbem iteration at fixed K ------
 * (cluster.cpp:203-226 loop body, reference loop structure: per-cluster addobs
 * then per-cluster Eloglike).  Returns F; q is updated in place. */
int orc_vbem_iteration(OrcModel *m, double prior, double *F) {
  return vbem(&m->X, &m->q, &m->w, &m->c, m->wkind, m->ckind, prior, 0, 0, F, NULL);
}
