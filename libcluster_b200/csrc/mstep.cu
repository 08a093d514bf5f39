// mstep.cu -- the M step of one VB iteration on the device.
//
// Reference: the posterior updates of src/cluster.cpp:203-217 -- GaussWish::update (src/distributions.cpp:316-337,
// logdet src/probutils.cpp:189-202), NormGamma::update (:441-464), Dirichlet / StickBreak / GDirichlet::update
// (:242-256, :124-168, :186-196) -- and the parameter part of the free energy (src/cluster.cpp:155-162,
// distributions.cpp:171-179, 199-215, 259-266, 388-399, 508-517).  host_model.cpp holds the same formulas for the
// host; this file evaluates them where the statistics are, in fp64, and writes the operands of the E pass in the
// layouts its kernels read (kernels.cuh, tc_kernels.cuh), so that an iteration needs no host arithmetic.
//
//   mstep_cluster_kernel   one CTA per cluster: un-centre the statistics, posterior parameters, Cholesky of iW in
//                          shared memory, in-place inverse of the factor, log det, cluster constant, Fc_k, operands
//   mstep_weights_kernel   one CTA per group: stick ordering (std::sort restated, mstep_math.hpp), E[log pi], Fw_j
//   mstep_finish_kernel    one CTA: quantities that need all clusters (mean constant, level-1 centring blocks and
//                          error constants, operand scale of the next scatter), the iteration record
#include "mstep.cuh"

#include <cuda_fp16.h>

#include <cfloat>
#include <cmath>

#include "mstep_math.hpp"

namespace lcb {
namespace dev {

namespace {
constexpr int kNT = 256;
constexpr double kPi = 3.141592653589793238462643383279502884;
enum { kGaussWish = 0, kNormGamma = 1 };
enum { kDirichlet = 0, kStickBreak = 1, kGDirichlet = 2 };
constexpr size_t kSmemWorkLimit = 200 * 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < kNT / 32; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = red[0];
  for (int i = 1; i < kNT / 32; ++i) t = fmax(t, red[i]);
  return t;
}

template <typename T> __device__ __forceinline__ void split_hi_lo(double v, T& hi, T& lo);
template <> __device__ __forceinline__ void split_hi_lo<float>(double v, float& hi, float& lo) {
  hi = (float)v;
  lo = (float)(v - (double)hi);
}
template <> __device__ __forceinline__ void split_hi_lo<double>(double v, double& hi, double& lo) {
  hi = v;
  lo = 0.0;
}

// offsets inside one cluster's tcgen05 operand blob (tc_kernels.cu: kAPart, kOffMean)
constexpr uint32_t kBlobBytes = 50176, kAPart = 16384, kOffMean = 49152;

template <typename T>
__global__ void __launch_bounds__(kNT) mstep_cluster_kernel(MStepArgs a, int work_in_smem) {
  extern __shared__ double sm[];
  if (a.stats[a.nstat] != 0.0) return;
  const int k = blockIdx.x, D = a.D, K = a.K, J = a.J, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const bool full = a.ckind == kGaussWish;
  const int lda = D | 1;
  double* c = sm;         // centre of this cluster's statistics, raw coordinates
  double* s0 = c + D;     // centred first moment
  double* xs = s0 + D;    // raw first moment
  double* xbar = xs + D;
  double* m = xbar + D;   // posterior mean, raw coordinates
  double* rel = m + D;    // m - data centre
  double* col = rel + D;
  double* Ld = col + D;   // diagonal of the Cholesky factor
  double* red = Ld + D;   // [16]
  double* A = work_in_smem ? red + 16 : a.work + (size_t)k * D * lda;
  const T* cenk = reinterpret_cast<const T*>(a.cen) + (size_t)k * a.cld;
  double* post = a.post + (size_t)k * kPostStride;
  const int64_t Sz = full ? (int64_t)D * D : D;
  double* rawk = a.raw + (size_t)k * (1 + D + Sz);

  // N_s over the groups that contribute (cluster.cpp:69-70)
  if (tid == 0) {
    double n = 0;
    for (int j = 0; j < J; ++j)
      if (a.act == nullptr || a.act[(size_t)j * K + k]) n += a.stats[(size_t)j * K + k];
    red[0] = n;
  }
  __syncthreads();
  const double n = red[0];
  __syncthreads();
  const double* xsk = a.stats + (size_t)J * K + (size_t)k * D;
  const double* Sk = a.stats + (size_t)J * K + (size_t)K * D + (size_t)k * Sz;
  const double beta_p = 1.0;
  const double beta = beta_p + n;
  for (int d = tid; d < D; d += kNT) {
    const double cd = (double)cenk[d] + a.centre[d];
    const double sd = xsk[d];
    const double x = sd + n * cd;  // ClusterPost::add_centred_stats
    c[d] = cd;
    s0[d] = sd;
    xs[d] = x;
    xbar[d] = n > 0 ? x / n : 0.0;
    const double md = (beta_p * 0.0 + x) / beta;
    m[d] = md;
    rel[d] = md - a.centre[d];
    rawk[1 + d] = x;
  }
  if (tid == 0) rawk[0] = n;
  __syncthreads();

  double nu, logdW, cconst, Fc, cmax = 0, fail = 0;
  double s_scale = 1, t_scale = 1, rfro = 0, vmaxk = 0;
  if (full) {
    const double nu_p = (double)D;
    nu = nu_p + n;
    const double ipw = nu_p * a.prior;  // diagonal of iW_p
    const double w = beta_p * n / beta;
    for (int idx = tid; idx < D * D; idx += kNT) {
      const int i = idx / D, j = idx - i * D;
      const double xx = Sk[idx] + c[i] * s0[j] + s0[i] * c[j] + n * c[i] * c[j];
      rawk[1 + D + idx] = xx;
      if (j <= i) A[(size_t)i * lda + j] = (i == j ? ipw : 0.0) + (xx - xbar[i] * xs[j]) + w * xbar[i] * xbar[j];
    }
    __syncthreads();
    double cvmax = 0;
    for (int d = tid; d < D; d += kNT) cvmax = fmax(cvmax, A[(size_t)d * lda + d] / nu);
    cvmax = block_max(cvmax, red);
    // ---- Cholesky, right-looking, lower triangle in place; the diagonal of L goes to Ld ----
    bool bad = false;
    for (int j = 0; j < D; ++j) {
      const double dj = A[(size_t)j * lda + j];
      if (!(dj > 0.0)) {
        bad = true;
        break;
      }
      const double l = sqrt(dj), inv = 1.0 / l;
      if (tid == 0) Ld[j] = l;
      for (int i = j + 1 + tid; i < D; i += kNT) {
        const double v = A[(size_t)i * lda + j] * inv;
        A[(size_t)i * lda + j] = v;
        col[i] = v;
      }
      __syncthreads();
      for (int i = j + 1 + warp; i < D; i += kNT / 32) {
        const double li = col[i];
        double* row = A + (size_t)i * lda;
        for (int cc = j + 1 + lane; cc <= i; cc += 32) row[cc] -= li * col[cc];
      }
      __syncthreads();
    }
    if (bad) {
      // Matrix A is not positive definite (probutils.cpp:198-199)
      if (tid == 0) {
        post[kPFail] = 1.0;
        post[kPN] = n;
        post[kPCconst] = 0.0;
        post[kPFc] = 0.0;
        post[kPVmax] = 0.0;
        post[kPCmax] = 0.0;
      }
      return;
    }
    double ld = 0;
    for (int d = tid; d < D; d += kNT) ld += 2.0 * log(Ld[d]);
    logdW = -block_sum(ld, red);
    // ---- inverse of the factor in place, last column first: two lanes share a row ----
    for (int j = D - 1; j >= 0; --j) {
      for (int p = j + 1 + tid; p < D; p += kNT) col[p] = A[(size_t)p * lda + j];
      __syncthreads();
      const double dinv = 1.0 / Ld[j];
      // the trip count is rounded up to whole warps: every lane of a warp that enters reaches the full-mask shuffle
      // (lanes past the last row carry acc = 0); a partial warp at the shuffle deadlocks
      const int rcount = (2 * (D - 1 - j) + 2 + 31) & ~31;
      for (int r = tid; r < rcount; r += kNT) {
        const int i = j + 1 + (r >> 1), h = r & 1;  // the pair (2 lanes) of row i
        double acc = 0;
        if (i < D) {
          const double* row = A + (size_t)i * lda;
          for (int p = j + 1 + h; p <= i; p += 2) acc += row[p] * col[p];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (i < D && h == 0) A[(size_t)i * lda + j] = -acc * dinv;
      }
      if (tid == 0) A[(size_t)j * lda + j] = dinv;
      __syncthreads();
    }
    // ---- per-row sums over L^-1 ----
    const double sqn = sqrt(nu);
    double tr = 0, mh = 0, fro = 0, rmax = 0;
    for (int i = tid; i < D; i += kNT) {
      const double* row = A + (size_t)i * lda;
      double z = 0, zr = 0, t = 0, f = 0, rm = 0;
      for (int p = 0; p <= i; ++p) {
        const double li = row[p];
        t += li * li * ipw;
        z += li * m[p];
        const double r = sqn * li;
        zr += r * rel[p];
        f += r * r;
        rm = fmax(rm, fabs(r));
      }
      tr += t;
      mh += z * z;
      fro += f;
      rmax = fmax(rmax, rm);
      col[i] = zr;  // (R rel)_i
    }
    tr = block_sum(tr, red);
    mh = block_sum(mh, red);
    fro = block_sum(fro, red);
    rmax = block_max(rmax, red);
    double sumpsi = 0, slg = 0;
    for (int l = 1 + tid; l <= D; l += kNT) {
      sumpsi += mm::digamma((nu + 1 - l) / 2);
      slg += lgamma((nu + 1 - l) / 2);
    }
    sumpsi = block_sum(sumpsi, red);
    slg = block_sum(slg, red);
    cconst = 0.5 * (sumpsi + logdW - D * (1 / beta + log(kPi)));
    const double logdW_p = -D * log(nu_p * a.prior);
    Fc = a.Fp + (D * (beta_p / beta - 1 - nu - log(beta_p / beta)) + nu * (tr + beta_p * mh) + nu_p * (logdW_p - logdW) +
                 n * sumpsi) / 2 - slg;
    rfro = sqrt(fro);

    if (a.path == 0) {
      const int DP = a.cld;
      T* RT = reinterpret_cast<T*>(a.RT) + (size_t)k * DP * DP;
      for (int idx = tid; idx < DP * DP; idx += kNT) {
        const int d = idx / DP, i = idx - d * DP;
        RT[idx] = (i < D && d <= i) ? (T)(sqn * A[(size_t)i * lda + d]) : (T)0;
      }
    } else {
      // tcgen05 operands (Engine::ephase_tc): a = s (x - m), b = (t / s) R with power-of-two s, t
      int es = (int)lround(log2(32.0 / sqrt(fmax(cvmax, 1e-300))));
      es = min(60, max(-60, es));
      s_scale = ldexp(1.0, es);
      int et = 8 - mm::ceil_log2(fmax(rmax / s_scale, 1e-300));
      et = min(100, max(-100, et));
      t_scale = ldexp(1.0, et);
      const double bscale = t_scale / s_scale;
      uint8_t* out = a.blob + (size_t)k * kBlobBytes;
      // 16-byte chunks: 8 consecutive k entries of one row; K block 0: rows 0..127, K block 1: rows 64..127
      for (int idx = tid; idx < 128 * 8 + 64 * 8; idx += kNT) {
        const int kbk = idx >= 1024 ? 1 : 0;
        const int loc = idx - 1024 * kbk;
        const int r = loc >> 3, ch = loc & 7;
        const int i = r + 64 * kbk, d0 = 64 * kbk + 8 * ch;
        const uint32_t base_hi = kbk == 0 ? 0u : 2 * kAPart;
        const uint32_t base_lo = base_hi + (kbk == 0 ? kAPart : kAPart / 2);
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int d = d0 + e;
          const float v = (d <= i && i < D) ? (float)((sqn * A[(size_t)i * lda + d]) * bscale) : 0.f;
          hi[e] = __float2half_rn(v);
          lo[e] = __float2half_rn(v - __half2float(hi[e]));
        }
        const uint32_t off = (uint32_t)r * 128u + (((uint32_t)ch ^ ((uint32_t)r & 7u)) << 4);
        *reinterpret_cast<uint4*>(out + base_hi + off) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(out + base_lo + off) = *reinterpret_cast<const uint4*>(lo);
      }
      float* mh_out = reinterpret_cast<float*>(out + kOffMean);
      for (int d = tid; d < 128; d += kNT) {  // the blob keeps its 128-dimensional layout for D = 64 (zeros above)
        const double rd = d < D ? rel[d] : 0.0;
        const float hi = (float)rd;
        mh_out[d] = hi;
        mh_out[128 + d] = (float)(-(rd - (double)hi) * s_scale);
      }
      if (tid == 0) {
        a.as[k] = (float)s_scale;
        a.it2[k] = (float)(1.0 / (t_scale * t_scale));
      }
      if (a.two_level) {
        // accumulator of level 1 = s_g tau (R x - R m): the centring term
        double vm = 0;
        for (int i = tid; i < D; i += kNT) {
          const double val = a.sg * bscale * col[i];
          a.vaug[(size_t)k * D + i] = val;
          vm = fmax(vm, fabs(val));
        }
        vmaxk = block_max(vm, red);
        // a NaN must reach the finish kernel (fmax drops it)
        double nanflag = 0;
        for (int i = tid; i < D; i += kNT) nanflag += isfinite(col[i]) ? 0.0 : 1.0;
        if (block_sum(nanflag, red) > 0) vmaxk = INFINITY;
      }
    }
  } else {
    // ---- NormGamma (distributions.cpp:441-464) ----
    const double nu_p = 1.0, Lp = nu_p * a.prior;
    nu = nu_p + n / 2;
    double ll = 0, bad = 0, fa = 0, fb = 0;
    T* Aout = reinterpret_cast<T*>(a.RT) + (size_t)k * D;
    for (int d = tid; d < D; d += kNT) {
      const double xx = Sk[d] + 2 * c[d] * s0[d] + n * c[d] * c[d];
      rawk[1 + D + d] = xx;
      double Skd = 0;
      if (n > 0) Skd = xx - xs[d] * xs[d] / n;
      const double L = Lp + Skd / 2 + (beta_p * n / (2 * beta)) * xbar[d] * xbar[d];
      if (L <= 0) bad = 1;
      ll += log(L);
      fa += m[d] * m[d] / L;
      fb += Lp / L;
      const double r = sqrt(nu / L);
      Aout[d] = (T)(r * r);
    }
    ll = block_sum(ll, red);
    bad = block_sum(bad, red);
    fa = block_sum(fa, red);
    fb = block_sum(fb, red);
    if (bad > 0) fail = 2;
    logdW = ll;
    cconst = 0.5 * (D * (mm::digamma(nu) - log(2 * kPi) - 1 / beta) - logdW);
    const unsigned Du = (unsigned)D;
    const double logLp = D * log(Lp);
    Fc = Du * (lgamma(nu_p) - lgamma(nu) + n * mm::digamma(nu) / 2 - nu) +
         (Du / 2) * (log(beta) - log(beta_p) - 1 + beta_p / beta) + beta_p * nu / 2 * fa + nu_p * (logdW - logLp) + nu * fb;
  }

  // mean of the cluster in the operand tables, centre of the next statistics pass
  {
    T* cen_out = reinterpret_cast<T*>(a.cen) + (size_t)k * a.cld;
    T* mhi = a.path == 0 ? reinterpret_cast<T*>(a.mhi) + (size_t)k * a.cld : nullptr;
    T* mlo = a.path == 0 ? reinterpret_cast<T*>(a.mlo) + (size_t)k * a.cld : nullptr;
    for (int d = tid; d < a.cld; d += kNT) {
      T hi = 0, lo = 0;
      if (d < D) split_hi_lo<T>(rel[d], hi, lo);
      if (mhi) {
        mhi[d] = hi;
        mlo[d] = lo;
      }
      if (d < D) {
        if (n > 0) cen_out[d] = hi;
        cmax = fmax(cmax, fabs((double)(n > 0 ? hi : cenk[d])));
      }
    }
    cmax = block_max(cmax, red);
  }
  if (tid == 0) {
    post[kPN] = n;
    post[kPNu] = nu;
    post[kPBeta] = beta;
    post[kPLogdW] = logdW;
    post[kPCconst] = cconst;
    post[kPFc] = Fc;
    post[kPS] = s_scale;
    post[kPT] = t_scale;
    post[kPRfro] = rfro;
    post[kPVmax] = vmaxk;
    post[kPCmax] = cmax;
    post[kPFail] = fail;
  }
}

// One CTA per group: WeightPost::update and ::fenergy (host_model.cpp) on the device.
template <typename T>
__global__ void __launch_bounds__(128) mstep_weights_kernel(MStepArgs a) {
  extern __shared__ double sm[];
  if (a.stats[a.nstat] != 0.0) return;
  const int K = a.K, tid = threadIdx.x;
  double* Nk = sm;
  double* a1 = Nk + K;
  double* a2 = a1 + K;
  double* Elogv = a2 + K;
  double* Elognv = Elogv + K;
  double* Elogpi = Elognv + K;
  double* term = Elogpi + K;
  int* ord = reinterpret_cast<int*>(term + K);
  __shared__ double sh[4];
  for (int j = blockIdx.x; j < a.J; j += gridDim.x) {
    __syncthreads();
    for (int k = tid; k < K; k += 128) {
      Nk[k] = a.stats[(size_t)j * K + k];
      a1[k] = a.a1p + Nk[k];
      a2[k] = a.a2p;
      ord[k] = k;
    }
    __syncthreads();
    double Fw = 0;
    if (a.wkind == kDirichlet) {
      if (tid == 0) {
        double asum = 0;
        for (int k = 0; k < K; ++k) asum += a1[k];
        sh[0] = asum;
        sh[1] = mm::digamma(asum);
      }
      __syncthreads();
      const double psum = sh[1];
      for (int k = tid; k < K; k += 128) {
        Elogpi[k] = mm::digamma(a1[k]) - psum;
        term[k] = lgamma(a1[k]);
      }
      __syncthreads();
      if (tid == 0) {
        double esum = 0, t = 0;
        for (int k = 0; k < K; ++k) {
          esum += Elogpi[k];
          t += (a1[k] - 1) * Elogpi[k] - term[k];
        }
        Fw = lgamma(sh[0]) - (a.a1p - 1) * esum + t - lgamma(K * a.a1p) + K * lgamma(a.a1p);
      }
    } else {
      if (tid == 0) {
        double total = 0;
        for (int k = 0; k < K; ++k) total += Nk[k];
        mm::DescSorter srt{ord, Nk};
        srt.sort(K);
        double seen = 0;
        for (int r = 0; r < K; ++r) {
          const int k = ord[r];
          seen += Nk[k];
          a2[k] = a.a2p + (total - seen);
        }
      }
      __syncthreads();
      for (int k = tid; k < K; k += 128) {
        const double ps = mm::digamma(a1[k] + a2[k]);
        Elogv[k] = mm::digamma(a1[k]) - ps;
        Elognv[k] = mm::digamma(a2[k]) - ps;
      }
      __syncthreads();
      if (tid == 0) {
        double left = 0;
        for (int r = 0; r < K; ++r) {
          const int k = ord[r];
          Elogpi[k] = Elogv[k] + left;
          left += Elognv[k];
        }
        if (a.wkind == kGDirichlet) {
          const int s = ord[K - 1];
          Elogpi[s] -= Elogv[s];
          Elogv[s] = 0;
          Elognv[s] = 0;
        }
      }
      __syncthreads();
      for (int k = tid; k < K; k += 128)
        term[k] = lgamma(a1[k] + a2[k]) - lgamma(a1[k]) - lgamma(a2[k]) + (a1[k] - a.a1p) * Elogv[k] +
                  (a2[k] - a.a2p) * Elognv[k];
      __syncthreads();
      if (tid == 0) {
        double s = 0;
        if (a.wkind == kStickBreak) {
          for (int k = 0; k < K; ++k) s += term[k];
          Fw = K * a.Fwp + s;
        } else {
          for (int r = 0; r + 1 < K; ++r) s += term[ord[r]];
          Fw = (K - 1) * a.Fwp + s;
        }
      }
    }
    __syncthreads();
    if (a.path == 0) {
      T* lw = reinterpret_cast<T*>(a.lw) + (size_t)j * K;
      for (int k = tid; k < K; k += 128) lw[k] = (T)Elogpi[k];
    } else {
      for (int k = tid; k < K; k += 128) a.lwf[(size_t)j * K + k] = (float)Elogpi[k];
    }
    if (tid == 0) a.wscr[j] = Fw;
  }
}

template <typename T>
__global__ void __launch_bounds__(kNT) mstep_finish_kernel(MStepArgs a) {
  __shared__ double red[16];
  __shared__ double bc[8];
  const int K = a.K, D = a.D, tid = threadIdx.x;
  if (a.stats[a.nstat] != 0.0) {
    if (tid == 0) {
      a.iter[kItAbort] = 1.0;
      a.ctl[kCtlSkipE] = 1u;
    }
    return;
  }
  if (tid == 0) {
    double cbar = 0, Fc = 0, fail = 0, cmax = 0, vmax = 0, Fw = 0;
    for (int k = 0; k < K; ++k) {
      const double* p = a.post + (size_t)k * kPostStride;
      cbar += p[kPCconst];
      Fc += p[kPFc];
      if (p[kPFail] != 0 && fail == 0) fail = p[kPFail];
      cmax = fmax(cmax, p[kPCmax]);
      vmax = p[kPVmax] > vmax || !isfinite(p[kPVmax]) ? p[kPVmax] : vmax;
    }
    cbar /= K;
    for (int j = 0; j < a.J; ++j) Fw += a.wscr[j];
    bc[0] = cbar;
    bc[1] = vmax;
    a.iter[kItFc] = Fc;
    a.iter[kItFw] = Fw;
    a.iter[kItCbar] = cbar;
    a.iter[kItMFail] = fail;
    if (fail != 0) a.ctl[kCtlSkipE] = 1u;
    // operand scale of the tensor-core scatter: scale * max |x - c| <= 2^14 over the resident rows
    if (a.sscale != nullptr) {
      const double span = fmax(a.xabs_max + cmax, 1e-30);
      const int e = min(100, max(-100, mm::floor_log2(16384.0 / span)));
      *a.sscale = (float)ldexp(1.0, e);
    }
  }
  __syncthreads();
  const double cbar = bc[0];
  for (int k = tid; k < K; k += kNT) {
    const double ch = a.post[(size_t)k * kPostStride + kPCconst] - cbar;
    if (a.path == 0) reinterpret_cast<T*>(a.chat)[k] = (T)ch;
    else a.chatf[k] = (float)ch;
  }
  if (a.path != 1 || !a.two_level) return;
  // ---- level-1 operands of the two-level E pass (Engine::ephase_tc) ----
  const double vmax = bc[1];
  int aug_exp = 0;
  bool ok = isfinite(vmax) && a.xabs_max > 0;
  if (ok && vmax > 16384.0) aug_exp = mm::ceil_log2(vmax / 16384.0);
  if (aug_exp > 15) ok = false;
  if (!ok) {
    if (tid == 0) {
      a.iter[kItAugFail] = 1.0;
      a.iter[kItRerun] = 1.0;
      a.ctl[kCtlSkipL] = 1u;
      a.ctl[kCtlAugH] = 0x3c003c00u;
    }
    aug_exp = 0;
  } else if (tid == 0) {
    const uint32_t h = (uint32_t)__half_as_ushort(__float2half_rn(ldexpf(1.f, aug_exp)));
    a.ctl[kCtlAugH] = h | (h << 16);
  }
  const double p2 = ldexp(1.0, aug_exp);
  const double eps = 1.0625 * ldexp(1.0, -10);  // two fp16 roundings + fp32 accumulation of 144 terms
  const int K4 = (K + 3) / 4 * 4;
  for (int k = 0; k < K4; ++k) {
    uint8_t* blk = a.aug + (size_t)(k >> 2) * 16384u;
    const int j = k & 3;
    double worst = 0;
    if (tid < 128) {
      const int i = tid;
      __align__(16) __half hs[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) hs[e] = __float2half_rn(0.f);
      if (k < K && ok) {
        double rem = i < D ? -a.vaug[(size_t)k * D + i] / p2 : 0.0;
        for (int slot = 0; slot < 3; ++slot) {
          hs[slot] = __float2half_rn((float)rem);
          rem -= (double)__half2float(hs[slot]);
        }
        worst = fabs(rem);
      }
      const uint32_t c0 = 2u * (uint32_t)j, c1 = c0 + 1;
      const uint32_t row = (uint32_t)i * 128u;
      *reinterpret_cast<uint4*>(blk + row + ((c0 ^ ((uint32_t)i & 7u)) << 4)) = *reinterpret_cast<const uint4*>(hs);
      *reinterpret_cast<uint4*>(blk + row + ((c1 ^ ((uint32_t)i & 7u)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
    const double res = block_max(worst, red);
    if (tid == 0 && k < K) {
      const double* p = a.post + (size_t)k * kPostStride;
      const double unit = a.sg * (p[kPT] / p[kPS]);
      a.cpar[k] = (float)(1.0 / (unit * unit));
      a.cpar[(size_t)K + k] = (float)(eps * p[kPRfro] * (1.0 + 1e-6));
      a.cpar[2 * (size_t)K + k] =
          (float)((eps * p[kPRfro] * a.xabs_max * ldexp(1.0, -12) + sqrt((double)D) * res * p2 / unit) * (1.0 + 1e-6));
      a.cpar[3 * (size_t)K + k] = (float)(p[kPCconst] - cbar);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kNT)
centres_from_stats_kernel(const double* __restrict__ stats, int J, int K, int D, int cld,
                          const double* __restrict__ centre, T* __restrict__ cen) {
  const int k = blockIdx.x;
  double n = 0;
  for (int j = 0; j < J; ++j) n += stats[(size_t)j * K + k];
  if (!(n > 0)) return;
  const double* xs = stats + (size_t)J * K + (size_t)k * D;
  for (int d = threadIdx.x; d < D; d += kNT) {
    const double c = (double)cen[(size_t)k * cld + d] + centre[d];
    const double mean = (xs[d] + n * c) / n;
    cen[(size_t)k * cld + d] = (T)(mean - centre[d]);
  }
}

__global__ void build_act_kernel(const double* __restrict__ Njk, int64_t n, double cutoff, uint8_t* __restrict__ act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) act[i] = Njk[i] >= cutoff ? 1 : 0;
}

__global__ void list_plan_kernel(const long long* __restrict__ tot, int K, long long cap, double too_many,
                                 long long* __restrict__ koff, int32_t* __restrict__ itoff,
                                 long long* __restrict__ nitems_out, double* __restrict__ iter, int slot_n, int slot_max,
                                 int over_slot, unsigned* __restrict__ ctl, int skip_word, double* abort_slot,
                                 int vote_slot) {
  if (threadIdx.x != 0) return;
  long long nnz = 0, nitems = 0, maxcnt = 0;
  for (int k = 0; k < K; ++k) {
    const long long t = tot[k];
    koff[k] = nnz;
    if (itoff) itoff[k] = (int32_t)(nitems < 2000000000LL ? nitems : 2000000000LL);
    nnz += t;
    nitems += (t + 127) / 128;
    maxcnt = t > maxcnt ? t : maxcnt;
  }
  if (itoff) itoff[K] = (int32_t)(nitems < 2000000000LL ? nitems : 2000000000LL);
  iter[slot_n] = (double)nnz;
  iter[slot_max] = (double)maxcnt;
  if (itoff) iter[kItItems] = (double)nitems;
  const bool over = nnz > cap || (too_many >= 0 && (double)nnz > too_many) || nitems > 2000000000LL;
  if (over) {
    ctl[skip_word] = 1u;
    iter[over_slot] = 1.0;
    if (abort_slot) *abort_slot = 1.0;
    if (vote_slot >= 0) iter[vote_slot] = 1.0;
  }
  if (nitems_out) *nitems_out = over ? 0 : nitems;
}

}  // namespace

size_t mstep_work_doubles(int D) { return (size_t)D * (size_t)(D | 1); }

template <typename T>
cudaError_t mstep(cudaStream_t st, const MStepArgs& a) {
  if (a.K <= 0) return cudaSuccess;
  const bool full = a.ckind == kGaussWish;
  const size_t vec = (size_t)(8 * a.D + 16) * sizeof(double);
  const size_t work = full ? mstep_work_doubles(a.D) * sizeof(double) : 0;
  const int in_smem = full && vec + work <= kSmemWorkLimit;
  if (full && !in_smem && a.work == nullptr) return cudaErrorInvalidValue;
  const size_t smem = vec + (in_smem ? work : 0);
  cudaError_t e = cudaFuncSetAttribute(mstep_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  mstep_cluster_kernel<T><<<a.K, kNT, smem, st>>>(a, in_smem);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t wsm = (size_t)a.K * (7 * sizeof(double) + sizeof(int)) + 16;
  if (wsm > 48 * 1024) {
    e = cudaFuncSetAttribute(mstep_weights_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm);
    if (e != cudaSuccess) return e;
  }
  mstep_weights_kernel<T><<<a.J < 1024 ? a.J : 1024, 128, wsm, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  mstep_finish_kernel<T><<<1, kNT, 0, st>>>(a);
  return cudaGetLastError();
}
template cudaError_t mstep<float>(cudaStream_t, const MStepArgs&);
template cudaError_t mstep<double>(cudaStream_t, const MStepArgs&);

template <typename T>
cudaError_t centres_from_stats(cudaStream_t st, const double* stats, int J, int K, int D, int cld, const double* centre,
                               T* cen) {
  if (K <= 0) return cudaSuccess;
  centres_from_stats_kernel<T><<<K, kNT, 0, st>>>(stats, J, K, D, cld, centre, cen);
  return cudaGetLastError();
}
template cudaError_t centres_from_stats<float>(cudaStream_t, const double*, int, int, int, int, const double*, float*);
template cudaError_t centres_from_stats<double>(cudaStream_t, const double*, int, int, int, int, const double*, double*);

cudaError_t build_act(cudaStream_t st, const double* Njk, int64_t n, double cutoff, uint8_t* act) {
  if (n <= 0) return cudaSuccess;
  build_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Njk, n, cutoff, act);
  return cudaGetLastError();
}

cudaError_t list_plan(cudaStream_t st, const long long* tot, int K, long long cap, double too_many, long long* koff,
                      int32_t* itoff, long long* nitems_out, double* iter, int slot_n, int slot_max, int over_slot,
                      unsigned* ctl, int skip_word, double* abort_slot, int vote_slot) {
  list_plan_kernel<<<1, 32, 0, st>>>(tot, K, cap, too_many, koff, itoff, nitems_out, iter, slot_n, slot_max, over_slot,
                                     ctl, skip_word, abort_slot, vote_slot);
  return cudaGetLastError();
}

void sort_desc_like_std(const double* v, int n, int* ids) {
  for (int i = 0; i < n; ++i) ids[i] = i;
  mm::DescSorter s{ids, v};
  s.sort(n);
}

}  // namespace dev
}  // namespace lcb
