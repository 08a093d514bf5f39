// kernels.cu -- hand-written sm_100a kernels of the VB pass (SIMT tier).
//
// Two passes per VB iteration (see DESIGN.md "kernels"):
//   sstat_*  : sufficient statistics of the stored responsibilities q about
//              per-cluster centres  (the reference's updateSS/addobs,
//              src/cluster.cpp:53-82, src/distributions.cpp:301-313,426-438)
//   estep_*  : expected log-likelihood of every point under every cluster,
//              row soft-max, -sum log Z   (vbexpectation, src/cluster.cpp:91-138;
//              Eloglike src/distributions.cpp:356-370,483-492; mahaldist/logsumexp
//              src/probutils.cpp:113-150)
// Both are register-tiled GEMM-shaped loops on CUDA cores, templated on the
// arithmetic type so the same code runs as the fp32 measured path and as the
// fp64 exact path.  The tcgen05 tier for D in {64,128} lives in tc_kernels.cu.
#include "kernels.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace lcb {
namespace dev {

namespace {

constexpr int kThreads = 256;

template <typename T> struct Vec4 { T v[4]; };

template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }
template <typename T> __device__ __forceinline__ T t_neg_inf();
template <> __device__ __forceinline__ float t_neg_inf<float>() { return -INFINITY; }
template <> __device__ __forceinline__ double t_neg_inf<double>() { return -(double)INFINITY; }

__device__ __forceinline__ double block_sum_double(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  double s = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += scratch[i];
  return s;  // valid on thread 0
}

// ---------------------------------------------------------------------------
// Shared tail of both E kernels.  Ls [TM][KP1] holds the logits of one tile.
//   kEWrite    : q = softmax_k(logit), written to global; Fz += logZ
//   kEScore    : H_k += q_stored * logit   (split ranking, cluster.cpp:401-415)
//   kERawLogit : q[n][k] = logit            (operator-level Eloglike)
// ---------------------------------------------------------------------------
template <typename T, int TM>
__device__ __forceinline__ void tile_tail(T* Ls, int KP1, int K, int64_t n0, int64_t N, T* q, int64_t ldq, int mode,
                                          double& fz_acc, double* h_acc) {
  const int tid = threadIdx.x;
  if (mode == kEWrite) {
    const int n = tid >> 2, sub = tid & 3;
    if (n < TM) {
      T* row = Ls + (size_t)n * KP1;
      T mx = t_neg_inf<T>();
      for (int k = sub; k < K; k += 4) mx = fmax(mx, row[k]);
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      T se = 0;
      for (int k = sub; k < K; k += 4) se += t_exp<T>(row[k] - mx);
      se += __shfl_xor_sync(0xffffffffu, se, 1);
      se += __shfl_xor_sync(0xffffffffu, se, 2);
      const T lz = t_log<T>(se) + mx;
      const bool valid = (n0 + n) < N;
      for (int k = sub; k < K; k += 4) row[k] = valid ? t_exp<T>(row[k] - lz) : (T)0;
      if (sub == 0 && valid) fz_acc += (double)lz;
    } else {
      // keep the shuffles convergent for TM*4 < blockDim
      T d = 0;
      d = __shfl_xor_sync(0xffffffffu, d, 1);
      d = __shfl_xor_sync(0xffffffffu, d, 2);
      d = __shfl_xor_sync(0xffffffffu, d, 1);
      d = __shfl_xor_sync(0xffffffffu, d, 2);
      (void)d;
    }
    __syncthreads();
    for (int idx = tid; idx < TM * K; idx += kThreads) {
      const int n = idx / K, k = idx - n * K;
      if (n0 + n < N) q[(n0 + n) * ldq + k] = Ls[(size_t)n * KP1 + k];
    }
  } else if (mode == kERawLogit) {
    for (int idx = tid; idx < TM * K; idx += kThreads) {
      const int n = idx / K, k = idx - n * K;
      if (n0 + n < N) q[(n0 + n) * ldq + k] = Ls[(size_t)n * KP1 + k];
    }
  } else {
    for (int idx = tid; idx < TM * K; idx += kThreads) {
      const int n = idx / K, k = idx - n * K;
      T v = 0;
      if (n0 + n < N) {
        const T qv = q[(n0 + n) * ldq + k];
        v = (qv == (T)0) ? (T)0 : qv * Ls[(size_t)n * KP1 + k];
      }
      Ls[(size_t)n * KP1 + k] = v;
    }
    __syncthreads();
    for (int k = tid, s = 0; k < K; k += kThreads, ++s) {
      double a = 0;
      for (int n = 0; n < TM; ++n) a += (double)Ls[(size_t)n * KP1 + k];
      h_acc[s] += a;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// E step, full covariance.  One CTA = 256 threads as 16 (tx: whitened
// dimensions, interleaved by 16) x 16 (ty: PT points each).  Per cluster k:
//   XcT = (X_tile - m_k)^T  -> Y = Xc R_k^T by DC-deep chunks of R_k^T staged in
//   shared memory, skipping the column groups that the triangular R_k leaves
//   zero -> logit = chat_k + lw - 0.5 |y|^2.
// ---------------------------------------------------------------------------
template <typename T, int TN, int PT>
__global__ void __launch_bounds__(kThreads)
estep_full_kernel(const T* __restrict__ X, int64_t N, int D, int64_t ldx, const int32_t* __restrict__ gid, int K,
                  const T* __restrict__ RT, const T* __restrict__ mhi, const T* __restrict__ mlo,
                  const T* __restrict__ chat, const T* __restrict__ lw, const uint8_t* __restrict__ act,
                  T* __restrict__ q, int64_t ldq, int mode, double* __restrict__ Fz, double* __restrict__ H,
                  const unsigned* __restrict__ skip) {
  constexpr int DP = 16 * TN, TM = 16 * PT, XS = TM + 4, DC = 8;
  if (skip != nullptr && *skip != 0u) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* XT = reinterpret_cast<T*>(smem_raw);  // [DP][XS] tile, dimension-major
  T* XcT = XT + DP * XS;                   // [DP][XS] centred on the current cluster
  T* Rs = XcT + DP * XS;                   // [DC][DP]
  const int KP1 = K | 1;
  T* Ls = Rs + DC * DP;                    // [TM][KP1]
  int* gs = reinterpret_cast<int*>(Ls + (size_t)TM * KP1);
  __shared__ double red[kThreads / 32];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t ntiles = (N + TM - 1) / TM;
  double fz_acc = 0;
  double h_acc[2] = {0, 0};

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t n0 = tile * TM;
    for (int idx = tid; idx < TM * DP; idx += kThreads) {
      const int n = idx / DP, d = idx - n * DP;
      T v = 0;
      if (n0 + n < N && d < D) v = X[(n0 + n) * ldx + d];
      XT[d * XS + n] = v;
    }
    if (tid < TM) gs[tid] = (gid != nullptr && n0 + tid < N) ? gid[n0 + tid] : 0;
    __syncthreads();

    for (int k = 0; k < K; ++k) {
      const T* mh = mhi + (size_t)k * DP;
      const T* ml = mlo + (size_t)k * DP;
      for (int idx = tid; idx < TM * DP; idx += kThreads) {
        const int d = idx / TM, n = idx - d * TM;
        XcT[d * XS + n] = (XT[d * XS + n] - mh[d]) - ml[d];
      }
      T acc[PT][TN];
#pragma unroll
      for (int p = 0; p < PT; ++p)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[p][c] = 0;

      const T* Rk = RT + (size_t)k * DP * DP;
      for (int d0 = 0; d0 < DP; d0 += DC) {
        __syncthreads();
        for (int idx = tid; idx < DC * DP; idx += kThreads) Rs[idx] = Rk[(size_t)d0 * DP + idx];
        __syncthreads();
        T a[DC][PT];
#pragma unroll
        for (int dd = 0; dd < DC; ++dd)
#pragma unroll
          for (int p = 0; p < PT; ++p) a[dd][p] = XcT[(d0 + dd) * XS + ty * PT + p];
        const int c0 = d0 >> 4;  // column groups below the chunk are structurally zero
#pragma unroll
        for (int c = 0; c < TN; ++c) {
          if (c >= c0) {
#pragma unroll
            for (int dd = 0; dd < DC; ++dd) {
              const T b = Rs[dd * DP + tx + 16 * c];
#pragma unroll
              for (int p = 0; p < PT; ++p) acc[p][c] = fma(a[dd][p], b, acc[p][c]);
            }
          }
        }
      }
      const T ck = chat[k];
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        T s = 0;
#pragma unroll
        for (int c = 0; c < TN; ++c) s = fma(acc[p][c], acc[p][c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        if (tx == 0) {
          const int n = ty * PT + p;
          const int g = gs[n];
          T l = ck + lw[(size_t)g * K + k] - (T)0.5 * s;
          if (act != nullptr && !act[(size_t)g * K + k]) l = t_neg_inf<T>();
          Ls[(size_t)n * KP1 + k] = l;
        }
      }
      __syncthreads();
    }
    tile_tail<T, TM>(Ls, KP1, K, n0, N, q, ldq, mode, fz_acc, h_acc);
  }

  if (mode == kEWrite) {
    const double s = block_sum_double(fz_acc, red);
    if (tid == 0) atomicAdd(Fz, s);
  } else if (mode == kEScore) {
    for (int k = tid, s = 0; k < K; k += kThreads, ++s) atomicAdd(H + k, h_acc[s]);
  }
}

// ---------------------------------------------------------------------------
// E step, diagonal covariance: logit = chat + lw - 0.5 sum_d A_kd (x_d - m_kd)^2
// 16 (tx: 4 clusters) x 16 (ty: 4 points) threads, D swept in chunks of 32.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
estep_diag_kernel(const T* __restrict__ X, int64_t N, int D, int64_t ldx, const int32_t* __restrict__ gid, int K,
                  const T* __restrict__ A, const T* __restrict__ mhi, const T* __restrict__ mlo,
                  const T* __restrict__ chat, const T* __restrict__ lw, const uint8_t* __restrict__ act,
                  T* __restrict__ q, int64_t ldq, int mode, double* __restrict__ Fz, double* __restrict__ H,
                  const unsigned* __restrict__ skip) {
  constexpr int TM = 64, KT = 64, DC = 32, XS = TM + 4, KS = KT + 4;
  if (skip != nullptr && *skip != 0u) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Xch = reinterpret_cast<T*>(smem_raw);  // [DC][XS]
  T* Ach = Xch + DC * XS;                   // [DC][KS]
  T* Mh = Ach + DC * KS;
  T* Ml = Mh + DC * KS;
  const int KP1 = K | 1;
  T* Ls = Ml + DC * KS;  // [TM][KP1]
  int* gs = reinterpret_cast<int*>(Ls + (size_t)TM * KP1);
  __shared__ double red[kThreads / 32];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t ntiles = (N + TM - 1) / TM;
  double fz_acc = 0;
  double h_acc[2] = {0, 0};

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t n0 = tile * TM;
    if (tid < TM) gs[tid] = (gid != nullptr && n0 + tid < N) ? gid[n0 + tid] : 0;
    for (int kt0 = 0; kt0 < K; kt0 += KT) {
      T acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[p][c] = 0;
      for (int d0 = 0; d0 < D; d0 += DC) {
        __syncthreads();
        for (int idx = tid; idx < TM * DC; idx += kThreads) {
          const int n = idx / DC, dd = idx - n * DC;
          T v = 0;
          if (n0 + n < N && d0 + dd < D) v = X[(n0 + n) * ldx + d0 + dd];
          Xch[dd * XS + n] = v;
        }
        for (int idx = tid; idx < KT * DC; idx += kThreads) {
          const int kk = idx / DC, dd = idx - kk * DC;
          T a = 0, h = 0, l = 0;
          if (kt0 + kk < K && d0 + dd < D) {
            const size_t o = (size_t)(kt0 + kk) * D + d0 + dd;
            a = A[o];
            h = mhi[o];
            l = mlo[o];
          }
          Ach[dd * KS + kk] = a;
          Mh[dd * KS + kk] = h;
          Ml[dd * KS + kk] = l;
        }
        __syncthreads();
        if constexpr (sizeof(T) == 4) {
          // fp32: packed f32x2 instructions, two clusters per lane-operation (the loop is bound by the issue rate of
          // its four floating-point instructions per (point, cluster, dimension); the same IEEE operations as below)
#pragma unroll 4
          for (int dd = 0; dd < DC; ++dd) {
            const float4 xv = *reinterpret_cast<const float4*>(Xch + dd * XS + ty * 4);
            const float4 av = *reinterpret_cast<const float4*>(Ach + dd * KS + tx * 4);
            const float4 hv = *reinterpret_cast<const float4*>(Mh + dd * KS + tx * 4);
            const float4 lv = *reinterpret_cast<const float4*>(Ml + dd * KS + tx * 4);
            const float xs4[4] = {xv.x, xv.y, xv.z, xv.w};
            unsigned long long a2[2], h2[2], l2[2];
            asm("mov.b64 %0, {%1, %2};" : "=l"(a2[0]) : "f"(av.x), "f"(av.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(a2[1]) : "f"(av.z), "f"(av.w));
            asm("mov.b64 %0, {%1, %2};" : "=l"(h2[0]) : "f"(hv.x), "f"(hv.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(h2[1]) : "f"(hv.z), "f"(hv.w));
            asm("mov.b64 %0, {%1, %2};" : "=l"(l2[0]) : "f"(lv.x), "f"(lv.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(l2[1]) : "f"(lv.z), "f"(lv.w));
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              unsigned long long x2;
              asm("mov.b64 %0, {%1, %1};" : "=l"(x2) : "f"(xs4[p]));
#pragma unroll
              for (int cp = 0; cp < 2; ++cp) {
                unsigned long long t2, u2, acc2;
                asm("mov.b64 %0, {%1, %2};" : "=l"(acc2) : "f"((float)acc[p][2 * cp]), "f"((float)acc[p][2 * cp + 1]));
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t2) : "l"(x2), "l"(h2[cp]));
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t2) : "l"(t2), "l"(l2[cp]));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u2) : "l"(a2[cp]), "l"(t2));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2) : "l"(u2), "l"(t2));
                float r0, r1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(acc2));
                acc[p][2 * cp] = (T)r0;
                acc[p][2 * cp + 1] = (T)r1;
              }
            }
          }
        } else {
#pragma unroll 4
        for (int dd = 0; dd < DC; ++dd) {
          T x[4], a[4], h[4], l[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) x[p] = Xch[dd * XS + ty * 4 + p];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            a[c] = Ach[dd * KS + tx * 4 + c];
            h[c] = Mh[dd * KS + tx * 4 + c];
            l[c] = Ml[dd * KS + tx * 4 + c];
          }
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const T t = (x[p] - h[c]) - l[c];
              acc[p][c] = fma(a[c] * t, t, acc[p][c]);
            }
        }
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int n = ty * 4 + p, k = kt0 + tx * 4 + c;
          if (k < K) {
            const int g = gs[n];
            T lg = chat[k] + lw[(size_t)g * K + k] - (T)0.5 * acc[p][c];
            if (act != nullptr && !act[(size_t)g * K + k]) lg = t_neg_inf<T>();
            Ls[(size_t)n * KP1 + k] = lg;
          }
        }
    }
    __syncthreads();
    tile_tail<T, TM>(Ls, KP1, K, n0, N, q, ldq, mode, fz_acc, h_acc);
  }
  if (mode == kEWrite) {
    const double s = block_sum_double(fz_acc, red);
    if (tid == 0) atomicAdd(Fz, s);
  } else if (mode == kEScore) {
    for (int k = tid, s = 0; k < K; k += kThreads, ++s) atomicAdd(H + k, h_acc[s]);
  }
}

// ---------------------------------------------------------------------------
// Sufficient statistics, full covariance.  CTA (k, row chunk, block (bi,bj)):
//   S_k[bi,bj] += sum_n q_nk (x_n - c_k)_bi (x_n - c_k)_bj^T   thread tile TN x TN
//   xs_k      += sum_n q_nk (x_n - c_k)                         (bj == 0 CTAs)
// fp32 partial sums cover at most rows_per_cta rows before they are added into
// the fp64 global accumulators.
// ---------------------------------------------------------------------------
template <typename T, int TN>
__global__ void __launch_bounds__(kThreads)
sstat_full_kernel(const T* __restrict__ X, int64_t N, int D, int64_t ldx, const int32_t* __restrict__ gid,
                  const T* __restrict__ q, int64_t ldq, int K, int DPc, const T* __restrict__ cen,
                  const uint8_t* __restrict__ act, int rows_per_cta, int nb, double* __restrict__ xs,
                  double* __restrict__ S) {
  constexpr int BW = 16 * TN, TMS = sizeof(T) == 8 ? 16 : 32;
  __shared__ __align__(16) T XI[TMS * BW];
  __shared__ __align__(16) T XJ[TMS * BW];
  __shared__ T qs[TMS];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k = blockIdx.x;
  const int bi = blockIdx.z / nb, bj = blockIdx.z - bi * nb;
  const int i0 = bi * BW, j0 = bj * BW;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < N) ? r0 + rows_per_cta : N;
  const T* ck = cen + (size_t)k * DPc;

  T acc[TN][TN];
#pragma unroll
  for (int a = 0; a < TN; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0;
  T xacc = 0;

  for (int64_t t0 = r0; t0 < r1; t0 += TMS) {
    T qv = 0;
    if (tid < TMS) {
      const int64_t n = t0 + tid;
      if (n < r1) {
        qv = q[n * ldq + k];
        if (act != nullptr) {
          const int g = gid != nullptr ? gid[n] : 0;
          if (!act[(size_t)g * K + k]) qv = 0;
        }
      }
      qs[tid] = qv;
    }
    const int any = __syncthreads_or(qv != (T)0);
    if (!any) continue;
    for (int idx = tid; idx < TMS * BW; idx += kThreads) {
      const int n = idx / BW, d = idx - n * BW;
      T vi = 0, vj = 0;
      if (t0 + n < r1) {
        if (i0 + d < D) vi = X[(t0 + n) * ldx + i0 + d] - ck[i0 + d];
        if (j0 + d < D) vj = X[(t0 + n) * ldx + j0 + d] - ck[j0 + d];
      }
      XI[idx] = vi;
      XJ[idx] = vj;
    }
    __syncthreads();
    for (int n = 0; n < TMS; ++n) {
      const T qn = qs[n];
      if (qn == (T)0) continue;
      T a[TN], b[TN];
#pragma unroll
      for (int c = 0; c < TN; ++c) {
        a[c] = XI[n * BW + ty + 16 * c];
        b[c] = qn * XJ[n * BW + tx + 16 * c];
      }
#pragma unroll
      for (int ci = 0; ci < TN; ++ci)
#pragma unroll
        for (int cj = 0; cj < TN; ++cj) acc[ci][cj] = fma(a[ci], b[cj], acc[ci][cj]);
      if (bj == 0 && tid < BW) xacc = fma(qn, XI[n * BW + tid], xacc);
    }
    __syncthreads();
  }
#pragma unroll
  for (int ci = 0; ci < TN; ++ci)
#pragma unroll
    for (int cj = 0; cj < TN; ++cj) {
      const int i = i0 + ty + 16 * ci, j = j0 + tx + 16 * cj;
      if (i < D && j < D && acc[ci][cj] != (T)0) atomicAdd(&S[((size_t)k * D + i) * D + j], (double)acc[ci][cj]);
    }
  if (bj == 0 && tid < BW && i0 + tid < D && xacc != (T)0) atomicAdd(&xs[(size_t)k * D + i0 + tid], (double)xacc);
}


// ---------------------------------------------------------------------------
// Sufficient statistics over the NON-ZERO responsibilities only.
//   nz_count : per (row block, cluster) number of rows with q != 0 (after the
//              sparse-update mask)              -> blockcnt [nblocks][K]
//   nz_scan  : per cluster exclusive scan over the row blocks (in place) and the
//              per-cluster totals                -> total [K]
//   nz_fill  : row index and q of every non-zero into per-cluster lists
//   sstat_gather_full : per (cluster, 4096-row chunk of its list) the centred
//              scatter on a register tile; fp32 partial sums are folded into an
//              fp64 accumulator in shared memory every 64 rows so that the
//              accumulation error stays ~1e-8 relative whatever N_k is.
// ---------------------------------------------------------------------------
constexpr int kNzBlock = 2048;  // rows per counting block

// Column lanes: thread (kl, rl) walks rows rl, rl+rw, ... of the block for columns kl, kl+kw, ... (no per-element
// division).  When Njk != nullptr the same sweep also produces the per-group column sums (updateSS's Njk).
template <typename T>
__global__ void __launch_bounds__(kThreads)
nz_count_kernel(const T* __restrict__ q, int64_t ldq, int64_t N, int K, const int32_t* __restrict__ gid,
                const uint8_t* __restrict__ act, int kw, int32_t* __restrict__ blockcnt, double* __restrict__ Njk,
                int pred) {
  extern __shared__ int scnt[];
  for (int k = threadIdx.x; k < K; k += kThreads) scnt[k] = 0;
  __syncthreads();
  const int kl = threadIdx.x % kw, rl = threadIdx.x / kw, rw = kThreads / kw;
  const int64_t r0 = (int64_t)blockIdx.x * kNzBlock;
  const int64_t r1 = (r0 + kNzBlock < N) ? r0 + kNzBlock : N;
  for (int k = kl; k < K; k += kw) {
    int c = 0;
    double acc = 0;
    int gc = -1;
    for (int64_t n = r0 + rl; n < r1; n += rw) {
      T v = q[n * ldq + k];
      const int g = gid != nullptr ? gid[n] : 0;
      if (Njk != nullptr) {
        if (g != gc) {
          if (gc >= 0 && acc != 0) atomicAdd(&Njk[(size_t)gc * K + k], acc);
          gc = g;
          acc = 0;
        }
        acc += (double)v;
      }
      if (pred == kNzNotNegInf) {
        if (v != -(T)INFINITY) ++c;
        continue;
      }
      if (act != nullptr && !act[(size_t)g * K + k]) v = 0;
      if (v != (T)0) ++c;
    }
    if (Njk != nullptr && gc >= 0 && acc != 0) atomicAdd(&Njk[(size_t)gc * K + k], acc);
    if (c) atomicAdd(&scnt[k], c);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += kThreads) blockcnt[(size_t)blockIdx.x * K + k] = scnt[k];
}

// one CTA per cluster: exclusive scan of its column of blockcnt
__global__ void __launch_bounds__(kThreads)
nz_scan_kernel(int32_t* __restrict__ blockcnt, int64_t nblocks, int K, long long* __restrict__ total) {
  __shared__ long long part[kThreads];
  const int k = blockIdx.x, t = threadIdx.x;
  const int64_t per = (nblocks + kThreads - 1) / kThreads;
  const int64_t b0 = (int64_t)t * per, b1 = (b0 + per < nblocks) ? b0 + per : nblocks;
  long long s = 0;
  for (int64_t b = b0; b < b1; ++b) s += blockcnt[(size_t)b * K + k];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int i = 0; i < kThreads; ++i) {
      const long long v = part[i];
      part[i] = run;
      run += v;
    }
    total[k] = run;
  }
  __syncthreads();
  long long run = part[t];
  for (int64_t b = b0; b < b1; ++b) {
    const int v = blockcnt[(size_t)b * K + k];
    blockcnt[(size_t)b * K + k] = (int32_t)run;
    run += v;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
nz_fill_kernel(const T* __restrict__ q, int64_t ldq, int64_t N, int K, const int32_t* __restrict__ gid,
               const uint8_t* __restrict__ act, int kw, const int32_t* __restrict__ blockoff,
               const long long* __restrict__ koff, int32_t* __restrict__ lrow, T* __restrict__ lq, int pred,
               const unsigned* __restrict__ skip) {
  extern __shared__ int scnt[];
  if (skip != nullptr && *skip != 0u) return;
  for (int k = threadIdx.x; k < K; k += kThreads) scnt[k] = 0;
  __syncthreads();
  const int kl = threadIdx.x % kw, rl = threadIdx.x / kw, rw = kThreads / kw;
  const int64_t r0 = (int64_t)blockIdx.x * kNzBlock;
  const int64_t r1 = (r0 + kNzBlock < N) ? r0 + kNzBlock : N;
  for (int k = kl; k < K; k += kw) {
    const long long base = koff[k] + blockoff[(size_t)blockIdx.x * K + k];
    for (int64_t n = r0 + rl; n < r1; n += rw) {
      T v = q[n * ldq + k];
      if (pred == kNzNotNegInf) {
        if (v != -(T)INFINITY) {
          const int pos = atomicAdd(&scnt[k], 1);
          lrow[base + pos] = (int32_t)n;
        }
        continue;
      }
      if (act != nullptr && v != (T)0) {
        const int g = gid != nullptr ? gid[n] : 0;
        if (!act[(size_t)g * K + k]) v = 0;
      }
      if (v != (T)0) {
        const int pos = atomicAdd(&scnt[k], 1);
        lrow[base + pos] = (int32_t)n;
        lq[base + pos] = v;
      }
    }
  }
}

constexpr int kGatherChunk = 4096;  // list rows per CTA

template <typename T, int TN>
__global__ void __launch_bounds__(kThreads)
sstat_gather_full_kernel(const T* __restrict__ X, int D, int64_t ldx, const int32_t* __restrict__ lrow,
                         const T* __restrict__ lq, const long long* __restrict__ koff,
                         const long long* __restrict__ kcnt, int DPc, const T* __restrict__ cen, int nb,
                         double* __restrict__ xs, double* __restrict__ S, const unsigned* __restrict__ skip) {
  if (skip != nullptr && *skip != 0u) return;
  constexpr int BW = 16 * TN, TMS = sizeof(T) == 8 ? 16 : 32;
  constexpr bool kStage = sizeof(T) == 4;  // fp32: fold into fp64 shared accumulators every 2 sub-tiles
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* XI = reinterpret_cast<T*>(smem_raw);
  T* XJ = XI + TMS * BW;
  T* qs = XJ + TMS * BW;
  int* rs = reinterpret_cast<int*>(qs + TMS);
  double* Sacc = reinterpret_cast<double*>(smem_raw + (((size_t)(2 * TMS * BW + TMS) * sizeof(T) + TMS * 4 + 15) & ~(size_t)15));
  double* xsacc = Sacc + (kStage ? TN * TN * kThreads : 0);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k = blockIdx.y;
  const int bi = blockIdx.z / nb, bj = blockIdx.z - bi * nb;
  const int i0 = bi * BW, j0 = bj * BW;
  const long long cnt = kcnt[k];
  const long long base = koff[k];
  const T* ck = cen + (size_t)k * DPc;
  // the grid's x extent is sized from a bound on the longest list; chunks beyond it are walked by striding
  for (long long l0 = (long long)blockIdx.x * kGatherChunk; l0 < cnt; l0 += (long long)gridDim.x * kGatherChunk) {
  const long long l1 = (l0 + kGatherChunk < cnt) ? l0 + kGatherChunk : cnt;

  T acc[TN][TN];
#pragma unroll
  for (int a = 0; a < TN; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0;
  T xacc = 0;
  if (kStage) {
#pragma unroll
    for (int e = 0; e < TN * TN; ++e) Sacc[e * kThreads + tid] = 0.0;
    if (tid < BW) xsacc[tid] = 0.0;
  }
  int pending = 0;
  for (long long t0 = l0; t0 < l1; t0 += TMS) {
    __syncthreads();
    if (tid < TMS) {
      const long long l = t0 + tid;
      qs[tid] = (l < l1) ? lq[base + l] : (T)0;
      rs[tid] = (l < l1) ? lrow[base + l] : -1;
    }
    __syncthreads();
    for (int idx = tid; idx < TMS * BW; idx += kThreads) {
      const int n = idx / BW, d = idx - n * BW;
      const int r = rs[n];
      T vi = 0, vj = 0;
      if (r >= 0) {
        if (i0 + d < D) vi = X[(int64_t)r * ldx + i0 + d] - ck[i0 + d];
        if (j0 + d < D) vj = X[(int64_t)r * ldx + j0 + d] - ck[j0 + d];
      }
      XI[idx] = vi;
      XJ[idx] = vj;
    }
    __syncthreads();
#pragma unroll 4
    for (int n = 0; n < TMS; ++n) {
      const T qn = qs[n];
      T a[TN], b[TN];
#pragma unroll
      for (int c = 0; c < TN; ++c) {
        a[c] = XI[n * BW + ty + 16 * c];
        b[c] = qn * XJ[n * BW + tx + 16 * c];
      }
#pragma unroll
      for (int ci = 0; ci < TN; ++ci)
#pragma unroll
        for (int cj = 0; cj < TN; ++cj) acc[ci][cj] = fma(a[ci], b[cj], acc[ci][cj]);
      if (bj == 0 && tid < BW) xacc = fma(qn, XI[n * BW + tid], xacc);
    }
    if (kStage && ++pending == 2) {
      pending = 0;
#pragma unroll
      for (int ci = 0; ci < TN; ++ci)
#pragma unroll
        for (int cj = 0; cj < TN; ++cj) {
          Sacc[(ci * TN + cj) * kThreads + tid] += (double)acc[ci][cj];
          acc[ci][cj] = 0;
        }
      if (bj == 0 && tid < BW) {
        xsacc[tid] += (double)xacc;
        xacc = 0;
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < TN; ++ci)
#pragma unroll
    for (int cj = 0; cj < TN; ++cj) {
      const int i = i0 + ty + 16 * ci, j = j0 + tx + 16 * cj;
      double v = (double)acc[ci][cj];
      if (kStage) v += Sacc[(ci * TN + cj) * kThreads + tid];
      if (i < D && j < D && v != 0.0) atomicAdd(&S[((size_t)k * D + i) * D + j], v);
    }
  if (bj == 0 && tid < BW && i0 + tid < D) {
    double v = (double)xacc;
    if (kStage) v += xsacc[tid];
    if (v != 0.0) atomicAdd(&xs[(size_t)k * D + i0 + tid], v);
  }
  __syncthreads();
  }
}

// Sufficient statistics, diagonal, over the per-cluster lists of non-zero responsibilities (the lists of the
// full-covariance pass: nz_count / nz_scan / nz_fill):  xs_k += sum_e q_e (x_e - c_k),  S_k += sum_e q_e (x_e - c_k)^2.
// q is numerically sparse, so the pass costs one gathered read of every listed row (2 D flop per pair) instead of
// the K * D products per row of the dense kernel below -- at one pair per row it is bound by the read of X.
// A CTA takes chunks of one cluster's list; thread t owns dimensions t, t + kThreads, ... (D <= 4 * kThreads); fp32
// sums over 64 entries are folded into fp64 registers, one fp64 atomic per (cluster, dimension, CTA) at the end.
template <typename T>
__global__ void __launch_bounds__(kThreads)
sstat_gather_diag_kernel(const T* __restrict__ X, int D, int64_t ldx, const int32_t* __restrict__ lrow,
                         const T* __restrict__ lq, const long long* __restrict__ koff,
                         const long long* __restrict__ kcnt, const T* __restrict__ cen, double* __restrict__ xs,
                         double* __restrict__ S, const unsigned* __restrict__ skip) {
  if (skip != nullptr && *skip != 0u) return;
  constexpr int kChunk = 1024, kFold = 64, kDims = 4;
  const int k = blockIdx.y, tid = threadIdx.x;
  const long long cnt = kcnt[k], base = koff[k];
  T c[kDims];
  double A1[kDims], A2[kDims];
#pragma unroll
  for (int j = 0; j < kDims; ++j) {
    const int d = tid + j * kThreads;
    c[j] = d < D ? cen[(size_t)k * D + d] : (T)0;
    A1[j] = 0;
    A2[j] = 0;
  }
  bool any = false;
  for (long long l0 = (long long)blockIdx.x * kChunk; l0 < cnt; l0 += (long long)gridDim.x * kChunk) {
    const long long l1 = l0 + kChunk < cnt ? l0 + kChunk : cnt;
    any = true;
    for (long long f0 = l0; f0 < l1; f0 += kFold) {
      const int nf = (int)(f0 + kFold < l1 ? kFold : l1 - f0);
      T s1[kDims], s2[kDims];
#pragma unroll
      for (int j = 0; j < kDims; ++j) {
        s1[j] = 0;
        s2[j] = 0;
      }
      int e = 0;
      for (; e + 4 <= nf; e += 4) {  // four rows in flight per thread and dimension
        int32_t r[4];
        T w[4], xv[4][kDims];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          r[u] = __ldg(lrow + base + f0 + e + u);
          w[u] = __ldg(lq + base + f0 + e + u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < kDims; ++j) {
            const int d = tid + j * kThreads;
            xv[u][j] = d < D ? __ldg(X + (size_t)r[u] * ldx + d) : (T)0;
          }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < kDims; ++j) {
            const T xc = xv[u][j] - c[j];
            const T t = w[u] * xc;
            s1[j] += t;
            s2[j] = fma(t, xc, s2[j]);
          }
      }
      for (; e < nf; ++e) {
        const int32_t r = __ldg(lrow + base + f0 + e);
        const T w = __ldg(lq + base + f0 + e);
#pragma unroll
        for (int j = 0; j < kDims; ++j) {
          const int d = tid + j * kThreads;
          const T xc = (d < D ? __ldg(X + (size_t)r * ldx + d) : (T)0) - c[j];
          const T t = w * xc;
          s1[j] += t;
          s2[j] = fma(t, xc, s2[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < kDims; ++j) {
        A1[j] += (double)s1[j];
        A2[j] += (double)s2[j];
      }
    }
  }
  if (!any) return;
#pragma unroll
  for (int j = 0; j < kDims; ++j) {
    const int d = tid + j * kThreads;
    if (d < D) {
      if (A1[j] != 0.0) atomicAdd(&xs[(size_t)k * D + d], A1[j]);
      if (A2[j] != 0.0) atomicAdd(&S[(size_t)k * D + d], A2[j]);
    }
  }
}

// Sufficient statistics, diagonal: CTA covers 64 clusters x 64 dimensions.
template <typename T>
__global__ void __launch_bounds__(kThreads)
sstat_diag_kernel(const T* __restrict__ X, int64_t N, int D, int64_t ldx, const int32_t* __restrict__ gid,
                  const T* __restrict__ q, int64_t ldq, int K, const T* __restrict__ cen,
                  const uint8_t* __restrict__ act, int rows_per_cta, int ndt, double* __restrict__ xs,
                  double* __restrict__ S) {
  constexpr int KT = 64, DT = 64, NC = 32;
  __shared__ __align__(16) T Qs[NC * KT];
  __shared__ __align__(16) T Xs[NC * DT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int kt0 = (blockIdx.x / ndt) * KT, dt0 = (blockIdx.x % ndt) * DT;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < N) ? r0 + rows_per_cta : N;
  T c[4][4], s1[4][4], s2[4][4];
  double S1[4][4], S2[4][4];  // fp32 chunk sums are folded into fp64 every NC rows
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = kt0 + ty * 4 + a, d = dt0 + tx * 4 + b;
      c[a][b] = (k < K && d < D) ? cen[(size_t)k * D + d] : (T)0;
      s1[a][b] = 0;
      s2[a][b] = 0;
      S1[a][b] = 0;
      S2[a][b] = 0;
    }
  for (int64_t t0 = r0; t0 < r1; t0 += NC) {
    __syncthreads();
    for (int idx = tid; idx < NC * KT; idx += kThreads) {
      const int n = idx / KT, kk = idx - n * KT;
      T v = 0;
      if (t0 + n < r1 && kt0 + kk < K) {
        v = q[(t0 + n) * ldq + kt0 + kk];
        if (act != nullptr) {
          const int g = gid != nullptr ? gid[t0 + n] : 0;
          if (!act[(size_t)g * K + kt0 + kk]) v = 0;
        }
      }
      Qs[idx] = v;
    }
    for (int idx = tid; idx < NC * DT; idx += kThreads) {
      const int n = idx / DT, dd = idx - n * DT;
      T v = 0;
      if (t0 + n < r1 && dt0 + dd < D) v = X[(t0 + n) * ldx + dt0 + dd];
      Xs[idx] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int n = 0; n < NC; ++n) {
      T qv[4], xv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qv[a] = Qs[n * KT + ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) xv[b] = Xs[n * DT + tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const T xc = xv[b] - c[a][b];
          const T t = qv[a] * xc;
          s1[a][b] += t;
          s2[a][b] = fma(t, xc, s2[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        S1[a][b] += (double)s1[a][b];
        S2[a][b] += (double)s2[a][b];
        s1[a][b] = 0;
        s2[a][b] = 0;
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = kt0 + ty * 4 + a, d = dt0 + tx * 4 + b;
      if (k < K && d < D) {
        if (S1[a][b] != 0.0) atomicAdd(&xs[(size_t)k * D + d], S1[a][b]);
        if (S2[a][b] != 0.0) atomicAdd(&S[(size_t)k * D + d], S2[a][b]);
      }
    }
}

// Njk[g][k] = sum over rows of group g of q[n][k]
template <typename T>
__global__ void __launch_bounds__(kThreads)
colsum_kernel(const T* __restrict__ q, int64_t ldq, int64_t N, int K, const int32_t* __restrict__ gid, int kw,
              int rows_per_cta, double* __restrict__ Njk) {
  const int kl = threadIdx.x % kw, rl = threadIdx.x / kw, rw = kThreads / kw;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < N) ? r0 + rows_per_cta : N;
  for (int k = kl; k < K; k += kw) {
    double acc = 0;
    int gc = -1;
    for (int64_t n = r0 + rl; n < r1; n += rw) {
      const int g = gid != nullptr ? gid[n] : 0;
      if (g != gc) {
        if (gc >= 0 && acc != 0) atomicAdd(&Njk[(size_t)gc * K + k], acc);
        gc = g;
        acc = 0;
      }
      acc += (double)q[n * ldq + k];
    }
    if (gc >= 0 && acc != 0) atomicAdd(&Njk[(size_t)gc * K + k], acc);
  }
}

// ------------------------------------------------------------- utilities ---
template <typename T>
__global__ void convert_rows_kernel(const double* __restrict__ src, int64_t rows, int D, int64_t ld, int colmajor,
                                    const double* __restrict__ mean, T* __restrict__ dst, int64_t ldx) {
  const int64_t total = rows * ldx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / ldx;
    const int d = (int)(i - n * ldx);
    T v = 0;
    if (d < D) v = (T)((colmajor ? src[(int64_t)d * ld + n] : src[n * ld + d]) - mean[d]);
    dst[i] = v;
  }
}

template <typename T>
__global__ void convert_f32_kernel(const float* __restrict__ src, int64_t rows, int D, int64_t ld,
                                   const double* __restrict__ mean, T* __restrict__ dst, int64_t ldx) {
  const int64_t total = rows * ldx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / ldx;
    const int d = (int)(i - n * ldx);
    T v = 0;
    if (d < D) v = (T)((double)src[n * ld + d] - mean[d]);
    dst[i] = v;
  }
}

__global__ void colsum_f32_kernel(const float* __restrict__ src, int64_t rows, int D, int64_t ld, int rows_per_cta,
                                  double* __restrict__ sums) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < rows) ? r0 + rows_per_cta : rows;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double a = 0;
    for (int64_t n = r0; n < r1; ++n) a += (double)src[n * ld + d];
    atomicAdd(&sums[d], a);
  }
}

// max |x| over a buffer (positive floats order like their bit patterns)
template <typename T> __global__ void absmax_kernel(const T* __restrict__ x, int64_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf((float)x[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

template <typename T> __global__ void fill_ones_kernel(T* q, int64_t ldq, int64_t N) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x)
    q[n * ldq] = (T)1;
}

template <typename T>
__global__ void labels_to_q_kernel(const int32_t* __restrict__ lab, T* __restrict__ q, int64_t ldq, int64_t N, int K) {
  const int64_t total = N * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / K;
    const int k = (int)(i - n * K);
    q[n * ldq + k] = (lab[n] == k) ? (T)1 : (T)0;
  }
}

template <typename T>
__global__ void q_from_double_kernel(const double* __restrict__ src, int64_t rows, int K, T* __restrict__ q, int64_t ldq) {
  const int64_t total = rows * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / K;
    const int k = (int)(i - n * K);
    q[n * ldq + k] = (T)src[i];
  }
}

template <typename T>
__global__ void q_to_double_kernel(const T* __restrict__ q, int64_t ldq, int64_t rows, int K, double* __restrict__ dst,
                                   int64_t ld, int colmajor) {
  const int64_t total = rows * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / K;
    const int k = (int)(i - n * K);
    const double v = (double)q[n * ldq + k];
    if (colmajor) dst[(int64_t)k * ld + n] = v;
    else dst[n * ld + k] = v;
  }
}

constexpr int kMemberBlock = 1024;  // rows per compaction block

template <typename T>
__global__ void __launch_bounds__(kThreads)
member_counts_kernel(const T* __restrict__ q, int64_t ldq, int64_t N, int k, int32_t* __restrict__ blockcnt) {
  const int64_t r0 = (int64_t)blockIdx.x * kMemberBlock;
  int c = 0;
  for (int i = threadIdx.x; i < kMemberBlock; i += kThreads) {
    const int64_t n = r0 + i;
    if (n < N && q[n * ldq + k] > (T)0.5) ++c;
  }
  __shared__ int red[kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    blockcnt[blockIdx.x] = s;
  }
}

// exclusive scan of the per-block counts by a single CTA (nblocks is N/1024)
__global__ void __launch_bounds__(1024) scan_counts_kernel(int32_t* __restrict__ cnt, int64_t nblocks, int64_t* total) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const int64_t per = (nblocks + 1023) / 1024;
  const int64_t b0 = (int64_t)t * per, b1 = (b0 + per < nblocks) ? b0 + per : nblocks;
  long long s = 0;
  for (int64_t b = b0; b < b1; ++b) s += cnt[b];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int i = 0; i < 1024; ++i) {
      const long long v = part[i];
      part[i] = run;
      run += v;
    }
    *total = run;
  }
  __syncthreads();
  long long run = part[t];
  for (int64_t b = b0; b < b1; ++b) {
    const int v = cnt[b];
    cnt[b] = (int32_t)run;  // member offsets fit int32 per shard (N < 2^31 rows per GPU)
    run += v;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
gather_members_kernel(const T* __restrict__ q, int64_t ldq, int64_t N, int k, const int32_t* __restrict__ blockoff,
                      const T* __restrict__ X, int64_t ldx, int D, const int32_t* __restrict__ gid,
                      T* __restrict__ Xk, int32_t* __restrict__ gidk, int64_t* __restrict__ map) {
  __shared__ int pos[kMemberBlock];
  __shared__ int wsum[kThreads / 32];
  __shared__ int base;
  const int64_t r0 = (int64_t)blockIdx.x * kMemberBlock;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  // ordered ranks inside the block: 4 sweeps of 256 rows
  for (int sweep = 0; sweep < kMemberBlock / kThreads; ++sweep) {
    const int i = sweep * kThreads + threadIdx.x;
    const int64_t n = r0 + i;
    const int f = (n < N && q[n * ldq + k] > (T)0.5) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int before = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int j = 0; j < w; ++j) woff += wsum[j];
    pos[i] = f ? (base + woff + before) : -1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int j = 0; j < kThreads / 32; ++j) s += wsum[j];
      base += s;
    }
    __syncthreads();
  }
  const int64_t off = blockoff[blockIdx.x];
  for (int i = threadIdx.x; i < kMemberBlock; i += kThreads) {
    if (pos[i] >= 0) {
      map[off + pos[i]] = r0 + i;
      if (gidk != nullptr) gidk[off + pos[i]] = gid != nullptr ? gid[r0 + i] : 0;
    }
  }
  for (int i = 0; i < kMemberBlock; ++i) {
    const int p = pos[i];
    if (p < 0) continue;
    const T* src = X + (r0 + i) * ldx;
    T* dst = Xk + (off + p) * ldx;
    for (int d = threadIdx.x; d < (int)ldx; d += kThreads) dst[d] = src[d];
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
split_side_kernel(const T* __restrict__ Xk, int64_t M, int D, int64_t ldx, const T* __restrict__ mc,
                  const T* __restrict__ v, T* __restrict__ qref, int64_t ldq, uint8_t* __restrict__ flags,
                  unsigned long long* __restrict__ scount) {
  // one warp per row
  const int lane = threadIdx.x & 31;
  const int64_t wglobal = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * kThreads) >> 5;
  unsigned long long cnt = 0;
  for (int64_t m = wglobal; m < M; m += nw) {
    T s = 0;
    for (int d = lane; d < D; d += 32) s = fma(Xk[m * ldx + d] - mc[d], v[d], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const bool side = s >= (T)0;
    if (lane == 0) {
      if (qref != nullptr) {
        qref[m * ldq + 0] = side ? (T)1 : (T)0;
        qref[m * ldq + 1] = side ? (T)0 : (T)1;
      }
      if (flags != nullptr) flags[m] = side ? 1 : 0;
      cnt += side ? 1 : 0;
    }
  }
  if (scount != nullptr && lane == 0 && cnt) atomicAdd(scount, cnt);
}

template <typename T>
__global__ void copy_q_kernel(const T* __restrict__ q, T* __restrict__ qaug, int64_t lds, int64_t ldd, int64_t N, int K,
                              int Knew) {
  const int64_t total = N * Knew;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / Knew;
    const int k = (int)(i - n * Knew);
    qaug[n * ldd + k] = (k < K) ? q[n * lds + k] : (T)0;
  }
}

template <typename T>
__global__ void aug_labels_kernel(const T* __restrict__ qref, int64_t ldqr, const int64_t* __restrict__ map, int64_t M,
                                  const T* __restrict__ q, T* __restrict__ qaug, int64_t ldq, int k, int K) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    if (qref[m * ldqr + 1] > (T)0.5) {
      const int64_t n = map[m];
      qaug[n * ldq + K] = q[n * ldq + k];
      qaug[n * ldq + k] = (T)0;
    }
  }
}

template <typename T>
__global__ void prune_columns_kernel(T* __restrict__ q, int64_t ldq, int64_t N, const int32_t* __restrict__ keep, int newK) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    T* row = q + n * ldq;
    for (int j = 0; j < newK; ++j) row[j] = row[keep[j]];
  }
}

inline int grid_for(int64_t work, int threads, int cap = 148 * 16) {
  int64_t g = (work + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

// ------------------------------------------------------------ launchers ----
int full_dp(int D) {
  if (D <= 16) return 16;
  if (D <= 32) return 32;
  if (D <= 64) return 64;
  if (D <= 128) return 128;
  if (D <= 256) return 256;
  return 0;
}

template <typename T> static int full_pt(int DP) {
  // points per thread: 4 (64-row tiles) unless shared memory forces 2
  const long need4 = (long)sizeof(T) * (2L * DP * (64 + 4) + 8L * DP);
  return need4 > 150 * 1024 ? 2 : 4;
}

template <typename T> long estep_full_smem(int D, int K) {
  const int DP = full_dp(D);
  if (DP == 0) return -1;
  const int PT = full_pt<T>(DP), TM = 16 * PT;
  const long b = (long)sizeof(T) * (2L * DP * (TM + 4) + 8L * DP + (long)TM * (K | 1)) + 4L * TM;
  return b > 227 * 1024 ? -1 : b;
}
template <typename T> long estep_diag_smem(int D, int K) {
  (void)D;
  const long b = (long)sizeof(T) * (32L * 68 + 3L * 32 * 68 + 64L * (K | 1)) + 4L * 64;
  return b > 227 * 1024 ? -1 : b;
}

template <typename T, int TN, int PT>
static cudaError_t launch_estep_full(cudaStream_t st, int sms, long smem, const T* X, int64_t N, int D, int64_t ldx,
                                     const int32_t* gid, int K, const T* RT, const T* mhi, const T* mlo, const T* chat,
                                     const T* lw, const uint8_t* act, T* q, int64_t ldq, int mode, double* Fz,
                                     double* H, const unsigned* skip) {
  auto kern = estep_full_kernel<T, TN, PT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t ntiles = (N + 16 * PT - 1) / (16 * PT);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sms * occ;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  kern<<<(int)grid, kThreads, smem, st>>>(X, N, D, ldx, gid, K, RT, mhi, mlo, chat, lw, act, q, ldq, mode, Fz, H, skip);
  return cudaGetLastError();
}

template <typename T>
cudaError_t estep_full(cudaStream_t st, int sms, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid, int K,
                       const T* RT, const T* mhi, const T* mlo, const T* chat, const T* lw, const uint8_t* act, T* q,
                       int64_t ldq, int mode, double* Fz, double* H, const unsigned* skip) {
  if (N <= 0) return cudaSuccess;
  const int DP = full_dp(D);
  const long smem = estep_full_smem<T>(D, K);
  if (DP == 0 || smem < 0 || K > 512) return cudaErrorInvalidValue;
  const int PT = full_pt<T>(DP);
#define LCB_CASE(TNV)                                                                                              \
  case 16 * TNV:                                                                                                   \
    return PT == 4 ? launch_estep_full<T, TNV, 4>(st, sms, smem, X, N, D, ldx, gid, K, RT, mhi, mlo, chat, lw, act, \
                                                  q, ldq, mode, Fz, H, skip)                                       \
                   : launch_estep_full<T, TNV, 2>(st, sms, smem, X, N, D, ldx, gid, K, RT, mhi, mlo, chat, lw, act, \
                                                  q, ldq, mode, Fz, H, skip);
  switch (DP) {
    LCB_CASE(1)
    LCB_CASE(2)
    LCB_CASE(4)
    LCB_CASE(8)
    LCB_CASE(16)
  }
#undef LCB_CASE
  return cudaErrorInvalidValue;
}

template <typename T>
cudaError_t estep_diag(cudaStream_t st, int sms, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid, int K,
                       const T* A, const T* mhi, const T* mlo, const T* chat, const T* lw, const uint8_t* act, T* q,
                       int64_t ldq, int mode, double* Fz, double* H, const unsigned* skip) {
  if (N <= 0) return cudaSuccess;
  const long smem = estep_diag_smem<T>(D, K);
  if (smem < 0 || K > 512) return cudaErrorInvalidValue;
  auto kern = estep_diag_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t ntiles = (N + 63) / 64;
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sms * occ;
  if (grid > ntiles) grid = ntiles;
  kern<<<(int)grid, kThreads, smem, st>>>(X, N, D, ldx, gid, K, A, mhi, mlo, chat, lw, act, q, ldq, mode, Fz, H, skip);
  return cudaGetLastError();
}

// rows folded into one CTA's low-precision partial sums before the fp64 atomics
template <typename T> static int rows_per_cta() { return sizeof(T) == 4 ? 8192 : 65536; }

template <typename T>
cudaError_t sstat_full(cudaStream_t st, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid, const T* q,
                       int64_t ldq, int K, const T* cen, const uint8_t* act, double* xs, double* S) {
  if (N <= 0 || K <= 0) return cudaSuccess;
  const int DP = full_dp(D);
  if (DP == 0) return cudaErrorInvalidValue;
  const int rpc = rows_per_cta<T>();
  const int64_t chunks = (N + rpc - 1) / rpc;
  if (chunks > 65535) return cudaErrorInvalidValue;
  const int tn = DP >= 128 ? 8 : DP / 16;
  const int nb = (D + 16 * tn - 1) / (16 * tn);
  dim3 grid(K, (unsigned)chunks, nb * nb);
  switch (tn) {
    case 1: sstat_full_kernel<T, 1><<<grid, kThreads, 0, st>>>(X, N, D, ldx, gid, q, ldq, K, DP, cen, act, rpc, nb, xs, S); break;
    case 2: sstat_full_kernel<T, 2><<<grid, kThreads, 0, st>>>(X, N, D, ldx, gid, q, ldq, K, DP, cen, act, rpc, nb, xs, S); break;
    case 4: sstat_full_kernel<T, 4><<<grid, kThreads, 0, st>>>(X, N, D, ldx, gid, q, ldq, K, DP, cen, act, rpc, nb, xs, S); break;
    default: sstat_full_kernel<T, 8><<<grid, kThreads, 0, st>>>(X, N, D, ldx, gid, q, ldq, K, DP, cen, act, rpc, nb, xs, S); break;
  }
  return cudaGetLastError();
}


template <typename T>
cudaError_t nz_count(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int K, const int32_t* gid, const uint8_t* act,
                     int32_t* blockcnt, double* Njk, int pred) {
  if (N <= 0) return cudaSuccess;
  int kw = 1;
  while (kw < K && kw < kThreads) kw <<= 1;
  nz_count_kernel<T><<<(unsigned)nz_blocks(N), kThreads, sizeof(int) * K, st>>>(q, ldq, N, K, gid, act, kw, blockcnt, Njk,
                                                                             pred);
  return cudaGetLastError();
}
cudaError_t nz_scan(cudaStream_t st, int32_t* blockcnt, int64_t nblocks, int K, long long* total) {
  nz_scan_kernel<<<K, kThreads, 0, st>>>(blockcnt, nblocks, K, total);
  return cudaGetLastError();
}
template <typename T>
cudaError_t nz_fill(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int K, const int32_t* gid, const uint8_t* act,
                    const int32_t* blockoff, const long long* koff, int32_t* lrow, T* lq, int pred, const unsigned* skip) {
  if (N <= 0) return cudaSuccess;
  int kw = 1;
  while (kw < K && kw < kThreads) kw <<= 1;
  nz_fill_kernel<T><<<(unsigned)nz_blocks(N), kThreads, sizeof(int) * K, st>>>(q, ldq, N, K, gid, act, kw, blockoff, koff,
                                                                            lrow, lq, pred, skip);
  return cudaGetLastError();
}
int64_t nz_blocks(int64_t N) { return (N + kNzBlock - 1) / kNzBlock; }

template <typename T, int TN>
static cudaError_t launch_gather_full(cudaStream_t st, dim3 grid, const T* X, int D, int64_t ldx, const int32_t* lrow,
                                      const T* lq, const long long* koff, const long long* kcnt, int DP, const T* cen,
                                      int nb, double* xs, double* S, const unsigned* skip) {
  constexpr int BW = 16 * TN, TMS = sizeof(T) == 8 ? 16 : 32;
  size_t smem = (((size_t)(2 * TMS * BW + TMS) * sizeof(T) + TMS * 4 + 15) & ~(size_t)15);
  if (sizeof(T) == 4) smem += sizeof(double) * ((size_t)TN * TN * kThreads + BW);
  auto kern = sstat_gather_full_kernel<T, TN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, smem, st>>>(X, D, ldx, lrow, lq, koff, kcnt, DP, cen, nb, xs, S, skip);
  return cudaGetLastError();
}

template <typename T>
cudaError_t sstat_gather_full(cudaStream_t st, const T* X, int D, int64_t ldx, const int32_t* lrow, const T* lq,
                              const long long* koff, const long long* kcnt, long long maxcnt, int K, const T* cen,
                              double* xs, double* S, const unsigned* skip) {
  if (K <= 0 || maxcnt <= 0) return cudaSuccess;
  const int DP = full_dp(D);
  if (DP == 0) return cudaErrorInvalidValue;
  const int tn = DP >= 128 ? 8 : DP / 16;
  const int nb = (D + 16 * tn - 1) / (16 * tn);
  long long chunks = (maxcnt + kGatherChunk - 1) / kGatherChunk;
  if (chunks > 4096) chunks = 4096;  // longer lists are strided over
  if (K > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)chunks, K, nb * nb);
  switch (tn) {
    case 1: return launch_gather_full<T, 1>(st, grid, X, D, ldx, lrow, lq, koff, kcnt, DP, cen, nb, xs, S, skip);
    case 2: return launch_gather_full<T, 2>(st, grid, X, D, ldx, lrow, lq, koff, kcnt, DP, cen, nb, xs, S, skip);
    case 4: return launch_gather_full<T, 4>(st, grid, X, D, ldx, lrow, lq, koff, kcnt, DP, cen, nb, xs, S, skip);
    default: return launch_gather_full<T, 8>(st, grid, X, D, ldx, lrow, lq, koff, kcnt, DP, cen, nb, xs, S, skip);
  }
}

template <typename T>
cudaError_t sstat_gather_diag(cudaStream_t st, const T* X, int D, int64_t ldx, const int32_t* lrow, const T* lq,
                              const long long* koff, const long long* kcnt, long long maxcnt, int K, const T* cen,
                              double* xs, double* S, const unsigned* skip) {
  if (K <= 0 || maxcnt <= 0) return cudaSuccess;
  if (D > 4 * kThreads || K > 65535) return cudaErrorInvalidValue;
  long long chunks = (maxcnt + 1023) / 1024;
  const long long cap = std::max<long long>(1, 4096 / K);  // ~4096 CTAs in all; longer lists are strided over
  if (chunks > cap) chunks = cap;
  sstat_gather_diag_kernel<T><<<dim3((unsigned)chunks, (unsigned)K), kThreads, 0, st>>>(X, D, ldx, lrow, lq, koff, kcnt, cen,
                                                                                       xs, S, skip);
  return cudaGetLastError();
}

template <typename T>
cudaError_t sstat_diag(cudaStream_t st, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid, const T* q,
                       int64_t ldq, int K, const T* cen, const uint8_t* act, double* xs, double* S) {
  if (N <= 0 || K <= 0) return cudaSuccess;
  const int rpc = rows_per_cta<T>();
  const int64_t chunks = (N + rpc - 1) / rpc;
  if (chunks > 65535) return cudaErrorInvalidValue;
  const int nkt = (K + 63) / 64, ndt = (D + 63) / 64;
  dim3 grid(nkt * ndt, (unsigned)chunks);
  sstat_diag_kernel<T><<<grid, kThreads, 0, st>>>(X, N, D, ldx, gid, q, ldq, K, cen, act, rpc, ndt, xs, S);
  return cudaGetLastError();
}

template <typename T>
cudaError_t colsum(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int K, const int32_t* gid, double* Njk) {
  if (N <= 0 || K <= 0) return cudaSuccess;
  int kw = 1;
  while (kw < K && kw < kThreads) kw <<= 1;
  const int rpc = 2048;
  const int64_t chunks = (N + rpc - 1) / rpc;
  colsum_kernel<T><<<(unsigned)chunks, kThreads, 0, st>>>(q, ldq, N, K, gid, kw, rpc, Njk);
  return cudaGetLastError();
}

template <typename T>
cudaError_t convert_rows(cudaStream_t st, const double* src, int64_t rows, int D, int64_t ld, int colmajor,
                         const double* mean, T* dst, int64_t ldx) {
  if (rows <= 0) return cudaSuccess;
  convert_rows_kernel<T><<<grid_for(rows * ldx, 256), 256, 0, st>>>(src, rows, D, ld, colmajor, mean, dst, ldx);
  return cudaGetLastError();
}
template <typename T>
cudaError_t convert_f32(cudaStream_t st, const float* src, int64_t rows, int D, int64_t ld, const double* mean, T* dst,
                        int64_t ldx) {
  if (rows <= 0) return cudaSuccess;
  convert_f32_kernel<T><<<grid_for(rows * ldx, 256), 256, 0, st>>>(src, rows, D, ld, mean, dst, ldx);
  return cudaGetLastError();
}
cudaError_t colsum_f32(cudaStream_t st, const float* src, int64_t rows, int D, int64_t ld, double* sums) {
  if (rows <= 0) return cudaSuccess;
  const int rpc = 4096;
  colsum_f32_kernel<<<(unsigned)((rows + rpc - 1) / rpc), 128, 0, st>>>(src, rows, D, ld, rpc, sums);
  return cudaGetLastError();
}
template <typename T> cudaError_t absmax(cudaStream_t st, const T* x, int64_t n, unsigned* out_bits) {
  if (n <= 0) return cudaSuccess;
  absmax_kernel<T><<<grid_for(n, 256), 256, 0, st>>>(x, n, out_bits);
  return cudaGetLastError();
}
template <typename T> cudaError_t fill_ones(cudaStream_t st, T* q, int64_t ldq, int64_t N) {
  if (N <= 0) return cudaSuccess;
  fill_ones_kernel<T><<<grid_for(N, 256), 256, 0, st>>>(q, ldq, N);
  return cudaGetLastError();
}
template <typename T> cudaError_t labels_to_q(cudaStream_t st, const int32_t* lab, T* q, int64_t ldq, int64_t N, int K) {
  if (N <= 0) return cudaSuccess;
  labels_to_q_kernel<T><<<grid_for(N * K, 256), 256, 0, st>>>(lab, q, ldq, N, K);
  return cudaGetLastError();
}
template <typename T> cudaError_t q_from_double(cudaStream_t st, const double* src, int64_t rows, int K, T* q, int64_t ldq) {
  if (rows <= 0) return cudaSuccess;
  q_from_double_kernel<T><<<grid_for(rows * K, 256), 256, 0, st>>>(src, rows, K, q, ldq);
  return cudaGetLastError();
}
template <typename T>
cudaError_t q_to_double(cudaStream_t st, const T* q, int64_t ldq, int64_t rows, int K, double* dst, int64_t ld,
                        int colmajor) {
  if (rows <= 0) return cudaSuccess;
  q_to_double_kernel<T><<<grid_for(rows * K, 256), 256, 0, st>>>(q, ldq, rows, K, dst, ld, colmajor);
  return cudaGetLastError();
}
template <typename T>
cudaError_t member_counts(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int k, int32_t* blockcnt) {
  if (N <= 0) return cudaSuccess;
  member_counts_kernel<T><<<(unsigned)((N + kMemberBlock - 1) / kMemberBlock), kThreads, 0, st>>>(q, ldq, N, k, blockcnt);
  return cudaGetLastError();
}
cudaError_t scan_counts(cudaStream_t st, int32_t* blockcnt, int64_t nblocks, int64_t* total) {
  scan_counts_kernel<<<1, 1024, 0, st>>>(blockcnt, nblocks, total);
  return cudaGetLastError();
}
template <typename T>
cudaError_t gather_members(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int k, const int32_t* blockoff,
                           const T* X, int64_t ldx, int D, const int32_t* gid, T* Xk, int32_t* gidk, int64_t* map) {
  if (N <= 0) return cudaSuccess;
  gather_members_kernel<T><<<(unsigned)((N + kMemberBlock - 1) / kMemberBlock), kThreads, 0, st>>>(
      q, ldq, N, k, blockoff, X, ldx, D, gid, Xk, gidk, map);
  return cudaGetLastError();
}
template <typename T>
cudaError_t split_side(cudaStream_t st, const T* Xk, int64_t M, int D, int64_t ldx, const T* mc, const T* v, T* qref,
                       int64_t ldq, unsigned long long* scount) {
  if (M <= 0) return cudaSuccess;
  split_side_kernel<T><<<grid_for(M * 32, kThreads), kThreads, 0, st>>>(Xk, M, D, ldx, mc, v, qref, ldq, nullptr, scount);
  return cudaGetLastError();
}
template <typename T>
cudaError_t side_flags(cudaStream_t st, const T* Xk, int64_t M, int D, int64_t ldx, const T* mc, const T* v,
                       uint8_t* out) {
  if (M <= 0) return cudaSuccess;
  split_side_kernel<T><<<grid_for(M * 32, kThreads), kThreads, 0, st>>>(Xk, M, D, ldx, mc, v, nullptr, 0, out, nullptr);
  return cudaGetLastError();
}
template <typename T>
cudaError_t copy_q(cudaStream_t st, const T* q, T* qaug, int64_t lds, int64_t ldd, int64_t N, int K, int Knew) {
  if (N <= 0) return cudaSuccess;
  copy_q_kernel<T><<<grid_for(N * Knew, 256), 256, 0, st>>>(q, qaug, lds, ldd, N, K, Knew);
  return cudaGetLastError();
}
template <typename T>
cudaError_t aug_labels(cudaStream_t st, const T* qref, int64_t ldqr, const int64_t* map, int64_t M, const T* q,
                       T* qaug, int64_t ldq, int k, int K) {
  if (M <= 0) return cudaSuccess;
  aug_labels_kernel<T><<<grid_for(M, 256), 256, 0, st>>>(qref, ldqr, map, M, q, qaug, ldq, k, K);
  return cudaGetLastError();
}
template <typename T>
cudaError_t prune_columns(cudaStream_t st, T* q, int64_t ldq, int64_t N, const int32_t* keep, int newK) {
  if (N <= 0) return cudaSuccess;
  prune_columns_kernel<T><<<grid_for(N, 256), 256, 0, st>>>(q, ldq, N, keep, newK);
  return cudaGetLastError();
}

// explicit instantiations for both arithmetic types
#define LCB_INST(T)                                                                                                   \
  template long estep_full_smem<T>(int, int);                                                                         \
  template long estep_diag_smem<T>(int, int);                                                                         \
  template cudaError_t estep_full<T>(cudaStream_t, int, const T*, int64_t, int, int64_t, const int32_t*, int,         \
                                     const T*, const T*, const T*, const T*, const T*, const uint8_t*, T*, int64_t,  \
                                     int, double*, double*, const unsigned*);                                         \
  template cudaError_t estep_diag<T>(cudaStream_t, int, const T*, int64_t, int, int64_t, const int32_t*, int,         \
                                     const T*, const T*, const T*, const T*, const T*, const uint8_t*, T*, int64_t,  \
                                     int, double*, double*, const unsigned*);                                         \
  template cudaError_t sstat_full<T>(cudaStream_t, const T*, int64_t, int, int64_t, const int32_t*, const T*,         \
                                     int64_t, int, const T*, const uint8_t*, double*, double*);                       \
  template cudaError_t sstat_diag<T>(cudaStream_t, const T*, int64_t, int, int64_t, const int32_t*, const T*,         \
                                     int64_t, int, const T*, const uint8_t*, double*, double*);                       \
  template cudaError_t colsum<T>(cudaStream_t, const T*, int64_t, int64_t, int, const int32_t*, double*);             \
  template cudaError_t nz_count<T>(cudaStream_t, const T*, int64_t, int64_t, int, const int32_t*, const uint8_t*,     \
                                   int32_t*, double*, int);                                                           \
  template cudaError_t nz_fill<T>(cudaStream_t, const T*, int64_t, int64_t, int, const int32_t*, const uint8_t*,      \
                                  const int32_t*, const long long*, int32_t*, T*, int, const unsigned*);              \
  template cudaError_t sstat_gather_diag<T>(cudaStream_t, const T*, int, int64_t, const int32_t*, const T*,           \
                                            const long long*, const long long*, long long, int, const T*, double*,   \
                                            double*, const unsigned*);                                                \
  template cudaError_t sstat_gather_full<T>(cudaStream_t, const T*, int, int64_t, const int32_t*, const T*,           \
                                            const long long*, const long long*, long long, int, const T*, double*,   \
                                            double*, const unsigned*);                                                \
  template cudaError_t convert_rows<T>(cudaStream_t, const double*, int64_t, int, int64_t, int, const double*, T*,    \
                                       int64_t);                                                                      \
  template cudaError_t convert_f32<T>(cudaStream_t, const float*, int64_t, int, int64_t, const double*, T*, int64_t); \
  template cudaError_t fill_ones<T>(cudaStream_t, T*, int64_t, int64_t);                                              \
  template cudaError_t absmax<T>(cudaStream_t, const T*, int64_t, unsigned*);                                         \
  template cudaError_t labels_to_q<T>(cudaStream_t, const int32_t*, T*, int64_t, int64_t, int);                       \
  template cudaError_t q_from_double<T>(cudaStream_t, const double*, int64_t, int, T*, int64_t);                      \
  template cudaError_t q_to_double<T>(cudaStream_t, const T*, int64_t, int64_t, int, double*, int64_t, int);          \
  template cudaError_t member_counts<T>(cudaStream_t, const T*, int64_t, int64_t, int, int32_t*);                     \
  template cudaError_t gather_members<T>(cudaStream_t, const T*, int64_t, int64_t, int, const int32_t*, const T*,     \
                                         int64_t, int, const int32_t*, T*, int32_t*, int64_t*);                       \
  template cudaError_t split_side<T>(cudaStream_t, const T*, int64_t, int, int64_t, const T*, const T*, T*, int64_t,  \
                                     unsigned long long*);                                                            \
  template cudaError_t side_flags<T>(cudaStream_t, const T*, int64_t, int, int64_t, const T*, const T*, uint8_t*);    \
  template cudaError_t copy_q<T>(cudaStream_t, const T*, T*, int64_t, int64_t, int64_t, int, int);                    \
  template cudaError_t aug_labels<T>(cudaStream_t, const T*, int64_t, const int64_t*, int64_t, const T*, T*, int64_t, \
                                     int, int);                                                                       \
  template cudaError_t prune_columns<T>(cudaStream_t, T*, int64_t, int64_t, const int32_t*, int);
LCB_INST(float)
LCB_INST(double)
#undef LCB_INST

}  // namespace dev
}  // namespace lcb
