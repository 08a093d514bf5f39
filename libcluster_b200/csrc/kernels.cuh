// kernels.cuh -- launchers of the sm_100a kernels of the VB pass.
//
// Device data model (T = float for LCB_F32, double for LCB_F64):
//   X    [N x ldx]  row-major observations, already minus the global column mean
//   q    [N x ldq]  row-major responsibilities qZ
//   gid  [N]        int32 group of every row (NULL when J == 1), non-decreasing
// Per-iteration parameters, uploaded by the host M-step (engine.cpp):
//   full covariance (GaussWish):  RT [K][DP][DP] with RT[k][d][i] = R_k[i][d],
//       R_k = sqrt(nu_k) L_k^-1 lower-triangular, zero padded to DP = 16*TN;
//       mhi/mlo [K][DP] mean split in two T words; chat [K] = c_k - cbar
//   diagonal (NormGamma): A [K][D] = nu_k / L_kd ; mhi/mlo [K][D] ; chat [K]
//   lw [J][K] = E[log pi_jk];  act [J][K] optional sparse-update mask (uint8)
// Reductions leave the SMs as fp64 atomics into:
//   Fz (1)  sum_n logZ'_n ;  H [K]  sum_n q_nk * logit'_nk (split scores)
//   Njk [J][K] ;  xs [K][D], S [K][D][D] (or [K][D]) centred about cen[k]
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lcb {
namespace dev {

enum EMode { kEWrite = 0, kEScore = 1, kERawLogit = 2 };

// Padded dimension used by the full-covariance kernels; 0 if unsupported.
int full_dp(int D);
// Dynamic shared memory the E kernels need; <0 if the shape is unsupported.
template <typename T> long estep_full_smem(int D, int K);
template <typename T> long estep_diag_smem(int D, int K);

template <typename T>
cudaError_t estep_full(cudaStream_t st, int sms, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid,
                       int K, const T* RT, const T* mhi, const T* mlo, const T* chat, const T* lw,
                       const uint8_t* act, T* q, int64_t ldq, int mode, double* Fz, double* H,
                       const unsigned* skip = nullptr);

template <typename T>
cudaError_t estep_diag(cudaStream_t st, int sms, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid,
                       int K, const T* A, const T* mhi, const T* mlo, const T* chat, const T* lw,
                       const uint8_t* act, T* q, int64_t ldq, int mode, double* Fz, double* H,
                       const unsigned* skip = nullptr);
// `skip` (device word, may be NULL): the kernel returns at once when it is non-zero -- the iteration was aborted on
// the device (list overflow, failed M step) and the host repeats or reports it (mstep.cuh).

// Sufficient statistics about per-cluster centres cen [K][DP] (full) / [K][D] (diag).
template <typename T>
cudaError_t sstat_full(cudaStream_t st, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid, const T* q,
                       int64_t ldq, int K, const T* cen, const uint8_t* act, double* xs, double* S);
template <typename T>
cudaError_t sstat_diag(cudaStream_t st, const T* X, int64_t N, int D, int64_t ldx, const int32_t* gid, const T* q,
                       int64_t ldq, int K, const T* cen, const uint8_t* act, double* xs, double* S);
template <typename T>
cudaError_t colsum(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int K, const int32_t* gid, double* Njk);

// Non-zero responsibilities as per-cluster (row, q) lists and the full-covariance
// statistics over those lists (work proportional to nnz(q), not N*K).
int64_t nz_blocks(int64_t N);  // row blocks used by nz_count / nz_fill
// which entries go on the lists: q != 0 (responsibilities, after the sparse mask) or q != -inf (the candidate
// marking left by estep_coarse_tc128; the mask and Njk are ignored and lq may be NULL)
enum NzPred { kNzNonZero = 0, kNzNotNegInf = 1 };
template <typename T>
cudaError_t nz_count(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int K, const int32_t* gid, const uint8_t* act,
                     int32_t* blockcnt /* [nz_blocks][K] */, double* Njk /* optional fused column sums */,
                     int pred = kNzNonZero);
cudaError_t nz_scan(cudaStream_t st, int32_t* blockcnt, int64_t nblocks, int K, long long* total /* [K] */);
template <typename T>
cudaError_t nz_fill(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int K, const int32_t* gid, const uint8_t* act,
                    const int32_t* blockoff, const long long* koff, int32_t* lrow, T* lq, int pred = kNzNonZero,
                    const unsigned* skip = nullptr);
template <typename T>
cudaError_t sstat_gather_full(cudaStream_t st, const T* X, int D, int64_t ldx, const int32_t* lrow, const T* lq,
                              const long long* koff, const long long* kcnt, long long maxcnt, int K, const T* cen,
                              double* xs, double* S, const unsigned* skip = nullptr);

// the same for diagonal models: xs_k += sum q (x - c_k), S_k += sum q (x - c_k)^2 over the lists (D <= 1024)
template <typename T>
cudaError_t sstat_gather_diag(cudaStream_t st, const T* X, int D, int64_t ldx, const int32_t* lrow, const T* lq,
                              const long long* koff, const long long* kcnt, long long maxcnt, int K, const T* cen,
                              double* xs, double* S, const unsigned* skip = nullptr);

// ---- data movement / labels / split bookkeeping ---------------------------
// dst[n][d] = T(src[n][d] - mean[d]) from a staged block of doubles in either order
template <typename T>
cudaError_t convert_rows(cudaStream_t st, const double* src, int64_t rows, int D, int64_t ld, int colmajor,
                         const double* mean, T* dst, int64_t ldx);
template <typename T>
cudaError_t convert_f32(cudaStream_t st, const float* src, int64_t rows, int D, int64_t ld, const double* mean,
                        T* dst, int64_t ldx);
cudaError_t colsum_f32(cudaStream_t st, const float* src, int64_t rows, int D, int64_t ld, double* sums);
template <typename T> cudaError_t fill_ones(cudaStream_t st, T* q, int64_t ldq, int64_t N);
// bit pattern of max |x| over n elements, combined into *out_bits with atomicMax (zero it first)
template <typename T> cudaError_t absmax(cudaStream_t st, const T* x, int64_t n, unsigned* out_bits);
template <typename T> cudaError_t labels_to_q(cudaStream_t st, const int32_t* lab, T* q, int64_t ldq, int64_t N, int K);
template <typename T> cudaError_t q_from_double(cudaStream_t st, const double* src, int64_t rows, int K, T* q, int64_t ldq);
template <typename T>
cudaError_t q_to_double(cudaStream_t st, const T* q, int64_t ldq, int64_t rows, int K, double* dst, int64_t ld,
                        int colmajor);
// members of cluster k (q[n][k] > 0.5): ordered compaction of rows into Xk / gidk / map
template <typename T>
cudaError_t member_counts(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int k, int32_t* blockcnt);
cudaError_t scan_counts(cudaStream_t st, int32_t* blockcnt, int64_t nblocks, int64_t* total);
template <typename T>
cudaError_t gather_members(cudaStream_t st, const T* q, int64_t ldq, int64_t N, int k, const int32_t* blockoff,
                           const T* X, int64_t ldx, int D, const int32_t* gid, T* Xk, int32_t* gidk, int64_t* map);
// hard two-way split of the gathered rows by sign((x - m) . v); counts side==true
template <typename T>
cudaError_t split_side(cudaStream_t st, const T* Xk, int64_t M, int D, int64_t ldx, const T* mc, const T* v,
                       T* qref, int64_t ldq, unsigned long long* scount);
// 1 where (x - m) . v >= 0 (operator-level splitobs)
template <typename T>
cudaError_t side_flags(cudaStream_t st, const T* Xk, int64_t M, int D, int64_t ldx, const T* mc, const T* v,
                       uint8_t* out);
// qaug = q with a new zero column K; rows map[m] with qref[m][1] > 0.5 move column k -> K
template <typename T>
cudaError_t copy_q(cudaStream_t st, const T* q, T* qaug, int64_t ldq_src, int64_t ldq_dst, int64_t N, int K,
                   int Knew);
template <typename T>
cudaError_t aug_labels(cudaStream_t st, const T* qref, int64_t ldqr, const int64_t* map, int64_t M, const T* q,
                       T* qaug, int64_t ldq, int k, int K);
// in-place removal of pruned columns: new column j takes old column keep[j] (keep ascending)
template <typename T>
cudaError_t prune_columns(cudaStream_t st, T* q, int64_t ldq, int64_t N, const int32_t* keep, int newK);

}  // namespace dev
}  // namespace lcb
