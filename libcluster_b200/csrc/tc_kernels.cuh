// tc_kernels.cuh -- tcgen05 tier of the E step (see tc_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lcb {
namespace dev {

constexpr uint32_t kTcBlobBytes = 50176;  // pre-swizzled fp16 hi/lo operand of one cluster + its mean (hi, -s*lo)

// true when the fp32 engine can run the full-covariance E step on the tensor cores
bool tc_supported(int D, int64_t ldx);
// 128 or 64 when the device-resident iteration (engine_dev.cu) can run the tcgen05 kernels on rows of that width,
// else 0.  Every tensor-core launcher below takes that width as its last argument (`dim`, default 128); 64-wide rows
// use the same operand blobs with R_k in the upper-left 64 x 64 corner (mstep.cu) and skip everything above it.
int tc_dim(int D, int64_t ldx);

cudaError_t estep_tc128(cudaStream_t st, int sms, const float* X, int64_t N, const int32_t* gid, int K,
                        const uint8_t* blob, const float* ascale, const float* inv_t2, const float* chat,
                        const float* lw, const uint8_t* act, float* q, int64_t ldq, double* Fz, unsigned* err,
                        const unsigned* skip = nullptr, int dim = 128);
// Device-driven variants (mstep.cuh): `skip` points at a control word; the kernel returns at once when it is set
// (the two-level kernels look at skip[0] | skip[1]: aborted iteration, or the dense kernel takes over).

// ---- two-level E step ------------------------------------------------------------------
// level 1: one-product distances with a rigorous error bound.  q[n][k] = upper bound of the logit; cmask [N][W],
//          W = ceil(K / 32): bit k of row n set iff the pair can reach e^-margin of the row's best (a *candidate*).
//          cpar [4][K] = {1/(s_g tau_k)^2, Ek, Ea, chat}; augblob = ceil(K/4) blocks of kTcAugBlockBytes
//          (tc_pack_aug); aug_exp = P (A slot = 2^P)
constexpr uint32_t kTcAugBlockBytes = 16384;
constexpr int kTcCoarseMaxK = 256;
cudaError_t estep_coarse_tc128(cudaStream_t st, int sms, const float* X, const float* xnorm, int64_t N,
                               const int32_t* gid, int K, const uint8_t* blob, const uint8_t* augblob,
                               const float* cpar, const float* lw, const uint8_t* act, float sg, int aug_exp,
                               float margin, float* q, int64_t ldq, uint32_t* cmask, uint32_t sbase_hint,
                               unsigned* err, const unsigned* augh_dev = nullptr, const unsigned* skip = nullptr,
                               int dim = 128);
// err[1] of a launch whose sbase_hint did not match: 0x80000000 | the 1024-aligned shared-memory base to pass instead
constexpr uint32_t kTcSbaseDefault = 1024;
// candidate masks -> per-cluster row lists: mask_count, nz_scan (kernels.cuh), mask_fill
cudaError_t mask_count(cudaStream_t st, const uint32_t* cmask, int64_t N, int K, int32_t* blockcnt,
                       const unsigned* skip = nullptr);
cudaError_t mask_fill(cudaStream_t st, const uint32_t* cmask, int64_t N, int K, int32_t* blockoff, const long long* koff,
                      int32_t* lrow, const unsigned* skip = nullptr);
// level 2: exact logits of the candidate pairs, written into q[row][k].  itoff [K+1] = prefix of
//          ceil(kcnt[k] / 128); items = scratch of 16 bytes per 128-pair work item
cudaError_t estep_tc128_list(cudaStream_t st, int sms, const float* X, int64_t N, const int32_t* gid, int K,
                             const uint8_t* blob, const float* ascale, const float* inv_t2, const float* chat,
                             const float* lw, const int32_t* lrow, const long long* koff, const long long* kcnt,
                             const int32_t* itoff, int64_t nitems, void* items, float* q, int64_t ldq, unsigned* err,
                             const long long* nitems_dev = nullptr, const unsigned* skip = nullptr, int dim = 128);
// level 3: q = softmax over the candidate logits (0 elsewhere), Fz += sum log Z
//          Hk (optional, [K], accumulated): split scores sum_n q_nk * logit_nk
cudaError_t estep_finalize(cudaStream_t st, int sms, float* q, int64_t ldq, int64_t N, int K, const uint32_t* cmask,
                           double* Fz, const unsigned* skip = nullptr, double* Hk = nullptr);
// q = -inf outside the candidate mask (test modes only)
cudaError_t apply_candidate_mask(cudaStream_t st, float* q, int64_t ldq, int64_t N, int K, const uint32_t* cmask);
// lq[e] = q[lrow[e]][k] over the lists, Nk[k] += sum (grouped models, gid != NULL: Nk[gid[row]][k]): lets the
// statistics pass reuse the candidate lists
cudaError_t gather_list_q(cudaStream_t st, int sms, const float* q, int64_t ldq, const int32_t* lrow,
                          const long long* koff, const long long* kcnt, long long maxcnt, int K, float* lq, double* Nk,
                          const int32_t* gid = nullptr);
cudaError_t row_norm128(cudaStream_t st, int sms, const float* X, int64_t N, float* out, int dim = 128);
double tc_pack_aug(const double* w, int k, uint8_t* augblob);

// Centred scatter over the per-cluster non-zero lists (see tc_kernels.cu); cen [K][128] is relative to the data
// centre, scale a power of two with scale * max|x - c| <= 2^14.
constexpr int kTcScatterChunk = 512;  // list rows folded into the fp32 accumulators before the fp64 add (32 tensor-core additions)
cudaError_t sstat_tc128(cudaStream_t st, int sms, const float* X, const int32_t* lrow, const float* lq,
                        const long long* koff, const long long* kcnt, long long maxcnt, long long nnz, int K,
                        const float* cen, float scale, double* xs, double* S, unsigned* err,
                        const float* scale_dev = nullptr, const unsigned* skip = nullptr, int dim = 128);

void tc_pack_cluster(const double* R, double bscale, const double* mean_rel, double ascale, uint8_t* out);
int tc_pack_selftest();  // bytes differing between the F16C and the portable packing of one operand

}  // namespace dev
}  // namespace lcb
