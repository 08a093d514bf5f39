// tc_kernels.cuh -- tcgen05 tier of the E step (see tc_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lcb {
namespace dev {

constexpr uint32_t kTcBlobBytes = 50176;  // pre-swizzled fp16 hi/lo operand of one cluster + its mean (hi, -s*lo)

// true when the fp32 engine can run the full-covariance E step on the tensor cores
bool tc_supported(int D, int64_t ldx);

cudaError_t estep_tc128(cudaStream_t st, int sms, const float* X, int64_t N, const int32_t* gid, int K,
                        const uint8_t* blob, const float* ascale, const float* inv_t2, const float* chat,
                        const float* lw, const uint8_t* act, float* q, int64_t ldq, double* Fz, unsigned* err);

// Centred scatter over the per-cluster non-zero lists (see tc_kernels.cu); cen [K][128] is relative to the data
// centre, scale a power of two with scale * max|x - c| <= 2^14.
constexpr int kTcScatterChunk = 512;  // list rows folded into the fp32 accumulators before the fp64 add (32 tensor-core additions)
cudaError_t sstat_tc128(cudaStream_t st, const float* X, const int32_t* lrow, const float* lq, const long long* koff,
                        const long long* kcnt, long long maxcnt, long long nnz, int K, const float* cen, float scale,
                        double* xs, double* S, unsigned* err);

void tc_pack_cluster(const double* R, double bscale, const double* mean_rel, double ascale, uint8_t* out);

}  // namespace dev
}  // namespace lcb
