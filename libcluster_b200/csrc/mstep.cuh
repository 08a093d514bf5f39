// mstep.cuh -- the M step of the VB iteration on the device (see mstep.cu).
//
// Replaces, for the iterations of vbem(), the host-side posterior updates of host_model.cpp: the statistics never
// leave the GPU, the E-step operands are produced where they are consumed, and the host reads back one small record
// per iteration (IterRec) to run the reference's convergence tests (src/cluster.cpp:229-236).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lcb {
namespace dev {

// One record per iteration, written by the device and copied to the host in one piece (doubles).
enum IterSlot {
  kItSumLogZ = 0,   // sum_n log Z'_n of this rank; all-reduced together with kItRerun
  kItRerun = 1,     // ranks whose two-level E pass has to be repeated with the dense kernel
  kItFc = 2,        // sum_k Fc_k
  kItFw = 3,        // sum_j Fw_j
  kItCbar = 4,      // mean of the cluster constants (subtracted from the logits, added back to F)
  kItPairs = 5,     // candidate pairs of the two-level E pass
  kItItems = 6,     // 128-pair work items of level 2
  kItMaxCnt = 7,    // longest candidate list
  kItNnzS = 8,      // entries of the non-zero lists of the S pass
  kItMaxCntS = 9,   // longest of them
  kItAbort = 10,    // != 0: some rank's S-pass lists overflowed; the iteration did nothing and is repeated
  kItMFail = 11,    // 1: iW not positive definite, 2: NormGamma variance <= 0
  kItOverS = 12,    // this rank's S-pass lists overflowed (capacity to grow)
  kItOverE = 13,    // this rank's candidate lists overflowed or were not worth it
  kItAugFail = 14,  // the level-1 centring chunk cannot be represented: dense kernel instead
  kItCount = 16
};

// Control words read by the kernels of the iteration.
enum CtlWord {
  kCtlSkipS = 0,  // S pass: lists overflowed, skip nz_fill / the statistics kernels
  kCtlSkipE = 1,  // E pass: abort or M-step failure, leave q untouched
  kCtlSkipL = 2,  // two-level: skip list fill, level 2 and the soft-max (dense kernel follows)
  kCtlAugH = 3,   // fp16 bits of 2^P (A slots of the level-1 centring chunk), replicated in both halves
  kCtlCount = 8
};

struct MStepArgs {
  int J, K, D, ckind, wkind;
  int cld;        // leading dimension of the centre table (T)
  int path;       // 0: SIMT operands (RT | mhi | mlo | chat | lw in T), 1: tcgen05 operands (D == 128, fp32)
  int two_level;  // path 1: also the level-1 operands (aug blocks, cpar)
  double prior;   // clustwidth
  double Fp;      // sum_l lgamma((nu_p + 1 - l) / 2) of the GaussWish prior
  double a1p, a2p, Fwp;  // weight prior and its lgamma constant
  double sg, xabs_max;   // level-1 data scale and max |x| of the resident rows
  double ntot;           // rows over all ranks of the view
  int64_t nstat;         // doubles in stats before the abort slot
  const double* stats;   // [Njk (J*K) | xs (K*D) | S (K*Sz)] centred statistics, then the abort slot
  const uint8_t* act;    // optional sparse mask [J][K]
  const double* centre;  // [D] data centre subtracted at upload
  void* cen;             // T [K][cld]: centres of this S pass in, centres of the next one out
  double* raw;           // [K][1 + D + Sz] raw statistics N_s, x_s, xx_s (host model sync)
  double* post;          // [K][kPostStride]
  double* work;          // fp64 scratch [K][D * (D | 1)] when the factorisation does not fit shared memory
  double* iter;          // IterRec
  unsigned* ctl;
  float* sscale;         // operand scale of the tensor-core scatter of the next S pass
  // SIMT operands
  void* RT; void* mhi; void* mlo; void* chat; void* lw;
  // tcgen05 operands
  uint8_t* blob; float* as; float* it2; float* chatf; float* lwf; uint8_t* aug; float* cpar; double* vaug;
  double* wscr;          // weight scratch [J][6 K]
};
constexpr int kPostStride = 16;
enum PostSlot { kPN = 0, kPNu, kPBeta, kPLogdW, kPCconst, kPFc, kPS, kPT, kPRfro, kPVmax, kPCmax, kPFail };

template <typename T> cudaError_t mstep(cudaStream_t st, const MStepArgs& a);
// dynamic shared memory of the per-cluster kernel; > 200 KB means the global scratch is used instead
size_t mstep_work_doubles(int D);

// Centres of clusters from a probe statistics pass: cen[k] = weighted mean of the rows (where N_k > 0).
template <typename T>
cudaError_t centres_from_stats(cudaStream_t st, const double* stats, int J, int K, int D, int cld, const double* centre,
                               T* cen);
// act[j][k] = Njk >= cutoff (sparse updates, cluster.cpp:69-70)
cudaError_t build_act(cudaStream_t st, const double* Njk, int64_t n, double cutoff, uint8_t* act);
// Lists of the S pass / candidate lists of the E pass from the per-cluster totals: offsets, work items, capacity check.
//   tot [K] -> koff [K], itoff [K + 1] (128-entry items, may be NULL); iter[slot_n], iter[slot_max] <- totals;
//   overflow (sum > cap, or sum > too_many when too_many >= 0): ctl[skip_word] = 1, iter[over_slot] = 1,
//   *abort_slot = 1 (when given).  nitems_out (int64) receives the number of items (0 on overflow).
cudaError_t list_plan(cudaStream_t st, const long long* tot, int K, long long cap, double too_many, long long* koff,
                      int32_t* itoff, long long* nitems_out, double* iter, int slot_n, int slot_max, int over_slot,
                      unsigned* ctl, int skip_word, double* abort_slot, int vote_slot);

// host-side check of the std::sort restatement (mstep_math.hpp): sorts ids by count, greater first
void sort_desc_like_std(const double* v, int n, int* ids);

}  // namespace dev
}  // namespace lcb
