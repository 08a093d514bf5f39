// tc_kernels.cu -- tcgen05 (5th-gen tensor core) tier of the E step, sm_100a only.
//
// estep_tc128_kernel: full-covariance E step for D == 128 in the fp32 engine.
//   logit[n,k] = chat_k + lw[g,k] - 0.5 * | R_k (x_n - m_k) |^2
// The whitening Y_k = (X - m_k) R_k^T of one 128-point tile is a 128x128x128
// GEMM per cluster.  It runs on the tensor cores at fp32-equivalent accuracy
// by splitting both operands into fp16 (hi, lo) pairs and issuing the three
// significant products  hi*hi + hi*lo + lo*hi  into one fp32 TMEM accumulator:
//   A_k = s_k (X - m_k)   centred per cluster in fp32 *before* the split (the
//                         cancellation must not happen inside the accumulator)
//   B_k = (t_k / s_k) R_k lower-triangular; packed by the host as pre-swizzled
//                         fp16 hi/lo blobs (engine.cu: pack_tc_operands)
// with s_k, t_k powers of two chosen so that both stay in fp16's normal range.
// Because R_k is lower-triangular the K-chunk covering input dims [16c,16c+16)
// only feeds output columns i >= 16c: the MMA for that chunk runs with
// N = 128 - 16c (56 % of the dense work).
//
// One persistent CTA per SM, 16 warps:
//   warp 0      producer: cp.async.bulk (TMA engine) of the X tile and of the
//               per-cluster operand blobs into a 2-stage ring
//   warp 1      one lane issues tcgen05.mma, commits to mbarriers
//   warp 2      TMEM allocation (2 accumulators x 128 columns)
//   warps 4-7   epilogue: tcgen05.ld accumulator -> sum of squares -> logit;
//               after the last cluster: row soft-max, q and log Z
//   warps 8-15  operand builders: centre, scale, split, swizzled st.shared of
//               A (4 warps per 64-dim K block so the blocks pipeline with MMA)
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "tc_kernels.cuh"

#include <cstring>
#include <vector>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

namespace lcb {
namespace dev {

namespace {

constexpr int kD = 128;  // widest row of the tier (the kernels are templated on DIM = 128 or 64)
[[maybe_unused]] constexpr int kDUnused = kD;
constexpr int kTM = 128;
constexpr uint32_t kAPart = kTM * 128;            // one fp16 [128 x 64] swizzled block: 16384
constexpr uint32_t kBBlob = kTcBlobBytes;         // B: kb0 hi(16K) lo(16K) kb1 hi(8K) lo(8K) | mhi (512) | -s*mlo (512)
constexpr uint32_t kOffMean = 49152;              // offset of the mean vectors inside a stage
constexpr int kStages = 4;
constexpr uint32_t kOffBar = kStages * kBBlob;    // 200704
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;
constexpr uint32_t kTransRow = 272;               // list mode: 64 fp32 of a row + 16 B of padding (conflict-free transposition)
constexpr uint32_t kTransWarp = 32 * kTransRow;   // one builder warp's gather buffer
constexpr uint32_t kOffTrans = 2 * kBBlob;        // list mode: behind its two operand stages
static_assert(kOffTrans + 8 * kTransWarp <= kOffBar, "gather buffers overlap the barriers");
constexpr int kThreadsTc = 512;               // 4 control + 4 epilogue + 8 operand-builder warps
constexpr uint32_t kTmemCols = 512;               // D0 [0,128) D1 [128,256) A0 hi/lo [256,384) A1 hi/lo [384,512)

// barrier slots (8 bytes each) after kOffBar
enum { BB_FULL0 = 0, BB_EMPTY0 = 4, BA_FULL00 = 8 /* [buf][kb] */, BA_EMPTY00 = 12, BT_FULL0 = 16, BT_EMPTY0 = 18, B_COUNT = 20 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug becomes a trap (reported by the host) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned* err) {
  uint32_t done = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 22)) {
      if (err) atomicExch(err, 0xdead0000u | (bar & 0xffffu));
      __trap();
    }
  }
}
// The same for warps that wait a long time by design (stagers, producers): back off so that the polling does not
// take issue slots from the warps that share the scheduler.
__device__ __forceinline__ void mbar_wait_patient(uint32_t bar, uint32_t parity, unsigned* err) {
  uint32_t done = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > 2) __nanosleep(spin < 64 ? 32 : 256);
    if (spin > (1u << 21)) {
      if (err) atomicExch(err, 0xdead0000u | (bar & 0xffffu));
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::f16 (fp16 in, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  if (accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc)
        : "memory");
  }
}
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);       // start address
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t umma_idesc(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// 32 consecutive accumulator columns of this thread's TMEM lane, squared and summed (packed f32x2 FMAs)
__device__ __forceinline__ float tmem_sumsq32(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  unsigned long long acc0 = 0ull, acc1 = 0ull;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    unsigned long long p0, p1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "r"(r[i]), "r"(r[i + 1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "r"(r[i + 2]), "r"(r[i + 3]));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc0) : "l"(p0));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc1) : "l"(p1));
  }
  uint32_t a, b, c, d;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(acc0));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(c), "=r"(d) : "l"(acc1));
  return (__uint_as_float(a) + __uint_as_float(b)) + (__uint_as_float(c) + __uint_as_float(d));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long A = *reinterpret_cast<unsigned long long*>(&a), B = *reinterpret_cast<unsigned long long*>(&b), C;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(C) : "l"(A), "l"(B));
  return *reinterpret_cast<float2*>(&C);
}
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// (x - mh) * s + nml on two lanes at once; returns the pair
__device__ __forceinline__ float2 centre_scale2(float2 x, float2 mh, float2 s2, float2 nml) {
  unsigned long long X = *reinterpret_cast<unsigned long long*>(&x), M = *reinterpret_cast<unsigned long long*>(&mh),
                     S = *reinterpret_cast<unsigned long long*>(&s2), L = *reinterpret_cast<unsigned long long*>(&nml), T, A;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(T) : "l"(X), "l"(M));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(A) : "l"(T), "l"(S), "l"(L));
  return *reinterpret_cast<float2*>(&A);
}

// One unit of work of the list mode: 128 consecutive entries of one cluster's candidate list.
struct ListItem {
  int k;          // cluster
  int count;      // valid entries (<= 128)
  long long base; // first entry in lrow
};
__device__ __forceinline__ ListItem list_item(int64_t it, int K, const int32_t* __restrict__ itoff,
                                              const long long* __restrict__ koff, const long long* __restrict__ kcnt) {
  int lo = 0, hi = K - 1;  // last k with itoff[k] <= it
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((int64_t)__ldg(itoff + mid) <= it) lo = mid;
    else hi = mid - 1;
  }
  ListItem r;
  r.k = lo;
  const long long first = (long long)(it - __ldg(itoff + lo)) * kTM;
  const long long left = __ldg(kcnt + lo) - first;
  r.count = (int)(left < kTM ? left : kTM);
  r.base = __ldg(koff + lo) + first;
  return r;
}

// kList == false: every 128-row tile of X against all K clusters, row soft-max in the epilogue (q, log Z).
// kList == true : the work items are 128-entry chunks of per-cluster row lists (lrow; item -> cluster through
//                 itoff); the kernel writes the exact logit of every (row, cluster) pair into q[row][cluster]
//                 and leaves the soft-max to estep_finalize_kernel.
// DIM == 64: the same kernel on 64-dimensional rows.  The operand blob keeps its 128-dimensional layout with
// R_k in the upper-left 64 x 64 corner and zeros elsewhere, so K block 1 (input dimensions 64..127) contributes
// nothing and is skipped: no builders for it, no MMAs, N = 64 - 16c output columns for chunk c, 64 accumulator
// columns in the epilogue, and only the 17 KB of the blob that are read travel to shared memory.
template <bool kList, int DIM>
__global__ void __launch_bounds__(kThreadsTc, 1)
estep_tc128_kernel(const float* __restrict__ X, int64_t N, const int32_t* __restrict__ gid, int K,
                   const uint8_t* __restrict__ blob, const float* __restrict__ ascale,
                   const float* __restrict__ inv_t2, const float* __restrict__ chat, const float* __restrict__ lw,
                   const uint8_t* __restrict__ act, float* __restrict__ q, int64_t ldq, double* __restrict__ Fz,
                   unsigned* __restrict__ err, const int32_t* __restrict__ lrow, const int4* __restrict__ items,
                   int64_t nitems, const long long* __restrict__ nitems_dev, const unsigned* __restrict__ skip) {
  extern __shared__ unsigned char smem_dyn[];
  if (skip != nullptr && *skip != 0u) return;
  if (kList && nitems_dev != nullptr) nitems = (int64_t)*nitems_dev;  // planned on the device (list_plan)
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sgen = smem_dyn + (sbase - smem_u32(smem_dyn));
  const uint32_t sB = sbase, sBar = sbase + kOffBar;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + kOffBar + 8 * B_COUNT);
  auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = kList ? nitems : (N + kTM - 1) / kTM;
  // list mode keeps two operand stages and gives the rest of shared memory to the builders' gather buffers
  constexpr uint32_t NS = kList ? 2u : (uint32_t)kStages;
  // dense mode: tiles strided over the CTAs.  List mode: a contiguous range of items per CTA -- the items are sorted by
  // cluster, so consecutive items mostly share their operand blob, which is then loaded once per run of equal
  // clusters instead of once per item (50 KB from L2 each: it bounded the pass).
  const int64_t tbeg = kList ? (nitems * (int64_t)blockIdx.x) / gridDim.x : (int64_t)blockIdx.x;
  const int64_t tend = kList ? (nitems * ((int64_t)blockIdx.x + 1)) / gridDim.x : ntiles;
  const int64_t tstep = kList ? 1 : (int64_t)gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(BB_FULL0 + i), 1);
      mbar_init(bar(BB_EMPTY0 + i), DIM == 64 ? 5 : 9);  // MMA commit + the builder warps (4 per K block)
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(BA_FULL00 + i), 4);  // 4 lane quadrants
      mbar_init(bar(BA_EMPTY00 + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(BT_FULL0 + i), 1);
      mbar_init(bar(BT_EMPTY0 + i), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register budget per role (64K registers / SM): the control warps and the epilogue
  // hand registers to the operand builders, which keep 64 fp32 of X per thread.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    // ------------------------------------------------------------ producer --
    if (warp == 0 && lane == 0) {
      uint32_t cnt = 0;
      int kprev = -1;
      for (int64_t tile = tbeg; tile < tend; tile += tstep) {
        int k0 = 0, k1 = K;
        if constexpr (kList) {
          k0 = __ldg(items + tile).x;
          k1 = k0 + 1;
          if (k0 == kprev) continue;  // the blob of this cluster is already on its way or resident
          kprev = k0;
        }
        for (int k = k0; k < k1; ++k, ++cnt) {
          const uint32_t st = cnt % NS, ph = (cnt / NS) & 1;
          mbar_wait(bar(BB_EMPTY0 + st), ph ^ 1, err);
          const uint8_t* src = blob + (size_t)k * kBBlob;
          if constexpr (DIM == 128) {
            mbar_expect_tx(bar(BB_FULL0 + st), kBBlob);
            bulk_g2s(sB + st * kBBlob, src, kBBlob, bar(BB_FULL0 + st));
          } else {
            // rows 0..63 of K block 0 (hi, lo) and the mean vectors
            mbar_expect_tx(bar(BB_FULL0 + st), 8192u + 8192u + 1024u);
            bulk_g2s(sB + st * kBBlob, src, 8192u, bar(BB_FULL0 + st));
            bulk_g2s(sB + st * kBBlob + kAPart, src + kAPart, 8192u, bar(BB_FULL0 + st));
            bulk_g2s(sB + st * kBBlob + kOffMean, src + kOffMean, 1024u, bar(BB_FULL0 + st));
          }
        }
      }
    }
    // ---------------------------------------------------------- MMA issuer --
    // The whole warp runs the loop so that addresses and descriptors live in uniform
    // registers; only the tcgen05 instructions themselves are issued by one elected lane.
    if (warp == 1) {
      uint32_t cnt = 0;   // (tile, cluster) items processed so far: A and accumulator stages
      uint32_t bidx = 0;  // operand blobs used so far: B stages (dense mode: one per item)
      int kprev = -1;
      for (int64_t tile = tbeg; tile < tend; tile += tstep) {
        const int nk = kList ? 1 : K;
        bool newk = true, lastk = true;
        if constexpr (kList) {
          const int kc = __ldg(items + tile).x;
          newk = kc != kprev;
          kprev = kc;
          lastk = tile + 1 >= tend || __ldg(items + tile + 1).x != kc;
        }
        for (int k = 0; k < nk; ++k, ++cnt) {
          if (newk) ++bidx;
          const uint32_t bs = (bidx - 1) % NS, bph = ((bidx - 1) / NS) & 1;
          const uint32_t st = cnt & 1, ph = (cnt >> 1) & 1;
          mbar_wait(bar(BT_EMPTY0 + st), ph ^ 1, err);
          if (newk) mbar_wait(bar(BB_FULL0 + bs), bph, err);
          const uint32_t sBk = sB + bs * kBBlob;
          const uint32_t a_hi0 = tmem_base + 256 + st * 128, a_lo0 = a_hi0 + 64;
          const uint32_t d0 = tmem_base + st * 128;
          // Order of the 24 accumulating MMAs: the tensor core truncates when it adds into the fp32 accumulator, a
          // bias of about half an ulp of the running sum per addition that shrinks |y| systematically (and, unlike
          // rounding noise, pushes q the same way in every VB iteration: 1.7e-5 .. 3e-5 on q after three iterations
          // of a soft 64-cluster fit with the products interleaved chunk by chunk, against 3e-6 for IEEE fp32,
          // profiles/diag_soft_r02.log).  The cross products hi*lo and lo*hi are 2^-11 of the result: they go first,
          // while the accumulator is still small and their truncations cost nothing; the eight hi*hi products
          // follow, so only eight additions happen at full magnitude.
          auto chunk_args = [&](int kb, int c4, uint32_t& dcol, uint32_t& acol, uint64_t& off, uint32_t& id) {
            const uint32_t c = 4 * kb + c4, nc = DIM - 16 * c;
            // first needed row of the stored block and the 16 fp16 along K inside the 128-byte row
            off = (uint64_t)(((16 * c - 64 * kb) * 128 + 32 * c4) >> 4);
            id = umma_idesc(nc);
            dcol = d0 + 16 * c;
            acol = 8 * c;
          };
          uint64_t dbh[2], dbl[2];
#pragma unroll
          for (int kb = 0; kb < DIM / 64; ++kb) {
            mbar_wait(bar(BA_FULL00 + 2 * st + kb), ph, err);
            tc_fence_after();
            const uint32_t b_hi = sBk + (kb == 0 ? 0u : 2 * kAPart);
            const uint32_t b_lo = b_hi + (kb == 0 ? kAPart : kAPart / 2);
            dbh[kb] = umma_desc(b_hi);
            dbl[kb] = umma_desc(b_lo);
            if (elect_one()) {
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t dcol, acol, id;
                uint64_t off;
                chunk_args(kb, c4, dcol, acol, off, id);
                tc_mma_f16_ts(dcol, a_hi0 + acol, dbl[kb] + off, id, (kb | c4) ? 1u : 0u);
                tc_mma_f16_ts(dcol, a_lo0 + acol, dbh[kb] + off, id, 1u);
              }
            }
            __syncwarp();
          }
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < DIM / 64; ++kb) {
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t dcol, acol, id;
                uint64_t off;
                chunk_args(kb, c4, dcol, acol, off, id);
                tc_mma_f16_ts(dcol, a_hi0 + acol, dbh[kb] + off, id, 1u);
              }
              tc_commit(bar(BA_EMPTY00 + 2 * st + kb));
            }
          }
          __syncwarp();
          if (elect_one()) {
            if (lastk) tc_commit(bar(BB_EMPTY0 + bs));
            tc_commit(bar(BT_FULL0 + st));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // ------------------------------------------------------------- epilogue --
    const int ew = warp - 4;
    double fz = 0;
    uint32_t cnt = 0;
    // list mode: (cluster, row) of this thread for the next item are fetched one item ahead
    int nk = 0;
    int64_t nn = -1;
    auto fetch = [&](int64_t it) {
      nk = 0;
      nn = -1;
      if (it < tend) {
        const int4 item = __ldg(items + it);
        nk = item.x;
        if (ew * 32 + lane < item.y)
          nn = (int64_t)__ldg(lrow + (((long long)(unsigned)item.z) | ((long long)item.w << 32)) + ew * 32 + lane);
      }
    };
    if constexpr (kList) fetch(tbeg);
    for (int64_t tile = tbeg; tile < tend; tile += tstep) {
      int64_t n = tile * kTM + ew * 32 + lane;
      bool valid = n < N;
      int k0 = 0, k1 = K;
      if constexpr (kList) {
        k0 = nk;
        k1 = k0 + 1;
        valid = nn >= 0;
        n = valid ? nn : 0;
        fetch(tile + tstep);
      }
      const int g = (gid != nullptr && valid) ? gid[n] : 0;
      const float* lwg = lw + (size_t)g * K;
      const uint8_t* actg = act != nullptr ? act + (size_t)g * K : nullptr;
      float* qrow = q + (valid ? n : 0) * ldq;
      float mx = -INFINITY, se = 0.f;  // running max / sum of exp for the row soft-max
      for (int k = k0; k < k1; ++k, ++cnt) {
        const uint32_t st = cnt & 1, ph = (cnt >> 1) & 1;
        mbar_wait(bar(BT_FULL0 + st), ph, err);
        tc_fence_after();
        float s = 0.f;
#pragma unroll
        for (int cc = 0; cc < DIM / 32; ++cc) s += tmem_sumsq32(tmem_base + ((uint32_t)(ew * 32) << 16) + st * 128 + 32 * cc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BT_EMPTY0 + st));
        float l = chat[k] + lwg[k] - 0.5f * inv_t2[k] * s;
        if (actg != nullptr && !actg[k]) l = -INFINITY;
        if (valid) qrow[k] = l;
        if constexpr (kList) continue;
        if (l > mx) {
          se = se * expf(mx - l) + 1.f;
          mx = l;
        } else if (l > -INFINITY) {
          se += expf(l - mx);
        }
      }
      if (!kList && valid) {
        // q = exp(logit - logZ) from the logits parked in the q row (L2-resident)
        const float lz = logf(se) + mx;
        float4* q4 = reinterpret_cast<float4*>(qrow);
        int k = 0;
        for (; k + 4 <= K; k += 4) {
          float4 v = q4[k >> 2];
          v.x = expf(v.x - lz); v.y = expf(v.y - lz); v.z = expf(v.z - lz); v.w = expf(v.w - lz);
          q4[k >> 2] = v;
        }
        for (; k < K; ++k) qrow[k] = expf(qrow[k] - lz);
        fz += (double)lz;
      }
    }
    if constexpr (!kList) {
      for (int o = 16; o > 0; o >>= 1) fz += __shfl_xor_sync(0xffffffffu, fz, o);
      if (lane == 0) atomicAdd(Fz, fz);
    }
  } else if (DIM == 64 && warp >= 12) {
    // no K block 1 in 64 dimensions: these builder warps have nothing to do
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    // ------------------------------------------------------ operand builders --
    // thread <-> one point (TMEM lane); its 64 dims of K block kb stay in registers for the whole tile
    const int quad = warp & 3, kb = (warp - 8) >> 2;
    const int row = 32 * quad + lane;
    uint32_t cnt = 0, bidx = 0;
    int kprev = -1;
    // List mode: the work items and the row indices travel ahead of the rows themselves, all with cp.async into
    // the padding bytes of this warp's gather buffer, so that no lane ever waits on a dependent load:
    //   while item t is built:  rows of item t+1, row indices of item t+2 and item t+3 are in flight.
    // (A plain prefetch into registers stalls the warp at the first instruction that touches the loaded value --
    // the address arithmetic of the next, dependent load -- for two global latencies per item.)
    const uint32_t tbs_pad = sbase + kOffTrans + (uint32_t)(warp - 8) * kTransWarp + 256u;
    auto item_slot = [&](int64_t t) { return tbs_pad + (uint32_t)(t % 3) * kTransRow; };
    auto lrow_slot = [&](int64_t t) {
      return tbs_pad + (uint32_t)(8 + 8 * (int)(t & 1) + (lane >> 2)) * kTransRow + 4u * (uint32_t)(lane & 3);
    };
    auto item_base = [](const int4& it) { return ((long long)(unsigned)it.z) | ((long long)it.w << 32); };
    int4 it0 = make_int4(-1, 0, 0, 0), it1 = it0;  // items t and t+1 (cluster -1: no such item)
    // Gathered rows (list mode): a thread reading its own row makes every load instruction touch 32 cache lines
    // (the L1 data pipe ran at 74 % and bounded the kernel, profiles/ncu_r01_refine_v4_raw.csv).  Instead the warp
    // copies two rows per instruction (16 lanes x 16 B each) into its private shared-memory buffer with cp.async, one
    // item ahead, so the gather latency hides behind the operand build of the current item.
    auto gather_rows = [&](int64_t nrow) {
      const uint32_t tbs = sbase + kOffTrans + (uint32_t)(warp - 8) * kTransWarp;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int rr = 2 * j + (lane >> 4);
        const int64_t nr = __shfl_sync(0xffffffffu, nrow, rr);
        const bool ok = nr < N;
        const float* src = X + (ok ? nr : 0) * DIM + 64 * kb + 4 * (lane & 15);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tbs + (uint32_t)rr * kTransRow + 16u * (uint32_t)(lane & 15)),
                     "l"(src), "r"(ok ? 16u : 0u)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if constexpr (kList) {
      // prologue with plain loads: items t0, t0+1 in registers, item t0+2 and the row indices of t0+1 in their slots
      if (tbeg < tend) it0 = __ldg(items + tbeg);
      if (tbeg + 1 < tend) it1 = __ldg(items + tbeg + 1);
      const int64_t n0 = row < it0.y ? (int64_t)__ldg(lrow + item_base(it0) + row) : N;
      if (row < it1.y) {
        const int32_t r1 = __ldg(lrow + item_base(it1) + row);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(lrow_slot(tbeg + 1)), "r"(r1) : "memory");
      }
      if (lane == 0 && tbeg + 2 < tend) {
        const int4 i2 = __ldg(items + tbeg + 2);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(item_slot(tbeg + 2)), "r"(i2.x), "r"(i2.y), "r"(i2.z), "r"(i2.w) : "memory");
      }
      gather_rows(n0);
    }
    for (int64_t tile = tbeg; tile < tend; tile += tstep) {
      int64_t n = tile * kTM + row;
      int k0 = 0, k1 = K;
      bool newk = true, lastk = true;
      if constexpr (kList) {
        k0 = it0.x;
        k1 = k0 + 1;
        newk = k0 != kprev;
        kprev = k0;
        lastk = it1.x != k0;  // the next item of this CTA (if any) uses another operand blob
      }
      float4 x[16];
      if constexpr (kList) {
        // The rows of this item were requested while the previous item was being built (gather_rows below): wait
        // for them, then transpose through the warp's private buffer: lane r takes row r.
        unsigned char* tb = sgen + kOffTrans + (uint32_t)(warp - 8) * kTransWarp;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = *reinterpret_cast<const float4*>(tb + lane * kTransRow + 16 * j);
        // what arrived with the rows: the row index of item t+1 and item t+2
        int64_t n1 = N;
        if (row < it1.y) {
          int32_t r1;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r1) : "r"(lrow_slot(tile + 1)) : "memory");
          n1 = r1;
        }
        int4 it2 = make_int4(-1, 0, 0, 0);
        if (tile + 2 < tend)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(it2.x), "=r"(it2.y), "=r"(it2.z), "=r"(it2.w) : "r"(item_slot(tile + 2)) : "memory");
        __syncwarp();
        // next requests: item t+3, the row indices of item t+2, the rows of item t+1
        if (lane == 0 && tile + 3 < tend)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(item_slot(tile + 3)), "l"(items + tile + 3) : "memory");
        if (row < it2.y)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(lrow_slot(tile + 2)), "l"(lrow + item_base(it2) + row) : "memory");
        gather_rows(n1);
        it0 = it1;
        it1 = it2;
      } else if (n < N) {
        const float4* src = reinterpret_cast<const float4*>(X + n * DIM + 64 * kb);
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __ldg(src + j);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int k = k0; k < k1; ++k, ++cnt) {
        if (newk) ++bidx;
        const uint32_t bs = (bidx - 1) % NS, bph = ((bidx - 1) / NS) & 1;
        const uint32_t st = cnt & 1, ph = (cnt >> 1) & 1;
        const float sc = ascale[k];
        const float2 s2 = make_float2(sc, sc);
        if (newk) mbar_wait(bar(BB_FULL0 + bs), bph, err);
        mbar_wait(bar(BA_EMPTY00 + 2 * st + kb), ph ^ 1, err);
        tc_fence_after();
        const float4* mh4 = reinterpret_cast<const float4*>(sgen + bs * kBBlob + kOffMean + 256 * kb);
        const float4* nl4 = reinterpret_cast<const float4*>(sgen + bs * kBBlob + kOffMean + 512 + 256 * kb);
        const uint32_t tA = tmem_base + ((uint32_t)(32 * quad) << 16) + 256 + st * 128 + 32 * kb;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t rh[16], rl[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 xv = x[8 * h + j], mh = mh4[8 * h + j], nl = nl4[8 * h + j];
            const float2 a01 = centre_scale2(make_float2(xv.x, xv.y), make_float2(mh.x, mh.y), s2, make_float2(nl.x, nl.y));
            const float2 a23 = centre_scale2(make_float2(xv.z, xv.w), make_float2(mh.z, mh.w), s2, make_float2(nl.z, nl.w));
            const uint32_t h01 = pack_f16x2_sat(a01.x, a01.y), h23 = pack_f16x2_sat(a23.x, a23.y);
            const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
            const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
            rh[2 * j] = h01;
            rh[2 * j + 1] = h23;
            const float2 r01 = sub2(a01, f01), r23 = sub2(a23, f23);
            rl[2 * j] = pack_f16x2_sat(r01.x, r01.y);
            rl[2 * j + 1] = pack_f16x2_sat(r23.x, r23.y);
          }
          tmem_st16(tA + 16 * h, rh);
          tmem_st16(tA + 64 + 16 * h, rl);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(BA_FULL00 + 2 * st + kb));
          if (lastk) mbar_arrive(bar(BB_EMPTY0 + bs));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}


// ===========================================================================
// sstat_tc128_kernel: centred scatter over the per-cluster lists of non-zero
// responsibilities, D == 128:
//     S_k += sum_r (x_r - c_k) q_r (x_r - c_k)^T ,   xs_k += sum_r q_r (x_r - c_k)
// as a 128 x 128 x (rows) GEMM on tcgen05 with both operands equal to
// sqrt(q) s (X - c)^T [dims x rows]: A lives in TMEM, B (rows contiguous) in swizzled
// shared memory, split into fp16 (hi, lo) with the three significant products
// accumulated in fp32 TMEM accumulators.  s is a power of two that cannot saturate fp16 for any
// row of the data set (engine: 2^14 / max|x - c|).  A builder thread owns one
// dimension (= TMEM lane = B row) and 64 of the 128 rows of a tile; rows are
// gathered from X through the (row, q) lists.
//
// A work item is a chunk of at most chunk_rows list entries of one cluster: the
// accumulators cover that many rows (<= 32 tensor-core additions per element at
// 512) before they are added, in fp64, to the global statistics -- the tensor core
// truncates when it adds into fp32, and the bias of a long chain would show in the
// covariances.  One persistent CTA per SM walks the items; four flush warps move a
// finished accumulator to the fp64 statistics while the builders and the MMA warp
// already work on the next item (the first version launched one CTA per item and
// spent more than half of its time in prologue, first-gather latency and the
// flush, profiles/ncu_r01_sstat_v4_raw.csv).
//   warp 0      list loader: rows and sqrt(q) of the next 128 entries into shared memory
//   warp 1      MMA issue
//   warp 2      TMEM allocation
//   warps 4-7   flush: tcgen05.ld of both accumulators, fp64 atomics
//   warps 8-23  builders: thread = (dimension, quarter of the tile's rows).  The first version had 8 builder warps
//               with 64 rows per thread: two warps per scheduler could not cover the latency of the gathered loads
//               (long-scoreboard stalls 2.0 per issue, issue slots 55 % busy, tensor pipe 30 %,
//               profiles/ncu_r01_sstat_v6_raw.csv); 16 warps with 32 rows each double the loads in flight per
//               scheduler, and the centring / weighting / hi-lo split runs on packed f32x2 instructions.
// ===========================================================================
constexpr int kScatterThreads = 768;
constexpr uint32_t kSB_Stage = 65536; // per stage: K block 0 (hi 16K, lo 16K), K block 1 (hi, lo)
constexpr uint32_t kSOffList = 2 * kSB_Stage;
constexpr int kSListStages = 8;                               // list tiles the loader (and its L2 prefetch) runs ahead
constexpr uint32_t kSOffPre = kSOffList + kSListStages * 1024;  // item prefix per cluster: (kTcCoarseMaxK + 1) ints
constexpr uint32_t kSOffBar = kSOffPre + 4 * (kTcCoarseMaxK + 8);
constexpr uint32_t kSOffTri = kSOffBar + 256;                 // fp64 running sums of the current cluster, lower triangle
constexpr uint32_t kSTriBytes = 128 * 129 / 2 * 8;            // 66048
constexpr uint32_t kSSmemBytes = kSOffTri + kSTriBytes + 1024;
static_assert(kSOffTri % 8 == 0 && kSSmemBytes <= 232448, "S pass shared memory");
enum {
  SL_FULL0 = 0, SL_EMPTY0 = 8, SAB_FULL00 = 16 /* [stage][h] */, SAB_EMPTY00 = 20, SACC_FULL = 24, SACC_EMPTY = 25,
  SB_COUNT = 26
};
static_assert(8 * SB_COUNT + 8 <= 256 && kSListStages == 8, "barrier block of the S pass");

struct ScatterItem {
  int k, ntile;
  long long l0, l1, base;
};

// DIM == 64: dimensions 64..127 do not exist: their builder and flush warps idle, the MMAs run with N = 64 (the A
// rows 64..127 in TMEM are never written and the accumulator rows they produce are never read).
template <int DIM>
__global__ void __launch_bounds__(kScatterThreads, 1)
sstat_tc128_kernel(const float* __restrict__ X, const int32_t* __restrict__ lrow, const float* __restrict__ lq,
                   const long long* __restrict__ koff, const long long* __restrict__ kcnt, int K,
                   const float* __restrict__ cen, float scale, int chunk_rows, double* __restrict__ xs,
                   double* __restrict__ S, unsigned* __restrict__ err, const float* __restrict__ scale_dev,
                   const unsigned* __restrict__ skip, int pf) {
  extern __shared__ unsigned char smem_dyn[];
  if (skip != nullptr && *skip != 0u) return;
  if (scale_dev != nullptr) scale = *scale_dev;  // chosen by the device M step from the centres it produced
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sgen = smem_dyn + (sbase - smem_u32(smem_dyn));
  const uint32_t sBar = sbase + kSOffBar;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + kSOffBar + 8 * SB_COUNT);
  int* pre = reinterpret_cast<int*>(sgen + kSOffPre);  // pre[k] = items of the clusters before k
  auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSListStages; ++i) {
      mbar_init(bar(SL_FULL0 + i), 1);
      mbar_init(bar(SL_EMPTY0 + i), DIM / 8);   // builder warps: 4 quarters of the rows x DIM / 32 lane quadrants
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(SAB_FULL00 + i), DIM / 16);
      mbar_init(bar(SAB_EMPTY00 + i), 1);
    }
    mbar_init(bar(SACC_FULL), 1);
    mbar_init(bar(SACC_EMPTY), DIM / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int acc = 0;
    for (int k = 0; k < K; ++k) {
      pre[k] = acc;
      acc += (int)((kcnt[k] + chunk_rows - 1) / chunk_rows);
    }
    pre[K] = acc;
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nitems = pre[K];
  // A contiguous range of items per CTA: the items are ordered by cluster, so a CTA meets two or three clusters in a
  // pass and can keep the running sums of the current one on chip (the first version strided the items over the CTAs
  // and added every 512-row accumulator to the global statistics: 16 K fp64 atomics per item, 1.6 G per pass at
  // N = 50 M -- the pass ran at the L2's atomic rate, 145 G/s, whatever the builders did).
  const int it_beg = (int)(((long long)nitems * blockIdx.x) / gridDim.x);
  const int it_end = (int)(((long long)nitems * (blockIdx.x + 1)) / gridDim.x);
  // item -> (cluster, list range): the last cluster whose prefix is <= it
  auto item_of = [&](int it) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pre[mid] <= it) lo = mid;
      else hi = mid - 1;
    }
    ScatterItem r;
    r.k = lo;
    const long long cnt = kcnt[lo];
    r.l0 = (long long)(it - pre[lo]) * chunk_rows;
    r.l1 = r.l0 + chunk_rows < cnt ? r.l0 + chunk_rows : cnt;
    r.base = koff[lo];
    r.ntile = (int)((r.l1 - r.l0 + 127) / 128);
    return r;
  };

  if (warp < 4) {
    // Register budget: the CTA owns 768 x 80 registers and setmaxnreg can only move registers inside that pool
    // (24 warps x 80 = 1920 per lane): control 4 x 48, flush 4 x 104, builders 16 x 80 = 1888.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      // ---- list loader: (row, q) of the next 128 list entries into shared memory ----
      uint32_t tc = 0;
      for (int it = it_beg; it < it_end; ++it) {
        const ScatterItem w = item_of(it);
        for (int t = 0; t < w.ntile; ++t, ++tc) {
          // The loader runs up to kSListStages tiles ahead of the builders and asks the L2 for the rows of every
          // tile it lists: the builders' gathers, issued several tile times later, find them there.  (With a
          // two-deep list ring and no prefetch the pass was bound by the latency chain list load -> gather from
          // DRAM -> build: 4.5 us per tile against 1 us of tensor work, whatever the number of builder warps and
          // however few atomics the flush issued.)
          const uint32_t st = tc & (kSListStages - 1);
          mbar_wait_patient(bar(SL_EMPTY0 + st), ((tc / kSListStages) & 1) ^ 1, err);
          // [128 rows (int32)][128 sqrt(q) (fp32)]; entries past the end of the list read row 0 with weight 0
          int* drow = reinterpret_cast<int*>(sgen + kSOffList + st * 1024);
          float* dw = reinterpret_cast<float*>(drow + 128);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const long long l = w.l0 + (long long)t * 128 + e * 32 + lane;
            int row = 0;
            float wq = 0.f;
            if (l < w.l1) {
              row = lrow[w.base + l];
              wq = sqrtf(lq[w.base + l]);  // the builders work with sqrt(q), see below
              if (pf == 2) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(X + (size_t)row * DIM), "n"(DIM * 4) : "memory");
            }
            drow[e * 32 + lane] = row;
            dw[e * 32 + lane] = wq;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(SL_FULL0 + st));
        }
      }
    } else if (warp == 3 && pf == 1) {
      // ---- L2 prefetch (optional): the rows of the tile the loader has just listed, all 128-byte lines of a row
      // back to back, so that the builders' gathers -- one 128-byte piece of a row per warp instruction, the four
      // pieces from four different warps at four different times -- find the row in L2 ----
      uint32_t tc = 0;
      for (int it = it_beg; it < it_end; ++it) {
        const ScatterItem w = item_of(it);
        for (int t = 0; t < w.ntile; ++t, ++tc) {
          const uint32_t st = tc & (kSListStages - 1);
          mbar_wait_patient(bar(SL_FULL0 + st), (tc / kSListStages) & 1, err);
          const int* rows = reinterpret_cast<const int*>(sgen + kSOffList + st * 1024);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const char* p = reinterpret_cast<const char*>(X + (size_t)rows[e * 32 + lane] * DIM);
#pragma unroll
            for (int b = 0; b < DIM * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b) : "memory");
          }
        }
      }
    } else if (warp == 1) {
      // ---- MMA issuer ----
      uint32_t tc = 0, ic = 0;
      for (int it = it_beg; it < it_end; ++it, ++ic) {
        const ScatterItem w = item_of(it);
        // the flush warps must have read the previous item's accumulators
        mbar_wait(bar(SACC_EMPTY), (ic & 1) ^ 1, err);
        tc_fence_after();
        for (int t = 0; t < w.ntile; ++t, ++tc) {
          const uint32_t st = tc & 1;
          const uint32_t ph = (tc >> 1) & 1;
          const uint32_t a_hi0 = tmem_base + 256 + st * 128, a_lo0 = a_hi0 + 64;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar(SAB_FULL00 + 2 * st + h), ph, err);
            tc_fence_after();
            const uint32_t b_hi = sbase + st * kSB_Stage + h * 32768, b_lo = b_hi + 16384;
            const uint64_t dbh0 = umma_desc(b_hi), dbl0 = umma_desc(b_lo);
            const uint32_t id = umma_idesc(DIM);
            if (elect_one()) {
              // The hi*hi products and the 2^-11 smaller cross terms go to separate accumulators: the tensor core
              // truncates when it adds into the fp32 accumulator, and the bias of a chain of n additions (~ n/2 ulp
              // of the running sum) must stay ~1e-7 relative for the statistics, so the big chain is kept short.
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const uint64_t off = (uint64_t)((32 * c4) >> 4);
                const uint32_t acol = 32 * h + 8 * c4;
                const uint32_t first = (t == 0 && h == 0 && c4 == 0) ? 0u : 1u;
                tc_mma_f16_ts(tmem_base, a_hi0 + acol, dbh0 + off, id, first);
                tc_mma_f16_ts(tmem_base + 128, a_hi0 + acol, dbl0 + off, id, first);
                tc_mma_f16_ts(tmem_base + 128, a_lo0 + acol, dbh0 + off, id, 1u);
              }
              tc_commit(bar(SAB_EMPTY00 + 2 * st + h));
            }
            __syncwarp();
          }
        }
        if (elect_one()) tc_commit(bar(SACC_FULL));
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---- flush: accumulators -> fp64 running sums of the current cluster in shared memory (lower triangle: lane i
    // owns row i, columns j <= i); the sums go to the global statistics, both triangles, when the CTA's items move on
    // to another cluster ----
    const int quad = warp & 3;
    const int i = 32 * quad + lane;
    const double inv_s = 1.0 / (double)scale, inv_s2 = inv_s * inv_s;
    double* tri = reinterpret_cast<double*>(sgen + kSOffTri) + (size_t)i * (i + 1) / 2;
    const bool mine = i < DIM;
    if (mine)
      for (int j = 0; j <= i; ++j) tri[j] = 0.0;
    auto dump = [&](int k) {
      double* Sk = S + (size_t)k * DIM * DIM;
      for (int j = 0; j <= i; ++j) {
        const double v = tri[j] * inv_s2;
        tri[j] = 0.0;
        if (v != 0.0) {
          atomicAdd(&Sk[(size_t)i * DIM + j], v);
          if (j != i) atomicAdd(&Sk[(size_t)j * DIM + i], v);
        }
      }
    };
    uint32_t ic = 0;
    int kcur = -1;
    for (int it = it_beg; it < (mine ? it_end : it_beg); ++it, ++ic) {
      const ScatterItem w = item_of(it);
      if (w.k != kcur) {
        if (kcur >= 0) dump(kcur);
        kcur = w.k;
      }
      mbar_wait(bar(SACC_FULL), ic & 1, err);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc <= 2 * quad + 1; ++cc) {  // 16-column blocks that hold a j <= i for some lane of this warp
        uint32_t r[16], r2[16];
        tmem_ld16(tmem_base + ((uint32_t)(32 * quad) << 16) + 16 * cc, r);
        tmem_ld16(tmem_base + ((uint32_t)(32 * quad) << 16) + 128 + 16 * cc, r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cc == 2 * quad + 1) {
          // everything this warp needs has been read: the next item may overwrite the accumulators
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(SACC_EMPTY));
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int j = 16 * cc + jj;
          if (j <= i) tri[j] += (double)__uint_as_float(r[jj]) + (double)__uint_as_float(r2[jj]);
        }
      }
    }
    if (kcur >= 0) dump(kcur);
  } else {
    // ---- builders: thread = dimension i (TMEM lane, B row), rows [32 hg, 32 hg + 32) of the tile ----
    // The gathers run one half step ahead of the arithmetic: the 32 rows of a thread are two batches of 16
    // (8 float2 registers each); as soon as a batch has been consumed its registers are reloaded with the same
    // batch of the NEXT tile of this CTA (the list ring is kSListStages deep, so its rows are known), so the loads
    // of a batch are in flight during the arithmetic of the other batch, the operand stores and the barrier
    // hand-offs.  (Loading all 32 rows and then using them exposed the whole gather latency once per tile: the
    // first use of a loaded value and the load issue queue held half of all stall samples,
    // profiles/ncu_r02_sstat_v1_*.)
    const int quad = warp & 3, hg = (warp - 8) >> 2, h = hg >> 1, g = hg & 1;
    const int i = 32 * quad + lane;
    const double inv_s = 1.0 / (double)scale;
    const float2 sc2 = make_float2(scale, scale);
    const bool active = 32 * quad < DIM;
    // tiles of this CTA in order: tile counter tc -> list stage; the items only matter for the centre and the sums
    float2 a[16];
    auto load_batch = [&](uint32_t tcn, int b) {
      const uint32_t lst = tcn & (kSListStages - 1);
      const int4* rows4 = reinterpret_cast<const int4*>(sgen + kSOffList + lst * 1024) + 8 * hg + 4 * b;
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        const int4 rr = rows4[r4];
        a[8 * b + 2 * r4].x = __ldg(X + (size_t)rr.x * DIM + i);
        a[8 * b + 2 * r4].y = __ldg(X + (size_t)rr.y * DIM + i);
        a[8 * b + 2 * r4 + 1].x = __ldg(X + (size_t)rr.z * DIM + i);
        a[8 * b + 2 * r4 + 1].y = __ldg(X + (size_t)rr.w * DIM + i);
      }
    };
    // total tiles of this CTA (for the look-ahead)
    uint32_t ntiles_cta = 0;
    if (active)
      for (int it = it_beg; it < it_end; ++it) ntiles_cta += (uint32_t)item_of(it).ntile;
    uint32_t tc = 0;
    if (active && ntiles_cta > 0) {
      mbar_wait(bar(SL_FULL0 + 0), 0, err);
      load_batch(0, 0);
      load_batch(0, 1);
    }
    for (int it = it_beg; it < (active ? it_end : it_beg); ++it) {
      const ScatterItem w = item_of(it);
      const float ncs = -cen[(size_t)w.k * DIM + i] * scale;
      const float2 ncs2 = make_float2(ncs, ncs);
      double xs64 = 0.0;
      for (int t = 0; t < w.ntile; ++t, ++tc) {
        const uint32_t st = tc & 1;
        const uint32_t ph = (tc >> 1) & 1;
        const uint32_t lst = tc & (kSListStages - 1);
        const bool more = tc + 1 < ntiles_cta;
        const float2* wq2 = reinterpret_cast<const float2*>(sgen + kSOffList + lst * 1024 + 512) + 16 * hg;
        unsigned long long xs2 = 0ull;  // packed pair of fp32 partial sums of q (x - c) s
        mbar_wait(bar(SAB_EMPTY00 + 2 * st + h), ph ^ 1, err);
        tc_fence_after();
        const uint32_t tA = tmem_base + ((uint32_t)(32 * quad) << 16) + 256 + st * 128 + 32 * h + 16 * g;
        const uint32_t bRow = sbase + st * kSB_Stage + h * 32768 + (uint32_t)i * 128u;
        uint32_t ah[8], al[8];   // eight packed columns (16 rows) at a time: stored to TMEM after every second chunk
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t bh[4], bl[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            // S_k = sum_r (w a)(w a)^T with w = sqrt(q): both MMA operands are the same numbers (A in TMEM, B in
            // shared memory), so one scaling and one fp16 hi/lo split serve both.  Two rows per instruction.
            const int pr = 4 * c + p;
            const float2 wv = wq2[pr];
            unsigned long long X2 = *reinterpret_cast<const unsigned long long*>(&a[pr]);
            const unsigned long long S2 = *reinterpret_cast<const unsigned long long*>(&sc2);
            const unsigned long long N2 = *reinterpret_cast<const unsigned long long*>(&ncs2);
            const unsigned long long W2 = *reinterpret_cast<const unsigned long long*>(&wv);
            unsigned long long T2, V2, D2;
            asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(T2) : "l"(X2), "l"(S2), "l"(N2));
            asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(V2) : "l"(T2), "l"(W2));
            const float2 v = *reinterpret_cast<const float2*>(&V2);
            const uint32_t hh = pack_f16x2_sat(v.x, v.y);
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hh));
            const unsigned long long H2 = *reinterpret_cast<const unsigned long long*>(&hf);
            asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(D2) : "l"(V2), "l"(H2));
            const float2 dl = *reinterpret_cast<const float2*>(&D2);
            const uint32_t ll = pack_f16x2_sat(dl.x, dl.y);
            ah[pr & 7] = hh;
            al[pr & 7] = ll;
            bh[p] = hh;
            bl[p] = ll;
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(xs2) : "l"(V2), "l"(W2));
          }
          const uint32_t off = bRow + ((((uint32_t)(4 * g + c)) ^ ((uint32_t)i & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(off), "r"(bh[0]), "r"(bh[1]), "r"(bh[2]), "r"(bh[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(off + 16384u), "r"(bl[0]), "r"(bl[1]), "r"(bl[2]), "r"(bl[3]) : "memory");
          if (c & 1) {
            tmem_st8(tA + 4 * (c - 1), ah);
            tmem_st8(tA + 64 + 4 * (c - 1), al);
          }
          if (c == 1 && more) {
            // batch 0 is consumed: reload it with the next tile's rows (its list stage must have been filled)
            mbar_wait(bar(SL_FULL0 + ((tc + 1) & (kSListStages - 1))), ((tc + 1) / kSListStages) & 1, err);
            load_batch(tc + 1, 0);
          }
          if (c == 3 && more) load_batch(tc + 1, 1);
        }
        {
          const float2 xp = *reinterpret_cast<const float2*>(&xs2);
          xs64 += (double)(xp.x + xp.y);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(SAB_FULL00 + 2 * st + h));
          mbar_arrive(bar(SL_EMPTY0 + lst));
        }
      }
      if (xs64 != 0.0) atomicAdd(&xs[(size_t)w.k * DIM + i], xs64 * inv_s);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ===========================================================================
// Two-level E step, level 1: estep_coarse_tc128_kernel (D == 128, fp32 engine).
//
// Every (row, cluster) distance is first computed with ONE fp16 product
// (hi * hi, no per-cluster operand build): A = fp16(s_g x) is converted once per
// 128-row tile and serves all K clusters, B_k is the hi half of the dense kernel's
// blob, and the centring -R_k m_k enters the accumulator through one extra K chunk
// (A = 2^P in three slots, B = the fp16 hi/mid/lo split of -s_g tau_k (R_k m_k)_i / 2^P).
// The result has a rigorous error bound:
//     | d~ - d | <= E_nk = Ek[k] * |x_n| + Ea[k],     d = | R_k (x_n - m_k) |
// so each logit is bracketed, LB <= logit <= UB.  A pair (n, k) is a *candidate*
// iff UB_nk >= max_j LB_nj - margin; only candidates can have q > e^-margin.  The
// kernel leaves UB in q and the candidates as a bit mask per row (cmask); level 2
// (the list mode of estep_tc128_kernel) recomputes the candidates at
// fp32-equivalent accuracy and estep_finalize_kernel forms q and log Z.
//
// One persistent CTA per SM works on groups of kCT = 3 tiles (384 rows) so that a
// cluster's operand, streamed from L2 by the TMA engine, is used by three MMAs
// chains: 16 warps =
//   warp 0      producer: cp.async.bulk of B_k (24 KB, 3-stage ring) and of the
//               aug blocks (16 KB per 4 clusters, 2-stage ring)
//   warps 1-3   MMA issue, one warp per tile slot: per (cluster, tile) item 1 aug +
//               8 triangular chunk MMAs (coarse_mma_issuer); warp 2 also allocates TMEM
//   warps 4-7   stagers: X -> fp16 staging in shared memory during the previous
//               group, staging -> TMEM A at the group boundary; they also turn the
//               parked UB rows of the finished group into the candidate marking
//   warps 8-15  two epilogue groups, one per accumulator: tcgen05.ld the 128
//               accumulator columns, release the accumulator, sum of squares,
//               bounds, park UB in the q row
// TMEM: accumulators [0,128) [128,256); A of tile slot s at 256 + 72 s (64 columns
// of data + 8 of the aug chunk).
// ===========================================================================
constexpr int kCT = 3;
constexpr int kCRows = kCT * kTM;                              // 384 rows per group
constexpr int kCStages = 3;
constexpr uint32_t kCBStage = 24576;                           // hi block 0 (16 KB) + hi block 1 (8 KB)
constexpr uint32_t kCOffAug = kCStages * kCBStage;             // 73728
constexpr uint32_t kCOffStageA = kCOffAug + 2 * kTcAugBlockBytes;  // 106496
constexpr uint32_t kCStageA = 32768;                           // one tile as fp16 [128][128], 16-byte chunks XOR (row & 15)
constexpr uint32_t kCOffPar = kCOffStageA + kCT * kCStageA;    // 204800: per-cluster floats [4][256]
constexpr uint32_t kCOffLb = kCOffPar + 4 * 256 * 4;           // 208896: [2 parities][2 groups][384] lower-bound maxima
constexpr uint32_t kCOffBar = kCOffLb + 2 * 2 * kCRows * 4;    // 215040
constexpr uint32_t kCSmemBytes = kCOffBar + 512 + 1024;
constexpr uint32_t kCAcol = 72;                                // TMEM columns of one A slot
enum {
  CB_FULL0 = 0, CB_EMPTY0 = 3, CG_FULL0 = 6, CG_EMPTY0 = 8, CA_READY0 = 10, CA_FREE0 = 13, CT_FULL0 = 16 /* 4 */,
  CL_FULL0 = 20, CL_FREE0 = 22, C_COUNT = 24
};

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}
// sum of squares of NV (128 or 64) fp32 register values: eight independent chains of packed FMAs, packed adds to
// fold them
template <int NV>
__device__ __forceinline__ float sumsq_regs(const uint32_t* r) {
  unsigned long long acc[8] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int i = 0; i < NV; i += 16) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      unsigned long long p;
      asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "r"(r[i + 2 * c]), "r"(r[i + 2 * c + 1]));
      asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc[c]) : "l"(p));
    }
  }
#pragma unroll
  for (int h = 4; h >= 1; h >>= 1) {
#pragma unroll
    for (int c = 0; c < h; ++c) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[c]) : "l"(acc[c + h]));
  }
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(acc[0]));
  return __uint_as_float(a) + __uint_as_float(b);
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------- MMA issuers of the level-1 kernel --
// Three issuing warps, one per tile slot S of the group; item (k, S) is the ic-th item of this CTA,
// ic = 3 (K g + k) + S, and uses accumulator ic & 1.  One issuer needs ~850 cycles of dependent uniform-datapath
// instructions per item (waits, descriptor arithmetic, nine MMAs, commits) against ~370 tensor cycles: a single
// issuer left the tensor pipe 57 % idle, two issuers 33 % (profiles/ncu_r01_coarse_v2_raw.csv, ..._v3_raw.csv).
// Shared-memory addresses derive from the kernel parameter sbase_hint and the loop counters only (TMEM base = 0: the
// CTA owns all 512 columns), which keeps the loop on the uniform datapath.
template <uint32_t S, int DIM>
__device__ __forceinline__ void coarse_mma_issuer(uint32_t sb, int K, int64_t ngroups, unsigned* err) {
  constexpr uint32_t a0 = 256u + kCAcol * S;
  const uint32_t barb = sb + kCOffBar;
  const uint32_t blo_base = umma_desc_lo(sb), glo_base = umma_desc_lo(sb + kCOffAug);
  uint32_t ic = S, gcnt = 0, acnt = 0;
  uint32_t bs = 0, bph = 0, as = 0;
  for (int64_t gi = blockIdx.x; gi < ngroups; gi += gridDim.x, ++gcnt) {
    mbar_wait(barb + 8u * (CA_READY0 + S), gcnt & 1, err);
    for (int k = 0; k < K; ++k, ic += 3) {
      if ((k & 3) == 0) {
        as = acnt & 1;
        mbar_wait(barb + 8u * (CG_FULL0 + as), (acnt >> 1) & 1, err);
        ++acnt;
      }
      mbar_wait(barb + 8u * (CB_FULL0 + bs), bph, err);
      const uint32_t lo0 = blo_base + bs * (kCBStage >> 4), lo1 = lo0 + (16384u >> 4);
      const uint32_t log_ = glo_base + as * (kTcAugBlockBytes >> 4) + 2u * (uint32_t)(k & 3);
      // accumulators: two of 128 columns, or -- 64 dimensions -- four of 64 columns, so that twice as many items are
      // in flight between issue and drain (at 112 tensor clocks per item the hand-off latency is what counts)
      constexpr uint32_t NACC = DIM == 64 ? 4u : 2u, LGA = DIM == 64 ? 2u : 1u, ACCW = DIM == 64 ? 64u : 128u;
      const uint32_t acc = ic & (NACC - 1u);
      const bool last = k == K - 1, aug_done = (k & 3) == 3 || last;
      if (elect_one()) {
        // One asm block: wait until the epilogue has drained the accumulator -- the previous ic >> 1 items that used
        // it, four arrivals each, counted in shared memory: the three issuers take turns on an accumulator, and a
        // phase-parity wait cannot tell "the item before the previous one is still being drained" from "the
        // previous one has been drained" --, then
        //   accumulator = -s_g tau_k R_k m_k (aug chunk, K = 16), followed by the eight triangular chunks (R_k is
        //   lower-triangular, so the K chunk of input dimensions [16c, 16c+16) only feeds output columns >= 16c:
        //   N = 128 - 16c), and the commit.
        // Instruction descriptors: D = f32, A = B = f16, K-major, M = 128, N as above.  The spin is bounded: a
        // protocol bug traps instead of hanging the device.
        if constexpr (DIM == 128) {
        asm volatile(
            "{\n\t"
            ".reg .pred p, q, r;\n\t"
            ".reg .b64 bd;\n\t"
            ".reg .b32 t, c;\n\t"
            "setp.ne.b32 p, 1, 0;\n\t"
            "mov.u32 c, 0;\n\t"
            "CW_WAIT_%=:\n\t"
            "add.u32 c, c, 1;\n\t"
            "setp.gt.u32 r, c, 67108864;\n\t"
            "@r trap;\n\t"
            "ld.acquire.cta.shared.u32 t, [%0];\n\t"
            "sub.u32 t, t, %1;\n\t"
            "setp.lt.s32 q, t, 0;\n\t"
            "@q bra CW_WAIT_%=;\n\t"
            "tcgen05.fence::after_thread_sync;\n\t"
            "mov.b64 bd, {%6, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2], [%3+64], bd, 0x8200010, !p;\n\t"
            "mov.b64 bd, {%4, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2], [%3], bd, 0x8200010, p;\n\t"
            "add.u32 t, %4, 0x82;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+16], [%3+8], bd, 0x81c0010, p;\n\t"
            "add.u32 t, %4, 0x104;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+32], [%3+16], bd, 0x8180010, p;\n\t"
            "add.u32 t, %4, 0x186;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+48], [%3+24], bd, 0x8140010, p;\n\t"
            "mov.b64 bd, {%5, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+64], [%3+32], bd, 0x8100010, p;\n\t"
            "add.u32 t, %5, 0x82;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+80], [%3+40], bd, 0x80c0010, p;\n\t"
            "add.u32 t, %5, 0x104;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+96], [%3+48], bd, 0x8080010, p;\n\t"
            "add.u32 t, %5, 0x186;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+112], [%3+56], bd, 0x8040010, p;\n\t"
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
            "}"
            ::"r"(barb + 8u * C_COUNT + 16u + 4u * acc), "r"(4u * (ic >> LGA)), "r"(ACCW * acc), "r"(a0), "r"(lo0), "r"(lo1),
              "r"(log_), "r"(kDescHi), "r"(barb + 8u * (CT_FULL0 + acc)), "r"(barb + 8u * (CB_EMPTY0 + bs))
            : "memory");
        } else {
          // 64 dimensions: the aug chunk and the four triangular chunks of K block 0, N = 64 - 16c
        asm volatile(
            "{\n\t"
            ".reg .pred p, q, r;\n\t"
            ".reg .b64 bd;\n\t"
            ".reg .b32 t, c;\n\t"
            "setp.ne.b32 p, 1, 0;\n\t"
            "mov.u32 c, 0;\n\t"
            "CW6_WAIT_%=:\n\t"
            "add.u32 c, c, 1;\n\t"
            "setp.gt.u32 r, c, 67108864;\n\t"
            "@r trap;\n\t"
            "ld.acquire.cta.shared.u32 t, [%0];\n\t"
            "sub.u32 t, t, %1;\n\t"
            "setp.lt.s32 q, t, 0;\n\t"
            "@q bra CW6_WAIT_%=;\n\t"
            "tcgen05.fence::after_thread_sync;\n\t"
            "mov.b64 bd, {%6, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2], [%3+64], bd, 0x8100010, !p;\n\t"
            "mov.b64 bd, {%4, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2], [%3], bd, 0x8100010, p;\n\t"
            "add.u32 t, %4, 0x82;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+16], [%3+8], bd, 0x80c0010, p;\n\t"
            "add.u32 t, %4, 0x104;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+32], [%3+16], bd, 0x8080010, p;\n\t"
            "add.u32 t, %4, 0x186;\n\t"
            "mov.b64 bd, {t, %7};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%2+48], [%3+24], bd, 0x8040010, p;\n\t"
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
            "}"
            ::"r"(barb + 8u * C_COUNT + 16u + 4u * acc), "r"(4u * (ic >> LGA)), "r"(ACCW * acc), "r"(a0), "r"(lo0), "r"(lo1),
              "r"(log_), "r"(kDescHi), "r"(barb + 8u * (CT_FULL0 + acc)), "r"(barb + 8u * (CB_EMPTY0 + bs))
            : "memory");
        }
        if (aug_done) tc_commit(barb + 8u * (CG_EMPTY0 + as));
        if (last) tc_commit(barb + 8u * (CA_FREE0 + S));
      }
      __syncwarp();
      if (++bs == kCStages) {
        bs = 0;
        bph ^= 1;
      }
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(kThreadsTc, 1)
estep_coarse_tc128_kernel(const float* __restrict__ X, const float* __restrict__ xnorm, int64_t N,
                          const int32_t* __restrict__ gid, int K, const uint8_t* __restrict__ blob,
                          const uint8_t* __restrict__ augblob, const float* __restrict__ cpar /* [4][K] */,
                          const float* __restrict__ lw, const uint8_t* __restrict__ act, float sg, uint32_t aug01,
                          uint32_t aug2, float margin, float* __restrict__ q, int64_t ldq,
                          uint32_t* __restrict__ cmask, uint32_t sbase_hint, unsigned* __restrict__ err,
                          const unsigned* __restrict__ augh_dev, const unsigned* __restrict__ skip) {
  extern __shared__ unsigned char smem_dyn[];
  if (augh_dev != nullptr) {  // 2^P of the centring chunk chosen by the device M step (both fp16 halves)
    aug01 = *augh_dev;
    aug2 = aug01 & 0xffffu;
  }
  // opaque copies: the compiler would otherwise rematerialise these from special registers (S2R / S2UR, ~50 cycles
  // each) inside the per-item loops, where registers are tight
  uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  asm volatile("" : "+r"(sbase));
  unsigned char* sgen = smem_dyn + (sbase - smem_u32(smem_dyn));
  asm volatile("" : "+l"(sgen));
  const uint32_t sBar = sbase + kCOffBar;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + kCOffBar + 8 * C_COUNT);
  auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };
  float* spar = reinterpret_cast<float*>(sgen + kCOffPar);  // [0] cinv2, [1] Ek, [2] Ea, [3] chat; stride 256
  float* slb = reinterpret_cast<float*>(sgen + kCOffLb);

  uint32_t tid = threadIdx.x;
  asm volatile("" : "+r"(tid));
  // The MMA warp addresses shared memory through sbase_hint, a kernel parameter (so that its address arithmetic
  // stays in uniform registers); a wrong hint is reported through err[1] and the host relaunches with the right one.
  if (sbase != sbase_hint) {
    if (tid == 0 && blockIdx.x == 0) err[1] = 0x80000000u | sbase;
    return;
  }
  if (skip != nullptr && (skip[0] | skip[1]) != 0u) return;  // aborted iteration, or the dense kernel runs instead
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31);
  const int64_t ngroups = (N + kCRows - 1) / kCRows;

  if (tid == 0) {
    for (int i = 0; i < kCStages; ++i) {
      mbar_init(bar(CB_FULL0 + i), 1);
      mbar_init(bar(CB_EMPTY0 + i), 3);  // the three MMA issuers
      mbar_init(bar(CA_READY0 + i), 4);  // 4 stager warps (lane quadrants)
      mbar_init(bar(CA_FREE0 + i), 1);   // the issuer of tile slot i
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(CG_FULL0 + i), 1);
      mbar_init(bar(CG_EMPTY0 + i), 3);  // the three MMA issuers
      mbar_init(bar(CT_FULL0 + i), 1);
      mbar_init(bar(CT_FULL0 + 2 + i), 1);  // accumulators 2 and 3 (64 dimensions)
      mbar_init(bar(CL_FULL0 + i), 8);   // 8 epilogue warps
      mbar_init(bar(CL_FREE0 + i), 4);   // 4 stager warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // drained-items counters of the two accumulators (4 arrivals per item), behind the TMEM slot
    for (int i = 0; i < 4; ++i) reinterpret_cast<volatile uint32_t*>(sgen + kCOffBar + 8 * C_COUNT + 16)[i] = 0u;
  }
  for (int i = (int)tid; i < 4 * 256; i += kThreadsTc) {
    const int a = i >> 8, k = i & 255;
    float val = k < K ? cpar[(size_t)a * K + k] : 0.f;
    if (a == 3 && gid == nullptr && k < K) val += lw[k];  // single group: fold E[log pi_k] into the constant
    spar[i] = val;
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base = *tmem_slot;
  asm volatile("" : "+r"(tmem_base));
  const bool tmem_ok = tmem_base == 0;  // all 512 columns are ours, so the allocation starts at column 0, lane 0
  if (!tmem_ok && tid == 0) atomicExch(err, 0xdead7e00u);

  if (!tmem_ok) {
    // nothing: fall through to the dealloc
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------- producer --
      uint32_t bcnt = 0, acnt = 0;
      for (int64_t gi = blockIdx.x; gi < ngroups; gi += gridDim.x) {
        for (int k = 0; k < K; ++k, ++bcnt) {
          if ((k & 3) == 0) {
            const uint32_t as = acnt & 1, aph = (acnt >> 1) & 1;
            mbar_wait_patient(bar(CG_EMPTY0 + as), aph ^ 1, err);
            mbar_expect_tx(bar(CG_FULL0 + as), kTcAugBlockBytes);
            bulk_g2s(sbase + kCOffAug + as * kTcAugBlockBytes, augblob + (size_t)(k >> 2) * kTcAugBlockBytes,
                     kTcAugBlockBytes, bar(CG_FULL0 + as));
            ++acnt;
          }
          const uint32_t bs = bcnt % kCStages, bph = (bcnt / kCStages) & 1;
          mbar_wait_patient(bar(CB_EMPTY0 + bs), bph ^ 1, err);
          const uint8_t* src = blob + (size_t)k * kBBlob;
          if constexpr (DIM == 128) {
            mbar_expect_tx(bar(CB_FULL0 + bs), kCBStage);
            bulk_g2s(sbase + bs * kCBStage, src, 16384u, bar(CB_FULL0 + bs));                  // hi, dims 0..63
            bulk_g2s(sbase + bs * kCBStage + 16384u, src + 32768u, 8192u, bar(CB_FULL0 + bs));  // hi, dims 64..127
          } else {
            mbar_expect_tx(bar(CB_FULL0 + bs), 8192u);
            bulk_g2s(sbase + bs * kCBStage, src, 8192u, bar(CB_FULL0 + bs));                   // hi, rows and dims 0..63
          }
        }
      }
    }
    if (warp == 1) coarse_mma_issuer<0, DIM>(sbase_hint, K, ngroups, err);
    if (warp == 2) coarse_mma_issuer<1, DIM>(sbase_hint, K, ngroups, err);
    if (warp == 3) coarse_mma_issuer<2, DIM>(sbase_hint, K, ngroups, err);
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // --------------------------------------------------------------- stagers --
    const int sw = warp - 4;  // == lane quadrant of the TMEM rows this warp may touch
    // X rows of group gi -> fp16(s_g x) in the staging buffers; a warp converts two rows per step
    auto stage_group = [&](int64_t gi) {
      for (int t = 0; t < kCT; ++t) {
        const int64_t n0 = gi * kCRows + (int64_t)t * kTM;
        unsigned char* dstT = sgen + kCOffStageA + (uint32_t)t * kCStageA;
        // lanes per row: 16 (128 dimensions, 8 floats each) or 8 (64 dimensions)
        constexpr int LPR = DIM / 8, RPW = 32 / LPR, NIT = kTM / (4 * RPW);
#pragma unroll 1
        for (int it0 = 0; it0 < NIT; it0 += 4) {
          float4 v[4][2];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = RPW * ((it0 + u) * 4 + sw) + lane / LPR;
            const int64_t n = n0 + r;
            if (n < N) {
              const float4* src = reinterpret_cast<const float4*>(X + n * DIM + 8 * (lane % LPR));
              v[u][0] = __ldg(src);
              v[u][1] = __ldg(src + 1);
            } else {
              v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = RPW * ((it0 + u) * 4 + sw) + lane / LPR;
            uint4 w;
            w.x = pack_f16x2_sat(v[u][0].x * sg, v[u][0].y * sg);
            w.y = pack_f16x2_sat(v[u][0].z * sg, v[u][0].w * sg);
            w.z = pack_f16x2_sat(v[u][1].x * sg, v[u][1].y * sg);
            w.w = pack_f16x2_sat(v[u][1].z * sg, v[u][1].w * sg);
            *reinterpret_cast<uint4*>(dstT + r * 256 + (((lane % LPR) ^ (r & 15)) << 4)) = w;
          }
        }
      }
    };
    // parked UB row of the finished group -> candidate bit mask (bit k of row n: UB_nk can reach e^-margin of the
    // row's best lower bound)
    auto mark_group = [&](int64_t gi, uint32_t parity) {
      const float* lb0 = slb + parity * 2 * kCRows;
      const int W = (K + 31) >> 5;
#pragma unroll 1
      for (int t = 0; t < kCT; ++t) {
        const int r = t * kTM + sw * 32 + lane;
        const int64_t n = gi * kCRows + r;
        if (n >= N) continue;
        const float thr = fmaxf(lb0[r], lb0[kCRows + r]) - margin;
        const float* qrow = q + n * ldq;
        const float4* q4 = reinterpret_cast<const float4*>(qrow);
        uint32_t* mrow = cmask + n * W;
        uint32_t word = 0;
        int k = 0, wi = 0;
        for (; k + 4 <= K; k += 4) {
          const float4 v = __ldcg(q4 + (k >> 2));
          const uint32_t b = (v.x >= thr ? 1u : 0u) | (v.y >= thr ? 2u : 0u) | (v.z >= thr ? 4u : 0u) | (v.w >= thr ? 8u : 0u);
          word |= b << (k & 31);
          if ((k & 31) == 28) {
            mrow[wi++] = word;
            word = 0;
          }
        }
        for (; k < K; ++k) {
          if (__ldcg(qrow + k) >= thr) word |= 1u << (k & 31);
        }
        if (K & 31) mrow[wi] = word;
      }
    };
    uint32_t gcnt = 0;
    int64_t gi = blockIdx.x, gprev = -1;
    if (gi < ngroups) stage_group(gi);
    named_bar_sync(1, 128);
    for (; gi < ngroups; gi += gridDim.x, ++gcnt) {
      // (a) staging -> TMEM A of every tile slot as soon as the previous group's MMAs released it
      const int row = sw * 32 + lane;
#pragma unroll 1
      for (int s = 0; s < kCT; ++s) {
        mbar_wait_patient(bar(CA_FREE0 + s), (gcnt & 1) ^ 1, err);
        tc_fence_after();
        const unsigned char* srcT = sgen + kCOffStageA + (uint32_t)s * kCStageA + row * 256;
        const uint32_t tA = tmem_base + ((uint32_t)(32 * sw) << 16) + 256u + kCAcol * (uint32_t)s;
#pragma unroll
        for (int h = 0; h < DIM / 32; ++h) {
          uint32_t rr[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 w = *reinterpret_cast<const uint4*>(srcT + (((4 * h + c) ^ (row & 15)) << 4));
            rr[4 * c] = w.x;
            rr[4 * c + 1] = w.y;
            rr[4 * c + 2] = w.z;
            rr[4 * c + 3] = w.w;
          }
          tmem_st16(tA + 16 * h, rr);
        }
        uint32_t ra[8] = {aug01, aug2, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(tA + 64u, ra);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(CA_READY0 + s));
      }
      named_bar_sync(1, 128);  // every stager is done reading the staging buffers
      // (b) candidate marking of the group that just finished
      if (gprev >= 0) {
        const uint32_t p = (gcnt - 1) & 1;
        mbar_wait_patient(bar(CL_FULL0 + p), ((gcnt - 1) >> 1) & 1, err);
        mark_group(gprev, p);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(CL_FREE0 + p));
      }
      // (c) stage the next group while this one runs
      if (gi + gridDim.x < ngroups) stage_group(gi + gridDim.x);
      named_bar_sync(1, 128);
      gprev = gi;
    }
    if (gprev >= 0) {
      const uint32_t p = (gcnt - 1) & 1;
      mbar_wait_patient(bar(CL_FULL0 + p), ((gcnt - 1) >> 1) & 1, err);
      mark_group(gprev, p);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    // -------------------------------------------------------------- epilogue --
    const int grp = (warp - 8) >> 2, quad = warp & 3;
    constexpr uint32_t NACC = DIM == 64 ? 4u : 2u, LGA = DIM == 64 ? 2u : 1u, ACCW = DIM == 64 ? 64u : 128u;
    const uint32_t tacc0 = tmem_base + ((uint32_t)(32 * quad) << 16);
    const bool grouped = gid != nullptr;
    const bool special = grouped || act != nullptr;  // per-row weights or an active mask: the rare, slower tail
    const uint32_t spar_s = sbase + kCOffPar;
    uint32_t icnt = 0, gcnt = 0;
    for (int64_t gi = blockIdx.x; gi < ngroups; gi += gridDim.x, ++gcnt) {
      float* qrow[kCT];
      float xn[kCT], lbmax[kCT];
      const float* lwg[kCT];
      const uint8_t* actg[kCT];
#pragma unroll
      for (int s = 0; s < kCT; ++s) {
        const int64_t n = gi * kCRows + s * kTM + 32 * quad + lane;
        const bool valid = n < N;
        qrow[s] = valid ? q + n * ldq : nullptr;
        xn[s] = valid ? __ldg(xnorm + n) : 0.f;
        const int g = (grouped && valid) ? gid[n] : 0;
        lwg[s] = lw + (size_t)g * K;
        actg[s] = act != nullptr ? act + (size_t)g * K : nullptr;
        lbmax[s] = -INFINITY;
      }
#pragma unroll 1
      for (int k = 0; k < K; ++k) {
        float cinv2, ek, ea, ch;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cinv2) : "r"(spar_s + 4u * (uint32_t)k));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ek) : "r"(spar_s + 4u * (uint32_t)k + 1024u));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ea) : "r"(spar_s + 4u * (uint32_t)k + 2048u));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ch) : "r"(spar_s + 4u * (uint32_t)k + 3072u));
#pragma unroll
        for (int s = 0; s < kCT; ++s, ++icnt) {
          // the icnt-th item of this CTA lives in accumulator icnt & (NACC - 1) (see the MMA issuers); this group
          // drains the accumulators of its parity
          if ((int)(icnt & 1) != grp) continue;
          const uint32_t acc = icnt & (NACC - 1u);
          const uint32_t tacc = tacc0 + ACCW * acc;
          const uint32_t bfull = bar(CT_FULL0 + (int)acc), drained = sBar + 8u * C_COUNT + 16u + 4u * acc;
          // everything that does not depend on the accumulator first: its latency hides behind the wait
          float c = ch;
          bool off = false;
          if (special) {
            if (grouped) c += __ldg(lwg[s] + k);
            off = actg[s] != nullptr && !__ldg(actg[s] + k);
          }
          const float e = fmaf(ek, xn[s], ea);
          mbar_wait(bfull, (icnt >> LGA) & 1, err);
          tc_fence_after();
          uint32_t r[DIM];
          tmem_ld64(tacc, r);
          if constexpr (DIM == 128) tmem_ld64(tacc + 64u, r + 64);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("red.release.cta.shared.add.u32 [%0], 1;" ::"r"(drained) : "memory");
          const float ss = sumsq_regs<DIM>(r);
          if (qrow[s] != nullptr) {
            float d;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(ss * cinv2));
            const float dlo = fmaxf(d - e, 0.f), dhi = d + e;
            float ub = fmaf(-0.5f * dlo, dlo, c), lb = fmaf(-0.5f * dhi, dhi, c);
            ub += 2e-6f * fabsf(ub) + 1e-3f;   // fp32 rounding of the bound itself (approximate square root)
            lb -= 2e-6f * fabsf(lb) + 1e-3f;
            if (off) ub = lb = -INFINITY;
            qrow[s][k] = ub;
            lbmax[s] = fmaxf(lbmax[s], lb);
          }
        }
      }
      // hand the per-row maxima of the lower bounds to the stagers
      const uint32_t p = gcnt & 1;
      mbar_wait(bar(CL_FREE0 + p), ((gcnt >> 1) & 1) ^ 1, err);
      float* dst = slb + (p * 2 + grp) * kCRows;
#pragma unroll
      for (int s = 0; s < kCT; ++s) dst[s * kTM + 32 * quad + lane] = lbmax[s];
      __threadfence_block();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(CL_FULL0 + p));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Level 3: row soft-max over the candidate logits in q (cmask: W words per row):
// q = exp(logit - log Z) for candidates, 0 elsewhere; Fz += sum_n log Z_n.
// The first version gave every 16-byte group of a row to a lane and ran the whole
// soft-max arithmetic on all of them; with one or two candidates per row it was
// issue-bound (84 % issue slots, 143 warp instructions per row,
// profiles/ncu_r01_finalize_v5_raw.csv).  Now a warp takes 32 rows:
//   phase A  lane = row: walk the set bits of the row's mask, max and sum of
//            exponentials over its candidates only; exp(logit - max) is parked
//            in place
//   phase B  GP lanes per row: rewrite the row as parked value / sum under the
//            mask, zeros elsewhere (only groups that hold a candidate are read)
// so the pass costs the row write plus a sector or two of reads per row.
// ---------------------------------------------------------------------------
template <int LGP>  // log2 of the lanes that share a row in phase B: 4 << LGP >= K
__global__ void __launch_bounds__(256)
estep_finalize_kernel(float* __restrict__ q, int64_t ldq, int64_t N, int K, const uint32_t* __restrict__ cmask, int W,
                      double* __restrict__ Fz, const unsigned* __restrict__ skip, double* __restrict__ Hk) {
  constexpr int GP = 1 << LGP, WMAX = (4 * GP + 31) / 32;
  if (skip != nullptr && (skip[0] | skip[1]) != 0u) return;
  // Optional split scores H_k = sum_n q_nk * logit_nk (the ranking of split_gr, src/cluster.cpp:401-415, from the
  // quantities this pass holds anyway; only fits ask for them): per-CTA fp64 sums in shared memory, one global atomic
  // per cluster and CTA.
  __shared__ double sH[4 * GP];
  if (Hk != nullptr) {
    for (int k = threadIdx.x; k < 4 * GP; k += blockDim.x) sH[k] = 0.0;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double fz = 0;
  for (int64_t r0 = warp0 * 32; r0 < N; r0 += nwarps * 32) {
    // ---- phase A ----
    const int64_t n = r0 + lane;
    uint32_t mw[WMAX];
#pragma unroll
    for (int w = 0; w < WMAX; ++w) mw[w] = (n < N && w < W) ? __ldg(cmask + n * W + w) : 0u;
    float* qrow = q + (n < N ? n : 0) * ldq;
    float mx = -INFINITY;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      uint32_t word = mw[w];
      while (word) {
        const int k = 32 * w + __ffs(word) - 1;
        word &= word - 1;
        mx = fmaxf(mx, qrow[k]);
      }
    }
    float se = 0.f;
    if (mx > -INFINITY) {
#pragma unroll
      for (int w = 0; w < WMAX; ++w) {
        uint32_t word = mw[w];
        while (word) {
          const int k = 32 * w + __ffs(word) - 1;
          word &= word - 1;
          const float e = expf(qrow[k] - mx);
          qrow[k] = e;
          se += e;
        }
      }
      fz += (double)(logf(se) + mx);
    }
    const float inv = se > 0.f ? 1.0f / se : 0.f;
    if (Hk != nullptr && se > 0.f) {
#pragma unroll
      for (int w = 0; w < WMAX; ++w) {
        uint32_t word = mw[w];
        while (word) {
          const int k = 32 * w + __ffs(word) - 1;
          word &= word - 1;
          const float e = qrow[k];   // exp(logit - max), parked above
          if (e > 0.f) atomicAdd(&sH[k], (double)(e * inv) * ((double)mx + (double)logf(e)));
        }
      }
    }
    __syncwarp();  // the parked values are read by other lanes below
    // ---- phase B ----
#pragma unroll 4
    for (int idx = lane; idx < 32 * GP; idx += 32) {
      const int row = idx >> LGP, g = idx & (GP - 1);
      const float invr = __shfl_sync(0xffffffffu, inv, row);
      uint32_t word = 0;
#pragma unroll
      for (int w = 0; w < WMAX; ++w) {
        const uint32_t t = __shfl_sync(0xffffffffu, mw[w], row);
        if ((g >> 3) == w) word = t;
      }
      const uint32_t nib = (word >> ((4 * g) & 31)) & 0xFu;
      const int64_t nr = r0 + row;
      const int k = 4 * g;
      if (nr < N && k < K) {
        float* dst = q + nr * ldq + k;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nib) {
          const float4 v = *reinterpret_cast<const float4*>(dst);
          o.x = (nib & 1u) ? v.x * invr : 0.f;
          o.y = (nib & 2u) ? v.y * invr : 0.f;
          o.z = (nib & 4u) ? v.z * invr : 0.f;
          o.w = (nib & 8u) ? v.w * invr : 0.f;
        }
        if (k + 4 <= K) {
          *reinterpret_cast<float4*>(dst) = o;
        } else {
          dst[0] = o.x;
          if (k + 1 < K) dst[1] = o.y;
          if (k + 2 < K) dst[2] = o.z;
        }
      }
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) fz += __shfl_xor_sync(0xffffffffu, fz, o);
  if (lane == 0 && fz != 0.0) atomicAdd(Fz, fz);
  if (Hk != nullptr) {
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x)
      if (sH[k] != 0.0) atomicAdd(&Hk[k], sH[k]);
  }
}

// q[n][k] = -inf where the candidate bit is clear (only the LCB_TC_STAGE test modes look at this)
__global__ void __launch_bounds__(256)
apply_mask_kernel(float* __restrict__ q, int64_t ldq, int64_t N, int K, const uint32_t* __restrict__ cmask, int W) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int64_t n = i / K;
  const int k = (int)(i - n * K);
  if (!((cmask[n * W + (k >> 5)] >> (k & 31)) & 1u)) q[n * ldq + k] = -INFINITY;
}

// ---------------------------------------------------------------------------
// Candidate masks -> per-cluster row lists (same block structure as nz_count / nz_fill in kernels.cu, so that
// nz_scan serves both): counts per (row block, cluster), then the rows at their scanned offsets.
// ---------------------------------------------------------------------------
constexpr int kMaskBlock = 2048;  // == kNzBlock
template <bool kFill>
__global__ void __launch_bounds__(256)
mask_lists_kernel(const uint32_t* __restrict__ cmask, int W, int64_t N, int K, int32_t* __restrict__ blockcnt,
                  const long long* __restrict__ koff, int32_t* __restrict__ lrow, const unsigned* __restrict__ skip) {
  extern __shared__ int scnt[];
  if (skip != nullptr && (skip[0] | skip[1]) != 0u) return;
  for (int k = threadIdx.x; k < K; k += 256) scnt[k] = 0;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * kMaskBlock;
  const int64_t r1 = (r0 + kMaskBlock < N) ? r0 + kMaskBlock : N;
  for (int64_t n = r0 + threadIdx.x; n < r1; n += 256) {
    for (int w = 0; w < W; ++w) {
      uint32_t word = cmask[n * W + w];
      while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        const int k = 32 * w + b;
        if (k >= K) break;
        const int pos = atomicAdd(&scnt[k], 1);
        if (kFill) lrow[koff[k] + blockcnt[(size_t)blockIdx.x * K + k] + pos] = (int32_t)n;
      }
    }
  }
  if (!kFill) {
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) blockcnt[(size_t)blockIdx.x * K + k] = scnt[k];
  }
}

// Work items of level 2: 128-entry chunks of the per-cluster lists, resolved once instead of by every warp role
__global__ void __launch_bounds__(256)
build_items_kernel(const int32_t* __restrict__ itoff, const long long* __restrict__ koff,
                   const long long* __restrict__ kcnt, int K, int64_t nitems, int4* __restrict__ items,
                   const long long* __restrict__ nitems_dev, const unsigned* __restrict__ skip) {
  if (skip != nullptr && *skip != 0u) return;
  if (nitems_dev != nullptr) {
    const int64_t cap = nitems;
    nitems = (int64_t)*nitems_dev;
    if (nitems > cap) nitems = cap;
  }
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += (int64_t)gridDim.x * blockDim.x) {
    const ListItem r = list_item(it, K, itoff, koff, kcnt);
    items[it] = make_int4(r.k, r.count, (int)(r.base & 0xffffffffLL), (int)(r.base >> 32));
  }
}

// lq[e] = q[lrow[e]][k] for every entry of cluster k's list and Nk[k] = sum_e lq[e]: the statistics pass can then
// reuse the candidate lists of the E pass instead of sweeping q again (single group, no sparse mask).
__global__ void __launch_bounds__(256)
gather_list_q_kernel(const float* __restrict__ q, int64_t ldq, const int32_t* __restrict__ lrow,
                     const long long* __restrict__ koff, const long long* __restrict__ kcnt, float* __restrict__ lq,
                     double* __restrict__ Nk, const int32_t* __restrict__ gid, int K) {
  const int k = blockIdx.y;
  const long long cnt = kcnt[k], base = koff[k];
  if (gid != nullptr) {
    // grouped model: Nk is [J][K].  The lists are in row order and rows are grouped, so a warp nearly always sees
    // one group: one warp reduction and one atomic per 32 entries; a warp that straddles a group boundary adds
    // entry by entry.
    const int lane = threadIdx.x & 31;
    for (long long e0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); e0 < cnt;
         e0 += (long long)gridDim.x * blockDim.x) {
      const long long e = e0 + lane;
      const bool valid = e < cnt;
      float v = 0.f;
      int g = -1;
      if (valid) {
        const int32_t row = lrow[base + e];
        v = q[(int64_t)row * ldq + k];
        lq[base + e] = v;
        g = gid[row];
      }
      const int g0 = __shfl_sync(0xffffffffu, g, 0);
      if (__all_sync(0xffffffffu, !valid || g == g0)) {
        double a = (double)v;
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && a != 0.0) atomicAdd(&Nk[(size_t)g0 * K + k], a);
      } else if (valid && v != 0.f) {
        atomicAdd(&Nk[(size_t)g * K + k], (double)v);
      }
    }
    return;
  }
  double acc = 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (long long)gridDim.x * blockDim.x) {
    const float v = q[(int64_t)lrow[base + e] * ldq + k];
    lq[base + e] = v;
    acc += (double)v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += part[i];
    if (t != 0.0) atomicAdd(&Nk[k], t);
  }
}

// Euclidean norm of every (centred) row, D == 128 or 64: one warp per row, float4 (float2) per lane
__global__ void __launch_bounds__(256)
row_norm128_kernel(const float* __restrict__ X, int64_t N, float* __restrict__ out, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp0; n < N; n += nwarps) {
    float4 v;
    if (dim == 128) {
      v = __ldg(reinterpret_cast<const float4*>(X + n * 128) + lane);
    } else {
      const float2 h = __ldg(reinterpret_cast<const float2*>(X + n * 64) + lane);
      v = make_float4(h.x, h.y, 0.f, 0.f);
    }
    float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[n] = sqrtf(s);
  }
}

}  // namespace

bool tc_supported(int D, int64_t ldx) { return D == 128 && ldx == 128; }
int tc_dim(int D, int64_t ldx) { return (D == 128 && ldx == 128) ? 128 : (D == 64 && ldx == 64) ? 64 : 0; }

cudaError_t estep_tc128(cudaStream_t st, int sms, const float* X, int64_t N, const int32_t* gid, int K,
                        const uint8_t* blob, const float* ascale, const float* inv_t2, const float* chat,
                        const float* lw, const uint8_t* act, float* q, int64_t ldq, double* Fz, unsigned* err,
                        const unsigned* skip, int dim) {
  if (N <= 0) return cudaSuccess;
  if (dim != 128 && dim != 64) return cudaErrorInvalidValue;
  auto kern = dim == 128 ? estep_tc128_kernel<false, 128> : estep_tc128_kernel<false, 64>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return e;
  const int64_t ntiles = (N + kTM - 1) / kTM;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  kern<<<grid, kThreadsTc, kSmemBytes, st>>>(X, N, gid, K, blob, ascale, inv_t2, chat, lw, act, q, ldq, Fz, err, nullptr,
                                             nullptr, 0, nullptr, skip);
  return cudaGetLastError();
}

cudaError_t estep_tc128_list(cudaStream_t st, int sms, const float* X, int64_t N, const int32_t* gid, int K,
                             const uint8_t* blob, const float* ascale, const float* inv_t2, const float* chat,
                             const float* lw, const int32_t* lrow, const long long* koff, const long long* kcnt,
                             const int32_t* itoff, int64_t nitems, void* items, float* q, int64_t ldq, unsigned* err,
                             const long long* nitems_dev, const unsigned* skip, int dim) {
  if (N <= 0 || (nitems <= 0 && nitems_dev == nullptr)) return cudaSuccess;
  if (dim != 128 && dim != 64) return cudaErrorInvalidValue;
  // nitems_dev != NULL: the item count lives on the device; `nitems` is then the capacity of `items`
  const int64_t bgrid = nitems_dev ? (int64_t)sms * 8 : (nitems + 255) / 256;
  build_items_kernel<<<(unsigned)bgrid, 256, 0, st>>>(itoff, koff, kcnt, K, nitems, (int4*)items, nitems_dev, skip);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  auto kern = dim == 128 ? estep_tc128_kernel<true, 128> : estep_tc128_kernel<true, 64>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return e;
  const int grid = (int)(nitems < sms && nitems_dev == nullptr ? nitems : sms);
  kern<<<grid, kThreadsTc, kSmemBytes, st>>>(X, N, gid, K, blob, ascale, inv_t2, chat, lw, nullptr, q, ldq, nullptr, err,
                                             lrow, (const int4*)items, nitems, nitems_dev, skip);
  return cudaGetLastError();
}

cudaError_t estep_coarse_tc128(cudaStream_t st, int sms, const float* X, const float* xnorm, int64_t N,
                               const int32_t* gid, int K, const uint8_t* blob, const uint8_t* augblob,
                               const float* cpar, const float* lw, const uint8_t* act, float sg, int aug_exp,
                               float margin, float* q, int64_t ldq, uint32_t* cmask, uint32_t sbase_hint,
                               unsigned* err, const unsigned* augh_dev, const unsigned* skip, int dim) {
  if (N <= 0) return cudaSuccess;
  if (K < 1 || K > kTcCoarseMaxK || aug_exp < 0 || aug_exp > 15 || (dim != 128 && dim != 64)) return cudaErrorInvalidValue;
  auto kern = dim == 128 ? estep_coarse_tc128_kernel<128> : estep_coarse_tc128_kernel<64>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCSmemBytes);
  if (e != cudaSuccess) return e;
  const int64_t ngroups = (N + kCRows - 1) / kCRows;
  const int grid = (int)(ngroups < sms ? ngroups : sms);
  const uint32_t h = (uint32_t)__half_as_ushort(__float2half_rn(ldexpf(1.f, aug_exp)));
  kern<<<grid, kThreadsTc, kCSmemBytes, st>>>(X, xnorm, N, gid, K, blob, augblob, cpar, lw, act, sg, h | (h << 16), h,
                                              margin, q, ldq, cmask, sbase_hint, err, augh_dev, skip);
  return cudaGetLastError();
}

cudaError_t estep_finalize(cudaStream_t st, int sms, float* q, int64_t ldq, int64_t N, int K, const uint32_t* cmask,
                           double* Fz, const unsigned* skip, double* Hk) {
  if (N <= 0) return cudaSuccess;
  if (K > 256 || (ldq & 3)) return cudaErrorInvalidValue;
  const int W = (K + 31) / 32;
  const int grid = sms * 16;
  if (K <= 32) estep_finalize_kernel<3><<<grid, 256, 0, st>>>(q, ldq, N, K, cmask, W, Fz, skip, Hk);
  else if (K <= 64) estep_finalize_kernel<4><<<grid, 256, 0, st>>>(q, ldq, N, K, cmask, W, Fz, skip, Hk);
  else if (K <= 128) estep_finalize_kernel<5><<<grid, 256, 0, st>>>(q, ldq, N, K, cmask, W, Fz, skip, Hk);
  else estep_finalize_kernel<6><<<grid, 256, 0, st>>>(q, ldq, N, K, cmask, W, Fz, skip, Hk);
  return cudaGetLastError();
}

cudaError_t apply_candidate_mask(cudaStream_t st, float* q, int64_t ldq, int64_t N, int K, const uint32_t* cmask) {
  if (N <= 0) return cudaSuccess;
  const int64_t total = N * K;
  apply_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(q, ldq, N, K, cmask, (K + 31) / 32);
  return cudaGetLastError();
}

cudaError_t mask_count(cudaStream_t st, const uint32_t* cmask, int64_t N, int K, int32_t* blockcnt, const unsigned* skip) {
  if (N <= 0) return cudaSuccess;
  mask_lists_kernel<false><<<(unsigned)((N + kMaskBlock - 1) / kMaskBlock), 256, sizeof(int) * K, st>>>(
      cmask, (K + 31) / 32, N, K, blockcnt, nullptr, nullptr, skip);
  return cudaGetLastError();
}

cudaError_t mask_fill(cudaStream_t st, const uint32_t* cmask, int64_t N, int K, int32_t* blockoff, const long long* koff,
                      int32_t* lrow, const unsigned* skip) {
  if (N <= 0) return cudaSuccess;
  mask_lists_kernel<true><<<(unsigned)((N + kMaskBlock - 1) / kMaskBlock), 256, sizeof(int) * K, st>>>(
      cmask, (K + 31) / 32, N, K, blockoff, koff, lrow, skip);
  return cudaGetLastError();
}

cudaError_t gather_list_q(cudaStream_t st, int sms, const float* q, int64_t ldq, const int32_t* lrow,
                          const long long* koff, const long long* kcnt, long long maxcnt, int K, float* lq, double* Nk,
                          const int32_t* gid) {
  if (K <= 0 || maxcnt <= 0) return cudaSuccess;
  long long bx = (maxcnt + 255) / 256;
  const long long cap = std::max<long long>(1, (long long)sms * 8 / K);
  if (bx > cap) bx = cap;
  gather_list_q_kernel<<<dim3((unsigned)bx, (unsigned)K), 256, 0, st>>>(q, ldq, lrow, koff, kcnt, lq, Nk, gid, K);
  return cudaGetLastError();
}

cudaError_t row_norm128(cudaStream_t st, int sms, const float* X, int64_t N, float* out, int dim) {
  if (N <= 0) return cudaSuccess;
  if (dim != 128 && dim != 64) return cudaErrorInvalidValue;
  const int64_t want = (N + 7) / 8;
  const int grid = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
  row_norm128_kernel<<<grid, 256, 0, st>>>(X, N, out, dim);
  return cudaGetLastError();
}

// Host-side packing of the aug chunk of cluster k into its 4-cluster block: row i (output dimension), K slots
// 0..2 of the 16-wide chunk (k & 3) hold the fp16 hi/mid/lo split of w[i]; same swizzled layout as the B blocks.
// Returns the largest residual |w[i] - (h1 + h2 + h3)|.
double tc_pack_aug(const double* w /* [128] */, int k, uint8_t* augblob) {
  uint8_t* blk = augblob + (size_t)(k >> 2) * kTcAugBlockBytes;
  const int j = k & 3;
  double worst = 0;
  for (int i = 0; i < 128; ++i) {
    double rem = w[i];
    for (int slot = 0; slot < 16; ++slot) {
      __half h = __float2half_rn(0.f);
      if (slot < 3) {
        h = __float2half_rn((float)rem);
        rem -= (double)__half2float(h);
      }
      const uint32_t kk = 16u * (uint32_t)j + (uint32_t)slot;
      const uint32_t off = (uint32_t)i * 128u + ((((kk >> 3) ^ ((uint32_t)i & 7u))) << 4) + ((kk & 7u) << 1);
      *reinterpret_cast<__half*>(blk + off) = h;
    }
    worst = std::max(worst, std::fabs(rem));
  }
  return worst;
}

cudaError_t sstat_tc128(cudaStream_t st, int sms, const float* X, const int32_t* lrow, const float* lq,
                        const long long* koff, const long long* kcnt, long long maxcnt, long long nnz, int K,
                        const float* cen, float scale, double* xs, double* S, unsigned* err, const float* scale_dev,
                        const unsigned* skip, int dim) {
  if (K <= 0 || maxcnt <= 0) return cudaSuccess;
  static const int pf = [] {
    const char* e = std::getenv("LCB_SSTAT_PREFETCH");  // 0 none, 1 prefetch.global.L2 by an idle warp, 2 bulk prefetch
    return e ? std::atoi(e) : 0;
  }();
  if (dim != 128 && dim != 64) return cudaErrorInvalidValue;
  // rows folded into one fp32 TMEM accumulator before the fp64 add: fewer for small problems (more accurate, and the
  // extra atomics are free there), kTcScatterChunk for large ones
  const int chunk_rows = nnz <= (1LL << 20) ? 128 : kTcScatterChunk;
  auto kern = dim == 128 ? sstat_tc128_kernel<128> : sstat_tc128_kernel<64>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSSmemBytes);
  if (e != cudaSuccess) return e;
  // persistent CTAs walk the (cluster, chunk) items; no more CTAs than items
  const long long items_max = (nnz + chunk_rows - 1) / chunk_rows + K;
  if (K > kTcCoarseMaxK || items_max > 2000000000LL) return cudaErrorInvalidValue;
  const int grid = (int)(items_max < sms ? items_max : sms);
  kern<<<grid, kScatterThreads, kSSmemBytes, st>>>(X, lrow, lq, koff, kcnt, K, cen, scale, chunk_rows, xs, S, err, scale_dev,
                                                   skip, pf);
  return cudaGetLastError();
}

// Host-side packing of one cluster's B operand: Bs[i][d] = R[i][d] * bscale,
// i >= d, split into fp16 hi/lo and laid out exactly as the kernel's shared
// memory expects (row pitch 128 B = 64 fp16 of one K block, 16-byte chunks
// XOR-swizzled with row % 8; K block 1 stores rows 64..127 only).
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__CUDA_ARCH__)
// eight consecutive K entries of one row are one 16-byte chunk of the swizzled layout: convert them with F16C
__attribute__((target("avx2,f16c"))) static void pack_rows_f16c(const double* R, double bscale, uint8_t* out) {
  for (int kbk = 0; kbk < 2; ++kbk) {
    const uint32_t base_hi = kbk == 0 ? 0u : 2 * kAPart;
    const uint32_t base_lo = base_hi + (kbk == 0 ? kAPart : kAPart / 2);
    for (int i = 64 * kbk; i < 128; ++i) {
      const int r = i - 64 * kbk;
      for (int ch = 0; ch < 8; ++ch) {
        const int d0 = 64 * kbk + 8 * ch;
        if (d0 > i) break;
        float v[8];
        for (int e = 0; e < 8; ++e) v[e] = d0 + e <= i ? (float)(R[(size_t)i * 128 + d0 + e] * bscale) : 0.f;
        const __m256 x = _mm256_loadu_ps(v);
        const __m128i h = _mm256_cvtps_ph(x, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
        const __m256 back = _mm256_cvtph_ps(h);
        const __m128i l = _mm256_cvtps_ph(_mm256_sub_ps(x, back), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
        const uint32_t off = (uint32_t)r * 128u + (((uint32_t)ch ^ ((uint32_t)r & 7u)) << 4);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + base_hi + off), h);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + base_lo + off), l);
      }
    }
  }
}
static bool have_f16c() {
  static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("f16c");
  return ok;
}
#else
static bool have_f16c() { return false; }
static void pack_rows_f16c(const double*, double, uint8_t*) {}
#endif

static bool g_pack_force_scalar = false;

// Packs one operand with the F16C path and with the portable path; returns the number of differing bytes (0 where
// F16C is not available: there is only one path then).
int tc_pack_selftest() {
  std::vector<double> R((size_t)128 * 128, 0.0), rel(128);
  uint64_t st = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(int64_t)(st >> 11) / 9007199254740992.0 * 2.0 - 1.0;
  };
  for (int i = 0; i < 128; ++i) {
    rel[i] = 3.0 * rnd();
    for (int j = 0; j <= i; ++j) R[(size_t)i * 128 + j] = std::ldexp(rnd(), (int)(8.0 * rnd()));  // wide dynamic range
  }
  std::vector<uint8_t> a(kBBlob), b(kBBlob);
  tc_pack_cluster(R.data(), 37.5, rel.data(), 0.25, a.data());
  g_pack_force_scalar = true;
  tc_pack_cluster(R.data(), 37.5, rel.data(), 0.25, b.data());
  g_pack_force_scalar = false;
  int diff = 0;
  for (uint32_t i = 0; i < kBBlob; ++i) diff += a[i] != b[i];
  return diff;
}

void tc_pack_cluster(const double* R /* [128][128] row-major lower-triangular */, double bscale,
                     const double* mean_rel /* [128] cluster mean minus the data centre */, double ascale, uint8_t* out) {
  std::memset(out, 0, kBBlob);
  float* mh = reinterpret_cast<float*>(out + kOffMean);
  float* nl = mh + 128;
  for (int d = 0; d < 128; ++d) {
    const float hi = (float)mean_rel[d];
    mh[d] = hi;
    nl[d] = (float)(-(mean_rel[d] - (double)hi) * ascale);
  }
  if (have_f16c() && !g_pack_force_scalar) {
    pack_rows_f16c(R, bscale, out);
    return;
  }
  for (int kbk = 0; kbk < 2; ++kbk) {
    const uint32_t base_hi = kbk == 0 ? 0u : 2 * kAPart;
    const uint32_t base_lo = base_hi + (kbk == 0 ? kAPart : kAPart / 2);
    for (int i = 64 * kbk; i < 128; ++i) {
      const int r = i - 64 * kbk;
      for (int kk = 0; kk < 64; ++kk) {
        const int dcol = 64 * kbk + kk;
        if (dcol > i) continue;
        const float v = (float)(R[(size_t)i * 128 + dcol] * bscale);
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)kk >> 3) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)kk & 7u) << 1);
        *reinterpret_cast<__half*>(out + base_hi + off) = h;
        *reinterpret_cast<__half*>(out + base_lo + off) = l;
      }
    }
  }
}

}  // namespace dev
}  // namespace lcb
