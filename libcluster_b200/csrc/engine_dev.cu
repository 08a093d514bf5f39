// engine_dev.cu -- the VB iteration against a device-resident model (Engine::dev_*).
//
// One iteration of vbem() (src/cluster.cpp:203-226) as a fixed sequence of launches on the engine's stream:
//
//   S pass (statistics of q)  ->  all-reduce of the packed statistics (the one data-sized collective, 8.45 MB at
//   K = 64, D = 128)  ->  M step on the device (mstep.cu: posteriors, Cholesky, E-step operands, Fc, Fw)  ->
//   E pass (new q, sum log Z)  ->  all-reduce of {sum log Z, rerun votes} (16 bytes)  ->  one 192-byte record to the
//   host and the only synchronisation of the iteration.
//
// Everything data-dependent that the host used to decide between kernels (list lengths, work-item counts, operand
// scales, the exponent of the level-1 centring chunk) is decided on the device (list_plan, mstep_finish); the host
// only learns about it afterwards, from the record, and repairs the rare cases after the fact:
//   * non-zero lists of the S pass longer than their buffer: every later kernel of the iteration returns at once
//     (q untouched), all ranks see the abort flag in the all-reduced statistics, the buffer grows, the iteration
//     is repeated;
//   * candidate lists of the two-level E pass too long (or not worth it): this rank repeats its E pass with the
//     dense kernel, all ranks repeat the 16-byte all-reduce.
// The host copies of the posteriors (weights_, clusters_) are rebuilt from the device's raw statistics when a
// vbem() ends or a getter asks for them (dev_sync_host).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.hpp"
#include "kernels.cuh"
#include "mstep.cuh"
#include "tc_kernels.cuh"

namespace lcb {

namespace {
enum { kF32 = 0, kF64 = 1 };
inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
constexpr size_t kIterBytes = 256;  // iteration record (32 doubles: 16 slots, 2 reduced, spare)
constexpr size_t kCtlOff = 256;     // control words behind it
constexpr size_t kScaleOff = 320;   // operand scale of the tensor-core scatter
constexpr int kReduced = 16;        // slot of the all-reduced {sum log Z, rerun votes}
}  // namespace

Engine::TcLayout Engine::tc_layout(int J, int K, bool two) const {
  TcLayout l;
  const size_t nblob = (size_t)K * dev::kTcBlobBytes;
  const size_t nfl = 3 * (size_t)K + (size_t)J * K;  // ascale, inv_t2, chat, lw
  l.naug = two ? (size_t)((K + 3) / 4) * dev::kTcAugBlockBytes : 0;
  const size_t ncpar = two ? 4 * (size_t)K : 0;
  l.off_f = nblob;
  l.off_aug = (l.off_f + nfl * sizeof(float) + 1023) / 1024 * 1024;
  l.off_cpar = l.off_aug + l.naug;
  l.total = l.off_cpar + ncpar * sizeof(float) + 16;
  return l;
}

void Engine::ensure_host_model() {
  if (host_stale_ && dev_live_ && dev_view_ == &main_) dev_sync_host(main_, weights_, clusters_);
  host_stale_ = false;
}

void Engine::dev_drop() {
  ensure_host_model();
  dev_live_ = false;
}

// Buffers of the device model and the centres of the first statistics pass (the host's choice, as in
// Engine::iteration: hint, posterior mean, or -- for fresh clusters without a hint -- the weighted means of a probe
// pass about the data centre).
void Engine::dev_begin(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                       std::vector<std::vector<double>>& hints) {
  (void)weights;
  const int J = v.J, K = v.K, D = v.D;
  const bool full = ckind_ == kGaussWish;
  const int cld = full ? dev::full_dp(D) : D;
  if (full && cld == 0) throw_invalid("full-covariance models support D <= 256");
  const size_t es = prec_ == kF32 ? 4 : 8;
  const int64_t Sz = full ? (int64_t)D * D : D;
  const int64_t nstat = (int64_t)J * K + (int64_t)K * D + K * Sz;
  reserve(d_stats_, sizeof(double) * (size_t)(nstat + 2));
  reserve(d_cen_, (size_t)K * cld * es);
  reserve(d_raw_, sizeof(double) * (size_t)K * (1 + D + Sz));
  reserve(d_post_, sizeof(double) * (size_t)K * dev::kPostStride);
  reserve(d_iter_, 512);
  reserve(d_centre_, sizeof(double) * (size_t)D + 16);
  reserve(d_wscr_, sizeof(double) * (size_t)(J + 2));
  reserve(d_err_, 16);
  if (full) {
    const size_t vec = (size_t)(8 * D + 16) * sizeof(double), work = dev::mstep_work_doubles(D) * sizeof(double);
    if (vec + work > 200 * 1024) reserve(d_work_, (size_t)K * work);
  }

  std::vector<std::vector<double>> centres(K);
  bool blind = false;
  for (int k = 0; k < K; ++k) {
    const bool have_hint = k < (int)hints.size() && (int)hints[k].size() == D;
    if (hints_first_ && have_hint) centres[k] = hints[k];
    else if (clusters[k].getN() > 0) centres[k] = clusters[k].mean();
    else if (have_hint) centres[k] = hints[k];
    else centres[k] = centre_;
    if (!(clusters[k].getN() > 0) && !have_hint) blind = true;
  }
  hints_first_ = false;
  blind = blind && K > 1;

  unsigned char* h = (unsigned char*)pinned((size_t)K * cld * es + sizeof(double) * (size_t)D + 64);
  std::memset(h, 0, (size_t)K * cld * es);
  double cmax = 0;
  for (int k = 0; k < K; ++k)
    for (int d = 0; d < D; ++d) {
      const double rel = centres[k][d] - centre_[d];
      double used;
      if (prec_ == kF32) {
        const float f = (float)rel;
        ((float*)h)[(size_t)k * cld + d] = f;
        used = (double)f;
      } else {
        ((double*)h)[(size_t)k * cld + d] = rel;
        used = rel;
      }
      cmax = std::max(cmax, std::fabs(used));
    }
  double* hc = reinterpret_cast<double*>(h + round_up((int64_t)((size_t)K * cld * es), 16));
  std::memcpy(hc, centre_.data(), sizeof(double) * (size_t)D);
  float* hs = reinterpret_cast<float*>(hc + D);
  auto scale_for = [&](double cm) {
    const double span = std::max(xabs_max_ + cm, 1e-30);
    return (float)std::ldexp(1.0, std::min(100, std::max(-100, (int)std::floor(std::log2(16384.0 / span)))));
  };
  hs[0] = scale_for(cmax);
  hs[1] = scale_for(xabs_max_);  // after a probe pass the centres are weighted means of rows: |c| <= max |x|
  check(cudaMemcpyAsync(d_cen_.p, h, (size_t)K * cld * es, cudaMemcpyHostToDevice, stream_), "H2D centres");
  check(cudaMemcpyAsync(d_centre_.p, hc, sizeof(double) * (size_t)D, cudaMemcpyHostToDevice, stream_), "H2D centre");
  float* d_sscale = reinterpret_cast<float*>((unsigned char*)d_iter_.p + kScaleOff);
  check(cudaMemcpyAsync(d_sscale, hs, sizeof(float), cudaMemcpyHostToDevice, stream_), "H2D scale");
  dev_view_ = &v;
  dev_K_ = K;
  list_valid_ = false;
  elist_cap_ = slist_cap_ = 0;  // list capacities are sized per view (d_list_ itself only ever grows)
  last_pairs_ = last_nnz_s_ = 0;
  if (blind) {
    // one preliminary statistics pass about the data centre gives the weighted means of the fresh clusters, which
    // then serve as their centres (the statistics are only as accurate as the centre is close to the cluster)
    double* it = (double*)d_iter_.p;
    for (int attempt = 0;; ++attempt) {
      if (attempt > 8) throw_runtime("probe statistics pass does not settle");
      check(cudaMemsetAsync(d_iter_.p, 0, kIterBytes + 32, stream_), "memset");
      dev_sphase(v);
      check(cudaMemcpyAsync(h_iter_, it, sizeof(double) * 16, cudaMemcpyDeviceToHost, stream_), "D2H iter");
      sync();
      // the abort slot of the statistics is all-reduced: every rank takes the same branch
      double ab = 0;
      check(cudaMemcpy(&ab, (double*)d_stats_.p + nstat, sizeof(double), cudaMemcpyDeviceToHost), "D2H abort");
      if (ab == 0) break;
      if (h_iter_[dev::kItOverS] != 0)
        slist_cap_ = std::min<long long>((long long)v.N * K, std::max<long long>(2 * slist_cap_, (long long)(1.5 * h_iter_[dev::kItNnzS])));
    }
    if (prec_ == kF32)
      check(dev::centres_from_stats<float>(stream_, (const double*)d_stats_.p, J, K, D, cld, (const double*)d_centre_.p, (float*)d_cen_.p), "centres");
    else
      check(dev::centres_from_stats<double>(stream_, (const double*)d_stats_.p, J, K, D, cld, (const double*)d_centre_.p, (double*)d_cen_.p), "centres");
    ++launches_;
    check(cudaMemcpyAsync(d_sscale, hs + 1, sizeof(float), cudaMemcpyHostToDevice, stream_), "H2D scale");
  }
  sync();  // the page-locked staging is reused by later calls
}

// Statistics of the stored q about the device's centres: d_stats_ = [Njk | xs | S | abort], all-reduced.
void Engine::dev_sphase(View& v) {
  const int J = v.J, K = v.K, D = v.D;
  const bool full = ckind_ == kGaussWish;
  const int64_t Sz = full ? (int64_t)D * D : D;
  const int cld = full ? dev::full_dp(D) : D;
  const int64_t nJK = (int64_t)J * K, nstat = nJK + (int64_t)K * D + K * Sz;
  const size_t es = prec_ == kF32 ? 4 : 8;
  double* d_njk = (double*)d_stats_.p;
  double* d_xs = d_njk + nJK;
  double* d_S = d_xs + (int64_t)K * D;
  double* d_abort = d_njk + nstat;
  double* it = (double*)d_iter_.p;
  unsigned* ctl = reinterpret_cast<unsigned*>((unsigned char*)d_iter_.p + kCtlOff);
  const unsigned* skipS = ctl + dev::kCtlSkipS;
  const float* d_sscale = reinterpret_cast<const float*>((unsigned char*)d_iter_.p + kScaleOff);
  (void)cld;
  check(cudaMemsetAsync(d_njk, 0, sizeof(double) * (size_t)(nstat + 1), stream_), "memset stats");
  check(cudaEventRecord(ev_[0], stream_), "event");
  const int tdim = dev::tc_dim(D, v.ldx);
  const bool tc_s = full && prec_ == kF32 && use_tc_ && use_tc_sstat_ && tdim != 0 && K <= dev::kTcCoarseMaxK;
  const bool reuse = list_valid_ && list_q_ == v.q && list_K_ == K && list_N_ == v.N && tc_s && !sparse_;
  list_valid_ = false;
  // diagonal models take the list route too (the statistics of the listed pairs only) up to 1024 dimensions
  const bool lists = full || D <= 1024;
  const bool fuse_counts = lists && !sparse_ && v.N > 0;  // nz_count / gather_list_q produce Njk in the same sweep
  if (!fuse_counts) {
    if (prec_ == kF32) check(dev::colsum<float>(stream_, (const float*)v.q, v.ldq, v.N, K, v.gid, d_njk), "colsum");
    else check(dev::colsum<double>(stream_, (const double*)v.q, v.ldq, v.N, K, v.gid, d_njk), "colsum");
    ++launches_;
  }
  const uint8_t* d_act = nullptr;
  if (sparse_) {
    // the mask needs the counts of all ranks: cluster.cpp:69-70
    allreduce(d_njk, nJK);
    reserve(d_act_, (size_t)nJK);
    check(dev::build_act(stream_, d_njk, nJK, kZeroCutoff, (uint8_t*)d_act_.p), "build_act");
    ++launches_;
    d_act = (const uint8_t*)d_act_.p;
    act_.assign(1, 1);  // "a mask exists": its content lives on the device
  } else {
    act_.clear();
  }
  cudaError_t ke = cudaSuccess;
  if (lists && v.N > 0) {
    long long* d_tot = (long long*)d_nzoff_.p;
    if (reuse) {
      // the candidate lists of the last E pass cover every non-zero of q
      long long* d_koff = d_tot + K;
      const size_t rows_bytes = (size_t)round_up((int64_t)elist_cap_ * 4, 256);
      int32_t* lrow = (int32_t*)d_list_.p;
      float* lq = (float*)((unsigned char*)d_list_.p + rows_bytes);
      check(dev::gather_list_q(stream_, sms_, (const float*)v.q, v.ldq, lrow, d_koff, d_tot, list_maxcnt_, K, lq, d_njk,
                               v.gid),
            "gather_list_q");
      ++launches_;
      ke = dev::sstat_tc128(stream_, sms_, (const float*)v.X, lrow, lq, d_koff, d_tot, list_maxcnt_, list_nnz_, K,
                            (const float*)d_cen_.p, 0.f, d_xs, d_S, (unsigned*)d_err_.p, d_sscale, skipS, tdim);
      ++launches_;
    } else {
      const int64_t nb = dev::nz_blocks(v.N);
      reserve(d_nzcnt_, sizeof(int32_t) * (size_t)nb * K);
      reserve(d_nzoff_, sizeof(long long) * (size_t)(2 * K + 6) + sizeof(int32_t) * (size_t)(K + 2));
      d_tot = (long long*)d_nzoff_.p;
      long long* d_koff = d_tot + K;
      int32_t* d_cnt = (int32_t*)d_nzcnt_.p;
      double* d_fused = fuse_counts ? d_njk : nullptr;
      if (prec_ == kF32) check(dev::nz_count<float>(stream_, (const float*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_fused), "nz_count");
      else check(dev::nz_count<double>(stream_, (const double*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_fused), "nz_count");
      check(dev::nz_scan(stream_, d_cnt, nb, K, d_tot), "nz_scan");
      // list buffer: every pair when that is small, else a few entries per row (grown after an overflow)
      const long long all = (long long)v.N * K;
      if (slist_cap_ <= 0 || slist_cap_ > all) slist_cap_ = all <= (32LL << 20) ? all : std::min(all, std::max<long long>(4 * (long long)v.N, (long long)(1.5 * last_nnz_s_)));
      if (all <= (32LL << 20)) slist_cap_ = all;
      const size_t rows_bytes = (size_t)round_up((int64_t)slist_cap_ * 4, 256);
      reserve(d_list_, rows_bytes + (size_t)slist_cap_ * es + 256);
      int32_t* lrow = (int32_t*)d_list_.p;
      void* lq = (unsigned char*)d_list_.p + rows_bytes;
      check(dev::list_plan(stream_, d_tot, K, slist_cap_, -1.0, d_koff, nullptr, nullptr, it, dev::kItNnzS,
                           dev::kItMaxCntS, dev::kItOverS, ctl, dev::kCtlSkipS, d_abort, -1),
            "list_plan");
      launches_ += 3;
      const long long nnz_hint = std::max<long long>((long long)v.N, (long long)last_nnz_s_);
      if (prec_ == kF32) {
        check(dev::nz_fill<float>(stream_, (const float*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_koff, lrow, (float*)lq,
                                  dev::kNzNonZero, skipS), "nz_fill");
        if (!full)
          ke = dev::sstat_gather_diag<float>(stream_, (const float*)v.X, D, v.ldx, lrow, (const float*)lq, d_koff, d_tot,
                                             (long long)v.N, K, (const float*)d_cen_.p, d_xs, d_S, skipS);
        else if (tc_s)
          ke = dev::sstat_tc128(stream_, sms_, (const float*)v.X, lrow, (const float*)lq, d_koff, d_tot, (long long)v.N,
                                nnz_hint, K, (const float*)d_cen_.p, 0.f, d_xs, d_S, (unsigned*)d_err_.p, d_sscale, skipS,
                                tdim);
        else
          ke = dev::sstat_gather_full<float>(stream_, (const float*)v.X, D, v.ldx, lrow, (const float*)lq, d_koff, d_tot,
                                             (long long)v.N, K, (const float*)d_cen_.p, d_xs, d_S, skipS);
      } else {
        check(dev::nz_fill<double>(stream_, (const double*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_koff, lrow, (double*)lq,
                                   dev::kNzNonZero, skipS), "nz_fill");
        if (!full)
          ke = dev::sstat_gather_diag<double>(stream_, (const double*)v.X, D, v.ldx, lrow, (const double*)lq, d_koff, d_tot,
                                              (long long)v.N, K, (const double*)d_cen_.p, d_xs, d_S, skipS);
        else
          ke = dev::sstat_gather_full<double>(stream_, (const double*)v.X, D, v.ldx, lrow, (const double*)lq, d_koff, d_tot,
                                              (long long)v.N, K, (const double*)d_cen_.p, d_xs, d_S, skipS);
      }
      launches_ += 2;
    }
  } else if (!full && v.N > 0) {
    if (prec_ == kF32)
      ke = dev::sstat_diag<float>(stream_, (const float*)v.X, v.N, D, v.ldx, v.gid, (const float*)v.q, v.ldq, K,
                                  (const float*)d_cen_.p, d_act, d_xs, d_S);
    else
      ke = dev::sstat_diag<double>(stream_, (const double*)v.X, v.N, D, v.ldx, v.gid, (const double*)v.q, v.ldq, K,
                                   (const double*)d_cen_.p, d_act, d_xs, d_S);
    ++launches_;
  }
  check(ke, "sstat kernel");
  check(cudaEventRecord(ev_[1], stream_), "event");
  if (sparse_) allreduce(d_xs, nstat - nJK + 1);
  else allreduce(d_njk, nstat + 1);
}

// Two-level E pass with the list bookkeeping on the device (Engine::ephase_two_level is the
// host-planned form).
void Engine::dev_two_level(View& v, int K, const TcLayout& lay, bool first_try) {
  (void)first_try;
  const float kMargin = 24.f;  // pairs that cannot reach e^-24 of the row's best get q = 0
  double* it = (double*)d_iter_.p;
  unsigned* ctl = reinterpret_cast<unsigned*>((unsigned char*)d_iter_.p + kCtlOff);
  const unsigned* skip = ctl + dev::kCtlSkipE;  // [0] abort / failed M step, [1] dense kernel instead
  if (v.xnorm == nullptr && v.N > 0) {
    cudaError_t e = cudaMalloc((void**)&v.xnorm, sizeof(float) * (size_t)v.N);
    if (e != cudaSuccess) throw Error{5, std::string("cudaMalloc: ") + cudaGetErrorString(e)};
    check(dev::row_norm128(stream_, sms_, (const float*)v.X, v.N, v.xnorm, dev::tc_dim(v.D, v.ldx)), "row_norm128");
    ++launches_;
  }
  float* q = (float*)v.q;
  const int W = (K + 31) / 32;
  reserve(d_cmask_, sizeof(uint32_t) * (size_t)std::max<int64_t>(v.N, 1) * W);
  uint32_t* cmask = (uint32_t*)d_cmask_.p;
  const uint8_t* d_blob = (const uint8_t*)d_tc_.p;
  const float* df = reinterpret_cast<const float*>(d_blob + lay.off_f);
  const float* d_as = df;
  const float* d_it2 = d_as + K;
  const float* d_chat = d_it2 + K;
  const float* d_lw = d_chat + K;
  const uint8_t* d_aug = d_blob + lay.off_aug;
  const float* d_cpar = reinterpret_cast<const float*>(d_blob + lay.off_cpar);
  const uint8_t* d_act = sparse_ ? (const uint8_t*)d_act_.p : nullptr;
  unsigned* d_err = (unsigned*)d_err_.p;
  const double xspan = std::max(xabs_max_, 1e-30);
  const float sg = (float)std::ldexp(1.0, std::min(100, std::max(-100, (int)std::floor(std::log2(256.0 / xspan)))));
  for (int attempt = 0;; ++attempt) {
    check(cudaEventRecord(ev_[4], stream_), "event");
    check(dev::estep_coarse_tc128(stream_, sms_, (const float*)v.X, v.xnorm, v.N, v.gid, K, d_blob, d_aug, d_cpar, d_lw,
                                  d_act, sg, 0, kMargin, q, v.ldq, cmask, coarse_sbase_hint_, d_err, ctl + dev::kCtlAugH,
                                  skip, dev::tc_dim(v.D, v.ldx)),
          "estep_coarse_tc128 launch");
    ++launches_;
    check(cudaEventRecord(ev_[5], stream_), "event");
    if (coarse_hint_ok_ || v.N <= 0) break;
    // first launch on this engine: the kernel reports the shared-memory base it expected as a parameter
    unsigned rep[2] = {0, 0};
    check(cudaMemcpyAsync(rep, d_err, sizeof(rep), cudaMemcpyDeviceToHost, stream_), "D2H err");
    sync();
    if (!(rep[1] & 0x80000000u)) {
      coarse_hint_ok_ = true;
      break;
    }
    if (attempt > 0) throw_runtime("estep_coarse_tc128: shared-memory base does not settle");
    coarse_sbase_hint_ = rep[1] & 0x7fffffffu;
    check(cudaMemsetAsync(d_err, 0, 8, stream_), "memset");
  }
  if (tc_stage_ == 1) {
    check(dev::apply_candidate_mask(stream_, q, v.ldq, v.N, K, cmask), "apply_candidate_mask");
    for (int i = 6; i <= 8; ++i) check(cudaEventRecord(ev_[i], stream_), "event");
    return;
  }
  // candidate pairs as per-cluster row lists
  const int64_t nb = dev::nz_blocks(v.N);
  reserve(d_nzcnt_, sizeof(int32_t) * (size_t)std::max<int64_t>(nb, 1) * K);
  reserve(d_nzoff_, sizeof(long long) * (size_t)(2 * K + 6) + sizeof(int32_t) * (size_t)(K + 2));
  int32_t* d_cnt = (int32_t*)d_nzcnt_.p;
  long long* d_tot = (long long*)d_nzoff_.p;
  long long* d_koff = d_tot + K;
  long long* d_nitems = d_koff + K + 2;
  int32_t* d_itoff = (int32_t*)(d_nitems + 2);
  check(dev::mask_count(stream_, cmask, v.N, K, d_cnt, skip), "mask_count");
  check(dev::nz_scan(stream_, d_cnt, nb, K, d_tot), "nz_scan");
  // each candidate costs about three products plus a gather; level 1 cost one product for all K
  const double limit = 0.4 * (double)K * (double)v.N;
  const long long cap_max = (long long)limit + 1024;
  if (elist_cap_ <= 0) elist_cap_ = std::min<long long>(cap_max, std::max<long long>(2 * (long long)v.N + 1024, (long long)(1.5 * last_pairs_)));
  if ((long long)v.N * K <= (32LL << 20)) elist_cap_ = std::max(elist_cap_, cap_max);  // small views: room for every pair the pass may keep
  elist_cap_ = std::min(elist_cap_, std::max<long long>(cap_max, 1024));
  if (tc_stage_ != 0) elist_cap_ = std::max<long long>((long long)v.N * K, 1024);  // test stages keep every candidate
  const size_t rows_bytes = (size_t)round_up((int64_t)elist_cap_ * 4, 256);
  const int64_t items_cap = elist_cap_ / 128 + K + 1;
  reserve(d_list_, 2 * rows_bytes + 256);
  reserve(d_items_, 16 * (size_t)items_cap);
  int32_t* lrow = (int32_t*)d_list_.p;
  check(dev::list_plan(stream_, d_tot, K, elist_cap_, tc_stage_ == 0 ? limit : -1.0, d_koff, d_itoff, d_nitems, it,
                       dev::kItPairs, dev::kItMaxCnt, dev::kItOverE, ctl, dev::kCtlSkipL, nullptr, dev::kItRerun),
        "list_plan");
  check(dev::mask_fill(stream_, cmask, v.N, K, d_cnt, d_koff, lrow, skip), "mask_fill");
  launches_ += 4;
  check(cudaEventRecord(ev_[6], stream_), "event");
  check(dev::estep_tc128_list(stream_, sms_, (const float*)v.X, v.N, v.gid, K, d_blob, d_as, d_it2, d_chat, d_lw, lrow,
                              d_koff, d_tot, d_itoff, items_cap, d_items_.p, q, v.ldq, d_err, d_nitems, skip,
                              dev::tc_dim(v.D, v.ldx)),
        "estep_tc128_list launch");
  launches_ += 2;
  check(cudaEventRecord(ev_[7], stream_), "event");
  if (tc_stage_ == 2) {
    check(dev::apply_candidate_mask(stream_, q, v.ldq, v.N, K, cmask), "apply_candidate_mask");
    check(cudaEventRecord(ev_[8], stream_), "event");
    return;
  }
  double* d_H = nullptr;
  if (want_scores_ && !sparse_) {
    reserve(d_score_, sizeof(double) * (size_t)K);
    d_H = (double*)d_score_.p;
    check(cudaMemsetAsync(d_H, 0, sizeof(double) * (size_t)K, stream_), "memset scores");
  }
  check(dev::estep_finalize(stream_, sms_, q, v.ldq, v.N, K, cmask, it + dev::kItSumLogZ, skip, d_H), "estep_finalize");
  ++launches_;
  check(cudaEventRecord(ev_[8], stream_), "event");
}

void Engine::dev_iteration(View& v, const std::vector<WeightPost>& weights, double* F) {
  const int J = v.J, K = v.K, D = v.D;
  const bool full = ckind_ == kGaussWish;
  const int cld = full ? dev::full_dp(D) : D;
  const size_t es = prec_ == kF32 ? 4 : 8;
  const int64_t Sz = full ? (int64_t)D * D : D;
  const int64_t nJK = (int64_t)J * K, nstat = nJK + (int64_t)K * D + K * Sz;
  const int tdim = dev::tc_dim(D, v.ldx);
  const bool tc = prec_ == kF32 && full && use_tc_ && tdim != 0;
  double* it = (double*)d_iter_.p;
  unsigned* ctl = reinterpret_cast<unsigned*>((unsigned char*)d_iter_.p + kCtlOff);
  const unsigned* skipE = ctl + dev::kCtlSkipE;
  unsigned* d_err = (unsigned*)d_err_.p;
  double rec[24];
  bool try_two = false, two_done = false;
  score_valid_ = false;
  TcLayout lay{};
  for (double& x : estep_detail_) x = 0;
  for (int attempt = 0;; ++attempt) {
    if (attempt > 8) throw_runtime("device iteration does not settle");
    check(cudaMemsetAsync(d_iter_.p, 0, kIterBytes + 32, stream_), "memset");
    check(cudaMemsetAsync(d_err, 0, 16, stream_), "memset");
    dev_sphase(v);

    // ---- M step on the device: posteriors and the operands of the E pass ----
    dev::MStepArgs a{};
    a.J = J;
    a.K = K;
    a.D = D;
    a.ckind = ckind_;
    a.wkind = wkind_;
    a.cld = cld;
    a.prior = prior_;
    a.Fp = 0;
    if (full)
      for (int l = 1; l <= D; ++l) a.Fp += std::lgamma(((double)D + 1 - l) / 2);
    a.a1p = weights[0].prior1();
    a.a2p = weights[0].prior2();
    a.Fwp = weights[0].prior_fenergy();
    a.xabs_max = xabs_max_;
    a.ntot = (double)v_ntot_;
    a.nstat = nstat;
    a.stats = (const double*)d_stats_.p;
    a.act = sparse_ ? (const uint8_t*)d_act_.p : nullptr;
    a.centre = (const double*)d_centre_.p;
    a.cen = d_cen_.p;
    a.raw = (double*)d_raw_.p;
    a.post = (double*)d_post_.p;
    a.work = (double*)d_work_.p;
    a.iter = it;
    a.ctl = ctl;
    a.sscale = reinterpret_cast<float*>((unsigned char*)d_iter_.p + kScaleOff);
    a.wscr = (double*)d_wscr_.p;
    unsigned char* opbase = nullptr;
    size_t oM = 0, oL = 0, oC = 0, oW = 0;
    if (tc) {
      try_two = use_two_level_ && K >= 8 && K <= dev::kTcCoarseMaxK && v_ntot_ >= 1024 && (two_level_skip_ == 0 || tc_stage_ != 0);
      if (!try_two && two_level_skip_ > 0) --two_level_skip_;
      lay = tc_layout(J, K, try_two);
      reserve(d_tc_, lay.total);
      reserve(d_vaug_, sizeof(double) * (size_t)K * D);
      uint8_t* base = (uint8_t*)d_tc_.p;
      float* df = reinterpret_cast<float*>(base + lay.off_f);
      a.path = 1;
      a.two_level = try_two ? 1 : 0;
      const double xspan = std::max(xabs_max_, 1e-30);
      a.sg = std::ldexp(1.0, std::min(100, std::max(-100, (int)std::floor(std::log2(256.0 / xspan)))));
      a.blob = base;
      a.as = df;
      a.it2 = df + K;
      a.chatf = df + 2 * (size_t)K;
      a.lwf = df + 3 * (size_t)K;
      a.aug = base + lay.off_aug;
      a.cpar = reinterpret_cast<float*>(base + lay.off_cpar);
      a.vaug = (double*)d_vaug_.p;
    } else {
      const size_t nR = full ? (size_t)K * cld * cld : (size_t)K * D;
      const size_t nM = (size_t)K * cld;
      oM = nR;
      oL = nR + nM;
      oC = nR + 2 * nM;
      oW = oC + K;
      reserve(d_RT_, (oW + (size_t)J * K) * es);
      opbase = (unsigned char*)d_RT_.p;
      a.path = 0;
      a.RT = opbase;
      a.mhi = opbase + oM * es;
      a.mlo = opbase + oL * es;
      a.chat = opbase + oC * es;
      a.lw = opbase + oW * es;
    }
    if (prec_ == kF32) check(dev::mstep<float>(stream_, a), "mstep");
    else check(dev::mstep<double>(stream_, a), "mstep");
    launches_ += 3;

    // ---- E pass ----
    two_done = false;
    check(cudaEventRecord(ev_[2], stream_), "event");
    const uint8_t* d_act = sparse_ ? (const uint8_t*)d_act_.p : nullptr;
    if (tc) {
      if (try_two) {
        dev_two_level(v, K, lay, attempt == 0);
        two_done = true;
      } else {
        check(dev::estep_tc128(stream_, sms_, (const float*)v.X, v.N, v.gid, K, a.blob, a.as, a.it2, a.chatf, a.lwf, d_act,
                               (float*)v.q, v.ldq, it + dev::kItSumLogZ, d_err, skipE, tdim),
              "estep_tc128 launch");
        ++launches_;
      }
    } else {
      cudaError_t ke;
      double* d_fz = it + dev::kItSumLogZ;
      double* d_H = it + 24;  // unused split scores of this mode
      if (prec_ == kF32) {
        ke = full ? dev::estep_full<float>(stream_, sms_, (const float*)v.X, v.N, D, v.ldx, v.gid, K, (const float*)a.RT,
                                            (const float*)a.mhi, (const float*)a.mlo, (const float*)a.chat,
                                            (const float*)a.lw, d_act, (float*)v.q, v.ldq, dev::kEWrite, d_fz, d_H, skipE)
                  : dev::estep_diag<float>(stream_, sms_, (const float*)v.X, v.N, D, v.ldx, v.gid, K, (const float*)a.RT,
                                            (const float*)a.mhi, (const float*)a.mlo, (const float*)a.chat,
                                            (const float*)a.lw, d_act, (float*)v.q, v.ldq, dev::kEWrite, d_fz, d_H, skipE);
      } else {
        ke = full ? dev::estep_full<double>(stream_, sms_, (const double*)v.X, v.N, D, v.ldx, v.gid, K, (const double*)a.RT,
                                             (const double*)a.mhi, (const double*)a.mlo, (const double*)a.chat,
                                             (const double*)a.lw, d_act, (double*)v.q, v.ldq, dev::kEWrite, d_fz, d_H, skipE)
                  : dev::estep_diag<double>(stream_, sms_, (const double*)v.X, v.N, D, v.ldx, v.gid, K, (const double*)a.RT,
                                             (const double*)a.mhi, (const double*)a.mlo, (const double*)a.chat,
                                             (const double*)a.lw, d_act, (double*)v.q, v.ldq, dev::kEWrite, d_fz, d_H, skipE);
      }
      if (ke == cudaErrorInvalidValue) throw_invalid("the CUDA-core E-step kernels take at most 512 clusters and 256 dimensions (full covariance); beyond "
                                                   "that only D = 128 or 64 in LCB_F32 (tensor-core tier) is supported");
      check(ke, "estep kernel");
      ++launches_;
    }
    check(cudaEventRecord(ev_[3], stream_), "event");
    allreduce2(it + dev::kItSumLogZ, it + kReduced, 2);
    check(cudaMemcpyAsync(h_iter_, it, sizeof(double) * 24, cudaMemcpyDeviceToHost, stream_), "D2H iteration record");
    sync();
    std::memcpy(rec, h_iter_, sizeof(rec));
    if (rec[dev::kItAbort] != 0) {
      // some rank's non-zero lists did not fit: nothing was written; grow (if it was this rank) and repeat
      if (rec[dev::kItOverS] != 0)
        slist_cap_ = std::min<long long>((long long)v.N * K,
                                         std::max<long long>(2 * slist_cap_, (long long)(1.5 * rec[dev::kItNnzS])));
      continue;
    }
    break;
  }
  if (rec[dev::kItMFail] == 1) throw_domain("Matrix A is not positive definite.");
  if (rec[dev::kItMFail] == 2) throw_invalid("Calc log(L): Variance is zero or less!");
  last_nnz_s_ = rec[dev::kItNnzS];
  if (rec[kReduced + 1] > 0) {
    // some rank's two-level pass gave up: those ranks run the dense kernel, everybody repeats the small all-reduce
    if (try_two && (rec[dev::kItOverE] != 0 || rec[dev::kItAugFail] != 0)) {
      two_done = false;
      const double limit = 0.4 * (double)K * (double)v.N;
      if (rec[dev::kItAugFail] != 0 || rec[dev::kItPairs] > limit) two_level_skip_ = 8;
      else elist_cap_ = std::min<long long>((long long)limit + 1024, std::max<long long>(2 * elist_cap_, (long long)(1.5 * rec[dev::kItPairs])));
      estep_detail_[5] = 2;
      check(cudaMemsetAsync(it, 0, 2 * sizeof(double), stream_), "memset");
      const uint8_t* base = (const uint8_t*)d_tc_.p;
      const float* df = reinterpret_cast<const float*>(base + lay.off_f);
      const uint8_t* d_act = sparse_ ? (const uint8_t*)d_act_.p : nullptr;
      check(dev::estep_tc128(stream_, sms_, (const float*)v.X, v.N, v.gid, K, base, df, df + K, df + 2 * (size_t)K,
                             df + 3 * (size_t)K, d_act, (float*)v.q, v.ldq, it + dev::kItSumLogZ, d_err, skipE, tdim),
            "estep_tc128 launch");
      ++launches_;
      check(cudaEventRecord(ev_[3], stream_), "event");
    }
    allreduce2(it + dev::kItSumLogZ, it + kReduced, 2);
    check(cudaMemcpyAsync(h_iter_, it + kReduced, sizeof(double) * 2, cudaMemcpyDeviceToHost, stream_), "D2H Fz");
    sync();
    rec[kReduced] = h_iter_[0];
  }
  if (tc && try_two) {
    estep_detail_[4] = rec[dev::kItPairs];
    estep_detail_[6] = rec[dev::kItItems];
    if (two_done) {
      estep_detail_[5] = 1;
      last_pairs_ = rec[dev::kItPairs];
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev_[4], ev_[5]) == cudaSuccess) estep_detail_[0] = ms;
      if (cudaEventElapsedTime(&ms, ev_[5], ev_[6]) == cudaSuccess) estep_detail_[1] = ms;
      if (cudaEventElapsedTime(&ms, ev_[6], ev_[7]) == cudaSuccess) estep_detail_[2] = ms;
      if (cudaEventElapsedTime(&ms, ev_[7], ev_[8]) == cudaSuccess) estep_detail_[3] = ms;
      cudaGetLastError();
      if (want_scores_ && !sparse_ && tc_stage_ == 0) {
        // the split ranking of this model on these responsibilities came with the soft-max pass
        score_valid_ = true;
        score_q_ = v.q;
        score_K_ = K;
        score_cbar_ = rec[dev::kItCbar];
      }
      if (tc_stage_ == 0 && !sparse_ && rec[dev::kItPairs] > 0) {
        list_valid_ = true;
        list_q_ = v.q;
        list_K_ = K;
        list_N_ = v.N;
        list_nnz_ = (long long)rec[dev::kItPairs];
        list_maxcnt_ = (long long)rec[dev::kItMaxCnt];
      }
    } else {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev_[4], ev_[5]) == cudaSuccess) estep_detail_[0] = ms;
      cudaGetLastError();
    }
  }
  // unsigned err word of the tensor-core kernels (TMEM allocation, barrier time-outs)
  const double Fz = -(rec[kReduced] + (double)v_ntot_ * rec[dev::kItCbar]);
  *F = rec[dev::kItFc] + rec[dev::kItFw] + Fz;
}

// weights / clusters on the host from the device's raw statistics (what Engine::sphase + ClusterPost::update leave)
void Engine::dev_sync_host(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters) {
  const int J = v.J, K = v.K, D = v.D;
  const bool full = ckind_ == kGaussWish;
  const int64_t Sz = full ? (int64_t)D * D : D;
  const size_t per = (size_t)(1 + D + Sz);
  double* h = (double*)pinned(sizeof(double) * ((size_t)K * per + (size_t)J * K));
  check(cudaMemcpyAsync(h, d_raw_.p, sizeof(double) * (size_t)K * per, cudaMemcpyDeviceToHost, stream_), "D2H raw statistics");
  check(cudaMemcpyAsync(h + (size_t)K * per, d_stats_.p, sizeof(double) * (size_t)J * K, cudaMemcpyDeviceToHost, stream_), "D2H Njk");
  sync();
  const double* Njk = h + (size_t)K * per;
  for (int j = 0; j < J && j < (int)weights.size(); ++j) weights[j].update(Njk + (size_t)j * K, K);
  int bad = 0;
  Error first{0, ""};
#pragma omp parallel for schedule(dynamic) num_threads(host_threads_) if (K >= 8)
  for (int k = 0; k < K; ++k) {
    try {
      const double* r = h + (size_t)k * per;
      clusters[k].clearobs();
      clusters[k].set_stats(r[0], r + 1, r + 1 + D);
      clusters[k].update();
    } catch (const Error& e) {
#pragma omp critical
      if (!bad) {
        bad = 1;
        first = e;
      }
    }
  }
  host_stale_ = false;
  if (bad) throw first;
}

}  // namespace lcb
