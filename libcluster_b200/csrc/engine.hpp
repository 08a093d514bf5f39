// engine.hpp -- host control flow of the VB fit over device-resident data.
//
// Mirrors src/cluster.cpp of the reference (vbem :177-239, split_gr :367-495,
// prune_clusters :505-552, cluster :564-629) but every O(N) step is a kernel
// launch from kernels.cu on this engine's stream; the host keeps only the
// K-sized posteriors (host_model.hpp) and the control decisions.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "host_model.hpp"

namespace lcb {

typedef int (*HostAllreduceFn)(double* buf, int64_t count, void* ctx);

// A set of rows resident on the device together with two responsibility
// buffers (current + candidate).
struct View {
  int64_t N = 0;  // local rows
  int D = 0;
  int64_t ldx = 0;
  int J = 1;
  void* X = nullptr;
  int32_t* gid = nullptr;
  void* q = nullptr;
  void* q2 = nullptr;
  int64_t ldq = 0;
  int K = 0;
  bool owns_x = false;
  float* xnorm = nullptr;  // |x_n| of every row (fp32 engine, D == 128; filled on first use by the two-level E step)
};

struct DeviceBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

class Engine {
 public:
  Engine(int device, int precision);
  ~Engine();

  void set_data_host(int J, const double* const* X, const int64_t* Nj, int D, const int64_t* ld, int layout);
  void set_data_device_f32(const float* X, int64_t N, int D, int64_t ld, const int32_t* gid, int J);

  void learn(int model, double prior, double wprior, int maxclusters, bool sparse, bool verbose, unsigned nthreads,
             double* F, int* K);
  void model_init(int model, double prior, double wprior, bool sparse);
  void set_qz(const double* q0, int K);
  void set_labels_device(const int32_t* labels, int K);
  void vbem_public(int maxit, double* F, int* iters);
  void vbem_step(double* F);

  int num_clusters() const { return (int)clusters_.size(); }
  int num_groups() const { return main_.J; }
  int64_t num_rows(int j) const;
  void get_qz(int j, double* out, int64_t ld, int layout);
  void get_group_weights(int j, double* Nk, double* Elogw, double* fen);
  void get_cluster(int k, double* N_s, double* x_s, double* xx_s, double* N, double* mean, double* cov, double* fen);
  const std::vector<double>& trace_F() const { return trace_F_; }
  const std::vector<int>& trace_K() const { return trace_K_; }
  void get_step_timing(double out[4]);
  void get_step_counts(double out[4]);  // launches, collectives, host synchronisations of the last vbem_step
  void get_estep_detail(double out[8]) const;
  cudaStream_t stream() const { return stream_; }

  void comm_init_nccl(const char id[128], int rank, int world);
  void comm_init_host(HostAllreduceFn fn, void* ctx, int rank, int world);

  // operator-level ops on host buffers (distributions.h surface)
  void op_addobs(ClusterPost& c, const double* qk, const double* X, int64_t N, int64_t ld, int layout);
  void op_eloglike(const ClusterPost& c, const double* X, int64_t N, int64_t ld, int layout, double* out);
  void op_splitobs(const ClusterPost& c, const double* X, int64_t N, int64_t ld, int layout, uint8_t* out);

 private:
  void upload_rows_f32(View& v, const double* const* X, const int64_t* Nj, const int64_t* ld, int J,
                       const std::vector<double>& mean);
  template <typename T> void upload_rows(View& v, const double* const* X, const int64_t* Nj, const int64_t* ld, int J,
                                         int layout, const std::vector<double>& mean);
  void free_view(View& v);
  void ensure_q(View& v, int K);
  void swap_q(View& v) { std::swap(v.q, v.q2); }
  void reserve(DeviceBuf& b, size_t bytes);
  void* pinned(size_t bytes);

  // one vbem() on a view; clusters/weights are resized like the reference does
  double vbem(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
              std::vector<std::vector<double>>& hints, int maxit, bool record, int* iters);
  void iteration(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                 std::vector<std::vector<double>>& hints, double* F);
  void sphase(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
              const std::vector<std::vector<double>>& centres);
  double ephase(View& v, const std::vector<WeightPost>& weights, const std::vector<ClusterPost>& clusters, int mode,
                std::vector<double>* H);
  double ephase_tc(View& v, const std::vector<WeightPost>& weights, const std::vector<ClusterPost>& clusters);
  bool ephase_two_level(View& v, int K, const uint8_t* d_blob, const float* d_as, const float* d_it2,
                        const float* d_chat, const float* d_lw, const uint8_t* d_act, const unsigned char* d_aug,
                        size_t aug_bytes, const float* d_cpar, float sg, int aug_exp, double* d_fz, unsigned* d_err);
  void group_counts(View& v, std::vector<double>& Njk);
  bool prune(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters);
  bool split_gr(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                std::vector<int>& tally, double F, int maxclusters);
  void build_act(const std::vector<double>& Njk, int J, int K);
  void allreduce(double* dev, int64_t count);
  void allreduce2(const double* src, double* dst, int64_t count);  // out of place (device pointers)

  // ---- device-resident M step (engine_dev.cu, mstep.cu): the default path of vbem() and vbem_step() ----
  struct TcLayout {
    size_t off_f, off_aug, off_cpar, total, naug;
  };
  TcLayout tc_layout(int J, int K, bool two) const;
  void dev_begin(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                 std::vector<std::vector<double>>& hints);
  void dev_iteration(View& v, const std::vector<WeightPost>& weights, double* F);
  void dev_sphase(View& v);
  void dev_two_level(View& v, int K, const TcLayout& lay, bool first_try);
  void dev_sync_host(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters);
  void dev_drop();          // host objects up to date, device model forgotten
  void ensure_host_model(); // getters: refresh the host objects from the device statistics if they are stale
  bool use_dev_mstep_ = true;   // LCB_HOST_MSTEP=1 keeps the posterior updates on the host (A/B checks)
  bool dev_live_ = false;       // the device holds the posterior of (dev_view_, dev_K_): its centres feed the next pass
  bool host_stale_ = false;     // weights_/clusters_ lag the device model
  const View* dev_view_ = nullptr;
  int dev_K_ = 0;
  long long elist_cap_ = 0, slist_cap_ = 0;   // capacity (entries) of the candidate lists / the non-zero lists
  double last_pairs_ = 0, last_nnz_s_ = 0;
  double* h_iter_ = nullptr;    // page-locked copy of the iteration record
  // split scores sum_n q_nk logit_nk left by the last two-level E pass of a fit (Engine::split_gr uses them instead of
  // a ranking pass over the data while they still describe the current q and model)
  bool want_scores_ = false, score_valid_ = false;
  const void* score_q_ = nullptr;
  int score_K_ = 0;
  double score_cbar_ = 0;
  DeviceBuf d_score_;
  DeviceBuf d_raw_, d_post_, d_work_, d_iter_, d_centre_, d_wscr_, d_vaug_;
  void allreduce_host(double* host, int64_t count);
  void share_host_threads();
  void check(cudaError_t e, const char* what) const;
  void sync();

  int device_, prec_, sms_;
  bool use_tc_ = true;  // tcgen05 tier of the E step (LCB_DISABLE_TC=1 forces the SIMT tier)
  bool debug_sync_ = false;   // LCB_DEBUG_SYNC=1: synchronise after every launch (names the failing kernel)
  bool use_tc_sstat_ = true;  // tcgen05 scatter of the S pass (LCB_TC_SSTAT=0: SIMT gather kernel, for A/B checks)
  // Two-level E step (one-product distances for all pairs, exact logits for the candidates only).
  // LCB_TC_TWO_LEVEL=0/1 overrides the default; LCB_TC_STAGE=coarse|refine stops after that level (tests).
  bool use_two_level_ = true;
  bool dist_mstep_ = false;     // ranks share the operand packing (NCCL only) ...
  bool dist_factor_ = false;    // ... and the factorisations of the M step
  int host_threads_ = 1;        // threads of the host-side posterior updates (engine.cu)
  int tc_stage_ = 0;            // 0 full, 1 stop after level 1, 2 stop after level 2
  int two_level_skip_ = 0;      // iterations left before the two-level path is tried again after it did not pay
  uint32_t coarse_sbase_hint_ = 1024;  // shared-memory base the level-1 kernel takes as a parameter (tc_kernels.cuh)
  bool coarse_hint_ok_ = false;
  double estep_detail_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t stream_ = nullptr;
  View main_;
  std::vector<int64_t> Nj_;       // local rows per group
  std::vector<double> centre_;    // global column mean subtracted at upload
  int64_t N_total_ = 0;           // rows over all ranks
  double xabs_max_ = 0;           // max |x - centre| over all rows and ranks (bounds the fp16 operand scale)
  void measure_absmax(const View& v);

  // model state of the current fit
  int model_ = -1, wkind_ = 0, ckind_ = 0;
  double prior_ = 1.0, wprior_ = -1.0;
  bool sparse_ = false, verbose_ = false;
  std::vector<WeightPost> weights_;
  std::vector<ClusterPost> clusters_;
  std::vector<std::vector<double>> hints_;
  std::vector<double> trace_F_;
  std::vector<int> trace_K_;
  double last_F_ = 0;
  int64_t v_ntot_ = 0;        // rows (all ranks) of the view the current vbem runs on
  bool hints_first_ = false;  // first iteration takes its statistic centres from the hints

  // device scratch
  DeviceBuf d_RT_, d_mhi_, d_mlo_, d_chat_, d_lw_, d_act_, d_cen_, d_stats_, d_small_, d_tmp_, d_mean_, d_tc_, d_nzcnt_, d_nzoff_, d_list_, d_err_, d_cmask_, d_items_;
  // per-cluster row lists left by the two-level E pass; the next statistics pass reuses them while q is untouched
  bool list_valid_ = false;
  const void* list_q_ = nullptr;
  int list_K_ = 0;
  int64_t list_N_ = 0;
  long long list_nnz_ = 0, list_maxcnt_ = 0;
  std::vector<uint8_t> act_;  // host copy of the sparse mask (J*K), empty if unused
  void* h_pin_ = nullptr;
  size_t h_pin_bytes_ = 0;

  // timing of the last step
  cudaEvent_t ev_[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double t_s_ = 0, t_e_ = 0, t_all_ = 0;
  long launches_ = 0, step_launches_ = 0;
  int collectives_ = 0, step_collectives_ = 0, syncs_ = 0, step_syncs_ = 0;

  // communication
  int rank_ = 0, world_ = 1;
  void* nccl_comm_ = nullptr;
  HostAllreduceFn host_ar_ = nullptr;
  void* host_ar_ctx_ = nullptr;
};

// NCCL resolved at run time (dlopen) so that the library loads without it.
int nccl_get_unique_id(char out[128], std::string* err);

}  // namespace lcb
