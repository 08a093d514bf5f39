// c_api.cpp -- extern "C" boundary declared in include/libcluster_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>
#include <exception>
#include <new>
#include <string>

#include "../../include/libcluster_b200.h"
#include "engine.hpp"
#include "tc_kernels.cuh"
#include "mstep.cuh"
#include "host_model.hpp"

using namespace lcb;

struct lcb_engine { Engine* e; };
struct lcb_weights { WeightPost* w; };
struct lcb_cluster { ClusterPost* c; };

static thread_local std::string g_last_error;

template <typename F> static int guard(F&& f) {
  try {
    f();
    return LCB_OK;
  } catch (const Error& e) {
    g_last_error = e.what;
    return e.status;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of host memory";
    return LCB_ENOMEM;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return LCB_ERUNTIME;
  } catch (...) {
    g_last_error = "unknown error";
    return LCB_ERUNTIME;
  }
}
static int bad(const char* m) {
  g_last_error = m;
  return LCB_EINVAL;
}

extern "C" {

const char* lcb_last_error(void) { return g_last_error.c_str(); }
const char* lcb_version(void) { return "libcluster_b200 0.1 (sm_100a)"; }
int lcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int lcb_create(lcb_engine** out, int device, int precision) {
  if (!out) return bad("null output pointer");
  *out = nullptr;
  return guard([&] {
    Engine* e = new Engine(device, precision);
    *out = new lcb_engine{e};
  });
}
void lcb_destroy(lcb_engine* e) {
  if (!e) return;
  delete e->e;
  delete e;
}

int lcb_set_data(lcb_engine* e, int J, const double* const* X, const int64_t* Nj, int D, const int64_t* ld, int layout) {
  if (!e) return bad("null engine");
  if (layout != LCB_ROW_MAJOR && layout != LCB_COL_MAJOR) return bad("unknown layout");
  return guard([&] { e->e->set_data_host(J, X, Nj, D, ld, layout); });
}
int lcb_set_data_device_f32(lcb_engine* e, const float* X_dev, int64_t N, int D, int64_t ld, const int32_t* gid_dev, int J) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->set_data_device_f32(X_dev, N, D, ld, gid_dev, J); });
}
int lcb_learn(lcb_engine* e, int model, double clusterprior, double weight_prior, int maxclusters, int sparse,
              int verbose, unsigned nthreads, double* F, int* K) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->learn(model, clusterprior, weight_prior, maxclusters, sparse != 0, verbose != 0, nthreads, F, K); });
}
int lcb_model_init(lcb_engine* e, int model, double clusterprior, double weight_prior, int sparse) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->model_init(model, clusterprior, weight_prior, sparse != 0); });
}
int lcb_set_qz(lcb_engine* e, const double* q0, int K) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->set_qz(q0, K); });
}
int lcb_set_labels_device(lcb_engine* e, const int32_t* labels_dev, int K) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->set_labels_device(labels_dev, K); });
}
int lcb_vbem(lcb_engine* e, int maxit, double* F, int* iters) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->vbem_public(maxit, F, iters); });
}
int lcb_vbem_step(lcb_engine* e, double* F) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->vbem_step(F); });
}
int lcb_num_clusters(const lcb_engine* e) { return e ? e->e->num_clusters() : 0; }
int lcb_num_groups(const lcb_engine* e) { return e ? e->e->num_groups() : 0; }
int64_t lcb_num_rows(const lcb_engine* e, int j) { return e ? e->e->num_rows(j) : -1; }
int lcb_get_qz(lcb_engine* e, int j, double* out, int64_t ld, int layout) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->get_qz(j, out, ld, layout); });
}
int lcb_get_group_weights(lcb_engine* e, int j, double* Nk, double* Elogweight, double* fenergy) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->get_group_weights(j, Nk, Elogweight, fenergy); });
}
int lcb_get_cluster(lcb_engine* e, int k, double* N_s, double* x_s, double* xx_s, double* N, double* mean, double* cov,
                    double* fenergy) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->get_cluster(k, N_s, x_s, xx_s, N, mean, cov, fenergy); });
}
int lcb_trace_len(const lcb_engine* e) { return e ? (int)e->e->trace_F().size() : 0; }
int lcb_get_trace(const lcb_engine* e, double* F, int* K) {
  if (!e) return bad("null engine");
  const auto& f = e->e->trace_F();
  const auto& k = e->e->trace_K();
  if (F) std::memcpy(F, f.data(), sizeof(double) * f.size());
  if (K) std::memcpy(K, k.data(), sizeof(int) * k.size());
  return LCB_OK;
}
int lcb_get_step_timing(lcb_engine* e, double out[4]) {
  if (!e || !out) return bad("null argument");
  e->e->get_step_timing(out);
  return LCB_OK;
}
int lcb_get_estep_detail(lcb_engine* e, double out[8]) {
  if (!e || !out) return bad("null argument");
  e->e->get_estep_detail(out);
  return LCB_OK;
}
void* lcb_stream(lcb_engine* e) { return e ? (void*)e->e->stream() : nullptr; }
int lcb_selftest_host_packing(void) { return lcb::dev::tc_pack_selftest(); }
int lcb_get_step_counts(lcb_engine* e, double out[4]) {
  if (!e || !out) return bad("null argument");
  e->e->get_step_counts(out);
  return LCB_OK;
}
int lcb_selftest_stick_order(const double* counts, int n, int* order) {
  if (!counts || !order || n < 0) return bad("null argument");
  lcb::dev::sort_desc_like_std(counts, n, order);
  // the reference's own call (distributions.cpp:146): std::sort on (index, count) pairs, greater count first
  std::vector<std::pair<int, double>> ov((size_t)n);
  for (int k = 0; k < n; ++k) ov[(size_t)k] = std::make_pair(k, counts[k]);
  std::sort(ov.begin(), ov.end(),
            [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.second > b.second; });
  int diff = 0;
  for (int k = 0; k < n; ++k) diff += ov[(size_t)k].first != order[k];
  return diff;
}

int lcb_nccl_unique_id(char out[128]) {
  std::string err;
  const int rc = nccl_get_unique_id(out, &err);
  if (rc) g_last_error = err;
  return rc;
}
int lcb_comm_init_nccl(lcb_engine* e, const char id[128], int rank, int world) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->comm_init_nccl(id, rank, world); });
}
int lcb_comm_init_host(lcb_engine* e, lcb_allreduce_fn fn, void* ctx, int rank, int world) {
  if (!e) return bad("null engine");
  return guard([&] { e->e->comm_init_host(fn, ctx, rank, world); });
}

// ---- operator surface -----------------------------------------------------
int lcb_weights_create(lcb_weights** out, int kind, double prior) {
  if (!out) return bad("null output pointer");
  *out = nullptr;
  return guard([&] { *out = new lcb_weights{new WeightPost(kind, prior)}; });
}
void lcb_weights_destroy(lcb_weights* w) {
  if (!w) return;
  delete w->w;
  delete w;
}
int lcb_weights_update(lcb_weights* w, const double* Nk, int K) {
  if (!w || !Nk || K < 1) return bad("weights update: bad arguments");
  return guard([&] { w->w->update(Nk, K); });
}
int lcb_weights_size(const lcb_weights* w) { return w ? w->w->size() : 0; }
int lcb_weights_elogweight(const lcb_weights* w, double* out) {
  if (!w || !out) return bad("null argument");
  const auto& v = w->w->Elogweight();
  std::memcpy(out, v.data(), sizeof(double) * v.size());
  return LCB_OK;
}
int lcb_weights_getnk(const lcb_weights* w, double* out) {
  if (!w || !out) return bad("null argument");
  const auto& v = w->w->getNk();
  std::memcpy(out, v.data(), sizeof(double) * v.size());
  return LCB_OK;
}
double lcb_weights_fenergy(const lcb_weights* w) { return w ? w->w->fenergy() : 0.0; }

int lcb_cluster_create(lcb_cluster** out, int kind, double clustwidth, int D) {
  if (!out) return bad("null output pointer");
  *out = nullptr;
  return guard([&] { *out = new lcb_cluster{new ClusterPost(kind, clustwidth, D)}; });
}
void lcb_cluster_destroy(lcb_cluster* c) {
  if (!c) return;
  delete c->c;
  delete c;
}
int lcb_cluster_addobs(lcb_engine* e, lcb_cluster* c, const double* qZk, const double* X, int64_t N, int64_t ld, int layout) {
  if (!e || !c) return bad("null argument");
  return guard([&] { e->e->op_addobs(*c->c, qZk, X, N, ld, layout); });
}
int lcb_cluster_update(lcb_cluster* c) {
  if (!c) return bad("null cluster");
  return guard([&] { c->c->update(); });
}
int lcb_cluster_clearobs(lcb_cluster* c) {
  if (!c) return bad("null cluster");
  c->c->clearobs();
  return LCB_OK;
}
int lcb_cluster_eloglike(lcb_engine* e, const lcb_cluster* c, const double* X, int64_t N, int64_t ld, int layout, double* out) {
  if (!e || !c) return bad("null argument");
  return guard([&] { e->e->op_eloglike(*c->c, X, N, ld, layout, out); });
}
int lcb_cluster_splitobs(lcb_engine* e, const lcb_cluster* c, const double* X, int64_t N, int64_t ld, int layout, uint8_t* out) {
  if (!e || !c) return bad("null argument");
  return guard([&] { e->e->op_splitobs(*c->c, X, N, ld, layout, out); });
}
double lcb_cluster_fenergy(const lcb_cluster* c) { return c ? c->c->fenergy() : 0.0; }
double lcb_cluster_getn(const lcb_cluster* c) { return c ? c->c->getN() : 0.0; }
double lcb_cluster_getprior(const lcb_cluster* c) { return c ? c->c->getprior() : 0.0; }
int lcb_cluster_dim(const lcb_cluster* c) { return c ? c->c->dim() : 0; }
int lcb_cluster_getmean(const lcb_cluster* c, double* out) {
  if (!c || !out) return bad("null argument");
  std::memcpy(out, c->c->mean().data(), sizeof(double) * c->c->dim());
  return LCB_OK;
}
int lcb_cluster_getcov(const lcb_cluster* c, double* out) {
  if (!c || !out) return bad("null argument");
  std::vector<double> cv = c->c->cov();
  std::memcpy(out, cv.data(), sizeof(double) * cv.size());
  return LCB_OK;
}
int lcb_cluster_get_stats(const lcb_cluster* c, double* N_s, double* x_s, double* xx_s) {
  if (!c) return bad("null cluster");
  if (N_s) *N_s = c->c->N_s();
  if (x_s) std::memcpy(x_s, c->c->x_s().data(), sizeof(double) * c->c->x_s().size());
  if (xx_s) std::memcpy(xx_s, c->c->xx_s().data(), sizeof(double) * c->c->xx_s().size());
  return LCB_OK;
}
int lcb_cluster_set_stats(lcb_cluster* c, double N_s, const double* x_s, const double* xx_s) {
  if (!c || !x_s || !xx_s) return bad("null argument");
  c->c->set_stats(N_s, x_s, xx_s);
  return LCB_OK;
}

// ---- host-only iteration pieces --------------------------------------------
int64_t lcb_packed_len(int model, int J, int K, int D) {
  int wk, ck;
  try {
    model_kinds(model, &wk, &ck);
  } catch (...) {
    return -1;
  }
  return packed_len(ck, J, K, D);
}
int lcb_host_mstep(int model, double clusterprior, double weight_prior, int J, int K, int D, const double* packed,
                   double* Fparams, double* Elogweight, double* means, double* covs) {
  if (!packed || J < 1 || K < 1 || D < 1) return bad("host_mstep: bad arguments");
  return guard([&] {
    int wk, ck;
    model_kinds(model, &wk, &ck);
    const int64_t blk = stat_block(ck, D), Sz = blk - 1 - D;
    double F = 0;
    for (int j = 0; j < J; ++j) {
      WeightPost w(wk, weight_prior);
      w.update(packed + (int64_t)j * K, K);
      F += w.fenergy();
      if (Elogweight) std::memcpy(Elogweight + (int64_t)j * K, w.Elogweight().data(), sizeof(double) * K);
    }
    const double* cs = packed + (int64_t)J * K;
    for (int k = 0; k < K; ++k) {
      ClusterPost c(ck, clusterprior, D);
      const double* b = cs + (int64_t)k * blk;
      c.set_stats(b[0], b + 1, b + 1 + D);
      c.update();
      F += c.fenergy();
      if (means) std::memcpy(means + (int64_t)k * D, c.mean().data(), sizeof(double) * D);
      if (covs) {
        std::vector<double> cv = c.cov();
        std::memcpy(covs + (int64_t)k * Sz, cv.data(), sizeof(double) * Sz);
      }
    }
    if (Fparams) *Fparams = F;
  });
}
void lcb_shard_rows(int64_t N, int rank, int world, int64_t* begin, int64_t* end) {
  if (world < 1) world = 1;
  const int64_t base = N / world, rem = N % world;
  const int64_t b = rank * base + (rank < rem ? rank : rem);
  if (begin) *begin = b;
  if (end) *end = b + base + (rank < rem ? 1 : 0);
}

}  // extern "C"
