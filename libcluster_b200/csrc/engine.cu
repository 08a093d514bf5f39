// engine.cu -- see engine.hpp.
#include "engine.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <thread>

#include "kernels.cuh"
#include "tc_kernels.cuh"

namespace lcb {

namespace {
enum { kF32 = 0, kF64 = 1 };

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ---- NCCL by dlopen ------------------------------------------------------
struct NcclId { char internal[128]; };
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void**, int, NcclId, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);
struct NcclApi {
  void* h = nullptr;
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_allgather allgather = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  if (api.h == nullptr) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
      api.get_id = (fn_get_id)dlsym(api.h, "ncclGetUniqueId");
      api.init_rank = (fn_init_rank)dlsym(api.h, "ncclCommInitRank");
      api.allreduce = (fn_allreduce)dlsym(api.h, "ncclAllReduce");
      api.allgather = (fn_allgather)dlsym(api.h, "ncclAllGather");
      api.destroy = (fn_destroy)dlsym(api.h, "ncclCommDestroy");
      api.errstr = (fn_errstr)dlsym(api.h, "ncclGetErrorString");
      api.ok = api.get_id && api.init_rank && api.allreduce && api.destroy;
    }
  }
  return api;
}
const int kNcclDouble = 8, kNcclSum = 0, kNcclUint8 = 1;
}  // namespace

int nccl_get_unique_id(char out[128], std::string* err) {
  NcclApi& a = nccl();
  if (!a.ok) {
    *err = "libnccl.so.2 could not be loaded";
    return 4;
  }
  NcclId id;
  const int rc = a.get_id(&id);
  if (rc != 0) {
    *err = std::string("ncclGetUniqueId: ") + (a.errstr ? a.errstr(rc) : "error");
    return 4;
  }
  std::memcpy(out, id.internal, 128);
  return 0;
}

// ---------------------------------------------------------------- plumbing --
void Engine::check(cudaError_t e, const char* what) const {
  if (e != cudaSuccess) throw Error{4, std::string(what) + ": " + cudaGetErrorString(e)};
  if (debug_sync_ && stream_ != nullptr) {
    // LCB_DEBUG_SYNC=1: wait after every checked call, so that a failing kernel is reported under its own name,
    // together with the error word the tensor-core kernels leave (0xdead0000 | barrier address: a barrier time-out)
    const cudaError_t s = cudaStreamSynchronize(stream_);
    if (s != cudaSuccess) {
      char buf[64] = "";
      unsigned w[2] = {0, 0};
      if (d_err_.p != nullptr && cudaMemcpy(w, d_err_.p, sizeof(w), cudaMemcpyDeviceToHost) == cudaSuccess)
        std::snprintf(buf, sizeof(buf), " (err words %08x %08x)", w[0], w[1]);
      throw Error{4, std::string(what) + " [after sync]: " + cudaGetErrorString(s) + buf};
    }
  }
}
void Engine::sync() {
  ++syncs_;
  check(cudaStreamSynchronize(stream_), "stream synchronize");
}

Engine::Engine(int device, int precision) : device_(device), prec_(precision) {
  if (precision != kF32 && precision != kF64) throw_invalid("precision must be LCB_F32 or LCB_F64");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
    throw Error{4, "no CUDA device is visible: the VB engine has no CPU fallback"};
  if (device < 0 || device >= n) throw_invalid("device index out of range");
  check(cudaSetDevice(device_), "cudaSetDevice");
  cudaDeviceProp prop;
  check(cudaGetDeviceProperties(&prop, device_), "cudaGetDeviceProperties");
  sms_ = prop.multiProcessorCount;
  check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  for (auto& e : ev_) check(cudaEventCreate(&e), "cudaEventCreate");
  const char* no_tc = std::getenv("LCB_DISABLE_TC");
  use_tc_ = !(no_tc && no_tc[0] && no_tc[0] != '0');
  if (const char* ts = std::getenv("LCB_TC_SSTAT")) use_tc_sstat_ = !(ts[0] == '0');
  if (const char* ds = std::getenv("LCB_DEBUG_SYNC")) debug_sync_ = ds[0] == '1';
  if (const char* tl = std::getenv("LCB_TC_TWO_LEVEL")) {
    if (tl[0]) use_two_level_ = tl[0] != '0';
  }
  // host threads of the O(K D^3) posterior updates: the machine's cores (not OMP_NUM_THREADS, which launchers often
  // pin to 1 for unrelated reasons); LCB_HOST_THREADS overrides
  host_threads_ = (int)std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  if (const char* ht = std::getenv("LCB_HOST_THREADS")) {
    const int n = std::atoi(ht);
    if (n >= 1 && n <= 1024) host_threads_ = n;
  }
  if (const char* sg = std::getenv("LCB_TC_STAGE")) {
    if (std::strcmp(sg, "coarse") == 0) tc_stage_ = 1;
    else if (std::strcmp(sg, "refine") == 0) tc_stage_ = 2;
  }
  if (const char* hm = std::getenv("LCB_HOST_MSTEP")) use_dev_mstep_ = !(hm[0] && hm[0] != '0');
  check(cudaMallocHost((void**)&h_iter_, sizeof(double) * 64), "cudaMallocHost");
}

Engine::~Engine() {
  cudaSetDevice(device_);
  if (stream_) cudaStreamSynchronize(stream_);
  free_view(main_);
  DeviceBuf* bufs[] = {&d_RT_, &d_mhi_, &d_mlo_, &d_chat_, &d_lw_, &d_act_, &d_cen_, &d_stats_, &d_small_, &d_tmp_, &d_mean_, &d_tc_, &d_nzcnt_, &d_nzoff_, &d_list_, &d_err_, &d_cmask_, &d_items_,
                       &d_raw_, &d_post_, &d_work_, &d_iter_, &d_centre_, &d_wscr_, &d_vaug_};
  for (DeviceBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (h_pin_) cudaFreeHost(h_pin_);
  if (h_iter_) cudaFreeHost(h_iter_);
  for (auto& e : ev_)
    if (e) cudaEventDestroy(e);
  if (nccl_comm_ && nccl().ok) nccl().destroy(nccl_comm_);
  if (stream_) cudaStreamDestroy(stream_);
}

void Engine::reserve(DeviceBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return;
  check(cudaSetDevice(device_), "cudaSetDevice");
  sync();
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) throw Error{5, std::string("cudaMalloc: ") + cudaGetErrorString(e)};
  b.bytes = want;
}

void* Engine::pinned(size_t bytes) {
  if (bytes > h_pin_bytes_) {
    sync();
    if (h_pin_) cudaFreeHost(h_pin_);
    h_pin_ = nullptr;
    h_pin_bytes_ = 0;
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&h_pin_, want);
    if (e != cudaSuccess) throw Error{5, std::string("cudaMallocHost: ") + cudaGetErrorString(e)};
    h_pin_bytes_ = want;
  }
  return h_pin_;
}

void Engine::free_view(View& v) {
  list_valid_ = false;
  if (dev_view_ == &v) {
    dev_live_ = false;
    host_stale_ = false;
  }
  if (v.owns_x && v.X) cudaFree(v.X);
  if (v.owns_x && v.gid) cudaFree(v.gid);
  if (v.q) cudaFree(v.q);
  if (v.q2) cudaFree(v.q2);
  if (v.xnorm) cudaFree(v.xnorm);
  v = View();
}

static void dev_alloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes > 0 ? bytes : 16);
  if (e != cudaSuccess) throw Error{5, std::string("cudaMalloc: ") + cudaGetErrorString(e)};
}

// Make room for K responsibility columns in both buffers, keeping q's content.
void Engine::ensure_q(View& v, int K) {
  if (K <= v.ldq && v.q && v.q2) return;
  list_valid_ = false;
  const size_t es = prec_ == kF32 ? 4 : 8;
  int64_t nld = round_up(std::max<int64_t>(K, 1), 8);
  if (v.ldq > 0) nld = std::max(nld, std::min<int64_t>(2 * v.ldq, 512));
  void *nq = nullptr, *nq2 = nullptr;
  const size_t bytes = (size_t)std::max<int64_t>(v.N, 1) * nld * es;
  dev_alloc(&nq, bytes);
  dev_alloc(&nq2, bytes);
  if (v.q && v.K > 0 && v.N > 0) {
    if (prec_ == kF32)
      check(dev::copy_q<float>(stream_, (const float*)v.q, (float*)nq, v.ldq, nld, v.N, v.K, v.K), "copy_q");
    else
      check(dev::copy_q<double>(stream_, (const double*)v.q, (double*)nq, v.ldq, nld, v.N, v.K, v.K), "copy_q");
    sync();
  }
  if (v.q) cudaFree(v.q);
  if (v.q2) cudaFree(v.q2);
  v.q = nq;
  v.q2 = nq2;
  v.ldq = nld;
}

// ------------------------------------------------------------ communication --
void Engine::comm_init_nccl(const char id[128], int rank, int world) {
  if (world < 1 || rank < 0 || rank >= world) throw_invalid("bad rank/world");
  NcclApi& a = nccl();
  if (!a.ok) throw Error{4, "libnccl.so.2 could not be loaded"};
  check(cudaSetDevice(device_), "cudaSetDevice");
  NcclId nid;
  std::memcpy(nid.internal, id, 128);
  void* comm = nullptr;
  const int rc = a.init_rank(&comm, world, nid, rank);
  if (rc != 0) throw Error{4, std::string("ncclCommInitRank: ") + (a.errstr ? a.errstr(rc) : "error")};
  nccl_comm_ = comm;
  host_ar_ = nullptr;
  rank_ = rank;
  world_ = world;
  share_host_threads();
}

void Engine::comm_init_host(HostAllreduceFn fn, void* ctx, int rank, int world) {
  if (world < 1 || rank < 0 || rank >= world || fn == nullptr) throw_invalid("bad rank/world/callback");
  host_ar_ = fn;
  host_ar_ctx_ = ctx;
  rank_ = rank;
  world_ = world;
  share_host_threads();
}

// One process per GPU on one box: every rank runs the same replicated host-side updates, so the ranks share the
// cores instead of each starting hardware_concurrency() threads.
void Engine::share_host_threads() {
  // With NCCL the ranks also split the O(K D^3) half of the M step and the operand packing instead of repeating them
  // (iteration(), ephase_tc()); LCB_DIST_MSTEP=0 keeps the replicated form.
  // The factor exchange moves 8 (1 + D^2) K bytes through the host each iteration, which pays from four ranks
  // on; LCB_DIST_MSTEP=0 keeps everything replicated, =1 splits the factorisations for any world size.
  dist_mstep_ = world_ > 1 && nccl_comm_ != nullptr && nccl().allgather != nullptr;
  dist_factor_ = dist_mstep_ && world_ >= 4;
  if (const char* dm = std::getenv("LCB_DIST_MSTEP")) {
    if (dm[0] == '0') dist_mstep_ = dist_factor_ = false;
    if (dm[0] == '1') dist_factor_ = dist_mstep_;
  }
  if (std::getenv("LCB_HOST_THREADS") != nullptr || world_ <= 1) return;
  const int hw = (int)std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  host_threads_ = std::max(2, hw / std::min(world_, 8));
}

void Engine::allreduce2(const double* src, double* dst, int64_t count) {
  if (count <= 0) return;
  if (world_ == 1) {
    check(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyDeviceToDevice, stream_), "D2D");
    return;
  }
  ++collectives_;
  if (nccl_comm_) {
    const int rc = nccl().allreduce(src, dst, (size_t)count, kNcclDouble, kNcclSum, nccl_comm_, stream_);
    if (rc != 0) throw Error{4, "ncclAllReduce failed"};
    return;
  }
  std::vector<double> h((size_t)count);
  check(cudaMemcpyAsync(h.data(), src, sizeof(double) * count, cudaMemcpyDeviceToHost, stream_), "D2H allreduce");
  sync();
  if (host_ar_(h.data(), count, host_ar_ctx_) != 0) throw Error{4, "host all-reduce callback failed"};
  check(cudaMemcpyAsync(dst, h.data(), sizeof(double) * count, cudaMemcpyHostToDevice, stream_), "H2D allreduce");
  sync();
}

void Engine::allreduce(double* dev, int64_t count) {
  if (world_ == 1 || count <= 0) return;
  ++collectives_;
  if (nccl_comm_) {
    const int rc = nccl().allreduce(dev, dev, (size_t)count, kNcclDouble, kNcclSum, nccl_comm_, stream_);
    if (rc != 0) throw Error{4, "ncclAllReduce failed"};
    return;
  }
  double* h = (double*)pinned(sizeof(double) * count);
  check(cudaMemcpyAsync(h, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, stream_), "D2H allreduce");
  sync();
  if (host_ar_(h, count, host_ar_ctx_) != 0) throw Error{4, "host all-reduce callback failed"};
  check(cudaMemcpyAsync(dev, h, sizeof(double) * count, cudaMemcpyHostToDevice, stream_), "H2D allreduce");
  sync();
}

void Engine::allreduce_host(double* host, int64_t count) {
  if (world_ == 1 || count <= 0) return;
  if (nccl_comm_) {
    reserve(d_tmp_, sizeof(double) * count);
    check(cudaMemcpyAsync(d_tmp_.p, host, sizeof(double) * count, cudaMemcpyHostToDevice, stream_), "H2D");
    allreduce((double*)d_tmp_.p, count);
    check(cudaMemcpyAsync(host, d_tmp_.p, sizeof(double) * count, cudaMemcpyDeviceToHost, stream_), "D2H");
    sync();
    return;
  }
  if (host_ar_(host, count, host_ar_ctx_) != 0) throw Error{4, "host all-reduce callback failed"};
}

// ------------------------------------------------------------------- data ---
// Rows travel host -> device through two staging slots so that the copy of
// block i+1 overlaps the conversion kernel of block i.  Row-major blocks in
// page-locked caller memory are copied straight from the caller's buffer.
template <typename T>
void Engine::upload_rows(View& v, const double* const* X, const int64_t* Nj, const int64_t* ld, int J, int layout,
                         const std::vector<double>& mean) {
  const int D = v.D;
  reserve(d_mean_, sizeof(double) * D);
  double* d_mean = (double*)d_mean_.p;
  check(cudaMemcpyAsync(d_mean, mean.data(), sizeof(double) * D, cudaMemcpyHostToDevice, stream_), "H2D mean");
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(64u << 20) / (8 * (int64_t)D));
  const size_t slot_bytes = sizeof(double) * chunk_rows * D;
  void* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  double* h = (double*)pinned(2 * slot_bytes);
  auto cleanup = [&]() {
    for (int i = 0; i < 2; ++i) {
      if (d_stage[i]) cudaFree(d_stage[i]);
      if (done[i]) cudaEventDestroy(done[i]);
    }
  };
  try {
    for (int i = 0; i < 2; ++i) {
      dev_alloc(&d_stage[i], slot_bytes);
      check(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming), "event");
    }
    int64_t row0 = 0;
    int slot = 0;
    bool used[2] = {false, false};
    for (int j = 0; j < J; ++j) {
      const int64_t ldj = ld ? ld[j] : (layout == 0 ? D : Nj[j]);
      bool pinned_src = false;
      if (Nj[j] > 0 && layout == 0 && ldj == D) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, X[j]) == cudaSuccess) pinned_src = at.type == cudaMemoryTypeHost;
        else cudaGetLastError();
      }
      for (int64_t r = 0; r < Nj[j]; r += chunk_rows) {
        const int64_t rows = std::min(chunk_rows, Nj[j] - r);
        if (used[slot]) check(cudaEventSynchronize(done[slot]), "event sync");
        const double* src;
        if (pinned_src) {
          src = X[j] + r * ldj;
        } else {
          double* hs = h + (size_t)slot * chunk_rows * D;
          if (layout == 0) {
            if (ldj == D) std::memcpy(hs, X[j] + r * ldj, sizeof(double) * rows * D);
            else for (int64_t n = 0; n < rows; ++n) std::memcpy(hs + n * D, X[j] + (r + n) * ldj, sizeof(double) * D);
          } else {
            for (int d = 0; d < D; ++d) std::memcpy(hs + (int64_t)d * rows, X[j] + (int64_t)d * ldj + r, sizeof(double) * rows);
          }
          src = hs;
        }
        check(cudaMemcpyAsync(d_stage[slot], src, sizeof(double) * rows * D, cudaMemcpyHostToDevice, stream_), "H2D rows");
        check(dev::convert_rows<T>(stream_, (const double*)d_stage[slot], rows, D, layout == 0 ? D : rows, layout != 0,
                                   d_mean, (T*)v.X + (row0 + r) * v.ldx, v.ldx),
              "convert_rows");
        check(cudaEventRecord(done[slot], stream_), "event record");
        used[slot] = true;
        slot ^= 1;
      }
      row0 += Nj[j];
    }
    sync();
  } catch (...) {
    cudaStreamSynchronize(stream_);
    cleanup();
    throw;
  }
  cleanup();
}

// fp32 engine, row-major groups.  The link is the bottleneck of an upload (51 GB of fp64 for the 50M x 128 case), so
// the host's threads turn the rows into centred fp32 in page-locked staging and only those 4 bytes per value cross
// PCIe, straight to their final place in X.  The conversions rarely keep the link busy on their own; when the
// caller's buffer is page-locked, a second lane sends raw fp64 blocks from the other end of the group (converted by
// convert_rows on the device), each sized to fill the link time the last conversion left idle.  Both lanes produce
// (float)(x - mean) with the subtraction in fp64, so the resident X does not depend on the split.
void Engine::upload_rows_f32(View& v, const double* const* X, const int64_t* Nj, const int64_t* ld, int J,
                             const std::vector<double>& mean) {
  const int D = v.D;
  const int64_t ldx = v.ldx;
  constexpr double kLinkBytesPerSec = 50e9;  // page-locked H2D rate assumed when sizing the raw blocks
  constexpr int kSlots = 3;
  const int64_t crow = std::max<int64_t>(1, (int64_t)(16u << 20) / (4 * ldx));
  const size_t slot_floats = (size_t)crow * ldx;
  struct Res {
    cudaStream_t s32 = nullptr, sraw = nullptr;
    cudaEvent_t e32[kSlots] = {nullptr, nullptr, nullptr}, eraw[2] = {nullptr, nullptr};
    void* dstage[2] = {nullptr, nullptr};
    ~Res() {
      if (s32) cudaStreamSynchronize(s32);
      if (sraw) cudaStreamSynchronize(sraw);
      for (cudaEvent_t e : e32) if (e) cudaEventDestroy(e);
      for (cudaEvent_t e : eraw) if (e) cudaEventDestroy(e);
      for (void* p : dstage) if (p) cudaFree(p);
      if (s32) cudaStreamDestroy(s32);
      if (sraw) cudaStreamDestroy(sraw);
    }
  } r;
  float* hslot = (float*)pinned(kSlots * slot_floats * sizeof(float));
  reserve(d_mean_, sizeof(double) * D);
  double* d_mean = (double*)d_mean_.p;
  check(cudaMemcpyAsync(d_mean, mean.data(), sizeof(double) * D, cudaMemcpyHostToDevice, stream_), "H2D mean");
  sync();
  check(cudaStreamCreateWithFlags(&r.s32, cudaStreamNonBlocking), "stream");
  check(cudaStreamCreateWithFlags(&r.sraw, cudaStreamNonBlocking), "stream");
  for (int i = 0; i < kSlots; ++i) check(cudaEventCreateWithFlags(&r.e32[i], cudaEventDisableTiming), "event");
  for (int i = 0; i < 2; ++i) check(cudaEventCreateWithFlags(&r.eraw[i], cudaEventDisableTiming), "event");
  bool used32[kSlots] = {false, false, false}, usedraw[2] = {false, false};
  int slot = 0, rslot = 0;
  const double* mu = mean.data();
  int64_t row0 = 0;
  for (int j = 0; j < J; ++j) {
    const int64_t n = Nj[j];
    const int64_t ldj = ld ? ld[j] : D;
    bool pinned_src = false;
    if (n > 0 && ldj == D) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, X[j]) == cudaSuccess) pinned_src = at.type == cudaMemoryTypeHost;
      else cudaGetLastError();
    }
    int64_t front = 0, back = n;  // rows [front, back) of the group are still to be sent
    while (front < back) {
      // ---- fp32 lane: the block at the back of what is left ----
      const int64_t rows = std::min(crow, back - front), b0 = back - rows;
      if (used32[slot]) check(cudaEventSynchronize(r.e32[slot]), "event sync");
      float* hs = hslot + (size_t)slot * slot_floats;
      const double* src0 = X[j] + b0 * ldj;
      const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) num_threads(host_threads_)
      for (int64_t i = 0; i < rows; ++i) {
        const double* src = src0 + i * ldj;
        float* dst = hs + i * ldx;
        for (int d = 0; d < D; ++d) dst[d] = (float)(src[d] - mu[d]);
        for (int64_t d = D; d < ldx; ++d) dst[d] = 0.f;
      }
      const double tc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      check(cudaMemcpyAsync((float*)v.X + (row0 + b0) * ldx, hs, sizeof(float) * (size_t)rows * ldx, cudaMemcpyHostToDevice,
                            r.s32),
            "H2D fp32 rows");
      check(cudaEventRecord(r.e32[slot], r.s32), "event record");
      used32[slot] = true;
      slot = (slot + 1) % kSlots;
      back = b0;
      // ---- raw lane: as many fp64 rows from the front as fit into the link time that conversion left idle ----
      if (pinned_src && front < back) {
        const double idle = tc - (double)rows * ldx * 4 / kLinkBytesPerSec;
        const int64_t fit = idle > 0 ? (int64_t)(idle * kLinkBytesPerSec / (8.0 * D)) : 0;
        const int64_t rrows = std::min(std::min(fit, 2 * crow), back - front);
        if (rrows >= 256) {
          if (!r.dstage[rslot]) dev_alloc(&r.dstage[rslot], sizeof(double) * (size_t)(2 * crow) * D);
          if (usedraw[rslot]) check(cudaEventSynchronize(r.eraw[rslot]), "event sync");
          check(cudaMemcpyAsync(r.dstage[rslot], X[j] + front * ldj, sizeof(double) * (size_t)rrows * D,
                                cudaMemcpyHostToDevice, r.sraw),
                "H2D raw rows");
          check(dev::convert_rows<float>(r.sraw, (const double*)r.dstage[rslot], rrows, D, D, 0, d_mean,
                                         (float*)v.X + (row0 + front) * ldx, ldx),
                "convert_rows");
          check(cudaEventRecord(r.eraw[rslot], r.sraw), "event record");
          usedraw[rslot] = true;
          rslot ^= 1;
          front += rrows;
        }
      }
    }
    row0 += n;
  }
  check(cudaStreamSynchronize(r.s32), "upload sync");
  check(cudaStreamSynchronize(r.sraw), "upload sync");
}

void Engine::set_data_host(int J, const double* const* X, const int64_t* Nj, int D, const int64_t* ld, int layout) {
  if (J < 1 || D < 1 || X == nullptr || Nj == nullptr) throw_invalid("set_data: bad arguments");
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_live_ = false;
  host_stale_ = false;
  int64_t N = 0;
  for (int j = 0; j < J; ++j) {
    if (Nj[j] < 0 || (Nj[j] > 0 && X[j] == nullptr)) throw_invalid("set_data: bad group");
    N += Nj[j];
  }
  // A matrix of the shape of the resident one (batches streamed through one engine) reuses its device buffers:
  // cudaFree + cudaMalloc of tens of GB cost more than 100 ms, a tenth of the upload itself.
  View keep;
  if (main_.owns_x && main_.X && J == 1 && main_.J == 1 && main_.N == N && main_.D == D && N > 0) {
    keep = main_;
    if (keep.xnorm) cudaFree(keep.xnorm);  // row norms belong to the old rows
    main_ = View();
    main_.X = keep.X;
    main_.q = keep.q;
    main_.q2 = keep.q2;
    main_.ldq = keep.ldq;
    main_.owns_x = true;
    list_valid_ = false;
  } else {
    free_view(main_);
  }
  Nj_.assign(Nj, Nj + J);
  // Centre subtracted at upload: the column mean of (at most) the first 2^20
  // local rows, averaged over ranks.  Any vector near the data works -- it only
  // conditions the fp32 arithmetic -- so a sample keeps large uploads one-pass.
  std::vector<double> sums(D + 2, 0.0);
  {
    int64_t left = (int64_t)1 << 20;
    for (int j = 0; j < J && left > 0; ++j) {
      const int64_t ldj = ld ? ld[j] : (layout == 0 ? D : Nj[j]);
      const int64_t take = std::min(left, Nj[j]);
      if (layout == 0) {
        const int nb = (int)std::min<int64_t>(64, (take + 4095) / 4096);
        std::vector<double> part((size_t)nb * D, 0.0);
#pragma omp parallel for schedule(static) num_threads(host_threads_)
        for (int b = 0; b < nb; ++b) {
          double* ps = part.data() + (size_t)b * D;
          for (int64_t n = take * b / nb; n < take * (b + 1) / nb; ++n) {
            const double* row = X[j] + n * ldj;
            for (int d = 0; d < D; ++d) ps[d] += row[d];
          }
        }
        for (int b = 0; b < nb; ++b)
          for (int d = 0; d < D; ++d) sums[d] += part[(size_t)b * D + d];
      } else {
        for (int d = 0; d < D; ++d) {
          double s = 0;
          const double* col = X[j] + (int64_t)d * ldj;
          for (int64_t n = 0; n < take; ++n) s += col[n];
          sums[d] += s;
        }
      }
      sums[D] += (double)take;
      left -= take;
    }
  }
  sums[D + 1] = (double)N;
  allreduce_host(sums.data(), D + 2);
  N_total_ = (int64_t)std::llround(sums[D + 1]);
  centre_.assign(D, 0.0);
  if (sums[D] > 0)
    for (int d = 0; d < D; ++d) centre_[d] = sums[d] / sums[D];

  main_.N = N;
  main_.D = D;
  main_.J = J;
  main_.ldx = round_up(D, 4);
  main_.owns_x = true;
  const size_t es = prec_ == kF32 ? 4 : 8;
  if (main_.X == nullptr) dev_alloc(&main_.X, (size_t)std::max<int64_t>(N, 1) * main_.ldx * es);
  if (J > 1) {
    dev_alloc((void**)&main_.gid, sizeof(int32_t) * std::max<int64_t>(N, 1));
    std::vector<int32_t> g((size_t)N);
    int64_t o = 0;
    for (int j = 0; j < J; ++j)
      for (int64_t n = 0; n < Nj[j]; ++n) g[o++] = j;
    check(cudaMemcpy(main_.gid, g.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice), "H2D gid");
  }
  if (prec_ == kF32 && layout == 0) upload_rows_f32(main_, X, Nj, ld, J, centre_);
  else if (prec_ == kF32) upload_rows<float>(main_, X, Nj, ld, J, layout, centre_);
  else upload_rows<double>(main_, X, Nj, ld, J, layout, centre_);
  measure_absmax(main_);
  clusters_.clear();
  weights_.clear();
  model_ = -1;
}

void Engine::set_data_device_f32(const float* X, int64_t N, int D, int64_t ld, const int32_t* gid, int J) {
  if (N < 0 || D < 1 || ld < D || J < 1 || (J > 1 && gid == nullptr)) throw_invalid("set_data_device: bad arguments");
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_live_ = false;
  host_stale_ = false;
  free_view(main_);
  // column sums on the device, then the mean over all ranks
  reserve(d_tmp_, sizeof(double) * (D + 1));
  check(cudaMemsetAsync(d_tmp_.p, 0, sizeof(double) * (D + 1), stream_), "memset");
  check(dev::colsum_f32(stream_, X, N, D, ld, (double*)d_tmp_.p), "colsum_f32");
  std::vector<double> sums(D + 1, 0.0);
  check(cudaMemcpyAsync(sums.data(), d_tmp_.p, sizeof(double) * D, cudaMemcpyDeviceToHost, stream_), "D2H sums");
  sync();
  sums[D] = (double)N;
  allreduce_host(sums.data(), D + 1);
  N_total_ = (int64_t)std::llround(sums[D]);
  centre_.assign(D, 0.0);
  if (N_total_ > 0)
    for (int d = 0; d < D; ++d) centre_[d] = sums[d] / sums[D];

  main_.N = N;
  main_.D = D;
  main_.J = J;
  main_.ldx = round_up(D, 4);
  main_.owns_x = true;
  const size_t es = prec_ == kF32 ? 4 : 8;
  dev_alloc(&main_.X, (size_t)std::max<int64_t>(N, 1) * main_.ldx * es);
  reserve(d_tmp_, sizeof(double) * D);
  check(cudaMemcpyAsync(d_tmp_.p, centre_.data(), sizeof(double) * D, cudaMemcpyHostToDevice, stream_), "H2D mean");
  if (prec_ == kF32)
    check(dev::convert_f32<float>(stream_, X, N, D, ld, (const double*)d_tmp_.p, (float*)main_.X, main_.ldx), "convert");
  else
    check(dev::convert_f32<double>(stream_, X, N, D, ld, (const double*)d_tmp_.p, (double*)main_.X, main_.ldx), "convert");
  Nj_.assign(J, 0);
  if (J > 1) {
    dev_alloc((void**)&main_.gid, sizeof(int32_t) * std::max<int64_t>(N, 1));
    check(cudaMemcpyAsync(main_.gid, gid, sizeof(int32_t) * N, cudaMemcpyDeviceToDevice, stream_), "D2D gid");
    std::vector<int32_t> g((size_t)N);
    check(cudaMemcpyAsync(g.data(), gid, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, stream_), "D2H gid");
    sync();
    for (int64_t n = 0; n < N; ++n) {
      if (g[n] < 0 || g[n] >= J || (n > 0 && g[n] < g[n - 1])) throw_invalid("group ids must be non-decreasing in [0,J)");
      Nj_[g[n]]++;
    }
  } else {
    Nj_[0] = N;
  }
  sync();
  measure_absmax(main_);
  clusters_.clear();
  weights_.clear();
  model_ = -1;
}

int64_t Engine::num_rows(int j) const {
  if (j < 0) return main_.N;
  if (j >= (int)Nj_.size()) return -1;
  return Nj_[j];
}

// max |x| of the resident (centred) data over all ranks: no product of the fp16 operand
// scale with a data value may overflow half precision in the tensor-core scatter
void Engine::measure_absmax(const View& v) {
  reserve(d_err_, 16);
  check(cudaMemsetAsync(d_err_.p, 0, 16, stream_), "memset");
  const int64_t n = v.N * v.ldx;
  if (prec_ == kF32) check(dev::absmax<float>(stream_, (const float*)v.X, n, (unsigned*)d_err_.p + 1), "absmax");
  else check(dev::absmax<double>(stream_, (const double*)v.X, n, (unsigned*)d_err_.p + 1), "absmax");
  unsigned bits = 0;
  check(cudaMemcpyAsync(&bits, (unsigned*)d_err_.p + 1, 4, cudaMemcpyDeviceToHost, stream_), "D2H absmax");
  sync();
  float f;
  std::memcpy(&f, &bits, 4);
  std::vector<double> slots((size_t)world_, 0.0);
  slots[(size_t)rank_] = (double)f;
  allreduce_host(slots.data(), world_);
  xabs_max_ = 0;
  for (double sv : slots) xabs_max_ = std::max(xabs_max_, sv);
}

// ------------------------------------------------------------ VB iteration --
void Engine::group_counts(View& v, std::vector<double>& Njk) {
  const int J = v.J, K = v.K;
  reserve(d_stats_, sizeof(double) * (size_t)J * K);
  double* d = (double*)d_stats_.p;
  check(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)J * K, stream_), "memset");
  if (prec_ == kF32) check(dev::colsum<float>(stream_, (const float*)v.q, v.ldq, v.N, K, v.gid, d), "colsum");
  else check(dev::colsum<double>(stream_, (const double*)v.q, v.ldq, v.N, K, v.gid, d), "colsum");
  ++launches_;
  allreduce(d, (int64_t)J * K);
  Njk.resize((size_t)J * K);
  check(cudaMemcpyAsync(Njk.data(), d, sizeof(double) * (size_t)J * K, cudaMemcpyDeviceToHost, stream_), "D2H Njk");
  sync();
}

// sparse-update mask: group j contributes to / competes for cluster k only if
// Njk >= ZEROCUTOFF (cluster.cpp:69-70, :111-112)
void Engine::build_act(const std::vector<double>& Njk, int J, int K) {
  act_.assign((size_t)J * K, 1);
  if (!sparse_) {
    act_.clear();
    return;
  }
  for (size_t i = 0; i < act_.size(); ++i) act_[i] = Njk[i] >= kZeroCutoff ? 1 : 0;
  reserve(d_act_, act_.size());
  check(cudaMemcpyAsync(d_act_.p, act_.data(), act_.size(), cudaMemcpyHostToDevice, stream_), "H2D act");
}

void Engine::sphase(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                    const std::vector<std::vector<double>>& centres) {
  const int J = v.J, K = v.K, D = v.D;
  const bool full = ckind_ == kGaussWish;
  const int64_t Sz = full ? (int64_t)D * D : D;
  const int cld = full ? dev::full_dp(D) : D;  // leading dimension of the centre table
  if (full && cld == 0) throw_invalid("full-covariance models support D <= 256");
  const int64_t nJK = (int64_t)J * K, nstat = nJK + (int64_t)K * D + K * Sz;
  reserve(d_stats_, sizeof(double) * nstat);
  double* d_njk = (double*)d_stats_.p;
  double* d_xs = d_njk + nJK;
  double* d_S = d_xs + (int64_t)K * D;
  check(cudaMemsetAsync(d_njk, 0, sizeof(double) * nstat, stream_), "memset stats");

  // centres, rounded to the device type; the host un-centres with the same values
  const size_t es = prec_ == kF32 ? 4 : 8;
  std::vector<double> craw((size_t)K * D);
  {
    unsigned char* h = (unsigned char*)pinned((size_t)K * cld * es);
    std::memset(h, 0, (size_t)K * cld * es);
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
        craw[(size_t)k * D + d] = used + centre_[d];
      }
    reserve(d_cen_, (size_t)K * cld * es);
    check(cudaMemcpyAsync(d_cen_.p, h, (size_t)K * cld * es, cudaMemcpyHostToDevice, stream_), "H2D centres");
  }

  check(cudaEventRecord(ev_[0], stream_), "event");
  const bool fuse_counts = full && !sparse_ && v.N > 0;  // nz_count produces Njk in the same sweep over q
  if (!fuse_counts) {
    if (prec_ == kF32) check(dev::colsum<float>(stream_, (const float*)v.q, v.ldq, v.N, K, v.gid, d_njk), "colsum");
    else check(dev::colsum<double>(stream_, (const double*)v.q, v.ldq, v.N, K, v.gid, d_njk), "colsum");
    ++launches_;
  }
  std::vector<double> host_njk;  // sparse mode reads the counts early
  const uint8_t* d_act = nullptr;
  if (sparse_) {
    allreduce(d_njk, nJK);
    host_njk.resize((size_t)nJK);
    check(cudaMemcpyAsync(host_njk.data(), d_njk, sizeof(double) * nJK, cudaMemcpyDeviceToHost, stream_), "D2H Njk");
    sync();
    build_act(host_njk, J, K);
    d_act = (const uint8_t*)d_act_.p;
  }
  cudaError_t ke = cudaSuccess;
  if (full) {
    // statistics over the non-zero responsibilities only: per-cluster (row, q) lists, then a gathered scatter
    const bool reuse = list_valid_ && list_q_ == v.q && list_K_ == K && list_N_ == v.N && prec_ == kF32 && J == 1 &&
                       !sparse_ && use_tc_ && dev::tc_supported(D, v.ldx) && K <= dev::kTcCoarseMaxK;
    list_valid_ = false;
    if (v.N > 0 && reuse) {
      // the candidate lists of the last E pass cover every non-zero of q: gather their q (and N_k) instead of
      // sweeping q twice more
      long long* d_tot = (long long*)d_nzoff_.p;
      long long* d_koff = d_tot + K;
      const size_t rows_bytes = (size_t)round_up((int64_t)list_nnz_ * 4, 256);
      int32_t* lrow = (int32_t*)d_list_.p;
      float* lq = (float*)((unsigned char*)d_list_.p + rows_bytes);
      check(dev::gather_list_q(stream_, sms_, (const float*)v.q, v.ldq, lrow, d_koff, d_tot, list_maxcnt_, K, lq, d_njk),
            "gather_list_q");
      ++launches_;
      double cmax = 0;
      for (int k = 0; k < K; ++k)
        for (int d = 0; d < D; ++d) cmax = std::max(cmax, std::fabs(craw[(size_t)k * D + d] - centre_[d]));
      const double span = std::max(xabs_max_ + cmax, 1e-30);
      const float scale = (float)std::ldexp(1.0, std::min(100, std::max(-100, (int)std::floor(std::log2(16384.0 / span)))));
      reserve(d_err_, 16);
      ke = dev::sstat_tc128(stream_, sms_, (const float*)v.X, lrow, lq, d_koff, d_tot, list_maxcnt_, list_nnz_, K,
                            (const float*)d_cen_.p, scale, d_xs, d_S, (unsigned*)d_err_.p);
    } else if (v.N > 0) {
      const int64_t nb = dev::nz_blocks(v.N);
      reserve(d_nzcnt_, sizeof(int32_t) * (size_t)nb * K);
      reserve(d_nzoff_, sizeof(long long) * (size_t)(2 * K + 2));
      int32_t* d_cnt = (int32_t*)d_nzcnt_.p;
      long long* d_tot = (long long*)d_nzoff_.p;
      long long* d_koff = d_tot + K;
      double* d_fused = fuse_counts ? d_njk : nullptr;
      if (prec_ == kF32) check(dev::nz_count<float>(stream_, (const float*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_fused), "nz_count");
      else check(dev::nz_count<double>(stream_, (const double*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_fused), "nz_count");
      check(dev::nz_scan(stream_, d_cnt, nb, K, d_tot), "nz_scan");
      std::vector<long long> tot(K), koff(K);
      check(cudaMemcpyAsync(tot.data(), d_tot, sizeof(long long) * K, cudaMemcpyDeviceToHost, stream_), "D2H nz totals");
      sync();
      long long nnz = 0, maxcnt = 0;
      for (int k = 0; k < K; ++k) {
        koff[k] = nnz;
        nnz += tot[k];
        maxcnt = std::max(maxcnt, tot[k]);
      }
      launches_ += 2;
      if (nnz > 0) {
        check(cudaMemcpyAsync(d_koff, koff.data(), sizeof(long long) * K, cudaMemcpyHostToDevice, stream_), "H2D koff");
        const size_t rows_bytes = (size_t)round_up((int64_t)nnz * 4, 256);
        reserve(d_list_, rows_bytes + (size_t)nnz * es + 256);
        int32_t* lrow = (int32_t*)d_list_.p;
        void* lq = (unsigned char*)d_list_.p + rows_bytes;
        if (prec_ == kF32) {
          check(dev::nz_fill<float>(stream_, (const float*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_koff, lrow, (float*)lq), "nz_fill");
          if (use_tc_ && dev::tc_supported(D, v.ldx) && K <= dev::kTcCoarseMaxK) {
            // power-of-two operand scale: scale * max|x - c| <= 2^14 for every row of the data set
            double cmax = 0;
            for (int k = 0; k < K; ++k)
              for (int d = 0; d < D; ++d) cmax = std::max(cmax, std::fabs(craw[(size_t)k * D + d] - centre_[d]));
            const double span = std::max(xabs_max_ + cmax, 1e-30);
            const float scale = (float)std::ldexp(1.0, std::min(100, std::max(-100, (int)std::floor(std::log2(16384.0 / span)))));
            reserve(d_err_, 16);
            ke = dev::sstat_tc128(stream_, sms_, (const float*)v.X, lrow, (const float*)lq, d_koff, d_tot, maxcnt, nnz, K,
                                  (const float*)d_cen_.p, scale, d_xs, d_S, (unsigned*)d_err_.p);
          } else {
            ke = dev::sstat_gather_full<float>(stream_, (const float*)v.X, D, v.ldx, lrow, (const float*)lq, d_koff, d_tot,
                                               maxcnt, K, (const float*)d_cen_.p, d_xs, d_S);
          }
        } else {
          check(dev::nz_fill<double>(stream_, (const double*)v.q, v.ldq, v.N, K, v.gid, d_act, d_cnt, d_koff, lrow, (double*)lq), "nz_fill");
          ke = dev::sstat_gather_full<double>(stream_, (const double*)v.X, D, v.ldx, lrow, (const double*)lq, d_koff,
                                              d_tot, maxcnt, K, (const double*)d_cen_.p, d_xs, d_S);
        }
        ++launches_;
      }
    }
  } else if (prec_ == kF32) {
    ke = dev::sstat_diag<float>(stream_, (const float*)v.X, v.N, D, v.ldx, v.gid, (const float*)v.q, v.ldq, K,
                                (const float*)d_cen_.p, d_act, d_xs, d_S);
  } else {
    ke = dev::sstat_diag<double>(stream_, (const double*)v.X, v.N, D, v.ldx, v.gid, (const double*)v.q, v.ldq, K,
                                 (const double*)d_cen_.p, d_act, d_xs, d_S);
  }
  check(ke, "sstat kernel");
  ++launches_;
  check(cudaEventRecord(ev_[1], stream_), "event");
  if (sparse_) allreduce(d_xs, nstat - nJK);
  else allreduce(d_njk, nstat);
  double* host = (double*)pinned(sizeof(double) * (size_t)nstat);
  check(cudaMemcpyAsync(host, d_njk, sizeof(double) * nstat, cudaMemcpyDeviceToHost, stream_), "D2H stats");
  sync();

  const double* Njk = host;
  const double* xs = Njk + nJK;
  const double* S = xs + (int64_t)K * D;
  for (int j = 0; j < J; ++j) weights[j].update(Njk + (int64_t)j * K, K);
#pragma omp parallel for schedule(dynamic) num_threads(host_threads_) if (K >= 8)
  for (int k = 0; k < K; ++k) {
    double n = 0;
    for (int j = 0; j < J; ++j)
      if (act_.empty() || act_[(size_t)j * K + k]) n += Njk[(int64_t)j * K + k];
    clusters[k].add_centred_stats(n, xs + (int64_t)k * D, S + (int64_t)k * Sz, &craw[(size_t)k * D]);
  }
}

double Engine::ephase(View& v, const std::vector<WeightPost>& weights, const std::vector<ClusterPost>& clusters,
                      int mode, std::vector<double>* H) {
  const int J = v.J, K = v.K, D = v.D;
  const bool full = ckind_ == kGaussWish;
  if (mode == dev::kEWrite) list_valid_ = false;
  if (mode == dev::kEWrite && prec_ == kF32 && full && use_tc_ && dev::tc_supported(D, v.ldx) && v.N > 0)
    return ephase_tc(v, weights, clusters);
  const size_t es = prec_ == kF32 ? 4 : 8;
  const int DP = full ? dev::full_dp(D) : D;
  if (full && DP == 0) throw_invalid("full-covariance models support D <= 256");
  const size_t nR = full ? (size_t)K * DP * DP : (size_t)K * D;
  const size_t nM = (size_t)K * DP;
  // pinned staging: [R | mhi | mlo | chat | lw]
  const size_t total = (nR + 2 * nM + K + (size_t)J * K) * es;
  unsigned char* h = (unsigned char*)pinned(total);
  std::memset(h, 0, total);
  auto put = [&](size_t off_elems, size_t idx, double val) {
    if (prec_ == kF32) ((float*)h)[off_elems + idx] = (float)val;
    else ((double*)h)[off_elems + idx] = val;
  };
  std::vector<double> cc(K);
  double cbar = 0;
  for (int k = 0; k < K; ++k) {
    cc[k] = clusters[k].cconst();
    cbar += cc[k];
  }
  cbar /= K;
  if (mode == dev::kERawLogit) cbar = 0;
  std::vector<double> R;
  const size_t oM = nR, oL = nR + nM, oC = nR + 2 * nM, oW = oC + K;
  for (int k = 0; k < K; ++k) {
    clusters[k].whitener(R);
    if (full) {
      for (int i = 0; i < D; ++i)
        for (int d = 0; d <= i; ++d) put(0, ((size_t)k * DP + d) * DP + i, R[(size_t)i * D + d]);
    } else {
      for (int d = 0; d < D; ++d) put(0, (size_t)k * D + d, R[d] * R[d]);
    }
    const std::vector<double>& m = clusters[k].mean();
    for (int d = 0; d < D; ++d) {
      const double rel = m[d] - centre_[d];
      if (prec_ == kF32) {
        const float hi = (float)rel;
        ((float*)h)[oM + (size_t)k * DP + d] = hi;
        ((float*)h)[oL + (size_t)k * DP + d] = (float)(rel - (double)hi);
      } else {
        ((double*)h)[oM + (size_t)k * DP + d] = rel;
      }
    }
    put(oC, k, mode == dev::kERawLogit ? 0.0 : cc[k] - cbar);
  }
  for (int j = 0; j < J; ++j) {
    const std::vector<double>& e = weights[j].Elogweight();
    for (int k = 0; k < K; ++k) put(oW, (size_t)j * K + k, mode == dev::kERawLogit ? 0.0 : e[k]);
  }
  reserve(d_RT_, total);
  check(cudaMemcpyAsync(d_RT_.p, h, total, cudaMemcpyHostToDevice, stream_), "H2D params");
  unsigned char* base = (unsigned char*)d_RT_.p;
  const void* dR = base;
  const void* dMh = base + oM * es;
  const void* dMl = base + oL * es;
  const void* dC = base + oC * es;
  const void* dW = base + oW * es;
  // the mask belongs to the E step proper: the split ranking (cluster.cpp:401-415) and operator calls see every cluster
  const uint8_t* d_act = (sparse_ && !act_.empty() && mode == dev::kEWrite) ? (const uint8_t*)d_act_.p : nullptr;

  reserve(d_small_, sizeof(double) * (K + 2));
  double* d_fz = (double*)d_small_.p;
  double* d_H = d_fz + 1;
  check(cudaMemsetAsync(d_fz, 0, sizeof(double) * (K + 1), stream_), "memset");
  check(cudaEventRecord(ev_[2], stream_), "event");
  cudaError_t ke;
  if (prec_ == kF32) {
    ke = full ? dev::estep_full<float>(stream_, sms_, (const float*)v.X, v.N, D, v.ldx, v.gid, K, (const float*)dR,
                                        (const float*)dMh, (const float*)dMl, (const float*)dC, (const float*)dW, d_act,
                                        (float*)v.q, v.ldq, mode, d_fz, d_H)
              : dev::estep_diag<float>(stream_, sms_, (const float*)v.X, v.N, D, v.ldx, v.gid, K, (const float*)dR,
                                        (const float*)dMh, (const float*)dMl, (const float*)dC, (const float*)dW, d_act,
                                        (float*)v.q, v.ldq, mode, d_fz, d_H);
  } else {
    ke = full ? dev::estep_full<double>(stream_, sms_, (const double*)v.X, v.N, D, v.ldx, v.gid, K, (const double*)dR,
                                         (const double*)dMh, (const double*)dMl, (const double*)dC, (const double*)dW,
                                         d_act, (double*)v.q, v.ldq, mode, d_fz, d_H)
              : dev::estep_diag<double>(stream_, sms_, (const double*)v.X, v.N, D, v.ldx, v.gid, K, (const double*)dR,
                                         (const double*)dMh, (const double*)dMl, (const double*)dC, (const double*)dW,
                                         d_act, (double*)v.q, v.ldq, mode, d_fz, d_H);
  }
  if (ke == cudaErrorInvalidValue) throw_invalid("the CUDA-core E-step kernels take at most 512 clusters and 256 dimensions (full covariance); beyond "
                                                   "that only D = 128 or 64 in LCB_F32 (tensor-core tier) is supported");
  check(ke, "estep kernel");
  ++launches_;
  check(cudaEventRecord(ev_[3], stream_), "event");
  if (mode == dev::kERawLogit) return 0.0;
  allreduce(d_fz, K + 1);
  std::vector<double> out(K + 1);
  check(cudaMemcpyAsync(out.data(), d_fz, sizeof(double) * (K + 1), cudaMemcpyDeviceToHost, stream_), "D2H Fz");
  sync();
  if (H) {
    H->assign(out.begin() + 1, out.end());
    H->push_back(cbar);  // caller adds cbar * Nk
  }
  return -(out[0] + (double)v_ntot_ * cbar);
}

// E step on the tensor cores (D == 128, fp32 engine): operands are packed on the
// host as pre-swizzled fp16 hi/lo blobs with per-cluster power-of-two scales.
double Engine::ephase_tc(View& v, const std::vector<WeightPost>& weights, const std::vector<ClusterPost>& clusters) {
  const int J = v.J, K = v.K, D = v.D;
  const size_t nblob = (size_t)K * dev::kTcBlobBytes;
  const size_t nfl = 3 * (size_t)K + (size_t)J * K;  // ascale, inv_t2, chat, lw
  // the two-level path adds the aug blocks and its per-cluster constants behind the dense operands
  const bool try_two = use_two_level_ && K >= 8 && K <= dev::kTcCoarseMaxK && v.N >= 1024 && (two_level_skip_ == 0 || tc_stage_ != 0);
  if (!try_two && two_level_skip_ > 0) --two_level_skip_;
  const size_t naug = try_two ? (size_t)((K + 3) / 4) * dev::kTcAugBlockBytes : 0;
  const size_t ncpar = try_two ? 4 * (size_t)K : 0;
  const size_t off_f = nblob, off_aug = (off_f + nfl * sizeof(float) + 1023) / 1024 * 1024;
  const size_t off_cpar = off_aug + naug, total = off_cpar + ncpar * sizeof(float) + 16;
  unsigned char* h = (unsigned char*)pinned(total);
  float* f = reinterpret_cast<float*>(h + off_f);
  float* h_as = f;
  float* h_it2 = h_as + K;
  float* h_chat = h_it2 + K;
  float* h_lw = h_chat + K;
  unsigned char* h_aug = h + off_aug;
  float* h_cpar = reinterpret_cast<float*>(h + off_cpar);
  if (naug) std::memset(h_aug, 0, naug);
  std::vector<double> cc(K);
  double cbar = 0;
  for (int k = 0; k < K; ++k) {
    cc[k] = clusters[k].cconst();
    cbar += cc[k];
  }
  cbar /= K;
  // ranks pack contiguous blocks of the operand blobs and all-gather them on the device (K % world == 0)
  const bool split_pack = dist_mstep_ && K % world_ == 0 && K >= world_ && &v == &main_;
  const int own0 = split_pack ? K / world_ * rank_ : 0, own1 = split_pack ? K / world_ * (rank_ + 1) : K;
  // level-1 operand scale: s_g max|x| <= 2^8 keeps fp16(s_g x) far from saturation and its small entries normal
  const double xspan = std::max(xabs_max_, 1e-30);
  const double sg = std::ldexp(1.0, std::min(100, std::max(-100, (int)std::floor(std::log2(256.0 / xspan)))));
  std::vector<double> vaug(try_two ? (size_t)K * D : 0), tau(K), rfro(K);
  double vmax = 0;
#pragma omp parallel for schedule(dynamic) reduction(max : vmax) num_threads(host_threads_) if (K >= 8)
  for (int k = 0; k < K; ++k) {
    std::vector<double> R;
    clusters[k].whitener(R);
    // a = s (x - m): s maps the widest posterior standard deviation to ~32
    // widest posterior variance = largest diagonal entry of cov() = iW / nu
    const std::vector<double>& iW = clusters[k].iW();
    double cvmax = 0, rmax = 0;
    for (int d = 0; d < D; ++d) cvmax = std::max(cvmax, iW[(size_t)d * D + d] / clusters[k].nu());
    for (size_t i = 0; i < R.size(); ++i) rmax = std::max(rmax, std::fabs(R[i]));
    int es_ = (int)std::lround(std::log2(32.0 / std::sqrt(std::max(cvmax, 1e-300))));
    es_ = std::min(60, std::max(-60, es_));
    const double s = std::ldexp(1.0, es_);
    // b = (t / s) R: t puts the largest entry of R / s near 2^8
    int et = 8 - (int)std::ceil(std::log2(std::max(rmax / s, 1e-300)));
    et = std::min(100, std::max(-100, et));
    const double t = std::ldexp(1.0, et);
    const std::vector<double>& m = clusters[k].mean();
    std::vector<double> rel(D);
    for (int d = 0; d < D; ++d) rel[d] = m[d] - centre_[d];
    if (!split_pack || (k >= own0 && k < own1))
      dev::tc_pack_cluster(R.data(), t / s, rel.data(), s, h + (size_t)k * dev::kTcBlobBytes);
    h_as[k] = (float)s;
    h_it2[k] = (float)(1.0 / (t * t));
    h_chat[k] = (float)(cc[k] - cbar);
    if (try_two) {
      // accumulator of level 1 = s_g tau (R x - R m): the centring term and |R|_F for the error bound
      tau[k] = t / s;
      double fro = 0;
      for (int i = 0; i < D; ++i) {
        double acc = 0;
        for (int d = 0; d <= i; ++d) {
          const double r = R[(size_t)i * D + d];
          acc += r * rel[d];
          fro += r * r;
        }
        const double val = sg * tau[k] * acc;
        vaug[(size_t)k * D + i] = val;
        vmax = std::max(vmax, std::fabs(val));
      }
      rfro[k] = std::sqrt(fro);
    }
  }
  for (int j = 0; j < J; ++j) {
    const std::vector<double>& e = weights[j].Elogweight();
    for (int k = 0; k < K; ++k) h_lw[(size_t)j * K + k] = (float)e[k];
  }
  bool two_ok = try_two;
  int aug_exp = 0;
  if (try_two) {
    // A slot of the aug chunk = 2^P, B slots = -v / 2^P split three ways: |v| / 2^P <= 2^14
    aug_exp = vmax > 16384.0 ? (int)std::ceil(std::log2(vmax / 16384.0)) : 0;
    if (aug_exp > 15 || !(xabs_max_ > 0) || !std::isfinite(vmax)) two_ok = false;
  }
  if (two_ok) {
    const double p2 = std::ldexp(1.0, aug_exp);
    const double eps = 1.0625 * std::ldexp(1.0, -10);  // two fp16 roundings + fp32 accumulation of 144 terms
    for (int k = 0; k < K; ++k) {
      std::vector<double> w(D);
      for (int i = 0; i < D; ++i) w[i] = -vaug[(size_t)k * D + i] / p2;
      const double res = dev::tc_pack_aug(w.data(), k, h_aug);  // in accumulator units / 2^P
      const double unit = sg * tau[k];
      h_cpar[k] = (float)(1.0 / (unit * unit));
      h_cpar[(size_t)K + k] = (float)(eps * rfro[k] * (1.0 + 1e-6));
      h_cpar[2 * (size_t)K + k] =
          (float)((eps * rfro[k] * xabs_max_ * std::ldexp(1.0, -12) + std::sqrt((double)D) * res * p2 / unit) * (1.0 + 1e-6));
      h_cpar[3 * (size_t)K + k] = h_chat[k];
    }
  }
  reserve(d_tc_, total);
  check(cudaMemcpyAsync(d_tc_.p, h, total, cudaMemcpyHostToDevice, stream_), "H2D tc params");
  if (split_pack) {
    const size_t chunk = (size_t)(K / world_) * dev::kTcBlobBytes;
    const int rc = nccl().allgather((const unsigned char*)d_tc_.p + chunk * (size_t)rank_, d_tc_.p, chunk, kNcclUint8,
                                    nccl_comm_, stream_);
    if (rc != 0) throw Error{4, "ncclAllGather failed"};
  }
  const uint8_t* d_blob = (const uint8_t*)d_tc_.p;
  const float* df = reinterpret_cast<const float*>(d_blob + off_f);
  const float* d_as = df;
  const float* d_it2 = d_as + K;
  const float* d_chat = d_it2 + K;
  const float* d_lw = d_chat + K;
  const uint8_t* d_act = (sparse_ && !act_.empty()) ? (const uint8_t*)d_act_.p : nullptr;

  reserve(d_small_, sizeof(double) * (K + 4));
  double* d_fz = (double*)d_small_.p;
  unsigned* d_err = (unsigned*)(d_fz + 1);
  check(cudaMemsetAsync(d_fz, 0, sizeof(double) * 2, stream_), "memset");
  check(cudaEventRecord(ev_[2], stream_), "event");
  bool done = false;
  for (double& x : estep_detail_) x = 0;
  if (two_ok)
    done = ephase_two_level(v, K, d_blob, d_as, d_it2, d_chat, d_lw, d_act, d_blob + off_aug, naug,
                            reinterpret_cast<const float*>(d_blob + off_cpar), (float)sg, aug_exp, d_fz, d_err);
  if (!done) {
    check(dev::estep_tc128(stream_, sms_, (const float*)v.X, v.N, v.gid, K, d_blob, d_as, d_it2, d_chat, d_lw,
                           d_act, (float*)v.q, v.ldq, d_fz, d_err),
          "estep_tc128 launch");
    ++launches_;
  }
  check(cudaEventRecord(ev_[3], stream_), "event");
  allreduce(d_fz, 1);
  double out[2] = {0, 0};
  check(cudaMemcpyAsync(out, d_fz, sizeof(double) * 2, cudaMemcpyDeviceToHost, stream_), "D2H Fz");
  sync();
  if (estep_detail_[5] == 2) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev_[4], ev_[5]) == cudaSuccess) estep_detail_[0] = ms;  // level 1 ran, then the dense kernel
    else cudaGetLastError();
  }
  if (estep_detail_[5] == 1) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_[4], ev_[5]);
    estep_detail_[0] = ms;
    cudaEventElapsedTime(&ms, ev_[5], ev_[6]);
    estep_detail_[1] = ms;
    cudaEventElapsedTime(&ms, ev_[6], ev_[7]);
    estep_detail_[2] = ms;
    cudaEventElapsedTime(&ms, ev_[7], ev_[8]);
    estep_detail_[3] = ms;
  }
  return -(out[0] + (double)v_ntot_ * cbar);
}

// Two-level E step on the tensor cores.  Returns false when the dense kernel has to run
// instead (too many candidate pairs for the two levels to pay); q then holds scratch values.
bool Engine::ephase_two_level(View& v, int K, const uint8_t* d_blob, const float* d_as, const float* d_it2,
                              const float* d_chat, const float* d_lw, const uint8_t* d_act, const unsigned char* d_aug,
                              size_t aug_bytes, const float* d_cpar, float sg, int aug_exp, double* d_fz,
                              unsigned* d_err) {
  (void)aug_bytes;
  const float kMargin = 24.f;  // pairs that cannot reach e^-24 of the row's best get q = 0
  if (v.xnorm == nullptr) {
    dev_alloc((void**)&v.xnorm, sizeof(float) * (size_t)std::max<int64_t>(v.N, 1));
    check(dev::row_norm128(stream_, sms_, (const float*)v.X, v.N, v.xnorm), "row_norm128");
    ++launches_;
  }
  float* q = (float*)v.q;
  const int W = (K + 31) / 32;
  reserve(d_cmask_, sizeof(uint32_t) * (size_t)std::max<int64_t>(v.N, 1) * W);
  uint32_t* cmask = (uint32_t*)d_cmask_.p;
  for (int attempt = 0;; ++attempt) {
    check(cudaEventRecord(ev_[4], stream_), "event");
    check(dev::estep_coarse_tc128(stream_, sms_, (const float*)v.X, v.xnorm, v.N, v.gid, K, d_blob, d_aug, d_cpar, d_lw,
                                  d_act, sg, aug_exp, kMargin, q, v.ldq, cmask, coarse_sbase_hint_, d_err),
          "estep_coarse_tc128 launch");
    ++launches_;
    check(cudaEventRecord(ev_[5], stream_), "event");
    if (coarse_hint_ok_) break;
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
  estep_detail_[5] = 1;
  if (tc_stage_ == 1) {
    check(dev::apply_candidate_mask(stream_, q, v.ldq, v.N, K, cmask), "apply_candidate_mask");
    for (int i = 6; i <= 8; ++i) check(cudaEventRecord(ev_[i], stream_), "event");
    return true;
  }
  // candidate pairs as per-cluster row lists
  const int64_t nb = dev::nz_blocks(v.N);
  reserve(d_nzcnt_, sizeof(int32_t) * (size_t)nb * K);
  reserve(d_nzoff_, sizeof(long long) * (size_t)(2 * K + 2) + sizeof(int32_t) * (size_t)(K + 2));
  int32_t* d_cnt = (int32_t*)d_nzcnt_.p;
  long long* d_tot = (long long*)d_nzoff_.p;
  long long* d_koff = d_tot + K;
  int32_t* d_itoff = (int32_t*)(d_koff + K + 2);
  check(dev::mask_count(stream_, cmask, v.N, K, d_cnt), "mask_count");
  check(dev::nz_scan(stream_, d_cnt, nb, K, d_tot), "nz_scan");
  launches_ += 2;
  std::vector<long long> tot(K), koff(K);
  check(cudaMemcpyAsync(tot.data(), d_tot, sizeof(long long) * K, cudaMemcpyDeviceToHost, stream_), "D2H candidate totals");
  sync();
  std::vector<int32_t> itoff(K + 1);
  long long npairs = 0, nitems = 0, maxcnt = 0;
  for (int k = 0; k < K; ++k) {
    koff[k] = npairs;
    itoff[k] = (int32_t)nitems;
    npairs += tot[k];
    nitems += (tot[k] + 127) / 128;
    maxcnt = std::max(maxcnt, tot[k]);
  }
  itoff[K] = (int32_t)nitems;
  estep_detail_[4] = (double)npairs;
  estep_detail_[6] = (double)nitems;
  // each candidate costs about three products plus a gather; level 1 cost one product for all K
  if (tc_stage_ == 0 && ((double)npairs > 0.4 * (double)K * (double)v.N || nitems > 2000000000LL)) {
    estep_detail_[5] = 2;
    two_level_skip_ = 8;
    return false;
  }
  check(cudaMemcpyAsync(d_koff, koff.data(), sizeof(long long) * K, cudaMemcpyHostToDevice, stream_), "H2D koff");
  check(cudaMemcpyAsync(d_itoff, itoff.data(), sizeof(int32_t) * (K + 1), cudaMemcpyHostToDevice, stream_), "H2D itoff");
  // room for the row lists and, behind them, the responsibilities the statistics pass gathers later
  const size_t rows_bytes = (size_t)round_up((int64_t)std::max<long long>(npairs, 1) * 4, 256);
  reserve(d_list_, 2 * rows_bytes + 256);
  reserve(d_items_, 16 * (size_t)std::max<long long>(nitems, 1));
  int32_t* lrow = (int32_t*)d_list_.p;
  check(dev::mask_fill(stream_, cmask, v.N, K, d_cnt, d_koff, lrow), "mask_fill");
  ++launches_;
  check(cudaEventRecord(ev_[6], stream_), "event");
  check(dev::estep_tc128_list(stream_, sms_, (const float*)v.X, v.N, v.gid, K, d_blob, d_as, d_it2, d_chat, d_lw, lrow,
                              d_koff, d_tot, d_itoff, nitems, d_items_.p, q, v.ldq, d_err),
        "estep_tc128_list launch");
  launches_ += 2;
  check(cudaEventRecord(ev_[7], stream_), "event");
  if (tc_stage_ == 2) {
    check(dev::apply_candidate_mask(stream_, q, v.ldq, v.N, K, cmask), "apply_candidate_mask");
    check(cudaEventRecord(ev_[8], stream_), "event");
    return true;
  }
  check(dev::estep_finalize(stream_, sms_, q, v.ldq, v.N, K, cmask, d_fz), "estep_finalize");
  ++launches_;
  check(cudaEventRecord(ev_[8], stream_), "event");
  if (v.J == 1 && !sparse_ && npairs > 0) {
    list_valid_ = true;
    list_q_ = v.q;
    list_K_ = K;
    list_N_ = v.N;
    list_nnz_ = npairs;
    list_maxcnt_ = maxcnt;
  }
  return true;
}

void Engine::iteration(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                       std::vector<std::vector<double>>& hints, double* F) {
  const int K = v.K, D = v.D;
  std::vector<std::vector<double>> centres(K);
  for (int k = 0; k < K; ++k) {
    const bool have_hint = k < (int)hints.size() && (int)hints[k].size() == D;
    if (hints_first_ && have_hint) centres[k] = hints[k];
    else if (clusters[k].getN() > 0) centres[k] = clusters[k].mean();
    else if (have_hint) centres[k] = hints[k];
    else centres[k] = centre_;
  }
  hints_first_ = false;
  // Fresh clusters without a hint (caller-supplied responsibilities): one preliminary statistics pass
  // about the data centre gives their weighted means, which then serve as centres -- the statistics
  // are only as accurate as the centre is close to the cluster (section 6 of DESIGN.md).
  bool blind = false;
  for (int k = 0; k < K; ++k) {
    const bool have_hint = k < (int)hints.size() && (int)hints[k].size() == D;
    if (!(clusters[k].getN() > 0) && !have_hint) blind = true;
  }
  if (blind && K > 1) {
    std::vector<ClusterPost> probe;
    for (int k = 0; k < K; ++k) probe.emplace_back(ckind_, prior_, D);
    std::vector<WeightPost> wprobe(weights);
    sphase(v, wprobe, probe, centres);
    for (int k = 0; k < K; ++k)
      if (probe[k].N_s() > 0) {
        centres[k] = probe[k].x_s();
        for (int d = 0; d < D; ++d) centres[k][d] /= probe[k].N_s();
      }
  }
  for (int k = 0; k < K; ++k) clusters[k].clearobs();
  sphase(v, weights, clusters, centres);
  // VBM for the clusters (cluster.cpp:215-217 runs this loop under OpenMP as well)
  const bool split_factor = dist_factor_ && ckind_ == kGaussWish && K >= world_ && &v == &main_;
  {
    int bad = 0;
    Error first{0, ""};
#pragma omp parallel for schedule(dynamic) num_threads(host_threads_) if (K >= 8)
    for (int k = 0; k < K; ++k) {
      try {
        if (!split_factor) {
          clusters[k].update();
        } else {
          clusters[k].update_params();
          if (k % world_ == rank_) clusters[k].factor();
        }
      } catch (const Error& e) {
#pragma omp critical
        if (!bad) {
          bad = 1;
          first = e;
        }
      }
    }
    if (split_factor) {
      // every rank factorised K / world of the matrices: exchange {log det, L^-1} (zero elsewhere, summed), and the
      // failure count so that all ranks raise the same error
      const size_t per = 1 + (size_t)D * D, n = (size_t)K * per + 1;
      double* buf = (double*)pinned(sizeof(double) * n);
      std::memset(buf, 0, sizeof(double) * n);
      if (!bad)
        for (int k = rank_; k < K; k += world_) clusters[k].export_factor(buf + (size_t)k * per);
      buf[n - 1] = bad ? 1.0 : 0.0;
      reserve(d_tmp_, sizeof(double) * n);
      check(cudaMemcpyAsync(d_tmp_.p, buf, sizeof(double) * n, cudaMemcpyHostToDevice, stream_), "H2D factors");
      allreduce((double*)d_tmp_.p, (int64_t)n);
      check(cudaMemcpyAsync(buf, d_tmp_.p, sizeof(double) * n, cudaMemcpyDeviceToHost, stream_), "D2H factors");
      sync();
      if (buf[n - 1] > 0 && !bad) {
        bad = 1;
        first = Error{3, "Matrix A is not positive definite."};
      }
      if (!bad)
        for (int k = 0; k < K; ++k)
          if (k % world_ != rank_) clusters[k].import_factor(buf + (size_t)k * per);
    }
    if (bad) throw first;
  }
  const double Fz = ephase(v, weights, clusters, dev::kEWrite, nullptr);
  double Fw = 0, Fc = 0;
  for (size_t j = 0; j < weights.size(); ++j) Fw += weights[j].fenergy();
  std::vector<double> fck(K);
#pragma omp parallel for schedule(dynamic) num_threads(host_threads_) if (K >= 8)
  for (int k = 0; k < K; ++k) fck[k] = clusters[k].fenergy();
  for (int k = 0; k < K; ++k) Fc += fck[k];
  *F = Fc + Fw + Fz;
}

double Engine::vbem(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                    std::vector<std::vector<double>>& hints, int maxit, bool record, int* iters) {
  const int J = v.J, K = v.K;
  while ((int)weights.size() < J) weights.emplace_back(wkind_, -1.0);  // weights.resize(J, W()), cluster.cpp:192
  weights.resize(J, WeightPost(wkind_, -1.0));
  while ((int)clusters.size() > K) clusters.pop_back();
  while ((int)clusters.size() < K) clusters.emplace_back(ckind_, prior_, v.D);  // :193
  hints.resize(K);
  double F = DBL_MAX, Fold;
  int i = 0, n = 0;
  // The iterations run against the device-resident model (statistics, posteriors and E-step operands never leave
  // the GPU; the host reads one small record per iteration); the host objects are refreshed once at the end.
  // every rank must take the same path (the collectives differ): a rank without rows of this view -- the members of
  // a split candidate can all live elsewhere -- runs the device iteration on zero rows
  const bool on_device = use_dev_mstep_ && (v.N > 0 || world_ > 1);
  if (on_device) {
    dev_drop();
    dev_begin(v, weights, clusters, hints);
  }
  try {
    do {
      Fold = F;
      if (on_device) dev_iteration(v, weights, &F);
      else iteration(v, weights, clusters, hints, &F);
      ++n;
      if (record) {
        trace_F_.push_back(F);
        trace_K_.push_back(K);
      }
      if ((F - Fold) / std::fabs(Fold) > kFengyDel) throw_runtime("Free energy increase!");
      if (verbose_) std::cout << '-' << std::flush;
    } while ((std::fabs((Fold - F) / Fold) > kConverge) && ((i++ < maxit) || (maxit < 0)));
  } catch (...) {
    dev_live_ = false;
    host_stale_ = false;
    throw;
  }
  if (on_device) {
    dev_sync_host(v, weights, clusters);
    dev_live_ = false;
  }
  if (iters) *iters = n;
  return F;
}

bool Engine::prune(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters) {
  const int K = (int)clusters.size();
  std::vector<int32_t> keep;
  for (int k = 0; k < K; ++k)
    if (!(clusters[k].getN() < kZeroCutoff)) keep.push_back(k);
  if ((int)keep.size() == K) return false;
  if (verbose_) std::cout << '*' << std::flush;
  list_valid_ = false;
  score_valid_ = false;  // columns go away and the weights are updated: the ranking is recomputed
  std::vector<ClusterPost> nc;
  std::vector<std::vector<double>> nh;
  for (int32_t k : keep) {
    nc.push_back(clusters[k]);
    nh.push_back(k < (int)hints_.size() ? hints_[k] : std::vector<double>());
  }
  clusters.swap(nc);
  if (&clusters == &clusters_) hints_.swap(nh);
  const int newK = (int)keep.size();
  reserve(d_tmp_, sizeof(int32_t) * std::max(newK, 1));
  if (newK > 0)
    check(cudaMemcpyAsync(d_tmp_.p, keep.data(), sizeof(int32_t) * newK, cudaMemcpyHostToDevice, stream_), "H2D keep");
  if (prec_ == kF32) check(dev::prune_columns<float>(stream_, (float*)v.q, v.ldq, v.N, (const int32_t*)d_tmp_.p, newK), "prune");
  else check(dev::prune_columns<double>(stream_, (double*)v.q, v.ldq, v.N, (const int32_t*)d_tmp_.p, newK), "prune");
  v.K = newK;
  std::vector<double> Njk;
  group_counts(v, Njk);
  for (int j = 0; j < v.J; ++j) weights[j].update(Njk.data() + (size_t)j * newK, newK);
  return true;
}

namespace {
struct GreedOrder {
  int k, tally;
  double Fk;
};
// comutils.h:60-68
bool greedcomp(const GreedOrder& i, const GreedOrder& j) {
  if (i.tally == j.tally) return i.Fk > j.Fk;
  return i.tally < j.tally;
}
bool anyempty(const std::vector<ClusterPost>& c) {  // comutils.h:114-123
  for (const auto& x : c)
    if (x.getN() <= 1) return true;
  return false;
}
}  // namespace

bool Engine::split_gr(View& v, std::vector<WeightPost>& weights, std::vector<ClusterPost>& clusters,
                      std::vector<int>& tally, double F, int maxclusters) {
  const int K = (int)clusters.size(), D = v.D, J = v.J;
  if (K >= maxclusters && maxclusters >= 0) return false;
  tally.resize(K, 0);

  // rank clusters by their approximate free-energy contribution (cluster.cpp:386-418)
  std::vector<double> H, Njk;
  v_ntot_ = N_total_;
  double cbar;
  // every rank must take the same branch (the fallback holds a collective): agree on the validity first
  double have = (score_valid_ && score_q_ == v.q && score_K_ == K && !host_stale_) ? 1.0 : 0.0;
  allreduce_host(&have, 1);
  if (have == (double)world_) {
    // the last E pass of the fit's vbem() already summed q * logit per cluster (estep_finalize_kernel)
    H.resize(K);
    check(cudaMemcpyAsync(H.data(), d_score_.p, sizeof(double) * (size_t)K, cudaMemcpyDeviceToHost, stream_), "D2H scores");
    sync();
    allreduce_host(H.data(), K);
    cbar = score_cbar_;
  } else {
    ephase(v, weights, clusters, dev::kEScore, &H);
    cbar = H.back();
  }
  score_valid_ = false;
  group_counts(v, Njk);
  std::vector<GreedOrder> ord(K);
  for (int k = 0; k < K; ++k) {
    double nk = 0;
    for (int j = 0; j < J; ++j) nk += Njk[(size_t)j * K + k];
    ord[k].k = k;
    ord[k].tally = tally[k];
    ord[k].Fk = clusters[k].fenergy() - (H[k] + cbar * nk);
  }
  std::sort(ord.begin(), ord.end(), greedcomp);

  const size_t es = prec_ == kF32 ? 4 : 8;
  for (const GreedOrder& cand : ord) {
    const int k = cand.k;
    ++tally[k];
    if (clusters[k].getN() < 4) continue;

    // members of cluster k (q > 0.5), gathered in order (partobs, comutils.cpp:56-72)
    const int64_t nblocks = (v.N + 1023) / 1024;
    reserve(d_tmp_, sizeof(int32_t) * (size_t)std::max<int64_t>(nblocks, 1) + 64);
    int32_t* d_cnt = (int32_t*)d_tmp_.p;
    reserve(d_small_, sizeof(double) * (K + 4));
    int64_t* d_total = (int64_t*)d_small_.p;
    int64_t M = 0;
    if (v.N > 0) {
      if (prec_ == kF32) check(dev::member_counts<float>(stream_, (const float*)v.q, v.ldq, v.N, k, d_cnt), "member_counts");
      else check(dev::member_counts<double>(stream_, (const double*)v.q, v.ldq, v.N, k, d_cnt), "member_counts");
      check(dev::scan_counts(stream_, d_cnt, nblocks, d_total), "scan_counts");
      check(cudaMemcpyAsync(&M, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, stream_), "D2H M");
      sync();
    }
    View sub;
    sub.N = M;
    sub.D = D;
    sub.ldx = v.ldx;
    sub.J = J;
    sub.owns_x = true;
    void* d_map = nullptr;
    struct Guard {
      Engine* e; View* s; void** map;
      ~Guard() { if (*map) cudaFree(*map); e->free_view(*s); }
    } guard{this, &sub, &d_map};
    dev_alloc(&sub.X, (size_t)std::max<int64_t>(M, 1) * sub.ldx * es);
    if (J > 1) dev_alloc((void**)&sub.gid, sizeof(int32_t) * std::max<int64_t>(M, 1));
    dev_alloc(&d_map, sizeof(int64_t) * std::max<int64_t>(M, 1));
    ensure_q(sub, 2);
    sub.K = 2;
    if (v.N > 0) {
      if (prec_ == kF32)
        check(dev::gather_members<float>(stream_, (const float*)v.q, v.ldq, v.N, k, d_cnt, (const float*)v.X, v.ldx, D,
                                         v.gid, (float*)sub.X, sub.gid, (int64_t*)d_map), "gather");
      else
        check(dev::gather_members<double>(stream_, (const double*)v.q, v.ldq, v.N, k, d_cnt, (const double*)v.X, v.ldx,
                                          D, v.gid, (double*)sub.X, sub.gid, (int64_t*)d_map), "gather");
    }
    // initial hard split perpendicular to the principal axis (splitobs)
    std::vector<double> dir;
    clusters[k].split_direction(dir);
    const std::vector<double>& mk = clusters[k].mean();
    {
      unsigned char* h = (unsigned char*)pinned(2 * (size_t)D * es);
      for (int d = 0; d < D; ++d) {
        if (prec_ == kF32) {
          ((float*)h)[d] = (float)(mk[d] - centre_[d]);
          ((float*)h)[D + d] = (float)dir[d];
        } else {
          ((double*)h)[d] = mk[d] - centre_[d];
          ((double*)h)[D + d] = dir[d];
        }
      }
      reserve(d_cen_, 2 * (size_t)D * es);
      check(cudaMemcpyAsync(d_cen_.p, h, 2 * (size_t)D * es, cudaMemcpyHostToDevice, stream_), "H2D split dir");
    }
    unsigned long long* d_sc = (unsigned long long*)((char*)d_small_.p + 16);
    check(cudaMemsetAsync(d_sc, 0, sizeof(unsigned long long), stream_), "memset");
    if (prec_ == kF32)
      check(dev::split_side<float>(stream_, (const float*)sub.X, M, D, sub.ldx, (const float*)d_cen_.p,
                                   (const float*)d_cen_.p + D, (float*)sub.q, sub.ldq, d_sc), "split_side");
    else
      check(dev::split_side<double>(stream_, (const double*)sub.X, M, D, sub.ldx, (const double*)d_cen_.p,
                                    (const double*)d_cen_.p + D, (double*)sub.q, sub.ldq, d_sc), "split_side");
    unsigned long long sc = 0;
    check(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, stream_), "D2H scount");
    sync();
    double cnts[2] = {(double)sc, (double)M};
    allreduce_host(cnts, 2);
    const int64_t scount = (int64_t)std::llround(cnts[0]), Mtot = (int64_t)std::llround(cnts[1]);
    if (scount < 2 || scount > Mtot - 2) continue;

    // refine the split on the members only (cluster.cpp:460-465)
    std::vector<WeightPost> wspl;
    std::vector<ClusterPost> cspl;
    std::vector<std::vector<double>> hspl(2, mk);
    v_ntot_ = Mtot;
    hints_first_ = true;
    vbem(sub, wspl, cspl, hspl, kSplitIter, true, nullptr);
    if (anyempty(cspl)) continue;

    // augment the labels of the whole data set (auglabels, comutils.cpp:75-104)
    list_valid_ = false;
    ensure_q(v, K + 1);
    if (prec_ == kF32) {
      check(dev::copy_q<float>(stream_, (const float*)v.q, (float*)v.q2, v.ldq, v.ldq, v.N, K, K + 1), "copy_q");
      check(dev::aug_labels<float>(stream_, (const float*)sub.q, sub.ldq, (const int64_t*)d_map, M, (const float*)v.q,
                                   (float*)v.q2, v.ldq, k, K), "aug_labels");
    } else {
      check(dev::copy_q<double>(stream_, (const double*)v.q, (double*)v.q2, v.ldq, v.ldq, v.N, K, K + 1), "copy_q");
      check(dev::aug_labels<double>(stream_, (const double*)sub.q, sub.ldq, (const int64_t*)d_map, M,
                                    (const double*)v.q, (double*)v.q2, v.ldq, k, K), "aug_labels");
    }
    swap_q(v);  // v.q = augmented labels, v.q2 = the labels we may have to go back to
    v.K = K + 1;
    std::vector<std::vector<double>> haug(K + 1);
    for (int i = 0; i < K; ++i) haug[i] = clusters[i].mean();
    haug[K] = cspl[1].mean();
    std::vector<double> new_hint = haug[K];
    v_ntot_ = N_total_;
    hints_first_ = true;
    double Fsplit;
    try {
      Fsplit = vbem(v, wspl, cspl, haug, 1, true, nullptr);
    } catch (...) {
      swap_q(v);
      v.K = K;
      throw;
    }
    bool accept = !anyempty(cspl);
    if (accept && verbose_) std::cout << '=' << std::flush;
    accept = accept && (Fsplit < F) && (std::fabs((F - Fsplit) / F) > kConverge);
    if (accept) {
      tally[k] = 0;
      if (&clusters == &clusters_) {
        hints_.resize(K + 1);
        hints_[K] = new_hint;
      }
      return true;
    }
    swap_q(v);
    v.K = K;
  }
  return false;
}

// ------------------------------------------------------------ public fits ---
void Engine::model_init(int model, double prior, double wprior, bool sparse) {
  if (main_.X == nullptr) throw_invalid("no observations loaded (call lcb_set_data first)");
  dev_live_ = false;
  host_stale_ = false;
  model_kinds(model, &wkind_, &ckind_);
  if (!(prior > 0)) throw_invalid("clustwidth must be > 0!");
  model_ = model;
  prior_ = prior;
  wprior_ = wprior;
  sparse_ = sparse;
  weights_.clear();
  clusters_.clear();
  hints_.clear();
  act_.clear();
  for (int j = 0; j < main_.J; ++j) weights_.emplace_back(wkind_, wprior_);
  trace_F_.clear();
  trace_K_.clear();
  main_.K = 0;
}

void Engine::learn(int model, double prior, double wprior, int maxclusters, bool sparse, bool verbose,
                   unsigned nthreads, double* F, int* K) {
  if (nthreads < 1) throw_invalid("Must specify at least one thread for execution!");
  check(cudaSetDevice(device_), "cudaSetDevice");
  // nthreads bounds the host-side part of the fit, as omp_set_num_threads(nthreads) does in cluster.cpp:578
  struct ThreadsGuard {
    int& ref;
    int saved;
    ~ThreadsGuard() { ref = saved; }
  } tguard{host_threads_, host_threads_};
  host_threads_ = (int)std::max(1u, std::min((unsigned)host_threads_, nthreads));
  struct ScoreGuard {
    bool& want;
    bool& valid;
    ~ScoreGuard() { want = false; valid = false; }
  } sguard{want_scores_, score_valid_};
  want_scores_ = true;
  model_init(model, prior, wprior, sparse);
  verbose_ = verbose;
  list_valid_ = false;
  ensure_q(main_, 1);
  if (prec_ == kF32) check(dev::fill_ones<float>(stream_, (float*)main_.q, main_.ldq, main_.N), "fill_ones");
  else check(dev::fill_ones<double>(stream_, (double*)main_.q, main_.ldq, main_.N), "fill_ones");
  main_.K = 1;
  std::vector<int> tally;
  bool issplit = true;
  double Fcur = 0;
  while (issplit) {
    v_ntot_ = N_total_;
    Fcur = vbem(main_, weights_, clusters_, hints_, -1, true, nullptr);
    prune(main_, weights_, clusters_);
    if (verbose_) std::cout << '<' << std::flush;
    issplit = split_gr(main_, weights_, clusters_, tally, Fcur, maxclusters);
    if (verbose_) std::cout << '>' << std::endl;
  }
  if (verbose_) {
    std::cout << "Finished!" << std::endl;
    std::cout << "Number of clusters = " << clusters_.size() << std::endl;
    std::cout << "Free energy = " << Fcur << std::endl;
  }
  last_F_ = Fcur;
  if (F) *F = Fcur;
  if (K) *K = (int)clusters_.size();
}

void Engine::set_qz(const double* q0, int K) {
  if (model_ < 0) throw_invalid("call lcb_model_init first");
  if (K < 1 || q0 == nullptr) throw_invalid("set_qz: bad arguments");
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_live_ = false;
  host_stale_ = false;
  main_.K = 0;
  list_valid_ = false;
  ensure_q(main_, K);
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(16u << 20) / (8 * (int64_t)K));
  reserve(d_tmp_, sizeof(double) * chunk * K);
  double* h = (double*)pinned(sizeof(double) * chunk * K);
  for (int64_t r = 0; r < main_.N; r += chunk) {
    const int64_t rows = std::min(chunk, main_.N - r);
    std::memcpy(h, q0 + r * K, sizeof(double) * rows * K);
    check(cudaMemcpyAsync(d_tmp_.p, h, sizeof(double) * rows * K, cudaMemcpyHostToDevice, stream_), "H2D q0");
    if (prec_ == kF32)
      check(dev::q_from_double<float>(stream_, (const double*)d_tmp_.p, rows, K, (float*)main_.q + r * main_.ldq, main_.ldq), "q_from_double");
    else
      check(dev::q_from_double<double>(stream_, (const double*)d_tmp_.p, rows, K, (double*)main_.q + r * main_.ldq, main_.ldq), "q_from_double");
    sync();
  }
  main_.K = K;
  clusters_.clear();
  hints_.clear();
}

void Engine::set_labels_device(const int32_t* labels, int K) {
  if (model_ < 0) throw_invalid("call lcb_model_init first");
  if (K < 1 || labels == nullptr) throw_invalid("set_labels: bad arguments");
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_live_ = false;
  host_stale_ = false;
  main_.K = 0;
  list_valid_ = false;
  ensure_q(main_, K);
  if (prec_ == kF32) check(dev::labels_to_q<float>(stream_, labels, (float*)main_.q, main_.ldq, main_.N, K), "labels_to_q");
  else check(dev::labels_to_q<double>(stream_, labels, (double*)main_.q, main_.ldq, main_.N, K), "labels_to_q");
  sync();
  main_.K = K;
  clusters_.clear();
  hints_.clear();
}

void Engine::vbem_public(int maxit, double* F, int* iters) {
  if (model_ < 0 || main_.K < 1) throw_invalid("model and responsibilities must be set first");
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_drop();
  trace_F_.clear();
  trace_K_.clear();
  v_ntot_ = N_total_;
  last_F_ = vbem(main_, weights_, clusters_, hints_, maxit, true, iters);
  if (F) *F = last_F_;
}

void Engine::vbem_step(double* F) {
  if (model_ < 0 || main_.K < 1) throw_invalid("model and responsibilities must be set first");
  check(cudaSetDevice(device_), "cudaSetDevice");
  const int J = main_.J, K = main_.K;
  while ((int)weights_.size() < J) weights_.emplace_back(wkind_, -1.0);
  while ((int)clusters_.size() > K) clusters_.pop_back();
  while ((int)clusters_.size() < K) clusters_.emplace_back(ckind_, prior_, main_.D);
  hints_.resize(K);
  v_ntot_ = N_total_;
  const bool on_device = use_dev_mstep_ && (main_.N > 0 || world_ > 1);
  if (on_device && !(dev_live_ && dev_view_ == &main_ && dev_K_ == K)) {
    dev_drop();
    dev_begin(main_, weights_, clusters_, hints_);
  }
  const long l0 = launches_;
  const int c0 = collectives_, s0 = syncs_;
  cudaEvent_t a, b;
  check(cudaEventCreate(&a), "event");
  check(cudaEventCreate(&b), "event");
  check(cudaEventRecord(a, stream_), "event");
  double f;
  if (on_device) {
    dev_iteration(main_, weights_, &f);
    dev_live_ = true;
    host_stale_ = true;
  } else {
    iteration(main_, weights_, clusters_, hints_, &f);
  }
  check(cudaEventRecord(b, stream_), "event");
  step_collectives_ = collectives_ - c0;
  step_syncs_ = syncs_ - s0;
  sync();
  float ms = 0;
  cudaEventElapsedTime(&ms, ev_[0], ev_[1]);
  t_s_ = ms;
  cudaEventElapsedTime(&ms, ev_[2], ev_[3]);
  t_e_ = ms;
  cudaEventElapsedTime(&ms, a, b);
  t_all_ = ms;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  step_launches_ = launches_ - l0;
  last_F_ = f;
  if (F) *F = f;
}

void Engine::get_estep_detail(double out[8]) const {
  for (int i = 0; i < 8; ++i) out[i] = estep_detail_[i];
}

void Engine::get_step_counts(double out[4]) {
  out[0] = (double)step_launches_;
  out[1] = (double)step_collectives_;
  out[2] = (double)step_syncs_;
  out[3] = use_dev_mstep_ ? 1.0 : 0.0;
}

void Engine::get_step_timing(double out[4]) {
  out[0] = t_s_;
  out[1] = t_e_;
  out[2] = t_all_;
  out[3] = (double)step_launches_;
}

// ---------------------------------------------------------------- results ---
// qZ of group j in the caller's layout as doubles (cluster.cpp:661,692,723: every entry point returns qZ).  The
// emit pass is bound by the PCIe link, so it is pipelined:
//  * row-major into page-locked memory: fp32 -> fp64 on the device into two alternating buffers, each DMA'd straight
//    into the caller's matrix by a copy stream while the next chunk converts;
//  * otherwise: the engine's element type crosses the link (4 bytes per value in LCB_F32) into two page-locked
//    staging buffers and the host's threads widen / transpose chunk i into `out` while chunk i + 1 is in flight.
void Engine::get_qz(int j, double* out, int64_t ld, int layout) {
  if (j < 0 || j >= main_.J || out == nullptr) throw_invalid("get_qz: bad group");
  check(cudaSetDevice(device_), "cudaSetDevice");
  const int K = main_.K;
  int64_t off = 0;
  for (int g = 0; g < j; ++g) off += Nj_[g];
  const int64_t Nj = Nj_[j];
  if (ld < (layout == 0 ? K : Nj)) throw_invalid("get_qz: leading dimension too small");
  if (Nj <= 0 || K <= 0) return;
  const size_t es = prec_ == kF32 ? 4 : 8;
  bool pinned_dst = false;
  if (layout == 0) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, out) == cudaSuccess) pinned_dst = at.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  struct Res {
    cudaStream_t cs = nullptr;
    cudaEvent_t a[2] = {nullptr, nullptr}, b[2] = {nullptr, nullptr};
    ~Res() {
      for (int i = 0; i < 2; ++i) {
        if (a[i]) cudaEventDestroy(a[i]);
        if (b[i]) cudaEventDestroy(b[i]);
      }
      if (cs) cudaStreamDestroy(cs);
    }
  } r;
  for (int i = 0; i < 2; ++i) {
    check(cudaEventCreateWithFlags(&r.a[i], cudaEventDisableTiming), "event");
    check(cudaEventCreateWithFlags(&r.b[i], cudaEventDisableTiming), "event");
  }
  if (pinned_dst) {
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(64u << 20) / (8 * (int64_t)K));
    reserve(d_tmp_, 2 * sizeof(double) * (size_t)chunk * K);
    check(cudaStreamCreateWithFlags(&r.cs, cudaStreamNonBlocking), "stream");
    int64_t i = 0;
    for (int64_t r0 = 0; r0 < Nj; r0 += chunk, ++i) {
      const int slot = (int)(i & 1);
      const int64_t rows = std::min(chunk, Nj - r0);
      double* dbuf = (double*)d_tmp_.p + (size_t)slot * chunk * K;
      if (i >= 2) check(cudaStreamWaitEvent(stream_, r.b[slot], 0), "wait");
      if (prec_ == kF32)
        check(dev::q_to_double<float>(stream_, (const float*)main_.q + (off + r0) * main_.ldq, main_.ldq, rows, K, dbuf, K, 0), "q_to_double");
      else
        check(dev::q_to_double<double>(stream_, (const double*)main_.q + (off + r0) * main_.ldq, main_.ldq, rows, K, dbuf, K, 0), "q_to_double");
      check(cudaEventRecord(r.a[slot], stream_), "event");
      check(cudaStreamWaitEvent(r.cs, r.a[slot], 0), "wait");
      check(cudaMemcpy2DAsync(out + r0 * ld, sizeof(double) * (size_t)ld, dbuf, sizeof(double) * (size_t)K,
                              sizeof(double) * (size_t)K, (size_t)rows, cudaMemcpyDeviceToHost, r.cs), "D2H q");
      check(cudaEventRecord(r.b[slot], r.cs), "event");
    }
    check(cudaStreamSynchronize(r.cs), "sync");
    sync();
    return;
  }
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(32u << 20) / ((int64_t)es * K));
  unsigned char* h = (unsigned char*)pinned(2 * es * (size_t)chunk * K);
  const int64_t nchunks = (Nj + chunk - 1) / chunk;
  auto widen = [&](int64_t c) {
    const int64_t r0 = c * chunk, rows = std::min(chunk, Nj - r0);
    const unsigned char* hs = h + (size_t)(c & 1) * es * (size_t)chunk * K;
    const int64_t nblk = (rows + 255) / 256;
#pragma omp parallel for schedule(static) num_threads(host_threads_) if (rows * K > 65536)
    for (int64_t b = 0; b < nblk; ++b) {
      const int64_t n0 = b * 256, n1 = std::min(rows, n0 + 256);
      if (layout == 0) {
        for (int64_t n = n0; n < n1; ++n) {
          double* dst = out + (r0 + n) * ld;
          if (prec_ == kF32) {
            const float* src = (const float*)hs + n * K;
            for (int k = 0; k < K; ++k) dst[k] = (double)src[k];
          } else {
            std::memcpy(dst, (const double*)hs + n * K, sizeof(double) * (size_t)K);
          }
        }
      } else {
        for (int k = 0; k < K; ++k) {
          double* dst = out + (int64_t)k * ld + r0;
          if (prec_ == kF32) {
            const float* src = (const float*)hs + k;
            for (int64_t n = n0; n < n1; ++n) dst[n] = (double)src[n * K];
          } else {
            const double* src = (const double*)hs + k;
            for (int64_t n = n0; n < n1; ++n) dst[n] = src[n * K];
          }
        }
      }
    }
  };
  for (int64_t c = 0; c < nchunks; ++c) {
    const int slot = (int)(c & 1);
    const int64_t r0 = c * chunk, rows = std::min(chunk, Nj - r0);
    // slot `slot` was widened two rounds ago (widen(c - 2) ran before this copy is issued)
    check(cudaMemcpy2DAsync(h + (size_t)slot * es * (size_t)chunk * K, es * (size_t)K,
                            (const unsigned char*)main_.q + es * (size_t)((off + r0) * main_.ldq), es * (size_t)main_.ldq,
                            es * (size_t)K, (size_t)rows, cudaMemcpyDeviceToHost, stream_), "D2H q");
    check(cudaEventRecord(r.a[slot], stream_), "event");
    if (c > 0) {
      check(cudaEventSynchronize(r.a[slot ^ 1]), "event sync");
      widen(c - 1);
    }
  }
  check(cudaEventSynchronize(r.a[(nchunks - 1) & 1]), "event sync");
  widen(nchunks - 1);
}

void Engine::get_group_weights(int j, double* Nk, double* Elogw, double* fen) {
  ensure_host_model();
  if (j < 0 || j >= (int)weights_.size()) throw_invalid("get_group_weights: bad group");
  const WeightPost& w = weights_[j];
  if (Nk) std::copy(w.getNk().begin(), w.getNk().end(), Nk);
  if (Elogw) std::copy(w.Elogweight().begin(), w.Elogweight().end(), Elogw);
  if (fen) *fen = w.fenergy();
}

void Engine::get_cluster(int k, double* N_s, double* x_s, double* xx_s, double* N, double* mean, double* cov,
                         double* fen) {
  ensure_host_model();
  if (k < 0 || k >= (int)clusters_.size()) throw_invalid("get_cluster: bad cluster");
  const ClusterPost& c = clusters_[k];
  if (N_s) *N_s = c.N_s();
  if (x_s) std::copy(c.x_s().begin(), c.x_s().end(), x_s);
  if (xx_s) std::copy(c.xx_s().begin(), c.xx_s().end(), xx_s);
  if (N) *N = c.getN();
  if (mean) std::copy(c.mean().begin(), c.mean().end(), mean);
  if (cov) {
    std::vector<double> cv = c.cov();
    std::copy(cv.begin(), cv.end(), cov);
  }
  if (fen) *fen = c.fenergy();
}

// -------------------------------------------------------- operator surface ---
namespace {
std::vector<double> host_colmean(const double* X, int64_t N, int D, int64_t ld, int layout) {
  std::vector<double> m(D, 0.0);
  if (N <= 0) return m;
  for (int d = 0; d < D; ++d) {
    double s = 0;
    if (layout == 0) for (int64_t n = 0; n < N; ++n) s += X[n * ld + d];
    else for (int64_t n = 0; n < N; ++n) s += X[(int64_t)d * ld + n];
    m[d] = s / (double)N;
  }
  return m;
}
}  // namespace

void Engine::op_addobs(ClusterPost& c, const double* qk, const double* X, int64_t N, int64_t ld, int layout) {
  const int D = c.dim();
  if (N < 0 || (N > 0 && (X == nullptr || qk == nullptr))) throw_invalid("addobs: bad arguments");
  if (N == 0) return;
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_drop();
  // temporary view with its own centring; engine state is saved and restored
  View saved = main_;
  std::vector<double> saved_centre = centre_;
  const int saved_ck = ckind_, saved_wk = wkind_;
  const bool saved_sparse = sparse_;
  std::vector<uint8_t> saved_act = act_;
  const int saved_world = world_;
  const double saved_absmax = xabs_max_;
  main_ = View();
  View tmp;
  try {
    centre_ = host_colmean(X, N, D, ld, layout);
    tmp.N = N; tmp.D = D; tmp.J = 1; tmp.ldx = round_up(D, 4); tmp.owns_x = true;
    const size_t es = prec_ == kF32 ? 4 : 8;
    dev_alloc(&tmp.X, (size_t)N * tmp.ldx * es);
    const double* Xs[1] = {X};
    const int64_t Ns[1] = {N};
    const int64_t lds[1] = {ld};
    if (prec_ == kF32 && layout == 0) upload_rows_f32(tmp, Xs, Ns, lds, 1, centre_);
    else if (prec_ == kF32) upload_rows<float>(tmp, Xs, Ns, lds, 1, layout, centre_);
    else upload_rows<double>(tmp, Xs, Ns, lds, 1, layout, centre_);
    ensure_q(tmp, 1);
    tmp.K = 1;
    const int64_t chunk = 1 << 20;
    reserve(d_tmp_, sizeof(double) * std::min(chunk, N));
    double* h = (double*)pinned(sizeof(double) * std::min(chunk, N));
    for (int64_t r = 0; r < N; r += chunk) {
      const int64_t rows = std::min(chunk, N - r);
      std::memcpy(h, qk + r, sizeof(double) * rows);
      check(cudaMemcpyAsync(d_tmp_.p, h, sizeof(double) * rows, cudaMemcpyHostToDevice, stream_), "H2D qk");
      if (prec_ == kF32) check(dev::q_from_double<float>(stream_, (const double*)d_tmp_.p, rows, 1, (float*)tmp.q + r * tmp.ldq, tmp.ldq), "q");
      else check(dev::q_from_double<double>(stream_, (const double*)d_tmp_.p, rows, 1, (double*)tmp.q + r * tmp.ldq, tmp.ldq), "q");
      sync();
    }
    ckind_ = c.kind();
    sparse_ = false;
    act_.clear();
    std::vector<WeightPost> w(1, WeightPost(kDirichlet, -1.0));
    std::vector<ClusterPost> cl(1, ClusterPost(c.kind(), c.getprior(), D));
    std::vector<std::vector<double>> cen(1, c.getN() > 0 ? c.mean() : centre_);
    world_ = 1;  // operator calls are local
    measure_absmax(tmp);
    sphase(tmp, w, cl, cen);
    c.add_stats(cl[0].N_s(), cl[0].x_s().data(), cl[0].xx_s().data());
  } catch (...) {
    free_view(tmp);
    main_ = saved; centre_ = saved_centre; ckind_ = saved_ck; wkind_ = saved_wk; sparse_ = saved_sparse; act_ = saved_act;
    world_ = saved_world; xabs_max_ = saved_absmax;
    throw;
  }
  free_view(tmp);
  main_ = saved; centre_ = saved_centre; ckind_ = saved_ck; wkind_ = saved_wk; sparse_ = saved_sparse; act_ = saved_act;
  world_ = saved_world; xabs_max_ = saved_absmax;
}

void Engine::op_eloglike(const ClusterPost& c, const double* X, int64_t N, int64_t ld, int layout, double* out) {
  const int D = c.dim();
  if (N < 0 || (N > 0 && (X == nullptr || out == nullptr))) throw_invalid("Eloglike: bad arguments");
  if (N == 0) return;
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_drop();
  View saved = main_;
  std::vector<double> saved_centre = centre_;
  const int saved_ck = ckind_;
  const bool saved_sparse = sparse_;
  main_ = View();
  View tmp;
  auto restore = [&]() { free_view(tmp); main_ = saved; centre_ = saved_centre; ckind_ = saved_ck; sparse_ = saved_sparse; };
  try {
    centre_ = c.mean();  // centre on the cluster: the device sees x - m directly
    tmp.N = N; tmp.D = D; tmp.J = 1; tmp.ldx = round_up(D, 4); tmp.owns_x = true;
    const size_t es = prec_ == kF32 ? 4 : 8;
    dev_alloc(&tmp.X, (size_t)N * tmp.ldx * es);
    const double* Xs[1] = {X};
    const int64_t Ns[1] = {N};
    const int64_t lds[1] = {ld};
    if (prec_ == kF32 && layout == 0) upload_rows_f32(tmp, Xs, Ns, lds, 1, centre_);
    else if (prec_ == kF32) upload_rows<float>(tmp, Xs, Ns, lds, 1, layout, centre_);
    else upload_rows<double>(tmp, Xs, Ns, lds, 1, layout, centre_);
    ensure_q(tmp, 1);
    tmp.K = 1;
    ckind_ = c.kind();
    sparse_ = false;
    std::vector<WeightPost> w(1, WeightPost(kDirichlet, -1.0));
    std::vector<ClusterPost> cl(1, c);
    v_ntot_ = N;
    ephase(tmp, w, cl, dev::kERawLogit, nullptr);
    const double cc = c.cconst();
    const int64_t chunk = 1 << 20;
    reserve(d_tmp_, sizeof(double) * std::min(chunk, N));
    double* h = (double*)pinned(sizeof(double) * std::min(chunk, N));
    for (int64_t r = 0; r < N; r += chunk) {
      const int64_t rows = std::min(chunk, N - r);
      if (prec_ == kF32) check(dev::q_to_double<float>(stream_, (const float*)tmp.q + r * tmp.ldq, tmp.ldq, rows, 1, (double*)d_tmp_.p, 1, 0), "q");
      else check(dev::q_to_double<double>(stream_, (const double*)tmp.q + r * tmp.ldq, tmp.ldq, rows, 1, (double*)d_tmp_.p, 1, 0), "q");
      check(cudaMemcpyAsync(h, d_tmp_.p, sizeof(double) * rows, cudaMemcpyDeviceToHost, stream_), "D2H");
      sync();
      for (int64_t n = 0; n < rows; ++n) out[r + n] = cc + h[n];
    }
  } catch (...) {
    restore();
    throw;
  }
  restore();
}

void Engine::op_splitobs(const ClusterPost& c, const double* X, int64_t N, int64_t ld, int layout, uint8_t* out) {
  const int D = c.dim();
  if (N < 0 || (N > 0 && (X == nullptr || out == nullptr))) throw_invalid("splitobs: bad arguments");
  if (N == 0) return;
  check(cudaSetDevice(device_), "cudaSetDevice");
  dev_drop();
  View tmp;
  void* d_flags = nullptr;
  try {
    std::vector<double> ctr = c.mean();
    tmp.N = N; tmp.D = D; tmp.J = 1; tmp.ldx = round_up(D, 4); tmp.owns_x = true;
    const size_t es = prec_ == kF32 ? 4 : 8;
    dev_alloc(&tmp.X, (size_t)N * tmp.ldx * es);
    const double* Xs[1] = {X};
    const int64_t Ns[1] = {N};
    const int64_t lds[1] = {ld};
    if (prec_ == kF32 && layout == 0) upload_rows_f32(tmp, Xs, Ns, lds, 1, ctr);
    else if (prec_ == kF32) upload_rows<float>(tmp, Xs, Ns, lds, 1, layout, ctr);
    else upload_rows<double>(tmp, Xs, Ns, lds, 1, layout, ctr);
    std::vector<double> dir;
    c.split_direction(dir);
    unsigned char* h = (unsigned char*)pinned(2 * (size_t)D * es);
    for (int d = 0; d < D; ++d) {
      if (prec_ == kF32) { ((float*)h)[d] = 0.f; ((float*)h)[D + d] = (float)dir[d]; }
      else { ((double*)h)[d] = 0.0; ((double*)h)[D + d] = dir[d]; }
    }
    reserve(d_cen_, 2 * (size_t)D * es);
    check(cudaMemcpyAsync(d_cen_.p, h, 2 * (size_t)D * es, cudaMemcpyHostToDevice, stream_), "H2D dir");
    dev_alloc(&d_flags, (size_t)N);
    if (prec_ == kF32)
      check(dev::side_flags<float>(stream_, (const float*)tmp.X, N, D, tmp.ldx, (const float*)d_cen_.p, (const float*)d_cen_.p + D, (uint8_t*)d_flags), "side_flags");
    else
      check(dev::side_flags<double>(stream_, (const double*)tmp.X, N, D, tmp.ldx, (const double*)d_cen_.p, (const double*)d_cen_.p + D, (uint8_t*)d_flags), "side_flags");
    check(cudaMemcpyAsync(out, d_flags, (size_t)N, cudaMemcpyDeviceToHost, stream_), "D2H flags");
    sync();
  } catch (...) {
    if (d_flags) cudaFree(d_flags);
    free_view(tmp);
    throw;
  }
  cudaFree(d_flags);
  free_view(tmp);
}

}  // namespace lcb
