/*
 * libcluster_b200.h -- C ABI of the B200-native variational E-step engine.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Each entry point names the reference interface it replaces
 * (paths relative to dsteinberg/libcluster @ c877625).  The reference-side
 * binding a maintainer would add on top of this is shown in INTEGRATION.md and
 * shipped as include/libcluster.h + include/distributions.h (Eigen shim).
 *
 * All O(N) work runs in hand-written sm_100a CUDA kernels; there is no CPU
 * fallback: every compute entry point returns LCB_ECUDA when no device or no
 * kernel image is available.
 */
#ifndef LIBCLUSTER_B200_H
#define LIBCLUSTER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes <-> the reference's exceptions (include/libcluster.h:171-175)
 *  LCB_EINVAL   std::invalid_argument  (cluster.cpp:576-577, distributions.cpp:107,234,282,415,460)
 *  LCB_ERUNTIME std::runtime_error     ("Free energy increase!" cluster.cpp:229-230)
 *  LCB_EDOMAIN  std::domain_error      (non-PD iW out of probutils::logdet, probutils.cpp:198-199)
 *  LCB_ECUDA    no reference analogue: device / driver / kernel-image failure
 */
enum { LCB_OK = 0, LCB_EINVAL = 1, LCB_ERUNTIME = 2, LCB_EDOMAIN = 3, LCB_ECUDA = 4, LCB_ENOMEM = 5 };

/* learnXXX entry points kept (include/libcluster.h:177,218,262,356,409,462) */
enum { LCB_VDP = 0, LCB_BGMM = 1, LCB_DGMM = 2, LCB_GMC = 3, LCB_SGMC = 4, LCB_DGMC = 5 };
/* weight / cluster operator classes kept (include/distributions.h:103,147,163,279,343) */
enum { LCB_W_DIRICHLET = 0, LCB_W_STICKBREAK = 1, LCB_W_GDIRICHLET = 2 };
enum { LCB_C_GAUSSWISH = 0, LCB_C_NORMGAMMA = 1 };
/* Eigen storage orders met at the boundary (CMakeLists.txt:59-61) */
enum { LCB_ROW_MAJOR = 0, LCB_COL_MAJOR = 1 };
/* device arithmetic: LCB_F32 is the measured path (fp32 data, fp32 per-tile
 * math, fp64 cross-tile accumulation); LCB_F64 is an exact fp64 variant of the
 * same kernels for decision-level parity and for validating F32 at full size. */
enum { LCB_F32 = 0, LCB_F64 = 1 };

/* namespace constants, include/libcluster.h:122-127 */
#define LCB_PRIORVAL 1.0
#define LCB_SPLITITER 15

typedef struct lcb_engine lcb_engine;
typedef struct lcb_weights lcb_weights; /* one WeightDist object  */
typedef struct lcb_cluster lcb_cluster; /* one ClusterDist object */

/* Thread-local message of the last non-OK status (the exception's what()). */
const char *lcb_last_error(void);
const char *lcb_version(void);
/* Number of CUDA devices visible; 0 on a host without a GPU (never fails). */
int lcb_device_count(void);

/* ------------------------------------------------------------------ engine */
int lcb_create(lcb_engine **out, int device, int precision);
void lcb_destroy(lcb_engine *e);

/* Observations: J groups, group j is an N_j x D matrix of doubles with leading
 * dimension ld[j] in the given storage order (replaces `const MatrixXd& X` /
 * `const vMatrixXd& X`, include/libcluster.h:178,357).  Copied to the device
 * (centred on the global column mean, converted to the engine precision);
 * the caller's buffers are not retained. */
int lcb_set_data(lcb_engine *e, int J, const double *const *X, const int64_t *Nj, int D,
                 const int64_t *ld, int layout);
/* Same, from a row-major fp32 matrix already resident on this engine's device
 * (synthetic benchmarks): rows [0,N) with leading dimension ld; gid (device,
 * int32, may be NULL when J == 1) gives the group of every row, group ids
 * non-decreasing.  N_total/row_offset describe this rank's shard of a larger
 * matrix (N_total == N, row_offset == 0 on one GPU). */
int lcb_set_data_device_f32(lcb_engine *e, const float *X_dev, int64_t N, int D, int64_t ld,
                            const int32_t *gid_dev, int J);

/* The model-selection fit: cluster<W,C>() (src/cluster.cpp:564-629) behind
 * learnVDP/BGMM/DGMM/GMC/SGMC/DGMC (:636-831).  weight_prior < 0 means a
 * default-constructed weight object (StickBreak()/Dirichlet()), 0 is LCB_EINVAL; nthreads has
 * no meaning on the GPU but nthreads < 1 is still LCB_EINVAL (:576-577). */
int lcb_learn(lcb_engine *e, int model, double clusterprior, double weight_prior, int maxclusters,
              int sparse, int verbose, unsigned nthreads, double *F, int *K);

/* ---- the primitive under learn: vbem<W,C>() (src/cluster.cpp:177-239) ---- */
/* Start a model of the given kind on the resident data. */
int lcb_model_init(lcb_engine *e, int model, double clusterprior, double weight_prior, int sparse);
/* Initial responsibilities: host row-major [N x K] doubles over all rows... */
int lcb_set_qz(lcb_engine *e, const double *q0, int K);
/* ...or hard labels in [0,K) resident on the device (one int32 per row). */
int lcb_set_labels_device(lcb_engine *e, const int32_t *labels_dev, int K);
/* Run vbem() from the current responsibilities (maxit as in cluster.cpp:183). */
int lcb_vbem(lcb_engine *e, int maxit, double *F, int *iters);
/* Exactly one loop body (cluster.cpp:203-226): suff. stats of the current
 * responsibilities -> posteriors -> new responsibilities -> F.  No convergence
 * or monotonicity test.  This is the timed "step" of bench.py. */
int lcb_vbem_step(lcb_engine *e, double *F);

/* ---- results ------------------------------------------------------------ */
int lcb_num_clusters(const lcb_engine *e);
int lcb_num_groups(const lcb_engine *e);
int64_t lcb_num_rows(const lcb_engine *e, int j); /* j < 0: all groups */
/* qZ of group j (include/libcluster.h:179) as doubles, N_j x K, given order. */
int lcb_get_qz(lcb_engine *e, int j, double *out, int64_t ld, int layout);
/* WeightDist results of group j: getNk(), Elogweight(), fenergy(). */
int lcb_get_group_weights(lcb_engine *e, int j, double *Nk, double *Elogweight, double *fenergy);
/* ClusterDist k: raw sufficient statistics N_s, x_s[D], xx_s[D*D | D]
 * (row-major) from which the shim rebuilds GaussWish/NormGamma, plus the
 * posterior getN(), getmean()[D], getcov()[D*D | D] and fenergy().  Any
 * output pointer may be NULL. */
int lcb_get_cluster(lcb_engine *e, int k, double *N_s, double *x_s, double *xx_s, double *N,
                    double *mean, double *cov, double *fenergy);
/* F after every vbem iteration of the last learn/vbem call (and K there). */
int lcb_trace_len(const lcb_engine *e);
int lcb_get_trace(const lcb_engine *e, double *F, int *K);
/* Device time of the kernels of the last lcb_vbem_step, in milliseconds,
 * from CUDA events on the engine's stream:
 *   out[0] suff-stat pass, out[1] E-step pass, out[2] whole step (device),
 *   out[3] launches issued in the step. */
int lcb_get_step_timing(lcb_engine *e, double out[4]);
/* Break-down of the last E pass when it ran on the tensor-core tier (D = 128, LCB_F32):
 *   out[0..3] device ms of level 1 (one-product distances + candidate marking), of the candidate lists,
 *   of level 2 (exact logits of the candidates) and of the row soft-max;
 *   out[4] candidate (row, cluster) pairs, out[5] path (0 dense kernel, 1 two-level, 2 two-level abandoned for
 *   the dense kernel because too many pairs were candidates), out[6] 128-pair work items of level 2. */
int lcb_get_estep_detail(lcb_engine *e, double out[8]);
/* The CUDA stream all engine work is issued on (cudaStream_t as void*). */
void *lcb_stream(lcb_engine *e);
/* Host-only self-test (no device needed): the operand packing has a run-time-dispatched F16C path; returns the number
 * of bytes in which it differs from the portable path on a test matrix (0 = identical, also where F16C is absent). */
int lcb_selftest_host_packing(void);
/* Counters of the last lcb_vbem_step: out[0] kernel launches, out[1] collectives (NCCL or host callback),
 * out[2] host synchronisations inside the step, out[3] 1 when the M step ran on the device (default; the
 * environment variable LCB_HOST_MSTEP=1 keeps the posterior updates of cluster.cpp:211,217 on the host). */
int lcb_get_step_counts(lcb_engine *e, double out[4]);
/* Host-only self-test: the device M step orders the sticks of StickBreak / GDirichlet (distributions.cpp:146) with a
 * restatement of libstdc++'s std::sort so that exact ties fall as in the reference; order[] receives the indices of
 * counts[] sorted greater-first by that code; the return value is the number of positions in which it differs from
 * std::sort on the same input (0 = identical order). */
int lcb_selftest_stick_order(const double *counts, int n, int *order);

/* ---- multi-GPU: rows sharded over ranks, one all-reduce of the packed
 * sufficient statistics per VB iteration.  No reference
 * analogue (the reference is single-process OpenMP, cluster.cpp:207-223). */
int lcb_nccl_unique_id(char out[128]);
int lcb_comm_init_nccl(lcb_engine *e, const char id[128], int rank, int world);
/* Host-side reduction hook (gloo tests, custom transports): called with the
 * packed fp64 buffer; must sum it element-wise across ranks in place. */
typedef int (*lcb_allreduce_fn)(double *buf, int64_t count, void *ctx);
int lcb_comm_init_host(lcb_engine *e, lcb_allreduce_fn fn, void *ctx, int rank, int world);

/* ---------------------------------------------- operator surface (L1) ----
 * WeightDist: update / Elogweight / getNk / fenergy (distributions.h:60-97).
 * prior < 0 selects the default constructor, prior == 0 is LCB_EINVAL (distributions.cpp:107,234).
 * Host-side, K-length math. */
int lcb_weights_create(lcb_weights **out, int kind, double prior);
void lcb_weights_destroy(lcb_weights *w);
int lcb_weights_update(lcb_weights *w, const double *Nk, int K);
int lcb_weights_size(const lcb_weights *w);
int lcb_weights_elogweight(const lcb_weights *w, double *out);
int lcb_weights_getnk(const lcb_weights *w, double *out);
double lcb_weights_fenergy(const lcb_weights *w);

/* ClusterDist: addobs / update / clearobs / Eloglike / fenergy / splitobs /
 * getN / getprior (distributions.h:200-273), GaussWish (:279-337) and
 * NormGamma (:343-400).  addobs, Eloglike and splitobs stream X through the
 * GPU kernels of engine `e` (host buffers in, host buffers out). */
int lcb_cluster_create(lcb_cluster **out, int kind, double clustwidth, int D);
void lcb_cluster_destroy(lcb_cluster *c);
int lcb_cluster_addobs(lcb_engine *e, lcb_cluster *c, const double *qZk, const double *X, int64_t N,
                       int64_t ld, int layout);
int lcb_cluster_update(lcb_cluster *c);
int lcb_cluster_clearobs(lcb_cluster *c);
int lcb_cluster_eloglike(lcb_engine *e, const lcb_cluster *c, const double *X, int64_t N, int64_t ld,
                         int layout, double *out);
int lcb_cluster_splitobs(lcb_engine *e, const lcb_cluster *c, const double *X, int64_t N, int64_t ld,
                         int layout, uint8_t *out);
double lcb_cluster_fenergy(const lcb_cluster *c);
double lcb_cluster_getn(const lcb_cluster *c);
double lcb_cluster_getprior(const lcb_cluster *c);
int lcb_cluster_dim(const lcb_cluster *c);
int lcb_cluster_getmean(const lcb_cluster *c, double *out);
int lcb_cluster_getcov(const lcb_cluster *c, double *out);
/* raw sufficient statistics in/out (row-major), for the Eigen shim */
int lcb_cluster_get_stats(const lcb_cluster *c, double *N_s, double *x_s, double *xx_s);
int lcb_cluster_set_stats(lcb_cluster *c, double N_s, const double *x_s, const double *xx_s);

/* ---- host-only pieces of the VB iteration, exposed for multi-rank tests ----
 * Packed statistics layout of one iteration (what is all-reduced):
 *   [ Njk (J*K) | per cluster k: N_s, x_s[D], xx_s[D*D | D] ]  (doubles)
 * lcb_packed_len gives its length; lcb_host_mstep runs the weight and cluster
 * posterior updates (cluster.cpp:211,217) on such a buffer and returns the
 * parameter part of the free energy sum_j Fw_j + sum_k Fc_k (cluster.cpp:155-162). */
int64_t lcb_packed_len(int model, int J, int K, int D);
int lcb_host_mstep(int model, double clusterprior, double weight_prior, int J, int K, int D,
                   const double *packed, double *Fparams, double *Elogweight /* J*K */,
                   double *means /* K*D */, double *covs /* K*(D*D | D) */);
/* Contiguous row shard [begin,end) of rank r of `world` over N rows. */
void lcb_shard_rows(int64_t N, int rank, int world, int64_t *begin, int64_t *end);

#ifdef __cplusplus
}
#endif
#endif /* LIBCLUSTER_B200_H */
