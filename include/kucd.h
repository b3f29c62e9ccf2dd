/* kucd.h - C ABI of libkucd.so, the B200 (sm_100a) contrastive-divergence engine that sits behind
 * keras_unsupervised's ku.ebm.RBM / ku.ebm.DBN.
 *
 * The reference has no FFI of its own: its boundary is the Python class surface of
 * /root/reference/ku/ebm/rbm.py and dbn.py, and everything below that line executes inside
 * TensorFlow through K.function objects.  Each entry point here replaces one of those K.function
 * objects (file:line given per function); the Python shim in keras_unsupervised_b200/ebm binds them
 * with ctypes and keeps the reference's method names and argument meaning.
 *
 * Conventions
 *   - plain C: pointers, sizes and PODs only; no C++ / torch types cross this boundary.
 *   - every function returns KUCD_OK (0) or a negative kucd_status; the message of the last failure on
 *     the calling thread is returned by kucd_last_error().  Nothing throws, nothing aborts.
 *   - tensors are described by kucd_tensor, a flattened DLTensor (a DLManagedTensor's dl_tensor maps
 *     field by field).  Data may live in host memory (pageable or pinned) or on the context's GPU; the
 *     engine copies as needed on its own stream.  The caller owns every buffer it passes and must keep
 *     it alive until the call returns (all entry points that touch caller memory are synchronous with
 *     respect to that memory; device work on engine-owned state may still be in flight - kucd_sync).
 *   - there is no CPU fallback.  A device that is not compute capability 10.x yields KUCD_ERR_NOT_SM100.
 *   - one context = one GPU = one host thread at a time.  Data-parallel training is one process (and
 *     one context) per GPU, joined by kucd_ctx_comm_init.
 */
#ifndef KUCD_H_
#define KUCD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KUCD_ABI_VERSION 1

typedef enum kucd_status {
  KUCD_OK = 0,
  KUCD_ERR_INVALID_ARG = -1,
  KUCD_ERR_SHAPE_MISMATCH = -2,
  KUCD_ERR_UNSUPPORTED_DTYPE = -3,
  KUCD_ERR_CUDA = -4,
  KUCD_ERR_NCCL = -5,
  KUCD_ERR_NOT_SM100 = -6,
  KUCD_ERR_NOT_BUILT = -7
} kucd_status;

/* DLPack device types / dtype codes (same numeric values as dlpack.h) */
#define KUCD_DEV_CPU 1
#define KUCD_DEV_CUDA 2
#define KUCD_DEV_CUDA_HOST 3
#define KUCD_DT_INT 0
#define KUCD_DT_UINT 1
#define KUCD_DT_FLOAT 2
#define KUCD_DT_BFLOAT 4

/* A 2-D (or 1-D: shape[1] = 1) strided tensor.  strides are in elements; the innermost stride must
 * be 1.  Accepted element types: float32 (the reference's K.floatx()), bfloat16, uint8 and - for 0/1
 * matrices (binarised data, sampled states) - packed bits: dtype_code = KUCD_DT_UINT, bits = 1, column j
 * of a row is bit (j % 8) of byte (j / 8) of that row (numpy.packbits(x, axis=1, bitorder="little")),
 * shape[1] is the number of columns and strides[0] the row pitch IN BITS (a multiple of 8).  Packed
 * inputs are accepted wherever a visible / hidden matrix is read; packed outputs wherever the result is
 * a 0/1 state (KUCD_ERR_UNSUPPORTED_DTYPE for probabilities, free energies, Gaussian visibles). */
typedef struct kucd_tensor {
  void* data;
  int32_t device_type;
  int32_t device_id;
  int32_t dtype_code;
  int32_t bits;
  int64_t shape[2];
  int64_t strides[2];
} kucd_tensor;

/* Describe a DLPack tensor.  `dl_managed_tensor` is the `DLManagedTensor*` a producer hands out inside its "dltensor"
 * capsule (tf.experimental.dlpack.to_dlpack(x), x.__dlpack__() of torch / jax / cupy / numpy >= 1.22): its dl_tensor maps
 * onto kucd_tensor field by field (data + byte_offset, device, dtype, shape, strides in elements; NULL strides = compact
 * row-major).  1-D tensors become (n, 1).  The descriptor BORROWS the producer's memory: the caller keeps the capsule
 * alive for the duration of the engine call and lets its deleter run afterwards.  Nothing is copied; no device is
 * touched (works without a GPU).  KUCD_ERR_INVALID_ARG: more than two dimensions, vector lanes, an innermost stride
 * other than 1; KUCD_ERR_UNSUPPORTED_DTYPE: not float32 / bfloat16 / uint8 / bool. */
int kucd_tensor_from_dlpack(const void* dl_managed_tensor, kucd_tensor* out);

/* visible-unit mode: constants of rbm.py:14-16 */
#define KUCD_MODE_VISIBLE_BERNOULLI 0
#define KUCD_MODE_VISIBLE_GAUSSIAN 1

/* arithmetic of the contractions */
#define KUCD_COMPUTE_BF16 0  /* W rounded to bf16 (RNE); fp32 accumulation in tensor memory */
#define KUCD_COMPUTE_F32X3 1 /* fp32 operands carried as three bf16 terms: fp32-grade products   */

/* which parameters a step may write (rbm.py:214-216 runs three single-parameter updates) */
#define KUCD_UPDATE_W 1
#define KUCD_UPDATE_C 2 /* hidden bias  */
#define KUCD_UPDATE_B 4 /* visible bias */
#define KUCD_UPDATE_ALL 7

typedef struct kucd_hparams {
  float lr;            /* hps['lr'] (rbm.py:128); multiplies batch SUMS unless normalize != 0       */
  int32_t k;           /* Gibbs steps of the negative chain; the reference is k = 1 (rbm.py:119-124) */
  int32_t persistent;  /* != 0: negative chain starts from the engine-held chains (PCD)               */
  float momentum;      /* extension, 0 = reference                                                    */
  float weight_decay;  /* extension, 0 = reference                                                    */
  int32_t normalize;   /* 0 = batch sum (reference), 1 = divide by the GLOBAL batch rows              */
  int32_t update_mask; /* KUCD_UPDATE_*                                                               */
  int32_t want_stats;  /* 1: fill kucd_step_stats (two free-energy passes + a chain); 2: recon_err only   */
} kucd_hparams;

/* Injected random draws for parity runs: arrays of float32 uniforms on the caller's side of the
 * boundary (host or device), one per sampling node of the chain, in chain order
 *   u_h[0] : (rows, H) for h_pos            (rbm.py:46, the draw inside self.transform)
 *   u_v[t] : (rows, V) for the t-th v_neg, t = 1..k (rbm.py:121 is t = 1; u_v[0] is unused)
 *   u_h[t] : (rows, H) for the t-th intermediate h (CD-k extension; t = 1..k-1)
 *   u_hc   : (rows, H) for the first h of a persistent chain (PCD only)
 * Any pointer may be NULL: that node then draws from the engine's Philox stream. */
#define KUCD_MAX_K 32
typedef struct kucd_inject {
  const kucd_tensor* u_h[KUCD_MAX_K];
  const kucd_tensor* u_v[KUCD_MAX_K];
  const kucd_tensor* u_hc;
} kucd_inject;

typedef struct kucd_step_stats {
  float score;     /* mean |F(v) - F(v_neg)| with the post-update parameters (rbm.py:227-233)        */
  float recon_err; /* mean (v - v_neg)^2 over the batch and the visible units                         */
  float fe_mean;   /* mean F(v)                                                                       */
  int32_t rows;
} kucd_step_stats;

typedef struct kucd_epoch_stats {
  int64_t steps;
  int64_t rows;
  float device_ms; /* CUDA-event time of the epoch on the engine stream */
  float last_score;
  float last_recon_err;
} kucd_epoch_stats;

typedef struct kucd_timings {
  int64_t gemm_launches;    /* tcgen05 contraction launches enqueued (directly or while capturing)  */
  int64_t chain_launches;   /* of those: whole-chain launches (all projections of a minibatch in one)  */
  int64_t chain_dw_launches;/* of those: chain launches that also carried the dW contraction            */
  int64_t aux_launches;     /* update / split / reduction / conversion launches enqueued            */
  int64_t graph_launches;   /* CUDA-graph replays (each replays a whole CD step)                     */
  int64_t graph_kernel_launches; /* kernels those replays launched                                    */
  int64_t allreduce_calls;
  int64_t fused_reduce_steps; /* steps whose dW/db/dc exchange ran over peer-mapped memory instead of NCCL  */
  int64_t h2d_bytes, d2h_bytes;
  /* per-launch CUDA-event timing of directly launched contractions, on while kucd_ctx_set_profile(1) */
  int64_t proj_timed, dw_timed; /* launches timed: projections (v.W, h.W^T), dW contractions          */
  float proj_ms, dw_ms;         /* their summed device time                                           */
  float last_gemm_ms;
  int64_t unit_steps;     /* steps that ran unit-sharded (states exchanged as bits, no dW on the wire)          */
  int64_t unit_exchanges; /* bit exchanges those steps enqueued (pack + peer stores, flag barrier, expansion)   */
  int64_t xchg_timed, upd_timed; /* with kucd_ctx_set_profile(1): bit exchanges / parameter updates timed        */
  float xchg_ms, upd_ms;         /* their summed device time (pack + barrier + expansion; update kernels + barrier) */
} kucd_timings;

typedef struct kucd_ctx kucd_ctx;
typedef struct kucd_rbm kucd_rbm;
typedef struct kucd_dataset kucd_dataset;

int kucd_abi_version(void);
const char* kucd_last_error(void);

/* ---- context ------------------------------------------------------------------------------------- */
int kucd_ctx_create(kucd_ctx** out, int device_id, uint64_t seed);
int kucd_ctx_destroy(kucd_ctx* ctx);
int kucd_sync(kucd_ctx* ctx);
int kucd_get_timings(kucd_ctx* ctx, kucd_timings* out, int reset);
int kucd_ctx_set_profile(kucd_ctx* ctx, int enable);
/* the CUstream the engine launches on (for callers that time with their own events) */
int kucd_ctx_stream(kucd_ctx* ctx, void** stream_out);

/* Data-parallel group (one process per GPU).  Rank 0 calls kucd_comm_unique_id and ships the 128
 * bytes to every rank by any means (the Python shim uses torch.distributed); then every rank calls
 * kucd_ctx_comm_init.  Afterwards every training step all-reduces dW/db/dc over NCCL. */
int kucd_comm_unique_id(void* id128);
int kucd_ctx_comm_init(kucd_ctx* ctx, const void* id128, int rank, int world);

/* ---- model: RBM.__init__/build (rbm.py:22-40) ----------------------------------------------------- */
int kucd_rbm_create(kucd_ctx* ctx, int64_t n_visible, int64_t n_hidden, int mode, int compute,
                    kucd_rbm** out);
int kucd_rbm_destroy(kucd_rbm* rbm);
/* rbm_weight (V,H), rbm_visible_bias (V), rbm_hidden_bias (H): float32 (rbm.py:30-40) */
int kucd_rbm_set_params(kucd_rbm* rbm, const kucd_tensor* W, const kucd_tensor* b, const kucd_tensor* c);
int kucd_rbm_get_params(kucd_rbm* rbm, kucd_tensor* W, kucd_tensor* b, kucd_tensor* c);
/* Philox stream of this model.  A draw is philox4x32-10(key = seed, counter = (column / 4, global row,
 * draw id lo, draw id hi)), component column % 4, mapped to the 2^-23 lattice in [0,1) and compared
 * with a strict < (K.random_uniform + K.less, rbm.py:46).  Draw ids: training step s (counted from
 * `step_count`) uses 64 s + phase with phase 0 = h_pos, 1 = first h of a persistent chain, 2t = t-th
 * v_neg, 2t+1 = t-th negative h; the n-th transform / inv_transform call uses 2^63 + n; the n-th score
 * chain uses 2^62 + 2n (h) and 2^62 + 2n + 1 (v).  Defaults: the context seed, all counters 0. */
int kucd_rbm_set_seed(kucd_rbm* rbm, uint64_t seed, uint64_t step_count);
/* what a checkpoint needs besides the parameters: the stream position and the number of stored chains */
int kucd_rbm_get_counters(kucd_rbm* rbm, uint64_t* seed, uint64_t* step_count, int64_t* n_chains);
/* ... and, for a resumed fit to continue exactly where an uninterrupted one would be: the positions of the inference
 * (2^63 + n) and score (2^62 + 2n) streams, which kucd_rbm_set_seed resets to 0 ... */
int kucd_rbm_get_draw_counters(kucd_rbm* rbm, uint64_t* infer_draws, uint64_t* score_draws);
int kucd_rbm_set_draw_counters(kucd_rbm* rbm, uint64_t infer_draws, uint64_t score_draws);
/* ... and the momentum buffers (kucd_hparams.momentum != 0; the reference has none): mW (V,H), mb (V), mc (H), float32.
 * get: zeros while no momentum step has run; *present (optional) tells which.  set: allocates them.  On several ranks
 * these are the LOCAL buffers (the fused and the unit-sharded exchange keep only the slices a rank owns up to date):
 * a checkpoint stores them per rank and restores them into the same group layout. */
int kucd_rbm_get_momentum(kucd_rbm* rbm, kucd_tensor* mW, kucd_tensor* mb, kucd_tensor* mc, int* present);
int kucd_rbm_set_momentum(kucd_rbm* rbm, const kucd_tensor* mW, const kucd_tensor* mb, const kucd_tensor* mc);

/* Fused reduction (optional, bf16 compute, 2..8 ranks on one NVLink domain).  After kucd_ctx_comm_init every rank
 * calls kucd_rbm_peer_export on its model (128 bytes of CUDA IPC handles out), the ranks exchange them, and every
 * rank calls kucd_rbm_peer_attach with all of them in rank order (world x 128 bytes).  From then on the dW
 * contraction stores each output row directly into its owner rank's memory (reduce-scatter inside the epilogue,
 * NVLink stores overlapped with the MMAs), the owner updates its rows and stores the refreshed bf16 rows into every
 * rank's operand plane, and two flag barriers per step replace the NCCL all-reduce.  The sum over ranks runs in a
 * fixed order: every rank (and every run) gets the same bits.  fp32 master rows are re-gathered (ncclBroadcast) when
 * a training call returns. */
int kucd_rbm_peer_export(kucd_rbm* rbm, void* handle128);
int kucd_rbm_peer_attach(kucd_rbm* rbm, const void* handles);
/* back to the NCCL all-reduce (every rank must do the same: the exchange is collective) */
int kucd_rbm_peer_detach(kucd_rbm* rbm);

/* ---- inference ------------------------------------------------------------------------------------- */
/* transform_func (rbm.py:45-48, 88-89) and RBM.call (rbm.py:80-83):  h = 1[u < sigmoid(v.W + c)]
 * (Gaussian mode, rbm.py:58-59: relu instead of sigmoid).  h_out (rows,H) and p_out (rows,H, nullable,
 * float32 probabilities) are caller buffers; u (rows,H, nullable) injects the uniforms. */
int kucd_rbm_transform(kucd_rbm* rbm, const kucd_tensor* v, kucd_tensor* h_out, kucd_tensor* p_out,
                       const kucd_tensor* u);
/* inv_transform_func (rbm.py:51-54, 91-92):  v' = 1[u < sigmoid(h.W^T + b)]
 * (Gaussian mode, rbm.py:64-66: v' = h.W^T + b + n, n ~ N(0,1); `u` then carries the normals). */
int kucd_rbm_inv_transform(kucd_rbm* rbm, const kucd_tensor* h, kucd_tensor* v_out, kucd_tensor* p_out,
                           const kucd_tensor* u);
/* free_energy_func (rbm.py:73-76, 97-98):  F = -(v.b + sum_j log(1 + exp((v.W + c)_j))), (rows,) f32 */
int kucd_rbm_free_energy(kucd_rbm* rbm, const kucd_tensor* v, kucd_tensor* fe_out);

/* ---- training -------------------------------------------------------------------------------------- */
/* One minibatch: the graph of rbm.py:119-134 (CD-1), generalised to CD-k / PCD, followed by the
 * updates selected by hp->update_mask.  `global_row0` is the index of this rank's first row inside the
 * global minibatch (0 on a single GPU): Philox draws are keyed by global row, so n ranks sample exactly
 * what one rank would. */
int kucd_rbm_cd_step(kucd_rbm* rbm, const kucd_tensor* v_batch, const kucd_hparams* hp,
                     const kucd_inject* inj, int64_t global_row0, kucd_step_stats* stats);
/* rbm.py:225-233: score = mean|F(v) - F(v_neg)| with a fresh chain (the reference's "run D").
 * u_h (rows,H) / u_v (rows,V) nullable. */
int kucd_rbm_score(kucd_rbm* rbm, const kucd_tensor* v_batch, const kucd_tensor* u_h,
                   const kucd_tensor* u_v, float* score_out);
/* statistics of the most recent cd_step, for parity tests: dW (V,H), db (V), dc (H) float32 - the
 * un-scaled batch sums of rbm.py:125-126,131,134 (after the all-reduce when a group is attached; under the
 * fused exchange dW is never assembled on one rank and this call returns stale values for it - db, dc are global);
 * and the sampled states of that step: h_pos (rows,H), v_neg (rows,V), h_neg (rows,H).  Any NULL. */
int kucd_rbm_last_stats(kucd_rbm* rbm, kucd_tensor* dW, kucd_tensor* db, kucd_tensor* dc,
                        kucd_tensor* h_pos, kucd_tensor* v_neg, kucd_tensor* h_neg);
/* Fine-tuning after the path (an extension: the reference stops at greedy pretraining, dbn.py:34-55).  One delta-rule
 * step of a directed sigmoid layer that shares the RBM's parameter layout - the building block of the up-down
 * (wake-sleep) fine-tuning of a stack (Hinton, Osindero & Teh 2006, appendix B):
 *   forward != 0:  p = sigmoid(in.W + c)     W += lr in^T (target - p)     c += lr sum_rows (target - p)
 *   forward == 0:  p = sigmoid(in.W^T + b)   W += lr (target - p)^T in     b += lr sum_rows (target - p)
 * in: (rows, V) forward, (rows, H) backward; target: (rows, H) forward, (rows, V) backward; normalize != 0 divides
 * the sums by rows.  Runs the CD path's own kernels (projection with the probability epilogue, the two-segment dW
 * contraction, the update).  Single rank. */
int kucd_rbm_delta_rule(kucd_rbm* rbm, int forward, const kucd_tensor* in, const kucd_tensor* target, float lr,
                        int normalize);
/* persistent chains (PCD): (n_chains, V) states; get/set for checkpointing and parity */
int kucd_rbm_set_chains(kucd_rbm* rbm, const kucd_tensor* v_chains);
int kucd_rbm_get_chains(kucd_rbm* rbm, kucd_tensor* v_chains);

/* ---- device-resident data sets: the minibatch loop of rbm.py:110-113,163,211-223 ------------------ */
/* Uploads (or adopts, when already on the GPU in engine layout) a (N, dim) matrix once. */
int kucd_dataset_create(kucd_ctx* ctx, const kucd_tensor* data, int compute, kucd_dataset** out);
int kucd_dataset_destroy(kucd_dataset* ds);
int kucd_dataset_shape(kucd_dataset* ds, int64_t* rows, int64_t* dim);
int kucd_dataset_read(kucd_dataset* ds, kucd_tensor* out);
/* Epoch shuffling (an extension: the reference walks the rows in order, rbm.py:218).  Row i of the result
 * is row perm(i) of `in`, perm = a keyed bijection of [0, rows): a 6-round balanced Feistel network over
 * the smallest even-width power of two >= rows with cycle walking, round keys = the first six words of
 * Philox4x32-10(counter = (0x53485546, j, epoch lo, epoch hi), key = seed), j = 0, 1, round function
 * murmur3's fmix32(R ^ k_r) (restated in oracle/cd_oracle.py:feistel_permutation).  *inout == NULL: a new
 * data set is created; otherwise *inout (same shape) is overwritten, so the captured step graph of
 * kucd_rbm_fit_epoch, which is keyed by the data set, stays valid from epoch to epoch. */
int kucd_dataset_shuffle(kucd_dataset* in, uint64_t seed, uint64_t epoch, kucd_dataset** inout);
/* One epoch over sequential slices of `batch` rows, remainder last, no shuffle (rbm.py:163,211,218);
 * every step is one replay of a captured CUDA graph.  With a data-parallel group attached every rank
 * passes its own shard of every global minibatch: `batch` is the per-rank row count and
 * `global_row0` = rank * batch (on the remainder minibatch the engine keys the draws by
 * rank * remaining rows, so n ranks still sample exactly what one rank would). */
int kucd_rbm_fit_epoch(kucd_rbm* rbm, kucd_dataset* ds, int64_t batch, const kucd_hparams* hp,
                       int64_t global_row0, kucd_epoch_stats* stats);
/* Minibatches [step_begin, step_end) of that epoch (step_end < 0: to the end). */
int kucd_rbm_fit_range(kucd_rbm* rbm, kucd_dataset* ds, int64_t batch, const kucd_hparams* hp,
                       int64_t global_row0, int64_t step_begin, int64_t step_end, kucd_epoch_stats* stats);
/* The same loop over a HOST-resident matrix (the numpy array RBM.fit receives, rbm.py:100): minibatch
 * i+1 is copied to the GPU while minibatch i runs.  step_recon (nullable, one float per minibatch) receives
 * every step's reconstruction error, read back asynchronously. */
int kucd_rbm_fit_host(kucd_rbm* rbm, const kucd_tensor* V_all, int64_t batch, const kucd_hparams* hp,
                      int64_t global_row0, float* step_recon, kucd_epoch_stats* stats);
/* DBN.fit's inter-layer step (dbn.py:55): V_p <- transform(V_p) on the whole data set, on device. */
int kucd_rbm_transform_dataset(kucd_rbm* rbm, kucd_dataset* in, kucd_dataset** out);
int kucd_rbm_inv_transform_dataset(kucd_rbm* rbm, kucd_dataset* in, kucd_dataset** out);

#ifdef __cplusplus
}
#endif
#endif /* KUCD_H_ */
