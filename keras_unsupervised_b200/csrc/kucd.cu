// libkucd.so - host side of the contrastive-divergence engine and its C ABI (include/kucd.h).
//
// What the reference does with eight K.function graph executions per minibatch
// (/root/reference/ku/ebm/rbm.py:214-231) is here a fixed sequence of launches on one stream:
//
//   column sums of v0                              -> db            (rbm.py:134, positive half)
//   h0 = 1[u < sigmoid(v0.W + c)]       tcgen05    -> dc += sum h0  (rbm.py:120 = :46-47)
//   k x { v = 1[u < sigmoid(h.W^T + b)] ; h = sample or, last, sigmoid(v.W + c) }   (rbm.py:121-124)
//        - for bf16 Bernoulli training all 2k+1 projections are ONE persistent launch (chain.cuh);
//          otherwise one launch per projection, as two row-half chains on two streams when large
//   dW = v0^T h0 - vk^T hk              tcgen05, both phases in one TMEM accumulator (rbm.py:125-126)
//   exchange between data-parallel ranks: dW rows stored straight into their owner's memory by the
//        contraction's epilogue (peer-mapped, NVLink) + flag barriers, or ncclAllReduce(dW | db | dc)
//   W += lr dW ; b += lr db ; c += lr dc           one fused launch, bf16 operand planes refreshed (rbm.py:127-134)
//
// Parameters, operand planes, chains and workspaces stay in HBM between calls; under fit_epoch the
// sequence is captured once as a CUDA graph and replayed per minibatch, with the minibatch offset,
// the remainder-row count and the Philox draw counter living in device memory (StepDyn); fit_host
// streams a host array through a double-buffered staging area, copies overlapped with the chains.
#include "../../include/kucd.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "chain.cuh"
#include "launch.cuh"

using namespace kucd;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(KUCD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define KU_TRY(expr)         \
  do {                       \
    int rc_ = (expr);        \
    if (rc_ != KUCD_OK) return rc_; \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time (the process usually already holds torch's libnccl.so.2)
// ------------------------------------------------------------------------------------------------
struct NcclUid {
  char internal[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.lib != nullptr) return KUCD_OK;
  const char* env = getenv("KUCD_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so", "/usr/local/cuda/lib64/libnccl.so.2"};
  void* lib = nullptr;
  for (const char* n : names) {
    if (n == nullptr || *n == 0) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib != nullptr) break;
  }
  if (lib == nullptr) return fail(KUCD_ERR_NCCL, "libnccl.so.2 not found (set KUCD_NCCL_LIB): %s", dlerror());
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  g_nccl.Broadcast = reinterpret_cast<decltype(g_nccl.Broadcast)>(dlsym(lib, "ncclBroadcast"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(lib, "ncclAllGather"));
  g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(dlsym(lib, "ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce ||
      !g_nccl.GetErrorString || !g_nccl.Broadcast || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd)
    return fail(KUCD_ERR_NCCL, "libnccl lacks a required symbol");
  g_nccl.lib = lib;
  return KUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  // `shareable`: the buffer may be exported with cudaIpcGetMemHandle.  The driver carves small allocations out of
  // shared 2 MiB blocks and an IPC handle maps the whole block, so exported buffers get whole blocks of their own.
  int ensure(size_t n, bool zero = false, bool shareable = false) {
    if (shareable) n = std::max<size_t>((n + (2u << 20) - 1) / (2u << 20) * (2u << 20), 2u << 20);
    if (n <= bytes) return KUCD_OK;
    if (p != nullptr) cudaFree(p);
    p = nullptr;
    bytes = 0;
    CU_TRY(cudaMalloc(&p, n));
    bytes = n;
    if (zero) CU_TRY(cudaMemset(p, 0, n));
    return KUCD_OK;
  }
  // Stream-ordered variant (data-set planes): cudaMallocAsync / cudaFreeAsync on the engine stream.  The device's memory
  // pool keeps freed blocks (release threshold = unlimited, set at context creation), so a transform of a small data set
  // does not pay ~0.3 ms of cudaMalloc + a device-wide synchronising cudaFree around a 30 us contraction.
  cudaStream_t pool_stream = nullptr;
  int ensure_async(size_t n, cudaStream_t st) {
    if (n <= bytes) return KUCD_OK;
    release();
    CU_TRY(cudaMallocAsync(&p, n, st));
    bytes = n;
    pool_stream = st;
    return KUCD_OK;
  }
  void release() {
    if (p != nullptr) {
      if (pool_stream != nullptr) cudaFreeAsync(p, pool_stream);  // ordered after every kernel enqueued so far
      else cudaFree(p);
    }
    p = nullptr;
    bytes = 0;
    pool_stream = nullptr;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

// temporaries of one API call: released on every return path
struct ScopedBufs {
  std::vector<DevBuf> v;
  explicit ScopedBufs(size_t n = 0) : v(n) {}
  ScopedBufs(const ScopedBufs&) = delete;
  ScopedBufs& operator=(const ScopedBufs&) = delete;
  ~ScopedBufs() {
    for (auto& b : v) b.release();  // cudaFree waits for work that still reads the buffer
  }
  DevBuf& operator[](size_t i) { return v[i]; }
};

struct kucd_ctx {
  int device = 0;
  int num_sms = 0;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // second Gibbs chain of a split minibatch
  cudaStream_t stream3 = nullptr;  // unit-sharded step: the positive-phase projection, beside the negative chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_units[4] = {};  // unit-sharded step: exchanges on stream2 next to independent projections
  cudaStream_t copy_stream = nullptr;  // host -> device staging of the next minibatch (fit_host)
  // slab-pipelined all-reduce (KUCD_AR_SLABS): dW leaves in row slabs on comm_stream while the next slab is contracted
  static constexpr int kMaxSlabs = 16;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_slab[kMaxSlabs] = {}, ev_red[kMaxSlabs] = {};
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  DevBuf stage_raw[2];
  bool split = true;               // KUCD_SPLIT=0 turns the two-chain schedule off, 2 forces it (tests)
  bool split_force = false;
  bool chain_force = false;        // KUCD_CHAIN=2: use it at any size (tests)
  bool chain_dw = false;           // KUCD_CHAIN_DW=1
  bool chain = true;               // KUCD_CHAIN=0: launch the projections one by one instead of the chain kernel
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  kucd_timings tm{};
  void* comm = nullptr;
  int rank = 0, world = 1;
  // optional per-launch CUDA-event timing of the contractions (kucd_ctx_set_profile)
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;
  struct Mark {
    int cls;  // 0 = projection, 1 = dW, 2 = bit exchange, 3 = parameter update
    size_t e0, e1;
    int n;  // launches the interval covers
  };
  std::vector<Mark> marks;
  size_t ev_used = 0;
  DevBuf stage_in, stage_u, stage_out;  // raw caller-dtype staging for host tensors
  float* pinned_stats = nullptr;        // page-locked landing area of fit_host's per-step statistics (grown, never shrunk)
  size_t pinned_stats_cap = 0;          // floats
};

// up to three bf16 term planes of one (rows, cols) matrix, leading dimension ld
struct Planes {
  __nv_bfloat16* p[3] = {nullptr, nullptr, nullptr};
  int n = 1;  // live terms (1: the first plane is exact)
  int64_t rows = 0, cols = 0, ld = 0;
  __nv_bfloat16* mid() const { return n == 3 ? p[1] : nullptr; }
  __nv_bfloat16* lo() const { return n == 3 ? p[2] : nullptr; }
};

struct PlaneBuf {
  DevBuf buf[3];
  int64_t rows = 0, ld = 0;
  int ensure(int64_t r, int64_t ld_, int nplanes, bool shareable = false) {
    for (int i = 0; i < nplanes; ++i) KU_TRY(buf[i].ensure(static_cast<size_t>(r) * ld_ * 2, false, shareable));
    rows = r;
    ld = ld_;
    return KUCD_OK;
  }
  Planes view(int64_t r, int64_t cols, int n) const {
    Planes v;
    for (int i = 0; i < 3; ++i) v.p[i] = buf[i].as<__nv_bfloat16>();
    v.n = n;
    v.rows = r;
    v.cols = cols;
    v.ld = ld;
    return v;
  }
  void release() {
    for (auto& b : buf) b.release();
  }
};

static std::atomic<uint64_t> g_next_dataset_id{1};  // contexts on different host threads create data sets concurrently

struct kucd_dataset {
  kucd_ctx* ctx = nullptr;
  uint64_t id = g_next_dataset_id++;  // a freed data set's address may be handed out again: graphs are keyed by id
  PlaneBuf planes;
  int64_t rows = 0, dim = 0;
  int nparts = 1;  // live terms
  Planes view() const { return planes.view(rows, dim, nparts); }
};

struct GraphKey {
  const kucd_dataset* ds = nullptr;
  uint64_t ds_id = 0;
  const void* ds_ptr = nullptr;
  int64_t batch = 0, global_row0 = 0, ds_rows = 0;
  kucd_hparams hp{};
  int ds_parts = 0;
  bool fused = false;
  bool units = false;
  const void* chains_g = nullptr;
  bool operator==(const GraphKey& o) const {
    return ds == o.ds && ds_id == o.ds_id && ds_ptr == o.ds_ptr && batch == o.batch && global_row0 == o.global_row0 && ds_rows == o.ds_rows &&
           ds_parts == o.ds_parts && fused == o.fused && units == o.units && chains_g == o.chains_g &&
           memcmp(&hp, &o.hp, sizeof hp) == 0;
  }
};

// key of the captured step that a streamed fit (kucd_rbm_fit_host) replays over its full minibatches
struct HostGraphKey {
  int64_t batch = 0, global_row0 = 0;
  kucd_hparams hp{};
  bool fused = false;
  int live = 0;
  const void* vin = nullptr;
  const void* log = nullptr;
  int64_t chunk_rows = 0;  // > 0: the step reads a resident chunk of that many rows at StepDyn::row_off
  bool operator==(const HostGraphKey& o) const {
    return batch == o.batch && global_row0 == o.global_row0 && fused == o.fused && live == o.live && vin == o.vin &&
           log == o.log && chunk_rows == o.chunk_rows && memcmp(&hp, &o.hp, sizeof hp) == 0;
  }
};

struct kucd_rbm {
  kucd_ctx* ctx = nullptr;
  int64_t V = 0, H = 0, ldV = 0, ldH = 0;
  int mode = KUCD_MODE_VISIBLE_BERNOULLI;
  int compute = KUCD_COMPUTE_BF16;
  int wparts = 1;
  // parameters
  DevBuf W32;        // (V, ldH) fp32 master
  PlaneBuf Wp;       // bf16 operand planes of W
  DevBuf b32, c32;   // biases, zero-padded so an epilogue may read a whole tile of them
  DevBuf mW, mb, mc; // momentum (allocated on first use)
  // gradient block: [dW (V*ldH) | db (ldV) | dc (ldH)] - one all-reduce covers it
  DevBuf grad;
  // workspaces, sized for `cap` rows
  int64_t cap = 0;
  PlaneBuf vin, vin2, h0, hk, vk;  // vin2: second staging slot of the host-streaming fit
  PlaneBuf chunk;   // chunked streaming (KUCD_STREAM_CHUNK): operand planes of several staged minibatches
  PlaneBuf ft_in, ft_t, ft_p;  // kucd_rbm_delta_rule: input states, targets, predicted probabilities
  PlaneBuf chains;  // persistent chains (n_chains, ldV)
  int64_t n_chains = 0;
  DevBuf fe0, fe1, sp0, sp1, pstage, stats, flag;
  DevBuf dyn;
  DevBuf chain_done;  // (stage, row block) completion counters of the chain kernel
  // fused reduction over peer-mapped memory (data-parallel ranks on one NVLink domain)
  bool peer_on = false;    // peer memory is mapped
  bool fused_now = false;  // the current training call exchanges through it (decided per call, see choose_exchange)
  bool wire16 = false;     // partial dW sums cross NVLink as bf16 (KUCD_WIRE_BF16=1 at creation, bf16 compute; changes the
                           // arithmetic): bf16 slots in the fused exchange, a bf16 ncclAllReduce otherwise
  DevBuf grad16;           // (V, ldH) bf16 dW of this rank for the bf16 all-reduce
  int slabs_now = 1;       // > 1: this training call contracts, all-reduces and applies dW in that many row slabs
  int64_t slab_rows = 0;   // rows of W per slab (multiple of 256)
  int slabs_inflight = 0;  // slabs whose all-reduce was enqueued by enqueue_cd and not yet joined by apply_update
  DevBuf arena;               // [n dW slots | n bias slots | flags | epoch]
  PeerSet ps{};
  void* peer_open[16] = {};   // pointers obtained from cudaIpcOpenMemHandle (to be closed)
  int n_peer_open = 0;
  int64_t rows_per = 0, slice_elems = 0;
  int bias_len = 0;
  // unit-sharded exchange (enqueue_cd_units): this rank computes the hidden units [me*H/n, (me+1)*H/n) and the visible
  // units [me*V/n, (me+1)*V/n) for ALL rows of the global minibatch; 0/1 states travel between the ranks as bits
  bool units_ok = false;     // the arena has bit slots and V, H split into whole 128-unit groups per rank
  bool units_now = false;    // the current training call runs unit-sharded (decided per call, see choose_exchange)
  PlaneBuf chains_g;         // persistent chains of the GLOBAL minibatch (every rank holds all of them in this mode)
  bool chains_g_valid = false;
  int64_t chains_g_rows = 0;
  bool units_pcd = false;    // the last unit-sharded step ran persistent chains (its v_neg lives in chains_g)
  DevBuf gtmp;               // staging of the master gather at the end of a unit-sharded training call
  uint32_t* epoch = nullptr;
  uint64_t seed = 0;  // Philox key; draws are (seed, draw id, global row, column)
  uint64_t step_count = 0, infer_draws = 0, score_draws = 0;
  int64_t last_rows = 0;
  int last_vk_parts = 1, last_hk_parts = 1;
  // captured CD step
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  GraphKey graph_key;
  int64_t graph_kernels = 0;  // kernels one replay launches
  int64_t graph_unit_ex = 0, graph_fused = 0, graph_ar = 0;  // bit exchanges / fused exchanges / all-reduces per replay
  // captured step of the streamed fit (reads the staged minibatch in `vin`)
  cudaGraph_t hgraph = nullptr;
  cudaGraphExec_t hgraph_exec = nullptr;
  HostGraphKey hgraph_key;
  int64_t hgraph_kernels = 0;

  float* dW() const { return grad.as<float>(); }
  float* db() const { return grad.as<float>() + V * ldH; }
  float* dc() const { return grad.as<float>() + V * ldH + ldVb(); }
  int64_t ldVb() const { return round_up(V, 256) + 256; }
  int64_t ldHb() const { return round_up(H, 256) + 256; }
  int64_t grad_elems() const { return V * ldH + ldVb() + ldHb(); }
};

static void drop_host_graph(kucd_rbm* r) {
  if (r->hgraph_exec != nullptr) cudaGraphExecDestroy(r->hgraph_exec);
  if (r->hgraph != nullptr) cudaGraphDestroy(r->hgraph);
  r->hgraph_exec = nullptr;
  r->hgraph = nullptr;
}

// The planes of a data set come from the device's stream-ordered memory pool (DevBuf::ensure_async).  Every kernel that
// touches data-set planes runs in order on the context's stream (forked work joins it before a step ends), so a block
// freed in stream order can be handed out again without waiting for anything.
static int dataset_planes(kucd_ctx* ctx, PlaneBuf& pb, int64_t rows, int64_t ld, int nplanes) {
  for (int i = 0; i < nplanes; ++i) KU_TRY(pb.buf[i].ensure_async(static_cast<size_t>(rows) * ld * 2, ctx->stream));
  pb.rows = rows;
  pb.ld = ld;
  return KUCD_OK;
}
static void dataset_planes_release(kucd_ctx*, PlaneBuf& pb) { pb.release(); }

// this training call all-reduces dW as bf16 through NCCL (the fused exchange has its own bf16 slots)
static bool nccl16(const kucd_rbm* r) { return r->wire16 && !r->fused_now && !r->units_now && r->ctx->comm != nullptr; }

static bool capturing(const kucd_ctx* ctx) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return st != cudaStreamCaptureStatusNone;
}

static size_t prof_event(kucd_ctx* ctx, cudaStream_t st = nullptr) {
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev_pool.push_back(e);
  }
  cudaEventRecord(ctx->ev_pool[ctx->ev_used], st != nullptr ? st : ctx->stream);
  return ctx->ev_used++;
}

// fold the recorded marks into the timing totals (synchronises the stream)
static void prof_collect(kucd_ctx* ctx) {
  if (ctx->marks.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (const auto& m : ctx->marks) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_pool[m.e0], ctx->ev_pool[m.e1]) != cudaSuccess) continue;
    if (m.cls == 0) {
      ctx->tm.proj_ms += ms;
      ctx->tm.proj_timed += m.n;
    } else if (m.cls == 1) {
      ctx->tm.dw_ms += ms;
      ctx->tm.dw_timed += m.n;
    } else if (m.cls == 2) {
      ctx->tm.xchg_ms += ms;
      ctx->tm.xchg_timed += m.n;
    } else {
      ctx->tm.upd_ms += ms;
      ctx->tm.upd_timed += m.n;
    }
    ctx->tm.last_gemm_ms = ms;
  }
  ctx->marks.clear();
  ctx->ev_used = 0;
}

static int grid_for(const kucd_ctx* ctx, int64_t work_items, int threads) {
  const int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(ctx->num_sms) * 8;
  return static_cast<int>(std::max<int64_t>(1, std::min(blocks, cap)));
}

// ------------------------------------------------------------------------------------------------
// tensors at the boundary
// ------------------------------------------------------------------------------------------------
// bit-packed 0/1 matrix: strides[0] counts bits (a multiple of 8), column j is bit j % 8 of byte j / 8 of its row
static bool is_packed(const kucd_tensor* t) { return t->dtype_code == KUCD_DT_UINT && t->bits == 1; }
// bytes from one row to the next, and bytes one row occupies
static int64_t row_pitch_bytes(const kucd_tensor* t) {
  return is_packed(t) ? t->strides[0] / 8 : t->strides[0] * (t->bits / 8);
}
static int64_t row_bytes(const kucd_tensor* t) {
  return is_packed(t) ? (t->shape[1] + 7) / 8 : t->shape[1] * (t->bits / 8);
}

static bool dtype_ok(const kucd_tensor* t) {
  return (t->dtype_code == KUCD_DT_FLOAT && t->bits == 32) || (t->dtype_code == KUCD_DT_BFLOAT && t->bits == 16) ||
         (t->dtype_code == KUCD_DT_UINT && t->bits == 8) || is_packed(t);
}

// vectors may arrive as (n,1) with unit stride: view them as one row of n
static kucd_tensor as_matrix(const kucd_tensor* t) {
  kucd_tensor m = *t;
  if (m.shape[1] == 1 && m.shape[0] > 1 && m.strides[0] == 1) {
    m.shape[1] = m.shape[0];
    m.shape[0] = 1;
    m.strides[0] = m.shape[1];
    m.strides[1] = 1;
  }
  return m;
}

static int check_tensor(const kucd_ctx* ctx, const kucd_tensor* t, int64_t rows, int64_t cols, const char* name,
                        bool f32_only = false) {
  if (t == nullptr) return fail(KUCD_ERR_INVALID_ARG, "%s is NULL", name);
  if (t->data == nullptr && t->shape[0] * t->shape[1] != 0) return fail(KUCD_ERR_INVALID_ARG, "%s has no data", name);
  if (!dtype_ok(t))
    return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "%s: dtype code %d / %d bits is not float32, bfloat16, uint8 or packed bits",
                name, t->dtype_code, t->bits);
  if (is_packed(t) && (t->strides[0] % 8 != 0 || t->strides[0] < t->shape[1]) && t->shape[0] > 1)
    return fail(KUCD_ERR_INVALID_ARG, "%s: rows of a bit-packed matrix must start on byte boundaries", name);
  if (f32_only && !(t->dtype_code == KUCD_DT_FLOAT && t->bits == 32))
    return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "%s must be float32", name);
  if (t->device_type != KUCD_DEV_CPU && t->device_type != KUCD_DEV_CUDA && t->device_type != KUCD_DEV_CUDA_HOST)
    return fail(KUCD_ERR_INVALID_ARG, "%s: unsupported device type %d", name, t->device_type);
  if (t->device_type == KUCD_DEV_CUDA && t->device_id != ctx->device)
    return fail(KUCD_ERR_INVALID_ARG, "%s lives on GPU %d, the context on GPU %d", name, t->device_id, ctx->device);
  if (t->shape[1] > 1 && t->strides[1] != 1) return fail(KUCD_ERR_INVALID_ARG, "%s: innermost stride must be 1", name);
  if ((rows >= 0 && t->shape[0] != rows) || (cols >= 0 && t->shape[1] != cols))
    return fail(KUCD_ERR_SHAPE_MISMATCH, "%s has shape (%lld, %lld), expected (%lld, %lld)", name,
                (long long)t->shape[0], (long long)t->shape[1], (long long)rows, (long long)cols);
  return KUCD_OK;
}

static bool on_device(const kucd_tensor* t) { return t->device_type == KUCD_DEV_CUDA; }

// rows [r0, r0+n) of a caller tensor as a device pointer + row stride (elements)
static int fetch_rows(kucd_ctx* ctx, const kucd_tensor* t, int64_t r0, int64_t n, DevBuf& staging, const void** ptr,
                      int64_t* ld) {
  const int64_t pitch = row_pitch_bytes(t), rb = row_bytes(t);
  const char* src = static_cast<const char*>(t->data) + r0 * pitch;
  if (on_device(t)) {
    *ptr = src;
    *ld = t->strides[0];
    return KUCD_OK;
  }
  KU_TRY(staging.ensure(static_cast<size_t>(n) * rb));
  if (pitch == rb || n == 1)
    CU_TRY(cudaMemcpyAsync(staging.p, src, static_cast<size_t>(n) * rb, cudaMemcpyHostToDevice, ctx->stream));
  else
    CU_TRY(cudaMemcpy2DAsync(staging.p, rb, src, pitch, rb, n, cudaMemcpyHostToDevice, ctx->stream));
  ctx->tm.h2d_bytes += n * rb;
  *ptr = staging.p;
  *ld = is_packed(t) ? rb * 8 : t->shape[1];  // elements (bits, when packed) from one staged row to the next
  return KUCD_OK;
}

template <typename F>
static int by_dtype(const kucd_tensor* t, F&& f) {
  if (t->dtype_code == KUCD_DT_FLOAT) return f(static_cast<float*>(nullptr));
  if (t->dtype_code == KUCD_DT_BFLOAT) return f(static_cast<__nv_bfloat16*>(nullptr));
  return f(static_cast<uint8_t*>(nullptr));
}

// caller rows -> bf16 planes in `dst` (rows [0,n) of dst).  nparts: planes to write.
static int ingest_rows(kucd_ctx* ctx, const kucd_tensor* t, int64_t r0, int64_t n, const Planes& dst, int nparts,
                       int* inexact_flag_dev) {
  if (n == 0) return KUCD_OK;
  const void* src;
  int64_t ld;
  KU_TRY(fetch_rows(ctx, t, r0, n, ctx->stage_in, &src, &ld));
  const int64_t groups = n * (dst.ld / 8);
  const int grid = grid_for(ctx, groups, 256);
  if (is_packed(t)) {
    ingest_bits_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(src), ld / 8, n, t->shape[1], dst.p[0],
                                                      dst.p[1], dst.p[2], dst.ld, nparts);
    ctx->tm.aux_launches++;
    CU_TRY(cudaGetLastError());
    return KUCD_OK;
  }
  by_dtype(t, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    ingest_kernel<T><<<grid, 256, 0, ctx->stream>>>(static_cast<const T*>(src), ld, n, t->shape[1], dst.p[0], dst.p[1],
                                                    dst.p[2], dst.ld, nparts, inexact_flag_dev);
    return 0;
  });
  ctx->tm.aux_launches++;
  CU_TRY(cudaGetLastError());
  return KUCD_OK;
}

// device rows -> caller tensor rows [r0, r0+n)
static int deliver_planes(kucd_ctx* ctx, const Planes& src, int64_t n, kucd_tensor* out, int64_t r0) {
  if (n == 0) return KUCD_OK;
  const int64_t pitch = row_pitch_bytes(out), rb = row_bytes(out);
  const int64_t cols = out->shape[1];
  char* dst = static_cast<char*>(out->data) + r0 * pitch;
  void* kdst = dst;
  int64_t kpitch = pitch;  // bytes
  if (!on_device(out)) {
    KU_TRY(ctx->stage_out.ensure(static_cast<size_t>(n) * rb));
    kdst = ctx->stage_out.p;
    kpitch = rb;
  }
  if (is_packed(out)) {  // 0/1 states as bits
    export_bits_kernel<<<grid_for(ctx, n * rb, 256), 256, 0, ctx->stream>>>(src.p[0], src.ld, n, cols,
                                                                            static_cast<uint8_t*>(kdst), kpitch);
  } else {
    const int grid = grid_for(ctx, n * cols, 256);
    const int64_t kld = kpitch / (out->bits / 8);
    by_dtype(out, [&](auto* tag) {
      using T = std::remove_pointer_t<decltype(tag)>;
      export_kernel<T><<<grid, 256, 0, ctx->stream>>>(src.p[0], src.mid(), src.lo(), src.ld, n, cols, static_cast<T*>(kdst),
                                                      kld);
      return 0;
    });
  }
  ctx->tm.aux_launches++;
  CU_TRY(cudaGetLastError());
  if (!on_device(out)) {
    CU_TRY(cudaMemcpy2DAsync(dst, pitch, kdst, rb, rb, n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->tm.d2h_bytes += n * rb;
    CU_TRY(cudaStreamSynchronize(ctx->stream));  // stage_out is reused by the next delivery
  }
  return KUCD_OK;
}

static int deliver_f32(kucd_ctx* ctx, const float* src, int64_t ld, int64_t n, kucd_tensor* out, int64_t r0) {
  if (n == 0) return KUCD_OK;
  if (is_packed(out)) return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "real-valued outputs cannot be delivered as packed bits");
  const int es = out->bits / 8;
  const int64_t cols = out->shape[1];
  char* dst = static_cast<char*>(out->data) + r0 * out->strides[0] * es;
  if (cols == 1 && ld == 1 && out->strides[0] == 1 && out->dtype_code == KUCD_DT_FLOAT) {  // a plain vector
    CU_TRY(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDefault, ctx->stream));
    if (!on_device(out)) ctx->tm.d2h_bytes += n * 4;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return KUCD_OK;
  }
  if (!on_device(out) && out->dtype_code == KUCD_DT_FLOAT) {  // straight strided copy
    CU_TRY(cudaMemcpy2DAsync(dst, out->strides[0] * es, src, ld * 4, cols * 4, n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->tm.d2h_bytes += n * cols * 4;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return KUCD_OK;
  }
  void* kdst = dst;
  int64_t kld = out->strides[0];
  if (!on_device(out)) {
    KU_TRY(ctx->stage_out.ensure(static_cast<size_t>(n) * cols * es));
    kdst = ctx->stage_out.p;
    kld = cols;
  }
  const int grid = grid_for(ctx, n * cols, 256);
  by_dtype(out, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    export_f32_kernel<T><<<grid, 256, 0, ctx->stream>>>(src, ld, n, cols, static_cast<T*>(kdst), kld);
    return 0;
  });
  ctx->tm.aux_launches++;
  CU_TRY(cudaGetLastError());
  if (!on_device(out)) {
    CU_TRY(cudaMemcpy2DAsync(dst, out->strides[0] * es, kdst, cols * es, cols * es, n, cudaMemcpyDeviceToHost,
                             ctx->stream));
    ctx->tm.d2h_bytes += n * cols * es;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return KUCD_OK;
}

// injected uniforms: float32 rows on the device
static int fetch_u(kucd_ctx* ctx, const kucd_tensor* u, int64_t r0, int64_t n, int64_t cols, const float** ptr,
                   int64_t* ld) {
  *ptr = nullptr;
  *ld = 0;
  if (u == nullptr) return KUCD_OK;
  KU_TRY(check_tensor(ctx, u, -1, cols, "injected draws", true));
  if (u->shape[0] < r0 + n) return fail(KUCD_ERR_SHAPE_MISMATCH, "injected draws have too few rows");
  const void* p;
  KU_TRY(fetch_rows(ctx, u, r0, n, ctx->stage_u, &p, ld));
  *ptr = static_cast<const float*>(p);
  return KUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// contractions
// ------------------------------------------------------------------------------------------------
// term pairs (ia, ib) with ia + ib <= 2: the products that matter at fp32 precision.  Smallest terms
// first: the tensor core truncates when it aligns an addend to the running fp32 accumulator, so the
// 2^-16 and 2^-8 terms are summed while the accumulator is still small and only the leading hi*hi
// products are added at full magnitude.
static int term_pairs(int na, int nb, int (*out)[2], int only_order = -1) {
  int n = 0;
  for (int s = 2; s >= 0; --s) {
    if (only_order >= 0 && s != only_order) continue;
    for (int ia = 0; ia <= s; ++ia) {
      const int ib = s - ia;
      if (ia < na && ib < nb) {
        out[n][0] = ia;
        out[n][1] = ib;
        ++n;
      }
    }
  }
  return n;
}

struct EpiArgs {
  int epi = kEpiSample;
  Planes out;                  // state / probability planes
  float* out_f32 = nullptr;    // optional fp32 copy (probabilities or raw)
  int64_t ld_f32 = 0;
  const float* u = nullptr;    // injected draws
  int64_t ld_u = 0;
  float* colsum = nullptr;
  float colsum_sign = 1.f;
  int64_t colsum_rows = 0;  // rows adding into `colsum` over all launches that share it (0: the rows of this launch)
  float* rowsum = nullptr;
  uint64_t draw = 0;
  uint64_t draw_stride = 0;
  int64_t row0 = 0;
  int32_t m_valid = -1;
  const StepDyn* dyn = nullptr;
  bool a_dyn = false;  // A rows are offset by dyn->row_off (A is a whole data set)
  int32_t row_base = 0;  // first minibatch row this launch covers (second chain of a split minibatch)
  cudaStream_t stream = nullptr;  // default: the context stream
  bool no_prof = false;
  // unit-sharded launches: only the output units [n_lo, n_lo + n_cnt) are computed (n_cnt = 0: all of them); `out`,
  // the bias and the column sums are still addressed by absolute unit index.  static_rows: with `dyn` set, only the
  // draw counter comes from the device-resident step state - rows and row indices are the host's (the launch covers the
  // gathered GLOBAL minibatch, not this rank's shard).
  int64_t n_lo = 0, n_cnt = 0;
  bool static_rows = false;
};

// forward: (rows,V).W + c -> (rows,H) ; backward: (rows,H).W^T + b -> (rows,V)
static int project(kucd_rbm* r, bool forward, const Planes& a, int64_t rows, const EpiArgs& e) {
  kucd_ctx* ctx = r->ctx;
  GemmOperands ops;
  ops.a_mn = false;
  ops.b_mn = forward;  // W is (V,H) row-major: (K,N) forward, (N,K) backward
  ops.M = rows;
  const int64_t n_all = forward ? r->H : r->V;
  const int64_t n_lo = e.n_cnt > 0 ? e.n_lo : 0, n_cnt = e.n_cnt > 0 ? e.n_cnt : n_all;
  ops.N = n_cnt;
  ops.K = forward ? r->V : r->H;
  int pairs[9][2];
  const int np = term_pairs(a.n, r->wparts, pairs);
  ops.num_seg = np;
  uint32_t dyn_mask = 0;
  for (int s = 0; s < np; ++s) {
    ops.a[s] = MatView{a.p[pairs[s][0]], a.rows, a.cols, a.ld};
    const __nv_bfloat16* w = r->Wp.buf[pairs[s][1]].as<__nv_bfloat16>();
    // the units of this launch: columns of W going forward ((K, N) view), rows of W going back ((N, K) view)
    ops.b[s] = forward ? MatView{w + n_lo, r->V, n_cnt, r->ldH} : MatView{w + n_lo * r->ldH, n_cnt, r->H, r->ldH};
    if (e.a_dyn) dyn_mask |= 1u << s;
  }
  GemmParams p;
  memset(&p, 0, sizeof p);
  p.bias = (forward ? r->c32.as<float>() : r->b32.as<float>()) + n_lo;
  p.col_off = static_cast<int32_t>(n_lo);
  p.out_bf16 = e.out.p[0] + n_lo;
  p.out_mid = (e.out.n == 3) ? e.out.p[1] + n_lo : nullptr;
  p.out_lo = (e.out.n == 3) ? e.out.p[2] + n_lo : nullptr;
  p.ld_bf16 = e.out.ld;
  p.out_f32 = e.out_f32 != nullptr ? e.out_f32 + n_lo : nullptr;
  p.ld_f32 = e.ld_f32;
  p.u_inject = e.u;
  p.ld_u = e.ld_u;
  p.colsum = e.colsum != nullptr ? e.colsum + n_lo : nullptr;
  p.colsum_sign = e.colsum_sign;
  p.colsum_rows = static_cast<int32_t>(e.colsum_rows > 0 ? e.colsum_rows : rows);
  p.rowsum = e.rowsum;
  p.seed = r->seed;
  p.draw = e.draw;
  p.draw_stride = e.draw_stride;
  p.row0 = e.row0;
  p.m_valid = e.m_valid < 0 ? static_cast<int32_t>(rows) : e.m_valid;
  p.dyn = e.dyn;
  p.dyn_rows = (e.dyn != nullptr && !e.static_rows) ? 1 : 0;
  p.dyn_rank = ctx->rank;
  p.dyn_row_base = e.row_base;
  p.a_dyn_mask = dyn_mask;
  std::string err;
  const bool prof = ctx->profile && e.dyn == nullptr && !e.no_prof;
  const size_t pe0 = prof ? prof_event(ctx, e.stream) : 0;
  if (!launch_gemm(p, ops, e.epi, ctx->num_sms, e.stream != nullptr ? e.stream : ctx->stream, &err, 0,
                   r->compute == KUCD_COMPUTE_F32X3))
    return fail(KUCD_ERR_CUDA, "%s", err.c_str());
  if (prof) ctx->marks.push_back({0, pe0, prof_event(ctx, e.stream), 1});
  ctx->tm.gemm_launches++;
  return KUCD_OK;
}

// dW = v0^T h0 - vk^T hk   (rbm.py:125-126): contraction over the minibatch rows, every operand read
// in place (MN-major descriptors), both phases accumulated into the same tensor-memory tile.
// [w_row0, w_row0 + w_rows): the rows of W this launch covers (a column range of the visible states); sm_reserve: SMs
// left to a concurrent collective (slab-pipelined all-reduce).
static int delta_w(kucd_rbm* r, const Planes& v0, const Planes& h0, const Planes& vk, const Planes& hk, int64_t rows,
                   const StepDyn* dyn, bool v0_dyn, int64_t w_row0 = 0, int64_t w_rows = -1, int sm_reserve = 0,
                   int64_t h_lo = 0, int64_t h_cnt = -1) {
  kucd_ctx* ctx = r->ctx;
  if (w_rows < 0) w_rows = r->V - w_row0;
  if (h_cnt < 0) h_cnt = r->H - h_lo;  // [h_lo, h_lo + h_cnt): the columns of W this launch covers (unit-sharded exchange)
  GemmOperands ops;
  ops.a_mn = true;
  ops.b_mn = true;
  ops.M = w_rows;
  ops.N = h_cnt;
  ops.K = rows;
  int pairs[9][2];
  int ns = 0;
  uint32_t neg = 0, dyn_mask = 0;
  for (int order = 2; order >= 0; --order) {  // small terms of both phases first, see term_pairs
    int np = term_pairs(v0.n, h0.n, pairs, order);
    if (ns + np > kMaxSeg) return fail(KUCD_ERR_INVALID_ARG, "too many operand terms");
    for (int s = 0; s < np; ++s, ++ns) {
      ops.a[ns] = MatView{v0.p[pairs[s][0]] + w_row0, v0.rows, w_rows, v0.ld};
      ops.b[ns] = MatView{h0.p[pairs[s][1]] + h_lo, h0.rows, h_cnt, h0.ld};
      if (v0_dyn) dyn_mask |= 1u << ns;
    }
    np = term_pairs(vk.n, hk.n, pairs, order);
    if (ns + np > kMaxSeg) return fail(KUCD_ERR_INVALID_ARG, "too many operand terms");
    for (int s = 0; s < np; ++s, ++ns) {
      ops.a[ns] = MatView{vk.p[pairs[s][0]] + w_row0, vk.rows, w_rows, vk.ld};
      ops.b[ns] = MatView{hk.p[pairs[s][1]] + h_lo, hk.rows, h_cnt, hk.ld};
      neg |= 1u << ns;
    }
  }
  ops.num_seg = ns;
  ops.neg_mask = neg;
  GemmParams p;
  memset(&p, 0, sizeof p);
  p.out_f32 = r->dW() + w_row0 * r->ldH + h_lo;
  p.ld_f32 = r->ldH;
  p.m_valid = static_cast<int32_t>(w_rows);
  p.dyn = dyn;
  p.a_dyn_mask = dyn_mask;
  int epi = kEpiRaw;
  if (r->fused_now) {  // each output row goes straight into its owner's slot for this rank (whole-matrix launches only)
    p.push_rows = static_cast<int32_t>(r->rows_per);
    for (int o = 0; o < ctx->world; ++o) {
      if (r->wire16)  // bf16 slots: same element pitch, half the bytes
        p.push_base[o] = reinterpret_cast<float*>(reinterpret_cast<__nv_bfloat16*>(r->ps.dw_slot[o]) +
                                                  ctx->rank * r->slice_elems);
      else
        p.push_base[o] = r->ps.dw_slot[o] + ctx->rank * r->slice_elems;
    }
    if (r->wire16) epi = kEpiRawPush16;
  } else if (nccl16(r)) {  // the whole dW as bf16 into the local all-reduce buffer: one "owner" holding every row
    p.push_rows = static_cast<int32_t>(w_rows);
    p.push_base[0] = reinterpret_cast<float*>(r->grad16.as<__nv_bfloat16>() + w_row0 * r->ldH);
    epi = kEpiRawPush16;
  }
  std::string err;
  const bool prof = ctx->profile && dyn == nullptr;
  const size_t pe0 = prof ? prof_event(ctx) : 0;
  if (!launch_gemm(p, ops, epi, std::max(2, ctx->num_sms - sm_reserve), ctx->stream, &err, 0,
                   r->compute == KUCD_COMPUTE_F32X3))
    return fail(KUCD_ERR_CUDA, "%s", err.c_str());
  if (prof) ctx->marks.push_back({1, pe0, prof_event(ctx), 1});
  ctx->tm.gemm_launches++;
  return KUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// workspaces
// ------------------------------------------------------------------------------------------------
static int ensure_workspace(kucd_rbm* r, int64_t rows) {
  if (rows <= r->cap) return KUCD_OK;
  const int64_t cap = round_up(rows, 128);
  const int np = 3;  // keep all term planes available: Gaussian visibles and fp32 probabilities need them
  KU_TRY(r->vin.ensure(cap, r->ldV, r->compute == KUCD_COMPUTE_F32X3 ? np : 1));
  KU_TRY(r->h0.ensure(cap, r->ldH, 1));
  KU_TRY(r->hk.ensure(cap, r->ldH, r->compute == KUCD_COMPUTE_F32X3 ? np : 1));
  KU_TRY(r->vk.ensure(cap, r->ldV,
                      (r->compute == KUCD_COMPUTE_F32X3 && r->mode == KUCD_MODE_VISIBLE_GAUSSIAN) ? np : 1));
  KU_TRY(r->fe0.ensure(cap * 4));
  KU_TRY(r->fe1.ensure(cap * 4));
  KU_TRY(r->sp0.ensure(cap * 4));
  KU_TRY(r->sp1.ensure(cap * 4));
  KU_TRY(r->pstage.ensure(static_cast<size_t>(cap) * std::max(r->ldV, r->ldH) * 4));
  r->cap = cap;
  // a captured graph holds the old pointers
  drop_host_graph(r);
  if (r->graph_exec != nullptr) {
    cudaGraphExecDestroy(r->graph_exec);
    cudaGraphDestroy(r->graph);
    r->graph_exec = nullptr;
    r->graph = nullptr;
  }
  return KUCD_OK;
}

static int vis_parts_out(const kucd_rbm* r) {
  return (r->compute == KUCD_COMPUTE_F32X3 && r->mode == KUCD_MODE_VISIBLE_GAUSSIAN) ? 3 : 1;
}
static int prob_parts_out(const kucd_rbm* r) { return r->compute == KUCD_COMPUTE_F32X3 ? 3 : 1; }

// ------------------------------------------------------------------------------------------------
// the CD step (all launches on ctx->stream; capturable)
// ------------------------------------------------------------------------------------------------
struct StepInject {
  const float* u_h[KUCD_MAX_K] = {};
  int64_t ld_h[KUCD_MAX_K] = {};
  const float* u_v[KUCD_MAX_K] = {};
  int64_t ld_v[KUCD_MAX_K] = {};
  const float* u_hc = nullptr;
  int64_t ld_hc = 0;
};

// `adv`: when the caller replays this step as a graph, the step-state advance rides on the update launch
// (returns *adv_done = true) instead of being a launch of its own.
static int apply_update(kucd_rbm* r, const kucd_hparams* hp, int64_t rows_global, StepDyn* adv = nullptr,
                        int32_t adv_batch = 0, int64_t adv_total = 0, bool* adv_done = nullptr) {
  kucd_ctx* ctx = r->ctx;
  if (adv_done != nullptr) *adv_done = false;
  const float scale = hp->normalize ? 1.0f / static_cast<float>(std::max<int64_t>(rows_global, 1)) : 1.0f;
  // replayed step + mean normalisation: the kernels divide by the current minibatch's rows (device state); the
  // step-state advance then has to be a launch of its own, after every reader
  const StepDyn* sdyn = (hp->normalize && adv != nullptr) ? adv : nullptr;
  const float world = static_cast<float>(ctx->world);
  if (sdyn != nullptr) adv = nullptr;
  const bool use_mom = hp->momentum != 0.f;
  if (use_mom) {
    KU_TRY(r->mW.ensure(static_cast<size_t>(r->V) * r->ldH * 4, true));
    KU_TRY(r->mb.ensure(r->ldVb() * 4, true));
    KU_TRY(r->mc.ensure(r->ldHb() * 4, true));
  }
  if (r->units_now) {
    // unit-sharded step: this rank owns the columns [h_lo, h_lo + Hs) of W, c[h_lo ...] and b[v_lo ...]; its statistics
    // for them already cover the whole global minibatch, so there is nothing to reduce
    const int n = ctx->world, me = ctx->rank;
    const int64_t Hs = r->H / n, Vs = r->V / n, h_lo = me * Hs, v_lo = me * Vs;
    const bool prof = ctx->profile && !capturing(ctx);
    const size_t pe0 = prof ? prof_event(ctx) : 0;
    if (hp->update_mask & KUCD_UPDATE_W) {
      update_w_units_kernel<<<grid_for(ctx, r->V * Hs / 4, 256), 256, 0, ctx->stream>>>(
          r->W32.as<float>(), r->dW(), use_mom ? r->mW.as<float>() : nullptr, r->ps, me, r->ldH, r->V, h_lo, Hs, Vs, hp->lr,
          scale, hp->momentum, hp->weight_decay, sdyn, world);
      ctx->tm.aux_launches++;
    }
    if (hp->update_mask & KUCD_UPDATE_C) {
      update_bias_kernel<<<(Hs + 255) / 256, 256, 0, ctx->stream>>>(r->c32.as<float>() + h_lo, r->dc() + h_lo,
                                                                    use_mom ? r->mc.as<float>() + h_lo : nullptr, Hs, hp->lr,
                                                                    scale, hp->momentum, sdyn, world);
      ctx->tm.aux_launches++;
    }
    if (hp->update_mask & KUCD_UPDATE_B) {
      update_bias_kernel<<<(Vs + 255) / 256, 256, 0, ctx->stream>>>(r->b32.as<float>() + v_lo, r->db() + v_lo,
                                                                    use_mom ? r->mb.as<float>() + v_lo : nullptr, Vs, hp->lr,
                                                                    scale, hp->momentum, sdyn, world);
      ctx->tm.aux_launches++;
    }
    // closes the step: every rank's weights for this rank's back-projection have landed, and every rank is done with the
    // bit slots (the next step starts again at slot 0)
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(r->epoch, r->ps, me, n);
    ctx->tm.aux_launches++;
    if (prof) ctx->marks.push_back({3, pe0, prof_event(ctx), 1});
    CU_TRY(cudaGetLastError());
    return KUCD_OK;
  }
  if (r->fused_now) {
    // global [db | dc] = sum of the ranks' slots; the biases are updated identically on every rank
    const int n = ctx->world, me = ctx->rank;
    reduce_bias_kernel<<<(r->bias_len + 255) / 256, 256, 0, ctx->stream>>>(r->ps.bias_slot[me], n, r->bias_len, r->db());
    ctx->tm.aux_launches++;
    if (hp->update_mask & KUCD_UPDATE_W) {
      const int64_t r0 = me * r->rows_per;
      const int64_t rows = std::max<int64_t>(0, std::min<int64_t>(r->rows_per, r->V - r0));
      const int64_t n4 = rows * r->ldH / 4;
      if (n4 > 0) {
        auto kern = r->wire16 ? update_w_sharded_kernel<true> : update_w_sharded_kernel<false>;
        kern<<<grid_for(ctx, n4, 256), 256, 0, ctx->stream>>>(
            r->W32.as<float>(), r->ps.dw_slot[me], r->slice_elems, n, use_mom ? r->mW.as<float>() : nullptr, r->ps,
            r0 * r->ldH, n4, hp->lr, scale, hp->momentum, hp->weight_decay, sdyn, world);
        ctx->tm.aux_launches++;
      }
    }
  } else {
    // one launch: W (if selected), both biases (if selected) and, under graph replay, the step-state advance
    const int64_t n4 = (hp->update_mask & KUCD_UPDATE_W) ? r->V * r->ldH / 4 : 0;
    UpdateTail tail{};
    if (hp->update_mask & KUCD_UPDATE_B) {
      tail.b = r->b32.as<float>();
      tail.db = r->db();
      tail.mb = use_mom ? r->mb.as<float>() : nullptr;
      tail.nb = static_cast<int32_t>(r->V);
    }
    if (hp->update_mask & KUCD_UPDATE_C) {
      tail.c = r->c32.as<float>();
      tail.dc = r->dc();
      tail.mc = use_mom ? r->mc.as<float>() : nullptr;
      tail.nc = static_cast<int32_t>(r->H);
    }
    tail.dyn = adv;
    tail.adv_batch = adv_batch;
    tail.adv_total = adv_total;
    const bool d16 = nccl16(r);
    auto kern = d16 ? update_w_kernel<true> : update_w_kernel<false>;
    if (r->slabs_inflight > 0) {
      // slab-pipelined exchange: slab i is updated as soon as its all-reduce (comm stream) is done, while the later
      // slabs are still on the wire; the bias statistics travelled with slab 0, the tail rides on the last slab
      const int S = r->slabs_inflight;
      r->slabs_inflight = 0;
      if (n4 == 0) {  // W is not updated by this call: just join the collective stream
        CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_red[S - 1], 0));
      } else {
        for (int i = 0; i < S; ++i) {
          CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_red[i], 0));
          const int64_t w0 = i * r->slab_rows, wn = std::min<int64_t>(r->slab_rows, r->V - w0);
          const int64_t off = w0 * r->ldH, m4 = wn * r->ldH / 4;
          const float* dsrc = d16 ? reinterpret_cast<const float*>(r->grad16.as<__nv_bfloat16>() + off) : r->dW() + off;
          kern<<<grid_for(ctx, m4, 256), 256, 0, ctx->stream>>>(
              r->W32.as<float>() + off, dsrc, use_mom ? r->mW.as<float>() + off : nullptr,
              r->Wp.buf[0].as<__nv_bfloat16>() + off, nullptr, nullptr, m4, hp->lr, scale, hp->momentum, hp->weight_decay,
              i == S - 1 ? tail : UpdateTail{}, sdyn, world);
          ctx->tm.aux_launches++;
        }
        if (adv_done != nullptr) *adv_done = adv != nullptr;
        CU_TRY(cudaGetLastError());
        return KUCD_OK;
      }
    }
    kern<<<grid_for(ctx, std::max<int64_t>(n4, 1), 256), 256, 0, ctx->stream>>>(
        r->W32.as<float>(), d16 ? r->grad16.as<float>() : r->dW(), use_mom ? r->mW.as<float>() : nullptr,
        r->Wp.buf[0].as<__nv_bfloat16>(),
        r->wparts == 3 ? r->Wp.buf[1].as<__nv_bfloat16>() : nullptr,
        r->wparts == 3 ? r->Wp.buf[2].as<__nv_bfloat16>() : nullptr, n4, hp->lr, scale, hp->momentum, hp->weight_decay,
        tail, sdyn, world);
    ctx->tm.aux_launches++;
    if (adv_done != nullptr) *adv_done = adv != nullptr;
    CU_TRY(cudaGetLastError());
    return KUCD_OK;
  }
  // fused exchange: the biases are updated redundantly on every rank from the reduced statistics
  if (hp->update_mask & KUCD_UPDATE_C) {
    update_bias_kernel<<<(r->H + 255) / 256, 256, 0, ctx->stream>>>(r->c32.as<float>(), r->dc(),
                                                                     use_mom ? r->mc.as<float>() : nullptr, r->H, hp->lr,
                                                                     scale, hp->momentum, sdyn, world);
    ctx->tm.aux_launches++;
  }
  if (hp->update_mask & KUCD_UPDATE_B) {
    update_bias_kernel<<<(r->V + 255) / 256, 256, 0, ctx->stream>>>(r->b32.as<float>(), r->db(),
                                                                     use_mom ? r->mb.as<float>() : nullptr, r->V, hp->lr,
                                                                     scale, hp->momentum, sdyn, world);
    ctx->tm.aux_launches++;
  }
  if (r->fused_now) {
    // every rank's rows of the new bf16 W must have landed in this rank's operand plane before the next chain
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(r->epoch, r->ps, ctx->rank, ctx->world);
    ctx->tm.aux_launches++;
  }
  CU_TRY(cudaGetLastError());
  return KUCD_OK;
}

// Fused exchange or NCCL all-reduce for this training call?  The dW contraction can hide its NVLink stores only if
// it runs longer than they take: time_gemm = 4 B V H / ~1.2e15 s against 4 V H (n-1)/n / ~6e11 s of stores, i.e. a
// per-rank minibatch of ~2000 rows.  Measured: C3 (4096 rows per rank) 2.59 ms fused vs 2.76 ms NCCL at 8 ranks; C4
// (1024 rows per rank, 512 MiB of dW) 3.13 ms fused vs 2.95 ms NCCL.  Fixed for the whole call: the two paths keep the
// fp32 master differently between steps.  KUCD_FUSED_MIN_ROWS overrides the threshold.
// With bf16 partial sums (KUCD_WIRE_BF16=1) half the bytes cross, so the break-even moves to ~1000 rows per rank.
//
// Third option, the UNIT-SHARDED step (enqueue_cd_units): no dW crosses the wire at all.  Rank r computes the hidden units
// [r H/n, (r+1) H/n) and the visible units [r V/n, (r+1) V/n) for all rows of the GLOBAL minibatch, the sampled 0/1
// states are exchanged as bits after every projection, dW and the update are local to the owned columns of W, and only
// the refreshed bf16 weights a peer needs for its back-projection travel (V x H/n x 2 bytes per rank and step).  Against
// the reduce-scatter + all-gather of V x H partial sums it wins when W is large next to the minibatch - BASELINE's C4
// (16384 x 8192, 8192 rows over 8 GPUs: 512 MiB of fp32 dW per rank and step against ~60 MiB of bits and weights) - and
// loses when the minibatch is (C3 weak scaling: 21 exchanges of a 32768-row state matrix).  The estimate below compares
// the two; KUCD_EXCHANGE=units|dp overrides it.  `hp` == nullptr or !uniform (a remainder minibatch in the range, injected
// draws, statistics requested): data-parallel.
static void choose_exchange(kucd_rbm* r, int64_t rows_per_rank, const kucd_hparams* hp = nullptr, bool uniform = false) {
  static const int64_t min_rows_env = [] {
    const char* e = getenv("KUCD_FUSED_MIN_ROWS");
    return e != nullptr ? static_cast<int64_t>(atoll(e)) : static_cast<int64_t>(-1);
  }();
  const int units_env = [] {  // -1 auto, 0 data-parallel, 1 unit-sharded wherever it can run (read per training call)
    const char* e = getenv("KUCD_EXCHANGE");
    if (e == nullptr) return -1;
    return strcmp(e, "units") == 0 ? 1 : (strcmp(e, "dp") == 0 ? 0 : -1);
  }();
  const int64_t min_rows = min_rows_env >= 0 ? min_rows_env : (r->wire16 ? 1024 : 2048);
  r->fused_now = r->peer_on && rows_per_rank >= min_rows;
  // Shorter shards of a SMALL weight matrix (one minibatch strong-scaled over the ranks): the pushes are few bytes even
  // though the contraction is short, and an all-reduce of the same matrix is all latency and exposure.  Measured at C3
  // (64 MiB of fp32 dW) with 512 rows per rank on 8 GPUs, same pod: 0.833 ms per step fused (dW with its pushes 0.109 ms)
  // against 1.002 ms with ncclAllReduce (profiles/r02_call12_n8_c3_strong_fused.json, r02_call11_n8_c3_strong.json).
  // Only where the shard's projections run as a mid chain (enqueue_cd: > 256 rows, >= 32 tiles of 128 x 256 per stage):
  // smaller problems keep NCCL, because their whole step is one small-tile chain launch with dW as its last stage, which
  // cannot push rows to peers.
  const int64_t mid_tiles = ((rows_per_rank + kBlockM - 1) / kBlockM) * ((std::min(r->V, r->H) + 255) / 256);
  if (!r->fused_now && r->peer_on && min_rows_env < 0 && rows_per_rank > 256 && mid_tiles >= 32 &&
      static_cast<int64_t>(r->V) * r->ldH * 4 <= (int64_t{128} << 20))
    r->fused_now = true;
  r->units_now = false;
  const int n = r->ctx->world;
  const int64_t Bg = rows_per_rank * n;
  if (r->units_ok && hp != nullptr && uniform && units_env != 0 && rows_per_rank % 128 == 0 &&
      Bg * std::max(r->V, r->H) / 8 <= kBitSlotBytes) {
    const int k = hp->k, pcd = hp->persistent ? 1 : 0;
    // data-parallel: (4 + 2) V H (n-1)/n bytes over NVLink at the ~350 GB/s the epilogue stores / NCCL reached at C4
    const double t_dp = 6.0 * r->V * r->H * (n - 1) / n / 350e9;
    // unit-sharded: every exchanged state matrix is expanded to bf16 once (HBM writes), plus ~20 us per exchange
    const double t_un = 2.0 * Bg * (static_cast<double>(r->V) * (k + 1) + static_cast<double>(r->H) * (k + pcd)) / 4e12 +
                        (2 * k + 2 + pcd) * 20e-6;
    if (units_env == 1 || t_un < t_dp) {
      r->units_now = true;
      r->fused_now = false;
    }
  }
}

// what the chosen exchange needs allocated before a step is enqueued (or captured)
//
// KUCD_AR_SLABS = S > 1 (NCCL path; off by default until measured): the dW contraction runs as S launches over row slabs
// of W, each slab's ncclAllReduce is enqueued on a second stream as soon as its contraction is, and the update kernel
// follows slab by slab - contraction, all-reduce and update overlap instead of running back to back (at C4 the three
// are 0.45 + 1.24 + 0.3 ms of a 2.95 ms step).  The contraction leaves KUCD_AR_RESERVE_SMS SMs (default 16) to NCCL.
static int prepare_exchange(kucd_rbm* r, int64_t rows_per_rank, const kucd_hparams* hp = nullptr, bool uniform = false) {
  kucd_ctx* ctx = r->ctx;
  choose_exchange(r, rows_per_rank, hp, uniform);
  if (nccl16(r)) KU_TRY(r->grad16.ensure(static_cast<size_t>(r->V) * r->ldH * 2, true));
  static const int slabs_env = [] {
    const char* e = getenv("KUCD_AR_SLABS");
    return e != nullptr ? atoi(e) : 1;
  }();
  r->slabs_now = 1;
  // (On a single rank the same slabs were tried as a way to hide the HBM-bound update behind the contraction of the next
  // slab: measured -1.3 % at C3 with two slabs, +3 ... +7 % at C4's share with 2 ... 8 - profiles/r02_call1_switches.log - and
  // removed.)
  static const int64_t slab_min_elems = [] {  // tests lower it to exercise the path at small sizes
    const char* e = getenv("KUCD_AR_SLABS_MIN_ELEMS");
    return e != nullptr ? static_cast<int64_t>(atoll(e)) : (int64_t{1} << 22);
  }();
  if (slabs_env > 1 && ctx->comm != nullptr && !r->fused_now && !r->units_now && r->compute == KUCD_COMPUTE_BF16 &&
      r->V * r->ldH >= slab_min_elems) {
    const int want = std::min(slabs_env, kucd_ctx::kMaxSlabs);
    const int64_t rows = round_up((r->V + want - 1) / want, 256);
    const int n = static_cast<int>((r->V + rows - 1) / rows);
    if (n > 1) {
      if (ctx->comm_stream == nullptr) CU_TRY(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
      for (int i = 0; i < n; ++i) {
        if (ctx->ev_slab[i] == nullptr) CU_TRY(cudaEventCreateWithFlags(&ctx->ev_slab[i], cudaEventDisableTiming));
        if (ctx->ev_red[i] == nullptr) CU_TRY(cudaEventCreateWithFlags(&ctx->ev_red[i], cudaEventDisableTiming));
      }
      r->slabs_now = n;
      r->slab_rows = rows;
    }
  }
  return KUCD_OK;
}

// fp32 master rows updated by their owners -> every rank (NCCL broadcasts, at API boundaries only)
static int refresh_planes(kucd_rbm* r) {
  kucd_ctx* ctx = r->ctx;
  const int64_t n4 = r->V * r->ldH / 4;
  refresh_planes_kernel<<<grid_for(ctx, n4, 256), 256, 0, ctx->stream>>>(
      r->W32.as<float>(), r->Wp.buf[0].as<__nv_bfloat16>(), r->wparts == 3 ? r->Wp.buf[1].as<__nv_bfloat16>() : nullptr,
      r->wparts == 3 ? r->Wp.buf[2].as<__nv_bfloat16>() : nullptr, n4);
  ctx->tm.aux_launches++;
  CU_TRY(cudaGetLastError());
  return KUCD_OK;
}


// end of a unit-sharded training call: every rank holds the current values of the columns of W (and the slices of b, c)
// it owns; bring the fp32 masters and the bf16 operand plane of every rank up to date (API boundary only)
static int gather_units(kucd_rbm* r) {
  kucd_ctx* ctx = r->ctx;
  const int n = ctx->world, me = ctx->rank;
  const int64_t Hs = r->H / n, Vs = r->V / n;
  const size_t slice = static_cast<size_t>(r->V) * Hs;  // floats
  KU_TRY(r->gtmp.ensure(slice * 4 * (n + 1)));
  float* mine = r->gtmp.as<float>();
  float* all = mine + slice;
  CU_TRY(cudaMemcpy2DAsync(mine, Hs * 4, r->W32.as<float>() + me * Hs, r->ldH * 4, Hs * 4, r->V, cudaMemcpyDeviceToDevice,
                           ctx->stream));
  int rc = g_nccl.GroupStart();
  if (rc == 0) rc = g_nccl.AllGather(mine, all, slice, /*ncclFloat32*/ 7, ctx->comm, ctx->stream);
  if (rc == 0)
    rc = g_nccl.AllGather(r->b32.as<float>() + me * Vs, r->b32.as<float>(), static_cast<size_t>(Vs), 7, ctx->comm, ctx->stream);
  if (rc == 0)
    rc = g_nccl.AllGather(r->c32.as<float>() + me * Hs, r->c32.as<float>(), static_cast<size_t>(Hs), 7, ctx->comm, ctx->stream);
  const int rc2 = g_nccl.GroupEnd();
  if (rc != 0 || rc2 != 0) return fail(KUCD_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(rc != 0 ? rc : rc2));
  for (int j = 0; j < n; ++j) {
    if (j == me) continue;
    CU_TRY(cudaMemcpy2DAsync(r->W32.as<float>() + j * Hs, r->ldH * 4, all + j * slice, Hs * 4, Hs * 4, r->V,
                             cudaMemcpyDeviceToDevice, ctx->stream));
  }
  KU_TRY(refresh_planes(r));
  if (r->chains_g_valid && r->n_chains > 0) {  // this rank's chains back into the buffer the API reads
    const int64_t b = r->chains_g_rows / n;
    CU_TRY(cudaMemcpyAsync(r->chains.buf[0].p, r->chains_g.buf[0].as<__nv_bfloat16>() + me * b * r->ldV,
                           static_cast<size_t>(std::min(b, r->n_chains)) * r->ldV * 2, cudaMemcpyDeviceToDevice, ctx->stream));
    r->chains_g_valid = false;  // the next unit-sharded call gathers them again (the caller may set new ones)
  }
  // nobody may start overwriting a plane that a slower rank's refresh... each rank refreshes its OWN plane from its own
  // master: no cross-rank hazard; the next training call starts with its own barrier
  return KUCD_OK;
}

static int gather_master(kucd_rbm* r) {
  kucd_ctx* ctx = r->ctx;
  if (r->units_now) return gather_units(r);
  if (!r->fused_now) return KUCD_OK;
  int rc = g_nccl.GroupStart();
  for (int o = 0; o < ctx->world && rc == 0; ++o) {
    const int64_t r0 = o * r->rows_per;
    const int64_t rows = std::max<int64_t>(0, std::min<int64_t>(r->rows_per, r->V - r0));
    if (rows == 0) continue;
    float* slice = r->W32.as<float>() + r0 * r->ldH;
    rc = g_nccl.Broadcast(slice, slice, static_cast<size_t>(rows * r->ldH), /*ncclFloat32*/ 7, o, ctx->comm, ctx->stream);
  }
  const int rc2 = g_nccl.GroupEnd();
  if (rc != 0 || rc2 != 0) return fail(KUCD_ERR_NCCL, "ncclBroadcast: %s", g_nccl.GetErrorString(rc != 0 ? rc : rc2));
  return KUCD_OK;
}

// One launch of chain_kernel<BN, CG, GAUSS>.  Its CTAs (CTA pairs) wait on each other's tiles, so all of them must be
// resident at once: one per SM (pair per TPC) at this shared-memory size, never more than the device can hold.
template <int BN, int CG, bool GAUSS, int CH = 0>
static int launch_chain_kernel(kucd_ctx* ctx, const ChainParams& p, int total, bool prof, size_t* pe0) {
  using Cfg = GemmCfg<BN / CG>;
  const void* kern = chain_kernel_ptr<BN, CG, GAUSS, CH>();
  void* args[] = {const_cast<ChainParams*>(&p)};  // one __grid_constant__ parameter block
  const int dev = current_device_slot();
  // per device: 0 = not prepared, 1 = plain launches, 2 = cooperative launches
  static int state[kMaxDevices] = {};
  static int max_clusters[kMaxDevices] = {};
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if constexpr (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  if (state[dev] == 0) {
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if constexpr (CG == 2) {
      cfg.gridDim = dim3(ctx->num_sms / 2 * 2);
      cfg.numAttrs = na;
      int mc = 0;
      if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) != cudaSuccess || mc <= 0) {
        cudaGetLastError();
        mc = ctx->num_sms / 2;
      }
      max_clusters[dev] = mc;
    }
    // The CTAs of this kernel wait on each other's tiles, so all of them must be resident at once.  A COOPERATIVE launch
    // makes the driver guarantee that (the grid is gang-scheduled: if other work holds SMs the launch waits instead of
    // running a part of the grid that would spin on the rest).  Whether this device / driver accepts the attribute for this
    // kernel - together with the cluster dimension - is found out once, with an empty launch (no tiles: every role falls
    // through) on a side stream, so that a refusal can never hit a stream that is being captured.  KUCD_COOP=0: plain.
    state[dev] = 1;
    const char* e = getenv("KUCD_COOP");
    int coop_ok = 0;
    if (!(e != nullptr && e[0] == '0') &&
        cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, ctx->device) == cudaSuccess && coop_ok) {
      ChainParams empty;
      memset(&empty, 0, sizeof empty);
      attr[na].id = cudaLaunchAttributeCooperative;
      attr[na].val.cooperative = 1;
      cfg.numAttrs = na + 1;
      cfg.gridDim = dim3(CG);
      cfg.stream = ctx->copy_stream;
      void* empty_args[] = {&empty};
      if (cudaLaunchKernelExC(&cfg, kern, empty_args) == cudaSuccess && cudaStreamSynchronize(ctx->copy_stream) == cudaSuccess)
        state[dev] = 2;
      cudaGetLastError();
    }
  }
  int units = std::min(total, ctx->num_sms / CG);
  if constexpr (CG == 2) units = std::min(units, max_clusters[dev]);
  if (state[dev] == 2) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  cfg.numAttrs = na;
  cfg.gridDim = dim3(units * CG);
  cfg.stream = ctx->stream;
  *pe0 = prof ? prof_event(ctx) : 0;
  cudaError_t le = cudaLaunchKernelExC(&cfg, kern, args);
  if (le != cudaSuccess && state[dev] == 2 && !capturing(ctx)) {
    // the cooperative form was refused for this launch (seen under ncu, which cannot replay it): plain launches from here
    // on - all CTAs still fit at once on an otherwise idle device, which is all the kernel needs
    cudaGetLastError();
    state[dev] = 1;
    cfg.numAttrs = na - 1;
    le = cudaLaunchKernelExC(&cfg, kern, args);
  }
  if (le != cudaSuccess)
    return fail(KUCD_ERR_CUDA, "chain kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
  return KUCD_OK;
}

// The 2k+1 (+1 with persistent chains) projections of one minibatch as one persistent kernel (chain.cuh).
// `small`: 128 x 64 tiles on single CTAs (latency-bound sizes) instead of 256 x 256 tiles on CTA pairs.
static int launch_chain(kucd_rbm* r, const Planes& v0, int64_t batch, const kucd_hparams* hp, int64_t global_row0,
                        uint64_t draw0, uint64_t stride, const StepDyn* dyn, bool v0_dyn, bool with_dw, int variant) {
  // variant 0: CTA pairs on 256 x 256 tiles; 1: single CTAs on 128 x 64 tiles (latency-bound sizes); 2: single CTAs on
  // 128 x 256 tiles (minibatches too short to give every CTA pair a 256-row tile, e.g. a 512-row shard of 4096 -> 4096)
  // variant 3: single CTAs on 128 x 128 tiles (twice the tiles of variant 2 for the same stage)
  const bool small = variant == 1;
  const int bn = small ? 64 : (variant == 3 ? 128 : kChainBN), cg = variant == 0 ? 2 : 1;
  const int tile_m = kBlockM * cg;
  kucd_ctx* ctx = r->ctx;
  const int k = hp->k;
  const bool pcd = hp->persistent != 0;
  ChainParams p;
  memset(&p, 0, sizeof p);
  std::string err;
  const Planes h0 = r->h0.view(batch, r->H, 1), hk = r->hk.view(batch, r->H, 1), vk = r->vk.view(batch, r->V, 1);
  auto amap = [&](int i, const Planes& q) {
    return make_tmap_bf16(&p.maps[i], MatView{q.p[0], q.rows, q.cols, q.ld}, kBlockM, &err);
  };
  bool ok = amap(0, v0) && amap(1, h0) && amap(2, hk) && amap(3, vk);
  if (ok && pcd) ok = amap(4, r->chains.view(batch, r->V, 1));
  if (ok && !pcd) p.maps[4] = p.maps[0];
  const MatView Wv{r->Wp.buf[0].p, r->V, r->H, r->ldH};
  ok = ok && make_tmap_bf16(&p.maps[5], Wv, 64u, &err)                 // v.W   : W as (K,N), boxes {64 n, 64 k}
          && make_tmap_bf16(&p.maps[6], Wv, bn / cg, &err);             // h.W^T : W as (N,K), boxes {64 k, bn/cg n}
  if (!ok) return fail(KUCD_ERR_CUDA, "%s", err.c_str());
  p.maps[7] = p.maps[6];
  const bool f32 = r->compute == KUCD_COMPUTE_F32X3;  // three term planes of W: W = hi + mid + lo (float32-grade mode)
  if (f32) {
    for (int t = 1; t <= 2 && ok; ++t) {
      const MatView Wt{r->Wp.buf[t].p, r->V, r->H, r->ldH};
      ok = make_tmap_bf16(&p.maps[11 + t], Wt, 64u, &err) && make_tmap_bf16(&p.maps[13 + t], Wt, bn / cg, &err);
    }
    if (!ok) return fail(KUCD_ERR_CUDA, "%s", err.c_str());
  } else {
    for (int i = 12; i < kChainMaps; ++i) p.maps[i] = p.maps[5];
  }
  // the same matrices as MN-major operands of the dW contraction (boxes {64 units, 64 minibatch rows})
  auto mnmap = [&](int i, const Planes& q) {
    return make_tmap_bf16(&p.maps[i], MatView{q.p[0], q.rows, q.cols, q.ld}, 64u, &err);
  };
  if (!(mnmap(8, v0) && mnmap(9, h0) && mnmap(10, vk) && mnmap(11, hk))) return fail(KUCD_ERR_CUDA, "%s", err.c_str());
  const int num_m_batch = static_cast<int>((batch + tile_m - 1) / tile_m);

  auto kind = [&](int i, bool fwd, int map_a, __nv_bfloat16* out, int epi, float* colsum, float sign, bool a_dyn) {
    ChainKind& q = p.kinds[i];
    q.bias = fwd ? r->c32.as<float>() : r->b32.as<float>();
    q.out_bf16 = out;
    q.ld_bf16 = fwd ? r->ldH : r->ldV;
    q.colsum = colsum;
    q.colsum_sign = sign;
    q.colsum_rows = static_cast<int32_t>(batch);
    q.epi = epi;
    q.M = static_cast<int32_t>(batch);
    q.N = static_cast<int32_t>(fwd ? r->H : r->V);
    q.seed = r->seed;
    q.kblocks = static_cast<int32_t>(((fwd ? r->V : r->H) + kBlockK - 1) / kBlockK);
    q.b_mn = fwd ? 1 : 0;
    q.map_a = map_a;
    q.map_b = fwd ? 5 : 6;
    q.a_dyn = a_dyn ? 1 : 0;
    q.num_n = (q.N + bn - 1) / bn;
    q.num_m = num_m_batch;
    q.batch_rows = 1;
    q.nseg = 1;
    if (f32) {  // smallest term of W first (term_pairs): lo, mid, hi
      q.nseg = 3;
      q.map_b = fwd ? 13 : 15;
      q.map_b2 = fwd ? 12 : 14;
      q.map_b3 = fwd ? 5 : 6;
    }
  };
  // Gaussian-visible mode (rbm.py:139-145): relu-threshold hiddens, v = mean + N(0,1), final h still the sigmoid
  const bool gaussian = r->mode == KUCD_MODE_VISIBLE_GAUSSIAN;
  const int eh = gaussian ? kEpiReluSample : kEpiSample, ev = gaussian ? kEpiGaussian : kEpiSample;
  kind(0, true, 0, h0.p[0], eh, r->dc(), 1.f, v0_dyn);                  // h_pos from v0            rbm.py:120
  kind(1, true, 4, hk.p[0], eh, nullptr, 0.f, false);                    // first h of a stored chain (PCD)
  kind(2, false, 1, vk.p[0], ev, nullptr, 0.f, false);                   // v from h_pos             rbm.py:121-123
  kind(3, false, 2, vk.p[0], ev, nullptr, 0.f, false);                   // v from a later h
  kind(4, false, 1, vk.p[0], ev, r->db(), -1.f, false);                  // last v from h_pos (k = 1)
  kind(5, false, 2, vk.p[0], ev, r->db(), -1.f, false);                  // last v from a later h
  kind(6, true, 3, hk.p[0], eh, nullptr, 0.f, false);                    // intermediate h (CD-k)
  kind(7, true, 3, hk.p[0], kEpiProb, r->dc(), -1.f, false);             // final h: probability     rbm.py:124
  if (f32) {  // the probability leaves as three bf16 terms: the dW contraction multiplies it at float32 grade
    p.kinds[7].out_mid = r->hk.buf[1].as<__nv_bfloat16>();
    p.kinds[7].out_lo = r->hk.buf[2].as<__nv_bfloat16>();
  }

  int ns = 0;
  auto stage = [&](int kd, int dep, uint32_t phase) {
    p.stages[ns].kind = static_cast<int16_t>(kd);
    p.stages[ns].dep = static_cast<int16_t>(dep);
    p.stages[ns].phase = phase;
    return ns++;
  };
  int hsrc = stage(0, -1, 0);
  bool from_h0 = true;
  if (pcd) {
    hsrc = stage(1, -1, 1);
    from_h0 = false;
  }
  for (int t = 1; t <= k; ++t) {
    const bool last = t == k;
    const int vs = stage(from_h0 ? (last ? 4 : 2) : (last ? 5 : 3), hsrc, 2u * t);
    hsrc = stage(last ? 7 : 6, vs, 2u * t + 1);
    from_h0 = false;
  }
  if (with_dw) {
    // dW = v0^T h0 - vk^T hk (rbm.py:125-126) as the last stage: its first K-segment only needs stage 0, so the
    // pairs that run out of projection tiles start it while the last projection is still finishing elsewhere
    ChainKind& q = p.kinds[8];
    q.out_f32 = r->dW();
    q.ld_f32 = r->ldH;
    q.epi = kEpiRaw;
    q.M = static_cast<int32_t>(r->V);
    q.N = static_cast<int32_t>(r->H);
    q.seed = r->seed;
    q.kblocks = static_cast<int32_t>((batch + kBlockK - 1) / kBlockK);
    q.a_mn = 1;
    q.b_mn = 1;
    q.nseg = 2;
    q.map_a = 8;
    q.map_b = 9;
    q.map_a2 = 10;
    q.map_b2 = 11;
    q.a_dyn = v0_dyn ? 1 : 0;
    q.num_n = static_cast<int32_t>((r->H + bn - 1) / bn);
    q.num_m = static_cast<int32_t>((r->V + tile_m - 1) / tile_m);
    q.batch_rows = 0;
    q.dep2 = hsrc;  // the final h stage (which itself waited for the final v stage, row block by row block)
    stage(8, 0, 0);
  }
  const int num_m = std::max(num_m_batch, with_dw ? p.kinds[8].num_m : 0);  // stride of the counter array
  int total = 0;
  for (int i = 0; i < ns; ++i) total += p.kinds[p.stages[i].kind].num_m * p.kinds[p.stages[i].kind].num_n;
  p.num_stages = ns;
  p.M = static_cast<int32_t>(batch);
  p.m_valid = static_cast<int32_t>(batch);
  p.total_tiles = total;
  KU_TRY(r->chain_done.ensure(static_cast<size_t>(kMaxChainStages) * num_m * 4));
  p.done = r->chain_done.as<uint32_t>();
  p.done_stride = num_m;
  p.num_m_batch = num_m_batch;
  p.draw = draw0;
  p.draw_stride = stride;
  p.row0 = global_row0;
  p.dyn = dyn;
  p.dyn_rank = ctx->rank;
  CU_TRY(cudaMemsetAsync(p.done, 0, static_cast<size_t>(ns) * num_m * 4, ctx->stream));

  const bool prof = ctx->profile && dyn == nullptr;
  size_t pe0 = 0;
  int rc;
  if (variant == 2)
    rc = launch_chain_kernel<kChainBN, 1, false>(ctx, p, total, prof, &pe0);
  else if (variant == 3)
    rc = launch_chain_kernel<128, 1, false>(ctx, p, total, prof, &pe0);
  else if (f32)
    rc = small ? launch_chain_kernel<64, 1, false, kPreciseCH>(ctx, p, total, prof, &pe0)
               : launch_chain_kernel<kChainBN, 2, false, kPreciseCH>(ctx, p, total, prof, &pe0);
  else if (small)
    rc = gaussian ? launch_chain_kernel<64, 1, true>(ctx, p, total, prof, &pe0)
                  : launch_chain_kernel<64, 1, false>(ctx, p, total, prof, &pe0);
  else
    rc = gaussian ? launch_chain_kernel<kChainBN, 2, true>(ctx, p, total, prof, &pe0)
                  : launch_chain_kernel<kChainBN, 2, false>(ctx, p, total, prof, &pe0);
  KU_TRY(rc);
  if (prof) ctx->marks.push_back({0, pe0, prof_event(ctx), ns});
  ctx->tm.gemm_launches++;
  ctx->tm.chain_launches++;
  if (with_dw) ctx->tm.chain_dw_launches++;
  return KUCD_OK;
}

// One exchange of the unit-sharded step: the rectangle (rows x [col_lo, col_lo + cols)) of `src` goes as bits into every
// rank's slot, the ranks meet, and the whole gathered bit matrix is expanded into the bf16 plane G (all rows, all units).
// in_place: the rectangle was computed into G itself (a slice of the units, all rows) - it is not expanded again.
static int units_exchange(kucd_rbm* r, int& ex, const __nv_bfloat16* src, int64_t src_ld, const StepDyn* src_dyn,
                          int64_t rows, int64_t dst_row0, int64_t col_lo, int64_t cols, const Planes& G,
                          bool in_place = false, cudaStream_t st = nullptr) {
  kucd_ctx* ctx = r->ctx;
  if (st == nullptr) st = ctx->stream;
  const int slot = ex++ & 1;
  const int64_t pitch = G.cols / 8;
  const int64_t total = rows * cols / 8;
  const bool prof = ctx->profile && !capturing(ctx);
  const size_t pe0 = prof ? prof_event(ctx, st) : 0;
  pack_push_kernel<<<grid_for(ctx, total, 256), 256, 0, st>>>(src, src_ld, src_dyn, rows, col_lo, cols, r->ps, ctx->world,
                                                             slot, dst_row0, pitch);
  peer_barrier_kernel<<<1, 32, 0, st>>>(r->epoch, r->ps, ctx->rank, ctx->world);
  ingest_bits_kernel<<<grid_for(ctx, G.rows * (G.ld / 8), 256), 256, 0, st>>>(
      r->ps.bits[ctx->rank] + static_cast<int64_t>(slot) * kBitSlotBytes, pitch, G.rows, G.cols, G.p[0], nullptr, nullptr,
      G.ld, 1, in_place ? col_lo : 0, in_place ? col_lo + cols : 0);
  ctx->tm.aux_launches += 3;
  ctx->tm.unit_exchanges++;
  if (prof) ctx->marks.push_back({2, pe0, prof_event(ctx, st), 1});
  CU_TRY(cudaGetLastError());
  return KUCD_OK;
}

// The CD step of rbm.py:119-126 sharded by UNITS instead of rows (see choose_exchange).  v0_local: this rank's b rows
// of the global minibatch (rank-contiguous shards: global row = rank * b + i).  Every projection below covers all n*b
// rows but only this rank's slice of the output units; Philox draws are keyed by (global row, absolute unit), so the
// sampled states equal a single-GPU run's bit for bit.
static int enqueue_cd_units(kucd_rbm* r, const Planes& v0_local, int64_t b, const kucd_hparams* hp, int64_t global_row0,
                            uint64_t step, const StepDyn* dyn, bool v0_dyn) {
  kucd_ctx* ctx = r->ctx;
  const int n = ctx->world, me = ctx->rank, k = hp->k;
  const bool pcd = hp->persistent != 0;
  const int64_t Bg = b * n, Hs = r->H / n, Vs = r->V / n, h_lo = me * Hs, v_lo = me * Vs;
  const int64_t base_row0 = global_row0 - me * b;
  const uint64_t draw0 = step * 64, stride = dyn != nullptr ? 64 : 0;
  const Planes Gv0 = r->vin.view(Bg, r->V, 1), Gh0 = r->h0.view(Bg, r->H, 1), Ghk = r->hk.view(Bg, r->H, 1);
  // with persistent chains the negative phase's visible states ARE the chains: the chain's first projection reads them,
  // the last back-projection overwrites them - no copy
  const Planes Gvk = pcd ? r->chains_g.view(Bg, r->V, 1) : r->vk.view(Bg, r->V, 1);
  int ex = 0;
  auto stage = [&](bool forward, const Planes& a, const Planes& out, int epi, int phase, float* colsum, float sign,
                   cudaStream_t st = nullptr) -> int {
    EpiArgs e;
    e.epi = epi;
    e.out = out;
    e.colsum = colsum;
    e.colsum_sign = sign;
    e.draw = draw0 + phase;
    e.draw_stride = stride;
    e.row0 = base_row0;
    e.dyn = dyn;
    e.static_rows = true;
    e.n_lo = forward ? h_lo : v_lo;
    e.n_cnt = forward ? Hs : Vs;
    e.stream = st;
    return project(r, forward, a, Bg, e);
  };
  auto share_h = [&](const Planes& G) { return units_exchange(r, ex, G.p[0], G.ld, nullptr, Bg, 0, h_lo, Hs, G, true); };
  auto share_v = [&](const Planes& G) { return units_exchange(r, ex, G.p[0], G.ld, nullptr, Bg, 0, v_lo, Vs, G, true); };
  // With persistent chains the positive phase (gather the minibatch, project it) and the negative chain (project the
  // stored chains, exchange, project back, exchange, project) only meet in dW.  The negative chain runs on the main
  // stream; the gather of the minibatch goes to the second stream and the positive-phase projection to the third: its
  // CTAs take the SMs the chain's contractions leave idle - their last, partial wave (128 tiles on 74 CTA pairs) and the
  // gaps in which the chain waits for an exchange.  KUCD_UNITS_OVERLAP=0: everything in order on one stream.
  static const bool overlap_env = [] {
    const char* e = getenv("KUCD_UNITS_OVERLAP");
    return !(e != nullptr && e[0] == '0');
  }();
  const bool overlap = overlap_env && pcd;
  cudaStream_t s2 = overlap ? ctx->stream2 : ctx->stream, s3 = overlap ? ctx->stream3 : ctx->stream;
  CU_TRY(cudaMemsetAsync(r->db(), 0, (r->ldVb() + r->ldHb()) * 4, ctx->stream));
  if (overlap) {
    CU_TRY(cudaEventRecord(ctx->ev_units[0], ctx->stream));
    CU_TRY(cudaStreamWaitEvent(s2, ctx->ev_units[0], 0));
  }
  // the global minibatch on every rank
  KU_TRY(units_exchange(r, ex, v0_local.p[0], v0_local.ld, v0_dyn ? dyn : nullptr, b, me * b, 0, r->V, Gv0, false, s2));
  {  // db[own visibles] += sum_rows v0   (rbm.py:134)
    dim3 grid(static_cast<unsigned>((Vs / 2 + 1 + 127) / 128), static_cast<unsigned>((Bg + 63) / 64));
    colsum_kernel<<<grid, 128, 0, s2>>>(Gv0.p[0] + v_lo, nullptr, nullptr, Gv0.ld, 0, static_cast<int32_t>(Bg),
                                        static_cast<int32_t>(Vs), nullptr, 1.f, r->db() + v_lo);
    ctx->tm.aux_launches++;
    CU_TRY(cudaGetLastError());
  }
  if (overlap) {
    CU_TRY(cudaEventRecord(ctx->ev_units[1], s2));
    CU_TRY(cudaStreamWaitEvent(s3, ctx->ev_units[1], 0));  // (also forks the third stream off the capture)
  }
  // h_pos   rbm.py:120   (with persistent chains nobody back-projects it: only this rank's slice is ever read, by dW and dc)
  KU_TRY(stage(true, Gv0, Gh0, kEpiSample, 0, r->dc(), 1.f, s3));
  if (overlap) CU_TRY(cudaEventRecord(ctx->ev_units[2], s3));
  if (!pcd) KU_TRY(share_h(Gh0));
  Planes hcur = Gh0;
  if (pcd) {  // the negative chain starts at the stored fantasy particles
    KU_TRY(stage(true, Gvk, Ghk, kEpiSample, 1, nullptr, 0.f));
    // every rank must run its flag barriers in the same order (they share one epoch counter): the gather's barrier on
    // the second stream comes before this exchange's - it finished long ago, the projection above is several times longer
    if (overlap) CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_units[1], 0));
    KU_TRY(share_h(Ghk));
    hcur = Ghk;
  }
  for (int t = 1; t <= k; ++t) {
    const bool last = t == k;
    KU_TRY(stage(false, hcur, Gvk, kEpiSample, 2 * t, last ? r->db() : nullptr, -1.f));  // rbm.py:121-123
    KU_TRY(share_v(Gvk));
    KU_TRY(stage(true, Gvk, Ghk, last ? kEpiProb : kEpiSample, 2 * t + 1, last ? r->dc() : nullptr, -1.f));  // :124
    if (!last) KU_TRY(share_h(Ghk));
    hcur = Ghk;
  }
  if (overlap) CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_units[2], 0));  // the positive phase joins
  // dW[:, own hidden units] = v0^T h0 - vk^T hk over the whole global minibatch   (rbm.py:125-126)
  KU_TRY(delta_w(r, Gv0, Gh0, Gvk, Ghk, Bg, nullptr, false, 0, -1, 0, h_lo, Hs));
  r->last_rows = Bg;
  r->last_vk_parts = 1;
  r->last_hk_parts = 1;
  r->units_pcd = pcd;
  ctx->tm.unit_steps++;
  return KUCD_OK;
}

// v0: the minibatch operand (rows [0,batch) of it, or - with v0_dyn - the rows at dyn->row_off of a data set)
static int enqueue_cd(kucd_rbm* r, const Planes& v0, int64_t batch, const kucd_hparams* hp, const StepInject* inj,
                      int64_t global_row0, uint64_t step, const StepDyn* dyn, bool v0_dyn) {
  kucd_ctx* ctx = r->ctx;
  const int k = hp->k;
  const bool gaussian = r->mode == KUCD_MODE_VISIBLE_GAUSSIAN;
  const int epi_h = gaussian ? kEpiReluSample : kEpiSample;
  const int epi_v = gaussian ? kEpiGaussian : kEpiSample;
  const uint64_t draw0 = step * 64;
  const uint64_t stride = dyn != nullptr ? 64 : 0;
  if (r->units_now) return enqueue_cd_units(r, v0, batch, hp, global_row0, step, dyn, v0_dyn);

  static const bool merge_small = [] {
    const char* e = getenv("KUCD_MERGE");
    return !(e != nullptr && e[0] == '0');
  }();
  if (batch <= 128 && merge_small) {
    // smallest minibatches: one launch stores db = sum_rows v0 (rbm.py:134) and clears dc.  Measured (same call,
    // KUCD_MERGE=1/0): 43.1 vs 51.2 us per step at C1 (batch 128) but 286 vs 171 us per 3-layer step at C2 (batch
    // 256), hence the threshold.
    const int32_t cols_pad = static_cast<int32_t>(round_up(r->V, 2));
    const int blocks = std::max<int>(1, (cols_pad / 2 + 31) / 32);
    colsum_store_kernel<<<blocks, dim3(32, 16), 0, ctx->stream>>>(v0.p[0], v0.mid(), v0.lo(), v0.ld, static_cast<int32_t>(batch),
                                                           static_cast<int32_t>(r->V), cols_pad, v0_dyn ? dyn : nullptr,
                                                           r->db(), r->dc(), static_cast<int32_t>(r->ldHb()));
    ctx->tm.aux_launches++;
    CU_TRY(cudaGetLastError());
  } else {
    CU_TRY(cudaMemsetAsync(r->db(), 0, (r->ldVb() + r->ldHb()) * 4, ctx->stream));
    // db += sum_rows v0   (rbm.py:134)
    dim3 grid(static_cast<unsigned>((r->V / 2 + 1 + 127) / 128), static_cast<unsigned>((batch + 63) / 64));
    colsum_kernel<<<grid, 128, 0, ctx->stream>>>(v0.p[0], v0.mid(), v0.lo(), v0.ld, 0, static_cast<int32_t>(batch),
                                                 static_cast<int32_t>(r->V), v0_dyn ? dyn : nullptr, 1.f, r->db());
    ctx->tm.aux_launches++;
    CU_TRY(cudaGetLastError());
  }
  // rows of the state buffers beyond the valid ones are written as zeros by the epilogues, so the
  // remainder minibatch contributes nothing to the batch-contracted dW
  const Planes h0 = r->h0.view(batch, r->H, 1);
  const int vparts = vis_parts_out(r);
  const int pparts = prob_parts_out(r);
  const Planes vk = r->vk.view(batch, r->V, vparts);
  const Planes hk = r->hk.view(batch, r->H, pparts);

  auto rows_of = [](const Planes& p, int64_t r0, int64_t n, int parts) {
    Planes q = p;
    for (int i = 0; i < 3; ++i)
      if (q.p[i] != nullptr) q.p[i] += r0 * q.ld;
    q.rows = n;
    q.n = parts;
    return q;
  };
  auto u_at = [](const float* u, int64_t ld, int64_t r0) { return u != nullptr ? u + r0 * ld : nullptr; };

  // One Gibbs chain over minibatch rows [r0, r0 + n): rows are independent given W, b, c (rbm.py:119-124).
  auto chain = [&](int64_t r0, int64_t n, cudaStream_t st, bool timed) -> int {
    const int64_t g0 = global_row0 + r0;
    const int32_t base = static_cast<int32_t>(r0);
    // v0 of a data set stays whole (the kernel offsets its TMA coordinate); a staged minibatch is sliced
    const Planes v0s = v0_dyn ? v0 : rows_of(v0, r0, n, v0.n);
    const Planes h0s = rows_of(h0, r0, n, 1);
    auto common = [&](EpiArgs& e, int phase) {
      e.draw = draw0 + phase;
      e.draw_stride = stride;
      e.row0 = g0;
      e.dyn = dyn;
      e.row_base = base;
      e.stream = st;
      e.no_prof = !timed;
      e.colsum_rows = batch;  // both chains of a split minibatch add into the same statistics: one grid
    };
    {
      EpiArgs e;
      e.epi = epi_h;
      e.out = h0s;
      e.u = inj ? u_at(inj->u_h[0], inj->ld_h[0], r0) : nullptr;
      e.ld_u = inj ? inj->ld_h[0] : 0;
      e.colsum = r->dc();
      e.colsum_sign = 1.f;
      common(e, 0);
      e.a_dyn = v0_dyn;
      KU_TRY(project(r, true, v0s, n, e));
    }
    Planes hcur = h0s;
    if (hp->persistent) {
      // negative chain starts at the stored fantasy particles
      const Planes ch = rows_of(r->chains.view(batch, r->V, r->last_vk_parts), r0, n, r->last_vk_parts);
      EpiArgs e;
      e.epi = epi_h;
      e.out = rows_of(hk, r0, n, 1);
      e.u = inj ? u_at(inj->u_hc, inj->ld_hc, r0) : nullptr;
      e.ld_u = inj ? inj->ld_hc : 0;
      common(e, 1);
      KU_TRY(project(r, true, ch, n, e));
      hcur = e.out;
    }
    const Planes vks = rows_of(vk, r0, n, vparts);
    for (int t = 1; t <= k; ++t) {
      {
        EpiArgs e;
        e.epi = epi_v;
        e.out = vks;
        e.u = inj ? u_at(inj->u_v[t], inj->ld_v[t], r0) : nullptr;
        e.ld_u = inj ? inj->ld_v[t] : 0;
        if (t == k) {
          e.colsum = r->db();
          e.colsum_sign = -1.f;
        }
        common(e, 2 * t);
        KU_TRY(project(r, false, hcur, n, e));
      }
      {
        EpiArgs e;
        const bool last = t == k;
        e.epi = last ? kEpiProb : epi_h;  // rbm.py:124: the final hidden term is the probability
        e.out = rows_of(hk, r0, n, last ? pparts : 1);
        e.u = (!last && inj) ? u_at(inj->u_h[t], inj->ld_h[t], r0) : nullptr;
        e.ld_u = (!last && inj) ? inj->ld_h[t] : 0;
        if (last) {
          e.colsum = r->dc();
          e.colsum_sign = -1.f;
        }
        common(e, 2 * t + 1);
        KU_TRY(project(r, true, vks, n, e));
        hcur = e.out;
      }
    }
    return KUCD_OK;
  };

  // Two chains on two streams when each half still fills the GPU: a 128 x 256-tiled projection of a
  // 4096-row minibatch is 512 tiles = 3.46 waves of 148 CTAs, and a lone kernel idles 13 % of the SMs in
  // its last wave.  Split by rows, the tail of one chain's kernel is filled by the other chain's next
  // kernel (their CTAs become resident as SMs drain), so the tensor pipes stay busy across launches.
  const int64_t half = round_up((batch + 1) / 2, kBlockM);
  const int64_t tiles_half = (half / kBlockM) * ((std::min(r->V, r->H) + 255) / 256);
  const bool two = ctx->split && half < batch && (tiles_half >= ctx->num_sms || ctx->split_force);
  const int n_proj = 2 * k + 1 + (hp->persistent ? 1 : 0);
  // one persistent kernel for the whole chain when every stage fills the chip with 256 x 256 tiles
  const int64_t pair_tiles = ((batch + 255) / 256) * ((std::min(r->V, r->H) + 255) / 256);
  // (float32-grade mode: Bernoulli visibles only - every state operand is then a single exact bf16 plane and a
  // projection is the three term planes of W as three K-segments, accumulated piecewise; chain.cuh, CH > 0)
  const bool f32 = r->compute == KUCD_COMPUTE_F32X3;
  const bool chain_able = ctx->chain && inj == nullptr && v0.n == 1 && (!f32 || !gaussian) &&
                          (!hp->persistent || r->last_vk_parts == 1);
  const bool whole_chain = chain_able && (pair_tiles >= ctx->num_sms / 2 || ctx->chain_force);
  // KUCD_CHAIN_DW=1 appends the dW contraction to the chain kernel as a final two-segment stage.  Measured: no gain
  // at C3 (2.529 vs 2.522 ms per step) and a loss at C4 (577 k vs 605 k samples/s) - its second segment has to wait
  // for every row block of the last projection, so it hides little - which is why it is off by default.
  const bool slabbed = r->slabs_now > 1 && !r->fused_now;
  const bool chain_dw = ctx->chain_dw && !r->fused_now && !nccl16(r) && !slabbed && !f32;
  // latency-bound sizes: the whole step's contractions (projections and dW) as one launch of the small-tile variant
  static const bool small_chain_env = [] {
    const char* e = getenv("KUCD_SMALL_CHAIN");
    return !(e != nullptr && e[0] == '0');
  }();
  // In between - a minibatch (or a data-parallel shard of one) too short for 74 pair tiles per stage but long enough
  // for >= 32 tiles of 128 x 256: the same flattened chain on single CTAs.  A 512-row shard of 4096 -> 4096 (C3 strong
  // scaling over 8 GPUs) is 64 such tiles per stage; launched one projection at a time each of its 21 projections cost
  // 58 us for 17 GFLOP (profiles/README.md, round 2).  The dW contraction stays a launch of its own (it may push rows to peers).
  const int64_t mid_tiles = ((batch + kBlockM - 1) / kBlockM) * ((std::min(r->V, r->H) + 255) / 256);
  static const bool mid_chain_env = [] {
    const char* e = getenv("KUCD_MID_CHAIN");
    return !(e != nullptr && e[0] == '0');
  }();
  const bool mid_chain = !whole_chain && mid_chain_env && chain_able && !f32 && !gaussian && batch > 256 && mid_tiles >= 32;
  const bool small_chain = !whole_chain && !mid_chain && small_chain_env && chain_able && batch <= 512 &&
                           !r->fused_now && !nccl16(r) && !slabbed;
  // the small-tile variant carries the dW contraction as its last stage, except at float32 grade (four term products)
  const bool small_dw = small_chain && !f32;
  if (whole_chain) {
    KU_TRY(launch_chain(r, v0, batch, hp, global_row0, draw0, stride, dyn, v0_dyn, chain_dw, 0));
  } else if (mid_chain) {
    // 128 x 256 tiles, or 128 x 128 when that still leaves SMs idle (KUCD_MID_BN=128|256 forces one)
    static const int mid_bn_env = [] {
      const char* e = getenv("KUCD_MID_BN");
      return e != nullptr ? atoi(e) : 0;
    }();
    const bool narrow = mid_bn_env == 128 || (mid_bn_env != 256 && mid_tiles < ctx->num_sms * 3 / 4);
    KU_TRY(launch_chain(r, v0, batch, hp, global_row0, draw0, stride, dyn, v0_dyn, false, narrow ? 3 : 2));
  } else if (small_chain) {
    KU_TRY(launch_chain(r, v0, batch, hp, global_row0, draw0, stride, dyn, v0_dyn, small_dw, 1));
  } else if (!two) {
    KU_TRY(chain(0, batch, ctx->stream, true));
  } else {
    const bool prof = ctx->profile && dyn == nullptr;
    const size_t pe0 = prof ? prof_event(ctx) : 0;
    CU_TRY(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU_TRY(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    // interleave the enqueue order so both streams always have a kernel queued
    int rc = chain(0, half, ctx->stream, false);
    if (rc == KUCD_OK) rc = chain(half, batch - half, ctx->stream2, false);
    // always re-join: a forked capture that is not joined cannot be ended
    cudaEventRecord(ctx->ev_join, ctx->stream2);
    cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    KU_TRY(rc);
    if (prof) ctx->marks.push_back({0, pe0, prof_event(ctx), 2 * n_proj});
  }
  if (slabbed) {
    // contraction of slab i+1 (compute stream, a few SMs left free) overlaps the all-reduce of slab i (comm stream)
    static const int reserve = [] {
      const char* e = getenv("KUCD_AR_RESERVE_SMS");
      return e != nullptr ? std::max(0, atoi(e)) : 16;
    }();
    const bool d16 = nccl16(r);
    const int S = r->slabs_now;
    for (int i = 0; i < S; ++i) {
      const int64_t w0 = i * r->slab_rows, wn = std::min<int64_t>(r->slab_rows, r->V - w0);
      KU_TRY(delta_w(r, v0, h0, vk, hk, batch, dyn, v0_dyn, w0, wn, reserve));
      CU_TRY(cudaEventRecord(ctx->ev_slab[i], ctx->stream));
      CU_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_slab[i], 0));
      const size_t off = static_cast<size_t>(w0) * r->ldH, cnt = static_cast<size_t>(wn) * r->ldH;
      int rc = i == 0 ? g_nccl.GroupStart() : 0;
      if (rc == 0) {
        if (d16) {
          __nv_bfloat16* g = r->grad16.as<__nv_bfloat16>() + off;
          rc = g_nccl.AllReduce(g, g, cnt, /*ncclBfloat16*/ 9, /*ncclSum*/ 0, ctx->comm, ctx->comm_stream);
        } else {
          float* g = r->dW() + off;
          rc = g_nccl.AllReduce(g, g, cnt, /*ncclFloat32*/ 7, /*ncclSum*/ 0, ctx->comm, ctx->comm_stream);
        }
      }
      if (i == 0) {  // the bias statistics are final since the chain ended: they travel with the first slab
        if (rc == 0)
          rc = g_nccl.AllReduce(r->db(), r->db(), static_cast<size_t>(r->ldVb() + r->ldHb()), /*ncclFloat32*/ 7,
                                /*ncclSum*/ 0, ctx->comm, ctx->comm_stream);
        const int rc2 = g_nccl.GroupEnd();
        if (rc == 0) rc = rc2;
      }
      if (rc != 0) return fail(KUCD_ERR_NCCL, "ncclAllReduce (slab %d): %s", i, g_nccl.GetErrorString(rc));
      CU_TRY(cudaEventRecord(ctx->ev_red[i], ctx->comm_stream));
    }
    r->slabs_inflight = S;
    ctx->tm.allreduce_calls++;
  } else if (!((whole_chain && chain_dw) || small_dw)) {
    KU_TRY(delta_w(r, v0, h0, vk, hk, batch, dyn, v0_dyn));
  }
  r->last_rows = batch;
  r->last_vk_parts = vparts;
  r->last_hk_parts = pparts;
  (void)n_proj;

  if (hp->persistent) {
    for (int i = 0; i < vparts; ++i) {
      copy_rows_kernel<<<grid_for(ctx, batch * (r->ldV / 8), 256), 256, 0, ctx->stream>>>(
          vk.p[i], r->chains.buf[i].as<__nv_bfloat16>(), r->ldV, static_cast<int32_t>(batch), dyn);
      ctx->tm.aux_launches++;
    }
    CU_TRY(cudaGetLastError());
  }
  if (r->fused_now) {
    // The dW contraction has already stored every row into its owner's slot (reduce-scatter fused into the
    // epilogue, NVLink stores overlapped with the MMAs).  Ship the small bias statistics the same way, then meet:
    // after the barrier every slot of this rank holds this step's contributions of all ranks.
    push_bias_kernel<<<(r->bias_len + 255) / 256, 256, 0, ctx->stream>>>(r->db(), r->ps, ctx->rank, ctx->world, r->bias_len);
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(r->epoch, r->ps, ctx->rank, ctx->world);
    ctx->tm.aux_launches += 2;
    ctx->tm.fused_reduce_steps++;
    CU_TRY(cudaGetLastError());
  } else if (ctx->comm != nullptr && !slabbed) {
    int rc;
    if (nccl16(r)) {  // dW as bf16 (half the bytes; NCCL adds in bf16), the small bias statistics as fp32, one group
      rc = g_nccl.GroupStart();
      if (rc == 0)
        rc = g_nccl.AllReduce(r->grad16.p, r->grad16.p, static_cast<size_t>(r->V * r->ldH), /*ncclBfloat16*/ 9,
                              /*ncclSum*/ 0, ctx->comm, ctx->stream);
      if (rc == 0)
        rc = g_nccl.AllReduce(r->db(), r->db(), static_cast<size_t>(r->ldVb() + r->ldHb()), /*ncclFloat32*/ 7,
                              /*ncclSum*/ 0, ctx->comm, ctx->stream);
      const int rc2 = g_nccl.GroupEnd();
      if (rc == 0) rc = rc2;
    } else {
      rc = g_nccl.AllReduce(r->grad.p, r->grad.p, static_cast<size_t>(r->grad_elems()), /*ncclFloat32*/ 7,
                            /*ncclSum*/ 0, ctx->comm, ctx->stream);
    }
    if (rc != 0) return fail(KUCD_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(rc));
    ctx->tm.allreduce_calls++;
  }
  return KUCD_OK;
}

static int free_energy_into(kucd_rbm* r, const Planes& v, int64_t rows, float* sp, float* out, const StepDyn* dyn,
                            bool v_dyn) {
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaMemsetAsync(sp, 0, rows * 4, ctx->stream));
  EpiArgs e;
  e.epi = kEpiFreeEnergy;
  e.rowsum = sp;
  e.dyn = dyn;
  e.a_dyn = v_dyn;
  KU_TRY(project(r, true, v, rows, e));
  const int threads = 256;
  const int blocks = static_cast<int>((rows * 32 + threads - 1) / threads);
  free_energy_finish_kernel<<<blocks, threads, 0, ctx->stream>>>(v.p[0], v.mid(), v.lo(), v.ld, 0,
                                                                 static_cast<int32_t>(rows), static_cast<int32_t>(r->V),
                                                                 v_dyn ? dyn : nullptr, r->b32.as<float>(), sp, out);
  ctx->tm.aux_launches++;
  CU_TRY(cudaGetLastError());
  return KUCD_OK;
}

// rbm.py:225-233: fe = F(v); fresh chain v -> h -> v'; fe_p = F(v'); stats[0] = mean|fe - fe_p|
static int enqueue_score(kucd_rbm* r, const Planes& v0, int64_t rows, const float* u_h, int64_t ld_h, const float* u_v,
                         int64_t ld_v, int64_t row0) {
  kucd_ctx* ctx = r->ctx;
  const bool gaussian = r->mode == KUCD_MODE_VISIBLE_GAUSSIAN;
  const uint64_t draw = (1ull << 62) | (2 * r->score_draws++);
  KU_TRY(free_energy_into(r, v0, rows, r->sp0.as<float>(), r->fe0.as<float>(), nullptr, false));
  const Planes h = r->h0.view(rows, r->H, 1);
  {
    EpiArgs e;
    e.epi = gaussian ? kEpiReluSample : kEpiSample;
    e.out = h;
    e.u = u_h;
    e.ld_u = ld_h;
    e.draw = draw;
    e.row0 = row0;
    KU_TRY(project(r, true, v0, rows, e));
  }
  const Planes vn = r->vk.view(rows, r->V, vis_parts_out(r));
  {
    EpiArgs e;
    e.epi = gaussian ? kEpiGaussian : kEpiSample;
    e.out = vn;
    e.u = u_v;
    e.ld_u = ld_v;
    e.draw = draw + 1;
    e.row0 = row0;
    KU_TRY(project(r, false, h, rows, e));
  }
  KU_TRY(free_energy_into(r, vn, rows, r->sp1.as<float>(), r->fe1.as<float>(), nullptr, false));
  score_kernel<<<1, 1024, 0, ctx->stream>>>(r->fe0.as<float>(), r->fe1.as<float>(), static_cast<int32_t>(rows), nullptr,
                                            r->stats.as<float>());
  ctx->tm.aux_launches++;
  CU_TRY(cudaGetLastError());
  return KUCD_OK;
}

// Before a unit-sharded training call (never inside a capture): workspaces for the GLOBAL minibatch and, with persistent
// chains, every rank's chains gathered on every rank (chains_g).
static int prepare_units(kucd_rbm* r, int64_t b, const kucd_hparams* hp);
static int ensure_chains(kucd_rbm* r, int64_t rows) {
  if (r->n_chains >= rows) return KUCD_OK;
  return fail(KUCD_ERR_INVALID_ARG,
              "persistent CD needs %lld chains but %lld are set (kucd_rbm_set_chains first)", (long long)rows,
              (long long)r->n_chains);
}


// v0_dyn != nullptr: v0 is a resident block of rows and the minibatch starts at v0_dyn->row_off (graph replay)
// host_log != nullptr (streamed fit under graph replay): the statistic also goes into the caller-visible page-locked log,
// slot = index of the minibatch (log_dyn, log_batch: see log_stat_kernel).
static int enqueue_recon(kucd_rbm* r, const Planes& v0, int64_t rows, const StepDyn* v0_dyn = nullptr,
                         float* host_log = nullptr, const StepDyn* log_dyn = nullptr, int32_t log_batch = 0) {
  kucd_ctx* ctx = r->ctx;
  Planes vk = r->vk.view(rows, r->V, r->last_vk_parts);
  if (r->units_now) {  // the last step's v_neg of the GLOBAL minibatch: this rank's rows sit at rank * rows
    vk = (r->chains_g_rows == r->last_rows && r->chains_g.buf[0].p != nullptr && r->units_pcd ? r->chains_g : r->vk)
             .view(rows, r->V, 1);
    vk.p[0] += static_cast<int64_t>(ctx->rank) * (r->last_rows / ctx->world) * vk.ld;
  }
  if (rows * r->V <= (1 << 18) && v0.n == 1 && vk.n == 1) {  // latency-bound: one launch, a few small blocks
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(rows, kReconBlocks)));
    recon_small_kernel<<<blocks, 128, 0, ctx->stream>>>(v0.p[0], v0.ld, vk.p[0], vk.ld, static_cast<int32_t>(rows),
                                                        static_cast<int32_t>(r->V), v0_dyn, r->stats.as<float>(), log_dyn,
                                                        log_batch, host_log, r->stats.as<float>() + 16,
                                                        r->stats.as<unsigned int>() + 15);
    ctx->tm.aux_launches++;
    CU_TRY(cudaGetLastError());
    return KUCD_OK;
  }
  float* acc = r->stats.as<float>() + 8;
  CU_TRY(cudaMemsetAsync(acc, 0, 4, ctx->stream));
  recon_kernel<<<grid_for(ctx, rows * r->V, 256), 256, 0, ctx->stream>>>(
      v0.p[0], v0.mid(), v0.lo(), v0.ld, 0, vk.p[0], vk.mid(), vk.lo(), vk.ld, static_cast<int32_t>(rows),
      static_cast<int32_t>(r->V), v0_dyn, acc);
  recon_finish_kernel<<<1, 1, 0, ctx->stream>>>(acc, static_cast<int32_t>(rows), static_cast<int32_t>(r->V), v0_dyn,
                                                r->stats.as<float>());
  ctx->tm.aux_launches += 2;
  if (host_log != nullptr) {
    log_stat_kernel<<<1, 1, 0, ctx->stream>>>(r->stats.as<float>() + 1, log_dyn, log_batch, host_log);
    ctx->tm.aux_launches++;
  }
  CU_TRY(cudaGetLastError());
  return KUCD_OK;
}

static int prepare_units(kucd_rbm* r, int64_t b, const kucd_hparams* hp) {
  kucd_ctx* ctx = r->ctx;
  const int n = ctx->world, me = ctx->rank;
  const int64_t Bg = b * n;
  KU_TRY(ensure_workspace(r, Bg));
  if (!hp->persistent) return KUCD_OK;
  if (r->chains_g_valid && r->chains_g_rows == Bg) return KUCD_OK;
  KU_TRY(ensure_chains(r, b));
  if (r->last_vk_parts != 1) return fail(KUCD_ERR_INVALID_ARG, "unit-sharded steps need 0/1 chains");
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  const void* before = r->chains_g.buf[0].p;
  KU_TRY(r->chains_g.ensure(round_up(Bg, 128), r->ldV, 1));
  if (before != r->chains_g.buf[0].p && r->graph_exec != nullptr) {  // a captured step holds the old pointer
    cudaGraphExecDestroy(r->graph_exec);
    cudaGraphDestroy(r->graph);
    r->graph_exec = nullptr;
    r->graph = nullptr;
  }
  int ex = 0;
  const Planes G = r->chains_g.view(Bg, r->V, 1);
  KU_TRY(units_exchange(r, ex, r->chains.buf[0].as<__nv_bfloat16>(), r->ldV, nullptr, b, me * b, 0, r->V, G));
  // slot parity restarts at 0 inside the step: close this exchange like a step is closed
  peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(r->epoch, r->ps, me, n);
  ctx->tm.aux_launches++;
  CU_TRY(cudaGetLastError());
  r->chains_g_valid = true;
  r->chains_g_rows = Bg;
  return KUCD_OK;
}

// caller batch -> vin planes; returns the live term count
static int ingest_batch(kucd_rbm* r, const kucd_tensor* v, int64_t r0, int64_t n, Planes* out) {
  kucd_ctx* ctx = r->ctx;
  const bool x3 = r->compute == KUCD_COMPUTE_F32X3;
  Planes dst = r->vin.view(n, r->V, x3 ? 3 : 1);
  int live = 1;
  const bool may_be_inexact = x3 && v->dtype_code == KUCD_DT_FLOAT;
  if (may_be_inexact) CU_TRY(cudaMemsetAsync(r->flag.p, 0, 4, ctx->stream));
  KU_TRY(ingest_rows(ctx, v, r0, n, dst, x3 ? 3 : 1, may_be_inexact ? r->flag.as<int>() : nullptr));
  if (may_be_inexact) {
    int flag = 0;
    CU_TRY(cudaMemcpyAsync(&flag, r->flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    live = flag ? 3 : 1;
  }
  dst.n = live;
  *out = dst;
  return KUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int kucd_abi_version(void) { return KUCD_ABI_VERSION; }
const char* kucd_last_error(void) { return g_err.c_str(); }

// dlpack.h (v0.8 / v1.0 share this prefix): the part of DLManagedTensor this library reads
namespace {
struct DlDevice {
  int32_t device_type, device_id;
};
struct DlDataType {
  uint8_t code, bits;
  uint16_t lanes;
};
struct DlTensor {
  void* data;
  DlDevice device;
  int32_t ndim;
  DlDataType dtype;
  int64_t* shape;
  int64_t* strides;
  uint64_t byte_offset;
};
}  // namespace

int kucd_tensor_from_dlpack(const void* dl_managed_tensor, kucd_tensor* out) {
  if (dl_managed_tensor == nullptr || out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  const DlTensor& t = *static_cast<const DlTensor*>(dl_managed_tensor);  // dl_tensor is the first member
  if (t.ndim < 1 || t.ndim > 2) return fail(KUCD_ERR_INVALID_ARG, "DLPack tensor has %d dimensions, expected 1 or 2", t.ndim);
  if (t.dtype.lanes != 1) return fail(KUCD_ERR_INVALID_ARG, "DLPack tensor has %d vector lanes", (int)t.dtype.lanes);
  int code = t.dtype.code, bits = t.dtype.bits;
  if (code == 6 /* kDLBool */ && bits == 8) code = KUCD_DT_UINT;
  const bool ok = (code == KUCD_DT_FLOAT && bits == 32) || (code == KUCD_DT_BFLOAT && bits == 16) ||
                  (code == KUCD_DT_UINT && bits == 8);
  if (!ok)
    return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "DLPack dtype code %d / %d bits is not float32, bfloat16, uint8 or bool",
                (int)t.dtype.code, bits);
  if (t.device.device_type != KUCD_DEV_CPU && t.device.device_type != KUCD_DEV_CUDA &&
      t.device.device_type != KUCD_DEV_CUDA_HOST)
    return fail(KUCD_ERR_INVALID_ARG, "DLPack device type %d is not CPU, CUDA or CUDA-pinned host", t.device.device_type);
  const int64_t rows = t.shape[0], cols = t.ndim == 2 ? t.shape[1] : 1;
  int64_t s0 = t.ndim == 2 ? cols : 1, s1 = 1;
  if (t.strides != nullptr) {
    s0 = t.strides[0];
    s1 = t.ndim == 2 ? t.strides[1] : 1;
  }
  if (cols > 1 && s1 != 1) return fail(KUCD_ERR_INVALID_ARG, "DLPack tensor: innermost stride %lld, expected 1", (long long)s1);
  if (rows <= 1) s0 = std::max<int64_t>(cols, 1);
  out->data = static_cast<char*>(t.data) + t.byte_offset;
  out->device_type = t.device.device_type;
  out->device_id = t.device.device_id;
  out->dtype_code = code;
  out->bits = bits;
  out->shape[0] = rows;
  out->shape[1] = cols;
  out->strides[0] = s0;
  out->strides[1] = 1;
  return KUCD_OK;
}

int kucd_ctx_create(kucd_ctx** out, int device_id, uint64_t seed) {
  if (out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KUCD_ERR_CUDA, "no CUDA device is visible; the engine has no CPU path");
  }
  if (device_id < 0 || device_id >= ndev) return fail(KUCD_ERR_INVALID_ARG, "device %d of %d", device_id, ndev);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10)
    return fail(KUCD_ERR_NOT_SM100, "device %d (%s) is sm_%d%d; the kernels are sm_100a only", device_id, prop.name,
                prop.major, prop.minor);
  CU_TRY(cudaSetDevice(device_id));
  {  // freed data-set planes stay in the device's pool for the next data set (see dataset_planes)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess && pool != nullptr) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  kucd_ctx* c = new (std::nothrow) kucd_ctx();
  if (c == nullptr) return fail(KUCD_ERR_INVALID_ARG, "out of host memory");
  c->device = device_id;
  c->num_sms = prop.multiProcessorCount;
  c->seed = seed;
  auto init = [&]() -> int {
    CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CU_TRY(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
    }
    for (auto& e : c->ev_units) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CU_TRY(cudaEventCreate(&c->ev0));
    CU_TRY(cudaEventCreate(&c->ev1));
    return KUCD_OK;
  };
  const int rc = init();
  if (rc != KUCD_OK) {
    const std::string why = g_err;
    kucd_ctx_destroy(c);  // tolerates the handles that were never created
    cudaGetLastError();
    g_err = why;
    return rc;
  }
  {
    const char* sp = getenv("KUCD_SPLIT");
    c->split = !(sp != nullptr && sp[0] == '0');
    c->split_force = sp != nullptr && sp[0] == '2';
    const char* ch = getenv("KUCD_CHAIN");
    c->chain = !(ch != nullptr && ch[0] == '0');
    c->chain_force = ch != nullptr && ch[0] == '2';
    const char* cd = getenv("KUCD_CHAIN_DW");
    c->chain_dw = cd != nullptr && cd[0] == '1';
  }
  *out = c;
  return KUCD_OK;
}

int kucd_ctx_destroy(kucd_ctx* ctx) {
  if (ctx == nullptr) return KUCD_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream != nullptr) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm != nullptr && g_nccl.CommDestroy != nullptr) g_nccl.CommDestroy(ctx->comm);
  ctx->stage_in.release();
  ctx->stage_u.release();
  ctx->stage_out.release();
  if (ctx->pinned_stats != nullptr) cudaFreeHost(ctx->pinned_stats);
  auto drop_event = [](cudaEvent_t e) {
    if (e != nullptr) cudaEventDestroy(e);
  };
  auto drop_stream = [](cudaStream_t s) {
    if (s != nullptr) cudaStreamDestroy(s);
  };
  for (cudaEvent_t e : ctx->ev_pool) drop_event(e);
  drop_event(ctx->ev0);
  drop_event(ctx->ev1);
  for (int i = 0; i < 2; ++i) {
    drop_event(ctx->ev_copied[i]);
    drop_event(ctx->ev_consumed[i]);
    ctx->stage_raw[i].release();
  }
  drop_event(ctx->ev_fork);
  drop_event(ctx->ev_join);
  for (cudaEvent_t e : ctx->ev_units) drop_event(e);
  for (int i = 0; i < kucd_ctx::kMaxSlabs; ++i) {
    drop_event(ctx->ev_slab[i]);
    drop_event(ctx->ev_red[i]);
  }
  drop_stream(ctx->comm_stream);
  drop_stream(ctx->copy_stream);
  drop_stream(ctx->stream3);
  drop_stream(ctx->stream2);
  drop_stream(ctx->stream);
  delete ctx;
  return KUCD_OK;
}

int kucd_sync(kucd_ctx* ctx) {
  if (ctx == nullptr) return fail(KUCD_ERR_INVALID_ARG, "ctx is NULL");
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_get_timings(kucd_ctx* ctx, kucd_timings* out, int reset) {
  if (ctx == nullptr || out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  prof_collect(ctx);
  *out = ctx->tm;
  if (reset) ctx->tm = kucd_timings{};
  return KUCD_OK;
}

int kucd_ctx_set_profile(kucd_ctx* ctx, int enable) {
  if (ctx == nullptr) return fail(KUCD_ERR_INVALID_ARG, "ctx is NULL");
  prof_collect(ctx);
  ctx->profile = enable != 0;
  return KUCD_OK;
}

int kucd_ctx_stream(kucd_ctx* ctx, void** stream_out) {
  if (ctx == nullptr || stream_out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  *stream_out = ctx->stream;
  return KUCD_OK;
}

int kucd_comm_unique_id(void* id128) {
  if (id128 == nullptr) return fail(KUCD_ERR_INVALID_ARG, "id128 is NULL");
  KU_TRY(nccl_load());
  NcclUid id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(KUCD_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
  memcpy(id128, &id, 128);
  return KUCD_OK;
}

int kucd_ctx_comm_init(kucd_ctx* ctx, const void* id128, int rank, int world) {
  if (ctx == nullptr || id128 == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(KUCD_ERR_INVALID_ARG, "rank %d of %d", rank, world);
  if (world == 1) {
    ctx->rank = 0;
    ctx->world = 1;
    return KUCD_OK;
  }
  KU_TRY(nccl_load());
  CU_TRY(cudaSetDevice(ctx->device));
  NcclUid id;
  memcpy(&id, id128, 128);
  void* comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, world, id, rank);
  if (rc != 0) return fail(KUCD_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return KUCD_OK;
}

// ---- model ---------------------------------------------------------------------------------------
int kucd_rbm_create(kucd_ctx* ctx, int64_t V, int64_t H, int mode, int compute, kucd_rbm** out) {
  if (ctx == nullptr || out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  if (V <= 0 || H <= 0 || V > (1 << 24) || H > (1 << 24))
    return fail(KUCD_ERR_INVALID_ARG, "bad dimensions V=%lld H=%lld", (long long)V, (long long)H);
  if (mode != KUCD_MODE_VISIBLE_BERNOULLI && mode != KUCD_MODE_VISIBLE_GAUSSIAN)
    return fail(KUCD_ERR_INVALID_ARG, "mode %d is not implemented (rbm.py:16 leaves MODE_COMPLEX as a TODO too)", mode);
  if (compute != KUCD_COMPUTE_BF16 && compute != KUCD_COMPUTE_F32X3)
    return fail(KUCD_ERR_INVALID_ARG, "compute %d", compute);
  CU_TRY(cudaSetDevice(ctx->device));
  kucd_rbm* r = new (std::nothrow) kucd_rbm();
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "out of host memory");
  r->ctx = ctx;
  r->seed = ctx->seed;
  r->V = V;
  r->H = H;
  r->ldV = round_up(V, 64);
  r->ldH = round_up(H, 64);
  r->mode = mode;
  r->compute = compute;
  r->wparts = compute == KUCD_COMPUTE_F32X3 ? 3 : 1;
  {  // opt-in, fixed for the life of the model (exchange slots and graphs are laid out for one element size)
    const char* e = getenv("KUCD_WIRE_BF16");
    r->wire16 = e != nullptr && e[0] == '1' && compute == KUCD_COMPUTE_BF16;
  }
  int rc = KUCD_OK;
  auto T = [&](int x) {
    if (rc == KUCD_OK) rc = x;
  };
  T(r->W32.ensure(static_cast<size_t>(V) * r->ldH * 4, true));
  T(r->Wp.ensure(V, r->ldH, r->wparts, /*shareable=*/true));
  for (int i = 0; i < r->wparts && rc == KUCD_OK; ++i)
    if (cudaMemset(r->Wp.buf[i].p, 0, r->Wp.buf[i].bytes) != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "memset");
  T(r->b32.ensure(r->ldVb() * 4, true));
  T(r->c32.ensure(r->ldHb() * 4, true));
  T(r->grad.ensure(r->grad_elems() * 4, true));
  T(r->stats.ensure(64 + 4 * kReconBlocks, true));  // [0..8] statistics and accumulators, [15] recon ticket, [16..] recon partials
  T(r->flag.ensure(16, true));
  T(r->dyn.ensure(sizeof(StepDyn), true));
  if (rc != KUCD_OK) {
    kucd_rbm_destroy(r);
    return rc;
  }
  *out = r;
  return KUCD_OK;
}

int kucd_rbm_destroy(kucd_rbm* r) {
  if (r == nullptr) return KUCD_OK;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  if (r->graph_exec != nullptr) cudaGraphExecDestroy(r->graph_exec);
  if (r->graph != nullptr) cudaGraphDestroy(r->graph);
  drop_host_graph(r);
  for (int i = 0; i < r->n_peer_open; ++i) cudaIpcCloseMemHandle(r->peer_open[i]);
  r->arena.release();
  for (DevBuf* b : {&r->W32, &r->b32, &r->c32, &r->mW, &r->mb, &r->mc, &r->grad, &r->fe0, &r->fe1, &r->sp0, &r->sp1,
                    &r->pstage, &r->stats, &r->flag, &r->dyn, &r->chain_done, &r->grad16, &r->gtmp})
    b->release();
  for (PlaneBuf* p : {&r->Wp, &r->vin, &r->vin2, &r->chunk, &r->ft_in, &r->ft_t, &r->ft_p, &r->h0, &r->hk, &r->vk,
                      &r->chains, &r->chains_g})
    p->release();
  delete r;
  return KUCD_OK;
}

int kucd_rbm_set_params(kucd_rbm* r, const kucd_tensor* W, const kucd_tensor* b, const kucd_tensor* c) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  if (W != nullptr) {
    KU_TRY(check_tensor(ctx, W, r->V, r->H, "rbm_weight", true));
    CU_TRY(cudaMemcpy2DAsync(r->W32.p, r->ldH * 4, W->data, W->strides[0] * 4, r->H * 4, r->V, cudaMemcpyDefault,
                             ctx->stream));
    KU_TRY(refresh_planes(r));
  }
  if (b != nullptr) {
    const kucd_tensor m = as_matrix(b);
    KU_TRY(check_tensor(ctx, &m, 1, r->V, "rbm_visible_bias", true));
    CU_TRY(cudaMemcpyAsync(r->b32.p, m.data, r->V * 4, cudaMemcpyDefault, ctx->stream));
  }
  if (c != nullptr) {
    const kucd_tensor m = as_matrix(c);
    KU_TRY(check_tensor(ctx, &m, 1, r->H, "rbm_hidden_bias", true));
    CU_TRY(cudaMemcpyAsync(r->c32.p, m.data, r->H * 4, cudaMemcpyDefault, ctx->stream));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_rbm_set_seed(kucd_rbm* r, uint64_t seed, uint64_t step_count) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  r->seed = seed;
  r->step_count = step_count;
  r->infer_draws = 0;
  r->score_draws = 0;
  return KUCD_OK;
}

// ---- fused reduction: peer-mapped exchange buffers ---------------------------------------------------
int kucd_rbm_peer_export(kucd_rbm* r, void* handle128) {
  if (r == nullptr || handle128 == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  const int n = ctx->world;
  if (n < 2 || n > 8) return fail(KUCD_ERR_INVALID_ARG, "the fused reduction needs 2..8 ranks (have %d)", n);
  if (ctx->comm == nullptr) return fail(KUCD_ERR_INVALID_ARG, "join the data-parallel group first");
  if (r->compute != KUCD_COMPUTE_BF16) return fail(KUCD_ERR_INVALID_ARG, "the fused reduction is built for bf16 compute");
  r->rows_per = (r->V + n - 1) / n;
  r->slice_elems = r->rows_per * r->ldH;
  r->bias_len = static_cast<int>(r->ldVb() + r->ldHb());
  // unit-sharded exchange: possible when every rank gets whole 128-unit groups of both layers (a 16-byte piece of a
  // packed row then never straddles two ranks' slices); costs two bit slots in the arena
  r->units_ok = r->mode == KUCD_MODE_VISIBLE_BERNOULLI && r->V % (128 * n) == 0 && r->H % (128 * n) == 0;
  const size_t bytes = (static_cast<size_t>(n) * r->slice_elems + static_cast<size_t>(n) * r->bias_len) * 4 + 1024 +
                       (r->units_ok ? 2 * static_cast<size_t>(kBitSlotBytes) : 0);
  KU_TRY(r->arena.ensure(bytes, false, /*shareable=*/true));
  CU_TRY(cudaMemset(r->arena.p, 0, r->arena.bytes));
  {  // canary checked by the attaching side: the mapping must start where this allocation starts
    const uint32_t magic = 0x4b554344u + static_cast<uint32_t>(ctx->rank);
    uint32_t* flags = reinterpret_cast<uint32_t*>(r->arena.as<float>() + static_cast<int64_t>(n) * r->slice_elems +
                                                  static_cast<int64_t>(n) * r->bias_len);
    CU_TRY(cudaMemcpy(flags + 128, &magic, 4, cudaMemcpyHostToDevice));
  }
  cudaIpcMemHandle_t h[2];
  CU_TRY(cudaIpcGetMemHandle(&h[0], r->arena.p));
  CU_TRY(cudaIpcGetMemHandle(&h[1], r->Wp.buf[0].p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(handle128, h, 128);
  return KUCD_OK;
}

int kucd_rbm_peer_attach(kucd_rbm* r, const void* handles) {
  if (r == nullptr || handles == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  const int n = ctx->world, me = ctx->rank;
  if (r->arena.p == nullptr) return fail(KUCD_ERR_INVALID_ARG, "kucd_rbm_peer_export first");
  const char* hs = static_cast<const char*>(handles);
  for (int j = 0; j < n; ++j) {
    void *pa = nullptr, *pw = nullptr;
    if (j == me) {
      pa = r->arena.p;
      pw = r->Wp.buf[0].p;
    } else {
      cudaIpcMemHandle_t h[2];
      memcpy(h, hs + 128 * j, 128);
      CU_TRY(cudaIpcOpenMemHandle(&pa, h[0], cudaIpcMemLazyEnablePeerAccess));
      r->peer_open[r->n_peer_open++] = pa;
      CU_TRY(cudaIpcOpenMemHandle(&pw, h[1], cudaIpcMemLazyEnablePeerAccess));
      r->peer_open[r->n_peer_open++] = pw;
    }
    r->ps.dw_slot[j] = static_cast<float*>(pa);
    r->ps.bias_slot[j] = r->ps.dw_slot[j] + static_cast<int64_t>(n) * r->slice_elems;
    r->ps.flags[j] = reinterpret_cast<uint32_t*>(r->ps.bias_slot[j] + static_cast<int64_t>(n) * r->bias_len);
    r->ps.bits[j] = r->units_ok ? reinterpret_cast<uint8_t*>(r->ps.flags[j]) + 1024 : nullptr;
    r->ps.wp[j] = static_cast<__nv_bfloat16*>(pw);
    uint32_t magic = 0;
    CU_TRY(cudaMemcpy(&magic, r->ps.flags[j] + 128, 4, cudaMemcpyDefault));
    if (magic != 0x4b554344u + static_cast<uint32_t>(j))
      return fail(KUCD_ERR_CUDA, "peer mapping of rank %d does not show its canary (got %08x)", j, magic);
  }
  r->epoch = r->ps.flags[me] + 64;
  r->peer_on = true;

  drop_host_graph(r);
  if (r->graph_exec != nullptr) {
    cudaGraphExecDestroy(r->graph_exec);
    cudaGraphDestroy(r->graph);
    r->graph_exec = nullptr;
    r->graph = nullptr;
  }
  return KUCD_OK;
}

int kucd_rbm_peer_detach(kucd_rbm* r) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  CU_TRY(cudaSetDevice(r->ctx->device));
  CU_TRY(cudaStreamSynchronize(r->ctx->stream));
  for (int i = 0; i < r->n_peer_open; ++i) cudaIpcCloseMemHandle(r->peer_open[i]);
  r->n_peer_open = 0;
  r->peer_on = false;
  r->units_ok = false;
  drop_host_graph(r);
  if (r->graph_exec != nullptr) {
    cudaGraphExecDestroy(r->graph_exec);
    cudaGraphDestroy(r->graph);
    r->graph_exec = nullptr;
    r->graph = nullptr;
  }
  return KUCD_OK;
}

int kucd_rbm_get_counters(kucd_rbm* r, uint64_t* seed, uint64_t* step_count, int64_t* n_chains) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  if (seed != nullptr) *seed = r->seed;
  if (step_count != nullptr) *step_count = r->step_count;
  if (n_chains != nullptr) *n_chains = r->n_chains;
  return KUCD_OK;
}

int kucd_rbm_get_draw_counters(kucd_rbm* r, uint64_t* infer_draws, uint64_t* score_draws) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  if (infer_draws != nullptr) *infer_draws = r->infer_draws;
  if (score_draws != nullptr) *score_draws = r->score_draws;
  return KUCD_OK;
}

int kucd_rbm_set_draw_counters(kucd_rbm* r, uint64_t infer_draws, uint64_t score_draws) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  r->infer_draws = infer_draws;
  r->score_draws = score_draws;
  return KUCD_OK;
}

// momentum buffers in and out (checkpoints): same layouts as the parameters (mW padded to ldH like W32)
int kucd_rbm_get_momentum(kucd_rbm* r, kucd_tensor* mW, kucd_tensor* mb, kucd_tensor* mc, int* present) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  const bool have = r->mW.p != nullptr && r->mb.p != nullptr && r->mc.p != nullptr;
  if (present != nullptr) *present = have ? 1 : 0;
  if (!have && (mW != nullptr || mb != nullptr || mc != nullptr)) {
    // asked for buffers that no momentum step has created yet: create them (zeroed), so that there is something to copy
    KU_TRY(r->mW.ensure(static_cast<size_t>(r->V) * r->ldH * 4, true));
    KU_TRY(r->mb.ensure(r->ldVb() * 4, true));
    KU_TRY(r->mc.ensure(r->ldHb() * 4, true));
  }
  if (mW != nullptr) {
    KU_TRY(check_tensor(ctx, mW, r->V, r->H, "momentum of rbm_weight", true));
    CU_TRY(cudaMemcpy2DAsync(mW->data, mW->strides[0] * 4, r->mW.p, r->ldH * 4, r->H * 4, r->V, cudaMemcpyDefault,
                             ctx->stream));
  }
  if (mb != nullptr) {
    const kucd_tensor m = as_matrix(mb);
    KU_TRY(check_tensor(ctx, &m, 1, r->V, "momentum of rbm_visible_bias", true));
    CU_TRY(cudaMemcpyAsync(m.data, r->mb.p, r->V * 4, cudaMemcpyDefault, ctx->stream));
  }
  if (mc != nullptr) {
    const kucd_tensor m = as_matrix(mc);
    KU_TRY(check_tensor(ctx, &m, 1, r->H, "momentum of rbm_hidden_bias", true));
    CU_TRY(cudaMemcpyAsync(m.data, r->mc.p, r->H * 4, cudaMemcpyDefault, ctx->stream));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_rbm_set_momentum(kucd_rbm* r, const kucd_tensor* mW, const kucd_tensor* mb, const kucd_tensor* mc) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  // all three exist together (apply_update allocates them together, zeroed)
  KU_TRY(r->mW.ensure(static_cast<size_t>(r->V) * r->ldH * 4, true));
  KU_TRY(r->mb.ensure(r->ldVb() * 4, true));
  KU_TRY(r->mc.ensure(r->ldHb() * 4, true));
  if (mW != nullptr) {
    KU_TRY(check_tensor(ctx, mW, r->V, r->H, "momentum of rbm_weight", true));
    CU_TRY(cudaMemcpy2DAsync(r->mW.p, r->ldH * 4, mW->data, mW->strides[0] * 4, r->H * 4, r->V, cudaMemcpyDefault,
                             ctx->stream));
  }
  if (mb != nullptr) {
    const kucd_tensor m = as_matrix(mb);
    KU_TRY(check_tensor(ctx, &m, 1, r->V, "momentum of rbm_visible_bias", true));
    CU_TRY(cudaMemcpyAsync(r->mb.p, m.data, r->V * 4, cudaMemcpyDefault, ctx->stream));
  }
  if (mc != nullptr) {
    const kucd_tensor m = as_matrix(mc);
    KU_TRY(check_tensor(ctx, &m, 1, r->H, "momentum of rbm_hidden_bias", true));
    CU_TRY(cudaMemcpyAsync(r->mc.p, m.data, r->H * 4, cudaMemcpyDefault, ctx->stream));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_rbm_get_params(kucd_rbm* r, kucd_tensor* W, kucd_tensor* b, kucd_tensor* c) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  if (W != nullptr) {
    KU_TRY(check_tensor(ctx, W, r->V, r->H, "rbm_weight", true));
    CU_TRY(cudaMemcpy2DAsync(W->data, W->strides[0] * 4, r->W32.p, r->ldH * 4, r->H * 4, r->V, cudaMemcpyDefault,
                             ctx->stream));
  }
  if (b != nullptr) {
    const kucd_tensor m = as_matrix(b);
    KU_TRY(check_tensor(ctx, &m, 1, r->V, "rbm_visible_bias", true));
    CU_TRY(cudaMemcpyAsync(m.data, r->b32.p, r->V * 4, cudaMemcpyDefault, ctx->stream));
  }
  if (c != nullptr) {
    const kucd_tensor m = as_matrix(c);
    KU_TRY(check_tensor(ctx, &m, 1, r->H, "rbm_hidden_bias", true));
    CU_TRY(cudaMemcpyAsync(m.data, r->c32.p, r->H * 4, cudaMemcpyDefault, ctx->stream));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

// ---- inference -------------------------------------------------------------------------------------
static const int64_t kChunkRows = 32768;
// streamed fits replay a captured step for latency-bound minibatches (see kucd_rbm_fit_host); KUCD_STREAM_GRAPH overrides
static const bool kStreamGraphDefault = false;

// forward = transform (rbm.py:45-48), !forward = inv_transform (rbm.py:51-54)
static int sample_api(kucd_rbm* r, bool forward, const kucd_tensor* in, kucd_tensor* s_out, kucd_tensor* p_out,
                      const kucd_tensor* u) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  const int64_t K = forward ? r->V : r->H, N = forward ? r->H : r->V;
  KU_TRY(check_tensor(ctx, in, -1, K, forward ? "v" : "h"));
  const int64_t rows = in->shape[0];
  if (s_out != nullptr) KU_TRY(check_tensor(ctx, s_out, rows, N, "sample output"));
  if (p_out != nullptr) KU_TRY(check_tensor(ctx, p_out, rows, N, "probability output"));
  if (p_out != nullptr && is_packed(p_out))
    return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "probabilities cannot be delivered as packed bits");
  if (s_out != nullptr && is_packed(s_out) && !forward && r->mode == KUCD_MODE_VISIBLE_GAUSSIAN)
    return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "Gaussian visible units are real-valued: no packed-bit output");
  if (u != nullptr) KU_TRY(check_tensor(ctx, u, rows, N, "injected draws", true));
  const bool gaussian = r->mode == KUCD_MODE_VISIBLE_GAUSSIAN;
  const uint64_t draw = (1ull << 63) | r->infer_draws++;
  const bool x3 = r->compute == KUCD_COMPUTE_F32X3;
  for (int64_t r0 = 0; r0 < rows; r0 += kChunkRows) {
    const int64_t n = std::min(kChunkRows, rows - r0);
    KU_TRY(ensure_workspace(r, n));
    Planes a;
    if (forward) {
      KU_TRY(ingest_batch(r, in, r0, n, &a));
    } else {
      // hidden states are 0/1 in every mode: one plane
      a = r->hk.view(n, r->H, 1);
      if (x3 && in->dtype_code == KUCD_DT_FLOAT) {
        Planes a3 = r->hk.view(n, r->H, 3);
        CU_TRY(cudaMemsetAsync(r->flag.p, 0, 4, ctx->stream));
        KU_TRY(ingest_rows(ctx, in, r0, n, a3, 3, r->flag.as<int>()));
        int flag = 0;
        CU_TRY(cudaMemcpyAsync(&flag, r->flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        a = a3;
        a.n = flag ? 3 : 1;
      } else {
        KU_TRY(ingest_rows(ctx, in, r0, n, a, 1, nullptr));
      }
    }
    EpiArgs e;
    if (forward) {
      e.epi = gaussian ? kEpiReluSample : kEpiSample;
      e.out = r->h0.view(n, r->H, 1);
    } else {
      e.epi = gaussian ? kEpiGaussian : kEpiSample;
      e.out = r->vk.view(n, r->V, vis_parts_out(r));
    }
    if (p_out != nullptr) {
      e.out_f32 = r->pstage.as<float>();
      e.ld_f32 = forward ? r->ldH : r->ldV;
    }
    KU_TRY(fetch_u(ctx, u, r0, n, N, &e.u, &e.ld_u));
    e.draw = draw;
    e.row0 = r0;
    KU_TRY(project(r, forward, a, n, e));
    if (s_out != nullptr) KU_TRY(deliver_planes(ctx, e.out, n, s_out, r0));
    if (p_out != nullptr) KU_TRY(deliver_f32(ctx, e.out_f32, e.ld_f32, n, p_out, r0));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_rbm_transform(kucd_rbm* r, const kucd_tensor* v, kucd_tensor* h_out, kucd_tensor* p_out,
                       const kucd_tensor* u) {
  return sample_api(r, true, v, h_out, p_out, u);
}

int kucd_rbm_inv_transform(kucd_rbm* r, const kucd_tensor* h, kucd_tensor* v_out, kucd_tensor* p_out,
                           const kucd_tensor* u) {
  return sample_api(r, false, h, v_out, p_out, u);
}

int kucd_rbm_free_energy(kucd_rbm* r, const kucd_tensor* v, kucd_tensor* fe_out) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_tensor(ctx, v, -1, r->V, "v"));
  const int64_t rows = v->shape[0];
  if (fe_out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "fe_out is NULL");
  kucd_tensor fo = *fe_out;  // (rows,) or (rows,1) float32
  if (!(fo.dtype_code == KUCD_DT_FLOAT && fo.bits == 32)) return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "fe_out must be float32");
  if (fo.shape[0] != rows || fo.shape[1] != 1)
    return fail(KUCD_ERR_SHAPE_MISMATCH, "fe_out has shape (%lld, %lld), expected (%lld, 1)", (long long)fo.shape[0],
                (long long)fo.shape[1], (long long)rows);
  for (int64_t r0 = 0; r0 < rows; r0 += kChunkRows) {
    const int64_t n = std::min(kChunkRows, rows - r0);
    KU_TRY(ensure_workspace(r, n));
    Planes a;
    KU_TRY(ingest_batch(r, v, r0, n, &a));
    KU_TRY(free_energy_into(r, a, n, r->sp0.as<float>(), r->fe0.as<float>(), nullptr, false));
    KU_TRY(deliver_f32(ctx, r->fe0.as<float>(), 1, n, &fo, r0));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

// ---- training --------------------------------------------------------------------------------------
static int check_hparams(const kucd_hparams* hp) {
  if (hp == nullptr) return fail(KUCD_ERR_INVALID_ARG, "hparams is NULL");
  if (hp->k < 1 || hp->k >= KUCD_MAX_K) return fail(KUCD_ERR_INVALID_ARG, "k = %d is outside [1, %d)", hp->k, KUCD_MAX_K);
  if ((hp->update_mask & ~KUCD_UPDATE_ALL) != 0) return fail(KUCD_ERR_INVALID_ARG, "update_mask %d", hp->update_mask);
  return KUCD_OK;
}

static int read_stats(kucd_rbm* r, kucd_step_stats* stats, int64_t rows) {
  float host[4];
  CU_TRY(cudaMemcpyAsync(host, r->stats.p, sizeof host, cudaMemcpyDeviceToHost, r->ctx->stream));
  CU_TRY(cudaStreamSynchronize(r->ctx->stream));
  r->ctx->tm.d2h_bytes += sizeof host;
  stats->score = host[0];
  stats->recon_err = host[1];
  stats->fe_mean = host[2];
  stats->rows = static_cast<int32_t>(rows);
  return KUCD_OK;
}

int kucd_rbm_cd_step(kucd_rbm* r, const kucd_tensor* v_batch, const kucd_hparams* hp, const kucd_inject* inj,
                     int64_t global_row0, kucd_step_stats* stats) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_hparams(hp));
  KU_TRY(check_tensor(ctx, v_batch, -1, r->V, "v_batch"));
  const int64_t rows = v_batch->shape[0];
  if (rows == 0) return KUCD_OK;  // an empty minibatch updates nothing
  if (rows > (1 << 22)) return fail(KUCD_ERR_INVALID_ARG, "minibatch of %lld rows", (long long)rows);
  KU_TRY(ensure_workspace(r, rows));
  if (hp->persistent) KU_TRY(ensure_chains(r, rows));
  // (unit-sharded only without injected draws and per-step statistics; every rank passes a shard of the same size)
  KU_TRY(prepare_exchange(r, rows, hp, inj == nullptr && !(stats != nullptr && hp->want_stats)));
  Planes v0;
  if (r->units_now) {
    // the gathered global minibatch lives in `vin`; this rank's rows are staged in `vin2`
    KU_TRY(prepare_units(r, rows, hp));
    KU_TRY(r->vin2.ensure(round_up(rows, 128), r->ldV, 1));
    v0 = r->vin2.view(rows, r->V, 1);
    KU_TRY(ingest_rows(ctx, v_batch, 0, rows, v0, 1, nullptr));
  } else {
    KU_TRY(ingest_batch(r, v_batch, 0, rows, &v0));
  }

  // injected draws: each node's array is staged to the device up front
  StepInject si;
  ScopedBufs keep(inj != nullptr ? 2 * KUCD_MAX_K + 1 : 0);
  if (inj != nullptr) {
    int slot = 0;
    auto stage = [&](const kucd_tensor* u, int64_t cols, const float** p, int64_t* ld) -> int {
      if (u == nullptr) return KUCD_OK;
      KU_TRY(check_tensor(ctx, u, rows, cols, "injected draws", true));
      const void* q;
      KU_TRY(fetch_rows(ctx, u, 0, rows, keep[slot++], &q, ld));
      *p = static_cast<const float*>(q);
      return KUCD_OK;
    };
    for (int t = 0; t < hp->k; ++t) KU_TRY(stage(inj->u_h[t], r->H, &si.u_h[t], &si.ld_h[t]));
    for (int t = 1; t <= hp->k; ++t) KU_TRY(stage(inj->u_v[t], r->V, &si.u_v[t], &si.ld_v[t]));
    KU_TRY(stage(inj->u_hc, r->H, &si.u_hc, &si.ld_hc));
  }
  KU_TRY(enqueue_cd(r, v0, rows, hp, inj ? &si : nullptr, global_row0, r->step_count, nullptr, false));
  r->step_count++;
  if (stats != nullptr && hp->want_stats) KU_TRY(enqueue_recon(r, v0, rows));
  KU_TRY(apply_update(r, hp, rows * ctx->world));
  if (stats != nullptr && hp->want_stats) {
    // the score chain reuses the state buffers, which is why last_stats must be read before asking for stats
    if (hp->want_stats == 1) KU_TRY(enqueue_score(r, v0, rows, nullptr, 0, nullptr, 0, global_row0));
    KU_TRY(read_stats(r, stats, rows));
  }
  KU_TRY(gather_master(r));
  const bool host_in = !on_device(v_batch) || inj != nullptr;
  if (host_in) CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

// Fine-tuning after the path.  One delta-rule step of a directed sigmoid layer laid out like the
// RBM: p = sigmoid(in.W + c) [forward] or sigmoid(in.W^T + b) [backward], then W += lr in^T (t - p) [or (t - p)^T in]
// and the matching bias += lr sum_rows (t - p).  in^T t - in^T p is the CD statistic v0^T h0 - vk^T hk with
// (v0, h0, vk, hk) = (in, t, in, p) [forward] or (t, in, p, in) [backward], so the step is the projection kernel with
// the probability epilogue, the two-segment dW contraction and the update kernel of the CD path.
int kucd_rbm_delta_rule(kucd_rbm* r, int forward, const kucd_tensor* in, const kucd_tensor* target, float lr,
                        int normalize) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  if (ctx->comm != nullptr) return fail(KUCD_ERR_INVALID_ARG, "delta_rule runs on a single rank (no data-parallel group)");
  if (r->mode != KUCD_MODE_VISIBLE_BERNOULLI && !forward)
    return fail(KUCD_ERR_INVALID_ARG, "delta_rule towards Gaussian visible units is not implemented (sigmoid units only)");
  const int64_t K = forward ? r->V : r->H, N = forward ? r->H : r->V;
  KU_TRY(check_tensor(ctx, in, -1, K, "in"));
  const int64_t rows = in->shape[0];
  KU_TRY(check_tensor(ctx, target, rows, N, "target"));
  if (rows == 0) return KUCD_OK;
  if (rows > (1 << 22)) return fail(KUCD_ERR_INVALID_ARG, "minibatch of %lld rows", (long long)rows);
  const bool x3 = r->compute == KUCD_COMPUTE_F32X3;
  const int np = x3 ? 3 : 1;
  const int64_t cap = round_up(rows, 128);
  KU_TRY(r->ft_in.ensure(cap, forward ? r->ldV : r->ldH, np));
  KU_TRY(r->ft_t.ensure(cap, forward ? r->ldH : r->ldV, np));
  KU_TRY(r->ft_p.ensure(cap, forward ? r->ldH : r->ldV, np));
  // caller rows -> operand planes; in fp32-grade mode float data is checked for being exact in one bf16 term (0/1
  // states are), which keeps the contraction at 1 + 3 term products
  auto ingest = [&](const kucd_tensor* t, PlaneBuf& buf, int64_t cols, Planes* out) -> int {
    Planes dst = buf.view(rows, cols, np);
    const bool maybe_inexact = x3 && t->dtype_code == KUCD_DT_FLOAT;
    if (maybe_inexact) CU_TRY(cudaMemsetAsync(r->flag.p, 0, 4, ctx->stream));
    KU_TRY(ingest_rows(ctx, t, 0, rows, dst, np, maybe_inexact ? r->flag.as<int>() : nullptr));
    int live = 1;
    if (maybe_inexact) {
      int flag = 0;
      CU_TRY(cudaMemcpyAsync(&flag, r->flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
      CU_TRY(cudaStreamSynchronize(ctx->stream));
      live = flag ? 3 : 1;
    }
    dst.n = live;
    *out = dst;
    return KUCD_OK;
  };
  Planes A, T;
  KU_TRY(ingest(in, r->ft_in, K, &A));
  KU_TRY(ingest(target, r->ft_t, N, &T));
  CU_TRY(cudaMemsetAsync(r->db(), 0, (r->ldVb() + r->ldHb()) * 4, ctx->stream));
  float* bias_stat = forward ? r->dc() : r->db();
  const Planes P = r->ft_p.view(rows, N, np);
  {
    EpiArgs e;
    e.epi = kEpiProb;
    e.out = P;
    e.colsum = bias_stat;
    e.colsum_sign = -1.f;  // - sum_rows p
    KU_TRY(project(r, forward != 0, A, rows, e));
  }
  {  // + sum_rows t
    dim3 grid(static_cast<unsigned>((N / 2 + 1 + 127) / 128), static_cast<unsigned>((rows + 63) / 64));
    colsum_kernel<<<grid, 128, 0, ctx->stream>>>(T.p[0], T.mid(), T.lo(), T.ld, 0, static_cast<int32_t>(rows),
                                                 static_cast<int32_t>(N), nullptr, 1.f, bias_stat);
    ctx->tm.aux_launches++;
    CU_TRY(cudaGetLastError());
  }
  r->fused_now = false;
  r->units_now = false;
  r->slabs_now = 1;
  r->slabs_inflight = 0;
  if (forward) KU_TRY(delta_w(r, A, T, A, P, rows, nullptr, false));
  else KU_TRY(delta_w(r, T, A, P, A, rows, nullptr, false));
  kucd_hparams hp{};
  hp.lr = lr;
  hp.k = 1;
  hp.normalize = normalize ? 1 : 0;
  hp.update_mask = KUCD_UPDATE_W | (forward ? KUCD_UPDATE_C : KUCD_UPDATE_B);
  KU_TRY(apply_update(r, &hp, rows));
  if (!on_device(in) || !on_device(target)) CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_rbm_score(kucd_rbm* r, const kucd_tensor* v_batch, const kucd_tensor* u_h, const kucd_tensor* u_v,
                   float* score_out) {
  if (r == nullptr || score_out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_tensor(ctx, v_batch, -1, r->V, "v_batch"));
  const int64_t rows = v_batch->shape[0];
  if (rows == 0) {
    *score_out = 0.f;
    return KUCD_OK;
  }
  KU_TRY(ensure_workspace(r, rows));
  Planes v0;
  KU_TRY(ingest_batch(r, v_batch, 0, rows, &v0));
  ScopedBufs ku(2);
  DevBuf &ku_h = ku[0], &ku_v = ku[1];
  const float *ph = nullptr, *pv = nullptr;
  int64_t ldh = 0, ldv = 0;
  if (u_h != nullptr) {
    KU_TRY(check_tensor(ctx, u_h, rows, r->H, "u_h", true));
    const void* q;
    KU_TRY(fetch_rows(ctx, u_h, 0, rows, ku_h, &q, &ldh));
    ph = static_cast<const float*>(q);
  }
  if (u_v != nullptr) {
    KU_TRY(check_tensor(ctx, u_v, rows, r->V, "u_v", true));
    const void* q;
    KU_TRY(fetch_rows(ctx, u_v, 0, rows, ku_v, &q, &ldv));
    pv = static_cast<const float*>(q);
  }
  KU_TRY(enqueue_score(r, v0, rows, ph, ldh, pv, ldv, 0));
  kucd_step_stats st;
  KU_TRY(read_stats(r, &st, rows));
  *score_out = st.score;
  return KUCD_OK;
}

int kucd_rbm_last_stats(kucd_rbm* r, kucd_tensor* dW, kucd_tensor* db, kucd_tensor* dc, kucd_tensor* h_pos,
                        kucd_tensor* v_neg, kucd_tensor* h_neg) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  const int64_t rows = r->last_rows;
  if (dW != nullptr) {
    KU_TRY(check_tensor(ctx, dW, r->V, r->H, "dW", true));
    CU_TRY(cudaMemcpy2DAsync(dW->data, dW->strides[0] * 4, r->dW(), r->ldH * 4, r->H * 4, r->V, cudaMemcpyDefault,
                             ctx->stream));
  }
  if (db != nullptr) {
    const kucd_tensor m = as_matrix(db);
    KU_TRY(check_tensor(ctx, &m, 1, r->V, "db", true));
    CU_TRY(cudaMemcpyAsync(m.data, r->db(), r->V * 4, cudaMemcpyDefault, ctx->stream));
  }
  if (dc != nullptr) {
    const kucd_tensor m = as_matrix(dc);
    KU_TRY(check_tensor(ctx, &m, 1, r->H, "dc", true));
    CU_TRY(cudaMemcpyAsync(m.data, r->dc(), r->H * 4, cudaMemcpyDefault, ctx->stream));
  }
  if (h_pos != nullptr) {
    KU_TRY(check_tensor(ctx, h_pos, rows, r->H, "h_pos"));
    KU_TRY(deliver_planes(ctx, r->h0.view(rows, r->H, 1), rows, h_pos, 0));
  }
  if (v_neg != nullptr) {
    KU_TRY(check_tensor(ctx, v_neg, rows, r->V, "v_neg"));
    if (is_packed(v_neg) && r->mode == KUCD_MODE_VISIBLE_GAUSSIAN)
      return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "Gaussian visible units are real-valued: no packed-bit output");
    KU_TRY(deliver_planes(ctx, r->vk.view(rows, r->V, r->last_vk_parts), rows, v_neg, 0));
  }
  if (h_neg != nullptr) {
    KU_TRY(check_tensor(ctx, h_neg, rows, r->H, "h_neg"));
    if (is_packed(h_neg)) return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "h_neg holds probabilities: no packed-bit output");
    KU_TRY(deliver_planes(ctx, r->hk.view(rows, r->H, r->last_hk_parts), rows, h_neg, 0));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_rbm_set_chains(kucd_rbm* r, const kucd_tensor* v_chains) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_tensor(ctx, v_chains, -1, r->V, "v_chains"));
  const int64_t n = v_chains->shape[0];
  const int np = vis_parts_out(r);
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  KU_TRY(r->chains.ensure(round_up(std::max<int64_t>(n, 1), 128), r->ldV, np));
  for (int i = 0; i < np; ++i) CU_TRY(cudaMemsetAsync(r->chains.buf[i].p, 0, r->chains.buf[i].bytes, ctx->stream));
  KU_TRY(ingest_rows(ctx, v_chains, 0, n, r->chains.view(n, r->V, np), np, nullptr));
  r->n_chains = n;
  r->chains_g_valid = false;
  r->last_vk_parts = np;
  drop_host_graph(r);
  if (r->graph_exec != nullptr) {  // chain buffers may have moved
    cudaGraphExecDestroy(r->graph_exec);
    cudaGraphDestroy(r->graph);
    r->graph_exec = nullptr;
    r->graph = nullptr;
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_rbm_get_chains(kucd_rbm* r, kucd_tensor* v_chains) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_tensor(ctx, v_chains, r->n_chains, r->V, "v_chains"));
  if (is_packed(v_chains) && r->mode == KUCD_MODE_VISIBLE_GAUSSIAN)
    return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "Gaussian visible units are real-valued: no packed-bit output");
  KU_TRY(deliver_planes(ctx, r->chains.view(r->n_chains, r->V, vis_parts_out(r)), r->n_chains, v_chains, 0));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

// ---- data sets -------------------------------------------------------------------------------------
int kucd_dataset_create(kucd_ctx* ctx, const kucd_tensor* data, int compute, kucd_dataset** out) {
  if (ctx == nullptr || out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_tensor(ctx, data, -1, -1, "data"));
  const int64_t rows = data->shape[0], dim = data->shape[1];
  if (dim <= 0) return fail(KUCD_ERR_INVALID_ARG, "data has no columns");
  kucd_dataset* ds = new (std::nothrow) kucd_dataset();
  if (ds == nullptr) return fail(KUCD_ERR_INVALID_ARG, "out of host memory");
  ds->ctx = ctx;
  ds->rows = rows;
  ds->dim = dim;
  const bool x3 = compute == KUCD_COMPUTE_F32X3 && data->dtype_code == KUCD_DT_FLOAT;
  const int np = x3 ? 3 : 1;
  int rc = dataset_planes(ctx, ds->planes, std::max<int64_t>(rows, 1), round_up(dim, 64), np);
  DevBuf flag;
  if (rc == KUCD_OK) rc = flag.ensure(16, true);
  for (int64_t r0 = 0; r0 < rows && rc == KUCD_OK; r0 += kChunkRows) {
    const int64_t n = std::min(kChunkRows, rows - r0);
    Planes dst = ds->planes.view(n, dim, np);
    for (int i = 0; i < np; ++i) dst.p[i] += r0 * dst.ld;
    rc = ingest_rows(ctx, data, r0, n, dst, np, x3 ? flag.as<int>() : nullptr);
    // the host staging buffer is reused by the next chunk
    if (rc == KUCD_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "sync failed");
  }
  int live = 1;
  if (rc == KUCD_OK && x3) {
    int f = 0;
    if (cudaMemcpy(&f, flag.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "flag read failed");
    live = f ? 3 : 1;
  }
  flag.release();
  if (rc != KUCD_OK) {
    dataset_planes_release(ctx, ds->planes);
    delete ds;
    return rc;
  }
  ds->nparts = live;
  *out = ds;
  return KUCD_OK;
}

int kucd_dataset_destroy(kucd_dataset* ds) {
  if (ds == nullptr) return KUCD_OK;
  cudaSetDevice(ds->ctx->device);
  dataset_planes_release(ds->ctx, ds->planes);  // freed in stream order: no synchronisation
  delete ds;
  return KUCD_OK;
}

int kucd_dataset_shape(kucd_dataset* ds, int64_t* rows, int64_t* dim) {
  if (ds == nullptr) return fail(KUCD_ERR_INVALID_ARG, "dataset is NULL");
  if (rows != nullptr) *rows = ds->rows;
  if (dim != nullptr) *dim = ds->dim;
  return KUCD_OK;
}

int kucd_dataset_read(kucd_dataset* ds, kucd_tensor* out) {
  if (ds == nullptr) return fail(KUCD_ERR_INVALID_ARG, "dataset is NULL");
  kucd_ctx* ctx = ds->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_tensor(ctx, out, ds->rows, ds->dim, "out"));
  if (is_packed(out) && ds->nparts != 1)
    return fail(KUCD_ERR_UNSUPPORTED_DTYPE, "this data set holds real values: no packed-bit output");
  for (int64_t r0 = 0; r0 < ds->rows; r0 += kChunkRows) {
    const int64_t n = std::min(kChunkRows, ds->rows - r0);
    Planes src = ds->view();
    for (int i = 0; i < 3; ++i)
      if (src.p[i] != nullptr) src.p[i] += r0 * src.ld;
    KU_TRY(deliver_planes(ctx, src, n, out, r0));
  }
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return KUCD_OK;
}

int kucd_dataset_shuffle(kucd_dataset* in, uint64_t seed, uint64_t epoch, kucd_dataset** inout) {
  if (in == nullptr || inout == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  kucd_ctx* ctx = in->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  kucd_dataset* out = *inout;
  if (out == in) return fail(KUCD_ERR_INVALID_ARG, "a data set cannot be shuffled in place");
  if (out != nullptr && (out->ctx != ctx || out->rows != in->rows || out->dim != in->dim || out->planes.ld != in->planes.ld))
    return fail(KUCD_ERR_SHAPE_MISMATCH, "the destination data set has another shape");
  bool created = false;
  if (out == nullptr) {
    out = new (std::nothrow) kucd_dataset();
    if (out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "out of host memory");
    out->ctx = ctx;
    out->rows = in->rows;
    out->dim = in->dim;
    created = true;
  }
  int rc = KUCD_OK;
  if (created || out->nparts < in->nparts)
    rc = dataset_planes(ctx, out->planes, std::max<int64_t>(in->rows, 1), in->planes.ld, in->nparts);
  if (rc == KUCD_OK && in->rows > 0) {
    const FeistelKey key = make_feistel_key(seed, epoch, in->rows);
    const bool three = in->nparts == 3;
    const int grid = static_cast<int>(std::min<int64_t>(in->rows, static_cast<int64_t>(ctx->num_sms) * 16));
    permute_rows_kernel<<<grid, 128, 0, ctx->stream>>>(
        in->planes.buf[0].as<__nv_bfloat16>(), three ? in->planes.buf[1].as<__nv_bfloat16>() : nullptr,
        three ? in->planes.buf[2].as<__nv_bfloat16>() : nullptr, out->planes.buf[0].as<__nv_bfloat16>(),
        three ? out->planes.buf[1].as<__nv_bfloat16>() : nullptr, three ? out->planes.buf[2].as<__nv_bfloat16>() : nullptr,
        in->planes.ld, in->rows, key);
    ctx->tm.aux_launches++;
    if (cudaGetLastError() != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "permute_rows_kernel launch failed");
  }
  if (rc != KUCD_OK) {
    if (created) {
      dataset_planes_release(ctx, out->planes);
      delete out;
    }
    return rc;
  }
  out->nparts = in->nparts;
  *inout = out;
  return KUCD_OK;
}

static int fit_range_impl(kucd_rbm* r, kucd_dataset* ds, int64_t batch, const kucd_hparams* hp, int64_t global_row0,
                          int64_t step_begin, int64_t step_end, kucd_epoch_stats* stats) {
  if (r == nullptr || ds == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_hparams(hp));
  if (ds->ctx != ctx) return fail(KUCD_ERR_INVALID_ARG, "data set belongs to another context");
  if (ds->dim != r->V)
    return fail(KUCD_ERR_SHAPE_MISMATCH, "data set has %lld columns, the RBM %lld visible units", (long long)ds->dim,
                (long long)r->V);
  if (batch < 1 || batch > (1 << 22)) return fail(KUCD_ERR_INVALID_ARG, "batch_size %lld", (long long)batch);
  const int64_t N = ds->rows;
  const int64_t epoch_steps = (N + batch - 1) / batch;  // rbm.py:110-111
  if (stats != nullptr) memset(stats, 0, sizeof *stats);
  if (step_end < 0) step_end = epoch_steps;
  if (step_begin < 0 || step_begin > step_end || step_end > epoch_steps)
    return fail(KUCD_ERR_INVALID_ARG, "minibatch range [%lld, %lld) of %lld", (long long)step_begin, (long long)step_end,
                (long long)epoch_steps);
  const int64_t steps = step_end - step_begin;
  if (steps == 0) return KUCD_OK;
  // unit-sharded only when every minibatch of the range is a full one and the data set is a single 0/1 plane
  KU_TRY(prepare_exchange(r, batch, hp, step_end * batch <= N && ds->nparts == 1));
  KU_TRY(ensure_workspace(r, batch));
  if (hp->persistent) KU_TRY(ensure_chains(r, std::min(batch, N)));
  if (r->units_now) KU_TRY(prepare_units(r, batch, hp));
  if (hp->momentum != 0.f) {  // allocate outside the capture
    KU_TRY(r->mW.ensure(static_cast<size_t>(r->V) * r->ldH * 4, true));
    KU_TRY(r->mb.ensure(r->ldVb() * 4, true));
    KU_TRY(r->mc.ensure(r->ldHb() * 4, true));
  }

  GraphKey key;
  key.ds = ds;
  key.ds_id = ds->id;
  key.ds_ptr = ds->planes.buf[0].p;
  key.batch = batch;
  key.global_row0 = global_row0;
  key.ds_rows = N;
  key.ds_parts = ds->nparts;
  key.hp = *hp;
  key.hp.want_stats = 0;
  key.fused = r->fused_now;
  key.units = r->units_now;
  key.chains_g = r->units_now ? r->chains_g.buf[0].p : nullptr;
  const Planes v0 = ds->view();
  StepDyn* dyn = r->dyn.as<StepDyn>();
  if (r->graph_exec == nullptr || !(r->graph_key == key)) {
    if (r->graph_exec != nullptr) {
      cudaGraphExecDestroy(r->graph_exec);
      cudaGraphDestroy(r->graph);
      r->graph_exec = nullptr;
      r->graph = nullptr;
    }
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    const int64_t k0 = ctx->tm.gemm_launches + ctx->tm.aux_launches;
    const kucd_timings t0 = ctx->tm;
    CU_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    int rc = enqueue_cd(r, v0, batch, hp, nullptr, global_row0, 0, dyn, true);
    bool advanced = false;
    if (rc == KUCD_OK) rc = apply_update(r, hp, batch * ctx->world, dyn, static_cast<int32_t>(batch), N, &advanced);
    if (rc == KUCD_OK && !advanced) {
      advance_dyn_kernel<<<1, 1, 0, ctx->stream>>>(dyn, static_cast<int32_t>(batch), N);
      ctx->tm.aux_launches++;
    }
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
    if (rc != KUCD_OK) {
      if (g != nullptr) cudaGraphDestroy(g);
      return rc;
    }
    if (ce != cudaSuccess) return fail(KUCD_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    cudaGraphExec_t ge = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
    if (ie != cudaSuccess) {
      cudaGraphDestroy(g);
      return fail(KUCD_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
    }
    r->graph = g;
    r->graph_exec = ge;
    r->graph_key = key;
    // the launches counted while capturing were recorded, not run: they are what one replay launches
    r->graph_kernels = ctx->tm.gemm_launches + ctx->tm.aux_launches - k0;
    // the exchange counters were bumped by the capture, which ran nothing: they count executed steps (below)
    r->graph_unit_ex = ctx->tm.unit_exchanges - t0.unit_exchanges;
    r->graph_fused = ctx->tm.fused_reduce_steps - t0.fused_reduce_steps;
    r->graph_ar = ctx->tm.allreduce_calls - t0.allreduce_calls;
    ctx->tm.unit_exchanges = t0.unit_exchanges;
    ctx->tm.unit_steps = t0.unit_steps;
    ctx->tm.fused_reduce_steps = t0.fused_reduce_steps;
    ctx->tm.allreduce_calls = t0.allreduce_calls;
  }
  set_dyn_kernel<<<1, 1, 0, ctx->stream>>>(dyn, step_begin * batch,
                                           static_cast<int32_t>(std::min(batch, N - step_begin * batch)), r->step_count);
  ctx->tm.aux_launches++;
  CU_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int64_t s = 0; s < steps; ++s) CU_TRY(cudaGraphLaunch(r->graph_exec, ctx->stream));
  CU_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
  KU_TRY(gather_master(r));
  ctx->tm.graph_launches += steps;
  ctx->tm.graph_kernel_launches += steps * r->graph_kernels;
  ctx->tm.unit_exchanges += steps * r->graph_unit_ex;
  ctx->tm.unit_steps += r->graph_unit_ex > 0 ? steps : 0;
  ctx->tm.fused_reduce_steps += steps * r->graph_fused;
  ctx->tm.allreduce_calls += steps * r->graph_ar;
  r->step_count += steps;
  r->last_rows = r->units_now ? batch * ctx->world : batch;
  if (stats != nullptr) {
    stats->steps = steps;
    stats->rows = N;
    if (hp->want_stats) {
      // score of the last minibatch with a fresh chain, as the reference prints it (rbm.py:227-234)
      const int64_t r0 = (step_end - 1) * batch, n = std::min(batch, N - r0);
      Planes last = v0;
      for (int i = 0; i < 3; ++i)
        if (last.p[i] != nullptr) last.p[i] += r0 * last.ld;
      last.rows = n;
      KU_TRY(enqueue_recon(r, last, n));
      KU_TRY(enqueue_score(r, last, n, nullptr, 0, nullptr, 0, global_row0));
      kucd_step_stats st;
      KU_TRY(read_stats(r, &st, n));
      stats->last_score = st.score;
      stats->last_recon_err = st.recon_err;
    }
    CU_TRY(cudaEventSynchronize(ctx->ev1));
    CU_TRY(cudaEventElapsedTime(&stats->device_ms, ctx->ev0, ctx->ev1));
  }
  return KUCD_OK;
}

// Chunked variant of the streamed fit for latency-bound minibatches (C = 8 by default, KUCD_STREAM_CHUNK): C minibatches
// travel per copy, one ingest launch expands them into resident operand planes, and their
// steps replay the captured graph of kucd_rbm_fit_range (minibatch offset, draw counter in StepDyn) - one copy ->
// ingest -> step hand-over per C steps instead of per step.  The raw staging is double-buffered, so the copy of
// chunk j+1 overlaps the steps of chunk j.  The remainder minibatch (at most one, rbm.py:211) runs as direct
// launches.  Same draws, same arithmetic, same per-step statistics as the per-minibatch stream.
static int fit_host_chunked(kucd_rbm* r, const kucd_tensor* V_all, int64_t batch, const kucd_hparams* hp,
                            int64_t global_row0, float* host_stats, int64_t chunk_steps) {
  kucd_ctx* ctx = r->ctx;
  const int64_t N = V_all->shape[0];
  const int64_t steps = (N + batch - 1) / batch;
  const int64_t pitch = row_pitch_bytes(V_all), rb = row_bytes(V_all);
  const bool packed = is_packed(V_all);
  const int64_t cols = r->V;
  const bool x3 = r->compute == KUCD_COMPUTE_F32X3;
  const int nplanes = x3 ? 3 : 1;
  const int live = (x3 && V_all->dtype_code == KUCD_DT_FLOAT) ? 3 : 1;
  const int64_t chunk_rows = chunk_steps * batch;
  const int64_t nchunks = (N + chunk_rows - 1) / chunk_rows;
  for (int i = 0; i < 2; ++i) KU_TRY(ctx->stage_raw[i].ensure(static_cast<size_t>(chunk_rows) * rb));
  KU_TRY(r->chunk.ensure(chunk_rows, r->ldV, nplanes));
  if (hp->momentum != 0.f) {  // allocate outside the capture
    KU_TRY(r->mW.ensure(static_cast<size_t>(r->V) * r->ldH * 4, true));
    KU_TRY(r->mb.ensure(r->ldVb() * 4, true));
    KU_TRY(r->mc.ensure(r->ldHb() * 4, true));
  }
  StepDyn* dyn = r->dyn.as<StepDyn>();
  Planes block = r->chunk.view(chunk_rows, r->V, nplanes);
  block.n = live;

  HostGraphKey key;
  key.batch = batch;
  key.global_row0 = global_row0;
  key.hp = *hp;
  key.hp.want_stats = 0;
  key.fused = r->fused_now;
  key.live = live;
  key.vin = r->chunk.buf[0].p;
  key.log = host_stats;
  key.chunk_rows = chunk_rows;
  if (r->hgraph_exec == nullptr || !(r->hgraph_key == key)) {
    drop_host_graph(r);
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    const int64_t k0 = ctx->tm.gemm_launches + ctx->tm.aux_launches;
    const int64_t no_end = int64_t{1} << 60;  // replayed steps are full minibatches: rows_valid stays `batch`
    CU_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    int crc = enqueue_cd(r, block, batch, hp, nullptr, global_row0, 0, dyn, true);
    if (crc == KUCD_OK && host_stats != nullptr) {
      crc = enqueue_recon(r, block, batch, dyn, host_stats, dyn, static_cast<int32_t>(batch));
    }
    bool advanced = false;
    if (crc == KUCD_OK) crc = apply_update(r, hp, batch * ctx->world, dyn, static_cast<int32_t>(batch), no_end, &advanced);
    if (crc == KUCD_OK && !advanced) {
      advance_dyn_kernel<<<1, 1, 0, ctx->stream>>>(dyn, static_cast<int32_t>(batch), no_end);
      ctx->tm.aux_launches++;
    }
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
    if (crc != KUCD_OK) {
      if (g != nullptr) cudaGraphDestroy(g);
      return crc;
    }
    if (ce != cudaSuccess) return fail(KUCD_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    cudaGraphExec_t ge = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
    if (ie != cudaSuccess) {
      cudaGraphDestroy(g);
      return fail(KUCD_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
    }
    r->hgraph = g;
    r->hgraph_exec = ge;
    r->hgraph_key = key;
    r->hgraph_kernels = ctx->tm.gemm_launches + ctx->tm.aux_launches - k0;  // recorded while capturing, not run
  }

  auto rows_of_chunk = [&](int64_t j) { return std::min(chunk_rows, N - j * chunk_rows); };
  auto copy_in = [&](int64_t j) -> int {
    const int slot = static_cast<int>(j & 1);
    if (j >= 2) CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[slot], 0));
    const char* src = static_cast<const char*>(V_all->data) + j * chunk_rows * pitch;
    const int64_t n = rows_of_chunk(j);
    if (pitch == rb) {
      CU_TRY(cudaMemcpyAsync(ctx->stage_raw[slot].p, src, static_cast<size_t>(n) * rb, cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    } else {
      CU_TRY(cudaMemcpy2DAsync(ctx->stage_raw[slot].p, rb, src, pitch, rb, n, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    CU_TRY(cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
    ctx->tm.h2d_bytes += n * rb;
    return KUCD_OK;
  };
  auto ingest = [&](const void* src, int64_t n) {
    const int grid = grid_for(ctx, n * (block.ld / 8), 256);
    if (packed) {
      ingest_bits_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(src), rb, n, cols, block.p[0],
                                                        block.p[1], block.p[2], block.ld, nplanes);
    } else {
      by_dtype(V_all, [&](auto* tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        ingest_kernel<T><<<grid, 256, 0, ctx->stream>>>(static_cast<const T*>(src), cols, n, cols, block.p[0], block.p[1],
                                                        block.p[2], block.ld, nplanes, nullptr);
        return 0;
      });
    }
    ctx->tm.aux_launches++;
  };

  int rc = KUCD_OK;
  if (cudaEventRecord(ctx->ev0, ctx->stream) != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "event record failed");
  if (rc == KUCD_OK) rc = copy_in(0);
  for (int64_t j = 0; j < nchunks && rc == KUCD_OK; ++j) {
    const int slot = static_cast<int>(j & 1);
    const int64_t n = rows_of_chunk(j);
    if (j + 1 < nchunks) rc = copy_in(j + 1);
    if (rc != KUCD_OK) break;
    if (cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[slot], 0) != cudaSuccess) {
      rc = fail(KUCD_ERR_CUDA, "stream wait failed");
      break;
    }
    ingest(ctx->stage_raw[slot].p, n);  // ordered after the previous chunk's steps: they read the same planes
    cudaEventRecord(ctx->ev_consumed[slot], ctx->stream);  // the raw staging slot is free again
    const int64_t full = n / batch;
    if (full > 0) {
      set_dyn_kernel<<<1, 1, 0, ctx->stream>>>(dyn, 0, static_cast<int32_t>(batch), r->step_count,
                                               static_cast<int32_t>(j * chunk_steps));
      ctx->tm.aux_launches++;
      for (int64_t s2 = 0; s2 < full && rc == KUCD_OK; ++s2)
        if (cudaGraphLaunch(r->hgraph_exec, ctx->stream) != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "graph launch failed");
      r->step_count += full;
      ctx->tm.graph_launches += full;
      ctx->tm.graph_kernel_launches += full * r->hgraph_kernels;
      if (host_stats != nullptr) ctx->tm.d2h_bytes += 4 * full;
    }
    const int64_t rem = n - full * batch;
    if (rem > 0 && rc == KUCD_OK) {  // the remainder minibatch (last chunk only)
      Planes v0 = block;
      for (int i = 0; i < 3; ++i)
        if (v0.p[i] != nullptr) v0.p[i] += full * batch * v0.ld;
      v0.rows = rem;
      rc = enqueue_cd(r, v0, rem, hp, nullptr, global_row0, r->step_count, nullptr, false);
      r->step_count++;
      if (rc == KUCD_OK && host_stats != nullptr) {
        rc = enqueue_recon(r, v0, rem);
        if (rc == KUCD_OK && cudaMemcpyAsync(host_stats + (steps - 1), r->stats.as<float>() + 1, 4, cudaMemcpyDeviceToHost,
                                             ctx->stream) != cudaSuccess)
          rc = fail(KUCD_ERR_CUDA, "statistic read-back failed");
        ctx->tm.d2h_bytes += 4;
      }
      if (rc == KUCD_OK) rc = apply_update(r, hp, rem * ctx->world);
    }
  }
  if (rc == KUCD_OK && cudaGetLastError() != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "chunked fit: launch failed");
  return rc;
}

// One pass over a HOST-resident matrix (rbm.py:163-223 with V as the caller's numpy array): the copy of
// minibatch i+1 (copy stream, double-buffered staging) overlaps the Gibbs chain of minibatch i; every
// step's reconstruction error is read back asynchronously into step_recon[i].
int kucd_rbm_fit_host(kucd_rbm* r, const kucd_tensor* V_all, int64_t batch, const kucd_hparams* hp, int64_t global_row0,
                      float* step_recon, kucd_epoch_stats* stats) {
  if (r == nullptr) return fail(KUCD_ERR_INVALID_ARG, "rbm is NULL");
  kucd_ctx* ctx = r->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  KU_TRY(check_hparams(hp));
  KU_TRY(check_tensor(ctx, V_all, -1, r->V, "V"));
  if (on_device(V_all)) return fail(KUCD_ERR_INVALID_ARG, "fit_host streams from host memory; use a data set for device data");
  if (batch < 1 || batch > (1 << 22)) return fail(KUCD_ERR_INVALID_ARG, "batch_size %lld", (long long)batch);
  const int64_t N = V_all->shape[0];
  const int64_t steps = (N + batch - 1) / batch;  // rbm.py:110-111
  if (stats != nullptr) memset(stats, 0, sizeof *stats);
  if (steps == 0) return KUCD_OK;
  KU_TRY(ensure_workspace(r, batch));
  if (hp->persistent) KU_TRY(ensure_chains(r, std::min(batch, N)));
  KU_TRY(prepare_exchange(r, batch));
  const int64_t pitch = row_pitch_bytes(V_all), rb = row_bytes(V_all);  // source row pitch / staged row size (bytes)
  const bool packed = is_packed(V_all);
  const int64_t cols = r->V;
  const bool x3 = r->compute == KUCD_COMPUTE_F32X3;
  const int nplanes = x3 ? 3 : 1;
  // Default: cudaMemcpyAsync on the copy stream into a raw staging slot, converted on the compute stream.
  // KUCD_ZEROCOPY=1 (pinned memory only): the ingest kernel reads the host array itself, straight over PCIe into
  // the operand planes on the copy stream.  Measured at C3 next to the chain kernel: DMA 2.62 ms per step vs
  // zero-copy 3.73 ms on one box; on another box of the pool the copy engine delivered only 4-14 GB/s under load
  // (tools/h2d_check.py), which is when the switch is worth trying.
  static const bool zc_env = [] {
    const char* e = getenv("KUCD_ZEROCOPY");
    return e != nullptr && e[0] == '1';
  }();
  const bool zero_copy = zc_env && V_all->device_type == KUCD_DEV_CUDA_HOST;
  if (zero_copy) {
    KU_TRY(r->vin2.ensure(r->cap, r->ldV, nplanes));
  } else {
    for (int i = 0; i < 2; ++i) KU_TRY(ctx->stage_raw[i].ensure(static_cast<size_t>(batch) * rb));
  }
  float* host_stats = nullptr;
  if (step_recon != nullptr) {
    if (ctx->pinned_stats_cap < static_cast<size_t>(steps)) {  // a page-locked allocation costs ~1 ms: keep it
      if (ctx->pinned_stats != nullptr) cudaFreeHost(ctx->pinned_stats);
      ctx->pinned_stats = nullptr;
      ctx->pinned_stats_cap = 0;
      const size_t cap = std::max<size_t>(static_cast<size_t>(steps), 1024);
      CU_TRY(cudaMallocHost(&ctx->pinned_stats, cap * sizeof(float)));
      ctx->pinned_stats_cap = cap;
    }
    host_stats = ctx->pinned_stats;
  }
  // Latency-bound minibatches travel and are ingested C at a time (fit_host_chunked): one copy -> ingest -> step
  // hand-over per C steps instead of per step.  Measured at the C1 shape (60000 x 784 float32 rows, minibatches of 128,
  // profiles/r02_call1_switches.log): 75.7 us per step unchunked, 61.5 / 61.1 / 61.6 / 61.4 us at C = 4 / 8 / 16 / 32.
  // KUCD_STREAM_CHUNK overrides C (0 or 1: per-minibatch stream).
  static const int64_t chunk_env = [] {
    const char* e = getenv("KUCD_STREAM_CHUNK");
    return e == nullptr ? int64_t{8} : static_cast<int64_t>(atoll(e));
  }();
  if (chunk_env > 1 && !zero_copy && N >= 2 * batch && batch * cols < (1 << 20)) {
    int rc = fit_host_chunked(r, V_all, batch, hp, global_row0, host_stats, std::min<int64_t>(chunk_env, 4096));
    cudaEventRecord(ctx->ev1, ctx->stream);
    if (rc == KUCD_OK) rc = gather_master(r);
    cudaStreamSynchronize(ctx->copy_stream);
    const cudaError_t se = cudaStreamSynchronize(ctx->stream);
    if (rc == KUCD_OK && se != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "fit_host: %s", cudaGetErrorString(se));
    if (host_stats != nullptr && rc == KUCD_OK) memcpy(step_recon, host_stats, steps * sizeof(float));
    if (rc == KUCD_OK && stats != nullptr) {
      stats->steps = steps;
      stats->rows = N;
      cudaEventElapsedTime(&stats->device_ms, ctx->ev0, ctx->ev1);
      if (step_recon != nullptr) stats->last_recon_err = step_recon[steps - 1];
    }
    return rc;
  }
  // fp32 data in fp32-grade mode is carried in all three terms (no per-step "is it exact" round trip)
  const int live = (x3 && V_all->dtype_code == KUCD_DT_FLOAT) ? 3 : 1;
  auto rows_of_step = [&](int64_t i) { return std::min(batch, N - i * batch); };
  auto slot_planes = [&](int slot, int64_t n) { return (slot == 0 ? r->vin : r->vin2).view(n, r->V, nplanes); };
  // src_ld: elements (bits, when packed) from one source row to the next
  auto ingest = [&](const void* src, int64_t src_ld, int64_t n, const Planes& dst, cudaStream_t st) {
    const int grid = grid_for(ctx, n * (dst.ld / 8), 256);
    if (packed) {
      ingest_bits_kernel<<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(src), src_ld / 8, n, cols, dst.p[0], dst.p[1],
                                               dst.p[2], dst.ld, nplanes);
      ctx->tm.aux_launches++;
      return;
    }
    by_dtype(V_all, [&](auto* tag) {
      using T = std::remove_pointer_t<decltype(tag)>;
      ingest_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(src), src_ld, n, cols, dst.p[0], dst.p[1], dst.p[2],
                                             dst.ld, nplanes, nullptr);
      return 0;
    });
    ctx->tm.aux_launches++;
  };
  auto copy_in = [&](int64_t i) -> int {
    const int slot = static_cast<int>(i & 1);
    if (i >= 2) CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[slot], 0));
    const char* src = static_cast<const char*>(V_all->data) + i * batch * pitch;
    const int64_t n = rows_of_step(i);
    if (zero_copy) {
      ingest(src, V_all->strides[0], n, slot_planes(slot, n), ctx->copy_stream);
      CU_TRY(cudaGetLastError());
    } else if (pitch == rb) {  // contiguous rows: one linear DMA
      CU_TRY(cudaMemcpyAsync(ctx->stage_raw[slot].p, src, static_cast<size_t>(n) * rb, cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    } else {
      CU_TRY(cudaMemcpy2DAsync(ctx->stage_raw[slot].p, rb, src, pitch, rb, n, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    CU_TRY(cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
    ctx->tm.h2d_bytes += n * rb;
    return KUCD_OK;
  };
  // Latency-bound minibatches: everything of a step after the ingest (column sums, chain, dW, statistic, update) is
  // captured once and replayed per full minibatch - one graph launch instead of nine dependent kernel launches - with
  // the draw counter in device memory (StepDyn) exactly as in kucd_rbm_fit_range.  The remainder minibatch runs through
  // the direct launches below.  KUCD_STREAM_GRAPH=1 turns it on for every size, 0 off.
  static const int graph_env = [] {
    const char* e = getenv("KUCD_STREAM_GRAPH");
    return e == nullptr ? -1 : atoi(e);
  }();
  const int64_t full_steps = N / batch;
  const bool use_graph = !zero_copy && full_steps >= 2 &&
                         (graph_env == 1 || (graph_env == -1 && kStreamGraphDefault && batch * cols < (1 << 20)));
  StepDyn* dyn = r->dyn.as<StepDyn>();
  if (use_graph) {
    if (hp->momentum != 0.f) {  // allocate outside the capture
      KU_TRY(r->mW.ensure(static_cast<size_t>(r->V) * r->ldH * 4, true));
      KU_TRY(r->mb.ensure(r->ldVb() * 4, true));
      KU_TRY(r->mc.ensure(r->ldHb() * 4, true));
    }
    HostGraphKey key;
    key.batch = batch;
    key.global_row0 = global_row0;
    key.hp = *hp;
    key.hp.want_stats = 0;
    key.fused = r->fused_now;
    key.live = live;
    key.vin = r->vin.buf[0].p;
    key.log = host_stats;
    if (r->hgraph_exec == nullptr || !(r->hgraph_key == key)) {
      drop_host_graph(r);
      CU_TRY(cudaStreamSynchronize(ctx->stream));
      const int64_t k0 = ctx->tm.gemm_launches + ctx->tm.aux_launches;
      const int64_t no_end = int64_t{1} << 60;  // the replayed steps are full minibatches: rows_valid stays `batch`
      CU_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
      Planes v0 = r->vin.view(batch, r->V, nplanes);
      v0.n = live;
      int crc = enqueue_cd(r, v0, batch, hp, nullptr, global_row0, 0, dyn, false);
      if (crc == KUCD_OK && host_stats != nullptr) {
        crc = enqueue_recon(r, v0, batch, nullptr, host_stats, dyn, static_cast<int32_t>(batch));
      }
      bool advanced = false;
      if (crc == KUCD_OK) crc = apply_update(r, hp, batch * ctx->world, dyn, static_cast<int32_t>(batch), no_end, &advanced);
      if (crc == KUCD_OK && !advanced) {
        advance_dyn_kernel<<<1, 1, 0, ctx->stream>>>(dyn, static_cast<int32_t>(batch), no_end);
        ctx->tm.aux_launches++;
      }
      cudaGraph_t g = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
      if (crc != KUCD_OK) {
        if (g != nullptr) cudaGraphDestroy(g);
        return crc;
      }
      if (ce != cudaSuccess) return fail(KUCD_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
      cudaGraphExec_t ge = nullptr;
      const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
      if (ie != cudaSuccess) {
        cudaGraphDestroy(g);
        return fail(KUCD_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
      }
      r->hgraph = g;
      r->hgraph_exec = ge;
      r->hgraph_key = key;
      r->hgraph_kernels = ctx->tm.gemm_launches + ctx->tm.aux_launches - k0;  // recorded while capturing, not run
    }
    set_dyn_kernel<<<1, 1, 0, ctx->stream>>>(dyn, 0, static_cast<int32_t>(batch), r->step_count);
    ctx->tm.aux_launches++;
    CU_TRY(cudaGetLastError());
  }
  int rc = KUCD_OK;
  if (cudaEventRecord(ctx->ev0, ctx->stream) != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "event record failed");
  if (rc == KUCD_OK) rc = copy_in(0);
  for (int64_t i = 0; i < steps && rc == KUCD_OK; ++i) {
    const int slot = static_cast<int>(i & 1);
    const int64_t n = rows_of_step(i);
    if (i + 1 < steps) rc = copy_in(i + 1);
    if (rc != KUCD_OK) break;
    if (cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[slot], 0) != cudaSuccess) {
      rc = fail(KUCD_ERR_CUDA, "stream wait failed");
      break;
    }
    Planes v0 = zero_copy ? slot_planes(slot, n) : r->vin.view(n, r->V, nplanes);
    if (!zero_copy) {
      ingest(ctx->stage_raw[slot].p, packed ? rb * 8 : cols, n, v0, ctx->stream);
      cudaEventRecord(ctx->ev_consumed[slot], ctx->stream);  // the raw staging slot is free again
    }
    v0.n = live;
    if (use_graph && n == batch) {  // the rest of the step is one replay
      if (cudaGraphLaunch(r->hgraph_exec, ctx->stream) != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "graph launch failed");
      r->step_count++;
      ctx->tm.graph_launches++;
      ctx->tm.graph_kernel_launches += r->hgraph_kernels;
      if (host_stats != nullptr) ctx->tm.d2h_bytes += 4;
      continue;
    }
    rc = enqueue_cd(r, v0, n, hp, nullptr, global_row0, r->step_count, nullptr, false);
    r->step_count++;
    if (rc == KUCD_OK && host_stats != nullptr) {
      rc = enqueue_recon(r, v0, n);
      if (rc == KUCD_OK &&
          cudaMemcpyAsync(host_stats + i, r->stats.as<float>() + 1, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        rc = fail(KUCD_ERR_CUDA, "statistic read-back failed");
      ctx->tm.d2h_bytes += 4;
    }
    if (zero_copy) cudaEventRecord(ctx->ev_consumed[slot], ctx->stream);  // the operand planes of this slot are free
    if (rc == KUCD_OK) rc = apply_update(r, hp, n * ctx->world);
  }
  cudaEventRecord(ctx->ev1, ctx->stream);
  if (rc == KUCD_OK) rc = gather_master(r);
  cudaStreamSynchronize(ctx->copy_stream);
  const cudaError_t se = cudaStreamSynchronize(ctx->stream);
  if (rc == KUCD_OK && se != cudaSuccess) rc = fail(KUCD_ERR_CUDA, "fit_host: %s", cudaGetErrorString(se));
  if (host_stats != nullptr && rc == KUCD_OK) memcpy(step_recon, host_stats, steps * sizeof(float));
  if (rc == KUCD_OK && stats != nullptr) {
    stats->steps = steps;
    stats->rows = N;
    cudaEventElapsedTime(&stats->device_ms, ctx->ev0, ctx->ev1);
    if (step_recon != nullptr) stats->last_recon_err = step_recon[steps - 1];
  }
  return rc;
}

int kucd_rbm_fit_epoch(kucd_rbm* r, kucd_dataset* ds, int64_t batch, const kucd_hparams* hp, int64_t global_row0,
                       kucd_epoch_stats* stats) {
  return fit_range_impl(r, ds, batch, hp, global_row0, 0, -1, stats);
}

int kucd_rbm_fit_range(kucd_rbm* r, kucd_dataset* ds, int64_t batch, const kucd_hparams* hp, int64_t global_row0,
                       int64_t step_begin, int64_t step_end, kucd_epoch_stats* stats) {
  return fit_range_impl(r, ds, batch, hp, global_row0, step_begin, step_end, stats);
}

// dbn.py:55 / :73 / :94 on a device-resident data set
static int map_dataset(kucd_rbm* r, bool forward, kucd_dataset* in, kucd_dataset** out) {
  if (r == nullptr || in == nullptr || out == nullptr) return fail(KUCD_ERR_INVALID_ARG, "NULL argument");
  kucd_ctx* ctx = r->ctx;
  *out = nullptr;
  CU_TRY(cudaSetDevice(ctx->device));
  const int64_t K = forward ? r->V : r->H, N = forward ? r->H : r->V;
  if (in->dim != K)
    return fail(KUCD_ERR_SHAPE_MISMATCH, "data set has %lld columns, expected %lld", (long long)in->dim, (long long)K);
  kucd_dataset* ds = new (std::nothrow) kucd_dataset();
  if (ds == nullptr) return fail(KUCD_ERR_INVALID_ARG, "out of host memory");
  ds->ctx = ctx;
  ds->rows = in->rows;
  ds->dim = N;
  const bool gaussian = r->mode == KUCD_MODE_VISIBLE_GAUSSIAN;
  const int np = forward ? 1 : vis_parts_out(r);
  ds->nparts = np;
  int rc = dataset_planes(ctx, ds->planes, std::max<int64_t>(in->rows, 1), round_up(N, 64), np);
  if (rc == KUCD_OK && in->rows > 0) {
    EpiArgs e;
    e.epi = forward ? (gaussian ? kEpiReluSample : kEpiSample) : (gaussian ? kEpiGaussian : kEpiSample);
    e.out = ds->view();
    e.draw = (1ull << 63) | r->infer_draws++;
    rc = project(r, forward, in->view(), in->rows, e);
  }
  if (rc != KUCD_OK) {
    dataset_planes_release(ctx, ds->planes);
    delete ds;
    return rc;
  }
  *out = ds;
  return KUCD_OK;
}

int kucd_rbm_transform_dataset(kucd_rbm* r, kucd_dataset* in, kucd_dataset** out) {
  return map_dataset(r, true, in, out);
}
int kucd_rbm_inv_transform_dataset(kucd_rbm* r, kucd_dataset* in, kucd_dataset** out) {
  return map_dataset(r, false, in, out);
}

}  // extern "C"
