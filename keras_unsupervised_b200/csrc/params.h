// Plain-data definitions shared by the kernels (gemm.cuh, chain.cuh), the host code (launch.cuh, kucd.cu) and - because
// this header has no device code - the test-only fake CUDA runtime (tools/dryrun), which decodes launch arguments with it.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdint>

namespace kucd {

constexpr int kMaxSeg = 10;

enum EpiMode : int {
  kEpiRaw = 0,     // out_f32 = D
  kEpiSample = 1,  // s = 1[u < sigmoid(D + bias)]  -> out_bf16 (0/1); optional p -> out_f32; optional colsum
  kEpiProb = 2,    // p = sigmoid(D + bias)         -> out_bf16 (+ mid/lo splits), optional out_f32, colsum
  kEpiFreeEnergy = 3,  // rowsum[m] += sum_n softplus(D + bias)
  kEpiReluSample = 4,  // s = 1[u < relu(D + bias)]   (Gaussian-visible mode, rbm.py:58-59)
  kEpiGaussian = 5,    // x = D + bias + N(0,1)       (Gaussian-visible mode, rbm.py:64-66) -> out_bf16 splits + out_f32
  kEpiRawPush16 = 6,   // D rounded to bf16 (RNE) into the owning rank's slot: the fused exchange with bf16 partial sums on
                       // the wire (opt-in, KUCD_WIRE_BF16=1; push_rows > 0 and BN >= 128 required)
};

// Per-step quantities that live in device memory so that a captured CUDA graph of one CD step can be
// replayed for every minibatch of an epoch: the kernels read them, a one-thread kernel advances them.
struct StepDyn {
  int64_t row_off;     // first data-set row of the current minibatch
  int32_t rows_valid;  // rows of the current minibatch (< batch on the remainder step, rbm.py:211)
  int32_t pad;         // chunked streaming: index of the first minibatch of the current chunk (log_stat_kernel); else 0
  uint64_t step;       // minibatch counter: offsets the Philox draw id
};

struct alignas(64) GemmParams {
  CUtensorMap tm_a[kMaxSeg];
  CUtensorMap tm_b[kMaxSeg];
  int32_t num_seg;
  uint32_t neg_mask;  // bit s: segment s enters with a minus sign
  int32_t M, N;
  int32_t kblocks;  // ceil(K / 64) per segment
  int32_t colsum_rows;  // rows that add into `colsum` over ALL launches sharing it (0: M) - sets the grid of stat_grid_round
  // ---- epilogue ----
  const float* bias;       // readable up to ceil(N/BN)*BN entries
  __nv_bfloat16* out_bf16;  // (M, ld_bf16), ld multiple of 8, >= round_up(N, 8)
  __nv_bfloat16* out_mid;   // optional bf16 split parts of a real-valued output (f32x3 mode)
  __nv_bfloat16* out_lo;
  int64_t ld_bf16;
  float* out_f32;  // (M, ld_f32), ld multiple of 4
  int64_t ld_f32;
  const float* u_inject;  // optional injected uniforms (M, ld_u): parity mode
  int64_t ld_u;
  float* colsum;  // optional (N): += colsum_sign * column sums of the stored output
  float colsum_sign;
  int32_t dyn_rows;  // != 0: the valid row count is min(m_valid, dyn->rows_valid) (minibatch-row outputs)
  int32_t dyn_rank;  // data-parallel rank: under graph replay row0 = dyn_rank * dyn->rows_valid, so that the
  int32_t dyn_row_base;  // remainder minibatch is keyed by global row exactly like the full ones;
                         // dyn_row_base = first minibatch row of this launch (second chain of a split minibatch)
  float* rowsum;  // free energy accumulator (M)
  uint64_t seed;
  uint64_t draw;  // draw id: distinct for every sampling launch
  int64_t row0;   // global row index of local row 0 (data-parallel shards sample identically)
  int32_t m_valid;  // rows < m_valid carry data; rows in [m_valid, M) are stored as zeros
  uint32_t a_dyn_mask;  // bit s: the row coordinate of segment s's A operand is offset by dyn->row_off
  const StepDyn* dyn;   // optional device-resident step state (graph replay)
  uint64_t draw_stride;  // draw += dyn->step * draw_stride
  // ---- fused reduce-scatter of dW over NVLink (data-parallel ranks, raw epilogue only) ----
  // Output row r belongs to rank o = r / push_rows; this rank's contribution to it is stored into
  // push_base[o] (rank o's slot for this rank, peer-mapped memory; push_base[me] is local), row r - o * push_rows.
  // With kEpiRawPush16 the slots hold bf16 rows of the same pitch (ld_f32 elements) and push_base[] carries
  // __nv_bfloat16 pointers (cast): half the bytes cross NVLink, the owner still sums the ranks' parts in fp32.
  float* push_base[8];
  int32_t push_rows;  // 0: off, rows are written to out_f32
  int32_t col_off;    // index of output column 0 among the layer's units: a launch that computes a SLICE of the units
                      // (unit-sharded exchange) draws what the whole-layer launch draws (Philox is keyed by unit index)
#ifdef KUCD_PROBE
  // ---- only in the bring-up probe's build (tools/probe_gemm.cu defines KUCD_PROBE): descriptor overrides (0 = default)
  // and switches that skip work.  libkucd.so is compiled without them: no field of the product's parameter block can
  // make a timed kernel skip its loads or its MMAs.
  uint32_t dbg_lbo_a, dbg_sbo_a, dbg_adv_a, dbg_lbo_b, dbg_sbo_b, dbg_adv_b;
  uint32_t dbg_flags;  // 1 = producer signals "full" without loading (MMA pacing alone),
                       // 2 = MMA thread commits without multiplying (TMA pacing alone)
#endif
};

constexpr int kMaxChainStages = 66;  // 2k+2 projections with k <= 31, + the dW contraction
constexpr int kMaxChainKinds = 9;
constexpr int kChainMaps = 16;  // 5 state operands, 3 views of W, 4 operands of dW, 4 more term planes of W (fp32-grade)

// The few distinct projections a chain is made of (member names shared with GemmParams: epilogue_chunk reads them).
struct ChainKind {
  const float* bias;
  __nv_bfloat16* out_bf16;
  __nv_bfloat16* out_mid;
  __nv_bfloat16* out_lo;
  int64_t ld_bf16;
  float* out_f32;
  int64_t ld_f32;
  const float* u_inject;
  int64_t ld_u;
  float* colsum;
  float colsum_sign;
  int32_t epi;  // kEpiSample / kEpiProb / kEpiRaw; with GAUSS: kEpiReluSample / kEpiGaussian / kEpiProb / kEpiRaw
  float* rowsum;
  int32_t M, N;
  uint64_t seed;
  int32_t kblocks;
  int32_t b_mn;   // 1: W read as (K,N) (v.W), 0: W read as (N,K) (h.W^T)
  int32_t map_a, map_b;
  int32_t a_dyn;  // A is the resident data set: rows offset by dyn->row_off
  int32_t num_n;
  int32_t num_m;       // row blocks of 256 output rows
  int32_t batch_rows;  // 1: output rows are minibatch rows (valid-row masking, global-row draws)
  // the dW contraction (rbm.py:125-126) as the chain's last stage: both operands MN-major (contraction over the
  // minibatch rows), two K-segments - v0^T h0, then vk^T hk with the negate-A bit
  int32_t a_mn;
  int32_t nseg;
  int32_t map_a2, map_b2;
  int32_t map_b3;      // float32-grade chain kernel: third term plane of W (map_b, map_b2, map_b3 = smallest term first)
  int32_t colsum_rows;  // as GemmParams::colsum_rows
  int32_t dep2;        // stage that must be complete in ALL its row blocks before segment 1 is loaded
  int32_t col_off;     // as GemmParams::col_off (always 0: chain launches cover whole layers)
};

struct ChainStageRef {
  int16_t kind;
  int16_t dep;     // stage whose row block must be complete before this stage reads it (-1: none); for an
                   // a_mn stage: ALL row blocks of that stage, before segment 0
  uint32_t phase;  // Philox draw id offset inside the step
};

struct alignas(64) ChainParams {
  CUtensorMap maps[kChainMaps];
  ChainKind kinds[kMaxChainKinds];
  ChainStageRef stages[kMaxChainStages];
  int32_t num_stages;
  int32_t M;        // minibatch rows (buffer capacity)
  int32_t m_valid;  // rows that carry data
  int32_t total_tiles;
  uint32_t* done;   // [num_stages][done_stride] tiles-finished counters, zeroed before the launch
  int32_t done_stride;
  int32_t num_m_batch;  // row blocks of the minibatch
  uint64_t draw;
  uint64_t draw_stride;
  int64_t row0;
  const StepDyn* dyn;
  int32_t dyn_rank;
  int32_t pad;
};

}  // namespace kucd
