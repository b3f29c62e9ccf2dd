// The one tensor-core kernel of the contrastive-divergence engine.
//
// Every dense contraction on the CD-k hot path of ku/ebm/rbm.py is an instance
// of   D(M,N) = sum_s  sign_s * A_s(M,K) * B_s(K,N)   followed by an elementwise
// epilogue, with bf16 operands and fp32 accumulation:
//
//   v.W + c   -> sigmoid -> Bernoulli h     (rbm.py:46-47, 120, 124)   A K-major,  B MN-major
//   h.W^T + b -> sigmoid -> Bernoulli v'    (rbm.py:52-53, 121-123)    A K-major,  B K-major
//   v0^T h0 - vk^T hk  = delta W            (rbm.py:125-126)           A MN-major, B MN-major,
//                                                                     two K-segments, 2nd negated
//   sum_j softplus(v.W + c)_j               (rbm.py:73-75)             free-energy epilogue
//
// so W is kept once, row-major (V,H), and no state matrix is ever transposed:
// operand major-ness is expressed in the UMMA shared-memory descriptors.  The
// "segments" are K-concatenations that accumulate into the same TMEM tile: they
// carry the positive/negative phases of delta W (negate-A bit of the
// instruction descriptor) and, in f32x3 mode, the bf16 hi/mid/lo splits of
// fp32 operands.
//
// Structure (one CTA per SM - or one CTA pair per TPC with CG = 2 - persistent over output tiles):
//   warp 0      TMA producer   global -> 128B-swizzled smem ring (STAGES deep)
//   warp 1      MMA issuer     one thread, tcgen05.mma cta_group::1 (M=128) or ::2 (M=256 over the pair), N=BN
//   warps 2..9  epilogue       tcgen05.ld TMEM -> registers -> fused math -> global
// The fp32 accumulator is double-buffered in TMEM (2 x BN columns) so the
// epilogue of tile i overlaps the MMAs of tile i+1.  chain.cuh runs the same pipeline over the
// flattened tile sequence of a whole Gibbs chain.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "params.h"
#include "ptx.cuh"
#include "rng_math.cuh"

namespace kucd {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + 32 * kNumEpiWarps;

// epilogues that transpose their tile through a per-warp shared-memory buffer
__host__ __device__ constexpr bool epi_stages(int epi) { return epi == kEpiRaw || epi == kEpiRawPush16; }

constexpr int kEpiStageBytes = kNumEpiWarps * 4096;  // one 32 x 32 fp32 transpose buffer per epilogue warp

template <int BN, int EXTRA = 0>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kCtrlBytes = 1024 + EXTRA;
  static constexpr int kSmemLimit = 232448 - 1024;  // 227 KB minus alignment slack
  static constexpr int kStagesRaw = (kSmemLimit - kCtrlBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kCtrlBytes + 1024;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn, bool negate_a) {
  return (1u << 4)                       // D format: f32
         | (1u << 7)                     // A format: bf16
         | (1u << 10)                    // B format: bf16
         | ((negate_a ? 1u : 0u) << 13)  // -A
         | ((a_mn ? 1u : 0u) << 15)      // A major: 0 = K, 1 = MN
         | ((b_mn ? 1u : 0u) << 16)      // B major
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// lane L ends up holding sum over the warp's 32 lanes of v[L] (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], uint32_t lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Box-Muller on two lattice uniforms (u1 shifted off zero).
__device__ __forceinline__ void box_muller(uint32_t b0, uint32_t b1, float& n0, float& n1) {
  const float u1 = (static_cast<float>(b0 >> 9) + 0.5f) * 1.1920928955078125e-07f;
  const float u2 = u01_from_bits(b1);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// where the raw (fp32) epilogue writes output row `row`
template <typename P>
__device__ __forceinline__ float* raw_row_ptr(const P& p, int row) {
  return p.out_f32 + static_cast<int64_t>(row) * p.ld_f32;
}
template <typename P>
__device__ __forceinline__ bool pushes_rows(const P&) {
  return false;
}
template <>
__device__ __forceinline__ bool pushes_rows<GemmParams>(const GemmParams& p) {
  return p.push_rows > 0;
}
template <>
__device__ __forceinline__ float* raw_row_ptr<GemmParams>(const GemmParams& p, int row) {
  if (p.push_rows > 0) {
    const int o = row / p.push_rows;
    return p.push_base[o] + static_cast<int64_t>(row - o * p.push_rows) * p.ld_f32;
  }
  return p.out_f32 + static_cast<int64_t>(row) * p.ld_f32;
}

// kEpiRawPush16: 64 columns (two tensor-memory chunks) of this warp's 32 rows, rounded to bf16 and stored into the
// owner's slot.  A thread owns one row; the 32 x 64 bf16 block is transposed through shared memory (32 rows of eight
// 16-byte units, unit index XOR row so that neither pass has bank conflicts) and leaves as one 128-byte store per row
// per eight lanes - the same NVLink write size as the fp32 path, for half the bytes.
__device__ __forceinline__ void push16_chunk(const GemmParams& p, const uint32_t (&a0)[32], const uint32_t (&a1)[32],
                                             int row, int col0, uint32_t lane, float* stage) {
  if (col0 >= p.N) return;  // warp-uniform
  uint4* st = reinterpret_cast<uint4*>(stage);
  auto unit = [&](const uint32_t (&a)[32], int q, int c0) {  // columns c0 + 8 q .. + 8 of this row, zero beyond N
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c0 + 8 * q + 2 * e;
      const float lo = c < p.N ? __uint_as_float(a[8 * q + 2 * e]) : 0.f;
      const float hi = c + 1 < p.N ? __uint_as_float(a[8 * q + 2 * e + 1]) : 0.f;
      w[e] = pack_bf16x2(lo, hi);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  };
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    st[lane * 8 + (static_cast<uint32_t>(q) ^ (lane & 7u))] = unit(a0, q, col0);
    st[lane * 8 + (static_cast<uint32_t>(q + 4) ^ (lane & 7u))] = unit(a1, q, col0 + 32);
  }
  __syncwarp();
  const int row_base = row - static_cast<int>(lane);
  const int u = static_cast<int>(lane & 7u);
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // eight lanes per row, four rows per instruction
    const int r = 4 * i + static_cast<int>(lane >> 3);
    const uint4 v = st[r * 8 + (u ^ (r & 7))];
    const int grow = row_base + r;
    if (grow < p.M && col0 + 8 * u < p.N) {
      const int o = grow / p.push_rows;
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.push_base[o]) +
                           static_cast<int64_t>(grow - o * p.push_rows) * p.ld_f32 + col0 + 8 * u;
      *reinterpret_cast<uint4*>(dst) = v;
    }
  }
  __syncwarp();
}

// One 32-column chunk of one output row (this thread's TMEM lane).
template <int EPI, typename P>
__device__ __forceinline__ void epilogue_chunk(const P& p, const uint32_t (&acc)[32], int row, int col0,
                                               bool row_ok, uint64_t draw, int64_t row0, uint32_t lane,
                                               float& row_acc, float* stage = nullptr) {
  if (col0 >= p.N) return;  // warp-uniform
  const bool row_st = row < p.M;  // rows in [m_valid, M) are stored as zeros: they are K-rows of the dW contraction

  if constexpr (EPI == kEpiRaw) {
    if (stage != nullptr && pushes_rows(p)) {
      // Rows may live in another GPU's memory.  A thread owns one row, so a direct store instruction would touch
      // 32 different rows with 16 bytes each - over NVLink that is 32 tiny writes.  Transpose the warp's 32 x 32
      // block through shared memory (XOR-swizzled, conflict-free) and store 128 contiguous bytes per 8 lanes.
      float4* st4 = reinterpret_cast<float4*>(stage);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]),
                               __uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3]));
        if (col0 + 4 * q + 1 >= p.N) v.y = 0.f;
        if (col0 + 4 * q + 2 >= p.N) v.z = 0.f;
        if (col0 + 4 * q + 3 >= p.N) v.w = 0.f;
        st4[lane * 8 + (q ^ (lane & 7u))] = v;
      }
      __syncwarp();
      const int row_base = row - static_cast<int>(lane);
      const int c4 = static_cast<int>(lane & 7u);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + static_cast<int>(lane >> 3);
        const float4 v = st4[r * 8 + (c4 ^ (r & 7))];
        const int grow = row_base + r;
        if (grow < p.M && col0 + 4 * c4 < p.N) *reinterpret_cast<float4*>(raw_row_ptr(p, grow) + col0 + 4 * c4) = v;
      }
      __syncwarp();
      return;
    }
    if (row_ok) {
      float* dst = raw_row_ptr(p, row) + col0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (col0 + 4 * q < p.N) {
          float4 v = make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]),
                                 __uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3]));
          if (col0 + 4 * q + 1 >= p.N) v.y = 0.f;
          if (col0 + 4 * q + 2 >= p.N) v.z = 0.f;
          if (col0 + 4 * q + 3 >= p.N) v.w = 0.f;
          *reinterpret_cast<float4*>(dst + 4 * q) = v;
        }
      }
    }
    return;
  }

  // pre-activation x = D + bias (bias is broadcast down the rows)
  float x[32];
  {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(b4 + q);
      x[4 * q + 0] = __uint_as_float(acc[4 * q + 0]) + b.x;
      x[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + b.y;
      x[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + b.z;
      x[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + b.w;
    }
  }

  if constexpr (EPI == kEpiFreeEnergy) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += (col0 + j < p.N) ? softplus_f32(x[j]) : 0.f;
    row_acc += s;
    return;
  }

  if constexpr (EPI == kEpiSample || EPI == kEpiReluSample) {
    // probabilities
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = (EPI == kEpiSample) ? sigmoid_f32(x[j]) : fmaxf(x[j], 0.f);
    if (p.out_f32 != nullptr && row_ok) {
      float* dst = p.out_f32 + static_cast<int64_t>(row) * p.ld_f32 + col0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (col0 + 4 * q < p.N)
          *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    }
    // uniforms and threshold: bit j of `bits` = sampled state of column col0 + j
    uint32_t bits = 0;
    if (p.u_inject != nullptr) {
      if (row_ok) {
        const float* up = p.u_inject + static_cast<int64_t>(row) * p.ld_u + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float u = (col0 + j < p.N) ? __ldg(up + j) : 2.0f;
          bits |= (u < x[j] ? 1u : 0u) << j;
        }
      }
    } else {
      const uint32_t grow = static_cast<uint32_t>(row0 + row);
      const uint32_t k0 = static_cast<uint32_t>(p.seed), k1 = static_cast<uint32_t>(p.seed >> 32);
      const uint32_t d0 = static_cast<uint32_t>(draw), d1 = static_cast<uint32_t>(draw >> 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const Philox4 r = philox4x32_10(static_cast<uint32_t>(((p.col_off + col0) >> 2) + q), grow, d0, d1, k0, k1);
        bits |= (u01_from_bits(r.x) < x[4 * q + 0] ? 1u : 0u) << (4 * q + 0);
        bits |= (u01_from_bits(r.y) < x[4 * q + 1] ? 1u : 0u) << (4 * q + 1);
        bits |= (u01_from_bits(r.z) < x[4 * q + 2] ? 1u : 0u) << (4 * q + 2);
        bits |= (u01_from_bits(r.w) < x[4 * q + 3] ? 1u : 0u) << (4 * q + 3);
      }
      const int ncol = p.N - col0;  // > 0
      if (ncol < 32) bits &= (1u << ncol) - 1u;
      if (!row_ok) bits = 0;
    }
    if (row_st) {
      uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + static_cast<int64_t>(row) * p.ld_bf16 + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (col0 + 8 * q < p.N) {
          const uint32_t b = bits >> (8 * q);
          uint4 v;  // bf16 1.0 = 0x3F80
          v.x = ((b & 1u) ? 0x3F80u : 0u) | ((b & 2u) ? 0x3F800000u : 0u);
          v.y = ((b & 4u) ? 0x3F80u : 0u) | ((b & 8u) ? 0x3F800000u : 0u);
          v.z = ((b & 16u) ? 0x3F80u : 0u) | ((b & 32u) ? 0x3F800000u : 0u);
          v.w = ((b & 64u) ? 0x3F80u : 0u) | ((b & 128u) ? 0x3F800000u : 0u);
          dst[q] = v;
        }
      }
    }
    if (p.colsum != nullptr) {
      // column sums of the 0/1 states over this warp's 32 rows: one ballot per column
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t b = __ballot_sync(0xffffffffu, (bits >> j) & 1u);
        if (lane == static_cast<uint32_t>(j)) mine = b;
      }
      const int cnt = __popc(mine);
      if (cnt != 0 && col0 + static_cast<int>(lane) < p.N)
        atomicAdd(p.colsum + col0 + lane, p.colsum_sign * static_cast<float>(cnt));
    }
    return;
  }

  if constexpr (EPI == kEpiProb || EPI == kEpiGaussian) {
    if constexpr (EPI == kEpiProb) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = sigmoid_f32(x[j]);
    } else {
      const uint32_t grow = static_cast<uint32_t>(row0 + row);
      const uint32_t k0 = static_cast<uint32_t>(p.seed), k1 = static_cast<uint32_t>(p.seed >> 32);
      const uint32_t d0 = static_cast<uint32_t>(draw), d1 = static_cast<uint32_t>(draw >> 32);
      if (p.u_inject != nullptr) {  // injected standard normals
        if (row_ok) {
          const float* up = p.u_inject + static_cast<int64_t>(row) * p.ld_u + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += (col0 + j < p.N) ? __ldg(up + j) : 0.f;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const Philox4 r = philox4x32_10(static_cast<uint32_t>(((p.col_off + col0) >> 2) + q), grow, d0, d1, k0, k1);
          float n0, n1, n2, n3;
          box_muller(r.x, r.y, n0, n1);
          box_muller(r.z, r.w, n2, n3);
          x[4 * q + 0] += n0;
          x[4 * q + 1] += n1;
          x[4 * q + 2] += n2;
          x[4 * q + 3] += n3;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j >= p.N || !row_ok) x[j] = 0.f;
    if (row_st) {
      if (p.out_f32 != nullptr && row_ok) {
        float* dst = p.out_f32 + static_cast<int64_t>(row) * p.ld_f32 + col0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (col0 + 4 * q < p.N)
            *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
      }
      // bf16 hi (+ optional mid, lo) parts: x = hi + mid + lo to ~2^-24
      float r1[32];
      {
        uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + static_cast<int64_t>(row) * p.ld_bf16 + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = x[8 * q + 2 * e], b = x[8 * q + 2 * e + 1];
            const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
            r1[8 * q + 2 * e] = a - __bfloat162float(ha);
            r1[8 * q + 2 * e + 1] = b - __bfloat162float(hb);
            w[e] = static_cast<uint32_t>(__bfloat16_as_ushort(ha)) |
                   (static_cast<uint32_t>(__bfloat16_as_ushort(hb)) << 16);
          }
          if (col0 + 8 * q < p.N) dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      if (p.out_mid != nullptr) {
        uint4* dm = reinterpret_cast<uint4*>(p.out_mid + static_cast<int64_t>(row) * p.ld_bf16 + col0);
        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + static_cast<int64_t>(row) * p.ld_bf16 + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t wm[4], wl[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = r1[8 * q + 2 * e], b = r1[8 * q + 2 * e + 1];
            const __nv_bfloat16 ma = __float2bfloat16_rn(a), mb = __float2bfloat16_rn(b);
            const __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ma));
            const __nv_bfloat16 lb = __float2bfloat16_rn(b - __bfloat162float(mb));
            wm[e] = static_cast<uint32_t>(__bfloat16_as_ushort(ma)) |
                    (static_cast<uint32_t>(__bfloat16_as_ushort(mb)) << 16);
            wl[e] = static_cast<uint32_t>(__bfloat16_as_ushort(la)) |
                    (static_cast<uint32_t>(__bfloat16_as_ushort(lb)) << 16);
          }
          if (col0 + 8 * q < p.N) {
            dm[q] = make_uint4(wm[0], wm[1], wm[2], wm[3]);
            dl[q] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
          }
        }
      }
    }
    if (p.colsum != nullptr) {
      // probabilities are bounded by 1 per row; the sum of Gaussian visibles (mean + N(0,1)) stays below the row count
      // for any data of unit scale (beyond it the adds merely stop being exact)
      const float s = stat_grid_round(warp_transpose_reduce(x, lane), p.colsum_rows > 0 ? p.colsum_rows : p.M);
      if (col0 + static_cast<int>(lane) < p.N) atomicAdd(p.colsum + col0 + lane, p.colsum_sign * s);
    }
    return;
  }
}

// CH = 0: the whole contraction of a tile accumulates in tensor memory.
// CH > 0: "precise" mode for fp32-grade (f32x3) runs.  The tensor core truncates when it aligns each
// MMA's result to the running fp32 accumulator, an error that grows with the length and the magnitude of
// the chain (measured ~2e-5 relative at K = 4096).  Here the chain is cut every CH k-blocks: each piece
// accumulates from zero in one of the two TMEM buffers and the epilogue warps sum the pieces in
// registers with IEEE fp32 adds while the next piece is being multiplied.
//
// CG = 2: the CTA pair of one TPC computes a 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA stages
// its own 128 rows of A and HALF of B (BN/2 columns) - the tensor cores of the pair exchange the B halves -
// and accumulates its 128 rows in its own tensor memory.  Per SM and per K = 16 step that is 8 KB read from
// shared memory (plus 8 KB written by TMA) instead of 12 + 12 KB: the single-CTA 128 x 256 kernel was
// shared-memory-bandwidth-bound at 75 % tensor-pipe activity (profiles/r01_c3_proj_ncu.md).  Rank 0 of the
// pair issues every MMA; TMA completions of both CTAs are counted on rank 0's barrier; rank 0's commits are
// multicast to both CTAs; the epilogue warps of both CTAs release the accumulator on rank 0's barrier.
template <int BN, bool A_MN, bool B_MN, int EPI, int CH = 0, int CG = 1>
__global__ void __launch_bounds__(kNumThreads, 1) gemm_bf16_kernel(const __grid_constant__ GemmParams p) {
  static_assert(CG == 1 || CG == 2, "cta_group");
  constexpr int kBNLocal = BN / CG;  // B columns staged by this CTA
  constexpr int kTileM = kBlockM * CG;
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  using Cfg = GemmCfg<kBNLocal, (epi_stages(EPI) ? kEpiStageBytes : 0)>;
  constexpr int kStages = Cfg::kStages;
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;  // a unit = one CTA, or one CTA pair

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ctrl = smem + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = ptx::lane_id();

  const int num_m = (p.M + kTileM - 1) / kTileM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;

  // step state: host-supplied for a direct launch, device-resident under graph replay
  int32_t dyn_row_off = 0;
  int32_t m_valid = p.m_valid;
  uint64_t draw = p.draw;
  int64_t row0 = p.row0;
  if (p.dyn != nullptr) {
    if (p.dyn_rows != 0) row0 = static_cast<int64_t>(p.dyn_rank) * p.dyn->rows_valid + p.dyn_row_base;
    dyn_row_off = static_cast<int32_t>(p.dyn->row_off) + p.dyn_row_base;
    if (p.dyn_rows != 0) {
      const int32_t left = p.dyn->rows_valid - p.dyn_row_base;
      m_valid = left < m_valid ? (left < 0 ? 0 : left) : m_valid;
    }
    draw += p.dyn->step * p.draw_stride;
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) {
      ptx::prefetch_tensormap(&p.tm_a[s]);
      ptx::prefetch_tensormap(&p.tm_b[s]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&tmem_full_bar[s], 1);
        ptx::mbar_init(&tmem_empty_bar[s], kNumEpiWarps * CG);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<CG>(tmem_slot, kTmemCols);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();  // peer barriers must exist before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      auto tma_ld = [](void* dst, const CUtensorMap* tm, uint64_t* bar, int32_t c0, int32_t c1) {
        if constexpr (CG == 2) ptx::tma_load_2d_pair(dst, tm, bar, c0, c1);  // bytes counted on rank 0's barrier
        else ptx::tma_load_2d(dst, tm, bar, c0, c1);
      };
      uint32_t stage = 0, phase = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int m0 = m_blk * kTileM + static_cast<int>(cta_rank) * kBlockM;  // this CTA's rows of A
        const int n0 = n_blk * BN + static_cast<int>(cta_rank) * kBNLocal;     // this CTA's columns of B
        for (int s = 0; s < p.num_seg; ++s) {
          const int a_off = ((p.a_dyn_mask >> s) & 1u) ? dyn_row_off : 0;  // A rows: M if K-major, K if MN-major
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
#ifdef KUCD_PROBE
            if (p.dbg_flags & 1u) {
              if (cta_rank == 0) ptx::mbar_arrive(&full_bar[stage]);
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1u;
              }
              continue;
            }
#endif
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * CG);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            const int k0 = kb * kBlockK;
            if constexpr (!A_MN) {
              tma_ld(sa, &p.tm_a[s], &full_bar[stage], k0, m0 + a_off);  // box {64 k, 128 m}
            } else {
#pragma unroll
              for (int j = 0; j < kBlockM / 64; ++j)  // boxes {64 m, 64 k}
                tma_ld(sa + j * (kBlockK * 128), &p.tm_a[s], &full_bar[stage], m0 + 64 * j, k0 + a_off);
            }
            if constexpr (!B_MN) {
              tma_ld(sb, &p.tm_b[s], &full_bar[stage], k0, n0);  // box {64 k, BN/CG n}
            } else {
#pragma unroll
              for (int j = 0; j < kBNLocal / 64; ++j)  // boxes {64 n, 64 k}
                tma_ld(sb + j * (kBlockK * 128), &p.tm_b[s], &full_bar[stage], n0 + 64 * j, k0);
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (rank 0 of a pair) =======================
    if (lane == 0 && cta_rank == 0) {
      // K-major, 128B swizzle : rows of 128 B, 8-row groups 1024 B apart (SBO); a K=16 slice is 32 B along the row.
      // MN-major, 128B swizzle: 64-element MN chunks kBlockK*128 B apart (LBO); 8-k groups 1024 B apart (SBO);
      //                         a K=16 slice is two 8-k groups = 2048 B.
#ifdef KUCD_PROBE
      const uint32_t lbo_a = p.dbg_lbo_a ? p.dbg_lbo_a : (A_MN ? kBlockK * 128u : 16u);
      const uint32_t sbo_a = p.dbg_sbo_a ? p.dbg_sbo_a : 1024u;
      const uint32_t adv_a = p.dbg_adv_a ? p.dbg_adv_a : (A_MN ? 2048u : 32u);
      const uint32_t lbo_b = p.dbg_lbo_b ? p.dbg_lbo_b : (B_MN ? kBlockK * 128u : 16u);
      const uint32_t sbo_b = p.dbg_sbo_b ? p.dbg_sbo_b : 1024u;
      const uint32_t adv_b = p.dbg_adv_b ? p.dbg_adv_b : (B_MN ? 2048u : 32u);
#else
      constexpr uint32_t lbo_a = A_MN ? kBlockK * 128u : 16u, sbo_a = 1024u, adv_a = A_MN ? 2048u : 32u;
      constexpr uint32_t lbo_b = B_MN ? kBlockK * 128u : 16u, sbo_b = 1024u, adv_b = B_MN ? 2048u : 32u;
#endif
      constexpr uint32_t idesc_pos = make_idesc(kTileM, BN, A_MN, B_MN, false);
      constexpr uint32_t idesc_neg = make_idesc(kTileM, BN, A_MN, B_MN, true);
      auto commit = [](uint64_t* bar) {
        if constexpr (CG == 2) ptx::mma_commit_pair(bar, 0b11);  // same-offset barrier in both CTAs
        else ptx::mma_commit(bar);
      };

      uint32_t stage = 0, phase = 0;
      uint32_t accn = 0;  // accumulator buffers handed to the epilogue so far
      const int total = p.num_seg * p.kblocks;
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        uint32_t as = 0, d_tmem = 0;
        int it = 0;
        for (int s = 0; s < p.num_seg; ++s) {
          const uint32_t idesc = ((p.neg_mask >> s) & 1u) ? idesc_neg : idesc_pos;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const int in_piece = CH > 0 ? it % (CH > 0 ? CH : 1) : it;
            if (in_piece == 0) {  // next accumulator buffer: wait until the epilogue has drained it
              as = accn & 1u;
              ptx::mbar_wait(&tmem_empty_bar[as], ((accn >> 1) & 1u) ^ 1u);
              ptx::tc_fence_after();
              d_tmem = tmem_base + as * BN;
            }
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t sb = sa + Cfg::kABytes;
            const uint64_t da = make_smem_desc(sa, lbo_a, sbo_a);
            const uint64_t db = make_smem_desc(sb, lbo_b, sbo_b);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
#ifdef KUCD_PROBE
              if (p.dbg_flags & 2u) break;
#endif
              ptx::mma_bf16<CG>(d_tmem, da + ((k * adv_a) >> 4), db + ((k * adv_b) >> 4), idesc,
                               (in_piece > 0 || k > 0) ? 1u : 0u);
            }
            commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
            ++it;
            if (CH > 0 ? (it % (CH > 0 ? CH : 1) == 0 || it == total) : it == total) {
              commit(&tmem_full_bar[as]);  // this piece of the accumulation is complete
              ++accn;
            }
          }
        }
      }
    }
  } else {
    // ======================= epilogue =======================
    const uint32_t ew = warp - 2;
    const uint32_t quarter = warp & 3u;  // TMEM lane quarter this warp may access
    const uint32_t half = ew >> 2;       // which half of the tile's columns
    constexpr int kColsPerWarp = BN / 2;
    float* epi_stage = epi_stages(EPI) ? reinterpret_cast<float*>(ctrl + 1024 + ew * 4096) : nullptr;
    uint32_t accn = 0;
    const int total = p.num_seg * p.kblocks;
    const int pieces = CH > 0 ? (total + (CH > 0 ? CH : 1) - 1) / (CH > 0 ? CH : 1) : 1;
    auto release = [&](uint32_t as) {  // this warp has drained accumulator buffer `as`
      if constexpr (CG == 2) ptx::mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      else ptx::mbar_arrive(&tmem_empty_bar[as]);
    };
    for (int tile = unit; tile < num_tiles; tile += num_units) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row = m_blk * kTileM + static_cast<int>(cta_rank) * kBlockM + quarter * 32 + lane;
      const bool row_ok = row < m_valid;
      float row_acc = 0.f;
      if constexpr (EPI == kEpiRawPush16) {
        static_assert(CH == 0 && kColsPerWarp % 64 == 0, "bf16 partial sums leave as 64-column (128-byte) row pieces");
        const uint32_t as = accn & 1u;
        ptx::mbar_wait(&tmem_full_bar[as], (accn >> 1) & 1u);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp; c += 64) {
          const int coff = half * kColsPerWarp + c;
          uint32_t a0[32], a1[32];
          ptx::tmem_ld_32x32(tmem_base + ((quarter * 32u) << 16) + as * BN + coff, a0);
          ptx::tmem_ld_32x32(tmem_base + ((quarter * 32u) << 16) + as * BN + coff + 32, a1);
          ptx::tmem_ld_wait();
          push16_chunk(p, a0, a1, row, n_blk * BN + coff, lane, epi_stage);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) release(as);
        ++accn;
      } else if constexpr (CH == 0) {
        const uint32_t as = accn & 1u;
        ptx::mbar_wait(&tmem_full_bar[as], (accn >> 1) & 1u);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp; c += 32) {
          const int coff = half * kColsPerWarp + c;
          uint32_t acc[32];
          ptx::tmem_ld_32x32(tmem_base + ((quarter * 32u) << 16) + as * BN + coff, acc);
          ptx::tmem_ld_wait();
          epilogue_chunk<EPI>(p, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc, epi_stage);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) release(as);
        ++accn;
      } else {
        float sum[kColsPerWarp];
#pragma unroll
        for (int j = 0; j < kColsPerWarp; ++j) sum[j] = 0.f;
        for (int piece = 0; piece < pieces; ++piece) {
          const uint32_t as = accn & 1u;
          ptx::mbar_wait(&tmem_full_bar[as], (accn >> 1) & 1u);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < kColsPerWarp; c += 32) {
            uint32_t acc[32];
            ptx::tmem_ld_32x32(tmem_base + ((quarter * 32u) << 16) + as * BN + half * kColsPerWarp + c, acc);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c + j] += __uint_as_float(acc[j]);
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) release(as);
          ++accn;
        }
#pragma unroll
        for (int c = 0; c < kColsPerWarp; c += 32) {
          uint32_t acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(sum[c + j]);
          epilogue_chunk<EPI>(p, acc, row, n_blk * BN + half * kColsPerWarp + c, row_ok, draw, row0, lane, row_acc, epi_stage);
        }
      }
      if constexpr (EPI == kEpiFreeEnergy) {
        if (row_ok) atomicAdd(p.rowsum + row, row_acc);
      }
    }
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();  // no CTA leaves while its peer may still signal it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

}  // namespace kucd
