// The HBM-bound kernels around the contractions of gemm.cuh: operand ingest (any caller dtype ->
// bf16 term planes), state export, the fused parameter update, column sums for the bias statistics,
// the v.b term of the free energy, the score reduction and the step-state bookkeeping used under
// CUDA-graph replay.  All of them are plain coalesced, vectorised streaming kernels sized in
// multiples of the SM count; none is worth a tensor core.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "gemm.cuh"

namespace kucd {

// x = hi + mid + lo with each term a bf16 (RNE): 24 mantissa bits in three 8-bit pieces.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  mid = __float2bfloat16_rn(r1);
  lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float load_as_float<uint8_t>(const uint8_t* p) {
  return static_cast<float>(__ldg(p));
}

// Grid-stride walk over (row, item-in-row) pairs without a division per element: the position of item index g is
// computed once, then advanced by a constant stride (64-bit divisions per element made the byte-sized kernels
// instruction-bound).
struct RowWalker {
  // 32-bit state (the launches cover at most 2^22 rows of at most 2^21 items): five registers, so the kernels keep the
  // occupancy that hides their load latency
  int32_t r, rem;          // current row, item inside the row
  int32_t dr, drem, per;   // per-step advance, items per row
  __device__ __forceinline__ RowWalker(int64_t g, int64_t step, int64_t per_row) {
    per = static_cast<int32_t>(per_row);
    const int64_t r0 = g / per_row, d0 = step / per_row;
    r = static_cast<int32_t>(r0 < 0x7fffffff ? r0 : 0x7fffffff);  // beyond the matrix anyway: the caller tests g < total
    rem = static_cast<int32_t>(g - r0 * per_row);
    dr = static_cast<int32_t>(d0 < 0x3fffffff ? d0 : 0x3fffffff);
    drem = static_cast<int32_t>(step - d0 * per_row);
  }
  __device__ __forceinline__ void advance() {
    r += dr;
    rem += drem;
    if (rem >= per) {
      rem -= per;
      ++r;
    }
  }
};

// Caller matrix (rows, cols), row stride `src_ld` elements -> `nparts` bf16 planes (rows, ld), columns
// [cols, ld) zeroed.  *inexact is raised when some value is not a bf16 number, which tells the host
// whether the first plane alone carries the data exactly (binary data always does).
template <typename T>
__global__ void __launch_bounds__(256, 8) ingest_kernel(const T* __restrict__ src, int64_t src_ld, int64_t rows, int64_t cols,
                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid,
                              __nv_bfloat16* __restrict__ lo, int64_t ld, int nparts, int* inexact) {
  const int64_t groups_per_row = ld / 8;
  const int64_t total = rows * groups_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool full = nparts == 3 || inexact != nullptr;  // otherwise only the leading term is wanted
  bool bad = false;
  int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (RowWalker w(g, stride, groups_per_row); g < total; g += stride, w.advance()) {
    const int64_t r = w.r, c0 = static_cast<int64_t>(w.rem) * 8;
    __align__(16) __nv_bfloat16 h[8], m[8], l[8];
    float x[8];
    const T* row = src + r * src_ld + c0;
    if (sizeof(T) == 1 && c0 + 8 <= cols && (reinterpret_cast<uintptr_t>(row) & 7u) == 0) {
      // uint8 input: the group's eight bytes in one load (eight byte loads per thread made this kernel LSU-bound)
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(row));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[j] = static_cast<float>((q.x >> (8 * j)) & 0xFFu);
        x[4 + j] = static_cast<float>((q.y >> (8 * j)) & 0xFFu);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (c0 + j < cols) ? load_as_float<T>(row + j) : 0.f;
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        split3(x[j], h[j], m[j], l[j]);
        bad |= (__bfloat162float(h[j]) != x[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __float2bfloat16_rn(x[j]);
    }
    *reinterpret_cast<uint4*>(hi + r * ld + c0) = *reinterpret_cast<const uint4*>(h);
    if (nparts == 3) {
      *reinterpret_cast<uint4*>(mid + r * ld + c0) = *reinterpret_cast<const uint4*>(m);
      *reinterpret_cast<uint4*>(lo + r * ld + c0) = *reinterpret_cast<const uint4*>(l);
    }
  }
  if (inexact != nullptr && bad) atomicOr(inexact, 1);
}

// Bit-packed 0/1 matrices: column j of a row is bit (j % 8) of byte (j / 8) - numpy.packbits(..., bitorder="little").
// The two conversions below are shared by the kernels and by tools/data_path_host.cu (host-side check against the oracle).
struct Bf16x8 {
  uint32_t w[4];  // eight bf16 values, element 2i in the low half of w[i]
};

// one byte -> eight bf16 values in {0, 1}, only the first `valid` columns kept (bf16 1.0 = 0x3F80)
__host__ __device__ __forceinline__ Bf16x8 bits_to_bf16x8(uint32_t b, int valid) {
  if (valid < 8) b &= (1u << (valid > 0 ? valid : 0)) - 1u;
  Bf16x8 v;
  v.w[0] = ((b & 1u) ? 0x3F80u : 0u) | ((b & 2u) ? 0x3F800000u : 0u);
  v.w[1] = ((b & 4u) ? 0x3F80u : 0u) | ((b & 8u) ? 0x3F800000u : 0u);
  v.w[2] = ((b & 16u) ? 0x3F80u : 0u) | ((b & 32u) ? 0x3F800000u : 0u);
  v.w[3] = ((b & 64u) ? 0x3F80u : 0u) | ((b & 128u) ? 0x3F800000u : 0u);
  return v;
}

// eight bf16 values -> one byte: bit j = (value j != +-0), only the first `valid` columns kept
__host__ __device__ __forceinline__ uint32_t bf16x8_to_bits(const Bf16x8& v, int valid) {
  uint32_t b = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    b |= ((v.w[j] & 0x7FFFu) != 0u ? 1u : 0u) << (2 * j);
    b |= ((v.w[j] & 0x7FFF0000u) != 0u ? 1u : 0u) << (2 * j + 1);
  }
  if (valid < 8) b &= (1u << (valid > 0 ? valid : 0)) - 1u;
  return b;
}

// Packed rows -> bf16 planes.  One source byte (8 columns = one 16-byte store) per thread and iteration, four
// iterations' loads in flight; 1/8 B read + 2 B written per unit, against 4 + 2 B for float32 input: a binary data
// set crosses PCIe and HBM 32x smaller than as float32.
// [skip_lo, skip_hi): columns (multiples of 8) that are left untouched - the unit-sharded exchange expands the peers'
// slices of a gathered state matrix around the slice this rank computed itself, which is already in place.
__global__ void ingest_bits_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int64_t rows, int64_t cols,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid,
                                   __nv_bfloat16* __restrict__ lo, int64_t ld, int nparts, int64_t skip_lo = 0,
                                   int64_t skip_hi = 0) {
  const int64_t groups_per_row = ld / 8;
  const int64_t total = rows * groups_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t g0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  RowWalker w[4] = {RowWalker(g0, 4 * stride, groups_per_row), RowWalker(g0 + stride, 4 * stride, groups_per_row),
                    RowWalker(g0 + 2 * stride, 4 * stride, groups_per_row),
                    RowWalker(g0 + 3 * stride, 4 * stride, groups_per_row)};
  for (; g0 < total; g0 += 4 * stride) {
    uint32_t b[4];
    int64_t off[4];
    int valid[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t g = g0 + u * stride;
      b[u] = 0u;
      valid[u] = -1;  // -1: no such group
      if (g < total) {
        const int64_t r = w[u].r, c0 = static_cast<int64_t>(w[u].rem) * 8;
        const int64_t left = cols - c0;  // columns of this group that exist
        if (c0 < skip_lo || c0 >= skip_hi) {
          if (left > 0) b[u] = __ldg(src + r * src_pitch + (c0 >> 3));
          valid[u] = left >= 8 ? 8 : (left > 0 ? static_cast<int>(left) : 0);
          off[u] = r * ld + c0;
        }
      }
      w[u].advance();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (valid[u] < 0) continue;
      const Bf16x8 v = bits_to_bf16x8(b[u], valid[u]);
      *reinterpret_cast<uint4*>(hi + off[u]) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
      if (nparts == 3) {
        *reinterpret_cast<uint4*>(mid + off[u]) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(lo + off[u]) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

// 0/1 state plane -> packed rows; one destination byte per thread and iteration, four 16-byte loads in flight
__global__ void export_bits_kernel(const __nv_bfloat16* __restrict__ hi, int64_t ld, int64_t rows, int64_t cols,
                                   uint8_t* __restrict__ dst, int64_t dst_pitch) {
  const int64_t bytes_per_row = (cols + 7) / 8;
  const int64_t total = rows * bytes_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  RowWalker w[4] = {RowWalker(i0, 4 * stride, bytes_per_row), RowWalker(i0 + stride, 4 * stride, bytes_per_row),
                    RowWalker(i0 + 2 * stride, 4 * stride, bytes_per_row),
                    RowWalker(i0 + 3 * stride, 4 * stride, bytes_per_row)};
  for (; i0 < total; i0 += 4 * stride) {
    uint4 q[4];
    int64_t out[4];
    int valid[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      valid[u] = -1;
      if (i < total) {
        const int64_t r = w[u].r, c0 = static_cast<int64_t>(w[u].rem) * 8;
        q[u] = *reinterpret_cast<const uint4*>(hi + r * ld + c0);  // ld is a multiple of 64: in bounds, aligned
        const int64_t left = cols - c0;
        valid[u] = left >= 8 ? 8 : static_cast<int>(left);
        out[u] = r * dst_pitch + (c0 >> 3);
      }
      w[u].advance();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (valid[u] < 0) continue;
      const Bf16x8 v{{q[u].x, q[u].y, q[u].z, q[u].w}};
      dst[out[u]] = static_cast<uint8_t>(bf16x8_to_bits(v, valid[u]));
    }
  }
}

template <typename T>
__device__ __forceinline__ void store_from_float(T* p, float x);
template <>
__device__ __forceinline__ void store_from_float<float>(float* p, float x) {
  *p = x;
}
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, float x) {
  *p = __float2bfloat16_rn(x);
}
template <>
__device__ __forceinline__ void store_from_float<uint8_t>(uint8_t* p, float x) {
  *p = static_cast<uint8_t>(x);
}

// bf16 planes (summed) -> caller matrix of type T
template <typename T>
__global__ void export_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ mid,
                              const __nv_bfloat16* __restrict__ lo, int64_t ld, int64_t rows, int64_t cols,
                              T* __restrict__ dst, int64_t dst_ld) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    float x = __bfloat162float(hi[r * ld + c]);
    if (mid != nullptr) x += __bfloat162float(mid[r * ld + c]) + __bfloat162float(lo[r * ld + c]);
    store_from_float<T>(dst + r * dst_ld + c, x);
  }
}

// float32 (rows, ld_src) -> caller matrix of type T (rows, cols)
template <typename T>
__global__ void export_f32_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols,
                                  T* __restrict__ dst, int64_t dst_ld) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    store_from_float<T>(dst + r * dst_ld + c, src[r * ld + c]);
  }
}

// normalize = mean under graph replay: the divisor is the CURRENT minibatch's global row count (the remainder
// minibatch is shorter), which only the device knows
__device__ __forceinline__ float step_scale(float scale, const StepDyn* sdyn, float world) {
  if (sdyn == nullptr) return scale;
  const int rows = sdyn->rows_valid > 0 ? sdyn->rows_valid : 1;
  return 1.0f / (static_cast<float>(rows) * world);
}

// Small work folded into the weight-update launch (done by its last block) so that a latency-bound step is one
// launch shorter per item: the two bias updates (rbm.py:129-134) and the step-state advance of graph replay.
struct UpdateTail {
  float* b;
  const float* db;
  float* mb;
  int32_t nb;  // visible bias entries (0: skip)
  float* c;
  const float* dc;
  float* mc;
  int32_t nc;  // hidden bias entries (0: skip)
  StepDyn* dyn;  // advance to the next minibatch (nullptr: skip)
  int32_t adv_batch;
  int64_t adv_total;
};

// The parameter update of rbm.py:127-134 for the weight matrix, generalised with momentum and weight
// decay (both 0 in the reference):
//     g = scale * dW - weight_decay * W ;  m = momentum * m + lr * g ;  W += m
// and, in the same pass, the refresh of the bf16 operand planes the contractions read.  One float4
// of W per thread per iteration: 4 B (dW) + 4 B (W) read, 4 B (W) + 2 B per plane written per weight.
// D16: dW is a bf16 array (the result of the bf16 all-reduce, KUCD_WIRE_BF16=1), widened here.
template <bool D16>
__global__ void update_w_kernel(float* __restrict__ W, const float* __restrict__ dW, float* __restrict__ mom,
                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid,
                                __nv_bfloat16* __restrict__ lo, int64_t n4, float lr, float scale, float momentum,
                                float weight_decay, UpdateTail tail, const StepDyn* sdyn, float world) {
  scale = step_scale(scale, sdyn, world);
  if (blockIdx.x == gridDim.x - 1) {
    for (int i = threadIdx.x; i < tail.nb; i += blockDim.x) {
      float s = lr * scale * tail.db[i];
      if (tail.mb != nullptr) {
        s = momentum * tail.mb[i] + s;
        tail.mb[i] = s;
      }
      tail.b[i] += s;
    }
    for (int i = threadIdx.x; i < tail.nc; i += blockDim.x) {
      float s = lr * scale * tail.dc[i];
      if (tail.mc != nullptr) {
        s = momentum * tail.mc[i] + s;
        tail.mc[i] = s;
      }
      tail.c[i] += s;
    }
    if (threadIdx.x == 0 && tail.dyn != nullptr) {
      // next minibatch: sequential slices, remainder last (rbm.py:163,211,218)
      const int64_t off = tail.dyn->row_off + tail.adv_batch;
      const int64_t left = tail.adv_total - off;
      tail.dyn->row_off = off;
      tail.dyn->rows_valid = static_cast<int32_t>(left < 0 ? 0 : (left < tail.adv_batch ? left : tail.adv_batch));
      tail.dyn->step += 1;
    }
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 w = reinterpret_cast<const float4*>(W)[i];
    float4 d;
    if constexpr (D16) {
      const uint2 h = reinterpret_cast<const uint2*>(dW)[i];
      d = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xFFFF0000u), __uint_as_float(h.y << 16),
                      __uint_as_float(h.y & 0xFFFF0000u));
    } else {
      d = reinterpret_cast<const float4*>(dW)[i];
    }
    float4 s;
    s.x = lr * (scale * d.x - weight_decay * w.x);
    s.y = lr * (scale * d.y - weight_decay * w.y);
    s.z = lr * (scale * d.z - weight_decay * w.z);
    s.w = lr * (scale * d.w - weight_decay * w.w);
    if (mom != nullptr) {
      float4 m = reinterpret_cast<const float4*>(mom)[i];
      m.x = momentum * m.x + s.x;
      m.y = momentum * m.y + s.y;
      m.z = momentum * m.z + s.z;
      m.w = momentum * m.w + s.w;
      reinterpret_cast<float4*>(mom)[i] = m;
      s = m;
    }
    w.x += s.x;
    w.y += s.y;
    w.z += s.z;
    w.w += s.w;
    reinterpret_cast<float4*>(W)[i] = w;
    __nv_bfloat16 h[4], m4[4], l4[4];
    split3(w.x, h[0], m4[0], l4[0]);
    split3(w.y, h[1], m4[1], l4[1]);
    split3(w.z, h[2], m4[2], l4[2]);
    split3(w.w, h[3], m4[3], l4[3]);
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<const uint2*>(h);
    if (mid != nullptr) {
      reinterpret_cast<uint2*>(mid)[i] = *reinterpret_cast<const uint2*>(m4);
      reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<const uint2*>(l4);
    }
  }
}

// fp32 master -> bf16 planes only (after set_params)
__global__ void refresh_planes_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo, int64_t n4) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 w = reinterpret_cast<const float4*>(W)[i];
    __nv_bfloat16 h[4], m4[4], l4[4];
    split3(w.x, h[0], m4[0], l4[0]);
    split3(w.y, h[1], m4[1], l4[1]);
    split3(w.z, h[2], m4[2], l4[2]);
    split3(w.w, h[3], m4[3], l4[3]);
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<const uint2*>(h);
    if (mid != nullptr) {
      reinterpret_cast<uint2*>(mid)[i] = *reinterpret_cast<const uint2*>(m4);
      reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<const uint2*>(l4);
    }
  }
}

// bias update (rbm.py:129-134): x += lr * scale * d (+ momentum), n entries
__global__ void update_bias_kernel(float* __restrict__ x, const float* __restrict__ d, float* __restrict__ mom,
                                   int64_t n, float lr, float scale, float momentum, const StepDyn* sdyn, float world) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  scale = step_scale(scale, sdyn, world);
  float s = lr * scale * d[i];
  if (mom != nullptr) {
    s = momentum * mom[i] + s;
    mom[i] = s;
  }
  x[i] += s;
}

// out[c] += sign * sum_r (hi + mid + lo)[row_off + r, c], r < rows_valid.  One bf16 pair per thread,
// 64 rows per block.y slab, one atomic per column per slab.
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ mid,
                              const __nv_bfloat16* __restrict__ lo, int64_t ld, int64_t row_off, int32_t rows,
                              int32_t cols, const StepDyn* dyn, float sign, float* __restrict__ out) {
  const int grid_rows = rows;  // the minibatch size every launch adding into `out` knows: one grid for all of them
  if (dyn != nullptr) {
    row_off += dyn->row_off;
    rows = dyn->rows_valid < rows ? dyn->rows_valid : rows;
  }
  const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= cols) return;
  const int r0 = blockIdx.y * 64;
  const int r1 = r0 + 64 < rows ? r0 + 64 : rows;
  float s0 = 0.f, s1 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int64_t off = (row_off + r) * ld + c;
    float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hi + off));
    if (mid != nullptr) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(mid + off));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(lo + off));
      v.x += a.x + b.x;
      v.y += a.y + b.y;
    }
    s0 += v.x;
    s1 += v.y;
  }
  if (r1 > r0) {
    // real-valued data: slab sums on the grid of stat_grid_round (rng_math.cuh), so that the atomics commute; 0/1 data
    // sums to integers and is unchanged
    atomicAdd(out + c, sign * stat_grid_round(s0, grid_rows));
    if (c + 1 < cols) atomicAdd(out + c + 1, sign * stat_grid_round(s1, grid_rows));
  }
}

// Small minibatches: each block owns 32 column pairs, its 16 warps stride over the rows and combine through shared
// memory, so the column sums are STORED (no memset, no atomics); the same launch clears `zero` (the dc accumulator
// the epilogues add into).  blockDim = (32, 16).
__global__ void colsum_store_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ mid,
                                    const __nv_bfloat16* __restrict__ lo, int64_t ld, int32_t rows, int32_t cols,
                                    int32_t cols_pad, const StepDyn* dyn, float* __restrict__ out,
                                    float* __restrict__ zero, int32_t zero_len) {
  __shared__ float2 part[16][32];
  int64_t row_off = 0;
  const int grid_rows = rows;  // as in colsum_kernel: the epilogues' atomics add to these sums and must find them on their grid
  if (dyn != nullptr) {
    row_off = dyn->row_off;
    rows = dyn->rows_valid < rows ? dyn->rows_valid : rows;
  }
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = blockIdx.x * 512 + tid; i < zero_len; i += gridDim.x * 512) zero[i] = 0.f;
  const int c = 2 * (blockIdx.x * 32 + threadIdx.x);
  float s0 = 0.f, s1 = 0.f;
  if (c < cols) {
    // eight independent loads in flight per thread: the rows of a minibatch are far apart in HBM
    for (int r = threadIdx.y; r < rows; r += 128) {
      __nv_bfloat162 h[8], m[8], l[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int rr = r + 16 * u;
        const int64_t off = (row_off + (rr < rows ? rr : r)) * ld + c;
        h[u] = *reinterpret_cast<const __nv_bfloat162*>(hi + off);
        if (mid != nullptr) {
          m[u] = *reinterpret_cast<const __nv_bfloat162*>(mid + off);
          l[u] = *reinterpret_cast<const __nv_bfloat162*>(lo + off);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r + 16 * u < rows) {
          float2 v = __bfloat1622float2(h[u]);
          if (mid != nullptr) {
            const float2 a = __bfloat1622float2(m[u]), b = __bfloat1622float2(l[u]);
            v.x += a.x + b.x;
            v.y += a.y + b.y;
          }
          s0 += v.x;
          s1 += v.y;
        }
      }
    }
  }
  part[threadIdx.y][threadIdx.x] = make_float2(s0, s1);
  __syncthreads();
  if (threadIdx.y == 0 && c < cols_pad) {
    float2 t = part[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      t.x += part[k][threadIdx.x].x;
      t.y += part[k][threadIdx.x].y;
    }
    out[c] = stat_grid_round(t.x, grid_rows);
    out[c + 1] = (c + 1 < cols) ? stat_grid_round(t.y, grid_rows) : 0.f;
  }
}

// Free energy tail (rbm.py:73-75):  F[r] = -( v[r,:].b + softplus_sum[r] ).  One warp per row.
__global__ void free_energy_finish_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ mid,
                                          const __nv_bfloat16* __restrict__ lo, int64_t ld, int64_t row_off,
                                          int32_t rows, int32_t cols, const StepDyn* dyn,
                                          const float* __restrict__ b, const float* __restrict__ softplus_sum,
                                          float* __restrict__ out) {
  if (dyn != nullptr) {
    row_off += dyn->row_off;
    rows = dyn->rows_valid < rows ? dyn->rows_valid : rows;
  }
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int64_t base = (row_off + warp) * ld;
  float s = 0.f;
  for (int c = 2 * lane; c < cols; c += 64) {
    float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hi + base + c));
    if (mid != nullptr) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(mid + base + c));
      const float2 l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(lo + base + c));
      v.x += a.x + l.x;
      v.y += a.y + l.y;
    }
    s += v.x * b[c];
    if (c + 1 < cols) s += v.y * b[c + 1];
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[warp] = -(s + softplus_sum[warp]);
}

// stats[0] = mean |fe - fe_p| (rbm.py:233) ; stats[2] = mean fe.  Single block: deterministic.
__global__ void score_kernel(const float* __restrict__ fe, const float* __restrict__ fe_p, int32_t rows,
                             const StepDyn* dyn, float* __restrict__ stats) {
  if (dyn != nullptr) rows = dyn->rows_valid < rows ? dyn->rows_valid : rows;
  __shared__ float sh_a[32], sh_b[32];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    a += fabsf(fe[i] - fe_p[i]);
    b += fe[i];
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh_a[threadIdx.x >> 5] = a;
    sh_b[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? sh_a[threadIdx.x] : 0.f;
    b = threadIdx.x < (blockDim.x >> 5) ? sh_b[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) {
      stats[0] = rows > 0 ? a / rows : 0.f;
      stats[2] = rows > 0 ? b / rows : 0.f;
      stats[3] = static_cast<float>(rows);
    }
  }
}

// acc[0] += sum (v0 - vk)^2 over the valid rows (reconstruction error numerator)
__global__ void recon_kernel(const __nv_bfloat16* __restrict__ a_hi, const __nv_bfloat16* __restrict__ a_mid,
                             const __nv_bfloat16* __restrict__ a_lo, int64_t a_ld, int64_t a_row_off,
                             const __nv_bfloat16* __restrict__ b_hi, const __nv_bfloat16* __restrict__ b_mid,
                             const __nv_bfloat16* __restrict__ b_lo, int64_t b_ld, int32_t rows, int32_t cols,
                             const StepDyn* dyn, float* __restrict__ acc) {
  if (dyn != nullptr) {
    a_row_off += dyn->row_off;
    rows = dyn->rows_valid < rows ? dyn->rows_valid : rows;
  }
  const int64_t total = static_cast<int64_t>(rows) * cols;
  float s = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    float a = __bfloat162float(a_hi[(a_row_off + r) * a_ld + c]);
    if (a_mid != nullptr)
      a += __bfloat162float(a_mid[(a_row_off + r) * a_ld + c]) + __bfloat162float(a_lo[(a_row_off + r) * a_ld + c]);
    float b = __bfloat162float(b_hi[r * b_ld + c]);
    if (b_mid != nullptr) b += __bfloat162float(b_mid[r * b_ld + c]) + __bfloat162float(b_lo[r * b_ld + c]);
    s += (a - b) * (a - b);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(acc, s);
}

// Latency-bound minibatches: the whole reconstruction statistic in ONE launch - the squared differences, the mean into
// stats[1] and, for a streamed fit under graph replay, the write into the caller-visible page-locked log (log_stat_kernel's
// job) - instead of memset + recon + finish + log.  Single planes only.  Up to kReconBlocks blocks take the rows round-robin
// (16-byte loads: eight units each; columns beyond `cols` are zero in both planes up to the next multiple of eight), leave
// their partial sums in `part` and draw a ticket; the block that draws the last one adds the partials IN INDEX ORDER (the
// statistic is the same in every run), writes the results and puts the ticket back to zero for the next launch.  As one
// block of 1024 threads walking bf16 pairs with a 64-bit division each, this took about as long as a fifth of a C1 step.
constexpr int kReconBlocks = 64;
__global__ void __launch_bounds__(128) recon_small_kernel(const __nv_bfloat16* __restrict__ a, int64_t a_ld,
                                                          const __nv_bfloat16* __restrict__ b, int64_t b_ld, int32_t rows,
                                                          int32_t cols, const StepDyn* a_dyn, float* __restrict__ stats,
                                                          const StepDyn* log_dyn, int32_t log_batch, float* host_log,
                                                          float* __restrict__ part, unsigned int* __restrict__ ticket) {
  __shared__ float warp_sum[4];
  int64_t a_row_off = 0;
  if (a_dyn != nullptr) {
    a_row_off = a_dyn->row_off;
    rows = a_dyn->rows_valid < rows ? a_dyn->rows_valid : rows;
  }
  const int nvec = (cols + 7) / 8;
  float s = 0.f;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const uint4* pa = reinterpret_cast<const uint4*>(a + (a_row_off + r) * a_ld);
    const uint4* pb = reinterpret_cast<const uint4*>(b + static_cast<int64_t>(r) * b_ld);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      const uint4 x = pa[v], y = pb[v];
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // a bf16 is the upper half of the float with the same value
        const float d0 = __uint_as_float(xs[q] << 16) - __uint_as_float(ys[q] << 16);
        const float d1 = __uint_as_float(xs[q] & 0xFFFF0000u) - __uint_as_float(ys[q] & 0xFFFF0000u);
        s += d0 * d0 + d1 * d1;
      }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    part[blockIdx.x] = (warp_sum[0] + warp_sum[1]) + (warp_sum[2] + warp_sum[3]);
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      float t = 0.f;
      for (unsigned int i = 0; i < gridDim.x; ++i) t += __ldcg(part + i);
      const float v = rows > 0 ? t / (static_cast<float>(rows) * cols) : 0.f;
      stats[1] = v;
      if (host_log != nullptr) host_log[log_dyn->pad + log_dyn->row_off / log_batch] = v;
      *ticket = 0u;
    }
  }
}

// stats[1] = acc / (rows * cols)
__global__ void recon_finish_kernel(const float* acc, int32_t rows, int32_t cols, const StepDyn* dyn, float* stats) {
  if (dyn != nullptr) rows = dyn->rows_valid < rows ? dyn->rows_valid : rows;
  stats[1] = rows > 0 ? acc[0] / (static_cast<float>(rows) * cols) : 0.f;
}

// valid rows of src -> dst (persistent chains keep the rows a remainder minibatch did not touch)
__global__ void copy_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t ld,
                                 int32_t rows, const StepDyn* dyn) {
  if (dyn != nullptr) rows = dyn->rows_valid < rows ? dyn->rows_valid : rows;
  const int64_t total = static_cast<int64_t>(rows) * (ld / 8);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

// ---------------------------------------------------------------------------------------------------
// Epoch shuffling (the reference never shuffles, rbm.py:218).  A keyed bijection of [0, n): a
// 6-round balanced Feistel network over the smallest even-width power of two >= n, walked until it lands below n
// (cycle walking keeps it a bijection).  Counter-based like the Philox draws: any row's source is computed from its
// index alone - no sort, no permutation array, nothing to exchange between ranks.
// ---------------------------------------------------------------------------------------------------
struct FeistelKey {
  uint32_t k[6];
  uint32_t half_bits;  // bits per half (width = 2 * half_bits)
  uint32_t pad;
};

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

__host__ __device__ __forceinline__ uint64_t feistel_permute(uint64_t i, uint64_t n, const FeistelKey& key) {
  const uint32_t hb = key.half_bits;
  const uint32_t mask = (hb >= 32u) ? 0xFFFFFFFFu : ((1u << hb) - 1u);
  uint64_t x = i;
  do {
    uint32_t L = static_cast<uint32_t>(x >> hb) & mask, R = static_cast<uint32_t>(x) & mask;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const uint32_t F = fmix32(R ^ key.k[r]) & mask;
      const uint32_t t = L ^ F;
      L = R;
      R = t;
    }
    x = (static_cast<uint64_t>(L) << hb) | R;
  } while (x >= n);
  return x;
}

// The key of epoch `epoch`'s row permutation: six 32-bit round keys from a model-independent Philox stream
// (counter = ("SHUF", j, epoch), key = seed), and the Feistel width for `rows` rows.
inline FeistelKey make_feistel_key(uint64_t seed, uint64_t epoch, int64_t rows) {
  FeistelKey key{};
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t e0 = static_cast<uint32_t>(epoch), e1 = static_cast<uint32_t>(epoch >> 32);
  const Philox4 a = philox4x32_10(0x53485546u, 0u, e0, e1, k0, k1);
  const Philox4 b = philox4x32_10(0x53485546u, 1u, e0, e1, k0, k1);
  key.k[0] = a.x;
  key.k[1] = a.y;
  key.k[2] = a.z;
  key.k[3] = a.w;
  key.k[4] = b.x;
  key.k[5] = b.y;
  int bits = 2;
  while ((int64_t{1} << bits) < rows) ++bits;
  if (bits & 1) ++bits;
  key.half_bits = static_cast<uint32_t>(bits / 2);
  return key;
}

// dst row i = src row perm(i), every live plane; one block per row, 16-byte copies.  2 x rows x ld x 2 B of traffic.
__global__ void permute_rows_kernel(const __nv_bfloat16* __restrict__ s0, const __nv_bfloat16* __restrict__ s1,
                                    const __nv_bfloat16* __restrict__ s2, __nv_bfloat16* __restrict__ d0,
                                    __nv_bfloat16* __restrict__ d1, __nv_bfloat16* __restrict__ d2, int64_t ld,
                                    int64_t rows, FeistelKey key) {
  const int64_t vec = ld / 8;
  for (int64_t i = blockIdx.x; i < rows; i += gridDim.x) {
    const int64_t j = static_cast<int64_t>(feistel_permute(static_cast<uint64_t>(i), static_cast<uint64_t>(rows), key));
    const uint4* a = reinterpret_cast<const uint4*>(s0 + j * ld);
    uint4* b = reinterpret_cast<uint4*>(d0 + i * ld);
    for (int64_t c = threadIdx.x; c < vec; c += blockDim.x) b[c] = a[c];
    if (s1 != nullptr) {
      const uint4* a1 = reinterpret_cast<const uint4*>(s1 + j * ld);
      const uint4* a2 = reinterpret_cast<const uint4*>(s2 + j * ld);
      uint4* b1 = reinterpret_cast<uint4*>(d1 + i * ld);
      uint4* b2 = reinterpret_cast<uint4*>(d2 + i * ld);
      for (int64_t c = threadIdx.x; c < vec; c += blockDim.x) {
        b1[c] = a1[c];
        b2[c] = a2[c];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Data-parallel exchange over peer-mapped memory (NVLink): see kucd.cu, "fused reduction"
// ---------------------------------------------------------------------------------------------------
struct PeerSet {
  float* dw_slot[8];      // rank j's dW slots: slot s (at s * slice_elems) holds rank s's contribution to rank j's rows
  float* bias_slot[8];    // rank j's bias slots: slot s (at s * bias_len) holds rank s's [db | dc]
  uint32_t* flags[8];     // rank j's arrival flags, one per source rank
  __nv_bfloat16* wp[8];   // rank j's bf16 operand plane of W
  uint8_t* bits[8];       // rank j's bit slots of the unit-sharded exchange (two slots of kBitSlotBytes; nullptr: none)
};

constexpr int64_t kBitSlotBytes = int64_t{32} << 20;  // 2^28 units: one gathered 0/1 matrix (rows x units) per slot

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All-to-all arrival barrier between the ranks' streams: everything this rank wrote into peer memory before this
// kernel is visible to a peer once the peer has seen this rank's flag.  One thread per peer.
__global__ void peer_barrier_kernel(uint32_t* epoch, PeerSet ps, int me, int n) {
  // one warp, no shared memory (see pack_push_kernel)
  uint32_t e = 0;
  if (threadIdx.x == 0) e = ++(*epoch);
  e = __shfl_sync(0xffffffffu, e, 0);
  const int j = threadIdx.x;
  if (j >= n) return;
  __threadfence_system();
  st_release_sys(ps.flags[j] + me, e);
  const uint32_t* mine = ps.flags[me] + j;
  if (ld_acquire_sys(mine) < e) {
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < e) {
      if (clock64() - t0 > 8000000000ll) {
        printf("kucd: peer barrier timed out (rank %d waits for rank %d, epoch %u, sees %u)\n", me, j, e,
               ld_acquire_sys(mine));
        __trap();
      }
    }
  }
}

// this rank's [db | dc] into every rank's bias slot for it
__global__ void push_bias_kernel(const float* __restrict__ mine, PeerSet ps, int me, int n, int len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const float v = mine[i];
  for (int j = 0; j < n; ++j) ps.bias_slot[j][static_cast<int64_t>(me) * len + i] = v;
}

// out[i] = sum over ranks of slot s (fixed order: the result is the same on every rank and in every run)
__global__ void reduce_bias_kernel(const float* __restrict__ slots, int n, int len, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += slots[static_cast<int64_t>(j) * len + i];
  out[i] = s;
}

// The owner's part of the update (rbm.py:127-128 on rows [r0, r0 + rows) of W): dW = sum of the n ranks' slots (fixed
// order), fp32 master and momentum updated locally, and the refreshed bf16 rows stored into EVERY rank's operand
// plane - the all-gather of the new W happens inside the update kernel, as plain NVLink stores.
// W16: the slots hold bf16 partial sums (kEpiRawPush16; `slots` is then a __nv_bfloat16 array, slot pitch slice_elems
// elements as before); they are widened and summed in fp32, in rank order.
template <bool W16>
__global__ void update_w_sharded_kernel(float* __restrict__ W, const float* __restrict__ slots, int64_t slice_elems,
                                        int n, float* __restrict__ mom, PeerSet ps, int64_t elem0, int64_t n4, float lr,
                                        float scale, float momentum, float weight_decay, const StepDyn* sdyn,
                                        float world) {
  scale = step_scale(scale, sdyn, world);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      float4 t;
      if constexpr (W16) {
        const uint2 h = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(slots) + j * slice_elems)[i];
        t = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xFFFF0000u), __uint_as_float(h.y << 16),
                        __uint_as_float(h.y & 0xFFFF0000u));
      } else {
        t = reinterpret_cast<const float4*>(slots + j * slice_elems)[i];
      }
      d.x += t.x;
      d.y += t.y;
      d.z += t.z;
      d.w += t.w;
    }
    float4* wp = reinterpret_cast<float4*>(W + elem0) + i;
    float4 w = *wp;
    float4 s;
    s.x = lr * (scale * d.x - weight_decay * w.x);
    s.y = lr * (scale * d.y - weight_decay * w.y);
    s.z = lr * (scale * d.z - weight_decay * w.z);
    s.w = lr * (scale * d.w - weight_decay * w.w);
    if (mom != nullptr) {
      float4* mp = reinterpret_cast<float4*>(mom + elem0) + i;
      float4 m = *mp;
      m.x = momentum * m.x + s.x;
      m.y = momentum * m.y + s.y;
      m.z = momentum * m.z + s.z;
      m.w = momentum * m.w + s.w;
      *mp = m;
      s = m;
    }
    w.x += s.x;
    w.y += s.y;
    w.z += s.z;
    w.w += s.w;
    *wp = w;
    __nv_bfloat16 h[4];
    h[0] = __float2bfloat16_rn(w.x);
    h[1] = __float2bfloat16_rn(w.y);
    h[2] = __float2bfloat16_rn(w.z);
    h[3] = __float2bfloat16_rn(w.w);
    const uint2 packed = *reinterpret_cast<const uint2*>(h);
    for (int j = 0; j < n; ++j) (reinterpret_cast<uint2*>(ps.wp[j] + elem0))[i] = packed;
  }
}

// ---------------------------------------------------------------------------------------------------
// Unit-sharded exchange (kucd.cu, enqueue_cd_units): every rank computes a slice of a layer's units for ALL rows of the
// global minibatch, and the 0/1 states travel between the ranks as BITS - 1/16 of their bf16 size, 1/32 of float32.
// ---------------------------------------------------------------------------------------------------
// The rectangle rows [0, rows) x units [col_lo, col_lo + cols) of a bf16 0/1 plane, packed (unit j of a row = bit j % 8
// of byte j / 8, as everywhere in this engine) and stored into EVERY rank's slot at rows dst_row0 + r of the gathered
// (*, pitch_bytes * 8) bit matrix.  A thread packs 8 units (one 16-byte load, coalesced); the block's 256 bytes leave as
// sixteen 16-byte stores per destination.  cols and col_lo are multiples of 128, so a 16-byte piece never straddles rows.
// `dyn` (graph replay): the source rows start at dyn->row_off of a resident data set.
__global__ void __launch_bounds__(256) pack_push_kernel(const __nv_bfloat16* __restrict__ src, int64_t ld,
                                                        const StepDyn* dyn, int64_t rows, int64_t col_lo, int64_t cols,
                                                        PeerSet ps, int n, int slot, int64_t dst_row0,
                                                        int64_t pitch_bytes) {
  // No shared memory (sixteen lanes assemble a 16-byte piece with shuffles): a block of this kernel then fits next to a
  // persistent contraction CTA that holds all but the last KB of an SM's shared memory, so an exchange on the second
  // stream really runs while an independent projection computes.
  const int64_t row_off = dyn != nullptr ? dyn->row_off : 0;
  const int64_t groups_per_row = cols / 8;  // bytes of one row of the rectangle (a multiple of 16)
  const int64_t total = rows * groups_per_row;
  const uint32_t lane = threadIdx.x & 31u;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * 256; base < total; base += static_cast<int64_t>(gridDim.x) * 256) {
    const int64_t g = base + threadIdx.x;  // total is a multiple of 16: a 16-lane group is inside or outside as a whole
    uint32_t b = 0;
    int64_t r = 0, c8 = 0;
    if (g < total) {
      r = g / groups_per_row;
      c8 = g - r * groups_per_row;
      const uint4 q = *reinterpret_cast<const uint4*>(src + (row_off + r) * ld + col_lo + 8 * c8);
      const Bf16x8 v{{q.x, q.y, q.z, q.w}};
      b = bf16x8_to_bits(v, 8);
    }
    // lane l (l % 4 == 0) collects bytes l .. l+3; lane l (l % 16 == 0) collects the four words of its 16 bytes
    uint32_t w = b | (__shfl_down_sync(0xffffffffu, b, 1) << 8) | (__shfl_down_sync(0xffffffffu, b, 2) << 16) |
                 (__shfl_down_sync(0xffffffffu, b, 3) << 24);
    uint4 piece;
    piece.x = w;
    piece.y = __shfl_down_sync(0xffffffffu, w, 4);
    piece.z = __shfl_down_sync(0xffffffffu, w, 8);
    piece.w = __shfl_down_sync(0xffffffffu, w, 12);
    if ((lane & 15u) == 0 && g < total) {
      const int64_t off = static_cast<int64_t>(slot) * kBitSlotBytes + (dst_row0 + r) * pitch_bytes + (col_lo >> 3) + c8;
      for (int j = 0; j < n; ++j) *reinterpret_cast<uint4*>(ps.bits[j] + off) = piece;
    }
  }
}

// The unit-sharded update (rbm.py:127-128 on the columns [h_lo, h_lo + hs) of W that this rank owns): fp32 master and
// momentum updated locally, the refreshed bf16 values stored into this rank's operand plane (it projects onto these
// hidden units) and into the plane of the rank that owns visible unit v = row (it projects back onto that visible
// unit and needs the whole row): the all-to-all that keeps the row slices current is the update kernel's own store
// stream - V x hs x 2 bytes leave per rank and step instead of a reduce-scatter of V x H partial sums.
__global__ void update_w_units_kernel(float* __restrict__ W, const float* __restrict__ dW, float* __restrict__ mom,
                                      PeerSet ps, int me, int64_t ld, int64_t rows, int64_t h_lo, int64_t hs,
                                      int64_t rows_per_rank, float lr, float scale, float momentum, float weight_decay,
                                      const StepDyn* sdyn, float world) {
  scale = step_scale(scale, sdyn, world);
  const int64_t q_per_row = hs / 4;
  const int64_t n4 = rows * q_per_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t v = i / q_per_row, c4 = i - v * q_per_row;
    const int64_t off = v * ld + h_lo + 4 * c4;
    const float4 d = *reinterpret_cast<const float4*>(dW + off);
    float4* wp = reinterpret_cast<float4*>(W + off);
    float4 w = *wp;
    float4 s;
    s.x = lr * (scale * d.x - weight_decay * w.x);
    s.y = lr * (scale * d.y - weight_decay * w.y);
    s.z = lr * (scale * d.z - weight_decay * w.z);
    s.w = lr * (scale * d.w - weight_decay * w.w);
    if (mom != nullptr) {
      float4* mp = reinterpret_cast<float4*>(mom + off);
      float4 m = *mp;
      m.x = momentum * m.x + s.x;
      m.y = momentum * m.y + s.y;
      m.z = momentum * m.z + s.z;
      m.w = momentum * m.w + s.w;
      *mp = m;
      s = m;
    }
    w.x += s.x;
    w.y += s.y;
    w.z += s.z;
    w.w += s.w;
    *wp = w;
    __nv_bfloat16 h[4];
    h[0] = __float2bfloat16_rn(w.x);
    h[1] = __float2bfloat16_rn(w.y);
    h[2] = __float2bfloat16_rn(w.z);
    h[3] = __float2bfloat16_rn(w.w);
    const uint2 packed = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(ps.wp[me] + off) = packed;
    const int owner = static_cast<int>(v / rows_per_rank);
    if (owner != me) *reinterpret_cast<uint2*>(ps.wp[owner] + off) = packed;
  }
}

// Streamed fit under graph replay: the step's statistic goes straight into the caller-visible page-locked array (a
// 4-byte posted write over PCIe), slot = index of the minibatch inside this pass (row_off advances by `batch` per step).
// A chunked fit restarts row_off at every chunk; dyn->pad then carries the index of the chunk's first minibatch.
__global__ void log_stat_kernel(const float* __restrict__ stat, const StepDyn* dyn, int32_t batch, float* host_log) {
  host_log[dyn->pad + dyn->row_off / batch] = *stat;
}

__global__ void set_dyn_kernel(StepDyn* dyn, int64_t row_off, int32_t rows_valid, uint64_t step, int32_t slot_base = 0) {
  dyn->row_off = row_off;
  dyn->rows_valid = rows_valid;
  dyn->pad = slot_base;
  dyn->step = step;
}

// next minibatch: sequential slices, remainder last (rbm.py:163,211,218)
__global__ void advance_dyn_kernel(StepDyn* dyn, int32_t batch, int64_t total_rows) {
  const int64_t off = dyn->row_off + batch;
  const int64_t left = total_rows - off;
  dyn->row_off = off;
  dyn->rows_valid = static_cast<int32_t>(left < 0 ? 0 : (left < batch ? left : batch));
  dyn->step += 1;
}

}  // namespace kucd
