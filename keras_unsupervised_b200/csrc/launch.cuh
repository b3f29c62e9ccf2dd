// Host side of gemm.cuh: TMA tensor maps and the launch dispatcher.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "gemm.cuh"

namespace kucd {

// A bf16 matrix in device memory as the engine lays it out: row-major (rows, cols),
// leading dimension `ld` elements (multiple of 64 so every row is 128-byte aligned).
struct MatView {
  const void* ptr = nullptr;
  int64_t rows = 0, cols = 0, ld = 0;
};

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Tensor map over a row-major bf16 matrix with a {64 x box_rows} box and 128B swizzle.
// Out-of-bounds elements read as zero, which is what makes ragged M/N/K tails exact.
//
// A descriptor is a pure function of (address, shape, leading dimension, box), and a training loop asks for the same
// dozen of them at every step (cuTensorMapEncodeTiled costs 1-2 us each, 12 per chain launch: a fifth of a
// latency-bound step's host time), so the last encodings are kept per host thread.
struct TmapCache {
  static constexpr int kSlots = 64;
  struct Key {
    const void* ptr;
    int64_t rows, cols, ld;
    uint32_t box_rows;
  };
  Key keys[kSlots];
  CUtensorMap maps[kSlots];
  int used = 0, next = 0;
};

inline bool make_tmap_bf16(CUtensorMap* tm, const MatView& m, uint32_t box_rows, std::string* err) {
  static thread_local TmapCache cache;
  for (int i = 0; i < cache.used; ++i) {
    const TmapCache::Key& k = cache.keys[i];
    if (k.ptr == m.ptr && k.rows == m.rows && k.cols == m.cols && k.ld == m.ld && k.box_rows == box_rows) {
      memcpy(tm, &cache.maps[i], sizeof(CUtensorMap));
      return true;
    }
  }
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) {
    if (err) *err = "cuTensorMapEncodeTiled unavailable";
    return false;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(m.cols), static_cast<cuuint64_t>(m.rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(m.ld) * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(m.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%lld cols=%lld ld=%lld box_rows=%u",
               static_cast<int>(r), m.ptr, (long long)m.rows, (long long)m.cols, (long long)m.ld, box_rows);
      *err = buf;
    }
    return false;
  }
  const int slot = cache.used < TmapCache::kSlots ? cache.used++ : (cache.next++ % TmapCache::kSlots);
  cache.keys[slot] = TmapCache::Key{m.ptr, m.rows, m.cols, m.ld, box_rows};
  memcpy(&cache.maps[slot], tm, sizeof(CUtensorMap));
  return true;
}

// One K-segment of a contraction.  `a_mn` / `b_mn` say how the operand sits in memory:
//   A K-major : a = (M, K) row-major        A MN-major: a = (K, M) row-major
//   B K-major : b = (N, K) row-major        B MN-major: b = (K, N) row-major
struct GemmOperands {
  MatView a[kMaxSeg], b[kMaxSeg];
  int num_seg = 1;
  uint32_t neg_mask = 0;
  bool a_mn = false, b_mn = false;
  int64_t M = 0, N = 0, K = 0;
};

constexpr int kMaxDevices = 64;
// index of the calling thread's current device (contexts on several GPUs may live in one process)
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

inline int pick_bn(int64_t M, int64_t N, int num_sms) {
  // Prefer the widest tile that still gives every SM a tile; small problems take narrower tiles.
  const int64_t m_tiles = (M + kBlockM - 1) / kBlockM;
  for (int bn : {256, 128}) {
    if (m_tiles * ((N + bn - 1) / bn) >= num_sms) return bn;
  }
  return 64;
}

constexpr int kPreciseBN = 128;  // precise mode keeps BN/2 = 64 partial sums per epilogue thread in registers
// k-blocks (of 64) per tensor-memory accumulation piece of the float32-grade mode.  Measured on one B200, same call
// (profiles/r02_ch_experiment.log; tools/diag/ch_experiment.py): max relative error of sigmoid(v.W + c) against float64,
// real-valued v, K = 4096 / 16384, and the C3 float32-grade step -
//   CH = 4: 1.2e-6 / 3.2e-6, 8.39 ms      CH = 8: 2.1e-6 / 4.6e-6, 7.74 ms      CH = 16: 3.7e-6 / 8.2e-6, 7.22 ms
// (binary inputs: 0.8e-6 / 1.6e-6 whatever CH).  8 keeps a factor two under the 1e-5 bar at C4's K = 16384.
#ifndef KUCD_PRECISE_CH
#define KUCD_PRECISE_CH 8
#endif
constexpr int kPreciseCH = KUCD_PRECISE_CH;

template <int BN, bool A_MN, bool B_MN, int EPI, int CH = 0, int CG = 1>
inline cudaError_t launch_one(const GemmParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN / CG, (epi_stages(EPI) ? kEpiStageBytes : 0)>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI, CH, CG>;
  static bool attr_set[kMaxDevices] = {};  // a function attribute belongs to the device it was set on
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int tile_m = kBlockM * CG;
  const int num_tiles = ((p.M + tile_m - 1) / tile_m) * ((p.N + BN - 1) / BN);
  const int units = num_sms / CG;  // persistent: one CTA (or CTA pair) per SM (or TPC)
  const int grid = (num_tiles < units ? num_tiles : units) * CG;
  if (grid <= 0) return cudaSuccess;
  if constexpr (CG == 1) {
    kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(p);
    return cudaGetLastError();
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // the pair shares a TPC
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
  }
}

// Operand-major combinations the engine instantiates per epilogue (keeps the fatbin small):
//   forward  v.W   : A K-major, B MN-major      backward h.W^T : A K-major, B K-major
//   delta W        : A MN-major, B MN-major     (raw epilogue also builds the 4th combo for the probe)
template <int EPI, bool A_MN, bool B_MN>
constexpr bool combo_built() {
  if (EPI == kEpiRaw) return true;
  if (A_MN) return false;
  if (EPI == kEpiFreeEnergy || EPI == kEpiReluSample) return B_MN;
  if (EPI == kEpiGaussian) return !B_MN;
  return true;  // sample / prob: both directions
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CH, int CG>
inline cudaError_t launch_if_built(const GemmParams& p, int num_sms, cudaStream_t s) {
  if constexpr (combo_built<EPI, A_MN, B_MN>())
    return launch_one<BN, A_MN, B_MN, EPI, CH, CG>(p, num_sms, s);
  else
    return cudaErrorInvalidValue;
}

template <int BN, int EPI, int CH, int CG = 1>
inline cudaError_t launch_major(const GemmParams& p, bool a_mn, bool b_mn, int num_sms, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_if_built<BN, false, false, EPI, CH, CG>(p, num_sms, s);
  if (!a_mn && b_mn) return launch_if_built<BN, false, true, EPI, CH, CG>(p, num_sms, s);
  if (a_mn && b_mn) return launch_if_built<BN, true, true, EPI, CH, CG>(p, num_sms, s);
  return launch_if_built<BN, true, false, EPI, CH, CG>(p, num_sms, s);
}

// (not `inline`: the product library is built from several translation units in parallel, KUCD_SPLIT_BUILD below, and an
// explicit instantiation declaration does not hold back an inline function)
template <int EPI>
cudaError_t launch_bn(const GemmParams& p, int bn, bool precise, int cg, bool a_mn, bool b_mn, int num_sms,
                      cudaStream_t s) {
  if (precise && cg == 2) return launch_major<256, EPI, kPreciseCH, 2>(p, a_mn, b_mn, num_sms, s);
  if (precise) return launch_major<kPreciseBN, EPI, kPreciseCH>(p, a_mn, b_mn, num_sms, s);
  if (cg == 2) return launch_major<256, EPI, 0, 2>(p, a_mn, b_mn, num_sms, s);
  switch (bn) {
    case 256: return launch_major<256, EPI, 0>(p, a_mn, b_mn, num_sms, s);
    case 128: return launch_major<128, EPI, 0>(p, a_mn, b_mn, num_sms, s);
    default: return launch_major<64, EPI, 0>(p, a_mn, b_mn, num_sms, s);
  }
}

#ifdef KUCD_SPLIT_BUILD
// libkucd.so is compiled as several translation units side by side (_lib.py: build): csrc/inst_*.cu hold the explicit
// instantiations of the contraction kernels per epilogue, kucd.cu only refers to them.  Same kernels, same SASS as the
// single-unit build (the dry-run build and the probe still compile everything in one unit).
#define KUCD_EXTERN_BN(E) \
  extern template cudaError_t launch_bn<E>(const GemmParams&, int, bool, int, bool, bool, int, cudaStream_t);
KUCD_EXTERN_BN(kEpiRaw)
KUCD_EXTERN_BN(kEpiSample)
KUCD_EXTERN_BN(kEpiProb)
KUCD_EXTERN_BN(kEpiFreeEnergy)
KUCD_EXTERN_BN(kEpiReluSample)
KUCD_EXTERN_BN(kEpiGaussian)
#undef KUCD_EXTERN_BN
#endif

// Fill the tensor maps / shape fields of `p` from `ops` (epilogue fields are the caller's) and launch.
// cta_group::2 (256 x 256 tiles on CTA pairs) once the problem fills the chip with such tiles
inline int pick_cg(int64_t M, int64_t N, int num_sms) {
  static const int env = [] {
    const char* e = getenv("KUCD_CG");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (env == 1 || env == 2) return env;
  const int64_t tiles = ((M + 255) / 256) * ((N + 255) / 256);
  return tiles >= num_sms / 2 ? 2 : 1;
}

#ifndef KUCD_INST_UNIT  // (an instantiation unit has no use for the dispatcher - and must not instantiate the bf16-push kernels it names)
inline bool launch_gemm(GemmParams& p, const GemmOperands& ops, int epi, int num_sms, cudaStream_t stream,
                        std::string* err, int force_bn = 0, bool precise = false, int force_cg = 0) {
  if (ops.num_seg < 1 || ops.num_seg > kMaxSeg) {
    if (err) *err = "bad segment count";
    return false;
  }
  int bn = precise ? kPreciseBN : (force_bn ? force_bn : pick_bn(ops.M, ops.N, num_sms));
  // (precise mode on CTA pairs: 256 x 256 tiles, each epilogue thread keeps 128 partial sums in registers)
  int cg = force_cg ? force_cg : ((force_bn == 0 || force_bn == 256) ? pick_cg(ops.M, ops.N, num_sms) : 1);
  if (cg == 2) bn = 256;
  p.num_seg = ops.num_seg;
  p.neg_mask = ops.neg_mask;
  p.M = static_cast<int32_t>(ops.M);
  p.N = static_cast<int32_t>(ops.N);
  p.kblocks = static_cast<int32_t>((ops.K + kBlockK - 1) / kBlockK);
  if (p.m_valid <= 0 || p.m_valid > p.M) p.m_valid = p.M;
  for (int s = 0; s < ops.num_seg; ++s) {
    if (!make_tmap_bf16(&p.tm_a[s], ops.a[s], ops.a_mn ? 64u : static_cast<uint32_t>(kBlockM), err)) return false;
    if (!make_tmap_bf16(&p.tm_b[s], ops.b[s], ops.b_mn ? 64u : static_cast<uint32_t>(bn / cg), err)) return false;
  }
  cudaError_t e;
  switch (epi) {
    case kEpiRawPush16:  // the dW contraction of the fused exchange with bf16 partial sums: MN-major operands, BN >= 128
      if (precise || !ops.a_mn || !ops.b_mn || p.push_rows <= 0) {
        if (err) *err = "bf16 partial sums: bf16 compute, MN-major operands and a peer-mapped destination only";
        return false;
      }
      if (cg == 2) e = launch_one<256, true, true, kEpiRawPush16, 0, 2>(p, num_sms, stream);
      else if (bn == 256) e = launch_one<256, true, true, kEpiRawPush16, 0, 1>(p, num_sms, stream);
      else e = launch_one<128, true, true, kEpiRawPush16, 0, 1>(p, num_sms, stream);
      break;
    case kEpiRaw: e = launch_bn<kEpiRaw>(p, bn, precise, cg, ops.a_mn, ops.b_mn, num_sms, stream); break;
    case kEpiSample: e = launch_bn<kEpiSample>(p, bn, precise, cg, ops.a_mn, ops.b_mn, num_sms, stream); break;
    case kEpiProb: e = launch_bn<kEpiProb>(p, bn, precise, cg, ops.a_mn, ops.b_mn, num_sms, stream); break;
    case kEpiFreeEnergy: e = launch_bn<kEpiFreeEnergy>(p, bn, precise, cg, ops.a_mn, ops.b_mn, num_sms, stream); break;
    case kEpiReluSample: e = launch_bn<kEpiReluSample>(p, bn, precise, cg, ops.a_mn, ops.b_mn, num_sms, stream); break;
    case kEpiGaussian: e = launch_bn<kEpiGaussian>(p, bn, precise, cg, ops.a_mn, ops.b_mn, num_sms, stream); break;
    default:
      if (err) *err = "bad epilogue mode";
      return false;
  }
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm launch failed: ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

#endif  // KUCD_INST_UNIT

}  // namespace kucd
