// The whole Gibbs chain of one minibatch as ONE persistent kernel.
//
// A CD-k step is 2k+1 projections, each consuming the previous one's output:
//   h0 = S(v0.W + c);  v1 = S(h0.W^T + b);  h1 = S(v1.W + c);  ...  vk;  hk = sigmoid(vk.W + c)   (rbm.py:119-124)
// Launched one by one, every projection pays a launch, a prologue (barriers, tensor-memory allocation, first TMA
// round trip), a last-epilogue tail during which the tensor pipe idles, and a partial last wave (512 tiles on 148
// SMs = 3.46 waves).  But the dependency between consecutive projections is only per ROW BLOCK: an output tile
// (rows R, columns C) of projection s+1 needs rows R of projection s - all its column tiles - and nothing else.
// So the chain is flattened into one tile sequence (stage-major, row-block-major inside a stage), CTA pairs walk it
// round-robin, and before loading the A operand of a tile the producer waits until the row block it reads has
// been completely written: a counter per (stage, row block) in global memory, incremented by every CTA that finishes
// a tile of that row block, read with acquire semantics.  The tail of stage s overlaps the head of stage s+1, and
// quantisation is paid once per minibatch (5376 tiles = 72.6 waves at C3) instead of once per projection.
//
// Every wait is on a tile that comes EARLIER in the sequence and every pair walks the sequence in order, so the
// smallest unfinished tile can always run: no deadlock as long as all pairs are resident (grid <= SM count).
// Memory ordering: epilogue stores (generic proxy) -> fence.proxy.async -> CTA barrier -> __threadfence +
// atomicAdd (release);  consumer: ld.acquire spin -> fence.proxy.async -> TMA load (async proxy).
//
// Same pipeline as gemm_bf16_kernel<256, ..., CG = 2>: TMA producer warp, one MMA-issuing thread on rank 0 of the
// pair (tcgen05.mma.cta_group::2, 256 x 256 tiles), eight epilogue warps per CTA, double-buffered TMEM.
#pragma once
#include "gemm.cuh"

namespace kucd {

constexpr int kChainBN = 256;

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// <256, 2>: CTA pairs on 256 x 256 tiles (large minibatches).  <64, 1>: single CTAs on 128 x 64 tiles - the
// latency-bound sizes (C1, C2), where a stage is a handful of tiles and what matters is that projections, dW and
// their dependencies cost no launches.
// GAUSS: the epilogue set of the Gaussian-visible mode (relu-threshold hiddens, rbm.py:58-59; v ~ N(h.W^T + b, I),
// rbm.py:64-66; the final hidden term stays the sigmoid, rbm.py:145) - a separate instantiation, so that the
// Bernoulli kernel's code and register allocation are exactly what they were.
// CH > 0: float32-grade mode (three bf16 term planes of W, see gemm.cuh): a projection is kd.nseg K-segments over the
// same A operand and the planes kd.map_b (smallest term first), kd.map_b2, kd.map_b3, and the accumulation is cut every
// CH k-blocks - each piece accumulates from zero in one of the two tensor-memory buffers, the epilogue warps add the
// pieces in registers with IEEE fp32 adds (BN / 2 partial sums per thread) and run the fused epilogue on the total.
template <int BN, int CG, bool GAUSS = false, int CH = 0>
__global__ void __launch_bounds__(kNumThreads, 1) chain_kernel(const __grid_constant__ ChainParams p) {
  constexpr int kBNLocal = BN / CG;
  constexpr int kTileM = kBlockM * CG;
  constexpr int kTmemCols = 2 * BN;
  using Cfg = GemmCfg<kBNLocal>;
  constexpr int kStages = Cfg::kStages;
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;

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
  const int done_stride = p.done_stride;

  int32_t dyn_row_off = 0;
  int32_t m_valid = p.m_valid;
  uint64_t draw_base = p.draw;
  int64_t row0 = p.row0;
  if (p.dyn != nullptr) {
    row0 = static_cast<int64_t>(p.dyn_rank) * p.dyn->rows_valid;
    dyn_row_off = static_cast<int32_t>(p.dyn->row_off);
    m_valid = p.dyn->rows_valid < m_valid ? p.dyn->rows_valid : m_valid;
    draw_base += p.dyn->step * p.draw_stride;
  }

  if (warp == 0 && lane == 0 && p.total_tiles > 0) {  // (an empty launch - the host's cooperative-launch probe - has no maps)
    for (int i = 0; i < kChainMaps; ++i) ptx::prefetch_tensormap(&p.maps[i]);
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
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // walk the flattened (stage, row block, column tile) sequence: q -> stage s, first tile `base` of that stage
  auto advance = [&](int q, int& s, int& base) {
    while (q >= base + p.kinds[p.stages[s].kind].num_m * p.kinds[p.stages[s].kind].num_n) {
      base += p.kinds[p.stages[s].kind].num_m * p.kinds[p.stages[s].kind].num_n;
      ++s;
    }
  };

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      auto tma_ld = [](void* dst, const CUtensorMap* tm, uint64_t* bar, int32_t c0, int32_t c1) {
        if constexpr (CG == 2) ptx::tma_load_2d_pair(dst, tm, bar, c0, c1);
        else ptx::tma_load_2d(dst, tm, bar, c0, c1);
      };
      uint32_t stage = 0, phase = 0;
      int s = 0, base = 0;
      for (int q = unit; q < p.total_tiles; q += num_units) {
        advance(q, s, base);
        const ChainKind& kd = p.kinds[p.stages[s].kind];
        const int t = q - base;
        const int m_blk = t / kd.num_n, n_blk = t % kd.num_n;
        auto wait_block = [&](int dep, int mb) {
          // row block `mb` of stage `dep` must have been written by every column tile (both CTAs of each pair)
          const uint32_t need = static_cast<uint32_t>(p.kinds[p.stages[dep].kind].num_n) * CG;
          const uint32_t* flag = p.done + dep * done_stride + mb;
          if (ld_acquire_gpu(flag) < need) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(flag) < need) {
              if (clock64() - t0 > KUCD_SPIN_LIMIT_CYCLES) {
                printf("kucd: chain dependency wait timed out (block %d stage %d waits for stage %d row block %d: %u of %u)\n",
                       blockIdx.x, s, dep, mb, ld_acquire_gpu(flag), need);
                __trap();
              }
            }
          }
        };
        const int dep = p.stages[s].dep;
        if (dep >= 0) {
          if (kd.a_mn) {
            for (int mb = 0; mb < p.num_m_batch; ++mb) wait_block(dep, mb);
          } else {
            wait_block(dep, m_blk);
          }
          ptx::fence_proxy_async_global();
        }
        const int m0 = m_blk * kTileM + static_cast<int>(cta_rank) * kBlockM;
        const int n0 = n_blk * BN + static_cast<int>(cta_rank) * kBNLocal;
        for (int seg = 0; seg < kd.nseg; ++seg) {
          if (CH == 0 && seg == 1) {
            for (int mb = 0; mb < p.num_m_batch; ++mb) wait_block(kd.dep2, mb);
            ptx::fence_proxy_async_global();
          }
          const CUtensorMap* ma = &p.maps[(CH > 0 || seg == 0) ? kd.map_a : kd.map_a2];
          const CUtensorMap* mb = &p.maps[seg == 0 ? kd.map_b : (seg == 1 ? kd.map_b2 : kd.map_b3)];
          // data-set rows: M if K-major, K if MN-major (the dW stage reads the data set in its first segment only; at
          // float32 grade every segment multiplies the same A operand)
          const int a_off = (kd.a_dyn && (CH > 0 || seg == 0)) ? dyn_row_off : 0;
          for (int kb = 0; kb < kd.kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * CG);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            const int k0 = kb * kBlockK;
            if (!kd.a_mn) {
              tma_ld(sa, ma, &full_bar[stage], k0, m0 + a_off);  // box {64 k, 128 rows}
            } else {
#pragma unroll
              for (int j = 0; j < kBlockM / 64; ++j)  // boxes {64 m, 64 k}
                tma_ld(sa + j * (kBlockK * 128), ma, &full_bar[stage], m0 + 64 * j, k0 + a_off);
            }
            if (kd.b_mn) {
#pragma unroll
              for (int j = 0; j < kBNLocal / 64; ++j)  // boxes {64 n, 64 k}
                tma_ld(sb + j * (kBlockK * 128), mb, &full_bar[stage], n0 + 64 * j, k0);
            } else {
              tma_ld(sb, mb, &full_bar[stage], k0, n0);  // box {64 k, 128 n}
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
    // ======================= MMA issuer (rank 0 of the pair) =======================
    if (lane == 0 && cta_rank == 0) {
      auto commit = [](uint64_t* bar) {
        if constexpr (CG == 2) ptx::mma_commit_pair(bar, 0b11);
        else ptx::mma_commit(bar);
      };
      uint32_t stage = 0, phase = 0;
      uint32_t accn = 0;
      int s = 0, base = 0;
      for (int q = unit; q < p.total_tiles; q += num_units) {
        advance(q, s, base);
        const ChainKind& kd = p.kinds[p.stages[s].kind];
        const bool a_mn = kd.a_mn != 0, b_mn = kd.b_mn != 0;
        const uint32_t idesc_pos = make_idesc(kTileM, BN, a_mn, b_mn, false);
        const uint32_t idesc_neg = make_idesc(kTileM, BN, a_mn, b_mn, true);
        const uint32_t lbo_a = a_mn ? kBlockK * 128u : 16u, adv_a = a_mn ? 2048u : 32u;
        const uint32_t lbo_b = b_mn ? kBlockK * 128u : 16u, adv_b = b_mn ? 2048u : 32u;
        if constexpr (CH == 0) {
          const uint32_t as = accn & 1u;
          ptx::mbar_wait(&tmem_empty_bar[as], ((accn >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int seg = 0; seg < kd.nseg; ++seg) {
            const uint32_t idesc = seg == 1 ? idesc_neg : idesc_pos;  // second segment: -vk^T hk
            for (int kb = 0; kb < kd.kblocks; ++kb) {
              ptx::mbar_wait(&full_bar[stage], phase);
              ptx::tc_fence_after();
              const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
              const uint32_t sb = sa + Cfg::kABytes;
              const uint64_t da = make_smem_desc(sa, lbo_a, 1024u);
              const uint64_t db = make_smem_desc(sb, lbo_b, 1024u);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                ptx::mma_bf16<CG>(d_tmem, da + ((k * adv_a) >> 4), db + ((k * adv_b) >> 4), idesc,
                                  (seg > 0 || kb > 0 || k > 0) ? 1u : 0u);
              commit(&empty_bar[stage]);
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
          commit(&tmem_full_bar[as]);
          ++accn;
        } else {
          // pieces of CH k-blocks, alternating between the two accumulator buffers
          const int total = kd.nseg * kd.kblocks;
          uint32_t as = 0, d_tmem = 0;
          for (int it = 0; it < total; ++it) {
            const int in_piece = it % (CH > 0 ? CH : 1);
            if (in_piece == 0) {
              as = accn & 1u;
              ptx::mbar_wait(&tmem_empty_bar[as], ((accn >> 1) & 1u) ^ 1u);
              ptx::tc_fence_after();
              d_tmem = tmem_base + as * BN;
            }
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t sb = sa + Cfg::kABytes;
            const uint64_t da = make_smem_desc(sa, lbo_a, 1024u);
            const uint64_t db = make_smem_desc(sb, lbo_b, 1024u);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              ptx::mma_bf16<CG>(d_tmem, da + ((k * adv_a) >> 4), db + ((k * adv_b) >> 4), idesc_pos,
                                (in_piece > 0 || k > 0) ? 1u : 0u);
            commit(&empty_bar[stage]);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
            if ((it + 1) % (CH > 0 ? CH : 1) == 0 || it + 1 == total) {
              commit(&tmem_full_bar[as]);
              ++accn;
            }
          }
        }
      }
    }
  } else {
    // ======================= epilogue =======================
    const uint32_t ew = warp - 2;
    const uint32_t quarter = warp & 3u;
    const uint32_t half = ew >> 2;
    constexpr int kColsPerWarp = BN / 2;
    uint32_t accn = 0;
    int s = 0, base = 0;
    for (int q = unit; q < p.total_tiles; q += num_units) {
      advance(q, s, base);
      const ChainKind& kd = p.kinds[p.stages[s].kind];
      const int t = q - base;
      const int m_blk = t / kd.num_n, n_blk = t % kd.num_n;
      const int row = m_blk * kTileM + static_cast<int>(cta_rank) * kBlockM + quarter * 32 + lane;
      const bool row_ok = row < (kd.batch_rows ? m_valid : kd.M);
      const uint64_t draw = draw_base + p.stages[s].phase;
      float row_acc = 0.f;
      auto run_epilogue = [&](const uint32_t (&acc)[32], int coff) {
        if constexpr (GAUSS) {
          if (kd.epi == kEpiReluSample)
            epilogue_chunk<kEpiReluSample>(kd, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc);
          else if (kd.epi == kEpiGaussian)
            epilogue_chunk<kEpiGaussian>(kd, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc);
          else if (kd.epi == kEpiProb)
            epilogue_chunk<kEpiProb>(kd, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc);
          else
            epilogue_chunk<kEpiRaw>(kd, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc);
        } else {
          if (kd.epi == kEpiSample)
            epilogue_chunk<kEpiSample>(kd, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc);
          else if (kd.epi == kEpiProb)
            epilogue_chunk<kEpiProb>(kd, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc);
          else
            epilogue_chunk<kEpiRaw>(kd, acc, row, n_blk * BN + coff, row_ok, draw, row0, lane, row_acc);
        }
      };
      auto release = [&](uint32_t as) {
        if constexpr (CG == 2) ptx::mbar_arrive_cluster(&tmem_empty_bar[as], 0);
        else ptx::mbar_arrive(&tmem_empty_bar[as]);
      };
      if constexpr (CH == 0) {
        const uint32_t as = accn & 1u;
        ptx::mbar_wait(&tmem_full_bar[as], (accn >> 1) & 1u);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp; c += 32) {
          const int coff = half * kColsPerWarp + c;
          uint32_t acc[32];
          ptx::tmem_ld_32x32(tmem_base + ((quarter * 32u) << 16) + as * BN + coff, acc);
          ptx::tmem_ld_wait();
          run_epilogue(acc, coff);
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_global();  // these stores will be read by other CTAs' TMA loads
        __syncwarp();
        if (lane == 0) release(as);
        ++accn;
      } else {
        float sum[kColsPerWarp];
#pragma unroll
        for (int j = 0; j < kColsPerWarp; ++j) sum[j] = 0.f;
        const int total = kd.nseg * kd.kblocks;
        const int pieces = (total + (CH > 0 ? CH : 1) - 1) / (CH > 0 ? CH : 1);
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
          run_epilogue(acc, half * kColsPerWarp + c);
        }
        ptx::fence_proxy_async_global();
        __syncwarp();
      }
      // this CTA's part of the tile is in global memory once all eight epilogue warps got here
      ptx::named_bar_sync(1, kNumEpiWarps * 32);
      if (ew == 0 && lane == 0) {
        __threadfence();
        atomicAdd(p.done + s * done_stride + m_blk, 1u);
      }
    }
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

// The host side launches a chain kernel through its address (cudaLaunchKernelExC), so that the instantiations can live
// in a translation unit of their own (csrc/inst_chain.cu, KUCD_SPLIT_BUILD).
template <int BN, int CG, bool GAUSS, int CH>
const void* chain_kernel_ptr() {
  return reinterpret_cast<const void*>(&chain_kernel<BN, CG, GAUSS, CH>);
}

#ifndef KUCD_PRECISE_CH
#define KUCD_PRECISE_CH 8  // launch.cuh: k-blocks per accumulation piece of the float32-grade mode
#endif
// every variant launch_chain (kucd.cu) uses
#define KUCD_CHAIN_VARIANTS(X)                                                                               \
  X(256, 2, false, 0) X(256, 2, true, 0) X(64, 1, false, 0) X(64, 1, true, 0) X(256, 1, false, 0) X(128, 1, false, 0) \
  X(256, 2, false, KUCD_PRECISE_CH) X(64, 1, false, KUCD_PRECISE_CH)
#ifdef KUCD_SPLIT_BUILD
#define KUCD_EXTERN_CHAIN(BN, CG, G, CH) extern template const void* chain_kernel_ptr<BN, CG, G, CH>();
KUCD_CHAIN_VARIANTS(KUCD_EXTERN_CHAIN)
#undef KUCD_EXTERN_CHAIN
#endif

}  // namespace kucd
