// A stand-in for libcudart (and libnccl) that lets the HOST side of libkucd run without a GPU.
//
// Test infrastructure, not product: tests/test_host_dryrun.py links csrc/kucd.cu against this library instead of the
// CUDA runtime and drives the engine's entry points on the CPU.  "Device" memory is host memory, copies and memsets
// really happen, streams / events / graphs are book-keeping, and a kernel launch is RECORDED instead of executed:
//   - every launch (kernel name, grid, block, dynamic shared memory, stream, captured or direct) goes to a log the test
//     reads back, so launch sequences and counts of a training call can be asserted;
//   - cudaMemcpy* / cudaMemset* ranges that touch a cudaMalloc'ed block must lie inside it;
//   - cuTensorMapEncodeTiled (handed out through cudaGetDriverEntryPoint) checks its arguments the way the driver
//     would: 16-byte aligned base inside an allocation, the described extent inside that allocation, strides multiples
//     of 16 bytes, box sizes 1..256, inner box of at most 128 bytes for the 128-byte swizzle;
//   - for the kernels whose parameter layout is simple (update_w_kernel, update_w_sharded_kernel) the pointer / length
//     arguments are checked against the allocation table too;
//   - ncclAllReduce / ncclBroadcast are recorded with type, count and stream, buffers bounds-checked;
//   - cudaIpcGetMemHandle / OpenMemHandle pass the pointer through, so two contexts of ONE process can attach to each
//     other's exchange buffers as rank 0 and rank 1.
// Nothing numerical is computed: outputs of a dry run are meaningless, only the host logic is exercised.
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../keras_unsupervised_b200/csrc/params.h"  // GemmParams / ChainParams: plain data, no device code

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

std::mutex g_mu;
std::map<uintptr_t, size_t> g_allocs;       // base -> bytes
std::map<const void*, std::string> g_funcs;  // host stub -> mangled kernel name
std::vector<std::string> g_log;
std::vector<std::string> g_errors;
size_t g_malloc_calls = 0, g_free_calls = 0, g_live_bytes = 0, g_peak_bytes = 0;
long g_async_allocs = 0;  // cudaMallocAsync calls
long g_fail_malloc_in = 0;  // > 0: that many cudaMalloc calls from now, one fails (fault injection for the error paths)
size_t g_decoded_launches = 0;  // contraction launches whose parameter block was decoded and checked
int g_device = 0;
// Stream capture, modelled on the rules the engine relies on: a capture starts on one stream; another stream joins it by
// waiting on an event recorded inside the capture (fork) and must hand its work back through an event the capture waits
// on (join) before the capture ends; a capturing stream cannot wait on an event recorded outside the capture; work on a
// stream that is not part of the capture is NOT captured - it runs (is logged) directly, which is how a forgotten fork
// shows up in a test.
int g_capturing = 0;                         // id of the capture in progress (0: none); one at a time is all the engine does
int g_capture_seq = 0;
cudaStream_t g_capture_origin = nullptr;
struct Member {
  long work = 0;    // pieces of work captured on this stream
  long joined = 0;  // ... of which this many are behind an event that a member of the capture has waited for
};
std::map<cudaStream_t, Member> g_capture_streams;
struct EventState {
  int capture;      // capture id at its last record (0: recorded outside any capture)
  cudaStream_t stream;
  long work;        // the recording stream's work count at that moment
};
std::map<cudaEvent_t, EventState> g_event_state;
std::vector<std::string> g_capture;          // launches of the capture in progress
std::map<uintptr_t, std::vector<std::string>> g_graphs;
uintptr_t g_next_handle = 0x1000;

struct CallCfg {
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
};
thread_local std::vector<CallCfg> g_cfg;

void err(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_errors.push_back(buf);
}

bool captured(cudaStream_t s) { return g_capturing != 0 && g_capture_streams.count(s) != 0; }

// work enqueued on stream s: into the capture when s is part of it, else it "runs"
void logw(cudaStream_t s, const char* fmt, ...) {
  char buf[768];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (captured(s)) {
    g_capture.push_back(buf);
    g_capture_streams[s].work++;
  } else {
    g_log.push_back(buf);
  }
}

// the allocation that contains p, or g_allocs.end()
std::map<uintptr_t, size_t>::iterator owner(const void* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  auto it = g_allocs.upper_bound(a);
  if (it == g_allocs.begin()) return g_allocs.end();
  --it;
  return a < it->first + it->second ? it : g_allocs.end();
}

// a range that starts inside a device allocation must end inside it (host ranges are not tracked)
void check_range(const void* p, size_t bytes, const char* what) {
  if (p == nullptr || bytes == 0) return;
  auto it = owner(p);
  if (it == g_allocs.end()) return;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if (a + bytes > it->first + it->second)
    err("%s: %zu bytes at offset %zu overrun an allocation of %zu bytes", what, bytes, static_cast<size_t>(a - it->first),
        it->second);
}

// device ranges a captured launch was checked against: re-validated at every replay of the graph
std::vector<std::pair<uintptr_t, size_t>> g_capture_ranges;
std::map<uintptr_t, std::vector<std::pair<uintptr_t, size_t>>> g_graph_ranges;

// [p, p + bytes) must be device memory
void check_device_range(const void* p, size_t bytes, const char* what) {
  if (bytes == 0) return;
  if (g_capturing != 0 && p != nullptr) g_capture_ranges.emplace_back(reinterpret_cast<uintptr_t>(p), bytes);
  if (p == nullptr || owner(p) == g_allocs.end()) {
    err("%s: %p is not inside a device allocation", what, p);
    return;
  }
  check_range(p, bytes, what);
}

const char* kname(const void* func) {
  auto it = g_funcs.find(func);
  return it == g_funcs.end() ? "?" : it->second.c_str();
}

// template arguments of a mangled kernel name, in order: ..ILi256ELb1ELb1ELi6ELi0ELi2EEEv.. -> {256, 1, 1, 6, 0, 2}
std::vector<long> template_args(const std::string& name) {
  std::vector<long> out;
  const size_t end = name.find("EEv");
  for (size_t i = name.find('I'); i != std::string::npos && i + 2 < name.size() && (end == std::string::npos || i < end); ++i) {
    if (name[i] == 'L' && (name[i + 1] == 'i' || name[i + 1] == 'b')) {
      size_t j = i + 2;
      long v = 0;
      while (j < name.size() && name[j] >= '0' && name[j] <= '9') v = 10 * v + (name[j++] - '0');
      if (j < name.size() && name[j] == 'E' && j > i + 2) out.push_back(v);
      i = j;
    }
  }
  return out;
}

long round_up_l(long x, long m) { return (x + m - 1) / m * m; }

// What one contraction's epilogue may touch, from the fields the kernel reads (gemm.cuh: epilogue_chunk / push16_chunk).
// P is GemmParams or a ChainKind (same member names).
template <typename P>
void check_epilogue(const char* kernel, const P& p, int epi, long M, long N, const float* const* push_base, long push_rows) {
  char what[160];
  auto W = [&](const char* field) {
    snprintf(what, sizeof what, "%s: %s", kernel, field);
    return what;
  };
  if (M <= 0 || N <= 0) return;
  const long n8 = round_up_l(N, 8), n4 = round_up_l(N, 4);
  if (epi != kucd::kEpiRaw && epi != kucd::kEpiRawPush16)
    check_device_range(p.bias, static_cast<size_t>(round_up_l(N, 32)) * 4, W("bias (read a whole 32-column chunk at a time)"));
  if (epi == kucd::kEpiSample || epi == kucd::kEpiReluSample || epi == kucd::kEpiProb || epi == kucd::kEpiGaussian) {
    const size_t ext = (static_cast<size_t>(M - 1) * p.ld_bf16 + n8) * 2;
    check_device_range(p.out_bf16, ext, W("out_bf16"));
    if (p.out_mid != nullptr) check_device_range(p.out_mid, ext, W("out_mid"));
    if (p.out_lo != nullptr) check_device_range(p.out_lo, ext, W("out_lo"));
    if (p.ld_bf16 % 8 != 0 || p.ld_bf16 < n8) err("%s: ld_bf16 = %lld for %ld columns", kernel, (long long)p.ld_bf16, N);
    if (p.out_f32 != nullptr)
      check_device_range(p.out_f32, (static_cast<size_t>(M - 1) * p.ld_f32 + n4) * 4, W("out_f32 (probabilities)"));
    if (p.u_inject != nullptr)
      check_device_range(p.u_inject, (static_cast<size_t>(M - 1) * p.ld_u + N) * 4, W("u_inject"));
  }
  if (epi == kucd::kEpiFreeEnergy) check_device_range(p.rowsum, static_cast<size_t>(M) * 4, W("rowsum"));
  if (epi == kucd::kEpiRaw || epi == kucd::kEpiRawPush16) {
    const size_t esz = epi == kucd::kEpiRaw ? 4 : 2;
    const long ncols = epi == kucd::kEpiRaw ? n4 : n8;
    if (p.ld_f32 % (epi == kucd::kEpiRaw ? 4 : 8) != 0 || p.ld_f32 < ncols)
      err("%s: output pitch %lld for %ld columns", kernel, (long long)p.ld_f32, N);
    if (push_rows > 0 && push_base != nullptr) {
      const long owners = (M + push_rows - 1) / push_rows;
      if (owners > 8) err("%s: %ld owners of %ld rows for %ld rows", kernel, owners, push_rows, M);
      for (long o = 0; o < owners && o < 8; ++o) {
        const long rows = std::min(push_rows, M - o * push_rows);
        check_device_range(push_base[o], (static_cast<size_t>(rows - 1) * p.ld_f32 + ncols) * esz, W("owner slot (push_base)"));
      }
    } else if (epi == kucd::kEpiRawPush16) {
      err("%s: the bf16 push epilogue without a destination", kernel);
    } else {
      check_device_range(p.out_f32, (static_cast<size_t>(M - 1) * p.ld_f32 + ncols) * 4, W("out_f32 (raw tile)"));
    }
  }
  if (p.colsum != nullptr) check_device_range(p.colsum, static_cast<size_t>(N) * 4, W("colsum"));
}

// a descriptor produced by fake_encode_tiled: what it describes must (still) be device memory
void check_tensor_map(const CUtensorMap& tm, const char* what) {
  uint64_t base, extent, magic;
  memcpy(&base, reinterpret_cast<const char*>(&tm), 8);
  memcpy(&extent, reinterpret_cast<const char*>(&tm) + 8, 8);
  memcpy(&magic, reinterpret_cast<const char*>(&tm) + 16, 8);
  if (magic != 0x74656e736f726d61ull) {
    err("%s: not a descriptor made by cuTensorMapEncodeTiled", what);
    return;
  }
  check_device_range(reinterpret_cast<const void*>(base), extent, what);
}

// kernels whose arguments are decoded and checked against the allocation table
void check_kernel_args(const std::string& name, void** args) {
  if (args == nullptr) return;
  auto ptr = [&](int i) { return *reinterpret_cast<void**>(args[i]); };
  if (name.find("gemm_bf16_kernel") != std::string::npos) {
    const std::vector<long> t = template_args(name);  // BN, A_MN, B_MN, EPI, CH, CG
    const kucd::GemmParams& p = *reinterpret_cast<const kucd::GemmParams*>(args[0]);
    if (t.size() != 6) err("gemm: could not read the template arguments of %s", name.c_str());
    if (t.size() == 6) {
      g_decoded_launches++;
      if (p.num_seg < 1 || p.num_seg > kucd::kMaxSeg || p.kblocks < 1) err("gemm: %d segments of %d k-blocks", p.num_seg, p.kblocks);
      if (t[3] == kucd::kEpiRawPush16 && t[0] < 128) err("gemm: the bf16 push epilogue needs BN >= 128");
      for (int sg = 0; sg < p.num_seg && sg < kucd::kMaxSeg; ++sg) {
        check_tensor_map(p.tm_a[sg], "gemm_bf16_kernel: A operand descriptor");
        check_tensor_map(p.tm_b[sg], "gemm_bf16_kernel: B operand descriptor");
      }
      check_epilogue("gemm_bf16_kernel", p, static_cast<int>(t[3]), p.M, p.N, p.push_base, p.push_rows);
      if (p.dyn != nullptr) check_device_range(p.dyn, sizeof(kucd::StepDyn), "gemm_bf16_kernel: dyn");
    }
    return;
  }
  if (name.find("chain_kernel") != std::string::npos) {
    const kucd::ChainParams& p = *reinterpret_cast<const kucd::ChainParams*>(args[0]);
    if (p.num_stages < 1 || p.num_stages > kucd::kMaxChainStages) {
      err("chain: %d stages", p.num_stages);
      return;
    }
    g_decoded_launches++;
    check_device_range(p.done, static_cast<size_t>(p.num_stages) * p.done_stride * 4, "chain_kernel: completion counters");
    if (p.dyn != nullptr) check_device_range(p.dyn, sizeof(kucd::StepDyn), "chain_kernel: dyn");
    long tiles = 0;
    for (int s2 = 0; s2 < p.num_stages; ++s2) {
      const int kd = p.stages[s2].kind;
      if (kd < 0 || kd >= kucd::kMaxChainKinds) {
        err("chain: stage %d has kind %d", s2, kd);
        return;
      }
      const kucd::ChainKind& q = p.kinds[kd];
      if (p.stages[s2].dep >= s2) err("chain: stage %d waits for stage %d (must be an earlier one)", s2, p.stages[s2].dep);
      if (q.nseg == 2 && q.dep2 >= s2) err("chain: stage %d's second segment waits for stage %d", s2, q.dep2);
      if (q.num_m > p.done_stride) err("chain: stage %d has %d row blocks, the counter stride is %d", s2, q.num_m, p.done_stride);
      tiles += static_cast<long>(q.num_m) * q.num_n;
      if (q.map_a < 0 || q.map_a >= kucd::kChainMaps || q.map_b < 0 || q.map_b >= kucd::kChainMaps) {
        err("chain: stage %d uses descriptors %d / %d", s2, q.map_a, q.map_b);
        return;
      }
      check_tensor_map(p.maps[q.map_a], "chain_kernel: A operand descriptor");
      check_tensor_map(p.maps[q.map_b], "chain_kernel: B operand descriptor");
      if (q.nseg == 2) {
        check_tensor_map(p.maps[q.map_a2], "chain_kernel: A operand descriptor (second segment)");
        check_tensor_map(p.maps[q.map_b2], "chain_kernel: B operand descriptor (second segment)");
      }
      check_epilogue("chain_kernel", q, q.epi, q.M, q.N, nullptr, 0);
    }
    if (tiles != p.total_tiles) err("chain: the stages have %ld tiles, total_tiles = %d", tiles, p.total_tiles);
    return;
  }
  auto i64 = [&](int i) { return *reinterpret_cast<int64_t*>(args[i]); };
  auto i32 = [&](int i) { return *reinterpret_cast<int32_t*>(args[i]); };
  auto planes = [&](int first, size_t bytes, const char* what) {  // hi must exist, mid / lo are optional
    check_device_range(ptr(first), bytes, what);
    if (ptr(first + 1) != nullptr) check_device_range(ptr(first + 1), bytes, what);
    if (ptr(first + 2) != nullptr) check_device_range(ptr(first + 2), bytes, what);
  };
  if (name.find("ingest_kernel") != std::string::npos) {
    // (src, src_ld, rows, cols, hi, mid, lo, ld, nparts, inexact): whole padded rows of the planes are written
    const size_t esz = name.find("ingest_kernelIfE") != std::string::npos ? 4 : (name.find("ingest_kernelIhE") != std::string::npos ? 1 : 2);
    const int64_t src_ld = i64(1), rows = i64(2), cols = i64(3), ld = i64(7);
    if (rows > 0) {
      check_range(ptr(0), (static_cast<size_t>(rows - 1) * src_ld + cols) * esz, "ingest_kernel source");
      const int np = *reinterpret_cast<int*>(args[8]);
      for (int k2 = 0; k2 < np && k2 < 3; ++k2)
        check_device_range(ptr(4 + k2), static_cast<size_t>(rows) * ld * 2, "ingest_kernel plane");
      if (ld % 8 != 0 || ld < cols) err("ingest_kernel: ld = %lld for %lld columns", (long long)ld, (long long)cols);
    }
    return;
  }
  if (name.find("ingest_bits_kernel") != std::string::npos) {  // (src, src_pitch, rows, cols, hi, mid, lo, ld, nparts)
    const int64_t pitch = i64(1), rows = i64(2), cols = i64(3), ld = i64(7);
    if (rows > 0) {
      check_range(ptr(0), static_cast<size_t>(rows - 1) * pitch + (cols + 7) / 8, "ingest_bits_kernel source");
      const int np = *reinterpret_cast<int*>(args[8]);
      for (int k2 = 0; k2 < np && k2 < 3; ++k2)
        check_device_range(ptr(4 + k2), static_cast<size_t>(rows) * ld * 2, "ingest_bits_kernel plane");
    }
    return;
  }
  if (name.find("export_bits_kernel") != std::string::npos) {  // (hi, ld, rows, cols, dst, dst_pitch)
    const int64_t ld = i64(1), rows = i64(2), cols = i64(3), pitch = i64(5);
    if (rows > 0) {
      check_device_range(ptr(0), (static_cast<size_t>(rows - 1) * ld + cols) * 2, "export_bits_kernel plane");
      check_range(ptr(4), static_cast<size_t>(rows - 1) * pitch + (cols + 7) / 8, "export_bits_kernel destination");
    }
    return;
  }
  if (name.find("export_kernel") != std::string::npos) {  // (hi, mid, lo, ld, rows, cols, dst, dst_ld)
    const size_t esz = name.find("export_kernelIfE") != std::string::npos ? 4 : (name.find("export_kernelIhE") != std::string::npos ? 1 : 2);
    const int64_t ld = i64(3), rows = i64(4), cols = i64(5), dst_ld = i64(7);
    if (rows > 0) {
      planes(0, (static_cast<size_t>(rows - 1) * ld + cols) * 2, "export_kernel plane");
      check_range(ptr(6), (static_cast<size_t>(rows - 1) * dst_ld + cols) * esz, "export_kernel destination");
    }
    return;
  }
  if (name.find("copy_rows_kernel") != std::string::npos) {  // (src, dst, ld, rows, dyn)
    const int64_t ld = i64(2);
    const int32_t rows = i32(3);
    if (rows > 0) {
      check_device_range(ptr(0), static_cast<size_t>(rows) * ld * 2, "copy_rows_kernel source");
      check_device_range(ptr(1), static_cast<size_t>(rows) * ld * 2, "copy_rows_kernel destination");
    }
    return;
  }
  if (name.find("permute_rows_kernel") != std::string::npos) {  // (s0, s1, s2, d0, d1, d2, ld, rows, key)
    const int64_t ld = i64(6), rows = i64(7);
    if (rows > 0) {
      planes(0, static_cast<size_t>(rows) * ld * 2, "permute_rows_kernel source");
      planes(3, static_cast<size_t>(rows) * ld * 2, "permute_rows_kernel destination");
      if ((ptr(1) == nullptr) != (ptr(4) == nullptr)) err("permute_rows_kernel: source and destination differ in planes");
    }
    return;
  }
  if (name.find("refresh_planes_kernel") != std::string::npos) {  // (W, hi, mid, lo, n4)
    const int64_t n4 = i64(4);
    check_device_range(ptr(0), static_cast<size_t>(n4) * 16, "refresh_planes_kernel W");
    planes(1, static_cast<size_t>(n4) * 8, "refresh_planes_kernel plane");
    return;
  }
  if (name.find("update_bias_kernel") != std::string::npos) {  // (x, d, mom, n, ...)
    const int64_t n = i64(3);
    check_device_range(ptr(0), static_cast<size_t>(n) * 4, "update_bias_kernel bias");
    check_device_range(ptr(1), static_cast<size_t>(n) * 4, "update_bias_kernel statistic");
    if (ptr(2) != nullptr) check_device_range(ptr(2), static_cast<size_t>(n) * 4, "update_bias_kernel momentum");
    return;
  }
  if (name.find("colsum_store_kernel") != std::string::npos) {
    // (hi, mid, lo, ld, rows, cols, cols_pad, dyn, out, zero, zero_len): with dyn the rows start at dyn->row_off (unknown here)
    const int64_t ld = i64(3);
    const int32_t rows = i32(4), cols = i32(5), cols_pad = i32(6);
    if (ptr(7) == nullptr && rows > 0) planes(0, (static_cast<size_t>(rows - 1) * ld + cols) * 2, "colsum_store_kernel plane");
    check_device_range(ptr(8), static_cast<size_t>(cols_pad) * 4, "colsum_store_kernel sums");
    if (ptr(9) != nullptr) check_device_range(ptr(9), static_cast<size_t>(i32(10)) * 4, "colsum_store_kernel cleared block");
    return;
  }
  if (name.find("colsum_kernel") != std::string::npos) {  // (hi, mid, lo, ld, row_off, rows, cols, dyn, sign, out)
    const int64_t ld = i64(3), row_off = i64(4);
    const int32_t rows = i32(5), cols = i32(6);
    if (ptr(7) == nullptr && rows > 0)
      planes(0, (static_cast<size_t>(row_off + rows - 1) * ld + cols) * 2, "colsum_kernel plane");
    check_device_range(ptr(9), static_cast<size_t>(cols) * 4, "colsum_kernel sums");
    return;
  }
  if (name.find("update_w_sharded_kernel") != std::string::npos) {
    // (W, slots, slice_elems, n, mom, PeerSet, elem0, n4, ...): the owner's rows of W, its n slots, every rank's plane
    struct PeerSetMirror {
      float* dw_slot[8];
      float* bias_slot[8];
      uint32_t* flags[8];
      void* wp[8];
    };
    const bool w16 = name.find("ILb1") != std::string::npos;
    const int64_t slice = i64(2), elem0 = i64(6), n4 = i64(7);
    const int n = *reinterpret_cast<int*>(args[3]);
    const PeerSetMirror& ps = *reinterpret_cast<const PeerSetMirror*>(args[5]);
    if (n4 > 0 && n >= 1 && n <= 8) {
      check_device_range(static_cast<char*>(ptr(0)) + elem0 * 4, static_cast<size_t>(n4) * 16, "update_w_sharded_kernel W rows");
      if (ptr(4) != nullptr)
        check_device_range(static_cast<char*>(ptr(4)) + elem0 * 4, static_cast<size_t>(n4) * 16, "update_w_sharded_kernel momentum");
      const size_t esz = w16 ? 2 : 4;
      for (int j = 0; j < n; ++j) {
        check_device_range(static_cast<char*>(ptr(1)) + static_cast<size_t>(j) * slice * esz, static_cast<size_t>(n4) * 4 * esz,
                           "update_w_sharded_kernel slot");
        check_device_range(static_cast<char*>(ps.wp[j]) + elem0 * 2, static_cast<size_t>(n4) * 8,
                           "update_w_sharded_kernel operand plane of a rank");
      }
    }
    return;
  }
  if (name.find("update_w_kernel") != std::string::npos) {
    // (W, dW, mom, hi, mid, lo, n4, ...): n4 float4 of W / dW (uint2 when dW is bf16), n4 uint2 of each plane
    const int64_t n4 = *reinterpret_cast<int64_t*>(args[6]);
    const bool d16 = name.find("ILb1") != std::string::npos;
    if (n4 > 0) {
      check_device_range(ptr(0), static_cast<size_t>(n4) * 16, "update_w_kernel W");
      check_device_range(ptr(1), static_cast<size_t>(n4) * (d16 ? 8 : 16), "update_w_kernel dW");
      if (ptr(2) != nullptr) check_device_range(ptr(2), static_cast<size_t>(n4) * 16, "update_w_kernel momentum");
      check_device_range(ptr(3), static_cast<size_t>(n4) * 8, "update_w_kernel bf16 plane");
      if (ptr(4) != nullptr) check_device_range(ptr(4), static_cast<size_t>(n4) * 8, "update_w_kernel mid plane");
      if (ptr(5) != nullptr) check_device_range(ptr(5), static_cast<size_t>(n4) * 8, "update_w_kernel lo plane");
    }
  }
}

void record_launch(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args, int cluster) {
  const std::string name = kname(func);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024)
    err("launch of %s with grid (%u,%u,%u) block (%u,%u,%u)", name.c_str(), grid.x, grid.y, grid.z, block.x, block.y, block.z);
  if (smem > 232448) err("launch of %s with %zu bytes of dynamic shared memory", name.c_str(), smem);
  if (cluster > 1 && grid.x % cluster != 0) err("launch of %s: grid %u is not a multiple of the cluster size %d", name.c_str(), grid.x, cluster);
  check_kernel_args(name, args);
  logw(stream, "launch %s grid=%u,%u,%u block=%u,%u,%u smem=%zu stream=%p cluster=%d", name.c_str(), grid.x, grid.y, grid.z,
       block.x, block.y, block.z, smem, static_cast<void*>(stream), cluster);
}

// ---- cuTensorMapEncodeTiled stand-in --------------------------------------------------------------------------
CUresult fake_encode_tiled(CUtensorMap* tm, CUtensorMapDataType dtype, cuuint32_t rank, void* base, const cuuint64_t* dims,
                           const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                           CUtensorMapInterleave, CUtensorMapSwizzle swizzle, CUtensorMapL2promotion,
                           CUtensorMapFloatOOBfill) {
  std::lock_guard<std::mutex> lk(g_mu);
  const size_t esz = (dtype == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 || dtype == CU_TENSOR_MAP_DATA_TYPE_FLOAT16) ? 2 : 4;
  bool ok = tm != nullptr && rank >= 1 && rank <= 5 && base != nullptr;
  if (ok && (reinterpret_cast<uintptr_t>(base) & 15u)) {
    err("tensor map: base %p is not 16-byte aligned", base);
    ok = false;
  }
  for (cuuint32_t i = 0; ok && i < rank; ++i) {
    if (dims[i] == 0 || dims[i] > (1ull << 32)) {
      err("tensor map: dim[%u] = %llu", i, (unsigned long long)dims[i]);
      ok = false;
    }
    if (box[i] == 0 || box[i] > 256) {
      err("tensor map: box[%u] = %u", i, box[i]);
      ok = false;
    }
    if (estr[i] == 0 || estr[i] > 8) ok = false;
    if (i + 1 < rank && (strides[i] % 16 != 0 || strides[i] == 0)) {
      err("tensor map: stride[%u] = %llu bytes is not a positive multiple of 16", i, (unsigned long long)strides[i]);
      ok = false;
    }
  }
  if (ok && swizzle == CU_TENSOR_MAP_SWIZZLE_128B && box[0] * esz > 128) {
    err("tensor map: inner box of %zu bytes with the 128-byte swizzle", box[0] * esz);
    ok = false;
  }
  if (ok && rank == 2) {
    if (dims[0] * esz > strides[0]) {
      err("tensor map: rows of %llu bytes overlap (pitch %llu)", (unsigned long long)(dims[0] * esz), (unsigned long long)strides[0]);
      ok = false;
    }
    // the last row ends inside the allocation the base points into
    const size_t extent = static_cast<size_t>(dims[1] - 1) * strides[0] + static_cast<size_t>(dims[0]) * esz;
    auto it = owner(base);
    if (it == g_allocs.end()) {
      err("tensor map: base %p is not device memory", base);
      ok = false;
    } else if (reinterpret_cast<uintptr_t>(base) + extent > it->first + it->second) {
      err("tensor map: %llu x %llu (pitch %llu) at offset %zu overruns an allocation of %zu bytes", (unsigned long long)dims[1],
          (unsigned long long)dims[0], (unsigned long long)strides[0],
          static_cast<size_t>(reinterpret_cast<uintptr_t>(base) - it->first), it->second);
      ok = false;
    }
  }
  if (!ok) return CUDA_ERROR_INVALID_VALUE;
  // the fake descriptor remembers what it describes, so that a launch (and a later graph replay) can be checked against
  // the allocations that are live at that moment
  memset(tm, 0, sizeof *tm);
  const uint64_t extent = rank == 2 ? static_cast<uint64_t>(dims[1] - 1) * strides[0] + dims[0] * esz : dims[0] * esz;
  const uint64_t magic = 0x74656e736f726d61ull;
  memcpy(reinterpret_cast<char*>(tm), &base, 8);
  memcpy(reinterpret_cast<char*>(tm) + 8, &extent, 8);
  memcpy(reinterpret_cast<char*>(tm) + 16, &magic, 8);
  return CUDA_SUCCESS;
}

}  // namespace

#define LOCK std::lock_guard<std::mutex> lk(g_mu)

extern "C" {

// ---- test interface -------------------------------------------------------------------------------------------
void fake_reset() {
  LOCK;
  g_log.clear();
  g_errors.clear();
  g_malloc_calls = g_free_calls = 0;
  g_async_allocs = 0;
  g_decoded_launches = 0;
  g_peak_bytes = g_live_bytes;
}
int fake_log_size() {
  LOCK;
  return static_cast<int>(g_log.size());
}
int fake_log_line(int i, char* buf, int n) {
  LOCK;
  if (i < 0 || i >= static_cast<int>(g_log.size())) return -1;
  snprintf(buf, n, "%s", g_log[i].c_str());
  return 0;
}
int fake_error_count() {
  LOCK;
  return static_cast<int>(g_errors.size());
}
int fake_error_line(int i, char* buf, int n) {
  LOCK;
  if (i < 0 || i >= static_cast<int>(g_errors.size())) return -1;
  snprintf(buf, n, "%s", g_errors[i].c_str());
  return 0;
}
long long fake_counter(int which) {  // 0 cudaMalloc calls, 1 cudaFree calls, 2 live bytes, 3 peak bytes, 4 live allocations,
                                     // 5 contraction launches whose parameters were decoded
  LOCK;
  switch (which) {
    case 0: return static_cast<long long>(g_malloc_calls);
    case 1: return static_cast<long long>(g_free_calls);
    case 2: return static_cast<long long>(g_live_bytes);
    case 3: return static_cast<long long>(g_peak_bytes);
    case 5: return static_cast<long long>(g_decoded_launches);
    case 6: return static_cast<long long>(g_async_allocs);
    default: return static_cast<long long>(g_allocs.size());
  }
}

void fake_fail_malloc_in(long n) {
  LOCK;
  g_fail_malloc_in = n;
}

// ---- registration of the kernels (called by the nvcc-generated module constructor) --------------------------
void** __cudaRegisterFatBinary(void*) {
  static void* handle = nullptr;
  return &handle;
}
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char* hostFun, char*, const char* deviceName, int, uint3*, uint3*, dim3*, dim3*, int*) {
  LOCK;
  g_funcs[hostFun] = deviceName;
}
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
  g_cfg.push_back({grid, block, smem, stream});
  return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream) {
  if (g_cfg.empty()) return cudaErrorInvalidConfiguration;
  const CallCfg c = g_cfg.back();
  g_cfg.pop_back();
  *grid = c.grid;
  *block = c.block;
  *smem = c.smem;
  *static_cast<cudaStream_t*>(stream) = c.stream;
  return cudaSuccess;
}

// ---- devices ------------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) {
  const char* e = getenv("FAKE_CUDA_DEVICES");
  *n = e != nullptr ? atoi(e) : 2;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) {
  g_device = d;
  return cudaSuccess;
}
cudaError_t cudaGetDevice(int* d) {
  *d = g_device;
  return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof *p);
  snprintf(p->name, sizeof p->name, "dry-run sm_100 (no GPU)");
  p->major = 10;
  p->minor = 0;
  p->multiProcessorCount = 148;
  p->totalGlobalMem = 180ull << 30;
  return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "dry-run error"; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int value) {
  if (value > 232448) {
    LOCK;
    err("cudaFuncSetAttribute: %d bytes of dynamic shared memory", value);
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}
cudaError_t cudaOccupancyMaxActiveClusters(int* n, const void*, const cudaLaunchConfig_t*) {
  *n = 74;
  return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* res) {
  if (strcmp(symbol, "cuTensorMapEncodeTiled") == 0) {
    *fn = reinterpret_cast<void*>(&fake_encode_tiled);
    if (res != nullptr) *res = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  if (res != nullptr) *res = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorNotSupported;
}

// ---- memory -------------------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t n) {
  LOCK;
  if (g_fail_malloc_in > 0 && --g_fail_malloc_in == 0) {
    *p = nullptr;
    return cudaErrorMemoryAllocation;
  }
  void* q = nullptr;
  if (posix_memalign(&q, 512, n == 0 ? 512 : n) != 0) return cudaErrorMemoryAllocation;
  // pages are not touched here: a dry run of a C4-sized model must not need C4-sized RAM
  g_allocs[reinterpret_cast<uintptr_t>(q)] = n;
  g_malloc_calls++;
  g_live_bytes += n;
  if (g_live_bytes > g_peak_bytes) g_peak_bytes = g_live_bytes;
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (p == nullptr) return cudaSuccess;
  LOCK;
  auto it = g_allocs.find(reinterpret_cast<uintptr_t>(p));
  if (it == g_allocs.end()) {
    err("cudaFree of %p, which is not the base of a live allocation", p);
    return cudaErrorInvalidValue;
  }
  g_live_bytes -= it->second;
  g_allocs.erase(it);
  g_free_calls++;
  free(p);
  return cudaSuccess;
}
// stream-ordered allocation: the same table (bounds checks see these blocks), separate call counters
cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) {
  const cudaError_t e = cudaMalloc(p, n);
  if (e == cudaSuccess) {
    LOCK;
    g_malloc_calls--;
    g_async_allocs++;
  }
  return e;
}
cudaError_t cudaFreeAsync(void* p, cudaStream_t) {
  const cudaError_t e = cudaFree(p);
  if (e == cudaSuccess && p != nullptr) {
    LOCK;
    g_free_calls--;
  }
  return e;
}
cudaError_t cudaDeviceGetAttribute(int* value, cudaDeviceAttr, int) {
  *value = 0;  // (no cooperative launches in the dry run: the engine then issues plain ones)
  return cudaSuccess;
}
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) {
  *pool = reinterpret_cast<cudaMemPool_t>(static_cast<uintptr_t>(0x9000));
  return cudaSuccess;
}
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  *p = malloc(n == 0 ? 1 : n);
  return *p != nullptr ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void* p) {
  free(p);
  return cudaSuccess;
}
// larger copies / fills are checked but not performed; FAKE_TOUCH_LIMIT=0 performs none (host-cost measurements)
static const size_t kTouchLimit = [] {
  const char* e = getenv("FAKE_TOUCH_LIMIT");
  return e != nullptr ? static_cast<size_t>(atoll(e)) : size_t{64} << 20;
}();
cudaError_t cudaMemset(void* p, int v, size_t n) {
  LOCK;
  check_device_range(p, n, "cudaMemset");
  if (n <= kTouchLimit && owner(p) != g_allocs.end()) memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t st) {
  LOCK;
  check_device_range(p, n, "cudaMemsetAsync");
  if (captured(st)) {
    logw(st, "memset %zu", n);
    return cudaSuccess;
  }
  if (n <= kTouchLimit && owner(p) != g_allocs.end()) memset(p, v, n);
  return cudaSuccess;
}
static cudaError_t do_copy(void* d, const void* s, size_t n, const char* what, cudaStream_t st, bool async) {
  LOCK;
  check_range(d, n, what);
  check_range(s, n, what);
  if (async && captured(st)) {
    logw(st, "memcpy %zu", n);
    return cudaSuccess;
  }
  if (!async && g_capturing) err("%s (synchronous) while a capture is in progress", what);
  if (n <= kTouchLimit) memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  return do_copy(d, s, n, "cudaMemcpy", nullptr, false);
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st) {
  return do_copy(d, s, n, "cudaMemcpyAsync", st, true);
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
  LOCK;
  if (h == 0 || w == 0) return cudaSuccess;
  if (w > dp || w > sp) {
    err("cudaMemcpy2DAsync: width %zu exceeds a pitch (%zu, %zu)", w, dp, sp);
    return cudaErrorInvalidValue;
  }
  check_range(d, (h - 1) * dp + w, "cudaMemcpy2DAsync dst");
  check_range(s, (h - 1) * sp + w, "cudaMemcpy2DAsync src");
  if (w * h <= kTouchLimit)
    for (size_t r = 0; r < h; ++r) memmove(static_cast<char*>(d) + r * dp, static_cast<const char*>(s) + r * sp, w);
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  memset(h, 0, sizeof *h);
  memcpy(h, &p, sizeof p);
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  memcpy(p, &h, sizeof *p);
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// ---- streams, events, graphs --------------------------------------------------------------------------------
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  LOCK;
  *s = reinterpret_cast<cudaStream_t>(g_next_handle += 16);
  return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t st) {
  LOCK;
  if (captured(st)) {
    err("cudaStreamSynchronize on a capturing stream");
    return cudaErrorStreamCaptureUnsupported;
  }
  return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned) {
  LOCK;
  if (e == nullptr) {
    err("cudaStreamWaitEvent on a NULL event");
    return cudaErrorInvalidResourceHandle;
  }
  auto it = g_event_state.find(e);
  if (it == g_event_state.end()) {
    err("cudaStreamWaitEvent on an event that was never recorded (the wait orders nothing)");
    return cudaSuccess;
  }
  const int ev_capture = it->second.capture;
  const cudaStream_t ev_stream = it->second.stream;
  if (g_capturing && ev_capture == g_capturing) {
    // an event of this capture: the waiting stream becomes (or stays) part of it, and the recording stream's work
    // up to that event is now depended upon
    {
      char buf[160];
      snprintf(buf, sizeof buf, "event_wait event=%p stream=%p", static_cast<void*>(e), static_cast<void*>(st));
      g_capture.push_back(buf);
    }
    g_capture_streams[st];
    Member& m = g_capture_streams[ev_stream];
    if (st != ev_stream && it->second.work > m.joined) m.joined = it->second.work;
    return cudaSuccess;
  }
  if (captured(st)) {
    err("a capturing stream waits on an event recorded outside the capture (cudaErrorStreamCaptureIsolation)");
    return cudaErrorStreamCaptureIsolation;
  }
  return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) {
  LOCK;
  *e = reinterpret_cast<cudaEvent_t>(g_next_handle += 16);
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st) {
  LOCK;
  if (e == nullptr) {
    err("cudaEventRecord on a NULL event");
    return cudaErrorInvalidResourceHandle;
  }
  g_event_state[e] = {captured(st) ? g_capturing : 0, st, captured(st) ? g_capture_streams[st].work : 0};
  if (captured(st)) {
    char buf[160];
    snprintf(buf, sizeof buf, "event_record event=%p stream=%p", static_cast<void*>(e), static_cast<void*>(st));
    g_capture.push_back(buf);
  }
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  *ms = 0.f;
  return cudaSuccess;
}
cudaError_t cudaStreamIsCapturing(cudaStream_t st, cudaStreamCaptureStatus* status) {
  LOCK;
  *status = captured(st) ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
  return cudaSuccess;
}
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode) {
  LOCK;
  if (g_capturing) {
    err("nested stream capture");
    return cudaErrorIllegalState;
  }
  g_capturing = ++g_capture_seq;
  g_capture_origin = st;
  g_capture_streams.clear();
  g_capture_streams[st];
  g_capture.clear();
  g_capture_ranges.clear();
  return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t* g) {
  LOCK;
  if (!g_capturing) return cudaErrorIllegalState;
  if (st != g_capture_origin) err("cudaStreamEndCapture on a stream that did not begin the capture");
  for (const auto& kv : g_capture_streams)
    if (kv.first != g_capture_origin && kv.second.work > kv.second.joined)
      err("the capture ends with work on a forked stream that was never joined (cudaErrorStreamCaptureUnjoined)");
  g_capturing = 0;
  g_capture_streams.clear();
  const uintptr_t h = (g_next_handle += 16);
  g_graphs[h] = g_capture;
  g_graph_ranges[h] = g_capture_ranges;
  *g = reinterpret_cast<cudaGraph_t>(h);
  return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* ge, cudaGraph_t g, unsigned long long) {
  *ge = reinterpret_cast<cudaGraphExec_t>(g);
  return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t ge, cudaStream_t) {
  LOCK;
  auto it = g_graphs.find(reinterpret_cast<uintptr_t>(ge));
  if (it == g_graphs.end()) {
    err("cudaGraphLaunch of an unknown or destroyed graph");
    return cudaErrorInvalidValue;
  }
  // what the captured launches were pointed at must still be device memory: a graph kept across a reallocation of one
  // of its buffers replays into freed memory
  for (const auto& r : g_graph_ranges[it->first]) {
    auto o = owner(reinterpret_cast<const void*>(r.first));
    if (o == g_allocs.end() || r.first + r.second > o->first + o->second) {
      err("cudaGraphLaunch: a captured launch points at %zu bytes that are no longer inside a live allocation", r.second);
      break;
    }
  }
  g_log.push_back("graph_launch nodes=" + std::to_string(it->second.size()));
  return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) {
  LOCK;
  g_graphs.erase(reinterpret_cast<uintptr_t>(g));
  g_graph_ranges.erase(reinterpret_cast<uintptr_t>(g));
  return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }

// the nodes of a captured graph, for the test (graph = the handle cudaStreamEndCapture returned most recently)
int fake_last_graph_size() {
  LOCK;
  return g_graphs.empty() ? -1 : static_cast<int>(g_graphs.rbegin()->second.size());
}
int fake_last_graph_line(int i, char* buf, int n) {
  LOCK;
  if (g_graphs.empty()) return -1;
  const auto& v = g_graphs.rbegin()->second;
  if (i < 0 || i >= static_cast<int>(v.size())) return -1;
  snprintf(buf, n, "%s", v[i].c_str());
  return 0;
}

// ---- launches -----------------------------------------------------------------------------------------------
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
  LOCK;
  record_launch(func, grid, block, smem, stream, args, 1);
  return cudaSuccess;
}
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t* cfg, const void* func, void** args) {
  LOCK;
  int cluster = 1;
  for (unsigned i = 0; i < cfg->numAttrs; ++i)
    if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) cluster = static_cast<int>(cfg->attrs[i].val.clusterDim.x);
  record_launch(func, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream, args, cluster);
  return cudaSuccess;
}

// ---- NCCL (KUCD_NCCL_LIB points at this library) --------------------------------------------------------------
struct FakeUid {
  char internal[128];
};
int ncclGetUniqueId(FakeUid* id) {
  memset(id, 7, sizeof *id);
  return 0;
}
int ncclCommInitRank(void** comm, int world, FakeUid, int rank) {
  *comm = reinterpret_cast<void*>(static_cast<uintptr_t>(0x7000 + world * 16 + rank));
  return 0;
}
int ncclCommDestroy(void*) { return 0; }
static size_t nccl_esize(int dtype) { return dtype == 9 || dtype == 6 ? 2 : (dtype == 8 || dtype == 4 || dtype == 5 ? 8 : (dtype <= 1 ? 1 : 4)); }
int ncclAllReduce(const void* s, void* d, size_t count, int dtype, int op, void* comm, cudaStream_t stream) {
  LOCK;
  if (comm == nullptr) err("ncclAllReduce without a communicator");
  check_device_range(s, count * nccl_esize(dtype), "ncclAllReduce send");
  check_device_range(d, count * nccl_esize(dtype), "ncclAllReduce recv");
  logw(stream, "allreduce count=%zu dtype=%d op=%d stream=%p", count, dtype, op, static_cast<void*>(stream));
  return 0;
}
int ncclBroadcast(const void* s, void* d, size_t count, int dtype, int root, void* comm, cudaStream_t stream) {
  LOCK;
  if (comm == nullptr) err("ncclBroadcast without a communicator");
  check_device_range(d, count * nccl_esize(dtype), "ncclBroadcast recv");
  (void)s;
  logw(stream, "broadcast count=%zu dtype=%d root=%d stream=%p", count, dtype, root, static_cast<void*>(stream));
  return 0;
}
int ncclAllGather(const void* s, void* d, size_t count, int dtype, void* comm, cudaStream_t stream) {
  LOCK;
  if (comm == nullptr) err("ncclAllGather without a communicator");
  const size_t world = (reinterpret_cast<uintptr_t>(comm) - 0x7000) / 16;  // see ncclCommInitRank
  check_device_range(s, count * nccl_esize(dtype), "ncclAllGather send");
  check_device_range(d, world * count * nccl_esize(dtype), "ncclAllGather recv");
  logw(stream, "allgather count=%zu dtype=%d world=%zu stream=%p", count, dtype, world, static_cast<void*>(stream));
  return 0;
}
int ncclGroupStart() { return 0; }
int ncclGroupEnd() { return 0; }
const char* ncclGetErrorString(int) { return "dry-run nccl"; }

}  // extern "C"
