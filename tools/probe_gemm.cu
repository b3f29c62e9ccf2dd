// Bring-up probe for csrc/gemm.cuh: runs the tcgen05 kernel against a naive
// CUDA-core contraction on integer-valued bf16 data (every partial sum is exact
// in fp32, so the comparison is bit-exact regardless of accumulation order).
//
//   probe_gemm <a_mn> <b_mn> <M> <N> <K> <bn> <nseg> <negmask> [timing_iters] [lboA sboA advA lboB sboB advB]
//
// One configuration per process: a protocol bug traps the context, and the next
// case must start from a clean one.
#define KUCD_PROBE 1  // the descriptor overrides and work-skipping switches exist only in this binary
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../keras_unsupervised_b200/csrc/launch.cuh"

using namespace kucd;

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

// logical A_s(m,k), B_s(k,n) -> memory per major-ness
__global__ void ref_kernel(const __nv_bfloat16* const* a, const __nv_bfloat16* const* b, int nseg, uint32_t negmask,
                           int a_mn, int b_mn, int M, int N, int K, int64_t lda, int64_t ldb, float* c, int64_t ldc) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int s = 0; s < nseg; ++s) {
    float part = 0.f;
    for (int k = 0; k < K; ++k) {
      const float av = __bfloat162float(a_mn ? a[s][(int64_t)k * lda + m] : a[s][(int64_t)m * lda + k]);
      const float bv = __bfloat162float(b_mn ? b[s][(int64_t)k * ldb + n] : b[s][(int64_t)n * ldb + k]);
      part += av * bv;
    }
    acc += ((negmask >> s) & 1u) ? -part : part;
  }
  c[(int64_t)m * ldc + n] = acc;
}

// SM clock actually delivered during the timing loop: cycle counter and wall clock of one SM, before and after
__global__ void clock_probe(long long* out) {
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  out[0] = clock64();
  out[1] = static_cast<long long>(t);
  out[2] = smid;
}

static uint32_t rng_state = 12345u;
static inline uint32_t xr() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 17;
  rng_state ^= rng_state << 5;
  return rng_state;
}

int main(int argc, char** argv) {
  if (argc < 9) {
    printf("usage: %s a_mn b_mn M N K bn nseg negmask [iters] [lboA sboA advA lboB sboB advB]\n", argv[0]);
    return 1;
  }
  const int a_mn = atoi(argv[1]), b_mn = atoi(argv[2]);
  const int M = atoi(argv[3]), N = atoi(argv[4]), K = atoi(argv[5]);
  const int bn = atoi(argv[6]), nseg = atoi(argv[7]);
  const uint32_t negmask = (uint32_t)strtoul(argv[8], nullptr, 0);
  const int iters = argc > 9 ? atoi(argv[9]) : 0;
  uint32_t dbg[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 6 && 10 + i < argc; ++i) dbg[i] = (uint32_t)strtoul(argv[10 + i], nullptr, 0);

  int dev = 0, num_sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));

  auto rup = [](int64_t x, int64_t m) { return (x + m - 1) / m * m; };
  // memory shapes
  const int64_t a_rows = a_mn ? K : M, a_cols = a_mn ? M : K, lda = rup(a_cols, 64);
  const int64_t b_rows = b_mn ? K : N, b_cols = b_mn ? N : K, ldb = rup(b_cols, 64);
  const int64_t ldc = rup(N, 64);

  std::vector<__nv_bfloat16*> dA(nseg), dB(nseg);
  for (int s = 0; s < nseg; ++s) {
    std::vector<__nv_bfloat16> hA(a_rows * lda), hB(b_rows * ldb);
    // pads are filled with garbage on purpose: TMA bounds (not the pad) must make tails exact
    const char* fill = getenv("KUCD_PROBE_FILL");  // "zero": all-zero operands, "rand": full-mantissa random bf16
    if (fill != nullptr && fill[0] == 'z') {
      for (auto& v : hA) v = __float2bfloat16(0.f);
      for (auto& v : hB) v = __float2bfloat16(0.f);
    } else if (fill != nullptr && fill[0] == 'r') {
      for (auto& v : hA) v = __float2bfloat16(((int)(xr() % 65536) - 32768) / 32768.0f);
      for (auto& v : hB) v = __float2bfloat16(((int)(xr() % 65536) - 32768) / 32768.0f);
    } else {
      for (auto& v : hA) v = __float2bfloat16((float)((int)(xr() % 5) - 2));
      for (auto& v : hB) v = __float2bfloat16((float)((int)(xr() % 5) - 2));
    }
    CK(cudaMalloc(&dA[s], hA.size() * 2));
    CK(cudaMalloc(&dB[s], hB.size() * 2));
    CK(cudaMemcpy(dA[s], hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB[s], hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  }
  __nv_bfloat16 **dAp, **dBp;
  CK(cudaMalloc(&dAp, nseg * sizeof(void*)));
  CK(cudaMalloc(&dBp, nseg * sizeof(void*)));
  CK(cudaMemcpy(dAp, dA.data(), nseg * sizeof(void*), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBp, dB.data(), nseg * sizeof(void*), cudaMemcpyHostToDevice));

  float *dC, *dRef;
  CK(cudaMalloc(&dC, (size_t)M * ldc * 4));
  CK(cudaMalloc(&dRef, (size_t)M * ldc * 4));
  CK(cudaMemset(dC, 0xFF, (size_t)M * ldc * 4));
  CK(cudaMemset(dRef, 0, (size_t)M * ldc * 4));

  ref_kernel<<<dim3((N + 127) / 128, M), 128>>>(dAp, dBp, nseg, negmask, a_mn, b_mn, M, N, K, lda, ldb, dRef, ldc);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  GemmOperands ops;
  ops.num_seg = nseg;
  ops.neg_mask = negmask;
  ops.a_mn = a_mn;
  ops.b_mn = b_mn;
  ops.M = M;
  ops.N = N;
  ops.K = K;
  for (int s = 0; s < nseg; ++s) {
    ops.a[s] = MatView{dA[s], a_rows, a_cols, lda};
    ops.b[s] = MatView{dB[s], b_rows, b_cols, ldb};
  }
  GemmParams p;
  memset(&p, 0, sizeof p);
  p.out_f32 = dC;
  p.ld_f32 = ldc;
  p.dbg_lbo_a = dbg[0];
  p.dbg_sbo_a = dbg[1];
  p.dbg_adv_a = dbg[2];
  p.dbg_lbo_b = dbg[3];
  p.dbg_sbo_b = dbg[4];
  p.dbg_adv_b = dbg[5];
  if (const char* f = getenv("KUCD_DBG_FLAGS")) p.dbg_flags = (uint32_t)atoi(f);
  std::string err;
  if (!launch_gemm(p, ops, kEpiRaw, num_sms, 0, &err, bn)) {
    printf("LAUNCH FAIL: %s\n", err.c_str());
    return 3;
  }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    printf("RESULT a_mn=%d b_mn=%d M=%d N=%d K=%d bn=%d nseg=%d neg=%u : KERNEL ERROR %s\n", a_mn, b_mn, M, N, K, bn,
           nseg, negmask, cudaGetErrorString(se));
    return 4;
  }
  std::vector<float> hC((size_t)M * ldc), hR((size_t)M * ldc);
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hR.data(), dRef, hR.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  double maxd = 0;
  int fm = -1, fn = -1;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      const float c = hC[(size_t)m * ldc + n], r = hR[(size_t)m * ldc + n];
      if (!(c == r)) {
        if (bad == 0) fm = m, fn = n;
        ++bad;
        const double d = fabs((double)c - (double)r);
        if (d > maxd || d != d) maxd = d;
      }
    }
  printf("RESULT a_mn=%d b_mn=%d M=%d N=%d K=%d bn=%d nseg=%d neg=%u dbg=%u,%u,%u,%u,%u,%u : %s  bad=%lld/%lld maxdiff=%g",
         a_mn, b_mn, M, N, K, bn, nseg, negmask, dbg[0], dbg[1], dbg[2], dbg[3], dbg[4], dbg[5],
         bad == 0 ? "PASS" : "FAIL", bad, (long long)M * N, maxd);
  if (bad) printf(" first=(%d,%d) got=%g want=%g", fm, fn, hC[(size_t)fm * ldc + fn], hR[(size_t)fm * ldc + fn]);
  printf("\n");

  if (iters > 0 && (bad == 0 || p.dbg_flags != 0 || getenv("KUCD_PROBE_FILL") != nullptr)) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    long long* dclk;
    CK(cudaMalloc(&dclk, 6 * sizeof(long long)));
    for (int i = 0; i < 3; ++i) launch_gemm(p, ops, kEpiRaw, num_sms, 0, &err, bn);
    clock_probe<<<1, 1>>>(dclk);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch_gemm(p, ops, kEpiRaw, num_sms, 0, &err, bn);
    CK(cudaEventRecord(e1));
    clock_probe<<<1, 1>>>(dclk + 3);
    CK(cudaEventSynchronize(e1));
    CK(cudaDeviceSynchronize());
    long long hclk[6];
    CK(cudaMemcpy(hclk, dclk, sizeof hclk, cudaMemcpyDeviceToHost));
    if (hclk[2] == hclk[5])
      printf("CLOCK sm %lld: %.0f MHz averaged over the timing loop\n", hclk[2],
             1e3 * double(hclk[3] - hclk[0]) / double(hclk[4] - hclk[1]));
    else
      printf("CLOCK probes landed on different SMs (%lld, %lld)\n", hclk[2], hclk[5]);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * M * N * (double)K * nseg;
    printf("TIMING a_mn=%d b_mn=%d M=%d N=%d K=%d bn=%d nseg=%d : %.3f us/launch  %.1f TFLOP/s\n", a_mn, b_mn, M, N, K,
           bn, nseg, ms * 1000.0 / iters, flop * iters / (ms * 1e-3) / 1e12);
  }
  return bad == 0 ? 0 : 5;
}
