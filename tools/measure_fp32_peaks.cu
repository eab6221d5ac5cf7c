// FP32 / issue-rate / latency micro-benchmarks on the box ("the builder must measure an FMA
// micro-benchmark on the box and use that").  Prints one JSON object:
//   fp32_tflops            independent FFMA, every SM full (the non-tensor FP32 roofline)
//   fp32x2_tflops          the same with packed FFMA2
//   issue_ginst_s          warp-instructions per second, whole chip (4 schedulers x 148 SMs x clock x IPC)
//   lat_*                  cycles per DEPENDENT instruction for ONE warp alone on its scheduler (what bounds the
//                          clock-recovery loop): FADD, FMUL, FFMA, packed FFMA2, IMAD, LOP3, FADD+LOP3 alternating,
//                          shared-memory load (pointer chase), and the issue interval of independent instructions
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/measure_fp32_peaks tools/measure_fp32_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int CHAIN = 4096;

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

// ---- dependent chains, one warp per CTA, one CTA per SM ------------------------------------------
template <int KIND>
__global__ void dep_chain(float* out, long long* cyc, float a, float b, unsigned m) {
  float x = a + threadIdx.x;
  unsigned u = m + threadIdx.x;
  float y0 = a, y1 = b, y2 = a + 1.f, y3 = b + 1.f;
  unsigned long long p = 0;
  if (KIND == 3) asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x), "f"(x + 1.f));
  unsigned long long pb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
  const long long t0 = clk();
#pragma unroll 1
  for (int it = 0; it < CHAIN / 64; it++) {
#pragma unroll
    for (int k = 0; k < 64; k++) {
      if (KIND == 0) x = __fadd_rn(x, b);
      if (KIND == 1) x = __fmul_rn(x, b);
      if (KIND == 2) x = __fmaf_rn(x, b, a);
      if (KIND == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(pb));
      if (KIND == 4) u = u * 3u + m;
      if (KIND == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(u) : "r"(m), "r"(0x80000000u));
      if (KIND == 6) {  // FADD -> LOP3 -> FADD ... (cross pipe)
        x = __fadd_rn(x, b);
        unsigned t = __float_as_uint(x);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(t) : "r"(m), "r"(0x80000000u));
        x = __uint_as_float(t);
      }
      if (KIND == 7) {  // dependent FADD with three independent FADDs in between (fills the latency?)
        x = __fadd_rn(x, b);
        y0 = __fadd_rn(y0, a); y1 = __fadd_rn(y1, a); y2 = __fadd_rn(y2, a);
      }
      if (KIND == 8) {  // four independent FADD chains: issue interval of one warp
        x = __fadd_rn(x, b); y0 = __fadd_rn(y0, a); y1 = __fadd_rn(y1, a); y2 = __fadd_rn(y2, a);
        y3 = __fadd_rn(y3, a); u = u + m;
        asm volatile("" : "+f"(x), "+f"(y0), "+f"(y1), "+f"(y2), "+f"(y3), "+r"(u));
      }
    }
  }
  const long long t1 = clk();
  if (KIND == 3) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); x = lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y0 + y1 + y2 + y3 + __uint_as_float(u);
  if (threadIdx.x == 0) cyc[blockIdx.x * (blockDim.x / 32) + 0] = t1 - t0;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = t1 - t0;
}

// shared-memory pointer chase: cycles per dependent LDS
__global__ void lds_chase(unsigned* out, long long* cyc, int stride_words) {
  __shared__ unsigned s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = ((i + stride_words * 32) & 4095) * 4u;
  __syncthreads();
  unsigned a = threadIdx.x * 4u;
  const unsigned base = (unsigned)__cvta_generic_to_shared(s);
  const long long t0 = clk();
#pragma unroll 1
  for (int it = 0; it < CHAIN / 64; it++) {
#pragma unroll
    for (int k = 0; k < 64; k++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(base + a));
  }
  const long long t1 = clk();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = t1 - t0;
}

// ---- throughput: every SM full of independent FFMA ---------------------------------------------
template <int PACKED>
__global__ void __launch_bounds__(1024) fma_tput(float* out, float a, float b, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = a + threadIdx.x + i;
  unsigned long long p[4], pa, pb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
#pragma unroll
  for (int i = 0; i < 4; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
      if (PACKED) {
#pragma unroll
        for (int i = 0; i < 4; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pa));
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], b, a);
      }
    }
  }
  float s = 0.f;
  if (PACKED) {
#pragma unroll
    for (int i = 0; i < 4; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

template <int KIND>
static double dep_cycles(int warps_per_cta, int ops_per_iter, float* d_out, long long* d_cyc, int nsm) {
  dep_chain<KIND><<<nsm, 32 * warps_per_cta>>>(d_out, d_cyc, 1.0f, 1.0000001f, 0x3f800000u);
  CK(cudaDeviceSynchronize());
  dep_chain<KIND><<<nsm, 32 * warps_per_cta>>>(d_out, d_cyc, 1.0f, 1.0000001f, 0x3f800000u);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(nsm * warps_per_cta);
  CK(cudaMemcpy(h.data(), d_cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  return (double)mx / ((double)CHAIN * ops_per_iter);
}

int main() {
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  const int nsm = pr.multiProcessorCount;
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, (size_t)nsm * 64 * 1024 * sizeof(float)));
  CK(cudaMalloc(&d_cyc, (size_t)nsm * 64 * sizeof(long long)));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_rate_mhz\": %.0f", pr.name, nsm, clock_khz / 1e3);
  // latencies: ONE warp per SM, then 6 warps per SM (the clock-recovery CTA shape: 2 warps alone on their scheduler)
  for (int w : {1, 4, 6, 8}) {
    printf(",\n \"lat_w%d\": {\"fadd\": %.2f, \"fmul\": %.2f, \"ffma\": %.2f, \"ffma2\": %.2f, \"imad\": %.2f, \"lop3\": %.2f, "
           "\"fadd_lop3_pair\": %.2f, \"fadd_dep_plus_3_indep\": %.2f, \"indep_6_issue_interval\": %.2f}",
           w, dep_cycles<0>(w, 1, d_out, d_cyc, nsm), dep_cycles<1>(w, 1, d_out, d_cyc, nsm), dep_cycles<2>(w, 1, d_out, d_cyc, nsm),
           dep_cycles<3>(w, 1, d_out, d_cyc, nsm), dep_cycles<4>(w, 1, d_out, d_cyc, nsm), dep_cycles<5>(w, 1, d_out, d_cyc, nsm),
           dep_cycles<6>(w, 1, d_out, d_cyc, nsm), dep_cycles<7>(w, 1, d_out, d_cyc, nsm), dep_cycles<8>(w, 6, d_out, d_cyc, nsm));
  }
  {
    unsigned* d_u; CK(cudaMalloc(&d_u, (size_t)nsm * 256 * 4));
    for (int w : {1, 6}) {
      lds_chase<<<nsm, 32 * w>>>(d_u, d_cyc, 1);
      CK(cudaDeviceSynchronize());
      lds_chase<<<nsm, 32 * w>>>(d_u, d_cyc, 1);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(nsm * w);
      CK(cudaMemcpy(h.data(), d_cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (auto v : h) mx = v > mx ? v : mx;
      printf(",\n \"lds_dependent_cycles_w%d\": %.2f", w, (double)mx / CHAIN);
    }
  }
  // throughput
  const int iters = 2000;
  const int ctas = nsm * 2;
  float ms = time_ms([&] { fma_tput<0><<<ctas, 1024>>>(d_out, 1.0f, 1.0000001f, iters); });
  const double flop = (double)ctas * 1024 * iters * 16 * 8 * 2;
  const double winst = (double)ctas * 32 * iters * 16 * 8;
  float ms2 = time_ms([&] { fma_tput<1><<<ctas, 1024>>>(d_out, 1.0f, 1.0000001f, iters); });
  const double flop2 = (double)ctas * 1024 * iters * 16 * 4 * 4;
  printf(",\n \"fp32_tflops\": %.2f, \"fp32_ffma_ms\": %.4f, \"issue_ginst_s\": %.1f, \"fp32x2_tflops\": %.2f, \"fp32x2_ms\": %.4f,\n"
         " \"how\": \"independent FFMA (8 accumulators per thread, 2 CTAs x 1024 threads per SM, %d x 128 FFMA per thread), best of 5, "
         "CUDA events; latencies = clock64 around 4096 dependent instructions, max over SMs\"}\n",
         flop / (ms * 1e-3) / 1e12, ms, winst / (ms * 1e-3) / 1e9, flop2 / (ms2 * 1e-3) / 1e12, ms2, iters);
  return 0;
}
