#pragma once
#include <cuda_runtime.h>

namespace grb {

// Packed FP32 (Blackwell FMUL2 / FFMA2 / FADD2): two independent IEEE roundings per instruction, so the results are
// the reference's bits.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false, so the
// accumulation acc + p is issued as fma(p, 1, acc) with the 1 coming from a kernel argument: p * 1 is exact, hence
// the FFMA2 rounds the same p + acc once (also for signed zeros), and ptxas cannot fold the multiply above into it.
typedef unsigned long long df_u64;
__device__ __forceinline__ df_u64 df_pack(float a0, float a1) {
  df_u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a0), "f"(a1));
  return r;
}
__device__ __forceinline__ void df_unpack(df_u64 v, float& a0, float& a1) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(v));
}
__device__ __forceinline__ df_u64 df_mul2(df_u64 a, df_u64 b) {
  df_u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ df_u64 df_add2(df_u64 a, df_u64 b) {
  df_u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ df_u64 df_acc2(df_u64 p, df_u64 ones, df_u64 acc) {  // acc + p, two lanes
  df_u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p), "l"(ones), "l"(acc));
  return r;
}

}  // namespace grb
