// Issue-rate microbenchmarks for the instructions the softmax passes of the attention kernels are made of
// (build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/micro/pipes.cu; run on a B200).
// Each kernel runs N independent chains per thread at full occupancy; reports warp-instructions / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(512) k(float* out, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  uint32_t acc = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      if (OP == 2) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(v[(i + 1) & 7])); acc ^= r; }
      if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]));
      if (OP == 5) { uint64_t a = ((uint64_t)__float_as_uint(v[i]) << 32) | __float_as_uint(v[(i + 1) & 7]), r;
                     asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(r) : "l"(a)); acc ^= (uint32_t)r; }
    }
  }
  float s = __uint_as_float(acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;
}
template <int OP>
void run(const char* name, int extra_per_op) {
  float* d; cudaMalloc(&d, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms * 4, 512>>>(d, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP><<<sms * 4, 512>>>(d, 0.5f);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr_per_sm = 4.0 * 16 * ITERS * 8;   // CTAs/SM x warps x iterations x ops
  printf("%-28s %.3f ms  -> %.2f warp-instr/clk/SM at the max clock %d MHz (x32 lanes = %.0f lanes/clk/SM)%s\n", name, ms,
         warp_instr_per_sm / (ms * 1e-3 * khz * 1e3), khz / 1000, 32 * warp_instr_per_sm / (ms * 1e-3 * khz * 1e3),
         extra_per_op ? "  [+1 LOP per op]" : "");
}
int main() {
  run<1>("fma.rn.f32", 0);
  run<0>("ex2.approx.ftz.f32", 0);
  run<3>("rcp.approx.ftz.f32", 0);
  run<2>("cvt.rn.bf16x2.f32", 1);
  run<4>("max.f32", 0);
  run<5>("fma.rn.f32x2", 1);
  return 0;
}
