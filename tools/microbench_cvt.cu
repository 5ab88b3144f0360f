// Issue-rate probes for the instructions of the int8 layers' epilogue (run under gpurun):
// I2FP / F2I conversions (XU pipe?), packed FFMA2, FMNMX, and the conversion-free replacements
// (magic-number int→float, FADD.RZ float→int).  Prints warp instructions per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_cvt microbench_cvt.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 16;
constexpr int kIters = 2048;

template <int kOp>
__global__ void __launch_bounds__(256) probe(uint32_t *out, uint32_t seed, float fa, float fb) {
  uint32_t r[kChains];
  uint64_t p[kChains / 2];
#pragma unroll
  for (int i = 0; i < kChains; ++i) r[i] = seed + threadIdx.x * 977u + i * 131u;
#pragma unroll
  for (int i = 0; i < kChains / 2; ++i) p[i] = (uint64_t(r[2 * i]) << 32) | r[2 * i + 1];
  const uint64_t a2 = (uint64_t(__float_as_uint(fa)) << 32) | __float_as_uint(fa), b2 = (uint64_t(__float_as_uint(fb)) << 32) | __float_as_uint(fb);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (kOp == 0) r[i] = __float_as_uint(__int2float_rn(int(r[i]))) ^ 0x5bd1e995u;               // I2F + LOP
      if (kOp == 1) r[i] = uint32_t(__float2int_rz(__uint_as_float((r[i] & 0x007fffffu) | 0x42000000u)));  // LOP + F2I
      if (kOp == 2) r[i] = (r[i] & 0x007fffffu) ^ 0x5bd1e995u;                                     // LOP only (baseline)
      if (kOp == 3) r[i] = __float_as_uint(fminf(fmaxf(__uint_as_float(r[i]), fa), fb));          // FMNMX x2
      if (kOp == 4) r[i] = __float_as_uint(__fadd_rz(__uint_as_float((r[i] & 0x007fffffu) | 0x42000000u), 8388608.0f));  // LOP + FADD.RZ
      if (kOp == 6) r[i] = __float_as_uint(__fadd_rn(__uint_as_float(((r[i] + 0x400000u) & 0x7fffffu) | 0x4B000000u), -12582912.0f));  // IADD+LOP+FADD
    }
    if (kOp == 5) {
#pragma unroll
      for (int i = 0; i < kChains / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(a2), "l"(b2));  // FFMA2
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= r[i];
#pragma unroll
  for (int i = 0; i < kChains / 2; ++i) s ^= uint32_t(p[i]) ^ uint32_t(p[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t *out;
  cudaMalloc(&out, sizeof(uint32_t) * sms * 8 * 256);
  const int grid = sms * 8;
  const double warp_iters = double(grid) * 8 * kIters * kChains;  // warp-level executions of the loop body element
  auto report = [&](const char *name, float ms, double instr_per_elem) {
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-34s %.3f ms  → %.2f cycles per warp-element per SMSP, i.e. %.2f warp-instr/clk/SM at %.1f instr per element\n", name, ms,
           clk / (warp_iters / sms / 4), warp_iters * instr_per_elem / sms / clk, instr_per_elem);
  };
  report("I2FP.F32.S32 + LOP3", time_ms([&] { probe<0><<<grid, 256>>>(out, 1, 0.5f, 2.f); }), 2);
  report("LOP3 + F2I.TRUNC", time_ms([&] { probe<1><<<grid, 256>>>(out, 1, 0.5f, 2.f); }), 2);
  report("LOP3 + LOP3 (baseline)", time_ms([&] { probe<2><<<grid, 256>>>(out, 1, 0.5f, 2.f); }), 2);
  report("FMNMX + FMNMX", time_ms([&] { probe<3><<<grid, 256>>>(out, 1, 0.5f, 2.f); }), 2);
  report("LOP3 + FADD.RZ (float→int magic)", time_ms([&] { probe<4><<<grid, 256>>>(out, 1, 0.5f, 2.f); }), 2);
  report("FFMA2 (0.5 per element)", time_ms([&] { probe<5><<<grid, 256>>>(out, 1, 0.9999f, 1e-3f); }), 0.5);
  report("IADD + LOP3 + FADD (int→float magic)", time_ms([&] { probe<6><<<grid, 256>>>(out, 1, 0.5f, 2.f); }), 3);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
