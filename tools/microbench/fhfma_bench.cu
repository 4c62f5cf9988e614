// fhfma_bench.cu -- issue rate of sm_100's mixed-precision FMA (fma.rn.f32.bf16 -> FHFMA.BF16 with .H0/.H1 operand selectors)
// against the ways the kernels widen-and-accumulate bf16 today.  Each thread runs 8 independent accumulator chains.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fhfma_bench fhfma_bench.cu && ./fhfma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
  float d; asm volatile("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c)); return d;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}

template <int MODE>
__global__ void k(const uint32_t *in, float *out, int iters) {
  uint32_t x[4], y[4];
  for (int i = 0; i < 4; ++i) { x[i] = in[threadIdx.x * 8 + i]; y[i] = in[threadIdx.x * 8 + 4 + i]; }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long acc2[4] = {0, 0, 0, 0};
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {            // 8 FHFMA: one 16-byte bf16 row dotted with another, no widening
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fhfma((unsigned short)(x[i] & 0xffff), (unsigned short)(y[i] & 0xffff), acc[2 * i]);
        acc[2 * i + 1] = fhfma((unsigned short)(x[i] >> 16), (unsigned short)(y[i] >> 16), acc[2 * i + 1]);
      }
    } else if (MODE == 1) {     // what the kernels do: widen one row (8 ALU ops) + 4 FFMA2 against an fp32x2 operand
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned long long v = pack(__uint_as_float(x[i] << 16), __uint_as_float(x[i] & 0xffff0000u));
        const unsigned long long w = pack(__uint_as_float(y[i]), __uint_as_float(y[i] ^ 0x1u));
        acc2[i] = ffma2(v, w, acc2[i]);
      }
    } else {                    // 4 FFMA2 alone (operands already fp32)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc2[i] = ffma2(pack(__uint_as_float(x[i]), __uint_as_float(y[i])), acc2[(i + 1) & 3], acc2[i]);
    }
    x[0] += it;                 // keep the loop from being hoisted
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  for (int i = 0; i < 4; ++i) s += (float)(acc2[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, const uint32_t *in, float *out) {
  const int iters = 20000, blocks = 148 * 8, threads = 256;
  k<MODE><<<blocks, threads>>>(in, out, 100);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(in, out, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double rows = (double)blocks * threads * iters;     // 8-channel row operations
  printf("%-44s %8.3f ms  %7.1f G row-ops/s  (%.2f channel-MACs per clk per SM at 1.965 GHz)\n", name, ms, rows / ms * 1e-6,
         rows * 8 / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
  uint32_t *in; float *out;
  cudaMalloc(&in, 256 * 8 * 4); cudaMemset(in, 0x3f, 256 * 8 * 4);
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("8 x FHFMA.BF16 (bf16 . bf16, fp32 accumulate)", in, out);
  run<1>("8 widen ops + 4 x FFMA2 (today's code)", in, out);
  run<2>("4 x FFMA2 only", in, out);
  return cudaDeviceSynchronize() != cudaSuccess;
}
