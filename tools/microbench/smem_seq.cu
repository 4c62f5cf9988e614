// smem_seq.cu -- probe for the shared-memory staged backward: one warp owns a window of fp32 rows and
// applies read-modify-write updates row by row (32 lanes = one 128-byte row), ILP independent rows per step.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_seq smem_seq.cu && ./smem_seq
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int ILP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_seq(float* out, int iters, int win_rows) {
  extern __shared__ __align__(128) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* win = sm + (size_t)warp * win_rows * 32;
  for (int i = lane; i < win_rows * 32; i += 32) win[i] = 0.f;
  __syncwarp();
  const uint32_t wid = blockIdx.x * WARPS + warp;
  float g = 1.0f + lane;
  for (int it = 0; it < iters; ++it) {
    // lane-parallel "geometry": every lane derives one sample's row + coefficient, then they are broadcast
    const uint32_t h = hash32(wid * 9781u + it * 32 + lane);
    const int my_row = (int)(h % (uint32_t)(win_rows - ILP));
    const float my_c = (float)(h >> 20) * 1e-3f;
#pragma unroll 1
    for (int s = 0; s < 32; ++s) {
      const int row = __shfl_sync(0xffffffffu, my_row, s);
      const float c = __shfl_sync(0xffffffffu, my_c, s);
      float v[ILP];
#pragma unroll
      for (int k = 0; k < ILP; ++k) v[k] = win[(row + k) * 32 + lane];      // ILP distinct rows (a sample's corners)
#pragma unroll
      for (int k = 0; k < ILP; ++k) win[(row + k) * 32 + lane] = fmaf(c, g, v[k]);
    }
  }
  __syncwarp();
  float acc = 0.f;
  for (int i = lane; i < win_rows * 32; i += 32) acc += win[i];
  out[blockIdx.x * WARPS * 32 + threadIdx.x] = acc;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * 4));
  const int iters = 200;
  auto run = [&](auto kern, int warps, int ctas_per_sm, int win_rows, int ilp, const char* name) {
    size_t smem = (size_t)warps * win_rows * 128;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    kern<<<sms * ctas_per_sm, warps * 32, smem>>>(out, iters, win_rows); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int r = 0; r < 5; ++r) kern<<<sms * ctas_per_sm, warps * 32, smem>>>(out, iters, win_rows);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 5;
    double rows = (double)sms * ctas_per_sm * warps * iters * 32 * ilp;
    printf("%-40s warps/CTA %d CTAs/SM %d win %4d rows: %7.3f ms  %7.2f Grows/s  %5.2f clk/row/SM@1.965GHz\n", name, warps, ctas_per_sm,
           win_rows, ms, rows / ms / 1e6, ms * 1e-3 * 1.965e9 * sms / rows);
  };
  run(k_seq<4, 4>, 4, 1, 300, 4, "seq RMW ILP4");
  run(k_seq<4, 4>, 4, 2, 200, 4, "seq RMW ILP4");
  run(k_seq<4, 8>, 8, 1, 200, 4, "seq RMW ILP4");
  run(k_seq<4, 8>, 8, 2, 100, 4, "seq RMW ILP4");
  run(k_seq<4, 16>, 16, 1, 100, 4, "seq RMW ILP4");
  run(k_seq<2, 8>, 8, 1, 200, 2, "seq RMW ILP2");
  run(k_seq<1, 8>, 8, 1, 200, 1, "seq RMW ILP1");
  run(k_seq<8, 8>, 8, 1, 200, 8, "seq RMW ILP8");
  return 0;
}
