// mix_bench.cu -- does moving the row reductions from the LSU (RED.v4) to the TMA unit (cp.reduce.async.bulk)
// help a kernel that ALSO gathers through L1?  Per iteration a warp gathers 8 random 64-byte rows (bf16-like, 4 lanes
// x 16 B each, twice) and sends 8 fp32 rows (128 B) to random rows of an L2-resident table.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mix_bench mix_bench.cu && ./mix_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE 0: gathers only   1: gathers + RED.v4 (8 lanes per row, whole rows)   2: gathers + TMA bulk reduce   3: RED only   4: TMA only
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_mix(const uint4* __restrict__ value, float* table, uint32_t vmask, uint32_t rmask, int iters, float* sink) {
  extern __shared__ __align__(128) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr int STAGES = 4;                 // 4 stages x 4 rows x 128 B per warp
  float* mine = sm + warp * (STAGES * 128);
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const uint32_t base = hash32(wid * 7919u + it);
    float4 pay = make_float4(1.f, 2.f, 3.f, 4.f);
    if (MODE <= 2) {
      // 2 x 8 random 64-byte rows: group of 4 lanes per row
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t row = hash32(base * 4 + r * 8 + (lane >> 2)) & vmask;
        const uint4 v = __ldg(value + (size_t)row * 4 + (lane & 3));
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
      }
      pay.x = acc;
    }
    if (MODE == 1 || MODE == 3) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t row = hash32(base * 2 + r * 4 + (lane >> 3) + 77) & rmask;
        red4(table + (size_t)row * 32 + (lane & 7) * 4, pay.x, pay.y, pay.z, pay.w);
      }
    }
    if (MODE == 2 || MODE == 4) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int st = (it * 2 + r) % STAGES;
        if (it * 2 + r >= STAGES) {
          if (lane < 4) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(STAGES - 1) : "memory");
          __syncwarp();
        }
        float* stage = mine + st * 128;
        *reinterpret_cast<float4*>(stage + lane * 4) = pay;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane < 4) {
          const uint32_t row = hash32(base * 2 + r * 4 + lane + 77) & rmask;
          float* dst = table + (size_t)row * 32;
          const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage + lane * 32);
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" :: "l"(dst), "r"(src) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
  }
  if (MODE == 2 || MODE == 4) { if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t vrows = 1u << 20, rrows = 1u << 19;       // 64 MB value rows (64 B each), 64 MB fp32 table
  uint4* value; float* table; float* sink;
  CK(cudaMalloc(&value, (size_t)vrows * 64)); CK(cudaMemset(value, 0, (size_t)vrows * 64));
  CK(cudaMalloc(&table, (size_t)rrows * 128)); CK(cudaMemset(table, 0, (size_t)rrows * 128));
  CK(cudaMalloc(&sink, 16));
  const int ctas = sms * 2, iters = 4000, smem = 8 * 4 * 512;
  auto run = [&](auto kern, const char* name) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    kern<<<ctas, 256, smem>>>(value, table, vrows - 1, rrows - 1, iters, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int r = 0; r < 3; ++r) kern<<<ctas, 256, smem>>>(value, table, vrows - 1, rrows - 1, iters, sink);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 3;
    const double rows = (double)ctas * 8 * iters * 8;
    printf("%-44s %8.3f ms   %6.2f G rows/s (8 reduce rows + 32 gather rows per warp-iteration)\n", name, ms, rows / ms / 1e6);
  };
  run(k_mix<0>, "gathers only");
  run(k_mix<3>, "RED.v4 only");
  run(k_mix<4>, "TMA bulk reduce only");
  run(k_mix<1>, "gathers + RED.v4");
  run(k_mix<2>, "gathers + TMA bulk reduce");
  return 0;
}
