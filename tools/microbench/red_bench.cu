// red_bench.cu -- scatter-add throughput probes for the backward design (not part of the product).
// Each probe adds 128-byte fp32 rows into a table of R rows at pseudo-random row indices.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_bench red_bench.cu && ./red_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}

// mode 0: RED.v4, 8 lanes per row (4 rows / instr)      mode 1: RED.32, 32 lanes per row
// mode 2: RED.v4 half-sector pattern (4 lanes per row-half, two instrs per row, like bf16 today)
// mode 3: RED.v2, 16 lanes per row                       mode 4: STG.128 plain scatter (no add)
// mode 5: RED.v4 two chunks/lane, 4 lanes per row, permuted so each instr covers full sectors
template <int MODE>
__global__ void __launch_bounds__(256) k_red(float* table, uint32_t rmask, int iters, uint32_t local_rows) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    const uint32_t base = hash32(wid * 7919u + it);
    if (MODE == 0) {
      const int g = lane >> 3, j = lane & 7;
      uint32_t row = local_rows ? ((wid * 64u + (hash32(base + g) % local_rows)) & rmask) : (hash32(base + g) & rmask);
      red4(table + (size_t)row * 32 + j * 4, 1.f, 2.f, 3.f, 4.f);
    } else if (MODE == 1) {
      uint32_t row = hash32(base) & rmask;
      atomicAdd(table + (size_t)row * 32 + lane, 1.f);
    } else if (MODE == 2) {
      const int g = lane >> 2, j = lane & 3;
      uint32_t row = hash32(base + g) & rmask;
      red4(table + (size_t)row * 32 + j * 8, 1.f, 2.f, 3.f, 4.f);
      red4(table + (size_t)row * 32 + j * 8 + 4, 1.f, 2.f, 3.f, 4.f);
    } else if (MODE == 3) {
      const int g = lane >> 4, j = lane & 15;
      uint32_t row = hash32(base + g) & rmask;
      red2(table + (size_t)row * 32 + j * 2, 1.f, 2.f);
    } else if (MODE == 4) {
      const int g = lane >> 3, j = lane & 7;
      uint32_t row = hash32(base + g) & rmask;
      *reinterpret_cast<float4*>(table + (size_t)row * 32 + j * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
    } else if (MODE == 5) {
      const int g = lane >> 2, j = lane & 3;
      uint32_t row = hash32(base + g) & rmask;
      red4(table + (size_t)row * 32 + j * 4, 1.f, 2.f, 3.f, 4.f);
      red4(table + (size_t)row * 32 + 16 + j * 4, 1.f, 2.f, 3.f, 4.f);
    }
  }
}

// TMA bulk reduce: each warp stages 4 rows (512 B) in smem, 4 lanes issue one 128-B bulk add each.
__global__ void __launch_bounds__(256) k_tma(float* table, uint32_t rmask, int iters, int rows_per_op) {
  extern __shared__ __align__(128) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr int STAGES = 4;
  float* mine = sm + warp * (STAGES * 128);   // 4 stages x 4 rows x 32 floats
  for (int it = 0; it < iters; ++it) {
    const int st = it % STAGES;
    if (it >= STAGES) {   // make sure the bulk op that read this stage has finished reading
      if (lane < 4) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(STAGES - 1) : "memory");
      __syncwarp();
    }
    const uint32_t base = hash32(wid * 7919u + it);
    float* stage = mine + st * 128;
    *reinterpret_cast<float4*>(stage + lane * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < 4) {
      uint32_t row = hash32(base + lane) & rmask;
      float* dst = table + (size_t)row * 32;
      uint32_t src = (uint32_t)__cvta_generic_to_shared(stage + lane * 32);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" :: "l"(dst), "r"(src) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// shared-memory fp32 atomicAdd (CAS loop) on a per-CTA window of `win` rows, 32 lanes per row
__global__ void __launch_bounds__(256) k_smem(float* table, int iters, int win) {
  extern __shared__ __align__(128) float sm[];
  for (int i = threadIdx.x; i < win * 32; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    uint32_t row = hash32(wid * 7919u + it) % (uint32_t)win;
    atomicAdd(&sm[row * 32 + lane], 1.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < win * 32; i += blockDim.x) atomicAdd(table + (size_t)blockIdx.x * win * 32 % (1u << 20) + i, sm[i]);
}

// shared-memory non-atomic RMW with float4 per lane (8 lanes per row): LDS.128 + FADD + STS.128 (racy across warps;
// measures the raw smem RMW rate an ownership scheme could reach)
__global__ void __launch_bounds__(256) k_smem_rmw(float* table, int iters, int win) {
  extern __shared__ __align__(128) float sm[];
  for (int i = threadIdx.x; i < win * 32; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    uint32_t row = hash32(wid * 7919u + it * 4 + g) % (uint32_t)win;
    float4* p = reinterpret_cast<float4*>(&sm[row * 32 + j * 4]);
    float4 v = *p; v.x += 1.f; v.y += 2.f; v.z += 3.f; v.w += 4.f; *p = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < win * 32; i += blockDim.x) atomicAdd(table + (size_t)blockIdx.x * win * 32 % (1u << 20) + i, sm[i]);
}

template <typename F>
float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t max_rows = 1u << 22;           // 512 MB table
  float* table; CK(cudaMalloc(&table, max_rows * 128)); CK(cudaMemset(table, 0, max_rows * 128));
  const int ctas = sms * 8, iters = 2000;
  const double warps = (double)ctas * 8;
  printf("SMs=%d ctas=%d iters=%d\n", sms, ctas, iters);
  for (uint32_t lg : {16u, 19u, 22u}) {
    const uint32_t rmask = (1u << lg) - 1;
    printf("-- table rows 2^%u (%.0f MB)\n", lg, (double)(1u << lg) * 128 / 1e6);
    auto report = [&](const char* name, float ms, double rows_per_warp_iter) {
      const double rows = warps * iters * rows_per_warp_iter;
      printf("  %-34s %8.3f ms  %7.2f Grows/s  %6.2f TB/s payload  %5.2f clk/row/SM@1.9GHz\n", name, ms, rows / ms / 1e6,
             rows * 128 / ms / 1e9, ms * 1e-3 * 1.9e9 * sms / rows);
    };
    report("RED.v4 full rows (8 lanes/row)", timeit([&] { k_red<0><<<ctas, 256>>>(table, rmask, iters, 0); }), 4);
    report("RED.32 (32 lanes/row)", timeit([&] { k_red<1><<<ctas, 256>>>(table, rmask, iters, 0); }), 1);
    report("RED.v4 half-sector x2 (bf16 now)", timeit([&] { k_red<2><<<ctas, 256>>>(table, rmask, iters, 0); }), 8);
    report("RED.v4 x2 permuted full sectors", timeit([&] { k_red<5><<<ctas, 256>>>(table, rmask, iters, 0); }), 8);
    report("RED.v2 (16 lanes/row)", timeit([&] { k_red<3><<<ctas, 256>>>(table, rmask, iters, 0); }), 2);
    report("STG.128 scatter (8 lanes/row)", timeit([&] { k_red<4><<<ctas, 256>>>(table, rmask, iters, 0); }), 4);
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4 * 512));
    report("TMA bulk reduce 128 B/op", timeit([&] { k_tma<<<ctas, 256, 8 * 4 * 512>>>(table, rmask, iters, 1); }), 4);
  }
  printf("-- locality: RED.v4 rows drawn from a per-warp window\n");
  for (uint32_t win : {16u, 64u, 512u}) {
    float ms = timeit([&] { k_red<0><<<ctas, 256>>>(table, (1u << 22) - 1, iters, win); });
    printf("  window %4u rows: %8.3f ms %7.2f Grows/s\n", win, ms, warps * iters * 4 / ms / 1e6);
  }
  printf("-- shared memory accumulation (per-CTA window)\n");
  for (int win : {256, 1024}) {
    CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, win * 128));
    CK(cudaFuncSetAttribute(k_smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, win * 128));
    int c2 = win >= 1024 ? sms : sms * 4;   // 128 KB windows: one CTA per SM
    double w2 = (double)c2 * 8;
    float ms = timeit([&] { k_smem<<<c2, 256, win * 128>>>(table, iters, win); });
    printf("  atomicAdd(float) smem, win %4d, %d CTAs: %8.3f ms %7.2f Grows/s  %5.2f clk/row/SM\n", win, c2, ms, w2 * iters / ms / 1e6, ms * 1e-3 * 1.9e9 * sms / (w2 * iters));
    ms = timeit([&] { k_smem_rmw<<<c2, 256, win * 128>>>(table, iters, win); });
    printf("  LDS.128+STS.128 RMW     win %4d, %d CTAs: %8.3f ms %7.2f Grows/s  %5.2f clk/row/SM\n", win, c2, ms, w2 * iters * 4 / ms / 1e6, ms * 1e-3 * 1.9e9 * sms / (w2 * iters * 4));
  }
  return 0;
}
