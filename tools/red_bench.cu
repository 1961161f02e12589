// Micro-benchmark: how fast can an SM push 48-byte (12 x f32) accumulator increments into L2?
//   mode 0: three red.global.add.v4.f32 per lane (what advance_p's mover phase does)
//   mode 1: stage 48 B per lane in shared memory, one cp.reduce.async.bulk (TMA reduce-add) per lane
//   mode 2: twelve scalar red.global.add.f32 per lane
// Targets: per-lane pseudo-random voxels within a window (spread) of a 105 MB accumulator array.
// Build+run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/red_bench tools/red_bench.cu && /tmp/red_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_v4(float *a, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *acc, int nvox, int iters, int window) {
  __shared__ __align__(16) float4 stage[8][3][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
  int base = (int)(((uint64_t)blockIdx.x * 7919u) % (uint64_t)(nvox - window - 1));
  for (int it = 0; it < iters; it++) {
    s = s * 1664525u + 1013904223u;
    const int v = base + (int)((s >> 8) % (uint32_t)window);
    float *a = acc + 12 * (size_t)v;
    const float x = 1e-6f * (float)(lane + it);
    if (MODE == 0) {
      red_v4(a, x, x, x, x); red_v4(a + 4, x, x, x, x); red_v4(a + 8, x, x, x, x);
    } else if (MODE == 2) {
#pragma unroll
      for (int c = 0; c < 12; c++) asm volatile("red.global.add.f32 [%0], %1;" :: "l"(a + c), "f"(x) : "memory");
    } else {
      // wait until the TMA engine has read the previous contents of the staging buffer
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      stage[w][0][lane] = make_float4(x, x, x, x);
      stage[w][1][lane] = make_float4(x, x, x, x);
      stage[w][2][lane] = make_float4(x, x, x, x);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      // plane-major staging keeps the STS conflict-free; one 16-byte reduce per plane would be 3 ops, so instead
      // each lane owns 48 contiguous bytes in a second, lane-major view:
      // (measure both: ops of 16 B x3 vs one op of 48 B needs lane-major layout)
      float *src = reinterpret_cast<float *>(&stage[w][0][0]) + 0;   // placeholder, see MODE 3
      (void)src;
#pragma unroll
      for (int pl = 0; pl < 3; pl++)
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 16;"
                     :: "l"(a + 4 * pl), "r"(smem_u32(&stage[w][pl][lane])) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    base += 3; if (base > nvox - window - 2) base = 0;
  }
  if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// MODE 3: lane-major staging (48 contiguous bytes per lane), ONE 48-byte bulk reduce per lane
__global__ void __launch_bounds__(256) k3(float *acc, int nvox, int iters, int window) {
  __shared__ __align__(16) float stage[8][32][12];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
  int base = (int)(((uint64_t)blockIdx.x * 7919u) % (uint64_t)(nvox - window - 1));
  for (int it = 0; it < iters; it++) {
    s = s * 1664525u + 1013904223u;
    const int v = base + (int)((s >> 8) % (uint32_t)window);
    float *a = acc + 12 * (size_t)v;
    const float x = 1e-6f * (float)(lane + it);
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    float4 *d = reinterpret_cast<float4 *>(&stage[w][lane][0]);
    d[0] = make_float4(x, x, x, x); d[1] = make_float4(x, x, x, x); d[2] = make_float4(x, x, x, x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 48;"
                 :: "l"(a), "r"(smem_u32(d)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    base += 3; if (base > nvox - window - 2) base = 0;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int nvox = 2197000, iters = 200;
  float *acc; cudaMalloc(&acc, 12ull * nvox * 4); cudaMemset(acc, 0, 12ull * nvox * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 4;
  for (int window : {4, 64, 4096}) {
    for (int mode = 0; mode < 4; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, 256>>>(acc, nvox, iters, window);
        else if (mode == 1) k<1><<<grid, 256>>>(acc, nvox, iters, window);
        else if (mode == 2) k<2><<<grid, 256>>>(acc, nvox, iters, window);
        else k3<<<grid, 256>>>(acc, nvox, iters, window);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        if (rep == 1) {
          const double deposits = (double)grid * 256 * iters;
          printf("window %5d mode %d: %8.3f ms  %7.2f G deposits(48B)/s  %6.1f cycles/warp-deposit/SM  %s\n", window, mode, ms,
                 deposits / ms / 1e6, ms * 1e-3 * 1.9e9 / (deposits / 32 / 148), err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
      }
    }
  }
  // checksum so the work is not optimised away and to verify all modes add the same total
  return 0;
}
