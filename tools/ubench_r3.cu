// Round-2 (second session) micro-benchmarks: what a straggler's 48-byte accumulator increment and a drifted
// particle's 72-byte interpolator gather cost on the L1 data pipe, and whether lanes that share a 32-byte sector
// are merged.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/ubench3 tools/ubench_r3.cu && /tmp/ubench3
//   E  three red.global.add.v4.f32 per lane to the lane's own voxel (stride 48 B), addresses drawn from the
//      neighbourhood a drifted row really touches (+-2 cells around a base voxel that walks along the array)
//      E0 as the kernel does it | E1 lane pairs arranged so that two of the three instructions write both halves of
//      one 32-byte sector | E2 only even lanes active | E3 one RED.v4 per lane | E4 stride 64 B (one voxel = 2 sectors)
//   F  interpolator gather for the same address pattern
//      F0 4 x LDG.128 + LDG.64 (read-only path) | F1 3 aligned 256-bit sector loads | F2 LDS from a CTA-shared tile
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_v4(float *a, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
constexpr int SY = 130, SZ = 130 * 130, NVOX = 130 * 130 * 130;

// voxel of this lane: base walks along x; the lane's particle has drifted up to +-R cells on every axis
__device__ __forceinline__ int drifted_voxel(uint32_t &s, int base, int R) {
  s = s * 1664525u + 1013904223u;
  const uint32_t h = s >> 8;
  const int w = 2 * R + 1;
  const int dx = (int)(h % w) - R, dy = (int)((h / w) % w) - R, dz = (int)((h / (w * w)) % w) - R;
  return base + dx + SY * dy + SZ * dz;
}

template <int MODE>
__global__ void __launch_bounds__(256) red_bench(float *acc, int iters, int R, int stride) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  uint32_t s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
  int base = 3 * SZ + 3 * SY + 3 + (int)(((uint64_t)gw * 7919u) % (uint64_t)(NVOX - 8 * SZ));
  for (int it = 0; it < iters; it++) {
    const int v = drifted_voxel(s, base, R);
    const float x = 1e-6f * (float)(lane + it);
    float *a = acc + (size_t)stride * v;
    if (MODE == 0 || MODE == 4) {
      red_v4(a, x, x, x, x); red_v4(a + 4, x, x, x, x); red_v4(a + 8, x, x, x, x);
    } else if (MODE == 1) {
      // pair (A = even lane, B = odd lane).  A voxel's 48 bytes are one full 32-byte sector plus half of a shared one:
      // even voxel -> full sector first, odd voxel -> full sector last.
      const int vo = __shfl_xor_sync(0xffffffffu, v, 1);
      const float y0 = __shfl_xor_sync(0xffffffffu, x, 1), y1 = __shfl_xor_sync(0xffffffffu, x + 1.f, 1);
      const float y2 = __shfl_xor_sync(0xffffffffu, x + 2.f, 1), y3 = __shfl_xor_sync(0xffffffffu, x + 3.f, 1);
      const int va = (lane & 1) ? vo : v, vb = (lane & 1) ? v : vo;          // voxels of lane A and lane B of this pair
      float *fa = acc + (size_t)12 * va + ((va & 1) ? 4 : 0);                // A's full sector (8 floats)
      float *fb = acc + (size_t)12 * vb + ((vb & 1) ? 4 : 0);                // B's full sector
      const int half = (lane & 1) ? 4 : 0;
      red_v4(fa + half, x, y0, y1, y2);                                      // both lanes of the pair: one sector
      red_v4(fb + half, y3, x, y0, y1);
      red_v4(a + ((v & 1) ? 0 : 8), x, x, x, x);                             // own leftover half sector
    } else if (MODE == 2) {
      if (!(lane & 1)) { red_v4(a, x, x, x, x); red_v4(a + 4, x, x, x, x); red_v4(a + 8, x, x, x, x); }
    } else if (MODE == 3) {
      red_v4(a, x, x, x, x);
    }
    if (it & 1) base++;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) gather_bench(const float *interp, float *out, int iters, int R) {
  extern __shared__ __align__(16) float tile[];                              // MODE 2: 512 voxels x 20 floats
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  uint32_t s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
  int base = 3 * SZ + 3 * SY + 3 + (int)(((uint64_t)gw * 7919u) % (uint64_t)(NVOX - 8 * SZ));
  if (MODE == 2) {
    for (int k = threadIdx.x; k < 512 * 20; k += 256) tile[k] = interp[k];
    __syncthreads();
  }
  float sum = 0.f;
  for (int it = 0; it < iters; it++) {
    const int v = drifted_voxel(s, base, R);
    if (MODE == 0) {
      const float4 *f = reinterpret_cast<const float4 *>(interp + (size_t)20 * v);
      const float4 a = __ldg(f), b = __ldg(f + 1), c = __ldg(f + 2), d = __ldg(f + 3);
      const float2 e = __ldg(reinterpret_cast<const float2 *>(f + 4));
      sum += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) + ((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)) + (e.x + e.y);
    } else if (MODE == 1) {
      // the 72 used bytes of a voxel lie in three consecutive aligned 32-byte sectors
      const char *p = reinterpret_cast<const char *>(interp) + (((size_t)80 * v) & ~(size_t)31);
      float r[24];
#pragma unroll
      for (int k = 0; k < 3; k++)
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r[8 * k]), "=f"(r[8 * k + 1]), "=f"(r[8 * k + 2]), "=f"(r[8 * k + 3]), "=f"(r[8 * k + 4]),
                       "=f"(r[8 * k + 5]), "=f"(r[8 * k + 6]), "=f"(r[8 * k + 7]) : "l"(p + 32 * k));
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 18; k++) q += r[k + ((v & 1) ? 4 : 0)];
      sum += q;
    } else {
      const uint32_t t = (uint32_t)(v - base + 2 * SZ + 2 * SY + 2) % 512u;   // any slot of the tile: timing only
      const float4 *f = reinterpret_cast<const float4 *>(tile + 20 * t);
      const float4 a = f[0], b = f[1], c = f[2], d = f[3];
      const float2 e = *reinterpret_cast<const float2 *>(f + 4);
      sum += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) + ((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)) + (e.x + e.y);
    }
    if (it & 1) base++;
  }
  if (sum == 1.2345e30f) out[0] = sum;
}

template <class F> static float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return best;
}

int main() {
  float *acc, *interp, *out;
  cudaMalloc(&acc, (size_t)NVOX * 16 * sizeof(float)); cudaMemset(acc, 0, (size_t)NVOX * 16 * sizeof(float));
  cudaMalloc(&interp, (size_t)NVOX * 20 * sizeof(float)); cudaMemset(interp, 0, (size_t)NVOX * 20 * sizeof(float));
  cudaMalloc(&out, 64);
  const int iters = 1024, grid = 148 * 4;                                    // 32 warps per SM like the push kernel
  const double rows = (double)grid * 8 * iters;
  for (int R : {0, 1, 2}) {
    float t0 = time_ms([&] { red_bench<0><<<grid, 256>>>(acc, iters, R, 12); });
    float t1 = time_ms([&] { red_bench<1><<<grid, 256>>>(acc, iters, R, 12); });
    float t2 = time_ms([&] { red_bench<2><<<grid, 256>>>(acc, iters, R, 12); });
    float t3 = time_ms([&] { red_bench<3><<<grid, 256>>>(acc, iters, R, 12); });
    float t4 = time_ms([&] { red_bench<4><<<grid, 256>>>(acc, iters, R, 16); });
    printf("E drift +-%d: 3xRED.v4/lane %.1f G incr/s | sector-paired %.1f | even lanes only %.1f | one RED.v4/lane %.1f G/s | stride 64 B %.1f\n",
           R, rows * 32 / t0 * 1e-6, rows * 32 / t1 * 1e-6, rows * 16 / t2 * 1e-6, rows * 32 / t3 * 1e-6, rows * 32 / t4 * 1e-6);
  }
  cudaFuncSetAttribute(gather_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 80);
  for (int R : {0, 1, 2}) {
    float t0 = time_ms([&] { gather_bench<0><<<grid, 256>>>(interp, out, iters, R); });
    float t1 = time_ms([&] { gather_bench<1><<<grid, 256>>>(interp, out, iters, R); });
    float t2 = time_ms([&] { gather_bench<2><<<grid, 256, 512 * 80>>>(interp, out, iters, R); });
    printf("F drift +-%d: 5 x LDG %.1f G gathers/s | 3 x LDG.256 %.1f | LDS from a shared tile %.1f\n",
           R, rows * 32 / t0 * 1e-6, rows * 32 / t1 * 1e-6, rows * 32 / t2 * 1e-6);
  }
  return 0;
}
