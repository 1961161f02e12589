// Shared helpers for libvpic_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vpic_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvpic_b200 is written for sm_100a (B200) only"
#endif

namespace vpb {

constexpr int kSMs = 148;                  // B200: 2 dies x 74 SMs

void set_error(const char *fmt, ...);
int  check_cuda(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);

#define VPB_CUDA(call) do { int _r = vpb::check_cuda((call), #call, __FILE__, __LINE__); if (_r) return _r; } while (0)
#define VPB_LAUNCH_CHECK() do { vpb::count_launch(); VPB_CUDA(cudaGetLastError()); } while (0)
#define VPB_REQUIRE(cond, ...) do { if (!(cond)) { vpb::set_error(__VA_ARGS__); return -1; } } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// VOXEL macro of the reference (src/grid/grid.h:136)
__host__ __device__ __forceinline__ int voxel(int x, int y, int z, int nx, int ny) {
  return x + (nx + 2) * (y + (ny + 2) * z);
}

// red.global.add.v4.f32 (sm_90+): one 16-byte reduction instead of four scalar ones
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float *addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" :: "l"(addr), "f"(a) : "memory");
}

// Warp-wide sum of N values per lane by reduce-scatter (N = 8 or 16): every butterfly step halves the number of values
// a lane carries, so N values cost N-1 shuffles (plus log2(32/N) for the tail) instead of 5N.  Afterwards the total of
// value c sits in v[0] of the 32/N lanes with lane / (32/N) == c.
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N]) {
  static_assert(N == 8 || N == 16, "N");
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = N / 2, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool hi = (lane & bit) != 0;
#pragma unroll
    for (int c = 0; c < half; c++) {
      const float keep = hi ? v[c + half] : v[c];
      const float send = hi ? v[c] : v[c + half];
      v[c] = keep + __shfl_xor_sync(full, send, bit);
    }
  }
#pragma unroll
  for (int bit = 16 / N; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(full, v[0], bit);
}

// Lanes of a warp that share a key, found cheaply when the whole warp shares one (voxel-sorted particles): returns the
// peer mask of this lane (0 for inactive lanes).
__device__ __forceinline__ unsigned warp_peers(bool active, int key) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned amask = __ballot_sync(full, active);
  if (amask == 0) return 0u;
  const int first = __shfl_sync(full, key, __ffs(amask) - 1);
  const bool uniform = __all_sync(full, !active || key == first);
  if (uniform) return active ? amask : 0u;
  const unsigned m = __match_any_sync(full, active ? key : (-1 - lane));
  return active ? m : 0u;
}

// 256-bit global accesses (LDG.E.256 / STG.E.256, new on sm_100): one particle_t per instruction, so a warp touches
// 1 KB of consecutive bytes with every sector fully used.  The address must be 32-byte aligned.
__device__ __forceinline__ void ld_particle(const float4 *p, float4 &r, float4 &u) {
  // particles are touched once per step: do not let them evict the interpolator lines from L1
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w), "=f"(u.x), "=f"(u.y), "=f"(u.z), "=f"(u.w) : "l"(p));
}
__device__ __forceinline__ void st_particle(float4 *p, const float4 &r, const float4 &u) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(p), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w), "f"(u.x), "f"(u.y), "f"(u.z), "f"(u.w) : "memory");
}

// streaming 128-bit accesses that do not allocate in L1
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace vpb
