// Interpolator / accumulator glue and per-particle diagnostics for sm_100a.  Compiled with -fmad=false so each
// expression rounds exactly as the reference's scalar pipelines do.
//
// Replaces (reference tree):
//   src/sf_interface/pipeline/interpolator_array_pipeline.cc:21-135   load_interpolator
//   src/sf_interface/pipeline/clear_array_pipeline.cc:40-67           clear_accumulator_array
//   src/sf_interface/pipeline/unload_accumulator_pipeline.cc:18-144   unload_accumulator_array
//   src/species_advance/standard/pipeline/energy_p_pipeline.cc:18-115 energy_p
//   src/species_advance/standard/pipeline/center_p_pipeline.cc:17-96, uncenter_p_pipeline.cc:17-98
// All are streaming, HBM-bound kernels: one thread per voxel (x fastest, so warps read consecutive field_t /
// accumulator_t records) or per particle, 128-bit accesses throughout.
#include "vpb_common.cuh"

namespace vpb {


// ---- load_interpolator -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) load_interpolator_kernel(float *__restrict__ interp, int istride,
                                                                const float *__restrict__ fld, int nx, int ny, int nz) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int y = blockIdx.y + 1, z = blockIdx.z + 1;
  if (x > nx) return;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2);
  const int v = voxel(x, y, z, nx, ny);
  const float4 *f = reinterpret_cast<const float4 *>(fld);
  // float4 #0 of field_t = {ex,ey,ez,div_e_err}, #1 = {cbx,cby,cbz,div_b_err}
  const float4 e0 = __ldg(f + 5 * (size_t)v), b0 = __ldg(f + 5 * (size_t)v + 1);
  const float4 ex_ = __ldg(f + 5 * (size_t)(v + 1)), bx_ = __ldg(f + 5 * (size_t)(v + 1) + 1);
  const float4 ey_ = __ldg(f + 5 * (size_t)(v + sy)), by_ = __ldg(f + 5 * (size_t)(v + sy) + 1);
  const float4 ez_ = __ldg(f + 5 * (size_t)(v + sz)), bz_ = __ldg(f + 5 * (size_t)(v + sz) + 1);
  const float4 eyz = __ldg(f + 5 * (size_t)(v + sy + sz));
  const float4 ezx = __ldg(f + 5 * (size_t)(v + sz + 1));
  const float4 exy = __ldg(f + 5 * (size_t)(v + 1 + sy));
  const float fourth = 0.25f, half = 0.5f;
  float w0, w1, w2, w3;
  float4 o;
  float4 *out = reinterpret_cast<float4 *>(interp + (size_t)v * istride);
  w0 = e0.x; w1 = ey_.x; w2 = ez_.x; w3 = eyz.x;                       // ex: neighbours +y, +z, +y+z
  o.x = fourth * ((w3 + w0) + (w1 + w2)); o.y = fourth * ((w3 - w0) + (w1 - w2));
  o.z = fourth * ((w3 - w0) - (w1 - w2)); o.w = fourth * ((w3 + w0) - (w1 + w2));
  out[0] = o;
  w0 = e0.y; w1 = ez_.y; w2 = ex_.y; w3 = ezx.y;                       // ey: +z, +x, +z+x
  o.x = fourth * ((w3 + w0) + (w1 + w2)); o.y = fourth * ((w3 - w0) + (w1 - w2));
  o.z = fourth * ((w3 - w0) - (w1 - w2)); o.w = fourth * ((w3 + w0) - (w1 + w2));
  out[1] = o;
  w0 = e0.z; w1 = ex_.z; w2 = ey_.z; w3 = exy.z;                       // ez: +x, +y, +x+y
  o.x = fourth * ((w3 + w0) + (w1 + w2)); o.y = fourth * ((w3 - w0) + (w1 - w2));
  o.z = fourth * ((w3 - w0) - (w1 - w2)); o.w = fourth * ((w3 + w0) - (w1 + w2));
  out[2] = o;
  o.x = half * (bx_.x + b0.x); o.y = half * (bx_.x - b0.x);             // cbx, dcbxdx
  o.z = half * (by_.y + b0.y); o.w = half * (by_.y - b0.y);             // cby, dcbydy
  out[3] = o;
  float2 o2; o2.x = half * (bz_.z + b0.z); o2.y = half * (bz_.z - b0.z);   // cbz, dcbzdz (padding untouched)
  *reinterpret_cast<float2 *>(out + 4) = o2;
}

// ---- unload_accumulator ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unload_accumulator_kernel(float *__restrict__ fld, const float *__restrict__ acc,
                                                                 int astride, int nx, int ny, int nz,
                                                                 float cx, float cy, float cz) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int y = blockIdx.y + 1, z = blockIdx.z + 1;
  if (x > nx + 1) return;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2);
  const int v = voxel(x, y, z, nx, ny);
#define A4(vv, k) __ldg(reinterpret_cast<const float4 *>(acc + (size_t)(vv) * astride) + (k))
  const float4 a0x = A4(v, 0), a0y = A4(v, 1), a0z = A4(v, 2);
  const float4 ayx = A4(v - sy, 0), azx = A4(v - sz, 0), ayzx = A4(v - sy - sz, 0);          // jx of -y, -z, -y-z
  const float4 azy = A4(v - sz, 1), axy = A4(v - 1, 1), azxy = A4(v - 1 - sz, 1);            // jy of -z, -x, -z-x
  const float4 axz = A4(v - 1, 2), ayz = A4(v - sy, 2), axyz = A4(v - 1 - sy, 2);            // jz of -x, -y, -x-y
#undef A4
  float4 *f3 = reinterpret_cast<float4 *>(fld) + 5 * (size_t)v + 3;   // {jfx,jfy,jfz,rhof}
  float4 j = *f3;
  j.x += cx * (((a0x.x + ayx.y) + azx.z) + ayzx.w);
  j.y += cy * (((a0y.x + azy.y) + axy.z) + azxy.w);
  j.z += cz * (((a0z.x + axz.y) + ayz.z) + axyz.w);
  *f3 = j;
}

// ---- per-particle helpers ----------------------------------------------------------------------------------
struct Interp { float4 ex, ey, ez, b0; float2 b1; };
__device__ __forceinline__ Interp load_interp(const float *interp, int istride, int ii) {
  const float4 *f = reinterpret_cast<const float4 *>(interp + (size_t)ii * istride);
  Interp r; r.ex = __ldg(f); r.ey = __ldg(f + 1); r.ez = __ldg(f + 2); r.b0 = __ldg(f + 3);
  r.b1 = __ldg(reinterpret_cast<const float2 *>(f + 4));
  return r;
}

// energy_p: one double per CTA, then one atomicAdd(double) per CTA into *en (pre-zeroed), finally scaled by cvac^2
__global__ void __launch_bounds__(256) energy_p_kernel(const float4 *__restrict__ p, int np, const float *__restrict__ interp,
                                                       int istride, float qdt_2mc, float msp, double *en) {
  __shared__ double s_part[8];
  double acc = 0.0;
  const float one = 1.0f;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < np; n += gridDim.x * blockDim.x) {
    const float4 r = p[2 * (size_t)n], u = p[2 * (size_t)n + 1];
    const Interp f = load_interp(interp, istride, __float_as_int(r.w));
    const float dx = r.x, dy = r.y, dz = r.z;
    float v0 = u.x + qdt_2mc * ((f.ex.x + dy * f.ex.y) + dz * (f.ex.z + dy * f.ex.w));
    float v1 = u.y + qdt_2mc * ((f.ey.x + dz * f.ey.y) + dx * (f.ey.z + dz * f.ey.w));
    float v2 = u.z + qdt_2mc * ((f.ez.x + dx * f.ez.y) + dy * (f.ez.z + dx * f.ez.w));
    v0 = (v0 * v0 + v1 * v1) + v2 * v2;
    v0 = (msp * u.w) * __fdiv_rn(v0, one + __fsqrt_rn(one + v0));
    acc += (double)v0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w];
    atomicAdd(en, t);
  }
}
__global__ void scale_double_kernel(double *x, double s) { *x *= s; }

// center_p (FORWARD=true): half E kick then half Boris rotation; uncenter_p: inverse half rotation then inverse kick
template <bool CENTER>
__global__ void __launch_bounds__(256) center_p_kernel(float4 *__restrict__ p, int np, const float *__restrict__ interp,
                                                       int istride, float qdt_2mc_in) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= np) return;
  const float qdt_2mc = CENTER ? qdt_2mc_in : -qdt_2mc_in;
  const float qdt_4mc = 0.5f * qdt_2mc;
  const float one = 1.0f, one_third = (float)(1.0 / 3.0), two_fifteenths = (float)(2.0 / 15.0);
  const float4 r = p[2 * (size_t)n];
  float4 u = p[2 * (size_t)n + 1];
  const Interp f = load_interp(interp, istride, __float_as_int(r.w));
  const float dx = r.x, dy = r.y, dz = r.z;
  const float hax = qdt_2mc * ((f.ex.x + dy * f.ex.y) + dz * (f.ex.z + dy * f.ex.w));
  const float hay = qdt_2mc * ((f.ey.x + dz * f.ey.y) + dx * (f.ey.z + dz * f.ey.w));
  const float haz = qdt_2mc * ((f.ez.x + dx * f.ez.y) + dy * (f.ez.z + dx * f.ez.w));
  const float cbx = f.b0.x + dx * f.b0.y, cby = f.b0.z + dy * f.b0.w, cbz = f.b1.x + dz * f.b1.y;
  float ux = u.x, uy = u.y, uz = u.z;
  if (CENTER) { ux += hax; uy += hay; uz += haz; }
  // the reference takes a double sqrt of the float argument here and rounds it to float before dividing
  float v0 = __fdiv_rn(qdt_4mc, (float)sqrt((double)(one + (ux * ux + (uy * uy + uz * uz)))));
  float v1 = cbx * cbx + (cby * cby + cbz * cbz);
  float v2 = (v0 * v0) * v1;
  float v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
  float v4 = __fdiv_rn(v3, one + v1 * (v3 * v3));
  v4 += v4;
  v0 = ux + v3 * (uy * cbz - uz * cby);
  v1 = uy + v3 * (uz * cbx - ux * cbz);
  v2 = uz + v3 * (ux * cby - uy * cbx);
  ux += v4 * (v1 * cbz - v2 * cby);
  uy += v4 * (v2 * cbx - v0 * cbz);
  uz += v4 * (v0 * cby - v1 * cbx);
  if (!CENTER) { ux += hax; uy += hay; uz += haz; }
  u.x = ux; u.y = uy; u.z = uz;
  p[2 * (size_t)n + 1] = u;
}

// accumulate_rho_p (rho_p.cc:22-113): trilinear node charges of every particle into field_t.rhof (float 15 of the
// 20-float field record).  One thread per particle; lanes of a warp that share a voxel (all of them, for voxel-sorted
// particles) sum their eight node charges with a reduce-scatter and issue eight REDs for the group instead of eight
// per lane; lanes in groups of fewer than four go straight to memory.
__global__ void __launch_bounds__(256) accumulate_rho_p_kernel(float *__restrict__ fld, const float4 *__restrict__ p,
                                                               int np, float q_8V, int sy, int sz) {
  const int lane = threadIdx.x & 31;
  const long long rows = ((long long)np + 31) / 32;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long n = row * 32 + lane;
    const bool valid = n < np;
    float w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int v = 0;
    if (valid) {
      const float4 r = p[2 * n];
      const float pw = p[2 * n + 1].w;
      float w0 = r.x, w1 = r.y, w2, w3, w4, w5, w6, w7 = pw * q_8V;
      const float dz = r.z;
      v = __float_as_int(r.w);
      w6 = w7 - w0 * w7; w7 = w7 + w0 * w7;
      w4 = w6 - w1 * w6; w5 = w7 - w1 * w7;
      w6 = w6 + w1 * w6; w7 = w7 + w1 * w7;
      w0 = w4 - dz * w4; w1 = w5 - dz * w5; w2 = w6 - dz * w6; w3 = w7 - dz * w7;
      w4 = w4 + dz * w4; w5 = w5 + dz * w5; w6 = w6 + dz * w6; w7 = w7 + dz * w7;
      w[0] = w0; w[1] = w1; w[2] = w2; w[3] = w3; w[4] = w4; w[5] = w5; w[6] = w6; w[7] = w7;
    }
    const unsigned peers = warp_peers(valid, v);
    const bool grouped = valid && __popc(peers) >= 4;
    float *f = fld + 15;
    if (valid && !grouped) {
      red_add(f + 20 * (size_t)(v), w[0]);           red_add(f + 20 * (size_t)(v + 1), w[1]);
      red_add(f + 20 * (size_t)(v + sy), w[2]);      red_add(f + 20 * (size_t)(v + sy + 1), w[3]);
      red_add(f + 20 * (size_t)(v + sz), w[4]);      red_add(f + 20 * (size_t)(v + sz + 1), w[5]);
      red_add(f + 20 * (size_t)(v + sz + sy), w[6]); red_add(f + 20 * (size_t)(v + sz + sy + 1), w[7]);
    }
    unsigned big = __ballot_sync(0xffffffffu, grouped);
    while (big) {
      const int leader = __ffs(big) - 1;
      const unsigned grp = __shfl_sync(0xffffffffu, peers, leader);
      const int gv = __shfl_sync(0xffffffffu, v, leader);
      const bool mine = grouped && peers == grp;
      float t[8];
#pragma unroll
      for (int c = 0; c < 8; c++) t[c] = mine ? w[c] : 0.0f;
      warp_reduce_scatter<8>(t);
      const int c = lane >> 2;                                     // node c = (c&1) + sy*(c>>1&1) + sz*(c>>2)
      if ((lane & 3) == 0) red_add(f + 20 * (size_t)(gv + (c & 1) + ((c & 2) ? sy : 0) + ((c & 4) ? sz : 0)), t[0]);
      big &= ~grp;
    }
  }
}

}  // namespace vpb

using namespace vpb;

extern "C" int vpb_accumulate_rho_p(float *fields, const void *p, int32_t np, float q, float r8V,
                                    int32_t nx, int32_t ny, int32_t nz, void *stream) {
  VPB_REQUIRE(fields && (p || np == 0) && nx > 0 && ny > 0 && nz > 0, "vpb_accumulate_rho_p: Bad args");
  if (np <= 0) return 0;
  const float q_8V = q * r8V;
  long long grid = (((long long)np + 31) / 32 + 7) / 8; if (grid > kSMs * 16) grid = kSMs * 16;
  accumulate_rho_p_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(fields, (const float4 *)p, np, q_8V,
                                                                   nx + 2, (nx + 2) * (ny + 2));
  VPB_LAUNCH_CHECK();
  return 0;
}


extern "C" int vpb_load_interpolator(float *interp, int32_t interp_stride, const float *fields,
                                     int32_t nx, int32_t ny, int32_t nz, void *stream) {
  VPB_REQUIRE(interp && fields && nx > 0 && ny > 0 && nz > 0 && interp_stride >= 18 && interp_stride % 4 == 0,
              "vpb_load_interpolator: Bad args");
  dim3 grid((nx + 255) / 256, ny, nz);
  VPB_REQUIRE(ny <= 65535 && nz <= 65535, "vpb_load_interpolator: grid too large");
  load_interpolator_kernel<<<grid, 256, 0, as_stream(stream)>>>(interp, interp_stride, fields, nx, ny, nz);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_clear_accumulator(float *accum, int32_t accum_stride, int32_t nx, int32_t ny, int32_t nz, void *stream) {
  VPB_REQUIRE(accum && nx > 0 && ny > 0 && nz > 0, "vpb_clear_accumulator: Bad args.");
  // same voxel window as the reference: [VOXEL(1,1,1) rounded down to even, VOXEL(nx,ny,nz)] rounded up to even length
  const int i0 = (voxel(1, 1, 1, nx, ny) / 2) * 2;
  const int na = (((voxel(nx, ny, nz, nx, ny) - i0 + 1) + 1) / 2) * 2;
  VPB_CUDA(cudaMemsetAsync(accum + (size_t)i0 * accum_stride, 0, (size_t)na * accum_stride * sizeof(float), as_stream(stream)));
  count_launch();
  return 0;
}

extern "C" int vpb_unload_accumulator(float *fields, const float *accum, int32_t accum_stride,
                                      int32_t nx, int32_t ny, int32_t nz,
                                      float rdx, float rdy, float rdz, float dt, void *stream) {
  VPB_REQUIRE(fields && accum && nx > 0 && ny > 0 && nz > 0, "vpb_unload_accumulator: Bad args");
  // the reference evaluates 0.25*rdy*rdz/dt in double and stores float (unload_accumulator_pipeline.cc:137-139)
  const float cx = (float)(0.25 * (double)rdy * (double)rdz / (double)dt);
  const float cy = (float)(0.25 * (double)rdz * (double)rdx / (double)dt);
  const float cz = (float)(0.25 * (double)rdx * (double)rdy / (double)dt);
  dim3 grid((nx + 1 + 255) / 256, ny + 1, nz + 1);
  VPB_REQUIRE(ny + 1 <= 65535 && nz + 1 <= 65535, "vpb_unload_accumulator: grid too large");
  unload_accumulator_kernel<<<grid, 256, 0, as_stream(stream)>>>(fields, accum, accum_stride, nx, ny, nz, cx, cy, cz);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_energy_p(const void *p, int32_t np, const float *interp, int32_t interp_stride,
                            float q, float m, float dt, float cvac, double *en_dev, void *stream) {
  VPB_REQUIRE(en_dev && interp && (p || np == 0), "vpb_energy_p: Bad args");
  cudaStream_t st = as_stream(stream);
  VPB_CUDA(cudaMemsetAsync(en_dev, 0, sizeof(double), st));
  if (np > 0) {
    const float qdt_2mc = (q * dt) / (2 * m * cvac);
    int grid = (np + 255) / 256; if (grid > kSMs * 8) grid = kSMs * 8;
    energy_p_kernel<<<grid, 256, 0, st>>>((const float4 *)p, np, interp, interp_stride, qdt_2mc, m, en_dev);
    VPB_LAUNCH_CHECK();
    scale_double_kernel<<<1, 1, 0, st>>>(en_dev, (double)cvac * (double)cvac);
    VPB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int vpb_center_p(void *p, int32_t np, const float *interp, int32_t interp_stride, float qdt_2mc, void *stream) {
  VPB_REQUIRE(interp && (p || np == 0), "vpb_center_p: Bad args.");
  if (np <= 0) return 0;
  center_p_kernel<true><<<(np + 255) / 256, 256, 0, as_stream(stream)>>>((float4 *)p, np, interp, interp_stride, qdt_2mc);
  VPB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vpb_uncenter_p(void *p, int32_t np, const float *interp, int32_t interp_stride, float qdt_2mc, void *stream) {
  VPB_REQUIRE(interp && (p || np == 0), "vpb_uncenter_p: Bad args.");
  if (np <= 0) return 0;
  center_p_kernel<false><<<(np + 255) / 256, 256, 0, as_stream(stream)>>>((float4 *)p, np, interp, interp_stride, qdt_2mc);
  VPB_LAUNCH_CHECK();
  return 0;
}
