// Hydro moments on the device: accumulate_hydro_p and synchronize_hydro_array, the diagnostics behind the reference's
// hydro dumps (src/vpic/dump.cc:229-270).  On the host a hydro dump walks every particle; with the particles resident
// in HBM that would mean pulling the whole particle array (8.6 GB at C2) across PCIe per dump.  Here only the hydro
// array (nv x 64 B) ever leaves the device.
//
// Replaces (reference tree):
//   src/species_advance/standard/pipeline/hydro_p_pipeline.cc:19-252   accumulate_hydro_p
//   src/sf_interface/hydro_array.cc:131-309                            synchronize_hydro_array (walls, periodic folds)
//   src/sf_interface/clear_array.cc / reduce_array.cc                  clear_hydro_array; reduce is the identity here
// Per-particle arithmetic follows the scalar pipeline (-fmad=false) up to the moments per unit weight; the node sums are
// formed per voxel group as a small matrix product out of shared memory (see the kernel) and added with vector REDs, so
// node values agree with the reference to fp32 summation-order tolerance.
#include "field_common.cuh"

namespace vpb {

constexpr int kHydroFloats = 16;
constexpr int kHydroWarps = 8;

// One row of 32 particles at a time per warp.  Node k of a particle receives w_k x {14 moments}, so the contribution of a
// group of P lanes that share a voxel to its 8 nodes is the product of a [8 x P] weight matrix with a [P x 16] moment
// matrix.  The warp stages both in shared memory (24 floats per lane) and every lane sums four consecutive moments of
// one node over the members of the group (two shared-memory loads and four FMAs per member), then issues one 16-byte
// RED.  (The first version reduced 8 x 16 values across the warp with shuffles: ~1000 instructions per row and 112
// scalar REDs per group, 8.1 ms per 134 M particles; this one takes 5.2 ms right after a sort and is bound by the
// shared-memory pipe — 1.2 G wavefronts per launch, ncu — and back at 8.8 ms seven steps later.  Two re-blockings of
// the inner loop, fixed trip count and two nodes per lane, measured no faster.)  The products are formed as w_k * (q v) instead of the reference's
// (q w_k) * v and summed with FMAs: node values agree with the reference to fp32 summation-order tolerance, as before.
__global__ void __launch_bounds__(32 * kHydroWarps, 4) accumulate_hydro_p_kernel(float *__restrict__ hydro, const float4 *__restrict__ p, int np,
                                                                 const float *__restrict__ interp, int istride,
                                                                 float qsp, float mspc, float c, float qdt_2mc, float qdt_4mc2,
                                                                 float r8V, int sy, int sz) {
  __shared__ __align__(16) float s_w[kHydroWarps][32][8];
  __shared__ __align__(16) float s_m[kHydroWarps][32][16];
  const float one = 1.0f, one_third = (float)(1.0 / 3.0);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int node = lane >> 2, quad = lane & 3;                   // this lane sums four consecutive moments of one node
  const int node_off = (node & 1) + ((node & 2) ? sy : 0) + ((node & 4) ? sz : 0);
  const long long rows = ((long long)np + 31) / 32;
  for (long long row = (long long)blockIdx.x * kHydroWarps + w; row < rows; row += (long long)gridDim.x * kHydroWarps) {
    const long long n = row * 32 + lane;
    const bool valid = n < np;
    int vox = -1 - lane;
    if (valid) {
      const float4 r = p[2 * n], u = p[2 * n + 1];
      float dx = r.x, dy = r.y, dz = r.z;
      vox = __float_as_int(r.w);
      const float4 *f = reinterpret_cast<const float4 *>(interp + (size_t)vox * istride);
      const float4 fex = __ldg(f), fey = __ldg(f + 1), fez = __ldg(f + 2), fb0 = __ldg(f + 3);
      const float2 fb1 = __ldg(reinterpret_cast<const float2 *>(f + 4));
      float ux = u.x, uy = u.y, uz = u.z;
      ux += qdt_2mc * ((fex.x + dy * fex.y) + dz * (fex.z + dy * fex.w));        // hydro_p_pipeline.cc:88-95
      uy += qdt_2mc * ((fey.x + dz * fey.y) + dx * (fey.z + dz * fey.w));
      uz += qdt_2mc * ((fez.x + dx * fez.y) + dy * (fez.z + dx * fez.w));
      float w5 = fb0.x + dx * fb0.y, w6 = fb0.z + dy * fb0.w, w7 = fb1.x + dz * fb1.y;
      float ke_mc = (ux * ux + uy * uy) + uz * uz;                                // :112-115
      float vz = __fsqrt_rn(one + ke_mc);
      ke_mc *= __fdiv_rn(c, vz + one);
      vz = __fdiv_rn(c, vz);
      float w0 = qdt_4mc2 * vz;                                                   // half Boris rotation, :121-136
      float w1 = (w5 * w5 + w6 * w6) + w7 * w7;
      float w2 = (w0 * w0) * w1;
      float w3 = w0 * (one + ((one_third) * w2) * (one + 0.4f * w2));
      float w4 = __fdiv_rn(w3, one + (w1 * w3) * w3);
      w4 += w4;
      w0 = ux + w3 * (uy * w7 - uz * w6);
      w1 = uy + w3 * (uz * w5 - ux * w7);
      w2 = uz + w3 * (ux * w6 - uy * w5);
      ux += w4 * (w1 * w7 - w2 * w6);
      uy += w4 * (w2 * w5 - w0 * w7);
      uz += w4 * (w0 * w6 - w1 * w5);
      const float vx = ux * vz, vy = uy * vz; vz = uz * vz;
      w0 = r8V * u.w;                                                             // trilinear weights, :152-172
      dx *= w0; w1 = w0 + dx; w0 -= dx;
      w3 = one + dy; w2 = w0 * w3; w3 *= w1;
      dy = one - dy; w0 *= dy; w1 *= dy;
      w7 = one + dz; w4 = w0 * w7; w5 = w1 * w7; w6 = w2 * w7; w7 *= w3;
      dz = one - dz; w0 *= dz; w1 *= dz; w2 *= dz; w3 *= dz;
      float4 *sw = reinterpret_cast<float4 *>(s_w[w][lane]);
      sw[0] = make_float4(w0, w1, w2, w3); sw[1] = make_float4(w4, w5, w6, w7);
      // the 14 moments per unit weight (ACCUM_HYDRO, :178-198), padded to the 16 floats of hydro_t
      const float tx = mspc * ux, ty = mspc * uy, tz = mspc * uz;
      float4 *sm = reinterpret_cast<float4 *>(s_m[w][lane]);
      sm[0] = make_float4(qsp * vx, qsp * vy, qsp * vz, qsp);
      sm[1] = make_float4(tx, ty, tz, mspc * ke_mc);
      sm[2] = make_float4(tx * vx, ty * vy, tz * vz, ty * vz);
      sm[3] = make_float4(tz * vx, tx * vy, 0.0f, 0.0f);
    }
    __syncwarp();
    unsigned rest = __ballot_sync(full, valid);
    while (rest) {                                               // one pass per voxel present in the row
      const int leader = __ffs(rest) - 1;
      const int gv = __shfl_sync(full, vox, leader);
      unsigned grp = __ballot_sync(full, valid && vox == gv);
      rest &= ~grp;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      while (grp) {
        const int l = __ffs(grp) - 1;
        grp &= grp - 1;
        const float wk = s_w[w][l][node];
        const float4 m = *reinterpret_cast<const float4 *>(&s_m[w][l][4 * quad]);
        a0 = __fmaf_rn(wk, m.x, a0); a1 = __fmaf_rn(wk, m.y, a1); a2 = __fmaf_rn(wk, m.z, a2); a3 = __fmaf_rn(wk, m.w, a3);
      }
      red_add_v4(hydro + (size_t)(gv + node_off) * kHydroFloats + 4 * quad, a0, a1, a2, a3);
    }
    __syncwarp();
  }
}

// ADJUST_HYDRO (hydro_array.cc:160-192): every local wall doubles the moments on its plane; a node on two walls is
// doubled twice, as the sequential reference does
__global__ void __launch_bounds__(256) adjust_hydro_kernel(float *hydro, FieldK k) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  if (x > k.nx + 1) return;
  const int n[3] = {k.nx, k.ny, k.nz}; const int cc[3] = {x, y, z};
  float s = 1.0f;
#pragma unroll
  for (int fc = 0; fc < 6; fc++)
    if (k.face[fc] < 0 && cc[fc % 3] == (fc < 3 ? 1 : n[fc % 3] + 1)) s *= 2.0f;
  if (s == 1.0f) return;
  float4 *h = reinterpret_cast<float4 *>(hydro + (size_t)voxel(x, y, z, k.nx, k.ny) * kHydroFloats);
#pragma unroll
  for (int q = 0; q < 4; q++) { float4 v = h[q]; v.x *= s; v.y *= s; v.z *= s; v.w *= s; h[q] = v; }   // powers of two: exact
}

// periodic fold of the two shared planes of one axis (hydro_array.cc:232-268)
__global__ void __launch_bounds__(256) sync_hydro_self_kernel(float *hydro, FieldK k, int X) {
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] + 1 || cz > n[Z] + 1) return;
  const int vl = 1 * s[X] + cy * s[Y] + cz * s[Z], vh = (n[X] + 1) * s[X] + cy * s[Y] + cz * s[Z];
  const float dX = X == 0 ? k.dx : X == 1 ? k.dy : k.dz;
  float rw = dX, lw = rw + dX;
  rw = __fdiv_rn(rw, lw); lw = __fdiv_rn(dX, lw); lw += lw; rw += rw;
  float *a = hydro + (size_t)vl * kHydroFloats, *b = hydro + (size_t)vh * kHydroFloats;
#pragma unroll
  for (int q = 0; q < 14; q++) {
    const float va = a[q], vb = b[q];
    a[q] = lw * va + rw * vb; b[q] = lw * vb + rw * va;
  }
}

// shared node plane of one face, 14 moments per node (the payload of hydro_array.cc:203-226): pack copies it out,
// unpack forms own + remote (equal cell sizes on both sides: lw = rw = 1, hydro_array.cc:236-262)
__global__ void __launch_bounds__(256) hydro_halo_kernel(float *hydro, FieldK k, int face, float *buf, bool unpack) {
  const int X = face % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] + 1 || cz > n[Z] + 1) return;
  const int plane = face < 3 ? 1 : n[X] + 1;
  float *h = hydro + (size_t)(plane * s[X] + cy * s[Y] + cz * s[Z]) * kHydroFloats;
  float *b = buf + 14 * ((size_t)(cz - 1) * (n[Y] + 1) + (cy - 1));
#pragma unroll
  for (int q = 0; q < 14; q++) { if (unpack) h[q] = h[q] + b[q]; else b[q] = h[q]; }
}

}  // namespace vpb

using namespace vpb;

extern "C" size_t vpb_hydro_halo_floats(int32_t nx, int32_t ny, int32_t nz, int axis) {
  const int n[3] = {nx, ny, nz};
  return 14 * (size_t)(n[(axis + 1) % 3] + 1) * (size_t)(n[(axis + 2) % 3] + 1);
}
extern "C" int vpb_hydro_halo_pack(float *hydro, const vpb_field_args_t *geometry, int face, float *buf, void *stream) {
  vpb_field_args_t g = *geometry; g.f = hydro;
  if (int r = check_field_args(&g, "vpb_hydro_halo_pack")) return r;
  VPB_REQUIRE(buf && face >= 0 && face < 6, "vpb_hydro_halo_pack: Bad args");
  hydro_halo_kernel<<<plane_grid(&g, face % 3, 1), 256, 0, as_stream(stream)>>>(hydro, to_k(&g), face, buf, false);
  VPB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vpb_hydro_halo_unpack(float *hydro, const vpb_field_args_t *geometry, int face, const float *buf, void *stream) {
  vpb_field_args_t g = *geometry; g.f = hydro;
  if (int r = check_field_args(&g, "vpb_hydro_halo_unpack")) return r;
  VPB_REQUIRE(buf && face >= 0 && face < 6, "vpb_hydro_halo_unpack: Bad args");
  hydro_halo_kernel<<<plane_grid(&g, face % 3, 1), 256, 0, as_stream(stream)>>>(hydro, to_k(&g), face, const_cast<float *>(buf), true);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_accumulate_hydro_p(float *hydro, const void *p, int32_t np, const float *interp, int32_t interp_stride,
                                      float q, float m, float dt, float cvac, float r8V,
                                      int32_t nx, int32_t ny, int32_t nz, void *stream) {
  VPB_REQUIRE(hydro && interp && (p || np == 0) && nx > 0 && ny > 0 && nz > 0 && interp_stride >= 18 && interp_stride % 4 == 0,
              "vpb_accumulate_hydro_p: Bad args.");
  if (np <= 0) return 0;
  const float qdt_2mc = (q * dt) / (2 * m * cvac);                 // hydro_p_pipeline.cc:238
  const float qdt_4mc2 = qdt_2mc / (2 * cvac);                     // :33-35
  const float mspc = cvac * m;
  const long long rows = ((long long)np + 31) / 32;
  long long grid = (rows + kHydroWarps - 1) / kHydroWarps; if (grid > kSMs * 16) grid = kSMs * 16;
  accumulate_hydro_p_kernel<<<(int)grid, 32 * kHydroWarps, 0, as_stream(stream)>>>(hydro, (const float4 *)p, np, interp, interp_stride,
                                                                     q, mspc, cvac, qdt_2mc, qdt_4mc2, r8V,
                                                                     nx + 2, (nx + 2) * (ny + 2));
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_clear_hydro(float *hydro, int32_t nx, int32_t ny, int32_t nz, void *stream) {
  VPB_REQUIRE(hydro && nx > 0 && ny > 0 && nz > 0, "vpb_clear_hydro: Bad args");
  const size_t nv = (size_t)(nx + 2) * (ny + 2) * (nz + 2);
  VPB_CUDA(cudaMemsetAsync(hydro, 0, nv * kHydroFloats * sizeof(float), as_stream(stream)));
  count_launch();
  return 0;
}

extern "C" int vpb_synchronize_hydro(float *hydro, const vpb_field_args_t *geometry, void *stream) {
  vpb_field_args_t g = *geometry;
  g.f = hydro;                                                     // only the geometry of the argument block is used
  const vpb_field_args_t *a = &g;
  if (int r = check_field_args(a, "vpb_synchronize_hydro")) return r;
  // faces shared with another rank: the caller exchanges their node planes (vpb_hydro_halo_pack / _unpack)
  cudaStream_t st = as_stream(stream);
  bool any_local = false;
  for (int i = 0; i < 6; i++) any_local |= a->face[i] < 0;
  if (any_local) {
    dim3 grid((a->nx + 1 + 255) / 256, a->ny + 1, a->nz + 1);
    adjust_hydro_kernel<<<grid, 256, 0, st>>>(hydro, to_k(a)); VPB_LAUNCH_CHECK();
  }
  for (int X = 0; X < 3; X++) {
    if (a->face[X] == VPB_FACE_PERIODIC_SELF && a->face[X + 3] == VPB_FACE_PERIODIC_SELF) {
      sync_hydro_self_kernel<<<plane_grid(a, X, 1), 256, 0, st>>>(hydro, to_k(a), X); VPB_LAUNCH_CHECK();
    } else if (a->face[X] != VPB_FACE_REMOTE && a->face[X + 3] != VPB_FACE_REMOTE) {
      VPB_REQUIRE(a->face[X] != VPB_FACE_PERIODIC_SELF && a->face[X + 3] != VPB_FACE_PERIODIC_SELF,
                  "vpb_synchronize_hydro: axis %d is periodic on one side only", X);
    }
  }
  return 0;
}
