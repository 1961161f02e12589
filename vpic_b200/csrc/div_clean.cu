// Divergence cleaning (Marder passes) and shared-face synchronisation of the standard field advance on sm_100a:
// the infrequent field operations of vpic_simulation::advance (src/vpic/advance.cc:138-176), kept on the device so
// that a cleaning step does not pull the field array across PCIe.  Compiled with -fmad=false; every update rounds as
// the reference's scalar code does, so the field array stays bit-identical to a host-side cleaning.
//
// Replaces (reference tree, src/field_advance/standard/):
//   sfa.cc:239-256                                         clear_rhof
//   remote.cc:534-620 + local.cc:376-444                   synchronize_rho (local_adjust_rhof / rhob, periodic folds)
//   pipeline/vacuum_compute_div_e_err_pipeline.{h,cc}      vacuum_compute_div_e_err (+ norm-E ghosts remote.cc:136-206,
//                                                          local.cc:128-179; local_adjust_div_e local.cc:298-330)
//   pipeline/compute_rms_div_e_err_pipeline.cc             compute_rms_div_e_err
//   pipeline/vacuum_clean_div_e_pipeline.{h,cc}            vacuum_clean_div_e (+ local_adjust_tang_e)
//   pipeline/compute_div_b_err_pipeline.cc                 compute_div_b_err
//   pipeline/compute_rms_div_b_err_pipeline.cc             compute_rms_div_b_err
//   pipeline/clean_div_b_pipeline.cc                       clean_div_b (+ div-B ghosts remote.cc:208-282,
//                                                          local.cc:181-217; local_adjust_norm_b)
//   remote.cc:298-416                                      synchronize_tang_e_norm_b
//   pipeline/vacuum_compute_rhob_pipeline.{h,cc}           vacuum_compute_rhob (initialisation)
// As in field_advance.cu the reference's interior-pipeline + host-strip split collapses into one kernel per update
// over the node box with per-component range predicates.  Faces shared with another rank (VPB_FACE_REMOTE) are not
// handled here yet; the drop-in layer leaves such field arrays to the reference.
#include "field_common.cuh"

namespace vpb {

__device__ __forceinline__ float *fslot(float4 *f, int v, int slot) { return reinterpret_cast<float *>(f) + 20 * (size_t)v + slot; }
enum { S_EX = 0, S_DIVE = 3, S_CBX = 4, S_DIVB = 7, S_TCAX = 8, S_RHOB = 11, S_JFX = 12, S_RHOF = 15 };

__device__ __forceinline__ float axis_d(const FieldK &k, int X) { return X == 0 ? k.dx : X == 1 ? k.dy : k.dz; }

// interpolation weights of a ghost plane filled from a neighbour whose cell size is `rem` (remote.cc:183-186)
__device__ __forceinline__ void ghost_weights(float rem, float dX, float &rw, float &lw) {
  rw = (float)((2. * (double)dX) / (double)(rem + dX));
  lw = __fdiv_rn(rem - dX, rem + dX);
}

__global__ void __launch_bounds__(256) clear_rhof_kernel(float4 *f, int nv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv) *fslot(f, v, S_RHOF) = 0.0f;
}

// local_adjust_rhof / local_adjust_rhob (local.cc:376-444): one thread per wall node applies the six faces in the
// reference's order, so a node on two walls is doubled twice or zeroed-then-doubled exactly as the host loop does.
__global__ void __launch_bounds__(256) adjust_rho_kernel(FieldK k) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  if (x > k.nx + 1) return;
  const int n[3] = {k.nx, k.ny, k.nz}; const int cc[3] = {x, y, z};
  bool on = false;
#pragma unroll
  for (int fc = 0; fc < 6; fc++) on |= (k.face[fc] < 0) && (cc[fc % 3] == (fc < 3 ? 1 : n[fc % 3] + 1));
  if (!on) return;
  const int v = voxel(x, y, z, k.nx, k.ny);
  float rhof = *fslot(k.f, v, S_RHOF), rhob = *fslot(k.f, v, S_RHOB);
#pragma unroll
  for (int fc = 0; fc < 6; fc++) {
    const int bc = k.face[fc];
    if (bc >= 0 || cc[fc % 3] != (fc < 3 ? 1 : n[fc % 3] + 1)) continue;
    rhof = (bc == -1) ? 0.0f : rhof * 2.0f;
    if (bc == -1) rhob = 0.0f;
  }
  *fslot(k.f, v, S_RHOF) = rhof; *fslot(k.f, v, S_RHOB) = rhob;
}

// synchronize_rho along one axis that is periodic onto this domain (remote.cc:566-585)
__global__ void __launch_bounds__(256) sync_rho_self_kernel(FieldK k, int X) {
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] + 1 || cz > n[Z] + 1) return;
  const int vl = 1 * s[X] + cy * s[Y] + cz * s[Z], vh = (n[X] + 1) * s[X] + cy * s[Y] + cz * s[Z];
  const float dX = axis_d(k, X);
  float hrw = dX, hlw = hrw + dX;
  hrw = __fdiv_rn(hrw, hlw); hlw = __fdiv_rn(dX, hlw);
  const float lw = hlw + hlw, rw = hrw + hrw;
  const float fl = *fslot(k.f, vl, S_RHOF), fh = *fslot(k.f, vh, S_RHOF);
  const float bl = *fslot(k.f, vl, S_RHOB), bh = *fslot(k.f, vh, S_RHOB);
  *fslot(k.f, vl, S_RHOF) = lw * fl + rw * fh;   *fslot(k.f, vh, S_RHOF) = lw * fh + rw * fl;
  *fslot(k.f, vl, S_RHOB) = hlw * bl + hrw * bh; *fslot(k.f, vh, S_RHOB) = hlw * bh + hrw * bl;
}

// Normal-E ghost planes: one thread per (face, Y, Z) node of the ghost plane.
__global__ void __launch_bounds__(256) ghost_norm_e_kernel(FieldK k) {
  const int fc = blockIdx.z;
  const int bc = k.face[fc];
  if (bc == VPB_FACE_REMOTE) return;
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] + 1 || cz > n[Z] + 1) return;
  const bool low = fc < 3;
  const int v = (low ? 0 : n[X] + 1) * s[X] + cy * s[Y] + cz * s[Z];
  const int in1 = v + (low ? s[X] : -s[X]), in2 = in1 + (low ? s[X] : -s[X]);
  float *e = fslot(k.f, v, S_EX + X), *t = fslot(k.f, v, S_TCAX + X);
  if (bc == VPB_FACE_PERIODIC_SELF) {
    const int src = (low ? n[X] : 1) * s[X] + cy * s[Y] + cz * s[Z];
    float rw, lw; ghost_weights(axis_d(k, X), axis_d(k, X), rw, lw);
    *e = rw * *fslot(k.f, src, S_EX + X) + lw * *fslot(k.f, in1, S_EX + X);
  } else if (bc == -1) {
    *e = *fslot(k.f, in1, S_EX + X); *t = *fslot(k.f, in1, S_TCAX + X);
  } else if (bc == -2 || bc == -3) {
    *e = -*fslot(k.f, in1, S_EX + X); *t = -*fslot(k.f, in1, S_TCAX + X);
  } else {                                                            // absorb_fields: linear extrapolation
    *e = 2 * *fslot(k.f, in1, S_EX + X) - *fslot(k.f, in2, S_EX + X);
    *t = 2 * *fslot(k.f, in1, S_TCAX + X) - *fslot(k.f, in2, S_TCAX + X);
  }
}

struct DivECoef { float nc, px, py, pz, cj; };

__global__ void __launch_bounds__(256) div_e_err_kernel(FieldK k, DivECoef c) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  const float4 e0 = FQ(v, 0);
  const float exm = FQ(v - 1, 0).x, eym = FQ(v - sy, 0).y, ezm = FQ(v - sz, 0).z;
  const float rhof = FQ(v, 3).w, rhob = FQ(v, 2).w;
  float err = c.nc * (((c.px * (e0.x - exm) + c.py * (e0.y - eym)) + c.pz * (e0.z - ezm)) - c.cj * (rhof + rhob));
  // local_adjust_div_e: the error is defined to vanish on pec and absorbing walls
  const int n[3] = {nx, ny, nz}; const int cc[3] = {x, y, z};
#pragma unroll
  for (int fc = 0; fc < 6; fc++)
    if ((k.face[fc] == -1 || k.face[fc] == -4) && cc[fc % 3] == (fc < 3 ? 1 : n[fc % 3] + 1)) err = 0.0f;
  *fslot(f, v, S_DIVE) = err;
}

// vacuum_compute_rhob: the bound charge that makes div E consistent with rhof (initialisation), pec walls zeroed
__global__ void __launch_bounds__(256) compute_rhob_kernel(FieldK k, DivECoef c) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  const float4 e0 = FQ(v, 0);
  const float exm = FQ(v - 1, 0).x, eym = FQ(v - sy, 0).y, ezm = FQ(v - sz, 0).z;
  float rhob = c.nc * (((c.px * (e0.x - exm) + c.py * (e0.y - eym)) + c.pz * (e0.z - ezm)) - FQ(v, 3).w);
  const int n[3] = {nx, ny, nz}; const int cc[3] = {x, y, z};
#pragma unroll
  for (int fc = 0; fc < 6; fc++)
    if (k.face[fc] == -1 && cc[fc % 3] == (fc < 3 ? 1 : n[fc % 3] + 1)) rhob = 0.0f;
  *fslot(f, v, S_RHOB) = rhob;
}

__device__ __forceinline__ void block_sum_to(double acc, double *out) {
  __shared__ double s_part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w];
    atomicAdd(out, t);
  }
}

// sum over nodes of w * div_e_err^2, w = 1 inside, 1/2 on walls, 1/4 on edges, 1/8 at corners (rms pipeline :38,96-160)
__global__ void __launch_bounds__(256) rms_div_e_kernel(FieldK k, double *out) {
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  const float4 *f = k.f;
  double acc = 0;
  const long long total = (long long)(nx + 1) * (ny + 1) * (nz + 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % (nx + 1)) + 1, y = (int)((t / (nx + 1)) % (ny + 1)) + 1, z = (int)(t / ((long long)(nx + 1) * (ny + 1))) + 1;
    const float e = FQ(voxel(x, y, z, nx, ny), 0).w;
    const int on = (x == 1 || x == nx + 1) + (y == 1 || y == ny + 1) + (z == 1 || z == nz + 1);
    if (on == 0) acc += (double)(e * e);
    else acc += (on == 1 ? 0.5 : on == 2 ? 0.25 : 0.125) * (double)e * (double)e;
  }
  block_sum_to(acc, out);
}

__global__ void __launch_bounds__(256) rms_div_b_kernel(FieldK k, double *out) {
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  const float4 *f = k.f;
  double acc = 0;
  const long long total = (long long)nx * ny * nz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % nx) + 1, y = (int)((t / nx) % ny) + 1, z = (int)(t / ((long long)nx * ny)) + 1;
    const float e = FQ(voxel(x, y, z, nx, ny), 1).w;
    acc += (double)(e * e);
  }
  block_sum_to(acc, out);
}

// pec walls zero the tangential E and TCA on their plane (local_adjust_tang_e, local.cc:224-265)
__device__ __forceinline__ void pec_tang_e(const FieldK &k, const int cc[3], float4 &e, float4 &t, bool &t_dirty) {
  const int n[3] = {k.nx, k.ny, k.nz};
#pragma unroll
  for (int fc = 0; fc < 6; fc++) {
    if (k.face[fc] != -1) continue;
    const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    if (cc[X] != (fc < 3 ? 1 : n[X] + 1)) continue;
    if (cc[Y] <= n[Y]) { set_comp(e, Y, 0.0f); set_comp(t, Y, 0.0f); t_dirty = true; }
    if (cc[Z] <= n[Z]) { set_comp(e, Z, 0.0f); set_comp(t, Z, 0.0f); t_dirty = true; }
  }
}
// symmetric walls zero the normal B on their plane (local_adjust_norm_b, local.cc:266-297)
__device__ __forceinline__ void sym_norm_b(const FieldK &k, const int cc[3], float4 &b) {
  const int n[3] = {k.nx, k.ny, k.nz};
#pragma unroll
  for (int fc = 0; fc < 6; fc++) {
    if (k.face[fc] != -2) continue;
    const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    if (cc[X] == (fc < 3 ? 1 : n[X] + 1) && cc[Y] <= n[Y] && cc[Z] <= n[Z]) set_comp(b, X, 0.0f);
  }
}

__global__ void __launch_bounds__(256) clean_div_e_kernel(FieldK k, float px, float py, float pz) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  float4 e = FQ(v, 0), t = FQ(v, 2);
  const float d0 = e.w;
  if (x <= nx) e.x += px * (FQ(v + 1, 0).w - d0);
  if (y <= ny) e.y += py * (FQ(v + sy, 0).w - d0);
  if (z <= nz) e.z += pz * (FQ(v + sz, 0).w - d0);
  const int cc[3] = {x, y, z};
  bool t_dirty = false;
  pec_tang_e(k, cc, e, t, t_dirty);
  *fslot(f, v, S_EX) = e.x; *fslot(f, v, S_EX + 1) = e.y; *fslot(f, v, S_EX + 2) = e.z;    // .w is read by neighbours
  if (t_dirty) { *fslot(f, v, S_TCAX) = t.x; *fslot(f, v, S_TCAX + 1) = t.y; *fslot(f, v, S_TCAX + 2) = t.z; }
}

__global__ void __launch_bounds__(256) div_b_err_kernel(FieldK k, float px, float py, float pz) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny;
  if (x > nx) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  const float4 b0 = FQ(v, 1);
  *fslot(f, v, S_DIVB) = (px * (FQ(v + 1, 1).x - b0.x) + py * (FQ(v + sy, 1).y - b0.y)) + pz * (FQ(v + sz, 1).z - b0.z);
}

// div-B-error ghost cells: one thread per (face, Y, Z) cell of the ghost plane
__global__ void __launch_bounds__(256) ghost_div_b_kernel(FieldK k) {
  const int fc = blockIdx.z;
  const int bc = k.face[fc];
  if (bc == VPB_FACE_REMOTE) return;
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] || cz > n[Z]) return;
  const bool low = fc < 3;
  const int v = (low ? 0 : n[X] + 1) * s[X] + cy * s[Y] + cz * s[Z];
  const int in1 = v + (low ? s[X] : -s[X]);
  float *g = fslot(k.f, v, S_DIVB);
  if (bc == VPB_FACE_PERIODIC_SELF) {
    const int src = (low ? n[X] : 1) * s[X] + cy * s[Y] + cz * s[Z];
    float rw, lw; ghost_weights(axis_d(k, X), axis_d(k, X), rw, lw);
    *g = rw * *fslot(k.f, src, S_DIVB) + lw * *fslot(k.f, in1, S_DIVB);
  } else if (bc == -1) *g = *fslot(k.f, in1, S_DIVB);
  else if (bc == -2 || bc == -3) *g = -*fslot(k.f, in1, S_DIVB);
  else *g = 0.0f;
}

__global__ void __launch_bounds__(256) clean_div_b_kernel(FieldK k, float px, float py, float pz) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  float4 b = FQ(v, 1);
  const float d0 = b.w;
  if (y <= ny && z <= nz) b.x += px * (d0 - FQ(v - 1, 1).w);
  if (z <= nz && x <= nx) b.y += py * (d0 - FQ(v - sy, 1).w);
  if (x <= nx && y <= ny) b.z += pz * (d0 - FQ(v - sz, 1).w);
  const int cc[3] = {x, y, z};
  sym_norm_b(k, cc, b);
  *fslot(f, v, S_CBX) = b.x; *fslot(f, v, S_CBX + 1) = b.y; *fslot(f, v, S_CBX + 2) = b.z;
}

// synchronize_tang_e_norm_b, first half: the local adjusts on every wall node
__global__ void __launch_bounds__(256) adjust_tang_e_norm_b_kernel(FieldK k) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  if (x > k.nx + 1) return;
  const int n[3] = {k.nx, k.ny, k.nz}; const int cc[3] = {x, y, z};
  bool on = false;
#pragma unroll
  for (int fc = 0; fc < 6; fc++) on |= (k.face[fc] == -1 || k.face[fc] == -2) && (cc[fc % 3] == (fc < 3 ? 1 : n[fc % 3] + 1));
  if (!on) return;
  float4 *f = k.f;
  const int v = voxel(x, y, z, k.nx, k.ny);
  float4 e = FQ(v, 0), b = FQ(v, 1), t = FQ(v, 2);
  bool t_dirty = false;
  pec_tang_e(k, cc, e, t, t_dirty);
  sym_norm_b(k, cc, b);
  FQ(v, 0) = e; FQ(v, 1) = b; FQ(v, 2) = t;
}

// second half, one axis that is periodic onto this domain: the two shared planes are averaged (remote.cc:340-372);
// err collects (w1-w2)^2 of cbX, eY and eZ once per receiving plane, i.e. twice per pair
__global__ void __launch_bounds__(256) sync_tang_e_norm_b_self_kernel(FieldK k, int X, double *err) {
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  double acc = 0;
  if (cy <= n[Y] + 1 && cz <= n[Z] + 1) {
    const int vl = 1 * s[X] + cy * s[Y] + cz * s[Z], vh = (n[X] + 1) * s[X] + cy * s[Y] + cz * s[Z];
    auto avg = [&](int slot, bool count) {
      float *pl = fslot(k.f, vl, slot), *ph = fslot(k.f, vh, slot);
      const double w1 = *pl, w2 = *ph;
      const float m = (float)(0.5 * (w1 + w2));
      *pl = m; *ph = m;
      if (count) acc += 2.0 * ((w1 - w2) * (w1 - w2));
    };
    if (cy <= n[Y] && cz <= n[Z]) avg(S_CBX + X, true);
    if (cy <= n[Y]) { avg(S_EX + Y, true); avg(S_TCAX + Y, false); }
    if (cz <= n[Z]) { avg(S_EX + Z, true); avg(S_TCAX + Z, false); }
  }
  block_sum_to(acc, err);
}

// Halo planes of a face shared with another rank.  The layouts follow the reference's messages without the one-float
// cell-size header (slabs have equal cells by construction, so lw/rw are the equal-cell constants):
//   RHO            {rhof, rhob} per node of the shared plane, nodes (Y 1..nY+1, Z 1..nZ+1)
//   NORM_E         e_X per node of the first interior plane -> neighbour's ghost plane
//   DIV_B          div_b_err per cell (Y 1..nY, Z 1..nZ) of the first interior plane -> neighbour's ghost cells
//   TANG_E_NORM_B  cb_X per face cell, then {e_Y, tca_Y} per Y edge, then {e_Z, tca_Z} per Z edge of the shared plane
__global__ void __launch_bounds__(256) halo_clean_kernel(FieldK k, int kind, int fc, float *buf, bool pack, double *err) {
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  const bool in_nodes = cy <= n[Y] + 1 && cz <= n[Z] + 1;
  const bool low = fc < 3;
  double acc = 0;
  if (in_nodes) {
    const int at = cy * s[Y] + cz * s[Z];
    const int i_node = (cy - 1) + (n[Y] + 1) * (cz - 1), i_cell = (cy - 1) + n[Y] * (cz - 1);
    const bool in_cell = cy <= n[Y] && cz <= n[Z];
    if (kind == VPB_HALO_RHO) {
      const int v = (low ? 1 : n[X] + 1) * s[X] + at;
      float *rf = fslot(k.f, v, S_RHOF), *rb = fslot(k.f, v, S_RHOB);
      if (pack) { buf[2 * i_node] = *rf; buf[2 * i_node + 1] = *rb; }
      else { *rf = 1.0f * *rf + 1.0f * buf[2 * i_node]; *rb = 0.5f * *rb + 0.5f * buf[2 * i_node + 1]; }
    } else if (kind == VPB_HALO_NORM_E) {
      if (pack) buf[i_node] = *fslot(k.f, (low ? 1 : n[X]) * s[X] + at, S_EX + X);
      else {
        const int v = (low ? 0 : n[X] + 1) * s[X] + at, in1 = v + (low ? s[X] : -s[X]);
        *fslot(k.f, v, S_EX + X) = 1.0f * buf[i_node] + 0.0f * *fslot(k.f, in1, S_EX + X);
      }
    } else if (kind == VPB_HALO_DIV_B) {
      if (in_cell) {
        if (pack) buf[i_cell] = *fslot(k.f, (low ? 1 : n[X]) * s[X] + at, S_DIVB);
        else {
          const int v = (low ? 0 : n[X] + 1) * s[X] + at, in1 = v + (low ? s[X] : -s[X]);
          *fslot(k.f, v, S_DIVB) = 1.0f * buf[i_cell] + 0.0f * *fslot(k.f, in1, S_DIVB);
        }
      }
    } else {                                                         // VPB_HALO_TANG_E_NORM_B
      const int v = (low ? 1 : n[X] + 1) * s[X] + at;
      float *sec_b = buf, *sec_y = buf + n[Y] * n[Z], *sec_z = sec_y + 2 * n[Y] * (n[Z] + 1);
      const int i_ey = (cy - 1) + n[Y] * (cz - 1), i_ez = (cy - 1) + (n[Y] + 1) * (cz - 1);
      auto one = [&](int slot, float *b, bool count) {
        float *own = fslot(k.f, v, slot);
        if (pack) { *b = *own; return; }
        const double w1 = *b, w2 = *own;
        *own = (float)(0.5 * (w1 + w2));
        if (count) acc += (w1 - w2) * (w1 - w2);
      };
      if (in_cell) one(S_CBX + X, sec_b + i_cell, true);
      if (cy <= n[Y]) { one(S_EX + Y, sec_y + 2 * i_ey, true); one(S_TCAX + Y, sec_y + 2 * i_ey + 1, false); }
      if (cz <= n[Z]) { one(S_EX + Z, sec_z + 2 * i_ez, true); one(S_TCAX + Z, sec_z + 2 * i_ez + 1, false); }
    }
  }
  if (err) block_sum_to(acc, err);
}

size_t halo_clean_floats(int nx, int ny, int nz, int axis, int kind) {
  const int n[3] = {nx, ny, nz};
  const size_t nY = n[(axis + 1) % 3], nZ = n[(axis + 2) % 3];
  switch (kind) {
    case VPB_HALO_RHO: return 2 * (nY + 1) * (nZ + 1);
    case VPB_HALO_NORM_E: return (nY + 1) * (nZ + 1);
    case VPB_HALO_DIV_B: return nY * nZ;
    case VPB_HALO_TANG_E_NORM_B: return nY * nZ + 2 * nY * (nZ + 1) + 2 * (nY + 1) * nZ;
  }
  return 0;
}

int halo_clean(const vpb_field_args_t *a, int kind, int face, float *buf, bool pack, double *err_dev, void *stream) {
  halo_clean_kernel<<<plane_grid(a, face % 3, 1), 256, 0, as_stream(stream)>>>(to_k(a), kind, face, buf, pack, err_dev);
  VPB_LAUNCH_CHECK();
  return 0;
}

// faces shared with another rank are exchanged by the caller (vpb_halo_pack / unpack); only a half-periodic axis is an error
static int no_remote_faces(const vpb_field_args_t *a, const char *who) {
  for (int X = 0; X < 3; X++)
    VPB_REQUIRE((a->face[X] == VPB_FACE_PERIODIC_SELF) == (a->face[X + 3] == VPB_FACE_PERIODIC_SELF),
                "%s: axis %d is periodic on one side only", who, X);
  return 0;
}
static dim3 node_grid(const vpb_field_args_t *a) { return dim3((a->nx + 1 + 255) / 256, a->ny + 1, a->nz + 1); }
static void marder_coefficients(const vpb_field_args_t *a, float &px, float &py, float &pz) {
  px = (a->nx > 1) ? a->rdx : 0; py = (a->ny > 1) ? a->rdy : 0; pz = (a->nz > 1) ? a->rdz : 0;
  const float alphadt = (float)(0.3888889 / (double)(px * px + py * py + pz * pz));   // clean_div_b_pipeline.cc:113
  px *= alphadt; py *= alphadt; pz *= alphadt;
}

}  // namespace vpb

using namespace vpb;

#define DIV_ENTRY(who)                                          \
  if (int r = check_field_args(a, who)) return r;               \
  if (int r = no_remote_faces(a, who)) return r;                \
  cudaStream_t st = as_stream(stream); (void)st

extern "C" int vpb_clear_rhof(const vpb_field_args_t *a, void *stream) {
  DIV_ENTRY("vpb_clear_rhof");
  const int nv = (a->nx + 2) * (a->ny + 2) * (a->nz + 2);
  clear_rhof_kernel<<<(nv + 255) / 256, 256, 0, st>>>((float4 *)a->f, nv);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_synchronize_rho(const vpb_field_args_t *a, void *stream) {
  DIV_ENTRY("vpb_synchronize_rho");
  bool any_local = false;
  for (int i = 0; i < 6; i++) any_local |= a->face[i] < 0;
  if (any_local) { adjust_rho_kernel<<<node_grid(a), 256, 0, st>>>(to_k(a)); VPB_LAUNCH_CHECK(); }
  for (int X = 0; X < 3; X++)
    if (a->face[X] == VPB_FACE_PERIODIC_SELF) { sync_rho_self_kernel<<<plane_grid(a, X, 1), 256, 0, st>>>(to_k(a), X); VPB_LAUNCH_CHECK(); }
  return 0;
}

extern "C" int vpb_vacuum_compute_div_e_err(const vpb_field_args_t *a, void *stream) {
  DIV_ENTRY("vpb_vacuum_compute_div_e_err");
  ghost_norm_e_kernel<<<max_plane_grid(a, 6), 256, 0, st>>>(to_k(a));
  VPB_LAUNCH_CHECK();
  const bool hm = a->has_material != 0;
  DivECoef c;                                                     // vacuum_compute_div_e_err_pipeline.h:22-26
  c.nc = hm ? a->material[9] : 1.0f;
  c.px = ((a->nx > 1) ? a->rdx : 0) * (hm ? a->material[10] : 1.0f);
  c.py = ((a->ny > 1) ? a->rdy : 0) * (hm ? a->material[11] : 1.0f);
  c.pz = ((a->nz > 1) ? a->rdz : 0) * (hm ? a->material[12] : 1.0f);
  c.cj = (float)(1. / (double)a->eps0);
  div_e_err_kernel<<<node_grid(a), 256, 0, st>>>(to_k(a), c);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_vacuum_compute_rhob(const vpb_field_args_t *a, void *stream) {
  DIV_ENTRY("vpb_vacuum_compute_rhob");
  ghost_norm_e_kernel<<<max_plane_grid(a, 6), 256, 0, st>>>(to_k(a));
  VPB_LAUNCH_CHECK();
  const bool hm = a->has_material != 0;
  DivECoef c;                                                     // vacuum_compute_rhob_pipeline.h:23-26
  c.nc = hm ? a->material[9] : 1.0f;
  c.px = (a->nx > 1) ? a->eps0 * (hm ? a->material[10] : 1.0f) * a->rdx : 0;
  c.py = (a->ny > 1) ? a->eps0 * (hm ? a->material[11] : 1.0f) * a->rdy : 0;
  c.pz = (a->nz > 1) ? a->eps0 * (hm ? a->material[12] : 1.0f) * a->rdz : 0;
  c.cj = 0;
  compute_rhob_kernel<<<node_grid(a), 256, 0, st>>>(to_k(a), c);
  VPB_LAUNCH_CHECK();
  return 0;
}

static int reduce_grid(long long total) { int g = (int)((total + 255) / 256); return g > kSMs * 4 ? kSMs * 4 : (g < 1 ? 1 : g); }

extern "C" int vpb_compute_rms_div_e_err(const vpb_field_args_t *a, double *sum_dev, void *stream) {
  DIV_ENTRY("vpb_compute_rms_div_e_err");
  VPB_REQUIRE(sum_dev, "vpb_compute_rms_div_e_err: Bad args");
  VPB_CUDA(cudaMemsetAsync(sum_dev, 0, sizeof(double), st));
  rms_div_e_kernel<<<reduce_grid((long long)(a->nx + 1) * (a->ny + 1) * (a->nz + 1)), 256, 0, st>>>(to_k(a), sum_dev);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_vacuum_clean_div_e(const vpb_field_args_t *a, void *stream) {
  DIV_ENTRY("vpb_vacuum_clean_div_e");
  const bool hm = a->has_material != 0;
  const float rdx = (a->nx > 1) ? a->rdx : 0, rdy = (a->ny > 1) ? a->rdy : 0, rdz = (a->nz > 1) ? a->rdz : 0;
  const float alphadt = (float)(0.3888889 / (double)(rdx * rdx + rdy * rdy + rdz * rdz));   // vacuum_clean_div_e_pipeline.h:27-30
  const float px = (alphadt * rdx) * (hm ? a->material[1] : 1.0f);
  const float py = (alphadt * rdy) * (hm ? a->material[3] : 1.0f);
  const float pz = (alphadt * rdz) * (hm ? a->material[5] : 1.0f);
  clean_div_e_kernel<<<node_grid(a), 256, 0, st>>>(to_k(a), px, py, pz);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_compute_div_b_err(const vpb_field_args_t *a, void *stream) {
  DIV_ENTRY("vpb_compute_div_b_err");
  const float px = (a->nx > 1) ? a->rdx : 0, py = (a->ny > 1) ? a->rdy : 0, pz = (a->nz > 1) ? a->rdz : 0;
  dim3 grid((a->nx + 255) / 256, a->ny, a->nz);
  div_b_err_kernel<<<grid, 256, 0, st>>>(to_k(a), px, py, pz);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_compute_rms_div_b_err(const vpb_field_args_t *a, double *sum_dev, void *stream) {
  DIV_ENTRY("vpb_compute_rms_div_b_err");
  VPB_REQUIRE(sum_dev, "vpb_compute_rms_div_b_err: Bad args");
  VPB_CUDA(cudaMemsetAsync(sum_dev, 0, sizeof(double), st));
  rms_div_b_kernel<<<reduce_grid((long long)a->nx * a->ny * a->nz), 256, 0, st>>>(to_k(a), sum_dev);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_clean_div_b(const vpb_field_args_t *a, void *stream) {
  DIV_ENTRY("vpb_clean_div_b");
  ghost_div_b_kernel<<<max_plane_grid(a, 6), 256, 0, st>>>(to_k(a));
  VPB_LAUNCH_CHECK();
  float px, py, pz; marder_coefficients(a, px, py, pz);
  clean_div_b_kernel<<<node_grid(a), 256, 0, st>>>(to_k(a), px, py, pz);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_synchronize_tang_e_norm_b(const vpb_field_args_t *a, double *err_dev, void *stream) {
  DIV_ENTRY("vpb_synchronize_tang_e_norm_b");
  VPB_REQUIRE(err_dev, "vpb_synchronize_tang_e_norm_b: Bad args");
  VPB_CUDA(cudaMemsetAsync(err_dev, 0, sizeof(double), st));
  bool any_local = false;
  for (int i = 0; i < 6; i++) any_local |= (a->face[i] == -1 || a->face[i] == -2);
  if (any_local) { adjust_tang_e_norm_b_kernel<<<node_grid(a), 256, 0, st>>>(to_k(a)); VPB_LAUNCH_CHECK(); }
  for (int X = 0; X < 3; X++)
    if (a->face[X] == VPB_FACE_PERIODIC_SELF) {
      sync_tang_e_norm_b_self_kernel<<<plane_grid(a, X, 1), 256, 0, st>>>(to_k(a), X, err_dev);
      VPB_LAUNCH_CHECK();
    }
  return 0;
}
