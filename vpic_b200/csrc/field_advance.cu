// Standard (Yee FDTD) field advance for one vacuum material on sm_100a, so fields never leave HBM between
// particle pushes.  Compiled with -fmad=false; every update rounds as the reference's scalar pipeline does.
//
// Replaces (reference tree, src/field_advance/standard/):
//   pipeline/advance_b_pipeline.cc:20-125 (+ .h:26-59)          advance_b and its surface strips
//   pipeline/vacuum_advance_e_pipeline.cc:20-332 (+ .h:18-69)   vacuum_advance_e: interior, strips, exterior
//   local.cc:50-130,224-297,335-366                             local_ghost_tang_b, local_adjust_tang_e/norm_b/jf
//   remote.cc:61-134,417-508                                    tang-B ghost planes, synchronize_jf
//   sfa.cc:231-237                                              clear_jf
//   pipeline/vacuum_energy_f_pipeline.cc:12-97 (+ .h:24-75)     vacuum_energy_f
// The reference splits each update into a pipelined interior plus a dozen host-side strips; on the GPU each update
// is ONE kernel over the (nx+1)(ny+1)(nz+1) node box with per-component range predicates — every voxel's update
// is independent once the tangential-B ghost planes are in place.
#include "field_common.cuh"

namespace vpb {

// ---- advance_b ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) advance_b_kernel(FieldK k, float px, float py, float pz) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  const float4 e0 = FQ(v, 0);
  float4 b = FQ(v, 1);
  // neighbours that are never dereferenced outside their component's range stay in-bounds: x+1 <= nx+2-1 etc.
  const float4 ex_ = (x <= nx) ? FQ(v + 1, 0) : e0;
  const float4 ey_ = (y <= ny) ? FQ(v + sy, 0) : e0;
  const float4 ez_ = (z <= nz) ? FQ(v + sz, 0) : e0;
  if (y <= ny && z <= nz) b.x -= (py * (ey_.z - e0.z) - pz * (ez_.y - e0.y));
  if (z <= nz && x <= nx) b.y -= (pz * (ez_.x - e0.x) - px * (ex_.z - e0.z));
  if (x <= nx && y <= ny) b.z -= (px * (ex_.y - e0.y) - py * (ey_.x - e0.x));
  // local_adjust_norm_b: a symmetric_fields (-2) wall zeroes the normal B on its face plane
  if (x == 1 && k.face[0] == -2 && y <= ny && z <= nz) b.x = 0;
  if (x == nx + 1 && k.face[3] == -2 && y <= ny && z <= nz) b.x = 0;
  if (y == 1 && k.face[1] == -2 && z <= nz && x <= nx) b.y = 0;
  if (y == ny + 1 && k.face[4] == -2 && z <= nz && x <= nx) b.y = 0;
  if (z == 1 && k.face[2] == -2 && x <= nx && y <= ny) b.z = 0;
  if (z == nz + 1 && k.face[5] == -2 && x <= nx && y <= ny) b.z = 0;
  FQ(v, 1) = b;
}

// ---- tangential-B ghost planes -----------------------------------------------------------------------------
// One thread per (face, Y, Z) node of the ghost plane.  Periodic-self faces copy the opposite interior plane with the
// reference's interpolation weights (remote.cc:105-117); pec copies the adjacent plane, symmetric/pmc negate it
// (local.cc:74-83).  VPB_FACE_REMOTE planes are written by vpb_halo_unpack instead.
__global__ void __launch_bounds__(256) ghost_tang_b_kernel(FieldK k) {
  const int fc = blockIdx.z;
  const int bc = k.face[fc];
  if (bc == VPB_FACE_REMOTE) return;
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] + 1 || cz > n[Z] + 1) return;
  float4 *f = k.f;
  const bool low = fc < 3;
  const int ghost = low ? 0 : n[X] + 1;
  const int v = ghost * s[X] + cy * s[Y] + cz * s[Z];
  int src; float rw, lw; bool negate = false;
  const int inward = v + (low ? s[X] : -s[X]);
  float *gs = reinterpret_cast<float *>(&FQ(v, 1));
  if (bc == -4) {
    // absorb_fields: 2nd-order accurate 1st-order Higdon ABC on the ghost tangential B (local.cc:84-112); the ghost
    // value is a state variable here (decay * previous ghost + drive * interior - dE/dn + dE/dt terms)
    const float rd[3] = {k.rdx, k.rdy, k.rdz};
    const float cdt_dX = k.cvac * k.dt * rd[X], cdt_dY = k.cvac * k.dt * rd[Y], cdt_dZ = k.cvac * k.dt * rd[Z];
    const float higend = (k.nx > 1 || k.ny > 1 || k.nz > 1) ? (float)1.03527618 : 1.0f;
    float drive = cdt_dX * higend;
    const float decay = (1 - drive) / (1 + drive);
    drive = 2 * drive / (1 + drive);
    const int face = (low ? 1 : n[X] + 1) * s[X] + cy * s[Y] + cz * s[Z];
    const int face_in = face + (low ? s[X] : -s[X]);
    const float4 g = FQ(v, 1), h = FQ(inward, 1);
    const float4 e_f = FQ(face, 0), e_fi = FQ(face_in, 0), e_h = FQ(inward, 0);
    if (cz <= n[Z]) {                                           // cbY over Y in 1..nY+1, Z in 1..nZ
      float t1 = cdt_dX * (comp(e_fi, Z) - comp(e_f, Z));
      t1 = low ? t1 : -t1;
      float t2 = comp(FQ(inward + s[Z], 0), X);
      t2 = cdt_dZ * (t2 - comp(e_h, X));
      gs[Y] = ((decay * comp(g, Y) + drive * comp(h, Y)) - t1) + t2;
    }
    if (cy <= n[Y]) {                                           // cbZ over Y in 1..nY, Z in 1..nZ+1
      float t1 = cdt_dX * (comp(e_fi, Y) - comp(e_f, Y));
      t1 = low ? t1 : -t1;
      float t2 = comp(FQ(inward + s[Y], 0), X);
      t2 = cdt_dY * (t2 - comp(e_h, X));
      gs[Z] = ((decay * comp(g, Z) + drive * comp(h, Z)) + t1) - t2;
    }
    return;
  }
  if (bc == VPB_FACE_PERIODIC_SELF) {
    // the ghost at X=0 receives the plane X=n sent out of the +X port of this same domain, and vice versa
    src = (low ? n[X] : 1) * s[X] + cy * s[Y] + cz * s[Z];
    const float dX = X == 0 ? k.dx : X == 1 ? k.dy : k.dz;
    const float rem = dX;                                       // "remote" cell size
    rw = (float)((2. * (double)dX) / (double)(rem + dX));
    lw = (rem - dX) / (rem + dX);
  } else {
    src = inward; rw = 1.0f; lw = 0.0f; negate = (bc != -1);
  }
  // Scalar stores: the corner line (X ghost, Y = nY+1) also belongs to the Y face's ghost plane, which writes
  // a different component of the same field_t — a float4 read-modify-write would lose one of the two.
  const float4 sv = FQ(src, 1), iv = FQ(inward, 1);
  if (cz <= n[Z]) {                                             // cbY over Y in 1..nY+1, Z in 1..nZ
    float val = comp(sv, Y);
    val = (bc == VPB_FACE_PERIODIC_SELF) ? (rw * val + lw * comp(iv, Y)) : (negate ? -val : val);
    gs[Y] = val;
  }
  if (cy <= n[Y]) {                                             // cbZ over Y in 1..nY, Z in 1..nZ+1
    float val = comp(sv, Z);
    val = (bc == VPB_FACE_PERIODIC_SELF) ? (rw * val + lw * comp(iv, Z)) : (negate ? -val : val);
    gs[Z] = val;
  }
}

// ---- vacuum_advance_e --------------------------------------------------------------------------------------
struct ECoef { float px_muz, px_muy, py_mux, py_muz, pz_muy, pz_mux, cj, damp, decayx, drivex, decayy, drivey, decayz, drivez; };

__global__ void __launch_bounds__(256) vacuum_advance_e_kernel(FieldK k, ECoef c) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  float4 e = FQ(v, 0), t = FQ(v, 2);
  const float4 b0 = FQ(v, 1), j = FQ(v, 3);
  const float4 bx = FQ(v - 1, 1), by = FQ(v - sy, 1), bz = FQ(v - sz, 1);
  // decay/drive of the single material (sfa.cc:119-136); 1 and 1 in true vacuum, kept as explicit multiplies
  if (x <= nx) {
    t.x = (c.py_muz * (b0.z - by.z) - c.pz_muy * (b0.y - bz.y)) - c.damp * t.x;
    e.x = c.decayx * e.x + c.drivex * (t.x - c.cj * j.x);
  }
  if (y <= ny) {
    t.y = (c.pz_mux * (b0.x - bz.x) - c.px_muz * (b0.z - bx.z)) - c.damp * t.y;
    e.y = c.decayy * e.y + c.drivey * (t.y - c.cj * j.y);
  }
  if (z <= nz) {
    t.z = (c.px_muy * (b0.y - bx.y) - c.py_mux * (b0.x - by.x)) - c.damp * t.z;
    e.z = c.decayz * e.z + c.drivez * (t.z - c.cj * j.z);
  }
  // local_adjust_tang_e: a pec (-1) wall zeroes tangential E and TCA on its face plane (local.cc:236-247)
  const int n[3] = {nx, ny, nz}; const int cc[3] = {x, y, z};
#pragma unroll
  for (int fc = 0; fc < 6; fc++) {
    if (k.face[fc] != -1) continue;
    const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    if (cc[X] != (fc < 3 ? 1 : n[X] + 1)) continue;
    if (cc[Y] <= n[Y]) { set_comp(e, Y, 0.0f); set_comp(t, Y, 0.0f); }     // eY over Y 1..nY, Z 1..nZ+1
    if (cc[Z] <= n[Z]) { set_comp(e, Z, 0.0f); set_comp(t, Z, 0.0f); }     // eZ over Y 1..nY+1, Z 1..nZ
  }
  FQ(v, 0) = e; FQ(v, 2) = t;
}

// vacuum_compute_curl_b (vacuum_compute_curl_b_pipeline.{h,cc}): TCA = curl B, no damping history — initialisation
__global__ void __launch_bounds__(256) compute_curl_b_kernel(FieldK k, ECoef c) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2), v = voxel(x, y, z, nx, ny);
  float4 e = FQ(v, 0), t = FQ(v, 2);
  const float4 b0 = FQ(v, 1);
  const float4 bx = FQ(v - 1, 1), by = FQ(v - sy, 1), bz = FQ(v - sz, 1);
  if (x <= nx) t.x = (c.py_muz * (b0.z - by.z) - c.pz_muy * (b0.y - bz.y));
  if (y <= ny) t.y = (c.pz_mux * (b0.x - bz.x) - c.px_muz * (b0.z - bx.z));
  if (z <= nz) t.z = (c.px_muy * (b0.y - bx.y) - c.py_mux * (b0.x - by.x));
  const int n[3] = {nx, ny, nz}; const int cc[3] = {x, y, z};
  bool e_dirty = false;
#pragma unroll
  for (int fc = 0; fc < 6; fc++) {                                 // local_adjust_tang_e
    if (k.face[fc] != -1) continue;
    const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    if (cc[X] != (fc < 3 ? 1 : n[X] + 1)) continue;
    if (cc[Y] <= n[Y]) { set_comp(e, Y, 0.0f); set_comp(t, Y, 0.0f); e_dirty = true; }
    if (cc[Z] <= n[Z]) { set_comp(e, Z, 0.0f); set_comp(t, Z, 0.0f); e_dirty = true; }
  }
  if (e_dirty) FQ(v, 0) = e;
  FQ(v, 2) = t;
}

// ---- jf ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clear_jf_kernel(float4 *f, int nv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  float4 j = FQ(v, 3); j.x = 0; j.y = 0; j.z = 0; FQ(v, 3) = j;
}

// local_adjust_jf (local.cc:335-366): pec zeroes tangential jf on the wall, symmetric/pmc/absorbing double it.
// One thread per node applies the six faces in the reference's order (-x,-y,-z,+x,+y,+z): an edge shared by two
// walls is adjusted twice, exactly as the sequential reference does.
__global__ void __launch_bounds__(256) adjust_jf_kernel(FieldK k) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  if (x > nx + 1) return;
  const int n[3] = {nx, ny, nz}; const int cc[3] = {x, y, z};
  bool on = false;
#pragma unroll
  for (int fc = 0; fc < 6; fc++) on |= (k.face[fc] < 0) && (cc[fc % 3] == (fc < 3 ? 1 : n[fc % 3] + 1));
  if (!on) return;
  float4 *f = k.f;
  const int v = voxel(x, y, z, nx, ny);
  float4 j = FQ(v, 3);
#pragma unroll
  for (int fc = 0; fc < 6; fc++) {
    const int bc = k.face[fc];
    const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    if (bc >= 0 || cc[X] != (fc < 3 ? 1 : n[X] + 1)) continue;
    if (cc[Y] <= n[Y]) set_comp(j, Y, bc == -1 ? 0.0f : comp(j, Y) * 2.0f);     // jfY over Y 1..nY, Z 1..nZ+1
    if (cc[Z] <= n[Z]) set_comp(j, Z, bc == -1 ? 0.0f : comp(j, Z) * 2.0f);     // jfZ over Y 1..nY+1, Z 1..nZ
  }
  FQ(v, 3) = j;
}

// synchronize_jf for one axis whose two faces are periodic onto this domain: both shared planes end up with
// lw*own + rw*other (remote.cc:451-470); both old values are read before either is written.
__global__ void __launch_bounds__(256) sync_jf_self_kernel(FieldK k, int X) {
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] + 1 || cz > n[Z] + 1) return;
  float4 *f = k.f;
  const int vl = 1 * s[X] + cy * s[Y] + cz * s[Z], vh = (n[X] + 1) * s[X] + cy * s[Y] + cz * s[Z];
  const float dX = X == 0 ? k.dx : X == 1 ? k.dy : k.dz;
  float rw = dX, lw = rw + dX;
  rw = __fdiv_rn(rw, lw); lw = __fdiv_rn(dX, lw); lw += lw; rw += rw;
  float4 a = FQ(vl, 3), b = FQ(vh, 3);
  const float4 a0 = a, b0 = b;
  if (cy <= n[Y]) { set_comp(a, Y, lw * comp(a0, Y) + rw * comp(b0, Y)); set_comp(b, Y, lw * comp(b0, Y) + rw * comp(a0, Y)); }
  if (cz <= n[Z]) { set_comp(a, Z, lw * comp(a0, Z) + rw * comp(b0, Z)); set_comp(b, Z, lw * comp(b0, Z) + rw * comp(a0, Z)); }
  FQ(vl, 3) = a; FQ(vh, 3) = b;
}

// ---- halo planes for faces shared with another GPU ---------------------------------------------------------
// Layout of a plane buffer: [ (nY+1)*nZ values of component Y | nY*(nZ+1) values of component Z ] for tang-B,
// and [ nY*(nZ+1) of jfY | (nY+1)*nZ of jfZ ] for jf — the reference's message order without its 1-float header
// (cell sizes are equal across slabs by construction).
__global__ void __launch_bounds__(256) halo_kernel(FieldK k, int kind, int fc, float *buf, bool pack) {
  const int n[3] = {k.nx, k.ny, k.nz};
  const int s[3] = {1, k.nx + 2, (k.nx + 2) * (k.ny + 2)};
  const int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int cy = blockIdx.x * blockDim.x + threadIdx.x + 1, cz = blockIdx.y + 1;
  if (cy > n[Y] + 1 || cz > n[Z] + 1) return;
  float4 *f = k.f;
  const bool low = fc < 3;
  int plane, q;
  if (kind == VPB_HALO_TANG_B) { q = 1; plane = pack ? (low ? 1 : n[X]) : (low ? 0 : n[X] + 1); }
  else                         { q = 3; plane = low ? 1 : n[X] + 1; }
  const int v = plane * s[X] + cy * s[Y] + cz * s[Z];
  float4 val = FQ(v, q);
  // section A holds the component whose range is (Y 1..nY+1, Z 1..nZ); section B the one with (Y 1..nY, Z 1..nZ+1)
  const int cA = (kind == VPB_HALO_TANG_B) ? Y : Z, cB = (kind == VPB_HALO_TANG_B) ? Z : Y;
  const int nA = (n[Y] + 1) * n[Z];
  float *bufA = (kind == VPB_HALO_TANG_B) ? buf : buf + n[Y] * (n[Z] + 1);
  float *bufB = (kind == VPB_HALO_TANG_B) ? buf + nA : buf;
  const bool inA = cz <= n[Z], inB = cy <= n[Y];
  const int iA = (cy - 1) + (n[Y] + 1) * (cz - 1), iB = (cy - 1) + n[Y] * (cz - 1);
  if (pack) {
    if (inA) bufA[iA] = comp(val, cA);
    if (inB) bufB[iB] = comp(val, cB);
  } else if (kind == VPB_HALO_TANG_B) {
    if (inA) set_comp(val, cA, bufA[iA]);
    if (inB) set_comp(val, cB, bufB[iB]);
    FQ(v, q) = val;
  } else {                                                       // jf: own + remote (lw = rw = 1)
    if (inA) set_comp(val, cA, comp(val, cA) + bufA[iA]);
    if (inB) set_comp(val, cB, comp(val, cB) + bufB[iB]);
    FQ(v, q) = val;
  }
}

// ---- vacuum_energy_f ---------------------------------------------------------------------------------------
struct EnCoef { float qepsx, qepsy, qepsz, hrmux, hrmuy, hrmuz; };

__global__ void __launch_bounds__(256) energy_f_kernel(FieldK k, EnCoef m, double *en6) {
  __shared__ double s_part[8][6];
  const int nx = k.nx, ny = k.ny, nz = k.nz;
  const float4 *f = k.f;
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2);
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const long long total = (long long)nx * ny * nz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % nx) + 1, y = (int)((t / nx) % ny) + 1, z = (int)(t / ((long long)nx * ny)) + 1;
    const int v = voxel(x, y, z, nx, ny);
    const float4 e0 = FQ(v, 0), ex = FQ(v + 1, 0), ey = FQ(v + sy, 0), ez = FQ(v + sz, 0);
    const float4 eyz = FQ(v + sy + sz, 0), ezx = FQ(v + sz + 1, 0), exy = FQ(v + 1 + sy, 0);
    const float4 b0 = FQ(v, 1), bx = FQ(v + 1, 1), by = FQ(v + sy, 1), bz = FQ(v + sz, 1);
    acc[0] += (double)(m.qepsx * (((e0.x * e0.x + ey.x * ey.x) + ez.x * ez.x) + eyz.x * eyz.x));
    acc[1] += (double)(m.qepsy * (((e0.y * e0.y + ez.y * ez.y) + ex.y * ex.y) + ezx.y * ezx.y));
    acc[2] += (double)(m.qepsz * (((e0.z * e0.z + ex.z * ex.z) + ey.z * ey.z) + exy.z * exy.z));
    acc[3] += (double)(m.hrmux * (b0.x * b0.x + bx.x * bx.x));
    acc[4] += (double)(m.hrmuy * (b0.y * b0.y + by.y * by.y));
    acc[5] += (double)(m.hrmuz * (b0.z * b0.z + bz.z * bz.z));
  }
#pragma unroll
  for (int c = 0; c < 6; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][c] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w][threadIdx.x];
    atomicAdd(en6 + threadIdx.x, t);
  }
}
__global__ void scale6_kernel(double *en6, double s) { if (threadIdx.x < 6) en6[threadIdx.x] *= s; }

}  // namespace vpb

using namespace vpb;

extern "C" int vpb_advance_b(const vpb_field_args_t *a, float frac, void *stream) {
  if (int r = check_field_args(a, "vpb_advance_b")) return r;
  const float px = (a->nx > 1) ? frac * a->cvac * a->dt * a->rdx : 0;      // advance_b_pipeline.h:26-28
  const float py = (a->ny > 1) ? frac * a->cvac * a->dt * a->rdy : 0;
  const float pz = (a->nz > 1) ? frac * a->cvac * a->dt * a->rdz : 0;
  dim3 grid((a->nx + 1 + 255) / 256, a->ny + 1, a->nz + 1);
  advance_b_kernel<<<grid, 256, 0, as_stream(stream)>>>(to_k(a), px, py, pz);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_vacuum_advance_e(const vpb_field_args_t *a, float frac, void *stream) {
  if (int r = check_field_args(a, "vpb_vacuum_advance_e")) return r;
  VPB_REQUIRE(frac == 1, "standard advance_e does not support frac != 1 yet");   // vacuum_advance_e_pipeline.cc:58-61
  cudaStream_t st = as_stream(stream);
  ghost_tang_b_kernel<<<max_plane_grid(a, 6), 256, 0, st>>>(to_k(a));
  VPB_LAUNCH_CHECK();
  ECoef c;
  const float damp = a->damp;
  const bool hm = a->has_material != 0;
  const float *m = a->material;
  const float rmux = hm ? m[6] : 1.0f, rmuy = hm ? m[7] : 1.0f, rmuz = hm ? m[8] : 1.0f;
  c.px_muz = ((a->nx > 1) ? (1 + damp) * a->cvac * a->dt * a->rdx : 0) * rmuz;   // vacuum_advance_e_pipeline.h:18-33
  c.px_muy = ((a->nx > 1) ? (1 + damp) * a->cvac * a->dt * a->rdx : 0) * rmuy;
  c.py_mux = ((a->ny > 1) ? (1 + damp) * a->cvac * a->dt * a->rdy : 0) * rmux;
  c.py_muz = ((a->ny > 1) ? (1 + damp) * a->cvac * a->dt * a->rdy : 0) * rmuz;
  c.pz_muy = ((a->nz > 1) ? (1 + damp) * a->cvac * a->dt * a->rdz : 0) * rmuy;
  c.pz_mux = ((a->nz > 1) ? (1 + damp) * a->cvac * a->dt * a->rdz : 0) * rmux;
  c.cj = a->dt / a->eps0;
  c.damp = damp;
  c.decayx = hm ? m[0] : 1.0f; c.drivex = hm ? m[1] : 1.0f;
  c.decayy = hm ? m[2] : 1.0f; c.drivey = hm ? m[3] : 1.0f;
  c.decayz = hm ? m[4] : 1.0f; c.drivez = hm ? m[5] : 1.0f;
  dim3 grid((a->nx + 1 + 255) / 256, a->ny + 1, a->nz + 1);
  vacuum_advance_e_kernel<<<grid, 256, 0, st>>>(to_k(a), c);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_vacuum_compute_curl_b(const vpb_field_args_t *a, void *stream) {
  if (int r = check_field_args(a, "vpb_vacuum_compute_curl_b")) return r;
  cudaStream_t st = as_stream(stream);
  ghost_tang_b_kernel<<<max_plane_grid(a, 6), 256, 0, st>>>(to_k(a));
  VPB_LAUNCH_CHECK();
  const bool hm = a->has_material != 0;
  const float *m = a->material;
  const float rmux = hm ? m[6] : 1.0f, rmuy = hm ? m[7] : 1.0f, rmuz = hm ? m[8] : 1.0f;
  ECoef c;
  memset(&c, 0, sizeof c);
  c.px_muz = ((a->nx > 1) ? a->cvac * a->dt * a->rdx : 0) * rmuz;   // vacuum_compute_curl_b_pipeline.h:23-28
  c.px_muy = ((a->nx > 1) ? a->cvac * a->dt * a->rdx : 0) * rmuy;
  c.py_mux = ((a->ny > 1) ? a->cvac * a->dt * a->rdy : 0) * rmux;
  c.py_muz = ((a->ny > 1) ? a->cvac * a->dt * a->rdy : 0) * rmuz;
  c.pz_muy = ((a->nz > 1) ? a->cvac * a->dt * a->rdz : 0) * rmuy;
  c.pz_mux = ((a->nz > 1) ? a->cvac * a->dt * a->rdz : 0) * rmux;
  dim3 grid((a->nx + 1 + 255) / 256, a->ny + 1, a->nz + 1);
  compute_curl_b_kernel<<<grid, 256, 0, st>>>(to_k(a), c);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_clear_jf(const vpb_field_args_t *a, void *stream) {
  if (int r = check_field_args(a, "vpb_clear_jf")) return r;
  const int nv = (a->nx + 2) * (a->ny + 2) * (a->nz + 2);
  clear_jf_kernel<<<(nv + 255) / 256, 256, 0, as_stream(stream)>>>((float4 *)a->f, nv);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_synchronize_jf(const vpb_field_args_t *a, void *stream) {
  if (int r = check_field_args(a, "vpb_synchronize_jf")) return r;
  cudaStream_t st = as_stream(stream);
  bool any_local = false;
  for (int i = 0; i < 6; i++) any_local |= a->face[i] < 0;
  if (any_local) {
    dim3 grid((a->nx + 1 + 255) / 256, a->ny + 1, a->nz + 1);
    adjust_jf_kernel<<<grid, 256, 0, st>>>(to_k(a)); VPB_LAUNCH_CHECK();
  }
  for (int X = 0; X < 3; X++) {
    if (a->face[X] == VPB_FACE_PERIODIC_SELF && a->face[X + 3] == VPB_FACE_PERIODIC_SELF) {
      sync_jf_self_kernel<<<plane_grid(a, X, 1), 256, 0, st>>>(to_k(a), X);
      VPB_LAUNCH_CHECK();
    } else {
      VPB_REQUIRE(a->face[X] != VPB_FACE_PERIODIC_SELF && a->face[X + 3] != VPB_FACE_PERIODIC_SELF,
                  "vpb_synchronize_jf: axis %d is periodic on one side only", X);
      // VPB_FACE_REMOTE axes are exchanged by the caller with vpb_halo_pack / NCCL / vpb_halo_unpack, in axis order
    }
  }
  return 0;
}

extern "C" int vpb_vacuum_energy_f(const vpb_field_args_t *a, double *en6_dev, void *stream) {
  if (int r = check_field_args(a, "vpb_vacuum_energy_f")) return r;
  VPB_REQUIRE(en6_dev, "vpb_vacuum_energy_f: Bad args");
  cudaStream_t st = as_stream(stream);
  VPB_CUDA(cudaMemsetAsync(en6_dev, 0, 6 * sizeof(double), st));
  long long total = (long long)a->nx * a->ny * a->nz;
  int grid = (int)((total + 255) / 256); if (grid > kSMs * 4) grid = kSMs * 4;
  const bool hm = a->has_material != 0;                       // vacuum_energy_f_pipeline.h:24-29
  EnCoef m;
  m.qepsx = 0.25f * (hm ? a->material[10] : 1.0f); m.qepsy = 0.25f * (hm ? a->material[11] : 1.0f);
  m.qepsz = 0.25f * (hm ? a->material[12] : 1.0f);
  m.hrmux = 0.5f * (hm ? a->material[6] : 1.0f); m.hrmuy = 0.5f * (hm ? a->material[7] : 1.0f);
  m.hrmuz = 0.5f * (hm ? a->material[8] : 1.0f);
  energy_f_kernel<<<grid, 256, 0, st>>>(to_k(a), m, en6_dev);
  VPB_LAUNCH_CHECK();
  scale6_kernel<<<1, 32, 0, st>>>(en6_dev, 0.5 * (double)a->eps0 * (double)a->dV);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t vpb_halo_floats(int32_t nx, int32_t ny, int32_t nz, int axis) {
  const int n[3] = {nx, ny, nz};
  const int Y = (axis + 1) % 3, Z = (axis + 2) % 3;
  return (size_t)(n[Y] + 1) * n[Z] + (size_t)n[Y] * (n[Z] + 1);
}

extern "C" size_t vpb_halo_floats_kind(int32_t nx, int32_t ny, int32_t nz, int axis, int kind) {
  if (kind == VPB_HALO_TANG_B || kind == VPB_HALO_JF) return vpb_halo_floats(nx, ny, nz, axis);
  return halo_clean_floats(nx, ny, nz, axis, kind);
}

extern "C" int vpb_halo_unpack_sync(const vpb_field_args_t *a, int face, const float *buf, double *err_dev, void *stream) {
  if (int r = check_field_args(a, "vpb_halo_unpack_sync")) return r;
  VPB_REQUIRE(buf && err_dev && face >= 0 && face < 6, "vpb_halo_unpack_sync: Bad args");
  return halo_clean(a, VPB_HALO_TANG_E_NORM_B, face, (float *)buf, false, err_dev, stream);
}

static int halo_common(const vpb_field_args_t *a, int kind, int face, float *buf, bool pack, void *stream) {
  if (int r = check_field_args(a, "vpb_halo")) return r;
  if (kind >= VPB_HALO_RHO && kind <= VPB_HALO_TANG_E_NORM_B) {
    VPB_REQUIRE(buf && face >= 0 && face < 6, "vpb_halo: Bad args");
    VPB_REQUIRE(pack || kind != VPB_HALO_TANG_E_NORM_B, "vpb_halo_unpack: use vpb_halo_unpack_sync for VPB_HALO_TANG_E_NORM_B");
    return halo_clean(a, kind, face, buf, pack, nullptr, stream);
  }
  VPB_REQUIRE(buf && face >= 0 && face < 6 && (kind == VPB_HALO_TANG_B || kind == VPB_HALO_JF), "vpb_halo: Bad args");
  halo_kernel<<<plane_grid(a, face % 3, 1), 256, 0, as_stream(stream)>>>(to_k(a), kind, face, buf, pack);
  VPB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vpb_halo_pack(const vpb_field_args_t *a, int kind, int face, float *buf, void *stream) {
  return halo_common(a, kind, face, buf, true, stream);
}
extern "C" int vpb_halo_unpack(const vpb_field_args_t *a, int kind, int face, const float *buf, void *stream) {
  return halo_common(a, kind, face, (float *)buf, false, stream);
}
