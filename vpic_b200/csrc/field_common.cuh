// Shared by field_advance.cu and div_clean.cu: the kernel-side view of vpb_field_args_t and the plane/launch helpers.
#pragma once
#include "vpb_common.cuh"

namespace vpb {

struct FieldK {
  float4 *f; int nx, ny, nz;
  float dt, cvac, eps0, damp, dx, dy, dz, rdx, rdy, rdz;
  int face[6];
};

static inline FieldK to_k(const vpb_field_args_t *a) {
  FieldK k; k.f = (float4 *)a->f; k.nx = a->nx; k.ny = a->ny; k.nz = a->nz;
  k.dt = a->dt; k.cvac = a->cvac; k.eps0 = a->eps0; k.damp = a->damp;
  k.dx = a->dx; k.dy = a->dy; k.dz = a->dz; k.rdx = a->rdx; k.rdy = a->rdy; k.rdz = a->rdz;
  for (int i = 0; i < 6; i++) k.face[i] = a->face[i];
  return k;
}

// field_t as five float4: 0 {ex,ey,ez,div_e_err} 1 {cbx,cby,cbz,div_b_err} 2 {tcax,tcay,tcaz,rhob} 3 {jfx,jfy,jfz,rhof} 4 materials
#define FQ(v, q) f[5 * (size_t)(v) + (q)]

__device__ __forceinline__ float comp(const float4 &v, int c) { return c == 0 ? v.x : c == 1 ? v.y : v.z; }
__device__ __forceinline__ void set_comp(float4 &v, int c, float x) { if (c == 0) v.x = x; else if (c == 1) v.y = x; else v.z = x; }

static inline int check_field_args(const vpb_field_args_t *a, const char *who) {
  VPB_REQUIRE(a && a->f && a->nx > 0 && a->ny > 0 && a->nz > 0, "%s: Bad args", who);
  VPB_REQUIRE(a->ny + 1 <= 65535 && a->nz + 1 <= 65535 && a->nx + 1 <= 65535, "%s: grid too large", who);
  for (int i = 0; i < 6; i++)
    VPB_REQUIRE(a->face[i] == VPB_FACE_PERIODIC_SELF || a->face[i] == VPB_FACE_REMOTE || (a->face[i] <= -1 && a->face[i] >= -4),
                "%s: Bad boundary condition encountered (face %d = %d)", who, i, a->face[i]);
  return 0;
}

static inline dim3 plane_grid(const vpb_field_args_t *a, int X, int nz_blocks) {
  const int n[3] = {a->nx, a->ny, a->nz};
  const int Y = (X + 1) % 3, Z = (X + 2) % 3;
  return dim3((n[Y] + 1 + 255) / 256, n[Z] + 1, nz_blocks);
}
static inline dim3 max_plane_grid(const vpb_field_args_t *a, int nz_blocks) {
  dim3 g(1, 1, nz_blocks);
  for (int X = 0; X < 3; X++) { dim3 t = plane_grid(a, X, 1); if (t.x > g.x) g.x = t.x; if (t.y > g.y) g.y = t.y; }
  return g;
}


// div_clean.cu: halo planes of the cleaning / synchronisation kinds (VPB_HALO_RHO .. VPB_HALO_TANG_E_NORM_B)
int halo_clean(const vpb_field_args_t *a, int kind, int face, float *buf, bool pack, double *err_dev, void *stream);
size_t halo_clean_floats(int nx, int ny, int nz, int axis, int kind);

}  // namespace vpb
