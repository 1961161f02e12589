// Device functions shared by the particle kernels (advance_p.cu, boundary_p.cu).  Every translation unit that
// includes this is compiled with -fmad=false: the arithmetic follows the reference's scalar code exactly.
#pragma once
#include "vpb_common.cuh"

namespace vpb {

// Closed-form neighbour lookup (see vpb_neighbor_rule_t).  Coordinates come from the voxel index by float-reciprocal
// division with a remainder step and a +-1 fix-up.
struct NbRule {
  int use, sy, sz, nx, ny, nz;
  float inv_sy, inv_sz;
  long long act[6], delta[6];
};

__device__ __forceinline__ int fast_div(int v, int d, float inv_d) {
  // two float-reciprocal steps (the second divides the small remainder of the first) and a +-1 fix-up: exact for
  // every voxel index the reference allows (nv < 2^31 / 6, src/grid/grid.h:98-103)
  int q = __float2int_rz(__int2float_rn(v) * inv_d);
  int r = v - q * d;
  const int q2 = __float2int_rd(__int2float_rn(r) * inv_d);
  q += q2; r -= q2 * d;
  if (r < 0) q--; else if (r >= d) q++;
  return q;
}

struct PushK {
  float4 *p; int np; int first;
  int *keys;                      // optional: voxel every particle ends the push with (for a following index sort)
  float4 *pout; const int *perm;  // stores go to pout (== p unless an index sort is being applied: loads from p[perm[k]])
  int4 *pm; int max_nm; int *counters;
  const float *interp; int istride;
  float *accum; int astride;
  const long long *neighbor; long long rangel, rangeh;
  float qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;
  int dbg;
  int span;                       // rows of 32 particles a warp of the linear kernel takes at a time
  NbRule nb;
};

// grid_t.neighbor[6*vox + face] as a global voxel id or a negative boundary code
__device__ __forceinline__ long long neighbor_of(const PushK &a, int vox, int face) {
  if (!a.nb.use) return __ldg(a.neighbor + 6ll * vox + face);
  const int axis = face < 3 ? face : face - 3;
  const int d = face < 3 ? -1 : 1;
  int c, n, stride;
  if (axis == 0) { const int q = fast_div(vox, a.nb.sy, a.nb.inv_sy); c = vox - q * a.nb.sy; n = a.nb.nx; stride = 1; }
  else if (axis == 1) {
    const int qz = fast_div(vox, a.nb.sz, a.nb.inv_sz);
    c = fast_div(vox - qz * a.nb.sz, a.nb.sy, a.nb.inv_sy); n = a.nb.ny; stride = a.nb.sy;
  } else { c = fast_div(vox, a.nb.sz, a.nb.inv_sz); n = a.nb.nz; stride = a.nb.sz; }
  const bool inside = d < 0 ? (c > 1) : (c < n);
  if (inside) return a.rangel + vox + d * stride;
  const long long act = a.nb.act[face];
  return act < 0 ? act : a.rangel + vox + a.nb.delta[face];
}


// The 12 accumulator increments of one straight streak inside one voxel (advance_p_pipeline.cc:172-208,
// move_p.cc:277-305).  q = charge*weight, (ux,uy,uz) = half displacement, (dx,dy,dz) = streak midpoint.
__device__ __forceinline__ void streak_currents(float q, float ux, float uy, float uz,
                                                float dx, float dy, float dz, float v5, float (&j)[12]) {
  float v0, v1, v2, v3, v4;
#define VPB_ACC(uX, dY, dZ, o)                                     \
  v4 = q * uX; v1 = v4 * dY; v0 = v4 - v1; v1 += v4;               \
  v4 = 1.0f + dZ; v2 = v0 * v4; v3 = v1 * v4;                      \
  v4 = 1.0f - dZ; v0 *= v4; v1 *= v4;                              \
  v0 += v5; v1 -= v5; v2 -= v5; v3 += v5;                          \
  j[o] = v0; j[o + 1] = v1; j[o + 2] = v2; j[o + 3] = v3;
  VPB_ACC(ux, dy, dz, 0)
  VPB_ACC(uy, dz, dx, 4)
  VPB_ACC(uz, dx, dy, 8)
#undef VPB_ACC
}

__device__ __forceinline__ void deposit_red_v4(float *a, const float (&j)[12]) {
  red_add_v4(a, j[0], j[1], j[2], j[3]);
  red_add_v4(a + 4, j[4], j[5], j[6], j[7]);
  red_add_v4(a + 8, j[8], j[9], j[10], j[11]);
}

// Warp-level segmented reduction by voxel.  Lanes whose voxel is shared by >= kMinGroup lanes are summed with a
// reduce-scatter butterfly (15 shuffles for 12 values, not 60) and one lane per component issues the RED;
// stragglers (drifted particles) go straight to 3 vector REDs.
constexpr int kMinGroup = 6;

__device__ __forceinline__ void deposit_warp_segmented(float *accum, int astride, int vox, bool active,
                                                       const float (&j)[12], int min_group = kMinGroup) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // Voxel-sorted rows hold one or two voxels (three at low ppc): peel off up to three voxels in lane order with a
  // shuffle and a vote each.  (match.any would find every group at once, but it costs ~7 cycles per distinct key —
  // 211 cycles for a fully drifted row, tools/ubench_r2.cu — and drifted rows have nothing worth grouping.)
  unsigned rest = __ballot_sync(full, active);
  bool todo = active;
#pragma unroll 1
  for (int g = 0; g < 3 && rest; g++) {
    const int gv = __shfl_sync(full, vox, __ffs(rest) - 1);
    const bool mine = todo && vox == gv;
    const unsigned grp = __ballot_sync(full, mine);
    rest &= ~grp;
    if (__popc(grp) < min_group) continue;
    float v[16];
#pragma unroll
    for (int c = 0; c < 12; c++) v[c] = mine ? j[c] : 0.0f;
#pragma unroll
    for (int c = 12; c < 16; c++) v[c] = 0.0f;
    // reduce-scatter: after the step with lane-bit b, each lane keeps the half of its values selected by bit b
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
      const bool hi = (lane & bit) != 0;
#pragma unroll
      for (int c = 0; c < half; c++) {
        const float keep = hi ? v[c + half] : v[c];
        const float send = hi ? v[c] : v[c + half];
        v[c] = keep + __shfl_xor_sync(full, send, bit);
      }
    }
    const float tot = v[0] + __shfl_xor_sync(full, v[0], 1);
    const int comp = lane >> 1;
    if (!(lane & 1) && comp < 12) red_add(accum + (size_t)gv * astride + comp, tot);
    if (mine) todo = false;
  }
  if (todo) deposit_red_v4(accum + (size_t)vox * astride, j);          // stragglers: three vector REDs
}

// One streak of move_p, scalar variant of the reference (move_p.cc:233-375), on registers: advance the particle to
// the first cell face on its way (or to the end of its displacement), produce the 12 accumulator increments of that
// streak for voxel dep_vox, and apply the face's boundary action.
// r = {dx,dy,dz,-}, u = {ux,uy,uz,w}, vox = current voxel.  Returns 0 = displacement used up, 1 = the particle left
// the local domain (vox = 8*voxel+face), 2 = more streaks to do.
__device__ __forceinline__ int streak_step(const PushK &a, float q, float4 &r, float4 &u, int &vox,
                                           float &dispx, float &dispy, float &dispz, float (&j)[12]) {
  float s_midx = r.x, s_midy = r.y, s_midz = r.z;
  float s_dispx = dispx, s_dispy = dispy, s_dispz = dispz;
  const float dirx = (s_dispx > 0.0f) ? 1.0f : -1.0f;
  const float diry = (s_dispy > 0.0f) ? 1.0f : -1.0f;
  const float dirz = (s_dispz > 0.0f) ? 1.0f : -1.0f;
  const float v0 = (s_dispx == 0.0f) ? 3.4e38f : __fdiv_rn(dirx - s_midx, s_dispx);
  const float v1 = (s_dispy == 0.0f) ? 3.4e38f : __fdiv_rn(diry - s_midy, s_dispy);
  const float v2 = (s_dispz == 0.0f) ? 3.4e38f : __fdiv_rn(dirz - s_midz, s_dispz);
  float v3 = 2.0f; int axis = 3;
  if (v0 < v3) { v3 = v0; axis = 0; }
  if (v1 < v3) { v3 = v1; axis = 1; }
  if (v2 < v3) { v3 = v2; axis = 2; }
  v3 *= 0.5f;
  s_dispx *= v3; s_dispy *= v3; s_dispz *= v3;
  s_midx += s_dispx; s_midy += s_dispy; s_midz += s_dispz;
  // the reference multiplies by the double constant 1.0/3.0 here (move_p.cc:277)
  const float v5 = (float)((double)(((q * s_dispx) * s_dispy) * s_dispz) * (1.0 / 3.0));
  streak_currents(q, s_dispx, s_dispy, s_dispz, s_midx, s_midy, s_midz, v5, j);
  dispx -= s_dispx; dispy -= s_dispy; dispz -= s_dispz;
  r.x += s_dispx + s_dispx; r.y += s_dispy + s_dispy; r.z += s_dispz + s_dispz;
  if (axis == 3) return 0;
  const float dir = (axis == 0) ? dirx : (axis == 1) ? diry : dirz;
  if (axis == 0) r.x = dir; else if (axis == 1) r.y = dir; else r.z = dir;   // exactly on the face
  const int face = axis + ((dir > 0.0f) ? 3 : 0);
  const long long nb = neighbor_of(a, vox, face);
  if (nb == -1) {                                        // reflect_particles
    if (axis == 0) { u.x = -u.x; dispx = -dispx; }
    else if (axis == 1) { u.y = -u.y; dispy = -dispy; }
    else { u.z = -u.z; dispz = -dispz; }
    return 2;
  }
  if (nb < a.rangel || nb > a.rangeh) { vox = 8 * vox + face; return 1; }
  vox = (int)(nb - a.rangel);
  if (axis == 0) r.x = -dir; else if (axis == 1) r.y = -dir; else r.z = -dir;
  return 2;
}

// move_p for one particle per thread: every streak goes to memory as three vector REDs.
template <bool DBG = false>
__device__ __forceinline__ int move_p_dev(const PushK &a, float4 &r, float4 &u, float &dispx, float &dispy, float &dispz) {
  const float q = a.qsp * u.w;
  int vox = __float_as_int(r.w);
  int st;
  do {
    float j[12];
    const int dep_vox = vox;
    st = streak_step(a, q, r, u, vox, dispx, dispy, dispz, j);
    if (DBG && (a.dbg & 32)) {                           // profiling: all the arithmetic, none of the REDs
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < 12; c++) s += j[c];
      if (s == 1.2345e30f) a.counters[3] = 1;
    } else {
      deposit_red_v4(a.accum + (size_t)dep_vox * a.astride, j);
    }
  } while (st == 2);
  r.w = __int_as_float(vox);
  return st;
}

// move_p for a warp-wide batch.  FIRST_SEG: every mover's first streak lies in the voxel it was queued from, and a
// batch is queued from two or three neighbouring voxels, so the first streaks are summed across the warp by voxel
// (one RED per sum) — the later streaks scatter to the face neighbours and go out as per-lane vector REDs.
// ALL_SEG: every round is summed across the warp (kept for comparison).
template <bool FIRST_SEG, bool ALL_SEG>
__device__ __forceinline__ int move_p_warp(const PushK &a, bool active, float4 &r, float4 &u,
                                           float &dispx, float &dispy, float &dispz) {
  const float q = a.qsp * u.w;
  int vox = __float_as_int(r.w);
  int st = active ? 2 : 0;
  bool first = true;
  while (__any_sync(0xffffffffu, st == 2)) {
    float j[12];
    const int dep_vox = vox;
    const bool dep = (st == 2);
    if (dep) st = streak_step(a, q, r, u, vox, dispx, dispy, dispz, j);
    if (ALL_SEG || (FIRST_SEG && first)) deposit_warp_segmented(a.accum, a.astride, dep_vox, dep, j);
    else if (dep) deposit_red_v4(a.accum + (size_t)dep_vox * a.astride, j);
    first = false;
    if (!ALL_SEG && FIRST_SEG) break;                    // the remaining streaks run per lane below
  }
  if (!ALL_SEG && FIRST_SEG) {
    while (st == 2) {
      float j[12];
      const int dep_vox = vox;
      st = streak_step(a, q, r, u, vox, dispx, dispy, dispz, j);
      deposit_red_v4(a.accum + (size_t)dep_vox * a.astride, j);
    }
  }
  r.w = __int_as_float(vox);
  return st;
}

static inline PushK to_push_k(const vpb_push_args_t *args) {
  PushK k;
  k.p = (float4 *)args->p; k.np = args->np; k.first = args->p_first;
  k.perm = args->perm; k.pout = args->perm ? (float4 *)args->p_out : k.p;
  k.keys = args->keys_out;
  k.pm = (int4 *)args->pm; k.max_nm = args->max_nm; k.counters = args->counters;
  k.interp = args->interp; k.istride = args->interp_stride;
  k.accum = args->accum; k.astride = args->accum_stride;
  k.neighbor = (const long long *)args->neighbor; k.rangel = args->rangel; k.rangeh = args->rangeh;
  k.qdt_2mc = args->qdt_2mc; k.cdt_dx = args->cdt_dx; k.cdt_dy = args->cdt_dy; k.cdt_dz = args->cdt_dz; k.qsp = args->qsp;
  k.dbg = args->debug_skip;
  k.span = 64;
  k.nb.use = 0;
  // grid geometry (always): strides of the voxel index and their float reciprocals for fast_div
  k.nb.nx = args->nx; k.nb.ny = args->ny; k.nb.nz = args->nz;
  k.nb.sy = args->nx + 2; k.nb.sz = (args->nx + 2) * (args->ny + 2);
  k.nb.inv_sy = 1.0f / (float)k.nb.sy; k.nb.inv_sz = 1.0f / (float)k.nb.sz;
  for (int f = 0; f < 6; f++) { k.nb.act[f] = 0; k.nb.delta[f] = 0; }
  const vpb_neighbor_rule_t *nr = args->neighbor_rule;
  if (nr && nr->valid && nr->nx == args->nx && nr->ny == args->ny && nr->nz == args->nz) {
    k.nb.use = 1;
    for (int f = 0; f < 6; f++) { k.nb.act[f] = nr->act[f]; k.nb.delta[f] = nr->delta[f]; }
  }
  return k;
}

}  // namespace vpb
