// advance_p for sm_100a: relativistic Boris push + charge-conserving current deposit + move_p.
//
// Replaces (reference tree):
//   src/species_advance/standard/pipeline/advance_p_pipeline.cc:20-340   push, in-bounds deposit, mover hand-off
//   src/species_advance/standard/move_p.cc:216-378                       cell-crossing streaks and boundary codes
//
// This file is compiled with -fmad=false: the push follows the reference's SCALAR pipeline operation for
// operation (same association, IEEE sqrt/divide, float constants rounded as the reference rounds them), so
// particle state (offsets, voxel index, momentum) is bit-identical to the scalar CPU build.  Deposited currents
// are the same per-particle values; only the order of the fp32 sums differs (atomics).
//
// Mapping to the machine (HBM-bound, no tensor cores — see DESIGN.md):
//   * one thread per particle, a CTA walks tiles of kTile consecutive particles (voxel-sorted by sort_p), so a
//     warp's 32 particles sit in one or two voxels: its five 128-bit interpolator loads are L1 broadcasts;
//   * particles move as two 128-bit halves ({dx,dy,dz,i} and {ux,uy,uz,w}); a warp's two loads cover the same
//     eight 128-byte lines, so HBM sees each line once;
//   * particles that leave their voxel are NOT moved inline (one crossing lane would stall its 31 neighbours
//     through the whole streak loop): they are queued in shared memory and finished by the CTA as a dense batch
//     after the tile, which keeps both phases convergent;
//   * deposit strategies (args.variant) — see deposit.cuh.
#include "vpb_common.cuh"

namespace vpb {

constexpr int kBlock = 256;
constexpr int kPPT   = 4;                 // particles per thread per tile
constexpr int kTile  = kBlock * kPPT;

struct PushK {
  float4 *p; int np;
  int4 *pm; int max_nm; int *counters;
  const float *interp; int istride;
  float *accum; int astride;
  const long long *neighbor; long long rangel, rangeh;
  float qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;
};

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
                                                       const float (&j)[12]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int key = active ? vox : (-1 - lane);
  const unsigned peers = __match_any_sync(full, key);
  const bool grouped = active && (__popc(peers) >= kMinGroup);
  if (active && !grouped) deposit_red_v4(accum + (size_t)vox * astride, j);
  unsigned big = __ballot_sync(full, grouped);
  while (big) {
    const int leader = __ffs(big) - 1;
    const unsigned grp = __shfl_sync(full, peers, leader);
    const int gv = __shfl_sync(full, vox, leader);
    const bool mine = grouped && (peers == grp);
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
    float tot = v[0] + __shfl_xor_sync(full, v[0], 1);
    const int comp = lane >> 1;
    if (!(lane & 1) && comp < 12) red_add(accum + (size_t)gv * astride + comp, tot);
    big &= ~grp;
  }
}

template <int VARIANT>
__device__ __forceinline__ void deposit(const PushK &a, int vox, bool active, const float (&j)[12]) {
  if (VARIANT == VPB_DEPOSIT_WARP_SEG) {
    deposit_warp_segmented(a.accum, a.astride, vox, active, j);
  } else {
    if (active) deposit_red_v4(a.accum + (size_t)vox * a.astride, j);
  }
}

// move_p, scalar variant of the reference (move_p.cc:216-378), on registers.
// r = {dx,dy,dz,i}, u = {ux,uy,uz,w}; returns 1 when the particle left the local domain (r.w = 8*voxel+face).
__device__ __forceinline__ int move_p_dev(const PushK &a, float4 &r, float4 &u, float &dispx, float &dispy, float &dispz) {
  const float q = a.qsp * u.w;
  int vox = __float_as_int(r.w);
  int ret = 0;
  for (;;) {
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
    float j[12];
    streak_currents(q, s_dispx, s_dispy, s_dispz, s_midx, s_midy, s_midz, v5, j);
    deposit_red_v4(a.accum + (size_t)vox * a.astride, j);
    dispx -= s_dispx; dispy -= s_dispy; dispz -= s_dispz;
    r.x += s_dispx + s_dispx; r.y += s_dispy + s_dispy; r.z += s_dispz + s_dispz;
    if (axis == 3) break;
    const float dir = (axis == 0) ? dirx : (axis == 1) ? diry : dirz;
    if (axis == 0) r.x = dir; else if (axis == 1) r.y = dir; else r.z = dir;   // exactly on the face
    const int face = axis + ((dir > 0.0f) ? 3 : 0);
    const long long nb = __ldg(a.neighbor + 6ll * vox + face);
    if (nb == -1) {                                        // reflect_particles
      if (axis == 0) { u.x = -u.x; dispx = -dispx; }
      else if (axis == 1) { u.y = -u.y; dispy = -dispy; }
      else { u.z = -u.z; dispz = -dispz; }
      continue;
    }
    if (nb < a.rangel || nb > a.rangeh) { vox = 8 * vox + face; ret = 1; break; }
    vox = (int)(nb - a.rangel);
    if (axis == 0) r.x = -dir; else if (axis == 1) r.y = -dir; else r.z = -dir;
  }
  r.w = __int_as_float(vox);
  return ret;
}

template <int VARIANT>
__global__ void __launch_bounds__(kBlock) advance_p_kernel(const PushK a) {
  __shared__ int4 s_mv[kTile];          // queued movers of this tile: {disp bits x3, particle index}
  __shared__ int s_nmv;

  const float one = 1.0f;
  const float one_third = (float)(1.0 / 3.0);
  const float two_fifteenths = (float)(2.0 / 15.0);
  const int tid = threadIdx.x;
  const int ntiles = (a.np + kTile - 1) / kTile;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid == 0) s_nmv = 0;
    __syncthreads();

#pragma unroll
    for (int k = 0; k < kPPT; k++) {
      const int i = tile * kTile + k * kBlock + tid;
      const bool valid = i < a.np;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f), u = r;
      if (valid) { r = a.p[2 * (size_t)i]; u = a.p[2 * (size_t)i + 1]; }
      const int ii = __float_as_int(r.w);
      bool inb = false;
      float j[12];
      if (valid) {
        const float4 *f = reinterpret_cast<const float4 *>(a.interp + (size_t)ii * a.istride);
        const float4 fex = __ldg(f), fey = __ldg(f + 1), fez = __ldg(f + 2), fb0 = __ldg(f + 3);
        const float2 fb1 = __ldg(reinterpret_cast<const float2 *>(f + 4));
        const float dx = r.x, dy = r.y, dz = r.z;
        const float hax = a.qdt_2mc * ((fex.x + dy * fex.y) + dz * (fex.z + dy * fex.w));
        const float hay = a.qdt_2mc * ((fey.x + dz * fey.y) + dx * (fey.z + dz * fey.w));
        const float haz = a.qdt_2mc * ((fez.x + dx * fez.y) + dy * (fez.z + dx * fez.w));
        const float cbx = fb0.x + dx * fb0.y;
        const float cby = fb0.z + dy * fb0.w;
        const float cbz = fb1.x + dz * fb1.y;
        float ux = u.x, uy = u.y, uz = u.z;
        ux += hax; uy += hay; uz += haz;
        float v0 = __fdiv_rn(a.qdt_2mc, __fsqrt_rn(one + (ux * ux + (uy * uy + uz * uz))));
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
        ux += hax; uy += hay; uz += haz;
        u.x = ux; u.y = uy; u.z = uz;
        a.p[2 * (size_t)i + 1] = u;                                   // momentum is stored in either case
        v0 = __fdiv_rn(one, __fsqrt_rn(one + (ux * ux + (uy * uy + uz * uz))));
        ux *= a.cdt_dx; uy *= a.cdt_dy; uz *= a.cdt_dz;
        ux *= v0; uy *= v0; uz *= v0;                                 // half displacement in cell units
        v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;                     // streak midpoint
        v3 = v0 + ux; v4 = v1 + uy; const float v5n = v2 + uz;        // new position
        inb = (v3 <= one) && (v4 <= one) && (v5n <= one) && (-v3 <= one) && (-v4 <= one) && (-v5n <= one);
        if (inb) {
          const float q = u.w * a.qsp;
          a.p[2 * (size_t)i] = make_float4(v3, v4, v5n, r.w);
          const float v5 = (((q * ux) * uy) * uz) * one_third;
          streak_currents(q, ux, uy, uz, v0, v1, v2, v5, j);
        } else {
          const int slot = atomicAdd(&s_nmv, 1);
          s_mv[slot] = make_int4(__float_as_int(ux), __float_as_int(uy), __float_as_int(uz), i);
        }
      }
      deposit<VARIANT>(a, ii, inb, j);
    }
    __syncthreads();

    // finish the queued movers as a dense batch
    const int nmv = s_nmv;
    for (int m = tid; m < nmv; m += kBlock) {
      const int4 mv = s_mv[m];
      const int i = mv.w;
      float4 r = a.p[2 * (size_t)i], u = a.p[2 * (size_t)i + 1];
      float dispx = __int_as_float(mv.x), dispy = __int_as_float(mv.y), dispz = __int_as_float(mv.z);
      const float ux0 = u.x, uy0 = u.y, uz0 = u.z;
      const int left = move_p_dev(a, r, u, dispx, dispy, dispz);
      if (left) {
        const int slot = atomicAdd(a.counters, 1);
        if (slot < a.max_nm) {
          a.pm[slot] = make_int4(__float_as_int(dispx), __float_as_int(dispy), __float_as_int(dispz), i);
        } else {
          atomicAdd(a.counters + 1, 1);                               // lost mover: keep p.i a valid voxel
          r.w = __int_as_float(__float_as_int(r.w) >> 3);
        }
      }
      a.p[2 * (size_t)i] = r;
      if (u.x != ux0 || u.y != uy0 || u.z != uz0) a.p[2 * (size_t)i + 1] = u;   // reflected
    }
    __syncthreads();
  }
}

}  // namespace vpb

using namespace vpb;

extern "C" int vpb_advance_p(const vpb_push_args_t *args, void *stream) {
  VPB_REQUIRE(args && args->p && args->interp && args->accum && args->neighbor && args->counters,
              "vpb_advance_p: Bad args.");
  VPB_REQUIRE(args->pm || args->max_nm == 0, "vpb_advance_p: mover array missing");
  VPB_REQUIRE(args->interp_stride % 4 == 0 && args->accum_stride % 4 == 0,
              "vpb_advance_p: strides must keep 16-byte alignment (got %d, %d)", args->interp_stride, args->accum_stride);
  VPB_REQUIRE((((uintptr_t)args->p | (uintptr_t)args->interp | (uintptr_t)args->accum | (uintptr_t)args->pm) & 15) == 0,
              "vpb_advance_p: arrays must be 16-byte aligned");
  if (args->np <= 0) return 0;
  PushK k;
  k.p = (float4 *)args->p; k.np = args->np;
  k.pm = (int4 *)args->pm; k.max_nm = args->max_nm; k.counters = args->counters;
  k.interp = args->interp; k.istride = args->interp_stride;
  k.accum = args->accum; k.astride = args->accum_stride;
  k.neighbor = (const long long *)args->neighbor; k.rangel = args->rangel; k.rangeh = args->rangeh;
  k.qdt_2mc = args->qdt_2mc; k.cdt_dx = args->cdt_dx; k.cdt_dy = args->cdt_dy; k.cdt_dz = args->cdt_dz; k.qsp = args->qsp;
  const int ntiles = (args->np + kTile - 1) / kTile;
  const int grid = ntiles < kSMs * 8 ? ntiles : kSMs * 8;
  int variant = args->variant == VPB_DEPOSIT_DEFAULT ? VPB_DEPOSIT_WARP_SEG : args->variant;
  switch (variant) {
    case VPB_DEPOSIT_RED_V4:
      advance_p_kernel<VPB_DEPOSIT_RED_V4><<<grid, kBlock, 0, as_stream(stream)>>>(k); break;
    case VPB_DEPOSIT_WARP_SEG:
      advance_p_kernel<VPB_DEPOSIT_WARP_SEG><<<grid, kBlock, 0, as_stream(stream)>>>(k); break;
    default:
      VPB_REQUIRE(false, "vpb_advance_p: unknown deposit variant %d", variant);
  }
  VPB_LAUNCH_CHECK();
  return 0;
}
