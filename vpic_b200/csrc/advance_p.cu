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
#include "push_common.cuh"

namespace vpb {

constexpr int kBlock = 256;
constexpr int kPPT   = 4;                 // particles per thread per tile
constexpr int kTile  = kBlock * kPPT;
constexpr int kMinBlocks = 3;            // resident CTAs per SM the register budget is tuned for

template <int VARIANT>
__device__ __forceinline__ void deposit(const PushK &a, int vox, bool active, const float (&j)[12]) {
  if (VARIANT == VPB_DEPOSIT_WARP_SEG) {
    deposit_warp_segmented(a.accum, a.astride, vox, active, j);
  } else {
    if (active) deposit_red_v4(a.accum + (size_t)vox * a.astride, j);
  }
}

template <int VARIANT>
__global__ void __launch_bounds__(kBlock, kMinBlocks) advance_p_kernel(const PushK a) {
  // Every warp is autonomous: it owns warp-tiles of 32*kPPT consecutive particles, keeps its own mover queue in
  // shared memory and never meets a block-wide barrier, so a warp waiting on HBM does not hold up its neighbours.
  __shared__ int4 s_mv[kBlock / 32][32 * kPPT];   // queued movers of the warp-tile: {disp bits x3, particle index}

  const float one = 1.0f;
  const float one_third = (float)(1.0 / 3.0);
  const float two_fifteenths = (float)(2.0 / 15.0);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wtile = 32 * kPPT;
  const int n_wtiles = (a.np + wtile - 1) / wtile;
  const int warps_total = gridDim.x * (kBlock / 32);
  int4 *q = s_mv[w];

  for (int wt = blockIdx.x * (kBlock / 32) + w; wt < n_wtiles; wt += warps_total) {
    const int base = wt * wtile;
    // issue every particle load of the warp-tile before touching any of them (kPPT 1-KB requests in flight)
    float4 r[kPPT], u[kPPT];
#pragma unroll
    for (int k = 0; k < kPPT; k++) {
      const int i = base + k * 32 + lane;
      r[k] = make_float4(0.f, 0.f, 0.f, 0.f); u[k] = r[k];
      if (i < a.np) ld_particle(a.p + 2 * (size_t)i, r[k], u[k]);
    }
    int nq = 0;                                                         // warp-uniform queue length

#pragma unroll
    for (int k = 0; k < kPPT; k++) {
      const int i = base + k * 32 + lane;
      const bool valid = i < a.np;
      const int ii = __float_as_int(r[k].w);
      bool inb = false;
      float j[12];
      float mux = 0.f, muy = 0.f, muz = 0.f;
      if (valid) {
        const float4 *f = reinterpret_cast<const float4 *>(a.interp + (size_t)ii * a.istride);
        float4 fex = make_float4(.01f, .02f, .03f, .04f), fey = fex, fez = fex, fb0 = fex;
        float2 fb1 = make_float2(.01f, .02f);
        if (!(a.dbg & 8)) {
          fex = __ldg(f); fey = __ldg(f + 1); fez = __ldg(f + 2); fb0 = __ldg(f + 3);
          fb1 = __ldg(reinterpret_cast<const float2 *>(f + 4));
        }
        const float dx = r[k].x, dy = r[k].y, dz = r[k].z;
        const float hax = a.qdt_2mc * ((fex.x + dy * fex.y) + dz * (fex.z + dy * fex.w));
        const float hay = a.qdt_2mc * ((fey.x + dz * fey.y) + dx * (fey.z + dz * fey.w));
        const float haz = a.qdt_2mc * ((fez.x + dx * fez.y) + dy * (fez.z + dx * fez.w));
        const float cbx = fb0.x + dx * fb0.y;
        const float cby = fb0.z + dy * fb0.w;
        const float cbz = fb1.x + dz * fb1.y;
        float ux = u[k].x, uy = u[k].y, uz = u[k].z;
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
        float4 un = u[k];
        un.x = ux; un.y = uy; un.z = uz;                               // momentum is stored in either case
        v0 = __fdiv_rn(one, __fsqrt_rn(one + (ux * ux + (uy * uy + uz * uz))));
        ux *= a.cdt_dx; uy *= a.cdt_dy; uz *= a.cdt_dz;
        ux *= v0; uy *= v0; uz *= v0;                                 // half displacement in cell units
        v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;                     // streak midpoint
        v3 = v0 + ux; v4 = v1 + uy; const float v5n = v2 + uz;        // new position
        inb = (v3 <= one) && (v4 <= one) && (v5n <= one) && (-v3 <= one) && (-v4 <= one) && (-v5n <= one);
        if (a.dbg & 2) inb = true;
        if (inb) {
          const float qw = un.w * a.qsp;
          if (!(a.dbg & 4)) st_particle(a.p + 2 * (size_t)i, make_float4(v3, v4, v5n, r[k].w), un);
          const float v5 = (((qw * ux) * uy) * uz) * one_third;
          streak_currents(qw, ux, uy, uz, v0, v1, v2, v5, j);
        } else {
          a.p[2 * (size_t)i + 1] = un;
          mux = ux; muy = uy; muz = uz;
        }
      }
      // queue the leavers of this row behind the ones already queued
      {
        const bool leave = valid && !inb;
        const unsigned lm = __ballot_sync(full, leave);
        if (leave) q[nq + __popc(lm & ((1u << lane) - 1u))] =
            make_int4(__float_as_int(mux), __float_as_int(muy), __float_as_int(muz), i);
        nq += __popc(lm);
      }
      deposit<VARIANT>(a, ii, inb && !(a.dbg & 1), j);
    }
    __syncwarp();

    // finish the queued movers as dense warp-wide batches
    for (int m0 = 0; m0 < nq; m0 += 32) {
      const int m = m0 + lane;
      const bool act = m < nq;
      int i = 0;
      float4 rr = make_float4(0.f, 0.f, 0.f, 0.f), uu = rr;
      float dispx = 0.f, dispy = 0.f, dispz = 0.f;
      if (act) {
        const int4 mv = q[m];
        i = mv.w;
        ld_particle(a.p + 2 * (size_t)i, rr, uu);
        dispx = __int_as_float(mv.x); dispy = __int_as_float(mv.y); dispz = __int_as_float(mv.z);
      }
      int left;
      if (VARIANT == VPB_DEPOSIT_WARP_SEG) left = move_p_warp(a, act, rr, uu, dispx, dispy, dispz);
      else left = act ? move_p_dev(a, rr, uu, dispx, dispy, dispz) : 0;
      if (act) {
        if (left) {
          const int slot = atomicAdd(a.counters, 1);
          if (slot < a.max_nm) {
            a.pm[slot] = make_int4(__float_as_int(dispx), __float_as_int(dispy), __float_as_int(dispz), i);
          } else {
            atomicAdd(a.counters + 1, 1);                             // lost mover: keep p.i a valid voxel
            rr.w = __int_as_float(__float_as_int(rr.w) >> 3);
          }
        }
        st_particle(a.p + 2 * (size_t)i, rr, uu);
      }
    }
    __syncwarp();
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
  const PushK k = to_push_k(args);
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
