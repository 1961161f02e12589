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
// Mapping to the machine (no tensor cores: there is no contraction on this path — see DESIGN.md §3.1):
//   * one thread per particle; every warp is autonomous (no block-wide barrier) and takes spans of kSpan consecutive
//     rows of 32 voxel-sorted particles, so its five read-only interpolator loads are mostly L1 broadcasts/hits;
//   * a particle is one 32-byte sector: one 256-bit load (no L1 allocation) and one 256-bit store each; the next
//     row is requested before the current one is processed;
//   * in-voxel streaks are summed across the warp by voxel before any atomic (deposit_warp_segmented);
//   * particles that leave their voxel are NOT moved inline (one crossing lane would stall its 31 neighbours
//     through the whole streak loop): their full state goes to a per-warp queue in shared memory and is finished
//     32 at a time by the same warp, which keeps both phases convergent;
//   * args.variant selects the deposit strategy (VPB_DEPOSIT_*), args.debug_skip the profiling ablations.
#include "push_common.cuh"
#include <string.h>

namespace vpb {

constexpr int kBlock = 256;
constexpr int kMinBlocks = 3;            // resident CTAs per SM the register budget is tuned for

template <int VARIANT>
__device__ __forceinline__ void deposit(const PushK &a, int vox, bool active, const float (&j)[12], int min_group) {
  if (VARIANT == VPB_DEPOSIT_WARP_SEG || VARIANT == VPB_DEPOSIT_WARP_SEG_MOVERS || VARIANT == VPB_DEPOSIT_WARP_SEG_FIRST) {
    deposit_warp_segmented(a.accum, a.astride, vox, active, j, min_group);
  } else {
    if (active) deposit_red_v4(a.accum + (size_t)vox * a.astride, j);
  }
}

constexpr int kWarps = kBlock / 32;
constexpr int kSpan  = 64;                 // consecutive rows a warp takes before jumping ahead
constexpr int kQCap  = 64;                 // a warp's queue holds at most 31 carried-over + 32 new movers
constexpr size_t kSmemBytes = (size_t)kWarps * 3 * kQCap * sizeof(int4);

// One dense batch of queued movers [start, start+count) of this warp's queue, count <= 32.
template <int VARIANT, bool DBG>
__device__ __forceinline__ void run_movers(const PushK &a, const int4 *q0, const int4 *q1, const int4 *q2,
                                           int start, int count, int lane) {
  const bool act = lane < count;
  int i = 0;
  float4 rr = make_float4(0.f, 0.f, 0.f, 0.f), uu = rr;
  float dispx = 0.f, dispy = 0.f, dispz = 0.f;
  if (act) {
    const int4 w0 = q0[start + lane], w1 = q1[start + lane], w2 = q2[start + lane];
    rr = make_float4(__int_as_float(w0.x), __int_as_float(w0.y), __int_as_float(w0.z), __int_as_float(w0.w));
    uu = make_float4(__int_as_float(w1.x), __int_as_float(w1.y), __int_as_float(w1.z), __int_as_float(w1.w));
    dispx = __int_as_float(w2.x); dispy = __int_as_float(w2.y); dispz = __int_as_float(w2.z);
    i = w2.w;
  }
  int left;
  if (VARIANT == VPB_DEPOSIT_WARP_SEG_MOVERS) left = move_p_warp<false, true>(a, act, rr, uu, dispx, dispy, dispz);
  else if (VARIANT == VPB_DEPOSIT_WARP_SEG_FIRST) left = move_p_warp<true, false>(a, act, rr, uu, dispx, dispy, dispz);
  else left = act ? move_p_dev<DBG>(a, rr, uu, dispx, dispy, dispz) : 0;
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
    st_particle(a.pout + 2 * (size_t)i, rr, uu);
    if (a.keys) a.keys[i] = __float_as_int(rr.w);
  }
}

// DBG = false (every production launch) compiles the profiling switches of args.debug_skip out of the loop.
// GATHER applies the order of an index sort on the fly: position k is loaded from p[perm[k]] (perm one row further
// ahead than the particles) and stored to pout[k].
template <int VARIANT, bool DBG, bool GATHER = false>
__global__ void __launch_bounds__(kBlock, GATHER ? 4 : kMinBlocks) advance_p_kernel(const PushK a) {
  // Every warp is autonomous: it walks rows of 32 consecutive particles (rows of one CTA are adjacent, so its
  // warps share interpolator lines in L1), keeps its own mover queue in shared memory and never meets a block-wide
  // barrier.  The next row's particles are requested before the current row is processed.  A queued mover carries
  // its whole state {r, u, disp, index} (three 16-byte planes, conflict-free), so finishing it needs no reload;
  // movers are finished 32 at a time, leftovers ride along to the warp's next row.
  extern __shared__ int4 s_q[];
  const int dbg = DBG ? a.dbg : 0;
  const int min_group = ((dbg >> 8) & 0xff) ? ((dbg >> 8) & 0xff) : kMinGroup;
  const float one = 1.0f;
  const float one_third = (float)(1.0 / 3.0);
  const float two_fifteenths = (float)(2.0 / 15.0);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_rows = (a.np + 31) / 32;
  const int warps_total = gridDim.x * kWarps;
  int4 *q0 = s_q + (size_t)w * 3 * kQCap, *q1 = q0 + kQCap, *q2 = q1 + kQCap;
  int nq = 0;                                                           // warp-uniform queue length

  // Rows are dealt to warps in spans of kSpan consecutive rows: consecutive rows of voxel-sorted particles share
  // their interpolator (64 ppc = 2 rows per voxel), so a warp's next gather usually hits the lines it just used.
  const int span_rows = a.span;                                         // kSpan, fewer for small species (host)
  const int n_spans = (n_rows + span_rows - 1) / span_rows;
  int span = blockIdx.x * kWarps + w;
  int row = span * span_rows;
  float4 rn = make_float4(0.f, 0.f, 0.f, 0.f), un_next = rn;
  // the row after `r` of this warp: same span, or the first row of its next span
  auto step_row = [&](int &r, int &s) { r++; if (r == (s + 1) * span_rows || r >= n_rows) { s += warps_total; r = s * span_rows; } };
  int src_next = 0;                                                     // GATHER: perm entry of the row after `row`
  if (span < n_spans && row * 32 + lane < a.np)
    ld_particle(a.p + 2 * (size_t)(GATHER ? __ldg(a.perm + row * 32 + lane) : a.first + row * 32 + lane), rn, un_next);
  if (GATHER) {
    int r1 = row, s1 = span;
    step_row(r1, s1);
    if (s1 < n_spans && r1 * 32 + lane < a.np) src_next = __ldg(a.perm + r1 * 32 + lane);
  }

#pragma unroll 1
  while (span < n_spans) {
    const float4 r = rn, u = un_next;
    const int i = a.first + row * 32 + lane;
    const bool valid = row * 32 + lane < a.np;
    // advance to the next row of this warp (same span, or the first row of its next span) and request it now
    int next_row = row, next_span = span;
    step_row(next_row, next_span);
    if (next_span < n_spans && next_row * 32 + lane < a.np)
      ld_particle(a.p + 2 * (size_t)(GATHER ? src_next : a.first + next_row * 32 + lane), rn, un_next);
    if (GATHER) {
      int r2 = next_row, s2 = next_span;
      step_row(r2, s2);
      if (s2 < n_spans && r2 * 32 + lane < a.np) src_next = __ldg(a.perm + r2 * 32 + lane);
    }
    const int ii = __float_as_int(r.w);
    bool inb = false;
    float j[12];
    float mux = 0.f, muy = 0.f, muz = 0.f;
    float4 un = u;
    if (valid) {
      const float4 *f = reinterpret_cast<const float4 *>(a.interp + (size_t)ii * a.istride);
      float4 fex = make_float4(.01f, .02f, .03f, .04f), fey = fex, fez = fex, fb0 = fex;
      float2 fb1 = make_float2(.01f, .02f);
      if (!(dbg & 8)) {
        fex = __ldg(f); fey = __ldg(f + 1); fez = __ldg(f + 2); fb0 = __ldg(f + 3);
        fb1 = __ldg(reinterpret_cast<const float2 *>(f + 4));
      }
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
      un.x = ux; un.y = uy; un.z = uz;                               // new momentum, kept in either case
      v0 = __fdiv_rn(one, __fsqrt_rn(one + (ux * ux + (uy * uy + uz * uz))));
      ux *= a.cdt_dx; uy *= a.cdt_dy; uz *= a.cdt_dz;
      ux *= v0; uy *= v0; uz *= v0;                                 // half displacement in cell units
      v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;                     // streak midpoint
      v3 = v0 + ux; v4 = v1 + uy; const float v5n = v2 + uz;        // new position
      inb = (v3 <= one) && (v4 <= one) && (v5n <= one) && (-v3 <= one) && (-v4 <= one) && (-v5n <= one);
      if (dbg & 2) inb = true;
      if (inb) {
        const float qw = un.w * a.qsp;
        if (!(dbg & 4)) st_particle(a.pout + 2 * (size_t)i, make_float4(v3, v4, v5n, r.w), un);
        if (a.keys) a.keys[i] = ii;
        const float v5 = (((qw * ux) * uy) * uz) * one_third;
        streak_currents(qw, ux, uy, uz, v0, v1, v2, v5, j);
      } else {
        mux = ux; muy = uy; muz = uz;
      }
    }
    // queue the leavers of this row (old position, new momentum, displacement) behind those already queued
    {
      const bool leave = valid && !inb;
      const unsigned lm = __ballot_sync(full, leave);
      if (leave) {
        const int slot = nq + __popc(lm & ((1u << lane) - 1u));
        q0[slot] = make_int4(__float_as_int(r.x), __float_as_int(r.y), __float_as_int(r.z), ii);
        q1[slot] = make_int4(__float_as_int(un.x), __float_as_int(un.y), __float_as_int(un.z), __float_as_int(un.w));
        q2[slot] = make_int4(__float_as_int(mux), __float_as_int(muy), __float_as_int(muz), i);
      }
      nq += __popc(lm);
    }
    deposit<VARIANT>(a, ii, inb && !(dbg & 1), j, min_group);
    if (nq >= 32) {                                                     // a full batch; the rest carries over
      __syncwarp();
      nq -= 32;
      run_movers<VARIANT, DBG>(a, q0, q1, q2, nq, 32, lane);
      __syncwarp();
    }
    row = next_row; span = next_span;
  }
  __syncwarp();
  if (nq > 0) run_movers<VARIANT, DBG>(a, q0, q1, q2, 0, nq, lane);
}

// Check the closed-form neighbour rule against the table for every interior voxel and face.
__global__ void __launch_bounds__(256) verify_neighbor_rule_kernel(PushK a, int *mismatch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1, y = blockIdx.y + 1, z = blockIdx.z + 1;
  if (x > a.nb.nx) return;
  const int vox = voxel(x, y, z, a.nb.nx, a.nb.ny);
  int bad = 0;
#pragma unroll
  for (int f = 0; f < 6; f++) bad += (neighbor_of(a, vox, f) != __ldg(a.neighbor + 6ll * vox + f));
  if (bad) atomicAdd(mismatch, bad);
}

int advance_p_brick(const vpb_push_args_t *args, const PushK &k, cudaStream_t st);   // advance_p_brick.cu

}  // namespace vpb

using namespace vpb;

extern "C" int vpb_neighbor_rule_derive(const int64_t *neighbor_dev, int32_t nx, int32_t ny, int32_t nz, int64_t rangel,
                                        vpb_neighbor_rule_t *rule, void *stream) {
  VPB_REQUIRE(neighbor_dev && rule && nx > 0 && ny > 0 && nz > 0, "vpb_neighbor_rule_derive: Bad args");
  cudaStream_t st = as_stream(stream);
  memset(rule, 0, sizeof *rule);
  rule->nx = nx; rule->ny = ny; rule->nz = nz;
  const int64_t nv = (int64_t)(nx + 2) * (ny + 2) * (nz + 2);
  if (nv >= (1ll << 31) / 6) return 0;               // beyond the reference's own neighbour-index limit
  const int vlo = voxel(1, 1, 1, nx, ny), vhi = voxel(nx, ny, nz, nx, ny);
  long long lo[6], hi[6];
  VPB_CUDA(cudaMemcpyAsync(lo, neighbor_dev + 6ll * vlo, sizeof lo, cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaMemcpyAsync(hi, neighbor_dev + 6ll * vhi, sizeof hi, cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaStreamSynchronize(st));
  for (int f = 0; f < 6; f++) {
    const long long act = f < 3 ? lo[f] : hi[f];
    const int vref = f < 3 ? vlo : vhi;
    rule->act[f] = act;
    rule->delta[f] = act < 0 ? 0 : act - (rangel + vref);
  }
  rule->valid = 1;                                   // provisional, for the check below
  vpb_push_args_t pa; memset(&pa, 0, sizeof pa);
  pa.neighbor = neighbor_dev; pa.rangel = rangel; pa.nx = nx; pa.ny = ny; pa.nz = nz; pa.neighbor_rule = rule;
  const PushK k = to_push_k(&pa);
  int *d_bad = nullptr, h_bad = 0;
  VPB_CUDA(cudaMalloc(&d_bad, sizeof(int)));
  VPB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  dim3 grid((nx + 255) / 256, ny, nz);
  verify_neighbor_rule_kernel<<<grid, 256, 0, st>>>(k, d_bad);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_bad);
  VPB_CUDA(e);
  rule->valid = (h_bad == 0);
  return 0;
}


extern "C" int vpb_advance_p(const vpb_push_args_t *args, void *stream) {
  VPB_REQUIRE(args && args->p && args->interp && args->accum && args->neighbor && args->counters,
              "vpb_advance_p: Bad args.");
  VPB_REQUIRE(args->pm || args->max_nm == 0, "vpb_advance_p: mover array missing");
  VPB_REQUIRE(args->interp_stride % 4 == 0 && args->accum_stride % 4 == 0,
              "vpb_advance_p: strides must keep 16-byte alignment (got %d, %d)", args->interp_stride, args->accum_stride);
  VPB_REQUIRE((((uintptr_t)args->interp | (uintptr_t)args->accum | (uintptr_t)args->pm) & 15) == 0 &&
              ((uintptr_t)args->p & 31) == 0 && args->p_first >= 0,
              "vpb_advance_p: particles must be 32-byte aligned, the other arrays 16-byte aligned");
  VPB_REQUIRE(!args->perm || (args->p_out && args->p_out != args->p && ((uintptr_t)args->p_out & 31) == 0 && args->p_first == 0 &&
                              args->debug_skip == 0),
              "vpb_advance_p: perm needs a separate 32-byte aligned p_out and p_first == 0");
  if (args->np <= 0) return 0;
  PushK k = to_push_k(args);
  // Default strategy: the linear kernel below (warp-segmented reduction + vector REDs).  The brick/tile kernel is
  // bit-identical in particle state but measured slower on B200 in round 2, so it runs only when
  // asked for (variant VPB_DEPOSIT_BRICK_TILE, or VPB_BRICK_DEFAULT=1 in the environment).
  static int brick_default = -1;
  if (brick_default < 0) { const char *e = getenv("VPB_BRICK_DEFAULT"); brick_default = e && atoi(e) != 0; }
  if (!args->perm && !args->keys_out && (args->variant == VPB_DEPOSIT_BRICK_TILE || (args->variant == VPB_DEPOSIT_DEFAULT && brick_default))) {
    VPB_REQUIRE(args->nx > 0 && args->ny > 0 && args->nz > 0, "vpb_advance_p: Bad grid");
    const int served = (args->debug_skip == 0) ? advance_p_brick(args, k, as_stream(stream)) : 0;
    if (served < 0) return -1;
    if (served > 0) return 0;
  }
  static size_t extra_smem = 0;          // profiling: VPB_EXTRA_SMEM_KB pads the CTA's shared memory to lower occupancy
  static bool env_done = false;
  if (!env_done) { if (const char *e = getenv("VPB_EXTRA_SMEM_KB")) extra_smem = (size_t)atoi(e) * 1024; env_done = true; }
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = kSMs; }
  const bool dbg = args->debug_skip != 0;
  const int n_rows = (args->np + 31) / 32;
  // spans of kSpan rows per warp; small species get shorter spans so that every SM still has warps to run
  int span_rows = ((args->debug_skip >> 16) & 0xff) ? ((args->debug_skip >> 16) & 0xff) : kSpan;
  if (!((args->debug_skip >> 16) & 0xff)) {
    const int fill = n_rows / (sms * 32);
    if (fill < span_rows) span_rows = fill < 4 ? 4 : fill;
  }
  k.span = span_rows;
  const int n_spans = (n_rows + span_rows - 1) / span_rows;
  // one warp per span of rows, capped at a grid of many short CTAs (side-stream kernels slot in between them)
  const int gmul = ((args->debug_skip >> 24) & 0xff) ? ((args->debug_skip >> 24) & 0xff) : 32;     // profiling override
  const int need = (n_spans + kWarps - 1) / kWarps;
  const int grid = need < sms * gmul ? need : sms * gmul;
  int variant = (args->variant == VPB_DEPOSIT_DEFAULT || args->variant == VPB_DEPOSIT_BRICK_TILE || args->perm) ? VPB_DEPOSIT_WARP_SEG : args->variant;
  const size_t smem = kSmemBytes + (variant == VPB_DEPOSIT_WARP_SEG ? extra_smem : 0);
#define VPB_LAUNCH_AP(V, D) do {                                                                                      \
    static bool attr_done = false;                                                                                    \
    if (!attr_done) { VPB_CUDA(cudaFuncSetAttribute(advance_p_kernel<V, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; } \
    advance_p_kernel<V, D><<<grid, kBlock, smem, as_stream(stream)>>>(k); } while (0)
  switch (variant) {
    case VPB_DEPOSIT_RED_V4:          if (dbg) VPB_LAUNCH_AP(VPB_DEPOSIT_RED_V4, true); else VPB_LAUNCH_AP(VPB_DEPOSIT_RED_V4, false); break;
    case VPB_DEPOSIT_WARP_SEG:
      if (args->perm) {
        static bool gattr = false;
        if (!gattr) { VPB_CUDA(cudaFuncSetAttribute(advance_p_kernel<VPB_DEPOSIT_WARP_SEG, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); gattr = true; }
        advance_p_kernel<VPB_DEPOSIT_WARP_SEG, false, true><<<grid, kBlock, smem, as_stream(stream)>>>(k);
      } else if (dbg) VPB_LAUNCH_AP(VPB_DEPOSIT_WARP_SEG, true); else VPB_LAUNCH_AP(VPB_DEPOSIT_WARP_SEG, false);
      break;
    case VPB_DEPOSIT_WARP_SEG_MOVERS: if (dbg) VPB_LAUNCH_AP(VPB_DEPOSIT_WARP_SEG_MOVERS, true); else VPB_LAUNCH_AP(VPB_DEPOSIT_WARP_SEG_MOVERS, false); break;
    case VPB_DEPOSIT_WARP_SEG_FIRST:  if (dbg) VPB_LAUNCH_AP(VPB_DEPOSIT_WARP_SEG_FIRST, true); else VPB_LAUNCH_AP(VPB_DEPOSIT_WARP_SEG_FIRST, false); break;
    default:
      VPB_REQUIRE(false, "vpb_advance_p: unknown deposit variant %d", variant);
  }
#undef VPB_LAUNCH_AP
  VPB_LAUNCH_CHECK();
  return 0;
}
