// advance_p for voxel-sorted particles, round-2 design: bricks of voxels, a warp-private accumulator tile in shared
// memory, two particles per thread.
//
// Replaces the same reference code as advance_p.cu (advance_p_pipeline.cc:20-340, move_p.cc:216-378) and produces the
// same particle bytes; it is chosen by vpb_advance_p when the caller passes the partition[] of the species' last
// sort_p (vpb_push_args_t.partition).  What changes is where the current deposits go.
//
//   * Work item = one BRICK of BX x BY x BZ voxels.  partition[] tells where the particles that were in those voxels
//     at the last sort sit in the array: one contiguous segment per x-line of the brick.  A warp walks the brick's
//     segments as one dense sequence of rows of 32 particles (256-bit loads/stores, one particle per sector).
//   * The warp owns an accumulator TILE in shared memory covering the brick plus a halo of H voxels on every side.
//     Deposits that land in the tile are plain shared-memory read-modify-writes (LDS.128/FADD/STS.128) — no atomics:
//     the tile is private to the warp, and lanes that hit the same voxel in the same row are either summed first
//     (warp reduce-scatter, the sorted case) or take turns (a 128-entry tag table decides who goes in each round).
//     Shared-memory float atomics are CAS loops on sm_100a and even native ATOMS costs 2 cycles per lane, more than a
//     global RED — that is why the tile is warp-private instead of CTA-shared (tools/ubench_r2.cu: 165 G tile
//     increments/s against 59 G through global REDs).
//   * When the brick is done the tile is added to the global accumulator array with one TMA bulk reduce per tile row
//     (cp.reduce.async.bulk ... add.f32, SASS UBLKRED; measured 5.3 TB/s chip-wide), 6 bytes per particle instead of
//     a 48-byte RED per drifted particle.  Particles that have drifted out of the tile since the last sort, and
//     particles that are not covered by partition[] (appended by boundary_p since the sort), fall back to global REDs.
//     Correctness never depends on partition[] being accurate: the bricks' segments plus the tail cover every index
//     exactly once as long as partition[] is monotone.
//   * Every thread advances TWO particles (rows r and r+1 of the brick): independent dependency chains instead of
//     resident warps hide the latency (the tile costs occupancy), and the two pushes can run as packed FFMA2
//     instructions (packed_f32.cuh) that halve the issue slots of the arithmetic.
//   * Movers (particles that leave their voxel) are queued per warp as in advance_p.cu and finished 32 at a time;
//     their streaks deposit into the tile as well.
#include "push_common.cuh"
#include "packed_f32.cuh"
#include <string.h>
#include <stdlib.h>

namespace vpb {

// ---------------------------------------------------------------------------------------------------------------
// arithmetic on pairs of particles: two scalar instruction streams, or packed FFMA2
struct ArithScalar {
  struct V { float a, b; };
  __device__ __forceinline__ V make(float a, float b) const { V v; v.a = a; v.b = b; return v; }
  __device__ __forceinline__ V bc(float c) const { return make(c, c); }
  __device__ __forceinline__ float lo(V v) const { return v.a; }
  __device__ __forceinline__ float hi(V v) const { return v.b; }
  __device__ __forceinline__ V mul(V x, V y) const { return make(x.a * y.a, x.b * y.b); }
  __device__ __forceinline__ V add(V x, V y) const { return make(x.a + y.a, x.b + y.b); }
  __device__ __forceinline__ V sub(V x, V y) const { return make(x.a - y.a, x.b - y.b); }
  template <bool SAFE> __device__ __forceinline__ V div(V x, V y) const { return make(__fdiv_rn(x.a, y.a), __fdiv_rn(x.b, y.b)); }
  __device__ __forceinline__ V sqrt(V x) const { return make(__fsqrt_rn(x.a), __fsqrt_rn(x.b)); }
};

struct ArithPacked {
  typedef f2 V;
  F2Const k;
  __device__ __forceinline__ V make(float a, float b) const { return pk(a, b); }
  __device__ __forceinline__ V bc(float c) const { return pk(c, c); }
  __device__ __forceinline__ float lo(V v) const { return lo_of(v); }
  __device__ __forceinline__ float hi(V v) const { return hi_of(v); }
  __device__ __forceinline__ V mul(V x, V y) const { return mul2(k, x, y); }
  __device__ __forceinline__ V add(V x, V y) const { return add2(k, x, y); }
  __device__ __forceinline__ V sub(V x, V y) const { return sub2(k, x, y); }
  template <bool SAFE> __device__ __forceinline__ V div(V x, V y) const { return div2<SAFE>(k, x, y); }
  __device__ __forceinline__ V sqrt(V x) const { return sqrt2(k, x); }
};

// The 12 accumulator increments of a straight streak for a pair of particles (same expressions as streak_currents).
template <class AR>
__device__ __forceinline__ void streak_currents_pair(const AR &ar, typename AR::V q, typename AR::V ux, typename AR::V uy,
                                                     typename AR::V uz, typename AR::V dx, typename AR::V dy,
                                                     typename AR::V dz, typename AR::V v5, typename AR::V (&j)[12]) {
  typedef typename AR::V V;
  const V one = ar.bc(1.0f);
  V v0, v1, v2, v3, v4;
#define VPB_ACC2(uX, dY, dZ, o)                                                     \
  v4 = ar.mul(q, uX); v1 = ar.mul(v4, dY); v0 = ar.sub(v4, v1); v1 = ar.add(v1, v4); \
  v4 = ar.add(one, dZ); v2 = ar.mul(v0, v4); v3 = ar.mul(v1, v4);                   \
  v4 = ar.sub(one, dZ); v0 = ar.mul(v0, v4); v1 = ar.mul(v1, v4);                   \
  v0 = ar.add(v0, v5); v1 = ar.sub(v1, v5); v2 = ar.sub(v2, v5); v3 = ar.add(v3, v5); \
  j[o] = v0; j[o + 1] = v1; j[o + 2] = v2; j[o + 3] = v3;
  VPB_ACC2(ux, dy, dz, 0)
  VPB_ACC2(uy, dz, dx, 4)
  VPB_ACC2(uz, dx, dy, 8)
#undef VPB_ACC2
}

// ---------------------------------------------------------------------------------------------------------------
struct BrickK {
  const int *part;            // partition[nv+1] of the last sort_p
  int nbx, nby, nbz, n_bricks;
  int tail0, n_tail;          // particles [tail0, np) are not covered by partition[]: n_tail linear spans
  int *work;                  // work-item counter (zero at launch)
  f2 one2, nz2, mone2;
};

constexpr int kTailRows = 64;         // rows per tail span
constexpr int kTagSlots = 128;
constexpr int kMaxSeg = 40;           // (BY+2)*(BZ+2) for the largest brick cross-section used (4x4 -> 36)
constexpr int kBQCap = 64;            // mover queue: at most 31 carried over + 32 new

// Per-warp shared-memory context
struct Tile {
  float *acc;                 // tile accumulators, TX*TY*TZ voxels of astride floats
  int *tag;                   // kTagSlots ints
  int on;                     // 0: this work item has no tile (tail spans), every deposit goes to global memory
  int cx, cy, cz;             // coordinates of the tile's corner voxel (may lie outside the array at the domain edge)
};

// Tile slot of a voxel, or -1 when it lies outside the tile.  The voxel's own coordinates are decoded (not the offset
// from the tile corner), so a tile that sticks out of a thin grid can never alias two voxels.
template <int TX, int TY, int TZ>
__device__ __forceinline__ int tile_index(const PushK &a, const Tile &t, int vox) {
  if (!t.on) return -1;
  const int z = fast_div(vox, a.nb.sz, a.nb.inv_sz);
  const int r = vox - z * a.nb.sz;
  const int y = fast_div(r, a.nb.sy, a.nb.inv_sy);
  const int qx = r - y * a.nb.sy - t.cx, qy = y - t.cy, qz = z - t.cz;
  if ((unsigned)qx >= (unsigned)TX || (unsigned)qy >= (unsigned)TY || (unsigned)qz >= (unsigned)TZ) return -1;
  return (qz * TY + qy) * TX + qx;
}

__device__ __forceinline__ void rmw48(float *t, const float (&j)[12]) {
  float4 *v = reinterpret_cast<float4 *>(t);
  float4 a0 = v[0], a1 = v[1], a2 = v[2];
  a0.x += j[0]; a0.y += j[1]; a0.z += j[2]; a0.w += j[3];
  a1.x += j[4]; a1.y += j[5]; a1.z += j[6]; a1.w += j[7];
  a2.x += j[8]; a2.y += j[9]; a2.z += j[10]; a2.w += j[11];
  v[0] = a0; v[1] = a1; v[2] = a2;
}

// Sum the 12 values of the lanes in `grp` (reduce-scatter butterfly) and add the totals to voxel gv: into the tile
// when tidx >= 0, else as scalar REDs.
__device__ __forceinline__ void deposit_group(const PushK &a, const Tile &t, unsigned grp, bool mine, int gv, int tidx,
                                              const float (&j)[12]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  float v[16];
#pragma unroll
  for (int c = 0; c < 12; c++) v[c] = mine ? j[c] : 0.0f;
#pragma unroll
  for (int c = 12; c < 16; c++) v[c] = 0.0f;
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
  if (!(lane & 1) && comp < 12) {
    if (tidx >= 0) t.acc[tidx * a.astride + comp] += tot;
    else red_add(a.accum + (size_t)gv * a.astride + comp, tot);
  }
  (void)grp;
}

// Warp-collective deposit of one streak per active lane into voxel `vox`.
//   * the two most common voxels of the row (voxel-sorted rows hold one or two) are summed across the warp when they
//     have >= kMinGroup lanes;
//   * every other lane deposits on its own: into the tile by read-modify-write, lanes that share a voxel taking turns
//     (tag table), or straight to global memory with three vector REDs when the voxel is outside the tile.
template <int TX, int TY, int TZ>
__device__ __forceinline__ void deposit_any(const PushK &a, const Tile &t, int vox, bool active, const float (&j)[12]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned amask = __ballot_sync(full, active);
  if (amask == 0) return;
  const int tidx = active ? tile_index<TX, TY, TZ>(a, t, vox) : -1;
  bool done = !active;
  // first and second voxel of the row
  unsigned rest = amask;
#pragma unroll 1
  for (int g = 0; g < 2 && rest; g++) {
    const int leader = __ffs(rest) - 1;
    const int gv = __shfl_sync(full, vox, leader);
    const int gt = __shfl_sync(full, tidx, leader);
    const bool mine = !done && vox == gv;
    const unsigned grp = __ballot_sync(full, mine);
    if (__popc(grp) >= kMinGroup) {
      deposit_group(a, t, grp, mine, gv, gt, j);
      if (mine) done = true;
    }
    rest &= ~grp;
  }
  // stragglers
  if (!done && tidx < 0) { deposit_red_v4(a.accum + (size_t)vox * a.astride, j); done = true; }
  unsigned pend = __ballot_sync(full, !done);
  const int slot = tidx & (kTagSlots - 1);
#pragma unroll 1
  while (pend) {
    if (!done) t.tag[slot] = lane;
    __syncwarp();
    const bool go = !done && t.tag[slot] == lane;
    __syncwarp();
    if (go) { rmw48(t.acc + tidx * a.astride, j); done = true; }
    pend = __ballot_sync(full, !done);
  }
  __syncwarp();
}

// One dense batch of queued movers [start, start+count), count <= 32: the reference's move_p streak loop, every streak
// deposited through deposit_any (MOVER_TILE) or as three vector REDs.
template <int TX, int TY, int TZ, bool MOVER_TILE>
__device__ __forceinline__ void run_movers_b(const PushK &a, const Tile &t, const int4 *q0, const int4 *q1, const int4 *q2,
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
  const float q = a.qsp * uu.w;
  int vox = __float_as_int(rr.w);
  int st = act ? 2 : 0;
  if (MOVER_TILE) {
    while (__any_sync(0xffffffffu, st == 2)) {
      float j[12];
      const int dep_vox = vox;
      const bool dep = (st == 2);
      if (dep) st = streak_step(a, q, rr, uu, vox, dispx, dispy, dispz, j);
      deposit_any<TX, TY, TZ>(a, t, dep_vox, dep, j);
    }
  } else {
    while (st == 2) {
      float j[12];
      const int dep_vox = vox;
      st = streak_step(a, q, rr, uu, vox, dispx, dispy, dispz, j);
      deposit_red_v4(a.accum + (size_t)dep_vox * a.astride, j);
    }
  }
  rr.w = __int_as_float(vox);
  if (act) {
    if (st == 1) {
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

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------------------
template <int BX, int BY, int BZ, int H, int WARPS, int MINB, class AR, bool MOVER_TILE>
__global__ void __launch_bounds__(WARPS * 32, MINB) advance_p_brick_kernel(const PushK a, const BrickK b) {
  constexpr int TX = BX + 2 * H, TY = BY + 2 * H, TZ = BZ + 2 * H, TVOX = TX * TY * TZ;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // per-warp shared memory: tile | mover queue (3 planes) | tag table | segment tables
  const size_t tile_bytes = (size_t)TVOX * a.astride * sizeof(float);
  const size_t warp_bytes = tile_bytes + 3 * kBQCap * sizeof(int4) + kTagSlots * sizeof(int) + (2 * kMaxSeg + 8) * sizeof(int);
  unsigned char *base = s_raw + (size_t)w * warp_bytes;
  Tile t;
  t.acc = reinterpret_cast<float *>(base);
  int4 *q0 = reinterpret_cast<int4 *>(base + tile_bytes), *q1 = q0 + kBQCap, *q2 = q1 + kBQCap;
  t.tag = reinterpret_cast<int *>(q2 + kBQCap);
  int *seg_start = t.tag + kTagSlots, *seg_cum = seg_start + kMaxSeg;      // seg_cum has kMaxSeg+1 entries
  AR ar;
  if constexpr (sizeof(AR) > 1) { ar.k.one = b.one2; ar.k.nz = b.nz2; ar.k.mone = b.mone2; }
  typedef typename AR::V V;
  const float one = 1.0f;
  const V ONE = ar.bc(1.0f), OT = ar.bc((float)(1.0 / 3.0)), TF = ar.bc((float)(2.0 / 15.0));
  const V Q = ar.bc(a.qdt_2mc), QSP = ar.bc(a.qsp), CX = ar.bc(a.cdt_dx), CY = ar.bc(a.cdt_dy), CZ = ar.bc(a.cdt_dz);
  const int sy = a.nb.sy, sz = a.nb.sz, nx = a.nb.nx, ny = a.nb.ny, nz = a.nb.nz;
  const int n_items = b.n_bricks + b.n_tail;
  int nq = 0;                                                            // warp-uniform mover-queue length
  // zero the tile once; every flush re-zeroes it
  for (int k = lane; k < (int)(tile_bytes / 16); k += 32) reinterpret_cast<float4 *>(t.acc)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();

#pragma unroll 1
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(b.work, 1);
    item = __shfl_sync(full, item, 0);
    if (item >= n_items) break;
    // ---- describe the work item: segments of the particle array, and where the tile sits
    int nseg, n_v;
    const bool tiled = item < b.n_bricks;
    if (tiled) {
      const int bx = item % b.nbx, by = (item / b.nbx) % b.nby, bz = item / (b.nbx * b.nby);
      const int x0 = 1 + bx * BX, y0 = 1 + by * BY, z0 = 1 + bz * BZ;
      // the first and last brick of an axis also take the ghost layer, so that the bricks cover every voxel
      const int xlo = bx == 0 ? 0 : x0, xhi = bx == b.nbx - 1 ? nx + 2 : x0 + BX;
      const int ylo = by == 0 ? 0 : y0, yhi = by == b.nby - 1 ? ny + 2 : y0 + BY;
      const int zlo = bz == 0 ? 0 : z0, zhi = bz == b.nbz - 1 ? nz + 2 : z0 + BZ;
      const int nyl = yhi - ylo;
      nseg = nyl * (zhi - zlo);
      for (int s = lane; s < nseg; s += 32) {
        const int yl = ylo + s % nyl, zl = zlo + s / nyl;
        const int vlo = xlo + sy * yl + sz * zl, vhi = (xhi - 1) + sy * yl + sz * zl + 1;
        const int p0 = min(__ldg(b.part + vlo), a.np), p1 = min(__ldg(b.part + vhi), a.np);
        seg_start[s] = p0;
        seg_cum[s + 1] = max(p1 - p0, 0);
      }
      t.cx = x0 - H; t.cy = y0 - H; t.cz = z0 - H; t.on = 1;
    } else {
      const int s0 = b.tail0 + (item - b.n_bricks) * kTailRows * 32;
      nseg = 1;
      if (lane == 0) { seg_start[0] = s0; seg_cum[1] = min(kTailRows * 32, a.np - s0); }
      t.on = 0; t.cx = t.cy = t.cz = 0;                                 // no tile: every deposit goes to global memory
    }
    __syncwarp();
    if (lane == 0) {
      int c = 0;
      seg_cum[0] = 0;
      for (int s = 0; s < nseg; s++) { c += seg_cum[s + 1]; seg_cum[s + 1] = c; }
    }
    __syncwarp();
    n_v = seg_cum[nseg];
    const int n_pr = (n_v + 63) >> 6;

    // ---- walk the item's particles two rows at a time
    int js = 0;                                                          // this lane's segment cursor (prefetch stream)
    int idxA = -1, idxB = -1;                                            // array index of the prefetched pair, -1 = none
    float4 rA, uA, rB, uB;
    rA = uA = rB = uB = make_float4(0.f, 0.f, 0.f, 0.f);
    auto locate = [&](int k) -> int {
      if (k >= n_v) return -1;
      while (k >= seg_cum[js + 1]) js++;
      return seg_start[js] + (k - seg_cum[js]);
    };
    idxA = locate(lane);
    idxB = locate(lane + 32);
    if (idxA >= 0) ld_particle(a.p + 2 * (size_t)idxA, rA, uA);
    if (idxB >= 0) ld_particle(a.p + 2 * (size_t)idxB, rB, uB);

#pragma unroll 1
    for (int pr = 0; pr < n_pr; pr++) {
      const int iA = idxA, iB = idxB;
      const bool validA = iA >= 0, validB = iB >= 0;
      const float4 r0A = rA, u0A = uA, r0B = rB, u0B = uB;
      // request the next pair of rows now
      idxA = locate((pr + 1) * 64 + lane);
      idxB = locate((pr + 1) * 64 + 32 + lane);
      if (idxA >= 0) ld_particle(a.p + 2 * (size_t)idxA, rA, uA);
      if (idxB >= 0) ld_particle(a.p + 2 * (size_t)idxB, rB, uB);

      // a lane without a particle computes on its partner's voxel (results unused)
      const int voxA = validA ? __float_as_int(r0A.w) : (validB ? __float_as_int(r0B.w) : 0);
      const int voxB = validB ? __float_as_int(r0B.w) : voxA;
      const float4 *fA = reinterpret_cast<const float4 *>(a.interp + (size_t)voxA * a.istride);
      const float4 *fB = reinterpret_cast<const float4 *>(a.interp + (size_t)voxB * a.istride);
      const float4 exA = __ldg(fA), eyA = __ldg(fA + 1), ezA = __ldg(fA + 2), b0A = __ldg(fA + 3);
      const float2 b1A = __ldg(reinterpret_cast<const float2 *>(fA + 4));
      const float4 exB = __ldg(fB), eyB = __ldg(fB + 1), ezB = __ldg(fB + 2), b0B = __ldg(fB + 3);
      const float2 b1B = __ldg(reinterpret_cast<const float2 *>(fB + 4));

      // ---- Boris push of both particles, the reference's scalar association (advance_p_pipeline.cc:91-162)
      const V dx = ar.make(r0A.x, r0B.x), dy = ar.make(r0A.y, r0B.y), dz = ar.make(r0A.z, r0B.z);
      V ux = ar.make(u0A.x, u0B.x), uy = ar.make(u0A.y, u0B.y), uz = ar.make(u0A.z, u0B.z);
      const V hax = ar.mul(Q, ar.add(ar.add(ar.make(exA.x, exB.x), ar.mul(dy, ar.make(exA.y, exB.y))),
                                     ar.mul(dz, ar.add(ar.make(exA.z, exB.z), ar.mul(dy, ar.make(exA.w, exB.w))))));
      const V hay = ar.mul(Q, ar.add(ar.add(ar.make(eyA.x, eyB.x), ar.mul(dz, ar.make(eyA.y, eyB.y))),
                                     ar.mul(dx, ar.add(ar.make(eyA.z, eyB.z), ar.mul(dz, ar.make(eyA.w, eyB.w))))));
      const V haz = ar.mul(Q, ar.add(ar.add(ar.make(ezA.x, ezB.x), ar.mul(dx, ar.make(ezA.y, ezB.y))),
                                     ar.mul(dy, ar.add(ar.make(ezA.z, ezB.z), ar.mul(dx, ar.make(ezA.w, ezB.w))))));
      const V cbx = ar.add(ar.make(b0A.x, b0B.x), ar.mul(dx, ar.make(b0A.y, b0B.y)));
      const V cby = ar.add(ar.make(b0A.z, b0B.z), ar.mul(dy, ar.make(b0A.w, b0B.w)));
      const V cbz = ar.add(ar.make(b1A.x, b1B.x), ar.mul(dz, ar.make(b1A.y, b1B.y)));
      ux = ar.add(ux, hax); uy = ar.add(uy, hay); uz = ar.add(uz, haz);
      V v0 = ar.template div<true>(Q, ar.sqrt(ar.add(ONE, ar.add(ar.mul(ux, ux), ar.add(ar.mul(uy, uy), ar.mul(uz, uz))))));
      V v1 = ar.add(ar.mul(cbx, cbx), ar.add(ar.mul(cby, cby), ar.mul(cbz, cbz)));
      V v2 = ar.mul(ar.mul(v0, v0), v1);
      V v3 = ar.mul(v0, ar.add(ONE, ar.mul(v2, ar.add(OT, ar.mul(v2, TF)))));
      V v4 = ar.template div<false>(v3, ar.add(ONE, ar.mul(v1, ar.mul(v3, v3))));
      v4 = ar.add(v4, v4);
      v0 = ar.add(ux, ar.mul(v3, ar.sub(ar.mul(uy, cbz), ar.mul(uz, cby))));
      v1 = ar.add(uy, ar.mul(v3, ar.sub(ar.mul(uz, cbx), ar.mul(ux, cbz))));
      v2 = ar.add(uz, ar.mul(v3, ar.sub(ar.mul(ux, cby), ar.mul(uy, cbx))));
      ux = ar.add(ux, ar.mul(v4, ar.sub(ar.mul(v1, cbz), ar.mul(v2, cby))));
      uy = ar.add(uy, ar.mul(v4, ar.sub(ar.mul(v2, cbx), ar.mul(v0, cbz))));
      uz = ar.add(uz, ar.mul(v4, ar.sub(ar.mul(v0, cby), ar.mul(v1, cbx))));
      ux = ar.add(ux, hax); uy = ar.add(uy, hay); uz = ar.add(uz, haz);
      const float4 unA = make_float4(ar.lo(ux), ar.lo(uy), ar.lo(uz), u0A.w);     // new momentum, kept in either case
      const float4 unB = make_float4(ar.hi(ux), ar.hi(uy), ar.hi(uz), u0B.w);
      v0 = ar.template div<true>(ONE, ar.sqrt(ar.add(ONE, ar.add(ar.mul(ux, ux), ar.add(ar.mul(uy, uy), ar.mul(uz, uz))))));
      ux = ar.mul(ux, CX); uy = ar.mul(uy, CY); uz = ar.mul(uz, CZ);
      ux = ar.mul(ux, v0); uy = ar.mul(uy, v0); uz = ar.mul(uz, v0);              // half displacement in cell units
      const V mx = ar.add(dx, ux), my = ar.add(dy, uy), mz = ar.add(dz, uz);      // streak midpoint
      const V px = ar.add(mx, ux), py = ar.add(my, uy), pz = ar.add(mz, uz);      // new position
      const float pxA = ar.lo(px), pyA = ar.lo(py), pzA = ar.lo(pz), pxB = ar.hi(px), pyB = ar.hi(py), pzB = ar.hi(pz);
      const bool inbA = validA && (pxA <= one) && (pyA <= one) && (pzA <= one) && (-pxA <= one) && (-pyA <= one) && (-pzA <= one);
      const bool inbB = validB && (pxB <= one) && (pyB <= one) && (pzB <= one) && (-pxB <= one) && (-pyB <= one) && (-pzB <= one);
      if (inbA) st_particle(a.p + 2 * (size_t)iA, make_float4(pxA, pyA, pzA, r0A.w), unA);
      if (inbB) st_particle(a.p + 2 * (size_t)iB, make_float4(pxB, pyB, pzB, r0B.w), unB);

      // ---- queue the leavers (old position, new momentum, half displacement) behind those already queued
      {
        const bool leaveA = validA && !inbA, leaveB = validB && !inbB;
        const unsigned lmA = __ballot_sync(full, leaveA), lmB = __ballot_sync(full, leaveB);
        if (leaveA) {
          const int slot = nq + __popc(lmA & ((1u << lane) - 1u));
          q0[slot] = make_int4(__float_as_int(r0A.x), __float_as_int(r0A.y), __float_as_int(r0A.z), voxA);
          q1[slot] = make_int4(__float_as_int(unA.x), __float_as_int(unA.y), __float_as_int(unA.z), __float_as_int(unA.w));
          q2[slot] = make_int4(__float_as_int(ar.lo(ux)), __float_as_int(ar.lo(uy)), __float_as_int(ar.lo(uz)), iA);
        }
        nq += __popc(lmA);
        if (nq + __popc(lmB) > kBQCap) {                                // rare: make room for B's leavers first
          __syncwarp();
          nq -= 32;
          run_movers_b<TX, TY, TZ, MOVER_TILE>(a, t, q0, q1, q2, nq, 32, lane);
          __syncwarp();
        }
        if (leaveB) {
          const int slot = nq + __popc(lmB & ((1u << lane) - 1u));
          q0[slot] = make_int4(__float_as_int(r0B.x), __float_as_int(r0B.y), __float_as_int(r0B.z), voxB);
          q1[slot] = make_int4(__float_as_int(unB.x), __float_as_int(unB.y), __float_as_int(unB.z), __float_as_int(unB.w));
          q2[slot] = make_int4(__float_as_int(ar.hi(ux)), __float_as_int(ar.hi(uy)), __float_as_int(ar.hi(uz)), iB);
        }
        nq += __popc(lmB);
      }
      // ---- in-voxel deposits (advance_p_pipeline.cc:172-208)
      {
        float jA[12], jB[12];
        {
          const V qw = ar.mul(ar.make(u0A.w, u0B.w), QSP);
          const V v5 = ar.mul(ar.mul(ar.mul(ar.mul(qw, ux), uy), uz), OT);
          V j[12];
          streak_currents_pair(ar, qw, ux, uy, uz, mx, my, mz, v5, j);
#pragma unroll
          for (int c = 0; c < 12; c++) { jA[c] = ar.lo(j[c]); jB[c] = ar.hi(j[c]); }
        }
        // one voxel for the whole pair of rows (common right after a sort): one reduction serves both particles
        const unsigned am = __ballot_sync(full, inbA), bm = __ballot_sync(full, inbB);
        if (am | bm) {
          const int vfirst = am ? __shfl_sync(full, voxA, __ffs(am) - 1) : __shfl_sync(full, voxB, __ffs(bm) - 1);
          const bool same = __all_sync(full, (!inbA || voxA == vfirst) && (!inbB || voxB == vfirst));
          if (same) {
            float jj[12];
#pragma unroll
            for (int c = 0; c < 12; c++) jj[c] = (inbA ? jA[c] : 0.0f) + (inbB ? jB[c] : 0.0f);
            deposit_any<TX, TY, TZ>(a, t, vfirst, inbA || inbB, jj);
          } else {
            deposit_any<TX, TY, TZ>(a, t, voxA, inbA, jA);
            deposit_any<TX, TY, TZ>(a, t, voxB, inbB, jB);
          }
        }
      }
#pragma unroll 1
      while (nq >= 32) {
        __syncwarp();
        nq -= 32;
        run_movers_b<TX, TY, TZ, MOVER_TILE>(a, t, q0, q1, q2, nq, 32, lane);
        __syncwarp();
      }
    }
    // ---- the item's leftover movers deposit into this item's tile too
    __syncwarp();
    if (nq > 0) { run_movers_b<TX, TY, TZ, MOVER_TILE>(a, t, q0, q1, q2, 0, nq, lane); nq = 0; }
    __syncwarp();
    // ---- flush the tile: one bulk reduce-add per tile row that lies inside the array, then re-zero
    if (tiled) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      const int xl = max(t.cx, 0), xr = min(t.cx + TX, nx + 2);
      const unsigned bytes = (unsigned)((xr - xl) * a.astride * sizeof(float));
      for (int row = lane; row < TY * TZ; row += 32) {
        const int y = t.cy + row % TY, z = t.cz + row / TY;
        if (y < 0 || y > ny + 1 || z < 0 || z > nz + 1) continue;
        float *gp = a.accum + (size_t)(xl + sy * y + sz * z) * a.astride;
        const float *sp = t.acc + (size_t)(row * TX + (xl - t.cx)) * a.astride;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                     :: "l"(gp), "r"(smem_addr(sp)), "r"(bytes) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      for (int k = lane; k < (int)(tile_bytes / 16); k += 32) reinterpret_cast<float4 *>(t.acc)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
template <int BX, int BY, int BZ, int H, int WARPS, int MINB, class AR, bool MT>
static int launch_brick(const PushK &k, BrickK b, int astride, int sms, cudaStream_t st) {
  constexpr int TVOX = (BX + 2 * H) * (BY + 2 * H) * (BZ + 2 * H);
  const size_t warp_bytes = (size_t)TVOX * astride * sizeof(float) + 3 * kBQCap * sizeof(int4) + kTagSlots * sizeof(int) +
                            (2 * kMaxSeg + 8) * sizeof(int);
  const size_t smem = warp_bytes * WARPS;
  auto kern = advance_p_brick_kernel<BX, BY, BZ, H, WARPS, MINB, AR, MT>;
  static bool attr_done = false;
  if (!attr_done) {
    VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    attr_done = true;
  }
  int per_sm = 0;
  VPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
  VPB_REQUIRE(per_sm > 0, "advance_p brick kernel does not fit (%zu bytes of shared memory per CTA)", smem);
  const int n_items = b.n_bricks + b.n_tail;
  int grid = sms * per_sm;
  const int need = (n_items + WARPS - 1) / WARPS;
  if (grid > need) grid = need;
  kern<<<grid, WARPS * 32, smem, st>>>(k, b);
  VPB_LAUNCH_CHECK();
  return 0;
}

// VPB_BRICK_CFG (profiling and tests): bit 0 scalar instead of packed arithmetic, bit 1 movers deposit with global
// REDs, bits 2-3 tile geometry / warps per CTA (see the switch below), 0x100 use bricks however few particles they hold.
// Read on every call so that tests can switch it.
int brick_config() { const char *e = getenv("VPB_BRICK_CFG"); return e ? atoi(e) : 0; }

// Host side of the brick path; returns 1 when the call was served, 0 when the caller should use the linear kernel.
int advance_p_brick(const vpb_push_args_t *args, const PushK &k, cudaStream_t st) {
  if (!args->partition || args->p_first != 0) return 0;
  if (args->accum_stride != 12 && args->accum_stride != 16) return 0;
  const int nx = args->nx, ny = args->ny, nz = args->nz;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = kSMs; }
  const int cfg = brick_config();
  const int BX = 4, BY = 4, BZ = 4, Hmax = 2;
  // the tile index decode needs the tile to be narrower than a grid line / plane
  if (nx + 2 < BX + 2 * Hmax || (long long)(ny + 2) < 1 || nz < 1) return 0;
  BrickK b;
  memset(&b, 0, sizeof b);
  b.part = args->partition;
  b.nbx = (nx + BX - 1) / BX; b.nby = (ny + BY - 1) / BY; b.nbz = (nz + BZ - 1) / BZ;
  b.n_bricks = b.nbx * b.nby * b.nbz;
  // bricks pay for themselves only when they hold a few rows of particles
  if ((long long)args->np < 256ll * b.n_bricks && !(cfg & 0x100)) return 0;
  const int np_sorted = args->partition_np < args->np ? args->partition_np : args->np;
  b.tail0 = np_sorted < 0 ? 0 : np_sorted;
  b.n_tail = (args->np - b.tail0 + kTailRows * 32 - 1) / (kTailRows * 32);
  b.work = args->counters + 2;
  VPB_CUDA(cudaMemsetAsync(b.work, 0, sizeof(int), st));
  b.one2 = 0x3f8000003f800000ull; b.nz2 = 0x8000000080000000ull; b.mone2 = 0xbf800000bf800000ull;
  // packed arithmetic needs a numerator the fast division path can take without a range check
  const float aq = fabsf(args->qdt_2mc);
  const bool packed_ok = (aq == 0.0f) || (aq > 1e-15f && aq < 1e15f);
  const int as = args->accum_stride;
  switch (packed_ok ? (cfg & 0xff) : ((cfg & 0xff) | 1)) {
    // bit 0: scalar arithmetic; bit 1: movers deposit with global REDs; bits 2-3: geometry
    case 0:  return launch_brick<4, 4, 4, 1, 8, 2, ArithPacked, true>(k, b, as, sms, st) ? -1 : 1;
    case 1:  return launch_brick<4, 4, 4, 1, 8, 2, ArithScalar, true>(k, b, as, sms, st) ? -1 : 1;
    case 2:  return launch_brick<4, 4, 4, 1, 8, 2, ArithPacked, false>(k, b, as, sms, st) ? -1 : 1;
    case 3:  return launch_brick<4, 4, 4, 1, 8, 2, ArithScalar, false>(k, b, as, sms, st) ? -1 : 1;
    case 4:  return launch_brick<4, 4, 4, 2, 8, 1, ArithPacked, true>(k, b, as, sms, st) ? -1 : 1;
    case 5:  return launch_brick<4, 4, 4, 2, 8, 1, ArithScalar, true>(k, b, as, sms, st) ? -1 : 1;
    case 6:  return launch_brick<4, 4, 4, 2, 4, 2, ArithPacked, true>(k, b, as, sms, st) ? -1 : 1;
    case 7:  return launch_brick<4, 4, 4, 2, 4, 2, ArithScalar, true>(k, b, as, sms, st) ? -1 : 1;
    case 8:  return launch_brick<4, 4, 4, 1, 4, 4, ArithPacked, true>(k, b, as, sms, st) ? -1 : 1;
    case 9:  return launch_brick<4, 4, 4, 1, 4, 4, ArithScalar, true>(k, b, as, sms, st) ? -1 : 1;
    default: set_error("vpb_advance_p: unknown VPB_BRICK_CFG %d", cfg); return -1;
  }
}

}  // namespace vpb
