// sort_p for sm_100a: stable counting sort of particles by voxel index, and of movers by particle index.
//
// Replaces src/species_advance/standard/pipeline/sort_p_pipeline.cc:30-371 (coarse_count / coarse_sort / subsort
// over pthread pipelines).  The reference's two stable passes give exactly "stable sort by p.i"; so does this:
// a least-significant-digit radix sort over only the bits nv needs (11-bit digits when that saves a pass, else
// 8-bit), each pass a stable split
//   (1) per-CTA digit histogram over the CTA's contiguous chunk,
//   (2) exclusive scan of the [digit][CTA] matrix (digit-major),
//   (3) stable scatter: inside a CTA, items are ranked warp by warp with __match_any_sync against per-warp digit
//       counters in shared memory, so equal keys keep their input order.
// vpb_sort_p moves whole 32-byte particles (one DRAM sector each, 256-bit accesses) in every pass.  vpb_sort_p_index
// (further down) sorts 8-byte (voxel, index) pairs with the same passes and leaves the particle movement to the
// advance_p that follows (vpb_push_args_t.perm) — the default of both host layers.  partition[] is the exclusive scan of
// the per-voxel counts the first histogram gathers on its way.
#include "vpb_common.cuh"

namespace vpb {

constexpr int kSortBlock = 256;
constexpr int kSortWarps = kSortBlock / 32;
constexpr int kIPT = 4;                                 // items per thread per sub-tile
constexpr int kSubTile = kSortBlock * kIPT;
constexpr int kMaxSortBlocks = 4096;

// ---- exclusive scan of int32 (in place), n up to 8192 * 8192 ----------------------------------------------
constexpr int kScanBlock = 1024;
constexpr int kScanIPT = 8;
constexpr int kScanTile = kScanBlock * kScanIPT;

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = lane < nw ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    if (lane < nw) s_warp[lane] = s;
  }
  __syncthreads();
  const int woff = w ? s_warp[w - 1] : 0;
  total = s_warp[nw - 1];
  __syncthreads();
  return woff + x - v;
}

__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(const int *in, int n, int *block_sums) {
  __shared__ int s_warp[32];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanIPT;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanIPT; k++) if (base + k < n) s += in[base + k];
  int total;
  block_exclusive_scan(s, s_warp, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of up to kScanTile block sums
__global__ void __launch_bounds__(kScanBlock) scan_top_kernel(int *sums, int n) {
  __shared__ int s_warp[32];
  const int base = threadIdx.x * kScanIPT;
  int v[kScanIPT], s = 0;
#pragma unroll
  for (int k = 0; k < kScanIPT; k++) { v[k] = (base + k < n) ? sums[base + k] : 0; s += v[k]; }
  int total;
  int off = block_exclusive_scan(s, s_warp, total);
#pragma unroll
  for (int k = 0; k < kScanIPT; k++) { if (base + k < n) sums[base + k] = off; off += v[k]; }
}

__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(int *data, int n, const int *block_offs) {
  __shared__ int s_warp[32];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanIPT;
  int v[kScanIPT], s = 0;
#pragma unroll
  for (int k = 0; k < kScanIPT; k++) { v[k] = (base + k < n) ? data[base + k] : 0; s += v[k]; }
  int total;
  int off = block_exclusive_scan(s, s_warp, total) + block_offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanIPT; k++) { if (base + k < n) data[base + k] = off; off += v[k]; }
}

static int exclusive_scan_inplace(int *data, int n, int *tmp /* >= ceil(n/kScanTile) ints */, cudaStream_t st) {
  const int nb = (n + kScanTile - 1) / kScanTile;
  VPB_REQUIRE(nb <= kScanTile, "scan: %d elements exceed the two-level scan", n);
  scan_reduce_kernel<<<nb, kScanBlock, 0, st>>>(data, n, tmp);  VPB_LAUNCH_CHECK();
  scan_top_kernel<<<1, kScanBlock, 0, st>>>(tmp, nb);           VPB_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, kScanBlock, 0, st>>>(data, n, tmp);   VPB_LAUNCH_CHECK();
  return 0;
}

// ---- radix passes ----------------------------------------------------------------------------------------
// Items are VEC int4 words; the key is word KEYW's .w (particle_t.i and particle_mover_t.i sit in word 0,
// the destination class of a particle_injector_t rides in word 2).

// key_count (optional, first pass of sort_p only): key_count[k] += number of items with full key k, 0 <= k < n_keys.
// Its exclusive scan is partition[] (sort_p_pipeline.cc:183-193,335-340), so the sorted array never has to be read
// again to build it.  Equal keys imply equal digits, so one match on the full key serves both histograms.
template <int VEC, int KEYW, int BITS>
__global__ void __launch_bounds__(kSortBlock) radix_hist_kernel(const int4 *items, int n, int per_block, int shift,
                                                                int *hist /* [1<<BITS][gridDim.x] */,
                                                                int *key_count = nullptr, int n_keys = 0) {
  constexpr int R = 1 << BITS;
  __shared__ int s_hist[R];
  for (int d = threadIdx.x; d < R; d += kSortBlock) s_hist[d] = 0;
  __syncthreads();
  const int lo = blockIdx.x * per_block;
  const int hi = min(n, lo + per_block);
  const int lane = threadIdx.x & 31;
  for (int i0 = lo; i0 < hi; i0 += kSortBlock) {
    const int i = i0 + threadIdx.x;
    const bool valid = i < hi;
    const int key = valid ? items[(size_t)i * VEC + KEYW].w : 0;
    const int d = valid ? ((key >> shift) & (R - 1)) : (R + lane);
    if (key_count) {
      const unsigned peers = __match_any_sync(0xffffffffu, valid ? key : (-1 - lane));
      if (valid && lane == __ffs(peers) - 1) {
        atomicAdd(&s_hist[d], __popc(peers));
        if ((unsigned)key < (unsigned)n_keys) atomicAdd(&key_count[key], __popc(peers));
      }
    } else {
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_hist[d], __popc(peers));
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < R; d += kSortBlock) hist[d * gridDim.x + blockIdx.x] = s_hist[d];
}

// Stable scatter of one digit.  Shared memory: next output slot per digit (s_base), per-warp digit counters of the
// current sub-tile (s_wcount) and the range of digits the sub-tile touched.  Only that range is prefixed and re-zeroed,
// so an 11-bit digit (2048 bins x 8 warps) costs no more per sub-tile than an 8-bit one when the input is nearly
// sorted (a sub-tile of voxel-ordered particles touches a few dozen consecutive digits).
// next_hist (optional): the [digit][CTA] histogram of the NEXT pass, filled from the output positions this pass
// assigns (an item written to position o belongs to CTA o / per_block of the next pass), so the next pass needs no
// read of its own to count.  Must be zero on entry.
template <int VEC, int KEYW, int BITS>
__global__ void __launch_bounds__(kSortBlock) radix_scatter_kernel(const int4 *src, int4 *dst, int n, int per_block,
                                                                   int shift, const int *offs /* scanned hist */,
                                                                   int *next_hist = nullptr) {
  constexpr int R = 1 << BITS;
  extern __shared__ int s_dyn[];
  int *s_base = s_dyn;                                  // [R]
  int *s_wcount = s_dyn + R;                            // [kSortWarps][R]
  int *s_range = s_dyn + R + kSortWarps * R;            // [2]: lowest and highest digit the current sub-tile touched
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int d = tid; d < R; d += kSortBlock) s_base[d] = offs[d * gridDim.x + blockIdx.x];
  for (int d = tid; d < R * kSortWarps; d += kSortBlock) s_wcount[d] = 0;
  if (tid == 0) { s_range[0] = R; s_range[1] = -1; }
  const int lo = blockIdx.x * per_block;
  const int hi = min(n, lo + per_block);
  __syncthreads();

  for (int t0 = lo; t0 < hi; t0 += kSubTile) {
    int4 it[kIPT][VEC];
    int digit[kIPT], rank[kIPT];
    int wlo = R, whi = -1;                                 // digits this warp touches in this sub-tile
    // warp-striped: warp w owns items [t0 + w*32*kIPT, +32*kIPT); step j covers 32 consecutive items
#pragma unroll
    for (int j = 0; j < kIPT; j++) {
      const int i = t0 + w * 32 * kIPT + j * 32 + lane;
      const bool valid = i < hi;
      if (valid) {
        if (VEC == 2) {                                  // a particle is one 32-byte sector: one 256-bit access
          float4 lo, hi;
          ld_particle(reinterpret_cast<const float4 *>(src) + 2 * (size_t)i, lo, hi);
          it[j][0] = make_int4(__float_as_int(lo.x), __float_as_int(lo.y), __float_as_int(lo.z), __float_as_int(lo.w));
          it[j][VEC - 1] = make_int4(__float_as_int(hi.x), __float_as_int(hi.y), __float_as_int(hi.z), __float_as_int(hi.w));
        } else {
#pragma unroll
          for (int v = 0; v < VEC; v++) it[j][v] = src[(size_t)i * VEC + v];
        }
      }
      const int d = valid ? ((it[j][KEYW].w >> shift) & (R - 1)) : (R + lane);
      digit[j] = valid ? d : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      int before = 0;
      if (valid) before = s_wcount[w * R + d];
      __syncwarp();
      if (valid && lane == __ffs(peers) - 1) s_wcount[w * R + d] = before + __popc(peers);
      __syncwarp();
      rank[j] = before + __popc(peers & lt_mask);
      wlo = min(wlo, __reduce_min_sync(0xffffffffu, valid ? d : R));
      whi = max(whi, __reduce_max_sync(0xffffffffu, valid ? d : -1));
    }
    if (lane == 0 && whi >= 0) { atomicMin(&s_range[0], wlo); atomicMax(&s_range[1], whi); }
    __syncthreads();
    // exclusive prefix over warps for every digit in the touched range, one thread per digit: a sub-tile of nearly
    // sorted items touches a few dozen CONSECUTIVE digits (a bitmap walk left them all to one or two threads)
    const int dlo = s_range[0], dhi = s_range[1];
    for (int d = dlo + tid; d <= dhi; d += kSortBlock) {
      int run = s_base[d];
#pragma unroll
      for (int ww = 0; ww < kSortWarps; ww++) { const int c = s_wcount[ww * R + d]; s_wcount[ww * R + d] = run; run += c; }
      s_base[d] = run;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kIPT; j++) {
      const int o = digit[j] >= 0 ? s_wcount[w * R + digit[j]] + rank[j] : 0;
      if (next_hist) {
        // (next digit, next CTA) pairs repeat along a run of voxel-ordered items: one atomic per distinct pair per warp
        const int nd = (it[j][KEYW].w >> (shift + BITS)) & (R - 1);
        const int cell = digit[j] >= 0 ? nd * (int)gridDim.x + o / per_block : -1 - lane;
        const unsigned pp = __match_any_sync(0xffffffffu, cell);
        if (digit[j] >= 0 && lane == __ffs(pp) - 1) atomicAdd(&next_hist[cell], __popc(pp));
      }
      if (digit[j] >= 0) {
        if (VEC == 2) {
          const int4 a0 = it[j][0], a1 = it[j][VEC - 1];
          st_particle(reinterpret_cast<float4 *>(dst) + 2 * (size_t)o,
                      make_float4(__int_as_float(a0.x), __int_as_float(a0.y), __int_as_float(a0.z), __int_as_float(a0.w)),
                      make_float4(__int_as_float(a1.x), __int_as_float(a1.y), __int_as_float(a1.z), __int_as_float(a1.w)));
        } else {
#pragma unroll
          for (int v = 0; v < VEC; v++) dst[(size_t)o * VEC + v] = it[j][v];
        }
      }
    }
    __syncthreads();
    // re-zero the counters this sub-tile used
    for (int d = dlo + tid; d <= dhi; d += kSortBlock) {
#pragma unroll
      for (int ww = 0; ww < kSortWarps; ww++) s_wcount[ww * R + d] = 0;
    }
    if (tid == 0) { s_range[0] = R; s_range[1] = -1; }
    __syncthreads();
  }
}

// partition[v] = number of particles with p.i < v, for v in [0, nv]  (sort_p_pipeline.cc:183-193,335-340)
__global__ void partition_kernel(const int4 *p, int np, int nv, int *partition) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > np) return;
  int lo = (i == 0) ? 0 : p[2 * (size_t)(i - 1)].w + 1;
  int hi = (i == np) ? nv : p[2 * (size_t)i].w;
  if (lo < 0) lo = 0;
  if (hi > nv) hi = nv;
  for (int v = lo; v <= hi; v++) partition[v] = i;
}

static int ceil_log2(int64_t x) { int b = 0; while (((int64_t)1 << b) < x) b++; return b; }

struct SortPlan { int nblocks, per_block, scan_tmp; size_t hist_bytes, total_bytes; };

static SortPlan plan_sort(int n, int bits = 11, int n_keys = 0, int sub_tile = kSubTile, bool pow2 = false) {
  const int R = 1 << bits;
  SortPlan s;
  int nb = (n + sub_tile - 1) / sub_tile;
  if (nb < 1) nb = 1;
  if (nb > kMaxSortBlocks) nb = kMaxSortBlocks;
  int per = (n + nb - 1) / nb;
  per = ((per + sub_tile - 1) / sub_tile) * sub_tile;          // whole sub-tiles per CTA
  if (per < sub_tile) per = sub_tile;
  if (pow2) { int q = sub_tile; while (q < per) q <<= 1; per = q; }   // output position -> CTA of the next pass by a shift
  nb = (n + per - 1) / per; if (nb < 1) nb = 1;
  s.nblocks = nb; s.per_block = per;
  s.scan_tmp = (R * nb + kScanTile - 1) / kScanTile;
  const int key_tmp = (n_keys + 1 + kScanTile - 1) / kScanTile;           // the partition scan shares the temporary
  if (key_tmp > s.scan_tmp) s.scan_tmp = key_tmp;
  s.hist_bytes = (size_t)R * nb * sizeof(int);
  // two histograms: the pass that runs and the next one, which the running scatter fills
  s.total_bytes = 2 * (((s.hist_bytes + 255) / 256) * 256) + (size_t)s.scan_tmp * sizeof(int) + 256;
  return s;
}

// LSD radix sort of n items on the low key_bits of the key.  WIDE: 11-bit digits (two passes cover the 22 bits of a
// 128^3 grid), else 8-bit digits (small arrays: movers, injectors).  The result ends in a or b (*result_in_b).
// Only the first pass reads its input to count; every scatter fills the histogram of the pass after it.
// key_count / n_keys: see radix_hist_kernel (zeroed by the caller).
template <int VEC, int KEYW, bool WIDE>
static int radix_sort(int4 *a, int4 *b, int n, int key_bits, void *scratch, size_t scratch_bytes, cudaStream_t st,
                      bool *result_in_b, int *key_count = nullptr, int n_keys = 0) {
  constexpr int BITS = WIDE ? 11 : 8;
  constexpr int R = 1 << BITS;
  constexpr size_t smem = ((size_t)R * (1 + kSortWarps) + R / 32) * sizeof(int);
  const SortPlan pl = plan_sort(n, BITS, n_keys);
  VPB_REQUIRE(scratch && scratch_bytes >= pl.total_bytes, "sort: scratch too small (%zu < %zu)", scratch_bytes, pl.total_bytes);
  static bool attr_done = false;
  if (!attr_done) {
    VPB_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<VEC, KEYW, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const size_t hb = ((pl.hist_bytes + 255) / 256) * 256;
  int *hist[2] = {(int *)scratch, (int *)((char *)scratch + hb)};
  int *tmp = (int *)((char *)scratch + 2 * hb);
  int4 *src = a, *dst = b;
  int cur = 0;
  radix_hist_kernel<VEC, KEYW, BITS><<<pl.nblocks, kSortBlock, 0, st>>>(src, n, pl.per_block, 0, hist[cur], key_count, n_keys);
  VPB_LAUNCH_CHECK();
  for (int shift = 0; shift < key_bits; shift += BITS) {
    const bool more = shift + BITS < key_bits;
    if (more) VPB_CUDA(cudaMemsetAsync(hist[cur ^ 1], 0, pl.hist_bytes, st));
    int r = exclusive_scan_inplace(hist[cur], R * pl.nblocks, tmp, st); if (r) return r;
    radix_scatter_kernel<VEC, KEYW, BITS><<<pl.nblocks, kSortBlock, smem, st>>>(src, dst, n, pl.per_block, shift, hist[cur],
                                                                               more ? hist[cur ^ 1] : nullptr);
    VPB_LAUNCH_CHECK();
    int4 *t = src; src = dst; dst = t;
    cur ^= 1;
  }
  *result_in_b = (src == b);
  return 0;
}

// ---- index sort: the order of sort_p without moving the particles -----------------------------------------------
// vpb_sort_p_index sorts (voxel, index) pairs — 8 bytes per particle and pass instead of 64 — and leaves perm[k] = the
// index of the particle that belongs at position k.  The particles themselves move once, inside the advance_p that
// follows (it reads p[perm[k]] and writes position k of the other buffer), or in vpb_permute_p when none follows.
//   key_hist_kernel      voxel of every particle (read from the particles, or from a key array a previous advance_p
//                        left) -> compact key array, digit-0 histogram per CTA, per-voxel counts for partition[]
//   pair_scatter_kernel  one stable LSD pass; SRC 0: keys (index implicit) / 1: pairs, DST 0: pairs / 1: index only
// Items per thread and sub-tile: the first pass reads bare keys (the index is the position), so it can hold twice as
// many per thread in the same registers — and its sub-tiles touch digits all over the range (the y and z neighbours of
// a voxel differ in the low bits), so every sub-tile pays a walk over the whole [warp][digit] counter matrix.
#ifndef VPB_PAIR_IPT0
#define VPB_PAIR_IPT0 16
#endif
template <int SRC> struct PairTile { static constexpr int kIPT = SRC == 0 ? VPB_PAIR_IPT0 : 8; static constexpr int kItems = kSortBlock * kIPT; };
constexpr int kPairSubTile = PairTile<0>::kItems;       // per-CTA chunks are whole multiples of the larger sub-tile

template <int BITS>
__global__ void __launch_bounds__(kSortBlock) key_hist_kernel(const int4 *p, const int *keys_in, int *keys_out, int n,
                                                              int per_block, int *hist, int *key_count, int n_keys) {
  constexpr int R = 1 << BITS;
  __shared__ int s_hist[R];
  for (int d = threadIdx.x; d < R; d += kSortBlock) s_hist[d] = 0;
  __syncthreads();
  const int lo = blockIdx.x * per_block;
  const int hi = min(n, lo + per_block);
  const int lane = threadIdx.x & 31;
  for (int i0 = lo; i0 < hi; i0 += kSortBlock) {
    const int i = i0 + threadIdx.x;
    const bool valid = i < hi;
    int key = 0;
    if (valid) {
      if (keys_in) key = keys_in[i];
      else { key = p[2 * (size_t)i].w; keys_out[i] = key; }
    }
    const int d = key & (R - 1);
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? key : (-1 - lane));
    if (valid && lane == __ffs(peers) - 1) {
      atomicAdd(&s_hist[d], __popc(peers));
      if ((unsigned)key < (unsigned)n_keys) atomicAdd(&key_count[key], __popc(peers));
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < R; d += kSortBlock) hist[d * gridDim.x + blockIdx.x] = s_hist[d];
}

template <int BITS, int SRC, int DST>
__global__ void __launch_bounds__(kSortBlock, (SRC == 0 && VPB_PAIR_IPT0 > 8) ? 2 : 3) pair_scatter_kernel(const void *src_, void *dst_, int n, int per_block,
                                                                  int shift, const int *offs, int *next_hist) {
  constexpr int R = 1 << BITS;
  constexpr int IPT = PairTile<SRC>::kIPT;
  extern __shared__ int s_dyn[];
  int *s_base = s_dyn;                                  // [R]
  int *s_wcount = s_dyn + R;                            // [kSortWarps][R]
  int *s_range = s_dyn + R + kSortWarps * R;            // [2]: lowest and highest digit the current sub-tile touched
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int d = tid; d < R; d += kSortBlock) s_base[d] = offs[d * gridDim.x + blockIdx.x];
  for (int d = tid; d < R * kSortWarps; d += kSortBlock) s_wcount[d] = 0;
  if (tid == 0) { s_range[0] = R; s_range[1] = -1; }
  const int lo = blockIdx.x * per_block;
  const int hi = min(n, lo + per_block);
  const int pb_shift = 31 - __clz(per_block);
  __syncthreads();

  for (int t0 = lo; t0 < hi; t0 += PairTile<SRC>::kItems) {
    int key[IPT], idx[SRC == 0 ? 1 : IPT], rank[IPT];
    // warp-striped as in radix_scatter_kernel: warp w owns [t0 + w*32*IPT, +32*IPT), step j covers 32 consecutive items
    const int i0 = t0 + w * 32 * IPT + lane;
#pragma unroll
    for (int j = 0; j < IPT; j++) {
      const int i = i0 + j * 32;
      if (SRC == 0) { key[j] = i < hi ? static_cast<const int *>(src_)[i] : 0; }
      else if (i < hi) { const int2 v = static_cast<const int2 *>(src_)[i]; key[j] = v.x; idx[j] = v.y; }
      else { key[j] = 0; idx[j] = -1; }
    }
    int wlo = R, whi = -1;                                 // digits this warp touches in this sub-tile
#pragma unroll
    for (int j = 0; j < IPT; j++) {
      const bool valid = i0 + j * 32 < hi;
      const int d = valid ? ((key[j] >> shift) & (R - 1)) : (R + lane);
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      int before = 0;
      if (valid) before = s_wcount[w * R + d];
      __syncwarp();
      if (valid && lane == __ffs(peers) - 1) s_wcount[w * R + d] = before + __popc(peers);
      __syncwarp();
      rank[j] = before + __popc(peers & lt_mask);
      wlo = min(wlo, __reduce_min_sync(0xffffffffu, valid ? d : R));
      whi = max(whi, __reduce_max_sync(0xffffffffu, valid ? d : -1));
    }
    if (lane == 0 && whi >= 0) { atomicMin(&s_range[0], wlo); atomicMax(&s_range[1], whi); }
    __syncthreads();
    // exclusive prefix over warps for every digit in the touched range, one thread per digit
    const int dlo = s_range[0], dhi = s_range[1];
    for (int d = dlo + tid; d <= dhi; d += kSortBlock) {
      int run = s_base[d];
#pragma unroll
      for (int ww = 0; ww < kSortWarps; ww++) { const int c = s_wcount[ww * R + d]; s_wcount[ww * R + d] = run; run += c; }
      s_base[d] = run;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; j++) {
      const bool valid = i0 + j * 32 < hi;
      const int o = valid ? s_wcount[w * R + ((key[j] >> shift) & (R - 1))] + rank[j] : 0;
      if (next_hist) {
        const int nd = (key[j] >> (shift + BITS)) & (R - 1);
        const int cell = valid ? nd * (int)gridDim.x + (o >> pb_shift) : -1 - lane;     // per_block is a power of two here
        const unsigned pp = __match_any_sync(0xffffffffu, cell);
        if (valid && lane == __ffs(pp) - 1) atomicAdd(&next_hist[cell], __popc(pp));
      }
      if (valid) {
        const int id = SRC == 0 ? i0 + j * 32 : idx[SRC == 0 ? 0 : j];
        if (DST == 0) static_cast<int2 *>(dst_)[o] = make_int2(key[j], id);
        else static_cast<int *>(dst_)[o] = id;
      }
    }
    __syncthreads();
    // re-zero the counters this sub-tile used
    for (int d = dlo + tid; d <= dhi; d += kSortBlock) {
#pragma unroll
      for (int ww = 0; ww < kSortWarps; ww++) s_wcount[ww * R + d] = 0;
    }
    if (tid == 0) { s_range[0] = R; s_range[1] = -1; }
    __syncthreads();
  }
}

template <int BITS, int SRC, int DST>
static int launch_pair_scatter(const SortPlan &pl, const void *src, void *dst, int n, int shift, const int *offs,
                               int *next_hist, cudaStream_t st) {
  constexpr int R = 1 << BITS;
  constexpr size_t smem = ((size_t)R * (1 + kSortWarps) + R / 32) * sizeof(int);
  static bool attr_done = false;
  if (!attr_done) {
    VPB_CUDA(cudaFuncSetAttribute(pair_scatter_kernel<BITS, SRC, DST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  pair_scatter_kernel<BITS, SRC, DST><<<pl.nblocks, kSortBlock, smem, st>>>(src, dst, n, pl.per_block, shift, offs, next_hist);
  VPB_LAUNCH_CHECK();
  return 0;
}

static size_t index_sort_work_bytes(int n) { return (((size_t)n * 4 + 255) / 256) * 256 + 2 * ((((size_t)n * 8 + 255) / 256) * 256); }

template <int BITS>
static int index_sort(const int4 *p, const int *keys_in, int n, int key_bits, int *perm, char *work, void *scratch,
                      size_t scratch_bytes, cudaStream_t st, int *key_count, int n_keys, cudaEvent_t partition_ready) {
  constexpr int R = 1 << BITS;
  const SortPlan pl = plan_sort(n, BITS, n_keys, kPairSubTile, true);
  VPB_REQUIRE(scratch && scratch_bytes >= pl.total_bytes, "sort: scratch too small (%zu < %zu)", scratch_bytes, pl.total_bytes);
  const size_t hb = ((pl.hist_bytes + 255) / 256) * 256;
  int *hist[2] = {(int *)scratch, (int *)((char *)scratch + hb)};
  int *tmp = (int *)((char *)scratch + 2 * hb);
  int *keys = (int *)work;
  const size_t kb_bytes = (((size_t)n * 4 + 255) / 256) * 256, pb_bytes = (((size_t)n * 8 + 255) / 256) * 256;
  void *pairs[2] = {work + kb_bytes, work + kb_bytes + pb_bytes};
  key_hist_kernel<BITS><<<pl.nblocks, kSortBlock, 0, st>>>(p, keys_in, keys, n, pl.per_block, hist[0], key_count, n_keys);
  VPB_LAUNCH_CHECK();
  // partition[] = exclusive scan of the per-voxel counts: final here, before the scatter passes, so that a caller who has
  // to hand it to the host can copy it while they run (partition_ready)
  { int r = exclusive_scan_inplace(key_count, n_keys + 1, tmp, st); if (r) return r; }
  if (partition_ready) VPB_CUDA(cudaEventRecord(partition_ready, st));
  const void *src = keys_in ? (const void *)keys_in : (const void *)keys;
  int cur = 0, pc = 0;
  for (int shift = 0; shift < key_bits; shift += BITS) {
    const bool first = shift == 0, last = shift + BITS >= key_bits;
    if (!last) VPB_CUDA(cudaMemsetAsync(hist[cur ^ 1], 0, pl.hist_bytes, st));
    int r = exclusive_scan_inplace(hist[cur], R * pl.nblocks, tmp, st); if (r) return r;
    void *dst = last ? (void *)perm : pairs[pc];
    int *nh = last ? nullptr : hist[cur ^ 1];
    if (first && last)       r = launch_pair_scatter<BITS, 0, 1>(pl, src, dst, n, shift, hist[cur], nh, st);
    else if (first)          r = launch_pair_scatter<BITS, 0, 0>(pl, src, dst, n, shift, hist[cur], nh, st);
    else if (last)           r = launch_pair_scatter<BITS, 1, 1>(pl, src, dst, n, shift, hist[cur], nh, st);
    else                     r = launch_pair_scatter<BITS, 1, 0>(pl, src, dst, n, shift, hist[cur], nh, st);
    if (r) return r;
    src = dst; pc ^= 1; cur ^= 1;
  }
  return 0;
}

// p[perm[k]] = src[k]: puts particles given in sorted positions back where the unsorted array holds them
__global__ void __launch_bounds__(256) unpermute_p_kernel(float4 *p, const int *perm, const float4 *src, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float4 r, u;
  ld_particle(src + 2 * (size_t)k, r, u);
  st_particle(p + 2 * (size_t)__ldg(perm + k), r, u);
}

// keys[k] = p[k].i for a few particles (the ends of an array whose other keys a push left behind)
__global__ void __launch_bounds__(256) extract_keys_kernel(const int4 *p, int *keys, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) keys[k] = p[2 * (size_t)k].w;
}

// dst[k] = p[perm[k]]: one 32-byte sector gathered per particle, consecutive stores
__global__ void __launch_bounds__(256) permute_p_kernel(const float4 *p, const int *perm, float4 *dst, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float4 r, u;
  ld_particle(p + 2 * (size_t)__ldg(perm + k), r, u);
  st_particle(dst + 2 * (size_t)k, r, u);
}

// class start offsets of a one-pass split: scanned hist[c * nblocks] is where class c begins
__global__ void split_offsets_kernel(const int *hist, int nblocks, int n, int *offs9) {
  const int c = threadIdx.x;
  if (c < 8) offs9[c] = hist[c * nblocks];
  if (c == 8) offs9[8] = n;
}

size_t radix_split_scratch_bytes(int n) { return plan_sort(n > 0 ? n : 1, 8).total_bytes; }

// Stable split of particle_injector_t records (3 words) by the class in word 2 .w (0..7): one radix pass a -> b.
int radix_split_injectors(int4 *a, int4 *b, int n, void *scratch, size_t scratch_bytes, cudaStream_t st,
                          int *class_offsets_dev) {
  bool in_b = false;
  int r = radix_sort<3, 2, false>(a, b, n, 3, scratch, scratch_bytes, st, &in_b);
  if (r) return r;
  const SortPlan pl = plan_sort(n, 8);
  split_offsets_kernel<<<1, 32, 0, st>>>((const int *)scratch, pl.nblocks, n, class_offsets_dev);
  VPB_LAUNCH_CHECK();
  return 0;
}

}  // namespace vpb

using namespace vpb;

extern "C" size_t vpb_sort_scratch_bytes(int32_t n_items, int32_t n_keys_hint) {
  // the wider plan covers both digit widths; n_keys_hint = nv sizes the partition scan's temporary
  return plan_sort(n_items > 0 ? n_items : 1, 11, n_keys_hint > 0 ? n_keys_hint : 0).total_bytes;
}

extern "C" int vpb_sort_p(void *p, int32_t np, void *aux, int32_t *partition, int32_t nx, int32_t ny, int32_t nz,
                          void *scratch, size_t scratch_bytes, void *stream) {
  VPB_REQUIRE(p && partition && (aux || np == 0), "vpb_sort_p: Bad args");
  VPB_REQUIRE(np >= 0 && nx > 0 && ny > 0 && nz > 0, "vpb_sort_p: Bad args");
  cudaStream_t st = as_stream(stream);
  const int64_t nv64 = (int64_t)(nx + 2) * (ny + 2) * (nz + 2);
  VPB_REQUIRE(nv64 < (1ll << 31), "vpb_sort_p: too many voxels");
  const int nv = (int)nv64;
  if (np > 1) {
    bool in_aux = false;
    // 11-bit digits when they save a pass (e.g. 22 key bits: 2 passes instead of 3), else 8-bit digits
    const int kb = ceil_log2(nv);
    const bool wide = (kb + 10) / 11 < (kb + 7) / 8;
    // partition[] = exclusive scan of the per-voxel counts, which the first histogram pass gathers on its way
    VPB_CUDA(cudaMemsetAsync(partition, 0, ((size_t)nv + 1) * sizeof(int), st));
    int r = wide ? radix_sort<2, 0, true>((int4 *)p, (int4 *)aux, np, kb, scratch, scratch_bytes, st, &in_aux, partition, nv)
                 : radix_sort<2, 0, false>((int4 *)p, (int4 *)aux, np, kb, scratch, scratch_bytes, st, &in_aux, partition, nv);
    if (r) return r;
    if (in_aux) VPB_CUDA(cudaMemcpyAsync(p, aux, (size_t)np * 32, cudaMemcpyDeviceToDevice, st));
    const SortPlan pl = plan_sort(np, wide ? 11 : 8, nv);
    int *tmp = (int *)((char *)scratch + 2 * (((pl.hist_bytes + 255) / 256) * 256));
    r = exclusive_scan_inplace(partition, nv + 1, tmp, st); if (r) return r;
    return 0;
  }
  const int threads = 256;
  partition_kernel<<<(np + 1 + threads - 1) / threads, threads, 0, st>>>((const int4 *)p, np, nv, partition);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t vpb_sort_index_work_bytes(int32_t np) { return index_sort_work_bytes(np > 0 ? np : 1); }

extern "C" size_t vpb_sort_index_scratch_bytes(int32_t n_items, int32_t n_keys_hint) {
  return plan_sort(n_items > 0 ? n_items : 1, 11, n_keys_hint > 0 ? n_keys_hint : 0, kPairSubTile, true).total_bytes;
}

extern "C" int vpb_sort_p_index(const void *p, const int32_t *keys, int32_t np, int32_t *perm, int32_t *partition,
                                int32_t nx, int32_t ny, int32_t nz, void *work, size_t work_bytes,
                                void *scratch, size_t scratch_bytes, void *stream, void *partition_ready_event) {
  VPB_REQUIRE((p || keys) && perm && partition && np >= 0 && nx > 0 && ny > 0 && nz > 0, "vpb_sort_p_index: Bad args");
  cudaStream_t st = as_stream(stream);
  cudaEvent_t ev = reinterpret_cast<cudaEvent_t>(partition_ready_event);
  const int64_t nv64 = (int64_t)(nx + 2) * (ny + 2) * (nz + 2);
  VPB_REQUIRE(nv64 < (1ll << 31), "vpb_sort_p_index: too many voxels");
  const int nv = (int)nv64;
  if (np == 0) {
    VPB_CUDA(cudaMemsetAsync(partition, 0, ((size_t)nv + 1) * sizeof(int), st));
    if (ev) VPB_CUDA(cudaEventRecord(ev, st));
    return 0;
  }
  VPB_REQUIRE(work && work_bytes >= index_sort_work_bytes(np), "vpb_sort_p_index: work area too small (%zu < %zu)",
              work_bytes, index_sort_work_bytes(np));
  const int kb = ceil_log2(nv) > 0 ? ceil_log2(nv) : 1;
  const bool wide = (kb + 10) / 11 < (kb + 7) / 8;
  VPB_CUDA(cudaMemsetAsync(partition, 0, ((size_t)nv + 1) * sizeof(int), st));
  return wide ? index_sort<11>((const int4 *)p, keys, np, kb, perm, (char *)work, scratch, scratch_bytes, st, partition, nv, ev)
              : index_sort<8>((const int4 *)p, keys, np, kb, perm, (char *)work, scratch, scratch_bytes, st, partition, nv, ev);
}

extern "C" int vpb_permute_p(const void *p, int32_t np, const int32_t *perm, void *dst, void *stream) {
  VPB_REQUIRE(np >= 0 && (np == 0 || (p && perm && dst && p != dst)), "vpb_permute_p: Bad args");
  if (np == 0) return 0;
  permute_p_kernel<<<(np + 255) / 256, 256, 0, as_stream(stream)>>>((const float4 *)p, perm, (float4 *)dst, np);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_extract_keys(const void *p, int32_t n, int32_t *keys, void *stream) {
  VPB_REQUIRE(n >= 0 && (n == 0 || (p && keys)), "vpb_extract_keys: Bad args");
  if (n == 0) return 0;
  extract_keys_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>((const int4 *)p, keys, n);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_unpermute_p(void *p, int32_t n, const int32_t *perm, const void *src, void *stream) {
  VPB_REQUIRE(n >= 0 && (n == 0 || (p && perm && src)), "vpb_unpermute_p: Bad args");
  if (n == 0) return 0;
  unpermute_p_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>((float4 *)p, perm, (const float4 *)src, n);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t vpb_sort_movers_scratch_bytes(int32_t nm) {
  if (nm < 1) nm = 1;
  return (((size_t)nm * 16 + 255) / 256) * 256 + plan_sort(nm, 8).total_bytes;
}

// Movers: sort ascending by particle index (unique keys).  nm is a host value here; the drop-in layer reads the
// device counter first (it has to report sp->nm to the host anyway).
extern "C" int vpb_sort_movers(void *pm, int32_t nm, void *scratch, size_t scratch_bytes, void *stream) {
  if (nm <= 1) return 0;
  VPB_REQUIRE(pm && scratch, "vpb_sort_movers: Bad args");
  cudaStream_t st = as_stream(stream);
  // scratch = [aux movers | sort scratch]
  const size_t aux_bytes = (((size_t)nm * 16 + 255) / 256) * 256;
  VPB_REQUIRE(scratch_bytes > aux_bytes, "vpb_sort_movers: scratch too small");
  bool in_aux = false;
  int r = radix_sort<1, 0, false>((int4 *)pm, (int4 *)scratch, nm, 31, (char *)scratch + aux_bytes, scratch_bytes - aux_bytes, st, &in_aux);
  if (r) return r;
  if (in_aux) VPB_CUDA(cudaMemcpyAsync(pm, scratch, (size_t)nm * 16, cudaMemcpyDeviceToDevice, st));
  return 0;
}
