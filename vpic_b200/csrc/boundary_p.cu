// boundary_p on the device: what happens to particles whose move_p ended on a face of the local domain.
//
// Replaces the particle-side work of src/boundary/boundary_p.cc:41-750 for slab-decomposed multi-GPU runs:
//   :257-371  walk the movers, turn particles bound for a neighbouring rank into particle_injector_t records
//             (coordinate on the crossed axis forced to -/+1, voxel made relative to the receiver's range),
//             drop absorbed ones, and back-fill the holes they leave in the particle array;
//   :595-711  append received injectors, give each a mover and finish its move_p.
// The message exchange itself (boundary_p.cc:392-446, MPI there) is NCCL send/recv in vpic_b200/parallel.py.
// Custom particle-boundary handlers (host callbacks, boundary_p.cc:332-346) are out of scope on the device.
//
// Determinism: movers arrive sorted by particle index (vpb_sort_movers); injectors are grouped per destination
// face by a STABLE split, so every buffer is in ascending particle order and the back-fill reproduces the
// reference's sequential "p[i] = p[--np]" result exactly (checked against the reference's own boundary_p in
// tests/test_gpu_parity.py::test_boundary_p_absorbing_walls_match_reference).
#include "push_common.cuh"

namespace vpb {

int radix_split_injectors(int4 *a, int4 *b, int n, void *scratch, size_t scratch_bytes, cudaStream_t st,
                          int *class_offsets_dev);
size_t radix_split_scratch_bytes(int n);

struct BoundK {
  float4 *p; int np;
  const int4 *pm; int nm;
  const long long *neighbor;
  long long rangel, rangeh, rangem;
  long long face_range[6];
  int sp_id;
  float *fields; float q_r8V; int nx, ny, nz;
  int absorb_all;
};

// class of a mover: 0..5 = send through that face, 6 = absorbed, 7 = no device handler (dropped with a count)
__global__ void __launch_bounds__(256) bp_build_injectors_kernel(BoundK k, int4 *inj /* 3 words per record */) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= k.nm) return;
  const int4 mv = k.pm[m];
  const int i = mv.w;
  float4 r = k.p[2 * (size_t)i];
  const float4 u = k.p[2 * (size_t)i + 1];
  int voxel = __float_as_int(r.w);
  const int face = voxel & 7;
  voxel >>= 3;
  const long long nn = __ldg(k.neighbor + 6ll * voxel + face);
  int cls = 7;
  int dst_voxel = voxel;
  if (nn == -2 /* absorb_particles, grid.h:30 */ || k.absorb_all) {
    cls = 6;
    if (k.fields) {                       // accumulate_rhob (rho_p.cc:126-213): the absorbed charge stays on the wall
      const int sy = k.nx + 2, sz = (k.nx + 2) * (k.ny + 2);
      float w0 = r.x, w1 = r.y, w2, w3, w4, w5, w6, w7 = k.q_r8V * u.w;
      const float dz = r.z;
      w6 = w7 - w0 * w7; w7 = w7 + w0 * w7;
      w4 = w6 - w1 * w6; w5 = w7 - w1 * w7;
      w6 = w6 + w1 * w6; w7 = w7 + w1 * w7;
      w0 = w4 - dz * w4; w1 = w5 - dz * w5; w2 = w6 - dz * w6; w3 = w7 - dz * w7;
      w4 = w4 + dz * w4; w5 = w5 + dz * w5; w6 = w6 + dz * w6; w7 = w7 + dz * w7;
      int x = voxel; const int z = x / sz;
      if (z == 1)    { w0 += w0; w1 += w1; w2 += w2; w3 += w3; }
      if (z == k.nz) { w4 += w4; w5 += w5; w6 += w6; w7 += w7; }
      x -= sz * z; const int y = x / sy;
      if (y == 1)    { w0 += w0; w1 += w1; w4 += w4; w5 += w5; }
      if (y == k.ny) { w2 += w2; w3 += w3; w6 += w6; w7 += w7; }
      x -= sy * y;
      if (x == 1)    { w0 += w0; w2 += w2; w4 += w4; w6 += w6; }
      if (x == k.nx) { w1 += w1; w3 += w3; w5 += w5; w7 += w7; }
      float *f = k.fields + 11;           // field_t.rhob
      red_add(f + 20 * (size_t)(voxel), w0);           red_add(f + 20 * (size_t)(voxel + 1), w1);
      red_add(f + 20 * (size_t)(voxel + sy), w2);      red_add(f + 20 * (size_t)(voxel + sy + 1), w3);
      red_add(f + 20 * (size_t)(voxel + sz), w4);      red_add(f + 20 * (size_t)(voxel + sz + 1), w5);
      red_add(f + 20 * (size_t)(voxel + sz + sy), w6); red_add(f + 20 * (size_t)(voxel + sz + sy + 1), w7);
    }
  }
  else if (((nn >= 0) && (nn < k.rangel)) || ((nn > k.rangeh) && (nn <= k.rangem))) {
    if (face < 6 && k.face_range[face] >= 0) {
      cls = face;
      dst_voxel = (int)(nn - k.face_range[face]);
      const float dir = face < 3 ? 1.0f : -1.0f;                  // where the sending face sits on the receiver
      const int axis = face % 3;
      if (axis == 0) r.x = dir; else if (axis == 1) r.y = dir; else r.z = dir;
    }
  }
  inj[3 * (size_t)m + 0] = make_int4(__float_as_int(r.x), __float_as_int(r.y), __float_as_int(r.z), dst_voxel);
  inj[3 * (size_t)m + 1] = make_int4(__float_as_int(u.x), __float_as_int(u.y), __float_as_int(u.z), __float_as_int(u.w));
  inj[3 * (size_t)m + 2] = make_int4(mv.x, mv.y, mv.z, cls);      // class rides in the sp_id slot until the split
}

__global__ void __launch_bounds__(256) bp_set_sp_id_kernel(int4 *inj, int n, int sp_id) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n) inj[3 * (size_t)m + 2].w = sp_id;
}

__device__ __forceinline__ int lower_bound_mover(const int4 *pm, int n, int key) {   // first m with pm[m].i >= key
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (pm[mid].w < key) lo = mid + 1; else hi = mid; }
  return lo;
}

// Back-fill: all nm mover particles leave the array.  Sequentially the reference does, for movers in DEscending
// index order, p[i] = p[--np] (boundary_p.cc:354-371).  Step j (0-based) therefore reads position L_j = np-1-j and
// writes the j-th largest mover index.  A position that is itself a mover index was overwritten at its own (earlier)
// step with the content of ITS L, so the element that finally lands in a hole is found by following that chain
// upwards until it leaves the mover set; chains are short (a tail position is a mover with probability nm/np).
// Holes at or above np' = np - nm are dead afterwards and need no write; sources are never written, so every
// thread can resolve and copy independently.
__global__ void __launch_bounds__(256) bp_backfill_kernel(BoundK k) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;          // ascending position in the mover list
  if (m >= k.nm) return;
  const int hole = k.pm[m].w;
  if (hole >= k.np - k.nm) return;
  int L = k.np - 1 - (k.nm - 1 - m);
  for (;;) {
    const int lb = lower_bound_mover(k.pm, k.nm, L);
    if (lb < k.nm && k.pm[lb].w == L) L = k.np - 1 - (k.nm - 1 - lb); else break;
  }
  k.p[2 * (size_t)hole] = k.p[2 * (size_t)L];
  k.p[2 * (size_t)hole + 1] = k.p[2 * (size_t)L + 1];
}

// Injection: record n_f - 1 - j of a face buffer lands at p[base + j'] in the reference's reverse walk, i.e. the
// LAST record is appended first (boundary_p.cc:640-644).  Each new particle gets a mover and finishes its move.
__global__ void __launch_bounds__(128) bp_inject_kernel(PushK a, const int4 *inj, int n, int base) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int4 w0 = inj[3 * (size_t)(n - 1 - j)], w1 = inj[3 * (size_t)(n - 1 - j) + 1], w2 = inj[3 * (size_t)(n - 1 - j) + 2];
  const int i = base + j;
  float4 r = make_float4(__int_as_float(w0.x), __int_as_float(w0.y), __int_as_float(w0.z), __int_as_float(w0.w));
  float4 u = make_float4(__int_as_float(w1.x), __int_as_float(w1.y), __int_as_float(w1.z), __int_as_float(w1.w));
  float dispx = __int_as_float(w2.x), dispy = __int_as_float(w2.y), dispz = __int_as_float(w2.z);
  const int left = move_p_dev(a, r, u, dispx, dispy, dispz);
  if (left) {
    const int slot = atomicAdd(a.counters, 1);
    if (slot < a.max_nm) a.pm[slot] = make_int4(__float_as_int(dispx), __float_as_int(dispy), __float_as_int(dispz), i);
    else { atomicAdd(a.counters + 1, 1); r.w = __int_as_float(__float_as_int(r.w) >> 3); }
  }
  a.p[2 * (size_t)i] = r;
  a.p[2 * (size_t)i + 1] = u;
}

// ---- fixed-capacity messages with the count in a 16-byte header (the reference reserves the same header in front of
// its injector buffers, boundary_p.cc:205-211) -------------------------------------------------------------------
// A message is int4 header {count, sp_id, capacity, 0} followed by `cap` particle_injector_t slots.  Sender and
// receiver agree on `cap` ahead of time, so the exchange needs no count handshake and no host synchronisation: the
// count travels inside the message and is read by the injection kernel on the device.
__global__ void __launch_bounds__(256) bp_stage_kernel(const int4 *inj, const int *class_offsets, int face, int cap, int sp_id,
                                                       int4 *msg, int *status) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int first = class_offsets[face], n = class_offsets[face + 1] - first;
  if (j == 0) {
    msg[0] = make_int4(n, sp_id, cap, 0);
    if (n > cap) atomicOr(status, 1);                              // capacity exceeded: reported, never silent
    atomicMax(status + 1, n);                                      // largest message of the step (sizes the next ones)
  }
  if (j >= n || j >= cap) return;
  const int4 *src = inj + 3 * (size_t)(first + j);
  int4 *dst = msg + 1 + 3 * (size_t)j;
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
}

// Injection of a received message: record count-1-j lands at p[base + *added + j] (the reference's reverse walk).
__global__ void __launch_bounds__(128) bp_inject_msg_kernel(PushK a, const int4 *msg, int cap, int base, int max_np,
                                                            const int *added, int *status) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(msg[0].x, cap);
  if (j >= n) return;
  const int i = base + *added + j;
  if (i >= max_np) { atomicOr(status, 2); return; }                // no room left in the particle array
  const int4 *rec = msg + 1 + 3 * (size_t)(n - 1 - j);
  const int4 w0 = rec[0], w1 = rec[1], w2 = rec[2];
  float4 r = make_float4(__int_as_float(w0.x), __int_as_float(w0.y), __int_as_float(w0.z), __int_as_float(w0.w));
  float4 u = make_float4(__int_as_float(w1.x), __int_as_float(w1.y), __int_as_float(w1.z), __int_as_float(w1.w));
  float dispx = __int_as_float(w2.x), dispy = __int_as_float(w2.y), dispz = __int_as_float(w2.z);
  const int left = move_p_dev(a, r, u, dispx, dispy, dispz);
  if (left) {
    const int slot = atomicAdd(a.counters, 1);
    if (slot < a.max_nm) a.pm[slot] = make_int4(__float_as_int(dispx), __float_as_int(dispy), __float_as_int(dispz), i);
    else { atomicAdd(a.counters + 1, 1); r.w = __int_as_float(__float_as_int(r.w) >> 3); }
  }
  a.p[2 * (size_t)i] = r;
  a.p[2 * (size_t)i + 1] = u;
}
__global__ void bp_bump_kernel(int *added, const int4 *msg, int cap) { *added += min(msg[0].x, cap); }

// move_p for ONE particle (src/species_advance/standard/move_p.cc:216-378): mover = {dispx, dispy, dispz, i} in and out,
// result[0] = the reference's return value (1: the particle is still in use, p.i = 8*voxel+face).
__global__ void move_p_single_kernel(PushK a, int4 *mover, int *result) {
  const int4 mv = mover[0];
  const int i = mv.w;
  float4 r = a.p[2 * (size_t)i], u = a.p[2 * (size_t)i + 1];
  float dispx = __int_as_float(mv.x), dispy = __int_as_float(mv.y), dispz = __int_as_float(mv.z);
  const int left = move_p_dev(a, r, u, dispx, dispy, dispz);
  a.p[2 * (size_t)i] = r;
  a.p[2 * (size_t)i + 1] = u;
  mover[0] = make_int4(__float_as_int(dispx), __float_as_int(dispy), __float_as_int(dispz), i);
  result[0] = left;
}

}  // namespace vpb

using namespace vpb;

extern "C" size_t vpb_boundary_msg_bytes(int32_t cap) { return 16 + (size_t)(cap < 0 ? 0 : cap) * 48; }

extern "C" int vpb_boundary_p_stage(const void *inj, const int32_t *class_offsets, int32_t face, int32_t cap, int32_t sp_id,
                                    void *msg, int32_t *status, void *stream) {
  VPB_REQUIRE(class_offsets && msg && status && face >= 0 && face < 6 && cap >= 0, "vpb_boundary_p_stage: Bad args.");
  bp_stage_kernel<<<(cap + 256) / 256, 256, 0, as_stream(stream)>>>((const int4 *)inj, class_offsets, face, cap, sp_id,
                                                                    (int4 *)msg, status);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_boundary_p_inject_msg(const vpb_push_args_t *push, const void *msg, int32_t cap, int32_t max_np,
                                         int32_t *added, int32_t *status, void *stream) {
  VPB_REQUIRE(push && push->p && push->accum && push->neighbor && push->counters && msg && added && status && cap >= 0,
              "vpb_boundary_p_inject_msg: Bad args.");
  if (cap == 0) return 0;
  const PushK k = to_push_k(push);
  cudaStream_t st = as_stream(stream);
  bp_inject_msg_kernel<<<(cap + 127) / 128, 128, 0, st>>>(k, (const int4 *)msg, cap, push->np, max_np, added, status);
  VPB_LAUNCH_CHECK();
  bp_bump_kernel<<<1, 1, 0, st>>>(added, (const int4 *)msg, cap);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_move_p(const vpb_push_args_t *push, void *mover_dev, int32_t *result_dev, void *stream) {
  VPB_REQUIRE(push && push->p && push->accum && push->neighbor && mover_dev && result_dev, "vpb_move_p: Bad args.");
  const PushK k = to_push_k(push);
  move_p_single_kernel<<<1, 1, 0, as_stream(stream)>>>(k, (int4 *)mover_dev, result_dev);
  VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t vpb_boundary_scratch_bytes(int32_t nm) {
  if (nm < 1) nm = 1;
  return (((size_t)nm * 48 + 255) / 256) * 256 + radix_split_scratch_bytes(nm);
}

extern "C" int vpb_boundary_p_pack(const vpb_boundary_args_t *b, void *stream) {
  VPB_REQUIRE(b && b->p && b->neighbor && b->class_offsets && (b->nm == 0 || (b->pm && b->inj && b->scratch)),
              "vpb_boundary_p_pack: Bad args.");
  VPB_REQUIRE(b->nm >= 0 && b->nm <= b->np, "vpb_boundary_p_pack: nm out of range");
  cudaStream_t st = as_stream(stream);
  if (b->nm == 0) { VPB_CUDA(cudaMemsetAsync(b->class_offsets, 0, 9 * sizeof(int), st)); return 0; }
  VPB_REQUIRE(b->scratch_bytes >= vpb_boundary_scratch_bytes(b->nm), "vpb_boundary_p_pack: scratch too small");
  BoundK k;
  k.p = (float4 *)b->p; k.np = b->np; k.pm = (const int4 *)b->pm; k.nm = b->nm;
  k.neighbor = (const long long *)b->neighbor; k.rangel = b->rangel; k.rangeh = b->rangeh; k.rangem = b->rangem;
  for (int f = 0; f < 6; f++) k.face_range[f] = b->face_range[f];
  k.sp_id = b->sp_id;
  k.fields = b->fields; k.q_r8V = b->q_r8V; k.nx = b->nx; k.ny = b->ny; k.nz = b->nz;
  k.absorb_all = b->absorb_all;
  VPB_REQUIRE(!b->fields || (b->nx > 0 && b->ny > 0 && b->nz > 0), "vpb_boundary_p_pack: grid size missing for rhob");
  int4 *tmp = (int4 *)b->scratch;
  void *sort_scratch = (char *)b->scratch + (((size_t)b->nm * 48 + 255) / 256) * 256;
  const int blocks = (b->nm + 255) / 256;
  bp_build_injectors_kernel<<<blocks, 256, 0, st>>>(k, tmp);                        VPB_LAUNCH_CHECK();
  int r = radix_split_injectors(tmp, (int4 *)b->inj, b->nm, sort_scratch, b->scratch_bytes - ((char *)sort_scratch - (char *)b->scratch),
                                st, b->class_offsets);
  if (r) return r;
  bp_set_sp_id_kernel<<<blocks, 256, 0, st>>>((int4 *)b->inj, b->nm, b->sp_id);     VPB_LAUNCH_CHECK();
  bp_backfill_kernel<<<blocks, 256, 0, st>>>(k);                                    VPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vpb_boundary_p_inject(const vpb_push_args_t *push, const void *inj, int32_t n, void *stream) {
  VPB_REQUIRE(push && push->p && push->accum && push->neighbor && push->counters && (inj || n == 0),
              "vpb_boundary_p_inject: Bad args.");
  if (n <= 0) return 0;
  const PushK k = to_push_k(push);
  bp_inject_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(k, (const int4 *)inj, n, push->np);
  VPB_LAUNCH_CHECK();
  return 0;
}
