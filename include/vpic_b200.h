/* vpic_b200.h — C-ABI of libvpic_b200.so (hand-written CUDA for sm_100a).
 *
 * Two layers, both plain C (pointers and sizes only, no torch or C++ types):
 *
 *  1. DEVICE LAYER  vpb_*   — every pointer is a DEVICE pointer; `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream).  This is what the
 *     drop-in layer, bench.py and the GPU tests call.  Each entry point names
 *     the reference routine it replaces (file:line under the reference tree).
 *
 *  2. DROP-IN LAYER (vpic_b200_dropin.h) — the reference's own extern "C"
 *     symbols (advance_p, sort_p, load_interpolator_array, ...) taking the
 *     reference's host structs; they mirror host arrays on the device and call
 *     layer 1.  That is the link-time seam described in INTEGRATION.md.
 *
 * Every function returns 0 on success, else a negative vpb error or a positive
 * cudaError_t; vpb_last_error() gives the message.  This layer has no CPU
 * fallback: without a usable CUDA device every call fails loudly.
 *
 * Array layouts are the reference's (vpic_b200_abi.h) with runtime strides so
 * one binary serves every SIMD padding of the host build:
 *   interp_stride  floats per interpolator_t  (20 | 24 | 32)
 *   accum_stride   floats per accumulator_t   (12 | 16)
 *   field_t is always 20 floats (80 B).
 */
#ifndef VPIC_B200_H
#define VPIC_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPB_VERSION 100

/* ---- runtime ------------------------------------------------------------ */
int         vpb_version(void);
const char *vpb_last_error(void);
int         vpb_device_count(int *count);
int         vpb_set_device(int device);
int         vpb_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_bytes);
int         vpb_malloc(void **dptr, size_t bytes);
int         vpb_free(void *dptr);
int         vpb_malloc_host(void **hptr, size_t bytes);        /* pinned host memory */
int         vpb_free_host(void *hptr);
int         vpb_memset(void *dptr, int value, size_t bytes, void *stream);
int         vpb_memcpy_h2d(void *dptr, const void *hptr, size_t bytes, void *stream);
int         vpb_memcpy_d2h(void *hptr, const void *dptr, size_t bytes, void *stream);
int         vpb_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int         vpb_stream_sync(void *stream);
int         vpb_device_sync(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t     vpb_launch_count(void);

/* ---- particle advance ----------------------------------------------------
 * Replaces advance_p_pipeline + advance_p_pipeline_scalar + move_p
 * (src/species_advance/standard/pipeline/advance_p_pipeline.cc:20-340,
 *  src/species_advance/standard/move_p.cc:216-378).
 * Arithmetic follows the reference's SCALAR pipeline operation for operation
 * (no FMA contraction, IEEE sqrt and divide) so particle state is bit-exact;
 * accumulator sums differ only by fp32 atomic ordering.
 */
/* Closed form of grid_t.neighbor for the regular grids the reference builds (size_grid / join_grid / set_pbc,
 * src/grid/ops.cc:19-211): inside the domain the neighbour is voxel +- stride; on each of the six walls it is one
 * action for the whole wall — a negative particle-boundary code, or "this voxel plus a constant" (periodic wrap
 * onto this rank, or the matching voxel of the rank behind the wall).  vpb_neighbor_rule_derive reads the six
 * actions out of the table and then CHECKS every interior voxel and face against the table on the device; only if
 * all 6*nx*ny*nz entries agree is valid set, and only then do the kernels skip the table lookup. */
typedef struct vpb_neighbor_rule {
  int32_t valid;
  int32_t nx, ny, nz;
  int64_t act[6];      /* < 0: boundary code of the wall; >= 0: the wall leads to another voxel */
  int64_t delta[6];    /* neighbour (global id) = rangel + voxel + delta[face] for walls with act >= 0 */
} vpb_neighbor_rule_t;
int vpb_neighbor_rule_derive(const int64_t *neighbor_dev, int32_t nx, int32_t ny, int32_t nz, int64_t rangel,
                             vpb_neighbor_rule_t *rule_out, void *stream);

typedef struct vpb_push_args {
  void          *p;              /* particle_t array, 32 B each, 32 B aligned         */
  int32_t        np;             /* particles [p_first, p_first + np) are advanced     */
  void          *pm;             /* particle_mover_t[max_nm]: movers that left the domain */
  int32_t        max_nm;
  int32_t       *counters;       /* int32[4] device: [0] += movers emitted (may exceed max_nm),
                                    [1] += movers dropped for lack of room (p.i restored, advance_p_pipeline.cc:223-236),
                                    [2] scratch of the brick kernel (work counter, reset by vpb_advance_p itself) */
  const float   *interp;  int32_t interp_stride;
  float         *accum;   int32_t accum_stride;   /* block 0 of the accumulator array */
  const int64_t *neighbor;       /* grid_t.neighbor, [6*nv]                          */
  int64_t        rangel, rangeh;
  float          qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;   /* computed by the caller in float, advance_p_pipeline.cc:279-283 */
  int32_t        nx, ny, nz;
  int32_t        variant;        /* deposit strategy, see VPB_DEPOSIT_*              */
  int32_t        p_first;        /* first particle of this call (chunked host pipelines); mover .i stay global */
  const vpb_neighbor_rule_t *neighbor_rule;   /* optional (host pointer): verified closed form of `neighbor` */
  int32_t        debug_skip;     /* must be 0.  Profiling only (results become INVALID): bit0 skip deposits,
                                    bit1 skip the mover phase, bit2 skip particle stores, bit3 skip the interpolator gather */
  /* Optional (NULL = not available): the partition[nv+1] the species' last vpb_sort_p wrote (device pointer), and the
   * particle count of that sort (partition[nv]).  With it vpb_advance_p walks the array brick by brick and collects
   * the deposits of each brick in a shared-memory accumulator tile (advance_p_brick.cu).  The array may have been
   * changed since the sort (back-fill, appended particles, np != partition_np): partition[] only has to be monotone
   * with entries in [0, partition_np]; results do not depend on how accurate it still is. */
  const int32_t *partition;
  int32_t        partition_np;
  /* Optional (NULL = off): apply the order of a vpb_sort_p_index on the fly.  The particle advanced into position k is
   * p[perm[k]] and it is written to p_out[k] (p_out != p, 32 B aligned, np particles; p_first must be 0).  The result in
   * p_out, the movers and the accumulators are exactly those of vpb_sort_p followed by vpb_advance_p on p. */
  const int32_t *perm;
  void          *p_out;
  /* Optional (NULL = off): int32[np], receives the voxel index every particle ends the push with (what sort_p would read
   * back out of p.i), in output positions.  A vpb_sort_p_index that follows with no change to the array in between
   * takes it as `keys` and never reads the particles.  Particles that left the domain store their 8*voxel+face code. */
  int32_t       *keys_out;
} vpb_push_args_t;

#define VPB_DEPOSIT_DEFAULT      0   /* library's best measured strategy                                  */
#define VPB_DEPOSIT_RED_V4       1   /* every particle: 3 x red.global.add.v4.f32                         */
#define VPB_DEPOSIT_WARP_SEG     2   /* in-voxel streaks: warp-level segmented reduction by voxel, one RED per sum;
                                        crossing streaks (move_p): 3 vector REDs each                     */
#define VPB_DEPOSIT_WARP_SEG_MOVERS 3 /* as 2, and move_p runs warp-synchronously with the same segmented reduction */
#define VPB_DEPOSIT_WARP_SEG_FIRST  4 /* as 2, and the FIRST streak of a mover batch (still in the source voxels) is
                                        summed across the warp; later streaks go out as vector REDs          */
#define VPB_DEPOSIT_BRICK_TILE      5 /* bricks of voxels, warp-private shared-memory accumulator tile flushed by TMA
                                        bulk reduce, two particles per thread; needs `partition`.  DEFAULT picks this
                                        when `partition` is given and falls back to 2 otherwise                   */

int vpb_advance_p(const vpb_push_args_t *args, void *stream);

/* Sort the emitted movers ascending by particle index (boundary_p needs that order,
 * src/boundary/boundary_p.cc:248-255).  nm = min(counters[0], max_nm), read back by the caller.
 * scratch: >= vpb_sort_movers_scratch_bytes(nm). */
size_t vpb_sort_scratch_bytes(int32_t n_items, int32_t n_keys_hint);
size_t vpb_sort_movers_scratch_bytes(int32_t nm);
int    vpb_sort_movers(void *pm, int32_t nm, void *scratch, size_t scratch_bytes, void *stream);

/* ---- boundary_p (particle side) ---------------------------------------------
 * Replaces the mover walk, back-fill and injection of src/boundary/boundary_p.cc:257-371,595-711 for runs that are
 * decomposed over several GPUs; the message exchange between them is NCCL send/recv in the host layer.
 * vpb_boundary_p_pack: every one of the nm movers' particles leaves the array (np becomes np - nm): particles bound
 * for a neighbour become particle_injector_t records in `inj`, grouped by destination class in ascending particle
 * order; class_offsets[c]..[c+1] delimits class c: 0..5 = face -x,-y,-z,+x,+y,+z, 6 = absorbed, 7 = no device handler.
 * vpb_boundary_p_inject: appends n received injectors at p[push->np ...) (last record first, like the reference),
 * gives each a mover and finishes its move_p; movers still in use are appended through push->counters. */
typedef struct vpb_boundary_args {
  void          *p;   int32_t np;
  const void    *pm;  int32_t nm;        /* movers, ascending by particle index */
  const int64_t *neighbor;
  int64_t        rangel, rangeh, rangem; /* rangem = range[world_size] */
  int64_t        face_range[6];          /* range[rank behind face] for faces shared with another rank, else -1 */
  int32_t        sp_id;
  void          *inj;                    /* out: particle_injector_t[nm] */
  int32_t       *class_offsets;          /* out: int32[9], device */
  void          *scratch; size_t scratch_bytes;   /* >= vpb_boundary_scratch_bytes(nm) */
  float         *fields;                 /* optional: field_t array; absorbed particles leave their charge in rhob
                                            (accumulate_rhob, rho_p.cc:126-213) */
  float          q_r8V;                  /* species charge * grid r8V, for rhob */
  int32_t        nx, ny, nz;
  int32_t        absorb_all;             /* 1: every mover is treated as absorbed — what vpic_simulation::advance does with
                                            movers that are still unresolved after the last communication round
                                            (src/vpic/advance.cc:78-101: charge into rhob, particle removed) */
} vpb_boundary_args_t;
size_t vpb_boundary_scratch_bytes(int32_t nm);
int    vpb_boundary_p_pack(const vpb_boundary_args_t *args, void *stream);
int    vpb_boundary_p_inject(const vpb_push_args_t *push, const void *inj, int32_t n, void *stream);

/* Fixed-capacity migration messages (no count handshake, no host synchronisation inside a round).  A message is a
 * 16-byte header {int32 count, sp_id, capacity, 0} — the header the reference reserves in front of its own injector
 * buffers, src/boundary/boundary_p.cc:205-211 — followed by `cap` particle_injector_t slots; vpb_boundary_msg_bytes(cap)
 * bytes in all.  Both neighbours use the same cap, so the send and the receive can be posted without knowing the count.
 *   vpb_boundary_p_stage       copies class `face` of a vpb_boundary_p_pack result into msg and writes the header.
 *   vpb_boundary_p_inject_msg  appends the records of a received message at p[push->np + *added ...) in the reference's
 *                              order, finishes their moves, then adds the count to *added (device int).
 * status (device int32[2]): [0] |= 1 when a count exceeded cap (records beyond cap were NOT sent), |= 2 when the particle
 * array was full; [1] = largest count staged since it was last cleared.  The caller reads it once per step. */
/* move_p for one particle (src/species_advance/species_advance.h:152-157, move_p.cc:216-378): *mover_dev is a
 * particle_mover_t on the device (in: displacement and particle index, out: what is left of it), *result_dev receives
 * the reference's return value.  push supplies p, accum, neighbor, rangel/rangeh, qsp and the grid size. */
int    vpb_move_p(const vpb_push_args_t *push, void *mover_dev, int32_t *result_dev, void *stream);

size_t vpb_boundary_msg_bytes(int32_t cap);
int    vpb_boundary_p_stage(const void *inj, const int32_t *class_offsets, int32_t face, int32_t cap, int32_t sp_id,
                            void *msg, int32_t *status, void *stream);
int    vpb_boundary_p_inject_msg(const vpb_push_args_t *push, const void *msg, int32_t cap, int32_t max_np,
                                 int32_t *added, int32_t *status, void *stream);

/* ---- sort_p ---------------------------------------------------------------
 * Replaces sort_p_pipeline (src/species_advance/standard/pipeline/sort_p_pipeline.cc:220-371):
 * stable counting sort of particles by voxel index p.i; writes partition[0..nv] (partition[nv] = np).
 * aux: particle_t[np] scratch; scratch: >= vpb_sort_scratch_bytes(np, nv) bytes. */
int vpb_sort_p(void *p, int32_t np, void *aux, int32_t *partition,
               int32_t nx, int32_t ny, int32_t nz,
               void *scratch, size_t scratch_bytes, void *stream);
/* The same order and partition[] without moving the particles: perm[k] = index in p of the particle that sort_p would
 * put at position k (a stable LSD sort of 8-byte (voxel, index) pairs instead of 32-byte particles).  The particles
 * move once, inside the next vpb_advance_p (push args perm / p_out), or in vpb_permute_p when no push follows.
 * keys (optional): int32[np], the voxel of every particle, when the caller already holds it; else p is read.
 * work: >= vpb_sort_index_work_bytes(np) bytes (the species' aux particle array is large enough);
 * scratch: >= vpb_sort_index_scratch_bytes(np, nv). */
size_t vpb_sort_index_work_bytes(int32_t np);
size_t vpb_sort_index_scratch_bytes(int32_t n_items, int32_t n_keys_hint);
int vpb_sort_p_index(const void *p, const int32_t *keys, int32_t np, int32_t *perm, int32_t *partition,
                     int32_t nx, int32_t ny, int32_t nz, void *work, size_t work_bytes,
                     void *scratch, size_t scratch_bytes, void *stream, void *partition_ready_event);
/* partition_ready_event: optional cudaEvent_t, recorded on `stream` as soon as partition[] is final (before the scatter
 * passes), so that a caller can copy it elsewhere on another stream while the order is still being computed. */
/* dst[k] = p[perm[k]] for k < np (dst != p) */
int vpb_permute_p(const void *p, int32_t np, const int32_t *perm, void *dst, void *stream);
/* keys[k] = voxel index of p[k] for k < n (refreshes part of a keys_out array after the host edited those particles) */
int vpb_extract_keys(const void *p, int32_t n, int32_t *keys, void *stream);
/* the inverse for a few particles: p[perm[k]] = src[k] for k < n (perm may point into the middle of an order) */
int vpb_unpermute_p(void *p, int32_t n, const int32_t *perm, const void *src, void *stream);

/* ---- interpolator / accumulator glue ---------------------------------------
 * load_interpolator_pipeline_scalar  (src/sf_interface/pipeline/interpolator_array_pipeline.cc:21-135)
 * clear_accumulator_array_pipeline   (src/sf_interface/pipeline/clear_array_pipeline.cc:40-67)
 * unload_accumulator_pipeline_scalar (src/sf_interface/pipeline/unload_accumulator_pipeline.cc:18-144)
 * reduce_accumulator_array is the identity here: the device keeps ONE accumulator block. */
int vpb_load_interpolator(float *interp, int32_t interp_stride, const float *fields,
                          int32_t nx, int32_t ny, int32_t nz, void *stream);
int vpb_clear_accumulator(float *accum, int32_t accum_stride, int32_t nx, int32_t ny, int32_t nz, void *stream);
int vpb_unload_accumulator(float *fields, const float *accum, int32_t accum_stride,
                           int32_t nx, int32_t ny, int32_t nz,
                           float rdx, float rdy, float rdz, float dt, void *stream);

/* ---- particle diagnostics / centering --------------------------------------
 * energy_p_pipeline (src/species_advance/standard/pipeline/energy_p_pipeline.cc:18-115): result (double, already
 * times cvac^2) is written to *en_dev (device).  center_p / uncenter_p (center_p_pipeline.cc:17-96,
 * uncenter_p_pipeline.cc:17-98). */
int vpb_energy_p(const void *p, int32_t np, const float *interp, int32_t interp_stride,
                 float q, float m, float dt, float cvac, double *en_dev, void *stream);
int vpb_center_p(void *p, int32_t np, const float *interp, int32_t interp_stride, float qdt_2mc, void *stream);
int vpb_uncenter_p(void *p, int32_t np, const float *interp, int32_t interp_stride, float qdt_2mc, void *stream);

/* accumulate_rho_p (src/species_advance/standard/rho_p.cc:22-113): node charge density of a species into
 * field_t.rhof (used by the divergence cleaning and by rho diagnostics). */
int vpb_accumulate_rho_p(float *fields, const void *p, int32_t np, float q, float r8V,
                         int32_t nx, int32_t ny, int32_t nz, void *stream);

/* ---- standard field advance, vacuum material -------------------------------
 * advance_b (src/field_advance/standard/pipeline/advance_b_pipeline.cc:20-125),
 * vacuum_advance_e (.../vacuum_advance_e_pipeline.cc:20-332), clear_jf (sfa.cc:231-237),
 * synchronize_jf (remote.cc:417-508), vacuum_energy_f (.../vacuum_energy_f_pipeline.cc:12-97),
 * with the local boundary conditions of local.cc and the periodic/remote ghost handling of remote.cc.
 * face[f], f = -x,-y,-z,+x,+y,+z:  VPB_FACE_PERIODIC_SELF  ghost plane copied from the opposite side of this domain,
 *                                  VPB_FACE_REMOTE         ghost plane filled by the caller (NCCL halo exchange),
 *                                  <0                      local field BC code (grid.h:20-26): -1 pec, -2 symmetric, -3 pmc,
 *                                                          -4 absorbing (Higdon, local.cc:84-112). */
#define VPB_FACE_PERIODIC_SELF 0
#define VPB_FACE_REMOTE        1

typedef struct vpb_field_args {
  float  *f;                       /* field_t[nv], 20 floats each */
  int32_t nx, ny, nz;
  float   dt, cvac, eps0, damp;
  float   dx, dy, dz, dV;
  float   rdx, rdy, rdz;
  int32_t face[6];
  /* The single material that fills space (the reference's vacuum_* kernels are used whenever the material list has
   * one entry, whatever its coefficients — sfa.cc:202-211): material_coefficient_t order (sfa_private.h:14-25)
   * decayx drivex decayy drivey decayz drivez rmux rmuy rmuz nonconductive epsx epsy epsz.  has_material == 0 means
   * true vacuum (all ones). */
  int32_t has_material;
  float   material[13];
} vpb_field_args_t;

int vpb_advance_b(const vpb_field_args_t *a, float frac, void *stream);
int vpb_vacuum_advance_e(const vpb_field_args_t *a, float frac, void *stream);
int vpb_clear_jf(const vpb_field_args_t *a, void *stream);
int vpb_synchronize_jf(const vpb_field_args_t *a, void *stream);
int vpb_vacuum_energy_f(const vpb_field_args_t *a, double *en6_dev, void *stream);

/* Divergence cleaning and shared-face synchronisation (src/vpic/advance.cc:138-176) on the same field array:
 * clear_rhof (sfa.cc:239-256), synchronize_rho (remote.cc:534-620), vacuum_compute_div_e_err / vacuum_clean_div_e
 * (pipeline/vacuum_{compute_div_e_err,clean_div_e}_pipeline.*), compute_div_b_err / clean_div_b
 * (pipeline/{compute_div_b_err,clean_div_b}_pipeline.cc), synchronize_tang_e_norm_b (remote.cc:298-416).
 * The two rms functions and synchronize_tang_e_norm_b leave a partial result in device memory: *sum_dev is the
 * weighted sum of squared errors over this domain (the caller finishes eps0*sqrt(sum*dV / (nx*ny*nz*dV)) after its
 * cross-rank sum, compute_rms_div_e_err_pipeline.cc:170-183), *err_dev the local desynchronisation error.
 * Faces must be local walls or periodic-self (no VPB_FACE_REMOTE yet). */
int vpb_clear_rhof(const vpb_field_args_t *a, void *stream);
int vpb_synchronize_rho(const vpb_field_args_t *a, void *stream);
int vpb_vacuum_compute_div_e_err(const vpb_field_args_t *a, void *stream);
int vpb_compute_rms_div_e_err(const vpb_field_args_t *a, double *sum_dev, void *stream);
int vpb_vacuum_clean_div_e(const vpb_field_args_t *a, void *stream);
int vpb_compute_div_b_err(const vpb_field_args_t *a, void *stream);
int vpb_compute_rms_div_b_err(const vpb_field_args_t *a, double *sum_dev, void *stream);
int vpb_clean_div_b(const vpb_field_args_t *a, void *stream);
int vpb_synchronize_tang_e_norm_b(const vpb_field_args_t *a, double *err_dev, void *stream);
/* initialisation-time entries (src/vpic/initialize.cc): vacuum_compute_rhob, vacuum_compute_curl_b */
int vpb_vacuum_compute_rhob(const vpb_field_args_t *a, void *stream);
int vpb_vacuum_compute_curl_b(const vpb_field_args_t *a, void *stream);

/* Hydro moments (diagnostics behind the reference's hydro dumps): accumulate_hydro_p
 * (src/species_advance/standard/pipeline/hydro_p_pipeline.cc:19-252) adds the 14 moments of every particle to the 8
 * nodes of its voxel; synchronize_hydro_array (src/sf_interface/hydro_array.cc:131-309) doubles wall nodes and folds
 * periodic planes.  hydro = hydro_t[nv], 16 floats each (jx jy jz rho px py pz ke txx tyy tzz tyz tzx txy pad pad). */
int vpb_accumulate_hydro_p(float *hydro, const void *p, int32_t np, const float *interp, int32_t interp_stride,
                           float q, float m, float dt, float cvac, float r8V,
                           int32_t nx, int32_t ny, int32_t nz, void *stream);
int vpb_clear_hydro(float *hydro, int32_t nx, int32_t ny, int32_t nz, void *stream);
int vpb_synchronize_hydro(float *hydro, const vpb_field_args_t *geometry, void *stream);   /* f of geometry is ignored */
/* node planes of faces shared with another rank (hydro_array.cc:203-262): pack the plane the neighbour behind `face`
 * shares with us, unpack adds the plane received through `face` (own + remote); vpb_hydro_halo_floats per plane */
size_t vpb_hydro_halo_floats(int32_t nx, int32_t ny, int32_t nz, int axis);
int vpb_hydro_halo_pack(float *hydro, const vpb_field_args_t *geometry, int face, float *buf, void *stream);
int vpb_hydro_halo_unpack(float *hydro, const vpb_field_args_t *geometry, int face, const float *buf, void *stream);

/* Halo planes for VPB_FACE_REMOTE faces (the payload of begin/end_remote_ghost_tang_b, remote.cc:61-134, and of
 * synchronize_jf, remote.cc:417-508).  pack copies the plane a neighbour needs into buf; unpack applies a received
 * plane.  floats per plane: vpb_halo_floats(). */
#define VPB_HALO_TANG_B 0
#define VPB_HALO_JF     1
/* the periodic extras of divergence cleaning and shared-face synchronisation:
 *   RHO             rhof, rhob of the shared node plane (synchronize_rho, remote.cc:547-585): own + remote / mean
 *   NORM_E          normal E of the first interior node plane into the neighbour's ghost plane (remote.cc:136-206);
 *                   exchange BEFORE vpb_vacuum_compute_div_e_err / vpb_vacuum_compute_rhob
 *   DIV_B           div-B error of the first interior cell plane into the neighbour's ghost cells (remote.cc:208-282);
 *                   exchange BEFORE vpb_clean_div_b
 *   TANG_E_NORM_B   normal B, tangential E and TCA of the shared plane, averaged with the neighbour's
 *                   (remote.cc:298-416); vpb_halo_unpack_sync adds the squared differences to *err_dev */
#define VPB_HALO_RHO           2
#define VPB_HALO_NORM_E        3
#define VPB_HALO_DIV_B         4
#define VPB_HALO_TANG_E_NORM_B 5
size_t vpb_halo_floats(int32_t nx, int32_t ny, int32_t nz, int axis);                 /* TANG_B and JF planes */
size_t vpb_halo_floats_kind(int32_t nx, int32_t ny, int32_t nz, int axis, int kind);  /* any kind */
int vpb_halo_pack(const vpb_field_args_t *a, int kind, int face, float *buf, void *stream);
int vpb_halo_unpack(const vpb_field_args_t *a, int kind, int face, const float *buf, void *stream);
int vpb_halo_unpack_sync(const vpb_field_args_t *a, int face, const float *buf, double *err_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VPIC_B200_H */
