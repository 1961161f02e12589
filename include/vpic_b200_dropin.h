/* vpic_b200_dropin.h — the link-time seam: the reference's own extern "C" hot-path symbols, re-implemented over
 * the device layer (vpic_b200.h).  Signatures are the reference's (file:line below, under the reference tree); the
 * struct types are the layout-compatible restatements in vpic_b200_abi.h.  A host program built from the unmodified
 * reference sources gets the GPU path by linking libvpic_b200.so in place of the seven source files listed in
 * INTEGRATION.md (or, for a shared-library build of the reference, by LD_PRELOADing it).
 *
 * Ownership stays with the host: every array is allocated, resized and checkpointed by the reference host code.
 * This layer keeps a device mirror per host array, looked up by the CURRENT host pointer and size on every call.
 *
 * Coherence modes (vpic_b200_set_mode, or env VPIC_B200_MODE=auto|coherent|resident):
 *   VPB_MODE_AUTO (default)      arrays stay on the device between calls AND host code may still read or write any
 *                                array at any time: the pages of a device-owned array are access-protected, the
 *                                first host touch of a 2 MB chunk faults and the chunk is copied back before the
 *                                access is retried (vpic_b200/csrc/lazy_pages.h).  Arrays below VPIC_B200_LAZY_MIN
 *                                bytes (default 32 MB + 4 KB) are copied on every call as in coherent mode, so small
 *                                decks and tests behave exactly as VPB_MODE_COHERENT.
 *   VPB_MODE_COHERENT            every call copies its inputs host->device and its outputs device->host.
 *   VPB_MODE_RESIDENT            arrays stay on the device between calls; host copies go stale until
 *                                vpic_b200_sync_to_host(ptr) and host writes need vpic_b200_invalidate(ptr).
 *
 * Errors follow the reference's convention (src/util/util_base.h:267-273): message on stderr as
 * "Error at file(line)[rank]:", then exit(1).  A missing or failing CUDA device is such an error: the particle path
 * (advance_p, sort_p, the interpolator/accumulator glue, energy_p, center_p, ...) has NO CPU fallback.
 *
 * Forwards.  A few entry points exist here only because LD_PRELOAD interposes whole symbols: boundary_p,
 * synchronize_hydro_array and the reference's field kernels (advance_b, vacuum_advance_e, clear_jf, synchronize_jf,
 * vacuum_energy_f, the divergence-cleaning kernels).  When this library cannot serve one of them on the device —
 * several MPI ranks without the NCCL seam (vpic_b200_mp.h), more than one material, custom particle-boundary handlers,
 * VPIC_B200_FIELDS=0 / VPIC_B200_BOUNDARY_P=0 — the call is handed to the HOST PROGRAM's own definition of that symbol
 * (dlsym(RTLD_NEXT)), i.e. the reference's CPU code runs.  That is never silent: the first forward of each symbol
 * prints "Warning ...: <symbol> is not served on the device (<why>): forwarding to the host program's own CPU
 * implementation", VPIC_B200_TRACE=1 counts them, and VPIC_B200_STRICT=1 turns the first forward into an Error.
 */
#ifndef VPIC_B200_DROPIN_H
#define VPIC_B200_DROPIN_H

#include "vpic_b200_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VPB_MODE_COHERENT 0
#define VPB_MODE_RESIDENT 1
#define VPB_MODE_AUTO 2

/* src/species_advance/species_advance.h:73-76 */
void advance_p(vpb_species_t *sp, vpb_accumulator_array_t *aa, const vpb_interpolator_array_t *ia);
/* src/species_advance/species_advance.h:65-66 */
void sort_p(vpb_species_t *sp);
/* src/species_advance/species_advance.h:152-157 — single particles from host code (inject_particle, emitters): runs
 * on the device while the particle's memory is device-owned, otherwise it is the host program's own move_p (forward) */
int move_p(vpb_particle_t *p0, vpb_particle_mover_t *pm, vpb_accumulator_t *a0, const vpb_grid_t *g, const float qsp);
/* src/species_advance/species_advance.h:90-107 */
void center_p(vpb_species_t *sp, const vpb_interpolator_array_t *ia);
void uncenter_p(vpb_species_t *sp, const vpb_interpolator_array_t *ia);
double energy_p(const vpb_species_t *sp, const vpb_interpolator_array_t *ia);
/* src/species_advance/species_advance.h:117-119 */
void accumulate_rho_p(vpb_field_array_t *fa, const vpb_species_t *sp);
/* src/boundary/boundary.h:33-38.  Served on the device for a single rank without custom particle-boundary handlers
 * (movers are then particles that hit an absorbing wall); everything else is forwarded to the reference's own
 * boundary_p (dlsym RTLD_NEXT).  pbc_list is the reference's particle_bc_t list (opaque here). */
void boundary_p(void *pbc_list, vpb_species_t *sp_list, vpb_field_array_t *fa, vpb_accumulator_array_t *aa);
/* hydro moments — src/species_advance/species_advance.h:139-148, src/sf_interface/sf_interface.h:216-240 */
void accumulate_hydro_p(vpb_hydro_array_t *ha, const vpb_species_t *sp, const vpb_interpolator_array_t *ia);
void clear_hydro_array(vpb_hydro_array_t *ha);
void reduce_hydro_array(vpb_hydro_array_t *ha);
void synchronize_hydro_array(vpb_hydro_array_t *ha);
/* src/sf_interface/sf_interface.h:99-101 */
void load_interpolator_array(vpb_interpolator_array_t *ia, const vpb_field_array_t *fa);
/* src/sf_interface/sf_interface.h:147-148,158-159,172-174 */
void clear_accumulator_array(vpb_accumulator_array_t *aa);
void reduce_accumulator_array(vpb_accumulator_array_t *aa);
void unload_accumulator_array(vpb_field_array_t *fa, const vpb_accumulator_array_t *aa);

/* The reference's own field kernels (src/field_advance/standard/sfa_private.h:40,51-53,86-88,117-119,383): its
 * kernel table is filled from these C symbols (sfa.cc:27-64,202-211), so under LD_PRELOAD they take over without a
 * deck change.  Served on the device when the field array has one material and no face shared with another rank;
 * otherwise (or with VPIC_B200_FIELDS=0) the call falls through to the reference's definition (dlsym RTLD_NEXT). */
void advance_b(vpb_field_array_t *fa, float frac);
void vacuum_advance_e(vpb_field_array_t *fa, float frac);
void clear_jf(vpb_field_array_t *fa);
void synchronize_jf(vpb_field_array_t *fa);
void vacuum_energy_f(double *en6, const vpb_field_array_t *fa);
/* the divergence-cleaning and shared-face entries of the same table (sfa_private.h; advance.cc:138-176) */
void   clear_rhof(vpb_field_array_t *fa);
void   synchronize_rho(vpb_field_array_t *fa);
void   vacuum_compute_div_e_err(vpb_field_array_t *fa);
double compute_rms_div_e_err(const vpb_field_array_t *fa);
void   vacuum_clean_div_e(vpb_field_array_t *fa);
void   compute_div_b_err(vpb_field_array_t *fa);
double compute_rms_div_b_err(const vpb_field_array_t *fa);
void   clean_div_b(vpb_field_array_t *fa);
double synchronize_tang_e_norm_b(vpb_field_array_t *fa);
void   vacuum_compute_rhob(vpb_field_array_t *fa);
void   vacuum_compute_curl_b(vpb_field_array_t *fa);

/* Device-backed entries for the field_advance_kernels_t table (src/field_advance/field_advance.h:170-218);
 * vpic_b200_install_field_kernels(fa) repoints fa->kernel[0] at them (one material, single rank) — for host builds
 * where the symbols above cannot be interposed (static linking with the reference's field sources kept). */
void vpic_b200_advance_b(vpb_field_array_t *fa, float frac);
void vpic_b200_advance_e(vpb_field_array_t *fa, float frac);
void vpic_b200_clear_jf(vpb_field_array_t *fa);
void vpic_b200_synchronize_jf(vpb_field_array_t *fa);
void vpic_b200_energy_f(double *en6, const vpb_field_array_t *fa);
void   vpic_b200_clear_rhof(vpb_field_array_t *fa);
void   vpic_b200_synchronize_rho(vpb_field_array_t *fa);
void   vpic_b200_compute_div_e_err(vpb_field_array_t *fa);
double vpic_b200_compute_rms_div_e_err(const vpb_field_array_t *fa);
void   vpic_b200_clean_div_e(vpb_field_array_t *fa);
void   vpic_b200_compute_div_b_err(vpb_field_array_t *fa);
double vpic_b200_compute_rms_div_b_err(const vpb_field_array_t *fa);
void   vpic_b200_clean_div_b(vpb_field_array_t *fa);
double vpic_b200_synchronize_tang_e_norm_b(vpb_field_array_t *fa);
void   vpic_b200_compute_rhob(vpb_field_array_t *fa);
void   vpic_b200_compute_curl_b(vpb_field_array_t *fa);
void vpic_b200_install_field_kernels(vpb_field_array_t *fa);

/* coherence control (new; no reference counterpart) */
void vpic_b200_set_mode(int mode);
void vpic_b200_sync_to_host(const void *host_ptr);     /* device -> host for the mirror of host_ptr (NULL: all) */
void vpic_b200_invalidate(const void *host_ptr);       /* host copy was modified: re-upload on next use (NULL: all) */
void vpic_b200_release(const void *host_ptr);          /* drop the mirror (call before the host frees/reallocs) */
/* bytes moved by the drop-in layer since load: [0] host->device, [1] device->host */
void vpic_b200_transfer_bytes(uint64_t out[2]);
/* VPB_MODE_AUTO: make host memory [p, p+bytes) current and host-owned before handing it to code that does not fault
 * in user space (system calls, RDMA, MPI on registered memory).  fwrite/fread — what the reference's dumps and
 * checkpoints use, src/util/io/StandardIOPolicy.h:133-145 — are interposed by this library and need no call.
 * Returns the number of tracked arrays the range overlapped. */
int vpic_b200_host_access(const void *p, size_t bytes);
/* VPB_MODE_AUTO counters: [0] page faults served, [1] bytes copied back on host accesses, [2] arrays found remapped
 * by the host allocator, [3] arrays currently tracked */
void vpic_b200_lazy_stats(uint64_t out[4]);
/* VPB_MODE_AUTO: smallest array (bytes) that is tracked by page protection; same as env VPIC_B200_LAZY_MIN.  Applies
 * to arrays first seen after the call. */
void vpic_b200_set_lazy_min(size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
