/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference hot path.
 *
 * Plain C, single-threaded, written from the reference's scalar pipelines
 * (the bit-reproducible numerics reference, SURVEY.md §7 hard part 3).  Each
 * function cites the reference file:line it follows.  Compiled with
 * -ffp-contract=off so no FMA is formed, like the reference's scalar build.
 *
 * PARITY PINNED: tests/test_oracle_vs_ref.py checks every function here
 * bit-for-bit against the UNMODIFIED reference compiled into
 * oracle/_ref/libvpic_ref_scalar.so (recipe: oracle/Makefile) on seeded
 * inputs (push, move_p, sort, glue, field advance with every wall kind and a
 * dielectric material, divergence cleaning, hydro moments); the reference
 * build itself is pinned by its own known-answer decks and golden energy
 * tests, which pass when built with oracle/Makefile (tests/test_dropin_decks.py
 * runs them on the CPU path before the GPU path).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  The product (vpic_b200/) never does.
 *
 * All arrays use the reference layouts with runtime strides so every SIMD
 * padding variant is covered: interpolator stride 20/24/32 floats,
 * accumulator stride 12/16 floats (sf_interface.h:27-53).
 */
#ifndef VPIC_ORACLE_H
#define VPIC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vpo_particle { float dx, dy, dz; int32_t i; float ux, uy, uz, w; } vpo_particle_t;
typedef struct vpo_mover    { float dispx, dispy, dispz; int32_t i; } vpo_mover_t;

typedef struct vpo_push_args {
  vpo_particle_t *p;   int32_t np;
  vpo_mover_t    *pm;  int32_t max_nm;
  const float    *interp; int32_t interp_stride;   /* floats per voxel */
  float          *accum;  int32_t accum_stride;    /* floats per voxel */
  const int64_t  *neighbor;                        /* [6*nv] */
  int64_t         rangel, rangeh;
  float           qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;
} vpo_push_args_t;

/* advance_p_pipeline_scalar, advance_p_pipeline.cc:20-245.  Returns nm (movers kept); *n_ignored = lost movers. */
int32_t vpo_advance_p(const vpo_push_args_t *a, int32_t *n_ignored);

/* move_p scalar variant, move_p.cc:216-378.  Returns 1 if the mover is still in use. */
int vpo_move_p(vpo_particle_t *p0, vpo_mover_t *pm, float *accum, int32_t accum_stride,
               const int64_t *neighbor, int64_t rangel, int64_t rangeh, float qsp);

/* sort_p_pipeline, sort_p_pipeline.cc:220-371: stable counting sort by p.i; partition[0..nv) filled, incl. ghosts. */
void vpo_sort_p(vpo_particle_t *p, int32_t np, vpo_particle_t *aux, int32_t *partition,
                int32_t nx, int32_t ny, int32_t nz);

/* load_interpolator_pipeline_scalar, interpolator_array_pipeline.cc:21-135.  fields: 20 floats (80 B) per voxel. */
void vpo_load_interpolator(float *interp, int32_t interp_stride, const float *fields,
                           int32_t nx, int32_t ny, int32_t nz);

/* clear_accumulator_array_pipeline, clear_array_pipeline.cc:40-67 (block 0 only). */
void vpo_clear_accumulator(float *accum, int32_t accum_stride, int32_t nx, int32_t ny, int32_t nz);

/* unload_accumulator_pipeline_scalar, unload_accumulator_pipeline.cc:18-82,137-139. */
void vpo_unload_accumulator(float *fields, const float *accum, int32_t accum_stride,
                            int32_t nx, int32_t ny, int32_t nz,
                            float rdx, float rdy, float rdz, float dt);

/* energy_p_pipeline_scalar + driver, energy_p_pipeline.cc:18-115 (without the cross-rank sum). */
double vpo_energy_p(const vpo_particle_t *p, int32_t np, const float *interp, int32_t interp_stride,
                    float q, float m, float dt, float cvac);

/* center_p / uncenter_p scalar pipelines, center_p_pipeline.cc:17-96, uncenter_p_pipeline.cc:17-98. */
void vpo_center_p(vpo_particle_t *p, int32_t np, const float *interp, int32_t interp_stride, float qdt_2mc);
void vpo_uncenter_p(vpo_particle_t *p, int32_t np, const float *interp, int32_t interp_stride, float qdt_2mc);

/* accumulate_rho_p (rho_p.cc:22-113): trilinear charge of every particle into field_t.rhof; accumulate_rhob
 * (rho_p.cc:126-213): one particle into field_t.rhob with doubled weights on domain walls. */
void vpo_accumulate_rho_p(float *fields, const vpo_particle_t *p, int32_t np, float q, float r8V,
                          int32_t nx, int32_t ny, int32_t nz);
void vpo_accumulate_rhob(float *fields, const vpo_particle_t *p, float qsp, float r8V,
                         int32_t nx, int32_t ny, int32_t nz);

/* ---- standard field advance, vacuum material, one local domain whose six faces are either
 *      periodic onto itself or a local field BC (pec = -1).  bc6[f] for faces -x,-y,-z,+x,+y,+z:
 *      >=0 periodic-self, -1 pec.  (sfa: advance_b_pipeline.cc:20-125, vacuum_advance_e_pipeline.cc:20-332,
 *      local.cc:50-444, remote.cc:61-134,417-508) ------------------------------------------------ */
typedef struct vpo_field_args {
  float  *f;                 /* [nv*20] field_t array */
  int32_t nx, ny, nz;
  float   dt, cvac, eps0, damp;
  float   dx, dy, dz, dV;    /* as stored in grid_t (partition.cc:55-72) */
  float   rdx, rdy, rdz;
  int32_t bc6[6];
  /* coefficients of the single material that fills space, in material_coefficient_t order (sfa_private.h:14-25):
   * decayx drivex decayy drivey decayz drivez rmux rmuy rmuz nonconductive epsx epsy epsz.  Used when has_material
   * is nonzero; otherwise true vacuum (all ones). */
  int32_t has_material;
  float   material[13];
} vpo_field_args_t;

void vpo_advance_b(const vpo_field_args_t *a, float frac);
void vpo_vacuum_advance_e(const vpo_field_args_t *a, float frac);
void vpo_clear_jf(const vpo_field_args_t *a);
void vpo_synchronize_jf(const vpo_field_args_t *a);
void vpo_vacuum_energy_f(const vpo_field_args_t *a, double en[6]);

/* ---- divergence cleaning and shared-face synchronisation (advance.cc:138-176), same single-domain setting ---- */
void   vpo_clear_rhof(const vpo_field_args_t *a);                 /* sfa.cc:239-256 */
void   vpo_synchronize_rho(const vpo_field_args_t *a);            /* remote.cc:534-620, local.cc:376-444 */
void   vpo_vacuum_compute_div_e_err(const vpo_field_args_t *a);   /* vacuum_compute_div_e_err_pipeline.{h,cc} */
double vpo_compute_rms_div_e_err(const vpo_field_args_t *a);      /* compute_rms_div_e_err_pipeline.cc */
void   vpo_vacuum_clean_div_e(const vpo_field_args_t *a);         /* vacuum_clean_div_e_pipeline.{h,cc} */
void   vpo_compute_div_b_err(const vpo_field_args_t *a);          /* compute_div_b_err_pipeline.cc */
double vpo_compute_rms_div_b_err(const vpo_field_args_t *a);      /* compute_rms_div_b_err_pipeline.cc */
void   vpo_clean_div_b(const vpo_field_args_t *a);                /* clean_div_b_pipeline.cc */
double vpo_synchronize_tang_e_norm_b(const vpo_field_args_t *a);  /* remote.cc:298-416 */
/* initialisation-time kernels of the same table (initialize.cc): bound charge from div E, TCA from curl B */
void   vpo_vacuum_compute_rhob(const vpo_field_args_t *a);        /* vacuum_compute_rhob_pipeline.{h,cc} */
void   vpo_vacuum_compute_curl_b(const vpo_field_args_t *a);      /* vacuum_compute_curl_b_pipeline.{h,cc} */

/* ---- hydro moments (diagnostics): accumulate_hydro_p (hydro_p_pipeline.cc:19-214) for one pipeline, i.e. particles
 *      summed in array order; synchronize_hydro_array (hydro_array.cc:131-309) on a single domain.  hydro_t is
 *      16 floats per voxel: jx jy jz rho px py pz ke txx tyy tzz tyz tzx txy pad pad. ------------------------------ */
void vpo_accumulate_hydro_p(float *hydro, const vpo_particle_t *p, int32_t np, const float *interp, int32_t interp_stride,
                            float q, float m, float qdt_2mc, float cvac, float r8V, int32_t nx, int32_t ny, int32_t nz);
void vpo_synchronize_hydro(float *hydro, const vpo_field_args_t *geometry);   /* uses nx ny nz, bc6, dx dy dz */

#ifdef __cplusplus
}
#endif
#endif
