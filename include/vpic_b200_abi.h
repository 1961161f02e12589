/* vpic_b200_abi.h — binary layouts the drop-in boundary must honour.
 *
 * The reference host program (lanl/vpic) owns every array on the hot path and
 * hands raw struct pointers to the extern "C" entry points this library
 * replaces.  These declarations restate those layouts (field order, widths,
 * padding) so that libvpic_b200 can be built where the reference headers are
 * not present.  They are layouts, not behaviour; each cites the reference
 * definition it must stay byte-compatible with.  tests/test_abi.py checks
 * sizeof/offsetof of every struct below against values probed from the
 * reference headers (tests/golden/abi_layout.json).
 *
 * Padding of interpolator/accumulator depends on the host build's SIMD width
 * (src/sf_interface/sf_interface.h:27-53).  Build this library with the same
 * -DVPB_SIMD_WIDTH={4,8,16} as the host (4 = scalar or V4-only builds).
 */
#ifndef VPIC_B200_ABI_H
#define VPIC_B200_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef VPB_SIMD_WIDTH
#define VPB_SIMD_WIDTH 4
#endif

#if VPB_SIMD_WIDTH == 16
#  define VPB_INTERPOLATOR_PAD 14
#  define VPB_ACCUMULATOR_PAD   4
#elif VPB_SIMD_WIDTH == 8
#  define VPB_INTERPOLATOR_PAD  6
#  define VPB_ACCUMULATOR_PAD   4
#else
#  define VPB_INTERPOLATOR_PAD  2
#  define VPB_ACCUMULATOR_PAD   0
#endif

/* ---- particles: src/species_advance/species_advance_aos.h:21-52 ---------- */

typedef struct vpb_particle {        /* 32 B, two 128-bit halves */
  float   dx, dy, dz;                /* offset in voxel, each on [-1,1]          */
  int32_t i;                         /* voxel; 8*voxel+face while awaiting boundary_p */
  float   ux, uy, uz;                /* normalised momentum                      */
  float   w;                         /* weight                                   */
} vpb_particle_t;

typedef struct vpb_particle_mover {  /* 16 B */
  float   dispx, dispy, dispz;       /* remaining displacement, cell units       */
  int32_t i;                         /* index of the particle in the species     */
} vpb_particle_mover_t;

typedef struct vpb_particle_injector { /* 48 B = particle ++ mover(sp_id in .i slot) */
  float   dx, dy, dz; int32_t i;
  float   ux, uy, uz, w;
  float   dispx, dispy, dispz; int32_t sp_id;
} vpb_particle_injector_t;

/* ---- grid: src/grid/grid.h:73-131 ---------------------------------------- */

typedef struct vpb_grid {
  float    dt, cvac, eps0;
  int64_t  step;
  double   t0;
  float    x0, y0, z0, x1, y1, z1;
  int32_t  nx, ny, nz;
  float    dx, dy, dz, dV;
  float    rdx, rdy, rdz, r8V;
  int32_t  sx, sy, sz, nv;
  int32_t  bc[27];
  int64_t *range;                    /* [world_size+1] global voxel ranges       */
  int64_t *neighbor;                 /* [6*nv]; <0 = particle boundary code      */
  int64_t  rangel, rangeh;
  void    *mp;                       /* opaque message-passing ports             */
} vpb_grid_t;

/* particle boundary codes stored in neighbor[] (grid.h:29-30); custom handlers <= -3 */
#define VPB_REFLECT_PARTICLES (-1)
#define VPB_ABSORB_PARTICLES  (-2)

/* ---- species: src/species_advance/species_advance_aos.h:54-94 ------------ */

typedef struct vpb_species {
  char                 *name;
  float                 q, m;
  int32_t               np, max_np;
  vpb_particle_t       *p;
  int32_t               nm, max_nm;
  vpb_particle_mover_t *pm;
  int64_t               last_sorted;
  int32_t               sort_interval;
  int32_t               sort_out_of_place;
  int32_t              *partition;   /* [nv+1] first particle of each voxel after sort_p */
  vpb_grid_t           *g;
  int32_t               id;
  struct vpb_species   *next;
} vpb_species_t;

/* ---- interpolator / accumulator: src/sf_interface/sf_interface.h:62-131 --- */

typedef struct vpb_interpolator {
  float ex, dexdy, dexdz, d2exdydz;
  float ey, deydz, deydx, d2eydzdx;
  float ez, dezdx, dezdy, d2ezdxdy;
  float cbx, dcbxdx;
  float cby, dcbydy;
  float cbz, dcbzdz;
  float pad_[VPB_INTERPOLATOR_PAD];
} vpb_interpolator_t;

typedef struct vpb_interpolator_array {
  vpb_interpolator_t *i;             /* [nv] */
  vpb_grid_t         *g;
} vpb_interpolator_array_t;

typedef struct vpb_accumulator {
  float jx[4];                       /* 4x charge through the (y,z) = (-,-),(+,-),(-,+),(+,+) quarter faces */
  float jy[4];                       /* same, (z,x) ordering */
  float jz[4];                       /* same, (x,y) ordering */
#if VPB_ACCUMULATOR_PAD
  float pad_[VPB_ACCUMULATOR_PAD];
#endif
} vpb_accumulator_t;

typedef struct vpb_accumulator_array {
  vpb_accumulator_t *a;              /* [(n_pipeline+1) * stride]; block 0 is the total after reduce */
  int32_t            n_pipeline;
  int32_t            stride;
  vpb_grid_t        *g;
} vpb_accumulator_array_t;

/* ---- fields: src/field_advance/field_advance.h:152-229 -------------------- */

typedef struct vpb_field {           /* 80 B */
  float   ex, ey, ez, div_e_err;
  float   cbx, cby, cbz, div_b_err;
  float   tcax, tcay, tcaz, rhob;
  float   jfx, jfy, jfz, rhof;
  int16_t ematx, ematy, ematz, nmat;
  int16_t fmatx, fmaty, fmatz, cmat;
} vpb_field_t;

struct vpb_field_array;

typedef struct vpb_field_advance_kernels {   /* 17 entry points, field_advance.h:170-218 */
  void   (*delete_fa)(struct vpb_field_array *);
  void   (*advance_b)(struct vpb_field_array *, float frac);
  void   (*advance_e)(struct vpb_field_array *, float frac);
  void   (*energy_f)(double *en6, const struct vpb_field_array *);
  void   (*clear_jf)(struct vpb_field_array *);
  void   (*synchronize_jf)(struct vpb_field_array *);
  void   (*clear_rhof)(struct vpb_field_array *);
  void   (*synchronize_rho)(struct vpb_field_array *);
  void   (*compute_rhob)(struct vpb_field_array *);
  void   (*compute_curl_b)(struct vpb_field_array *);
  double (*synchronize_tang_e_norm_b)(struct vpb_field_array *);
  void   (*compute_div_e_err)(struct vpb_field_array *);
  double (*compute_rms_div_e_err)(const struct vpb_field_array *);
  void   (*clean_div_e)(struct vpb_field_array *);
  void   (*compute_div_b_err)(struct vpb_field_array *);
  double (*compute_rms_div_b_err)(const struct vpb_field_array *);
  void   (*clean_div_b)(struct vpb_field_array *);
} vpb_field_advance_kernels_t;

typedef struct vpb_field_array {
  vpb_field_t                *f;     /* [nv] */
  vpb_grid_t                 *g;
  void                       *params;
  vpb_field_advance_kernels_t kernel[1];
} vpb_field_array_t;

/* hydro moments: src/sf_interface/sf_interface.h:185-200 (hydro_t is 64 B for every SIMD padding) */
typedef struct vpb_hydro {
  float jx, jy, jz, rho;
  float px, py, pz, ke;
  float txx, tyy, tzz;
  float tyz, tzx, txy;
  float pad_[2];
} vpb_hydro_t;

typedef struct vpb_hydro_array {
  vpb_hydro_t *h;        /* [(n_pipeline+1) * stride], block 0 = host */
  int32_t      n_pipeline;
  int32_t      stride;
  vpb_grid_t  *g;
} vpb_hydro_array_t;

/* standard field advance parameters: src/field_advance/standard/sfa_private.h:14-34 */
typedef struct vpb_material_coefficient {
  float decayx, drivex;
  float decayy, drivey;
  float decayz, drivez;
  float rmux, rmuy, rmuz;
  float nonconductive;
  float epsx, epsy, epsz;
  float pad_[3];
} vpb_material_coefficient_t;

typedef struct vpb_sfa_params {
  vpb_material_coefficient_t *mc;
  int32_t                     n_mc;
  float                       damp;
} vpb_sfa_params_t;

#ifdef __cplusplus
}
#endif
#endif /* VPIC_B200_ABI_H */
