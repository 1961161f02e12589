/* TEST INFRASTRUCTURE ONLY — see vpic_oracle.h.  CPU restatement of the
 * reference hot path (scalar pipelines), one function per reference routine,
 * citing the reference file:line it follows.  Built with -ffp-contract=off. */
#include "vpic_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define VOX(x, y, z) ((x) + (nx + 2) * ((y) + (ny + 2) * (z)))   /* grid.h:136 */

/* interpolator member offsets (sf_interface.h:62-74) */
enum { I_EX = 0, I_DEXDY, I_DEXDZ, I_D2EXDYDZ, I_EY, I_DEYDZ, I_DEYDX, I_D2EYDZDX,
       I_EZ, I_DEZDX, I_DEZDY, I_D2EZDXDY, I_CBX, I_DCBXDX, I_CBY, I_DCBYDY, I_CBZ, I_DCBZDZ };
/* field_t member offsets in floats (field_advance.h:152-160) */
enum { F_EX = 0, F_EY, F_EZ, F_DIVE, F_CBX, F_CBY, F_CBZ, F_DIVB,
       F_TCAX, F_TCAY, F_TCAZ, F_RHOB, F_JFX, F_JFY, F_JFZ, F_RHOF, F_STRIDE = 20 };

/* ------------------------------------------------------------------------ */
/* move_p, scalar variant: move_p.cc:216-378                                 */

int vpo_move_p(vpo_particle_t *p0, vpo_mover_t *pm, float *accum, int32_t accum_stride,
               const int64_t *neighbor, int64_t rangel, int64_t rangeh, float qsp) {
  float s_midx, s_midy, s_midz, s_dispx, s_dispy, s_dispz, s_dir[3];
  float v0, v1, v2, v3, v4, v5, q;
  int axis, face;
  int64_t nb;
  float *a;
  vpo_particle_t *p = p0 + pm->i;

  q = qsp * p->w;                                                   /* :233 */
  for (;;) {
    s_midx = p->dx; s_midy = p->dy; s_midz = p->dz;
    s_dispx = pm->dispx; s_dispy = pm->dispy; s_dispz = pm->dispz;

    s_dir[0] = (s_dispx > 0.0f) ? 1.0f : -1.0f;                     /* :245-247 */
    s_dir[1] = (s_dispy > 0.0f) ? 1.0f : -1.0f;
    s_dir[2] = (s_dispz > 0.0f) ? 1.0f : -1.0f;

    v0 = (s_dispx == 0.0f) ? 3.4e38f : (s_dir[0] - s_midx) / s_dispx;   /* :251-253 */
    v1 = (s_dispy == 0.0f) ? 3.4e38f : (s_dir[1] - s_midy) / s_dispy;
    v2 = (s_dispz == 0.0f) ? 3.4e38f : (s_dir[2] - s_midz) / s_dispz;

    v3 = 2.0f; axis = 3;                                             /* :262-266 */
    if (v0 < v3) { v3 = v0; axis = 0; }
    if (v1 < v3) { v3 = v1; axis = 1; }
    if (v2 < v3) { v3 = v2; axis = 2; }
    v3 *= 0.5f;

    s_dispx *= v3; s_dispy *= v3; s_dispz *= v3;                     /* :269-275 */
    s_midx += s_dispx; s_midy += s_dispy; s_midz += s_dispz;

    /* :280 — the 1/3 is a double constant: one double multiply, then back to float */
    v5 = (float)((double)(q * s_dispx * s_dispy * s_dispz) * (1.0 / 3.0));

    a = accum + (size_t)p->i * accum_stride;
#   define ACC(sd, mY, mZ)                                            \
    v4 = q * (sd); v1 = v4 * (mY); v0 = v4 - v1; v1 += v4;            \
    v4 = 1 + (mZ); v2 = v0 * v4; v3 = v1 * v4;                        \
    v4 = 1 - (mZ); v0 *= v4; v1 *= v4;                                \
    v0 += v5; v1 -= v5; v2 -= v5; v3 += v5;                           \
    a[0] += v0; a[1] += v1; a[2] += v2; a[3] += v3
    ACC(s_dispx, s_midy, s_midz); a += 4;                            /* :284-305 */
    ACC(s_dispy, s_midz, s_midx); a += 4;
    ACC(s_dispz, s_midx, s_midy);
#   undef ACC

    pm->dispx -= s_dispx; pm->dispy -= s_dispy; pm->dispz -= s_dispz;   /* :310-312 */
    p->dx += s_dispx + s_dispx;                                      /* :315-317 */
    p->dy += s_dispy + s_dispy;
    p->dz += s_dispz + s_dispz;

    if (axis == 3) break;                                            /* :323 */

    v0 = s_dir[axis];
    (&p->dx)[axis] = v0;                                             /* :337 exactly on the face */
    face = axis;
    if (v0 > 0.0f) face += 3;
    nb = neighbor[6 * (int64_t)p->i + face];                         /* :347 */

    if (nb == -1) {                                                  /* reflect_particles :349-358 */
      (&p->ux)[axis] = -(&p->ux)[axis];
      (&pm->dispx)[axis] = -(&pm->dispx)[axis];
      continue;
    }
    if (nb < rangel || nb > rangeh) {                                /* :360-369 */
      p->i = 8 * p->i + face;
      return 1;
    }
    p->i = (int32_t)(nb - rangel);                                   /* :374-376 */
    (&p->dx)[axis] = -v0;
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* advance_p_pipeline_scalar: advance_p_pipeline.cc:20-245 (one pipeline, one accumulator block) */

int32_t vpo_advance_p(const vpo_push_args_t *A, int32_t *n_ignored) {
  const float qdt_2mc = A->qdt_2mc, cdt_dx = A->cdt_dx, cdt_dy = A->cdt_dy, cdt_dz = A->cdt_dz, qsp = A->qsp;
  const float one = 1.0, one_third = 1.0 / 3.0, two_fifteenths = 2.0 / 15.0;   /* :39-42 */
  float dx, dy, dz, ux, uy, uz, q, hax, hay, haz, cbx, cby, cbz, v0, v1, v2, v3, v4, v5;
  int32_t ii, nm = 0, ign = 0;
  vpo_particle_t *p = A->p;
  vpo_mover_t local_pm[1];

  for (int32_t n = 0; n < A->np; n++, p++) {
    dx = p->dx; dy = p->dy; dz = p->dz; ii = p->i;                  /* :91-94 */
    const float *f = A->interp + (size_t)ii * A->interp_stride;

    hax = qdt_2mc * ((f[I_EX] + dy * f[I_DEXDY]) + dz * (f[I_DEXDZ] + dy * f[I_D2EXDYDZ]));   /* :98-105 */
    hay = qdt_2mc * ((f[I_EY] + dz * f[I_DEYDZ]) + dx * (f[I_DEYDX] + dz * f[I_D2EYDZDX]));
    haz = qdt_2mc * ((f[I_EZ] + dx * f[I_DEZDX]) + dy * (f[I_DEZDY] + dx * f[I_D2EZDXDY]));
    cbx = f[I_CBX] + dx * f[I_DCBXDX];                               /* :107-109 */
    cby = f[I_CBY] + dy * f[I_DCBYDY];
    cbz = f[I_CBZ] + dz * f[I_DCBZDZ];

    ux = p->ux; uy = p->uy; uz = p->uz; q = p->w;
    ux += hax; uy += hay; uz += haz;                                 /* :116-118 */

    v0 = qdt_2mc / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));     /* :120 */
    v1 = cbx * cbx + (cby * cby + cbz * cbz);                        /* :123-127 */
    v2 = (v0 * v0) * v1;
    v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
    v4 = v3 / (one + v1 * (v3 * v3));
    v4 += v4;
    v0 = ux + v3 * (uy * cbz - uz * cby);                            /* :129-131 */
    v1 = uy + v3 * (uz * cbx - ux * cbz);
    v2 = uz + v3 * (ux * cby - uy * cbx);
    ux += v4 * (v1 * cbz - v2 * cby);                                /* :133-135 */
    uy += v4 * (v2 * cbx - v0 * cbz);
    uz += v4 * (v0 * cby - v1 * cbx);
    ux += hax; uy += hay; uz += haz;                                 /* :137-139 */
    p->ux = ux; p->uy = uy; p->uz = uz;                              /* :141-143 */

    v0 = one / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));         /* :145 */
    ux *= cdt_dx; uy *= cdt_dy; uz *= cdt_dz;
    ux *= v0; uy *= v0; uz *= v0;
    v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;                        /* :156-158 */
    v3 = v0 + ux; v4 = v1 + uy; v5 = v2 + uz;                        /* :160-162 */

    if (v3 <= one && v4 <= one && v5 <= one && -v3 <= one && -v4 <= one && -v5 <= one) {   /* :165-166 */
      q *= qsp;
      p->dx = v3; p->dy = v4; p->dz = v5;
      dx = v0; dy = v1; dz = v2;
      v5 = q * ux * uy * uz * one_third;                             /* :183 */
      float *a = A->accum + (size_t)ii * A->accum_stride;
#     define ACCUMULATE_J(uX, dY, dZ, off)                            \
      v4 = q * (uX); v1 = v4 * (dY); v0 = v4 - v1; v1 += v4;          \
      v4 = one + (dZ); v2 = v0 * v4; v3 = v1 * v4;                    \
      v4 = one - (dZ); v0 *= v4; v1 *= v4;                            \
      v0 += v5; v1 -= v5; v2 -= v5; v3 += v5;                         \
      a[off + 0] += v0; a[off + 1] += v1; a[off + 2] += v2; a[off + 3] += v3
      ACCUMULATE_J(ux, dy, dz, 0);                                   /* :187-208 */
      ACCUMULATE_J(uy, dz, dx, 4);
      ACCUMULATE_J(uz, dx, dy, 8);
#     undef ACCUMULATE_J
    } else {                                                         /* :213-238 */
      local_pm->dispx = ux; local_pm->dispy = uy; local_pm->dispz = uz;
      local_pm->i = (int32_t)(p - A->p);
      if (vpo_move_p(A->p, local_pm, A->accum, A->accum_stride, A->neighbor, A->rangel, A->rangeh, qsp)) {
        if (nm < A->max_nm) A->pm[nm++] = local_pm[0];
        else { ign++; p->i = p->i >> 3; }
      }
    }
  }
  if (n_ignored) *n_ignored = ign;
  return nm;
}

/* ------------------------------------------------------------------------ */
/* sort_p_pipeline, single-subsort branch: sort_p_pipeline.cc:142-214,345-370.
 * The multi-threaded branch is two stable passes, so the result is the same
 * stable counting sort for any thread count. */

void vpo_sort_p(vpo_particle_t *p, int32_t np, vpo_particle_t *aux, int32_t *partition,
                int32_t nx, int32_t ny, int32_t nz) {
  const int32_t nv = (nx + 2) * (ny + 2) * (nz + 2);
  const int32_t vl = VOX(1, 1, 1), vh = VOX(nx, ny, nz);
  int32_t *next = (int32_t *)calloc((size_t)nv + 1, sizeof(int32_t));
  int32_t sum = 0, v, i;
  for (i = 0; i < np; i++) next[p[i].i]++;                           /* :178-181 */
  for (v = vl; v < vh + 1; v++) {                                    /* :184-191 */
    int32_t count = next[v];
    next[v] = sum; partition[v] = sum; sum += count;
  }
  partition[vh + 1] = sum;                                           /* :193 */
  for (i = 0; i < np; i++) aux[next[p[i].i]++] = p[i];               /* :196-212 */
  for (v = 0; v < vl; v++) partition[v] = 0;                         /* :357 */
  for (v = vh + 1; v < nv; v++) partition[v] = np;                   /* :359-362 */
  memcpy(p, aux, (size_t)np * sizeof *p);                            /* :369 */
  free(next);
}

/* ------------------------------------------------------------------------ */
/* load_interpolator_pipeline_scalar: interpolator_array_pipeline.cc:21-135 */

void vpo_load_interpolator(float *interp, int32_t is, const float *fld, int32_t nx, int32_t ny, int32_t nz) {
  const float fourth = 0.25, half = 0.50;
  for (int z = 1; z <= nz; z++) for (int y = 1; y <= ny; y++) for (int x = 1; x <= nx; x++) {
    float *pi = interp + (size_t)VOX(x, y, z) * is;
    const float *pf0 = fld + (size_t)VOX(x, y, z) * F_STRIDE;
    const float *pfx = fld + (size_t)VOX(x + 1, y, z) * F_STRIDE;
    const float *pfy = fld + (size_t)VOX(x, y + 1, z) * F_STRIDE;
    const float *pfz = fld + (size_t)VOX(x, y, z + 1) * F_STRIDE;
    const float *pfyz = fld + (size_t)VOX(x, y + 1, z + 1) * F_STRIDE;
    const float *pfzx = fld + (size_t)VOX(x + 1, y, z + 1) * F_STRIDE;
    const float *pfxy = fld + (size_t)VOX(x + 1, y + 1, z) * F_STRIDE;
    float w0, w1, w2, w3;
    w0 = pf0[F_EX]; w1 = pfy[F_EX]; w2 = pfz[F_EX]; w3 = pfyz[F_EX];              /* :67-79 */
    pi[I_EX] = fourth * ((w3 + w0) + (w1 + w2));
    pi[I_DEXDY] = fourth * ((w3 - w0) + (w1 - w2));
    pi[I_DEXDZ] = fourth * ((w3 - w0) - (w1 - w2));
    pi[I_D2EXDYDZ] = fourth * ((w3 + w0) - (w1 + w2));
    w0 = pf0[F_EY]; w1 = pfz[F_EY]; w2 = pfx[F_EY]; w3 = pfzx[F_EY];              /* :82-90 */
    pi[I_EY] = fourth * ((w3 + w0) + (w1 + w2));
    pi[I_DEYDZ] = fourth * ((w3 - w0) + (w1 - w2));
    pi[I_DEYDX] = fourth * ((w3 - w0) - (w1 - w2));
    pi[I_D2EYDZDX] = fourth * ((w3 + w0) - (w1 + w2));
    w0 = pf0[F_EZ]; w1 = pfx[F_EZ]; w2 = pfy[F_EZ]; w3 = pfxy[F_EZ];              /* :93-101 */
    pi[I_EZ] = fourth * ((w3 + w0) + (w1 + w2));
    pi[I_DEZDX] = fourth * ((w3 - w0) + (w1 - w2));
    pi[I_DEZDY] = fourth * ((w3 - w0) - (w1 - w2));
    pi[I_D2EZDXDY] = fourth * ((w3 + w0) - (w1 + w2));
    w0 = pf0[F_CBX]; w1 = pfx[F_CBX]; pi[I_CBX] = half * (w1 + w0); pi[I_DCBXDX] = half * (w1 - w0);   /* :104-121 */
    w0 = pf0[F_CBY]; w1 = pfy[F_CBY]; pi[I_CBY] = half * (w1 + w0); pi[I_DCBYDY] = half * (w1 - w0);
    w0 = pf0[F_CBZ]; w1 = pfz[F_CBZ]; pi[I_CBZ] = half * (w1 + w0); pi[I_DCBZDZ] = half * (w1 - w0);
  }
}

/* ------------------------------------------------------------------------ */
/* clear_accumulator_array_pipeline (block 0): clear_array_pipeline.cc:40-67 */

void vpo_clear_accumulator(float *accum, int32_t as, int32_t nx, int32_t ny, int32_t nz) {
  int32_t i0 = (VOX(1, 1, 1) / 2) * 2;
  int32_t na = (((VOX(nx, ny, nz) - i0 + 1) + 1) / 2) * 2;
  memset(accum + (size_t)i0 * as, 0, (size_t)na * as * sizeof(float));
}

/* unload_accumulator_pipeline_scalar: unload_accumulator_pipeline.cc:18-82; coefficients :137-139 */

void vpo_unload_accumulator(float *fld, const float *accum, int32_t as, int32_t nx, int32_t ny, int32_t nz,
                            float rdx, float rdy, float rdz, float dt) {
  const float cx = 0.25 * rdy * rdz / dt;   /* double 0.25 * float ... evaluated in double, stored to float */
  const float cy = 0.25 * rdz * rdx / dt;
  const float cz = 0.25 * rdx * rdy / dt;
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++) {
    float *f0 = fld + (size_t)VOX(x, y, z) * F_STRIDE;
    const float *a0 = accum + (size_t)VOX(x, y, z) * as;
    const float *ax = accum + (size_t)VOX(x - 1, y, z) * as;
    const float *ay = accum + (size_t)VOX(x, y - 1, z) * as;
    const float *az = accum + (size_t)VOX(x, y, z - 1) * as;
    const float *ayz = accum + (size_t)VOX(x, y - 1, z - 1) * as;
    const float *azx = accum + (size_t)VOX(x - 1, y, z - 1) * as;
    const float *axy = accum + (size_t)VOX(x - 1, y - 1, z) * as;
    f0[F_JFX] += cx * (a0[0] + ay[1] + az[2] + ayz[3]);              /* :65-67 */
    f0[F_JFY] += cy * (a0[4] + az[5] + ax[6] + azx[7]);
    f0[F_JFZ] += cz * (a0[8] + ax[9] + ay[10] + axy[11]);
  }
}

/* ------------------------------------------------------------------------ */
/* energy_p: energy_p_pipeline.cc:18-115 (single pipeline order, no cross-rank sum) */

double vpo_energy_p(const vpo_particle_t *p, int32_t np, const float *interp, int32_t is,
                    float q, float m, float dt, float cvac) {
  const float qdt_2mc = (q * dt) / (2 * m * cvac);                  /* :98 */
  const float msp = m, one = 1.0;
  double en = 0.0;
  for (int32_t n = 0; n < np; n++) {
    float dx = p[n].dx, dy = p[n].dy, dz = p[n].dz, v0, v1, v2;
    const float *f = interp + (size_t)p[n].i * is;
    v0 = p[n].ux + qdt_2mc * ((f[I_EX] + dy * f[I_DEXDY]) + dz * (f[I_DEXDZ] + dy * f[I_D2EXDYDZ]));
    v1 = p[n].uy + qdt_2mc * ((f[I_EY] + dz * f[I_DEYDZ]) + dx * (f[I_DEYDX] + dz * f[I_D2EYDZDX]));
    v2 = p[n].uz + qdt_2mc * ((f[I_EZ] + dx * f[I_DEZDX]) + dy * (f[I_DEZDY] + dx * f[I_D2EZDXDY]));
    v0 = v0 * v0 + v1 * v1 + v2 * v2;                                /* :62 */
    v0 = (msp * p[n].w) * (v0 / (one + sqrtf(one + v0)));            /* :64 */
    en += (double)v0;
  }
  return en * ((double)cvac * (double)cvac);
}

/* center_p / uncenter_p scalar pipelines: center_p_pipeline.cc:17-96, uncenter_p_pipeline.cc:17-98 */

void vpo_center_p(vpo_particle_t *p, int32_t np, const float *interp, int32_t is, float qdt_2mc) {
  const float qdt_4mc = 0.5 * qdt_2mc, one = 1.0, one_third = 1.0 / 3.0, two_fifteenths = 2.0 / 15.0;
  for (int32_t n = 0; n < np; n++, p++) {
    float dx = p->dx, dy = p->dy, dz = p->dz, hax, hay, haz, cbx, cby, cbz, ux, uy, uz, v0, v1, v2, v3, v4;
    const float *f = interp + (size_t)p->i * is;
    hax = qdt_2mc * ((f[I_EX] + dy * f[I_DEXDY]) + dz * (f[I_DEXDZ] + dy * f[I_D2EXDYDZ]));
    hay = qdt_2mc * ((f[I_EY] + dz * f[I_DEYDZ]) + dx * (f[I_DEYDX] + dz * f[I_D2EYDZDX]));
    haz = qdt_2mc * ((f[I_EZ] + dx * f[I_DEZDX]) + dy * (f[I_DEZDY] + dx * f[I_D2EZDXDY]));
    cbx = f[I_CBX] + dx * f[I_DCBXDX]; cby = f[I_CBY] + dy * f[I_DCBYDY]; cbz = f[I_CBZ] + dz * f[I_DCBZDZ];
    ux = p->ux; uy = p->uy; uz = p->uz;
    ux += hax; uy += hay; uz += haz;
    v0 = qdt_4mc / (float)sqrt(one + (ux * ux + (uy * uy + uz * uz)));   /* double sqrt of a float expr */
    v1 = cbx * cbx + (cby * cby + cbz * cbz);
    v2 = (v0 * v0) * v1;
    v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
    v4 = v3 / (one + v1 * (v3 * v3)); v4 += v4;
    v0 = ux + v3 * (uy * cbz - uz * cby);
    v1 = uy + v3 * (uz * cbx - ux * cbz);
    v2 = uz + v3 * (ux * cby - uy * cbx);
    ux += v4 * (v1 * cbz - v2 * cby);
    uy += v4 * (v2 * cbx - v0 * cbz);
    uz += v4 * (v0 * cby - v1 * cbx);
    p->ux = ux; p->uy = uy; p->uz = uz;
  }
}

void vpo_uncenter_p(vpo_particle_t *p, int32_t np, const float *interp, int32_t is, float qdt_2mc_in) {
  const float qdt_2mc = -qdt_2mc_in;            /* uncenter_p_pipeline.cc: caller passes -q dt/2mc ... see .cc:120 */
  const float qdt_4mc = 0.5 * qdt_2mc, one = 1.0, one_third = 1.0 / 3.0, two_fifteenths = 2.0 / 15.0;
  for (int32_t n = 0; n < np; n++, p++) {
    float dx = p->dx, dy = p->dy, dz = p->dz, hax, hay, haz, cbx, cby, cbz, ux, uy, uz, v0, v1, v2, v3, v4;
    const float *f = interp + (size_t)p->i * is;
    hax = qdt_2mc * ((f[I_EX] + dy * f[I_DEXDY]) + dz * (f[I_DEXDZ] + dy * f[I_D2EXDYDZ]));
    hay = qdt_2mc * ((f[I_EY] + dz * f[I_DEYDZ]) + dx * (f[I_DEYDX] + dz * f[I_D2EYDZDX]));
    haz = qdt_2mc * ((f[I_EZ] + dx * f[I_DEZDX]) + dy * (f[I_DEZDY] + dx * f[I_D2EZDXDY]));
    cbx = f[I_CBX] + dx * f[I_DCBXDX]; cby = f[I_CBY] + dy * f[I_DCBYDY]; cbz = f[I_CBZ] + dz * f[I_DCBZDZ];
    ux = p->ux; uy = p->uy; uz = p->uz;
    v0 = qdt_4mc / (float)sqrt(one + (ux * ux + (uy * uy + uz * uz)));
    v1 = cbx * cbx + (cby * cby + cbz * cbz);
    v2 = (v0 * v0) * v1;
    v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
    v4 = v3 / (one + v1 * (v3 * v3)); v4 += v4;
    v0 = ux + v3 * (uy * cbz - uz * cby);
    v1 = uy + v3 * (uz * cbx - ux * cbz);
    v2 = uz + v3 * (ux * cby - uy * cbx);
    ux += v4 * (v1 * cbz - v2 * cby);
    uy += v4 * (v2 * cbx - v0 * cbz);
    uz += v4 * (v0 * cby - v1 * cbx);
    ux += hax; uy += hay; uz += haz;
    p->ux = ux; p->uy = uy; p->uz = uz;
  }
}

/* trilinear node weights shared by accumulate_rho_p / accumulate_rhob (rho_p.cc:60-75,140-153) */
static void trilinear8(float w0, float w1, float dz, float w7, float w[8]) {
  float w2, w3, w4, w5, w6;
  w6 = w7 - w0 * w7; w7 = w7 + w0 * w7;
  w4 = w6 - w1 * w6; w5 = w7 - w1 * w7;
  w6 = w6 + w1 * w6; w7 = w7 + w1 * w7;
  w0 = w4 - dz * w4; w1 = w5 - dz * w5; w2 = w6 - dz * w6; w3 = w7 - dz * w7;
  w4 = w4 + dz * w4; w5 = w5 + dz * w5; w6 = w6 + dz * w6; w7 = w7 + dz * w7;
  w[0] = w0; w[1] = w1; w[2] = w2; w[3] = w3; w[4] = w4; w[5] = w5; w[6] = w6; w[7] = w7;
}

void vpo_accumulate_rho_p(float *F, const vpo_particle_t *p, int32_t np, float q, float r8V,
                          int32_t nx, int32_t ny, int32_t nz) {
  const float q_8V = q * r8V;                                        /* rho_p.cc:32 */
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2);
  (void)nz;
  for (int32_t n = 0; n < np; n++) {
    float w[8];
    const int v = p[n].i;
    trilinear8(p[n].dx, p[n].dy, p[n].dz, p[n].w * q_8V, w);
    F[(size_t)(v) * F_STRIDE + F_RHOF] += w[0];           F[(size_t)(v + 1) * F_STRIDE + F_RHOF] += w[1];
    F[(size_t)(v + sy) * F_STRIDE + F_RHOF] += w[2];      F[(size_t)(v + sy + 1) * F_STRIDE + F_RHOF] += w[3];
    F[(size_t)(v + sz) * F_STRIDE + F_RHOF] += w[4];      F[(size_t)(v + sz + 1) * F_STRIDE + F_RHOF] += w[5];
    F[(size_t)(v + sz + sy) * F_STRIDE + F_RHOF] += w[6]; F[(size_t)(v + sz + sy + 1) * F_STRIDE + F_RHOF] += w[7];
  }
}

void vpo_accumulate_rhob(float *F, const vpo_particle_t *p, float qsp, float r8V, int32_t nx, int32_t ny, int32_t nz) {
  const int sy = nx + 2, sz = (nx + 2) * (ny + 2);
  float w[8];
  const int v = p->i;
  trilinear8(p->dx, p->dy, p->dz, (qsp * r8V) * p->w, w);            /* rho_p.cc:139 */
  int x = v, z = x / sz, y;
  if (z == 1)  { w[0] += w[0]; w[1] += w[1]; w[2] += w[2]; w[3] += w[3]; }
  if (z == nz) { w[4] += w[4]; w[5] += w[5]; w[6] += w[6]; w[7] += w[7]; }
  x -= sz * z; y = x / sy;
  if (y == 1)  { w[0] += w[0]; w[1] += w[1]; w[4] += w[4]; w[5] += w[5]; }
  if (y == ny) { w[2] += w[2]; w[3] += w[3]; w[6] += w[6]; w[7] += w[7]; }
  x -= sy * y;
  if (x == 1)  { w[0] += w[0]; w[2] += w[2]; w[4] += w[4]; w[6] += w[6]; }
  if (x == nx) { w[1] += w[1]; w[3] += w[3]; w[5] += w[5]; w[7] += w[7]; }
  F[(size_t)(v) * F_STRIDE + F_RHOB] += w[0];           F[(size_t)(v + 1) * F_STRIDE + F_RHOB] += w[1];
  F[(size_t)(v + sy) * F_STRIDE + F_RHOB] += w[2];      F[(size_t)(v + sy + 1) * F_STRIDE + F_RHOB] += w[3];
  F[(size_t)(v + sz) * F_STRIDE + F_RHOB] += w[4];      F[(size_t)(v + sz + 1) * F_STRIDE + F_RHOB] += w[5];
  F[(size_t)(v + sz + sy) * F_STRIDE + F_RHOB] += w[6]; F[(size_t)(v + sz + sy + 1) * F_STRIDE + F_RHOB] += w[7];
}

/* ======================================================================== */
/* Standard field advance, vacuum material, single local domain.            */

typedef struct { int n[3]; int s[3]; } dims_t;   /* n = cells per axis, s = voxel stride per axis */

static dims_t mkdims(const vpo_field_args_t *a) {
  dims_t d; d.n[0] = a->nx; d.n[1] = a->ny; d.n[2] = a->nz;
  d.s[0] = 1; d.s[1] = a->nx + 2; d.s[2] = (a->nx + 2) * (a->ny + 2);
  return d;
}
static float axis_d(const vpo_field_args_t *a, int X) { return X == 0 ? a->dx : X == 1 ? a->dy : a->dz; }
static float axis_rd(const vpo_field_args_t *a, int X) { return X == 0 ? a->rdx : X == 1 ? a->rdy : a->rdz; }

/* Loop over plane X=xp with Y in [yl,yh], Z in [zl,zh] (cyclic axis naming as in local.cc/remote.cc macros:
 * the reference's XYZ_LOOP always iterates z outer, y middle, x inner in *physical* axes; message packing
 * order therefore depends on the physical order, reproduced here). */
#define PLANE_LOOP(X, xp, Y, yl, yh, Z, zl, zh, ...) do {                                   \
    int lo_[3], hi_[3], c_[3];                                                               \
    lo_[X] = hi_[X] = (xp); lo_[Y] = (yl); hi_[Y] = (yh); lo_[Z] = (zl); hi_[Z] = (zh);      \
    for (c_[2] = lo_[2]; c_[2] <= hi_[2]; c_[2]++)                                           \
      for (c_[1] = lo_[1]; c_[1] <= hi_[1]; c_[1]++)                                         \
        for (c_[0] = lo_[0]; c_[0] <= hi_[0]; c_[0]++) {                                     \
          int v = c_[0] * d.s[0] + c_[1] * d.s[1] + c_[2] * d.s[2]; (void)v; __VA_ARGS__; } } while (0)

void vpo_clear_jf(const vpo_field_args_t *a) {                       /* sfa.cc:231-237 */
  int nv = (a->nx + 2) * (a->ny + 2) * (a->nz + 2);
  for (int v = 0; v < nv; v++) { float *f = a->f + (size_t)v * F_STRIDE; f[F_JFX] = 0; f[F_JFY] = 0; f[F_JFZ] = 0; }
}

/* advance_b: advance_b_pipeline.cc:20-125, stencil advance_b_pipeline.h:26-28,57-59; local_adjust_norm_b local.cc:268-297 */
void vpo_advance_b(const vpo_field_args_t *a, float frac) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const float px = (nx > 1) ? frac * a->cvac * a->dt * a->rdx : 0;
  const float py = (ny > 1) ? frac * a->cvac * a->dt * a->rdy : 0;
  const float pz = (nz > 1) ? frac * a->cvac * a->dt * a->rdz : 0;
  float *F = a->f;
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++) {
    if (y <= ny && z <= nz)
      FF(x, y, z, F_CBX) -= (py * (FF(x, y + 1, z, F_EZ) - FF(x, y, z, F_EZ)) - pz * (FF(x, y, z + 1, F_EY) - FF(x, y, z, F_EY)));
    if (z <= nz && x <= nx)
      FF(x, y, z, F_CBY) -= (pz * (FF(x, y, z + 1, F_EX) - FF(x, y, z, F_EX)) - px * (FF(x + 1, y, z, F_EZ) - FF(x, y, z, F_EZ)));
    if (x <= nx && y <= ny)
      FF(x, y, z, F_CBZ) -= (px * (FF(x + 1, y, z, F_EY) - FF(x, y, z, F_EY)) - py * (FF(x, y + 1, z, F_EX) - FF(x, y, z, F_EX)));
  }
  /* local_adjust_norm_b: only symmetric_fields (-2) zeroes the normal B on the face */
  dims_t d = mkdims(a);
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] != -2) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z], F[(size_t)v * F_STRIDE + F_CBX + X] = 0);
  }
# undef FF
}

/* Tangential-B ghosts: begin/end_remote_ghost_tang_b (remote.cc:61-134) for faces that are periodic onto this
 * same domain, local_ghost_tang_b (local.cc:50-130) for pec / symmetric / pmc faces. */
static void ghost_tang_b(const vpo_field_args_t *a) {
  dims_t d = mkdims(a);
  float *F = a->f;
  float *msg[6] = {0};
  /* pack ("send") first for every periodic face, as the reference does before any ghost is written */
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] < 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X];
    float *p = msg[fc] = (float *)malloc(sizeof(float) * (size_t)(1 + d.n[Y] * (d.n[Z] + 1) + d.n[Z] * (d.n[Y] + 1)));
    *p++ = axis_d(a, X);
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], *p++ = F[(size_t)v * F_STRIDE + F_CBX + Y]);   /* ZY edge loop: cbY */
    PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, *p++ = F[(size_t)v * F_STRIDE + F_CBX + Z]);   /* YZ edge loop: cbZ */
  }
  /* local boundary conditions */
  const float higend = (a->nx > 1 || a->ny > 1 || a->nz > 1) ? 1.03527618 : 1.;          /* local.cc:66 */
  for (int fc = 0; fc < 6; fc++) {
    int bc = a->bc6[fc];
    if (bc >= 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    int ghost = fc < 3 ? 0 : d.n[X] + 1, step = fc < 3 ? d.s[X] : -d.s[X];   /* f(x-i) = one cell inward */
    if (bc == -4) {                                                 /* absorb_fields: local.cc:84-112 */
      const float cdt_dX = a->cvac * a->dt * axis_rd(a, X), cdt_dY = a->cvac * a->dt * axis_rd(a, Y),
                  cdt_dZ = a->cvac * a->dt * axis_rd(a, Z);
      float drive = cdt_dX * higend, decay = (1 - drive) / (1 + drive);
      drive = 2 * drive / (1 + drive);
      const int face = fc < 3 ? 1 : d.n[X] + 1;
      const int to_face = (face - ghost) * d.s[X];                  /* ghost voxel -> same (Y,Z) on the face plane */
      PLANE_LOOP(X, ghost, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], {
        float *fg = F + (size_t)v * F_STRIDE, *fh = F + (size_t)(v + step) * F_STRIDE;
        const float *ff = F + (size_t)(v + to_face) * F_STRIDE, *ffi = F + (size_t)(v + to_face + step) * F_STRIDE;
        float t1 = cdt_dX * (ffi[F_EX + Z] - ff[F_EX + Z]);
        t1 = fc < 3 ? t1 : -t1;
        float t2 = F[(size_t)(v + step + d.s[Z]) * F_STRIDE + F_EX + X];
        t2 = cdt_dZ * (t2 - fh[F_EX + X]);
        fg[F_CBX + Y] = decay * fg[F_CBX + Y] + drive * fh[F_CBX + Y] - t1 + t2;
      });
      PLANE_LOOP(X, ghost, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, {
        float *fg = F + (size_t)v * F_STRIDE, *fh = F + (size_t)(v + step) * F_STRIDE;
        const float *ff = F + (size_t)(v + to_face) * F_STRIDE, *ffi = F + (size_t)(v + to_face + step) * F_STRIDE;
        float t1 = cdt_dX * (ffi[F_EX + Y] - ff[F_EX + Y]);
        t1 = fc < 3 ? t1 : -t1;
        float t2 = F[(size_t)(v + step + d.s[Y]) * F_STRIDE + F_EX + X];
        t2 = cdt_dY * (t2 - fh[F_EX + X]);
        fg[F_CBX + Z] = decay * fg[F_CBX + Z] + drive * fh[F_CBX + Z] + t1 - t2;
      });
      continue;
    }
    float sgn = (bc == -1) ? 1.0f : -1.0f;                          /* pec copies, symmetric/pmc negate */
    if (bc != -1 && bc != -2 && bc != -3) abort();
    PLANE_LOOP(X, ghost, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z],
               F[(size_t)v * F_STRIDE + F_CBX + Y] = (sgn > 0) ? F[(size_t)(v + step) * F_STRIDE + F_CBX + Y] : -F[(size_t)(v + step) * F_STRIDE + F_CBX + Y]);
    PLANE_LOOP(X, ghost, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1,
               F[(size_t)v * F_STRIDE + F_CBX + Z] = (sgn > 0) ? F[(size_t)(v + step) * F_STRIDE + F_CBX + Z] : -F[(size_t)(v + step) * F_STRIDE + F_CBX + Z]);
  }
  /* unpack ("recv"): the message sent out of port fc lands in the opposite ghost plane */
  for (int fc = 0; fc < 6; fc++) {
    if (!msg[fc]) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    int ghost = fc < 3 ? d.n[X] + 1 : 0, step = fc < 3 ? -d.s[X] : d.s[X];   /* field(x+i) with i = sender port dir */
    float *p = msg[fc];
    float lw = *p++, dX = axis_d(a, X);
    float rw = (2. * dX) / (lw + dX);
    lw = (lw - dX) / (lw + dX);
    PLANE_LOOP(X, ghost, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z],
               F[(size_t)v * F_STRIDE + F_CBX + Y] = rw * (*p++) + lw * F[(size_t)(v + step) * F_STRIDE + F_CBX + Y]);
    PLANE_LOOP(X, ghost, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1,
               F[(size_t)v * F_STRIDE + F_CBX + Z] = rw * (*p++) + lw * F[(size_t)(v + step) * F_STRIDE + F_CBX + Z]);
    free(msg[fc]);
  }
}

/* vacuum_advance_e: vacuum_advance_e_pipeline.cc:20-332, stencil .h:18-69; local_adjust_tang_e local.cc:224-265 */
void vpo_vacuum_advance_e(const vpo_field_args_t *a, float frac) {
  if (frac != 1) abort();                                           /* .cc:58-61 */
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const float damp = a->damp;
  /* the single material's coefficients (vacuum_advance_e_pipeline.h:18-26); true vacuum: decay=1, drive=1/eps=1,
     rmu=1 (sfa.cc:119-136 with eps=mu=1, sigma=0) */
  const float *m = a->material;
  const int hm = a->has_material;
  const float decayx = hm ? m[0] : 1, drivex = hm ? m[1] : 1, decayy = hm ? m[2] : 1, drivey = hm ? m[3] : 1;
  const float decayz = hm ? m[4] : 1, drivez = hm ? m[5] : 1, rmux = hm ? m[6] : 1, rmuy = hm ? m[7] : 1, rmuz = hm ? m[8] : 1;
  const float px_muz = ((nx > 1) ? (1 + damp) * a->cvac * a->dt * a->rdx : 0) * rmuz;
  const float px_muy = ((nx > 1) ? (1 + damp) * a->cvac * a->dt * a->rdx : 0) * rmuy;
  const float py_mux = ((ny > 1) ? (1 + damp) * a->cvac * a->dt * a->rdy : 0) * rmux;
  const float py_muz = ((ny > 1) ? (1 + damp) * a->cvac * a->dt * a->rdy : 0) * rmuz;
  const float pz_muy = ((nz > 1) ? (1 + damp) * a->cvac * a->dt * a->rdz : 0) * rmuy;
  const float pz_mux = ((nz > 1) ? (1 + damp) * a->cvac * a->dt * a->rdz : 0) * rmux;
  const float cj = a->dt / a->eps0;
  float *F = a->f;

  ghost_tang_b(a);

# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++) {
    if (x <= nx) {
      FF(x, y, z, F_TCAX) = (py_muz * (FF(x, y, z, F_CBZ) - FF(x, y - 1, z, F_CBZ)) -
                             pz_muy * (FF(x, y, z, F_CBY) - FF(x, y, z - 1, F_CBY))) - damp * FF(x, y, z, F_TCAX);
      FF(x, y, z, F_EX) = decayx * FF(x, y, z, F_EX) + drivex * (FF(x, y, z, F_TCAX) - cj * FF(x, y, z, F_JFX));
    }
    if (y <= ny) {
      FF(x, y, z, F_TCAY) = (pz_mux * (FF(x, y, z, F_CBX) - FF(x, y, z - 1, F_CBX)) -
                             px_muz * (FF(x, y, z, F_CBZ) - FF(x - 1, y, z, F_CBZ))) - damp * FF(x, y, z, F_TCAY);
      FF(x, y, z, F_EY) = decayy * FF(x, y, z, F_EY) + drivey * (FF(x, y, z, F_TCAY) - cj * FF(x, y, z, F_JFY));
    }
    if (z <= nz) {
      FF(x, y, z, F_TCAZ) = (px_muy * (FF(x, y, z, F_CBY) - FF(x - 1, y, z, F_CBY)) -
                             py_mux * (FF(x, y, z, F_CBX) - FF(x, y - 1, z, F_CBX))) - damp * FF(x, y, z, F_TCAZ);
      FF(x, y, z, F_EZ) = decayz * FF(x, y, z, F_EZ) + drivez * (FF(x, y, z, F_TCAZ) - cj * FF(x, y, z, F_JFZ));
    }
  }
# undef FF
  /* local_adjust_tang_e: pec zeroes tangential e and tca on the face */
  dims_t d = mkdims(a);
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] != -1) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, (F[(size_t)v * F_STRIDE + F_EX + Y] = 0, F[(size_t)v * F_STRIDE + F_TCAX + Y] = 0));
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], (F[(size_t)v * F_STRIDE + F_EX + Z] = 0, F[(size_t)v * F_STRIDE + F_TCAX + Z] = 0));
  }
}

/* synchronize_jf: remote.cc:417-508 with local_adjust_jf local.cc:335-366 */
void vpo_synchronize_jf(const vpo_field_args_t *a) {
  dims_t d = mkdims(a);
  float *F = a->f;
  for (int fc = 0; fc < 6; fc++) {                                  /* local_adjust_jf, order -x,-y,-z,+x,+y,+z */
    int bc = a->bc6[fc];
    if (bc >= 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    if (bc == -1) {
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, F[(size_t)v * F_STRIDE + F_JFX + Y] = 0);
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], F[(size_t)v * F_STRIDE + F_JFX + Z] = 0);
    } else {
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, F[(size_t)v * F_STRIDE + F_JFX + Y] *= 2.);
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], F[(size_t)v * F_STRIDE + F_JFX + Z] *= 2.);
    }
  }
  for (int X = 0; X < 3; X++) {                                     /* exchange one axis at a time */
    int Y = (X + 1) % 3, Z = (X + 2) % 3;
    float *msg[2] = {0, 0};
    for (int side = 0; side < 2; side++) {                          /* pack both before either is applied */
      int fc = X + 3 * side;
      if (a->bc6[fc] < 0) continue;
      int face = side == 0 ? 1 : d.n[X] + 1;
      float *p = msg[side] = (float *)malloc(sizeof(float) * (size_t)(1 + d.n[Y] * (d.n[Z] + 1) + d.n[Z] * (d.n[Y] + 1)));
      *p++ = axis_d(a, X);
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, *p++ = F[(size_t)v * F_STRIDE + F_JFX + Y]);
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], *p++ = F[(size_t)v * F_STRIDE + F_JFX + Z]);
    }
    for (int side = 0; side < 2; side++) {
      if (!msg[side]) continue;
      float *p = msg[side];
      float rw = *p++, dX = axis_d(a, X), lw = rw + dX;
      rw /= lw; lw = dX / lw; lw += lw; rw += rw;
      int face = side == 0 ? d.n[X] + 1 : 1;                        /* lands on the opposite shared plane */
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1,
                 F[(size_t)v * F_STRIDE + F_JFX + Y] = lw * F[(size_t)v * F_STRIDE + F_JFX + Y] + rw * (*p++));
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z],
                 F[(size_t)v * F_STRIDE + F_JFX + Z] = lw * F[(size_t)v * F_STRIDE + F_JFX + Z] + rw * (*p++));
      free(msg[side]);
    }
  }
}

/* vacuum_energy_f: vacuum_energy_f_pipeline.cc:12-97, stencil .h:24-75 (single pipeline order) */
void vpo_vacuum_energy_f(const vpo_field_args_t *a, double en[6]) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const float *m = a->material;
  const int hm = a->has_material;                              /* .h:24-29: 0.25*eps, 0.50*rmu of the single material */
  const float qepsx = 0.25 * (hm ? m[10] : 1.0f), qepsy = 0.25 * (hm ? m[11] : 1.0f), qepsz = 0.25 * (hm ? m[12] : 1.0f);
  const float hrmux = 0.50 * (hm ? m[6] : 1.0f), hrmuy = 0.50 * (hm ? m[7] : 1.0f), hrmuz = 0.50 * (hm ? m[8] : 1.0f);
  double e0 = 0, e1 = 0, e2 = 0, b0 = 0, b1 = 0, b2 = 0;
  const float *F = a->f;
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz; z++) for (int y = 1; y <= ny; y++) for (int x = 1; x <= nx; x++) {
    e0 += qepsx * (FF(x, y, z, F_EX) * FF(x, y, z, F_EX) + FF(x, y + 1, z, F_EX) * FF(x, y + 1, z, F_EX) +
                  FF(x, y, z + 1, F_EX) * FF(x, y, z + 1, F_EX) + FF(x, y + 1, z + 1, F_EX) * FF(x, y + 1, z + 1, F_EX));
    e1 += qepsy * (FF(x, y, z, F_EY) * FF(x, y, z, F_EY) + FF(x, y, z + 1, F_EY) * FF(x, y, z + 1, F_EY) +
                  FF(x + 1, y, z, F_EY) * FF(x + 1, y, z, F_EY) + FF(x + 1, y, z + 1, F_EY) * FF(x + 1, y, z + 1, F_EY));
    e2 += qepsz * (FF(x, y, z, F_EZ) * FF(x, y, z, F_EZ) + FF(x + 1, y, z, F_EZ) * FF(x + 1, y, z, F_EZ) +
                  FF(x, y + 1, z, F_EZ) * FF(x, y + 1, z, F_EZ) + FF(x + 1, y + 1, z, F_EZ) * FF(x + 1, y + 1, z, F_EZ));
    b0 += hrmux * (FF(x, y, z, F_CBX) * FF(x, y, z, F_CBX) + FF(x + 1, y, z, F_CBX) * FF(x + 1, y, z, F_CBX));
    b1 += hrmuy * (FF(x, y, z, F_CBY) * FF(x, y, z, F_CBY) + FF(x, y + 1, z, F_CBY) * FF(x, y + 1, z, F_CBY));
    b2 += hrmuz * (FF(x, y, z, F_CBZ) * FF(x, y, z, F_CBZ) + FF(x, y, z + 1, F_CBZ) * FF(x, y, z + 1, F_CBZ));
  }
# undef FF
  double v0 = 0.5 * a->eps0 * a->dV;                            /* .cc:84 */
  en[0] = e0 * v0; en[1] = e1 * v0; en[2] = e2 * v0; en[3] = b0 * v0; en[4] = b1 * v0; en[5] = b2 * v0;
}


/* ======================================================================== */
/* Divergence cleaning (Marder passes) and shared-face synchronisation.     */

#define FV(v, m) F[(size_t)(v) * F_STRIDE + (m)]

static void adjust_tang_e(const vpo_field_args_t *a) {             /* local_adjust_tang_e, local.cc:224-265 */
  dims_t d = mkdims(a);
  float *F = a->f;
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] != -1) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, (FV(v, F_EX + Y) = 0, FV(v, F_TCAX + Y) = 0));
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], (FV(v, F_EX + Z) = 0, FV(v, F_TCAX + Z) = 0));
  }
}

static void adjust_norm_b(const vpo_field_args_t *a) {             /* local_adjust_norm_b, local.cc:266-297 */
  dims_t d = mkdims(a);
  float *F = a->f;
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] != -2) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z], FV(v, F_CBX + X) = 0);
  }
}

void vpo_clear_rhof(const vpo_field_args_t *a) {
  int nv = (a->nx + 2) * (a->ny + 2) * (a->nz + 2);
  for (int v = 0; v < nv; v++) a->f[(size_t)v * F_STRIDE + F_RHOF] = 0;
}

void vpo_synchronize_rho(const vpo_field_args_t *a) {
  dims_t d = mkdims(a);
  float *F = a->f;
  for (int fc = 0; fc < 6; fc++) {                                  /* local_adjust_rhof: pec zeroes, others double */
    int bc = a->bc6[fc];
    if (bc >= 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    if (bc == -1) PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, FV(v, F_RHOF) = 0);
    else          PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, FV(v, F_RHOF) *= 2);
  }
  for (int fc = 0; fc < 6; fc++) {                                  /* local_adjust_rhob: pec zeroes */
    if (a->bc6[fc] != -1) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, FV(v, F_RHOB) = 0);
  }
  for (int X = 0; X < 3; X++) {                                     /* one axis at a time; both sides packed first */
    int Y = (X + 1) % 3, Z = (X + 2) % 3;
    float *msg[2] = {0, 0};
    for (int side = 0; side < 2; side++) {
      if (a->bc6[X + 3 * side] < 0) continue;
      int face = side == 0 ? 1 : d.n[X] + 1;
      float *p = msg[side] = (float *)malloc(sizeof(float) * (size_t)(1 + 2 * (d.n[Y] + 1) * (d.n[Z] + 1)));
      *p++ = axis_d(a, X);
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, (*p++ = FV(v, F_RHOF), *p++ = FV(v, F_RHOB)));
    }
    for (int side = 0; side < 2; side++) {
      if (!msg[side]) continue;
      float *p = msg[side];
      float hrw = *p++, dX = axis_d(a, X), hlw = hrw + dX;
      hrw /= hlw; hlw = dX / hlw;
      float lw = hlw + hlw, rw = hrw + hrw;
      int face = side == 0 ? d.n[X] + 1 : 1;
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, {
        float rf = *p++, rb = *p++;
        FV(v, F_RHOF) = lw * FV(v, F_RHOF) + rw * rf;
        FV(v, F_RHOB) = hlw * FV(v, F_RHOB) + hrw * rb;
      });
      free(msg[side]);
    }
  }
}

/* Normal-E ghosts: begin/end_remote_ghost_norm_e (remote.cc:136-206) for periodic-self faces, local_ghost_norm_e
 * (local.cc:128-179) for local walls. */
static void ghost_norm_e(const vpo_field_args_t *a) {
  dims_t d = mkdims(a);
  float *F = a->f;
  float *msg[6] = {0};
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] < 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X];
    float *p = msg[fc] = (float *)malloc(sizeof(float) * (size_t)(1 + (d.n[Y] + 1) * (d.n[Z] + 1)));
    *p++ = axis_d(a, X);
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, *p++ = FV(v, F_EX + X));
  }
  for (int fc = 0; fc < 6; fc++) {
    int bc = a->bc6[fc];
    if (bc >= 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    int ghost = fc < 3 ? 0 : d.n[X] + 1, in1 = fc < 3 ? d.s[X] : -d.s[X];
    if (bc == -1)
      PLANE_LOOP(X, ghost, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1,
                 (FV(v, F_EX + X) = FV(v + in1, F_EX + X), FV(v, F_TCAX + X) = FV(v + in1, F_TCAX + X)));
    else if (bc == -2 || bc == -3)
      PLANE_LOOP(X, ghost, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1,
                 (FV(v, F_EX + X) = -FV(v + in1, F_EX + X), FV(v, F_TCAX + X) = -FV(v + in1, F_TCAX + X)));
    else if (bc == -4)
      PLANE_LOOP(X, ghost, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1,
                 (FV(v, F_EX + X) = 2 * FV(v + in1, F_EX + X) - FV(v + 2 * in1, F_EX + X),
                  FV(v, F_TCAX + X) = 2 * FV(v + in1, F_TCAX + X) - FV(v + 2 * in1, F_TCAX + X)));
    else abort();
  }
  for (int fc = 0; fc < 6; fc++) {
    if (!msg[fc]) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    int ghost = fc < 3 ? d.n[X] + 1 : 0, step = fc < 3 ? -d.s[X] : d.s[X];
    float *p = msg[fc];
    float lw = *p++, dX = axis_d(a, X);
    float rw = (2. * dX) / (lw + dX);
    lw = (lw - dX) / (lw + dX);
    PLANE_LOOP(X, ghost, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, FV(v, F_EX + X) = rw * (*p++) + lw * FV(v + step, F_EX + X));
    free(msg[fc]);
  }
}

void vpo_vacuum_compute_div_e_err(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const int hm = a->has_material;
  const float *m = a->material;
  const float nc = hm ? m[9] : 1.0f;
  const float px = ((nx > 1) ? a->rdx : 0) * (hm ? m[10] : 1.0f);
  const float py = ((ny > 1) ? a->rdy : 0) * (hm ? m[11] : 1.0f);
  const float pz = ((nz > 1) ? a->rdz : 0) * (hm ? m[12] : 1.0f);
  const float cj = 1. / a->eps0;
  float *F = a->f;
  ghost_norm_e(a);
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++)
    FF(x, y, z, F_DIVE) = nc * (px * (FF(x, y, z, F_EX) - FF(x - 1, y, z, F_EX)) +
                                py * (FF(x, y, z, F_EY) - FF(x, y - 1, z, F_EY)) +
                                pz * (FF(x, y, z, F_EZ) - FF(x, y, z - 1, F_EZ)) -
                                cj * (FF(x, y, z, F_RHOF) + FF(x, y, z, F_RHOB)));
# undef FF
  dims_t d = mkdims(a);                                              /* local_adjust_div_e, local.cc:298-330 */
  for (int fc = 0; fc < 6; fc++) {
    int bc = a->bc6[fc];
    if (bc != -1 && bc != -4) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, FV(v, F_DIVE) = 0);
  }
}

double vpo_compute_rms_div_e_err(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const float *F = a->f;
  double err = 0;
# define D(x, y, z) F[(size_t)VOX(x, y, z) * F_STRIDE + F_DIVE]
  /* interior nodes: float product summed in double (pipeline :38); walls, edges and corners weigh 1/2, 1/4, 1/8 and
     multiply in double (:96-160) */
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++) {
    int on = (x == 1 || x == nx + 1) + (y == 1 || y == ny + 1) + (z == 1 || z == nz + 1);
    float e = D(x, y, z);
    if (on == 0) err += e * e;
    else err += (on == 1 ? 0.5 : on == 2 ? 0.25 : 0.125) * (double)e * (double)e;
  }
# undef D
  double l0 = err * a->dV, l1 = (nx * ny * nz) * a->dV;
  return a->eps0 * sqrt(l0 / l1);
}

void vpo_vacuum_clean_div_e(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const int hm = a->has_material;
  const float *m = a->material;
  const float rdx = (nx > 1) ? a->rdx : 0, rdy = (ny > 1) ? a->rdy : 0, rdz = (nz > 1) ? a->rdz : 0;
  const float alphadt = 0.3888889 / (rdx * rdx + rdy * rdy + rdz * rdz);
  const float px = (alphadt * rdx) * (hm ? m[1] : 1.0f);
  const float py = (alphadt * rdy) * (hm ? m[3] : 1.0f);
  const float pz = (alphadt * rdz) * (hm ? m[5] : 1.0f);
  float *F = a->f;
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++) {
    if (x <= nx) FF(x, y, z, F_EX) += px * (FF(x + 1, y, z, F_DIVE) - FF(x, y, z, F_DIVE));
    if (y <= ny) FF(x, y, z, F_EY) += py * (FF(x, y + 1, z, F_DIVE) - FF(x, y, z, F_DIVE));
    if (z <= nz) FF(x, y, z, F_EZ) += pz * (FF(x, y, z + 1, F_DIVE) - FF(x, y, z, F_DIVE));
  }
# undef FF
  adjust_tang_e(a);
}

void vpo_compute_div_b_err(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const float px = (nx > 1) ? a->rdx : 0, py = (ny > 1) ? a->rdy : 0, pz = (nz > 1) ? a->rdz : 0;
  float *F = a->f;
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz; z++) for (int y = 1; y <= ny; y++) for (int x = 1; x <= nx; x++)
    FF(x, y, z, F_DIVB) = px * (FF(x + 1, y, z, F_CBX) - FF(x, y, z, F_CBX)) +
                          py * (FF(x, y + 1, z, F_CBY) - FF(x, y, z, F_CBY)) +
                          pz * (FF(x, y, z + 1, F_CBZ) - FF(x, y, z, F_CBZ));
# undef FF
}

double vpo_compute_rms_div_b_err(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const float *F = a->f;
  double err = 0;
  for (int z = 1; z <= nz; z++) for (int y = 1; y <= ny; y++) for (int x = 1; x <= nx; x++) {
    float e = F[(size_t)VOX(x, y, z) * F_STRIDE + F_DIVB];
    err += e * e;
  }
  double l0 = err * a->dV, l1 = (nx * ny * nz) * a->dV;
  return a->eps0 * sqrt(l0 / l1);
}

/* div-B-error ghosts: begin/end_remote_ghost_div_b (remote.cc:208-282), local_ghost_div_b (local.cc:181-217) */
static void ghost_div_b(const vpo_field_args_t *a) {
  dims_t d = mkdims(a);
  float *F = a->f;
  float *msg[6] = {0};
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] < 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X];
    float *p = msg[fc] = (float *)malloc(sizeof(float) * (size_t)(1 + d.n[Y] * d.n[Z]));
    *p++ = axis_d(a, X);
    PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z], *p++ = FV(v, F_DIVB));
  }
  for (int fc = 0; fc < 6; fc++) {
    int bc = a->bc6[fc];
    if (bc >= 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    int ghost = fc < 3 ? 0 : d.n[X] + 1, in1 = fc < 3 ? d.s[X] : -d.s[X];
    if (bc == -1)                  PLANE_LOOP(X, ghost, Y, 1, d.n[Y], Z, 1, d.n[Z], FV(v, F_DIVB) = FV(v + in1, F_DIVB));
    else if (bc == -2 || bc == -3) PLANE_LOOP(X, ghost, Y, 1, d.n[Y], Z, 1, d.n[Z], FV(v, F_DIVB) = -FV(v + in1, F_DIVB));
    else if (bc == -4)             PLANE_LOOP(X, ghost, Y, 1, d.n[Y], Z, 1, d.n[Z], FV(v, F_DIVB) = 0);
    else abort();
  }
  for (int fc = 0; fc < 6; fc++) {
    if (!msg[fc]) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3;
    int ghost = fc < 3 ? d.n[X] + 1 : 0, step = fc < 3 ? -d.s[X] : d.s[X];
    float *p = msg[fc];
    float lw = *p++, dX = axis_d(a, X);
    float rw = (2. * dX) / (lw + dX);
    lw = (lw - dX) / (lw + dX);
    PLANE_LOOP(X, ghost, Y, 1, d.n[Y], Z, 1, d.n[Z], FV(v, F_DIVB) = rw * (*p++) + lw * FV(v + step, F_DIVB));
    free(msg[fc]);
  }
}

void vpo_clean_div_b(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  float px = (nx > 1) ? a->rdx : 0, py = (ny > 1) ? a->rdy : 0, pz = (nz > 1) ? a->rdz : 0;
  float alphadt = 0.3888889 / (px * px + py * py + pz * pz);
  px *= alphadt; py *= alphadt; pz *= alphadt;
  float *F = a->f;
  ghost_div_b(a);
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++) {
    if (y <= ny && z <= nz) FF(x, y, z, F_CBX) += px * (FF(x, y, z, F_DIVB) - FF(x - 1, y, z, F_DIVB));
    if (z <= nz && x <= nx) FF(x, y, z, F_CBY) += py * (FF(x, y, z, F_DIVB) - FF(x, y - 1, z, F_DIVB));
    if (x <= nx && y <= ny) FF(x, y, z, F_CBZ) += pz * (FF(x, y, z, F_DIVB) - FF(x, y, z - 1, F_DIVB));
  }
# undef FF
  adjust_norm_b(a);
}

double vpo_synchronize_tang_e_norm_b(const vpo_field_args_t *a) {
  dims_t d = mkdims(a);
  float *F = a->f;
  double err = 0;
  adjust_tang_e(a);
  adjust_norm_b(a);
  for (int X = 0; X < 3; X++) {
    int Y = (X + 1) % 3, Z = (X + 2) % 3;
    float *msg[2] = {0, 0};
    for (int side = 0; side < 2; side++) {
      if (a->bc6[X + 3 * side] < 0) continue;
      int face = side == 0 ? 1 : d.n[X] + 1;
      float *p = msg[side] = (float *)malloc(sizeof(float) *
          (size_t)(2 * d.n[Y] * (d.n[Z] + 1) + 2 * d.n[Z] * (d.n[Y] + 1) + d.n[Y] * d.n[Z]));
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z], *p++ = FV(v, F_CBX + X));
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1, (*p++ = FV(v, F_EX + Y), *p++ = FV(v, F_TCAX + Y)));
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z], (*p++ = FV(v, F_EX + Z), *p++ = FV(v, F_TCAX + Z)));
    }
    /* the reference receives the -X message first, then +X (remote.cc:376-383); the error sum follows that order */
    for (int side = 0; side < 2; side++) {
      if (!msg[side]) continue;
      float *p = msg[side];
      int face = side == 0 ? d.n[X] + 1 : 1;
      double w1, w2;
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z],
                 (w1 = *p++, w2 = FV(v, F_CBX + X), FV(v, F_CBX + X) = 0.5 * (w1 + w2), err += (w1 - w2) * (w1 - w2)));
      PLANE_LOOP(X, face, Y, 1, d.n[Y], Z, 1, d.n[Z] + 1,
                 (w1 = *p++, w2 = FV(v, F_EX + Y), FV(v, F_EX + Y) = 0.5 * (w1 + w2), err += (w1 - w2) * (w1 - w2),
                  w1 = *p++, w2 = FV(v, F_TCAX + Y), FV(v, F_TCAX + Y) = 0.5 * (w1 + w2)));
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z],
                 (w1 = *p++, w2 = FV(v, F_EX + Z), FV(v, F_EX + Z) = 0.5 * (w1 + w2), err += (w1 - w2) * (w1 - w2),
                  w1 = *p++, w2 = FV(v, F_TCAX + Z), FV(v, F_TCAX + Z) = 0.5 * (w1 + w2)));
      free(msg[side]);
    }
  }
  return err;
}

void vpo_vacuum_compute_rhob(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const int hm = a->has_material;
  const float *m = a->material;
  const float nc = hm ? m[9] : 1.0f;
  const float px = (nx > 1) ? a->eps0 * (hm ? m[10] : 1.0f) * a->rdx : 0;
  const float py = (ny > 1) ? a->eps0 * (hm ? m[11] : 1.0f) * a->rdy : 0;
  const float pz = (nz > 1) ? a->eps0 * (hm ? m[12] : 1.0f) * a->rdz : 0;
  float *F = a->f;
  ghost_norm_e(a);
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++)
    FF(x, y, z, F_RHOB) = nc * (px * (FF(x, y, z, F_EX) - FF(x - 1, y, z, F_EX)) +
                                py * (FF(x, y, z, F_EY) - FF(x, y - 1, z, F_EY)) +
                                pz * (FF(x, y, z, F_EZ) - FF(x, y, z - 1, F_EZ)) - FF(x, y, z, F_RHOF));
# undef FF
  dims_t d = mkdims(a);                                              /* local_adjust_rhob */
  for (int fc = 0; fc < 6; fc++) {
    if (a->bc6[fc] != -1) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, FV(v, F_RHOB) = 0);
  }
}

void vpo_vacuum_compute_curl_b(const vpo_field_args_t *a) {
  const int nx = a->nx, ny = a->ny, nz = a->nz;
  const int hm = a->has_material;
  const float *m = a->material;
  const float rmux = hm ? m[6] : 1, rmuy = hm ? m[7] : 1, rmuz = hm ? m[8] : 1;
  const float px_muz = ((nx > 1) ? a->cvac * a->dt * a->rdx : 0) * rmuz, px_muy = ((nx > 1) ? a->cvac * a->dt * a->rdx : 0) * rmuy;
  const float py_mux = ((ny > 1) ? a->cvac * a->dt * a->rdy : 0) * rmux, py_muz = ((ny > 1) ? a->cvac * a->dt * a->rdy : 0) * rmuz;
  const float pz_muy = ((nz > 1) ? a->cvac * a->dt * a->rdz : 0) * rmuy, pz_mux = ((nz > 1) ? a->cvac * a->dt * a->rdz : 0) * rmux;
  float *F = a->f;
  ghost_tang_b(a);
# define FF(x, y, z, m) F[(size_t)VOX(x, y, z) * F_STRIDE + (m)]
  for (int z = 1; z <= nz + 1; z++) for (int y = 1; y <= ny + 1; y++) for (int x = 1; x <= nx + 1; x++) {
    if (x <= nx) FF(x, y, z, F_TCAX) = (py_muz * (FF(x, y, z, F_CBZ) - FF(x, y - 1, z, F_CBZ)) -
                                        pz_muy * (FF(x, y, z, F_CBY) - FF(x, y, z - 1, F_CBY)));
    if (y <= ny) FF(x, y, z, F_TCAY) = (pz_mux * (FF(x, y, z, F_CBX) - FF(x, y, z - 1, F_CBX)) -
                                        px_muz * (FF(x, y, z, F_CBZ) - FF(x - 1, y, z, F_CBZ)));
    if (z <= nz) FF(x, y, z, F_TCAZ) = (px_muy * (FF(x, y, z, F_CBY) - FF(x - 1, y, z, F_CBY)) -
                                        py_mux * (FF(x, y, z, F_CBX) - FF(x, y - 1, z, F_CBX)));
  }
# undef FF
  adjust_tang_e(a);
}


/* ======================================================================== */
/* Hydro moments.                                                           */
#define H_STRIDE 16

void vpo_accumulate_hydro_p(float *hydro, const vpo_particle_t *p, int32_t np, const float *interp, int32_t istride,
                            float qsp, float msp, float qdt_2mc, float cvac, float r8V, int32_t nx, int32_t ny, int32_t nz) {
  const float qdt_4mc2 = qdt_2mc / (2 * cvac);                       /* hydro_p_pipeline.cc:33-38 */
  const float mspc = cvac * msp, c = cvac;
  const float one = 1.0, one_third = 1.0 / 3.0;
  const int sx = 1, sy = nx + 2, sz = (nx + 2) * (ny + 2);
  (void)nz;
  for (int n = 0; n < np; n++) {
    float dx = p[n].dx, dy = p[n].dy, dz = p[n].dz;
    int i = p[n].i;
    float ux = p[n].ux, uy = p[n].uy, uz = p[n].uz, w = p[n].w;
    const float *f = interp + (size_t)i * istride;
    ux += qdt_2mc * ((f[0] + dy * f[1]) + dz * (f[2] + dy * f[3]));            /* half E kick, :88-95 */
    uy += qdt_2mc * ((f[4] + dz * f[5]) + dx * (f[6] + dz * f[7]));
    uz += qdt_2mc * ((f[8] + dx * f[9]) + dy * (f[10] + dx * f[11]));
    float w5 = f[12] + dx * f[13], w6 = f[14] + dy * f[15], w7 = f[16] + dz * f[17];
    float ke_mc = ux * ux + uy * uy + uz * uz;                                   /* :112-115 */
    float vz = sqrtf(one + ke_mc);
    ke_mc *= c / (vz + one);
    vz = c / vz;
    float w0 = qdt_4mc2 * vz;                                                    /* half Boris rotation, :121-136 */
    float w1 = w5 * w5 + w6 * w6 + w7 * w7;
    float w2 = w0 * w0 * w1;
    float w3 = w0 * (one + (one_third) * w2 * (one + 0.4f * w2));
    float w4 = w3 / (one + w1 * w3 * w3);
    w4 += w4;
    w0 = ux + w3 * (uy * w7 - uz * w6);
    w1 = uy + w3 * (uz * w5 - ux * w7);
    w2 = uz + w3 * (ux * w6 - uy * w5);
    ux += w4 * (w1 * w7 - w2 * w6);
    uy += w4 * (w2 * w5 - w0 * w7);
    uz += w4 * (w0 * w6 - w1 * w5);
    float vx = ux * vz, vy = uy * vz;
    vz = uz * vz;
    w0 = r8V * w;                                                                /* trilinear weights, :152-172 */
    dx *= w0; w1 = w0 + dx; w0 -= dx;
    w3 = one + dy; w2 = w0 * w3; w3 *= w1;
    dy = one - dy; w0 *= dy; w1 *= dy;
    w7 = one + dz; w4 = w0 * w7; w5 = w1 * w7; w6 = w2 * w7; w7 *= w3;
    dz = one - dz; w0 *= dz; w1 *= dz; w2 *= dz; w3 *= dz;
    const float wn[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
    const int node[8] = {i, i + sx, i + sy, i + sy + sx, i + sz, i + sz + sx, i + sz + sy, i + sz + sy + sx};
    for (int k = 0; k < 8; k++) {                                                /* ACCUM_HYDRO, :178-198 */
      float *h = hydro + (size_t)node[k] * H_STRIDE;
      float t = qsp * wn[k];
      h[0] += t * vx; h[1] += t * vy; h[2] += t * vz; h[3] += t;
      t = mspc * wn[k];
      float tx = t * ux, ty = t * uy, tz = t * uz;
      h[4] += tx; h[5] += ty; h[6] += tz; h[7] += t * ke_mc;
      h[8] += tx * vx; h[9] += ty * vy; h[10] += tz * vz;
      h[11] += ty * vz; h[12] += tz * vx; h[13] += tx * vy;
    }
  }
}

void vpo_synchronize_hydro(float *H, const vpo_field_args_t *a) {
  dims_t d = mkdims(a);
  for (int fc = 0; fc < 6; fc++) {                                   /* ADJUST_HYDRO: every local wall doubles its nodes */
    if (a->bc6[fc] >= 0) continue;
    int X = fc % 3, Y = (X + 1) % 3, Z = (X + 2) % 3, face = fc < 3 ? 1 : d.n[X] + 1;
    PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, { for (int k = 0; k < 14; k++) H[(size_t)v * H_STRIDE + k] *= 2; });
  }
  for (int X = 0; X < 3; X++) {
    int Y = (X + 1) % 3, Z = (X + 2) % 3;
    float *msg[2] = {0, 0};
    for (int side = 0; side < 2; side++) {
      if (a->bc6[X + 3 * side] < 0) continue;
      int face = side == 0 ? 1 : d.n[X] + 1;
      float *p = msg[side] = (float *)malloc(sizeof(float) * (size_t)(1 + 14 * (d.n[Y] + 1) * (d.n[Z] + 1)));
      *p++ = axis_d(a, X);
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1, { for (int k = 0; k < 14; k++) *p++ = H[(size_t)v * H_STRIDE + k]; });
    }
    for (int side = 0; side < 2; side++) {
      if (!msg[side]) continue;
      float *p = msg[side];
      float rw = *p++, dX = axis_d(a, X), lw = rw + dX;
      rw /= lw; lw = dX / lw; lw += lw; rw += rw;
      int face = side == 0 ? d.n[X] + 1 : 1;
      PLANE_LOOP(X, face, Y, 1, d.n[Y] + 1, Z, 1, d.n[Z] + 1,
                 { for (int k = 0; k < 14; k++) { float *h = H + (size_t)v * H_STRIDE + k; *h = lw * (*h) + rw * (*p++); } });
      free(msg[side]);
    }
  }
}
