/*
 * rbc3d_oracle_walls.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See rbc3d_oracle.h.
 *
 * CPU restatement of the wall part of RBC3D's Ewald boundary-integral operator:
 * ModIntOnWalls.F90 (AddIntOnWalls, PrepareSingIntOnWall, SingIntOnWall, Tri_Int_Regular,
 * Tri_Int_Duffy, MinDistToTri), the wall branches of ModSourceList.F90:127-146 and
 * ModPME.F90:105-129, Wall_ComputeGeometry (ModWall.F90:118-145) and gqTri7
 * (ModQuadRule.F90:70-94).  The PETSc SeqAIJ matrix of t_Wall%lhs is restated as a block-row
 * sparse matrix (one 3x3 block per (target vertex, source vertex)); MatMult sums a row in
 * ascending column order, which is reproduced (component-major, then vertex).
 * PARITY UNPINNED (no reference tests or runnable reference exist; see header).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rbc3d_oracle.h"

#define THRD (1.0 / 3)

static inline double wnint(double x) { return round(x); }

/* ---------------------------------------------------------------------- */
/* ModQuadRule.F90:70-94: 7-point Gauss rule on the reference triangle */
static void gq_tri7(double rs[7][2], double w[7]) {
  double r = (6. - sqrt(15.)) / 21., s = r, t = 1 - r - s;
  rs[0][0] = r, rs[1][0] = s, rs[2][0] = t;
  rs[0][1] = s, rs[1][1] = t, rs[2][1] = r;
  w[0] = w[1] = w[2] = (155. - sqrt(15.)) / 2400.;
  r = (6. + sqrt(15.)) / 21., s = r, t = 1 - r - s;
  rs[3][0] = r, rs[4][0] = s, rs[5][0] = t;
  rs[3][1] = s, rs[4][1] = t, rs[5][1] = r;
  w[3] = w[4] = w[5] = (155. + sqrt(15.)) / 2400.;
  r = 1. / 3., s = 1. / 3.;
  rs[6][0] = r, rs[6][1] = s;
  w[6] = 9. / 80.;
}
void orc_gq_tri7(double *rs /* [7][2] */, double *w) {
  double a[7][2];
  gq_tri7(a, w);
  memcpy(rs, a, sizeof(a));
}

/* ModBasicMath.F90:132-145 TriArea; x[l*3+d] = x(l+1, d+1) */
static double tri_area(const double *x) {
  double a[3], b[3], c[3];
  for (int d = 0; d < 3; d++) {
    a[d] = x[3 + d] - x[d];
    b[d] = x[6 + d] - x[d];
  }
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  return 0.5 * sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
}

/* ModWall.F90:118-145 Wall_ComputeGeometry: area and epsDist of every element */
void orc_wall_compute_geometry(const orc_walls *W, double *area, double *epsDist) {
  int NV = 0, NE = 0;
  for (int w = 0; w < W->nwall; w++) NV += W->nvert[w], NE += W->nele[w];
  int v0 = 0, e0 = 0;
  for (int w = 0; w < W->nwall; w++) {
    for (int e = 0; e < W->nele[w]; e++) {
      double xe[9];
      for (int l = 0; l < 3; l++) {
        int iv = v0 + W->e2v[(size_t)l * NE + e0 + e] - 1;
        for (int d = 0; d < 3; d++) xe[3 * l + d] = W->x[(size_t)d * NV + iv];
      }
      double x12[3], x13[3], a3[3];
      for (int d = 0; d < 3; d++) x12[d] = xe[3 + d] - xe[d], x13[d] = xe[6 + d] - xe[d];
      a3[0] = x12[1] * x13[2] - x12[2] * x13[1];
      a3[1] = x12[2] * x13[0] - x12[0] * x13[2];
      a3[2] = x12[0] * x13[1] - x12[1] * x13[0];
      double a3N = sqrt(a3[0] * a3[0] + a3[1] * a3[1] + a3[2] * a3[2]);
      area[e0 + e] = 0.5 * a3N;
      epsDist[e0 + e] = sqrt(area[e0 + e]);
    }
    v0 += W->nvert[w];
    e0 += W->nele[w];
  }
}

/* ModIntOnWalls.F90:480-577 MinDistToTri (Eberly).  Returns the distance; s0,t0,x0 may be NULL. */
double orc_min_dist_to_tri(const double xTar[3], const double *x, double *s0, double *t0, double *x0) {
  double x12[3], x13[3], x1Tar[3];
  for (int k = 0; k < 3; k++) {
    x12[k] = x[3 + k] - x[k];
    x13[k] = x[6 + k] - x[k];
  }
  double a = x12[0] * x12[0] + x12[1] * x12[1] + x12[2] * x12[2];
  double b = x12[0] * x13[0] + x12[1] * x13[1] + x12[2] * x13[2];
  double c = x13[0] * x13[0] + x13[1] * x13[1] + x13[2] * x13[2];
  double det = a * c - b * b;
  double invDet = 1. / det;
  for (int k = 0; k < 3; k++) x1Tar[k] = x[k] - xTar[k];
  double d = x12[0] * x1Tar[0] + x12[1] * x1Tar[1] + x12[2] * x1Tar[2];
  double e = x13[0] * x1Tar[0] + x13[1] * x1Tar[1] + x13[2] * x1Tar[2];
  double f = x1Tar[0] * x1Tar[0] + x1Tar[1] * x1Tar[1] + x1Tar[2] * x1Tar[2];
  double s = b * e - c * d;
  double t = b * d - a * e;
  int region;
  if (s + t <= det) {
    if (s < 0)
      region = (t < 0) ? 4 : 3;
    else if (t < 0)
      region = 5;
    else
      region = 0;
  } else {
    if (s < 0)
      region = 2;
    else if (t < 0)
      region = 6;
    else
      region = 1;
  }
  if (region == 2)
    region = (-(c + e) < 0) ? 3 : 1;
  else if (region == 4)
    region = (d < 0) ? 5 : 3;
  else if (region == 6)
    region = (b + e - a - d < 0) ? 1 : 5;
  switch (region) {
    case 0:
      s = invDet * s;
      t = invDet * t;
      break;
    case 1:
      s = (c + e - b - d) / (a - 2 * b + c);
      s = fmin(1., fmax(0., s));
      t = 1. - s;
      break;
    case 3:
      s = 0.;
      t = -e / c;
      t = fmin(1., fmax(0., t));
      break;
    case 5:
      t = 0.;
      s = -d / a;
      s = fmin(1., fmax(0., s));
      break;
  }
  double dist = sqrt(a * s * s + 2 * b * s * t + c * t * t + 2 * d * s + 2 * e * t + f);
  if (s0) *s0 = s;
  if (t0) *t0 = t;
  if (x0)
    for (int k = 0; k < 3; k++) x0[k] = (1 - s - t) * x[k] + s * x[3 + k] + t * x[6 + k];
  return dist;
}

/* shared tail of both quadratures (ModIntOnWalls.F90:344-360, 424-458) */
static inline void tri_accum(const orc_params *prm, const double xTar[3], const double xGq[3], double fGq[3],
                             double dsGq, double wl0, double wl1, double wl2, double rhs[3], double *lhs) {
  double xx[3], EA, EB;
  for (int k = 0; k < 3; k++) fGq[k] = dsGq * fGq[k];
  for (int k = 0; k < 3; k++) xx[k] = xTar[k] - xGq[k];
  double rr = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]);
  orc_ewald_coeff_sl(prm, rr, &EA, &EB);
  double dot = xx[0] * fGq[0] + xx[1] * fGq[1] + xx[2] * fGq[2];
  for (int k = 0; k < 3; k++) rhs[k] = rhs[k] + (EA * xx[k] * dot + EB * fGq[k]);
  if (lhs) {
    double tmp[3][3];
    for (int ii = 0; ii < 3; ii++)
      for (int jj = 0; jj < 3; jj++) tmp[ii][jj] = EA * xx[ii] * xx[jj];
    for (int ii = 0; ii < 3; ii++) tmp[ii][ii] = tmp[ii][ii] + EB;
    for (int ii = 0; ii < 3; ii++)
      for (int jj = 0; jj < 3; jj++) {
        tmp[ii][jj] = dsGq * tmp[ii][jj];
        lhs[0 * 9 + ii * 3 + jj] += wl0 * tmp[ii][jj];
        lhs[1 * 9 + ii * 3 + jj] += wl1 * tmp[ii][jj];
        lhs[2 * 9 + ii * 3 + jj] += wl2 * tmp[ii][jj];
      }
  }
}

/* ModIntOnWalls.F90:319-363 Tri_Int_Regular.  x,f: [l*3+d]; lhs[(l*3+ii)*3+jj] = lhs(l,ii,jj) or NULL */
void orc_tri_int_regular(const orc_params *prm, const double *x, const double *f, const double xtar[3],
                         double rhs[3], double *lhs) {
  double rs[7][2], w[7];
  gq_tri7(rs, w);
  double detJ = 2 * tri_area(x);
  rhs[0] = rhs[1] = rhs[2] = 0.;
  if (lhs) memset(lhs, 0, sizeof(double) * 27);
  for (int q = 0; q < 7; q++) {
    double s = rs[q][0], t = rs[q][1], xGq[3], fGq[3];
    for (int k = 0; k < 3; k++) {
      xGq[k] = (1. - s - t) * x[k] + s * x[3 + k] + t * x[6 + k];
      fGq[k] = (1. - s - t) * f[k] + s * f[3 + k] + t * f[6 + k];
    }
    double dsGq = w[q] * detJ;
    tri_accum(prm, xtar, xGq, fGq, dsGq, 1 - s - t, s, t, rhs, lhs);
  }
}

/* ModIntOnWalls.F90:373-465 Tri_Int_Duffy: 3 sub-triangles around (s0,t0), 4x4 Gauss-Legendre each */
void orc_tri_int_duffy(const orc_params *prm, const double *x, const double *f, const double xtar[3], double s0,
                       double t0, double rhs[3], double *lhs) {
  enum { nGq = 4 };
  double rGq[nGq], wGq[nGq];
  orc_gauleg(0., 1., nGq, rGq, wGq);
  double x0[3], f0[3];
  for (int k = 0; k < 3; k++) {
    x0[k] = (1 - s0 - t0) * x[k] + s0 * x[3 + k] + t0 * x[6 + k];
    f0[k] = (1 - s0 - t0) * f[k] + s0 * f[3 + k] + t0 * f[6 + k];
  }
  rhs[0] = rhs[1] = rhs[2] = 0.;
  if (lhs) memset(lhs, 0, sizeof(double) * 27);
  for (int n = 0; n < 3; n++) {
    const double *x1 = x + 3 * n, *x2 = x + 3 * ((n + 1) % 3);
    const double *f1 = f + 3 * n, *f2 = f + 3 * ((n + 1) % 3);
    double u[3], v[3], nr[3];
    for (int k = 0; k < 3; k++) u[k] = x1[k] - x0[k], v[k] = x2[k] - x0[k];
    nr[0] = u[1] * v[2] - u[2] * v[1];
    nr[1] = u[2] * v[0] - u[0] * v[2];
    nr[2] = u[0] * v[1] - u[1] * v[0];
    double detJ = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
    double s1, t1, s2, t2;
    if (n == 0) {
      s1 = 0., t1 = 0., s2 = 1., t2 = 0.;
    } else if (n == 1) {
      s1 = 1., t1 = 0., s2 = 0., t2 = 1.;
    } else {
      s1 = 0., t1 = 1., s2 = 0., t2 = 0.;
    }
    for (int i = 0; i < nGq; i++)
      for (int j = 0; j < nGq; j++) {
        double s = rGq[i], t = s * rGq[j], xGq[3], fGq[3];
        for (int k = 0; k < 3; k++) {
          xGq[k] = (1. - s) * x0[k] + (s - t) * x1[k] + t * x2[k];
          fGq[k] = (1. - s) * f0[k] + (s - t) * f1[k] + t * f2[k];
        }
        double dsGq = wGq[i] * wGq[j] * detJ * s;
        double sGlb = (1. - s) * s0 + (s - t) * s1 + t * s2;
        double tGlb = (1. - s) * t0 + (s - t) * t1 + t * t2;
        tri_accum(prm, xtar, xGq, fGq, dsGq, 1. - sGlb - tGlb, sGlb, tGlb, rhs, lhs);
      }
  }
}

/* ---------------------------------------------------------------------- */
typedef struct {
  int NV, NE;
  int *voff, *eoff;   /* [nwall+1] */
  int *ewall;         /* [NE] wall index of an element */
  double *xc;         /* SoA(3,NE) centroids = slist_wall%x (ModSourceList.F90:131-143) */
  int *hoc, *next;
} wall_lists;

static void wall_lists_init(const orc_params *prm, const orc_walls *W, wall_lists *L) {
  L->voff = (int *)calloc(W->nwall + 1, sizeof(int));
  L->eoff = (int *)calloc(W->nwall + 1, sizeof(int));
  for (int w = 0; w < W->nwall; w++) {
    L->voff[w + 1] = L->voff[w] + W->nvert[w];
    L->eoff[w + 1] = L->eoff[w] + W->nele[w];
  }
  const int NV = L->NV = L->voff[W->nwall], NE = L->NE = L->eoff[W->nwall];
  L->ewall = (int *)malloc(sizeof(int) * (NE > 0 ? NE : 1));
  L->xc = (double *)malloc(sizeof(double) * 3 * (NE > 0 ? NE : 1));
  for (int w = 0; w < W->nwall; w++)
    for (int e = L->eoff[w]; e < L->eoff[w + 1]; e++) {
      L->ewall[e] = w;
      for (int d = 0; d < 3; d++) {
        double xe[3];
        for (int l = 0; l < 3; l++) xe[l] = W->x[(size_t)d * NV + L->voff[w] + W->e2v[(size_t)l * NE + e] - 1];
        L->xc[(size_t)d * NE + e] = THRD * (xe[0] + xe[1] + xe[2]); /* THRD*sum(xele,dim=1) */
      }
    }
  const int n1 = prm->Nc[0] + 2, n2 = prm->Nc[1] + 2, n3 = prm->Nc[2] + 2;
  L->hoc = (int *)malloc(sizeof(int) * (size_t)n1 * n2 * n3);
  L->next = (int *)malloc(sizeof(int) * (NE > 0 ? NE : 1));
  orc_hash_build(prm, NE, L->xc, L->hoc, L->next);
}
static void wall_lists_free(wall_lists *L) {
  free(L->voff), free(L->eoff), free(L->ewall), free(L->xc), free(L->hoc), free(L->next);
}

void orc_wall_centroids(const orc_params *prm, const orc_walls *W, double *xc) {
  wall_lists L;
  wall_lists_init(prm, W, &L);
  memcpy(xc, L.xc, sizeof(double) * 3 * L.NE);
  wall_lists_free(&L);
}

/* element e (global), translated close to xi (ModIntOnWalls.F90:96-107) */
static inline void load_element(const orc_params *prm, const orc_walls *W, const wall_lists *L, int e,
                                const double xi[3], double *xele, double *fele, int *ivert) {
  const int w = L->ewall[e], NV = L->NV, NE = L->NE;
  for (int l = 0; l < 3; l++) {
    int iv = L->voff[w] + W->e2v[(size_t)l * NE + e] - 1;
    if (ivert) ivert[l] = iv;
    for (int d = 0; d < 3; d++) {
      xele[3 * l + d] = W->x[(size_t)d * NV + iv];
      if (fele) fele[3 * l + d] = W->f ? W->f[(size_t)d * NV + iv] : 0.;
    }
  }
  double xx[3];
  for (int d = 0; d < 3; d++) xx[d] = wnint((xi[d] - xele[d]) * prm->iLb[d]) * prm->Lb[d];
  for (int l = 0; l < 3; l++)
    for (int d = 0; d < 3; d++) xele[3 * l + d] = xele[3 * l + d] + xx[d];
}

#define HOC3(i1, i2, i3) L.hoc[(i1) + (size_t)n1 * ((i2) + (size_t)n2 * (i3))]
#define FOR_NEIGHBOUR_CELLS                                                                              \
  for (int j1 = (i1 - 1 > 0 ? i1 - 1 : 0); j1 <= (i1 + 1 < Nc[0] + 1 ? i1 + 1 : Nc[0] + 1); j1++)       \
    for (int j2 = (i2 - 1 > 0 ? i2 - 1 : 0); j2 <= (i2 + 1 < Nc[1] + 1 ? i2 + 1 : Nc[1] + 1); j2++)     \
      for (int j3 = (i3 - 1 > 0 ? i3 - 1 : 0); j3 <= (i3 + 1 < Nc[2] + 1 ? i3 + 1 : Nc[2] + 1); j3++)

/* ---------------------------------------------------------------------- */
/* t_Wall%lhs (PETSc SeqAIJ, 3nvert x 3nvert) as block rows */
struct orc_wallmat {
  int nvert;
  int *rowptr; /* [nvert+1] */
  int *col;    /* [nblk] source vertex (0-based, local to the wall), ascending inside a row */
  double *val; /* [nblk][3][3]: val[ii*3+jj] couples target component ii with source component jj */
};

typedef struct {
  int col;
  int seq;
  double v[9];
} blk_entry;
static int blk_cmp(const void *a, const void *b) {
  const blk_entry *p = (const blk_entry *)a, *q = (const blk_entry *)b;
  if (p->col != q->col) return p->col < q->col ? -1 : 1;
  return p->seq < q->seq ? -1 : (p->seq > q->seq);
}

/* ModIntOnWalls.F90:181-308 PrepareSingIntOnWall.  active: [nvert of wall iwall] or NULL (all) */
orc_wallmat *orc_prepare_sing_int_on_wall(const orc_params *prm, const orc_walls *W, int iwall, const int *active) {
  wall_lists L;
  wall_lists_init(prm, W, &L);
  const int *Nc = prm->Nc;
  const int n1 = Nc[0] + 2, n2 = Nc[1] + 2;
  const int nvert = W->nvert[iwall], NV = L.NV;
  orc_wallmat *M = (orc_wallmat *)calloc(1, sizeof(orc_wallmat));
  M->nvert = nvert;
  M->rowptr = (int *)calloc(nvert + 1, sizeof(int));
  blk_entry **rows = (blk_entry **)calloc(nvert, sizeof(blk_entry *));
  int *nrow = (int *)calloc(nvert, sizeof(int));
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < nvert; i++) {
    if (active && !active[i]) continue;
    double xi[3];
    for (int d = 0; d < 3; d++) xi[d] = W->x[(size_t)d * NV + L.voff[iwall] + i];
    int i1, i2, i3, cap = 64, n = 0;
    blk_entry *ent = (blk_entry *)malloc(sizeof(blk_entry) * cap);
    orc_hash_index(prm, xi, &i1, &i2, &i3);
    FOR_NEIGHBOUR_CELLS {
      for (int j = HOC3(j1, j2, j3); j >= 0; j = L.next[j]) {
        if (L.ewall[j] != iwall) continue; /* :263 */
        double xele[9], fele[9], s0, t0, rhs[3], lhs[27];
        int ivert[3];
        load_element(prm, W, &L, j, xi, xele, NULL, ivert);
        memset(fele, 0, sizeof(fele)); /* dummy, :276 */
        double rr = orc_min_dist_to_tri(xi, xele, &s0, &t0, NULL);
        if (rr > prm->rc) continue;
        if (rr > W->epsDist[j])
          orc_tri_int_regular(prm, xele, fele, xi, rhs, lhs);
        else
          orc_tri_int_duffy(prm, xele, fele, xi, s0, t0, rhs, lhs);
        for (int l = 0; l < 3; l++) {
          if (n == cap) ent = (blk_entry *)realloc(ent, sizeof(blk_entry) * (cap *= 2));
          ent[n].col = ivert[l] - L.voff[iwall];
          ent[n].seq = n;
          memcpy(ent[n].v, lhs + 9 * l, sizeof(double) * 9);
          n++;
        }
      }
    }
    /* MatSetValues(..., ADD_VALUES): duplicates are summed in insertion order */
    qsort(ent, n, sizeof(blk_entry), blk_cmp);
    int m = 0;
    for (int k = 0; k < n; k++) {
      if (m > 0 && ent[m - 1].col == ent[k].col) {
        for (int q = 0; q < 9; q++) ent[m - 1].v[q] += ent[k].v[q];
      } else {
        if (m != k) ent[m] = ent[k];
        m++;
      }
    }
    rows[i] = ent;
    nrow[i] = m;
  }
  for (int i = 0; i < nvert; i++) M->rowptr[i + 1] = M->rowptr[i] + nrow[i];
  const int nblk = M->rowptr[nvert];
  M->col = (int *)malloc(sizeof(int) * (nblk > 0 ? nblk : 1));
  M->val = (double *)malloc(sizeof(double) * 9 * (nblk > 0 ? nblk : 1));
  for (int i = 0; i < nvert; i++) {
    for (int k = 0; k < nrow[i]; k++) {
      M->col[M->rowptr[i] + k] = rows[i][k].col;
      memcpy(M->val + 9 * (size_t)(M->rowptr[i] + k), rows[i][k].v, sizeof(double) * 9);
    }
    free(rows[i]);
  }
  free(rows), free(nrow);
  wall_lists_free(&L);
  return M;
}

void orc_wallmat_free(orc_wallmat *M) {
  if (!M) return;
  free(M->rowptr), free(M->col), free(M->val), free(M);
}
int orc_wallmat_nblk(const orc_wallmat *M) { return M->rowptr[M->nvert]; }
void orc_wallmat_get(const orc_wallmat *M, int *rowptr, int *col, double *val) {
  memcpy(rowptr, M->rowptr, sizeof(int) * (M->nvert + 1));
  memcpy(col, M->col, sizeof(int) * M->rowptr[M->nvert]);
  memcpy(val, M->val, sizeof(double) * 9 * M->rowptr[M->nvert]);
}

/* ModIntOnWalls.F90:136-172 SingIntOnWall: v = c1 * lhs * f; f, v SoA(3,nvert) of that wall */
void orc_sing_int_on_wall(const orc_wallmat *M, double c1, const double *f, double *v) {
  const int nv = M->nvert;
#pragma omp parallel for
  for (int i = 0; i < nv; i++)
    for (int ii = 0; ii < 3; ii++) {
      double s = 0.;
      for (int jj = 0; jj < 3; jj++) /* AIJ row: columns ascending = component-major */
        for (int k = M->rowptr[i]; k < M->rowptr[i + 1]; k++)
          s += M->val[9 * (size_t)k + ii * 3 + jj] * f[(size_t)jj * nv + M->col[k]];
      v[(size_t)ii * nv + i] = c1 * s;
    }
}

/* ModIntOnWalls.F90:33-130 AddIntOnWalls.  mats: [nwall] self matrices (needed when tl is the wall list) */
void orc_add_int_on_walls(const orc_params *prm, const orc_walls *W, double c1, const orc_targets *tl, double *v,
                          orc_wallmat *const *mats) {
  if (W->nwall == 0) return;
  wall_lists L;
  wall_lists_init(prm, W, &L);
  const size_t nt = tl->n;
  const int NV = L.NV;
  const int first_surfId = tl->indx[0];
  if (first_surfId == W->id0) { /* self-interactions, :54-77 */
    for (int w = 0; w < W->nwall; w++) {
      const int nv = W->nvert[w], p = L.voff[w];
      double *fw = (double *)malloc(sizeof(double) * 3 * nv), *vt = (double *)malloc(sizeof(double) * 3 * nv);
      for (int d = 0; d < 3; d++)
        for (int i = 0; i < nv; i++) fw[(size_t)d * nv + i] = W->f[(size_t)d * NV + p + i];
      orc_sing_int_on_wall(mats[w], c1, fw, vt);
      for (int d = 0; d < 3; d++)
        for (int i = 0; i < nv; i++) v[d * nt + p + i] = v[d * nt + p + i] + vt[(size_t)d * nv + i] / tl->Acoef[p + i];
      free(fw), free(vt);
    }
    if (W->nwall == 1) {
      wall_lists_free(&L);
      return;
    }
  }
  const int *Nc = prm->Nc;
  const int n1 = Nc[0] + 2, n2 = Nc[1] + 2;
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t i = 0; i < nt; i++) {
    if (!tl->active[i]) continue;
    double xi[3] = {tl->x[i], tl->x[nt + i], tl->x[2 * nt + i]};
    int i1, i2, i3;
    orc_hash_index(prm, xi, &i1, &i2, &i3);
    FOR_NEIGHBOUR_CELLS {
      for (int j = HOC3(j1, j2, j3); j >= 0; j = L.next[j]) {
        if (tl->indx[i] == W->id0 + L.ewall[j]) continue; /* :92 */
        double xele[9], fele[9], s0, t0, dv[3];
        load_element(prm, W, &L, j, xi, xele, fele, NULL);
        double rr = orc_min_dist_to_tri(xi, xele, &s0, &t0, NULL);
        if (rr > prm->rc) continue;
        if (rr < W->epsDist[j])
          orc_tri_int_duffy(prm, xele, fele, xi, s0, t0, dv, NULL);
        else
          orc_tri_int_regular(prm, xele, fele, xi, dv, NULL);
        for (int d = 0; d < 3; d++) v[d * nt + i] = v[d * nt + i] + c1 * dv[d] / tl->Acoef[i];
      }
    }
  }
  wall_lists_free(&L);
}

/* in-range (target, element) sets: count and order-independent checksum per target (tests: the GPU
 * neighbour indexing must be bit-exact).  self_skip = 1 applies the same-surface exclusion of :92 */
void orc_wall_neighbor_signature(const orc_params *prm, const orc_walls *W, const orc_targets *tl, int self_skip,
                                 int *count, unsigned long long *sig, int *nduffy) {
  wall_lists L;
  wall_lists_init(prm, W, &L);
  const size_t nt = tl->n;
  const int *Nc = prm->Nc;
  const int n1 = Nc[0] + 2, n2 = Nc[1] + 2;
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t i = 0; i < nt; i++) {
    count[i] = 0;
    sig[i] = 0;
    if (nduffy) nduffy[i] = 0;
    if (!tl->active[i]) continue;
    double xi[3] = {tl->x[i], tl->x[nt + i], tl->x[2 * nt + i]};
    int i1, i2, i3;
    orc_hash_index(prm, xi, &i1, &i2, &i3);
    FOR_NEIGHBOUR_CELLS {
      for (int j = HOC3(j1, j2, j3); j >= 0; j = L.next[j]) {
        if (self_skip && tl->indx[i] == W->id0 + L.ewall[j]) continue;
        double xele[9], s0, t0;
        load_element(prm, W, &L, j, xi, xele, NULL, NULL);
        double rr = orc_min_dist_to_tri(xi, xele, &s0, &t0, NULL);
        if (rr > prm->rc) continue;
        count[i]++;
        unsigned long long z = (unsigned long long)j + 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        sig[i] += z ^ (z >> 31);
        if (nduffy && rr < W->epsDist[j]) nduffy[i]++;
      }
    }
  }
  wall_lists_free(&L);
}

/* ModPME.F90:105-129: wall sources of the PME spread = element centroid, THRD*sum(fele)*area */
void orc_pme_distrib_walls(orc_pme *pme, const orc_params *prm, double c1, double c2, const orc_walls *W,
                           int accumulate) {
  wall_lists L;
  wall_lists_init(prm, W, &L);
  const int NE = L.NE, NV = L.NV;
  double *ft = (double *)malloc(sizeof(double) * 3 * (NE > 0 ? NE : 1));
  for (int e = 0; e < NE; e++) {
    const int w = L.ewall[e];
    for (int d = 0; d < 3; d++) {
      double fe[3];
      for (int l = 0; l < 3; l++) fe[l] = W->f[(size_t)d * NV + L.voff[w] + W->e2v[(size_t)l * NE + e] - 1];
      ft[(size_t)d * NE + e] = THRD * (fe[0] + fe[1] + fe[2]) * W->area[e];
    }
  }
  /* ttmp = 0: no double-layer density on walls (:124); c2 only keeps flag_doub_lay of the same call */
  orc_pme_distrib_source(pme, c1, c2, NE, L.xc, ft, NULL, NULL, NULL, accumulate);
  free(ft);
  wall_lists_free(&L);
}

/* ====================================================================== */
/* ModRepulsion.F90: closest-neighbour queries on the cell lists of the path's own source
 * lists, and the displacement InterCellRepulsion derives from them.  Lives here because it needs both the
 * cell source list and the wall element list. */

/* ModRepulsion.F90:480-546 Closest_Neighbor_Cell for one point.  hoc/next: cell list of all cell points
 * (slist_rbc).  Returns dist0 (HUGE when no other cell has a point in the 27 neighbouring list cells). */
static double closest_neighbor_cell(const orc_params *prm, const orc_cells *C, const int *hoc, const int *next,
                                    const double xi[3], int surfId_i, double epsDist, double x0[3]) {
  const int *Nc = prm->Nc;
  const int n1 = Nc[0] + 2, n2 = Nc[1] + 2;
  const int nlat = C->nlat, nlon = C->nlon, npc = nlat * nlon;
  const size_t Np = (size_t)C->ncell * npc;
  double dist0 = HUGE_VAL;
  if (C->ncell == 0) return dist0;
  int i1, i2, i3, j0 = -1;
  orc_hash_index(prm, xi, &i1, &i2, &i3);
  FOR_NEIGHBOUR_CELLS {
    for (int j = hoc[j1 + (size_t)n1 * (j2 + (size_t)n2 * j3)]; j >= 0; j = next[j]) {
      if (j / npc + 1 == surfId_i) continue; /* :507 */
      double xx[3], rr = 0;
      for (int d = 0; d < 3; d++) {
        xx[d] = xi[d] - C->x[d * Np + j];
        xx[d] = xx[d] - wnint(xx[d] * prm->iLb[d]) * prm->Lb[d];
        rr += xx[d] * xx[d];
      }
      rr = sqrt(rr);
      if (rr < dist0) dist0 = rr, j0 = j;
    }
  }
  if (dist0 <= 2 * epsDist) { /* :525-542: refine by projecting on the neighbour's spline surface */
    const int jc = j0 / npc, ilon0 = (j0 % npc) / nlat, ilat0 = (j0 % npc) % nlat;
    double th0 = C->th[ilat0], phi0 = C->phi[ilon0], xtar[3], xx[3];
    for (int d = 0; d < 3; d++) {
      x0[d] = C->x[d * Np + j0];
      xx[d] = xi[d] - x0[d];
      xx[d] = xx[d] - wnint(xx[d] * prm->iLb[d]) * prm->Lb[d];
      xtar[d] = x0[d] + xx[d];
    }
    const size_t sp3 = (size_t)4 * 3 * nlon * 2 * nlat;
    orc_spline_find_projection(C->spx + sp3 * jc, 2 * nlat, nlon, xtar, &th0, &phi0, x0);
    dist0 = 0;
    for (int d = 0; d < 3; d++) dist0 += (xtar[d] - x0[d]) * (xtar[d] - x0[d]);
    dist0 = sqrt(dist0);
  }
  return dist0;
}

/* ModRepulsion.F90:556-613 Closest_Neighbor_Wall for one point */
static double closest_neighbor_wall(const orc_params *prm, const orc_walls *W, const wall_lists *Lp, const double xi[3],
                                    int surfId_i, double x0[3]) {
  const wall_lists L = *Lp;
  const int *Nc = prm->Nc;
  const int n1 = Nc[0] + 2, n2 = Nc[1] + 2;
  double dist0 = HUGE_VAL;
  if (W->nwall == 0) return dist0;
  int i1, i2, i3;
  orc_hash_index(prm, xi, &i1, &i2, &i3);
  FOR_NEIGHBOUR_CELLS {
    for (int j = HOC3(j1, j2, j3); j >= 0; j = L.next[j]) {
      if (W->id0 + L.ewall[j] == surfId_i) continue; /* :582 */
      const int w = L.ewall[j];
      double xele[9], xtar[3], xj[3], s0, t0;
      for (int l = 0; l < 3; l++) {
        int iv = L.voff[w] + W->e2v[(size_t)l * L.NE + j] - 1;
        for (int d = 0; d < 3; d++) xele[3 * l + d] = W->x[(size_t)d * L.NV + iv];
      }
      for (int d = 0; d < 3; d++) { /* :593-595: translate xi as close to the triangle as possible */
        double xx = xi[d] - xele[d];
        xx = xx - wnint(xx * prm->iLb[d]) * prm->Lb[d];
        xtar[d] = xele[d] + xx;
      }
      double rr = orc_min_dist_to_tri(xtar, xele, &s0, &t0, xj);
      if (rr < dist0) {
        dist0 = rr;
        for (int d = 0; d < 3; d++) x0[d] = xj[d];
      }
    }
  }
  return dist0;
}

/* Both queries for a list of points: x SoA(3,n), surfId [n]; outputs dist_cell/dist_wall [n] (HUGE_VAL =
 * none), x0_cell/x0_wall SoA(3,n) (undefined where the distance is HUGE_VAL).  W may be NULL. */
void orc_closest_neighbors(const orc_params *prm, const orc_cells *C, const orc_walls *W, int n, const double *x,
                           const int *surfId, double epsDist, double *dist_cell, double *x0_cell, double *dist_wall,
                           double *x0_wall) {
  const int n1 = prm->Nc[0] + 2, n2 = prm->Nc[1] + 2, n3 = prm->Nc[2] + 2;
  const size_t Np = (size_t)C->ncell * C->nlat * C->nlon;
  int *hoc = (int *)malloc(sizeof(int) * (size_t)n1 * n2 * n3), *next = (int *)malloc(sizeof(int) * (Np > 0 ? Np : 1));
  orc_hash_build(prm, (int)Np, C->x, hoc, next);
  wall_lists L;
  const int have_walls = W && W->nwall > 0;
  if (have_walls) wall_lists_init(prm, W, &L);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++) {
    const double xi[3] = {x[i], x[(size_t)n + i], x[2 * (size_t)n + i]};
    double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    dist_cell[i] = closest_neighbor_cell(prm, C, hoc, next, xi, surfId[i], epsDist, a);
    dist_wall[i] = have_walls ? closest_neighbor_wall(prm, W, &L, xi, surfId[i], b) : HUGE_VAL;
    for (int d = 0; d < 3; d++) x0_cell[(size_t)d * n + i] = a[d], x0_wall[(size_t)d * n + i] = b[d];
  }
  if (have_walls) wall_lists_free(&L);
  free(hoc), free(next);
}

/* ModRepulsion.F90:270-402 InterCellRepulsion up to the displacement field: for every active cell point the
 * closer of the two neighbours; if it is closer than epsDist the point is pushed away by half the deficit,
 * dx = 0.5 xx (epsDist - rr) / rr with xx the minimum image of xi - xj (:317-325).  dx SoA(3,Np) (zero rows for
 * points that do not move); returns the number of moved points, *dist_min = smallest separation seen.  The
 * caller adds dx to rbc%x (:350) -- cells below viscRatThresh; the rigid-cell branch (:352-385) averages dx. */
int orc_inter_cell_repulsion(const orc_params *prm, const orc_cells *C, const orc_walls *W, const int *active,
                             double epsDist, double *dx, double *dist_min) {
  const int npc = C->nlat * C->nlon;
  const int Np = C->ncell * npc;
  double *dc = (double *)malloc(sizeof(double) * (Np > 0 ? Np : 1)), *dw = (double *)malloc(sizeof(double) * (Np > 0 ? Np : 1));
  double *xc = (double *)malloc(sizeof(double) * 3 * (Np > 0 ? Np : 1)), *xw = (double *)malloc(sizeof(double) * 3 * (Np > 0 ? Np : 1));
  int *sid = (int *)calloc(Np > 0 ? Np : 1, sizeof(int));
  for (int i = 0; i < Np; i++) sid[i] = i / npc + 1;
  orc_closest_neighbors(prm, C, W, Np, C->x, sid, epsDist, dc, xc, dw, xw);
  int cnt = 0;
  double dmin = HUGE_VAL;
  memset(dx, 0, sizeof(double) * 3 * (size_t)Np);
  for (int i = 0; i < Np; i++) {
    if (active && !active[i]) continue;
    const double *x0 = dc[i] < dw[i] ? xc : xw; /* :304-310 */
    const double dist = dc[i] < dw[i] ? dc[i] : dw[i];
    if (dist < dmin) dmin = dist;
    if (dist < epsDist) {
      double xx[3], rr = 0;
      for (int d = 0; d < 3; d++) {
        xx[d] = C->x[(size_t)d * Np + i] - x0[(size_t)d * Np + i];
        xx[d] = xx[d] - wnint(xx[d] * prm->iLb[d]) * prm->Lb[d];
        rr += xx[d] * xx[d];
      }
      rr = sqrt(rr);
      cnt++;
      for (int d = 0; d < 3; d++) dx[(size_t)d * Np + i] = 0.5 * xx[d] * (epsDist - rr) / rr;
    }
  }
  *dist_min = dmin;
  free(dc), free(dw), free(xc), free(xw), free(sid);
  return cnt;
}
