/* oracle/adv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see header).
 *
 * Line-by-line C restatement of the reference's tracer advection (FESOM/fesom2 @ e3c3d9d).
 * Every function cites the Fortran it follows.  Loop order and floating-point expression order
 * are those of the Fortran source (left-to-right evaluation, explicit parentheses kept); build the
 * parity flavour with -O2 -ffp-contract=off so that no FMA contraction is introduced.
 */
#include "adv_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define R_EARTH 6367500.0 /* src/oce_modules.F90:29 */

/* 1-based column-major accessors */
#define IX2(ld, i, j) (((size_t)(j)-1) * (size_t)(ld) + (size_t)((i)-1))
#define IX3(d1, d2, i, j, k) ((((size_t)(k)-1) * (size_t)(d2) + (size_t)((j)-1)) * (size_t)(d1) + (size_t)((i)-1))

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* ------------------------------------------------------------------------------------------
 * horizontal fluxes: src/oce_adv_tra_hor.F90
 * ------------------------------------------------------------------------------------------ */
#define EDGE_PROLOGUE                                                                    \
    const int L = m->nl - 1;                                                             \
    const int E = m->myDim_edge2D;                                                       \
    if (init_zero) {                                                                     \
        for (int edge = 1; edge <= E; ++edge)                                            \
            for (int nz = 1; nz <= L; ++nz) flux[IX2(L, nz, edge)] = 0.0;                \
    }

#define EDGE_SETUP                                                                       \
    const int en1 = m->edges[IX2(2, 1, edge)], en2 = m->edges[IX2(2, 2, edge)];          \
    const int el1 = m->edge_tri[IX2(2, 1, edge)], el2 = m->edge_tri[IX2(2, 2, edge)];    \
    const int nl1 = m->nlevels[el1 - 1] - 1;                                             \
    const int nu1 = m->ulevels[el1 - 1];                                                 \
    const double deltaX1 = m->edge_cross_dxdy[IX2(4, 1, edge)];                          \
    const double deltaY1 = m->edge_cross_dxdy[IX2(4, 2, edge)];                          \
    double a = R_EARTH * m->elem_cos[el1 - 1];                                           \
    double deltaX2 = 0.0, deltaY2 = 0.0;                                                 \
    int nl2 = 0, nu2 = 0;                                                                \
    if (el2 > 0) {                                                                       \
        deltaX2 = m->edge_cross_dxdy[IX2(4, 3, edge)];                                   \
        deltaY2 = m->edge_cross_dxdy[IX2(4, 4, edge)];                                   \
        nl2 = m->nlevels[el2 - 1] - 1;                                                   \
        nu2 = m->ulevels[el2 - 1];                                                       \
        a = 0.5 * (a + R_EARTH * m->elem_cos[el2 - 1]);                                  \
    }                                                                                    \
    const int nl12 = imin(nl1, nl2);                                                     \
    const int nu12 = imax(nu1, nu2);                                                     \
    (void)a; (void)deltaX2; (void)deltaY2;

#define VEL(c, nz, el) vel[IX3(2, L, c, nz, el)]
#define HELEM(nz, el) m->helem[IX2(L, nz, el)]
#define VFLUX1(nz) ((-VEL(2, nz, el1) * deltaX1 + VEL(1, nz, el1) * deltaY1) * HELEM(nz, el1))
#define VFLUX2(nz) ((VEL(2, nz, el2) * deltaX2 - VEL(1, nz, el2) * deltaY2) * HELEM(nz, el2))
#define TTF(nz, n) ttf[IX2(L, nz, n)]

/* oce_adv_tra_hor.F90:64-257 */
void ora_adv_tra_hor_upw1(const ora_mesh_t *m, const double *vel, const double *ttf,
                          double *flux, int init_zero)
{
    EDGE_PROLOGUE
    for (int edge = 1; edge <= E; ++edge) {
        EDGE_SETUP
        double vflux;
#define UPW1_BODY                                                                           \
    flux[IX2(L, nz, edge)] = -0.5 * (TTF(nz, en1) * (vflux + fabs(vflux)) +                 \
                                     TTF(nz, en2) * (vflux - fabs(vflux))) -                \
                             flux[IX2(L, nz, edge)];
        /* (A) :167-178 */
        for (int nz = nu1; nz <= nu12 - 1; ++nz) { vflux = VFLUX1(nz); UPW1_BODY }
        /* (B) :185-199 */
        if (nu2 > 0)
            for (int nz = nu2; nz <= nu12 - 1; ++nz) { vflux = VFLUX2(nz); UPW1_BODY }
        /* (C) :207-217 */
        for (int nz = nu12; nz <= nl12; ++nz) { vflux = VFLUX1(nz) + VFLUX2(nz); UPW1_BODY }
        /* (D) :223-233 */
        for (int nz = nl12 + 1; nz <= nl1; ++nz) { vflux = VFLUX1(nz); UPW1_BODY }
        /* (E) :239-248 */
        for (int nz = nl12 + 1; nz <= nl2; ++nz) { vflux = VFLUX2(nz); UPW1_BODY }
#undef UPW1_BODY
    }
}

#define GRAD(k, nz, e) edge_up_dn_grad[IX3(4, L, k, nz, e)]
/* Tmean2/Tmean1 of oce_adv_tra_hor.F90:446-461 (MUSCL, with c_lo) and :736-751 (MFCT, c_lo = 1:
 * the MFCT source has no `*c_lo` factor at all, so no multiplication is performed there) */
#define HO_RECON(USE_CLO)                                                                   \
    double Tmean2 = TTF(nz, en2) -                                                          \
        (2.0 * (TTF(nz, en2) - TTF(nz, en1)) + edx * a * GRAD(2, nz, edge) +                \
         edy * R_EARTH * GRAD(4, nz, edge)) / 6.0 USE_CLO(2);                               \
    double Tmean1 = TTF(nz, en1) +                                                          \
        (2.0 * (TTF(nz, en2) - TTF(nz, en1)) + edx * a * GRAD(1, nz, edge) +                \
         edy * R_EARTH * GRAD(3, nz, edge)) / 6.0 USE_CLO(1);
#define HO_FLUX                                                                             \
    {                                                                                       \
        double cHO = (vflux + fabs(vflux)) * Tmean1 + (vflux - fabs(vflux)) * Tmean2;       \
        flux[IX2(L, nz, edge)] = -0.5 * (1.0 - num_ord) * cHO -                             \
                                 vflux * num_ord * 0.5 * (Tmean1 + Tmean2) -                \
                                 flux[IX2(L, nz, edge)];                                    \
    }
#define CLO_MUL(k) *c_lo##k
#define CLO_NONE(k)

/* c_lo(k) = real(max(sign(1, nboundary_lay(enodes(k)) - nz), 0))   (:411-412) */
static inline double clo(int nb, int nz) { return (nb - nz >= 0) ? 1.0 : 0.0; }

/* oce_adv_tra_hor.F90:261-542 */
void ora_adv_tra_hor_muscl(const ora_mesh_t *m, const double *vel, const double *ttf,
                           double num_ord, double *flux, const double *edge_up_dn_grad,
                           const int *nboundary_lay, int init_zero)
{
    EDGE_PROLOGUE
    for (int edge = 1; edge <= E; ++edge) {
        EDGE_SETUP
        const double edx = m->edge_dxdy[IX2(2, 1, edge)], edy = m->edge_dxdy[IX2(2, 2, edge)];
        const int nb1 = nboundary_lay[en1 - 1], nb2 = nboundary_lay[en2 - 1];
        double vflux;
#define MUSCL_BODY                                                                          \
    {                                                                                       \
        const double c_lo1 = clo(nb1, nz), c_lo2 = clo(nb2, nz);                            \
        HO_RECON(CLO_MUL) HO_FLUX                                                           \
    }
        for (int nz = nu1; nz <= nu12 - 1; ++nz) { vflux = VFLUX1(nz); MUSCL_BODY }          /* A :355 */
        if (nu2 > 0)
            for (int nz = nu2; nz <= nu12 - 1; ++nz) { vflux = VFLUX2(nz); MUSCL_BODY }      /* B :381 */
        for (int nz = nu12; nz <= nl12; ++nz) { vflux = VFLUX1(nz) + VFLUX2(nz); MUSCL_BODY }/* C :410 */
        for (int nz = nl12 + 1; nz <= nl1; ++nz) { vflux = VFLUX1(nz); MUSCL_BODY }          /* D :494 */
        for (int nz = nl12 + 1; nz <= nl2; ++nz) { vflux = VFLUX2(nz); MUSCL_BODY }          /* E :518 */
#undef MUSCL_BODY
    }
}

/* oce_adv_tra_hor.F90:546-834 */
void ora_adv_tra_hor_mfct(const ora_mesh_t *m, const double *vel, const double *ttf,
                          double num_ord, double *flux, const double *edge_up_dn_grad,
                          int init_zero)
{
    EDGE_PROLOGUE
    for (int edge = 1; edge <= E; ++edge) {
        EDGE_SETUP
        const double edx = m->edge_dxdy[IX2(2, 1, edge)], edy = m->edge_dxdy[IX2(2, 2, edge)];
        double vflux;
#define MFCT_BODY { HO_RECON(CLO_NONE) HO_FLUX }
        for (int nz = nu1; nz <= nu12 - 1; ++nz) { vflux = VFLUX1(nz); MFCT_BODY }           /* A :651 */
        if (nu2 > 0)
            for (int nz = nu2; nz <= nu12 - 1; ++nz) { vflux = VFLUX2(nz); MFCT_BODY }       /* B :675 */
        for (int nz = nu12; nz <= nl12; ++nz) { vflux = VFLUX1(nz) + VFLUX2(nz); MFCT_BODY } /* C :703 */
        for (int nz = nl12 + 1; nz <= nl1; ++nz) { vflux = VFLUX1(nz); MFCT_BODY }           /* D :785 */
        for (int nz = nl12 + 1; nz <= nl2; ++nz) { vflux = VFLUX2(nz); MFCT_BODY }           /* E :808 */
#undef MFCT_BODY
    }
}

/* ------------------------------------------------------------------------------------------
 * vertical fluxes: src/oce_adv_tra_ver.F90
 * ------------------------------------------------------------------------------------------ */
#define NODE_PROLOGUE                                                                    \
    const int nl = m->nl;                                                                \
    const int L = nl - 1;                                                                \
    const int N = m->myDim_nod2D;                                                        \
    (void)L;                                                                             \
    if (init_zero) {                                                                     \
        for (int n = 1; n <= N; ++n)                                                     \
            for (int nz = 1; nz <= nl; ++nz) flux[IX2(nl, nz, n)] = 0.0;                 \
    }
#define W(nz, n) w[IX2(nl, nz, n)]
#define AREA(nz, n) m->area[IX2(nl, nz, n)]
#define FLUXV(nz, n) flux[IX2(nl, nz, n)]

/* oce_adv_tra_ver.F90:244-328 */
void ora_adv_tra_ver_upw1(const ora_mesh_t *m, const double *w, const double *ttf,
                          double *flux, int init_zero)
{
    NODE_PROLOGUE
    for (int n = 1; n <= N; ++n) {
        const int nzmax = m->nlevels_nod2D[n - 1];
        const int nzmin = m->ulevels_nod2D[n - 1];
        int nz = nzmin;                                                              /* :300 */
        FLUXV(nz, n) = -W(nz, n) * TTF(nz, n) * AREA(nz, n) - FLUXV(nz, n);
        nz = nzmax;                                                                  /* :305 */
        FLUXV(nz, n) = 0.0 - FLUXV(nz, n);
        for (nz = nzmin + 1; nz <= nzmax - 1; ++nz)                                  /* :315 */
            FLUXV(nz, n) = -0.5 * (TTF(nz, n) * (W(nz, n) + fabs(W(nz, n))) +
                                   TTF(nz - 1, n) * (W(nz, n) - fabs(W(nz, n)))) * AREA(nz, n) -
                           FLUXV(nz, n);
    }
}

/* oce_adv_tra_ver.F90:332-434 */
void ora_adv_tra_ver_qr4c(const ora_mesh_t *m, const double *w, const double *ttf,
                          double num_ord, double *flux, int init_zero)
{
    NODE_PROLOGUE
#define Z3(nz, n) m->Z_3d_n[IX2(L, nz, n)]
#define ZB3(nz, n) m->zbar_3d_n[IX2(nl, nz, n)]
    for (int n = 1; n <= N; ++n) {
        const int nzmax = m->nlevels_nod2D[n - 1];
        const int nzmin = m->ulevels_nod2D[n - 1];
        int nz = nzmin;                                                              /* :390 */
        FLUXV(nz, n) = -TTF(nz, n) * W(nz, n) * AREA(nz, n) - FLUXV(nz, n);
        nz = nzmin + 1;                                                              /* :395 */
        FLUXV(nz, n) = -0.5 * (TTF(nz - 1, n) + TTF(nz, n)) * W(nz, n) * AREA(nz, n) - FLUXV(nz, n);
        nz = nzmax - 1;                                                              /* :400 */
        FLUXV(nz, n) = -0.5 * (TTF(nz - 1, n) + TTF(nz, n)) * W(nz, n) * AREA(nz, n) - FLUXV(nz, n);
        nz = nzmax;                                                                  /* :405 */
        FLUXV(nz, n) = 0.0 - FLUXV(nz, n);
        for (nz = nzmin + 2; nz <= nzmax - 2; ++nz) {                                /* :415 */
            const double qc = (TTF(nz - 1, n) - TTF(nz, n)) / (Z3(nz - 1, n) - Z3(nz, n));
            const double qu = (TTF(nz, n) - TTF(nz + 1, n)) / (Z3(nz, n) - Z3(nz + 1, n));
            const double qd = (TTF(nz - 2, n) - TTF(nz - 1, n)) / (Z3(nz - 2, n) - Z3(nz - 1, n));
            const double Tmean1 = TTF(nz, n) + (2 * qc + qu) * (ZB3(nz, n) - Z3(nz, n)) / 3.0;
            const double Tmean2 = TTF(nz - 1, n) + (2 * qc + qd) * (ZB3(nz, n) - Z3(nz - 1, n)) / 3.0;
            const double Tmean = (W(nz, n) + fabs(W(nz, n))) * Tmean1 + (W(nz, n) - fabs(W(nz, n))) * Tmean2;
            FLUXV(nz, n) = (-0.5 * (1.0 - num_ord) * Tmean -
                            num_ord * (0.5 * (Tmean1 + Tmean2)) * W(nz, n)) * AREA(nz, n) -
                           FLUXV(nz, n);
        }
    }
}

static inline double dsign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); } /* Fortran sign() */
static inline double dmin3(double a, double b, double c) { return dmin(dmin(a, b), c); }

/* oce_adv_tra_ver.F90:438-631 */
void ora_adv_tra_vert_ppm(const ora_mesh_t *m, double dt, const double *w, const double *ttf,
                          double *flux, int init_zero)
{
    NODE_PROLOGUE
#define HN(nz, n) m->hnode[IX2(L, nz, n)]
#define HNN(nz, n) m->hnode_new[IX2(L, nz, n)]
    double *tv = (double *)malloc(sizeof(double) * (size_t)(nl + 2));
    double *tvert = (double *)malloc(sizeof(double) * (size_t)(nl + 2));
    for (int n = 1; n <= N; ++n) {
        const int nzmax = m->nlevels_nod2D[n - 1];
        const int nzmin = m->ulevels_nod2D[n - 1];
        tv[nzmin] = TTF(nzmin, n);                                                   /* :496 */
        tv[nzmin + 1] = 0.5 * (TTF(nzmin, n) + TTF(nzmin + 1, n));                   /* :500 */
        tv[nzmax - 1] = 0.5 * (TTF(nzmax - 2, n) + TTF(nzmax - 1, n));               /* :504 */
        tv[nzmax] = TTF(nzmax - 1, n);                                               /* :507 */
        for (int nz = nzmin + 1; nz <= nzmax - 3; ++nz) {                            /* :514 */
            const double dzjm1 = HNN(nz - 1, n), dzj = HNN(nz, n), dzjp1 = HNN(nz + 1, n), dzjp2 = HNN(nz + 2, n);
            double deltaj = dzj / (dzjm1 + dzj + dzjp1) *
                            ((2.0 * dzjm1 + dzj) / (dzjp1 + dzj) * (TTF(nz + 1, n) - TTF(nz, n)) +
                             (dzj + 2.0 * dzjp1) / (dzjm1 + dzj) * (TTF(nz, n) - TTF(nz - 1, n)));
            double deltajp1 = dzjp1 / (dzj + dzjp1 + dzjp2) *
                              ((2.0 * dzj + dzjp1) / (dzjp2 + dzjp1) * (TTF(nz + 2, n) - TTF(nz + 1, n)) +
                               (dzjp1 + 2.0 * dzjp2) / (dzj + dzjp1) * (TTF(nz + 1, n) - TTF(nz, n)));
            if ((TTF(nz + 1, n) - TTF(nz, n)) * (TTF(nz, n) - TTF(nz - 1, n)) > 0.0)   /* :551 */
                deltaj = dmin3(fabs(deltaj), 2.0 * fabs(TTF(nz + 1, n) - TTF(nz, n)),
                               2.0 * fabs(TTF(nz, n) - TTF(nz - 1, n))) * dsign(1.0, deltaj);
            else
                deltaj = 0.0;
            if ((TTF(nz + 2, n) - TTF(nz + 1, n)) * (TTF(nz + 1, n) - TTF(nz, n)) > 0.0) /* :559 */
                deltajp1 = dmin3(fabs(deltajp1), 2.0 * fabs(TTF(nz + 2, n) - TTF(nz + 1, n)),
                                 2.0 * fabs(TTF(nz + 1, n) - TTF(nz, n))) * dsign(1.0, deltajp1);
            else
                deltajp1 = 0.0;
            tv[nz + 1] = TTF(nz, n) + dzj / (dzj + dzjp1) * (TTF(nz + 1, n) - TTF(nz, n)) +   /* :571 */
                         1.0 / (dzjm1 + dzj + dzjp1 + dzjp2) *
                             ((2.0 * dzjp1 * dzj) / (dzj + dzjp1) *
                                  ((dzjm1 + dzj) / (2.0 * dzj + dzjp1) - (dzjp2 + dzjp1) / (2.0 * dzjp1 + dzj)) *
                                  (TTF(nz + 1, n) - TTF(nz, n)) -
                              dzj * (dzjm1 + dzj) / (2.0 * dzj + dzjp1) * deltajp1 +
                              dzjp1 * (dzjp1 + dzjp2) / (dzj + 2.0 * dzjp1) * deltaj);
        }
        for (int nz = 1; nz <= nzmax; ++nz) tvert[nz] = 0.0;                         /* :583 */
        for (int nz = nzmin; nz <= nzmax - 1; ++nz) {                                /* :585 */
            if ((W(nz, n) <= 0.0) && (W(nz + 1, n) >= 0.0)) continue;
            double aL = tv[nz], aR = tv[nz + 1];
            if ((aR - TTF(nz, n)) * (TTF(nz, n) - aL) <= 0.0) { aL = TTF(nz, n); aR = TTF(nz, n); }
            if ((aR - aL) * (TTF(nz, n) - 0.5 * (aL + aR)) > (aR - aL) * (aR - aL) / 6.0)
                aL = 3.0 * TTF(nz, n) - 2.0 * aR;
            if ((aR - aL) * (TTF(nz, n) - 0.5 * (aR + aL)) < -((aR - aL) * (aR - aL)) / 6.0)
                aR = 3.0 * TTF(nz, n) - 2.0 * aL;
            const double dzj = HN(nz, n);                                            /* :603 */
            const double aj = 6.0 * (TTF(nz, n) - 0.5 * (aL + aR));
            if (W(nz, n) > 0.0) {
                const double x = dmin(W(nz, n) * dt / dzj, 1.0);
                tvert[nz] = (-aL - 0.5 * x * (aR - aL + (1.0 - 2.0 / 3.0 * x) * aj));
                tvert[nz] = tvert[nz] * AREA(nz, n) * W(nz, n);
            }
            if (W(nz + 1, n) < 0.0) {
                const double x = dmin(-W(nz + 1, n) * dt / dzj, 1.0);
                tvert[nz + 1] = (-aR + 0.5 * x * (aR - aL - (1.0 - 2.0 / 3.0 * x) * aj));
                tvert[nz + 1] = tvert[nz + 1] * AREA(nz + 1, n) * W(nz + 1, n);
            }
        }
        tvert[nzmin] = -tv[nzmin] * W(nzmin, n) * AREA(nzmin, n);                    /* :623 */
        tvert[nzmax] = 0.0;
        for (int nz = nzmin; nz <= nzmax; ++nz) FLUXV(nz, n) = tvert[nz] - FLUXV(nz, n); /* :626 */
    }
    free(tv);
    free(tvert);
}

/* oce_adv_tra_ver.F90:635-695.  NB nzmax = nlevels-1 here and the bottom interface nzmax+1 is
 * never written to flux. */
void ora_adv_tra_ver_cdiff(const ora_mesh_t *m, const double *w, const double *ttf,
                           double *flux, int init_zero)
{
    NODE_PROLOGUE
    double *tvert = (double *)malloc(sizeof(double) * (size_t)(nl + 2));
    for (int n = 1; n <= N; ++n) {
        const int nzmax = m->nlevels_nod2D[n - 1] - 1;
        const int nzmin = m->ulevels_nod2D[n - 1];
        tvert[nzmin] = -W(nzmin, n) * TTF(nzmin, n) * AREA(nzmin, n);
        tvert[nzmax + 1] = 0.0;
        for (int nz = nzmin + 1; nz <= nzmax; ++nz) {
            const double tvv = 0.5 * (TTF(nz - 1, n) + TTF(nz, n));
            tvert[nz] = -tvv * W(nz, n) * AREA(nz, n);
        }
        for (int nz = nzmin; nz <= nzmax; ++nz) FLUXV(nz, n) = tvert[nz] - FLUXV(nz, n);
    }
    free(tvert);
}

/* oce_adv_tra_ver.F90:90-240 (implicit upwind part of the w-split; Thomas algorithm) */
void ora_adv_tra_vert_impl(const ora_mesh_t *m, double dt, const double *w, double *ttf)
{
    const int nl = m->nl, L = nl - 1, N = m->myDim_nod2D;
#define AVOL(nz, n) m->areasvol[IX2(nl, nz, n)]
    double *a = (double *)malloc(sizeof(double) * (size_t)(nl + 2) * 6);
    double *b = a + (nl + 2), *c = b + (nl + 2), *tr = c + (nl + 2), *cp = tr + (nl + 2), *tp = cp + (nl + 2);
    for (int n = 1; n <= N; ++n) {
        for (int k = 0; k < nl + 2; ++k) a[k] = b[k] = c[k] = tr[k] = tp[k] = cp[k] = 0.0;
        const int nzmax = m->nlevels_nod2D[n - 1];
        const int nzmin = m->ulevels_nod2D[n - 1];
        /* zbar_n / Z_n (:142-150) are computed by the reference but never used afterwards */
        int nz = nzmin;
        const double zinv = 1.0 * dt;                                                /* :157 */
        double v_adv;
        a[nz] = 0.0;
        v_adv = zinv * AREA(nz, n) / AVOL(nz, n);
        b[nz] = HNN(nz, n) + W(nz, n) * v_adv;
        v_adv = zinv * AREA(nz + 1, n) / AVOL(nz, n);
        b[nz] = b[nz] - dmin(0.0, W(nz + 1, n)) * v_adv;
        c[nz] = -dmax(0.0, W(nz + 1, n)) * v_adv;
        for (nz = nzmin + 1; nz <= nzmax - 2; ++nz) {                                /* :174 */
            v_adv = zinv * AREA(nz, n) / AVOL(nz, n);
            a[nz] = dmin(0.0, W(nz, n)) * v_adv;
            b[nz] = HNN(nz, n) + dmax(0.0, W(nz, n)) * v_adv;
            v_adv = zinv * AREA(nz + 1, n) / AVOL(nz, n);
            b[nz] = b[nz] - dmin(0.0, W(nz + 1, n)) * v_adv;
            c[nz] = -dmax(0.0, W(nz + 1, n)) * v_adv;
        }
        nz = nzmax - 1;                                                              /* :187 */
        v_adv = zinv * AREA(nz, n) / AVOL(nz, n);
        a[nz] = dmin(0.0, W(nz, n)) * v_adv;
        b[nz] = HNN(nz, n) + dmax(0.0, W(nz, n)) * v_adv;
        c[nz] = 0.0;
        nz = nzmin;                                                                  /* :198 */
        double dz = HNN(nz, n);
        tr[nz] = -(b[nz] - dz) * TTF(nz, n) - c[nz] * TTF(nz + 1, n);
        for (nz = nzmin + 1; nz <= nzmax - 2; ++nz) {
            dz = HNN(nz, n);
            tr[nz] = -a[nz] * TTF(nz - 1, n) - (b[nz] - dz) * TTF(nz, n) - c[nz] * TTF(nz + 1, n);
        }
        nz = nzmax - 1;
        dz = HNN(nz, n);
        tr[nz] = -a[nz] * TTF(nz - 1, n) - (b[nz] - dz) * TTF(nz, n);
        nz = nzmin;                                                                  /* :211 */
        cp[nz] = c[nz] / b[nz];
        tp[nz] = tr[nz] / b[nz];
        for (nz = nzmin + 1; nz <= nzmax - 1; ++nz) {
            const double mm = b[nz] - cp[nz - 1] * a[nz];
            cp[nz] = c[nz] / mm;
            tp[nz] = (tr[nz] - tp[nz - 1] * a[nz]) / mm;
        }
        tr[nzmax - 1] = tp[nzmax - 1];                                               /* :224 */
        for (nz = nzmax - 2; nz >= nzmin; --nz) tr[nz] = tp[nz] - cp[nz] * tr[nz + 1];
        for (nz = nzmin; nz <= nzmax - 1; ++nz) ttf[IX2(L, nz, n)] = TTF(nz, n) + tr[nz]; /* :233 */
    }
    free(a);
}

/* ------------------------------------------------------------------------------------------
 * FCT limiter: src/oce_adv_tra_fct.F90:72-512
 * ------------------------------------------------------------------------------------------ */
void ora_oce_tra_adv_fct(const ora_mesh_t *m, double dt, const double *ttf, const double *lo,
                         double *adf_h, double *adf_v, double *fct_ttf_min, double *fct_ttf_max,
                         double *fct_plus, double *fct_minus, double *AUX,
                         double *tvert_max, double *tvert_min,
                         ora_exchange_fn xchg, void *user)
{
    const int nl = m->nl, L = nl - 1;
    const int N = m->myDim_nod2D, Nh = N + m->eDim_nod2D, T = m->myDim_elem2D, E = m->myDim_edge2D;
    const double flux_eps = 1e-16;  /* :99  */
    const double bignumber = 1e3;   /* :100 */
#define LO(nz, n) lo[IX2(L, nz, n)]
#define TMAX(nz, n) fct_ttf_max[IX2(L, nz, n)]
#define TMIN(nz, n) fct_ttf_min[IX2(L, nz, n)]
#define PLUS(nz, n) fct_plus[IX2(L, nz, n)]
#define MINUS(nz, n) fct_minus[IX2(L, nz, n)]
#define AUXA(k, nz, e) AUX[IX3(4, L, k, nz, e)]
#define TVMAX(nz, n) tvert_max[IX2(L, nz, n)]
#define TVMIN(nz, n) tvert_min[IX2(L, nz, n)]
#define ADFH(nz, e) adf_h[IX2(L, nz, e)]
#define ADFV(nz, n) adf_v[IX2(nl, nz, n)]
    /* a1 :124-133 */
    for (int n = 1; n <= Nh; ++n) {
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        for (int nz = nu1; nz <= nl1 - 1; ++nz) {
            TMAX(nz, n) = dmax(LO(nz, n), TTF(nz, n));
            TMIN(nz, n) = dmin(LO(nz, n), TTF(nz, n));
        }
    }
    /* a2 :148-178 */
    for (int elem = 1; elem <= T; ++elem) {
        const int e1 = m->elem2D_nodes[IX2(3, 1, elem)], e2 = m->elem2D_nodes[IX2(3, 2, elem)],
                  e3 = m->elem2D_nodes[IX2(3, 3, elem)];
        const int nu1 = m->ulevels[elem - 1], nl1 = m->nlevels[elem - 1];
        if (nu1 > 1)
            for (int nz = 1; nz <= nu1 - 1; ++nz) { AUXA(1, nz, elem) = -bignumber; AUXA(2, nz, elem) = bignumber; }
        for (int nz = nu1; nz <= nl1 - 1; ++nz) {
            AUXA(1, nz, elem) = dmax(dmax(TMAX(nz, e1), TMAX(nz, e2)), TMAX(nz, e3));
            AUXA(2, nz, elem) = dmin(dmin(TMIN(nz, e1), TMIN(nz, e2)), TMIN(nz, e3));
        }
        if (nl1 <= nl - 1)
            for (int nz = nl1; nz <= nl - 1; ++nz) { AUXA(1, nz, elem) = -bignumber; AUXA(2, nz, elem) = bignumber; }
    }
    /* a3 :195-215 */
    for (int n = 1; n <= N; ++n) {
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        for (int nz = nu1; nz <= nl1 - 1; ++nz) {
            const int ld = m->nod_in_elem_ld;
            TVMAX(nz, n) = AUXA(1, nz, m->nod_in_elem2D[IX2(ld, 1, n)]);
            TVMIN(nz, n) = AUXA(2, nz, m->nod_in_elem2D[IX2(ld, 1, n)]);
            for (int k = 2; k <= m->nod_in_elem2D_num[n - 1]; ++k) {
                TVMAX(nz, n) = dmax(TVMAX(nz, n), AUXA(1, nz, m->nod_in_elem2D[IX2(ld, k, n)]));
                TVMIN(nz, n) = dmin(TVMIN(nz, n), AUXA(2, nz, m->nod_in_elem2D[IX2(ld, k, n)]));
            }
        }
    }
    /* :227-248 */
    for (int n = 1; n <= N; ++n) {
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        TMAX(nu1, n) = TVMAX(nu1, n) - LO(nu1, n);
        TMIN(nu1, n) = TVMIN(nu1, n) - LO(nu1, n);
        for (int nz = nu1 + 1; nz <= nl1 - 2; ++nz) {
            TMAX(nz, n) = dmax(dmax(TVMAX(nz - 1, n), TVMAX(nz, n)), TVMAX(nz + 1, n)) - LO(nz, n);
            TMIN(nz, n) = dmin(dmin(TVMIN(nz - 1, n), TVMIN(nz, n)), TVMIN(nz + 1, n)) - LO(nz, n);
        }
        const int nz = nl1 - 1;
        TMAX(nz, n) = TVMAX(nz, n) - LO(nz, n);
        TMIN(nz, n) = TVMIN(nz, n) - LO(nz, n);
    }
    /* b1 :265-295 */
    for (int n = 1; n <= N; ++n) {
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        for (int nz = nu1; nz <= nl1 - 1; ++nz) { PLUS(nz, n) = 0.0; MINUS(nz, n) = 0.0; }
    }
    for (int n = 1; n <= N; ++n) {
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        for (int nz = nu1; nz <= nl1 - 1; ++nz) {
            PLUS(nz, n) = PLUS(nz, n) + (dmax(0.0, ADFV(nz, n)) + dmax(0.0, -ADFV(nz + 1, n)));
            MINUS(nz, n) = MINUS(nz, n) + (dmin(0.0, ADFV(nz, n)) + dmin(0.0, -ADFV(nz + 1, n)));
        }
    }
    /* :312-377 */
    for (int edge = 1; edge <= E; ++edge) {
        const int en1 = m->edges[IX2(2, 1, edge)], en2 = m->edges[IX2(2, 2, edge)];
        const int el1 = m->edge_tri[IX2(2, 1, edge)], el2 = m->edge_tri[IX2(2, 2, edge)];
        const int nl1 = m->nlevels[el1 - 1] - 1, nu1 = m->ulevels[el1 - 1];
        int nl2 = 0, nu2 = 0;
        if (el2 > 0) { nl2 = m->nlevels[el2 - 1] - 1; nu2 = m->ulevels[el2 - 1]; }
        const int nl12 = imax(nl1, nl2);
        int nu12 = nu1;
        if (nu2 > 0) nu12 = imin(nu1, nu2);
        for (int nz = nu12; nz <= nl12; ++nz) {
            PLUS(nz, en1) = PLUS(nz, en1) + dmax(0.0, ADFH(nz, edge));
            MINUS(nz, en1) = MINUS(nz, en1) + dmin(0.0, ADFH(nz, edge));
            PLUS(nz, en2) = PLUS(nz, en2) + dmax(0.0, -ADFH(nz, edge));
            MINUS(nz, en2) = MINUS(nz, en2) + dmin(0.0, -ADFH(nz, edge));
        }
    }
    /* b2 :394-405 */
    for (int n = 1; n <= N; ++n) {
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        for (int nz = nu1; nz <= nl1 - 1; ++nz) {
            double flux = PLUS(nz, n) * dt / AVOL(nz, n) / HNN(nz, n) + flux_eps;
            PLUS(nz, n) = dmin(1.0, TMAX(nz, n) / flux);
            flux = MINUS(nz, n) * dt / AVOL(nz, n) / HNN(nz, n) - flux_eps;
            MINUS(nz, n) = dmin(1.0, TMIN(nz, n) / flux);
        }
    }
    /* :413 exchange_nod(fct_plus, fct_minus) */
    if (xchg) { xchg(user, fct_plus, L); xchg(user, fct_minus, L); }
    /* b3 vertical :425-456 */
    for (int n = 1; n <= N; ++n) {
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        int nz = nu1;
        double ae = 1.0;
        double flux = ADFV(nz, n);
        if (flux >= 0.0) ae = dmin(ae, PLUS(nz, n));
        else ae = dmin(ae, MINUS(nz, n));
        ADFV(nz, n) = ae * ADFV(nz, n);
        for (nz = nu1 + 1; nz <= nl1 - 1; ++nz) {
            ae = 1.0;
            flux = ADFV(nz, n);
            if (flux >= 0.0) { ae = dmin(ae, MINUS(nz - 1, n)); ae = dmin(ae, PLUS(nz, n)); }
            else { ae = dmin(ae, PLUS(nz - 1, n)); ae = dmin(ae, MINUS(nz, n)); }
            ADFV(nz, n) = ae * ADFV(nz, n);
        }
    }
    /* b3 horizontal :468-500 */
    for (int edge = 1; edge <= E; ++edge) {
        const int en1 = m->edges[IX2(2, 1, edge)], en2 = m->edges[IX2(2, 2, edge)];
        const int el1 = m->edge_tri[IX2(2, 1, edge)], el2 = m->edge_tri[IX2(2, 2, edge)];
        const int nu1 = m->ulevels[el1 - 1], nl1 = m->nlevels[el1 - 1] - 1;
        int nl2 = 0, nu2 = 0;
        if (el2 > 0) { nu2 = m->ulevels[el2 - 1]; nl2 = m->nlevels[el2 - 1] - 1; }
        const int nl12 = imax(nl1, nl2);
        int nu12 = nu1;
        if (nu2 > 0) nu12 = imin(nu1, nu2);
        for (int nz = nu12; nz <= nl12; ++nz) {
            double ae = 1.0;
            const double flux = ADFH(nz, edge);
            if (flux >= 0.0) { ae = dmin(ae, PLUS(nz, en1)); ae = dmin(ae, MINUS(nz, en2)); }
            else { ae = dmin(ae, MINUS(nz, en1)); ae = dmin(ae, PLUS(nz, en2)); }
            ADFH(nz, edge) = ae * ADFH(nz, edge);
        }
    }
}

/* oce_adv_tra_driver.F90:494-646 */
void ora_oce_tra_adv_flux2dtracer(const ora_mesh_t *m, double dt, double *dttf_h, double *dttf_v,
                                  double *flux_h, double *flux_v, int use_lo,
                                  const double *ttf, const double *lo)
{
    const int nl = m->nl, L = nl - 1, N = m->myDim_nod2D, E = m->myDim_edge2D;
#define DH(nz, n) dttf_h[IX2(L, nz, n)]
#define DV(nz, n) dttf_v[IX2(L, nz, n)]
    if (use_lo) {                                                                    /* :522-545 */
        for (int n = 1; n <= N; ++n) {
            const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
            for (int nz = nu1; nz <= nl1 - 1; ++nz)
                DV(nz, n) = DV(nz, n) - TTF(nz, n) * HN(nz, n) + LO(nz, n) * HNN(nz, n);
        }
    }
    for (int n = 1; n <= N; ++n) {                                                   /* :551-559 */
        const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
        for (int nz = nu1; nz <= nl1 - 1; ++nz)
            DV(nz, n) = DV(nz, n) + (flux_v[IX2(nl, nz, n)] - flux_v[IX2(nl, nz + 1, n)]) * dt / AVOL(nz, n);
    }
    for (int edge = 1; edge <= E; ++edge) {                                          /* :575-633 */
        const int en1 = m->edges[IX2(2, 1, edge)], en2 = m->edges[IX2(2, 2, edge)];
        const int el1 = m->edge_tri[IX2(2, 1, edge)], el2 = m->edge_tri[IX2(2, 2, edge)];
        const int nl1 = m->nlevels[el1 - 1] - 1, nu1 = m->ulevels[el1 - 1];
        int nl2 = 0, nu2 = 0;
        if (el2 > 0) { nl2 = m->nlevels[el2 - 1] - 1; nu2 = m->ulevels[el2 - 1]; }
        const int nl12 = imax(nl1, nl2);
        int nu12 = nu1;
        if (nu2 > 0) nu12 = imin(nu1, nu2);
        for (int nz = nu12; nz <= nl12; ++nz) {
            DH(nz, en1) = DH(nz, en1) + flux_h[IX2(L, nz, edge)] * dt / AVOL(nz, en1);
            DH(nz, en2) = DH(nz, en2) - flux_h[IX2(L, nz, edge)] * dt / AVOL(nz, en2);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * driver: src/oce_adv_tra_driver.F90:46-490 (ltra_diag = .false., ldiag_DVD = .false.)
 * ------------------------------------------------------------------------------------------ */
int ora_do_oce_adv_tra(const ora_mesh_t *m, ora_work_t *wk, double dt,
                       const double *vel, const double *w, const double *wi, const double *we,
                       int use_wsplit,
                       const double *ttf, const double *ttfAB, const double *edge_up_dn_grad,
                       int hor, int ver, int lim, double opth, double optv,
                       double *dttf_h, double *dttf_v,
                       ora_exchange_fn xchg, void *user)
{
    const int nl = m->nl, L = nl - 1;
    const int N = m->myDim_nod2D, Nh = N + m->eDim_nod2D, E = m->myDim_edge2D;
    double *fct_LO = wk->fct_LO, *adv_flux_hor = wk->adv_flux_hor, *adv_flux_ver = wk->adv_flux_ver;
    const int fct = (lim == ORA_LIM_FCT);

    if (fct) {
        ora_adv_tra_hor_upw1(m, vel, ttf, adv_flux_hor, 1);                          /* :115 */
        for (int n = 1; n <= Nh; ++n)                                                /* :122-126 */
            for (int nz = 1; nz <= L; ++nz) fct_LO[IX2(L, nz, n)] = 0.0;
        for (int e = 1; e <= E; ++e) {                                               /* :142-201 */
            const int en1 = m->edges[IX2(2, 1, e)], en2 = m->edges[IX2(2, 2, e)];
            const int el1 = m->edge_tri[IX2(2, 1, e)], el2 = m->edge_tri[IX2(2, 2, e)];
            const int nl1 = m->nlevels[el1 - 1] - 1, nu1 = m->ulevels[el1 - 1];
            int nl2 = 0, nu2 = 0;
            if (el2 > 0) { nl2 = m->nlevels[el2 - 1] - 1; nu2 = m->ulevels[el2 - 1]; }
            const int nl12 = imax(nl1, nl2);
            int nu12 = nu1;
            if (nu2 > 0) nu12 = imin(nu1, nu2);
            for (int nz = nu12; nz <= nl12; ++nz) {
                fct_LO[IX2(L, nz, en1)] = fct_LO[IX2(L, nz, en1)] + adv_flux_hor[IX2(L, nz, e)];
                fct_LO[IX2(L, nz, en2)] = fct_LO[IX2(L, nz, en2)] - adv_flux_hor[IX2(L, nz, e)];
            }
        }
        if (wk->tra_advhoriz)                                                        /* :221-229 ltra_diag: LO, horizontal part */
            for (int n = 1; n <= Nh; ++n) {
                const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
                for (int nz = nu1; nz <= nl1 - 1; ++nz)
                    wk->tra_advhoriz[IX2(L, nz, n)] = fct_LO[IX2(L, nz, n)] * dt / AVOL(nz, n) / HNN(nz, n);
            }
        ora_adv_tra_ver_upw1(m, we, ttf, adv_flux_ver, 1);                           /* :235 */
        for (int n = 1; n <= N; ++n) {                                               /* :243-252 */
            const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
            for (int nz = nu1; nz <= nl1 - 1; ++nz)
                fct_LO[IX2(L, nz, n)] =
                    (TTF(nz, n) * HN(nz, n) +
                     (fct_LO[IX2(L, nz, n)] + (adv_flux_ver[IX2(nl, nz, n)] - adv_flux_ver[IX2(nl, nz + 1, n)])) *
                         dt / AVOL(nz, n)) /
                    HNN(nz, n);
        }
        if (wk->dvd_trflx_hor)                                                       /* :263-281 ldiag_DVD: LO fluxes */
            for (int e = 1; e <= E; ++e)
                for (int nz = 1; nz <= L; ++nz) wk->dvd_trflx_hor[IX2(L, nz, e)] = adv_flux_hor[IX2(L, nz, e)];
        if (wk->dvd_trflx_ver)                                                       /* :283-296 */
            for (int n = 1; n <= N; ++n)
                for (int nz = 1; nz <= nl; ++nz) wk->dvd_trflx_ver[IX2(nl, nz, n)] = adv_flux_ver[IX2(nl, nz, n)];
        if (wk->tra_advvert)                                                         /* :307-318 ltra_diag: LO, vertical part */
            for (int n = 1; n <= N; ++n) {
                const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
                for (int nz = nu1; nz <= nl1 - 1; ++nz)
                    wk->tra_advvert[IX2(L, nz, n)] =
                        (adv_flux_ver[IX2(nl, nz, n)] - adv_flux_ver[IX2(nl, nz + 1, n)]) * dt / AVOL(nz, n) / HNN(nz, n);
            }
        if (use_wsplit) {                                                            /* :323-334 */
            ora_adv_tra_vert_impl(m, dt, wi, fct_LO);
            ora_adv_tra_ver_upw1(m, w, ttf, adv_flux_ver, 1);
        }
        if (xchg) xchg(user, fct_LO, L);                                             /* :335 */
    }
    const int do_zero_flux = fct ? 0 : 1;                                            /* :339-340 */
    switch (hor) {                                                                   /* :343-354 */
    case ORA_HOR_MUSCL:
        ora_adv_tra_hor_muscl(m, vel, ttfAB, opth, adv_flux_hor, edge_up_dn_grad, wk->nboundary_lay, do_zero_flux);
        break;
    case ORA_HOR_MFCT:
        ora_adv_tra_hor_mfct(m, vel, ttfAB, opth, adv_flux_hor, edge_up_dn_grad, do_zero_flux);
        break;
    case ORA_HOR_UPW1:
        ora_adv_tra_hor_upw1(m, vel, ttfAB, adv_flux_hor, do_zero_flux);
        break;
    default:
        return 1;
    }
    const double *pwvel = fct ? w : we;                                              /* :355-359 */
    switch (ver) {                                                                   /* :363-379 */
    case ORA_VER_QR4C:  ora_adv_tra_ver_qr4c(m, pwvel, ttfAB, optv, adv_flux_ver, do_zero_flux); break;
    case ORA_VER_CDIFF: ora_adv_tra_ver_cdiff(m, pwvel, ttfAB, adv_flux_ver, do_zero_flux); break;
    case ORA_VER_PPM:   ora_adv_tra_vert_ppm(m, dt, pwvel, ttfAB, adv_flux_ver, do_zero_flux); break;
    case ORA_VER_UPW1:  ora_adv_tra_ver_upw1(m, pwvel, ttfAB, adv_flux_ver, do_zero_flux); break;
    default:
        return 1;
    }
    if (fct) {                                                                       /* :382-388 */
        ora_oce_tra_adv_fct(m, dt, ttf, fct_LO, adv_flux_hor, adv_flux_ver, wk->fct_ttf_min, wk->fct_ttf_max,
                            wk->fct_plus, wk->fct_minus, wk->AUX, wk->tvert_max, wk->tvert_min, xchg, user);
        ora_oce_tra_adv_flux2dtracer(m, dt, dttf_h, dttf_v, adv_flux_hor, adv_flux_ver, 1, ttf, fct_LO);
    } else {
        ora_oce_tra_adv_flux2dtracer(m, dt, dttf_h, dttf_v, adv_flux_hor, adv_flux_ver, 0, NULL, NULL);
    }
    if (wk->dvd_trflx_hor)                                                           /* :395-458: + the (limited) antidiffusive flux, or the HO flux itself */
        for (int e = 1; e <= E; ++e)
            for (int nz = 1; nz <= L; ++nz)
                wk->dvd_trflx_hor[IX2(L, nz, e)] = fct ? wk->dvd_trflx_hor[IX2(L, nz, e)] + adv_flux_hor[IX2(L, nz, e)] : adv_flux_hor[IX2(L, nz, e)];
    if (wk->dvd_trflx_ver)
        for (int n = 1; n <= N; ++n)
            for (int nz = 1; nz <= nl; ++nz)
                wk->dvd_trflx_ver[IX2(nl, nz, n)] = fct ? wk->dvd_trflx_ver[IX2(nl, nz, n)] + adv_flux_ver[IX2(nl, nz, n)] : adv_flux_ver[IX2(nl, nz, n)];
    if (wk->tra_advhoriz || wk->tra_advvert)                                         /* :464-488 ltra_diag */
        for (int n = 1; n <= Nh; ++n) {
            const int nu1 = m->ulevels_nod2D[n - 1], nl1 = m->nlevels_nod2D[n - 1];
            for (int nz = nu1; nz <= nl1 - 1; ++nz) {
                if (wk->tra_advhoriz) {
                    if (fct) wk->tra_advhoriz[IX2(L, nz, n)] = wk->tra_advhoriz[IX2(L, nz, n)] + dttf_h[IX2(L, nz, n)] / HNN(nz, n);   /* :472 */
                    else wk->tra_advhoriz[IX2(L, nz, n)] = dttf_h[IX2(L, nz, n)] / HNN(nz, n);                                         /* :482 */
                }
                if (wk->tra_advvert) {
                    if (fct) wk->tra_advvert[IX2(L, nz, n)] = wk->tra_advvert[IX2(L, nz, n)] + dttf_v[IX2(L, nz, n)] / HNN(nz, n);     /* :473 */
                    else wk->tra_advvert[IX2(L, nz, n)] = dttf_v[IX2(L, nz, n)] / HNN(nz, n);                                          /* :483 */
                }
            }
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * multi-rank runner (threads stand in for MPI ranks)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int nranks;
    ora_rank_t *ranks;
    pthread_barrier_t bar;
    double **slot; /* published field pointer per rank */
} ora_shared_t;

typedef struct {
    ora_shared_t *sh;
    int rank;
    double dt;
    int nsteps, mode;
    double t0, t1;
} ora_thread_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* exchange_nod3D (gen_halo_exchange.F90:432-517): halo columns := owner's columns */
static void ora_thread_exchange(void *user, double *field, int nlev)
{
    ora_thread_t *th = (ora_thread_t *)user;
    ora_shared_t *sh = th->sh;
    if (sh->nranks == 1) return;
    ora_rank_t *rk = &sh->ranks[th->rank];
    sh->slot[th->rank] = field;
    pthread_barrier_wait(&sh->bar);
    const int N = rk->mesh.myDim_nod2D, eN = rk->mesh.eDim_nod2D;
    for (int k = 0; k < eN; ++k) {
        const double *src = sh->slot[rk->halo_owner[k]] + (size_t)(rk->halo_owner_idx[k] - 1) * (size_t)nlev;
        memcpy(field + (size_t)(N + k) * (size_t)nlev, src, sizeof(double) * (size_t)nlev);
    }
    pthread_barrier_wait(&sh->bar);
}

static void *ora_thread_main(void *arg)
{
    ora_thread_t *th = (ora_thread_t *)arg;
    ora_shared_t *sh = th->sh;
    ora_rank_t *rk = &sh->ranks[th->rank];
    const ora_mesh_t *m = &rk->mesh;
    const int L = m->nl - 1, N = m->myDim_nod2D, Nh = N + m->eDim_nod2D;
    pthread_barrier_wait(&sh->bar);
    th->t0 = now_s();
    for (int step = 0; step < th->nsteps; ++step) {
        for (int t = 0; t < rk->ntr; ++t) {
            if (th->mode == 1) { /* fesom.F90:88-95 */
                memset(rk->dttf_h[t], 0, sizeof(double) * (size_t)L * (size_t)Nh);
                memset(rk->dttf_v[t], 0, sizeof(double) * (size_t)L * (size_t)Nh);
            }
            rk->work.tra_advhoriz = rk->tra_advhoriz ? rk->tra_advhoriz[t] : NULL;
            rk->work.tra_advvert = rk->tra_advvert ? rk->tra_advvert[t] : NULL;
            rk->work.dvd_trflx_hor = rk->dvd_trflx_hor ? rk->dvd_trflx_hor[t] : NULL;
            rk->work.dvd_trflx_ver = rk->dvd_trflx_ver ? rk->dvd_trflx_ver[t] : NULL;
            ora_do_oce_adv_tra(m, &rk->work, th->dt, rk->vel, rk->w, rk->wi, rk->we, rk->use_wsplit,
                               rk->values[t], rk->valuesAB[t], rk->edge_up_dn_grad[t],
                               rk->hor[t], rk->ver[t], rk->lim[t], rk->opth[t], rk->optv[t],
                               rk->dttf_h[t], rk->dttf_v[t], ora_thread_exchange, th);
            if (th->mode == 1) {
                /* fesom.F90:105-127 with del_ttf reset every step (init_tracers_AB,
                 * oce_tracer_mod.F90:28-34): values += (advhoriz + advvert)/hnode_new on owned nodes */
                double *val = rk->values[t];
                for (int n = 1; n <= N; ++n) {
                    const int nzmax = m->nlevels_nod2D[n - 1] - 1, nzmin = m->ulevels_nod2D[n - 1];
                    for (int nz = nzmin; nz <= nzmax; ++nz) {
                        const double del = 0.0 + rk->dttf_h[t][IX2(L, nz, n)] + rk->dttf_v[t][IX2(L, nz, n)];
                        val[IX2(L, nz, n)] = val[IX2(L, nz, n)] + del / m->hnode_new[IX2(L, nz, n)];
                    }
                }
                ora_thread_exchange(th, val, L);
            }
        }
    }
    pthread_barrier_wait(&sh->bar);
    th->t1 = now_s();
    return NULL;
}

double ora_run_ranks(int nranks, ora_rank_t *ranks, double dt, int nsteps, int mode)
{
    ora_shared_t sh;
    sh.nranks = nranks;
    sh.ranks = ranks;
    sh.slot = (double **)calloc((size_t)nranks, sizeof(double *));
    pthread_barrier_init(&sh.bar, NULL, (unsigned)nranks);
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nranks);
    ora_thread_t *th = (ora_thread_t *)malloc(sizeof(ora_thread_t) * (size_t)nranks);
    for (int r = 0; r < nranks; ++r) {
        th[r].sh = &sh; th[r].rank = r; th[r].dt = dt; th[r].nsteps = nsteps; th[r].mode = mode;
        th[r].t0 = th[r].t1 = 0.0;
        pthread_create(&tid[r], NULL, ora_thread_main, &th[r]);
    }
    double t0 = 1e300, t1 = 0.0;
    for (int r = 0; r < nranks; ++r) {
        pthread_join(tid[r], NULL);
        if (th[r].t0 < t0) t0 = th[r].t0;
        if (th[r].t1 > t1) t1 = th[r].t1;
    }
    pthread_barrier_destroy(&sh.bar);
    free(tid); free(th); free(sh.slot);
    return t1 - t0;
}


/* ====================================================================================================
 * tracer_gradient_elements, src/oce_tracer_mod.F90:146-188
 * ==================================================================================================== */
void ora_tracer_gradient_elements(int nl, int myDim_elem2D, const int *elem2D_nodes, const int *nlevels,
                                  const int *ulevels, const double *gradient_sca, const double *ttf, double *tr_xy)
{
    const int L = nl - 1;
#define G_TTF(nz, n) ttf[(size_t)((n) - 1) * L + ((nz) - 1)]
#define G_TRXY(c, nz, e) tr_xy[((size_t)((e) - 1) * L + ((nz) - 1)) * 2 + ((c) - 1)]
#define G_SCA(k, e) gradient_sca[(size_t)((e) - 1) * 6 + ((k) - 1)]
    for (int elem = 1; elem <= myDim_elem2D; ++elem) {                         /* :171 */
        const int n1 = elem2D_nodes[3 * (elem - 1)], n2 = elem2D_nodes[3 * (elem - 1) + 1], n3 = elem2D_nodes[3 * (elem - 1) + 2];
        const int nzmin = ulevels[elem - 1], nzmax = nlevels[elem - 1];         /* :173-174 */
        for (int nz = nzmin; nz <= nzmax - 1; ++nz) {                           /* :176 */
            G_TRXY(1, nz, elem) = G_SCA(1, elem) * G_TTF(nz, n1) + G_SCA(2, elem) * G_TTF(nz, n2) + G_SCA(3, elem) * G_TTF(nz, n3);   /* :177 */
            G_TRXY(2, nz, elem) = G_SCA(4, elem) * G_TTF(nz, n1) + G_SCA(5, elem) * G_TTF(nz, n2) + G_SCA(6, elem) * G_TTF(nz, n3);   /* :178 */
        }
    }
#undef G_TTF
#undef G_SCA
}

/* ====================================================================================================
 * fill_up_dn_grad, src/oce_muscl_adv.F90:356-525
 * ==================================================================================================== */
static void ora_node_mean_gradient(int L, int node, int nz, const int *nod_in_elem2D, int ld, const int *nod_in_elem2D_num,
                                   const int *nlevels, const int *ulevels, const double *elem_area, const double *tr_xy,
                                   double *gx, double *gy)
{
    double tvol = 0.0, tx = 0.0, ty = 0.0;                                      /* :391-393 */
    for (int k = 1; k <= nod_in_elem2D_num[node - 1]; ++k) {                    /* :397 */
        const int elem = nod_in_elem2D[(size_t)(node - 1) * ld + (k - 1)];
        if (nlevels[elem - 1] - 1 < nz || nz < ulevels[elem - 1]) continue;     /* :400 */
        tvol = tvol + elem_area[elem - 1];
        tx = tx + G_TRXY(1, nz, elem) * elem_area[elem - 1];
        ty = ty + G_TRXY(2, nz, elem) * elem_area[elem - 1];
    }
    *gx = tx / tvol;                                                            /* :405 */
    *gy = ty / tvol;
}

void ora_fill_up_dn_grad(int nl, int myDim_edge2D, const int *edges, const int *edge_up_dn_tri,
                         const int *nod_in_elem2D, int ld, const int *nod_in_elem2D_num,
                         const int *nlevels, const int *ulevels, const int *nlevels_nod2D, const int *ulevels_nod2D,
                         const int *nlevels_nod2D_min, const int *ulevels_nod2D_max,
                         const double *elem_area, const double *tr_xy, double *edge_up_dn_grad)
{
    const int L = nl - 1;
#define G_GRAD(c, nz, e) edge_up_dn_grad[((size_t)((e) - 1) * L + ((nz) - 1)) * 4 + ((c) - 1)]
#define NODE_MEAN(node, nz, cx, cy) do { double gx_, gy_; \
        ora_node_mean_gradient(L, node, nz, nod_in_elem2D, ld, nod_in_elem2D_num, nlevels, ulevels, elem_area, tr_xy, &gx_, &gy_); \
        G_GRAD(cx, nz, edge) = gx_; G_GRAD(cy, nz, edge) = gy_; } while (0)
    for (int edge = 1; edge <= myDim_edge2D; ++edge) {                          /* :378 */
        const int e1 = edges[2 * (edge - 1)], e2 = edges[2 * (edge - 1) + 1];
        const int t1 = edge_up_dn_tri[2 * (edge - 1)], t2 = edge_up_dn_tri[2 * (edge - 1) + 1];
        if (t1 != 0 && t2 != 0) {                                               /* :382 */
            const int a = ulevels_nod2D_max[e1 - 1], b = ulevels_nod2D_max[e2 - 1];
            const int c = nlevels_nod2D_min[e1 - 1], d = nlevels_nod2D_min[e2 - 1];
            const int nzmin = a > b ? a : b, nzmax = c < d ? c : d;             /* :383-384 */
            for (int nz = ulevels_nod2D[e1 - 1]; nz <= nzmin - 1; ++nz) NODE_MEAN(e1, nz, 1, 3);      /* :388-407 */
            for (int nz = ulevels_nod2D[e2 - 1]; nz <= nzmin - 1; ++nz) NODE_MEAN(e2, nz, 2, 4);      /* :411-430 */
            for (int nz = nzmin; nz <= nzmax - 1; ++nz) {                       /* :435-440 */
                G_GRAD(1, nz, edge) = G_TRXY(1, nz, t1); G_GRAD(2, nz, edge) = G_TRXY(1, nz, t2);
                G_GRAD(3, nz, edge) = G_TRXY(2, nz, t1); G_GRAD(4, nz, edge) = G_TRXY(2, nz, t2);
            }
            for (int nz = nzmax; nz <= nlevels_nod2D[e1 - 1] - 1; ++nz) NODE_MEAN(e1, nz, 1, 3);      /* :445-463 */
            for (int nz = nzmax; nz <= nlevels_nod2D[e2 - 1] - 1; ++nz) NODE_MEAN(e2, nz, 2, 4);      /* :467-485 */
        } else {                                                                /* :489: surface boundary edge */
            for (int nz = ulevels_nod2D[e1 - 1]; nz <= nlevels_nod2D[e1 - 1] - 1; ++nz) NODE_MEAN(e1, nz, 1, 3);   /* :491-508 */
            for (int nz = ulevels_nod2D[e2 - 1]; nz <= nlevels_nod2D[e2 - 1] - 1; ++nz) NODE_MEAN(e2, nz, 2, 4);   /* :509-524 */
        }
    }
#undef NODE_MEAN
#undef G_GRAD
}
#undef G_TRXY


/* ====================================================================================================
 * vert_vel_ale, continuity part, src/oce_ale.F90:2164-2310 (linfs, Fer_GM = .false.)
 * ==================================================================================================== */
void ora_vert_vel_ale_core(const ora_mesh_t *m, const double *UV, double *Wvel)
{
    const int nl = m->nl, L = nl - 1;
    const int Nh = m->myDim_nod2D + m->eDim_nod2D;
#define V_UV(c, nz, e) UV[((size_t)((e) - 1) * L + ((nz) - 1)) * 2 + ((c) - 1)]
#define V_HE(nz, e) m->helem[(size_t)((e) - 1) * L + ((nz) - 1)]
#define V_W(nz, n) Wvel[(size_t)((n) - 1) * nl + ((nz) - 1)]
    for (int n = 1; n <= Nh; ++n)
        for (int nz = 1; nz <= nl; ++nz) V_W(nz, n) = 0.0;                          /* :2165 */
    for (int ed = 1; ed <= m->myDim_edge2D; ++ed) {                                 /* :2171 */
        const int n1 = m->edges[2 * (ed - 1)], n2 = m->edges[2 * (ed - 1) + 1];
        const int el1 = m->edge_tri[2 * (ed - 1)], el2 = m->edge_tri[2 * (ed - 1) + 1];
        const double dX1 = m->edge_cross_dxdy[4 * (ed - 1)], dY1 = m->edge_cross_dxdy[4 * (ed - 1) + 1];
        int nzmin = m->ulevels[el1 - 1], nzmax = m->nlevels[el1 - 1] - 1;           /* :2178-2179 */
        for (int nz = nzmax; nz >= nzmin; --nz) {                                   /* :2180-2185, :2192,:2201 */
            const double c1 = (V_UV(2, nz, el1) * dX1 - V_UV(1, nz, el1) * dY1) * V_HE(nz, el1);
            V_W(nz, n1) = V_W(nz, n1) + c1;
            V_W(nz, n2) = V_W(nz, n2) - c1;
        }
        if (el2 > 0) {                                                              /* :2213 */
            const double dX2 = m->edge_cross_dxdy[4 * (ed - 1) + 2], dY2 = m->edge_cross_dxdy[4 * (ed - 1) + 3];
            nzmin = m->ulevels[el2 - 1]; nzmax = m->nlevels[el2 - 1] - 1;
            for (int nz = nzmax; nz >= nzmin; --nz) {                               /* :2218-2223, :2230,:2239 */
                const double c1 = -(V_UV(2, nz, el2) * dX2 - V_UV(1, nz, el2) * dY2) * V_HE(nz, el2);
                V_W(nz, n1) = V_W(nz, n1) + c1;
                V_W(nz, n2) = V_W(nz, n2) - c1;
            }
        }
    }
    for (int n = 1; n <= m->myDim_nod2D; ++n) {                                     /* :2277-2286 */
        const int nzmin = m->ulevels_nod2D[n - 1], nzmax = m->nlevels_nod2D[n - 1] - 1;
        for (int nz = nzmax; nz >= nzmin; --nz) V_W(nz, n) = V_W(nz, n) + V_W(nz + 1, n);
    }
    for (int n = 1; n <= m->myDim_nod2D; ++n) {                                     /* :2301-2308 */
        const int nzmin = m->ulevels_nod2D[n - 1], nzmax = m->nlevels_nod2D[n - 1] - 1;
        for (int nz = nzmin; nz <= nzmax; ++nz) V_W(nz, n) = V_W(nz, n) / m->area[(size_t)(n - 1) * nl + (nz - 1)];
    }
#undef V_UV
#undef V_HE
#undef V_W
}

/* ------------------------------------------------------------------------------------------------
 * find_up_downwind_triangles (src/oce_muscl_adv.F90:162-352), single rank (coord_elem / e_nodes are then
 * plain gathers of coord_nod2D / elem2D_nodes; the exchange_elem calls of :192-216 only fill halo
 * elements).  Loop for loop: per edge, the elements around edges(1,e) are tested against -x, those around
 * edges(2,e) against +x; every element that passes one of the three tests overwrites the previous hit
 * (`cycle` only skips the remaining tests), so the LAST passing element wins.  Second, independent
 * restatement of what fesom2_b200/fields.py::find_up_downwind_triangles computes whole-array.
 * ---------------------------------------------------------------------------------------------- */
void ora_find_up_downwind_triangles(int myDim_edge2D, const int *edges, const int *elem2D_nodes,
                                    const int *nod_in_elem2D, int ld, const int *nod_in_elem2D_num,
                                    const double *coord_nod2D /* (2,N) radians */, double cyclic_length,
                                    int *edge_up_dn_tri /* (2,E) */)
{
    for (int n = 1; n <= myDim_edge2D; ++n) {
        edge_up_dn_tri[2 * (n - 1)] = 0;
        edge_up_dn_tri[2 * (n - 1) + 1] = 0;
        const int ednodes[2] = {edges[2 * (n - 1)], edges[2 * (n - 1) + 1]};
        double x[2];
        x[0] = coord_nod2D[2 * (ednodes[1] - 1)] - coord_nod2D[2 * (ednodes[0] - 1)];
        x[1] = coord_nod2D[2 * (ednodes[1] - 1) + 1] - coord_nod2D[2 * (ednodes[0] - 1) + 1];
        if (x[0] > cyclic_length / 2.0) x[0] = x[0] - cyclic_length;                 /* :222-223 */
        if (x[0] < -cyclic_length / 2.0) x[0] = x[0] + cyclic_length;
        for (int side = 0; side < 2; ++side) {
            x[0] = -x[0]; x[1] = -x[1];                                              /* :224 and :277 */
            const int node = ednodes[side];
            for (int k = 1; k <= nod_in_elem2D_num[node - 1]; ++k) {
                const int elem = nod_in_elem2D[(size_t)(node - 1) * ld + (k - 1)];
                const int *en = elem2D_nodes + 3 * (size_t)(elem - 1);
                int i0, ib, ic;                                                      /* :229-241 */
                if (en[0] == node) { i0 = 0; ib = 1; ic = 2; }
                else if (en[1] == node) { i0 = 1; ib = 0; ic = 2; }
                else { i0 = 2; ib = 0; ic = 1; }
                double b[2], c[2];
                for (int d = 0; d < 2; ++d) {
                    b[d] = coord_nod2D[2 * (size_t)(en[ib] - 1) + d] - coord_nod2D[2 * (size_t)(en[i0] - 1) + d];
                    c[d] = coord_nod2D[2 * (size_t)(en[ic] - 1) + d] - coord_nod2D[2 * (size_t)(en[i0] - 1) + d];
                }
                if (b[0] > cyclic_length / 2.0) b[0] = b[0] - cyclic_length;         /* :242-245 */
                if (b[0] < -cyclic_length / 2.0) b[0] = b[0] + cyclic_length;
                if (c[0] > cyclic_length / 2.0) c[0] = c[0] - cyclic_length;
                if (c[0] < -cyclic_length / 2.0) c[0] = c[0] + cyclic_length;
                const double cr = c[0] * c[0] + c[1] * c[1];                         /* :247-253 */
                const double bx = (b[0] * c[0] + b[1] * c[1]) / cr;
                const double by = (-b[0] * c[1] + b[1] * c[0]) / cr;
                const double xx = (x[0] * c[0] + x[1] * c[1]) / cr;
                const double xy = (-x[0] * c[1] + x[1] * c[0]) / cr;
                const double ab = atan2(by, bx), ax = atan2(xy, xx);
                int hit = 0;
                if (ab > 0.0 && ax > 0.0 && ax < ab) hit = 1;                        /* :255-258 */
                else if (ab < 0.0 && ax < 0.0 && ax > ab) hit = 1;                   /* :260-263 */
                else if (ab == ax || ax == 0.0) hit = 1;                             /* :265-268 */
                if (hit) edge_up_dn_tri[2 * (n - 1) + side] = elem;
            }
        }
    }
}

/* ====================================================================================================
 * compute_CFLz (src/oce_ale.F90:2906-2998, without the diagnostic print) and compute_Wvel_split
 * (src/oce_ale.F90:3001-3049) for all myDim+eDim nodes.  CFL_z, Wvel, Wvel_e, Wvel_i are (nl, Nh).
 * ==================================================================================================== */
void ora_compute_cflz(const ora_mesh_t *m, double dt, const double *Wvel, double *CFL_z)
{
    const int nl = m->nl, L = nl - 1;
    const int Nh = m->myDim_nod2D + m->eDim_nod2D;
#define C_W(nz, n) Wvel[(size_t)((n) - 1) * nl + ((nz) - 1)]
#define C_C(nz, n) CFL_z[(size_t)((n) - 1) * nl + ((nz) - 1)]
#define C_H(nz, n) m->hnode_new[(size_t)((n) - 1) * L + ((nz) - 1)]
    for (int n = 1; n <= Nh; ++n)
        for (int nz = 1; nz <= nl; ++nz) C_C(nz, n) = 0.0;                          /* :2933 */
    for (int n = 1; n <= Nh; ++n) {                                                 /* :2939-2952 */
        const int nzmin = m->ulevels_nod2D[n - 1], nzmax = m->nlevels_nod2D[n - 1] - 1;
        for (int nz = nzmin; nz <= nzmax; ++nz) {
            const double c1 = fabs(C_W(nz, n) * dt / C_H(nz, n));
            const double c2 = fabs(C_W(nz + 1, n) * dt / C_H(nz, n));
            C_C(nz, n) = C_C(nz, n) + c1;
            C_C(nz + 1, n) = c2;
        }
    }
#undef C_H
}

void ora_compute_wvel_split(const ora_mesh_t *m, int use_wsplit, double wsplit_maxcfl, const double *Wvel,
                            const double *CFL_z, double *Wvel_e, double *Wvel_i)
{
    const int nl = m->nl;
    const int Nh = m->myDim_nod2D + m->eDim_nod2D;
    for (int n = 1; n <= Nh; ++n) {                                                 /* :3033-3047 */
        const int nzmin = m->ulevels_nod2D[n - 1], nzmax = m->nlevels_nod2D[n - 1];
        for (int nz = nzmin; nz <= nzmax; ++nz) {
            const size_t i = (size_t)(n - 1) * nl + (nz - 1);
            Wvel_e[i] = C_W(nz, n);
            Wvel_i[i] = 0.0;
            if (use_wsplit && C_C(nz, n) > wsplit_maxcfl) {
                const double dd = fmax(C_C(nz, n) - wsplit_maxcfl, 0.0) / fmax(wsplit_maxcfl, 1.e-12);
                Wvel_e[i] = (1.0 / (1.0 + dd)) * C_W(nz, n);
                Wvel_i[i] = (dd / (1.0 + dd)) * C_W(nz, n);
            }
        }
    }
#undef C_W
#undef C_C
}

/* ====================================================================================================
 * vert_vel_ale, the 'zstar' correction for the free surface (src/oce_ale.F90:2539-2603): the change of the
 * elevation hbar - hbar_old is distributed over the layers above the shallowest bottom around the node, the
 * surface fresh-water flux closes the continuity at the top.  Owned nodes without a cavity (ulevels_nod2D = 1).
 * Wvel (nl, Nh) in/out, hnode_new (nl-1, Nh) in/out (only the stretched layers are written).
 * ==================================================================================================== */
void ora_vert_vel_ale_zstar(const ora_mesh_t *m, double dt, const int *nlevels_nod2D_min, const double *hbar,
                            const double *hbar_old, const double *water_flux, double *Wvel, double *hnode_new)
{
    const int nl = m->nl, L = nl - 1;
#define Z_W(nz, n) Wvel[(size_t)((n) - 1) * nl + ((nz) - 1)]
#define Z_ZB(nz, n) m->zbar_3d_n[(size_t)((n) - 1) * nl + ((nz) - 1)]
    for (int n = 1; n <= m->myDim_nod2D; ++n) {
        const int nzmin = m->ulevels_nod2D[n - 1];
        const int nzmax = nlevels_nod2D_min[n - 1] - 1;
        if (nzmin != 1) continue;                                                   /* :2550 */
        const double dd1 = Z_ZB(nzmax, n);                                          /* :2558 */
        double dd = Z_ZB(nzmin, n) - dd1;                                           /* :2562 */
        dd = (hbar[n - 1] - hbar_old[n - 1]) / dd;                                  /* :2566 */
        const double dddt = dd / dt;                                                /* :2570 */
        for (int nz = nzmin; nz <= nzmax - 1; ++nz) {                               /* :2574-2589 */
            Z_W(nz, n) = Z_W(nz, n) - (Z_ZB(nz, n) - dd1) * dddt;
            hnode_new[(size_t)(n - 1) * L + (nz - 1)] = m->hnode[(size_t)(n - 1) * L + (nz - 1)] + (Z_ZB(nz, n) - Z_ZB(nz + 1, n)) * dd;
        }
        Z_W(nzmin, n) = Z_W(nzmin, n) - water_flux[n - 1];                          /* :2595 */
    }
#undef Z_W
#undef Z_ZB
}

/* ====================================================================================================
 * vert_vel_ale, the free-surface correction for which_ALE = 'zlevel' (src/oce_ale.F90:2336-2538): the elevation
 * change goes into the surface layer; when that layer would get thinner than min_hnode times its rest thickness the
 * change is spread over the first lzstar_lev layers ("local zstar", :2371-2454), and a later rise refills the
 * subsurface layers first (:2455-2510).  zbar = mesh%zbar (nl), CFL_z = the previous step's (nl, Nh) array.
 * Loop for loop, including the reference's pairwise (not cumulative) "cumsum" at :2397-2398.
 * ==================================================================================================== */
void ora_vert_vel_ale_zlevel(const ora_mesh_t *m, double dt, const int *nlevels_nod2D_min, const double *hbar,
                             const double *hbar_old, const double *water_flux, const double *zbar, const double *CFL_z,
                             double min_hnode, int lzstar_lev, double *Wvel, double *hnode_new)
{
    const int nl = m->nl, L = nl - 1, lz = lzstar_lev;
    double max_dhbar2distr[256], distrib_dhbar[256], cumsum_maxdhbar[256];
#define Z_W(nz, n) Wvel[(size_t)((n) - 1) * nl + ((nz) - 1)]
#define Z_H(nz, n) m->hnode[(size_t)((n) - 1) * L + ((nz) - 1)]
#define Z_HN(nz, n) hnode_new[(size_t)((n) - 1) * L + ((nz) - 1)]
#define Z_C(nz, n) CFL_z[(size_t)((n) - 1) * nl + ((nz) - 1)]
#define ZB(nz) zbar[(nz) - 1]
    for (int n = 1; n <= m->myDim_nod2D; ++n) {
        const int nzmin = m->ulevels_nod2D[n - 1];
        int nzmax = nlevels_nod2D_min[n - 1] - 1;
        if (nzmin != 1) continue;                                                   /* :2354 cavity: like linfs */
        const double dhbar_total = hbar[n - 1] - hbar_old[n - 1];                   /* :2357 */
        if (dhbar_total < 0.0 && Z_H(nzmin, n) + dhbar_total <= (ZB(nzmin) - ZB(nzmin + 1)) * min_hnode) {   /* :2367 */
            for (int k = 1; k <= lz; ++k) {                                         /* :2374-2383 */
                double v = (ZB(nzmin + k - 1) - ZB(nzmin + k)) * min_hnode - Z_H(nzmin + k - 1, n);
                if (v >= 0.0) v = 0.0;
                if (Z_C(nzmin + k - 1, n) >= 0.95) v = 0.0;
                max_dhbar2distr[k - 1] = v;
            }
            cumsum_maxdhbar[0] = max_dhbar2distr[0];                                /* :2397 */
            for (int k = 2; k <= lz; ++k) cumsum_maxdhbar[k - 1] = max_dhbar2distr[k - 1] + max_dhbar2distr[k - 2];   /* :2398 */
            int nz = 2147483647;                                                    /* minval of an empty pack = huge */
            for (int k = lz; k >= 1; --k) if (cumsum_maxdhbar[k - 1] < dhbar_total) nz = k;                           /* :2399 */
            if (nz > lz) nz = lz;                                                   /* :2400 */
            for (int k = 0; k < lz; ++k) distrib_dhbar[k] = 0.0;
            double dhbar_rest = dhbar_total;
            nzmax = nz < nzmax - 1 ? nz : nzmax - 1;                                /* :2411 */
            for (nz = 1; nz <= nzmax; ++nz) {                                       /* :2412-2416 */
                distrib_dhbar[nz - 1] = dhbar_rest > max_dhbar2distr[nz - 1] ? dhbar_rest : max_dhbar2distr[nz - 1];
                dhbar_rest = dhbar_rest - distrib_dhbar[nz - 1];
                dhbar_rest = 0.0 < dhbar_rest ? 0.0 : dhbar_rest;
            }
            double distrib_dhbar_int = 0.0;
            for (nz = nzmax; nz >= 1; --nz) {                                       /* :2438-2449 */
                distrib_dhbar_int = distrib_dhbar_int + distrib_dhbar[nz - 1];
                Z_W(nzmin + nz - 1, n) = Z_W(nzmin + nz - 1, n) - distrib_dhbar_int / dt;
                Z_HN(nzmin + nz - 1, n) = Z_H(nzmin + nz - 1, n) + distrib_dhbar[nz - 1];
            }
        } else {
            int any_ne = 0, last_ne = -2147483647;                                  /* maxval of an empty pack = -huge */
            if (dhbar_total > 0.0)
                for (int k = 2; k <= lz; ++k)                                       /* :2462-2463 */
                    if (Z_H(nzmin + k - 1, n) != ZB(nzmin + k - 1) - ZB(nzmin + k)) any_ne = 1;
            if (dhbar_total > 0.0 && any_ne) {
                for (int k = 1; k <= lz; ++k)                                       /* :2471 */
                    max_dhbar2distr[k - 1] = (ZB(nzmin + k - 1) - ZB(nzmin + k)) - Z_H(nzmin + k - 1, n);
                max_dhbar2distr[0] = 1000.0;                                        /* :2475 */
                for (int k = 1; k <= lz; ++k)                                       /* :2482 */
                    if (Z_H(nzmin + k - 1, n) != ZB(nzmin + k - 1) - ZB(nzmin + k)) last_ne = k;
                int nz = last_ne;
                nzmax = nz < nzmax - 1 ? nz : nzmax - 1;                            /* :2488 */
                double dhbar_rest = dhbar_total, distrib_dhbar_int = 0.0;
                for (nz = nzmax; nz >= 1; --nz) {                                   /* :2496-2510 */
                    const double d = dhbar_rest < max_dhbar2distr[nz - 1] ? dhbar_rest : max_dhbar2distr[nz - 1];
                    dhbar_rest = dhbar_rest - d;
                    dhbar_rest = 0.0 > dhbar_rest ? 0.0 : dhbar_rest;
                    distrib_dhbar_int = distrib_dhbar_int + d;
                    Z_W(nzmin + nz - 1, n) = Z_W(nzmin + nz - 1, n) - distrib_dhbar_int / dt;
                    Z_HN(nzmin + nz - 1, n) = Z_H(nzmin + nz - 1, n) + d;
                }
            } else {                                                                /* :2519-2520 the normal zlevel case */
                Z_W(nzmin, n) = Z_W(nzmin, n) - dhbar_total / dt;
                Z_HN(nzmin, n) = Z_H(nzmin, n) + dhbar_total;
            }
        }
        Z_W(nzmin, n) = Z_W(nzmin, n) - water_flux[n - 1];                          /* :2527 */
    }
#undef Z_W
#undef Z_H
#undef Z_HN
#undef Z_C
#undef ZB
}
