/* oracle/adv_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, single thread per rank) of the reference's tracer-advection path
 * (FESOM/fesom2 @ e3c3d9d): src/oce_adv_tra_driver.F90, src/oce_adv_tra_hor.F90,
 * src/oce_adv_tra_ver.F90, src/oce_adv_tra_fct.F90.  Same loop order, same expression order,
 * column-major (nz fastest) arrays with 1-based index semantics via macros.
 *
 * PARITY UNPINNED: the reference cannot be compiled in this environment (no Fortran compiler,
 * no MPI) and ships no golden vector / known-answer test for this path.
 * The restatement is pinned only by invariants and by an independently written NumPy
 * restatement (oracle/numpy_ref.py, tests/test_oracle_pinning.py, tests/test_multirank_cpu.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#ifndef ADV_ORACLE_H
#define ADV_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* scheme codes (the reference dispatches on strings, oce_adv_tra_driver.F90:343-379) */
enum { ORA_HOR_UPW1 = 0, ORA_HOR_MUSCL = 1, ORA_HOR_MFCT = 2 };
enum { ORA_VER_UPW1 = 0, ORA_VER_QR4C = 1, ORA_VER_PPM = 2, ORA_VER_CDIFF = 3 };
enum { ORA_LIM_NONE = 0, ORA_LIM_FCT = 1 };

/* slice of t_mesh / t_partit the path reads (MOD_MESH.F90:22-175, MOD_PARTIT.F90:35-120).
 * All index values 1-based; all arrays column-major with the Fortran shapes noted. */
typedef struct {
    int nl;                     /* interfaces; layers L = nl-1 */
    int myDim_nod2D, eDim_nod2D;
    int myDim_elem2D, eDim_elem2D;
    int myDim_edge2D;
    int nod_in_elem_ld;         /* leading dim of nod_in_elem2D */
    const int *edges;           /* (2,E) */
    const int *edge_tri;        /* (2,E)  <=0 : no element */
    const int *elem2D_nodes;    /* (3,T) */
    const int *nod_in_elem2D;   /* (ld,N) */
    const int *nod_in_elem2D_num; /* (N) */
    const int *nlevels, *ulevels;           /* (T+eT) */
    const int *nlevels_nod2D, *ulevels_nod2D; /* (Nh) */
    const double *edge_cross_dxdy; /* (4,E) */
    const double *edge_dxdy;       /* (2,E) */
    const double *elem_cos;        /* (T+eT) */
    const double *area;            /* (nl,Nh) */
    const double *areasvol;        /* (nl,Nh) */
    /* ALE state refreshed every step */
    const double *helem;           /* (L,T+eT) */
    const double *hnode, *hnode_new; /* (L,Nh) */
    const double *zbar_3d_n;       /* (nl,Nh) */
    const double *Z_3d_n;          /* (L,Nh) */
    const double *zbar_n_bot;      /* (Nh)  only adv_tra_vert_impl */
} ora_mesh_t;

/* t_tracer_work scratch + one tracer (MOD_TRACER.F90:10-109) */
typedef struct {
    double *fct_LO;          /* (L,Nh) */
    double *adv_flux_hor;    /* (L,E)  */
    double *adv_flux_ver;    /* (nl,N) */
    double *fct_ttf_min, *fct_ttf_max; /* (L,Nh) */
    double *fct_plus, *fct_minus;      /* (L,Nh) */
    double *tvert_max, *tvert_min;     /* (L,Nh) automatic arrays of oce_tra_adv_fct */
    double *AUX;             /* (4,L,E) FCT scratch; the reference aliases edge_up_dn_grad */
    const int *nboundary_lay;/* (Nh) */
    /* optional diagnostics of the CURRENT tracer (NULL = off), oce_adv_tra_driver.F90:221-229,:263-296,:307-318,:395-458,:464-488 */
    double *tra_advhoriz, *tra_advvert;     /* (L,Nh) tracers%data(tr_num)%ltra_diag (default .true., MOD_TRACER.F90:25) */
    double *dvd_trflx_hor, *dvd_trflx_ver;  /* (L,E), (nl,N) ldiag_DVD .and. tr_num <= 2 */
} ora_work_t;

typedef void (*ora_exchange_fn)(void *user, double *field, int nlev); /* exchange_nod3D */

/* the individual routines ------------------------------------------------------------- */
void ora_adv_tra_hor_upw1(const ora_mesh_t *m, const double *vel, const double *ttf,
                          double *flux, int init_zero);
void ora_adv_tra_hor_muscl(const ora_mesh_t *m, const double *vel, const double *ttf,
                           double num_ord, double *flux, const double *edge_up_dn_grad,
                           const int *nboundary_lay, int init_zero);
void ora_adv_tra_hor_mfct(const ora_mesh_t *m, const double *vel, const double *ttf,
                          double num_ord, double *flux, const double *edge_up_dn_grad,
                          int init_zero);
void ora_adv_tra_ver_upw1(const ora_mesh_t *m, const double *w, const double *ttf,
                          double *flux, int init_zero);
void ora_adv_tra_ver_qr4c(const ora_mesh_t *m, const double *w, const double *ttf,
                          double num_ord, double *flux, int init_zero);
void ora_adv_tra_vert_ppm(const ora_mesh_t *m, double dt, const double *w, const double *ttf,
                          double *flux, int init_zero);
void ora_adv_tra_ver_cdiff(const ora_mesh_t *m, const double *w, const double *ttf,
                           double *flux, int init_zero);
void ora_adv_tra_vert_impl(const ora_mesh_t *m, double dt, const double *w, double *ttf);
void ora_oce_tra_adv_fct(const ora_mesh_t *m, double dt, const double *ttf, const double *lo,
                         double *adf_h, double *adf_v, double *fct_ttf_min, double *fct_ttf_max,
                         double *fct_plus, double *fct_minus, double *AUX,
                         double *tvert_max, double *tvert_min,
                         ora_exchange_fn xchg, void *user);
void ora_oce_tra_adv_flux2dtracer(const ora_mesh_t *m, double dt, double *dttf_h, double *dttf_v,
                                  double *flux_h, double *flux_v, int use_lo,
                                  const double *ttf, const double *lo);

/* ---- the producer of edge_up_dn_grad ------------------------------
 * tracer_gradient_elements (src/oce_tracer_mod.F90:146-188): tr_xy(1:2,nz,elem) for elem <= myDim_elem2D,
 * layers ulevels(elem) .. nlevels(elem)-1.  `sum()` of the three products is taken left to right.
 * tr_xy is (2, nl-1, >= myDim_elem2D); entries outside the loop bounds are left untouched. */
void ora_tracer_gradient_elements(int nl, int myDim_elem2D, const int *elem2D_nodes, const int *nlevels,
                                  const int *ulevels, const double *gradient_sca /* (6,T) */,
                                  const double *ttf /* (nl-1,Nh) */, double *tr_xy);
/* fill_up_dn_grad (src/oce_muscl_adv.F90:356-525): edge_up_dn_grad(1:4,nz,edge) for edge <= myDim_edge2D
 * from tr_xy (which the caller has halo-exchanged, src/oce_tracer_mod.F90:140) -- the up/down-wind
 * triangle's gradient on the levels both end nodes share, the area-weighted mean over the wet elements
 * around the end node elsewhere.  Entries the reference does not write are left untouched. */
void ora_fill_up_dn_grad(int nl, int myDim_edge2D, const int *edges, const int *edge_up_dn_tri /* (2,E) */,
                         const int *nod_in_elem2D, int ld, const int *nod_in_elem2D_num,
                         const int *nlevels, const int *ulevels,
                         const int *nlevels_nod2D, const int *ulevels_nod2D,
                         const int *nlevels_nod2D_min, const int *ulevels_nod2D_max,
                         const double *elem_area, const double *tr_xy, double *edge_up_dn_grad /* (4,nl-1,E) */);

/* find_up_downwind_triangles (src/oce_muscl_adv.F90:162-352), single rank: edge_up_dn_tri(1:2,edge), 0 = none;
 * the last element around the end node whose sector contains the (reversed) edge direction wins. */
void ora_find_up_downwind_triangles(int myDim_edge2D, const int *edges, const int *elem2D_nodes,
                                    const int *nod_in_elem2D, int ld, const int *nod_in_elem2D_num,
                                    const double *coord_nod2D /* (2,N) */, double cyclic_length,
                                    int *edge_up_dn_tri /* (2,E) */);

/* ---- SURVEY.md section 8f row 3, oracle first: the continuity part of vert_vel_ale
 * (src/oce_ale.F90:2164-2310, linfs, no Fer_GM): edge transports scattered to the two end nodes in edge
 * order (element 1 then element 2 of each edge), bottom-up cumulative sum over the node column, division
 * by area.  Wvel is (nl, Nh); only owned nodes are completed (the reference exchanges it afterwards). */
void ora_vert_vel_ale_core(const ora_mesh_t *m, const double *UV /* (2,nl-1,T) */, double *Wvel);

/* do_oce_adv_tra (oce_adv_tra_driver.F90:46-490) for one tracer.  Returns 0, or 1 for an unknown
 * scheme (the reference calls par_ex there). */
int ora_do_oce_adv_tra(const ora_mesh_t *m, ora_work_t *wk, double dt,
                       const double *vel, const double *w, const double *wi, const double *we,
                       int use_wsplit,
                       const double *ttf, const double *ttfAB, const double *edge_up_dn_grad,
                       int hor, int ver, int lim, double opth, double optv,
                       double *dttf_h, double *dttf_v,
                       ora_exchange_fn xchg, void *user);

/* ---- multi-rank runner: one pthread per partition, shared-memory halo copies in place of MPI
 * (the "MPI CPU path" stand-in of SURVEY.md section 8d).  Each rank runs `nsteps` dwarf iterations
 * (dwarf/dwarf_tracer/dwarf_ini/fesom.F90:85-128) over its `ntr` tracers. */
typedef struct {
    ora_mesh_t mesh;
    ora_work_t work;
    const double *vel, *w, *wi, *we;
    int use_wsplit;
    int ntr;
    double **values;          /* [ntr] (L,Nh)  updated in place when update_values != 0 */
    double **valuesAB;        /* [ntr] (L,Nh) */
    double **edge_up_dn_grad; /* [ntr] (4,L,E) */
    double **dttf_h, **dttf_v;/* [ntr] (L,Nh) del_ttf_advhoriz / advvert */
    const int *hor, *ver, *lim; /* [ntr] */
    const double *opth, *optv;  /* [ntr] */
    /* halo description: for halo node k (0-based, k < eDim_nod2D): owner rank and the owner's
     * 1-based local node index */
    const int *halo_owner, *halo_owner_idx;
    /* optional per-tracer diagnostics, each NULL or [ntr] pointers (entries may be NULL) */
    double **tra_advhoriz, **tra_advvert, **dvd_trflx_hor, **dvd_trflx_ver;
} ora_rank_t;

/* returns wall seconds of the timed region (all ranks, barrier to barrier). mode: 0 = only
 * do_oce_adv_tra per step (the timed region of SURVEY 8d); 1 = full dwarf iteration (zero
 * del_ttf, advect, values += (dttf_h+dttf_v)/hnode_new, exchange values). */
double ora_run_ranks(int nranks, ora_rank_t *ranks, double dt, int nsteps, int mode);

/* compute_CFLz (src/oce_ale.F90:2906-2998, no print) and compute_Wvel_split (:3001-3049), all myDim+eDim nodes;
 * every array (nl, Nh) */
/* vert_vel_ale, zstar correction (src/oce_ale.F90:2539-2603): Wvel (nl,Nh) and hnode_new (nl-1,Nh) updated in place */
void ora_vert_vel_ale_zstar(const ora_mesh_t *m, double dt, const int *nlevels_nod2D_min, const double *hbar,
                            const double *hbar_old, const double *water_flux, double *Wvel, double *hnode_new);
/* vert_vel_ale, zlevel correction (src/oce_ale.F90:2336-2538): zbar = mesh%zbar (nl), CFL_z = the previous step's (nl,Nh) */
void ora_vert_vel_ale_zlevel(const ora_mesh_t *m, double dt, const int *nlevels_nod2D_min, const double *hbar,
                             const double *hbar_old, const double *water_flux, const double *zbar, const double *CFL_z,
                             double min_hnode, int lzstar_lev, double *Wvel, double *hnode_new);
void ora_compute_cflz(const ora_mesh_t *m, double dt, const double *Wvel, double *CFL_z);
void ora_compute_wvel_split(const ora_mesh_t *m, int use_wsplit, double wsplit_maxcfl, const double *Wvel,
                            const double *CFL_z, double *Wvel_e, double *Wvel_i);

#ifdef __cplusplus
}
#endif
#endif
