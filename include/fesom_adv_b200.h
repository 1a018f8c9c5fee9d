/* include/fesom_adv_b200.h -- C ABI of the B200 tracer-advection library (libfesom_adv_b200.so).
 *
 * Drop-in boundary for the reference's Fortran procedures (FESOM/fesom2 @ e3c3d9d); every entry
 * point names the reference interface it replaces.  The reference has no FFI for this path: the
 * seam is the external-procedure interface `do_oce_adv_tra` (src/oce_adv_tra_driver.F90:1-21) and
 * `exchange_nod` (src/gen_halo_exchange.F90:2731-2768).  A Fortran host binds these functions with
 * ISO_C_BINDING, passing c_loc() of the allocatable components of t_mesh / t_partit / t_tracer /
 * t_dyn plus scalar dimensions (see INTEGRATION.md and fesom2_b200/fortran/oce_adv_tra_b200.F90).
 *
 * Conventions
 *   - All arrays keep the reference's memory layout: column-major, level index fastest,
 *     e.g. values(nl-1, myDim_nod2D+eDim_nod2D), edge_up_dn_grad(4, nl-1, myDim_edge2D).
 *   - All index VALUES are 1-based local indices, as stored by the Fortran host.
 *   - Reals are IEEE binary64 (WP = real64, src/oce_modules.F90:17); integers are 32 bit.
 *   - Every function returns 0 on success, a negative ADV_E* code otherwise (the reference has no
 *     status codes: an unknown scheme ends in par_ex -> MPI_ABORT, oce_adv_tra_driver.F90:351-353;
 *     the Fortran shim calls par_ex when it sees ADV_ESCHEME).
 *   - `where` tells whether the data pointers of a call are HOST or DEVICE pointers.  Host
 *     pointers are copied H2D / D2H inside the call; device pointers are used in place (the
 *     OpenACC `!$ACC HOST_DATA USE_DEVICE` convention of the reference's GPU build).
 *   - No CPU fallback exists: without a CUDA device every call fails with ADV_ECUDA.
 */
#ifndef FESOM_ADV_B200_H
#define FESOM_ADV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADV_OK        0
#define ADV_EINVAL   -1   /* bad argument / inconsistent mesh */
#define ADV_ECUDA    -2   /* CUDA runtime error (adv_last_error() has the text) */
#define ADV_ESCHEME  -3   /* unknown tra_adv_hor / tra_adv_ver / tra_adv_lim string */
#define ADV_ENCCL    -4   /* NCCL error */
#define ADV_ESTATE   -5   /* call order violated (e.g. no state set) */

#define ADV_HOST      0
#define ADV_DEVICE    1

typedef struct adv_ctx adv_ctx_t;

/* Static mesh + partition description: the slice of t_mesh (src/MOD_MESH.F90:22-175, views in
 * src/associate_mesh_ass.h:9-79) and t_partit (src/MOD_PARTIT.F90:35-120) that the path reads.
 * All pointers are HOST pointers; the arrays are copied and may be released after the call. */
typedef struct {
    int32_t nl;                 /* mesh%nl: number of interfaces, layers = nl-1 */
    int32_t myDim_nod2D, eDim_nod2D;
    int32_t myDim_elem2D, eDim_elem2D; /* vel/helem/nlevels/... carry myDim+eDim element columns */
    int32_t myDim_edge2D;
    int32_t nod_in_elem2D_ld;   /* leading dimension of nod_in_elem2D */
    const int32_t *edges;             /* (2, myDim_edge2D)                          */
    const int32_t *edge_tri;          /* (2, myDim_edge2D), <= 0: no element        */
    const int32_t *elem2D_nodes;      /* (3, myDim_elem2D)                          */
    const int32_t *nod_in_elem2D;     /* (ld, myDim_nod2D)                          */
    const int32_t *nod_in_elem2D_num; /* (myDim_nod2D)                              */
    const int32_t *nlevels, *ulevels;             /* (myDim_elem2D [+eDim])         */
    const int32_t *nlevels_nod2D, *ulevels_nod2D; /* (myDim_nod2D+eDim_nod2D)       */
    const double  *edge_cross_dxdy;   /* (4, myDim_edge2D)                          */
    const double  *edge_dxdy;         /* (2, myDim_edge2D)                          */
    const double  *elem_cos;          /* (myDim_elem2D [+eDim])                     */
    const double  *area;              /* (nl, myDim_nod2D+eDim_nod2D)               */
    const double  *areasvol;          /* (nl, myDim_nod2D+eDim_nod2D)               */
    const int32_t *nboundary_lay;     /* (myDim_nod2D+eDim_nod2D) t_tracer_work, may be NULL if MUSCL unused */
    /* com_nod2D (src/MOD_PARTIT.F90:18-33); all NULL / 0 on a single rank */
    int32_t mype, npes;
    int32_t rPEnum; const int32_t *rPE, *rptr, *rlist;
    int32_t sPEnum; const int32_t *sPE, *sptr, *slist;
} adv_mesh_desc_t;

/* Per-step ocean state: what the reference refreshes with `!$ACC UPDATE DEVICE` every step
 * (src/oce_ale_tracer.F90:260-262): dynamics%uv,w,w_e,w_i and mesh%helem,hnode,hnode_new,
 * zbar_3d_n,Z_3d_n.  zbar_n_bot may be NULL unless use_wsplit. */
typedef struct {
    const double *uv;         /* (2, nl-1, myDim_elem2D+eDim_elem2D) */
    const double *w, *w_e, *w_i; /* (nl, Nh) */
    const double *helem;      /* (nl-1, myDim_elem2D+eDim_elem2D)    */
    const double *hnode, *hnode_new; /* (nl-1, Nh) */
    const double *zbar_3d_n;  /* (nl, Nh)   */
    const double *Z_3d_n;     /* (nl-1, Nh) */
    const double *zbar_n_bot; /* (Nh)       */
    int32_t use_wsplit;       /* dynamics%use_wsplit (src/MOD_DYN.F90:156) */
} adv_state_desc_t;

/* One tracer of the batch: t_tracer_data (src/MOD_TRACER.F90:10-45) + its work slices. */
typedef struct {
    const double *values;          /* ttf   (nl-1, Nh) in    */
    const double *valuesAB;        /* ttfAB (nl-1, Nh) in    */
    const double *edge_up_dn_grad; /* (4, nl-1, myDim_edge2D) in (NOT clobbered, unlike the
                                      reference which reuses it as FCT scratch,
                                      oce_adv_tra_driver.F90:383-384); may be NULL for UPW1 */
    double *del_ttf_advhoriz;      /* (nl-1, Nh) inout, accumulated */
    double *del_ttf_advvert;       /* (nl-1, Nh) inout, accumulated */
    const char *tra_adv_hor;       /* "MUSCL" | "MFCT" | "UPW1"          (blank padded or NUL terminated) */
    const char *tra_adv_ver;       /* "QR4C" | "CDIFF" | "PPM" | "UPW1"  */
    const char *tra_adv_lim;       /* "FCT" or anything else = no limiter */
    double tra_adv_ph, tra_adv_pv; /* num_ord of the horizontal / vertical scheme */
    /* tracers%data(tr_num)%ltra_diag (default .true., src/MOD_TRACER.F90:25): the tracer's slices of
     * tracers%work%tra_advhoriz / tra_advvert, (nl-1, Nh) each, or NULL for ltra_diag = .false.  On return the wet layers of
     * the OWNED nodes hold what src/oce_adv_tra_driver.F90:221-229, :307-318 and :464-488 leave there -- with FCT the
     * low-order tendency plus (del_ttf_advhoriz | del_ttf_advvert after the call) / hnode_new, otherwise that quotient alone; everything else
     * (halo nodes, where the reference stores partial edge sums nobody reads, and dry layers) is left untouched. */
    double *tra_advhoriz, *tra_advvert;
    /* ldiag_DVD (src/gen_modules_diag.F90:101, default .false.; the reference does it for tr_num <= 2 only): the tracer's
     * slices of tracers%work%dvd_trflx_hor (nl-1, myDim_edge2D) and dvd_trflx_ver (nl, myDim_nod2D), or NULL.  On return
     * they hold the total flux through every mid-edge face / scalar-cell face: with FCT the low-order flux plus the limited
     * antidiffusive flux, otherwise the high-order flux (src/oce_adv_tra_driver.F90:263-296, :395-458).  Every entry is
     * written; the layers a boundary edge has above a cavity top are written as zero (the reference keeps fluxes of
     * cells it never uses there). */
    double *dvd_trflx_hor, *dvd_trflx_ver;
} adv_tracer_desc_t;

/* --- life cycle ------------------------------------------------------------------------------ */
/* Replaces oce_adv_tra_fct_init (src/oce_adv_tra_fct.F90:35-67) and the `!$ACC ENTER DATA` of the
 * static mesh (src/fesom_module.F90:759-805): uploads the mesh, builds the node->edge gather
 * lists, FCT cluster lists and halo pack lists, allocates all work arrays for `max_tracers`
 * tracers per batched call. */
int adv_ctx_create(adv_ctx_t **ctx, const adv_mesh_desc_t *mesh, int device, int max_tracers);
int adv_ctx_destroy(adv_ctx_t *ctx);
const char *adv_last_error(void);

/* Replaces par_init/init_mpi_types for this path (src/gen_modules_partitioning.F90:41-85,
 * :416-514): one NCCL communicator over the npes ranks.  Rank 0 calls adv_comm_unique_id and the
 * host broadcasts the 128 bytes (MPI_Bcast in the Fortran host, torch.distributed in the tests). */
int adv_comm_unique_id(char id[128]);
int adv_ctx_comm_init(adv_ctx_t *ctx, const char id[128]);

/* In-process alternative to adv_ctx_comm_init: the n contexts of ONE process (ctxs[r] created with mype = r,
 * npes = n; one per GPU, or several on one GPU) exchange their halos with direct device-to-device copies --
 * peer copies over NVLink between GPUs -- ordered by CUDA events, without NCCL.  Every context must then be
 * driven by its own host thread, all threads making the same sequence of calls (each exchange is a
 * rendezvous of the n threads, like the MPI exchange it replaces). */
int adv_ctx_comm_init_local(adv_ctx_t *const *ctxs, int n);

/* --- per step ------------------------------------------------------------------------------- */
/* Hands over the state of a model step.  HOST pointers: uploaded now (asynchronously; the arrays are page-locked
 * with cudaHostRegister the first time they are seen, see adv_ctx_set_host_register).  DEVICE pointers: used in
 * place.  RULE: the volume flux Q(nz,edge) derived from uv/helem is cached from the first tracer call after
 * adv_ctx_set_state until the next adv_ctx_set_state[_step]; a caller that changes uv/helem/w... -- also in place
 * on the device -- MUST call adv_ctx_set_state again before the next adv_do_oce_adv_tra (with DEVICE pointers the
 * call copies nothing). */
int adv_ctx_set_state(adv_ctx_t *ctx, const adv_state_desc_t *st, int where);
/* Same with the model's step counter (mstep, src/oce_modules.F90:23): a repeated call with the same `step`, the
 * same `where` and the same pointers is a no-op.  The reference refreshes the state once per step
 * (src/oce_ale_tracer.F90:260-262) and then calls do_oce_adv_tra once per tracer (:280-312); the drop-in wrapper
 * of that signature calls this function every time, so that upload and Q are shared by the tracer loop. */
int adv_ctx_set_state_step(adv_ctx_t *ctx, const adv_state_desc_t *st, int where, int64_t step);
/* on = 0: never page-lock the caller's HOST arrays (default 1: cudaHostRegister on first use, undone by
 * adv_ctx_destroy; ranges that are already pinned are left alone). */
int adv_ctx_set_host_register(adv_ctx_t *ctx, int on);

/* Replaces `do_oce_adv_tra(dt, vel, w, wi, we, tr_num, dynamics, tracers, partit, mesh)`
 * (src/oce_adv_tra_driver.F90:46-490) for a BATCH of ntr tracers (ntr = 1 reproduces one Fortran
 * call), including its two internal halo exchanges (:335 and src/oce_adv_tra_fct.F90:413).
 * vel/w/wi/we are those of the last adv_ctx_set_state.  Accumulates into del_ttf_advhoriz /
 * del_ttf_advvert.  Blocking: results are complete on return. */
int adv_do_oce_adv_tra(adv_ctx_t *ctx, double dt, int ntr, const adv_tracer_desc_t *tr, int where);

/* Same, asynchronous on the context's stream, DEVICE pointers only (dwarf convention: operands
 * stay resident in one `!$ACC DATA` region, dwarf/dwarf_tracer/dwarf_ini/fesom.F90:71-130). */
int adv_do_oce_adv_tra_async(adv_ctx_t *ctx, double dt, int ntr, const adv_tracer_desc_t *tr);
int adv_ctx_synchronize(adv_ctx_t *ctx);

/* Stream ordering.  The library computes on its own non-blocking stream (adv_ctx_stream), which is NOT ordered
 * against any stream of the caller.  DEVICE-pointer inputs that the caller produces asynchronously (OpenACC
 * async queues -- acc_get_cuda_stream --, torch streams) must be ordered before the call with
 * adv_ctx_wait_for(ctx, stream): the context waits for everything submitted so far to `stream` (0 = the legacy
 * default stream).  Outputs of the *_async entry points are ordered into a caller stream with
 * adv_ctx_signal(ctx, stream): `stream` waits for the context's work submitted so far; or call
 * adv_ctx_synchronize.  The blocking entry points with HOST pointers need neither. */
int adv_ctx_wait_for(adv_ctx_t *ctx, void *stream);
int adv_ctx_signal(adv_ctx_t *ctx, void *stream);

/* Replaces `exchange_nod(field, partit)` for nfields 3-D node fields with nlev levels
 * (src/gen_halo_exchange.F90:432-633): packed-halo NCCL send/recv.  DEVICE pointers. */
int adv_exchange_nod(adv_ctx_t *ctx, int nfields, double *const *fields, int nlev);

/* Replaces `exchange_elem(field, partit)` for nfields element fields with nwords doubles per element column
 * and myDim+eDim+eXDim columns (src/gen_halo_exchange.F90:769-980, the com_elem2D_full branch), e.g. tr_xy with
 * nwords = 2*(nl-1) (src/oce_tracer_mod.F90:140).  Needs com_elem2D_full from adv_ctx_set_gradient_mesh.
 * DEVICE pointers. */
int adv_exchange_elem(adv_ctx_t *ctx, int nfields, double *const *fields, int nwords);

/* The dwarf's epilogue (dwarf/dwarf_tracer/dwarf_ini/fesom.F90:105-127) with del_ttf reset per
 * step as in the model (src/oce_tracer_mod.F90:28-34): values += (advhoriz+advvert)/hnode_new on
 * owned nodes, then exchange_nod(values).  DEVICE pointers. */
int adv_update_values(adv_ctx_t *ctx, int ntr, double *const *values,
                      const double *const *del_ttf_advhoriz, const double *const *del_ttf_advvert);

/* The vertical velocities the path consumes, for the linear free surface: the continuity
 * part of `vert_vel_ale` (src/oce_ale.F90:2164-2310, which_ALE = 'linfs', no Fer_GM) on the owned nodes, its
 * exchange_nod(Wvel) (:2654), `compute_CFLz` (:2906-2998 without the diagnostic print) and `compute_Wvel_split`
 * (:3001-3049) on all myDim+eDim nodes: from uv / helem / hnode_new of the last adv_ctx_set_state to w, w_e, w_i and
 * cfl_z (each (nl, Nh); cfl_z may be NULL).  dt = g_config dt, wsplit_maxcfl = dynamics%wsplit_maxcfl.  Entries the
 * reference leaves alone in w_e / w_i stay untouched.  DEVICE pointers, asynchronous on the context's stream.
 * The result becomes the path's input with the next adv_ctx_set_state. */
int adv_vert_vel_ale(adv_ctx_t *ctx, double dt, int use_wsplit, double wsplit_maxcfl,
                     double *w, double *w_e, double *w_i, double *cfl_z);
/* The same for which_ALE = 'zstar' (the reference's default, config/namelist.config:71): after the continuity part the
 * elevation change hbar - hbar_old is distributed over the layers above the shallowest bottom around each owned,
 * cavity-free node -- Wvel and hnode_new, src/oce_ale.F90:2539-2603 --, the surface fresh-water flux closes the
 * continuity at the top, then exchange_nod(Wvel) and exchange_nod(hnode_new) (:2654-2655); compute_CFLz uses the new
 * hnode_new.  All arrays are DEVICE arrays. */
typedef struct {
    const double  *hbar, *hbar_old;      /* (Nh) mesh%hbar, mesh%hbar_old                              */
    const double  *water_flux;           /* (Nh) o_ARRAYS water_flux                                   */
    const int32_t *nlevels_nod2D_min;    /* (Nh) mesh%nlevels_nod2D_min                                */
    double        *hnode_new;            /* (nl-1, Nh) inout: the stretched layers of the owned nodes are rewritten from hnode */
} adv_zstar_desc_t;
int adv_vert_vel_ale_zstar(adv_ctx_t *ctx, double dt, int use_wsplit, double wsplit_maxcfl, const adv_zstar_desc_t *z,
                           double *w, double *w_e, double *w_i, double *cfl_z);
/* The same for which_ALE = 'zlevel' (src/oce_ale.F90:2336-2538): the elevation change goes into the surface layer of every
 * owned, cavity-free node; where that layer would become thinner than min_hnode times its rest thickness the change is
 * spread over the first lzstar_lev layers (local zstar, :2367-2449, with the previous step's CFL_z as a brake), and a
 * later rise refills the squeezed subsurface layers first (:2461-2510); then the fresh-water flux, both exchanges,
 * compute_CFLz with the new thickness and compute_Wvel_split as above.  `cfl_z` is IN/OUT here and must not be NULL: on
 * entry the previous step's CFL_z (the reference reads its module array before compute_CFLz rewrites it).  min_hnode and
 * lzstar_lev are the namelist values (src/gen_modules_config.F90:71,:75: 0.5 and 4).  All arrays are DEVICE arrays. */
typedef struct {
    const double  *hbar, *hbar_old;      /* (Nh) mesh%hbar, mesh%hbar_old                              */
    const double  *water_flux;           /* (Nh) o_ARRAYS water_flux                                   */
    const int32_t *nlevels_nod2D_min;    /* (Nh) mesh%nlevels_nod2D_min                                */
    double        *hnode_new;            /* (nl-1, Nh) inout: the layers that take part are rewritten from hnode */
    const double  *zbar;                 /* (nl) mesh%zbar: the rest interfaces                        */
    double         min_hnode;            /* g_config min_hnode                                         */
    int32_t        lzstar_lev;           /* g_config lzstar_lev (1 .. 16)                              */
} adv_zlevel_desc_t;
int adv_vert_vel_ale_zlevel(adv_ctx_t *ctx, double dt, int use_wsplit, double wsplit_maxcfl, const adv_zlevel_desc_t *z,
                            double *w, double *w_e, double *w_i, double *cfl_z);

/* The prologue of the tracer step, `init_tracers_AB(tr_num, tracers, partit, mesh)`
 * (src/oce_tracer_mod.F90:13-123) without its gradient calls, for ntr tracers: zeroes del_ttf /
 * del_ttf_advhoriz / del_ttf_advvert (each may be NULL), sets valuesAB from the Adams-Bashforth
 * extrapolation of order ab_order (2: -(0.5+eps)*valuesold(1) + (1.5+eps)*values; 3: (5*valuesold(2)
 * - 16*valuesold(1) + 23*values)/12) and rotates valuesold.  valuesold has the reference layout
 * (ab_order-1, nl-1, Nh), first index fastest.  epsilon = o_PARAM epsilon (src/oce_modules.F90:105).
 * Any other order returns ADV_EINVAL (the reference stops with par_ex).  DEVICE pointers, asynchronous
 * on the context's stream. */
int adv_init_tracers_AB(adv_ctx_t *ctx, int ntr, int ab_order, double epsilon,
                        const double *const *values, double *const *valuesold, double *const *valuesAB,
                        double *const *del_ttf, double *const *del_ttf_advhoriz, double *const *del_ttf_advvert);

/* --- the producer of edge_up_dn_grad (the caller's side of the path) ----
 * Static inputs of tracer_gradient_elements / fill_up_dn_grad that adv_ctx_create does not get.  Index
 * values 1-based, reference layouts.  The arrays are copied to the device. */
typedef struct {
    int32_t n_elem;                 /* elements tr_xy covers: myDim_elem2D + eDim_elem2D + eXDim_elem2D        */
    int32_t n_nod_in_elem;          /* nodes nod_in_elem2D(_num) cover: every end node of a local edge          */
    int32_t nod_in_elem2D_ld;
    const int32_t *nod_in_elem2D;       /* (ld, n_nod_in_elem)                                                  */
    const int32_t *nod_in_elem2D_num;   /* (n_nod_in_elem)                                                      */
    const int32_t *nlevels, *ulevels;   /* (n_elem)                                                             */
    const int32_t *edge_up_dn_tri;      /* (2, myDim_edge2D)  t_tracer_work%edge_up_dn_tri, 0 = none            */
    const int32_t *nlevels_nod2D_min;   /* (myDim_nod2D + eDim_nod2D)                                           */
    const int32_t *ulevels_nod2D_max;   /* (myDim_nod2D + eDim_nod2D)                                           */
    const double *gradient_sca;         /* (6, myDim_elem2D)                                                    */
    const double *elem_area;            /* (n_elem)                                                             */
    /* com_elem2D_full (src/MOD_PARTIT.F90:18-33,:62; built by communication_elemn, src/gen_comm.F90:222-527): the
     * halo of tr_xy, i.e. what exchange_elem3D moves for an array with myDim+eDim+eXDim element columns
     * (src/gen_halo_exchange.F90:769-980).  All NULL / 0 on a single rank. */
    int32_t rPEnum; const int32_t *rPE, *rptr, *rlist;
    int32_t sPEnum; const int32_t *sPE, *sptr, *slist;
} adv_gradient_mesh_desc_t;
int adv_ctx_set_gradient_mesh(adv_ctx_t *ctx, const adv_gradient_mesh_desc_t *g);

/* Replaces `tracer_gradient_elements(ttf, partit, mesh)` (src/oce_tracer_mod.F90:146-188) for ntr tracers:
 * tr_xy(1:2, nz, elem) for elem <= myDim_elem2D.  tr_xy[i] is (2, nl-1, n_elem); entries the reference does
 * not write are left untouched.  DEVICE pointers, asynchronous on the context's stream.  The halo part of
 * tr_xy (exchange_elem, src/oce_tracer_mod.F90:140) remains the caller's job on more than one rank. */
int adv_tracer_gradient_elements(adv_ctx_t *ctx, int ntr, const double *const *ttf, double *const *tr_xy);

/* Replaces `fill_up_dn_grad(twork, partit, mesh)` (src/oce_muscl_adv.F90:356-525) for ntr tracers:
 * edge_up_dn_grad(1:4, nz, edge) for edge <= myDim_edge2D from the (exchanged) tr_xy.  Entries the reference
 * does not write are left untouched.  DEVICE pointers, asynchronous on the context's stream. */
int adv_fill_up_dn_grad(adv_ctx_t *ctx, int ntr, const double *const *tr_xy, double *const *edge_up_dn_grad);
/* After adv_ctx_set_gradient_mesh a tracer descriptor may also carry edge_up_dn_grad = NULL (gradient-based
 * scheme): adv_do_oce_adv_tra then runs the two routines above on `values` itself before the step -- with
 * adv_exchange_elem(tr_xy) between them on more than one rank -- which saves the caller's gradient sweeps and
 * the 4 E (nl-1) words of host-to-device copy per tracer on the HOST-pointer path. */

/* --- introspection (tests, profiling) -------------------------------------------------------- */
/* Copies an internal work array of tracer slot `slot` to a HOST buffer.  name is one of
 * "fct_LO" (nl-1,Nh), "adv_flux_hor" (nl-1,E), "adv_flux_ver" (nl,N), "fct_plus", "fct_minus"
 * (nl-1,Nh), "edge_volflux" (nl-1,E; slot ignored). */
int adv_ctx_get_work(adv_ctx_t *ctx, const char *name, int slot, double *out);
/* number of kernels launched by this context since creation */
int64_t adv_ctx_launch_count(const adv_ctx_t *ctx);
/* raw CUDA stream (cudaStream_t) the context computes on */
void *adv_ctx_stream(adv_ctx_t *ctx);
/* device-side duration [ms] of the last adv_do_oce_adv_tra[_async] call, measured with CUDA
 * events on the context's compute stream (valid after synchronisation) */
int adv_ctx_last_elapsed_ms(adv_ctx_t *ctx, float *ms);
/* Halo-exchange accounting of the last adv_do_oce_adv_tra[_async] call with an FCT tracer on more than one rank
 * (synchronises): bytes this rank sent in the call's two exchanges (fct_LO: 0; fct_plus/fct_minus: 1), the
 * duration of each on the communication stream (pack + transfer), and the time the compute stream had to wait for
 * it after running out of interior work (the exposed part; 0 = fully hidden). */
int adv_ctx_halo_stats(adv_ctx_t *ctx, int64_t *bytes_sent, float comm_ms[2], float exposed_ms[2]);
/* per-phase kernel time of the last call [ms]: 0 edge fluxes (+Q), 1 LO solution, 2 bounds/R, 3 update; valid only when profiling was switched on with adv_ctx_set_profiling(ctx, 1) */
int adv_ctx_set_profiling(adv_ctx_t *ctx, int on);
int adv_ctx_phase_ms(adv_ctx_t *ctx, float ms[8]);

/* Self-test of the library's exact reciprocal-based division against the IEEE `/` on `count`
 * pseudo-random operand pairs (mode 0: divisor 6, 1: divisor 3, 2: random divisor); writes the
 * number of results that differ (must be 0). */
int adv_selftest_div(uint64_t count, uint64_t seed, int mode, uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* FESOM_ADV_B200_H */
