// fesom_host.hpp -- the compiled host side above the C ABI: a C++ mirror of the reference's interface for the tracer-advection
// path.  The reference's seam is the Fortran external procedure
//     subroutine do_oce_adv_tra(dt, vel, w, wi, we, tr_num, dynamics, tracers, partit, mesh)   (src/oce_adv_tra_driver.F90:46)
// over the derived types t_mesh / t_partit / t_tracer / t_dyn.  A Fortran compiler exists neither in this image nor on the GPU
// box, so the ISO_C_BINDING shim (fesom2_b200/fortran/oce_adv_tra_b200.F90) cannot be compiled here; this file is the same
// thin layer in C++: the types carry the components the path reads, under the reference's names and shapes (column-major,
// level fastest, 1-based index VALUES), and the procedures keep the reference's names, argument order and error behaviour.
// Everything is HOST memory owned by the caller, like the allocatable components of the Fortran types; the library
// page-locks and uploads what it is handed (ADV_HOST).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "fesom_adv_b200.h"

namespace fesom {

using WP = double;                                        // src/oce_modules.F90:17

struct com_struct {                                       // src/MOD_PARTIT.F90:18-33
    int rPEnum = 0, sPEnum = 0;
    std::vector<int32_t> rPE, rptr, rlist, sPE, sptr, slist;
};

struct t_partit {                                         // src/MOD_PARTIT.F90:42-120 (what the path reads)
    int mype = 0, npes = 1;
    int myDim_nod2D = 0, eDim_nod2D = 0, myDim_elem2D = 0, eDim_elem2D = 0, myDim_edge2D = 0;
    com_struct com_nod2D;
};

struct t_mesh {                                           // src/MOD_MESH.F90:22-175 (what the path reads)
    int nl = 0;
    int nod_in_elem2D_ld = 0;                             // first extent of nod_in_elem2D
    std::vector<int32_t> edges, edge_tri;                 // (2, myDim_edge2D)
    std::vector<int32_t> elem2D_nodes;                    // (3, myDim_elem2D+eDim)
    std::vector<int32_t> nod_in_elem2D, nod_in_elem2D_num;
    std::vector<int32_t> nlevels, ulevels, nlevels_nod2D, ulevels_nod2D;
    std::vector<WP> edge_cross_dxdy, edge_dxdy, elem_cos; // (4,E) (2,E) (T)
    std::vector<WP> area, areasvol;                       // (nl, Nh)
    std::vector<WP> helem;                                // (nl-1, T)
    std::vector<WP> hnode, hnode_new, Z_3d_n;             // (nl-1, Nh)
    std::vector<WP> zbar_3d_n;                            // (nl, Nh)
    std::vector<WP> zbar_n_bot;                           // (Nh)
};

struct t_tracer_data {                                    // src/MOD_TRACER.F90:10-45
    std::vector<WP> values, valuesAB;                     // (nl-1, Nh)
    bool ltra_diag = true;                                // :25
    std::string tra_adv_hor = "MFCT", tra_adv_ver = "QR4C", tra_adv_lim = "FCT";   // :15-17 defaults
    WP tra_adv_ph = 1.0, tra_adv_pv = 1.0;
    int ID = 0;
};

struct t_tracer_work {                                    // src/MOD_TRACER.F90:47-109 (what the caller sees of it)
    std::vector<WP> del_ttf, del_ttf_advhoriz, del_ttf_advvert;   // (nl-1, Nh)
    std::vector<WP> edge_up_dn_grad;                      // (4, nl-1, myDim_edge2D): ONE set, refilled per tracer by the caller
    std::vector<int32_t> nboundary_lay;                   // (Nh)
    std::vector<WP> tra_advhoriz, tra_advvert;            // (nl-1, Nh, num_tracers)
    std::vector<WP> dvd_trflx_hor, dvd_trflx_ver;         // (nl-1, E, 2), (nl, N, 2); empty unless ldiag_DVD
    adv_ctx_t* b200 = nullptr;                            // the device context that replaces the seven FCT work arrays
};

struct t_tracer {
    int num_tracers = 0;
    std::vector<t_tracer_data> data;
    t_tracer_work work;
};

struct t_dyn {                                            // src/MOD_DYN.F90:65- (uv :68, w / w_e / w_i :75, use_wsplit :156)
    std::vector<WP> uv;                                   // (2, nl-1, T)
    std::vector<WP> w, w_e, w_i;                          // (nl, Nh)
    bool use_wsplit = false;
};

// o_PARAM mstep (src/oce_modules.F90:23), the model's step counter, and diagnostics ldiag_DVD (src/gen_modules_diag.F90:101).
// thread_local: a host that runs several ranks as threads of one process (one context each, adv_ctx_comm_init_local)
// gives every rank its own copy, like the module variables of separate MPI processes.
extern thread_local int mstep;
extern thread_local bool ldiag_DVD;

// par_ex(comm, mype, abort) (src/gen_modules_partitioning.F90:87-123): the reference's way out of an unrecoverable error
[[noreturn]] void par_ex(int mype, int abort_code);

// oce_adv_tra_fct_init(twork, partit, mesh) (src/oce_adv_tra_fct.F90:35-67): the reference allocates the FCT work arrays
// here; this one creates the device context (mesh upload, gather lists, work arrays for `max_tracers` per call)
void oce_adv_tra_fct_init(t_tracer_work& twork, t_partit& partit, const t_mesh& mesh, int device = 0, int max_tracers = 2);
void oce_adv_tra_fct_final(t_tracer_work& twork);
// N ranks inside one process (one host thread per rank, every twork initialised with its own partit): links their contexts
// so that the halo exchanges inside do_oce_adv_tra run as device-to-device copies (adv_ctx_comm_init_local); with MPI ranks
// in separate processes the host broadcasts adv_comm_unique_id and calls adv_ctx_comm_init instead
void par_init_local(std::vector<t_tracer_work*>& tworks, t_partit& partit0);

// do_oce_adv_tra(dt, vel, w, wi, we, tr_num, dynamics, tracers, partit, mesh) (src/oce_adv_tra_driver.F90:46-490): one tracer
// per call, tr_num 1-based, accumulating into tracers%work%del_ttf_advhoriz / del_ttf_advvert; the state arrays are uploaded
// (and the edge volume flux computed) by the first call of a model step only (mstep).  Unknown scheme: message + par_ex.
void do_oce_adv_tra(WP dt, const WP* vel, const WP* w, const WP* wi, const WP* we, int tr_num, t_dyn& dynamics,
                    t_tracer& tracers, t_partit& partit, const t_mesh& mesh);

}  // namespace fesom
