// fesom_host.cpp -- see fesom_host.hpp.  Thin: fills the C-ABI descriptors from the mirrored derived types and forwards.
#include "fesom_host.hpp"

#include <cstdio>
#include <cstdlib>

namespace fesom {

thread_local int mstep = 0;
thread_local bool ldiag_DVD = false;

[[noreturn]] void par_ex(int mype, int abort_code)
{
    // the reference: MPI_ABORT(MPI_COMM_FESOM, 1) when abort is present, otherwise barrier + finalize + stop
    std::fflush(stdout);
    if (mype == 0) std::fprintf(stderr, "par_ex: run finished unexpectedly (abort %d)\n", abort_code);
    std::exit(abort_code ? abort_code : 1);
}

static void check(int rc, const t_partit& partit)
{
    if (rc == ADV_OK) return;
    if (partit.mype == 0) std::fprintf(stderr, "fesom_adv_b200: %s\n", adv_last_error());
    par_ex(partit.mype, 1);                               // src/oce_adv_tra_driver.F90:351-353
}

void oce_adv_tra_fct_init(t_tracer_work& twork, t_partit& partit, const t_mesh& mesh, int device, int max_tracers)
{
    adv_mesh_desc_t d{};
    d.nl = mesh.nl;
    d.myDim_nod2D = partit.myDim_nod2D; d.eDim_nod2D = partit.eDim_nod2D;
    d.myDim_elem2D = partit.myDim_elem2D; d.eDim_elem2D = partit.eDim_elem2D;
    d.myDim_edge2D = partit.myDim_edge2D;
    d.nod_in_elem2D_ld = mesh.nod_in_elem2D_ld;
    d.edges = mesh.edges.data(); d.edge_tri = mesh.edge_tri.data(); d.elem2D_nodes = mesh.elem2D_nodes.data();
    d.nod_in_elem2D = mesh.nod_in_elem2D.data(); d.nod_in_elem2D_num = mesh.nod_in_elem2D_num.data();
    d.nlevels = mesh.nlevels.data(); d.ulevels = mesh.ulevels.data();
    d.nlevels_nod2D = mesh.nlevels_nod2D.data(); d.ulevels_nod2D = mesh.ulevels_nod2D.data();
    d.edge_cross_dxdy = mesh.edge_cross_dxdy.data(); d.edge_dxdy = mesh.edge_dxdy.data(); d.elem_cos = mesh.elem_cos.data();
    d.area = mesh.area.data(); d.areasvol = mesh.areasvol.data();
    d.nboundary_lay = twork.nboundary_lay.empty() ? nullptr : twork.nboundary_lay.data();
    const com_struct& c = partit.com_nod2D;
    d.mype = partit.mype; d.npes = partit.npes;
    d.rPEnum = c.rPEnum; d.rPE = c.rPE.data(); d.rptr = c.rptr.data(); d.rlist = c.rlist.data();
    d.sPEnum = c.sPEnum; d.sPE = c.sPE.data(); d.sptr = c.sptr.data(); d.slist = c.slist.data();
    check(adv_ctx_create(&twork.b200, &d, device, max_tracers), partit);
}

void par_init_local(std::vector<t_tracer_work*>& tworks, t_partit& partit0)
{
    std::vector<adv_ctx_t*> ctxs;
    for (t_tracer_work* w : tworks) ctxs.push_back(w->b200);
    check(adv_ctx_comm_init_local(ctxs.data(), (int)ctxs.size()), partit0);
}

void oce_adv_tra_fct_final(t_tracer_work& twork)
{
    if (twork.b200) adv_ctx_destroy(twork.b200);
    twork.b200 = nullptr;
}

void do_oce_adv_tra(WP dt, const WP* vel, const WP* w, const WP* wi, const WP* we, int tr_num, t_dyn& dynamics,
                    t_tracer& tracers, t_partit& partit, const t_mesh& mesh)
{
    t_tracer_work& wk = tracers.work;
    adv_state_desc_t st{};
    st.uv = vel; st.w = w; st.w_e = we; st.w_i = wi;
    st.helem = mesh.helem.data(); st.hnode = mesh.hnode.data(); st.hnode_new = mesh.hnode_new.data();
    st.zbar_3d_n = mesh.zbar_3d_n.data(); st.Z_3d_n = mesh.Z_3d_n.data();
    st.zbar_n_bot = mesh.zbar_n_bot.empty() ? nullptr : mesh.zbar_n_bot.data();
    st.use_wsplit = dynamics.use_wsplit ? 1 : 0;
    // the reference refreshes the state once per step (src/oce_ale_tracer.F90:260-262) and then loops over the tracers
    // (:280-312): a repeated call with the same mstep and the same arrays is a no-op
    check(adv_ctx_set_state_step(wk.b200, &st, ADV_HOST, (int64_t)mstep), partit);

    t_tracer_data& td = tracers.data[(size_t)tr_num - 1];
    const size_t nLN = (size_t)(mesh.nl - 1) * (size_t)(partit.myDim_nod2D + partit.eDim_nod2D);
    adv_tracer_desc_t t{};
    t.values = td.values.data(); t.valuesAB = td.valuesAB.data();
    t.edge_up_dn_grad = wk.edge_up_dn_grad.empty() ? nullptr : wk.edge_up_dn_grad.data();
    t.del_ttf_advhoriz = wk.del_ttf_advhoriz.data(); t.del_ttf_advvert = wk.del_ttf_advvert.data();
    t.tra_adv_hor = td.tra_adv_hor.c_str(); t.tra_adv_ver = td.tra_adv_ver.c_str(); t.tra_adv_lim = td.tra_adv_lim.c_str();
    t.tra_adv_ph = td.tra_adv_ph; t.tra_adv_pv = td.tra_adv_pv;
    if (td.ltra_diag && !wk.tra_advhoriz.empty()) {       // the tracer's slice: the tracer index is the slowest one
        t.tra_advhoriz = wk.tra_advhoriz.data() + (size_t)(tr_num - 1) * nLN;
        t.tra_advvert = wk.tra_advvert.data() + (size_t)(tr_num - 1) * nLN;
    }
    if (ldiag_DVD && tr_num <= 2 && !wk.dvd_trflx_hor.empty()) {       // src/oce_adv_tra_driver.F90:263, :395
        t.dvd_trflx_hor = wk.dvd_trflx_hor.data() + (size_t)(tr_num - 1) * (size_t)(mesh.nl - 1) * (size_t)partit.myDim_edge2D;
        t.dvd_trflx_ver = wk.dvd_trflx_ver.data() + (size_t)(tr_num - 1) * (size_t)mesh.nl * (size_t)partit.myDim_nod2D;
    }
    check(adv_do_oce_adv_tra(wk.b200, dt, 1, &t, ADV_HOST), partit);
}

}  // namespace fesom
