// dwarf_tracer.cpp -- the reference's tracer-advection dwarf (dwarf/dwarf_tracer/dwarf_ini/fesom.F90:85-128) as a compiled
// host program above the C++ mirror of the reference interface (fesom_host.hpp): no Python, no torch -- pageable host arrays
// in, host arrays out, the call sequence of the model (state refreshed once per step, do_oce_adv_tra once per tracer).
//
//     dwarf_tracer_b200 <case.bin> <result.bin>
//     dwarf_tracer_b200 --restart <npepath> <result.bin> [nsteps [dt]]
//     dwarf_tracer_b200 --ranks N <case-prefix> <result-prefix>
//
// The third form runs N ranks of a partitioned mesh as N host threads of this process (files <case-prefix>.<rank> written
// with the partition's local numbering and com_nod2D lists): one context per rank, linked by the library's in-process
// communicator, the halo exchanges inside do_oce_adv_tra on the device, exchange_nod(values) of the dwarf's epilogue
// between the ranks' host arrays.
//
// The second form is the dwarf as the reference ships it: it reads the derived-type binary restarts of rank 0 of 1 from
// <npepath> (read_all_bin_restarts, fesom.F90:64), advects tracer 1 ten times with dt = 1.e-3, accumulates del_ttf over the
// iterations exactly as fesom.F90:105-125 does, and writes values / del_ttf / del_ttf_advhoriz / del_ttf_advvert.
//
// case.bin (little endian, written by tests/test_gpu_host_cpp.py): 'FADV', int32 {nl, myDim_nod2D, eDim_nod2D, myDim_elem2D,
// eDim_elem2D, myDim_edge2D, nod_in_elem2D_ld, num_tracers, nsteps, use_wsplit, ldiag_DVD}, double dt, then the arrays of
// t_mesh / t_dyn / t_tracer in the order read below, reference layouts.  One rank (the reference's restart reader and
// MPI are outside the path; the N-rank launch sequence is covered through the C ABI by tests/test_gpu_multirank.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <numeric>

#include "fesom_host.hpp"
#include "fesom_restart.hpp"

#include <pthread.h>
#include <stdexcept>
#include <thread>

using namespace fesom;

template <class T>
static void rd(FILE* f, std::vector<T>& v, size_t n)
{
    v.resize(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "dwarf_tracer: short read\n"); std::exit(2); }
}
template <class T>
static void wr(FILE* f, const std::vector<T>& v)
{
    if (!v.empty() && std::fwrite(v.data(), sizeof(T), v.size(), f) != v.size()) { std::fprintf(stderr, "dwarf_tracer: short write\n"); std::exit(2); }
}

// dwarf/dwarf_tracer/dwarf_ini/fesom.F90:64-128, literally (tracer 1, del_ttf carried over the iterations)
static int dwarf_from_restarts(const char* npepath, const char* result, int nsteps, double dt)
{
    t_mesh mesh;
    t_partit partit;
    t_dyn dyn;
    t_tracer tracers;
    try {
        read_all_bin_restarts(npepath, 0, 1, partit, mesh, dyn, tracers);                     // :64
    } catch (const std::exception& ex) {
        std::fprintf(stderr, " -ERROR-> %s\n", ex.what());
        par_ex(0, 1);
    }
    t_tracer_work& wk = tracers.work;
    const size_t L = (size_t)mesh.nl - 1, N = (size_t)partit.myDim_nod2D, Nh = N + (size_t)partit.eDim_nod2D;
    oce_adv_tra_fct_init(wk, partit, mesh, 0, 1);
    std::vector<WP>& val = tracers.data[0].values;
    for (int i = 1; i <= nsteps; ++i) {                                                       // :85
        mstep = i;
        std::fill(wk.del_ttf_advhoriz.begin(), wk.del_ttf_advhoriz.end(), 0.0);               // :88-95
        std::fill(wk.del_ttf_advvert.begin(), wk.del_ttf_advvert.end(), 0.0);
        do_oce_adv_tra(dt, dyn.uv.data(), dyn.w.data(), dyn.w_i.data(), dyn.w_e.data(), 1, dyn, tracers, partit, mesh);   // :97
        const auto mm = std::minmax_element(val.begin(), val.end());
        std::printf("%.17g %.17g %.17g\n", *mm.first, *mm.second, std::accumulate(val.begin(), val.end(), 0.0));          // :99
        for (size_t n = 0; n < Nh; ++n)                                                       // :105-113
            for (int nz = mesh.ulevels_nod2D[n]; nz <= mesh.nlevels_nod2D[n] - 1; ++nz) {
                const size_t o = n * L + (size_t)(nz - 1);
                wk.del_ttf[o] = wk.del_ttf[o] + wk.del_ttf_advhoriz[o] + wk.del_ttf_advvert[o];
            }
        for (size_t n = 0; n < N; ++n)                                                        // :116-124
            for (int nz = mesh.ulevels_nod2D[n]; nz <= mesh.nlevels_nod2D[n] - 1; ++nz) {
                const size_t o = n * L + (size_t)(nz - 1);
                val[o] = val[o] + wk.del_ttf[o] / mesh.hnode_new[o];
            }
        // exchange_nod(values): one rank (:127)
    }
    oce_adv_tra_fct_final(wk);
    FILE* g = std::fopen(result, "wb");
    if (!g) { std::perror(result); return 2; }
    wr(g, val); wr(g, wk.del_ttf); wr(g, wk.del_ttf_advhoriz); wr(g, wk.del_ttf_advvert);
    std::fclose(g);
    return 0;
}

// what read_all_bin_restarts delivered, as text (no GPU call): dimensions, scheme strings, sums of the arrays
static int dump_restart(const char* npepath)
{
    t_mesh mesh;
    t_partit partit;
    t_dyn dyn;
    t_tracer tracers;
    try {
        read_all_bin_restarts(npepath, 0, 1, partit, mesh, dyn, tracers);
    } catch (const std::exception& ex) {
        std::fprintf(stderr, " -ERROR-> %s\n", ex.what());
        return 1;
    }
    auto sumd = [](const std::vector<WP>& a) { return std::accumulate(a.begin(), a.end(), 0.0); };
    auto sumi = [](const std::vector<int32_t>& a) { return (long long)std::accumulate(a.begin(), a.end(), 0LL); };
    std::printf("dims %d %d %d %d %d %d %d %d %d\n", mesh.nl, partit.myDim_nod2D, partit.eDim_nod2D, partit.myDim_elem2D, partit.eDim_elem2D,
                partit.myDim_edge2D, mesh.nod_in_elem2D_ld, tracers.num_tracers, dyn.use_wsplit ? 1 : 0);
    std::printf("ints %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", sumi(mesh.edges), sumi(mesh.edge_tri), sumi(mesh.elem2D_nodes),
                sumi(mesh.nod_in_elem2D), sumi(mesh.nod_in_elem2D_num), sumi(mesh.nlevels), sumi(mesh.ulevels), sumi(mesh.nlevels_nod2D),
                sumi(mesh.ulevels_nod2D), sumi(tracers.work.nboundary_lay));
    std::printf("reals %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", sumd(mesh.edge_cross_dxdy),
                sumd(mesh.edge_dxdy), sumd(mesh.elem_cos), sumd(mesh.area), sumd(mesh.areasvol), sumd(mesh.helem), sumd(mesh.hnode),
                sumd(mesh.hnode_new), sumd(mesh.zbar_3d_n), sumd(mesh.Z_3d_n), sumd(mesh.zbar_n_bot), sumd(dyn.uv), sumd(dyn.w), sumd(dyn.w_e),
                sumd(dyn.w_i));
    for (const t_tracer_data& t : tracers.data)
        std::printf("tracer %d %s %s %s %.17g %.17g %.17g %.17g\n", t.ID, t.tra_adv_hor.c_str(), t.tra_adv_ver.c_str(), t.tra_adv_lim.c_str(),
                    t.tra_adv_ph, t.tra_adv_pv, sumd(t.values), sumd(t.valuesAB));
    std::printf("work %.17g %zu\n", sumd(tracers.work.edge_up_dn_grad), tracers.work.del_ttf.size());
    return 0;
}

struct Case {
    t_mesh mesh;
    t_partit partit;
    t_dyn dyn;
    t_tracer tracers;
    std::vector<std::vector<WP>> grad;                    // what the caller's fill_up_dn_grad produced for each tracer
    int nsteps = 1;
    double dt = 0.0;
    bool dvd = false;
};

// 'FADV': one rank; 'FADW': + int32 {mype, npes, rPEnum, sPEnum, len(rlist), len(slist)} and the com_nod2D lists after the header
static void read_case(const char* path, Case& c)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    char magic[4];
    int32_t h[11];
    if (std::fread(magic, 1, 4, f) != 4 || (std::memcmp(magic, "FADV", 4) != 0 && std::memcmp(magic, "FADW", 4) != 0) ||
        std::fread(h, 4, 11, f) != 11 || std::fread(&c.dt, 8, 1, f) != 1) {
        std::fprintf(stderr, "dwarf_tracer: bad header in %s\n", path);
        std::exit(2);
    }
    t_mesh& mesh = c.mesh; t_partit& partit = c.partit; t_dyn& dyn = c.dyn; t_tracer& tracers = c.tracers;
    mesh.nl = h[0];
    partit.myDim_nod2D = h[1]; partit.eDim_nod2D = h[2]; partit.myDim_elem2D = h[3]; partit.eDim_elem2D = h[4]; partit.myDim_edge2D = h[5];
    mesh.nod_in_elem2D_ld = h[6];
    tracers.num_tracers = h[7];
    c.nsteps = h[8];
    dyn.use_wsplit = h[9] != 0;
    c.dvd = h[10] != 0;
    if (magic[3] == 'W') {
        int32_t p[6];
        if (std::fread(p, 4, 6, f) != 6) { std::fprintf(stderr, "dwarf_tracer: bad partition header\n"); std::exit(2); }
        partit.mype = p[0]; partit.npes = p[1];
        com_struct& cm = partit.com_nod2D;
        cm.rPEnum = p[2]; cm.sPEnum = p[3];
        rd(f, cm.rPE, (size_t)p[2]); rd(f, cm.rptr, (size_t)p[2] + 1); rd(f, cm.rlist, (size_t)p[4]);
        rd(f, cm.sPE, (size_t)p[3]); rd(f, cm.sptr, (size_t)p[3] + 1); rd(f, cm.slist, (size_t)p[5]);
    }
    const size_t nl = mesh.nl, L = nl - 1, N = partit.myDim_nod2D, Nh = N + partit.eDim_nod2D, T = partit.myDim_elem2D + partit.eDim_elem2D, E = partit.myDim_edge2D;
    const size_t ntr = tracers.num_tracers;
    rd(f, mesh.edges, 2 * E); rd(f, mesh.edge_tri, 2 * E); rd(f, mesh.elem2D_nodes, 3 * T);
    rd(f, mesh.nod_in_elem2D, (size_t)mesh.nod_in_elem2D_ld * Nh); rd(f, mesh.nod_in_elem2D_num, Nh);
    rd(f, mesh.nlevels, T); rd(f, mesh.ulevels, T); rd(f, mesh.nlevels_nod2D, Nh); rd(f, mesh.ulevels_nod2D, Nh);
    rd(f, tracers.work.nboundary_lay, Nh);
    rd(f, mesh.edge_cross_dxdy, 4 * E); rd(f, mesh.edge_dxdy, 2 * E); rd(f, mesh.elem_cos, T);
    rd(f, mesh.area, nl * Nh); rd(f, mesh.areasvol, nl * Nh);
    rd(f, dyn.uv, 2 * L * T); rd(f, dyn.w, nl * Nh); rd(f, dyn.w_e, nl * Nh); rd(f, dyn.w_i, nl * Nh);
    rd(f, mesh.helem, L * T); rd(f, mesh.hnode, L * Nh); rd(f, mesh.hnode_new, L * Nh);
    rd(f, mesh.zbar_3d_n, nl * Nh); rd(f, mesh.Z_3d_n, L * Nh); rd(f, mesh.zbar_n_bot, Nh);
    tracers.data.resize(ntr);
    c.grad.resize(ntr);
    for (size_t k = 0; k < ntr; ++k) {
        t_tracer_data& td = tracers.data[k];
        char s[24];
        int32_t diag;
        double pp[2];
        if (std::fread(s, 1, 24, f) != 24 || std::fread(pp, 8, 2, f) != 2 || std::fread(&diag, 4, 1, f) != 1) { std::fprintf(stderr, "dwarf_tracer: bad tracer header\n"); std::exit(2); }
        auto str = [&](int o) { std::string x(s + o, 8); x.erase(x.find_last_not_of(' ') + 1); return x; };
        td.tra_adv_hor = str(0); td.tra_adv_ver = str(8); td.tra_adv_lim = str(16);
        td.tra_adv_ph = pp[0]; td.tra_adv_pv = pp[1];
        td.ltra_diag = diag != 0;
        td.ID = (int)k + 1;
        rd(f, td.values, L * Nh); rd(f, td.valuesAB, L * Nh); rd(f, c.grad[k], 4 * L * E);
    }
    std::fclose(f);
    t_tracer_work& wk = tracers.work;
    wk.del_ttf.assign(L * Nh, 0.0); wk.del_ttf_advhoriz.assign(L * Nh, 0.0); wk.del_ttf_advvert.assign(L * Nh, 0.0);
    wk.edge_up_dn_grad.assign(4 * L * E, 0.0);
    wk.tra_advhoriz.assign(L * Nh * ntr, 0.0); wk.tra_advvert.assign(L * Nh * ntr, 0.0);      // src/oce_setup_step.F90:505-507
    if (c.dvd) { wk.dvd_trflx_hor.assign(L * E * 2, 0.0); wk.dvd_trflx_ver.assign(nl * N * 2, 0.0); }
}

// exchange_nod(values) between ranks that are threads of this process (the library's exchanges work on device arrays; the
// dwarf's epilogue updates HOST arrays): halo columns := the owner's columns, along com_nod2D (gen_halo_exchange.F90:432-517)
struct HostComm {
    pthread_barrier_t bar;
    std::vector<Case*> ranks;
    std::vector<WP*> field;
};
static void host_exchange_nod(HostComm& hc, int me, WP* field, size_t nlev)
{
    if (hc.ranks.size() < 2) return;
    hc.field[(size_t)me] = field;
    pthread_barrier_wait(&hc.bar);
    const com_struct& cm = hc.ranks[(size_t)me]->partit.com_nod2D;
    for (int i = 0; i < cm.rPEnum; ++i) {
        const int pe = cm.rPE[(size_t)i];
        const com_struct& pc = hc.ranks[(size_t)pe]->partit.com_nod2D;
        int j = 0;
        while (pc.sPE[(size_t)j] != me) ++j;
        const int n = cm.rptr[(size_t)i + 1] - cm.rptr[(size_t)i];
        for (int k = 0; k < n; ++k) {
            const size_t dst = (size_t)cm.rlist[(size_t)(cm.rptr[(size_t)i] - 1 + k)] - 1, src = (size_t)pc.slist[(size_t)(pc.sptr[(size_t)j] - 1 + k)] - 1;
            std::memcpy(field + dst * nlev, hc.field[(size_t)pe] + src * nlev, nlev * sizeof(WP));
        }
    }
    pthread_barrier_wait(&hc.bar);
}

// the dwarf iteration of one rank (fesom.F90:85-128 with the model's tracer loop and per-step del_ttf reset)
static void dwarf_loop(Case& c, HostComm* hc, std::vector<std::vector<WP>>& last_h, std::vector<std::vector<WP>>& last_v)
{
    t_mesh& mesh = c.mesh; t_partit& partit = c.partit; t_dyn& dyn = c.dyn; t_tracer& tracers = c.tracers;
    t_tracer_work& wk = tracers.work;
    const size_t L = (size_t)mesh.nl - 1, N = partit.myDim_nod2D, ntr = tracers.num_tracers;
    ldiag_DVD = c.dvd;
    last_h.resize(ntr); last_v.resize(ntr);
    for (int i = 1; i <= c.nsteps; ++i) {                  // fesom.F90:85
        mstep = i;
        for (int tr = 1; tr <= (int)ntr; ++tr) {           // src/oce_ale_tracer.F90:280-312
            std::fill(wk.del_ttf_advhoriz.begin(), wk.del_ttf_advhoriz.end(), 0.0);           // fesom.F90:88-95
            std::fill(wk.del_ttf_advvert.begin(), wk.del_ttf_advvert.end(), 0.0);
            wk.edge_up_dn_grad = c.grad[(size_t)tr - 1];   // the gradient calls of init_tracers_AB, src/oce_tracer_mod.F90:125-141
            do_oce_adv_tra(c.dt, dyn.uv.data(), dyn.w.data(), dyn.w_i.data(), dyn.w_e.data(), tr, dyn, tracers, partit, mesh);   // :97
            std::vector<WP>& val = tracers.data[(size_t)tr - 1].values;
            if (partit.mype == 0) {                        // :99
                const auto mm = std::minmax_element(val.begin(), val.end());
                std::printf("%d %d %.17g %.17g %.17g\n", i, tr, *mm.first, *mm.second, std::accumulate(val.begin(), val.end(), 0.0));
            }
            // :105-125 with del_ttf reset per step as in the model (src/oce_tracer_mod.F90:28-34)
            for (size_t n = 0; n < N; ++n) {
                const int nzmax = mesh.nlevels_nod2D[n] - 1, nzmin = mesh.ulevels_nod2D[n];
                for (int nz = nzmin; nz <= nzmax; ++nz) {
                    const size_t o = n * L + (size_t)(nz - 1);
                    wk.del_ttf[o] = 0.0 + wk.del_ttf_advhoriz[o] + wk.del_ttf_advvert[o];
                    val[o] = val[o] + wk.del_ttf[o] / mesh.hnode_new[o];
                }
            }
            if (hc) host_exchange_nod(*hc, partit.mype, val.data(), L);                       // exchange_nod(values), :127
            if (i == c.nsteps) { last_h[(size_t)tr - 1] = wk.del_ttf_advhoriz; last_v[(size_t)tr - 1] = wk.del_ttf_advvert; }
        }
    }
}

static void write_result(const char* path, const Case& c, const std::vector<std::vector<WP>>& last_h, const std::vector<std::vector<WP>>& last_v)
{
    FILE* g = std::fopen(path, "wb");
    if (!g) { std::perror(path); std::exit(2); }
    const t_tracer_work& wk = c.tracers.work;
    for (size_t k = 0; k < (size_t)c.tracers.num_tracers; ++k) { wr(g, c.tracers.data[k].values); wr(g, last_h[k]); wr(g, last_v[k]); }
    wr(g, wk.tra_advhoriz); wr(g, wk.tra_advvert); wr(g, wk.dvd_trflx_hor); wr(g, wk.dvd_trflx_ver);
    std::fclose(g);
}

// N ranks as N host threads
static int dwarf_ranks(int n, const char* case_prefix, const char* result_prefix)
{
    std::vector<Case> cases((size_t)n);
    HostComm hc;
    hc.field.assign((size_t)n, nullptr);
    pthread_barrier_init(&hc.bar, nullptr, (unsigned)n);
    std::vector<t_tracer_work*> tworks;
    for (int r = 0; r < n; ++r) {
        read_case((std::string(case_prefix) + "." + std::to_string(r)).c_str(), cases[(size_t)r]);
        if (cases[(size_t)r].partit.mype != r || cases[(size_t)r].partit.npes != n) { std::fprintf(stderr, "dwarf_tracer: %s.%d is not rank %d of %d\n", case_prefix, r, r, n); return 2; }
        hc.ranks.push_back(&cases[(size_t)r]);
        oce_adv_tra_fct_init(cases[(size_t)r].tracers.work, cases[(size_t)r].partit, cases[(size_t)r].mesh, 0, 1);
        tworks.push_back(&cases[(size_t)r].tracers.work);
    }
    par_init_local(tworks, cases[0].partit);
    std::vector<std::vector<std::vector<WP>>> lh((size_t)n), lv((size_t)n);
    std::vector<std::thread> th;
    for (int r = 0; r < n; ++r) th.emplace_back([&, r] { dwarf_loop(cases[(size_t)r], &hc, lh[(size_t)r], lv[(size_t)r]); });
    for (std::thread& t : th) t.join();
    for (int r = 0; r < n; ++r) {
        oce_adv_tra_fct_final(cases[(size_t)r].tracers.work);
        write_result((std::string(result_prefix) + "." + std::to_string(r)).c_str(), cases[(size_t)r], lh[(size_t)r], lv[(size_t)r]);
    }
    pthread_barrier_destroy(&hc.bar);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc == 3 && std::strcmp(argv[1], "--dump-restart") == 0) return dump_restart(argv[2]);
    if (argc >= 4 && std::strcmp(argv[1], "--restart") == 0)
        return dwarf_from_restarts(argv[2], argv[3], argc > 4 ? std::atoi(argv[4]) : 10, argc > 5 ? std::atof(argv[5]) : 1.e-3);
    if (argc == 5 && std::strcmp(argv[1], "--ranks") == 0) return dwarf_ranks(std::atoi(argv[2]), argv[3], argv[4]);
    if (argc != 3) { std::fprintf(stderr, "usage: %s case.bin result.bin | --restart npepath result.bin [nsteps [dt]] | --ranks N case-prefix result-prefix\n", argv[0]); return 2; }
    Case c;
    read_case(argv[1], c);
    oce_adv_tra_fct_init(c.tracers.work, c.partit, c.mesh, 0, 1);
    std::vector<std::vector<WP>> last_h, last_v;
    dwarf_loop(c, nullptr, last_h, last_v);
    oce_adv_tra_fct_final(c.tracers.work);
    write_result(argv[2], c, last_h, last_v);
    return 0;
}
