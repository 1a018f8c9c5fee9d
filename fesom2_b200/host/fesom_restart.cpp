// fesom_restart.cpp -- see fesom_restart.hpp
#include "fesom_restart.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>

namespace fesom {
namespace {

// item kinds: i default integer, r real(WP), l default logical, c<N> character(N), ai<R> / ar<R> write_bin_array of rank R
// (extents first, all 0 = unallocated), si write1d_int_static (count first)
struct Field { const char* kind; const char* name; };

const Field T_MESH[] = {
    {"i", "nod2D"}, {"r", "ocean_area"}, {"r", "ocean_areawithcav"}, {"i", "edge2D"}, {"i", "edge2D_in"}, {"i", "elem2D"},
    {"ai2", "elem2D_nodes"}, {"ai2", "edges"}, {"ai2", "edge_tri"}, {"ai2", "elem_edges"}, {"ar1", "elem_area"},
    {"ar2", "edge_dxdy"}, {"ar2", "edge_cross_dxdy"}, {"ar1", "elem_cos"}, {"ar1", "metric_factor"},
    {"ai2", "elem_neighbors"}, {"ai2", "nod_in_elem2D"}, {"ar2", "x_corners"}, {"ar2", "y_corners"},
    {"ai1", "nod_in_elem2D_num"}, {"ar1", "depth"}, {"ar2", "gradient_vec"}, {"ar2", "gradient_sca"},
    {"ai1", "bc_index_nod2D"}, {"i", "nl"}, {"ar1", "zbar"}, {"ar1", "Z"}, {"ar1", "elem_depth"}, {"ai1", "ulevels"},
    {"ai1", "ulevels_nod2D"}, {"ai1", "ulevels_nod2D_max"}, {"ai1", "nlevels"}, {"ai1", "nlevels_nod2D"},
    {"ai1", "nlevels_nod2D_min"}, {"ar2", "area"}, {"ar2", "area_inv"}, {"ar2", "areasvol"}, {"ar2", "areasvol_inv"},
    {"ar1", "mesh_resolution"}, {"ai1", "cavity_flag_n"}, {"ai1", "cavity_flag_e"}, {"ar1", "cavity_depth"},
    {"ar2", "cavity_nrst_cavlpnt_xyz"}, {"i", "ssh_stiff%dim"}, {"i", "ssh_stiff%nza"}, {"ai1", "ssh_stiff%rowptr"},
    {"ai1", "ssh_stiff%colind"}, {"ar1", "ssh_stiff%values"}, {"ai1", "ssh_stiff%colind_loc"},
    {"ai1", "ssh_stiff%rowptr_loc"}, {"ar1", "lump2d_south"}, {"ar1", "lump2d_north"}, {"ai1", "ind_south"},
    {"ai1", "ind_north"}, {"i", "nn_size"}, {"ai1", "nn_num"}, {"ai2", "nn_pos"}, {"ar2", "hnode"}, {"ar2", "hnode_new"},
    {"ar2", "zbar_3d_n"}, {"ar2", "Z_3d_n"}, {"ar2", "Z_3d_n_ib"}, {"ar2", "helem"}, {"ar1", "bottom_elem_thickness"},
    {"ar1", "bottom_node_thickness"}, {"ar1", "dhe"}, {"ar1", "hbar"}, {"ar1", "hbar_old"}, {"ar1", "zbar_n_bot"},
    {"ar1", "zbar_e_bot"}, {"ar1", "zbar_n_srf"}, {"ar1", "zbar_e_srf"}, {"ar1", "coriolis"}, {"ar1", "coriolis_node"}};
const Field T_COM_STRUCT[] = {{"i", "rPEnum"}, {"si", "rPE"}, {"si", "rptr"}, {"ai1", "rlist"}, {"i", "sPEnum"}, {"si", "sPE"},
                              {"si", "sptr"}, {"ai1", "slist"}, {"i", "nreq"}};
const Field T_PARTIT_TAIL[] = {{"i", "npes"}, {"i", "mype"}, {"i", "maxPEnum"}, {"ai1", "part"}, {"i", "myDim_nod2D"},
                               {"i", "eDim_nod2D"}, {"ai1", "myList_nod2D"}, {"i", "myDim_elem2D"}, {"i", "eDim_elem2D"},
                               {"i", "eXDim_elem2D"}, {"ai1", "myList_elem2D"}, {"i", "myDim_edge2D"}, {"i", "eDim_edge2D"},
                               {"ai1", "myList_edge2D"}, {"i", "pe_status"}};
const Field T_TRACER_DATA[] = {{"ar2", "values"}, {"ar3", "valuesold"}, {"ar2", "valuesAB"}, {"l", "smooth_bh_tra"},
                               {"r", "gamma0_tra"}, {"r", "gamma1_tra"}, {"r", "gamma2_tra"}, {"l", "i_vert_diff"},
                               {"c20", "tra_adv_hor"}, {"c20", "tra_adv_ver"}, {"c20", "tra_adv_lim"}, {"r", "tra_adv_ph"},
                               {"r", "tra_adv_pv"}, {"i", "ID"}};
const Field T_TRACER_WORK[] = {{"ar2", "del_ttf"}, {"ar2", "del_ttf_advhoriz"}, {"ar2", "del_ttf_advvert"}, {"ar3", "dvd_trflx_hor"},
                               {"ar3", "dvd_trflx_ver"}, {"ar2", "fct_LO"}, {"ar2", "adv_flux_hor"}, {"ar2", "adv_flux_ver"},
                               {"ar2", "fct_ttf_max"}, {"ar2", "fct_ttf_min"}, {"ar2", "fct_plus"}, {"ar2", "fct_minus"},
                               {"ai1", "nboundary_lay"}, {"ai2", "edge_up_dn_tri"}, {"ar3", "edge_up_dn_grad"}};
const Field T_SOLVERINFO[] = {{"i", "ident"}, {"i", "maxiter"}, {"i", "restart"}, {"i", "fillin"}, {"i", "lutype"}, {"r", "droptol"},
                              {"r", "soltol"}, {"ar1", "rr"}, {"ar1", "zz"}, {"ar1", "pp"}, {"ar1", "App"}};
const Field T_DYN_WORK[] = {{"ar3", "uvnode_rhs"}, {"ar2", "u_c"}, {"ar2", "v_c"}, {"ar2", "u_b"}, {"ar2", "v_b"}};
const Field T_DYN_HEAD[] = {{"i", "opt_visc"}, {"r", "visc_gamma0"}, {"r", "visc_gamma1"}, {"r", "visc_gamma2"},
                            {"r", "visc_easybsreturn"}, {"l", "use_ivertvisc"}, {"i", "momadv_opt"}, {"l", "use_freeslip"},
                            {"l", "use_wsplit"}, {"r", "wsplit_maxcfl"}, {"l", "use_ssh_se_subcycl"}};
const Field T_DYN_ARRAYS[] = {{"ar3", "uv"}, {"ar3", "uv_rhs"}, {"ar4", "uv_rhsAB"}, {"ar3", "uvnode"}, {"ar2", "w"}, {"ar2", "w_e"},
                              {"ar2", "w_i"}, {"ar2", "cfl_z"}};
const Field T_DYN_FER[] = {{"ar2", "fer_w"}, {"ar3", "fer_uv"}};
const Field T_DYN_SE[] = {{"ar3", "se_uvh"}, {"ar2", "se_uvBT_rhs"}, {"ar2", "se_uvBT_4AB"}, {"ar2", "se_uvBT"}, {"ar2", "se_uvBT_theta"},
                          {"ar2", "se_uvBT_mean"}, {"ar2", "se_uvBT_12"}, {"ar2", "se_uvBT_stab_hvisc"}, {"ar1", "se_uvBT_stab_bdrag"}};

struct Item {
    std::vector<int32_t> i;
    std::vector<double> r;
    std::vector<int> dims;          // Fortran order (first extent fastest)
    std::string s;
};
using Items = std::map<std::string, Item>;

// payload of a Fortran sequential unformatted file with the record markers stripped
struct Stream {
    std::string path;
    std::vector<unsigned char> buf;
    size_t pos = 0;
    explicit Stream(const std::string& p) : path(p)
    {
        FILE* f = std::fopen(p.c_str(), "rb");
        if (!f) throw std::runtime_error(p + ": cannot open");
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        std::vector<unsigned char> raw((size_t)n);
        if (n && std::fread(raw.data(), 1, (size_t)n, f) != (size_t)n) { std::fclose(f); throw std::runtime_error(p + ": short read"); }
        std::fclose(f);
        size_t q = 0;
        while (q < raw.size()) {
            int32_t ln, tail;
            if (q + 4 > raw.size()) throw std::runtime_error(p + ": truncated record marker at byte " + std::to_string(q));
            std::memcpy(&ln, &raw[q], 4);
            const size_t size = (size_t)(ln < 0 ? -(int64_t)ln : ln), end = q + 4 + size;
            if (end + 4 > raw.size()) throw std::runtime_error(p + ": record at byte " + std::to_string(q) + " runs past the end of the file");
            std::memcpy(&tail, &raw[end], 4);
            if ((size_t)(tail < 0 ? -(int64_t)tail : tail) != size)
                throw std::runtime_error(p + ": record markers disagree at byte " + std::to_string(q) + ": not a sequential unformatted file");
            buf.insert(buf.end(), raw.begin() + (long)(q + 4), raw.begin() + (long)end);
            q = end + 4;
        }
    }
    const unsigned char* take(size_t n)
    {
        if (pos + n > buf.size()) throw std::runtime_error(path + ": payload ends after " + std::to_string(buf.size()) + " bytes, item needs bytes " + std::to_string(pos) + ".." + std::to_string(pos + n));
        const unsigned char* p = buf.data() + pos;
        pos += n;
        return p;
    }
    int32_t i4() { int32_t v; std::memcpy(&v, take(4), 4); return v; }
    bool done() const { return pos == buf.size(); }
};

int static_len(const std::string& name) { return name == "rptr" ? 33 : 32; }   // MAX_NEIGHBOR_PARTITIONS = 32, src/MOD_PARTIT.F90:15,:20-25

template <size_t NF>
void read_items(Stream& s, const Field (&schema)[NF], Items& out, const std::string& prefix = "")
{
    for (const Field& f : schema) {
        Item& it = out[prefix + f.name];
        const std::string kind = f.kind;
        if (kind == "i" || kind == "l") it.i.assign(1, s.i4());
        else if (kind == "r") { it.r.resize(1); std::memcpy(it.r.data(), s.take(8), 8); }
        else if (kind[0] == 'c') {
            const size_t n = (size_t)std::stoi(kind.substr(1));
            it.s.assign(reinterpret_cast<const char*>(s.take(n)), n);
            it.s.erase(it.s.find_last_not_of(' ') + 1);
        } else if (kind == "si") {
            const int n = s.i4();
            if (n != static_len(f.name)) throw std::runtime_error(s.path + ": " + prefix + f.name + " has " + std::to_string(n) + " entries, the reference declares " + std::to_string(static_len(f.name)));
            it.i.resize((size_t)n);
            std::memcpy(it.i.data(), s.take(4 * (size_t)n), 4 * (size_t)n);
        } else {                                        // write_bin_array
            const int rank = kind[2] - '0';
            size_t cnt = 1;
            it.dims.resize((size_t)rank);
            for (int k = 0; k < rank; ++k) {
                it.dims[(size_t)k] = s.i4();
                if (it.dims[(size_t)k] < 0) throw std::runtime_error(s.path + ": negative extent for " + prefix + f.name);
                cnt *= (size_t)it.dims[(size_t)k];
            }
            if (cnt == 0) continue;
            if (kind[1] == 'i') { it.i.resize(cnt); std::memcpy(it.i.data(), s.take(4 * cnt), 4 * cnt); }
            else { it.r.resize(cnt); std::memcpy(it.r.data(), s.take(8 * cnt), 8 * cnt); }
        }
    }
}

void left_over(const Stream& s, const char* what)
{
    if (!s.done()) throw std::runtime_error(s.path + ": " + std::to_string(s.buf.size() - s.pos) + " bytes left after " + what + " (different FESOM version, or Fer_GM?)");
}

// the first `ncol` columns of a (ld, *) array
template <class T>
std::vector<T> head_cols(const std::vector<T>& a, size_t ld, size_t ncol)
{
    if (a.size() < ld * ncol) throw std::runtime_error("restart array shorter than the partition's dimensions");
    return std::vector<T>(a.begin(), a.begin() + (long)(ld * ncol));
}

void fill_com(const Items& d, const std::string& n, com_struct& c)
{
    c.rPEnum = d.at(n + "%rPEnum").i[0]; c.sPEnum = d.at(n + "%sPEnum").i[0];
    const auto& rPE = d.at(n + "%rPE").i; const auto& rptr = d.at(n + "%rptr").i;
    const auto& sPE = d.at(n + "%sPE").i; const auto& sptr = d.at(n + "%sptr").i;
    c.rPE.assign(rPE.begin(), rPE.begin() + c.rPEnum); c.rptr.assign(rptr.begin(), rptr.begin() + c.rPEnum + 1);
    c.sPE.assign(sPE.begin(), sPE.begin() + c.sPEnum); c.sptr.assign(sptr.begin(), sptr.begin() + c.sPEnum + 1);
    c.rlist = d.at(n + "%rlist").i; c.slist = d.at(n + "%slist").i;
}

}  // namespace

std::string mpirank_to_txt(int mype, int npes)
{
    const int width = (int)std::log10((double)npes) + 1;
    char b[32];
    std::snprintf(b, sizeof b, "%0*d", width, mype);
    return b;
}

void read_all_bin_restarts(const std::string& path_in, int mype, int npes, t_partit& partit, t_mesh& mesh, t_dyn& dynamics,
                           t_tracer& tracers, bool fer_gm)
{
    const std::string sfx = mpirank_to_txt(mype, npes);
    Items tp, tm, td;
    {                                                   // t_partit: the three communicators, then the scalars and lists
        Stream s(path_in + "/t_partit." + sfx);
        for (const char* com : {"com_nod2D", "com_elem2D", "com_elem2D_full"}) read_items(s, T_COM_STRUCT, tp, std::string(com) + "%");
        read_items(s, T_PARTIT_TAIL, tp);
        left_over(s, "t_partit");
    }
    partit.npes = tp.at("npes").i[0]; partit.mype = tp.at("mype").i[0];
    partit.myDim_nod2D = tp.at("myDim_nod2D").i[0]; partit.eDim_nod2D = tp.at("eDim_nod2D").i[0];
    partit.myDim_elem2D = tp.at("myDim_elem2D").i[0]; partit.eDim_elem2D = tp.at("eDim_elem2D").i[0];
    partit.myDim_edge2D = tp.at("myDim_edge2D").i[0];
    fill_com(tp, "com_nod2D", partit.com_nod2D);
    const size_t N = (size_t)partit.myDim_nod2D, Nh = N + (size_t)partit.eDim_nod2D;
    const size_t T = (size_t)partit.myDim_elem2D + (size_t)partit.eDim_elem2D, E = (size_t)partit.myDim_edge2D;
    {
        Stream s(path_in + "/t_mesh." + sfx);
        read_items(s, T_MESH, tm);
        left_over(s, "t_mesh");
    }
    mesh.nl = tm.at("nl").i[0];
    const size_t nl = (size_t)mesh.nl, L = nl - 1;
    mesh.edges = head_cols(tm.at("edges").i, 2, E); mesh.edge_tri = head_cols(tm.at("edge_tri").i, 2, E);
    mesh.elem2D_nodes = head_cols(tm.at("elem2D_nodes").i, 3, T);
    mesh.nod_in_elem2D_ld = tm.at("nod_in_elem2D").dims.at(0);
    mesh.nod_in_elem2D = head_cols(tm.at("nod_in_elem2D").i, (size_t)mesh.nod_in_elem2D_ld, Nh);
    mesh.nod_in_elem2D_num = head_cols(tm.at("nod_in_elem2D_num").i, 1, Nh);
    mesh.nlevels = head_cols(tm.at("nlevels").i, 1, T); mesh.ulevels = head_cols(tm.at("ulevels").i, 1, T);
    mesh.nlevels_nod2D = head_cols(tm.at("nlevels_nod2D").i, 1, Nh); mesh.ulevels_nod2D = head_cols(tm.at("ulevels_nod2D").i, 1, Nh);
    mesh.edge_cross_dxdy = head_cols(tm.at("edge_cross_dxdy").r, 4, E); mesh.edge_dxdy = head_cols(tm.at("edge_dxdy").r, 2, E);
    mesh.elem_cos = head_cols(tm.at("elem_cos").r, 1, T);
    mesh.area = head_cols(tm.at("area").r, nl, Nh); mesh.areasvol = head_cols(tm.at("areasvol").r, nl, Nh);
    mesh.helem = head_cols(tm.at("helem").r, L, T);
    mesh.hnode = head_cols(tm.at("hnode").r, L, Nh); mesh.hnode_new = head_cols(tm.at("hnode_new").r, L, Nh);
    mesh.zbar_3d_n = head_cols(tm.at("zbar_3d_n").r, nl, Nh); mesh.Z_3d_n = head_cols(tm.at("Z_3d_n").r, L, Nh);
    mesh.zbar_n_bot = tm.at("zbar_n_bot").r;
    {
        Stream s(path_in + "/t_tracer." + sfx);
        tracers.num_tracers = s.i4();                   // src/MOD_TRACER.F90:209
        tracers.data.assign((size_t)tracers.num_tracers, t_tracer_data());
        for (t_tracer_data& x : tracers.data) {
            Items d;
            read_items(s, T_TRACER_DATA, d);
            x.values = head_cols(d.at("values").r, L, Nh); x.valuesAB = head_cols(d.at("valuesAB").r, L, Nh);
            x.tra_adv_hor = d.at("tra_adv_hor").s; x.tra_adv_ver = d.at("tra_adv_ver").s; x.tra_adv_lim = d.at("tra_adv_lim").s;
            x.tra_adv_ph = d.at("tra_adv_ph").r[0]; x.tra_adv_pv = d.at("tra_adv_pv").r[0];
            x.ID = d.at("ID").i[0];
        }
        Items w;
        read_items(s, T_TRACER_WORK, w);
        left_over(s, "t_tracer");
        t_tracer_work& wk = tracers.work;
        auto or_zero = [](const std::vector<double>& a, size_t n) { return a.empty() ? std::vector<double>(n, 0.0) : a; };
        wk.del_ttf = or_zero(w.at("del_ttf").r, L * Nh);
        wk.del_ttf_advhoriz = or_zero(w.at("del_ttf_advhoriz").r, L * Nh); wk.del_ttf_advvert = or_zero(w.at("del_ttf_advvert").r, L * Nh);
        wk.edge_up_dn_grad = w.at("edge_up_dn_grad").r;
        wk.nboundary_lay = w.at("nboundary_lay").i;
    }
    {
        Stream s(path_in + "/t_dynamics." + sfx);
        read_items(s, T_DYN_HEAD, td);
        read_items(s, T_SOLVERINFO, td, "solverinfo%");
        read_items(s, T_DYN_WORK, td, "work%");
        read_items(s, T_DYN_ARRAYS, td);
        if (fer_gm) read_items(s, T_DYN_FER, td);
        if (td.at("use_ssh_se_subcycl").i[0] != 0) read_items(s, T_DYN_SE, td);
        left_over(s, "t_dynamics");
    }
    dynamics.use_wsplit = td.at("use_wsplit").i[0] != 0;
    dynamics.uv = head_cols(td.at("uv").r, 2 * L, T);
    dynamics.w = head_cols(td.at("w").r, nl, Nh); dynamics.w_e = head_cols(td.at("w_e").r, nl, Nh); dynamics.w_i = head_cols(td.at("w_i").r, nl, Nh);
}

}  // namespace fesom
