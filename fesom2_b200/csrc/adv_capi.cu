// fesom2_b200/csrc/adv_capi.cu -- C ABI (include/fesom_adv_b200.h): context, gather-list
// construction, launch sequencing with halo/compute overlap, packed-halo NCCL exchange.
//
// There is deliberately no CPU path in this file: every entry point needs a CUDA device.
#include "../../include/fesom_adv_b200.h"
#include "adv_kernels.cuh"
#include "adv_pipe.cuh"

#include <nccl.h>   // types only: the library is resolved at run time (see NcclApi)
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

using namespace adv;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ADV_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)
#define NC(call)                                                                              \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess)                                                                \
            return fail(ADV_ENCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_));    \
    } while (0)

static const double R_EARTH = 6367500.0;  // src/oce_modules.F90:29

// NCCL is bound with dlopen at the first communicator call instead of at link time: inside a
// Python process torch has already loaded its own libnccl.so.2 (same SONAME) and that copy must be
// the one in use; a Fortran host gets the system library.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;
static int nccl_load()
{
    if (g_nccl.ok) return ADV_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(ADV_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return fail(ADV_ENCCL, "missing NCCL symbol " name);
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.ok = true;
    return ADV_OK;
}

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;              // owns device memory: never copied
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    cudaError_t alloc(size_t count, bool zero = true)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        // cudaMemset on device memory is asynchronous on the LEGACY default stream, which the context's non-blocking
        // streams do not synchronise with: without the wait a kernel could write the buffer before the zeroing lands
        // (seen under compute-sanitizer's timing: work arrays allocated inside the first call came back zeroed)
        if (zero) { e = cudaMemset(p, 0, count * sizeof(T)); if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy); }
        return e;
    }
    cudaError_t upload(const std::vector<T>& h)
    {
        cudaError_t e = alloc(h.size(), false);
        if (e != cudaSuccess || h.empty()) return e;
        // a pageable-source cudaMemcpy may return while the DMA from its staging buffer is still in flight on the
        // legacy stream: wait, for the same reason as above
        e = cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

struct Slot {  // per-tracer staging: host pointers, unaligned device pointers, library-computed gradients
    DevBuf<double> ttf, ttfAB, grad, dh, dv, tr_xy, gmean, dgh, dgv, dvdh, dvdv;
};

// work arrays of one chunk of <= 2 tracers (t_tracer_work, allocated by oce_adv_tra_fct_init in the
// reference), tracer-interleaved, always sized for 2 tracers
struct ChunkBuf {
    DevBuf<double> lo, pm, adf_h, adf_v;
    DevBuf<double> sendbuf;                   // send columns x L x 4
    int tb_last = 0;                          // tracers per chunk of the last use: the [column][level][TB] interleave depends on it
};

struct TrLoc { int chunk = -1, pos = 0, tb = 1; };  // where tracer i of the last call lives

struct Peer { int pe; int off, cnt; };  // segment of slist / halo tail, in columns

// one com_struct (src/MOD_PARTIT.F90:18-33) on the device: com_nod2D or com_elem2D_full
struct HaloSet {
    DevBuf<int> slist, rlist;                 // 0-based local ids; rlist only when the receive side needs an unpack
    std::vector<Peer> rpeers, speers;
    int send_cols = 0, recv_cols = 0;
    int recv_base = -1;                       // >= 0: the received columns are the contiguous tail starting here (no unpack)
};
enum { HALO_NOD = 0, HALO_ELEM = 1 };

// In-process communicator (adv_ctx_comm_init_local): the contexts of one process exchange halos with direct
// device-to-device copies -- peer copies over NVLink between GPUs, plain copies on one GPU -- ordered by CUDA
// events; two host rendezvous per exchange make the events visible to the peers.
struct LocalComm {
    int n = 0;
    std::vector<adv_ctx*> ctx;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    unsigned gen = 0;
    bool broken = false;
    bool barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        if (broken) return false;
        const unsigned g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return true; }
        if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g || broken; })) { broken = true; cv.notify_all(); return false; }
        return !broken;
    }
};

}  // namespace

struct adv_ctx {
    int device = 0;
    int max_tr = 0;
    MeshDev m{};
    int mype = 0, npes = 1;
    // topology
    DevBuf<int> ne_ptr, nboundary_lay, list_S, list_I, list_SH;
    HaloSet halo[2];                          // HALO_NOD: com_nod2D, HALO_ELEM: com_elem2D_full (adv_ctx_set_gradient_mesh)
    DevBuf<int4> ne_ent;
    DevBuf<int2> edge_el;
    DevBuf<int4> edge_meta;
    DevBuf<uchar4> node_lev, edge_lev;
    DevBuf<uint2> node_rec;
    DevBuf<int4> ne_ell;
    DevBuf<double4> edge_cross;
    DevBuf<double2> edge_c;
    DevBuf<double> area, areasvol, r_areasvol, Q;
    // gradient producer (adv_ctx_set_gradient_mesh)
    DevBuf<int> g_nie, g_nie_num, g_nlevels, g_ulevels, g_tri, g_nmin, g_umax, g_elem_nodes;
    DevBuf<double> g_sca, g_earea;
    DevBuf<int4> edge_g;
    int fuse_grad = 1;                        // edge_up_dn_grad = NULL: reconstruct the gradients inside the edge kernel (ADV_FUSE_GRAD)
    GradMeshDev gm{};
    bool grad_mesh_set = false, elem_halo_set = false;
    int nS = 0, nI = 0, nSH = 0;
    int pf_dist = 200;                        // L2 prefetch distance of the node kernels in CTAs (ADV_PF; 0 = off)
    int g_lo = 3, g_k2 = 1, g_k3 = 2;         // gather batch sizes (tunable: ADV_G_LO / ADV_G_K2 / ADV_G_K3)
    int bulk = 1;                             // bulk-copy edge kernel (adv_pipe.cuh); 0 = register-gather k_edge_flux (ADV_BULK)
    int e1_ng = 8, e1_depth = 2, e1_il = 0;   // edge groups per CTA / stages / grid-strided groups (ADV_E1_NG, ADV_E1_D, ADV_E1_IL)
    int i_identity = 1;                       // interior range as identity range + skip flags (ADV_I_IDENTITY)
    int e1_pf = 100;                          // metadata prefetch distance of the bulk edge kernel in CTAs (ADV_E1_PF)
    int force_tb1 = 0;                        // experiments: one tracer per chunk (ADV_TB1)
    // wet-level compaction: CTA partitions of the node ranges and edge groups (built in adv_ctx_create)
    int cta_threads = 224;                    // threads per CTA of the FCT node kernels (ADV_CTA_THREADS) ...
    int cta_n1 = 288, cta_k2 = 256, cta_k3 = 288; // ... per kernel (ADV_CTA_N1 / _K2 / _K3; 0 = cta_threads): the size at which the register
                                              //     file holds most warps -- N1 (56 registers) 4 x 9 = 36, K2 (64) 4 x 8 = 32, K3 (56) 4 x 9 = 36
    struct Part { DevBuf<int> first; int ncta = 0; };
    struct PartSet { int threads = 0; Part all, inner, s, sh; };   // the CTA partitions of the four node ranges for one CTA size
    PartSet parts[3];                         // one per distinct CTA size in use (at most three kernels)
    int nparts = 0;
    const PartSet& parts_for(int threads) const
    {
        for (int i = 0; i < nparts; ++i) if (parts[i].threads == threads) return parts[i];
        return parts[0];
    }
    int max_smem_optin = 0;
    // state (ADV_HOST staging)
    DevBuf<double> uv, helem, w, we, wi, hnode, hnode_new, zbar3d, Z3d, zbar_n_bot;
    bool state_set = false, q_valid = false;
    int64_t state_step = -1;                  // adv_ctx_set_state_step: step number of the resident state (-1: none)
    int state_where = -1;
    adv_state_desc_t state_desc{};
    bool host_register = true;                // page-lock ADV_HOST arrays on first use (adv_ctx_set_host_register)
    std::map<const void*, size_t> registered; // ranges this context registered (0 bytes: left alone)
    std::set<const void*> registered_foreign;
    std::vector<Slot> slots;
    std::vector<std::unique_ptr<ChunkBuf>> cbufs;
    std::vector<TrLoc> trloc;
    DevBuf<double> xbuf, xrbuf;  // adv_exchange_nod / adv_exchange_elem pack and unpack buffers
    // in-process communicator
    std::shared_ptr<LocalComm> lc;
    cudaEvent_t ev_packed = nullptr, ev_consumed = nullptr, ev_order = nullptr;
    std::vector<const double*> pub_send;      // send buffers of the exchange in progress, read by the peers
    // halo-exchange accounting (adv_ctx_halo_stats)
    int64_t halo_bytes_sent = 0, halo_exchanges = 0, halo_bytes_last = 0;
    cudaEvent_t ev_x0[2] = {nullptr, nullptr}, ev_x1[2] = {nullptr, nullptr}, ev_w0[2] = {nullptr, nullptr}, ev_w1[2] = {nullptr, nullptr};
    bool x_valid = false;
    cudaStream_t s_comp = nullptr, s_comm = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_ph[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool profiling = false, ph_valid = false;
    ncclComm_t comm = nullptr;
    int64_t launches = 0;
    bool timed = false;
};

const char* adv_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------------
static int parse_scheme(const char* s, const char* const* names, const int* codes, int n)
{
    if (!s) return -1;
    char buf[16];
    int k = 0;
    while (k < 15 && s[k] && s[k] != ' ') { buf[k] = s[k]; ++k; }   // Fortran strings are blank padded
    buf[k] = 0;
    for (int i = 0; i < n; ++i)
        if (strcmp(buf, names[i]) == 0) return codes[i];
    return -1;
}

enum RangeId { R_ALL = 0, R_S, R_I, R_SH, R_ALLH };

int adv_ctx_create(adv_ctx_t** out, const adv_mesh_desc_t* d, int device, int max_tracers)
{
    if (!out || !d || max_tracers < 1) return fail(ADV_EINVAL, "adv_ctx_create: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ADV_ECUDA, "no CUDA device: this library has no CPU fallback");
    CU(cudaSetDevice(device));
    const int nl = d->nl, L = nl - 1;
    const int N = d->myDim_nod2D, Nh = N + d->eDim_nod2D, T = d->myDim_elem2D, E = d->myDim_edge2D;
    if (L < 1 || L > kBlock || nl > 255) return fail(ADV_EINVAL, "nl out of the supported range (2..255)");
    if (N < 1 || T < 1 || E < 1) return fail(ADV_EINVAL, "empty mesh");

    // ---- host-side gather lists ----------------------------------------------------------------
    std::vector<int2> edge_el(E);
    std::vector<int4> edge_meta(E);
    std::vector<uchar4> edge_lev(E);
    std::vector<double4> edge_cross(E);
    std::vector<double2> edge_c(E);
    std::vector<int> deg(Nh + 1, 0);
    for (int e = 0; e < E; ++e) {
        const int n1 = d->edges[2 * e] - 1, n2 = d->edges[2 * e + 1] - 1;
        const int el1 = d->edge_tri[2 * e] - 1, el2 = d->edge_tri[2 * e + 1] > 0 ? d->edge_tri[2 * e + 1] - 1 : -1;
        if (n1 < 0 || n1 >= Nh || n2 < 0 || n2 >= Nh || el1 < 0 || el1 >= T || el2 >= T)
            return fail(ADV_EINVAL, "edge " + std::to_string(e + 1) + ": index out of range");
        const int nu1 = d->ulevels[el1], nl1 = d->nlevels[el1] - 1;
        int nu2 = 0, nl2 = 0;
        double a = R_EARTH * d->elem_cos[el1];                       // oce_adv_tra_hor.F90:325,338
        if (el2 >= 0) {
            nu2 = d->ulevels[el2]; nl2 = d->nlevels[el2] - 1;
            a = 0.5 * (a + R_EARTH * d->elem_cos[el2]);
            if (std::max(nu1, nu2) > std::min(nl1, nl2) + 1)
                return fail(ADV_EINVAL, "edge " + std::to_string(e + 1) + ": adjacent elements have disjoint level ranges");
        }
        if (nu1 < 1 || nl1 > L || nl2 > L || nu1 > nl1) return fail(ADV_EINVAL, "element levels out of range");
        edge_el[e] = make_int2(el1, el2);
        edge_meta[e] = make_int4(n1, n2, el1, el2);
        edge_lev[e] = make_uchar4((unsigned char)nu1, (unsigned char)nl1, (unsigned char)nu2, (unsigned char)nl2);
        edge_cross[e] = make_double4(d->edge_cross_dxdy[4 * e], d->edge_cross_dxdy[4 * e + 1],
                                     d->edge_cross_dxdy[4 * e + 2], d->edge_cross_dxdy[4 * e + 3]);
        edge_c[e] = make_double2(d->edge_dxdy[2 * e] * a, d->edge_dxdy[2 * e + 1] * R_EARTH);
        ++deg[n1 + 1]; ++deg[n2 + 1];
    }
    std::vector<int> ne_ptr(Nh + 1, 0);
    for (int n = 0; n < Nh; ++n) ne_ptr[n + 1] = ne_ptr[n] + deg[n + 1];
    std::vector<int4> ne_ent(ne_ptr[Nh]);
    {
        std::vector<int> fill(ne_ptr.begin(), ne_ptr.end() - 1);
        for (int e = 0; e < E; ++e) {   // ascending e => every node's list is ascending (serial scatter order)
            const int n1 = d->edges[2 * e] - 1, n2 = d->edges[2 * e + 1] - 1;
            const uchar4 lv = edge_lev[e];
            // scatter range of oce_adv_tra_driver.F90:154-156: [min(nu1, nu2>0), max(nl1, nl2)]
            const int lo = lv.z > 0 ? std::min<int>(lv.x, lv.z) : lv.x;
            const int hi = std::max<int>(lv.y, lv.w);
            ne_ent[fill[n1]++] = make_int4(e, n2, lo | (hi << 8) | (0 << 16), 0);
            ne_ent[fill[n2]++] = make_int4(e, n1, lo | (hi << 8) | (1 << 16), 0);
        }
    }
    // FCT clusters (oce_adv_tra_fct.F90:148-215 collapsed): for owned node n the distinct nodes of
    // its elements, each with the union of the level ranges of the elements that contain it.  On a
    // triangulation that is the node itself plus its edge neighbours with the edge's scatter range,
    // so the kernels reuse the edge slots; the equivalence is verified here for every owned node.
    std::vector<uchar4> node_lev(Nh);
    std::vector<uint2> node_rec(Nh);
    int ell_w = 1;
    for (int n = 0; n < Nh; ++n) ell_w = std::max(ell_w, ne_ptr[n + 1] - ne_ptr[n]);
    if (ell_w > 255) return fail(ADV_EINVAL, "node degree > 255");
    std::vector<int4> ne_ell((size_t)Nh * ell_w, make_int4(0, 0, 0xff, 0));   // padding: lo = 255 > hi = 0
    for (int n = 0; n < Nh; ++n)
        for (int k = ne_ptr[n]; k < ne_ptr[n + 1]; ++k) ne_ell[(size_t)n * ell_w + (k - ne_ptr[n])] = ne_ent[k];
    const int ld = d->nod_in_elem2D_ld;
    struct Iv { int node, lo, hi; };
    std::vector<Iv> iv, cl;
    for (int n = 0; n < Nh; ++n) {
        int pad_lo = 0, pad_hi = 255, self_lo = 1, self_hi = 0;
        if (n < N) {
            iv.clear(); cl.clear();
            const int num = d->nod_in_elem2D_num[n];
            if (num < 1 || num > ld) return fail(ADV_EINVAL, "nod_in_elem2D_num out of range");
            for (int k = 0; k < num; ++k) {
                const int el = d->nod_in_elem2D[(size_t)n * ld + k] - 1;
                if (el < 0 || el >= T) return fail(ADV_EINVAL, "nod_in_elem2D out of range");
                const int lo = d->ulevels[el], hi = d->nlevels[el] - 1;
                pad_lo = std::max(pad_lo, lo); pad_hi = std::min(pad_hi, hi);
                for (int j = 0; j < 3; ++j) iv.push_back({d->elem2D_nodes[3 * el + j] - 1, lo, hi});
            }
            std::sort(iv.begin(), iv.end(), [](const Iv& a, const Iv& b) { return a.node != b.node ? a.node < b.node : a.lo < b.lo; });
            for (size_t i = 0; i < iv.size();) {
                Iv cur = iv[i];
                size_t j = i + 1;
                for (; j < iv.size() && iv[j].node == cur.node && iv[j].lo <= cur.hi + 1; ++j) cur.hi = std::max(cur.hi, iv[j].hi);
                cl.push_back(cur);
                i = j;
            }
            // compare with {self} + edge slots
            std::vector<Iv> ref;
            for (int k = ne_ptr[n]; k < ne_ptr[n + 1]; ++k) ref.push_back({ne_ent[k].y, ne_ent[k].z & 0xff, (ne_ent[k].z >> 8) & 0xff});
            std::sort(ref.begin(), ref.end(), [](const Iv& a, const Iv& b) { return a.node < b.node; });
            bool ok = cl.size() == ref.size() + 1;
            self_lo = 0; self_hi = -1;
            for (size_t i = 0, j = 0; ok && i < cl.size(); ++i) {
                if (cl[i].node == n) { self_lo = cl[i].lo; self_hi = cl[i].hi; continue; }
                ok = j < ref.size() && ref[j].node == cl[i].node && ref[j].lo == cl[i].lo && ref[j].hi == cl[i].hi;
                ++j;
            }
            if (!ok || self_hi < self_lo)
                return fail(ADV_EINVAL, "node " + std::to_string(n + 1) + ": FCT cluster is not {node} + edge neighbours "
                                        "(non-triangular mesh or non-contiguous element level ranges)");
        }
        node_rec[n] = make_uint2((unsigned)d->ulevels_nod2D[n] | ((unsigned)d->nlevels_nod2D[n] << 8) | ((unsigned)pad_lo << 16) | ((unsigned)pad_hi << 24),
                                 (unsigned)self_lo | ((unsigned)self_hi << 8) | ((unsigned)(ne_ptr[n + 1] - ne_ptr[n]) << 16));
        node_lev[n] = make_uchar4((unsigned char)d->ulevels_nod2D[n], (unsigned char)d->nlevels_nod2D[n],
                                  (unsigned char)pad_lo, (unsigned char)pad_hi);
        if (d->nlevels_nod2D[n] > nl || d->ulevels_nod2D[n] < 1) return fail(ADV_EINVAL, "node levels out of range");
    }

    adv_ctx* c = new adv_ctx();
    c->device = device; c->max_tr = max_tracers; c->mype = d->mype; c->npes = std::max(1, d->npes);
    if (const char* v = getenv("ADV_PF")) c->pf_dist = std::max(0, atoi(v));
    if (const char* v = getenv("ADV_G_LO")) c->g_lo = atoi(v);
    if (const char* v = getenv("ADV_G_K2")) c->g_k2 = atoi(v);
    if (const char* v = getenv("ADV_G_K3")) c->g_k3 = atoi(v);
    if (const char* v = getenv("ADV_BULK")) c->bulk = atoi(v);
    if (const char* v = getenv("ADV_E1_NG")) c->e1_ng = std::max(1, atoi(v));
    if (const char* v = getenv("ADV_E1_IL")) c->e1_il = atoi(v) ? 1 : 0;
    if (const char* v = getenv("ADV_E1_D")) c->e1_depth = std::max(2, std::min(4, atoi(v)));
    if (const char* v = getenv("ADV_TB1")) c->force_tb1 = atoi(v);
    if (const char* v = getenv("ADV_I_IDENTITY")) c->i_identity = atoi(v) ? 1 : 0;
    if (const char* v = getenv("ADV_E1_PF")) c->e1_pf = std::max(0, atoi(v));
    if (const char* v = getenv("ADV_CTA_THREADS")) { c->cta_threads = atoi(v); c->cta_n1 = 0; c->cta_k2 = 0; c->cta_k3 = 0; }
    if (const char* v = getenv("ADV_CTA_N1")) c->cta_n1 = atoi(v);
    if (const char* v = getenv("ADV_CTA_K2")) c->cta_k2 = atoi(v);
    if (const char* v = getenv("ADV_CTA_K3")) c->cta_k3 = atoi(v);
    if (const char* v = getenv("ADV_FUSE_GRAD")) c->fuse_grad = atoi(v) ? 1 : 0;
    auto clamp_cta = [&](int t) { return std::max(((L + 31) / 32) * 32, std::min(1024, (t / 32) * 32)); };   // whole warps, at least one column
    c->cta_threads = clamp_cta(c->cta_threads);
    c->cta_n1 = c->cta_n1 > 0 ? clamp_cta(c->cta_n1) : c->cta_threads;
    c->cta_k2 = c->cta_k2 > 0 ? clamp_cta(c->cta_k2) : c->cta_threads;
    c->cta_k3 = c->cta_k3 > 0 ? clamp_cta(c->cta_k3) : c->cta_threads;
    cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { delete c; return fail(ADV_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    CUF(c->ne_ptr.upload(ne_ptr)); CUF(c->ne_ent.upload(ne_ent));
    CUF(c->node_lev.upload(node_lev)); CUF(c->node_rec.upload(node_rec)); CUF(c->ne_ell.upload(ne_ell));
    CUF(c->edge_el.upload(edge_el)); CUF(c->edge_meta.upload(edge_meta)); CUF(c->edge_lev.upload(edge_lev));
    CUF(c->edge_cross.upload(edge_cross)); CUF(c->edge_c.upload(edge_c));
    {
        std::vector<int> nb(Nh, L);
        if (d->nboundary_lay) nb.assign(d->nboundary_lay, d->nboundary_lay + Nh);
        CUF(c->nboundary_lay.upload(nb));
        std::vector<double> tmp(d->area, d->area + (size_t)nl * Nh);
        CUF(c->area.upload(tmp));
        tmp.assign(d->areasvol, d->areasvol + (size_t)nl * Nh);
        CUF(c->areasvol.upload(tmp));
        for (double& x : tmp) x = 1.0 / x;                 // IEEE division on the host == the device's 1.0 / av
        CUF(c->r_areasvol.upload(tmp));
    }
    CUF(c->Q.alloc((size_t)L * E));
    {
        std::vector<int> en(d->elem2D_nodes, d->elem2D_nodes + (size_t)3 * T);
        CUF(c->g_elem_nodes.upload(en));
    }

    // ---- halo bookkeeping (com_nod2D) ------------------------------------------------------------
    if (c->npes > 1) {
        HaloSet& hn = c->halo[HALO_NOD];
        std::vector<char> isS(N, 0);
        std::vector<int> sl;
        for (int i = 0; i < d->sPEnum; ++i) {
            Peer p{d->sPE[i], d->sptr[i] - 1, d->sptr[i + 1] - d->sptr[i]};
            hn.speers.push_back(p);
            for (int k = 0; k < p.cnt; ++k) {
                const int n = d->slist[p.off + k] - 1;
                if (n < 0 || n >= N) { delete c; return fail(ADV_EINVAL, "slist entry is not an owned node"); }
                isS[n] = 1; sl.push_back(n);
            }
        }
        hn.send_cols = (int)sl.size();
        for (int i = 0; i < d->rPEnum; ++i) {
            Peer p{d->rPE[i], d->rptr[i] - 1, d->rptr[i + 1] - d->rptr[i]};
            // the halo tail must be contiguous per source rank: rlist(k) = myDim_nod2D + k (oce_local.F90:41)
            for (int k = 0; k < p.cnt; ++k)
                if (d->rlist[p.off + k] != N + p.off + k + 1) { delete c; return fail(ADV_EINVAL, "rlist is not the identity on the halo tail"); }
            hn.rpeers.push_back(p);
            hn.recv_cols += p.cnt;
        }
        hn.recv_base = N;
        // a node also belongs to the boundary set when one of its edge neighbours is a halo node
        for (int e = 0; e < E; ++e) {
            const int n1 = d->edges[2 * e] - 1, n2 = d->edges[2 * e + 1] - 1;
            if (n1 < N && n2 >= N) isS[n1] = 1;
            if (n2 < N && n1 >= N) isS[n2] = 1;
        }
        std::vector<int> S, I, SH;
        for (int n = 0; n < N; ++n) (isS[n] ? S : I).push_back(n);
        for (int n = 0; n < N; ++n) if (isS[n]) node_rec[n].y |= 1u << 24;       // boundary-set flag (NodeRange::skip_s)
        CUF(c->node_rec.upload(node_rec));
        SH = S;
        for (int n = N; n < Nh; ++n) SH.push_back(n);
        c->nS = (int)S.size(); c->nI = (int)I.size(); c->nSH = (int)SH.size();
        CUF(c->list_S.upload(S)); CUF(c->list_I.upload(I)); CUF(c->list_SH.upload(SH)); CUF(hn.slist.upload(sl));
    }
    // ---- wet-level compaction: pack whole columns into CTAs (one thread per wet layer), once per CTA size in use -----
    for (int threads : {c->cta_n1, c->cta_k2, c->cta_k3}) {
        bool have = false;
        for (int i = 0; i < c->nparts; ++i) have = have || c->parts[i].threads == threads;
        if (have) continue;
        adv_ctx::PartSet& ps = c->parts[c->nparts++];
        ps.threads = threads;
        auto pack = [&](int count, auto wet_of, adv_ctx::Part& out) -> cudaError_t {
            std::vector<int> first(1, 0);
            int used = 0, ncol = 0;
            for (int i = 0; i < count; ++i) {
                const int w = wet_of(i);
                if (ncol > 0 && (used + w > threads || ncol == kMaxCols)) { first.push_back(i); used = 0; ncol = 0; }
                used += w; ++ncol;
            }
            first.push_back(count);
            out.ncta = count > 0 ? (int)first.size() - 1 : 0;
            return out.first.upload(first);
        };
        auto wet_node = [&](int n) { return d->nlevels_nod2D[n] - d->ulevels_nod2D[n]; };
        CUF(pack(N, wet_node, ps.all));
        if (c->npes > 1) {
            std::vector<int> S, SH;
            for (int n = 0; n < N; ++n) if ((node_rec[n].y >> 24) & 1u) S.push_back(n);
            SH = S;
            for (int n = N; n < Nh; ++n) SH.push_back(n);
            CUF(pack(N, [&](int n) { return ((node_rec[n].y >> 24) & 1u) ? 0 : wet_node(n); }, ps.inner));
            CUF(pack((int)S.size(), [&](int i) { return wet_node(S[i]); }, ps.s));
            CUF(pack((int)SH.size(), [&](int i) { return wet_node(SH[i]); }, ps.sh));
        }
    }
    c->slots.resize(max_tracers);
    c->trloc.resize(max_tracers);
    CUF(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    CUF(cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking));
    for (cudaEvent_t* ev : {&c->ev_a, &c->ev_b, &c->ev_c, &c->ev_d}) CUF(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    CUF(cudaEventCreate(&c->ev_t0)); CUF(cudaEventCreate(&c->ev_t1));
    for (auto& ev : c->ev_ph) CUF(cudaEventCreate(&ev));
    CUF(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming)); CUF(cudaEventCreateWithFlags(&c->ev_consumed, cudaEventDisableTiming)); CUF(cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) { CUF(cudaEventCreate(&c->ev_x0[i])); CUF(cudaEventCreate(&c->ev_x1[i])); CUF(cudaEventCreate(&c->ev_w0[i])); CUF(cudaEventCreate(&c->ev_w1[i])); }
#undef CUF
    MeshDev& m = c->m;
    m.L = L; m.nl = nl; m.N = N; m.Nh = Nh; m.T = T; m.E = E;
    m.div_magic = ((1u << 20) + L - 1) / L;
    for (unsigned t = 0; t < 1024; ++t)
        if (((t * m.div_magic) >> 20) != t / L) { delete c; return fail(ADV_EINVAL, "internal: division magic"); }
    m.ne_ptr = c->ne_ptr.p; m.ne_ent = c->ne_ent.p;
    m.node_lev = c->node_lev.p; m.node_rec = c->node_rec.p; m.ne_ell = c->ne_ell.p; m.ell_w = ell_w;
    m.edge_meta = c->edge_meta.p; m.edge_el = c->edge_el.p; m.edge_lev = c->edge_lev.p;
    m.edge_cross = c->edge_cross.p; m.edge_c = c->edge_c.p; m.nboundary_lay = c->nboundary_lay.p;
    m.area = c->area.p; m.areasvol = c->areasvol.p; m.r_areasvol = c->r_areasvol.p; m.Q = c->Q.p;
    *out = c;
    return ADV_OK;
}

int adv_ctx_destroy(adv_ctx_t* c)
{
    if (!c) return ADV_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    for (auto& kv : c->registered) if (kv.second) cudaHostUnregister(const_cast<void*>(kv.first));
    for (cudaEvent_t ev : {c->ev_a, c->ev_b, c->ev_c, c->ev_d, c->ev_t0, c->ev_t1}) if (ev) cudaEventDestroy(ev);
    for (auto ev : c->ev_ph) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : {c->ev_packed, c->ev_consumed, c->ev_order, c->ev_x0[0], c->ev_x0[1], c->ev_x1[0], c->ev_x1[1], c->ev_w0[0], c->ev_w0[1], c->ev_w1[0], c->ev_w1[1]})
        if (ev) cudaEventDestroy(ev);
    if (c->lc) {   // a communicator with a destroyed member can no longer rendezvous
        std::lock_guard<std::mutex> lk(c->lc->mu);
        c->lc->broken = true;
        c->lc->cv.notify_all();
    }
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_comm) cudaStreamDestroy(c->s_comm);
    delete c;
    return ADV_OK;
}

int adv_comm_unique_id(char id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (int rc = nccl_load()) return rc;
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, 128);
    return ADV_OK;
}

int adv_ctx_comm_init(adv_ctx_t* c, const char id[128])
{
    if (!c) return fail(ADV_EINVAL, "null ctx");
    if (c->npes == 1) return ADV_OK;
    CU(cudaSetDevice(c->device));
    if (int rc = nccl_load()) return rc;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    NC(g_nccl.CommInitRank(&c->comm, c->npes, u, c->mype));
    return ADV_OK;
}

int adv_ctx_comm_init_local(adv_ctx_t* const* ctxs, int n)
{
    if (!ctxs || n < 1) return fail(ADV_EINVAL, "adv_ctx_comm_init_local: bad argument");
    for (int r = 0; r < n; ++r) {
        if (!ctxs[r]) return fail(ADV_EINVAL, "adv_ctx_comm_init_local: null context");
        if (ctxs[r]->npes != n || ctxs[r]->mype != r) return fail(ADV_EINVAL, "adv_ctx_comm_init_local: context " + std::to_string(r) + " was created with mype/npes = " + std::to_string(ctxs[r]->mype) + "/" + std::to_string(ctxs[r]->npes));
        if (ctxs[r]->comm || ctxs[r]->lc) return fail(ADV_ESTATE, "adv_ctx_comm_init_local: context already has a communicator");
    }
    // what r receives from p must be what p sends to r
    for (int r = 0; r < n; ++r)
        for (const Peer& p : ctxs[r]->halo[HALO_NOD].rpeers) {
            bool ok = p.pe >= 0 && p.pe < n;
            if (ok) { ok = false; for (const Peer& sp : ctxs[p.pe]->halo[HALO_NOD].speers) ok = ok || (sp.pe == r && sp.cnt == p.cnt); }
            if (!ok) return fail(ADV_EINVAL, "adv_ctx_comm_init_local: com_nod2D of ranks " + std::to_string(r) + " and " + std::to_string(p.pe) + " do not pair up");
        }
    auto lc = std::make_shared<LocalComm>();
    lc->n = n;
    lc->ctx.assign(ctxs, ctxs + n);
    for (int r = 0; r < n; ++r) {
        for (int q = 0; q < n; ++q) {   // peer copies between different GPUs: direct access where the topology allows it
            if (ctxs[q]->device == ctxs[r]->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device) == cudaSuccess && can) {
                cudaSetDevice(ctxs[r]->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e != cudaSuccess) cudaGetLastError();   // already enabled
            }
        }
        ctxs[r]->lc = lc;
    }
    return ADV_OK;
}

// ------------------------------------------------------------------------------------------------
// ADV_HOST arrays are pageable allocatables of the Fortran host: page-lock each range the first time it is seen
// (cudaHostRegister) so that the per-step copies run at pinned-memory speed and asynchronously.  Ranges that are
// already pinned (or cannot be registered) are simply used as they are.
static void host_register(adv_ctx* c, const void* p, size_t bytes)
{
    if (!c->host_register || !p || bytes < (size_t)1 << 16) return;
    auto it = c->registered.find(p);
    if (it != c->registered.end() && it->second >= bytes) return;
    if (it != c->registered.end()) { cudaHostUnregister(const_cast<void*>(p)); c->registered.erase(it); }
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) { c->registered[p] = 0; c->registered_foreign.insert(p); return; }
    cudaGetLastError();
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) c->registered[p] = bytes;
    else { cudaGetLastError(); c->registered[p] = 0; c->registered_foreign.insert(p); }
}

int adv_ctx_set_host_register(adv_ctx_t* c, int on)
{
    if (!c) return fail(ADV_EINVAL, "null ctx");
    c->host_register = on != 0;
    return ADV_OK;
}

static int set_state_impl(adv_ctx_t* c, const adv_state_desc_t* st, int where);

int adv_ctx_set_state(adv_ctx_t* c, const adv_state_desc_t* st, int where)
{
    if (c) c->state_step = -1;
    return set_state_impl(c, st, where);
}

int adv_ctx_set_state_step(adv_ctx_t* c, const adv_state_desc_t* st, int where, int64_t step)
{
    if (!c || !st) return fail(ADV_EINVAL, "null argument");
    if (step < 0) return fail(ADV_EINVAL, "adv_ctx_set_state_step: step must be >= 0");
    const bool same = c->state_set && c->state_step == step && c->state_where == where &&
                      memcmp(&c->state_desc, st, sizeof(adv_state_desc_t)) == 0;
    if (same) return ADV_OK;            // the tracer loop of one model step: state, uploads and Q stay valid
    if (int rc = set_state_impl(c, st, where)) return rc;
    c->state_step = step; c->state_where = where;
    memcpy(&c->state_desc, st, sizeof(adv_state_desc_t));
    return ADV_OK;
}

static int set_state_impl(adv_ctx_t* c, const adv_state_desc_t* st, int where)
{
    if (!c || !st) return fail(ADV_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    MeshDev& m = c->m;
    const size_t L = m.L, nl = m.nl, Nh = m.Nh, T = m.T;
    if (!st->uv || !st->w || !st->w_e || !st->helem || !st->hnode || !st->hnode_new || !st->zbar_3d_n || !st->Z_3d_n)
        return fail(ADV_EINVAL, "adv_ctx_set_state: null field");
    if (st->use_wsplit && !st->w_i) return fail(ADV_EINVAL, "use_wsplit needs w_i");
    if (where == ADV_DEVICE) {
        m.uv = st->uv; m.helem = st->helem; m.w = st->w; m.we = st->w_e; m.wi = st->w_i;
        if ((uintptr_t)st->uv & 15u) {       // the kernels read (u,v) pairs as 16-byte words: stage an aligned copy
            if (c->uv.n != 2 * L * T) CU(c->uv.alloc(2 * L * T, false));
            CU(cudaMemcpyAsync(c->uv.p, st->uv, 2 * L * T * sizeof(double), cudaMemcpyDeviceToDevice, c->s_comp));
            m.uv = c->uv.p;
        }
        m.hnode = st->hnode; m.hnode_new = st->hnode_new; m.zbar3d = st->zbar_3d_n; m.Z3d = st->Z_3d_n;
    } else {
        struct { DevBuf<double>* b; const double* src; size_t n; const double** dst; } cp[] = {
            {&c->uv, st->uv, 2 * L * T, &m.uv}, {&c->helem, st->helem, L * T, &m.helem},
            {&c->w, st->w, nl * Nh, &m.w}, {&c->we, st->w_e, nl * Nh, &m.we},
            {&c->wi, st->w_i, st->w_i ? nl * Nh : 0, &m.wi},
            {&c->hnode, st->hnode, L * Nh, &m.hnode}, {&c->hnode_new, st->hnode_new, L * Nh, &m.hnode_new},
            {&c->zbar3d, st->zbar_3d_n, nl * Nh, &m.zbar3d}, {&c->Z3d, st->Z_3d_n, L * Nh, &m.Z3d}};
        for (auto& x : cp) {
            if (x.n == 0) { *x.dst = nullptr; continue; }
            host_register(c, x.src, x.n * sizeof(double));
            if (x.b->n != x.n) CU(x.b->alloc(x.n, false));
            CU(cudaMemcpyAsync(x.b->p, x.src, x.n * sizeof(double), cudaMemcpyHostToDevice, c->s_comp));
            *x.dst = x.b->p;
        }
    }
    m.use_wsplit = st->use_wsplit ? 1 : 0;
    c->state_set = true;
    c->q_valid = false;
    return ADV_OK;
}

// ------------------------------------------------------------------------------------------------
namespace {

struct Group { int fct, hor, ver, gs; std::vector<int> idx; };
struct ChunkSel { int fct, hor, ver, gs, tb; int idx[2]; int buf; };

// columns per CTA of the kernels that keep the fixed (column, level) thread map (non-FCT branch, gradients,
// register-gather edge kernel): as many as fit into 224 threads; one column when a column alone is longer
inline int cols_per_block(int L) { return std::max(1, 224 / L); }
inline int nblocks(int count, int cpb) { return (count + cpb - 1) / cpb; }

struct TrPtrs {   // device pointers of the call's tracers
    std::vector<const double*> ttf, ttfAB, grad, txy, gmean;   // grad == nullptr && txy != nullptr: fused gradients
    std::vector<double*> dh, dv;
    std::vector<double*> dgh, dgv;                             // ltra_diag outputs (nullptr = off)
    std::vector<double*> dvdh, dvdv;                           // ldiag_DVD outputs (nullptr = off)
};

template <int TB>
Chunk<TB> make_chunk(adv_ctx* c, const TrPtrs& p, const adv_tracer_desc_t* tr, const ChunkSel& ch)
{
    Chunk<TB> b;
    ChunkBuf& cb = *c->cbufs[ch.buf];
    for (int t = 0; t < TB; ++t) {
        const int i = ch.idx[t];
        b.ttf[t] = p.ttf[i]; b.ttfAB[t] = p.ttfAB[i]; b.grad[t] = p.grad[i]; b.txy[t] = p.txy[i]; b.gmean[t] = p.gmean[i];
        b.dttf_h[t] = p.dh[i]; b.dttf_v[t] = p.dv[i];
        b.dgh[t] = p.dgh[i]; b.dgv[t] = p.dgv[i];
        b.ph[t] = tr[i].tra_adv_ph; b.pv[t] = tr[i].tra_adv_pv;
    }
    b.lo = cb.lo.p; b.adf_h = cb.adf_h.p; b.adf_v = cb.adf_v.p; b.pm = cb.pm.p;
    b.nolo = ch.fct ? 0 : 1;
    return b;
}

enum Phase { PH_E1, PH_N1, PH_K2, PH_K3, PH_NOFCT };

// dynamic shared memory above 48 KB is an opt-in per kernel; also ask for the largest carve-out
template <class K>
static cudaError_t smem_optin(K kern, size_t bytes)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

static NodeRange node_range(const adv_ctx* c, int rid, int cpb)
{
    const MeshDev& m = c->m;
    switch (rid) {
    case R_S: return NodeRange{c->list_S.p, 0, c->nS, cpb, c->pf_dist, 0};
    // interior = all owned nodes minus the boundary set: an identity range with the flagged columns skipped
    // (no list indirection: one dependent load less per thread, block prefetch of the own columns)
    case R_I: return c->i_identity ? NodeRange{nullptr, 0, m.N, cpb, c->pf_dist, 1} : NodeRange{c->list_I.p, 0, c->nI, cpb, c->pf_dist, 0};
    case R_SH: return NodeRange{c->list_SH.p, 0, c->nSH, cpb, c->pf_dist, 0};
    case R_ALLH: return NodeRange{nullptr, 0, m.Nh, cpb, c->pf_dist, 0};
    default: return NodeRange{nullptr, 0, m.N, cpb, c->pf_dist, 0};
    }
}

// the pipelined kernels copy 16-byte cells from the caller's edge_up_dn_grad / uv (cp.async needs the
// alignment); anything else runs on the register-gather kernels
template <int TB>
static bool chunk_aligned(const MeshDev& m, const Chunk<TB>& b)
{
    uintptr_t x = (uintptr_t)m.uv;
    for (int t = 0; t < TB; ++t) x |= (uintptr_t)b.grad[t] | (uintptr_t)b.txy[t] | (uintptr_t)b.gmean[t];
    return (x & 15u) == 0;
}

static NodePart node_part(const adv_ctx* c, int rid, int threads)
{
    const adv_ctx::PartSet& ps = c->parts_for(threads);
    switch (rid) {
    case R_S: return NodePart{ps.s.first.p, c->list_S.p, ps.s.ncta, c->pf_dist, 0};
    // interior = all owned nodes minus the boundary set: an identity range in which the flagged columns get no threads
    case R_I: return NodePart{ps.inner.first.p, nullptr, ps.inner.ncta, c->pf_dist, 1};
    case R_SH: return NodePart{ps.sh.first.p, c->list_SH.p, ps.sh.ncta, c->pf_dist, 0};
    default: return NodePart{ps.all.first.p, nullptr, ps.all.ncta, c->pf_dist, 0};
    }
}

template <int TB>
int launch_phase(adv_ctx* c, Phase ph, int hor, int ver, bool q_stored, const Chunk<TB>& b, int rid, double dt)
{
    const MeshDev& m = c->m;
    cudaStream_t s = c->s_comp;
    cudaError_t se = cudaSuccess;
    bool piped = false;
    int grid = 0, nthr = 0;
    size_t smem = 0;
    if (ph == PH_E1) {
        const int epb = cols_per_block(m.L);
        nthr = epb * m.L;
        if (c->bulk && hor != HOR_UPW1 && chunk_aligned<TB>(m, b)) {
            // fixed (edge column, level) thread map: the compacted variant (groups of edges filling the CTA, measured in
            // profiles/r4_compaction.md) lost more to the extra shared-memory lookups per group than it gained in lanes
            const int ng = std::max(1, std::min(c->e1_ng, nthr / epb));
            const int nge = epb * ng, D = c->e1_depth;
            grid = nblocks(m.E, nge);
            const int gs = b.grad[0] ? 0 : 1;
#define E1B(H, Q, DD, GS) if (!piped && hor == H && (q_stored ? 1 : 0) == Q && D == DD && gs == GS) { \
                smem = e1b_smem_bytes<TB, Q, GS>(nge, nthr, DD); \
                if ((int)smem <= c->max_smem_optin) { \
                    se = smem_optin(k_edge_flux_b<H, TB, Q, DD, GS>, smem); \
                    if (se == cudaSuccess) k_edge_flux_b<H, TB, Q, DD, GS><<<grid, nthr, smem, s>>>(m, b, epb, ng, c->e1_il, c->e1_pf); \
                    piped = true; } }
#define E1BD(H, Q) E1B(H, Q, 2, 0) E1B(H, Q, 3, 0) E1B(H, Q, 4, 0) E1B(H, Q, 2, 1) E1B(H, Q, 3, 1) E1B(H, Q, 4, 1)
            E1BD(HOR_MUSCL, 0) E1BD(HOR_MUSCL, 1) E1BD(HOR_MFCT, 0) E1BD(HOR_MFCT, 1)
#undef E1BD
#undef E1B
        }
        if (!piped && hor != HOR_UPW1 && !b.grad[0]) return fail(ADV_ECUDA, "internal: fused-gradient chunk cannot run on the bulk edge kernel");
        if (!piped) {
            grid = nblocks(m.E, epb);
#define E1(H) if (hor == H) { if (q_stored) k_edge_flux<H, TB, 1><<<grid, nthr, 0, s>>>(m, b, epb); \
                              else k_edge_flux<H, TB, 0><<<grid, nthr, 0, s>>>(m, b, epb); }
            E1(HOR_UPW1) E1(HOR_MUSCL) E1(HOR_MFCT)
#undef E1
        }
    } else if (ph == PH_NOFCT) {
        const NodeRange r = node_range(c, rid, cols_per_block(m.L));
        if (r.count <= 0) return ADV_OK;
        nthr = r.cpb * m.L;
        const size_t sm1 = (size_t)(2 * TB + 6) * nthr * sizeof(double);
        grid = nblocks(r.count, r.cpb);
#define NV(V) if (ver == V) k_nofct_update<V, TB><<<grid, nthr, sm1, s>>>(m, b, r, dt);
        NV(VER_UPW1) NV(VER_QR4C) NV(VER_PPM) NV(VER_CDIFF)
#undef NV
    } else {
        // FCT node kernels: wet-level compaction, CTA partition of the range built in adv_ctx_create
        nthr = ph == PH_N1 ? c->cta_n1 : ph == PH_K2 ? c->cta_k2 : c->cta_k3;
        const NodePart r = node_part(c, rid, nthr);
        if (r.ncta <= 0) return ADV_OK;
        grid = r.ncta;
        const size_t hdr = node_smem_header(m.ell_w);
        // dynamic shared memory above 48 KB needs the opt-in attribute (set once per instantiation: cheap, idempotent)
#define NODE_LAUNCH(K, BYTES) do { smem = (BYTES); if (smem > 48 * 1024) se = smem_optin(K, smem); \
                                   if (se == cudaSuccess) K<<<grid, nthr, smem, s>>>(m, b, r, dt); } while (0)
        if (ph == PH_N1) {
#define N1(V) if (ver == V) { const size_t smn = hdr + (size_t)n1_smem_arrays<V, TB>() * nthr * sizeof(double); \
                              if (c->g_lo == 6) NODE_LAUNCH((k_node_lo<V, TB, 6>), smn); \
                              else if (c->g_lo == 2) NODE_LAUNCH((k_node_lo<V, TB, 2>), smn); \
                              else NODE_LAUNCH((k_node_lo<V, TB, 3>), smn); }
            N1(VER_UPW1) N1(VER_QR4C) N1(VER_PPM) N1(VER_CDIFF)
#undef N1
        } else if (ph == PH_K2) {
            const size_t sm2 = hdr + (size_t)(6 * TB + 1) * nthr * sizeof(double);   // tvert exchange area + parked own-column operands
            if (c->g_k2 == 6) NODE_LAUNCH((k_fct_bounds<TB, 6>), sm2);
            else if (c->g_k2 == 3) NODE_LAUNCH((k_fct_bounds<TB, 3>), sm2);
            else if (c->g_k2 == 1) NODE_LAUNCH((k_fct_bounds<TB, 1>), sm2);
            else NODE_LAUNCH((k_fct_bounds<TB, 2>), sm2);
        } else {
            if (c->g_k3 == 6) NODE_LAUNCH((k_fct_update<TB, 6>), hdr);
            else if (c->g_k3 == 3) NODE_LAUNCH((k_fct_update<TB, 3>), hdr);
            else if (c->g_k3 == 1) NODE_LAUNCH((k_fct_update<TB, 1>), hdr);
            else NODE_LAUNCH((k_fct_update<TB, 2>), hdr);
        }
#undef NODE_LAUNCH
    }
    ++c->launches;
    static const char* names[] = {"k_edge_flux", "k_node_lo", "k_fct_bounds", "k_fct_update", "k_nofct_update"};
    if (se != cudaSuccess) return fail(ADV_ECUDA, std::string(names[ph]) + " attributes: " + cudaGetErrorString(se));
    if (cudaError_t e = cudaGetLastError()) {
        return fail(ADV_ECUDA, std::string("launch ") + names[ph] + (piped ? "_b" : "") + " (grid " + std::to_string(grid) + ", block " + std::to_string(nthr) +
                                   ", smem " + std::to_string(smem) + ", hor " + std::to_string(hor) + ", ver " + std::to_string(ver) + ", tb " + std::to_string(TB) + "): " + cudaGetErrorString(e));
    }
    return ADV_OK;
}

// one exchange_nod / exchange_elem for `nf` fields of nlev[f] doubles per column over the halo set `kind`:
// pack the send columns per field, move them (NCCL: one grouped send/recv per (peer, field); in-process
// communicator: one device-to-device copy per (peer, field), pulled by the receiver), unpack where the
// received columns are not a contiguous tail.  recvbufs may be null when the set has a contiguous tail.
int halo_exchange(adv_ctx* c, int kind, cudaStream_t s, int nf, double* const* fields, double* const* sendbufs,
                  double* const* recvbufs, const int* nlev)
{
    HaloSet& h = c->halo[kind];
    const int cols = h.send_cols;
    const bool direct = h.recv_base >= 0;
    if (!direct && !recvbufs) return fail(ADV_EINVAL, "internal: halo set needs receive buffers");
    for (int f = 0; f < nf; ++f) {
        if (cols > 0) {
            k_pack_halo<<<cols, std::min(kBlock, ((nlev[f] + 31) / 32) * 32), 0, s>>>(fields[f], h.slist.p, nlev[f], sendbufs[f]);
            ++c->launches;
        }
        c->halo_bytes_sent += (int64_t)cols * nlev[f] * 8;
    }
    ++c->halo_exchanges;
    auto dst_of = [&](int f, const Peer& p) {
        return direct ? fields[f] + ((size_t)h.recv_base + p.off) * nlev[f] : recvbufs[f] + (size_t)p.off * nlev[f];
    };
    if (c->lc) {
        LocalComm& lc = *c->lc;
        CU(cudaEventRecord(c->ev_packed, s));
        c->pub_send.assign(sendbufs, sendbufs + nf);
        if (!lc.barrier()) return fail(ADV_ESTATE, "in-process communicator: the peers did not reach the exchange (every context needs its own host thread and the same call sequence)");
        for (const Peer& p : h.rpeers) {
            adv_ctx* q = lc.ctx[p.pe];
            CU(cudaStreamWaitEvent(s, q->ev_packed, 0));
            int off = -1;
            for (const Peer& sp : q->halo[kind].speers) if (sp.pe == c->mype) { off = sp.off; if (sp.cnt != p.cnt) off = -1; }
            if (off < 0 || (int)q->pub_send.size() != nf) return fail(ADV_EINVAL, "in-process communicator: send/recv lists of ranks " + std::to_string(c->mype) + " and " + std::to_string(p.pe) + " do not pair up");
            for (int f = 0; f < nf; ++f)
                CU(cudaMemcpyAsync(dst_of(f, p), q->pub_send[f] + (size_t)off * nlev[f], (size_t)p.cnt * nlev[f] * 8, cudaMemcpyDefault, s));
        }
        CU(cudaEventRecord(c->ev_consumed, s));
        if (!lc.barrier()) return fail(ADV_ESTATE, "in-process communicator: the peers did not finish the exchange");
        for (const Peer& p : h.speers) CU(cudaStreamWaitEvent(s, lc.ctx[p.pe]->ev_consumed, 0));   // the send buffer may be packed again
    } else {
        NC(g_nccl.GroupStart());
        for (int f = 0; f < nf; ++f) {
            for (const Peer& p : h.rpeers)
                NC(g_nccl.Recv(dst_of(f, p), (size_t)p.cnt * nlev[f], ncclDouble, p.pe, c->comm, s));
            for (const Peer& p : h.speers)
                NC(g_nccl.Send(sendbufs[f] + (size_t)p.off * nlev[f], (size_t)p.cnt * nlev[f], ncclDouble, p.pe, c->comm, s));
        }
        NC(g_nccl.GroupEnd());
    }
    if (!direct && h.recv_cols > 0) {
        for (int f = 0; f < nf; ++f) {
            k_unpack_halo<<<h.recv_cols, std::min(kBlock, ((nlev[f] + 31) / 32) * 32), 0, s>>>(recvbufs[f], h.rlist.p, nlev[f], fields[f]);
            ++c->launches;
        }
    }
    CU(cudaGetLastError());
    return ADV_OK;
}

}  // namespace

static int ensure_chunk_bufs(adv_ctx* c, int count)
{
    const MeshDev& m = c->m;
    while ((int)c->cbufs.size() < count) {
        c->cbufs.emplace_back(new ChunkBuf());
        ChunkBuf& b = *c->cbufs.back();
        CU(b.lo.alloc((size_t)m.L * m.Nh * 2)); CU(b.pm.alloc((size_t)m.L * m.Nh * 4));
        CU(b.adf_h.alloc((size_t)m.L * m.E * 2)); CU(b.adf_v.alloc((size_t)m.nl * m.N * 2));
        if (c->npes > 1) CU(b.sendbuf.alloc((size_t)c->halo[HALO_NOD].send_cols * m.L * 4));
    }
    return ADV_OK;
}

static int run_batch(adv_ctx* c, double dt, int ntr, const adv_tracer_desc_t* tr, const TrPtrs& p)
{
    static const char* hn[] = {"MUSCL", "MFCT", "UPW1"};
    static const int hc[] = {HOR_MUSCL, HOR_MFCT, HOR_UPW1};
    static const char* vn[] = {"QR4C", "CDIFF", "PPM", "UPW1"};
    static const int vc[] = {VER_QR4C, VER_CDIFF, VER_PPM, VER_UPW1};
    MeshDev& m = c->m;
    std::vector<Group> groups;
    for (int i = 0; i < ntr; ++i) {
        const int hor = parse_scheme(tr[i].tra_adv_hor, hn, hc, 3);
        const int ver = parse_scheme(tr[i].tra_adv_ver, vn, vc, 4);
        if (hor < 0) return fail(ADV_ESCHEME, std::string("Unknown horizontal advection type ") + (tr[i].tra_adv_hor ? tr[i].tra_adv_hor : "(null)"));
        if (ver < 0) return fail(ADV_ESCHEME, std::string("Unknown vertical advection type ") + (tr[i].tra_adv_ver ? tr[i].tra_adv_ver : "(null)"));
        static const char* ln[] = {"FCT"};
        static const int lc[] = {1};
        const int fct = parse_scheme(tr[i].tra_adv_lim, ln, lc, 1) == 1 ? 1 : 0;   // driver :111: anything else = no limiter
        if (hor != HOR_UPW1 && !p.grad[i] && !p.txy[i]) return fail(ADV_EINVAL, "edge_up_dn_grad is NULL for a gradient-based scheme");
        const int gs = (hor != HOR_UPW1 && !p.grad[i]) ? 1 : 0;
        if (gs && !fct) return fail(ADV_EINVAL, "internal: fused gradients are for FCT tracers only");
        bool found = false;
        for (auto& g : groups)
            if (g.fct == fct && g.hor == hor && g.ver == ver && g.gs == gs) { g.idx.push_back(i); found = true; break; }
        if (!found) groups.push_back(Group{fct, hor, ver, gs, {i}});
    }
    std::vector<ChunkSel> chunks;
    for (auto& g : groups) {
        size_t i = 0;
        for (; !c->force_tb1 && i + 2 <= g.idx.size(); i += 2) chunks.push_back(ChunkSel{g.fct, g.hor, g.ver, g.gs, 2, {g.idx[i], g.idx[i + 1]}, 0});
        for (; i < g.idx.size(); ++i) chunks.push_back(ChunkSel{g.fct, g.hor, g.ver, g.gs, 1, {g.idx[i], g.idx[i]}, 0});
    }
    if (int rc = ensure_chunk_bufs(c, (int)chunks.size())) return rc;
    for (size_t k = 0; k < chunks.size(); ++k) {
        chunks[k].buf = (int)k;
        // the kernels only write wet levels and rely on the zeros of the allocation below the sea floor; a buffer
        // reused with another interleave (TB 1 <-> 2) would show stale values there: clear it
        ChunkBuf& cb = *c->cbufs[k];
        if (cb.tb_last != 0 && cb.tb_last != chunks[k].tb) {
            for (DevBuf<double>* x : {&cb.lo, &cb.pm, &cb.adf_h, &cb.adf_v}) CU(cudaMemsetAsync(x->p, 0, x->n * sizeof(double), c->s_comp));
        }
        cb.tb_last = chunks[k].tb;
        for (int t = 0; t < chunks[k].tb; ++t) c->trloc[chunks[k].idx[t]] = TrLoc{(int)k, t, chunks[k].tb};
    }
    cudaStream_t sc = c->s_comp, sx = c->s_comm;
    const int cpb = cols_per_block(m.L);
    const bool multi = c->npes > 1;
    if (multi && !c->comm && !c->lc) return fail(ADV_ESTATE, "npes > 1 but neither adv_ctx_comm_init nor adv_ctx_comm_init_local was called");
    const bool prof = c->profiling && !multi;
    auto mark = [&](int i) { if (prof) cudaEventRecord(c->ev_ph[i], sc); };

    const int rAll = R_ALL, rS = R_S, rI = R_I, rSH = R_SH, rAllH = R_ALLH;
    int launch_rc = ADV_OK;
    std::string launch_msg;
    auto run = [&](Phase ph, const ChunkSel& ch, int r) {
        const bool qs = c->q_valid;
        int rc;
        if (ch.tb == 2) rc = launch_phase<2>(c, ph, ch.hor, ch.ver, qs, make_chunk<2>(c, p, tr, ch), r, dt);
        else rc = launch_phase<1>(c, ph, ch.hor, ch.ver, qs, make_chunk<1>(c, p, tr, ch), r, dt);
        if (rc && !launch_rc) { launch_rc = rc; launch_msg = g_err; }
        if (ph == PH_E1) c->q_valid = true;
    };
    std::vector<double*> f1, s1, f2, s2;   // exchange field / send buffer lists over all FCT chunks
    std::vector<int> n1, n2;
    bool any_fct = false;
    for (auto& ch : chunks) {
        if (ch.fct) {
            any_fct = true;
            ChunkBuf& cb = *c->cbufs[ch.buf];
            f1.push_back(cb.lo.p); s1.push_back(cb.sendbuf.p); n1.push_back(m.L * ch.tb);
            f2.push_back(cb.pm.p); s2.push_back(cb.sendbuf.p); n2.push_back(m.L * ch.tb * 2);
        }
    }
    mark(0);
    // ---- phase 0: edge fluxes, every chunk (FCT: antidiffusive HO - LO; otherwise the HO flux itself); the first
    //      launch of a step also produces Q
    for (auto& ch : chunks) run(PH_E1, ch, rAll);
    // non-FCT tracers: one more sweep, no exchange inside the path
    for (auto& ch : chunks)
        if (!ch.fct) run(PH_NOFCT, ch, multi ? rAllH : rAll);
    mark(1);
    if (any_fct) {
        const bool overlap = multi;
        // adv_tra_vert_impl on fct_LO (driver :320-333, use_wsplit only), same node range as the LO launch before it
        auto vimpl = [&](int rid) {
            if (!m.use_wsplit) return;
            for (auto& ch : chunks) {
                if (!ch.fct) continue;
                const NodeRange r = node_range(c, rid, cpb);
                if (r.count <= 0) continue;
                const int nthr = r.cpb * m.L;
                double* lo = c->cbufs[ch.buf]->lo.p;
                if (ch.tb == 2) k_vert_impl<2><<<nblocks(r.count, r.cpb), nthr, (size_t)(3 + 2) * nthr * sizeof(double), sc>>>(m, lo, r, dt);
                else k_vert_impl<1><<<nblocks(r.count, r.cpb), nthr, (size_t)(3 + 1) * nthr * sizeof(double), sc>>>(m, lo, r, dt);
                ++c->launches;
            }
        };
        const int64_t bytes0 = c->halo_bytes_sent;
        if (overlap) {
            for (auto& ch : chunks) if (ch.fct) run(PH_N1, ch, rS);
            vimpl(rS);
            CU(cudaEventRecord(c->ev_a, sc));
            CU(cudaStreamWaitEvent(sx, c->ev_a, 0));
            CU(cudaEventRecord(c->ev_x0[0], sx));
            if (int rc = halo_exchange(c, HALO_NOD, sx, (int)f1.size(), f1.data(), s1.data(), nullptr, n1.data())) return rc;   // driver :335
            CU(cudaEventRecord(c->ev_x1[0], sx));
            CU(cudaEventRecord(c->ev_b, sx));
            for (auto& ch : chunks) if (ch.fct) run(PH_N1, ch, rI);
            vimpl(rI);
            CU(cudaEventRecord(c->ev_w0[0], sc));
            CU(cudaStreamWaitEvent(sc, c->ev_b, 0));
            CU(cudaEventRecord(c->ev_w1[0], sc));
        } else {
            for (auto& ch : chunks) if (ch.fct) run(PH_N1, ch, rAll);
            vimpl(rAll);
        }
        mark(2);
        // ---- phase 2: bounds + R+/R- (boundary set first, then start exchange 2)
        if (multi) {
            for (auto& ch : chunks) if (ch.fct) run(PH_K2, ch, rS);
            CU(cudaEventRecord(c->ev_c, sc));
            CU(cudaStreamWaitEvent(sx, c->ev_c, 0));
            CU(cudaEventRecord(c->ev_x0[1], sx));
            if (int rc = halo_exchange(c, HALO_NOD, sx, (int)f2.size(), f2.data(), s2.data(), nullptr, n2.data())) return rc;   // fct :413
            CU(cudaEventRecord(c->ev_x1[1], sx));
            CU(cudaEventRecord(c->ev_d, sx));
            for (auto& ch : chunks) if (ch.fct) run(PH_K2, ch, rI);
            // ---- phase 3: interior update overlaps exchange 2
            for (auto& ch : chunks) if (ch.fct) run(PH_K3, ch, rI);
            CU(cudaEventRecord(c->ev_w0[1], sc));
            CU(cudaStreamWaitEvent(sc, c->ev_d, 0));
            CU(cudaEventRecord(c->ev_w1[1], sc));
            for (auto& ch : chunks) if (ch.fct) run(PH_K3, ch, rSH);
            c->halo_bytes_last = c->halo_bytes_sent - bytes0;
            c->x_valid = true;
        } else {
            for (auto& ch : chunks) if (ch.fct) run(PH_K2, ch, rAll);
            mark(3);
            for (auto& ch : chunks) if (ch.fct) run(PH_K3, ch, rAll);
        }
    } else { mark(2); mark(3); }
    // ---- ldiag_DVD (driver :263-296, :395-458): two optional sweeps over the edges / the owned nodes of the chunks that asked
    for (auto& ch : chunks) {
        bool any = false;
        for (int t = 0; t < ch.tb; ++t) any = any || p.dvdh[ch.idx[t]] || p.dvdv[ch.idx[t]];
        if (!any) continue;
        const NodeRange rn = node_range(c, rAll, cpb);
        auto go = [&](auto chunk) {
            constexpr int TBc = sizeof(chunk.ttf) / sizeof(chunk.ttf[0]);
            DvdPtrs<TBc> d;
            for (int t = 0; t < TBc; ++t) { d.hor[t] = p.dvdh[ch.idx[t]]; d.ver[t] = p.dvdv[ch.idx[t]]; }
            k_dvd_hor<TBc><<<nblocks(m.E, cpb), cpb * m.L, 0, sc>>>(m, chunk, d, cpb, ch.fct);
            k_dvd_ver<TBc><<<nblocks(rn.count, rn.cpb), rn.cpb * m.L, 0, sc>>>(m, chunk, d, rn, ch.fct);
            c->launches += 2;
        };
        if (ch.tb == 2) go(make_chunk<2>(c, p, tr, ch)); else go(make_chunk<1>(c, p, tr, ch));
    }
    mark(4);
    c->ph_valid = prof;
    if (launch_rc) { cudaGetLastError(); return fail(launch_rc, launch_msg); }
    CU(cudaGetLastError());
    return ADV_OK;
}

static int do_adv(adv_ctx* c, double dt, int ntr, const adv_tracer_desc_t* tr, int where, bool blocking)
{
    if (!c || !tr || ntr < 1) return fail(ADV_EINVAL, "null argument");
    if (ntr > c->max_tr) return fail(ADV_EINVAL, "ntr exceeds max_tracers of the context");
    if (!c->state_set) return fail(ADV_ESTATE, "adv_ctx_set_state has not been called");
    CU(cudaSetDevice(c->device));
    MeshDev& m = c->m;
    const size_t nLN = (size_t)m.L * m.Nh, nLE = (size_t)m.L * m.E;
    TrPtrs p;
    p.ttf.resize(ntr); p.ttfAB.resize(ntr); p.grad.resize(ntr); p.txy.assign(ntr, nullptr); p.gmean.assign(ntr, nullptr); p.dh.resize(ntr); p.dv.resize(ntr);
    p.dgh.assign(ntr, nullptr); p.dgv.assign(ntr, nullptr); p.dvdh.assign(ntr, nullptr); p.dvdv.assign(ntr, nullptr);
    for (int i = 0; i < ntr; ++i) {
        if (!tr[i].values || !tr[i].valuesAB || !tr[i].del_ttf_advhoriz || !tr[i].del_ttf_advvert)
            return fail(ADV_EINVAL, "tracer " + std::to_string(i + 1) + ": null field");
        if (where == ADV_DEVICE) {
            p.ttf[i] = tr[i].values; p.ttfAB[i] = tr[i].valuesAB; p.grad[i] = tr[i].edge_up_dn_grad;
            p.dh[i] = tr[i].del_ttf_advhoriz; p.dv[i] = tr[i].del_ttf_advvert;
            p.dgh[i] = tr[i].tra_advhoriz; p.dgv[i] = tr[i].tra_advvert;
            p.dvdh[i] = tr[i].dvd_trflx_hor; p.dvdv[i] = tr[i].dvd_trflx_ver;
            if (tr[i].edge_up_dn_grad && ((uintptr_t)tr[i].edge_up_dn_grad & 15u)) {
                // edge_up_dn_grad(1:4,nz,e) is read as 16-byte words / bulk copies: stage an aligned copy
                Slot& s = c->slots[i];
                if (s.grad.n != 4 * nLE) CU(s.grad.alloc(4 * nLE, false));
                CU(cudaMemcpyAsync(s.grad.p, tr[i].edge_up_dn_grad, 4 * nLE * 8, cudaMemcpyDeviceToDevice, c->s_comp));
                p.grad[i] = s.grad.p;
            }
        } else {
            Slot& s = c->slots[i];
            host_register(c, tr[i].values, nLN * 8); host_register(c, tr[i].valuesAB, nLN * 8);
            host_register(c, tr[i].del_ttf_advhoriz, nLN * 8); host_register(c, tr[i].del_ttf_advvert, nLN * 8);
            if (tr[i].edge_up_dn_grad) host_register(c, tr[i].edge_up_dn_grad, 4 * nLE * 8);
            if (s.ttf.n != nLN) { CU(s.ttf.alloc(nLN, false)); CU(s.ttfAB.alloc(nLN, false)); CU(s.dh.alloc(nLN, false)); CU(s.dv.alloc(nLN, false)); }
            CU(cudaMemcpyAsync(s.ttf.p, tr[i].values, nLN * 8, cudaMemcpyHostToDevice, c->s_comp));
            CU(cudaMemcpyAsync(s.ttfAB.p, tr[i].valuesAB, nLN * 8, cudaMemcpyHostToDevice, c->s_comp));
            CU(cudaMemcpyAsync(s.dh.p, tr[i].del_ttf_advhoriz, nLN * 8, cudaMemcpyHostToDevice, c->s_comp));
            CU(cudaMemcpyAsync(s.dv.p, tr[i].del_ttf_advvert, nLN * 8, cudaMemcpyHostToDevice, c->s_comp));
            p.grad[i] = nullptr;
            if (tr[i].edge_up_dn_grad) {
                if (s.grad.n != 4 * nLE) CU(s.grad.alloc(4 * nLE, false));
                CU(cudaMemcpyAsync(s.grad.p, tr[i].edge_up_dn_grad, 4 * nLE * 8, cudaMemcpyHostToDevice, c->s_comp));
                p.grad[i] = s.grad.p;
            }
            p.ttf[i] = s.ttf.p; p.ttfAB[i] = s.ttfAB.p; p.dh[i] = s.dh.p; p.dv[i] = s.dv.p;
            // ltra_diag arrays: staged both ways (the entries the path does not write keep the caller's values)
            double* const hd[2] = {tr[i].tra_advhoriz, tr[i].tra_advvert};
            DevBuf<double>* const sd[2] = {&s.dgh, &s.dgv};
            for (int k = 0; k < 2; ++k) {
                if (!hd[k]) continue;
                host_register(c, hd[k], nLN * 8);
                if (sd[k]->n != nLN) CU(sd[k]->alloc(nLN, false));
                CU(cudaMemcpyAsync(sd[k]->p, hd[k], nLN * 8, cudaMemcpyHostToDevice, c->s_comp));
                (k == 0 ? p.dgh[i] : p.dgv[i]) = sd[k]->p;
            }
            // ldiag_DVD arrays: every entry is written by the library, so they are only copied back
            if (tr[i].dvd_trflx_hor) { host_register(c, tr[i].dvd_trflx_hor, nLE * 8); if (s.dvdh.n != nLE) CU(s.dvdh.alloc(nLE, false)); p.dvdh[i] = s.dvdh.p; }
            if (tr[i].dvd_trflx_ver) {
                const size_t nNN = (size_t)m.nl * m.N;
                host_register(c, tr[i].dvd_trflx_ver, nNN * 8); if (s.dvdv.n != nNN) CU(s.dvdv.alloc(nNN, false)); p.dvdv[i] = s.dvdv.p;
            }
        }
    }
    // edge_up_dn_grad == NULL for a gradient-based scheme: the library runs the caller's tracer_gradient_elements,
    // exchange_elem(tr_xy) and fill_up_dn_grad itself (adv_ctx_set_gradient_mesh).  Saves the 4 E L words of
    // H2D per tracer, and the caller's own gradient sweeps.
    {
        std::vector<int> need;
        for (int i = 0; i < ntr; ++i) {
            if (p.grad[i] || !c->grad_mesh_set) continue;
            char hbuf[16]; int k = 0;
            for (const char* q = tr[i].tra_adv_hor; q && *q && *q != ' ' && k < 15; ++q) hbuf[k++] = *q;
            hbuf[k] = 0;
            if (strcmp(hbuf, "UPW1") != 0) need.push_back(i);
        }
        if (!need.empty()) {
            if (c->npes > 1 && !c->elem_halo_set)
                return fail(ADV_EINVAL, "edge_up_dn_grad = NULL on more than one rank needs com_elem2D_full in adv_ctx_set_gradient_mesh");
            const size_t nxy = (size_t)2 * m.L * c->gm.n_elem, nmean = (size_t)2 * m.L * m.Nh;
            const int cpb = cols_per_block(m.L);
            std::vector<const double*> ttfs, cxy;
            std::vector<double*> xy, gr;
            std::vector<int> fused;
            for (int i : need) {
                Slot& s = c->slots[i];
                if (s.tr_xy.n != nxy) CU(s.tr_xy.alloc(nxy));
                ttfs.push_back(p.ttf[i]); xy.push_back(s.tr_xy.p);
                // FCT tracers: the edge kernel reconstructs the gradients itself (k_edge_flux_b<.., GS = 1>) from tr_xy
                // and the node means; the one-sweep non-FCT kernel reads a materialised edge_up_dn_grad
                char lbuf[16]; int k = 0;
                for (const char* q = tr[i].tra_adv_lim; q && *q && *q != ' ' && k < 15; ++q) lbuf[k++] = *q;
                lbuf[k] = 0;
                const bool fuse = c->fuse_grad && c->bulk && strcmp(lbuf, "FCT") == 0 && c->gm.n_nie >= m.Nh;
                if (fuse) fused.push_back(i);
                else {
                    if (s.grad.n != 4 * nLE) CU(s.grad.alloc(4 * nLE));
                    cxy.push_back(s.tr_xy.p); gr.push_back(s.grad.p);
                    p.grad[i] = s.grad.p;
                }
            }
            if (int rc = adv_tracer_gradient_elements(c, (int)need.size(), ttfs.data(), xy.data())) return rc;
            if (c->npes > 1) if (int rc = adv_exchange_elem(c, (int)need.size(), xy.data(), 2 * m.L)) return rc;
            if (!cxy.empty()) if (int rc = adv_fill_up_dn_grad(c, (int)cxy.size(), cxy.data(), gr.data())) return rc;
            for (size_t k = 0; k < fused.size();) {           // node means: two tracers per launch share the element walk
                const int i0 = fused[k], i1 = k + 1 < fused.size() ? fused[k + 1] : -1;
                for (int i : {i0, i1}) {
                    if (i < 0) continue;
                    Slot& s = c->slots[i];
                    if (s.gmean.n != nmean) CU(s.gmean.alloc(nmean));
                    p.txy[i] = s.tr_xy.p; p.gmean[i] = s.gmean.p;
                }
                if (i1 >= 0) {
                    PtrPack<2> pk{{c->slots[i0].tr_xy.p, c->slots[i1].tr_xy.p}, {c->slots[i0].gmean.p, c->slots[i1].gmean.p}};
                    k_node_mean_grad<2><<<nblocks(m.Nh, cpb), cpb * m.L, 0, c->s_comp>>>(m, c->gm, cpb, pk);
                    k += 2;
                } else {
                    PtrPack<1> pk{{c->slots[i0].tr_xy.p}, {c->slots[i0].gmean.p}};
                    k_node_mean_grad<1><<<nblocks(m.Nh, cpb), cpb * m.L, 0, c->s_comp>>>(m, c->gm, cpb, pk);
                    k += 1;
                }
                ++c->launches;
            }
            CU(cudaGetLastError());
        }
    }
    CU(cudaEventRecord(c->ev_t0, c->s_comp));
    if (int rc = run_batch(c, dt, ntr, tr, p)) return rc;
    CU(cudaEventRecord(c->ev_t1, c->s_comp));
    c->timed = true;
    if (where == ADV_HOST) {
        for (int i = 0; i < ntr; ++i) {
            CU(cudaMemcpyAsync(tr[i].del_ttf_advhoriz, p.dh[i], nLN * 8, cudaMemcpyDeviceToHost, c->s_comp));
            CU(cudaMemcpyAsync(tr[i].del_ttf_advvert, p.dv[i], nLN * 8, cudaMemcpyDeviceToHost, c->s_comp));
            if (tr[i].tra_advhoriz) CU(cudaMemcpyAsync(tr[i].tra_advhoriz, p.dgh[i], nLN * 8, cudaMemcpyDeviceToHost, c->s_comp));
            if (tr[i].tra_advvert) CU(cudaMemcpyAsync(tr[i].tra_advvert, p.dgv[i], nLN * 8, cudaMemcpyDeviceToHost, c->s_comp));
            if (tr[i].dvd_trflx_hor) CU(cudaMemcpyAsync(tr[i].dvd_trflx_hor, p.dvdh[i], nLE * 8, cudaMemcpyDeviceToHost, c->s_comp));
            if (tr[i].dvd_trflx_ver) CU(cudaMemcpyAsync(tr[i].dvd_trflx_ver, p.dvdv[i], (size_t)m.nl * m.N * 8, cudaMemcpyDeviceToHost, c->s_comp));
        }
    }
    if (blocking) CU(cudaStreamSynchronize(c->s_comp));
    return ADV_OK;
}

int adv_do_oce_adv_tra(adv_ctx_t* c, double dt, int ntr, const adv_tracer_desc_t* tr, int where)
{
    return do_adv(c, dt, ntr, tr, where, true);
}

int adv_do_oce_adv_tra_async(adv_ctx_t* c, double dt, int ntr, const adv_tracer_desc_t* tr)
{
    return do_adv(c, dt, ntr, tr, ADV_DEVICE, false);
}

int adv_ctx_wait_for(adv_ctx_t* c, void* stream)
{
    if (!c) return fail(ADV_EINVAL, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev_order, (cudaStream_t)stream));
    CU(cudaStreamWaitEvent(c->s_comp, c->ev_order, 0));
    return ADV_OK;
}

int adv_ctx_signal(adv_ctx_t* c, void* stream)
{
    if (!c) return fail(ADV_EINVAL, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev_order, c->s_comp));
    CU(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_order, 0));
    return ADV_OK;
}

int adv_ctx_synchronize(adv_ctx_t* c)
{
    if (!c) return fail(ADV_EINVAL, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_comp));
    CU(cudaStreamSynchronize(c->s_comm));
    return ADV_OK;
}

static int exchange_fields(adv_ctx_t* c, int kind, int nfields, double* const* fields, int nlev)
{
    CU(cudaSetDevice(c->device));
    const HaloSet& h = c->halo[kind];
    const size_t need = (size_t)nfields * h.send_cols * nlev, needr = h.recv_base >= 0 ? 0 : (size_t)nfields * h.recv_cols * nlev;
    if (c->xbuf.n < need || c->xrbuf.n < needr) {
        CU(cudaStreamSynchronize(c->s_comp));
        if (c->xbuf.n < need) CU(c->xbuf.alloc(need, false));
        if (c->xrbuf.n < needr) CU(c->xrbuf.alloc(needr, false));
    }
    std::vector<double*> sb(nfields), rb(nfields);
    for (int f = 0; f < nfields; ++f) { sb[f] = c->xbuf.p + (size_t)f * h.send_cols * nlev; rb[f] = c->xrbuf.p + (size_t)f * h.recv_cols * nlev; }
    std::vector<int> nl(nfields, nlev);
    return halo_exchange(c, kind, c->s_comp, nfields, fields, sb.data(), h.recv_base >= 0 ? nullptr : rb.data(), nl.data());
}

int adv_exchange_nod(adv_ctx_t* c, int nfields, double* const* fields, int nlev)
{
    if (!c || !fields || nfields < 1 || nlev < 1) return fail(ADV_EINVAL, "bad argument");
    if (c->npes == 1) return ADV_OK;
    if (!c->comm && !c->lc) return fail(ADV_ESTATE, "adv_ctx_comm_init was not called");
    return exchange_fields(c, HALO_NOD, nfields, fields, nlev);
}

int adv_exchange_elem(adv_ctx_t* c, int nfields, double* const* fields, int nwords)
{
    if (!c || !fields || nfields < 1 || nwords < 1) return fail(ADV_EINVAL, "bad argument");
    if (c->npes == 1) return ADV_OK;
    if (!c->comm && !c->lc) return fail(ADV_ESTATE, "adv_ctx_comm_init was not called");
    if (!c->grad_mesh_set || !c->elem_halo_set) return fail(ADV_ESTATE, "adv_ctx_set_gradient_mesh with com_elem2D_full has not been called");
    return exchange_fields(c, HALO_ELEM, nfields, fields, nwords);
}

static int vert_vel_ale_impl(adv_ctx_t* c, double dt, int use_wsplit, double wsplit_maxcfl, const adv_zstar_desc_t* z,
                             double* w, double* w_e, double* w_i, double* cfl_z, const adv_zlevel_desc_t* zl = nullptr)
{
    if (!c || !w || !w_e || !w_i) return fail(ADV_EINVAL, "adv_vert_vel_ale: null argument");
    if (zl) {
        if (!zl->hbar || !zl->hbar_old || !zl->water_flux || !zl->nlevels_nod2D_min || !zl->hnode_new || !zl->zbar)
            return fail(ADV_EINVAL, "adv_vert_vel_ale_zlevel: null field in the zlevel descriptor");
        if (!cfl_z) return fail(ADV_EINVAL, "adv_vert_vel_ale_zlevel: cfl_z must hold the previous step's CFL_z (it is read before it is recomputed)");
        if (zl->lzstar_lev < 1 || zl->lzstar_lev > kMaxLzstar || zl->lzstar_lev > c->m.L)
            return fail(ADV_EINVAL, "adv_vert_vel_ale_zlevel: lzstar_lev out of range (1 .. min(16, nl-1))");
    }
    if (z && (!z->hbar || !z->hbar_old || !z->water_flux || !z->nlevels_nod2D_min || !z->hnode_new))
        return fail(ADV_EINVAL, "adv_vert_vel_ale_zstar: null field in the zstar descriptor");
    if (!c->state_set) return fail(ADV_ESTATE, "adv_ctx_set_state has not been called");
    if (c->npes > 1 && !c->comm && !c->lc) return fail(ADV_ESTATE, "adv_ctx_comm_init was not called");
    CU(cudaSetDevice(c->device));
    const MeshDev& m = c->m;
    const int cpb = cols_per_block(m.L);
    const NodeRange rAll{nullptr, 0, m.N, cpb, 0, 0};
    k_vert_vel_ale<<<nblocks(m.N, cpb), cpb * m.L, (size_t)cpb * m.L * sizeof(double), c->s_comp>>>(m, rAll, w);
    ++c->launches;
    if (z) {                                                      // which_ALE = 'zstar', src/oce_ale.F90:2539-2603
        k_vert_vel_zstar<<<nblocks(m.N, cpb), cpb * m.L, 0, c->s_comp>>>(m, rAll, dt, z->nlevels_nod2D_min, z->hbar, z->hbar_old,
                                                                         z->water_flux, w, z->hnode_new);
        ++c->launches;
    }
    if (zl) {                                                     // which_ALE = 'zlevel', src/oce_ale.F90:2336-2538
        k_vert_vel_zlevel<<<nblocks(m.N, 256), 256, 0, c->s_comp>>>(m, dt, zl->nlevels_nod2D_min, zl->hbar, zl->hbar_old, zl->water_flux,
                                                                    zl->zbar, cfl_z, zl->min_hnode, zl->lzstar_lev, w, zl->hnode_new);
        ++c->launches;
    }
    CU(cudaGetLastError());
    double* hnew = z ? z->hnode_new : zl ? zl->hnode_new : nullptr;
    if (c->npes > 1) {                                            // exchange_nod(Wvel), exchange_nod(hnode_new): :2654-2655
        double* f[1] = {w};
        if (int rc = exchange_fields(c, HALO_NOD, 1, f, m.nl)) return rc;
        if (hnew) { double* h[1] = {hnew}; if (int rc = exchange_fields(c, HALO_NOD, 1, h, m.L)) return rc; }
    }
    const size_t n = (size_t)m.Nh * m.nl;
    k_cflz_wsplit<<<(unsigned)((n + 255) / 256), 256, 0, c->s_comp>>>(m, hnew ? hnew : m.hnode_new, dt, use_wsplit ? 1 : 0, wsplit_maxcfl,
                                                                     w, w_e, w_i, cfl_z);
    ++c->launches;
    CU(cudaGetLastError());
    return ADV_OK;
}

int adv_vert_vel_ale(adv_ctx_t* c, double dt, int use_wsplit, double wsplit_maxcfl, double* w, double* w_e, double* w_i, double* cfl_z)
{
    return vert_vel_ale_impl(c, dt, use_wsplit, wsplit_maxcfl, nullptr, w, w_e, w_i, cfl_z);
}

int adv_vert_vel_ale_zstar(adv_ctx_t* c, double dt, int use_wsplit, double wsplit_maxcfl, const adv_zstar_desc_t* z,
                           double* w, double* w_e, double* w_i, double* cfl_z)
{
    if (!z) return fail(ADV_EINVAL, "adv_vert_vel_ale_zstar: null descriptor");
    return vert_vel_ale_impl(c, dt, use_wsplit, wsplit_maxcfl, z, w, w_e, w_i, cfl_z);
}

int adv_vert_vel_ale_zlevel(adv_ctx_t* c, double dt, int use_wsplit, double wsplit_maxcfl, const adv_zlevel_desc_t* z,
                            double* w, double* w_e, double* w_i, double* cfl_z)
{
    if (!z) return fail(ADV_EINVAL, "adv_vert_vel_ale_zlevel: null descriptor");
    return vert_vel_ale_impl(c, dt, use_wsplit, wsplit_maxcfl, nullptr, w, w_e, w_i, cfl_z, z);
}

int adv_update_values(adv_ctx_t* c, int ntr, double* const* values, const double* const* dh, const double* const* dv)
{
    if (!c || !values || !dh || !dv || ntr < 1) return fail(ADV_EINVAL, "bad argument");
    if (!c->state_set) return fail(ADV_ESTATE, "adv_ctx_set_state has not been called");
    CU(cudaSetDevice(c->device));
    const int cpb = cols_per_block(c->m.L);
    const NodeRange rAll{nullptr, 0, c->m.N, cpb, 0, 0};
    for (int i = 0; i < ntr; ++i) {
        k_update_values<<<nblocks(c->m.N, cpb), cpb * c->m.L, 0, c->s_comp>>>(c->m, rAll, values[i], dh[i], dv[i]);
        ++c->launches;
    }
    CU(cudaGetLastError());
    if (c->npes > 1) return adv_exchange_nod(c, ntr, values, c->m.L);
    return ADV_OK;
}

int adv_init_tracers_AB(adv_ctx_t* c, int ntr, int ab_order, double epsilon, const double* const* values,
                        double* const* valuesold, double* const* valuesAB, double* const* del_ttf,
                        double* const* del_ttf_advhoriz, double* const* del_ttf_advvert)
{
    if (!c || !values || !valuesold || !valuesAB || ntr < 1) return fail(ADV_EINVAL, "bad argument");
    if (ab_order != 2 && ab_order != 3)
        return fail(ADV_EINVAL, "Adams-Bashfort tracer order must be 2 or 3, others are not supported (AB_order = " + std::to_string(ab_order) + ")");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->m.L * c->m.Nh;
    const int grid = (int)((n + 255) / 256);
    for (int i = 0; i < ntr; ++i) {
        if (!values[i] || !valuesold[i] || !valuesAB[i]) return fail(ADV_EINVAL, "tracer " + std::to_string(i + 1) + ": null field");
        double* d0 = del_ttf ? del_ttf[i] : nullptr;
        double* d1 = del_ttf_advhoriz ? del_ttf_advhoriz[i] : nullptr;
        double* d2 = del_ttf_advvert ? del_ttf_advvert[i] : nullptr;
        if (ab_order == 2) k_init_tracers_AB<2><<<grid, 256, 0, c->s_comp>>>(n, epsilon, values[i], valuesold[i], valuesAB[i], d0, d1, d2);
        else k_init_tracers_AB<3><<<grid, 256, 0, c->s_comp>>>(n, epsilon, values[i], valuesold[i], valuesAB[i], d0, d1, d2);
        ++c->launches;
    }
    CU(cudaGetLastError());
    return ADV_OK;
}

int adv_ctx_set_gradient_mesh(adv_ctx_t* c, const adv_gradient_mesh_desc_t* g)
{
    if (!c || !g) return fail(ADV_EINVAL, "null argument");
    if (!g->nod_in_elem2D || !g->nod_in_elem2D_num || !g->nlevels || !g->ulevels || !g->edge_up_dn_tri ||
        !g->nlevels_nod2D_min || !g->ulevels_nod2D_max || !g->gradient_sca || !g->elem_area)
        return fail(ADV_EINVAL, "adv_ctx_set_gradient_mesh: null array");
    CU(cudaSetDevice(c->device));
    const MeshDev& m = c->m;
    const int ne = g->n_elem, nn = g->n_nod_in_elem, ld = g->nod_in_elem2D_ld;
    if (ne < m.T || nn < 1 || nn > m.Nh || ld < 1) return fail(ADV_EINVAL, "adv_ctx_set_gradient_mesh: dimensions out of range");
    for (int e = 0; e < m.E; ++e)
        for (int k = 0; k < 2; ++k)
            if (g->edge_up_dn_tri[2 * e + k] < 0 || g->edge_up_dn_tri[2 * e + k] > ne)
                return fail(ADV_EINVAL, "edge_up_dn_tri out of range at edge " + std::to_string(e + 1));
    for (int n = 0; n < nn; ++n) {
        if (g->nod_in_elem2D_num[n] < 0 || g->nod_in_elem2D_num[n] > ld) return fail(ADV_EINVAL, "nod_in_elem2D_num out of range");
        for (int k = 0; k < g->nod_in_elem2D_num[n]; ++k)
            if (g->nod_in_elem2D[(size_t)n * ld + k] < 1 || g->nod_in_elem2D[(size_t)n * ld + k] > ne)
                return fail(ADV_EINVAL, "nod_in_elem2D out of range at node " + std::to_string(n + 1));
    }
    // every end node of a local edge needs its element neighbourhood (src/oce_muscl_adv.F90:397,:420)
    {
        std::vector<int4> em(m.E);
        CU(cudaMemcpy(em.data(), m.edge_meta, sizeof(int4) * m.E, cudaMemcpyDeviceToHost));
        for (int e = 0; e < m.E; ++e)
            if (em[e].x >= nn || em[e].y >= nn)
                return fail(ADV_EINVAL, "nod_in_elem2D does not cover the end nodes of edge " + std::to_string(e + 1));
    }
    auto up_i = [&](DevBuf<int>& b, const int32_t* p, size_t n) { return b.upload(std::vector<int>(p, p + n)); };
    auto up_d = [&](DevBuf<double>& b, const double* p, size_t n) { return b.upload(std::vector<double>(p, p + n)); };
    CU(up_i(c->g_nie, g->nod_in_elem2D, (size_t)nn * ld)); CU(up_i(c->g_nie_num, g->nod_in_elem2D_num, nn));
    CU(up_i(c->g_nlevels, g->nlevels, ne)); CU(up_i(c->g_ulevels, g->ulevels, ne));
    CU(up_i(c->g_tri, g->edge_up_dn_tri, (size_t)2 * m.E));
    CU(up_i(c->g_nmin, g->nlevels_nod2D_min, m.Nh)); CU(up_i(c->g_umax, g->ulevels_nod2D_max, m.Nh));
    CU(up_d(c->g_sca, g->gradient_sca, (size_t)6 * m.T)); CU(up_d(c->g_earea, g->elem_area, ne));
    GradMeshDev& gm = c->gm;
    gm.n_elem = ne; gm.n_nie = nn; gm.ld = ld;
    gm.nie = c->g_nie.p; gm.nie_num = c->g_nie_num.p; gm.nlevels = c->g_nlevels.p; gm.ulevels = c->g_ulevels.p;
    gm.up_dn_tri = c->g_tri.p; gm.nmin = c->g_nmin.p; gm.umax = c->g_umax.p; gm.elem_nodes = c->g_elem_nodes.p;
    gm.gsca = c->g_sca.p; gm.earea = c->g_earea.p;
    {   // per edge: the two triangles and the layers on which fill_up_dn_grad takes their gradients (:431-440)
        std::vector<int4> em(m.E), eg(m.E);
        CU(cudaMemcpy(em.data(), m.edge_meta, sizeof(int4) * m.E, cudaMemcpyDeviceToHost));
        for (int e = 0; e < m.E; ++e) {
            const int t1 = g->edge_up_dn_tri[2 * e] - 1, t2 = g->edge_up_dn_tri[2 * e + 1] - 1;
            int lo = 1, hi = 0;
            if (t1 >= 0 && t2 >= 0) {
                lo = std::max(g->ulevels_nod2D_max[em[e].x], g->ulevels_nod2D_max[em[e].y]);
                hi = std::min(g->nlevels_nod2D_min[em[e].x], g->nlevels_nod2D_min[em[e].y]) - 1;
            }
            eg[e] = make_int4(t1, t2, lo, hi);
        }
        CU(c->edge_g.upload(eg));
        c->m.edge_g = c->edge_g.p;
    }
    // com_elem2D_full: halo of tr_xy (exchange_elem, src/oce_tracer_mod.F90:140)
    c->elem_halo_set = false;
    if (c->npes > 1 && (g->rPEnum > 0 || g->sPEnum > 0)) {
        HaloSet& he = c->halo[HALO_ELEM];
        he.rpeers.clear(); he.speers.clear(); he.send_cols = he.recv_cols = 0; he.recv_base = -1;
        if ((g->rPEnum > 0 && (!g->rPE || !g->rptr || !g->rlist)) || (g->sPEnum > 0 && (!g->sPE || !g->sptr || !g->slist)))
            return fail(ADV_EINVAL, "adv_ctx_set_gradient_mesh: com_elem2D_full has null arrays");
        std::vector<int> sl, rl;
        for (int i = 0; i < g->sPEnum; ++i) {
            Peer p{g->sPE[i], g->sptr[i] - 1, g->sptr[i + 1] - g->sptr[i]};
            he.speers.push_back(p);
            for (int k = 0; k < p.cnt; ++k) {
                const int el = g->slist[p.off + k] - 1;
                if (el < 0 || el >= m.T) return fail(ADV_EINVAL, "com_elem2D_full%slist entry is not an own element");
                sl.push_back(el);
            }
        }
        for (int i = 0; i < g->rPEnum; ++i) {
            Peer p{g->rPE[i], g->rptr[i] - 1, g->rptr[i + 1] - g->rptr[i]};
            he.rpeers.push_back(p);
            for (int k = 0; k < p.cnt; ++k) {
                const int el = g->rlist[p.off + k] - 1;
                if (el < m.T || el >= ne) return fail(ADV_EINVAL, "com_elem2D_full%rlist entry is not a halo element");
                rl.push_back(el);
            }
        }
        he.send_cols = (int)sl.size(); he.recv_cols = (int)rl.size();
        CU(he.slist.upload(sl)); CU(he.rlist.upload(rl));
        c->elem_halo_set = true;
    }
    c->grad_mesh_set = true;
    return ADV_OK;
}

int adv_tracer_gradient_elements(adv_ctx_t* c, int ntr, const double* const* ttf, double* const* tr_xy)
{
    if (!c || !ttf || !tr_xy || ntr < 1) return fail(ADV_EINVAL, "bad argument");
    if (!c->grad_mesh_set) return fail(ADV_ESTATE, "adv_ctx_set_gradient_mesh has not been called");
    CU(cudaSetDevice(c->device));
    const MeshDev& m = c->m;
    const int cpb = cols_per_block(m.L);
    for (int i = 0; i < ntr; ++i) {
        if (!ttf[i] || !tr_xy[i]) return fail(ADV_EINVAL, "tracer " + std::to_string(i + 1) + ": null field");
        if ((uintptr_t)tr_xy[i] & 15u) return fail(ADV_EINVAL, "tr_xy must be 16-byte aligned");
    }
    for (int i = 0; i < ntr;) {                               // two tracers per launch share the element's metadata
        if (i + 1 < ntr) {
            PtrPack<2> pk{{ttf[i], ttf[i + 1]}, {tr_xy[i], tr_xy[i + 1]}};
            k_tracer_gradient_elements<2><<<nblocks(m.T, cpb), cpb * m.L, 0, c->s_comp>>>(m, c->gm, cpb, pk);
            i += 2;
        } else {
            PtrPack<1> pk{{ttf[i]}, {tr_xy[i]}};
            k_tracer_gradient_elements<1><<<nblocks(m.T, cpb), cpb * m.L, 0, c->s_comp>>>(m, c->gm, cpb, pk);
            i += 1;
        }
        ++c->launches;
    }
    CU(cudaGetLastError());
    return ADV_OK;
}

int adv_fill_up_dn_grad(adv_ctx_t* c, int ntr, const double* const* tr_xy, double* const* edge_up_dn_grad)
{
    if (!c || !tr_xy || !edge_up_dn_grad || ntr < 1) return fail(ADV_EINVAL, "bad argument");
    if (!c->grad_mesh_set) return fail(ADV_ESTATE, "adv_ctx_set_gradient_mesh has not been called");
    CU(cudaSetDevice(c->device));
    const MeshDev& m = c->m;
    const int cpb = cols_per_block(m.L);
    for (int i = 0; i < ntr; ++i) {
        if (!tr_xy[i] || !edge_up_dn_grad[i]) return fail(ADV_EINVAL, "tracer " + std::to_string(i + 1) + ": null field");
        if ((uintptr_t)tr_xy[i] & 15u) return fail(ADV_EINVAL, "tr_xy must be 16-byte aligned");
        k_fill_up_dn_grad<<<nblocks(m.E, cpb), cpb * m.L, 0, c->s_comp>>>(m, c->gm, cpb, tr_xy[i], edge_up_dn_grad[i]);
        ++c->launches;
    }
    CU(cudaGetLastError());
    return ADV_OK;
}

int adv_ctx_get_work(adv_ctx_t* c, const char* name, int slot, double* out)
{
    if (!c || !name || !out) return fail(ADV_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_comp));
    CU(cudaStreamSynchronize(c->s_comm));
    const MeshDev& m = c->m;
    const std::string s(name);
    if (s == "edge_volflux") {
        CU(cudaMemcpy(out, c->Q.p, (size_t)m.L * m.E * sizeof(double), cudaMemcpyDeviceToHost));
        return ADV_OK;
    }
    if (slot < 0 || slot >= c->max_tr || c->trloc[slot].chunk < 0 || c->trloc[slot].chunk >= (int)c->cbufs.size())
        return fail(ADV_EINVAL, "slot out of range / tracer not part of the last call");
    const TrLoc loc = c->trloc[slot];
    ChunkBuf& cb = *c->cbufs[loc.chunk];
    // the work arrays interleave the chunk's tracers (and R+ with R-): copy the strided slice
    const double* src = nullptr;
    size_t n = 0, stride = loc.tb, off = loc.pos;
    if (s == "fct_LO") { src = cb.lo.p; n = (size_t)m.L * m.Nh; }
    else if (s == "fct_plus") { src = cb.pm.p; n = (size_t)m.L * m.Nh; stride = 2 * loc.tb; off = 2 * loc.pos; }
    else if (s == "fct_minus") { src = cb.pm.p; n = (size_t)m.L * m.Nh; stride = 2 * loc.tb; off = 2 * loc.pos + 1; }
    else if (s == "adv_flux_hor") { src = cb.adf_h.p; n = (size_t)m.L * m.E; }
    else if (s == "adv_flux_ver") { src = cb.adf_v.p; n = (size_t)m.nl * m.N; }
    else return fail(ADV_EINVAL, "unknown work array " + s);
    CU(cudaMemcpy2D(out, sizeof(double), src + off, stride * sizeof(double), sizeof(double), n, cudaMemcpyDeviceToHost));
    return ADV_OK;
}

int adv_selftest_div(uint64_t count, uint64_t seed, int mode, uint64_t* mismatches)
{
    if (!mismatches) return fail(ADV_EINVAL, "null argument");
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, sizeof(unsigned long long)));
    CU(cudaMemset(d, 0, sizeof(unsigned long long)));
    k_selftest_div<<<148 * 8, 256>>>(count, seed, mode, d);
    unsigned long long h = 0;
    CU(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = h;
    return ADV_OK;
}

int64_t adv_ctx_launch_count(const adv_ctx_t* c) { return c ? c->launches : 0; }
void* adv_ctx_stream(adv_ctx_t* c) { return c ? (void*)c->s_comp : nullptr; }

int adv_ctx_last_elapsed_ms(adv_ctx_t* c, float* ms)
{
    if (!c || !ms || !c->timed) return fail(ADV_ESTATE, "no timed call yet");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->ev_t1));
    CU(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return ADV_OK;
}

int adv_ctx_halo_stats(adv_ctx_t* c, int64_t* bytes_sent, float comm_ms[2], float exposed_ms[2])
{
    if (!c || !bytes_sent || !comm_ms || !exposed_ms) return fail(ADV_EINVAL, "null argument");
    if (!c->x_valid) return fail(ADV_ESTATE, "no multi-rank FCT call yet");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_comp));
    CU(cudaStreamSynchronize(c->s_comm));
    *bytes_sent = c->halo_bytes_last;
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventElapsedTime(&comm_ms[i], c->ev_x0[i], c->ev_x1[i]));
        CU(cudaEventElapsedTime(&exposed_ms[i], c->ev_w0[i], c->ev_w1[i]));
    }
    return ADV_OK;
}

int adv_ctx_set_profiling(adv_ctx_t* c, int on)
{
    if (!c) return fail(ADV_EINVAL, "null ctx");
    c->profiling = on != 0;
    c->ph_valid = false;
    return ADV_OK;
}

int adv_ctx_phase_ms(adv_ctx_t* c, float ms[8])
{
    if (!c || !ms || !c->ph_valid) return fail(ADV_ESTATE, "no profiled call yet");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->ev_ph[4]));
    for (int i = 0; i < 8; ++i) ms[i] = 0.f;
    for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&ms[i], c->ev_ph[i], c->ev_ph[i + 1]));
    return ADV_OK;
}
