// fesom2_b200/csrc/adv_kernels.cuh -- sm_100a kernels of the tracer-advection path.
//
// Hand-written CUDA (FP64, no tensor cores: a bandwidth-bound sparse stencil).  The reference's
// ~22 sweeps per tracer are recast as node-centred, deterministic gathers:
//
//   k_edge_volflux   Q(nz,e): the tracer-independent volume flux of adv_tra_hor_* (once per step)
//   k_fct_lo_adf     D1-D5 + D7/D8 + F5-F7: LO solution, antidiffusive fluxes HO-LO, P+/P- sums
//   k_vert_impl      adv_tra_vert_impl (use_wsplit only)
//   k_fct_bounds     F1-F4 + F8: cluster bounds and the limiter factors R+/R-
//   k_fct_update     F10-F11 + U1-U3: limit and accumulate del_ttf_advhoriz / del_ttf_advvert
//   k_nofct          D7/D8 + U2-U3 when tra_adv_lim /= 'FCT'
//
// Thread mapping: a CTA owns `cpb` node columns, thread = (column, layer).  Fields keep the
// reference layout (level fastest, src/associate_mesh_ass.h:9-79) so the layer index of adjacent
// lanes is contiguous in HBM: every field access is a coalesced run of 8-byte words per column and
// the 4-component gradient is one 32-byte vector per thread.  Edge->node scatters of the reference
// (oce_adv_tra_driver.F90:142-201,:575-633; oce_adv_tra_fct.F90:312-377) become gathers over a
// node->edge CSR sorted by ascending edge id, which reproduces the serial summation order bit for
// bit.  Compile with -fmad=false: parity is checked against a non-contracted
// CPU restatement.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace adv {

enum { HOR_UPW1 = 0, HOR_MUSCL = 1, HOR_MFCT = 2 };
enum { VER_UPW1 = 0, VER_QR4C = 1, VER_PPM = 2, VER_CDIFF = 3 };

constexpr int kBlock = 256;

struct MeshDev {
    int L, nl, N, Nh, T, E;
    // topology (built in adv_ctx_create)
    const int*    ne_ptr;    // (Nh+1) node -> incident edges, ascending edge id
    const int4*   ne_ent;    // {edge, other node, lo | hi<<8 | flags<<16, 0}; flags bit0: node is edges(2,e); bit1: writes adf_h
    const int*    cl_ptr;    // (N+1) FCT cluster of an owned node
    const int2*   cl_ent;    // {node, lo | hi<<8}
    const uchar4* node_lev;  // (Nh) {ulevels_nod2D, nlevels_nod2D, pad_lo, pad_hi}
    const int2*   edge_el;   // (E) {el1, el2} 0-based, -1 none
    const uchar4* edge_lev;  // (E) {nu1, nl1, nu2, nl2}  (nl = nlevels-1; 0,0 for a missing el2)
    const double4* edge_cross; // (E) edge_cross_dxdy
    const double2* edge_c;   // (E) {edge_dxdy(1)*a, edge_dxdy(2)*r_earth}
    const int*    nboundary_lay; // (Nh)
    const double* area;      // (nl,Nh)
    const double* areasvol;  // (nl,Nh)
    // state
    const double *uv, *helem;            // (2,L,T) (L,T)
    const double *w, *we, *wi;           // (nl,Nh)
    const double *hnode, *hnode_new;     // (L,Nh)
    const double *zbar3d, *Z3d;          // (nl,Nh) (L,Nh)
    int use_wsplit;
    double* Q;               // (L,E) volume flux
};

template <int TB>
struct TrBatch {
    const double* ttf[TB];
    const double* ttfAB[TB];
    const double* grad[TB];
    double* lo[TB];
    double* adf_h[TB];
    double* adf_v[TB];
    double* plus[TB];
    double* minus[TB];
    double* dttf_h[TB];
    double* dttf_v[TB];
    double ph[TB], pv[TB];
};

struct NodeRange {
    const int* list;  // optional indirection (0-based node ids); nullptr = identity
    int begin, count;
    int cpb;          // columns per CTA
    int max_slots;    // max over CTAs of the number of (column, incident edge) pairs
};

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

// Correctly rounded x / b from y = RN(1/b): q0 = x*y is refined twice with exact FMA residuals
// (Markstein: a faithful quotient plus one residual step with the correctly rounded reciprocal
// is the correctly rounded quotient; the first step makes q faithful).  Bit-identical to the IEEE
// division of the CPU restatement at a fraction of its ~25 instructions; checked against `/` on
// 2^28 operands per divisor class by adv_selftest_div (tests/test_gpu_division.py).
__device__ __forceinline__ double div_rcp(double x, double b, double y)
{
    double q = x * y;
    double r = fma(-b, q, x);
    q = fma(r, y, q);
    r = fma(-b, q, x);
    return fma(r, y, q);
}
constexpr double kInv6 = 1.0 / 6.0, kInv3 = 1.0 / 3.0;   // RN(1/6), RN(1/3): compile-time IEEE division
__device__ __forceinline__ double div6(double x) { return div_rcp(x, 6.0, kInv6); }
__device__ __forceinline__ double div3(double x) { return div_rcp(x, 3.0, kInv3); }

// ----------------------------------------------------------------------------------------------
// Q(nz,e): vflux of oce_adv_tra_hor.F90:170,190,211-212,226,242 on the level ranges A-E (:127-160)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_edge_volflux(MeshDev m)
{
    const int L = m.L;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m.E * L) return;
    const int e = (int)(idx / L);
    const int nz = (int)(idx - (long long)e * L) + 1;
    const int2 el = m.edge_el[e];
    const uchar4 lv = m.edge_lev[e];
    const int nu1 = lv.x, nl1 = lv.y, nu2 = lv.z, nl2 = lv.w;
    const int nl12 = min(nl1, nl2), nu12 = max(nu1, nu2);
    const bool inA = nz >= nu1 && nz <= nu12 - 1;
    const bool inB = nu2 > 0 && nz >= nu2 && nz <= nu12 - 1;
    const bool inC = nz >= nu12 && nz <= nl12;
    const bool inD = nz >= nl12 + 1 && nz <= nl1;
    const bool inE = nz >= nl12 + 1 && nz <= nl2;
    const bool use1 = inA || inC || inD, use2 = inB || inC || inE;
    const double4 cr = m.edge_cross[e];
    double v1 = 0.0, v2 = 0.0;
    if (use1) {
        const size_t o = (size_t)el.x * L + (nz - 1);
        const double2 uv = reinterpret_cast<const double2*>(m.uv)[o];
        v1 = (-uv.y * cr.x + uv.x * cr.y) * m.helem[o];
    }
    if (use2) {
        const size_t o = (size_t)el.y * L + (nz - 1);
        const double2 uv = reinterpret_cast<const double2*>(m.uv)[o];
        v2 = (uv.y * cr.z - uv.x * cr.w) * m.helem[o];
    }
    double q = 0.0;
    if (use1 && use2) q = v1 + v2;
    else if (use1) q = v1;
    else if (use2) q = v2;
    m.Q[idx] = q;
}

// ----------------------------------------------------------------------------------------------
// vertical interface fluxes.  Each returns the value the reference leaves in flux(k,n) given the
// incoming value `fin` (o_init_zero=.false.: flux := new - flux), applying the reference's
// assignments to interface k in source order, so overlapping special levels of short columns
// behave as in the Fortran.  k is 1-based; col pointers address level 1.
// ----------------------------------------------------------------------------------------------
struct ColV {
    const double* w;     // (nl) vertical velocity column
    const double* area;  // (nl)
    const double* ttf;   // (L)
    const double* Z;     // (L)   Z_3d_n
    const double* zbar;  // (nl)  zbar_3d_n
    const double* hnode; // (L)
    const double* hnode_new; // (L)
    int nzmin, nzmax;    // ulevels_nod2D, nlevels_nod2D
    double dt, num_ord;
};
#define CW(k) c.w[(k)-1]
#define CA(k) c.area[(k)-1]
#define CT(k) c.ttf[(k)-1]
#define CZ(k) c.Z[(k)-1]
#define CZB(k) c.zbar[(k)-1]
#define CHN(k) c.hnode[(k)-1]
#define CHNN(k) c.hnode_new[(k)-1]

// adv_tra_ver_upw1, oce_adv_tra_ver.F90:293-321
__device__ __forceinline__ double ver_upw1(const ColV& c, int k, double fin)
{
    double v = fin;
    if (k == c.nzmin) v = -CW(k) * CT(k) * CA(k) - v;
    if (k == c.nzmax) v = 0.0 - v;
    if (k >= c.nzmin + 1 && k <= c.nzmax - 1) {
        const double w = CW(k);
        v = -0.5 * (CT(k) * (w + fabs(w)) + CT(k - 1) * (w - fabs(w))) * CA(k) - v;
    }
    return v;
}

// adv_tra_ver_qr4c, oce_adv_tra_ver.F90:384-427
__device__ __forceinline__ double ver_qr4c(const ColV& c, int k, double fin)
{
    double v = fin;
    if (k == c.nzmin) v = -CT(k) * CW(k) * CA(k) - v;
    if (k == c.nzmin + 1) v = -0.5 * (CT(k - 1) + CT(k)) * CW(k) * CA(k) - v;
    if (k == c.nzmax - 1) v = -0.5 * (CT(k - 1) + CT(k)) * CW(k) * CA(k) - v;
    if (k == c.nzmax) v = 0.0 - v;
    if (k >= c.nzmin + 2 && k <= c.nzmax - 2) {
        const double t0 = CT(k), tm1 = CT(k - 1), tm2 = CT(k - 2), tp1 = CT(k + 1);
        const double z0 = CZ(k), zm1 = CZ(k - 1), zm2 = CZ(k - 2), zp1 = CZ(k + 1);
        const double qc = (tm1 - t0) / (zm1 - z0);
        const double qu = (t0 - tp1) / (z0 - zp1);
        const double qd = (tm2 - tm1) / (zm2 - zm1);
        const double zb = CZB(k);
        const double Tmean1 = t0 + div3((2 * qc + qu) * (zb - z0));
        const double Tmean2 = tm1 + div3((2 * qc + qd) * (zb - zm1));
        const double w = CW(k);
        const double Tmean = (w + fabs(w)) * Tmean1 + (w - fabs(w)) * Tmean2;
        v = (-0.5 * (1.0 - c.num_ord) * Tmean - c.num_ord * (0.5 * (Tmean1 + Tmean2)) * w) * CA(k) - v;
    }
    return v;
}

// adv_tra_ver_cdiff, oce_adv_tra_ver.F90:670-692 (bottom interface nzmax is never written)
__device__ __forceinline__ double ver_cdiff(const ColV& c, int k, double fin)
{
    double v = fin;
    const int nzmax = c.nzmax - 1;  // the routine's own nzmax = nlevels-1
    if (k >= c.nzmin && k <= nzmax) {
        double tvert;
        if (k == c.nzmin) tvert = -CW(k) * CT(k) * CA(k);
        else { const double tv = 0.5 * (CT(k - 1) + CT(k)); tvert = -tv * CW(k) * CA(k); }
        v = tvert - v;
    }
    return v;
}

__device__ __forceinline__ double dsign1(double b) { return b >= 0.0 ? 1.0 : -1.0; }  // sign(1.0,b)
__device__ __forceinline__ double dmin3(double a, double b, double c) { return dmin(dmin(a, b), c); }

// interface value tv(i) of adv_tra_vert_ppm, oce_adv_tra_ver.F90:496-581; later assignments win
__device__ double ppm_tv(const ColV& c, int i)
{
    if (i == c.nzmax) return CT(c.nzmax - 1);
    if (i == c.nzmax - 1) return 0.5 * (CT(c.nzmax - 2) + CT(c.nzmax - 1));
    if (i == c.nzmin + 1) return 0.5 * (CT(c.nzmin) + CT(c.nzmin + 1));
    if (i == c.nzmin) return CT(c.nzmin);
    const int nz = i - 1;  // loop index of :514, writes tv(nz+1)
    const double dzjm1 = CHNN(nz - 1), dzj = CHNN(nz), dzjp1 = CHNN(nz + 1), dzjp2 = CHNN(nz + 2);
    const double tm1 = CT(nz - 1), t0 = CT(nz), tp1 = CT(nz + 1), tp2 = CT(nz + 2);
    double deltaj = dzj / (dzjm1 + dzj + dzjp1) *
                    ((2.0 * dzjm1 + dzj) / (dzjp1 + dzj) * (tp1 - t0) + (dzj + 2.0 * dzjp1) / (dzjm1 + dzj) * (t0 - tm1));
    double deltajp1 = dzjp1 / (dzj + dzjp1 + dzjp2) *
                      ((2.0 * dzj + dzjp1) / (dzjp2 + dzjp1) * (tp2 - tp1) + (dzjp1 + 2.0 * dzjp2) / (dzj + dzjp1) * (tp1 - t0));
    if ((tp1 - t0) * (t0 - tm1) > 0.0)
        deltaj = dmin3(fabs(deltaj), 2.0 * fabs(tp1 - t0), 2.0 * fabs(t0 - tm1)) * dsign1(deltaj);
    else
        deltaj = 0.0;
    if ((tp2 - tp1) * (tp1 - t0) > 0.0)
        deltajp1 = dmin3(fabs(deltajp1), 2.0 * fabs(tp2 - tp1), 2.0 * fabs(tp1 - t0)) * dsign1(deltajp1);
    else
        deltajp1 = 0.0;
    return t0 + dzj / (dzj + dzjp1) * (tp1 - t0) +
           1.0 / (dzjm1 + dzj + dzjp1 + dzjp2) *
               ((2.0 * dzjp1 * dzj) / (dzj + dzjp1) *
                    ((dzjm1 + dzj) / (2.0 * dzj + dzjp1) - (dzjp2 + dzjp1) / (2.0 * dzjp1 + dzj)) * (tp1 - t0) -
                dzj * (dzjm1 + dzj) / (2.0 * dzj + dzjp1) * deltajp1 +
                dzjp1 * (dzjp1 + dzjp2) / (dzj + 2.0 * dzjp1) * deltaj);
}

// limited parabola edge values of layer j, oce_adv_tra_ver.F90:588-601
__device__ __forceinline__ void ppm_parabola(const ColV& c, int j, double& aL, double& aR)
{
    aL = ppm_tv(c, j);
    aR = ppm_tv(c, j + 1);
    const double t = CT(j);
    if ((aR - t) * (t - aL) <= 0.0) { aL = t; aR = t; }
    if ((aR - aL) * (t - 0.5 * (aL + aR)) > (aR - aL) * (aR - aL) / 6.0) aL = 3.0 * t - 2.0 * aR;
    if ((aR - aL) * (t - 0.5 * (aR + aL)) < -((aR - aL) * (aR - aL)) / 6.0) aR = 3.0 * t - 2.0 * aL;
}

// adv_tra_vert_ppm, oce_adv_tra_ver.F90:487-627: tvert(k) is written by layer k (W(k)>0) or by
// layer k-1 (W(k)<0), never by both; layers with W(j)<=0 and W(j+1)>=0 are skipped (:586).
__device__ double ver_ppm(const ColV& c, int k, double fin)
{
    if (k < c.nzmin || k > c.nzmax) return fin;
    double tvert = 0.0;
    if (k == c.nzmax) tvert = 0.0;
    else if (k == c.nzmin) tvert = -ppm_tv(c, c.nzmin) * CW(k) * CA(k);
    else {
        const double wk = CW(k);
        if (wk > 0.0) {            // layer j = k, its upper interface
            const int j = k;
            double aL, aR;
            ppm_parabola(c, j, aL, aR);
            const double aj = 6.0 * (CT(j) - 0.5 * (aL + aR));
            const double x = dmin(wk * c.dt / CHN(j), 1.0);
            tvert = (-aL - 0.5 * x * (aR - aL + (1.0 - 2.0 / 3.0 * x) * aj));
            tvert = tvert * CA(k) * wk;
        } else if (wk < 0.0) {     // layer j = k-1, its lower interface
            const int j = k - 1;
            double aL, aR;
            ppm_parabola(c, j, aL, aR);
            const double aj = 6.0 * (CT(j) - 0.5 * (aL + aR));
            const double x = dmin(-wk * c.dt / CHN(j), 1.0);
            tvert = (-aR + 0.5 * x * (aR - aL - (1.0 - 2.0 / 3.0 * x) * aj));
            tvert = tvert * CA(k) * wk;
        }
    }
    return tvert - fin;
}

template <int VER>
__device__ __forceinline__ double ver_flux(const ColV& c, int k, double fin)
{
    if (VER == VER_UPW1) return ver_upw1(c, k, fin);
    if (VER == VER_QR4C) return ver_qr4c(c, k, fin);
    if (VER == VER_PPM) return ver_ppm(c, k, fin);
    return ver_cdiff(c, k, fin);
}

// ----------------------------------------------------------------------------------------------
// horizontal edge fluxes for one (edge, layer): LO = adv_tra_hor_upw1 body (:214-216), HO = the
// MUSCL/MFCT body (:446-461,:488-489 / :736-751,:777-778) or upw1 on ttfAB.  (t1,t2)/(a1,a2) are
// the values at edges(1,e), edges(2,e).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double hor_lo(double t1, double t2, double qp, double qm)
{
    return -0.5 * (t1 * qp + t2 * qm);
}

template <int HOR>
__device__ __forceinline__ double hor_ho(double a1, double a2, double q, double qp, double qm, double2 ec,
                                         const double* __restrict__ g, double num_ord, double clo1, double clo2,
                                         double fin)
{
    if (HOR == HOR_UPW1) return -0.5 * (a1 * qp + a2 * qm) - fin;
    const double2 g12 = __ldg(reinterpret_cast<const double2*>(g));      // gx_up, gx_dn
    const double2 g34 = __ldg(reinterpret_cast<const double2*>(g) + 1);  // gy_up, gy_dn
    const double d = 2.0 * (a2 - a1);
    double Tmean2, Tmean1;
    if (HOR == HOR_MUSCL) {
        Tmean2 = a2 - div6(d + ec.x * g12.y + ec.y * g34.y) * clo2;
        Tmean1 = a1 + div6(d + ec.x * g12.x + ec.y * g34.x) * clo1;
    } else {
        Tmean2 = a2 - div6(d + ec.x * g12.y + ec.y * g34.y);
        Tmean1 = a1 + div6(d + ec.x * g12.x + ec.y * g34.x);
    }
    const double cHO = qp * Tmean1 + qm * Tmean2;
    return -0.5 * (1.0 - num_ord) * cHO - q * num_ord * 0.5 * (Tmean1 + Tmean2) - fin;
}

// thread -> (column, layer) decode shared by the node kernels
struct ThreadCol {
    int n, nz0, nzmin, nzmax;
    bool active;
};
__device__ __forceinline__ ThreadCol decode(const MeshDev& m, const NodeRange& r)
{
    ThreadCol t;
    const int L = m.L;
    const int col = threadIdx.x / L;
    t.nz0 = threadIdx.x - col * L;
    const int i = blockIdx.x * r.cpb + col;
    t.active = (col < r.cpb) && (i < r.count);
    t.n = 0; t.nzmin = 1; t.nzmax = 0;
    if (t.active) {
        t.n = r.list ? r.list[r.begin + i] : r.begin + i;
        const uchar4 lv = m.node_lev[t.n];
        t.nzmin = lv.x; t.nzmax = lv.y;
    }
    return t;
}

// ----------------------------------------------------------------------------------------------
// K1: LO solution + antidiffusive fluxes + P+/P- (owned nodes)
//   reference: oce_adv_tra_driver.F90:115-252 (D1-D5), :343-379 (D7-D8),
//              oce_adv_tra_fct.F90:265-377 (b1)
// outputs: lo(nz,n), adf_v(1:nl,n), adf_h(nz,e) (written by the designated end node), raw P+/P-
// sums into plus/minus.
//
// Two phases per CTA so that no thread walks its edges serially behind dependent loads:
//   A  every (incident-edge slot, layer) pair of the CTA's columns is one work item: its thread
//      loads Q, both end values, the 4 gradients and stores the LO and antidiffusive edge flux in
//      shared memory -- all loads of all slots are in flight together;
//   B  thread (column, layer) adds the slots in ascending-edge order (the serial order of the
//      reference's scatter loops), so the sums are bit-identical to the CPU result.
// ----------------------------------------------------------------------------------------------
template <int TB>
__host__ __device__ inline size_t k1_smem_bytes(int L, int cpb, int max_slots)
{
    return ((size_t)2 * TB * cpb * L + (size_t)max_slots * TB * 2 * L) * sizeof(double) +
           (size_t)(3 * cpb + 2 + 2 * max_slots) * sizeof(int);
}

// launch with blockDim.x == cpb * L: thread = (column g, layer nz0), no idle threads
template <int HOR, int VER, int TB>
__global__ void __launch_bounds__(kBlock) k_fct_lo_adf(MeshDev m, TrBatch<TB> b, NodeRange r, double dt)
{
    extern __shared__ double sm[];
    const int L = m.L, nl = m.nl, nthr = blockDim.x;
    double* s_vert = sm;                                   // [2*TB][nthr] LO(we) flux, adf_v at the top interface
    double* s_flux = sm + (size_t)2 * TB * nthr;           // [slot][TB][2][L]
    int2* s_slot = reinterpret_cast<int2*>(s_flux + (size_t)r.max_slots * TB * 2 * L);  // [slot] {node, CSR index}
    int* s_node = reinterpret_cast<int*>(s_slot + r.max_slots);
    int* s_k0 = s_node + r.cpb;                            // first CSR entry of the column
    int* s_off = s_k0 + r.cpb;                             // [cpb+1] slot offset of the column
    const int c0 = blockIdx.x * r.cpb;
    const int ncols = min(r.cpb, r.count - c0);
    const int g = threadIdx.x / L;                         // the only integer division of the kernel
    const int nz0 = threadIdx.x - g * L, nz = nz0 + 1;
    if (threadIdx.x < ncols) {
        const int n = r.list ? r.list[r.begin + c0 + threadIdx.x] : r.begin + c0 + threadIdx.x;
        s_node[threadIdx.x] = n;
        s_k0[threadIdx.x] = m.ne_ptr[n];
        s_off[threadIdx.x + 1] = m.ne_ptr[n + 1] - m.ne_ptr[n];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int c = 0; c < ncols; ++c) { const int cnt = s_off[c + 1]; s_off[c] = off; off += cnt; }
        s_off[ncols] = off;
    }
    __syncthreads();
    const int nslots = s_off[ncols];
    for (int s = threadIdx.x; s < nslots; s += nthr) {
        int c = 0;
        while (c + 1 < ncols && s >= s_off[c + 1]) ++c;
        s_slot[s] = make_int2(s_node[c], s_k0[c] + (s - s_off[c]));
    }
    __syncthreads();

    // ---- phase A: edge fluxes; column group g takes slots g, g+cpb, ... at its layer ----------
    for (int s = g; s < nslots; s += r.cpb) {
        const int2 sl = s_slot[s];
        const int4 ent = __ldg(&m.ne_ent[sl.y]);
        const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
        if (nz < lo || nz > hi) continue;
        const int e = ent.x;
        const bool second = (ent.z >> 16) & 1, writer = (ent.z >> 17) & 1;
        const int i1 = second ? ent.y : sl.x, i2 = second ? sl.x : ent.y;      // edges(1,e), edges(2,e)
        const unsigned oe = (unsigned)e * L + nz0, o1 = (unsigned)i1 * L + nz0, o2 = (unsigned)i2 * L + nz0;
        const double q = __ldg(&m.Q[oe]);
        const double aq = fabs(q), qp = q + aq, qm = q - aq;
        double2 ec = make_double2(0.0, 0.0);
        double clo1 = 1.0, clo2 = 1.0;
        if (HOR != HOR_UPW1) ec = __ldg(&m.edge_c[e]);
        if (HOR == HOR_MUSCL) {
            clo1 = (__ldg(&m.nboundary_lay[i1]) - nz >= 0) ? 1.0 : 0.0;   // oce_adv_tra_hor.F90:411-412
            clo2 = (__ldg(&m.nboundary_lay[i2]) - nz >= 0) ? 1.0 : 0.0;
        }
        double* f = s_flux + (size_t)s * TB * 2 * L + nz0;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double t1 = __ldg(&b.ttf[t][o1]), t2 = __ldg(&b.ttf[t][o2]);
            const double a1 = __ldg(&b.ttfAB[t][o1]), a2 = __ldg(&b.ttfAB[t][o2]);
            const double flo = hor_lo(t1, t2, qp, qm);                                  // driver :115
            const double* gr = (HOR != HOR_UPW1) ? (b.grad[t] + (size_t)oe * 4) : nullptr;
            const double adf = hor_ho<HOR>(a1, a2, q, qp, qm, ec, gr, b.ph[t], clo1, clo2, flo);  // driver :343-354
            f[(2 * t) * L] = flo;
            f[(2 * t + 1) * L] = adf;
            if (writer) b.adf_h[t][oe] = adf;
        }
    }

    // ---- vertical fluxes at the thread's top interface ------------------------------------------
    const bool active = g < ncols;
    int n = 0, nzmin = 1, nzmax = 0;
    if (active) {
        n = s_node[g];
        const uchar4 lv = m.node_lev[n];
        nzmin = lv.x; nzmax = lv.y;
    }
    const size_t oL = (size_t)n * L + nz0;
    const size_t cL = (size_t)n * L, cN = (size_t)n * nl;
    const bool valid = active && nz >= nzmin && nz <= nzmax - 1;
    double flo_top[TB], adfv_top[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { flo_top[t] = 0.0; adfv_top[t] = 0.0; }
    if (active && nz >= nzmin && nz <= nzmax) {
        ColV c;
        c.area = m.area + cN; c.Z = m.Z3d + cL; c.zbar = m.zbar3d + cN;
        c.hnode = m.hnode + cL; c.hnode_new = m.hnode_new + cL;
        c.nzmin = nzmin; c.nzmax = nzmax; c.dt = dt;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            c.ttf = b.ttf[t] + cL; c.w = m.we + cN; c.num_ord = 0.0;
            const double fe = ver_upw1(c, nz, 0.0);                     // driver :235
            double flo = fe;
            if (m.use_wsplit) { c.w = m.w + cN; flo = ver_upw1(c, nz, 0.0); }  // driver :333
            c.ttf = b.ttfAB[t] + cL; c.w = m.w + cN; c.num_ord = b.pv[t];
            flo_top[t] = fe;
            adfv_top[t] = ver_flux<VER>(c, nz, flo);                    // driver :363-379
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        s_vert[(2 * t) * nthr + threadIdx.x] = flo_top[t];
        s_vert[(2 * t + 1) * nthr + threadIdx.x] = adfv_top[t];
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            b.adf_v[t][cN + nz0] = adfv_top[t];
            if (nz0 == L - 1) b.adf_v[t][cN + L] = 0.0;  // interface nl is always the (zero) bottom
        }
    }
    if (!valid) return;

    // ---- phase B: ordered accumulation -----------------------------------------------------------
    double losum[TB], pp[TB], pm[TB];
    const bool has_below = nz0 + 1 < L;
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const double flo_bot = has_below ? s_vert[(2 * t) * nthr + threadIdx.x + 1] : 0.0;
        const double adfv_bot = has_below ? s_vert[(2 * t + 1) * nthr + threadIdx.x + 1] : 0.0;
        flo_top[t] = flo_top[t] - flo_bot;                                             // fv(nz)-fv(nz+1)
        pp[t] = 0.0 + (dmax(0.0, adfv_top[t]) + dmax(0.0, -adfv_bot));                 // fct :291
        pm[t] = 0.0 + (dmin(0.0, adfv_top[t]) + dmin(0.0, -adfv_bot));                 // fct :292
        losum[t] = 0.0;
    }
    const int s1 = s_off[g + 1];
    for (int s = s_off[g]; s < s1; ++s) {
        const int z = __ldg(&m.ne_ent[s_slot[s].y]).z;
        const int lo = z & 0xff, hi = (z >> 8) & 0xff;
        if (nz < lo || nz > hi) continue;
        const bool second = (z >> 16) & 1;
        const double* f = s_flux + (size_t)s * TB * 2 * L + nz0;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double flo = f[(2 * t) * L], adf = f[(2 * t + 1) * L];
            if (!second) {
                losum[t] = losum[t] + flo;                                              // driver :175
                pp[t] = pp[t] + dmax(0.0, adf);                                         // fct :342
                pm[t] = pm[t] + dmin(0.0, adf);                                         // fct :346
            } else {
                losum[t] = losum[t] - flo;                                              // driver :188
                pp[t] = pp[t] + dmax(0.0, -adf);                                        // fct :360
                pm[t] = pm[t] + dmin(0.0, -adf);                                        // fct :364
            }
        }
    }
    const double av = m.areasvol[cN + nz0], hn = m.hnode[oL], hnn = m.hnode_new[oL];
    const double r_av = 1.0 / av, r_hnn = 1.0 / hnn;
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const double num = b.ttf[t][oL] * hn + div_rcp((losum[t] + flo_top[t]) * dt, av, r_av);
        b.lo[t][oL] = div_rcp(num, hnn, r_hnn);                                         // driver :249
        b.plus[t][oL] = pp[t];
        b.minus[t][oL] = pm[t];
    }
}

// ----------------------------------------------------------------------------------------------
// adv_tra_vert_impl (oce_adv_tra_ver.F90:120-236): one thread per owned column, Thomas algorithm.
// cp/tp are kept in the (L,N) scratch arrays `cp`,`tp`.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_vert_impl(MeshDev m, double* __restrict__ ttf, double* __restrict__ cp,
                                                   double* __restrict__ tp, double dt)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= m.N) return;
    const int L = m.L, nl = m.nl;
    const uchar4 lv = m.node_lev[n];
    const int nzmin = lv.x, nzmax = lv.y;
    const double* W = m.wi + (size_t)n * nl;
    const double* area = m.area + (size_t)n * nl;
    const double* avol = m.areasvol + (size_t)n * nl;
    const double* hnn = m.hnode_new + (size_t)n * L;
    double* T = ttf + (size_t)n * L;
    double* CP = cp + (size_t)n * L;
    double* TP = tp + (size_t)n * L;
    const double zinv = 1.0 * dt;
#define AW(k) W[(k)-1]
#define AA(k) area[(k)-1]
#define AV(k) avol[(k)-1]
#define AH(k) hnn[(k)-1]
#define AT(k) T[(k)-1]
    double cp_prev = 0.0, tp_prev = 0.0;
    for (int nz = nzmin; nz <= nzmax - 1; ++nz) {
        double a, bb, c, tr, v_adv;
        if (nz == nzmin) {                                      // :154-170, :198-200
            a = 0.0;
            v_adv = zinv * AA(nz) / AV(nz);
            bb = AH(nz) + AW(nz) * v_adv;
            v_adv = zinv * AA(nz + 1) / AV(nz);
            bb = bb - dmin(0.0, AW(nz + 1)) * v_adv;
            c = -dmax(0.0, AW(nz + 1)) * v_adv;
            tr = -(bb - AH(nz)) * AT(nz) - c * AT(nz + 1);
        } else if (nz <= nzmax - 2) {                           // :174-183, :202-205
            v_adv = zinv * AA(nz) / AV(nz);
            a = dmin(0.0, AW(nz)) * v_adv;
            bb = AH(nz) + dmax(0.0, AW(nz)) * v_adv;
            v_adv = zinv * AA(nz + 1) / AV(nz);
            bb = bb - dmin(0.0, AW(nz + 1)) * v_adv;
            c = -dmax(0.0, AW(nz + 1)) * v_adv;
            tr = -a * AT(nz - 1) - (bb - AH(nz)) * AT(nz) - c * AT(nz + 1);
        } else {                                                // :187-195, :206-208
            v_adv = zinv * AA(nz) / AV(nz);
            a = dmin(0.0, AW(nz)) * v_adv;
            bb = AH(nz) + dmax(0.0, AW(nz)) * v_adv;
            c = 0.0;
            tr = -a * AT(nz - 1) - (bb - AH(nz)) * AT(nz);
        }
        if (nz == nzmin) { cp_prev = c / bb; tp_prev = tr / bb; }                       // :211-213
        else { const double mm = bb - cp_prev * a; cp_prev = c / mm; tp_prev = (tr - tp_prev * a) / mm; }
        CP[nz - 1] = cp_prev; TP[nz - 1] = tp_prev;
    }
    double trn = TP[nzmax - 2];                                                         // :224
    AT(nzmax - 1) = AT(nzmax - 1) + trn;
    for (int nz = nzmax - 2; nz >= nzmin; --nz) {                                       // :227-235
        trn = TP[nz - 1] - CP[nz - 1] * trn;
        AT(nz) = AT(nz) + trn;
    }
#undef AW
#undef AA
#undef AV
#undef AH
#undef AT
}

// ----------------------------------------------------------------------------------------------
// K2: FCT bounds and limiter factors (owned nodes); needs lo on the halo.
//   reference: oce_adv_tra_fct.F90:124-248 (a1-a3), :394-405 (b2)
// plus/minus hold the raw P+/P- sums on entry and R+/R- on exit.
// ----------------------------------------------------------------------------------------------
template <int TB>
__global__ void __launch_bounds__(kBlock) k_fct_bounds(MeshDev m, TrBatch<TB> b, NodeRange r, double dt)
{
    extern __shared__ double sm[];  // [2*TB][blockDim]: tvert_max, tvert_min
    const int L = m.L, nl = m.nl;
    const ThreadCol tc = decode(m, r);
    const int n = tc.n, nz0 = tc.nz0, nz = nz0 + 1;
    const bool valid = tc.active && nz >= tc.nzmin && nz <= tc.nzmax - 1;
    double tmax[TB], tmin[TB];
    if (valid) {
        const uchar4 lv = m.node_lev[n];
        const bool padded = nz < lv.z || nz > lv.w;   // some element of the cluster is dry at nz
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            tmax[t] = padded ? -1.0e3 : -CUDART_INF;  // bignumber, oce_adv_tra_fct.F90:100,159-176
            tmin[t] = padded ? 1.0e3 : CUDART_INF;
        }
        const int k1 = m.cl_ptr[n + 1];
        // branch-free body (an out-of-range entry is loaded and discarded) so that the unrolled
        // iterations issue their loads together
#pragma unroll 4
        for (int k = m.cl_ptr[n]; k < k1; ++k) {
            const int2 ent = __ldg(&m.cl_ent[k]);
            const int lo = ent.y & 0xff, hi = (ent.y >> 8) & 0xff;
            const bool inr = nz >= lo && nz <= hi;
            const size_t o = (size_t)ent.x * L + nz0;
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const double a = __ldg(&b.lo[t][o]), c = __ldg(&b.ttf[t][o]);
                const double hi2 = dmax(a, c), lo2 = dmin(a, c);
                tmax[t] = (inr && hi2 > tmax[t]) ? hi2 : tmax[t];   // a1 :129, a2 :166, a3 :209
                tmin[t] = (inr && lo2 < tmin[t]) ? lo2 : tmin[t];
            }
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            sm[(2 * t) * blockDim.x + threadIdx.x] = tmax[t];
            sm[(2 * t + 1) * blockDim.x + threadIdx.x] = tmin[t];
        }
    }
    __syncthreads();
    if (!valid) return;
    const size_t oL = (size_t)n * L + nz0;
    const double av = m.areasvol[(size_t)n * nl + nz0], hnn = m.hnode_new[oL];
    const double r_av = 1.0 / av, r_hnn = 1.0 / hnn;
    const bool edge_layer = (nz == tc.nzmin) || (nz == tc.nzmax - 1);   // :233-234, :245-247
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        double vmax = tmax[t], vmin = tmin[t];
        if (!edge_layer) {                                               // :238-241
            const double* smax = sm + (2 * t) * blockDim.x + threadIdx.x;
            const double* smin = sm + (2 * t + 1) * blockDim.x + threadIdx.x;
            vmax = dmax(dmax(smax[-1], vmax), smax[1]);
            vmin = dmin(dmin(smin[-1], vmin), smin[1]);
        }
        const double lo = b.lo[t][oL];
        const double inc_max = vmax - lo, inc_min = vmin - lo;
        double flux = div_rcp(div_rcp(b.plus[t][oL] * dt, av, r_av), hnn, r_hnn) + 1e-16;   // b2 :399
        b.plus[t][oL] = dmin(1.0, inc_max / flux);
        flux = div_rcp(div_rcp(b.minus[t][oL] * dt, av, r_av), hnn, r_hnn) - 1e-16;         // :401
        b.minus[t][oL] = dmin(1.0, inc_min / flux);
    }
}

// ----------------------------------------------------------------------------------------------
// K3: limit the antidiffusive fluxes and accumulate the tendencies.  Owned nodes: vertical +
// horizontal; halo nodes: the partial horizontal sums the reference's edge scatter leaves there.
//   reference: oce_adv_tra_fct.F90:425-500 (b3), oce_adv_tra_driver.F90:529-633 (U1-U3)
// ----------------------------------------------------------------------------------------------
template <int TB>
__global__ void __launch_bounds__(kBlock) k_fct_update(MeshDev m, TrBatch<TB> b, NodeRange r, double dt)
{
    extern __shared__ double sm[];  // [TB][blockDim]: limited vertical flux at the top interface
    const int L = m.L, nl = m.nl;
    const ThreadCol tc = decode(m, r);
    const int n = tc.n, nz0 = tc.nz0, nz = nz0 + 1;
    const bool owned = n < m.N;
    const bool valid = tc.active && nz >= tc.nzmin && nz <= tc.nzmax - 1;
    const size_t oL = (size_t)n * L + nz0, cN = (size_t)n * nl;
    double fv_top[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) fv_top[t] = 0.0;
    if (tc.active && owned && nz >= tc.nzmin && nz <= tc.nzmax) {
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double f = b.adf_v[t][cN + nz0];
            double ae = 1.0;
            if (nz == tc.nzmin) {                                         // fct :430-438
                ae = (f >= 0.0) ? dmin(ae, b.plus[t][oL]) : dmin(ae, b.minus[t][oL]);
            } else if (nz <= tc.nzmax - 1) {                              // :442-453
                if (f >= 0.0) { ae = dmin(ae, b.minus[t][oL - 1]); ae = dmin(ae, b.plus[t][oL]); }
                else { ae = dmin(ae, b.plus[t][oL - 1]); ae = dmin(ae, b.minus[t][oL]); }
            }                                                             // bottom interface untouched
            fv_top[t] = ae * f;
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) sm[t * blockDim.x + threadIdx.x] = fv_top[t];
    __syncthreads();
    if (!valid) return;
    const double av = m.areasvol[cN + nz0];
    const double r_av = 1.0 / av;
    const bool has_below = nz0 + 1 < L;
    double dh[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) dh[t] = b.dttf_h[t][oL];
    if (owned) {
        const double hn = m.hnode[oL], hnn = m.hnode_new[oL];
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double fv_bot = has_below ? sm[t * blockDim.x + threadIdx.x + 1] : 0.0;
            double dv = b.dttf_v[t][oL];
            dv = dv - b.ttf[t][oL] * hn + b.lo[t][oL] * hnn;              // driver :535
            dv = dv + div_rcp((fv_top[t] - fv_bot) * dt, av, r_av);       // driver :556
            b.dttf_v[t][oL] = dv;
        }
    }
    const int k1 = m.ne_ptr[n + 1];
    double pn[TB], mn[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { pn[t] = __ldg(&b.plus[t][oL]); mn[t] = __ldg(&b.minus[t][oL]); }
    // branch-free body: all four factors are loaded whatever the sign of the flux, out-of-range
    // entries are loaded and discarded, so the unrolled iterations issue their loads together
#pragma unroll 3
    for (int k = m.ne_ptr[n]; k < k1; ++k) {
        const int4 ent = __ldg(&m.ne_ent[k]);
        const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
        const bool inr = nz >= lo && nz <= hi;
        const bool second = (ent.z >> 16) & 1;
        const size_t oe = (size_t)ent.x * L + nz0, om = (size_t)ent.y * L + nz0;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double f = __ldg(&b.adf_h[t][oe]);
            const double pmo = __ldg(&b.plus[t][om]), mmo = __ldg(&b.minus[t][om]);
            const double p1 = second ? pmo : pn[t], m1 = second ? mmo : mn[t];   // factors at edges(1,e)
            const double p2 = second ? pn[t] : pmo, m2 = second ? mn[t] : mmo;   // factors at edges(2,e)
            double ae = 1.0;
            if (f >= 0.0) { ae = dmin(ae, p1); ae = dmin(ae, m2); }       // fct :489-491
            else { ae = dmin(ae, m1); ae = dmin(ae, p2); }                // :493-494
            const double term = div_rcp(ae * f * dt, av, r_av);           // fct :497, driver :607,:620
            const double nd = second ? dh[t] - term : dh[t] + term;
            dh[t] = inr ? nd : dh[t];
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) b.dttf_h[t][oL] = dh[t];
}

// ----------------------------------------------------------------------------------------------
// tra_adv_lim /= 'FCT': HO fluxes with o_init_zero=.true. and the plain tendency update
//   reference: oce_adv_tra_driver.F90:339-379, :387, :551-633 (vertical velocity is `we`, :358)
// ----------------------------------------------------------------------------------------------
template <int HOR, int VER, int TB>
__global__ void __launch_bounds__(kBlock) k_nofct(MeshDev m, TrBatch<TB> b, NodeRange r, double dt)
{
    extern __shared__ double sm[];  // [TB][blockDim]
    const int L = m.L, nl = m.nl;
    const ThreadCol tc = decode(m, r);
    const int n = tc.n, nz0 = tc.nz0, nz = nz0 + 1;
    const bool owned = n < m.N;
    const bool valid = tc.active && nz >= tc.nzmin && nz <= tc.nzmax - 1;
    const size_t oL = (size_t)n * L + nz0, cL = (size_t)n * L, cN = (size_t)n * nl;
    double fv_top[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) fv_top[t] = 0.0;
    if (tc.active && owned && nz >= tc.nzmin && nz <= tc.nzmax) {
        ColV c;
        c.area = m.area + cN; c.Z = m.Z3d + cL; c.zbar = m.zbar3d + cN;
        c.hnode = m.hnode + cL; c.hnode_new = m.hnode_new + cL;
        c.nzmin = tc.nzmin; c.nzmax = tc.nzmax; c.dt = dt; c.w = m.we + cN;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            c.ttf = b.ttfAB[t] + cL; c.num_ord = b.pv[t];
            fv_top[t] = ver_flux<VER>(c, nz, 0.0);
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) sm[t * blockDim.x + threadIdx.x] = fv_top[t];
    __syncthreads();
    if (tc.active && owned) {
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            b.adf_v[t][cN + nz0] = fv_top[t];
            if (nz0 == L - 1) b.adf_v[t][cN + L] = 0.0;
        }
    }
    if (!valid) return;
    const double av = m.areasvol[cN + nz0];
    const double r_av = 1.0 / av;
    const bool has_below = nz0 + 1 < L;
    double dh[TB], tabn[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { dh[t] = b.dttf_h[t][oL]; tabn[t] = b.ttfAB[t][oL]; }
    if (owned) {
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double fv_bot = has_below ? sm[t * blockDim.x + threadIdx.x + 1] : 0.0;
            b.dttf_v[t][oL] = b.dttf_v[t][oL] + div_rcp((fv_top[t] - fv_bot) * dt, av, r_av);   // driver :556
        }
    }
    const int nb_n = (HOR == HOR_MUSCL) ? m.nboundary_lay[n] : 0;
    const int k1 = m.ne_ptr[n + 1];
    for (int k = m.ne_ptr[n]; k < k1; ++k) {
        const int4 ent = m.ne_ent[k];
        const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
        if (nz < lo || nz > hi) continue;
        const int e = ent.x, mo = ent.y;
        const bool second = (ent.z >> 16) & 1, writer = (ent.z >> 17) & 1;
        const size_t oe = (size_t)e * L + nz0, om = (size_t)mo * L + nz0;
        const double q = m.Q[oe];
        const double aq = fabs(q), qp = q + aq, qm = q - aq;
        double2 ec = make_double2(0.0, 0.0);
        double clo_n = 1.0, clo_m = 1.0;
        if (HOR != HOR_UPW1) ec = m.edge_c[e];
        if (HOR == HOR_MUSCL) {
            clo_n = (nb_n - nz >= 0) ? 1.0 : 0.0;
            clo_m = (m.nboundary_lay[mo] - nz >= 0) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double tabm = b.ttfAB[t][om];
            const double a1 = second ? tabm : tabn[t], a2 = second ? tabn[t] : tabm;
            const double* g = (HOR != HOR_UPW1) ? (b.grad[t] + oe * 4) : nullptr;
            const double f = hor_ho<HOR>(a1, a2, q, qp, qm, ec, g, b.ph[t], second ? clo_m : clo_n,
                                         second ? clo_n : clo_m, 0.0);
            const double term = div_rcp(f * dt, av, r_av);                          // driver :607,:620
            dh[t] = second ? dh[t] - term : dh[t] + term;
            if (writer && owned) b.adf_h[t][oe] = f;
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) b.dttf_h[t][oL] = dh[t];
}

// ----------------------------------------------------------------------------------------------
// halo pack (replaces the MPI_TYPE_INDEXED send types, gen_modules_partitioning.F90:462-473):
// out[(i*nlev)+nz] = field[(slist[i])*nlev + nz]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_pack_halo(const double* __restrict__ field, const int* __restrict__ slist,
                                                      int count, int nlev, double* __restrict__ out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)count * nlev) return;
    const int i = (int)(idx / nlev), nz0 = (int)(idx - (long long)i * nlev);
    out[idx] = field[(size_t)slist[i] * nlev + nz0];
}

// dwarf epilogue (fesom.F90:105-125 with del_ttf reset per step): values += (dh+dv)/hnode_new
__global__ void __launch_bounds__(kBlock) k_update_values(MeshDev m, double* __restrict__ values,
                                                          const double* __restrict__ dh, const double* __restrict__ dv)
{
    const int L = m.L;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m.N * L) return;
    const int n = (int)(idx / L), nz = (int)(idx - (long long)n * L) + 1;
    const uchar4 lv = m.node_lev[n];
    if (nz < lv.x || nz > lv.y - 1) return;
    const double del = 0.0 + dh[idx] + dv[idx];
    values[idx] = values[idx] + del / m.hnode_new[idx];
}

// self-test of div_rcp against the IEEE division: returns the number of mismatching results over
// `count` pseudo-random operand pairs (mode 0: b = 6, 1: b = 3, 2: random b in [1e-3, 1e13])
__global__ void k_selftest_div(unsigned long long count, unsigned long long seed, int mode, unsigned long long* bad)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long nbad = 0;
    for (; i < count; i += stride) {
        unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        unsigned long long z2 = (z + 0x632BE59BD9B4E019ull) * 0xD6E8FEB86659FD93ull; z2 ^= z2 >> 32;
        // x: random significand, exponent in [-40, 40], random sign
        const int ex = (int)((z >> 52) % 81) - 40;
        double x = ldexp(1.0 + (double)(z & 0xFFFFFFFFFFFFFull) * 0x1p-52, ex);
        if (z2 & 1) x = -x;
        double b = 6.0;
        if (mode == 1) b = 3.0;
        if (mode == 2) b = ldexp(1.0 + (double)((z2 >> 1) & 0xFFFFFFFFFFFFFull) * 0x1p-52, (int)((z2 >> 54) % 54) - 10);
        const double y = (mode == 0) ? kInv6 : (mode == 1) ? kInv3 : 1.0 / b;
        if (div_rcp(x, b, y) != x / b) ++nbad;
    }
    if (nbad) atomicAdd(bad, nbad);
}

}  // namespace adv
