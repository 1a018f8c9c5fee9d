// fesom2_b200/csrc/adv_kernels.cuh -- sm_100a kernels of the tracer-advection path.
//
// Hand-written CUDA (FP64, no tensor cores: a bandwidth-bound sparse stencil).  The reference's
// ~22 sweeps per tracer become four passes per tracer chunk, separated only
// by the two dependency fronts that carry a halo exchange:
//
//   k_edge_flux    D1 + D7 (+ the tracer-independent volume flux Q): one thread per (edge, layer)
//                  computes the low-order and high-order edge flux ONCE and stores the antidiffusive
//                  flux HO-LO; a pure streaming kernel, every operand read exactly once.  The shipped
//                  path is its bulk-copy form k_edge_flux_b (adv_pipe.cuh); this register-gather form
//                  serves UPW1 and ADV_BULK=0
//   k_node_lo      D2-D5 + D8: LO solution by an ordered gather of the upwind edge fluxes
//                  (recomputed from Q: 3 flops instead of a stored field), vertical LO/HO fluxes
//   ------------------------------------------------------------ exchange_nod(fct_LO)
//   k_fct_bounds   F1-F8: cluster bounds, P+/P- sums and the limiter factors R+/R-
//   ------------------------------------------------------------ exchange_nod(fct_plus, fct_minus)
//   k_fct_update   F10-F11 + U1-U3: limit and accumulate del_ttf_advhoriz / del_ttf_advvert
//   k_nofct_update D8 + U2-U3 when tra_adv_lim /= 'FCT' (D7 by the edge kernel with Chunk::nolo)
//   k_vert_impl    adv_tra_vert_impl (use_wsplit only)
//   k_tracer_gradient_elements, k_fill_up_dn_grad   the producer of edge_up_dn_grad
//   k_init_tracers_AB, k_update_values              prologue / epilogue of the dwarf iteration (row 2)
//
// Thread mapping: a CTA owns `cpb` whole columns (nodes or edges), thread = (column, layer).  All
// caller-visible fields keep the reference layout (level fastest, src/associate_mesh_ass.h:9-79),
// so adjacent lanes touch adjacent 8-byte words.  The library-owned work arrays (fct_LO, adv_flux_*,
// fct_plus/minus) interleave the TB tracers of a chunk -- and R+ with R- -- at every (level, column)
// so that one 16/32-byte vector access serves the whole chunk.
//
// Edge->node scatters of the reference (oce_adv_tra_driver.F90:142-201,:575-633;
// oce_adv_tra_fct.F90:312-377) are gathers over node->edge ELL rows sorted by ascending edge id, which
// reproduces the serial summation order bit for bit.  Gathers are written as
// "load a batch of G slots into registers, then accumulate in order" so that G x fields loads are
// in flight per thread; the records the NEXT CTAs will wait for first (node_rec, ELL rows, own columns)
// are pulled into L2 a few hundred CTAs ahead (node_thread, prefetch_own_columns).  Compile with
// -fmad=false: parity is checked against a non-contracted CPU restatement.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace adv {

enum { HOR_UPW1 = 0, HOR_MUSCL = 1, HOR_MFCT = 2 };
enum { VER_UPW1 = 0, VER_QR4C = 1, VER_PPM = 2, VER_CDIFF = 3 };

constexpr int kBlock = 256;
#ifndef ADV_QR4C_RCP
#define ADV_QR4C_RCP 1   // k_node_lo: one reciprocal per thread shared through smem instead of six IEEE divisions
#endif
#ifndef ADV_PF_OWN
#define ADV_PF_OWN 0     // bit mask N1|K2|K3: also pull the next CTAs' own columns into L2.  Round 1 / early round 2: 5 (K3 1.37 -> 1.26 ms,
                         // N1 -1 %); with the per-kernel CTA sizes of DESIGN decision 17 it costs N1 6 % and K3 1.5 % (profiles/r8p_*): off
#endif
// minimum resident CTAs per SM (register caps): measured on B200, see DESIGN.md section 4 --
// occupancy beats register-resident batching: 4-5 CTAs of 7 warps with a few spilled words run
// 20-30 % faster than 2 CTAs without spills (gpurun_out/exp_mb*.log, profiles/r1_tuning.md)
// Register budgets of the node kernels.  The CTAs are cpb * L threads (210 for L = 70, i.e. 7 warps): a
// cap of 72 registers still fits 4 CTAs per SM there (65536 / (224 * 72)), 56 fits 5, and the 8 extra
// registers over the __launch_bounds__(256, n) caps remove most of the spills (profiles/r1_tuning.md).
#ifndef ADV_N1_REGS
#define ADV_N1_REGS 56
#endif
#ifndef ADV_K2_REGS
#define ADV_K2_REGS 64   // with 256-thread CTAs and one slot per gather batch: 4 x 8 = 32 resident warps (72 registers: 28), 8 bytes spilled
#endif
#ifndef ADV_K3_REGS
#define ADV_K3_REGS 56   // since the vertical part of the update moved into k_fct_bounds: no spills at 56, 5 CTAs per SM (0.84 -> 0.77 ms)
#endif
#if ADV_N1_REGS > 0
#define ADV_N1_BOUNDS __maxnreg__(ADV_N1_REGS)
#else
#define ADV_N1_BOUNDS __launch_bounds__(kBlock, ADV_N1_MINB)
#endif
#if ADV_K2_REGS > 0
#define ADV_K2_BOUNDS __maxnreg__(ADV_K2_REGS)
#else
#define ADV_K2_BOUNDS __launch_bounds__(kBlock, ADV_K2_MINB)
#endif
#if ADV_K3_REGS > 0
#define ADV_K3_BOUNDS __maxnreg__(ADV_K3_REGS)
#else
#define ADV_K3_BOUNDS __launch_bounds__(kBlock, ADV_K3_MINB)
#endif
#ifndef ADV_E1_MINB
#define ADV_E1_MINB 4
#endif
#ifndef ADV_N1_MINB
#define ADV_N1_MINB 5
#endif
#ifndef ADV_K2_MINB
#define ADV_K2_MINB 4
#endif
#ifndef ADV_K3_MINB
#define ADV_K3_MINB 4
#endif

struct MeshDev {
    int L, nl, N, Nh, T, E;
    unsigned div_magic;      // ceil(2^20 / L): (tid * div_magic) >> 20 == tid / L for tid < 1024
    // topology (built in adv_ctx_create)
    const int*    ne_ptr;    // (Nh+1) node -> incident edges, ascending edge id
    const int4*   ne_ent;    // {edge, other node, lo | hi<<8 | flags<<16, 0}; flags bit0: node is edges(2,e)
    const int4*   ne_ell;    // (Nh, ell_w) the same entries padded to the maximum degree (ELL): no pointer chase
    int ell_w;
    const uchar4* node_lev;  // (Nh) {ulevels_nod2D, nlevels_nod2D, pad_lo, pad_hi}
    const uint2*  node_rec;  // (Nh) bytes {ulev, nlev, pad_lo, pad_hi, self_lo, self_hi, degree, 0}: pad = levels where
                             //      every element of the FCT cluster is wet; self = range covered by the node's own elements
    const int4*   edge_meta; // (E) {edges(1,e), edges(2,e), el1, el2} 0-based, el2 = -1: none
    const int2*   edge_el;   // (E) {el1, el2} 0-based, -1 none
    const uchar4* edge_lev;  // (E) {nu1, nl1, nu2, nl2}  (nl = nlevels-1; 0,0 for a missing el2)
    const double4* edge_cross; // (E) edge_cross_dxdy
    const double2* edge_c;   // (E) {edge_dxdy(1)*a, edge_dxdy(2)*r_earth}
    const int4*   edge_g;    // (E) fused gradients (adv_ctx_set_gradient_mesh): {up triangle, down triangle (0-based, -1 none),
                             //     first, last layer on which the triangles' gradients are used (first > last: none)}
    const int*    nboundary_lay; // (Nh)
    const double* area;      // (nl,Nh)
    const double* areasvol;  // (nl,Nh)
    const double* r_areasvol;  // (nl,Nh) RN(1/areasvol): the exact-division helper y of div_rcp, precomputed (static geometry)
    // state
    const double *uv, *helem;            // (2,L,T) (L,T)
    const double *w, *we, *wi;           // (nl,Nh)
    const double *hnode, *hnode_new;     // (L,Nh)
    const double *zbar3d, *Z3d;          // (nl,Nh) (L,Nh)
    int use_wsplit;
    double* Q;               // (L,E) volume flux
};

// One chunk of TB tracers.  Work arrays are tracer-interleaved:
//   lo[(n*L+nz0)*TB+t], adf_h[(e*L+nz0)*TB+t], adf_v[(n*nl+k0)*TB+t], pm[((n*L+nz0)*TB+t)*2+{0:plus,1:minus}]
template <int TB>
struct Chunk {
    const double* ttf[TB];
    const double* ttfAB[TB];
    const double* grad[TB];
    const double* txy[TB];    // fused gradients (grad[t] == nullptr): tr_xy (2,L,n_elem) and the node-mean gradients (2,L,Nh)
    const double* gmean[TB];
    double* dttf_h[TB];
    double* dttf_v[TB];
    double* dgh[TB];          // tra_advhoriz / tra_advvert (nl-1, Nh) of tracers with ltra_diag (oce_adv_tra_driver.F90:221-229,
    double* dgv[TB];          //   :307-318, :464-488), owned nodes; nullptr = off
    double ph[TB], pv[TB];
    double *lo, *adf_h, *adf_v, *pm;
    int nolo;                 // 1: tra_adv_lim /= 'FCT' -- the edge kernel stores the high-order flux itself (o_init_zero = .true.,
                              //    oce_adv_tra_driver.F90:339-346), not HO - LO
};

struct NodeRange {
    const int* list;  // optional indirection (0-based node ids); nullptr = identity
    int begin, count;
    int cpb;          // columns per CTA
    int pf;           // metadata prefetch distance in CTAs (0 = off)
    int skip_s;       // 1: columns flagged "boundary set" in node_rec (bit 24 of .y) are left to another launch
};

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

// x or -x (mask = 0 / 0x80000000): the IEEE negation is a flip of the sign bit, one integer instruction on the high
// word instead of DADD + two selects
__device__ __forceinline__ double flip_sign(double x, int mask)
{
    return __hiloint2double(__double2hiint(x) ^ mask, __double2loint(x));
}
// L2 software prefetch of [p, p+bytes), widened to the 16-byte granularity of the bulk-prefetch
// instruction (sm_90+: one instruction per column instead of one per 128-byte line).  CTAs
// prefetch the operands of the CTA `pf` launches ahead so that the dependent metadata -> data
// load chain of a gather finds its lines in the 126 MB L2 instead of paying HBM latency twice.
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes)
{
    const unsigned long long a = (unsigned long long)p, a0 = a & ~15ull;
    const unsigned sz = (unsigned)(((a + bytes + 15ull) & ~15ull) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(sz) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// vector access to the tracer-interleaved work arrays: TB doubles at p (16-byte aligned for TB=2)
template <int TB> __device__ __forceinline__ void ldv(const double* __restrict__ p, double (&v)[TB]);
template <> __device__ __forceinline__ void ldv<1>(const double* __restrict__ p, double (&v)[1]) { v[0] = __ldg(p); }
template <> __device__ __forceinline__ void ldv<2>(const double* __restrict__ p, double (&v)[2])
{
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    v[0] = t.x; v[1] = t.y;
}
template <int TB> __device__ __forceinline__ void stv(double* __restrict__ p, const double (&v)[TB]);
template <> __device__ __forceinline__ void stv<1>(double* __restrict__ p, const double (&v)[1]) { p[0] = v[0]; }
template <> __device__ __forceinline__ void stv<2>(double* __restrict__ p, const double (&v)[2])
{
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}
// {plus, minus} pairs of the TB tracers at p
template <int TB> __device__ __forceinline__ void ldpm(const double* __restrict__ p, double (&pl)[TB], double (&mi)[TB])
{
#ifndef ADV_NO_LDG256
    if constexpr (TB == 2) {   // the {R+,R-} pairs of both tracers are one 32-byte record: a single 256-bit load (sm_100: LDG.E.256)
        double a, b, c, d;
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
        pl[0] = a; mi[0] = b; pl[TB - 1] = c; mi[TB - 1] = d;
    } else
#endif
    {
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(p) + t);
            pl[t] = v.x; mi[t] = v.y;
        }
    }
}
// store TB {a, b} pairs at p (32-byte aligned for TB = 2: one 256-bit store)
template <int TB> __device__ __forceinline__ void stpm(double* __restrict__ p, const double (&a)[TB], const double (&b)[TB])
{
#ifndef ADV_NO_LDG256
    if constexpr (TB == 2) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a[0]), "d"(b[0]), "d"(a[TB - 1]), "d"(b[TB - 1]) : "memory");
    } else
#endif
    {
#pragma unroll
        for (int t = 0; t < TB; ++t) reinterpret_cast<double2*>(p)[t] = make_double2(a[t], b[t]);
    }
}

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

// thread -> (column, layer) of a CTA that owns cpb whole columns: blockDim.x == cpb * L
struct ColThread { int g, nz0; };
__device__ __forceinline__ ColThread col_thread(const MeshDev& m)
{
    ColThread c;
    c.g = (int)((threadIdx.x * m.div_magic) >> 20);
    c.nz0 = (int)threadIdx.x - c.g * m.L;
    return c;
}

// the level ranges A-E of an edge column (oce_adv_tra_hor.F90:127-160) -> which element(s)
// contribute to the volume flux at level nz
__device__ __forceinline__ void edge_use(uchar4 lv, int nz, bool& use1, bool& use2)
{
    const int nu1 = lv.x, nl1 = lv.y, nu2 = lv.z, nl2 = lv.w;
    const int nl12 = min(nl1, nl2), nu12 = max(nu1, nu2);
    const bool inA = nz >= nu1 && nz <= nu12 - 1;
    const bool inB = nu2 > 0 && nz >= nu2 && nz <= nu12 - 1;
    const bool inC = nz >= nu12 && nz <= nl12;
    const bool inD = nz >= nl12 + 1 && nz <= nl1;
    const bool inE = nz >= nl12 + 1 && nz <= nl2;
    use1 = inA || inC || inD;
    use2 = inB || inC || inE;
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

// adv_tra_ver_qr4c with the three slopes of :417-419 taken from a per-column array: S(j) =
// (ttf(j-1) - ttf(j)) / (Z(j-1) - Z(j)) for j in [nzmin+1, nzmax-1], so that qc = S(k), qu = S(k+1),
// qd = S(k-1).  Every slope is evaluated once per column (by the thread of layer j) instead of three
// times, with the correctly rounded division, hence bit-identical to ver_qr4c.
__device__ __forceinline__ double ver_qr4c_s(const ColV& c, const double* S, int k, double fin)
{
    double v = fin;
    if (k == c.nzmin) v = -CT(k) * CW(k) * CA(k) - v;
    if (k == c.nzmin + 1) v = -0.5 * (CT(k - 1) + CT(k)) * CW(k) * CA(k) - v;
    if (k == c.nzmax - 1) v = -0.5 * (CT(k - 1) + CT(k)) * CW(k) * CA(k) - v;
    if (k == c.nzmax) v = 0.0 - v;
    if (k >= c.nzmin + 2 && k <= c.nzmax - 2) {
        const double qc = S[k - 1], qu = S[k], qd = S[k - 2];
        const double zb = CZB(k);
        const double Tmean1 = CT(k) + div3((2 * qc + qu) * (zb - CZ(k)));
        const double Tmean2 = CT(k - 1) + div3((2 * qc + qd) * (zb - CZ(k - 1)));
        const double w = CW(k);
        const double Tmean = (w + fabs(w)) * Tmean1 + (w - fabs(w)) * Tmean2;
        v = (-0.5 * (1.0 - c.num_ord) * Tmean - c.num_ord * (0.5 * (Tmean1 + Tmean2)) * w) * CA(k) - v;
    }
    return v;
}

// adv_tra_ver_qr4c with the reciprocals of the three layer-centre distances taken from a per-column
// array R(j) = RN(1 / (Z(j-1) - Z(j))), j in [nzmin+1, nzmax-1]: each thread computes ONE reciprocal
// (the only IEEE division left) and the six slopes of :417-419 (3 per tracer) become div_rcp, which is
// bit-identical to the division (adv_selftest_div).
__device__ __forceinline__ double ver_qr4c_r(const ColV& c, const double* R, int k, double fin)
{
    double v = fin;
    if (k == c.nzmin) v = -CT(k) * CW(k) * CA(k) - v;
    if (k == c.nzmin + 1) v = -0.5 * (CT(k - 1) + CT(k)) * CW(k) * CA(k) - v;
    if (k == c.nzmax - 1) v = -0.5 * (CT(k - 1) + CT(k)) * CW(k) * CA(k) - v;
    if (k == c.nzmax) v = 0.0 - v;
    if (k >= c.nzmin + 2 && k <= c.nzmax - 2) {
        const double t0 = CT(k), tm1 = CT(k - 1), tm2 = CT(k - 2), tp1 = CT(k + 1);
        const double z0 = CZ(k), zm1 = CZ(k - 1), zm2 = CZ(k - 2), zp1 = CZ(k + 1);
        const double qc = div_rcp(tm1 - t0, zm1 - z0, R[k - 1]);
        const double qu = div_rcp(t0 - tp1, z0 - zp1, R[k]);
        const double qd = div_rcp(tm2 - tm1, zm2 - zm1, R[k - 2]);
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
                                         double2 g12, double2 g34, double num_ord, double clo1, double clo2,
                                         double fin)
{
    // g12 = (gx_up, gx_dn), g34 = (gy_up, gy_dn): edge_up_dn_grad(1:4,nz,e), oce_adv_tra_hor.F90:431-434
    if (HOR == HOR_UPW1) return -0.5 * (a1 * qp + a2 * qm) - fin;
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

// node decode shared by the node kernels
struct NodeThread {
    int n, nz0, nzmin, nzmax;
    int pad_lo, pad_hi, self_lo, self_hi, deg;
    bool active;
};
__device__ __forceinline__ NodeThread node_thread(const MeshDev& m, const NodeRange& r)
{
    NodeThread t;
    const ColThread c = col_thread(m);
    // Metadata prefetch: every CTA starts with two dependent metadata loads (node_rec, then the ELL
    // row) that nothing else overlaps (ncu: 11-15 % of the node kernels' stall samples).  Thread j < cpb
    // pulls the record and the ELL row of column j of the CTA `pf` launches ahead into L2, so that
    // those loads become L2 hits (1.30 / 1.21 / 1.43 -> 1.25 / 1.14 / 1.37 ms for N1 / K2 / K3).
    if (r.pf > 0 && (int)threadIdx.x < r.cpb) {
        const long long i = ((long long)blockIdx.x + r.pf) * r.cpb + threadIdx.x;
        if (i < r.count) {
            const int nf = r.list ? __ldg(&r.list[r.begin + i]) : r.begin + (int)i;
            const char* row = reinterpret_cast<const char*>(m.ne_ell + (size_t)nf * m.ell_w);
            l2_prefetch_line(&m.node_rec[nf]);
            l2_prefetch_line(row);
            l2_prefetch_line(row + m.ell_w * 16 - 1);
        }
    }
    t.n = 0; t.nzmin = 1; t.nzmax = 0;
    t.pad_lo = 0; t.pad_hi = 255; t.self_lo = 1; t.self_hi = 0; t.deg = 0;
    t.nz0 = c.nz0;
    const int i = blockIdx.x * r.cpb + c.g;
    t.active = i < r.count;
    if (t.active) {
        t.n = r.list ? __ldg(&r.list[r.begin + i]) : r.begin + i;
        const uint2 rec = __ldg(&m.node_rec[t.n]);
        t.nzmin = rec.x & 0xff; t.nzmax = (rec.x >> 8) & 0xff;
        t.pad_lo = (rec.x >> 16) & 0xff; t.pad_hi = rec.x >> 24;
        t.self_lo = rec.y & 0xff; t.self_hi = (rec.y >> 8) & 0xff; t.deg = (rec.y >> 16) & 0xff;
        if (r.skip_s && ((rec.y >> 24) & 1u)) { t.active = false; t.nzmin = 1; t.nzmax = 0; t.deg = 0; }
    }
    return t;
}

// ----------------------------------------------------------------------------------------------
// Wet-level compaction of the FCT node kernels (N1, K2, K3).  A CTA owns a run of WHOLE columns chosen on the
// host so that their wet layers fill the CTA (adv_ctx_create: NodePartHost), and every wet (column, layer) gets
// one thread: no lane is spent below the sea floor (round 1: 23-24 of 32 lanes active).  Warp 0 decodes the
// records of the CTA's columns and a prefix sum of their depths; the adjacency (ELL) rows of all columns are
// staged in shared memory by the whole CTA at the same time (for identity ranges the node id is known
// without any load, so record and rows are fetched in parallel: one dependent wait less than round 1's
// record -> row -> operands chain), then each thread finds its column with a few shared-memory compares.
// ----------------------------------------------------------------------------------------------
constexpr int kMaxCols = 16;       // columns per CTA (<= 32: the prefix sum runs in one warp)
struct NodePart {
    const int* cta_first;          // (ncta + 1) position in the range of the first column of every CTA
    const int* list;               // optional position -> node id (0-based); nullptr = identity
    int ncta;
    int pf;                        // prefetch distance in CTAs (0 = off)
    int skip_s;                    // 1: columns flagged "boundary set" in node_rec get no threads
};
struct NodeSmem {                  // header of the node kernels' dynamic shared memory
    uint2 rec[kMaxCols];
    int n[kMaxCols];
    int off[kMaxCols + 1];         // first thread of column j; off[ncols] = number of threads in use
    int pad_[3];
};
static_assert(sizeof(NodeSmem) % 16 == 0, "NodeSmem must keep the 16-byte alignment of what follows");
__host__ __device__ inline size_t node_smem_header(int ell_w) { return sizeof(NodeSmem) + (size_t)kMaxCols * ell_w * sizeof(int4); }

struct NodeCta {
    int n, nz, nzmin, nzmax;       // node (0-based), layer (1-based), ulevels_nod2D, nlevels_nod2D
    int pad_lo, pad_hi, self_lo, self_hi, deg;
    int base;                      // first thread of this thread's column: column-local index of layer k is base + k - nzmin
    bool valid;                    // this thread owns a wet (column, layer)
    const int4* ell;               // the column's adjacency row in shared memory
};

// prefetch (L2) what the CTA `pf` launches ahead will wait for first: its column records and adjacency rows
// (one contiguous block each for identity ranges) and the wet part of its own columns of `arr`
struct PfArr { const void* base; unsigned col_bytes, lev_bytes; int iface; };   // column stride, bytes per level, 1: nl interfaces per column
template <int NA, class F>
__device__ __forceinline__ void node_cta_prefetch(const MeshDev& m, const NodePart& r, F arr, bool own)
{
    if (r.pf <= 0 || (int)threadIdx.x < 32) return;      // warp 1 onwards: warp 0 is busy with the prefix sum
    const int cta = (int)blockIdx.x + r.pf;
    if (cta >= r.ncta) return;
    const int c0 = __ldg(&r.cta_first[cta]), c1 = __ldg(&r.cta_first[cta + 1]);
    const int na = own ? NA : 0;
    for (int k = (int)threadIdx.x - 32; k < (na + 1) * kMaxCols; k += (int)blockDim.x - 32) {
        const int a = k / kMaxCols, j = k - a * kMaxCols;     // a == 0: metadata, a >= 1: array a-1; column j
        if (c0 + j >= c1) continue;
        const int n = r.list ? __ldg(&r.list[c0 + j]) : c0 + j;
        if (a == 0) {
            const char* row = reinterpret_cast<const char*>(m.ne_ell + (size_t)n * m.ell_w);
            l2_prefetch_line(&m.node_rec[n]);
            l2_prefetch_line(row);
            l2_prefetch_line(row + m.ell_w * 16 - 1);
            continue;
        }
        if (n >= m.N) continue;                                // halo columns (K3): not every array covers them
        const uint2 rec = __ldg(&m.node_rec[n]);
        if (r.skip_s && ((rec.y >> 24) & 1u)) continue;
        const unsigned lo = rec.x & 0xff, hi = (rec.x >> 8) & 0xff;      // layers lo .. hi-1, interfaces lo .. hi
        PfArr x = arr(0);               // arr(q) with a compile-time q: nothing is materialised in local memory
#pragma unroll
        for (int q = 1; q < NA; ++q) if (a - 1 == q) x = arr(q);
        l2_prefetch(reinterpret_cast<const char*>(x.base) + (size_t)n * x.col_bytes + (size_t)(lo - 1) * x.lev_bytes,
                    (hi - lo + (x.iface ? 1u : 0u)) * x.lev_bytes);
    }
}

__device__ __forceinline__ NodeCta node_cta(const MeshDev& m, const NodePart& r, unsigned char* smem_raw)
{
    NodeSmem* h = reinterpret_cast<NodeSmem*>(smem_raw);
    int4* s_ell = reinterpret_cast<int4*>(smem_raw + sizeof(NodeSmem));
    const int tid = threadIdx.x, ell_w = m.ell_w;
    const int c0 = __ldg(&r.cta_first[blockIdx.x]), c1 = __ldg(&r.cta_first[blockIdx.x + 1]);
    const int ncols = c1 - c0;
    if (tid < 32) {
        int n = 0, wet = 0;
        uint2 rec = make_uint2(0u, 0u);
        if (tid < ncols) {
            n = r.list ? __ldg(&r.list[c0 + tid]) : c0 + tid;
            rec = __ldg(&m.node_rec[n]);
            wet = (int)((rec.x >> 8) & 0xff) - (int)(rec.x & 0xff);            // layers ulev .. nlev-1
            if (r.skip_s && ((rec.y >> 24) & 1u)) wet = 0;
        }
        int incl = wet;
#pragma unroll
        for (int d = 1; d < kMaxCols; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (tid >= d) incl += v;
        }
        if (tid < ncols) { h->n[tid] = n; h->rec[tid] = rec; h->off[tid] = incl - wet; }
        if (tid == ncols - 1) h->off[ncols] = incl;
    }
    for (int i = tid; i < ncols * ell_w; i += blockDim.x) {
        const int col = i / ell_w;
        const int n = r.list ? __ldg(&r.list[c0 + col]) : c0 + col;
        s_ell[i] = __ldg(&m.ne_ell[(size_t)n * ell_w + (i - col * ell_w)]);
    }
    __syncthreads();
    int g = 0;
    for (int j = 1; j < ncols; ++j) g += (tid >= h->off[j]) ? 1 : 0;
    NodeCta t;
    const uint2 rec = h->rec[g];
    t.n = h->n[g];
    t.base = h->off[g];
    t.nzmin = rec.x & 0xff; t.nzmax = (rec.x >> 8) & 0xff;
    t.pad_lo = (rec.x >> 16) & 0xff; t.pad_hi = rec.x >> 24;
    t.self_lo = rec.y & 0xff; t.self_hi = (rec.y >> 8) & 0xff; t.deg = (rec.y >> 16) & 0xff;
    t.nz = t.nzmin + (tid - t.base);
    t.valid = tid < h->off[ncols] && t.nz <= t.nzmax - 1;
    if (!t.valid) t.deg = 0;
    t.ell = s_ell + g * ell_w;
    return t;
}

// an empty gather slot: lo = 255 > hi = 0, never in range (nl <= 255)
#define ADV_EMPTY_SLOT make_int4(0, 0, 0xff, 0)

// ----------------------------------------------------------------------------------------------
// E1: antidiffusive horizontal flux HO-LO per (edge, layer), every operand read once.
//   reference: oce_adv_tra_driver.F90:115 (LO = adv_tra_hor_upw1 on ttf), :343-354 (HO on ttfAB
//   with o_init_zero=.false.: flux := HO - LO), bodies oce_adv_tra_hor.F90:214-216, :446-461,
//   :488-489, :736-751, :777-778.
// QMODE 0: compute the volume flux Q from uv/helem and store it (first chunk of a step);
// QMODE 1: read the stored Q (later chunks of the same step: geometry reads amortised).
// ----------------------------------------------------------------------------------------------
template <int HOR, int TB, int QMODE>
__global__ void __launch_bounds__(kBlock, ADV_E1_MINB) k_edge_flux(MeshDev m, Chunk<TB> b, int epb)
{
    const ColThread c = col_thread(m);
    const int L = m.L;
    const int e = blockIdx.x * epb + c.g;
    if (e >= m.E) return;
    const int nz0 = c.nz0, nz = nz0 + 1;
    // ---- wait 1: all per-edge metadata, issued together ---------------------------------------
    const int4 em = __ldg(&m.edge_meta[e]);          // {edges(1,e), edges(2,e), el1, el2} 0-based, el2 = -1: none
    const uchar4 lv = __ldg(&m.edge_lev[e]);
    const double2 cr12 = __ldg(reinterpret_cast<const double2*>(&m.edge_cross[e]));
    const double2 cr34 = __ldg(reinterpret_cast<const double2*>(&m.edge_cross[e]) + 1);
    double2 ec = make_double2(0.0, 0.0);
    if (HOR != HOR_UPW1) ec = __ldg(&m.edge_c[e]);
    const unsigned oe = (unsigned)e * L + nz0, o1 = (unsigned)em.x * L + nz0, o2 = (unsigned)em.y * L + nz0;
    // scatter range of oce_adv_tra_driver.F90:154-156: [min(nu1, nu2>0), max(nl1, nl2)]
    const int lo = lv.z > 0 ? min((int)lv.x, (int)lv.z) : (int)lv.x;
    const int hi = max((int)lv.y, (int)lv.w);
    const bool inr = nz >= lo && nz <= hi;
    bool use1 = false, use2 = false;
    if (QMODE == 0) edge_use(lv, nz, use1, use2);
    // ---- wait 2: all operand loads, issued together -------------------------------------------
    double t1[TB], t2[TB], a1[TB], a2[TB];
    double2 g12[TB], g34[TB];
    double2 uv1 = make_double2(0.0, 0.0), uv2 = uv1;
    double he1 = 0.0, he2 = 0.0, q = 0.0, clo1 = 1.0, clo2 = 1.0;
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        t1[t] = t2[t] = a1[t] = a2[t] = 0.0;
        g12[t] = g34[t] = make_double2(0.0, 0.0);
    }
    if (inr) {
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            if (HOR != HOR_UPW1) {
                const double2* gp = reinterpret_cast<const double2*>(b.grad[t]) + (size_t)oe * 2;
                g12[t] = __ldg(gp);
                g34[t] = __ldg(gp + 1);
            }
            t1[t] = __ldg(&b.ttf[t][o1]); t2[t] = __ldg(&b.ttf[t][o2]);
            a1[t] = __ldg(&b.ttfAB[t][o1]); a2[t] = __ldg(&b.ttfAB[t][o2]);
        }
        if (HOR == HOR_MUSCL) {
            clo1 = (__ldg(&m.nboundary_lay[em.x]) - nz >= 0) ? 1.0 : 0.0;   // oce_adv_tra_hor.F90:411-412
            clo2 = (__ldg(&m.nboundary_lay[em.y]) - nz >= 0) ? 1.0 : 0.0;
        }
        if (QMODE == 1) q = __ldg(&m.Q[oe]);
    }
    if (QMODE == 0) {
        if (use1) {
            const unsigned o = (unsigned)em.z * L + nz0;
            uv1 = __ldg(reinterpret_cast<const double2*>(m.uv) + o); he1 = __ldg(&m.helem[o]);
        }
        if (use2) {
            const unsigned o = (unsigned)em.w * L + nz0;
            uv2 = __ldg(reinterpret_cast<const double2*>(m.uv) + o); he2 = __ldg(&m.helem[o]);
        }
        // Q(nz,e): vflux of oce_adv_tra_hor.F90:170,190,211-212,226,242
        const double v1 = (-uv1.y * cr12.x + uv1.x * cr12.y) * he1;
        const double v2 = (uv2.y * cr34.x - uv2.x * cr34.y) * he2;
        q = 0.0;
        if (use1 && use2) q = v1 + v2;
        else if (use1) q = v1;
        else if (use2) q = v2;
        m.Q[oe] = q;
    }
    double out[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) out[t] = 0.0;
    if (inr) {
        const double aq = fabs(q), qp = q + aq, qm = q - aq;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double flo = b.nolo ? 0.0 : hor_lo(t1[t], t2[t], qp, qm);
            out[t] = hor_ho<HOR>(a1[t], a2[t], q, qp, qm, ec, g12[t], g34[t], b.ph[t], clo1, clo2, flo);
        }
    }
    stv<TB>(b.adf_h + (size_t)oe * TB, out);
}

// ----------------------------------------------------------------------------------------------
// N1: LO solution + vertical antidiffusive flux (owned nodes)
//   reference: oce_adv_tra_driver.F90:115-252 (D1-D5), :363-379 (D8)
// The upwind edge flux is recomputed from Q and the two end values (3 flops) and added in
// ascending-edge order, the serial order of the reference's scatter loop (:142-201).
// Every thread first issues ALL its global loads (own level of the column operands + the first
// gather batch), parks the column operands in shared memory and evaluates the vertical stencils
// from there: one memory wait per thread instead of one per stencil point.
// ----------------------------------------------------------------------------------------------
template <int VER, int TB>
__host__ __device__ constexpr int n1_smem_arrays() { return 3 * TB + 5 + (VER == VER_PPM ? 2 : 0) + (VER == VER_QR4C && ADV_QR4C_RCP ? 1 : 0); }

template <int VER, int TB, int G>
__global__ void ADV_N1_BOUNDS k_node_lo(MeshDev m, Chunk<TB> b, NodePart r, double dt)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // after the header: n1_smem_arrays() arrays of [nthr] doubles; element index == threadIdx.x (compact wet layers)
    double* sm = reinterpret_cast<double*>(smem_raw + node_smem_header(m.ell_w));
    const int L = m.L, nl = m.nl, nthr = blockDim.x, tid = threadIdx.x;
    node_cta_prefetch<2 * TB + 8>(m, r, [&](int q) {
        const unsigned cl = (unsigned)L * 8u, cn = (unsigned)nl * 8u;
        if (q < 2 * TB) return PfArr{(q & 1) ? b.ttfAB[q >> 1] : b.ttf[q >> 1], cl, 8u, 0};
        switch (q - 2 * TB) {
        case 0: return PfArr{m.hnode, cl, 8u, 0};
        case 1: return PfArr{m.hnode_new, cl, 8u, 0};
        case 2: return PfArr{m.Z3d, cl, 8u, 0};
        case 3: return PfArr{m.zbar3d, cn, 8u, 1};
        case 4: return PfArr{m.w, cn, 8u, 1};
        case 5: return PfArr{m.we, cn, 8u, 1};
        case 6: return PfArr{m.area, cn, 8u, 1};
        default: return PfArr{m.areasvol, cn, 8u, 1};
        }
    }, (ADV_PF_OWN & 1) != 0);
    const NodeCta th = node_cta(m, r, smem_raw);
    const int n = th.n, nz = th.nz, nz0 = nz - 1, nzmin = th.nzmin, nzmax = th.nzmax;
    const bool valid = th.valid;
    const unsigned oL = (unsigned)n * L + nz0;
    const size_t cN = (size_t)n * nl;
    double* s_flo = sm;                               // [TB][nthr] LO vertical flux at the top interface
    double* s_ttf = sm + (size_t)TB * nthr;           // [TB][nthr]
    double* s_tab = s_ttf + (size_t)TB * nthr;        // [TB][nthr]
    double* s_Z = s_tab + (size_t)TB * nthr;
    double* s_zbar = s_Z + nthr;
    double* s_w = s_zbar + nthr;
    double* s_we = s_w + nthr;
    double* s_area = s_we + nthr;
    double* s_hn = s_area + nthr;                     // PPM only
    double* s_hnn = s_hn + nthr;                      // PPM only
    double* s_rdz = s_area + nthr;                    // QR4C only (aliases s_hn, which QR4C does not use)

    // ---- all global loads ---------------------------------------------------------------------
    int4 ent[G];
    double q[G], to[G][TB];
    bool in[G];
    double tn[TB], tab[TB];
    double av = 1.0, hn = 0.0, hnn = 1.0, zz = 0.0, zb = 0.0, ww = 0.0, wwe = 0.0, ar = 0.0;
#pragma unroll
    for (int j = 0; j < G; ++j) ent[j] = (j < th.deg) ? th.ell[j] : ADV_EMPTY_SLOT;
#pragma unroll
    for (int t = 0; t < TB; ++t) { tn[t] = 0.0; tab[t] = 0.0; }
    if (valid) {
#pragma unroll
        for (int t = 0; t < TB; ++t) { tn[t] = __ldg(&b.ttf[t][oL]); tab[t] = __ldg(&b.ttfAB[t][oL]); }
        av = __ldg(&m.areasvol[cN + nz0]); hn = __ldg(&m.hnode[oL]); hnn = __ldg(&m.hnode_new[oL]);
        zz = __ldg(&m.Z3d[oL]); zb = __ldg(&m.zbar3d[cN + nz0]);
        ww = __ldg(&m.w[cN + nz0]); wwe = __ldg(&m.we[cN + nz0]); ar = __ldg(&m.area[cN + nz0]);
    }
    double zup = 0.0;                                 // Z of the layer above (an L1 hit: the thread above loads it too)
    const bool has_rdz = VER == VER_QR4C && ADV_QR4C_RCP && valid && nz > nzmin;
    if (has_rdz) zup = __ldg(&m.Z3d[oL - 1]);
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
        in[j] = nz >= lo && nz <= hi;
        q[j] = 0.0;
#pragma unroll
        for (int t = 0; t < TB; ++t) to[j][t] = 0.0;
        if (in[j]) {
            q[j] = __ldg(&m.Q[(unsigned)ent[j].x * L + nz0]);
            const unsigned oo = (unsigned)ent[j].y * L + nz0;
#pragma unroll
            for (int t = 0; t < TB; ++t) to[j][t] = __ldg(&b.ttf[t][oo]);
        }
    }
    // ---- park the column operands in shared memory ---------------------------------------------
#pragma unroll
    for (int t = 0; t < TB; ++t) { s_ttf[t * nthr + tid] = tn[t]; s_tab[t * nthr + tid] = tab[t]; }
    s_Z[tid] = zz; s_zbar[tid] = zb; s_w[tid] = ww; s_we[tid] = wwe; s_area[tid] = ar;
    if (VER == VER_PPM) { s_hn[tid] = hn; s_hnn[tid] = hnn; }
    if (VER == VER_QR4C && ADV_QR4C_RCP) s_rdz[tid] = has_rdz ? 1.0 / (zup - zz) : 0.0;
    __syncthreads();

    // ---- horizontal LO gather: ordered accumulation ----------------------------------------------
    double losum[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) losum[t] = 0.0;
    if (valid) {
        for (int j0 = 0;;) {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if (!in[j]) continue;
                const bool second = (ent[j].z >> 16) & 1;
                const double aq = fabs(q[j]), qp = q[j] + aq, qm = q[j] - aq;
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (!second) losum[t] = losum[t] + hor_lo(tn[t], to[j][t], qp, qm);        // driver :175
                    else losum[t] = losum[t] - hor_lo(to[j][t], tn[t], qp, qm);                // driver :188
                }
            }
            j0 += G;
            if (j0 >= th.deg) break;
            // further batches (degree > G)
#pragma unroll
            for (int j = 0; j < G; ++j) ent[j] = (j0 + j < th.deg) ? th.ell[j0 + j] : ADV_EMPTY_SLOT;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
                in[j] = nz >= lo && nz <= hi;
                if (in[j]) {
                    q[j] = __ldg(&m.Q[(unsigned)ent[j].x * L + nz0]);
                    const unsigned oo = (unsigned)ent[j].y * L + nz0;
#pragma unroll
                    for (int t = 0; t < TB; ++t) to[j][t] = __ldg(&b.ttf[t][oo]);
                }
            }
        }
    }

    // ---- vertical fluxes at the thread's top interface, stencils read from shared memory --------
    // (the bottom interface nzmax always carries +0.0: every scheme writes 0 - 0 there)
    double flo_top[TB], adfv_top[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { flo_top[t] = 0.0; adfv_top[t] = 0.0; }
    if (valid) {
        const int c0 = th.base - (nzmin - 1);         // shared-memory index of (level 1) of this thread's column
        ColV c;
        c.area = s_area + c0; c.Z = s_Z + c0; c.zbar = s_zbar + c0;
        c.hnode = s_hn + c0; c.hnode_new = s_hnn + c0;
        c.nzmin = nzmin; c.nzmax = nzmax; c.dt = dt;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            c.ttf = s_ttf + t * nthr + c0; c.w = s_we + c0; c.num_ord = 0.0;
            const double fe = ver_upw1(c, nz, 0.0);                     // driver :235
            double flo = fe;
            if (m.use_wsplit) { c.w = s_w + c0; flo = ver_upw1(c, nz, 0.0); }  // driver :333
            c.ttf = s_tab + t * nthr + c0; c.w = s_w + c0; c.num_ord = b.pv[t];
            flo_top[t] = fe;
            if (VER == VER_QR4C && ADV_QR4C_RCP) adfv_top[t] = ver_qr4c_r(c, s_rdz + c0, nz, flo);
            else adfv_top[t] = ver_flux<VER>(c, nz, flo);               // driver :363-379
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) s_flo[t * nthr + tid] = flo_top[t];
    __syncthreads();
    if (!valid) return;
    stv<TB>(b.adf_v + (cN + nz0) * TB, adfv_top);
    const bool has_below = nz + 1 <= nzmax - 1;              // the next thread is the layer below of the same column
    if (!has_below) {                                        // interface nzmax: the (zero) bottom
        double z[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) z[t] = 0.0;
        stv<TB>(b.adf_v + (cN + nz0 + 1) * TB, z);
    }
    const double r_av = 1.0 / av, r_hnn = 1.0 / hnn;
    double lo_out[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const double flo_bot = has_below ? s_flo[t * nthr + tid + 1] : 0.0;
        const double fv = flo_top[t] - flo_bot;                                          // fv(nz)-fv(nz+1)
        const double num = tn[t] * hn + div_rcp((losum[t] + fv) * dt, av, r_av);
        lo_out[t] = div_rcp(num, hnn, r_hnn);                                            // driver :249
        // ltra_diag: the low-order parts of tra_advhoriz / tra_advvert (driver :221-229, :307-318); plain IEEE divisions
        if (b.dgh[t]) b.dgh[t][oL] = losum[t] * dt / av / hnn;
        if (b.dgv[t]) b.dgv[t][oL] = fv * dt / av / hnn;
    }
    stv<TB>(b.lo + (size_t)oL * TB, lo_out);
}

// ----------------------------------------------------------------------------------------------
// adv_tra_vert_impl (oce_adv_tra_ver.F90:120-236) on the TB tracer-interleaved fct_LO columns of a chunk.
// Thread = (column, layer) like the other node kernels, three phases:
//   A  every thread evaluates the tridiagonal coefficients a, b, c of its layer (tracer-independent) and the
//      right-hand sides tr[t] from coalesced loads and parks them in shared memory;
//   B  ONE thread per column runs the Thomas recurrences -- the forward elimination cp(nz) = c/(b - cp(nz-1) a),
//      tp(nz) = (tr - tp(nz-1) a)/(b - cp(nz-1) a) and the back substitution -- in the reference's order: the
//      chain is serial by definition and any reassociation (cyclic reduction) would change the rounding;
//      cp is shared by the TB tracers;
//   C  every thread adds its layer's increment to fct_LO with a coalesced store.
// Replaces round 1's one-thread-per-column kernel (stride-L accesses, scratch arrays in global memory).
// ----------------------------------------------------------------------------------------------
template <int TB>
__global__ void __launch_bounds__(kBlock) k_vert_impl(MeshDev m, double* __restrict__ lo, NodeRange r, double dt)
{
    extern __shared__ double sm[];            // [3 + TB][blockDim]: a, b, c (later cp), tr[t] (later tp[t], then the increment)
    const int L = m.L, nl = m.nl, nthr = blockDim.x, tid = threadIdx.x;
    const NodeThread th = node_thread(m, r);
    const int n = th.n, nz0 = th.nz0, nz = nz0 + 1, nzmin = th.nzmin, nzmax = th.nzmax;   // nzmax = nlevels_nod2D
    const bool valid = th.active && n < m.N && nz >= nzmin && nz <= nzmax - 1;
    double* s_a = sm; double* s_b = sm + nthr; double* s_c = sm + 2 * nthr; double* s_t = sm + 3 * nthr;
    const size_t oL = (size_t)n * L + nz0, cN = (size_t)n * nl + nz0;
    if (valid) {
        const double zinv = 1.0 * dt;
        const double hn = __ldg(&m.hnode_new[oL]), av = __ldg(&m.areasvol[cN]);
        const double w0 = __ldg(&m.wi[cN]), a0 = __ldg(&m.area[cN]);
        double a, bb, c;
        double tm[TB], t0[TB], tp[TB], tr[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            t0[t] = lo[oL * TB + t];
            tm[t] = nz > nzmin ? lo[(oL - 1) * TB + t] : 0.0;
            tp[t] = nz <= nzmax - 2 ? lo[(oL + 1) * TB + t] : 0.0;
        }
        if (nz == nzmin) {                                      // :154-170, :198-200
            const double w1 = __ldg(&m.wi[cN + 1]), a1 = __ldg(&m.area[cN + 1]);
            a = 0.0;
            double v_adv = zinv * a0 / av;
            bb = hn + w0 * v_adv;
            v_adv = zinv * a1 / av;
            bb = bb - dmin(0.0, w1) * v_adv;
            c = -dmax(0.0, w1) * v_adv;
#pragma unroll
            for (int t = 0; t < TB; ++t) tr[t] = -(bb - hn) * t0[t] - c * tp[t];
        } else if (nz <= nzmax - 2) {                           // :174-183, :202-205
            const double w1 = __ldg(&m.wi[cN + 1]), a1 = __ldg(&m.area[cN + 1]);
            double v_adv = zinv * a0 / av;
            a = dmin(0.0, w0) * v_adv;
            bb = hn + dmax(0.0, w0) * v_adv;
            v_adv = zinv * a1 / av;
            bb = bb - dmin(0.0, w1) * v_adv;
            c = -dmax(0.0, w1) * v_adv;
#pragma unroll
            for (int t = 0; t < TB; ++t) tr[t] = -a * tm[t] - (bb - hn) * t0[t] - c * tp[t];
        } else {                                                // :187-195, :206-208
            const double v_adv = zinv * a0 / av;
            a = dmin(0.0, w0) * v_adv;
            bb = hn + dmax(0.0, w0) * v_adv;
            c = 0.0;
#pragma unroll
            for (int t = 0; t < TB; ++t) tr[t] = -a * tm[t] - (bb - hn) * t0[t];
        }
        s_a[tid] = a; s_b[tid] = bb; s_c[tid] = c;
#pragma unroll
        for (int t = 0; t < TB; ++t) s_t[t * nthr + tid] = tr[t];
    }
    __syncthreads();
    if (valid && nz == nzmin) {                                 // the column's serial part
        const int nlay = nzmax - nzmin;                         // layers nzmin .. nzmax-1 live at tid .. tid+nlay-1
        double cp_prev = 0.0, tp_prev[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) tp_prev[t] = 0.0;
        for (int k = 0; k < nlay; ++k) {                        // :211-221
            const double a = s_a[tid + k], bb = s_b[tid + k], c = s_c[tid + k];
            if (k == 0) {
                cp_prev = c / bb;
#pragma unroll
                for (int t = 0; t < TB; ++t) tp_prev[t] = s_t[t * nthr + tid] / bb;
            } else {
                const double mm = bb - cp_prev * a;
                cp_prev = c / mm;
#pragma unroll
                for (int t = 0; t < TB; ++t) tp_prev[t] = (s_t[t * nthr + tid + k] - tp_prev[t] * a) / mm;
            }
            s_c[tid + k] = cp_prev;
#pragma unroll
            for (int t = 0; t < TB; ++t) s_t[t * nthr + tid + k] = tp_prev[t];
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) {                          // :224-235: the last layer's increment is tp itself
            double trn = s_t[t * nthr + tid + nlay - 1];
            for (int k = nlay - 2; k >= 0; --k) {
                trn = s_t[t * nthr + tid + k] - s_c[tid + k] * trn;
                s_t[t * nthr + tid + k] = trn;
            }
        }
    }
    __syncthreads();
    if (valid) {
#pragma unroll
        for (int t = 0; t < TB; ++t) lo[oL * TB + t] = lo[oL * TB + t] + s_t[t * nthr + tid];
    }
}

// ----------------------------------------------------------------------------------------------
// vert_vel_ale, continuity part for the linear free surface (src/oce_ale.F90:2164-2310, no Fer_GM):
// the reference scatters, edge after edge, the transport through the two half-edges (element 1, then element 2)
// into Wvel of both end nodes, sums each column bottom-up and divides by the cell area.  Here: an ordered gather
// over the node's edge slots (ascending edge = the serial scatter order; per edge first the element-1 term, then
// the element-2 term, each its own addition), the bottom-up running sum by one thread per column (a serial
// chain, kept in the reference's order), the division in parallel.  Owned columns; every other entry of the
// column is zeroed like :2165.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_vert_vel_ale(MeshDev m, NodeRange r, double* __restrict__ W)
{
    extern __shared__ double sm[];            // [blockDim]
    const int L = m.L, nl = m.nl, tid = threadIdx.x;
    const NodeThread th = node_thread(m, r);
    const int n = th.n, nz0 = th.nz0, nz = nz0 + 1, nzmin = th.nzmin, nzmax = th.nzmax - 1;   // nzmax: last layer
    const bool wet = th.active && nz >= nzmin && nz <= nzmax;
    double acc = 0.0;
    if (wet) {
        const int4* ell = m.ne_ell + (size_t)n * m.ell_w;
        for (int j = 0; j < th.deg; ++j) {
            const int4 ent = __ldg(&ell[j]);
            const int e = ent.x;
            const bool second = (ent.z >> 16) & 1;
            const uchar4 lv = __ldg(&m.edge_lev[e]);
            const int2 el = __ldg(&m.edge_el[e]);
            const double2 cr12 = __ldg(reinterpret_cast<const double2*>(&m.edge_cross[e]));
            const double2 cr34 = __ldg(reinterpret_cast<const double2*>(&m.edge_cross[e]) + 1);
            if (nz >= (int)lv.x && nz <= (int)lv.y) {                                   // :2178-2202
                const unsigned o = (unsigned)el.x * L + nz0;
                const double2 uv = __ldg(reinterpret_cast<const double2*>(m.uv) + o);
                const double c1 = (uv.y * cr12.x - uv.x * cr12.y) * __ldg(&m.helem[o]);
                acc = second ? acc - c1 : acc + c1;
            }
            if (el.y >= 0 && nz >= (int)lv.z && nz <= (int)lv.w) {                      // :2213-2240
                const unsigned o = (unsigned)el.y * L + nz0;
                const double2 uv = __ldg(reinterpret_cast<const double2*>(m.uv) + o);
                const double c1 = -(uv.y * cr34.x - uv.x * cr34.y) * __ldg(&m.helem[o]);
                acc = second ? acc - c1 : acc + c1;
            }
        }
    }
    sm[tid] = acc;
    __syncthreads();
    if (wet && nz == nzmin) {                                   // :2277-2286: W(nz) = W(nz) + W(nz+1), bottom-up
        double below = 0.0;
        for (int k = nzmax - nzmin; k >= 0; --k) { below = sm[tid + k] + below; sm[tid + k] = below; }
    }
    __syncthreads();
    if (!th.active) return;
    const size_t cN = (size_t)n * nl + nz0;
    W[cN] = wet ? sm[tid] / __ldg(&m.area[cN]) : 0.0;           // :2301-2308
    if (nz0 == L - 1) W[cN + 1] = 0.0;                          // interface nl
}

// vert_vel_ale, 'zstar' correction (src/oce_ale.F90:2539-2603): the elevation change hbar - hbar_old is distributed over
// the layers above the shallowest bottom around the node (Wvel and hnode_new), the surface fresh-water flux closes
// the continuity at the top.  Owned, cavity-free columns; one thread per (column, layer), no dependence between layers.
__global__ void __launch_bounds__(kBlock) k_vert_vel_zstar(MeshDev m, NodeRange r, double dt, const int* __restrict__ nmin,
                                                           const double* __restrict__ hbar, const double* __restrict__ hbar_old,
                                                           const double* __restrict__ wflux, double* __restrict__ W, double* __restrict__ hnode_new)
{
    const NodeThread th = node_thread(m, r);
    if (!th.active || th.nzmin != 1) return;                                        // :2550
    const int n = th.n, nz = th.nz0 + 1, nzmin = th.nzmin, nzmax = __ldg(&nmin[n]) - 1;
    const double* zb = m.zbar3d + (size_t)n * m.nl;
    const double dd1 = __ldg(&zb[nzmax - 1]);
    double dd = __ldg(&zb[nzmin - 1]) - dd1;
    dd = (__ldg(&hbar[n]) - __ldg(&hbar_old[n])) / dd;
    const double dddt = dd / dt;
    const size_t cN = (size_t)n * m.nl + th.nz0, oL = (size_t)n * m.L + th.nz0;
    if (nz >= nzmin && nz <= nzmax - 1) {                                           // :2574-2589
        double w = W[cN];
        w = w - (__ldg(&zb[nz - 1]) - dd1) * dddt;
        if (nz == nzmin) w = w - __ldg(&wflux[n]);                                  // :2595
        W[cN] = w;
        hnode_new[oL] = __ldg(&m.hnode[oL]) + (__ldg(&zb[nz - 1]) - __ldg(&zb[nz])) * dd;
    } else if (nz == nzmin) {
        W[cN] = W[cN] - __ldg(&wflux[n]);                                           // empty stretch range: only the flux
    }
}

// which_ALE = 'zlevel' (src/oce_ale.F90:2336-2538): one thread per owned, cavity-free column -- the correction touches at
// most the first lzstar_lev (namelist default 4) layers of a column and is a short serial recurrence over them.  The
// elevation change goes into the surface layer (:2519-2520) unless that layer would shrink below min_hnode times its
// rest thickness: then it is spread downwards over the layers that still have room ("local zstar", :2367-2449; the
// reference's pairwise "cumsum" at :2397-2398 decides how many layers take part), and a later rise refills the squeezed
// subsurface layers before the surface layer (:2461-2510).  zbar = mesh%zbar (nl); cfl_old = CFL_z of the previous step.
constexpr int kMaxLzstar = 16;
__global__ void __launch_bounds__(256) k_vert_vel_zlevel(MeshDev m, double dt, const int* __restrict__ nmin, const double* __restrict__ hbar,
                                                         const double* __restrict__ hbar_old, const double* __restrict__ wflux,
                                                         const double* __restrict__ zbar, const double* __restrict__ cfl_old, double min_hnode,
                                                         int lz, double* __restrict__ W, double* __restrict__ hnode_new)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= m.N) return;
    if (m.node_lev[n].x != 1) return;                                               // :2354: a cavity is treated like linfs
    double* w = W + (size_t)n * m.nl;
    double* hn = hnode_new + (size_t)n * m.L;
    const double* h = m.hnode + (size_t)n * m.L;
    int nzmax = __ldg(&nmin[n]) - 1;
    const double dh = __ldg(&hbar[n]) - __ldg(&hbar_old[n]);                        // :2357
    double mx[kMaxLzstar];
    if (dh < 0.0 && h[0] + dh <= (zbar[0] - zbar[1]) * min_hnode) {                 // :2367 local zstar
#pragma unroll 1
        for (int k = 0; k < lz; ++k) {                                              // :2374-2383
            double v = (zbar[k] - zbar[k + 1]) * min_hnode - h[k];
            if (v >= 0.0) v = 0.0;
            if (__ldg(&cfl_old[(size_t)n * m.nl + k]) >= 0.95) v = 0.0;
            mx[k] = v;
        }
        int nz = lz;                                                                // :2397-2400
        for (int k = lz - 1; k >= 0; --k) {
            const double cs = k == 0 ? mx[0] : mx[k] + mx[k - 1];
            if (cs < dh) nz = k + 1;
        }
        nzmax = min(nz, nzmax - 1);                                                 // :2411
        double rest = dh, distrib[kMaxLzstar];
        for (int k = 0; k < nzmax; ++k) {                                           // :2412-2416
            distrib[k] = dmax(rest, mx[k]);
            rest = rest - distrib[k];
            rest = dmin(0.0, rest);
        }
        double integ = 0.0;
        for (int k = nzmax - 1; k >= 0; --k) {                                      // :2438-2449
            integ = integ + distrib[k];
            w[k] = w[k] - integ / dt;
            hn[k] = h[k] + distrib[k];
        }
    } else {
        int last_ne = -1;
        bool any_ne = false;
        if (dh > 0.0)
            for (int k = 0; k < lz; ++k)
                if (h[k] != zbar[k] - zbar[k + 1]) { last_ne = k + 1; any_ne = any_ne || k >= 1; }   // :2462-2463, :2482
        if (dh > 0.0 && any_ne) {                                                   // refill, :2461-2510
            nzmax = min(last_ne, nzmax - 1);                                        // :2488
            double rest = dh, integ = 0.0;
            for (int k = nzmax - 1; k >= 0; --k) {
                const double cap = k == 0 ? 1000.0 : (zbar[k] - zbar[k + 1]) - h[k];     // :2471-2475
                const double d = dmin(rest, cap);
                rest = rest - d;
                rest = dmax(0.0, rest);
                integ = integ + d;
                w[k] = w[k] - integ / dt;
                hn[k] = h[k] + d;
            }
        } else {                                                                    // :2519-2520
            w[0] = w[0] - dh / dt;
            hn[0] = h[0] + dh;
        }
    }
    w[0] = w[0] - __ldg(&wflux[n]);                                                 // :2527
}

// compute_CFLz (src/oce_ale.F90:2933-2952, without the diagnostic print) and compute_Wvel_split (:3033-3047):
// one thread per (interface, node) over all myDim+eDim columns.  CFL_z(nz) = c2 of the layer above, then + c1 of
// the layer below, in that order; W_e / W_i are written for nzmin..nlevels_nod2D only, like the reference.
__global__ void __launch_bounds__(256) k_cflz_wsplit(MeshDev m, const double* __restrict__ hnode_new, double dt, int use_wsplit, double maxcfl,
                                                     const double* __restrict__ W, double* __restrict__ We, double* __restrict__ Wi, double* __restrict__ cflz)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)m.Nh * m.nl) return;
    const int n = (int)(idx / m.nl), nz = (int)(idx - (size_t)n * m.nl) + 1;
    const uchar4 lv = __ldg(&m.node_lev[n]);
    const int nzmin = lv.x, nlev = lv.y, nzmax = nlev - 1;
    const double w = W[idx];
    double cfl = 0.0;
    if (nz - 1 >= nzmin && nz - 1 <= nzmax) cfl = fabs(w * dt / __ldg(&hnode_new[(size_t)n * m.L + nz - 2]));            // c2 of layer nz-1
    if (nz >= nzmin && nz <= nzmax) cfl = cfl + fabs(w * dt / __ldg(&hnode_new[(size_t)n * m.L + nz - 1]));              // + c1 of layer nz
    if (cflz) cflz[idx] = cfl;
    if (nz < nzmin || nz > nlev) return;
    double we = w, wi = 0.0;
    if (use_wsplit && cfl > maxcfl) {
        const double dd = dmax(cfl - maxcfl, 0.0) / dmax(maxcfl, 1.e-12);
        we = (1.0 / (1.0 + dd)) * w;
        wi = (dd / (1.0 + dd)) * w;
    }
    We[idx] = we; Wi[idx] = wi;
}

// limited vertical antidiffusive flux at interface k of a column (oce_adv_tra_fct.F90:425-455):
// f = adf_v(k); (pa, ma) = R+/R- of layer k-1, (pk, mk) of layer k
__device__ __forceinline__ double limit_v(double f, int k, int nzmin, int nzmax, double pa, double ma, double pk, double mk)
{
    double ae = 1.0;
    if (k == nzmin) {                                                 // :430-438
        ae = (f >= 0.0) ? dmin(ae, pk) : dmin(ae, mk);
    } else if (k <= nzmax - 1) {                                      // :442-453
        if (f >= 0.0) { ae = dmin(ae, ma); ae = dmin(ae, pk); }
        else { ae = dmin(ae, pa); ae = dmin(ae, mk); }
    }                                                                 // bottom interface untouched
    return ae * f;
}

// ----------------------------------------------------------------------------------------------
// K2: FCT bounds, P+/P- sums and limiter factors (owned nodes); needs lo on the halo.
//   reference: oce_adv_tra_fct.F90:124-248 (a1-a3), :265-377 (b1), :394-405 (b2)
// The FCT cluster of a node (all nodes of its elements) is the node itself plus its edge
// neighbours, each with the level range of the element(s) they share -- which is exactly the
// scatter range stored with the edge slot -- so one gather serves the bounds and the P sums.
// Output: pm = {R+, R-} per tracer.
// ----------------------------------------------------------------------------------------------
template <int TB, int G>
__global__ void ADV_K2_BOUNDS k_fct_bounds(MeshDev m, Chunk<TB> b, NodePart r, double dt)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw + node_smem_header(m.ell_w));   // [2*TB][blockDim]: tvert_max, tvert_min
    const int L = m.L, nl = m.nl;
    node_cta_prefetch<2 * TB + 5>(m, r, [&](int q) {
        if (q < 2 * TB) return PfArr{(q & 1) ? (const void*)b.dttf_v[q >> 1] : (const void*)b.ttf[q >> 1], (unsigned)L * 8u, 8u, 0};
        switch (q - 2 * TB) {
        case 0: return PfArr{b.lo, (unsigned)L * TB * 8u, TB * 8u, 0};
        case 1: return PfArr{b.adf_v, (unsigned)nl * TB * 8u, TB * 8u, 1};
        case 2: return PfArr{m.areasvol, (unsigned)nl * 8u, 8u, 1};
        case 3: return PfArr{m.hnode, (unsigned)L * 8u, 8u, 0};
        default: return PfArr{m.hnode_new, (unsigned)L * 8u, 8u, 0};
        }
    }, (ADV_PF_OWN & 2) != 0);
    const NodeCta th = node_cta(m, r, smem_raw);
    const int n = th.n, nz = th.nz, nz0 = nz - 1;
    const bool valid = th.valid;
    const unsigned oL = (unsigned)n * L + nz0;
    double tmax[TB], tmin[TB], pp[TB], pn[TB], lo_n[TB];
    double av = 1.0, hnn = 1.0;
    // own-column operands that are needed again only after the gather (vertical update at the end) are parked in
    // shared memory instead of being held in registers across it: [4*TB + 1][blockDim] after the tvert area
    double* park = sm + (size_t)2 * TB * blockDim.x + threadIdx.x;
    const unsigned pst = blockDim.x;
    if (valid) {
        // ---- all global loads of the first batch + the own column ------------------------------
        int4 ent[G];
        double lo_o[G][TB], t_o[G][TB], f[G][TB];
        bool in[G];
#pragma unroll
        for (int j = 0; j < G; ++j) ent[j] = (j < th.deg) ? th.ell[j] : ADV_EMPTY_SLOT;
        const size_t cN = (size_t)n * nl + nz0;
        double tn[TB], vt[TB], vb[TB], dvv[TB];
        ldv<TB>(b.lo + (size_t)oL * TB, lo_n);
#pragma unroll
        for (int t = 0; t < TB; ++t) { tn[t] = __ldg(&b.ttf[t][oL]); dvv[t] = n < m.N ? b.dttf_v[t][oL] : 0.0; }
        ldv<TB>(b.adf_v + cN * TB, vt);
        ldv<TB>(b.adf_v + (cN + 1) * TB, vb);
        av = __ldg(&m.areasvol[cN]); hnn = __ldg(&m.hnode_new[oL]);
        const double hn = __ldg(&m.hnode[oL]);
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
            in[j] = nz >= lo && nz <= hi;
#pragma unroll
            for (int t = 0; t < TB; ++t) { lo_o[j][t] = 0.0; t_o[j][t] = 0.0; f[j][t] = 0.0; }
            if (in[j]) {
                const unsigned oo = (unsigned)ent[j].y * L + nz0;
                ldv<TB>(b.adf_h + ((size_t)(unsigned)ent[j].x * L + nz0) * TB, f[j]);
                ldv<TB>(b.lo + (size_t)oo * TB, lo_o[j]);
#pragma unroll
                for (int t = 0; t < TB; ++t) t_o[j][t] = __ldg(&b.ttf[t][oo]);
            }
        }
        const bool padded = nz < th.pad_lo || nz > th.pad_hi;   // some element of the cluster is dry at nz
        const bool self_in = nz >= th.self_lo && nz <= th.self_hi;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            tmax[t] = padded ? -1.0e3 : -CUDART_INF;        // bignumber, oce_adv_tra_fct.F90:100,159-176
            tmin[t] = padded ? 1.0e3 : CUDART_INF;
            const double hi2 = dmax(lo_n[t], tn[t]), lo2 = dmin(lo_n[t], tn[t]);          // a1 :129-130
            tmax[t] = (self_in && hi2 > tmax[t]) ? hi2 : tmax[t];
            tmin[t] = (self_in && lo2 < tmin[t]) ? lo2 : tmin[t];
            pp[t] = 0.0 + (dmax(0.0, vt[t]) + dmax(0.0, -vb[t]));                          // fct :291
            pn[t] = 0.0 + (dmin(0.0, vt[t]) + dmin(0.0, -vb[t]));                          // fct :292
            park[(4 * t) * pst] = tn[t]; park[(4 * t + 1) * pst] = vt[t]; park[(4 * t + 2) * pst] = vb[t]; park[(4 * t + 3) * pst] = dvv[t];
        }
        park[(4 * TB) * pst] = hn;
        for (int j0 = 0;;) {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if (!in[j]) continue;
                const int smask = (ent[j].z & 0x10000) << 15;                 // bit 31 set: this node is edges(2,e)
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    const double hi2 = dmax(lo_o[j][t], t_o[j][t]), lo2 = dmin(lo_o[j][t], t_o[j][t]);
                    tmax[t] = hi2 > tmax[t] ? hi2 : tmax[t];                  // a2 :166, a3 :209
                    tmin[t] = lo2 < tmin[t] ? lo2 : tmin[t];
                    const double a = flip_sign(f[j][t], smask);               // fct :342,:346 / :360,:364: -f for edges(2,e)
                    // pp += max(0,a); pn += min(0,a): exactly one addend is non-zero, and adding +0.0 changes nothing
                    // (pp >= +0, pn <= +0 and never -0.0: both start as 0.0 + x), so one compare and two predicated adds
                    if (a > 0.0) pp[t] = pp[t] + a; else pn[t] = pn[t] + a;
                }
            }
            j0 += G;
            if (j0 >= th.deg) break;
#pragma unroll
            for (int j = 0; j < G; ++j) ent[j] = (j0 + j < th.deg) ? th.ell[j0 + j] : ADV_EMPTY_SLOT;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
                in[j] = nz >= lo && nz <= hi;
                if (in[j]) {
                    const unsigned oo = (unsigned)ent[j].y * L + nz0;
                    ldv<TB>(b.adf_h + ((size_t)(unsigned)ent[j].x * L + nz0) * TB, f[j]);
                    ldv<TB>(b.lo + (size_t)oo * TB, lo_o[j]);
#pragma unroll
                    for (int t = 0; t < TB; ++t) t_o[j][t] = __ldg(&b.ttf[t][oo]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            sm[(2 * t) * blockDim.x + threadIdx.x] = tmax[t];
            sm[(2 * t + 1) * blockDim.x + threadIdx.x] = tmin[t];
        }
    }
    __syncthreads();
    const double r_av = 1.0 / av, r_hnn = 1.0 / hnn;
    const bool edge_layer = (nz == th.nzmin) || (nz == th.nzmax - 1);   // :233-234, :245-247
    double rp[TB], rm[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { rp[t] = 1.0; rm[t] = 1.0; }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        if (!valid) break;
        double vmax = tmax[t], vmin = tmin[t];
        if (!edge_layer) {                                               // :238-241 (layers nz-1, nz+1 = the threads before / after)
            const double* smax = sm + (2 * t) * blockDim.x + threadIdx.x;
            const double* smin = sm + (2 * t + 1) * blockDim.x + threadIdx.x;
            vmax = dmax(dmax(smax[-1], vmax), smax[1]);
            vmin = dmin(dmin(smin[-1], vmin), smin[1]);
        }
        const double inc_max = vmax - lo_n[t], inc_min = vmin - lo_n[t];
        const double fp = div_rcp(div_rcp(pp[t] * dt, av, r_av), hnn, r_hnn) + 1e-16;   // b2 :399
        const double fm = div_rcp(div_rcp(pn[t] * dt, av, r_av), hnn, r_hnn) - 1e-16;   // :401
        rp[t] = dmin(1.0, inc_max / fp); rm[t] = dmin(1.0, inc_min / fm);
    }
    if (valid) stpm<TB>(b.pm + (size_t)oL * TB * 2, rp, rm);
    // ---- vertical part of the update (oce_adv_tra_fct.F90:425-455 b3 vertical, driver :529-556 U1-U2): the limited
    // vertical fluxes of a column need R+/R- of THIS column only, which this CTA has just computed, so the update of
    // del_ttf_advvert happens here and k_fct_update no longer re-reads adf_v, fct_LO, ttf, hnode, hnode_new.
    __syncthreads();                                  // everybody is done with the tvert exchange area
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        sm[(2 * t) * blockDim.x + threadIdx.x] = rp[t];
        sm[(2 * t + 1) * blockDim.x + threadIdx.x] = rm[t];
    }
    __syncthreads();
    if (!valid || n >= m.N) return;
    const bool above = nz > th.nzmin, below = nz + 1 <= th.nzmax - 1;
    const bool has_below = nz0 + 1 < L;
    const double hn = park[(4 * TB) * pst];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const double tn_t = park[(4 * t) * pst], vt_t = park[(4 * t + 1) * pst], vb_t = park[(4 * t + 2) * pst], dv_t = park[(4 * t + 3) * pst];
        const double* sp = sm + (2 * t) * blockDim.x + threadIdx.x;
        const double* sn = sm + (2 * t + 1) * blockDim.x + threadIdx.x;
        const double pa = above ? sp[-1] : 1.0, ma = above ? sn[-1] : 1.0;
        const double pb = below ? sp[1] : 1.0, mb = below ? sn[1] : 1.0;
        const double fv_top = limit_v(vt_t, nz, th.nzmin, th.nzmax, pa, ma, rp[t], rm[t]);
        const double fv_bot = has_below ? limit_v(vb_t, nz + 1, th.nzmin, th.nzmax, rp[t], rm[t], pb, mb) : 0.0;
        double d = dv_t;
        d = d - tn_t * hn + lo_n[t] * hnn;                            // driver :535
        d = d + div_rcp((fv_top - fv_bot) * dt, av, r_av);            // driver :556
        b.dttf_v[t][oL] = d;
        if (b.dgv[t]) b.dgv[t][oL] = b.dgv[t][oL] + d / hnn;          // ltra_diag, driver :473
    }
}

// ----------------------------------------------------------------------------------------------
// K3: limit the horizontal antidiffusive fluxes and accumulate del_ttf_advhoriz (the vertical part of the update
// needs R+/R- of the own column only and runs at the end of k_fct_bounds).  Owned nodes and, on more than one rank,
// halo nodes: the partial horizontal sums the reference's edge scatter leaves there.
//   reference: oce_adv_tra_fct.F90:425-500 (b3), oce_adv_tra_driver.F90:529-633 (U1-U3)
// ----------------------------------------------------------------------------------------------
template <int TB, int G>
__global__ void ADV_K3_BOUNDS k_fct_update(MeshDev m, Chunk<TB> b, NodePart r, double dt)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int L = m.L, nl = m.nl;
    node_cta_prefetch<TB + 2>(m, r, [&](int q) {
        if (q < TB) return PfArr{(const void*)b.dttf_h[q], (unsigned)L * 8u, 8u, 0};
        if (q == TB) return PfArr{b.pm, (unsigned)L * TB * 16u, TB * 16u, 0};
        return PfArr{m.areasvol, (unsigned)nl * 8u, 8u, 1};
    }, (ADV_PF_OWN & 4) != 0);
    const NodeCta th = node_cta(m, r, smem_raw);
    const int n = th.n, nz = th.nz, nz0 = nz - 1;
    if (!th.valid) return;
    const unsigned oL = (unsigned)n * L + nz0;
    const size_t cN = (size_t)n * nl + nz0;
    // ---- gather metadata first (its latency hides behind the vertical part) ----------------------
    int4 ent[G];
    double f[G][TB], po[G][TB], mo[G][TB];
    bool in[G];
#pragma unroll
    for (int j = 0; j < G; ++j) ent[j] = (j < th.deg) ? th.ell[j] : ADV_EMPTY_SLOT;
    const double av = __ldg(&m.areasvol[cN]);
    double pk[TB], mk[TB], dh[TB];
    ldpm<TB>(b.pm + (size_t)oL * TB * 2, pk, mk);
#pragma unroll
    for (int t = 0; t < TB; ++t) dh[t] = b.dttf_h[t][oL];
    const double r_av = 1.0 / av;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
        in[j] = nz >= lo && nz <= hi;
#pragma unroll
        for (int t = 0; t < TB; ++t) { f[j][t] = 0.0; po[j][t] = 1.0; mo[j][t] = 1.0; }
        if (in[j]) {
            ldv<TB>(b.adf_h + ((size_t)(unsigned)ent[j].x * L + nz0) * TB, f[j]);
            ldpm<TB>(b.pm + ((size_t)(unsigned)ent[j].y * L + nz0) * TB * 2, po[j], mo[j]);
        }
    }
    for (int j0 = 0;;) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (!in[j]) continue;
            // fct :489-494: ae = min(1, R+/R- of edges(1,e), R-/R+ of edges(2,e)) by the sign of the flux; R+/R- are
            // themselves min(1, .) (k_fct_bounds), so min(1, A, B) == min(A, B) bit for bit (also for NaN operands).
            // One (nearly warp-uniform) branch on which end this node is replaces four double selects
            if (!((ent[j].z >> 16) & 1)) {                 // this node is edges(1,e)
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    const double ff = f[j][t];
                    const bool pos = ff >= 0.0;
                    const double A = pos ? pk[t] : mk[t], B = pos ? mo[j][t] : po[j][t];
                    const double ae = dmin(A, B);
                    dh[t] = dh[t] + div_rcp(ae * ff * dt, av, r_av);          // fct :497, driver :607
                }
            } else {                                       // this node is edges(2,e)
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    const double ff = f[j][t];
                    const bool pos = ff >= 0.0;
                    const double A = pos ? po[j][t] : mo[j][t], B = pos ? mk[t] : pk[t];
                    const double ae = dmin(A, B);
                    dh[t] = dh[t] - div_rcp(ae * ff * dt, av, r_av);          // driver :620
                }
            }
        }
        j0 += G;
        if (j0 >= th.deg) break;
#pragma unroll
        for (int j = 0; j < G; ++j) ent[j] = (j0 + j < th.deg) ? th.ell[j0 + j] : ADV_EMPTY_SLOT;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
            in[j] = nz >= lo && nz <= hi;
            if (in[j]) {
                ldv<TB>(b.adf_h + ((size_t)(unsigned)ent[j].x * L + nz0) * TB, f[j]);
                ldpm<TB>(b.pm + ((size_t)(unsigned)ent[j].y * L + nz0) * TB * 2, po[j], mo[j]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        b.dttf_h[t][oL] = dh[t];
        if (b.dgh[t] && n < m.N) b.dgh[t][oL] = b.dgh[t][oL] + dh[t] / __ldg(&m.hnode_new[oL]);   // ltra_diag, driver :472
    }
}

// ----------------------------------------------------------------------------------------------
// tra_adv_lim /= 'FCT': the edge kernel has stored the high-order horizontal flux (Chunk::nolo, o_init_zero=.true.);
// this pass computes the high-order vertical flux and accumulates both tendencies -- an ordered gather of the edge
// slots like k_fct_update, without limiter.  Round 1 recomputed every edge flux at both end nodes from a CSR walk and
// read edge_up_dn_grad twice; now every flux is evaluated once, by the streaming edge kernel.
//   reference: oce_adv_tra_driver.F90:339-379, :387, :551-633 (vertical velocity is `we`, :358)
// ----------------------------------------------------------------------------------------------
template <int VER, int TB>
__global__ void __launch_bounds__(kBlock) k_nofct_update(MeshDev m, Chunk<TB> b, NodeRange r, double dt)
{
    extern __shared__ double sm[];  // [TB][blockDim] vertical flux, then [6 + TB][blockDim] own-column operands
    const int L = m.L, nl = m.nl, nthr = blockDim.x, tid = threadIdx.x;
    const NodeThread tc = node_thread(m, r);
    const int n = tc.n, nz0 = tc.nz0, nz = nz0 + 1;
    const bool owned = n < m.N;
    const bool valid = tc.active && nz >= tc.nzmin && nz <= tc.nzmax - 1;
    const size_t oL = (size_t)n * L + nz0, cN = (size_t)n * nl;
    // own-column operands of the vertical stencils: one coalesced load per thread, parked in shared memory (as k_node_lo)
    double* s_we = sm + (size_t)TB * nthr; double* s_area = s_we + nthr; double* s_Z = s_area + nthr; double* s_zbar = s_Z + nthr;
    double* s_hn = s_zbar + nthr; double* s_hnn = s_hn + nthr; double* s_tab = s_hnn + nthr;   // [TB][nthr]
    if (tc.active && owned && nz >= tc.nzmin && nz <= tc.nzmax - 1) {
        s_we[tid] = __ldg(&m.we[cN + nz0]); s_area[tid] = __ldg(&m.area[cN + nz0]); s_Z[tid] = __ldg(&m.Z3d[oL]);
        s_zbar[tid] = __ldg(&m.zbar3d[cN + nz0]); s_hn[tid] = __ldg(&m.hnode[oL]); s_hnn[tid] = __ldg(&m.hnode_new[oL]);
#pragma unroll
        for (int t = 0; t < TB; ++t) s_tab[t * nthr + tid] = __ldg(&b.ttfAB[t][oL]);
    }
    __syncthreads();
    double fv_top[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) fv_top[t] = 0.0;
    if (tc.active && owned && nz >= tc.nzmin && nz <= tc.nzmax) {
        const int c0 = tid - nz0;                     // first element of this thread's column
        ColV c;
        c.area = s_area + c0; c.Z = s_Z + c0; c.zbar = s_zbar + c0;
        c.hnode = s_hn + c0; c.hnode_new = s_hnn + c0;
        c.nzmin = tc.nzmin; c.nzmax = tc.nzmax; c.dt = dt; c.w = s_we + c0;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            c.ttf = s_tab + t * nthr + c0; c.num_ord = b.pv[t];
            fv_top[t] = ver_flux<VER>(c, nz, 0.0);
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) sm[t * blockDim.x + threadIdx.x] = fv_top[t];
    __syncthreads();
    if (tc.active && owned) {
        stv<TB>(b.adf_v + (cN + nz0) * TB, fv_top);
        if (nz0 == L - 1) {
            double z[TB];
#pragma unroll
            for (int t = 0; t < TB; ++t) z[t] = 0.0;
            stv<TB>(b.adf_v + (cN + L) * TB, z);
        }
    }
    if (!valid) return;
    const double av = m.areasvol[cN + nz0];
    const double r_av = 1.0 / av;
    const bool has_below = nz0 + 1 < L;
    double dh[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) dh[t] = b.dttf_h[t][oL];
    if (owned) {
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double fv_bot = has_below ? sm[t * blockDim.x + threadIdx.x + 1] : 0.0;
            b.dttf_v[t][oL] = b.dttf_v[t][oL] + div_rcp((fv_top[t] - fv_bot) * dt, av, r_av);   // driver :556
        }
    }
    const int4* ell = m.ne_ell + (size_t)n * m.ell_w;
    for (int j = 0; j < tc.deg; ++j) {
        const int4 ent = __ldg(&ell[j]);
        const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
        if (nz < lo || nz > hi) continue;
        double f[TB];
        ldv<TB>(b.adf_h + ((size_t)(unsigned)ent.x * L + nz0) * TB, f);
        const bool second = (ent.z >> 16) & 1;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double term = div_rcp(f[t] * dt, av, r_av);                       // driver :607,:620
            dh[t] = second ? dh[t] - term : dh[t] + term;
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        b.dttf_h[t][oL] = dh[t];
        if (owned && (b.dgh[t] || b.dgv[t])) {                        // ltra_diag without FCT, driver :482-483
            const double hnn = __ldg(&m.hnode_new[oL]);
            if (b.dgh[t]) b.dgh[t][oL] = dh[t] / hnn;
            if (b.dgv[t]) b.dgv[t][oL] = b.dttf_v[t][oL] / hnn;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// ldiag_DVD (oce_adv_tra_driver.F90:263-296, :395-458): the total tracer fluxes through the mid-edge faces and through the
// upper / lower faces of the scalar cells, for the discrete-variance-decay diagnostic.  With FCT the reference stores the
// low-order flux first and adds the LIMITED antidiffusive flux at the end; neither is materialised by the four passes
// (the low-order flux is recomputed from Q, the limiter is applied on the fly), so two optional sweeps rebuild them from
// what is at hand after the last pass: Q, ttf, the unlimited antidiffusive fluxes and R+/R- (owned and halo nodes).
// Without FCT the stored high-order fluxes are the answer.  Every entry of the outputs is written (zeros outside the
// wet range, like the reference's zeroed flux arrays); the layers a boundary edge has above a cavity top, where the
// reference keeps fluxes of cells nobody uses, are written as zero.
// ----------------------------------------------------------------------------------------------
template <int TB> struct DvdPtrs { double* hor[TB]; double* ver[TB]; };

template <int TB>
__global__ void __launch_bounds__(kBlock) k_dvd_hor(MeshDev m, Chunk<TB> b, DvdPtrs<TB> d, int epb, int fct)
{
    const ColThread c = col_thread(m);
    const int L = m.L;
    const int e = blockIdx.x * epb + c.g;
    if (e >= m.E) return;
    const int nz0 = c.nz0, nz = nz0 + 1;
    const int4 em = __ldg(&m.edge_meta[e]);
    const uchar4 lv = __ldg(&m.edge_lev[e]);
    const int lo = lv.z > 0 ? min((int)lv.x, (int)lv.z) : (int)lv.x;      // scatter range, oce_adv_tra_driver.F90:154-156
    const int hi = max((int)lv.y, (int)lv.w);
    const size_t oe = (size_t)e * L + nz0;
    double out[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) out[t] = 0.0;
    if (nz >= lo && nz <= hi) {
        double f[TB];
        ldv<TB>(b.adf_h + oe * TB, f);
        if (fct) {
            const size_t o1 = (size_t)em.x * L + nz0, o2 = (size_t)em.y * L + nz0;
            const double q = __ldg(&m.Q[oe]), aq = fabs(q), qp = q + aq, qm = q - aq;
            double p1[TB], m1[TB], p2[TB], m2[TB];
            ldpm<TB>(b.pm + o1 * TB * 2, p1, m1);
            ldpm<TB>(b.pm + o2 * TB * 2, p2, m2);
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const double flo = hor_lo(__ldg(&b.ttf[t][o1]), __ldg(&b.ttf[t][o2]), qp, qm);   // driver :115, :273
                double ae = 1.0;                                                                 // fct :489-494
                if (f[t] >= 0.0) { ae = dmin(ae, p1[t]); ae = dmin(ae, m2[t]); }
                else { ae = dmin(ae, m1[t]); ae = dmin(ae, p2[t]); }
                out[t] = flo + ae * f[t];                                                        // driver :404
            }
        } else {
#pragma unroll
            for (int t = 0; t < TB; ++t) out[t] = f[t];                                          // driver :437
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) if (d.hor[t]) d.hor[t][oe] = out[t];
}

template <int TB>
__global__ void __launch_bounds__(kBlock) k_dvd_ver(MeshDev m, Chunk<TB> b, DvdPtrs<TB> d, NodeRange r, int fct)
{
    const NodeThread th = node_thread(m, r);
    if (!th.active) return;
    const int L = m.L, nl = m.nl, n = th.n, nzmin = th.nzmin, nzmax = th.nzmax;
    const size_t cN = (size_t)n * nl;
    ColV c;
    c.w = m.we + cN; c.area = m.area + cN; c.nzmin = nzmin; c.nzmax = nzmax; c.dt = 0.0; c.num_ord = 0.0;
    c.Z = nullptr; c.zbar = nullptr; c.hnode = nullptr; c.hnode_new = nullptr;
    // thread nz0 owns interface nz0 + 1; the thread of the last layer also owns interface nl
    for (int k = th.nz0 + 1; k <= (th.nz0 == L - 1 ? nl : th.nz0 + 1); ++k) {
        double out[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) out[t] = 0.0;
        if (k >= nzmin && k <= nzmax) {
            double f[TB];
            ldv<TB>(b.adf_v + (cN + k - 1) * TB, f);
            if (fct) {
                double pa[TB], ma[TB], pk[TB], mk[TB];
#pragma unroll
                for (int t = 0; t < TB; ++t) { pa[t] = ma[t] = pk[t] = mk[t] = 1.0; }
                if (k - 1 >= nzmin) ldpm<TB>(b.pm + ((size_t)n * L + k - 2) * TB * 2, pa, ma);      // R+/R- of layer k-1
                if (k <= nzmax - 1) ldpm<TB>(b.pm + ((size_t)n * L + k - 1) * TB * 2, pk, mk);      // of layer k
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    c.ttf = b.ttf[t] + (size_t)n * L;
                    const double flo = ver_upw1(c, k, 0.0);                                          // driver :235, :288
                    out[t] = flo + limit_v(f[t], k, nzmin, nzmax, pa[t], ma[t], pk[t], mk[t]);       // driver :418
                }
            } else {
#pragma unroll
                for (int t = 0; t < TB; ++t) out[t] = f[t];                                          // driver :449
            }
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) if (d.ver[t]) d.ver[t][cN + k - 1] = out[t];
    }
}

// ----------------------------------------------------------------------------------------------
// halo pack (replaces the MPI_TYPE_INDEXED send types, gen_modules_partitioning.F90:462-473):
// one CTA per send column, out[i*nlev + k] = field[slist[i]*nlev + k]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_pack_halo(const double* __restrict__ field, const int* __restrict__ slist,
                                                      int nlev, double* __restrict__ out)
{
    const size_t src = (size_t)slist[blockIdx.x] * nlev, dst = (size_t)blockIdx.x * nlev;
    for (int k = threadIdx.x; k < nlev; k += blockDim.x) out[dst + k] = field[src + k];
}

// halo unpack (the MPI_TYPE_INDEXED receive types of the element halo, gen_modules_partitioning.F90:217-410:
// com_elem2D_full%rlist is not a contiguous tail): field[rlist[i]*nlev + k] = in[i*nlev + k]
__global__ void __launch_bounds__(kBlock) k_unpack_halo(const double* __restrict__ in, const int* __restrict__ rlist,
                                                        int nlev, double* __restrict__ field)
{
    const size_t dst = (size_t)rlist[blockIdx.x] * nlev, src = (size_t)blockIdx.x * nlev;
    for (int k = threadIdx.x; k < nlev; k += blockDim.x) field[dst + k] = in[src + k];
}

// dwarf epilogue (fesom.F90:105-125 with del_ttf reset per step): values += (dh+dv)/hnode_new
__global__ void __launch_bounds__(kBlock) k_update_values(MeshDev m, NodeRange r, double* __restrict__ values,
                                                          const double* __restrict__ dh, const double* __restrict__ dv)
{
    const NodeThread th = node_thread(m, r);
    const int nz = th.nz0 + 1;
    if (!th.active || nz < th.nzmin || nz > th.nzmax - 1) return;
    const size_t idx = (size_t)th.n * m.L + th.nz0;
    const double del = 0.0 + dh[idx] + dv[idx];
    values[idx] = values[idx] + del / m.hnode_new[idx];
}

// init_tracers_AB (src/oce_tracer_mod.F90:28-34, :45-54, :97-122): one thread per (layer, node) of ALL
// Nh columns; evaluation order of the Fortran expressions, no contraction.
template <int ORDER>
__global__ void __launch_bounds__(256) k_init_tracers_AB(size_t n, double eps, const double* __restrict__ values,
                                                         double* __restrict__ vold, double* __restrict__ vab,
                                                         double* __restrict__ d0, double* __restrict__ d1, double* __restrict__ d2)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = values[i];
    if (ORDER == 2) {
        const double o1 = vold[i];
        vab[i] = -(0.5 + eps) * o1 + (1.5 + eps) * v;
        vold[i] = v;
    } else {
        const double o1 = vold[2 * i], o2 = vold[2 * i + 1];
        const double t = 5.0 * o2 - 16.0 * o1 + 23.0 * v;
        vab[i] = t / 12.0;
        vold[2 * i + 1] = o1; vold[2 * i] = v;
    }
    if (d0) d0[i] = 0.0;
    if (d1) d1[i] = 0.0;
    if (d2) d2[i] = 0.0;
}

// ----------------------------------------------------------------------------------------------
// The producer of edge_up_dn_grad, first version: correct and coalesced,
// not yet fused into the edge kernel.
// ----------------------------------------------------------------------------------------------
struct GradMeshDev {
    int n_elem, n_nie, ld;
    const int* nie;          // (ld, n_nie) nod_in_elem2D, 1-based
    const int* nie_num;      // (n_nie)
    const int* nlevels;      // (n_elem)
    const int* ulevels;      // (n_elem)
    const int* up_dn_tri;    // (2, E) 1-based, 0 = none
    const int* nmin;         // (Nh) nlevels_nod2D_min
    const int* umax;         // (Nh) ulevels_nod2D_max
    const int* elem_nodes;   // (3, T) 1-based
    const double* gsca;      // (6, T)
    const double* earea;     // (n_elem)
};

// tracer_gradient_elements, src/oce_tracer_mod.F90:171-180: one thread per (element, layer), NT tracers per launch
// (the element's node ids, level range and gradient_sca row are loaded once for all of them)
template <int NT> struct PtrPack { const double* in[NT]; double* out[NT]; };
template <int NT>
__global__ void __launch_bounds__(kBlock) k_tracer_gradient_elements(MeshDev m, GradMeshDev g, int cpb, PtrPack<NT> p)
{
    const ColThread c = col_thread(m);
    const int el = blockIdx.x * cpb + c.g;
    if (c.g >= cpb || el >= m.T) return;
    const int nz = c.nz0 + 1;
    if (nz < __ldg(&g.ulevels[el]) || nz > __ldg(&g.nlevels[el]) - 1) return;
    const int n1 = __ldg(&g.elem_nodes[3 * el]) - 1, n2 = __ldg(&g.elem_nodes[3 * el + 1]) - 1, n3 = __ldg(&g.elem_nodes[3 * el + 2]) - 1;
    const double* gs = g.gsca + (size_t)el * 6;
    const double g0 = __ldg(&gs[0]), g1 = __ldg(&gs[1]), g2 = __ldg(&gs[2]), g3 = __ldg(&gs[3]), g4 = __ldg(&gs[4]), g5 = __ldg(&gs[5]);
    double t1[NT], t2[NT], t3[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        t1[t] = __ldg(&p.in[t][(size_t)n1 * m.L + c.nz0]); t2[t] = __ldg(&p.in[t][(size_t)n2 * m.L + c.nz0]); t3[t] = __ldg(&p.in[t][(size_t)n3 * m.L + c.nz0]);
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const double tx = g0 * t1[t] + g1 * t2[t] + g2 * t3[t];      // sum() left to right
        const double ty = g3 * t1[t] + g4 * t2[t] + g5 * t3[t];
        reinterpret_cast<double2*>(p.out[t])[(size_t)el * m.L + c.nz0] = make_double2(tx, ty);
    }
}

// area-weighted mean of the element gradients around `node` at layer nz, slots in their order
// (src/oce_muscl_adv.F90:391-406), for NT tracers at once (the element walk is shared)
template <int NT>
__device__ __forceinline__ void node_mean_gradient(const MeshDev& m, const GradMeshDev& g, int node, int nz,
                                                   const double* const (&tr_xy)[NT], double2 (&out)[NT])
{
    double tvol = 0.0, tx[NT], ty[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) { tx[t] = 0.0; ty[t] = 0.0; }
    const int num = __ldg(&g.nie_num[node]);
    const int* row = g.nie + (size_t)node * g.ld;
    for (int k = 0; k < num; ++k) {
        const int el = __ldg(&row[k]) - 1;
        if (__ldg(&g.nlevels[el]) - 1 < nz || nz < __ldg(&g.ulevels[el])) continue;
        const double a = __ldg(&g.earea[el]);
        tvol = tvol + a;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(tr_xy[t]) + (size_t)el * m.L + (nz - 1));
            tx[t] = tx[t] + v.x * a;
            ty[t] = ty[t] + v.y * a;
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) out[t] = make_double2(tx[t] / tvol, ty[t] / tvol);
}
__device__ __forceinline__ double2 node_mean_gradient(const MeshDev& m, const GradMeshDev& g, int node, int nz,
                                                       const double* __restrict__ tr_xy)
{
    const double* const p[1] = {tr_xy};
    double2 o[1];
    node_mean_gradient<1>(m, g, node, nz, p, o);
    return o[0];
}

// fill_up_dn_grad, src/oce_muscl_adv.F90:378-522: one thread per (edge, layer); writes exactly the entries
// the reference writes
__global__ void __launch_bounds__(kBlock) k_fill_up_dn_grad(MeshDev m, GradMeshDev g, int cpb, const double* __restrict__ tr_xy,
                                                            double* __restrict__ grad)
{
    const ColThread c = col_thread(m);
    const int e = blockIdx.x * cpb + c.g;
    if (c.g >= cpb || e >= m.E) return;
    const int nz = c.nz0 + 1;
    const int4 em = __ldg(&m.edge_meta[e]);                       // {edges(1,e), edges(2,e), ..} 0-based
    const int t1 = __ldg(&g.up_dn_tri[2 * e]), t2 = __ldg(&g.up_dn_tri[2 * e + 1]);
    const uchar4 l1 = __ldg(&m.node_lev[em.x]), l2 = __ldg(&m.node_lev[em.y]);   // {ulevels_nod2D, nlevels_nod2D, ..}
    const bool both = t1 != 0 && t2 != 0;
    int nzmin = 0, nzmax = 0;
    if (both) {
        nzmin = max(__ldg(&g.umax[em.x]), __ldg(&g.umax[em.y]));
        nzmax = min(__ldg(&g.nmin[em.x]), __ldg(&g.nmin[em.y]));
    }
    const bool shared = both && nz >= nzmin && nz <= nzmax - 1;
    double* out = grad + ((size_t)e * m.L + c.nz0) * 4;
    const double2* txy = reinterpret_cast<const double2*>(tr_xy);
    if (shared) {                                                 // :435-440
        const double2 up = __ldg(&txy[(size_t)(t1 - 1) * m.L + c.nz0]), dn = __ldg(&txy[(size_t)(t2 - 1) * m.L + c.nz0]);
        out[0] = up.x; out[1] = dn.x; out[2] = up.y; out[3] = dn.y;
        return;
    }
    // loop bounds of :388,:411 (above the shared range), :445,:467 (below it) and :493,:511 (boundary edge)
    const int u1 = l1.x, n1 = (int)l1.y - 1, u2 = l2.x, n2 = (int)l2.y - 1;
    const bool w1 = both ? ((nz >= u1 && nz <= nzmin - 1) || (nz >= nzmax && nz <= n1)) : (nz >= u1 && nz <= n1);
    const bool w2 = both ? ((nz >= u2 && nz <= nzmin - 1) || (nz >= nzmax && nz <= n2)) : (nz >= u2 && nz <= n2);
    if (w1) { const double2 v = node_mean_gradient(m, g, em.x, nz, tr_xy); out[0] = v.x; out[2] = v.y; }
    if (w2) { const double2 v = node_mean_gradient(m, g, em.y, nz, tr_xy); out[1] = v.x; out[3] = v.y; }
}

// The node part of fill_up_dn_grad (src/oce_muscl_adv.F90:391-406,:414-429 and the below-bottom / boundary-edge
// twins): the area-weighted mean of the element gradients around a node depends on (node, layer) only, so it is
// evaluated ONCE per node instead of once per incident edge; the fused edge kernel (k_edge_flux_b<.., GS = 1>)
// picks it up on the layers where the reference uses it.  One thread per (node, layer) of all local nodes.
template <int NT>
__global__ void __launch_bounds__(kBlock) k_node_mean_grad(MeshDev m, GradMeshDev g, int cpb, PtrPack<NT> p)
{
    const ColThread c = col_thread(m);
    const int n = blockIdx.x * cpb + c.g;
    if (c.g >= cpb || n >= g.n_nie) return;
    const int nz = c.nz0 + 1;
    const uchar4 lv = __ldg(&m.node_lev[n]);
    double2 v[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) v[t] = make_double2(0.0, 0.0);
    if (nz >= (int)lv.x && nz <= (int)lv.y - 1) node_mean_gradient<NT>(m, g, n, nz, p.in, v);
#pragma unroll
    for (int t = 0; t < NT; ++t) reinterpret_cast<double2*>(p.out[t])[(size_t)n * m.L + c.nz0] = v[t];
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
