// fesom2_b200/csrc/adv_lean.cuh -- multi-group register-gather node kernels with an instruction diet.
//
// ncu of the first-generation kernels (profiles/r1f, r1h): 870-950 warp instructions per
// (32 node-layers x 2 tracers), only ~15 % of them FP64 arithmetic; 20 % are FSEL pairs of
// double-precision selects, 19 % integer address arithmetic; and every thread serialises eight
// memory waits (record -> adjacency row -> one wait per gather slot).  This file keeps the
// thread = (node, layer) mapping and the exact order of operations, and removes the overheads:
//   * a CTA walks `ng` groups of `cpb` nodes; node records and ELL rows of the whole CTA are
//     loaded to shared memory once, with the column offsets (edge * L, node * L) premultiplied;
//   * gather slots are loaded in batches of G (two batches cover degree 6): 2 waits, not 8;
//   * the "which end of the edge am I" selects collapse into one (nearly warp-uniform) branch;
//   * 1/areasvol comes from a precomputed array (div_rcp stays bit-identical to IEEE division).
#pragma once
#include "adv_kernels.cuh"

namespace adv {

#ifndef ADV_K3L_REGS
#define ADV_K3L_REGS 72
#endif

__host__ __device__ constexpr size_t lean_align16(size_t x) { return (x + 15) & ~(size_t)15; }
// shared metadata: int4 ell[cn][W] = {edge * L, other node * L, lo | hi << 8 | flags << 16, 0},
//                  int4 nd[cn] = {n (or -1), n * L, node_rec.x, node_rec.y}
__host__ __device__ inline size_t lean_meta_bytes(int cn, int W) { return (size_t)cn * W * 16 + (size_t)cn * 16; }

struct LeanMeta { const int4* ell; const int4* nd; };
__device__ __forceinline__ LeanMeta lean_meta_load(unsigned char* base, const MeshDev& m, const NodeRange& r, int cn)
{
    int4* s_ell = reinterpret_cast<int4*>(base);
    int4* s_nd = s_ell + (size_t)cn * m.ell_w;
    const int i0 = blockIdx.x * cn;
    for (int i = threadIdx.x; i < cn; i += blockDim.x) {
        int4 nd = make_int4(-1, 0, 1, 0);             // nzmin = 1 > nzmax = 0: no valid layer, degree 0
        if (i0 + i < r.count) {
            const int n = r.list ? __ldg(&r.list[r.begin + i0 + i]) : r.begin + i0 + i;
            const uint2 rec = __ldg(&m.node_rec[n]);
            nd = make_int4(n, n * m.L, (int)rec.x, (int)rec.y);
        }
        s_nd[i] = nd;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < cn * m.ell_w; k += blockDim.x) {
        const int i = k / m.ell_w, n = s_nd[i].x;
        int4 e = ADV_EMPTY_SLOT;
        if (n >= 0) {
            e = __ldg(&m.ne_ell[(size_t)n * m.ell_w + (k - i * m.ell_w)]);
            e.x *= m.L; e.y *= m.L;
        }
        s_ell[k] = e;
    }
    __syncthreads();
    LeanMeta s; s.ell = s_ell; s.nd = s_nd;
    return s;
}

// ----------------------------------------------------------------------------------------------
// K3 (lean): formulas and citations as k_fct_update.
// ----------------------------------------------------------------------------------------------
template <int TB, int G>
__global__ void __maxnreg__(ADV_K3L_REGS) k_fct_update_l(MeshDev m, Chunk<TB> b, NodeRange r, int ng, double dt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = m.L, W = m.ell_w;
    const int cn = r.cpb * ng;
    const LeanMeta ms = lean_meta_load(smem_raw, m, r, cn);
    const ColThread c = col_thread(m);
    const int nz0 = c.nz0, nz = nz0 + 1;
    const int left = r.count - blockIdx.x * cn;
    const int ngrp = min(ng, (left + r.cpb - 1) / r.cpb);
    const double2* pm2 = reinterpret_cast<const double2*>(b.pm);
    for (int i = 0; i < ngrp; ++i) {
        const int li = i * r.cpb + c.g;
        const int4 nd = ms.nd[li];
        const int n = nd.x;
        const int nzmin = nd.z & 0xff, nzmax = (nd.z >> 8) & 0xff, deg = (nd.w >> 16) & 0xff;
        if (nz < nzmin || nz > nzmax - 1) continue;
        const unsigned oL = (unsigned)nd.y + nz0;
        const size_t cN = (size_t)oL + n;                       // n * nl + nz0, nl = L + 1
        const int4* ell = ms.ell + li * W;
        int4 ent[G];
        double f[G][TB], po[G][TB], mo[G][TB];
        bool in[G];
        // ---- first gather batch, own column, vertical operands: all loads before the first use ------
#pragma unroll
        for (int j = 0; j < G; ++j) {
            ent[j] = ell[j];
            const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
            in[j] = j < deg && nz >= lo && nz <= hi;
            if (in[j]) {
                ldv<TB>(b.adf_h + (size_t)((unsigned)ent[j].x + nz0) * TB, f[j]);
                ldpm<TB>(b.pm + (size_t)((unsigned)ent[j].y + nz0) * TB * 2, po[j], mo[j]);
            }
        }
        double pk[TB], mk[TB], dh[TB];
        ldpm<TB>(b.pm + (size_t)oL * TB * 2, pk, mk);
#pragma unroll
        for (int t = 0; t < TB; ++t) dh[t] = b.dttf_h[t][oL];
        const double av = __ldg(&m.areasvol[cN]), r_av = __ldg(&m.r_areasvol[cN]);
        if (n < m.N) {
            double vt[TB], vb[TB], lo_n[TB];
            const bool above = nz > nzmin, below = nz + 1 <= nzmax - 1, has_below = nz0 + 1 < L;
            ldv<TB>(b.adf_v + cN * TB, vt);
            ldv<TB>(b.adf_v + (cN + 1) * TB, vb);
            ldv<TB>(b.lo + (size_t)oL * TB, lo_n);
            const double hn = __ldg(&m.hnode[oL]), hnn = __ldg(&m.hnode_new[oL]);
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                double pa = 1.0, ma = 1.0, pb = 1.0, mb = 1.0;
                if (above) { const double2 v = __ldg(pm2 + (size_t)(oL - 1) * TB + t); pa = v.x; ma = v.y; }
                if (below) { const double2 v = __ldg(pm2 + (size_t)(oL + 1) * TB + t); pb = v.x; mb = v.y; }
                const double fv_top = limit_v(vt[t], nz, nzmin, nzmax, pa, ma, pk[t], mk[t]);
                const double fv_bot = has_below ? limit_v(vb[t], nz + 1, nzmin, nzmax, pk[t], mk[t], pb, mb) : 0.0;
                double d = b.dttf_v[t][oL];
                d = d - __ldg(&b.ttf[t][oL]) * hn + lo_n[t] * hnn;                // driver :535
                d = d + div_rcp((fv_top - fv_bot) * dt, av, r_av);                // driver :556
                b.dttf_v[t][oL] = d;
            }
        }
        for (int j0 = 0;;) {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if (!in[j]) continue;
                if (!((ent[j].z >> 16) & 1)) {                 // this node is edges(1,e)
#pragma unroll
                    for (int t = 0; t < TB; ++t) {
                        const double ff = f[j][t];
                        const bool pos = ff >= 0.0;
                        const double A = pos ? pk[t] : mk[t], B = pos ? mo[j][t] : po[j][t];   // fct :489-494
                        const double ae = dmin(dmin(1.0, A), B);
                        dh[t] = dh[t] + div_rcp(ae * ff * dt, av, r_av);          // fct :497, driver :607
                    }
                } else {                                       // this node is edges(2,e)
#pragma unroll
                    for (int t = 0; t < TB; ++t) {
                        const double ff = f[j][t];
                        const bool pos = ff >= 0.0;
                        const double A = pos ? po[j][t] : mo[j][t], B = pos ? mk[t] : pk[t];
                        const double ae = dmin(dmin(1.0, A), B);
                        dh[t] = dh[t] - div_rcp(ae * ff * dt, av, r_av);          // driver :620
                    }
                }
            }
            j0 += G;
            if (j0 >= deg) break;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                in[j] = false;
                if (j0 + j < deg) {
                    ent[j] = ell[j0 + j];
                    const int lo = ent[j].z & 0xff, hi = (ent[j].z >> 8) & 0xff;
                    in[j] = nz >= lo && nz <= hi;
                    if (in[j]) {
                        ldv<TB>(b.adf_h + (size_t)((unsigned)ent[j].x + nz0) * TB, f[j]);
                        ldpm<TB>(b.pm + (size_t)((unsigned)ent[j].y + nz0) * TB * 2, po[j], mo[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) b.dttf_h[t][oL] = dh[t];
    }
}

}  // namespace adv
