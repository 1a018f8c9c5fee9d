// fesom2_b200/csrc/adv_slot.cuh -- slot-parallel versions of the node kernels.
//
// The register-gather kernels (adv_kernels.cuh) give one thread a whole (node, layer): it walks the
// node's edge slots one after the other, so its life is a chain of dependent memory waits
// (record -> adjacency row -> slot 0 operands -> slot 1 operands ...) and the registers needed to
// batch the slots cap the occupancy (ncu: 11-16 stall cycles per issued instruction on the long
// scoreboard at 26-33 resident warps).  Here a CTA owns ONE node column and runs S x L threads:
// thread (s, layer) evaluates the contribution of slot s only -- one short, independent load ->
// compute -> shared-memory store -- so all slots of a node are in flight at once, every thread
// needs few registers and 40-56 warps stay resident.  The ORDER of the reference's scatter sums
// (ascending edge id) is kept by a second phase in which one thread per (layer,
// tracer, part) adds the finished terms in slot order; the terms themselves are computed exactly
// as before, so results stay bit-identical.
#pragma once
#include "adv_kernels.cuh"

namespace adv {

constexpr int kSlotBlock = 512;
#ifndef ADV_K3T_MINB
#define ADV_K3T_MINB 3
#endif

// thread -> (slot lane s, layer): blockDim.x == S * L
struct SlotThread { int s, nz0; };
__device__ __forceinline__ SlotThread slot_thread(const MeshDev& m)
{
    SlotThread c;
    c.s = (int)((threadIdx.x * m.div_magic) >> 20);
    c.nz0 = (int)threadIdx.x - c.s * m.L;
    return c;
}

// ----------------------------------------------------------------------------------------------
// K3 (slot-parallel): limit the antidiffusive fluxes and accumulate the tendencies; formulas and
// citations as k_fct_update.
//   phase 1, thread (s, layer): for slots j = s, s+S, ...: the limited, signed contribution
//            +-(ae * adf_h * dt) / areasvol of edge slot j           (fct :489-497, driver :607,:620)
//   phase 2, thread rows 0..TB-1:   del_ttf_advhoriz(t) += terms in ascending slot order
//            thread rows TB..2TB-1: del_ttf_advvert(t)   (fct :425-455, driver :535,:556; owned nodes)
// Dynamic shared memory: double term[W][TB][L].
// ----------------------------------------------------------------------------------------------
template <int TB>
__global__ void __launch_bounds__(kSlotBlock, ADV_K3T_MINB) k_fct_update_t(MeshDev m, Chunk<TB> b, NodeRange r, int S, double dt)
{
    extern __shared__ double s_term[];
    const int L = m.L, nl = m.nl, W = m.ell_w;
    const SlotThread c = slot_thread(m);
    const int nz0 = c.nz0, nz = nz0 + 1;
    const int n = r.list ? __ldg(&r.list[r.begin + blockIdx.x]) : r.begin + blockIdx.x;
    const uint2 rec = __ldg(&m.node_rec[n]);
    const int nzmin = rec.x & 0xff, nzmax = (rec.x >> 8) & 0xff, deg = (rec.y >> 16) & 0xff;
    const bool valid = nz >= nzmin && nz <= nzmax - 1;
    const unsigned oL = (unsigned)n * L + nz0;
    const size_t cN = (size_t)n * nl + nz0;
    const int4* ell = m.ne_ell + (size_t)n * W;
    double pk[TB], mk[TB];
    double av = 1.0, r_av = 1.0;
#pragma unroll
    for (int t = 0; t < TB; ++t) { pk[t] = 1.0; mk[t] = 1.0; }
    if (valid) {
        ldpm<TB>(b.pm + (size_t)oL * TB * 2, pk, mk);
        av = __ldg(&m.areasvol[cN]); r_av = __ldg(&m.r_areasvol[cN]);
        for (int j = c.s; j < deg; j += S) {
            const int4 ent = __ldg(&ell[j]);
            const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
            if (nz < lo || nz > hi) continue;
            const bool second = (ent.z >> 16) & 1;
            double f[TB], po[TB], mo[TB];
            ldv<TB>(b.adf_h + ((size_t)(unsigned)ent.x * L + nz0) * TB, f);
            ldpm<TB>(b.pm + ((size_t)(unsigned)ent.y * L + nz0) * TB * 2, po, mo);
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const double ff = f[t];
                const double p1 = second ? po[t] : pk[t], m1 = second ? mo[t] : mk[t];   // factors at edges(1,e)
                const double p2 = second ? pk[t] : po[t], m2 = second ? mk[t] : mo[t];   // factors at edges(2,e)
                double ae = 1.0;
                if (ff >= 0.0) { ae = dmin(ae, p1); ae = dmin(ae, m2); }      // fct :489-491
                else { ae = dmin(ae, m1); ae = dmin(ae, p2); }                // :493-494
                const double term = div_rcp(ae * ff * dt, av, r_av);          // fct :497, driver :607,:620
                s_term[(j * TB + t) * L + nz0] = second ? -term : term;       // x - term == x + (-term) exactly
            }
        }
    }
    __syncthreads();
    if (!valid) return;
    for (int role = c.s; role < 2 * TB; role += S) {
        const int t = role % TB;
        if (role < TB) {
            double dh = b.dttf_h[t][oL];
            for (int j = 0; j < deg; ++j) {
                const int4 ent = __ldg(&ell[j]);
                const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
                if (nz >= lo && nz <= hi) dh = dh + s_term[(j * TB + t) * L + nz0];
            }
            b.dttf_h[t][oL] = dh;
        } else if (n < m.N) {
            const bool above = nz > nzmin, below = nz + 1 <= nzmax - 1, has_below = nz0 + 1 < L;
            double pa = 1.0, ma = 1.0, pb = 1.0, mb = 1.0, vb = 0.0;
            const double vt = __ldg(&b.adf_v[cN * TB + t]);
            if (has_below) vb = __ldg(&b.adf_v[(cN + 1) * TB + t]);
            const double lo_n = __ldg(&b.lo[(size_t)oL * TB + t]);
            const double tn = __ldg(&b.ttf[t][oL]);
            const double hn = __ldg(&m.hnode[oL]), hnn = __ldg(&m.hnode_new[oL]);
            double d = b.dttf_v[t][oL];
            if (above) { const double2 v = __ldg(reinterpret_cast<const double2*>(b.pm) + (size_t)(oL - 1) * TB + t); pa = v.x; ma = v.y; }
            if (below) { const double2 v = __ldg(reinterpret_cast<const double2*>(b.pm) + (size_t)(oL + 1) * TB + t); pb = v.x; mb = v.y; }
            const double fv_top = limit_v(vt, nz, nzmin, nzmax, pa, ma, pk[t], mk[t]);
            const double fv_bot = has_below ? limit_v(vb, nz + 1, nzmin, nzmax, pk[t], mk[t], pb, mb) : 0.0;
            d = d - tn * hn + lo_n * hnn;                                     // driver :535
            d = d + div_rcp((fv_top - fv_bot) * dt, av, r_av);                // driver :556
            b.dttf_v[t][oL] = d;
        }
    }
}

}  // namespace adv
