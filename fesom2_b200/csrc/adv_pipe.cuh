// fesom2_b200/csrc/adv_pipe.cuh -- multi-group, asynchronously fed versions of the four FCT kernels.
//
// Same arithmetic, same order of operations and same reference citations as adv_kernels.cuh (the
// per-(column, layer) formulas are shared); what changes is how operands reach the thread:
//
//   * a CTA owns `ng` consecutive groups of `cpb` columns.  All per-column metadata of the CTA
//     (edge records, node records, ELL adjacency rows) is loaded ONCE into shared memory, so the
//     dependent chain  index -> metadata -> operands  is paid once per CTA, not once per column;
//   * every operand of a (column, layer) thread -- its own column's words and all gather slots --
//     is fetched with per-thread asynchronous copies (cp.async, LDGSTS) into thread-private
//     shared-memory cells: no registers are held while the data is in flight, every load of a
//     group is issued before the first wait, and with D >= 2 stages the loads of the next group(s)
//     are in flight while the current group is computed.  Because a thread mostly reads the
//     cells it filled itself, the edge kernel needs no CTA barrier at all; the node kernels need
//     the one barrier their vertical neighbour accesses need anyway.
//
// Measured against the register-gather kernels in profiles/ (r1h).
#pragma once
#include "adv_kernels.cuh"

namespace adv {

#ifndef ADV_E1P_MINB
#define ADV_E1P_MINB 4      // 4 CTAs of 7 warps per SM: measured best for the bulk edge kernel (profiles/r1_tuning.md)
#endif
#ifndef ADV_NDP_MINB
#define ADV_NDP_MINB 2
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// 8-byte copies may only use .ca; 16-byte copies of streamed data bypass L1 (.cg)
__device__ __forceinline__ void cpa8(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16_stream(unsigned dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// TB doubles (8 or 16 bytes)
template <int TB> __device__ __forceinline__ void cpa_vec(unsigned dst, const double* src)
{
    if (TB == 2) cpa16(dst, src); else cpa8(dst, src);
}

__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ----------------------------------------------------------------------------------------------
// E1 (pipelined): antidiffusive horizontal flux per (edge, layer); formulas and citations as
// k_edge_flux.  Per-thread cells of one stage, [cell][thread]:
//   16-byte cells: g12[t], g34[t] (gradient schemes), uv1, uv2 (QMODE 0)
//    8-byte cells: t1[t], t2[t], a1[t], a2[t], then he1, he2 (QMODE 0) or q (QMODE 1)
// ----------------------------------------------------------------------------------------------
template <int HOR, int TB, int QMODE>
struct E1Cells {
    static constexpr int n16 = (HOR != HOR_UPW1 ? 2 * TB : 0) + (QMODE == 0 ? 2 : 0);
    static constexpr int n8 = 4 * TB + (QMODE == 0 ? 2 : 1);
    static constexpr int bytes = 16 * n16 + 8 * n8;     // per thread and stage
    static constexpr int c_g12 = 0, c_g34 = TB, c_uv = (HOR != HOR_UPW1 ? 2 * TB : 0);   // 16-byte cell ids
    static constexpr int c_t1 = 0, c_t2 = TB, c_a1 = 2 * TB, c_a2 = 3 * TB, c_x = 4 * TB;  // 8-byte cell ids
};
// metadata block: int4 em[nge], double2 cr[2 nge], double2 ec[nge], unsigned lv[nge], int2 nb[nge]
__host__ __device__ inline size_t e1p_meta_bytes(int nge) { return align16((size_t)nge * (16 + 32 + 16 + 8 + 4)); }
template <int HOR, int TB, int QMODE>
__host__ __device__ inline size_t e1p_smem_bytes(int nge, int nthr, int D)
{
    return e1p_meta_bytes(nge) + (size_t)D * align16((size_t)E1Cells<HOR, TB, QMODE>::bytes * nthr);
}

template <int HOR, int TB, int QMODE, int D>
__global__ void __launch_bounds__(kBlock, ADV_E1P_MINB) k_edge_flux_p(MeshDev m, Chunk<TB> b, int epb, int ng)
{
    using C = E1Cells<HOR, TB, QMODE>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = m.L, nthr = blockDim.x, tid = threadIdx.x;
    const int nge = epb * ng;
    const int e0 = blockIdx.x * nge;
    int4* s_em = reinterpret_cast<int4*>(smem_raw);
    double2* s_cr = reinterpret_cast<double2*>(s_em + nge);
    double2* s_ec = s_cr + 2 * nge;
    int2* s_nb = reinterpret_cast<int2*>(s_ec + nge);
    unsigned* s_lv = reinterpret_cast<unsigned*>(s_nb + nge);
    unsigned char* s_stage = smem_raw + e1p_meta_bytes(nge);
    // ---- per-CTA metadata, once ------------------------------------------------------------------
    for (int i = tid; i < nge; i += nthr) {
        const int e = e0 + i;
        if (e < m.E) {
            const int4 em = __ldg(&m.edge_meta[e]);      // {edges(1,e), edges(2,e), el1, el2} 0-based, el2 = -1: none
            s_em[i] = em;
            s_lv[i] = __ldg(reinterpret_cast<const unsigned*>(m.edge_lev) + e);
            const double2* cr = reinterpret_cast<const double2*>(&m.edge_cross[e]);
            s_cr[2 * i] = __ldg(cr); s_cr[2 * i + 1] = __ldg(cr + 1);
            if (HOR != HOR_UPW1) s_ec[i] = __ldg(&m.edge_c[e]);
            if (HOR == HOR_MUSCL) s_nb[i] = make_int2(__ldg(&m.nboundary_lay[em.x]), __ldg(&m.nboundary_lay[em.y]));
        }
    }
    __syncthreads();
    const ColThread c = col_thread(m);
    const int nz0 = c.nz0, nz = nz0 + 1;
    const int left = m.E - e0;                                   // edges from e0 on
    const int ngrp = min(ng, (left + epb - 1) / epb);
    const unsigned stage_bytes = (unsigned)align16((size_t)C::bytes * nthr);
    const unsigned sb0 = smem_u32(s_stage);
    const unsigned o16 = (unsigned)tid * 16u, o8 = (unsigned)C::n16 * 16u * nthr + (unsigned)tid * 8u;
    const unsigned st16 = 16u * nthr, st8 = 8u * nthr;

    auto issue = [&](int i) {
        const int li = i * epb + c.g;
        if (i < ngrp && li < left) {
            const int e = e0 + li;
            const int4 em = s_em[li];
            const unsigned lvw = s_lv[li];
            const int nu1 = lvw & 0xff, nl1 = (lvw >> 8) & 0xff, nu2 = (lvw >> 16) & 0xff, nl2 = lvw >> 24;
            const int lo = nu2 > 0 ? min(nu1, nu2) : nu1, hi = max(nl1, nl2);
            const bool inr = nz >= lo && nz <= hi;
            const unsigned sb = sb0 + (unsigned)(i % D) * stage_bytes;
            const unsigned oe = (unsigned)e * L + nz0, o1 = (unsigned)em.x * L + nz0, o2 = (unsigned)em.y * L + nz0;
            if (inr) {
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (HOR != HOR_UPW1) {
                        const double* gp = b.grad[t] + (size_t)oe * 4;
                        cpa16_stream(sb + (C::c_g12 + t) * st16 + o16, gp);
                        cpa16_stream(sb + (C::c_g34 + t) * st16 + o16, gp + 2);
                    }
                    cpa8(sb + o8 + (C::c_t1 + t) * st8, &b.ttf[t][o1]);
                    cpa8(sb + o8 + (C::c_t2 + t) * st8, &b.ttf[t][o2]);
                    cpa8(sb + o8 + (C::c_a1 + t) * st8, &b.ttfAB[t][o1]);
                    cpa8(sb + o8 + (C::c_a2 + t) * st8, &b.ttfAB[t][o2]);
                }
                if (QMODE == 1) cpa8(sb + o8 + C::c_x * st8, &m.Q[oe]);
            }
            if (QMODE == 0) {
                bool use1, use2;
                edge_use(make_uchar4(nu1, nl1, nu2, nl2), nz, use1, use2);
                if (use1) {
                    const unsigned o = (unsigned)em.z * L + nz0;
                    cpa16(sb + (C::c_uv + 0) * st16 + o16, m.uv + (size_t)o * 2);
                    cpa8(sb + o8 + (C::c_x + 0) * st8, &m.helem[o]);
                }
                if (use2) {
                    const unsigned o = (unsigned)em.w * L + nz0;
                    cpa16(sb + (C::c_uv + 1) * st16 + o16, m.uv + (size_t)o * 2);
                    cpa8(sb + o8 + (C::c_x + 1) * st8, &m.helem[o]);
                }
            }
        }
        cpa_commit();
    };

    // ---- prologue: D-1 groups in flight -------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < D - 1; ++s) issue(s);
    for (int i = 0; i < ngrp; ++i) {
        issue(i + D - 1);
        cpa_wait<D - 1>();                       // this thread's copies of group i have landed
        const int li = i * epb + c.g;
        if (li >= left) continue;
        const int e = e0 + li;
        const unsigned lvw = s_lv[li];
        const int nu1 = lvw & 0xff, nl1 = (lvw >> 8) & 0xff, nu2 = (lvw >> 16) & 0xff, nl2 = lvw >> 24;
        const int lo = nu2 > 0 ? min(nu1, nu2) : nu1, hi = max(nl1, nl2);
        const bool inr = nz >= lo && nz <= hi;
        const unsigned oe = (unsigned)e * L + nz0;
        const unsigned char* sp = s_stage + (size_t)(i % D) * stage_bytes;
        const double2* c16 = reinterpret_cast<const double2*>(sp) + tid;
        const double* c8 = reinterpret_cast<const double*>(sp + (size_t)C::n16 * 16 * nthr) + tid;
        double q = 0.0;
        if (QMODE == 0) {
            bool use1, use2;
            edge_use(make_uchar4(nu1, nl1, nu2, nl2), nz, use1, use2);
            const double2 cr12 = s_cr[2 * li], cr34 = s_cr[2 * li + 1];
            double v1 = 0.0, v2 = 0.0;
            // Q(nz,e): vflux of oce_adv_tra_hor.F90:170,190,211-212,226,242
            if (use1) {
                const double2 uv1 = c16[(C::c_uv + 0) * nthr];
                v1 = (-uv1.y * cr12.x + uv1.x * cr12.y) * c8[(C::c_x + 0) * nthr];
            }
            if (use2) {
                const double2 uv2 = c16[(C::c_uv + 1) * nthr];
                v2 = (uv2.y * cr34.x - uv2.x * cr34.y) * c8[(C::c_x + 1) * nthr];
            }
            if (use1 && use2) q = v1 + v2;
            else if (use1) q = v1;
            else if (use2) q = v2;
            m.Q[oe] = q;
        } else if (inr) q = c8[C::c_x * nthr];
        double out[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) out[t] = 0.0;
        if (inr) {
            double2 ec = make_double2(0.0, 0.0);
            double clo1 = 1.0, clo2 = 1.0;
            if (HOR != HOR_UPW1) ec = s_ec[li];
            if (HOR == HOR_MUSCL) {
                const int2 nb = s_nb[li];
                clo1 = (nb.x - nz >= 0) ? 1.0 : 0.0;   // oce_adv_tra_hor.F90:411-412
                clo2 = (nb.y - nz >= 0) ? 1.0 : 0.0;
            }
            const double aq = fabs(q), qp = q + aq, qm = q - aq;
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                double2 g12 = make_double2(0.0, 0.0), g34 = g12;
                if (HOR != HOR_UPW1) { g12 = c16[(C::c_g12 + t) * nthr]; g34 = c16[(C::c_g34 + t) * nthr]; }
                const double t1 = c8[(C::c_t1 + t) * nthr], t2 = c8[(C::c_t2 + t) * nthr];
                const double a1 = c8[(C::c_a1 + t) * nthr], a2 = c8[(C::c_a2 + t) * nthr];
                const double flo = hor_lo(t1, t2, qp, qm);
                out[t] = hor_ho<HOR>(a1, a2, q, qp, qm, ec, g12, g34, b.ph[t], clo1, clo2, flo);
            }
        }
        stv<TB>(b.adf_h + (size_t)oe * TB, out);
    }
}

// ----------------------------------------------------------------------------------------------
// E1 (bulk): as k_edge_flux_p, but the contiguous streams of a group -- edge_up_dn_grad of the
// group's epb consecutive edges (epb * L * 32 bytes per tracer, 64 % of the kernel's DRAM bytes)
// -- are fetched by ONE bulk asynchronous copy each
// (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), issued by one elected thread.  Bulk
// copies go through the TMA unit, not through the per-thread load/store path whose outstanding-miss
// capacity the gather kernels saturate; only the genuine gathers (node columns of ttf/ttfAB,
// element columns of uv/helem; mostly L1/L2 hits) stay per-thread cp.async cells.
// Stage layout: [grad t=0 | grad t=1] in global order, then the cells [cell][thread].
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* b, int cnt)
{
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
                 ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned long long* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

#ifndef ADV_E1B_DIRECT
#define ADV_E1B_DIRECT 1     // 1: ttf/ttfAB end values are loaded straight to registers in the compute phase (not staged)
#endif
template <int TB, int QMODE>
struct E1bCells {
    static constexpr int n16 = (QMODE == 0 ? 2 : 0);
    static constexpr int n8 = (ADV_E1B_DIRECT ? 0 : 4 * TB) + (QMODE == 0 ? 2 : 1);   // + he1, he2 (QMODE 0) or q (QMODE 1)
    static constexpr int bytes = 16 * n16 + 8 * n8;     // per thread and stage
    static constexpr int c_uv = 0;
    static constexpr int c_t1 = 0, c_t2 = TB, c_a1 = 2 * TB, c_a2 = 3 * TB, c_he = (ADV_E1B_DIRECT ? 0 : 4 * TB);
};
__host__ __device__ constexpr size_t align128(size_t x) { return (x + 127) & ~(size_t)127; }
template <int TB, int QMODE>
__host__ __device__ inline size_t e1b_bulk_bytes(int nthr) { return align128((size_t)nthr * 32 * TB); }
template <int TB, int QMODE>
__host__ __device__ inline size_t e1b_stage_bytes(int nthr) { return e1b_bulk_bytes<TB, QMODE>(nthr) + align128((size_t)E1bCells<TB, QMODE>::bytes * nthr); }
template <int TB, int QMODE>
__host__ __device__ inline size_t e1b_smem_bytes(int nge, int nthr, int D)
{
    return align128(e1p_meta_bytes(nge)) + (size_t)D * e1b_stage_bytes<TB, QMODE>(nthr) + 64;
}

template <int HOR, int TB, int QMODE, int D>
__global__ void __launch_bounds__(kBlock, ADV_E1P_MINB) k_edge_flux_b(MeshDev m, Chunk<TB> b, int epb, int ng, int il)
{
    static_assert(HOR != HOR_UPW1, "the bulk variant stages edge_up_dn_grad");
    using C = E1bCells<TB, QMODE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int L = m.L, nthr = blockDim.x, tid = threadIdx.x;
    const int nge = epb * ng;
    // group i of this CTA: consecutive groups (il = 0) or grid-strided (il = 1: at any time the resident
    // CTAs then work on one dense window of the edge arrays, which keeps DRAM pages open longer)
    const int g0 = il ? (int)blockIdx.x : (int)blockIdx.x * ng, gs = il ? (int)gridDim.x : 1;
    const int ngroups = (m.E + epb - 1) / epb;
    int4* s_em = reinterpret_cast<int4*>(smem_raw);
    double2* s_cr = reinterpret_cast<double2*>(s_em + nge);
    double2* s_ec = s_cr + 2 * nge;
    int2* s_nb = reinterpret_cast<int2*>(s_ec + nge);
    unsigned* s_lv = reinterpret_cast<unsigned*>(s_nb + nge);
    unsigned char* s_stage = smem_raw + align128(e1p_meta_bytes(nge));
    const unsigned stage_bytes = (unsigned)e1b_stage_bytes<TB, QMODE>(nthr);
    const unsigned bulk_bytes = (unsigned)e1b_bulk_bytes<TB, QMODE>(nthr);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(s_stage + (size_t)D * stage_bytes);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < nge; i += nthr) {
        const int gi = i / epb;
        const int e = (g0 + gi * gs) * epb + (i - gi * epb);
        if (e < m.E) {
            const int4 em = __ldg(&m.edge_meta[e]);
            s_em[i] = em;
            s_lv[i] = __ldg(reinterpret_cast<const unsigned*>(m.edge_lev) + e);
            const double2* cr = reinterpret_cast<const double2*>(&m.edge_cross[e]);
            s_cr[2 * i] = __ldg(cr); s_cr[2 * i + 1] = __ldg(cr + 1);
            s_ec[i] = __ldg(&m.edge_c[e]);
            if (HOR == HOR_MUSCL) s_nb[i] = make_int2(__ldg(&m.nboundary_lay[em.x]), __ldg(&m.nboundary_lay[em.y]));
        }
    }
    __syncthreads();
    const ColThread c = col_thread(m);
    const int nz0 = c.nz0, nz = nz0 + 1;
    int ngrp = 0;                                             // groups of this CTA that exist
    for (; ngrp < ng && g0 + ngrp * gs < ngroups; ++ngrp) {}
    const unsigned sb0 = smem_u32(s_stage);
    const unsigned o16 = bulk_bytes + (unsigned)tid * 16u, o8 = bulk_bytes + (unsigned)C::n16 * 16u * nthr + (unsigned)tid * 8u;
    const unsigned st16 = 16u * nthr, st8 = 8u * nthr;
    const unsigned colw = (unsigned)L * 32u;          // bytes of one edge column of edge_up_dn_grad

    auto issue = [&](int i) {
        if (i < ngrp) {
            const unsigned sb = sb0 + (unsigned)(i % D) * stage_bytes;
            const int eg = (g0 + i * gs) * epb;                // first edge of the group
            if (tid == 0) {                                    // the contiguous streams of the group: bulk copies
                const int ne = min(epb, m.E - eg);             // edges of this group
                const size_t ecol = (size_t)eg * L;
                const unsigned gb = (unsigned)ne * colw;
                mbar_expect(&full[i % D], gb * TB);
#pragma unroll
                for (int t = 0; t < TB; ++t) bulk_g2s(sb + (unsigned)t * epb * colw, b.grad[t] + ecol * 4, gb, &full[i % D]);
            }
            const int li = i * epb + c.g;
            if (eg + c.g < m.E) {
                const int4 em = s_em[li];
                const unsigned lvw = s_lv[li];
                const int nu1 = lvw & 0xff, nl1 = (lvw >> 8) & 0xff, nu2 = (lvw >> 16) & 0xff, nl2 = lvw >> 24;
                const int lo = nu2 > 0 ? min(nu1, nu2) : nu1, hi = max(nl1, nl2);
                const unsigned o1 = (unsigned)em.x * L + nz0, o2 = (unsigned)em.y * L + nz0;
                if (nz >= lo && nz <= hi) {
                    if (!ADV_E1B_DIRECT) {
#pragma unroll
                        for (int t = 0; t < TB; ++t) {
                            cpa8(sb + o8 + (C::c_t1 + t) * st8, &b.ttf[t][o1]);
                            cpa8(sb + o8 + (C::c_t2 + t) * st8, &b.ttf[t][o2]);
                            cpa8(sb + o8 + (C::c_a1 + t) * st8, &b.ttfAB[t][o1]);
                            cpa8(sb + o8 + (C::c_a2 + t) * st8, &b.ttfAB[t][o2]);
                        }
                    }
                    if (QMODE == 1) cpa8(sb + o8 + C::c_he * st8, &m.Q[(unsigned)(eg + c.g) * L + nz0]);
                }
                if (QMODE == 0) {
                    bool use1, use2;
                    edge_use(make_uchar4(nu1, nl1, nu2, nl2), nz, use1, use2);
                    if (use1) {
                        const unsigned o = (unsigned)em.z * L + nz0;
                        cpa16(sb + (C::c_uv + 0) * st16 + o16, m.uv + (size_t)o * 2);
                        cpa8(sb + o8 + (C::c_he + 0) * st8, &m.helem[o]);
                    }
                    if (use2) {
                        const unsigned o = (unsigned)em.w * L + nz0;
                        cpa16(sb + (C::c_uv + 1) * st16 + o16, m.uv + (size_t)o * 2);
                        cpa8(sb + o8 + (C::c_he + 1) * st8, &m.helem[o]);
                    }
                }
            }
        }
        cpa_commit();
    };

#pragma unroll
    for (int s = 0; s < D - 1; ++s) issue(s);
    for (int i = 0; i < ngrp; ++i) {
        // everybody is done with group i-1, whose stage the copies of group i+D-1 overwrite
        __syncthreads();
        if (tid == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(i + D - 1);
        cpa_wait<D - 1>();
        mbar_wait(&full[i % D], (unsigned)((i / D) & 1));
        const int li = i * epb + c.g;
        const int e = (g0 + i * gs) * epb + c.g;
        if (e >= m.E) continue;
        const unsigned lvw = s_lv[li];
        const int nu1 = lvw & 0xff, nl1 = (lvw >> 8) & 0xff, nu2 = (lvw >> 16) & 0xff, nl2 = lvw >> 24;
        const int lo = nu2 > 0 ? min(nu1, nu2) : nu1, hi = max(nl1, nl2);
        const bool inr = nz >= lo && nz <= hi;
        const unsigned oe = (unsigned)e * L + nz0;
        const unsigned char* sp = s_stage + (size_t)(i % D) * stage_bytes;
        const double2* c16 = reinterpret_cast<const double2*>(sp + bulk_bytes) + tid;
        const double* c8 = reinterpret_cast<const double*>(sp + bulk_bytes + (size_t)C::n16 * 16 * nthr) + tid;
        double t1[TB], t2[TB], a1[TB], a2[TB];
        if (inr) {
            if (ADV_E1B_DIRECT) {
                const int4 em = s_em[li];
                const unsigned o1 = (unsigned)em.x * L + nz0, o2 = (unsigned)em.y * L + nz0;
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    t1[t] = __ldg(&b.ttf[t][o1]); t2[t] = __ldg(&b.ttf[t][o2]);
                    a1[t] = __ldg(&b.ttfAB[t][o1]); a2[t] = __ldg(&b.ttfAB[t][o2]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    t1[t] = c8[(C::c_t1 + t) * nthr]; t2[t] = c8[(C::c_t2 + t) * nthr];
                    a1[t] = c8[(C::c_a1 + t) * nthr]; a2[t] = c8[(C::c_a2 + t) * nthr];
                }
            }
        }
        double q = 0.0;
        if (QMODE == 0) {
            bool use1, use2;
            edge_use(make_uchar4(nu1, nl1, nu2, nl2), nz, use1, use2);
            const double2 cr12 = s_cr[2 * li], cr34 = s_cr[2 * li + 1];
            double v1 = 0.0, v2 = 0.0;
            // Q(nz,e): vflux of oce_adv_tra_hor.F90:170,190,211-212,226,242
            if (use1) {
                const double2 uv1 = c16[(C::c_uv + 0) * nthr];
                v1 = (-uv1.y * cr12.x + uv1.x * cr12.y) * c8[(C::c_he + 0) * nthr];
            }
            if (use2) {
                const double2 uv2 = c16[(C::c_uv + 1) * nthr];
                v2 = (uv2.y * cr34.x - uv2.x * cr34.y) * c8[(C::c_he + 1) * nthr];
            }
            if (use1 && use2) q = v1 + v2;
            else if (use1) q = v1;
            else if (use2) q = v2;
            m.Q[oe] = q;
        } else if (inr) q = c8[C::c_he * nthr];
        double out[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) out[t] = 0.0;
        if (inr) {
            const double2 ec = s_ec[li];
            double clo1 = 1.0, clo2 = 1.0;
            if (HOR == HOR_MUSCL) {
                const int2 nb = s_nb[li];
                clo1 = (nb.x - nz >= 0) ? 1.0 : 0.0;   // oce_adv_tra_hor.F90:411-412
                clo2 = (nb.y - nz >= 0) ? 1.0 : 0.0;
            }
            const double aq = fabs(q), qp = q + aq, qm = q - aq;
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const double2* gp = reinterpret_cast<const double2*>(sp + (size_t)t * epb * colw) + (size_t)tid * 2;
                const double2 g12 = gp[0], g34 = gp[1];
                const double flo = hor_lo(t1[t], t2[t], qp, qm);
                out[t] = hor_ho<HOR>(a1[t], a2[t], q, qp, qm, ec, g12, g34, b.ph[t], clo1, clo2, flo);
            }
        }
        stv<TB>(b.adf_h + (size_t)oe * TB, out);
    }
}

// ----------------------------------------------------------------------------------------------
// node kernels: shared metadata of a CTA = cn = cpb * ng consecutive range entries
//   int n[cn] (0-based node id, -1 past the end), uint2 rec[cn] (node_rec), int4 ell[cn][ell_w]
// ----------------------------------------------------------------------------------------------
struct NodeMetaS {
    int* n; uint2* rec; int4* ell;
    unsigned char* end;
};
__host__ __device__ inline size_t ndp_meta_bytes(int cn, int ell_w)
{
    return align16((size_t)cn * 4) + align16((size_t)cn * 8) + (size_t)cn * ell_w * 16;
}
__device__ __forceinline__ NodeMetaS ndp_meta_load(unsigned char* base, const MeshDev& m, const NodeRange& r, int cn)
{
    NodeMetaS s;
    s.ell = reinterpret_cast<int4*>(base);
    s.rec = reinterpret_cast<uint2*>(base + (size_t)cn * m.ell_w * 16);
    s.n = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(s.rec) + align16((size_t)cn * 8));
    s.end = reinterpret_cast<unsigned char*>(s.n) + align16((size_t)cn * 4);
    const int i0 = blockIdx.x * cn;
    for (int i = threadIdx.x; i < cn; i += blockDim.x) {
        int n = -1;
        uint2 rec = make_uint2(1u, 0u);          // nzmin = 1 > nzmax = 0: no valid layer
        if (i0 + i < r.count) {
            n = r.list ? __ldg(&r.list[r.begin + i0 + i]) : r.begin + i0 + i;
            rec = __ldg(&m.node_rec[n]);
        }
        s.n[i] = n; s.rec[i] = rec;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < cn * m.ell_w; k += blockDim.x) {
        const int i = k / m.ell_w, n = s.n[i];
        s.ell[k] = n >= 0 ? __ldg(&m.ne_ell[(size_t)n * m.ell_w + (k - i * m.ell_w)]) : ADV_EMPTY_SLOT;
    }
    __syncthreads();
    return s;
}

// ----------------------------------------------------------------------------------------------
// K3 (pipelined): limit the antidiffusive fluxes and accumulate the tendencies; formulas and
// citations as k_fct_update.  Cells per thread and stage, every cell an array [thread]:
//   16-byte "P" cells, one per tracer: {R+,R-}  own (id 0), of the other end of slot j (id 1+j)
//   "V" cells of TB doubles: vt = adf_v at the top interface (id 0; owned nodes; the thread above
//        reads it as its bottom interface), lo = fct_LO (id 1), f_j = adf_h of slot j (id 2+j)
//   8-byte cells: dh[t], tn[t], dv[t], av, hn, hnn
// ----------------------------------------------------------------------------------------------
template <int TB>
struct K3Cells {
    static constexpr int V = 8 * TB;
    __host__ __device__ static constexpr int nP(int W) { return (1 + W) * TB; }          // 16-byte cells
    __host__ __device__ static constexpr int oV(int W) { return nP(W) * 16; }            // byte offset per thread (x nthr)
    __host__ __device__ static constexpr int o8(int W) { return oV(W) + (2 + W) * V; }
    __host__ __device__ static constexpr int bytes(int W) { return o8(W) + 8 * (3 * TB + 3); }
};
template <int TB> __device__ __forceinline__ void ldsv(const double* p, double (&v)[TB]);
template <> __device__ __forceinline__ void ldsv<1>(const double* p, double (&v)[1]) { v[0] = p[0]; }
template <> __device__ __forceinline__ void ldsv<2>(const double* p, double (&v)[2])
{
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x; v[1] = t.y;
}

template <int TB, int D>
__global__ void __launch_bounds__(kBlock, ADV_NDP_MINB) k_fct_update_p(MeshDev m, Chunk<TB> b, NodeRange r, int ng, double dt)
{
    using C = K3Cells<TB>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = m.L, nl = m.nl, nthr = blockDim.x, tid = threadIdx.x, W = m.ell_w;
    const int cn = r.cpb * ng;
    const NodeMetaS ms = ndp_meta_load(smem_raw, m, r, cn);
    unsigned char* s_stage = ms.end;
    const ColThread c = col_thread(m);
    const int nz0 = c.nz0, nz = nz0 + 1;
    const int left = r.count - blockIdx.x * cn;
    const int ngrp = min(ng, (left + r.cpb - 1) / r.cpb);
    const unsigned stage_bytes = (unsigned)align16((size_t)C::bytes(W) * nthr);
    const unsigned sb0 = smem_u32(s_stage);
    const unsigned st16 = 16u * nthr, stV = (unsigned)C::V * nthr, st8 = 8u * nthr;
    const unsigned oP = (unsigned)tid * 16u;                                    // + id * st16
    const unsigned oV = (unsigned)C::oV(W) * nthr + (unsigned)tid * C::V;       // + id * stV
    const unsigned o8 = (unsigned)C::o8(W) * nthr + (unsigned)tid * 8u;         // + id * st8

    auto issue = [&](int i) {
        if (i < ngrp) {
            const int li = i * r.cpb + c.g;
            const int n = ms.n[li];
            const uint2 rec = ms.rec[li];
            const int nzmin = rec.x & 0xff, nzmax = (rec.x >> 8) & 0xff;
            const bool valid = nz >= nzmin && nz <= nzmax - 1;
            const bool owned = n < m.N;
            const unsigned sb = sb0 + (unsigned)(i % D) * stage_bytes;
            const unsigned oL = (unsigned)n * L + nz0;
            const size_t cN = (size_t)n * nl + nz0;
            if (owned && nz >= nzmin && nz <= nzmax)      // interface nz of an owned column
                cpa_vec<TB>(sb + oV + 0 * stV, b.adf_v + cN * TB);
            if (valid) {
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    cpa16(sb + oP + t * st16, b.pm + ((size_t)oL * TB + t) * 2);
                    cpa8(sb + o8 + (0 * TB + t) * st8, &b.dttf_h[t][oL]);
                }
                cpa8(sb + o8 + (3 * TB + 0) * st8, &m.areasvol[cN]);
                if (owned) {
                    cpa_vec<TB>(sb + oV + 1 * stV, b.lo + (size_t)oL * TB);
#pragma unroll
                    for (int t = 0; t < TB; ++t) {
                        cpa8(sb + o8 + (1 * TB + t) * st8, &b.ttf[t][oL]);
                        cpa8(sb + o8 + (2 * TB + t) * st8, &b.dttf_v[t][oL]);
                    }
                    cpa8(sb + o8 + (3 * TB + 1) * st8, &m.hnode[oL]);
                    cpa8(sb + o8 + (3 * TB + 2) * st8, &m.hnode_new[oL]);
                }
                const int deg = (rec.y >> 16) & 0xff;
                const int4* ell = ms.ell + li * W;
                for (int j = 0; j < deg; ++j) {
                    const int4 ent = ell[j];
                    const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
                    if (nz >= lo && nz <= hi) {
                        cpa_vec<TB>(sb + oV + (2 + j) * stV, b.adf_h + ((size_t)(unsigned)ent.x * L + nz0) * TB);
                        const double* src = b.pm + ((size_t)(unsigned)ent.y * L + nz0) * TB * 2;
#pragma unroll
                        for (int t = 0; t < TB; ++t) cpa16(sb + oP + ((1 + j) * TB + t) * st16, src + 2 * t);
                    }
                }
            }
        }
        cpa_commit();
    };

    if (D >= 2) issue(0);
    for (int i = 0; i < ngrp; ++i) {
        if (D == 1) issue(i);
        cpa_wait<0>();
        __syncthreads();           // all cells of group i visible; everybody is done with group i-1
        if (D >= 2) issue(i + 1);
        const int li = i * r.cpb + c.g;
        const int n = ms.n[li];
        const uint2 rec = ms.rec[li];
        const int nzmin = rec.x & 0xff, nzmax = (rec.x >> 8) & 0xff;
        const bool valid = nz >= nzmin && nz <= nzmax - 1;
        if (valid) {
            const bool owned = n < m.N;
            const unsigned oL = (unsigned)n * L + nz0;
            const unsigned char* sp = s_stage + (size_t)(i % D) * stage_bytes;
            const double2* cP = reinterpret_cast<const double2*>(sp) + tid;                       // + id * nthr
            const double* cV = reinterpret_cast<const double*>(sp + (size_t)C::oV(W) * nthr) + (size_t)tid * TB;   // + id * nthr * TB
            const double* c8 = reinterpret_cast<const double*>(sp + (size_t)C::o8(W) * nthr) + tid;  // + id * nthr
            double pk[TB], mk[TB], dh[TB];
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const double2 v = cP[t * nthr];
                pk[t] = v.x; mk[t] = v.y;
                dh[t] = c8[(0 * TB + t) * nthr];
            }
            const double av = c8[(3 * TB + 0) * nthr];
            const double r_av = 1.0 / av;
            if (owned) {
                const bool above = nz > nzmin, below = nz + 1 <= nzmax - 1;
                const bool has_below = nz0 + 1 < L;
                const double hn = c8[(3 * TB + 1) * nthr], hnn = c8[(3 * TB + 2) * nthr];
                double vt[TB], vb[TB], lo_n[TB];
                ldsv<TB>(cV, vt);
                ldsv<TB>(cV + 1 * nthr * TB, lo_n);
#pragma unroll
                for (int t = 0; t < TB; ++t) vb[t] = 0.0;
                if (has_below) ldsv<TB>(cV + TB, vb);          // top interface of the thread below
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    double pa = 1.0, ma = 1.0, pb = 1.0, mb = 1.0;
                    if (above) { const double2 v = cP[t * nthr - 1]; pa = v.x; ma = v.y; }
                    if (below) { const double2 v = cP[t * nthr + 1]; pb = v.x; mb = v.y; }
                    const double fv_top = limit_v(vt[t], nz, nzmin, nzmax, pa, ma, pk[t], mk[t]);
                    const double fv_bot = has_below ? limit_v(vb[t], nz + 1, nzmin, nzmax, pk[t], mk[t], pb, mb) : 0.0;
                    double d = c8[(2 * TB + t) * nthr];
                    d = d - c8[(1 * TB + t) * nthr] * hn + lo_n[t] * hnn;             // driver :535
                    d = d + div_rcp((fv_top - fv_bot) * dt, av, r_av);                // driver :556
                    b.dttf_v[t][oL] = d;
                }
            }
            const int deg = (rec.y >> 16) & 0xff;
            const int4* ell = ms.ell + li * W;
            for (int j = 0; j < deg; ++j) {
                const int4 ent = ell[j];
                const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
                if (nz < lo || nz > hi) continue;
                const bool second = (ent.z >> 16) & 1;
                double f[TB];
                ldsv<TB>(cV + (size_t)(2 + j) * nthr * TB, f);
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    const double ff = f[t];
                    const double2 o = cP[((1 + j) * TB + t) * nthr];
                    const double p1 = second ? o.x : pk[t], m1 = second ? o.y : mk[t];   // factors at edges(1,e)
                    const double p2 = second ? pk[t] : o.x, m2 = second ? mk[t] : o.y;   // factors at edges(2,e)
                    double ae = 1.0;
                    if (ff >= 0.0) { ae = dmin(ae, p1); ae = dmin(ae, m2); }      // fct :489-491
                    else { ae = dmin(ae, m1); ae = dmin(ae, p2); }                // :493-494
                    const double term = div_rcp(ae * ff * dt, av, r_av);          // fct :497, driver :607,:620
                    dh[t] = second ? dh[t] - term : dh[t] + term;
                }
            }
#pragma unroll
            for (int t = 0; t < TB; ++t) b.dttf_h[t][oL] = dh[t];
        }
        if (D == 1) __syncthreads();   // the single stage is refilled by the next issue
    }
}
template <int TB>
__host__ inline size_t k3p_smem_bytes(int cn, int ell_w, int nthr, int D)
{
    return ndp_meta_bytes(cn, ell_w) + (size_t)D * align16((size_t)K3Cells<TB>::bytes(ell_w) * nthr);
}

}  // namespace adv
