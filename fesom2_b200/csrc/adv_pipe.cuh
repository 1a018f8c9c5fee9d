// fesom2_b200/csrc/adv_pipe.cuh -- the bulk-copy edge kernel (E1 of the FCT path).
//
// Same arithmetic, same order of operations and same reference citations as k_edge_flux in
// adv_kernels.cuh; what changes is how operands reach the thread:
//   * a CTA owns `ng` consecutive groups of `epb` edge columns; the per-edge metadata of the whole
//     CTA is loaded ONCE into shared memory, so the dependent chain  index -> metadata -> operands
//     is paid once per CTA, not once per column;
//   * the contiguous stream (edge_up_dn_grad, 64 % of the DRAM bytes) arrives by bulk asynchronous
//     copies (cp.async.bulk / UBLKCP + mbarrier) issued by one thread, two groups deep: no registers
//     are held while it is in flight; a stage is released through a second mbarrier, so the loop has
//     no CTA-wide barrier;
//   * everything several edges share (end values ttf/ttfAB, element columns uv/helem) is a plain load
//     issued before the wait for the bulk copy (ADV_E1B_DIRECT = 2; the measured alternatives that stage
//     them in per-thread cp.async cells are kept behind ADV_E1B_DIRECT = 0 / 1).
// Measured history (register gather 1.89 ms -> cp.async cells 1.77 -> bulk 1.50 -> direct loads 1.43)
// in profiles/r1_tuning.md.
#pragma once
#include "adv_kernels.cuh"

namespace adv {

#ifndef ADV_E1P_MINB
#define ADV_E1P_MINB 4      // 4 CTAs of 7 warps per SM: measured best for the bulk edge kernel (profiles/r1_tuning.md)
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

// per-CTA edge metadata block: int4 em[nge], double2 cr[2 nge], double2 ec[nge], int4 eg[nge], int2 nb[nge], unsigned lv[nge]
__host__ __device__ inline size_t e1p_meta_bytes(int nge) { return align16((size_t)nge * (16 + 32 + 16 + 16 + 8 + 4)); }

// ----------------------------------------------------------------------------------------------
// E1 (bulk): edge_up_dn_grad of the group's epb consecutive edges (up to epb * L * 32 bytes per tracer)
// is fetched by one bulk asynchronous copy per edge column, trimmed to the column's wet levels
// (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), issued by one elected thread.  Bulk
// copies go through the TMA unit, not through the per-thread load/store path.
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
__device__ __forceinline__ void mbar_arrive(unsigned long long* b)
{
    asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
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

#ifndef ADV_E1B_NOSYNC
#define ADV_E1B_NOSYNC 1     // stage release through an mbarrier instead of a CTA-wide barrier per group
#endif
#ifndef ADV_E1B_DIRECT
#define ADV_E1B_DIRECT 2     // 0: every operand staged in cp.async cells; 1: ttf/ttfAB end values loaded straight to registers
                             // in the compute phase; 2: uv/helem/Q too (only edge_up_dn_grad is staged, by bulk copy)
#endif
template <int TB, int QMODE>
struct E1bCells {
    static constexpr int n16 = (ADV_E1B_DIRECT == 2 ? 0 : (QMODE == 0 ? 2 : 0));
    static constexpr int n8 = ADV_E1B_DIRECT == 2 ? 0 : (ADV_E1B_DIRECT ? 0 : 4 * TB) + (QMODE == 0 ? 2 : 1);   // + he1, he2 (QMODE 0) or q (QMODE 1)
    static constexpr int bytes = 16 * n16 + 8 * n8;     // per thread and stage
    static constexpr int c_uv = 0;
    static constexpr int c_t1 = 0, c_t2 = TB, c_a1 = 2 * TB, c_a2 = 3 * TB, c_he = (ADV_E1B_DIRECT ? 0 : 4 * TB);
};
__host__ __device__ constexpr size_t align128(size_t x) { return (x + 127) & ~(size_t)127; }
// GS = 0: the stage holds the edge's edge_up_dn_grad column (32 bytes per level and tracer);
// GS = 1: the gradients are reconstructed on the fly (fill_up_dn_grad fused, src/oce_muscl_adv.F90:378-522): the stage
//         holds four 16-byte-per-level columns per edge and tracer -- tr_xy of the up- and of the down-wind triangle,
//         the node-mean gradients of the two end nodes -- and every (edge, layer) picks the pair the reference
//         would have stored: the triangles' on the layers both end nodes share, the node means elsewhere.
template <int TB, int QMODE, int GS>
__host__ __device__ inline size_t e1b_bulk_bytes(int nthr) { return align128((size_t)nthr * (GS ? 64 : 32) * TB); }
template <int TB, int QMODE, int GS>
__host__ __device__ inline size_t e1b_stage_bytes(int nthr) { return e1b_bulk_bytes<TB, QMODE, GS>(nthr) + align128((size_t)E1bCells<TB, QMODE>::bytes * nthr); }
template <int TB, int QMODE, int GS>
__host__ __device__ inline size_t e1b_smem_bytes(int nge, int nthr, int D)
{
    return align128(e1p_meta_bytes(nge)) + (size_t)D * e1b_stage_bytes<TB, QMODE, GS>(nthr) + 16 * D;
}

#ifndef ADV_E1P_REGS
#define ADV_E1P_REGS 0      // > 0: explicit register budget instead of the __launch_bounds__ cap (experiments)
#endif
#if ADV_E1P_REGS > 0
#define ADV_E1P_BOUNDS __maxnreg__(ADV_E1P_REGS)
#else
#define ADV_E1P_BOUNDS __launch_bounds__(kBlock, ADV_E1P_MINB)
#endif
template <int HOR, int TB, int QMODE, int D, int GS>
__global__ void ADV_E1P_BOUNDS k_edge_flux_b(MeshDev m, Chunk<TB> b, int epb, int ng, int il, int pf)
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
    int4* s_eg = reinterpret_cast<int4*>(s_ec + nge);
    int2* s_nb = reinterpret_cast<int2*>(s_eg + nge);
    unsigned* s_lv = reinterpret_cast<unsigned*>(s_nb + nge);
    unsigned char* s_stage = smem_raw + align128(e1p_meta_bytes(nge));
    const unsigned stage_bytes = (unsigned)e1b_stage_bytes<TB, QMODE, GS>(nthr);
    const unsigned bulk_bytes = (unsigned)e1b_bulk_bytes<TB, QMODE, GS>(nthr);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(s_stage + (size_t)D * stage_bytes);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) { mbar_init(&full[s], 1); mbar_init(&full[D + s], nthr); }   // full[D+s] = "stage s is free again"
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // L2 prefetch of the metadata block of the CTA `pf` launches ahead (the only dependent wait of this kernel)
    if (pf > 0 && il == 0 && tid >= 32 && tid < 64) {
        const long long ef = ((long long)blockIdx.x + pf) * nge;
        if (ef < m.E) {
            const int ne = (int)min((long long)nge, (long long)m.E - ef);
            const int k = tid - 32;                        // 128-byte lines: 0-7 edge_cross, 8-11 edge_meta, 12-15 edge_c, 16 edge_lev
            const char* base = nullptr; unsigned bytes = 0, off = 0;
            if (k < 8) { base = reinterpret_cast<const char*>(m.edge_cross + ef); bytes = ne * 32u; off = k * 128u; }
            else if (k < 12) { base = reinterpret_cast<const char*>(m.edge_meta + ef); bytes = ne * 16u; off = (k - 8) * 128u; }
            else if (k < 16) { base = reinterpret_cast<const char*>(m.edge_c + ef); bytes = ne * 16u; off = (k - 12) * 128u; }
            else if (k == 16) { base = reinterpret_cast<const char*>(m.edge_lev + ef); bytes = ne * 4u; off = 0; }
            if (base && off < bytes + 127u) l2_prefetch_line(base + min(off, bytes - 1u));
        }
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
            if (GS) s_eg[i] = __ldg(&m.edge_g[e]);
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
            if (tid == 0) {                                    // the contiguous streams of the group: bulk copies,
                const int ne = min(epb, m.E - eg);             // one per edge column, trimmed to its wet levels 1..hi
                unsigned total = 0;
                for (int k = 0; k < ne; ++k) {
                    const unsigned lvw = s_lv[i * epb + k];
                    const unsigned hi = max((lvw >> 8) & 0xff, lvw >> 24);
                    if (GS) { const int4 g4 = s_eg[i * epb + k]; total += hi * 16u * (2u + (g4.x >= 0 ? 1u : 0u) + (g4.y >= 0 ? 1u : 0u)); }
                    else total += hi * 32u;
                }
                mbar_expect(&full[i % D], total * TB);
                for (int k = 0; k < ne; ++k) {
                    const unsigned lvw = s_lv[i * epb + k];
                    const unsigned hi = max((lvw >> 8) & 0xff, lvw >> 24);
                    if (GS) {
                        const int4 g4 = s_eg[i * epb + k];
                        const int4 em = s_em[i * epb + k];
                        const unsigned cw = (unsigned)L * 16u, nb = hi * 16u;
#pragma unroll
                        for (int t = 0; t < TB; ++t) {
                            const unsigned st = sb + (unsigned)(t * 4 * epb + k) * cw;
                            if (g4.x >= 0) bulk_g2s(st, b.txy[t] + (size_t)g4.x * L * 2, nb, &full[i % D]);
                            if (g4.y >= 0) bulk_g2s(st + (unsigned)epb * cw, b.txy[t] + (size_t)g4.y * L * 2, nb, &full[i % D]);
                            bulk_g2s(st + 2u * epb * cw, b.gmean[t] + (size_t)em.x * L * 2, nb, &full[i % D]);
                            bulk_g2s(st + 3u * epb * cw, b.gmean[t] + (size_t)em.y * L * 2, nb, &full[i % D]);
                        }
                    } else {
                        const unsigned nb = hi * 32u;
                        const size_t ecol = (size_t)(eg + k) * L;
#pragma unroll
                        for (int t = 0; t < TB; ++t) bulk_g2s(sb + (unsigned)(t * epb + k) * colw, b.grad[t] + ecol * 4, nb, &full[i % D]);
                    }
                }
            }
            const int li = i * epb + c.g;
            if (ADV_E1B_DIRECT != 2 && eg + c.g < m.E) {
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
        if (ADV_E1B_DIRECT != 2) cpa_commit();
    };

#pragma unroll
    for (int s = 0; s < D - 1; ++s) issue(s);
    for (int i = 0; i < ngrp; ++i) {
        // the copies of group i+D-1 overwrite the stage of group i-1: the issuing thread waits until all
        // threads have released it (mbarrier, no CTA-wide barrier: the other warps run ahead)
#if ADV_E1B_NOSYNC
        if (tid == 0 && i >= 1 && i + D - 1 < ngrp) {
            mbar_wait(&full[D + (i - 1) % D], (unsigned)(((i - 1) / D) & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
#else
        __syncthreads();
        if (tid == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        issue(i + D - 1);
        if (ADV_E1B_DIRECT != 2) cpa_wait<D - 1>();
        const int li = i * epb + c.g;
        const int e = (g0 + i * gs) * epb + c.g;
        if (e >= m.E) {                                       // no edge (partial last CTA): keep in step with the others
            mbar_wait(&full[i % D], (unsigned)((i / D) & 1));
#if ADV_E1B_NOSYNC
            mbar_arrive(&full[D + i % D]);
#endif
            continue;
        }
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
        bool use1 = false, use2 = false;
        double2 uv1 = make_double2(0.0, 0.0), uv2 = uv1;
        double he1 = 0.0, he2 = 0.0;
        if (QMODE == 0) {
            edge_use(make_uchar4(nu1, nl1, nu2, nl2), nz, use1, use2);
            if (ADV_E1B_DIRECT == 2) {                       // element columns straight from L1/L2 (three edges share them)
                const int4 em = s_em[li];
                if (use1) { const unsigned o = (unsigned)em.z * L + nz0; uv1 = __ldg(reinterpret_cast<const double2*>(m.uv) + o); he1 = __ldg(&m.helem[o]); }
                if (use2) { const unsigned o = (unsigned)em.w * L + nz0; uv2 = __ldg(reinterpret_cast<const double2*>(m.uv) + o); he2 = __ldg(&m.helem[o]); }
            } else {
                if (use1) { uv1 = c16[(C::c_uv + 0) * nthr]; he1 = c8[(C::c_he + 0) * nthr]; }
                if (use2) { uv2 = c16[(C::c_uv + 1) * nthr]; he2 = c8[(C::c_he + 1) * nthr]; }
            }
        } else if (inr) q = ADV_E1B_DIRECT == 2 ? __ldg(&m.Q[oe]) : c8[C::c_he * nthr];
        mbar_wait(&full[i % D], (unsigned)((i / D) & 1));   // the bulk-copied gradient columns of group i
        if (QMODE == 0) {
            const double2 cr12 = s_cr[2 * li], cr34 = s_cr[2 * li + 1];
            double v1 = 0.0, v2 = 0.0;
            // Q(nz,e): vflux of oce_adv_tra_hor.F90:170,190,211-212,226,242
            if (use1) v1 = (-uv1.y * cr12.x + uv1.x * cr12.y) * he1;
            if (use2) v2 = (uv2.y * cr34.x - uv2.x * cr34.y) * he2;
            if (use1 && use2) q = v1 + v2;
            else if (use1) q = v1;
            else if (use2) q = v2;
            if (use1 || use2 || inr) m.Q[oe] = q;       // dry levels keep the zero of the allocation
        }
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
                double2 g12, g34;
                if (GS) {                                  // fill_up_dn_grad :435-440 (shared layers) / :391-406 (node means)
                    const int4 g4 = s_eg[li];
                    const bool sh = nz >= g4.z && nz <= g4.w;
                    const double2* base = reinterpret_cast<const double2*>(sp + (size_t)t * 4 * epb * L * 16);
                    const double2 up = base[(size_t)((sh ? 0 : 2) * epb + c.g) * L + nz0];
                    const double2 dn = base[(size_t)((sh ? 1 : 3) * epb + c.g) * L + nz0];
                    g12 = make_double2(up.x, dn.x); g34 = make_double2(up.y, dn.y);
                } else {
                    const double2* gp = reinterpret_cast<const double2*>(sp + (size_t)t * epb * colw) + (size_t)tid * 2;
                    g12 = gp[0]; g34 = gp[1];
                }
                const double flo = b.nolo ? 0.0 : hor_lo(t1[t], t2[t], qp, qm);
                out[t] = hor_ho<HOR>(a1[t], a2[t], q, qp, qm, ec, g12, g34, b.ph[t], clo1, clo2, flo);
            }
        }
        if (inr) stv<TB>(b.adf_h + (size_t)oe * TB, out);   // out-of-range levels are never read and stay zero
#if ADV_E1B_NOSYNC
        mbar_arrive(&full[D + i % D]);                         // this thread is done with the stage
#endif
    }
}

}  // namespace adv
