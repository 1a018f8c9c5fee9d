// fesom2_b200/csrc/adv_staged.cuh -- shared-memory staged versions of the four FCT-path kernels.
//
// Why: the register-gather kernels of adv_kernels.cuh are latency bound (ncu: 7-16 warps stalled on
// long_scoreboard per issue, 2.4-4.0 TB/s) because every byte in flight costs a register and the
// slot -> operand load chain is serial.  Here the operands of a CTA -- whole columns, each a
// contiguous run of L (or nl) doubles in the reference's level-fastest layout -- are fetched by
// 1-D bulk async copies (cp.async.bulk global -> shared, completion on an mbarrier; SASS UBLKCP),
// issued by a few threads right after the CTA's metadata arrives.  Memory-level parallelism no
// longer depends on registers or occupancy: a CTA has 30-60 KB in flight from one instruction per
// column, and the (column, layer) threads then compute from shared memory.
//
//   * only the wet level range of every column is copied (union over the CTA's users)
//   * columns used by several edges/nodes of the CTA are fetched once (in-CTA dedupe): the node
//     kernels read a neighbour column ~2x less often from L2 than the register-gather version
//   * values a thread needs only at its own (column, layer) are plain loads issued BEFORE the
//     barrier wait, so their latency overlaps the bulk copies
//
// Arithmetic is shared with adv_kernels.cuh (same helper functions, same expression order): the
// results are bit-identical to the register-gather kernels and to the CPU oracle.
//
// Alignment contract: cp.async.bulk needs 16-byte aligned source, destination and size.  Columns
// of L doubles start on 8-byte boundaries when L is odd, so a copy is widened to the enclosing
// 16-byte window and the slot keeps the source's alignment phase (`shift`).  All field base
// pointers must be 16-byte aligned (checked on the host); a widened copy may read up to 8 bytes
// past the last element of an array, which stays inside the allocation granule of cudaMalloc and
// of every caching allocator in use (>= 256 B).
#pragma once
#include "adv_kernels.cuh"

namespace adv {

#ifndef ADV_E1S_MINB
#define ADV_E1S_MINB 4
#endif
#ifndef ADV_N1S_MINB
#define ADV_N1S_MINB 4
#endif
#ifndef ADV_K2S_MINB
#define ADV_K2S_MINB 4
#endif
#ifndef ADV_K3S_MINB
#define ADV_K3S_MINB 4
#endif

// ---------------------------------------------------------------------------------------------
// async-copy primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ADV_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ADV_DONE_%=;\n"
        "bra ADV_WAIT_%=;\n"
        "ADV_DONE_%=:\n"
        "}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// a slot that can hold a column of `len` doubles in either alignment phase plus the widening
__host__ __device__ constexpr uint32_t slot_bytes(int len) { return (uint32_t)(((len * 8 + 15) & ~15) + 16); }

// Stage levels [k0, k1] (0-based, inclusive; w doubles per level) of the column starting at `col`
// into `slot`: element i (in doubles) of the column lands at slot + (col & 15) + 8 i.
// ISSUE = false: only return the byte count (for expect_tx); ISSUE = true: start the copy.
template <bool ISSUE>
__device__ __forceinline__ uint32_t stage_col(uint32_t slot, const double* col, int w, int k0, int k1, uint64_t* bar)
{
    if (k1 < k0) return 0;
    const unsigned long long base = (unsigned long long)col;
    const unsigned long long g0 = base + (unsigned long long)(k0 * w) * 8ull;
    const unsigned long long g1 = base + (unsigned long long)((k1 + 1) * w) * 8ull;
    const unsigned long long a0 = g0 & ~15ull, a1 = (g1 + 15ull) & ~15ull;
    const uint32_t bytes = (uint32_t)(a1 - a0);
    if (ISSUE) bulk_g2s(slot + (uint32_t)(base & 15ull) + (uint32_t)(k0 * w * 8) - (uint32_t)(g0 - a0), (const void*)a0, bytes, bar);
    return bytes;
}
// pointer to element 0 of a staged column
__device__ __forceinline__ const double* col_ptr(const unsigned char* slot, const double* col)
{
    return reinterpret_cast<const double*>(slot + ((unsigned long long)col & 15ull));
}

// ---------------------------------------------------------------------------------------------
// in-CTA column dedupe.  Entries {id, lo | hi << 8} (id < 0: inactive) live in shared memory;
// pass 1 finds the first entry with the same id (the leader) and the union of the level ranges,
// pass 2 turns leaders into dense slot numbers.  O(n) per entry, n <= ~80.
// ---------------------------------------------------------------------------------------------
struct DedupTab {
    int2* ent;
    unsigned short* rep;    // first entry with the same id
    unsigned short* ur;     // leader: union range lo | hi << 8
    unsigned char* lead;    // 1 = leader
    unsigned short* slot;   // dense slot of the entry's column
    int n;
};
__host__ __device__ inline uint32_t dedup_bytes(int n) { return (uint32_t)(((n * (8 + 2 + 2 + 2 + 1)) + 15) & ~15) + 64; }
__device__ __forceinline__ DedupTab dedup_carve(unsigned char* p, int n)
{
    DedupTab d;
    d.n = n;
    d.ent = reinterpret_cast<int2*>(p); p += (size_t)n * 8;
    d.rep = reinterpret_cast<unsigned short*>(p); p += (size_t)n * 2;
    d.ur = reinterpret_cast<unsigned short*>(p); p += (size_t)n * 2;
    d.slot = reinterpret_cast<unsigned short*>(p); p += (size_t)n * 2;
    d.lead = p;
    return d;
}
__device__ __forceinline__ void dedup_scan(const DedupTab& d, int i)
{
    const int2 me = d.ent[i];
    int rep = i, lo = me.y & 0xff, hi = (me.y >> 8) & 0xff;
    if (me.x >= 0) {
        for (int j = 0; j < d.n; ++j) {
            const int2 o = d.ent[j];
            if (o.x == me.x) {
                rep = min(rep, j);
                lo = min(lo, o.y & 0xff);
                hi = max(hi, (o.y >> 8) & 0xff);
            }
        }
    }
    d.rep[i] = (unsigned short)rep;
    d.ur[i] = (unsigned short)(lo | (hi << 8));
    d.lead[i] = (me.x >= 0 && rep == i) ? 1 : 0;
}
__device__ __forceinline__ int dedup_slot(const DedupTab& d, int i)
{
    const int rep = d.rep[i];
    int s = 0;
    for (int j = 0; j < rep; ++j) s += d.lead[j];
    d.slot[i] = (unsigned short)s;
    return s;
}

// =============================================================================================
// E1 staged: antidiffusive horizontal flux (same contract as k_edge_flux)
//   per edge: TB gradient columns (4 doubles per level), Q (QMODE 1);
//   per distinct end node: ttf[t], ttfAB[t]; per distinct element (QMODE 0): uv (2 per level), helem
// =============================================================================================
struct E1Layout {
    uint32_t meta, lev, ntab, etab, grad, node, elem, total;
    uint32_t SC, SG, SU, node_slot, elem_slot;
};
__host__ __device__ inline E1Layout e1_layout(int L, int epb, int TB, int HOR, int QMODE)
{
    E1Layout y;
    y.SC = slot_bytes(L); y.SG = slot_bytes(4 * L); y.SU = slot_bytes(2 * L);
    y.node_slot = 2 * TB * y.SC;
    y.elem_slot = (QMODE == 0) ? y.SU + y.SC : y.SC;   // QMODE 1: one Q column per edge
    uint32_t o = 16;
    y.meta = o; o += (uint32_t)epb * 16;
    y.lev = o; o += (uint32_t)((epb * 4 + 15) & ~15);
    y.ntab = o; o += dedup_bytes(2 * epb);
    y.etab = o; o += dedup_bytes(2 * epb);
    y.grad = o; o += (HOR != HOR_UPW1) ? (uint32_t)epb * TB * y.SG : 0;
    y.node = o; o += (uint32_t)(2 * epb) * y.node_slot;
    y.elem = o; o += (uint32_t)((QMODE == 0) ? 2 * epb : epb) * y.elem_slot;
    y.total = o;
    return y;
}

template <int HOR, int TB, int QMODE, bool ISSUE>
__device__ __forceinline__ uint32_t e1_items(const MeshDev& m, const Chunk<TB>& b, const E1Layout& y, unsigned char* smem,
                                             const DedupTab& nt, const DedupTab& et, int epb, uint64_t* bar)
{
    const int L = m.L, tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t s0 = smem_u32(smem);
    const int4* s_meta = reinterpret_cast<const int4*>(smem + y.meta);
    const uchar4* s_lev = reinterpret_cast<const uchar4*>(smem + y.lev);
    uint32_t bytes = 0;
    // node columns (leaders only)
    for (int i = tid; i < 2 * epb; i += nthr) {
        if (!nt.lead[i]) continue;
        const int id = nt.ent[i].x, k0 = (nt.ur[i] & 0xff) - 1, k1 = (nt.ur[i] >> 8) - 1;
        const uint32_t sl = s0 + y.node + (uint32_t)nt.slot[i] * y.node_slot;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            bytes += stage_col<ISSUE>(sl + t * y.SC, b.ttf[t] + (size_t)id * L, 1, k0, k1, bar);
            bytes += stage_col<ISSUE>(sl + (TB + t) * y.SC, b.ttfAB[t] + (size_t)id * L, 1, k0, k1, bar);
        }
    }
    // element columns (QMODE 0) / Q columns (QMODE 1) -- handled by the upper threads so that the
    // two classes are issued by different warps
    for (int i = nthr - 1 - tid; i < 2 * epb; i += nthr) {
        if (QMODE == 0) {
            if (!et.lead[i]) continue;
            const int id = et.ent[i].x, k0 = (et.ur[i] & 0xff) - 1, k1 = (et.ur[i] >> 8) - 1;
            const uint32_t sl = s0 + y.elem + (uint32_t)et.slot[i] * y.elem_slot;
            bytes += stage_col<ISSUE>(sl, m.uv + (size_t)id * L * 2, 2, k0, k1, bar);
            bytes += stage_col<ISSUE>(sl + y.SU, m.helem + (size_t)id * L, 1, k0, k1, bar);
        }
    }
    // per-edge columns: gradients, stored Q
    for (int i = (tid + nthr / 2) % nthr; i < epb * (TB + 1); i += nthr) {
        const int g = i / (TB + 1), t = i - g * (TB + 1);
        const int4 em = s_meta[g];
        if (em.x < 0) continue;
        const uchar4 lv = s_lev[g];
        const int lo = lv.z > 0 ? min((int)lv.x, (int)lv.z) : (int)lv.x, hi = max((int)lv.y, (int)lv.w);
        const int e = blockIdx.x * epb + g;
        if (t < TB) {
            if (HOR != HOR_UPW1)
                bytes += stage_col<ISSUE>(s0 + y.grad + (uint32_t)(g * TB + t) * y.SG, b.grad[t] + (size_t)e * L * 4, 4, lo - 1, hi - 1, bar);
        } else if (QMODE == 1) {
            bytes += stage_col<ISSUE>(s0 + y.elem + (uint32_t)g * y.elem_slot, m.Q + (size_t)e * L, 1, lo - 1, hi - 1, bar);
        }
    }
    return bytes;
}

template <int HOR, int TB, int QMODE>
__global__ void __launch_bounds__(kBlock, ADV_E1S_MINB) k_edge_flux_s(MeshDev m, Chunk<TB> b, int epb)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = m.L, tid = threadIdx.x, nthr = blockDim.x;
    const E1Layout y = e1_layout(L, epb, TB, HOR, QMODE);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    int4* s_meta = reinterpret_cast<int4*>(smem + y.meta);
    uchar4* s_lev = reinterpret_cast<uchar4*>(smem + y.lev);
    const DedupTab nt = dedup_carve(smem + y.ntab, 2 * epb), et = dedup_carve(smem + y.etab, 2 * epb);

    // ---- wait 1: per-edge metadata -----------------------------------------------------------------
    if (tid == 0) { mbar_init(bar, nthr); fence_async_smem(); }
    if (tid < epb) {
        const int e = blockIdx.x * epb + tid;
        int4 em = make_int4(-1, -1, -1, -1);
        uchar4 lv = make_uchar4(1, 0, 0, 0);
        if (e < m.E) { em = __ldg(&m.edge_meta[e]); lv = __ldg(&m.edge_lev[e]); }
        s_meta[tid] = em; s_lev[tid] = lv;
        // dedupe entries: end nodes with the edge's scatter range (oce_adv_tra_driver.F90:154-156),
        // elements with the levels where they contribute to Q (oce_adv_tra_hor.F90:127-160)
        const int lo = lv.z > 0 ? min((int)lv.x, (int)lv.z) : (int)lv.x, hi = max((int)lv.y, (int)lv.w);
        const int rng = (lo & 0xff) | ((hi & 0xff) << 8);
        const bool act = em.x >= 0 && lo <= hi;
        nt.ent[2 * tid] = make_int2(act ? em.x : -1, rng);
        nt.ent[2 * tid + 1] = make_int2(act ? em.y : -1, rng);
        const int lo1 = em.w >= 0 ? (int)lv.x : 1;   // boundary edge: range D runs from nz = 1
        et.ent[2 * tid] = make_int2(em.x >= 0 ? em.z : -1, (lo1 & 0xff) | ((int)lv.y << 8));
        et.ent[2 * tid + 1] = make_int2((em.x >= 0 && em.w >= 0) ? em.w : -1, (int)lv.z | ((int)lv.w << 8));
    }
    __syncthreads();
    const ColThread c = col_thread(m);
    const int g = c.g, nz0 = c.nz0, nz = nz0 + 1;
    const int4 em = (g < epb) ? s_meta[g] : make_int4(-1, -1, -1, -1);
    const bool have = em.x >= 0;
    const int e = blockIdx.x * epb + g;
    // per-edge constants: plain loads issued before the barrier wait
    double2 cr12 = make_double2(0.0, 0.0), cr34 = cr12, ec = cr12;
    double clo1 = 1.0, clo2 = 1.0;
    if (have) {
        if (QMODE == 0) {
            cr12 = __ldg(reinterpret_cast<const double2*>(&m.edge_cross[e]));
            cr34 = __ldg(reinterpret_cast<const double2*>(&m.edge_cross[e]) + 1);
        }
        if (HOR != HOR_UPW1) ec = __ldg(&m.edge_c[e]);
        if (HOR == HOR_MUSCL) {
            clo1 = (__ldg(&m.nboundary_lay[em.x]) - nz >= 0) ? 1.0 : 0.0;   // oce_adv_tra_hor.F90:411-412
            clo2 = (__ldg(&m.nboundary_lay[em.y]) - nz >= 0) ? 1.0 : 0.0;
        }
    }
    // ---- dedupe + issue ----------------------------------------------------------------------------
    for (int i = tid; i < 2 * epb; i += nthr) dedup_scan(nt, i);
    if (QMODE == 0) for (int i = nthr - 1 - tid; i < 2 * epb; i += nthr) dedup_scan(et, i);
    __syncthreads();
    for (int i = tid; i < 2 * epb; i += nthr) dedup_slot(nt, i);
    if (QMODE == 0) for (int i = nthr - 1 - tid; i < 2 * epb; i += nthr) dedup_slot(et, i);
    __syncthreads();
    {
        const uint32_t bytes = e1_items<HOR, TB, QMODE, false>(m, b, y, smem, nt, et, epb, bar);
        if (bytes) mbar_arrive_tx(bar, bytes); else mbar_arrive(bar);
        if (bytes) e1_items<HOR, TB, QMODE, true>(m, b, y, smem, nt, et, epb, bar);
    }
    mbar_wait(bar, 0);
    if (!have) return;

    // ---- compute from shared memory ----------------------------------------------------------------
    const uchar4 lv = s_lev[g];
    const int lo = lv.z > 0 ? min((int)lv.x, (int)lv.z) : (int)lv.x;
    const int hi = max((int)lv.y, (int)lv.w);
    const bool inr = nz >= lo && nz <= hi;
    bool use1 = false, use2 = false;
    if (QMODE == 0) edge_use(lv, nz, use1, use2);
    const unsigned oe = (unsigned)e * L + nz0;
    double q = 0.0;
    if (QMODE == 0) {
        double2 uv1 = make_double2(0.0, 0.0), uv2 = uv1;
        double he1 = 0.0, he2 = 0.0;
        if (use1) {
            const unsigned char* sl = smem + y.elem + (uint32_t)et.slot[2 * g] * y.elem_slot;
            uv1 = reinterpret_cast<const double2*>(col_ptr(sl, m.uv + (size_t)em.z * L * 2))[nz0];
            he1 = col_ptr(sl + y.SU, m.helem + (size_t)em.z * L)[nz0];
        }
        if (use2) {
            const unsigned char* sl = smem + y.elem + (uint32_t)et.slot[2 * g + 1] * y.elem_slot;
            uv2 = reinterpret_cast<const double2*>(col_ptr(sl, m.uv + (size_t)em.w * L * 2))[nz0];
            he2 = col_ptr(sl + y.SU, m.helem + (size_t)em.w * L)[nz0];
        }
        // Q(nz,e): vflux of oce_adv_tra_hor.F90:170,190,211-212,226,242
        const double v1 = (-uv1.y * cr12.x + uv1.x * cr12.y) * he1;
        const double v2 = (uv2.y * cr34.x - uv2.x * cr34.y) * he2;
        if (use1 && use2) q = v1 + v2;
        else if (use1) q = v1;
        else if (use2) q = v2;
        // outside [lo,hi] and the element ranges Q keeps the zero it was allocated with
        if (use1 || use2 || inr) m.Q[oe] = q;
    } else if (inr) {
        q = col_ptr(smem + y.elem + (uint32_t)g * y.elem_slot, m.Q + (size_t)e * L)[nz0];
    }
    if (!inr) return;   // adv_flux_hor outside the scatter range stays at its allocation-time zero
    const unsigned char* np1 = smem + y.node + (uint32_t)nt.slot[2 * g] * y.node_slot;
    const unsigned char* np2 = smem + y.node + (uint32_t)nt.slot[2 * g + 1] * y.node_slot;
    const double aq = fabs(q), qp = q + aq, qm = q - aq;
    double out[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const double t1 = col_ptr(np1 + t * y.SC, b.ttf[t] + (size_t)em.x * L)[nz0];
        const double t2 = col_ptr(np2 + t * y.SC, b.ttf[t] + (size_t)em.y * L)[nz0];
        const double a1 = col_ptr(np1 + (TB + t) * y.SC, b.ttfAB[t] + (size_t)em.x * L)[nz0];
        const double a2 = col_ptr(np2 + (TB + t) * y.SC, b.ttfAB[t] + (size_t)em.y * L)[nz0];
        double2 g12 = make_double2(0.0, 0.0), g34 = g12;
        if (HOR != HOR_UPW1) {
            const double2* gp = reinterpret_cast<const double2*>(smem + y.grad + (uint32_t)(g * TB + t) * y.SG) + nz0 * 2;
            g12 = gp[0]; g34 = gp[1];
        }
        const double flo = hor_lo(t1, t2, qp, qm);
        out[t] = hor_ho<HOR>(a1, a2, q, qp, qm, ec, g12, g34, b.ph[t], clo1, clo2, flo);
    }
    stv<TB>(b.adf_h + (size_t)oe * TB, out);
}

// =============================================================================================
// node kernels: shared prologue.  A CTA owns cpb node columns; the gather tables (ELL slots) are
// pulled into shared memory, neighbour-node and edge columns are deduplicated, the leaders issue
// the bulk copies.
// =============================================================================================
struct NodeLayout {
    uint32_t nid, rec, ent, ntab, etab, node, edge, own, extra, total;
    uint32_t node_slot, edge_slot, own_slot;
    int nn, ne;
};
// ns_max / es_max: upper bounds of distinct node / edge columns per CTA (host: range_max_distinct)
__host__ __device__ inline NodeLayout node_layout(int cpb, int ell_w, int ns_max, int es_max, uint32_t node_slot,
                                                  uint32_t edge_slot, uint32_t own_slot, uint32_t extra)
{
    NodeLayout y;
    y.nn = cpb * (1 + ell_w); y.ne = cpb * ell_w;
    y.node_slot = node_slot; y.edge_slot = edge_slot; y.own_slot = own_slot;
    uint32_t o = 16;
    y.nid = o; o += (uint32_t)((cpb * 4 + 15) & ~15);
    y.rec = o; o += (uint32_t)((cpb * 8 + 15) & ~15);
    y.ent = o; o += (uint32_t)y.ne * 16;
    y.ntab = o; o += dedup_bytes(y.nn);
    y.etab = o; o += dedup_bytes(y.ne);
    y.node = o; o += (uint32_t)ns_max * node_slot;
    y.edge = o; o += (uint32_t)es_max * edge_slot;
    y.own = o; o += (uint32_t)cpb * own_slot;
    y.extra = o; o += (extra + 15u) & ~15u;
    y.total = o;
    return y;
}

struct StageRange {   // NodeRange + the shared-memory sizing of the staged kernels
    NodeRange r;
    int ns_max, es_max;
};

struct NodeCtx {
    uint64_t* bar;
    int* s_nid;
    uint2* s_rec;
    int4* s_ent;
    DedupTab nt, et;
};

// steps 1-3 of the prologue: metadata -> shared memory, dedupe tables.  On return the tables are
// complete (two __syncthreads inside) and the caller issues its copies.
// halo_rows: K3 also runs on halo nodes, whose valid range is the node's own levels as well.
__device__ __forceinline__ NodeCtx node_prologue(const MeshDev& m, const NodeRange& r, const NodeLayout& y, unsigned char* smem)
{
    const int tid = threadIdx.x, nthr = blockDim.x, cpb = r.cpb, W = m.ell_w;
    NodeCtx c;
    c.bar = reinterpret_cast<uint64_t*>(smem);
    c.s_nid = reinterpret_cast<int*>(smem + y.nid);
    c.s_rec = reinterpret_cast<uint2*>(smem + y.rec);
    c.s_ent = reinterpret_cast<int4*>(smem + y.ent);
    c.nt = dedup_carve(smem + y.ntab, y.nn);
    c.et = dedup_carve(smem + y.etab, y.ne);
    if (tid == 0) { mbar_init(c.bar, nthr); fence_async_smem(); }
    // ---- wait 1: node ids, records and ELL rows (one pass; the list indirection adds a hop) ----------
    for (int i = tid; i < y.ne; i += nthr) {
        const int g = i / W, j = i - g * W;
        const int k = blockIdx.x * cpb + g;
        int4 ent = ADV_EMPTY_SLOT;
        if (k < r.count) {
            const int n = r.list ? __ldg(&r.list[r.begin + k]) : r.begin + k;
            ent = __ldg(&m.ne_ell[(size_t)n * W + j]);
            if (j == 0) { c.s_nid[g] = n; c.s_rec[g] = __ldg(&m.node_rec[n]); }
        } else if (j == 0) { c.s_nid[g] = -1; c.s_rec[g] = make_uint2(1u, 0u); }   // nzmin 1 > nzmax-1 = -1: empty
        c.s_ent[i] = ent;
    }
    __syncthreads();
    // ---- dedupe entries: self columns, neighbour columns, edge columns --------------------------------
    for (int i = tid; i < y.nn; i += nthr) {
        int2 en = make_int2(-1, 0xff);
        if (i < cpb) {
            const uint2 rec = c.s_rec[i];
            const int nzmin = rec.x & 0xff, nzmax = (rec.x >> 8) & 0xff;
            if (c.s_nid[i] >= 0 && nzmin <= nzmax - 1) en = make_int2(c.s_nid[i], nzmin | ((nzmax - 1) << 8));
        } else {
            const int k = i - cpb, g = k / W;
            const int4 ent = c.s_ent[k];
            const uint2 rec = c.s_rec[g];
            const int nzmin = rec.x & 0xff, nzmax = (rec.x >> 8) & 0xff;
            const int lo = max(ent.z & 0xff, nzmin), hi = min((ent.z >> 8) & 0xff, nzmax - 1);
            if (lo <= hi) {
                en = make_int2(ent.y, lo | (hi << 8));
                c.et.ent[k] = make_int2(ent.x, lo | (hi << 8));
            } else c.et.ent[k] = make_int2(-1, 0xff);
        }
        c.nt.ent[i] = en;
    }
    __syncthreads();
    for (int i = tid; i < y.nn; i += nthr) dedup_scan(c.nt, i);
    for (int i = nthr - 1 - tid; i < y.ne; i += nthr) dedup_scan(c.et, i);
    __syncthreads();
    for (int i = tid; i < y.nn; i += nthr) dedup_slot(c.nt, i);
    for (int i = nthr - 1 - tid; i < y.ne; i += nthr) dedup_slot(c.et, i);
    __syncthreads();
    return c;
}

// decode of the calling thread's (node, layer)
__device__ __forceinline__ NodeThread node_thread_s(const MeshDev& m, const NodeRange& r, const NodeCtx& c)
{
    NodeThread t;
    const ColThread ct = col_thread(m);
    t.nz0 = ct.nz0;
    const int g = ct.g;
    t.n = 0; t.nzmin = 1; t.nzmax = 0;
    t.pad_lo = 0; t.pad_hi = 255; t.self_lo = 1; t.self_hi = 0; t.deg = 0;
    t.active = g < r.cpb && c.s_nid[g < r.cpb ? g : 0] >= 0;
    if (t.active) {
        t.n = c.s_nid[g];
        const uint2 rec = c.s_rec[g];
        t.nzmin = rec.x & 0xff; t.nzmax = (rec.x >> 8) & 0xff;
        t.pad_lo = (rec.x >> 16) & 0xff; t.pad_hi = rec.x >> 24;
        t.self_lo = rec.y & 0xff; t.self_hi = (rec.y >> 8) & 0xff; t.deg = (rec.y >> 16) & 0xff;
    }
    return t;
}

// =============================================================================================
// N1 staged: LO solution + vertical antidiffusive flux (contract of k_node_lo)
//   node class  : ttf[t]                        (self + neighbours)
//   edge class  : Q
//   own columns : ttfAB[t], w, we, area, zbar (nl) and Z; PPM adds hnode, hnode_new
//   own level   : areasvol, hnode, hnode_new    (plain loads before the wait)
// =============================================================================================
template <int VER, int TB>
__host__ __device__ inline uint32_t n1_own_slot(int L, int nl)
{
    return (uint32_t)TB * slot_bytes(L) + 4 * slot_bytes(nl) + slot_bytes(L) + (VER == VER_PPM ? 2 * slot_bytes(L) : 0);
}
template <int VER, int TB>
__host__ __device__ inline NodeLayout n1_layout(int L, int nl, int cpb, int ell_w, int ns_max, int es_max)
{
    return node_layout(cpb, ell_w, ns_max, es_max, (uint32_t)TB * slot_bytes(L), slot_bytes(L), n1_own_slot<VER, TB>(L, nl),
                       (uint32_t)(TB * cpb * L * 8));
}

template <int VER, int TB, bool ISSUE>
__device__ __forceinline__ uint32_t n1_items(const MeshDev& m, const Chunk<TB>& b, const NodeLayout& y, unsigned char* smem, const NodeCtx& c, int cpb)
{
    const int L = m.L, nl = m.nl, tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t s0 = smem_u32(smem), SC = slot_bytes(L), SN = slot_bytes(nl);
    uint32_t bytes = 0;
    for (int i = tid; i < y.nn; i += nthr) {
        if (!c.nt.lead[i]) continue;
        const int id = c.nt.ent[i].x, k0 = (c.nt.ur[i] & 0xff) - 1, k1 = (c.nt.ur[i] >> 8) - 1;
        const uint32_t sl = s0 + y.node + (uint32_t)c.nt.slot[i] * y.node_slot;
#pragma unroll
        for (int t = 0; t < TB; ++t) bytes += stage_col<ISSUE>(sl + t * SC, b.ttf[t] + (size_t)id * L, 1, k0, k1, c.bar);
    }
    for (int i = nthr - 1 - tid; i < y.ne; i += nthr) {
        if (!c.et.lead[i]) continue;
        const int id = c.et.ent[i].x, k0 = (c.et.ur[i] & 0xff) - 1, k1 = (c.et.ur[i] >> 8) - 1;
        bytes += stage_col<ISSUE>(s0 + y.edge + (uint32_t)c.et.slot[i] * y.edge_slot, m.Q + (size_t)id * L, 1, k0, k1, c.bar);
    }
    constexpr int NOWN = TB + 5 + (VER == VER_PPM ? 2 : 0);
    for (int i = (tid + nthr / 2) % nthr; i < cpb * NOWN; i += nthr) {
        const int g = i / NOWN, a = i - g * NOWN;
        const int n = c.s_nid[g];
        if (n < 0) continue;
        const uint2 rec = c.s_rec[g];
        const int nzmin = rec.x & 0xff, nzmax = (rec.x >> 8) & 0xff;
        const uint32_t sl = s0 + y.own + (uint32_t)g * y.own_slot;
        const int kl0 = nzmin - 1, kl1 = nzmax - 2, kn1 = nzmax - 1;   // layers nzmin..nzmax-1, interfaces nzmin..nzmax
        if (a < TB) bytes += stage_col<ISSUE>(sl + a * SC, b.ttfAB[a] + (size_t)n * L, 1, kl0, kl1, c.bar);
        else if (a == TB) bytes += stage_col<ISSUE>(sl + TB * SC, m.w + (size_t)n * nl, 1, kl0, kn1, c.bar);
        else if (a == TB + 1) bytes += stage_col<ISSUE>(sl + TB * SC + SN, m.we + (size_t)n * nl, 1, kl0, kn1, c.bar);
        else if (a == TB + 2) bytes += stage_col<ISSUE>(sl + TB * SC + 2 * SN, m.area + (size_t)n * nl, 1, kl0, kn1, c.bar);
        else if (a == TB + 3) bytes += stage_col<ISSUE>(sl + TB * SC + 3 * SN, m.zbar3d + (size_t)n * nl, 1, kl0, kn1, c.bar);
        else if (a == TB + 4) bytes += stage_col<ISSUE>(sl + TB * SC + 4 * SN, m.Z3d + (size_t)n * L, 1, kl0, kl1, c.bar);
        else if (a == TB + 5) bytes += stage_col<ISSUE>(sl + (TB + 1) * SC + 4 * SN, m.hnode + (size_t)n * L, 1, kl0, kl1, c.bar);
        else bytes += stage_col<ISSUE>(sl + (TB + 2) * SC + 4 * SN, m.hnode_new + (size_t)n * L, 1, kl0, kl1, c.bar);
    }
    return bytes;
}

template <int VER, int TB>
__global__ void __launch_bounds__(kBlock, ADV_N1S_MINB) k_node_lo_s(MeshDev m, Chunk<TB> b, StageRange sr, double dt)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const NodeRange& r = sr.r;
    const int L = m.L, nl = m.nl, tid = threadIdx.x, nthr = blockDim.x, W = m.ell_w;
    const NodeLayout y = n1_layout<VER, TB>(L, nl, r.cpb, W, sr.ns_max, sr.es_max);
    const NodeCtx c = node_prologue(m, r, y, smem);
    const NodeThread th = node_thread_s(m, r, c);
    const int g = tid < r.cpb * L ? (int)((tid * m.div_magic) >> 20) : 0;
    const int n = th.n, nz0 = th.nz0, nz = nz0 + 1, nzmin = th.nzmin, nzmax = th.nzmax;
    const bool active = th.active;
    const bool valid = active && nz >= nzmin && nz <= nzmax - 1;
    const unsigned oL = (unsigned)n * L + nz0;
    const size_t cN = (size_t)n * nl;
    // own-level operands: plain loads in flight across the barrier wait
    double av = 1.0, hn = 0.0, hnn = 1.0;
    if (valid) { av = __ldg(&m.areasvol[cN + nz0]); hn = __ldg(&m.hnode[oL]); hnn = __ldg(&m.hnode_new[oL]); }
    {
        const uint32_t bytes = n1_items<VER, TB, false>(m, b, y, smem, c, r.cpb);
        if (bytes) mbar_arrive_tx(c.bar, bytes); else mbar_arrive(c.bar);
        if (bytes) n1_items<VER, TB, true>(m, b, y, smem, c, r.cpb);
    }
    mbar_wait(c.bar, 0);

    const uint32_t SC = slot_bytes(L), SN = slot_bytes(nl);
    const unsigned char* own = smem + y.own + (uint32_t)g * y.own_slot;
    const unsigned char* self = smem + y.node + (uint32_t)c.nt.slot[g] * y.node_slot;
    double* s_flo = reinterpret_cast<double*>(smem + y.extra);   // [TB][nthr]: LO vertical flux at the top interface

    // ---- horizontal LO gather: ordered accumulation (oce_adv_tra_driver.F90:142-201) --------------------
    double tn[TB], losum[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { tn[t] = 0.0; losum[t] = 0.0; }
    if (valid) {
#pragma unroll
        for (int t = 0; t < TB; ++t) tn[t] = col_ptr(self + t * SC, b.ttf[t] + (size_t)n * L)[nz0];
        const int4* ents = c.s_ent + g * W;
        const unsigned short* nsl = c.nt.slot + r.cpb + g * W;
        const unsigned short* esl = c.et.slot + g * W;
        for (int j = 0; j < th.deg; ++j) {
            const int4 ent = ents[j];
            const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
            if (nz < lo || nz > hi) continue;
            const bool second = (ent.z >> 16) & 1;
            const double q = col_ptr(smem + y.edge + (uint32_t)esl[j] * y.edge_slot, m.Q + (size_t)ent.x * L)[nz0];
            const unsigned char* np = smem + y.node + (uint32_t)nsl[j] * y.node_slot;
            const double aq = fabs(q), qp = q + aq, qm = q - aq;
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const double to = col_ptr(np + t * SC, b.ttf[t] + (size_t)ent.y * L)[nz0];
                if (!second) losum[t] = losum[t] + hor_lo(tn[t], to, qp, qm);        // driver :175
                else losum[t] = losum[t] - hor_lo(to, tn[t], qp, qm);                // driver :188
            }
        }
    }

    // ---- vertical fluxes at the thread's top interface, stencils read from the staged columns -------
    double flo_top[TB], adfv_top[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { flo_top[t] = 0.0; adfv_top[t] = 0.0; }
    if (active && nz >= nzmin && nz <= nzmax) {
        ColV cv;
        const double* s_w = col_ptr(own + TB * SC, m.w + cN);
        const double* s_we = col_ptr(own + TB * SC + SN, m.we + cN);
        cv.area = col_ptr(own + TB * SC + 2 * SN, m.area + cN);
        cv.zbar = col_ptr(own + TB * SC + 3 * SN, m.zbar3d + cN);
        cv.Z = col_ptr(own + TB * SC + 4 * SN, m.Z3d + (size_t)n * L);
        cv.hnode = col_ptr(own + (TB + 1) * SC + 4 * SN, m.hnode + (size_t)n * L);       // PPM only
        cv.hnode_new = col_ptr(own + (TB + 2) * SC + 4 * SN, m.hnode_new + (size_t)n * L);
        cv.nzmin = nzmin; cv.nzmax = nzmax; cv.dt = dt;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            cv.ttf = col_ptr(self + t * SC, b.ttf[t] + (size_t)n * L); cv.w = s_we; cv.num_ord = 0.0;
            const double fe = ver_upw1(cv, nz, 0.0);                     // driver :235
            double flo = fe;
            if (m.use_wsplit) { cv.w = s_w; flo = ver_upw1(cv, nz, 0.0); }  // driver :333
            cv.ttf = col_ptr(own + t * SC, b.ttfAB[t] + (size_t)n * L); cv.w = s_w; cv.num_ord = b.pv[t];
            flo_top[t] = fe;
            adfv_top[t] = ver_flux<VER>(cv, nz, flo);                    // driver :363-379
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) s_flo[t * nthr + tid] = flo_top[t];
    __syncthreads();
    if (active) {
        stv<TB>(b.adf_v + (cN + nz0) * TB, adfv_top);
        if (nz0 == L - 1) {                                  // interface nl is always the (zero) bottom
            double z[TB];
#pragma unroll
            for (int t = 0; t < TB; ++t) z[t] = 0.0;
            stv<TB>(b.adf_v + (cN + L) * TB, z);
        }
    }
    if (!valid) return;
    const bool has_below = nz0 + 1 < L;
    const double r_av = 1.0 / av, r_hnn = 1.0 / hnn;
    double lo_out[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const double flo_bot = has_below ? s_flo[t * nthr + tid + 1] : 0.0;
        const double fv = flo_top[t] - flo_bot;                                          // fv(nz)-fv(nz+1)
        const double num = tn[t] * hn + div_rcp((losum[t] + fv) * dt, av, r_av);
        lo_out[t] = div_rcp(num, hnn, r_hnn);                                            // driver :249
    }
    stv<TB>(b.lo + (size_t)oL * TB, lo_out);
}

// =============================================================================================
// K2 staged: FCT bounds, P+/P- and R+/R- (contract of k_fct_bounds)
//   node class : lo (TB-interleaved), ttf[t]      edge class : adf_h (TB-interleaved)
//   own level  : adf_v(nz), adf_v(nz+1), areasvol, hnode_new
// =============================================================================================
template <int TB>
__host__ __device__ inline NodeLayout k2_layout(int L, int cpb, int ell_w, int ns_max, int es_max)
{
    return node_layout(cpb, ell_w, ns_max, es_max, slot_bytes(L * TB) + (uint32_t)TB * slot_bytes(L), slot_bytes(L * TB), 0,
                       (uint32_t)(2 * TB * cpb * L * 8));
}
template <int TB, bool ISSUE>
__device__ __forceinline__ uint32_t k2_items(const MeshDev& m, const Chunk<TB>& b, const NodeLayout& y, unsigned char* smem, const NodeCtx& c)
{
    const int L = m.L, tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t s0 = smem_u32(smem), SC = slot_bytes(L), ST = slot_bytes(L * TB);
    uint32_t bytes = 0;
    for (int i = tid; i < y.nn; i += nthr) {
        if (!c.nt.lead[i]) continue;
        const int id = c.nt.ent[i].x, k0 = (c.nt.ur[i] & 0xff) - 1, k1 = (c.nt.ur[i] >> 8) - 1;
        const uint32_t sl = s0 + y.node + (uint32_t)c.nt.slot[i] * y.node_slot;
        bytes += stage_col<ISSUE>(sl, b.lo + (size_t)id * L * TB, TB, k0, k1, c.bar);
#pragma unroll
        for (int t = 0; t < TB; ++t) bytes += stage_col<ISSUE>(sl + ST + t * SC, b.ttf[t] + (size_t)id * L, 1, k0, k1, c.bar);
    }
    for (int i = nthr - 1 - tid; i < y.ne; i += nthr) {
        if (!c.et.lead[i]) continue;
        const int id = c.et.ent[i].x, k0 = (c.et.ur[i] & 0xff) - 1, k1 = (c.et.ur[i] >> 8) - 1;
        bytes += stage_col<ISSUE>(s0 + y.edge + (uint32_t)c.et.slot[i] * y.edge_slot, b.adf_h + (size_t)id * L * TB, TB, k0, k1, c.bar);
    }
    return bytes;
}

template <int TB>
__global__ void __launch_bounds__(kBlock, ADV_K2S_MINB) k_fct_bounds_s(MeshDev m, Chunk<TB> b, StageRange sr, double dt)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const NodeRange& r = sr.r;
    const int L = m.L, nl = m.nl, tid = threadIdx.x, nthr = blockDim.x, W = m.ell_w;
    const NodeLayout y = k2_layout<TB>(L, r.cpb, W, sr.ns_max, sr.es_max);
    const NodeCtx c = node_prologue(m, r, y, smem);
    const NodeThread th = node_thread_s(m, r, c);
    const int g = tid < r.cpb * L ? (int)((tid * m.div_magic) >> 20) : 0;
    const int n = th.n, nz0 = th.nz0, nz = nz0 + 1;
    const bool valid = th.active && nz >= th.nzmin && nz <= th.nzmax - 1;
    const unsigned oL = (unsigned)n * L + nz0;
    double av = 1.0, hnn = 1.0, vt[TB], vb[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { vt[t] = 0.0; vb[t] = 0.0; }
    if (valid) {
        const size_t cN = (size_t)n * nl + nz0;
        ldv<TB>(b.adf_v + cN * TB, vt);
        ldv<TB>(b.adf_v + (cN + 1) * TB, vb);
        av = __ldg(&m.areasvol[cN]); hnn = __ldg(&m.hnode_new[oL]);
    }
    {
        const uint32_t bytes = k2_items<TB, false>(m, b, y, smem, c);
        if (bytes) mbar_arrive_tx(c.bar, bytes); else mbar_arrive(c.bar);
        if (bytes) k2_items<TB, true>(m, b, y, smem, c);
    }
    mbar_wait(c.bar, 0);

    const uint32_t SC = slot_bytes(L), ST = slot_bytes(L * TB);
    double* sm = reinterpret_cast<double*>(smem + y.extra);   // [2*TB][nthr]: tvert_max, tvert_min
    double tmax[TB], tmin[TB], pp[TB], pn[TB], lo_n[TB];
    if (valid) {
        const unsigned char* self = smem + y.node + (uint32_t)c.nt.slot[g] * y.node_slot;
        const bool padded = nz < th.pad_lo || nz > th.pad_hi;   // some element of the cluster is dry at nz
        const bool self_in = nz >= th.self_lo && nz <= th.self_hi;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            lo_n[t] = col_ptr(self, b.lo + (size_t)n * L * TB)[nz0 * TB + t];
            const double tnv = col_ptr(self + ST + t * SC, b.ttf[t] + (size_t)n * L)[nz0];
            tmax[t] = padded ? -1.0e3 : -CUDART_INF;        // bignumber, oce_adv_tra_fct.F90:100,159-176
            tmin[t] = padded ? 1.0e3 : CUDART_INF;
            const double hi2 = dmax(lo_n[t], tnv), lo2 = dmin(lo_n[t], tnv);              // a1 :129-130
            tmax[t] = (self_in && hi2 > tmax[t]) ? hi2 : tmax[t];
            tmin[t] = (self_in && lo2 < tmin[t]) ? lo2 : tmin[t];
            pp[t] = 0.0 + (dmax(0.0, vt[t]) + dmax(0.0, -vb[t]));                          // fct :291
            pn[t] = 0.0 + (dmin(0.0, vt[t]) + dmin(0.0, -vb[t]));                          // fct :292
        }
        const int4* ents = c.s_ent + g * W;
        const unsigned short* nsl = c.nt.slot + r.cpb + g * W;
        const unsigned short* esl = c.et.slot + g * W;
        for (int j = 0; j < th.deg; ++j) {
            const int4 ent = ents[j];
            const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
            if (nz < lo || nz > hi) continue;
            const bool second = (ent.z >> 16) & 1;
            const unsigned char* np = smem + y.node + (uint32_t)nsl[j] * y.node_slot;
            const double* fp = col_ptr(smem + y.edge + (uint32_t)esl[j] * y.edge_slot, b.adf_h + (size_t)ent.x * L * TB) + nz0 * TB;
            const double* lp = col_ptr(np, b.lo + (size_t)ent.y * L * TB) + nz0 * TB;
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const double lo_o = lp[t], t_o = col_ptr(np + ST + t * SC, b.ttf[t] + (size_t)ent.y * L)[nz0];
                const double hi2 = dmax(lo_o, t_o), lo2 = dmin(lo_o, t_o);
                tmax[t] = hi2 > tmax[t] ? hi2 : tmax[t];                  // a2 :166, a3 :209
                tmin[t] = lo2 < tmin[t] ? lo2 : tmin[t];
                const double a = second ? -fp[t] : fp[t];                 // fct :342,:346 / :360,:364
                pp[t] = pp[t] + dmax(0.0, a);
                pn[t] = pn[t] + dmin(0.0, a);
            }
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            sm[(2 * t) * nthr + tid] = tmax[t];
            sm[(2 * t + 1) * nthr + tid] = tmin[t];
        }
    }
    __syncthreads();
    if (!valid) return;
    const double r_av = 1.0 / av, r_hnn = 1.0 / hnn;
    const bool edge_layer = (nz == th.nzmin) || (nz == th.nzmax - 1);   // :233-234, :245-247
    double* out = b.pm + (size_t)oL * TB * 2;
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        double vmax = tmax[t], vmin = tmin[t];
        if (!edge_layer) {                                               // :238-241
            const double* smax = sm + (2 * t) * nthr + tid;
            const double* smin = sm + (2 * t + 1) * nthr + tid;
            vmax = dmax(dmax(smax[-1], vmax), smax[1]);
            vmin = dmin(dmin(smin[-1], vmin), smin[1]);
        }
        const double inc_max = vmax - lo_n[t], inc_min = vmin - lo_n[t];
        const double fp = div_rcp(div_rcp(pp[t] * dt, av, r_av), hnn, r_hnn) + 1e-16;   // b2 :399
        const double fm = div_rcp(div_rcp(pn[t] * dt, av, r_av), hnn, r_hnn) - 1e-16;   // :401
        reinterpret_cast<double2*>(out)[t] = make_double2(dmin(1.0, inc_max / fp), dmin(1.0, inc_min / fm));
    }
}

// =============================================================================================
// K3 staged: limit the antidiffusive fluxes, accumulate the tendencies (contract of k_fct_update)
//   node class : pm = {R+, R-} x TB      edge class : adf_h
//   own level  : adf_v(nz), adf_v(nz+1), lo, ttf, del_ttf_*, areasvol, hnode, hnode_new
// =============================================================================================
template <int TB>
__host__ __device__ inline NodeLayout k3_layout(int L, int cpb, int ell_w, int ns_max, int es_max)
{
    return node_layout(cpb, ell_w, ns_max, es_max, slot_bytes(L * TB * 2), slot_bytes(L * TB), 0, 0);
}
template <int TB, bool ISSUE>
__device__ __forceinline__ uint32_t k3_items(const MeshDev& m, const Chunk<TB>& b, const NodeLayout& y, unsigned char* smem, const NodeCtx& c)
{
    const int L = m.L, tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t s0 = smem_u32(smem);
    uint32_t bytes = 0;
    for (int i = tid; i < y.nn; i += nthr) {
        if (!c.nt.lead[i]) continue;
        const int id = c.nt.ent[i].x, k0 = (c.nt.ur[i] & 0xff) - 1, k1 = (c.nt.ur[i] >> 8) - 1;
        bytes += stage_col<ISSUE>(s0 + y.node + (uint32_t)c.nt.slot[i] * y.node_slot, b.pm + (size_t)id * L * TB * 2, TB * 2, k0, k1, c.bar);
    }
    for (int i = nthr - 1 - tid; i < y.ne; i += nthr) {
        if (!c.et.lead[i]) continue;
        const int id = c.et.ent[i].x, k0 = (c.et.ur[i] & 0xff) - 1, k1 = (c.et.ur[i] >> 8) - 1;
        bytes += stage_col<ISSUE>(s0 + y.edge + (uint32_t)c.et.slot[i] * y.edge_slot, b.adf_h + (size_t)id * L * TB, TB, k0, k1, c.bar);
    }
    return bytes;
}

template <int TB>
__global__ void __launch_bounds__(kBlock, ADV_K3S_MINB) k_fct_update_s(MeshDev m, Chunk<TB> b, StageRange sr, double dt)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const NodeRange& r = sr.r;
    const int L = m.L, nl = m.nl, tid = threadIdx.x, W = m.ell_w;
    const NodeLayout y = k3_layout<TB>(L, r.cpb, W, sr.ns_max, sr.es_max);
    const NodeCtx c = node_prologue(m, r, y, smem);
    const NodeThread th = node_thread_s(m, r, c);
    const int g = tid < r.cpb * L ? (int)((tid * m.div_magic) >> 20) : 0;
    const int n = th.n, nz0 = th.nz0, nz = nz0 + 1;
    const bool valid = th.active && nz >= th.nzmin && nz <= th.nzmax - 1;
    const bool owned = n < m.N;
    const unsigned oL = (unsigned)n * L + nz0;
    const size_t cN = (size_t)n * nl + nz0;
    double av = 1.0, hn = 0.0, hnn = 0.0, dh[TB], dv[TB], vt[TB], vb[TB], lo_n[TB], tn[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { dh[t] = dv[t] = vt[t] = vb[t] = lo_n[t] = tn[t] = 0.0; }
    if (valid) {
        av = __ldg(&m.areasvol[cN]);
#pragma unroll
        for (int t = 0; t < TB; ++t) dh[t] = b.dttf_h[t][oL];
        if (owned) {
            ldv<TB>(b.adf_v + cN * TB, vt);
            ldv<TB>(b.adf_v + (cN + 1) * TB, vb);
            ldv<TB>(b.lo + (size_t)oL * TB, lo_n);
#pragma unroll
            for (int t = 0; t < TB; ++t) { tn[t] = __ldg(&b.ttf[t][oL]); dv[t] = b.dttf_v[t][oL]; }
            hn = __ldg(&m.hnode[oL]); hnn = __ldg(&m.hnode_new[oL]);
        }
    }
    {
        const uint32_t bytes = k3_items<TB, false>(m, b, y, smem, c);
        if (bytes) mbar_arrive_tx(c.bar, bytes); else mbar_arrive(c.bar);
        if (bytes) k3_items<TB, true>(m, b, y, smem, c);
    }
    mbar_wait(c.bar, 0);
    if (!valid) return;

    const double r_av = 1.0 / av;
    const double2* self = reinterpret_cast<const double2*>(smem + y.node + (uint32_t)c.nt.slot[g] * y.node_slot) + nz0 * TB;   // pm: 16-byte aligned
    double pk[TB], mk[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) { const double2 v = self[t]; pk[t] = v.x; mk[t] = v.y; }
    if (owned) {
        const bool above = nz > th.nzmin, below = nz + 1 <= th.nzmax - 1;
        const bool has_below = nz0 + 1 < L;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            double pa = 1.0, ma = 1.0, pb = 1.0, mb = 1.0;
            if (above) { const double2 v = self[t - TB]; pa = v.x; ma = v.y; }
            if (below) { const double2 v = self[t + TB]; pb = v.x; mb = v.y; }
            const double fv_top = limit_v(vt[t], nz, th.nzmin, th.nzmax, pa, ma, pk[t], mk[t]);
            const double fv_bot = has_below ? limit_v(vb[t], nz + 1, th.nzmin, th.nzmax, pk[t], mk[t], pb, mb) : 0.0;
            double d = dv[t];
            d = d - tn[t] * hn + lo_n[t] * hnn;                           // driver :535
            d = d + div_rcp((fv_top - fv_bot) * dt, av, r_av);            // driver :556
            b.dttf_v[t][oL] = d;
        }
    }
    const int4* ents = c.s_ent + g * W;
    const unsigned short* nsl = c.nt.slot + r.cpb + g * W;
    const unsigned short* esl = c.et.slot + g * W;
    for (int j = 0; j < th.deg; ++j) {
        const int4 ent = ents[j];
        const int lo = ent.z & 0xff, hi = (ent.z >> 8) & 0xff;
        if (nz < lo || nz > hi) continue;
        const bool second = (ent.z >> 16) & 1;
        const double2* op = reinterpret_cast<const double2*>(smem + y.node + (uint32_t)nsl[j] * y.node_slot) + nz0 * TB;
        const double* fp = col_ptr(smem + y.edge + (uint32_t)esl[j] * y.edge_slot, b.adf_h + (size_t)ent.x * L * TB) + nz0 * TB;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const double ff = fp[t];
            const double2 o = op[t];
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

}  // namespace adv
