// tools/bw_probe3.cu -- prototype: persistent CTAs, operands staged in shared memory by 1-D bulk
// TMA copies (cp.async.bulk + mbarrier), S-stage pipeline.  Same work as bw_probe2 "meta+gather+fp64".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// item = EPB consecutive edges; per edge: 2 x (L*32 B) streamed + 8 x (L*8 B) gathered node columns
template <int S, int NFP>
__global__ void __launch_bounds__(256) k_tma(const double* __restrict__ a0, const double* __restrict__ a1, const int4* __restrict__ meta,
                                             const double* __restrict__ nodes, double* __restrict__ out, int L, int E, int epb, unsigned magic, int nitems)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[S];
    const int colb = L * 8;                       // bytes of one 8-byte column
    const int edge_bytes = 2 * 4 * colb + 8 * colb;
    const int stage_bytes = epb * edge_bytes;
    const int g = (threadIdx.x * magic) >> 20, nz0 = threadIdx.x - g * L;
    const size_t nst = (size_t)(E / 3 + 2) * L;
    if (threadIdx.x == 0) for (int s = 0; s < S; ++s) mbar_init(&full[s], epb);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    auto issue = [&](int item, int stage, int4 em) {     // called by threads < epb
        const int e = item * epb + threadIdx.x;
        if (e >= E) { mbar_expect(&full[stage], 0); return; }
        unsigned char* base = smem + (size_t)stage * stage_bytes + (size_t)threadIdx.x * edge_bytes;
        mbar_expect(&full[stage], edge_bytes);
        bulk_g2s(base, a0 + (size_t)e * L * 4, 4 * colb, &full[stage]);
        bulk_g2s(base + 4 * colb, a1 + (size_t)e * L * 4, 4 * colb, &full[stage]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bulk_g2s(base + (8 + 2 * k) * colb, nodes + k * nst + (size_t)em.x * L, colb, &full[stage]);
            bulk_g2s(base + (9 + 2 * k) * colb, nodes + k * nst + (size_t)em.y * L, colb, &full[stage]);
        }
    };
    auto load_meta = [&](int item) {
        const int e = item * epb + threadIdx.x;
        return (threadIdx.x < epb && e < E) ? __ldg(&meta[e]) : make_int4(0, 0, 0, 0);
    };
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_n = first < nitems ? (nitems - first + stride - 1) / stride : 0;
    // prologue: fill S-1 stages
    for (int s = 0; s < S - 1 && s < my_n; ++s) {
        const int4 em = load_meta(first + s * stride);
        if (threadIdx.x < epb) issue(first + s * stride, s, em);
    }
    int4 em_next = (S - 1 < my_n) ? load_meta(first + (S - 1) * stride) : make_int4(0, 0, 0, 0);
    for (int it = 0; it < my_n; ++it) {
        const int stage = it % S;
        // issue item it+S-1 into the stage freed at the end of the previous iteration
        if (it + S - 1 < my_n) {
            if (threadIdx.x < epb) issue(first + (it + S - 1) * stride, (it + S - 1) % S, em_next);
            if (it + S < my_n) em_next = load_meta(first + (it + S) * stride);
        }
        mbar_wait(&full[stage], (it / S) & 1);
        const int e = (first + it * stride) * epb + g;
        if (e < E && g < epb) {
            const unsigned char* base = smem + (size_t)stage * stage_bytes + (size_t)g * edge_bytes;
            const double2* p0 = reinterpret_cast<const double2*>(base) + nz0 * 2;
            const double2* p1 = reinterpret_cast<const double2*>(base + 4 * colb) + nz0 * 2;
            const double2 x0 = p0[0], y0 = p0[1], x1 = p1[0], y1 = p1[1];
            double s0 = x0.x + y0.x, s1 = x1.x + y1.x;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const double t = reinterpret_cast<const double*>(base + (8 + k) * colb)[nz0]; s0 += t; s1 -= t; }
            s0 += x0.y * y0.y; s1 += x1.y * y1.y;
#pragma unroll
            for (int i = 0; i < NFP; ++i) { s0 = __dmul_rn(s0, 1.0000001); s0 = __dadd_rn(s0, 1e-9); s1 = __dmul_rn(s1, 0.9999999); s1 = __dadd_rn(s1, 1e-9); }
            reinterpret_cast<double2*>(out)[(size_t)e * L + nz0] = make_double2(s0, s1);
        }
        __syncthreads();   // everyone is done with `stage` before it is refilled
    }
}
int main()
{
    const int L = 70, E = 1127307;
    const size_t n4 = (size_t)E * L * 4;
    double *a0, *a1, *out, *nodes; int4* meta;
    CK(cudaMalloc(&a0, n4 * 8)); CK(cudaMalloc(&a1, n4 * 8)); CK(cudaMalloc(&out, n4 * 4));
    CK(cudaMalloc(&nodes, (size_t)(E / 3 + 2) * L * 4 * 8)); CK(cudaMalloc(&meta, (size_t)E * 16));
    CK(cudaMemset(a0, 0, n4 * 8)); CK(cudaMemset(a1, 0, n4 * 8)); CK(cudaMemset(out, 0, n4 * 4));
    CK(cudaMemset(nodes, 0, (size_t)(E / 3 + 2) * L * 4 * 8));
    {
        std::vector<int4> h(E);
        const int nx = 613;
        for (int e = 0; e < E; ++e) { int n = e / 3; int k = e % 3; int o = k == 0 ? n + 1 : k == 1 ? n + nx : n + nx + 1; if (o > E / 3) o = n; h[e] = make_int4(n, o, 0, 0); }
        CK(cudaMemcpy(meta, h.data(), (size_t)E * 16, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned magic = ((1u << 20) + L - 1) / L;
    const double bytes = (double)E * L * (64 + 16);
    auto timeit = [&](const char* name, auto kern, int S, int epb, int ctas_per_sm) {
        const int nitems = (E + epb - 1) / epb;
        const size_t sm = (size_t)S * epb * (16 * L * 8);
        if (sm * ctas_per_sm > 225 * 1024) { printf("%-28s S=%d epb=%d occ=%d: smem %zu too big\n", name, S, epb, ctas_per_sm, sm); return; }
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        const int grid = 148 * ctas_per_sm, nthr = ((epb * L + 31) / 32) * 32;
        for (int i = 0; i < 2; ++i) kern<<<grid, nthr, sm>>>(a0, a1, meta, nodes, out, L, E, epb, magic, nitems);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 5;
        for (int i = 0; i < reps; ++i) kern<<<grid, nthr, sm>>>(a0, a1, meta, nodes, out, L, E, epb, magic, nitems);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("%-28s S=%d epb=%d occ=%d smem/CTA=%3zuKB  %8.3f ms  %7.1f GB/s\n", name, S, epb, ctas_per_sm, sm / 1024, ms, bytes / ms / 1e6);
    };
    for (int epb : {1, 2, 3}) for (int occ : {2, 3, 4, 6, 8}) { timeit("tma fp64x50 S=2", k_tma<2, 50>, 2, epb, occ); }
    for (int epb : {1, 2, 3}) for (int occ : {2, 3, 4}) { timeit("tma fp64x50 S=3", k_tma<3, 50>, 3, epb, occ); }
    for (int epb : {1, 2}) for (int occ : {2, 4, 6}) { timeit("tma fp64x50 S=4", k_tma<4, 50>, 4, epb, occ); }
    for (int epb : {1, 3}) for (int occ : {3, 4, 6}) { timeit("tma fp64x0 S=2", k_tma<2, 0>, 2, epb, occ); }
    return 0;
}
