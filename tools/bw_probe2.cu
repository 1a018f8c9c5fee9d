// tools/bw_probe2.cu -- which ingredient of the edge-flux kernel costs HBM bandwidth?
// knobs: occupancy (dynamic smem to limit CTAs/SM), dependent index load, FP64 chain length,
// gather loads (node columns through an index)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int META, int NFP, int GATHER>
__global__ void k_e1(const double* __restrict__ a0, const double* __restrict__ a1, const int4* __restrict__ meta,
                     const double* __restrict__ nodes, double* __restrict__ out, int L, int E, int epb, unsigned magic)
{
    extern __shared__ double dummy[];
    const int g = (threadIdx.x * magic) >> 20, nz0 = threadIdx.x - g * L;
    int e = blockIdx.x * epb + g;
    if (e >= E) return;
    int4 em = make_int4(e / 3, e / 3 + 1, 0, 0);
    if (META) em = __ldg(&meta[e]);
    const size_t o = (size_t)e * L + nz0;
    const double2* p0 = reinterpret_cast<const double2*>(a0) + o * 2;
    const double2* p1 = reinterpret_cast<const double2*>(a1) + o * 2;
    double2 x0 = __ldg(p0), y0 = __ldg(p0 + 1), x1 = __ldg(p1), y1 = __ldg(p1 + 1);
    double s0 = x0.x + y0.x, s1 = x1.x + y1.x;
    if (GATHER) {
        const size_t o1 = (size_t)em.x * L + nz0, o2 = (size_t)em.y * L + nz0;
        const size_t st = (size_t)(E / 3 + 2) * L;
        double t[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { t[2 * k] = __ldg(nodes + k * st + o1); t[2 * k + 1] = __ldg(nodes + k * st + o2); }
#pragma unroll
        for (int k = 0; k < 8; ++k) { s0 += t[k]; s1 -= t[k]; }
    }
    s0 += x0.y * y0.y; s1 += x1.y * y1.y;
#pragma unroll
    for (int i = 0; i < NFP; ++i) { s0 = __dmul_rn(s0, 1.0000001) ; s0 = __dadd_rn(s0, 1e-9); s1 = __dmul_rn(s1, 0.9999999); s1 = __dadd_rn(s1, 1e-9); }
    reinterpret_cast<double2*>(out)[o] = make_double2(s0, s1);
}
int main()
{
    const int L = 70, E = 1127307, epb = 3;
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
    const int grid = (E + epb - 1) / epb, nthr = epb * L;
    const double bytes = (double)E * L * (64 + 16) ;
    auto timeit = [&](const char* name, int ctas_per_sm, auto kern) {
        // limit occupancy with dynamic shared memory
        size_t sm = ctas_per_sm >= 9 ? 0 : (size_t)(200 * 1024 / ctas_per_sm);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sm > 48 * 1024 ? sm : 48 * 1024)));
        for (int i = 0; i < 2; ++i) kern<<<grid, nthr, sm>>>(a0, a1, meta, nodes, out, L, E, epb, magic);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 5;
        for (int i = 0; i < reps; ++i) kern<<<grid, nthr, sm>>>(a0, a1, meta, nodes, out, L, E, epb, magic);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("%-40s occ %d CTA/SM  %8.3f ms  %7.1f GB/s (stream bytes only)\n", name, ctas_per_sm, ms, bytes / ms / 1e6);
    };
    for (int occ : {9, 6, 4, 3, 2}) timeit("plain", occ, k_e1<0, 0, 0>);
    for (int occ : {9, 4}) timeit("meta", occ, k_e1<1, 0, 0>);
    for (int occ : {9, 4}) timeit("fp64 x50", occ, k_e1<0, 50, 0>);
    for (int occ : {9, 4}) timeit("fp64 x100", occ, k_e1<0, 100, 0>);
    for (int occ : {9, 4}) timeit("gather", occ, k_e1<0, 0, 1>);
    for (int occ : {9, 4}) timeit("meta+gather", occ, k_e1<1, 0, 1>);
    for (int occ : {9, 6, 4, 3}) timeit("meta+gather+fp64 x50", occ, k_e1<1, 50, 1>);
    return 0;
}
