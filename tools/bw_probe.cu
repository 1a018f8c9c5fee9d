// tools/bw_probe.cu -- HBM bandwidth probe for the access shapes of the advection kernels
// (experiment harness, not product code).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// (a) E1-like: CTA = epb edges x L levels, thread reads 32 B of NA arrays, writes 16 B
template <int NA, int WR>
__global__ void k_cols(const double* __restrict__ a0, const double* __restrict__ a1, double* __restrict__ out, int L, int E, int epb, unsigned magic)
{
    const int g = (threadIdx.x * magic) >> 20, nz0 = threadIdx.x - g * L;
    const int e = blockIdx.x * epb + g;
    if (e >= E) return;
    const size_t o = (size_t)e * L + nz0;
    double s = 0;
    {
        const double2* p = reinterpret_cast<const double2*>(a0) + o * 2;
        double2 x = __ldg(p), y = __ldg(p + 1); s += x.x + x.y + y.x + y.y;
    }
    if (NA > 1) {
        const double2* p = reinterpret_cast<const double2*>(a1) + o * 2;
        double2 x = __ldg(p), y = __ldg(p + 1); s += x.x * x.y + y.x * y.y;
    }
    if (WR) reinterpret_cast<double2*>(out)[o] = make_double2(s, s);
    else if (s == 1.2345e-300) out[0] = s;
}
// (b) flat grid-stride, 16 B per thread per iteration, persistent
__global__ void k_flat(const double2* __restrict__ a, double2* __restrict__ out, size_t n, int wr)
{
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 x = __ldg(a + i);
        if (wr) out[i] = x; else s += x.x + x.y;
    }
    if (!wr && s == 1.2345e-300) out[0] = make_double2(s, s);
}
// (c) flat non-persistent, 8 B per thread (the node-kernel shape)
__global__ void k_flat8(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c, double* __restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __ldg(a + i) + __ldg(b + i) + __ldg(c + i);
}
int main()
{
    const int L = 70, E = 1127307, epb = 3;
    const size_t n4 = (size_t)E * L * 4;
    double *a0, *a1, *out;
    CK(cudaMalloc(&a0, n4 * 8)); CK(cudaMalloc(&a1, n4 * 8)); CK(cudaMalloc(&out, n4 * 8));
    CK(cudaMemset(a0, 0, n4 * 8)); CK(cudaMemset(a1, 0, n4 * 8)); CK(cudaMemset(out, 0, n4 * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned magic = ((1u << 20) + L - 1) / L;
    auto timeit = [&](const char* name, double bytes, auto&& launch) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 10;
        for (int i = 0; i < reps; ++i) launch();
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("%-44s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6);
    };
    const int grid = (E + epb - 1) / epb, nthr = epb * L;
    const double b1 = (double)E * L * 32;
    timeit("cols 1 array rd 32B/thr, no write", b1, [&] { k_cols<1, 0><<<grid, nthr>>>(a0, a1, out, L, E, epb, magic); });
    timeit("cols 2 arrays rd, no write", 2 * b1, [&] { k_cols<2, 0><<<grid, nthr>>>(a0, a1, out, L, E, epb, magic); });
    timeit("cols 2 arrays rd + 16B write", 2.5 * b1, [&] { k_cols<2, 1><<<grid, nthr>>>(a0, a1, out, L, E, epb, magic); });
    timeit("cols 1 array rd + 16B write", 1.5 * b1, [&] { k_cols<1, 1><<<grid, nthr>>>(a0, a1, out, L, E, epb, magic); });
    for (int epb2 : {1, 2, 6, 12}) {
        char nm[64]; snprintf(nm, 64, "cols 2 arrays rd + write, epb=%d (%d thr)", epb2, epb2 * L);
        timeit(nm, 2.5 * b1, [&] { k_cols<2, 1><<<(E + epb2 - 1) / epb2, epb2 * L>>>(a0, a1, out, L, E, epb2, magic); });
    }
    const size_t n2 = n4 / 2;
    for (int bl : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
        char nm[64]; snprintf(nm, 64, "flat persistent read 16B, %d CTAs x 512", bl);
        timeit(nm, n4 * 8.0, [&] { k_flat<<<bl, 512>>>((const double2*)a0, (double2*)out, n2, 0); });
        snprintf(nm, 64, "flat persistent copy 16B, %d CTAs x 512", bl);
        timeit(nm, 2 * n4 * 8.0, [&] { k_flat<<<bl, 512>>>((const double2*)a0, (double2*)out, n2, 1); });
    }
    {
        const size_t n = n4;
        timeit("flat 8B: 3 reads + 1 write, 256 thr", 4.0 * n * 8, [&] { k_flat8<<<(unsigned)((n + 255) / 256), 256>>>(a0, a1, out, out, n); });
        timeit("flat 8B: 3 reads + 1 write, 224 thr", 4.0 * n * 8, [&] { k_flat8<<<(unsigned)((n + 223) / 224), 224>>>(a0, a1, out, out, n); });
    }
    return 0;
}
