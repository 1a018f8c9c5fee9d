// tools/bw_probe4.cu -- does HBM throughput depend on the NUMBER OF CONCURRENT STREAMS and on the
// contiguous chunk a CTA reads from each?  (experiment harness, not product code)
// Every CTA of T threads reads one 8-byte (or 16-byte) word per thread from each of NA arrays at the
// same offset (chunk = T*8 bytes contiguous per array per CTA) and writes one word per thread.
// All loads of a thread are independent and issued before the first use.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct Ptrs { const double* a[16]; };

template <int NA, int VEC, int CHUNKS>   // CHUNKS: consecutive chunks handled by one CTA, one after the other
__global__ void __launch_bounds__(256) k_streams(Ptrs p, double* __restrict__ out, size_t nwords)
{
    for (int c = 0; c < CHUNKS; ++c) {
        const size_t i = ((size_t)blockIdx.x * CHUNKS + c) * blockDim.x + threadIdx.x;   // word (VEC doubles) index
        if (i * VEC >= nwords) return;
        double s = 0.0;
        if (VEC == 1) {
            double v[NA];
#pragma unroll
            for (int k = 0; k < NA; ++k) v[k] = __ldg(p.a[k] + i);
#pragma unroll
            for (int k = 0; k < NA; ++k) s += v[k];
            out[i] = s;
        } else {
            double2 v[NA];
#pragma unroll
            for (int k = 0; k < NA; ++k) v[k] = __ldg(reinterpret_cast<const double2*>(p.a[k]) + i);
#pragma unroll
            for (int k = 0; k < NA; ++k) s += v[k].x + v[k].y;
            reinterpret_cast<double2*>(out)[i] = make_double2(s, s);
        }
    }
}

int main()
{
    const size_t nwords = (size_t)375769 * 70;          // one node field of the bench mesh
    Ptrs p;
    for (int k = 0; k < 16; ++k) { double* a; CK(cudaMalloc(&a, nwords * 8)); CK(cudaMemset(a, 0, nwords * 8)); p.a[k] = a; }
    double* out; CK(cudaMalloc(&out, nwords * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, int na, auto&& launch) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 10;
        for (int i = 0; i < reps; ++i) launch();
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        const double bytes = (double)nwords * 8 * (na + 1);
        printf("%-44s %7.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6);
    };
#define RUN(NA, VEC, CH, T) { char nm[96]; snprintf(nm, 96, "arrays=%2d vec=%d chunks/CTA=%d threads=%d", NA, VEC, CH, T); \
        const size_t nw = nwords / VEC; const int grid = (int)((nw + (size_t)T * CH - 1) / ((size_t)T * CH)); \
        timeit(nm, NA, [&] { k_streams<NA, VEC, CH><<<grid, T>>>(p, out, nwords); }); }
    RUN(1, 1, 1, 210) RUN(2, 1, 1, 210) RUN(4, 1, 1, 210) RUN(8, 1, 1, 210) RUN(12, 1, 1, 210) RUN(16, 1, 1, 210)
    RUN(4, 2, 1, 210) RUN(8, 2, 1, 210) RUN(16, 2, 1, 210)
    RUN(8, 1, 1, 256) RUN(16, 1, 1, 256)
    RUN(8, 1, 8, 210) RUN(16, 1, 8, 210)
    RUN(16, 1, 1, 70) RUN(16, 1, 1, 140)
    CK(cudaGetLastError());
    return 0;
}
