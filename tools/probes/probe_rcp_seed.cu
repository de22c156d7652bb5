// Measures the relative error of the MUFU.RCP64H seed (rcp.approx.ftz.f64) on the device,
// to justify the Newton/cubic step counts in csrc/kem_math.cuh.  Run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe tools/probes/probe_rcp_seed.cu && /tmp/probe
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double* worst, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double w = 0.0, w2 = 0.0;
    for (int j = i; j < n; j += gridDim.x * blockDim.x) {
        double b = 1.0 + (double)j / (double)n;          // mantissa sweep over [1, 2)
        b *= (j & 1) ? 3.7e11 : 1.3e-7;
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
        double e = fabs(fma(-b, r, 1.0));
        w = e > w ? e : w;
        double q;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(b));
        double e2 = fabs(fma(-b * q, q, 1.0));          // 1 - b q^2 ~ 2 * relative error of q
        w2 = e2 > w2 ? e2 : w2;
    }
    // block max
    __shared__ double s[256];
    s[threadIdx.x] = w;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) worst[blockIdx.x] = s[0];
    __syncthreads();
    s[threadIdx.x] = w2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) worst[gridDim.x + blockIdx.x] = s[0];
}
int main()
{
    const int blocks = 1024, n = 1 << 28;
    double* d;
    cudaMalloc(&d, 2 * blocks * sizeof(double));
    k<<<blocks, 256>>>(d, n);
    double h[2 * blocks];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double w = 0, w2 = 0;
    for (int i = 0; i < blocks; ++i) { w = h[i] > w ? h[i] : w; w2 = h[blocks + i] > w2 ? h[blocks + i] : w2; }
    printf("rcp.approx.ftz.f64   max |1 - b*r|   = %.4e = 2^%.2f over %d samples\n", w, log2(w), n);
    printf("rsqrt.approx.ftz.f64 max |1 - b*q*q| = %.4e = 2^%.2f (relative error of q: half of it)\n", w2, log2(w2));
    return cudaGetLastError() != cudaSuccess;
}
