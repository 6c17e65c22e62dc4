#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void bench(double* out, long long* cyc, double a, double b, int mode) {
    double c[6] = {0, 0, 0, 0, 0, 0};
    double f[8];
    for (int j = 0; j < 8; ++j) f[j] = a + j;
    a += threadIdx.x * 1e-3; b -= threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 512; ++i) {
        if (mode == 0 || mode == 2) {           // 3 independent DMMA chains x 2
            dmma(c[0], c[1], a, b); dmma(c[2], c[3], a, b); dmma(c[4], c[5], a, b);
            dmma(c[0], c[1], b, a); dmma(c[2], c[3], b, a); dmma(c[4], c[5], b, a);
        }
        if (mode == 1 || mode == 2) {           // 8 independent DFMA chains x 2
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fma(f[j], b, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fma(f[j], b, b);
        }
        if (mode == 3) {                        // dependent DMMA chain
            dmma(c[0], c[1], a, b); dmma(c[0], c[1], a, b); dmma(c[0], c[1], a, b);
            dmma(c[0], c[1], a, b); dmma(c[0], c[1], a, b); dmma(c[0], c[1], a, b);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0; for (int j = 0; j < 6; ++j) s += c[j]; for (int j = 0; j < 8; ++j) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// layout check: D = A(8x4) * B(4x8)
__global__ void check(double* D, const double* A, const double* B) {
    int lane = threadIdx.x;
    double d0 = 0, d1 = 0;
    double a = A[(lane >> 2) * 4 + (lane & 3)];      // A[m][k], m = lane/4, k = lane%4
    double b = B[(lane & 3) * 8 + (lane >> 2)];      // B[k][n], k = lane%4, n = lane/4
    dmma(d0, d1, a, b);
    D[(lane >> 2) * 8 + 2 * (lane & 3)] = d0;
    D[(lane >> 2) * 8 + 2 * (lane & 3) + 1] = d1;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    const char* names[] = {"DMMA 3 chains (3072 mma)", "DFMA 8 chains (8192 fma)", "DMMA+DFMA interleaved", "DMMA dependent (3072 mma)"};
    for (int warps = 1; warps <= 16; warps *= 2)
        for (int m = 0; m < 4; ++m) {
            for (int rep = 0; rep < 2; ++rep) { bench<<<148, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, m); cudaDeviceSynchronize(); }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps/SM=%2d %-28s %8lld cycles\n", warps, names[m], c);
        }
    double hA[32], hB[32], hD[64], *dA, *dB, *dD;
    for (int i = 0; i < 32; ++i) { hA[i] = i + 1; hB[i] = 0.5 * i - 3; }
    cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dD, 512);
    cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
    check<<<1, 32>>>(dD, dA, dB); cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 8; ++m) for (int n = 0; n < 8; ++n) { double r = 0; for (int k = 0; k < 4; ++k) r += hA[m * 4 + k] * hB[k * 8 + n]; maxerr = fmax(maxerr, fabs(r - hD[m * 8 + n])); }
    printf("layout check max err %g (%s)\n", maxerr, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
