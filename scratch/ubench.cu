#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b, int mode) {
    double x = a, y = b, z = a + 1, w = b + 2;
    long long t0 = clock64();
    if (mode == 0) {            // dependent DFMA chain
#pragma unroll 1
        for (int i = 0; i < 256; ++i) { x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); }
    } else if (mode == 1) {     // 4 independent chains
#pragma unroll 1
        for (int i = 0; i < 256; ++i) { x = fma(x, y, y); z = fma(z, y, y); w = fma(w, y, y); a = fma(a, y, y); }
    } else if (mode == 2) {     // dependent division chain
#pragma unroll 1
        for (int i = 0; i < 256; ++i) { x = 1.0 / (x + y); }
    } else if (mode == 3) {     // dependent compare/select chain
#pragma unroll 1
        for (int i = 0; i < 256; ++i) { if (x > y) { x = y + z; } else { x = x * w; } }
    } else if (mode == 4) {    // dependent FFMA chain (fp32) for comparison
        float fx = (float)a, fy = (float)b;
#pragma unroll 1
        for (int i = 0; i < 256; ++i) { fx = fmaf(fx, fy, fy); fx = fmaf(fx, fy, fy); fx = fmaf(fx, fy, fy); fx = fmaf(fx, fy, fy); }
        x = fx;
    } else if (mode == 5) {    // 16 independent DFMA chains (throughput, 1 warp)
        double v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = a + j;
#pragma unroll 1
        for (int i = 0; i < 256; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fma(v[j], y, y);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) x += v[j];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + z + w + a;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const char* names[] = {"DFMA dependent (1024 ops)", "DFMA 4 chains (1024 ops)", "DDIV dependent (256 ops)", "DSETP+select dependent (256)", "FFMA dependent (1024 ops)", "DFMA 16 chains (4096 ops)"};
    int nops[] = {1024, 1024, 256, 256, 1024, 4096};
    for (int warps = 1; warps <= 16; warps *= 4)
    for (int m = 0; m < 6; ++m) {
        lat<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, m);
        cudaDeviceSynchronize();
        lat<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, m);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d %-32s %8lld cycles  %.1f cyc/op\n", warps, names[m], c, (double)c / nops[m]);
    }
    return 0;
}
