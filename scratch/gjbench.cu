#include <cstdio>
#include "../rpsmf_b200/csrc/psmf_filter.cuh"
using namespace psmf;
template <int R>
__global__ void gjkernel(long long* cyc, double* out, int variant) {
    __shared__ Smem<R> sh;
    const int tid = threadIdx.x;
    long long acc[4] = {0, 0, 0, 0};
    for (int rep = 0; rep < 50; ++rep) {
        // SPD-ish test matrix
        for (int idx = tid; idx < R * (2 * R + 1); idx += blockDim.x) {
            int i = idx / (2 * R + 1), c = idx % (2 * R + 1);
            sh.aug[0][i][c] = (c == i ? 3.0 + i : 0.0) + 0.01 * ((i * 7 + c * 3 + rep) % 11) + (c >= R ? 0.5 : 0.0);
        }
        __syncthreads();
        long long t0 = clock64();
        if (tid < GJ_THREADS) gauss_jordan_cta<R>(sh, tid);
        __syncthreads();
        long long t1 = clock64();
        acc[0] += t1 - t0;
    }
    if (tid == 0) { cyc[0] = acc[0] / 50; }
    out[tid] = sh.aug[R & 1][tid % R][R + tid % R];
}
int main() {
    long long* cyc; double* out; cudaMalloc(&cyc, 64); cudaMalloc(&out, 8192);
    const int thrs[3] = {192, 256, 480};
    for (int ti = 0; ti < 3; ++ti) {
        const int thr = thrs[ti];
        gjkernel<16><<<1, thr>>>(cyc, out, 0); cudaDeviceSynchronize();
        gjkernel<16><<<148, thr>>>(cyc, out, 0); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("R=16 threads=%d  GJ total %lld cycles (%.0f per pivot step) err=%s\n", thr, c, c / 16.0, cudaGetErrorString(cudaGetLastError()));
    }
    gjkernel<8><<<148, 256>>>(cyc, out, 0); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("R=8 GJ total %lld cycles (%.0f per step)\n", c, c / 8.0);
    return 0;
}
