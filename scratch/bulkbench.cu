// Microbenchmark of the streaming pattern: every CTA streams its private slab through a ring of smem slots
// with cp.async.bulk loads (+ optional bulk stores), one producer thread, consumers release immediately.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bload(void* d, const void* g, uint32_t n, uint64_t* b) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(g), "r"(n), "r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bstore(void* g, const void* s, uint32_t n) { asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s32(s)), "r"(n) : "memory"); }
__device__ __forceinline__ void bcommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bwaitr() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

template <int LAG>
__global__ void __launch_bounds__(64, 1) stream(char* buf, size_t slab, int chunk, int nslot, int passes, int do_store) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t full[64], done[64];
    const int tid = threadIdx.x;
    if (tid == 0) { for (int s = 0; s < nslot; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    char* base = buf + (size_t)blockIdx.x * slab;
    const int nch = (int)(slab / chunk);
    const long long total = (long long)passes * nch;
    if (tid == 0) {          // producer
        for (long long c = 0; c < total && c < nslot; ++c) { mbar_expect(&full[c], chunk); bload(sm + (size_t)c * chunk, base + (size_t)(c % nch) * chunk, chunk, &full[c]); }
        for (long long j = 0; j < total + LAG - 1; ++j) {
            if (j < total) {
                const int slot = (int)(j % nslot);
                mbar_wait(&done[slot], (uint32_t)((j / nslot) & 1));
                if (do_store) bstore(base + (size_t)(j % nch) * chunk, sm + (size_t)slot * chunk, chunk);
            }
            bcommit();
            const long long jr = j - (LAG - 1), c = jr + nslot;
            if (jr >= 0 && c < total) {
                bwaitr<LAG - 1>();
                const int slot = (int)(jr % nslot);
                mbar_expect(&full[slot], chunk);
                bload(sm + (size_t)slot * chunk, base + (size_t)(c % nch) * chunk, chunk, &full[slot]);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (tid == 32) {  // consumer: release immediately
        for (long long j = 0; j < total; ++j) {
            const int slot = (int)(j % nslot);
            mbar_wait(&full[slot], (uint32_t)((j / nslot) & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&done[slot]);
        }
    }
}
template <int LAG>
void run(char* buf, size_t slab, int chunk, int nslot, int passes, int do_store) {
    cudaFuncSetAttribute(stream<LAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, nslot * chunk);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    stream<LAG><<<148, 64, nslot * chunk>>>(buf, slab, chunk, nslot, 2, do_store);
    cudaEventRecord(a);
    stream<LAG><<<148, 64, nslot * chunk>>>(buf, slab, chunk, nslot, passes, do_store);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double bytes = (double)148 * slab * passes * (do_store ? 2 : 1);
    printf("chunk %6d B  slots %2d  lag %d  store %d : %7.1f us/pass  %7.1f GB/s  (%s)\n", chunk, nslot, LAG, do_store, ms * 1e3 / passes, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError())); fflush(stdout);
}
int main() {
    const size_t slab = 864 * 1024;   // ~ d=1M, r=16 fp64 per SM
    char* buf; cudaMalloc(&buf, 148 * slab); cudaMemset(buf, 1, 148 * slab);
    const int passes = 50;
    for (int st = 0; st < 2; ++st) {
        run<1>(buf, slab, 16384, 7, passes, st);
        run<3>(buf, slab, 16384, 7, passes, st);
        run<1>(buf, slab, 16384, 10, passes, st);
        run<3>(buf, slab, 16384, 10, passes, st);
        run<3>(buf, slab, 8192, 14, passes, st);
        run<3>(buf, slab, 8192, 20, passes, st);
        run<3>(buf, slab, 32768, 5, passes, st);
        run<3>(buf, slab, 4096, 28, passes, st);
        run<5>(buf, slab, 8192, 20, passes, st);
        run<5>(buf, slab, 4096, 40, passes, st);
    }
    return 0;
}
