// unit test of gather_ranks (psmf_filter.cuh), built and run by tests/test_multi_gpu.py::test_gather_ranks_unit:
//   nvcc -gencode arch=compute_100a,code=sm_100a -I rpsmf_b200/csrc tests/cuda/gather_test.cu -o /tmp/gather_test
#include <cstdio>
#include "psmf_filter.cuh"
using namespace psmf;
__global__ void k(KParams p, uint4* mbox, double* out, int world, int nst, int with_local) {
    const uint32_t tag = tag_of(1);
    if (blockIdx.x == 1) {                       // "peers": the other ranks write their cells with staggered delays
        for (int r = 0; r < world; ++r) {
            if (r == p.rank && !with_local) continue;
            long long t0 = clock64();
            while (clock64() - t0 < 20000LL * r) {}
            for (int e = threadIdx.x; e < nst; e += blockDim.x)
                cell_store(r == p.rank ? mbox + (size_t)MAX_PEERS * MBOX_SLOT + e : mbox + (size_t)r * MBOX_SLOT + e, 1000.0 * (r + 1) + e, tag + (r == p.rank ? 5 : 0));
        }
        return;
    }
    for (int e = threadIdx.x; e < nst; e += blockDim.x)
        out[e] = gather_ranks(p, mbox, e, tag, with_local ? mbox + (size_t)MAX_PEERS * MBOX_SLOT + e : nullptr, tag + 5, 1000.0 * (p.rank + 1) + e, 0);
}
int main() {
    for (int cfg = 0; cfg < 28; ++cfg) {
        const int world = 2 + cfg % 7, with_local = (cfg / 7) & 1, rank = cfg < 14 ? 0 : world - 1 - (cfg % 2);
        KParams p; memset(&p, 0, sizeof(p));
        p.world = world; p.rank = rank; p.spin_ns = 5000000000ULL;
        unsigned long long* bar; long long* status; uint4* mbox; double* out;
        cudaMalloc(&bar, 64); cudaMemset(bar, 0, 64); cudaMalloc(&status, 8); cudaMemset(status, 0xFF, 8);
        cudaMalloc(&mbox, (MAX_PEERS + 1) * MBOX_SLOT * 16); cudaMemset(mbox, 0, (MAX_PEERS + 1) * MBOX_SLOT * 16);
        cudaMalloc(&out, 256 * 8);
        p.bar = bar; p.status = status;
        k<<<2, 256>>>(p, mbox, out, world, 156, with_local);
        cudaError_t e = cudaDeviceSynchronize();
        double h[156]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 156; ++i) {
            double want = 0.0; for (int r = 0; r < world; ++r) want += 1000.0 * (r + 1) + i;
            if (h[i] != want) { if (bad < 3) printf("  world %d entry %d got %.1f want %.1f\n", world, i, h[i], want); ++bad; }
        }
        printf("world %d rank %d local_cell %d: %s (%d bad) %s\n", world, rank, with_local, bad ? "FAIL" : "ok", bad, cudaGetErrorString(e));
    }
}
