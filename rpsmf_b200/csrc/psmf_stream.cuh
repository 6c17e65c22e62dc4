// TMA-staged PSMF / rPSMF filter kernel (sm_100a): the large-d path.
//
// Same step as psmf_filter.cuh, but C never travels through registers from global memory.  A producer
// warp moves contiguous CHUNKS of the tiled C (plus the matching slices of y_t and m_t) between HBM and a
// ring of shared-memory slots with bulk asynchronous copies (cp.async.bulk, the 1-D TMA path: UBLKCP in
// SASS) that complete on mbarriers; the consumer warps read tiles from shared memory, apply the pending
// rank-1 update, accumulate the statistics, write the updated tile back into the slot, and the producer
// bulk-stores it to HBM.  Two regimes, chosen per CTA:
//
//   streaming  (chunks of the CTA > slots): the slots form a ring; per step every chunk is loaded once and
//              stored once -> HBM traffic = 2 d r s_C + d (s_y + 1) bytes per step.  While the consumers
//              sit in the reduction / grid barrier / r x r solve, the producer already fills the ring
//              with the first chunks of the next step.
//   resident   (chunks <= slots): C is loaded once, stays in shared memory for the whole launch and is
//              stored once at the end; only y_t / m_t stream.
//
// Requirements checked by the host (else the direct-load kernel of psmf_filter.cuh is used): 16-byte
// aligned Y / M base pointers and time strides, d a multiple of 16.
#pragma once
#include "psmf_filter.cuh"

namespace psmf {

constexpr int MAXSLOT = 64;

__host__ __device__ constexpr int s_threads() { return (V2_CWARPS + 1) * 32; }

__host__ __device__ constexpr size_t round128(size_t x) { return (x + 127) / 128 * 128; }
template <int R, typename T>
struct SlotLayout {
    static constexpr int TS = V2_TS;
    static constexpr size_t TILE_BYTES = (size_t)R * TILE * sizeof(T);
    static constexpr size_t CB = round128(TS * TILE_BYTES);
    static constexpr size_t YB = round128((size_t)TS * TILE * sizeof(T));
    static constexpr size_t MB = round128((size_t)TS * TILE);
    static constexpr size_t SLOT = CB + YB + MB;
};

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- consumer: one warp over its tiles of one pass (tile tl of the CTA belongs to warp tl % V2_CWARPS) ----
// FLUSH: apply the pending rank-1 update only (the pass after the last step).
template <int R, typename T, bool FLUSH>
__device__ __forceinline__ void s_warp_pass(const KParams& p, Smem<R>& sh, double* __restrict__ ebuf,
                                            unsigned char* __restrict__ slots, uint64_t* full, uint64_t* done,
                                            T* __restrict__ Yrec_t, bool masked, int tb, int nt, int nslot, int64_t pass,
                                            int warp, int lane) {
    using L = SlotLayout<R, T>;
    constexpr int NSP = nstat_pad(R), TS = L::TS;
    const double w1 = sh.w1, w0 = sh.w0;
    const int nchunks = (nt + TS - 1) / TS;
    const bool streaming = nchunks > nslot;
    TileAcc<R> acc;
    acc.zero();

    for (int tl = warp; tl < nt; tl += V2_CWARPS) {
        const int k = tl / TS, i = tl - k * TS;
        const int64_t kk = pass * nchunks + k;
        const int slot = streaming ? (int)(kk % nslot) : k;
        const uint32_t parity = (uint32_t)((streaming ? kk / nslot : pass) & 1);
        unsigned char* sb = slots + (size_t)slot * L::SLOT;
        mbar_wait(&full[slot], parity);
        const int64_t row = (int64_t)(tb + tl) * TILE + lane;
        const int rl = tl * TILE + lane;
        T* tile = reinterpret_cast<T*>(sb) + (size_t)i * (R * TILE);
        const double ep = ebuf[rl];
        if constexpr (FLUSH) {
#pragma unroll
            for (int j = 0; j < R; ++j) tile[tile_pos(j, lane)] = (T)fma(ep, sh.g[j], (double)tile[tile_pos(j, lane)]);
        } else {
            double c[R];
#pragma unroll
            for (int j = 0; j < R; ++j) c[j] = (double)tile[tile_pos(j, lane)];
            const bool inb = row < p.d;
            bool mi = inb;
            if (masked && inb) mi = (sb + L::CB + L::YB)[i * TILE + lane] != 0;
            const double yi = inb ? (double)reinterpret_cast<const T*>(sb + L::CB)[i * TILE + lane] : 0.0;
            double e, yh;
            row_stats<R>(acc, sh, c, ep, inb, mi, yi, w1, w0, e, yh);
#pragma unroll
            for (int j = 0; j < R; ++j) tile[tile_pos(j, lane)] = (T)c[j];
            ebuf[rl] = e;
            if (Yrec_t != nullptr && inb) Yrec_t[row] = (T)yh;
            const unsigned mbits = __ballot_sync(FULL, mi);
            __syncwarp();
            tile_gram<R, T>(acc, tile, mbits, lane);
        }
        // this warp is done with its tile: make the generic-proxy writes visible to the bulk store and
        // release the slot (the warp that owns the last tile of a partial chunk also signs for the
        // tiles that do not exist)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const int ntc = min(TS, nt - k * TS);
            mbar_arrive_n(&done[slot], (i == ntc - 1) ? (uint32_t)(TS - ntc + 1) : 1u);
        }
    }
    if constexpr (!FLUSH) acc_writeout<R>(acc, sh.red + warp * NSP, w1, lane);
}

// ---- producer: one thread drives all bulk copies of the CTA -------------------------------------------
template <int R, typename T>
__device__ void s_producer(const KParams& p, unsigned char* slots, uint64_t* full, uint64_t* done, T* Cs, int series,
                           int tb, int nt, int nslot) {
    using L = SlotLayout<R, T>;
    constexpr int TS = L::TS;
    const int nchunks = (nt + TS - 1) / TS;
    const bool streaming = nchunks > nslot;
    const int64_t npass = p.n_steps + 1;
    const bool masked = p.M != nullptr;
    int64_t kk = 0;
    for (int64_t pass = 0; pass < npass; ++pass) {
        const bool last = pass == p.n_steps;
        const T* Yt = last ? nullptr : reinterpret_cast<const T*>(p.Y) + (int64_t)series * p.ysst + pass * p.ldy;
        const uint8_t* Mt = (last || !masked) ? nullptr : p.M + (int64_t)series * p.msst + pass * p.ldm;
        for (int k = 0; k < nchunks; ++k, ++kk) {
            const int slot = streaming ? (int)(kk % nslot) : k;
            const int64_t use = streaming ? kk / nslot : pass;
            unsigned char* sb = slots + (size_t)slot * L::SLOT;
            if (use > 0) {
                mbar_wait(&done[slot], (uint32_t)((use - 1) & 1));        // consumers released the previous occupant
                if (streaming) {
                    const int pk = (int)((kk - nslot) % nchunks);
                    const int ptiles = min(TS, nt - pk * TS);
                    bulk_store(Cs + (size_t)(tb + pk * TS) * (R * TILE), sb, (uint32_t)(ptiles * L::TILE_BYTES));
                    bulk_commit();
                    bulk_wait_read<0>();                                   // slot may be overwritten
                    // the chunk loaded below was stored (nchunks - nslot) groups ago: make sure that store has
                    // fully completed before reading it back
                    if (nchunks - nslot >= 2) bulk_wait<2>(); else bulk_wait<0>();
                }
            }
            const int ntc = min(TS, nt - k * TS);
            const int64_t row0 = (int64_t)(tb + k * TS) * TILE;
            int64_t vrows = p.d - row0;
            vrows = vrows < 0 ? 0 : (vrows > (int64_t)ntc * TILE ? (int64_t)ntc * TILE : vrows);
            const bool loadC = streaming || pass == 0;
            const uint32_t cbytes = loadC ? (uint32_t)(ntc * L::TILE_BYTES) : 0u;
            const uint32_t ybytes = (Yt != nullptr) ? (uint32_t)(vrows * sizeof(T)) : 0u;
            const uint32_t mbytes = (Mt != nullptr) ? (uint32_t)vrows : 0u;
            const uint32_t tx = cbytes + ybytes + mbytes;
            if (tx == 0) {
                mbar_arrive(&full[slot]);
            } else {
                mbar_arrive_expect_tx(&full[slot], tx);
                if (cbytes) bulk_load(sb, Cs + (size_t)(tb + k * TS) * (R * TILE), cbytes, &full[slot]);
                if (ybytes) bulk_load(sb + L::CB, Yt + row0, ybytes, &full[slot]);
                if (mbytes) bulk_load(sb + L::CB + L::YB, Mt + row0, mbytes, &full[slot]);
            }
        }
    }
    // drain: store what is still only in shared memory
    if (streaming) {
        for (int64_t j = kk - nslot; j < kk; ++j) {
            const int slot = (int)(j % nslot);
            mbar_wait(&done[slot], (uint32_t)((j / nslot) & 1));
            const int pk = (int)(j % nchunks);
            const int ptiles = min(TS, nt - pk * TS);
            bulk_store(Cs + (size_t)(tb + pk * TS) * (R * TILE), slots + (size_t)slot * L::SLOT, (uint32_t)(ptiles * L::TILE_BYTES));
            bulk_commit();
        }
    } else {
        for (int k = 0; k < nchunks; ++k) {
            mbar_wait(&done[k], (uint32_t)((npass - 1) & 1));
            const int ptiles = min(TS, nt - k * TS);
            bulk_store(Cs + (size_t)(tb + k * TS) * (R * TILE), slots + (size_t)k * L::SLOT, (uint32_t)(ptiles * L::TILE_BYTES));
            bulk_commit();
        }
    }
    bulk_wait<0>();
}

template <int R, typename T>
__global__ void __launch_bounds__(s_threads(), 1) psmf_stream_kernel(const KParams p) {
    using L = SlotLayout<R, T>;
    constexpr int NSP = nstat_pad(R), NST = nstat(R);
    constexpr int NCW = V2_CWARPS, NCT = NCW * 32;
    extern __shared__ __align__(128) unsigned char dyn_smem_s[];
    __shared__ Smem<R> sh;
    __shared__ __align__(8) uint64_t full[MAXSLOT];
    __shared__ __align__(8) uint64_t done[MAXSLOT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int series = blockIdx.x / p.cps;
    const int part = blockIdx.x % p.cps;
    const int ntiles = (int)((p.d + TILE - 1) / TILE);
    const int tb = (int)((int64_t)ntiles * part / p.cps);
    const int te = (int)((int64_t)ntiles * (part + 1) / p.cps);
    const int nt = te - tb;
    const bool writer = part == 0;
    const int nslot = p.nslot;

    unsigned char* slots = dyn_smem_s;
    double* ebuf = reinterpret_cast<double*>(dyn_smem_s + (size_t)nslot * L::SLOT);
    T* Cs = reinterpret_cast<T*>(p.C) + (int64_t)series * p.c_series_stride;
    double* stg = p.state + (int64_t)series * st_size(R);

    if (tid == 0) {
        for (int s = 0; s < nslot; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&done[s], L::TS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < R * R; i += blockDim.x) {
        sh.P[i] = stg[st_P(R) + i];
        sh.V[i] = stg[st_V(R) + i];
        sh.Q[i] = stg[st_Q(R) + i];
    }
    if (tid < R) {
        sh.x[tid] = stg[st_x(R) + tid];
        sh.th[tid] = stg[st_theta(R) + tid];
        sh.g[tid] = 0.0;
        sh.grad[tid] = 0.0;
    }
    if (tid == 0) {
        sh.rho = stg[st_rho(R)];
        sh.lam = stg[st_lam(R)];
    }
    for (int i = tid; i < nt * TILE; i += blockDim.x) ebuf[i] = 0.0;
    __syncthreads();

    if (warp == NCW) {                       // ---- producer warp ----
        if (lane == 0) s_producer<R, T>(p, slots, full, done, Cs, series, tb, nt, nslot);
        return;
    }

    // ---- consumer warps ----
    const bool masked = p.M != nullptr;
    predict_cta<R>(p, sh, tid, p.k0, series, NCT);

    for (int64_t t = 0; t < p.n_steps; ++t) {
        T* Yrec_t = p.Yrec ? reinterpret_cast<T*>(p.Yrec) + (int64_t)series * p.recsst + t * p.ldrec : nullptr;
        stamp(p, t, 0);
        s_warp_pass<R, T, false>(p, sh, ebuf, slots, full, done, Yrec_t, masked, tb, nt, nslot, t, warp, lane);
        stamp(p, t, 1);
        sync_n(NCT);
        stamp(p, t, 2);
        if (tid < NST) {                     // CTA partial: fixed order over the warps
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NCW; ++w) s += sh.red[w * NSP + tid];
            sh.part[tid] = s;
        }
        grid_reduce<R>(p, sh, tid, lane, warp, t, series, part, NCT);
        stamp(p, t, 5);
        small_update<R>(p, sh, tid, lane, warp, series, t, writer, NCT);
        stamp(p, t, 6);
    }

    // flush pass: pending rank-1 update of the last step
    s_warp_pass<R, T, true>(p, sh, ebuf, slots, full, done, (T*)nullptr, masked, tb, nt, nslot, p.n_steps, warp, lane);
    if (writer) {
        for (int i = tid; i < R * R; i += NCT) {
            stg[st_P(R) + i] = sh.P[i];
            stg[st_V(R) + i] = sh.V[i];
            stg[st_Q(R) + i] = sh.Q[i];
        }
        if (tid < R) {
            stg[st_x(R) + tid] = sh.x[tid];
            if (p.grad_out != nullptr) p.grad_out[(int64_t)series * R + tid] = sh.grad[tid];
        }
        if (tid == 0) {
            stg[st_rho(R)] = sh.rho;
            stg[st_lam(R)] = sh.lam;
        }
    }
}

}  // namespace psmf
