// TMA-staged, software-pipelined PSMF / rPSMF filter kernel (sm_100a): the large-d path.
//
// Three warp groups per CTA (one CTA per SM, cooperative launch):
//
//   producer (1 warp, one thread)  moves contiguous CHUNKS of the tiled C (plus the matching slices of y and
//       m) between HBM and a ring of shared-memory slots with bulk asynchronous copies (cp.async.bulk, the
//       1-D TMA path: UBLKCP in SASS) that complete on mbarriers, and bulk-stores updated chunks back.
//       streaming (chunks of the CTA > slots): ring, every chunk loaded and stored once per step;
//       resident  (chunks <= slots): C is loaded once, stays in shared memory for the whole launch.
//   pass warps (12 / 9)  one warp per 32-row tile, all warps independent (psmf_filter.cuh: lane = row for the
//       rank-1 update / y_hat / e, fp64 DMMA fragments for the Gram-type sums).
//   reducer warp (1)  CTA partial -> deterministic grid reduction (+ NVLink exchange) of step t+1, while the
//   solver warps (2 / 5) run the r x r solve of step t.
//
// Software pipeline.  The statistics of step t are sums over C_t = C_{t-1} + e_{t-1} g_{t-1}', and g_{t-1}
// only exists after the solve of step t-1.  Expanding the rank-1 term,
//
//     A_t  = sum m_t c c' = A0 + u g' + g u' + kappa g g'        A0 = sum m_t c c',  u = sum m_t e c
//     h_t  = sum m_t y_t c = h0 + psi g                          h0 = sum m_t y_t c, psi = sum m_t y_t e
//     bu_t = sum m_t e_t c = h_t - A_t xbar_t                     (e_t = y_t - m_t c.xbar_t)
//     q1_t = sum m_t e_t^2 = gamma - 2 xbar_t' h_t + xbar_t' A_t xbar_t,   gamma = sum m_t y_t^2
//
// with c = rows of C_{t-1}, e = e_{t-1}, g = g_{t-1}: every sum on the right is independent of the solve of
// step t-1.  So pass P_t (which turns C_{t-2} in HBM into C_{t-1}, computes e_{t-1} and the sums for step t)
// only needs the solve of step t-2 and runs CONCURRENTLY with the reduction / solve of step t-1; the solve of
// step t assembles A_t, bu_t, q1_t from the reduced sums in O(r^2) and proceeds as in psmf_filter.cuh.  The
// step time becomes max(pass, (pass + reduce + solve) / 2) instead of pass + reduce + solve.  One extra
// read of y_t / m_t per step (+3 % HBM bytes) pays for it; C is still read and written once per step.
//
// Requirements checked by the host (else the direct-load kernel of psmf_filter.cuh is used): 16-byte
// aligned Y / M base pointers and time strides, d a multiple of 16.
#pragma once
#include "psmf_filter.cuh"

namespace psmf {

constexpr int MAXSLOT = 64;
// warp groups: warp 0 = reducer, warps 1..NSOLVE = solver, the next NPASS = V2_CWARPS - 1 - NSOLVE warps run
// the row pass, the last warp is the producer.  Two configurations are instantiated:
//   streaming (pass-bound):            NSOLVE = 2, 12 pass warps
//   resident  (latency-bound, multi-GPU): NSOLVE = 5,  9 pass warps
__host__ __device__ constexpr int s_npass(int NSOLVE) { return V2_CWARPS - 1 - NSOLVE; }

__host__ __device__ constexpr int s_threads() { return (V2_CWARPS + 1) * 32; }

__host__ __device__ constexpr size_t round128(size_t x) { return (x + 127) / 128 * 128; }
template <int R, typename T>
struct SlotLayout {
    static constexpr int TS = V2_TS;
    static constexpr size_t TILE_BYTES = (size_t)R * TILE * sizeof(T);
    static constexpr size_t SLOT = round128(TS * TILE_BYTES);     // a slot holds one chunk of C, nothing else
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

// shared state of the pipeline (in addition to Smem<R>)
template <int R>
struct PipeSmem {
    static constexpr int NSP2 = nstat2_pad(R);
    double tot2[2][NSP2];          // reduced sums of step t in tot2[t & 1]   (reducer -> solver)
    double part2[NSP2];            // reducer scratch
    double asm2[NSP2];             // solver scratch (assembled A_t, h_t)
    double par[2][2 * R];          // par[t & 1] = {g_t (R), xbar_{t+1} (R)} published by the solve of step t
    double xb0[R];                 // xbar_0 (from the state entering the launch)
    uint64_t full[MAXSLOT];        // slot loaded            (producer -> pass warps)
    uint64_t done[MAXSLOT];        // slot processed         (pass warps -> producer)
    uint64_t stats_full[2];        // partial sums of pass t written        (pass warps -> control)
    uint64_t red_free;             // partial sums consumed                  (control -> pass warps)
    uint64_t par_full[2];          // par[t & 1] published                   (solver -> pass warps)
    uint64_t tot_full[2];          // tot2[t & 1] written                    (reducer -> solver)
};

// per-warp accumulators of one pipelined pass
struct PassAcc {
    double g00[2], g01[2], g11[2];   // fragments of A0 = sum m_t c c'
    double u0[2], u1[2];             // fragments of [u | h0] for columns 0..7 / 8..15 of C
    double v[5];                     // per-lane: kappa, psi, gamma, q0, n_obs
    __device__ __forceinline__ void zero() {
        g00[0] = g00[1] = g01[0] = g01[1] = g11[0] = g11[1] = u0[0] = u0[1] = u1[0] = u1[1] = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = 0.0;
    }
};

// ---- pass warp: pass `pass` over the tiles tl = wp, wp + NPW, ... of the CTA ---------------------------
//   pass p in [0, n):  C_{p-2} -> C_{p-1} in the slot, e_{p-1} -> ebuf, Yrec_{p-1}, sums for step p
//   pass n (FLUSH):    C_{n-2} -> C_n (both pending rank-1 updates), Yrec_{n-1}
// y / m of steps p-1 and p are read with coalesced global loads, prefetched one tile ahead (the producer
// thread is latency-bound per bulk operation, so the slots carry C only).
template <typename T>
struct YM {              // RAW loaded values: converting at load time would stall on the load right away
    T yp, yc;            // y_{p-1}[row], y_p[row]
    unsigned char mp, mc;
};
template <typename T>
__device__ __forceinline__ YM<T> load_ym(const KParams& p, const T* __restrict__ Yb, const uint8_t* __restrict__ Mb, int64_t pass,
                                         int64_t row, bool has_prev, bool has_cur) {
    YM<T> r;
    r.yp = (T)0; r.yc = (T)0; r.mp = 1; r.mc = 1;
    if (row < p.d) {
        if (has_prev) {
            r.yp = __ldg(Yb + (pass - 1) * p.ldy + row);
            if (Mb != nullptr) r.mp = __ldg(Mb + (pass - 1) * p.ldm + row);
        }
        if (has_cur) {
            r.yc = __ldg(Yb + pass * p.ldy + row);
            if (Mb != nullptr) r.mc = __ldg(Mb + pass * p.ldm + row);
        }
    }
    return r;
}

template <int R, typename T, bool FLUSH, int NPW>
__device__ __forceinline__ void s_warp_pass(const KParams& p, PipeSmem<R>& ps, double* __restrict__ ebuf,
                                            unsigned char* __restrict__ slots, double* __restrict__ red, T* __restrict__ Yrec_prev,
                                            const T* __restrict__ Yb, const uint8_t* __restrict__ Mb, int tb, int nt, int nslot,
                                            int64_t pass, int wp, int lane) {
    using L = SlotLayout<R, T>;
    constexpr int TS = L::TS, NSP2 = nstat2_pad(R);
    const int nchunks = (nt + TS - 1) / TS;
    const bool streaming = nchunks > nslot;
    // g_{pass-2} and xbar_{pass-1} were published by the solve of step pass-2
    const double* gp = (pass >= 2) ? ps.par[(pass - 2) & 1] : ps.xb0;
    const double* xbp = (pass >= 2) ? ps.par[(pass - 2) & 1] + R : ps.xb0;
    const double* gl = ps.par[(pass - 1) & 1];                          // g_{n-1} (flush only)
    const bool has_prev = pass >= 1, has_cur = !FLUSH;
    PassAcc acc;
    acc.zero();
    long long wait_full = 0;

    YM<T> nx = load_ym<T>(p, Yb, Mb, pass, (int64_t)(tb + wp) * TILE + lane, has_prev, has_cur);
    for (int tl = wp; tl < nt; tl += NPW) {
        const YM<T> ym = nx;
        if (tl + NPW < nt)                                             // prefetch y / m of this warp's next tile
            nx = load_ym<T>(p, Yb, Mb, pass, (int64_t)(tb + tl + NPW) * TILE + lane, has_prev, has_cur);
        const int k = tl / TS, i = tl - k * TS;
        const int64_t kk = pass * nchunks + k;
        const int slot = streaming ? (int)(kk % nslot) : k;
        const uint32_t parity = (uint32_t)((streaming ? kk / nslot : pass) & 1);
        unsigned char* sb = slots + (size_t)slot * L::SLOT;
        const long long c0 = clock64();
        mbar_wait(&ps.full[slot], parity);
        wait_full += clock64() - c0;
        const int64_t row = (int64_t)(tb + tl) * TILE + lane;
        const int rl = tl * TILE + lane;
        const bool inb = row < p.d;
        T* tile = reinterpret_cast<T*>(sb) + (size_t)i * (R * TILE);

        if ((p.flags & F_DBG_NOCOMPUTE) != 0) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                const int ntc = min(TS, nt - k * TS);
                mbar_arrive_n(&ps.done[slot], (i == ntc - 1) ? (uint32_t)(TS - ntc + 1) : 1u);
            }
            continue;
        }
        // ---- phase 1, lane = row ----
        double c[R];
#pragma unroll
        for (int j = 0; j < R; ++j) c[j] = (double)tile[tile_pos(j, lane)];
        double e = 0.0;
        if (pass >= 2) {                                           // C_{p-1} = C_{p-2} + e_{p-2} g_{p-2}'   (rPSMF.py:111)
            const double epp = ebuf[rl];
#pragma unroll
            for (int j = 0; j < R; ++j) c[j] = fma(epp, gp[j], c[j]);
        }
        if (pass >= 1) {                                           // e_{p-1} = y_{p-1} - m_{p-1} C_{p-1} xbar_{p-1}
            double yh4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < R; ++j) yh4[j & 3] = fma(c[j], xbp[j], yh4[j & 3]);     // rPSMF.py:89
            const double yh = (yh4[0] + yh4[1]) + (yh4[2] + yh4[3]);
            e = (double)ym.yp - ((inb && ym.mp != 0) ? yh : 0.0);  // rPSMF.py:101
            if (Yrec_prev != nullptr && inb) Yrec_prev[row] = (T)yh;
        }
        if constexpr (FLUSH) {                                     // C_n = C_{n-1} + e_{n-1} g_{n-1}'
#pragma unroll
            for (int j = 0; j < R; ++j) c[j] = fma(e, gl[j], c[j]);
#pragma unroll
            for (int j = 0; j < R; ++j) tile[tile_pos(j, lane)] = (T)c[j];
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) tile[tile_pos(j, lane)] = (T)c[j];
            ebuf[rl] = e;
            const bool mi = inb && ym.mc != 0;
            const double yi = (double)ym.yc;
            acc.v[0] += mi ? e * e : 0.0;                          // kappa
            acc.v[1] += mi ? yi * e : 0.0;                         // psi
            acc.v[2] += mi ? yi * yi : 0.0;                        // gamma
            acc.v[3] += mi ? 0.0 : yi * yi;                        // q0 (rows missing at step p: e_p = y_p)
            acc.v[4] += mi ? 1.0 : 0.0;                            // n_obs
            const double me = mi ? e : 0.0, my = mi ? yi : 0.0;   // B columns 0 / 1 of the [u | h0] product
            __syncwarp();
            // ---- phase 2: A0 += sum m c c', [u | h0] += sum m c [e, y]  (fp64 DMMA, k = row) ----
            const unsigned mbits = __ballot_sync(FULL, mi);
            const int kq = lane & 3, mm = lane >> 2;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int r4 = 4 * s + kq;
                const bool mrow = (mbits >> r4) & 1u;
                const int pos = r4 ^ (mm << 2);
                const double a0 = (mm < R) ? (double)tile[mm * 32 + pos] : 0.0;
                const double b0 = mrow ? a0 : 0.0;
                const double se = __shfl_sync(FULL, me, r4), sy = __shfl_sync(FULL, my, r4);
                const double bx = (mm == 0) ? se : ((mm == 1) ? sy : 0.0);
                dmma884(acc.g00, a0, b0);
                dmma884(acc.u0, a0, bx);
                if constexpr (R > 8) {
                    const double a1 = (8 + mm < R) ? (double)tile[(8 + mm) * 32 + pos] : 0.0;
                    const double b1 = mrow ? a1 : 0.0;
                    dmma884(acc.g01, a0, b1);
                    dmma884(acc.g11, a1, b1);
                    dmma884(acc.u1, a1, bx);
                }
            }
        }
        // this warp is done with its tile: make the generic-proxy writes visible to the bulk store and
        // release the slot (the owner of the last tile of a partial chunk also signs for the missing tiles)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const int ntc = min(TS, nt - k * TS);
            mbar_arrive_n(&ps.done[slot], (i == ntc - 1) ? (uint32_t)(TS - ntc + 1) : 1u);
        }
    }
    if constexpr (!FLUSH) {
        if (wp == 0 && lane == 0) stamp_pass(p, pass, 10, (unsigned long long)wait_full, false);
        // the control warps have consumed the partial sums of the previous pass
        if (pass >= 1) mbar_wait(&ps.red_free, (uint32_t)((pass - 1) & 1));
        double* r0 = red + wp * NSP2;
        const int kq = lane & 3, mm = lane >> 2;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const int n = 2 * kq + x;
            if (mm <= n && n < R) r0[gram_off(R, mm) + (n - mm)] = acc.g00[x];
            if constexpr (R > 8) {
                if (8 + n < R) r0[gram_off(R, mm) + (8 + n - mm)] = acc.g01[x];
                if (mm <= n && 8 + n < R) r0[gram_off(R, 8 + mm) + (n - mm)] = acc.g11[x];
            }
        }
        if (kq == 0) {                                             // D[mm][0] = u_mm, D[mm][1] = h0_mm
            if (mm < R) {
                r0[ngram(R) + mm] = acc.u0[0];
                r0[ngram(R) + R + mm] = acc.u0[1];
            }
            if constexpr (R > 8) {
                if (8 + mm < R) {
                    r0[ngram(R) + 8 + mm] = acc.u1[0];
                    r0[ngram(R) + R + 8 + mm] = acc.u1[1];
                }
            }
        }
        int base = 0, lim = 5;
        bfly<5, 16, 5>(acc.v, lane, base, lim);
        if (base < lim) r0[ngram(R) + 2 * R + base] = acc.v[0];
        __syncwarp();
        if (lane == 0) mbar_arrive(&ps.stats_full[pass & 1]);
    }
}

// ---- producer: one thread drives all bulk copies of the CTA (C chunks only) ------------------------------
// The loop is latency-bound per iteration (~0.6 us measured, scratch/bulkbench.cu): one bulk store and one
// bulk load of a 16 KB chunk per iteration is what sustains the HBM rate, so nothing else goes through it.
template <int R, typename T>
__device__ void s_producer(const KParams& p, PipeSmem<R>& ps, unsigned char* slots, T* Cs, int tb, int nt, int nslot) {
    using L = SlotLayout<R, T>;
    constexpr int TS = L::TS;
    const int nchunks = (nt + TS - 1) / TS;
    const bool streaming = nchunks > nslot;
    const int64_t npass = p.n_steps + 1;
    const uint32_t last_bytes = (uint32_t)((nt - (nchunks - 1) * TS) * L::TILE_BYTES);
    const uint32_t full_bytes = (uint32_t)(TS * L::TILE_BYTES);
    T* Cb = Cs + (size_t)tb * (R * TILE);
    constexpr size_t CHUNK_ELEMS = (size_t)TS * R * TILE;
    if (streaming) {
        // Ring of nslot slots over the chunk sequence (consumption order: pass-major).  Chunk j is stored as soon
        // as the pass warps release it and its slot is reloaded with chunk j + nslot right after the store has
        // left shared memory.
        const int64_t total = npass * nchunks;
        const int lag = nchunks - nslot + 1;          // >= 2: store groups issued after the previous version of a chunk
        int lk = 0;                                   // chunk-in-pass index of the next load
        int64_t lc = 0;                               // global index of the next load
        for (; lc < total && lc < nslot; ++lc) {
            const uint32_t bytes = lk == nchunks - 1 ? last_bytes : full_bytes;
            mbar_arrive_expect_tx(&ps.full[lc], bytes);
            bulk_load(slots + (size_t)lc * L::SLOT, Cb + (size_t)lk * CHUNK_ELEMS, bytes, &ps.full[lc]);
            if (++lk == nchunks) lk = 0;
        }
        int sk = 0, slot = 0;                         // chunk-in-pass index / slot of the next store
        uint32_t par = 0;
        for (int64_t j = 0; j < total; ++j) {
            mbar_wait(&ps.done[slot], par);                                       // chunk j processed
            unsigned char* sb = slots + (size_t)slot * L::SLOT;
            if ((p.flags & F_DBG_NOSTORE) == 0)
                bulk_store(Cb + (size_t)sk * CHUNK_ELEMS, sb, sk == nchunks - 1 ? last_bytes : full_bytes);
            bulk_commit();
            if (lc < total) {
                bulk_wait_read<0>();                                               // the store has left shared memory
                // the previous version of the chunk loaded now (one pass ago) was stored `lag` groups ago
                if (lag >= 8) bulk_wait<8>(); else if (lag >= 4) bulk_wait<4>(); else if (lag >= 2) bulk_wait<2>(); else bulk_wait<0>();
                const uint32_t bytes = lk == nchunks - 1 ? last_bytes : full_bytes;
                mbar_arrive_expect_tx(&ps.full[slot], bytes);
                bulk_load(sb, Cb + (size_t)lk * CHUNK_ELEMS, bytes, &ps.full[slot]);
                if (++lk == nchunks) lk = 0;
                ++lc;
            }
            if (++sk == nchunks) sk = 0;
            if (++slot == nslot) { slot = 0; par ^= 1u; }
        }
    } else {
        for (int k = 0; k < nchunks; ++k) {
            const uint32_t bytes = k == nchunks - 1 ? last_bytes : full_bytes;
            mbar_arrive_expect_tx(&ps.full[k], bytes);
            bulk_load(slots + (size_t)k * L::SLOT, Cb + (size_t)k * CHUNK_ELEMS, bytes, &ps.full[k]);
        }
        for (int64_t pass = 1; pass < npass; ++pass)
            for (int k = 0; k < nchunks; ++k) {
                mbar_wait(&ps.done[k], (uint32_t)((pass - 1) & 1));               // pass warps are done with pass-1
                mbar_arrive(&ps.full[k]);                                          // C stays resident
            }
        for (int k = 0; k < nchunks; ++k) {
            mbar_wait(&ps.done[k], (uint32_t)((npass - 1) & 1));
            bulk_store(Cb + (size_t)k * CHUNK_ELEMS, slots + (size_t)k * L::SLOT, k == nchunks - 1 ? last_bytes : full_bytes);
            bulk_commit();
        }
    }
    bulk_wait<0>();
}

// ---- solver warps: pipelined sums of step t (+ g_{t-1}, xbar_t) -> statistics vector of psmf_filter.cuh ----
// sh.g = g_{t-1}, sh.xb = xbar_t, sh.w1/w0 for step t are current (left by the solve of step t-1).
template <int R, int BAR>
__device__ __forceinline__ void assemble_stats(Smem<R>& sh, PipeSmem<R>& ps, const double* __restrict__ t2, int tid, int nthr) {
    constexpr int NGm = ngram(R);
    const double kappa = t2[NGm + 2 * R + 0], psi = t2[NGm + 2 * R + 1], gamma = t2[NGm + 2 * R + 2];
    // A_t (packed upper triangle) -> part2[0..NGm), h_t -> part2[NGm..NGm+R)
    for (int idx = tid; idx < R * R; idx += nthr) {
        const int j = idx / R, k = idx % R;
        if (j <= k) {
            const double gj = sh.g[j], gk = sh.g[k];
            ps.asm2[gram_off(R, j) + (k - j)] = t2[gram_off(R, j) + (k - j)] + (t2[NGm + j] * gk + gj * t2[NGm + k]) + kappa * gj * gk;
        }
    }
    if (tid >= nthr - R) {
        const int j = tid - (nthr - R);
        ps.asm2[NGm + j] = fma(psi, sh.g[j], t2[NGm + R + j]);
    }
    sync_n<BAR>(nthr);
    // bu = h - A xbar ; q1 = gamma - 2 xbar'h + xbar'A xbar   (warp 0)
    if (tid < 32) {
        const int lane = tid;
        double ax = 0.0, hj = 0.0, xj = 0.0;
        if (lane < R) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int k = 0; k < R; k += 2) {
                const int lo = k < lane ? k : lane, hi = k < lane ? lane : k;
                a0 = fma(ps.asm2[gram_off(R, lo) + hi - lo], sh.xb[k], a0);
                if (k + 1 < R) {
                    const int lo1 = k + 1 < lane ? k + 1 : lane, hi1 = k + 1 < lane ? lane : k + 1;
                    a1 = fma(ps.asm2[gram_off(R, lo1) + hi1 - lo1], sh.xb[k + 1], a1);
                }
            }
            ax = a0 + a1;
            hj = ps.asm2[NGm + lane];
            xj = sh.xb[lane];
        }
        const double xh = warp_allsum(xj * hj);
        const double xax = warp_allsum(xj * ax);
        const double w1 = sh.w1, w0 = sh.w0;
        const double q1 = gamma - 2.0 * xh + xax;
        const double q0 = t2[NGm + 2 * R + 3];
        if (lane < R) sh.tot[NGm + lane] = w1 * (hj - ax);                     // b = w1 sum m e c
        if (lane == 0) {
            sh.tot[NGm + R + 0] = w1 * q1 + w0 * q0;                           // s = diff' Ri diff
            sh.tot[NGm + R + 1] = q1;
            sh.tot[NGm + R + 2] = q0;
            sh.tot[NGm + R + 3] = t2[NGm + 2 * R + 4];
        }
    }
    {
        const double w1 = sh.w1;
        for (int idx = tid; idx < NGm; idx += nthr) sh.tot[idx] = w1 * ps.asm2[idx];   // G = w1 A_t
    }
    sync_n<BAR>(nthr);
}

template <int R, typename T, int NSOLVE>
__global__ void __launch_bounds__(s_threads(), 1) psmf_stream_kernel(const KParams p) {
    using L = SlotLayout<R, T>;
    constexpr int NSP2 = nstat2_pad(R), NST2 = nstat2(R);
    constexpr int NPW = s_npass(NSOLVE);
    constexpr int NSV = NSOLVE * 32;                            // solver threads
    constexpr int BAR_RED = 0, BAR_GJ = 1, BAR_XB0 = 2, BAR_SOLVE = 3;
    extern __shared__ __align__(128) unsigned char dyn_smem_s[];
    __shared__ Smem<R> sh;
    __shared__ __align__(8) PipeSmem<R> ps;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int series = blockIdx.x / p.cps;
    const int part = blockIdx.x % p.cps;
    const int ntiles = (int)((p.d + TILE - 1) / TILE);
    const int tb = (int)((int64_t)ntiles * part / p.cps);
    const int te = (int)((int64_t)ntiles * (part + 1) / p.cps);
    const int nt = te - tb;
    const bool writer = part == 0;
    const int nslot = p.nslot;
    const int64_t n = p.n_steps;

    unsigned char* slots = dyn_smem_s;
    double* ebuf = reinterpret_cast<double*>(dyn_smem_s + (size_t)nslot * L::SLOT);
    T* Cs = reinterpret_cast<T*>(p.C) + (int64_t)series * p.c_series_stride;
    double* stg = p.state + (int64_t)series * st_size(R);

    if (tid == 0) {
        for (int s = 0; s < nslot; ++s) {
            mbar_init(&ps.full[s], 1);
            mbar_init(&ps.done[s], L::TS);
        }
        mbar_init(&ps.stats_full[0], NPW);
        mbar_init(&ps.stats_full[1], NPW);
        mbar_init(&ps.red_free, 1);
        mbar_init(&ps.par_full[0], 1);
        mbar_init(&ps.par_full[1], 1);
        mbar_init(&ps.tot_full[0], 1);
        mbar_init(&ps.tot_full[1], 1);
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

    if (warp == V2_CWARPS) {                 // ---- producer warp ----
        if (lane == 0) s_producer<R, T>(p, ps, slots, Cs, tb, nt, nslot);
        return;
    }

    if (warp == 0) {
        // ---- reducer warp: partial sums of pass t -> grid (and GPU) totals of step t ----
        for (int64_t t = 0; t < n; ++t) {
            stamp(p, t, 0);
            mbar_wait(&ps.stats_full[t & 1], (uint32_t)((t >> 1) & 1));
            stamp(p, t, 1);
            for (int e = lane; e < NST2; e += 32) {                     // CTA partial: fixed order over the pass warps
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NPW; ++w) s += sh.red[w * NSP2 + e];
                ps.part2[e] = s;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ps.red_free);
            stamp(p, t, 2);
            // tot2[t & 1] was last read by the solve of step t-2, which completed before pass t could start
            grid_reduce<NST2, NSP2, BAR_RED>(p, ps.part2, ps.tot2[t & 1], lane, lane, 0, t, series, part, 32);
            __threadfence_block();
            if (lane == 0) mbar_arrive(&ps.tot_full[t & 1]);
            stamp(p, t, 5);
        }
        return;
    }

    if (warp <= NSOLVE) {
        // ---- solver warps: r x r solve of step t (needs the totals of step t and the solve of step t-1) ----
        const int st = tid - 32, swarp = warp - 1;
        predict_cta<R, BAR_SOLVE>(p, sh, st, p.k0, series, NSV);       // xbar_0, Pbar_0, a_0, w1/w0
        if (st < R) ps.xb0[st] = sh.xb[st];
        __threadfence_block();
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_XB0), "r"(NSV + NPW * 32) : "memory");     // xbar_0 is published
        for (int64_t t = 0; t < n; ++t) {
            mbar_wait(&ps.tot_full[t & 1], (uint32_t)((t >> 1) & 1));
            if (st == 0) stamp_pass(p, t, 12, 0, true);
            assemble_stats<R, BAR_SOLVE>(sh, ps, ps.tot2[t & 1], st, NSV);
            small_update<R, NSV, BAR_SOLVE, BAR_GJ>(p, sh, st, lane, swarp, series, t, writer, NSV);
            // publish g_t and xbar_{t+1} for pass t+2 (small_update ended with a barrier over the solver threads)
            if (st < R) {
                ps.par[t & 1][st] = sh.g[st];
                ps.par[t & 1][R + st] = sh.xb[st];
            }
            sync_n<BAR_SOLVE>(NSV);
            if (st == 0) {
                mbar_arrive(&ps.par_full[t & 1]);
                stamp_pass(p, t, 13, 0, true);
            }
        }
        if (writer) {
            for (int i = st; i < R * R; i += NSV) {
                stg[st_P(R) + i] = sh.P[i];
                stg[st_V(R) + i] = sh.V[i];
                stg[st_Q(R) + i] = sh.Q[i];
            }
            if (st < R) {
                stg[st_x(R) + st] = sh.x[st];
                if (p.grad_out != nullptr) p.grad_out[(int64_t)series * R + st] = sh.grad[st];
            }
            if (st == 0) {
                stg[st_rho(R)] = sh.rho;
                stg[st_lam(R)] = sh.lam;
            }
        }
        return;
    }

    // ---- pass warps ----
    const int wp = warp - 1 - NSOLVE;
    const T* Yb = reinterpret_cast<const T*>(p.Y) + (int64_t)series * p.ysst;
    const uint8_t* Mb = p.M != nullptr ? p.M + (int64_t)series * p.msst : nullptr;
    asm volatile("bar.sync %0, %1;" ::"n"(BAR_XB0), "r"(NSV + NPW * 32) : "memory");   // xbar_0 available
    for (int64_t pass = 0; pass < n; ++pass) {
        // pass `pass` needs the solve of step pass-2 (g_{pass-2}, xbar_{pass-1})
        if (wp == 0 && lane == 0) stamp_pass(p, pass, 11, 0, true);
        if (pass >= 2) mbar_wait(&ps.par_full[(pass - 2) & 1], (uint32_t)(((pass - 2) >> 1) & 1));
        if (wp == 0 && lane == 0) stamp_pass(p, pass, 8, 0, true);
        T* Yrec_prev = (p.Yrec && pass >= 1) ? reinterpret_cast<T*>(p.Yrec) + (int64_t)series * p.recsst + (pass - 1) * p.ldrec : nullptr;
        s_warp_pass<R, T, false, NPW>(p, ps, ebuf, slots, sh.red, Yrec_prev, Yb, Mb, tb, nt, nslot, pass, wp, lane);
        if (wp == 0 && lane == 0) stamp_pass(p, pass, 9, 0, true);
    }
    // flush: both pending rank-1 updates -> C_n; needs the solves of steps n-2 and n-1
    if (n >= 2) mbar_wait(&ps.par_full[(n - 2) & 1], (uint32_t)(((n - 2) >> 1) & 1));
    mbar_wait(&ps.par_full[(n - 1) & 1], (uint32_t)(((n - 1) >> 1) & 1));
    {
        T* Yrec_prev = p.Yrec ? reinterpret_cast<T*>(p.Yrec) + (int64_t)series * p.recsst + (n - 1) * p.ldrec : nullptr;
        s_warp_pass<R, T, true, NPW>(p, ps, ebuf, slots, sh.red, Yrec_prev, Yb, Mb, tb, nt, nslot, n, wp, lane);
    }
}

}  // namespace psmf
