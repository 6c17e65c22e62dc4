// Kernels either side of the filter step: data ingest, the missing-segment generator, the
// post-sweep evaluation of C X and the forecast roll-out.  None of them is templated on the rank: they are
// bandwidth-trivial next to the filter itself and run once per sweep.
#pragma once
#include "psmf_common.cuh"

namespace psmf {

// ---- ingest: (d, n) row-major float64 with NaN = missing  ->  time-major (n, ld) ----------------------------
// The reference keeps Y as (d, n) and gathers the strided column Y[:, t] every step (rPSMF.py:101); the filter wants
// y_t contiguous.  One pass: 32 x 32 tiles through shared memory, coalesced on both sides.
//   mode bit 0 (INGEST_KEEP_NAN): missing entries stay NaN in Y_out (pairs with PSMF_NAN_MASK: no mask stream at all);
//                                 else they become 0 (rPSMF.py:200-202 / :163-164)
//   M_out (optional): 1 where the entry is observed (!isnan), the M of rPSMF.py:198
constexpr int INGEST_KEEP_NAN = 1;
template <typename T>
__global__ void ingest_kernel(const double* __restrict__ src, int64_t d, int64_t n, int mode, T* __restrict__ Y_out, int64_t ldy,
                              uint8_t* __restrict__ M_out, int64_t ldm) {
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.y * 32, t0 = (int64_t)blockIdx.x * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t i = i0 + k, t = t0 + threadIdx.x;
        tile[k][threadIdx.x] = (i < d && t < n) ? src[i * n + t] : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t t = t0 + k, i = i0 + threadIdx.x;
        if (t < n && i < d) {
            const double v = tile[threadIdx.x][k];
            const bool obs = !isnan(v);
            if (Y_out != nullptr) Y_out[t * ldy + i] = (T)((obs || (mode & INGEST_KEEP_NAN)) ? v : 0.0);
            if (M_out != nullptr) M_out[t * ldm + i] = obs ? 1 : 0;
        }
    }
}

// (d, n) row-major bytes -> time-major (n, ld) bytes, 1 where the source is non-zero (the reference's int / float 0-1
// masks M and Mmiss after a host-side narrowing to one byte)
__global__ void transpose_mask_kernel(const uint8_t* __restrict__ src, int64_t d, int64_t n, uint8_t* __restrict__ dst, int64_t ld) {
    __shared__ uint8_t tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.y * 32, t0 = (int64_t)blockIdx.x * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t i = i0 + k, t = t0 + threadIdx.x;
        tile[k][threadIdx.x] = (i < d && t < n) ? src[i * n + t] : 0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t t = t0 + k, i = i0 + threadIdx.x;
        if (t < n && i < d) dst[t * ld + i] = tile[threadIdx.x][k] != 0 ? 1 : 0;
    }
}

// ---- prepare_missing on the device (ExperimentImpute/common.py:50-76) ------------------------------------------
// One sweep of the reference's generator: row i loses the `seg` entries starting at starts[i] (those that are not
// already NaN): Y <- NaN, E <- 1.  The starts come from the caller's random stream (the host draws the d integers of
// a sweep exactly as the reference does; only the O(d * seg) application and the count run here); `count` receives
// the number of entries removed by this sweep (integer atomics: exact and order-independent).
template <typename T>
__global__ void missing_segments_kernel(T* __restrict__ Y, int64_t ldy, uint8_t* __restrict__ E, int64_t lde, int64_t d, int64_t n,
                                        const int64_t* __restrict__ starts, int seg, unsigned long long* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned removed = 0;
    if (i < d) {
        const int64_t s = starts[i];
        for (int64_t j = s; j < s + seg && j < n; ++j) {
            if (j < 0) continue;
            const T v = Y[j * ldy + i];
            if (!isnan(v)) {                                            // common.py:72-74
                Y[j * ldy + i] = (T)nan("");
                E[j * lde + i] = 1;
                ++removed;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    if ((threadIdx.x & 31) == 0 && removed) atomicAdd(count, (unsigned long long)removed);
}

// number of NaN entries of a time-major array (NumMissDefault of common.py:65)
template <typename T>
__global__ void count_nan_kernel(const T* __restrict__ Y, int64_t ldy, int64_t d, int64_t n, unsigned long long* __restrict__ count) {
    unsigned c = 0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < d * n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx / d, i = idx - t * d;
        c += isnan(Y[t * ldy + i]) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

// ---- products with the tiled dictionary: forecast (psmf.py:182-188) and the evaluation of C X (rPSMF.py:137-140) ----
// One warp per 32-row tile (lane = row, the row of C in registers), looping over the time / horizon index; the r-vectors
// are broadcast loads.  grid = (tile groups, series).
constexpr int TOOL_WARPS = 8;

template <typename T>
__device__ __forceinline__ void load_row(const T* __restrict__ Cs, int64_t tile, int R, int lane, double (&c)[MAXR]) {
    const T* gt = Cs + tile * (int64_t)(R * TILE);
#pragma unroll
    for (int j = 0; j < MAXR; ++j) c[j] = j < R ? (double)gt[tile_pos(j, lane)] : 0.0;
}

// Ypred[k][i] = c_i . Xp[k]      Xp: (n_series, n_pred, R), Ypred: (n_series, n_pred, ldp)
template <typename T>
__global__ void project_kernel(const T* __restrict__ C, int64_t c_series_stride, int64_t d, int R, const double* __restrict__ Xp,
                               int64_t n_pred, T* __restrict__ Ypred, int64_t ldp, int64_t psst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int series = blockIdx.y;
    const int64_t ntiles = (d + TILE - 1) / TILE;
    const T* Cs = C + (int64_t)series * c_series_stride;
    const double* X = Xp + (int64_t)series * n_pred * R;
    for (int64_t tile = (int64_t)blockIdx.x * TOOL_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * TOOL_WARPS) {
        double c[MAXR];
        load_row<T>(Cs, tile, R, lane, c);
        const int64_t row = tile * TILE + lane;
        for (int64_t k = 0; k < n_pred; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < MAXR; ++j)
                if (j < R) acc = fma(c[j], X[k * R + j], acc);
            if (row < d) Ypred[(int64_t)series * psst + k * ldp + row] = (T)acc;
        }
    }
}

// sum over the entries marked in E of (c_i . x_t - yorig_ti)^2 and their count: per-CTA partials, summed in fixed
// order by eval_full_finish (deterministic).  X: (n_series, n_steps, R) the filtered x_t of the sweep.
template <typename T>
__global__ void eval_full_kernel(const T* __restrict__ C, int64_t c_series_stride, int64_t d, int R, const double* __restrict__ X,
                                 int64_t n_steps, const T* __restrict__ Yorig, int64_t ldy, int64_t ysst,
                                 const uint8_t* __restrict__ E, int64_t lde, int64_t esst, double* __restrict__ partials) {
    __shared__ double red[TOOL_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int series = blockIdx.y;
    const int64_t ntiles = (d + TILE - 1) / TILE;
    const T* Cs = C + (int64_t)series * c_series_stride;
    const double* Xs = X + (int64_t)series * n_steps * R;
    double se = 0.0, ne = 0.0;
    for (int64_t tile = (int64_t)blockIdx.x * TOOL_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * TOOL_WARPS) {
        double c[MAXR];
        load_row<T>(Cs, tile, R, lane, c);
        const int64_t row = tile * TILE + lane;
        if (row >= d) continue;
        for (int64_t t = 0; t < n_steps; ++t) {
            if (E[(int64_t)series * esst + t * lde + row] == 0) continue;
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < MAXR; ++j)
                if (j < R) acc = fma(c[j], Xs[t * R + j], acc);
            const double dlt = acc - (double)Yorig[(int64_t)series * ysst + t * ldy + row];
            se = fma(dlt, dlt, se);
            ne += 1.0;
        }
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        ne += __shfl_xor_sync(0xffffffffu, ne, o);
    }
    if (lane == 0) {
        red[warp][0] = se;
        red[warp][1] = ne;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int w = 0; w < TOOL_WARPS; ++w) s += red[w][threadIdx.x];
        partials[((int64_t)series * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = s;
    }
}
__global__ void eval_full_finish(const double* __restrict__ partials, int nparts, double* __restrict__ out) {
    const int series = blockIdx.x;
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int k = 0; k < nparts; ++k) s += partials[((int64_t)series * nparts + k) * 2 + threadIdx.x];
        out[(int64_t)series * 2 + threadIdx.x] = s;
    }
}

// per-CTA evaluation records of a filter launch -> (n_series, NEVAL), fixed order over the CTAs of a series
__global__ void eval_part_finish(const double* __restrict__ part, int cps, double* __restrict__ out) {
    const int series = blockIdx.x;
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int k = 0; k < cps; ++k) s += part[((int64_t)series * cps + k) * NEVAL + threadIdx.x];
        out[(int64_t)series * NEVAL + threadIdx.x] = s;
    }
    if (threadIdx.x == 3) out[(int64_t)series * NEVAL + 3] = 0.0;
}

// forecast roll-out x_{k} = f(x_{k-1}), k = k0 .. k0 + n_pred - 1, from the engine's current x (psmf.py:182-186):
// one warp per series (r <= 16 values), sequential in k.
__global__ void rollout_kernel(const double* __restrict__ state, int R, int dynamics, const double* __restrict__ lin_A,
                               const double* __restrict__ lin_c, int64_t k0, int64_t n_pred, double* __restrict__ Xp) {
    const int series = blockIdx.x, lane = threadIdx.x;
    const double* stg = state + (int64_t)series * st_size(R);
    double x = lane < R ? stg[st_x(R) + lane] : 0.0;
    const double th = lane < R ? stg[st_theta(R) + lane] : 0.0;
    for (int64_t k = 0; k < n_pred; ++k) {
        double xn = x;
        if (dynamics == DYN_COS) {
            xn = cos(__dadd_rn(__dmul_rn(__dmul_rn(6.283185307179586, th), (double)(k0 + k)), x));
        } else if (dynamics == DYN_LINEAR) {
            double acc = (lin_c != nullptr && lane < R) ? lin_c[lane] : 0.0;
            for (int j = 0; j < R; ++j) {
                const double xj = __shfl_sync(0xffffffffu, x, j);
                if (lane < R) acc = fma(lin_A[lane * R + j], xj, acc);
            }
            xn = acc;
        }
        x = xn;
        if (lane < R) Xp[((int64_t)series * n_pred + k) * R + lane] = x;
    }
}

}  // namespace psmf
