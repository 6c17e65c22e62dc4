// Persistent PSMF / rPSMF filter kernel (sm_100a): one launch runs n_steps filter steps.
//
// Per step (reference ExperimentImpute/rPSMF.py:81-135, PSMF.py:60-84,
// pypsmf/psmf/psmf.py:90-165, rpsmf.py:116-171):
//
//   row pass   (all warps)  c_i += e_i(t-1) g(t-1)            rank-1 update of the PREVIOUS step, fused in
//                            yhat_i = c_i . xbar ; e_i = y_i - m_i yhat_i ; w_i = 1/(m_i rho + a)
//                            G += m w c c' ; b += m w e c ; s += w e^2 ; q1, q0, n_obs
//   reduce     warp butterfly -> CTA (fixed order) -> grid (fixed CTA order, one grid barrier)
//   small      K = (I + Pbar G)^-1 Pbar (Gauss-Jordan, partial pivoting, fp64), x, omega, P, eta, N, phi,
//              V, Q, rho, lambda, g = V xbar / N, then the predict half of the next step
//
// C is touched exactly once per step (read + write); e_i stays in shared memory between steps.
// Every CTA (and every GPU) derives the small state from bit-identical reduced statistics, so the
// replicated r x r state never diverges.
#pragma once
#include "psmf_common.cuh"

namespace psmf {

constexpr unsigned FULL = 0xffffffffu;

template <int R>
struct Smem {
    static constexpr int NG = ngroups_for(R);
    static constexpr int NSP = nstat_pad(R);
    double x[R];        // x_{t-1}, then x_t
    double xb[R];       // x_bar = f(x_{t-1})
    double fd[R];       // diagonal of F = df/dx (identity / cos dynamics)
    double vx[R];       // V x_bar
    double vxt[R];      // V' x_bar
    double g[R];        // rank-1 direction of the previous step
    double th[R];
    double P[R * R], V[R * R], Q[R * R], Pb[R * R];
    double aug[R][2 * R + 2];
    double tot[NSP];
    double part[NSP];
    double red[NG * NSP];
    double a, rho, lam;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// barrier 0 over the first `n` threads of the CTA (n == blockDim.x: plain __syncthreads; the streaming
// kernel excludes its producer warp)
__device__ __forceinline__ void sync_n(int n) { asm volatile("bar.sync 0, %0;" ::"r"(n) : "memory"); }

// All CTAs of the grid are co-resident (cooperative launch).  `target` grows monotonically.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target, int nthr) {
    sync_n(nthr);
    if (threadIdx.x == 0) {
        red_release_gpu(bar, 1ULL);
        while (ld_acquire_gpu(bar) < target) {
        }
    }
    sync_n(nthr);
}

__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;   // bit-identical on every lane (a+b == b+a at every stage)
}

// Transposing butterfly: reduces N per-lane accumulators over the 32 lanes with ~N shuffles instead of
// 5N.  On return lane l holds the full sums of entries [base, base + bfly_final(N)).
__host__ __device__ constexpr int bfly_final(int n) {
    for (int i = 0; i < 5; ++i) n = (n + 1) / 2;
    return n;
}
// Entries with index >= lim are padding (odd counts are rounded up at every stage) and must not be stored.
template <int N, int O, int NA>
__device__ __forceinline__ void bfly(double (&v)[NA], int lane, int& base, int& lim) {
    constexpr int H = (N + 1) / 2;
    const bool up = (lane & O) != 0;
    lim = min(lim, base + N);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const double lo = v[i];
        const double hi = (i + H < N) ? v[i + H] : 0.0;
        const double send = up ? lo : hi;
        const double keep = up ? hi : lo;
        v[i] = keep + __shfl_xor_sync(FULL, send, O);
    }
    if (up) base += H;
    if constexpr (O > 1) bfly<H, O / 2, NA>(v, lane, base, lim);
}

// ---- row pass of one role over the CTA's tiles ----------------------------------------------------
template <int R, int Q, typename T>
__device__ __forceinline__ void role_pass(const KParams& p, Smem<R>& sh, double* __restrict__ ebuf, T* __restrict__ Cs,
                                          const T* __restrict__ Yt, const uint8_t* __restrict__ Mt,
                                          T* __restrict__ Yrec_t, int tb, int te, int group, int lane) {
    constexpr int NS = nsplit_for(R), NG = ngroups_for(R), NSP = nstat_pad(R);
    constexpr int JB = split_begin(R, NS, Q), JE = split_begin(R, NS, Q + 1);
    constexpr int NGR = gram_off(R, JE) - gram_off(R, JB);
    constexpr int NACC = NGR + (Q == 0 ? R + 4 : 0);
    const double a = sh.a, rho = sh.rho;
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    for (int tile = tb + group; tile < te; tile += NG) {
        const int64_t row = (int64_t)tile * TILE + lane;
        const int rl = (tile - tb) * TILE + lane;
        T* ct = Cs + (int64_t)tile * (R * TILE) + lane;
        double c[R];
#pragma unroll
        for (int j = JB; j < R; ++j) c[j] = (double)ct[j * TILE];
        const double ep = ebuf[rl];
        const bool inb = row < p.d;
        bool mi = inb;
        if (Mt != nullptr && inb) mi = Mt[row] != 0;
        double yi = 0.0;
        if (Q == 0 && inb) yi = (double)Yt[row];
        // every role has read what it needs of this tile (C columns, e of the previous step) before any
        // role overwrites it
        if (NS > 1) named_bar_sync(1 + group, NS * 32);
#pragma unroll
        for (int j = JB; j < R; ++j) c[j] = fma(ep, sh.g[j], c[j]);          // rPSMF.py:111 (previous step)
#pragma unroll
        for (int j = JB; j < JE; ++j) ct[j * TILE] = (T)c[j];
        const double w = 1.0 / ((mi ? rho : 0.0) + a);                        // rPSMF.py:92,98,32
        const double mw = mi ? w : 0.0;
#pragma unroll
        for (int j = JB; j < JE; ++j) {
            const double cw = c[j] * mw;
#pragma unroll
            for (int k = j; k < R; ++k) {
                const int idx = gram_off(R, j) - gram_off(R, JB) + (k - j);
                acc[idx] = fma(cw, c[k], acc[idx]);                           // G = CM' Ri CM, rPSMF.py:35
            }
        }
        if (Q == 0) {
            double yh = 0.0;
#pragma unroll
            for (int j = 0; j < R; ++j) yh = fma(c[j], sh.xb[j], yh);         // rPSMF.py:89
            const double e = yi - (mi ? yh : 0.0);                            // rPSMF.py:101
            ebuf[rl] = e;
            if (Yrec_t != nullptr && inb) Yrec_t[row] = (T)yh;
            const double ew = e * mw;
#pragma unroll
            for (int j = 0; j < R; ++j) acc[NGR + j] = fma(ew, c[j], acc[NGR + j]);
            const double e2 = e * e;
            acc[NGR + R + 0] = fma(inb ? w : 0.0, e2, acc[NGR + R + 0]);      // diff' Ri diff
            acc[NGR + R + 1] += mi ? e2 : 0.0;                                // rPSMF.py:112-114 (observed rows)
            acc[NGR + R + 2] += mi ? 0.0 : e2;                                //                  (missing rows)
            acc[NGR + R + 3] += mi ? 1.0 : 0.0;
        }
    }

    int base = 0, lim = NACC;
    bfly<NACC, 16, NACC>(acc, lane, base, lim);
    constexpr int NF = bfly_final(NACC);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        const int li = base + i;
        if (li < lim) {
            int gi;
            if (Q == 0)
                gi = li < NGR ? li : ngram(R) + (li - NGR);
            else
                gi = gram_off(R, JB) + li;
            sh.red[group * NSP + gi] = acc[i];
        }
    }
}

// apply the pending rank-1 update of the last step so that C in HBM is the filtered C_T
template <int R, int Q, typename T>
__device__ __forceinline__ void role_flush(Smem<R>& sh, const double* __restrict__ ebuf, T* __restrict__ Cs, int tb,
                                           int te, int group, int lane) {
    constexpr int NS = nsplit_for(R), NG = ngroups_for(R);
    constexpr int JB = split_begin(R, NS, Q), JE = split_begin(R, NS, Q + 1);
    for (int tile = tb + group; tile < te; tile += NG) {
        const int rl = (tile - tb) * TILE + lane;
        T* ct = Cs + (int64_t)tile * (R * TILE) + lane;
        const double ep = ebuf[rl];
#pragma unroll
        for (int j = JB; j < JE; ++j) ct[j * TILE] = (T)fma(ep, sh.g[j], (double)ct[j * TILE]);
    }
}

template <int R, typename T, int Q>
__device__ __forceinline__ void dispatch_pass(int role, const KParams& p, Smem<R>& sh, double* ebuf, T* Cs, const T* Yt,
                                              const uint8_t* Mt, T* Yrec_t, int tb, int te, int group, int lane) {
    if (role == Q) {
        role_pass<R, Q, T>(p, sh, ebuf, Cs, Yt, Mt, Yrec_t, tb, te, group, lane);
        return;
    }
    if constexpr (Q + 1 < nsplit_for(R)) dispatch_pass<R, T, Q + 1>(role, p, sh, ebuf, Cs, Yt, Mt, Yrec_t, tb, te, group, lane);
}
template <int R, typename T, int Q>
__device__ __forceinline__ void dispatch_flush(int role, Smem<R>& sh, const double* ebuf, T* Cs, int tb, int te,
                                               int group, int lane) {
    if (role == Q) {
        role_flush<R, Q, T>(sh, ebuf, Cs, tb, te, group, lane);
        return;
    }
    if constexpr (Q + 1 < nsplit_for(R)) dispatch_flush<R, T, Q + 1>(role, sh, ebuf, Cs, tb, te, group, lane);
}

// ---- predict half: x_bar = f(x), P_bar = F P F' + Q, V x_bar, a (warp 0) ---------------------------
template <int R>
__device__ void predict(const KParams& p, Smem<R>& sh, int lane, int64_t k, int series) {
    const bool simp = (p.flags & F_SIMPLIFIED) != 0;
    if (lane < R) {
        double xb, fd = 1.0;
        if (p.dynamics == DYN_COS) {                                          // synthetic_psmf.py:105-106
            const double arg = __dadd_rn(__dmul_rn(__dmul_rn(6.283185307179586, sh.th[lane]), (double)k), sh.x[lane]);
            xb = cos(arg);
            fd = -sin(arg);
        } else if (p.dynamics == DYN_EXTERNAL) {
            xb = p.xbar_ext[(int64_t)series * R + lane];
        } else {
            xb = sh.x[lane];                                                  // rPSMF.py:86
        }
        sh.xb[lane] = xb;
        sh.fd[lane] = fd;
    }
    __syncwarp();
    double av = 0.0;
    if (lane < R) {
        const int j = lane;
        if (simp) {                                                           // synthetic_psmf.py:83-84
#pragma unroll
            for (int i = 0; i < R; ++i) sh.Pb[i * R + j] = sh.P[i * R + j];
        } else if (p.dynamics == DYN_EXTERNAL) {                              // psmf.py:115 with a dense F
            const double* F = p.F_ext + (int64_t)series * R * R;
            double tmp[R];
#pragma unroll
            for (int k2 = 0; k2 < R; ++k2) {
                double s = 0.0;
                for (int l = 0; l < R; ++l) s = fma(sh.P[k2 * R + l], F[j * R + l], s);
                tmp[k2] = s;
            }
            for (int i = 0; i < R; ++i) {
                double s = 0.0;
#pragma unroll
                for (int k2 = 0; k2 < R; ++k2) s = fma(F[i * R + k2], tmp[k2], s);
                sh.Pb[i * R + j] = s + sh.Q[i * R + j];
            }
        } else {                                                              // rPSMF.py:87 / psmf.py:115
#pragma unroll
            for (int i = 0; i < R; ++i) sh.Pb[i * R + j] = sh.fd[i] * sh.P[i * R + j] * sh.fd[j] + sh.Q[i * R + j];
        }
        double v1 = 0.0, v2 = 0.0;
#pragma unroll
        for (int k2 = 0; k2 < R; ++k2) {
            v1 = fma(sh.V[j * R + k2], sh.xb[k2], v1);
            v2 = fma(sh.V[k2 * R + j], sh.xb[k2], v2);
        }
        sh.vx[j] = v1;
        sh.vxt[j] = v2;
        av = sh.xb[j] * v1;
    }
    av = warp_allsum(av);                                                     // rPSMF.py:93
    if (lane == 0) sh.a = av;
    __syncwarp();
}

// Gauss-Jordan with partial pivoting on the R x (2R+1) augmented matrix [I + Pbar G | Pbar | Pbar b] (warp 0).
template <int R>
__device__ void gauss_jordan(double (*aug)[2 * R + 2], int lane) {
    constexpr int NC = 2 * R + 1;
    for (int k = 0; k < R; ++k) {
        double v = (lane >= k && lane < R) ? fabs(aug[lane][k]) : -1.0;
        int pi = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL, v, o);
            const int oi = __shfl_xor_sync(FULL, pi, o);
            if (ov > v || (ov == v && oi < pi)) {
                v = ov;
                pi = oi;
            }
        }
        const double inv = 1.0 / aug[pi][k];
        __syncwarp();
        for (int c = lane; c < NC; c += 32) {
            const double tk = aug[k][c], tp = aug[pi][c];
            aug[pi][c] = tk;
            aug[k][c] = tp * inv;
        }
        __syncwarp();
        double f[R];
#pragma unroll
        for (int i = 0; i < R; ++i) f[i] = aug[i][k];
        __syncwarp();
        for (int c = lane; c < NC; c += 32) {
            const double rk = aug[k][c];
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (i != k) aug[i][c] = fma(-f[i], rk, aug[i][c]);
        }
        __syncwarp();
    }
}

// ---- r x r part of the step (rPSMF.py:102-115,133-135), identical on every CTA ---------------------
template <int R>
__device__ void small_update(const KParams& p, Smem<R>& sh, int tid, int lane, int warp, int series, int64_t t,
                             bool writer, int nthr) {
    constexpr int NGm = ngram(R);
    const bool simp = (p.flags & F_SIMPLIFIED) != 0;
    const bool robust = (p.flags & F_ROBUST) != 0;
    const double* tot = sh.tot;
    if (!simp) {
        for (int idx = tid; idx < R * R; idx += nthr) {
            const int i = idx / R, j = idx % R;
            double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int lo = k < j ? k : j, hi = k < j ? j : k;
                acc = fma(sh.Pb[i * R + k], tot[gram_off(R, lo) + hi - lo], acc);
            }
            sh.aug[i][j] = acc;
            sh.aug[i][R + j] = sh.Pb[i * R + j];
        }
        if (tid < R) {                                   // right-hand side Pbar b: the solve then yields K b
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) acc = fma(sh.Pb[tid * R + k], tot[NGm + k], acc);
            sh.aug[tid][2 * R] = acc;
        }
    }
    sync_n(nthr);
    if (warp == 0) {
        if (!simp) gauss_jordan<R>(sh.aug, lane);     // aug[:, R:2R] = K, aug[:, 2R] = K b
        const double a = sh.a, rho = sh.rho, lam = sh.lam;
        const double s = tot[NGm + R], q1 = tot[NGm + R + 1], q0 = tot[NGm + R + 2], nobs = tot[NGm + R + 3];
        const double dg = (double)p.d_global;
        double xn = 0.0, bkb = 0.0, trpg = 0.0;
        if (lane < R) {
            if (simp) {
                xn = sh.xb[lane];                                              // synthetic_psmf.py:93-94
            } else {
                const double kb = sh.aug[lane][2 * R];
                xn = sh.xb[lane] + kb;                                         // rPSMF.py:104
                bkb = tot[NGm + lane] * kb;
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int lo = i < lane ? i : lane, hi = i < lane ? lane : i;
                    trpg = fma(sh.Pb[i * R + lane], tot[gram_off(R, lo) + hi - lo], trpg);
                }
            }
        }
        bkb = warp_allsum(bkb);
        trpg = warp_allsum(trpg);
        double sSe, eta;
        if (simp) {
            sSe = s;                                                           // synthetic_rpsmf.py:93-98
            eta = rho;                                                         // synthetic_psmf.py:86-87
        } else {
            sSe = s - bkb;                                                     // diff' CPinv diff
            eta = (rho * nobs + (rho + a) * trpg) / dg;                        // rPSMF.py:108
        }
        const double omega = robust ? (lam + sSe) / (lam + dg) : 1.0;          // rPSMF.py:105
        const double N = a + eta;                                              // rPSMF.py:109
        const double phi = robust ? (lam + q1 / (a + eta) + (q0 != 0.0 ? q0 / eta : 0.0)) / (lam + dg) : 1.0;
        const double aphi = p.alpha * phi, bom = p.beta * omega;
        if (lane < R) {
            const int j = lane;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const double pn = simp ? sh.Pb[i * R + j] : bom * sh.aug[i][R + j];          // rPSMF.py:106
                const double vn = aphi * (sh.V[i * R + j] - sh.vx[i] * sh.vxt[j] / N);       // rPSMF.py:115
                sh.P[i * R + j] = pn;
                sh.V[i * R + j] = vn;
                if (!simp) sh.Q[i * R + j] = omega * sh.Q[i * R + j];                        // rPSMF.py:133
            }
            sh.x[j] = xn;
            sh.g[j] = (((p.flags & F_CUPDATE_VT) != 0) ? sh.vx[j] : sh.vxt[j]) / N;          // rPSMF.py:111 / PSMF.py:80
            if (writer && p.X_out != nullptr) p.X_out[((int64_t)series * p.n_steps + t) * R + j] = xn;
        }
        if (lane == 0) {
            sh.rho = omega * rho;                                              // rPSMF.py:134
            sh.lam = (robust && (p.flags & F_FIXED_LAMBDA) == 0) ? lam + dg : lam;   // rPSMF.py:135
            if (writer) {
                if (p.scal_out != nullptr) {
                    double* so = p.scal_out + ((int64_t)series * p.n_steps + t) * NSCAL;
                    so[0] = a; so[1] = eta; so[2] = N; so[3] = omega; so[4] = phi; so[5] = sSe; so[6] = lam; so[7] = rho;
                }
                if (!(isfinite(N) && isfinite(omega) && isfinite(phi) && isfinite(xn)) || N == 0.0)
                    atomicCAS((unsigned long long*)p.status, ~0ULL, (unsigned long long)t);
            }
        }
        __syncwarp();
        if (t + 1 < p.n_steps) predict<R>(p, sh, lane, p.k0 + t + 1, series);
    }
    sync_n(nthr);
}

template <int R, typename T>
__global__ void __launch_bounds__(nsplit_for(R) * ngroups_for(R) * 32, 1) psmf_filter_kernel(const KParams p) {
    constexpr int NS = nsplit_for(R), NG = ngroups_for(R), NSP = nstat_pad(R), NST = nstat(R);
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ Smem<R> sh;
    double* ebuf = reinterpret_cast<double*>(dyn_smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = warp / NS;
    const int role = ((warp % NS) + group) % NS;      // spread the roles over the 4 SM sub-partitions
    const int series = blockIdx.x / p.cps;
    const int part = blockIdx.x % p.cps;
    const int ntiles = (int)((p.d + TILE - 1) / TILE);
    const int tb = (int)((int64_t)ntiles * part / p.cps);
    const int te = (int)((int64_t)ntiles * (part + 1) / p.cps);
    const bool writer = part == 0;

    T* Cs = reinterpret_cast<T*>(p.C) + (int64_t)series * p.c_series_stride;
    double* stg = p.state + (int64_t)series * st_size(R);

    for (int i = tid; i < R * R; i += blockDim.x) {
        sh.P[i] = stg[st_P(R) + i];
        sh.V[i] = stg[st_V(R) + i];
        sh.Q[i] = stg[st_Q(R) + i];
    }
    if (tid < R) {
        sh.x[tid] = stg[st_x(R) + tid];
        sh.th[tid] = stg[st_theta(R) + tid];
        sh.g[tid] = 0.0;
    }
    if (tid == 0) {
        sh.rho = stg[st_rho(R)];
        sh.lam = stg[st_lam(R)];
    }
    for (int i = tid; i < (te - tb) * TILE; i += blockDim.x) ebuf[i] = 0.0;
    __syncthreads();
    if (warp == 0) predict<R>(p, sh, lane, p.k0, series);
    __syncthreads();

    for (int64_t t = 0; t < p.n_steps; ++t) {
        const T* Yt = reinterpret_cast<const T*>(p.Y) + (int64_t)series * p.ysst + t * p.ldy;
        const uint8_t* Mt = p.M ? p.M + (int64_t)series * p.msst + t * p.ldm : nullptr;
        T* Yrec_t = p.Yrec ? reinterpret_cast<T*>(p.Yrec) + (int64_t)series * p.recsst + t * p.ldrec : nullptr;
        dispatch_pass<R, T, 0>(role, p, sh, ebuf, Cs, Yt, Mt, Yrec_t, tb, te, group, lane);
        __syncthreads();
        // CTA partial: fixed order over the row groups
        if (tid < NST) {
            double s = 0.0;
#pragma unroll
            for (int g = 0; g < NG; ++g) s += sh.red[g * NSP + tid];
            sh.part[tid] = s;
        }
        if (p.cps > 1) {
            const int parity = (int)(t & 1);
            double* mine = p.partials + ((size_t)parity * gridDim.x + blockIdx.x) * NSP;
            if (tid < NST) mine[tid] = sh.part[tid];
            grid_barrier(p.bar, (unsigned long long)gridDim.x * (unsigned long long)(t + 1), blockDim.x);
            const double* basep = p.partials + ((size_t)parity * gridDim.x + (size_t)series * p.cps) * NSP;
            if (tid < NST) {
                // fixed summation order: 4 interleaved chains over the CTAs, combined as (s0+s1)+(s2+s3)
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int c = 0;
                for (; c + 3 < p.cps; c += 4) {
                    s0 += __ldcg(basep + (size_t)(c + 0) * NSP + tid);
                    s1 += __ldcg(basep + (size_t)(c + 1) * NSP + tid);
                    s2 += __ldcg(basep + (size_t)(c + 2) * NSP + tid);
                    s3 += __ldcg(basep + (size_t)(c + 3) * NSP + tid);
                }
                for (; c < p.cps; ++c) s0 += __ldcg(basep + (size_t)c * NSP + tid);
                sh.tot[tid] = (s0 + s1) + (s2 + s3);
            }
        } else {
            if (tid < NST) sh.tot[tid] = sh.part[tid];
        }
        __syncthreads();
        small_update<R>(p, sh, tid, lane, warp, series, t, writer, blockDim.x);
    }

    dispatch_flush<R, T, 0>(role, sh, ebuf, Cs, tb, te, group, lane);
    if (writer) {
        for (int i = tid; i < R * R; i += blockDim.x) {
            stg[st_P(R) + i] = sh.P[i];
            stg[st_V(R) + i] = sh.V[i];
            stg[st_Q(R) + i] = sh.Q[i];
        }
        if (tid < R) stg[st_x(R) + tid] = sh.x[tid];
        if (tid == 0) {
            stg[st_rho(R)] = sh.rho;
            stg[st_lam(R)] = sh.lam;
        }
    }
}

}  // namespace psmf
