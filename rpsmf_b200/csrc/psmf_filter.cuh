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
    double aug[2][R][2 * R + 2];   // double-buffered augmented matrix of the r x r solve
    double tot[NSP];
    double part[NSP];
    double red[NG * NSP];          // per row-group partial statistics
    double a, rho, lam;
    double w1, w0;                 // 1/(rho + a), 1/a : the only two values of w_i (rPSMF.py:92,98,32)
    double sc[8];                  // omega, eta, N, phi, sSe, alpha*phi, beta*omega
    int perm[R];                   // pivot row of elimination step k
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

// debug phase stamps (CTA 0, thread 0): enabled when KParams.trace != nullptr
__device__ __forceinline__ void stamp(const KParams& p, int64_t t, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && t < p.trace_steps) {
        unsigned long long v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        p.trace[t * 8 + slot] = v;
    }
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
    const double w1 = sh.w1, w0 = sh.w0;
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
        const double w = mi ? w1 : w0;                                        // rPSMF.py:92,98,32
        const double mw = mi ? w1 : 0.0;
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
            double yh4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < R; ++j) yh4[j & 3] = fma(c[j], sh.xb[j], yh4[j & 3]);   // rPSMF.py:89
            const double yh = (yh4[0] + yh4[1]) + (yh4[2] + yh4[3]);
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

// ---- predict half: x_bar = f(x), P_bar = F P F' + Q, V x_bar, a (all `nthr` threads; ends with a barrier) ----
template <int R>
__device__ void predict_cta(const KParams& p, Smem<R>& sh, int tid, int64_t k, int series, int nthr) {
    const bool simp = (p.flags & F_SIMPLIFIED) != 0;
    const int lane = tid & 31;
    if (tid < R) {
        double xb, fd = 1.0;
        if (p.dynamics == DYN_COS) {                                          // synthetic_psmf.py:105-106
            const double arg = __dadd_rn(__dmul_rn(__dmul_rn(6.283185307179586, sh.th[tid]), (double)k), sh.x[tid]);
            xb = cos(arg);
            fd = -sin(arg);
        } else if (p.dynamics == DYN_EXTERNAL) {
            xb = p.xbar_ext[(int64_t)series * R + tid];
        } else {
            xb = sh.x[tid];                                                   // rPSMF.py:86
        }
        sh.xb[tid] = xb;
        sh.fd[tid] = fd;
    }
    sync_n(nthr);
    if (tid >= nthr - 32) {
        // last warp: V x_bar, V' x_bar, a = x_bar' V x_bar and the two possible row weights
        double v = 0.0;
        if (lane < 2 * R) {
            const int j = lane < R ? lane : lane - R;
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int k2 = 0; k2 < R; k2 += 2) {
                v0 = fma(lane < R ? sh.V[j * R + k2] : sh.V[k2 * R + j], sh.xb[k2], v0);
                if (k2 + 1 < R) v1 = fma(lane < R ? sh.V[j * R + k2 + 1] : sh.V[(k2 + 1) * R + j], sh.xb[k2 + 1], v1);
            }
            v = v0 + v1;
            if (lane < R) sh.vx[j] = v; else sh.vxt[j] = v;
        }
        double av = lane < R ? sh.xb[lane] * v : 0.0;
        av = warp_allsum(av);                                                 // rPSMF.py:93
        if (lane == 0) {
            sh.a = av;
            sh.w1 = 1.0 / (sh.rho + av);
            sh.w0 = 1.0 / av;
        }
    }
    for (int idx = tid; idx < R * R; idx += nthr) {
        const int i = idx / R, j = idx % R;
        double pb;
        if (simp) {                                                           // synthetic_psmf.py:83-84
            pb = sh.P[idx];
        } else if (p.dynamics == DYN_EXTERNAL) {                              // psmf.py:115 with a dense F
            const double* F = p.F_ext + (int64_t)series * R * R;
            double acc = 0.0;
            for (int k2 = 0; k2 < R; ++k2) {
                double t2 = 0.0;
                for (int l = 0; l < R; ++l) t2 = fma(sh.P[k2 * R + l], F[j * R + l], t2);
                acc = fma(F[i * R + k2], t2, acc);
            }
            pb = acc + sh.Q[idx];
        } else {                                                              // rPSMF.py:87 / psmf.py:115
            pb = sh.fd[i] * sh.P[idx] * sh.fd[j] + sh.Q[idx];
        }
        sh.Pb[idx] = pb;
    }
    sync_n(nthr);
}

// Gauss-Jordan with partial pivoting on the R x (2R+1) augmented matrix [I + Pbar G | Pbar | Pbar b], one
// element per thread and one barrier per elimination step.  Rows are not swapped: step k eliminates with
// the unused row of largest |a_ik| (every thread finds it redundantly -> no broadcast barrier) and
// perm[k] remembers it, so the solution row of unknown k is aug[R & 1][perm[k]].
template <int R>
__device__ void gauss_jordan_cta(Smem<R>& sh, int tid, int nthr) {
    constexpr int NC = 2 * R + 1;
    unsigned used = 0;
    for (int k = 0; k < R; ++k) {
        const int cur = k & 1, nxt = cur ^ 1;
        double v[R];
        int ix[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            v[i] = ((used >> i) & 1u) ? -1.0 : fabs(sh.aug[cur][i][k]);
            ix[i] = i;
        }
#pragma unroll
        for (int st = 1; st < R; st <<= 1) {
#pragma unroll
            for (int i = 0; i + st < R; i += 2 * st) {
                if (v[i + st] > v[i]) {                 // strict: ties keep the lower row index
                    v[i] = v[i + st];
                    ix[i] = ix[i + st];
                }
            }
        }
        const int pi = ix[0];
        const double inv = 1.0 / sh.aug[cur][pi][k];
        used |= 1u << pi;
        if (tid == 0) sh.perm[k] = pi;
        for (int idx = tid; idx < R * NC; idx += nthr) {
            const int i = idx / NC, c = idx - i * NC;
            const double rpc = sh.aug[cur][pi][c] * inv;
            const double aic = sh.aug[cur][i][c];
            const double aik = sh.aug[cur][i][k];
            sh.aug[nxt][i][c] = (i == pi) ? rpc : fma(-aik, rpc, aic);
        }
        sync_n(nthr);
    }
}

// ---- r x r part of the step (rPSMF.py:102-115,133-135), identical on every CTA; all `nthr` threads ----
template <int R>
__device__ void small_update(const KParams& p, Smem<R>& sh, int tid, int lane, int warp, int series, int64_t t,
                             bool writer, int nthr) {
    constexpr int NGm = ngram(R);
    constexpr int FIN = R & 1;                      // buffer holding the result of the elimination
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
            sh.aug[0][i][j] = acc;
            sh.aug[0][i][R + j] = sh.Pb[i * R + j];
        }
        if (tid >= nthr - R) {                           // right-hand side Pbar b: the solve then yields K b
            const int i = tid - (nthr - R);
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) acc = fma(sh.Pb[i * R + k], tot[NGm + k], acc);
            sh.aug[0][i][2 * R] = acc;
        }
        sync_n(nthr);
        gauss_jordan_cta<R>(sh, tid, nthr);              // aug[FIN][perm[k]][R..2R) = K[k][:], [..][2R] = (K b)[k]
    }
    stamp(p, t, 7);
    if (warp == 0) {
        const double a = sh.a, rho = sh.rho, lam = sh.lam;
        const double s = tot[NGm + R], q1 = tot[NGm + R + 1], q0 = tot[NGm + R + 2], nobs = tot[NGm + R + 3];
        const double dg = (double)p.d_global;
        double xn = 0.0, bkb = 0.0, trpg = 0.0;
        if (lane < R) {
            if (simp) {
                xn = sh.xb[lane];                                              // synthetic_psmf.py:93-94
            } else {
                const double kb = sh.aug[FIN][sh.perm[lane]][2 * R];
                xn = sh.xb[lane] + kb;                                         // rPSMF.py:104
                bkb = tot[NGm + lane] * kb;
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int i = 0; i < R; i += 2) {
                    const int lo = i < lane ? i : lane, hi = i < lane ? lane : i;
                    t0 = fma(sh.Pb[i * R + lane], tot[gram_off(R, lo) + hi - lo], t0);
                    if (i + 1 < R) {
                        const int lo1 = i + 1 < lane ? i + 1 : lane, hi1 = i + 1 < lane ? lane : i + 1;
                        t1 = fma(sh.Pb[(i + 1) * R + lane], tot[gram_off(R, lo1) + hi1 - lo1], t1);
                    }
                }
                trpg = t0 + t1;
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
        if (lane < R) {
            sh.x[lane] = xn;
            if (writer && p.X_out != nullptr) p.X_out[((int64_t)series * p.n_steps + t) * R + lane] = xn;
        }
        if (lane == 0) {
            sh.sc[0] = omega; sh.sc[1] = eta; sh.sc[2] = N; sh.sc[3] = phi; sh.sc[4] = sSe;
            sh.sc[5] = p.alpha * phi; sh.sc[6] = p.beta * omega;
            if (writer) {
                if (p.scal_out != nullptr) {
                    double* so = p.scal_out + ((int64_t)series * p.n_steps + t) * NSCAL;
                    so[0] = a; so[1] = eta; so[2] = N; so[3] = omega; so[4] = phi; so[5] = sSe; so[6] = lam; so[7] = rho;
                }
                if (!(isfinite(N) && isfinite(omega) && isfinite(phi) && isfinite(xn)) || N == 0.0)
                    atomicCAS((unsigned long long*)p.status, ~0ULL, (unsigned long long)t);
            }
        }
    }
    sync_n(nthr);
    {
        const double omega = sh.sc[0], N = sh.sc[2], aphi = sh.sc[5], bom = sh.sc[6];
        for (int idx = tid; idx < R * R; idx += nthr) {
            const int i = idx / R, j = idx % R;
            const double pn = simp ? sh.Pb[idx] : bom * sh.aug[FIN][sh.perm[i]][R + j];      // rPSMF.py:106
            const double vn = aphi * (sh.V[idx] - sh.vx[i] * sh.vxt[j] / N);                 // rPSMF.py:115
            sh.P[idx] = pn;
            sh.V[idx] = vn;
            if (!simp) sh.Q[idx] = omega * sh.Q[idx];                                        // rPSMF.py:133
        }
        if (tid >= nthr - R) {
            const int j = tid - (nthr - R);
            sh.g[j] = (((p.flags & F_CUPDATE_VT) != 0) ? sh.vx[j] : sh.vxt[j]) / N;          // rPSMF.py:111 / PSMF.py:80
        }
        if (tid == nthr - 32) {
            const double lam = sh.lam;
            sh.rho = omega * sh.rho;                                                         // rPSMF.py:134
            sh.lam = (robust && (p.flags & F_FIXED_LAMBDA) == 0) ? lam + (double)p.d_global : lam;   // rPSMF.py:135
        }
    }
    sync_n(nthr);
    if (t + 1 < p.n_steps) predict_cta<R>(p, sh, tid, p.k0 + t + 1, series, nthr);
}

// ---- cross-CTA reduction of the statistics, deterministic (fixed summation order) ---------------------
//   cps <= 16 : every CTA reads all partials after one grid barrier
//   cps  > 16 : reduce-scatter / all-gather through L2: CTA c sums entries {c, c + cps, ..} over all CTAs
//               (coalesced reads of a transposed partial array), a second barrier publishes the totals.
// sh.part -> sh.tot; ends with a barrier over the `nthr` threads.
template <int R>
__device__ __forceinline__ void grid_reduce(const KParams& p, Smem<R>& sh, int tid, int lane, int warp, int64_t t,
                                            int series, int part, int nthr) {
    constexpr int NSP = nstat_pad(R), NST = nstat(R);
    if (p.cps == 1) {
        if (tid < NST) sh.tot[tid] = sh.part[tid];
        sync_n(nthr);
        return;
    }
    const int parity = (int)(t & 1);
    const int cps = p.cps;
    if (cps <= 16) {
        double* mine = p.partials + ((size_t)parity * cps + part) * NSP;
        if (tid < NST) mine[tid] = sh.part[tid];
        stamp(p, t, 3);
        grid_barrier(p.bar, (unsigned long long)gridDim.x * (unsigned long long)(t + 1), nthr);
        stamp(p, t, 4);
        const double* basep = p.partials + (size_t)parity * cps * NSP;
        if (tid < NST) {
            double v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = c < cps ? __ldcg(basep + (size_t)c * NSP + tid) : 0.0;
            sh.tot[tid] = (((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]))) +
                          (((v[8] + v[9]) + (v[10] + v[11])) + ((v[12] + v[13]) + (v[14] + v[15])));
        }
    } else {
        const int pstr = (cps + 7) & ~7;
        double* pT = p.partials + (size_t)parity * NSP * (pstr + 1);
        double* totals = pT + (size_t)NSP * pstr;
        if (tid < NST) pT[(size_t)tid * pstr + part] = sh.part[tid];
        stamp(p, t, 3);
        grid_barrier(p.bar, (unsigned long long)gridDim.x * (unsigned long long)(2 * t + 1), nthr);
        const int nw = nthr >> 5;
        for (int e = part + warp * cps; e < NST; e += nw * cps) {
            const double* src = pT + (size_t)e * pstr;
            double s = 0.0;
            for (int c0 = 0; c0 < cps; c0 += 256) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + lane + 32 * j;
                    v[j] = c < cps ? __ldcg(src + c) : 0.0;
                }
                s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            }
            s = warp_allsum(s);
            if (lane == 0) totals[e] = s;
        }
        grid_barrier(p.bar, (unsigned long long)gridDim.x * (unsigned long long)(2 * t + 2), nthr);
        stamp(p, t, 4);
        if (tid < NST) sh.tot[tid] = __ldcg(totals + tid);
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
    predict_cta<R>(p, sh, tid, p.k0, series, blockDim.x);

    for (int64_t t = 0; t < p.n_steps; ++t) {
        const T* Yt = reinterpret_cast<const T*>(p.Y) + (int64_t)series * p.ysst + t * p.ldy;
        const uint8_t* Mt = p.M ? p.M + (int64_t)series * p.msst + t * p.ldm : nullptr;
        T* Yrec_t = p.Yrec ? reinterpret_cast<T*>(p.Yrec) + (int64_t)series * p.recsst + t * p.ldrec : nullptr;
        stamp(p, t, 0);
        dispatch_pass<R, T, 0>(role, p, sh, ebuf, Cs, Yt, Mt, Yrec_t, tb, te, group, lane);
        stamp(p, t, 1);
        __syncthreads();
        stamp(p, t, 2);
        // CTA partial: fixed order over the row groups
        if (tid < NST) {
            double s = 0.0;
#pragma unroll
            for (int g = 0; g < NG; ++g) s += sh.red[g * NSP + tid];
            sh.part[tid] = s;
        }
        grid_reduce<R>(p, sh, tid, lane, warp, t, series, part, blockDim.x);
        stamp(p, t, 5);
        small_update<R>(p, sh, tid, lane, warp, series, t, writer, blockDim.x);
        stamp(p, t, 6);
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
