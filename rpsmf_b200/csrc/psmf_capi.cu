// C ABI of the B200-native PSMF / rPSMF filter (include/psmf_b200.h).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <new>
#include <string>

#include "../../include/psmf_b200.h"
#include "psmf_common.cuh"

PSMF_DECLARE_R(1) PSMF_DECLARE_R(2) PSMF_DECLARE_R(3) PSMF_DECLARE_R(4)
PSMF_DECLARE_R(5) PSMF_DECLARE_R(6) PSMF_DECLARE_R(7) PSMF_DECLARE_R(8)
PSMF_DECLARE_R(9) PSMF_DECLARE_R(10) PSMF_DECLARE_R(11) PSMF_DECLARE_R(12)
PSMF_DECLARE_R(13) PSMF_DECLARE_R(14) PSMF_DECLARE_R(15) PSMF_DECLARE_R(16)

namespace psmf {

static const launch_fn LAUNCH[MAXR + 1] = {
    nullptr,           launch_filter_r1,  launch_filter_r2,  launch_filter_r3,  launch_filter_r4,  launch_filter_r5,
    launch_filter_r6,  launch_filter_r7,  launch_filter_r8,  launch_filter_r9,  launch_filter_r10, launch_filter_r11,
    launch_filter_r12, launch_filter_r13, launch_filter_r14, launch_filter_r15, launch_filter_r16};
static const launch_fn LAUNCH_S[MAXR + 1] = {
    nullptr,           launch_stream_r1,  launch_stream_r2,  launch_stream_r3,  launch_stream_r4,  launch_stream_r5,
    launch_stream_r6,  launch_stream_r7,  launch_stream_r8,  launch_stream_r9,  launch_stream_r10, launch_stream_r11,
    launch_stream_r12, launch_stream_r13, launch_stream_r14, launch_stream_r15, launch_stream_r16};
static const shape_fn SHAPE_S[MAXR + 1] = {
    nullptr,          shape_stream_r1,  shape_stream_r2,  shape_stream_r3,  shape_stream_r4,  shape_stream_r5,
    shape_stream_r6,  shape_stream_r7,  shape_stream_r8,  shape_stream_r9,  shape_stream_r10, shape_stream_r11,
    shape_stream_r12, shape_stream_r13, shape_stream_r14, shape_stream_r15, shape_stream_r16};
static const shape_fn SHAPE[MAXR + 1] = {
    nullptr,          shape_filter_r1,  shape_filter_r2,  shape_filter_r3,  shape_filter_r4,  shape_filter_r5,
    shape_filter_r6,  shape_filter_r7,  shape_filter_r8,  shape_filter_r9,  shape_filter_r10, shape_filter_r11,
    shape_filter_r12, shape_filter_r13, shape_filter_r14, shape_filter_r15, shape_filter_r16};

// (n_series, d, R) row-major  <->  tiled [tile][R][32]
template <typename T>
__global__ void pack_C(const T* __restrict__ src, T* __restrict__ dst, int64_t d, int R, int64_t ntiles, int64_t total) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(idx % TILE);
        const int j = (int)((idx / TILE) % R);
        const int64_t tile = (idx / (TILE * (int64_t)R)) % ntiles;
        const int64_t s = idx / (TILE * (int64_t)R * ntiles);
        const int64_t row = tile * TILE + (l ^ ((j & 7) << 2));        // inverse of tile_pos (XOR swizzle)
        dst[idx] = row < d ? src[(s * d + row) * R + j] : (T)0;
    }
}
template <typename T>
__global__ void unpack_C(const T* __restrict__ src, T* __restrict__ dst, int64_t d, int R, int64_t ntiles, int64_t total) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(idx % TILE);
        const int j = (int)((idx / TILE) % R);
        const int64_t tile = (idx / (TILE * (int64_t)R)) % ntiles;
        const int64_t s = idx / (TILE * (int64_t)R * ntiles);
        const int64_t row = tile * TILE + (l ^ ((j & 7) << 2));
        if (row < d) dst[(s * d + row) * R + j] = src[idx];
    }
}

}  // namespace psmf

using namespace psmf;

struct psmf_engine {
    psmf_config cfg;
    int R = 0;
    int S = 1;
    int64_t d = 0, ntiles = 0;
    size_t esize = 8;
    void* C = nullptr;
    double* state = nullptr;
    double* partials = nullptr;
    unsigned long long* bar = nullptr;
    double* gparams = nullptr;
    long long* status = nullptr;
    int cps = 1, threads = 0, launches_last = 0, num_sms = 0;
    size_t dyn_smem = 0;
    bool cooperative = false;
    // TMA-staged kernel configuration (cps2 == 0: not available for this shape)
    int cps2 = 0, threads2 = 0, nslot = 0;
    size_t dyn_smem2 = 0;
    bool resident2 = false;
    int last_kernel = 0;
    unsigned long long* trace = nullptr;
    int trace_steps = 0;
    // NVLink mailbox (world_size > 1): [2 parities][MAX_PEERS][192] tagged 16-byte cells (psmf_filter.cuh gpu_exchange)
    void* mbox = nullptr;
    void* peer_mbox[PSMF_MAX_PEERS] = {nullptr};
    bool connected = false;
    int agreed_kernel = 0;             // world_size > 1: kernel all ranks agreed on in psmf_mailbox_connect (1 or 2)
    unsigned long long spin_ns = 10ULL * 1000000000ULL;   // bounded waits of the kernels (env PSMF_SPIN_TIMEOUT_MS)
    unsigned long long step_base = 0;
    cudaStream_t last_stream = nullptr;
    std::string err;
};

static std::string g_create_error;

static size_t mbox_bytes() { return (size_t)2 * PSMF_MAX_PEERS * MBOX_SLOT * 16; }   // [2 parities][peers] slots of MBOX_SLOT cells

static int fail(psmf_engine* h, int code, const std::string& msg) {
    if (h)
        h->err = msg;
    else
        g_create_error = msg;
    return code;
}
#define CK(h, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return fail(h, PSMF_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                \
    } while (0)

extern "C" int psmf_version(void) { return 200; }

extern "C" const char* psmf_last_error(psmf_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static void free_engine(psmf_engine* e) {
    if (!e) return;
    cudaFree(e->C);
    cudaFree(e->state);
    cudaFree(e->partials);
    cudaFree(e->bar);
    cudaFree(e->gparams);
    cudaFree(e->status);
    for (int i = 0; i < PSMF_MAX_PEERS; ++i)
        if (e->peer_mbox[i] && i != e->cfg.rank) cudaIpcCloseMemHandle(e->peer_mbox[i]);
    cudaFree(e->mbox);
    delete e;
}

extern "C" int psmf_create(psmf_handle* out, const psmf_config* cfg) {
    if (!out || !cfg) return fail(nullptr, PSMF_E_INVALID, "null argument");
    *out = nullptr;
    if (cfg->r < 1 || cfg->r > PSMF_MAX_RANK) return fail(nullptr, PSMF_E_INVALID, "rank r must be in 1..16");
    if (cfg->d < 1) return fail(nullptr, PSMF_E_INVALID, "d must be >= 1");
    if (cfg->n_series < 1) return fail(nullptr, PSMF_E_INVALID, "n_series must be >= 1");
    if (cfg->dtype != PSMF_F64 && cfg->dtype != PSMF_F32) return fail(nullptr, PSMF_E_INVALID, "dtype must be PSMF_F64 or PSMF_F32");
    if (cfg->dynamics != PSMF_DYN_IDENTITY && cfg->dynamics != PSMF_DYN_COS && cfg->dynamics != PSMF_DYN_EXTERNAL)
        return fail(nullptr, PSMF_E_INVALID, "unknown dynamics id");
    if (cfg->world_size < 1 || cfg->world_size > PSMF_MAX_PEERS || cfg->rank < 0 || cfg->rank >= cfg->world_size)
        return fail(nullptr, PSMF_E_INVALID, "bad world_size / rank");
    if (cfg->world_size > 1 && cfg->n_series > 1)
        return fail(nullptr, PSMF_E_INVALID, "row sharding (world_size > 1) and batching (n_series > 1) are exclusive");
    const int64_t dg = cfg->d_global > 0 ? cfg->d_global : cfg->d;
    if (dg < cfg->d) return fail(nullptr, PSMF_E_INVALID, "d_global < d");

    cudaError_t ce = cudaSetDevice(cfg->device);
    if (ce != cudaSuccess) return fail(nullptr, PSMF_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
    psmf_engine* e = new (std::nothrow) psmf_engine();
    if (!e) return fail(nullptr, PSMF_E_NOMEM, "host allocation failed");
    e->cfg = *cfg;
    e->cfg.d_global = dg;
    e->R = cfg->r;
    e->S = cfg->n_series;
    e->d = cfg->d;
    e->ntiles = (cfg->d + TILE - 1) / TILE;
    e->esize = cfg->dtype == PSMF_F64 ? 8 : 4;
    int sms = 0, maxsmem = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    cudaDeviceGetAttribute(&maxsmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    e->num_sms = sms;
    if (const char* ev = getenv("PSMF_SPIN_TIMEOUT_MS")) {              // how long a kernel waits for a silent peer GPU / CTA
        const long long ms = atoll(ev);
        if (ms >= 1) e->spin_ns = (unsigned long long)ms * 1000000ULL;
    }

    // ---- direct-load kernel: V1_WARPS warps per CTA, one staging tile per warp + the residual buffer ----
    int cps;
    if (e->S > 1) {
        cps = 1;
    } else if (cfg->ctas > 0) {
        cps = cfg->ctas;
    } else {
        int64_t want = e->ntiles / (2 * V1_WARPS);
        cps = (int)(want < 1 ? 1 : (want > sms ? sms : want));
    }
    if ((int64_t)cps > e->ntiles) cps = (int)e->ntiles;
    const size_t stage_bytes = (size_t)V1_WARPS * e->R * TILE * sizeof(double);
    LaunchShape shp;
    for (;;) {
        const int64_t tiles_per = (e->ntiles + cps - 1) / cps + 1;
        e->dyn_smem = stage_bytes + (size_t)tiles_per * TILE * sizeof(double);
        ce = SHAPE[e->R](cfg->dtype, e->dyn_smem, &shp);
        if (ce == cudaSuccess && shp.max_ctas_per_sm >= 1) break;
        cudaGetLastError();
        // residual buffer does not fit: use more CTAs if allowed
        if (e->S == 1 && cfg->ctas <= 0 && cps < sms) {
            cps = cps * 2 > sms ? sms : cps * 2;
            continue;
        }
        free_engine(e);
        return fail(nullptr, PSMF_E_NOMEM, "d too large: per-CTA residual buffer exceeds shared memory");
    }
    if (e->S == 1 && cps > 1 && cps > sms * shp.max_ctas_per_sm) cps = sms * shp.max_ctas_per_sm;
    e->cps = cps;
    e->threads = shp.threads;
    e->cooperative = cps > 1;

    // ---- TMA-staged kernel: slots of V2_TS tiles + y/m slices, residual buffer behind them ----
    if (cfg->kernel != 1 && e->d % 16 == 0 && e->S == 1 && sms >= 2) {
        const int TS = V2_TS;
        auto r128 = [](size_t x) { return (x + 127) / 128 * 128; };
        const size_t slot = r128((size_t)TS * e->R * TILE * e->esize);     // psmf_stream.cuh SlotLayout: one chunk of C
        // data CTAs (one per SM) + one control CTA, all co-resident (cooperative launch)
        int cps2;
        if (cfg->ctas > 0) cps2 = cfg->ctas;
        else {
            int64_t want = e->ntiles / V2_CWARPS;                       // about one tile per pass warp and step
            cps2 = (int)(want < 1 ? 1 : want);
        }
        if ((int64_t)cps2 > e->ntiles) cps2 = (int)e->ntiles;
        if (cps2 > sms - 1) cps2 = sms - 1;
        if (cps2 > 256) cps2 = 256;                                    // control CTA sums <= 256 CTA partials per entry
        LaunchShape shp2;
        if (SHAPE_S[e->R](cfg->dtype, 0, &shp2) == cudaSuccess) {
            const int64_t tiles_max = (e->ntiles + cps2 - 1) / cps2;
            const int64_t nchunks_max = (tiles_max + TS - 1) / TS;
            const size_t ebuf = (size_t)tiles_max * TILE * sizeof(double);
            const int64_t avail = (int64_t)maxsmem - shp2.static_smem - (int64_t)ebuf - 256;
            int64_t nslot = avail > 0 ? avail / (int64_t)slot : 0;
            if (nslot > nchunks_max) nslot = nchunks_max;
            if (nslot > 64) nslot = 64;
#ifdef PSMF_DEBUG
            if (const char* ev = getenv("PSMF_NSLOT_MAX")) {            // debug knob (debug builds only)
                const int64_t cap = atoll(ev);
                if (cap >= 2 && nslot > cap) nslot = cap;
            }
#endif
            // a ring needs depth: with fewer than 5 slots the producer cannot keep loads, stores and the pass warps
            // apart (very large shards, where the residual buffer eats the shared memory) -> direct-load kernel
            const bool ring_ok = nchunks_max <= nslot || nslot >= 5;
            if (ring_ok && nslot >= (nchunks_max >= 2 ? 2 : 1)) {
                e->cps2 = cps2;
                e->nslot = (int)nslot;
                e->dyn_smem2 = (size_t)nslot * slot + ebuf;
                e->threads2 = shp2.threads;
                e->resident2 = nchunks_max <= nslot;
                if (SHAPE_S[e->R](cfg->dtype, e->dyn_smem2, &shp2) != cudaSuccess || shp2.max_ctas_per_sm < 1) e->cps2 = 0;
            }
        }
        cudaGetLastError();
    }
    if (cfg->kernel == 2 && e->cps2 == 0) {
        free_engine(e);
        return fail(nullptr, PSMF_E_INVALID, "TMA-staged kernel not available for this shape (needs d % 16 == 0, one series, and room for a ring of 5 chunk slots)");
    }

    const size_t cbytes = (size_t)e->S * e->ntiles * TILE * e->R * e->esize;
    const int nsp = (ngram(e->R) + 2 * e->R + 5 + 7) / 8 * 8;      // >= nstat_pad(R): pipelined statistics (nstat2_pad)
#define CKC(call)                                                                                            \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            free_engine(e);                                                                                  \
            return fail(nullptr, e__ == cudaErrorMemoryAllocation ? PSMF_E_NOMEM : PSMF_E_CUDA,              \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                                \
        }                                                                                                    \
    } while (0)
    CKC(cudaMalloc(&e->C, cbytes));
    CKC(cudaMemset(e->C, 0, cbytes));
    CKC(cudaMalloc(&e->state, (size_t)e->S * st_size(e->R) * sizeof(double)));
    CKC(cudaMemset(e->state, 0, (size_t)e->S * st_size(e->R) * sizeof(double)));
    {
        const int cmax = e->cps > e->cps2 ? e->cps : e->cps2;
        const size_t pstr = (size_t)((cmax + 7) & ~7) + 1;            // transposed partials + totals (grid_reduce)
        // direct kernel: doubles; pipelined kernel: [2][nstat2_pad][pstr] tagged 16-byte cells
        const size_t pbytes = (size_t)2 * nsp * (pstr > 17 ? pstr : 17) * 16;
        CKC(cudaMalloc(&e->partials, pbytes));
        CKC(cudaMemset(e->partials, 0, pbytes));
    }
    CKC(cudaMalloc(&e->bar, 8 * sizeof(unsigned long long)));
    CKC(cudaMalloc(&e->gparams, GPARAMS_BYTES));                        // 16-byte cells (psmf_stream.cuh): parameter sets [2][2R], totals [2][nstat2_pad]
    CKC(cudaMalloc(&e->status, sizeof(long long)));
    CKC(cudaMemset(e->status, 0xFF, sizeof(long long)));
    if (cfg->world_size > 1) {
        CKC(cudaMalloc(&e->mbox, mbox_bytes()));
        CKC(cudaMemset(e->mbox, 0, mbox_bytes()));
    }
#undef CKC
    *out = e;
    return PSMF_OK;
}

extern "C" int psmf_destroy(psmf_handle h) {
    if (!h) return PSMF_E_INVALID;
    cudaSetDevice(h->cfg.device);
    free_engine(h);
    return PSMF_OK;
}

static int copy_small(psmf_engine* h, bool set, double* user, int off, int n, cudaStream_t st) {
    if (!user) return PSMF_OK;
    const size_t pitch = (size_t)st_size(h->R) * sizeof(double);
    double* eng = h->state + off;
    if (set)
        CK(h, cudaMemcpy2DAsync(eng, pitch, user, n * sizeof(double), n * sizeof(double), h->S, cudaMemcpyDeviceToDevice, st));
    else
        CK(h, cudaMemcpy2DAsync(user, n * sizeof(double), eng, pitch, n * sizeof(double), h->S, cudaMemcpyDeviceToDevice, st));
    return PSMF_OK;
}

static int state_io(psmf_engine* h, bool set, void* C, double* V, double* P, double* x, double* Q, double* rho,
                    double* lambda, double* theta, cudaStream_t st) {
    if (!h) return PSMF_E_INVALID;
    CK(h, cudaSetDevice(h->cfg.device));
    const int R = h->R;
    if (C) {
        const int64_t total = (int64_t)h->S * h->ntiles * TILE * R;
        const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
        if (h->cfg.dtype == PSMF_F64) {
            if (set) pack_C<double><<<blocks, 256, 0, st>>>((const double*)C, (double*)h->C, h->d, R, h->ntiles, total);
            else unpack_C<double><<<blocks, 256, 0, st>>>((const double*)h->C, (double*)C, h->d, R, h->ntiles, total);
        } else {
            if (set) pack_C<float><<<blocks, 256, 0, st>>>((const float*)C, (float*)h->C, h->d, R, h->ntiles, total);
            else unpack_C<float><<<blocks, 256, 0, st>>>((const float*)h->C, (float*)C, h->d, R, h->ntiles, total);
        }
        CK(h, cudaGetLastError());
    }
    int rc;
    if ((rc = copy_small(h, set, x, st_x(R), R, st))) return rc;
    if ((rc = copy_small(h, set, P, st_P(R), R * R, st))) return rc;
    if ((rc = copy_small(h, set, V, st_V(R), R * R, st))) return rc;
    if ((rc = copy_small(h, set, Q, st_Q(R), R * R, st))) return rc;
    if ((rc = copy_small(h, set, theta, st_theta(R), R, st))) return rc;
    if ((rc = copy_small(h, set, rho, st_rho(R), 1, st))) return rc;
    if ((rc = copy_small(h, set, lambda, st_lam(R), 1, st))) return rc;
    return PSMF_OK;
}

extern "C" int psmf_set_state(psmf_handle h, const void* C, const double* V, const double* P, const double* x,
                              const double* Q, const double* rho, const double* lambda, const double* theta, void* stream) {
    return state_io(h, true, const_cast<void*>(C), const_cast<double*>(V), const_cast<double*>(P), const_cast<double*>(x),
                    const_cast<double*>(Q), const_cast<double*>(rho), const_cast<double*>(lambda),
                    const_cast<double*>(theta), (cudaStream_t)stream);
}

extern "C" int psmf_get_state(psmf_handle h, void* C, double* V, double* P, double* x, double* Q, double* rho,
                              double* lambda, double* theta, void* stream) {
    return state_io(h, false, C, V, P, x, Q, rho, lambda, theta, (cudaStream_t)stream);
}

extern "C" int psmf_run(psmf_handle h, const psmf_io* io, int64_t n_steps, int64_t k0, void* stream) {
    if (!h || !io) return PSMF_E_INVALID;
    if (n_steps < 1) return fail(h, PSMF_E_INVALID, "n_steps must be >= 1");
    if (!io->Y) return fail(h, PSMF_E_INVALID, "Y is NULL");
    if (io->ldy < h->d) return fail(h, PSMF_E_INVALID, "ldy < d");
    if (io->M && io->ldm < h->d) return fail(h, PSMF_E_INVALID, "ldm < d");
    if (io->Yrec_out && io->ldrec < h->d) return fail(h, PSMF_E_INVALID, "ldrec < d");
    if (h->cfg.dynamics == PSMF_DYN_EXTERNAL) {
        if (n_steps != 1) return fail(h, PSMF_E_INVALID, "PSMF_DYN_EXTERNAL runs one step per call");
        if (!io->xbar_ext || (!io->F_ext && !(h->cfg.flags & PSMF_SIMPLIFIED)))
            return fail(h, PSMF_E_INVALID, "PSMF_DYN_EXTERNAL needs xbar_ext and F_ext");
    }
    cudaStream_t st = (cudaStream_t)stream;
    CK(h, cudaSetDevice(h->cfg.device));
    KParams p;
    memset(&p, 0, sizeof(p));
    p.C = h->C;
    p.c_series_stride = h->ntiles * TILE * h->R;
    p.state = h->state;
    p.Y = io->Y; p.ldy = io->ldy; p.ysst = io->y_series_stride;
    p.M = io->M; p.ldm = io->ldm; p.msst = io->m_series_stride;
    p.X_out = io->X_out;
    p.Yrec = io->Yrec_out; p.ldrec = io->ldrec; p.recsst = io->rec_series_stride;
    p.scal_out = io->scal_out;
    p.xbar_ext = io->xbar_ext; p.F_ext = io->F_ext;
    p.grad_out = io->grad_out;
    p.partials = h->partials;
    p.bar = h->bar;
    p.status = h->status;
    p.d = h->d; p.d_global = h->cfg.d_global;
    p.n_steps = n_steps; p.k0 = k0;
    p.n_series = h->S; p.cps = h->cps;
    p.flags = h->cfg.flags; p.dynamics = h->cfg.dynamics;
#ifdef PSMF_DEBUG
    if (const char* ev = getenv("PSMF_DEBUG_FLAGS")) p.flags |= (int)strtol(ev, nullptr, 0) & (7 << 20);
#endif
    p.spin_ns = h->spin_ns;
    p.alpha = h->cfg.alpha; p.beta = h->cfg.beta;
    p.world = h->cfg.world_size; p.rank = h->cfg.rank;
    if (p.world > 1) {
        if (!h->connected) return fail(h, PSMF_E_STATE, "world_size > 1: call psmf_mailbox_connect before psmf_run");
        // the mailbox slots are nstat_pad(R) doubles apart (kernel indexing), inside a buffer sized for MAXR
        p.mbox_local = (double*)h->mbox;
        for (int i = 0; i < p.world; ++i) {
            p.mbox_peer[i] = (double*)h->peer_mbox[i];
        }
    }
    p.step_base = h->step_base;       // steps filtered by this engine so far: tags of the mailbox / partial cells
    p.trace = h->trace; p.trace_steps = h->trace_steps;
    CK(h, cudaMemsetAsync(h->bar, 0, 8 * sizeof(unsigned long long), st));
    CK(h, cudaMemsetAsync(h->gparams, 0, GPARAMS_BYTES, st));
    CK(h, cudaMemsetAsync(h->status, 0xFF, sizeof(long long), st));
    // the TMA-staged kernel needs 16-byte aligned rows of Y / M (bulk copies)
    const size_t es = h->esize;
    bool aligned = h->cps2 > 0 && ((uintptr_t)io->Y % 16 == 0) && ((size_t)io->ldy * es % 16 == 0) &&
                   ((size_t)io->y_series_stride * es % 16 == 0);
    if (io->M) aligned = aligned && ((uintptr_t)io->M % 16 == 0) && (io->ldm % 16 == 0) && (io->m_series_stride % 16 == 0);
    bool use2 = aligned && h->cfg.dynamics != PSMF_DYN_EXTERNAL &&
                (h->cfg.kernel == 2 || (h->cfg.kernel == 0 && h->ntiles >= 2 * (V2_CWARPS + 1)));
    if (h->cfg.kernel == 2 && !use2)
        return fail(h, PSMF_E_INVALID, "kernel=2 requested but Y/M are not 16-byte aligned (or dynamics is external)");
    if (p.world > 1) {
        // row sharding: the kernel is a collective decision (psmf_mailbox_connect) -- the two kernels exchange
        // different statistics vectors, so a rank must never pick one from local facts alone
        if (h->agreed_kernel == 2 && !aligned)
            return fail(h, PSMF_E_INVALID, "the ranks agreed on the TMA-staged kernel but this rank's Y/M are not 16-byte aligned "
                                           "(pad ldy/ldm to a multiple of 16 bytes, or create every engine with kernel=1)");
        use2 = h->agreed_kernel == 2;
    }
    if (use2) {
        p.cps = h->cps2;
        p.nslot = h->nslot;
        p.trace_cta = h->cps2;
        // streaming from HBM: 12 of the 14 pass warps (3 per scheduler; a multiple of the 4 tiles of a chunk) keep up
        // with the ring and leave issue slots to the producer -- measured optimum at r = 16 under the power cap;
        // resident in shared memory: every warp helps
        p.npw = h->resident2 ? V2_CWARPS : 12;
#ifdef PSMF_DEBUG
        if (const char* ev = getenv("PSMF_NPW")) { const int v = atoi(ev); if (v >= 1 && v <= V2_CWARPS) p.npw = v; }
#endif
        p.gparams = h->gparams;
        CK(h, LAUNCH_S[h->R](p, h->cfg.dtype, h->cps2 + 1, h->dyn_smem2, st, true));
        h->last_kernel = 2;
    } else {
        CK(h, LAUNCH[h->R](p, h->cfg.dtype, h->S * h->cps, h->dyn_smem, st, h->cooperative));
        h->last_kernel = 1;
    }
    h->launches_last = 1;
    h->last_stream = st;
    h->step_base += (unsigned long long)n_steps;
    return PSMF_OK;
}

extern "C" int psmf_status(psmf_handle h, int64_t* first_bad_step) {
    if (!h) return PSMF_E_INVALID;
    CK(h, cudaSetDevice(h->cfg.device));
    CK(h, cudaStreamSynchronize(h->last_stream));
    long long v = -1;
    CK(h, cudaMemcpy(&v, h->status, sizeof(v), cudaMemcpyDeviceToHost));
    if (v >= 0 && (v & (STATUS_TIMEOUT | STATUS_MISMATCH)) != 0) {
        const long long step = v & 0xFFFFFFFFFFFFLL;
        if (first_bad_step) *first_bad_step = (int64_t)step;
        if (v & STATUS_TIMEOUT) {
            static const char* const SITE[] = {"?", "tagged cell", "NVLink mailbox (peer GPU)", "mbarrier", "grid barrier",
                                               "chunk ring", "CTA partials"};
            const int where = (int)((v >> 48) & 0xFF);
            char buf[256];
            snprintf(buf, sizeof(buf), "a kernel wait made no progress for %.1f s at step %lld (%s): a peer GPU or CTA is gone; "
                     "the results of this launch are invalid", (double)h->spin_ns * 1e-9, step, SITE[where < 7 ? where : 0]);
            return fail(h, PSMF_E_STATE, buf);
        }
        return fail(h, PSMF_E_STATE, "a peer GPU runs a different kernel / statistics layout (mailbox header mismatch at step " +
                                         std::to_string(step) + ")");
    }
    if (first_bad_step) *first_bad_step = (int64_t)v;
    return PSMF_OK;
}

extern "C" int psmf_launch_info(psmf_handle h, int32_t* ctas, int32_t* threads, int32_t* smem_bytes, int32_t* launches) {
    if (!h) return PSMF_E_INVALID;
    const bool k2 = h->last_kernel == 2;
    if (ctas) *ctas = k2 ? h->cps2 + 1 : h->cps;
    if (threads) *threads = k2 ? h->threads2 : h->threads;
    if (smem_bytes) *smem_bytes = (int32_t)(k2 ? h->dyn_smem2 : h->dyn_smem);
    if (launches) *launches = h->launches_last;
    return PSMF_OK;
}

extern "C" int psmf_launch_info2(psmf_handle h, int32_t* kernel, int32_t* nslot, int32_t* resident) {
    if (!h) return PSMF_E_INVALID;
    if (kernel) *kernel = h->last_kernel;
    if (nslot) *nslot = h->last_kernel == 2 ? h->nslot : 0;
    if (resident) *resident = h->last_kernel == 2 && h->resident2 ? 1 : 0;
    return PSMF_OK;
}

extern "C" int psmf_set_trace(psmf_handle h, uint64_t* dev_buf, int32_t steps) {
    if (!h) return PSMF_E_INVALID;
    h->trace = (unsigned long long*)dev_buf;
    h->trace_steps = dev_buf ? steps : 0;
    return PSMF_OK;
}

// Plan record that travels with the IPC handle (bytes 64..127 of the blob): every rank sees every record and
// derives the same decision from them, so the ranks can never run different kernels against each other.
struct plan_record {
    int32_t version, r, dtype, flags, dynamics, world, kernel_pref, eligible2, auto2, rank;
    int64_t d_global;
    int32_t pad[4];
};
static_assert(sizeof(plan_record) <= 64, "plan record must fit the second half of the blob");

static plan_record make_plan(const psmf_engine* h) {
    plan_record r;
    memset(&r, 0, sizeof(r));
    r.version = psmf_version();
    r.r = h->R; r.dtype = h->cfg.dtype; r.flags = h->cfg.flags; r.dynamics = h->cfg.dynamics; r.world = h->cfg.world_size;
    r.kernel_pref = h->cfg.kernel; r.rank = h->cfg.rank; r.d_global = h->cfg.d_global;
    r.eligible2 = (h->cps2 > 0 && h->cfg.dynamics != PSMF_DYN_EXTERNAL) ? 1 : 0;
    r.auto2 = (r.eligible2 && h->ntiles >= 2 * (V2_CWARPS + 1)) ? 1 : 0;
    return r;
}

extern "C" int psmf_mailbox_export(psmf_handle h, void* blob_128B) {
    if (!h || !blob_128B) return PSMF_E_INVALID;
    if (h->cfg.world_size < 2 || !h->mbox) return fail(h, PSMF_E_STATE, "engine was created with world_size == 1");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(h, cudaSetDevice(h->cfg.device));
    cudaIpcMemHandle_t hd;
    CK(h, cudaIpcGetMemHandle(&hd, h->mbox));
    memset(blob_128B, 0, PSMF_MAILBOX_BLOB_BYTES);
    memcpy(blob_128B, &hd, sizeof(hd));
    const plan_record pr = make_plan(h);
    memcpy((char*)blob_128B + 64, &pr, sizeof(pr));
    return PSMF_OK;
}

extern "C" int psmf_mailbox_connect(psmf_handle h, const void* all_blobs, int32_t n) {
    if (!h || !all_blobs) return PSMF_E_INVALID;
    if (n != h->cfg.world_size || n < 2) return fail(h, PSMF_E_INVALID, "need one blob per rank");
    CK(h, cudaSetDevice(h->cfg.device));
    const char* blobs = (const char*)all_blobs;
    // ---- collective plan: same inputs on every rank -> same decision on every rank ----
    const plan_record mine = make_plan(h);
    bool any_forced1 = false, any_forced2 = false, all_eligible2 = true, all_auto2 = true;
    for (int i = 0; i < n; ++i) {
        plan_record pr;
        memcpy(&pr, blobs + (size_t)i * PSMF_MAILBOX_BLOB_BYTES + 64, sizeof(pr));
        if (pr.version != mine.version || pr.r != mine.r || pr.dtype != mine.dtype || pr.flags != mine.flags ||
            pr.dynamics != mine.dynamics || pr.world != mine.world || pr.d_global != mine.d_global || pr.rank != i)
            return fail(h, PSMF_E_INVALID, "rank " + std::to_string(i) + " was created with a different configuration "
                                           "(version / r / dtype / flags / dynamics / world_size / d_global / rank order)");
        any_forced1 = any_forced1 || pr.kernel_pref == 1;
        any_forced2 = any_forced2 || pr.kernel_pref == 2;
        all_eligible2 = all_eligible2 && pr.eligible2 != 0;
        all_auto2 = all_auto2 && (pr.auto2 != 0 || (pr.kernel_pref == 2 && pr.eligible2 != 0));
    }
    if (any_forced1 && any_forced2) return fail(h, PSMF_E_INVALID, "ranks disagree: kernel=1 and kernel=2 were both forced");
    if (any_forced2 && !all_eligible2)
        return fail(h, PSMF_E_INVALID, "kernel=2 was forced but not every rank's shard is eligible for the TMA-staged kernel");
    h->agreed_kernel = any_forced2 ? 2 : (any_forced1 ? 1 : (all_auto2 ? 2 : 1));
    for (int i = 0; i < n; ++i) {
        if (i == h->cfg.rank) {
            h->peer_mbox[i] = h->mbox;
            continue;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, blobs + (size_t)i * PSMF_MAILBOX_BLOB_BYTES, sizeof(hd));
        void* ptr = nullptr;
        CK(h, cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
        h->peer_mbox[i] = ptr;
    }
    h->connected = true;
    return PSMF_OK;
}
