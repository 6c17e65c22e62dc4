/*
 * Plain-C caller of the PSMF / rPSMF filter ABI (include/psmf_b200.h): the calls a host program in any
 * language with a C FFI makes.  Filters T steps of a synthetic d x r problem with 20 % missing entries and
 * prints the last filtered x_t.
 *
 *   gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/psmf_demo.c -o psmf_demo \
 *       -Lrpsmf_b200 -lpsmf_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/rpsmf_b200
 *   ./psmf_demo 100000 16 200
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "psmf_b200.h"

#define CK_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                   \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)
#define CK_PSMF(h, call)                                                                  \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != PSMF_OK) {                                                             \
            fprintf(stderr, "%s: error %d: %s\n", #call, rc_, psmf_last_error(h));        \
            return 3;                                                                     \
        }                                                                                 \
    } while (0)

static double urand(unsigned long long* s) {   /* xorshift64*: uniform in [0, 1) */
    *s ^= *s >> 12; *s ^= *s << 25; *s ^= *s >> 27;
    return (double)((*s * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
    const int64_t d = argc > 1 ? atoll(argv[1]) : 100000;
    const int r = argc > 2 ? atoi(argv[2]) : 16;
    const int64_t T = argc > 3 ? atoll(argv[3]) : 200;
    unsigned long long seed = 20261017ULL;

    /* host data: y_t = C_true x_t + noise, time-major, zero-filled where missing */
    double* Ct = malloc(sizeof(double) * d * r);
    double* C0 = malloc(sizeof(double) * d * r);
    double* Y = malloc(sizeof(double) * T * d);
    uint8_t* M = malloc((size_t)T * d);
    double x[PSMF_MAX_RANK], x0[PSMF_MAX_RANK];
    double V[PSMF_MAX_RANK * PSMF_MAX_RANK] = {0}, P[PSMF_MAX_RANK * PSMF_MAX_RANK] = {0}, Q[PSMF_MAX_RANK * PSMF_MAX_RANK] = {0};
    if (!Ct || !C0 || !Y || !M || r < 1 || r > PSMF_MAX_RANK) return 1;
    for (int64_t i = 0; i < d * r; ++i) { Ct[i] = 2.0 * urand(&seed) - 1.0; C0[i] = urand(&seed); }
    for (int j = 0; j < r; ++j) { x[j] = urand(&seed); x0[j] = urand(&seed); V[j * r + j] = 2.0; P[j * r + j] = 1.0; Q[j * r + j] = 0.1; }
    for (int64_t t = 0; t < T; ++t) {
        for (int j = 0; j < r; ++j) x[j] += 0.1 * (urand(&seed) - 0.5);
        for (int64_t i = 0; i < d; ++i) {
            double y = 0.3 * (urand(&seed) - 0.5);
            for (int j = 0; j < r; ++j) y += Ct[i * r + j] * x[j];
            M[t * d + i] = urand(&seed) >= 0.2;
            Y[t * d + i] = M[t * d + i] ? y : 0.0;
        }
    }

    /* device buffers are caller-owned */
    double *dC, *dY, *dX, *dV, *dP, *dQ, *dx, *drho, *dlam;
    uint8_t* dM;
    const double rho = 10.0, lam = 1.8;
    CK_CUDA(cudaMalloc((void**)&dC, sizeof(double) * d * r));
    CK_CUDA(cudaMalloc((void**)&dY, sizeof(double) * T * d));
    CK_CUDA(cudaMalloc((void**)&dM, (size_t)T * d));
    CK_CUDA(cudaMalloc((void**)&dX, sizeof(double) * T * r));
    CK_CUDA(cudaMalloc((void**)&dV, sizeof(double) * r * r));
    CK_CUDA(cudaMalloc((void**)&dP, sizeof(double) * r * r));
    CK_CUDA(cudaMalloc((void**)&dQ, sizeof(double) * r * r));
    CK_CUDA(cudaMalloc((void**)&dx, sizeof(double) * r));
    CK_CUDA(cudaMalloc((void**)&drho, sizeof(double)));
    CK_CUDA(cudaMalloc((void**)&dlam, sizeof(double)));
    CK_CUDA(cudaMemcpy(dC, C0, sizeof(double) * d * r, cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(dY, Y, sizeof(double) * T * d, cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(dM, M, (size_t)T * d, cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(dV, V, sizeof(double) * r * r, cudaMemcpyHostToDevice));      /* r x r row-major */
    CK_CUDA(cudaMemcpy(dP, P, sizeof(double) * r * r, cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(dQ, Q, sizeof(double) * r * r, cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(dx, x0, sizeof(double) * r, cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(drho, &rho, sizeof(double), cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(dlam, &lam, sizeof(double), cudaMemcpyHostToDevice));

    psmf_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.d = d; cfg.d_global = d; cfg.r = r; cfg.n_series = 1; cfg.dtype = PSMF_F64;
    cfg.flags = PSMF_ROBUST | PSMF_CUPDATE_VT;                 /* the masked rPSMF step of ExperimentImpute/rPSMF.py */
    cfg.dynamics = PSMF_DYN_IDENTITY; cfg.device = 0; cfg.world_size = 1; cfg.alpha = 1.0; cfg.beta = 1.0;
    psmf_handle h = NULL;
    CK_PSMF(NULL, psmf_create(&h, &cfg));
    CK_PSMF(h, psmf_set_state(h, dC, dV, dP, dx, dQ, drho, dlam, NULL, NULL));

    psmf_io io;
    memset(&io, 0, sizeof io);
    io.Y = dY; io.ldy = d; io.M = dM; io.ldm = d; io.X_out = dX;
    CK_PSMF(h, psmf_run(h, &io, T, 1, NULL));                  /* one launch filters all T steps */
    int64_t bad = 0;
    CK_PSMF(h, psmf_status(h, &bad));

    int32_t ctas, threads, smem, launches, kernel, nslot, resident;
    CK_PSMF(h, psmf_launch_info(h, &ctas, &threads, &smem, &launches));
    CK_PSMF(h, psmf_launch_info2(h, &kernel, &nslot, &resident));
    double xT[PSMF_MAX_RANK];
    CK_CUDA(cudaMemcpy(xT, dX + (T - 1) * r, sizeof(double) * r, cudaMemcpyDeviceToHost));
    printf("psmf_b200 v%d: d=%lld r=%d T=%lld  kernel=%s  %d CTAs x %d threads  first_bad_step=%lld\nx_T =",
           psmf_version(), (long long)d, r, (long long)T, kernel == PSMF_KERNEL_STREAM ? "tma" : (kernel == PSMF_KERNEL_BATCH ? "batch" : "direct"), ctas, threads, (long long)bad);
    for (int j = 0; j < r; ++j) printf(" %.6f", xT[j]);
    printf("\n");
    CK_PSMF(h, psmf_destroy(h));
    cudaFree(dC); cudaFree(dY); cudaFree(dM); cudaFree(dX); cudaFree(dV); cudaFree(dP); cudaFree(dQ); cudaFree(dx); cudaFree(drho); cudaFree(dlam);
    free(Ct); free(C0); free(Y); free(M);
    return bad < 0 ? 0 : 4;
}
