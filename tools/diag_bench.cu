// Stand-alone timing of the Cholesky diagonal-block kernel variants.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../bayesian_optimisation_b200/csrc -I. -o diag_bench.bin diag_bench.cu
#include "chol_diag.cuh"
#include "diag_variants.cuh"
#include <vector>
#include <cmath>
namespace bogp { void set_error(const char*, ...) {} }
using namespace bogp;
int main() {
    const int n = 64;
    std::vector<double> h(n * n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) h[i * n + j] = exp(-0.5 * (i - j) * (i - j) / 9.0) + (i == j ? 1e-2 : 0);
    double *a, *a0, *w, *ld; int* info;
    cudaMalloc(&a, n * n * 8); cudaMalloc(&a0, n * n * 8); cudaMalloc(&w, n * n * 8); cudaMalloc(&ld, 8); cudaMalloc(&info, 4);
    cudaMemcpy(a0, h.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaMemset(ld, 0, 8); cudaMemset(info, 0, 4);
    DiagArgs g{a, n, 0, w, n, 0, ld, info, 0};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 3; variant++) {
        float best = 1e9;
        for (int rep = 0; rep < 20; rep++) {
            cudaMemcpy(a, a0, n * n * 8, cudaMemcpyDeviceToDevice);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            if (variant == 0) chol_diag_kernel_v1<<<1, 256>>>(g); else if (variant == 1) chol_diag_kernel_v2<<<1, 256>>>(g); else chol_diag_kernel<<<1, 256>>>(g);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        std::vector<double> l(n * n), x(n * n);
        cudaMemcpy(l.data(), a, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(x.data(), w, n * n * 8, cudaMemcpyDeviceToHost);
        double err = 0, errx = 0;
        for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
            double s = 0, t = 0; for (int k = 0; k <= j; k++) s += l[i * n + k] * l[j * n + k];
            for (int k = j; k <= i; k++) t += l[i * n + k] * x[k * n + j];
            err = fmax(err, fabs(s - h[i * n + j])); errx = fmax(errx, fabs(t - (i == j)));
        }
#ifdef BOGP_DIAG_TRACE
        if (variant == 2) { long long tr[64]; cudaMemcpyFromSymbol(tr, g_diag_trace, sizeof(tr)); printf("groups (panel+barrier | update) cycles:"); for (int jb = 0; jb < 16; jb++) printf(" %lld|%lld", tr[3*jb+1]-tr[3*jb], tr[3*jb+2]-tr[3*jb+1]); printf("\n total loop %lld, epilogue %lld cycles\n", tr[48]-tr[0], tr[49]-tr[48]);
          long long t2[16][8]; cudaMemcpyFromSymbol(t2, g_diag_trace2, sizeof(t2)); printf("per group: panel | publish | barrier->next owner | next owner trailing update | (owner inverse update)\n"); for (int jb = 1; jb < 15; jb++) printf("  g%2d: %5lld %5lld %5lld %5lld (%5lld)   start-to-start %lld\n", jb, t2[jb][1]-t2[jb][0], t2[jb][2]-t2[jb][1], t2[jb][3]-t2[jb][2], t2[jb][4]-t2[jb][3], t2[jb][5]-t2[jb][2], t2[jb+1][0]-t2[jb][0]); }
#endif
        printf("variant %d: %.2f us   |LL^T-A| %.2e  |L X - I| %.2e  (%s)\n", variant, best * 1e3, err, errx, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
