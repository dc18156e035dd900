// Where does the cold start of a one-shot process go?  Times the steps of bringing up CUDA for libbogp.so.
// nvcc -O2 -o ctx_probe.bin ctx_probe.cu -lcuda
#include <cstdio>
#include <chrono>
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    double t0 = now();
    cuInit(0);
    double t1 = now();
    int n = 0; cudaGetDeviceCount(&n);
    double t2 = now();
    cudaSetDevice(0); cudaFree(0);
    double t3 = now();
    void* p; cudaMalloc(&p, 1 << 20);
    double t4 = now();
    cudaStream_t s; cudaStreamCreate(&s);
    double t5 = now();
    void* h = argc > 1 ? dlopen(argv[1], RTLD_NOW) : nullptr;
    double t6 = now();
    int rc = -99; void* ctx = nullptr;
    if (h) { auto f = (int (*)(int, void**))dlsym(h, "bogp_create"); if (f) rc = f(0, &ctx); }
    double t7 = now();
    printf("{\"cuInit_s\": %.3f, \"device_count_s\": %.3f, \"context_s\": %.3f, \"first_malloc_s\": %.3f, \"stream_s\": %.3f, \"dlopen_libbogp_s\": %.3f, \"bogp_create_s\": %.3f, \"bogp_create_rc\": %d, \"devices\": %d}\n",
           t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, rc, n);
    return 0;
}
