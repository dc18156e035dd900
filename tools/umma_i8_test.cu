// Stand-alone check of the tcgen05 kind::i8 building blocks used by the Ozaki acquisition path:
// shared-memory descriptors (K-major, no swizzle, [kchunk][row][16 B] layout), TMEM allocation,
// tcgen05.mma, tcgen05.commit -> mbarrier, tcgen05.ld.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_test.bin umma_i8_test.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // layout_type = 0 (no swizzle), base_offset = 0
}
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) /* D = s32 */ | (1u << 7) /* A = s8 */ | (1u << 10) /* B = s8 */ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int M = 128, NMAX = 256, K = 64;   // two k-steps of 32

__global__ void __launch_bounds__(128) test_kernel(const int8_t* A, const int8_t* B, int32_t* D, int N) {
    __shared__ __align__(128) int8_t sA[(K / 16) * M * 16];
    __shared__ __align__(128) int8_t sB[(K / 16) * NMAX * 16];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // [kchunk][row][16 B] layout
    for (int i = tid; i < M * K; i += 128) { int r = i / K, k = i % K; sA[(k / 16) * (M * 16) + r * 16 + k % 16] = A[i]; }
    for (int i = tid; i < N * K; i += 128) { int r = i / K, k = i % K; sB[(k / 16) * (N * 16) + r * 16 + k % 16] = B[i]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_i8(M, N);
        for (int ks = 0; ks < K / 32; ks++) {
            uint64_t da = make_desc(smem_u32(sA) + ks * 2 * (M * 16), M * 16, 128);
            uint64_t db = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
            umma_i8(tb + 64, da, db, idesc, ks > 0);                 // D at column offset 64
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + 64 + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) D[(warp * 32 + lane) * N + c0 + j] = (int32_t)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}

int main() {
    for (int N : {64, 192, 256}) {
        std::vector<int8_t> hA(M * K), hB(N * K);
        srand(N);
        for (auto& v : hA) v = (int8_t)(rand() % 129 - 64);
        for (auto& v : hB) v = (int8_t)(rand() % 129 - 64);
        int8_t *dA, *dB; int32_t* dD;
        cudaMalloc(&dA, M * K); cudaMalloc(&dB, N * K); cudaMalloc(&dD, M * N * 4);
        cudaMemcpy(dA, hA.data(), M * K, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, M * N * 4);
        test_kernel<<<1, 128>>>(dA, dB, dD, N);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<int32_t> hD(M * N);
        cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
            int32_t s = 0; for (int k = 0; k < K; k++) s += (int32_t)hA[m * K + k] * hB[n * K + k];
            if (s != hD[m * N + n]) { if (bad < 5) printf("  mismatch (%d,%d): got %d want %d\n", m, n, hD[m * N + n], s); bad++; }
        }
        printf("N=%d: %s, %ld mismatches of %d\n", N, cudaGetErrorString(e), bad, M * N);
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
