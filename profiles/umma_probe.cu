// Standalone probe of tcgen05.mma kind::tf32 operand layouts (sm_100a): one CTA, operands written to shared memory by
// ordinary stores in a chosen canonical layout, one accumulator tile read back with tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe profiles/umma_probe.cu && ./umma_probe
// Prints, per variant, the number of wrong accumulator elements against an integer-exact host product.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int M = 128, KT = 32;   // K total per stage = 4 MMAs of K = 8

struct Variant {
    int a_mn;            // 0: A K-major SWIZZLE_128B; 1: A MN-major SWIZZLE_128B (16-byte base); 2: A MN-major SWIZZLE_128B_BASE32B
    int n;               // UMMA N
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_kstep, b_kstep;   // bytes added to the start address per K = 8 step
    int nk;              // number of MMAs
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;       // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* A /*[M][KT] row-major logical*/, const float* B /*[N][KT]*/, float* D /*[M][N]*/, float* D2, Variant v, int* flag) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sa = sm;                    // 16 KB
    uint8_t* sb = sm + 16384;            // N * 128 bytes
    const int tid = threadIdx.x, warp = tid >> 5;

    // operands in their canonical SWIZZLE_128B layouts
    for (int i = tid; i < M * KT; i += 128) {
        const int m = i / KT, k = i % KT;
        uint32_t off;
        if (v.a_mn == 2) off = (m / 32) * 4096 + k * 128 + ((((m % 32) / 8) ^ (k % 4)) * 32) + (m % 8) * 4;   // Swizzle<2,5,2>
        else if (v.a_mn) off = (m / 32) * 4096 + k * 128 + ((((m % 32) / 4) ^ (k % 8)) * 16) + (m % 4) * 4;
        else        off = (m / 8) * 1024 + (m % 8) * 128 + (((k / 4) ^ (m % 8)) * 16) + (k % 4) * 4;
        *reinterpret_cast<float*>(sa + off) = A[i];
    }
    for (int i = tid; i < v.n * KT; i += 128) {
        const int n = i / KT, k = i % KT;
        const uint32_t off = (n / 8) * 1024 + (n % 8) * 128 + (((k / 4) ^ (n % 8)) * 16) + (k % 4) * 4;
        *reinterpret_cast<float*>(sb + off) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) flag[0] = (int)tmem;

    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(v.a_mn != 0) << 15) | ((uint32_t)(v.n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int k = 0; k < v.nk; ++k) {
            const uint64_t ad = smem_desc(base + k * v.a_kstep, v.a_lbo, v.a_sbo, v.a_mn == 2 ? 1u : 2u);
            const uint64_t bd = smem_desc(base + 16384 + k * v.b_kstep, v.b_lbo, v.b_sbo);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(k > 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the accumulator
    uint32_t ok = 0;
    for (int spins = 0; !ok && spins < (1 << 22); ++spins)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    if (!ok && tid == 0) flag[1] = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tid;                   // warp w reads TMEM lanes 32w..32w+31
    for (int c = 0; c < v.n; c += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[row * v.n + c + j] = __uint_as_float(r[j]);
    }
    // the same accumulator through 16-column loads that start at column 5 (unaligned): D2[row][c] = D[row][c + 5]
    if (D2) {
        for (int c = 0; c + 5 + 16 <= v.n; c += 16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c + 5)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) D2[row * v.n + c + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
    const int NMAX = 256;
    float *hA = (float*)malloc(M * KT * 4), *hB = (float*)malloc(NMAX * KT * 4), *hD = (float*)malloc(M * NMAX * 4);
    srand(7);
    for (int i = 0; i < M * KT; ++i) hA[i] = (float)(rand() % 15 - 7);
    for (int i = 0; i < NMAX * KT; ++i) hB[i] = (float)(rand() % 9 - 4);
    float *dA, *dB, *dD, *dD2;
    float* hD2 = (float*)malloc(M * NMAX * 4);
    int* dflag;
    CHECK(cudaMalloc(&dA, M * KT * 4)); CHECK(cudaMalloc(&dB, NMAX * KT * 4)); CHECK(cudaMalloc(&dD, M * NMAX * 4)); CHECK(cudaMalloc(&dD2, M * NMAX * 4));
    CHECK(cudaMalloc(&dflag, 8));
    CHECK(cudaMemcpy(dA, hA, M * KT * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, hB, NMAX * KT * 4, cudaMemcpyHostToDevice));
    const size_t smem = 16384 + NMAX * 128 + 1024;
    CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    struct { const char* name; Variant v; } vs[] = {
        {"A K-major  B K-major N=64  nk=1", {0, 64, 16, 1024, 16, 1024, 32, 32, 1}},
        {"A K-major  B K-major N=64  nk=4", {0, 64, 16, 1024, 16, 1024, 32, 32, 4}},
        {"A K-major  B K-major N=256 nk=4", {0, 256, 16, 1024, 16, 1024, 32, 32, 4}},
        {"A MN-major LBO=4096 SBO=1024 kstep=1024 N=64  nk=1", {1, 64, 4096, 1024, 16, 1024, 1024, 32, 1}},
        {"A MN-major LBO=4096 SBO=1024 kstep=1024 N=64  nk=4", {1, 64, 4096, 1024, 16, 1024, 1024, 32, 4}},
        {"A MN-major LBO=4096 SBO=1024 kstep=1024 N=256 nk=4", {1, 256, 4096, 1024, 16, 1024, 1024, 32, 4}},
        {"A MN-major LBO=1024 SBO=4096 (swapped)  N=64  nk=4", {1, 64, 1024, 4096, 16, 1024, 1024, 32, 4}},
        {"A MN-major BASE32B LBO=4096 SBO=512 kstep=1024 N=64  nk=1", {2, 64, 4096, 512, 16, 1024, 1024, 32, 1}},
        {"A MN-major BASE32B LBO=4096 SBO=512 kstep=1024 N=64  nk=4", {2, 64, 4096, 512, 16, 1024, 1024, 32, 4}},
        {"A MN-major BASE32B LBO=4096 SBO=512 kstep=1024 N=256 nk=4", {2, 256, 4096, 512, 16, 1024, 1024, 32, 4}},
        {"A MN-major BASE32B LBO=512 SBO=4096 (swapped)  N=64  nk=4", {2, 64, 512, 4096, 16, 1024, 1024, 32, 4}},
    };
    for (auto& t : vs) {
        const Variant& v = t.v;
        CHECK(cudaMemset(dD, 0xff, M * NMAX * 4));
        CHECK(cudaMemset(dflag, 0, 8));
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dD2, v, dflag);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-60s  CUDA error: %s\n", t.name, cudaGetErrorString(e)); return 1; }
        int hflag[2];
        CHECK(cudaMemcpy(hD, dD, M * v.n * 4, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(hflag, dflag, 8, cudaMemcpyDeviceToHost));
        int bad = 0, zeros = 0, first = -1;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < v.n; ++n) {
                double ref = 0;
                for (int k = 0; k < 8 * v.nk; ++k) ref += (double)hA[m * KT + k] * hB[n * KT + k];
                const float got = hD[m * v.n + n];
                if (got != (float)ref) { if (first < 0) first = m * v.n + n; ++bad; }
                if (got == 0.0f) ++zeros;
            }
        printf("%-60s  wrong %6d / %6d   zeros %6d  tmem_base 0x%x  timeout %d", t.name, bad, M * v.n, zeros, hflag[0], hflag[1]);
        if (first >= 0) printf("   first (m=%d,n=%d) got %g", first / v.n, first % v.n, hD[first]);
        CHECK(cudaMemcpy(hD2, dD2, M * v.n * 4, cudaMemcpyDeviceToHost));
        int bad2 = 0;
        for (int m = 0; m < M; ++m)
            for (int c = 0; c + 5 + 16 <= v.n; c += 16)
                for (int j = 0; j < 16; ++j) bad2 += hD2[m * v.n + c + j] != hD[m * v.n + c + j + 5];
        printf("   x16 loads from column 5: %d wrong", bad2);
        printf("\n");
    }
    return 0;
}
