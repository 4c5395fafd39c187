// umma_ts_probe.cu -- development probe: tcgen05.mma.kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st
// from registers) and B in shared memory (K-major, SWIZZLE_128B).  Checks D = A B^T (M = 128, N = 128, K = 32) against the
// host and measures the issue rate of the long-tap kernel's 24-MMA chunk in this mode.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_ts_probe umma_ts_probe.cu && ./umma_ts_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
    return p != 0;
}

constexpr int M = 128, N = 128, K = 32;

// a: [M][K], b: [N][K] row-major in global memory; d: [M][N]
template <int NN>
__global__ void __launch_bounds__(128, 1) probe(const float* a, const float* b, float* d, int rounds, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];      // B: [256 rows x 128 B] swizzled
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
        const int n = i / 32, k = i % 32;
        const float v = n < N ? b[n * K + k] : 0.f;
        *reinterpret_cast<float*>(smem + (n / 8) * 1024 + (n % 8) * 128 + (((k / 4) ^ (n % 8)) * 16) + (k % 4) * 4) = v;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_slot;
    const uint32_t a_col = 256;                           // A tiles live in columns [256, 512): 8 tiles of 32 columns
    // thread = row m (TMEM lane), 32 K values in 32 columns, written 8 columns at a time
    const int m = warp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int t = 0; t < 8; ++t)
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(t == 0 ? a[m * K + k0 + j] : 0.f);
            tmem_st8(lane_base + a_col + t * 32 + k0, v);
        }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        for (int k = 0; k < K / 8; ++k) umma_ts(tmem, tmem + a_col + k * 8, umma_desc(s_u32(smem) + k * 32), idesc, k ? 1u : 0u);
        umma_commit(s_u32(&bars[0]));
        bar_wait(s_u32(&bars[0]), 0);
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (d)
        for (int n0 = 0; n0 < NN; n0 += 8) {
            uint32_t v[8];
            tmem_ld8(lane_base + n0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) d[m * N + n0 + j] = __uint_as_float(v[j]);
        }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    // rate: rounds of 2 planes x 3 terms x 4 k-steps, A tiles rotating over the 8 TMEM tiles, 2 rounds in flight
    if (__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) == 0 && rounds > 0) {     // warp-uniform loop, one elected lane issues
        const bool leader = elect_one();
        asm volatile("tcgen05.fence::after_thread_sync;");
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            const int s = r & 1;
            if (r >= 2) bar_wait(s_u32(&bars[2 + s]), ((r >> 1) - 1) & 1);
            if (!leader) continue;
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int term = 0; term < 3; ++term)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_ts(tmem + p * NN, tmem + a_col + (s * 4 + 2 * p + (term == 1 ? 1 : 0)) * 32 + k * 8,
                                umma_desc(s_u32(smem) + ((r % 4) + (term == 2 ? 8 : 0)) * 1024 + k * 32), idesc, 1u);
            umma_commit(s_u32(&bars[2 + s]));
        }
        const int last = rounds - 1;
        bar_wait(s_u32(&bars[2 + (last & 1)]), (last >> 1) & 1);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && leader) *cycles = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static float tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

template <int NN>
void rate(const float* da, const float* db) {
    long long* dc;
    cudaMalloc(&dc, 8);
    const int rounds = 2000;
    cudaFuncSetAttribute(probe<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128);
    probe<NN><<<148, 128, 256 * 128>>>(da, db, nullptr, rounds, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("A in TMEM, N=%3d: %7.1f cycles per chunk of 24 MMAs (%5.1f per MMA; math floor %d)  %s\n", NN, (double)c / rounds, (double)c / rounds / 24,
           128 * NN / 256, cudaGetErrorString(e));
    cudaFree(dc);
}

int main() {
    std::vector<float> a(M * K), b(N * K), d(M * N);
    for (auto& x : a) x = tf32((float)rand() / RAND_MAX - 0.5f);
    for (auto& x : b) x = tf32((float)rand() / RAND_MAX - 0.5f);
    float *da, *db, *dd;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128);
    probe<N><<<1, 128, 256 * 128>>>(da, db, dd, 0, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)a[m * K + k] * b[n * K + k];
            worst = fmax(worst, fabs(s - d[m * N + n]));
            scale = fmax(scale, fabs(s));
        }
    printf("A-in-TMEM GEMM 128x128x32: max abs err %.3g (scale %.3g)  %s -> %s\n", worst, scale, cudaGetErrorString(e), worst <= 1e-5 * scale ? "PASS" : "FAIL");
    rate<32>(da, db);
    rate<64>(da, db);
    rate<128>(da, db);
    return 0;
}
