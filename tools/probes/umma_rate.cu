// umma_rate.cu -- development probe: issue rate of the long-tap kernel's MMA pattern with both operands in shared memory
// (tcgen05.mma.kind::tf32, M = 128, K = 8 per instruction): cycles per "chunk" of 2 planes x 3 terms x 4 k-steps for
// N = 32 .. 256, with and without eight warps storing operand tiles beside it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
    return p != 0;
}

constexpr int kOp = 128 * 128;

// mode 0: MMAs only; 1: warps 1..8 also store 64 floats per thread per round (the converters' traffic); order: 0 = (p, term, k), 1 = (p, k, term)
template <int N>
__global__ void __launch_bounds__(288, 1) rate(int rounds, int mode, int order, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_a = smem;                    // 2 sets x 4 tiles
    uint8_t* s_b = smem + 8 * kOp;          // 2 parts x 256 rows
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (8 * kOp + 2 * 256 * 128) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])));
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
    if (warp == 0) {
        // the whole warp runs the loop and ONE elected lane issues: descriptors stay in uniform registers.  (A lane == 0 branch
        // instead makes ptxas wrap every tcgen05.mma in an R2UR waterfall loop: measured 74.6 cycles per MMA for any N.)
        const bool leader = elect_one();
        {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const long long t0 = clock64();
            for (int r = 0; r < rounds; ++r) {
                const int s = r & 1;
                if (r >= 2) bar_wait(s_u32(&bars[s]), ((r >> 1) - 1) & 1);
                const uint32_t op = s_u32(s_a + s * 4 * kOp);
                const uint32_t bh = s_u32(s_b) + (r % 12) * 1024, bl = s_u32(s_b + 256 * 128) + (r % 12) * 1024;
                if (!leader) continue;
                if (order == 0) {
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int term = 0; term < 3; ++term)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_tf32(tmem + p * N, umma_desc(op + (2 * p + (term == 1 ? 1 : 0)) * kOp + k * 32), umma_desc((term == 2 ? bl : bh) + k * 32), idesc, 1u);
                } else {
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int term = 0; term < 3; ++term)      // hi hi, hi lo, lo hi: the same A twice in a row
                                umma_tf32(tmem + p * N, umma_desc(op + (2 * p + (term == 2 ? 1 : 0)) * kOp + k * 32), umma_desc((term == 1 ? bl : bh) + k * 32), idesc, 1u);
                }
                umma_commit(s_u32(&bars[s]));
            }
            bar_wait(s_u32(&bars[(rounds - 1) & 1]), ((rounds - 1) >> 1) & 1);
            if (rounds >= 2) bar_wait(s_u32(&bars[rounds & 1]), ((rounds - 2) >> 1) & 1);
            const long long t1 = clock64();
            if (blockIdx.x == 0 && leader) *cycles = t1 - t0;
        }
    } else if (mode == 1) {
        const int w = warp - 1;
        const int off0 = w * 128 + ((((lane >> 2) ^ w) << 4) | ((lane & 3) << 2));
        for (int r = 0; r < rounds; ++r) {      // unsynchronised: only the store traffic matters here
            uint8_t* op = s_a + (r & 1) * 4 * kOp + off0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
#pragma unroll
                for (int t = 0; t < 4; ++t) *reinterpret_cast<volatile float*>(op + t * kOp + i * 1024) = 0.f;
            __nanosleep(200);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int N>
void run(int mode, int order) {
    long long* d;
    cudaMalloc(&d, 8);
    const size_t smem = 8 * kOp + 2 * 256 * 128 + 1024;
    cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int rounds = 2000;
    rate<N><<<148, 288, smem>>>(rounds, mode, order, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d mode=%d order=%d: %7.1f cycles per chunk of 24 MMAs (%5.1f per MMA; math floor %d)  %s\n", N, mode, order, (double)c / rounds,
           (double)c / rounds / 24, 128 * N / 256, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int mode = 0; mode < 2; ++mode)
        for (int order = 0; order < 2; ++order) {
            run<32>(mode, order);
            run<64>(mode, order);
            run<128>(mode, order);
        }
    return 0;
}
