// umma_probe.cu -- development probe (not part of the product): one CTA, one tcgen05 TF32 tile.
// D[128 x N] = A[128 x K] * B[N x K]^T, A and B K-major in global memory, TMA (SWIZZLE_128B) -> shared memory,
// tcgen05.mma.kind::tf32 with the accumulator in TMEM, tcgen05.ld back to registers.  Purpose: establish the shared
// memory / instruction descriptors on this toolchain before a tensor-core long-tap FIR is attempted.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int M = 128, N = 64, K = 64, KB = 32;      // KB floats = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;                            // tf32: 32 bytes per instruction along K

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset >> 4 in [16,30) (unused for swizzled K-major: 1), stride byte offset >> 4 in [32,46) = 8 rows x 128 B,
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* d_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sa = smem;                                   // 2 k-blocks x [128 rows x 128 B]
    uint8_t* sb = smem + 2 * M * 128;                     // 2 k-blocks x [64 rows x 128 B]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + 2 * M * 128 + 2 * N * 128);
    uint64_t* bar_mma = bar_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                      // one warp allocates 64 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    if (threadIdx.x == 0) {
        const uint32_t bar = smem_u32(bar_full);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * M * 128 + 2 * N * 128) : "memory");
        for (int kb = 0; kb < K / KB; ++kb) {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sa + kb * M * 128)), "l"(&map_a), "r"(kb * KB), "r"(0), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sb + kb * N * 128)), "l"(&map_b), "r"(kb * KB), "r"(0), "r"(bar) : "memory");
        }
        // wait for the tiles
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;");
        // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format TF32 = 2 at [7,10) / [10,13),
        // K-major both, n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kb = 0; kb < K / KB; ++kb)
            for (int k = 0; k < KB / UMMA_K; ++k) {
                const uint64_t da = make_desc(smem_u32(sa + kb * M * 128) + k * UMMA_K * 4);
                const uint64_t db = make_desc(smem_u32(sb + kb * N * 128) + k * UMMA_K * 4);
                const uint32_t acc = (kb | k) ? 1u : 0u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)) : "memory");
    }
    // everyone waits for the MMA to finish
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar_mma)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    // warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows of D), 64 columns
    uint32_t v[64];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]),
          "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]),
          "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < N; ++n) d_out[(size_t)(warp * 32 + lane) * N + n] = __uint_as_float(v[n]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float tf32(float x) {      // round to 10 mantissa bits (what the tensor core sees, up to its own rounding mode)
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    std::vector<float> a((size_t)M * K), b((size_t)N * K), d((size_t)M * N), ref((size_t)M * N);
    srand(1);
    for (auto& x : a) x = tf32((float)rand() / RAND_MAX - 0.5f);
    for (auto& x : b) x = tf32((float)rand() / RAND_MAX - 0.5f);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)a[(size_t)m * K + k] * b[(size_t)n * K + k];
            ref[(size_t)m * N + n] = (float)s;
        }
    float *da, *db, *dd;
    CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dd, d.size() * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dd, 0, d.size() * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {KB, M}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
        cuuint64_t dimsb[2] = {(cuuint64_t)K, (cuuint64_t)N};
        cuuint32_t boxb[2] = {KB, N};
        r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    const size_t smem = 2 * M * 128 + 2 * N * 128 + 64 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 128, smem>>>(ma, mb, dd);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, scale = 0;
    for (size_t i = 0; i < d.size(); ++i) { worst = fmax(worst, fabs((double)d[i] - ref[i])); scale = fmax(scale, fabs((double)ref[i])); }
    printf("umma tf32 128x%dx%d: max |err| %.3g (max |ref| %.3g) -> %s\n", N, K, worst, scale, worst <= 1e-4 * scale ? "OK" : "MISMATCH");
    printf("d[0..3] = %g %g %g %g   ref = %g %g %g %g\n", d[0], d[1], d[2], d[3], ref[0], ref[1], ref[2], ref[3]);
    return 0;
}
