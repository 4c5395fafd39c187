// toeplitz_probe.cu -- development probe (not part of the product): one decimating-FIR output tile on tcgen05.
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


// y[r][n] = sum_i h[i] x[r][D n + T - 1 - i + ...] for R = 128 rows (streams) and N = 32 outputs, complex x (I,Q interleaved),
// real taps, as D_p[128 x 32] = X_p[128 x K] * B[32 x K]^T per plane p in {I, Q}, B the (constant) Toeplitz matrix of the taps,
// with the 3xTF32 split (x_hi h_hi + x_lo h_hi + x_hi h_lo) that keeps FP32-level accuracy.  The planes are de-interleaved by
// the conversion pass that also splits every sample into its TF32 high and low parts (a strided TMA box is limited to the
// 128-byte swizzle span of the BOUNDING box, i.e. 16 samples per plane: not worth it).
constexpr int R = 128, N = 32, T = 255, D = 4;
constexpr int KB = 32;                                   // floats per 128-byte swizzle row = one K chunk
constexpr int K = ((D * (N - 1) + T) + KB - 1) / KB * KB; // 384
constexpr int CH = K / KB;                               // 12 chunks
constexpr int UMMA_K = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

constexpr int kChunkA = R * 128;          // bytes of one [128 rows x 128 B] chunk
constexpr int kChunkB = N * 128;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_bh,
                                             const __grid_constant__ CUtensorMap map_bl, float2* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_bh = smem;                                 // CH chunks of [32 rows x 128 B]: taps, high part
    uint8_t* s_bl = s_bh + CH * kChunkB;                  // low part
    uint8_t* s_a = s_bl + CH * kChunkB;                   // I_hi, I_lo, Q_hi, Q_lo: [128 rows x 128 B] each, SWIZZLE_128B layout
    uint8_t* s_raw = s_a + 4 * kChunkA;                   // [128 rows x 64 floats]: 32 interleaved I,Q samples per row, plain
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_raw + 2 * kChunkA);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t bar_b = smem_u32(bars), bar_a = smem_u32(bars + 1), bar_mma = smem_u32(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(R >> 4) << 24);

    if (threadIdx.x == 0) {                                // the constant Toeplitz operand, once
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_b), "r"(2 * CH * kChunkB) : "memory");
        for (int c = 0; c < CH; ++c) {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(s_bh + c * kChunkB)), "l"(&map_bh), "r"(c * KB), "r"(0), "r"(bar_b) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(s_bl + c * kChunkB)), "l"(&map_bl), "r"(c * KB), "r"(0), "r"(bar_b) : "memory");
        }
    }
    mbar_wait(bar_b, 0);
    for (int c = 0; c < CH; ++c) {
        const uint32_t ph = c & 1;
        if (threadIdx.x == 0) {                            // chunk c: 32 samples x 128 rows, interleaved as in memory
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2 * kChunkA) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(s_raw)), "l"(&map_x), "r"(2 * c * KB), "r"(0), "r"(bar_a) : "memory");
        }
        mbar_wait(bar_a, ph);
        // de-interleave and split: thread -> (row, sample); element (row r, k) of a SWIZZLE_128B K-major operand lives at
        // (r / 8) * 1024 + (r % 8) * 128 + (((k / 4) ^ (r % 8)) * 16) + (k % 4) * 4
        for (int i = threadIdx.x; i < R * KB; i += 128) {
            const int r = i / KB, k = i % KB;
            const float2 x = reinterpret_cast<const float2*>(s_raw)[i];
            const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2));
            uint32_t hi, hq;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x.x));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hq) : "f"(x.y));
            *reinterpret_cast<float*>(s_a + 0 * kChunkA + off) = __uint_as_float(hi);
            *reinterpret_cast<float*>(s_a + 1 * kChunkA + off) = x.x - __uint_as_float(hi);
            *reinterpret_cast<float*>(s_a + 2 * kChunkA + off) = __uint_as_float(hq);
            *reinterpret_cast<float*>(s_a + 3 * kChunkA + off) = x.y - __uint_as_float(hq);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int p = 0; p < 2; ++p)
                for (int term = 0; term < 3; ++term)                       // hi*hi, lo*hi, hi*lo
                    for (int k = 0; k < KB / UMMA_K; ++k) {
                        const uint8_t* a = s_a + (2 * p + (term == 1 ? 1 : 0)) * kChunkA;
                        const uint8_t* b = (term == 2 ? s_bl : s_bh) + c * kChunkB;
                        mma_tf32(tmem + p * N, make_desc(smem_u32(a) + k * UMMA_K * 4), make_desc(smem_u32(b) + k * UMMA_K * 4), idesc,
                                 (c | term | k) ? 1u : 0u);
                    }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
        }
        mbar_wait(bar_mma, ph);                            // probe: no ring, the chunk buffers are reused right away
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
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
    for (int n = 0; n < N; ++n) out[(size_t)(warp * 32 + lane) * N + n] = make_float2(__uint_as_float(v[n]), __uint_as_float(v[N + n]));
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float tf32_rna(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    const int n_in = K;                                   // samples per row
    std::vector<float> x((size_t)R * n_in * 2), h(T), bh((size_t)N * K, 0.f), bl((size_t)N * K, 0.f);
    std::vector<float2> y((size_t)R * N);
    srand(2);
    for (auto& v : x) v = (float)(rand() % 65536 - 32768);          // int16-valued, like the radio's samples
    double hs = 0;
    for (int i = 0; i < T; ++i) { const double u = (i - (T - 1) / 2.0) / 40.0; h[i] = (float)(exp(-u * u) * (u == 0 ? 1 : sin(3 * u) / (3 * u))); hs += h[i]; }
    for (auto& v : h) v = (float)(v / hs);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const int i = D * n + T - 1 - k;
            if (i >= 0 && i < T) { bh[(size_t)n * K + k] = tf32_rna(h[i]); bl[(size_t)n * K + k] = h[i] - tf32_rna(h[i]); }
        }
    float *dx, *dbh, *dbl; float2* dy;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dbh, bh.size() * 4)); CK(cudaMalloc(&dbl, bl.size() * 4)); CK(cudaMalloc(&dy, y.size() * 8));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbh, bh.data(), bh.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbl, bl.data(), bl.size() * 4, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    CUtensorMap mx, mbh, mbl;
    {
        // float view [R][2 n_in] of the float2 rows; traversal stride 2 along the row picks one plane (I at even, Q at odd offsets)
        cuuint64_t dims[2] = {(cuuint64_t)(2 * n_in), (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)n_in * 8};
        cuuint32_t box[2] = {2 * KB, R}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode X failed %d\n", (int)r); return 1; }
        cuuint64_t dimsb[2] = {(cuuint64_t)K, (cuuint64_t)N}; cuuint64_t stridesb[1] = {(cuuint64_t)K * 4};
        cuuint32_t boxb[2] = {KB, N}; cuuint32_t es1[2] = {1, 1};
        r = enc(&mbh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dbh, dimsb, stridesb, boxb, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode Bh failed %d\n", (int)r); return 1; }
        r = enc(&mbl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dbl, dimsb, stridesb, boxb, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode Bl failed %d\n", (int)r); return 1; }
    }
    const size_t smem = 2 * CH * kChunkB + 6 * kChunkA + 64 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 128, smem>>>(mx, mbh, mbl, dy);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(y.data(), dy, y.size() * 8, cudaMemcpyDeviceToHost));
    double worst = 0, scale = 0;
    for (int r = 0; r < R; ++r)
        for (int n = 0; n < N; ++n) {
            double si = 0, sq = 0;
            for (int i = 0; i < T; ++i) {
                const int k = D * n + T - 1 - i;
                si += (double)h[i] * x[((size_t)r * n_in + k) * 2];
                sq += (double)h[i] * x[((size_t)r * n_in + k) * 2 + 1];
            }
            const float2 g = y[(size_t)r * N + n];
            worst = fmax(worst, fmax(fabs(g.x - si), fabs(g.y - sq)));
            scale = fmax(scale, fmax(fabs(si), fabs(sq)));
        }
    printf("toeplitz FIR on tcgen05 (3xTF32), %d rows x %d outputs, %d taps / %d: max |err| %.3g, max |ref| %.3g, relative %.2e -> %s\n", R, N, T, D,
           worst, scale, worst / scale, worst <= 1e-5 * scale ? "OK" : "MISMATCH");
    printf("smem %zu bytes, K = %d (%d chunks), useful fraction %.2f\n", smem, K, CH, (double)T / K);
    return 0;
}
