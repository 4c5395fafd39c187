// tma_probe.cu -- development probe: what HBM read bandwidth does the cascade kernel's access pattern
// (one TMA box of [32 streams x B bytes] per warp per step, rows 20 MB apart) reach with no compute at all?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_probe.cu
// Usage: ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
    CUtensorMap map;
    int groups, segs, steps_per_seg, box_floats, step_floats, stages, warps;
    int split;      // boxes per stage: the row is fetched as `split` boxes of box_floats / split floats each (swizzle experiments)
    float* sink;
};

__global__ void probe(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stage_bytes = 32 * a.box_floats * 4;
    const int stage_pitch = (stage_bytes + 127) & ~127;
    uint8_t* wbase = smem + (size_t)warp * a.stages * stage_pitch;
    const uint32_t wbase_s = smem_u32(wbase);
    const uint32_t bar0 = smem_u32(smem + (size_t)a.warps * a.stages * stage_pitch) + warp * a.stages * 8;
    if (lane == 0) {
        for (int s = 0; s < a.stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * s), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long wg = (long long)blockIdx.x * a.warps + warp;
    if (wg >= (long long)a.groups * a.segs) return;
    const int seg = (int)(wg / a.groups), row0 = (int)(wg % a.groups) * 32;
    const long long col0 = (long long)seg * a.steps_per_seg * a.step_floats;
    auto issue = [&](int t, int stage) {
        if (lane == 0) {
            const uint32_t bar = bar0 + 8 * stage;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
            const int part = a.box_floats / a.split;
            for (int b = 0; b < a.split; ++b)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(wbase_s + stage * stage_pitch + b * 32 * part * 4), "l"(&a.map),
                               "r"((int)(col0 + (long long)t * a.step_floats) + b * part), "r"(row0), "r"(bar) : "memory");
        }
    };
    for (int s = 0; s < a.stages; ++s) if (s < a.steps_per_seg) issue(s, s);
    int stage = 0; uint32_t parity = 0; float acc = 0.f;
    for (int t = 0; t < a.steps_per_seg; ++t) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                     ::"r"(bar0 + 8 * stage), "r"(parity) : "memory");
        acc += *reinterpret_cast<const float*>(wbase + stage * stage_pitch + lane * a.box_floats * 4);
        __syncwarp();
        if (t + a.stages < a.steps_per_seg) issue(t + a.stages, stage);
        if (++stage == a.stages) { stage = 0; parity ^= 1; }
    }
    if (acc == 123.456f) a.sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int S = 1024;
    const long long n = 2590000;      // samples per stream (float2)
    float* x; float* sink;
    CK(cudaMalloc(&x, (size_t)S * n * 8));
    CK(cudaMemset(x, 0, (size_t)S * n * 8));
    CK(cudaMalloc(&sink, 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int step_floats, box_floats, stages, warps, ctas_per_sm, promo, split; };
    const Cfg cfgs[] = {
        {112, 112, 3, 4, 1, 2, 1}, {112, 112, 4, 4, 1, 2, 1}, {112, 112, 4, 4, 1, 2, 7}, {112, 112, 3, 4, 1, 2, 7}, {112, 128, 3, 4, 1, 2, 4}, {112, 128, 3, 4, 1, 2, 1},
        {56, 56, 3, 4, 2, 2, 1},  {56, 56, 6, 4, 1, 2, 1},
        {112, 112, 3, 2, 2, 2, 1}, {112, 116, 3, 2, 2, 2, 1},
        {224, 224, 2, 3, 1, 2, 1}, {224, 224, 3, 2, 1, 2, 1},
    };
    for (const Cfg& c : cfgs) {
        Args a;
        cuuint64_t dims[2] = {(cuuint64_t)(2 * n), (cuuint64_t)S};
        cuuint64_t strides[1] = {(cuuint64_t)(n * 8)};
        cuuint32_t box[2] = {(cuuint32_t)(c.box_floats / c.split), 32};
        cuuint32_t es[2] = {1, 1};
        CUtensorMapL2promotion promo = c.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : c.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        CUresult r = enc(&a.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int total_warps = 148 * c.ctas_per_sm * c.warps;
        a.groups = S / 32;
        a.segs = total_warps / a.groups;
        const long long total_steps = 2 * n / c.step_floats;
        a.steps_per_seg = (int)(total_steps / a.segs);
        a.box_floats = c.box_floats; a.step_floats = c.step_floats; a.stages = c.stages; a.warps = c.warps; a.sink = sink; a.split = c.split;
        const int stage_pitch = (32 * c.box_floats * 4 + 127) & ~127;
        const size_t smem = (size_t)c.warps * c.stages * stage_pitch + c.warps * c.stages * 8;
        const int grid = (a.groups * a.segs + c.warps - 1) / c.warps;
        float best = 1e9f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            probe<<<grid, c.warps * 32, smem>>>(a);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it && ms < best) best = ms;
        }
        const double bytes = (double)a.groups * a.segs * 32.0 * a.steps_per_seg * c.step_floats * 4.0;
        printf("row %4d B (box %4d B x %d) stages %d warps/cta %d ctas/sm %d promo %d smem/cta %6zu segs %3d: %.3f ms  %.0f GB/s\n", c.step_floats * 4,
               c.box_floats / c.split * 4, c.split, c.stages, c.warps, c.ctas_per_sm, c.promo, smem, a.segs, best, bytes / best / 1e6);
    }
    return 0;
}
