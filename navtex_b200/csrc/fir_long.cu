// fir_long.cu -- register-tiled polyphase decimating FIR kernels for long tap sets (see fir_long.cuh).
#include "fir_long.cuh"

#include <math.h>

namespace nvx {

namespace {

// The taps of the launched stage and the reference NCO table travel in the kernel parameter block (constant bank 0,
// __grid_constant__): per engine, so engines with different tap sets can be alive on one device, and still warp-uniform
// constant-bank operands for the FFMA2 loops.

__device__ __forceinline__ float2 ffma2s(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }

__device__ __forceinline__ float2 load_sample(const LongArgs& a, int row, long long in_pitch, int H, long long g) {
    if (g < 0) return a.hist[(size_t)row * H + (H + g)];
    if (g >= a.n_in) return make_float2(0.f, 0.f);
    if (a.s16) {
        const short2 v = static_cast<const short2*>(a.in)[(size_t)row * in_pitch + g];
        return make_float2((float)v.x, (float)v.y);
    }
    return static_cast<const float2*>(a.in)[(size_t)row * in_pitch + g];
}

// Shared-memory layout of a tile: the inputs in their natural (interleaved) order, f = sample index relative to the
// tile's first input, with one pad slot after every 8 D samples: slot(f) = f + f / (8 D).  A thread's outputs are 8
// apart from its neighbour's, i.e. 8 D inputs: the pad turns that lane stride into 8 D + 1 (odd), so both the staging
// stores (consecutive f) and the window loads (fixed phase, lane stride 8 D + 1) are bank-conflict free, and the
// staging needs no de-interleaving arithmetic.
template <int D>
__device__ __forceinline__ int slot(int f) { return f + f / (kLongR * D); }

// grid (tiles, input rows).  STAGE 0 (the first stage) writes TWO output rows per input row: its 63 kHz outputs rotated by
// each channel's NCO (fir2cpp.C:112-128), so that the second stage is a plain FIR over channel rows.
template <int D, int STAGE>
__global__ void __launch_bounds__(long_threads(D)) fir_long_kernel(const LongArgs a, const int J, const long long in_pitch,
                                                                   const __grid_constant__ LongStageTaps tp) {
    constexpr int kLongThreads = long_threads(D), kLongTile = long_tile(D);
    extern __shared__ __align__(16) float2 s_x[];
    const int H = D * J;
    const int F = D * (kLongTile + J - 1);                    // inputs staged per tile
    const int row_in = blockIdx.y;
    const long long k0 = (long long)blockIdx.x * kLongTile;  // first output of the tile
    const long long n_out = a.n_in / D;
    const long long g0 = (long long)D * (k0 - J + 1);        // first input the tile needs (oldest tap of output k0)

    // ---- stage the tile: a pure copy, so float2 sources go through cp.async (every load of the tile in flight at once);
    // int16 input is converted on the way and keeps the register path ----
    if (a.s16) {
        constexpr int kBatch = 8;
        for (int base = 0; base < F; base += kLongThreads * kBatch) {
            float2 v[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                const int f = base + i * kLongThreads + (int)threadIdx.x;
                v[i] = f < F ? load_sample(a, row_in, in_pitch, H, g0 + f) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                const int f = base + i * kLongThreads + (int)threadIdx.x;
                if (f < F) s_x[slot<D>(f)] = v[i];
            }
        }
    } else {
        const float2* blk = static_cast<const float2*>(a.in) + (size_t)row_in * in_pitch;
        const float2* hist = a.hist + (size_t)row_in * H + H;
        if (g0 >= 0 && g0 + F <= a.n_in) {
            // interior tile (all but the first and the last of a row): no history, no zero fill, and the pad slot of every
            // kLongR * D samples kept incrementally -- the general loop below costs ~20 instructions per sample, which at 65 taps
            // was more than the filter itself (ncu: 4 000 instructions per warp and tile for 640 FFMA2)
            constexpr int kG = kLongR * D, kStepG = kLongThreads / kG, kStepR = kLongThreads % kG;
            int grp = (int)threadIdx.x / kG, rem = (int)threadIdx.x % kG;
            const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s_x);
            const float2* src = blk + g0 + threadIdx.x;
#pragma unroll 4
            for (int f = threadIdx.x; f < F; f += kLongThreads, src += kLongThreads) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s0 + 8u * (uint32_t)(f + grp)), "l"(src) : "memory");
                grp += kStepG;
                rem += kStepR;
                if (rem >= kG) { rem -= kG; ++grp; }
            }
        } else {
            for (int f = threadIdx.x; f < F; f += kLongThreads) {
                const long long g = g0 + f;
                float2* dst = s_x + slot<D>(f);
                if (g >= a.n_in) { *dst = make_float2(0.f, 0.f); continue; }
                const float2* src = g < 0 ? hist + g : blk + g;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    // ---- R consecutive outputs per thread ----
    // Output k0 + R t + u, tap j = jb + jj of phase p reads position R (q - 1) + (u - jj + R - 1) of that phase, q = t +
    // (J - jb) / R, i.e. input f = D pos + p.  The window win[i] holds positions R (q - 1) + i, i < 2 R - 1; since
    // D i + p < 8 D exactly when i < 8, its slots are (8 D + 1)(q - 1) + D i + p + (i >= 8).
    const int t = threadIdx.x;
    float2 acc[kLongR];
#pragma unroll
    for (int u = 0; u < kLongR; ++u) acc[u] = make_float2(0.f, 0.f);
    const float* taps = tp.h;
    constexpr int kLane = kLongR * D + 1;
#pragma unroll 1
    for (int p = 0; p < D; ++p) {
        const float* hp = taps + p * J;
        const float2* sp = s_x + kLane * (t + J / kLongR - 1) + p;     // window base of tap block 0
        float2 win[2 * kLongR - 1];
#pragma unroll
        for (int i = 0; i < 2 * kLongR - 1; ++i) win[i] = sp[D * i + (i >= kLongR)];
#pragma unroll 1
        for (int jb = 0; jb < J; jb += kLongR) {
#pragma unroll
            for (int jj = 0; jj < kLongR; ++jj) {
                const float h = hp[jb + jj];
#pragma unroll
                for (int u = 0; u < kLongR; ++u) acc[u] = ffma2s(win[u - jj + kLongR - 1], h, acc[u]);
            }
            // slide one tap block towards older samples
            sp -= kLane;
#pragma unroll
            for (int i = 2 * kLongR - 2; i >= kLongR; --i) win[i] = win[i - kLongR];
            if (jb + kLongR < J) {
#pragma unroll
                for (int i = 0; i < kLongR; ++i) win[i] = sp[D * i];
            }
        }
    }
    if (STAGE == 0 && !a.plain) {
        // NCO mix: output k sits at 63 kHz clock tick k_abs + k; channel c gets y1 * (cos - j sin)(2 pi tick f_c / 63000)
        float2* out0 = a.out + (size_t)(2 * row_in) * a.out_pitch + a.out_off + k0 + (long long)kLongR * t;
        float2* out1 = out0 + a.out_pitch;
        NcoParam np = {};
        if (a.nco) np = a.nco[row_in];
        const long long tick0 = a.k_abs + k0 + (long long)kLongR * t;
        long long r9 = tick0 % kNcoPeriod, rden = tick0 % kNcoDen;
        int k9 = (int)(r9 < 0 ? r9 + kNcoPeriod : r9), kden = (int)(rden < 0 ? rden + kNcoDen : rden);
#pragma unroll
        for (int u = 0; u < kLongR; ++u) {
            float2 rot[2];
            if (a.nco) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int ph = (int)(((long long)kden * np.num[c]) % kNcoDen);
                    float tt = (float)ph * (2.0f / kNcoDen);
                    if (tt > 1.0f) tt -= 2.0f;
                    float sn, cs;
                    sincospif(tt, &sn, &cs);
                    rot[c] = make_float2(cs, -sn);
                }
            } else {
                rot[0] = tp.nco[k9];
                rot[1] = make_float2(rot[0].x, -rot[0].y);     // "490": conjugate rotation (fir2cpp.C:121-124)
            }
            if (k0 + (long long)kLongR * t + u < n_out) {
                const float2 y = acc[u];
                out0[u] = make_float2(fmaf(-y.y, rot[0].y, y.x * rot[0].x), fmaf(y.x, rot[0].y, y.y * rot[0].x));
                out1[u] = make_float2(fmaf(-y.y, rot[1].y, y.x * rot[1].x), fmaf(y.x, rot[1].y, y.y * rot[1].x));
            }
            if (++k9 == kNcoPeriod) k9 = 0;
            if (++kden == kNcoDen) kden = 0;
        }
    } else {
        float2* out = a.out + (size_t)row_in * a.out_pitch + a.out_off + k0 + (long long)kLongR * t;
#pragma unroll
        for (int u = 0; u < kLongR; ++u)
            if (k0 + (long long)kLongR * t + u < n_out) out[u] = acc[u];
    }
}

template <typename Sample>
__global__ void long_carry_kernel(const float2* __restrict__ old_hist, const Sample* __restrict__ block, long long pitch,
                                  float2* __restrict__ new_hist, int rows, int H, long long n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * H) return;
    const int row = (int)(idx / H), i = (int)(idx % H);
    const long long src = n - H + i;
    float2 v;
    if (src >= 0) {
        const Sample s = block[(size_t)row * pitch + src];
        v = make_float2((float)s.x, (float)s.y);
    } else {
        v = old_hist[(size_t)row * H + (H + src)];
    }
    new_hist[(size_t)row * H + i] = v;
}

template <int D, int STAGE>
cudaError_t launch_one(const LongArgs& a, const LongStage& st, const LongStageTaps& tp, long long in_pitch, cudaStream_t stream) {
    constexpr int kLongThreads = long_threads(D), kLongTile = long_tile(D);
    const int F = D * (kLongTile + st.J - 1);
    const size_t smem = (size_t)(F + F / (kLongR * D) + 2) * sizeof(float2);
    {   // per device: not cached
        cudaError_t e = cudaFuncSetAttribute(fir_long_kernel<D, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long n_out = a.n_in / D;
    fir_long_kernel<D, STAGE><<<dim3((unsigned)((n_out + kLongTile - 1) / kLongTile), (unsigned)a.rows_in), kLongThreads, smem, stream>>>(
        a, st.J, in_pitch, tp);
    return cudaGetLastError();
}

}  // namespace

LongStage long_stage(int D, int T) {
    LongStage s;
    s.D = D;
    s.T = T;
    const int per_phase = (T + D - 1) / D;
    s.J = (per_phase + kLongR - 1) / kLongR * kLongR;
    s.H = D * s.J;
    return s;
}

// host side of one stage's tap block: [p * J + j] = h[D j + D - 1 - p] (0 beyond the set), plus the reference NCO table
bool long_fill_taps(const LongStage& st, const double* h, LongStageTaps* out) {
    if (st.T < 1 || st.T > kLongMaxTaps || st.D * st.J > kLongTapSlots) return false;
    for (int k = 0; k < kLongTapSlots; ++k) out->h[k] = 0.f;
    for (int p = 0; p < st.D; ++p)
        for (int j = 0; j < st.J; ++j) {
            const int i = st.D * j + st.D - 1 - p;
            out->h[p * st.J + j] = i < st.T ? (float)h[i] : 0.f;
        }
    for (int k = 0; k < kNcoPeriod; ++k)       // same expression as fir2cpp.C:105-106, rounded once to float
        out->nco[k] = make_float2((float)cos((2 * M_PI * k * 14000) / 63000), (float)-sin((2 * M_PI * k * 14000) / 63000));
    return true;
}

cudaError_t long_launch(const LongArgs& a, const LongStage& st, const LongStageTaps& tp, long long in_pitch, cudaStream_t stream) {
    switch (a.stage) {
        case 0: return launch_one<NVX_D1, 0>(a, st, tp, in_pitch, stream);
        case 1: return launch_one<NVX_D2, 1>(a, st, tp, in_pitch, stream);
        default: return launch_one<NVX_D3, 2>(a, st, tp, in_pitch, stream);
    }
}

cudaError_t long_carry(const float2* old_hist, const void* block, long long pitch, float2* new_hist, int rows, int H, long long n,
                       int s16, cudaStream_t stream) {
    const long long work = (long long)rows * H;
    const unsigned grid = (unsigned)((work + 255) / 256);
    if (s16)
        long_carry_kernel<short2><<<grid, 256, 0, stream>>>(old_hist, static_cast<const short2*>(block), pitch, new_hist, rows, H, n);
    else
        long_carry_kernel<float2><<<grid, 256, 0, stream>>>(old_hist, static_cast<const float2*>(block), pitch, new_hist, rows, H, n);
    return cudaGetLastError();
}

}  // namespace nvx
