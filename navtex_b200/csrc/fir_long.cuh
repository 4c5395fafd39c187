// fir_long.cuh -- the three decimating FIR stages as separate, register-tiled kernels for tap sets LONGER than the
// reference's 37 / 47 / 71 (BASELINE.json configs[4]: "long-tap FIR stress, 255+ taps").
//
// Same arithmetic definition as the reference stages:
//   y[k] = sum_{i < T} h[i] x[D (k + 1) - 1 - i]     fir1cpp.C:80-136 (D = 4), fir2cpp.C:131-215 (D = 7, after the
//                                                     NCO mix of fir2cpp.C:112-128), fir3cpp.C:22-60 (D = 10)
// but a different machine mapping from the fused cascade: with hundreds of taps the partial sums of the transposed
// form no longer fit the register file and the work is FP32-bound (~330 flop per input sample at 255 taps, 42 flop/B),
// so the intermediates may as well round-trip through HBM (+4.6 B per input sample) and every stage becomes a plain
// polyphase, register-tiled FIR:
//   * a CTA owns (row, tile of kLongTile consecutive outputs); the inputs that tile needs are staged in shared memory
//     de-interleaved by decimation phase (x_p[m] = x[D m + p]) so that lanes read consecutive addresses;
//   * a thread owns R = 8 consecutive outputs; per phase it slides an R-wide register window over x_p and applies the
//     taps of that phase (warp-uniform constant-bank operands) as packed FFMA2 on (I, Q): R * R FFMA2 per R loads;
//   * the stage-1 kernel applies the NCO rotation of both channels in its epilogue (9-entry table of the reference or
//     the exact per-stream phase of the general NCO) and writes one 63 kHz row per channel, so stages 2 and 3 are
//     plain FIRs over channel rows -- unless stage 2 runs on the streaming tensor-core kernel with the reference table:
//     then stage 1 writes ONE un-mixed row per stream and stage 2 rotates while it loads (LongArgs.plain / mix_in).
// History between blocks is carried per stage (last H inputs of each row), not recomputed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "fir_cascade.cuh"

namespace nvx {

constexpr int kLongR = 8;                       // outputs per thread
// threads per CTA (each owns kLongR outputs).  The tile a CTA stages is D (tile + J - 1) samples: at D = 10 a 1024-output tile is
// 84 KB of shared memory, two CTAs = 8 warps per SM, and the kernel ran at 42 % FMA-pipe activity for lack of warps; 64 threads
// (512 outputs, 43 KB) put five CTAs on an SM and also waste less on the ragged last tile of a 9250-output row.
__host__ __device__ constexpr int long_threads(int D) { return D == NVX_D3 ? 64 : 128; }
__host__ __device__ constexpr int long_tile(int D) { return kLongR * long_threads(D); }   // outputs per CTA
constexpr int kLongMaxTaps = 1024;              // per stage

struct LongStage {
    int D;            // decimation
    int T;            // taps
    int J;            // taps per phase, padded to a multiple of kLongR: D * J >= T
    int H;            // history samples carried per row: D * J
};

struct LongArgs {
    const void* in;           // [rows_in][in_pitch] float2 (or short2 for stage 1 with s16 input), this block
    const float2* hist;       // [rows_in][H] float2: the H samples that preceded the block (always float2)
    float2* out;              // [rows_out][out_pitch], written at out_off + k; stage 1: rows_out = 2 rows_in (one per channel)
    long long n_in;           // input samples per row in this block (multiple of D)
    long long out_pitch, out_off;
    int rows_in;              // streams (stage 1) or channels (stages 2, 3)
    int stage;                // 0, 1, 2
    int s16;                  // stage 1 only: input block is short2
    long long k_abs;          // stage 1: absolute index of the block's first OUTPUT sample (63 kHz clock), for the NCO
    const NcoParam* nco;      // stage 1: per-stream general NCO or null (reference table)
    // "mix on load" (reference NCO table only): stage 1 leaves its output un-mixed, ONE 63 kHz row per stream (plain = 1), and the
    // tensor-core stage 2 applies the rotation of its row's channel while it converts the samples (mix_in = 1: the input has
    // rows_in / 2 rows, k_abs is the absolute index of the block's first INPUT sample) -- the same FP32 products in the same order
    // as the stage-1 epilogue, with half the y1 traffic on both sides
    int plain, mix_in;
};

// taps of one stage regrouped per decimation phase + the reference NCO table: passed by value in the kernel parameter block
constexpr int kLongTapSlots = 1152;             // D * J <= 1024 + D * kLongR
struct LongStageTaps {
    float h[kLongTapSlots];
    float2 nco[kNcoPeriod];
};

LongStage long_stage(int D, int T);
bool long_fill_taps(const LongStage& st, const double* h, LongStageTaps* out);
cudaError_t long_launch(const LongArgs& a, const LongStage& st, const LongStageTaps& tp, long long in_pitch, cudaStream_t stream);
// new_hist = last H samples of (old_hist ++ block[.][0..n)) per row; block rows are `pitch` samples apart, short2 if s16
cudaError_t long_carry(const float2* old_hist, const void* block, long long pitch, float2* new_hist, int rows, int H, long long n,
                       int s16, cudaStream_t stream);

// ---- tensor-core variant of stages 1 and 2 (fir_long_tc.cu): tcgen05 3xTF32 Toeplitz GEMM, same contract as long_launch ----
struct LongTcStage;
struct TcBand {
    int T, N, chunks, J, copies;      // padded taps, outputs per tile, K chunks per tile, band rows per copy, copies
    bool streaming = false;           // served by the streaming kernel (every input chunk feeds the two tiles that contain it)
    std::vector<float> gh, gl;        // [copies][J][32] high / low TF32 parts
};
bool long_tc_band(int D, int T_taps, const double* h, TcBand* out);
// builds the Toeplitz operand of one stage (D = 4 or 7, T taps) on the device; nullptr when the stage does not fit the kernel
LongTcStage* long_tc_prepare(int D, int T, const double* h, cudaStream_t stream);
void long_tc_free(LongTcStage* s);
bool long_tc_mixes_on_load(const LongTcStage* s);      // stage 2 served by the streaming kernel: LongArgs.mix_in is available
cudaError_t long_tc_launch(const LongTcStage* s, const LongArgs& a, const LongStage& st, long long in_pitch, cudaStream_t stream);

}  // namespace nvx
