// demod.cuh -- per-channel persistent state of the FSK demodulator / bit synchroniser / SITOR-B
// state machine, and the launch arguments of demod_kernel (demod.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nvx {

constexpr int kSpb = 9;                 // samples per bit at 900 Hz (decoder.h:21)
constexpr int kCorrLen = 63 * kSpb;     // 567 (decoder.h:23-24)
constexpr int kCorrRing = 1024;         // power-of-two time-indexed history >= kCorrLen + 32
constexpr int kEvAbort = 0x18;          // event byte: message_abort() (nav_b_sm.C:44-52)

// Everything decoder.{h,C} and nav_b_sm.{h,C} keep per channel between samples, re-expressed so
// that ring positions are functions of the absolute 900 Hz sample count `seen`.
struct ChannelScalars {
    double prev_i, prev_q;    // decoder.C:54-55
    long long seen;           // 900 Hz samples consumed so far
    int last_pick;            // prev_offset (decoder.C:247), -1 = none
    // mark/space discriminator (decoder.C:73-137)
    int dstate, offs, next_offs, burned, used;
    float br, bi, yr, yi;
    // SITOR-B state machine (nav_b_sm.h:92-116), arrays packed into words
    int match;                // phasing pattern bits matched (status)
    int byte_state, figures, nbits, shift;
    unsigned dx_ring;         // 3 bytes, slot k at bits [8k, 8k+8)
    int dx_at, dx_full;
    unsigned err_mask;        // bit k set = error_buffer[k] holds an invalid code
    int err_at, err_full, err_count;
    int eoe_count, prev_dx_alpha, holdoff, enabled;
    int pad_;
};

struct ChannelState {
    double corr[kCorrRing];   // |mask correlation| history, value number v at [v & 1023] (decoder.C:170)
    double ang[8];            // last 8 discriminator angles, oldest first (decoder.C:147)
    double osum[kSpb];        // per-offset correlation sums (decoder.C:186-190)
    ChannelScalars sc;
};

struct DemodArgs {
    const float2* y3;         // [channels][y3_pitch]
    long long y3_pitch;
    long long y3_off;         // first new sample of every channel row
    int n_new;                // new 900 Hz samples per channel
    int channels;             // streams * 2
    ChannelState* state;      // [channels]
    // per-launch outputs
    uint8_t* events;          // [channels][ev_cap]: appended characters, '\n' = line complete, 0x18 = abort
    int* ev_count;            // [channels]
    int ev_cap;
    char* bits;               // optional [channels][bit_cap] 'B'/'Y' decisions (debug / parity taps), may be null
    float* disc;              // optional [channels][bit_cap][4] BR BI YR YI at each decision, may be null
    int* bit_count;           // [channels] (required if bits != null)
    int bit_cap;
};

cudaError_t demod_launch(const DemodArgs& a, cudaStream_t stream);
cudaError_t demod_init_state(ChannelState* state, int channels, cudaStream_t stream);

}  // namespace nvx
