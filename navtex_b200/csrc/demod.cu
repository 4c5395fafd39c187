// demod.cu -- FSK discriminator, bit synchroniser, mark/space decision and SITOR-B byte state
// machine for [channels] independent 900 Hz complex streams.
//
// Replaces, per channel and per block instead of per sample,
//   decoder::sample_in            receiver/decoder.C:42-59    (delay-conjugate product + atan2)
//   decoder::bs_decoded_sample_in receiver/decoder.C:142-255  (bit-sync timing recovery)
//   decoder::bd_decoded_sample_in receiver/decoder.C:73-137   (mark/space energy discriminator)
//   byte_state_machine::receive_bit / receive_rxdx_byte / message_byte_out
//                                 receiver/nav_b_sm.C:266-634, :150-262, :100-145
// Line assembly, the ZCZC/NNNN regexes and add_message (nav_b_sm.C:56-97, :44-52) stay on the
// host (message_assembler.cpp); the last kernel emits the character/line/abort event stream they
// consume.
//
// The reference interleaves everything per sample, but almost all of it is feed-forward in time.
// Three kernels per block:
//   angle_corr_kernel   thread = sample: FP64 angle of y[n] conj(y[n-1]) and the 9-tap transition
//                       mask correlation, in the reference's operation order;
//   sum_decide_kernel   thread = sample: the 63-term per-offset sum "as the 567-ring stood at that
//                       sample" (same ascending-slot summation order), the first-maximum arg max over
//                       the nine sums at evaluation samples, and -- for EVERY possible bit start --
//                       the mark/space decision of the 5-sample window starting there, with the
//                       reference's mixed float/double accumulator arithmetic (_rn intrinsics, no
//                       FMA contraction).  One byte per sample comes out;
//   symbol_clock_kernel thread = channel: the only truly sequential part -- slew-limited offset
//                       tracking, the WAIT/BIT_START/RECEIVING symbol clock (which merely selects
//                       which precomputed decision is a bit), and the byte state machine.
#include "demod.cuh"

#include <math.h>
#include <stdlib.h>

namespace nvx {

namespace {

constexpr int kTile = 256;

enum { DS_INIT = 0, DS_WAIT = 1, DS_PENDING = 2 };   // decoder.h:16-19 (BIT_START / RECEIVING folded into PENDING)
enum { BY_WAIT = 1, BY_GOT_DX = 2, BY_GOT_RX = 3 };                       // nav_b_sm.h:41-43

__constant__ unsigned char c_ltrs[128];
__constant__ unsigned char c_figs[128];
__constant__ float c_tone_r[5];
__constant__ float c_tone_i[5];
__constant__ double c_tone_rd[5];   // the same tones widened to double
__constant__ double c_tone_id[5];

struct Emit {
    uint8_t* ev; int* ev_n; int ev_cap; int n;
    bool writer;
    const unsigned char *ltrs, *figs;      // code tables in shared memory (per-lane indices would serialise the constant cache)
    __device__ __forceinline__ void put(int c) {
        if (writer && n < ev_cap) ev[n] = (uint8_t)c;
        ++n;
    }
};

// byte_state_machine::init, nav_b_sm.C:16-42 (the error ring contents survive, only its counters reset)
__device__ __forceinline__ void fsm_reset(ChannelState& s) {
    s.match = 0; s.byte_state = BY_WAIT; s.figures = 0; s.nbits = 0;
    s.dx_at = 0; s.dx_full = 0;
    s.err_count = 0; s.err_at = 0; s.err_full = 0;
    s.eoe_count = 0; s.prev_dx_alpha = 0;
    s.holdoff = 0; s.enabled = 0;
}

// message_byte_out, nav_b_sm.C:100-145; code 0 = "no valid copy" -> '*'
__device__ __forceinline__ void fsm_char(ChannelState& s, Emit& e, int code) {
    if (code == 0) { e.put('*'); return; }
    const int l = e.ltrs[code];
    if (l == 'l') { s.figures = 0; return; }
    if (l == 'f') { s.figures = 1; return; }
    if (l == 'n') { e.put('\n'); return; }
    if (l == 'r' || l == 'p' || l == 'q') return;
    e.put(s.figures ? e.figs[code] : l);
}

// message_abort, nav_b_sm.C:44-52: the host decides whether a message was in progress
__device__ __forceinline__ void fsm_abort(ChannelState& s, Emit& e) {
    e.put(kEvAbort);
    fsm_reset(s);
}

// receive_rxdx_byte, nav_b_sm.C:150-262
__device__ __forceinline__ void fsm_byte(ChannelState& s, Emit& e, int b) {
    if (s.byte_state == BY_WAIT) {
        if (b == 0x07) s.byte_state = BY_GOT_RX;
        if (b == 0x4c) s.byte_state = BY_GOT_DX;
    } else if (s.byte_state == BY_GOT_RX) {           // byte in the DX slot
        s.dx_ring = (s.dx_ring & ~(0xffu << (8 * s.dx_at))) | ((unsigned)b << (8 * s.dx_at));
        if (++s.dx_at == 3) { s.dx_at = 0; s.dx_full = 1; }
        bool stopped = false;
        if (b == 0x07) {
            if (s.prev_dx_alpha && ++s.eoe_count == 2) { fsm_abort(s, e); stopped = true; }   // end of emission
            if (!stopped) s.prev_dx_alpha = 1;
        } else {
            s.prev_dx_alpha = 0;
        }
        if (!stopped) s.byte_state = BY_GOT_DX;
    } else {                                           // BY_GOT_DX: byte in the RX slot
        if (s.dx_full) {
            const int dx = (s.dx_ring >> (8 * s.dx_at)) & 0x7f;
            if (e.ltrs[b] != '_') fsm_char(s, e, b);
            else if (e.ltrs[dx] != '_') fsm_char(s, e, dx);
            else fsm_char(s, e, 0);
        }
        s.byte_state = BY_GOT_RX;
    }
    // 20-byte sliding window of invalid codes (nav_b_sm.C:235-261)
    const unsigned bit = 1u << s.err_at;
    if (s.err_full && (s.err_mask & bit)) s.err_count--;
    const bool bad = e.ltrs[b] == '_';
    s.err_mask = bad ? (s.err_mask | bit) : (s.err_mask & ~bit);
    if (bad) s.err_count++;
    if (++s.err_at == 20) { s.err_at = 0; s.err_full = 1; }
    if (s.err_count > 12) {
        fsm_char(s, e, 0);
        fsm_abort(s, e);
    }
}

// receive_bit, nav_b_sm.C:266-634.  is_y: 'Y' (=1) else 'B'.
__device__ __forceinline__ void fsm_bit(ChannelState& s, Emit& e, bool is_y) {
    if (s.enabled) {
        s.shift = ((s.shift << 1) | (is_y ? 1 : 0)) & 0x7f;
        if (++s.nbits == 7) {
            fsm_byte(s, e, s.shift);
            s.nbits = 0;
            s.shift = 0;
        }
    }
    if (s.holdoff != 0) { s.holdoff--; return; }
    // 30-bit phasing pattern BBBBBB YYYY BB YY BBBBBB YYYY BB YY BB, bit k of the mask = 1 for 'Y'
    // positions of 'Y': 6-9, 12-13, 20-23, 26-27
    constexpr unsigned kYmask = (0xFu << 6) | (0x3u << 12) | (0xFu << 20) | (0x3u << 26);
    if (s.match == 29) {
        if (!is_y) { s.enabled = 1; s.nbits = 0; s.shift = 0; s.holdoff = 1100; }   // nav_b_sm.h:52
        s.match = 0;
    } else if ((((kYmask >> s.match) & 1u) != 0) == is_y) {
        s.match++;
    } else if (s.match != 6) {          // extra B's are tolerated only after the first BBBBBB (nav_b_sm.C:363-372)
        s.match = 0;
    }
}


__device__ __forceinline__ size_t pitch_y(int p_max) { return (size_t)kHistY + p_max; }
__device__ __forceinline__ size_t pitch_c(int p_max) { return (size_t)kHistC + p_max; }
__device__ __forceinline__ size_t pitch_d(int p_max) { return ((size_t)kHistD + p_max + 15) & ~(size_t)15; }

// decoder.C:48-52
__device__ __forceinline__ double angle_of(float2 cur, float2 prev) {
    const double yi = cur.x, yq = cur.y, pi = prev.x, pq = prev.y;
    const double re = __dadd_rn(__dmul_rn(yi, pi), __dmul_rn(yq, pq));
    const double im = __dsub_rn(__dmul_rn(yq, pi), __dmul_rn(yi, pq));
    return atan2(im, re);
}

// grid (tiles, channels), block kTile: |mask correlation| for samples [tile0, tile0 + kTile)
__global__ void __launch_bounds__(kTile) angle_corr_kernel(const DemodArgs a) {
    __shared__ double s_ang[kTile + 8];
    const int ch = blockIdx.y, tile0 = blockIdx.x * kTile, t = threadIdx.x;
    const float2* y = a.b.y3 + (size_t)ch * pitch_y(a.b.p_max) + kHistY;     // y[m], m >= -kHistY
    const int m = tile0 + t;
    if (m < a.n_new) s_ang[8 + t] = angle_of(y[m], y[m - 1]);
    if (t < 8) s_ang[t] = angle_of(y[tile0 - 8 + t], y[tile0 - 9 + t]);
    __syncthreads();
    if (m >= a.n_new) return;
    // mask {0,1,1,1,0,-1,-1,-1,0} over angles n-8 .. n, oldest first (decoder.C:161-170); s_ang[8 + t - k] = angle[m - k]
    double c = s_ang[t + 1];
    c = __dadd_rn(c, s_ang[t + 2]);
    c = __dadd_rn(c, s_ang[t + 3]);
    c = __dsub_rn(c, s_ang[t + 5]);
    c = __dsub_rn(c, s_ang[t + 6]);
    c = __dsub_rn(c, s_ang[t + 7]);
    a.b.corr[(size_t)ch * pitch_c(a.b.p_max) + kHistC + m] = fabs(c);
}

// mark/space decision of the window y[0..4] (decoder.C:109-133).  Returns true for 'Y'.
// Reference arithmetic per sample: acc = (float)((double)acc + ((double)((float)sR * f_a) +- sI * (double)f_b)).
// (float)sR is the stored float itself, (float)(-sR) * f = -((float)sR * f) exactly, and the float accumulators are
// carried as doubles holding float-representable values, so only the four roundings to float cost conversions.
__device__ __forceinline__ bool window_is_y(const float2* __restrict__ y, float* sums) {
    double br = 0.0, bi = 0.0, yr = 0.0, yi = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float srf = y[k].x;
        const double si = (double)y[k].y;
        const float fr = c_tone_r[k], fi = c_tone_i[k];
        const double pr = (double)__fmul_rn(srf, fr), pim = (double)__fmul_rn(srf, fi);
        const double qi = __dmul_rn(si, c_tone_id[k]), qr = __dmul_rn(si, c_tone_rd[k]);
        yr = (double)(float)__dadd_rn(yr, __dsub_rn(pr, qi));
        yi = (double)(float)__dadd_rn(yi, __dadd_rn(pim, qr));
        br = (double)(float)__dadd_rn(br, __dadd_rn(pr, qi));
        bi = (double)(float)__dadd_rn(bi, __dsub_rn(qr, pim));      // (-pim) + qr
    }
    const float brf = (float)br, bif = (float)bi, yrf = (float)yr, yif = (float)yi;
    if (sums) { sums[0] = brf; sums[1] = bif; sums[2] = yrf; sums[3] = yif; }
    const float eb = __fadd_rn(__fmul_rn(brf, brf), __fmul_rn(bif, bif));
    const float ey = __fadd_rn(__fmul_rn(yrf, yrf), __fmul_rn(yif, yif));
    return !(eb > ey);
}

// grid (tiles, channels), block kTile: one decision byte for samples w in [tile0 - 4, tile0 + kTile - 4)
__global__ void __launch_bounds__(kTile) sum_decide_kernel(const DemodArgs a) {
    constexpr int kBack = kCorrLen - 1 + 12;                 // corr history one tile needs before tile0
    __shared__ double s_corr[kBack + kTile];
    __shared__ double s_osum[kTile + 8];
    const int ch = blockIdx.y, tile0 = blockIdx.x * kTile, t = threadIdx.x;
    const double* corr = a.b.corr + (size_t)ch * pitch_c(a.b.p_max) + kHistC;      // corr[m], m >= -kHistC
    for (int k = t; k < kBack + kTile; k += kTile) {
        const int m = tile0 - kBack + k;
        s_corr[k] = (m >= -kHistC && m < a.n_new) ? corr[m] : 0.0;
    }
    __syncthreads();
    // per-offset sums for samples [tile0 - 12, tile0 + kTile - 4): thread t -> tile0 - 12 + t (and 8 more by t < 8)
    const int seen9 = (int)(a.seen % kSpb), seen567 = (int)(a.seen % kCorrLen);
    auto offset_sum = [&](int m) -> double {
        if (a.seen + m < kCorrLen + 7 || m >= a.n_new) return 0.0;
        // decoder.C:186-190: slots j, j+9, ... in ascending slot order, ring as it stood after this sample's write.
        // Slot i then held the value written at sample n - ((n - 8 - i) mod 567), n = seen + m.  In time order that
        // is: from the slot-j value forward to the newest one (q terms), then from the oldest one forward.
        // (n - 574) mod 9 and (n - 8 - j) mod 567 in 32-bit arithmetic; m >= -12.
        const int j = (seen9 + m + 27 - 7) % kSpb;                       // 574 = 63 * 9 + 7; m >= -12
        const int d0 = (seen567 + m + 2 * kCorrLen - 8 - j) % kCorrLen;   // multiple of 9 away from the newest slot
        const int q = d0 / kSpb + 1;
        const double* p = s_corr + (m - tile0 + kBack) - d0;             // slot-j value
        double acc = 0.0;
        int k = 0;
#pragma unroll 4
        for (; k < q; ++k) acc = __dadd_rn(acc, p[kSpb * k]);
        p -= kCorrLen;
#pragma unroll 4
        for (; k < 63; ++k) acc = __dadd_rn(acc, p[kSpb * k]);
        return acc;
    };
    s_osum[t] = offset_sum(tile0 - 12 + t);
    if (t < 8) s_osum[kTile + t] = offset_sum(tile0 - 12 + kTile + t);
    __syncthreads();
    const int w = tile0 - 4 + t;
    if (w >= a.n_new) return;
    unsigned out = 0;
    if (w >= 0 && a.seen + w >= kCorrLen + 15 && (seen9 + w) % kSpb == 6) {
        // first maximum of the nine sums, oldest first = offset index ascending (decoder.C:207-215); s_osum[t + 8 - k] = osum(w - k)
        double best = -1.0;
        int pick = 0;
#pragma unroll
        for (int i = 0; i < kSpb; ++i) {
            const double v = s_osum[t + i];
            if (v > best) { best = v; pick = i; }
        }
        out = (unsigned)pick;
    }
    if (w + 4 < a.n_new) {
        const float2* y = a.b.y3 + (size_t)ch * pitch_y(a.b.p_max) + kHistY + w;
        if (window_is_y(y, nullptr)) out |= 0x80u;
    }
    a.b.dec[(size_t)ch * pitch_d(a.b.p_max) + kHistD + w] = (uint8_t)out;
}

// lane = channel, warp = 32 channels.  The decision bytes are staged through shared memory in chunks
// (coalesced), then every lane runs its own symbol clock over the chunk, one iteration per BIT:
//   trigger t0 = first sample >= cur whose (n + 1) mod 9 equals the bit-sync offset (decoder.C:83-90);
//   the samples t0, t0+1, t0+2 are burnt, t0+3 .. t0+7 integrated, the bit is decided at t0+7
//   (decoder.C:91-135) = the precomputed decision of the window starting at t0+3;
//   every evaluation sample <= t0+7 has by then updated next_offs (bs_ runs before bd_, decoder.C:57-58).
constexpr int kChunk = 1152;                      // samples per staged chunk (multiple of 9 and 16)
constexpr int kChunkBack = 16;                    // bytes kept before the chunk (open windows, late evaluations)
constexpr int kRowPitch32 = (kChunk + kChunkBack) / 4 + 1;   // odd word pitch: conflict-free column reads

__global__ void __launch_bounds__(32) symbol_clock_kernel(const DemodArgs a) {
    __shared__ uint32_t s_dec[32 * kRowPitch32];
    __shared__ unsigned char s_ltrs[128], s_figs[128];
    const int lane = threadIdx.x;
    for (int k = lane; k < 128; k += 32) { s_ltrs[k] = c_ltrs[k]; s_figs[k] = c_figs[k]; }
    const int ch0 = blockIdx.x * 32, ch = ch0 + lane;
    const bool live = ch < a.channels;
    ChannelState s = {};
    if (live) s = a.b.state[ch];
    const float2* y = a.b.y3 + (size_t)(live ? ch : 0) * pitch_y(a.b.p_max) + kHistY;
    Emit em;
    em.ev = a.events + (size_t)(live ? ch : 0) * a.ev_cap; em.ev_cap = a.ev_cap; em.n = 0; em.writer = live;
    em.ltrs = s_ltrs; em.figs = s_figs;
    int nbits_out = 0;
    char* bits = a.bits && live ? a.bits + (size_t)ch * a.bit_cap : nullptr;
    float* disc = a.disc && live ? a.disc + (size_t)ch * a.bit_cap * 4 : nullptr;
    const uint8_t* my_row = reinterpret_cast<const uint8_t*>(s_dec + lane * kRowPitch32);

    // first evaluation sample of this block: absolute index >= 582 and == 6 (mod 9)  (decoder.C:204)
    int next_eval;
    {
        long long e = kCorrLen + 15 - a.seen;
        if (e < 0) e = 0;
        const int r = (int)((a.seen + e) % kSpb);
        e += (6 - r + kSpb) % kSpb;
        next_eval = e > a.n_new ? a.n_new : (int)e;
    }
    const int seen9 = (int)(a.seen % kSpb);

    for (int c0 = 0; c0 < a.n_new; c0 += kChunk) {
        const int lim = min(a.n_new, c0 + kChunk);
        // stage rows [c0 - 16, c0 + kChunk) of 32 channels
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            if (ch0 + r >= a.channels) break;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(a.b.dec + (size_t)(ch0 + r) * pitch_d(a.b.p_max) + kHistD + c0 - kChunkBack);
            const int words = (lim - c0 + kChunkBack + 3) / 4;
            for (int k = lane; k < words; k += 32) s_dec[r * kRowPitch32 + k] = src[k];
        }
        __syncwarp();
        auto byte_at = [&](int idx) -> unsigned { return my_row[idx - c0 + kChunkBack]; };
        auto do_eval = [&](int e) {
            int pick = (int)(byte_at(e) & 15u);
            if (s.last_pick != -1 && pick != s.last_pick) {        // slew one step the short way round (decoder.C:217-246)
                bool up;
                if (pick > s.last_pick) up = !(pick - s.last_pick > 4);
                else up = (s.last_pick - pick > 4);
                pick = up ? (s.last_pick + 1) % kSpb : (s.last_pick - 1 + kSpb) % kSpb;
            }
            s.last_pick = pick;
            const int offs = (pick + 5) % kSpb;                    // decoder.C:249
            if (s.dstate == DS_INIT) { s.dstate = DS_WAIT; s.offs = offs; s.cur = e; }   // decoder.C:62-70
            s.next_offs = offs;
        };
        while (live) {
            if (s.dstate == DS_INIT) {
                if (next_eval >= lim) break;
                do_eval(next_eval);
                next_eval += kSpb;
                if (s.dstate == DS_INIT) continue;
            }
            if (s.dstate == DS_WAIT) {
                const int tick = (seen9 + s.cur + 1) % kSpb;
                const int t0 = s.cur + (s.offs - tick + kSpb) % kSpb;
                if (t0 >= a.n_new) break;
                s.dstate = DS_PENDING;
                s.pend = t0;
            }
            const int td = s.pend + 7;
            if (td >= lim) break;
            while (next_eval <= td) { do_eval(next_eval); next_eval += kSpb; }
            const bool is_y = (byte_at(s.pend + 3) & 0x80u) != 0;
            if (bits && nbits_out < a.bit_cap) {
                bits[nbits_out] = is_y ? 'Y' : 'B';
                if (disc) window_is_y(y + s.pend + 3, disc + 4 * (size_t)nbits_out);
            }
            ++nbits_out;
            s.offs = s.next_offs;
            s.dstate = DS_WAIT;
            s.cur = td + 1;
            fsm_bit(s, em, is_y);
        }
        if (lim == a.n_new && live)
            while (next_eval < a.n_new) { do_eval(next_eval); next_eval += kSpb; }
    }
    if (!live) return;
    if (s.dstate == DS_PENDING) s.pend -= a.n_new;
    s.cur = s.cur > a.n_new ? s.cur - a.n_new : 0;
    s.seen = a.seen + a.n_new;
    a.b.state[ch] = s;
    a.ev_count[ch] = em.n;
    if (a.bit_count) a.bit_count[ch] = nbits_out;
}

// slide every history: the last H entries of [hist | new] become the next block's hist.  One CTA per channel.
__global__ void __launch_bounds__(256) carry_kernel(const DemodArgs a) {
    __shared__ double s_c[kHistC];
    __shared__ float2 s_y[kHistY];
    const int ch = blockIdx.x, t = threadIdx.x;
    double* corr = a.b.corr + (size_t)ch * pitch_c(a.b.p_max);
    const float2* y = a.b.y3 + (size_t)ch * pitch_y(a.b.p_max);
    float2* yn = a.y3_next + (size_t)ch * pitch_y(a.b.p_max);
    for (int k = t; k < kHistC; k += blockDim.x) s_c[k] = corr[k + a.n_new];
    if (t < kHistY) s_y[t] = y[t + a.n_new];
    __syncthreads();
    for (int k = t; k < kHistC; k += blockDim.x) corr[k] = s_c[k];
    if (t < kHistY) yn[t] = s_y[t];
}

__global__ void init_state_kernel(ChannelState* st, int channels) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= channels) return;
    ChannelState s = {};
    // decoder::decoder (decoder.C:6-39) and byte_state_machine::init (nav_b_sm.C:16-42)
    s.last_pick = -1;
    s.dstate = DS_INIT;
    s.byte_state = BY_WAIT;
    st[ch] = s;
}

}  // namespace

int demod_launches_per_block() { return 4; }
size_t demod_pitch_d(int p_max) { return ((size_t)kHistD + p_max + 15) & ~(size_t)15; }

cudaError_t demod_init_state(const DemodBuffers& b, int channels, cudaStream_t stream) {
    static bool tables_done = false;
    if (!tables_done) {
        // CCIR 476 tables as (code, letters, figures); see nav_b_sm.h:60-83 -- every other code is invalid '_'.
        static const struct { unsigned char code; char l, f; } codes[] = {
            {0x07, 'p', 'p'}, {0x0b, 'J', 'b'}, {0x0d, 'W', '2'}, {0x0e, 'A', '-'}, {0x13, 'F', '*'},
            {0x15, 'Y', '6'}, {0x16, 'S', '\''}, {0x19, '-', '-'}, {0x1a, 'D', '%'}, {0x1c, 'Z', '+'},
            {0x1d, '_', ' '}, {0x23, 'C', ':'}, {0x25, 'P', '0'}, {0x26, 'I', '8'}, {0x29, 'G', '*'},
            {0x2a, 'R', '4'}, {0x2c, 'L', ')'}, {0x31, 'M', '.'}, {0x32, 'N', ','}, {0x34, 'H', '*'},
            {0x38, 'O', '9'}, {0x43, 'K', '('}, {0x45, 'Q', '1'}, {0x46, 'U', '7'}, {0x49, 'f', 'f'},
            {0x4a, 'E', '3'}, {0x4c, 'q', 'q'}, {0x51, 'X', '/'}, {0x52, 'l', 'l'}, {0x58, 'B', '?'},
            {0x5c, ' ', ' '}, {0x61, 'V', '='}, {0x62, ' ', ' '}, {0x64, 'n', 'n'}, {0x68, 'T', '5'},
            {0x70, 'r', 'r'},
        };
        unsigned char ltrs[128], figs[128];
        for (int i = 0; i < 128; ++i) ltrs[i] = figs[i] = '_';
        for (const auto& c : codes) { ltrs[c.code] = (unsigned char)c.l; figs[c.code] = (unsigned char)c.f; }
        float tr[5], ti[5];
        for (int i = 0; i < 5; ++i) {
            const float ang = (float)((i * 2 * 3.1415 * 85) / 900);     // decoder.C:25: 3.1415, float angle
            tr[i] = cosf(ang);
            ti[i] = sinf(ang);
        }
        cudaError_t e;
        if ((e = cudaMemcpyToSymbol(c_ltrs, ltrs, sizeof ltrs)) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_figs, figs, sizeof figs)) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_tone_r, tr, sizeof tr)) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_tone_i, ti, sizeof ti)) != cudaSuccess) return e;
        double trd[5], tid[5];
        for (int i = 0; i < 5; ++i) { trd[i] = tr[i]; tid[i] = ti[i]; }
        if ((e = cudaMemcpyToSymbol(c_tone_rd, trd, sizeof trd)) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_tone_id, tid, sizeof tid)) != cudaSuccess) return e;
        tables_done = true;
    }
    cudaError_t e;
    if ((e = cudaMemsetAsync(b.y3, 0, sizeof(float2) * (size_t)channels * (kHistY + b.p_max), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(b.corr, 0, sizeof(double) * (size_t)channels * (kHistC + b.p_max), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(b.dec, 0, (size_t)channels * demod_pitch_d(b.p_max), stream)) != cudaSuccess) return e;
    init_state_kernel<<<(channels + 127) / 128, 128, 0, stream>>>(b.state, channels);
    return cudaGetLastError();
}

cudaError_t demod_launch(const DemodArgs& a, cudaStream_t stream) {
    if (a.n_new <= 0 || a.channels <= 0) return cudaSuccess;
    // These kernels run beside the NEXT block's cascade kernel.  NVX_DEMOD_THROTTLE=<bytes> asks for that much unused
    // dynamic shared memory, capping their occupancy in the ~48 KB the cascade leaves free (tuning knob; measured on
    // B200: throttling speeds the cascade up 10 % but stretches the demod past it, so the default is off).
    static int throttle = -1;
    if (throttle < 0) {
        const char* env = getenv("NVX_DEMOD_THROTTLE");
        throttle = env ? atoi(env) : 0;
    }
    const dim3 g1((a.n_new + kTile - 1) / kTile, a.channels);
    angle_corr_kernel<<<g1, kTile, throttle, stream>>>(a);
    const dim3 g2((a.n_new + 4 + kTile - 1) / kTile, a.channels);
    sum_decide_kernel<<<g2, kTile, throttle, stream>>>(a);
    symbol_clock_kernel<<<(a.channels + 31) / 32, 32, 0, stream>>>(a);
    carry_kernel<<<a.channels, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace nvx
