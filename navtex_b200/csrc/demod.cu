// demod.cu -- FSK discriminator, bit synchroniser, mark/space decision and SITOR-B byte state
// machine for [channels] independent 900 Hz complex streams, one WARP per channel.
//
// Replaces, per channel and per batch instead of per sample,
//   decoder::sample_in            receiver/decoder.C:42-59    (delay-conjugate product + atan2)
//   decoder::bs_decoded_sample_in receiver/decoder.C:142-255  (bit-sync timing recovery)
//   decoder::bd_decoded_sample_in receiver/decoder.C:73-137   (mark/space energy discriminator)
//   byte_state_machine::receive_bit / receive_rxdx_byte / message_byte_out
//                                 receiver/nav_b_sm.C:266-634, :150-262, :100-145
// Line assembly, the ZCZC/NNNN regexes and add_message (nav_b_sm.C:56-97, :44-52) stay on the
// host (message_assembler.cpp); this kernel emits the character/line/abort event stream they
// consume.
//
// Work split inside the warp, per batch of 32 consecutive 900 Hz samples (lane = sample):
//   feed-forward part, all lanes in parallel: angle (FP64 atan2), 9-tap transition-mask
//   correlation, and the 63-term per-offset sum, each in exactly the reference's operation order
//   (the 567-deep ring is kept as a time-indexed history so every lane sees the ring "as of" its
//   own sample);
//   feedback part, 32 short uniform iterations: lanes 0..8 own the nine offset sums, the argmax is
//   a shuffle reduction, then slew limiting, the WAIT/BIT_START/RECEIVING machine with the
//   reference's mixed float/double accumulator arithmetic (explicit _rn intrinsics, no FMA
//   contraction), and the byte state machine, evaluated redundantly by all lanes (no divergence).
#include "demod.cuh"

#include <math.h>

namespace nvx {

namespace {

constexpr int kWarps = 4;

enum { DS_INIT = 0, DS_WAIT = 1, DS_BIT_START = 2, DS_RECEIVING = 3 };   // decoder.h:16-19
enum { BY_WAIT = 1, BY_GOT_DX = 2, BY_GOT_RX = 3 };                       // nav_b_sm.h:41-43

__constant__ unsigned char c_ltrs[128];
__constant__ unsigned char c_figs[128];
__constant__ float c_tone_r[5];
__constant__ float c_tone_i[5];

struct Emit {
    uint8_t* ev; int* ev_n; int ev_cap; int n;
    bool writer;
    __device__ __forceinline__ void put(int c) {
        if (writer && n < ev_cap) ev[n] = (uint8_t)c;
        ++n;
    }
};

// byte_state_machine::init, nav_b_sm.C:16-42 (the error ring contents survive, only its counters reset)
__device__ __forceinline__ void fsm_reset(ChannelScalars& s) {
    s.match = 0; s.byte_state = BY_WAIT; s.figures = 0; s.nbits = 0;
    s.dx_at = 0; s.dx_full = 0;
    s.err_count = 0; s.err_at = 0; s.err_full = 0;
    s.eoe_count = 0; s.prev_dx_alpha = 0;
    s.holdoff = 0; s.enabled = 0;
}

// message_byte_out, nav_b_sm.C:100-145; code 0 = "no valid copy" -> '*'
__device__ __forceinline__ void fsm_char(ChannelScalars& s, Emit& e, int code) {
    if (code == 0) { e.put('*'); return; }
    const int l = c_ltrs[code];
    if (l == 'l') { s.figures = 0; return; }
    if (l == 'f') { s.figures = 1; return; }
    if (l == 'n') { e.put('\n'); return; }
    if (l == 'r' || l == 'p' || l == 'q') return;
    e.put(s.figures ? c_figs[code] : l);
}

// message_abort, nav_b_sm.C:44-52: the host decides whether a message was in progress
__device__ __forceinline__ void fsm_abort(ChannelScalars& s, Emit& e) {
    e.put(kEvAbort);
    fsm_reset(s);
}

// receive_rxdx_byte, nav_b_sm.C:150-262
__device__ __forceinline__ void fsm_byte(ChannelScalars& s, Emit& e, int b) {
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
            if (c_ltrs[b] != '_') fsm_char(s, e, b);
            else if (c_ltrs[dx] != '_') fsm_char(s, e, dx);
            else fsm_char(s, e, 0);
        }
        s.byte_state = BY_GOT_RX;
    }
    // 20-byte sliding window of invalid codes (nav_b_sm.C:235-261)
    const unsigned bit = 1u << s.err_at;
    if (s.err_full && (s.err_mask & bit)) s.err_count--;
    const bool bad = c_ltrs[b] == '_';
    s.err_mask = bad ? (s.err_mask | bit) : (s.err_mask & ~bit);
    if (bad) s.err_count++;
    if (++s.err_at == 20) { s.err_at = 0; s.err_full = 1; }
    if (s.err_count > 12) {
        fsm_char(s, e, 0);
        fsm_abort(s, e);
    }
}

// receive_bit, nav_b_sm.C:266-634.  is_y: 'Y' (=1) else 'B'.
__device__ __forceinline__ void fsm_bit(ChannelScalars& s, Emit& e, bool is_y) {
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

__global__ void __launch_bounds__(kWarps * 32) demod_kernel(const DemodArgs a) {
    __shared__ double s_corr[kWarps][kCorrRing];
    __shared__ double s_ang[kWarps][40];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kWarps + warp;
    if (ch >= a.channels) return;
    constexpr unsigned kAll = 0xffffffffu;

    ChannelState& g = a.state[ch];
    double* corr = s_corr[warp];
    double* ang = s_ang[warp];
    for (int k = lane; k < kCorrRing; k += 32) corr[k] = g.corr[k];
    if (lane < 8) ang[lane] = g.ang[lane];
    double my_osum = lane < kSpb ? g.osum[lane] : -2.0;

    // scalar state, held redundantly (and identically) by every lane
    ChannelScalars s = g.sc;
    __syncwarp();

    Emit em;
    em.ev = a.events + (size_t)ch * a.ev_cap; em.ev_cap = a.ev_cap; em.n = 0; em.writer = lane == 0;
    int nbits_out = 0;
    char* bits = a.bits ? a.bits + (size_t)ch * a.bit_cap : nullptr;
    float* disc = a.disc ? a.disc + (size_t)ch * a.bit_cap * 4 : nullptr;

    const float2* y3 = a.y3 + (size_t)ch * a.y3_pitch + a.y3_off;
    int m9 = (int)(s.seen % kSpb);

    for (int base = 0; base < a.n_new; base += 32) {
        const int cnt = min(32, a.n_new - base);
        const long long n = s.seen + lane;                 // absolute index of this lane's sample
        const bool active = lane < cnt;
        const float2 y = active ? y3[base + lane] : make_float2(0.f, 0.f);
        const double yi = (double)y.x, yq = (double)y.y;

        // ---- feed-forward, lane = sample --------------------------------------------------
        double pi = __shfl_up_sync(kAll, yi, 1), pq = __shfl_up_sync(kAll, yq, 1);
        if (lane == 0) { pi = s.prev_i; pq = s.prev_q; }
        const double re = __dadd_rn(__dmul_rn(yi, pi), __dmul_rn(yq, pq));       // decoder.C:48
        const double im = __dsub_rn(__dmul_rn(yq, pi), __dmul_rn(yi, pq));       // decoder.C:49
        const double angle = atan2(im, re);                                       // decoder.C:52
        ang[8 + lane] = angle;
        __syncwarp();
        if (active && n >= 8) {
            // mask {0,1,1,1,0,-1,-1,-1,0} over angles n-8 .. n, oldest first (decoder.C:161-170)
            double t = ang[lane + 1];
            t = __dadd_rn(t, ang[lane + 2]);
            t = __dadd_rn(t, ang[lane + 3]);
            t = __dsub_rn(t, ang[lane + 5]);
            t = __dsub_rn(t, ang[lane + 6]);
            t = __dsub_rn(t, ang[lane + 7]);
            corr[(int)((n - 8) & (kCorrRing - 1))] = fabs(t);
        }
        __syncwarp();
        double osum_new = 0.0;
        if (active && n >= kCorrLen + 7) {
            // decoder.C:186-190: sum ring slots j, j+9, ... in ascending slot order, as the ring stood
            // after this sample's write.  Slot i then held value number v - ((v - i) mod 567), v = n - 8.
            const long long v = n - 8;
            const int j = (int)((v - (kCorrLen - 1)) % kSpb);
            int d = (int)((v - j) % kCorrLen);
            double acc = 0.0;
#pragma unroll 9
            for (int k = 0; k < 63; ++k) {
                acc = __dadd_rn(acc, corr[(int)((v - d) & (kCorrRing - 1))]);
                d -= kSpb;
                if (d < 0) d += kCorrLen;
            }
            osum_new = acc;
        }
        __syncwarp();
        {   // slide the angle history and the previous-sample carry to the end of this batch
            const double keep = lane < 8 ? ang[cnt + lane] : 0.0;
            __syncwarp();
            if (lane < 8) ang[lane] = keep;
            s.prev_i = __shfl_sync(kAll, yi, cnt - 1);
            s.prev_q = __shfl_sync(kAll, yq, cnt - 1);
        }

        // ---- feedback, 32 uniform iterations -----------------------------------------------
        for (int k = 0; k < cnt; ++k) {
            const long long nk = s.seen + k;
            const double ov = __shfl_sync(kAll, osum_new, k);
            const double sr = __shfl_sync(kAll, yi, k), si = __shfl_sync(kAll, yq, k);
            if (nk >= kCorrLen + 7) {
                int j = m9 - 7; if (j < 0) j += kSpb;             // (nk - 574) mod 9
                if (lane == j) my_osum = ov;
                if (nk >= kCorrLen + 15 && m9 == 6) {             // every 9th sample from 582 on (decoder.C:204)
                    double bv = my_osum; int bi = lane;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const double ovv = __shfl_xor_sync(kAll, bv, off);
                        const int oi = __shfl_xor_sync(kAll, bi, off);
                        if (ovv > bv || (ovv == bv && oi < bi)) { bv = ovv; bi = oi; }
                    }
                    int pick = bi;                                 // first maximum, strict '>' (decoder.C:207-215)
                    if (s.last_pick != -1 && pick != s.last_pick) {
                        bool up;
                        if (pick > s.last_pick) up = !(pick - s.last_pick > 4);
                        else up = (s.last_pick - pick > 4);
                        pick = up ? (s.last_pick + 1) % kSpb : (s.last_pick - 1 + kSpb) % kSpb;
                    }
                    s.last_pick = pick;
                    const int offs = (pick + 5) % kSpb;            // decoder.C:249
                    if (s.dstate == DS_INIT) { s.dstate = DS_WAIT; s.offs = offs; }
                    s.next_offs = offs;
                }
            }
            // mark/space discriminator (decoder.C:73-137); bd_seq_nbr % 9 == (nk + 1) % 9
            const int tick = m9 + 1 == kSpb ? 0 : m9 + 1;
            if (s.dstate != DS_INIT) {
                if (s.dstate == DS_WAIT && tick == s.offs) { s.dstate = DS_BIT_START; s.burned = 0; }
                if (s.dstate == DS_BIT_START) {
                    if (s.burned == 2) {
                        s.dstate = DS_RECEIVING; s.used = 0;
                        s.br = s.bi = s.yr = s.yi = 0.f;
                    } else {
                        s.burned++;
                    }
                } else if (s.dstate == DS_RECEIVING) {
                    const float fr = c_tone_r[s.used], fi = c_tone_i[s.used];
                    const float srf = (float)sr, nsrf = (float)(-sr);
                    const double pr = (double)__fmul_rn(srf, fr), pim = (double)__fmul_rn(srf, fi);
                    const double npim = (double)__fmul_rn(nsrf, fi);
                    const double qi = __dmul_rn(si, (double)fi), qr = __dmul_rn(si, (double)fr);
                    s.yr = (float)__dadd_rn((double)s.yr, __dsub_rn(pr, qi));
                    s.yi = (float)__dadd_rn((double)s.yi, __dadd_rn(pim, qr));
                    s.br = (float)__dadd_rn((double)s.br, __dadd_rn(pr, qi));
                    s.bi = (float)__dadd_rn((double)s.bi, __dadd_rn(npim, qr));
                    if (++s.used == 5) {
                        const float eb = __fadd_rn(__fmul_rn(s.br, s.br), __fmul_rn(s.bi, s.bi));
                        const float ey = __fadd_rn(__fmul_rn(s.yr, s.yr), __fmul_rn(s.yi, s.yi));
                        const bool is_y = !(eb > ey);
                        if (bits && lane == 0 && nbits_out < a.bit_cap) {
                            bits[nbits_out] = is_y ? 'Y' : 'B';
                            if (disc) {
                                float* dd = disc + 4 * (size_t)nbits_out;
                                dd[0] = s.br; dd[1] = s.bi; dd[2] = s.yr; dd[3] = s.yi;
                            }
                        }
                        ++nbits_out;
                        s.dstate = DS_WAIT;
                        s.offs = s.next_offs;
                        fsm_bit(s, em, is_y);
                    }
                }
            }
            m9 = tick;
        }
        s.seen += cnt;
        __syncwarp();
    }

    // write back
    for (int k = lane; k < kCorrRing; k += 32) g.corr[k] = corr[k];
    if (lane < 8) g.ang[lane] = ang[lane];
    if (lane < kSpb) g.osum[lane] = my_osum;
    if (lane == 0) {
        g.sc = s;
        a.ev_count[ch] = em.n;
        if (a.bit_count) a.bit_count[ch] = nbits_out;
    }
}

__global__ void demod_init_kernel(ChannelState* st, int channels) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= channels) return;
    ChannelScalars& s = st[ch].sc;
    // decoder::decoder (decoder.C:6-39) and byte_state_machine::init (nav_b_sm.C:16-42); rings are zero
    s.last_pick = -1;
    s.dstate = DS_INIT;
    s.byte_state = BY_WAIT;
}

}  // namespace

cudaError_t demod_init_state(ChannelState* state, int channels, cudaStream_t stream) {
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
        tables_done = true;
    }
    cudaError_t e = cudaMemsetAsync(state, 0, sizeof(ChannelState) * (size_t)channels, stream);
    if (e != cudaSuccess) return e;
    demod_init_kernel<<<(channels + 127) / 128, 128, 0, stream>>>(state, channels);
    return cudaGetLastError();
}

cudaError_t demod_launch(const DemodArgs& a, cudaStream_t stream) {
    if (a.n_new <= 0 || a.channels <= 0) return cudaSuccess;
    demod_kernel<<<(a.channels + kWarps - 1) / kWarps, kWarps * 32, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace nvx
