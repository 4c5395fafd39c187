// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// ref_harness.cpp: driver + stage taps around the UNMODIFIED reference
// translation units (fir1cpp.C fir2cpp.C fir3cpp.C decoder.C nav_b_sm.C
// nav_sched.C wav.c), which the Makefile next to this file compiles straight
// from /root/reference/receiver into oracle/_ref/.  Nothing from the reference
// is copied into this repository; this file only *calls* it.
//
// What it adds (none of which exists in the reference):
//   * the WAV/raw -> sample_in_1 feed loop (the reference only has the SDRplay
//     consumer loop, capt_sched.c:484-528, calling sample_in_1 at :511);
//   * an in-memory add_message collector (the reference's sink is SQLite,
//     message_store.c:59-97; the DSP path only needs the symbol, nav_b_sm.C:4);
//   * GNU ld --wrap interposers that record every stage boundary:
//       sample_in_2            (stage-1 output @63 kHz,  fir1cpp.C:129)
//       fir_filter3::sample_in (stage-2 output @9 kHz,   fir2cpp.C:164,208)
//       decoder::sample_in     (stage-3 output @900 Hz,  fir3cpp.C:54)
//       byte_state_machine::receive_bit (bit decisions,  decoder.C:127,131)
//
// Private members (freq tag, discriminator sums) are read through the usual
// "#define private public" trick, in this TU only; object layout is unchanged.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>

#define private public
#include "fir2cpp.h"   // pulls fir3cpp.h -> decoder.h -> nav_b_sm.h
#undef private
#include "fir1cpp.h"
#include "nav_sched.h"

extern "C" {
#include "wav_c_api.h"
}

namespace {
struct Chan {
    std::vector<double> y2, y3;       // interleaved I,Q
    std::vector<char> bits;
    std::vector<int32_t> bitpos;      // 900 Hz sample index (1-based count) at decision time
    std::vector<float> disc;          // BR,BI,YR,YI at decision time
    long n3 = 0;
};
Chan g_ch[2];                         // 0 = 518 kHz (+14 kHz), 1 = 490 kHz (-14 kHz)
std::vector<double> g_y1;
bool g_record = true;
const void *g_ff3[2] = {nullptr, nullptr};
const void *g_dec[2] = {nullptr, nullptr};
decoder *g_cur_dec = nullptr;
int g_cur_ch = 0;

struct Msg { int freq; std::string bbbb, text; };
std::vector<Msg> g_msgs;

int slot_of(const void **tab, const void *p) {
    if (tab[0] == p) return 0;
    if (tab[1] == p) return 1;
    if (!tab[0]) { tab[0] = p; return 0; }    // 518 path runs first inside sample_in_2 (fir2cpp.C:115-124)
    tab[1] = p; return 1;
}
}  // namespace

// ---- sink -------------------------------------------------------------
extern "C" int add_message(char *bbbb, char *message, int freq) {
    g_msgs.push_back(Msg{freq, std::string(bbbb), std::string(message)});
    return 0;
}

// ---- interposers (left out of the timing binary ref_chain_timing: -DNVX_NO_TAPS, linked without --wrap) -------------
#ifndef NVX_NO_TAPS
extern "C" {
void __real__Z11sample_in_2dd(double, double);
void __wrap__Z11sample_in_2dd(double i, double q) {
    if (g_record) { g_y1.push_back(i); g_y1.push_back(q); }
    __real__Z11sample_in_2dd(i, q);
}
void __real__ZN11fir_filter39sample_inEdd(void *, double, double);
void __wrap__ZN11fir_filter39sample_inEdd(void *self, double i, double q) {
    int c = slot_of(g_ff3, self);
    if (g_record) { g_ch[c].y2.push_back(i); g_ch[c].y2.push_back(q); }
    __real__ZN11fir_filter39sample_inEdd(self, i, q);
}
void __real__ZN7decoder9sample_inEdd(void *, double, double);
void __wrap__ZN7decoder9sample_inEdd(void *self, double i, double q) {
    int c = slot_of(g_dec, self);
    g_cur_dec = static_cast<decoder *>(self);
    g_cur_ch = c;
    g_ch[c].n3++;
    if (g_record) { g_ch[c].y3.push_back(i); g_ch[c].y3.push_back(q); }
    __real__ZN7decoder9sample_inEdd(self, i, q);
}
void __real__ZN18byte_state_machine11receive_bitEc(void *, char);
void __wrap__ZN18byte_state_machine11receive_bitEc(void *self, char b) {
    int c = static_cast<byte_state_machine *>(self)->freq == 518 ? 0 : 1;
    if (g_record) {
        g_ch[c].bits.push_back(b);
        g_ch[c].bitpos.push_back((int32_t)g_ch[c].n3);
        const decoder *d = g_cur_dec;
        g_ch[c].disc.push_back(d->Brotated_samplesumR);
        g_ch[c].disc.push_back(d->Brotated_samplesumI);
        g_ch[c].disc.push_back(d->Yrotated_samplesumR);
        g_ch[c].disc.push_back(d->Yrotated_samplesumI);
    }
    __real__ZN18byte_state_machine11receive_bitEc(self, b);
}
}
#endif

namespace {
template <class T>
void dump(const std::string &path, const std::vector<T> &v) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

std::vector<double> load_input(const char *kind, const char *path) {
    std::vector<double> iq;
    if (!strcmp(kind, "--wav")) {
        // the wav.c reader half (wav.c:469, :494-528); stereo s16, ch0 = I, ch1 = Q
        WavFile *w = wav_open(path, WAV_OPEN_READ);
        if (!w || wav_get_num_channels(w) != 2 || wav_get_sample_size(w) != 2) {
            fprintf(stderr, "ref_chain: %s is not a stereo 16-bit PCM WAV\n", path);
            exit(2);
        }
        std::vector<int16_t> buf(2 * 65536);
        size_t got;
        while ((got = wav_read(w, buf.data(), 65536)) > 0)
            for (size_t k = 0; k < 2 * got; ++k) iq.push_back((double)buf[k]);
        wav_close(w);
        return iq;
    }
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (!strcmp(kind, "--s16")) {
        std::vector<int16_t> b(bytes / 2);
        if (fread(b.data(), 2, b.size(), f) != b.size()) exit(2);
        iq.assign(b.begin(), b.end());
    } else if (!strcmp(kind, "--f32")) {
        std::vector<float> b(bytes / 4);
        if (fread(b.data(), 4, b.size(), f) != b.size()) exit(2);
        iq.assign(b.begin(), b.end());
    } else {
        fprintf(stderr, "ref_chain: unknown input kind %s\n", kind);
        exit(2);
    }
    fclose(f);
    return iq;
}
}  // namespace

int main(int argc, char **argv) {
    const char *kind = nullptr, *in = nullptr, *out = nullptr;
    int passes = 1;
    bool verbose = false;
    for (int a = 1; a < argc; ++a) {
        if ((!strcmp(argv[a], "--wav") || !strcmp(argv[a], "--s16") || !strcmp(argv[a], "--f32")) && a + 1 < argc) {
            kind = argv[a]; in = argv[++a];
        } else if (!strcmp(argv[a], "--out") && a + 1 < argc) out = argv[++a];
        else if (!strcmp(argv[a], "--passes") && a + 1 < argc) passes = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--verbose")) verbose = true;
        else { fprintf(stderr, "usage: ref_chain (--wav|--s16|--f32) FILE [--out PREFIX] [--passes N] [--verbose]\n"); return 2; }
    }
    if (!in) { fprintf(stderr, "ref_chain: no input\n"); return 2; }
    // the reference chats on stdout (nav_b_sm.C:46,66,106,...); keep ours on stderr
    if (!verbose && !freopen("/dev/null", "w", stdout)) return 2;

    std::vector<double> iq = load_input(kind, in);
    const size_t n = iq.size() / 2;

    init_fir_filter1();      // capt_sched.c:554
    init_fir2_wrapper();     // capt_sched.c:612

    double best = 1e30, total = 0;
    std::vector<double> pass_s;
    for (int p = 0; p < passes; ++p) {
        g_record = (p == 0) && out != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        for (size_t k = 0; k < n; ++k) sample_in_1(iq[2 * k], iq[2 * k + 1]);   // capt_sched.c:511
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
        total += dt;
        pass_s.push_back(dt);
    }
    fprintf(stderr, "{\"samples\": %zu, \"passes\": %d, \"best_s\": %.6f, \"mean_s\": %.6f, \"msps_best\": %.3f, \"msps_mean\": %.3f, \"messages\": %zu, \"pass_s\": [",
            n, passes, best, total / passes, n / best * 1e-6, n * passes / total * 1e-6, g_msgs.size());
    for (size_t p = 0; p < pass_s.size(); ++p) fprintf(stderr, "%s%.6f", p ? ", " : "", pass_s[p]);
    fprintf(stderr, "]}\n");

    if (out) {
        std::string o(out);
        dump(o + ".y1", g_y1);
        const char *tag[2] = {"518", "490"};
        for (int c = 0; c < 2; ++c) {
            dump(o + ".y2_" + tag[c], g_ch[c].y2);
            dump(o + ".y3_" + tag[c], g_ch[c].y3);
            dump(o + ".bits_" + tag[c], g_ch[c].bits);
            dump(o + ".bitpos_" + tag[c], g_ch[c].bitpos);
            dump(o + ".disc_" + tag[c], g_ch[c].disc);
        }
        FILE *f = fopen((o + ".msgs").c_str(), "wb");
        for (const Msg &m : g_msgs) {
            fprintf(f, "%d|%s|%zu\n", m.freq, m.bbbb.c_str(), m.text.size());
            fwrite(m.text.data(), 1, m.text.size(), f);
            fputc('\n', f);
        }
        fclose(f);
    }
    return 0;
}
