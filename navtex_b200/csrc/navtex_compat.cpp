// navtex_compat.cpp -- one-stream adapter that keeps the reference's link-level names alive on top of
// the batched engine (see include/navtex_compat.h).
#include "../../include/navtex_compat.h"

#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../include/navtex_b200.h"

extern "C" int add_message(char* bbbb, char* message, int freq) __attribute__((weak));

namespace {
nvx_engine* g_engine = nullptr;
std::vector<float> g_buf;          // interleaved I,Q of the block being filled
navtex_sink_fn g_sink = nullptr;
int g_device = 0;
int g_tag[2] = {518, 490};

int deliver(void*, int, char* bbbb, char* message, int freq) {
    if (g_sink) return g_sink(bbbb, message, freq);
    if (add_message) return add_message(bbbb, message, freq);
    fprintf(stderr, "navtex_compat: message %s on %d dropped: host defines no add_message and set no sink\n", bbbb, freq);
    return -1;
}

void die(const char* what) {
    fprintf(stderr, "navtex_compat: %s: %s\n", what, nvx_last_error());
    abort();      // the reference's entry points return void and cannot fail; there is no CPU path to fall back to
}

void ensure_engine() {
    if (g_engine) return;
    nvx_config cfg;
    nvx_default_config(&cfg);
    cfg.device = g_device;
    cfg.n_streams = 1;
    cfg.max_block = NAVTEX_COMPAT_BLOCK;
    cfg.freq_tag[0] = g_tag[0];
    cfg.freq_tag[1] = g_tag[1];
    if (nvx_engine_create(&cfg, &g_engine) != 0) die("cannot create the GPU engine");
    nvx_engine_set_message_callback(g_engine, deliver, nullptr);
    g_buf.reserve(2 * NAVTEX_COMPAT_BLOCK);
}

void push(size_t samples) {
    if (nvx_engine_push_host_f32(g_engine, g_buf.data(), (long long)samples) != 0) die("push failed");
    g_buf.erase(g_buf.begin(), g_buf.begin() + 2 * (long)samples);
}

[[noreturn]] void no_stage_push(const char* name) {
    fprintf(stderr, "navtex_compat: %s: per-stage pushes do not exist on the GPU path (stages are fused); feed sample_in_1\n", name);
    abort();
}
}  // namespace

extern "C" {

void navtex_compat_set_sink(navtex_sink_fn fn) { g_sink = fn; }
void navtex_compat_set_device(int device) { g_device = device; }

// fir1cpp.C:65-77 resets stage 1 only; the fused engine resets the whole chain, which is what
// capt_sched.c:552-555,:612 obtain by calling both initialisers back to back at start-up.
void init_fir_filter1(void) {
    ensure_engine();
    g_buf.clear();
    if (nvx_engine_reset(g_engine) != 0) die("reset failed");
}

void init_fir2_wrapper(void) { ensure_engine(); }

void sample_in_1(double sample_I, double sample_Q) {
    ensure_engine();
    g_buf.push_back((float)sample_I);      // (double)short at capt_sched.c:511: exact in float
    g_buf.push_back((float)sample_Q);
    if (g_buf.size() == 2 * (size_t)NAVTEX_COMPAT_BLOCK) {
        push(NAVTEX_COMPAT_BLOCK);
        const int rc = nvx_engine_sync(g_engine);
        if (rc < 0 && rc != NVX_ERR_OVERFLOW) die("sync failed");
    }
}

int navtex_compat_flush(void) {
    if (!g_engine) return 0;
    const size_t whole = (g_buf.size() / 2 / NVX_BLOCK_ALIGN) * NVX_BLOCK_ALIGN;
    if (whole) push(whole);
    return nvx_engine_sync(g_engine);
}

void navtex_compat_shutdown(void) {
    if (g_engine) nvx_engine_destroy(g_engine);
    g_engine = nullptr;
    g_buf.clear();
}

}  // extern "C"

byte_state_machine::byte_state_machine(unsigned int frequency) : freq(frequency) {}
void byte_state_machine::receive_bit(char) { no_stage_push("byte_state_machine::receive_bit"); }
decoder::decoder(byte_state_machine* bsm) : output_bsm(bsm) {}
void decoder::sample_in(double, double) { no_stage_push("decoder::sample_in"); }
fir_filter3::fir_filter3(decoder* dec) : output_dec(dec) {}
void fir_filter3::sample_in(double, double) { no_stage_push("fir_filter3::sample_in"); }
// fir2cpp.C:90-110: remember which state machines sit behind the 518 / 490 branches (their tags label messages)
void init_fir_filter2(fir_filter3* ff3_518, fir_filter3* ff3_490) {
    int tag[2] = {g_tag[0], g_tag[1]};
    if (ff3_518 && ff3_518->output_dec && ff3_518->output_dec->output_bsm) tag[0] = (int)ff3_518->output_dec->output_bsm->freq;
    if (ff3_490 && ff3_490->output_dec && ff3_490->output_dec->output_bsm) tag[1] = (int)ff3_490->output_dec->output_bsm->freq;
    if (tag[0] == g_tag[0] && tag[1] == g_tag[1]) return;
    g_tag[0] = tag[0]; g_tag[1] = tag[1];
    // capt_sched.c calls init_fir_filter1 (:554) before init_fir2_wrapper (:612): an engine created with other tags is
    // replaced (nothing has been pushed yet at that point of the reference's start-up)
    if (g_engine) { nvx_engine_destroy(g_engine); g_engine = nullptr; g_buf.clear(); ensure_engine(); }
}
void sample_in_2(double, double) { no_stage_push("sample_in_2"); }
void fir_in_2(double, double) { no_stage_push("fir_in_2"); }
void fir_in_2_490(double, double) { no_stage_push("fir_in_2_490"); }
