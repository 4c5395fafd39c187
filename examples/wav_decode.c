/* wav_decode.c -- a plain C host on the C ABI of include/navtex_b200.h: decode one NAVTEX capture.
 *
 * The reference's equivalent is capt_sched.c's consumer loop (receiver/capt_sched.c:484-528) fed from a file through
 * wav.c (wav_open / wav_read, receiver/wav.c) instead of the radio: read stereo s16 frames (I = left, Q = right,
 * 252 kHz, the format PrepWav writes at capt_sched.c:87-96) and hand every I,Q pair to the chain; decoded messages
 * arrive through an add_message-shaped callback (receiver/nav_b_sm.C:4).
 *
 *   cc -std=c11 -Iinclude examples/wav_decode.c -Lnavtex_b200 -lnavtex_b200 -Wl,-rpath,$PWD/navtex_b200 -o wav_decode
 *   ./wav_decode capture.wav        ->  one line "freq|bbbb|length" + the message text per decoded message
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "navtex_b200.h"

static int on_message(void *user, int stream, char *bbbb, char *message, int freq) {
    (void)stream;
    ++*(int *)user;
    printf("%d|%s|%zu\n%s\n", freq, bbbb, strlen(message), message);
    return 0;
}

/* minimal RIFF/WAVE reader: finds "fmt " and "data", accepts PCM 16 bit stereo only */
static int16_t *read_wav(const char *path, long long *frames, unsigned *rate) {
    FILE *f = fopen(path, "rb");
    unsigned char hdr[12], ck[8];
    int16_t *data = NULL;
    int have_fmt = 0;
    if (!f) return NULL;
    if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4)) { fclose(f); return NULL; }
    while (fread(ck, 1, 8, f) == 8) {
        const unsigned len = ck[4] | ck[5] << 8 | ck[6] << 16 | (unsigned)ck[7] << 24;
        if (!memcmp(ck, "fmt ", 4)) {
            unsigned char fmt[16];
            if (len < 16 || fread(fmt, 1, 16, f) != 16) break;
            if ((fmt[0] | fmt[1] << 8) != 1 || (fmt[2] | fmt[3] << 8) != 2 || (fmt[14] | fmt[15] << 8) != 16) break;
            *rate = fmt[4] | fmt[5] << 8 | fmt[6] << 16 | (unsigned)fmt[7] << 24;
            have_fmt = 1;
            fseek(f, (long)(len - 16 + (len & 1)), SEEK_CUR);
        } else if (!memcmp(ck, "data", 4) && have_fmt) {
            data = (int16_t *)malloc(len ? len : 1);
            if (data && fread(data, 1, len, f) == len) *frames = len / 4;
            else { free(data); data = NULL; }
            break;
        } else {
            fseek(f, (long)(len + (len & 1)), SEEK_CUR);
        }
    }
    fclose(f);
    return data;
}

int main(int argc, char **argv) {
    long long frames = 0, done = 0;
    unsigned rate = 0;
    int n_messages = 0;
    const long long block = 25200 * 10;      /* 1 s per push */
    nvx_config cfg;
    nvx_engine *eng;
    int16_t *iq;
    if (argc < 2) { fprintf(stderr, "usage: %s capture.wav\n", argv[0]); return 2; }
    iq = read_wav(argv[1], &frames, &rate);
    if (!iq) { fprintf(stderr, "%s: not a PCM s16 stereo WAV\n", argv[1]); return 2; }
    if (rate != NVX_FS_HZ) fprintf(stderr, "warning: sample rate %u, the chain expects %d\n", rate, NVX_FS_HZ);

    nvx_default_config(&cfg);                /* one stream, 518 kHz at +14 kHz, 490 kHz at -14 kHz */
    cfg.max_block = block;
    if (nvx_engine_create(&cfg, &eng)) { fprintf(stderr, "%s\n", nvx_last_error()); return 1; }   /* no GPU: fails here, no CPU path */
    nvx_engine_set_message_callback(eng, on_message, &n_messages);
    while (frames - done >= NVX_BLOCK_ALIGN) {
        long long n = frames - done < block ? frames - done : block;
        n -= n % NVX_BLOCK_ALIGN;
        if (nvx_engine_push_host_s16(eng, iq + 2 * done, n)) { fprintf(stderr, "%s\n", nvx_last_error()); return 1; }
        done += n;
    }
    if (nvx_engine_sync(eng) < 0) { fprintf(stderr, "%s\n", nvx_last_error()); return 1; }    /* callbacks run here */
    fprintf(stderr, "%lld of %lld frames decoded, %d message(s)\n", done, frames, n_messages);
    nvx_engine_destroy(eng);
    free(iq);
    return 0;
}
