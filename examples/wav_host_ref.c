/* wav_host_ref.c -- the WAV -> sample_in_1 driver the north star names ("fed through the wav.c input path").  The reference
 * has no such driver (its only consumer loop reads the SDRplay ring, capt_sched.c:484-528, and wav.c is only ever used to
 * write a debug capture, capt_sched.c:87-101), so this is the ~15 lines a maintainer would add.  It uses nothing but the
 * reference's own interfaces: wav_open / wav_read / wav_close (receiver/wav.h:129-219, wav.c:469, :494-528),
 * init_fir_filter1 / sample_in_1 (fir1cpp.h:2-3), init_fir2_wrapper (nav_sched.h:1) and the add_message sink
 * (nav_b_sm.C:4, message_store.h:7).
 *
 * oracle/Makefile (target ref_host) compiles it together with the reference's UNMODIFIED receiver/nav_sched.C (against
 * include/compat/) and receiver/wav.c, and links the three against libnavtex_compat.so -- the GPU engine -- instead of
 * fir1cpp.o fir2cpp.o fir3cpp.o decoder.o nav_b_sm.o.  Stereo 16-bit PCM at 252 kS/s, left = I, right = Q (capt_sched.c:87-96).
 */
#include <stdio.h>
#include <string.h>

#include "wav.h"          /* the reference's receiver/wav.h */

void init_fir_filter1(void);                         /* fir1cpp.h:2, as capt_sched.c:17-18 declares them */
void sample_in_1(double sample_I, double sample_Q);
void init_fir2_wrapper(void);                        /* nav_sched.h:1 -- defined by the reference's nav_sched.C */
int navtex_compat_flush(void);                       /* GPU adapter: push what is left of the last 0.1 s block */

int add_message(char *bbbb, char *message, int freq) {   /* message_store.c:59's symbol, defined by the host */
    printf("%d|%s|%zu\n%s\n", freq, bbbb, strlen(message), message);
    return 0;
}

int main(int argc, char **argv) {
    static short buf[2 * 65536];
    size_t got, k;
    WavFile *w;
    if (argc < 2) { fprintf(stderr, "usage: %s capture.wav\n", argv[0]); return 2; }
    w = wav_open(argv[1], WAV_OPEN_READ);
    if (!w || wav_get_num_channels(w) != 2 || wav_get_sample_size(w) != 2) { fprintf(stderr, "%s: not a stereo 16-bit PCM WAV\n", argv[1]); return 2; }
    init_fir_filter1();                              /* capt_sched.c:554 */
    init_fir2_wrapper();                             /* capt_sched.c:612 */
    while ((got = wav_read(w, buf, 65536)) > 0)
        for (k = 0; k < got; ++k) sample_in_1((double)buf[2 * k], (double)buf[2 * k + 1]);   /* capt_sched.c:511 */
    wav_close(w);
    return navtex_compat_flush() < 0 ? 1 : 0;
}
