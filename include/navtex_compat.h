/* navtex_compat.h -- the reference's own link-level seam, served by the GPU engine.
 *
 * A host written against the reference headers (capt_sched.c:17-18, :511, :554, :612; nav_sched.C)
 * keeps calling exactly these names and links against libnavtex_compat.so instead of
 * fir1cpp.o / fir2cpp.o / fir3cpp.o / decoder.o / nav_b_sm.o (and, if it likes, nav_sched.o: the reference's own
 * nav_sched.C also compiles UNMODIFIED against the same-named shim headers in include/compat/ and links on top):
 *
 *   extern "C" void init_fir_filter1();                    receiver/fir1cpp.h:2
 *   extern "C" void sample_in_1(double I, double Q);       receiver/fir1cpp.h:3
 *   extern "C" void init_fir2_wrapper();                   receiver/nav_sched.h:1
 *   extern "C" int  add_message(char*, char*, int);        receiver/nav_b_sm.C:4 (defined BY THE HOST)
 *
 * sample_in_1 buffers one stream into blocks of NAVTEX_COMPAT_BLOCK samples (0.1 s) and pushes each
 * full block through a one-stream nvx_engine; decoded messages come back through the host's
 * add_message (resolved as a weak symbol at load time, or set with navtex_compat_set_sink), strings
 * borrowed for the duration of the call exactly as in the reference.  Results are therefore
 * delivered up to 0.1 s later than the per-sample CPU chain would have; navtex_compat_flush()
 * forces delivery of everything that is a whole multiple of 280 samples.
 */
#ifndef NAVTEX_COMPAT_H
#define NAVTEX_COMPAT_H
#ifdef __cplusplus
extern "C" {
#endif

#define NAVTEX_COMPAT_BLOCK 25200

void init_fir_filter1(void);
void sample_in_1(double sample_I, double sample_Q);
void init_fir2_wrapper(void);

typedef int (*navtex_sink_fn)(char *bbbb, char *message, int freq);
void navtex_compat_set_sink(navtex_sink_fn fn);   /* overrides the weak add_message lookup */
void navtex_compat_set_device(int device);        /* before init_fir_filter1; default 0 */
int navtex_compat_flush(void);                    /* 0 or a negative nvx_status */
void navtex_compat_shutdown(void);

#ifdef __cplusplus
}

/* Wiring-only counterparts of the reference classes so nav_sched.C-style code compiles unchanged
 * (receiver/nav_b_sm.h:126-127, decoder.h:84-85, fir3cpp.h:98-100, fir2cpp.h:3-6).  On the GPU path
 * the stages are fused into two kernels, so there is no per-stage push: constructing and wiring the
 * objects is supported (the frequency tags are honoured), calling the per-sample members is not and
 * terminates with a diagnostic instead of silently running a CPU path. */
class byte_state_machine {
  public:
    explicit byte_state_machine(unsigned int frequency);
    void receive_bit(char bit_received);
    unsigned int freq;
};
class decoder {
  public:
    explicit decoder(byte_state_machine *bsm);
    void sample_in(double sampleI, double sampleQ);
    byte_state_machine *output_bsm;
};
class fir_filter3 {
  public:
    explicit fir_filter3(decoder *dec);
    void sample_in(double sample_I, double sample_Q);
    decoder *output_dec;
};
void init_fir_filter2(fir_filter3 *ff3_518, fir_filter3 *ff3_490);   /* fir2cpp.h:3 */
void sample_in_2(double sample_I, double sample_Q);                    /* fir2cpp.h:4: aborts (no per-stage push) */
void fir_in_2(double sample_I, double sample_Q);                       /* fir2cpp.h:5: aborts */
void fir_in_2_490(double sample_I, double sample_Q);                   /* fir2cpp.h:6: aborts */
#endif
#endif
