/* fir1cpp.h -- same-named shim of the reference header receiver/fir1cpp.h:2-3 (init_fir_filter1, sample_in_1).
 * Put include/compat on the include path INSTEAD of the reference's receiver/ directory and the reference's own host
 * sources (nav_sched.C, capt_sched.c) compile unmodified against the GPU engine; link with -lnavtex_compat. */
#include "../navtex_compat.h"
