/* fir2cpp.h -- same-named shim of the reference header receiver/fir2cpp.h:3-6 (init_fir_filter2, sample_in_2, fir_in_2, fir_in_2_490).
 * Put include/compat on the include path INSTEAD of the reference's receiver/ directory and the reference's own host
 * sources (nav_sched.C, capt_sched.c) compile unmodified against the GPU engine; link with -lnavtex_compat. */
#include "../navtex_compat.h"
