/* nav_sched.h -- same-named shim of the reference header receiver/nav_sched.h:1 (init_fir2_wrapper).
 * Put include/compat on the include path INSTEAD of the reference's receiver/ directory and the reference's own host
 * sources (nav_sched.C, capt_sched.c) compile unmodified against the GPU engine; link with -lnavtex_compat. */
#include "../navtex_compat.h"
