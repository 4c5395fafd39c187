/* nav_b_sm.h -- same-named shim of the reference header receiver/nav_b_sm.h:56-128 (class byte_state_machine).
 * Put include/compat on the include path INSTEAD of the reference's receiver/ directory and the reference's own host
 * sources (nav_sched.C, capt_sched.c) compile unmodified against the GPU engine; link with -lnavtex_compat. */
#include "../navtex_compat.h"
