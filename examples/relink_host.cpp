// relink_host.cpp -- a host written against the REFERENCE's own names, unchanged: what capt_sched.c + nav_sched.C do
// (receiver/capt_sched.c:17-18, :511, :552-555, :612; receiver/nav_sched.C:10-22), linked against libnavtex_compat.so
// instead of fir1cpp.o fir2cpp.o fir3cpp.o decoder.o nav_b_sm.o nav_sched.o.
//
//   g++ -std=c++17 -Iinclude examples/relink_host.cpp -Lnavtex_b200 -lnavtex_compat -lnavtex_b200
//       -Wl,-rpath,$PWD/navtex_b200 -o relink_host                       (one command line)
//   ./relink_host capture.s16      (raw interleaved int16 I,Q at 252 kS/s)
#include <stdio.h>
#include <string.h>

#include <vector>

#include "navtex_compat.h"

// message_store.c's symbol, defined by the host exactly as in the reference (nav_b_sm.C:4)
extern "C" int add_message(char* bbbb, char* message, int freq) {
    printf("%d|%s|%zu\n%s\n", freq, bbbb, strlen(message), message);
    return 0;
}

// nav_sched.C:10-22, verbatim in spirit: build the object graph and wire it
static byte_state_machine bsm518(518), bsm490(490);
static decoder dec518(&bsm518), dec490(&bsm490);
static fir_filter3 ff3_518(&dec518), ff3_490(&dec490);
extern "C" void init_fir2_wrapper_host() { init_fir_filter2(&ff3_518, &ff3_490); }

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s capture.s16\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<short> buf(1 << 16);
    init_fir2_wrapper_host();
    init_fir_filter1();                 // capt_sched.c:554
    init_fir2_wrapper();                // capt_sched.c:612
    size_t got;
    while ((got = fread(buf.data(), sizeof(short), buf.size(), f)) > 1)
        for (size_t i = 0; i + 1 < got; i += 2) sample_in_1((double)buf[i], (double)buf[i + 1]);     // capt_sched.c:511
    fclose(f);
    const int rc = navtex_compat_flush();
    navtex_compat_shutdown();
    return rc < 0 ? 1 : 0;
}
