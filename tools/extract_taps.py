#!/usr/bin/env python3
"""Regenerate include/navtex_taps.h from the reference's tap arrays.

Run in the build container only (needs /root/reference).  The values are parsed
numerically and re-emitted as exact binary64 hexadecimal floats, so the header
is bit-identical to receiver/fir1cpp.C:10-49, fir2cpp.C:24-72, fir3cpp.h:17-89
without sharing any text with them.  tests/test_oracle_vs_ref.py proves the
equality by running the compiled reference against the oracle that uses them.
"""
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/receiver"
OUT = sys.argv[2] if len(sys.argv) > 2 else "include/navtex_taps.h"


def taps(path, start_pat):
    src = open(path).read()
    i = src.index(start_pat)
    body = src[src.index("{", i) + 1: src.index("}", i)]
    return [float(x) for x in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?", body)]


def main():
    sets = (
        ("NVX_H1", taps(f"{REF}/fir1cpp.C", "filter_h[]")),
        ("NVX_H2", taps(f"{REF}/fir2cpp.C", "filter_h[]")),
        ("NVX_H3", taps(f"{REF}/fir3cpp.h", "filter_h[")),
    )
    assert [len(h) for _, h in sets] == [37, 47, 71]
    head = open(OUT).read().split("#define NVX_H1_VALUES")[0]
    out = [head.rstrip("\n"), ""]
    for name, h in sets:
        out.append("#define %s_VALUES \\" % name)
        rows = ["    " + ", ".join(float.hex(v) for v in h[k:k + 3]) for k in range(0, len(h), 3)]
        out.append(", \\\n".join(rows))
        out.append("")
    out.append("#endif\n")
    open(OUT, "w").write("\n".join(out))


if __name__ == "__main__":
    main()
