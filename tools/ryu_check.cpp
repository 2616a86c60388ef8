// Host check of subphaser_b200/csrc/spk_format.cuh: reads binary doubles on stdin, prints py_repr of each, one per line.
//   g++ -O2 -o /tmp/ryu_check tools/ryu_check.cpp && python tools/ryu_check.py
#include <stdio.h>
#include "../subphaser_b200/csrc/spk_format.cuh"
int main() {
    double v;
    char buf[64];
    while (fread(&v, 8, 1, stdin) == 1) {
        const int n = spkfmt::py_repr(v, buf);
        buf[n] = 0;
        puts(buf);
    }
    return 0;
}
