#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
static inline double divk(double a, double K, double y) { double q0 = a*y; double r = fma(-K, q0, a); return fma(r, y, q0); }
int main() {
    uint64_t s = 88172645463325252ull; long bad3=0,bad5=0,bad6=0,bad90=0,n=0;
    for (long i=0;i<400000000;++i) {
        s ^= s<<13; s ^= s>>7; s ^= s<<17;
        uint32_t b = (uint32_t)(s>>20); float f; memcpy(&f,&b,4);
        if (!isfinite(f) || f==0) continue;
        double a = (double)f;
        // also products/sums of floats in double: use a second variant with full double mantissa
        double a2 = a * (1.0 + (double)(s & 0xfffff) * 1e-7);
        if (divk(a,3.0,1.0/3.0) != a/3.0) bad3++;
        if (divk(a,5.0,1.0/5.0) != a/5.0) bad5++;
        if (divk(a2,3.0,1.0/3.0) != a2/3.0) bad3++;
        if (divk(a2,5.0,1.0/5.0) != a2/5.0) bad5++;
        if (divk(a2,6.0,1.0/6.0) != a2/6.0) bad6++;
        if (divk(a2,90.0,1.0/90.0) != a2/90.0) bad90++;
        n++;
    }
    printf("n=%ld bad3=%ld bad5=%ld bad6=%ld bad90=%ld\n", n,bad3,bad5,bad6,bad90);
    return 0;
}
