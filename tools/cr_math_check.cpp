// Host check of msdfgen_b200/csrc/cr_math.cuh against quad precision (libquadmath): every result must be the correctly rounded
// double.  Build: g++ -O2 -ffp-contract=off -o /tmp/cr_math_check tools/cr_math_check.cpp -lquadmath ; run: /tmp/cr_math_check [samples]
// (tests/test_cr_math.py runs it with 300 000 samples per function)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <random>
#include <quadmath.h>
#include "../msdfgen_b200/csrc/cr_math.cuh"

int main(int argc, char **argv) {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> u(0, 1);
    long bad[3] = { 0, 0, 0 }, libm[3] = { 0, 0, 0 };
    const long N = argc > 1 ? atol(argv[1]) : 4000000;
    for (long i = 0; i < N; ++i) {
        // acos: mix of uniform arguments and arguments crowding the ends
        double x = 2*u(rng)-1;
        int sel = int(i%4);
        if (sel == 1) x = 1-ldexp(u(rng), -int(rng()%50));
        if (sel == 2) x = -1+ldexp(u(rng), -int(rng()%50));
        if (sel == 3) x = ldexp(2*u(rng)-1, -int(rng()%40));
        double want = (double) acosq((__float128) x);
        if (crm::cr_acos(x) != want) { if (bad[0]++ < 5) printf("acos(%a): got %a want %a\n", x, crm::cr_acos(x), want); }
        if (acos(x) != want) ++libm[0];
        // cos: the solver's range [-2.1, 3.2], plus arguments near pi/2 multiples
        double y = -2.2+5.5*u(rng);
        if (sel == 1) y = M_PI_2+ldexp(2*u(rng)-1, -int(rng()%45));
        if (sel == 2) y = 1/3.*(acos(x)+2*M_PI);
        if (sel == 3) y = 1/3.*(acos(x)-2*M_PI);
        want = (double) cosq((__float128) y);
        if (crm::cr_cos(y) != want) { if (bad[1]++ < 5) printf("cos(%a): got %a want %a\n", y, crm::cr_cos(y), want); }
        if (cos(y) != want) ++libm[1];
        // pow(x, 1/3.): magnitudes over 60 decades
        double z = ldexp(.5+u(rng), int(rng()%200)-100);
        want = (double) powq((__float128) z, (__float128) (1/3.));
        if (crm::cr_pow_third(z) != want) { if (bad[2]++ < 5) printf("pow(%a): got %a want %a\n", z, crm::cr_pow_third(z), want); }
        if (pow(z, 1/3.) != want) ++libm[2];
    }
    printf("%ld samples each: incorrectly rounded  cr_acos %ld  cr_cos %ld  cr_pow_third %ld   (host libm for comparison: %ld %ld %ld)\n",
           N, bad[0], bad[1], bad[2], libm[0], libm[1], libm[2]);
    return bad[0] || bad[1] || bad[2];
}
