// cr_math.cuh — correctly rounded acos, cos and pow(x, 1/3.) in double precision, for the literal cubic solver.
//
// The reference solves the foot-point cubic of a quadratic Bezier segment in closed form with the host libm's acos, cos and
// pow (/root/reference/core/equation-solver.cpp:45-52).  For almost every texel the result is insensitive to the last bit of
// those calls, but at a CUSP (control point on an end point, or collinear and doubling back) a root lands exactly on the
// boundary of (0,1) or multiplies a vanishing tangent, and the reference's texel is decided by that bit.  glibc's functions are
// correctly rounded in all but a vanishing fraction of arguments (errors <= 0.52 ulp); CUDA's are not (1-2 ulp).  These
// replacements evaluate in double-double arithmetic (~100 bits) and round once, i.e. they return the correctly rounded double
// unless the exact value lies within ~2^-45 ulp of a rounding boundary — which makes the device agree with the compiled
// reference on every cusp shape of the parity fuzz set (tests/test_fuzz.py family 8), bit for bit.
//
// Only segments flagged SEG_EXACT_ROOTS (and the fast path's rare fallbacks) reach these functions: speed is irrelevant, the
// code is kept small and out of line.  Everything is __host__ __device__ so that tests/ can check it against long double on
// the CPU (tools/cr_math_check.cpp).  fma() is the fused operation on both sides (-fmad=false does not affect explicit calls).
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define CRM_FN __host__ __device__ inline
#else
#define CRM_FN inline
#endif

namespace crm {

struct dd { double hi, lo; };

CRM_FN dd two_sum(double a, double b) {
    double s = a+b, bb = s-a;
    dd r; r.hi = s; r.lo = (a-(s-bb))+(b-bb);
    return r;
}
CRM_FN dd quick_two_sum(double a, double b) {       // |a| >= |b|
    double s = a+b;
    dd r; r.hi = s; r.lo = b-(s-a);
    return r;
}
CRM_FN dd two_prod(double a, double b) {
    double p = a*b;
    dd r; r.hi = p; r.lo = fma(a, b, -p);
    return r;
}
CRM_FN dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
CRM_FN dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo += a.lo;
    return quick_two_sum(s.hi, s.lo);
}
CRM_FN dd dd_neg(dd a) { dd r; r.hi = -a.hi; r.lo = -a.lo; return r; }
CRM_FN dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi*b.lo+a.lo*b.hi;
    return quick_two_sum(p.hi, p.lo);
}
CRM_FN dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo*b;
    return quick_two_sum(p.hi, p.lo);
}
CRM_FN dd dd_div_d(dd a, double b) {
    double q1 = a.hi/b;
    dd p = two_prod(q1, b);
    dd r = dd_add(a, dd_neg(p));
    double q2 = r.hi/b;
    p = two_prod(q2, b);
    r = dd_add(r, dd_neg(p));
    double q3 = r.hi/b;
    dd q = quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
CRM_FN dd dd_from(double a) { dd r; r.hi = a; r.lo = 0; return r; }

// sin and cos of a double-double |r| <= ~0.8 by their Taylor series (terms below 2^-110 of the leading one are dropped)
CRM_FN void dd_sincos_small(dd r, dd &s, dd &c) {
    const dd r2 = dd_mul(r, r);
    // cos: 1 - r^2/2! + r^4/4! ...   sin: r - r^3/3! + ...
    dd term = dd_from(1.0);
    c = term;
    dd sterm = r;
    s = r;
    for (int k = 1; k <= 16; ++k) {
        term = dd_div_d(dd_mul(term, r2), double((2*k-1)*(2*k)));
        sterm = dd_div_d(dd_mul(sterm, r2), double((2*k)*(2*k+1)));
        if (k&1) { c = dd_add(c, dd_neg(term)); s = dd_add(s, dd_neg(sterm)); }
        else { c = dd_add(c, term); s = dd_add(s, sterm); }
    }
}

// pi/2 to ~160 bits: three doubles
#define CRM_PIO2_1 1.5707963267948966
#define CRM_PIO2_2 6.123233995736766e-17
#define CRM_PIO2_3 -1.4973849048591698e-33

// sin and cos of a double x, |x| < ~100 (the solver's arguments lie in [-2.1, 3.2])
CRM_FN void dd_sincos(double x, dd &s, dd &c) {
    const double kf = rint(x*0.6366197723675814);       // x / (pi/2)
    const int k = int(kf);
    // r = x - k*pi/2 in double-double (k*PIO2_1 is exact for the |k| <= 2^20 that can occur; cancellation handled by two_sum)
    dd r = two_sum(x, -kf*CRM_PIO2_1);
    if (kf != 0) {
        // -kf*PIO2_1 may round for large k: recover the exact product
        dd p1 = two_prod(kf, CRM_PIO2_1);
        r = dd_add(dd_from(x), dd_neg(p1));
        r = dd_add(r, dd_neg(two_prod(kf, CRM_PIO2_2)));
        r = dd_add_d(r, -kf*CRM_PIO2_3);
    }
    dd sr, cr;
    dd_sincos_small(r, sr, cr);
    switch (k&3) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = dd_neg(sr); break;
        case 2: s = dd_neg(sr); c = dd_neg(cr); break;
        default: s = dd_neg(cr); c = sr; break;
    }
}

// correctly rounded cos(x)
CRM_FN double cr_cos(double x) {
    if (!(fabs(x) < 1e5))
        return cos(x);          // never reached by the solver; plain libm beyond the reduction's range
    dd s, c;
    dd_sincos(x, s, c);
    return c.hi+c.lo;
}

CRM_FN dd dd_sqrt(dd a) {
    if (!(a.hi > 0))
        return dd_from(0.0);
    const double s0 = sqrt(a.hi);
    dd r = dd_add(a, dd_neg(two_prod(s0, s0)));
    return dd_add_d(dd_from(s0), r.hi/(2*s0));
}

// correctly rounded acos(x), |x| <= 1: one Newton step (with its second-order term) from the double-precision estimate y0,
// the trigonometric functions of y0 evaluated in double-double.  |x| <= 0.7: on cos(y) = x.  Towards the ends cos is flat and
// that equation ill-conditioned, so there y = acos(|x|) is found from sin(y) = sqrt((1-|x|)(1+|x|)) — 1-|x| is exact — and
// acos(x) = pi - y for negative x.
CRM_FN double cr_acos(double x) {
    if (!(x >= -1 && x <= 1))
        return acos(x);         // NaN, as libm
    const double ax = fabs(x);
    if (ax <= 0.7) {
        const double y0 = acos(x);
        dd s, c;
        dd_sincos(y0, s, c);
        // cos(y0 + e) = x  =>  e = q - e^2 cot(y0)/2,  q = (cos(y0) - x)/sin(y0)
        dd q = dd_div_d(dd_add_d(c, -x), s.hi);
        q = dd_add(q, dd_neg(dd_mul_d(q, s.lo/s.hi)));          // divisor sin(y0) = s.hi + s.lo
        dd y = dd_add_d(q, y0);
        y = dd_add_d(y, -.5*(c.hi/s.hi)*q.hi*q.hi);
        return y.hi+y.lo;
    }
    const dd S = dd_sqrt(dd_mul_d(two_sum(1.0, ax), 1.0-ax));
    const double y0 = acos(ax);
    dd y = dd_from(0.0);
    if (y0 != 0) {
        dd s, c;
        dd_sincos(y0, s, c);
        // sin(y0 + e) = S  =>  e = q + e^2 tan(y0)/2,  q = (S - sin(y0))/cos(y0)
        dd q = dd_div_d(dd_add(S, dd_neg(s)), c.hi);
        q = dd_add(q, dd_neg(dd_mul_d(q, c.lo/c.hi)));
        y = dd_add_d(q, y0);
        y = dd_add_d(y, .5*(s.hi/c.hi)*q.hi*q.hi);
    } else
        y = S;                  // acos rounded to 0: y = asin(S) = S to far below the last bit
    if (x > 0)
        return y.hi+y.lo;
    // pi - y with pi = 2 * (PIO2_1 + PIO2_2 + PIO2_3)
    dd r = dd_add(dd_from(2*CRM_PIO2_1), dd_neg(y));
    r = dd_add_d(r, 2*CRM_PIO2_2);
    r = dd_add_d(r, 2*CRM_PIO2_3);
    return r.hi+r.lo;
}

// correctly rounded pow(x, 1/3.) for x > 0, where 1/3. is the double 0x3FD5555555555555 = 1/3 - 2^-54/3:
//   x^(1/3.) = cbrt(x) * exp(-(2^-54/3) ln x) = cbrt(x) * (1 - e ln x + O(1e-28)),  e = 2^-54/3
CRM_FN double cr_pow_third(double x) {
    if (!(x > 1e-290) || !(x < 1e300))
        return pow(x, 1/3.);    // 0, denormal range, inf, NaN, negative: libm (never reached by the solver with finite shapes)
    // cube root to double-double: one Newton step c1 = c0 - (c0^3 - x)/(3 c0^2) with the residual in double-double
    const double c0 = cbrt(x);
    dd c2 = two_prod(c0, c0);
    dd c3 = dd_mul_d(c2, c0);
    dd res = dd_add_d(c3, -x);
    dd corr = dd_div_d(res, 3.0*c2.hi);
    dd c = dd_add(dd_from(c0), dd_neg(corr));
    // second-order term of Newton for the cube root: the iterate overshoots by e^2/c
    c = dd_add_d(c, -corr.hi*corr.hi/c0);
    const double e = 1.850371707708594e-17;      // 2^-54 / 3
    dd f = dd_mul_d(c, -e*log(x));
    dd r = dd_add(c, f);
    return r.hi+r.lo;
}

}  // namespace crm
