// distance_core.cuh — device-side geometry + selector + contour-combiner math for the sm_100a kernels.
//
// One thread owns one query point and walks the glyph's prepared segment records; all arithmetic on the distance path is
// fp64 and the translation unit is compiled with -fmad=false so that every expression rounds exactly like the reference's
// scalar C++ (which must itself be built without FMA contraction).
//
// Reference semantics restated here (paths relative to /root/reference/):
//   segment distance kernels      core/edge-segments.cpp:173-277, direction/point :108-139, perpendicular conversion :28-52
//   closed-form solvers           core/equation-solver.cpp:9-70
//   selectors (cache-less)        core/edge-selectors.cpp:19-38 (true), :42-52,81-117,129-164 (perpendicular), :174-260 (multi)
//   contour combiners             core/contour-combiners.cpp:34-50 (simple), :57-134 (overlapping)
//   traversal order               core/ShapeDistanceFinder.hpp:36-58 (oneShotDistance: every edge evaluated, last edge first)
// What is NOT the reference's structure: per-segment constants (unit tangents at both ends, corner bisectors, the cubic
// coefficients of the quadratic-Bezier foot-point equation) are hoisted into a SegRec by one thread per segment instead of
// being recomputed per pixel per edge (that recomputation is 42 % of the reference's profile); the
// EdgeCache pruning is dropped (it is a serial-scan optimisation); per-contour selectors are streamed and folded into
// shape/inner/outer accumulators instead of living in a std::vector.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>
#include "cr_math.cuh"

namespace b200 {

struct d2 { double x, y; };

__device__ __forceinline__ d2 mk(double x, double y) { d2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ d2 operator+(d2 a, d2 b) { return mk(a.x+b.x, a.y+b.y); }
__device__ __forceinline__ d2 operator-(d2 a, d2 b) { return mk(a.x-b.x, a.y-b.y); }
__device__ __forceinline__ d2 operator*(double s, d2 a) { return mk(s*a.x, s*a.y); }
__device__ __forceinline__ double dot(d2 a, d2 b) { return a.x*b.x+a.y*b.y; }
__device__ __forceinline__ double cross(d2 a, d2 b) { return a.x*b.y-a.y*b.x; }
// IEEE fp64 square root and division expand to 25-30 instructions each; the distance kernels issue them from dozens of sites.
// B200_COMPACT_MATH=1 routes them through two out-of-line functions (same correctly rounded results), which keeps the kernels'
// hot code inside the 32 KB L1.5 instruction cache.
#ifndef B200_COMPACT_MATH
#define B200_COMPACT_MATH 1
#endif
#if B200_COMPACT_MATH
__device__ __noinline__ double dsqrt(double x) { return sqrt(x); }
__device__ __noinline__ double ddiv(double a, double b) { return a/b; }
#else
__device__ __forceinline__ double dsqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double ddiv(double a, double b) { return a/b; }
#endif
__device__ __forceinline__ double len(d2 a) { return dsqrt(a.x*a.x+a.y*a.y); }
__device__ __forceinline__ bool is_zero(d2 a) { return a.x == 0 && a.y == 0; }
// Vector2::normalize(allowZero): zero vector -> (0, !allowZero)
__device__ __forceinline__ d2 normalize(d2 a, bool allowZero) {
    double l = len(a);
    if (l != 0)
        return mk(ddiv(a.x, l), ddiv(a.y, l));
    return mk(0, allowZero ? 0.0 : 1.0);
}
// a/|a| when the length is already known (same sqrt value the reference recomputes)
__device__ __forceinline__ d2 normalize_known(d2 a, double l) {
    if (l != 0)
        return mk(ddiv(a.x, l), ddiv(a.y, l));
    return mk(0, 1.0);
}
// mix(a, b, w) = (1-w)*a + w*b
__device__ __forceinline__ d2 mix(d2 a, d2 b, double w) { return mk((1.0-w)*a.x+w*b.x, (1.0-w)*a.y+w*b.y); }
__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double dmax(double a, double b) { return a < b ? b : a; }
__device__ __forceinline__ double median3(double a, double b, double c) { return dmax(dmin(a, b), dmin(dmax(a, b), c)); }
__device__ __forceinline__ float fmin2(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float fmax2(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float median3f(float a, float b, float c) { return fmax2(fmin2(a, b), fmin2(fmax2(a, b), c)); }
// nonZeroSign(n)*m
__device__ __forceinline__ double signed_by(double n, double m) { return n > 0 ? m : -m; }
// SignedDistance operator<
__device__ __forceinline__ bool sd_less(double ad, double adot, double bd, double bdot) {
    double x = fabs(ad), y = fabs(bd);
    return x < y || (x == y && adot < bdot);
}

enum : int {
    SEG_LAST_IN_CONTOUR = 1,   // contour ends after this record (evaluation order)
    SEG_CUBIC_NORMED = 2,      // quadratic segment: foot-point cubic takes the normalised (trig/Cardano) branch
    SEG_EXACT_ROOTS = 4        // quadratic segment with a cusp (vanishing tangent in [0,1]): literal closed-form roots, see sd_quadratic
};

// One prepared edge segment, in EVALUATION order (per contour: last edge first, then edges 0..n-2).
struct __align__(16) SegRec {
    d2 p0, p1, p2, p3;     // control points (unused = 0)
    d2 pEnd;               // point(1) = p[order]
    d2 aDir, bDir;         // direction(0), direction(1), normalize(true)
    d2 aBis, bBis;         // (prevDir+aDir).normalize(true), (bDir+nextDir).normalize(true)
    d2 aDirNZ, bDirNZ;     // direction(0), direction(1), normalize()   (zero -> (0,1))
    // type-specific constants:
    //   linear:    e0,1 = ab   e2 = ab.ab   e3,4 = ab.getOrthonormal(false)
    //   quadratic: e0,1 = ab   e2,3 = br    e4 = br.br  e5 = 3 ab.br  e6 = 2 ab.ab  e7,8 = direction(0)  e9 = its sq. length
    //              e10,11 = direction(1)  e12 = its sq. length  e13 = e5/e4
    //   cubic:     e0,1 = ab   e2,3 = br    e4,5 = as   e7..e12 as quadratic
    double e[14];
    d2 bc;                 // bounding circle of the whole curve (centre, radius incl. safety margin): every point of the
    double br;             // segment lies within br of bc — used only to SKIP work that provably cannot matter
    double pad2;
    int order;             // 1, 2, 3
    int color;             // EdgeColor mask
    int flags;
    int winding;           // Contour::winding() of the owning contour
    int cornerL, cornerB;  // texel l,b of protectCorners if this edge starts a colour corner, else INT_MIN/2
    int storageIndex;      // index of the segment in the caller's arrays
    int contour;           // index of the owning contour within the glyph (empty contours counted)
};

// Per-glyph geometry, built on the host when a batch is uploaded.
struct GlyphDesc {
    double sx, sy, tx, ty;          // Projection scale, translate
    double dmScale, dmTranslate;    // DistanceMapping
    unsigned long long outOffset;   // float offset of the glyph's [h][w][N] block in the output buffer
    unsigned long long texelOffset; // texel offset (stencil buffer)
    int w, h;
    unsigned c0, c1;                // contour range
    unsigned s0, s1;                // segment range (storage == evaluation range)
    int flip;                       // output orientation != shape orientation (rows reversed in memory)
    int shapeYDown;                 // Shape::inverseYAxis
    int nonEmptyContours;
    int x0, y0;                     // pixel origin of the block inside the full image (regions; 0 otherwise)
    int recBase;                    // >= 0: prepared records live in the global big-shape buffer at this index; -1: staged in smem
    int bigIndex;                   // index among the batch's big shapes (recBase >= 0)
};

struct RawShapes {
    const double *points;
    const uint8_t *orders;
    const uint8_t *colors;      // may be null (all WHITE)
    const uint32_t *cbegin;
    const int8_t *windings;
};

// ---------------------------------------------------------------------------------------------------------------------
// Raw segment access (used by the per-segment preparation only)

struct RawSeg { d2 p[4]; int order; };

__device__ __forceinline__ RawSeg load_raw(const RawShapes &rs, unsigned s) {
    RawSeg r;
    const double2 *q = reinterpret_cast<const double2 *>(rs.points+8*size_t(s));
    double2 a = q[0], b = q[1], c = q[2], d = q[3];
    r.p[0] = mk(a.x, a.y); r.p[1] = mk(b.x, b.y); r.p[2] = mk(c.x, c.y); r.p[3] = mk(d.x, d.y);
    r.order = rs.orders[s];
    return r;
}

// EdgeSegment::point(param)
__device__ inline d2 seg_point(const RawSeg &s, double t) {
    if (s.order == 1)
        return mix(s.p[0], s.p[1], t);
    if (s.order == 2)
        return mix(mix(s.p[0], s.p[1], t), mix(s.p[1], s.p[2], t), t);
    d2 p12 = mix(s.p[1], s.p[2], t);
    return mix(mix(mix(s.p[0], s.p[1], t), p12, t), mix(p12, mix(s.p[2], s.p[3], t), t), t);
}

// EdgeSegment::direction(param)
__device__ inline d2 seg_direction(const RawSeg &s, double t) {
    if (s.order == 1)
        return s.p[1]-s.p[0];
    if (s.order == 2) {
        d2 tangent = mix(s.p[1]-s.p[0], s.p[2]-s.p[1], t);
        if (is_zero(tangent))
            return s.p[2]-s.p[0];
        return tangent;
    }
    d2 tangent = mix(mix(s.p[1]-s.p[0], s.p[2]-s.p[1], t), mix(s.p[2]-s.p[1], s.p[3]-s.p[2], t), t);
    if (is_zero(tangent)) {
        if (t == 0) return s.p[2]-s.p[0];
        if (t == 1) return s.p[3]-s.p[1];
    }
    return tangent;
}

// Contour::winding() for contour c (one thread).
__device__ inline int contour_winding(const RawShapes &rs, unsigned c) {
    unsigned b = rs.cbegin[c], e = rs.cbegin[c+1];
    unsigned n = e-b;
    if (n == 0)
        return 0;
    double total = 0;
    #define B200_SHOELACE(A, B) (((B).x-(A).x)*((A).y+(B).y))
    if (n == 1) {
        RawSeg s = load_raw(rs, b);
        d2 p = seg_point(s, 0), q = seg_point(s, 1/3.), r = seg_point(s, 2/3.);
        total += B200_SHOELACE(p, q);
        total += B200_SHOELACE(q, r);
        total += B200_SHOELACE(r, p);
    } else if (n == 2) {
        RawSeg s0 = load_raw(rs, b), s1 = load_raw(rs, b+1);
        d2 p = seg_point(s0, 0), q = seg_point(s0, .5), r = seg_point(s1, 0), t = seg_point(s1, .5);
        total += B200_SHOELACE(p, q);
        total += B200_SHOELACE(q, r);
        total += B200_SHOELACE(r, t);
        total += B200_SHOELACE(t, p);
    } else {
        d2 prev = seg_point(load_raw(rs, e-1), 0);
        for (unsigned i = b; i < e; ++i) {
            d2 cur = seg_point(load_raw(rs, i), 0);
            total += B200_SHOELACE(prev, cur);
            prev = cur;
        }
    }
    #undef B200_SHOELACE
    return (0 < total)-(total < 0);
}

// Builds the record for evaluation position j of a glyph (j counts from the glyph's first segment; positions map to
// contours exactly like storage indices do, only rotated inside each contour: position k of a contour [b,e) is edge e-1
// for k == 0 and edge b+k-1 otherwise — ShapeDistanceFinder.hpp:22-28).
__device__ inline void prepare_record(const RawShapes &rs, const GlyphDesc &g, unsigned j, SegRec &out) {
    unsigned s = g.s0+j;
    // contour containing storage index s (contours of a glyph: [c0, c1))
    unsigned lo = g.c0, hi = g.c1;
    while (hi-lo > 1) {
        unsigned mid = (lo+hi)>>1;
        if (rs.cbegin[mid] <= s) lo = mid; else hi = mid;
    }
    // skip over empty contours that share the same begin offset: the binary search lands on the LAST contour whose begin <= s
    unsigned c = lo;
    unsigned b = rs.cbegin[c], e = rs.cbegin[c+1];
    unsigned k = s-b;
    unsigned cur = k == 0 ? e-1 : s-1;
    unsigned prev = cur == b ? e-1 : cur-1;
    unsigned next = cur == e-1 ? b : cur+1;
    RawSeg sc = load_raw(rs, cur);
    RawSeg sp = load_raw(rs, prev);
    RawSeg sn = load_raw(rs, next);

    out.p0 = sc.p[0];
    out.p1 = sc.p[1];
    out.p2 = sc.order >= 2 ? sc.p[2] : mk(0, 0);
    out.p3 = sc.order >= 3 ? sc.p[3] : mk(0, 0);
    out.pEnd = seg_point(sc, 1);
    out.p0 = seg_point(sc, 0);    // point(0): identical to p[0] for finite input, kept literal
    d2 dir0 = seg_direction(sc, 0), dir1 = seg_direction(sc, 1);
    out.aDir = normalize(dir0, true);
    out.bDir = normalize(dir1, true);
    out.aDirNZ = normalize(dir0, false);
    out.bDirNZ = normalize(dir1, false);
    d2 prevDir = normalize(seg_direction(sp, 1), true);
    d2 nextDir = normalize(seg_direction(sn, 0), true);
    out.aBis = normalize(prevDir+out.aDir, true);
    out.bBis = normalize(out.bDir+nextDir, true);
    out.order = sc.order;
    int color = rs.colors ? rs.colors[cur] : 7;
    out.color = color;
    int flags = (k == e-b-1) ? SEG_LAST_IN_CONTOUR : 0;
    out.winding = rs.windings[c];
    out.storageIndex = int(cur);
    out.contour = int(c-g.c0);
    #pragma unroll
    for (int i = 0; i < 14; ++i)
        out.e[i] = 0;
    // p0 restored to the raw control point for the distance kernels (point(0) above only differs for non-finite input)
    d2 P0 = sc.p[0], P1 = sc.p[1], P2 = sc.p[2], P3 = sc.p[3];
    {
        // control-point hull: a Bezier curve lies inside the convex hull of its control points
        double xmin = fmin(P0.x, P1.x), xmax = fmax(P0.x, P1.x), ymin = fmin(P0.y, P1.y), ymax = fmax(P0.y, P1.y);
        if (sc.order >= 2) { xmin = fmin(xmin, P2.x); xmax = fmax(xmax, P2.x); ymin = fmin(ymin, P2.y); ymax = fmax(ymax, P2.y); }
        if (sc.order >= 3) { xmin = fmin(xmin, P3.x); xmax = fmax(xmax, P3.x); ymin = fmin(ymin, P3.y); ymax = fmax(ymax, P3.y); }
        d2 c = mk(.5*(xmin+xmax), .5*(ymin+ymax));
        double r2 = fmax(dot(P0-c, P0-c), dot(P1-c, P1-c));
        if (sc.order >= 2) r2 = fmax(r2, dot(P2-c, P2-c));
        if (sc.order >= 3) r2 = fmax(r2, dot(P3-c, P3-c));
        out.bc = c;
        out.br = sqrt(r2)*(1+1e-9);
        out.pad2 = 0;
    }
    if (sc.order == 1) {
        d2 ab = P1-P0;
        out.e[0] = ab.x; out.e[1] = ab.y;
        out.e[2] = dot(ab, ab);
        double l = len(ab);
        d2 on = l != 0 ? mk(ab.y/l, -ab.x/l) : mk(0, -1.0);    // getOrthonormal(false)
        out.e[3] = on.x; out.e[4] = on.y;
    } else {
        d2 ab = P1-P0;
        d2 br = P2-P1-ab;
        out.e[0] = ab.x; out.e[1] = ab.y;
        out.e[2] = br.x; out.e[3] = br.y;
        out.e[7] = dir0.x; out.e[8] = dir0.y; out.e[9] = dot(dir0, dir0);
        out.e[10] = dir1.x; out.e[11] = dir1.y; out.e[12] = dot(dir1, dir1);
        if (sc.order == 2) {
            double a = dot(br, br);
            double bq = 3*dot(ab, br);
            out.e[4] = a;
            out.e[5] = bq;
            out.e[6] = 2*dot(ab, ab);
            if (a != 0) {
                double bn = bq/a;
                out.e[13] = bn;
                out.p3.x = 1/a;     // quadratics do not use p3: reciprocal of the leading coefficient for the fast root path
                if (fabs(bn) < 1e6)
                    flags |= SEG_CUBIC_NORMED;
            }
            // Cusps: control point on an end point, or collinear and doubling back — exactly or within rounding (the same shape
            // after a scale or a rotation).  The tangent ab + t br (all but) vanishes at a parameter t* in [0,1] that is itself a
            // root of the foot-point cubic, and what the reference returns for texels nearest to the cusp hinges on the last bits
            // of that root (is it inside (0,1)? which sign has cross(~0, q)?).  The polished roots of the fast path are accurate
            // but not bit-faithful, so these segments take the literal closed form, which follows the reference wherever its own
            // outcome is not rounding noise.  Criterion: the shortest tangent over [0,1] is below 1e-4 of the tangents' scale
            // (ordinary outlines are nowhere near; a false positive only costs the slower solver).
            const double tq = a != 0 ? fmin(fmax(-dot(ab, br)/a, 0.0), 1.0) : 0.0;
            const d2 vq = ab+tq*br;
            if (dot(vq, vq) <= 1e-8*(dot(ab, ab)+a))
                flags |= SEG_EXACT_ROOTS;
            // Chord twins: another edge of the glyph runs from this curve's end point back to its start point — the same curve
            // traversed in the opposite direction (abutting outlines that share an edge, as font editors leave them), or a
            // two-edge contour (a sliver when the two nearly coincide).  Texels nearest to such a pair are equidistant from
            // both, or nearly so, with opposite signs, and the reference's pick between them hangs on the last bit of its
            // closed-form roots; only the literal solver reproduces that.  (Exact duplicates in the same direction tie
            // identically in every solver.)  Glyph-sized shapes only: the search is linear in the shape's segment count.
            if (!(flags&SEG_EXACT_ROOTS) && g.s1-g.s0 <= 16384u) {
                for (unsigned k = g.s0; k < g.s1; ++k) {
                    const double *q = rs.points+8*size_t(k);
                    if (q[0] == P2.x && q[1] == P2.y && k != cur) {
                        const int ok = rs.orders[k];
                        if (q[2*ok] == P0.x && q[2*ok+1] == P0.y) {
                            flags |= SEG_EXACT_ROOTS;
                            break;
                        }
                    }
                }
            }
        } else {
            d2 as = (P3-P2)-(P2-P1)-br;
            out.e[4] = as.x; out.e[5] = as.y;
        }
    }
    out.flags = flags;
    // protectCorners (MSDFErrorCorrection.cpp:121-151): corner where the colour changes between prev and cur
    int prevColor = rs.colors ? rs.colors[prev] : 7;
    int common = prevColor&color;
    int cl = INT_MIN/2, cb = INT_MIN/2;
    if (!(common&(common-1))) {
        d2 c0 = out.p0;
        double px = g.sx*(c0.x+g.tx), py = g.sy*(c0.y+g.ty);    // Projection::project
        double fl = floor(px-.5), fb = floor(py-.5);
        if (fl > -1e9 && fl < 1e9 && fb > -1e9 && fb < 1e9) {
            cl = int(fl);
            cb = int(fb);
        }
    }
    out.cornerL = cl;
    out.cornerB = cb;
    out.p0 = P0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Closed-form solvers (equation-solver.cpp). Roots are returned in registers.

struct Roots { double t0, t1, t2; int n; };

__device__ __forceinline__ Roots solve_quadratic(double a, double b, double c) {
    Roots r; r.t0 = r.t1 = r.t2 = 0; r.n = 0;
    if (a == 0 || fabs(b) > 1e12*fabs(a)) {
        if (b == 0) {
            r.n = c == 0 ? -1 : 0;
            return r;
        }
        r.t0 = ddiv(-c, b);
        r.n = 1;
        return r;
    }
    double dscr = b*b-4*a*c;
    if (dscr > 0) {
        dscr = dsqrt(dscr);
        r.t0 = ddiv(-b+dscr, 2*a);
        r.t1 = ddiv(-b-dscr, 2*a);
        r.n = 2;
    } else if (dscr == 0) {
        r.t0 = ddiv(-b, 2*a);
        r.n = 1;
    }
    return r;
}

// solveCubicNormed (equation-solver.cpp:34-61) with correctly rounded acos / cos / pow (cr_math.cuh): reproduces the reference's
// roots to the bit wherever the host libm rounds correctly.  Cold: cusp segments (SEG_EXACT_ROOTS) and the fast path's guard-band
// and range fallbacks — kept out of line.
__device__ unsigned long long g_exactRootCalls;     // how often the literal solver ran (b200msdf_debug_exact_root_calls)
__device__ __noinline__ Roots solve_cubic_normed(double a, double b, double c) {
    Roots x; x.t0 = x.t1 = x.t2 = 0;
    atomicAdd(&g_exactRootCalls, 1ull);
    double a2 = a*a;
    double q = 1/9.*(a2-3*b);
    double r = 1/54.*(a*(2*a2-9*b)+27*c);
    double r2 = r*r;
    double q3 = q*q*q;
    a *= 1/3.;
    if (r2 < q3) {
        double t = r/sqrt(q3);
        if (t < -1) t = -1;
        if (t > 1) t = 1;
        t = crm::cr_acos(t);
        q = -2*sqrt(q);
        x.t0 = q*crm::cr_cos(1/3.*t)-a;
        x.t1 = q*crm::cr_cos(1/3.*(t+2*3.14159265358979323846))-a;
        x.t2 = q*crm::cr_cos(1/3.*(t-2*3.14159265358979323846))-a;
        x.n = 3;
    } else {
        double u = (r < 0 ? 1 : -1)*crm::cr_pow_third(fabs(r)+sqrt(r2-q3));
        double v = u == 0 ? 0 : q/u;
        x.t0 = (u+v)-a;
        if (u == v || fabs(u-v) < 1e-12*fabs(u+v)) {
            x.t1 = -.5*(u+v)-a;
            x.n = 2;
        } else
            x.n = 1;
    }
    return x;
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast roots of the foot-point cubic t^3 + A t^2 + B t + C = 0 (quadratic segments).
//
// The reference evaluates the closed form in fp64 (acos + 3 cos, or pow — equation-solver.cpp:34-61); on the GPU those
// software transcendentals and the two normalising divisions are the largest single cost of the distance kernel.  The roots
// are only used as curve parameters at which the squared distance is STATIONARY, so the distance obtained from a root
// carries an error that is second order in the root's error: any t within ~1e-10 of the true root reproduces the
// reference's distance to the last bit or two of fp64, far below the fp32 texel.  So: closed form in fp32 (hardware
// rsqrt / cos approximations) for a ~1e-6 starting value, then Newton steps in fp64 on the cubic (reciprocal of the
// derivative in fp32 — Newton is self-correcting) down to ~1e-14.  Only roots near [0,1] are polished; coefficient ranges
// the fp32 stage cannot represent fall back to the exact path.  Compile with -DB200_EXACT_ROOTS=1 for the literal solver.
#ifndef B200_EXACT_ROOTS
#define B200_EXACT_ROOTS 0
#endif
#ifndef B200_POLISH_STEPS
#define B200_POLISH_STEPS 3
#endif

__device__ __noinline__ double cubic_polish(double t, double A, double B, double C) {
    #pragma unroll
    for (int i = 0; i < B200_POLISH_STEPS; ++i) {
        double f = fma(fma(t+A, t, B), t, C);
        double df = fma(fma(3.0, t, A+A), t, B);
        float rdf = __frcp_rn(float(df));
        // a vanishing derivative (double root) leaves t where it is: such roots are inflections of the distance, never its minimum
        if (isfinite(rdf))
            t = fma(-f, double(rdf), t);
    }
    return t;
}

__device__ __forceinline__ Roots solve_cubic_fast(double A, double B, double C, bool &ok) {
    Roots x; x.t0 = x.t1 = x.t2 = 0; x.n = 0;
    double a2 = A*A;
    double q = 1/9.*(a2-3*B);
    double r = 1/54.*(A*(2*a2-9*B)+27*C);
    ok = fabs(q) < 1e12 && fabs(r) < 1e18;      // keep q^3, r^2 inside fp32 range
    if (!ok)
        return x;
    float qf = float(q), rf = float(r);
    double a3 = A*(1/3.);
    double cand0, cand1 = 2, cand2 = 2;         // 2 = "no root" (outside the polishing window)
    if (r*r < q*q*q) {
        float sq = sqrtf(qf);
        float t = rf/(qf*sq);
        t = fminf(fmaxf(t, -1.f), 1.f);
        float th = acosf(t);
        float m = -2.f*sq;
        cand0 = double(m*__cosf(th*(1/3.f)))-a3;
        cand1 = double(m*__cosf((th+6.2831853f)*(1/3.f)))-a3;
        cand2 = double(m*__cosf((th-6.2831853f)*(1/3.f)))-a3;
    } else {
        float disc = sqrtf(fmaxf(rf*rf-qf*qf*qf, 0.f));
        float u = cbrtf(fabsf(rf)+disc);
        u = rf < 0 ? u : -u;
        float v = u == 0 ? 0.f : qf/u;
        cand0 = double(u+v)-a3;
        if (fabsf(u-v) <= 1e-3f*fabsf(u+v))
            cand1 = double(-.5f*(u+v))-a3;      // (near-)double root
    }
    // polish what can end up inside (0,1)
    if (cand0 > -.05 && cand0 < 1.05) { x.t0 = cubic_polish(cand0, A, B, C); x.n = 1; }
    if (cand1 > -.05 && cand1 < 1.05) { double t = cubic_polish(cand1, A, B, C); if (x.n == 0) x.t0 = t; else x.t1 = t; ++x.n; }
    if (cand2 > -.05 && cand2 < 1.05) { double t = cubic_polish(cand2, A, B, C); if (x.n == 0) x.t0 = t; else if (x.n == 1) x.t1 = t; else x.t2 = t; ++x.n; }
    return x;
}

// ---------------------------------------------------------------------------------------------------------------------
// EdgeSegment::signedDistance for the three segment types

// Result of EdgeSegment::signedDistance.  The reference also returns the curve parameter of the nearest point, but its
// only consumers are range tests (param < 0, param > 1, 0 <= param <= 1, param < .5); those are decided exactly from the
// numerator and denominator of the quotient (num/den < 0 <=> num < 0, num/den > 1 <=> num > den, num/den < .5 <=>
// num < .5*den for den > 0 under round-to-nearest), so the fp64 division itself is never issued.
// `dot` (the orthogonality measure that only breaks exact |distance| ties between edges meeting at a corner) is evaluated
// only when the edge can still become some channel's minimum, i.e. |d| <= need (SignedDistance::operator<): two fp64
// divisions saved for every other edge.  When it is skipped the returned value is never read.
struct SDist { double d, dot; int side; };   // side: 0 param in [0,1] (or NaN), 1 param < 0, 2 param > 1

struct Param { double num, den; };
__device__ __forceinline__ bool param_lt0(Param q) { return q.den > 0 ? q.num < 0 : (q.den == 0 && q.num != 0 && q.num < 0); }
__device__ __forceinline__ bool param_gt1(Param q) { return q.den > 0 ? q.num > q.den : (q.den == 0 && q.num > 0); }
__device__ __forceinline__ bool param_in01(Param q) { return q.den > 0 && q.num >= 0 && q.num <= q.den; }
__device__ __forceinline__ bool param_lt_half(Param q) { return q.den > 0 ? q.num < .5*q.den : (q.den == 0 && q.num < 0); }
__device__ __forceinline__ bool param_gt_half(Param q) { return q.den > 0 ? q.num > .5*q.den : (q.den == 0 && q.num > 0); }
__device__ __forceinline__ int param_side(Param q) { return param_lt0(q) ? 1 : param_gt1(q) ? 2 : 0; }

__device__ __forceinline__ SDist sd_linear(const SegRec &s, d2 p, double need) {
    d2 aq = p-s.p0;
    d2 ab = mk(s.e[0], s.e[1]);
    Param t;
    t.num = dot(aq, ab);
    t.den = s.e[2];
    SDist r;
    r.side = param_side(t);
    d2 ep = param_gt_half(t) ? s.p1 : s.p0;
    d2 eq = ep-p;
    double endpointDistance = len(eq);
    if (t.den > 0 && t.num > 0 && t.num < t.den) {    // param > 0 && param < 1
        double ortho = dot(mk(s.e[3], s.e[4]), aq);
        if (fabs(ortho) < endpointDistance) {
            r.d = ortho;
            r.dot = 0;
            return r;
        }
    }
    r.d = signed_by(cross(aq, ab), endpointDistance);
    r.dot = endpointDistance <= need ? fabs(dot(s.aDirNZ, normalize_known(eq, endpointDistance))) : 0.0;
    return r;
}

// One root of the foot-point cubic (edge-segments.cpp:209-216).  `unsure` is raised when the outcome hinges on the last bits of
// the root: a root within 1e-9 of the ends of (0,1), or a distance within 1e-13 (relative) of the running minimum — the fast
// path's polished roots are accurate to ~1e-14 but not bit-faithful, so such an edge is re-evaluated with the literal solver.
#ifndef B200_GUARD_BAND
#ifdef B200_GUARD_MODE
#define B200_GUARD_BAND (B200_GUARD_MODE != 0)
#else
#define B200_GUARD_BAND 1
#endif
#endif
__device__ __forceinline__ void quad_root(d2 qa, d2 ab, d2 br, double t, double &minDistance, bool &interior, bool &unsure) {
#if B200_GUARD_BAND
    if (fabs(t) < 1e-9 || fabs(t-1) < 1e-9)
        unsure = true;
#endif
    if (t > 0 && t < 1) {
        d2 qe = qa+(2*t)*ab+(t*t)*br;
        double distance = len(qe);
#if B200_GUARD_BAND
        if (fabs(distance-fabs(minDistance)) <= 1e-13*distance)
            unsure = true;
#endif
        if (distance <= fabs(minDistance)) {
            minDistance = signed_by(cross(ab+t*br, qe), distance);
            interior = true;
        }
    }
}

// The part of QuadraticSegment::signedDistance that follows the root solve (edge-segments.cpp:198-226)
__device__ __forceinline__ SDist quad_finish(const SegRec &s, d2 p, d2 qa, d2 ab, d2 br, const Roots &t, double need, bool &unsure) {
    d2 epDir = mk(s.e[7], s.e[8]);
    double lenA = len(qa);
    double minDistance = signed_by(cross(epDir, qa), lenA);
    Param param;
    param.num = -dot(qa, epDir);
    param.den = s.e[9];
    d2 qb = s.p2-p;
    double lenB = len(qb);
    if (lenB < fabs(minDistance)) {
        epDir = mk(s.e[10], s.e[11]);
        minDistance = signed_by(cross(epDir, qb), lenB);
        param.num = dot(p-s.p1, epDir);
        param.den = s.e[12];
    }
    bool interior = false;     // param became a root in (0,1)
    if (t.n > 0) quad_root(qa, ab, br, t.t0, minDistance, interior, unsure);
    if (t.n > 1) quad_root(qa, ab, br, t.t1, minDistance, interior, unsure);
    if (t.n > 2) quad_root(qa, ab, br, t.t2, minDistance, interior, unsure);

    SDist r;
    r.d = minDistance;
    r.dot = 0;
    if (interior || param_in01(param))
        r.side = 0;
    else {
        r.side = param_side(param);
        if (fabs(minDistance) <= need) {
            if (param_lt_half(param))
                r.dot = fabs(dot(s.aDirNZ, normalize_known(qa, lenA)));
            else
                r.dot = fabs(dot(s.bDirNZ, normalize_known(qb, lenB)));
        }
    }
    return r;
}

// Per-thread flag in shared memory: set when the guard band of quad_root was hit during the thread's current texel.
__device__ __forceinline__ unsigned char &guard_flag() {
    __shared__ unsigned char s_guard[1024];
    return s_guard[threadIdx.x];
}

// LITERAL = false: fast foot-point roots (fp32 closed form + fp64 polish); guard_flag() is raised when the guard band of
// quad_root was hit — k1b_exact then re-evaluates the texel with LITERAL = true (texel_literal, kernels.cuh), which takes the
// reference's closed form with correctly rounded acos / cos / pow for every quadratic segment.
#ifndef B200_NOINLINE_SD
#define B200_NOINLINE_SD 0
#endif
#if B200_NOINLINE_SD
#define B200_SD_INLINE __noinline__
#else
#define B200_SD_INLINE __forceinline__
#endif
template <bool LITERAL>
__device__ B200_SD_INLINE SDist sd_quadratic(const SegRec &s, d2 p, double need) {
    d2 qa = s.p0-p;
    d2 ab = mk(s.e[0], s.e[1]);
    d2 br = mk(s.e[2], s.e[3]);
    double a = s.e[4];
    double b = s.e[5];
    double c = s.e[6]+dot(qa, br);
    double d = dot(qa, ab);
    Roots t;
    bool fast = false;
    if (s.flags&SEG_CUBIC_NORMED) {
        bool ok = !LITERAL && !B200_EXACT_ROOTS && !(s.flags&SEG_EXACT_ROOTS);
        if (ok)
            t = solve_cubic_fast(s.e[13], c*s.p3.x, d*s.p3.x, ok);
        fast = ok;
        if (!ok)
            t = solve_cubic_normed(s.e[13], c/a, d/a);
    } else
        t = solve_quadratic(b, c, d);
    bool unsure = false;
    SDist r = quad_finish(s, p, qa, ab, br, t, need, unsure);
    if (fast && unsure)
        guard_flag() = 1;       // (a store on a rare path instead of a value carried through the callers' hot loops)
    return r;
}

__device__ __noinline__ SDist sd_cubic(const SegRec &s, d2 p, double need) {
    d2 qa = s.p0-p;
    d2 ab = mk(s.e[0], s.e[1]);
    d2 br = mk(s.e[2], s.e[3]);
    d2 as = mk(s.e[4], s.e[5]);

    d2 epDir = mk(s.e[7], s.e[8]);
    double lenA = len(qa);
    double minDistance = signed_by(cross(epDir, qa), lenA);
    Param param;
    param.num = -dot(qa, epDir);
    param.den = s.e[9];
    d2 qb = s.p3-p;
    double lenB = len(qb);
    if (lenB < fabs(minDistance)) {
        epDir = mk(s.e[10], s.e[11]);
        minDistance = signed_by(cross(epDir, qb), lenB);
        param.num = dot(epDir-qb, epDir);
        param.den = s.e[12];
    }
    bool interior = false;
    for (int i = 0; i <= 4; ++i) {      // MSDFGEN_CUBIC_SEARCH_STARTS
        double t = 1./4*i;
        d2 qe = qa+(3*t)*ab+(3*t*t)*br+(t*t*t)*as;
        d2 d1 = 3*ab+(6*t)*br+(3*t*t)*as;
        d2 d2v = 6*br+(6*t)*as;
        double improvedT = t-dot(qe, d1)/(dot(d1, d1)+dot(qe, d2v));
        if (improvedT > 0 && improvedT < 1) {
            int remainingSteps = 4;     // MSDFGEN_CUBIC_SEARCH_STEPS
            do {
                t = improvedT;
                qe = qa+(3*t)*ab+(3*t*t)*br+(t*t*t)*as;
                d1 = 3*ab+(6*t)*br+(3*t*t)*as;
                if (!--remainingSteps)
                    break;
                d2v = 6*br+(6*t)*as;
                improvedT = t-dot(qe, d1)/(dot(d1, d1)+dot(qe, d2v));
            } while (improvedT > 0 && improvedT < 1);
            double distance = len(qe);
            if (distance < fabs(minDistance)) {
                minDistance = signed_by(cross(d1, qe), distance);
                interior = true;    // param = t, which lies in (0,1)
            }
        }
    }
    SDist r;
    r.d = minDistance;
    r.dot = 0;
    if (interior || param_in01(param))
        r.side = 0;
    else {
        r.side = param_side(param);
        if (fabs(minDistance) <= need) {
            if (param_lt_half(param))
                r.dot = fabs(dot(s.aDirNZ, normalize_known(qa, lenA)));
            else
                r.dot = fabs(dot(s.bDirNZ, normalize_known(qb, lenB)));
        }
    }
    return r;
}

template <bool LITERAL = false>
__device__ __forceinline__ SDist signed_distance(const SegRec &s, d2 p, double need) {
    // s.order is uniform across the CTA (every thread walks the same record), so this switch does not diverge
    if (s.order == 1)
        return sd_linear(s, p, need);
    if (s.order == 2)
        return sd_quadratic<LITERAL>(s, p, need);
    return sd_cubic(s, p, need);
}

// ---------------------------------------------------------------------------------------------------------------------
// Selectors. MODE: 1 SDF (true distance), 2 PSDF (perpendicular), 3 MSDF (multi), 4 MTSDF (multi + true)

template <int MODE> struct ModeTraits {
    static const int NCH = MODE >= 3 ? 3 : 1;                   // selector channels
    static const int NOUT = MODE == 3 ? 3 : MODE == 4 ? 4 : 1;  // distances per point
};

// PerpendicularDistanceSelectorBase state. `near` packs the nearest record index and on which side of [0,1] its
// parameter fell (all that distanceToPerpendicularDistance needs): (index << 2) | side, side 1: param < 0, 2: param > 1.
struct Chan {
    double d, dot;      // minTrueDistance
    double neg, pos;    // min negative / positive perpendicular distance
    int near;           // -1 = NULL
};

__device__ __forceinline__ void chan_init(Chan &c) {
    c.d = -DBL_MAX; c.dot = 0; c.neg = -DBL_MAX; c.pos = DBL_MAX; c.near = -1;
}
__device__ __forceinline__ void chan_add_perp(Chan &c, double x) {
    if (x <= 0 && x > c.neg) c.neg = x;
    if (x >= 0 && x < c.pos) c.pos = x;
}
template <bool PERP>
__device__ __forceinline__ void chan_merge(Chan &c, const Chan &o) {
    if (sd_less(o.d, o.dot, c.d, c.dot)) { c.d = o.d; c.dot = o.dot; c.near = o.near; }
    if (PERP) {
        if (o.neg > c.neg) c.neg = o.neg;
        if (o.pos < c.pos) c.pos = o.pos;
    }
}
// computeDistance (edge-selectors.cpp:108-117), once per contour and channel: out of line (by value, so that the selector
// state stays in registers) to keep the kernels' instruction footprint small
__device__ __noinline__ double chan_distance_values(double cd, double neg, double pos, int near, const SegRec *recs, double px, double py) {
    d2 p = mk(px, py);
    double m = cd < 0 ? neg : pos;
    if (near >= 0) {
        double dd = cd;
        int side = near&3;
        if (side) {
            const SegRec &s = recs[near>>2];
            if (side == 1) {
                d2 aq = p-s.p0;
                double ts = dot(aq, s.aDirNZ);
                if (ts < 0) {
                    double pd = cross(aq, s.aDirNZ);
                    if (fabs(pd) <= fabs(dd))
                        dd = pd;
                }
            } else {
                d2 bq = p-s.pEnd;
                double ts = dot(bq, s.bDirNZ);
                if (ts > 0) {
                    double pd = cross(bq, s.bDirNZ);
                    if (fabs(pd) <= fabs(dd))
                        dd = pd;
                }
            }
        }
        if (fabs(dd) < fabs(m))
            m = dd;
    }
    return m;
}
__device__ __forceinline__ double chan_distance(const Chan &c, const SegRec *recs, d2 p) {
    return chan_distance_values(c.d, c.neg, c.pos, c.near, recs, p.x, p.y);
}

template <int MODE> struct Sel { Chan ch[ModeTraits<MODE>::NCH]; };

template <int MODE>
__device__ __forceinline__ void sel_init(Sel<MODE> &s) {
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
        chan_init(s.ch[k]);
}

template <int MODE>
__device__ __forceinline__ void sel_merge(Sel<MODE> &a, const Sel<MODE> &b) {
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
        chan_merge<MODE != 1>(a.ch[k], b.ch[k]);
}

template <int MODE>
__device__ __forceinline__ void sel_distance(const Sel<MODE> &s, const SegRec *recs, d2 p, double *out) {
    if (MODE == 1) {
        out[0] = s.ch[0].d;
        return;
    }
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
        out[k] = chan_distance(s.ch[k], recs, p);
    if (MODE == 4) {    // MultiDistanceSelector::trueDistance
        double td = s.ch[0].d, tdot = s.ch[0].dot;
        #pragma unroll
        for (int k = 1; k < ModeTraits<MODE>::NCH; ++k)
            if (sd_less(s.ch[k].d, s.ch[k].dot, td, tdot)) { td = s.ch[k].d; tdot = s.ch[k].dot; }
        out[ModeTraits<MODE>::NOUT-1] = td;
    }
}

// {True,Perpendicular,Multi}DistanceSelector::addEdge (edge-selectors.cpp:19-29,129-160,174-227), cache-less.
//
// Exact culling.  cap[k] is an upper bound of channel k's FINAL minimum true distance within the selector's scope (the
// distance from p to the start point of some edge of that channel).  With need = max over the edge's channels of
// min(cap, current minimum), an edge whose bounding circle is farther than need from p has |d| > need, hence
//   * it cannot become (or tie with) any of its channels' final minimum, so its signedDistance is not evaluated, and
//   * a perpendicular candidate pd of it matters only if |pd| is below the channel's final |minimum| (a larger |pd| can
//     at most become minNeg/minPos, which computeDistance then rejects in favour of the nearest edge's own distance); the
//     reference adds pd iff |pd| < |d|, which for |pd| below the circle's lower bound is certainly true, and for every
//     other |pd| is irrelevant by the previous sentence.
// The result is bit-identical to evaluating every edge (the reference's EdgeCache prunes on the same principle, using
// the previous pixel's distances instead of a bounding circle).
template <int MODE, bool LITERAL = false>
__device__ __forceinline__ void sel_add_edge(Sel<MODE> &sel, const SegRec &s, int recIndex, d2 p, double *lim) {
    // lim[k] = min(cap[k], |current minimum of channel k|), kept up to date by this function; mask is uniform across the
    // warp (every thread walks the same record), so the branches below do not diverge
    const int mask = (MODE == 1 || MODE == 2) ? 1 : s.color;
    double need = 0;
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
        if (mask&(1<<k))
            need = dmax(need, lim[k]);
    d2 pc = p-s.bc;
    const double dsq = dot(pc, pc);
    const double reach = s.br+need;
    const bool culled = dsq > reach*reach*(1+1e-9);
    double rd = 0;
    if (!culled) {
        SDist r = signed_distance<LITERAL>(s, p, need);
        rd = r.d;
        if (fabs(r.d) <= need) {
            int near = (recIndex<<2)|r.side;
            #pragma unroll
            for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
                if (mask&(1<<k))
                    if (sd_less(r.d, r.dot, sel.ch[k].d, sel.ch[k].dot)) {
                        sel.ch[k].d = r.d; sel.ch[k].dot = r.dot; sel.ch[k].near = near;
                        lim[k] = dmin(lim[k], fabs(r.d));
                    }
        }
    }
    if (MODE == 1)
        return;
    // Perpendicular candidates (getPerpendicularDistance + addEdgePerpendicularDistance).  A candidate can only influence
    // a channel's result if |pd| <= that channel's final minimum true distance (see above), of which `need` is an upper
    // bound for all channels of this edge — so the cheap |pd| test comes first and the domain tests only for survivors.
    d2 ap = p-s.p0;
    double pdA = cross(ap, s.aDir);
    if (fabs(pdA) <= need) {
        // getPerpendicularDistance(pd, ap, -aDir) followed by pd = -pd; the two negations cancel exactly
        double m = fabs(pdA)+s.br;
        if (dot(ap, s.aBis) > 0 && -dot(ap, s.aDir) > 0 && (culled ? m*m*(1+1e-9) < dsq : fabs(pdA) < fabs(rd))) {
            #pragma unroll
            for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
                if (mask&(1<<k))
                    chan_add_perp(sel.ch[k], pdA);
        }
    }
    d2 bp = p-s.pEnd;
    double pdB = cross(bp, s.bDir);
    if (fabs(pdB) <= need) {
        double m = fabs(pdB)+s.br;
        if (-dot(bp, s.bBis) > 0 && dot(bp, s.bDir) > 0 && (culled ? m*m*(1+1e-9) < dsq : fabs(pdB) < fabs(rd))) {
            #pragma unroll
            for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
                if (mask&(1<<k))
                    chan_add_perp(sel.ch[k], pdB);
        }
    }
}

// cap[] for the records [j0, j1]: per channel, the distance from p to the nearest START POINT of an edge of that channel
// (an upper bound of the channel's minimum true distance; the tiny inflation absorbs rounding differences between this
// expression and the segment kernels').
//
// Record iteration: the walk below visits records i0..i1 of an ordered sequence.  SeqIter: the sequence is the glyph's
// whole record array.  ListIter: an ordered SUBSET of it (tile-level culling of big shapes, k1_big_blocks) — valid whenever
// every dropped record provably has no effect on any texel the list is used for.
struct SeqIter {
    __device__ __forceinline__ int at(int i) const { return i; }
    // last position of the contour that starts at position i0
    __device__ __forceinline__ int group_last(const SegRec *recs, int i0, int n) const {
        int i1 = i0;
        while (i1 < n-1 && !(recs[i1].flags&SEG_LAST_IN_CONTOUR))
            ++i1;
        return i1;
    }
};
struct ListIter {
    const int *list;
    const int *groupEnd;    // groupEnd[g] = position after the last listed record of the g-th listed contour (ascending)
    mutable int group;      // next entry of groupEnd (the walk visits the groups in order, once)
    __device__ __forceinline__ int at(int i) const { return list[i]; }
    __device__ __forceinline__ int group_last(const SegRec *, int, int) const { return groupEnd[group++]-1; }
};

template <int MODE, typename IT>
__device__ __forceinline__ void distance_caps(const SegRec *recs, IT it, int j0, int j1, d2 p, double *cap) {
    double capsq[ModeTraits<MODE>::NCH];
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
        capsq[k] = DBL_MAX;
    for (int i = j0; i <= j1; ++i) {
        const int j = it.at(i);
        const int mask = (MODE == 1 || MODE == 2) ? 1 : recs[j].color;
        d2 v = p-recs[j].p0;
        double q = dot(v, v);
        #pragma unroll
        for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
            if (mask&(1<<k))
                capsq[k] = dmin(capsq[k], q);
    }
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k)
        cap[k] = sqrt(capsq[k])*(1+1e-12);
}

template <int MODE>
__device__ __forceinline__ double resolve(const double *d) {
    return MODE >= 3 ? median3(d[0], d[1], d[2]) : d[0];
}

// Per-contour results of the overlapping combiner, parked OUTSIDE the register file while the edges are walked: the
// contour's selector state (so that the shape / inner / outer selectors can be merged afterwards exactly like the
// reference does), its resolved distances and its winding.  A small per-thread local array serves typical glyphs; shapes
// with many contours use a slice of a global scratch buffer (stride = total threads, so accesses coalesce).
#define B200_LOCAL_CONTOURS 4
template <int MODE> struct StoreLayout {
    static const int CH = MODE == 1 ? 2 : 5;                                   // doubles per channel: d, dot [, neg, pos, near]
    static const int STATE = ModeTraits<MODE>::NCH*CH;
    static const int DIST = STATE;                                             // NOUT resolved distances
    static const int WINDING = STATE+ModeTraits<MODE>::NOUT;
    static const int PER_CONTOUR = WINDING+1;
};
struct ContourStore {
    double *base;
    size_t stride;
    int perContour;
    __device__ __forceinline__ double &at(int contour, int k) { return base[(size_t(contour)*perContour+k)*stride]; }
};

template <int MODE>
__device__ __forceinline__ void store_selector(ContourStore &st, int c, const Sel<MODE> &s) {
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k) {
        const int o = k*StoreLayout<MODE>::CH;
        st.at(c, o) = s.ch[k].d;
        st.at(c, o+1) = s.ch[k].dot;
        if (MODE != 1) {
            st.at(c, o+2) = s.ch[k].neg;
            st.at(c, o+3) = s.ch[k].pos;
            st.at(c, o+4) = double(s.ch[k].near);
        }
    }
}

template <int MODE>
__device__ __forceinline__ void load_selector(ContourStore &st, int c, Sel<MODE> &s) {
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k) {
        const int o = k*StoreLayout<MODE>::CH;
        s.ch[k].d = st.at(c, o);
        s.ch[k].dot = st.at(c, o+1);
        if (MODE != 1) {
            s.ch[k].neg = st.at(c, o+2);
            s.ch[k].pos = st.at(c, o+3);
            s.ch[k].near = int(st.at(c, o+4));
        } else {
            s.ch[k].neg = -DBL_MAX; s.ch[k].pos = DBL_MAX; s.ch[k].near = -1;
        }
    }
}

// OverlappingContourCombiner::distance (contour-combiners.cpp:78-134) over the nc parked contours of `store`.
template <int MODE>
__device__ __forceinline__ void combine_overlapping(const SegRec *recs, d2 p, ContourStore store, int nc, double *out) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    typedef StoreLayout<MODE> L;
    {
        // ---- OverlappingContourCombiner::distance (contour-combiners.cpp:78-134)
        double shapeD[NOUT], innerD[NOUT], outerD[NOUT], dist[NOUT];
        {
            Sel<MODE> shape, inner, outer;
            sel_init(shape); sel_init(inner); sel_init(outer);
            for (int c = 0; c < nc; ++c) {
                Sel<MODE> sc;
                load_selector<MODE>(store, c, sc);
                double cd[NOUT];
                #pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    cd[k] = store.at(c, L::DIST+k);
                double w = store.at(c, L::WINDING);
                sel_merge<MODE>(shape, sc);
                double rd = resolve<MODE>(cd);
                if (w > 0 && rd >= 0)
                    sel_merge<MODE>(inner, sc);
                if (w < 0 && rd <= 0)
                    sel_merge<MODE>(outer, sc);
            }
            sel_distance<MODE>(shape, recs, p, shapeD);
            sel_distance<MODE>(inner, recs, p, innerD);
            sel_distance<MODE>(outer, recs, p, outerD);
        }
        double innerScalar = resolve<MODE>(innerD), outerScalar = resolve<MODE>(outerD);
        double winding;
        if (innerScalar >= 0 && fabs(innerScalar) <= fabs(outerScalar)) {
            #pragma unroll
            for (int k = 0; k < NOUT; ++k) dist[k] = innerD[k];
            winding = 1;
        } else if (outerScalar <= 0 && fabs(outerScalar) < fabs(innerScalar)) {
            #pragma unroll
            for (int k = 0; k < NOUT; ++k) dist[k] = outerD[k];
            winding = -1;
        } else {
            #pragma unroll
            for (int k = 0; k < NOUT; ++k) out[k] = shapeD[k];
            return;
        }
        // Pass 0: same-winding contours refine the inner/outer pick.  Pass 1: the others may override it.
        for (int pass = 0; pass < 2; ++pass)
            for (int c = 0; c < nc; ++c) {
                double w = store.at(c, L::WINDING);
                double cd[NOUT];
                #pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    cd[k] = store.at(c, L::DIST+k);
                double rc = resolve<MODE>(cd), rdist = resolve<MODE>(dist);
                bool take;
                if (pass == 0) {
                    if (winding > 0)
                        take = w > 0 && fabs(rc) < fabs(outerScalar) && rc > rdist;
                    else
                        take = w < 0 && fabs(rc) < fabs(innerScalar) && rc < rdist;
                } else
                    take = w != winding && rc*rdist >= 0 && fabs(rc) < fabs(rdist);
                if (take) {
                    #pragma unroll
                    for (int k = 0; k < NOUT; ++k) dist[k] = cd[k];
                }
            }
        if (resolve<MODE>(dist) == resolve<MODE>(shapeD)) {
            #pragma unroll
            for (int k = 0; k < NOUT; ++k) dist[k] = shapeD[k];
        }
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) out[k] = dist[k];
    }
}

// ShapeDistanceFinder<Combiner<Selector>>::oneShotDistance at point p over the glyph's records [0, nRec).
// recs is shared memory (small glyphs) or the global prepared buffer (big shapes); every thread of the CTA walks the same
// record at the same time, so loads broadcast.  out gets NOUT raw (unmapped) fp64 distances.
template <int MODE, bool OVERLAP, typename IT, bool LITERAL = false>
__device__ __forceinline__ void shape_distance_over(const SegRec *recs, int nRec, IT it, d2 p, ContourStore store, double *out) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    typedef StoreLayout<MODE> L;
    if (!OVERLAP) {
        Sel<MODE> sel;
        sel_init(sel);
        double cap[ModeTraits<MODE>::NCH];
        distance_caps<MODE>(recs, it, 0, nRec-1, p, cap);
        for (int i = 0; i < nRec; ++i) {
            const int j = it.at(i);
            sel_add_edge<MODE, LITERAL>(sel, recs[j], j, p, cap);
        }
        sel_distance<MODE>(sel, recs, p, out);
    } else {
        // ---- walk the edges with ONE live selector; park each finished contour
        int nc = 0;
        {
            Sel<MODE> cur;
            sel_init(cur);
            int j0 = 0;
            while (j0 < nRec) {
                // the contour's records are at positions [j0, j1] (SeqIter: its closing record carries SEG_LAST_IN_CONTOUR)
                const int j1 = it.group_last(recs, j0, nRec);
                double cap[ModeTraits<MODE>::NCH];
                distance_caps<MODE>(recs, it, j0, j1, p, cap);
                for (int i = j0; i <= j1; ++i) {
                    const int j = it.at(i);
                    sel_add_edge<MODE, LITERAL>(cur, recs[j], j, p, cap);
                }
                double cd[NOUT];
                sel_distance<MODE>(cur, recs, p, cd);
                store_selector<MODE>(store, nc, cur);
                #pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    store.at(nc, L::DIST+k) = cd[k];
                store.at(nc, L::WINDING) = double(recs[it.at(j1)].winding);
                sel_init(cur);
                ++nc;
                j0 = j1+1;
            }
        }
        combine_overlapping<MODE>(recs, p, store, nc, out);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The overlapping combiner with the shape / inner / outer selectors FOLDED while the contours are walked (big shapes).
//
// OverlappingContourCombiner::distance (contour-combiners.cpp:78-134) merges every contour's selector into three selectors and
// afterwards only needs, per contour, its resolved distance(s) and winding.  Parking whole selector states per contour and texel
// (the small-glyph kernels do, for <= 4 contours in local memory) costs 32-150 bytes per contour and texel; a 100-contour shape
// at 4096x4096 then moves > 100 GB through L2/HBM.  Here a finished contour is merged at once into three accumulators held in
// registers and only its resolved distances (+ one byte of winding) are parked: 9 bytes per contour and texel for an SDF.
// A merged channel's computeDistance needs the nearest edge only through dd = the edge's distance after
// distanceToPerpendicularDistance, which depends on (d, edge, p) alone — so it is evaluated once per contour and travels with
// the minimum through the merges (same argument as combine_parked in distance_filter.cuh).
struct FoldStore {
    double *base;           // resolved distances: NOUT per contour
    signed char *wbase;     // windings
    size_t stride;          // element stride between consecutive contours' slots of one thread (1 = private arrays)
    int perContour;
    __device__ __forceinline__ double &cd(int c, int k) { return base[(size_t(c)*perContour+k)*stride]; }
    __device__ __forceinline__ signed char &w(int c) { return wbase[size_t(c)*stride]; }
};

template <int MODE> struct FoldAcc {
    double d[ModeTraits<MODE>::NCH], dot[ModeTraits<MODE>::NCH], dd[ModeTraits<MODE>::NCH], neg[ModeTraits<MODE>::NCH], pos[ModeTraits<MODE>::NCH];
};

template <int MODE>
__device__ __forceinline__ void fold_init(FoldAcc<MODE> &a) {
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k) { a.d[k] = -DBL_MAX; a.dot[k] = 0; a.dd[k] = DBL_MAX; a.neg[k] = -DBL_MAX; a.pos[k] = DBL_MAX; }
}

template <int MODE>
__device__ __forceinline__ void fold_merge(FoldAcc<MODE> &a, const Sel<MODE> &s, const double *dd) {
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NCH; ++k) {
        if (sd_less(s.ch[k].d, s.ch[k].dot, a.d[k], a.dot[k])) { a.d[k] = s.ch[k].d; a.dot[k] = s.ch[k].dot; a.dd[k] = dd[k]; }
        if (MODE != 1) {
            if (s.ch[k].neg > a.neg[k]) a.neg[k] = s.ch[k].neg;
            if (s.ch[k].pos < a.pos[k]) a.pos[k] = s.ch[k].pos;
        }
    }
}

template <int MODE>
__device__ __forceinline__ void fold_distance(const FoldAcc<MODE> &a, double *out) {
    const int NCH = ModeTraits<MODE>::NCH;
    double td = a.d[0], tdot = a.dot[0];
    #pragma unroll
    for (int k = 0; k < NCH; ++k) {
        double v = a.d[k];
        if (MODE != 1) {
            const double m = a.d[k] < 0 ? a.neg[k] : a.pos[k];
            v = fabs(a.dd[k]) < fabs(m) ? a.dd[k] : m;
        }
        out[k] = v;
        if (MODE == 4 && k > 0 && sd_less(a.d[k], a.dot[k], td, tdot)) { td = a.d[k]; tdot = a.dot[k]; }
    }
    if (MODE == 4)
        out[3] = td;
}

// the nearest edge's distance after distanceToPerpendicularDistance (edge-segments.cpp:28-52); DBL_MAX = no nearest edge
__device__ __noinline__ double near_edge_dd(double cd, int near, const SegRec *recs, double px, double py) {
    if (near < 0)
        return DBL_MAX;
    double dd = cd;
    const int side = near&3;
    if (side) {
        const d2 p = mk(px, py);
        const SegRec &s = recs[near>>2];
        if (side == 1) {
            d2 aq = p-s.p0;
            if (dot(aq, s.aDirNZ) < 0) {
                double pd = cross(aq, s.aDirNZ);
                if (fabs(pd) <= fabs(dd))
                    dd = pd;
            }
        } else {
            d2 bq = p-s.pEnd;
            if (dot(bq, s.bDirNZ) > 0) {
                double pd = cross(bq, s.bDirNZ);
                if (fabs(pd) <= fabs(dd))
                    dd = pd;
            }
        }
    }
    return dd;
}

// Second half of OverlappingContourCombiner::distance (contour-combiners.cpp:94-134) over the folded selectors and the parked
// contour distances.
template <int MODE>
__device__ __forceinline__ void fold_finish(const FoldAcc<MODE> &shape, const FoldAcc<MODE> &inner, const FoldAcc<MODE> &outer, FoldStore st, int nc,
                                            double *out) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    double shapeD[NOUT], innerD[NOUT], outerD[NOUT], dist[NOUT];
    fold_distance<MODE>(shape, shapeD);
    fold_distance<MODE>(inner, innerD);
    fold_distance<MODE>(outer, outerD);
    const double innerScalar = resolve<MODE>(innerD), outerScalar = resolve<MODE>(outerD);
    int winding;
    if (innerScalar >= 0 && fabs(innerScalar) <= fabs(outerScalar)) {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) dist[k] = innerD[k];
        winding = 1;
    } else if (outerScalar <= 0 && fabs(outerScalar) < fabs(innerScalar)) {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) dist[k] = outerD[k];
        winding = -1;
    } else {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) out[k] = shapeD[k];
        return;
    }
    // Pass 0: same-winding contours refine the inner/outer pick.  Pass 1: the others may override it.
    for (int pass = 0; pass < 2; ++pass)
        for (int c = 0; c < nc; ++c) {
            const int w = st.w(c);
            // (the winding test first: it spares the loads of the distances of half the contours)
            if (pass == 0 ? w != winding : w == winding)
                continue;
            double cd[NOUT];
            #pragma unroll
            for (int k = 0; k < NOUT; ++k)
                cd[k] = st.cd(c, k);
            const double rc = resolve<MODE>(cd), rdist = resolve<MODE>(dist);
            bool take;
            if (pass == 0) {
                if (winding > 0)
                    take = fabs(rc) < fabs(outerScalar) && rc > rdist;
                else
                    take = fabs(rc) < fabs(innerScalar) && rc < rdist;
            } else
                take = rc*rdist >= 0 && fabs(rc) < fabs(rdist);
            if (take) {
                #pragma unroll
                for (int k = 0; k < NOUT; ++k) dist[k] = cd[k];
            }
        }
    if (resolve<MODE>(dist) == resolve<MODE>(shapeD)) {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) dist[k] = shapeD[k];
    }
    #pragma unroll
    for (int k = 0; k < NOUT; ++k) out[k] = dist[k];
}

template <int MODE, typename IT>
__device__ __forceinline__ void shape_distance_fold(const SegRec *recs, int nRec, IT it, d2 p, FoldStore st, double *out) {
    const int NCH = ModeTraits<MODE>::NCH;
    const int NOUT = ModeTraits<MODE>::NOUT;
    FoldAcc<MODE> shape, inner, outer;
    fold_init(shape); fold_init(inner); fold_init(outer);
    int nc = 0;
    {
        Sel<MODE> cur;
        sel_init(cur);
        int j0 = 0;
        while (j0 < nRec) {
            const int j1 = it.group_last(recs, j0, nRec);
            double cap[NCH];
            distance_caps<MODE>(recs, it, j0, j1, p, cap);
            for (int i = j0; i <= j1; ++i) {
                const int j = it.at(i);
                sel_add_edge<MODE>(cur, recs[j], j, p, cap);
            }
            // the contour's own distances (computeDistance per channel) ...
            double dd[NCH], cd[NOUT];
            double td = cur.ch[0].d, tdot = cur.ch[0].dot;
            #pragma unroll
            for (int k = 0; k < NCH; ++k) {
                cd[k] = cur.ch[k].d;
                dd[k] = DBL_MAX;
                if (MODE != 1) {
                    dd[k] = near_edge_dd(cur.ch[k].d, cur.ch[k].near, recs, p.x, p.y);
                    const double m = cur.ch[k].d < 0 ? cur.ch[k].neg : cur.ch[k].pos;
                    cd[k] = fabs(dd[k]) < fabs(m) ? dd[k] : m;
                }
                if (MODE == 4 && k > 0 && sd_less(cur.ch[k].d, cur.ch[k].dot, td, tdot)) { td = cur.ch[k].d; tdot = cur.ch[k].dot; }
            }
            if (MODE == 4)
                cd[3] = td;
            // ... its merges (contour-combiners.cpp:84-92) ...
            const int w = recs[it.at(j1)].winding;
            const double rd = resolve<MODE>(cd);
            fold_merge<MODE>(shape, cur, dd);
            if (w > 0 && rd >= 0)
                fold_merge<MODE>(inner, cur, dd);
            if (w < 0 && rd <= 0)
                fold_merge<MODE>(outer, cur, dd);
            // ... and what the second half of the combiner reads back
            #pragma unroll
            for (int k = 0; k < NOUT; ++k)
                st.cd(nc, k) = cd[k];
            st.w(nc) = (signed char) w;
            sel_init(cur);
            ++nc;
            j0 = j1+1;
        }
    }
    fold_finish<MODE>(shape, inner, outer, st, nc, out);
}

template <int MODE, bool OVERLAP, bool LITERAL = false>
__device__ __forceinline__ void shape_distance(const SegRec *recs, int nRec, d2 p, ContourStore store, double *out) {
    shape_distance_over<MODE, OVERLAP, SeqIter, LITERAL>(recs, nRec, SeqIter(), p, store, out);
}

}  // namespace b200
