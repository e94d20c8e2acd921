/*
 * oracle/msdf_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C99, single-threaded, cache-less CPU restatement of the reference hot path (Chlumsky/msdfgen v1.13.0):
 * per-pixel nearest signed / perpendicular / multi-channel distance over a flat list of edge segments, the
 * overlapping-contour resolve, and the MSDF error-correction pass.  Every function cites the reference
 * file:line (relative to /root/reference/) it restates.  It is the portable checker: tests/, the smoke() entry and
 * bench.py's cpu_baseline leg are the ONLY callers; the product (msdfgen_b200/csrc) never links or loads it.
 *
 * Pinning: the reference ships no golden vectors, so this file is pinned against outputs of the
 * reference itself — live against oracle/_ref/libmsdfgen_ref.so (tests/test_oracle_vs_ref.py, bit-exact on floats)
 * and against the committed fixtures in tests/golden/ that oracle/_ref produced (tests/golden/make_golden.py).
 *
 * Differences from the reference in structure (not in results): no EdgeCache pruning — every pixel starts from the
 * selector constructor state, which is ShapeDistanceFinder::oneShotDistance semantics (core/ShapeDistanceFinder.hpp:36-58);
 * shapes arrive flattened (points[nSeg][8], orders, colors, contour offsets) instead of as pointer-chasing objects.
 * Build: gcc -O2 -ffp-contract=off -std=c99 (no FMA contraction, so fp64 expressions round exactly like the reference's).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { double x, y; } v2;

/* ---- Vector2 (core/Vector2.hpp) ---- */
static v2 V(double x, double y) { v2 r; r.x = x; r.y = y; return r; }
static v2 vadd(v2 a, v2 b) { return V(a.x+b.x, a.y+b.y); }
static v2 vsub(v2 a, v2 b) { return V(a.x-b.x, a.y-b.y); }
static v2 vneg(v2 a) { return V(-a.x, -a.y); }
static v2 smul(double s, v2 a) { return V(s*a.x, s*a.y); }
static double vdot(v2 a, v2 b) { return a.x*b.x+a.y*b.y; }          /* Vector2.hpp:103 */
static double vcross(v2 a, v2 b) { return a.x*b.y-a.y*b.x; }        /* Vector2.hpp:107 */
static double vlen(v2 a) { return sqrt(a.x*a.x+a.y*a.y); }          /* Vector2.hpp:37 */
static int vzero(v2 a) { return !a.x && !a.y; }
/* Vector2::normalize(allowZero), Vector2.hpp:42-46 */
static v2 vnorm(v2 a, int allowZero) {
    double len = vlen(a);
    if (len != 0)
        return V(a.x/len, a.y/len);
    return V(0, allowZero ? 0 : 1);
}
/* Vector2::getOrthonormal(polarity, allowZero=false), Vector2.hpp:53-57 */
static v2 vorthonormal(v2 a, int polarity) {
    double len = vlen(a);
    if (len != 0)
        return polarity ? V(-a.y/len, a.x/len) : V(a.y/len, -a.x/len);
    return polarity ? V(0, 1) : V(0, -1);
}
/* mix(), arithmetics.hpp:28-31 with S = double */
static v2 vmix(v2 a, v2 b, double w) { return V((1.0-w)*a.x+w*b.x, (1.0-w)*a.y+w*b.y); }
static int non_zero_sign(double n) { return 2*(n > 0)-1; }           /* arithmetics.hpp:59-61 */
static double dmin(double a, double b) { return b < a ? b : a; }     /* arithmetics.hpp:10-13 */
static double dmax(double a, double b) { return a < b ? b : a; }     /* arithmetics.hpp:16-19 */
static double dmedian(double a, double b, double c) { return dmax(dmin(a, b), dmin(dmax(a, b), c)); } /* arithmetics.hpp:22-25 */
static float fmin_(float a, float b) { return b < a ? b : a; }
static float fmax_(float a, float b) { return a < b ? b : a; }
static float fmedian(float a, float b, float c) { return fmax_(fmin_(a, b), fmin_(fmax_(a, b), c)); }

/* ---- equation solver (core/equation-solver.cpp) ---- */
int oracle_solve_quadratic(double x[2], double a, double b, double c) {   /* equation-solver.cpp:9-32 */
    if (a == 0 || fabs(b) > 1e12*fabs(a)) {
        if (b == 0) {
            if (c == 0)
                return -1;
            return 0;
        }
        x[0] = -c/b;
        return 1;
    }
    double dscr = b*b-4*a*c;
    if (dscr > 0) {
        dscr = sqrt(dscr);
        x[0] = (-b+dscr)/(2*a);
        x[1] = (-b-dscr)/(2*a);
        return 2;
    } else if (dscr == 0) {
        x[0] = -b/(2*a);
        return 1;
    }
    return 0;
}

static int solve_cubic_normed(double x[3], double a, double b, double c) {  /* equation-solver.cpp:34-61 */
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
        t = acos(t);
        q = -2*sqrt(q);
        x[0] = q*cos(1/3.*t)-a;
        x[1] = q*cos(1/3.*(t+2*M_PI))-a;
        x[2] = q*cos(1/3.*(t-2*M_PI))-a;
        return 3;
    } else {
        double u = (r < 0 ? 1 : -1)*pow(fabs(r)+sqrt(r2-q3), 1/3.);
        double v = u == 0 ? 0 : q/u;
        x[0] = (u+v)-a;
        if (u == v || fabs(u-v) < 1e-12*fabs(u+v)) {
            x[1] = -.5*(u+v)-a;
            return 2;
        }
        return 1;
    }
}

int oracle_solve_cubic(double x[3], double a, double b, double c, double d) {  /* equation-solver.cpp:63-70 */
    if (a != 0) {
        double bn = b/a;
        if (fabs(bn) < 1e6)
            return solve_cubic_normed(x, bn, c/a, d/a);
    }
    return oracle_solve_quadratic(x, b, c, d);
}

/* ---- edge segments (core/edge-segments.cpp); a segment is (order, p[0..order]) ---- */
static v2 seg_point(int order, const v2 *p, double t) {   /* edge-segments.cpp:108-119 */
    if (order == 1)
        return vmix(p[0], p[1], t);
    if (order == 2)
        return vmix(vmix(p[0], p[1], t), vmix(p[1], p[2], t), t);
    v2 p12 = vmix(p[1], p[2], t);
    return vmix(vmix(vmix(p[0], p[1], t), p12, t), vmix(p12, vmix(p[2], p[3], t), t), t);
}

static v2 seg_direction(int order, const v2 *p, double t) {   /* edge-segments.cpp:121-139 */
    if (order == 1)
        return vsub(p[1], p[0]);
    if (order == 2) {
        v2 tangent = vmix(vsub(p[1], p[0]), vsub(p[2], p[1]), t);
        if (vzero(tangent))
            return vsub(p[2], p[0]);
        return tangent;
    }
    v2 tangent = vmix(vmix(vsub(p[1], p[0]), vsub(p[2], p[1]), t), vmix(vsub(p[2], p[1]), vsub(p[3], p[2]), t), t);
    if (vzero(tangent)) {
        if (t == 0) return vsub(p[2], p[0]);
        if (t == 1) return vsub(p[3], p[1]);
    }
    return tangent;
}

typedef struct { double d, dot; } sdist;   /* SignedDistance, core/SignedDistance.hpp:11-36 */
static int sd_less(sdist a, sdist b) {     /* SignedDistance.hpp:22-24 */
    return fabs(a.d) < fabs(b.d) || (fabs(a.d) == fabs(b.d) && a.dot < b.dot);
}

static sdist linear_sd(const v2 *p, v2 origin, double *param) {   /* edge-segments.cpp:173-185 */
    v2 aq = vsub(origin, p[0]);
    v2 ab = vsub(p[1], p[0]);
    *param = vdot(aq, ab)/vdot(ab, ab);
    v2 eq = vsub(p[*param > .5], origin);
    double endpointDistance = vlen(eq);
    sdist r;
    if (*param > 0 && *param < 1) {
        double orthoDistance = vdot(vorthonormal(ab, 0), aq);
        if (fabs(orthoDistance) < endpointDistance) {
            r.d = orthoDistance; r.dot = 0;
            return r;
        }
    }
    r.d = non_zero_sign(vcross(aq, ab))*endpointDistance;
    r.dot = fabs(vdot(vnorm(ab, 0), vnorm(eq, 0)));
    return r;
}

static sdist quadratic_sd(const v2 *p, v2 origin, double *param) {   /* edge-segments.cpp:187-226 */
    v2 qa = vsub(p[0], origin);
    v2 ab = vsub(p[1], p[0]);
    v2 br = vsub(vsub(p[2], p[1]), ab);
    double a = vdot(br, br);
    double b = 3*vdot(ab, br);
    double c = 2*vdot(ab, ab)+vdot(qa, br);
    double d = vdot(qa, ab);
    double t[3];
    int solutions = oracle_solve_cubic(t, a, b, c, d);

    v2 epDir = seg_direction(2, p, 0);
    double minDistance = non_zero_sign(vcross(epDir, qa))*vlen(qa);
    *param = -vdot(qa, epDir)/vdot(epDir, epDir);
    {
        double distance = vlen(vsub(p[2], origin));
        if (distance < fabs(minDistance)) {
            epDir = seg_direction(2, p, 1);
            minDistance = non_zero_sign(vcross(epDir, vsub(p[2], origin)))*distance;
            *param = vdot(vsub(origin, p[1]), epDir)/vdot(epDir, epDir);
        }
    }
    for (int i = 0; i < solutions; ++i) {
        if (t[i] > 0 && t[i] < 1) {
            v2 qe = vadd(vadd(qa, smul(2*t[i], ab)), smul(t[i]*t[i], br));
            double distance = vlen(qe);
            if (distance <= fabs(minDistance)) {
                minDistance = non_zero_sign(vcross(vadd(ab, smul(t[i], br)), qe))*distance;
                *param = t[i];
            }
        }
    }
    sdist r;
    r.d = minDistance;
    if (*param >= 0 && *param <= 1)
        r.dot = 0;
    else if (*param < .5)
        r.dot = fabs(vdot(vnorm(seg_direction(2, p, 0), 0), vnorm(qa, 0)));
    else
        r.dot = fabs(vdot(vnorm(seg_direction(2, p, 1), 0), vnorm(vsub(p[2], origin), 0)));
    return r;
}

#define CUBIC_SEARCH_STARTS 4   /* edge-segments.h:11 */
#define CUBIC_SEARCH_STEPS 4    /* edge-segments.h:12 */

static sdist cubic_sd(const v2 *p, v2 origin, double *param) {   /* edge-segments.cpp:228-277 */
    v2 qa = vsub(p[0], origin);
    v2 ab = vsub(p[1], p[0]);
    v2 br = vsub(vsub(p[2], p[1]), ab);
    v2 as = vsub(vsub(vsub(p[3], p[2]), vsub(p[2], p[1])), br);

    v2 epDir = seg_direction(3, p, 0);
    double minDistance = non_zero_sign(vcross(epDir, qa))*vlen(qa);
    *param = -vdot(qa, epDir)/vdot(epDir, epDir);
    {
        double distance = vlen(vsub(p[3], origin));
        if (distance < fabs(minDistance)) {
            epDir = seg_direction(3, p, 1);
            minDistance = non_zero_sign(vcross(epDir, vsub(p[3], origin)))*distance;
            *param = vdot(vsub(epDir, vsub(p[3], origin)), epDir)/vdot(epDir, epDir);
        }
    }
    for (int i = 0; i <= CUBIC_SEARCH_STARTS; ++i) {
        double t = 1./CUBIC_SEARCH_STARTS*i;
        v2 qe = vadd(vadd(vadd(qa, smul(3*t, ab)), smul(3*t*t, br)), smul(t*t*t, as));
        v2 d1 = vadd(vadd(smul(3, ab), smul(6*t, br)), smul(3*t*t, as));
        v2 d2 = vadd(smul(6, br), smul(6*t, as));
        double improvedT = t-vdot(qe, d1)/(vdot(d1, d1)+vdot(qe, d2));
        if (improvedT > 0 && improvedT < 1) {
            int remainingSteps = CUBIC_SEARCH_STEPS;
            do {
                t = improvedT;
                qe = vadd(vadd(vadd(qa, smul(3*t, ab)), smul(3*t*t, br)), smul(t*t*t, as));
                d1 = vadd(vadd(smul(3, ab), smul(6*t, br)), smul(3*t*t, as));
                if (!--remainingSteps)
                    break;
                d2 = vadd(smul(6, br), smul(6*t, as));
                improvedT = t-vdot(qe, d1)/(vdot(d1, d1)+vdot(qe, d2));
            } while (improvedT > 0 && improvedT < 1);
            double distance = vlen(qe);
            if (distance < fabs(minDistance)) {
                minDistance = non_zero_sign(vcross(d1, qe))*distance;
                *param = t;
            }
        }
    }
    sdist r;
    r.d = minDistance;
    if (*param >= 0 && *param <= 1)
        r.dot = 0;
    else if (*param < .5)
        r.dot = fabs(vdot(vnorm(seg_direction(3, p, 0), 0), vnorm(qa, 0)));
    else
        r.dot = fabs(vdot(vnorm(seg_direction(3, p, 1), 0), vnorm(vsub(p[3], origin), 0)));
    return r;
}

static sdist seg_signed_distance(int order, const v2 *p, v2 origin, double *param) {
    return order == 1 ? linear_sd(p, origin, param) : order == 2 ? quadratic_sd(p, origin, param) : cubic_sd(p, origin, param);
}

/* EdgeSegment::distanceToPerpendicularDistance, edge-segments.cpp:28-52 */
static void seg_to_perpendicular(int order, const v2 *p, sdist *distance, v2 origin, double param) {
    if (param < 0) {
        v2 dir = vnorm(seg_direction(order, p, 0), 0);
        v2 aq = vsub(origin, seg_point(order, p, 0));
        double ts = vdot(aq, dir);
        if (ts < 0) {
            double pd = vcross(aq, dir);
            if (fabs(pd) <= fabs(distance->d)) {
                distance->d = pd;
                distance->dot = 0;
            }
        }
    } else if (param > 1) {
        v2 dir = vnorm(seg_direction(order, p, 1), 0);
        v2 bq = vsub(origin, seg_point(order, p, 1));
        double ts = vdot(bq, dir);
        if (ts > 0) {
            double pd = vcross(bq, dir);
            if (fabs(pd) <= fabs(distance->d)) {
                distance->d = pd;
                distance->dot = 0;
            }
        }
    }
}

/* ---- flat shape ---- */
typedef struct {
    int nContours;
    const uint32_t *cbegin;   /* [nContours+1] */
    const v2 *pts;            /* [nSeg][4] */
    const uint8_t *orders;
    const uint8_t *colors;    /* may be NULL => WHITE */
} flat_shape;

static double shoelace(v2 a, v2 b) { return (b.x-a.x)*(a.y+b.y); }   /* Contour.cpp:8-10 */

/* Contour::winding, Contour.cpp:57-81 */
static int contour_winding(const flat_shape *s, int c) {
    uint32_t b = s->cbegin[c], e = s->cbegin[c+1];
    uint32_t n = e-b;
    if (n == 0)
        return 0;
    double total = 0;
    if (n == 1) {
        const v2 *p = s->pts+4*b;
        int o = s->orders[b];
        v2 a = seg_point(o, p, 0), bb = seg_point(o, p, 1/3.), cc = seg_point(o, p, 2/3.);
        total += shoelace(a, bb);
        total += shoelace(bb, cc);
        total += shoelace(cc, a);
    } else if (n == 2) {
        v2 a = seg_point(s->orders[b], s->pts+4*b, 0), bb = seg_point(s->orders[b], s->pts+4*b, .5);
        v2 cc = seg_point(s->orders[b+1], s->pts+4*(b+1), 0), d = seg_point(s->orders[b+1], s->pts+4*(b+1), .5);
        total += shoelace(a, bb);
        total += shoelace(bb, cc);
        total += shoelace(cc, d);
        total += shoelace(d, a);
    } else {
        v2 prev = seg_point(s->orders[e-1], s->pts+4*(e-1), 0);
        for (uint32_t i = b; i < e; ++i) {
            v2 cur = seg_point(s->orders[i], s->pts+4*i, 0);
            total += shoelace(prev, cur);
            prev = cur;
        }
    }
    return (0 < total)-(total < 0);
}

int oracle_windings(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int *out) {
    flat_shape s = { nContours, cbegin, (const v2 *) points, orders, NULL };
    for (int c = 0; c < nContours; ++c)
        out[c] = contour_winding(&s, c);
    return 0;
}

/* ---- edge selectors (core/edge-selectors.cpp), cache removed ---- */
typedef struct {            /* PerpendicularDistanceSelectorBase state, edge-selectors.h:60-66 */
    sdist minTrue;
    double minNeg, minPos;
    int nearEdge;           /* flat segment index, -1 = NULL */
    double nearParam;
} chan_sel;

static void chan_init(chan_sel *c) {   /* ctor edge-selectors.cpp:54 + SignedDistance.hpp:17 */
    c->minTrue.d = -DBL_MAX; c->minTrue.dot = 0;
    c->minNeg = -DBL_MAX; c->minPos = DBL_MAX;
    c->nearEdge = -1; c->nearParam = 0;
}
static void chan_add_true(chan_sel *c, int edge, sdist d, double param) {   /* edge-selectors.cpp:81-87 */
    if (sd_less(d, c->minTrue)) { c->minTrue = d; c->nearEdge = edge; c->nearParam = param; }
}
static void chan_add_perp(chan_sel *c, double d) {   /* edge-selectors.cpp:89-94 */
    if (d <= 0 && d > c->minNeg) c->minNeg = d;
    if (d >= 0 && d < c->minPos) c->minPos = d;
}
static void chan_merge(chan_sel *c, const chan_sel *o) {   /* edge-selectors.cpp:96-106 */
    if (sd_less(o->minTrue, c->minTrue)) { c->minTrue = o->minTrue; c->nearEdge = o->nearEdge; c->nearParam = o->nearParam; }
    if (o->minNeg > c->minNeg) c->minNeg = o->minNeg;
    if (o->minPos < c->minPos) c->minPos = o->minPos;
}
static double chan_distance(const chan_sel *c, const flat_shape *s, v2 p) {   /* computeDistance, edge-selectors.cpp:108-117 */
    double m = c->minTrue.d < 0 ? c->minNeg : c->minPos;
    if (c->nearEdge >= 0) {
        sdist d = c->minTrue;
        seg_to_perpendicular(s->orders[c->nearEdge], s->pts+4*(size_t) c->nearEdge, &d, p, c->nearParam);
        if (fabs(d.d) < fabs(m))
            m = d.d;
    }
    return m;
}
/* getPerpendicularDistance, edge-selectors.cpp:42-52 */
static int get_perp(double *distance, v2 ep, v2 edgeDir) {
    double ts = vdot(ep, edgeDir);
    if (ts > 0) {
        double pd = vcross(ep, edgeDir);
        if (fabs(pd) < fabs(*distance)) { *distance = pd; return 1; }
    }
    return 0;
}

/* mode: 1 SDF (TrueDistanceSelector), 2 PSDF (PerpendicularDistanceSelector), 3 MSDF (Multi), 4 MTSDF (MultiAndTrue) */
typedef struct { chan_sel ch[3]; } selector;    /* ch[0] only for modes 1,2; SDF uses just ch[0].minTrue */

static void sel_init(selector *s) { chan_init(&s->ch[0]); chan_init(&s->ch[1]); chan_init(&s->ch[2]); }

/* {True,Perpendicular,Multi}DistanceSelector::addEdge with every edge relevant (edge-selectors.cpp:19-29,129-160,174-227) */
static void sel_add_edge(selector *sel, int mode, const flat_shape *s, int prev, int cur, int next, v2 p) {
    const v2 *cp = s->pts+4*(size_t) cur;
    int order = s->orders[cur];
    int color = s->colors ? s->colors[cur] : 7;
    double param;
    sdist d = seg_signed_distance(order, cp, p, &param);
    if (mode == 1) {
        if (sd_less(d, sel->ch[0].minTrue))
            sel->ch[0].minTrue = d;
        return;
    }
    int mask = mode == 2 ? 1 : color;    /* PSDF ignores colours */
    for (int k = 0; k < 3; ++k)
        if (mask&(1<<k))
            chan_add_true(&sel->ch[k], cur, d, param);
    v2 ap = vsub(p, seg_point(order, cp, 0));
    v2 bp = vsub(p, seg_point(order, cp, 1));
    v2 aDir = vnorm(seg_direction(order, cp, 0), 1);
    v2 bDir = vnorm(seg_direction(order, cp, 1), 1);
    v2 prevDir = vnorm(seg_direction(s->orders[prev], s->pts+4*(size_t) prev, 1), 1);
    v2 nextDir = vnorm(seg_direction(s->orders[next], s->pts+4*(size_t) next, 0), 1);
    double add = vdot(ap, vnorm(vadd(prevDir, aDir), 1));
    double bdd = -vdot(bp, vnorm(vadd(bDir, nextDir), 1));
    if (add > 0) {
        double pd = d.d;
        if (get_perp(&pd, ap, vneg(aDir))) {
            pd = -pd;
            for (int k = 0; k < 3; ++k)
                if (mask&(1<<k))
                    chan_add_perp(&sel->ch[k], pd);
        }
    }
    if (bdd > 0) {
        double pd = d.d;
        if (get_perp(&pd, bp, bDir))
            for (int k = 0; k < 3; ++k)
                if (mask&(1<<k))
                    chan_add_perp(&sel->ch[k], pd);
    }
}

static void sel_merge(selector *a, const selector *b, int mode) {
    if (mode == 1) {    /* TrueDistanceSelector::merge, edge-selectors.cpp:31-34 */
        if (sd_less(b->ch[0].minTrue, a->ch[0].minTrue))
            a->ch[0].minTrue = b->ch[0].minTrue;
        return;
    }
    int n = mode == 2 ? 1 : 3;
    for (int k = 0; k < n; ++k)
        chan_merge(&a->ch[k], &b->ch[k]);
}

/* selector.distance() -> up to 4 doubles (edge-selectors.cpp:36-38,162-164,235-260) */
static void sel_distance(const selector *sel, int mode, const flat_shape *s, v2 p, double *out) {
    if (mode == 1) { out[0] = sel->ch[0].minTrue.d; return; }
    if (mode == 2) { out[0] = chan_distance(&sel->ch[0], s, p); return; }
    out[0] = chan_distance(&sel->ch[0], s, p);
    out[1] = chan_distance(&sel->ch[1], s, p);
    out[2] = chan_distance(&sel->ch[2], s, p);
    if (mode == 4) {    /* MultiDistanceSelector::trueDistance, edge-selectors.cpp:243-250 */
        sdist t = sel->ch[0].minTrue;
        if (sd_less(sel->ch[1].minTrue, t)) t = sel->ch[1].minTrue;
        if (sd_less(sel->ch[2].minTrue, t)) t = sel->ch[2].minTrue;
        out[3] = t.d;
    }
}

static int mode_channels(int mode) { return mode == 3 ? 3 : mode == 4 ? 4 : 1; }
static double resolve(const double *d, int mode) {   /* contour-combiners.cpp:22-32 */
    return mode >= 3 ? dmedian(d[0], d[1], d[2]) : d[0];
}

/* Feeds one contour's edges to a selector in reference order: cur = last edge first (ShapeDistanceFinder.hpp:18-31). */
static void feed_contour(selector *sel, int mode, const flat_shape *s, int c, v2 p) {
    int b = (int) s->cbegin[c], e = (int) s->cbegin[c+1];
    int n = e-b;
    if (n <= 0)
        return;
    int prev = n >= 2 ? e-2 : b;
    int cur = e-1;
    for (int next = b; next < e; ++next) {
        sel_add_edge(sel, mode, s, prev, cur, next, p);
        prev = cur;
        cur = next;
    }
}

/* ShapeDistanceFinder<Combiner<Selector>>::oneShotDistance; out gets 1/1/3/4 raw (unmapped) distances. */
static void shape_distance(int mode, int overlap, const flat_shape *s, const int *windings, selector *contourSel, double *contourDist, v2 p, double *out) {
    int N = mode_channels(mode);
    if (!overlap) {   /* SimpleContourCombiner, contour-combiners.cpp:34-50 */
        selector sel;
        sel_init(&sel);
        for (int c = 0; c < s->nContours; ++c)
            feed_contour(&sel, mode, s, c, p);
        sel_distance(&sel, mode, s, p, out);
        return;
    }
    /* OverlappingContourCombiner::distance, contour-combiners.cpp:78-134 */
    selector shapeSel, innerSel, outerSel;
    sel_init(&shapeSel); sel_init(&innerSel); sel_init(&outerSel);
    for (int c = 0; c < s->nContours; ++c) {
        sel_init(&contourSel[c]);
        feed_contour(&contourSel[c], mode, s, c, p);
        double *cd = contourDist+4*(size_t) c;
        sel_distance(&contourSel[c], mode, s, p, cd);
        sel_merge(&shapeSel, &contourSel[c], mode);
        if (windings[c] > 0 && resolve(cd, mode) >= 0)
            sel_merge(&innerSel, &contourSel[c], mode);
        if (windings[c] < 0 && resolve(cd, mode) <= 0)
            sel_merge(&outerSel, &contourSel[c], mode);
    }
    double shapeD[4], innerD[4], outerD[4], dist[4];
    sel_distance(&shapeSel, mode, s, p, shapeD);
    sel_distance(&innerSel, mode, s, p, innerD);
    sel_distance(&outerSel, mode, s, p, outerD);
    double innerScalar = resolve(innerD, mode), outerScalar = resolve(outerD, mode);
    for (int k = 0; k < N; ++k) dist[k] = -DBL_MAX;
    int winding = 0;
    if (innerScalar >= 0 && fabs(innerScalar) <= fabs(outerScalar)) {
        memcpy(dist, innerD, sizeof(double)*N);
        winding = 1;
        for (int c = 0; c < s->nContours; ++c)
            if (windings[c] > 0) {
                const double *cd = contourDist+4*(size_t) c;
                if (fabs(resolve(cd, mode)) < fabs(outerScalar) && resolve(cd, mode) > resolve(dist, mode))
                    memcpy(dist, cd, sizeof(double)*N);
            }
    } else if (outerScalar <= 0 && fabs(outerScalar) < fabs(innerScalar)) {
        memcpy(dist, outerD, sizeof(double)*N);
        winding = -1;
        for (int c = 0; c < s->nContours; ++c)
            if (windings[c] < 0) {
                const double *cd = contourDist+4*(size_t) c;
                if (fabs(resolve(cd, mode)) < fabs(innerScalar) && resolve(cd, mode) < resolve(dist, mode))
                    memcpy(dist, cd, sizeof(double)*N);
            }
    } else {
        memcpy(out, shapeD, sizeof(double)*N);
        return;
    }
    for (int c = 0; c < s->nContours; ++c)
        if (windings[c] != winding) {
            const double *cd = contourDist+4*(size_t) c;
            if (resolve(cd, mode)*resolve(dist, mode) >= 0 && fabs(resolve(cd, mode)) < fabs(resolve(dist, mode)))
                memcpy(dist, cd, sizeof(double)*N);
        }
    if (resolve(dist, mode) == resolve(shapeD, mode))
        memcpy(dist, shapeD, sizeof(double)*N);
    memcpy(out, dist, sizeof(double)*N);
}

typedef struct {
    flat_shape s;
    int *windings;
    selector *contourSel;
    double *contourDist;
} finder;

static int finder_init(finder *f, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors) {
    f->s.nContours = nContours; f->s.cbegin = cbegin; f->s.pts = (const v2 *) points; f->s.orders = orders; f->s.colors = colors;
    f->windings = (int *) malloc(sizeof(int)*(size_t)(nContours+1));
    f->contourSel = (selector *) malloc(sizeof(selector)*(size_t)(nContours+1));
    f->contourDist = (double *) malloc(sizeof(double)*4*(size_t)(nContours+1));
    if (!f->windings || !f->contourSel || !f->contourDist)
        return -1;
    for (int c = 0; c < nContours; ++c)   /* OverlappingContourCombiner ctor, contour-combiners.cpp:58-63 */
        f->windings[c] = contour_winding(&f->s, c);
    return 0;
}
static void finder_free(finder *f) { free(f->windings); free(f->contourSel); free(f->contourDist); }

int oracle_distance(int mode, int overlap, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors,
                    int nPts, const double *pts, double *out) {
    finder f;
    if (finder_init(&f, nContours, cbegin, points, orders, colors))
        return -1;
    int N = mode_channels(mode);
    for (int i = 0; i < nPts; ++i)
        shape_distance(mode, overlap, &f.s, f.windings, f.contourSel, f.contourDist, V(pts[2*i], pts[2*i+1]), out+(size_t) N*i);
    finder_free(&f);
    return 0;
}

int oracle_segment_distances(int nSeg, const double *points, const uint8_t *orders, int nPts, const double *pts, double *out) {
    for (int i = 0; i < nPts; ++i)
        for (int s = 0; s < nSeg; ++s) {
            double param = 0;
            sdist d = seg_signed_distance(orders[s], (const v2 *) points+4*(size_t) s, V(pts[2*i], pts[2*i+1]), &param);
            double *o = out+3*((size_t) i*nSeg+s);
            o[0] = d.d; o[1] = d.dot; o[2] = param;
        }
    return 0;
}

/* ---- transformation: xf = scale.x, scale.y, translate.x, translate.y, dmScale, dmTranslate ---- */
static v2 unproject(const double *xf, v2 c) { return V(c.x/xf[0]-xf[2], c.y/xf[1]-xf[3]); }     /* Projection.cpp:14-16 */
static v2 project(const double *xf, v2 c) { return V(xf[0]*(c.x+xf[2]), xf[1]*(c.y+xf[3])); }   /* Projection.cpp:10-12 */
static v2 unproject_vector(const double *xf, v2 v) { return V(v.x/xf[0], v.y/xf[1]); }          /* Projection.cpp:22-24 */
static double dmap(const double *xf, double d) { return xf[4]*(d+xf[5]); }                      /* DistanceMapping.cpp:15-17 */

/* ---- error correction (core/MSDFErrorCorrection.cpp, core/msdf-error-correction.cpp) ---- */
#define ARTIFACT_T_EPSILON .01
#define PROTECTION_RADIUS_TOLERANCE 1.001
#define FLAG_ERROR 1
#define FLAG_PROTECTED 2
#define CLS_CANDIDATE 1
#define CLS_ARTIFACT 2

typedef struct {
    int N, w, h;
    const float *sdf; ptrdiff_t sdfStride;     /* shape-oriented view of the field */
    uint8_t *st; ptrdiff_t stStride;           /* shape-oriented view of the stencil */
    const double *xf;
    double minDeviationRatio, minImproveRatio;
    /* exact-distance checker (NULL shape => base classifier only) */
    finder *f; int overlap;
} ec_ctx;

static const float *TX(const ec_ctx *e, int x, int y) { return e->sdf+e->sdfStride*y+(ptrdiff_t) e->N*x; }
static uint8_t *ST(const ec_ctx *e, int x, int y) { return e->st+e->stStride*y+x; }
static float mixf(float a, float b, double t) { return (float) ((1.0-t)*a+t*b); }   /* mix<float,double>, arithmetics.hpp:28-31 */

/* protectCorners, MSDFErrorCorrection.cpp:121-151 */
static void ec_protect_corners(ec_ctx *e, const flat_shape *s) {
    for (int c = 0; c < s->nContours; ++c) {
        int b = (int) s->cbegin[c], en = (int) s->cbegin[c+1];
        if (en <= b)
            continue;
        int prev = en-1;
        for (int i = b; i < en; ++i) {
            int commonColor = (s->colors ? s->colors[prev] : 7)&(s->colors ? s->colors[i] : 7);
            if (!(commonColor&(commonColor-1))) {
                v2 p = project(e->xf, seg_point(s->orders[i], s->pts+4*(size_t) i, 0));
                int l = (int) floor(p.x-.5);
                int bb = (int) floor(p.y-.5);
                int r = l+1;
                int t = bb+1;
                if (l < e->w && bb < e->h && r >= 0 && t >= 0) {
                    if (l >= 0 && bb >= 0) *ST(e, l, bb) |= FLAG_PROTECTED;
                    if (r < e->w && bb >= 0) *ST(e, r, bb) |= FLAG_PROTECTED;
                    if (l >= 0 && t < e->h) *ST(e, l, t) |= FLAG_PROTECTED;
                    if (r < e->w && t < e->h) *ST(e, r, t) |= FLAG_PROTECTED;
                }
            }
            prev = i;
        }
    }
}

/* edgeBetweenTexelsChannel / edgeBetweenTexels / protectExtremeChannels, MSDFErrorCorrection.cpp:154-187 */
static int edge_between_channel(const float *a, const float *b, int ch) {
    double t = (a[ch]-.5)/(a[ch]-b[ch]);
    if (t > 0 && t < 1) {
        float c[3] = { mixf(a[0], b[0], t), mixf(a[1], b[1], t), mixf(a[2], b[2], t) };
        return fmedian(c[0], c[1], c[2]) == c[ch];
    }
    return 0;
}
static int edge_between(const float *a, const float *b) {
    return 1*edge_between_channel(a, b, 0)+2*edge_between_channel(a, b, 1)+4*edge_between_channel(a, b, 2);
}
static void protect_extreme(uint8_t *st, const float *msd, float m, int mask) {
    if ((mask&1 && msd[0] != m) || (mask&2 && msd[1] != m) || (mask&4 && msd[2] != m))
        *st |= FLAG_PROTECTED;
}

/* protectEdges<N>, MSDFErrorCorrection.cpp:189-250 */
static void ec_protect_edges(ec_ctx *e) {
    float radius;
    radius = (float) (PROTECTION_RADIUS_TOLERANCE*vlen(unproject_vector(e->xf, V(e->xf[4]*1, 0))));
    for (int y = 0; y < e->h; ++y)
        for (int x = 0; x < e->w-1; ++x) {
            const float *l = TX(e, x, y), *r = TX(e, x+1, y);
            float lm = fmedian(l[0], l[1], l[2]), rm = fmedian(r[0], r[1], r[2]);
            if (fabsf(lm-.5f)+fabsf(rm-.5f) < radius) {
                int mask = edge_between(l, r);
                protect_extreme(ST(e, x, y), l, lm, mask);
                protect_extreme(ST(e, x+1, y), r, rm, mask);
            }
        }
    radius = (float) (PROTECTION_RADIUS_TOLERANCE*vlen(unproject_vector(e->xf, V(0, e->xf[4]*1))));
    for (int y = 0; y < e->h-1; ++y)
        for (int x = 0; x < e->w; ++x) {
            const float *b = TX(e, x, y), *t = TX(e, x, y+1);
            float bm = fmedian(b[0], b[1], b[2]), tm = fmedian(t[0], t[1], t[2]);
            if (fabsf(bm-.5f)+fabsf(tm-.5f) < radius) {
                int mask = edge_between(b, t);
                protect_extreme(ST(e, x, y), b, bm, mask);
                protect_extreme(ST(e, x, y+1), t, tm, mask);
            }
        }
    radius = (float) (PROTECTION_RADIUS_TOLERANCE*vlen(unproject_vector(e->xf, V(e->xf[4]*1, e->xf[4]*1))));
    for (int y = 0; y < e->h-1; ++y)
        for (int x = 0; x < e->w-1; ++x) {
            const float *lb = TX(e, x, y), *rb = TX(e, x+1, y), *lt = TX(e, x, y+1), *rt = TX(e, x+1, y+1);
            float mlb = fmedian(lb[0], lb[1], lb[2]), mrb = fmedian(rb[0], rb[1], rb[2]);
            float mlt = fmedian(lt[0], lt[1], lt[2]), mrt = fmedian(rt[0], rt[1], rt[2]);
            if (fabsf(mlb-.5f)+fabsf(mrt-.5f) < radius) {
                int mask = edge_between(lb, rt);
                protect_extreme(ST(e, x, y), lb, mlb, mask);
                protect_extreme(ST(e, x+1, y+1), rt, mrt, mask);
            }
            if (fabsf(mrb-.5f)+fabsf(mlt-.5f) < radius) {
                int mask = edge_between(rb, lt);
                protect_extreme(ST(e, x+1, y), rb, mrb, mask);
                protect_extreme(ST(e, x, y+1), lt, mlt, mask);
            }
        }
}

static void ec_protect_all(ec_ctx *e) {   /* MSDFErrorCorrection.cpp:252-258 */
    for (int y = 0; y < e->h; ++y)
        for (int x = 0; x < e->w; ++x)
            *ST(e, x, y) |= FLAG_PROTECTED;
}

/* classifier = BaseArtifactClassifier (f == NULL) or ShapeDistanceChecker::ArtifactClassifier (MSDFErrorCorrection.cpp:26-102) */
typedef struct {
    const ec_ctx *e;
    double span;
    int protectedFlag;
    int useShape;
    v2 direction, shapeCoord, sdfCoord;
    const float *msd;
} classifier;

static int range_test(const classifier *k, double at, double bt, double xt, float am, float bm, float xm) {   /* :30-41 */
    if ((am > .5f && bm > .5f && xm <= .5f) || (am < .5f && bm < .5f && xm >= .5f) || (!k->protectedFlag && fmedian(am, bm, xm) != xm)) {
        double axSpan = (xt-at)*k->span, bxSpan = (bt-xt)*k->span;
        if (!(xm >= am-axSpan && xm <= am+axSpan && xm >= bm-bxSpan && xm <= bm+bxSpan))
            return CLS_CANDIDATE|CLS_ARTIFACT;
        return CLS_CANDIDATE;
    }
    return 0;
}

/* interpolate(), bitmap-interpolation.hpp:10-25 (clamp(n,b): arithmetics.hpp:40-42) */
static double clampd(double n, double b) { return n >= 0 && n <= b ? n : (double) (n > 0)*b; }
static int clampi(int n, int b) { return n >= 0 && n <= b ? n : (int) (n > 0)*b; }
static void interpolate(float *out, const ec_ctx *e, v2 pos) {
    pos.x = clampd(pos.x, (double) e->w);
    pos.y = clampd(pos.y, (double) e->h);
    pos.x -= .5; pos.y -= .5;
    int l = (int) floor(pos.x);
    int b = (int) floor(pos.y);
    int r = l+1;
    int t = b+1;
    double lr = pos.x-l;
    double bt = pos.y-b;
    l = clampi(l, e->w-1), r = clampi(r, e->w-1);
    b = clampi(b, e->h-1), t = clampi(t, e->h-1);
    for (int i = 0; i < e->N; ++i)
        out[i] = mixf(mixf(TX(e, l, b)[i], TX(e, r, b)[i], lr), mixf(TX(e, l, t)[i], TX(e, r, t)[i], lr), bt);
}

/* test-only statistic: number of exact shape-distance evaluations performed by the error-correction pass */
static long long g_exact_evals = 0;
static int g_stat_no_short_circuit = 0;   /* test-only: walk every neighbour test instead of stopping at the first hit (same flags) */
long long oracle_exact_eval_count(int reset) { long long v = g_exact_evals; if (reset) g_exact_evals = 0; return v; }
void oracle_stat_no_short_circuit(int on) { g_stat_no_short_circuit = on; }

static int cls_evaluate(const classifier *k, double t, float m, int flags) {   /* :47-49 and :59-82 */
    (void) m;
    if (!k->useShape)
        return (flags&2) != 0;
    if (flags&CLS_CANDIDATE) {
        if (flags&CLS_ARTIFACT)
            return 1;
        const ec_ctx *e = k->e;
        ++g_exact_evals;
        v2 tVector = smul(t, k->direction);
        float oldMSD[4], newMSD[3];
        v2 sdfCoord = vadd(k->sdfCoord, tVector);
        interpolate(oldMSD, e, sdfCoord);
        double aWeight = (1-fabs(tVector.x))*(1-fabs(tVector.y));
        float aPSD = fmedian(k->msd[0], k->msd[1], k->msd[2]);
        newMSD[0] = (float) (oldMSD[0]+aWeight*(aPSD-k->msd[0]));
        newMSD[1] = (float) (oldMSD[1]+aWeight*(aPSD-k->msd[1]));
        newMSD[2] = (float) (oldMSD[2]+aWeight*(aPSD-k->msd[2]));
        float oldPSD = fmedian(oldMSD[0], oldMSD[1], oldMSD[2]);
        float newPSD = fmedian(newMSD[0], newMSD[1], newMSD[2]);
        v2 texelSize = unproject_vector(e->xf, V(1, 1));
        v2 q = vadd(k->shapeCoord, V(tVector.x*texelSize.x, tVector.y*texelSize.y));
        double dist;
        shape_distance(2, e->overlap, &e->f->s, e->f->windings, e->f->contourSel, e->f->contourDist, q, &dist);
        float refPSD = (float) dmap(e->xf, dist);
        return e->minImproveRatio*fabsf(newPSD-refPSD) < (double) fabsf(oldPSD-refPSD);
    }
    return 0;
}

static float interp_median_lin(const float *a, const float *b, double t) {   /* :261-267 */
    return fmedian(mixf(a[0], b[0], t), mixf(a[1], b[1], t), mixf(a[2], b[2], t));
}
static float interp_median_quad(const float *a, const float *l, const float *q, double t) {   /* :269-275 */
    return (float) dmedian(t*(t*q[0]+l[0])+a[0], t*(t*q[1]+l[1])+a[1], t*(t*q[2]+l[2])+a[2]);
}

static int linear_artifact_inner(const classifier *k, float am, float bm, const float *a, const float *b, float dA, float dB) {   /* :279-289 */
    double t = (double) dA/(dA-dB);
    if (t > ARTIFACT_T_EPSILON && t < 1-ARTIFACT_T_EPSILON) {
        float xm = interp_median_lin(a, b, t);
        return cls_evaluate(k, t, xm, range_test(k, 0, 1, t, am, bm, xm));
    }
    return 0;
}

static int diagonal_artifact_inner(const classifier *k, float am, float dm, const float *a, const float *l, const float *q, float dA, float dBC, float dD, double tEx0, double tEx1) {   /* :293-329 */
    double t[2];
    int solutions = oracle_solve_quadratic(t, dD-dBC+dA, dBC-dA-dA, dA);
    for (int i = 0; i < solutions; ++i) {
        if (t[i] > ARTIFACT_T_EPSILON && t[i] < 1-ARTIFACT_T_EPSILON) {
            float xm = interp_median_quad(a, l, q, t[i]);
            int rangeFlags = range_test(k, 0, 1, t[i], am, dm, xm);
            double tEnd[2];
            float em[2];
            if (tEx0 > 0 && tEx0 < 1) {
                tEnd[0] = 0, tEnd[1] = 1;
                em[0] = am, em[1] = dm;
                tEnd[tEx0 > t[i]] = tEx0;
                em[tEx0 > t[i]] = interp_median_quad(a, l, q, tEx0);
                rangeFlags |= range_test(k, tEnd[0], tEnd[1], t[i], em[0], em[1], xm);
            }
            if (tEx1 > 0 && tEx1 < 1) {
                tEnd[0] = 0, tEnd[1] = 1;
                em[0] = am, em[1] = dm;
                tEnd[tEx1 > t[i]] = tEx1;
                em[tEx1 > t[i]] = interp_median_quad(a, l, q, tEx1);
                rangeFlags |= range_test(k, tEnd[0], tEnd[1], t[i], em[0], em[1], xm);
            }
            if (cls_evaluate(k, t[i], xm, rangeFlags))
                return 1;
        }
    }
    return 0;
}

static int has_linear_artifact(const classifier *k, float am, const float *a, const float *b) {   /* :333-345 */
    float bm = fmedian(b[0], b[1], b[2]);
    return fabsf(am-.5f) >= fabsf(bm-.5f) && (
        linear_artifact_inner(k, am, bm, a, b, a[1]-a[0], b[1]-b[0]) ||
        linear_artifact_inner(k, am, bm, a, b, a[2]-a[1], b[2]-b[1]) ||
        linear_artifact_inner(k, am, bm, a, b, a[0]-a[2], b[0]-b[2]));
}

static int has_diagonal_artifact(const classifier *k, float am, const float *a, const float *b, const float *c, const float *d) {   /* :349-381 */
    float dm = fmedian(d[0], d[1], d[2]);
    if (fabsf(am-.5f) >= fabsf(dm-.5f)) {
        float abc[3] = { a[0]-b[0]-c[0], a[1]-b[1]-c[1], a[2]-b[2]-c[2] };
        float l[3] = { -a[0]-abc[0], -a[1]-abc[1], -a[2]-abc[2] };
        float q[3] = { d[0]+abc[0], d[1]+abc[1], d[2]+abc[2] };
        double tEx[3] = { -.5*l[0]/q[0], -.5*l[1]/q[1], -.5*l[2]/q[2] };
        return
            diagonal_artifact_inner(k, am, dm, a, l, q, a[1]-a[0], b[1]-b[0]+c[1]-c[0], d[1]-d[0], tEx[0], tEx[1]) ||
            diagonal_artifact_inner(k, am, dm, a, l, q, a[2]-a[1], b[2]-b[1]+c[2]-c[1], d[2]-d[1], tEx[1], tEx[2]) ||
            diagonal_artifact_inner(k, am, dm, a, l, q, a[0]-a[2], b[0]-b[2]+c[0]-c[2], d[0]-d[2], tEx[2], tEx[0]);
    }
    return 0;
}

/* findErrors<N>(sdf) and findErrors<Combiner,N>(sdf, shape), MSDFErrorCorrection.cpp:383-457 */
static void ec_find_errors(ec_ctx *e, int useShape) {
    double dm1 = e->xf[4]*1;   /* distanceMapping(Delta(1)) */
    double hSpan = e->minDeviationRatio*vlen(unproject_vector(e->xf, V(dm1, 0)));
    double vSpan = e->minDeviationRatio*vlen(unproject_vector(e->xf, V(0, dm1)));
    double dSpan = e->minDeviationRatio*vlen(unproject_vector(e->xf, V(dm1, dm1)));
    int w = e->w, h = e->h;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            if (useShape && (*ST(e, x, y)&FLAG_ERROR))
                continue;
            const float *c = TX(e, x, y);
            float cm = fmedian(c[0], c[1], c[2]);
            classifier k;
            k.e = e; k.useShape = useShape;
            k.protectedFlag = (*ST(e, x, y)&FLAG_PROTECTED) != 0;
            k.shapeCoord = unproject(e->xf, V(x+.5, y+.5));
            k.sdfCoord = V(x+.5, y+.5);
            k.msd = c;
            const float *l = x > 0 ? TX(e, x-1, y) : NULL, *b = y > 0 ? TX(e, x, y-1) : NULL;
            const float *r = x < w-1 ? TX(e, x+1, y) : NULL, *t = y < h-1 ? TX(e, x, y+1) : NULL;
            int err = 0;
#define LIN(cond, nb, dx, dy, sp) if ((!err || g_stat_no_short_circuit) && (cond)) { k.direction = V(dx, dy); k.span = sp; err |= has_linear_artifact(&k, cm, c, nb); }
#define DIA(cond, n1, n2, nd, dx, dy) if ((!err || g_stat_no_short_circuit) && (cond)) { k.direction = V(dx, dy); k.span = dSpan; err |= has_diagonal_artifact(&k, cm, c, n1, n2, nd); }
            LIN(x > 0, l, -1, 0, hSpan)
            LIN(y > 0, b, 0, -1, vSpan)
            LIN(x < w-1, r, +1, 0, hSpan)
            LIN(y < h-1, t, 0, +1, vSpan)
            DIA(x > 0 && y > 0, l, b, TX(e, x-1, y-1), -1, -1)
            DIA(x < w-1 && y > 0, r, b, TX(e, x+1, y-1), +1, -1)
            DIA(x > 0 && y < h-1, l, t, TX(e, x-1, y+1), -1, +1)
            DIA(x < w-1 && y < h-1, r, t, TX(e, x+1, y+1), +1, +1)
#undef LIN
#undef DIA
            if (err)
                *ST(e, x, y) |= FLAG_ERROR;
        }
}

/*
 * msdfErrorCorrectionInner<N> (msdf-error-correction.cpp:12-48) followed by apply (MSDFErrorCorrection.cpp:459-479).
 * ecMode: 0 DISABLED, 1 INDISCRIMINATE, 2 EDGE_PRIORITY, 3 EDGE_ONLY; ecDcm: 0 DO_NOT_CHECK, 1 CHECK_AT_EDGE, 2 ALWAYS_CHECK
 * (generator-config.h:19-36).  shapeless != 0 restates msdfErrorCorrectionShapeless (:51-59): ecMode 3 <=> protectAll.
 * stencilOut (w*h bytes, rows in the sdf's own orientation... see note) may be NULL.
 * The sdf is given with its own orientation flag; like the reference, all work happens in the shape's orientation.
 */
int oracle_error_correction(int N, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int shapeYDown,
                            const double *xf, float *sdf, int w, int h, int rowStride, int sdfYDown,
                            int overlap, int ecMode, int ecDcm, double minDev, double minImp, int shapeless, uint8_t *stencilOut) {
    if (N != 3 && N != 4)
        return -1;
    if (ecMode == 0)
        return 0;
    uint8_t *stencil = stencilOut ? stencilOut : (uint8_t *) malloc((size_t) w*h);
    if (!stencil)
        return -1;
    memset(stencil, 0, (size_t) w*h);
    finder f;
    int haveFinder = 0;
    ec_ctx e;
    e.N = N; e.w = w; e.h = h; e.xf = xf;
    e.minDeviationRatio = minDev; e.minImproveRatio = minImp;
    e.f = NULL; e.overlap = overlap;
    /* Orientation: the stencil BitmapSection starts Y_UPWARD (BitmapRef.hpp:82) and is reoriented to the shape
       (protectCorners :122) or to the sdf (:192,:385); the sdf is reoriented to the shape in findErrors(sdf, shape) (:414)
       and to the stencil in apply (:461).  All texel pairings are orientation-symmetric, so one consistent view works:
       rows are addressed in SHAPE orientation for both buffers (shapeless: in sdf orientation, no flip). */
    int flipSdf = shapeless ? 0 : (sdfYDown != shapeYDown);
    int flipSt = shapeless ? (sdfYDown != 0) : (shapeYDown != 0);
    e.sdf = flipSdf ? sdf+(ptrdiff_t) rowStride*(h-1) : sdf;
    e.sdfStride = flipSdf ? -(ptrdiff_t) rowStride : rowStride;
    e.st = flipSt ? stencil+(ptrdiff_t) w*(h-1) : stencil;
    e.stStride = flipSt ? -(ptrdiff_t) w : w;
    flat_shape fs = { nContours, cbegin, (const v2 *) points, orders, colors };
    switch (ecMode) {
        case 2:
            if (!shapeless)
                ec_protect_corners(&e, &fs);
            ec_protect_edges(&e);
            break;
        case 3:
            ec_protect_all(&e);
            break;
    }
    if (shapeless) {
        ec_find_errors(&e, 0);
    } else {
        if (ecDcm == 0 || (ecDcm == 1 && ecMode != 3)) {
            ec_find_errors(&e, 0);
            if (ecDcm == 1)
                ec_protect_all(&e);
        }
        if (ecDcm == 2 || ecDcm == 1) {
            if (finder_init(&f, nContours, cbegin, points, orders, colors)) {
                if (!stencilOut) free(stencil);
                return -1;
            }
            haveFinder = 1;
            e.f = &f;
            ec_find_errors(&e, 1);
        }
    }
    /* apply, MSDFErrorCorrection.cpp:459-479 */
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            if (*ST(&e, x, y)&FLAG_ERROR) {
                float *px = (float *) TX(&e, x, y);
                float m = fmedian(px[0], px[1], px[2]);
                px[0] = m, px[1] = m, px[2] = m;
            }
    if (haveFinder)
        finder_free(&f);
    if (!stencilOut)
        free(stencil);
    return 0;
}

/*
 * generate{SDF,PSDF,MSDF,MTSDF}: generateDistanceField (core/msdfgen.cpp:52-76) + msdfErrorCorrection for modes 3,4 (:92-106).
 */
int oracle_generate(int mode, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int shapeYDown,
                    const double *xf, float *out, int w, int h, int rowStride, int outYDown,
                    int overlap, int ecMode, int ecDcm, double minDev, double minImp, uint8_t *stencilOut) {
    if (mode < 1 || mode > 4)
        return -1;
    int N = mode_channels(mode);
    finder f;
    if (finder_init(&f, nContours, cbegin, points, orders, colors))
        return -1;
    int flip = outYDown != shapeYDown;    /* output.reorient(shape.getYAxisOrientation()), msdfgen.cpp:55 */
    float *base = flip ? out+(ptrdiff_t) rowStride*(h-1) : out;
    ptrdiff_t stride = flip ? -(ptrdiff_t) rowStride : rowStride;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            v2 p = unproject(xf, V(x+.5, y+.5));
            double d[4];
            shape_distance(mode, overlap, &f.s, f.windings, f.contourSel, f.contourDist, p, d);
            float *px = base+stride*y+(ptrdiff_t) N*x;
            for (int k = 0; k < N; ++k)
                px[k] = (float) dmap(xf, d[k]);
        }
    finder_free(&f);
    if (mode >= 3)
        return oracle_error_correction(N, nContours, cbegin, points, orders, colors, shapeYDown, xf, out, w, h, rowStride, outYDown,
                                       overlap, ecMode, ecDcm, minDev, minImp, 0, stencilOut);
    return 0;
}

/* Batch loop for the cpu_baseline "port" leg: same-size glyphs, packed outputs, one shared or per-glyph transform. */
int oracle_generate_batch(int mode, int nGlyphs, const uint32_t *gbegin, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors,
                          const uint8_t *yDown, const double *xf, int xfPerGlyph, float *out, int w, int h,
                          int overlap, int ecMode, int ecDcm, double minDev, double minImp) {
    int N = mode_channels(mode);
    for (int g = 0; g < nGlyphs; ++g) {
        uint32_t c0 = gbegin[g], c1 = gbegin[g+1];
        uint32_t base = cbegin[c0];
        uint32_t *local = (uint32_t *) malloc(sizeof(uint32_t)*(c1-c0+1));
        if (!local)
            return -1;
        for (uint32_t c = c0; c <= c1; ++c)
            local[c-c0] = cbegin[c]-base;
        int rc = oracle_generate(mode, (int) (c1-c0), local, points+8*(size_t) base, orders+base, colors ? colors+base : NULL, yDown ? yDown[g] : 0,
                                 xf+(xfPerGlyph ? 6*(size_t) g : 0), out+(size_t) g*w*h*N, w, h, w*N, 0, overlap, ecMode, ecDcm, minDev, minImp, NULL);
        free(local);
        if (rc)
            return rc;
    }
    return 0;
}

/* ---- scanline sign correction (core/rasterization.cpp:19-96, Scanline.cpp, edge-segments.cpp:279-403) ---- */
static int isign(double n) { return (0 < n)-(n < 0); }

/* {Linear,Quadratic,Cubic}Segment::scanlineIntersections: x positions and crossing directions of the edge on the line y */
static int seg_scanline(int order, const v2 *p, double x[3], int dy[3], double y) {
    if (order == 1) {   /* edge-segments.cpp:279-287 */
        if ((y >= p[0].y && y < p[1].y) || (y >= p[1].y && y < p[0].y)) {
            double param = (y-p[0].y)/(p[1].y-p[0].y);
            x[0] = (1.0-param)*p[0].x+param*p[1].x;
            dy[0] = isign(p[1].y-p[0].y);
            return 1;
        }
        return 0;
    }
    int total = 0;
    int nextDY = y > p[0].y ? 1 : -1;
    x[total] = p[0].x;
    if (order == 2) {   /* edge-segments.cpp:289-341 */
        if (p[0].y == y) {
            if (p[0].y < p[1].y || (p[0].y == p[1].y && p[0].y < p[2].y))
                dy[total++] = 1;
            else
                nextDY = 1;
        }
        {
            v2 ab = vsub(p[1], p[0]);
            v2 br = vsub(vsub(p[2], p[1]), ab);
            double t[2];
            int solutions = oracle_solve_quadratic(t, br.y, 2*ab.y, p[0].y-y);
            double tmp;
            if (solutions >= 2 && t[0] > t[1])
                tmp = t[0], t[0] = t[1], t[1] = tmp;
            for (int i = 0; i < solutions && total < 2; ++i) {
                if (t[i] >= 0 && t[i] <= 1) {
                    x[total] = p[0].x+2*t[i]*ab.x+t[i]*t[i]*br.x;
                    if (nextDY*(ab.y+t[i]*br.y) >= 0) {
                        dy[total++] = nextDY;
                        nextDY = -nextDY;
                    }
                }
            }
        }
        if (p[2].y == y) {
            if (nextDY > 0 && total > 0) {
                --total;
                nextDY = -1;
            }
            if ((p[2].y < p[1].y || (p[2].y == p[1].y && p[2].y < p[0].y)) && total < 2) {
                x[total] = p[2].x;
                if (nextDY < 0) {
                    dy[total++] = -1;
                    nextDY = 1;
                }
            }
        }
        if (nextDY != (y >= p[2].y ? 1 : -1)) {
            if (total > 0)
                --total;
            else {
                if (fabs(p[2].y-y) < fabs(p[0].y-y))
                    x[total] = p[2].x;
                dy[total++] = nextDY;
            }
        }
        return total;
    }
    /* cubic, edge-segments.cpp:343-403 */
    if (p[0].y == y) {
        if (p[0].y < p[1].y || (p[0].y == p[1].y && (p[0].y < p[2].y || (p[0].y == p[2].y && p[0].y < p[3].y))))
            dy[total++] = 1;
        else
            nextDY = 1;
    }
    {
        v2 ab = vsub(p[1], p[0]);
        v2 br = vsub(vsub(p[2], p[1]), ab);
        v2 as = vsub(vsub(vsub(p[3], p[2]), vsub(p[2], p[1])), br);
        double t[3];
        int solutions = oracle_solve_cubic(t, as.y, 3*br.y, 3*ab.y, p[0].y-y);
        double tmp;
        if (solutions >= 2) {
            if (t[0] > t[1])
                tmp = t[0], t[0] = t[1], t[1] = tmp;
            if (solutions >= 3 && t[1] > t[2]) {
                tmp = t[1], t[1] = t[2], t[2] = tmp;
                if (t[0] > t[1])
                    tmp = t[0], t[0] = t[1], t[1] = tmp;
            }
        }
        for (int i = 0; i < solutions && total < 3; ++i) {
            if (t[i] >= 0 && t[i] <= 1) {
                x[total] = p[0].x+3*t[i]*ab.x+3*t[i]*t[i]*br.x+t[i]*t[i]*t[i]*as.x;
                if (nextDY*(ab.y+2*t[i]*br.y+t[i]*t[i]*as.y) >= 0) {
                    dy[total++] = nextDY;
                    nextDY = -nextDY;
                }
            }
        }
    }
    if (p[3].y == y) {
        if (nextDY > 0 && total > 0) {
            --total;
            nextDY = -1;
        }
        if ((p[3].y < p[2].y || (p[3].y == p[2].y && (p[3].y < p[1].y || (p[3].y == p[1].y && p[3].y < p[0].y)))) && total < 3) {
            x[total] = p[3].x;
            if (nextDY < 0) {
                dy[total++] = -1;
                nextDY = 1;
            }
        }
    }
    if (nextDY != (y >= p[3].y ? 1 : -1)) {
        if (total > 0)
            --total;
        else {
            if (fabs(p[3].y-y) < fabs(p[0].y-y))
                x[total] = p[3].x;
            dy[total++] = nextDY;
        }
    }
    return total;
}

/* interpretFillRule, Scanline.cpp:12-24: 0 NONZERO, 1 ODD, 2 POSITIVE, 3 NEGATIVE */
static int fill_rule(int intersections, int rule) {
    switch (rule) {
        case 0: return intersections != 0;
        case 1: return intersections&1;
        case 2: return intersections > 0;
        case 3: return intersections < 0;
    }
    return 0;
}

/* Scanline::filled(x) over Shape::scanline(y): the sum of the crossing directions of all intersections at or left of x
 * (Scanline.cpp:72-123 sorts and prefix-sums; the sum itself does not depend on the order) */
static int shape_filled(const flat_shape *s, double x, double y, int rule) {
    int sum = 0;
    uint32_t nSeg = s->cbegin[s->nContours];
    for (uint32_t e = 0; e < nSeg; ++e) {
        double ix[3];
        int dy[3];
        int n = seg_scanline(s->orders[e], s->pts+4*(size_t) e, ix, dy, y);
        for (int i = 0; i < n; ++i)
            if (ix[i] <= x)
                sum += dy[i];
    }
    return fill_rule(sum, rule);
}

/* rasterize (rasterization.cpp:9-17) */
int oracle_rasterize(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int shapeYDown,
                     const double *xf, float *out, int w, int h, int outYDown, int fillRule) {
    flat_shape fs = { nContours, cbegin, (const v2 *) points, orders, NULL };
    int flip = outYDown != shapeYDown;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            out[(size_t) (flip ? h-1-y : y)*w+x] = (float) shape_filled(&fs, (x+.5)/xf[0]-xf[2], (y+.5)/xf[1]-xf[3], fillRule);
    return 0;
}

/* distanceSignCorrection for N = 1 (rasterization.cpp:19-32) and N = 3, 4 (multiDistanceSignCorrection :34-96); in place.
 * xf: scale.x, scale.y, translate.x, translate.y (only the projection is used). */
int oracle_sign_correction(int N, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int shapeYDown,
                           const double *xf, float *sdf, int w, int h, int rowStride, int sdfYDown, float zeroValue, int fillRule) {
    (void) colors;
    if (!(N == 1 || N == 3 || N == 4))
        return -1;
    if (!(w && h))
        return 0;
    flat_shape fs = { nContours, cbegin, (const v2 *) points, orders, NULL };
    int flip = sdfYDown != shapeYDown;
    float *base = flip ? sdf+(ptrdiff_t) rowStride*(h-1) : sdf;
    ptrdiff_t stride = flip ? -(ptrdiff_t) rowStride : rowStride;
    float dbl = zeroValue+zeroValue;
    char *match = (char *) calloc((size_t) w*h, 1);
    if (!match)
        return -1;
    int ambiguous = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double px = (x+.5)/xf[0]-xf[2], py = (y+.5)/xf[1]-xf[3];   /* unprojectX / unprojectY */
            int fill = shape_filled(&fs, px, py, fillRule);
            float *msd = base+stride*y+(ptrdiff_t) N*x;
            if (N == 1) {
                if ((msd[0] > zeroValue) != fill)
                    msd[0] = dbl-msd[0];
                continue;
            }
            float sd = fmedian(msd[0], msd[1], msd[2]);
            if (sd == zeroValue)
                ambiguous = 1;
            else if ((sd > zeroValue) != fill) {
                msd[0] = dbl-msd[0];
                msd[1] = dbl-msd[1];
                msd[2] = dbl-msd[2];
                match[(size_t) y*w+x] = -1;
            } else
                match[(size_t) y*w+x] = 1;
            if (N >= 4 && (msd[3] > zeroValue) != fill)
                msd[3] = dbl-msd[3];
        }
    if (N > 1 && ambiguous)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                if (!match[(size_t) y*w+x]) {
                    int nm = 0;
                    if (x > 0) nm += match[(size_t) y*w+x-1];
                    if (x < w-1) nm += match[(size_t) y*w+x+1];
                    if (y > 0) nm += match[(size_t) (y-1)*w+x];
                    if (y < h-1) nm += match[(size_t) (y+1)*w+x];
                    if (nm < 0) {
                        float *msd = base+stride*y+(ptrdiff_t) N*x;
                        msd[0] = dbl-msd[0];
                        msd[1] = dbl-msd[1];
                        msd[2] = dbl-msd[2];
                    }
                }
    free(match);
    return 0;
}

/* ---- renderSDF / simulate8bit (core/render-sdf.cpp, bitmap-interpolation.hpp:11-27) ---- */
static int iclamp(int n, int b) { return n >= 0 && n <= b ? n : (n > 0)*b; }                 /* clamp(n, b), arithmetics.hpp:40-43 */
static double dclamp(double n, double b) { return n >= 0 && n <= b ? n : (double) (n > 0)*b; }
static float fmixd(float a, float b, double w) { return (float) ((1.0-w)*a+w*b); }           /* mix<float, double> */

static void interpolate_n(float *out, const float *bitmap, int w, int h, int N, double px, double py) {
    px = dclamp(px, (double) w);
    py = dclamp(py, (double) h);
    px -= .5;
    py -= .5;
    int l = (int) floor(px), b = (int) floor(py);
    int r = l+1, t = b+1;
    double lr = px-l, bt = py-b;
    l = iclamp(l, w-1), r = iclamp(r, w-1);
    b = iclamp(b, h-1), t = iclamp(t, h-1);
    for (int i = 0; i < N; ++i)
        out[i] = fmixd(fmixd(bitmap[((size_t) b*w+l)*N+i], bitmap[((size_t) b*w+r)*N+i], lr), fmixd(bitmap[((size_t) t*w+l)*N+i], bitmap[((size_t) t*w+r)*N+i], lr), bt);
}

/* renderSDF for (outN <- sdfN) in {1<-1, 3<-1, 1<-3, 3<-3, 1<-4, 4<-4} (render-sdf.cpp:14-173) */
int oracle_render_sdf(int outN, float *out, int ow, int oh, int sdfN, const float *sdf, int sw, int sh, double lower, double upper, float threshold) {
    if (!((outN == 1 && (sdfN == 1 || sdfN == 3 || sdfN == 4)) || (outN == 3 && (sdfN == 1 || sdfN == 3)) || (outN == 4 && sdfN == 4)))
        return -1;
    double sx = (double) sw/ow, sy = (double) sh/oh;
    int hard = lower == upper;
    double mScale = 1, mTranslate = 0;
    float bias = .5f-threshold;
    if (!hard) {
        double f = (double) (ow+oh)/(sw+sh);
        lower *= f;
        upper *= f;
        double width = upper-lower;                     /* DistanceMapping::inverse(range), DistanceMapping.cpp:6-9 */
        mScale = width;
        mTranslate = lower/(width ? width : 1);
    }
    for (int y = 0; y < oh; ++y)
        for (int x = 0; x < ow; ++x) {
            float sd[4];
            interpolate_n(sd, sdf, sw, sh, sdfN, sx*(x+.5), sy*(y+.5));
            float v[4];
            int nv = sdfN;
            if (outN == 1 && sdfN >= 3) {
                v[0] = fmedian(sd[0], sd[1], sd[2]);
                nv = 1;
            } else
                for (int i = 0; i < sdfN; ++i) v[i] = sd[i];
            for (int i = 0; i < nv; ++i) {
                if (hard)
                    v[i] = (float) (v[i] >= threshold);
                else {
                    double m = mScale*((double) (v[i]+bias)+mTranslate)+.5;       /* distVal */
                    v[i] = (float) (m >= 0 && m <= 1 ? m : (double) (m > 0));
                }
            }
            float *o = out+((size_t) y*ow+x)*outN;
            if (outN == 3 && sdfN == 1)
                o[0] = o[1] = o[2] = v[0];
            else
                for (int i = 0; i < outN; ++i) o[i] = v[i];
        }
    return 0;
}

/* simulate8bit: pixelByteToFloat(pixelFloatToByte(x)) (render-sdf.cpp:175-191, pixel-conversion.hpp:8-14) */
void oracle_simulate_8bit(float *pixels, int w, int h, int N) {
    size_t n = (size_t) w*h*N;
    for (size_t i = 0; i < n; ++i) {
        float x = pixels[i];
        float c = x >= 0.f && x <= 1.f ? x : (float) (x > 0.f);
        unsigned char b = (unsigned char) ~(int) (255.5f-255.f*c);
        pixels[i] = 1.f/255.f*(float) b;
    }
}

/* pixelFloatToByte, core/pixel-conversion.hpp:8-10 (clamp: arithmetics.hpp:34-37) */
void oracle_pixel_float_to_byte(const float *in, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float x = in[i];
        float c = x >= 0.f && x <= 1.f ? x : (float) (x > 0.f);
        out[i] = (uint8_t) ~(int) (255.5f-255.f*c);
    }
}
