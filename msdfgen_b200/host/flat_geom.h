// flat_geom.h — the host library's curve geometry, written ONCE against the wire format of the device path.
//
// The C ABI (include/msdfgen_b200.h) ships a shape as flat arrays: per segment an order (1 line, 2 quadratic, 3 cubic), up to
// four control points and a colour.  Everything the host has to compute per edge before the GPU takes over — points and
// tangents of a Bezier segment, bounds, the split in thirds, windings, cusp de-convergence — is defined here on exactly that
// representation (a `Seg`: control points + order + colour), order-generic wherever the arithmetic allows it: de Casteljau
// levels for the curve, its hodograph and its second difference instead of one hand-expanded formula per segment class.
// The msdfgen-compatible object model of msdfgen.h (EdgeSegment and its three subclasses, Contour, Shape) is a thin adaptor over
// these functions (shape_model.cpp), and the batch ingest never builds those objects at all (flat_prep.cpp).
//
// Every result must equal the reference's to the last bit (it is the INPUT of a device path compared bit for bit), so the order
// of the floating-point operations follows the reference's expressions; tests/test_host_model.py checks that against the compiled
// reference.  Behaviour restated from /root/reference/core/: edge-segments.cpp:54-139 (point, direction, directionChange, length),
// :405-458 (bound), :460-531 (reverse, move*Point, splitInThirds, convertToCubic), Contour.cpp:19-86 (bound, boundMiters, winding),
// equation-solver.cpp:9-32 (the quadratic behind the cubic's bounds), convergent-curve-ordering.cpp (ordering at a cusp).
#pragma once

#include "msdfgen.h"

#include <cstdint>

namespace msdfgen {
namespace flat {

/// One edge segment in wire form.
struct Seg {
    Point2 p[4];        // p[0..order]
    int order;          // 1, 2 or 3
    EdgeColor color;
};

inline Seg makeSeg(int order, Point2 p0, Point2 p1, Point2 p2 = Point2(), Point2 p3 = Point2(), EdgeColor color = WHITE) {
    Seg s;
    s.p[0] = p0; s.p[1] = p1; s.p[2] = p2; s.p[3] = p3;
    s.order = order;
    s.color = color;
    return s;
}

/// Segment s of a flat batch: 8 doubles per segment (x0 y0 .. x3 y3), one order byte, one colour byte (NULL = white).
inline Seg loadSeg(const double *points, const uint8_t *orders, const uint8_t *colors, size_t s) {
    const double *q = points+8*s;
    return makeSeg(orders[s], Point2(q[0], q[1]), Point2(q[2], q[3]), Point2(q[4], q[5]), Point2(q[6], q[7]), EdgeColor(colors ? colors[s] : (uint8_t) WHITE));
}

/// The inverse of loadSeg; control-point slots the order does not use are written as zeros.
inline void storeSeg(const Seg &seg, double *points, uint8_t *orders, uint8_t *colors, size_t s) {
    for (int k = 0; k < 4; ++k) {
        points[8*s+2*k] = k <= seg.order ? seg.p[k].x : 0;
        points[8*s+2*k+1] = k <= seg.order ? seg.p[k].y : 0;
    }
    orders[s] = (uint8_t) seg.order;
    colors[s] = (uint8_t) seg.color;
}

// ---- evaluation ----------------------------------------------------------------------------------------------------------------

/// de Casteljau: collapses n+1 values to the curve value at t, level by level.
inline Vector2 collapse(Vector2 *q, int n, double t) {
    for (int level = n; level > 0; --level)
        for (int i = 0; i < level; ++i)
            q[i] = mix(q[i], q[i+1], t);
    return q[0];
}

inline Point2 pointAt(const Point2 *p, int order, double t) {
    Vector2 q[4];
    for (int i = 0; i <= order; ++i)
        q[i] = p[i];
    return collapse(q, order, t);
}

/// Tangent (the derivative up to a positive factor).  Where it vanishes at a degenerate end, the chord to the next distinct
/// control point stands in: for quadratics at any parameter, for cubics only at the two ends.
inline Vector2 directionAt(const Point2 *p, int order, double t) {
    Vector2 d[3];
    for (int i = 0; i < order; ++i)
        d[i] = p[i+1]-p[i];
    const Vector2 tangent = collapse(d, order-1, t);
    if (order >= 2 && !tangent) {
        if (order == 2 || t == 0)
            return p[2]-p[0];
        if (t == 1)
            return p[3]-p[1];
    }
    return tangent;
}

/// Second difference of the control polygon along the curve (zero for lines).
inline Vector2 directionChangeAt(const Point2 *p, int order, double t) {
    if (order < 2)
        return Vector2();
    Vector2 e[2];
    for (int i = 0; i+1 < order; ++i)
        e[i] = (p[i+2]-p[i+1])-(p[i+1]-p[i]);
    return collapse(e, order-2, t);
}

/// Length of the polyline through `pieces`+1 equidistant parameters (what the ink-trap colouring calls the edge's length).
inline double polylineLength(const Point2 *p, int order, int pieces) {
    double total = 0;
    Point2 prev = pointAt(p, order, 0);
    for (int i = 1; i <= pieces; ++i) {
        Point2 cur = pointAt(p, order, 1./pieces*i);
        total += (cur-prev).length();
        prev = cur;
    }
    return total;
}

/// Exact arc length of a line (order 1) or a parabola arc (order 2, closed form).
inline double arcLength(const Point2 *p, int order) {
    if (order == 1)
        return (p[1]-p[0]).length();
    const Vector2 ab = p[1]-p[0];
    const Vector2 br = p[2]-p[1]-ab;
    const double abab = dotProduct(ab, ab), abbr = dotProduct(ab, br), brbr = dotProduct(br, br);
    const double abLen = std::sqrt(abab), brLen = std::sqrt(brbr);
    const double crs = crossProduct(ab, br);
    const double h = std::sqrt(abab+abbr+abbr+brbr);
    return (brLen*((abbr+brbr)*h-abbr*abLen)+crs*crs*std::log((brLen*h+abbr+brbr)/(brLen*abLen+abbr)))/(brbr*brLen);
}

// ---- bounds --------------------------------------------------------------------------------------------------------------------

inline void growBounds(Point2 q, double &xMin, double &yMin, double &xMax, double &yMax) {
    if (q.x < xMin) xMin = q.x;
    if (q.y < yMin) yMin = q.y;
    if (q.x > xMax) xMax = q.x;
    if (q.y > yMax) yMax = q.y;
}

/// Parameters where one coordinate of the curve is stationary, c[0..order] being that coordinate of the control points:
/// the root of a linear hodograph (quadratic segment) or the roots of a quadratic one (cubic segment).  Returns the count.
inline int stationaryParams(double ts[2], const double *c, int order) {
    if (order == 2) {
        const double bot = (c[1]-c[0])-(c[2]-c[1]);
        if (!bot)
            return 0;
        ts[0] = (c[1]-c[0])/bot;
        return 1;
    }
    if (order != 3)
        return 0;
    // hodograph / 3 = a0 + a1 t + a2 t^2
    const double a0 = c[1]-c[0];
    const double a1 = 2*(c[2]-c[1]-a0);
    const double a2 = c[3]-3*c[2]+3*c[1]-c[0];
    if (a2 == 0 || std::fabs(a1) > 1e12*std::fabs(a2)) {
        // (numerically) linear
        if (a1 == 0)
            return 0;
        ts[0] = -a0/a1;
        return 1;
    }
    const double discriminant = a1*a1-4*a2*a0;
    if (discriminant > 0) {
        const double root = std::sqrt(discriminant);
        ts[0] = (-a1+root)/(2*a2);
        ts[1] = (-a1-root)/(2*a2);
        return 2;
    }
    if (discriminant == 0) {
        ts[0] = -a1/(2*a2);
        return 1;
    }
    return 0;
}

/// Grows a box by the exact extent of the segment: both ends and the interior stationary points of x and of y.
inline void growBounds(const Point2 *p, int order, double &xMin, double &yMin, double &xMax, double &yMax) {
    growBounds(p[0], xMin, yMin, xMax, yMax);
    growBounds(p[order], xMin, yMin, xMax, yMax);
    for (int axis = 0; axis < 2 && order > 1; ++axis) {
        double c[4], ts[2];
        for (int i = 0; i <= order; ++i)
            c[i] = axis ? p[i].y : p[i].x;
        const int n = stationaryParams(ts, c, order);
        for (int i = 0; i < n; ++i)
            if (ts[i] > 0 && ts[i] < 1)
                growBounds(pointAt(p, order, ts[i]), xMin, yMin, xMax, yMax);
    }
}

// ---- editing -------------------------------------------------------------------------------------------------------------------

inline void reverse(Point2 *p, int order) {
    for (int i = 0; 2*i < order; ++i) {
        Point2 t = p[i];
        p[i] = p[order-i];
        p[order-i] = t;
    }
}

/// Moves the start (end == 0) or end (end == 1) point.  A quadratic slides its control point along the far leg so that the
/// tangent direction at the moved end is kept (and gives up if that would flip it); a cubic carries the adjacent handle along.
inline void moveEnd(Point2 *p, int order, int end, Point2 to) {
    const int e = end ? order : 0;
    if (order == 2) {
        const int far = end ? 0 : 2;
        const Vector2 origDir = p[e]-p[1];
        const Point2 origP1 = p[1];
        p[1] += crossProduct(p[e]-p[1], to-p[e])/crossProduct(p[e]-p[1], p[far]-p[1])*(p[far]-p[1]);
        p[e] = to;
        if (dotProduct(origDir, p[e]-p[1]) < 0)
            p[1] = origP1;
        return;
    }
    if (order == 3)
        p[end ? 2 : 1] += to-p[e];
    p[e] = to;
}

/// The same curve as a cubic (degree elevation of a quadratic).
inline Seg toCubic(const Seg &s) {
    if (s.order != 2)
        return s;
    return makeSeg(3, s.p[0], mix(s.p[0], s.p[1], 2/3.), mix(s.p[1], s.p[2], 1/3.), s.p[2], s.color);
}

/// The three sub-curves over [0,1/3], [1/3,2/3], [2/3,1], same order and colour.
inline void splitThirds(const Seg &s, Seg out[3]) {
    const Point2 *p = s.p;
    const Point2 a = pointAt(p, s.order, 1/3.), b = pointAt(p, s.order, 2/3.);
    switch (s.order) {
        case 1:
            out[0] = makeSeg(1, p[0], a);
            out[1] = makeSeg(1, a, b);
            out[2] = makeSeg(1, b, p[1]);
            break;
        case 2:
            out[0] = makeSeg(2, p[0], mix(p[0], p[1], 1/3.), a);
            out[1] = makeSeg(2, a, mix(mix(p[0], p[1], 5/9.), mix(p[1], p[2], 4/9.), .5), b);
            out[2] = makeSeg(2, b, mix(p[1], p[2], 2/3.), p[2]);
            break;
        default: {
            // first de Casteljau level at 1/3 (q) and at 2/3 (r); coincident end handles stay exactly coincident
            const Point2 q01 = mix(p[0], p[1], 1/3.), q12 = mix(p[1], p[2], 1/3.), q23 = mix(p[2], p[3], 1/3.);
            const Point2 r01 = mix(p[0], p[1], 2/3.), r12 = mix(p[1], p[2], 2/3.), r23 = mix(p[2], p[3], 2/3.);
            out[0] = makeSeg(3, p[0], p[0] == p[1] ? p[0] : q01, mix(q01, q12, 1/3.), a);
            out[1] = makeSeg(3, a, mix(mix(q01, q12, 1/3.), mix(q12, q23, 1/3.), 2/3.), mix(mix(r01, r12, 2/3.), mix(r12, r23, 2/3.), 1/3.), b);
            out[2] = makeSeg(3, b, mix(r12, r23, 2/3.), p[2] == p[3] ? p[3] : r23, p[3]);
        }
    }
    out[0].color = out[1].color = out[2].color = s.color;
}

// ---- closed contours: e[0..n) ----------------------------------------------------------------------------------------------------

/// Orientation of the contour: sign of the shoelace sum over sample points (edge starts; contours of one or two edges are
/// sampled inside their edges as well, since their starts alone span no area).
inline int winding(const Seg *e, int n) {
    if (n <= 0)
        return 0;
    Point2 samples[4];
    int count = 0;
    if (n == 1) {
        samples[0] = pointAt(e[0].p, e[0].order, 0);
        samples[1] = pointAt(e[0].p, e[0].order, 1/3.);
        samples[2] = pointAt(e[0].p, e[0].order, 2/3.);
        count = 3;
    } else if (n == 2) {
        samples[0] = pointAt(e[0].p, e[0].order, 0);
        samples[1] = pointAt(e[0].p, e[0].order, .5);
        samples[2] = pointAt(e[1].p, e[1].order, 0);
        samples[3] = pointAt(e[1].p, e[1].order, .5);
        count = 4;
    }
    double total = 0;
    if (count) {
        for (int i = 0; i < count; ++i) {
            const Point2 &a = samples[i], &b = samples[(i+1)%count];
            total += (b.x-a.x)*(a.y+b.y);
        }
    } else {
        Point2 prev = pointAt(e[n-1].p, e[n-1].order, 0);
        for (int i = 0; i < n; ++i) {
            const Point2 cur = pointAt(e[i].p, e[i].order, 0);
            total += (cur.x-prev.x)*(prev.y+cur.y);
            prev = cur;
        }
    }
    return sign(total);
}

/// Grows a box by the tips of the miter joins of an outline of half-width `border` (polarity selects convex or concave corners).
inline void growMiterBounds(const Seg *e, int n, double &xMin, double &yMin, double &xMax, double &yMax, double border, double miterLimit, int polarity) {
    if (n <= 0)
        return;
    Vector2 prevDir = directionAt(e[n-1].p, e[n-1].order, 1).normalize(true);
    for (int i = 0; i < n; ++i) {
        const Vector2 dir = -directionAt(e[i].p, e[i].order, 0).normalize(true);
        if (polarity*crossProduct(prevDir, dir) >= 0) {
            double miterLength = miterLimit;
            const double q = .5*(1-dotProduct(prevDir, dir));
            if (q > 0)
                miterLength = min(1/std::sqrt(q), miterLimit);
            growBounds(pointAt(e[i].p, e[i].order, 0)+border*miterLength*(prevDir+dir).normalize(true), xMin, yMin, xMax, yMax);
        }
        prevDir = directionAt(e[i].p, e[i].order, 1).normalize(true);
    }
}

}  // namespace flat
}  // namespace msdfgen
