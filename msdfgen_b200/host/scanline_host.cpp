// scanline_host.cpp — host side of the scanline machinery of msdfgen's public interface: equation solvers, Scanline,
// EdgeSegment::scanlineIntersections, Shape::scanline / orientContours, scanlineSDF and estimateSDFError.
//
// These are caller-side helpers (geometry preprocessing and a quality diagnostic), O(edges) or O(rows x width) host work
// that never sits on the per-texel path; the per-texel consumers of scanlines (distanceSignCorrection, rasterize) run on
// the device (csrc/scanline.cuh).  Behaviour follows /root/reference/core/{equation-solver,Scanline,edge-segments,Shape,
// sdf-error-estimation}.cpp; tests/test_host_scanline.py compares every function with the compiled reference.
#include "msdfgen.h"

#include <algorithm>
#include <cmath>

namespace msdfgen {

// ---- equation solvers (equation-solver.cpp) ---------------------------------------------------------------------------------

int solveQuadratic(double x[2], double a, double b, double c) {
    // a == 0, or b so much larger than a that the quadratic term is numerically irrelevant: linear equation
    if (a == 0 || fabs(b) > 1e12*fabs(a)) {
        if (b == 0)
            return c == 0 ? -1 : 0;     // 0 == 0: every x is a solution
        x[0] = -c/b;
        return 1;
    }
    double discriminant = b*b-4*a*c;
    if (discriminant > 0) {
        double root = sqrt(discriminant);
        x[0] = (-b+root)/(2*a);
        x[1] = (-b-root)/(2*a);
        return 2;
    }
    if (discriminant == 0) {
        x[0] = -b/(2*a);
        return 1;
    }
    return 0;
}

// x^3 + a x^2 + b x + c = 0: trigonometric form for three real roots, Cardano otherwise
static int solveNormalizedCubic(double x[3], double a, double b, double c) {
    double aa = a*a;
    double q = 1/9.*(aa-3*b);
    double r = 1/54.*(a*(2*aa-9*b)+27*c);
    double rr = r*r;
    double qqq = q*q*q;
    a *= 1/3.;
    if (rr < qqq) {
        double t = r/sqrt(qqq);
        if (t < -1) t = -1;
        if (t > 1) t = 1;
        t = acos(t);
        q = -2*sqrt(q);
        x[0] = q*cos(1/3.*t)-a;
        x[1] = q*cos(1/3.*(t+2*M_PI))-a;
        x[2] = q*cos(1/3.*(t-2*M_PI))-a;
        return 3;
    }
    double u = (r < 0 ? 1 : -1)*pow(fabs(r)+sqrt(rr-qqq), 1/3.);
    double v = u == 0 ? 0 : q/u;
    x[0] = (u+v)-a;
    if (u == v || fabs(u-v) < 1e-12*fabs(u+v)) {
        x[1] = -.5*(u+v)-a;
        return 2;
    }
    return 1;
}

int solveCubic(double x[3], double a, double b, double c, double d) {
    if (a != 0) {
        double bn = b/a;
        if (fabs(bn) < 1e6)     // beyond this ratio treating a as zero is the more accurate choice
            return solveNormalizedCubic(x, bn, c/a, d/a);
    }
    return solveQuadratic(x, b, c, d);
}

// ---- Scanline (Scanline.cpp) ------------------------------------------------------------------------------------------------

bool interpretFillRule(int intersections, FillRule fillRule) {
    switch (fillRule) {
        case FILL_NONZERO: return intersections != 0;
        case FILL_ODD: return (intersections&1) != 0;
        case FILL_POSITIVE: return intersections > 0;
        case FILL_NEGATIVE: return intersections < 0;
    }
    return false;
}

Scanline::Scanline() : lastIndex(0) { }

// Sorts by x and turns the directions into running totals: after this, intersections[i].direction is the winding sum of
// everything at or left of intersections[i].x (for entries sharing an x only the last one carries the full sum, and that is
// the one every query lands on).
void Scanline::preprocess() {
    lastIndex = 0;
    std::sort(intersections.begin(), intersections.end(), [](const Intersection &a, const Intersection &b) { return a.x < b.x; });
    int total = 0;
    for (size_t i = 0; i < intersections.size(); ++i) {
        total += intersections[i].direction;
        intersections[i].direction = total;
    }
}

void Scanline::setIntersections(const std::vector<Intersection> &list) {
    intersections = list;
    preprocess();
}

void Scanline::setIntersections(std::vector<Intersection> &&list) {
    intersections = (std::vector<Intersection> &&) list;
    preprocess();
}

// Index of the last intersection with x' <= x, or -1; remembers the position because queries usually sweep left to right.
int Scanline::moveTo(double x) const {
    if (intersections.empty())
        return -1;
    int index = lastIndex;
    if (x < intersections[index].x) {
        while (x < intersections[index].x) {
            if (index == 0) {
                lastIndex = 0;
                return -1;
            }
            --index;
        }
    } else {
        const int last = (int) intersections.size()-1;
        while (index < last && x >= intersections[index+1].x)
            ++index;
    }
    lastIndex = index;
    return index;
}

int Scanline::countIntersections(double x) const {
    return moveTo(x)+1;
}

int Scanline::sumIntersections(double x) const {
    int index = moveTo(x);
    return index >= 0 ? intersections[index].direction : 0;
}

bool Scanline::filled(double x, FillRule fillRule) const {
    return interpretFillRule(sumIntersections(x), fillRule);
}

// Length of the part of [xFrom, xTo] on which the two scanlines agree (both filled or both empty).  A two-way merge over the
// intersection lists; a cursor that has run out of intersections parks at xTo.
double Scanline::overlap(const Scanline &a, const Scanline &b, double xFrom, double xTo, FillRule fillRule) {
    struct Cursor {
        const std::vector<Intersection> &list;
        int index;
        double x;
        bool inside;
        Cursor(const std::vector<Intersection> &l, double xEnd) : list(l), index(0), x(l.empty() ? xEnd : l[0].x), inside(false) { }
        // consumes the intersection at the cursor if the sweep has reached it
        void step(double at, double xEnd, FillRule rule) {
            if (x == at && index < (int) list.size()) {
                inside = interpretFillRule(list[index].direction, rule);
                ++index;
                x = index < (int) list.size() ? list[index].x : xEnd;
            }
        }
    };
    Cursor ca(a.intersections, xTo), cb(b.intersections, xTo);
    // state at xFrom
    while (ca.x < xFrom || cb.x < xFrom) {
        double next = min(ca.x, cb.x);
        ca.step(next, xTo, fillRule);
        cb.step(next, xTo, fillRule);
    }
    double total = 0;
    double x = xFrom;
    while (ca.x < xTo || cb.x < xTo) {
        double next = min(ca.x, cb.x);
        if (ca.inside == cb.inside)
            total += next-x;
        ca.step(next, xTo, fillRule);
        cb.step(next, xTo, fillRule);
        x = next;
    }
    if (ca.inside == cb.inside)
        total += xTo-x;
    return total;
}

// ---- EdgeSegment::scanlineIntersections (edge-segments.cpp:279-403) ----------------------------------------------------------

int LinearSegment::scanlineIntersections(double x[3], int dy[3], double y) const {
    // half-open in y so that a vertex shared by two edges is counted once
    if ((y >= p[0].y && y < p[1].y) || (y >= p[1].y && y < p[0].y)) {
        double t = (y-p[0].y)/(p[1].y-p[0].y);
        x[0] = mix(p[0].x, p[1].x, t);
        dy[0] = sign(p[1].y-p[0].y);
        return 1;
    }
    return 0;
}

namespace {

// Shared body of the quadratic and cubic cases.  The curve is cut into monotone pieces by the roots of y(t) = y; `next` is
// the direction the next accepted crossing must have (pieces alternate between rising and falling).  c[0..ORDER] are the
// control points; ab, br, as the difference vectors of the power basis.
template <int ORDER>
int curveScanlineIntersections(const Point2 *c, double x[3], int dy[3], double y) {
    const Point2 &first = c[0], &last = c[ORDER];
    const Vector2 ab = c[1]-c[0];
    const Vector2 br = c[2]-c[1]-ab;
    const Vector2 as = ORDER == 3 ? (c[ORDER]-c[2])-(c[2]-c[1])-br : Vector2();
    int total = 0;
    int next = y > first.y ? 1 : -1;
    x[total] = first.x;
    if (first.y == y) {
        // leaves the start point upwards?  the first control point that is not level with it decides
        bool up = false;
        for (int k = 1; k <= ORDER; ++k)
            if (c[k].y != first.y) {
                up = first.y < c[k].y;
                break;
            }
        if (up)
            dy[total++] = 1;
        else
            next = 1;
    }
    double t[3];
    int solutions = ORDER == 3 ? solveCubic(t, as.y, 3*br.y, 3*ab.y, first.y-y) : solveQuadratic(t, br.y, 2*ab.y, first.y-y);
    // ascending order (at most three values; plain compare-and-swap keeps the behaviour defined for any input)
    for (int i = 1; i < solutions; ++i)
        for (int j = i; j > 0 && t[j-1] > t[j]; --j)
            std::swap(t[j-1], t[j]);
    for (int i = 0; i < solutions && total < ORDER; ++i) {
        if (t[i] >= 0 && t[i] <= 1) {
            double slope;
            if (ORDER == 3) {
                x[total] = first.x+3*t[i]*ab.x+3*t[i]*t[i]*br.x+t[i]*t[i]*t[i]*as.x;
                slope = ab.y+2*t[i]*br.y+t[i]*t[i]*as.y;
            } else {
                x[total] = first.x+2*t[i]*ab.x+t[i]*t[i]*br.x;
                slope = ab.y+t[i]*br.y;
            }
            if (next*slope >= 0) {
                dy[total++] = next;
                next = -next;
            }
        }
    }
    if (last.y == y) {
        if (next > 0 && total > 0) {
            --total;
            next = -1;
        }
        // arrives at the end point from above?  the last control point that is not level with it decides
        bool down = false;
        for (int k = ORDER-1; k >= 0; --k)
            if (c[k].y != last.y) {
                down = last.y < c[k].y;
                break;
            }
        if (down && total < ORDER) {
            x[total] = last.x;
            if (next < 0) {
                dy[total++] = -1;
                next = 1;
            }
        }
    }
    // the parity of the crossings has to match the sides the end points lie on
    if (next != (y >= last.y ? 1 : -1)) {
        if (total > 0)
            --total;
        else {
            if (fabs(last.y-y) < fabs(first.y-y))
                x[total] = last.x;
            dy[total++] = next;
        }
    }
    return total;
}

}  // namespace

int QuadraticSegment::scanlineIntersections(double x[3], int dy[3], double y) const {
    return curveScanlineIntersections<2>(p, x, dy, y);
}

int CubicSegment::scanlineIntersections(double x[3], int dy[3], double y) const {
    return curveScanlineIntersections<3>(p, x, dy, y);
}

// ---- Shape::scanline, Shape::orientContours (Shape.cpp:117-135, :144-199) -----------------------------------------------------

void Shape::scanline(Scanline &line, double y) const {
    std::vector<Scanline::Intersection> found;
    double x[3];
    int dy[3];
    for (size_t c = 0; c < contours.size(); ++c)
        for (size_t e = 0; e < contours[c].edges.size(); ++e) {
            int n = contours[c].edges[e]->scanlineIntersections(x, dy, y);
            for (int i = 0; i < n; ++i) {
                Scanline::Intersection hit = { x[i], dy[i] };
                found.push_back(hit);
            }
        }
    line.setIntersections((std::vector<Scanline::Intersection> &&) found);
}

void Shape::orientContours() {
    struct Hit {
        double x;
        int direction;
        int contour;
    };
    // an irrational fraction keeps the probe line away from corners and other special points
    const double ratio = .5*(sqrt(5)-1);
    std::vector<int> votes(contours.size(), 0);
    std::vector<Hit> hits;
    for (int i = 0; i < (int) contours.size(); ++i) {
        if (votes[i] || contours[i].edges.empty())
            continue;
        // a y that crosses contour i: between its first start point and the first edge end (or interior point) at another height
        const std::vector<EdgeHolder> &edges = contours[i].edges;
        double y0 = edges.front()->point(0).y;
        double y1 = y0;
        for (size_t e = 0; e < edges.size() && y0 == y1; ++e)
            y1 = edges[e]->point(1).y;
        for (size_t e = 0; e < edges.size() && y0 == y1; ++e)
            y1 = edges[e]->point(ratio).y;
        double y = mix(y0, y1, ratio);
        // all crossings of the whole shape with that line
        double x[3];
        int dy[3];
        for (int j = 0; j < (int) contours.size(); ++j)
            for (size_t e = 0; e < contours[j].edges.size(); ++e) {
                int n = contours[j].edges[e]->scanlineIntersections(x, dy, y);
                for (int k = 0; k < n; ++k) {
                    Hit hit = { x[k], dy[k], j };
                    hits.push_back(hit);
                }
            }
        if (hits.empty())
            continue;
        std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.x < b.x; });
        // crossings that coincide say nothing reliable
        for (size_t j = 1; j < hits.size(); ++j)
            if (hits[j].x == hits[j-1].x)
                hits[j].direction = hits[j-1].direction = 0;
        // walking left to right, crossings alternately enter and leave the filled area (even-odd); a contour whose direction
        // at an entering crossing is "down" is wound the non-zero way, otherwise it is not
        for (size_t j = 0; j < hits.size(); ++j)
            if (hits[j].direction)
                votes[hits[j].contour] += 2*(int((j&1) != 0)^int(hits[j].direction > 0))-1;
        hits.clear();
    }
    for (size_t i = 0; i < contours.size(); ++i)
        if (votes[i] < 0)
            contours[i].reverse();
}

// ---- scanlineSDF, estimateSDFError (sdf-error-estimation.cpp) -----------------------------------------------------------------

namespace {

// Row interpolation set-up shared by the single- and multi-channel versions: rows b and t of the field and the weight of t.
struct RowPair { int b, t; double weight; };

template <int N>
RowPair rowsAt(const BitmapConstSection<float, N> &sdf, const Projection &projection, double y, YAxisOrientation yAxisOrientation) {
    double pixelY = clamp(projection.projectY(y)-.5, double(sdf.height-1));
    if (yAxisOrientation == Y_DOWNWARD)
        pixelY = sdf.height-1-pixelY;
    RowPair rows;
    rows.b = (int) floor(pixelY);
    rows.t = rows.b+1;
    rows.weight = pixelY-rows.b;
    if (rows.t >= sdf.height) {
        rows.b = rows.t = sdf.height-1;
        rows.weight = 1;
    }
    return rows;
}

void scanlineSingle(Scanline &line, const BitmapConstSection<float, 1> &sdf, const Projection &projection, double y, YAxisOrientation yAxisOrientation) {
    std::vector<Scanline::Intersection> found;
    if (sdf.width > 0 && sdf.height > 0) {
        const RowPair rows = rowsAt(sdf, projection, y, yAxisOrientation);
        float right = mix(*sdf(0, rows.b), *sdf(0, rows.t), rows.weight);
        if (right > .5f) {
            Scanline::Intersection open = { -1e240, 1 };    // the row starts inside
            found.push_back(open);
        }
        for (int r = 1; r < sdf.width; ++r) {
            const float left = right;
            right = mix(*sdf(r, rows.b), *sdf(r, rows.t), rows.weight);
            if (left != right) {
                double frac = double(.5f-left)/double(right-left);
                if (frac >= 0 && frac <= 1) {
                    Scanline::Intersection hit = { projection.unprojectX(r-1+frac+.5), sign(right-left) };
                    found.push_back(hit);
                }
            }
        }
    }
    line.setIntersections((std::vector<Scanline::Intersection> &&) found);
}

template <int N>
void scanlineMulti(Scanline &line, const BitmapConstSection<float, N> &sdf, const Projection &projection, double y, YAxisOrientation yAxisOrientation) {
    std::vector<Scanline::Intersection> found;
    if (sdf.width > 0 && sdf.height > 0) {
        const RowPair rows = rowsAt(sdf, projection, y, yAxisOrientation);
        float left[3], right[3];
        for (int k = 0; k < 3; ++k)
            right[k] = mix(sdf(0, rows.b)[k], sdf(0, rows.t)[k], rows.weight);
        bool inside = median(right[0], right[1], right[2]) > .5f;
        if (inside) {
            Scanline::Intersection open = { -1e240, 1 };
            found.push_back(open);
        }
        for (int r = 1; r < sdf.width; ++r) {
            for (int k = 0; k < 3; ++k) {
                left[k] = right[k];
                right[k] = mix(sdf(r, rows.b)[k], sdf(r, rows.t)[k], rows.weight);
            }
            // a channel crossing .5 between the two columns is an edge of the median only where that channel IS the median
            Scanline::Intersection hits[3];
            int nHits = 0;
            for (int k = 0; k < 3; ++k) {
                if (left[k] == right[k])
                    continue;
                double frac = double(.5f-left[k])/double(right[k]-left[k]);
                if (!(frac >= 0 && frac <= 1))
                    continue;
                float at[3] = { mix(left[0], right[0], frac), mix(left[1], right[1], frac), mix(left[2], right[2], frac) };
                if (median(at[0], at[1], at[2]) == at[k]) {
                    hits[nHits].x = projection.unprojectX(r-1+frac+.5);
                    hits[nHits].direction = sign(right[k]-left[k]);
                    ++nHits;
                }
            }
            std::stable_sort(hits, hits+nHits, [](const Scanline::Intersection &a, const Scanline::Intersection &b) { return a.x < b.x; });
            // keep only crossings that toggle the state
            for (int i = 0; i < nHits; ++i)
                if ((hits[i].direction > 0) == !inside) {
                    found.push_back(hits[i]);
                    inside = !inside;
                }
            // the state has to agree with the median at the right column
            float scalar = median(right[0], right[1], right[2]);
            if ((scalar > .5f) != inside && scalar != .5f && !found.empty()) {
                found.pop_back();
                inside = !inside;
            }
        }
    }
    line.setIntersections((std::vector<Scanline::Intersection> &&) found);
}

template <int N>
double estimateError(const BitmapConstSection<float, N> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) {
    if (sdf.width <= 1 || sdf.height <= 1 || scanlinesPerRow < 1)
        return 0;
    const double step = 1./scanlinesPerRow;
    const double xFrom = projection.unprojectX(.5);
    const double xTo = projection.unprojectX(sdf.width-.5);
    const double perUnit = 1/(xTo-xFrom);
    double error = 0;
    Scanline exact, sampled;
    for (int row = 0; row < sdf.height-1; ++row)
        for (int sub = 0; sub < scanlinesPerRow; ++sub) {
            double y = projection.unprojectY(row+(sub+.5)*step+.5);
            shape.scanline(exact, y);
            scanlineSDF(sampled, sdf, projection, y, shape.getYAxisOrientation());
            error += 1-perUnit*Scanline::overlap(exact, sampled, xFrom, xTo, fillRule);
        }
    return error/((sdf.height-1)*scanlinesPerRow);
}

inline YAxisOrientation orientationOf(bool inverseYAxis) {
    return inverseYAxis ? MSDFGEN_Y_AXIS_NONDEFAULT_ORIENTATION : MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION;
}

}  // namespace

void scanlineSDF(Scanline &line, const BitmapConstSection<float, 1> &sdf, const Projection &projection, double y, YAxisOrientation yAxisOrientation) { scanlineSingle(line, sdf, projection, y, yAxisOrientation); }
void scanlineSDF(Scanline &line, const BitmapConstSection<float, 3> &sdf, const Projection &projection, double y, YAxisOrientation yAxisOrientation) { scanlineMulti(line, sdf, projection, y, yAxisOrientation); }
void scanlineSDF(Scanline &line, const BitmapConstSection<float, 4> &sdf, const Projection &projection, double y, YAxisOrientation yAxisOrientation) { scanlineMulti(line, sdf, projection, y, yAxisOrientation); }

double estimateSDFError(const BitmapConstSection<float, 1> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) { return estimateError(sdf, shape, projection, scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 3> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) { return estimateError(sdf, shape, projection, scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 4> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) { return estimateError(sdf, shape, projection, scanlinesPerRow, fillRule); }

void scanlineSDF(Scanline &line, const BitmapConstSection<float, 1> &sdf, const Projection &projection, double y, bool inverseYAxis) { scanlineSDF(line, sdf, projection, y, orientationOf(inverseYAxis)); }
void scanlineSDF(Scanline &line, const BitmapConstSection<float, 3> &sdf, const Projection &projection, double y, bool inverseYAxis) { scanlineSDF(line, sdf, projection, y, orientationOf(inverseYAxis)); }
void scanlineSDF(Scanline &line, const BitmapConstSection<float, 4> &sdf, const Projection &projection, double y, bool inverseYAxis) { scanlineSDF(line, sdf, projection, y, orientationOf(inverseYAxis)); }
void scanlineSDF(Scanline &line, const BitmapConstSection<float, 1> &sdf, const Vector2 &scale, const Vector2 &translate, bool inverseYAxis, double y) { scanlineSDF(line, sdf, Projection(scale, translate), y, inverseYAxis); }
void scanlineSDF(Scanline &line, const BitmapConstSection<float, 3> &sdf, const Vector2 &scale, const Vector2 &translate, bool inverseYAxis, double y) { scanlineSDF(line, sdf, Projection(scale, translate), y, inverseYAxis); }
void scanlineSDF(Scanline &line, const BitmapConstSection<float, 4> &sdf, const Vector2 &scale, const Vector2 &translate, bool inverseYAxis, double y) { scanlineSDF(line, sdf, Projection(scale, translate), y, inverseYAxis); }
double estimateSDFError(const BitmapConstSection<float, 1> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule) { return estimateSDFError(sdf, shape, Projection(scale, translate), scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 3> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule) { return estimateSDFError(sdf, shape, Projection(scale, translate), scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 4> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule) { return estimateSDFError(sdf, shape, Projection(scale, translate), scanlinesPerRow, fillRule); }

}  // namespace msdfgen
