// shape_model.cpp — host-side shape data model and O(edges) shape preparation for the B200 generator.
//
// Everything here runs once per shape on the CPU before the per-pixel work is shipped to the GPU: Bezier segment
// evaluation, bounds, contour winding, Shape::normalize (cusp de-convergence) and the edge colouring strategies.
// Results must equal the reference's exactly (the same doubles, the same colours), because they are the INPUT of the
// device path whose output is compared bit-for-bit; tests/test_host_model.py checks that against the compiled reference.
// Reference behaviour restated from (paths relative to /root/reference/core/):
//   edge-segments.cpp:10-26,54-139,405-531   EdgeHolder.cpp   Contour.cpp   Shape.cpp:29-115,190-206
//   convergent-curve-ordering.cpp   edge-coloring.cpp:19-258   equation-solver.cpp:9-32 (for cubic bounds)
#include "msdfgen.h"
#include "flat_prep.h"

#include <cstdlib>

namespace msdfgen {

// ---------------------------------------------------------------------------------------------------------------------
// small local helpers

namespace {

// Roots of a*x^2 + b*x + c (the closed form the reference's bounds code relies on; -1 = identity, 0..2 roots).
int quadraticRoots(double x[2], double a, double b, double c) {
    if (a == 0 || std::fabs(b) > 1e12*std::fabs(a)) {
        if (b == 0)
            return c == 0 ? -1 : 0;
        x[0] = -c/b;
        return 1;
    }
    double discriminant = b*b-4*a*c;
    if (discriminant > 0) {
        double root = std::sqrt(discriminant);
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

inline void growBounds(Point2 p, double &xMin, double &yMin, double &xMax, double &yMax) {
    if (p.x < xMin) xMin = p.x;
    if (p.y < yMin) yMin = p.y;
    if (p.x > xMax) xMax = p.x;
    if (p.y > yMax) yMax = p.y;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// EdgeSegment factory: degenerate curves collapse to the lowest order that represents them

EdgeSegment *EdgeSegment::create(Point2 p0, Point2 p1, EdgeColor edgeColor) {
    return new LinearSegment(p0, p1, edgeColor);
}

EdgeSegment *EdgeSegment::create(Point2 p0, Point2 p1, Point2 p2, EdgeColor edgeColor) {
    if (!crossProduct(p1-p0, p2-p1))
        return new LinearSegment(p0, p2, edgeColor);
    return new QuadraticSegment(p0, p1, p2, edgeColor);
}

EdgeSegment *EdgeSegment::create(Point2 p0, Point2 p1, Point2 p2, Point2 p3, EdgeColor edgeColor) {
    Vector2 p12 = p2-p1;
    if (!crossProduct(p1-p0, p12) && !crossProduct(p12, p3-p2))
        return new LinearSegment(p0, p3, edgeColor);
    if ((p12 = 1.5*p1-.5*p0) == 1.5*p2-.5*p3)
        return new QuadraticSegment(p0, p12, p3, edgeColor);
    return new CubicSegment(p0, p1, p2, p3, edgeColor);
}

// ---------------------------------------------------------------------------------------------------------------------
// LinearSegment

LinearSegment::LinearSegment(Point2 p0, Point2 p1, EdgeColor edgeColor) : EdgeSegment(edgeColor) { p[0] = p0; p[1] = p1; }
LinearSegment *LinearSegment::clone() const { return new LinearSegment(p[0], p[1], color); }
int LinearSegment::type() const { return (int) EDGE_TYPE; }
const Point2 *LinearSegment::controlPoints() const { return p; }
Point2 LinearSegment::point(double param) const { return mix(p[0], p[1], param); }
Vector2 LinearSegment::direction(double) const { return p[1]-p[0]; }
Vector2 LinearSegment::directionChange(double) const { return Vector2(); }
double LinearSegment::length() const { return (p[1]-p[0]).length(); }

void LinearSegment::bound(double &xMin, double &yMin, double &xMax, double &yMax) const {
    growBounds(p[0], xMin, yMin, xMax, yMax);
    growBounds(p[1], xMin, yMin, xMax, yMax);
}

void LinearSegment::reverse() { Point2 t = p[0]; p[0] = p[1]; p[1] = t; }
void LinearSegment::moveStartPoint(Point2 to) { p[0] = to; }
void LinearSegment::moveEndPoint(Point2 to) { p[1] = to; }

void LinearSegment::splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const {
    Point2 a = point(1/3.), b = point(2/3.);
    part0 = new LinearSegment(p[0], a, color);
    part1 = new LinearSegment(a, b, color);
    part2 = new LinearSegment(b, p[1], color);
}

// ---------------------------------------------------------------------------------------------------------------------
// QuadraticSegment

QuadraticSegment::QuadraticSegment(Point2 p0, Point2 p1, Point2 p2, EdgeColor edgeColor) : EdgeSegment(edgeColor) { p[0] = p0; p[1] = p1; p[2] = p2; }
QuadraticSegment *QuadraticSegment::clone() const { return new QuadraticSegment(p[0], p[1], p[2], color); }
int QuadraticSegment::type() const { return (int) EDGE_TYPE; }
const Point2 *QuadraticSegment::controlPoints() const { return p; }
Point2 QuadraticSegment::point(double param) const { return mix(mix(p[0], p[1], param), mix(p[1], p[2], param), param); }

Vector2 QuadraticSegment::direction(double param) const {
    Vector2 tangent = mix(p[1]-p[0], p[2]-p[1], param);
    return tangent ? tangent : p[2]-p[0];
}

Vector2 QuadraticSegment::directionChange(double) const { return (p[2]-p[1])-(p[1]-p[0]); }

double QuadraticSegment::length() const {
    Vector2 ab = p[1]-p[0];
    Vector2 br = p[2]-p[1]-ab;
    double abab = dotProduct(ab, ab), abbr = dotProduct(ab, br), brbr = dotProduct(br, br);
    double abLen = std::sqrt(abab), brLen = std::sqrt(brbr);
    double crs = crossProduct(ab, br);
    double h = std::sqrt(abab+abbr+abbr+brbr);
    return (brLen*((abbr+brbr)*h-abbr*abLen)+crs*crs*std::log((brLen*h+abbr+brbr)/(brLen*abLen+abbr)))/(brbr*brLen);
}

void QuadraticSegment::bound(double &xMin, double &yMin, double &xMax, double &yMax) const {
    growBounds(p[0], xMin, yMin, xMax, yMax);
    growBounds(p[2], xMin, yMin, xMax, yMax);
    Vector2 bot = (p[1]-p[0])-(p[2]-p[1]);
    if (bot.x) {
        double param = (p[1].x-p[0].x)/bot.x;
        if (param > 0 && param < 1)
            growBounds(point(param), xMin, yMin, xMax, yMax);
    }
    if (bot.y) {
        double param = (p[1].y-p[0].y)/bot.y;
        if (param > 0 && param < 1)
            growBounds(point(param), xMin, yMin, xMax, yMax);
    }
}

void QuadraticSegment::reverse() { Point2 t = p[0]; p[0] = p[2]; p[2] = t; }

void QuadraticSegment::moveStartPoint(Point2 to) {
    Vector2 origSDir = p[0]-p[1];
    Point2 origP1 = p[1];
    p[1] += crossProduct(p[0]-p[1], to-p[0])/crossProduct(p[0]-p[1], p[2]-p[1])*(p[2]-p[1]);
    p[0] = to;
    if (dotProduct(origSDir, p[0]-p[1]) < 0)
        p[1] = origP1;
}

void QuadraticSegment::moveEndPoint(Point2 to) {
    Vector2 origEDir = p[2]-p[1];
    Point2 origP1 = p[1];
    p[1] += crossProduct(p[2]-p[1], to-p[2])/crossProduct(p[2]-p[1], p[0]-p[1])*(p[0]-p[1]);
    p[2] = to;
    if (dotProduct(origEDir, p[2]-p[1]) < 0)
        p[1] = origP1;
}

void QuadraticSegment::splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const {
    Point2 a = point(1/3.), b = point(2/3.);
    part0 = new QuadraticSegment(p[0], mix(p[0], p[1], 1/3.), a, color);
    part1 = new QuadraticSegment(a, mix(mix(p[0], p[1], 5/9.), mix(p[1], p[2], 4/9.), .5), b, color);
    part2 = new QuadraticSegment(b, mix(p[1], p[2], 2/3.), p[2], color);
}

EdgeSegment *QuadraticSegment::convertToCubic() const {
    return new CubicSegment(p[0], mix(p[0], p[1], 2/3.), mix(p[1], p[2], 1/3.), p[2], color);
}

// ---------------------------------------------------------------------------------------------------------------------
// CubicSegment

CubicSegment::CubicSegment(Point2 p0, Point2 p1, Point2 p2, Point2 p3, EdgeColor edgeColor) : EdgeSegment(edgeColor) { p[0] = p0; p[1] = p1; p[2] = p2; p[3] = p3; }
CubicSegment *CubicSegment::clone() const { return new CubicSegment(p[0], p[1], p[2], p[3], color); }
int CubicSegment::type() const { return (int) EDGE_TYPE; }
const Point2 *CubicSegment::controlPoints() const { return p; }

Point2 CubicSegment::point(double param) const {
    Vector2 p12 = mix(p[1], p[2], param);
    return mix(mix(mix(p[0], p[1], param), p12, param), mix(p12, mix(p[2], p[3], param), param), param);
}

Vector2 CubicSegment::direction(double param) const {
    Vector2 tangent = mix(mix(p[1]-p[0], p[2]-p[1], param), mix(p[2]-p[1], p[3]-p[2], param), param);
    if (!tangent) {
        if (param == 0) return p[2]-p[0];
        if (param == 1) return p[3]-p[1];
    }
    return tangent;
}

Vector2 CubicSegment::directionChange(double param) const {
    return mix((p[2]-p[1])-(p[1]-p[0]), (p[3]-p[2])-(p[2]-p[1]), param);
}

void CubicSegment::bound(double &xMin, double &yMin, double &xMax, double &yMax) const {
    growBounds(p[0], xMin, yMin, xMax, yMax);
    growBounds(p[3], xMin, yMin, xMax, yMax);
    Vector2 a0 = p[1]-p[0];
    Vector2 a1 = 2*(p[2]-p[1]-a0);
    Vector2 a2 = p[3]-3*p[2]+3*p[1]-p[0];
    double params[2];
    int n = quadraticRoots(params, a2.x, a1.x, a0.x);
    for (int i = 0; i < n; ++i)
        if (params[i] > 0 && params[i] < 1)
            growBounds(point(params[i]), xMin, yMin, xMax, yMax);
    n = quadraticRoots(params, a2.y, a1.y, a0.y);
    for (int i = 0; i < n; ++i)
        if (params[i] > 0 && params[i] < 1)
            growBounds(point(params[i]), xMin, yMin, xMax, yMax);
}

void CubicSegment::reverse() {
    Point2 t = p[0]; p[0] = p[3]; p[3] = t;
    t = p[1]; p[1] = p[2]; p[2] = t;
}

void CubicSegment::moveStartPoint(Point2 to) { p[1] += to-p[0]; p[0] = to; }
void CubicSegment::moveEndPoint(Point2 to) { p[2] += to-p[3]; p[3] = to; }

void CubicSegment::splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const {
    // de Casteljau levels at 1/3 and 2/3
    Point2 q01 = mix(p[0], p[1], 1/3.), q12 = mix(p[1], p[2], 1/3.), q23 = mix(p[2], p[3], 1/3.);
    Point2 r01 = mix(p[0], p[1], 2/3.), r12 = mix(p[1], p[2], 2/3.), r23 = mix(p[2], p[3], 2/3.);
    Point2 a = point(1/3.), b = point(2/3.);
    part0 = new CubicSegment(p[0], p[0] == p[1] ? p[0] : q01, mix(q01, q12, 1/3.), a, color);
    part1 = new CubicSegment(a, mix(mix(q01, q12, 1/3.), mix(q12, q23, 1/3.), 2/3.), mix(mix(r01, r12, 2/3.), mix(r12, r23, 2/3.), 1/3.), b, color);
    part2 = new CubicSegment(b, mix(r12, r23, 2/3.), p[2] == p[3] ? p[3] : r23, p[3], color);
}

// ---------------------------------------------------------------------------------------------------------------------
// EdgeHolder

void EdgeHolder::swap(EdgeHolder &a, EdgeHolder &b) {
    EdgeSegment *t = a.edgeSegment;
    a.edgeSegment = b.edgeSegment;
    b.edgeSegment = t;
}

EdgeHolder::EdgeHolder(const EdgeHolder &orig) : edgeSegment(orig.edgeSegment ? orig.edgeSegment->clone() : NULL) { }
EdgeHolder::EdgeHolder(EdgeHolder &&orig) : edgeSegment(orig.edgeSegment) { orig.edgeSegment = NULL; }
EdgeHolder::~EdgeHolder() { delete edgeSegment; }

EdgeHolder &EdgeHolder::operator=(const EdgeHolder &orig) {
    if (this != &orig) {
        delete edgeSegment;
        edgeSegment = orig.edgeSegment ? orig.edgeSegment->clone() : NULL;
    }
    return *this;
}

EdgeHolder &EdgeHolder::operator=(EdgeHolder &&orig) {
    if (this != &orig) {
        delete edgeSegment;
        edgeSegment = orig.edgeSegment;
        orig.edgeSegment = NULL;
    }
    return *this;
}

// ---------------------------------------------------------------------------------------------------------------------
// Contour

void Contour::addEdge(const EdgeHolder &edge) { edges.push_back(edge); }
void Contour::addEdge(EdgeHolder &&edge) { edges.push_back((EdgeHolder &&) edge); }
EdgeHolder &Contour::addEdge() { edges.resize(edges.size()+1); return edges.back(); }

void Contour::bound(double &xMin, double &yMin, double &xMax, double &yMax) const {
    for (size_t i = 0; i < edges.size(); ++i)
        edges[i]->bound(xMin, yMin, xMax, yMax);
}

void Contour::boundMiters(double &xMin, double &yMin, double &xMax, double &yMax, double border, double miterLimit, int polarity) const {
    if (edges.empty())
        return;
    Vector2 prevDir = edges.back()->direction(1).normalize(true);
    for (size_t i = 0; i < edges.size(); ++i) {
        Vector2 dir = -edges[i]->direction(0).normalize(true);
        if (polarity*crossProduct(prevDir, dir) >= 0) {
            double miterLength = miterLimit;
            double q = .5*(1-dotProduct(prevDir, dir));
            if (q > 0)
                miterLength = min(1/std::sqrt(q), miterLimit);
            Point2 miter = edges[i]->point(0)+border*miterLength*(prevDir+dir).normalize(true);
            growBounds(miter, xMin, yMin, xMax, yMax);
        }
        prevDir = edges[i]->direction(1).normalize(true);
    }
}

int Contour::winding() const {
    if (edges.empty())
        return 0;
    // shoelace terms between consecutive sample points; 1- and 2-edge contours are sampled inside their edges
    struct Area { static double term(const Point2 &a, const Point2 &b) { return (b.x-a.x)*(a.y+b.y); } };
    double total = 0;
    if (edges.size() == 1) {
        Point2 a = edges[0]->point(0), b = edges[0]->point(1/3.), c = edges[0]->point(2/3.);
        total += Area::term(a, b);
        total += Area::term(b, c);
        total += Area::term(c, a);
    } else if (edges.size() == 2) {
        Point2 a = edges[0]->point(0), b = edges[0]->point(.5), c = edges[1]->point(0), d = edges[1]->point(.5);
        total += Area::term(a, b);
        total += Area::term(b, c);
        total += Area::term(c, d);
        total += Area::term(d, a);
    } else {
        Point2 prev = edges.back()->point(0);
        for (size_t i = 0; i < edges.size(); ++i) {
            Point2 cur = edges[i]->point(0);
            total += Area::term(prev, cur);
            prev = cur;
        }
    }
    return sign(total);
}

void Contour::reverse() {
    for (int i = (int) edges.size()/2; i > 0; --i)
        EdgeHolder::swap(edges[i-1], edges[edges.size()-i]);
    for (size_t i = 0; i < edges.size(); ++i)
        edges[i]->reverse();
}

// ---------------------------------------------------------------------------------------------------------------------
// Ordering of two curves that leave a shared corner in (numerically) the same direction: which one is on the left?
// Decided by the lowest-order non-vanishing term of cross(A(t)-P, B(t)-P) with both curves re-parameterised to unit speed.

namespace {

void reduceDegenerate(Point2 *cp, int &order) {
    if (order == 3 && (cp[1] == cp[0] || cp[1] == cp[3]) && (cp[2] == cp[0] || cp[2] == cp[3])) {
        cp[1] = cp[3];
        order = 1;
    }
    if (order == 2 && (cp[1] == cp[0] || cp[1] == cp[2])) {
        cp[1] = cp[2];
        order = 1;
    }
    if (order == 1 && cp[0] == cp[1])
        order = 0;
}

// corner[-before..-1] are the control points of the incoming curve (in order), corner[0] the shared point,
// corner[1..after] those of the outgoing curve.
int cornerCurveOrdering(const Point2 *corner, int before, int after) {
    if (!(before > 0 && after > 0))
        return 0;
    Vector2 a1, a2, a3, b1, b2, b3;
    a1 = corner[-1]-corner[0];
    b1 = corner[1]-corner[0];
    if (before >= 2)
        a2 = corner[-2]-corner[-1]-a1;
    if (after >= 2)
        b2 = corner[2]-corner[1]-b1;
    if (before >= 3) {
        a3 = corner[-3]-corner[-2]-(corner[-2]-corner[-1])-a2;
        a2 *= 3;
    }
    if (after >= 3) {
        b3 = corner[3]-corner[2]-(corner[2]-corner[1])-b2;
        b2 *= 3;
    }
    a1 *= before;
    b1 *= after;
    if (a1 && b1) {
        double as = a1.length(), bs = b1.length();
        double d;
        if ((d = as*crossProduct(a1, b2)+bs*crossProduct(a2, b1)))
            return sign(d);
        if ((d = as*as*crossProduct(a1, b3)+as*bs*crossProduct(a2, b2)+bs*bs*crossProduct(a3, b1)))
            return sign(d);
        if ((d = as*crossProduct(a2, b3)+bs*crossProduct(a3, b2)))
            return sign(d);
        return sign(crossProduct(a3, b3));
    }
    int flip = 1;
    if (a1) {
        // outgoing curve degenerate at the corner: swap roles and remember to negate
        b1 = a1;
        a1 = b2, b2 = a2, a2 = a1;
        a1 = b3, b3 = a3, a3 = a1;
        flip = -1;
    }
    if (b1) {
        double d;
        if ((d = crossProduct(a3, b1)))
            return flip*sign(d);
        if ((d = crossProduct(a2, b2)))
            return flip*sign(d);
        if ((d = crossProduct(a3, b2)))
            return flip*sign(d);
        if ((d = crossProduct(a2, b3)))
            return flip*sign(d);
        return flip*sign(crossProduct(a3, b3));
    }
    double d;
    if ((d = std::sqrt(a2.length())*crossProduct(a2, b3)+std::sqrt(b2.length())*crossProduct(a3, b2)))
        return sign(d);
    return sign(crossProduct(a3, b3));
}

int convergentOrdering(const EdgeSegment *a, const EdgeSegment *b) {
    Point2 buffer[12];
    Point2 *corner = buffer+4;
    Point2 *incoming = buffer+8;
    int aOrder = a->type(), bOrder = b->type();
    if (!(aOrder >= 1 && aOrder <= 3 && bOrder >= 1 && bOrder <= 3))
        return 0;
    for (int i = 0; i <= aOrder; ++i)
        incoming[i] = a->controlPoints()[i];
    for (int i = 0; i <= bOrder; ++i)
        corner[i] = b->controlPoints()[i];
    if (incoming[aOrder] != *corner)
        return 0;
    reduceDegenerate(incoming, aOrder);
    reduceDegenerate(corner, bOrder);
    for (int i = 0; i < aOrder; ++i)
        corner[i-aOrder] = incoming[i];
    return cornerCurveOrdering(corner, aOrder, bOrder);
}

// Nudges the control point next to the cusp sideways; quadratics are promoted to cubics first so only one end moves.
void pushControlPoint(EdgeHolder &edge, int end, Vector2 offset) {
    if (edge->type() == (int) QuadraticSegment::EDGE_TYPE)
        edge = static_cast<const QuadraticSegment *>(&*edge)->convertToCubic();
    if (edge->type() == (int) CubicSegment::EDGE_TYPE) {
        Point2 *p = static_cast<CubicSegment *>(&*edge)->p;
        if (end == 0)
            p[1] += (p[1]-p[0]).length()*offset;
        else
            p[2] += (p[2]-p[3]).length()*offset;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Shape

Shape::Shape() : inverseYAxis(false) { }
void Shape::addContour(const Contour &contour) { contours.push_back(contour); }
void Shape::addContour(Contour &&contour) { contours.push_back((Contour &&) contour); }
Contour &Shape::addContour() { contours.resize(contours.size()+1); return contours.back(); }

bool Shape::validate() const {
    for (size_t c = 0; c < contours.size(); ++c) {
        const std::vector<EdgeHolder> &edges = contours[c].edges;
        if (edges.empty())
            continue;
        Point2 corner = edges.back()->point(1);
        for (size_t i = 0; i < edges.size(); ++i) {
            if (!(const EdgeSegment *) edges[i])
                return false;
            if (edges[i]->point(0) != corner)
                return false;
            corner = edges[i]->point(1);
        }
    }
    return true;
}

void Shape::normalize() {
    const double cuspDot = MSDFGEN_CORNER_DOT_EPSILON-1;
    const double overshoot = 1.11111111111111111;
    for (size_t c = 0; c < contours.size(); ++c) {
        std::vector<EdgeHolder> &edges = contours[c].edges;
        if (edges.size() == 1) {
            EdgeSegment *parts[3] = { };
            edges[0]->splitInThirds(parts[0], parts[1], parts[2]);
            edges.clear();
            edges.push_back(EdgeHolder(parts[0]));
            edges.push_back(EdgeHolder(parts[1]));
            edges.push_back(EdgeHolder(parts[2]));
        } else if (!edges.empty()) {
            EdgeHolder *prevEdge = &edges.back();
            for (size_t i = 0; i < edges.size(); ++i) {
                EdgeHolder *edge = &edges[i];
                Vector2 prevDir = (*prevEdge)->direction(1).normalize();
                Vector2 curDir = (*edge)->direction(0).normalize();
                if (dotProduct(prevDir, curDir) < cuspDot) {
                    double factor = overshoot*std::sqrt(1-cuspDot*cuspDot)/cuspDot;
                    Vector2 axis = factor*(curDir-prevDir).normalize();
                    if (convergentOrdering(*prevEdge, *edge) < 0)
                        axis = -axis;
                    pushControlPoint(*prevEdge, 1, axis.getOrthogonal(true));
                    pushControlPoint(*edge, 0, axis.getOrthogonal(false));
                }
                prevEdge = edge;
            }
        }
    }
}

void Shape::bound(double &xMin, double &yMin, double &xMax, double &yMax) const {
    for (size_t c = 0; c < contours.size(); ++c)
        contours[c].bound(xMin, yMin, xMax, yMax);
}

void Shape::boundMiters(double &xMin, double &yMin, double &xMax, double &yMax, double border, double miterLimit, int polarity) const {
    for (size_t c = 0; c < contours.size(); ++c)
        contours[c].boundMiters(xMin, yMin, xMax, yMax, border, miterLimit, polarity);
}

Shape::Bounds Shape::getBounds(double border, double miterLimit, int polarity) const {
    const double huge = 1e240;
    Bounds bounds = { +huge, +huge, -huge, -huge };
    bound(bounds.l, bounds.b, bounds.r, bounds.t);
    if (border > 0) {
        bounds.l -= border, bounds.b -= border;
        bounds.r += border, bounds.t += border;
        if (miterLimit > 0)
            boundMiters(bounds.l, bounds.b, bounds.r, bounds.t, border, miterLimit, polarity);
    }
    return bounds;
}

int Shape::edgeCount() const {
    int total = 0;
    for (size_t c = 0; c < contours.size(); ++c)
        total += (int) contours[c].edges.size();
    return total;
}

YAxisOrientation Shape::getYAxisOrientation() const {
    return inverseYAxis ? MSDFGEN_Y_AXIS_NONDEFAULT_ORIENTATION : MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION;
}

void Shape::setYAxisOrientation(YAxisOrientation yAxisOrientation) {
    inverseYAxis = yAxisOrientation != MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION;
}

// ---------------------------------------------------------------------------------------------------------------------
// Edge colouring

namespace {

// The colour stream: a seed consumed bit by bit (or trit by trit) so that equal seeds give equal colourings.
struct ColorStream {
    unsigned long long seed;
    EdgeColor color;

    explicit ColorStream(unsigned long long s) : seed(s) {
        static const EdgeColor start[3] = { CYAN, MAGENTA, YELLOW };
        color = start[seed%3];
        seed /= 3;
    }
    // rotate the two-channel colour by one or two channel positions
    void advance() {
        int shifted = color<<(1+(int(seed)&1));
        seed >>= 1;
        color = EdgeColor((shifted|shifted>>3)&WHITE);
    }
    // like advance(), but never lands on a colour sharing two channels with `banned`
    void advance(EdgeColor banned) {
        EdgeColor shared = EdgeColor(color&banned);
        if (shared == RED || shared == GREEN || shared == BLUE)
            color = EdgeColor(shared^WHITE);
        else
            advance();
    }
};

bool sharpCorner(const Vector2 &inDir, const Vector2 &outDir, double crossThreshold) {
    return dotProduct(inDir, outDir) <= 0 || std::fabs(crossProduct(inDir, outDir)) > crossThreshold;
}

// -1 / 0 / +1 for the first / middle / last third of n positions, balanced so the three groups sum to zero
int thirdOf(int position, int n) {
    return int(3+2.875*position/(n-1)-1.4375+.5)-3;
}

double polylineLength(const EdgeSegment *edge) {
    double total = 0;
    Point2 prev = edge->point(0);
    for (int i = 1; i <= MSDFGEN_EDGE_LENGTH_PRECISION; ++i) {
        Point2 cur = edge->point(1./MSDFGEN_EDGE_LENGTH_PRECISION*i);
        total += (cur-prev).length();
        prev = cur;
    }
    return total;
}

// ---- the colouring of ONE closed contour, independent of what holds its geometry -------------------------------------------------
// Both entry paths (Shape objects; flat arrays of a batch) reduce a contour to the unit tangents at the two ends of every edge
// (tangents[2i] = direction(0).normalize(), tangents[2i+1] = direction(1).normalize() of edge i) and, for the ink-trap variant,
// a length estimate per edge; colours go out through a callback.  Colour corners are where consecutive tangents turn sharply.
struct ColorCorner {
    int index;
    double precedingLength;   // estimated length of the spline that ends at this corner (ink trap)
    bool minor;
    EdgeColor color;
};

// A contour with exactly one corner and fewer than three edges cannot carry three colours: the caller has to split its edges in
// thirds first.  paintContour then only reports the corner and the three colours (the stream has been advanced).
struct TeardropSplit {
    bool needed;
    int corner;
    EdgeColor colors[3];
};

template <typename LengthOf, typename SetColor>
TeardropSplit paintContour(int m, const Vector2 *tangents, int coloring, double crossThreshold, ColorStream &stream, std::vector<ColorCorner> &corners,
                           LengthOf lengthOf, SetColor setColor) {
    TeardropSplit split = { false, 0, { BLACK, BLACK, BLACK } };
    double splineLength = 0;
    corners.clear();
    Vector2 inDir = tangents[2*(m-1)+1];
    for (int i = 0; i < m; ++i) {
        if (sharpCorner(inDir, tangents[2*i], crossThreshold)) {
            ColorCorner corner = { i, splineLength, false, BLACK };
            corners.push_back(corner);
            splineLength = 0;
        }
        if (coloring == 2)
            splineLength += lengthOf(i);
        inDir = tangents[2*i+1];
    }
    const int cornerCount = (int) corners.size();
    if (cornerCount == 0) {
        // smooth loop: one colour
        stream.advance();
        for (int i = 0; i < m; ++i)
            setColor(i, stream.color);
    } else if (cornerCount == 1) {
        // teardrop: three colours around the loop, white in the middle
        stream.advance();
        split.colors[0] = stream.color;
        split.colors[1] = WHITE;
        stream.advance();
        split.colors[2] = stream.color;
        split.corner = corners[0].index;
        if (m >= 3) {
            for (int i = 0; i < m; ++i)
                setColor((split.corner+i)%m, split.colors[1+thirdOf(i, m)]);
        } else
            split.needed = true;
    } else if (coloring == 1) {
        // simple: a new colour at every corner; the last spline must also differ from the first
        int spline = 0;
        const int start = corners[0].index;
        stream.advance();
        const EdgeColor initialColor = stream.color;
        for (int i = 0; i < m; ++i) {
            const int index = (start+i)%m;
            if (spline+1 < cornerCount && corners[spline+1].index == index) {
                ++spline;
                stream.advance(EdgeColor((spline == cornerCount-1)*initialColor));
            }
            setColor(index, stream.color);
        }
    } else {
        // ink trap: corners that only enclose a short spline between two longer ones are "minor" and get the colour that
        // differs from both neighbours, so that the short spline does not cost a colour change of its own
        int majorCornerCount = cornerCount;
        if (cornerCount > 3) {
            corners[0].precedingLength += splineLength;
            for (int i = 0; i < cornerCount; ++i) {
                if (corners[i].precedingLength > corners[(i+1)%cornerCount].precedingLength &&
                    corners[(i+1)%cornerCount].precedingLength < corners[(i+2)%cornerCount].precedingLength) {
                    corners[i].minor = true;
                    --majorCornerCount;
                }
            }
        }
        EdgeColor initialColor = BLACK;
        for (int i = 0; i < cornerCount; ++i) {
            if (!corners[i].minor) {
                --majorCornerCount;
                stream.advance(EdgeColor(!majorCornerCount*initialColor));
                corners[i].color = stream.color;
                if (!initialColor)
                    initialColor = stream.color;
            }
        }
        for (int i = 0; i < cornerCount; ++i) {
            if (corners[i].minor) {
                EdgeColor nextColor = corners[(i+1)%cornerCount].color;
                corners[i].color = EdgeColor((stream.color&nextColor)^WHITE);
            } else
                stream.color = corners[i].color;
        }
        int spline = 0;
        const int start = corners[0].index;
        stream.color = corners[0].color;
        for (int i = 0; i < m; ++i) {
            const int index = (start+i)%m;
            if (spline+1 < cornerCount && corners[spline+1].index == index)
                stream.color = corners[++spline].color;
            setColor(index, stream.color);
        }
    }
    return split;
}

// One- or two-edge teardrop of a Shape: every edge becomes three, the six (or three) parts take the three colours in order
// starting at the corner.
void splitTeardrop(Contour &contour, const TeardropSplit &split) {
    std::vector<EdgeHolder> &edges = contour.edges;
    EdgeSegment *parts[7] = { };
    const int corner = split.corner;
    edges[0]->splitInThirds(parts[0+3*corner], parts[1+3*corner], parts[2+3*corner]);
    if (edges.size() >= 2) {
        edges[1]->splitInThirds(parts[3-3*corner], parts[4-3*corner], parts[5-3*corner]);
        for (int i = 0; i < 6; ++i)
            parts[i]->color = split.colors[i/2];
    } else {
        for (int i = 0; i < 3; ++i)
            parts[i]->color = split.colors[i];
    }
    edges.clear();
    for (int i = 0; parts[i]; ++i)
        edges.push_back(EdgeHolder(parts[i]));
}

// edgeColoringSimple (coloring 1) / edgeColoringInkTrap (2) of a Shape: the contours' tangents, then the shared core.
void colorShape(Shape &shape, int coloring, double angleThreshold, unsigned long long seed) {
    const double crossThreshold = std::sin(angleThreshold);
    ColorStream stream(seed);
    std::vector<ColorCorner> corners;
    std::vector<Vector2> tangents;
    for (size_t c = 0; c < shape.contours.size(); ++c) {
        Contour &contour = shape.contours[c];
        std::vector<EdgeHolder> &edges = contour.edges;
        const int m = (int) edges.size();
        if (m == 0)
            continue;
        tangents.resize(2*size_t(m));
        for (int i = 0; i < m; ++i) {
            tangents[2*size_t(i)] = edges[size_t(i)]->direction(0).normalize();
            tangents[2*size_t(i)+1] = edges[size_t(i)]->direction(1).normalize();
        }
        const TeardropSplit split = paintContour(m, tangents.data(), coloring, crossThreshold, stream, corners,
            [&edges](int i) { return polylineLength(edges[size_t(i)]); },
            [&edges](int i, EdgeColor color) { edges[size_t(i)]->color = color; });
        if (split.needed)
            splitTeardrop(contour, split);
    }
}

}  // namespace

void edgeColoringSimple(Shape &shape, double angleThreshold, unsigned long long seed) { colorShape(shape, 1, angleThreshold, seed); }
void edgeColoringInkTrap(Shape &shape, double angleThreshold, unsigned long long seed) { colorShape(shape, 2, angleThreshold, seed); }

// ---------------------------------------------------------------------------------------------------------------------
// Allocation-free preparation of one glyph on flat arrays (flat_prep.h).  Segments are materialised as STACK objects of the
// same classes the Shape path uses, so every direction / point / length is computed by the very same member functions, and the
// colours come from the same paintContour as the Shape path's.

namespace b200 {

namespace {

// A segment of any order living on the stack (no heap, no clone)
struct StackEdge {
    int order;
    Point2 p[4];
    StackEdge(const double *q, int o) : order(o) {
        for (int k = 0; k < 4; ++k)
            p[k] = Point2(q[2*k], q[2*k+1]);
    }
    Vector2 direction(double t) const {
        switch (order) {
            case 1: return LinearSegment(p[0], p[1]).direction(t);
            case 2: return QuadraticSegment(p[0], p[1], p[2]).direction(t);
            default: return CubicSegment(p[0], p[1], p[2], p[3]).direction(t);
        }
    }
    double polylineLength() const {
        switch (order) {
            case 1: { LinearSegment e(p[0], p[1]); return msdfgen::polylineLength(&e); }
            case 2: { QuadraticSegment e(p[0], p[1], p[2]); return msdfgen::polylineLength(&e); }
            default: { CubicSegment e(p[0], p[1], p[2], p[3]); return msdfgen::polylineLength(&e); }
        }
    }
};

}  // namespace

bool prepareGlyphFlat(const uint32_t *cbegin, uint32_t c0, uint32_t c1, const double *points, const uint8_t *orders, const uint8_t *inColors,
                      bool doNormalize, int coloring, double angleThreshold, unsigned long long seed, uint8_t *outColors) {
    const double cuspDot = MSDFGEN_CORNER_DOT_EPSILON-1;
    const double crossThreshold = std::sin(angleThreshold);
    // ---- pass 1: does anything change the geometry?  (Shape::normalize, paintTeardrop's split)
    // The unit tangents at both ends of every edge are kept for pass 2 (they are the bulk of the arithmetic).
    static thread_local std::vector<Vector2> tangents;      // [2*k] = direction(0).normalize(), [2*k+1] = direction(1).normalize()
    const uint32_t s0 = cbegin[c0];
    tangents.resize(2*size_t(cbegin[c1]-s0));
    for (uint32_t s = s0; s < cbegin[c1]; ++s) {
        const StackEdge edge(points+8*size_t(s), orders[s]);
        tangents[2*size_t(s-s0)] = edge.direction(0).normalize();
        tangents[2*size_t(s-s0)+1] = edge.direction(1).normalize();
    }
    for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t first = cbegin[c], n = cbegin[c+1]-first;
        if (n == 0)
            continue;
        if (doNormalize && n == 1)
            return false;                                   // single-edge contour: split in thirds
        int cornerCount = 0;
        Vector2 inDir = tangents[2*size_t(first+n-1-s0)+1];
        for (uint32_t i = 0; i < n; ++i) {
            const Vector2 outDir = tangents[2*size_t(first+i-s0)];
            if (doNormalize && n > 1 && dotProduct(inDir, outDir) < cuspDot)
                return false;                               // cusp: control points get pushed apart
            if (coloring && sharpCorner(inDir, outDir, crossThreshold))
                ++cornerCount;
            inDir = tangents[2*size_t(first+i-s0)+1];
        }
        if (coloring && cornerCount == 1 && n < 3)
            return false;                                   // teardrop of one or two edges: split to make room for three colours
    }
    // ---- pass 2: colours
    if (!coloring) {
        for (uint32_t s = cbegin[c0]; s < cbegin[c1]; ++s)
            outColors[s] = inColors ? inColors[s] : (uint8_t) WHITE;
        return true;
    }
    static thread_local std::vector<ColorCorner> corners;
    ColorStream stream(seed);
    for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t first = cbegin[c];
        const int m = int(cbegin[c+1]-first);
        if (m == 0)
            continue;
        uint8_t *colors = outColors+first;
        // (pass 1 has sent every contour that would need the teardrop split down the general path)
        paintContour(m, &tangents[2*size_t(first-s0)], coloring, crossThreshold, stream, corners,
            [=](int i) { return StackEdge(points+8*size_t(first+i), orders[first+i]).polylineLength(); },
            [=](int i, EdgeColor color) { colors[i] = (uint8_t) color; });
    }
    return true;
}

}  // namespace b200

const double ErrorCorrectionConfig::defaultMinDeviationRatio = 1.11111111111111111;
const double ErrorCorrectionConfig::defaultMinImproveRatio = 1.11111111111111111;

}  // namespace msdfgen
