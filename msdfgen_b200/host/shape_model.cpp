// shape_model.cpp — the msdfgen-compatible shape objects of msdfgen.h as an adaptor over the flat geometry.
//
// Callers written against msdfgen build Shape / Contour / EdgeHolder / EdgeSegment objects and call member functions on them.
// Here those classes own nothing but storage: every geometric question is answered by flat_geom.h on the segment's control
// points (the same functions the batch ingest uses on wire arrays), and Shape::normalize / edgeColoring* flatten a contour,
// run the list algorithms of flat_prep.cpp and write the result back.  Interface restated from /root/reference/core/:
// edge-segments.h, EdgeHolder.h, Contour.h, Shape.h, edge-coloring.h; results equal the compiled reference's exactly
// (tests/test_host_model.py, tests/test_dropin.py).
#include "msdfgen.h"
#include "flat_prep.h"

namespace msdfgen {

// ---------------------------------------------------------------------------------------------------------------------------------
// objects <-> flat segments

namespace {

// An edge of one of the three built-in classes as a flat segment (false for a null edge or a foreign subclass).
bool flatten(const EdgeSegment *edge, flat::Seg &seg) {
    const int order = edge ? edge->type() : 0;
    if (!(order >= 1 && order <= 3))
        return false;
    const Point2 *cp = edge->controlPoints();
    seg = flat::makeSeg(order, cp[0], cp[1], order >= 2 ? cp[2] : Point2(), order >= 3 ? cp[3] : Point2(), edge->color);
    return true;
}

bool flatten(const Contour &contour, std::vector<flat::Seg> &out) {
    out.resize(contour.edges.size());
    for (size_t i = 0; i < out.size(); ++i)
        if (!flatten(contour.edges[i], out[i]))
            return false;
    return true;
}

// No collapsing of degenerate curves here (that is EdgeSegment::create's job): the order is kept.
EdgeSegment *materialize(const flat::Seg &seg) {
    switch (seg.order) {
        case 1: return new LinearSegment(seg.p[0], seg.p[1], seg.color);
        case 2: return new QuadraticSegment(seg.p[0], seg.p[1], seg.p[2], seg.color);
        default: return new CubicSegment(seg.p[0], seg.p[1], seg.p[2], seg.p[3], seg.color);
    }
}

// Stores a processed edge list in the contour.  Edges that kept their order are updated in place, so pointers a caller holds to
// them stay valid exactly where they would with msdfgen (which edits such edges in place too); promoted edges are replaced, and
// a contour whose edge count changed is rebuilt.
void writeBack(Contour &contour, const std::vector<flat::Seg> &segs) {
    std::vector<EdgeHolder> &edges = contour.edges;
    if (edges.size() != segs.size()) {
        edges.clear();
        for (size_t i = 0; i < segs.size(); ++i)
            edges.push_back(EdgeHolder(materialize(segs[i])));
        return;
    }
    for (size_t i = 0; i < segs.size(); ++i) {
        if (edges[i]->type() != segs[i].order) {
            edges[i] = EdgeHolder(materialize(segs[i]));
            continue;
        }
        Point2 *cp = const_cast<Point2 *>(edges[i]->controlPoints());
        for (int k = 0; k <= segs[i].order; ++k)
            cp[k] = segs[i].p[k];
        edges[i]->color = segs[i].color;
    }
}

void splitInThirds(const flat::Seg &seg, EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) {
    flat::Seg parts[3];
    flat::splitThirds(seg, parts);
    part0 = materialize(parts[0]);
    part1 = materialize(parts[1]);
    part2 = materialize(parts[2]);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------------
// EdgeSegment and its three classes

// The factory lowers degenerate curves to the lowest order that represents them.
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

LinearSegment::LinearSegment(Point2 p0, Point2 p1, EdgeColor edgeColor) : EdgeSegment(edgeColor) { p[0] = p0; p[1] = p1; }
QuadraticSegment::QuadraticSegment(Point2 p0, Point2 p1, Point2 p2, EdgeColor edgeColor) : EdgeSegment(edgeColor) { p[0] = p0; p[1] = p1; p[2] = p2; }
CubicSegment::CubicSegment(Point2 p0, Point2 p1, Point2 p2, Point2 p3, EdgeColor edgeColor) : EdgeSegment(edgeColor) { p[0] = p0; p[1] = p1; p[2] = p2; p[3] = p3; }

LinearSegment *LinearSegment::clone() const { return new LinearSegment(*this); }
QuadraticSegment *QuadraticSegment::clone() const { return new QuadraticSegment(*this); }
CubicSegment *CubicSegment::clone() const { return new CubicSegment(*this); }

// Everything else is the flat geometry applied to the member array p[0..ORDER].
#define B200_SEGMENT_ADAPTOR(Class, ORDER) \
    int Class::type() const { return ORDER; } \
    const Point2 *Class::controlPoints() const { return p; } \
    Point2 Class::point(double param) const { return flat::pointAt(p, ORDER, param); } \
    Vector2 Class::direction(double param) const { return flat::directionAt(p, ORDER, param); } \
    Vector2 Class::directionChange(double param) const { return flat::directionChangeAt(p, ORDER, param); } \
    void Class::bound(double &xMin, double &yMin, double &xMax, double &yMax) const { flat::growBounds(p, ORDER, xMin, yMin, xMax, yMax); } \
    void Class::reverse() { flat::reverse(p, ORDER); } \
    void Class::moveStartPoint(Point2 to) { flat::moveEnd(p, ORDER, 0, to); } \
    void Class::moveEndPoint(Point2 to) { flat::moveEnd(p, ORDER, 1, to); } \
    void Class::splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const { \
        flat::Seg seg; \
        flatten(this, seg); \
        msdfgen::splitInThirds(seg, part0, part1, part2); \
    }

B200_SEGMENT_ADAPTOR(LinearSegment, 1)
B200_SEGMENT_ADAPTOR(QuadraticSegment, 2)
B200_SEGMENT_ADAPTOR(CubicSegment, 3)

#undef B200_SEGMENT_ADAPTOR

double LinearSegment::length() const { return flat::arcLength(p, 1); }
double QuadraticSegment::length() const { return flat::arcLength(p, 2); }

EdgeSegment *QuadraticSegment::convertToCubic() const {
    flat::Seg seg;
    flatten(this, seg);
    return materialize(flat::toCubic(seg));
}

// ---------------------------------------------------------------------------------------------------------------------------------
// EdgeHolder: owning pointer, deep copy

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
        EdgeHolder copy(orig);
        swap(*this, copy);
    }
    return *this;
}

EdgeHolder &EdgeHolder::operator=(EdgeHolder &&orig) {
    if (this != &orig) {
        EdgeHolder taken((EdgeHolder &&) orig);
        swap(*this, taken);
    }
    return *this;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Contour

void Contour::addEdge(const EdgeHolder &edge) { edges.push_back(edge); }
void Contour::addEdge(EdgeHolder &&edge) { edges.push_back((EdgeHolder &&) edge); }
EdgeHolder &Contour::addEdge() { edges.resize(edges.size()+1); return edges.back(); }

void Contour::bound(double &xMin, double &yMin, double &xMax, double &yMax) const {
    for (size_t i = 0; i < edges.size(); ++i)
        edges[i]->bound(xMin, yMin, xMax, yMax);
}

void Contour::boundMiters(double &xMin, double &yMin, double &xMax, double &yMax, double border, double miterLimit, int polarity) const {
    std::vector<flat::Seg> segs;
    if (flatten(*this, segs))
        flat::growMiterBounds(segs.data(), (int) segs.size(), xMin, yMin, xMax, yMax, border, miterLimit, polarity);
}

int Contour::winding() const {
    std::vector<flat::Seg> segs;
    return flatten(*this, segs) ? flat::winding(segs.data(), (int) segs.size()) : 0;
}

void Contour::reverse() {
    for (size_t i = 0, j = edges.size(); i+1 < j; ++i)
        EdgeHolder::swap(edges[i], edges[--j]);
    for (size_t i = 0; i < edges.size(); ++i)
        edges[i]->reverse();
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Shape

Shape::Shape() : inverseYAxis(false) { }
void Shape::addContour(const Contour &contour) { contours.push_back(contour); }
void Shape::addContour(Contour &&contour) { contours.push_back((Contour &&) contour); }
Contour &Shape::addContour() { contours.resize(contours.size()+1); return contours.back(); }

// every edge present, and starting where its predecessor ended
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
    std::vector<flat::Seg> segs;
    for (size_t c = 0; c < contours.size(); ++c)
        if (flatten(contours[c], segs) && flat::normalizeContour(segs))
            writeBack(contours[c], segs);
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
    Bounds box = { +huge, +huge, -huge, -huge };
    bound(box.l, box.b, box.r, box.t);
    if (border > 0) {
        box.l -= border, box.b -= border;
        box.r += border, box.t += border;
        if (miterLimit > 0)
            boundMiters(box.l, box.b, box.r, box.t, border, miterLimit, polarity);
    }
    return box;
}

int Shape::edgeCount() const {
    size_t total = 0;
    for (size_t c = 0; c < contours.size(); ++c)
        total += contours[c].edges.size();
    return (int) total;
}

YAxisOrientation Shape::getYAxisOrientation() const {
    return inverseYAxis ? MSDFGEN_Y_AXIS_NONDEFAULT_ORIENTATION : MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION;
}

void Shape::setYAxisOrientation(YAxisOrientation yAxisOrientation) {
    inverseYAxis = yAxisOrientation != MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Edge colouring of a Shape: contour by contour through the list form (one colour stream across the contours)

static void colorShape(Shape &shape, int strategy, double angleThreshold, unsigned long long seed) {
    flat::Coloring coloring(strategy, angleThreshold, seed);
    std::vector<flat::Seg> segs;
    for (size_t c = 0; c < shape.contours.size(); ++c) {
        if (!flatten(shape.contours[c], segs) || segs.empty())
            continue;
        coloring.colorContour(segs);
        writeBack(shape.contours[c], segs);
    }
}

void edgeColoringSimple(Shape &shape, double angleThreshold, unsigned long long seed) { colorShape(shape, 1, angleThreshold, seed); }
void edgeColoringInkTrap(Shape &shape, double angleThreshold, unsigned long long seed) { colorShape(shape, 2, angleThreshold, seed); }

const double ErrorCorrectionConfig::defaultMinDeviationRatio = 1.11111111111111111;
const double ErrorCorrectionConfig::defaultMinImproveRatio = 1.11111111111111111;

}  // namespace msdfgen
