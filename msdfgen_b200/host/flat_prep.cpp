// flat_prep.cpp — Shape::normalize and the edge colourings on flat segment lists (see flat_prep.h).
//
// Results must equal the reference's exactly — the same doubles, the same colours: they are the INPUT of the device path whose
// output is compared bit for bit; tests/test_host_model.py checks both entry paths against the compiled reference.  Behaviour
// restated from /root/reference/core/: Shape.cpp:29-92 (normalize, deconvergeEdge), convergent-curve-ordering.cpp (which of two
// curves leaving a cusp is on the left), edge-coloring.cpp:19-258 (colour switching, corner detection, the two strategies).
#include "flat_prep.h"

namespace msdfgen {
namespace flat {

// ---------------------------------------------------------------------------------------------------------------------------------
// Shape::normalize

namespace {

// A curve whose inner control points coincide with its ends is really of lower order at the corner: strip that.
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

// Two curves meet at corner[0] in (numerically) opposite directions; corner[-before..-1] are the control points of the incoming
// curve in order, corner[1..after] those of the outgoing one.  +1 / -1 by the side the outgoing curve leaves on: the sign of the
// lowest-order non-vanishing term of cross(A(t)-P, B(t)-P) with both curves re-parameterised to unit speed at the corner.
int orderingAtCorner(const Point2 *corner, int before, int after) {
    if (!(before > 0 && after > 0))
        return 0;
    // first, second and third forward differences on both sides (scaled like Taylor coefficients)
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
    double d;
    if (a1 && b1) {
        const double as = a1.length(), bs = b1.length();
        if ((d = as*crossProduct(a1, b2)+bs*crossProduct(a2, b1)))
            return sign(d);
        if ((d = as*as*crossProduct(a1, b3)+as*bs*crossProduct(a2, b2)+bs*bs*crossProduct(a3, b1)))
            return sign(d);
        if ((d = as*crossProduct(a2, b3)+bs*crossProduct(a3, b2)))
            return sign(d);
        return sign(crossProduct(a3, b3));
    }
    // at least one side starts with zero velocity
    int flip = 1;
    if (a1) {
        // it is the outgoing one: exchange the roles of the two curves (which negates the answer)
        b1 = a1;
        a1 = b2, b2 = a2, a2 = a1;
        a1 = b3, b3 = a3, a3 = a1;
        flip = -1;
    }
    if (b1) {
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
    if ((d = std::sqrt(a2.length())*crossProduct(a2, b3)+std::sqrt(b2.length())*crossProduct(a3, b2)))
        return sign(d);
    return sign(crossProduct(a3, b3));
}

int convergentOrdering(const Seg &a, const Seg &b) {
    // buffer[4] is the corner; the incoming curve is assembled in the upper part first because reduceDegenerate may shorten it
    Point2 buffer[12];
    Point2 *corner = buffer+4;
    Point2 *incoming = buffer+8;
    int aOrder = a.order, bOrder = b.order;
    if (!(aOrder >= 1 && aOrder <= 3 && bOrder >= 1 && bOrder <= 3))
        return 0;
    for (int i = 0; i <= aOrder; ++i)
        incoming[i] = a.p[i];
    for (int i = 0; i <= bOrder; ++i)
        corner[i] = b.p[i];
    if (incoming[aOrder] != *corner)
        return 0;
    reduceDegenerate(incoming, aOrder);
    reduceDegenerate(corner, bOrder);
    for (int i = 0; i < aOrder; ++i)
        corner[i-aOrder] = incoming[i];
    return orderingAtCorner(corner, aOrder, bOrder);
}

// Nudges the control point next to the cusp sideways, by `offset` times the handle length.  Quadratics are promoted to cubics
// first so that only this end of the curve moves; lines stay as they are.
void pushControlPoint(Seg &edge, int end, Vector2 offset) {
    edge = toCubic(edge);
    if (edge.order == 3) {
        Point2 *p = edge.p;
        if (end == 0)
            p[1] += (p[1]-p[0]).length()*offset;
        else
            p[2] += (p[2]-p[3]).length()*offset;
    }
}

}  // namespace

bool normalizeContour(std::vector<Seg> &edges) {
    if (edges.size() == 1) {
        Seg parts[3];
        splitThirds(edges[0], parts);
        edges.assign(parts, parts+3);
        return true;
    }
    const double cuspDot = MSDFGEN_CORNER_DOT_EPSILON-1;
    const double overshoot = 1.11111111111111111;
    bool changed = false;
    for (size_t i = 0, prev = edges.size()-1; i < edges.size(); prev = i++) {
        const Vector2 prevDir = directionAt(edges[prev].p, edges[prev].order, 1).normalize();
        const Vector2 curDir = directionAt(edges[i].p, edges[i].order, 0).normalize();
        if (dotProduct(prevDir, curDir) < cuspDot) {
            const double factor = overshoot*std::sqrt(1-cuspDot*cuspDot)/cuspDot;
            Vector2 axis = factor*(curDir-prevDir).normalize();
            if (convergentOrdering(edges[prev], edges[i]) < 0)
                axis = -axis;
            pushControlPoint(edges[prev], 1, axis.getOrthogonal(true));
            pushControlPoint(edges[i], 0, axis.getOrthogonal(false));
            changed = true;
        }
    }
    return changed;
}

void endTangents(const Seg *edges, int n, Vector2 *out) {
    for (int i = 0; i < n; ++i) {
        out[2*i] = directionAt(edges[i].p, edges[i].order, 0).normalize();
        out[2*i+1] = directionAt(edges[i].p, edges[i].order, 1).normalize();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Edge colouring

// The colour stream: a seed consumed trit, then bit by bit, so that equal seeds give equal colourings.
void Coloring::reset(int newStrategy, double angleThreshold, unsigned long long newSeed) {
    static const EdgeColor start[3] = { CYAN, MAGENTA, YELLOW };
    strategy = newStrategy;
    crossThreshold = std::sin(angleThreshold);
    color = start[newSeed%3];
    seed = newSeed/3;
}

// rotate the two-channel colour by one or two channel positions
void Coloring::advance() {
    const int shifted = color<<(1+(int(seed)&1));
    seed >>= 1;
    color = EdgeColor((shifted|shifted>>3)&WHITE);
}

// like advance(), but never lands on a colour sharing two channels with `banned`
void Coloring::advance(EdgeColor banned) {
    const EdgeColor shared = EdgeColor(color&banned);
    if (shared == RED || shared == GREEN || shared == BLUE)
        color = EdgeColor(shared^WHITE);
    else
        advance();
}

bool Coloring::sharp(const Vector2 &inDir, const Vector2 &outDir) const {
    return dotProduct(inDir, outDir) <= 0 || std::fabs(crossProduct(inDir, outDir)) > crossThreshold;
}

// -1 / 0 / +1 for the first / middle / last third of n positions, balanced so the three groups sum to zero
static int thirdOf(int position, int n) {
    return int(3+2.875*position/(n-1)-1.4375+.5)-3;
}

bool Coloring::keepsGeometry(int m, const Vector2 *t) const {
    if (m >= 3)
        return true;
    int cornerCount = 0;
    Vector2 inDir = t[2*(m-1)+1];
    for (int i = 0; i < m; ++i) {
        cornerCount += sharp(inDir, t[2*i]);
        inDir = t[2*i+1];
    }
    return cornerCount != 1;
}

// The colouring of one closed contour of m edges, independent of what holds its geometry: all it needs are the unit tangents at
// the two ends of every edge and, for the ink-trap strategy, a length estimate per edge; colours go out through a callback.
// Colour changes happen at corners — where consecutive tangents turn sharply.
template <typename LengthOf, typename SetColor>
Coloring::Teardrop Coloring::paint(int m, const Vector2 *t, LengthOf lengthOf, SetColor setColor) {
    Teardrop teardrop = { false, 0, { BLACK, BLACK, BLACK } };
    double splineLength = 0;
    corners.clear();
    Vector2 inDir = t[2*(m-1)+1];
    for (int i = 0; i < m; ++i) {
        if (sharp(inDir, t[2*i])) {
            Corner corner = { i, splineLength, false, BLACK };
            corners.push_back(corner);
            splineLength = 0;
        }
        if (strategy == 2)
            splineLength += lengthOf(i);
        inDir = t[2*i+1];
    }
    const int cornerCount = (int) corners.size();
    if (cornerCount == 0) {
        // smooth loop: one colour
        advance();
        for (int i = 0; i < m; ++i)
            setColor(i, color);
    } else if (cornerCount == 1) {
        // teardrop: three colours around the loop, white in the middle
        advance();
        teardrop.colors[0] = color;
        teardrop.colors[1] = WHITE;
        advance();
        teardrop.colors[2] = color;
        teardrop.corner = corners[0].index;
        if (m >= 3) {
            for (int i = 0; i < m; ++i)
                setColor((teardrop.corner+i)%m, teardrop.colors[1+thirdOf(i, m)]);
        } else
            teardrop.splitNeeded = true;
    } else if (strategy == 1) {
        // simple: a new colour at every corner; the last spline must also differ from the first
        int spline = 0;
        const int start = corners[0].index;
        advance();
        const EdgeColor initialColor = color;
        for (int i = 0; i < m; ++i) {
            const int index = (start+i)%m;
            if (spline+1 < cornerCount && corners[spline+1].index == index) {
                ++spline;
                advance(EdgeColor((spline == cornerCount-1)*initialColor));
            }
            setColor(index, color);
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
                advance(EdgeColor(!majorCornerCount*initialColor));
                corners[i].color = color;
                if (!initialColor)
                    initialColor = color;
            }
        }
        for (int i = 0; i < cornerCount; ++i) {
            if (corners[i].minor) {
                const EdgeColor nextColor = corners[(i+1)%cornerCount].color;
                corners[i].color = EdgeColor((color&nextColor)^WHITE);
            } else
                color = corners[i].color;
        }
        int spline = 0;
        const int start = corners[0].index;
        color = corners[0].color;
        for (int i = 0; i < m; ++i) {
            const int index = (start+i)%m;
            if (spline+1 < cornerCount && corners[spline+1].index == index)
                color = corners[++spline].color;
            setColor(index, color);
        }
    }
    return teardrop;
}

bool Coloring::colorContour(std::vector<Seg> &edges) {
    const int m = (int) edges.size();
    if (m == 0)
        return false;
    tangents.resize(2*size_t(m));
    endTangents(edges.data(), m, tangents.data());
    const Teardrop teardrop = paint(m, tangents.data(),
        [&edges](int i) { return polylineLength(edges[size_t(i)].p, edges[size_t(i)].order, MSDFGEN_EDGE_LENGTH_PRECISION); },
        [&edges](int i, EdgeColor c) { edges[size_t(i)].color = c; });
    if (!teardrop.splitNeeded)
        return false;
    // one or two edges, one corner: every edge becomes three and the parts take the three colours in order from the corner on
    Seg parts[6];
    int count = 3;
    splitThirds(edges[0], parts+3*teardrop.corner);
    if (m >= 2) {
        splitThirds(edges[1], parts+3-3*teardrop.corner);
        count = 6;
    }
    for (int i = 0; i < count; ++i)
        parts[i].color = teardrop.colors[i/(count/3)];
    edges.assign(parts, parts+count);
    return true;
}

void Coloring::colorContour(const double *points, const uint8_t *orders, uint32_t first, int m, const Vector2 *t, uint8_t *outColors) {
    if (m == 0)
        return;
    uint8_t *colors = outColors+first;
    paint(m, t,
        [=](int i) {
            const Seg edge = loadSeg(points, orders, NULL, size_t(first)+size_t(i));
            return polylineLength(edge.p, edge.order, MSDFGEN_EDGE_LENGTH_PRECISION);
        },
        [=](int i, EdgeColor c) { colors[i] = (uint8_t) c; });
}

// ---------------------------------------------------------------------------------------------------------------------------------
// One glyph of a flat batch

bool prepareGlyph(const uint32_t *cbegin, uint32_t c0, uint32_t c1, const double *points, const uint8_t *orders, const uint8_t *inColors,
                  bool doNormalize, int coloring, double angleThreshold, unsigned long long seed, uint8_t *outColors,
                  std::vector<Seg> &changed, std::vector<uint32_t> &changedCbegin, bool allowFastPath) {
    const double cuspDot = MSDFGEN_CORNER_DOT_EPSILON-1;
    const uint32_t s0 = cbegin[c0], s1 = cbegin[c1];
    static thread_local Coloring painter(0, 0, 0);
    painter.reset(coloring, angleThreshold, seed);
    changed.clear();
    changedCbegin.clear();
    // ---- does anything change the geometry?  The unit tangents at both ends of every edge decide that, and they are the bulk of
    // the colouring's arithmetic as well: computed once, kept for the colours.
    static thread_local std::vector<Vector2> tangents;
    bool keeps = allowFastPath;
    if (keeps) {
        tangents.resize(2*size_t(s1-s0));
        for (uint32_t s = s0; s < s1; ++s) {
            const Seg edge = loadSeg(points, orders, NULL, s);
            endTangents(&edge, 1, &tangents[2*size_t(s-s0)]);
        }
    }
    for (uint32_t c = c0; c < c1 && keeps; ++c) {
        const uint32_t first = cbegin[c], n = cbegin[c+1]-first;
        if (n == 0)
            continue;
        const Vector2 *t = &tangents[2*size_t(first-s0)];
        if (doNormalize) {
            keeps = n > 1;                                      // a single-edge contour is split in thirds
            Vector2 inDir = t[2*(n-1)+1];
            for (uint32_t i = 0; i < n && keeps; ++i) {
                keeps = !(dotProduct(inDir, t[2*i]) < cuspDot);     // cusp: control points get pushed apart
                inDir = t[2*i+1];
            }
        }
        if (keeps && coloring)
            keeps = painter.keepsGeometry((int) n, t);          // tiny teardrop: split to make room for three colours
    }
    if (keeps) {
        // ---- colours only, straight from / to the arrays
        if (!coloring) {
            for (uint32_t s = s0; s < s1; ++s)
                outColors[s] = inColors ? inColors[s] : (uint8_t) WHITE;
        } else {
            for (uint32_t c = c0; c < c1; ++c)
                painter.colorContour(points, orders, cbegin[c], int(cbegin[c+1]-cbegin[c]), &tangents[2*size_t(cbegin[c]-s0)], outColors);
        }
        return true;
    }
    // ---- general path: contour by contour through the list form
    static thread_local std::vector<Seg> edges;
    changedCbegin.push_back(0);
    for (uint32_t c = c0; c < c1; ++c) {
        edges.clear();
        for (uint32_t s = cbegin[c]; s < cbegin[c+1]; ++s)
            edges.push_back(loadSeg(points, orders, inColors, s));
        if (doNormalize && !edges.empty())
            normalizeContour(edges);
        if (coloring)
            painter.colorContour(edges);
        changed.insert(changed.end(), edges.begin(), edges.end());
        changedCbegin.push_back((uint32_t) changed.size());
    }
    return false;
}

}  // namespace flat
}  // namespace msdfgen
