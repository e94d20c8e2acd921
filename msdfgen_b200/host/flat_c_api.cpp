// flat_c_api.cpp — extern "C" access to the host shape preparation for callers that keep shapes as flat arrays
// (the Python package, bench.py): builds msdfgen::Shape objects from a flat batch, runs Shape::normalize() and an edge
// colouring on every glyph, and flattens the result again.  Host-only (no device work); this is the batch-ingest step that
// precedes b200msdf_generate_batch (SURVEY.md §8f "next" #1).
#include "msdfgen.h"
#include "flat_prep.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace msdfgen;

namespace {

void buildShape(Shape &shape, const uint32_t *cbegin, uint32_t c0, uint32_t c1, const double *points, const uint8_t *orders, const uint8_t *colors, bool yDown) {
    shape.contours.clear();
    shape.inverseYAxis = yDown;
    for (uint32_t c = c0; c < c1; ++c) {
        Contour &contour = shape.addContour();
        for (uint32_t s = cbegin[c]; s < cbegin[c+1]; ++s) {
            const double *p = points+8*size_t(s);
            EdgeColor col = EdgeColor(colors ? colors[s] : 7);
            switch (orders[s]) {
                case 1: contour.addEdge(EdgeHolder(new LinearSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), col))); break;
                case 2: contour.addEdge(EdgeHolder(new QuadraticSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), col))); break;
                default: contour.addEdge(EdgeHolder(new CubicSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), Point2(p[6], p[7]), col))); break;
            }
        }
    }
}

// Glyphs are independent, so the batch entry points split [0, n) into contiguous ranges over host threads (the GPU needs
// ~1.2 us per glyph; preparation costs ~5 us per glyph of CPU time plus as much again in allocator work, which would make it
// the slowest stage of an atlas build).  At most 16 threads by default; B200MSDF_HOST_THREADS overrides the count.  Measured in
// the 8-vCPU build container on 4096 glyphs: prepare 22 -> 5 ms, flatten 3 -> 1 ms, teardown 8 -> 5 ms with a warm heap; with a
// cold heap (fresh output arrays every call) the page faults of the allocator dominate and the split gains nothing.
template <typename F>
void parallelRanges(size_t n, size_t grain, F body) {
    size_t threads = std::thread::hardware_concurrency();
    if (threads > 16)
        threads = 16;
    if (const char *env = std::getenv("B200MSDF_HOST_THREADS"))
        threads = (size_t) std::atoi(env);
    if (threads > n/grain)
        threads = n/grain;
    if (threads <= 1) {
        body(size_t(0), n);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(threads-1);
    for (size_t t = 1; t < threads; ++t)
        pool.emplace_back(body, n*t/threads, n*(t+1)/threads);
    body(size_t(0), n/threads);
    for (size_t t = 0; t < pool.size(); ++t)
        pool[t].join();
}

}  // namespace

extern "C" {

/// Pass 1 of 2: prepares every glyph and returns the totals needed to size the output arrays. The prepared shapes are kept
/// in an opaque handle until b200host_prepare_fetch / b200host_prepare_free.
struct b200host_prepared {
    std::vector<Shape> shapes;          // glyphs that went through the general path (empty Shape for the others)
    std::vector<uint8_t> flat;          // 1: the glyph's geometry is unchanged, only colours were computed (flatColors)
    std::vector<uint8_t> flatColors;    // per input segment
    // the caller's input arrays: flat glyphs are copied from them by b200host_prepare_fetch, so they must stay alive until then
    const uint32_t *gbegin, *cbegin;
    const double *points;
    const uint8_t *orders;
};

int b200host_prepare_begin(int nGlyphs, const uint32_t *gbegin, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors,
                           const uint8_t *yDown, int doNormalize, int coloring, double angleThreshold, unsigned long long seed,
                           b200host_prepared **handle, int *outContours, int *outSegments) {
    b200host_prepared *h = new b200host_prepared;
    h->shapes.resize((size_t) nGlyphs);
    h->flat.assign((size_t) nGlyphs, 0);
    h->flatColors.resize(nGlyphs > 0 ? (size_t) cbegin[gbegin[nGlyphs]] : 0);
    h->gbegin = gbegin; h->cbegin = cbegin; h->points = points; h->orders = orders;
    // B200MSDF_HOST_FLAT=0 forces every glyph through the Shape objects (for testing the two paths against each other)
    const char *flatEnv = std::getenv("B200MSDF_HOST_FLAT");
    const bool tryFlat = !(flatEnv && flatEnv[0] == '0');
    parallelRanges((size_t) nGlyphs, 64, [&](size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; ++g) {
            // almost every glyph keeps its geometry: colours are then computed straight from the flat arrays, no allocation
            if (tryFlat && b200::prepareGlyphFlat(cbegin, gbegin[g], gbegin[g+1], points, orders, colors, doNormalize != 0, coloring, angleThreshold, seed, h->flatColors.data())) {
                h->flat[g] = 1;
                continue;
            }
            Shape &shape = h->shapes[g];
            buildShape(shape, cbegin, gbegin[g], gbegin[g+1], points, orders, colors, yDown ? yDown[g] != 0 : false);
            if (doNormalize)
                shape.normalize();
            if (coloring == 1)
                edgeColoringSimple(shape, angleThreshold, seed);
            else if (coloring == 2)
                edgeColoringInkTrap(shape, angleThreshold, seed);
        }
    });
    int nc = 0, ns = 0;
    for (int g = 0; g < nGlyphs; ++g) {
        if (h->flat[(size_t) g]) {
            nc += (int) (gbegin[g+1]-gbegin[g]);
            ns += (int) (cbegin[gbegin[g+1]]-cbegin[gbegin[g]]);
        } else {
            nc += (int) h->shapes[(size_t) g].contours.size();
            ns += h->shapes[(size_t) g].edgeCount();
        }
    }
    *handle = h;
    *outContours = nc;
    *outSegments = ns;
    return 0;
}

int b200host_prepare_fetch(b200host_prepared *h, uint32_t *gbegin, uint32_t *cbegin, double *points, uint8_t *orders, uint8_t *colors) {
    // offsets first (serial, O(glyphs)), then every thread flattens its own glyph range
    const size_t n = h->shapes.size();
    std::vector<uint32_t> segBegin(n+1, 0);
    gbegin[0] = 0;
    cbegin[0] = 0;
    for (size_t g = 0; g < n; ++g) {
        if (h->flat[g]) {
            gbegin[g+1] = gbegin[g]+(h->gbegin[g+1]-h->gbegin[g]);
            segBegin[g+1] = segBegin[g]+(h->cbegin[h->gbegin[g+1]]-h->cbegin[h->gbegin[g]]);
        } else {
            gbegin[g+1] = gbegin[g]+(uint32_t) h->shapes[g].contours.size();
            segBegin[g+1] = segBegin[g]+(uint32_t) h->shapes[g].edgeCount();
        }
    }
    parallelRanges(n, 64, [&](size_t g0, size_t g1) {
    for (size_t g = g0; g < g1; ++g) {
        uint32_t c = gbegin[g], s = segBegin[g];
        if (h->flat[g]) {
            // unchanged geometry: block copy of the input segments (unused control-point slots zeroed like the general path)
            const uint32_t ic0 = h->gbegin[g], ic1 = h->gbegin[g+1];
            const uint32_t is0 = h->cbegin[ic0], is1 = h->cbegin[ic1];
            for (uint32_t ic = ic0; ic < ic1; ++ic)
                cbegin[++c] = s+(h->cbegin[ic+1]-is0);
            for (uint32_t is = is0; is < is1; ++is, ++s) {
                const int order = h->orders[is];
                for (int k = 0; k < 8; ++k)
                    points[8*size_t(s)+k] = k < 2*(order+1) ? h->points[8*size_t(is)+k] : 0;
                orders[s] = h->orders[is];
                colors[s] = h->flatColors[is];
            }
            continue;
        }
        const Shape &shape = h->shapes[g];
        for (size_t ci = 0; ci < shape.contours.size(); ++ci) {
            const std::vector<EdgeHolder> &edges = shape.contours[ci].edges;
            for (size_t i = 0; i < edges.size(); ++i) {
                const EdgeSegment *edge = edges[i];
                int order = edge->type();
                const Point2 *cp = edge->controlPoints();
                for (int k = 0; k < 4; ++k) {
                    points[8*size_t(s)+2*k] = k <= order ? cp[k].x : 0;
                    points[8*size_t(s)+2*k+1] = k <= order ? cp[k].y : 0;
                }
                orders[s] = (uint8_t) order;
                colors[s] = (uint8_t) edge->color;
                ++s;
            }
            cbegin[++c] = s;
        }
    }
    });
    return 0;
}

void b200host_prepare_free(b200host_prepared *h) {
    if (!h)
        return;
    // tearing down ~36 heap-allocated segments per glyph is as expensive as building them: same split over threads
    parallelRanges(h->shapes.size(), 64, [&](size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; ++g)
            std::vector<Contour>().swap(h->shapes[g].contours);
    });
    delete h;
}

/// Contour::winding() of every contour; Shape::validate() and Shape::getBounds() per glyph.
int b200host_windings(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int *out) {
    Shape shape;
    buildShape(shape, cbegin, 0, (uint32_t) nContours, points, orders, NULL, false);
    for (int c = 0; c < nContours; ++c)
        out[c] = shape.contours[(size_t) c].winding();
    return 0;
}

int b200host_validate(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders) {
    Shape shape;
    buildShape(shape, cbegin, 0, (uint32_t) nContours, points, orders, NULL, false);
    return shape.validate() ? 1 : 0;
}

int b200host_bounds(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, double border, double miterLimit, int polarity, double *lbrt) {
    Shape shape;
    buildShape(shape, cbegin, 0, (uint32_t) nContours, points, orders, NULL, false);
    Shape::Bounds b = shape.getBounds(border, miterLimit, polarity);
    lbrt[0] = b.l; lbrt[1] = b.b; lbrt[2] = b.r; lbrt[3] = b.t;
    return 0;
}

}
