// flat_c_api.cpp — extern "C" access to the host shape preparation for callers that keep shapes as flat arrays
// (the Python package, bench.py): Shape::normalize() and an edge colouring for every glyph of a flat batch, computed on the
// arrays themselves (flat_prep.h) — no msdfgen::Shape is built unless B200MSDF_HOST_FLAT=0 asks for the object path.
// Host-only (no device work); this is the batch-ingest step that precedes b200msdf_generate_batch (SURVEY.md §8f "next" #1).
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
// ~1.1 us per glyph; preparation costs ~2 us per glyph of CPU time, which alone would make it the slowest stage of an atlas
// build).  At most 16 threads by default; B200MSDF_HOST_THREADS overrides the count.  Measured in the 8-vCPU build container on
// 4096 glyphs (begin + fetch): 8.5 ms on one thread, 3.1 ms on eight.
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

/// Pass 1 of 2: prepares every glyph and returns the totals needed to size the output arrays.  Glyphs whose geometry changed
/// are kept (as flat segment lists) in an opaque handle until b200host_prepare_fetch / b200host_prepare_free.
struct b200host_prepared {
    struct Changed {
        std::vector<flat::Seg> segs;
        std::vector<uint32_t> cbegin;       // local contour offsets into segs, from 0
    };
    std::vector<Changed> changed;       // per glyph; empty cbegin = the glyph's geometry is unchanged, only colours were computed
    std::vector<uint8_t> flatColors;    // per input segment (unchanged glyphs)
    // the caller's input arrays: unchanged glyphs are copied from them by b200host_prepare_fetch, so they must stay alive until then
    const uint32_t *gbegin, *cbegin;
    const double *points;
    const uint8_t *orders;
};

int b200host_prepare_begin(int nGlyphs, const uint32_t *gbegin, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors,
                           const uint8_t *yDown, int doNormalize, int coloring, double angleThreshold, unsigned long long seed,
                           b200host_prepared **handle, int *outContours, int *outSegments) {
    b200host_prepared *h = new b200host_prepared;
    h->changed.resize((size_t) nGlyphs);
    h->flatColors.resize(nGlyphs > 0 ? (size_t) cbegin[gbegin[nGlyphs]] : 0);
    h->gbegin = gbegin; h->cbegin = cbegin; h->points = points; h->orders = orders;
    // B200MSDF_HOST_FLAT=0 sends every glyph through the msdfgen-compatible Shape objects instead (Shape::normalize,
    // edgeColoring*): the two entry paths share their algorithms but not their plumbing, and tests compare them
    const char *flatEnv = std::getenv("B200MSDF_HOST_FLAT");
    const bool viaObjects = flatEnv && flatEnv[0] == '0';
    parallelRanges((size_t) nGlyphs, 64, [&](size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; ++g) {
            b200host_prepared::Changed &out = h->changed[g];
            if (!viaObjects) {
                // almost every glyph keeps its geometry: colours are then computed straight from the arrays, no allocation
                flat::prepareGlyph(cbegin, gbegin[g], gbegin[g+1], points, orders, colors, doNormalize != 0, coloring, angleThreshold, seed,
                                   h->flatColors.data(), out.segs, out.cbegin);
                continue;
            }
            Shape shape;
            buildShape(shape, cbegin, gbegin[g], gbegin[g+1], points, orders, colors, yDown ? yDown[g] != 0 : false);
            if (doNormalize)
                shape.normalize();
            if (coloring == 1)
                edgeColoringSimple(shape, angleThreshold, seed);
            else if (coloring == 2)
                edgeColoringInkTrap(shape, angleThreshold, seed);
            out.cbegin.push_back(0);
            for (size_t ci = 0; ci < shape.contours.size(); ++ci) {
                const std::vector<EdgeHolder> &edges = shape.contours[ci].edges;
                for (size_t i = 0; i < edges.size(); ++i) {
                    const Point2 *cp = edges[i]->controlPoints();
                    const int order = edges[i]->type();
                    out.segs.push_back(flat::makeSeg(order, cp[0], cp[1], order >= 2 ? cp[2] : Point2(), order >= 3 ? cp[3] : Point2(), edges[i]->color));
                }
                out.cbegin.push_back((uint32_t) out.segs.size());
            }
        }
    });
    int nc = 0, ns = 0;
    for (int g = 0; g < nGlyphs; ++g) {
        const b200host_prepared::Changed &out = h->changed[(size_t) g];
        if (out.cbegin.empty()) {
            nc += (int) (gbegin[g+1]-gbegin[g]);
            ns += (int) (cbegin[gbegin[g+1]]-cbegin[gbegin[g]]);
        } else {
            nc += (int) out.cbegin.size()-1;
            ns += (int) out.segs.size();
        }
    }
    *handle = h;
    *outContours = nc;
    *outSegments = ns;
    return 0;
}

int b200host_prepare_fetch(b200host_prepared *h, uint32_t *gbegin, uint32_t *cbegin, double *points, uint8_t *orders, uint8_t *colors) {
    // offsets first (serial, O(glyphs)), then every thread writes its own glyph range
    const size_t n = h->changed.size();
    std::vector<uint32_t> segBegin(n+1, 0);
    gbegin[0] = 0;
    cbegin[0] = 0;
    for (size_t g = 0; g < n; ++g) {
        const b200host_prepared::Changed &out = h->changed[g];
        if (out.cbegin.empty()) {
            gbegin[g+1] = gbegin[g]+(h->gbegin[g+1]-h->gbegin[g]);
            segBegin[g+1] = segBegin[g]+(h->cbegin[h->gbegin[g+1]]-h->cbegin[h->gbegin[g]]);
        } else {
            gbegin[g+1] = gbegin[g]+(uint32_t) out.cbegin.size()-1;
            segBegin[g+1] = segBegin[g]+(uint32_t) out.segs.size();
        }
    }
    parallelRanges(n, 64, [&](size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; ++g) {
            const b200host_prepared::Changed &out = h->changed[g];
            uint32_t c = gbegin[g], s = segBegin[g];
            if (out.cbegin.empty()) {
                // unchanged geometry: the input segments with their new colours (unused control-point slots zeroed)
                const uint32_t ic0 = h->gbegin[g], ic1 = h->gbegin[g+1];
                const uint32_t is0 = h->cbegin[ic0], is1 = h->cbegin[ic1];
                for (uint32_t ic = ic0; ic < ic1; ++ic)
                    cbegin[++c] = s+(h->cbegin[ic+1]-is0);
                for (uint32_t is = is0; is < is1; ++is, ++s)
                    flat::storeSeg(flat::loadSeg(h->points, h->orders, h->flatColors.data(), is), points, orders, colors, s);
                continue;
            }
            for (size_t ci = 0; ci+1 < out.cbegin.size(); ++ci)
                cbegin[++c] = s+out.cbegin[ci+1];
            for (size_t i = 0; i < out.segs.size(); ++i)
                flat::storeSeg(out.segs[i], points, orders, colors, s+i);
        }
    });
    return 0;
}

void b200host_prepare_free(b200host_prepared *h) {
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
