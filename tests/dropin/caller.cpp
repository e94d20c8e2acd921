// caller.cpp — a msdfgen USER program, written only against the public msdfgen.h interface (README usage pattern,
// /root/reference/README.md:114-142).  tests/test_dropin.py compiles this SAME file twice:
//   * against the unmodified reference headers + core sources (-I/root/reference)          -> libcaller_ref.so
//   * against msdfgen_b200/host/msdfgen.h + libmsdfgen_b200_host.so (the B200 backend)     -> libcaller_b200.so
// and compares the outputs.  That it compiles unchanged against both is the drop-in claim; nothing here is B200-specific.
#include "msdfgen.h"

#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using namespace msdfgen;

static void buildShape(Shape &shape, int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int yDown) {
    for (int c = 0; c < nContours; ++c) {
        Contour &contour = shape.addContour();
        for (uint32_t s = cbegin[c]; s < cbegin[c+1]; ++s) {
            const double *p = pts+8*s;
            if (orders[s] == 1)
                contour.addEdge(EdgeHolder(Point2(p[0], p[1]), Point2(p[2], p[3])));
            else if (orders[s] == 2)
                contour.addEdge(EdgeHolder(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5])));
            else
                contour.addEdge(EdgeHolder(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), Point2(p[6], p[7])));
        }
    }
    if (yDown)
        shape.setYAxisOrientation(Y_DOWNWARD);
}

static SDFTransformation autoframe(const Shape &shape, int w, int h, double pxRange) {
    Shape::Bounds b = shape.getBounds();
    double sx = w/(b.r-b.l)*.9, sy = h/(b.t-b.b)*.9;
    double s = min(sx, sy);
    Vector2 translate(-b.l+.5*(w/s-(b.r-b.l)), -b.b+.5*(h/s-(b.t-b.b)));
    return SDFTransformation(Projection(Vector2(s), translate), Range(pxRange/s));
}

extern "C" {

// README flow: normalize, colour, generateMSDF with the default config into an owning Bitmap.
int caller_msdf(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int yDown, int w, int h, float *out) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, yDown);
    if (!shape.validate())
        return -1;
    shape.normalize();
    edgeColoringSimple(shape, 3.0);
    Bitmap<float, 3> msdf(w, h);
    generateMSDF(msdf, shape, autoframe(shape, w, h, 4.0));
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            memcpy(out+3*(size_t(w)*y+x), msdf(x, y), 3*sizeof(float));
    return shape.edgeCount();
}

// Four generators into sections of one atlas-like buffer (row strides, Y_DOWNWARD section), plus the legacy overloads.
// out: float[4][h][w][4] (unused channels stay at the fill value).
int caller_all_modes(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int w, int h, int overlap, float *out) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, 0);
    shape.normalize();
    edgeColoringInkTrap(shape, 3.0, 3);
    SDFTransformation t = autoframe(shape, w, h, 3.0);
    Projection proj(t.projectVector(Vector2(1)), -t.unproject(Point2(0)));
    Range range(3.0/t.projectVector(Vector2(1)).x);
    size_t plane = size_t(w)*h*4;
    // 1-channel results are written with a row stride of 4*w floats (every 4th float of the plane is a texel row start)
    generateSDF(BitmapSection<float, 1>(out, w, h, 4*w), shape, t, GeneratorConfig(overlap != 0));
    generatePSDF(BitmapSection<float, 1>(out+plane, w, h, 4*w, Y_DOWNWARD), shape, proj, range, GeneratorConfig(overlap != 0));
    generateMSDF(BitmapSection<float, 3>(out+2*plane, w, h, 4*w), shape, range, t.projectVector(Vector2(1)), -t.unproject(Point2(0)), ErrorCorrectionConfig(), overlap != 0);
    Bitmap<float, 4> mtsdf(w, h);
    generateMTSDF(mtsdf, shape, t, MSDFGeneratorConfig(overlap != 0, ErrorCorrectionConfig(ErrorCorrectionConfig::EDGE_ONLY, ErrorCorrectionConfig::ALWAYS_CHECK_DISTANCE)));
    memcpy(out+3*plane, (const float *) mtsdf, plane*sizeof(float));
    return 0;
}

// Distance pass without correction, then the stand-alone correction entry points.
int caller_error_correction(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int w, int h, float *out, uint8_t *stencil) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, 0);
    shape.normalize();
    edgeColoringSimple(shape, 3.0, 5);
    SDFTransformation t = autoframe(shape, w, h, 4.0);
    size_t plane = size_t(w)*h*3;
    MSDFGeneratorConfig off(true, ErrorCorrectionConfig(ErrorCorrectionConfig::DISABLED));
    for (int i = 0; i < 3; ++i)
        generateMSDF(BitmapSection<float, 3>(out+i*plane, w, h), shape, t, off);
    MSDFGeneratorConfig withBuffer(true, ErrorCorrectionConfig(ErrorCorrectionConfig::EDGE_PRIORITY, ErrorCorrectionConfig::CHECK_DISTANCE_AT_EDGE,
        ErrorCorrectionConfig::defaultMinDeviationRatio, ErrorCorrectionConfig::defaultMinImproveRatio, stencil));
    msdfErrorCorrection(BitmapSection<float, 3>(out, w, h), shape, t, withBuffer);
    msdfFastDistanceErrorCorrection(BitmapSection<float, 3>(out+plane, w, h), t);
    msdfFastEdgeErrorCorrection(BitmapSection<float, 3>(out+2*plane, w, h), t, 1.2);
    return 0;
}

// The command line's scanline flow (no geometry preprocessing: contours are NOT oriented, signs come from the fill):
// distance pass with correction postponed, distanceSignCorrection, msdfErrorCorrection; plus rasterize and the legacy overloads.
// out: float[h][w][4] mtsdf | float[h][w][3] msdf (legacy scale/translate overload, FILL_ODD) | float[h][w] sdf | float[h][w] raster
int caller_scanline(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int w, int h, float *out) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, 0);
    // reverse every other contour: a shape as an un-preprocessed font outline may deliver it
    for (size_t i = 1; i < shape.contours.size(); i += 2)
        shape.contours[i].reverse();
    edgeColoringSimple(shape, 3.0);
    SDFTransformation t = autoframe(shape, w, h, 4.0);
    MSDFGeneratorConfig off(true, ErrorCorrectionConfig(ErrorCorrectionConfig::DISABLED));
    float *mtsdf = out, *msdf = out+size_t(w)*h*4, *sdf = msdf+size_t(w)*h*3, *raster = sdf+size_t(w)*h;
    generateMTSDF(BitmapSection<float, 4>(mtsdf, w, h), shape, t, off);
    distanceSignCorrection(BitmapSection<float, 4>(mtsdf, w, h), shape, t);
    msdfErrorCorrection(BitmapSection<float, 4>(mtsdf, w, h), shape, t);
    generateMSDF(BitmapSection<float, 3>(msdf, w, h), shape, t, off);
    distanceSignCorrection(BitmapSection<float, 3>(msdf, w, h), shape, t.projectVector(Vector2(1)), -t.unproject(Point2(0)), FILL_ODD);
    generateSDF(BitmapSection<float, 1>(sdf, w, h, w, Y_DOWNWARD), shape, t);
    distanceSignCorrection(BitmapSection<float, 1>(sdf, w, h, w, Y_DOWNWARD), shape, t, .5f, FILL_POSITIVE);
    rasterize(BitmapSection<float, 1>(raster, w, h), shape, t, FILL_NONZERO);
    return 0;
}

// Preview path of the command line (-testrender, main.cpp:1330-1390): generate, snap to 8 bits, render at another size.
// out: float[2h][2w][3] render of the msdf | float[h][w] hard-threshold render of the mtsdf's median
int caller_render(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int w, int h, float *out) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, 0);
    shape.normalize();
    edgeColoringSimple(shape, 3.0);
    SDFTransformation t = autoframe(shape, w, h, 4.0);
    Bitmap<float, 3> msdf(w, h);
    generateMSDF(msdf, shape, t);
    simulate8bit(msdf);
    renderSDF(BitmapSection<float, 3>(out, 2*w, 2*h), msdf, Range(4.0));
    Bitmap<float, 4> mtsdf(w, h);
    generateMTSDF(mtsdf, shape, t);
    renderSDF(BitmapSection<float, 1>(out+size_t(2*w)*(2*h)*3, w, h), mtsdf);
    return 0;
}


// The command line's error estimate (main.cpp:1323-1379): generate, then estimateSDFError of the result.  out[0..2] = the estimates
// of the sdf, msdf and mtsdf of the shape.
int caller_estimate(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int w, int h, double *out) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, 0);
    shape.normalize();
    edgeColoringSimple(shape, 3.0);
    SDFTransformation t = autoframe(shape, w, h, 4.0);
    Bitmap<float, 1> sdf(w, h);
    generateSDF(sdf, shape, t);
    out[0] = estimateSDFError(sdf, shape, t, 4, FILL_NONZERO);
    Bitmap<float, 3> msdf(w, h);
    generateMSDF(msdf, shape, t);
    out[1] = estimateSDFError(msdf, shape, t, 10, FILL_NONZERO);
    Bitmap<float, 4> mtsdf(w, h);
    generateMTSDF(mtsdf, shape, t);
    out[2] = estimateSDFError(mtsdf, shape, t, 1, FILL_ODD);
    return 0;
}

// The shape data model as a user sees it (no generator involved, host only): every member function of the edge, contour and
// shape objects that returns geometry, probed on the caller's shape.  out receives a stream of doubles (cap of them at most);
// returns how many were written, or -1 if cap is too small.  `raw` builds the edges without the factory's order lowering.
namespace {
struct Stream {
    double *out; int cap, n;
    void put(double v) { if (n < cap) out[n] = v; ++n; }
    void put(Vector2 v) { put(v.x); put(v.y); }
    void put(const EdgeSegment *e) {
        put((double) e->type());
        put((double) (int) e->color);
        for (int i = 0; i <= e->type(); ++i)
            put(e->controlPoints()[i]);
    }
};
}

int caller_shape_model(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int raw, double *out, int cap) {
    static const double params[] = { 0, .25, 1/3., .5, .8125, 1 };
    Stream o = { out, cap, 0 };
    Shape shape;
    if (raw) {
        for (int c = 0; c < nContours; ++c) {
            Contour &contour = shape.addContour();
            for (uint32_t s = cbegin[c]; s < cbegin[c+1]; ++s) {
                const double *p = pts+8*s;
                if (orders[s] == 1)
                    contour.addEdge(EdgeHolder(new LinearSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), YELLOW)));
                else if (orders[s] == 2)
                    contour.addEdge(EdgeHolder(new QuadraticSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), CYAN)));
                else
                    contour.addEdge(EdgeHolder(new CubicSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), Point2(p[6], p[7]), MAGENTA)));
            }
        }
    } else
        buildShape(shape, nContours, cbegin, pts, orders, 0);
    o.put((double) shape.validate());
    o.put((double) shape.edgeCount());
    for (size_t c = 0; c < shape.contours.size(); ++c) {
        const Contour &contour = shape.contours[c];
        o.put((double) contour.winding());
        for (size_t i = 0; i < contour.edges.size(); ++i) {
            const EdgeSegment *edge = contour.edges[i];
            o.put(edge);
            for (size_t k = 0; k < sizeof(params)/sizeof(*params); ++k) {
                o.put(edge->point(params[k]));
                o.put(edge->direction(params[k]));
                o.put(edge->directionChange(params[k]));
            }
            double box[4] = { 1e9, 1e9, -1e9, -1e9 };
            edge->bound(box[0], box[1], box[2], box[3]);
            for (int k = 0; k < 4; ++k)
                o.put(box[k]);
            if (edge->type() == (int) LinearSegment::EDGE_TYPE)
                o.put(static_cast<const LinearSegment *>(edge)->length());
            if (edge->type() == (int) QuadraticSegment::EDGE_TYPE) {
                o.put(static_cast<const QuadraticSegment *>(edge)->length());
                EdgeHolder cubic(static_cast<const QuadraticSegment *>(edge)->convertToCubic());
                o.put(cubic);
            }
            EdgeSegment *parts[3] = { };
            edge->splitInThirds(parts[0], parts[1], parts[2]);
            for (int k = 0; k < 3; ++k) {
                o.put(parts[k]);
                delete parts[k];
            }
            // edits on copies (EdgeHolder's copy is deep)
            const Point2 target = edge->point(.5)+Vector2(.375, -.25);
            EdgeHolder a(contour.edges[i]), b(contour.edges[i]), r = contour.edges[i];
            a->moveStartPoint(target);
            b->moveEndPoint(target);
            r->reverse();
            o.put(a); o.put(b); o.put(r);
            o.put(edge);    // the original is untouched
        }
        for (int polarity = -1; polarity <= 1; ++polarity) {
            double box[4] = { 1e9, 1e9, -1e9, -1e9 };
            contour.bound(box[0], box[1], box[2], box[3]);
            contour.boundMiters(box[0], box[1], box[2], box[3], .5, 2.5, polarity);
            for (int k = 0; k < 4; ++k)
                o.put(box[k]);
        }
        Contour reversed(contour);
        reversed.reverse();
        o.put((double) reversed.winding());
        for (size_t i = 0; i < reversed.edges.size(); ++i)
            o.put(reversed.edges[i]);
    }
    if (shape.edgeCount()) {
        const Shape::Bounds plain = shape.getBounds(), wide = shape.getBounds(.25), mitered = shape.getBounds(.25, 3, 1);
        const Shape::Bounds *all[3] = { &plain, &wide, &mitered };
        for (int k = 0; k < 3; ++k) {
            o.put(all[k]->l); o.put(all[k]->b); o.put(all[k]->r); o.put(all[k]->t);
        }
    }
    // preparation through the objects, both colouring strategies
    for (int strategy = 1; strategy <= 2; ++strategy) {
        Shape prepared(shape);
        prepared.normalize();
        if (strategy == 1)
            edgeColoringSimple(prepared, 3.0, 5);
        else
            edgeColoringInkTrap(prepared, 3.0, 5);
        o.put((double) prepared.edgeCount());
        for (size_t c = 0; c < prepared.contours.size(); ++c)
            for (size_t i = 0; i < prepared.contours[c].edges.size(); ++i)
                o.put(prepared.contours[c].edges[i]);
    }
    return o.n <= cap ? o.n : -1;
}


// Re-entrancy (the reference's generators are free functions without shared state, msdfgen.h:47-69): nThreads host threads
// generate concurrently, each its own Shape object and its own bitmap of its own size (w+thread, h+thread), `reps` times.
// out: the threads' [h+t][w+t][3] images back to back.  Returns the number of generator calls made.
int caller_threads(int nThreads, int reps, int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int w, int h, float *out) {
    std::vector<std::thread> threads;
    std::vector<size_t> offsets;
    size_t off = 0;
    for (int t = 0; t < nThreads; ++t) {
        offsets.push_back(off);
        off += size_t(w+t)*(h+t)*3;
    }
    std::vector<int> calls((size_t) nThreads, 0);
    for (int t = 0; t < nThreads; ++t)
        threads.push_back(std::thread([=, &calls]() {
            Shape shape;
            buildShape(shape, nContours, cbegin, pts, orders, 0);
            shape.normalize();
            edgeColoringSimple(shape, 3.0, (unsigned long long) t);
            const int tw = w+t, th = h+t;
            Bitmap<float, 3> msdf(tw, th);
            for (int r = 0; r < reps; ++r) {
                generateMSDF(msdf, shape, autoframe(shape, tw, th, 4.0));
                ++calls[(size_t) t];
            }
            memcpy(out+offsets[(size_t) t], (const float *) msdf, size_t(tw)*th*3*sizeof(float));
        }));
    int total = 0;
    for (int t = 0; t < nThreads; ++t) {
        threads[(size_t) t].join();
        total += calls[(size_t) t];
    }
    return total;
}

}
