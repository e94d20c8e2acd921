/*
 * oracle/ref_wrapper.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin extern "C" wrapper over the UNMODIFIED reference (Chlumsky/msdfgen, core/ only), compiled
 * straight from the sources where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libmsdfgen_ref.so (git-ignored).  Nothing here restates the algorithm; every function
 * builds a msdfgen::Shape from the flat exchange format and calls the reference's own public API:
 *   generateSDF/PSDF/MSDF/MTSDF            (/root/reference/msdfgen.h:47-56)
 *   msdfErrorCorrection / msdfFast*         (/root/reference/core/msdf-error-correction.h)
 *   Shape::normalize, edgeColoringSimple    (/root/reference/core/Shape.cpp:65, edge-coloring.cpp:68)
 *   Contour::winding                        (/root/reference/core/Contour.cpp:57)
 *   ShapeDistanceFinder::oneShotDistance    (/root/reference/core/ShapeDistanceFinder.hpp:36)
 *   readShapeDescription                    (/root/reference/core/shape-description.cpp:235)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Flat exchange format (identical to include/msdfgen_b200.h's b200msdf_shapes for one glyph):
 *   points  double[nSeg][8]   control points p0..p3 (x,y), unused slots ignored
 *   orders  uint8[nSeg]       EdgeSegment::type(): 1 linear, 2 quadratic, 3 cubic
 *   colors  uint8[nSeg]       EdgeColor bitmask
 *   cbegin  uint32[nContours+1]  segment index range of every contour (contour->edges order)
 *   yDown   int               Shape::inverseYAxis
 */
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "msdfgen.h"
#include "core/ShapeDistanceFinder.h"
#include "core/contour-combiners.h"
#include "core/edge-selectors.h"
#include "core/MSDFErrorCorrection.h"
#include "core/pixel-conversion.hpp"
#include "core/equation-solver.h"
#include "core/render-sdf.h"
#include "core/sdf-error-estimation.h"

using namespace msdfgen;

namespace {

void buildShape(Shape &shape, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int yDown) {
    shape.contours.clear();
    shape.inverseYAxis = yDown != 0;
    for (int c = 0; c < nContours; ++c) {
        Contour &contour = shape.addContour();
        for (uint32_t s = cbegin[c]; s < cbegin[c+1]; ++s) {
            const double *p = points+8*size_t(s);
            EdgeColor col = EdgeColor(colors ? colors[s] : 7);
            // Construct the concrete segment classes directly: EdgeSegment::create() may demote degenerate curves.
            switch (orders[s]) {
                case 1: contour.addEdge(EdgeHolder(new LinearSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), col))); break;
                case 2: contour.addEdge(EdgeHolder(new QuadraticSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), col))); break;
                default: contour.addEdge(EdgeHolder(new CubicSegment(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), Point2(p[6], p[7]), col))); break;
            }
        }
    }
}

// Returns number of segments, or -1 if capacity is insufficient.
int flattenShape(const Shape &shape, int capSeg, int capContours, uint32_t *cbegin, double *points, uint8_t *orders, uint8_t *colors) {
    if ((int) shape.contours.size() > capContours)
        return -1;
    int s = 0;
    cbegin[0] = 0;
    for (size_t c = 0; c < shape.contours.size(); ++c) {
        const Contour &contour = shape.contours[c];
        for (size_t e = 0; e < contour.edges.size(); ++e) {
            if (s >= capSeg)
                return -1;
            const EdgeSegment *edge = contour.edges[e];
            int n = edge->type()+1;
            const Point2 *cp = edge->controlPoints();
            for (int i = 0; i < 4; ++i) {
                points[8*size_t(s)+2*i] = i < n ? cp[i].x : 0;
                points[8*size_t(s)+2*i+1] = i < n ? cp[i].y : 0;
            }
            orders[s] = (uint8_t) edge->type();
            colors[s] = (uint8_t) edge->color;
            ++s;
        }
        cbegin[c+1] = (uint32_t) s;
    }
    return s;
}

SDFTransformation makeTransformation(const double *xf) {
    // xf = scale.x, scale.y, translate.x, translate.y, range.lower, range.upper
    return SDFTransformation(Projection(Vector2(xf[0], xf[1]), Vector2(xf[2], xf[3])), DistanceMapping(Range(xf[4], xf[5])));
}

MSDFGeneratorConfig makeConfig(int overlap, int ecMode, int ecDcm, double minDev, double minImp, uint8_t *ecBuffer) {
    ErrorCorrectionConfig ec((ErrorCorrectionConfig::Mode) ecMode, (ErrorCorrectionConfig::DistanceCheckMode) ecDcm, minDev, minImp, (byte *) ecBuffer);
    return MSDFGeneratorConfig(overlap != 0, ec);
}

int generateOne(int mode, const Shape &shape, const SDFTransformation &t, float *out, int w, int h, int rowStride, int outYDown, const MSDFGeneratorConfig &config) {
    YAxisOrientation yo = outYDown ? Y_DOWNWARD : Y_UPWARD;
    switch (mode) {
        case 1: generateSDF(BitmapSection<float, 1>(out, w, h, rowStride, yo), shape, t, GeneratorConfig(config.overlapSupport)); return 0;
        case 2: generatePSDF(BitmapSection<float, 1>(out, w, h, rowStride, yo), shape, t, GeneratorConfig(config.overlapSupport)); return 0;
        case 3: generateMSDF(BitmapSection<float, 3>(out, w, h, rowStride, yo), shape, t, config); return 0;
        case 4: generateMTSDF(BitmapSection<float, 4>(out, w, h, rowStride, yo), shape, t, config); return 0;
    }
    return -1;
}

}

extern "C" {

int ref_info(int *ompEnabled, int *maxThreads) {
    // ompEnabled: 2 = the reference's own row-parallel pragmas are compiled in (MSDFGEN_USE_OPENMP), 1 = only this wrapper's
    // outer loops are parallel (glyph-parallel harness), 0 = single-threaded build.  maxThreads = what an OpenMP region would use.
#if defined(MSDFGEN_USE_OPENMP)
    *ompEnabled = 2;
    *maxThreads = omp_get_max_threads();
#elif defined(_OPENMP)
    *ompEnabled = 1;
    *maxThreads = omp_get_max_threads();
#else
    *ompEnabled = 0;
    *maxThreads = 1;
#endif
    return MSDFGEN_VERSION_MAJOR*10000+MSDFGEN_VERSION_MINOR*100+MSDFGEN_VERSION_REVISION;
}

/// Sets the OpenMP team size of this build's parallel regions (the environment may have been read before the caller could set
/// OMP_NUM_THREADS — e.g. torchrun exports OMP_NUM_THREADS=1); returns what a region will now use.
int ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void) n;
    return 1;
#endif
}

/// generate{SDF,PSDF,MSDF,MTSDF}(BitmapSection, Shape, SDFTransformation, config) for one shape.
int ref_generate(int mode, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int yDown,
                 const double *xf, float *out, int w, int h, int rowStride, int outYDown,
                 int overlap, int ecMode, int ecDcm, double minDev, double minImp, uint8_t *ecBuffer) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, colors, yDown);
    return generateOne(mode, shape, makeTransformation(xf), out, w, h, rowStride, outYDown, makeConfig(overlap, ecMode, ecDcm, minDev, minImp, ecBuffer));
}

/// Loops the reference generator over a batch of glyphs (same w,h; packed outputs). Shape construction is outside the timed
/// region; returns seconds spent inside generate* via *seconds. glyphParallel != 0 adds an outer OpenMP loop over glyphs
/// (NOT the reference's own parallelisation; labelled as such by callers).
int ref_generate_batch(int mode, int nGlyphs, const uint32_t *gbegin, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors,
                       const uint8_t *yDown, const double *xf, int xfPerGlyph, float *out, int w, int h,
                       int overlap, int ecMode, int ecDcm, double minDev, double minImp, int glyphParallel, double *seconds) {
    int N = mode == 3 ? 3 : mode == 4 ? 4 : 1;
    std::vector<Shape> shapes((size_t) nGlyphs);
    for (int g = 0; g < nGlyphs; ++g) {
        uint32_t c0 = gbegin[g], c1 = gbegin[g+1];
        std::vector<uint32_t> local(c1-c0+1);
        uint32_t base = cbegin[c0];
        for (uint32_t c = c0; c <= c1; ++c)
            local[c-c0] = cbegin[c]-base;
        buildShape(shapes[g], int(c1-c0), local.data(), points+8*size_t(base), orders+base, colors ? colors+base : NULL, yDown ? yDown[g] : 0);
    }
    MSDFGeneratorConfig config = makeConfig(overlap, ecMode, ecDcm, minDev, minImp, NULL);
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    int rc = 0;
    if (glyphParallel) {
#ifdef _OPENMP
        #pragma omp parallel for schedule(dynamic, 4)
#endif
        for (int g = 0; g < nGlyphs; ++g)
            generateOne(mode, shapes[g], makeTransformation(xf+(xfPerGlyph ? 6*size_t(g) : 0)), out+size_t(g)*w*h*N, w, h, w*N, 0, config);
    } else {
        for (int g = 0; g < nGlyphs; ++g)
            rc |= generateOne(mode, shapes[g], makeTransformation(xf+(xfPerGlyph ? 6*size_t(g) : 0)), out+size_t(g)*w*h*N, w, h, w*N, 0, config);
    }
    std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    if (seconds)
        *seconds = std::chrono::duration<double>(t1-t0).count();
    return rc;
}

/// msdfErrorCorrection(sdf, shape, transformation, config) on an existing N-channel field (N = 3 or 4).
int ref_error_correction(int N, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int yDown,
                         const double *xf, float *sdf, int w, int h, int rowStride, int sdfYDown,
                         int overlap, int ecMode, int ecDcm, double minDev, double minImp, uint8_t *ecBuffer) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, colors, yDown);
    MSDFGeneratorConfig config = makeConfig(overlap, ecMode, ecDcm, minDev, minImp, ecBuffer);
    YAxisOrientation yo = sdfYDown ? Y_DOWNWARD : Y_UPWARD;
    if (N == 3)
        msdfErrorCorrection(BitmapSection<float, 3>(sdf, w, h, rowStride, yo), shape, makeTransformation(xf), config);
    else if (N == 4)
        msdfErrorCorrection(BitmapSection<float, 4>(sdf, w, h, rowStride, yo), shape, makeTransformation(xf), config);
    else
        return -1;
    return 0;
}

/// msdfFastDistanceErrorCorrection (edge == 0) / msdfFastEdgeErrorCorrection (edge != 0): shapeless variants.
int ref_fast_error_correction(int N, int edge, const double *xf, float *sdf, int w, int h, int rowStride, double minDev) {
    SDFTransformation t = makeTransformation(xf);
    if (N == 3) {
        BitmapSection<float, 3> s(sdf, w, h, rowStride);
        if (edge) msdfFastEdgeErrorCorrection(s, t, minDev); else msdfFastDistanceErrorCorrection(s, t, minDev);
    } else if (N == 4) {
        BitmapSection<float, 4> s(sdf, w, h, rowStride);
        if (edge) msdfFastEdgeErrorCorrection(s, t, minDev); else msdfFastDistanceErrorCorrection(s, t, minDev);
    } else
        return -1;
    return 0;
}

/// Shape::normalize() and/or edgeColoringSimple() by the reference's own code; writes the processed shape back in flat form.
int ref_prepare(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int yDown,
                int doNormalize, int coloring, double angleThreshold, unsigned long long seed,
                int capSeg, int capContours, uint32_t *outCbegin, double *outPoints, uint8_t *outOrders, uint8_t *outColors, int *outContours) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, colors, yDown);
    if (doNormalize)
        shape.normalize();
    switch (coloring) {
        case 1: edgeColoringSimple(shape, angleThreshold, seed); break;
        case 2: edgeColoringInkTrap(shape, angleThreshold, seed); break;
        case 3: edgeColoringByDistance(shape, angleThreshold, seed); break;
    }
    *outContours = (int) shape.contours.size();
    return flattenShape(shape, capSeg, capContours, outCbegin, outPoints, outOrders, outColors);
}

int ref_validate(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, NULL, 0);
    return shape.validate() ? 1 : 0;
}

int ref_windings(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int *out) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, NULL, 0);
    for (int c = 0; c < nContours; ++c)
        out[c] = shape.contours[c].winding();
    return 0;
}

/// ShapeDistanceFinder<Combiner<Selector>>::oneShotDistance at arbitrary points; raw (unmapped) fp64 distances, N per point.
int ref_oneshot_distance(int mode, int overlap, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors,
                         int nPts, const double *pts, double *out) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, colors, 0);
    if (mode < 1 || mode > 4)
        return -1;
    // (independent query points: the glyph-parallel harness build spreads them over the host threads)
#ifdef _OPENMP
    #pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int i = 0; i < nPts; ++i) {
        Point2 p(pts[2*i], pts[2*i+1]);
        switch (mode) {
            case 1:
                out[i] = overlap ? ShapeDistanceFinder<OverlappingContourCombiner<TrueDistanceSelector> >::oneShotDistance(shape, p)
                                 : ShapeDistanceFinder<SimpleContourCombiner<TrueDistanceSelector> >::oneShotDistance(shape, p);
                break;
            case 2:
                out[i] = overlap ? ShapeDistanceFinder<OverlappingContourCombiner<PerpendicularDistanceSelector> >::oneShotDistance(shape, p)
                                 : ShapeDistanceFinder<SimpleContourCombiner<PerpendicularDistanceSelector> >::oneShotDistance(shape, p);
                break;
            case 3: {
                MultiDistance d = overlap ? ShapeDistanceFinder<OverlappingContourCombiner<MultiDistanceSelector> >::oneShotDistance(shape, p)
                                          : ShapeDistanceFinder<SimpleContourCombiner<MultiDistanceSelector> >::oneShotDistance(shape, p);
                out[3*i] = d.r, out[3*i+1] = d.g, out[3*i+2] = d.b;
                break;
            }
            case 4: {
                MultiAndTrueDistance d = overlap ? ShapeDistanceFinder<OverlappingContourCombiner<MultiAndTrueDistanceSelector> >::oneShotDistance(shape, p)
                                                 : ShapeDistanceFinder<SimpleContourCombiner<MultiAndTrueDistanceSelector> >::oneShotDistance(shape, p);
                out[4*i] = d.r, out[4*i+1] = d.g, out[4*i+2] = d.b, out[4*i+3] = d.a;
                break;
            }
        }
    }
    return 0;
}

/// EdgeSegment::signedDistance for every segment of the flat list at every point: out[(i*nSeg+s)*3 + {0,1,2}] = distance, dot, param.
int ref_segment_distances(int nSeg, const double *points, const uint8_t *orders, int nPts, const double *pts, double *out) {
    uint32_t cb[2] = { 0, (uint32_t) nSeg };
    Shape shape;
    buildShape(shape, 1, cb, points, orders, NULL, 0);
    for (int i = 0; i < nPts; ++i)
        for (int s = 0; s < nSeg; ++s) {
            double param = 0;
            SignedDistance sd = shape.contours[0].edges[s]->signedDistance(Point2(pts[2*i], pts[2*i+1]), param);
            double *o = out+3*(size_t(i)*nSeg+s);
            o[0] = sd.distance, o[1] = sd.dot, o[2] = param;
        }
    return 0;
}

int ref_solve_cubic(double *x, double a, double b, double c, double d) {
    return solveCubic(x, a, b, c, d);
}

int ref_solve_quadratic(double *x, double a, double b, double c) {
    return solveQuadratic(x, a, b, c);
}

/// readShapeDescription(const char *) -> flat form. Returns segment count, -1 on capacity, -2 on parse error.
int ref_read_shape_description(const char *text, int capSeg, int capContours, uint32_t *outCbegin, double *outPoints, uint8_t *outOrders, uint8_t *outColors,
                               int *outContours, int *outYDown, int *colorsSpecified) {
    Shape shape;
    bool cs = false;
    if (!readShapeDescription(text, shape, &cs))
        return -2;
    *outContours = (int) shape.contours.size();
    *outYDown = shape.inverseYAxis ? 1 : 0;
    *colorsSpecified = cs ? 1 : 0;
    return flattenShape(shape, capSeg, capContours, outCbegin, outPoints, outOrders, outColors);
}

/// distanceSignCorrection(sdf, shape, projection, sdfZeroValue, fillRule) (core/rasterization.h), N = 1, 3 or 4; in place.
int ref_sign_correction(int N, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, int yDown,
                        const double *xf, float *sdf, int w, int h, int rowStride, int sdfYDown, float zeroValue, int fillRule) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, colors, yDown);
    Projection projection(Vector2(xf[0], xf[1]), Vector2(xf[2], xf[3]));
    YAxisOrientation yo = sdfYDown ? Y_DOWNWARD : Y_UPWARD;
    switch (N) {
        case 1: distanceSignCorrection(BitmapSection<float, 1>(sdf, w, h, rowStride, yo), shape, projection, zeroValue, (FillRule) fillRule); return 0;
        case 3: distanceSignCorrection(BitmapSection<float, 3>(sdf, w, h, rowStride, yo), shape, projection, zeroValue, (FillRule) fillRule); return 0;
        case 4: distanceSignCorrection(BitmapSection<float, 4>(sdf, w, h, rowStride, yo), shape, projection, zeroValue, (FillRule) fillRule); return 0;
    }
    return -1;
}

/// rasterize(output, shape, projection, fillRule): 1.0 / 0.0 coverage at texel centres.
int ref_rasterize(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int yDown,
                  const double *xf, float *out, int w, int h, int outYDown, int fillRule) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, NULL, yDown);
    rasterize(BitmapSection<float, 1>(out, w, h, w, outYDown ? Y_DOWNWARD : Y_UPWARD), shape, Projection(Vector2(xf[0], xf[1]), Vector2(xf[2], xf[3])), (FillRule) fillRule);
    return 0;
}

/// renderSDF(output, sdf, sdfPxRange, sdThreshold) (core/render-sdf.h:11-16) for the six channel combinations it declares.
int ref_render_sdf(int outN, float *out, int ow, int oh, int sdfN, const float *sdf, int sw, int sh, double lower, double upper, float threshold) {
    Range range(lower, upper);
    #define RENDER(NO, NI) if (outN == NO && sdfN == NI) { renderSDF(BitmapSection<float, NO>(out, ow, oh), BitmapConstSection<float, NI>(sdf, sw, sh), range, threshold); return 0; }
    RENDER(1, 1) RENDER(3, 1) RENDER(1, 3) RENDER(3, 3) RENDER(1, 4) RENDER(4, 4)
    #undef RENDER
    return -1;
}

/// simulate8bit (core/render-sdf.h:19-21), in place on n floats.
/// estimateSDFError(sdf, shape, projection, scanlinesPerRow, fillRule)  (core/sdf-error-estimation.h:19-21); xf: scale, translate
double ref_estimate_sdf_error(int N, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int yDown,
                              const double *xf, const float *sdf, int w, int h, int scanlinesPerRow, int fillRule) {
    Shape shape;
    buildShape(shape, nContours, cbegin, points, orders, NULL, yDown);
    Projection projection(Vector2(xf[0], xf[1]), Vector2(xf[2], xf[3]));
    switch (N) {
        case 1: return estimateSDFError(BitmapConstSection<float, 1>(sdf, w, h), shape, projection, scanlinesPerRow, (FillRule) fillRule);
        case 3: return estimateSDFError(BitmapConstSection<float, 3>(sdf, w, h), shape, projection, scanlinesPerRow, (FillRule) fillRule);
        case 4: return estimateSDFError(BitmapConstSection<float, 4>(sdf, w, h), shape, projection, scanlinesPerRow, (FillRule) fillRule);
    }
    return -1;
}

void ref_simulate_8bit(float *pixels, int w, int h, int N) {
    if (N == 1) simulate8bit(BitmapSection<float, 1>(pixels, w, h));
    else if (N == 3) simulate8bit(BitmapSection<float, 3>(pixels, w, h));
    else if (N == 4) simulate8bit(BitmapSection<float, 4>(pixels, w, h));
}

void ref_pixel_float_to_byte(const float *in, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; ++i)
        out[i] = pixelFloatToByte(in[i]);
}

}
