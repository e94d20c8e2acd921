// host_caller.cpp — a user of msdfgen's HOST-side scanline interface (Scanline, EdgeSegment::scanlineIntersections,
// Shape::scanline, Shape::orientContours, scanlineSDF, estimateSDFError, the equation solvers), written only against the public
// headers.  tests/test_host_scanline.py compiles this SAME file against the unmodified reference and against
// msdfgen_b200/host/msdfgen.h and compares the results.  No device work: runs on any box.
#include "msdfgen.h"
#if defined(__has_include)
#if __has_include("core/equation-solver.h")
#include "core/equation-solver.h"      // the reference keeps the solvers in their own header; the mirror declares them in msdfgen.h
#endif
#endif

#include <cstdint>
#include <cstring>
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

template <int N>
static void sdfScanline(const float *sdf, int w, int h, const double *xf, double y, int yDown, int nq, const double *xs, int *sum) {
    Scanline line;
    scanlineSDF(line, BitmapConstSection<float, N>(sdf, w, h), Projection(Vector2(xf[0], xf[1]), Vector2(xf[2], xf[3])), y, yDown ? Y_DOWNWARD : Y_UPWARD);
    for (int i = 0; i < nq; ++i)
        sum[i] = line.sumIntersections(xs[i]);
}

extern "C" {

int hc_solve_quadratic(double *x, double a, double b, double c) { return solveQuadratic(x, a, b, c); }
int hc_solve_cubic(double *x, double a, double b, double c, double d) { return solveCubic(x, a, b, c, d); }

// EdgeSegment::scanlineIntersections of one segment at several y
void hc_intersections(int order, const double *p, int ny, const double *ys, int *counts, double *xs, int *dys) {
    EdgeHolder edge = order == 1 ? EdgeHolder(Point2(p[0], p[1]), Point2(p[2], p[3]))
                    : order == 2 ? EdgeHolder(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]))
                                 : EdgeHolder(Point2(p[0], p[1]), Point2(p[2], p[3]), Point2(p[4], p[5]), Point2(p[6], p[7]));
    for (int i = 0; i < ny; ++i) {
        double x[3] = { 0, 0, 0 };
        int dy[3] = { 0, 0, 0 };
        counts[i] = edge->scanlineIntersections(x, dy, ys[i]);
        for (int k = 0; k < 3; ++k) {
            xs[3*i+k] = k < counts[i] ? x[k] : 0;
            dys[3*i+k] = k < counts[i] ? dy[k] : 0;
        }
    }
}

// Shape::scanline at y, queried at xs
void hc_scanline(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, double y, int nq, const double *xs, int fillRule,
                 int *count, int *sum, int *filled) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, 0);
    Scanline line;
    shape.scanline(line, y);
    for (int i = 0; i < nq; ++i) {
        count[i] = line.countIntersections(xs[i]);
        sum[i] = line.sumIntersections(xs[i]);
        filled[i] = line.filled(xs[i], FillRule(fillRule));
    }
}

// Scanline::overlap of two explicit intersection lists
double hc_overlap(int na, const double *ax, const int *ad, int nb, const double *bx, const int *bd, double xFrom, double xTo, int fillRule) {
    std::vector<Scanline::Intersection> a((size_t) na), b((size_t) nb);
    for (int i = 0; i < na; ++i) { a[(size_t) i].x = ax[i]; a[(size_t) i].direction = ad[i]; }
    for (int i = 0; i < nb; ++i) { b[(size_t) i].x = bx[i]; b[(size_t) i].direction = bd[i]; }
    Scanline sa, sb;
    sa.setIntersections(a);
    sb.setIntersections(b);
    return Scanline::overlap(sa, sb, xFrom, xTo, FillRule(fillRule));
}

// Shape::orientContours; the (possibly reversed) control points and orders come back in the input layout
void hc_orient(int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, double *outPts, uint8_t *outOrders) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, 0);
    shape.orientContours();
    size_t s = 0;
    for (size_t c = 0; c < shape.contours.size(); ++c)
        for (size_t e = 0; e < shape.contours[c].edges.size(); ++e, ++s) {
            const EdgeSegment &edge = *shape.contours[c].edges[e];
            const Point2 *p = edge.controlPoints();
            int order = edge.type();
            for (int k = 0; k < 4; ++k) {
                outPts[8*s+2*k] = k <= order ? p[k].x : 0;
                outPts[8*s+2*k+1] = k <= order ? p[k].y : 0;
            }
            outOrders[s] = (uint8_t) order;
        }
}

void hc_scanline_sdf(int N, const float *sdf, int w, int h, const double *xf, double y, int yDown, int nq, const double *xs, int *sum) {
    if (N == 1) sdfScanline<1>(sdf, w, h, xf, y, yDown, nq, xs, sum);
    else if (N == 3) sdfScanline<3>(sdf, w, h, xf, y, yDown, nq, xs, sum);
    else sdfScanline<4>(sdf, w, h, xf, y, yDown, nq, xs, sum);
}

double hc_estimate_error(int N, const float *sdf, int w, int h, int nContours, const uint32_t *cbegin, const double *pts, const uint8_t *orders, int yDown,
                         const double *xf, int scanlinesPerRow, int fillRule) {
    Shape shape;
    buildShape(shape, nContours, cbegin, pts, orders, yDown);
    Projection projection(Vector2(xf[0], xf[1]), Vector2(xf[2], xf[3]));
    if (N == 1) return estimateSDFError(BitmapConstSection<float, 1>(sdf, w, h), shape, projection, scanlinesPerRow, FillRule(fillRule));
    if (N == 3) return estimateSDFError(BitmapConstSection<float, 3>(sdf, w, h), shape, projection, scanlinesPerRow, FillRule(fillRule));
    return estimateSDFError(BitmapConstSection<float, 4>(sdf, w, h), shape, Vector2(xf[0], xf[1]), Vector2(xf[2], xf[3]), scanlinesPerRow, FillRule(fillRule));   // legacy overload
}

}
