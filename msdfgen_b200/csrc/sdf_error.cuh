// sdf_error.cuh — estimateSDFError on the device (/root/reference/core/sdf-error-estimation.cpp:134-154).
//
// The reference's validation metric: for every bitmap row and `scanlinesPerRow` sub-rows it intersects the horizontal line with
// the SHAPE (Shape::scanline: every edge's scanlineIntersections, sorted, directions prefix-summed — Shape.cpp:153-165,
// Scanline.cpp:72-84) and with the DISTANCE FIELD (scanlineSDF: the zero crossings of the bilinearly interpolated median along
// the row, :8-121), and measures the length on which the two disagree about inside/outside (Scanline::overlap,
// Scanline.cpp:26-60).  Scanlines are independent, so one warp takes one scanline:
//   * the lanes share the edges: crossings (scanline_crossings, scanline.cuh) are appended to a shared-memory list, which is then
//     sorted by x with a bitonic network; ties need no care — between equal abscissae the merge below adds zero length and the
//     state after the last of them is the cumulative sum, whatever their order;
//   * lane 0 runs the two sequential parts exactly as the reference does: the walk along the field's row with its inside/outside
//     state and consistency check, and the two-list merge;
//   * result: 1 - overlap/(xTo - xFrom) per scanline; the host adds them in the reference's order (row-major) in fp64.
#pragma once

#include "scanline.cuh"

namespace b200 {

#define SE_THREADS 32
#define SE_CAP 2048             // crossings per scanline (shape side); a scanline with more is reported as NaN

struct SdfErrorParams {
    RawShapes raw;
    const GlyphDesc *glyphs;
    const float *sdf;               // packed [h][w][N] blocks at GlyphDesc::outOffset, rows as in the caller's memory
    const unsigned *scanBase;       // [nGlyphs+1] first scanline of every glyph
    int nGlyphs;
    int scanlinesPerRow, fillRule;
    double *out;                    // [total scanlines]
    double *slabX;                  // per CTA: slabStride abscissae / directions of the field's intersections
    int *slabDir;
    size_t slabStride;              // >= 3*max width + 4
    unsigned totalScanlines;
};

template <int N>
__device__ __forceinline__ float se_row_value(const float *rowB, const float *rowT, int x, int ch, double bt) {
    return float((1.0-bt)*rowB[size_t(x)*N+ch]+bt*rowT[size_t(x)*N+ch]);       // mix<float, double>
}

template <int N>
__global__ void __launch_bounds__(SE_THREADS) k7_sdf_error(SdfErrorParams P) {
    __shared__ double s_x[SE_CAP];
    __shared__ int s_dir[SE_CAP];
    __shared__ int s_count;
    const unsigned lane = threadIdx.x;
    for (unsigned scan = blockIdx.x; scan < P.totalScanlines; scan += gridDim.x) {
    __syncwarp();
    // glyph of this scanline
    int lo = 0, hi = P.nGlyphs;
    while (hi-lo > 1) {
        int mid = (lo+hi)>>1;
        if (P.scanBase[mid] <= scan) lo = mid; else hi = mid;
    }
    const GlyphDesc g = P.glyphs[lo];
    const unsigned local = scan-P.scanBase[lo];
    const int row = int(local/unsigned(P.scanlinesPerRow)), subRow = int(local%unsigned(P.scanlinesPerRow));
    const double subRowSize = 1./P.scanlinesPerRow;
    const double bt0 = (subRow+.5)*subRowSize;
    const double y = (row+bt0+.5)/g.sy-g.ty;            // projection.unprojectY(row+bt+.5)
    const double xFrom = .5/g.sx-g.tx;                  // projection.unprojectX(.5)
    const double xTo = (g.w-.5)/g.sx-g.tx;
    if (!(xFrom < xTo)) {
        // mirrored or degenerate projection: the reference's merge loop does not terminate for xTo <= xFrom; report NaN
        if (lane == 0)
            P.out[scan] = __longlong_as_double(0x7ff8000000000000ll);
        continue;
    }
    if (lane == 0)
        s_count = 0;
    __syncwarp();
    // ---- shape side: all crossings of the line with the edges
    for (unsigned e = g.s0+lane; e < g.s1; e += SE_THREADS) {
        Crossings cr = scanline_crossings(load_raw(P.raw, e), y);
        for (int i = 0; i < cr.n; ++i) {
            int at = atomicAdd(&s_count, 1);
            if (at < SE_CAP) {
                s_x[at] = cr.x[i];
                s_dir[at] = cr.dy[i];
            }
        }
    }
    __syncwarp();
    const int nA = s_count;
    if (nA > SE_CAP) {
        if (lane == 0)
            P.out[scan] = __longlong_as_double(0x7ff8000000000000ll);
        continue;
    }
    // bitonic sort by x (NaN abscissae sort last, like values the reference's comparator calls "equal" they cannot be ordered;
    // shapes with finite coordinates produce none)
    int m = 1;
    while (m < nA) m <<= 1;
    for (int i = nA+int(lane); i < m; i += SE_THREADS) {
        s_x[i] = __longlong_as_double(0x7ff0000000000000ll);       // +inf padding
        s_dir[i] = 0;
    }
    __syncwarp();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k>>1; j > 0; j >>= 1) {
            for (int i = int(lane); i < m; i += SE_THREADS) {
                int l = i^j;
                if (l > i) {
                    const bool up = (i&k) == 0;
                    const double a = s_x[i], b = s_x[l];
                    if (up ? (a > b || (a != a && b == b)) : (a < b || (b != b && a == a))) {
                        s_x[i] = b; s_x[l] = a;
                        int t = s_dir[i]; s_dir[i] = s_dir[l]; s_dir[l] = t;
                    }
                }
            }
            __syncwarp();
        }
    if (lane != 0)
        continue;
    // Scanline::preprocess: directions -> running sums
    {
        int total = 0;
        for (int i = 0; i < nA; ++i) {
            total += s_dir[i];
            s_dir[i] = total;
        }
    }
    // ---- field side and merge, sequential like the reference
    const int w = g.w, h = g.h;
    double pixelY = (y+g.ty)*g.sy-.5;                                   // projection.projectY(y)-.5
    pixelY = pixelY >= 0 && pixelY <= double(h-1) ? pixelY : double(pixelY > 0)*double(h-1);       // clamp(pixelY, h-1)
    if (g.shapeYDown)
        pixelY = h-1-pixelY;
    int b = int(floor(pixelY));
    int t = b+1;
    double bt = pixelY-b;
    if (t >= h) {
        b = h-1;
        t = h-1;
        bt = 1;
    }
    const float *rowB = P.sdf+g.outOffset+size_t(b)*w*N, *rowT = P.sdf+g.outOffset+size_t(t)*w*N;
    // field intersections: this CTA's slab (the walk may pop what it pushed, so the list is completed before it is merged)
    double *const bxs = P.slabX+size_t(blockIdx.x)*P.slabStride;
    int *const bds = P.slabDir+size_t(blockIdx.x)*P.slabStride;
    int nB = 0;
    bool inside = false;
    if (N == 1) {
        float lv, rv = se_row_value<N>(rowB, rowT, 0, 0, bt);
        if ((inside = rv > .5f)) {
            bxs[nB] = -1e240; bds[nB] = 1; ++nB;
        }
        for (int l = 0, r = 1; r < w; ++l, ++r) {
            lv = rv;
            rv = se_row_value<N>(rowB, rowT, r, 0, bt);
            if (lv != rv) {
                double lr = double(.5f-lv)/double(rv-lv);
                if (lr >= 0 && lr <= 1) {
                    bxs[nB] = (l+lr+.5)/g.sx-g.tx;             // projection.unprojectX(l+lr+.5)
                    bds[nB] = (rv-lv > 0)-(rv-lv < 0);
                    ++nB;
                }
            }
        }
    } else {
        float lv[3], rv[3];
        rv[0] = se_row_value<N>(rowB, rowT, 0, 0, bt);
        rv[1] = se_row_value<N>(rowB, rowT, 0, 1, bt);
        rv[2] = se_row_value<N>(rowB, rowT, 0, 2, bt);
        if ((inside = median3f(rv[0], rv[1], rv[2]) > .5f)) {
            bxs[nB] = -1e240; bds[nB] = 1; ++nB;
        }
        for (int l = 0, r = 1; r < w; ++l, ++r) {
            lv[0] = rv[0], lv[1] = rv[1], lv[2] = rv[2];
            rv[0] = se_row_value<N>(rowB, rowT, r, 0, bt);
            rv[1] = se_row_value<N>(rowB, rowT, r, 1, bt);
            rv[2] = se_row_value<N>(rowB, rowT, r, 2, bt);
            double nx[4];
            int nd[4];
            int n = 0;
            for (int i = 0; i < 3; ++i) {
                if (lv[i] != rv[i]) {
                    double lr = double(.5f-lv[i])/double(rv[i]-lv[i]);
                    if (lr >= 0 && lr <= 1) {
                        float v0 = float((1.0-lr)*lv[0]+lr*rv[0]), v1 = float((1.0-lr)*lv[1]+lr*rv[1]), v2 = float((1.0-lr)*lv[2]+lr*rv[2]);
                        float vi = i == 0 ? v0 : i == 1 ? v1 : v2;
                        if (median3f(v0, v1, v2) == vi) {
                            nx[n] = (l+lr+.5)/g.sx-g.tx;
                            nd[n] = (rv[i]-lv[i] > 0)-(rv[i]-lv[i] < 0);
                            ++n;
                        }
                    }
                }
            }
            // sort the new intersections
            if (n >= 2) {
                if (nx[0] > nx[1]) { nx[3] = nx[0]; nd[3] = nd[0]; nx[0] = nx[1]; nd[0] = nd[1]; nx[1] = nx[3]; nd[1] = nd[3]; }
                if (n >= 3 && nx[1] > nx[2]) {
                    nx[3] = nx[1]; nd[3] = nd[1]; nx[1] = nx[2]; nd[1] = nd[2]; nx[2] = nx[3]; nd[2] = nd[3];
                    if (nx[0] > nx[1]) { nx[3] = nx[0]; nd[3] = nd[0]; nx[0] = nx[1]; nd[0] = nd[1]; nx[1] = nx[3]; nd[1] = nd[3]; }
                }
            }
            for (int i = 0; i < n; ++i) {
                if ((nd[i] > 0) == !inside) {
                    bxs[nB] = nx[i]; bds[nB] = nd[i]; ++nB;
                    inside = !inside;
                }
            }
            // consistency check
            float rvScalar = median3f(rv[0], rv[1], rv[2]);
            if ((rvScalar > .5f) != inside && rvScalar != .5f && nB > 0) {
                --nB;
                inside = !inside;
            }
        }
    }
    // Scanline::preprocess of the field's list: qsort by x, then running sums.  The abscissae are produced in ascending order for
    // a projection with positive x scale; otherwise (mirrored projection) they descend — an insertion sort restores the order the
    // reference's qsort would (ties: see the header).
    for (int i = 1; i < nB; ++i) {
        double kx = bxs[i];
        int kd = bds[i];
        int j = i-1;
        while (j >= 0 && bxs[j] > kx) {
            bxs[j+1] = bxs[j]; bds[j+1] = bds[j];
            --j;
        }
        bxs[j+1] = kx; bds[j+1] = kd;
    }
    {
        int run = 0;
        for (int i = 0; i < nB; ++i) {
            run += bds[i];
            bds[i] = run;
        }
    }
    // Scanline::overlap(ref, sdf, xFrom, xTo, fillRule), verbatim
    double total = 0;
    bool aInside = false, bInside = false;
    int ai = 0, bi = 0;
    double ax = nA > 0 ? s_x[ai] : xTo;
    double bx = nB > 0 ? bxs[bi] : xTo;
    while (ax < xFrom || bx < xFrom) {
        double xNext = bx < ax ? bx : ax;           // min(ax, bx)
        if (ax == xNext && ai < nA) {
            aInside = interpret_fill_rule(s_dir[ai], P.fillRule);
            ax = ++ai < nA ? s_x[ai] : xTo;
        }
        if (bx == xNext && bi < nB) {
            bInside = interpret_fill_rule(bds[bi], P.fillRule);
            bx = ++bi < nB ? bxs[bi] : xTo;
        }
    }
    double x = xFrom;
    while (ax < xTo || bx < xTo) {
        double xNext = bx < ax ? bx : ax;
        if (aInside == bInside)
            total += xNext-x;
        if (ax == xNext && ai < nA) {
            aInside = interpret_fill_rule(s_dir[ai], P.fillRule);
            ax = ++ai < nA ? s_x[ai] : xTo;
        }
        if (bx == xNext && bi < nB) {
            bInside = interpret_fill_rule(bds[bi], P.fillRule);
            bx = ++bi < nB ? bxs[bi] : xTo;
        }
        x = xNext;
    }
    if (aInside == bInside)
        total += xTo-x;
    const double overlapFactor = 1/(xTo-xFrom);
    P.out[scan] = 1-overlapFactor*total;
    }
}

}  // namespace b200
