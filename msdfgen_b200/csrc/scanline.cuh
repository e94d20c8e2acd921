// scanline.cuh — scanline fill test and distance sign correction on the device.
//
//   k5_scanline_sign<N>       distanceSignCorrection / multiDistanceSignCorrection, first sweep
//                             (/root/reference/core/rasterization.cpp:19-32, :34-74)
//   k5_resolve_ambiguous<N>   second sweep over texels whose median sits exactly on the zero value (:75-94)
//   k5_rasterize              rasterize (:9-17)
//
// The reference builds, per bitmap row, a sorted list of the shape's intersections with the horizontal line through the
// texel centres (Shape::scanline -> EdgeSegment::scanlineIntersections, Shape.cpp:153-165, edge-segments.cpp:279-403) and asks
// Scanline::filled(x) per texel (Scanline.cpp:108-123), which is the fill rule applied to the SUM of crossing directions
// of all intersections with intersection.x <= x.  The sum does not depend on the order, so no sort is needed here:
//
//   * a tile is the same 1024 consecutive texels of one glyph as in k1; one CTA per tile;
//   * phase A: the (row, segment) pairs of the tile are spread over the CTA's threads; each pair yields up to three
//     intersections; an intersection adds its direction to every texel of its row from the first column whose centre is at
//     or right of it — one atomicAdd into a shared-memory DIFFERENCE array (and one at the end of the interval);
//   * phase B: one CTA-wide inclusive scan of the difference array turns it into per-texel sums (row starts are
//     subtracted out), then every thread fixes the signs of its texels with coalesced loads/stores.
//
// Work per tile is rows x segments root solves instead of texels x segments: for a 32x32 glyph 32x fewer than the
// distance kernel, so the pass is bound by reading and writing the field once (HBM), not by the solver.
#pragma once

#include "kernels.cuh"

namespace b200 {

// solveCubic (equation-solver.cpp:63-70)
__device__ __forceinline__ Roots solve_cubic(double a, double b, double c, double d) {
    if (a != 0) {
        double bn = b/a;
        if (fabs(bn) < 1e6)
            return solve_cubic_normed(bn, c/a, d/a);
    }
    return solve_quadratic(b, c, d);
}

struct Crossings { double x[3]; int dy[3]; int n; };

// EdgeSegment::scanlineIntersections for the three segment types.  Curves share one body: the monotone pieces of the
// curve alternate between rising and falling, `next` is the direction the next accepted crossing must have.
__device__ inline Crossings scanline_crossings(const RawSeg &s, double y) {
    Crossings out;
    out.n = 0;
    const d2 a = s.p[0];
    if (s.order == 1) {
        const d2 b = s.p[1];
        if ((y >= a.y && y < b.y) || (y >= b.y && y < a.y)) {
            double t = (y-a.y)/(b.y-a.y);
            out.x[0] = (1.0-t)*a.x+t*b.x;
            out.dy[0] = (0 < b.y-a.y)-(b.y-a.y < 0);
            out.n = 1;
        }
        return out;
    }
    const bool cubic = s.order == 3;
    const int cap = cubic ? 3 : 2;
    const d2 end = cubic ? s.p[3] : s.p[2];
    const d2 ab = s.p[1]-s.p[0];
    const d2 br = s.p[2]-s.p[1]-ab;
    d2 as = mk(0, 0);
    if (cubic)
        as = (s.p[3]-s.p[2])-(s.p[2]-s.p[1])-br;
    int next = y > a.y ? 1 : -1;
    out.x[0] = a.x;
    if (a.y == y) {
        // does the curve leave its start point upwards?  (first control point that is not level with it decides)
        bool up = cubic ? (a.y < s.p[1].y || (a.y == s.p[1].y && (a.y < s.p[2].y || (a.y == s.p[2].y && a.y < s.p[3].y))))
                        : (a.y < s.p[1].y || (a.y == s.p[1].y && a.y < s.p[2].y));
        if (up)
            out.dy[out.n++] = 1;
        else
            next = 1;
    }
    Roots r = cubic ? solve_cubic(as.y, 3*br.y, 3*ab.y, a.y-y) : solve_quadratic(br.y, 2*ab.y, a.y-y);
    // ascending order
    if (r.n >= 2) {
        double tmp;
        if (r.t0 > r.t1) { tmp = r.t0; r.t0 = r.t1; r.t1 = tmp; }
        if (r.n >= 3 && r.t1 > r.t2) {
            tmp = r.t1; r.t1 = r.t2; r.t2 = tmp;
            if (r.t0 > r.t1) { tmp = r.t0; r.t0 = r.t1; r.t1 = tmp; }
        }
    }
    #pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < r.n && out.n < cap) {
            double t = i == 0 ? r.t0 : i == 1 ? r.t1 : r.t2;
            if (t >= 0 && t <= 1) {
                double x, slope;
                if (cubic) {
                    x = a.x+3*t*ab.x+3*t*t*br.x+t*t*t*as.x;
                    slope = ab.y+2*t*br.y+t*t*as.y;
                } else {
                    x = a.x+2*t*ab.x+t*t*br.x;
                    slope = ab.y+t*br.y;
                }
                out.x[out.n] = x;
                if (next*slope >= 0) {
                    out.dy[out.n++] = next;
                    next = -next;
                }
            }
        }
    }
    if (end.y == y) {
        if (next > 0 && out.n > 0) {
            --out.n;
            next = -1;
        }
        // does the curve arrive at its end point from above?
        bool down = cubic ? (end.y < s.p[2].y || (end.y == s.p[2].y && (end.y < s.p[1].y || (end.y == s.p[1].y && end.y < a.y))))
                          : (end.y < s.p[1].y || (end.y == s.p[1].y && end.y < a.y));
        if (down && out.n < cap) {
            out.x[out.n] = end.x;
            if (next < 0) {
                out.dy[out.n++] = -1;
                next = 1;
            }
        }
    }
    if (next != (y >= end.y ? 1 : -1)) {
        if (out.n > 0)
            --out.n;
        else {
            if (fabs(end.y-y) < fabs(a.y-y))
                out.x[out.n] = end.x;
            out.dy[out.n++] = next;
        }
    }
    return out;
}

// interpretFillRule (Scanline.cpp:12-24): 0 FILL_NONZERO, 1 FILL_ODD, 2 FILL_POSITIVE, 3 FILL_NEGATIVE
__device__ __forceinline__ bool interpret_fill_rule(int intersections, int rule) {
    switch (rule) {
        case 0: return intersections != 0;
        case 1: return (intersections&1) != 0;
        case 2: return intersections > 0;
        case 3: return intersections < 0;
    }
    return false;
}

// Projection::unprojectX at the centre of column c of the glyph's block
__device__ __forceinline__ double column_x(const GlyphDesc &g, int c) { return (g.x0+c+.5)/g.sx-g.tx; }

// Per-texel crossing sums of the tile [pix0, tileEnd) of glyph g into s_sum[0 .. tileEnd-pix0).  All threads of the CTA.
__device__ inline void tile_crossing_sums(const RawShapes &raw, const GlyphDesc &g, unsigned pix0, unsigned tileEnd, int *s_sum, int *s_warp) {
    const int w = g.w;
    const unsigned nPix = tileEnd-pix0;
    for (unsigned i = threadIdx.x; i < TILE_PIXELS; i += blockDim.x)
        s_sum[i] = 0;
    __syncthreads();
    const unsigned rowFirst = pix0/unsigned(w), rowLast = (tileEnd-1)/unsigned(w);
    const unsigned nSeg = g.s1-g.s0;
    const unsigned long long pairs = (unsigned long long) (rowLast-rowFirst+1)*nSeg;
    const bool rising = !(g.sx < 0);       // column centres move right with the column index
    for (unsigned long long k = threadIdx.x; k < pairs; k += blockDim.x) {
        unsigned rowOut = rowFirst+unsigned(k/nSeg);
        unsigned e = unsigned(k%nSeg);
        int y = g.flip ? g.h-1-int(rowOut) : int(rowOut);
        double py = (g.y0+y+.5)/g.sy-g.ty;      // Projection::unprojectY
        Crossings cr = scanline_crossings(load_raw(raw, g.s0+e), py);
        // columns of this row that lie inside the tile: [cs, ce)
        long long rowBase = (long long) rowOut*w;
        int cs = rowBase < (long long) pix0 ? int((long long) pix0-rowBase) : 0;
        int ce = rowBase+w > (long long) tileEnd ? int((long long) tileEnd-rowBase) : w;
        for (int i = 0; i < cr.n; ++i) {
            double ix = cr.x[i];
            if (!(ix == ix))
                continue;       // NaN <= x is false for every texel
            // Columns with ix <= column_x(c): an interval [lo, hi) because column_x is monotonic in c (correctly rounded
            // division and subtraction preserve order).  Estimate the boundary, then settle it with the exact comparison.
            double est = rising ? ceil((ix+g.tx)*g.sx-.5-g.x0) : floor((ix+g.tx)*g.sx-.5-g.x0)+1;
            int b = est < 0 ? 0 : est > w ? w : int(est);      // (NaN estimate -> w; fixed below)
            if (!(est == est)) b = 0;
            int lo, hi;
            if (rising) {
                while (b > 0 && ix <= column_x(g, b-1)) --b;
                while (b < w && !(ix <= column_x(g, b))) ++b;
                lo = b; hi = w;
            } else {
                while (b < w && ix <= column_x(g, b)) ++b;
                while (b > 0 && !(ix <= column_x(g, b-1))) --b;
                lo = 0; hi = b;
            }
            if (lo < cs) lo = cs;
            if (lo >= ce || lo >= hi)
                continue;
            atomicAdd(&s_sum[unsigned(rowBase+lo)-pix0], cr.dy[i]);
            if (hi < ce)
                atomicAdd(&s_sum[unsigned(rowBase+hi)-pix0], -cr.dy[i]);
        }
    }
    __syncthreads();
    // CTA-wide inclusive scan of the difference array (4 entries per thread at 256 threads)
    const int PER = TILE_PIXELS/CTA_THREADS;
    int v[PER];
    int run = 0;
    #pragma unroll
    for (int j = 0; j < PER; ++j) {
        run += s_sum[threadIdx.x*PER+j];
        v[j] = run;
    }
    int incl = run;
    const unsigned lane = threadIdx.x&31, warp = threadIdx.x>>5;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= unsigned(d)) incl += o;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    int before = incl-run;
    for (unsigned k = 0; k < warp; ++k)
        before += s_warp[k];
    #pragma unroll
    for (int j = 0; j < PER; ++j)
        s_sum[threadIdx.x*PER+j] = v[j]+before;
    __syncthreads();
    (void) nPix;
}

// Sum of crossing directions at or left of texel i (tile-local index j = i - pix0): scan value minus what earlier rows left.
__device__ __forceinline__ int texel_crossing_sum(const int *s_sum, unsigned i, unsigned pix0, int w) {
    unsigned rowStart = (i/unsigned(w))*unsigned(w);
    unsigned j = i-pix0;
    int sum = s_sum[j];
    if (rowStart > pix0)
        sum -= s_sum[rowStart-pix0-1];
    return sum;
}

// First sweep.  src -> dst (may alias): [h][w][N] blocks in output orientation (k1's layout).  match: per texel, +1 sign agreed, -1 flipped,
// 0 median exactly on the zero value (N >= 3 only).  *ambiguous is set when any texel of the batch got 0.
template <int N>
__global__ void __launch_bounds__(CTA_THREADS) k5_scanline_sign(KernelParams P, const float *src, float *dst, int8_t *match, int *ambiguous, float zero, int fillRule) {
    __shared__ int s_sum[TILE_PIXELS];
    __shared__ int s_warp[CTA_THREADS/32];
    __shared__ GlyphDesc s_g;
    const TileDesc td = P.tiles[blockIdx.x];
    if (threadIdx.x == 0)
        s_g = P.glyphs[td.glyph];
    __syncthreads();
    const int w = s_g.w;
    const unsigned total = unsigned(w)*unsigned(s_g.h);
    const unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
    tile_crossing_sums(P.raw, s_g, td.pix0, tileEnd, s_sum, s_warp);
    const float doubleZero = zero+zero;
    bool sawZero = false;
    for (unsigned i = td.pix0+threadIdx.x; i < tileEnd; i += blockDim.x) {
        bool fill = interpret_fill_rule(texel_crossing_sum(s_sum, i, td.pix0, w), fillRule);
        const float *in = src+s_g.outOffset+size_t(i)*N;
        float *px = dst+s_g.outOffset+size_t(i)*N;
        if (N == 1) {
            float sd = in[0];
            px[0] = (sd > zero) != fill ? doubleZero-sd : sd;
        } else {
            float r = in[0], g = in[1], b = in[2];
            float sd = median3f(r, g, b);
            int8_t m = 1;
            if (sd == zero) {
                m = 0;
                sawZero = true;
            } else if ((sd > zero) != fill) {
                r = doubleZero-r;
                g = doubleZero-g;
                b = doubleZero-b;
                m = -1;
            }
            px[0] = r;
            px[1] = g;
            px[2] = b;
            match[s_g.texelOffset+i] = m;
            if (N >= 4) {
                float a = in[3];
                px[3] = (a > zero) != fill ? doubleZero-a : a;
            }
        }
    }
    if (N > 1 && __syncthreads_or(sawZero) && threadIdx.x == 0)
        *ambiguous = 1;
}

// Second sweep: a texel whose median was exactly the zero value follows the majority of its 4 neighbours.
template <int N>
__global__ void __launch_bounds__(CTA_THREADS) k5_resolve_ambiguous(KernelParams P, float *field, const int8_t *match, const int *ambiguous, float zero) {
    if (*ambiguous == 0)
        return;
    const TileDesc td = P.tiles[blockIdx.x];
    const GlyphDesc g = P.glyphs[td.glyph];
    const int w = g.w, h = g.h;
    const unsigned total = unsigned(w)*unsigned(h);
    const unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
    const int8_t *m = match+g.texelOffset;
    const float doubleZero = zero+zero;
    for (unsigned i = td.pix0+threadIdx.x; i < tileEnd; i += blockDim.x) {
        if (m[i])
            continue;
        int x = int(i%unsigned(w)), y = int(i/unsigned(w));
        int votes = 0;
        if (x > 0) votes += m[i-1];
        if (x < w-1) votes += m[i+1];
        if (y > 0) votes += m[i-unsigned(w)];
        if (y < h-1) votes += m[i+unsigned(w)];
        if (votes < 0) {
            float *px = field+g.outOffset+size_t(i)*N;
            px[0] = doubleZero-px[0];
            px[1] = doubleZero-px[1];
            px[2] = doubleZero-px[2];
        }
    }
}

// rasterize: 1.0f inside, 0.0f outside at texel centres; out: [h][w] blocks laid out like a 1-channel field.
__global__ void __launch_bounds__(CTA_THREADS) k5_rasterize(KernelParams P, float *out, int fillRule) {
    __shared__ int s_sum[TILE_PIXELS];
    __shared__ int s_warp[CTA_THREADS/32];
    __shared__ GlyphDesc s_g;
    const TileDesc td = P.tiles[blockIdx.x];
    if (threadIdx.x == 0)
        s_g = P.glyphs[td.glyph];
    __syncthreads();
    const int w = s_g.w;
    const unsigned total = unsigned(w)*unsigned(s_g.h);
    const unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
    tile_crossing_sums(P.raw, s_g, td.pix0, tileEnd, s_sum, s_warp);
    for (unsigned i = td.pix0+threadIdx.x; i < tileEnd; i += blockDim.x)
        out[s_g.outOffset+i] = interpret_fill_rule(texel_crossing_sum(s_sum, i, td.pix0, w), fillRule) ? 1.f : 0.f;
}

}  // namespace b200
