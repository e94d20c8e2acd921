// kernels.cuh — the sm_100a kernels of the hot path and their launch parameters.
//
//   k0_windings          one thread per contour: Contour::winding()  (/root/reference/core/Contour.cpp:57-81)
//   k0_prepare_big       one thread per segment of shapes too large for shared memory: SegRec -> global buffer
//   k1_distance_field    generateDistanceField<Combiner<Selector>>    (/root/reference/core/msdfgen.cpp:52-76)
//   k2_error_correction  msdfErrorCorrectionInner<N> + apply         (/root/reference/core/msdf-error-correction.cpp:12-48,
//                                                                      core/MSDFErrorCorrection.cpp:121-479)
//
// Work decomposition: a tile is up to TILE_PIXELS consecutive output texels of one glyph; persistent CTAs of CTA_THREADS
// threads pull tiles from an atomic counter (greedy list scheduling: glyph cost varies 8x with its segment count), stage
// the glyph's prepared segment records in shared memory once per tile, then every thread owns one texel per pass and walks
// all records (broadcast shared-memory reads, no divergence on the segment type).  Texels are staged in shared memory and
// written back as coalesced float4 rows.
#pragma once

#include "distance_core.cuh"

namespace b200 {

#define CTA_THREADS 256
#define TILE_PIXELS 1024
#define SMEM_RECORDS 96

struct TileDesc { unsigned glyph, pix0; };

struct KernelParams {
    RawShapes raw;
    const GlyphDesc *glyphs;
    const TileDesc *tiles;
    int nTiles;
    int nContours;
    const SegRec *bigRecs;        // prepared records of big shapes
    int *counters;                // [0] k1 tile counter, [1] k2 tile counter
    float *field;                 // k1 output (k2 input)
    float *out;                   // k2 output
    uint8_t *stencil;             // k2: per-texel flags (1 = ERROR, 2 = PROTECTED), rows in Y-up order like the reference's buffer
    double *scratch;              // per-thread contour-distance slices for many-contour shapes
    size_t scratchStride;         // = gridDim.x * CTA_THREADS
    // error correction config
    int ecMode, ecDcm;
    double minDeviationRatio, minImproveRatio;
    int shapeless;                // msdfFast*ErrorCorrection: no shape, no corner protection, base classifier only
};

// ---------------------------------------------------------------------------------------------------------------------

__global__ void k0_windings(RawShapes raw, int nContours, int8_t *windings, int *counters) {
    int c = blockIdx.x*blockDim.x+threadIdx.x;
    if (c == 0) {
        counters[0] = 0;
        counters[1] = 0;
    }
    if (c < nContours)
        windings[c] = (int8_t) contour_winding(raw, (unsigned) c);
}

// items: (glyph index, evaluation position) pairs are implied by a prefix table: bigGlyphs[i] = glyph index, bigBase[i] = first record
__global__ void k0_prepare_big(RawShapes raw, const GlyphDesc *glyphs, const int *bigGlyphs, int nBig, SegRec *recs) {
    for (int i = blockIdx.y; i < nBig; i += gridDim.y) {
        GlyphDesc g = glyphs[bigGlyphs[i]];
        unsigned n = g.s1-g.s0;
        for (unsigned j = blockIdx.x*blockDim.x+threadIdx.x; j < n; j += gridDim.x*blockDim.x)
            prepare_record(raw, g, j, recs[size_t(g.recBase)+j]);
    }
}

// Fetches the next tile for the CTA (uniform) and (re)stages the glyph's records when the glyph changes.
struct TileCursor {
    int tile;
    int glyph;
};

__device__ __forceinline__ const SegRec *stage_glyph(const KernelParams &P, const GlyphDesc &g, SegRec *smemRecs) {
    if (g.recBase >= 0)
        return P.bigRecs+g.recBase;
    unsigned n = g.s1-g.s0;
    for (unsigned j = threadIdx.x; j < n; j += blockDim.x)
        prepare_record(P.raw, g, j, smemRecs[j]);
    return smemRecs;
}

// Coalesced write-back of nFloats staged floats to dst (16-byte aligned when aligned == true).
__device__ __forceinline__ void store_tile(float *dst, const float *staged, int nFloats) {
    if ((reinterpret_cast<uintptr_t>(dst)&15) == 0) {
        int n4 = nFloats>>2;
        const float4 *s4 = reinterpret_cast<const float4 *>(staged);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int i = threadIdx.x; i < n4; i += blockDim.x)
            d4[i] = s4[i];
        for (int i = (n4<<2)+threadIdx.x; i < nFloats; i += blockDim.x)
            dst[i] = staged[i];
    } else {
        for (int i = threadIdx.x; i < nFloats; i += blockDim.x)
            dst[i] = staged[i];
    }
}

template <int MODE, bool OVERLAP>
__global__ void __launch_bounds__(CTA_THREADS) k1_distance_field(KernelParams P) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    __shared__ SegRec s_recs[SMEM_RECORDS];
    __shared__ __align__(16) float s_out[CTA_THREADS*NOUT];
    __shared__ GlyphDesc s_g;
    __shared__ int s_tile;

    double localStore[OVERLAP ? B200_LOCAL_CONTOURS*NOUT : 1];
    int stagedGlyph = -1;
    const SegRec *recs = nullptr;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            s_tile = atomicAdd(&P.counters[0], 1);
        __syncthreads();
        int tile = s_tile;
        if (tile >= P.nTiles)
            break;
        TileDesc td = P.tiles[tile];
        if ((int) td.glyph != stagedGlyph) {
            if (threadIdx.x == 0)
                s_g = P.glyphs[td.glyph];
            __syncthreads();
            recs = stage_glyph(P, s_g, s_recs);
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const int w = s_g.w;
        const unsigned total = unsigned(w)*unsigned(s_g.h);
        const int nRec = int(s_g.s1-s_g.s0);
        unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
        ContourStore store;
        if (OVERLAP && s_g.nonEmptyContours > B200_LOCAL_CONTOURS) {
            store.base = P.scratch+(size_t(blockIdx.x)*CTA_THREADS+threadIdx.x);
            store.stride = P.scratchStride;
        } else {
            store.base = localStore;
            store.stride = 1;
        }
        for (unsigned passBase = td.pix0; passBase < tileEnd; passBase += CTA_THREADS) {
            unsigned i = passBase+threadIdx.x;
            if (i < tileEnd) {
                int x = int(i%unsigned(w));
                int yOut = int(i/unsigned(w));
                int y = s_g.flip ? s_g.h-1-yOut : yOut;
                // Projection::unproject(Point2(x+.5, y+.5))
                d2 p = mk((x+.5)/s_g.sx-s_g.tx, (y+.5)/s_g.sy-s_g.ty);
                double dist[NOUT];
                shape_distance<MODE, OVERLAP>(recs, nRec, s_g.nonEmptyContours, p, store, dist);
                #pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    s_out[threadIdx.x*NOUT+k] = float(s_g.dmScale*(dist[k]+s_g.dmTranslate));    // DistanceMapping::operator()
            }
            __syncthreads();
            unsigned nPix = tileEnd-passBase < CTA_THREADS ? tileEnd-passBase : CTA_THREADS;
            store_tile(P.field+s_g.outOffset+size_t(passBase)*NOUT, s_out, int(nPix)*NOUT);
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Error correction

#define EC_FLAG_ERROR 1
#define EC_FLAG_PROTECTED 2
#define ARTIFACT_T_EPSILON .01
#define PROTECTION_RADIUS_TOLERANCE 1.001
#define CLS_CANDIDATE 1
#define CLS_ARTIFACT 2

// mix<float, double>(a, b, t)
__device__ __forceinline__ float mixf(float a, float b, double t) { return float((1.0-t)*a+t*b); }

// Read-only view of one glyph's N-channel field in SHAPE orientation.
template <int N>
struct FieldView {
    const float *base;     // first float of the glyph block
    int w, h, flip;
    __device__ __forceinline__ const float *at(int x, int y) const {
        int row = flip ? h-1-y : y;
        return base+(size_t(row)*w+x)*N;
    }
};

__device__ __forceinline__ bool edge_between_channel(const float *a, const float *b, int ch) {
    double t = (a[ch]-.5)/(a[ch]-b[ch]);
    if (t > 0 && t < 1) {
        float c0 = mixf(a[0], b[0], t), c1 = mixf(a[1], b[1], t), c2 = mixf(a[2], b[2], t);
        float cc = ch == 0 ? c0 : ch == 1 ? c1 : c2;
        return median3f(c0, c1, c2) == cc;
    }
    return false;
}

// Would protectEdges mark texel `a` (median am) because of its pair with neighbour `b` (median bm)?
__device__ __forceinline__ bool pair_protects(const float *a, float am, const float *b, float bm, float radius) {
    if (fabsf(am-.5f)+fabsf(bm-.5f) < radius) {
        // edgeBetweenTexels(first, second) is evaluated by the reference with the pair in scan order; the channel test is
        // symmetric in (a, b) only up to rounding of t, so callers pass the pair in the reference's order via `swap`.
        return true;
    }
    return false;
}

template <int N>
struct EcCtx {
    FieldView<N> f;
    const GlyphDesc *g;
    const SegRec *recs;
    int nRec;
    ContourStore store;
    double minImproveRatio;
    int overlap;
};

struct Classifier {
    double span;
    bool protectedFlag;
    bool useShape;
    d2 direction;
    d2 shapeCoord, sdfCoord;
    const float *msd;
};

__device__ __forceinline__ int range_test(const Classifier &k, double at, double bt, double xt, float am, float bm, float xm) {
    if ((am > .5f && bm > .5f && xm <= .5f) || (am < .5f && bm < .5f && xm >= .5f) || (!k.protectedFlag && median3f(am, bm, xm) != xm)) {
        double axSpan = (xt-at)*k.span, bxSpan = (bt-xt)*k.span;
        if (!(xm >= am-axSpan && xm <= am+axSpan && xm >= bm-bxSpan && xm <= bm+bxSpan))
            return CLS_CANDIDATE|CLS_ARTIFACT;
        return CLS_CANDIDATE;
    }
    return 0;
}

__device__ __forceinline__ double clampd(double n, double b) { return n >= 0 && n <= b ? n : double(n > 0)*b; }
__device__ __forceinline__ int clampi(int n, int b) { return n >= 0 && n <= b ? n : int(n > 0)*b; }

// interpolate() of bitmap-interpolation.hpp:10-25 (first three channels are all the caller uses)
template <int N>
__device__ __forceinline__ void interpolate3(float *out, const FieldView<N> &f, d2 pos) {
    pos.x = clampd(pos.x, double(f.w));
    pos.y = clampd(pos.y, double(f.h));
    pos.x -= .5; pos.y -= .5;
    int l = int(floor(pos.x));
    int b = int(floor(pos.y));
    int r = l+1;
    int t = b+1;
    double lr = pos.x-l;
    double bt = pos.y-b;
    l = clampi(l, f.w-1), r = clampi(r, f.w-1);
    b = clampi(b, f.h-1), t = clampi(t, f.h-1);
    const float *lb = f.at(l, b), *rb = f.at(r, b), *lt = f.at(l, t), *rt = f.at(r, t);
    #pragma unroll
    for (int i = 0; i < 3; ++i)
        out[i] = mixf(mixf(lb[i], rb[i], lr), mixf(lt[i], rt[i], lr), bt);
}

// The exact PSDF shape distance used by ShapeDistanceChecker (rare, divergent): kept out of line.
template <bool OVERLAP>
__device__ __noinline__ double exact_psdf(const SegRec *recs, int nRec, int nContours, double px, double py, double *storeBase, size_t storeStride) {
    ContourStore store;
    store.base = storeBase;
    store.stride = storeStride;
    double d;
    shape_distance<2, OVERLAP>(recs, nRec, nContours, mk(px, py), store, &d);
    return d;
}

template <int N>
__device__ __forceinline__ bool cls_evaluate(const EcCtx<N> &e, const Classifier &k, double t, float m, int flags) {
    (void) m;
    if (!k.useShape)
        return (flags&2) != 0;
    if (flags&CLS_CANDIDATE) {
        if (flags&CLS_ARTIFACT)
            return true;
        d2 tVector = t*k.direction;
        float oldMSD[3], newMSD[3];
        d2 sdfCoord = k.sdfCoord+tVector;
        interpolate3<N>(oldMSD, e.f, sdfCoord);
        double aWeight = (1-fabs(tVector.x))*(1-fabs(tVector.y));
        float aPSD = median3f(k.msd[0], k.msd[1], k.msd[2]);
        newMSD[0] = float(oldMSD[0]+aWeight*(aPSD-k.msd[0]));
        newMSD[1] = float(oldMSD[1]+aWeight*(aPSD-k.msd[1]));
        newMSD[2] = float(oldMSD[2]+aWeight*(aPSD-k.msd[2]));
        float oldPSD = median3f(oldMSD[0], oldMSD[1], oldMSD[2]);
        float newPSD = median3f(newMSD[0], newMSD[1], newMSD[2]);
        // texelSize = unprojectVector(Vector2(1))
        double qx = k.shapeCoord.x+tVector.x*(1.0/e.g->sx), qy = k.shapeCoord.y+tVector.y*(1.0/e.g->sy);
        double dist = e.overlap ? exact_psdf<true>(e.recs, e.nRec, e.g->nonEmptyContours, qx, qy, e.store.base, e.store.stride)
                                : exact_psdf<false>(e.recs, e.nRec, e.g->nonEmptyContours, qx, qy, e.store.base, e.store.stride);
        float refPSD = float(e.g->dmScale*(dist+e.g->dmTranslate));
        return e.minImproveRatio*fabsf(newPSD-refPSD) < double(fabsf(oldPSD-refPSD));
    }
    return false;
}

__device__ __forceinline__ float interp_median_lin(const float *a, const float *b, double t) {
    return median3f(mixf(a[0], b[0], t), mixf(a[1], b[1], t), mixf(a[2], b[2], t));
}
__device__ __forceinline__ float interp_median_quad(const float *a, const float *l, const float *q, double t) {
    return float(median3(t*(t*q[0]+l[0])+a[0], t*(t*q[1]+l[1])+a[1], t*(t*q[2]+l[2])+a[2]));
}

template <int N>
__device__ __forceinline__ bool linear_artifact_inner(const EcCtx<N> &e, const Classifier &k, float am, float bm, const float *a, const float *b, float dA, float dB) {
    double t = double(dA)/(dA-dB);
    if (t > ARTIFACT_T_EPSILON && t < 1-ARTIFACT_T_EPSILON) {
        float xm = interp_median_lin(a, b, t);
        return cls_evaluate<N>(e, k, t, xm, range_test(k, 0, 1, t, am, bm, xm));
    }
    return false;
}

template <int N>
__device__ __forceinline__ bool diagonal_artifact_inner(const EcCtx<N> &e, const Classifier &k, float am, float dm, const float *a, const float *l, const float *q,
                                                        float dA, float dBC, float dD, double tEx0, double tEx1) {
    Roots r = solve_quadratic(dD-dBC+dA, dBC-dA-dA, dA);
    for (int i = 0; i < r.n; ++i) {
        double ti = i == 0 ? r.t0 : r.t1;
        if (ti > ARTIFACT_T_EPSILON && ti < 1-ARTIFACT_T_EPSILON) {
            float xm = interp_median_quad(a, l, q, ti);
            int rangeFlags = range_test(k, 0, 1, ti, am, dm, xm);
            if (tEx0 > 0 && tEx0 < 1) {
                double tEnd0 = 0, tEnd1 = 1;
                float em0 = am, em1 = dm;
                float ex = interp_median_quad(a, l, q, tEx0);
                if (tEx0 > ti) { tEnd1 = tEx0; em1 = ex; } else { tEnd0 = tEx0; em0 = ex; }
                rangeFlags |= range_test(k, tEnd0, tEnd1, ti, em0, em1, xm);
            }
            if (tEx1 > 0 && tEx1 < 1) {
                double tEnd0 = 0, tEnd1 = 1;
                float em0 = am, em1 = dm;
                float ex = interp_median_quad(a, l, q, tEx1);
                if (tEx1 > ti) { tEnd1 = tEx1; em1 = ex; } else { tEnd0 = tEx1; em0 = ex; }
                rangeFlags |= range_test(k, tEnd0, tEnd1, ti, em0, em1, xm);
            }
            if (cls_evaluate<N>(e, k, ti, xm, rangeFlags))
                return true;
        }
    }
    return false;
}

template <int N>
__device__ __forceinline__ bool has_linear_artifact(const EcCtx<N> &e, const Classifier &k, float am, const float *a, const float *b) {
    float bm = median3f(b[0], b[1], b[2]);
    return fabsf(am-.5f) >= fabsf(bm-.5f) && (
        linear_artifact_inner<N>(e, k, am, bm, a, b, a[1]-a[0], b[1]-b[0]) ||
        linear_artifact_inner<N>(e, k, am, bm, a, b, a[2]-a[1], b[2]-b[1]) ||
        linear_artifact_inner<N>(e, k, am, bm, a, b, a[0]-a[2], b[0]-b[2]));
}

template <int N>
__device__ __forceinline__ bool has_diagonal_artifact(const EcCtx<N> &e, const Classifier &k, float am, const float *a, const float *b, const float *c, const float *d) {
    float dm = median3f(d[0], d[1], d[2]);
    if (fabsf(am-.5f) >= fabsf(dm-.5f)) {
        float abc[3] = { a[0]-b[0]-c[0], a[1]-b[1]-c[1], a[2]-b[2]-c[2] };
        float l[3] = { -a[0]-abc[0], -a[1]-abc[1], -a[2]-abc[2] };
        float q[3] = { d[0]+abc[0], d[1]+abc[1], d[2]+abc[2] };
        double tEx[3] = { -.5*l[0]/q[0], -.5*l[1]/q[1], -.5*l[2]/q[2] };
        return
            diagonal_artifact_inner<N>(e, k, am, dm, a, l, q, a[1]-a[0], b[1]-b[0]+c[1]-c[0], d[1]-d[0], tEx[0], tEx[1]) ||
            diagonal_artifact_inner<N>(e, k, am, dm, a, l, q, a[2]-a[1], b[2]-b[1]+c[2]-c[1], d[2]-d[1], tEx[1], tEx[2]) ||
            diagonal_artifact_inner<N>(e, k, am, dm, a, l, q, a[0]-a[2], b[0]-b[2]+c[0]-c[2], d[0]-d[2], tEx[2], tEx[0]);
    }
    return false;
}

// findErrors for one texel (base classifier or shape-distance classifier), MSDFErrorCorrection.cpp:383-457
template <int N>
__device__ __noinline__ bool find_errors_texel(const EcCtx<N> &e, int x, int y, bool protectedFlag, bool useShape, double hSpan, double vSpan, double dSpan) {
    const int w = e.f.w, h = e.f.h;
    const float *c = e.f.at(x, y);
    float cm = median3f(c[0], c[1], c[2]);
    Classifier k;
    k.protectedFlag = protectedFlag;
    k.useShape = useShape;
    k.sdfCoord = mk(x+.5, y+.5);
    k.shapeCoord = mk((x+.5)/e.g->sx-e.g->tx, (y+.5)/e.g->sy-e.g->ty);
    k.msd = c;
    const float *l = x > 0 ? e.f.at(x-1, y) : nullptr, *b = y > 0 ? e.f.at(x, y-1) : nullptr;
    const float *r = x < w-1 ? e.f.at(x+1, y) : nullptr, *t = y < h-1 ? e.f.at(x, y+1) : nullptr;
    if (x > 0) { k.direction = mk(-1, 0); k.span = hSpan; if (has_linear_artifact<N>(e, k, cm, c, l)) return true; }
    if (y > 0) { k.direction = mk(0, -1); k.span = vSpan; if (has_linear_artifact<N>(e, k, cm, c, b)) return true; }
    if (x < w-1) { k.direction = mk(+1, 0); k.span = hSpan; if (has_linear_artifact<N>(e, k, cm, c, r)) return true; }
    if (y < h-1) { k.direction = mk(0, +1); k.span = vSpan; if (has_linear_artifact<N>(e, k, cm, c, t)) return true; }
    k.span = dSpan;
    if (x > 0 && y > 0) { k.direction = mk(-1, -1); if (has_diagonal_artifact<N>(e, k, cm, c, l, b, e.f.at(x-1, y-1))) return true; }
    if (x < w-1 && y > 0) { k.direction = mk(+1, -1); if (has_diagonal_artifact<N>(e, k, cm, c, r, b, e.f.at(x+1, y-1))) return true; }
    if (x > 0 && y < h-1) { k.direction = mk(-1, +1); if (has_diagonal_artifact<N>(e, k, cm, c, l, t, e.f.at(x-1, y+1))) return true; }
    if (x < w-1 && y < h-1) { k.direction = mk(+1, +1); if (has_diagonal_artifact<N>(e, k, cm, c, r, t, e.f.at(x+1, y+1))) return true; }
    return false;
}

// protectEdges restated as a gather: is texel (x,y) protected by any of the (up to 8) texel pairs it belongs to?
// (MSDFErrorCorrection.cpp:189-250; pairs are evaluated with the same (first, second) roles as the reference's scan.)
template <int N>
__device__ __forceinline__ bool protect_edges_texel(const FieldView<N> &f, int x, int y, float hRadius, float vRadius, float dRadius) {
    const float *c = f.at(x, y);
    float cm = median3f(c[0], c[1], c[2]);
    const int w = f.w, h = f.h;
    bool prot = false;
    // pair (first, second): returns whether `c` (which is first when cIsFirst) gets protected
    #define B200_PAIR(nx, ny, radius, cIsFirst) do { \
        const float *o = f.at(nx, ny); \
        float om = median3f(o[0], o[1], o[2]); \
        if (fabsf((cIsFirst ? cm : om)-.5f)+fabsf((cIsFirst ? om : cm)-.5f) < radius) { \
            const float *a = cIsFirst ? c : o, *bb = cIsFirst ? o : c; \
            int mask = (edge_between_channel(a, bb, 0) ? 1 : 0)|(edge_between_channel(a, bb, 1) ? 2 : 0)|(edge_between_channel(a, bb, 2) ? 4 : 0); \
            if (((mask&1) && c[0] != cm) || ((mask&2) && c[1] != cm) || ((mask&4) && c[2] != cm)) \
                prot = true; \
        } \
    } while (0)
    if (x < w-1) B200_PAIR(x+1, y, hRadius, true);               // horizontal: left = c
    if (x > 0) B200_PAIR(x-1, y, hRadius, false);                // horizontal: right = c
    if (y < h-1) B200_PAIR(x, y+1, vRadius, true);               // vertical: bottom = c
    if (y > 0) B200_PAIR(x, y-1, vRadius, false);                // vertical: top = c
    if (x < w-1 && y < h-1) B200_PAIR(x+1, y+1, dRadius, true);  // diagonal lb(c) - rt
    if (x > 0 && y > 0) B200_PAIR(x-1, y-1, dRadius, false);     // diagonal lb - rt(c)
    if (x > 0 && y < h-1) B200_PAIR(x-1, y+1, dRadius, true);    // diagonal rb(c) - lt
    if (x < w-1 && y > 0) B200_PAIR(x+1, y-1, dRadius, false);   // diagonal rb - lt(c)
    #undef B200_PAIR
    return prot;
}

template <int N, bool OVERLAP>
__global__ void __launch_bounds__(CTA_THREADS) k2_error_correction(KernelParams P) {
    __shared__ SegRec s_recs[SMEM_RECORDS];
    __shared__ __align__(16) float s_out[CTA_THREADS*N];
    __shared__ GlyphDesc s_g;
    __shared__ int s_tile;

    double localStore[B200_LOCAL_CONTOURS];
    int stagedGlyph = -1;
    const SegRec *recs = nullptr;
    const bool needShape = !P.shapeless && (P.ecDcm == 1 || P.ecDcm == 2);
    const bool needCorners = !P.shapeless && P.ecMode == 2;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            s_tile = atomicAdd(&P.counters[1], 1);
        __syncthreads();
        int tile = s_tile;
        if (tile >= P.nTiles)
            break;
        TileDesc td = P.tiles[tile];
        if ((int) td.glyph != stagedGlyph) {
            if (threadIdx.x == 0)
                s_g = P.glyphs[td.glyph];
            __syncthreads();
            if (needShape || needCorners)
                recs = stage_glyph(P, s_g, s_recs);
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const int w = s_g.w, h = s_g.h;
        const unsigned total = unsigned(w)*unsigned(h);
        const int nRec = (needShape || needCorners) ? int(s_g.s1-s_g.s0) : 0;
        unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;

        EcCtx<N> e;
        e.f.base = P.field+s_g.outOffset;
        e.f.w = w; e.f.h = h; e.f.flip = s_g.flip;
        e.g = &s_g;
        e.recs = recs;
        e.nRec = nRec;
        e.minImproveRatio = P.minImproveRatio;
        e.overlap = OVERLAP;
        if (OVERLAP && s_g.nonEmptyContours > B200_LOCAL_CONTOURS) {
            e.store.base = P.scratch+(size_t(blockIdx.x)*CTA_THREADS+threadIdx.x);
            e.store.stride = P.scratchStride;
        } else {
            e.store.base = localStore;
            e.store.stride = 1;
        }
        // distanceMapping(Delta(1)) = dmScale; unprojectVector(v) = v/scale
        const double dm1 = s_g.dmScale*1;
        const double hLen = len(mk(dm1/s_g.sx, 0/s_g.sy));
        const double vLen = len(mk(0/s_g.sx, dm1/s_g.sy));
        const double dLen = len(mk(dm1/s_g.sx, dm1/s_g.sy));
        const float hRadius = float(PROTECTION_RADIUS_TOLERANCE*hLen);
        const float vRadius = float(PROTECTION_RADIUS_TOLERANCE*vLen);
        const float dRadius = float(PROTECTION_RADIUS_TOLERANCE*dLen);
        const double hSpan = P.minDeviationRatio*hLen, vSpan = P.minDeviationRatio*vLen, dSpan = P.minDeviationRatio*dLen;

        for (unsigned passBase = td.pix0; passBase < tileEnd; passBase += CTA_THREADS) {
            unsigned i = passBase+threadIdx.x;
            if (i < tileEnd) {
                int x = int(i%unsigned(w));
                int yOut = int(i/unsigned(w));
                int y = s_g.flip ? h-1-yOut : yOut;
                int flags = 0;
                // msdfErrorCorrectionInner: protection stage
                if (P.ecMode == 2) {
                    bool prot = false;
                    if (needCorners)
                        for (int j = 0; j < nRec; ++j) {
                            int cl = recs[j].cornerL, cb = recs[j].cornerB;
                            if (unsigned(x-cl) <= 1u && unsigned(y-cb) <= 1u)
                                prot = true;
                        }
                    if (!prot)
                        prot = protect_edges_texel<N>(e.f, x, y, hRadius, vRadius, dRadius);
                    if (prot)
                        flags |= EC_FLAG_PROTECTED;
                } else if (P.ecMode == 3)
                    flags |= EC_FLAG_PROTECTED;
                // detection stage
                if (P.shapeless) {
                    if (find_errors_texel<N>(e, x, y, (flags&EC_FLAG_PROTECTED) != 0, false, hSpan, vSpan, dSpan))
                        flags |= EC_FLAG_ERROR;
                } else {
                    if (P.ecDcm == 0 || (P.ecDcm == 1 && P.ecMode != 3)) {
                        if (find_errors_texel<N>(e, x, y, (flags&EC_FLAG_PROTECTED) != 0, false, hSpan, vSpan, dSpan))
                            flags |= EC_FLAG_ERROR;
                        if (P.ecDcm == 1)
                            flags |= EC_FLAG_PROTECTED;    // protectAll
                    }
                    if ((P.ecDcm == 2 || P.ecDcm == 1) && !(flags&EC_FLAG_ERROR)) {
                        if (find_errors_texel<N>(e, x, y, (flags&EC_FLAG_PROTECTED) != 0, true, hSpan, vSpan, dSpan))
                            flags |= EC_FLAG_ERROR;
                    }
                }
                // apply
                const float *c = e.f.at(x, y);
                float o0 = c[0], o1 = c[1], o2 = c[2];
                if (flags&EC_FLAG_ERROR) {
                    float m = median3f(o0, o1, o2);
                    o0 = m; o1 = m; o2 = m;
                }
                s_out[threadIdx.x*N+0] = o0;
                s_out[threadIdx.x*N+1] = o1;
                s_out[threadIdx.x*N+2] = o2;
                if (N == 4)
                    s_out[threadIdx.x*N+3] = c[3];
                if (P.stencil) {
                    int srow = s_g.shapeYDown ? h-1-y : y;
                    P.stencil[s_g.texelOffset+size_t(srow)*w+x] = (uint8_t) flags;
                }
            }
            __syncthreads();
            unsigned nPix = tileEnd-passBase < CTA_THREADS ? tileEnd-passBase : CTA_THREADS;
            store_tile(P.out+s_g.outOffset+size_t(passBase)*N, s_out, int(nPix)*N);
            __syncthreads();
        }
    }
}

}  // namespace b200
