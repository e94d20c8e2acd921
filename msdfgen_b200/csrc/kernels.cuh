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

// A sub-texel point whose exact shape distance decides whether its texel is an artifact (ShapeDistanceChecker, rare).
// The producing thread finishes everything that only needs the field (interpolation, before/after medians); the exact
// PSDF distance at (qx, qy) is evaluated later by whichever thread picks the record up, so that a warp evaluates 32
// different candidates instead of idling 31 lanes behind one.
struct __align__(16) Candidate {
    double qx, qy;
    float oldPSD, newPSD;
    int texel;      // thread index of the owning texel in the current pass
    int pad;
};

template <int N>
struct EcCtx {
    FieldView<N> f;
    const GlyphDesc *g;
    const SegRec *recs;
    int nRec;
    ContourStore store;
    double minImproveRatio;
    Candidate *queue;
    int *queueCount;
    int queueCapacity;
};

// Per-texel classification state shared by the base pass (findErrors<N>(sdf)) and the shape pass (findErrors(sdf, shape)):
// both walk the same neighbour tests and differ only in the protected flag fed to rangeTest and in what a CANDIDATE means.
struct TexelState {
    bool useBase, protBase;      // BaseArtifactClassifier pass enabled / its protectedFlag
    bool useShape, protShape;    // ShapeDistanceChecker pass enabled / its protectedFlag
    bool errBase, errShape;      // decided without an exact distance
    double span;
    d2 direction;
    d2 shapeCoord, sdfCoord;
    const float *msd;
    int texel;
};

__device__ __forceinline__ int range_test(bool protectedFlag, double span, double at, double bt, double xt, float am, float bm, float xm) {
    if ((am > .5f && bm > .5f && xm <= .5f) || (am < .5f && bm < .5f && xm >= .5f) || (!protectedFlag && median3f(am, bm, xm) != xm)) {
        double axSpan = (xt-at)*span, bxSpan = (bt-xt)*span;
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

// The exact PSDF shape distance used by ShapeDistanceChecker: kept out of line.
template <bool OVERLAP>
__device__ __noinline__ double exact_psdf(const SegRec *recs, int nRec, int nContours, double px, double py, double *storeBase, size_t storeStride) {
    ContourStore store;
    store.base = storeBase;
    store.stride = storeStride;
    double d;
    shape_distance<2, OVERLAP>(recs, nRec, nContours, mk(px, py), store, &d);
    return d;
}

// ArtifactClassifier::evaluate's comparison once the exact distance is known (MSDFErrorCorrection.cpp:79)
template <int N, bool OVERLAP>
__device__ __forceinline__ bool candidate_is_artifact(const EcCtx<N> &e, const Candidate &c) {
    double dist = exact_psdf<OVERLAP>(e.recs, e.nRec, e.g->nonEmptyContours, c.qx, c.qy, e.store.base, e.store.stride);
    float refPSD = float(e.g->dmScale*(dist+e.g->dmTranslate));
    return e.minImproveRatio*fabsf(c.newPSD-refPSD) < double(fabsf(c.oldPSD-refPSD));
}

// Queues the exact-distance check of ArtifactClassifier::evaluate (MSDFErrorCorrection.cpp:59-79) — rare, out of line.
template <int N, bool OVERLAP>
__device__ __noinline__ void push_candidate(const EcCtx<N> &e, TexelState &k, double t) {
    d2 tVector = t*k.direction;
    float oldMSD[3], newMSD[3];
    interpolate3<N>(oldMSD, e.f, k.sdfCoord+tVector);
    double aWeight = (1-fabs(tVector.x))*(1-fabs(tVector.y));
    float aPSD = median3f(k.msd[0], k.msd[1], k.msd[2]);
    newMSD[0] = float(oldMSD[0]+aWeight*(aPSD-k.msd[0]));
    newMSD[1] = float(oldMSD[1]+aWeight*(aPSD-k.msd[1]));
    newMSD[2] = float(oldMSD[2]+aWeight*(aPSD-k.msd[2]));
    Candidate c;
    c.oldPSD = median3f(oldMSD[0], oldMSD[1], oldMSD[2]);
    c.newPSD = median3f(newMSD[0], newMSD[1], newMSD[2]);
    // shapeCoord + tVector*texelSize, texelSize = unprojectVector(Vector2(1))
    c.qx = k.shapeCoord.x+tVector.x*(1.0/e.g->sx);
    c.qy = k.shapeCoord.y+tVector.y*(1.0/e.g->sy);
    c.texel = k.texel;
    c.pad = 0;
    int slot = atomicAdd(e.queueCount, 1);
    if (slot < e.queueCapacity)
        e.queue[slot] = c;
    else if (candidate_is_artifact<N, OVERLAP>(e, c))    // queue full: evaluate in place (slow path, same result)
        k.errShape = true;
}

// Classifier::evaluate for both passes at interpolation ratio t with the two passes' range flags.
template <int N, bool OVERLAP>
__device__ __forceinline__ void classify_point(const EcCtx<N> &e, TexelState &k, double t, int flagsBase, int flagsShape) {
    if (k.useBase && (flagsBase&CLS_ARTIFACT))
        k.errBase = true;
    if (k.useShape && (flagsShape&CLS_CANDIDATE)) {
        if (flagsShape&CLS_ARTIFACT)
            k.errShape = true;
        else if (!(k.errBase || k.errShape))    // once an error, the exact distance cannot change the outcome
            push_candidate<N, OVERLAP>(e, k, t);
    }
}

__device__ __forceinline__ float interp_median_lin(const float *a, const float *b, double t) {
    return median3f(mixf(a[0], b[0], t), mixf(a[1], b[1], t), mixf(a[2], b[2], t));
}
__device__ __forceinline__ float interp_median_quad(const float *a, const float *l, const float *q, double t) {
    return float(median3(t*(t*q[0]+l[0])+a[0], t*(t*q[1]+l[1])+a[1], t*(t*q[2]+l[2])+a[2]));
}

template <int N, bool OVERLAP>
__device__ __forceinline__ void linear_artifact_inner(const EcCtx<N> &e, TexelState &k, float am, float bm, const float *a, const float *b, float dA, float dB) {
    // t = dA/(dA-dB) lies strictly inside (0,1) only if dA and dB have strictly opposite signs; in every other case the
    // reference's range test (t > .01 && t < .99) fails as well, so the fp64 division is skipped without changing a result.
    if (!((dA > 0 && dB < 0) || (dA < 0 && dB > 0)))
        return;
    double t = double(dA)/(dA-dB);
    if (t > ARTIFACT_T_EPSILON && t < 1-ARTIFACT_T_EPSILON) {
        float xm = interp_median_lin(a, b, t);
        int fb = k.useBase ? range_test(k.protBase, k.span, 0, 1, t, am, bm, xm) : 0;
        int fs = k.useShape ? range_test(k.protShape, k.span, 0, 1, t, am, bm, xm) : 0;
        classify_point<N, OVERLAP>(e, k, t, fb, fs);
    }
}

template <int N, bool OVERLAP>
__device__ __forceinline__ void diagonal_artifact_inner(const EcCtx<N> &e, TexelState &k, float am, float dm, const float *a, const float *l, const float *q,
                                                        float dA, float dBC, float dD, int ex0, int ex1) {
    Roots r = solve_quadratic(dD-dBC+dA, dBC-dA-dA, dA);
    for (int i = 0; i < r.n; ++i) {
        double ti = i == 0 ? r.t0 : r.t1;
        if (ti > ARTIFACT_T_EPSILON && ti < 1-ARTIFACT_T_EPSILON) {
            float xm = interp_median_quad(a, l, q, ti);
            int fb = k.useBase ? range_test(k.protBase, k.span, 0, 1, ti, am, dm, xm) : 0;
            int fs = k.useShape ? range_test(k.protShape, k.span, 0, 1, ti, am, dm, xm) : 0;
            #pragma unroll
            for (int ex = 0; ex < 2; ++ex) {
                int ch = ex == 0 ? ex0 : ex1;
                // local extreme of channel ch along the diagonal: 2*q*t + l == 0
                float lc = ch == 0 ? l[0] : ch == 1 ? l[1] : l[2];
                float qc = ch == 0 ? q[0] : ch == 1 ? q[1] : q[2];
                double tEx = -.5*lc/qc;
                if (tEx > 0 && tEx < 1) {
                    double tEnd0 = 0, tEnd1 = 1;
                    float em0 = am, em1 = dm;
                    float exm = interp_median_quad(a, l, q, tEx);
                    if (tEx > ti) { tEnd1 = tEx; em1 = exm; } else { tEnd0 = tEx; em0 = exm; }
                    if (k.useBase) fb |= range_test(k.protBase, k.span, tEnd0, tEnd1, ti, em0, em1, xm);
                    if (k.useShape) fs |= range_test(k.protShape, k.span, tEnd0, tEnd1, ti, em0, em1, xm);
                }
            }
            classify_point<N, OVERLAP>(e, k, ti, fb, fs);
        }
    }
}

// findErrors<N>(sdf) and findErrors<Combiner,N>(sdf, shape) for one texel in one walk over its 8 neighbours
// (MSDFErrorCorrection.cpp:333-457).  The reference short-circuits its ||-chain; the tests are pure, so evaluating all of
// them and OR-ing gives the same flag.  Outcomes that need an exact distance are queued as Candidates.  The neighbour
// loops stay rolled: the body is large and divergent, and a compact loop reconverges the warp at every neighbour.
template <int N, bool OVERLAP>
__device__ __noinline__ void find_errors_texel(const EcCtx<N> &e, TexelState &k, int x, int y, double hSpan, double vSpan, double dSpan) {
    const int w = e.f.w, h = e.f.h;
    const float *c = e.f.at(x, y);
    const float a[3] = { c[0], c[1], c[2] };
    float am = median3f(a[0], a[1], a[2]);
    k.sdfCoord = mk(x+.5, y+.5);
    k.shapeCoord = mk((x+.5)/e.g->sx-e.g->tx, (y+.5)/e.g->sy-e.g->ty);
    k.msd = c;
    // horizontally / vertically adjacent texels: hasLinearArtifact (:333-345)
    #pragma unroll 1
    for (int n = 0; n < 4; ++n) {
        int dx = n == 0 ? -1 : n == 2 ? 1 : 0;
        int dy = n == 1 ? -1 : n == 3 ? 1 : 0;
        int nx = x+dx, ny = y+dy;
        if (nx < 0 || ny < 0 || nx >= w || ny >= h)
            continue;
        const float *bp = e.f.at(nx, ny);
        const float b[3] = { bp[0], bp[1], bp[2] };
        float bm = median3f(b[0], b[1], b[2]);
        if (!(fabsf(am-.5f) >= fabsf(bm-.5f)))
            continue;
        k.direction = mk(dx, dy);
        k.span = dx ? hSpan : vSpan;
        linear_artifact_inner<N, OVERLAP>(e, k, am, bm, a, b, a[1]-a[0], b[1]-b[0]);
        linear_artifact_inner<N, OVERLAP>(e, k, am, bm, a, b, a[2]-a[1], b[2]-b[1]);
        linear_artifact_inner<N, OVERLAP>(e, k, am, bm, a, b, a[0]-a[2], b[0]-b[2]);
    }
    // diagonally adjacent texels: hasDiagonalArtifact (:349-381) with b = horizontal, c = vertical neighbour
    k.span = dSpan;
    #pragma unroll 1
    for (int n = 0; n < 4; ++n) {
        int dx = (n&1) ? 1 : -1;
        int dy = (n&2) ? 1 : -1;
        int nx = x+dx, ny = y+dy;
        if (nx < 0 || ny < 0 || nx >= w || ny >= h)
            continue;
        const float *dp = e.f.at(nx, ny);
        const float d[3] = { dp[0], dp[1], dp[2] };
        float dm = median3f(d[0], d[1], d[2]);
        if (!(fabsf(am-.5f) >= fabsf(dm-.5f)))
            continue;
        const float *bp = e.f.at(nx, y), *cp = e.f.at(x, ny);
        const float b[3] = { bp[0], bp[1], bp[2] };
        const float cc[3] = { cp[0], cp[1], cp[2] };
        k.direction = mk(dx, dy);
        float abc[3] = { a[0]-b[0]-cc[0], a[1]-b[1]-cc[1], a[2]-b[2]-cc[2] };
        float l[3] = { -a[0]-abc[0], -a[1]-abc[1], -a[2]-abc[2] };
        float q[3] = { d[0]+abc[0], d[1]+abc[1], d[2]+abc[2] };
        diagonal_artifact_inner<N, OVERLAP>(e, k, am, dm, a, l, q, a[1]-a[0], b[1]-b[0]+cc[1]-cc[0], d[1]-d[0], 0, 1);
        diagonal_artifact_inner<N, OVERLAP>(e, k, am, dm, a, l, q, a[2]-a[1], b[2]-b[1]+cc[2]-cc[1], d[2]-d[1], 1, 2);
        diagonal_artifact_inner<N, OVERLAP>(e, k, am, dm, a, l, q, a[0]-a[2], b[0]-b[2]+cc[0]-cc[2], d[0]-d[2], 2, 0);
    }
}

// edgeBetweenTexelsChannel (MSDFErrorCorrection.cpp:154-169)
__device__ __forceinline__ bool edge_between_channel(float a0, float a1, float a2, float b0, float b1, float b2, int ch) {
    float ac = ch == 0 ? a0 : ch == 1 ? a1 : a2;
    float bc = ch == 0 ? b0 : ch == 1 ? b1 : b2;
    double num = ac-.5;
    float den = ac-bc;
    // t = num/den is in (0,1) only if both are non-zero with equal signs: otherwise skip the division (same outcome)
    if (!((num > 0 && den > 0) || (num < 0 && den < 0)))
        return false;
    double t = num/den;
    if (t > 0 && t < 1) {
        float c0 = mixf(a0, b0, t), c1 = mixf(a1, b1, t), c2 = mixf(a2, b2, t);
        float cc = ch == 0 ? c0 : ch == 1 ? c1 : c2;
        return median3f(c0, c1, c2) == cc;
    }
    return false;
}

// protectEdges restated as a gather: is texel (x,y) protected by any of the (up to 8) texel pairs it belongs to?
// (MSDFErrorCorrection.cpp:189-250; each pair is evaluated with the same (first, second) roles as the reference's scan:
// first = left / bottom / the lower texel of a diagonal.)
template <int N>
__device__ __forceinline__ bool protect_edges_texel(const FieldView<N> &f, int x, int y, float hRadius, float vRadius, float dRadius) {
    const float *c = f.at(x, y);
    const float c0 = c[0], c1 = c[1], c2 = c[2];
    float cm = median3f(c0, c1, c2);
    const int w = f.w, h = f.h;
    bool prot = false;
    #pragma unroll 1
    for (int n = 0; n < 8; ++n) {
        // neighbour offsets; the texel with the smaller y (then smaller x) is the pair's first element
        int dx = n == 0 ? 1 : n == 1 ? -1 : n == 2 || n == 3 ? 0 : (n == 4 || n == 7) ? 1 : -1;
        int dy = n < 2 ? 0 : n == 2 ? 1 : n == 3 ? -1 : (n == 4 || n == 6) ? 1 : -1;
        int nx = x+dx, ny = y+dy;
        if (nx < 0 || ny < 0 || nx >= w || ny >= h)
            continue;
        float radius = n < 2 ? hRadius : n < 4 ? vRadius : dRadius;
        bool cIsFirst = dy > 0 || (dy == 0 && dx > 0);
        const float *o = f.at(nx, ny);
        const float o0 = o[0], o1 = o[1], o2 = o[2];
        float om = median3f(o0, o1, o2);
        float sum = cIsFirst ? fabsf(cm-.5f)+fabsf(om-.5f) : fabsf(om-.5f)+fabsf(cm-.5f);
        if (!(sum < radius))
            continue;
        float a0 = cIsFirst ? c0 : o0, a1 = cIsFirst ? c1 : o1, a2 = cIsFirst ? c2 : o2;
        float b0 = cIsFirst ? o0 : c0, b1 = cIsFirst ? o1 : c1, b2 = cIsFirst ? o2 : c2;
        // protectExtremeChannels: a channel matters only if it is not the texel's median
        if (c0 != cm && edge_between_channel(a0, a1, a2, b0, b1, b2, 0)) prot = true;
        if (c1 != cm && edge_between_channel(a0, a1, a2, b0, b1, b2, 1)) prot = true;
        if (c2 != cm && edge_between_channel(a0, a1, a2, b0, b1, b2, 2)) prot = true;
        if (prot)
            break;
    }
    return prot;
}

#define K2_QUEUE_CAPACITY 2048

// Dynamic shared memory layout of k2 (bytes): records | output staging | candidate queue
#define K2_SMEM_BYTES (sizeof(SegRec)*SMEM_RECORDS+sizeof(float)*CTA_THREADS*4+sizeof(Candidate)*K2_QUEUE_CAPACITY)

template <int N, bool OVERLAP>
__global__ void __launch_bounds__(CTA_THREADS) k2_error_correction(KernelParams P) {
    extern __shared__ __align__(16) unsigned char k2_smem[];
    SegRec *s_recs = reinterpret_cast<SegRec *>(k2_smem);
    float *s_out = reinterpret_cast<float *>(k2_smem+sizeof(SegRec)*SMEM_RECORDS);
    Candidate *s_queue = reinterpret_cast<Candidate *>(k2_smem+sizeof(SegRec)*SMEM_RECORDS+sizeof(float)*CTA_THREADS*4);
    __shared__ GlyphDesc s_g;
    __shared__ int s_tile;
    __shared__ int s_queueCount;
    __shared__ int s_err[CTA_THREADS];

    double localStore[B200_LOCAL_CONTOURS];
    int stagedGlyph = -1;
    const SegRec *recs = nullptr;
    const bool useBase = P.shapeless || P.ecDcm == 0 || (P.ecDcm == 1 && P.ecMode != 3);
    const bool useShape = !P.shapeless && (P.ecDcm == 1 || P.ecDcm == 2);
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
            if (useShape || needCorners)
                recs = stage_glyph(P, s_g, s_recs);
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const int w = s_g.w, h = s_g.h;
        const unsigned total = unsigned(w)*unsigned(h);
        const int nRec = (useShape || needCorners) ? int(s_g.s1-s_g.s0) : 0;
        unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;

        EcCtx<N> e;
        e.f.base = P.field+s_g.outOffset;
        e.f.w = w; e.f.h = h; e.f.flip = s_g.flip;
        e.g = &s_g;
        e.recs = recs;
        e.nRec = nRec;
        e.minImproveRatio = P.minImproveRatio;
        e.queue = s_queue;
        e.queueCount = &s_queueCount;
        e.queueCapacity = K2_QUEUE_CAPACITY;
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
            int x = 0, y = 0, flags = 0;
            bool errShape = false;
            s_err[threadIdx.x] = 0;
            if (threadIdx.x == 0)
                s_queueCount = 0;
            __syncthreads();
            // ---- phase A: protection + every test that needs only the field; queue the exact-distance candidates
            if (i < tileEnd) {
                x = int(i%unsigned(w));
                int yOut = int(i/unsigned(w));
                y = s_g.flip ? h-1-yOut : yOut;
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
                TexelState k;
                k.useBase = useBase;
                k.protBase = (flags&EC_FLAG_PROTECTED) != 0;
                k.useShape = useShape;
                // protectAll() runs between the two passes when distanceCheckMode == CHECK_DISTANCE_AT_EDGE
                k.protShape = k.protBase || (useBase && P.ecDcm == 1);
                k.errBase = k.errShape = false;
                k.texel = int(threadIdx.x);
                find_errors_texel<N, OVERLAP>(e, k, x, y, hSpan, vSpan, dSpan);
                if (k.errBase)
                    flags |= EC_FLAG_ERROR;
                if (k.protShape && useShape)
                    flags |= EC_FLAG_PROTECTED;
                errShape = k.errShape;
            }
            __syncthreads();
            // ---- phase B: exact distances, one candidate per thread
            int nCand = s_queueCount < K2_QUEUE_CAPACITY ? s_queueCount : K2_QUEUE_CAPACITY;
            for (int c = threadIdx.x; c < nCand; c += CTA_THREADS) {
                Candidate cand = s_queue[c];
                if (candidate_is_artifact<N, OVERLAP>(e, cand))
                    s_err[cand.texel] = 1;
            }
            __syncthreads();
            // ---- phase C: apply
            if (i < tileEnd) {
                if (errShape || s_err[threadIdx.x])
                    flags |= EC_FLAG_ERROR;
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
        }
    }
}

}  // namespace b200
