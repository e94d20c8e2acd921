// error_correction.cuh — the MSDF error-correction pass as two kernels.
//
// Reference (paths relative to /root/reference/core/): msdfErrorCorrectionInner<N> msdf-error-correction.cpp:12-48 drives
// MSDFErrorCorrection.cpp: protectCorners :121-151, protectEdges :189-250, protectAll :252-258, findErrors<N>(sdf)
// :383-410, findErrors<Combiner,N>(sdf, shape) :412-457 (ShapeDistanceChecker :50-102), apply :459-479.
//
// The reference makes several sweeps over the bitmap that OR flags into a stencil and finally rewrites flagged texels.
// All detection sweeps read the UNMODIFIED field and each texel's flags depend only on its 3x3 neighbourhood, so here one
// thread owns one texel and derives its final flags in a single walk over its 8 neighbours (scatter-ORs become gathers,
// the base pass and the shape pass share the walk), reading the distance kernel's output buffer and writing a second one:
//
//   k2_classify        per texel: protection, every test decidable from the field alone (the out-of-line ones compacted per
//                      warp and run 32 at a time), apply; the rare sub-texel points
//                      whose verdict needs the exact shape distance (about 0.006 per texel in the default configuration)
//                      are written to per-tile candidate slots.
//   k3_exact_distance  one warp per tile: stages the glyph's segment records, evaluates the exact perpendicular distance
//                      for the tile's candidates with all lanes busy, and patches the texels found to be artifacts.
//
// Why two kernels: an exact distance costs as much as a texel of the distance kernel (a walk over every segment, a
// latency-bound fp64 chain), while the classification is a few hundred instructions; inside one kernel a single candidate
// stalls its whole warp (or CTA) for the length of a distance evaluation and the heavy code inflates the register count of
// the light path.  Split, k2 runs at high occupancy and k3 evaluates candidates densely.
#pragma once

#include "kernels.cuh"

namespace b200 {

#define EC_FLAG_ERROR 1
#define EC_FLAG_PROTECTED 2
#define EC_FLAG_RECHECK 0x80     // internal: candidate slots overflowed, k3 re-walks this texel with in-place evaluation
#define ARTIFACT_T_EPSILON .01
#define PROTECTION_RADIUS_TOLERANCE 1.001
#define CLS_CANDIDATE 1
#define CLS_ARTIFACT 2
#define EC_WARPS (CTA_THREADS/32)

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

// A sub-texel point whose exact shape distance decides whether its texel is an artifact (ShapeDistanceChecker).  The
// producer finishes everything that needs only the field (interpolation, before/after medians).
struct __align__(16) Candidate {
    double qx, qy;          // point in shape coordinates
    float oldPSD, newPSD;   // interpolated median before / after the hypothetical correction
    unsigned texel;         // texel index inside the glyph, output orientation
    int pad;
};

template <int N>
struct EcCtx {
    FieldView<N> f;
    const GlyphDesc *g;
    double minImproveRatio;
    // SINK 0 (k2): candidate slots of the current tile
    Candidate *slots;
    int *slotCount;
    int slotCap;
    // SINK 1 (k3): in-place exact evaluation
    const SegRec *recs;
    int nRec;
    ContourStore store;
};

// Per-texel classification state shared by the base pass (findErrors<N>(sdf)) and the shape pass (findErrors(sdf, shape)):
// both walk the same neighbour tests and differ only in the protected flag fed to rangeTest and in what a CANDIDATE means.
struct TexelState {
    bool useBase, protBase;      // BaseArtifactClassifier pass enabled / its protectedFlag
    bool useShape, protShape;    // ShapeDistanceChecker pass enabled / its protectedFlag
    bool errBase, errShape;      // decided without an exact distance
    bool overflow;               // a candidate could not be queued
    double span;
    d2 direction;
    d2 shapeCoord, sdfCoord;
    const float *msd;
    unsigned texel;
    // k2 (SINK 0): the neighbour walk only ENUMERATES the tests that survive the prefilters (owner texel, kind, neighbour,
    // channel pair packed in 16 bits) into a queue shared by the tile; the CTA then runs them 256 at a time
    unsigned short *queue;
    int *queueCount;
    int tileTexel;
};

// Tile-wide queue of out-of-line tests: texel index inside the tile (10 bits), kind (bit 10), neighbour (bits 11-12), channel pair
// (bits 13-14).  A full queue makes the enqueuing lane run its test in place.
#define K2_QUEUE 6144
__device__ __forceinline__ bool enqueue_test(TexelState &k, int kind, int n, int pr) {
    int at = atomicAdd(k.queueCount, 1);
    if (at >= K2_QUEUE)
        return false;
    k.queue[at] = (unsigned short) (k.tileTexel|(kind<<10)|(n<<11)|(pr<<13));
    return true;
}

// BaseArtifactClassifier::rangeTest (:30-41)
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

// ArtifactClassifier::evaluate's comparison once the exact distance is known (:79)
template <bool OVERLAP>
__device__ __forceinline__ bool candidate_is_artifact(const SegRec *recs, int nRec, ContourStore store, const GlyphDesc &g, double minImproveRatio, const Candidate &c) {
    double dist;
    shape_distance<2, OVERLAP>(recs, nRec, mk(c.qx, c.qy), store, &dist);     // ShapeDistanceFinder<Combiner<PerpendicularDistanceSelector>>
    float refPSD = float(g.dmScale*(dist+g.dmTranslate));
    return minImproveRatio*fabsf(c.newPSD-refPSD) < double(fabsf(c.oldPSD-refPSD));
}

// The part of ArtifactClassifier::evaluate (:59-79) that precedes the distance evaluation — rare, out of line.
template <int N, int SINK, bool OVERLAP>
__device__ __noinline__ void handle_candidate(const EcCtx<N> &e, TexelState &k, double t) {
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
    if (SINK == 0) {
        int slot = atomicAdd(e.slotCount, 1);
        if (slot < e.slotCap)
            e.slots[slot] = c;
        else
            k.overflow = true;
    } else if (candidate_is_artifact<OVERLAP>(e.recs, e.nRec, e.store, *e.g, e.minImproveRatio, c))
        k.errShape = true;
}

// Classifier::evaluate for both passes at interpolation ratio t with the two passes' range flags.
template <int N, int SINK, bool OVERLAP>
__device__ __forceinline__ void classify_point(const EcCtx<N> &e, TexelState &k, double t, int flagsBase, int flagsShape) {
    if (k.useBase && (flagsBase&CLS_ARTIFACT))
        k.errBase = true;
    if (k.useShape && (flagsShape&CLS_CANDIDATE)) {
        if (flagsShape&CLS_ARTIFACT)
            k.errShape = true;
        else if (!(k.errBase || k.errShape))    // once an error, the exact distance cannot change the outcome
            handle_candidate<N, SINK, OVERLAP>(e, k, t);
    }
}

__device__ __forceinline__ float interp_median_lin(const float *a, const float *b, double t) {
    return median3f(mixf(a[0], b[0], t), mixf(a[1], b[1], t), mixf(a[2], b[2], t));
}
__device__ __forceinline__ float interp_median_quad(const float *a, const float *l, const float *q, double t) {
    return float(median3(t*(t*q[0]+l[0])+a[0], t*(t*q[1]+l[1])+a[1], t*(t*q[2]+l[2])+a[2]));
}

// hasLinearArtifactInner (:279-289)
template <int N, int SINK, bool OVERLAP>
__device__ __noinline__ void linear_artifact_inner(const EcCtx<N> &e, TexelState &k, float am, float bm, const float *a, const float *b, float dA, float dB) {
    double t = double(dA)/(dA-dB);
    if (t > ARTIFACT_T_EPSILON && t < 1-ARTIFACT_T_EPSILON) {
        float xm = interp_median_lin(a, b, t);
        int fb = k.useBase ? range_test(k.protBase, k.span, 0, 1, t, am, bm, xm) : 0;
        int fs = k.useShape ? range_test(k.protShape, k.span, 0, 1, t, am, bm, xm) : 0;
        classify_point<N, SINK, OVERLAP>(e, k, t, fb, fs);
    }
}

// Local extremes of the three channels along one diagonal (tEx of hasDiagonalArtifact :364-369) and the interpolated medians
// there, computed on first use and shared by the three channel-pair tests of that diagonal.
struct DiagonalExtremes {
    double tEx[3];
    float median[3];
    int have;       // bit k: tEx[k] (and, if inside (0,1), median[k]) are valid
};

// hasDiagonalArtifactInner (:293-329)
template <int N, int SINK, bool OVERLAP>
__device__ __noinline__ void diagonal_artifact_inner(const EcCtx<N> &e, TexelState &k, float am, float dm, const float *a, const float *l, const float *q,
                                                        float dA, float dBC, float dD, int ex0, int ex1, DiagonalExtremes &x) {
    const double qa = dD-dBC+dA, qb = dBC-dA-dA, qc = dA;
    Roots r = solve_quadratic(qa, qb, qc);
    for (int i = 0; i < r.n; ++i) {
        double ti = i == 0 ? r.t0 : r.t1;
        if (ti > ARTIFACT_T_EPSILON && ti < 1-ARTIFACT_T_EPSILON) {
            float xm = interp_median_quad(a, l, q, ti);
            int fb = k.useBase ? range_test(k.protBase, k.span, 0, 1, ti, am, dm, xm) : 0;
            int fs = k.useShape ? range_test(k.protShape, k.span, 0, 1, ti, am, dm, xm) : 0;
            #pragma unroll 1
            for (int ex = 0; ex < 2; ++ex) {
                int ch = ex == 0 ? ex0 : ex1;
                if (!((x.have>>ch)&1)) {
                    // local extreme of channel ch along the diagonal: 2*q*t + l == 0
                    double t = -.5*l[ch]/q[ch];
                    x.tEx[ch] = t;
                    if (t > 0 && t < 1)
                        x.median[ch] = interp_median_quad(a, l, q, t);
                    x.have |= 1<<ch;
                }
                double tEx = x.tEx[ch];
                if (tEx > 0 && tEx < 1) {
                    double tEnd0 = 0, tEnd1 = 1;
                    float em0 = am, em1 = dm;
                    float exm = x.median[ch];
                    if (tEx > ti) { tEnd1 = tEx; em1 = exm; } else { tEnd0 = tEx; em0 = exm; }
                    if (k.useBase) fb |= range_test(k.protBase, k.span, tEnd0, tEnd1, ti, em0, em1, xm);
                    if (k.useShape) fs |= range_test(k.protShape, k.span, tEnd0, tEnd1, ti, em0, em1, xm);
                }
            }
            classify_point<N, SINK, OVERLAP>(e, k, ti, fb, fs);
        }
    }
}

// findErrors<N>(sdf) and findErrors<Combiner,N>(sdf, shape) for one texel in one walk over its 8 neighbours (:333-457).
// The reference short-circuits its ||-chain; the tests are pure, so evaluating all of them and OR-ing gives the same flag.
// The neighbour loops stay rolled: the body is large and divergent, a compact loop reconverges the warp at every neighbour.
template <int N, int SINK, bool OVERLAP>
__device__ __forceinline__ void find_errors_texel(const EcCtx<N> &e, TexelState &k, int x, int y, double hSpan, double vSpan, double dSpan) {
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
        // channel pairs (g-r, b-g, r-b); the loop stays rolled to keep the kernel's instruction footprint small
        #pragma unroll 1
        for (int pr = 0; pr < 3; ++pr) {
            float a0 = pr == 0 ? a[0] : pr == 1 ? a[1] : a[2], a1 = pr == 0 ? a[1] : pr == 1 ? a[2] : a[0];
            float b0 = pr == 0 ? b[0] : pr == 1 ? b[1] : b[2], b1 = pr == 0 ? b[1] : pr == 1 ? b[2] : b[0];
            // t = dA/(dA-dB) lies strictly inside (0,1) only if dA and dB have strictly opposite signs; in every other case
            // the reference's range test (t > .01 && t < .99) fails as well, so the out-of-line test with its fp64 division
            // is skipped without changing a result.
            float dA = a1-a0, dB = b1-b0;
            if ((dA > 0 && dB < 0) || (dA < 0 && dB > 0))
                if (SINK != 0 || !enqueue_test(k, 0, n, pr))
                    linear_artifact_inner<N, SINK, OVERLAP>(e, k, am, bm, a, b, dA, dB);
        }
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
        DiagonalExtremes extremes;
        extremes.have = 0;
        #pragma unroll 1
        for (int pr = 0; pr < 3; ++pr) {
            const int i0 = pr, i1 = pr == 2 ? 0 : pr+1;
            #define B200_PICK(v, i) ((i) == 0 ? v[0] : (i) == 1 ? v[1] : v[2])
            float a0 = B200_PICK(a, i0), a1 = B200_PICK(a, i1), b0 = B200_PICK(b, i0), b1 = B200_PICK(b, i1);
            float c0 = B200_PICK(cc, i0), c1 = B200_PICK(cc, i1), d0 = B200_PICK(d, i0), d1 = B200_PICK(d, i1);
            #undef B200_PICK
            // Two channels are equal where f(t) = qa t^2 + qb t + qc vanishes.  f has a zero strictly inside (0,1) only if
            // it changes sign between t = 0 and t = 1 or has its vertex inside (0,1); otherwise the reference's roots all
            // fall outside its (.01, .99) window, so the out-of-line test with the square root and divisions of
            // solveQuadratic is skipped without changing a result.
            float dA = a1-a0, dBC = b1-b0+c1-c0, dD = d1-d0;
            const double qa = dD-dBC+dA, qb = dBC-dA-dA, qc = dA;
            const double f1 = qa+qb+qc;
            const bool signChange = (qc > 0 && f1 < 0) || (qc < 0 && f1 > 0);
            // without a sign change both roots must lie inside [0,1): vertex inside, f(0) not on the far side of the
            // parabola's opening (f(0) == 0 is the frequent "two channels equal at the texel" case: the other root may be
            // inside), and a non-negative discriminant (the very expression solveQuadratic evaluates)
            const bool vertexInside = (qa > 0 && qb < 0 && -qb < 2*qa && qc >= 0) || (qa < 0 && qb > 0 && qb < -2*qa && qc <= 0);
            if (signChange || (vertexInside && qb*qb-4*qa*qc >= 0))
                if (SINK != 0 || !enqueue_test(k, 1, n, pr))
                    diagonal_artifact_inner<N, SINK, OVERLAP>(e, k, am, dm, a, l, q, dA, dBC, dD, i0, i1, extremes);
        }
    }
}

// One queued test: rebuilds exactly the operands find_errors_texel had at the enqueue site for texel
// (x, y), neighbour n and channel pair pr, and runs the out-of-line test.  k carries the owner's protection flags in and the
// verdict bits out.
template <int N>
__device__ __forceinline__ void run_queued_test(const EcCtx<N> &e, TexelState &k, int x, int y, int kind, int n, int pr, double hSpan, double vSpan, double dSpan) {
    const float *c = e.f.at(x, y);
    const float a[3] = { c[0], c[1], c[2] };
    const float am = median3f(a[0], a[1], a[2]);
    k.sdfCoord = mk(x+.5, y+.5);
    k.shapeCoord = mk((x+.5)/e.g->sx-e.g->tx, (y+.5)/e.g->sy-e.g->ty);
    k.msd = c;
    #define B200_PICK(v, i) ((i) == 0 ? v[0] : (i) == 1 ? v[1] : v[2])
    if (kind == 0) {
        const int dx = n == 0 ? -1 : n == 2 ? 1 : 0;
        const int dy = n == 1 ? -1 : n == 3 ? 1 : 0;
        const float *bp = e.f.at(x+dx, y+dy);
        const float b[3] = { bp[0], bp[1], bp[2] };
        const float bm = median3f(b[0], b[1], b[2]);
        k.direction = mk(dx, dy);
        k.span = dx ? hSpan : vSpan;
        const int i0 = pr, i1 = pr == 2 ? 0 : pr+1;
        const float dA = B200_PICK(a, i1)-B200_PICK(a, i0), dB = B200_PICK(b, i1)-B200_PICK(b, i0);
        linear_artifact_inner<N, 0, false>(e, k, am, bm, a, b, dA, dB);
    } else {
        const int dx = (n&1) ? 1 : -1;
        const int dy = (n&2) ? 1 : -1;
        const float *dp = e.f.at(x+dx, y+dy), *bp = e.f.at(x+dx, y), *cp = e.f.at(x, y+dy);
        const float d[3] = { dp[0], dp[1], dp[2] };
        const float b[3] = { bp[0], bp[1], bp[2] };
        const float cc[3] = { cp[0], cp[1], cp[2] };
        const float dm = median3f(d[0], d[1], d[2]);
        k.direction = mk(dx, dy);
        k.span = dSpan;
        float abc[3] = { a[0]-b[0]-cc[0], a[1]-b[1]-cc[1], a[2]-b[2]-cc[2] };
        float l[3] = { -a[0]-abc[0], -a[1]-abc[1], -a[2]-abc[2] };
        float q[3] = { d[0]+abc[0], d[1]+abc[1], d[2]+abc[2] };
        const int i0 = pr, i1 = pr == 2 ? 0 : pr+1;
        const float a0 = B200_PICK(a, i0), a1 = B200_PICK(a, i1), b0 = B200_PICK(b, i0), b1 = B200_PICK(b, i1);
        const float c0 = B200_PICK(cc, i0), c1 = B200_PICK(cc, i1), d0 = B200_PICK(d, i0), d1 = B200_PICK(d, i1);
        const float dA = a1-a0, dBC = b1-b0+c1-c0, dD = d1-d0;
        DiagonalExtremes extremes;
        extremes.have = 0;
        diagonal_artifact_inner<N, 0, false>(e, k, am, dm, a, l, q, dA, dBC, dD, i0, i1, extremes);
    }
    #undef B200_PICK
}

// edgeBetweenTexelsChannel (:154-169)
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
// (:189-250; each pair is evaluated with the same (first, second) roles as the reference's scan: first = left / bottom /
// the lower texel of a diagonal.)
template <int N>
__device__ __forceinline__ bool protect_edges_texel(const FieldView<N> &f, int x, int y, float hRadius, float vRadius, float dRadius) {
    const float *c = f.at(x, y);
    const float c0 = c[0], c1 = c[1], c2 = c[2];
    float cm = median3f(c0, c1, c2);
    const int w = f.w, h = f.h;
    bool prot = false;
    #pragma unroll 1
    for (int n = 0; n < 8; ++n) {
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

// protectCorners (:121-151) for one segment: texel (l,b) of the colour corner at its start point, or a sentinel.
__device__ __forceinline__ int2 corner_of(const RawShapes &rs, const GlyphDesc &g, unsigned j) {
    unsigned s = g.s0+j;
    unsigned lo = g.c0, hi = g.c1;
    while (hi-lo > 1) {
        unsigned mid = (lo+hi)>>1;
        if (rs.cbegin[mid] <= s) lo = mid; else hi = mid;
    }
    unsigned b = rs.cbegin[lo], e = rs.cbegin[lo+1];
    unsigned prev = s == b ? e-1 : s-1;
    int color = rs.colors ? rs.colors[s] : 7, prevColor = rs.colors ? rs.colors[prev] : 7;
    int common = prevColor&color;
    int2 r = make_int2(INT_MIN/2, INT_MIN/2);
    if (!(common&(common-1))) {
        RawSeg sc = load_raw(rs, s);
        d2 c0 = seg_point(sc, 0);
        double px = g.sx*(c0.x+g.tx), py = g.sy*(c0.y+g.ty);    // Projection::project
        double fl = floor(px-.5), fb = floor(py-.5);
        if (fl > -1e9 && fl < 1e9 && fb > -1e9 && fb < 1e9)
            r = make_int2(int(fl), int(fb));
    }
    return r;
}

struct EcSpans {
    float hRadius, vRadius, dRadius;
    double hSpan, vSpan, dSpan;
};

__device__ __forceinline__ EcSpans ec_spans(const GlyphDesc &g, double minDeviationRatio) {
    // distanceMapping(Delta(1)) = dmScale; unprojectVector(v) = v/scale
    const double dm1 = g.dmScale*1;
    const double hLen = len(mk(dm1/g.sx, 0/g.sy));
    const double vLen = len(mk(0/g.sx, dm1/g.sy));
    const double dLen = len(mk(dm1/g.sx, dm1/g.sy));
    EcSpans s;
    s.hRadius = float(PROTECTION_RADIUS_TOLERANCE*hLen);
    s.vRadius = float(PROTECTION_RADIUS_TOLERANCE*vLen);
    s.dRadius = float(PROTECTION_RADIUS_TOLERANCE*dLen);
    s.hSpan = minDeviationRatio*hLen; s.vSpan = minDeviationRatio*vLen; s.dSpan = minDeviationRatio*dLen;
    return s;
}

// ---------------------------------------------------------------------------------------------------------------------
// k2: classification + apply, three phases per 1024-texel tile (two CTA barriers):
//   1  every warp pulls 32-texel chunks (8x4 blocks where the tile allows): protection flags per texel into shared memory, and
//      the neighbour walk of both findErrors passes, which only ENUMERATES the out-of-line tests (the ones with fp64 divisions,
//      square roots, interpolations) that survive the exact prefilters into a tile-wide queue — about 0.4 per texel;
//   2  the whole CTA runs the queued tests, 256 at a time with every lane busy, and ORs the verdicts into the owners' flags;
//   3  apply (error texels <- median), stencil and the float4 write-back.
#ifndef K2_MIN_CTAS
#define K2_MIN_CTAS 3
#endif
template <int N>
__global__ void __launch_bounds__(CTA_THREADS, K2_MIN_CTAS) k2_classify(KernelParams P) {
    __shared__ int2 s_corner[SMEM_RECORDS];
    __shared__ __align__(16) float s_out[EC_WARPS][32*4];
    __shared__ GlyphDesc s_g;
    __shared__ int s_tile;
    __shared__ int s_chunk;
    __shared__ unsigned short s_queue[K2_QUEUE];
    __shared__ int s_queueCount;
    __shared__ int s_state[TILE_PIXELS];         // per texel of the tile: bit 0 protBase, 1 protShape; 2 errBase, 3 errShape, 4 overflow

    const int warp = threadIdx.x>>5, lane = threadIdx.x&31;
    int stagedGlyph = -1;
    const bool useBase = P.shapeless || P.ecDcm == 0 || (P.ecDcm == 1 && P.ecMode != 3);
    const bool useShape = !P.shapeless && (P.ecDcm == 1 || P.ecDcm == 2);
    const bool needCorners = !P.shapeless && P.ecMode == 2;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_tile = P.tileBegin+atomicAdd(&P.counters[P.counterIndex], 1);
            s_chunk = 0;
            s_queueCount = 0;
        }
        __syncthreads();
        int tile = s_tile;
        if (tile >= P.tileEnd)
            break;
        TileDesc td = P.tiles[tile];
        if ((int) td.glyph != stagedGlyph) {
            if (threadIdx.x == 0)
                s_g = P.glyphs[td.glyph];
            __syncthreads();
            if (needCorners && s_g.recBase < 0)
                for (unsigned j = threadIdx.x; j < s_g.s1-s_g.s0; j += blockDim.x)
                    s_corner[j] = corner_of(P.raw, s_g, j);
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const int w = s_g.w, h = s_g.h;
        const unsigned total = unsigned(w)*unsigned(h);
        const int nRec = needCorners ? int(s_g.s1-s_g.s0) : 0;
        const SegRec *bigRecs = s_g.recBase >= 0 ? P.bigRecs+s_g.recBase : nullptr;
        unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
        const bool blocked = (w&7) == 0 && w <= 256 && td.pix0%unsigned(w) == 0 && (tileEnd-td.pix0)%(4u*unsigned(w)) == 0;

        EcCtx<N> e;
        e.f.base = P.field+s_g.outOffset;
        e.f.w = w; e.f.h = h; e.f.flip = s_g.flip;
        e.g = &s_g;
        e.minImproveRatio = P.minImproveRatio;
        e.slots = P.candidates+size_t(tile)*P.candCap;
        e.slotCount = P.candCount+tile;
        e.slotCap = P.candCap;
        e.recs = nullptr; e.nRec = 0;
        const EcSpans sp = ec_spans(s_g, P.minDeviationRatio);

        // ---- phase 0: protectCorners (:121-151) as a scatter — every colour corner marks its 2x2 texels (bit 0 of the state)
        for (unsigned t = threadIdx.x; t < TILE_PIXELS; t += blockDim.x)
            s_state[t] = 0;
        __syncthreads();
        if (P.ecMode == 2 && needCorners) {
            for (int j = threadIdx.x; j < nRec; j += blockDim.x) {
                int cl, cb;
                if (bigRecs) { cl = bigRecs[j].cornerL; cb = bigRecs[j].cornerB; }
                else { int2 c = s_corner[j]; cl = c.x; cb = c.y; }
                #pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const int x = cl+(d&1), y = cb+(d>>1);
                    if (x >= 0 && y >= 0 && x < w && y < h) {
                        const unsigned i = unsigned(s_g.flip ? h-1-y : y)*unsigned(w)+unsigned(x);
                        if (i >= td.pix0 && i < tileEnd)
                            atomicOr(&s_state[i-td.pix0], 1);
                    }
                }
            }
            __syncthreads();
        }

        // ---- phase 1: protection + enumeration
        for (;;) {
            int chunk = 0;
            if (lane == 0)
                chunk = atomicAdd(&s_chunk, 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            unsigned chunkBase = td.pix0+unsigned(chunk)*32u;
            if (chunkBase >= tileEnd)
                break;
            unsigned i = chunkBase+lane;
            // 8x4 texel blocks instead of 32x1 strips when the tile is a whole number of 4-row groups: lanes of a warp then
            // share most of their 3x3 neighbourhoods (cache lines) and take more similar branches
            if (blocked) {
                const unsigned blocksAcross = unsigned(w)>>3;
                unsigned bx = unsigned(chunk)%blocksAcross, by = unsigned(chunk)/blocksAcross;
                i = td.pix0+by*4u*unsigned(w)+bx*8u+unsigned(lane>>3)*unsigned(w)+unsigned(lane&7);
            }
            if (i < tileEnd) {
                const int x = int(i%unsigned(w));
                const int yOut = int(i/unsigned(w));
                const int y = s_g.flip ? h-1-yOut : yOut;
                bool prot = (s_state[i-td.pix0]&1) != 0;
                // msdfErrorCorrectionInner: protection stage
                if (P.ecMode == 2) {
                    if (!prot)
                        prot = protect_edges_texel<N>(e.f, x, y, sp.hRadius, sp.vRadius, sp.dRadius);
                } else if (P.ecMode == 3)
                    prot = true;
                // detection stage: both passes in one walk
                TexelState k;
                k.useBase = useBase;
                k.protBase = prot;
                k.useShape = useShape;
                // protectAll() runs between the two passes when distanceCheckMode == CHECK_DISTANCE_AT_EDGE
                k.protShape = k.protBase || (useBase && P.ecDcm == 1);
                k.errBase = k.errShape = k.overflow = false;
                k.texel = i;
                k.queue = s_queue;
                k.queueCount = &s_queueCount;
                k.tileTexel = int(i-td.pix0);
                find_errors_texel<N, 0, false>(e, k, x, y, sp.hSpan, sp.vSpan, sp.dSpan);     // enumerates (runs a test in place only if the queue is full)
                s_state[i-td.pix0] = (k.protBase ? 1 : 0)|(k.protShape ? 2 : 0)|(k.errBase ? 4 : 0)|(k.errShape ? 8 : 0)|(k.overflow ? 16 : 0);
            }
        }
        __syncthreads();
        // ---- phase 2: the queued tests, dense
        {
            const int nQueued = s_queueCount < K2_QUEUE ? s_queueCount : K2_QUEUE;
            for (int q = threadIdx.x; q < nQueued; q += blockDim.x) {
                const unsigned item = s_queue[q];
                const int owner = int(item&1023u), kind = int((item>>10)&1u), n = int((item>>11)&3u), pr = int((item>>13)&3u);
                const unsigned io = td.pix0+unsigned(owner);
                const int ox = int(io%unsigned(w));
                const int oyOut = int(io/unsigned(w));
                const int oy = s_g.flip ? h-1-oyOut : oyOut;
                const int state = *(volatile int *) &s_state[owner];
                TexelState t;
                t.useBase = useBase; t.useShape = useShape;
                t.protBase = (state&1) != 0; t.protShape = (state&2) != 0;
                // an owner already found in error needs no exact-distance candidates (classify_point's short cut)
                const bool knownError = (state&12) != 0;
                t.errBase = knownError; t.errShape = false; t.overflow = false;
                t.texel = io;
                t.queue = nullptr; t.queueCount = nullptr; t.tileTexel = owner;
                run_queued_test<N>(e, t, ox, oy, kind, n, pr, sp.hSpan, sp.vSpan, sp.dSpan);
                const int bits = ((t.errBase && !knownError) ? 4 : 0)|(t.errShape ? 8 : 0)|(t.overflow ? 16 : 0);
                if (bits)
                    atomicOr(&s_state[owner], bits);
            }
        }
        __syncthreads();
        // ---- phase 3: apply (:459-479), stencil, coalesced write-back
        for (unsigned chunkBase = td.pix0+unsigned(warp)*32u; chunkBase < tileEnd; chunkBase += EC_WARPS*32u) {
            const unsigned i = chunkBase+lane;
            if (i < tileEnd) {
                const int x = int(i%unsigned(w));
                const int yOut = int(i/unsigned(w));
                const int y = s_g.flip ? h-1-yOut : yOut;
                const int state = s_state[i-td.pix0];
                int flags = 0;
                if (P.ecMode == 3 || (state&1))
                    flags |= EC_FLAG_PROTECTED;
                if (state&12)
                    flags |= EC_FLAG_ERROR;
                if ((state&2) && useShape)
                    flags |= EC_FLAG_PROTECTED;
                if ((state&16) && !(flags&EC_FLAG_ERROR))
                    flags |= EC_FLAG_RECHECK;
                const float *c = e.f.base+size_t(i)*N;
                float o0 = c[0], o1 = c[1], o2 = c[2];
                if (flags&EC_FLAG_ERROR) {
                    float m = median3f(o0, o1, o2);
                    o0 = m; o1 = m; o2 = m;
                }
                s_out[warp][lane*N+0] = o0;
                s_out[warp][lane*N+1] = o1;
                s_out[warp][lane*N+2] = o2;
                if (N == 4)
                    s_out[warp][lane*N+3] = c[3];
                int srow = s_g.shapeYDown ? h-1-y : y;
                P.stencil[s_g.texelOffset+size_t(srow)*w+x] = (uint8_t) flags;
            }
            __syncwarp();
            // 32 consecutive texels of the glyph block: float4 when aligned
            unsigned nPix = tileEnd-chunkBase < 32u ? tileEnd-chunkBase : 32u;
            int nFloats = int(nPix)*N;
            float *dst = P.out+s_g.outOffset+size_t(chunkBase)*N;
            if ((reinterpret_cast<uintptr_t>(dst)&15) == 0) {
                int n4 = nFloats>>2;
                const float4 *s4 = reinterpret_cast<const float4 *>(s_out[warp]);
                float4 *d4 = reinterpret_cast<float4 *>(dst);
                for (int q = lane; q < n4; q += 32)
                    d4[q] = s4[q];
                for (int q = (n4<<2)+lane; q < nFloats; q += 32)
                    dst[q] = s_out[warp][q];
            } else {
                for (int q = lane; q < nFloats; q += 32)
                    dst[q] = s_out[warp][q];
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k3: exact-distance verdicts.  One warp (a 32-thread CTA) per tile that has candidates: stage the glyph's records,
// evaluate 32 candidates at a time, patch artifacts (texel <- median, stencil |= ERROR).  Many small CTAs per SM keep the
// latency-bound distance evaluations overlapped.
#define K3_THREADS 32

template <int N>
__device__ __forceinline__ void patch_texel(const KernelParams &P, const GlyphDesc &g, unsigned i, int clearRecheck) {
    int w = g.w, h = g.h;
    int x = int(i%unsigned(w)), yOut = int(i/unsigned(w));
    int y = g.flip ? h-1-yOut : yOut;
    const float *c = P.field+g.outOffset+size_t(i)*N;
    float m = median3f(c[0], c[1], c[2]);
    float *o = P.out+g.outOffset+size_t(i)*N;
    o[0] = m; o[1] = m; o[2] = m;
    int srow = g.shapeYDown ? h-1-y : y;
    uint8_t *st = P.stencil+g.texelOffset+size_t(srow)*w+x;
    uint8_t f = *st;
    f |= EC_FLAG_ERROR;
    if (clearRecheck)
        f &= (uint8_t) ~EC_FLAG_RECHECK;
    *st = f;
}

template <int N, bool OVERLAP, bool BIG>
__global__ void __launch_bounds__(K3_THREADS) k3_exact_distance(KernelParams P) {
    __shared__ SegRec s_recs[BIG ? 1 : SMEM_RECORDS];
    __shared__ __align__(16) double s_points[BIG ? 2 : SMEM_RECORDS*8];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ GlyphDesc s_g;
    double localStore[B200_LOCAL_CONTOURS*StoreLayout<2>::PER_CONTOUR];
    const int lane = threadIdx.x;
    unsigned phase = 0;
    if (!BIG && lane == 0)
        mbar_init(&s_mbar);
    __syncwarp();
    int stagedGlyph = -1;
    for (;;) {
        int tile = 0;
        if (lane == 0)
            tile = P.tileBegin+atomicAdd(&P.counters[P.counterIndex], 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= P.tileEnd)
            break;
        int pushed = P.candCount[tile];
        if (pushed == 0)
            continue;
        TileDesc td = P.tiles[tile];
        if ((int) td.glyph != stagedGlyph) {
            __syncwarp();
            if (lane == 0)
                s_g = P.glyphs[td.glyph];
            __syncwarp();
            if (!BIG)
                stage_glyph(P, s_g, s_recs, s_points, &s_mbar, phase);
            stagedGlyph = (int) td.glyph;
            __syncwarp();
        }
        const SegRec *recs = BIG ? P.bigRecs+s_g.recBase : s_recs;
        const int nRec = int(s_g.s1-s_g.s0);
        ContourStore store;
        store.perContour = StoreLayout<2>::PER_CONTOUR;
        if (OVERLAP && s_g.nonEmptyContours > B200_LOCAL_CONTOURS) {
            store.base = P.scratch+(size_t(blockIdx.x)*K3_THREADS+threadIdx.x);
            store.stride = P.scratchStride;
        } else {
            store.base = localStore;
            store.stride = 1;
        }
        const int w = s_g.w, h = s_g.h;
        int nCand = pushed < P.candCap ? pushed : P.candCap;
        const Candidate *slots = P.candidates+size_t(tile)*P.candCap;
        for (int c = lane; c < nCand; c += K3_THREADS) {
            Candidate cand = slots[c];
            // skip candidates of texels that the field-only tests already flagged (their verdict cannot change)
            int x = int(cand.texel%unsigned(w)), yOut = int(cand.texel/unsigned(w));
            int y = s_g.flip ? h-1-yOut : yOut;
            int srow = s_g.shapeYDown ? h-1-y : y;
            if (P.stencil[s_g.texelOffset+size_t(srow)*w+x]&EC_FLAG_ERROR)
                continue;
            if (candidate_is_artifact<OVERLAP>(recs, nRec, store, s_g, P.minImproveRatio, cand))
                patch_texel<N>(P, s_g, cand.texel, 0);
        }
        if (pushed > P.candCap) {
            // slots overflowed: re-walk the texels that lost a candidate, evaluating exact distances in place
            __syncwarp();
            const unsigned total = unsigned(w)*unsigned(h);
            unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
            EcCtx<N> e;
            e.f.base = P.field+s_g.outOffset;
            e.f.w = w; e.f.h = h; e.f.flip = s_g.flip;
            e.g = &s_g;
            e.minImproveRatio = P.minImproveRatio;
            e.slots = nullptr; e.slotCount = nullptr; e.slotCap = 0;
            e.recs = recs; e.nRec = nRec; e.store = store;
            const EcSpans sp = ec_spans(s_g, P.minDeviationRatio);
            for (unsigned i = td.pix0+lane; i < tileEnd; i += K3_THREADS) {
                int x = int(i%unsigned(w)), yOut = int(i/unsigned(w));
                int y = s_g.flip ? h-1-yOut : yOut;
                int srow = s_g.shapeYDown ? h-1-y : y;
                uint8_t *st = P.stencil+s_g.texelOffset+size_t(srow)*w+x;
                int flags = *st;
                if (!(flags&EC_FLAG_RECHECK))
                    continue;
                TexelState k;
                k.useBase = false; k.protBase = false;
                k.useShape = true;
                k.protShape = (flags&EC_FLAG_PROTECTED) != 0;
                k.errBase = k.errShape = k.overflow = false;
                k.texel = i;
                if (!(flags&EC_FLAG_ERROR))
                    find_errors_texel<N, 1, OVERLAP>(e, k, x, y, sp.hSpan, sp.vSpan, sp.dSpan);
                if (k.errShape)
                    patch_texel<N>(P, s_g, i, 1);
                else
                    *st = (uint8_t) (flags&~EC_FLAG_RECHECK);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k4: fused 8-bit quantisation + atlas placement.  pixelFloatToByte (core/pixel-conversion.hpp:8-10) of every final float;
// bytes go either to the glyph's rectangle in the caller's (pinned, device-visible) atlas page or to a packed byte buffer.
struct ByteTarget {
    uint8_t *base;          // first byte of row 0 in the caller's orientation (or packed buffer position)
    long long rowStride;    // bytes
};

__device__ __forceinline__ uint8_t pixel_float_to_byte(float x) {
    float c = x >= 0.f && x <= 1.f ? x : float(x > 0.f);       // clamp()
    return (uint8_t) ~int(255.5f-255.f*c);
}

template <int N>
__global__ void __launch_bounds__(CTA_THREADS) k4_quantize(const float *src, const GlyphDesc *glyphs, const TileDesc *tiles, int nTiles, const ByteTarget *targets) {
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        TileDesc td = tiles[tile];
        GlyphDesc g = glyphs[td.glyph];
        ByteTarget t = targets[td.glyph];
        const unsigned total = unsigned(g.w)*unsigned(g.h);
        unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
        for (unsigned i = td.pix0+threadIdx.x; i < tileEnd; i += CTA_THREADS) {
            unsigned x = i%unsigned(g.w), row = i/unsigned(g.w);     // row in the caller's row order
            const float *f = src+g.outOffset+size_t(i)*N;
            uint8_t *d = t.base+t.rowStride*(long long) row+(long long) x*N;
            #pragma unroll
            for (int k = 0; k < N; ++k)
                d[k] = pixel_float_to_byte(f[k]);
        }
    }
}

}  // namespace b200
