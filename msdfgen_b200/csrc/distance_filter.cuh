// distance_filter.cuh — the per-texel walk of k1_glyph_tiles: an fp32 FILTER in front of the exact fp64 evaluation.
//
// What the reference computes per texel, contour and channel (core/edge-selectors.cpp:81-117,129-164,174-227) is
//   * the minimum SignedDistance over the channel's edges (strict <, so the first edge in evaluation order wins exact ties), and
//   * the smallest negative / positive perpendicular distance among the edges' end-point candidates.
// Every edge takes part in it, but only very few can decide it.  The walk therefore runs in three passes per contour (per
// shape for the simple combiner); passes 1 and 2 are single precision and only ever DROP work that provably cannot change the
// reference's result, pass 3 is the unchanged fp64 evaluation (distance_core.cuh) of what is left:
//
//   pass 1  U[k]  = min over the channel's edges of an UPPER bound of the distance the reference computes for the edge:
//                   the distance from p to the edge's chord plus the curve's maximal deviation from the chord (linear and
//                   quadratic edges: the reference's closed forms find the true minimum), or the distance to the nearer end
//                   point (cubic edges: the reference's Newton search is approximate, but never worse than the end points,
//                   core/edge-segments.cpp:228-277).  Hence U[k] >= the channel's final minimum |distance|.
//   pass 2  an edge whose LOWER bound (chord distance minus deviation) exceeds U[k] for all of its channels cannot become or tie
//           with a minimum: its signed distance is never evaluated.  Its perpendicular candidates matter only if |pd| <= the
//           final minimum (a larger one is discarded by computeDistance in favour of the nearest edge's own distance,
//           core/edge-selectors.cpp:108-117), and since |pd| <= U < lower bound <= |d(edge)| the reference's acceptance test
//           |pd| < |d(edge)| is then known to hold — so such a candidate is added iff its domain tests hold and |pd| <= U.
//           Survivors and candidates are recorded as bit masks (32 records per round).
//   pass 3  (a second kernel, see below) survivors: EdgeSegment::signedDistance in fp64, selector update, perpendicular candidates exactly as the reference;
//           candidates of dropped edges: the fp64 domain tests and cross product.  The warp runs through the UNION of its lanes'
//           masks in evaluation order (record-major: uniform segment type, broadcast loads), so ties resolve as in the full walk.
//
// Rigour of the single-precision bounds: coordinates are taken relative to the centre O of the glyph's texel block, M bounds
// every local coordinate (control points and texel centres), and every bound carries a slack of 2^-16 M + 2^-44 |O| — an
// order of magnitude above the accumulated rounding of the filter arithmetic (inputs 2^-24 M each, about ten operations).
// All rejections are written as !(x > y), so a NaN or an overflow anywhere keeps the edge (it is then evaluated exactly).
#pragma once

#include "distance_core.cuh"

namespace b200 {

struct __align__(16) FRec {
    float ax, ay, abx, aby;       // chord: start point (local coordinates) and vector to the end point
    float adx, ady, bdx, bdy;     // direction(0), direction(1), normalize(true)
    float aqx, aqy, bqx, bqy;     // corner bisectors at the start and at the end
    float inv;                    // 1 / |ab|^2 (0 for a degenerate chord)
    float eps;                    // deviation of the curve from its chord + slack
    int mask;                     // channels (EdgeColor), bit 8: cubic
    int pad;
};

#define FREC_CUBIC 256

__device__ __forceinline__ float sqrt_approx(float x) {
#ifdef B200_HOST_SHIM
    return b200_shim_sqrt_approx(x);        // tests/devcore: the host stand-in, with a selectable error of the instruction's size
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// float not below |x| (x finite double): round to nearest, then one step up
__device__ __forceinline__ float float_up(double x) { return __double2float_ru(fabs(x)); }

// Largest |local coordinate| of the record's control points (for M).
__device__ __forceinline__ float record_extent(const SegRec &r, d2 O) {
    double m = fmax(fabs(r.p0.x-O.x), fabs(r.p0.y-O.y));
    m = fmax(m, fmax(fabs(r.p1.x-O.x), fabs(r.p1.y-O.y)));
    if (r.order >= 2) m = fmax(m, fmax(fabs(r.p2.x-O.x), fabs(r.p2.y-O.y)));
    if (r.order >= 3) m = fmax(m, fmax(fabs(r.p3.x-O.x), fabs(r.p3.y-O.y)));
    return m == m ? float_up(m) : __int_as_float(0x7f800000);
}

template <int MODE>
__device__ __forceinline__ void make_filter_record(const SegRec &r, d2 O, float slack, FRec &f) {
    const d2 a = r.p0, b = r.pEnd;
    f.ax = float(a.x-O.x); f.ay = float(a.y-O.y);
    f.abx = float(b.x-a.x); f.aby = float(b.y-a.y);
    f.adx = float(r.aDir.x); f.ady = float(r.aDir.y);
    f.bdx = float(r.bDir.x); f.bdy = float(r.bDir.y);
    f.aqx = float(r.aBis.x); f.aqy = float(r.aBis.y);
    f.bqx = float(r.bBis.x); f.bqy = float(r.bBis.y);
    const float len2 = f.abx*f.abx+f.aby*f.aby;
    float extra = 0.f;
    if (len2 > 1e-30f && len2 < 1e30f)
        f.inv = 1.f/len2;
    else {
        f.inv = 0.f;                                    // the chord is treated as its start point
        extra = float_up(len(b-a))*1.0001f;
    }
    double h = 0;
    if (r.order == 2) {
        // B(t) - lerp(p0, p2, t) = -t(1-t) (p0 - 2 p1 + p2): at most |br|/4
        h = .25*len(mk(r.e[2], r.e[3]));
    } else if (r.order == 3) {
        // B(t) - lerp(p0, p3, t) = t(1-t)^2 (3p1 - 2p0 - p3) + t^2(1-t) (3p2 - p0 - 2p3): at most max(|v1|, |v2|)/4
        d2 v1 = mk(3*r.p1.x-2*r.p0.x-r.p3.x, 3*r.p1.y-2*r.p0.y-r.p3.y);
        d2 v2 = mk(3*r.p2.x-r.p0.x-2*r.p3.x, 3*r.p2.y-r.p0.y-2*r.p3.y);
        h = .25*fmax(len(v1), len(v2));
    }
    f.eps = float_up(h)*1.0001f+extra+slack;
    f.mask = ((MODE == 1 || MODE == 2) ? 1 : (r.color&7))|(r.order == 3 ? FREC_CUBIC : 0);
    f.pad = 0;
}

// Per-glyph data of the filter, staged in shared memory next to the records.
struct FilterFrame {
    d2 O;               // origin of the local coordinates
    float slack;        // 2^-16 M + 2^-44 |O|
    int nGroups;        // non-empty contours
};

// Origin of the filter's local coordinates: the centre of the glyph's texel block in shape units
__device__ __forceinline__ d2 filter_origin(const GlyphDesc &g) {
    return mk((g.x0+.5*g.w)/g.sx-g.tx, (g.y0+.5*g.h)/g.sy-g.ty);
}

// slack of the fp32 filter for glyph g given the extent of its control points (M of the header comment: the larger of that
// extent and the half size of the texel block)
__device__ __forceinline__ float filter_slack(const GlyphDesc &g, d2 O, float extent) {
    float M = extent;
    M = fmaxf(M, float_up((.5*g.w+1)/fabs(g.sx)));
    M = fmaxf(M, float_up((.5*g.h+1)/fabs(g.sy)));
    return M*1.52587890625e-5f+float_up(fmax(fabs(O.x), fabs(O.y)))*5.7e-14f+1e-37f;       // 2^-16 M + 2^-44 |O|
}

// ---------------------------------------------------------------------------------------------------------------------
// The walk is split over two kernels.  k1a_filter runs passes 1 and 2 (single precision, few registers, a loop body of a few
// dozen instructions that stays in the L0 instruction cache, high occupancy) and leaves per texel a short record in HBM/L2:
//     for every contour (group):  U[0..NCH)  as float bits,  then for every round of 32 records:  surv [, candA, candB]
// k1b_exact replays the same loop structure, reads the words and runs pass 3 plus the combiner in double precision.  The words
// of the 32 texels of a unit are interleaved (word w of lane l at base + 32 w + l), so both sides move whole 128-byte lines.

// Number of filter words per texel of a glyph whose non-empty contours have the given record counts.
template <int MODE>
__host__ __device__ inline unsigned filter_words_group(unsigned nRecords) {
    return unsigned(ModeTraits<MODE>::NCH)+((nRecords+31u)/32u)*(MODE == 1 ? 1u : 3u);
}

// Passes 1 and 2 at the local point (plx, ply).  out: the unit's word block, already offset by the lane.
template <int MODE, bool OVERLAP>
__device__ __forceinline__ void filter_texel(const FRec *frecs, int nRec, const int *groupEnd, int nGroupsAll, float slack, float plx, float ply,
                                             unsigned *out) {
    const int NCH = ModeTraits<MODE>::NCH;
    const int nGroups = OVERLAP ? nGroupsAll : (nRec > 0 ? 1 : 0);
    int j0 = 0;
    for (int g = 0; g < nGroups; ++g) {
        const int jEnd = OVERLAP ? groupEnd[g] : nRec;
        // ---- pass 1: upper bounds of the channels' minima
        float U[NCH];
        #pragma unroll
        for (int k = 0; k < NCH; ++k)
            U[k] = 3.0e38f;
        #pragma unroll 2
        for (int j = j0; j < jEnd; ++j) {
            const float4 c = reinterpret_cast<const float4 *>(frecs+j)[0];
            const float4 m = reinterpret_cast<const float4 *>(frecs+j)[3];       // inv, eps, mask, -
            const int mask = __float_as_int(m.z);
            const float apx = plx-c.x, apy = ply-c.y;
            float t = __saturatef(__fmaf_rn(apx, c.z, apy*c.w)*m.x);
            if (mask&FREC_CUBIC)
                t = t >= .5f ? 1.f : 0.f;       // cubic edges: only the end points bound the reference's result from above
            const float qx = __fmaf_rn(-t, c.z, apx), qy = __fmaf_rn(-t, c.w, apy);
            const float u = __fmaf_rn(sqrt_approx(__fmaf_rn(qx, qx, qy*qy)), 1.000004f, m.y);
            #pragma unroll
            for (int k = 0; k < NCH; ++k)
                if (mask&(1<<k))
                    U[k] = fminf(U[k], u);
        }
        #pragma unroll
        for (int k = 0; k < NCH; ++k) {
            *out = __float_as_uint(U[k]);
            out += 32;
        }
        // ---- pass 2, 32 records per round
        for (int c0 = j0; c0 < jEnd; c0 += 32) {
            const int c1 = c0+32 < jEnd ? c0+32 : jEnd;
            unsigned surv = 0, candA = 0, candB = 0;
            #pragma unroll 1
            for (int j = c0; j < c1; ++j) {
                const float4 c = reinterpret_cast<const float4 *>(frecs+j)[0];
                const float4 m = reinterpret_cast<const float4 *>(frecs+j)[3];
                const int mask = __float_as_int(m.z);
                const unsigned bit = 1u<<(j-c0);
                float Um = 0.f;
                #pragma unroll
                for (int k = 0; k < NCH; ++k)
                    if (mask&(1<<k))
                        Um = fmaxf(Um, U[k]);
                const float apx = plx-c.x, apy = ply-c.y;
                const float t = __saturatef(__fmaf_rn(apx, c.z, apy*c.w)*m.x);
                const float qx = __fmaf_rn(-t, c.z, apx), qy = __fmaf_rn(-t, c.w, apy);
                const float d2f = __fmaf_rn(qx, qx, qy*qy);
                const float thr = Um+m.y;
                const bool keep = !(d2f > thr*thr*1.00001f);
                if (keep)
                    surv |= bit;
                if (MODE != 1) {
                    const float4 dir = reinterpret_cast<const float4 *>(frecs+j)[1];
                    const float4 bis = reinterpret_cast<const float4 *>(frecs+j)[2];
                    // |pd| limit: a dropped edge's candidate matters only below U; a survivor's is accepted by the reference iff
                    // |pd| < |d(edge)|, and |d(edge)| <= chord distance + deviation (no such bound for the cubic Newton search)
                    const float lim = (keep ? ((mask&FREC_CUBIC) ? 3.0e38f : __fmaf_rn(sqrt_approx(d2f), 1.000004f, m.y)) : Um)+slack;
                    // A: cross(ap, aDir), dot(ap, aBis) > 0, -dot(ap, aDir) > 0
                    const float pdA = __fmaf_rn(apx, dir.y, -apy*dir.x);
                    if (!(fabsf(pdA) > lim) && !(__fmaf_rn(apx, bis.x, apy*bis.y) < -slack) && !(__fmaf_rn(apx, dir.x, apy*dir.y) > slack))
                        candA |= bit;
                    // B: cross(bp, bDir), -dot(bp, bBis) > 0, dot(bp, bDir) > 0
                    const float bpx = apx-c.z, bpy = apy-c.w;
                    const float pdB = __fmaf_rn(bpx, dir.w, -bpy*dir.z);
                    if (!(fabsf(pdB) > lim) && !(__fmaf_rn(bpx, bis.z, bpy*bis.w) > slack) && !(__fmaf_rn(bpx, dir.z, bpy*dir.w) < -slack))
                        candB |= bit;
                }
            }
            *out = surv;
            out += 32;
            if (MODE != 1) {
                out[0] = candA;
                out[32] = candB;
                out += 64;
            }
        }
        j0 = jEnd;
    }
}

// ---- the overlapping combiner of k1b_exact, compact form ------------------------------------------------------------------
// computeDistance (edge-selectors.cpp:108-117) of a channel is  m = d < 0 ? minNeg : minPos;  dd = the nearest edge's distance
// after distanceToPerpendicularDistance;  |dd| < |m| ? dd : m.  dd depends only on (d, nearest edge, p), so when selectors are
// MERGED (contour-combiners.cpp:84-92) the merged channel's dd is the dd of the contour that won the merge: it is computed once
// per contour and channel here and parked instead of the edge reference.  Parked per contour: per channel d, dot [, dd, neg,
// pos], then the NOUT resolved distances and the winding (StoreLayout).  dd = DBL_MAX stands for "no nearest edge".
__device__ __noinline__ double near_edge_distance(double cd, int near, const SegRec *recs, double px, double py) {
    if (near < 0)
        return DBL_MAX;
    double dd = cd;
    const int side = near&3;
    if (side) {
        const d2 p = mk(px, py);
        const SegRec &s = recs[near>>2];
        if (side == 1) {
            d2 aq = p-s.p0;
            if (dot(aq, s.aDirNZ) < 0) {
                double pd = cross(aq, s.aDirNZ);
                if (fabs(pd) <= fabs(dd))
                    dd = pd;
            }
        } else {
            d2 bq = p-s.pEnd;
            if (dot(bq, s.bDirNZ) > 0) {
                double pd = cross(bq, s.bDirNZ);
                if (fabs(pd) <= fabs(dd))
                    dd = pd;
            }
        }
    }
    return dd;
}

template <int MODE>
__device__ __forceinline__ void park_contour(ContourStore &st, int c, const Sel<MODE> &s, const SegRec *recs, d2 p, double winding) {
    const int NCH = ModeTraits<MODE>::NCH;
    typedef StoreLayout<MODE> L;
    double td = s.ch[0].d, tdot = s.ch[0].dot;
    #pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int o = k*L::CH;
        st.at(c, o) = s.ch[k].d;
        st.at(c, o+1) = s.ch[k].dot;
        double cd = s.ch[k].d;
        if (MODE != 1) {
            const double dd = near_edge_distance(s.ch[k].d, s.ch[k].near, recs, p.x, p.y);
            const double m = s.ch[k].d < 0 ? s.ch[k].neg : s.ch[k].pos;
            cd = fabs(dd) < fabs(m) ? dd : m;
            st.at(c, o+2) = dd;
            st.at(c, o+3) = s.ch[k].neg;
            st.at(c, o+4) = s.ch[k].pos;
        }
        st.at(c, L::DIST+k) = cd;
        if (MODE == 4 && k > 0 && sd_less(s.ch[k].d, s.ch[k].dot, td, tdot)) { td = s.ch[k].d; tdot = s.ch[k].dot; }
    }
    if (MODE == 4)
        st.at(c, L::DIST+3) = td;      // MultiAndTrueDistanceSelector: the true distance rides along as the fourth value
    st.at(c, L::WINDING) = winding;
}

// OverlappingContourCombiner::distance (contour-combiners.cpp:78-134) over the nc parked contours.
template <int MODE>
__device__ __forceinline__ void combine_parked(ContourStore &st, int nc, double *out) {
    const int NCH = ModeTraits<MODE>::NCH;
    const int NOUT = ModeTraits<MODE>::NOUT;
    typedef StoreLayout<MODE> L;
    if (nc == 1) {
        // one contour: the shape, inner and outer selectors are that contour's or empty, and every branch of the combiner ends
        // at the contour's own distance (an infinite distance takes the general path, whose result is then an empty selector's)
        double cd[NOUT];
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) cd[k] = st.at(0, L::DIST+k);
        if (!(fabs(resolve<MODE>(cd)) > DBL_MAX)) {
            #pragma unroll
            for (int k = 0; k < NOUT; ++k) out[k] = cd[k];
            return;
        }
    }
    // shape (which = 0), inner (1), outer (2) selectors: merged one after the other to keep the register footprint small
    double D[3][NOUT];
    #pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        double ad[NCH], adot[NCH], add[NCH], aneg[NCH], apos[NCH];
        #pragma unroll
        for (int k = 0; k < NCH; ++k) { ad[k] = -DBL_MAX; adot[k] = 0; add[k] = DBL_MAX; aneg[k] = -DBL_MAX; apos[k] = DBL_MAX; }
        #pragma unroll 1
        for (int c = 0; c < nc; ++c) {
            if (which) {
                double cd[NOUT];
                #pragma unroll
                for (int k = 0; k < NOUT; ++k) cd[k] = st.at(c, L::DIST+k);
                const double w = st.at(c, L::WINDING), rd = resolve<MODE>(cd);
                if (!(which == 1 ? (w > 0 && rd >= 0) : (w < 0 && rd <= 0)))
                    continue;
            }
            #pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int o = k*L::CH;
                const double od = st.at(c, o), odot = st.at(c, o+1);
                if (sd_less(od, odot, ad[k], adot[k])) {
                    ad[k] = od; adot[k] = odot;
                    if (MODE != 1) add[k] = st.at(c, o+2);
                }
                if (MODE != 1) {
                    const double on = st.at(c, o+3), op = st.at(c, o+4);
                    if (on > aneg[k]) aneg[k] = on;
                    if (op < apos[k]) apos[k] = op;
                }
            }
        }
        double td = ad[0], tdot = adot[0];
        #pragma unroll
        for (int k = 0; k < NCH; ++k) {
            double v = ad[k];
            if (MODE != 1) {
                const double m = ad[k] < 0 ? aneg[k] : apos[k];
                v = fabs(add[k]) < fabs(m) ? add[k] : m;
            }
            D[which][k] = v;
            if (MODE == 4 && k > 0 && sd_less(ad[k], adot[k], td, tdot)) { td = ad[k]; tdot = adot[k]; }
        }
        if (MODE == 4)
            D[which][3] = td;
    }
    const double *shapeD = D[0], *innerD = D[1], *outerD = D[2];
    double dist[NOUT];
    const double innerScalar = resolve<MODE>(innerD), outerScalar = resolve<MODE>(outerD);
    double winding;
    if (innerScalar >= 0 && fabs(innerScalar) <= fabs(outerScalar)) {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) dist[k] = innerD[k];
        winding = 1;
    } else if (outerScalar <= 0 && fabs(outerScalar) < fabs(innerScalar)) {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) dist[k] = outerD[k];
        winding = -1;
    } else {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) out[k] = shapeD[k];
        return;
    }
    // Pass 0: same-winding contours refine the inner/outer pick.  Pass 1: the others may override it.
    #pragma unroll 1
    for (int pass = 0; pass < 2; ++pass)
        #pragma unroll 1
        for (int c = 0; c < nc; ++c) {
            const double w = st.at(c, L::WINDING);
            double cd[NOUT];
            #pragma unroll
            for (int k = 0; k < NOUT; ++k)
                cd[k] = st.at(c, L::DIST+k);
            const double rc = resolve<MODE>(cd), rdist = resolve<MODE>(dist);
            bool take;
            if (pass == 0) {
                if (winding > 0)
                    take = w > 0 && fabs(rc) < fabs(outerScalar) && rc > rdist;
                else
                    take = w < 0 && fabs(rc) < fabs(innerScalar) && rc < rdist;
            } else
                take = w != winding && rc*rdist >= 0 && fabs(rc) < fabs(rdist);
            if (take) {
                #pragma unroll
                for (int k = 0; k < NOUT; ++k) dist[k] = cd[k];
            }
        }
    if (resolve<MODE>(dist) == resolve<MODE>(shapeD)) {
        #pragma unroll
        for (int k = 0; k < NOUT; ++k) dist[k] = shapeD[k];
    }
    #pragma unroll
    for (int k = 0; k < NOUT; ++k) out[k] = dist[k];
}

// Pass 3 and the contour combiner at p: ShapeDistanceFinder<Combiner<Selector>>::oneShotDistance restricted to what the filter
// kept.  in: the unit's word block offset by the lane.  Must be called by ALL 32 lanes of the warp (warp-wide votes inside).
// guard_flag() tells afterwards whether the texel has to be re-evaluated with the literal root solver (sd_quadratic).
template <int MODE, bool OVERLAP>
__device__ __forceinline__ void exact_texel(const SegRec *recs, int nRec, const int *groupEnd, int nGroupsAll, const unsigned *in,
                                            d2 p, ContourStore store, double *out) {
    const int NCH = ModeTraits<MODE>::NCH;
    const int NOUT = ModeTraits<MODE>::NOUT;
    typedef StoreLayout<MODE> L;
    const int nGroups = OVERLAP ? nGroupsAll : (nRec > 0 ? 1 : 0);

    Sel<MODE> cur;
    sel_init(cur);
    int nc = 0;
    int j0 = 0;
    // The words of the NEXT contour (bounds + first round of masks) are requested while the current one is evaluated.
    const int WPR = MODE == 1 ? 1 : 3;          // words per round
    unsigned nU[NCH], nS = 0, nA = 0, nB = 0;
    if (nGroups > 0) {
        #pragma unroll
        for (int k = 0; k < NCH; ++k)
            nU[k] = __ldcs(in+32*k);
        nS = __ldcs(in+32*NCH);
        if (MODE != 1) {
            nA = __ldcs(in+32*(NCH+1));
            nB = __ldcs(in+32*(NCH+2));
        }
    }
    for (int g = 0; g < nGroups; ++g) {
        const int jEnd = OVERLAP ? groupEnd[g] : nRec;
        float U[NCH];
        #pragma unroll
        for (int k = 0; k < NCH; ++k)
            U[k] = __uint_as_float(nU[k]);
        unsigned surv = nS, candA = nA, candB = nB;
        const unsigned *more = in+32*(NCH+WPR);                     // this contour's further rounds
        in += 32*(NCH+WPR*((jEnd-j0+31)>>5));                        // the next contour's words
        if (g+1 < nGroups) {
            #pragma unroll
            for (int k = 0; k < NCH; ++k)
                nU[k] = __ldcs(in+32*k);
            nS = __ldcs(in+32*NCH);
            if (MODE != 1) {
                nA = __ldcs(in+32*(NCH+1));
                nB = __ldcs(in+32*(NCH+2));
            }
        }
        for (int c0 = j0; c0 < jEnd; c0 += 32) {
            if (c0 != j0) {
                surv = __ldcs(more);
                if (MODE != 1) {
                    candA = __ldcs(more+32);
                    candB = __ldcs(more+64);
                }
                more += 32*WPR;
            }
            // ---- pass 3a: survivors, exactly as the reference (edge-selectors.cpp:129-160 / 174-227).  Record-major over the
            // union of the warp's survivors: the record, hence the segment type, is uniform across the warp and its loads
            // broadcast; a lane without the bit idles.
            unsigned survAny = __reduce_or_sync(0xffffffffu, surv);
            while (survAny) {
                const int b = __ffs(survAny)-1;
                survAny &= survAny-1;
                if (!((surv>>b)&1u))
                    continue;
                const int j = c0+b;
                const SegRec &s = recs[j];
                const int mask = (MODE == 1 || MODE == 2) ? 1 : s.color;
                // the orthogonality measure only breaks ties between candidates for a channel's minimum: |d| above every
                // channel's current minimum (or above its upper bound) cannot become one
                double need = 0;
                #pragma unroll
                for (int k = 0; k < NCH; ++k)
                    if (mask&(1<<k))
                        need = dmax(need, dmin(fabs(cur.ch[k].d), double(U[k])*1.000001));
                const SDist r = signed_distance(s, p, need);
                if (fabs(r.d) <= need) {
                    const int near = (j<<2)|r.side;
                    #pragma unroll
                    for (int k = 0; k < NCH; ++k)
                        if (mask&(1<<k))
                            if (sd_less(r.d, r.dot, cur.ch[k].d, cur.ch[k].dot)) {
                                cur.ch[k].d = r.d; cur.ch[k].dot = r.dot; cur.ch[k].near = near;
                            }
                }
                if (MODE != 1 && ((candA>>b)&1u)) {
                    const d2 ap = p-s.p0;
                    const double pdA = cross(ap, s.aDir);
                    if (dot(ap, s.aBis) > 0 && -dot(ap, s.aDir) > 0 && fabs(pdA) < fabs(r.d)) {
                        #pragma unroll
                        for (int k = 0; k < NCH; ++k)
                            if (mask&(1<<k))
                                chan_add_perp(cur.ch[k], pdA);
                    }
                }
                if (MODE != 1 && ((candB>>b)&1u)) {
                    const d2 bp = p-s.pEnd;
                    const double pdB = cross(bp, s.bDir);
                    if (-dot(bp, s.bBis) > 0 && dot(bp, s.bDir) > 0 && fabs(pdB) < fabs(r.d)) {
                        #pragma unroll
                        for (int k = 0; k < NCH; ++k)
                            if (mask&(1<<k))
                                chan_add_perp(cur.ch[k], pdB);
                    }
                }
            }
            // ---- pass 3b: perpendicular candidates of dropped edges
            if (MODE != 1) {
                candA &= ~surv;         // (the survivors' candidates were handled with their exact distances above)
                candB &= ~surv;
                unsigned candAny = __reduce_or_sync(0xffffffffu, candA|candB);
                while (candAny) {
                    const int b = __ffs(candAny)-1;
                    candAny &= candAny-1;
                    const bool isA = (candA>>b)&1u, isB = (candB>>b)&1u;
                    if (!isA && !isB)
                        continue;
                    const SegRec &s = recs[c0+b];
                    const int mask = MODE == 2 ? 1 : s.color;
                    double Um = 0;
                    #pragma unroll
                    for (int k = 0; k < NCH; ++k)
                        if (mask&(1<<k))
                            Um = dmax(Um, double(U[k]));
                    if (isA) {
                        const d2 ap = p-s.p0;
                        const double pdA = cross(ap, s.aDir);
                        if (dot(ap, s.aBis) > 0 && -dot(ap, s.aDir) > 0 && fabs(pdA) <= Um) {
                            #pragma unroll
                            for (int k = 0; k < NCH; ++k)
                                if (mask&(1<<k))
                                    chan_add_perp(cur.ch[k], pdA);
                        }
                    }
                    if (isB) {
                        const d2 bp = p-s.pEnd;
                        const double pdB = cross(bp, s.bDir);
                        if (-dot(bp, s.bBis) > 0 && dot(bp, s.bDir) > 0 && fabs(pdB) <= Um) {
                            #pragma unroll
                            for (int k = 0; k < NCH; ++k)
                                if (mask&(1<<k))
                                    chan_add_perp(cur.ch[k], pdB);
                        }
                    }
                }
            }
        }
        if (OVERLAP) {
            // park the finished contour (channel minima, resolved distances, winding) and start the next one
            park_contour<MODE>(store, nc, cur, recs, p, double(recs[jEnd-1].winding));
            sel_init(cur);
            ++nc;
        }
        j0 = jEnd;
    }
    if (!OVERLAP)
        sel_distance<MODE>(cur, recs, p, out);
    else
        combine_parked<MODE>(store, nc, out);
}

// ---------------------------------------------------------------------------------------------------------------------
// Big shapes (k1_big_blocks): the same filter in front of the per-texel walk over a warp's record sub-list.  One kernel, no mask
// words: pass 1 gives U[k] over the contour's listed records, pass 2 evaluates a survivor — or a dropped edge's perpendicular
// candidate — exactly on the spot (records of big shapes live in global memory; their FRecs are 64 bytes instead of 352).
// OVERLAP: contours are folded into the shape / inner / outer accumulators as they finish (shape_distance_fold).
template <int MODE, bool OVERLAP, typename IT>
__device__ __forceinline__ void shape_distance_list_filtered(const SegRec *recs, const FRec *frecs, int nRec, IT it, d2 O, float slack, d2 p,
                                                             FoldStore st, double *out) {
    const int NCH = ModeTraits<MODE>::NCH;
    const int NOUT = ModeTraits<MODE>::NOUT;
    const float plx = float(p.x-O.x), ply = float(p.y-O.y);
    FoldAcc<MODE> shape, inner, outer;
    if (OVERLAP) { fold_init(shape); fold_init(inner); fold_init(outer); }
    int nc = 0;
    Sel<MODE> cur;
    sel_init(cur);
    int j0 = 0;
    while (j0 < nRec) {
        const int j1 = OVERLAP ? it.group_last(recs, j0, nRec) : nRec-1;
        // ---- pass 1
        float U[NCH];
        #pragma unroll
        for (int k = 0; k < NCH; ++k)
            U[k] = 3.0e38f;
        for (int i = j0; i <= j1; ++i) {
            const int j = it.at(i);
            const float4 c = __ldg(reinterpret_cast<const float4 *>(frecs+j));
            const float4 m = __ldg(reinterpret_cast<const float4 *>(frecs+j)+3);
            const int mask = __float_as_int(m.z);
            const float apx = plx-c.x, apy = ply-c.y;
            float t = __saturatef(__fmaf_rn(apx, c.z, apy*c.w)*m.x);
            if (mask&FREC_CUBIC)
                t = t >= .5f ? 1.f : 0.f;
            const float qx = __fmaf_rn(-t, c.z, apx), qy = __fmaf_rn(-t, c.w, apy);
            const float u = __fmaf_rn(sqrt_approx(__fmaf_rn(qx, qx, qy*qy)), 1.000004f, m.y);
            #pragma unroll
            for (int k = 0; k < NCH; ++k)
                if (mask&(1<<k))
                    U[k] = fminf(U[k], u);
        }
        // ---- pass 2 with the exact work in place
        for (int i = j0; i <= j1; ++i) {
            const int j = it.at(i);
            const float4 c = __ldg(reinterpret_cast<const float4 *>(frecs+j));
            const float4 m = __ldg(reinterpret_cast<const float4 *>(frecs+j)+3);
            const int fmask = __float_as_int(m.z);
            float Um = 0.f;
            #pragma unroll
            for (int k = 0; k < NCH; ++k)
                if (fmask&(1<<k))
                    Um = fmaxf(Um, U[k]);
            const float apx = plx-c.x, apy = ply-c.y;
            const float t = __saturatef(__fmaf_rn(apx, c.z, apy*c.w)*m.x);
            const float qx = __fmaf_rn(-t, c.z, apx), qy = __fmaf_rn(-t, c.w, apy);
            const float d2f = __fmaf_rn(qx, qx, qy*qy);
            const float thr = Um+m.y;
            if (!(d2f > thr*thr*1.00001f)) {
                // survivor: exactly as the reference (edge-selectors.cpp:129-160 / 174-227)
                const SegRec &s = recs[j];
                const int mask = (MODE == 1 || MODE == 2) ? 1 : s.color;
                double need = 0;
                #pragma unroll
                for (int k = 0; k < NCH; ++k)
                    if (mask&(1<<k))
                        need = dmax(need, dmin(fabs(cur.ch[k].d), double(U[k])*1.000001));
                const SDist r = signed_distance(s, p, need);
                if (fabs(r.d) <= need) {
                    const int near = (j<<2)|r.side;
                    #pragma unroll
                    for (int k = 0; k < NCH; ++k)
                        if (mask&(1<<k))
                            if (sd_less(r.d, r.dot, cur.ch[k].d, cur.ch[k].dot)) {
                                cur.ch[k].d = r.d; cur.ch[k].dot = r.dot; cur.ch[k].near = near;
                            }
                }
                if (MODE != 1) {
                    const d2 ap = p-s.p0;
                    const double pdA = cross(ap, s.aDir);
                    if (dot(ap, s.aBis) > 0 && -dot(ap, s.aDir) > 0 && fabs(pdA) < fabs(r.d)) {
                        #pragma unroll
                        for (int k = 0; k < NCH; ++k)
                            if (mask&(1<<k))
                                chan_add_perp(cur.ch[k], pdA);
                    }
                    const d2 bp = p-s.pEnd;
                    const double pdB = cross(bp, s.bDir);
                    if (-dot(bp, s.bBis) > 0 && dot(bp, s.bDir) > 0 && fabs(pdB) < fabs(r.d)) {
                        #pragma unroll
                        for (int k = 0; k < NCH; ++k)
                            if (mask&(1<<k))
                                chan_add_perp(cur.ch[k], pdB);
                    }
                }
            } else if (MODE != 1) {
                // dropped edge: its perpendicular candidates (see the header of this file)
                const float4 dir = __ldg(reinterpret_cast<const float4 *>(frecs+j)+1);
                const float4 bis = __ldg(reinterpret_cast<const float4 *>(frecs+j)+2);
                const float lim = Um+slack;
                const float pdAf = __fmaf_rn(apx, dir.y, -apy*dir.x);
                const bool cA = !(fabsf(pdAf) > lim) && !(__fmaf_rn(apx, bis.x, apy*bis.y) < -slack) && !(__fmaf_rn(apx, dir.x, apy*dir.y) > slack);
                const float bpx = apx-c.z, bpy = apy-c.w;
                const float pdBf = __fmaf_rn(bpx, dir.w, -bpy*dir.z);
                const bool cB = !(fabsf(pdBf) > lim) && !(__fmaf_rn(bpx, bis.z, bpy*bis.w) > slack) && !(__fmaf_rn(bpx, dir.z, bpy*dir.w) < -slack);
                if (cA || cB) {
                    const SegRec &s = recs[j];
                    const int mask = MODE == 2 ? 1 : s.color;
                    const double UmD = double(Um);
                    if (cA) {
                        const d2 ap = p-s.p0;
                        const double pdA = cross(ap, s.aDir);
                        if (dot(ap, s.aBis) > 0 && -dot(ap, s.aDir) > 0 && fabs(pdA) <= UmD) {
                            #pragma unroll
                            for (int k = 0; k < NCH; ++k)
                                if (mask&(1<<k))
                                    chan_add_perp(cur.ch[k], pdA);
                        }
                    }
                    if (cB) {
                        const d2 bp = p-s.pEnd;
                        const double pdB = cross(bp, s.bDir);
                        if (-dot(bp, s.bBis) > 0 && dot(bp, s.bDir) > 0 && fabs(pdB) <= UmD) {
                            #pragma unroll
                            for (int k = 0; k < NCH; ++k)
                                if (mask&(1<<k))
                                    chan_add_perp(cur.ch[k], pdB);
                        }
                    }
                }
            }
        }
        if (OVERLAP) {
            // the contour's own distances, its merges, and what the second half of the combiner reads back (shape_distance_fold)
            double dd[NCH], cd[NOUT];
            double td = cur.ch[0].d, tdot = cur.ch[0].dot;
            #pragma unroll
            for (int k = 0; k < NCH; ++k) {
                cd[k] = cur.ch[k].d;
                dd[k] = DBL_MAX;
                if (MODE != 1) {
                    dd[k] = near_edge_dd(cur.ch[k].d, cur.ch[k].near, recs, p.x, p.y);
                    const double mm = cur.ch[k].d < 0 ? cur.ch[k].neg : cur.ch[k].pos;
                    cd[k] = fabs(dd[k]) < fabs(mm) ? dd[k] : mm;
                }
                if (MODE == 4 && k > 0 && sd_less(cur.ch[k].d, cur.ch[k].dot, td, tdot)) { td = cur.ch[k].d; tdot = cur.ch[k].dot; }
            }
            if (MODE == 4)
                cd[3] = td;
            const int w = recs[it.at(j1)].winding;
            const double rd = resolve<MODE>(cd);
            fold_merge<MODE>(shape, cur, dd);
            if (w > 0 && rd >= 0)
                fold_merge<MODE>(inner, cur, dd);
            if (w < 0 && rd <= 0)
                fold_merge<MODE>(outer, cur, dd);
            #pragma unroll
            for (int k = 0; k < NOUT; ++k)
                st.cd(nc, k) = cd[k];
            st.w(nc) = (signed char) w;
            sel_init(cur);
            ++nc;
        }
        j0 = j1+1;
    }
    if (!OVERLAP)
        sel_distance<MODE>(cur, recs, p, out);
    else
        fold_finish<MODE>(shape, inner, outer, st, nc, out);
}

}  // namespace b200
