// devcore_host.cpp — TEST INFRASTRUCTURE: the texel semantics of the CUDA path compiled for the host.
//
// msdfgen_b200/csrc/distance_core.cuh holds everything one GPU thread computes for one texel: segment records, the three
// signed-distance kernels with their root solvers (fast path, guard band, literal path with correctly rounded acos / cos / pow),
// the True / Perpendicular / Multi selectors, the Simple and Overlapping contour combiners (parked and folded form).  None of it
// cooperates across threads, so the SAME SOURCE compiles as plain C++ behind a shim of <cuda_runtime.h> (tests/devcore/shim).
// tests/test_devcore.py runs it on the CPU against the oracle, bit for bit: the device arithmetic is then covered by the
// `-m "not gpu"` tier too, on machines without a GPU.  Nothing in the product links or loads this file; the product has no CPU path.
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -Itests/devcore/shim -Imsdfgen_b200/csrc (-ffp-contract=off mirrors nvcc -fmad=false;
// explicit fma() calls are fused on both sides).
#include "error_correction.cuh"     // (pulls in kernels.cuh, distance_filter.cuh, distance_core.cuh)

#include <vector>

using namespace b200;

int b200_shim_sqrt_bias = 0;
int b200_shim_cos_bias = 0;

namespace {

struct Prepared {
    std::vector<SegRec> recs;
    std::vector<int8_t> windings;
    int nonEmpty;
};

Prepared prepare(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors) {
    Prepared out;
    out.windings.resize((size_t) nContours+1);
    RawShapes rs = { points, orders, colors, cbegin, out.windings.data() };
    out.nonEmpty = 0;
    for (int c = 0; c < nContours; ++c) {
        out.windings[(size_t) c] = (int8_t) contour_winding(rs, (unsigned) c);
        out.nonEmpty += cbegin[c+1] > cbegin[c];
    }
    GlyphDesc g = { };
    g.sx = g.sy = 1;
    g.c0 = 0; g.c1 = (unsigned) nContours;
    g.s0 = 0; g.s1 = cbegin[nContours];
    g.nonEmptyContours = out.nonEmpty;
    g.recBase = -1;
    out.recs.resize(g.s1);
    for (unsigned j = 0; j < g.s1; ++j)
        prepare_record(rs, g, j, out.recs[j]);
    return out;
}

// policy 0: fast root path only; 1: fast path, texels that hit the guard band are redone with the literal path (what the kernels
// do: recheck launch); 2: literal path for every texel; 3: fast path + guard redo through the FOLDED overlapping combiner (the
// big-shape kernel's; overlap only)
template <int MODE, bool OVERLAP>
int run(const Prepared &pr, int policy, int nPts, const double *pts, double *out) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    const int nRec = (int) pr.recs.size();
    const size_t nc = (size_t) pr.nonEmpty+1;
    std::vector<double> parked(nc*StoreLayout<MODE>::PER_CONTOUR), folded(nc*NOUT);
    std::vector<signed char> foldedW(nc);
    ContourStore store = { parked.data(), 1, StoreLayout<MODE>::PER_CONTOUR };
    FoldStore fold = { folded.data(), foldedW.data(), 1, NOUT };
    int guardHits = 0;
    for (int i = 0; i < nPts; ++i) {
        const d2 p = mk(pts[2*i], pts[2*i+1]);
        double *dist = out+(size_t) i*NOUT;
        guard_flag() = 0;
        if (policy == 2)
            shape_distance<MODE, OVERLAP, true>(pr.recs.data(), nRec, p, store, dist);
        else if (policy == 3 && OVERLAP)
            shape_distance_fold<MODE, SeqIter>(pr.recs.data(), nRec, SeqIter(), p, fold, dist);
        else
            shape_distance<MODE, OVERLAP, false>(pr.recs.data(), nRec, p, store, dist);
        if (guard_flag()) {
            ++guardHits;
            guard_flag() = 0;
            if (policy == 1 || policy == 3)
                shape_distance<MODE, OVERLAP, true>(pr.recs.data(), nRec, p, store, dist);
        }
    }
    return guardHits;
}

// The two-kernel path of small glyphs, one texel at a time: k1a_filter's passes 1 and 2 (fp32 bounds, survivor / candidate masks)
// write the texel's filter words, k1b_exact's pass 3 + combiner read them.  The frame (origin of the local coordinates, extent,
// slack, contour groups) is set up as the kernels do (kernels.cuh: k1a_filter / build_groups), for a w x h texel block under the
// projection (sx, sy, tx, ty); the evaluated points are arbitrary, as long as they lie inside that block.
template <int MODE, bool OVERLAP>
int runFiltered(const Prepared &pr, const GlyphDesc &g, const uint32_t *cbegin, int policy, int nPts, const double *pts, double *out, double *survivorsPerTexel) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    const int nRec = (int) pr.recs.size();
    const d2 O = filter_origin(g);
    float extent = 0;
    for (int j = 0; j < nRec; ++j)
        extent = fmaxf(extent, record_extent(pr.recs[(size_t) j], O));
    const float slack = filter_slack(g, O, extent);
    std::vector<FRec> frecs((size_t) nRec+1);
    for (int j = 0; j < nRec; ++j)
        make_filter_record<MODE>(pr.recs[(size_t) j], O, slack, frecs[(size_t) j]);
    std::vector<int> groupEnd;
    unsigned words = 0;
    for (unsigned c = g.c0; c < g.c1; ++c)
        if (cbegin[c+1] > cbegin[c]) {
            groupEnd.push_back((int) (cbegin[c+1]-g.s0));
            words += filter_words_group<MODE>(cbegin[c+1]-cbegin[c]);
        }
    if (!OVERLAP)
        words = nRec ? filter_words_group<MODE>((unsigned) nRec) : 0u;
    groupEnd.push_back(0);
    std::vector<unsigned> block((size_t) words*32u+32u);
    const size_t nc = (size_t) pr.nonEmpty+1;
    std::vector<double> parked(nc*StoreLayout<MODE>::PER_CONTOUR), folded(nc*NOUT);
    std::vector<signed char> foldedW(nc);
    ContourStore store = { parked.data(), 1, StoreLayout<MODE>::PER_CONTOUR };
    FoldStore fold = { folded.data(), foldedW.data(), 1, NOUT };
    int guardHits = 0;
    double survivors = 0;
    for (int i = 0; i < nPts; ++i) {
        const d2 p = mk(pts[2*i], pts[2*i+1]);
        const unsigned lane = (unsigned) i&31u;         // the words of a unit's 32 texels are interleaved: use every lane slot
        double *dist = out+(size_t) i*NOUT;
        guard_flag() = 0;
        if (policy == 3 || policy == 4)
            shape_distance_list_filtered<MODE, OVERLAP, SeqIter>(pr.recs.data(), frecs.data(), nRec, SeqIter(), O, slack, p, fold, dist);
        else {
            std::fill(block.begin(), block.end(), 0xdeadbeefu);
            filter_texel<MODE, OVERLAP>(frecs.data(), nRec, groupEnd.data(), pr.nonEmpty, slack, float(p.x-O.x), float(p.y-O.y), block.data()+lane);
            // statistics: survivors of this texel (mask words sit behind each contour's NCH bounds)
            {
                const unsigned *w = block.data()+lane;
                int j0 = 0;
                const int nGroups = OVERLAP ? pr.nonEmpty : (nRec > 0 ? 1 : 0);
                for (int gi = 0; gi < nGroups; ++gi) {
                    const int jEnd = OVERLAP ? groupEnd[(size_t) gi] : nRec;
                    w += 32*ModeTraits<MODE>::NCH;
                    for (int c0 = j0; c0 < jEnd; c0 += 32) {
                        survivors += __builtin_popcount(w[0]);
                        w += 32*(MODE == 1 ? 1 : 3);
                    }
                    j0 = jEnd;
                }
            }
            exact_texel<MODE, OVERLAP>(pr.recs.data(), nRec, groupEnd.data(), pr.nonEmpty, block.data()+lane, p, store, dist);
        }
        if (guard_flag()) {
            ++guardHits;
            guard_flag() = 0;
            if (policy != 0 && policy != 4)
                shape_distance<MODE, OVERLAP, true>(pr.recs.data(), nRec, p, store, dist);
        }
    }
    if (survivorsPerTexel)
        *survivorsPerTexel = nPts ? survivors/nPts : 0;
    return guardHits;
}

// msdfErrorCorrection of one glyph's field, texel by texel through the device's classification code with the IN-PLACE sink (the one
// k3 uses for texels whose candidate slots overflowed): protectCorners as a scatter over the segments' colour corners, protectEdges,
// then find_errors_texel with both findErrors passes in one walk and the exact shape distance evaluated on the spot, then apply.
// The flag logic around the calls follows k2_classify (error_correction.cuh) line for line.
template <int N, bool OVERLAP>
void runErrorCorrection(const Prepared &pr, const RawShapes &rs, const GlyphDesc &g, int ecMode, int ecDcm, double minDeviationRatio,
                        double minImproveRatio, float *field, uint8_t *stencil) {
    const int w = g.w, h = g.h;
    const bool useBase = ecDcm == 0 || (ecDcm == 1 && ecMode != 3);
    const bool useShape = ecDcm == 1 || ecDcm == 2;
    std::vector<int> state((size_t) w*h, 0);
    if (ecMode == 2)
        for (unsigned j = 0; j < g.s1-g.s0; ++j) {
            const int2 c = corner_of(rs, g, j);
            for (int d = 0; d < 4; ++d) {
                const int x = c.x+(d&1), y = c.y+(d>>1);
                if (x >= 0 && y >= 0 && x < w && y < h)
                    state[(size_t) y*w+x] |= 1;
            }
        }
    std::vector<double> parked(((size_t) pr.nonEmpty+1)*StoreLayout<2>::PER_CONTOUR);
    EcCtx<N> e;
    e.f.base = field; e.f.w = w; e.f.h = h; e.f.flip = 0;
    e.g = &g;
    e.minImproveRatio = minImproveRatio;
    e.slots = nullptr; e.slotCount = nullptr; e.slotCap = 0;
    e.recs = pr.recs.data(); e.nRec = (int) pr.recs.size();
    e.store.base = parked.data(); e.store.stride = 1; e.store.perContour = StoreLayout<2>::PER_CONTOUR;
    const EcSpans sp = ec_spans(g, minDeviationRatio);
    std::vector<uint8_t> flags((size_t) w*h, 0);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t i = (size_t) y*w+x;
            bool prot = (state[i]&1) != 0;
            if (ecMode == 2) {
                if (!prot)
                    prot = protect_edges_texel<N>(e.f, x, y, sp.hRadius, sp.vRadius, sp.dRadius);
            } else if (ecMode == 3)
                prot = true;
            TexelState k;
            k.useBase = useBase;
            k.protBase = prot;
            k.useShape = useShape;
            k.protShape = k.protBase || (useBase && ecDcm == 1);
            k.errBase = k.errShape = k.overflow = false;
            k.texel = (unsigned) i;
            k.queue = nullptr; k.queueCount = nullptr; k.tileTexel = 0;
            guard_flag() = 0;
            find_errors_texel<N, 1, OVERLAP>(e, k, x, y, sp.hSpan, sp.vSpan, sp.dSpan);
            int f = 0;
            if (ecMode == 3 || prot)
                f |= EC_FLAG_PROTECTED;
            if (k.errBase || k.errShape)
                f |= EC_FLAG_ERROR;
            if (k.protShape && useShape)
                f |= EC_FLAG_PROTECTED;
            flags[i] = (uint8_t) f;
        }
    // apply: the classification above read the UNCORRECTED field everywhere, as the reference does (findErrors only marks)
    for (size_t i = 0; i < flags.size(); ++i) {
        if (flags[i]&EC_FLAG_ERROR) {
            float *c = field+i*N;
            const float m = median3f(c[0], c[1], c[2]);
            c[0] = c[1] = c[2] = m;
        }
        if (stencil)
            stencil[i] = flags[i];
    }
}

}  // namespace

extern "C" {

/// msdfErrorCorrection (mode 1 INDISCRIMINATE / 2 EDGE_PRIORITY / 3 EDGE_ONLY, distance check 0 / 1 / 2) of field[h][w][channels] in
/// place, rows in shape orientation, under the transformation xf = (sx, sy, tx, ty, range lower, range upper); stencil[h][w] may
/// be NULL.  channels 3 or 4.
int devcore_error_correction(int channels, int overlap, int ecMode, int ecDcm, double minDeviationRatio, double minImproveRatio, int nContours,
                             const uint32_t *cbegin, const double *points, const uint8_t *orders, const uint8_t *colors, const double *xf, int w,
                             int h, float *field, uint8_t *stencil) {
    if (ecMode == 0)
        return 0;
    const Prepared pr = prepare(nContours, cbegin, points, orders, colors);
    RawShapes rs = { points, orders, colors, cbegin, pr.windings.data() };
    GlyphDesc g = { };
    g.sx = xf[0]; g.sy = xf[1]; g.tx = xf[2]; g.ty = xf[3];
    g.dmScale = 1/(xf[5]-xf[4]); g.dmTranslate = -xf[4];
    g.w = w; g.h = h;
    g.c0 = 0; g.c1 = (unsigned) nContours;
    g.s0 = 0; g.s1 = cbegin[nContours];
    g.nonEmptyContours = pr.nonEmpty;
    g.recBase = -1;
    switch (channels*2+(overlap != 0)) {
        case 6: runErrorCorrection<3, false>(pr, rs, g, ecMode, ecDcm, minDeviationRatio, minImproveRatio, field, stencil); break;
        case 7: runErrorCorrection<3, true>(pr, rs, g, ecMode, ecDcm, minDeviationRatio, minImproveRatio, field, stencil); break;
        case 8: runErrorCorrection<4, false>(pr, rs, g, ecMode, ecDcm, minDeviationRatio, minImproveRatio, field, stencil); break;
        case 9: runErrorCorrection<4, true>(pr, rs, g, ecMode, ecDcm, minDeviationRatio, minImproveRatio, field, stencil); break;
        default: return -1;
    }
    return 0;
}

/// The filtered path (k1a_filter + k1b_exact; policy 3: the big-shape kernel's single-pass form, 4: the same without the literal
/// redo of texels that hit the guard band — what k1_big_blocks does today) at nPts points inside the w x h
/// texel block of the projection xf = (sx, sy, tx, ty).  sqrtBias: error of the stand-ins for sqrt.approx (in units of 2^-22 relative) and __cosf (in units
/// of 1e-6 absolute): -1, 0, +1.  Otherwise as devcore_distance; survivorsPerTexel receives the mean number of exact edge evaluations per texel.
int devcore_distance_filtered(int mode, int overlap, int policy, int sqrtBias, int nContours, const uint32_t *cbegin, const double *points,
                              const uint8_t *orders, const uint8_t *colors, const double *xf, int w, int h, int nPts, const double *pts, double *out,
                              double *survivorsPerTexel) {
    const Prepared pr = prepare(nContours, cbegin, points, orders, colors);
    GlyphDesc g = { };
    g.sx = xf[0]; g.sy = xf[1]; g.tx = xf[2]; g.ty = xf[3];
    g.w = w; g.h = h;
    g.c0 = 0; g.c1 = (unsigned) nContours;
    g.s0 = 0; g.s1 = cbegin[nContours];
    g.nonEmptyContours = pr.nonEmpty;
    g.recBase = -1;
    b200_shim_sqrt_bias = sqrtBias;
    b200_shim_cos_bias = sqrtBias;          // one knob: both approximate instructions err the same way, then the other way
    int rc = -1;
    switch (mode*2+(overlap != 0)) {
        case 2: rc = runFiltered<1, false>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
        case 3: rc = runFiltered<1, true>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
        case 4: rc = runFiltered<2, false>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
        case 5: rc = runFiltered<2, true>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
        case 6: rc = runFiltered<3, false>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
        case 7: rc = runFiltered<3, true>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
        case 8: rc = runFiltered<4, false>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
        case 9: rc = runFiltered<4, true>(pr, g, cbegin, policy, nPts, pts, out, survivorsPerTexel); break;
    }
    b200_shim_sqrt_bias = b200_shim_cos_bias = 0;
    return rc;
}

/// Raw (unmapped) distances of the shape at nPts points, NOUT doubles per point (1, 1, 3, 4 for mode 1..4 = sdf, psdf, msdf,
/// mtsdf) — the quantity ShapeDistanceFinder::oneShotDistance returns.  Returns the number of points that hit the guard band of
/// the fast root path, or -1 for bad arguments.
int devcore_distance(int mode, int overlap, int policy, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders,
                     const uint8_t *colors, int nPts, const double *pts, double *out) {
    const Prepared pr = prepare(nContours, cbegin, points, orders, colors);
    switch (mode*2+(overlap != 0)) {
        case 2: return run<1, false>(pr, policy, nPts, pts, out);
        case 3: return run<1, true>(pr, policy, nPts, pts, out);
        case 4: return run<2, false>(pr, policy, nPts, pts, out);
        case 5: return run<2, true>(pr, policy, nPts, pts, out);
        case 6: return run<3, false>(pr, policy, nPts, pts, out);
        case 7: return run<3, true>(pr, policy, nPts, pts, out);
        case 8: return run<4, false>(pr, policy, nPts, pts, out);
        case 9: return run<4, true>(pr, policy, nPts, pts, out);
    }
    return -1;
}

/// EdgeSegment::signedDistance of every segment (storage order) at every point: out[(i*nSeg+s)*3 + {0,1,2}] = distance, dot (the
/// orthogonality tie-breaker; evaluated with need = infinity), side (0 param in [0,1], 1 below, 2 above) and the segment's flags in
/// flagsOut[s].  literal != 0: the literal root solver for quadratics.  Returns the number of evaluations that hit the guard band.
int devcore_segment_distances(int literal, int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int nPts,
                              const double *pts, double *out, int *flagsOut) {
    const Prepared pr = prepare(nContours, cbegin, points, orders, nullptr);
    const int nSeg = (int) pr.recs.size();
    int guardHits = 0;
    for (int j = 0; j < nSeg; ++j) {
        const SegRec &rec = pr.recs[(size_t) j];
        if (flagsOut)
            flagsOut[rec.storageIndex] = rec.flags;
        for (int i = 0; i < nPts; ++i) {
            guard_flag() = 0;
            const d2 p = mk(pts[2*i], pts[2*i+1]);
            const SDist r = literal ? signed_distance<true>(rec, p, DBL_MAX) : signed_distance<false>(rec, p, DBL_MAX);
            guardHits += guard_flag();
            double *o = out+3*((size_t) i*nSeg+rec.storageIndex);
            o[0] = r.d; o[1] = r.dot; o[2] = r.side;
        }
    }
    guard_flag() = 0;
    return guardHits;
}

/// solveCubicNormed as the literal path computes it: roots of x^3 + a x^2 + b x + c into x[3]; returns their number.
int devcore_solve_cubic_normed(double *x, double a, double b, double c) {
    const Roots r = solve_cubic_normed(a, b, c);
    x[0] = r.t0; x[1] = r.t1; x[2] = r.t2;
    return r.n;
}

/// Contour::winding() as the device computes it (k0_windings).
int devcore_windings(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int *out) {
    const Prepared pr = prepare(nContours, cbegin, points, orders, nullptr);
    for (int c = 0; c < nContours; ++c)
        out[c] = pr.windings[(size_t) c];
    return 0;
}

/// How often the literal closed-form cubic solver ran since the last call.
unsigned long long devcore_exact_root_calls() {
    const unsigned long long n = g_exactRootCalls;
    g_exactRootCalls = 0;
    return n;
}

}
