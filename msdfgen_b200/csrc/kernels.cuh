// kernels.cuh — the sm_100a kernels of the hot path and their launch parameters.
//
//   k0_windings          one thread per contour: Contour::winding()  (/root/reference/core/Contour.cpp:57-81)
//   k0_prepare_big       one thread per segment of shapes too large for shared memory: SegRec (+ compact cull records) -> global
//   k1_distance_field    generateDistanceField<Combiner<Selector>>    (/root/reference/core/msdfgen.cpp:52-76)
//   k1_big_blocks        the same for shapes with more than SMEM_RECORDS segments: 32x32 blocks, two-level record lists
//   k2_classify / k3_exact_distance / k4_quantize   msdfErrorCorrectionInner<N> + apply, 8-bit output  (error_correction.cuh)
//   k5_* (scanline.cuh) distanceSignCorrection / rasterize;  k6_* (render.cuh) renderSDF / simulate8bit
//
// Work decomposition: a tile is up to TILE_PIXELS consecutive output texels of one glyph; persistent CTAs of CTA_THREADS
// threads pull tiles from an atomic counter (greedy list scheduling, most expensive tiles first: glyph cost varies 8x with its
// segment count), stage
// the glyph's prepared segment records in shared memory once per tile, then every thread owns one texel per pass and walks
// all records (broadcast shared-memory reads, no divergence on the segment type).  Texels are staged in shared memory and
// written back as coalesced float4 rows.
#pragma once

#include "distance_core.cuh"
#include "distance_filter.cuh"

namespace b200 {

#define CTA_THREADS 256
#define TILE_PIXELS 1024
#define SMEM_RECORDS 96

struct TileDesc { unsigned glyph, pix0; };

// Compact copies of the fields the tile-level culling of big shapes reads (48 B each instead of a 352 B SegRec)
struct CullRec { d2 bc; double br; int color, contour; d2 p0; };
struct PerpRec { d2 pEnd, aDir, bDir, aBis, bBis; };

struct KernelParams {
    RawShapes raw;
    const GlyphDesc *glyphs;
    const TileDesc *tiles;
    int nTiles;                   // tiles [0, nSmallTiles) belong to glyphs staged in shared memory, the rest to big shapes
    int nSmallTiles;
    int tileBegin, tileEnd;       // tile range of this launch
    int counterIndex;             // which counter this launch pulls tiles from
    int nContours;
    const SegRec *bigRecs;        // prepared records of big shapes
    const CullRec *bigCull;       // same indexing as bigRecs
    const PerpRec *bigPerp;
    const FRec *bigFrecs;         // fp32 filter records of big shapes (same indexing as bigRecs)
    unsigned *bigExtent;          // per big glyph: float bits of the largest |local coordinate| of its control points
    int *counters;                // tile counters: [0],[3] k1 small/big, [1] k2, [2],[4] k3 small/big
    struct Candidate *candidates; // k2 -> k3: per-tile slots of exact-distance checks, candCap per tile
    int *candCount;               // per-tile number of pushes (may exceed candCap: overflow texels carry EC_FLAG_RECHECK)
    int candCap;
    float *field;                 // k1 output (k2 input)
    float *out;                   // k2 output
    uint8_t *stencil;             // k2: per-texel flags (1 = ERROR, 2 = PROTECTED), rows in Y-up order like the reference's buffer
    // big shapes (> SMEM_RECORDS segments): 2-D blocks of BLOCK_W x BLOCK_H texels with per-block record lists
    const TileDesc *blocks;       // {glyph, block index (row-major over the glyph's block grid)}
    int nBlocks;
    int *blockList;               // per-CTA workspace: listStride record indices, then the group ends
    unsigned long long *blockCaps;// per-CTA workspace: capStride squared distances (as ordered bit patterns)
    size_t listStride, listPerCta, capStride;   // list capacity; ints per CTA (list + group ends); cap slots per CTA
    double *scratch;              // per-thread contour-distance slices for many-contour shapes
    signed char *scratchW;        // k1_big_blocks: per-thread winding slices (same indexing)
    size_t scratchStride;         // = gridDim.x * CTA_THREADS
    unsigned *recheckMask;        // B200_GUARD_MODE 2: k1b -> recheck launch, one bit per texel of the batch
    unsigned char *recheckTiles;  // one byte per small tile: the tile has such texels
    int *recheckAny;              // any texel flagged at all (the recheck launch returns at once otherwise)
    unsigned *filt;               // k1a -> k1b: per-unit filter words (distance_filter.cuh)
    const unsigned long long *tileFilt;   // word offset of every small tile's first unit in filt
    // error correction config
    int ecMode, ecDcm;
    double minDeviationRatio, minImproveRatio;
    int shapeless;                // msdfFast*ErrorCorrection: no shape, no corner protection, base classifier only
};

// ---------------------------------------------------------------------------------------------------------------------

__global__ void k0_windings(RawShapes raw, int nContours, int8_t *windings, int *counters) {
    int c = blockIdx.x*blockDim.x+threadIdx.x;
    if (c == 0) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            counters[i] = 0;
    }
    if (c < nContours)
        windings[c] = (int8_t) contour_winding(raw, (unsigned) c);
}

// items: (glyph index, evaluation position) pairs are implied by a prefix table: bigGlyphs[i] = glyph index, bigBase[i] = first record
__global__ void k0_prepare_big(RawShapes raw, const GlyphDesc *glyphs, const int *bigGlyphs, int nBig, SegRec *recs, CullRec *cull, PerpRec *perp,
                               unsigned *extent) {
    for (int i = blockIdx.y; i < nBig; i += gridDim.y) {
        GlyphDesc g = glyphs[bigGlyphs[i]];
        unsigned n = g.s1-g.s0;
        const d2 O = filter_origin(g);
        for (unsigned j = blockIdx.x*blockDim.x+threadIdx.x; j < n; j += gridDim.x*blockDim.x) {
            SegRec &r = recs[size_t(g.recBase)+j];
            prepare_record(raw, g, j, r);
            atomicMax(&extent[i], __float_as_uint(record_extent(r, O)));
            CullRec c;
            c.bc = r.bc; c.br = r.br; c.color = r.color; c.contour = r.contour; c.p0 = r.p0;
            cull[size_t(g.recBase)+j] = c;
            PerpRec q;
            q.pEnd = r.pEnd; q.aDir = r.aDir; q.bDir = r.bDir; q.aBis = r.aBis; q.bBis = r.bBis;
            perp[size_t(g.recBase)+j] = q;
        }
    }
}

// second preparation pass of big shapes: the filter records (need the shape's extent, known after k0_prepare_big)
__global__ void k0_prepare_big_filter(const GlyphDesc *glyphs, const int *bigGlyphs, int nBig, const SegRec *recs, FRec *frecs, const unsigned *extent,
                                      int singleChannel) {
    for (int i = blockIdx.y; i < nBig; i += gridDim.y) {
        GlyphDesc g = glyphs[bigGlyphs[i]];
        unsigned n = g.s1-g.s0;
        const d2 O = filter_origin(g);
        const float slack = filter_slack(g, O, __uint_as_float(extent[i]));
        for (unsigned j = blockIdx.x*blockDim.x+threadIdx.x; j < n; j += gridDim.x*blockDim.x) {
            FRec f;
            make_filter_record<3>(recs[size_t(g.recBase)+j], O, slack, f);
            if (singleChannel)
                f.mask = (f.mask&~7)|1;
            frecs[size_t(g.recBase)+j] = f;
        }
    }
}

// Fetches the next tile for the CTA (uniform) and (re)stages the glyph's records when the glyph changes.
struct TileCursor {
    int tile;
    int glyph;
};

// ---- TMA (bulk async copy) staging of a glyph's raw control points ---------------------------------------------------------
// One elected thread issues cp.async.bulk (global -> shared, completion counted in bytes on an mbarrier); the CTA waits on
// the barrier's phase and then prepares the records from shared memory.  SASS: UBLKCP + SYNCS.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void tma_load_1d(void *dstSmem, const void *srcGlobal, unsigned bytes, unsigned long long *mbar) {
    unsigned bar = smem_addr(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, unsigned phase) {
    unsigned bar = smem_addr(mbar);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
}

// Small glyphs (<= SMEM_RECORDS segments): the raw control points arrive in shared memory by one TMA bulk copy, then every
// thread of the CTA prepares records from them.  `phase` is the CTA-uniform parity of the mbarrier (flipped per use).
__device__ __forceinline__ void stage_glyph(const KernelParams &P, const GlyphDesc &g, SegRec *smemRecs, double *smemPoints,
                                            unsigned long long *mbar, unsigned &phase) {
    unsigned n = g.s1-g.s0;
    if (n == 0)
        return;
    if (threadIdx.x == 0)
        tma_load_1d(smemPoints, P.raw.points+8*size_t(g.s0), n*64u, mbar);
    mbar_wait(mbar, phase);
    phase ^= 1u;
    RawShapes rs = P.raw;
    rs.points = smemPoints-8*size_t(g.s0);       // prepare_record indexes segments by their global index
    for (unsigned j = threadIdx.x; j < n; j += blockDim.x)
        prepare_record(rs, g, j, smemRecs[j]);
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

// The full walk (every texel evaluates every record with the exact fp64 culling of sel_add_edge): the distance kernel of
// round 1, kept for two jobs.  RECHECK = false: A/B baseline (B200MSDF_K1_LEGACY=1).  RECHECK = true (B200_GUARD_MODE 2): launched
// after k1b_exact, it recomputes — with the LITERAL root solver for every quadratic segment — exactly the texels k1b flagged
// (guard band of the fast root path) and overwrites them; it returns at once when nothing was flagged.
template <int MODE, bool OVERLAP, bool RECHECK>
__global__ void __launch_bounds__(CTA_THREADS, 2) k1_distance_field(KernelParams P) {
    const bool BIG = false;
    if (RECHECK && !*P.recheckAny)
        return;
    const int NOUT = ModeTraits<MODE>::NOUT;
    __shared__ SegRec s_recs[BIG ? 1 : SMEM_RECORDS];
    __shared__ __align__(16) double s_points[BIG ? 2 : SMEM_RECORDS*8];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ __align__(16) float s_out[CTA_THREADS*NOUT];
    __shared__ GlyphDesc s_g;
    __shared__ int s_tile;
    unsigned phase = 0;
    if (!BIG && threadIdx.x == 0)
        mbar_init(&s_mbar);

    double localStore[OVERLAP ? B200_LOCAL_CONTOURS*StoreLayout<MODE>::PER_CONTOUR : 1];
    int stagedGlyph = -1;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            s_tile = P.tileBegin+atomicAdd(&P.counters[P.counterIndex], 1);
        __syncthreads();
        int tile = s_tile;
        if (tile >= P.tileEnd)
            break;
        if (RECHECK && !P.recheckTiles[tile])
            continue;
        TileDesc td = P.tiles[tile];
        if ((int) td.glyph != stagedGlyph) {
            if (threadIdx.x == 0)
                s_g = P.glyphs[td.glyph];
            __syncthreads();
            if (!BIG)
                stage_glyph(P, s_g, s_recs, s_points, &s_mbar, phase);
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const SegRec *recs = BIG ? P.bigRecs+s_g.recBase : s_recs;
        const int w = s_g.w;
        const unsigned total = unsigned(w)*unsigned(s_g.h);
        const int nRec = int(s_g.s1-s_g.s0);
        unsigned tileEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
        ContourStore store;
        store.perContour = StoreLayout<MODE>::PER_CONTOUR;
        if (OVERLAP && s_g.nonEmptyContours > B200_LOCAL_CONTOURS) {
            store.base = P.scratch+(size_t(blockIdx.x)*CTA_THREADS+threadIdx.x);
            store.stride = P.scratchStride;
        } else {
            store.base = localStore;
            store.stride = 1;
        }
        // Thread -> texel mapping inside a pass of CTA_THREADS consecutive texels.  When the pass is a whole number of rows
        // (w in {8,16,32,64}) a warp takes an 8x4 block instead of a 32x1 strip: neighbouring lanes then see nearly the same
        // set of relevant edges, so the culling branches and the segment kernels' branches diverge less.
        const bool blocked = (w&7) == 0 && w <= 64 && (CTA_THREADS%w) == 0 && ((CTA_THREADS/w)&3) == 0;
        for (unsigned passBase = td.pix0; passBase < tileEnd; passBase += CTA_THREADS) {
            unsigned slot = threadIdx.x;        // position of this thread's texel inside the pass (row-major)
            if (blocked && passBase+CTA_THREADS <= tileEnd) {
                const unsigned warp = threadIdx.x>>5, lane = threadIdx.x&31;
                const unsigned blocksAcross = unsigned(w)>>3;
                unsigned bx = warp%blocksAcross, by = warp/blocksAcross;
                slot = (by*4+(lane>>3))*unsigned(w)+bx*8+(lane&7);
            }
            unsigned i = passBase+slot;
            bool mine = i < tileEnd;
            if (RECHECK && mine) {
                const size_t t = size_t(s_g.texelOffset)+i;
                mine = (P.recheckMask[t>>5]>>(t&31))&1u;
            }
            if (mine) {
                int x = int(i%unsigned(w));
                int yOut = int(i/unsigned(w));
                int y = s_g.flip ? s_g.h-1-yOut : yOut;
                // Projection::unproject(Point2(x+.5, y+.5))
                d2 p = mk((s_g.x0+x+.5)/s_g.sx-s_g.tx, (s_g.y0+y+.5)/s_g.sy-s_g.ty);
                double dist[NOUT];
                shape_distance<MODE, OVERLAP, RECHECK>(recs, nRec, p, store, dist);
                if (RECHECK) {
                    #pragma unroll
                    for (int k = 0; k < NOUT; ++k)
                        P.field[s_g.outOffset+size_t(i)*NOUT+k] = float(s_g.dmScale*(dist[k]+s_g.dmTranslate));
                } else {
                    #pragma unroll
                    for (int k = 0; k < NOUT; ++k)
                        s_out[slot*NOUT+k] = float(s_g.dmScale*(dist[k]+s_g.dmTranslate));    // DistanceMapping::operator()
                }
            }
            if (!RECHECK) {
                __syncthreads();
                unsigned nPix = tileEnd-passBase < CTA_THREADS ? tileEnd-passBase : CTA_THREADS;
                store_tile(P.field+s_g.outOffset+size_t(passBase)*NOUT, s_out, int(nPix)*NOUT);
                __syncthreads();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k1a_filter + k1b_exact — the distance kernels of glyphs whose records fit in shared memory (<= SMEM_RECORDS segments).
//
// Both are persistent 128-thread CTAs pulling 1024-texel tiles from an atomic counter.  Per glyph change one elected thread
// TMA-copies the raw control points into shared memory and the CTA prepares what the kernel needs there: k1a the fp32 filter
// records (FRec), k1b the SegRecs (fp64 constants of the exact evaluation).  Inside a tile there is NO CTA barrier: every warp
// pulls 32-texel units from a shared-memory counter (8x4 texel blocks when the tile is a whole number of rows of a width that
// is a multiple of 8 — neighbouring lanes then keep nearly the same edges —, 32x1 strips otherwise).
//   k1a_filter  passes 1 and 2 of distance_filter.cuh: fp32, ~40 registers, a loop body that fits the L0 instruction cache;
//               writes the unit's survivor / candidate masks and upper bounds (interleaved words, whole 128-byte lines).
//   k1b_exact   pass 3 + contour combiner in fp64 for what the filter kept; stages its 32 texels in the warp's shared-memory
//               slice and writes them back as float4 rows.
#define K1_THREADS 128
#define K1_WARPS (K1_THREADS/32)

// Unit geometry of a tile (uniform across the CTA).
struct UnitGrid {
    bool blocked;
    int row0, nRows, blocksAcross, nUnits;
    unsigned pixEnd;
    __device__ __forceinline__ UnitGrid(const GlyphDesc &g, const TileDesc &td) {
        const int w = g.w;
        const unsigned total = unsigned(w)*unsigned(g.h);
        pixEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
        blocked = (w&7) == 0 && w <= TILE_PIXELS && (TILE_PIXELS%w) == 0;
        row0 = blocked ? int(td.pix0/unsigned(w)) : 0;
        nRows = blocked ? int((pixEnd-td.pix0)/unsigned(w)) : 0;
        blocksAcross = w>>3;
        nUnits = blocked ? blocksAcross*((nRows+3)>>2) : int((pixEnd-td.pix0+31u)>>5);
    }
    // texel of (unit, lane): returns validity; idle lanes shadow a texel of the unit (the exact walk votes warp-wide)
    __device__ __forceinline__ bool texel(const GlyphDesc &g, const TileDesc &td, int u, unsigned lane, int &x, int &yOut) const {
        bool valid;
        if (blocked) {
            const int bx = u%blocksAcross, by = u/blocksAcross;
            x = bx*8+int(lane&7);
            yOut = row0+by*4+int(lane>>3);
            valid = yOut < row0+nRows;
            if (!valid)
                yOut = row0+nRows-1;
        } else {
            unsigned i = td.pix0+unsigned(u)*32u+lane;
            valid = i < pixEnd;
            if (!valid)
                i = pixEnd-1;
            x = int(i%unsigned(g.w));
            yOut = int(i/unsigned(g.w));
        }
        return valid;
    }
};

// Projection::unproject(Point2(x+.5, y+.5)) of output texel (x, yOut)
__device__ __forceinline__ d2 texel_point(const GlyphDesc &g, int x, int yOut) {
    const int y = g.flip ? g.h-1-yOut : yOut;
    return mk(ddiv(g.x0+x+.5, g.sx)-g.tx, ddiv(g.y0+y+.5, g.sy)-g.ty);
}

// group ends (positions after the last record of every non-empty contour) and filter words per texel; one thread
template <int MODE, bool OVERLAP>
__device__ __forceinline__ void build_groups(const KernelParams &P, const GlyphDesc &g, int *groupEnd, int *nGroups, unsigned *wordsPerTexel) {
    int ng = 0;
    unsigned words = 0;
    for (unsigned c = g.c0; c < g.c1; ++c) {
        const unsigned b = P.raw.cbegin[c], e = P.raw.cbegin[c+1];
        if (e > b) {
            groupEnd[ng++] = int(e-g.s0);
            words += filter_words_group<MODE>(e-b);
        }
    }
    if (!OVERLAP)
        words = g.s1 > g.s0 ? filter_words_group<MODE>(g.s1-g.s0) : 0u;
    *nGroups = ng;
    *wordsPerTexel = words;
}

template <int MODE, bool OVERLAP>
__global__ void __launch_bounds__(K1_THREADS, 8) k1a_filter(KernelParams P) {
    __shared__ __align__(16) double s_points[SMEM_RECORDS*8];
    __shared__ __align__(16) FRec s_frecs[SMEM_RECORDS];
    __shared__ int s_groupEnd[SMEM_RECORDS];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ GlyphDesc s_g;
    __shared__ float s_slack;
    __shared__ int s_nGroups, s_tile, s_unit;
    __shared__ unsigned s_words, s_extent;
    unsigned phase = 0;
    if (threadIdx.x == 0)
        mbar_init(&s_mbar);
    const unsigned lane = threadIdx.x&31;
    int stagedGlyph = -1;
    for (;;) {
        __syncthreads();        // every warp has left the previous tile
        if (threadIdx.x == 0) {
            s_tile = P.tileBegin+atomicAdd(&P.counters[P.counterIndex], 1);
            s_unit = 0;
        }
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.tileEnd)
            break;
        const TileDesc td = P.tiles[tile];
        if ((int) td.glyph != stagedGlyph) {
            if (threadIdx.x == 0) {
                s_g = P.glyphs[td.glyph];
                s_extent = 0;
            }
            __syncthreads();
            const unsigned n = s_g.s1-s_g.s0;
            const d2 O = filter_origin(s_g);
            if (n) {
                if (threadIdx.x == 0)
                    tma_load_1d(s_points, P.raw.points+8*size_t(s_g.s0), n*64u, &s_mbar);
                mbar_wait(&s_mbar, phase);
                phase ^= 1u;
                // M: largest |local coordinate| of any control point
                for (unsigned j = threadIdx.x; j < n; j += blockDim.x) {
                    const int order = P.raw.orders[s_g.s0+j];
                    double m = 0;
                    for (int i = 0; i <= order; ++i)
                        m = fmax(m, fmax(fabs(s_points[8*j+2*i]-O.x), fabs(s_points[8*j+2*i+1]-O.y)));
                    atomicMax(&s_extent, __float_as_uint(m == m ? float_up(m) : __int_as_float(0x7f800000)));
                }
            }
            if (threadIdx.x == 0)
                build_groups<MODE, OVERLAP>(P, s_g, s_groupEnd, &s_nGroups, &s_words);
            __syncthreads();
            float M = __uint_as_float(s_extent);
            M = fmaxf(M, float_up((.5*s_g.w+1)/fabs(s_g.sx)));
            M = fmaxf(M, float_up((.5*s_g.h+1)/fabs(s_g.sy)));
            const float slack = M*1.52587890625e-5f+float_up(fmax(fabs(O.x), fabs(O.y)))*5.7e-14f+1e-37f;       // 2^-16 M + 2^-44 |O|
            if (threadIdx.x == 0)
                s_slack = slack;
            RawShapes rs = P.raw;
            rs.points = s_points-8*size_t(s_g.s0);          // prepare_record indexes segments by their global index
            for (unsigned j = threadIdx.x; j < n; j += blockDim.x) {
                SegRec tmp;
                prepare_record(rs, s_g, j, tmp);
                make_filter_record<MODE>(tmp, O, slack, s_frecs[j]);
            }
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const UnitGrid grid(s_g, td);
        const int nRec = int(s_g.s1-s_g.s0);
        const d2 O = filter_origin(s_g);
        unsigned *const tileWords = P.filt+P.tileFilt[tile];
        const unsigned words = s_words;
        for (;;) {
            int u = 0;
            if (lane == 0)
                u = atomicAdd(&s_unit, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= grid.nUnits)
                break;
            int x, yOut;
            grid.texel(s_g, td, u, lane, x, yOut);
            const d2 p = texel_point(s_g, x, yOut);
            filter_texel<MODE, OVERLAP>(s_frecs, nRec, s_groupEnd, s_nGroups, s_slack, float(p.x-O.x), float(p.y-O.y),
                                        tileWords+size_t(u)*words*32u+lane);
        }
    }
}

// How texels whose fast-path roots hit the guard band (distance_core.cuh: quad_root) get their literal re-evaluation:
//   0  no guard band;  1  in k1b_exact, by an out-of-line call per flagged texel;  2  by a second launch of the full-walk kernel over
//   the flagged texels (k1b only records them).  Measured on B200 (headline atlas, k1a+k1b): see profiles/README.md.
#ifndef B200_GUARD_MODE
#define B200_GUARD_MODE 2
#endif

// ShapeDistanceFinder::oneShotDistance by the full walk with LITERAL roots (distance_core.cuh), mapped and stored as floats.
template <int MODE, bool OVERLAP>
__device__ __noinline__ void texel_literal(const SegRec *recs, int nRec, double px, double py, ContourStore store, double dmScale, double dmTranslate,
                                           float *out) {
    double dist[ModeTraits<MODE>::NOUT];
    shape_distance<MODE, OVERLAP, true>(recs, nRec, mk(px, py), store, dist);
    #pragma unroll
    for (int k = 0; k < ModeTraits<MODE>::NOUT; ++k)
        out[k] = float(dmScale*(dist[k]+dmTranslate));
}

#ifndef K1B_MIN_CTAS
#define K1B_MIN_CTAS 4
#endif
template <int MODE, bool OVERLAP>
__global__ void __launch_bounds__(K1_THREADS, K1B_MIN_CTAS) k1b_exact(KernelParams P) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    __shared__ SegRec s_recs[SMEM_RECORDS];
    __shared__ __align__(16) double s_points[SMEM_RECORDS*8];
    __shared__ int s_groupEnd[SMEM_RECORDS];
    __shared__ __align__(16) float s_out[K1_WARPS][32*NOUT];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ GlyphDesc s_g;
    __shared__ int s_nGroups, s_tile, s_unit;
    __shared__ unsigned s_words;
    unsigned phase = 0;
    if (threadIdx.x == 0)
        mbar_init(&s_mbar);
    double localStore[OVERLAP ? B200_LOCAL_CONTOURS*StoreLayout<MODE>::PER_CONTOUR : 1];
    const unsigned lane = threadIdx.x&31, warp = threadIdx.x>>5;
    guard_flag() = 0;
    int stagedGlyph = -1;
    for (;;) {
        __syncthreads();        // every warp has left the previous tile
        if (threadIdx.x == 0) {
            s_tile = P.tileBegin+atomicAdd(&P.counters[P.counterIndex], 1);
            s_unit = 0;
        }
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.tileEnd)
            break;
        const TileDesc td = P.tiles[tile];
        if ((int) td.glyph != stagedGlyph) {
            if (threadIdx.x == 0)
                s_g = P.glyphs[td.glyph];
            __syncthreads();
            stage_glyph(P, s_g, s_recs, s_points, &s_mbar, phase);
            if (threadIdx.x == 0)
                build_groups<MODE, OVERLAP>(P, s_g, s_groupEnd, &s_nGroups, &s_words);
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const UnitGrid grid(s_g, td);
        const int w = s_g.w;
        const int nRec = int(s_g.s1-s_g.s0);
        ContourStore store;
        store.perContour = StoreLayout<MODE>::PER_CONTOUR;
        if (OVERLAP && s_g.nonEmptyContours > B200_LOCAL_CONTOURS) {
            store.base = P.scratch+(size_t(blockIdx.x)*K1_THREADS+threadIdx.x);
            store.stride = P.scratchStride;
        } else {
            store.base = localStore;
            store.stride = 1;
        }
        float *const tileOut = P.field+s_g.outOffset;
        const unsigned *const tileWords = P.filt+P.tileFilt[tile];
        const unsigned words = s_words;
        for (;;) {
            int u = 0;
            if (lane == 0)
                u = atomicAdd(&s_unit, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= grid.nUnits)
                break;
            int x, yOut;
            const bool valid = grid.texel(s_g, td, u, lane, x, yOut);
            const d2 p = texel_point(s_g, x, yOut);
            double dist[NOUT];
            exact_texel<MODE, OVERLAP>(s_recs, nRec, s_groupEnd, s_nGroups, tileWords+size_t(u)*words*32u+lane, p, store, dist);
            if (valid) {
                #pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    s_out[warp][lane*NOUT+k] = float(s_g.dmScale*(dist[k]+s_g.dmTranslate));      // DistanceMapping::operator()
            }
#if B200_GUARD_MODE == 1
            if (guard_flag()) {
                // guard band of the fast foot-point roots hit (a handful of texels per atlas): the texel is recomputed by the full
                // walk with the literal, correctly rounded root solver for every quadratic segment — out of line, cold
                guard_flag() = 0;
                if (valid)
                    texel_literal<MODE, OVERLAP>(s_recs, nRec, p.x, p.y, store, s_g.dmScale, s_g.dmTranslate, &s_out[warp][lane*NOUT]);
            }
#endif
            __syncwarp();
            if (grid.blocked) {
                // 4 rows of 8 texels: 2*NOUT float4 per row, rows contiguous in the warp's slice
                const int by = u/grid.blocksAcross, bx = u%grid.blocksAcross;
                const int r = int(lane)/(2*NOUT), q = int(lane)%(2*NOUT);
                if (r < 4 && by*4+r < grid.nRows) {
                    float4 *dst = reinterpret_cast<float4 *>(tileOut+(size_t(grid.row0+by*4+r)*w+bx*8)*NOUT)+q;
                    *dst = reinterpret_cast<const float4 *>(s_out[warp])[lane];
                }
            } else {
                const unsigned i0 = td.pix0+unsigned(u)*32u;
                const int nFloats = int((grid.pixEnd-i0 < 32u ? grid.pixEnd-i0 : 32u))*NOUT;
                float *dst = tileOut+size_t(i0)*NOUT;
                for (int i = int(lane); i < (nFloats>>2); i += 32)
                    reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(s_out[warp])[i];
                for (int i = (nFloats&~3)+int(lane); i < nFloats; i += 32)
                    dst[i] = s_out[warp][i];
            }
            __syncwarp();
#if B200_GUARD_MODE == 2
            // guard band hit by some lane (rare): record the texel for the recheck launch (k1_distance_field<.., true>), which
            // recomputes it with the literal solver.  The texel is re-derived from (unit, lane) here, so that nothing extra has
            // to stay live across exact_texel (keeping x, yOut and the tile index alive there cost 0.2 ms per atlas).
            if (__any_sync(0xffffffffu, guard_flag() != 0) && guard_flag()) {
                guard_flag() = 0;
                int fx, fy;
                if (grid.texel(s_g, td, u, lane, fx, fy)) {
                    const size_t t = size_t(s_g.texelOffset)+size_t(fy)*s_g.w+fx;
                    atomicOr(&P.recheckMask[t>>5], 1u<<(t&31));
                    P.recheckTiles[s_tile] = 1;
                    *P.recheckAny = 1;
                }
            }
#endif
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Big shapes: two-level tile culling.
//
// A shape with thousands of segments (BASELINE configs[3]: 10 000) makes every texel walk every record, although the exact
// per-texel culling of sel_add_edge then skips the math of nearly all of them.  Here a tile is a BLOCK_W x BLOCK_H block of
// texels.  Level 1: the CTA builds the ordered list of records that can matter ANYWHERE in the block.  Level 2: each warp
// refines that list for the 8x4 sub-block it evaluates in the current pass (one lane tests one record, so the refinement
// costs 1/32 of a walk).  The texels walk only the sub-list (ListIter).  A record is dropped for a region R (block or
// sub-block) only if, for every texel p of R, sel_add_edge would take its `culled` path AND reject both perpendicular
// candidates, i.e. it provably has no effect:
//
//   c = centre of R, hd = half diagonal of R's texel centres (so |p - c| <= hd), and per contour (OVERLAP) or per shape
//   (simple combiner) and channel k:  capT[k] = min over the candidate records with channel k of |c - p0| + hd.
//   For every texel, cap[k](p) <= capT[k]: the record attaining the minimum is itself kept (its circle contains p0), and
//   |p - p0| <= |c - p0| + hd.  With needT = max over the record's channels of capT[k] >= need(p):
//     * |c - bc| - hd - br > needT   =>  |p - bc| > br + need(p)                     => culled at p;
//     * | cross(c - p0, aDir) | - hd > needT  =>  |pdA(p)| > need(p)                 => candidate A rejected (aDir is a unit
//       vector or zero, so pdA is 1-Lipschitz); same for B with pEnd, bDir.  A candidate is also rejected when one of its
//       domain tests (dot(ap, aBis) > 0, -dot(ap, aDir) > 0; B: -dot(bp, bBis) > 0, dot(bp, bDir) > 0) fails by more than
//       hd at c, since then it fails at every p of R.
//   All comparisons carry a relative slack of 1e-8 (rounding is ~1e-16) and NaNs compare as "keep".
// The lists preserve the evaluation order, so ties resolve exactly as in the full walk.  Results are bit-identical.
#define BLOCK_W 32
#define BLOCK_H 32
#define SUB_CAP 3072        // record indices per warp sub-list (dynamic shared memory); overflow -> the warp walks the block list
#define SUB_GROUPS 256      // contour groups per warp sub-list

template <int MODE>
__device__ __forceinline__ bool region_keeps(const CullRec &cr, const PerpRec *perp, int j, d2 c, double hd, double needT) {
    d2 v = c-cr.bc;
    const double reach = needT+cr.br+hd;
    bool keep = !(dot(v, v) > reach*reach*(1+1e-8));
    if (MODE != 1 && !keep) {
        const double lim = needT+hd;
        // a candidate also has to pass the domain tests of sel_add_edge somewhere in the region: dot(ap, aBis) > 0 and
        // -dot(ap, aDir) > 0 (B: -dot(bp, bBis) > 0 and dot(bp, bDir) > 0); all four are 1-Lipschitz in p
        const PerpRec q = perp[j];
        const double hs = hd*(1+1e-8);
        const d2 ap = c-cr.p0, bp = c-q.pEnd;
        const bool a = !(fabs(cross(ap, q.aDir)) > lim) && !(dot(ap, q.aBis) < -hs) && !(dot(ap, q.aDir) > hs);
        const bool b = !(fabs(cross(bp, q.bDir)) > lim) && !(dot(bp, q.bBis) > hs) && !(dot(bp, q.bDir) < -hs);
        keep = a || b;
    }
    return keep;
}

template <int MODE, bool OVERLAP>
__global__ void __launch_bounds__(CTA_THREADS, 2) k1_big_blocks(KernelParams P) {
    const int NOUT = ModeTraits<MODE>::NOUT;
    const int NCH = ModeTraits<MODE>::NCH;
    const int ROWS = CTA_THREADS/BLOCK_W;           // texel rows per pass
    const int WARPS = CTA_THREADS/32;
    __shared__ __align__(16) float s_out[CTA_THREADS*NOUT];
    __shared__ GlyphDesc s_g;
    __shared__ int s_tile;
    __shared__ int s_warpKeep[WARPS];
    extern __shared__ int s_sub[];          // WARPS x SUB_CAP
    __shared__ int s_subEnd[OVERLAP ? WARPS : 1][OVERLAP ? SUB_GROUPS : 1];
    int *list = P.blockList+size_t(blockIdx.x)*P.listPerCta;
    unsigned long long *caps = P.blockCaps+size_t(blockIdx.x)*P.capStride;
    double localCd[OVERLAP ? B200_LOCAL_CONTOURS*NOUT : 1];
    signed char localW[OVERLAP ? B200_LOCAL_CONTOURS : 1];
    const unsigned lane = threadIdx.x&31, warp = threadIdx.x>>5;
    int stagedGlyph = -1;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            s_tile = atomicAdd(&P.counters[P.counterIndex], 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.nBlocks)
            break;
        const TileDesc td = P.blocks[tile];
        if ((int) td.glyph != stagedGlyph) {
            if (threadIdx.x == 0)
                s_g = P.glyphs[td.glyph];
            stagedGlyph = (int) td.glyph;
            __syncthreads();
        }
        const SegRec *recs = P.bigRecs+s_g.recBase;
        const CullRec *cull = P.bigCull+s_g.recBase;
        const PerpRec *perp = P.bigPerp+s_g.recBase;
        const FRec *frecs = P.bigFrecs+s_g.recBase;
        const d2 O = filter_origin(s_g);
        const float fslack = filter_slack(s_g, O, __uint_as_float(P.bigExtent[s_g.bigIndex]));
        const int nRec = int(s_g.s1-s_g.s0);
        const int w = s_g.w, h = s_g.h;
        const int blocksX = (w+BLOCK_W-1)/BLOCK_W;
        const int bx0 = int(td.pix0%unsigned(blocksX))*BLOCK_W, by0 = int(td.pix0/unsigned(blocksX))*BLOCK_H;
        const double slack = 1e-12*(fabs(s_g.tx)+fabs(s_g.ty));
        // block centre and half diagonal in shape units (texel centres span +-(BLOCK-1)/2 pixels around the centre)
        const double cpx = s_g.x0+bx0+.5*BLOCK_W;
        const double cpy = s_g.y0+(s_g.flip ? h-by0-.5*BLOCK_H : by0+.5*BLOCK_H);
        const d2 c = mk(cpx/s_g.sx-s_g.tx, cpy/s_g.sy-s_g.ty);
        const double hx = .5*(BLOCK_W-1)/fabs(s_g.sx), hy = .5*(BLOCK_H-1)/fabs(s_g.sy);
        const double hd = sqrt(hx*hx+hy*hy)*(1+1e-9)+1e-12*(fabs(c.x)+fabs(c.y))+slack;

        // ---- level 1, phase 1: capT per (contour | shape, channel)
        const int nGroups = OVERLAP ? int(s_g.c1-s_g.c0) : 1;
        for (int i = threadIdx.x; i < nGroups*NCH; i += CTA_THREADS)
            caps[i] = ~0ull;
        __syncthreads();
        for (int j = threadIdx.x; j < nRec; j += CTA_THREADS) {
            const CullRec cr = cull[j];
            const int mask = (MODE == 1 || MODE == 2) ? 1 : cr.color;
            const int grp = OVERLAP ? cr.contour : 0;
            d2 v = c-cr.p0;
            unsigned long long q = (unsigned long long) __double_as_longlong(dot(v, v));      // >= 0: bit patterns order like values
            #pragma unroll
            for (int k = 0; k < NCH; ++k)
                if (mask&(1<<k))
                    atomicMin(&caps[grp*NCH+k], q);
        }
        __syncthreads();
        // ---- level 1, phase 2: ordered list of the records that can matter in this block
        int nList = 0;
        for (int base = 0; base < nRec; base += CTA_THREADS) {
            const int j = base+int(threadIdx.x);
            bool keep = false;
            if (j < nRec) {
                const CullRec cr = cull[j];
                const int mask = (MODE == 1 || MODE == 2) ? 1 : cr.color;
                const int grp = OVERLAP ? cr.contour : 0;
                double needT = 0;
                #pragma unroll
                for (int k = 0; k < NCH; ++k)
                    if (mask&(1<<k))
                        needT = dmax(needT, sqrt(__longlong_as_double((long long) caps[grp*NCH+k])));
                needT = (needT+hd)*(1+1e-8);
                keep = region_keeps<MODE>(cr, perp, j, c, hd, needT);
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (lane == 0)
                s_warpKeep[warp] = __popc(ballot);
            __syncthreads();
            int before = 0, total = 0;
            #pragma unroll
            for (int k = 0; k < WARPS; ++k) {
                const int n = s_warpKeep[k];
                if (k < int(warp)) before += n;
                total += n;
            }
            if (keep)
                list[nList+before+__popc(ballot&((1u<<lane)-1u))] = j;
            nList += total;
            __syncthreads();
        }
        // (the list stores above are ordered before the reads below by the last barrier of the loop)
        // group ends: position i closes a contour's run when the next listed record belongs to another contour; the runs are
        // numbered by counting the closing positions before them (ordered compaction again, 256 positions per round)
        int *groupEnd = list+P.listStride;
        int nListGroups = 1;
        if (OVERLAP) {
            nListGroups = 0;
            for (int base = 0; base < nList; base += CTA_THREADS) {
                const int i = base+int(threadIdx.x);
                const bool closes = i < nList && (i == nList-1 || cull[list[i+1]].contour != cull[list[i]].contour);
                const unsigned ballot = __ballot_sync(0xffffffffu, closes);
                if (lane == 0)
                    s_warpKeep[warp] = __popc(ballot);
                __syncthreads();
                int before = 0, total = 0;
                #pragma unroll
                for (int k = 0; k < WARPS; ++k) {
                    const int n = s_warpKeep[k];
                    if (k < int(warp)) before += n;
                    total += n;
                }
                if (closes)
                    groupEnd[nListGroups+before+__popc(ballot&((1u<<lane)-1u))] = i+1;
                nListGroups += total;
                __syncthreads();
            }
        }
        // parked per contour and texel: the resolved distance(s) and the winding (shape_distance_fold)
        FoldStore store;
        store.perContour = NOUT;
        if (OVERLAP && s_g.nonEmptyContours > B200_LOCAL_CONTOURS) {
            store.base = P.scratch+(size_t(blockIdx.x)*CTA_THREADS+threadIdx.x);
            store.wbase = P.scratchW+(size_t(blockIdx.x)*CTA_THREADS+threadIdx.x);
            store.stride = P.scratchStride;
        } else {
            store.base = localCd;
            store.wbase = localW;
            store.stride = 1;
        }
        // ---- the texels; a warp takes an 8x4 sub-block per pass, like k1_distance_field
        for (int pass = 0; pass < BLOCK_H/ROWS; ++pass) {
            const int sx0 = int(warp%(BLOCK_W/8))*8, sy0 = pass*ROWS+int(warp/(BLOCK_W/8))*4;      // sub-block origin inside the block
            // ---- level 2: the warp's sub-list for its sub-block (warp-synchronous; no CTA barrier inside)
            int nSub = 0, nSubGroups = 0;
            bool overflow = false;
            {
                const double scx = s_g.x0+bx0+sx0+4.0;
                const double scy = s_g.y0+(s_g.flip ? h-(by0+sy0)-2.0 : by0+sy0+2.0);
                const d2 sc = mk(scx/s_g.sx-s_g.tx, scy/s_g.sy-s_g.ty);
                const double shx = 3.5/fabs(s_g.sx), shy = 1.5/fabs(s_g.sy);
                const double shd = sqrt(shx*shx+shy*shy)*(1+1e-9)+1e-12*(fabs(sc.x)+fabs(sc.y))+slack;
                for (int g = 0; g < nListGroups; ++g) {
                    const int i0 = OVERLAP ? (g ? groupEnd[g-1] : 0) : 0;
                    const int i1 = OVERLAP ? groupEnd[g] : nList;
                    double capsq[NCH];
                    #pragma unroll
                    for (int k = 0; k < NCH; ++k)
                        capsq[k] = DBL_MAX;
                    for (int i = i0+int(lane); i < i1; i += 32) {
                        const CullRec cr = cull[list[i]];
                        const int mask = (MODE == 1 || MODE == 2) ? 1 : cr.color;
                        d2 v = sc-cr.p0;
                        const double q = dot(v, v);
                        #pragma unroll
                        for (int k = 0; k < NCH; ++k)
                            if (mask&(1<<k))
                                capsq[k] = dmin(capsq[k], q);
                    }
                    #pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        #pragma unroll
                        for (int d = 16; d > 0; d >>= 1)
                            capsq[k] = dmin(capsq[k], __shfl_xor_sync(0xffffffffu, capsq[k], d));
                        capsq[k] = sqrt(capsq[k]);      // (DBL_MAX for a channel without records: never consulted)
                    }
                    const int before = nSub;
                    for (int ib = i0; ib < i1; ib += 32) {
                        const int i = ib+int(lane);
                        bool keep = false;
                        int j = 0;
                        if (i < i1) {
                            j = list[i];
                            const CullRec cr = cull[j];
                            const int mask = (MODE == 1 || MODE == 2) ? 1 : cr.color;
                            double needT = 0;
                            #pragma unroll
                            for (int k = 0; k < NCH; ++k)
                                if (mask&(1<<k))
                                    needT = dmax(needT, capsq[k]);
                            needT = (needT+shd)*(1+1e-8);
                            keep = region_keeps<MODE>(cr, perp, j, sc, shd, needT);
                        }
                        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                        const int pos = nSub+__popc(ballot&((1u<<lane)-1u));
                        if (keep && pos < SUB_CAP)
                            s_sub[warp*SUB_CAP+pos] = j;
                        nSub += __popc(ballot);
                    }
                    if (OVERLAP && nSub > before) {
                        if (nSubGroups < SUB_GROUPS) {
                            if (lane == 0)
                                s_subEnd[warp][nSubGroups] = nSub;
                        } else
                            overflow = true;
                        ++nSubGroups;
                    }
                }
                overflow = overflow || nSub > SUB_CAP;
                __syncwarp();
            }
            ListIter it;
            it.list = overflow ? list : s_sub+warp*SUB_CAP;
            it.groupEnd = overflow ? groupEnd : s_subEnd[OVERLAP ? warp : 0];
            it.group = 0;
            const int nWalk = overflow ? nList : nSub;
            const int xl = sx0+int(lane&7);
            const int yl = sy0-pass*ROWS+int(lane>>3);         // row inside the pass: 0 .. ROWS-1
            const int x = bx0+xl, yOut = by0+pass*ROWS+yl;
            if (x < w && yOut < h) {
                const int y = s_g.flip ? h-1-yOut : yOut;
                d2 p = mk((s_g.x0+x+.5)/s_g.sx-s_g.tx, (s_g.y0+y+.5)/s_g.sy-s_g.ty);
                double dist[NOUT];
                shape_distance_list_filtered<MODE, OVERLAP>(recs, frecs, nWalk, it, O, fslack, p, store, dist);
                #pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    s_out[(yl*BLOCK_W+xl)*NOUT+k] = float(s_g.dmScale*(dist[k]+s_g.dmTranslate));
            }
            __syncthreads();
            for (int i = threadIdx.x; i < ROWS*BLOCK_W*NOUT; i += CTA_THREADS) {
                const int row = i/(BLOCK_W*NOUT), col = i%(BLOCK_W*NOUT);
                const int yo = by0+pass*ROWS+row;
                if (yo < h && bx0+col/NOUT < w)
                    P.field[s_g.outOffset+(size_t(yo)*w+bx0)*NOUT+col] = s_out[i];
            }
            __syncthreads();
        }
    }
}

}  // namespace b200
