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
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            counters[i] = 0;
    }
    if (c < nContours)
        windings[c] = (int8_t) contour_winding(raw, (unsigned) c);
}

// items: (glyph index, evaluation position) pairs are implied by a prefix table: bigGlyphs[i] = glyph index, bigBase[i] = first record
__global__ void k0_prepare_big(RawShapes raw, const GlyphDesc *glyphs, const int *bigGlyphs, int nBig, SegRec *recs, CullRec *cull, PerpRec *perp) {
    for (int i = blockIdx.y; i < nBig; i += gridDim.y) {
        GlyphDesc g = glyphs[bigGlyphs[i]];
        unsigned n = g.s1-g.s0;
        for (unsigned j = blockIdx.x*blockDim.x+threadIdx.x; j < n; j += gridDim.x*blockDim.x) {
            SegRec &r = recs[size_t(g.recBase)+j];
            prepare_record(raw, g, j, r);
            CullRec c;
            c.bc = r.bc; c.br = r.br; c.color = r.color; c.contour = r.contour; c.p0 = r.p0;
            cull[size_t(g.recBase)+j] = c;
            PerpRec q;
            q.pEnd = r.pEnd; q.aDir = r.aDir; q.bDir = r.bDir; q.aBis = r.aBis; q.bBis = r.bBis;
            perp[size_t(g.recBase)+j] = q;
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

// BIG = false: tiles of glyphs whose records are prepared into shared memory by the CTA (the compiler sees a __shared__
// array: LDS).  BIG = true: tiles of shapes whose records were prepared into the global buffer by k0_prepare_big.
template <int MODE, bool OVERLAP, bool BIG>
__global__ void __launch_bounds__(CTA_THREADS, 2) k1_distance_field(KernelParams P) {
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
            if (i < tileEnd) {
                int x = int(i%unsigned(w));
                int yOut = int(i/unsigned(w));
                int y = s_g.flip ? s_g.h-1-yOut : yOut;
                // Projection::unproject(Point2(x+.5, y+.5))
                d2 p = mk((s_g.x0+x+.5)/s_g.sx-s_g.tx, (s_g.y0+y+.5)/s_g.sy-s_g.ty);
                double dist[NOUT];
                shape_distance<MODE, OVERLAP>(recs, nRec, p, store, dist);
                #pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    s_out[slot*NOUT+k] = float(s_g.dmScale*(dist[k]+s_g.dmTranslate));    // DistanceMapping::operator()
            }
            __syncthreads();
            unsigned nPix = tileEnd-passBase < CTA_THREADS ? tileEnd-passBase : CTA_THREADS;
            store_tile(P.field+s_g.outOffset+size_t(passBase)*NOUT, s_out, int(nPix)*NOUT);
            __syncthreads();
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
    double localStore[OVERLAP ? B200_LOCAL_CONTOURS*StoreLayout<MODE>::PER_CONTOUR : 1];
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
        ContourStore store;
        store.perContour = StoreLayout<MODE>::PER_CONTOUR;
        if (OVERLAP && s_g.nonEmptyContours > B200_LOCAL_CONTOURS) {
            store.base = P.scratch+(size_t(blockIdx.x)*CTA_THREADS+threadIdx.x);
            store.stride = P.scratchStride;
        } else {
            store.base = localStore;
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
                shape_distance_over<MODE, OVERLAP>(recs, nWalk, it, p, store, dist);
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
