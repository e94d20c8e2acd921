// kernels.cuh — the sm_100a kernels of the hot path and their launch parameters.
//
//   k0_windings          one thread per contour: Contour::winding()  (/root/reference/core/Contour.cpp:57-81)
//   k0_prepare_big       one thread per segment of shapes too large for shared memory: SegRec -> global buffer
//   k1_distance_field    generateDistanceField<Combiner<Selector>>    (/root/reference/core/msdfgen.cpp:52-76)
//   k2_classify / k3_exact_distance   msdfErrorCorrectionInner<N> + apply  (error_correction.cuh)
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
    int nTiles;                   // tiles [0, nSmallTiles) belong to glyphs staged in shared memory, the rest to big shapes
    int nSmallTiles;
    int tileBegin, tileEnd;       // tile range of this launch
    int counterIndex;             // which counter this launch pulls tiles from
    int nContours;
    const SegRec *bigRecs;        // prepared records of big shapes
    int *counters;                // tile counters: [0],[3] k1 small/big, [1] k2, [2],[4] k3 small/big
    struct Candidate *candidates; // k2 -> k3: per-tile slots of exact-distance checks, candCap per tile
    int *candCount;               // per-tile number of pushes (may exceed candCap: overflow texels carry EC_FLAG_RECHECK)
    int candCap;
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
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            counters[i] = 0;
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

}  // namespace b200
