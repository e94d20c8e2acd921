// c_abi.cu — implementation of include/msdfgen_b200.h: context, HBM-resident batches, launches, host<->device transfers.
// No torch, no oracle, no CPU compute path: every entry point that produces texels does so with the sm_100a kernels in
// kernels.cuh or fails.
#include "../../include/msdfgen_b200.h"
#include "error_correction.cuh"
#include "scanline.cuh"
#include "render.cuh"
#include "sdf_error.cuh"

#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <sched.h>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <unordered_map>
#include <new>
#include <string>
#include <algorithm>
#include <chrono>
#include <functional>
#include <vector>

using namespace b200;

namespace {

thread_local std::string g_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(B200MSDF_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap)
            return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes+bytes/4+256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200MSDF_ERR_ALLOC, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap)
            return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes+bytes/4+256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200MSDF_ERR_ALLOC, "cudaHostAlloc(%zu): %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// B200MSDF_TRACE=1: host-side phase timing of the host-buffer entry points on stderr (microseconds since the call started)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(getenv("B200MSDF_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) { }
    void mark(const char *what) const {
        if (on)
            fprintf(stderr, "[b200msdf] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now()-t0).count());
    }
};

// exact-distance candidate slots per 1024-texel tile (the default configuration produces about 7; ALWAYS_CHECK_DISTANCE about 230)
const int EC_CANDIDATES_PER_TILE = 384;

inline int channels_of(int mode) { return mode == B200MSDF_MSDF ? 3 : mode == B200MSDF_MTSDF ? 4 : 1; }

}  // namespace

struct b200msdf_batch {
    int mode = 0, N = 0;
    int nGlyphs = 0, nContours = 0, nSegments = 0, nTiles = 0, nSmallTiles = 0;
    int nBig = 0;
    size_t nBigRecs = 0;
    size_t outFloats = 0, texels = 0;
    int maxContours = 0;
    bool hasColors = false;
    bool launched = false;
    bool regions = false;
    float *outOverride = nullptr;   // final texels go here instead of `out` (device-accessible pinned host memory of the caller)
    bool ecRan = false;
    float *current = nullptr;       // device pointer to the most recent result (raw field after k1, final output after a full launch)
    DevBuf bigFrecs, bigExtent;
    DevBuf points, orders, colors, cbegin, windings, glyphs, tiles, bigGlyphs, bigRecs, bigCull, bigPerp, field, out, stencil, candidates, candCount;
    DevBuf signMatch;             // scanline pass of multi-channel fields: per-texel match flags (never the EC stencil)
    DevBuf bytes, byteTargets;    // 8-bit output path
    DevBuf blocks, blockWork;     // big shapes: 2-D block tiles of k1_big_blocks and its per-CTA list / cap workspace
    int nBlocks = 0, maxBigSegs = 0, maxBigContours = 0;
    std::vector<TileDesc> hBlocks;
    DevBuf k1Counters;            // one tile counter per k1a / k1b launch of a round (+ the recheck launch)
    DevBuf recheck;               // guard-band texel mask + tile flags
    DevBuf filt, tileFilt;        // k1a -> k1b filter words (sized at launch) and the small tiles' word offsets: [overlap][tile], [simple][tile]
    size_t filtWords[2] = { 0, 0 };
    std::vector<unsigned long long> hTileFilt;
    DevBuf scratchW;
    DevBuf counters, scratch;     // per batch, so that batches launched on different streams never share scheduling state
    std::vector<GlyphDesc> hGlyphs;
    std::vector<TileDesc> hTiles;
    std::vector<int> hBig;
    void release() {
        points.release(); orders.release(); colors.release(); cbegin.release(); windings.release(); glyphs.release(); tiles.release();
        bigGlyphs.release(); bigRecs.release(); bigCull.release(); bigPerp.release(); bigFrecs.release(); bigExtent.release(); field.release(); out.release(); stencil.release(); candidates.release(); candCount.release(); counters.release(); scratch.release(); scratchW.release(); bytes.release(); byteTargets.release(); signMatch.release();
        blocks.release(); blockWork.release(); filt.release(); tileFilt.release(); k1Counters.release(); recheck.release();
    }
};

struct b200msdf_ctx {
    int device = 0;
    int numSMs = 0;
    cudaStream_t own = nullptr, stream = nullptr;
    cudaStream_t aux[2] = { nullptr, nullptr };     // double-buffered chunks of large host batches (copy/compute overlap)
    PinBuf pinOut, pinIn, pinDesc;  // pinDesc: staging of per-batch descriptors (glyphs, tiles) so that their H2D copies are true async DMAs
    cudaEvent_t descDone = nullptr; // the last copy out of pinDesc
    uint64_t launches = 0;
    unsigned k1Parts = 3;           // per-kernel timing (b200msdf_batch_launch_stages): bit 0 k1a_filter, bit 1 k1b_exact
    // resident CTAs per SM of every kernel launched through this context, and the dynamic-shared-memory opt-in that goes with
    // it: both are per DEVICE, and a context is bound to one device
    std::unordered_map<const void *, int> occupancy;
    b200msdf_batch pool;        // reused by the host-buffer entry points
    b200msdf_batch chunk[2];    // reused by the chunked host path
    b200msdf_batch *cur = nullptr;
};

namespace {

int bind(b200msdf_ctx *ctx) {
    if (!ctx)
        return fail(B200MSDF_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return 0;
}

int check_shapes(const b200msdf_shapes *s) {
    if (!s)
        return fail(B200MSDF_ERR_ARG, "null shapes");
    if (s->n_glyphs < 0 || s->n_contours < 0 || s->n_segments < 0)
        return fail(B200MSDF_ERR_ARG, "negative counts");
    if (s->n_glyphs > 0 && !s->glyph_contour_begin)
        return fail(B200MSDF_ERR_ARG, "glyph_contour_begin is null");
    if ((s->n_contours > 0 || s->n_glyphs > 0) && !s->contour_segment_begin)
        return fail(B200MSDF_ERR_ARG, "contour_segment_begin is null");
    if (s->n_segments > 0 && (!s->points || !s->orders))
        return fail(B200MSDF_ERR_ARG, "points/orders are null");
    for (int g = 0; g < s->n_glyphs; ++g)
        if (s->glyph_contour_begin[g] > s->glyph_contour_begin[g+1])
            return fail(B200MSDF_ERR_ARG, "glyph_contour_begin not monotonic at %d", g);
    if (s->n_glyphs > 0 && s->glyph_contour_begin[s->n_glyphs] > (uint32_t) s->n_contours)
        return fail(B200MSDF_ERR_ARG, "glyph_contour_begin exceeds n_contours");
    for (int c = 0; c < s->n_contours; ++c)
        if (s->contour_segment_begin[c] > s->contour_segment_begin[c+1])
            return fail(B200MSDF_ERR_ARG, "contour_segment_begin not monotonic at %d", c);
    if (s->contour_segment_begin && s->contour_segment_begin[s->n_contours] > (uint32_t) s->n_segments)
        return fail(B200MSDF_ERR_ARG, "contour_segment_begin exceeds n_segments");
    // EdgeSegment::type(): 1, 2 or 3 — anything else would index past the control points on the device
    unsigned bad = 0;
    for (int i = 0; i < s->n_segments; ++i)
        bad |= (unsigned) (s->orders[i]-1u) > 2u;
    if (bad)
        return fail(B200MSDF_ERR_ARG, "orders must be 1 (linear), 2 (quadratic) or 3 (cubic)");
    return 0;
}

// Builds per-glyph descriptors + tiles and uploads everything. flips[g] = output orientation != shape orientation.
int upload_batch(b200msdf_ctx *ctx, b200msdf_batch *b, int mode, const b200msdf_shapes *s, const b200msdf_transform *xf, int xfPerGlyph,
                 const int32_t *dims, int dimsPerGlyph, const uint8_t *outYDown, const int32_t *origins = nullptr, int originsPerGlyph = 0) {
    if (mode < 1 || mode > 4)
        return fail(B200MSDF_ERR_ARG, "mode %d out of range", mode);
    int rc = check_shapes(s);
    if (rc) return rc;
    if (!xf || !dims)
        return fail(B200MSDF_ERR_ARG, "null transformation/dims");
    // the raw shape arrays go first: with pinned caller memory their DMA runs while the host builds the descriptors below
    cudaStream_t st = ctx->stream;
    const size_t nSeg = (size_t) s->n_segments, nCon = (size_t) s->n_contours;
    if ((rc = b->points.reserve(nSeg*64+64))) return rc;
    if ((rc = b->orders.reserve(nSeg+16))) return rc;
    if ((rc = b->colors.reserve(nSeg+16))) return rc;
    if ((rc = b->cbegin.reserve((nCon+1)*4))) return rc;
    if (nSeg) {
        CUDA_TRY(cudaMemcpyAsync(b->points.p, s->points, nSeg*64, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->orders.p, s->orders, nSeg, cudaMemcpyHostToDevice, st));
        if (s->colors)
            CUDA_TRY(cudaMemcpyAsync(b->colors.p, s->colors, nSeg, cudaMemcpyHostToDevice, st));
    }
    if (s->contour_segment_begin)
        CUDA_TRY(cudaMemcpyAsync(b->cbegin.p, s->contour_segment_begin, (nCon+1)*4, cudaMemcpyHostToDevice, st));
    else
        CUDA_TRY(cudaMemsetAsync(b->cbegin.p, 0, 4, st));
    b->mode = mode;
    b->N = channels_of(mode);
    b->nGlyphs = s->n_glyphs;
    b->nContours = s->n_contours;
    b->nSegments = s->n_segments;
    b->hasColors = s->colors != nullptr;
    b->launched = false;
    b->regions = origins != nullptr;
    b->hGlyphs.resize((size_t) s->n_glyphs);
    b->hTiles.clear();
    b->hBig.clear();
    b->hBlocks.clear();
    b->maxBigSegs = b->maxBigContours = 0;
    size_t outFloats = 0, texels = 0, bigRecs = 0;
    int maxContours = 0;
    std::vector<TileDesc> bigTiles;
    for (int g = 0; g < s->n_glyphs; ++g) {
        GlyphDesc &d = b->hGlyphs[(size_t) g];
        const b200msdf_transform &t = xf[xfPerGlyph ? g : 0];
        d.sx = t.scale_x; d.sy = t.scale_y; d.tx = t.translate_x; d.ty = t.translate_y;
        d.dmScale = t.dm_scale; d.dmTranslate = t.dm_translate;
        d.w = dims[dimsPerGlyph ? 2*g : 0];
        d.h = dims[dimsPerGlyph ? 2*g+1 : 1];
        if (d.w < 0 || d.h < 0 || (long long) d.w*d.h > (long long) UINT_MAX-TILE_PIXELS)
            return fail(B200MSDF_ERR_ARG, "bad dimensions for glyph %d", g);
        d.c0 = s->glyph_contour_begin[g];
        d.c1 = s->glyph_contour_begin[g+1];
        d.s0 = s->contour_segment_begin[d.c0];
        d.s1 = s->contour_segment_begin[d.c1];
        int nonEmpty = 0;
        for (unsigned c = d.c0; c < d.c1; ++c)
            nonEmpty += s->contour_segment_begin[c+1] > s->contour_segment_begin[c];
        d.nonEmptyContours = nonEmpty;
        if (nonEmpty > maxContours) maxContours = nonEmpty;
        d.shapeYDown = s->y_downward ? (s->y_downward[g] != 0) : 0;
        int oyd = outYDown ? (outYDown[g] != 0) : 0;
        d.flip = oyd != d.shapeYDown;
        d.x0 = origins ? origins[originsPerGlyph ? 2*g : 0] : 0;
        d.y0 = origins ? origins[originsPerGlyph ? 2*g+1 : 1] : 0;
        if (origins)
            d.flip = 0;     // regions are generated in the shape's orientation
        d.outOffset = outFloats;
        d.texelOffset = texels;
        size_t px = (size_t) d.w*(size_t) d.h;
        outFloats += (px*(size_t) b->N+3)&~size_t(3);
        texels += px;
        unsigned nSeg = d.s1-d.s0;
        if (nSeg > SMEM_RECORDS) {
            d.recBase = (int) bigRecs;
            d.bigIndex = (int) b->hBig.size();
            bigRecs += nSeg;
            b->hBig.push_back(g);
            if ((int) nSeg > b->maxBigSegs) b->maxBigSegs = (int) nSeg;
            if ((int) (d.c1-d.c0) > b->maxBigContours) b->maxBigContours = (int) (d.c1-d.c0);
            // the distance kernel walks big shapes in BLOCK_W x BLOCK_H blocks (tile-level culling)
            if (px) {
                unsigned nb = (unsigned) ((d.w+BLOCK_W-1)/BLOCK_W)*(unsigned) ((d.h+BLOCK_H-1)/BLOCK_H);
                for (unsigned k = 0; k < nb; ++k) {
                    TileDesc td;
                    td.glyph = (unsigned) g;
                    td.pix0 = k;
                    b->hBlocks.push_back(td);
                }
            }
        } else {
            d.recBase = -1;
            d.bigIndex = -1;
        }
        for (size_t p0 = 0; p0 < px; p0 += TILE_PIXELS) {
            TileDesc td;
            td.glyph = (unsigned) g;
            td.pix0 = (unsigned) p0;
            (d.recBase < 0 ? b->hTiles : bigTiles).push_back(td);
        }
    }
    // Longest-processing-time-first: the persistent CTAs pull tiles from an atomic counter (greedy list scheduling), so handing
    // out the expensive tiles first leaves only cheap ones for the tail.  Cost model: segments x texels of the tile, quantised
    // to 256 classes; a stable counting sort (O(tiles), ~10 us for 4096 tiles — this runs inside every host-buffer call) keeps
    // equal-cost tiles, e.g. the tiles of one glyph, in their original order.
    if (b->hTiles.size() > 1) {
        const size_t n = b->hTiles.size();
        const int CLASSES = 256;
        std::vector<uint8_t> cls(n);
        size_t count[CLASSES+1] = { 0 };
        for (size_t t = 0; t < n; ++t) {
            const GlyphDesc &d = b->hGlyphs[b->hTiles[t].glyph];
            size_t px = (size_t) d.w*(size_t) d.h-b->hTiles[t].pix0;
            size_t cost = (size_t) (d.s1-d.s0)*(px < TILE_PIXELS ? px : (size_t) TILE_PIXELS);      // <= SMEM_RECORDS*TILE_PIXELS
            size_t c = cost*(CLASSES-1)/((size_t) SMEM_RECORDS*TILE_PIXELS);
            cls[t] = (uint8_t) (CLASSES-1-(c < (size_t) CLASSES ? c : (size_t) CLASSES-1));         // class 0 = most expensive
            ++count[cls[t]+1];
        }
        for (int c = 0; c < CLASSES; ++c)
            count[c+1] += count[c];
        std::vector<TileDesc> sorted(n);
        for (size_t t = 0; t < n; ++t)
            sorted[count[cls[t]]++] = b->hTiles[t];
        b->hTiles.swap(sorted);
    }
    b->nSmallTiles = (int) b->hTiles.size();
    // filter words of the small tiles (k1a_filter -> k1b_exact), for the overlapping and for the simple combiner: per unit of
    // 32 texels 32 x words(glyph) — the unit count mirrors UnitGrid in kernels.cuh
    {
        const size_t nT = b->hTiles.size();
        b->hTileFilt.resize(2*nT);
        std::vector<unsigned> wordsOv((size_t) s->n_glyphs), wordsSimple((size_t) s->n_glyphs);
        for (int g = 0; g < s->n_glyphs; ++g) {
            const GlyphDesc &d = b->hGlyphs[(size_t) g];
            unsigned ov = 0;
            for (unsigned c = d.c0; c < d.c1; ++c) {
                unsigned n = s->contour_segment_begin[c+1]-s->contour_segment_begin[c];
                if (n)
                    ov += mode == 1 ? filter_words_group<1>(n) : mode == 2 ? filter_words_group<2>(n) : filter_words_group<3>(n);
            }
            unsigned n = d.s1-d.s0;
            wordsOv[(size_t) g] = ov;
            wordsSimple[(size_t) g] = !n ? 0u : mode == 1 ? filter_words_group<1>(n) : mode == 2 ? filter_words_group<2>(n) : filter_words_group<3>(n);
        }
        size_t offOv = 0, offSimple = 0;
        for (size_t t = 0; t < nT; ++t) {
            const TileDesc &td = b->hTiles[t];
            const GlyphDesc &d = b->hGlyphs[td.glyph];
            const size_t total = (size_t) d.w*(size_t) d.h;
            const size_t pixEnd = td.pix0+TILE_PIXELS < total ? td.pix0+TILE_PIXELS : total;
            const bool blocked = (d.w&7) == 0 && d.w <= TILE_PIXELS && (TILE_PIXELS%d.w) == 0;
            size_t units;
            if (blocked) {
                size_t nRows = (pixEnd-td.pix0)/(size_t) d.w;
                units = (size_t) (d.w>>3)*((nRows+3)>>2);
            } else
                units = (pixEnd-td.pix0+31)>>5;
            b->hTileFilt[t] = offOv;
            b->hTileFilt[nT+t] = offSimple;
            offOv += units*32*wordsOv[td.glyph];
            offSimple += units*32*wordsSimple[td.glyph];
        }
        b->filtWords[0] = offOv;
        b->filtWords[1] = offSimple;
    }
    b->hTiles.insert(b->hTiles.end(), bigTiles.begin(), bigTiles.end());
    b->outFloats = outFloats;
    b->texels = texels;
    b->maxContours = maxContours;
    b->nTiles = (int) b->hTiles.size();
    b->nBig = (int) b->hBig.size();
    b->nBigRecs = bigRecs;
    b->nBlocks = (int) b->hBlocks.size();

    if ((rc = b->windings.reserve(nCon+16))) return rc;
    if ((rc = b->glyphs.reserve(sizeof(GlyphDesc)*(b->hGlyphs.size()+1)))) return rc;
    if ((rc = b->tiles.reserve(sizeof(TileDesc)*(b->hTiles.size()+1)))) return rc;
    if ((rc = b->tileFilt.reserve(sizeof(unsigned long long)*(b->hTileFilt.size()+1)))) return rc;
    if ((rc = b->out.reserve(outFloats*4+64))) return rc;
    if ((rc = b->counters.reserve(64))) return rc;
    if (b->N >= 3) {
        if ((rc = b->field.reserve(outFloats*4+64))) return rc;
        if ((rc = b->stencil.reserve(texels+64))) return rc;
        if ((rc = b->candidates.reserve(sizeof(Candidate)*(size_t) EC_CANDIDATES_PER_TILE*(b->hTiles.size()+1)))) return rc;
        if ((rc = b->candCount.reserve(sizeof(int)*(b->hTiles.size()+1)))) return rc;
    }
    if (b->nBig) {
        if ((rc = b->bigGlyphs.reserve(sizeof(int)*b->hBig.size()))) return rc;
        if ((rc = b->bigRecs.reserve(sizeof(SegRec)*bigRecs))) return rc;
        if ((rc = b->bigCull.reserve(sizeof(CullRec)*bigRecs))) return rc;
        if ((rc = b->bigPerp.reserve(sizeof(PerpRec)*bigRecs))) return rc;
        if ((rc = b->bigFrecs.reserve(sizeof(FRec)*bigRecs))) return rc;
        if ((rc = b->bigExtent.reserve(sizeof(unsigned)*b->hBig.size()))) return rc;
        CUDA_TRY(cudaMemcpyAsync(b->bigGlyphs.p, b->hBig.data(), sizeof(int)*b->hBig.size(), cudaMemcpyHostToDevice, st));
        if ((rc = b->blocks.reserve(sizeof(TileDesc)*(b->hBlocks.size()+1)))) return rc;
        if (!b->hBlocks.empty())
            CUDA_TRY(cudaMemcpyAsync(b->blocks.p, b->hBlocks.data(), sizeof(TileDesc)*b->hBlocks.size(), cudaMemcpyHostToDevice, st));
    }
    // glyph and tile descriptors go through pinned staging (a cudaMemcpyAsync from a pageable std::vector is staged and
    // synchronous on the host)
    {
        const size_t gBytes = sizeof(GlyphDesc)*b->hGlyphs.size(), tBytes = sizeof(TileDesc)*b->hTiles.size();
        const size_t fBytes = sizeof(unsigned long long)*b->hTileFilt.size();
        if (gBytes+tBytes) {
            if (ctx->descDone)
                CUDA_TRY(cudaEventSynchronize(ctx->descDone));      // an earlier batch may still be reading the staging buffer
            if ((rc = ctx->pinDesc.reserve(gBytes+tBytes+fBytes+64))) return rc;
            char *stage = static_cast<char *>(ctx->pinDesc.p);
            if (gBytes) {
                memcpy(stage, b->hGlyphs.data(), gBytes);
                CUDA_TRY(cudaMemcpyAsync(b->glyphs.p, stage, gBytes, cudaMemcpyHostToDevice, st));
            }
            if (tBytes) {
                memcpy(stage+gBytes, b->hTiles.data(), tBytes);
                CUDA_TRY(cudaMemcpyAsync(b->tiles.p, stage+gBytes, tBytes, cudaMemcpyHostToDevice, st));
            }
            if (fBytes) {
                memcpy(stage+gBytes+tBytes, b->hTileFilt.data(), fBytes);
                CUDA_TRY(cudaMemcpyAsync(b->tileFilt.p, stage+gBytes+tBytes, fBytes, cudaMemcpyHostToDevice, st));
            }
            if (!ctx->descDone)
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->descDone, cudaEventDisableTiming));
            CUDA_TRY(cudaEventRecord(ctx->descDone, st));
        }
    }
    return 0;
}

template <typename K>
int grid_for(b200msdf_ctx *ctx, K kernel, int nTiles, size_t dynSmem, int threads, int *grid) {
    // resident CTAs per SM of this kernel on this context's device: queried once and cached (the query costs microseconds
    // per launch); cudaFuncSetAttribute is per device too, so the cache lives in the context
    std::unordered_map<const void *, int> &cache = ctx->occupancy;
    int perSM = 0;
    auto it = cache.find((const void *) kernel);
    if (it != cache.end())
        perSM = it->second;
    else {
        if (dynSmem)
            CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dynSmem));
        // One shared-memory carve-out for every kernel of the pipeline: consecutive kernels with different carve-outs make the SMs
        // drain and reconfigure their L1/shared split at every kernel boundary (B200MSDF_CARVEOUT=percent, -1 = driver's choice)
        {
            static const int carveout = getenv("B200MSDF_CARVEOUT") ? atoi(getenv("B200MSDF_CARVEOUT")) : -1;
            if (carveout >= 0)
                CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
        }
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, threads, dynSmem));
        cache[(const void *) kernel] = perSM;
    }
    if (perSM < 1) perSM = 1;
    long long g = (long long) perSM*ctx->numSMs;
    *grid = (int) (g < nTiles ? g : nTiles);
    if (*grid < 1) *grid = 1;
    return 0;
}

template <typename K>
int launch_tiles(b200msdf_ctx *ctx, K kernel, KernelParams &P, int perThreadDoubles, size_t dynSmem = 0, int threads = CTA_THREADS) {
    int grid = 1, rc;
    if (P.tileEnd <= P.tileBegin)
        return 0;
    if ((rc = grid_for(ctx, kernel, P.tileEnd-P.tileBegin, dynSmem, threads, &grid))) return rc;
    P.scratchStride = (size_t) grid*threads;
    if (perThreadDoubles > 0) {
        if ((rc = ctx->cur->scratch.reserve(P.scratchStride*(size_t) perThreadDoubles*sizeof(double)))) return rc;
        P.scratch = ctx->cur->scratch.as<double>();
    } else
        P.scratch = nullptr;
    kernel<<<grid, threads, dynSmem, ctx->stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    ++ctx->launches;
    return 0;
}

template <int MODE, bool OVERLAP>
int launch_k1_mode(b200msdf_ctx *ctx, KernelParams &P, int scratch) {
    int rc;
    P.tileBegin = 0; P.tileEnd = P.nSmallTiles; P.counterIndex = 0;
    if (getenv("B200MSDF_K1_LEGACY")) {
        if ((rc = launch_tiles(ctx, k1_distance_field<MODE, OVERLAP, false>, P, scratch))) return rc;
    } else if (P.nSmallTiles > 0) {
        // k1a_filter (fp32 bounds -> survivor masks) then k1b_exact (fp64 evaluation of the survivors + contour combiner).
        // B200MSDF_K1_ROUND=n runs them in rounds of n tiles that reuse one filter-word buffer (L2-resident words, less memory);
        // measured on B200: the tail of every persistent launch costs more than the L2 hits save (4096-glyph atlas: one round
        // 2.58 ms, 2048 tiles 2.75, 1024 tiles 2.99), so the default is ONE round up to 1 GiB of filter words
        b200msdf_batch *bb = ctx->cur;
        const unsigned long long *hOff = bb->hTileFilt.data()+(OVERLAP ? 0 : (size_t) P.nSmallTiles);
        const size_t totalWords = bb->filtWords[OVERLAP ? 0 : 1];
        P.tileFilt = bb->tileFilt.as<unsigned long long>()+(OVERLAP ? 0 : (size_t) P.nSmallTiles);
        int roundTiles = P.nSmallTiles;
        if (totalWords > (size_t) 256*1024*1024)
            roundTiles = (int) ((double) P.nSmallTiles*(256.0*1024*1024)/(double) totalWords)+1;
        if (const char *env = getenv("B200MSDF_K1_ROUND"))
            roundTiles = atoi(env) > 0 ? atoi(env) : P.nSmallTiles;
        const int rounds = (P.nSmallTiles+roundTiles-1)/roundTiles;
        // every round reuses the same buffer (its lines then never leave the L2): sized for the largest round
        size_t roundWords = 0;
        for (int r = 0; r < rounds; ++r) {
            const size_t t0 = (size_t) r*roundTiles, t1 = t0+roundTiles < (size_t) P.nSmallTiles ? t0+roundTiles : (size_t) P.nSmallTiles;
            const size_t w = (t1 < (size_t) P.nSmallTiles ? hOff[t1] : totalWords)-hOff[t0];
            if (w > roundWords) roundWords = w;
        }
        if ((rc = bb->filt.reserve(roundWords*sizeof(unsigned)+256))) return rc;
        if ((rc = bb->k1Counters.reserve(sizeof(int)*(2*(size_t) rounds+1)))) return rc;
        CUDA_TRY(cudaMemsetAsync(bb->k1Counters.p, 0, sizeof(int)*(2*(size_t) rounds+1), ctx->stream));
#if B200_GUARD_MODE == 2
        // guard band of the fast foot-point roots: k1b flags the texels (and tiles) whose outcome hinged on the last bits of a root
        const size_t maskBytes = ((bb->texels+31)/32)*4, tileBytes = ((size_t) P.nSmallTiles+3)&~size_t(3);
        if ((rc = bb->recheck.reserve(maskBytes+tileBytes+16))) return rc;
        CUDA_TRY(cudaMemsetAsync(bb->recheck.p, 0, maskBytes+tileBytes+4, ctx->stream));
        P.recheckMask = bb->recheck.as<unsigned>();
        P.recheckTiles = bb->recheck.as<unsigned char>()+maskBytes;
        P.recheckAny = reinterpret_cast<int *>(bb->recheck.as<unsigned char>()+maskBytes+tileBytes);
#endif
        int *const batchCounters = P.counters;
        P.counters = bb->k1Counters.as<int>();
        for (int r = 0; r < rounds && !rc; ++r) {
            P.tileBegin = r*roundTiles;
            P.tileEnd = P.tileBegin+roundTiles < P.nSmallTiles ? P.tileBegin+roundTiles : P.nSmallTiles;
            P.counterIndex = 2*r;
            P.filt = bb->filt.as<unsigned>()-hOff[P.tileBegin];       // tileFilt offsets are batch-wide; the round starts at 0
            if (ctx->k1Parts&1u)
                rc = launch_tiles(ctx, k1a_filter<MODE, OVERLAP>, P, 0, 0, K1_THREADS);
            P.counterIndex = 2*r+1;
            if (!rc && (ctx->k1Parts&2u)) rc = launch_tiles(ctx, k1b_exact<MODE, OVERLAP>, P, scratch, 0, K1_THREADS);
        }
#if B200_GUARD_MODE == 2
        // ... and the flagged texels are recomputed with the literal solver (a handful per atlas; tiles without any are skipped)
        P.tileBegin = 0; P.tileEnd = P.nSmallTiles; P.counterIndex = 2*rounds;
        if (!rc && (ctx->k1Parts&2u)) rc = launch_tiles(ctx, k1_distance_field<MODE, OVERLAP, true>, P, scratch);
#endif
        P.counters = batchCounters;
        P.tileBegin = 0; P.tileEnd = P.nSmallTiles; P.counterIndex = 0;
        if (rc) return rc;
    }
    // big shapes: 2-D blocks with per-block record lists; each resident CTA owns a slice of the list / cap workspace
    b200msdf_batch *b = ctx->cur;
    if (b->nBlocks > 0) {
        int grid = 1;
        const size_t subListBytes = (size_t) (CTA_THREADS/32)*SUB_CAP*sizeof(int);      // the warps' sub-lists (dynamic shared memory)
        if ((rc = grid_for(ctx, k1_big_blocks<MODE, OVERLAP>, b->nBlocks, subListBytes, CTA_THREADS, &grid))) return rc;
        if (OVERLAP && b->maxBigContours > B200_LOCAL_CONTOURS) {
            // the parked per-contour state below grows with the contour count of the largest shape, not with the work: keep it
            // under 8 GiB by launching fewer CTAs (blocks are pulled from a counter, so any grid size covers them)
            const size_t perCta = (size_t) CTA_THREADS*(size_t) b->maxBigContours*(ModeTraits<MODE>::NOUT*sizeof(double)+1);
            const size_t cap = size_t(8)<<30;
            if ((size_t) grid*perCta > cap)
                grid = cap/perCta > 1 ? (int) (cap/perCta) : 1;
        }
        P.blocks = b->blocks.as<TileDesc>();
        P.nBlocks = b->nBlocks;
        P.listStride = ((size_t) b->maxBigSegs+3)&~size_t(3);
        P.capStride = (size_t) (b->maxBigContours > 0 ? b->maxBigContours : 1)*ModeTraits<MODE>::NCH;
        size_t capBytes = (size_t) grid*P.capStride*sizeof(unsigned long long);
        // per CTA: listStride record indices followed by up to maxBigContours group ends
        P.listPerCta = P.listStride+(((size_t) b->maxBigContours+4)&~size_t(3));
        if ((rc = b->blockWork.reserve(capBytes+(size_t) grid*P.listPerCta*sizeof(int)))) return rc;
        P.blockCaps = b->blockWork.as<unsigned long long>();
        P.blockList = reinterpret_cast<int *>(static_cast<char *>(b->blockWork.p)+capBytes);
        // parked per contour and texel (shape_distance_fold): NOUT resolved distances + one byte of winding
        P.scratchStride = (size_t) grid*CTA_THREADS;
        if (OVERLAP && b->maxBigContours > B200_LOCAL_CONTOURS) {
            const size_t slots = P.scratchStride*(size_t) b->maxBigContours;
            if ((rc = b->scratch.reserve(slots*ModeTraits<MODE>::NOUT*sizeof(double)))) return rc;
            if ((rc = b->scratchW.reserve(slots))) return rc;
            P.scratch = b->scratch.as<double>();
            P.scratchW = b->scratchW.as<signed char>();
        } else {
            P.scratch = nullptr;
            P.scratchW = nullptr;
        }
        P.counterIndex = 3;
        k1_big_blocks<MODE, OVERLAP><<<grid, CTA_THREADS, subListBytes, ctx->stream>>>(P);
        CUDA_TRY(cudaGetLastError());
        ++ctx->launches;
    }
    P.tileBegin = 0; P.tileEnd = P.nTiles; P.counterIndex = 0;
    return 0;
}

int launch_k1(b200msdf_ctx *ctx, int mode, bool overlap, KernelParams &P, int maxContours) {
    const int per[5] = { 0, StoreLayout<1>::PER_CONTOUR, StoreLayout<2>::PER_CONTOUR, StoreLayout<3>::PER_CONTOUR, StoreLayout<4>::PER_CONTOUR };
    if (mode < 1 || mode > 4)
        return fail(B200MSDF_ERR_ARG, "bad mode");
    int scratch = (overlap && maxContours > B200_LOCAL_CONTOURS) ? maxContours*per[mode] : 0;
    switch (mode*2+(overlap ? 1 : 0)) {
        case 2: return launch_k1_mode<1, false>(ctx, P, scratch);
        case 3: return launch_k1_mode<1, true>(ctx, P, scratch);
        case 4: return launch_k1_mode<2, false>(ctx, P, scratch);
        case 5: return launch_k1_mode<2, true>(ctx, P, scratch);
        case 6: return launch_k1_mode<3, false>(ctx, P, scratch);
        case 7: return launch_k1_mode<3, true>(ctx, P, scratch);
        case 8: return launch_k1_mode<4, false>(ctx, P, scratch);
        case 9: return launch_k1_mode<4, true>(ctx, P, scratch);
    }
    return fail(B200MSDF_ERR_ARG, "bad mode");
}

// error correction = k2_classify (all texels) + k3_exact_distance (candidate verdicts; only when the configuration has a
// shape-distance pass)
int launch_k2(b200msdf_ctx *ctx, b200msdf_batch *b, int N, bool overlap, KernelParams &P, int maxContours) {
    int rc;
    P.candidates = b->candidates.as<Candidate>();
    P.candCount = b->candCount.as<int>();
    P.candCap = EC_CANDIDATES_PER_TILE;
    bool shapePass = !P.shapeless && (P.ecDcm == 1 || P.ecDcm == 2);
    if (shapePass)
        CUDA_TRY(cudaMemsetAsync(P.candCount, 0, sizeof(int)*(size_t) P.nTiles, ctx->stream));
    P.tileBegin = 0; P.tileEnd = P.nTiles; P.counterIndex = 1;
    rc = N == 3 ? launch_tiles(ctx, k2_classify<3>, P, 0) : launch_tiles(ctx, k2_classify<4>, P, 0);
    if (rc || !shapePass)
        return rc;
    int scratch = (overlap && maxContours > B200_LOCAL_CONTOURS) ? maxContours*StoreLayout<2>::PER_CONTOUR : 0;
    for (int big = 0; big < 2 && !rc; ++big) {
        P.tileBegin = big ? P.nSmallTiles : 0;
        P.tileEnd = big ? P.nTiles : P.nSmallTiles;
        P.counterIndex = big ? 4 : 2;
        if (N == 3) {
            if (overlap) rc = big ? launch_tiles(ctx, k3_exact_distance<3, true, true>, P, scratch, 0, K3_THREADS) : launch_tiles(ctx, k3_exact_distance<3, true, false>, P, scratch, 0, K3_THREADS);
            else rc = big ? launch_tiles(ctx, k3_exact_distance<3, false, true>, P, scratch, 0, K3_THREADS) : launch_tiles(ctx, k3_exact_distance<3, false, false>, P, scratch, 0, K3_THREADS);
        } else {
            if (overlap) rc = big ? launch_tiles(ctx, k3_exact_distance<4, true, true>, P, scratch, 0, K3_THREADS) : launch_tiles(ctx, k3_exact_distance<4, true, false>, P, scratch, 0, K3_THREADS);
            else rc = big ? launch_tiles(ctx, k3_exact_distance<4, false, true>, P, scratch, 0, K3_THREADS) : launch_tiles(ctx, k3_exact_distance<4, false, false>, P, scratch, 0, K3_THREADS);
        }
    }
    P.tileBegin = 0; P.tileEnd = P.nTiles; P.counterIndex = 0;
    return rc;
}

void fill_params(b200msdf_ctx *ctx, b200msdf_batch *b, const b200msdf_config *cfg, KernelParams &P) {
    memset(&P, 0, sizeof(P));
    P.raw.points = b->points.as<double>();
    P.raw.orders = b->orders.as<uint8_t>();
    P.raw.colors = b->hasColors ? b->colors.as<uint8_t>() : nullptr;
    P.raw.cbegin = b->cbegin.as<uint32_t>();
    P.raw.windings = b->windings.as<int8_t>();
    P.glyphs = b->glyphs.as<GlyphDesc>();
    P.tiles = b->tiles.as<TileDesc>();
    P.nTiles = b->nTiles;
    P.nSmallTiles = b->nSmallTiles;
    P.tileBegin = 0;
    P.tileEnd = b->nTiles;
    P.counterIndex = 0;
    P.nContours = b->nContours;
    P.bigRecs = b->bigRecs.as<SegRec>();
    P.bigCull = b->bigCull.as<CullRec>();
    P.bigPerp = b->bigPerp.as<PerpRec>();
    P.bigFrecs = b->bigFrecs.as<FRec>();
    P.bigExtent = b->bigExtent.as<unsigned>();
    P.counters = b->counters.as<int>();
    ctx->cur = b;
    P.ecMode = cfg->ec_mode;
    P.ecDcm = cfg->ec_distance_check_mode;
    P.minDeviationRatio = cfg->min_deviation_ratio;
    P.minImproveRatio = cfg->min_improve_ratio;
}

// windings (+ counter reset) and big-shape record preparation
int launch_prologue(b200msdf_ctx *ctx, b200msdf_batch *b, KernelParams &P) {
    int blocks = (b->nContours+127)/128;
    if (blocks < 1) blocks = 1;
    k0_windings<<<blocks, 128, 0, ctx->stream>>>(P.raw, b->nContours, b->windings.as<int8_t>(), P.counters);
    CUDA_TRY(cudaGetLastError());
    ++ctx->launches;
    if (b->nBig) {
        dim3 grid(64, (unsigned) (b->nBig < 1024 ? b->nBig : 1024));
        CUDA_TRY(cudaMemsetAsync(b->bigExtent.p, 0, sizeof(unsigned)*(size_t) b->nBig, ctx->stream));
        k0_prepare_big<<<grid, 128, 0, ctx->stream>>>(P.raw, P.glyphs, b->bigGlyphs.as<int>(), b->nBig, b->bigRecs.as<SegRec>(), b->bigCull.as<CullRec>(), b->bigPerp.as<PerpRec>(),
                                                      b->bigExtent.as<unsigned>());
        CUDA_TRY(cudaGetLastError());
        k0_prepare_big_filter<<<grid, 128, 0, ctx->stream>>>(P.glyphs, b->bigGlyphs.as<int>(), b->nBig, b->bigRecs.as<SegRec>(), b->bigFrecs.as<FRec>(),
                                                             b->bigExtent.as<unsigned>(), b->mode <= 2 ? 1 : 0);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 2;
    }
    return 0;
}

int check_config(const b200msdf_config *cfg) {
    if (!cfg)
        return fail(B200MSDF_ERR_ARG, "null config");
    if (cfg->ec_mode < 0 || cfg->ec_mode > 3 || cfg->ec_distance_check_mode < 0 || cfg->ec_distance_check_mode > 2)
        return fail(B200MSDF_ERR_ARG, "error-correction mode out of range");
    if (cfg->scanline_pass && (cfg->fill_rule < 0 || cfg->fill_rule > 3))
        return fail(B200MSDF_ERR_ARG, "fill rule out of range");
    return 0;
}

// distanceSignCorrection over every glyph of the batch: src -> dst (device pointers in the batch's output layout; may alias).
int launch_sign(b200msdf_ctx *ctx, b200msdf_batch *b, KernelParams &P, const float *src, float *dst, float zero, int fillRule) {
    if (fillRule < 0 || fillRule > 3)
        return fail(B200MSDF_ERR_ARG, "fill rule out of range");
    if (b->nTiles == 0)
        return 0;
    if (b->N >= 3 && b->regions)
        return fail(B200MSDF_ERR_ARG, "multi-channel sign correction is not available for region batches (needs neighbouring tiles)");
    int *ambiguous = b->counters.as<int>()+8;
    int8_t *match = nullptr;
    if (b->N >= 3) {
        int rc = b->signMatch.reserve(b->texels+64);
        if (rc) return rc;
        match = b->signMatch.as<int8_t>();
        CUDA_TRY(cudaMemsetAsync(ambiguous, 0, sizeof(int), ctx->stream));
    }
    switch (b->N) {
        case 1: k5_scanline_sign<1><<<b->nTiles, CTA_THREADS, 0, ctx->stream>>>(P, src, dst, match, ambiguous, zero, fillRule); break;
        case 3: k5_scanline_sign<3><<<b->nTiles, CTA_THREADS, 0, ctx->stream>>>(P, src, dst, match, ambiguous, zero, fillRule); break;
        default: k5_scanline_sign<4><<<b->nTiles, CTA_THREADS, 0, ctx->stream>>>(P, src, dst, match, ambiguous, zero, fillRule); break;
    }
    CUDA_TRY(cudaGetLastError());
    ++ctx->launches;
    if (b->N >= 3) {
        if (b->N == 3)
            k5_resolve_ambiguous<3><<<b->nTiles, CTA_THREADS, 0, ctx->stream>>>(P, dst, match, ambiguous, zero);
        else
            k5_resolve_ambiguous<4><<<b->nTiles, CTA_THREADS, 0, ctx->stream>>>(P, dst, match, ambiguous, zero);
        CUDA_TRY(cudaGetLastError());
        ++ctx->launches;
    }
    return 0;
}

int launch_batch(b200msdf_ctx *ctx, b200msdf_batch *b, const b200msdf_config *cfg, unsigned stages = 15u) {
    int rc = check_config(cfg);
    if (rc) return rc;
    if (b->nGlyphs == 0 || b->nTiles == 0) {
        b->launched = true;
        return 0;
    }
    KernelParams P;
    fill_params(ctx, b, cfg, P);
    bool ec = b->N >= 3 && cfg->ec_mode != B200MSDF_EC_DISABLED;
    if (ec && b->regions)
        return fail(B200MSDF_ERR_ARG, "error correction is not available for region batches (needs neighbouring tiles): disable it or generate whole images");
    float *finalOut = b->outOverride ? b->outOverride : b->out.as<float>();
    const bool scan = cfg->scanline_pass != 0;
    // k1's target: the field buffer when error correction follows; otherwise the final output — except that a scanline pass
    // in front of caller-owned (zero-copy) memory reads device memory and makes the final store itself
    P.field = ec ? b->field.as<float>() : scan ? b->out.as<float>() : finalOut;
    P.out = finalOut;
    P.stencil = ec ? b->stencil.as<uint8_t>() : nullptr;
    if (stages&1u)
        if ((rc = launch_prologue(ctx, b, P))) return rc;
    if (!(stages&1u)) {
        // the tile counters are normally reset by the prologue kernel
        CUDA_TRY(cudaMemsetAsync(P.counters, 0, 8*sizeof(int), ctx->stream));
    }
    if (stages&(2u|16u|32u)) {
        // stage bits 16 / 32: only k1a_filter / only k1b_exact (the latter reads the words a previous launch left) — kernel timing
        ctx->k1Parts = (stages&2u) ? 3u : ((stages&16u) ? 1u : 0u)|((stages&32u) ? 2u : 0u);
        rc = launch_k1(ctx, b->mode, cfg->overlap_support != 0, P, b->maxContours);
        ctx->k1Parts = 3u;
        if (rc) return rc;
    }
    b->current = P.field;
    if (scan && (stages&8u))
        if ((rc = launch_sign(ctx, b, P, P.field, ec ? P.field : finalOut, cfg->sdf_zero_value, cfg->fill_rule))) return rc;
    if (scan && !ec)
        b->current = finalOut;
    if (ec && (stages&4u)) {
        if ((rc = launch_k2(ctx, b, b->N, cfg->overlap_support != 0, P, b->maxContours))) return rc;
        b->current = finalOut;
    }
    b->launched = true;
    b->ecRan = ec;
    return 0;
}

// Copies device results (src: device pointer to the packed output) to caller bitmaps.
int download_to_bitmaps(b200msdf_ctx *ctx, b200msdf_batch *b, const float *dSrc, const b200msdf_bitmap *bm) {
    if (b->nGlyphs == 0 || b->outFloats == 0)
        return 0;
    int N = b->N;
    // Fast path: the caller's sections are one packed block in device layout -> a single DMA straight into it.
    bool packed = true;
    for (int g = 0; g < b->nGlyphs && packed; ++g) {
        const GlyphDesc &d = b->hGlyphs[(size_t) g];
        if ((size_t) d.w*d.h == 0)
            continue;
        // (device offsets are padded to 4 floats: a glyph whose float count is not a multiple of 4 would drag padding into
        // caller memory that belongs to no bitmap)
        if (bm[g].row_stride != d.w*N || bm[g].pixels != bm[0].pixels+(d.outOffset-b->hGlyphs[0].outOffset) || (((size_t) d.w*d.h*N)&3))
            packed = false;
    }
    if (packed && bm[0].pixels) {
        const GlyphDesc &last = b->hGlyphs[(size_t) b->nGlyphs-1];
        size_t floats = last.outOffset+(size_t) last.w*last.h*N;
        CUDA_TRY(cudaMemcpyAsync(bm[0].pixels, dSrc, floats*4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    int rc = ctx->pinOut.reserve(b->outFloats*4);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(ctx->pinOut.p, dSrc, b->outFloats*4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const float *h = static_cast<const float *>(ctx->pinOut.p);
    for (int g = 0; g < b->nGlyphs; ++g) {
        const GlyphDesc &d = b->hGlyphs[(size_t) g];
        if (!bm[g].pixels)
            continue;
        for (int r = 0; r < d.h; ++r)
            memcpy(bm[g].pixels+(ptrdiff_t) bm[g].row_stride*r, h+d.outOffset+(size_t) r*d.w*N, sizeof(float)*(size_t) d.w*N);
    }
    return 0;
}

int check_bitmaps(const b200msdf_bitmap *bm, int n) {
    if (n > 0 && !bm)
        return fail(B200MSDF_ERR_ARG, "null outputs");
    for (int g = 0; g < n; ++g) {
        if (bm[g].width < 0 || bm[g].height < 0)
            return fail(B200MSDF_ERR_ARG, "negative bitmap size at %d", g);
        if ((size_t) bm[g].width*bm[g].height > 0 && !bm[g].pixels)
            return fail(B200MSDF_ERR_ARG, "null pixels at %d", g);
    }
    return 0;
}

void dims_from_bitmaps(const b200msdf_bitmap *bm, int n, std::vector<int32_t> &dims, std::vector<uint8_t> &yd) {
    dims.resize((size_t) n*2);
    yd.resize((size_t) n);
    for (int g = 0; g < n; ++g) {
        dims[2*(size_t) g] = bm[g].width;
        dims[2*(size_t) g+1] = bm[g].height;
        yd[(size_t) g] = bm[g].y_downward != 0;
    }
}

}  // namespace

extern "C" {

int b200msdf_version(void) { return B200MSDF_VERSION; }

const char *b200msdf_last_error(void) { return g_error.c_str(); }

void b200msdf_default_config(b200msdf_config *cfg) {
    if (!cfg) return;
    cfg->overlap_support = 1;
    cfg->ec_mode = B200MSDF_EC_EDGE_PRIORITY;
    cfg->ec_distance_check_mode = B200MSDF_DCM_CHECK_AT_EDGE;
    cfg->min_deviation_ratio = 1.11111111111111111;
    cfg->min_improve_ratio = 1.11111111111111111;
    cfg->ec_buffer = nullptr;
    cfg->scanline_pass = 0;
    cfg->fill_rule = B200MSDF_FILL_NONZERO;
    cfg->sdf_zero_value = .5f;
}

int b200msdf_create(int device, b200msdf_ctx **out) {
    if (!out)
        return fail(B200MSDF_ERR_ARG, "null ctx pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(B200MSDF_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= count)
        return fail(B200MSDF_ERR_ARG, "device %d out of range (%d devices)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    b200msdf_ctx *ctx = new (std::nothrow) b200msdf_ctx;
    if (!ctx)
        return fail(B200MSDF_ERR_ALLOC, "out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete ctx; return fail(B200MSDF_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    ctx->numSMs = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return fail(B200MSDF_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    ctx->stream = ctx->own;
    for (int i = 0; i < 2; ++i) {
        e = cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(B200MSDF_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    }
    *out = ctx;
    return 0;
}

void b200msdf_destroy(b200msdf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->aux[i]) { cudaStreamSynchronize(ctx->aux[i]); cudaStreamDestroy(ctx->aux[i]); }
        ctx->chunk[i].release();
    }
    ctx->pool.release();
    ctx->pinOut.release();
    ctx->pinIn.release();
    ctx->pinDesc.release();
    if (ctx->descDone) cudaEventDestroy(ctx->descDone);
    if (ctx->own) cudaStreamDestroy(ctx->own);
    delete ctx;
}

int b200msdf_set_stream(b200msdf_ctx *ctx, void *cuda_stream, int use_own) {
    if (!ctx)
        return fail(B200MSDF_ERR_ARG, "null context");
    ctx->stream = use_own ? ctx->own : static_cast<cudaStream_t>(cuda_stream);
    return 0;
}

int b200msdf_synchronize(b200msdf_ctx *ctx) {
    int rc = bind(ctx);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

uint64_t b200msdf_launch_count(const b200msdf_ctx *ctx) { return ctx ? ctx->launches : 0; }

int b200msdf_debug_exact_root_calls(b200msdf_ctx *ctx, int reset, uint64_t *calls) {
    int rc = bind(ctx);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpyFromSymbol(&v, g_exactRootCalls, sizeof(v)));
    if (calls) *calls = v;
    if (reset) {
        v = 0;
        CUDA_TRY(cudaMemcpyToSymbol(g_exactRootCalls, &v, sizeof(v)));
    }
    return 0;
}

int b200msdf_bind_host_thread(b200msdf_ctx *ctx) {
    // Pins the calling thread to the CPUs next to the context's GPU (sysfs local_cpulist of its PCI function).  Pinned host
    // buffers the thread allocates afterwards land on that NUMA node, and the zero-copy stores / DMA of the host-buffer entry
    // points no longer cross the inter-socket link — with one process per GPU on a two-socket box this is what keeps the
    // end-to-end rate from dropping when all GPUs run at once.
    int rc = bind(ctx);
    if (rc) return rc;
    char busId[32] = { 0 };
    CUDA_TRY(cudaDeviceGetPCIBusId(busId, (int) sizeof(busId), ctx->device));
    for (char *c = busId; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = (char) (*c-'A'+'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", busId);
    FILE *f = fopen(path, "r");
    if (!f)
        return fail(B200MSDF_ERR_ARG, "cannot read %s", path);
    char list[1024] = { 0 };
    const bool got = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!got)
        return fail(B200MSDF_ERR_ARG, "%s is empty", path);
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (const char *c = list; *c && *c != '\n';) {
        char *end;
        long a = strtol(c, &end, 10), b2 = a;
        if (end == c) break;
        if (*end == '-') { c = end+1; b2 = strtol(c, &end, 10); }
        for (long k = a; k <= b2 && k < CPU_SETSIZE; ++k) { CPU_SET((int) k, &set); ++n; }
        c = *end == ',' ? end+1 : end;
    }
    if (n == 0)
        return fail(B200MSDF_ERR_ARG, "no CPUs listed in %s", path);
    // keep to what the process is allowed to use (cgroup / taskset)
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        cpu_set_t both;
        CPU_AND(&both, &set, &allowed);
        if (CPU_COUNT(&both) > 0)
            set = both;
    }
    if (sched_setaffinity(0, sizeof(set), &set) != 0)
        return fail(B200MSDF_ERR_ARG, "sched_setaffinity failed");
    return CPU_COUNT(&set);
}

int b200msdf_host_alloc(void **ptr, size_t bytes) {
    if (!ptr)
        return fail(B200MSDF_ERR_ARG, "null pointer");
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200MSDF_ERR_ALLOC, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return 0;
}

void b200msdf_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

namespace {

// A glyph range [g0, g1) of a batch as an independent batch (index arrays rebased on the host: O(contours)).
struct SubShapes {
    b200msdf_shapes view;
    std::vector<uint32_t> gcb, csb;
    SubShapes(const b200msdf_shapes *s, int g0, int g1) {
        uint32_t c0 = s->glyph_contour_begin[g0], c1 = s->glyph_contour_begin[g1];
        uint32_t s0 = s->contour_segment_begin[c0], s1 = s->contour_segment_begin[c1];
        gcb.resize((size_t) (g1-g0)+1);
        for (int g = g0; g <= g1; ++g)
            gcb[(size_t) (g-g0)] = s->glyph_contour_begin[g]-c0;
        csb.resize((size_t) (c1-c0)+1);
        for (uint32_t k = c0; k <= c1; ++k)
            csb[k-c0] = s->contour_segment_begin[k]-s0;
        view.n_glyphs = g1-g0;
        view.n_contours = (int32_t) (c1-c0);
        view.n_segments = (int32_t) (s1-s0);
        view.glyph_contour_begin = gcb.data();
        view.contour_segment_begin = csb.data();
        view.points = s->points ? s->points+8*(size_t) s0 : nullptr;
        view.orders = s->orders ? s->orders+s0 : nullptr;
        view.colors = s->colors ? s->colors+s0 : nullptr;
        view.y_downward = s->y_downward ? s->y_downward+g0 : nullptr;
    }
};

// Are the caller's sections one packed block in device layout (every glyph [h][w][N] contiguous, blocks back to back at
// 16-byte multiples)?  Then results can be DMA'd straight into them, chunk by chunk.
bool packed_outputs(const b200msdf_bitmap *bm, int n, int N) {
    size_t off = 0;
    for (int g = 0; g < n; ++g) {
        size_t fl = (size_t) bm[g].width*bm[g].height*N;
        if (fl == 0)
            continue;
        if (bm[g].row_stride != bm[g].width*N || bm[g].pixels != bm[0].pixels+off || (fl&3))
            return false;
        off += fl;
    }
    return n > 0 && bm[0].pixels != nullptr;
}

}  // namespace

int b200msdf_generate_batch(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                            const b200msdf_bitmap *outputs, const b200msdf_config *config) {
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_shapes(shapes))) return rc;
    if ((rc = check_config(config))) return rc;
    if ((rc = check_bitmaps(outputs, shapes->n_glyphs))) return rc;
    if (shapes->n_glyphs == 0)
        return 0;
    if (mode < 1 || mode > 4)
        return fail(B200MSDF_ERR_ARG, "mode %d out of range", mode);
    std::vector<int32_t> dims;
    std::vector<uint8_t> yd;
    dims_from_bitmaps(outputs, shapes->n_glyphs, dims, yd);
    const int n = shapes->n_glyphs;
    // Large atlases into one packed block: split into chunks on two streams so that the D2H of one chunk overlaps the
    // kernels of the next (the output is 5x the input, so the transfer is almost entirely D2H).
    // (measured on B200 / PCIe Gen5: the per-chunk launch and tail overhead cancels the overlap for the headline workload, so
    // chunking is opt-in; the zero-copy output path below is what removes the D2H from the critical path)
    int chunks = 1;
    if (const char *env = getenv("B200MSDF_CHUNKS"))
        chunks = atoi(env) > 0 ? atoi(env) : 1;
    if (chunks > 1 && ctx->stream == ctx->own && packed_outputs(outputs, n, channels_of(mode))) {
        cudaStream_t saved = ctx->stream;
        rc = 0;
        for (int c = 0; c < chunks && !rc; ++c) {
            int g0 = (int) ((long long) n*c/chunks), g1 = (int) ((long long) n*(c+1)/chunks);
            if (g1 <= g0)
                continue;
            b200msdf_batch *b = &ctx->chunk[c&1];
            ctx->stream = ctx->aux[c&1];
            SubShapes sub(shapes, g0, g1);
            rc = upload_batch(ctx, b, mode, &sub.view, xf+(xfPerGlyph ? g0 : 0), xfPerGlyph, dims.data()+2*(size_t) g0, 1, yd.data()+g0);
            if (!rc) rc = launch_batch(ctx, b, config);
            if (!rc && b->outFloats) {
                const GlyphDesc &last = b->hGlyphs.back();
                size_t floats = last.outOffset+(size_t) last.w*last.h*b->N;
                cudaError_t e = cudaMemcpyAsync(outputs[g0].pixels, b->out.p, floats*4, cudaMemcpyDeviceToHost, ctx->stream);
                if (e != cudaSuccess)
                    rc = fail(B200MSDF_ERR_CUDA, "cudaMemcpyAsync D2H: %s", cudaGetErrorString(e));
            }
        }
        ctx->stream = saved;
        cudaError_t e0 = cudaStreamSynchronize(ctx->aux[0]), e1 = cudaStreamSynchronize(ctx->aux[1]);
        if (!rc && (e0 != cudaSuccess || e1 != cudaSuccess))
            rc = fail(B200MSDF_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
        return rc;
    }
    b200msdf_batch *b = &ctx->pool;
    Trace trace;
    if ((rc = upload_batch(ctx, b, mode, shapes, xf, xfPerGlyph, dims.data(), 1, yd.data()))) return rc;
    trace.mark("descriptors + H2D enqueued");
    // Zero-copy output: when the caller's sections form one packed block of pinned, device-accessible host memory (e.g. from
    // b200msdf_host_alloc), the kernels store the final texels straight into it over PCIe (coalesced float4 rows, posted
    // writes) while they are still computing, instead of a separate D2H copy after the last kernel.
    float *direct = nullptr;
    if (packed_outputs(outputs, n, channels_of(mode)) && !getenv("B200MSDF_NO_ZEROCOPY")) {
        cudaPointerAttributes attr;
        // (the kernels write float4 rows: a packed block that starts at a 4-, 8- or 12-byte offset takes the staged path)
        if (cudaPointerGetAttributes(&attr, outputs[0].pixels) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer &&
            (reinterpret_cast<uintptr_t>(attr.devicePointer)&15) == 0)
            direct = static_cast<float *>(attr.devicePointer);
        else
            cudaGetLastError();
    }
    b->outOverride = direct;
    rc = launch_batch(ctx, b, config);
    b->outOverride = nullptr;
    if (rc) return rc;
    trace.mark("kernels enqueued");
    if (!direct)
        if ((rc = download_to_bitmaps(ctx, b, b->out.as<float>(), outputs))) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    trace.mark(direct ? "done (zero-copy output)" : "done (D2H copy)");
    return 0;
}

int b200msdf_generate_batch_u8(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                               const b200msdf_bitmap_u8 *outputs, const b200msdf_config *config) {
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_shapes(shapes))) return rc;
    if ((rc = check_config(config))) return rc;
    const int n = shapes->n_glyphs;
    if (n == 0)
        return 0;
    if (!outputs)
        return fail(B200MSDF_ERR_ARG, "null outputs");
    if (mode < 1 || mode > 4)
        return fail(B200MSDF_ERR_ARG, "mode %d out of range", mode);
    const int N = channels_of(mode);
    std::vector<int32_t> dims((size_t) n*2);
    std::vector<uint8_t> yd((size_t) n);
    bool direct = !getenv("B200MSDF_NO_ZEROCOPY");
    for (int g = 0; g < n; ++g) {
        if (outputs[g].width < 0 || outputs[g].height < 0 || ((size_t) outputs[g].width*outputs[g].height > 0 && !outputs[g].pixels))
            return fail(B200MSDF_ERR_ARG, "bad output bitmap %d", g);
        dims[2*(size_t) g] = outputs[g].width;
        dims[2*(size_t) g+1] = outputs[g].height;
        yd[(size_t) g] = outputs[g].y_downward != 0;
    }
    b200msdf_batch *b = &ctx->pool;
    if ((rc = upload_batch(ctx, b, mode, shapes, xf, xfPerGlyph, dims.data(), 1, yd.data()))) return rc;
    if ((rc = launch_batch(ctx, b, config))) return rc;
    // destinations: the caller's rectangles when they are device-visible pinned memory, else a packed device buffer
    std::vector<ByteTarget> targets((size_t) n);
    for (int g = 0; g < n && direct; ++g) {
        targets[(size_t) g].base = nullptr;
        targets[(size_t) g].rowStride = 0;
        if ((size_t) outputs[g].width*outputs[g].height == 0)
            continue;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, outputs[g].pixels) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
            targets[(size_t) g].base = static_cast<uint8_t *>(attr.devicePointer);
            targets[(size_t) g].rowStride = outputs[g].row_stride;
        } else {
            cudaGetLastError();
            direct = false;
        }
    }
    if (!direct) {
        if ((rc = b->bytes.reserve(b->outFloats+64))) return rc;
        for (int g = 0; g < n; ++g) {
            targets[(size_t) g].base = b->bytes.as<uint8_t>()+b->hGlyphs[(size_t) g].outOffset;
            targets[(size_t) g].rowStride = (long long) b->hGlyphs[(size_t) g].w*N;
        }
    }
    if ((rc = b->byteTargets.reserve(sizeof(ByteTarget)*(size_t) n))) return rc;
    CUDA_TRY(cudaMemcpyAsync(b->byteTargets.p, targets.data(), sizeof(ByteTarget)*(size_t) n, cudaMemcpyHostToDevice, ctx->stream));
    if (b->nTiles) {
        int grid = b->nTiles < ctx->numSMs*8 ? b->nTiles : ctx->numSMs*8;
        const float *src = b->out.as<float>();
        switch (N) {
            case 1: k4_quantize<1><<<grid, CTA_THREADS, 0, ctx->stream>>>(src, b->glyphs.as<GlyphDesc>(), b->tiles.as<TileDesc>(), b->nTiles, b->byteTargets.as<ByteTarget>()); break;
            case 3: k4_quantize<3><<<grid, CTA_THREADS, 0, ctx->stream>>>(src, b->glyphs.as<GlyphDesc>(), b->tiles.as<TileDesc>(), b->nTiles, b->byteTargets.as<ByteTarget>()); break;
            default: k4_quantize<4><<<grid, CTA_THREADS, 0, ctx->stream>>>(src, b->glyphs.as<GlyphDesc>(), b->tiles.as<TileDesc>(), b->nTiles, b->byteTargets.as<ByteTarget>()); break;
        }
        CUDA_TRY(cudaGetLastError());
        ++ctx->launches;
    }
    if (!direct && b->outFloats) {
        if ((rc = ctx->pinOut.reserve(b->outFloats))) return rc;
        CUDA_TRY(cudaMemcpyAsync(ctx->pinOut.p, b->bytes.p, b->outFloats, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        const uint8_t *h = static_cast<const uint8_t *>(ctx->pinOut.p);
        for (int g = 0; g < n; ++g) {
            const GlyphDesc &d = b->hGlyphs[(size_t) g];
            for (int r = 0; r < d.h; ++r)
                memcpy(outputs[g].pixels+(ptrdiff_t) outputs[g].row_stride*r, h+d.outOffset+(size_t) r*d.w*N, (size_t) d.w*N);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200msdf_generate(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shape, const b200msdf_transform *xf, const b200msdf_bitmap *output,
                      const b200msdf_config *config) {
    if (!shape || shape->n_glyphs != 1)
        return fail(B200MSDF_ERR_ARG, "b200msdf_generate takes exactly one shape");
    int rc = b200msdf_generate_batch(ctx, mode, shape, xf, 0, output, config);
    if (rc) return rc;
    if (config->ec_buffer && mode >= 3 && config->ec_mode != B200MSDF_EC_DISABLED && ctx->pool.texels) {
        CUDA_TRY(cudaMemcpyAsync(config->ec_buffer, ctx->pool.stencil.p, ctx->pool.texels, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

static int ec_common(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                     const b200msdf_bitmap *sdfs, const b200msdf_config *config, int shapeless) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (channels != 3 && channels != 4)
        return fail(B200MSDF_ERR_ARG, "error correction needs 3 or 4 channels");
    if ((rc = check_shapes(shapes))) return rc;
    if ((rc = check_config(config))) return rc;
    if ((rc = check_bitmaps(sdfs, shapes->n_glyphs))) return rc;
    if (shapes->n_glyphs == 0 || config->ec_mode == B200MSDF_EC_DISABLED)
        return 0;
    std::vector<int32_t> dims;
    std::vector<uint8_t> yd;
    dims_from_bitmaps(sdfs, shapes->n_glyphs, dims, yd);
    b200msdf_batch *b = &ctx->pool;
    int mode = channels == 3 ? B200MSDF_MSDF : B200MSDF_MTSDF;
    if (shapeless) {
        // msdfErrorCorrectionShapeless works in the sdf's own orientation: present the (empty) shape with the same orientation
        for (int g = 0; g < shapes->n_glyphs; ++g) yd[(size_t) g] = 0;
    }
    if ((rc = upload_batch(ctx, b, mode, shapes, xf, xfPerGlyph, dims.data(), 1, yd.data()))) return rc;
    // gather the caller's fields into the device layout
    if ((rc = ctx->pinIn.reserve(b->outFloats*4))) return rc;
    float *h = static_cast<float *>(ctx->pinIn.p);
    for (int g = 0; g < b->nGlyphs; ++g) {
        const GlyphDesc &d = b->hGlyphs[(size_t) g];
        for (int r = 0; r < d.h; ++r)
            memcpy(h+d.outOffset+(size_t) r*d.w*channels, sdfs[g].pixels+(ptrdiff_t) sdfs[g].row_stride*r, sizeof(float)*(size_t) d.w*channels);
    }
    CUDA_TRY(cudaMemcpyAsync(b->field.p, h, b->outFloats*4, cudaMemcpyHostToDevice, ctx->stream));
    KernelParams P;
    fill_params(ctx, b, config, P);
    P.field = b->field.as<float>();
    P.out = b->out.as<float>();
    P.stencil = b->stencil.as<uint8_t>();
    P.shapeless = shapeless;
    if ((rc = launch_prologue(ctx, b, P))) return rc;
    if ((rc = launch_k2(ctx, b, channels, config->overlap_support != 0, P, b->maxContours))) return rc;
    if ((rc = download_to_bitmaps(ctx, b, b->out.as<float>(), sdfs))) return rc;
    if (!shapeless && config->ec_buffer && shapes->n_glyphs == 1 && b->texels) {
        CUDA_TRY(cudaMemcpyAsync(config->ec_buffer, b->stencil.p, b->texels, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200msdf_error_correction(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                              const b200msdf_bitmap *sdfs, const b200msdf_config *config) {
    return ec_common(ctx, channels, shapes, xf, xfPerGlyph, sdfs, config, 0);
}

int b200msdf_fast_error_correction(b200msdf_ctx *ctx, int channels, int edge, const b200msdf_transform *xf, const b200msdf_bitmap *sdf,
                                   double min_deviation_ratio) {
    static const uint32_t zero[2] = { 0, 0 };
    b200msdf_shapes empty;
    memset(&empty, 0, sizeof(empty));
    empty.n_glyphs = 1;
    empty.glyph_contour_begin = zero;
    empty.contour_segment_begin = zero;
    b200msdf_config cfg;
    b200msdf_default_config(&cfg);
    cfg.ec_mode = edge ? B200MSDF_EC_EDGE_ONLY : B200MSDF_EC_INDISCRIMINATE;   // protectAll or nothing (msdf-error-correction.cpp:51-59)
    cfg.ec_distance_check_mode = B200MSDF_DCM_DO_NOT_CHECK;
    cfg.min_deviation_ratio = min_deviation_ratio;
    return ec_common(ctx, channels, &empty, xf, 0, sdf, &cfg, 1);
}

// Host-buffer scanline passes: upload, gather the caller's fields (sign correction), run, scatter back.
static int scanline_common(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                           const b200msdf_bitmap *bitmaps, float zero, int fillRule, bool rasterizeOnly) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (channels != 1 && channels != 3 && channels != 4)
        return fail(B200MSDF_ERR_ARG, "sign correction needs 1, 3 or 4 channels");
    if (fillRule < 0 || fillRule > 3)
        return fail(B200MSDF_ERR_ARG, "fill rule out of range");
    if ((rc = check_shapes(shapes))) return rc;
    if ((rc = check_bitmaps(bitmaps, shapes->n_glyphs))) return rc;
    if (shapes->n_glyphs == 0)
        return 0;
    std::vector<int32_t> dims;
    std::vector<uint8_t> yd;
    dims_from_bitmaps(bitmaps, shapes->n_glyphs, dims, yd);
    b200msdf_batch *b = &ctx->pool;
    int mode = channels == 1 ? B200MSDF_SDF : channels == 3 ? B200MSDF_MSDF : B200MSDF_MTSDF;
    if ((rc = upload_batch(ctx, b, mode, shapes, xf, xfPerGlyph, dims.data(), 1, yd.data()))) return rc;
    if (b->nTiles == 0)
        return 0;
    b200msdf_config cfg;
    b200msdf_default_config(&cfg);
    KernelParams P;
    fill_params(ctx, b, &cfg, P);
    float *dev = b->out.as<float>();
    if (rasterizeOnly) {
        k5_rasterize<<<b->nTiles, CTA_THREADS, 0, ctx->stream>>>(P, dev, fillRule);
        CUDA_TRY(cudaGetLastError());
        ++ctx->launches;
    } else {
        if ((rc = ctx->pinIn.reserve(b->outFloats*4))) return rc;
        float *h = static_cast<float *>(ctx->pinIn.p);
        for (int g = 0; g < b->nGlyphs; ++g) {
            const GlyphDesc &d = b->hGlyphs[(size_t) g];
            for (int r = 0; r < d.h; ++r)
                memcpy(h+d.outOffset+(size_t) r*d.w*channels, bitmaps[g].pixels+(ptrdiff_t) bitmaps[g].row_stride*r, sizeof(float)*(size_t) d.w*channels);
        }
        CUDA_TRY(cudaMemcpyAsync(dev, h, b->outFloats*4, cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = launch_sign(ctx, b, P, dev, dev, zero, fillRule))) return rc;
    }
    if ((rc = download_to_bitmaps(ctx, b, dev, bitmaps))) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200msdf_sign_correction(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                             const b200msdf_bitmap *sdfs, float sdf_zero_value, int fill_rule) {
    return scanline_common(ctx, channels, shapes, xf, xfPerGlyph, sdfs, sdf_zero_value, fill_rule, false);
}

int b200msdf_rasterize(b200msdf_ctx *ctx, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                       const b200msdf_bitmap *outputs, int fill_rule) {
    return scanline_common(ctx, 1, shapes, xf, xfPerGlyph, outputs, 0.f, fill_rule, true);
}

// Host image <-> packed device block helpers of the render entry points (pool buffers: field = input, out = output)
static int upload_image(b200msdf_ctx *ctx, DevBuf &dst, const b200msdf_bitmap *bm, int N) {
    size_t floats = (size_t) bm->width*bm->height*N;
    int rc = dst.reserve(floats*4+64);
    if (rc) return rc;
    if ((rc = ctx->pinIn.reserve(floats*4+64))) return rc;
    float *h = static_cast<float *>(ctx->pinIn.p);
    for (int r = 0; r < bm->height; ++r)
        memcpy(h+(size_t) r*bm->width*N, bm->pixels+(ptrdiff_t) bm->row_stride*r, sizeof(float)*(size_t) bm->width*N);
    CUDA_TRY(cudaMemcpyAsync(dst.p, h, floats*4, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

static int download_image(b200msdf_ctx *ctx, const DevBuf &src, const b200msdf_bitmap *bm, int N) {
    size_t floats = (size_t) bm->width*bm->height*N;
    int rc = ctx->pinOut.reserve(floats*4+64);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(ctx->pinOut.p, src.p, floats*4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const float *h = static_cast<const float *>(ctx->pinOut.p);
    for (int r = 0; r < bm->height; ++r)
        memcpy(bm->pixels+(ptrdiff_t) bm->row_stride*r, h+(size_t) r*bm->width*N, sizeof(float)*(size_t) bm->width*N);
    return 0;
}

int b200msdf_render_sdf(b200msdf_ctx *ctx, const b200msdf_bitmap *output, int outN, const b200msdf_bitmap *sdf, int sdfN,
                        double lower, double upper, float threshold) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!output || !sdf)
        return fail(B200MSDF_ERR_ARG, "null bitmap");
    if (!((outN == 1 && (sdfN == 1 || sdfN == 3 || sdfN == 4)) || (outN == 3 && (sdfN == 1 || sdfN == 3)) || (outN == 4 && sdfN == 4)))
        return fail(B200MSDF_ERR_ARG, "renderSDF has no overload for %d <- %d channels", outN, sdfN);
    if ((rc = check_bitmaps(output, 1)) || (rc = check_bitmaps(sdf, 1))) return rc;
    if ((size_t) output->width*output->height == 0)
        return 0;
    if ((size_t) sdf->width*sdf->height == 0)
        return fail(B200MSDF_ERR_ARG, "empty distance field");
    b200msdf_batch *b = &ctx->pool;
    if ((rc = upload_image(ctx, b->field, sdf, sdfN))) return rc;
    if ((rc = b->out.reserve((size_t) output->width*output->height*outN*4+64))) return rc;
    RenderParams R;
    R.sdf = b->field.as<float>();
    R.out = b->out.as<float>();
    R.sw = sdf->width; R.sh = sdf->height; R.ow = output->width; R.oh = output->height;
    R.scaleX = (double) sdf->width/output->width;
    R.scaleY = (double) sdf->height/output->height;
    R.hard = lower == upper;
    R.threshold = threshold;
    R.bias = .5f-threshold;
    R.mapScale = 1; R.mapTranslate = 0;
    if (!R.hard) {
        // sdfPxRange *= (output.width+output.height)/(sdf.width+sdf.height); DistanceMapping::inverse(range)
        double f = (double) (output->width+output->height)/(sdf->width+sdf->height);
        lower *= f;
        upper *= f;
        double width = upper-lower;
        R.mapScale = width;
        R.mapTranslate = lower/(width ? width : 1);
    }
    size_t total = (size_t) output->width*output->height;
    int grid = (int) ((total+CTA_THREADS-1)/CTA_THREADS < (size_t) ctx->numSMs*8 ? (total+CTA_THREADS-1)/CTA_THREADS : (size_t) ctx->numSMs*8);
    switch (outN*8+sdfN) {
        case 1*8+1: k6_render_sdf<1, 1><<<grid, CTA_THREADS, 0, ctx->stream>>>(R); break;
        case 3*8+1: k6_render_sdf<1, 3><<<grid, CTA_THREADS, 0, ctx->stream>>>(R); break;
        case 1*8+3: k6_render_sdf<3, 1><<<grid, CTA_THREADS, 0, ctx->stream>>>(R); break;
        case 3*8+3: k6_render_sdf<3, 3><<<grid, CTA_THREADS, 0, ctx->stream>>>(R); break;
        case 1*8+4: k6_render_sdf<4, 1><<<grid, CTA_THREADS, 0, ctx->stream>>>(R); break;
        default:    k6_render_sdf<4, 4><<<grid, CTA_THREADS, 0, ctx->stream>>>(R); break;
    }
    CUDA_TRY(cudaGetLastError());
    ++ctx->launches;
    return download_image(ctx, b->out, output, outN);
}

int b200msdf_simulate_8bit(b200msdf_ctx *ctx, const b200msdf_bitmap *bitmap, int channels) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!bitmap || channels < 1 || channels > 4)
        return fail(B200MSDF_ERR_ARG, "bad bitmap / channel count");
    if ((rc = check_bitmaps(bitmap, 1))) return rc;
    size_t n = (size_t) bitmap->width*bitmap->height*channels;
    if (!n)
        return 0;
    b200msdf_batch *b = &ctx->pool;
    if ((rc = upload_image(ctx, b->out, bitmap, channels))) return rc;
    int grid = (int) ((n+CTA_THREADS-1)/CTA_THREADS < (size_t) ctx->numSMs*8 ? (n+CTA_THREADS-1)/CTA_THREADS : (size_t) ctx->numSMs*8);
    k6_simulate_8bit<<<grid, CTA_THREADS, 0, ctx->stream>>>(b->out.as<float>(), n);
    CUDA_TRY(cudaGetLastError());
    ++ctx->launches;
    return download_image(ctx, b->out, bitmap, channels);
}

int b200msdf_estimate_sdf_error(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                                const b200msdf_bitmap *sdfs, int scanlines_per_row, int fill_rule, double *errors) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (channels != 1 && channels != 3 && channels != 4)
        return fail(B200MSDF_ERR_ARG, "estimateSDFError needs 1, 3 or 4 channels");
    if (fill_rule < 0 || fill_rule > 3)
        return fail(B200MSDF_ERR_ARG, "fill rule out of range");
    if (!errors)
        return fail(B200MSDF_ERR_ARG, "null result array");
    if ((rc = check_shapes(shapes))) return rc;
    if ((rc = check_bitmaps(sdfs, shapes->n_glyphs))) return rc;
    const int n = shapes->n_glyphs;
    if (n == 0)
        return 0;
    std::vector<int32_t> dims;
    std::vector<uint8_t> yd;
    dims_from_bitmaps(sdfs, n, dims, yd);
    b200msdf_batch *b = &ctx->pool;
    const int mode = channels == 1 ? B200MSDF_SDF : channels == 3 ? B200MSDF_MSDF : B200MSDF_MTSDF;
    if ((rc = upload_batch(ctx, b, mode, shapes, xf, xfPerGlyph, dims.data(), 1, nullptr))) return rc;
    // scanlines: (height-1) x scanlinesPerRow per bitmap; sdf-error-estimation.cpp:135-136 returns 0 for degenerate inputs
    std::vector<unsigned> scanBase((size_t) n+1, 0u);
    int maxW = 0;
    for (int g = 0; g < n; ++g) {
        const GlyphDesc &d = b->hGlyphs[(size_t) g];
        const bool empty = d.w <= 1 || d.h <= 1 || scanlines_per_row < 1;
        scanBase[(size_t) g+1] = scanBase[(size_t) g]+(empty ? 0u : (unsigned) (d.h-1)*(unsigned) scanlines_per_row);
        if (d.w > maxW) maxW = d.w;
    }
    const unsigned total = scanBase[(size_t) n];
    for (int g = 0; g < n; ++g)
        errors[g] = 0;
    if (total == 0)
        return 0;
    // the caller's fields in the batch's packed layout (rows as they lie in the caller's memory)
    if ((rc = ctx->pinIn.reserve(b->outFloats*4))) return rc;
    float *h = static_cast<float *>(ctx->pinIn.p);
    for (int g = 0; g < n; ++g) {
        const GlyphDesc &d = b->hGlyphs[(size_t) g];
        for (int r = 0; r < d.h; ++r)
            memcpy(h+d.outOffset+(size_t) r*d.w*channels, sdfs[g].pixels+(ptrdiff_t) sdfs[g].row_stride*r, sizeof(float)*(size_t) d.w*channels);
    }
    CUDA_TRY(cudaMemcpyAsync(b->out.p, h, b->outFloats*4, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (int) (total < (unsigned) ctx->numSMs*16u ? total : (unsigned) ctx->numSMs*16u);
    SdfErrorParams E;
    memset(&E, 0, sizeof(E));
    KernelParams P;
    b200msdf_config cfg;
    b200msdf_default_config(&cfg);
    fill_params(ctx, b, &cfg, P);
    E.raw = P.raw;
    E.glyphs = P.glyphs;
    E.sdf = b->out.as<float>();
    E.nGlyphs = n;
    E.scanlinesPerRow = scanlines_per_row;
    E.fillRule = fill_rule;
    E.slabStride = (size_t) 3*maxW+8;
    E.totalScanlines = total;
    // scratch: scan bases, per-scanline results, per-CTA slabs
    const size_t baseBytes = ((size_t) n+1)*sizeof(unsigned), outBytes = (size_t) total*sizeof(double);
    const size_t slabX = (size_t) grid*E.slabStride*sizeof(double), slabD = (size_t) grid*E.slabStride*sizeof(int);
    const size_t offOut = (baseBytes+15)&~size_t(15), offX = offOut+((outBytes+15)&~size_t(15)), offD = offX+((slabX+15)&~size_t(15));
    if ((rc = b->scratch.reserve(offD+slabD+64))) return rc;
    char *sc = static_cast<char *>(b->scratch.p);
    CUDA_TRY(cudaMemcpyAsync(sc, scanBase.data(), baseBytes, cudaMemcpyHostToDevice, ctx->stream));
    E.scanBase = reinterpret_cast<const unsigned *>(sc);
    E.out = reinterpret_cast<double *>(sc+offOut);
    E.slabX = reinterpret_cast<double *>(sc+offX);
    E.slabDir = reinterpret_cast<int *>(sc+offD);
    switch (channels) {
        case 1: k7_sdf_error<1><<<grid, SE_THREADS, 0, ctx->stream>>>(E); break;
        case 3: k7_sdf_error<3><<<grid, SE_THREADS, 0, ctx->stream>>>(E); break;
        default: k7_sdf_error<4><<<grid, SE_THREADS, 0, ctx->stream>>>(E); break;
    }
    CUDA_TRY(cudaGetLastError());
    ++ctx->launches;
    std::vector<double> per((size_t) total);
    CUDA_TRY(cudaMemcpyAsync(per.data(), E.out, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // error += ... in the reference's order (rows, then sub-rows), then the mean
    for (int g = 0; g < n; ++g) {
        const unsigned s0 = scanBase[(size_t) g], s1 = scanBase[(size_t) g+1];
        if (s1 == s0)
            continue;
        double error = 0;
        for (unsigned k = s0; k < s1; ++k)
            error += per[k];
        errors[g] = error/(double) ((b->hGlyphs[(size_t) g].h-1)*scanlines_per_row);
    }
    return 0;
}

int b200msdf_batch_sign_correction(b200msdf_ctx *ctx, b200msdf_batch *batch, float sdf_zero_value, int fill_rule) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!batch)
        return fail(B200MSDF_ERR_ARG, "null batch");
    if (!batch->launched || !batch->current)
        return fail(B200MSDF_ERR_ARG, "the batch has no result yet: launch it first");
    if (batch->nTiles == 0)
        return 0;
    b200msdf_config cfg;
    b200msdf_default_config(&cfg);
    KernelParams P;
    fill_params(ctx, batch, &cfg, P);
    return launch_sign(ctx, batch, P, batch->current, batch->current, sdf_zero_value, fill_rule);
}

int b200msdf_batch_upload(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                          const int32_t *dims, int dimsPerGlyph, const uint8_t *outYDown, b200msdf_batch **out) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!out)
        return fail(B200MSDF_ERR_ARG, "null batch pointer");
    *out = nullptr;
    b200msdf_batch *b = new (std::nothrow) b200msdf_batch;
    if (!b)
        return fail(B200MSDF_ERR_ALLOC, "out of host memory");
    rc = upload_batch(ctx, b, mode, shapes, xf, xfPerGlyph, dims, dimsPerGlyph, outYDown);
    if (rc) {
        b->release();
        delete b;
        return rc;
    }
    *out = b;
    return 0;
}

int b200msdf_batch_upload_regions(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *xf, int xfPerGlyph,
                                  const int32_t *dims, int dimsPerGlyph, const int32_t *origins, int originsPerGlyph, b200msdf_batch **out) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!out || !origins)
        return fail(B200MSDF_ERR_ARG, "null batch pointer / origins");
    *out = nullptr;
    b200msdf_batch *b = new (std::nothrow) b200msdf_batch;
    if (!b)
        return fail(B200MSDF_ERR_ALLOC, "out of host memory");
    rc = upload_batch(ctx, b, mode, shapes, xf, xfPerGlyph, dims, dimsPerGlyph, nullptr, origins, originsPerGlyph);
    if (rc) {
        b->release();
        delete b;
        return rc;
    }
    *out = b;
    return 0;
}

int b200msdf_batch_launch(b200msdf_ctx *ctx, b200msdf_batch *batch, const b200msdf_config *config) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!batch)
        return fail(B200MSDF_ERR_ARG, "null batch");
    return launch_batch(ctx, batch, config);
}

int b200msdf_batch_launch_stages(b200msdf_ctx *ctx, b200msdf_batch *batch, const b200msdf_config *config, unsigned stages) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!batch)
        return fail(B200MSDF_ERR_ARG, "null batch");
    return launch_batch(ctx, batch, config, stages);
}

float *b200msdf_batch_device_output(const b200msdf_batch *batch) { return batch ? batch->out.as<float>() : nullptr; }
uint8_t *b200msdf_batch_device_stencil(const b200msdf_batch *batch) { return batch ? batch->stencil.as<uint8_t>() : nullptr; }
size_t b200msdf_batch_output_floats(const b200msdf_batch *batch) { return batch ? batch->outFloats : 0; }
size_t b200msdf_batch_glyph_offset(const b200msdf_batch *batch, int glyph) {
    if (!batch || glyph < 0 || glyph >= batch->nGlyphs) return 0;
    return batch->hGlyphs[(size_t) glyph].outOffset;
}

int b200msdf_batch_download(b200msdf_ctx *ctx, b200msdf_batch *batch, float *host_dst) {
    int rc = bind(ctx);
    if (rc) return rc;
    if (!batch || !host_dst)
        return fail(B200MSDF_ERR_ARG, "null batch/destination");
    if (batch->outFloats)
        CUDA_TRY(cudaMemcpyAsync(host_dst, batch->out.p, batch->outFloats*4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

void b200msdf_batch_free(b200msdf_ctx *ctx, b200msdf_batch *batch) {
    if (!batch) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    batch->release();
    delete batch;
}

}  // extern "C"
