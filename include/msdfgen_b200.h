/*
 * msdfgen_b200.h — C ABI of the B200-native MSDF generator (libmsdfgen_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of Chlumsky/msdfgen: per-pixel signed-distance-field generation plus
 * the MSDF error-correction pass.  The reference exposes that path as C++ free functions (there is no FFI in the
 * reference; substitution is at link time), so each entry point below names the reference interface it replaces:
 *
 *   b200msdf_generate            <-> generateSDF / generatePSDF / generateMSDF / generateMTSDF
 *                                    (/root/reference/msdfgen.h:47-56, core/msdfgen.cpp:78-106)
 *   b200msdf_generate_batch      <-> the loop a caller (e.g. msdf-atlas-gen) runs over those four for a font atlas
 *   b200msdf_error_correction    <-> msdfErrorCorrection            (core/msdf-error-correction.h:15-18, .cpp:61-72)
 *   b200msdf_fast_error_correction <-> msdfFastDistanceErrorCorrection / msdfFastEdgeErrorCorrection (.h:21-34)
 *   b200msdf_batch_*             <-> same path with inputs/outputs resident in HBM (no reference counterpart)
 *
 * Plain pointers and sizes only; no C++ or torch types.  All functions return 0 on success and a negative
 * b200msdf_status on failure; b200msdf_last_error() gives the message.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with B200MSDF_ERR_CUDA.
 *
 * Semantics match the reference bit-for-bit in intent: fp64 distance arithmetic, fp32 texels, identical treatment of
 * Y orientation (output.reorient(shape.getYAxisOrientation()), core/msdfgen.cpp:55), negative row strides, empty
 * shapes (all texels -inf) and empty contours (skipped, core/ShapeDistanceFinder.hpp:19).
 */
#ifndef MSDFGEN_B200_H
#define MSDFGEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MSDF_VERSION 101 /* 0.1.1 */

typedef enum {
    B200MSDF_OK = 0,
    B200MSDF_ERR_ARG = -1,     /* invalid argument */
    B200MSDF_ERR_CUDA = -2,    /* CUDA runtime / no device */
    B200MSDF_ERR_ALLOC = -3    /* host or device allocation failed */
} b200msdf_status;

/* Which generator: selects the edge selector (core/msdfgen.cpp:78-106) and the channel count N = 1,1,3,4. */
typedef enum {
    B200MSDF_SDF = 1,    /* TrueDistanceSelector          -> generateSDF   */
    B200MSDF_PSDF = 2,   /* PerpendicularDistanceSelector -> generatePSDF  */
    B200MSDF_MSDF = 3,   /* MultiDistanceSelector         -> generateMSDF  (+ error correction) */
    B200MSDF_MTSDF = 4   /* MultiAndTrueDistanceSelector  -> generateMTSDF (+ error correction) */
} b200msdf_mode;

/* ErrorCorrectionConfig::Mode / ::DistanceCheckMode, core/generator-config.h:19-36 (same numeric values). */
enum { B200MSDF_EC_DISABLED = 0, B200MSDF_EC_INDISCRIMINATE = 1, B200MSDF_EC_EDGE_PRIORITY = 2, B200MSDF_EC_EDGE_ONLY = 3 };
enum { B200MSDF_DCM_DO_NOT_CHECK = 0, B200MSDF_DCM_CHECK_AT_EDGE = 1, B200MSDF_DCM_ALWAYS_CHECK = 2 };

/*
 * A batch of shapes, flattened from Shape -> Contour -> EdgeHolder -> EdgeSegment
 * (core/Shape.h:15-58, Contour.h:10-32, edge-segments.h:15-144).  One glyph = one Shape.
 */
typedef struct {
    int32_t n_glyphs;
    int32_t n_contours;                      /* total over all glyphs */
    int32_t n_segments;                      /* total over all glyphs */
    const uint32_t *glyph_contour_begin;     /* [n_glyphs+1]   contour range of glyph g (Shape::contours)           */
    const uint32_t *contour_segment_begin;   /* [n_contours+1] segment range of contour c, in Contour::edges order  */
    const double *points;                    /* [n_segments][8] control points p0..p3 (x,y); unused slots ignored   */
    const uint8_t *orders;                   /* [n_segments]   EdgeSegment::type(): 1 linear, 2 quadratic, 3 cubic  */
    const uint8_t *colors;                   /* [n_segments]   EdgeColor bitmask (EdgeColor.h); NULL = all WHITE    */
    const uint8_t *y_downward;               /* [n_glyphs]     Shape::inverseYAxis; NULL = all Y_UPWARD             */
} b200msdf_shapes;

/* SDFTransformation = Projection{scale, translate} + DistanceMapping{scale, translate} (core/Projection.h:31-33,
 * core/DistanceMapping.h:28-30).  For DistanceMapping(Range r): dm_scale = 1/(r.upper-r.lower), dm_translate = -r.lower. */
typedef struct {
    double scale_x, scale_y;
    double translate_x, translate_y;
    double dm_scale, dm_translate;
} b200msdf_transform;

/* BitmapSection<float, N> (core/BitmapRef.hpp:74-111): host memory, row_stride in floats (may be negative). */
typedef struct {
    float *pixels;
    int32_t width, height;
    int32_t row_stride;
    int32_t y_downward;                      /* BitmapSection::yOrientation == Y_DOWNWARD */
} b200msdf_bitmap;

/* MSDFGeneratorConfig = GeneratorConfig{overlapSupport} + ErrorCorrectionConfig (core/generator-config.h:13-64). */
typedef struct {
    int32_t overlap_support;                 /* default 1 */
    int32_t ec_mode;                         /* default B200MSDF_EC_EDGE_PRIORITY */
    int32_t ec_distance_check_mode;          /* default B200MSDF_DCM_CHECK_AT_EDGE */
    double min_deviation_ratio;              /* default 1.11111111111111111 */
    double min_improve_ratio;                /* default 1.11111111111111111 */
    uint8_t *ec_buffer;                      /* ErrorCorrectionConfig::buffer: optional, >= width*height bytes; single-shape calls
                                                only; receives the stencil like the reference's does */
    /* Scanline pass between generation and error correction — what the reference's command line does for `-scanline`
     * (main.cpp:1223-1297: generate with error correction postponed, distanceSignCorrection, msdfErrorCorrection). */
    int32_t scanline_pass;                   /* default 0 */
    int32_t fill_rule;                       /* B200MSDF_FILL_*, default NONZERO */
    float sdf_zero_value;                    /* default 0.5f */
} b200msdf_config;

/* FillRule, core/Scanline.h:9-14 */
#define B200MSDF_FILL_NONZERO 0
#define B200MSDF_FILL_ODD 1
#define B200MSDF_FILL_POSITIVE 2
#define B200MSDF_FILL_NEGATIVE 3

typedef struct b200msdf_ctx b200msdf_ctx;       /* one per (thread, GPU): stream, pooled device + pinned buffers */
typedef struct b200msdf_batch b200msdf_batch;   /* a shape batch resident in HBM with its output buffers */

/* ---- lifecycle ---------------------------------------------------------------------------------------------------- */
int b200msdf_version(void);
const char *b200msdf_last_error(void);                       /* thread-local message of the last failing call */
void b200msdf_default_config(b200msdf_config *cfg);          /* MSDFGeneratorConfig() defaults */
int b200msdf_create(int device, b200msdf_ctx **ctx);         /* device = CUDA ordinal */
void b200msdf_destroy(b200msdf_ctx *ctx);
/* Run on the caller's cudaStream_t (NULL = the CUDA default stream); use_own != 0 switches back to the context's own stream. */
int b200msdf_set_stream(b200msdf_ctx *ctx, void *cuda_stream, int use_own);
int b200msdf_synchronize(b200msdf_ctx *ctx);
/* Number of kernels this context has launched so far (for bench.py's gpu_launches). */
uint64_t b200msdf_launch_count(const b200msdf_ctx *ctx);
/* Pins the calling host thread to the CPUs local to the context's GPU (its NUMA node); pinned buffers allocated by the thread
 * afterwards are local to the GPU too.  Returns the number of CPUs in the new affinity mask (> 0) or a negative error code.
 * Optional: one process per GPU on a multi-socket host should call it once before allocating its host buffers. */
int b200msdf_bind_host_thread(b200msdf_ctx *ctx);
/* Diagnostics: how many times the literal closed-form cubic solver (equation-solver.cpp:34-61 with correctly rounded acos / cos /
 * pow) has run on this device since the last reset — cusp segments, and the guard-band fallbacks of the fast foot-point root
 * path (a polished root within 1e-9 of 0 or 1, or a distance within 1e-13 of a tie).  Synchronises the context's stream. */
int b200msdf_debug_exact_root_calls(b200msdf_ctx *ctx, int reset, uint64_t *calls);
/* Pinned host memory for callers that want DMA-speed transfers. */
int b200msdf_host_alloc(void **ptr, size_t bytes);
void b200msdf_host_free(void *ptr);

/* ---- host-buffer entry points (synchronous; results are in host memory on return) ----------------------------------- */

/* generate{SDF,PSDF,MSDF,MTSDF}(output, shape, transformation, config): shapes->n_glyphs must be 1. */
int b200msdf_generate(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shape, const b200msdf_transform *transformation,
                      const b200msdf_bitmap *output, const b200msdf_config *config);

/* One generate* call per glyph of the batch. transformations: [1] shared or [n_glyphs] (transform_per_glyph != 0);
 * outputs: [n_glyphs] BitmapSections (any sizes/strides/orientations). config->ec_buffer is ignored. */
int b200msdf_generate_batch(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                            int transform_per_glyph, const b200msdf_bitmap *outputs, const b200msdf_config *config);

/* 8-bit atlas output (SURVEY.md §8f "next" #2): the same generation followed by a fused pixelFloatToByte
 * (core/pixel-conversion.hpp:8-10) and placement store — every glyph's bytes land at its rectangle of the caller's atlas page
 * (pixels + row_stride*y + N*x, row_stride in bytes).  When the destinations are pinned, device-accessible host memory the
 * quantise kernel writes them directly; otherwise a packed byte buffer is copied back and scattered on the host.  The transfer
 * is a quarter of the float path's. */
typedef struct {
    uint8_t *pixels;
    int32_t width, height;
    int32_t row_stride;                      /* bytes, may be negative */
    int32_t y_downward;
} b200msdf_bitmap_u8;
int b200msdf_generate_batch_u8(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                               int transform_per_glyph, const b200msdf_bitmap_u8 *outputs, const b200msdf_config *config);

/* msdfErrorCorrection(sdf, shape, transformation, config) on existing N-channel fields (N = 3 or 4), in place. */
int b200msdf_error_correction(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                              int transform_per_glyph, const b200msdf_bitmap *sdfs, const b200msdf_config *config);

/* msdfFastDistanceErrorCorrection (edge == 0) / msdfFastEdgeErrorCorrection (edge != 0): shapeless variants, in place. */
int b200msdf_fast_error_correction(b200msdf_ctx *ctx, int channels, int edge, const b200msdf_transform *transformation,
                                   const b200msdf_bitmap *sdf, double min_deviation_ratio);

/* distanceSignCorrection(sdf, shape, projection, sdfZeroValue, fillRule) (core/rasterization.h:14-18, rasterization.cpp:19-108):
 * makes the sign of every texel agree with the shape's scanline fill. channels 1, 3 or 4; in place; one field per glyph.
 * Only the projection part of the transformations is used. */
int b200msdf_sign_correction(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                             int transform_per_glyph, const b200msdf_bitmap *sdfs, float sdf_zero_value, int fill_rule);

/* rasterize(output, shape, projection, fillRule) (core/rasterization.h:12, rasterization.cpp:9-17): 1.0f where the texel centre
 * is filled, else 0.0f; outputs are 1-channel sections. */
int b200msdf_rasterize(b200msdf_ctx *ctx, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                       int transform_per_glyph, const b200msdf_bitmap *outputs, int fill_rule);

/* renderSDF(output, sdf, sdfPxRange, sdThreshold) (core/render-sdf.h:11-16, render-sdf.cpp:14-173): reconstructs the shape's
 * appearance from a distance field by bilinear sampling.  (out_channels <- sdf_channels) is one of 1<-1, 3<-1, 1<-3, 3<-3, 1<-4,
 * 4<-4 like the reference's overloads; px_range_lower == px_range_upper selects the hard threshold.  Orientation flags are
 * ignored, as in the reference. */
int b200msdf_render_sdf(b200msdf_ctx *ctx, const b200msdf_bitmap *output, int out_channels, const b200msdf_bitmap *sdf, int sdf_channels,
                        double px_range_lower, double px_range_upper, float sd_threshold);

/* estimateSDFError (core/sdf-error-estimation.h:19-21, sdf-error-estimation.cpp:134-154) for every bitmap of a batch: the fraction
 * of the scanned length on which the distance field (1, 3 or 4 channels) and the shape disagree about inside / outside,
 * scanlines_per_row sub-rows per bitmap row.  errors[n_glyphs] receives the values (0 for bitmaps narrower or lower than 2 texels,
 * NaN if one scanline crosses more than 2048 edges or the projection's x scale is not positive — the reference does not terminate
 * for a mirrored projection).  Like the reference, the field's rows are taken as they lie in memory and
 * flipped only by the SHAPE's y_downward flag; bitmap.y_downward is ignored. */
int b200msdf_estimate_sdf_error(b200msdf_ctx *ctx, int channels, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                                int transform_per_glyph, const b200msdf_bitmap *sdfs, int scanlines_per_row, int fill_rule, double *errors);

/* simulate8bit (core/render-sdf.h:19-21): snaps every float of the bitmap to the value an 8-bit image would hold. */
int b200msdf_simulate_8bit(b200msdf_ctx *ctx, const b200msdf_bitmap *bitmap, int channels);

/* ---- HBM-resident batches (asynchronous on the context's stream) ---------------------------------------------------- */

/* Uploads shapes + per-glyph geometry and allocates outputs. dims: [2] shared (w,h) or [2*n_glyphs] (dims_per_glyph != 0).
 * out_y_downward: [n_glyphs] or NULL.  Output layout: glyph g is a packed [h][w][N] float block at
 * b200msdf_batch_glyph_offset(batch, g) floats from the start of the output buffer (offsets are multiples of 4 floats).
 * INPUT LIFETIME: the copies out of the caller's arrays are enqueued on the context's stream; when those arrays are pinned
 * (page-locked) memory the call returns before they have been read, so they must stay untouched until the stream has drained
 * (b200msdf_synchronize, or any later host-buffer call on this context).  Pageable arrays are consumed before the call returns. */
int b200msdf_batch_upload(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                          int transform_per_glyph, const int32_t *dims, int dims_per_glyph, const uint8_t *out_y_downward,
                          b200msdf_batch **batch);
/* Same as b200msdf_batch_upload for REGIONS of larger images (tiling one huge shape across GPUs): origins = [2] shared or
 * [2*n_glyphs] pixel coordinates (x0, y0) of each output block inside its full image; texel (x, y) of the block is evaluated at
 * unproject(x0+x+.5, y0+y+.5), i.e. with full-image pixel coordinates, so tiles are bit-identical to the corresponding part of
 * the untiled image.  Regions require the output orientation to equal the shape's and carry no error correction (a correction
 * pass needs the neighbouring tiles' texels): mode must be SDF or PSDF, or the config's ec_mode DISABLED at launch. */
int b200msdf_batch_upload_regions(b200msdf_ctx *ctx, int mode, const b200msdf_shapes *shapes, const b200msdf_transform *transformations,
                                  int transform_per_glyph, const int32_t *dims, int dims_per_glyph, const int32_t *origins,
                                  int origins_per_glyph, b200msdf_batch **batch);
/* Enqueues the whole path (contour windings -> distance field -> error correction) for the batch. */
int b200msdf_batch_launch(b200msdf_ctx *ctx, b200msdf_batch *batch, const b200msdf_config *config);
/* Same, restricted to some stages (for per-kernel timing): bit 0 windings/prepare, bit 1 distance field, bit 2 error correction,
 * bit 3 scanline pass (only when config->scanline_pass), bits 4 / 5 the two kernels of the distance field one at a time (k1a_filter /
 * k1b_exact; small glyphs only).
 * Later stages read what earlier stages of a previous full launch left in HBM. */
int b200msdf_batch_launch_stages(b200msdf_ctx *ctx, b200msdf_batch *batch, const b200msdf_config *config, unsigned stages);
/* Sign-corrects the batch's most recent result in place on the device (after stages 1|2 it is the raw field, after a full
 * launch the final output).  Not available for multi-channel region batches (the second sweep needs neighbouring tiles). */
int b200msdf_batch_sign_correction(b200msdf_ctx *ctx, b200msdf_batch *batch, float sdf_zero_value, int fill_rule);
float *b200msdf_batch_device_output(const b200msdf_batch *batch);      /* device pointer, valid until batch_free */
uint8_t *b200msdf_batch_device_stencil(const b200msdf_batch *batch);   /* device pointer to per-texel EC flags (1 byte/texel) */
size_t b200msdf_batch_output_floats(const b200msdf_batch *batch);      /* total floats incl. alignment padding */
size_t b200msdf_batch_glyph_offset(const b200msdf_batch *batch, int glyph);
/* D2H of the whole output buffer into host_dst (>= output_floats*4 bytes) and stream sync. */
int b200msdf_batch_download(b200msdf_ctx *ctx, b200msdf_batch *batch, float *host_dst);
void b200msdf_batch_free(b200msdf_ctx *ctx, b200msdf_batch *batch);

#ifdef __cplusplus
}
#endif

#endif
