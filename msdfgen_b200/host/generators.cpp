// generators.cpp — msdfgen's generator and error-correction entry points, implemented by marshalling the host Shape into
// the flat arrays of include/msdfgen_b200.h and calling the CUDA library.  There is no CPU distance code here.
//
// Replaces (reference paths relative to /root/reference/): core/msdfgen.cpp:78-162 (generate* incl. the two generations of
// compatibility overloads) and core/msdf-error-correction.cpp:61-113.  The *_legacy variants are out of scope.
#include "msdfgen.h"
#include "../../include/msdfgen_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstdint>

namespace msdfgen {

namespace {

thread_local int g_device = 0;
thread_local b200msdf_ctx *g_ctx = NULL;
thread_local int g_ctxDevice = -1;
b200::ErrorHandler g_handler = NULL;

void report(int status) {
    const char *msg = b200msdf_last_error();
    if (g_handler) {
        g_handler(status, msg);
        return;
    }
    // The msdfgen API is void and has no error channel; a silent CPU fallback is ruled out, so fail loudly.
    std::fprintf(stderr, "msdfgen_b200: device generator failed (%d): %s\n", status, msg);
    std::abort();
}

struct ContextReleaser {
    ~ContextReleaser() { if (g_ctx) b200msdf_destroy(g_ctx); g_ctx = NULL; }
};

b200msdf_ctx *context() {
    static thread_local ContextReleaser releaser;
    (void) releaser;
    if (g_ctx && g_ctxDevice != g_device) {
        b200msdf_destroy(g_ctx);
        g_ctx = NULL;
    }
    if (!g_ctx) {
        int rc = b200msdf_create(g_device, &g_ctx);
        if (rc) {
            g_ctx = NULL;
            report(rc);
            return NULL;
        }
        g_ctxDevice = g_device;
    }
    return g_ctx;
}

// Shape -> flat arrays (appended, so several shapes can share one pool)
struct FlatPool {
    std::vector<uint32_t> glyphContourBegin, contourSegmentBegin;
    std::vector<double> points;
    std::vector<uint8_t> orders, colors, yDownward;

    FlatPool() { glyphContourBegin.push_back(0); contourSegmentBegin.push_back(0); }

    void add(const Shape &shape) {
        for (size_t c = 0; c < shape.contours.size(); ++c) {
            const std::vector<EdgeHolder> &edges = shape.contours[c].edges;
            for (size_t i = 0; i < edges.size(); ++i) {
                const EdgeSegment *edge = edges[i];
                int order = edge->type();
                const Point2 *cp = edge->controlPoints();
                for (int k = 0; k < 4; ++k) {
                    points.push_back(k <= order ? cp[k].x : 0);
                    points.push_back(k <= order ? cp[k].y : 0);
                }
                orders.push_back((uint8_t) order);
                colors.push_back((uint8_t) edge->color);
            }
            contourSegmentBegin.push_back((uint32_t) orders.size());
        }
        glyphContourBegin.push_back((uint32_t) (contourSegmentBegin.size()-1));
        yDownward.push_back(shape.inverseYAxis ? 1 : 0);
    }

    b200msdf_shapes view() const {
        b200msdf_shapes s;
        s.n_glyphs = (int32_t) yDownward.size();
        s.n_contours = (int32_t) (contourSegmentBegin.size()-1);
        s.n_segments = (int32_t) orders.size();
        s.glyph_contour_begin = glyphContourBegin.data();
        s.contour_segment_begin = contourSegmentBegin.data();
        s.points = points.data();
        s.orders = orders.data();
        s.colors = colors.data();
        s.y_downward = yDownward.data();
        return s;
    }
};

b200msdf_transform marshal(const SDFTransformation &t) {
    b200msdf_transform x;
    x.scale_x = t.getScale().x; x.scale_y = t.getScale().y;
    x.translate_x = t.getTranslate().x; x.translate_y = t.getTranslate().y;
    x.dm_scale = t.distanceMapping.getScale();
    x.dm_translate = t.distanceMapping.getTranslate();
    return x;
}

template <int N>
b200msdf_bitmap marshal(const BitmapSection<float, N> &b) {
    b200msdf_bitmap m;
    m.pixels = b.pixels;
    m.width = b.width; m.height = b.height;
    m.row_stride = b.rowStride;
    m.y_downward = b.yOrientation == Y_DOWNWARD;
    return m;
}

struct ScanlinePass { bool enabled; int fillRule; float zero; };
thread_local ScanlinePass g_scanline = { false, 0, .5f };

b200msdf_config marshal(bool overlapSupport, const ErrorCorrectionConfig *ec) {
    b200msdf_config c;
    b200msdf_default_config(&c);
    c.overlap_support = overlapSupport ? 1 : 0;
    if (ec) {
        c.ec_mode = (int) ec->mode;
        c.ec_distance_check_mode = (int) ec->distanceCheckMode;
        c.min_deviation_ratio = ec->minDeviationRatio;
        c.min_improve_ratio = ec->minImproveRatio;
        c.ec_buffer = ec->buffer;
    }
    return c;
}

template <int N>
void generateOne(int mode, const BitmapSection<float, N> &output, const Shape &shape, const SDFTransformation &transformation, bool overlapSupport, const ErrorCorrectionConfig *ec) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    FlatPool pool;
    pool.add(shape);
    b200msdf_shapes s = pool.view();
    b200msdf_transform t = marshal(transformation);
    b200msdf_bitmap b = marshal(output);
    b200msdf_config c = marshal(overlapSupport, ec);
    int rc = b200msdf_generate(ctx, mode, &s, &t, &b, &c);
    if (rc)
        report(rc);
}

template <int N>
void generateMany(int mode, const std::vector<BitmapSection<float, N> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, bool overlapSupport, const ErrorCorrectionConfig *ec) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    if (outputs.size() != shapes.size() || (transformations.size() != shapes.size() && transformations.size() != 1)) {
        std::fprintf(stderr, "msdfgen_b200: batch size mismatch (%zu outputs, %zu shapes, %zu transformations)\n", outputs.size(), shapes.size(), transformations.size());
        std::abort();
    }
    FlatPool pool;
    std::vector<b200msdf_bitmap> bitmaps(outputs.size());
    std::vector<b200msdf_transform> xf(transformations.size());
    for (size_t i = 0; i < shapes.size(); ++i) {
        pool.add(*shapes[i]);
        bitmaps[i] = marshal(outputs[i]);
    }
    for (size_t i = 0; i < transformations.size(); ++i)
        xf[i] = marshal(transformations[i]);
    b200msdf_shapes s = pool.view();
    b200msdf_config c = marshal(overlapSupport, ec);
    c.ec_buffer = NULL;
    c.scanline_pass = g_scanline.enabled;
    c.fill_rule = g_scanline.fillRule;
    c.sdf_zero_value = g_scanline.zero;
    int rc = b200msdf_generate_batch(ctx, mode, &s, xf.data(), transformations.size() > 1, bitmaps.data(), &c);
    if (rc)
        report(rc);
}

template <int N>
void generateManyBytes(int mode, const std::vector<BitmapSection<byte, N> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, bool overlapSupport, const ErrorCorrectionConfig *ec) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    if (outputs.size() != shapes.size() || (transformations.size() != shapes.size() && transformations.size() != 1)) {
        std::fprintf(stderr, "msdfgen_b200: batch size mismatch (%zu outputs, %zu shapes, %zu transformations)\n", outputs.size(), shapes.size(), transformations.size());
        std::abort();
    }
    FlatPool pool;
    std::vector<b200msdf_bitmap_u8> bitmaps(outputs.size());
    std::vector<b200msdf_transform> xf(transformations.size());
    for (size_t i = 0; i < shapes.size(); ++i) {
        pool.add(*shapes[i]);
        bitmaps[i].pixels = outputs[i].pixels;
        bitmaps[i].width = outputs[i].width; bitmaps[i].height = outputs[i].height;
        bitmaps[i].row_stride = outputs[i].rowStride;
        bitmaps[i].y_downward = outputs[i].yOrientation == Y_DOWNWARD;
    }
    for (size_t i = 0; i < transformations.size(); ++i)
        xf[i] = marshal(transformations[i]);
    b200msdf_shapes s = pool.view();
    b200msdf_config c = marshal(overlapSupport, ec);
    c.ec_buffer = NULL;
    c.scanline_pass = g_scanline.enabled;
    c.fill_rule = g_scanline.fillRule;
    c.sdf_zero_value = g_scanline.zero;
    int rc = b200msdf_generate_batch_u8(ctx, mode, &s, xf.data(), transformations.size() > 1, bitmaps.data(), &c);
    if (rc)
        report(rc);
}

template <int N>
void correctOne(const BitmapSection<float, N> &sdf, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config) {
    if (config.errorCorrection.mode == ErrorCorrectionConfig::DISABLED)
        return;
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    FlatPool pool;
    pool.add(shape);
    b200msdf_shapes s = pool.view();
    b200msdf_transform t = marshal(transformation);
    b200msdf_bitmap b = marshal(sdf);
    b200msdf_config c = marshal(config.overlapSupport, &config.errorCorrection);
    int rc = b200msdf_error_correction(ctx, N, &s, &t, 0, &b, &c);
    if (rc)
        report(rc);
}

template <int N>
void correctShapeless(const BitmapSection<float, N> &sdf, const SDFTransformation &transformation, double minDeviationRatio, bool edge) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    b200msdf_transform t = marshal(transformation);
    b200msdf_bitmap b = marshal(sdf);
    int rc = b200msdf_fast_error_correction(ctx, N, edge ? 1 : 0, &t, &b, minDeviationRatio);
    if (rc)
        report(rc);
}

template <int N>
void signCorrectOne(const BitmapSection<float, N> &sdf, const Shape &shape, const Projection &projection, float sdfZeroValue, FillRule fillRule) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    FlatPool pool;
    pool.add(shape);
    b200msdf_shapes s = pool.view();
    b200msdf_transform t = marshal(SDFTransformation(projection, Range(1)));
    b200msdf_bitmap b = marshal(sdf);
    int rc = b200msdf_sign_correction(ctx, N, &s, &t, 0, &b, sdfZeroValue, (int) fillRule);
    if (rc)
        report(rc);
}

template <int NO, int NI>
void renderOne(const BitmapSection<float, NO> &output, const BitmapConstSection<float, NI> &sdf, Range sdfPxRange, float sdThreshold) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    b200msdf_bitmap o = marshal(output);
    b200msdf_bitmap i;
    i.pixels = const_cast<float *>(sdf.pixels);       // read only
    i.width = sdf.width; i.height = sdf.height;
    i.row_stride = sdf.rowStride;
    i.y_downward = sdf.yOrientation == Y_DOWNWARD;
    int rc = b200msdf_render_sdf(ctx, &o, NO, &i, NI, sdfPxRange.lower, sdfPxRange.upper, sdThreshold);
    if (rc)
        report(rc);
}

template <int N>
void snapOne(const BitmapSection<float, N> &bitmap) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    // the reference walks width*height*N consecutive floats from the first pixel (render-sdf.cpp:175-191)
    b200msdf_bitmap b;
    b.pixels = bitmap.pixels;
    b.width = bitmap.width; b.height = bitmap.height;
    b.row_stride = N*bitmap.width;
    b.y_downward = 0;
    int rc = b200msdf_simulate_8bit(ctx, &b, N);
    if (rc)
        report(rc);
}

}  // namespace

// ---- generators ------------------------------------------------------------------------------------------------------------

void generateSDF(const BitmapSection<float, 1> &output, const Shape &shape, const SDFTransformation &transformation, const GeneratorConfig &config) {
    generateOne(B200MSDF_SDF, output, shape, transformation, config.overlapSupport, NULL);
}

void generatePSDF(const BitmapSection<float, 1> &output, const Shape &shape, const SDFTransformation &transformation, const GeneratorConfig &config) {
    generateOne(B200MSDF_PSDF, output, shape, transformation, config.overlapSupport, NULL);
}

void generateMSDF(const BitmapSection<float, 3> &output, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config) {
    generateOne(B200MSDF_MSDF, output, shape, transformation, config.overlapSupport, &config.errorCorrection);
}

void generateMTSDF(const BitmapSection<float, 4> &output, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config) {
    generateOne(B200MSDF_MTSDF, output, shape, transformation, config.overlapSupport, &config.errorCorrection);
}

void generateSDF(const BitmapSection<float, 1> &output, const Shape &shape, const Projection &projection, Range range, const GeneratorConfig &config) {
    generateSDF(output, shape, SDFTransformation(projection, range), config);
}

void generatePSDF(const BitmapSection<float, 1> &output, const Shape &shape, const Projection &projection, Range range, const GeneratorConfig &config) {
    generatePSDF(output, shape, SDFTransformation(projection, range), config);
}

void generatePseudoSDF(const BitmapSection<float, 1> &output, const Shape &shape, const Projection &projection, Range range, const GeneratorConfig &config) {
    generatePSDF(output, shape, SDFTransformation(projection, range), config);
}

void generateMSDF(const BitmapSection<float, 3> &output, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config) {
    generateMSDF(output, shape, SDFTransformation(projection, range), config);
}

void generateMTSDF(const BitmapSection<float, 4> &output, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config) {
    generateMTSDF(output, shape, SDFTransformation(projection, range), config);
}

void generateSDF(const BitmapSection<float, 1> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, bool overlapSupport) {
    generateSDF(output, shape, SDFTransformation(Projection(scale, translate), range), GeneratorConfig(overlapSupport));
}

void generatePSDF(const BitmapSection<float, 1> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, bool overlapSupport) {
    generatePSDF(output, shape, SDFTransformation(Projection(scale, translate), range), GeneratorConfig(overlapSupport));
}

void generatePseudoSDF(const BitmapSection<float, 1> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, bool overlapSupport) {
    generatePSDF(output, shape, SDFTransformation(Projection(scale, translate), range), GeneratorConfig(overlapSupport));
}

void generateMSDF(const BitmapSection<float, 3> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, const ErrorCorrectionConfig &errorCorrectionConfig, bool overlapSupport) {
    generateMSDF(output, shape, SDFTransformation(Projection(scale, translate), range), MSDFGeneratorConfig(overlapSupport, errorCorrectionConfig));
}

void generateMTSDF(const BitmapSection<float, 4> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, const ErrorCorrectionConfig &errorCorrectionConfig, bool overlapSupport) {
    generateMTSDF(output, shape, SDFTransformation(Projection(scale, translate), range), MSDFGeneratorConfig(overlapSupport, errorCorrectionConfig));
}

// ---- error correction ------------------------------------------------------------------------------------------------------

void msdfErrorCorrection(const BitmapSection<float, 3> &sdf, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config) { correctOne(sdf, shape, transformation, config); }
void msdfErrorCorrection(const BitmapSection<float, 4> &sdf, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config) { correctOne(sdf, shape, transformation, config); }
void msdfErrorCorrection(const BitmapSection<float, 3> &sdf, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config) { correctOne(sdf, shape, SDFTransformation(projection, range), config); }
void msdfErrorCorrection(const BitmapSection<float, 4> &sdf, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config) { correctOne(sdf, shape, SDFTransformation(projection, range), config); }

void msdfFastDistanceErrorCorrection(const BitmapSection<float, 3> &sdf, const SDFTransformation &transformation, double minDeviationRatio) { correctShapeless(sdf, transformation, minDeviationRatio, false); }
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 4> &sdf, const SDFTransformation &transformation, double minDeviationRatio) { correctShapeless(sdf, transformation, minDeviationRatio, false); }
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 3> &sdf, const Projection &projection, Range range, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(projection, range), minDeviationRatio, false); }
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 4> &sdf, const Projection &projection, Range range, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(projection, range), minDeviationRatio, false); }
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 3> &sdf, Range pxRange, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(Projection(), pxRange), minDeviationRatio, false); }
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 4> &sdf, Range pxRange, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(Projection(), pxRange), minDeviationRatio, false); }

void msdfFastEdgeErrorCorrection(const BitmapSection<float, 3> &sdf, const SDFTransformation &transformation, double minDeviationRatio) { correctShapeless(sdf, transformation, minDeviationRatio, true); }
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 4> &sdf, const SDFTransformation &transformation, double minDeviationRatio) { correctShapeless(sdf, transformation, minDeviationRatio, true); }
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 3> &sdf, const Projection &projection, Range range, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(projection, range), minDeviationRatio, true); }
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 4> &sdf, const Projection &projection, Range range, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(projection, range), minDeviationRatio, true); }
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 3> &sdf, Range pxRange, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(Projection(), pxRange), minDeviationRatio, true); }
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 4> &sdf, Range pxRange, double minDeviationRatio) { correctShapeless(sdf, SDFTransformation(Projection(), pxRange), minDeviationRatio, true); }

// ---- scanline pass ---------------------------------------------------------------------------------------------------------

void rasterize(BitmapSection<float, 1> output, const Shape &shape, const Projection &projection, FillRule fillRule) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return;
    FlatPool pool;
    pool.add(shape);
    b200msdf_shapes s = pool.view();
    b200msdf_transform t = marshal(SDFTransformation(projection, Range(1)));
    b200msdf_bitmap b = marshal(output);
    int rc = b200msdf_rasterize(ctx, &s, &t, 0, &b, (int) fillRule);
    if (rc)
        report(rc);
}

void distanceSignCorrection(BitmapSection<float, 1> sdf, const Shape &shape, const Projection &projection, float sdfZeroValue, FillRule fillRule) { signCorrectOne(sdf, shape, projection, sdfZeroValue, fillRule); }
void distanceSignCorrection(BitmapSection<float, 3> sdf, const Shape &shape, const Projection &projection, float sdfZeroValue, FillRule fillRule) { signCorrectOne(sdf, shape, projection, sdfZeroValue, fillRule); }
void distanceSignCorrection(BitmapSection<float, 4> sdf, const Shape &shape, const Projection &projection, float sdfZeroValue, FillRule fillRule) { signCorrectOne(sdf, shape, projection, sdfZeroValue, fillRule); }

void rasterize(const BitmapSection<float, 1> &output, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule) { rasterize(output, shape, Projection(scale, translate), fillRule); }
void distanceSignCorrection(BitmapSection<float, 1> sdf, const Shape &shape, const Projection &projection, FillRule fillRule) { signCorrectOne(sdf, shape, projection, .5f, fillRule); }
void distanceSignCorrection(BitmapSection<float, 3> sdf, const Shape &shape, const Projection &projection, FillRule fillRule) { signCorrectOne(sdf, shape, projection, .5f, fillRule); }
void distanceSignCorrection(BitmapSection<float, 4> sdf, const Shape &shape, const Projection &projection, FillRule fillRule) { signCorrectOne(sdf, shape, projection, .5f, fillRule); }
void distanceSignCorrection(const BitmapSection<float, 1> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule) { signCorrectOne(sdf, shape, Projection(scale, translate), .5f, fillRule); }
void distanceSignCorrection(const BitmapSection<float, 3> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule) { signCorrectOne(sdf, shape, Projection(scale, translate), .5f, fillRule); }
void distanceSignCorrection(const BitmapSection<float, 4> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule) { signCorrectOne(sdf, shape, Projection(scale, translate), .5f, fillRule); }

// ---- error estimation ------------------------------------------------------------------------------------------------------

template <int N>
double estimateOne(const BitmapConstSection<float, N> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) {
    b200msdf_ctx *ctx = context();
    if (!ctx)
        return 0;
    FlatPool pool;
    pool.add(shape);
    b200msdf_shapes s = pool.view();
    b200msdf_transform t = marshal(SDFTransformation(projection, Range(1)));
    // (the reference reads the field's rows as they lie in memory: no reorientation of the section)
    b200msdf_bitmap b;
    b.pixels = const_cast<float *>(sdf.pixels);
    b.width = sdf.width; b.height = sdf.height; b.row_stride = (int32_t) sdf.rowStride; b.y_downward = 0;
    double error = 0;
    int rc = b200msdf_estimate_sdf_error(ctx, N, &s, &t, 0, &b, scanlinesPerRow, (int) fillRule, &error);
    if (rc)
        report(rc);
    return error;
}

double estimateSDFError(const BitmapConstSection<float, 1> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) { return estimateOne(sdf, shape, projection, scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 3> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) { return estimateOne(sdf, shape, projection, scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 4> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule) { return estimateOne(sdf, shape, projection, scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 1> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule) { return estimateOne(sdf, shape, Projection(scale, translate), scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 3> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule) { return estimateOne(sdf, shape, Projection(scale, translate), scanlinesPerRow, fillRule); }
double estimateSDFError(const BitmapConstSection<float, 4> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule) { return estimateOne(sdf, shape, Projection(scale, translate), scanlinesPerRow, fillRule); }

// ---- preview rendering -----------------------------------------------------------------------------------------------------

void renderSDF(const BitmapSection<float, 1> &output, const BitmapConstSection<float, 1> &sdf, Range sdfPxRange, float sdThreshold) { renderOne(output, sdf, sdfPxRange, sdThreshold); }
void renderSDF(const BitmapSection<float, 3> &output, const BitmapConstSection<float, 1> &sdf, Range sdfPxRange, float sdThreshold) { renderOne(output, sdf, sdfPxRange, sdThreshold); }
void renderSDF(const BitmapSection<float, 1> &output, const BitmapConstSection<float, 3> &sdf, Range sdfPxRange, float sdThreshold) { renderOne(output, sdf, sdfPxRange, sdThreshold); }
void renderSDF(const BitmapSection<float, 3> &output, const BitmapConstSection<float, 3> &sdf, Range sdfPxRange, float sdThreshold) { renderOne(output, sdf, sdfPxRange, sdThreshold); }
void renderSDF(const BitmapSection<float, 1> &output, const BitmapConstSection<float, 4> &sdf, Range sdfPxRange, float sdThreshold) { renderOne(output, sdf, sdfPxRange, sdThreshold); }
void renderSDF(const BitmapSection<float, 4> &output, const BitmapConstSection<float, 4> &sdf, Range sdfPxRange, float sdThreshold) { renderOne(output, sdf, sdfPxRange, sdThreshold); }
void simulate8bit(const BitmapSection<float, 1> &bitmap) { snapOne(bitmap); }
void simulate8bit(const BitmapSection<float, 3> &bitmap) { snapOne(bitmap); }
void simulate8bit(const BitmapSection<float, 4> &bitmap) { snapOne(bitmap); }

// ---- B200 additions --------------------------------------------------------------------------------------------------------

namespace b200 {

void setDevice(int device) { g_device = device; }

ErrorHandler setErrorHandler(ErrorHandler handler) {
    ErrorHandler prev = g_handler;
    g_handler = handler;
    return prev;
}

void setScanlinePass(bool enabled, FillRule fillRule, float sdfZeroValue) {
    g_scanline.enabled = enabled;
    g_scanline.fillRule = (int) fillRule;
    g_scanline.zero = sdfZeroValue;
}

void generateMSDFBatch(const std::vector<BitmapSection<float, 3> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config) {
    generateMany(B200MSDF_MSDF, outputs, shapes, transformations, config.overlapSupport, &config.errorCorrection);
}

void generateMTSDFBatch(const std::vector<BitmapSection<float, 4> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config) {
    generateMany(B200MSDF_MTSDF, outputs, shapes, transformations, config.overlapSupport, &config.errorCorrection);
}

void generateSDFBatch(const std::vector<BitmapSection<float, 1> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const GeneratorConfig &config) {
    generateMany(B200MSDF_SDF, outputs, shapes, transformations, config.overlapSupport, NULL);
}

void generatePSDFBatch(const std::vector<BitmapSection<float, 1> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const GeneratorConfig &config) {
    generateMany(B200MSDF_PSDF, outputs, shapes, transformations, config.overlapSupport, NULL);
}

void generateSDFBatch(const std::vector<BitmapSection<byte, 1> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const GeneratorConfig &config) {
    generateManyBytes(B200MSDF_SDF, outputs, shapes, transformations, config.overlapSupport, NULL);
}

void generateMSDFBatch(const std::vector<BitmapSection<byte, 3> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config) {
    generateManyBytes(B200MSDF_MSDF, outputs, shapes, transformations, config.overlapSupport, &config.errorCorrection);
}

void generateMTSDFBatch(const std::vector<BitmapSection<byte, 4> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config) {
    generateManyBytes(B200MSDF_MTSDF, outputs, shapes, transformations, config.overlapSupport, &config.errorCorrection);
}

}  // namespace b200

}  // namespace msdfgen
