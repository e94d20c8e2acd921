/* msdfgen_b200_host.h — C entry points of libmsdfgen_b200_host.so for callers that keep shapes as flat arrays.
 *
 * The host library is the CPU side of the B200 generator: what a msdfgen caller does BEFORE generate* — load glyph outlines,
 * Shape::normalize, edge colouring, windings, bounds — on whole batches, in the wire format of include/msdfgen_b200.h
 * (per segment 8 doubles x0 y0 .. x3 y3, one order byte 1/2/3, one colour byte; cbegin[c] = first segment of contour c,
 * gbegin[g] = first contour of glyph g).  No device work happens here and nothing here computes a distance.
 * C++ callers of msdfgen's own interface use msdfgen_b200/host/msdfgen.h and import-font.h instead.
 *
 * Reference behaviour restated (paths relative to /root/reference): core/Shape.cpp:29-115 (validate, normalize, bounds),
 * core/Contour.cpp:40-86 (winding), core/edge-coloring.cpp:19-258, ext/import-font.cpp:62-161,204-243.
 * All functions return 0 on success unless stated otherwise.
 */
#ifndef MSDFGEN_B200_HOST_H
#define MSDFGEN_B200_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- shape preparation of a batch: Shape::normalize() (doNormalize) then edgeColoringSimple (coloring 1) / InkTrap (2) / none (0)
 * for every glyph.  Two calls because normalisation can change segment counts: _begin prepares and reports the output sizes,
 * _fetch writes the arrays (the INPUT arrays must stay alive until then), _free releases the handle.  colors may be NULL
 * (white); yDown may be NULL. */
typedef struct b200host_prepared b200host_prepared;
int b200host_prepare_begin(int nGlyphs, const uint32_t *gbegin, const uint32_t *cbegin, const double *points, const uint8_t *orders,
                           const uint8_t *colors, const uint8_t *yDown, int doNormalize, int coloring, double angleThreshold,
                           unsigned long long seed, b200host_prepared **handle, int *outContours, int *outSegments);
int b200host_prepare_fetch(b200host_prepared *handle, uint32_t *gbegin, uint32_t *cbegin, double *points, uint8_t *orders, uint8_t *colors);
void b200host_prepare_free(b200host_prepared *handle);

/* Contour::winding() of every contour; Shape::validate() (returns 1 / 0); Shape::getBounds(border, miterLimit, polarity) -> l b r t */
int b200host_windings(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, int *out);
int b200host_validate(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders);
int b200host_bounds(int nContours, const uint32_t *cbegin, const double *points, const uint8_t *orders, double border, double miterLimit,
                    int polarity, double *lbrt);

/* ---- TrueType glyph ingest (no FreeType): loadFontData / getGlyphIndex / loadGlyph for many glyphs at once.
 * b200host_font_open copies the data and returns NULL for anything but a TrueType font with glyf outlines.
 * info[0..7] = numGlyphs, unitsPerEm, ascender, descender, lineHeight, underlinePosition, underlineThickness, has-kern-table.
 * scaling: 0 font units, 1 em-normalised, 2 legacy (1/64) — FontCoordinateScaling.  Glyphs that cannot be loaded come back empty
 * with loaded[g] = 0.  advances / loaded may be NULL.  b200host_font_kerning returns 1 if the pair has a kerning value. */
typedef struct b200host_glyphs b200host_glyphs;
void *b200host_font_open(const uint8_t *data, int length);
void b200host_font_close(void *font);
int b200host_font_info(void *font, int *info);
int b200host_font_glyph_indices(void *font, int n, const uint32_t *unicodes, uint32_t *glyphs);
b200host_glyphs *b200host_font_glyphs_begin(void *font, int n, const uint32_t *glyphs, int scaling, int *outContours, int *outSegments);
int b200host_font_glyphs_fetch(const b200host_glyphs *handle, uint32_t *gbegin, uint32_t *cbegin, double *points, uint8_t *orders,
                               uint8_t *colors, double *advances, uint8_t *loaded);
void b200host_font_glyphs_free(b200host_glyphs *handle);
int b200host_font_kerning(void *font, uint32_t left, uint32_t right, int *value);

#ifdef __cplusplus
}
#endif

#endif
