// import-font.h — glyph ingest for the B200 generator: msdfgen's font-import interface without FreeType.
//
// msdfgen-ext loads glyph outlines through FreeType (ext/import-font.h, ext/import-font.cpp:62-243: FT_Load_Glyph with
// FT_LOAD_NO_SCALE, FT_Outline_Decompose into move / line / conic / cubic callbacks that append EdgeHolders).  This image has no
// FreeType, and an atlas build wants thousands of glyphs as ONE flat batch for b200msdf_generate_batch rather than thousands of
// Shape objects — so the ingest is native here: a TrueType (`glyf`) reader in libmsdfgen_b200_host that walks the outline points
// exactly as FT_Outline_Decompose does (implied on-curve midpoints with FreeType's integer halving, contours that start on a
// control point, the closing line; the outline moved so that its origin is where the `hmtx` left side bearing puts it) and applies
// msdfgen's callbacks' rules (zero-length edges dropped, degenerate quadratics lowered to lines by the EdgeHolder factory, empty
// contours not kept), straight into flat segment arrays.
//
// The functions below keep msdfgen's names and meaning so that a caller of import-font.h compiles unchanged; FreetypeHandle is an
// empty token.  Covered: TrueType outlines incl. composite glyphs, cmap formats 4 / 6 / 12, hmtx advances, hhea / post metrics,
// kern format 0.  Not covered (loadFont fails or the call returns false): CFF / OpenType-PS outlines, variable-font axes,
// the FreeType-typed helpers (adoptFreetypeFont, readFreetypeOutline).
#pragma once

#include "msdfgen.h"

#include <cstdint>

#define MSDFGEN_LEGACY_FONT_COORDINATE_SCALE (1/64.)

namespace msdfgen {

typedef unsigned unicode_t;

class FreetypeHandle;
class FontHandle;

class GlyphIndex {
public:
    explicit GlyphIndex(unsigned index = 0);
    unsigned getIndex() const;
private:
    unsigned index;
};

/// Global metrics of a typeface (in font units scaled by the chosen FontCoordinateScaling).
struct FontMetrics {
    double emSize;
    double ascenderY, descenderY;
    double lineHeight;
    double underlineY, underlineThickness;
};

/// What glyph coordinates are multiplied by on load.
enum FontCoordinateScaling {
    FONT_SCALING_NONE,              // font units as stored
    FONT_SCALING_EM_NORMALIZED,     // 1 = one em
    FONT_SCALING_LEGACY             // 1/64 (msdfgen before 1.12)
};

FreetypeHandle *initializeFreetype();
void deinitializeFreetype(FreetypeHandle *library);

FontHandle *loadFont(FreetypeHandle *library, const char *filename);
FontHandle *loadFontData(FreetypeHandle *library, const byte *data, int length);
void destroyFont(FontHandle *font);

bool getFontMetrics(FontMetrics &metrics, FontHandle *font, FontCoordinateScaling coordinateScaling = FONT_SCALING_LEGACY);
bool getFontWhitespaceWidth(double &spaceAdvance, double &tabAdvance, FontHandle *font, FontCoordinateScaling coordinateScaling = FONT_SCALING_LEGACY);
bool getGlyphCount(unsigned &output, FontHandle *font);
bool getGlyphIndex(GlyphIndex &glyphIndex, FontHandle *font, unicode_t unicode);
bool loadGlyph(Shape &output, FontHandle *font, GlyphIndex glyphIndex, FontCoordinateScaling coordinateScaling, double *outAdvance = NULL);
bool loadGlyph(Shape &output, FontHandle *font, unicode_t unicode, FontCoordinateScaling coordinateScaling, double *outAdvance = NULL);
bool loadGlyph(Shape &output, FontHandle *font, GlyphIndex glyphIndex, double *outAdvance = NULL);
bool loadGlyph(Shape &output, FontHandle *font, unicode_t unicode, double *outAdvance = NULL);
bool getKerning(double &output, FontHandle *font, GlyphIndex glyphIndex0, GlyphIndex glyphIndex1, FontCoordinateScaling coordinateScaling = FONT_SCALING_LEGACY);
bool getKerning(double &output, FontHandle *font, unicode_t unicode0, unicode_t unicode1, FontCoordinateScaling coordinateScaling = FONT_SCALING_LEGACY);

}  // namespace msdfgen
