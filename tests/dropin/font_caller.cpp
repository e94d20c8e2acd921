// font_caller.cpp — a user of msdfgen's font-import interface (ext/import-font.h usage as in /root/reference/README.md:114-142:
// initializeFreetype, loadFont, loadGlyph, shape.normalize, edgeColoringSimple), compiled against the B200 host mirror only (the
// reference's ext/ needs FreeType, which this image does not have).  tests/test_font_ingest.py compares what comes out of the
// Shape objects here with the flat batch ingest.
#include "msdfgen.h"
#include "import-font.h"

#include <cstdint>

using namespace msdfgen;

extern "C" int font_caller_glyph(const char *path, unsigned unicode, int scaling, int prepare, int capSeg, int capContours, uint32_t *cbegin,
                                 double *points, uint8_t *orders, uint8_t *colors, double *advance, double *metrics) {
    FreetypeHandle *ft = initializeFreetype();
    if (!ft)
        return -1;
    int result = -2;
    if (FontHandle *font = loadFont(ft, path)) {
        Shape shape;
        FontMetrics fm;
        if (loadGlyph(shape, font, unicode, (FontCoordinateScaling) scaling, advance) && getFontMetrics(fm, font, (FontCoordinateScaling) scaling)) {
            metrics[0] = fm.emSize; metrics[1] = fm.ascenderY; metrics[2] = fm.descenderY; metrics[3] = fm.lineHeight;
            metrics[4] = fm.underlineY; metrics[5] = fm.underlineThickness;
            if (prepare) {
                shape.normalize();
                edgeColoringSimple(shape, 3.0);
            }
            int s = 0;
            result = (int) shape.contours.size() <= capContours ? 0 : -3;
            cbegin[0] = 0;
            for (size_t c = 0; c < shape.contours.size() && result == 0; ++c) {
                for (size_t i = 0; i < shape.contours[c].edges.size(); ++i, ++s) {
                    if (s >= capSeg) { result = -3; break; }
                    const EdgeSegment *edge = shape.contours[c].edges[i];
                    for (int k = 0; k < 4; ++k) {
                        points[8*s+2*k] = k <= edge->type() ? edge->controlPoints()[k].x : 0;
                        points[8*s+2*k+1] = k <= edge->type() ? edge->controlPoints()[k].y : 0;
                    }
                    orders[s] = (uint8_t) edge->type();
                    colors[s] = (uint8_t) edge->color;
                }
                cbegin[c+1] = (uint32_t) s;
            }
            if (result == 0)
                result = (int) shape.contours.size();
        }
        destroyFont(font);
    }
    deinitializeFreetype(ft);
    return result;
}
