// flat_prep.h — internal to libmsdfgen_b200_host: allocation-free shape preparation on flat arrays (not part of msdfgen.h).
#pragma once

#include <cstdint>

namespace msdfgen {
namespace b200 {

/// Shape::normalize() (if doNormalize) followed by edgeColoringSimple (coloring 1) / edgeColoringInkTrap (coloring 2) / nothing
/// (coloring 0) for ONE glyph given as flat arrays (contours [c0, c1) of cbegin, 8 doubles + order per segment), without building
/// a Shape — valid whenever the preparation leaves the geometry untouched, which is the case for almost every glyph: no
/// single-edge contour to split, no cusp to push apart, no one-corner contour of fewer than three edges.  Writes one colour per
/// segment to outColors[s] (s = global segment index) and returns true; returns false without side effects that matter when the
/// glyph needs the general path (the caller then builds a Shape and runs the member functions).
bool prepareGlyphFlat(const uint32_t *cbegin, uint32_t c0, uint32_t c1, const double *points, const uint8_t *orders, const uint8_t *inColors,
                      bool doNormalize, int coloring, double angleThreshold, unsigned long long seed, uint8_t *outColors);

}
}
