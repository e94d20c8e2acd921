// flat_prep.h — internal to libmsdfgen_b200_host: shape preparation (Shape::normalize + edge colouring) on flat segment lists.
//
// One implementation serves both entry paths: the msdfgen-compatible objects (Shape::normalize, edgeColoringSimple / InkTrap in
// shape_model.cpp flatten a contour, call in here and write the result back) and the batch ingest of flat_c_api.cpp, which never
// builds an object.  Not part of msdfgen.h.
#pragma once

#include "flat_geom.h"

#include <cstdint>
#include <vector>

namespace msdfgen {
namespace flat {

/// Shape::normalize() of ONE closed contour, in place: a single-edge contour becomes its three thirds; edges that meet at a
/// (numerically) 180-degree cusp get their adjacent control points pushed apart (quadratics are promoted to cubics for that).
/// Returns true if anything changed.
bool normalizeContour(std::vector<Seg> &edges);

/// State of an edge colouring across the contours of one shape: the seed-driven colour stream plus scratch space.
class Coloring {
public:
    /// strategy 1 = edgeColoringSimple, 2 = edgeColoringInkTrap
    Coloring(int strategy, double angleThreshold, unsigned long long seed) { reset(strategy, angleThreshold, seed); }
    /// Starts the colouring of another shape (scratch space is kept: a long-lived object colours without allocating).
    void reset(int strategy, double angleThreshold, unsigned long long seed);
    /// Colours one closed contour (e[i].color).  A contour with exactly one corner and fewer than three edges is split in thirds
    /// first so that it can carry three colours; returns true if that happened (the edge list was replaced).
    bool colorContour(std::vector<Seg> &edges);
    /// The same for a contour that is known to need no split (see below), straight from / to wire arrays: segments
    /// [first, first+m) of points / orders, tangents[2i], tangents[2i+1] = unit tangents at the start and end of edge first+i.
    void colorContour(const double *points, const uint8_t *orders, uint32_t first, int m, const Vector2 *tangents, uint8_t *outColors);
    /// Does a contour with these end tangents keep its geometry under colouring (no one-corner contour of fewer than 3 edges)?
    bool keepsGeometry(int m, const Vector2 *tangents) const;

private:
    struct Corner {
        int index;
        double precedingLength;   // estimated length of the spline that ends at this corner (ink trap)
        bool minor;
        EdgeColor color;
    };
    struct Teardrop {
        bool splitNeeded;
        int corner;
        EdgeColor colors[3];
    };
    int strategy;
    double crossThreshold;
    unsigned long long seed;
    EdgeColor color;
    std::vector<Corner> corners;
    std::vector<Vector2> tangents;

    void advance();
    void advance(EdgeColor banned);
    bool sharp(const Vector2 &inDir, const Vector2 &outDir) const;
    template <typename LengthOf, typename SetColor>
    Teardrop paint(int m, const Vector2 *tangents, LengthOf lengthOf, SetColor setColor);
};

/// Unit tangents at both ends of every edge: out[2i] at the start, out[2i+1] at the end of edge i.
void endTangents(const Seg *edges, int n, Vector2 *out);

/// Shape::normalize() (if doNormalize) followed by the colouring (0 none, 1 simple, 2 ink trap) of ONE glyph of a flat batch
/// (contours [c0, c1) of cbegin).  Almost every glyph keeps its geometry: then only colours are computed, straight from the
/// arrays and without allocation — one colour per segment into outColors[s] (s = global segment index), `changed` left empty,
/// returns true.  Otherwise returns false and `changed` / `changedCbegin` (local contour offsets, from 0) hold the glyph's new
/// segments.
bool prepareGlyph(const uint32_t *cbegin, uint32_t c0, uint32_t c1, const double *points, const uint8_t *orders, const uint8_t *inColors,
                  bool doNormalize, int coloring, double angleThreshold, unsigned long long seed, uint8_t *outColors,
                  std::vector<Seg> &changed, std::vector<uint32_t> &changedCbegin, bool allowFastPath = true);

}
}
