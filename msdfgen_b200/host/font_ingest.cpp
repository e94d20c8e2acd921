// font_ingest.cpp — TrueType glyph ingest without FreeType (see import-font.h): sfnt tables -> outline points -> the edge list
// msdfgen's FreeType callbacks would have produced, as flat segments.
//
// Behaviour restated from /root/reference/ext/import-font.cpp:62-161,229-243 (callbacks, scaling, loadGlyph) and from what
// FT_Load_Glyph(FT_LOAD_NO_SCALE) + FT_Outline_Decompose are documented to do for TrueType outlines (the TrueType / OpenType `glyf`,
// `loca`, `cmap`, `hmtx`, `hhea`, `post`, `kern` table formats are public specifications).  All reads are bounds-checked: a
// malformed font makes loadFont / loadGlyph fail, never read outside the buffer.
#include "import-font.h"
#include "flat_geom.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace msdfgen {

namespace {

struct IPoint { long x, y; bool on; };

// (a * b) / 65536 rounded to nearest, halves away from zero — the 16.16 product used for the transforms of composite glyphs
long mulFix(long a, long b) {
    const long long ab = (long long) a*(long long) b;
    return (long) ((ab+0x8000LL-(ab < 0)) >> 16);
}

struct Table { uint32_t offset, length; };

}  // namespace

class FontHandle {
public:
    std::vector<uint8_t> data;
    Table glyf, loca, hmtx, kern;
    uint32_t cmapSubtable;
    int cmapFormat;
    int unitsPerEm, numGlyphs, numHMetrics;
    bool longLoca;
    int ascender, descender, lineHeight, underlinePosition, underlineThickness;

    bool parse();
    unsigned glyphIndex(unicode_t unicode) const;
    int advance(unsigned glyph) const;
    int leftSideBearing(unsigned glyph) const;
    bool outline(unsigned glyph, std::vector<IPoint> &points, std::vector<int> &contourEnds, int depth = 0) const;
    bool kerning(unsigned left, unsigned right, int &value) const;
    // the edges of a glyph as msdfgen's FreeType callbacks build them; cbegin gets one entry per contour END (appended)
    bool decompose(unsigned glyph, double scale, std::vector<flat::Seg> &segs, std::vector<uint32_t> &cbegin) const;

private:
    bool has(size_t offset, size_t length) const { return offset <= data.size() && length <= data.size()-offset; }
    unsigned u8(size_t o) const { return has(o, 1) ? data[o] : 0u; }
    unsigned u16(size_t o) const { return has(o, 2) ? (unsigned) data[o] << 8 | data[o+1] : 0u; }
    int i16(size_t o) const { return (int16_t) u16(o); }
    uint32_t u32(size_t o) const { return has(o, 4) ? (uint32_t) u16(o) << 16 | u16(o+2) : 0u; }
    bool table(const char *tag, Table &out) const;
    bool glyphRange(unsigned glyph, size_t &begin, size_t &end) const;
};

bool FontHandle::table(const char *tag, Table &out) const {
    const unsigned count = u16(4);
    for (unsigned i = 0; i < count; ++i) {
        const size_t rec = 12+16*(size_t) i;
        if (has(rec, 16) && !memcmp(&data[rec], tag, 4)) {
            out.offset = u32(rec+8);
            out.length = u32(rec+12);
            return has(out.offset, out.length);
        }
    }
    out.offset = out.length = 0;
    return false;
}

bool FontHandle::parse() {
    const uint32_t version = u32(0);
    if (version != 0x00010000u && version != 0x74727565u /* 'true' */)
        return false;                   // CFF-flavoured OpenType ('OTTO'), collections etc.: not TrueType outlines
    Table head, maxp, hhea, cmap, post;
    if (!(table("head", head) && table("maxp", maxp) && table("loca", loca) && table("glyf", glyf) && table("hhea", hhea) && table("hmtx", hmtx)))
        return false;
    if (head.length < 54 || maxp.length < 6 || hhea.length < 36)
        return false;
    unitsPerEm = (int) u16(head.offset+18);
    longLoca = i16(head.offset+50) != 0;
    numGlyphs = (int) u16(maxp.offset+4);
    if ((size_t) loca.length < ((size_t) numGlyphs+1)*(longLoca ? 4 : 2))
        return false;
    ascender = i16(hhea.offset+4);
    descender = i16(hhea.offset+6);
    lineHeight = ascender-descender+i16(hhea.offset+8);
    numHMetrics = (int) u16(hhea.offset+34);
    underlinePosition = underlineThickness = 0;
    if (table("post", post) && post.length >= 12) {
        underlineThickness = i16(post.offset+10);
        underlinePosition = i16(post.offset+8)-underlineThickness/2;       // FreeType reports the centre of the line
    }
    table("kern", kern);
    // Unicode character map: full repertoire (3/10, 0/4+) before BMP (3/1, 0/x)
    cmapSubtable = 0;
    cmapFormat = -1;
    if (table("cmap", cmap) && cmap.length >= 4) {
        int bestRank = 0;
        const unsigned count = u16(cmap.offset+2);
        for (unsigned i = 0; i < count; ++i) {
            const size_t rec = cmap.offset+4+8*(size_t) i;
            const unsigned platform = u16(rec), encoding = u16(rec+2);
            const uint32_t sub = cmap.offset+u32(rec+4);
            const int format = (int) u16(sub);
            if (!(format == 4 || format == 6 || format == 12))
                continue;
            int rank = 0;
            if (platform == 3 && encoding == 10) rank = 4;
            else if (platform == 0 && encoding >= 4) rank = 3;
            else if (platform == 3 && encoding == 1) rank = 2;
            else if (platform == 0) rank = 1;
            if (rank > bestRank) {
                bestRank = rank;
                cmapSubtable = sub;
                cmapFormat = format;
            }
        }
    }
    return unitsPerEm > 0;
}

unsigned FontHandle::glyphIndex(unicode_t unicode) const {
    const size_t t = cmapSubtable;
    unsigned glyph = 0;
    if (cmapFormat == 4 && unicode <= 0xffffu) {
        const unsigned segCount = u16(t+6)/2;
        const size_t ends = t+14, starts = ends+2*(size_t) segCount+2, deltas = starts+2*(size_t) segCount, ranges = deltas+2*(size_t) segCount;
        for (unsigned i = 0; i < segCount; ++i) {
            if (unicode > u16(ends+2*(size_t) i))
                continue;
            const unsigned start = u16(starts+2*(size_t) i);
            if (unicode < start)
                break;
            const unsigned rangeOffset = u16(ranges+2*(size_t) i);
            if (!rangeOffset)
                glyph = (unicode+u16(deltas+2*(size_t) i))&0xffffu;
            else {
                glyph = u16(ranges+2*(size_t) i+rangeOffset+2*(size_t) (unicode-start));
                if (glyph)
                    glyph = (glyph+u16(deltas+2*(size_t) i))&0xffffu;
            }
            break;
        }
    } else if (cmapFormat == 6) {
        const unsigned first = u16(t+6), count = u16(t+8);
        if (unicode >= first && unicode-first < count)
            glyph = u16(t+10+2*(size_t) (unicode-first));
    } else if (cmapFormat == 12) {
        const uint32_t groups = u32(t+12);
        uint32_t lo = 0, hi = groups;
        while (lo < hi) {
            const uint32_t mid = lo+(hi-lo)/2;
            const size_t g = t+16+12*(size_t) mid;
            if (unicode < u32(g)) hi = mid;
            else if (unicode > u32(g+4)) lo = mid+1;
            else { glyph = u32(g+8)+(unicode-u32(g)); break; }
        }
    }
    return glyph < (unsigned) numGlyphs ? glyph : 0;
}

int FontHandle::advance(unsigned glyph) const {
    if (numHMetrics <= 0)
        return 0;
    const unsigned i = glyph < (unsigned) numHMetrics ? glyph : (unsigned) numHMetrics-1;
    return (int) u16(hmtx.offset+4*(size_t) i);
}

int FontHandle::leftSideBearing(unsigned glyph) const {
    if (numHMetrics <= 0)
        return 0;
    if (glyph < (unsigned) numHMetrics)
        return i16(hmtx.offset+4*(size_t) glyph+2);
    return i16(hmtx.offset+4*(size_t) numHMetrics+2*(size_t) (glyph-(unsigned) numHMetrics));
}

bool FontHandle::glyphRange(unsigned glyph, size_t &begin, size_t &end) const {
    if (glyph >= (unsigned) numGlyphs)
        return false;
    if (longLoca) {
        begin = u32(loca.offset+4*(size_t) glyph);
        end = u32(loca.offset+4*(size_t) glyph+4);
    } else {
        begin = 2*(size_t) u16(loca.offset+2*(size_t) glyph);
        end = 2*(size_t) u16(loca.offset+2*(size_t) glyph+2);
    }
    if (begin > end || end > glyf.length)
        return false;
    begin += glyf.offset;
    end += glyf.offset;
    return true;
}

// The glyph's outline in font units: points with their on-curve flags, and the index of the last point of every contour.
// Composite glyphs are assembled from their components (offsets, 2x2 transforms in 16.16 arithmetic, point matching).
bool FontHandle::outline(unsigned glyph, std::vector<IPoint> &points, std::vector<int> &contourEnds, int depth) const {
    size_t g, gEnd;
    if (!glyphRange(glyph, g, gEnd))
        return false;
    if (gEnd == g)
        return true;                    // no outline (space)
    if (gEnd-g < 10)
        return false;
    const int nContours = i16(g);
    size_t p = g+10;
    if (nContours >= 0) {
        const size_t base = points.size();
        if (!has(p, 2*(size_t) nContours+2))
            return false;
        int nPoints = 0;
        for (int c = 0; c < nContours; ++c) {
            const int end = (int) u16(p+2*(size_t) c);
            if (end+1 < nPoints || (c == 0 && end < 0))
                return false;
            nPoints = end+1;
            contourEnds.push_back((int) base+end);
        }
        p += 2*(size_t) nContours;
        p += 2+(size_t) u16(p);         // hinting program
        std::vector<uint8_t> flags((size_t) nPoints);
        for (int i = 0; i < nPoints;) {
            if (!has(p, 1))
                return false;
            const uint8_t f = data[p++];
            flags[(size_t) i++] = f;
            if (f&8) {
                if (!has(p, 1))
                    return false;
                for (unsigned r = data[p++]; r && i < nPoints; --r)
                    flags[(size_t) i++] = f;
            }
        }
        points.resize(base+(size_t) nPoints);
        for (int axis = 0; axis < 2; ++axis) {
            const uint8_t isShort = axis ? 4 : 2, sameOrPositive = axis ? 32 : 16;
            long v = 0;
            for (int i = 0; i < nPoints; ++i) {
                const uint8_t f = flags[(size_t) i];
                if (f&isShort) {
                    if (!has(p, 1))
                        return false;
                    const long d = data[p++];
                    v += (f&sameOrPositive) ? d : -d;
                } else if (!(f&sameOrPositive)) {
                    if (!has(p, 2))
                        return false;
                    v += i16(p);
                    p += 2;
                }
                IPoint &pt = points[base+(size_t) i];
                (axis ? pt.y : pt.x) = v;
                pt.on = (f&1) != 0;
            }
        }
        return true;
    }
    // composite
    if (depth > 8)
        return false;
    const size_t parentBase = points.size();
    for (;;) {
        if (!has(p, 4))
            return false;
        const unsigned flags = u16(p), component = u16(p+2);
        p += 4;
        long arg1, arg2;
        if (flags&1) {
            if (!has(p, 4)) return false;
            arg1 = (flags&2) ? i16(p) : (long) u16(p);
            arg2 = (flags&2) ? i16(p+2) : (long) u16(p+2);
            p += 4;
        } else {
            if (!has(p, 2)) return false;
            arg1 = (flags&2) ? (long) (int8_t) data[p] : (long) data[p];
            arg2 = (flags&2) ? (long) (int8_t) data[p+1] : (long) data[p+1];
            p += 2;
        }
        long xx = 0x10000, xy = 0, yx = 0, yy = 0x10000;       // 16.16
        bool transformed = true;
        if (flags&8) {
            if (!has(p, 2)) return false;
            xx = yy = (long) i16(p)*4;
            p += 2;
        } else if (flags&0x40) {
            if (!has(p, 4)) return false;
            xx = (long) i16(p)*4;
            yy = (long) i16(p+2)*4;
            p += 4;
        } else if (flags&0x80) {
            if (!has(p, 8)) return false;
            xx = (long) i16(p)*4;
            yx = (long) i16(p+2)*4;
            xy = (long) i16(p+4)*4;
            yy = (long) i16(p+6)*4;
            p += 8;
        } else
            transformed = false;
        const size_t childBase = points.size();
        if (!outline(component, points, contourEnds, depth+1))
            return false;
        if (transformed)
            for (size_t i = childBase; i < points.size(); ++i) {
                const long x = points[i].x, y = points[i].y;
                points[i].x = mulFix(x, xx)+mulFix(y, xy);
                points[i].y = mulFix(x, yx)+mulFix(y, yy);
            }
        long dx, dy;
        if (flags&2) {
            dx = arg1;
            dy = arg2;
            if (transformed && (flags&0x800)) {
                // SCALED_COMPONENT_OFFSET (Apple's convention): the offset is stretched by the lengths of the transformed unit
                // vectors (FreeType computes them with its integer hypot; a correctly rounded one stands in here)
                dx = mulFix(dx, std::lround(std::hypot((double) xx, (double) xy)));
                dy = mulFix(dy, std::lround(std::hypot((double) yy, (double) yx)));
            }
        } else {
            // point matching: parent point arg1 and component point arg2 coincide
            if (parentBase+(size_t) arg1 >= childBase || childBase+(size_t) arg2 >= points.size())
                return false;
            dx = points[parentBase+(size_t) arg1].x-points[childBase+(size_t) arg2].x;
            dy = points[parentBase+(size_t) arg1].y-points[childBase+(size_t) arg2].y;
        }
        for (size_t i = childBase; i < points.size(); ++i) {
            points[i].x += dx;
            points[i].y += dy;
        }
        if (!(flags&0x20))
            break;
    }
    return true;
}

bool FontHandle::kerning(unsigned left, unsigned right, int &value) const {
    value = 0;
    if (!kern.length || u16(kern.offset) != 0)
        return false;
    const unsigned nTables = u16(kern.offset+2);
    size_t t = kern.offset+4;
    bool found = false;
    for (unsigned i = 0; i < nTables && has(t, 6); ++i) {
        const size_t length = u16(t+2);
        const unsigned coverage = u16(t+4);
        if ((coverage>>8) == 0 && (coverage&1) && !(coverage&2) && has(t, 14)) {       // format 0, horizontal, kerning values
            const uint32_t key = (uint32_t) left<<16|right;
            uint32_t lo = 0, hi = u16(t+6);
            while (lo < hi) {
                const uint32_t mid = lo+(hi-lo)/2;
                const size_t pair = t+14+6*(size_t) mid;
                const uint32_t k = u32(pair);
                if (k < key) lo = mid+1;
                else if (k > key) hi = mid;
                else {
                    value = (coverage&8) ? i16(pair+4) : value+i16(pair+4);        // override or accumulate
                    found = true;
                    break;
                }
            }
        }
        if (!length)
            break;
        t += length;
    }
    return found;
}

namespace {

// The receiving end of the outline walk: msdfgen's FreeType callbacks (import-font.cpp:66-106) on flat segments.
struct EdgeSink {
    double scale;
    std::vector<flat::Seg> &segs;
    std::vector<uint32_t> &contourEnds;
    size_t contourStart;        // first segment of the open contour
    bool open;
    Point2 position;

    Point2 point(const IPoint &v) const { return Point2(scale*v.x, scale*v.y); }
    void moveTo(const IPoint &to) {
        // a contour that received no edge is reused instead of being kept empty
        if (open && segs.size() > contourStart)
            contourEnds.push_back((uint32_t) segs.size());
        contourStart = segs.size();
        open = true;
        position = point(to);
    }
    void lineTo(const IPoint &to) {
        const Point2 end = point(to);
        if (end != position) {
            segs.push_back(flat::makeSeg(1, position, end));
            position = end;
        }
    }
    void conicTo(const IPoint &control, const IPoint &to) {
        const Point2 end = point(to);
        if (end != position) {
            const Point2 c = point(control);
            // the EdgeHolder factory: a control point on the chord's line makes it a line segment
            if (!crossProduct(c-position, end-c))
                segs.push_back(flat::makeSeg(1, position, end));
            else
                segs.push_back(flat::makeSeg(2, position, c, end));
            position = end;
        }
    }
    void finish() {
        if (open && segs.size() > contourStart)
            contourEnds.push_back((uint32_t) segs.size());
        open = false;
    }
};

IPoint halfway(const IPoint &a, const IPoint &b) {
    IPoint m = { (a.x+b.x)/2, (a.y+b.y)/2, true };      // integer halving, towards zero
    return m;
}

}  // namespace

bool FontHandle::decompose(unsigned glyph, double scale, std::vector<flat::Seg> &segs, std::vector<uint32_t> &cbegin) const {
    std::vector<IPoint> points;
    std::vector<int> ends;
    if (!outline(glyph, points, ends))
        return false;
    // The glyph's origin is where its left side bearing says it is: FreeType moves every TrueType outline by the difference
    // between the xMin of the glyph's header and the lsb of `hmtx` (zero in a well-formed font).
    size_t g0, g1;
    if (glyphRange(glyph, g0, g1) && g1-g0 >= 10) {
        const long shift = (long) i16(g0+2)-leftSideBearing(glyph);
        for (size_t i = 0; shift && i < points.size(); ++i)
            points[i].x -= shift;
    }
    EdgeSink sink = { scale, segs, cbegin, segs.size(), false, Point2() };
    int first = 0;
    for (size_t c = 0; c < ends.size(); ++c) {
        const int last = ends[c];
        if (last < first || (size_t) last >= points.size())
            return false;
        // A contour is a cyclic sequence of on-curve points and quadratic control points; two consecutive control points imply
        // an on-curve point halfway.  It starts at its first point if that is on the curve, else at its last point if that is,
        // else halfway between the two.
        int limit = last, i = first;
        IPoint start = points[(size_t) first];
        if (!start.on) {
            if (points[(size_t) last].on) {
                start = points[(size_t) last];
                --limit;
            } else
                start = halfway(start, points[(size_t) last]);
            --i;                        // the first point is a control point: the walk below must visit it
        }
        sink.moveTo(start);
        bool closed = false;
        while (i < limit) {
            const IPoint &pt = points[(size_t) ++i];
            if (pt.on) {
                sink.lineTo(pt);
                continue;
            }
            IPoint control = pt;
            for (;;) {
                if (i >= limit) {
                    sink.conicTo(control, start);
                    closed = true;
                    break;
                }
                const IPoint &next = points[(size_t) ++i];
                if (next.on) {
                    sink.conicTo(control, next);
                    break;
                }
                sink.conicTo(control, halfway(control, next));
                control = next;
            }
            if (closed)
                break;
        }
        if (!closed)
            sink.lineTo(start);
        first = last+1;
    }
    sink.finish();
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// msdfgen's interface

class FreetypeHandle { };

GlyphIndex::GlyphIndex(unsigned index) : index(index) { }
unsigned GlyphIndex::getIndex() const { return index; }

FreetypeHandle *initializeFreetype() { return new FreetypeHandle; }
void deinitializeFreetype(FreetypeHandle *library) { delete library; }

FontHandle *loadFontData(FreetypeHandle *library, const byte *data, int length) {
    if (!library || !data || length <= 0)
        return NULL;
    FontHandle *font = new FontHandle;
    font->data.assign(data, data+length);
    if (!font->parse()) {
        delete font;
        return NULL;
    }
    return font;
}

FontHandle *loadFont(FreetypeHandle *library, const char *filename) {
    if (!library || !filename)
        return NULL;
    FILE *f = fopen(filename, "rb");
    if (!f)
        return NULL;
    std::vector<byte> bytes;
    byte chunk[65536];
    for (size_t n; (n = fread(chunk, 1, sizeof(chunk), f)) > 0;)
        bytes.insert(bytes.end(), chunk, chunk+n);
    fclose(f);
    return bytes.size() <= 0x7fffffffu ? loadFontData(library, bytes.data(), (int) bytes.size()) : NULL;
}

void destroyFont(FontHandle *font) { delete font; }

static double coordinateScale(const FontHandle *font, FontCoordinateScaling scaling) {
    switch (scaling) {
        case FONT_SCALING_NONE: return 1;
        case FONT_SCALING_EM_NORMALIZED: return 1./(font->unitsPerEm ? font->unitsPerEm : 1);
        case FONT_SCALING_LEGACY: return MSDFGEN_LEGACY_FONT_COORDINATE_SCALE;
    }
    return 1;
}

bool getFontMetrics(FontMetrics &metrics, FontHandle *font, FontCoordinateScaling coordinateScaling) {
    if (!font)
        return false;
    const double scale = coordinateScale(font, coordinateScaling);
    metrics.emSize = scale*font->unitsPerEm;
    metrics.ascenderY = scale*font->ascender;
    metrics.descenderY = scale*font->descender;
    metrics.lineHeight = scale*font->lineHeight;
    metrics.underlineY = scale*font->underlinePosition;
    metrics.underlineThickness = scale*font->underlineThickness;
    return true;
}

bool getFontWhitespaceWidth(double &spaceAdvance, double &tabAdvance, FontHandle *font, FontCoordinateScaling coordinateScaling) {
    if (!font)
        return false;
    const double scale = coordinateScale(font, coordinateScaling);
    spaceAdvance = scale*font->advance(font->glyphIndex(' '));
    tabAdvance = scale*font->advance(font->glyphIndex('\t'));
    return true;
}

bool getGlyphCount(unsigned &output, FontHandle *font) {
    if (!font)
        return false;
    output = (unsigned) font->numGlyphs;
    return true;
}

bool getGlyphIndex(GlyphIndex &glyphIndex, FontHandle *font, unicode_t unicode) {
    glyphIndex = GlyphIndex(font ? font->glyphIndex(unicode) : 0);
    return glyphIndex.getIndex() != 0;
}

bool loadGlyph(Shape &output, FontHandle *font, GlyphIndex glyphIndex, FontCoordinateScaling coordinateScaling, double *outAdvance) {
    if (!font)
        return false;
    const double scale = coordinateScale(font, coordinateScaling);
    std::vector<flat::Seg> segs;
    std::vector<uint32_t> ends;
    if (!font->decompose(glyphIndex.getIndex(), scale, segs, ends))
        return false;
    if (outAdvance)
        *outAdvance = scale*font->advance(glyphIndex.getIndex());
    output.contours.clear();
    output.setYAxisOrientation(Y_UPWARD);
    size_t s = 0;
    for (size_t c = 0; c < ends.size(); ++c) {
        Contour &contour = output.addContour();
        for (; s < ends[c]; ++s) {
            const flat::Seg &seg = segs[s];
            contour.addEdge(seg.order == 1 ? EdgeHolder(seg.p[0], seg.p[1]) : EdgeHolder(seg.p[0], seg.p[1], seg.p[2]));
        }
    }
    return true;
}

bool loadGlyph(Shape &output, FontHandle *font, unicode_t unicode, FontCoordinateScaling coordinateScaling, double *outAdvance) {
    return font && loadGlyph(output, font, GlyphIndex(font->glyphIndex(unicode)), coordinateScaling, outAdvance);
}

bool loadGlyph(Shape &output, FontHandle *font, GlyphIndex glyphIndex, double *outAdvance) {
    return loadGlyph(output, font, glyphIndex, FONT_SCALING_LEGACY, outAdvance);
}

bool loadGlyph(Shape &output, FontHandle *font, unicode_t unicode, double *outAdvance) {
    return loadGlyph(output, font, unicode, FONT_SCALING_LEGACY, outAdvance);
}

bool getKerning(double &output, FontHandle *font, GlyphIndex glyphIndex0, GlyphIndex glyphIndex1, FontCoordinateScaling coordinateScaling) {
    int value = 0;
    if (!font || !font->kerning(glyphIndex0.getIndex(), glyphIndex1.getIndex(), value)) {
        output = 0;
        return false;
    }
    output = coordinateScale(font, coordinateScaling)*value;
    return true;
}

bool getKerning(double &output, FontHandle *font, unicode_t unicode0, unicode_t unicode1, FontCoordinateScaling coordinateScaling) {
    if (!font) {
        output = 0;
        return false;
    }
    return getKerning(output, font, GlyphIndex(font->glyphIndex(unicode0)), GlyphIndex(font->glyphIndex(unicode1)), coordinateScaling);
}

}  // namespace msdfgen

// ---------------------------------------------------------------------------------------------------------------------------------
// extern "C": a font's glyphs as ONE flat batch (the input format of b200msdf_batch_upload / b200host_prepare_begin)

using namespace msdfgen;

extern "C" {

struct b200host_glyphs {
    std::vector<flat::Seg> segs;
    std::vector<uint32_t> gbegin, cbegin;       // gbegin[g] = first contour of glyph g; cbegin[c] = first segment of contour c
    std::vector<double> advances;
    std::vector<uint8_t> loaded;
};

void *b200host_font_open(const uint8_t *data, int length) {
    static FreetypeHandle token;
    return loadFontData(&token, data, length);
}

void b200host_font_close(void *font) { destroyFont(static_cast<FontHandle *>(font)); }

/// info[0..7] = numGlyphs, unitsPerEm, ascender, descender, lineHeight, underlinePosition, underlineThickness, has kern table
int b200host_font_info(void *font, int *info) {
    const FontHandle *f = static_cast<const FontHandle *>(font);
    if (!f)
        return -1;
    info[0] = f->numGlyphs; info[1] = f->unitsPerEm; info[2] = f->ascender; info[3] = f->descender; info[4] = f->lineHeight;
    info[5] = f->underlinePosition; info[6] = f->underlineThickness; info[7] = f->kern.length != 0;
    return 0;
}

int b200host_font_glyph_indices(void *font, int n, const uint32_t *unicodes, uint32_t *glyphs) {
    const FontHandle *f = static_cast<const FontHandle *>(font);
    if (!f)
        return -1;
    for (int i = 0; i < n; ++i)
        glyphs[i] = f->glyphIndex(unicodes[i]);
    return 0;
}

/// Decomposes n glyphs (scaling: 0 none, 1 em-normalised, 2 legacy 1/64).  Glyphs that fail to load become empty shapes with
/// loaded[g] = 0.  Returns a handle for b200host_font_glyphs_fetch / _free, or NULL.
b200host_glyphs *b200host_font_glyphs_begin(void *font, int n, const uint32_t *glyphs, int scaling, int *outContours, int *outSegments) {
    const FontHandle *f = static_cast<const FontHandle *>(font);
    if (!f || n < 0)
        return NULL;
    const double scale = coordinateScale(f, (FontCoordinateScaling) scaling);
    b200host_glyphs *h = new b200host_glyphs;
    h->gbegin.push_back(0);
    h->cbegin.push_back(0);
    for (int g = 0; g < n; ++g) {
        const size_t segs0 = h->segs.size(), contours0 = h->cbegin.size();
        bool ok = f->decompose(glyphs[g], scale, h->segs, h->cbegin);
        if (!ok) {
            h->segs.resize(segs0);
            h->cbegin.resize(contours0);
        }
        h->loaded.push_back(ok);
        h->advances.push_back(scale*f->advance(glyphs[g]));
        h->gbegin.push_back((uint32_t) h->cbegin.size()-1);
    }
    *outContours = (int) h->cbegin.size()-1;
    *outSegments = (int) h->segs.size();
    return h;
}

int b200host_font_glyphs_fetch(const b200host_glyphs *h, uint32_t *gbegin, uint32_t *cbegin, double *points, uint8_t *orders, uint8_t *colors,
                               double *advances, uint8_t *loaded) {
    if (!h)
        return -1;
    memcpy(gbegin, h->gbegin.data(), h->gbegin.size()*sizeof(uint32_t));
    memcpy(cbegin, h->cbegin.data(), h->cbegin.size()*sizeof(uint32_t));
    for (size_t s = 0; s < h->segs.size(); ++s)
        flat::storeSeg(h->segs[s], points, orders, colors, s);
    if (advances)
        memcpy(advances, h->advances.data(), h->advances.size()*sizeof(double));
    if (loaded)
        memcpy(loaded, h->loaded.data(), h->loaded.size());
    return 0;
}

void b200host_font_glyphs_free(b200host_glyphs *h) { delete h; }

int b200host_font_kerning(void *font, uint32_t left, uint32_t right, int *value) {
    const FontHandle *f = static_cast<const FontHandle *>(font);
    return f && f->kerning(left, right, *value) ? 1 : 0;
}

}
