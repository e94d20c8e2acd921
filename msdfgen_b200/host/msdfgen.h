// msdfgen.h — host-side C++ mirror of the msdfgen public interface for the B200-native generator.
//
// A caller written against Chlumsky/msdfgen's msdfgen.h (namespace msdfgen: Vector2, Shape/Contour/EdgeHolder/EdgeSegment,
// Projection/Range/DistanceMapping/SDFTransformation, Bitmap/BitmapSection, edgeColoringSimple, generateSDF / generatePSDF /
// generateMSDF / generateMTSDF, msdfErrorCorrection…) compiles against this header unchanged and links
// libmsdfgen_b200_host.so + libmsdfgen_b200.so instead of msdfgen-core.  The data model and the O(edges) shape
// preparation (Shape::normalize, Contour::winding, edge colouring, bounds) stay on the host; the per-pixel work — the whole
// of generate* and msdfErrorCorrection* — is marshalled across the C ABI (include/msdfgen_b200.h) into CUDA kernels.
// There is no CPU implementation of the distance path here: EdgeSegment deliberately has no signedDistance().
//
// Interface parity notes (reference paths relative to /root/reference/):
//   value types      core/Vector2.hpp, arithmetics.hpp, Range.hpp, Projection.h, DistanceMapping.h, SDFTransformation.h
//   bitmaps          core/BitmapRef.hpp, Bitmap.h, YAxisOrientation.h, pixel-conversion.hpp
//   shape model      core/EdgeColor.h, edge-segments.h, EdgeHolder.h, Contour.h, Shape.h, edge-coloring.h
//   generators       msdfgen.h:47-69, core/generator-config.h, core/msdf-error-correction.h
// Out of scope: the host-side Scanline class and its users (Shape::scanline, scanlineSDF, Shape::orientContours), the
// equation solvers as a public API, *_legacy generators, file I/O, SVG/FreeType import.
#pragma once

#include <cmath>
#include <cstddef>
#include <vector>

#define MSDFGEN_B200 1

namespace msdfgen {

typedef unsigned char byte;

// ---- arithmetics -------------------------------------------------------------------------------------------------------
template <typename T> inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> inline T max(T a, T b) { return a < b ? b : a; }
template <typename T> inline T median(T a, T b, T c) { return max(min(a, b), min(max(a, b), c)); }
template <typename T, typename S> inline T mix(T a, T b, S weight) { return T((S(1)-weight)*a+weight*b); }
template <typename T> inline T clamp(T n) { return n >= T(0) && n <= T(1) ? n : T(n > T(0)); }
template <typename T> inline T clamp(T n, T b) { return n >= T(0) && n <= b ? n : T(n > T(0))*b; }
template <typename T> inline T clamp(T n, T a, T b) { return n >= a && n <= b ? n : n < a ? a : b; }
template <typename T> inline int sign(T n) { return (T(0) < n)-(n < T(0)); }
template <typename T> inline int nonZeroSign(T n) { return 2*(n > T(0))-1; }

// ---- Vector2 -----------------------------------------------------------------------------------------------------------
struct Vector2 {
    double x, y;
    inline Vector2(double val = 0) : x(val), y(val) { }
    inline Vector2(double x, double y) : x(x), y(y) { }
    inline void reset() { x = 0; y = 0; }
    inline void set(double nx, double ny) { x = nx; y = ny; }
    inline double squaredLength() const { return x*x+y*y; }
    inline double length() const { return std::sqrt(x*x+y*y); }
    /// Unit vector; the zero vector maps to (0, 1), or (0, 0) if allowZero.
    inline Vector2 normalize(bool allowZero = false) const {
        double len = length();
        return len != 0 ? Vector2(x/len, y/len) : Vector2(0, allowZero ? 0 : 1);
    }
    inline Vector2 getOrthogonal(bool polarity = true) const { return polarity ? Vector2(-y, x) : Vector2(y, -x); }
    inline Vector2 getOrthonormal(bool polarity = true, bool allowZero = false) const {
        double len = length();
        if (len != 0)
            return polarity ? Vector2(-y/len, x/len) : Vector2(y/len, -x/len);
        return polarity ? Vector2(0, allowZero ? 0 : 1) : Vector2(0, allowZero ? 0 : -1);
    }
    inline explicit operator bool() const { return x || y; }
    inline Vector2 &operator+=(const Vector2 o) { x += o.x; y += o.y; return *this; }
    inline Vector2 &operator-=(const Vector2 o) { x -= o.x; y -= o.y; return *this; }
    inline Vector2 &operator*=(const Vector2 o) { x *= o.x; y *= o.y; return *this; }
    inline Vector2 &operator/=(const Vector2 o) { x /= o.x; y /= o.y; return *this; }
    inline Vector2 &operator*=(double v) { x *= v; y *= v; return *this; }
    inline Vector2 &operator/=(double v) { x /= v; y /= v; return *this; }
};
typedef Vector2 Point2;

inline double dotProduct(const Vector2 a, const Vector2 b) { return a.x*b.x+a.y*b.y; }
inline double crossProduct(const Vector2 a, const Vector2 b) { return a.x*b.y-a.y*b.x; }
inline bool operator==(const Vector2 a, const Vector2 b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(const Vector2 a, const Vector2 b) { return a.x != b.x || a.y != b.y; }
inline Vector2 operator+(const Vector2 v) { return v; }
inline Vector2 operator-(const Vector2 v) { return Vector2(-v.x, -v.y); }
inline bool operator!(const Vector2 v) { return !v.x && !v.y; }
inline Vector2 operator+(const Vector2 a, const Vector2 b) { return Vector2(a.x+b.x, a.y+b.y); }
inline Vector2 operator-(const Vector2 a, const Vector2 b) { return Vector2(a.x-b.x, a.y-b.y); }
inline Vector2 operator*(const Vector2 a, const Vector2 b) { return Vector2(a.x*b.x, a.y*b.y); }
inline Vector2 operator/(const Vector2 a, const Vector2 b) { return Vector2(a.x/b.x, a.y/b.y); }
inline Vector2 operator*(double a, const Vector2 b) { return Vector2(a*b.x, a*b.y); }
inline Vector2 operator/(double a, const Vector2 b) { return Vector2(a/b.x, a/b.y); }
inline Vector2 operator*(const Vector2 a, double b) { return Vector2(a.x*b, a.y*b); }
inline Vector2 operator/(const Vector2 a, double b) { return Vector2(a.x/b, a.y/b); }

// ---- SignedDistance (part of the data model; the distance path itself runs on the device) -----------------------------
class SignedDistance {
public:
    double distance, dot;
    inline SignedDistance() : distance(-1.7976931348623157e308), dot(0) { }
    inline SignedDistance(double dist, double d) : distance(dist), dot(d) { }
};
inline bool operator<(const SignedDistance a, const SignedDistance b) { return std::fabs(a.distance) < std::fabs(b.distance) || (std::fabs(a.distance) == std::fabs(b.distance) && a.dot < b.dot); }
inline bool operator>(const SignedDistance a, const SignedDistance b) { return std::fabs(a.distance) > std::fabs(b.distance) || (std::fabs(a.distance) == std::fabs(b.distance) && a.dot > b.dot); }
inline bool operator<=(const SignedDistance a, const SignedDistance b) { return std::fabs(a.distance) < std::fabs(b.distance) || (std::fabs(a.distance) == std::fabs(b.distance) && a.dot <= b.dot); }
inline bool operator>=(const SignedDistance a, const SignedDistance b) { return std::fabs(a.distance) > std::fabs(b.distance) || (std::fabs(a.distance) == std::fabs(b.distance) && a.dot >= b.dot); }

// ---- Range / Projection / DistanceMapping / SDFTransformation --------------------------------------------------------------
struct Range {
    double lower, upper;
    inline Range(double symmetricalWidth = 0) : lower(-.5*symmetricalWidth), upper(.5*symmetricalWidth) { }
    inline Range(double lowerBound, double upperBound) : lower(lowerBound), upper(upperBound) { }
    inline Range &operator*=(double f) { lower *= f; upper *= f; return *this; }
    inline Range &operator/=(double d) { lower /= d; upper /= d; return *this; }
    inline Range operator*(double f) const { return Range(lower*f, upper*f); }
    inline Range operator/(double d) const { return Range(lower/d, upper/d); }
};
inline Range operator*(double f, const Range &r) { return Range(f*r.lower, f*r.upper); }

class Projection {
public:
    inline Projection() : scale(1), translate(0) { }
    inline Projection(const Vector2 &scale, const Vector2 &translate) : scale(scale), translate(translate) { }
    inline Point2 project(const Point2 &coord) const { return scale*(coord+translate); }
    inline Point2 unproject(const Point2 &coord) const { return coord/scale-translate; }
    inline Vector2 projectVector(const Vector2 &v) const { return scale*v; }
    inline Vector2 unprojectVector(const Vector2 &v) const { return v/scale; }
    inline double projectX(double x) const { return scale.x*(x+translate.x); }
    inline double projectY(double y) const { return scale.y*(y+translate.y); }
    inline double unprojectX(double x) const { return x/scale.x-translate.x; }
    inline double unprojectY(double y) const { return y/scale.y-translate.y; }
    // B200 additions: the raw parameters, which cross the C ABI verbatim
    inline const Vector2 &getScale() const { return scale; }
    inline const Vector2 &getTranslate() const { return translate; }
private:
    Vector2 scale, translate;
};

class DistanceMapping {
public:
    class Delta {
    public:
        double value;
        inline explicit Delta(double distanceDelta) : value(distanceDelta) { }
        inline operator double() const { return value; }
    };
    static inline DistanceMapping inverse(Range range) {
        double width = range.upper-range.lower;
        return DistanceMapping(width, range.lower/(width ? width : 1));
    }
    inline DistanceMapping() : scale(1), translate(0) { }
    inline DistanceMapping(Range range) : scale(1/(range.upper-range.lower)), translate(-range.lower) { }
    inline double operator()(double d) const { return scale*(d+translate); }
    inline double operator()(Delta d) const { return scale*d.value; }
    inline DistanceMapping inverse() const { return DistanceMapping(1/scale, -scale*translate); }
    // B200 additions
    inline double getScale() const { return scale; }
    inline double getTranslate() const { return translate; }
private:
    double scale, translate;
    inline DistanceMapping(double scale, double translate) : scale(scale), translate(translate) { }
};

class SDFTransformation : public Projection {
public:
    DistanceMapping distanceMapping;
    inline SDFTransformation() { }
    inline SDFTransformation(const Projection &projection, const DistanceMapping &distanceMapping) : Projection(projection), distanceMapping(distanceMapping) { }
};

// ---- bitmaps -----------------------------------------------------------------------------------------------------------
enum YAxisOrientation { Y_UPWARD, Y_DOWNWARD };

}  // namespace msdfgen

#define MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION msdfgen::Y_UPWARD
#define MSDFGEN_Y_AXIS_NONDEFAULT_ORIENTATION (MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION == msdfgen::Y_DOWNWARD ? msdfgen::Y_UPWARD : msdfgen::Y_DOWNWARD)

namespace msdfgen {

template <typename T, int N = 1> struct BitmapRef;
template <typename T, int N = 1> struct BitmapConstRef;
template <typename T, int N = 1> struct BitmapSection;
template <typename T, int N = 1> struct BitmapConstSection;

#define MSDFGEN_B200_SECTION_OPS(SectionType, PtrType) \
    inline PtrType *operator()(int x, int y) const { return pixels+rowStride*y+N*x; } \
    inline SectionType<T, N> getSection(int xMin, int yMin, int xMax, int yMax) const { return SectionType<T, N>(pixels+rowStride*yMin+N*xMin, xMax-xMin, yMax-yMin, rowStride, yOrientation); } \
    inline BitmapConstSection<T, N> getConstSection(int xMin, int yMin, int xMax, int yMax) const { return BitmapConstSection<T, N>(pixels+rowStride*yMin+N*xMin, xMax-xMin, yMax-yMin, rowStride, yOrientation); } \
    inline void reorient(YAxisOrientation newOrientation) { \
        if (yOrientation != newOrientation) { pixels += rowStride*(height-1); rowStride = -rowStride; yOrientation = newOrientation; } \
    }

/// Reference to a contiguous bitmap.
template <typename T, int N>
struct BitmapRef {
    T *pixels;
    int width, height;
    YAxisOrientation yOrientation;
    inline BitmapRef() : pixels(NULL), width(0), height(0), yOrientation(MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) { }
    inline BitmapRef(T *pixels, int width, int height, YAxisOrientation yOrientation = MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) : pixels(pixels), width(width), height(height), yOrientation(yOrientation) { }
    inline T *operator()(int x, int y) const { return pixels+N*(width*y+x); }
    inline BitmapSection<T, N> getSection(int xMin, int yMin, int xMax, int yMax) const { return BitmapSection<T, N>(pixels+N*(width*yMin+xMin), xMax-xMin, yMax-yMin, N*width, yOrientation); }
    inline BitmapConstSection<T, N> getConstSection(int xMin, int yMin, int xMax, int yMax) const { return BitmapConstSection<T, N>(pixels+N*(width*yMin+xMin), xMax-xMin, yMax-yMin, N*width, yOrientation); }
};

template <typename T, int N>
struct BitmapConstRef {
    const T *pixels;
    int width, height;
    YAxisOrientation yOrientation;
    inline BitmapConstRef() : pixels(NULL), width(0), height(0), yOrientation(MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) { }
    inline BitmapConstRef(const T *pixels, int width, int height, YAxisOrientation yOrientation = MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) : pixels(pixels), width(width), height(height), yOrientation(yOrientation) { }
    inline BitmapConstRef(const BitmapRef<T, N> &o) : pixels(o.pixels), width(o.width), height(o.height), yOrientation(o.yOrientation) { }
    inline const T *operator()(int x, int y) const { return pixels+N*(width*y+x); }
    inline BitmapConstSection<T, N> getSection(int xMin, int yMin, int xMax, int yMax) const { return BitmapConstSection<T, N>(pixels+N*(width*yMin+xMin), xMax-xMin, yMax-yMin, N*width, yOrientation); }
    inline BitmapConstSection<T, N> getConstSection(int xMin, int yMin, int xMax, int yMax) const { return getSection(xMin, yMin, xMax, yMax); }
};

/// Reference to a rectangular section of a bitmap: rows are rowStride elements apart (may be negative).
template <typename T, int N>
struct BitmapSection {
    T *pixels;
    int width, height;
    int rowStride;
    YAxisOrientation yOrientation;
    inline BitmapSection() : pixels(NULL), width(0), height(0), rowStride(0), yOrientation(MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) { }
    inline BitmapSection(T *pixels, int width, int height, YAxisOrientation yOrientation = MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) : pixels(pixels), width(width), height(height), rowStride(N*width), yOrientation(yOrientation) { }
    inline BitmapSection(T *pixels, int width, int height, int rowStride, YAxisOrientation yOrientation = MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) : pixels(pixels), width(width), height(height), rowStride(rowStride), yOrientation(yOrientation) { }
    inline BitmapSection(const BitmapRef<T, N> &o) : pixels(o.pixels), width(o.width), height(o.height), rowStride(N*o.width), yOrientation(o.yOrientation) { }
    MSDFGEN_B200_SECTION_OPS(BitmapSection, T)
};

template <typename T, int N>
struct BitmapConstSection {
    const T *pixels;
    int width, height;
    int rowStride;
    YAxisOrientation yOrientation;
    inline BitmapConstSection() : pixels(NULL), width(0), height(0), rowStride(0), yOrientation(MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) { }
    inline BitmapConstSection(const T *pixels, int width, int height, YAxisOrientation yOrientation = MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) : pixels(pixels), width(width), height(height), rowStride(N*width), yOrientation(yOrientation) { }
    inline BitmapConstSection(const T *pixels, int width, int height, int rowStride, YAxisOrientation yOrientation = MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) : pixels(pixels), width(width), height(height), rowStride(rowStride), yOrientation(yOrientation) { }
    inline BitmapConstSection(const BitmapRef<T, N> &o) : pixels(o.pixels), width(o.width), height(o.height), rowStride(N*o.width), yOrientation(o.yOrientation) { }
    inline BitmapConstSection(const BitmapConstRef<T, N> &o) : pixels(o.pixels), width(o.width), height(o.height), rowStride(N*o.width), yOrientation(o.yOrientation) { }
    inline BitmapConstSection(const BitmapSection<T, N> &o) : pixels(o.pixels), width(o.width), height(o.height), rowStride(o.rowStride), yOrientation(o.yOrientation) { }
    MSDFGEN_B200_SECTION_OPS(BitmapConstSection, const T)
    inline BitmapConstSection<T, N> getSection(int xMin, int yMin, int xMax, int yMax, int) const { return getConstSection(xMin, yMin, xMax, yMax); }
};

#undef MSDFGEN_B200_SECTION_OPS

/// An owning 2D image with N channels of T.
template <typename T, int N = 1>
class Bitmap {
public:
    Bitmap() : pixels(NULL), w(0), h(0), yOrientation(MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) { }
    Bitmap(int width, int height, YAxisOrientation yOrientation = MSDFGEN_Y_AXIS_DEFAULT_ORIENTATION) : w(width), h(height), yOrientation(yOrientation) { pixels = new T[size_t(N)*w*h]; }
    explicit Bitmap(const BitmapConstRef<T, N> &o) : w(o.width), h(o.height), yOrientation(o.yOrientation) { pixels = new T[size_t(N)*w*h]; copyRows(o.pixels, N*o.width); }
    explicit Bitmap(const BitmapConstSection<T, N> &o) : w(o.width), h(o.height), yOrientation(o.yOrientation) { pixels = new T[size_t(N)*w*h]; copyRows(o.pixels, o.rowStride); }
    Bitmap(const Bitmap<T, N> &o) : w(o.w), h(o.h), yOrientation(o.yOrientation) { pixels = new T[size_t(N)*w*h]; copyRows(o.pixels, N*o.w); }
    Bitmap(Bitmap<T, N> &&o) : pixels(o.pixels), w(o.w), h(o.h), yOrientation(o.yOrientation) { o.pixels = NULL; o.w = 0; o.h = 0; }
    ~Bitmap() { delete[] pixels; }
    Bitmap<T, N> &operator=(const BitmapConstRef<T, N> &o) { assign(o.pixels, o.width, o.height, N*o.width, o.yOrientation); return *this; }
    Bitmap<T, N> &operator=(const BitmapConstSection<T, N> &o) { assign(o.pixels, o.width, o.height, o.rowStride, o.yOrientation); return *this; }
    Bitmap<T, N> &operator=(const Bitmap<T, N> &o) { if (this != &o) assign(o.pixels, o.w, o.h, N*o.w, o.yOrientation); return *this; }
    Bitmap<T, N> &operator=(Bitmap<T, N> &&o) {
        if (this != &o) { delete[] pixels; pixels = o.pixels; w = o.w; h = o.h; yOrientation = o.yOrientation; o.pixels = NULL; o.w = 0; o.h = 0; }
        return *this;
    }
    int width() const { return w; }
    int height() const { return h; }
    T *operator()(int x, int y) { return pixels+N*(size_t(w)*y+x); }
    const T *operator()(int x, int y) const { return pixels+N*(size_t(w)*y+x); }
    explicit operator T *() { return pixels; }
    explicit operator const T *() const { return pixels; }
    operator BitmapRef<T, N>() { return BitmapRef<T, N>(pixels, w, h, yOrientation); }
    operator BitmapConstRef<T, N>() const { return BitmapConstRef<T, N>(pixels, w, h, yOrientation); }
    operator BitmapSection<T, N>() { return BitmapSection<T, N>(pixels, w, h, yOrientation); }
    operator BitmapConstSection<T, N>() const { return BitmapConstSection<T, N>(pixels, w, h, yOrientation); }
    BitmapSection<T, N> getSection(int xMin, int yMin, int xMax, int yMax) { return BitmapSection<T, N>(pixels+N*(size_t(w)*yMin+xMin), xMax-xMin, yMax-yMin, N*w, yOrientation); }
    BitmapConstSection<T, N> getConstSection(int xMin, int yMin, int xMax, int yMax) const { return BitmapConstSection<T, N>(pixels+N*(size_t(w)*yMin+xMin), xMax-xMin, yMax-yMin, N*w, yOrientation); }
private:
    T *pixels;
    int w, h;
    YAxisOrientation yOrientation;
    void copyRows(const T *src, int srcStride) {
        for (int y = 0; y < h; ++y)
            for (int i = 0; i < N*w; ++i)
                pixels[size_t(N)*w*y+i] = src[ptrdiff_t(srcStride)*y+i];
    }
    void assign(const T *src, int nw, int nh, int srcStride, YAxisOrientation o) {
        if (nw != w || nh != h) { delete[] pixels; w = nw; h = nh; pixels = new T[size_t(N)*w*h]; }
        yOrientation = o;
        copyRows(src, srcStride);
    }
};

inline byte pixelFloatToByte(float x) { return byte(~int(255.5f-255.f*clamp(x))); }
inline float pixelByteToFloat(byte x) { return 1.f/255.f*float(x); }

/// Bilinear sample of a bitmap at pos (texel centres at half-integers).
template <typename T, int N>
inline void interpolate(T *output, const BitmapConstSection<T, N> &bitmap, Point2 pos) {
    pos.x = clamp(pos.x, double(bitmap.width));
    pos.y = clamp(pos.y, double(bitmap.height));
    pos -= .5;
    int l = (int) std::floor(pos.x), b = (int) std::floor(pos.y);
    int r = l+1, t = b+1;
    double lr = pos.x-l, bt = pos.y-b;
    l = clamp(l, bitmap.width-1), r = clamp(r, bitmap.width-1);
    b = clamp(b, bitmap.height-1), t = clamp(t, bitmap.height-1);
    for (int i = 0; i < N; ++i)
        output[i] = mix(mix(bitmap(l, b)[i], bitmap(r, b)[i], lr), mix(bitmap(l, t)[i], bitmap(r, t)[i], lr), bt);
}

// ---- fill rule (core/Scanline.h:9-14) ---------------------------------------------------------------------------------------
/// How the winding total along a scanline decides inside / outside.
enum FillRule {
    FILL_NONZERO,
    FILL_ODD,           // even-odd
    FILL_POSITIVE,
    FILL_NEGATIVE
};

// ---- shape model -------------------------------------------------------------------------------------------------------
enum EdgeColor { BLACK = 0, RED = 1, GREEN = 2, YELLOW = 3, BLUE = 4, MAGENTA = 5, CYAN = 6, WHITE = 7 };

/// An abstract edge segment (a Bezier curve of order type()).
class EdgeSegment {
public:
    EdgeColor color;
    static EdgeSegment *create(Point2 p0, Point2 p1, EdgeColor edgeColor = WHITE);
    static EdgeSegment *create(Point2 p0, Point2 p1, Point2 p2, EdgeColor edgeColor = WHITE);
    static EdgeSegment *create(Point2 p0, Point2 p1, Point2 p2, Point2 p3, EdgeColor edgeColor = WHITE);
    EdgeSegment(EdgeColor edgeColor = WHITE) : color(edgeColor) { }
    virtual ~EdgeSegment() { }
    virtual EdgeSegment *clone() const = 0;
    virtual int type() const = 0;
    virtual const Point2 *controlPoints() const = 0;
    virtual Point2 point(double param) const = 0;
    virtual Vector2 direction(double param) const = 0;
    virtual Vector2 directionChange(double param) const = 0;
    virtual void bound(double &xMin, double &yMin, double &xMax, double &yMax) const = 0;
    virtual void reverse() = 0;
    virtual void moveStartPoint(Point2 to) = 0;
    virtual void moveEndPoint(Point2 to) = 0;
    virtual void splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const = 0;
};

class LinearSegment : public EdgeSegment {
public:
    enum EdgeType { EDGE_TYPE = 1 };
    Point2 p[2];
    LinearSegment(Point2 p0, Point2 p1, EdgeColor edgeColor = WHITE);
    LinearSegment *clone() const;
    int type() const;
    const Point2 *controlPoints() const;
    Point2 point(double param) const;
    Vector2 direction(double param) const;
    Vector2 directionChange(double param) const;
    double length() const;
    void bound(double &xMin, double &yMin, double &xMax, double &yMax) const;
    void reverse();
    void moveStartPoint(Point2 to);
    void moveEndPoint(Point2 to);
    void splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const;
};

class QuadraticSegment : public EdgeSegment {
public:
    enum EdgeType { EDGE_TYPE = 2 };
    Point2 p[3];
    QuadraticSegment(Point2 p0, Point2 p1, Point2 p2, EdgeColor edgeColor = WHITE);
    QuadraticSegment *clone() const;
    int type() const;
    const Point2 *controlPoints() const;
    Point2 point(double param) const;
    Vector2 direction(double param) const;
    Vector2 directionChange(double param) const;
    double length() const;
    void bound(double &xMin, double &yMin, double &xMax, double &yMax) const;
    void reverse();
    void moveStartPoint(Point2 to);
    void moveEndPoint(Point2 to);
    void splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const;
    EdgeSegment *convertToCubic() const;
};

class CubicSegment : public EdgeSegment {
public:
    enum EdgeType { EDGE_TYPE = 3 };
    Point2 p[4];
    CubicSegment(Point2 p0, Point2 p1, Point2 p2, Point2 p3, EdgeColor edgeColor = WHITE);
    CubicSegment *clone() const;
    int type() const;
    const Point2 *controlPoints() const;
    Point2 point(double param) const;
    Vector2 direction(double param) const;
    Vector2 directionChange(double param) const;
    void bound(double &xMin, double &yMin, double &xMax, double &yMax) const;
    void reverse();
    void moveStartPoint(Point2 to);
    void moveEndPoint(Point2 to);
    void splitInThirds(EdgeSegment *&part0, EdgeSegment *&part1, EdgeSegment *&part2) const;
};

/// Owning smart pointer to an EdgeSegment (deep-copies on copy).
class EdgeHolder {
public:
    static void swap(EdgeHolder &a, EdgeHolder &b);
    inline EdgeHolder() : edgeSegment(NULL) { }
    inline EdgeHolder(EdgeSegment *segment) : edgeSegment(segment) { }
    inline EdgeHolder(Point2 p0, Point2 p1, EdgeColor edgeColor = WHITE) : edgeSegment(EdgeSegment::create(p0, p1, edgeColor)) { }
    inline EdgeHolder(Point2 p0, Point2 p1, Point2 p2, EdgeColor edgeColor = WHITE) : edgeSegment(EdgeSegment::create(p0, p1, p2, edgeColor)) { }
    inline EdgeHolder(Point2 p0, Point2 p1, Point2 p2, Point2 p3, EdgeColor edgeColor = WHITE) : edgeSegment(EdgeSegment::create(p0, p1, p2, p3, edgeColor)) { }
    EdgeHolder(const EdgeHolder &orig);
    EdgeHolder(EdgeHolder &&orig);
    ~EdgeHolder();
    EdgeHolder &operator=(const EdgeHolder &orig);
    EdgeHolder &operator=(EdgeHolder &&orig);
    inline EdgeSegment &operator*() { return *edgeSegment; }
    inline const EdgeSegment &operator*() const { return *edgeSegment; }
    inline EdgeSegment *operator->() { return edgeSegment; }
    inline const EdgeSegment *operator->() const { return edgeSegment; }
    inline operator EdgeSegment *() { return edgeSegment; }
    inline operator const EdgeSegment *() const { return edgeSegment; }
private:
    EdgeSegment *edgeSegment;
};

/// A closed sequence of edges.
class Contour {
public:
    std::vector<EdgeHolder> edges;
    void addEdge(const EdgeHolder &edge);
    void addEdge(EdgeHolder &&edge);
    EdgeHolder &addEdge();
    void bound(double &xMin, double &yMin, double &xMax, double &yMax) const;
    void boundMiters(double &xMin, double &yMin, double &xMax, double &yMax, double border, double miterLimit, int polarity) const;
    /// -1, 0 or +1: orientation of the contour (sign of the shoelace sum over edge start points).
    int winding() const;
    void reverse();
};

#define MSDFGEN_CORNER_DOT_EPSILON .000001

/// A vector shape: a set of contours.
class Shape {
public:
    struct Bounds { double l, b, r, t; };
    std::vector<Contour> contours;
    bool inverseYAxis;
    Shape();
    void addContour(const Contour &contour);
    void addContour(Contour &&contour);
    Contour &addContour();
    /// Splits single-edge contours in three and pushes apart edges that meet at a (numerically) 180-degree cusp.
    void normalize();
    bool validate() const;
    void bound(double &xMin, double &yMin, double &xMax, double &yMax) const;
    void boundMiters(double &xMin, double &yMin, double &xMax, double &yMax, double border, double miterLimit, int polarity) const;
    Bounds getBounds(double border = 0, double miterLimit = 0, int polarity = 0) const;
    int edgeCount() const;
    YAxisOrientation getYAxisOrientation() const;
    void setYAxisOrientation(YAxisOrientation yAxisOrientation);
};

#define MSDFGEN_EDGE_LENGTH_PRECISION 4

/// Assigns channel colours to edges so that every corner separates two edges with only one channel in common.
void edgeColoringSimple(Shape &shape, double angleThreshold, unsigned long long seed = 0);
/// Ink-trap aware variant of the above.
void edgeColoringInkTrap(Shape &shape, double angleThreshold, unsigned long long seed = 0);

// ---- generator configuration -------------------------------------------------------------------------------------------
struct ErrorCorrectionConfig {
    static const double defaultMinDeviationRatio;
    static const double defaultMinImproveRatio;
    enum Mode { DISABLED, INDISCRIMINATE, EDGE_PRIORITY, EDGE_ONLY } mode;
    enum DistanceCheckMode { DO_NOT_CHECK_DISTANCE, CHECK_DISTANCE_AT_EDGE, ALWAYS_CHECK_DISTANCE } distanceCheckMode;
    double minDeviationRatio;
    double minImproveRatio;
    byte *buffer;
    inline explicit ErrorCorrectionConfig(Mode mode = EDGE_PRIORITY, DistanceCheckMode distanceCheckMode = CHECK_DISTANCE_AT_EDGE, double minDeviationRatio = defaultMinDeviationRatio, double minImproveRatio = defaultMinImproveRatio, byte *buffer = NULL) : mode(mode), distanceCheckMode(distanceCheckMode), minDeviationRatio(minDeviationRatio), minImproveRatio(minImproveRatio), buffer(buffer) { }
};

struct GeneratorConfig {
    bool overlapSupport;
    inline explicit GeneratorConfig(bool overlapSupport = true) : overlapSupport(overlapSupport) { }
};

struct MSDFGeneratorConfig : GeneratorConfig {
    ErrorCorrectionConfig errorCorrection;
    inline MSDFGeneratorConfig() { }
    inline explicit MSDFGeneratorConfig(bool overlapSupport, const ErrorCorrectionConfig &errorCorrection = ErrorCorrectionConfig()) : GeneratorConfig(overlapSupport), errorCorrection(errorCorrection) { }
};

// ---- generators (device) -------------------------------------------------------------------------------------------------
void generateSDF(const BitmapSection<float, 1> &output, const Shape &shape, const SDFTransformation &transformation, const GeneratorConfig &config = GeneratorConfig());
void generatePSDF(const BitmapSection<float, 1> &output, const Shape &shape, const SDFTransformation &transformation, const GeneratorConfig &config = GeneratorConfig());
void generateMSDF(const BitmapSection<float, 3> &output, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void generateMTSDF(const BitmapSection<float, 4> &output, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());

void generateSDF(const BitmapSection<float, 1> &output, const Shape &shape, const Projection &projection, Range range, const GeneratorConfig &config = GeneratorConfig());
void generatePSDF(const BitmapSection<float, 1> &output, const Shape &shape, const Projection &projection, Range range, const GeneratorConfig &config = GeneratorConfig());
void generatePseudoSDF(const BitmapSection<float, 1> &output, const Shape &shape, const Projection &projection, Range range, const GeneratorConfig &config = GeneratorConfig());
void generateMSDF(const BitmapSection<float, 3> &output, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void generateMTSDF(const BitmapSection<float, 4> &output, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());

void generateSDF(const BitmapSection<float, 1> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, bool overlapSupport = true);
void generatePSDF(const BitmapSection<float, 1> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, bool overlapSupport = true);
void generatePseudoSDF(const BitmapSection<float, 1> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, bool overlapSupport = true);
void generateMSDF(const BitmapSection<float, 3> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, const ErrorCorrectionConfig &errorCorrectionConfig = ErrorCorrectionConfig(), bool overlapSupport = true);
void generateMTSDF(const BitmapSection<float, 4> &output, const Shape &shape, Range range, const Vector2 &scale, const Vector2 &translate, const ErrorCorrectionConfig &errorCorrectionConfig = ErrorCorrectionConfig(), bool overlapSupport = true);

// ---- error correction (device) ------------------------------------------------------------------------------------------
void msdfErrorCorrection(const BitmapSection<float, 3> &sdf, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void msdfErrorCorrection(const BitmapSection<float, 4> &sdf, const Shape &shape, const SDFTransformation &transformation, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void msdfErrorCorrection(const BitmapSection<float, 3> &sdf, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void msdfErrorCorrection(const BitmapSection<float, 4> &sdf, const Shape &shape, const Projection &projection, Range range, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());

void msdfFastDistanceErrorCorrection(const BitmapSection<float, 3> &sdf, const SDFTransformation &transformation, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 4> &sdf, const SDFTransformation &transformation, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 3> &sdf, const Projection &projection, Range range, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 4> &sdf, const Projection &projection, Range range, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 3> &sdf, Range pxRange, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastDistanceErrorCorrection(const BitmapSection<float, 4> &sdf, Range pxRange, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);

void msdfFastEdgeErrorCorrection(const BitmapSection<float, 3> &sdf, const SDFTransformation &transformation, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 4> &sdf, const SDFTransformation &transformation, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 3> &sdf, const Projection &projection, Range range, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 4> &sdf, const Projection &projection, Range range, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 3> &sdf, Range pxRange, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);
void msdfFastEdgeErrorCorrection(const BitmapSection<float, 4> &sdf, Range pxRange, double minDeviationRatio = ErrorCorrectionConfig::defaultMinDeviationRatio);

// ---- scanline pass (core/Scanline.h:9-14, core/rasterization.h) -----------------------------------------------------

/// 1 inside, 0 outside, per texel centre (scanline pass on the device).
void rasterize(BitmapSection<float, 1> output, const Shape &shape, const Projection &projection, FillRule fillRule = FILL_NONZERO);
/// Flips texels whose side of sdfZeroValue disagrees with the scanline fill of the shape (scanline pass on the device).
void distanceSignCorrection(BitmapSection<float, 1> sdf, const Shape &shape, const Projection &projection, float sdfZeroValue = .5f, FillRule fillRule = FILL_NONZERO);
void distanceSignCorrection(BitmapSection<float, 3> sdf, const Shape &shape, const Projection &projection, float sdfZeroValue = .5f, FillRule fillRule = FILL_NONZERO);
void distanceSignCorrection(BitmapSection<float, 4> sdf, const Shape &shape, const Projection &projection, float sdfZeroValue = .5f, FillRule fillRule = FILL_NONZERO);
// overloads of older msdfgen releases
void rasterize(const BitmapSection<float, 1> &output, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule = FILL_NONZERO);
void distanceSignCorrection(BitmapSection<float, 1> sdf, const Shape &shape, const Projection &projection, FillRule fillRule);
void distanceSignCorrection(BitmapSection<float, 3> sdf, const Shape &shape, const Projection &projection, FillRule fillRule);
void distanceSignCorrection(BitmapSection<float, 4> sdf, const Shape &shape, const Projection &projection, FillRule fillRule);
void distanceSignCorrection(const BitmapSection<float, 1> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule = FILL_NONZERO);
void distanceSignCorrection(const BitmapSection<float, 3> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule = FILL_NONZERO);
void distanceSignCorrection(const BitmapSection<float, 4> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, FillRule fillRule = FILL_NONZERO);

// ---- preview rendering (core/render-sdf.h) --------------------------------------------------------------------------------

/// Samples the distance field at the output's resolution and thresholds it (optionally with a soft edge of sdfPxRange).
void renderSDF(const BitmapSection<float, 1> &output, const BitmapConstSection<float, 1> &sdf, Range sdfPxRange = 0, float sdThreshold = .5f);
void renderSDF(const BitmapSection<float, 3> &output, const BitmapConstSection<float, 1> &sdf, Range sdfPxRange = 0, float sdThreshold = .5f);
void renderSDF(const BitmapSection<float, 1> &output, const BitmapConstSection<float, 3> &sdf, Range sdfPxRange = 0, float sdThreshold = .5f);
void renderSDF(const BitmapSection<float, 3> &output, const BitmapConstSection<float, 3> &sdf, Range sdfPxRange = 0, float sdThreshold = .5f);
void renderSDF(const BitmapSection<float, 1> &output, const BitmapConstSection<float, 4> &sdf, Range sdfPxRange = 0, float sdThreshold = .5f);
void renderSDF(const BitmapSection<float, 4> &output, const BitmapConstSection<float, 4> &sdf, Range sdfPxRange = 0, float sdThreshold = .5f);
/// Rounds every value to what it would be after a trip through an 8-bit texture.
void simulate8bit(const BitmapSection<float, 1> &bitmap);
void simulate8bit(const BitmapSection<float, 3> &bitmap);
void simulate8bit(const BitmapSection<float, 4> &bitmap);

// ---- error estimation (core/sdf-error-estimation.h:19-26): evaluated on the device (sdf_error.cuh) ---------------------------------

/// Fraction of the area whose fill, reconstructed from the SDF, differs from the shape's (scanlinesPerRow scanlines per texel row).
double estimateSDFError(const BitmapConstSection<float, 1> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule = FILL_NONZERO);
double estimateSDFError(const BitmapConstSection<float, 3> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule = FILL_NONZERO);
double estimateSDFError(const BitmapConstSection<float, 4> &sdf, const Shape &shape, const Projection &projection, int scanlinesPerRow, FillRule fillRule = FILL_NONZERO);
// overloads of older msdfgen releases
double estimateSDFError(const BitmapConstSection<float, 1> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule = FILL_NONZERO);
double estimateSDFError(const BitmapConstSection<float, 3> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule = FILL_NONZERO);
double estimateSDFError(const BitmapConstSection<float, 4> &sdf, const Shape &shape, const Vector2 &scale, const Vector2 &translate, int scanlinesPerRow, FillRule fillRule = FILL_NONZERO);

// ---- B200 additions: atlas-sized batches and device control ------------------------------------------------------------------
namespace b200 {

/// Selects the CUDA device used by the calling thread's generator context (default 0).
void setDevice(int device);

/// Thrown-free error reporting: the msdfgen API is void, so failures of the device path abort with a message unless a
/// handler is installed. Returns the previous handler.
typedef void (*ErrorHandler)(int status, const char *message);
ErrorHandler setErrorHandler(ErrorHandler handler);

/// One generate* call per shape, executed as ONE device batch (shapes.size() == outputs.size()).
void generateMSDFBatch(const std::vector<BitmapSection<float, 3> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void generateMTSDFBatch(const std::vector<BitmapSection<float, 4> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void generateSDFBatch(const std::vector<BitmapSection<float, 1> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const GeneratorConfig &config = GeneratorConfig());
void generatePSDFBatch(const std::vector<BitmapSection<float, 1> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const GeneratorConfig &config = GeneratorConfig());

/// The command line's -scanline flow for a whole batch in ONE device pass: generation with error correction postponed,
/// distanceSignCorrection, then msdfErrorCorrection (main.cpp:1223-1297).  Applies to the Batch calls of the calling thread
/// made while enabled.
void setScanlinePass(bool enabled, FillRule fillRule = FILL_NONZERO, float sdfZeroValue = .5f);

/// 8-bit atlas output: generation + fused pixelFloatToByte + placement into byte sections (e.g. rectangles of one atlas page).
void generateSDFBatch(const std::vector<BitmapSection<byte, 1> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const GeneratorConfig &config = GeneratorConfig());
void generateMSDFBatch(const std::vector<BitmapSection<byte, 3> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());
void generateMTSDFBatch(const std::vector<BitmapSection<byte, 4> > &outputs, const std::vector<const Shape *> &shapes, const std::vector<SDFTransformation> &transformations, const MSDFGeneratorConfig &config = MSDFGeneratorConfig());

}  // namespace b200

}  // namespace msdfgen
