"""Shared helpers for the parity tests."""
import numpy as np


def compare(a, b):
    """Parity metrics between two float32 arrays: (#differing bit patterns, max |delta|, #texels over 1e-4)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    diff_bits = int((a.view(np.uint32) != b.view(np.uint32)).sum())
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    d = np.where(both_inf, 0.0, np.abs(a.astype(np.float64) - b.astype(np.float64)))
    d = np.nan_to_num(d, nan=np.inf)
    return diff_bits, float(d.max()) if d.size else 0.0, int((d > 1e-4).sum())


def to_bytes(a):
    """pixelFloatToByte (/root/reference/core/pixel-conversion.hpp:8-10) in numpy."""
    a = np.asarray(a, dtype=np.float32)
    c = np.where((a >= 0) & (a <= 1), a, (a > 0).astype(np.float32)).astype(np.float32)
    v = (np.float32(255.5) - np.float32(255.0) * c).astype(np.float32)
    return (~v.astype(np.int32)).astype(np.uint8)


def fuzz_shape(seed):
    """Adversarial closed shapes for parity fuzzing: (family, cbegin, points[n,8], orders, colors, xf[6], w, h).

    Families (seed % 9): 0 integer-grid polygons with texel centres ON the vertices/edges (ties everywhere); 1 quadratics with
    collinear control points inside the chord (straight-line curves); 2 cubics with coincident control points, cusps and loops; 3 contours with
    zero-length edges and repeated points; 4 huge coordinates (1e6) ; 5 tiny coordinates (1e-6); 6 one- and two-edge
    contours and an empty contour; 7 self-intersecting star polygons with mixed curve types; 8 quadratic CUSPS: the control point lies on the
    chord line outside the chord (hairpin: the curve doubles back) or coincides with an end point, so the tangent vanishes at a
    parameter t* in [0,1].  t* is then a root of the foot-point cubic, and what the reference returns for texels whose nearest
    point is the cusp is decided by the last bit of its libm's acos/cos/pow: whether the computed root falls inside (0,1) when
    t* is 0 or 1, and nonZeroSign(cross(~0, q)) when it is a hairpin tip.  No implementation with a different libm (another
    compiler, another platform, a GPU) reproduces that noise, so only magnitudes are compared for this family.
    Colours are random non-black masks (black edges take part in no channel; one family-3 edge is black on purpose)."""
    rng = np.random.default_rng(1000003 * seed + 17)
    fam = seed % 9
    contours = []       # list of lists of (order, [points...])

    def poly_to_edges(vs, kinds=None):
        edges = []
        n = len(vs)
        for i in range(n):
            a, b = vs[i], vs[(i + 1) % n]
            k = 1 if kinds is None else kinds[i]
            if k == 1:
                edges.append((1, [a, b]))
            elif k == 2:
                c = (a + b) / 2 + rng.normal(0, 0.25 * np.linalg.norm(b - a) + 1e-12, 2)
                edges.append((2, [a, c, b]))
            else:
                d = b - a
                c1 = a + d / 3 + rng.normal(0, 0.3 * np.linalg.norm(d) + 1e-12, 2)
                c2 = a + 2 * d / 3 + rng.normal(0, 0.3 * np.linalg.norm(d) + 1e-12, 2)
                edges.append((3, [a, c1, c2, b]))
        return edges

    w = h = 16
    xf = np.array([1.0, 1.0, 0.5, 0.5, -2.0, 2.0])          # texel centres = integer coordinates
    if fam == 0:
        for _ in range(int(rng.integers(1, 4))):
            n = int(rng.integers(3, 8))
            vs = rng.integers(1, 15, (n, 2)).astype(np.float64)
            contours.append(poly_to_edges(vs))
    elif fam == 1 or fam == 8:
        vs = rng.integers(1, 15, (5, 2)).astype(np.float64)
        edges = []
        for i in range(5):
            a, b = vs[i], vs[(i + 1) % 5]
            t = rng.choice([0.5, 0.25, 0.75, 0.5]) if fam == 1 else rng.choice([-0.5, 1.5, 0.0, 1.0, -0.25])
            edges.append((2, [a, a + t * (b - a), b]))       # control point on the chord line (or on an end point)
        contours.append(edges)
    elif fam == 2:
        vs = rng.integers(1, 15, (4, 2)).astype(np.float64)
        edges = []
        for i in range(4):
            a, b = vs[i], vs[(i + 1) % 4]
            kind = int(rng.integers(0, 5))
            if kind == 0:
                edges.append((3, [a, a, b, b]))
            elif kind == 1:
                edges.append((3, [a, a, a + 0.5 * (b - a) + rng.normal(0, 2, 2), b]))
            elif kind == 2:
                edges.append((3, [a, b + rng.normal(0, 3, 2), a + rng.normal(0, 3, 2), b]))      # loop / cusp candidates
            elif kind == 3:
                m = (a + b) / 2
                edges.append((3, [a, m, m, b]))
            else:
                edges.append((3, [a, a + (b - a) / 3, a + 2 * (b - a) / 3, b]))                # a straight cubic
        contours.append(edges)
    elif fam == 3:
        vs = rng.integers(2, 14, (5, 2)).astype(np.float64)
        edges = poly_to_edges(vs, kinds=rng.integers(1, 4, 5))
        k = int(rng.integers(0, 5))
        edges.insert(k, (1, [edges[k][1][0], edges[k][1][0]]))           # zero-length line
        edges.insert(0, (2, [edges[0][1][0]] * 3))                       # zero-length quadratic
        contours.append(edges)
    elif fam == 4 or fam == 5:
        s = 1e6 if fam == 4 else 1e-6
        vs = rng.uniform(1, 15, (6, 2))
        order = np.argsort(np.arctan2(vs[:, 1] - 8, vs[:, 0] - 8))
        vs = vs[order] * s + (3e6 if fam == 4 else 0.0)
        contours.append(poly_to_edges(vs, kinds=rng.integers(1, 4, 6)))
        off = 3e6 if fam == 4 else 0.0
        xf = np.array([1.0 / s, 1.0 / s, -off + 0.0, -off + 0.0, -2.0 * s, 2.0 * s])
    elif fam == 6:
        a, b = rng.uniform(2, 14, 2), rng.uniform(2, 14, 2)
        contours.append([(3, [a, a + rng.normal(0, 4, 2), b + rng.normal(0, 4, 2), a])])           # one closed cubic
        contours.append([])                                                                        # empty contour
        c, d = rng.uniform(2, 14, 2), rng.uniform(2, 14, 2)
        contours.append([(2, [c, (c + d) / 2 + rng.normal(0, 3, 2), d]), (1, [d, c])])
        xf = np.array([1.0, 1.0, 0.0, 0.0, -2.0, 2.0])
    else:
        n = 7
        ang = 2 * np.pi * (np.arange(n) * 3 % n) / n
        vs = 8 + 6.5 * np.stack([np.cos(ang), np.sin(ang)], axis=1)
        contours.append(poly_to_edges(vs, kinds=rng.integers(1, 4, n)))
        xf = np.array([1.0, 1.0, 0.0, 0.0, -3.0, 3.0])
    cbegin, pts, orders = [0], [], []
    for edges in contours:
        for order, ps in edges:
            row = np.zeros(8)
            row[: 2 * (order + 1)] = np.concatenate(ps)
            pts.append(row)
            orders.append(order)
        cbegin.append(len(orders))
    n = len(orders)
    colors = rng.choice([1, 2, 3, 4, 5, 6, 7], n).astype(np.uint8)
    if fam == 3 and n:
        colors[int(rng.integers(0, n))] = 0
    return (fam, np.array(cbegin, np.uint32), np.array(pts, np.float64).reshape(n, 8), np.array(orders, np.uint8), colors, xf, w, h)


def fuzz_configs(seed):
    """(key, mode name, oracle Config) triples exercised per fuzz seed."""
    from oracle.bindings import Config
    out = []
    for mode in ("sdf", "psdf", "msdf", "mtsdf"):
        for ov in (True, False):
            out.append((f"{mode}/ov{int(ov)}/ec0", mode, Config(overlap=ov, ec_mode=0)))
            if mode in ("msdf", "mtsdf"):
                out.append((f"{mode}/ov{int(ov)}/ec{1 + seed % 3}{seed % 3}", mode, Config(overlap=ov, ec_mode=1 + seed % 3, ec_dcm=seed % 3)))
    return out


def digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).hexdigest()[:24]
