"""Test helper: writes minimal TrueType fonts (head, hhea, maxp, hmtx, loca, glyf, cmap, post, name, optionally kern) from glyph
descriptions, so that the ingest's less common paths — contours that start on a control point, contours made of control points only,
composite glyphs with offsets / scales / 2x2 matrices / nested components, cmap formats 4 and 12, kern format 0 — can be exercised
and compared with FreeType (the one bundled in Pillow) on the same bytes."""
import struct

ON, OFF = 1, 0


def simple(contours):
    """contours: list of lists of (x, y, on) -> glyf entry bytes"""
    pts = [p for c in contours for p in c]
    if not pts:
        return b""
    ends, n = [], 0
    for c in contours:
        n += len(c)
        ends.append(n - 1)
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    out = struct.pack(">hhhhh", len(contours), min(xs), min(ys), max(xs), max(ys))
    out += b"".join(struct.pack(">H", e) for e in ends) + struct.pack(">H", 0)
    out += bytes(1 if p[2] else 0 for p in pts)
    for axis in (0, 1):
        prev = 0
        for p in pts:
            out += struct.pack(">h", p[axis] - prev)
            prev = p[axis]
    return out


def composite(components, bbox=(0, 0, 1000, 1000)):
    """components: list of dicts {glyph, dx, dy, [scale] | [xscale, yscale] | [matrix=(xx, yx, xy, yy)], [scaled_offset], [match=(parent, child)]}"""
    out = struct.pack(">hhhhh", -1, *bbox)
    for i, c in enumerate(components):
        flags = 0x0001                                  # ARG_1_AND_2_ARE_WORDS
        if "match" not in c:
            flags |= 0x0002                             # ARGS_ARE_XY_VALUES
        if i + 1 < len(components):
            flags |= 0x0020
        if c.get("scaled_offset"):
            flags |= 0x0800
        tr = b""
        f2 = lambda v: struct.pack(">h", int(round(v * 16384)))
        if "scale" in c:
            flags |= 0x0008
            tr = f2(c["scale"])
        elif "xscale" in c:
            flags |= 0x0040
            tr = f2(c["xscale"]) + f2(c["yscale"])
        elif "matrix" in c:
            flags |= 0x0080
            tr = b"".join(f2(v) for v in c["matrix"])
        a1, a2 = c["match"] if "match" in c else (c.get("dx", 0), c.get("dy", 0))
        out += struct.pack(">HH", flags, c["glyph"]) + (struct.pack(">HH", a1, a2) if "match" in c else struct.pack(">hh", a1, a2)) + tr
    return out


def build(glyphs, cmap, units_per_em=1000, advances=None, kern=None, cmap12=False, ascender=800, descender=-200, lsb=None):
    """glyphs: list of glyf entry bytes (index 0 = .notdef); cmap: {code point: glyph index}; kern: {(left, right): value};
    lsb: left side bearings (default: each glyph's header xMin, as in a well-formed font)"""
    n = len(glyphs)
    advances = advances or [units_per_em] * n
    lsb = lsb or [struct.unpack(">h", g[2:4])[0] if g else 0 for g in glyphs]
    loca, glyf = [0], b""
    for g in glyphs:
        glyf += g + b"\0" * (-len(g) % 4)
        loca.append(len(glyf))
    tables = {
        "head": struct.pack(">IIIIHHqqhhhhHHhhh", 0x00010000, 0x00010000, 0, 0x5F0F3CF5, 0x000B, units_per_em, 0, 0, 0, descender, units_per_em,
                            ascender, 0, 8, 2, 1, 0),
        "hhea": struct.pack(">IhhhHhhhhhhhhhhhH", 0x00010000, ascender, descender, 0, max(advances), 0, 0, units_per_em, 1, 0, 0, 0, 0, 0, 0, 0, n),
        "maxp": struct.pack(">IHHHHHHHHHHHHHH", 0x00010000, n, 256, 32, 512, 64, 2, 0, 0, 0, 0, 64, 0, 8, 4),
        "hmtx": b"".join(struct.pack(">Hh", a, b) for a, b in zip(advances, lsb)),
        "loca": b"".join(struct.pack(">I", o) for o in loca),
        "glyf": glyf if glyf else b"\0\0\0\0",
        "post": struct.pack(">IIhhIIIII", 0x00030000, 0, -100, 50, 0, 0, 0, 0, 0),
        "name": struct.pack(">HHH", 0, 0, 6),
    }
    # cmap: format 4 (BMP) and optionally format 12
    bmp = sorted((c, g) for c, g in cmap.items() if c < 0xffff)
    segs = [(c, c, (g - c) & 0xffff) for c, g in bmp] + [(0xffff, 0xffff, 1)]
    sc = len(segs)
    sr = 2 * (1 << (sc.bit_length() - 1))
    f4 = struct.pack(">HHHHHHH", 4, 16 + 8 * sc, 0, 2 * sc, sr, sc.bit_length() - 1, 2 * sc - sr)
    f4 += b"".join(struct.pack(">H", s[1]) for s in segs) + struct.pack(">H", 0)
    f4 += b"".join(struct.pack(">H", s[0]) for s in segs) + b"".join(struct.pack(">H", s[2]) for s in segs) + struct.pack(">H", 0) * sc
    subtables = [(3, 1, f4)]
    if cmap12:
        groups = sorted(cmap.items())
        f12 = struct.pack(">HHIII", 12, 0, 16 + 12 * len(groups), 0, len(groups)) + b"".join(struct.pack(">III", c, c, g) for c, g in groups)
        subtables.append((3, 10, f12))
    head = struct.pack(">HH", 0, len(subtables))
    off = 4 + 8 * len(subtables)
    body = b""
    for plat, enc, data in subtables:
        head += struct.pack(">HHI", plat, enc, off + len(body))
        body += data
    tables["cmap"] = head + body
    if kern:
        pairs = sorted(kern.items())
        sub = struct.pack(">HHHHHHH", 0, 14 + 6 * len(pairs), 0x0001, len(pairs), 0, 0, 0) + b"".join(struct.pack(">HHh", l, r, v) for (l, r), v in pairs)
        tables["kern"] = struct.pack(">HH", 0, 1) + sub
    tags = sorted(tables)
    nt = len(tags)
    sr = 16 * (1 << (nt.bit_length() - 1))
    out = struct.pack(">IHHHH", 0x00010000, nt, sr, nt.bit_length() - 1, 16 * nt - sr)
    offset = 12 + 16 * nt
    blobs = b""
    for t in tags:
        d = tables[t]
        pad = d + b"\0" * (-len(d) % 4)
        checksum = sum(struct.unpack(">%dI" % (len(pad) // 4), pad)) & 0xffffffff
        out += struct.pack(">4sIII", t.encode(), checksum, offset + len(blobs), len(d))
        blobs += pad
    return out + blobs
