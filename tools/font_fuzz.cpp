// font_fuzz.cpp — robustness check of the TrueType reader (msdfgen_b200/host/font_ingest.cpp): 3000 mutations of a font file
// (random bytes, truncation, a scrambled table directory, a scrambled 64-byte window) through open / info / all glyphs / cmap /
// kerning / loadGlyph under AddressSanitizer + UBSan.  Every table read is bounds-checked; this is the test of that claim.
//   H=msdfgen_b200/host; g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -std=c++11 -I$H -o /tmp/font_fuzz \
//       tools/font_fuzz.cpp $H/font_ingest.cpp $H/shape_model.cpp $H/flat_prep.cpp && /tmp/font_fuzz some.ttf
// Last run (fontawesome-webfont.ttf from the CUDA toolkit's documentation): 1583 of 3000 mutants opened, 60.5 M segments
// decomposed, no report.
#include "import-font.h"
#include <cstdio>
#include <vector>
#include <cstdint>
using namespace msdfgen;
extern "C" { void *b200host_font_open(const uint8_t*, int); void b200host_font_close(void*); struct b200host_glyphs; b200host_glyphs *b200host_font_glyphs_begin(void*, int, const uint32_t*, int, int*, int*); void b200host_font_glyphs_free(b200host_glyphs*); int b200host_font_info(void*, int*); int b200host_font_glyph_indices(void*, int, const uint32_t*, uint32_t*); int b200host_font_kerning(void*, uint32_t, uint32_t, int*); }
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb"); std::vector<uint8_t> data(200000); data.resize(fread(data.data(), 1, data.size(), f)); fclose(f);
    uint64_t s = 88172645463325252ull; auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    int opened = 0; long glyphs = 0;
    for (int it = 0; it < 3000; ++it) {
        std::vector<uint8_t> d = data;
        switch (it%4) {
            case 0: for (int k = rnd()%40+1; k--;) d[rnd()%d.size()] = (uint8_t) rnd(); break;
            case 1: d.resize(12+rnd()%(d.size()-12)); break;
            case 2: for (int k = 200; k--;) d[rnd()%4096] = (uint8_t) rnd(); break;
            default: { size_t a = rnd()%(d.size()-64); for (int k = 0; k < 64; ++k) d[a+k] = (uint8_t) rnd(); }
        }
        void *font = b200host_font_open(d.data(), (int) d.size());
        if (!font) continue;
        ++opened;
        int info[8]; b200host_font_info(font, info);
        int n = info[0] < 3000 ? info[0] : 3000;
        std::vector<uint32_t> idx(n); for (int i = 0; i < n; ++i) idx[i] = i;
        int nc, ns; b200host_glyphs *h = b200host_font_glyphs_begin(font, n, idx.data(), 1, &nc, &ns); glyphs += ns; b200host_font_glyphs_free(h);
        std::vector<uint32_t> u(512), g(512); for (int i = 0; i < 512; ++i) u[i] = 0xf000+i*37; b200host_font_glyph_indices(font, 512, u.data(), g.data());
        int kv; b200host_font_kerning(font, 3, 4, &kv);
        Shape shape; double adv; loadGlyph(shape, (FontHandle *) font, GlyphIndex(rnd()%800), FONT_SCALING_EM_NORMALIZED, &adv);
        b200host_font_close(font);
    }
    printf("opened %d segments %ld\n", opened, glyphs);
}
