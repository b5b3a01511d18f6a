// PNG writer (8-bit RGB) — replaces the reference's vendored stb_image_write for saveImage (reference
// src/Render.cpp:257-265): adaptive row filters + zlib's deflate.  Not on the hot path.
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <vector>

#include <zlib.h>

namespace cornelis {
namespace {

std::uint32_t crc32(std::uint32_t crc, unsigned char const *data, std::size_t n) {
    static std::uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (std::uint32_t i = 0; i < 256; i++) {
            std::uint32_t c = i;
            for (int k = 0; k < 8; k++)
                c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (std::size_t i = 0; i < n; i++)
        crc = table[(crc ^ data[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}

void put32(std::vector<unsigned char> &out, std::uint32_t v) {
    for (int s = 24; s >= 0; s -= 8)
        out.push_back(static_cast<unsigned char>(v >> s));
}

void chunk(std::vector<unsigned char> &out, char const tag[4], std::vector<unsigned char> const &payload) {
    put32(out, static_cast<std::uint32_t>(payload.size()));
    std::size_t const start = out.size();
    out.insert(out.end(), tag, tag + 4);
    out.insert(out.end(), payload.begin(), payload.end());
    put32(out, crc32(0, out.data() + start, out.size() - start));
}

} // namespace

// Rows are filtered (PNG filter types 0-4, the one with the smallest sum of absolute residuals per row, as stb's
// encoder and libpng's heuristic do) and deflated by zlib.
bool writePngRgb8(std::string const &path, int width, int height, unsigned char const *rgb) {
    std::size_t const stride = 3u * static_cast<std::size_t>(width);
    std::vector<unsigned char> raw;
    raw.reserve(static_cast<std::size_t>(height) * (stride + 1u));
    std::vector<unsigned char> candidate(stride), best(stride), zeros(stride, 0);
    auto paeth = [](int a, int b, int c) {
        int const p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
        return pa <= pb && pa <= pc ? a : pb <= pc ? b : c;
    };
    for (int j = 0; j < height; j++) {
        unsigned char const *row = rgb + static_cast<std::size_t>(j) * stride;
        unsigned char const *up = j ? row - stride : zeros.data();
        unsigned long bestCost = ~0ul;
        int bestType = 0;
        for (int type = 0; type < 5; type++) {
            unsigned long cost = 0;
            for (std::size_t i = 0; i < stride; i++) {
                int const left = i >= 3 ? row[i - 3] : 0, above = up[i], diag = i >= 3 ? up[i - 3] : 0;
                int predicted = 0;
                switch (type) {
                case 1: predicted = left; break;
                case 2: predicted = above; break;
                case 3: predicted = (left + above) / 2; break;
                case 4: predicted = paeth(left, above, diag); break;
                default: break;
                }
                unsigned char const r = static_cast<unsigned char>(row[i] - predicted);
                candidate[i] = r;
                cost += static_cast<unsigned long>(r < 128 ? r : 256 - r);
            }
            if (cost < bestCost) {
                bestCost = cost;
                bestType = type;
                best.swap(candidate);
            }
        }
        raw.push_back(static_cast<unsigned char>(bestType));
        raw.insert(raw.end(), best.begin(), best.end());
    }
    uLongf zsize = compressBound(static_cast<uLong>(raw.size()));
    std::vector<unsigned char> z(zsize);
    if (compress2(z.data(), &zsize, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK)
        return false;
    z.resize(zsize);

    std::vector<unsigned char> file{0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<unsigned char> header;
    put32(header, static_cast<std::uint32_t>(width));
    put32(header, static_cast<std::uint32_t>(height));
    header.insert(header.end(), {8, 2, 0, 0, 0}); // 8 bits, colour type RGB
    chunk(file, "IHDR", header);
    chunk(file, "IDAT", z);
    chunk(file, "IEND", {});
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f)
        return false;
    bool const ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
    std::fclose(f);
    return ok;
}

} // namespace cornelis
