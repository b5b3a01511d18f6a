// Minimal PNG writer (8-bit RGB, stored deflate blocks) — replaces the reference's vendored stb_image_write for
// saveImage (reference src/Render.cpp:257-265).  No compression: the encoder is not on the hot path.
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

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

bool writePngRgb8(std::string const &path, int width, int height, unsigned char const *rgb) {
    std::vector<unsigned char> raw;
    raw.reserve(static_cast<std::size_t>(height) * (3u * width + 1u));
    for (int j = 0; j < height; j++) {
        raw.push_back(0); // filter: none
        raw.insert(raw.end(), rgb + static_cast<std::size_t>(j) * 3u * width, rgb + static_cast<std::size_t>(j + 1) * 3u * width);
    }
    std::vector<unsigned char> z{0x78, 0x01};
    std::uint32_t a = 1, b = 0;
    for (std::size_t pos = 0; pos < raw.size();) {
        std::size_t const n = std::min<std::size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back(static_cast<unsigned char>(n & 0xFF));
        z.push_back(static_cast<unsigned char>(n >> 8));
        z.push_back(static_cast<unsigned char>(~n & 0xFF));
        z.push_back(static_cast<unsigned char>((~n >> 8) & 0xFF));
        for (std::size_t i = 0; i < n; i++) {
            a = (a + raw[pos + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        z.insert(z.end(), raw.begin() + static_cast<std::ptrdiff_t>(pos), raw.begin() + static_cast<std::ptrdiff_t>(pos + n));
        pos += n;
    }
    put32(z, (b << 16) | a);

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
