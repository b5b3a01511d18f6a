// FrameTiling (reference src/Tiles.cpp) with partial tiles clipped at the frame edge.
#include <cornelis/Tiles.hpp>

namespace cornelis {

FrameTiling::FrameTiling(PixelRect dimensions, PixelRect maxTileSize) {
    auto const W = dimensions.width(), H = dimensions.height();
    auto const tw = maxTileSize.width(), th = maxTileSize.height();
    auto const columns = (W + tw - 1) / tw, rows = (H + th - 1) / th;
    tiles_.reserve(static_cast<std::size_t>(columns) * static_cast<std::size_t>(rows));
    std::size_t number = 0;
    for (PixelRect::element_type r = 0; r < rows; r++)
        for (PixelRect::element_type c = 0; c < columns; c++) {
            PixelCoord const lo{c * tw, r * th};
            PixelCoord const hi{std::min((c + 1) * tw, W) - 1, std::min((r + 1) * th, H) - 1};
            tiles_.emplace_back(number++, PixelRect{lo, hi});
        }
}

} // namespace cornelis
