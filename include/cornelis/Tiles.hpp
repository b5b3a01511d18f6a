// Frame tiling (reference include/cornelis/Tiles.hpp, src/Tiles.cpp).  On the CPU a tile is the unit of scheduling
// and owns a PRNG; on the GPU path samples, not tiles, are the unit of work and random numbers are keyed by
// (pixel, sample, depth), so TileInfo only carries its number and bounds.  FrameTiling is kept for callers that
// partition frames (progress display, region renders).
//
// Unlike the reference (Tiles.cpp:21-24, where the last column / row of a frame that is not a multiple of the tile
// gets max = spill - 1 measured from 0 and overlaps the first tiles), partial tiles here end at the frame edge:
// every pixel belongs to exactly one tile.
#pragma once

#include <cstddef>
#include <vector>

#include <cornelis/Math.hpp>

namespace cornelis {

using TileCoord = PixelCoord;

struct TileInfo {
    explicit TileInfo(std::size_t number, PixelRect pBounds) : tileNumber(number), bounds(pBounds) {}
    std::size_t tileNumber; // unique identifier, left-to-right then top-to-bottom
    PixelRect bounds;       // pixels that belong to this tile (inclusive)
};

class FrameTiling {
  public:
    using container_type = std::vector<TileInfo>;
    using iterator = container_type::iterator;
    using const_iterator = container_type::const_iterator;
    using value_type = container_type::value_type;
    using reference = container_type::reference;

    explicit FrameTiling(PixelRect dimensions, PixelRect maxTileSize = PixelRect{32, 32});

    const_iterator begin() const noexcept { return tiles_.begin(); }
    const_iterator end() const noexcept { return tiles_.end(); }
    iterator begin() noexcept { return tiles_.begin(); }
    iterator end() noexcept { return tiles_.end(); }
    std::size_t size() const noexcept { return tiles_.size(); }
    TileInfo const &operator[](std::size_t k) const { return tiles_[k]; }

  private:
    container_type tiles_;
};

} // namespace cornelis
