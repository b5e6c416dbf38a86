// Decoded RGBA8 image of a scene texture. Stands in for the reference's STBImage (reference src/STBImage.hpp, a wrapper around
// the vendored stb_image, which always expands to 4 channels: src/STBImage.cpp). Decoders written here: PNG (grey / grey+alpha / RGB /
// RGBA / palette at every legal bit depth incl. tRNS keys, non-interlaced; 16-bit samples keep their high byte; inflate through zlib), Netpbm P6 (RGB) and P7 (RGB_ALPHA). JPEG and the other stb_image
// formats are not decoded: load() fails and Scene::loadScene substitutes the blank image the reference uses for textures it cannot read.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace vkx {

struct Image {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> pixels; // width * height * 4, row-major
    bool load(const std::string& path, std::string* error = nullptr);
    bool savePam(const std::string& path) const;
    static Image blank(); // 1 x 1 opaque white
};

} // namespace vkx
