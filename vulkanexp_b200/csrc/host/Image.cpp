#include "Image.hpp"
#include "../../../include/vkx.h"
#include <zlib.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <new>
#include <stdexcept>
#include <sstream>

namespace vkx {

namespace {

bool fail(std::string* error, const std::string& what) { if (error) *error = what; return false; }

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }

// PNG (ISO/IEC 15948): signature, IHDR, optional PLTE / tRNS, concatenated IDAT = zlib stream of filtered scanlines.
bool decodePng(const std::vector<uint8_t>& f, Image& img, std::string* error) {
    size_t off = 8;
    uint32_t w = 0, h = 0; int depth = 0, colour = -1, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    while (off + 12 <= f.size()) {
        const uint32_t len = be32(&f[off]);
        const char* type = reinterpret_cast<const char*>(&f[off + 4]);
        if (off + 12 + size_t(len) > f.size()) return fail(error, "truncated PNG chunk");
        const uint8_t* data = &f[off + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) { w = be32(data); h = be32(data + 4); depth = data[8]; colour = data[9]; interlace = data[12]; }
        else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        off += 12 + size_t(len);
    }
    if (w == 0 || h == 0 || interlace != 0) return fail(error, "unsupported PNG (interlaced or empty)");
    if (w > 16384u || h > 16384u) return fail(error, "PNG larger than 16384 x 16384 (VKX_MAX_TEXTURE_SIZE)"); // IHDR values are attacker-controlled: no multi-GB allocations
    int channels;
    switch (colour) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 4: channels = 2; break; case 6: channels = 4; break; default: return fail(error, "unsupported PNG colour type"); }
    const bool depthOk = depth == 8 || depth == 16 || ((colour == 0 || colour == 3) && (depth == 1 || depth == 2 || depth == 4));
    if (!depthOk || (colour == 3 && depth == 16)) return fail(error, "unsupported PNG bit depth");
    const size_t stride = (size_t(w) * size_t(channels) * size_t(depth) + 7) / 8; // bytes per scanline
    const size_t bpp = std::max<size_t>(1, size_t(channels) * size_t(depth) / 8); // filter distance
    if (idat.size() < 2 || (stride + 1) * size_t(h) > idat.size() * 1040 + 64) return fail(error, "PNG IDAT too small for the image size"); // deflate expands at most ~1032x
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawLen = uLongf(raw.size());
    if (uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size())) != Z_OK || rawLen != raw.size()) return fail(error, "PNG inflate failed");
    std::vector<uint8_t> cur(stride), prev(stride, 0);
    std::vector<uint16_t> samples(size_t(w) * channels);
    img.width = w; img.height = h; img.pixels.assign(size_t(w) * h * 4, 255);
    // tRNS for grey / RGB images: one colour that becomes transparent (compared at the file's bit depth)
    const bool keyed = (colour == 0 && trns.size() >= 2) || (colour == 2 && trns.size() >= 6);
    uint16_t key[3] = {0, 0, 0};
    if (keyed) for (int k = 0; k < (colour == 0 ? 1 : 3); ++k) key[k] = uint16_t((trns[2 * k] << 8) | trns[2 * k + 1]);
    const int scale = depth == 1 ? 255 : depth == 2 ? 85 : depth == 4 ? 17 : 1; // grey samples below 8 bits are stretched to 0..255
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* line = &raw[(stride + 1) * y];
        const int filter = line[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred = 0;
            switch (filter) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) / 2; break;
                case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: return fail(error, "bad PNG filter type");
            }
            cur[i] = uint8_t(line[1 + i] + pred);
        }
        for (size_t k = 0; k < samples.size(); ++k) { // samples at the file's depth
            if (depth == 8) samples[k] = cur[k];
            else if (depth == 16) samples[k] = uint16_t((cur[2 * k] << 8) | cur[2 * k + 1]);
            else { const size_t bit = k * size_t(depth); samples[k] = uint16_t((cur[bit / 8] >> (8 - depth - int(bit % 8))) & ((1 << depth) - 1)); }
        }
        auto to8 = [&](uint16_t v) { return uint8_t(depth == 16 ? (v >> 8) : v); }; // 16-bit samples keep their high byte (as stbi_load's 8-bit interface does)
        for (uint32_t x = 0; x < w; ++x) {
            uint8_t* o = &img.pixels[(size_t(y) * w + x) * 4];
            const uint16_t* sp = &samples[size_t(x) * channels];
            switch (colour) {
                case 0: o[0] = o[1] = o[2] = uint8_t(to8(sp[0]) * (depth < 8 ? scale : 1)); if (keyed && sp[0] == key[0]) o[3] = 0; break;
                case 2: o[0] = to8(sp[0]); o[1] = to8(sp[1]); o[2] = to8(sp[2]); if (keyed && sp[0] == key[0] && sp[1] == key[1] && sp[2] == key[2]) o[3] = 0; break;
                case 3: { const size_t k = sp[0]; if (3 * k + 2 >= palette.size()) return fail(error, "PNG palette index out of range"); o[0] = palette[3 * k]; o[1] = palette[3 * k + 1]; o[2] = palette[3 * k + 2]; if (k < trns.size()) o[3] = trns[k]; break; }
                case 4: o[0] = o[1] = o[2] = to8(sp[0]); o[3] = to8(sp[1]); break;
                case 6: o[0] = to8(sp[0]); o[1] = to8(sp[1]); o[2] = to8(sp[2]); o[3] = to8(sp[3]); break;
            }
        }
        prev.swap(cur);
    }
    return true;
}

// Netpbm: P6 (binary RGB, maxval 255) and P7 / PAM (DEPTH 4, MAXVAL 255)
bool decodeNetpbm(const std::vector<uint8_t>& f, Image& img, std::string* error) {
    if (f[1] == '7') {
        const char* end = static_cast<const char*>(memmem(f.data(), f.size(), "ENDHDR\n", 7));
        if (!end) return fail(error, "PAM header without ENDHDR");
        std::istringstream hdr(std::string(reinterpret_cast<const char*>(f.data()), size_t(end - reinterpret_cast<const char*>(f.data()))));
        std::string key; uint32_t w = 0, h = 0, depth = 0, maxval = 0;
        hdr >> key; // P7
        while (hdr >> key) {
            if (key == "WIDTH") hdr >> w; else if (key == "HEIGHT") hdr >> h; else if (key == "DEPTH") hdr >> depth; else if (key == "MAXVAL") hdr >> maxval;
            else { std::string rest; std::getline(hdr, rest); }
        }
        const size_t data = size_t(end - reinterpret_cast<const char*>(f.data())) + 7;
        if (depth != 4 || maxval != 255 || w == 0 || h == 0 || f.size() < data + size_t(w) * h * 4) return fail(error, "unsupported PAM (needs DEPTH 4, MAXVAL 255)");
        img.width = w; img.height = h; img.pixels.assign(f.begin() + long(data), f.begin() + long(data + size_t(w) * h * 4));
        return true;
    }
    size_t off = 2; uint32_t vals[3]; int got = 0;
    while (got < 3 && off < f.size()) {
        if (f[off] == '#') { while (off < f.size() && f[off] != '\n') ++off; continue; }
        if (isspace(f[off])) { ++off; continue; }
        uint32_t v = 0; while (off < f.size() && isdigit(f[off])) v = v * 10 + uint32_t(f[off++] - '0');
        vals[got++] = v;
    }
    ++off; // the single whitespace after maxval
    if (got < 3 || vals[2] != 255 || vals[0] == 0 || vals[1] == 0 || f.size() < off + size_t(vals[0]) * vals[1] * 3) return fail(error, "unsupported PPM (needs P6, maxval 255)");
    img.width = vals[0]; img.height = vals[1]; img.pixels.resize(size_t(img.width) * img.height * 4);
    for (size_t i = 0; i < size_t(img.width) * img.height; ++i) { img.pixels[4 * i] = f[off + 3 * i]; img.pixels[4 * i + 1] = f[off + 3 * i + 1]; img.pixels[4 * i + 2] = f[off + 3 * i + 2]; img.pixels[4 * i + 3] = 255; }
    return true;
}

} // namespace

bool Image::load(const std::string& path, std::string* error) {
    std::ifstream file(path, std::ios::binary | std::ios::ate);
    if (!file) return fail(error, "could not open '" + path + "'");
    const std::streamsize size = file.tellg();
    file.seekg(0, std::ios::beg);
    std::vector<uint8_t> f(static_cast<size_t>(size));
    if (size < 16 || !file.read(reinterpret_cast<char*>(f.data()), size)) return fail(error, "could not read '" + path + "'");
    static const uint8_t pngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (!std::memcmp(f.data(), pngSig, 8)) return decodePng(f, *this, error);
    if (f[0] == 'P' && (f[1] == '6' || f[1] == '7')) return decodeNetpbm(f, *this, error);
    return fail(error, "'" + path + "': not a PNG / P6 / P7 image");
}

bool Image::savePam(const std::string& path) const {
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    f << "P7\nWIDTH " << width << "\nHEIGHT " << height << "\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n";
    f.write(reinterpret_cast<const char*>(pixels.data()), std::streamsize(pixels.size()));
    return bool(f);
}

Image Image::blank() { Image i; i.width = i.height = 1; i.pixels = {255, 255, 255, 255}; return i; }

} // namespace vkx

// C ABI: the image decoder on its own, for callers that bind the library without the C++ facade (include/vkx.h).
extern "C" int vkx_image_decode(const char* path, uint8_t* rgba, size_t rgbaBytes, uint32_t* width, uint32_t* height) {
    if (!path || !width || !height) return VKX_E_INVALID;
    try { // no exception may cross the C boundary
        vkx::Image img;
        std::string err;
        if (!img.load(path, &err)) { std::fprintf(stderr, "vkx_image_decode: %s\n", err.c_str()); return VKX_E_UNSUPPORTED; }
        *width = img.width; *height = img.height;
        if (!rgba) return VKX_OK;
        if (rgbaBytes < img.pixels.size()) return VKX_E_INVALID;
        std::memcpy(rgba, img.pixels.data(), img.pixels.size());
        return VKX_OK;
    } catch (const std::bad_alloc&) { return VKX_E_NOMEM;
    } catch (const std::exception& e) { std::fprintf(stderr, "vkx_image_decode: %s\n", e.what()); return VKX_E_UNSUPPORTED; }
}
