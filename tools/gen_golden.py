"""Generates tests/golden/*.json from the reference checkout (/root/reference). Run in the build container only;
the fixtures are committed because /root/reference does not exist on the GPU box.

  border_tables.json  the constant copy tables of src/shaders/probesCopyBorders.comp:21-220, parsed from the shader text
  glm_pin.json        glm::sphericalRand / genBasis / mat3-transpose results computed by the reference's vendored GLM
                      0.9.9.8 (ext/glm) through a tiny driver program compiled here, with std::rand replaced by the MSVC
                      LCG the reference runs on (Windows-only project)
  stb_pin.json + img/ small PNG / PPM files (written here with Pillow, deterministic content) and the SHA-256 of what the reference's
                      vendored decoder returns for them: stbi_load(path, &x, &y, &n, 4) exactly as src/STBImage.hpp:25 calls it
                      (ext/stb_image.h compiled from where it lies). Pins the library's own decoders (csrc/host/Image.cpp)
  layouts.json        sizeof/offsetof of the POD structs, measured by compiling the reference's own headers where they
                      are self-contained (Vertex.hpp needs Vulkan, so those offsets are computed from its member list)
"""
import json
import os
import re
import subprocess
import sys
import tempfile

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def border_tables():
    src = open(os.path.join(REF, "src/shaders/probesCopyBorders.comp")).read()
    out = {}
    for name in ("irradianceCopiesDst", "irradianceCopiesSrc", "depthCopiesDst", "depthCopiesSrc"):
        m = re.search(r"ivec2\s+" + name + r"\[[^\]]*\]\s*=\s*\{(.*?)\};", src, re.S)
        body = re.sub(r"//[^\n]*", "", m.group(1))
        out[name] = [[int(a), int(b)] for a, b in re.findall(r"ivec2\(\s*(\d+)\s*,\s*(\d+)\s*\)", body)]
    assert len(out["irradianceCopiesDst"]) == 28 and len(out["depthCopiesDst"]) == 60
    return out


GLM_DRIVER = r"""
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
// MSVC CRT rand(): the reference only builds with Visual Studio (VulkanExp.vcxproj). Defining rand() in the executable
// interposes libc's, so the unmodified GLM headers below draw from the MSVC sequence (seed 1, never seeded by the reference).
static uint32_t g_x = 1;
extern "C" int rand(void) noexcept { g_x = g_x * 214013u + 2531011u; return int((g_x >> 16) & 0x7fff); }
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtc/random.hpp>
static void genBasis(const glm::vec3& n, glm::vec3& b1, glm::vec3& b2) { // src/IrradianceProbes.cpp:347-355 (restated)
    if (n.x > 0.9f) b1 = glm::vec3(0.0f, 1.0f, 0.0f); else b1 = glm::vec3(1.0f, 0.0f, 0.0f);
    b1 -= n * glm::dot(b1, n);
    b1 = glm::normalize(b1);
    b2 = glm::cross(n, b1);
}
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main() {
    printf("[");
    for (int i = 0; i < 24; ++i) {
        glm::vec3 Z = glm::sphericalRand(1.0f);
        glm::vec3 X, Y; genBasis(Z, X, Y);
        glm::mat4 M = glm::mat4(glm::transpose(glm::mat3(X, Y, Z)));
        printf("%s{\"Z\":[%u,%u,%u],\"M\":[", i ? "," : "", bits(Z.x), bits(Z.y), bits(Z.z));
        const float* p = &M[0][0];
        for (int k = 0; k < 16; ++k) printf("%s%u", k ? "," : "", bits(p[k]));
        printf("]}");
    }
    printf("]\n");
    return 0;
}
"""


def glm_pin():
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "glm_pin.cpp")
        open(src, "w").write(GLM_DRIVER)
        exe = os.path.join(td, "glm_pin")
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(REF, "ext/glm"), src, "-o", exe])
        return json.loads(subprocess.check_output([exe]).decode())


STB_DRIVER = r"""
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include <cstdio>
int main(int argc, char** argv) { // prints: width height channels-in-file, then the RGBA bytes as hex
    for (int i = 1; i < argc; ++i) {
        int x = 0, y = 0, n = 0;
        unsigned char* d = stbi_load(argv[i], &x, &y, &n, 4); // src/STBImage.hpp:25 ("Force 4 channels")
        if (!d) { printf("FAIL\n"); continue; }
        printf("%d %d %d ", x, y, n);
        for (int k = 0; k < x * y * 4; ++k) printf("%02x", d[k]);
        printf("\n");
        stbi_image_free(d);
    }
    return 0;
}
"""


def stb_pin():
    import hashlib

    import numpy as np
    from PIL import Image

    img_dir = os.path.join(OUT, "img")
    os.makedirs(img_dir, exist_ok=True)
    rng = np.random.default_rng(0x57B)
    rgba = rng.integers(0, 256, (13, 19, 4), dtype=np.uint8)
    smooth = np.zeros((16, 24, 4), dtype=np.uint8)  # gradients: exercises the Sub / Up / Average / Paeth filters
    yy, xx = np.mgrid[0:16, 0:24]
    smooth[..., 0], smooth[..., 1], smooth[..., 2], smooth[..., 3] = xx * 10, yy * 15, (xx + yy) * 6, 255 - xx * 9
    files = {}
    Image.fromarray(rgba, "RGBA").save(os.path.join(img_dir, "rgba_noise.png"))
    Image.fromarray(smooth, "RGBA").save(os.path.join(img_dir, "rgba_smooth.png"), optimize=True)
    Image.fromarray(rgba[..., :3].copy(), "RGB").save(os.path.join(img_dir, "rgb.png"))
    Image.fromarray(smooth[..., 0].copy(), "L").save(os.path.join(img_dir, "grey.png"))
    Image.fromarray(np.stack([smooth[..., 1], smooth[..., 3]], axis=-1).copy(), "LA").save(os.path.join(img_dir, "grey_alpha.png"))
    pal = Image.fromarray(smooth[..., :3].copy(), "RGB").quantize(16)
    pal.save(os.path.join(img_dir, "palette.png"))
    pal.save(os.path.join(img_dir, "palette_trns.png"), transparency=3)
    Image.fromarray((smooth[..., 0].astype(np.uint16) * 257 + 3), "I;16").save(os.path.join(img_dir, "grey16.png"))  # 16 bits per sample
    Image.fromarray(((xx + yy) % 3 == 0)).convert("1").save(os.path.join(img_dir, "bw1.png"))                   # 1 bit per sample
    key = tuple(int(v) for v in smooth[5, 7, :3])
    Image.fromarray(smooth[..., :3].copy(), "RGB").save(os.path.join(img_dir, "rgb_keyed.png"), transparency=key)  # tRNS colour key
    Image.fromarray((smooth[..., 1] // 64 * 85).copy(), "L").quantize(4).convert("L").save(os.path.join(img_dir, "grey_coarse.png"))
    with open(os.path.join(img_dir, "rgb.ppm"), "wb") as f:
        f.write(b"P6\n19 13\n255\n" + rgba[..., :3].tobytes())
    names = sorted(os.listdir(img_dir))
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "stb_pin.cpp")
        open(src, "w").write(STB_DRIVER)
        exe = os.path.join(td, "stb_pin")
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-w", "-I", os.path.join(REF, "ext"), src, "-o", exe])
        lines = subprocess.check_output([exe] + [os.path.join(img_dir, n) for n in names]).decode().splitlines()
    for n, line in zip(names, lines):
        w, h, c, hexbytes = line.split()
        files[n] = {"width": int(w), "height": int(h), "channels_in_file": int(c), "sha256": hashlib.sha256(bytes.fromhex(hexbytes)).hexdigest()}
    return files


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; fixtures are already committed")
    os.makedirs(OUT, exist_ok=True)
    json.dump(border_tables(), open(os.path.join(OUT, "border_tables.json"), "w"))
    json.dump(glm_pin(), open(os.path.join(OUT, "glm_pin.json"), "w"))
    json.dump(stb_pin(), open(os.path.join(OUT, "stb_pin.json"), "w"), indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
