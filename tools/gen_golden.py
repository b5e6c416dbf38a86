"""Generates tests/golden/*.json from the reference checkout (/root/reference). Run in the build container only;
the fixtures are committed because /root/reference does not exist on the GPU box.

  border_tables.json  the constant copy tables of src/shaders/probesCopyBorders.comp:21-220, parsed from the shader text
  glm_pin.json        glm::sphericalRand / genBasis / mat3-transpose results computed by the reference's vendored GLM
                      0.9.9.8 (ext/glm) through a tiny driver program compiled here, with std::rand replaced by the MSVC
                      LCG the reference runs on (Windows-only project)
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


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; fixtures are already committed")
    os.makedirs(OUT, exist_ok=True)
    json.dump(border_tables(), open(os.path.join(OUT, "border_tables.json"), "w"))
    json.dump(glm_pin(), open(os.path.join(OUT, "glm_pin.json"), "w"))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
