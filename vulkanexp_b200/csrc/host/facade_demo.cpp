// Drives the C++ facade the way Editor::run / initVulkan / mainLoop drive the reference classes:
//   vkx_facade_demo <scene.scene> <rx> <ry> <rz> <raysPerProbe> <frames> <out.bin> [probesPerUpdate [device|host [skinMesh]]]
// (`device` selects the on-device scheduler instead of the state read-back + host loop of the reference; skinMesh adds a skinned
// renderer of that mesh on the root node, posed anew before every update: three joints, joint j scaled by 1 + j / 8 and moved by
// (frame * (j + 1) / 32, frame / 64, 0), vertex v bound to joints (v + k) mod 3 with weights 0.4, 0.3, 0.2, 0.1)
// Writes irradiance, depth and state arrays (u32) to out.bin so tests can compare them with the C-ABI path.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "IrradianceProbes.hpp"

int main(int argc, char** argv) {
    if (argc < 8) { std::fprintf(stderr, "usage: %s scene rx ry rz rays frames out.bin [probesPerUpdate [device]]\n", argv[0]); return 2; }
    try {
        vkx::Scene scene;
        if (!scene.load(argv[1])) return 1;
        scene.update(); // the first Scene::update of the main loop (propagates transforms)
        const int skinMesh = argc > 10 ? std::atoi(argv[10]) : -1;
        if (skinMesh >= 0) {
            vkx::SkinnedMeshRenderer r;
            r.node = scene.getRoot(); r.meshIndex = uint32_t(skinMesh); r.materialIndex = scene.getMeshes().at(size_t(skinMesh)).defaultMaterialIndex;
            const size_t nv = scene.getMeshes().at(size_t(skinMesh)).vertices.size();
            const float w[4] = {0.4f, 0.3f, 0.2f, 0.1f};
            for (size_t v = 0; v < nv; ++v) for (int k = 0; k < 4; ++k) { r.joints.push_back(uint16_t((v + size_t(k)) % 3)); r.weights.push_back(w[k]); }
            scene.getSkinnedRenderers().push_back(r);
        }
        auto pose = [](int frame) {
            std::vector<vkx::mat4> js(3);
            for (int j = 0; j < 3; ++j) {
                const float s = 1.0f + float(j) / 8.0f;
                js[size_t(j)].m[0][0] = js[size_t(j)].m[1][1] = js[size_t(j)].m[2][2] = s;
                js[size_t(j)].m[3][0] = float(frame) * float(j + 1) / 32.0f; js[size_t(j)].m[3][1] = float(frame) / 64.0f;
            }
            return std::vector<std::vector<vkx::mat4>>{js};
        };
        vkx::Device device(0);
        vkx::Renderer renderer; renderer.setDevice(device); renderer.setScene(scene);
        renderer.allocateMeshes();
        renderer.createAccelerationStructures();
        vkx::LightBuffer light;
        vkx::IrradianceProbes probes;
        probes.GridParameters.resolution[0] = std::atoi(argv[2]); probes.GridParameters.resolution[1] = std::atoi(argv[3]); probes.GridParameters.resolution[2] = std::atoi(argv[4]);
        probes.GridParameters.raysPerProbe = unsigned(std::atoi(argv[5]));
        if (argc > 8) probes.ProbesPerUpdate = unsigned(std::atoi(argv[8]));
        if (argc > 9) probes.DeviceScheduler = std::string(argv[9]) == "device";
        probes.init(device, scene.getBounds().min, scene.getBounds().max);
        probes.createPipeline();
        probes.writeDescriptorSet(renderer, light);
        probes.initProbes();
        const int frames = std::atoi(argv[6]);
        for (int f = 0; f < frames; ++f) {
            if (skinMesh >= 0) { renderer.updateSkinnedVertexBuffer(pose(f)); renderer.updateSkinnedBLAS(); } // Editor::mainLoop: skinning before the probe update
            probes.update();
        }
        std::vector<uint32_t> irr, dep, st;
        probes.download(irr, dep, st);
        FILE* fp = std::fopen(argv[7], "wb");
        if (!fp) return 1;
        std::fwrite(irr.data(), 4, irr.size(), fp); std::fwrite(dep.data(), 4, dep.size(), fp); std::fwrite(st.data(), 4, st.size(), fp);
        std::fclose(fp);
        auto info = renderer.getTLAS();
        std::printf("facade ok: %u nodes, %u triangles, last update %u probes, hysteresis %.4f, compute %.3f ms\n", info.numNodes, info.numTriangles, probes.lastUpdatedProbeCount(),
                    probes.GridParameters.hysteresis, probes.getComputeTimes().last());
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
