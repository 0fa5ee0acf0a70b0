// drawer_main.cpp -- the reference's Awake()/Update() sequence driven from C++ through usrt_host.hpp.
//   drawer_main <triangles.bin> <W> <H> <near> <tanHalfFov> <16 matrix floats...> <hits.bin>
// Reads n packed 128-byte Triangles, builds step by step as the reference does, rebuilds fused, traces and
// writes the hit records. Used by tests/test_cpp_host.py, which compares the output with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "usrt_host.hpp"

int main(int argc, char** argv) {
    if (argc != 23) { std::fprintf(stderr, "usage: %s tris.bin W H near tanHalfFov m[16] hits.bin\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
        const size_t bytes = (size_t)f.tellg();
        std::vector<usrt_triangle> mesh(bytes / sizeof(usrt_triangle));
        f.seekg(0); f.read(reinterpret_cast<char*>(mesh.data()), bytes);
        const int W = std::atoi(argv[2]), H = std::atoi(argv[3]);
        const float nearPlane = (float)std::atof(argv[4]), fov = (float)std::atof(argv[5]);
        float m[16];
        for (int i = 0; i < 16; ++i) m[i] = (float)std::atof(argv[6 + i]);

        usrt::RaytracingMeshDrawer drawer;
        drawer.Awake(mesh);                                        // step by step (RaytracingMeshDrawer.cs:34-51)
        std::vector<usrt_raycast_result> a = drawer.Update(W, H, nearPlane, fov, m);
        drawer.Rebuild();                                          // fused
        std::vector<usrt_raycast_result> b = drawer.Update(W, H, nearPlane, fov, m);
        if (a.size() != b.size() || std::memcmp(a.data(), b.data(), a.size() * sizeof(a[0])) != 0) {
            std::fprintf(stderr, "step-by-step and fused rebuild disagree\n"); return 1;
        }
        std::vector<uint32_t> keys(drawer.container().TrianglesLength());
        drawer.container().Keys().GetData(keys);
        for (size_t i = 1; i < keys.size(); ++i)
            if (keys[i] <= keys[i - 1]) { std::fprintf(stderr, "distributed keys not strictly increasing at %zu\n", i); return 1; }
        std::ofstream o(argv[22], std::ios::binary);
        o.write(reinterpret_cast<const char*>(b.data()), b.size() * sizeof(b[0]));
        std::printf("ok n=%zu rays=%zu\n", mesh.size(), b.size());
        return 0;
    } catch (const usrt::Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
