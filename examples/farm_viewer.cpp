// farm_viewer.cpp — BASELINE config 5 in miniature through the C++ layer: many camera views of one mesh rendered on
// every GPU of the box (view i on GPU i mod N, several frames in flight per GPU), each finished colour buffer pushed by
// the copy engine into GPU 0's frame store over NVLink (FrameFarm / edx_set_frame_sink). The reference renders one
// view per iteration of its viewer loop on the host cores (RealtimeViewer/Main.cpp:65-75).
// Usage: farm_viewer [views] [gpus (0 = all)] [slices]. Exit code 0 iff every farmed frame equals the same view
// rendered by a single Renderer on GPU 0.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/edxraster/Renderer.h"

using namespace edx_b200;

int main(int argc, char** argv)
{
    const int views = argc > 1 ? std::atoi(argv[1]) : 64;
    const int gpus = argc > 2 ? std::atoi(argv[2]) : 0;
    const int slices = argc > 3 ? std::atoi(argv[3]) : 400;
    const int W = 1280, H = 720;

    FrameFarm farm(gpus, 3);
    if (farm.Gpus() == 0) { std::fprintf(stderr, "no B200 visible\n"); return 2; }
    farm.Initialize(W, H);
    farm.SetPixelShader(PixelShaderKind::BlinnPhong);
    farm.LoadMeshes([&](Mesh& m) { m.LoadSphere(Vector3(0, 0, 0), Vector3(1, 1, 1), Vector3(0, 0, 0), 1.2f, slices, slices); });

    std::vector<ViewTransform> xf(views);
    for (int i = 0; i < views; i++) {
        const float a = 6.2831853f * float(i) / float(views);
        Camera cam;
        cam.Init(Vector3(5.0f * std::sin(a), 1.5f * std::cos(2 * a), -5.0f * std::cos(a)), Vector3(0, 0, 0), Vector3(0, 1, 0), W, H, 65, 0.01f);
        xf[i].modelView = cam.GetViewMatrix(); xf[i].proj = cam.GetProjMatrix(); xf[i].toRaster = cam.GetRasterMatrix();
    }
    if (!farm.Render(xf)) { std::fprintf(stderr, "farm render failed\n"); return 1; }          // warm-up: uploads, queue growth
    const auto t0 = std::chrono::steady_clock::now();
    if (!farm.Render(xf)) { std::fprintf(stderr, "farm render failed\n"); return 1; }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%d views of %d triangles on %d GPU(s): %.2f ms, %.0f frames/s, frames gathered on GPU 0\n",
                views, 2 * slices * slices, farm.Gpus(), s * 1e3, views / s);

    // every frame must equal the single-GPU rendering of the same view
    Renderer single(0);
    single.Initialize(W, H);
    single.SetPixelShader(PixelShaderKind::BlinnPhong);
    Mesh mesh;
    mesh.LoadSphere(Vector3(0, 0, 0), Vector3(1, 1, 1), Vector3(0, 0, 0), 1.2f, slices, slices);
    std::vector<_byte> got((size_t)W * H * 4);
    int bad = 0;
    for (int i = 0; i < views; i++) {
        single.SetTransform(xf[i].modelView, xf[i].proj, xf[i].toRaster);
        single.RenderMesh(mesh);
        const _byte* want = single.GetBackBuffer();
        if (!farm.GetFrame((size_t)i, got.data()) || !want || std::memcmp(want, got.data(), got.size()) != 0) bad++;
    }
    std::printf("%d of %d farmed frames differ from the single-GPU frames\n", bad, views);
    return bad ? 1 : 0;
}
