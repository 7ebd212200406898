// headless_viewer.cpp — the reference's RealtimeViewer (RealtimeViewer/Main.cpp) without the window:
// same calls in the same order (OnInit :32-62, OnRender :65-75), frames go to a BMP instead of
// glDrawPixels. Usage: headless_viewer [frames] [out.bmp] [dump.bin] [mesh.obj] [msaa_log2] [frames_in_flight] [texture_filter] [mtl]
// With frames_in_flight > 1 the loop keeps that many frames on the GPU (FrameRing) and reads each back in order.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/edxraster/Renderer.h"

using namespace edx_b200;

int main(int argc, char** argv)
{
    const int frames = argc > 1 ? std::atoi(argv[1]) : 100;
    const char* out = argc > 2 ? argv[2] : "frame.bmp";
    const int W = 1280, H = 720;                                   // Main.cpp:18-19

    Renderer renderer;                                             // Main.cpp:37
    renderer.Initialize(W, H);                                     // Main.cpp:38
    if (renderer.LastStatus() != EDX_OK) { std::fprintf(stderr, "init failed: %s\n", renderer.LastError()); return 1; }
    Camera camera;
    camera.Init(Vector3(0, 0, -5), Vector3(0, 0, 0), Vector3(0, 1, 0), W, H, 65, 0.01f);   // Main.cpp:39
    Mesh mesh;
    if (argc > 4 && argv[4][0]) {
        if (!mesh.LoadMesh(Vector3(0, 0, 0), Vector3(1, 1, 1), Vector3(0, 30, 0), argv[4])) {   // Main.cpp:44-51 (LoadMesh variants)
            std::fprintf(stderr, "cannot load %s\n", argv[4]);
            return 1;
        }
    } else {
        mesh.LoadSphere(Vector3(0, 0, 0), Vector3(1, 1, 1), Vector3(0, 0, 0), 1.2f);        // Main.cpp:42
    }
    if (argc > 5) renderer.SetMSAAMode(std::atoi(argv[5]));                                 // Main.cpp:97
    renderer.SetPixelShader(PixelShaderKind::BlinnPhong);
    if (argc > 7) {
        // the reference's default shader (LambertianAlbedoPixelShader, Renderer.cpp:41) on an image texture; the
        // texels are a formula so that a test can rebuild them: 64 x 32 RGBA8
        std::vector<_byte> tex(64 * 32 * 4);
        for (int y = 0; y < 32; y++)
            for (int x = 0; x < 64; x++) {
                _byte* p = &tex[4 * (y * 64 + x)];
                p[0] = (_byte)((x * 37 + y * 11) & 255); p[1] = (_byte)(((x / 4 + y / 4) & 1) ? 230 : 40); p[2] = (_byte)((x * y * 3) & 255); p[3] = 255;
            }
        if (argc <= 8) {                           // argv[8] ("mtl"): keep the textures and slots the OBJ's materials gave the mesh
            mesh.AddImageTexture(tex.data(), 64, 32);
            std::vector<uint> ids(mesh.GetIndexBuffer()->GetTriangleCount());
            for (size_t i = 0; i < ids.size(); i++) ids[i] = (i / 5) % 2 ? (uint)(mesh.GetTextureCount() - 1) : 0u;      // image / the mesh's first slot
            mesh.SetTextureIds(ids);
        }
        renderer.SetPixelShader(PixelShaderKind::LambertianAlbedo);
        renderer.SetTextureFilter(TextureFilter(std::atoi(argv[7])));                       // Main.cpp:108
    }

    const int inFlight = argc > 6 ? std::atoi(argv[6]) : 1;
    std::vector<_byte> ringFrame;                                  // last frame that came out of the ring
    bool ringSame = true;
    auto t0 = std::chrono::steady_clock::now();
    if (inFlight > 1) {
        FrameRing ring(inFlight);
        ring.Initialize(W, H);
        if (argc > 5) ring.SetMSAAMode(std::atoi(argv[5]));
        ring.SetPixelShader(PixelShaderKind::BlinnPhong);
        for (int f = 0; f < frames + inFlight - 1; f++) {
            if (f < frames) ring.Submit(mesh, camera.GetViewMatrix(), camera.GetProjMatrix(), camera.GetRasterMatrix());
            if (f < inFlight - 1) continue;
            const _byte* px = ring.GetBackBuffer((size_t)(f - inFlight + 1));
            if (!px) { std::fprintf(stderr, "frame failed: %s\n", ring.Lane(f - inFlight + 1).LastError()); return 1; }
            if (!ringFrame.empty() && std::memcmp(ringFrame.data(), px, ringFrame.size()) != 0) ringSame = false;
            ringFrame.assign(px, px + (size_t)W * H * 4);
        }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("%.1f frames/s (render + read-back, %d frames in flight)\n", frames / s, inFlight);
        // show the last frame through the plain renderer path below as well (same image)
        t0 = std::chrono::steady_clock::now();
    }
    for (int f = 0; f < frames; f++) {                             // OnRender
        camera.Transform();
        renderer.SetTransform(camera.GetViewMatrix(), camera.GetProjMatrix(), camera.GetRasterMatrix());   // Main.cpp:71
        renderer.RenderMesh(mesh);                                 // Main.cpp:72
        const _byte* px = renderer.GetBackBuffer();                // Main.cpp:75
        if (!px) { std::fprintf(stderr, "frame failed: %s\n", renderer.LastError()); return 1; }
    }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("Image Res: %i, %i\nTriangle Count: %u\n%.1f frames/s (render + read-back)\n", W, H,
                mesh.GetIndexBuffer()->GetTriangleCount(), frames / s);
    if (inFlight > 1) {
        const _byte* px = renderer.GetBackBuffer();
        ringSame = ringSame && px && std::memcmp(ringFrame.data(), px, ringFrame.size()) == 0;
        std::printf("ring frames identical to the single-frame path: %s\n", ringSame ? "yes" : "NO");
        if (!ringSame) return 1;
    }
    if (!renderer.WriteFrame(out)) { std::fprintf(stderr, "write failed: %s\n", renderer.LastError()); return 1; }
    if (argc > 3) {                                                // raw inputs, so a test can replay them
        FILE* f = std::fopen(argv[3], "wb");
        uint32_t hdr[4] = { (uint32_t)W, (uint32_t)H, mesh.GetVertexBuffer()->GetVertexCount(), mesh.GetIndexBuffer()->GetTriangleCount() };
        std::fwrite(hdr, 4, 4, f);
        std::fwrite(camera.GetViewMatrix().Data(), 4, 16, f);
        std::fwrite(camera.GetProjMatrix().Data(), 4, 16, f);
        std::fwrite(camera.GetRasterMatrix().Data(), 4, 16, f);
        std::fwrite(mesh.GetVertexBuffer()->GetBuffer(), 32, hdr[2], f);
        std::fwrite(mesh.GetIndexBuffer()->GetBuffer(), 12, hdr[3], f);
        // trailer: the mesh's texture table and per-triangle slots (count; per slot kind, colour, size, texels; ids)
        const uint32_t nTex = (uint32_t)mesh.GetTextureCount();
        std::fwrite(&nTex, 4, 1, f);
        for (uint32_t k = 0; k < nTex; k++) {
            int kind; float color[3]; uint w, h; const _byte* px;
            mesh.GetTexture(k, kind, color, w, h, px);
            std::fwrite(&kind, 4, 1, f); std::fwrite(color, 4, 3, f); std::fwrite(&w, 4, 1, f); std::fwrite(&h, 4, 1, f);
            if (kind == EDX_TEXTURE_IMAGE) std::fwrite(px, 4, (size_t)w * h, f);
        }
        const uint32_t nIds = (uint32_t)mesh.GetTextureIds().size();
        std::fwrite(&nIds, 4, 1, f);
        std::fwrite(mesh.GetTextureIds().data(), 4, nIds, f);
        std::fclose(f);
    }
    return 0;
}
