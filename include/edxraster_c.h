/* edxraster_c.h — C ABI of the B200-native EDXRaster raster hot path.
 *
 * The reference has no FFI: its boundary is the C++ class EDX::RasterRenderer::Renderer in a static
 * library (EDXRaster/Core/Renderer.h:36-50) fed by Mesh / IVertexBuffer / IndexBuffer
 * (EDXRaster/Utils/Mesh.h:29-68, Utils/InputBuffer.h:43-66,148-194). Each entry point below names the
 * reference member it replaces. The C++ classes with the reference's own names live in
 * include/edxraster/Renderer.h and forward here; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns an edx_status (0 = ok, <0 = error) unless
 *    it returns a pointer; edx_last_error() gives the message. Nothing aborts or throws.
 *  - matrices: 16 floats, row-major, column-vector convention (clip = Proj * ModelView * p), exactly
 *    what Renderer::SetTransform receives (Core/Renderer.cpp:85-92).
 *  - vertices: the reference's 32-byte submission format, position(3) normal(3) texcoord(2) floats
 *    (Utils/InputBuffer.h:16-28); indices: uint32 x 3 per triangle (InputBuffer.h:151,175-179).
 *  - frame buffer: RGBA8, x fastest, row 0 = BOTTOM scanline (Core/FrameBuffer.cpp:41, Main.cpp:75).
 *    Depth and winner-id read-backs use the same bottom-up order.
 *  - one context per GPU; a context is not thread-safe, distinct contexts are independent.
 *  - there is NO CPU fallback: every call fails with EDX_ERR_NO_DEVICE / EDX_ERR_CUDA without a GPU.
 */
#ifndef EDXRASTER_C_H
#define EDXRASTER_C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct edx_context edx_context;
typedef struct edx_mesh edx_mesh;

typedef enum edx_status {
    EDX_OK = 0,
    EDX_ERR_INVALID = -1,      /* bad argument */
    EDX_ERR_CUDA = -2,         /* CUDA runtime error, see edx_last_error */
    EDX_ERR_OOM = -3,          /* device or pinned-host allocation failed */
    EDX_ERR_OVERFLOW = -4,     /* an internal queue could not be grown, or frames submitted before the last one
                                  (with no synchronising call after them) overflowed and are incomplete */
    EDX_ERR_UNSUPPORTED = -5,  /* feature outside the round's scope (e.g. a resolution beyond the int32 edge range) */
    EDX_ERR_NO_DEVICE = -6     /* no CUDA device / wrong architecture */
} edx_status;

/* Pixel shaders. The reference hard-codes LambertianAlbedoPixelShader (Core/Renderer.cpp:41) and has
 * no setter; the selector is our one API extension on the render path. */
typedef enum edx_shader {
    EDX_SHADER_DEPTH_ONLY = 0,      /* depth test only, colour buffer left cleared */
    EDX_SHADER_BLINN_PHONG = 1,     /* Core/Shader.h:246-282 */
    EDX_SHADER_LAMBERT = 2,         /* Core/Shader.h:185-207 */
    EDX_SHADER_LAMBERT_ALBEDO = 3   /* Core/Shader.h:209-244 with the constant texture of Utils/Mesh.cpp:48,67 */
} edx_shader;

typedef struct edx_stats {
    uint64_t submitted_tris;   /* triangles in the last RenderMesh call */
    uint64_t clipped_tris;     /* triangles that went through the polygon clipper */
    uint64_t binned_tris;      /* post-setup triangles routed to the tile (large-triangle) path */
    uint64_t clip_records;     /* fan triangles emitted by the clipper */
    uint32_t regrow_count;     /* times an internal queue was grown and the frame re-run */
    uint32_t tile_pairs;       /* (triangle, 64x64 bin) pairs that survived the bin-level culls: load of the tile path */
    uint64_t mid_tris;         /* post-setup triangles rasterised one warp per triangle (the mid-size path) */
    uint64_t bin_pairs;        /* (triangle, bin) entries of the per-bin lists built for a long tile-path list; 0 = shared list */
    float    stage_ms[8];      /* valid with profiling on: geom, clip, tile, total; rest 0 */
} edx_stats;

/* ---- context ------------------------------------------------------------------------------ */
/* new Renderer (RealtimeViewer/Main.cpp:37); `device` = CUDA ordinal. */
int edx_create(int device, edx_context** out);
/* Renderer::~Renderer (Core/Renderer.cpp:365-371) */
void edx_destroy(edx_context* ctx);
const char* edx_last_error(const edx_context* ctx);
/* version string and the architecture the kernels were built for ("sm_100a") */
const char* edx_version(void);

/* Renderer::Initialize (Core/Renderer.cpp:22-62) / Renderer::Resize (:64-83) */
int edx_initialize(edx_context* ctx, uint32_t width, uint32_t height);
int edx_resize(edx_context* ctx, uint32_t width, uint32_t height);
/* Renderer::SetTransform (Core/Renderer.cpp:85-92). Derives MVP = proj * model_view and the
 * model-view inverse exactly as the reference does. */
int edx_set_transform(edx_context* ctx, const float model_view[16], const float proj[16], const float to_raster[16]);
/* Renderer::SetMSAAMode (Core/Renderer.cpp:94-98): 2^log2 samples per pixel, log2 in 0..5 (sample tables
 * of Core/FrameBuffer.cpp:107-191). Re-creates the frame buffer; a no-op when the level is unchanged.
 * With MSAA the back buffer is the box-filtered resolve (FrameBuffer.cpp:70-87). */
int edx_set_msaa_mode(edx_context* ctx, int sample_count_log2);
/* Renderer::SetTextureFilter (Core/Renderer.h:48): 0 nearest, 1 linear, 2 trilinear (default, RenderStates.h:60),
 * 3 / 4 / 5 anisotropic 4x / 8x / 16x. Read by EDX_SHADER_LAMBERT_ALBEDO on meshes that own image textures. */
int edx_set_texture_filter(edx_context* ctx, int filter);
/* Renderer::SetHierarchicalRasterize (Core/Renderer.h:49). Off = per-pixel tests only; same image. */
int edx_set_hierarchical_rasterize(edx_context* ctx, int enabled);
/* Renderer::SetWriteFrames / WriteFrameToFile (Core/Renderer.h:50, Renderer.cpp:352-358): 24-bit BMP. */
int edx_write_frame_to_file(edx_context* ctx, const char* path);
/* extension: the reference hard-codes LambertianAlbedoPixelShader (Core/Renderer.cpp:41), which is
 * also the default here; its other shaders (Core/Shader.h:185-282) and a depth-only mode are selectable. */
int edx_set_pixel_shader(edx_context* ctx, int shader);
int edx_set_albedo(edx_context* ctx, float r, float g, float b);

/* ---- meshes -------------------------------------------------------------------------------- */
/* CreateVertexBuffer + CreateIndexBuffer (Utils/InputBuffer.h:136-146,196-205): copies the caller's
 * host arrays to the device (SoA streams). tex_ids may be NULL (Mesh::GetTextureIds, Mesh.h:56-59). */
int edx_mesh_create(edx_context* ctx, const void* vertices_pnt32, uint32_t vertex_count,
                    const uint32_t* indices, uint32_t triangle_count, const uint32_t* tex_ids, edx_mesh** out);
/* A mesh is read-only while it renders: after edx_mesh_create returns, any context on the same device may
 * render it, concurrently (several frames in flight on one GPU = several contexts, each on its own stream,
 * sharing the meshes). */
/* Re-upload into an existing mesh of the same or smaller size (streaming geometry). Asynchronous on ctx's
 * stream: contexts other than ctx that share the mesh must not render it until edx_synchronize(ctx). */
int edx_mesh_update(edx_context* ctx, edx_mesh* mesh, const void* vertices_pnt32, uint32_t vertex_count,
                    const uint32_t* indices, uint32_t triangle_count);
/* Mesh::mTextures and GetTextureIds (Utils/Mesh.h:23,54-59; Mesh.cpp:26-29,47,66): the textures a mesh owns and
 * one slot index per triangle (NULL = keep the ids given to edx_mesh_create, or slot 0 everywhere), read by EDX_SHADER_LAMBERT_ALBEDO
 * (LambertianAlbedoPixelShader, Core/Shader.h:209-244) under the filter of edx_set_texture_filter. A constant
 * texture is ConstantTexture2D<Color>; an image texture is ImageTexture<Color, Color4b>(path, gamma 1) given as
 * decoded RGBA8 texels, row 0 at v = 0, repeat addressing; its mip chain is built on the device. A mesh with no
 * textures is shaded with the context's constant albedo (edx_set_albedo). Replaces any previous set; count = 0
 * removes them. Synchronous. The sampler is our definition (EDXUtil's Texture2D is absent): DESIGN.md shims 19-24. */
typedef enum edx_texture_kind { EDX_TEXTURE_CONSTANT = 0, EDX_TEXTURE_IMAGE = 1 } edx_texture_kind;
typedef struct edx_texture_desc {
    int kind;               /* edx_texture_kind */
    float color[3];         /* constant textures */
    const uint8_t* rgba8;   /* image textures: width * height * 4 bytes */
    uint32_t width, height;
} edx_texture_desc;
int edx_mesh_set_textures(edx_context* ctx, edx_mesh* mesh, const edx_texture_desc* textures, uint32_t count,
                          const uint32_t* triangle_texture_ids);
/* diagnostics: one mip level of an image texture back to the host (out_rgba8 may be NULL to query the size) */
int edx_mesh_read_texture_level(edx_context* ctx, const edx_mesh* mesh, uint32_t slot, uint32_t level,
                                uint8_t* out_rgba8, uint32_t* out_width, uint32_t* out_height);
/* Mesh::Release (Utils/Mesh.cpp:72-78) */
int edx_mesh_destroy(edx_context* ctx, edx_mesh* mesh);

/* ---- the hot path -------------------------------------------------------------------------- */
/* Renderer::RenderMesh (Core/Renderer.cpp:100-118): clear, vertex transform, clip + setup, raster,
 * depth test, shade, frame-buffer update. Asynchronous on the context's stream. */
int edx_render_mesh(edx_context* ctx, const edx_mesh* mesh);
/* Renderer::GetBackBuffer (Core/Renderer.cpp:360-363): waits for the frame, copies it to a pinned
 * host mirror and returns a borrowed pointer valid until the next RenderMesh / Resize. NULL on error. */
const uint8_t* edx_get_back_buffer(edx_context* ctx);
/* Wait for all queued work on the context's stream and vet the frames submitted since the last synchronising
 * call (edx_synchronize, edx_get_back_buffer, edx_read_*): if the last frame overflowed an internal queue the
 * queue is grown and the frame rendered again; if an earlier one did, the queues are grown and the call
 * returns EDX_ERR_OVERFLOW once - those frames must be submitted again. */
int edx_synchronize(edx_context* ctx);

/* ---- read-backs for parity (ours; the reference keeps depth private, FrameBuffer.h:20) ------ */
int edx_read_depth(edx_context* ctx, float* out_w_times_h);
/* per pixel: submitted-triangle index * 8 + fan index of the fragment that owns it, 0xFFFFFFFF = none.
 * Needs edx_set_capture_ids(ctx, 1) before the frame. */
int edx_set_capture_ids(edx_context* ctx, int enabled);
int edx_read_winner_ids(edx_context* ctx, uint32_t* out_w_times_h);
/* per-sample read-back with MSAA: depth and/or owner ids of sample `sample` (either pointer may be NULL);
 * edx_read_depth / edx_read_winner_ids return sample 0 */
int edx_read_sample(edx_context* ctx, int sample, float* depth_w_times_h, uint32_t* ids_w_times_h);
/* stage dumps: clip-space vertices (vertex_count x 4 floats) of stage a1, Core/Renderer.cpp:120-127 */
int edx_debug_clip_vertices(edx_context* ctx, const edx_mesh* mesh, float* out_xyzw);
/* post-setup triangles (stages a3-a6) sorted by prim id; ints: prim,v0x,v0y,v1x,v1y,v2x,v2y;
 * floats: z0,z1,z2,invW0,invW1,invW2,invDet. Returns the count through *count (<= capacity). */
int edx_debug_raster_triangles(edx_context* ctx, const edx_mesh* mesh, uint64_t capacity,
                               int32_t* ints7, float* floats7, uint64_t* count);
/* MVP, eye position and normalised light direction the shaders will use */
int edx_get_derived_state(const edx_context* ctx, float mvp[16], float eye[3], float light[3]);

/* ---- device-side access and measurement ----------------------------------------------------- */
/* device pointers of the current colour (RGBA8) / depth (f32) buffers, for NCCL gathers without a
 * host round trip */
void* edx_device_color(edx_context* ctx);
void* edx_device_depth(edx_context* ctx);
/* render into caller-owned device buffers (width*height RGBA8 / float32), e.g. torch tensors that
 * an NCCL gather then sends without a copy; NULL restores the context's own buffer. */
int edx_set_render_target(edx_context* ctx, void* device_color, void* device_depth);
/* Frame-parallel gather (the call site it replaces is the per-frame read-back loop of
 * RealtimeViewer/Main.cpp:71-75 run once per GPU): after every frame this context renders, the finished colour
 * and / or depth buffer (width*height RGBA8 / float32) is pushed to `remote_color` / `remote_depth` by the COPY
 * ENGINE, ordered behind the frame (no SM time; see edx_flush_frame_sink). The addresses are device pointers
 * this GPU can reach - its own memory, or a peer's mapped over NVLink (cudaDeviceEnablePeerAccess, or a
 * symmetric-memory mapping) - typically this rank's slot of the root GPU's frame store. NULL, NULL switches it off.
 * Single-sample only. */
int edx_set_frame_sink(edx_context* ctx, void* remote_color, void* remote_depth);
/* ... and the protocol, if the caller wants one without a collective: `remote_word` (a 4-byte-aligned device address
 * this GPU can reach, usually on the root GPU next to the frame store) receives, behind the pushes of every frame and
 * in stream order, the number of frames this context has finished (and pushed, if a sink is set) since this call
 * (1, 2, 3 ...; system-scope store) - the root GPU's own contexts render straight into the store (edx_set_render_target).
 * The consumer polls the word; slot reuse is the caller's ring discipline. NULL switches it off. */
int edx_set_frame_sink_signal(edx_context* ctx, void* remote_word);
/* The pushes run beside the context's stream (the next frame's geometry does not wait for them; only its final pass,
 * which overwrites the buffers, does). edx_synchronize waits for them; work that the CALLER queues on the context's
 * stream (edx_set_stream) and that must see the pushes finished calls this first: the stream then waits for every
 * push issued so far. No host synchronisation. */
int edx_flush_frame_sink(edx_context* ctx);
/* Helpers for a frame farm inside ONE process (one context per GPU; include/edxraster/Renderer.h FrameFarm). They
 * only wrap what a caller without the CUDA runtime needs around edx_set_frame_sink: */
int edx_device_count(void);                                            /* B200 devices visible to the process */
int edx_enable_peer_access(edx_context* ctx, int peer_device);         /* ctx's GPU may write peer_device's memory (NVLink) */
int edx_device_alloc(edx_context* ctx, size_t bytes, void** out);      /* device memory on ctx's GPU, e.g. the root's frame store */
int edx_device_free(edx_context* ctx, void* p);
int edx_read_device(edx_context* ctx, void* host_dst, const void* device_src, size_t bytes);   /* stream-ordered D2H on ctx's stream, then synchronise */
/* Sort-first split of ONE frame over several contexts / GPUs: this context rasterises, resolves
 * and writes only the 64x64-pixel bins b (row-major) with b % parts == part; the rest of its frame buffer is left
 * untouched. Every context still runs the geometry stages on the whole mesh. parts = 1 restores the full frame. */
int edx_set_screen_partition(edx_context* ctx, int part, int parts);
/* use an existing cudaStream_t (e.g. torch's current stream) instead of the context's own */
int edx_set_stream(edx_context* ctx, void* cuda_stream);
/* CUDA-event timing on the context's stream: begin, N x render, end -> elapsed ms (synchronises) */
int edx_timer_begin(edx_context* ctx);
int edx_timer_end(edx_context* ctx, float* elapsed_ms);
/* per-stage CUDA events inside RenderMesh (adds event records; off by default) */
int edx_set_profiling(edx_context* ctx, int enabled);
int edx_get_stats(edx_context* ctx, edx_stats* out);
/* tuning knobs: "small_max" / "small_max_clip" (largest pixel-centre box side rasterised directly by the
 * geometry / clip kernels, defaults 32 / 8),
 * "hiz" (hierarchical-Z culling of the tile path, default 1), "cluster_cull" (frustum-cull 256-triangle
 * clusters: 0 off, 1 = only for meshes whose triangle order is spatially coherent (default), 2 always),
 * "pdl" (programmatic dependent launch, default 1), "fuse_clip" (clip single-plane straddlers inside the geometry
 * kernel, default 0), "clip_carveout" (shared-memory carve-out the clipper asks for: 0 = follow the tile path's load
 * (default), 1 = prefer L1, 2 = prefer shared memory), "lean_resolve" (shared-memory-free resolve
 * kernel ahead of the tile kernel: 0 never (default), 1 when the last vetted frame had an empty tile path, 2 always).
 * None of them changes a pixel. */
int edx_set_option(edx_context* ctx, const char* name, int value);
/* diagnostics: CTAs of the tile kernel's shape (512 threads, 97 KB shared memory) an SM holds at once (expected 2) */
int edx_debug_tile_residency(edx_context* ctx, int* ctas_per_sm);
/* number of kernel launches issued by the last RenderMesh (for bench.py's gpu_launches) */
int edx_last_launch_count(const edx_context* ctx);
/* their names, comma-separated, in launch order (valid until the next RenderMesh) */
const char* edx_last_launch_list(const edx_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* EDXRASTER_C_H */
