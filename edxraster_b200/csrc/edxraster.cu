// edxraster.cu — C ABI (include/edxraster_c.h) over the sm_100a kernels in edx_kernels.cuh.
//
// Host-side frame orchestration: the counterpart of Renderer::RenderMesh (Core/Renderer.cpp:100-118)
// and of the Renderer's buffer ownership (Renderer.h:18-31). No CPU fallback exists: without a
// CUDA device every entry point reports an error.
#include "../../include/edxraster_c.h"
#include "edx_host_math.h"
#include "edx_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace edx;

struct edx_mesh {
    float4* pos4 = nullptr;
    float4* nrm4 = nullptr;
    uint32_t* i0 = nullptr; uint32_t* i1 = nullptr; uint32_t* i2 = nullptr;
    // Mesh::mTextures + per-triangle slot (Utils/Mesh.h:23,54-59); nTex == 0: the context's constant albedo
    TexDesc* texDesc = nullptr; uchar4* texels = nullptr; uint32_t* texIds = nullptr; uint32_t nTex = 0;
    float4* clusterBox = nullptr;                          // 2 x float4 per 256-triangle cluster
    float4* vclusterBox = nullptr;                         // 2 x float4 per 256-vertex cluster (list front end)
    void* staging = nullptr; size_t stagingBytes = 0;     // device-side landing area for the AoS upload
    uint32_t nVerts = 0, nTris = 0, capVerts = 0, capTris = 0;
    bool coherent = false;                                 // triangle order is spatially coherent: cluster culling pays
    int device = 0;                                        // so the mesh can be released without its context
};

struct edx_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = true;
    uint32_t width = 0, height = 0, binsX = 0, binsY = 0, keyStride = 0;
    bool initialized = false;

    // Renderer::SetTransform state (RenderStates.h:15-19)
    edx_host::Mat4 mv, mvInv, proj, mvp, raster;
    float eye[3], light[3], albedo[3];
    int shader = EDX_SHADER_LAMBERT_ALBEDO;   // the reference installs LambertianAlbedoPixelShader (Renderer.cpp:41)
    int msaaLog2 = 0, texFilter = 2, hierarchical = 1, captureIds = 0, profiling = 0;
    int smallMax = 8, smallMaxClip = 8, hiz = 1, fuseClip = 0, pdl = 1, clusterCull = 1, part = 0, parts = 1;
    MidRec* mid = nullptr; uint32_t midCap = 0;
    int midCtasPerSm = 16;
    uint32_t hintTris = 0, hintVerts = 0;    // size of the previously enqueued mesh: the counters its frame published hint what a frame of an equal-sized mesh needs
    int skipIdle = 1;                        // edx_set_option("skip_idle", 0 | 1): leave out kernels the previous frame of the same mesh had no work for
    int midMax = 64;                         // boxes from smallMax up to this go to mid_kernel (one warp per triangle); 0 = none
    int frontEnd = -1;                       // -1 auto, 0 geom_kernel, 1 cull + list, 2 cull + per-vertex stage + list (FrameParams::frontEnd)
    int frontEndUsed = 0;
    uint32_t* workList = nullptr; uint32_t workListCap = 0;
    uint32_t* vcFlag = nullptr; uint32_t vcFlagCap = 0;
    int4* vrec = nullptr; uint32_t vrecCap = 0;
    int leanResolve = 0;                     // 0 never (default: measured slower with frames in flight), 1 when the last vetted frame had an empty tile path, 2 always
    int clipCarveout = 1;                    // 1 default carve-out (large L1; the clipper's polygons live in local memory), 2 prefer shared memory
    bool colorDirty = false;

    unsigned long long* keys = nullptr;
    uchar4* color = nullptr; float* depth = nullptr; uint32_t* ids = nullptr;
    uchar4* extColor = nullptr; float* extDepth = nullptr;      // caller-owned render targets (optional)
    void* sinkColor = nullptr; void* sinkDepth = nullptr;       // edx_set_frame_sink: where the copy engine pushes each finished frame
    uint32_t* sinkSignal = nullptr; uint32_t sinkSerial = 0;    // edx_set_frame_sink_signal: word that receives the count of frames pushed so far
    // The pushes run on a stream of their own, behind the frame that produced the buffers; the NEXT frame's geometry and
    // clipping proceed meanwhile and only its final pass, which overwrites the buffers, waits for the push (evPushDone).
    cudaStream_t sinkStream = nullptr; cudaEvent_t evFrameDone = nullptr, evPushDone = nullptr; bool pushPending = false;
    BigRec* big = nullptr; uint32_t bigCap = 0;
    uint32_t* bigBox = nullptr; uint32_t bigBoxCap = 0;
    uint32_t* bigOrder = nullptr; uint32_t* bigKey = nullptr; uint32_t* bigBoxSorted = nullptr; uint32_t* bigBound = nullptr; uint32_t bigSortCap = 0;   // nearest-first view (sort_big_kernel)
    int sortBig = 1;                         // edx_set_option("sort_big", 0 | 1)
    int skipTile = 1;                        // edx_set_option("skip_tile", 0 | 1): frames without large triangles end in lean_resolve_kernel instead of tile_kernel
    bool midShrunk = false; int midAuto = 1;  // edx_set_option("mid_auto", 0 | 1): see enqueue_frame
    int binMin = 16384;                      // edx_set_option("bin_min", n): tile-path lists at least this long get per-bin lists (0 = never)
    uint32_t* binCursor = nullptr; uint32_t* binList = nullptr; uint32_t* binKey = nullptr; uint32_t binCursorCap = 0, binListCap = 0, binKeyCap = 0;
    ClipItem* clipQueue = nullptr; uint32_t clipQueueCap = 0;
    ClipRec* clipRecs = nullptr; uint32_t clipRecCap = 0;
    Counters* counters = nullptr;
    Counters* hostCounters = nullptr;        // pinned + device-mapped
    Counters* hostCountersDev = nullptr;     // device view of the same memory
    uint8_t* hostColor = nullptr;            // pinned mirror behind GetBackBuffer
    size_t hostColorBytes = 0;

    uint32_t* clipSlot = nullptr; uint32_t clipSlotCap = 0;   // per triangle: first ClipRec of its fan (this frame)
    uint32_t seenOverFrames = 0;             // Counters::overFrames already accounted for
    // What the pending frame was submitted with. A caller may change transform, shader, targets ... before the next
    // synchronising call; if that call has to re-run the frame (queue overflow) it must be the frame as submitted.
    struct Submitted {
        edx_host::Mat4 mvp, raster; float eye[3], light[3], albedo[3];
        int shader, texFilter, hierarchical, captureIds, part, parts;
        uchar4* extColor; float* extDepth; void* sinkColor; void* sinkDepth; uint32_t* sinkSignal; uint32_t sinkSerial;
    } submitted;
    const edx_mesh* lastMesh = nullptr;
    bool framePending = false;
    int launches = 0;
    std::string launchList;                  // kernels of the last frame, in launch order
    // The frame's kernels as a CUDA graph (one per context, rebuilt when the launch sequence changes shape): a frame is
    // five to eight launches, ~3.5 us of host time each - more than a small frame takes on the GPU. Replaying the
    // captured sequence with fresh kernel parameters costs one launch.
    struct LaunchDesc { const void* func; dim3 grid, block; uint32_t smem; int which; int stage; const char* name; };
    std::vector<LaunchDesc> seq, graphSeq;
    std::vector<cudaGraphNode_t> graphNodes;
    std::vector<cudaKernelNodeParams> graphNodeParams;
    cudaGraph_t graph = nullptr; cudaGraphExec_t graphExec = nullptr;
    int graphPdl = -1, graphCarveGen = -1;
    int useGraphs = 1;                       // edx_set_option("graphs", 0 never | 1 small meshes (launch-bound frames) | 2 always)
    edx_stats stats;
    cudaEvent_t evTimer[2] = { nullptr, nullptr };
    cudaEvent_t evStage[4] = { nullptr, nullptr, nullptr, nullptr };
    std::string error;
};

namespace {

int g_carveGeneration[64];       // bumped whenever clip_kernel's carve-out preference changes: captured graphs keep the old one

int fail(edx_context* c, int code, const std::string& msg)
{
    if (c) c->error = msg;
    return code;
}

#define EDX_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? EDX_ERR_OOM : EDX_ERR_CUDA,           \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                         \
    } while (0)

template <typename T> void dev_free(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

int bind(edx_context* c) { EDX_CUDA(c, cudaSetDevice(c->device)); return EDX_OK; }

void release_frame_buffers(edx_context* c)
{
    dev_free(c->keys); dev_free(c->color); dev_free(c->depth); dev_free(c->ids);
    if (c->hostColor) { cudaFreeHost(c->hostColor); c->hostColor = nullptr; }
}

int allocate_frame_buffers(edx_context* c, uint32_t w, uint32_t h)
{
    release_frame_buffers(c);
    c->width = w; c->height = h;
    c->binsX = (w + BIN - 1) / BIN;                    // cf. Renderer.cpp:26-27 (32-px tiles there)
    c->binsY = (h + BIN - 1) / BIN;
    // one key / depth / id plane per sample (FrameBuffer::Init, FrameBuffer.cpp:12-28); colour is the resolved buffer
    const size_t S = (size_t)1 << c->msaaLog2;
    c->keyStride = (uint32_t)((size_t)c->binsX * c->binsY * KEYS_PER_BIN);
    const size_t nKeys = (size_t)c->keyStride * S;
    const size_t nPix = (size_t)w * h;
    EDX_CUDA(c, cudaMalloc(&c->keys, nKeys * sizeof(unsigned long long)));
    EDX_CUDA(c, cudaMalloc(&c->color, nPix * sizeof(uchar4)));
    EDX_CUDA(c, cudaMalloc(&c->depth, nPix * S * sizeof(float)));
    EDX_CUDA(c, cudaMalloc(&c->ids, nPix * S * sizeof(uint32_t)));
    c->hostColorBytes = nPix * 4;
    EDX_CUDA(c, cudaMallocHost(&c->hostColor, c->hostColorBytes));
    const size_t pairs = nKeys / 2;
    fill_keys_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<ulonglong2*>(c->keys), pairs);
    EDX_CUDA(c, cudaGetLastError());
    EDX_CUDA(c, cudaMemsetAsync(c->color, 0, nPix * sizeof(uchar4), c->stream));     // FrameBuffer.cpp:91-95
    EDX_CUDA(c, cudaMemsetAsync(c->ids, 0xFF, nPix * S * sizeof(uint32_t), c->stream));
    {
        // depth reads before the first frame see the reference's clear value 1.0 (FrameBuffer::Init / Clear, FrameBuffer.cpp:103)
        const size_t n = nPix * S;
        fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->depth, n, 1.0f);
        EDX_CUDA(c, cudaGetLastError());
    }
    c->colorDirty = false;
    return EDX_OK;
}

template <typename T> int grow(edx_context* c, T*& p, uint32_t& cap, uint64_t need)
{
    if (need <= cap && p) return EDX_OK;
    uint64_t want = std::max<uint64_t>(need + need / 4 + 1024, 4096);
    if (want > 0xFFFFFFF0ull) return fail(c, EDX_ERR_OVERFLOW, "queue would exceed 2^32 entries");
    dev_free(p);
    EDX_CUDA(c, cudaMalloc(&p, (size_t)want * sizeof(T)));
    cap = (uint32_t)want;
    return EDX_OK;
}

void fill_params(const edx_context* c, const edx_mesh* m, FrameParams& P)
{
    memset(&P, 0, sizeof(P));
    memcpy(P.mvp, c->mvp.m, 64);
    memcpy(P.raster, c->raster.m, 64);
    memcpy(P.eye, c->eye, 12); memcpy(P.light, c->light, 12); memcpy(P.albedo, c->albedo, 12);
    P.width = (int)c->width; P.height = (int)c->height; P.binsX = (int)c->binsX; P.binsY = (int)c->binsY;
    P.shader = c->shader; P.smallMax = c->smallMax; P.smallMaxClip = c->smallMaxClip; P.hiz = c->hiz; P.hierarchical = c->hierarchical;
    P.captureIds = c->captureIds; P.dump = 0; P.midMax = c->midMax; P.mid = c->mid; P.midCap = c->midCap;
    P.part = c->part; P.parts = c->parts;
    P.fuseClip = c->fuseClip; P.clusterCull = (c->clusterCull == 1 && m->coherent) || c->clusterCull == 2;
    P.msLevel = c->msaaLog2; P.samples = 1 << c->msaaLog2; P.keyStride = c->keyStride;
    const float* Rm = c->raster.m;
    P.rasterAffineXY = (Rm[2] == 0.0f && Rm[6] == 0.0f && Rm[12] == 0.0f && Rm[13] == 0.0f && Rm[14] == 0.0f && Rm[15] == 1.0f) ? 1 : 0;
    P.pos4 = m->pos4; P.nrm4 = m->nrm4; P.i0 = m->i0; P.i1 = m->i1; P.i2 = m->i2; P.clusterBox = m->clusterBox;
    P.nTris = m->nTris; P.nVerts = m->nVerts;
    P.vclusterBox = m->vclusterBox; P.nTriClusters = (m->nTris + 255) / 256; P.nVertClusters = (m->nVerts + 255) / 256;
    P.frontEnd = 0; P.workList = c->workList; P.vcFlag = c->vcFlag; P.vrec = nullptr;
    P.tex = m->texDesc; P.texels = m->texels; P.texIds = m->texIds; P.nTex = m->nTex; P.texFilter = c->texFilter;
    P.keys = c->keys;
    P.big = c->big; P.bigCap = c->bigCap; P.bigBox = c->bigBox;
    P.bigOrder = c->bigOrder; P.bigKey = c->bigKey; P.bigBoxSorted = c->bigBoxSorted; P.bigBound = c->bigBound;
    P.binCursor = c->binCursor; P.binList = c->binList; P.binKey = c->binKey; P.binListCap = c->binListCap; P.binMin = c->binMin < 0 ? -c->binMin : c->binMin; P.binForce = c->binMin < 0 ? 1 : 0;
    P.clipQueue = c->clipQueue; P.clipQueueCap = c->clipQueueCap;
    P.clipRecs = c->clipRecs; P.clipRecCap = c->clipRecCap;
    P.clipSlot = c->clipSlot;
    P.counters = c->counters; P.hostCounters = c->hostCountersDev;
    P.color = c->extColor ? c->extColor : c->color; P.depth = c->extDepth ? c->extDepth : c->depth; P.ids = c->ids;
}

// One frame on the stream. dumpBuf != nullptr routes post-setup triangles to the debug dump as well.
int enqueue_frame(edx_context* c, const edx_mesh* m, DumpRec* dumpBuf, uint32_t dumpCap)
{
    // initial queue sizes; grown on demand after a frame reports it needed more
    if (int r = grow(c, c->big, c->bigCap, std::max<uint64_t>(1u << 16, m->nTris / 8))) return r;
    if (int r = grow(c, c->bigBox, c->bigBoxCap, c->bigCap)) return r;
    if (!c->bigOrder) {                      // the sorted view never holds more than SORT_MAX entries
        uint32_t cap = 0;
        if (int r = grow(c, c->bigOrder, cap, SORT_MAX)) return r;
        cap = 0; if (int r = grow(c, c->bigKey, cap, SORT_MAX)) return r;
        cap = 0; if (int r = grow(c, c->bigBoxSorted, cap, SORT_MAX)) return r;
        cap = 0; if (int r = grow(c, c->bigBound, cap, SORT_MAX)) return r;
        c->bigSortCap = cap;
    }
    if (int r = grow(c, c->mid, c->midCap, std::max<uint64_t>(1u << 16, m->nTris / 8))) return r;
    if (int r = grow(c, c->clipQueue, c->clipQueueCap, std::max<uint64_t>(1u << 14, m->nTris / 32))) return r;
    if (int r = grow(c, c->clipRecs, c->clipRecCap, std::max<uint64_t>(1u << 16, m->nTris / 8))) return r;
    {
        const uint32_t before = c->clipSlotCap;
        if (int r = grow(c, c->clipSlot, c->clipSlotCap, m->nTris)) return r;  // in the context, so a mesh is read-only while it renders
        // slots are written for straddlers only; a fresh allocation must not hold addresses (an overflowed frame's
        // resolve may look one up before the frame is re-run)
        if (c->clipSlotCap != before) EDX_CUDA(c, cudaMemsetAsync(c->clipSlot, 0, (size_t)c->clipSlotCap * 4, c->stream));
    }

    // Front end: large meshes cull their clusters once, on the device, into a work list; meshes that share vertices
    // between triangles (at least two corners per vertex) additionally do the per-vertex work once per vertex.
    int fe = c->frontEnd;
    if (fe < 0) {
        const bool large = m->nTris >= (1u << 18);
        const bool shared = 2ull * m->nVerts <= 3ull * m->nTris;
        const bool cull = (c->clusterCull == 1 && m->coherent) || c->clusterCull == 2;
        fe = !large ? 0 : (shared ? 2 : (cull ? 1 : 0));
    }
    if (fe) {
        if (int r = grow(c, c->workList, c->workListCap, (m->nTris + 255) / 256)) return r;
        if (int r = grow(c, c->vcFlag, c->vcFlagCap, (m->nVerts + 255) / 256)) return r;
        if (fe == 2) if (int r = grow(c, c->vrec, c->vrecCap, m->nVerts)) return r;
    }
    c->frontEndUsed = fe;

    // Per-bin lists for a long tile-path list (bin_*_kernel): launched when the previous frame of an equal-sized mesh had
    // such a list (or nothing is known yet and the mesh could produce one); the kernels themselves look at this frame's
    // list and do nothing if it is short. Room for the (triangle, bin) pairs follows the demand the last frame published;
    // a frame that finds too little keeps the shared list and the next one has the room.
    const uint32_t binMin = (uint32_t)(c->binMin < 0 ? -c->binMin : c->binMin);      // (negative: tests force lists that would not pay)
    bool launchBin = binMin > 0 && m->nTris >= binMin && !dumpBuf;
    if (launchBin && c->skipIdle && c->hintTris == m->nTris && c->hintVerts == m->nVerts) {
        const volatile Counters* h = c->hostCounters;
        if (h->frameSerial != 0) launchBin = h->nBig >= binMin;
    }
    if (launchBin) {
        const volatile Counters* h = c->hostCounters;
        const uint64_t pairs = std::max<uint64_t>({ 1u << 20, (uint64_t)h->nBinPairs, 8ull * std::min<uint64_t>(h->nBig, c->bigCap) });
        if (int r = grow(c, c->binList, c->binListCap, std::min<uint64_t>(pairs, 1ull << 30))) return r;
        if (int r = grow(c, c->binCursor, c->binCursorCap, (uint64_t)c->binsX * c->binsY * BIN_LEVELS)) return r;
        if (int r = grow(c, c->binKey, c->binKeyCap, c->bigCap)) return r;
    }

    FrameParams P;
    fill_params(c, m, P);
    P.frontEnd = fe;
    if (fe == 2) P.vrec = c->vrec;
    if (dumpBuf) { P.dump = 1; P.dumpBuf = dumpBuf; P.dumpCap = dumpCap; }

    auto wait_for_push = [&]() -> cudaError_t {       // before anything of this frame writes colour or depth
        if (!c->pushPending) return cudaSuccess;
        c->pushPending = false;
        return cudaStreamWaitEvent(c->stream, c->evPushDone, 0);
    };
    if (c->shader == EDX_SHADER_DEPTH_ONLY && c->colorDirty) {
        EDX_CUDA(c, wait_for_push());
        // depth-only frames never touch colour: restore the cleared state once (FrameBuffer.cpp:91-95)
        EDX_CUDA(c, cudaMemsetAsync(c->extColor ? c->extColor : c->color, 0, (size_t)c->width * c->height * 4, c->stream));
        c->colorDirty = false;
    }
    if (c->shader != EDX_SHADER_DEPTH_ONLY) c->colorDirty = true;

    // ---- the frame's launch sequence (stage: 0 geometry, 1 clip + mid, 2 tile / shade / end) ----
    typedef edx_context::LaunchDesc LaunchDesc;
    std::vector<LaunchDesc>& seq = c->seq;
    seq.clear();
    auto add = [&](const char* name, auto kernel, dim3 grid, dim3 block, size_t smem, int which, int stage) {
        seq.push_back(LaunchDesc{ (const void*)kernel, grid, block, (uint32_t)smem, which, stage, name });
    };
    const bool textured = c->shader == EDX_SHADER_LAMBERT_ALBEDO && m->nTex != 0;
    bool lean = c->msaaLog2 == 0 && (c->leanResolve == 2 || (c->leanResolve == 1 && c->stats.binned_tris == 0));
    P.leanResolve = lean ? 1 : 0;
    // tile_kernel resolves depth (+ owner ids); a shaded single-sample frame's colour is a pass of its own over those
    // ids (shade_kernel), so the tile kernel sees the frame as depth-only with id capture (parameter block T)
    const bool shaded = c->shader != EDX_SHADER_DEPTH_ONLY;
    FrameParams T = P;
    if (shaded && c->msaaLog2 == 0) { T.shader = EDX_SHADER_DEPTH_ONLY; T.captureIds = 1; }
    if (m->nTris && fe == 0) {
        add("geom_kernel", geom_kernel, dim3((m->nTris + 255) / 256), dim3(256), 0, 0, 0);
    } else if (m->nTris) {
        const uint32_t nTC = (m->nTris + 255) / 256, nVC = (m->nVerts + 255) / 256;
        add("cull_kernel", cull_kernel, dim3((nTC + (fe == 2 ? nVC : 0) + 255) / 256), dim3(256), 0, 0, 0);
        if (fe == 2) add("vertex_kernel", vertex_kernel, dim3(std::min(nVC, 148u * 8u)), dim3(256), 0, 0, 0);
        if (fe == 2) add("geom_list_kernel", geom_list_kernel<true>, dim3(std::min(nTC, 148u * 5u)), dim3(256), 0, 0, 0);
        else add("geom_list_kernel", geom_list_kernel<false>, dim3(std::min(nTC, 148u * 5u)), dim3(256), 0, 0, 0);
    }
    // Two kernels of the chain are often idle: mid_kernel (no mid-size triangles: C2) and sort_big_kernel (short or
    // empty tile-path list: C1, C2). An idle kernel still costs its launch and a grid of CTAs that read one counter and
    // exit - 2.5 us per C2 frame. The counters the previous frame of this context published (pinned host memory; no
    // synchronisation, a stale value is as good) say whether a mesh of the same size had work for them; if not they are left out.
    // That is always safe: without mid_kernel a mid-size triangle takes the tile path (and counts itself, so the next
    // frame has the kernel again), without the sort the tile kernel reads the list as appended.
    bool launchMid = c->midMax > 0, launchSort = c->sortBig && c->hiz && c->hierarchical;
    if (c->skipIdle && c->hintTris == m->nTris && c->hintVerts == m->nVerts && !dumpBuf) {
        const volatile Counters* h = c->hostCounters;
        if (h->frameSerial != 0) {
            launchMid = launchMid && (h->nMid != 0 || h->nMidDiverted != 0);
            launchSort = launchSort && h->nBig >= (uint32_t)SORT_MIN;
        }
    }
    // The warp-per-triangle path has no occlusion culling: it tests every pixel of every box. That is the cheapest way
    // through a frame whose mid-size triangles cover the screen a few times (C4: 7x), and the wrong one when they cover it
    // hundreds of times (the M1 stress: 740x) - there the tile path, which culls hierarchically and reads per-bin lists,
    // wins even at its higher cost per triangle. mid_kernel publishes the pixels its boxes spanned; above 128 screens the
    // following frames of an equal-sized mesh keep only boxes below 32 px on the path (until another mesh arrives: the
    // measure itself changes with the routing, so the decision is not revisited every frame).
    if (c->hintTris != m->nTris || c->hintVerts != m->nVerts) c->midShrunk = false;
    else if (c->midAuto && c->hostCounters->frameSerial != 0 &&
             ((const volatile Counters*)c->hostCounters)->midArea > 128ull * c->width * c->height) c->midShrunk = true;
    if (c->midShrunk && c->midMax > 32) { P.midMax = 32; T.midMax = 32; }
    // The tile kernel as well: a frame without large triangles needs it only for the final pass over the keys, and its
    // CTAs own half an SM each (512 threads x 64 registers, 100 KB of shared memory) while they wait on one 32 KB copy -
    // no geometry CTA of another frame in flight fits beside them. If the previous frame of this mesh put nothing on the
    // tile path, lean_resolve_kernel (256-thread CTAs, no shared memory) ends the frame instead. A large triangle that
    // turns up in such a frame is rasterised by mid_kernel - correct at any size - and reported (nBigDiverted), so the
    // next frame has the tile kernel again.
    bool launchTile = true;
    if (c->skipIdle && c->skipTile && c->hintTris == m->nTris && c->hintVerts == m->nVerts && !dumpBuf && c->msaaLog2 == 0 && c->midMax > 0 && c->leanResolve == 0) {
        const volatile Counters* h = c->hostCounters;
        if (h->frameSerial != 0 && h->nBig == 0 && h->nBigDiverted == 0) { launchTile = false; launchMid = true; lean = true; }
    }
    P.leanResolve = lean ? 1 : 0; T.leanResolve = P.leanResolve;
    P.tileLaunched = launchTile ? 1 : 0; T.tileLaunched = P.tileLaunched;
    c->hintTris = m->nTris; c->hintVerts = m->nVerts;
    P.midLaunched = launchMid ? 1 : 0; T.midLaunched = P.midLaunched;
    if (m->nTris) add("clip_kernel", clip_kernel, dim3(148 * 4), dim3(128), 0, 0, 1);
    // one warp per mid-size triangle: as many warps as the chip holds (148 SMs x 16 CTAs x 4 warps) so that a queue of
    // tens of thousands is a couple of triangles per warp, not a serial walk whose every step waits on a record load
    // (when it is only the catch-all of a frame without tile_kernel - the previous frame had nothing for it - one CTA per
    // SM will do: the grid-stride loop is correct at any grid size, and 2,368 CTAs that read a counter and exit cost 1-2 us)
    bool midIdle = false;
    if (launchMid && !launchTile && c->skipIdle && c->hostCounters->frameSerial != 0) {
        const volatile Counters* h = c->hostCounters;
        midIdle = h->nMid == 0 && h->nMidDiverted == 0 && h->nBigDiverted == 0;
    }
    if (m->nTris && launchMid) add("mid_kernel", mid_kernel, dim3(148 * (midIdle ? 1u : (unsigned)c->midCtasPerSm)), dim3(128), 0, 0, 1);
    if (m->nTris && launchSort) add("sort_big_kernel", sort_big_kernel, dim3(1), dim3(1024), 0, 0, 1);
    if (launchBin) {
        add("bin_keys_kernel", bin_keys_kernel, dim3(148 * 4), dim3(256), 0, 0, 1);
        add("bin_fill_kernel", bin_fill_kernel<false>, dim3(148 * 8), dim3(256), 0, 0, 1);
        add("bin_scan_kernel", bin_scan_kernel, dim3(1), dim3(1024), 0, 0, 1);
        add("bin_fill_kernel", bin_fill_kernel<true>, dim3(148 * 8), dim3(256), 0, 0, 1);
    }
    const dim3 leanGrid((c->binsX * c->binsY * 16 + 7) / 8);
    if (c->msaaLog2 == 0) {
        // (a shaded frame without tile_kernel needs no resolve kernel either: shade_kernel<., true> reads the keys itself)
        const bool fusedShade = shaded && !launchTile;
        if (fusedShade) { }
        else if (lean && !shaded && !c->captureIds) add("lean_resolve_kernel", lean_resolve_kernel<true>, leanGrid, dim3(256), 0, 1, 2);
        else if (lean) add("lean_resolve_kernel", lean_resolve_kernel<false>, leanGrid, dim3(256), 0, 1, 2);
        if (launchTile) add("tile_kernel", tile_kernel<false>, dim3(c->binsX * c->binsY), dim3(TILE_THREADS), sizeof(TileShared), 1, 2);
        if (shaded) {
            const uint32_t tiles = ((c->width + TILE_PX - 1) / TILE_PX) * ((c->height + TILE_PX - 1) / TILE_PX);   // one CTA each
            if (fusedShade) {
                if (textured) add("shade_kernel", shade_kernel<true, true>, dim3(tiles), dim3(256), 0, 0, 2);
                else add("shade_kernel", shade_kernel<false, true>, dim3(tiles), dim3(256), 0, 0, 2);
            } else if (textured) add("shade_kernel", shade_kernel<true, false>, dim3(tiles), dim3(256), 0, 0, 2);
            else add("shade_kernel", shade_kernel<false, false>, dim3(tiles), dim3(256), 0, 0, 2);
        }
    } else {
        // one CTA per (bin, sample), then the per-pixel resolve that also ends the frame
        add("tile_kernel", tile_kernel<true>, dim3(c->binsX * c->binsY, 1u << c->msaaLog2), dim3(TILE_THREADS), sizeof(TileShared), 0, 2);
        add("msaa_resolve_kernel", msaa_resolve_kernel, dim3((c->keyStride + 255) / 256), dim3(256), 0, 0, 2);
    }
    add("frame_end_kernel", frame_end_kernel, dim3(1), dim3(32), 0, 0, 2);       // counters -> pinned host memory, reset for the next frame

    c->launches = (int)seq.size();
    c->launchList.clear();
    for (const LaunchDesc& d : seq) { if (!c->launchList.empty()) c->launchList += ","; c->launchList += d.name; }
    FrameParams* blocks[2] = { &P, &T };
    // Frame kernels are launched with programmatic stream serialization (PDL): each may be scheduled while
    // its predecessor drains and blocks in cudaGridDependencySynchronize() until that one has completed.
    auto launch_one = [&](const LaunchDesc& d) -> cudaError_t {
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = c->pdl ? 1 : 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = d.grid; cfg.blockDim = d.block; cfg.dynamicSmemBytes = d.smem; cfg.stream = c->stream;
        cfg.attrs = attr; cfg.numAttrs = 1;
        void* args[1] = { blocks[d.which] };
        return cudaLaunchKernelExC(&cfg, d.func, args);
    };
    // Measured on B200: replay saves ~8 us of host time per frame (C1 with three frames in flight: 25.5 -> 17.1 us)
    // but a replayed frame takes 4-6 % longer on the GPU than the same kernels launched one by one (C2: 46.1 -> 48.5 us),
    // so by default only meshes small enough to be launch-bound go through the graph.
    bool viaGraph = (c->useGraphs == 2 || (c->useGraphs == 1 && m->nTris <= (1u << 18))) && !c->profiling && !dumpBuf;
    if (viaGraph) {
        bool same = c->graphExec && c->graphPdl == c->pdl && c->graphCarveGen == g_carveGeneration[c->device & 63] && c->graphSeq.size() == seq.size();
        for (size_t i = 0; same && i < seq.size(); i++) {
            const LaunchDesc& a = seq[i]; const LaunchDesc& b = c->graphSeq[i];
            same = a.func == b.func && a.grid.x == b.grid.x && a.grid.y == b.grid.y && a.block.x == b.block.x && a.smem == b.smem && a.which == b.which;
        }
        if (!same) {
            // (re)capture: the same launches, recorded instead of executed
            if (c->graphExec) { cudaGraphExecDestroy(c->graphExec); c->graphExec = nullptr; }
            if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
            c->graphNodes.clear(); c->graphNodeParams.clear();
            bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            for (size_t i = 0; ok && i < seq.size(); i++) {
                ok = launch_one(seq[i]) == cudaSuccess;
                cudaStreamCaptureStatus st; const cudaGraphNode_t* deps = nullptr; const cudaGraphEdgeData* edges = nullptr; size_t nDeps = 0;
                ok = ok && cudaStreamGetCaptureInfo_v3(c->stream, &st, nullptr, nullptr, &deps, &edges, &nDeps) == cudaSuccess && nDeps == 1;
                if (ok) c->graphNodes.push_back(deps[0]);
            }
            cudaGraph_t g = nullptr;
            const cudaError_t endErr = cudaStreamEndCapture(c->stream, &g);      // always end the capture, even after a failure
            ok = ok && endErr == cudaSuccess && g;
            if (ok) { c->graph = g; ok = cudaGraphInstantiate(&c->graphExec, c->graph, 0) == cudaSuccess; }
            else if (g) cudaGraphDestroy(g);
            for (size_t i = 0; ok && i < c->graphNodes.size(); i++) {
                cudaKernelNodeParams np;
                ok = cudaGraphKernelNodeGetParams(c->graphNodes[i], &np) == cudaSuccess;
                c->graphNodeParams.push_back(np);
            }
            if (ok) { c->graphSeq = seq; c->graphPdl = c->pdl; c->graphCarveGen = g_carveGeneration[c->device & 63]; }
            else {
                cudaGetLastError();                       // graphs are an optimisation: fall back to plain launches for good
                if (c->graphExec) { cudaGraphExecDestroy(c->graphExec); c->graphExec = nullptr; }
                c->useGraphs = 0; viaGraph = false;
            }
        } else {
            for (size_t i = 0; i < seq.size(); i++) {
                cudaKernelNodeParams np = c->graphNodeParams[i];
                void* args[1] = { blocks[seq[i].which] };
                np.kernelParams = args; np.extra = nullptr;
                EDX_CUDA(c, cudaGraphExecKernelNodeSetParams(c->graphExec, c->graphNodes[i], &np));
            }
        }
        if (viaGraph) { EDX_CUDA(c, wait_for_push()); EDX_CUDA(c, cudaGraphLaunch(c->graphExec, c->stream)); }
    }
    if (!viaGraph) {
        int stage = 0;
        if (c->profiling) EDX_CUDA(c, cudaEventRecord(c->evStage[0], c->stream));
        for (const LaunchDesc& d : seq) {
            while (c->profiling && stage < d.stage) EDX_CUDA(c, cudaEventRecord(c->evStage[++stage], c->stream));
            if (d.stage == 2) EDX_CUDA(c, wait_for_push());      // the final pass overwrites the buffers a previous frame's push may still read
            EDX_CUDA(c, launch_one(d));
        }
        while (c->profiling && stage < 2) EDX_CUDA(c, cudaEventRecord(c->evStage[++stage], c->stream));
    }
    // frame sink (edx_set_frame_sink): the copy engine pushes the finished buffers, e.g. into the root GPU's memory over NVLink
    const bool pushColor = c->sinkColor && c->shader != EDX_SHADER_DEPTH_ONLY && c->msaaLog2 == 0, pushDepth = c->sinkDepth && c->msaaLog2 == 0;
    cudaStream_t tail = c->stream;                    // where the signal goes: behind the pushes, or behind the frame
    if (pushColor || pushDepth) {
        if (!c->sinkStream) {
            EDX_CUDA(c, cudaStreamCreateWithFlags(&c->sinkStream, cudaStreamNonBlocking));
            EDX_CUDA(c, cudaEventCreateWithFlags(&c->evFrameDone, cudaEventDisableTiming));
            EDX_CUDA(c, cudaEventCreateWithFlags(&c->evPushDone, cudaEventDisableTiming));
        }
        EDX_CUDA(c, cudaEventRecord(c->evFrameDone, c->stream));
        EDX_CUDA(c, cudaStreamWaitEvent(c->sinkStream, c->evFrameDone, 0));
        if (pushColor) EDX_CUDA(c, cudaMemcpyAsync(c->sinkColor, P.color, (size_t)c->width * c->height * 4, cudaMemcpyDeviceToDevice, c->sinkStream));
        if (pushDepth) EDX_CUDA(c, cudaMemcpyAsync(c->sinkDepth, P.depth, (size_t)c->width * c->height * 4, cudaMemcpyDeviceToDevice, c->sinkStream));
        tail = c->sinkStream;
    }
    // ... and a one-thread kernel behind the pushes publishes how many frames this context has finished and pushed
    // (system-scope store into the same GPU's or a peer's memory): what a consumer on the root GPU polls instead of a
    // collective. (No sink needed: the root's own contexts render straight into the store, edx_set_render_target.)
    if (c->sinkSignal && c->msaaLog2 == 0) {
        sink_signal_kernel<<<1, 1, 0, tail>>>(c->sinkSignal, c->sinkSerial);
        c->launches++; c->launchList += ",sink_signal_kernel";
    }
    if (pushColor || pushDepth) { EDX_CUDA(c, cudaEventRecord(c->evPushDone, c->sinkStream)); c->pushPending = true; }
    if (c->profiling) EDX_CUDA(c, cudaEventRecord(c->evStage[3], c->stream));
    EDX_CUDA(c, cudaGetLastError());
    return EDX_OK;
}

// Which shared-memory carve-out clip_kernel asks for (`edx_set_option("clip_carveout", 1 | 2)`, an experiment knob).
// Round 1 found that C3 ran 0.8 or 1.65 ms per frame depending on this preference and followed the tile path's load
// with a heuristic. The cause is now known: the preference changes how fast the clipper runs -
// its polygons live in local memory, so it is 2.4x faster with the large L1 - and thereby the ORDER in which it
// appends the large triangles; the tile kernel's hierarchical Z used to cull progressively, in list order, and on
// the bins that no single triangle covers a bad order left thousands of survivors. The bound is now complete before
// anything is admitted and long lists are walked nearest first, so the order no longer matters and the clipper
// simply keeps the default (large L1) configuration.
void tune_clip_carveout(edx_context* c, uint32_t)
{
    static int current[64];                                    // function attributes are per device (0 = not set yet)
    const int want = c->clipCarveout == 2 ? 2 : 1;
    int& cur = current[c->device & 63];
    if (want == cur) return;
    cudaFuncSetAttribute(clip_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         want == 2 ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
    cur = want;
    g_carveGeneration[c->device & 63]++;
}

// Wait for the pending frame; if a queue overflowed, grow it and run the frame again. Frames submitted earlier
// without a synchronising call in between cannot be run again (their transform and target are gone): if the
// device counted more overflowed frames than the ones repaired here, the queues are grown to the largest demand
// seen and the call reports EDX_ERR_OVERFLOW once, so the caller knows those frames are incomplete.
void save_submitted(edx_context* c, edx_context::Submitted& s)
{
    s.mvp = c->mvp; s.raster = c->raster;
    memcpy(s.eye, c->eye, 12); memcpy(s.light, c->light, 12); memcpy(s.albedo, c->albedo, 12);
    s.shader = c->shader; s.texFilter = c->texFilter; s.hierarchical = c->hierarchical; s.captureIds = c->captureIds;
    s.part = c->part; s.parts = c->parts;
    s.extColor = c->extColor; s.extDepth = c->extDepth; s.sinkColor = c->sinkColor; s.sinkDepth = c->sinkDepth; s.sinkSignal = c->sinkSignal; s.sinkSerial = c->sinkSerial;
}

void load_submitted(edx_context* c, const edx_context::Submitted& s)
{
    c->mvp = s.mvp; c->raster = s.raster;
    memcpy(c->eye, s.eye, 12); memcpy(c->light, s.light, 12); memcpy(c->albedo, s.albedo, 12);
    c->shader = s.shader; c->texFilter = s.texFilter; c->hierarchical = s.hierarchical; c->captureIds = s.captureIds;
    c->part = s.part; c->parts = s.parts;
    c->extColor = s.extColor; c->extDepth = s.extDepth; c->sinkColor = s.sinkColor; c->sinkDepth = s.sinkDepth; c->sinkSignal = s.sinkSignal; c->sinkSerial = s.sinkSerial;
}

int finish_frame(edx_context* c)
{
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->sinkStream) EDX_CUDA(c, cudaStreamSynchronize(c->sinkStream));      // a synchronised frame has also been pushed
    if (!c->framePending) return EDX_OK;
    uint32_t repaired = 0;
    for (int attempt = 0; attempt < 8; attempt++) {
        const Counters k = *c->hostCounters;
        const bool over = k.nBig > c->bigCap || k.nClipQueue > c->clipQueueCap || k.nClipRecs > c->clipRecCap || k.nMid > c->midCap;
        c->stats.binned_tris = k.nBig; c->stats.clipped_tris = k.nClipQueue; c->stats.clip_records = k.nClipRecs;
        c->stats.tile_pairs = k.tilePairs; c->stats.mid_tris = k.nMid; c->stats.bin_pairs = k.binned ? k.nBinPairs : 0;
#ifdef EDX_DEBUG_STATS
        if (getenv("EDX_DEBUG_PRINT")) {
            uint32_t h[512];
            if (cudaMemcpyFromSymbol(h, g_tileResident, sizeof(h)) == cudaSuccess) {
                int hist[4] = { 0, 0, 0, 0 };
                for (int i = 0; i < 256; i++) hist[std::min(3u, h[256 + i])]++;
                fprintf(stderr, "[edx dbg] tile_kernel CTAs resident per SM (max seen): %d SMs x1, %d SMs x2, %d SMs x3+\n", hist[1], hist[2], hist[3]);
                memset(h, 0, sizeof(h));
                cudaMemcpyToSymbol(g_tileResident, h, sizeof(h));
            }
        }
        if (getenv("EDX_DEBUG_PRINT")) {
            unsigned long long cd[4];
            if (cudaMemcpyFromSymbol(cd, g_clipDbg, sizeof(cd)) == cudaSuccess) {
                fprintf(stderr, "[edx dbg] clip_kernel, longest warp (cycles): single-plane loop %llu, multi-plane loop %llu (of it, until the polygons are clipped: %llu), body %llu\n", cd[0], cd[1], cd[3], cd[2]);
                memset(cd, 0, sizeof(cd));
                cudaMemcpyToSymbol(g_clipDbg, cd, sizeof(cd));
            }
        }
        if (getenv("EDX_DEBUG_PRINT") && k.dbg[6]) {
            static unsigned long long hb[8192][10];
            if (cudaMemcpyFromSymbol(hb, g_binDbg, sizeof(hb)) == cudaSuccess) {
                const int nb = std::min<int>(8192, c->binsX * c->binsY);
                std::vector<int> order(nb);
                for (int i = 0; i < nb; i++) order[i] = i;
                auto tot = [&](int i) { return hb[i][0] + hb[i][1] + hb[i][2] + hb[i][3]; };
                std::sort(order.begin(), order.end(), [&](int a, int b) { return tot(a) > tot(b); });
                unsigned long long sum = 0;
                for (int i = 0; i < nb; i++) sum += tot(i);
                fprintf(stderr, "[edx dbg] per-bin cycles: mean %llu; heaviest bins (bin: cand sweep raster resolve | survivors)\n", sum / std::max(nb, 1));
                for (int i = 0; i < std::min(nb, 10); i++) {
                    const int b = order[i];
                    fprintf(stderr, "[edx dbg]   bin %4d (%2d,%2d): %7llu %7llu %7llu %7llu | %llu | iterations/flushes/candidates %llu cycles until: classified %llu counted %llu slabs done %llu (flushes in them %llu); flush at loop top %llu\n", b, b % (int)c->binsX, b / (int)c->binsX, hb[b][0], hb[b][1], hb[b][2], hb[b][3], hb[b][5], hb[b][4] & 0xFFFFFFFFull, hb[b][4] >> 32, hb[b][6], hb[b][7], hb[b][9], hb[b][8]);
                }
                memset(hb, 0, sizeof(hb));
                cudaMemcpyToSymbol(g_binDbg, hb, sizeof(hb));
            }
        }
        if (getenv("EDX_DEBUG_PRINT") && k.dbg[6])
            fprintf(stderr, "[edx dbg] bins=%llu  mean cycles/bin: cand=%llu sweep=%llu flush=%llu final=%llu resolve=%llu | flushes=%llu survivors=%llu\n",
                    k.dbg[6], k.dbg[0] / k.dbg[6], k.dbg[1] / k.dbg[6], k.dbg[2] / k.dbg[6], k.dbg[3] / k.dbg[6], k.dbg[7] / k.dbg[6], k.dbg[4], k.dbg[5]);
#endif
        if (!over) {
            if (c->profiling) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, c->evStage[0], c->evStage[1]); c->stats.stage_ms[0] = ms;
                cudaEventElapsedTime(&ms, c->evStage[1], c->evStage[2]); c->stats.stage_ms[1] = ms;
                cudaEventElapsedTime(&ms, c->evStage[2], c->evStage[3]); c->stats.stage_ms[2] = ms;
                cudaEventElapsedTime(&ms, c->evStage[0], c->evStage[3]); c->stats.stage_ms[3] = ms;
            }
            c->framePending = false;
            tune_clip_carveout(c, k.tilePairs);
            const uint32_t lost = k.overFrames - c->seenOverFrames - repaired;
            c->seenOverFrames = k.overFrames;
            if (lost) {
                if (int r = grow(c, c->big, c->bigCap, (uint64_t)k.maxBig + (k.maxClipQueue > c->clipQueueCap ? 7ull * k.maxClipQueue : 0))) return r;
                if (int r = grow(c, c->bigBox, c->bigBoxCap, c->bigCap)) return r;
                if (int r = grow(c, c->mid, c->midCap, (uint64_t)k.maxMid + (k.maxClipQueue > c->clipQueueCap ? 7ull * k.maxClipQueue : 0))) return r;
                if (int r = grow(c, c->clipRecs, c->clipRecCap, std::max<uint64_t>(k.maxClipRecs, 7ull * std::min<uint64_t>(k.maxClipQueue, c->clipQueueCap)))) return r;
                if (int r = grow(c, c->clipQueue, c->clipQueueCap, k.maxClipQueue)) return r;
                c->stats.regrow_count++;
                return fail(c, EDX_ERR_OVERFLOW, std::to_string(lost) + " frame(s) submitted before the last one overflowed the internal queues and are "
                            "incomplete (no synchronising call followed them, so they could not be re-run); the queues have been grown: render them again");
            }
            return EDX_OK;
        }
        repaired++;
        // a clip-queue overflow hides fan triangles, so size the dependent queues generously too
        if (int r = grow(c, c->big, c->bigCap, (uint64_t)k.nBig + (k.nClipQueue > c->clipQueueCap ? 7ull * k.nClipQueue : 0))) return r;
        if (int r = grow(c, c->bigBox, c->bigBoxCap, c->bigCap)) return r;
        if (int r = grow(c, c->mid, c->midCap, (uint64_t)k.nMid + (k.nClipQueue > c->clipQueueCap ? 7ull * k.nClipQueue : 0))) return r;
        if (int r = grow(c, c->clipQueue, c->clipQueueCap, k.nClipQueue)) return r;
        if (int r = grow(c, c->clipRecs, c->clipRecCap, std::max<uint64_t>(k.nClipRecs, 7ull * std::min<uint64_t>(k.nClipQueue, c->clipQueueCap)))) return r;
        c->stats.regrow_count++;
        {
            // re-run the frame AS SUBMITTED: the caller may have set another transform / shader / target since
            edx_context::Submitted now;
            save_submitted(c, now);
            load_submitted(c, c->submitted);
            const int r = enqueue_frame(c, c->lastMesh, nullptr, 0);
            load_submitted(c, now);
            if (r) return r;
        }
        EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return fail(c, EDX_ERR_OVERFLOW, "internal queues still overflow after 8 regrow attempts");
}

int upload_mesh(edx_context* c, edx_mesh* m, const void* vertices, uint32_t nv, const uint32_t* indices, uint32_t nt)
{
    const size_t vb = (size_t)nv * 32, ib = (size_t)nt * 12;
    const size_t need = std::max(vb, ib);
    if (need > m->stagingBytes) {
        dev_free(m->staging);
        EDX_CUDA(c, cudaMalloc(&m->staging, std::max<size_t>(need, 256)));
        m->stagingBytes = std::max<size_t>(need, 256);
    }
    m->nVerts = nv; m->nTris = nt;
    if (nv) {
        EDX_CUDA(c, cudaMemcpyAsync(m->staging, vertices, vb, cudaMemcpyHostToDevice, c->stream));
        split_vertices_kernel<<<(nv + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(m->staging), m->pos4, m->nrm4, nv);
        vertex_cluster_bounds_kernel<<<(nv + 255) / 256, 256, 0, c->stream>>>(m->pos4, nv, m->vclusterBox);
    }
    if (nt) {
        EDX_CUDA(c, cudaMemcpyAsync(m->staging, indices, ib, cudaMemcpyHostToDevice, c->stream));
        split_indices_kernel<<<(nt + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(m->staging), m->i0, m->i1, m->i2, nt);
        cluster_bounds_kernel<<<(nt + 255) / 256, 256, 0, c->stream>>>(m->pos4, m->i0, m->i1, m->i2, nt, m->clusterBox);
    }
    EDX_CUDA(c, cudaGetLastError());
    return EDX_OK;
}

} // namespace

extern "C" {

const char* edx_version(void) { return "edxraster_b200 0.1.0 (sm_100a)"; }

int edx_create(int device, edx_context** out)
{
    if (!out) return EDX_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return EDX_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return EDX_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return EDX_ERR_CUDA;
    if (prop.major != 10) return EDX_ERR_NO_DEVICE;           // kernels are built for sm_100a only
    edx_context* c = new edx_context;
    c->device = device;
    memset(&c->stats, 0, sizeof(c->stats));
    c->mv = c->mvInv = c->proj = c->mvp = c->raster = edx_host::identity();
    c->eye[0] = c->eye[1] = c->eye[2] = 0.0f;
    const float l[3] = { 1.0f, 1.0f, -1.0f };                   // Renderer.cpp:290
    edx_host::normalize3(l, c->light);
    c->albedo[0] = c->albedo[1] = c->albedo[2] = 0.9f;          // Mesh.cpp:48,67
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&c->counters, sizeof(Counters)) == cudaSuccess &&
              cudaHostAlloc(&c->hostCounters, sizeof(Counters), cudaHostAllocMapped) == cudaSuccess &&
              cudaHostGetDevicePointer((void**)&c->hostCountersDev, c->hostCounters, 0) == cudaSuccess &&
              cudaMemset(c->counters, 0, sizeof(Counters)) == cudaSuccess &&
              cudaFuncSetAttribute(tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared)) == cudaSuccess &&
              cudaFuncSetAttribute(tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared)) == cudaSuccess &&
              cudaFuncSetAttribute(tile_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
              cudaFuncSetAttribute(tile_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) == cudaSuccess;
    for (int i = 0; ok && i < 2; i++) ok = cudaEventCreate(&c->evTimer[i]) == cudaSuccess;
    for (int i = 0; ok && i < 4; i++) ok = cudaEventCreate(&c->evStage[i]) == cudaSuccess;
    if (!ok) { edx_destroy(c); return EDX_ERR_CUDA; }
    memset(c->hostCounters, 0, sizeof(Counters));
    *out = c;
    return EDX_OK;
}

void edx_destroy(edx_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    release_frame_buffers(c);
    dev_free(c->big); dev_free(c->bigBox); dev_free(c->bigOrder); dev_free(c->bigKey); dev_free(c->bigBoxSorted); dev_free(c->bigBound); dev_free(c->binCursor); dev_free(c->binList); dev_free(c->binKey); dev_free(c->clipQueue); dev_free(c->clipRecs); dev_free(c->clipSlot); dev_free(c->counters);
    dev_free(c->workList); dev_free(c->vcFlag); dev_free(c->vrec); dev_free(c->mid);
    if (c->graphExec) cudaGraphExecDestroy(c->graphExec);
    if (c->graph) cudaGraphDestroy(c->graph);
    if (c->hostCounters) cudaFreeHost(c->hostCounters);
    for (auto& e : c->evTimer) if (e) cudaEventDestroy(e);
    for (auto& e : c->evStage) if (e) cudaEventDestroy(e);
    if (c->evFrameDone) cudaEventDestroy(c->evFrameDone);
    if (c->evPushDone) cudaEventDestroy(c->evPushDone);
    if (c->sinkStream) cudaStreamDestroy(c->sinkStream);
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* edx_last_error(const edx_context* c) { return c ? c->error.c_str() : "null context"; }

int edx_initialize(edx_context* c, uint32_t width, uint32_t height)
{
    if (!c || width == 0 || height == 0 || width > 16384 || height > 16384) return fail(c, EDX_ERR_INVALID, "bad size");
    // 28.4 edge functions are int32: (16 W)(16 H) must stay below 2^31
    if ((uint64_t)(width * 16ull) * (uint64_t)(height * 16ull) >= (1ull << 31))
        return fail(c, EDX_ERR_UNSUPPORTED, "resolution exceeds the int32 range of the 28.4 edge functions (16W * 16H must stay below 2^31)");
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    // RenderStates::DefaultSettings (RenderStates.h:55-61)
    c->msaaLog2 = 0; c->hierarchical = 1; c->texFilter = 2;
    c->framePending = false;
    if (int r = allocate_frame_buffers(c, width, height)) return r;
    c->initialized = true;
    return EDX_OK;
}

int edx_resize(edx_context* c, uint32_t width, uint32_t height)
{
    if (!c || !c->initialized) return fail(c, EDX_ERR_INVALID, "not initialised");
    if (width == 0 || height == 0 || width > 16384 || height > 16384) return fail(c, EDX_ERR_INVALID, "bad size");
    if ((uint64_t)(width * 16ull) * (uint64_t)(height * 16ull) >= (1ull << 31))
        return fail(c, EDX_ERR_UNSUPPORTED, "resolution exceeds the int32 range of the 28.4 edge functions (16W * 16H must stay below 2^31)");
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    c->framePending = false;
    return allocate_frame_buffers(c, width, height);
}

int edx_set_transform(edx_context* c, const float mv[16], const float proj[16], const float raster[16])
{
    if (!c || !mv || !proj || !raster) return fail(c, EDX_ERR_INVALID, "null matrix");
    memcpy(c->mv.m, mv, 64); memcpy(c->proj.m, proj, 64); memcpy(c->raster.m, raster, 64);
    c->mvInv = edx_host::inverse(c->mv);                    // Renderer.cpp:88
    c->mvp = edx_host::multiply(c->proj, c->mv);            // Renderer.cpp:90
    const float zero[3] = { 0.0f, 0.0f, 0.0f };
    edx_host::transform_point3(c->mvInv, zero, c->eye);     // Renderer.cpp:289
    return EDX_OK;
}

int edx_set_msaa_mode(edx_context* c, int log2)
{
    if (!c) return EDX_ERR_INVALID;
    if (log2 < 0 || log2 > 5) return fail(c, EDX_ERR_INVALID, "sample_count_log2 must be 0..5 (FrameBuffer.cpp:107-191 has tables up to 32x)");
    if (log2 == c->msaaLog2) return EDX_OK;          // the viewer calls this every frame (Main.cpp:97)
    if (c->extColor || c->extDepth || c->sinkColor || c->sinkDepth) {
        if (log2 != 0) return fail(c, EDX_ERR_UNSUPPORTED, "caller-owned render targets and frame sinks are single-sample only");
    }
    c->msaaLog2 = log2;
    if (!c->initialized) return EDX_OK;
    // Renderer::SetMSAAMode re-creates the frame buffer (Renderer.cpp:94-98 -> Resize)
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    c->framePending = false;
    return allocate_frame_buffers(c, c->width, c->height);
}

int edx_set_texture_filter(edx_context* c, int filter)
{
    if (!c || filter < 0 || filter > 5) return fail(c, EDX_ERR_INVALID, "filter out of range");
    c->texFilter = filter;
    return EDX_OK;
}

int edx_set_hierarchical_rasterize(edx_context* c, int enabled)
{
    if (!c) return EDX_ERR_INVALID;
    c->hierarchical = enabled ? 1 : 0;
    return EDX_OK;
}

int edx_set_pixel_shader(edx_context* c, int shader)
{
    if (!c || shader < 0 || shader > 3) return fail(c, EDX_ERR_INVALID, "unknown shader");
    c->shader = shader;
    return EDX_OK;
}

int edx_set_albedo(edx_context* c, float r, float g, float b)
{
    if (!c) return EDX_ERR_INVALID;
    c->albedo[0] = r; c->albedo[1] = g; c->albedo[2] = b;
    return EDX_OK;
}

int edx_set_option(edx_context* c, const char* name, int value)
{
    if (!c || !name) return EDX_ERR_INVALID;
    if (!strcmp(name, "small_max")) { if (value < 0 || value > 64) return fail(c, EDX_ERR_INVALID, "small_max in [0,64]"); c->smallMax = value; return EDX_OK; }
    if (!strcmp(name, "mid_ctas")) { if (value < 1 || value > 32) return fail(c, EDX_ERR_INVALID, "mid_ctas in [1,32]"); c->midCtasPerSm = value; return EDX_OK; }
    if (!strcmp(name, "mid_max")) { if (value < 0 || value > 1024) return fail(c, EDX_ERR_INVALID, "mid_max in [0,1024]"); c->midMax = value; return EDX_OK; }
    if (!strcmp(name, "small_max_clip")) { if (value < 0 || value > 64) return fail(c, EDX_ERR_INVALID, "small_max_clip in [0,64]"); c->smallMaxClip = value; return EDX_OK; }
    if (!strcmp(name, "cluster_cull")) { if (value < 0 || value > 2) return fail(c, EDX_ERR_INVALID, "cluster_cull: 0 off, 1 auto, 2 always"); c->clusterCull = value; return EDX_OK; }
    if (!strcmp(name, "front_end")) { if (value < -1 || value > 2) return fail(c, EDX_ERR_INVALID, "front_end: -1 auto, 0 per-cluster CTAs, 1 cull + work list, 2 cull + per-vertex stage + work list"); c->frontEnd = value; return EDX_OK; }
    if (!strcmp(name, "skip_idle")) { c->skipIdle = value ? 1 : 0; return EDX_OK; }
    if (!strcmp(name, "sort_big")) { c->sortBig = value ? 1 : 0; return EDX_OK; }
    if (!strcmp(name, "bin_min")) { c->binMin = value; return EDX_OK; }
    if (!strcmp(name, "skip_tile")) { c->skipTile = value ? 1 : 0; return EDX_OK; }
    if (!strcmp(name, "mid_auto")) { c->midAuto = value ? 1 : 0; c->midShrunk = false; return EDX_OK; }
    if (!strcmp(name, "graphs")) { if (value < 0 || value > 2) return fail(c, EDX_ERR_INVALID, "graphs: 0 never, 1 small meshes, 2 always"); c->useGraphs = value; return EDX_OK; }
    if (!strcmp(name, "pdl")) { c->pdl = value ? 1 : 0; return EDX_OK; }
    if (!strcmp(name, "fuse_clip")) { c->fuseClip = value ? 1 : 0; return EDX_OK; }
    if (!strcmp(name, "hiz")) { c->hiz = value ? 1 : 0; return EDX_OK; }
    if (!strcmp(name, "lean_resolve")) { if (value < 0 || value > 2) return fail(c, EDX_ERR_INVALID, "lean_resolve: 0 never, 1 auto, 2 always"); c->leanResolve = value; return EDX_OK; }
    if (!strcmp(name, "clip_carveout")) { if (value < 0 || value > 2) return fail(c, EDX_ERR_INVALID, "clip_carveout: 1 L1 (default), 2 shared"); if (value == 0) value = 1; c->clipCarveout = value; if (int r = bind(c)) return r; tune_clip_carveout(c, c->stats.tile_pairs); return EDX_OK; }
    return fail(c, EDX_ERR_INVALID, std::string("unknown option ") + name);
}

int edx_mesh_create(edx_context* c, const void* vertices, uint32_t nv, const uint32_t* indices, uint32_t nt,
                    const uint32_t* tex_ids, edx_mesh** out)
{
    if (!c || !out || (nv && !vertices) || (nt && !indices)) return fail(c, EDX_ERR_INVALID, "null buffer");
    if (nt > (1u << 29) - 1) return fail(c, EDX_ERR_UNSUPPORTED, "more than 2^29-1 triangles per mesh (prim id = tri*8 + fan)");
    if (int r = bind(c)) return r;
    edx_mesh* m = new edx_mesh;
    m->device = c->device;
    m->capVerts = std::max(nv, 1u); m->capTris = std::max(nt, 1u);
    cudaError_t e = cudaMalloc(&m->pos4, (size_t)m->capVerts * 16);
    if (e == cudaSuccess) e = cudaMalloc(&m->nrm4, (size_t)m->capVerts * 16);
    if (e == cudaSuccess) e = cudaMalloc(&m->i0, (size_t)m->capTris * 4);
    if (e == cudaSuccess) e = cudaMalloc(&m->i1, (size_t)m->capTris * 4);
    if (e == cudaSuccess) e = cudaMalloc(&m->i2, (size_t)m->capTris * 4);
    if (e == cudaSuccess) e = cudaMalloc(&m->clusterBox, (size_t)((m->capTris + 255) / 256) * 32);
    if (e == cudaSuccess) e = cudaMalloc(&m->vclusterBox, (size_t)((m->capVerts + 255) / 256) * 32);
    if (e != cudaSuccess) { edx_mesh_destroy(c, m); return fail(c, EDX_ERR_OOM, cudaGetErrorString(e)); }
    if (int r = upload_mesh(c, m, vertices, nv, indices, nt)) { edx_mesh_destroy(c, m); return r; }
    if (tex_ids && nt) {                 // Mesh::GetTextureIds (Mesh.h:58): kept until edx_mesh_set_textures passes its own
        if (cudaMalloc(&m->texIds, (size_t)m->capTris * 4) != cudaSuccess ||
            cudaMemcpyAsync(m->texIds, tex_ids, (size_t)nt * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
            edx_mesh_destroy(c, m);
            return fail(c, EDX_ERR_OOM, "texture id upload failed");
        }
    }
    // copy semantics of CreateVertexBuffer / CreateIndexBuffer: the caller's arrays are free on return
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { edx_mesh_destroy(c, m); return fail(c, EDX_ERR_CUDA, "mesh upload failed"); }
    // Is the triangle order spatially coherent? Cluster culling (geom_kernel prologue) only pays when a
    // cluster of 256 consecutive triangles is small compared with the mesh; decided once per upload.
    {
        const uint32_t nc = (nt + 255) / 256;
        std::vector<float4> boxes(2 * (size_t)nc);
        m->coherent = false;
        if (nc >= 8 && cudaMemcpy(boxes.data(), m->clusterBox, boxes.size() * sizeof(float4), cudaMemcpyDeviceToHost) == cudaSuccess) {
            float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
            for (uint32_t k = 0; k < nc; k++) {
                const float* a = &boxes[2 * k].x; const float* b = &boxes[2 * k + 1].x;
                for (int ax = 0; ax < 3; ax++) { lo[ax] = std::min(lo[ax], a[ax]); hi[ax] = std::max(hi[ax], b[ax]); }
            }
            double sum = 0.0;
            for (uint32_t k = 0; k < nc; k++) {
                const float* a = &boxes[2 * k].x; const float* b = &boxes[2 * k + 1].x;
                // geometric mean of the per-axis extent ratios (a thin strip of a terrain is coherent even
                // though it spans most of the height range)
                double r = 1.0; int axes = 0;
                for (int ax = 0; ax < 3; ax++) {
                    const double ext = (double)hi[ax] - lo[ax];
                    if (ext > 0) { r *= std::max(1e-6, ((double)b[ax] - a[ax]) / ext); axes++; }
                }
                sum += axes ? std::pow(r, 1.0 / axes) : 1.0;
            }
            m->coherent = sum / nc < 0.25;
        }
    }
    *out = m;
    return EDX_OK;
}

int edx_mesh_update(edx_context* c, edx_mesh* m, const void* vertices, uint32_t nv, const uint32_t* indices, uint32_t nt)
{
    if (!c || !m || (nv && !vertices) || (nt && !indices)) return fail(c, EDX_ERR_INVALID, "null buffer");
    if (nv > m->capVerts || nt > m->capTris) return fail(c, EDX_ERR_INVALID, "mesh_update larger than the mesh's capacity");
    if (int r = bind(c)) return r;
    return upload_mesh(c, m, vertices, nv, indices, nt);
}

int edx_mesh_set_textures(edx_context* c, edx_mesh* m, const edx_texture_desc* textures, uint32_t count, const uint32_t* tri_tex_ids)
{
    if (!c || !m || (count && !textures)) return fail(c, EDX_ERR_INVALID, "null argument");
    if (count > 4096) return fail(c, EDX_ERR_UNSUPPORTED, "more than 4096 texture slots");
    if (int r = bind(c)) return r;
    std::vector<TexDesc> descs(count);
    size_t total = 0;
    for (uint32_t i = 0; i < count; i++) {
        const edx_texture_desc& t = textures[i];
        TexDesc& d = descs[i];
        memset(&d, 0, sizeof(d));
        if (t.kind == EDX_TEXTURE_CONSTANT) {
            d.kind = 0; d.r = t.color[0]; d.g = t.color[1]; d.b = t.color[2];
        } else if (t.kind == EDX_TEXTURE_IMAGE) {
            if (!t.rgba8 || t.width == 0 || t.height == 0 || t.width > 32768 || t.height > 32768)
                return fail(c, EDX_ERR_INVALID, "image texture needs pixels and 1..32768 texels per side");
            d.kind = 1; d.w = t.width; d.h = t.height;
            uint32_t w = t.width, h = t.height;
            for (;;) {                                              // shim 19: halve (floor, min 1) down to 1x1
                if (total + (size_t)w * h > 0xFFFFFFF0ull) return fail(c, EDX_ERR_UNSUPPORTED, "texel pool would exceed 2^32 texels");
                d.off[d.levels++] = (uint32_t)total;
                total += (size_t)w * h;
                if (w == 1 && h == 1) break;
                w = std::max(1u, w >> 1); h = std::max(1u, h >> 1);
            }
        } else {
            return fail(c, EDX_ERR_INVALID, "unknown texture kind");
        }
    }
    if (tri_tex_ids)
        for (uint32_t i = 0; i < m->nTris; i++)
            if (tri_tex_ids[i] >= std::max(count, 1u)) return fail(c, EDX_ERR_INVALID, "texture id out of range");
    // the mesh may be rendering (shared, frames in flight): wait for the device before replacing its tables
    EDX_CUDA(c, cudaDeviceSynchronize());
    dev_free(m->texDesc); dev_free(m->texels);
    if (tri_tex_ids) dev_free(m->texIds);
    m->nTex = 0;
    if (count == 0) return EDX_OK;
    EDX_CUDA(c, cudaMalloc(&m->texDesc, count * sizeof(TexDesc)));
    EDX_CUDA(c, cudaMalloc(&m->texels, std::max<size_t>(total, 1) * sizeof(uchar4)));
    EDX_CUDA(c, cudaMemcpyAsync(m->texDesc, descs.data(), count * sizeof(TexDesc), cudaMemcpyHostToDevice, c->stream));
    for (uint32_t i = 0; i < count; i++) {
        const TexDesc& d = descs[i];
        if (d.kind != 1) continue;
        EDX_CUDA(c, cudaMemcpyAsync(m->texels + d.off[0], textures[i].rgba8, (size_t)d.w * d.h * 4, cudaMemcpyHostToDevice, c->stream));
        uint32_t w = d.w, h = d.h;
        for (uint32_t l = 1; l < d.levels; l++) {
            const uint32_t dw = std::max(1u, w >> 1), dh = std::max(1u, h >> 1);
            mip_kernel<<<(dw * dh + 255) / 256, 256, 0, c->stream>>>(m->texels, d.off[l - 1], (int)w, (int)h, d.off[l], (int)dw, (int)dh);
            w = dw; h = dh;
        }
    }
    if (tri_tex_ids && m->nTris) {
        // sized for the mesh's capacity (edx_mesh_update may raise the triangle count later); the rest reads slot 0
        EDX_CUDA(c, cudaMalloc(&m->texIds, (size_t)m->capTris * 4));
        EDX_CUDA(c, cudaMemsetAsync(m->texIds, 0, (size_t)m->capTris * 4, c->stream));
        EDX_CUDA(c, cudaMemcpyAsync(m->texIds, tri_tex_ids, (size_t)m->nTris * 4, cudaMemcpyHostToDevice, c->stream));
    }
    EDX_CUDA(c, cudaGetLastError());
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));        // copy semantics: the caller's arrays are free on return
    m->nTex = count;
    return EDX_OK;
}

int edx_mesh_read_texture_level(edx_context* c, const edx_mesh* m, uint32_t slot, uint32_t level, uint8_t* out_rgba8, uint32_t* out_w, uint32_t* out_h)
{
    if (!c || !m || slot >= m->nTex) return fail(c, EDX_ERR_INVALID, "no such texture slot");
    if (int r = bind(c)) return r;
    TexDesc d;
    EDX_CUDA(c, cudaMemcpy(&d, m->texDesc + slot, sizeof(d), cudaMemcpyDeviceToHost));
    if (d.kind != 1 || level >= d.levels) return fail(c, EDX_ERR_INVALID, "no such mip level");
    const uint32_t w = std::max(1u, d.w >> level), h = std::max(1u, d.h >> level);
    if (out_w) *out_w = w;
    if (out_h) *out_h = h;
    if (out_rgba8) EDX_CUDA(c, cudaMemcpy(out_rgba8, m->texels + d.off[level], (size_t)w * h * 4, cudaMemcpyDeviceToHost));
    return EDX_OK;
}

int edx_mesh_destroy(edx_context* c, edx_mesh* m)
{
    if (!m) return EDX_OK;
    // A mesh may be shared by several contexts of its device (frames in flight), and ctx may be NULL (the mesh
    // outlived its context): wait for the whole device. Contexts other than ctx that rendered it last must be
    // synchronised by the caller first (their overflow re-run would need the mesh).
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    if (c && c->lastMesh == m) { c->lastMesh = nullptr; c->framePending = false; }
    dev_free(m->pos4); dev_free(m->nrm4); dev_free(m->i0); dev_free(m->i1); dev_free(m->i2); dev_free(m->clusterBox); dev_free(m->vclusterBox);
    dev_free(m->texDesc); dev_free(m->texels); dev_free(m->texIds);
    if (m->staging) cudaFree(m->staging);
    delete m;
    return EDX_OK;
}

int edx_render_mesh(edx_context* c, const edx_mesh* m)
{
    if (!c || !m) return fail(c, EDX_ERR_INVALID, "null mesh");
    if (!c->initialized) return fail(c, EDX_ERR_INVALID, "Initialize has not been called");
    if (m->device != c->device) return fail(c, EDX_ERR_INVALID, "the mesh lives on another device");
    if (int r = bind(c)) return r;
    c->lastMesh = m;
    c->stats.submitted_tris = m->nTris;
    c->sinkSerial++;                 // (a frame that has to be re-run after a queue overflow keeps its number)
    save_submitted(c, c->submitted);
    if (int r = enqueue_frame(c, m, nullptr, 0)) return r;
    c->framePending = true;
    return EDX_OK;
}

int edx_synchronize(edx_context* c)
{
    if (!c) return EDX_ERR_INVALID;
    if (int r = bind(c)) return r;
    return finish_frame(c);
}

const uint8_t* edx_get_back_buffer(edx_context* c)
{
    if (!c || !c->initialized) return nullptr;
    if (bind(c) || finish_frame(c)) return nullptr;
    if (cudaMemcpyAsync(c->hostColor, c->extColor ? c->extColor : c->color, c->hostColorBytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        c->error = "frame read-back failed";
        return nullptr;
    }
    return c->hostColor;
}

int edx_read_depth(edx_context* c, float* out)
{
    if (!c || !out || !c->initialized) return fail(c, EDX_ERR_INVALID, "bad argument");
    if (int r = bind(c)) return r;
    if (int r = finish_frame(c)) return r;
    EDX_CUDA(c, cudaMemcpyAsync(out, c->extDepth ? c->extDepth : c->depth, (size_t)c->width * c->height * 4, cudaMemcpyDeviceToHost, c->stream));
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    return EDX_OK;
}

int edx_read_sample(edx_context* c, int sample, float* depth, uint32_t* ids)
{
    if (!c || !c->initialized || sample < 0 || sample >= (1 << c->msaaLog2) || (!depth && !ids)) return fail(c, EDX_ERR_INVALID, "bad argument");
    if (ids && !c->captureIds) return fail(c, EDX_ERR_INVALID, "edx_set_capture_ids(ctx, 1) must be set before the frame");
    if (c->extDepth && depth) return fail(c, EDX_ERR_UNSUPPORTED, "per-sample read-back needs the context's own depth buffer");
    if (int r = bind(c)) return r;
    if (int r = finish_frame(c)) return r;
    const size_t plane = (size_t)c->width * c->height;
    if (depth) EDX_CUDA(c, cudaMemcpyAsync(depth, c->depth + (size_t)sample * plane, plane * 4, cudaMemcpyDeviceToHost, c->stream));
    if (ids) EDX_CUDA(c, cudaMemcpyAsync(ids, c->ids + (size_t)sample * plane, plane * 4, cudaMemcpyDeviceToHost, c->stream));
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    return EDX_OK;
}

int edx_set_capture_ids(edx_context* c, int enabled)
{
    if (!c) return EDX_ERR_INVALID;
    c->captureIds = enabled ? 1 : 0;
    return EDX_OK;
}

int edx_read_winner_ids(edx_context* c, uint32_t* out)
{
    if (!c || !out || !c->initialized) return fail(c, EDX_ERR_INVALID, "bad argument");
    if (!c->captureIds) return fail(c, EDX_ERR_INVALID, "edx_set_capture_ids(ctx, 1) must be set before the frame");
    if (int r = bind(c)) return r;
    if (int r = finish_frame(c)) return r;
    EDX_CUDA(c, cudaMemcpyAsync(out, c->ids, (size_t)c->width * c->height * 4, cudaMemcpyDeviceToHost, c->stream));
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    return EDX_OK;
}

int edx_debug_clip_vertices(edx_context* c, const edx_mesh* m, float* out)
{
    if (!c || !m || !out) return fail(c, EDX_ERR_INVALID, "bad argument");
    if (int r = bind(c)) return r;
    if (!m->nVerts) return EDX_OK;
    float4* d = nullptr;
    EDX_CUDA(c, cudaMalloc(&d, (size_t)m->nVerts * 16));
    FrameParams P;
    fill_params(c, m, P);
    vertex_transform_kernel<<<(m->nVerts + 255) / 256, 256, 0, c->stream>>>(P, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, (size_t)m->nVerts * 16, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(c, EDX_ERR_CUDA, cudaGetErrorString(e));
    return EDX_OK;
}

int edx_debug_raster_triangles(edx_context* c, const edx_mesh* m, uint64_t capacity, int32_t* ints7, float* floats7, uint64_t* count)
{
    if (!c || !m || !ints7 || !floats7 || !count || !c->initialized) return fail(c, EDX_ERR_INVALID, "bad argument");
    if (capacity == 0 || capacity > 0x7FFFFFFFull) return fail(c, EDX_ERR_INVALID, "bad capacity");
    if (int r = bind(c)) return r;
    if (int r = finish_frame(c)) return r;
    DumpRec* d = nullptr;
    EDX_CUDA(c, cudaMalloc(&d, (size_t)capacity * sizeof(DumpRec)));
    int rc = EDX_OK;
    uint32_t n = 0;
    for (int attempt = 0; attempt < 8; attempt++) {
        c->lastMesh = m;
        rc = enqueue_frame(c, m, d, (uint32_t)capacity);
        if (rc) break;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(c, EDX_ERR_CUDA, "dump frame failed"); break; }
        const Counters& k = *c->hostCounters;
        n = k.nDump;
        if (k.nClipQueue > c->clipQueueCap) {        // the dump must see every clipped triangle
            if ((rc = grow(c, c->clipQueue, c->clipQueueCap, k.nClipQueue))) break;
            continue;
        }
        break;
    }
    if (!rc && n > capacity) rc = fail(c, EDX_ERR_OVERFLOW, "dump capacity too small");
    if (!rc) {
        std::vector<DumpRec> h(n);
        if (n && cudaMemcpy(h.data(), d, (size_t)n * sizeof(DumpRec), cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(c, EDX_ERR_CUDA, "dump copy failed");
        else {
            std::sort(h.begin(), h.end(), [](const DumpRec& a, const DumpRec& b) { return (uint32_t)a.i[0] < (uint32_t)b.i[0]; });
            for (uint32_t i = 0; i < n; i++) { memcpy(ints7 + 7 * (size_t)i, h[i].i, 28); memcpy(floats7 + 7 * (size_t)i, h[i].f, 28); }
            *count = n;
        }
    }
    cudaFree(d);
    c->seenOverFrames = c->hostCounters->overFrames;     // attempts that overflowed were repeated right here: nothing was lost
    c->framePending = true;              // a regular frame was rendered alongside; let finish_frame vet its queues
    if (!rc) rc = finish_frame(c);
    return rc;
}

int edx_get_derived_state(const edx_context* c, float mvp[16], float eye[3], float light[3])
{
    if (!c) return EDX_ERR_INVALID;
    if (mvp) memcpy(mvp, c->mvp.m, 64);
    if (eye) memcpy(eye, c->eye, 12);
    if (light) memcpy(light, c->light, 12);
    return EDX_OK;
}

void* edx_device_color(edx_context* c) { return c ? (c->extColor ? c->extColor : c->color) : nullptr; }
void* edx_device_depth(edx_context* c) { return c ? (c->extDepth ? c->extDepth : c->depth) : nullptr; }

int edx_device_count(void)
{
    int n = 0, ok = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ok++;
    }
    return ok;
}

int edx_enable_peer_access(edx_context* c, int peer)
{
    if (!c) return EDX_ERR_INVALID;
    if (peer == c->device) return EDX_OK;
    if (int r = bind(c)) return r;
    int can = 0;
    EDX_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, peer));
    if (!can) return fail(c, EDX_ERR_UNSUPPORTED, "no peer access between the two devices");
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return EDX_OK; }
    EDX_CUDA(c, e);
    return EDX_OK;
}

int edx_device_alloc(edx_context* c, size_t bytes, void** out)
{
    if (!c || !out || !bytes) return fail(c, EDX_ERR_INVALID, "bad argument");
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaMalloc(out, bytes));
    return EDX_OK;
}

int edx_device_free(edx_context* c, void* p)
{
    if (!c) return EDX_ERR_INVALID;
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaFree(p));
    return EDX_OK;
}

int edx_read_device(edx_context* c, void* dst, const void* src, size_t bytes)
{
    if (!c || !dst || !src) return fail(c, EDX_ERR_INVALID, "bad argument");
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    return EDX_OK;
}

int edx_set_frame_sink(edx_context* c, void* color, void* depth)
{
    if (!c) return EDX_ERR_INVALID;
    if ((color || depth) && c->msaaLog2 != 0) return fail(c, EDX_ERR_UNSUPPORTED, "frame sinks are single-sample only");
    c->sinkColor = color;
    c->sinkDepth = depth;
    return EDX_OK;
}

int edx_flush_frame_sink(edx_context* c)
{
    if (!c) return EDX_ERR_INVALID;
    if (int r = bind(c)) return r;
    if (c->pushPending) { c->pushPending = false; EDX_CUDA(c, cudaStreamWaitEvent(c->stream, c->evPushDone, 0)); }
    return EDX_OK;
}

int edx_set_frame_sink_signal(edx_context* c, void* word)
{
    if (!c) return EDX_ERR_INVALID;
    if (word && ((uintptr_t)word & 3u)) return fail(c, EDX_ERR_INVALID, "the signal word must be 4-byte aligned");
    c->sinkSignal = (uint32_t*)word;
    c->sinkSerial = 0;
    return EDX_OK;
}

int edx_set_render_target(edx_context* c, void* color, void* depth)
{
    if (!c) return EDX_ERR_INVALID;
    if ((color || depth) && c->msaaLog2 != 0) return fail(c, EDX_ERR_UNSUPPORTED, "caller-owned render targets are single-sample only");
    c->extColor = (uchar4*)color;
    c->extDepth = (float*)depth;
    return EDX_OK;
}

int edx_set_screen_partition(edx_context* c, int part, int parts)
{
    if (!c || parts < 1 || parts > 256 || part < 0 || part >= parts) return fail(c, EDX_ERR_INVALID, "need 0 <= part < parts <= 256");
    c->part = part; c->parts = parts;
    return EDX_OK;
}

int edx_set_stream(edx_context* c, void* s)
{
    if (!c) return EDX_ERR_INVALID;
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->ownStream = false;
    return EDX_OK;
}

int edx_timer_begin(edx_context* c)
{
    if (!c) return EDX_ERR_INVALID;
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaEventRecord(c->evTimer[0], c->stream));
    return EDX_OK;
}

int edx_timer_end(edx_context* c, float* ms)
{
    if (!c || !ms) return EDX_ERR_INVALID;
    if (int r = bind(c)) return r;
    EDX_CUDA(c, cudaEventRecord(c->evTimer[1], c->stream));
    EDX_CUDA(c, cudaEventSynchronize(c->evTimer[1]));
    EDX_CUDA(c, cudaEventElapsedTime(ms, c->evTimer[0], c->evTimer[1]));
    return finish_frame(c);
}

int edx_set_profiling(edx_context* c, int enabled)
{
    if (!c) return EDX_ERR_INVALID;
    c->profiling = enabled ? 1 : 0;
    return EDX_OK;
}

int edx_get_stats(edx_context* c, edx_stats* out)
{
    if (!c || !out) return EDX_ERR_INVALID;
    *out = c->stats;
    return EDX_OK;
}

const char* edx_last_launch_list(const edx_context* c) { return c ? c->launchList.c_str() : ""; }

int edx_last_launch_count(const edx_context* c) { return c ? c->launches : 0; }

int edx_debug_tile_residency(edx_context* c, int* ctas_per_sm)
{
    if (!c || !ctas_per_sm) return EDX_ERR_INVALID;
    if (int r = bind(c)) return r;
    uint32_t* d = nullptr;
    EDX_CUDA(c, cudaMalloc(&d, 1024 * 4));
    EDX_CUDA(c, cudaMemsetAsync(d, 0, 1024 * 4, c->stream));
    static bool attr[64];                                       // function attributes are per device
    if (!attr[c->device & 63]) { cudaFuncSetAttribute(residency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared)); attr[c->device & 63] = true; }
    residency_kernel<<<2048, TILE_THREADS, sizeof(TileShared), c->stream>>>(d);
    uint32_t h[1024];
    cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(c, EDX_ERR_CUDA, cudaGetErrorString(e));
    int m = 0, sms = 0;
    for (int i = 0; i < 256; i++) { m = std::max(m, (int)h[256 + i]); sms += h[256 + i] ? 1 : 0; }
    *ctas_per_sm = m;
    if (getenv("EDX_DEBUG_PRINT")) fprintf(stderr, "[edx dbg] residency: %d CTAs/SM on %d SMs; 100000 SM cycles took %u ns (%.0f MHz)\n", m, sms, h[512], h[512] ? 1e8 / h[512] : 0.0);
    return EDX_OK;
}

int edx_write_frame_to_file(edx_context* c, const char* path)
{
    if (!c || !path) return fail(c, EDX_ERR_INVALID, "null path");
    const uint8_t* px = edx_get_back_buffer(c);
    if (!px) return EDX_ERR_CUDA;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(c, EDX_ERR_INVALID, std::string("cannot open ") + path);
    const uint32_t w = c->width, h = c->height, rowBytes = (w * 3 + 3) & ~3u, size = 54 + rowBytes * h;
    uint8_t hdr[54] = { 'B', 'M' };
    auto put32 = [&](int at, uint32_t v) { hdr[at] = v & 255; hdr[at + 1] = (v >> 8) & 255; hdr[at + 2] = (v >> 16) & 255; hdr[at + 3] = (v >> 24) & 255; };
    put32(2, size); put32(10, 54); put32(14, 40); put32(18, w); put32(22, h);
    hdr[26] = 1; hdr[28] = 24; put32(34, rowBytes * h);
    fwrite(hdr, 1, 54, f);
    std::vector<uint8_t> row(rowBytes, 0);
    for (uint32_t y = 0; y < h; y++) {          // both BMP and the back buffer are bottom-up
        const uint8_t* src = px + (size_t)y * w * 4;
        for (uint32_t x = 0; x < w; x++) { row[3 * x] = src[4 * x + 2]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x]; }
        fwrite(row.data(), 1, rowBytes, f);
    }
    fclose(f);
    return EDX_OK;
}

} // extern "C"
