// ort_multi.cu -- the multi-GPU side of the C ABI.
//
// The radiance loop shards with no data-path dependency: the scene is replicated, every (pixel, chunk)
// sample stream is independent, and each GPU adds its chunk sums into its own int64 fixed-point
// framebuffer.  The only exchange step is the sum of those framebuffers.  It is done here by ONE kernel
// on the root GPU that reads the peers' framebuffers over NVLink peer memory (P2P loads through NVSwitch),
// adds them and resolves to float3 pixels / RGBE words in the same pass -- no staging copy, no separate
// collective.  Two ways to get at the peers' memory:
//   * one process, N devices  (OrtMulti):  cudaDeviceEnablePeerAccess, plain pointers;
//   * one process per GPU (torchrun, bench.py):  CUDA IPC handles of the framebuffers, exchanged by the
//     host-side plumbing (torch.distributed), opened once on the root.
// Integer addition is associative, so the N-GPU image is bit-identical to the 1-GPU image.
//
// Replaces: the nine tile workers + ThreadWorkQueue of the reference (code/macos_main.mm:165-240, 574-671,
// code/platform.h:307-339), whose unit of work is a 32x32 tile on one shared-memory machine.
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "ort_internal.h"

extern "C" int ort_scene_device(const OrtScene *s, int *device);

namespace {

#define ORT_MAX_PEERS 16

struct PeerList { const long long *p[ORT_MAX_PEERS]; uint32_t n; };

// radiance RGBE word, identical to kernels.cuh: rgbe_encode (v3_to_rgbe, macos_main.mm:242-261)
__device__ __forceinline__ uint32_t rgbe_word(float x, float y, float z)
{
    float m = x > y ? x : y; m = m > z ? m : z;
    if(!(m >= 1e-32f)) return 0u;
    int e;
    float denom = frexpf(m, &e) * 255.0f / m;
    return ((uint32_t)roundf(x * denom) << 0) | ((uint32_t)roundf(y * denom) << 8) |
           ((uint32_t)roundf(z * denom) << 16) | ((uint32_t)(e + 128) << 24);
}

// acc (this GPU's sums) += sum over peers; optionally resolved in the same pass:
//   rgb  != 0: float3 pixels = sum / 2^24 / spp           (ray.cpp:1428)
//   rgbe != 0: RGBE words in .hdr file order              (macos_main.mm:682-707)
// One thread per pixel: 32 B from every framebuffer as two 128-bit loads (the peers' over NVLink).
__global__ void __launch_bounds__(256)
k_accum_reduce_resolve(long long *__restrict__ acc, PeerList peers, int32_t w, int32_t h, uint32_t spp,
                       float *__restrict__ rgb, uint32_t *__restrict__ rgbe, int write_back)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= (size_t)w * h) return;
    longlong2 a = reinterpret_cast<const longlong2 *>(acc)[2 * i], b = reinterpret_cast<const longlong2 *>(acc)[2 * i + 1];
    for(uint32_t k = 0; k < peers.n; ++k)
    {
        longlong2 pa = reinterpret_cast<const longlong2 *>(peers.p[k])[2 * i], pb = reinterpret_cast<const longlong2 *>(peers.p[k])[2 * i + 1];
        a.x += pa.x; a.y += pa.y; b.x += pb.x; b.y += pb.y;
    }
    if(write_back && peers.n)
    {
        reinterpret_cast<longlong2 *>(acc)[2 * i] = a;
        reinterpret_cast<longlong2 *>(acc)[2 * i + 1] = b;
    }
    if(rgb || rgbe)
    {
        const double inv = 1.0 / (double)(1 << ORT_ACCUM_FRAC_BITS);
        float r = (float)((double)a.x * inv) / (float)spp, g = (float)((double)a.y * inv) / (float)spp, bl = (float)((double)b.x * inv) / (float)spp;
        if(rgb) { rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = bl; }
        if(rgbe)
        {
            int x = (int)(i % (size_t)w), y = (int)(i / (size_t)w);
            rgbe[(size_t)(h - 1 - y) * w + x] = rgbe_word(r, g, bl);
        }
    }
}

int launch_reduce(int device, long long *acc, const void *const *peer_accums, uint32_t n_peers, int32_t w, int32_t h,
                  uint32_t spp, float *rgb, uint32_t *rgbe, int write_back, cudaStream_t st)
{
    if(n_peers > ORT_MAX_PEERS) return ort::fail_with(ORT_ERR_ARG, "too many peers");
    ORT_CUDA_TRY(cudaSetDevice(device));
    PeerList pl; pl.n = n_peers;
    for(uint32_t k = 0; k < n_peers; ++k) pl.p[k] = (const long long *)peer_accums[k];
    size_t n = (size_t)w * h;
    k_accum_reduce_resolve<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, pl, w, h, spp ? spp : 1u, rgb, rgbe, write_back);
    ORT_CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------------
struct OrtMulti
{
    std::vector<OrtScene *> scenes;
    std::vector<int> devices;
    std::vector<long long *> accum;      // one int64[h*w*4] framebuffer per device
    std::vector<cudaStream_t> streams;
    std::vector<bool> direct;            // root can read this device's memory (peer access); else staged copy
    long long *staging = 0;              // on the root, for devices without peer access
    float *d_rgb = 0; uint32_t *d_rgbe = 0;
    size_t pixels = 0;
    std::string thread_error;
};

// progressive / dynamic dispatch state
struct OrtProgress
{
    OrtMulti *multi;
    OrtCamera camera;
    OrtRenderParams params;
    uint32_t n_chunks;
    std::vector<uint8_t> done;
    std::atomic<uint32_t> cursor;
    bool dirty;                          // peers hold sums not yet folded into the root's framebuffer
};

namespace {

int multi_ensure(OrtMulti *m, size_t pixels)
{
    if(m->pixels >= pixels) return ORT_OK;
    for(size_t i = 0; i < m->devices.size(); ++i)
    {
        ORT_CUDA_TRY(cudaSetDevice(m->devices[i]));
        if(m->accum[i]) cudaFree(m->accum[i]);
        m->accum[i] = 0;
        ORT_CUDA_TRY(cudaMalloc((void **)&m->accum[i], pixels * 4 * sizeof(long long)));
    }
    ORT_CUDA_TRY(cudaSetDevice(m->devices[0]));
    cudaFree(m->d_rgb); cudaFree(m->d_rgbe); cudaFree(m->staging);
    m->d_rgb = 0; m->d_rgbe = 0; m->staging = 0;
    ORT_CUDA_TRY(cudaMalloc((void **)&m->d_rgb, pixels * 3 * sizeof(float)));
    ORT_CUDA_TRY(cudaMalloc((void **)&m->d_rgbe, pixels * sizeof(uint32_t)));
    bool need_staging = false;
    for(size_t i = 1; i < m->devices.size(); ++i) if(!m->direct[i]) need_staging = true;
    if(need_staging) ORT_CUDA_TRY(cudaMalloc((void **)&m->staging, pixels * 4 * sizeof(long long)));
    m->pixels = pixels;
    return ORT_OK;
}

// folds every peer's framebuffer into the root's (and zeroes nothing: callers decide); optional resolve
int multi_reduce(OrtMulti *m, int32_t w, int32_t h, uint32_t spp, bool want_rgb, bool want_rgbe)
{
    const size_t n = m->devices.size();
    cudaStream_t st = m->streams[0];
    std::vector<const void *> direct_peers;
    for(size_t i = 1; i < n; ++i)
    {
        if(m->direct[i]) direct_peers.push_back(m->accum[i]);
        else
        {
            // no peer access between these two devices: one peer copy into the root, then the same kernel
            ORT_CUDA_TRY(cudaSetDevice(m->devices[0]));
            ORT_CUDA_TRY(cudaMemcpyPeerAsync(m->staging, m->devices[0], m->accum[i], m->devices[i], (size_t)w * h * 4 * sizeof(long long), st));
            const void *p = m->staging;
            int rc = launch_reduce(m->devices[0], m->accum[0], &p, 1, w, h, spp, 0, 0, 1, st);
            if(rc != ORT_OK) return rc;
        }
    }
    int rc = launch_reduce(m->devices[0], m->accum[0], direct_peers.data(), (uint32_t)direct_peers.size(), w, h, spp,
                           want_rgb ? m->d_rgb : 0, want_rgbe ? m->d_rgbe : 0, 1, st);
    if(rc != ORT_OK) return rc;
    ORT_CUDA_TRY(cudaStreamSynchronize(st));
    return ORT_OK;
}

int zero_accum(OrtMulti *m, size_t i, int32_t w, int32_t h)
{
    ORT_CUDA_TRY(cudaSetDevice(m->devices[i]));
    ORT_CUDA_TRY(cudaMemsetAsync(m->accum[i], 0, (size_t)w * h * 4 * sizeof(long long), m->streams[i]));
    return ORT_OK;
}

void add_stats(OrtRenderStats *total, const OrtRenderStats &s)
{
    total->samples += s.samples; total->rays += s.rays; total->node_visits += s.node_visits;
    total->box_tests += s.box_tests; total->shape_tests += s.shape_tests; total->kernel_launches += s.kernel_launches;
    if(s.device_ms > total->device_ms) total->device_ms = s.device_ms;      // devices run concurrently: the slowest
    total->extend_ms += s.extend_ms; total->shade_ms += s.shade_ms; total->sort_ms += s.sort_ms;
}

uint32_t chunk_count_of(const OrtRenderParams *P)
{
    uint32_t spp = P->ray_per_pixel_count, c = P->chunk_spp ? P->chunk_spp : spp;
    if(c > spp) c = spp;
    return (spp + c - 1) / c;
}

} // namespace

extern "C" {

int ort_accum_reduce_resolve_device(OrtScene *s, void *accum_device, const void *const *peer_accums, uint32_t n_peers,
                                    int32_t width, int32_t height, uint32_t ray_per_pixel_count,
                                    void *rgb_device, void *rgbe_device, void *stream)
{
    if(!s || !accum_device || width <= 0 || height <= 0 || (n_peers && !peer_accums)) return ort::fail_with(ORT_ERR_ARG, "bad argument");
    if((rgb_device || rgbe_device) && ray_per_pixel_count == 0) return ort::fail_with(ORT_ERR_ARG, "ray_per_pixel_count must be > 0 to resolve");
    int device = 0;
    ort_scene_device(s, &device);
    return launch_reduce(device, (long long *)accum_device, peer_accums, n_peers, width, height, ray_per_pixel_count,
                         (float *)rgb_device, (uint32_t *)rgbe_device, 1, (cudaStream_t)stream);
}

int ort_accum_alloc_device(OrtScene *s, int32_t width, int32_t height, void **accum_device)
{
    if(!s || !accum_device || width <= 0 || height <= 0) return ort::fail_with(ORT_ERR_ARG, "bad argument");
    int device = 0;
    ort_scene_device(s, &device);
    ORT_CUDA_TRY(cudaSetDevice(device));
    *accum_device = 0;
    ORT_CUDA_TRY(cudaMalloc(accum_device, (size_t)width * height * 4 * sizeof(long long)));
    ORT_CUDA_TRY(cudaMemset(*accum_device, 0, (size_t)width * height * 4 * sizeof(long long)));
    return ORT_OK;
}

int ort_accum_free_device(OrtScene *s, void *accum_device)
{
    if(!s) return ort::fail_with(ORT_ERR_ARG, "null scene");
    int device = 0;
    ort_scene_device(s, &device);
    ORT_CUDA_TRY(cudaSetDevice(device));
    ORT_CUDA_TRY(cudaFree(accum_device));
    return ORT_OK;
}

int ort_accum_ipc_export(OrtScene *s, const void *accum_device, uint8_t handle[ORT_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == ORT_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    if(!s || !accum_device || !handle) return ort::fail_with(ORT_ERR_ARG, "null argument");
    int device = 0;
    ort_scene_device(s, &device);
    ORT_CUDA_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    ORT_CUDA_TRY(cudaIpcGetMemHandle(&h, (void *)accum_device));
    memcpy(handle, &h, sizeof(h));
    return ORT_OK;
}

int ort_accum_ipc_open(OrtScene *s, const uint8_t handle[ORT_IPC_HANDLE_BYTES], void **peer_accum)
{
    if(!s || !handle || !peer_accum) return ort::fail_with(ORT_ERR_ARG, "null argument");
    int device = 0;
    ort_scene_device(s, &device);
    ORT_CUDA_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *peer_accum = 0;
    ORT_CUDA_TRY(cudaIpcOpenMemHandle(peer_accum, h, cudaIpcMemLazyEnablePeerAccess));
    return ORT_OK;
}

int ort_accum_ipc_close(OrtScene *s, void *peer_accum)
{
    if(!s) return ort::fail_with(ORT_ERR_ARG, "null scene");
    int device = 0;
    ort_scene_device(s, &device);
    ORT_CUDA_TRY(cudaSetDevice(device));
    if(peer_accum) ORT_CUDA_TRY(cudaIpcCloseMemHandle(peer_accum));
    return ORT_OK;
}

// ---- one process, N devices --------------------------------------------------------------------------
int ort_multi_create(OrtScene *const *scenes, uint32_t n, OrtMulti **out)
{
    ORT_GUARD_BEGIN
    if(!scenes || !out || n == 0 || n > ORT_MAX_PEERS + 1) return ort::fail_with(ORT_ERR_ARG, "bad argument");
    *out = 0;
    OrtMulti *m = new OrtMulti();
    for(uint32_t i = 0; i < n; ++i)
    {
        int dev = 0;
        if(!scenes[i] || ort_scene_device(scenes[i], &dev) != ORT_OK) { delete m; return ort::fail_with(ORT_ERR_ARG, "null scene in the list"); }
        for(int d : m->devices) if(d == dev) { delete m; return ort::fail_with(ORT_ERR_ARG, "two scenes on the same device"); }
        m->scenes.push_back(scenes[i]); m->devices.push_back(dev);
    }
    m->accum.assign(n, (long long *)0); m->streams.assign(n, (cudaStream_t)0); m->direct.assign(n, true);
    for(uint32_t i = 0; i < n; ++i)
    {
        cudaError_t e = cudaSetDevice(m->devices[i]);
        if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking);
        if(e != cudaSuccess) { ort_multi_destroy(m); return ort::fail_with(ORT_ERR_CUDA, cudaGetErrorString(e)); }
    }
    // the root reads the peers' framebuffers in place over NVLink
    cudaSetDevice(m->devices[0]);
    for(uint32_t i = 1; i < n; ++i)
    {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->devices[0], m->devices[i]);
        if(can)
        {
            cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[i], 0);
            if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
            cudaGetLastError();
        }
        m->direct[i] = can != 0;
    }
    *out = m;
    return ORT_OK;
    ORT_GUARD_END
}

int ort_multi_destroy(OrtMulti *m)
{
    if(!m) return ORT_OK;
    for(size_t i = 0; i < m->devices.size(); ++i)
    {
        cudaSetDevice(m->devices[i]);
        if(m->streams[i]) { cudaStreamSynchronize(m->streams[i]); cudaStreamDestroy(m->streams[i]); }
        cudaFree(m->accum[i]);
    }
    if(!m->devices.empty()) { cudaSetDevice(m->devices[0]); cudaFree(m->d_rgb); cudaFree(m->d_rgbe); cudaFree(m->staging); }
    delete m;
    return ORT_OK;
}

int ort_multi_device_count(const OrtMulti *m, uint32_t *n, uint32_t *n_peer_direct)
{
    if(!m || !n) return ort::fail_with(ORT_ERR_ARG, "null argument");
    *n = (uint32_t)m->devices.size();
    if(n_peer_direct) { uint32_t c = 0; for(size_t i = 1; i < m->direct.size(); ++i) c += m->direct[i] ? 1u : 0u; *n_peer_direct = c; }
    return ORT_OK;
}

// renders chunks [lo, hi) of the frame, split statically over the devices, into the per-device framebuffers
static int multi_render_range(OrtMulti *m, const OrtCamera *camera, const OrtRenderParams *P, uint32_t lo, uint32_t hi, OrtRenderStats *stats)
{
    const uint32_t n = (uint32_t)m->devices.size(), count = hi - lo;
    std::vector<int> rcs(n, ORT_OK);
    std::vector<std::string> errs(n);
    std::vector<OrtRenderStats> sts(n);
    auto work = [&](uint32_t i)
    {
        memset(&sts[i], 0, sizeof(OrtRenderStats));
        uint32_t base = count / n, rem = count % n;
        uint32_t b = lo + i * base + (i < rem ? i : rem), e = b + base + (i < rem ? 1u : 0u);
        if(e <= b) return;
        OrtRenderParams Q = *P;
        if(Q.chunk_spp == 0) Q.chunk_spp = Q.ray_per_pixel_count;      // explicit, so that (0, 0) is not read as "all"
        Q.chunk_begin = b; Q.chunk_end = e;
        cudaSetDevice(m->devices[i]);
        rcs[i] = ort_render_accumulate_device(m->scenes[i], camera, &Q, m->accum[i], m->streams[i], &sts[i]);
        if(rcs[i] != ORT_OK) errs[i] = ort_last_error();
    };
    std::vector<std::thread> threads;
    for(uint32_t i = 1; i < n; ++i) threads.emplace_back(work, i);
    work(0);
    for(auto &t : threads) t.join();
    for(uint32_t i = 0; i < n; ++i)
    {
        if(rcs[i] != ORT_OK) return ort::fail_with(rcs[i], "device " + std::to_string(m->devices[i]) + ": " + errs[i]);
        if(stats) add_stats(stats, sts[i]);
    }
    return ORT_OK;
}

int ort_multi_render(OrtMulti *m, const OrtCamera *camera, const OrtRenderParams *P, ort_v3 *output_buffer, OrtRenderStats *stats)
{
    ORT_GUARD_BEGIN
    if(!m || !camera || !P || !output_buffer) return ort::fail_with(ORT_ERR_ARG, "null argument");
    if(P->output_width <= 0 || P->output_height <= 0 || P->ray_per_pixel_count == 0) return ort::fail_with(ORT_ERR_ARG, "bad params");
    if(P->tile_min_x != 0 || P->tile_min_y != 0 || P->tile_one_past_max_x != P->output_width || P->tile_one_past_max_y != P->output_height)
        return ort::fail_with(ORT_ERR_ARG, "ort_multi_render renders whole images: the tile must cover the output");
    const int32_t w = P->output_width, h = P->output_height;
    int rc = multi_ensure(m, (size_t)w * h);
    if(rc != ORT_OK) return rc;
    if(stats) memset(stats, 0, sizeof(*stats));
    for(size_t i = 0; i < m->devices.size(); ++i) { rc = zero_accum(m, i, w, h); if(rc != ORT_OK) return rc; }
    // the image must not depend on the number of devices: fewer chunks than devices is fine (idle devices),
    // but a single-chunk frame would take the float-sum path on one GPU and the fixed-point path here --
    // chunking is part of the seed schedule, so it is the caller's choice, not ours
    const uint32_t n_chunks = chunk_count_of(P);
    uint32_t lo = P->chunk_begin, hi = P->chunk_end;
    if(lo == 0 && hi == 0) hi = n_chunks;
    if(hi > n_chunks) hi = n_chunks;
    if(lo > hi) lo = hi;
    rc = multi_render_range(m, camera, P, lo, hi, stats);
    if(rc != ORT_OK) return rc;
    rc = multi_reduce(m, w, h, P->ray_per_pixel_count, true, false);
    if(rc != ORT_OK) return rc;
    ORT_CUDA_TRY(cudaSetDevice(m->devices[0]));
    ORT_CUDA_TRY(cudaMemcpy(output_buffer, m->d_rgb, (size_t)w * h * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return ORT_OK;
    ORT_GUARD_END
}

// ---- progressive accumulation with dynamic chunk dispatch and checkpoints ------------------------------
int ort_progress_create(OrtMulti *m, const OrtCamera *camera, const OrtRenderParams *P, OrtProgress **out)
{
    ORT_GUARD_BEGIN
    if(!m || !camera || !P || !out) return ort::fail_with(ORT_ERR_ARG, "null argument");
    if(P->output_width <= 0 || P->output_height <= 0 || P->ray_per_pixel_count == 0) return ort::fail_with(ORT_ERR_ARG, "bad params");
    if(P->tile_min_x != 0 || P->tile_min_y != 0 || P->tile_one_past_max_x != P->output_width || P->tile_one_past_max_y != P->output_height)
        return ort::fail_with(ORT_ERR_ARG, "progressive renders cover whole images");
    *out = 0;
    int rc = multi_ensure(m, (size_t)P->output_width * P->output_height);
    if(rc != ORT_OK) return rc;
    OrtProgress *p = new OrtProgress();
    p->multi = m; p->camera = *camera; p->params = *P;
    if(p->params.chunk_spp == 0) p->params.chunk_spp = P->ray_per_pixel_count;
    p->n_chunks = chunk_count_of(&p->params);
    p->done.assign(p->n_chunks, 0);
    p->cursor = 0; p->dirty = false;
    for(size_t i = 0; i < m->devices.size(); ++i)
    {
        rc = zero_accum(m, i, P->output_width, P->output_height);
        if(rc == ORT_OK && cudaStreamSynchronize(m->streams[i]) != cudaSuccess) rc = ort::fail_with(ORT_ERR_CUDA, "stream synchronize failed");
        if(rc != ORT_OK) { delete p; return rc; }
    }
    *out = p;
    return ORT_OK;
    ORT_GUARD_END
}

int ort_progress_destroy(OrtProgress *p) { delete p; return ORT_OK; }

int ort_progress_state(const OrtProgress *p, uint32_t *chunks_done, uint32_t *chunks_total, uint32_t *spp_done)
{
    if(!p) return ort::fail_with(ORT_ERR_ARG, "null argument");
    uint32_t d = 0, s = 0;
    const uint32_t spp = p->params.ray_per_pixel_count, cs = p->params.chunk_spp;
    for(uint32_t c = 0; c < p->n_chunks; ++c)
        if(p->done[c]) { ++d; s += ((c + 1) * cs > spp) ? spp - c * cs : cs; }
    if(chunks_done) *chunks_done = d;
    if(chunks_total) *chunks_total = p->n_chunks;
    if(spp_done) *spp_done = s;
    return ORT_OK;
}

// every device pulls the next chunk not yet rendered from a shared counter until `max_chunks` more are done
// (0 = all that remain): a slow or busy GPU simply takes fewer chunks
int ort_progress_render(OrtProgress *p, uint32_t max_chunks, OrtRenderStats *stats)
{
    ORT_GUARD_BEGIN
    if(!p) return ort::fail_with(ORT_ERR_ARG, "null argument");
    OrtMulti *m = p->multi;
    std::vector<uint32_t> todo;
    for(uint32_t c = 0; c < p->n_chunks; ++c) if(!p->done[c]) todo.push_back(c);
    if(max_chunks && todo.size() > max_chunks) todo.resize(max_chunks);
    if(stats) memset(stats, 0, sizeof(*stats));
    if(todo.empty()) return ORT_OK;
    const uint32_t n = (uint32_t)m->devices.size();
    p->cursor = 0;
    std::vector<int> rcs(n, ORT_OK);
    std::vector<std::string> errs(n);
    std::vector<OrtRenderStats> sts(n);
    std::vector<std::vector<uint32_t>> mine(n);
    auto work = [&](uint32_t i)
    {
        memset(&sts[i], 0, sizeof(OrtRenderStats));
        cudaSetDevice(m->devices[i]);
        for(;;)
        {
            uint32_t k = p->cursor.fetch_add(1);
            if(k >= todo.size()) break;
            OrtRenderParams Q = p->params;
            Q.chunk_begin = todo[k]; Q.chunk_end = todo[k] + 1;
            OrtRenderStats one; memset(&one, 0, sizeof(one));
            rcs[i] = ort_render_accumulate_device(m->scenes[i], &p->camera, &Q, m->accum[i], m->streams[i], &one);
            if(rcs[i] != ORT_OK) { errs[i] = ort_last_error(); break; }
            float ms = sts[i].device_ms + one.device_ms;
            add_stats(&sts[i], one);
            sts[i].device_ms = ms;
            mine[i].push_back(todo[k]);
        }
    };
    std::vector<std::thread> threads;
    for(uint32_t i = 1; i < n; ++i) threads.emplace_back(work, i);
    work(0);
    for(auto &t : threads) t.join();
    // a chunk counts as done only if its device finished it; a failed device's other chunks were completed
    // (each call is synchronous), so they stay valid
    for(uint32_t i = 0; i < n; ++i) for(uint32_t c : mine[i]) p->done[c] = 1;
    if(n > 1) p->dirty = true;
    for(uint32_t i = 0; i < n; ++i)
    {
        if(stats) add_stats(stats, sts[i]);
        if(rcs[i] != ORT_OK) return ort::fail_with(rcs[i], "device " + std::to_string(m->devices[i]) + ": " + errs[i]);
    }
    return ORT_OK;
    ORT_GUARD_END
}

// folds the peers' sums into the root's framebuffer and clears them, so that the root holds the frame so far
static int progress_fold(OrtProgress *p, bool want_rgb, uint32_t spp_done)
{
    OrtMulti *m = p->multi;
    const int32_t w = p->params.output_width, h = p->params.output_height;
    int rc = multi_reduce(m, w, h, spp_done, want_rgb, false);
    if(rc != ORT_OK) return rc;
    if(p->dirty)
    {
        for(size_t i = 1; i < m->devices.size(); ++i)
        {
            rc = zero_accum(m, i, w, h);
            if(rc != ORT_OK) return rc;
            ORT_CUDA_TRY(cudaStreamSynchronize(m->streams[i]));
        }
        p->dirty = false;
    }
    return ORT_OK;
}

// image of the chunks rendered so far: sum / (samples per pixel done so far)
int ort_progress_resolve(OrtProgress *p, ort_v3 *output_buffer)
{
    ORT_GUARD_BEGIN
    if(!p || !output_buffer) return ort::fail_with(ORT_ERR_ARG, "null argument");
    uint32_t spp_done = 0;
    ort_progress_state(p, 0, 0, &spp_done);
    if(spp_done == 0) return ort::fail_with(ORT_ERR_ARG, "no chunk rendered yet");
    int rc = progress_fold(p, true, spp_done);
    if(rc != ORT_OK) return rc;
    OrtMulti *m = p->multi;
    ORT_CUDA_TRY(cudaSetDevice(m->devices[0]));
    ORT_CUDA_TRY(cudaMemcpy(output_buffer, m->d_rgb, (size_t)p->params.output_width * p->params.output_height * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return ORT_OK;
    ORT_GUARD_END
}

// Checkpoint file (little endian):
//   char[8]  "ORTPROG1"
//   u32      sizeof(OrtRenderParams), then the OrtRenderParams of the frame (seed schedule included)
//   u32      sizeof(OrtCamera), then the OrtCamera
//   u32      n_chunks, then n_chunks bytes: 1 = chunk already in the sums
//   i64      [height * width * 4] fixed-point sums (ORT_ACCUM_FRAC_BITS fractional bits, 4th lane unused)
int ort_progress_save(OrtProgress *p, const char *path)
{
    ORT_GUARD_BEGIN
    if(!p || !path) return ort::fail_with(ORT_ERR_ARG, "null argument");
    int rc = progress_fold(p, false, 1);
    if(rc != ORT_OK) return rc;
    OrtMulti *m = p->multi;
    const size_t n = (size_t)p->params.output_width * p->params.output_height * 4;
    std::vector<long long> host(n);
    ORT_CUDA_TRY(cudaSetDevice(m->devices[0]));
    ORT_CUDA_TRY(cudaMemcpy(host.data(), m->accum[0], n * sizeof(long long), cudaMemcpyDeviceToHost));
    FILE *f = fopen(path, "wb");
    if(!f) return ort::fail_with(ORT_ERR_IO, std::string("cannot write ") + path);
    uint32_t sp = (uint32_t)sizeof(OrtRenderParams), sc = (uint32_t)sizeof(OrtCamera);
    bool ok = fwrite("ORTPROG1", 1, 8, f) == 8 && fwrite(&sp, 4, 1, f) == 1 && fwrite(&p->params, sp, 1, f) == 1 &&
              fwrite(&sc, 4, 1, f) == 1 && fwrite(&p->camera, sc, 1, f) == 1 && fwrite(&p->n_chunks, 4, 1, f) == 1 &&
              (p->n_chunks == 0 || fwrite(p->done.data(), 1, p->n_chunks, f) == p->n_chunks) &&
              fwrite(host.data(), sizeof(long long), n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if(!ok) return ort::fail_with(ORT_ERR_IO, std::string("short write to ") + path);
    return ORT_OK;
    ORT_GUARD_END
}

int ort_progress_load(OrtMulti *m, const char *path, OrtProgress **out)
{
    ORT_GUARD_BEGIN
    if(!m || !path || !out) return ort::fail_with(ORT_ERR_ARG, "null argument");
    *out = 0;
    FILE *f = fopen(path, "rb");
    if(!f) return ort::fail_with(ORT_ERR_IO, std::string("cannot read ") + path);
    char magic[8]; uint32_t sp = 0, sc = 0, nc = 0;
    OrtRenderParams P; OrtCamera cam;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "ORTPROG1", 8) == 0 && fread(&sp, 4, 1, f) == 1 && sp == sizeof(OrtRenderParams) &&
              fread(&P, sp, 1, f) == 1 && fread(&sc, 4, 1, f) == 1 && sc == sizeof(OrtCamera) && fread(&cam, sc, 1, f) == 1 && fread(&nc, 4, 1, f) == 1;
    if(!ok || P.output_width <= 0 || P.output_height <= 0 || (int64_t)P.output_width * P.output_height >= (1ll << 29) || nc != chunk_count_of(&P))
    {
        fclose(f);
        return ort::fail_with(ORT_ERR_PARSE, std::string("not a checkpoint of this library: ") + path);
    }
    OrtProgress *p = 0;
    int rc = ort_progress_create(m, &cam, &P, &p);
    if(rc != ORT_OK) { fclose(f); return rc; }
    const size_t n = (size_t)P.output_width * P.output_height * 4;
    std::vector<long long> host(n);
    ok = (nc == 0 || fread(p->done.data(), 1, nc, f) == nc) && fread(host.data(), sizeof(long long), n, f) == n;
    fclose(f);
    if(!ok) { delete p; return ort::fail_with(ORT_ERR_PARSE, std::string("truncated checkpoint: ") + path); }
    for(uint32_t c = 0; c < nc; ++c) p->done[c] = p->done[c] ? 1 : 0;
    cudaError_t e = cudaSetDevice(m->devices[0]);
    if(e == cudaSuccess) e = cudaMemcpy(m->accum[0], host.data(), n * sizeof(long long), cudaMemcpyHostToDevice);
    if(e != cudaSuccess) { delete p; return ort::fail_with(ORT_ERR_CUDA, cudaGetErrorString(e)); }
    *out = p;
    return ORT_OK;
    ORT_GUARD_END
}

} // extern "C"
