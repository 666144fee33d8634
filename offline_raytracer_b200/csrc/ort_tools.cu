// ort_tools.cu -- entry points around the hot path that are not the hot path:
//   * roofline denominators measured on the box (L2 read bandwidth; FP32 lives in ort_b200.cu),
//   * the DEVICE differential harness: the device build of csrc/core_math.h (intersectors) and
//     csrc/path.h (BSDF) run on explicit inputs, so that the reference's golden vectors
//     (tests/golden/reference_vectors.npz) are checked against the very code the render kernels
//     inline -- not only against its host build,
//   * the explicit ray buffers of BASELINE config 2 generated on the device (reference camera model,
//     code/ray.cpp:1215-1246, and uniform incoherent rays), one xorshift stream per ray.
#include <vector>

#include "ort_internal.h"
#include "path.h"
#include "scene_flatten.h"

using namespace ort;

namespace {

int pick_device(int device)
{
    int n = 0;
    if(ort_device_count(&n) != ORT_OK) return ORT_ERR_CUDA;
    if(device < 0 || device >= n) return fail_with(ORT_ERR_ARG, "device ordinal out of range");
    ORT_CUDA_TRY(cudaSetDevice(device));
    return ORT_OK;
}

// ---- L2 read bandwidth ---------------------------------------------------------------------------
// Every block streams the whole buffer (chunks of 256 x 16 B, starting at a block-dependent offset) with
// ld.global.cg -- cached in L2 only, the way a wide node that misses L1 is served.  The buffer is far
// smaller than the 126 MB L2 and far larger than one SM's L1, and every byte is read gridDim.x times, so
// after the first pass all traffic is L2 -> SM.
__global__ void __launch_bounds__(256)
k_l2_read(const uint4 *__restrict__ buf, uint32_t n_chunks, uint32_t passes, uint32_t *sink)
{
    uint32_t acc = 0u;
    const uint32_t total = n_chunks * passes;
    uint32_t c = (blockIdx.x * 7919u) % n_chunks;
    for(uint32_t it = 0; it < total; it += 8u)
    {
        uint4 v[8];
#pragma unroll
        for(int k = 0; k < 8; ++k)
        {
            uint32_t cc = c + (uint32_t)k; if(cc >= n_chunks) cc -= n_chunks;
            const uint4 *p = buf + (size_t)cc * 256u + threadIdx.x;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p));
        }
#pragma unroll
        for(int k = 0; k < 8; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
        c += 8u; if(c >= n_chunks) c -= n_chunks;
    }
    if(acc == 0x12345678u) sink[0] = acc;        // never true for the zero-filled buffer; keeps the loads alive
}

// ---- device differential harness -------------------------------------------------------------------
// one case per thread; `in` rows: triangle v0 v1 v2 o d (15) | sphere c r o d (10) | box min max o d (12) |
// cylinder base rot(9) |axis| r o d (20, rotation formed on the host exactly as the flattener does);
// out rows: t, n.x, n.y, n.z (unnormalised, as IntersectionTestResult), inner flag
__global__ void k_st_intersect(uint32_t kind, uint32_t n, const float *__restrict__ in, float *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    exact::Hit h; int inner = 0;
    if(kind == PRIM_TRIANGLE)
    {
        const float *p = in + 15u * i;
        h = exact::triangle(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]), mk3(p[9], p[10], p[11]), mk3(p[12], p[13], p[14]));
    }
    else if(kind == PRIM_SPHERE)
    {
        const float *p = in + 10u * i;
        h = exact::sphere(mk3(p[0], p[1], p[2]), p[3], mk3(p[4], p[5], p[6]), mk3(p[7], p[8], p[9]), &inner);
    }
    else if(kind == PRIM_AAB)
    {
        const float *p = in + 12u * i;
        h = exact::aab(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]), mk3(p[9], p[10], p[11]));
    }
    else
    {
        const float *p = in + 20u * i;
        exact::m3 rot; rot.r0 = mk3(p[3], p[4], p[5]); rot.r1 = mk3(p[6], p[7], p[8]); rot.r2 = mk3(p[9], p[10], p[11]);
        h = exact::cylinder_pre(mk3(p[0], p[1], p[2]), rot, p[12], p[13], mk3(p[14], p[15], p[16]), mk3(p[17], p[18], p[19]));
    }
    float *o = out + 5u * i;
    o[0] = h.t; o[1] = h.n.x; o[2] = h.n.y; o[3] = h.n.z; o[4] = (float)inner;
}

// BSDF: sample_brdf / pdf_brdf / eval_scattering of path.h on explicit tuples; material i = DevMaterial i
__global__ void k_st_bsdf(uint32_t n, const q4 *__restrict__ materials, const float *__restrict__ N, const float *__restrict__ wo,
                          const float *__restrict__ wi, const uint32_t *__restrict__ state, const float *__restrict__ dist, float roughness,
                          float *__restrict__ sample_wi, int32_t *__restrict__ is_t, uint32_t *__restrict__ state_after,
                          float *__restrict__ pdf, float *__restrict__ eval)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    PathConsts c; c.materials = materials;
    Mat m = load_material(c, i);
    f3 n3 = mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]);
    f3 o3 = mk3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
    f3 i3 = mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
    uint32_t s = state[i];
    SampleBRDF sb = sample_brdf(&s, n3, o3, roughness, m);
    sample_wi[3 * i] = sb.wi.x; sample_wi[3 * i + 1] = sb.wi.y; sample_wi[3 * i + 2] = sb.wi.z;
    is_t[i] = sb.is_transmission; state_after[i] = s;
    pdf[i] = pdf_brdf(n3, i3, o3, roughness, m);
    f3 e = eval_scattering(n3, i3, o3, m, roughness, dist[i]);
    eval[3 * i] = e.x; eval[3 * i + 1] = e.y; eval[3 * i + 2] = e.z;
}


// RNG (code/random.h) on the device: per seed 64 xorshift states, then -- each from the seed again -- 8 x
// random_between_0_1, 8 x random_between(0, hi), 8 x random_between_u32(0, n)
__global__ void k_st_rng(uint32_t n, const uint32_t *__restrict__ seeds, float hi, uint32_t one_past_max,
                         uint32_t *__restrict__ states, float *__restrict__ f01, float *__restrict__ between, uint32_t *__restrict__ u32)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    uint32_t x = seeds[i];
    for(int k = 0; k < 64; ++k) { xor_shift_32(&x); states[64u * i + k] = x; }
    x = seeds[i];
    for(int k = 0; k < 8; ++k) f01[8u * i + k] = random_between_0_1(&x);
    x = seeds[i];
    for(int k = 0; k < 8; ++k) between[8u * i + k] = random_between(&x, 0.0f, hi);
    x = seeds[i];
    for(int k = 0; k < 8; ++k) u32[8u * i + k] = random_between_u32(&x, 0u, one_past_max);
}
// sample_random_lights (ray.cpp:537-601) as the render kernels run it: only its RNG side effects survive
__global__ void k_st_light_pick(uint32_t n, PathConsts pc, uint32_t state, uint32_t *__restrict__ out)
{
    if(blockIdx.x != 0 || threadIdx.x != 0) return;
    for(uint32_t k = 0; k < n; ++k) { sample_random_lights_rng(pc, &state); out[k] = state; }
}

// ---- ray buffers of BASELINE config 2 ---------------------------------------------------------------
// ray i draws from its own xorshift stream ort_stream_seed(seed, i, 0).  Coherent: pixel (i % w, i / w) of a
// w x h grid, lens sample as ray.cpp:1232-1237 (generate_primary of path.h, the render kernels' own code).
__global__ void k_gen_camera_rays(PathConsts pc, uint32_t seed, unsigned long long n, uint32_t w,
                                  float *__restrict__ origins, float *__restrict__ dirs)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    int x = (int)(i % w), y = (int)(i / w);
    Path p; p.series = ort_stream_seed(seed, (uint32_t)i, (uint32_t)(i >> 32));
    generate_primary(pc, pixel_focal_point(pc, x, y), &p);
    origins[3 * i] = p.origin.x; origins[3 * i + 1] = p.origin.y; origins[3 * i + 2] = p.origin.z;
    dirs[3 * i] = p.dir.x; dirs[3 * i + 1] = p.dir.y; dirs[3 * i + 2] = p.dir.z;
}

// incoherent: origin uniform in the box, direction uniform on the sphere (z uniform in [-1, 1], azimuth uniform)
__global__ void k_gen_random_rays(f3 lo, f3 hi, uint32_t seed, unsigned long long n, float *__restrict__ origins, float *__restrict__ dirs)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    uint32_t s = ort_stream_seed(seed, (uint32_t)i, (uint32_t)(i >> 32));
    float ox = lo.x + (hi.x - lo.x) * random_between_0_1(&s);
    float oy = lo.y + (hi.y - lo.y) * random_between_0_1(&s);
    float oz = lo.z + (hi.z - lo.z) * random_between_0_1(&s);
    float z = 2.0f * random_between_0_1(&s) - 1.0f;
    float phi = 2.0f * ORT_PI_32 * random_between_0_1(&s);
    float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    f3 d = normalize(mk3(r * cosf(phi), r * sinf(phi), z));
    if(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f) d = mk3(0.0f, 0.0f, 1.0f);
    origins[3 * i] = ox; origins[3 * i + 1] = oy; origins[3 * i + 2] = oz;
    dirs[3 * i] = d.x; dirs[3 * i + 1] = d.y; dirs[3 * i + 2] = d.z;
}


// ---- mesh bake (code/macos_main.mm:382-413) ----------------------------------------------
// v *= scale; v = quaternion_rotation(q, quaternion_rotation((0,1,0), degree, v)); v += translate -- the IEEE
// operations of code/math.h:746-795 in their order, no contraction, so the baked vertices have the bits the
// reference's (and this library's host) bake produces.  The axis quaternion (cosf / sinf of the half angle) is
// a per-mesh constant and is formed on the host, like every libm value that reaches the geometry.
__device__ __forceinline__ f3 quat_rotate_dev(float4 q, f3 v)          // math.h:774-795 (q = x, y, z, w)
{
    float m00 = 1.0f - 2 * q.y * q.y - 2 * q.z * q.z;
    float m01 = 2 * q.x * q.y - 2 * q.w * q.z;
    float m02 = 2 * q.x * q.z + 2 * q.w * q.y;
    float m10 = 2 * q.x * q.y + 2 * q.w * q.z;
    float m11 = 1.0f - 2 * q.x * q.x - 2 * q.z * q.z;
    float m12 = 2 * q.y * q.z - 2 * q.w * q.x;
    float m20 = 2 * q.x * q.z - 2 * q.w * q.y;
    float m21 = 2 * q.y * q.z + 2 * q.w * q.x;
    float m22 = 1 - 2 * q.x * q.x - 2 * q.y * q.y;
    return mk3(m00 * v.x + m01 * v.y + m02 * v.z, m10 * v.x + m11 * v.y + m12 * v.z, m20 * v.x + m21 * v.y + m22 * v.z);
}
// order-preserving map float -> uint32 for atomicMin / atomicMax
__device__ __forceinline__ uint32_t f_ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float f_unord(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

// pass 1: bake + value bounds (ord[0..2] = min, ord[3..5] = max, as ordered integers)
__global__ void k_bake_mesh(uint32_t n, const float *__restrict__ in, float *__restrict__ out, float scale, float4 q_axis, float4 q,
                            f3 translate, uint32_t *__restrict__ ord)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    f3 v = mk3(in[3 * i] * scale, in[3 * i + 1] * scale, in[3 * i + 2] * scale);
    v = quat_rotate_dev(q, quat_rotate_dev(q_axis, v));
    v = v + translate;
    out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
    atomicMin(&ord[0], f_ord(v.x)); atomicMin(&ord[1], f_ord(v.y)); atomicMin(&ord[2], f_ord(v.z));
    atomicMax(&ord[3], f_ord(v.x)); atomicMax(&ord[4], f_ord(v.y)); atomicMax(&ord[5], f_ord(v.z));
}
// pass 2: the reference folds the vertices sequentially with `(a < b) ? a : b` (types.h:50-51), which keeps the
// LATER of two equal values -- visible only when the bound is a zero whose sign differs between vertices.
// last[k] = highest vertex index whose component EQUALS the bound (IEEE ==, so +0 == -0).
__global__ void k_bake_last(uint32_t n, const float *__restrict__ out, const uint32_t *__restrict__ ord, uint32_t *__restrict__ last)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
        float v = out[3 * i + k];
        if(v == f_unord(ord[k])) atomicMax(&last[k], i + 1u);
        if(v == f_unord(ord[3 + k])) atomicMax(&last[3 + k], i + 1u);
    }
}

struct DevBuf
{
    void *p = 0;
    ~DevBuf() { if(p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T *as() { return (T *)p; }
};

} // namespace

namespace ort { void material_constants_public(DevMaterial *dm); }

extern "C" {

int ort_measure_l2_bandwidth(int device, uint32_t buffer_mib, float *gb_per_s)
{
    if(!gb_per_s) return fail_with(ORT_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    if(buffer_mib == 0) buffer_mib = 32;
    if(buffer_mib > 96) return fail_with(ORT_ERR_ARG, "buffer must stay well inside the 126 MB L2 (<= 96 MiB)");
    int sms = 0;
    ORT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const size_t bytes = (size_t)buffer_mib << 20;
    const uint32_t n_chunks = (uint32_t)(bytes / (256 * 16));
    const uint32_t passes = 2;
    DevBuf buf, sink;
    ORT_CUDA_TRY(buf.alloc(bytes));
    ORT_CUDA_TRY(sink.alloc(16));
    ORT_CUDA_TRY(cudaMemset(buf.p, 0, bytes));
    cudaEvent_t e0, e1;
    ORT_CUDA_TRY(cudaEventCreate(&e0)); ORT_CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms * 8;
    float best = 0.f;
    for(int rep = 0; rep < 6; ++rep)
    {
        ORT_CUDA_TRY(cudaEventRecord(e0, 0));
        k_l2_read<<<blocks, 256>>>(buf.as<uint4>(), n_chunks, passes, sink.as<uint32_t>());
        ORT_CUDA_TRY(cudaEventRecord(e1, 0));
        ORT_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        ORT_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double moved = (double)blocks * (double)n_chunks * passes * 256.0 * 16.0;
        float gbs = (float)(moved / (ms * 1e-3) / 1e9);
        if(rep > 0 && gbs > best) best = gbs;       // rep 0 warms the L2
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gb_per_s = best;
    return ORT_OK;
}

int ort_selftest_intersect(int device, uint32_t kind, uint32_t n, const float *cases, float *out)
{
    ORT_GUARD_BEGIN
    if(n && (!cases || !out)) return fail_with(ORT_ERR_ARG, "null argument");
    if(kind > PRIM_CYLINDER) return fail_with(ORT_ERR_ARG, "kind: 0 triangle, 1 sphere, 2 box, 3 cylinder");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    if(n == 0) return ORT_OK;
    const uint32_t in_stride[4] = { 15u, 10u, 12u, 13u }, dev_stride[4] = { 15u, 10u, 12u, 20u };
    std::vector<float> staged;
    const float *src = cases;
    if(kind == PRIM_CYLINDER)
    {
        // base axis r o d  ->  base rot |axis| r o d: rotation_matrix_along_z (ray.cpp:8-33) depends on the
        // cylinder only and is evaluated on the host with the same IEEE operations, as the flattener does
        staged.resize((size_t)n * 20u);
        for(uint32_t i = 0; i < n; ++i)
        {
            const float *p = cases + 13u * i; float *q = staged.data() + 20u * i;
            f3 axis = mk3(p[3], p[4], p[5]);
            exact::m3 rot = exact::rotation_matrix_along_z(axis);
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
            q[3] = rot.r0.x; q[4] = rot.r0.y; q[5] = rot.r0.z; q[6] = rot.r1.x; q[7] = rot.r1.y; q[8] = rot.r1.z;
            q[9] = rot.r2.x; q[10] = rot.r2.y; q[11] = rot.r2.z;
            q[12] = length(axis); q[13] = p[6];
            for(int k = 0; k < 6; ++k) q[14 + k] = p[7 + k];
        }
        src = staged.data();
    }
    (void)in_stride;
    DevBuf d_in, d_out;
    ORT_CUDA_TRY(d_in.alloc((size_t)n * dev_stride[kind] * sizeof(float)));
    ORT_CUDA_TRY(d_out.alloc((size_t)n * 5 * sizeof(float)));
    ORT_CUDA_TRY(cudaMemcpy(d_in.p, src, (size_t)n * dev_stride[kind] * sizeof(float), cudaMemcpyHostToDevice));
    k_st_intersect<<<(n + 127u) / 128u, 128>>>(kind, n, d_in.as<float>(), d_out.as<float>());
    ORT_CUDA_TRY(cudaGetLastError());
    ORT_CUDA_TRY(cudaMemcpy(out, d_out.p, (size_t)n * 5 * sizeof(float), cudaMemcpyDeviceToHost));
    return ORT_OK;
    ORT_GUARD_END
}

int ort_selftest_bsdf(int device, uint32_t n, const float *mat10, const float *N, const float *wo, const float *wi,
                      const uint32_t *state, const float *dist, float roughness,
                      float *sample_wi, int32_t *is_transmission, uint32_t *state_after, float *pdf, float *eval)
{
    ORT_GUARD_BEGIN
    if(n && (!mat10 || !N || !wo || !wi || !state || !dist || !sample_wi || !is_transmission || !state_after || !pdf || !eval))
        return fail_with(ORT_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    if(n == 0) return ORT_OK;
    std::vector<DevMaterial> mats(n);
    memset(mats.data(), 0, mats.size() * sizeof(DevMaterial));
    for(uint32_t i = 0; i < n; ++i)
    {
        const float *m = mat10 + 10u * i;
        DevMaterial &dm = mats[i];
        for(int k = 0; k < 3; ++k) { dm.diffuse[k] = m[k]; dm.specular[k] = m[3 + k]; dm.transmission[k] = m[6 + k]; }
        dm.ior = m[9];
        material_constants_public(&dm);
    }
    DevBuf d_m, d_N, d_wo, d_wi, d_st, d_dist, o_wi, o_t, o_st, o_pdf, o_ev;
    const size_t v3 = (size_t)n * 3 * sizeof(float), s1 = (size_t)n * sizeof(float);
    ORT_CUDA_TRY(d_m.alloc(mats.size() * sizeof(DevMaterial)));
    ORT_CUDA_TRY(d_N.alloc(v3)); ORT_CUDA_TRY(d_wo.alloc(v3)); ORT_CUDA_TRY(d_wi.alloc(v3));
    ORT_CUDA_TRY(d_st.alloc(s1)); ORT_CUDA_TRY(d_dist.alloc(s1));
    ORT_CUDA_TRY(o_wi.alloc(v3)); ORT_CUDA_TRY(o_t.alloc(s1)); ORT_CUDA_TRY(o_st.alloc(s1)); ORT_CUDA_TRY(o_pdf.alloc(s1)); ORT_CUDA_TRY(o_ev.alloc(v3));
    ORT_CUDA_TRY(cudaMemcpy(d_m.p, mats.data(), mats.size() * sizeof(DevMaterial), cudaMemcpyHostToDevice));
    ORT_CUDA_TRY(cudaMemcpy(d_N.p, N, v3, cudaMemcpyHostToDevice));
    ORT_CUDA_TRY(cudaMemcpy(d_wo.p, wo, v3, cudaMemcpyHostToDevice));
    ORT_CUDA_TRY(cudaMemcpy(d_wi.p, wi, v3, cudaMemcpyHostToDevice));
    ORT_CUDA_TRY(cudaMemcpy(d_st.p, state, s1, cudaMemcpyHostToDevice));
    ORT_CUDA_TRY(cudaMemcpy(d_dist.p, dist, s1, cudaMemcpyHostToDevice));
    k_st_bsdf<<<(n + 127u) / 128u, 128>>>(n, d_m.as<q4>(), d_N.as<float>(), d_wo.as<float>(), d_wi.as<float>(), d_st.as<uint32_t>(),
                                          d_dist.as<float>(), roughness, o_wi.as<float>(), o_t.as<int32_t>(), o_st.as<uint32_t>(),
                                          o_pdf.as<float>(), o_ev.as<float>());
    ORT_CUDA_TRY(cudaGetLastError());
    ORT_CUDA_TRY(cudaMemcpy(sample_wi, o_wi.p, v3, cudaMemcpyDeviceToHost));
    ORT_CUDA_TRY(cudaMemcpy(is_transmission, o_t.p, s1, cudaMemcpyDeviceToHost));
    ORT_CUDA_TRY(cudaMemcpy(state_after, o_st.p, s1, cudaMemcpyDeviceToHost));
    ORT_CUDA_TRY(cudaMemcpy(pdf, o_pdf.p, s1, cudaMemcpyDeviceToHost));
    ORT_CUDA_TRY(cudaMemcpy(eval, o_ev.p, v3, cudaMemcpyDeviceToHost));
    return ORT_OK;
    ORT_GUARD_END
}


int ort_selftest_rng(int device, uint32_t n, const uint32_t *seeds, float between_hi, uint32_t u32_one_past_max,
                     uint32_t *states64, float *f01_8, float *between_8, uint32_t *u32_8)
{
    ORT_GUARD_BEGIN
    if(n && (!seeds || !states64 || !f01_8 || !between_8 || !u32_8)) return fail_with(ORT_ERR_ARG, "null argument");
    if(u32_one_past_max == 0) return fail_with(ORT_ERR_ARG, "u32_one_past_max must be > 0");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    if(n == 0) return ORT_OK;
    DevBuf d_s, d_st, d_f, d_b, d_u;
    ORT_CUDA_TRY(d_s.alloc(n * 4)); ORT_CUDA_TRY(d_st.alloc((size_t)n * 64 * 4)); ORT_CUDA_TRY(d_f.alloc((size_t)n * 8 * 4));
    ORT_CUDA_TRY(d_b.alloc((size_t)n * 8 * 4)); ORT_CUDA_TRY(d_u.alloc((size_t)n * 8 * 4));
    ORT_CUDA_TRY(cudaMemcpy(d_s.p, seeds, n * 4, cudaMemcpyHostToDevice));
    k_st_rng<<<(n + 63u) / 64u, 64>>>(n, d_s.as<uint32_t>(), between_hi, u32_one_past_max, d_st.as<uint32_t>(), d_f.as<float>(), d_b.as<float>(), d_u.as<uint32_t>());
    ORT_CUDA_TRY(cudaGetLastError());
    ORT_CUDA_TRY(cudaMemcpy(states64, d_st.p, (size_t)n * 64 * 4, cudaMemcpyDeviceToHost));
    ORT_CUDA_TRY(cudaMemcpy(f01_8, d_f.p, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    ORT_CUDA_TRY(cudaMemcpy(between_8, d_b.p, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    ORT_CUDA_TRY(cudaMemcpy(u32_8, d_u.p, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    return ORT_OK;
    ORT_GUARD_END
}

int ort_selftest_light_pick(int device, uint32_t light_count, const uint8_t *light_is_sphere, uint32_t state, uint32_t n, uint32_t *states_out)
{
    ORT_GUARD_BEGIN
    if((light_count && !light_is_sphere) || (n && !states_out)) return fail_with(ORT_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    if(n == 0) return ORT_OK;
    DevBuf d_l, d_o;
    ORT_CUDA_TRY(d_l.alloc(light_count)); ORT_CUDA_TRY(d_o.alloc((size_t)n * 4));
    if(light_count) ORT_CUDA_TRY(cudaMemcpy(d_l.p, light_is_sphere, light_count, cudaMemcpyHostToDevice));
    PathConsts pc; memset(&pc, 0, sizeof(pc));
    pc.light_count = light_count; pc.light_is_sphere = d_l.as<uint8_t>();
    k_st_light_pick<<<1, 32>>>(n, pc, state, d_o.as<uint32_t>());
    ORT_CUDA_TRY(cudaGetLastError());
    ORT_CUDA_TRY(cudaMemcpy(states_out, d_o.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return ORT_OK;
    ORT_GUARD_END
}

int ort_bake_mesh(int device, uint32_t vertex_count, const ort_v3 *vertices_in, ort_v3 *vertices_out,
                  float scale, float degree, ort_v4 quaternion, ort_v3 translate, ort_v3 *aabb_min, ort_v3 *aabb_max)
{
    ORT_GUARD_BEGIN
    if(vertex_count && (!vertices_in || !vertices_out)) return fail_with(ORT_ERR_ARG, "null argument");
    if(!aabb_min || !aabb_max) return fail_with(ORT_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    // the empty fold: (FLT_MAX, FLT_MIN) -- FLT_MIN being the smallest POSITIVE float (macos_main.mm:383)
    aabb_min->x = aabb_min->y = aabb_min->z = FLT_MAX;
    aabb_max->x = aabb_max->y = aabb_max->z = FLT_MIN;
    if(vertex_count == 0) return ORT_OK;
    // quaternion_rotation(v3 axis = (0,1,0), rad, v), math.h:746-772: the quaternion of the axis rotation, on the host
    const float rad = 0.0174533f * degree;
    float4 qa; qa.w = cosf(rad / 2); qa.x = 0.0f * sinf(rad / 2); qa.y = 1.0f * sinf(rad / 2); qa.z = 0.0f * sinf(rad / 2);
    float4 q; q.x = quaternion.x; q.y = quaternion.y; q.z = quaternion.z; q.w = quaternion.w;
    DevBuf d_in, d_out, d_ord, d_last;
    const size_t bytes = (size_t)vertex_count * 3 * sizeof(float);
    ORT_CUDA_TRY(d_in.alloc(bytes)); ORT_CUDA_TRY(d_out.alloc(bytes)); ORT_CUDA_TRY(d_ord.alloc(6 * 4)); ORT_CUDA_TRY(d_last.alloc(6 * 4));
    uint32_t init[6] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u };
    ORT_CUDA_TRY(cudaMemcpy(d_in.p, vertices_in, bytes, cudaMemcpyHostToDevice));
    ORT_CUDA_TRY(cudaMemcpy(d_ord.p, init, sizeof(init), cudaMemcpyHostToDevice));
    ORT_CUDA_TRY(cudaMemset(d_last.p, 0, 6 * 4));
    const unsigned grid = (vertex_count + 255u) / 256u;
    k_bake_mesh<<<grid, 256>>>(vertex_count, d_in.as<float>(), d_out.as<float>(), scale, qa, q, mk3(translate.x, translate.y, translate.z), d_ord.as<uint32_t>());
    k_bake_last<<<grid, 256>>>(vertex_count, d_out.as<float>(), d_ord.as<uint32_t>(), d_last.as<uint32_t>());
    ORT_CUDA_TRY(cudaGetLastError());
    ORT_CUDA_TRY(cudaMemcpy(vertices_out, d_out.p, bytes, cudaMemcpyDeviceToHost));
    uint32_t last[6];
    ORT_CUDA_TRY(cudaMemcpy(last, d_last.p, sizeof(last), cudaMemcpyDeviceToHost));
    float *mn = &aabb_min->x, *mx = &aabb_max->x;
    for(int k = 0; k < 3; ++k)
    {
        // every vertex is finite here or the bound has no last index: then the fold's start value stands
        if(last[k]) { float v = (&vertices_out[last[k] - 1u].x)[k]; if(!(FLT_MAX < v)) mn[k] = v; }
        if(last[3 + k]) { float v = (&vertices_out[last[3 + k] - 1u].x)[k]; if(!(FLT_MIN > v)) mx[k] = v; }
    }
    return ORT_OK;
    ORT_GUARD_END
}

int ort_generate_camera_rays_device(int device, const OrtCamera *cam, const OrtRenderParams *P, uint32_t seed, uint64_t n,
                                    float *origins_device, float *dirs_device, void *stream)
{
    if(!cam || !P || (n && (!origins_device || !dirs_device))) return fail_with(ORT_ERR_ARG, "null argument");
    if(P->output_width <= 0 || P->output_height <= 0) return fail_with(ORT_ERR_ARG, "grid size must be positive");
    if(n > (uint64_t)P->output_width * (uint64_t)P->output_height) return fail_with(ORT_ERR_ARG, "more rays than grid cells");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    if(n == 0) return ORT_OK;
    PathConsts c; memset(&c, 0, sizeof(c));
    c.cam_p = mk3(cam->p.x, cam->p.y, cam->p.z);
    c.cam_x = mk3(cam->x_axis.x, cam->x_axis.y, cam->x_axis.z);
    c.cam_y = mk3(cam->y_axis.x, cam->y_axis.y, cam->y_axis.z);
    c.cam_z = mk3(cam->z_axis.x, cam->z_axis.y, cam->z_axis.z);
    c.focal_length = length(c.cam_p - mk3(P->focus_target[0], P->focus_target[1], P->focus_target[2]));
    c.aperture_radius = P->aperture_radius; c.lens_z_offset = P->lens_z_offset;
    c.width = P->output_width; c.height = P->output_height;
    k_gen_camera_rays<<<(unsigned)((n + 255ull) / 256ull), 256, 0, (cudaStream_t)stream>>>(c, seed, n, (uint32_t)P->output_width,
                                                                                             origins_device, dirs_device);
    ORT_CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

int ort_generate_random_rays_device(int device, const float box_min[3], const float box_max[3], uint32_t seed, uint64_t n,
                                    float *origins_device, float *dirs_device, void *stream)
{
    if(!box_min || !box_max || (n && (!origins_device || !dirs_device))) return fail_with(ORT_ERR_ARG, "null argument");
    int rc = pick_device(device);
    if(rc != ORT_OK) return rc;
    if(n == 0) return ORT_OK;
    k_gen_random_rays<<<(unsigned)((n + 255ull) / 256ull), 256, 0, (cudaStream_t)stream>>>(
        mk3(box_min[0], box_min[1], box_min[2]), mk3(box_max[0], box_max[1], box_max[2]), seed, n, origins_device, dirs_device);
    ORT_CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

} // extern "C"
