// path.h -- the integrator's per-sample logic ("generate" and "shade"), the BSDF,
// the light pick and the xorshift RNG, written once for device and host (host =
// tests/sim only).  Replaces, function by function:
//
//   xor_shift_32, random_between_0_1, random_between, random_between_u32,
//   random_spherical_coordinate                    code/random.h:5-38,47-53,74-81,99-117
//   sample_random_lights                           code/ray.cpp:537-601
//   fresnel, ggx_distribution, geometry,
//   get_radicand, get_beer_n, eval_scattering      code/ray.cpp:825-1005
//   pdf_brdf                                       code/ray.cpp:1007-1063
//   sample_lobe, sample_brdf                       code/ray.cpp:1065-1161
//   the per-sample body of tiled_raytrace_bvh      code/ray.cpp:1211-1426
//
// Everything keeps the reference's operation order and its quirks (they are the
// specification): pdf_brdf's `pt = ps` default and doubled
// no*wo_dot_m term, fresnel as written, absolute(+0) = -0, pdf/eval evaluated at
// the NEW hit with the OLD wo, the light pick whose only effect is on the RNG.
// The sample stream draws in the reference's order:
//   [2] lens angle -> first hit -> { [1] roulette -> [1 (+4 if sphere)] light
//   pick -> [3] e0,e1,choice -> hit -> ... }.
// Compile with contraction off (see core_math.h).  sinf/cosf/atan2f/powf/logf
// come from the platform's libm (CUDA's on the device), which is why converged
// images, not sample streams, are compared with the CPU reference.
#pragma once

#include "bvh.h"
#include "ort_b200.h"

namespace ort {

#define ORT_EULER 2.71828182845904523536028747135266249f   // ray.cpp:4

// Out-of-line building blocks: the BSDF code calls these many times (normalize x12,
// geometry x4, ggx_distribution x3, ...); inlining every copy makes SHADE ~77 KB of SASS,
// more than the instruction caches hold (ncu: icc hit rate 63 %, "no_instruction" the top
// stall).  One shared copy each keeps the kernel resident.
#if defined(__CUDA_ARCH__) && !defined(ORT_SHADE_INLINE_ALL)
#define ORT_HD_BIG __device__ __noinline__
#else
#define ORT_HD_BIG ORT_HD
#endif
ORT_HD_BIG f3 nrm(f3 a) { return normalize(a); }

// ---- RNG -------------------------------------------------------------------
ORT_HD void xor_shift_32(uint32_t *s)                                    // random.h:5-16
{
    uint32_t x = *s;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x >> 5;      // sic: right shift
    *s = x;
}
ORT_HD float random_between_0_1(uint32_t *s)                             // random.h:32-38
{
    xor_shift_32(s);
    return (float)*s / (float)UINT32_MAX;
}
ORT_HD float random_between(uint32_t *s, float mn, float mx)             // random.h:47-53 (two steps)
{
    xor_shift_32(s);
    return mn + (mx - mn) * random_between_0_1(s);
}
ORT_HD uint32_t random_between_u32(uint32_t *s, uint32_t mn, uint32_t one_past_max)   // random.h:75-81
{
    xor_shift_32(s);
    return (uint32_t)(*s % (one_past_max - mn) + mn);
}

// ---- per-launch constants ----------------------------------------------------
struct PathConsts
{
    f3 cam_p, cam_x, cam_y, cam_z;     // Camera, code/ray.h:42-49 (axes pre-scaled, macos_main.mm:552-556)
    float focal_length;                // |cam.p - focus_target|, ray.cpp:1198
    float aperture_radius;             // ray.cpp:1199
    float lens_z_offset;               // ray.cpp:1234
    float roughness;                   // ray.cpp:1194
    float eps;                         // dont_get_too_close_epsilon, ray.cpp:1196
    float rr;                          // russian_roulette_value
    int32_t width, height;
    uint32_t light_count;
    const uint8_t *light_is_sphere;    // one byte per light-list entry
    const q4 *materials;               // DevMaterial[] viewed as q4[6*n]
};

struct Mat
{
    f3 Kd, Ks, Kt, emit;
    float ior;
    int is_light;
    // per-material constants formed once by the flattener (scene_flatten.h, DevMaterial) with the
    // operations of ray.cpp:1013-1019 / 940 -- the values pdf_brdf, sample_brdf and
    // eval_scattering would otherwise recompute at every bounce
    float pd_c, ps_c, pt_c;
    f3 Ed;
    uint32_t lobes;                    // bit 0/1/2: length_square(Kd/Ks/Kt) > 0
};
// DevMaterial is 96 bytes at a 96-byte stride from a 256-byte-aligned base: three 256-bit read-only loads on the
// device (sm_100: LDG.E.256.CONSTANT) instead of six 128-bit ones (measured variant, profiles/README.md)
#ifndef ORT_MAT_LD256
#define ORT_MAT_LD256 0
#endif
ORT_HD Mat load_material(const PathConsts &c, uint32_t index)
{
    const q4 *p = c.materials + 6u * index;
#if ORT_MAT_LD256 && defined(__CUDA_ARCH__)
    q4 a, b, t, e, k, d;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w), "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w) : "l"(p + 2));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(k.x), "=f"(k.y), "=f"(k.z), "=f"(k.w), "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(p + 4));
#else
    q4 a = ldq(p), b = ldq(p + 1), t = ldq(p + 2), e = ldq(p + 3), k = ldq(p + 4), d = ldq(p + 5);
#endif
    Mat m;
    m.Kd = q3(a); m.is_light = (int)f2u(a.w);
    m.Ks = q3(b); m.ior = b.w;
    m.Kt = q3(t); m.lobes = f2u(t.w);
    m.emit = q3(e);
    m.pd_c = k.x; m.ps_c = k.y; m.pt_c = k.z;
    m.Ed = q3(d);
    return m;
}

// sample_random_lights: only the RNG side effects survive (its result feeds an
// `#if 0` block, ray.cpp:1285-1327).  One step picks the entry; a sphere entry
// draws random_spherical_coordinate = two random_between = four more steps.
ORT_HD void sample_random_lights_rng(const PathConsts &c, uint32_t *s)
{
    if(c.light_count == 0u)
    {
        // the reference would divide by zero here (random.h:80); defined instead
        // as "one step, no pick" (include/ort_b200.h)
        xor_shift_32(s);
        return;
    }
    uint32_t idx = random_between_u32(s, 0u, c.light_count);
    if(c.light_is_sphere[idx])
    {
        xor_shift_32(s); xor_shift_32(s); xor_shift_32(s); xor_shift_32(s);
    }
}

// ---- BSDF --------------------------------------------------------------------
// The reference calls powf with the constant exponents 5 and 4 and with base e
// (ray.cpp:829, 857, 964-966).  On the device these are evaluated as repeated
// multiplication and expf: the same functions to within the few ulps by which
// CUDA's libm differs from glibc's anyway (path.h header), at a fraction of the
// instructions and code size of the general powf.  The host build (tests/sim) keeps
// powf so that it stays bit-identical to the oracle.
#if defined(__CUDA_ARCH__)
ORT_HD float pow5_ref(float x) { float x2 = x * x; return x2 * x2 * x; }
ORT_HD float pow4_ref(float x) { float x2 = x * x; return x2 * x2; }
ORT_HD float exp_ref(float y) { return expf(y); }
#else
ORT_HD float pow5_ref(float x) { return powf(x, 5.0f); }
ORT_HD float pow4_ref(float x) { return powf(x, 4.0f); }
ORT_HD float exp_ref(float y) { return powf(ORT_EULER, y); }
#endif

ORT_HD_BIG f3 fresnel(f3 Ks, float l_dot_h)                                  // ray.cpp:825-831
{
    return Ks + (1 - pow5_ref(1.0f - absolute(l_dot_h))) * (mk3(1.0f, 1.0f, 1.0f) - Ks);
}

ORT_HD_BIG float ggx_distribution(f3 N, f3 H, float roughness)               // ray.cpp:834-865
{
    float result = 0.0f;
    float n_dot_h = dot(N, H);
    if(n_dot_h > 0.0f)
    {
        float r2 = square(roughness);
        float tan_theta = sqrtf(1.0f - square(n_dot_h)) / n_dot_h;
        float nom = r2;
        float denom = ORT_PI_32 * pow4_ref(n_dot_h) * square(r2 + square(tan_theta));
        if(!compare_equal_f32(denom, 0.0f)) result = nom / denom;
    }
    return result;
}

ORT_HD_BIG float geometry(f3 w, f3 N, f3 m, float roughness)                 // ray.cpp:868-897
{
    float result = 0.0f;
    float w_dot_n = dot(w, N);
    float w_dot_m = dot(w, m);
    if(!compare_equal_f32(w_dot_m, 0.0f) && (w_dot_n / w_dot_m) > 0)
    {
        if(w_dot_m > 1.0f) result = 1.0f;
        else
        {
            float tan_theta = sqrtf(1.0f - square(w_dot_n)) / w_dot_n;
            if(!compare_equal_f32(tan_theta, 0.0f))
            {
                float r2 = square(roughness);
                result = 2.0f / (1.0f + sqrtf(1 + r2 * square(tan_theta)));
            }
        }
    }
    return result;
}

ORT_HD float get_radicand(f3 m, f3 wo, float n)                          // ray.cpp:899-904
{
    return 1 - square(n) * (1 - square(dot(wo, m)));
}

struct Beern { float ni, no, n; };
ORT_HD Beern get_beer_n(f3 N, f3 wo, float ior)                          // ray.cpp:914-933
{
    Beern r;
    if(dot(N, wo) >= 0.0f) { r.ni = 1.0f; r.no = ior; }
    else { r.ni = ior; r.no = 1.0f; }
    r.n = r.ni / r.no;
    return r;
}

// The half vector H and the refraction half vector m are only formed for materials whose
// specular / transmission lobe can use them (the reference computes them unconditionally and then
// leaves them unused: ray.cpp:943-951, 1027-1030, 1041-1047) -- a diffuse surface, i.e. most
// bounces, skips two normalisations per call.  Every value that reaches the result is computed
// exactly as before.
ORT_HD f3 eval_scattering(f3 N, f3 wi, f3 wo, const Mat &mt, float roughness, float distance)   // ray.cpp:936-1005
{
    const f3 Ks = mt.Ks, Kt = mt.Kt;
    const float ior = mt.ior;
    f3 Ed = mt.Ed;
    f3 Es = mk3(0.0f, 0.0f, 0.0f);
    float wi_dot_n = dot(wi, N);
    float wo_dot_n = dot(wo, N);
    const bool has_Ks = (mt.lobes & 2u) != 0u, has_Kt = (mt.lobes & 4u) != 0u;
    f3 H = mk3(0.0f, 0.0f, 0.0f);
    float wi_dot_h = 0.0f;
    if(has_Ks)
    {
        H = ref_sign(dot(wi, N)) * nrm(wo + wi);
        wi_dot_h = dot(wi, H);
    }
    if(has_Ks && wi_dot_h > 0.0f)
    {
        f3 F = fresnel(Ks, wi_dot_h);
        float D = ggx_distribution(N, H, roughness);
        float G = geometry(wi, N, H, roughness) * geometry(wo, N, H, roughness);
        Es = ((D * G) / (4.0f * absolute(wi_dot_n) * absolute(wo_dot_n))) * F;
    }
    f3 Et = mk3(0.0f, 0.0f, 0.0f);
    if(has_Kt)
    {
        f3 At = mk3(1.0f, 1.0f, 1.0f);
        if(wo_dot_n < 0)
        {
            At.x = exp_ref(distance * logf(Kt.x));
            At.y = exp_ref(distance * logf(Kt.y));
            At.z = exp_ref(distance * logf(Kt.z));
        }
        Beern bn = get_beer_n(N, wo, ior);
        f3 m = nrm(-(bn.ni * wi + bn.no * wo));
        float r = get_radicand(m, wo, bn.n);
        if(r < 0.0f)
        {
            if(has_Ks) Et = hadamard(At, Es);     // total internal reflection
        }
        else
        {
            float wi_dot_m = dot(wi, m);
            float wo_dot_m = dot(wo, m);
            f3 F = mk3(1.0f, 1.0f, 1.0f) - fresnel(Ks, wi_dot_m);
            float D = ggx_distribution(N, m, roughness);
            float G = geometry(wi, N, m, roughness) * geometry(wo, N, m, roughness);
            float denom = (absolute(wi_dot_n) * absolute(wo_dot_n) * square(bn.ni * wi_dot_m + bn.no * wo_dot_m));
            if(!compare_equal_f32(denom, 0.0f))
            {
                f3 nom = (D * G * absolute(wi_dot_m) * absolute(wo_dot_m) * square(bn.no)) * F;
                Et = hadamard(At, nom / denom);
            }
        }
    }
    return absolute(wi_dot_n) * (Ed + Es + Et);
}

ORT_HD float pdf_brdf(f3 N, f3 wi, f3 wo, float roughness, const Mat &mt)   // ray.cpp:1007-1063
{
    const float pd_c = mt.pd_c, ps_c = mt.ps_c, pt_c = mt.pt_c, ior = mt.ior;
    float pd = absolute(dot(wi, N)) / ORT_PI_32;
    float ps = 0.0f;
    if(ps_c > 0.0f)
    {
        f3 H = ref_sign(dot(N, wi)) * nrm(wo + wi);
        float n_dot_h = dot(N, H);
        float wi_dot_h = dot(wi, H);
        float denom = (4.0f * absolute(wi_dot_h));
        if(!compare_equal_f32(denom, 0.0f))
        {
            float D = ggx_distribution(N, H, roughness);
            ps = D * absolute(n_dot_h) / denom;
        }
    }
    float pt = ps;                                                       // sic, ray.cpp:1046
    if(pt_c > 0.0f)
    {
      Beern bn = get_beer_n(N, wo, ior);
      f3 m = nrm(-(bn.ni * wi + bn.no * wo));
      float r = get_radicand(m, wo, bn.n);
      if(r >= 0.0f)
      {
        float n_dot_m = dot(N, m);
        float wi_dot_m = dot(wi, m);
        float wo_dot_m = dot(wo, m);
        float denom = square(bn.no * wo_dot_m + bn.no * wo_dot_m);       // sic, ray.cpp:1054
        if(!compare_equal_f32(denom, 0.0f))
        {
            float D = ggx_distribution(N, m, roughness);
            pt = D * absolute(n_dot_m) * square(bn.no) * absolute(wi_dot_m) / denom;
        }
      }
    }
    return pd_c * pd + ps_c * ps + pt_c * pt;
}

ORT_HD_BIG f3 sample_lobe(f3 N, float c, float phi)                          // ray.cpp:1065-1091
{
    N = nrm(N);
    float s = sqrtf(1.0f - c * c);
    f3 K = mk3(s * cosf(phi), s * sinf(phi), c);
    if(absolute(N.z - 1.0f) < 0.0001f) return K;
    if(absolute(N.z + 1.0f) < 0.0001f) return mk3(K.x, -K.y, -K.z);
    f3 B = nrm(mk3(-N.y, N.x, 0.0f));
    f3 C = cross(N, B);
    return K.x * B + K.y * C + K.z * N;
}

struct SampleBRDF { f3 wi; int is_transmission; };
ORT_HD SampleBRDF sample_brdf(uint32_t *series, f3 N, f3 wo, float roughness, const Mat &mt)   // ray.cpp:1100-1161
{
    SampleBRDF res; res.wi = mk3(0.0f, 0.0f, 0.0f); res.is_transmission = 0;
    const float pd_c = mt.pd_c, ps_c = mt.ps_c, ior = mt.ior;
    float e0 = random_between_0_1(series);
    float e1 = random_between_0_1(series);
    float choice = random_between_0_1(series);
    if(choice < pd_c)
    {
        res.wi = sample_lobe(N, sqrtf(e0), 2.0f * ORT_PI_32 * e1);
    }
    else
    {
        float ggx_cos = cosf(atan2f(roughness * sqrtf(e0), sqrtf(1.0f - e0)));
        f3 m = sample_lobe(N, ggx_cos, 2.0f * ORT_PI_32 * e1);
        bool reflect = true;
        if(!(choice >= pd_c && choice < pd_c + ps_c))
        {
            // transmission lobe
            Beern bn = get_beer_n(N, wo, ior);
            float r = get_radicand(m, wo, bn.n);
            if(!(r < 0.0f))
            {
                res.wi = (bn.n * dot(wo, m) - ref_sign(dot(wo, N)) * sqrtf(r)) * m - bn.n * wo;
                res.is_transmission = 1;
                reflect = false;
            }
        }
        if(reflect) res.wi = 2.0f * absolute(dot(wo, m)) * m - wo;
    }
    res.wi = nrm(res.wi);
    return res;
}

// ---- per-sample path logic -----------------------------------------------------
struct Path
{
    f3 origin, dir;     // the ray to extend next
    f3 wo;              // outgoing direction at the surface the path sits on
    f3 weight;          // throughput
    f3 normal;          // normalised normal of that surface
    uint32_t mat;       // its material index
    uint32_t series;    // RandomSeries.next_random of the sample stream
};

// per-pixel part of "generate": ray.cpp:1215-1221
ORT_HD f3 pixel_focal_point(const PathConsts &c, int x, int y)
{
    float pixel_x = (2.0f * x / (float)c.width) - 1.0f;
    float pixel_y = (2.0f * y / (float)c.height) - 1.0f;
    f3 camera_to_pixel = nrm(pixel_x * c.cam_x + pixel_y * c.cam_y - c.cam_z);
    return c.cam_p + c.focal_length * camera_to_pixel;
}

// per-sample part of "generate": ray.cpp:1232-1246.  Two RNG steps.
ORT_HD void generate_primary(const PathConsts &c, f3 focal_point, Path *p)
{
    float random_rad = random_between(&p->series, 0.0f, 2 * ORT_PI_32);
    f3 lens = c.cam_p + c.aperture_radius * cosf(random_rad) * c.cam_x
                      + c.aperture_radius * sinf(random_rad) * c.cam_y - c.lens_z_offset * c.cam_z;
    p->dir = nrm(focal_point - lens);
    p->wo = -nrm(p->dir);          // normalised twice, as ray.cpp:1237,1240
    p->origin = lens;
    p->weight = mk3(1.0f, 1.0f, 1.0f);
    p->normal = mk3(0.0f, 0.0f, 0.0f);
    p->mat = 0u;
}

// "shade" of the FIRST hit, ray.cpp:1251-1277.  Returns true while the path is alive.
ORT_HD bool shade_primary(const PathConsts &c, Path *p, float hit_t, uint32_t hit_mat, f3 hit_normal, f3 *color)
{
    if(!hit_mat)
    {
        // primary miss: the reference dereferences a null material here
        // (ray.cpp:1251,1329); defined as "path ends" (include/ort_b200.h)
        return false;
    }
    Mat m = load_material(c, hit_mat);
    if(m.is_light)
    {
        *color = *color + m.emit;                                        // ray.cpp:1257
        return false;
    }
    p->origin = p->origin + (hit_t - c.eps) * p->dir;                    // ray.cpp:1262
    p->normal = hit_normal;
    p->mat = hit_mat;
    if(m.lobes & 1u) p->weight = hadamard(p->weight, m.Kd);
    return true;
}

// head of the bounce loop, ray.cpp:1280-1349: roulette, light pick, BSDF sample.
// Returns true when (p->origin, p->dir) is the next ray to extend.
ORT_HD bool next_bounce(const PathConsts &c, Path *p)
{
    if(!(random_between_0_1(&p->series) < c.rr)) return false;
    sample_random_lights_rng(c, &p->series);
    Mat m = load_material(c, p->mat);
    SampleBRDF sb = sample_brdf(&p->series, p->normal, p->wo, c.roughness, m);
    if(sb.is_transmission) p->origin = p->origin + (2.0f * c.eps) * p->dir;   // p->dir is still previous_ray_dir
    p->dir = sb.wi;
    return true;
}

// "shade" of a bounce hit, ray.cpp:1355-1421.  Returns true while the path is alive.
ORT_HD bool shade_bounce(const PathConsts &c, Path *p, float hit_t, uint32_t hit_mat, f3 hit_normal, f3 *color)
{
    if(!hit_mat) return false;                                           // ray.cpp:1420
    Mat m = load_material(c, hit_mat);
    if(m.is_light)
    {
        f3 e = hadamard(p->weight, m.emit);                              // ray.cpp:1361
        if(!is_nan3(e) && !is_inf3(e)) *color = *color + e;
        return false;
    }
    f3 wi = p->dir;
    float pdf = pdf_brdf(hit_normal, wi, p->wo, c.roughness, m) * c.rr;   // ray.cpp:1380
    if(pdf > 0.000001f)
    {
        f3 f = eval_scattering(hit_normal, wi, p->wo, m, c.roughness, hit_t);
        p->weight = hadamard(f / pdf, p->weight);                        // ray.cpp:1403
    }
    p->origin = p->origin + (hit_t - c.eps) * wi;                        // ray.cpp:1411
    p->normal = hit_normal;
    p->mat = hit_mat;
    p->wo = -wi;
    return true;
}

// chunk sum -> 64-bit fixed point (include/ort_b200.h, OrtRenderParams.chunk_spp)
ORT_HD long long to_fixed(float v)
{
    if(v != v) return 0;
    const float sat = (float)(1 << ORT_ACCUM_SAT_BITS);
    if(v > sat) v = sat;
    if(v < -sat) v = -sat;
    float s = v * (float)(1 << ORT_ACCUM_FRAC_BITS);       // exact: power-of-two scaling
#if defined(__CUDA_ARCH__)
    return __float2ll_rn(s);
#else
    return llrintf(s);
#endif
}

} // namespace ort
