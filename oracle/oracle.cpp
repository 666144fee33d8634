// oracle/oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement, in plain scalar C++ + libm, of the hot path of
// gyuhyun-lee/offline_raytracer: the per-pixel radiance loop tiled_raytrace_bvh
// (code/ray.cpp:1178-1466) and everything it calls.  Every function cites the
// reference file:line it follows and keeps the reference's operation ORDER
// (IEEE f32, no FMA contraction: build with -ffp-contract=off), so that on the
// same libm it is bit-identical to the reference -- which is how it is pinned:
// tests/test_oracle_vs_ref.py compares it with oracle/_ref/libref.so (the
// unmodified reference sources compiled for Linux, oracle/Makefile) on whole
// images, ray batches, and function-by-function; tests/golden/ holds vectors
// generated from the reference for boxes where /root/reference is absent.
// PARITY PINNED: yes (against the reference itself run here), see DESIGN.md.
//
// It adds exactly three things the reference lacks:
//   * the winning record's RANK in raycast results (the reference reports no
//     primitive id, code/ray.cpp:613-622).  Rank = position of the record in
//     the order raycast_bvh would test records with no culling: breadth-first
//     over nodes (children 0..7 in order, empty leaves skipped, ray.cpp:780-811)
//     then push-buffer order (ray.cpp:637-774);
//   * a brute-force closest hit over all records (the structure-free definition
//     of the answer);
//   * the chunked per-(pixel, chunk) sample-stream mode of the boundary
//     (include/ort_b200.h, OrtRenderParams.chunk_spp).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product never does.

#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ort_b200.h"

namespace {

typedef ort_v3 V3;
typedef uint32_t u32;
typedef uint64_t u64;
typedef float f32;

#define PI_32 3.14159265358979323846264338327950288419716939937510582097494459230f  // platform.h:45
#define EULER 2.71828182845904523536028747135266249f                                // ray.cpp:4
#define HIT_T_THRESHOLD 0.000001f                                                   // ray.cpp:5

// ---- vector math, code/math.h -------------------------------------------
static inline V3 mk(f32 x, f32 y, f32 z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }      // math.h:212
static inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }      // math.h:224
static inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }                           // math.h:200
static inline V3 mul(f32 s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }           // math.h:267
static inline V3 divs(V3 a, f32 s) { return mk(a.x / s, a.y / s, a.z / s); }          // math.h:235
static inline V3 had(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }      // math.h:326
static inline f32 dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // math.h:320
static inline f32 len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }            // math.h:293
static inline f32 len(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }      // math.h:174
static inline V3 cross(V3 a, V3 b)                                                   // math.h:281
{
    return mk(a.y * b.z - b.y * a.z, b.x * a.z - a.x * b.z, a.x * b.y - b.x * a.y);
}
static inline bool eq_f32(f32 a, f32 b)                                              // math.h:10 compare_equal_f32
{
    f32 diff = a - b;
    return diff >= -0.000001f && diff < 0.000001f;
}
static inline bool is_zero3(V3 v)                                                    // math.h:332 compare_0
{
    const f32 tol = 0.000001f;
    return v.x >= -tol && v.x < tol && v.y >= -tol && v.y < tol && v.z >= -tol && v.z < tol;
}
static inline V3 normalize(V3 a)                                                     // math.h:299
{
    f32 l = len(a);
    if(!eq_f32(l, 0.0f)) return divs(a, l);
    return mk(0, 0, 0);
}
static inline bool nan3(V3 v) { return isnan(v.x) || isnan(v.y) || isnan(v.z); }     // math.h:363
static inline bool inf3(V3 v) { return isinf(v.x) || isinf(v.y) || isinf(v.z); }     // math.h:375
// types.h:50-51: macros, so NaN picks the SECOND operand
#define MAXIMUM(a, b) (((a) > (b)) ? (a) : (b))
#define MINIMUM(a, b) (((a) < (b)) ? (a) : (b))
static inline f32 absolute(f32 v) { return (v <= 0.0f) ? v * -1.0f : v; }            // intrinsic.h:132 (maps +0 to -0)
static inline f32 square(f32 v) { return v * v; }                                     // intrinsic.h:145
static inline f32 signf_(f32 a) { return (a >= 0.0f) ? 1.0f : -1.0f; }               // types.h:52
static inline bool in_rect(V3 p, V3 mn, V3 mx)                                       // math.h:1157 (half-open)
{
    return (p.x >= mn.x && p.x < mx.x) && (p.y >= mn.y && p.y < mx.y) && (p.z >= mn.z && p.z < mx.z);
}

// ---- RNG, code/random.h --------------------------------------------------
static inline void xor_shift_32(u32 *s)                                              // random.h:5-16
{
    u32 x = *s;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x >> 5;   // sic: right shift
    *s = x;
}
static inline f32 rnd01(u32 *s)                                                      // random.h:32-38
{
    xor_shift_32(s);
    return (f32)*s / (f32)UINT32_MAX;
}
static inline f32 rnd_between(u32 *s, f32 mn, f32 mx)                                // random.h:47-53 (steps twice)
{
    xor_shift_32(s);
    return mn + (mx - mn) * rnd01(s);
}
static inline u32 rnd_between_u32(u32 *s, u32 mn, u32 one_past_max)                  // random.h:75-81
{
    xor_shift_32(s);
    return (u32)(*s % (one_past_max - mn) + mn);
}
static inline V3 rnd_spherical(u32 *s, f32 phi_min, f32 phi_max, f32 th_min, f32 th_max) // random.h:99-117
{
    f32 phi = rnd_between(s, phi_min, phi_max);
    f32 theta = rnd_between(s, th_min, th_max);
    f32 sp = sinf(phi), cp = cosf(phi), st = sinf(theta), ct = cosf(theta);
    return mk(cp * ct, cp * st, sp);
}

// ---- intersectors, code/ray.cpp:8-352 ------------------------------------
struct Isect { f32 t; V3 n; int inner; };

static Isect isect_triangle(V3 v0, V3 v1, V3 v2, V3 o, V3 d)                         // ray.cpp:63-115
{
    Isect r; r.t = -1.0f; r.n = mk(0, 0, 0); r.inner = 0;
    V3 e1 = sub(v1, v0);
    V3 e2 = sub(v2, v0);
    V3 p = cross(d, e2);
    f32 det = dot(p, e1);
    V3 T = sub(o, v0);
    const f32 tol = 0.000001f;
    if(det <= -tol || det >= tol)
    {
        V3 a = cross(T, e1);
        f32 t = dot(a, e2) / det;
        f32 u = dot(p, T) / det;
        f32 v = dot(a, d) / det;
        if(t >= HIT_T_THRESHOLD && u >= 0.0f && v >= 0.0f && u + v <= 1.0f)
        {
            r.t = t;
            r.n = cross(e1, e2);
        }
    }
    return r;
}

static Isect isect_sphere(V3 c, f32 rad, V3 o, V3 d)                                 // ray.cpp:132-190
{
    Isect r; r.t = -1.0f; r.n = mk(0, 0, 0); r.inner = 0;
    V3 rel = sub(o, c);
    f32 a = dot(d, d);
    f32 b = dot(d, rel);
    f32 cc = dot(rel, rel) - rad * rad;
    f32 root = b * b - a * cc;
    const f32 tol = 0.00001f;
    if(root >= tol)
    {
        f32 sq = sqrtf(root);
        f32 tn = (-b - sq) / a;
        f32 tp = (-b + sq) / a;
        f32 t;
        if(tn < 0.0f) { t = tp; r.inner = 1; }
        else t = tn;
        if(t > HIT_T_THRESHOLD)
        {
            r.t = t;
            r.n = mul(1.0f, sub(add(o, mul(r.t, d)), c));
        }
    }
    else if(root < tol && root > -tol)
    {
        f32 t = (-b) / (2 * a);
        if(t > HIT_T_THRESHOLD)
        {
            r.t = t;
            r.n = sub(add(o, mul(r.t, d)), c);
        }
    }
    return r;
}

static Isect isect_aab(V3 mn, V3 mx, V3 o, V3 d)                                     // ray.cpp:206-283
{
    Isect r; r.t = -1.0f; r.n = mk(0, 0, 0); r.inner = 0;
    V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    V3 t0 = had(sub(mn, o), inv);
    V3 t1 = had(sub(mx, o), inv);
    V3 tmin = mk(MINIMUM(t0.x, t1.x), MINIMUM(t0.y, t1.y), MINIMUM(t0.z, t1.z));
    V3 tmax = mk(MAXIMUM(t0.x, t1.x), MAXIMUM(t0.y, t1.y), MAXIMUM(t0.z, t1.z));
    f32 max_of_min = MAXIMUM(MAXIMUM(tmin.x, tmin.y), tmin.z);
    f32 min_of_max = MINIMUM(MINIMUM(tmax.x, tmax.y), tmax.z);
    if(min_of_max >= max_of_min)
    {
        f32 tx = t0.x; V3 nx = mk(-1, 0, 0);
        if(tx > t1.x) { tx = t1.x; nx = mk(1, 0, 0); }
        f32 ty = t0.y; V3 ny = mk(0, -1, 0);
        if(ty > t1.y) { ty = t1.y; ny = mk(0, 1, 0); }
        f32 tz = t0.z; V3 nz = mk(0, 0, -1);
        if(tz > t1.z) { tz = t1.z; nz = mk(0, 0, 1); }
        f32 best = tx; V3 bn = nx;
        if(best < ty) { best = ty; bn = ny; }
        if(best < tz) { best = tz; bn = nz; }
        r.t = max_of_min;
        r.n = bn;
    }
    return r;
}

struct M3 { V3 r0, r1, r2; };
static inline V3 m3_mul(M3 m, V3 v) { return mk(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); } // math.h:957
static inline M3 m3_transpose(M3 m)                                                   // math.h:981
{
    M3 t;
    t.r0 = mk(m.r0.x, m.r1.x, m.r2.x);
    t.r1 = mk(m.r0.y, m.r1.y, m.r2.y);
    t.r2 = mk(m.r0.z, m.r1.z, m.r2.z);
    return t;
}
static M3 rotation_matrix_along_z(V3 src)                                            // ray.cpp:8-33
{
    M3 m; m.r0 = mk(1, 0, 0); m.r1 = mk(0, 1, 0); m.r2 = mk(0, 0, 1);
    if(!is_zero3(cross(src, mk(0, 0, 1))))
    {
        V3 a = normalize(src);
        V3 b = normalize(cross(mk(0, 0, 1), a));
        if(is_zero3(b)) b = normalize(cross(mk(1, 0, 0), a));
        V3 c = cross(a, b);
        m.r0 = b; m.r1 = c; m.r2 = a;
    }
    return m;
}

static Isect isect_cylinder(V3 base, V3 axis, f32 radius, V3 o, V3 d)                // ray.cpp:286-352
{
    Isect r; r.t = -1.0f; r.n = mk(0, 0, 0); r.inner = 0;
    M3 rot = rotation_matrix_along_z(axis);
    o = m3_mul(rot, sub(o, base));
    d = m3_mul(rot, d);
    f32 t_bottom = (-o.z) / d.z;
    f32 t_top = (len(axis) - o.z) / d.z;
    f32 t_slab_min = MINIMUM(t_bottom, t_top);
    f32 t_slab_max = MAXIMUM(t_bottom, t_top);
    f32 a = d.x * d.x + d.y * d.y;
    f32 b = d.x * o.x + d.y * o.y;
    f32 c = (o.x * o.x + o.y * o.y) - radius * radius;
    f32 det = b * b - a * c;
    if(det >= 0.0f)
    {
        f32 sq = sqrtf(det);
        f32 tc_min = (-b - sq) / a;
        f32 tc_max = (-b + sq) / a;
        f32 t_max_in_min = MAXIMUM(t_slab_min, tc_min);
        f32 t_min_in_max = MINIMUM(t_slab_max, tc_max);
        if(t_max_in_min <= t_min_in_max)
        {
            r.t = t_max_in_min;
            V3 n = mk(0, 1, 0);   // sic: cap normal is +Y of the rotated frame (ray.cpp:330)
            if(t_slab_min < tc_min)
            {
                V3 hp = add(o, mul(r.t, d));
                n = mk(hp.x, hp.y, 0);
            }
            r.n = m3_mul(m3_transpose(rot), n);
        }
    }
    return r;
}

// ---- octree access -------------------------------------------------------
struct Record
{
    u32 type;      // OrtShapeType
    u32 mat;
    V3 a, b, c;    // triangle: v0,v1,v2; sphere: center,(r,0,0); aab: min,max; cylinder: base,axis,(r,0,0)
};

static inline u32 payload_size(u32 type)
{
    switch(type)
    {
        case ORT_SHAPE_SPHERE:   return (u32)sizeof(OrtSphere);
        case ORT_SHAPE_AAB:      return (u32)sizeof(OrtAAB);
        case ORT_SHAPE_CYLINDER: return (u32)sizeof(OrtCylinder);
        case ORT_SHAPE_TRIANGLE: return (u32)sizeof(OrtTriangle);
        case ORT_SHAPE_CSG:      return (u32)sizeof(OrtCSG);
        default: return 0;
    }
}

static inline bool decode_record(const uint8_t *p, Record *r)
{
    u32 type; memcpy(&type, p, 4);
    r->type = type;
    const uint8_t *q = p + 4;
    switch(type)
    {
        case ORT_SHAPE_SPHERE:
        {
            OrtSphere s; memcpy(&s, q, sizeof(s));
            r->a = s.center; r->b = mk(s.r, 0, 0); r->c = mk(0, 0, 0); r->mat = s.mat_index;
        } return true;
        case ORT_SHAPE_AAB:
        {
            OrtAAB s; memcpy(&s, q, sizeof(s));
            r->a = s.min; r->b = s.max; r->c = mk(0, 0, 0); r->mat = s.mat_index;
        } return true;
        case ORT_SHAPE_CYLINDER:
        {
            OrtCylinder s; memcpy(&s, q, sizeof(s));
            r->a = s.base; r->b = s.axis; r->c = mk(s.r, 0, 0); r->mat = s.mat_index;
        } return true;
        case ORT_SHAPE_TRIANGLE:
        {
            OrtTriangle s; memcpy(&s, q, sizeof(s));
            r->a = s.mesh->vertices[s.i_0];                                     // ray.cpp:702-704
            r->b = s.mesh->vertices[s.i_1];
            r->c = s.mesh->vertices[s.i_2];
            r->mat = s.mesh->mat_index;
        } return true;
        case ORT_SHAPE_CSG:
            r->mat = 0; return true;                                            // inert, ray.cpp:718-767
        default:
            return false;
    }
}

static inline Isect isect_record(const Record &rec, V3 o, V3 d)
{
    switch(rec.type)
    {
        case ORT_SHAPE_SPHERE:   return isect_sphere(rec.a, rec.b.x, o, d);
        case ORT_SHAPE_AAB:      return isect_aab(rec.a, rec.b, o, d);
        case ORT_SHAPE_CYLINDER: return isect_cylinder(rec.a, rec.b, rec.c.x, o, d);
        case ORT_SHAPE_TRIANGLE: return isect_triangle(rec.a, rec.b, rec.c, o, d);
        default: { Isect r; r.t = -1.0f; r.n = mk(0, 0, 0); r.inner = 0; return r; }
    }
}

struct OracleScene
{
    const OrtWorld *world;
    const OrtBVHOctreeNode *root;
    std::unordered_map<const OrtBVHOctreeNode *, u32> first_rank;  // node -> rank of its first record
    std::vector<Record> records;                                     // all records, indexed by rank
    std::vector<uint8_t> light_is_sphere;
    u32 node_count;
    u32 max_depth;
};

struct RayHit
{
    f32 t;          // FLT_MAX on a miss (ray.cpp:627)
    V3 n;           // normalised (ray.cpp:817)
    u32 mat;        // 0 on a miss
    u32 rank;       // ORT_MISS_RANK on a miss
    int inner;
    u64 shape_tests, box_tests, node_visits;
};

static inline bool child_is_live(const OrtBVHOctreeNode *c)                          // ray.cpp:788
{
    return (c->is_leaf && c->push_buffer.used) || c->first_child;
}

// raycast_top_most_node + raycast_bvh, code/ray.cpp:624-822, 1165-1176
static void raycast_bfs(const OracleScene &sc, std::vector<const OrtBVHOctreeNode *> &queue, V3 o, V3 d, RayHit *res)
{
    res->t = FLT_MAX; res->n = mk(0, 0, 0); res->mat = 0; res->rank = ORT_MISS_RANK; res->inner = 0;
    res->shape_tests = res->box_tests = res->node_visits = 0;
    queue.clear();
    queue.push_back(sc.root);                                                        // ray.cpp:1170
    for(size_t cursor = 0; cursor < queue.size(); ++cursor)
    {
        const OrtBVHOctreeNode *node = queue[cursor];
        res->node_visits++;
        const uint8_t *base = (const uint8_t *)node->push_buffer.base;
        u32 rank = 0;
        if(node->push_buffer.used) rank = sc.first_rank.find(node)->second;
        for(size_t consumed = 0; consumed < node->push_buffer.used; )
        {
            Record rec;
            if(!decode_record(base + consumed, &rec)) break;
            consumed += 4 + payload_size(rec.type);
            if(rec.type != ORT_SHAPE_CSG)
            {
                Isect is = isect_record(rec, o, d);
                if(is.t >= HIT_T_THRESHOLD && is.t < res->t)                         // ray.cpp:653,670,686,708
                {
                    res->t = is.t; res->n = is.n; res->mat = rec.mat; res->inner = is.inner; res->rank = rank;
                }
                res->shape_tests++;
            }
            rank++;
        }
        if(node->first_child)
        {
            for(u32 ci = 0; ci < 8; ++ci)
            {
                const OrtBVHOctreeNode *child = node->first_child + ci;
                if(child_is_live(child))
                {
                    bool should_add = false;
                    if(in_rect(o, child->aabb_min, child->aabb_max)) should_add = true;   // ray.cpp:792
                    else
                    {
                        Isect is = isect_aab(child->aabb_min, child->aabb_max, o, d);     // ray.cpp:799
                        res->box_tests++;
                        if(is.t >= HIT_T_THRESHOLD && is.t < res->t) should_add = true;
                    }
                    if(should_add) queue.push_back(child);
                }
            }
        }
    }
    res->n = normalize(res->n);                                                      // ray.cpp:817
}

// structure-free closest hit: argmin over all records of (t, rank)
static void raycast_brute(const OracleScene &sc, V3 o, V3 d, RayHit *res)
{
    res->t = FLT_MAX; res->n = mk(0, 0, 0); res->mat = 0; res->rank = ORT_MISS_RANK; res->inner = 0;
    res->shape_tests = res->box_tests = res->node_visits = 0;
    for(size_t i = 0; i < sc.records.size(); ++i)
    {
        const Record &rec = sc.records[i];
        if(rec.type == ORT_SHAPE_CSG) continue;
        Isect is = isect_record(rec, o, d);
        res->shape_tests++;
        if(is.t >= HIT_T_THRESHOLD && is.t < res->t)
        {
            res->t = is.t; res->n = is.n; res->mat = rec.mat; res->inner = is.inner; res->rank = (u32)i;
        }
    }
    res->n = normalize(res->n);
}

// ---- BSDF, code/ray.cpp:825-1161 ------------------------------------------
static V3 fresnel(V3 Ks, f32 l_dot_h)                                                // ray.cpp:825-831
{
    return add(Ks, mul(1 - powf(1.0f - absolute(l_dot_h), 5.0f), sub(mk(1, 1, 1), Ks)));
}

static f32 ggx_distribution(V3 N, V3 H, f32 roughness)                               // ray.cpp:834-865
{
    f32 result = 0.0f;
    f32 n_dot_h = dot(N, H);
    if(n_dot_h > 0.0f)
    {
        f32 r2 = square(roughness);
        f32 tan_theta = sqrtf(1.0f - square(n_dot_h)) / n_dot_h;
        f32 nom = r2;
        f32 denom = PI_32 * powf(n_dot_h, 4.0f) * square(r2 + square(tan_theta));
        if(!eq_f32(denom, 0.0f)) result = nom / denom;
    }
    return result;
}

static f32 geometry(V3 w, V3 N, V3 m, f32 roughness)                                 // ray.cpp:868-897
{
    f32 result = 0.0f;
    f32 w_dot_n = dot(w, N);
    f32 w_dot_m = dot(w, m);
    if(!eq_f32(w_dot_m, 0.0f) && (w_dot_n / w_dot_m) > 0)
    {
        if(w_dot_m > 1.0f) result = 1.0f;
        else
        {
            f32 tan_theta = sqrtf(1.0f - square(w_dot_n)) / w_dot_n;
            if(!eq_f32(tan_theta, 0.0f))
            {
                f32 r2 = square(roughness);
                result = 2.0f / (1.0f + sqrtf(1 + r2 * square(tan_theta)));
            }
        }
    }
    return result;
}

static inline f32 get_radicand(V3 m, V3 wo, f32 n)                                   // ray.cpp:899-904
{
    return 1 - square(n) * (1 - square(dot(wo, m)));
}

struct Beern { f32 ni, no, n; };
static inline Beern get_beer_n(V3 N, V3 wo, f32 ior)                                 // ray.cpp:914-933
{
    Beern r;
    if(dot(N, wo) >= 0.0f) { r.ni = 1.0f; r.no = ior; }
    else { r.ni = ior; r.no = 1.0f; }
    r.n = r.ni / r.no;
    return r;
}

static V3 eval_scattering(V3 N, V3 wi, V3 wo, V3 Kd, V3 Ks, V3 Kt, f32 ior, f32 roughness, f32 distance) // ray.cpp:936-1005
{
    V3 Ed = divs(Kd, PI_32);
    V3 H = mul(signf_(dot(wi, N)), normalize(add(wo, wi)));
    f32 wi_dot_h = dot(wi, H);
    V3 Es = mk(0, 0, 0);
    f32 wi_dot_n = dot(wi, N);
    f32 wo_dot_n = dot(wo, N);
    if(wi_dot_h > 0.0f && len2(Ks) > 0.0f)
    {
        V3 F = fresnel(Ks, wi_dot_h);
        f32 D = ggx_distribution(N, H, roughness);
        f32 G = geometry(wi, N, H, roughness) * geometry(wo, N, H, roughness);
        Es = mul((D * G) / (4.0f * absolute(wi_dot_n) * absolute(wo_dot_n)), F);
    }
    V3 Et = mk(0, 0, 0);
    if(len2(Kt) > 0.0f)
    {
        V3 At = mk(1, 1, 1);
        if(wo_dot_n < 0)
        {
            At.x = powf(EULER, distance * logf(Kt.x));
            At.y = powf(EULER, distance * logf(Kt.y));
            At.z = powf(EULER, distance * logf(Kt.z));
        }
        Beern bn = get_beer_n(N, wo, ior);
        V3 m = normalize(neg(add(mul(bn.ni, wi), mul(bn.no, wo))));
        f32 r = get_radicand(m, wo, bn.n);
        if(r < 0.0f)
        {
            if(len2(Ks) > 0.0f) Et = had(At, Es);
        }
        else
        {
            f32 wi_dot_m = dot(wi, m);
            f32 wo_dot_m = dot(wo, m);
            V3 F = sub(mk(1, 1, 1), fresnel(Ks, wi_dot_m));
            f32 D = ggx_distribution(N, m, roughness);
            f32 G = geometry(wi, N, m, roughness) * geometry(wo, N, m, roughness);
            f32 denom = (absolute(wi_dot_n) * absolute(wo_dot_n) * square(bn.ni * wi_dot_m + bn.no * wo_dot_m));
            if(!eq_f32(denom, 0.0f))
            {
                V3 nom = mul(D * G * absolute(wi_dot_m) * absolute(wo_dot_m) * square(bn.no), F);
                Et = had(At, divs(nom, denom));
            }
        }
    }
    return mul(absolute(wi_dot_n), add(add(Ed, Es), Et));
}

static f32 pdf_brdf(V3 N, V3 wi, V3 wo, f32 roughness, V3 Kd, V3 Ks, V3 Kt, f32 ior)    // ray.cpp:1007-1063
{
    f32 Kd_l = len(Kd), Ks_l = len(Ks), Kt_l = len(Kt);
    f32 s = Kd_l + Ks_l + Kt_l;
    f32 pd_c = Kd_l / s, ps_c = Ks_l / s, pt_c = Kt_l / s;
    f32 pd = absolute(dot(wi, N)) / PI_32;
    V3 H = mul(signf_(dot(N, wi)), normalize(add(wo, wi)));
    f32 n_dot_h = dot(N, H);
    f32 wi_dot_h = dot(wi, H);
    f32 ps = 0.0f;
    if(ps_c > 0.0f)
    {
        f32 denom = (4.0f * absolute(wi_dot_h));
        if(!eq_f32(denom, 0.0f))
        {
            f32 D = ggx_distribution(N, H, roughness);
            ps = D * absolute(n_dot_h) / denom;
        }
    }
    Beern bn = get_beer_n(N, wo, ior);
    V3 m = normalize(neg(add(mul(bn.ni, wi), mul(bn.no, wo))));
    f32 r = get_radicand(m, wo, bn.n);
    f32 pt = ps;                                                                     // sic, ray.cpp:1046
    if(pt_c > 0.0f && r >= 0.0f)
    {
        f32 n_dot_m = dot(N, m);
        f32 wi_dot_m = dot(wi, m);
        f32 wo_dot_m = dot(wo, m);
        f32 denom = square(bn.no * wo_dot_m + bn.no * wo_dot_m);                     // sic, ray.cpp:1054
        if(!eq_f32(denom, 0.0f))
        {
            f32 D = ggx_distribution(N, m, roughness);
            pt = D * absolute(n_dot_m) * square(bn.no) * absolute(wi_dot_m) / denom;
        }
    }
    return pd_c * pd + ps_c * ps + pt_c * pt;
}

static V3 sample_lobe(V3 N, f32 c, f32 phi)                                          // ray.cpp:1065-1091
{
    N = normalize(N);
    f32 s = sqrtf(1.0f - c * c);
    V3 K = mk(s * cosf(phi), s * sinf(phi), c);
    if(absolute(N.z - 1.0f) < 0.0001f) return K;
    if(absolute(N.z + 1.0f) < 0.0001f) return mk(K.x, -K.y, -K.z);
    V3 B = normalize(mk(-N.y, N.x, 0));
    V3 C = cross(N, B);
    return add(add(mul(K.x, B), mul(K.y, C)), mul(K.z, N));
}

struct SampleBRDF { V3 wi; int is_transmission; };
static SampleBRDF sample_brdf(u32 *series, V3 N, V3 wo, f32 roughness, V3 Kd, V3 Ks, V3 Kt, f32 ior) // ray.cpp:1100-1161
{
    SampleBRDF res; res.wi = mk(0, 0, 0); res.is_transmission = 0;
    f32 Kd_l = len(Kd), Ks_l = len(Ks), Kt_l = len(Kt);
    f32 s = Kd_l + Ks_l + Kt_l;
    f32 pd_c = Kd_l / s, ps_c = Ks_l / s;
    f32 e0 = rnd01(series);
    f32 e1 = rnd01(series);
    f32 choice = rnd01(series);
    if(choice < pd_c)
    {
        res.wi = sample_lobe(N, sqrtf(e0), 2.0f * PI_32 * e1);
    }
    else if(choice >= pd_c && choice < pd_c + ps_c)
    {
        f32 ggx_cos = cosf(atan2f(roughness * sqrtf(e0), sqrtf(1.0f - e0)));
        V3 m = sample_lobe(N, ggx_cos, 2.0f * PI_32 * e1);
        res.wi = sub(mul(2.0f * absolute(dot(wo, m)), m), wo);
    }
    else
    {
        f32 ggx_cos = cosf(atan2f(roughness * sqrtf(e0), sqrtf(1.0f - e0)));
        V3 m = sample_lobe(N, ggx_cos, 2.0f * PI_32 * e1);
        Beern bn = get_beer_n(N, wo, ior);
        f32 r = get_radicand(m, wo, bn.n);
        if(r < 0.0f)
        {
            res.wi = sub(mul(2.0f * absolute(dot(wo, m)), m), wo);
        }
        else
        {
            res.wi = sub(mul(bn.n * dot(wo, m) - signf_(dot(wo, N)) * sqrtf(r), m), mul(bn.n, wo));
            res.is_transmission = 1;
        }
    }
    res.wi = normalize(res.wi);
    return res;
}

// sample_random_lights, code/ray.cpp:537-601: only its RNG side effects survive
// (the NEE block that would use the result is #if 0, ray.cpp:1285-1327).
static void sample_random_lights_rng(const OracleScene &sc, u32 *series)
{
    u32 light_count = (u32)sc.light_is_sphere.size();
    if(light_count == 0)
    {
        // the reference divides by zero here (random.h:80); the boundary defines
        // "one RNG step, no pick" instead.  No shipped scene has 0 lights.
        xor_shift_32(series);
        return;
    }
    u32 idx = rnd_between_u32(series, 0, light_count);
    if(sc.light_is_sphere[idx])
        (void)rnd_spherical(series, -PI_32 / 2.0f, PI_32 / 2.0f, 0, 2.0f * PI_32);
}

// ---- integrator, code/ray.cpp:1178-1466 -----------------------------------
struct Counters { u64 rays, shape_tests, box_tests, node_visits; };

// One sample stream: n samples of pixel (x, y) drawn sequentially from *series.
// Returns the float radiance SUM in sample order (the reference's `color`).
static V3 trace_stream(const OracleScene &sc, const OrtCamera *cam, const OrtRenderParams *P,
                       int x, int y, u32 *series, u32 n_samples,
                       std::vector<const OrtBVHOctreeNode *> &queue, Counters *cnt, int brute)
{
    const OrtMaterial *mats = sc.world->materials;
    f32 roughness = P->roughness;
    f32 eps = P->dont_get_too_close_epsilon;
    f32 rr = P->russian_roulette_value;
    f32 focal_length = len(sub(cam->p, mk(P->focus_target[0], P->focus_target[1], P->focus_target[2]))); // ray.cpp:1198
    f32 aperture_radius = P->aperture_radius;

    V3 color = mk(0, 0, 0);
    f32 pixel_x = (2.0f * x / (f32)P->output_width) - 1.0f;                         // ray.cpp:1215-1216
    f32 pixel_y = (2.0f * y / (f32)P->output_height) - 1.0f;
    V3 camera_to_pixel = normalize(sub(add(mul(pixel_x, cam->x_axis), mul(pixel_y, cam->y_axis)), cam->z_axis));
    V3 focal_point = add(cam->p, mul(focal_length, camera_to_pixel));

    for(u32 ray_index = 0; ray_index < n_samples; ++ray_index)
    {
        f32 random_rad = rnd_between(series, 0.0f, 2 * PI_32);                       // ray.cpp:1232
        V3 lens = sub(add(add(cam->p, mul(aperture_radius * cosf(random_rad), cam->x_axis)),
                          mul(aperture_radius * sinf(random_rad), cam->y_axis)),
                      mul(P->lens_z_offset, cam->z_axis));                            // ray.cpp:1233-1234
        V3 initial_dir = normalize(sub(focal_point, lens));
        V3 wo = neg(normalize(initial_dir));                                         // ray.cpp:1240 (normalised twice)
        V3 origin = lens;
        bool alive = true;
        V3 prev_dir = mk(0, 0, 0);
        V3 hit_normal = mk(0, 0, 0);
        V3 weight = mk(1, 1, 1);
        const OrtMaterial *hit_mat = 0;

        RayHit h;
        if(brute) raycast_brute(sc, origin, initial_dir, &h); else raycast_bfs(sc, queue, origin, initial_dir, &h);
        cnt->rays++; cnt->shape_tests += h.shape_tests; cnt->box_tests += h.box_tests; cnt->node_visits += h.node_visits;
        if(h.mat)
        {
            const OrtMaterial *mat = mats + h.mat;
            if(mat->is_light)
            {
                color = add(color, mat->emit_color);                                  // ray.cpp:1257
                alive = false;
            }
            else
            {
                origin = add(origin, mul(h.t - eps, initial_dir));                    // ray.cpp:1262
                hit_normal = h.n;
                hit_mat = mat;
                prev_dir = initial_dir;
                if(len2(hit_mat->diffuse) > 0.0f) weight = had(weight, hit_mat->diffuse);
            }
        }
        else
        {
            // The reference leaves ray_alive set with hit_mat == 0 and then
            // dereferences it (ray.cpp:1251,1329).  The boundary defines a primary
            // miss as "path ends, contributes nothing".  Closed scenes never get here.
            alive = false;
        }

        while(alive && rnd01(series) < rr)                                           // ray.cpp:1280
        {
            sample_random_lights_rng(sc, series);                                    // ray.cpp:1283
            V3 Ks = mk(hit_mat->specular.x, hit_mat->specular.y, hit_mat->specular.z);
            SampleBRDF sb = sample_brdf(series, hit_normal, wo, roughness, hit_mat->diffuse, Ks,
                                        hit_mat->transmission, hit_mat->ior);        // ray.cpp:1335
            V3 wi = sb.wi;
            if(sb.is_transmission) origin = add(origin, mul(2.0f * eps, prev_dir));  // ray.cpp:1347

            if(brute) raycast_brute(sc, origin, wi, &h); else raycast_bfs(sc, queue, origin, wi, &h);
            cnt->rays++; cnt->shape_tests += h.shape_tests; cnt->box_tests += h.box_tests; cnt->node_visits += h.node_visits;
            if(h.mat)
            {
                const OrtMaterial *m2 = mats + h.mat;
                if(m2->is_light)
                {
                    V3 c = had(weight, m2->emit_color);                               // ray.cpp:1361
                    if(!nan3(c) && !inf3(c)) color = add(color, c);
                    alive = false;
                }
                else
                {
                    V3 Ks2 = mk(m2->specular.x, m2->specular.y, m2->specular.z);
                    f32 p = pdf_brdf(h.n, wi, wo, roughness, m2->diffuse, Ks2, m2->transmission, m2->ior) * rr; // ray.cpp:1380
                    if(p > 0.000001f)
                    {
                        V3 f = eval_scattering(h.n, wi, wo, m2->diffuse, Ks2, m2->transmission, m2->ior, roughness, h.t);
                        weight = had(divs(f, p), weight);                             // ray.cpp:1403
                    }
                    origin = add(origin, mul(h.t - eps, wi));                         // ray.cpp:1411
                    hit_normal = h.n;
                    hit_mat = m2;
                    prev_dir = wi;
                    wo = neg(wi);
                }
            }
            else alive = false;                                                      // ray.cpp:1420
        }
    }
    return color;
}

static inline int64_t to_fixed(f32 v)
{
    // round-to-nearest-even fixed point with ORT_ACCUM_FRAC_BITS fractional bits,
    // saturating at +-2^ORT_ACCUM_SAT_BITS; NaN -> 0  (include/ort_b200.h)
    double s = (double)v * (double)(1 << ORT_ACCUM_FRAC_BITS);
    const double sat = (double)((int64_t)1 << (ORT_ACCUM_SAT_BITS + ORT_ACCUM_FRAC_BITS));
    if(!(s == s)) return 0;
    if(s > sat) s = sat;
    if(s < -sat) s = -sat;
    return (int64_t)llrint(s);
}

} // namespace

extern "C" {

void *oracle_scene_create(const OrtWorld *world, const OrtBVHOctreeNode *root)
{
    OracleScene *sc = new OracleScene();
    sc->world = world;
    sc->root = root;
    sc->node_count = 0;
    sc->max_depth = 0;
    // full breadth-first walk == the order raycast_bvh visits nodes when nothing is culled
    std::vector<std::pair<const OrtBVHOctreeNode *, u32> > q;
    q.push_back(std::make_pair(root, 0u));
    for(size_t cursor = 0; cursor < q.size(); ++cursor)
    {
        const OrtBVHOctreeNode *node = q[cursor].first;
        u32 depth = q[cursor].second;
        sc->node_count++;
        if(depth > sc->max_depth) sc->max_depth = depth;
        if(node->push_buffer.used)
        {
            sc->first_rank[node] = (u32)sc->records.size();
            const uint8_t *base = (const uint8_t *)node->push_buffer.base;
            for(size_t consumed = 0; consumed < node->push_buffer.used; )
            {
                Record rec;
                if(!decode_record(base + consumed, &rec)) break;
                consumed += 4 + payload_size(rec.type);
                sc->records.push_back(rec);
            }
        }
        if(node->first_child)
            for(u32 ci = 0; ci < 8; ++ci)
            {
                const OrtBVHOctreeNode *child = node->first_child + ci;
                if(child_is_live(child)) q.push_back(std::make_pair(child, depth + 1));
            }
    }
    // light list: packed (u32 type, void *ptr) pairs, parser.cpp:1144-1182
    const uint8_t *lb = (const uint8_t *)world->light_push_buffer.base;
    for(size_t consumed = 0; consumed + 12 <= world->light_push_buffer.used; consumed += 12)
    {
        u32 type; memcpy(&type, lb + consumed, 4);
        sc->light_is_sphere.push_back(type == ORT_SHAPE_SPHERE ? 1 : 0);
    }
    return sc;
}

void oracle_scene_destroy(void *h) { delete (OracleScene *)h; }

// info[0..3] = record count, visited node count, max depth, light count
void oracle_scene_info(void *h, uint32_t *info)
{
    OracleScene *sc = (OracleScene *)h;
    info[0] = (u32)sc->records.size();
    info[1] = sc->node_count;
    info[2] = sc->max_depth;
    info[3] = (u32)sc->light_is_sphere.size();
}

// record dump for flattening tests: per rank {type, mat} and 9 floats
void oracle_scene_records(void *h, uint32_t *type_mat, float *geom)
{
    OracleScene *sc = (OracleScene *)h;
    for(size_t i = 0; i < sc->records.size(); ++i)
    {
        const Record &r = sc->records[i];
        type_mat[2*i] = r.type; type_mat[2*i+1] = r.mat;
        const V3 *v[3] = { &r.a, &r.b, &r.c };
        for(int k = 0; k < 3; ++k)
        {
            geom[9*i + 3*k + 0] = (r.type == ORT_SHAPE_CSG) ? 0.0f : v[k]->x;
            geom[9*i + 3*k + 1] = (r.type == ORT_SHAPE_CSG) ? 0.0f : v[k]->y;
            geom[9*i + 3*k + 2] = (r.type == ORT_SHAPE_CSG) ? 0.0f : v[k]->z;
        }
    }
}

// mode 0: the reference's BFS octree traversal; mode 1: brute force over all records.
// counters[0..2] += shape tests, box tests, node visits.
void oracle_raycast_batch(void *h, uint64_t n, const float *origins, const float *dirs, int mode,
                          float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, float *hit_normal,
                          int32_t *inner_hit, uint64_t *counters, int n_threads)
{
    OracleScene *sc = (OracleScene *)h;
    if(n_threads < 1) n_threads = 1;
    std::atomic<unsigned long long> c0(0), c1(0), c2(0);
    auto work = [&](int tid)
    {
        std::vector<const OrtBVHOctreeNode *> queue;
        queue.reserve(4096);
        u64 a = 0, b = 0, c = 0;
        for(u64 i = (u64)tid; i < n; i += (u64)n_threads)
        {
            V3 o = mk(origins[3*i], origins[3*i+1], origins[3*i+2]);
            V3 d = mk(dirs[3*i], dirs[3*i+1], dirs[3*i+2]);
            RayHit r;
            if(mode == 1) raycast_brute(*sc, o, d, &r); else raycast_bfs(*sc, queue, o, d, &r);
            if(hit_t) hit_t[i] = r.t;
            if(prim_rank) prim_rank[i] = r.rank;
            if(mat_index) mat_index[i] = r.mat;
            if(hit_normal) { hit_normal[3*i] = r.n.x; hit_normal[3*i+1] = r.n.y; hit_normal[3*i+2] = r.n.z; }
            if(inner_hit) inner_hit[i] = r.inner;
            a += r.shape_tests; b += r.box_tests; c += r.node_visits;
        }
        c0 += a; c1 += b; c2 += c;
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work, i);
    work(0);
    for(auto &t : threads) t.join();
    if(counters) { counters[0] += c0.load(); counters[1] += c1.load(); counters[2] += c2.load(); }
}

// The boundary's render semantics (include/ort_b200.h: ort_render) on the CPU.
// counters[0..3] += rays, shape tests, box tests, node visits.
void oracle_render(void *h, const OrtCamera *cam, const OrtRenderParams *P, ort_v3 *out,
                   uint64_t *counters, int n_threads, int brute)
{
    OracleScene *sc = (OracleScene *)h;
    if(n_threads < 1) n_threads = 1;
    u32 spp = P->ray_per_pixel_count;
    u32 chunk_spp = P->chunk_spp ? P->chunk_spp : spp;
    if(chunk_spp > spp) chunk_spp = spp;
    u32 n_chunks = spp ? (spp + chunk_spp - 1) / chunk_spp : 0;
    u32 c_begin = P->chunk_begin, c_end = P->chunk_end;
    if(c_begin == 0 && c_end == 0) c_end = n_chunks;
    if(c_end > n_chunks) c_end = n_chunks;
    std::atomic<int> next_row(P->tile_min_y);
    std::atomic<unsigned long long> c0(0), c1(0), c2(0), c3(0);
    auto work = [&]()
    {
        std::vector<const OrtBVHOctreeNode *> queue;
        queue.reserve(4096);
        Counters cnt = { 0, 0, 0, 0 };
        for(;;)
        {
            int y = next_row.fetch_add(1);
            if(y >= P->tile_one_past_max_y) break;
            for(int x = P->tile_min_x; x < P->tile_one_past_max_x; ++x)
            {
                u32 pixel_index = (u32)(y * P->output_width + x);
                ort_v3 *pixel = out + (size_t)y * P->output_width + x;
                if(n_chunks == 1)
                {
                    u32 series = ort_stream_seed(P->base_seed, pixel_index, 0);
                    V3 color = trace_stream(*sc, cam, P, x, y, &series, spp, queue, &cnt, brute);
                    *pixel = divs(color, (f32)spp);                                   // ray.cpp:1428
                }
                else
                {
                    int64_t acc[3] = { 0, 0, 0 };
                    for(u32 c = c_begin; c < c_end; ++c)
                    {
                        u32 n = chunk_spp;
                        if((c + 1) * chunk_spp > spp) n = spp - c * chunk_spp;
                        u32 series = ort_stream_seed(P->base_seed, pixel_index, c);
                        V3 color = trace_stream(*sc, cam, P, x, y, &series, n, queue, &cnt, brute);
                        acc[0] += to_fixed(color.x); acc[1] += to_fixed(color.y); acc[2] += to_fixed(color.z);
                    }
                    const double inv = 1.0 / (double)(1 << ORT_ACCUM_FRAC_BITS);
                    pixel->x = (f32)((double)acc[0] * inv) / (f32)spp;
                    pixel->y = (f32)((double)acc[1] * inv) / (f32)spp;
                    pixel->z = (f32)((double)acc[2] * inv) / (f32)spp;
                }
            }
        }
        c0 += cnt.rays; c1 += cnt.shape_tests; c2 += cnt.box_tests; c3 += cnt.node_visits;
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work);
    work();
    for(auto &t : threads) t.join();
    if(counters) { counters[0] += c0.load(); counters[1] += c1.load(); counters[2] += c2.load(); counters[3] += c3.load(); }
}

// Same semantics as ort_render_accumulate_device (include/ort_b200.h): ADDS the fixed-point
// chunk sums of [chunk_begin, chunk_end) into accum, an int64[height*width*4] buffer.
void oracle_render_accum(void *h, const OrtCamera *cam, const OrtRenderParams *P, int64_t *accum, int n_threads)
{
    OracleScene *sc = (OracleScene *)h;
    if(n_threads < 1) n_threads = 1;
    u32 spp = P->ray_per_pixel_count;
    u32 chunk_spp = P->chunk_spp ? P->chunk_spp : spp;
    if(chunk_spp > spp) chunk_spp = spp;
    u32 n_chunks = spp ? (spp + chunk_spp - 1) / chunk_spp : 0;
    u32 c_begin = P->chunk_begin, c_end = P->chunk_end;
    if(c_begin == 0 && c_end == 0) c_end = n_chunks;
    if(c_end > n_chunks) c_end = n_chunks;
    std::atomic<int> next_row(P->tile_min_y);
    auto work = [&]()
    {
        std::vector<const OrtBVHOctreeNode *> queue;
        Counters cnt = { 0, 0, 0, 0 };
        for(;;)
        {
            int y = next_row.fetch_add(1);
            if(y >= P->tile_one_past_max_y) break;
            for(int x = P->tile_min_x; x < P->tile_one_past_max_x; ++x)
            {
                u32 pixel_index = (u32)(y * P->output_width + x);
                int64_t *dst = accum + 4 * (size_t)pixel_index;
                for(u32 c = c_begin; c < c_end; ++c)
                {
                    u32 n = chunk_spp;
                    if((c + 1) * chunk_spp > spp) n = spp - c * chunk_spp;
                    u32 series = ort_stream_seed(P->base_seed, pixel_index, c);
                    V3 color = trace_stream(*sc, cam, P, x, y, &series, n, queue, &cnt, 0);
                    dst[0] += to_fixed(color.x); dst[1] += to_fixed(color.y); dst[2] += to_fixed(color.z);
                }
            }
        }
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work);
    work();
    for(auto &t : threads) t.join();
}

// resolve of the fixed-point buffer, as ort_accum_resolve_device
void oracle_accum_resolve(const int64_t *accum, int width, int height, uint32_t spp, ort_v3 *out)
{
    const double inv = 1.0 / (double)(1 << ORT_ACCUM_FRAC_BITS);
    for(size_t i = 0; i < (size_t)width * height; ++i)
    {
        out[i].x = (f32)((double)accum[4 * i + 0] * inv) / (f32)spp;
        out[i].y = (f32)((double)accum[4 * i + 1] * inv) / (f32)spp;
        out[i].z = (f32)((double)accum[4 * i + 2] * inv) / (f32)spp;
    }
}

// ---- single functions, same signatures as oracle/_ref's ref_* --------------
static void pack_isect(Isect r, float *out)
{
    out[0] = r.t; out[1] = r.n.x; out[2] = r.n.y; out[3] = r.n.z; out[4] = (float)r.inner;
}
void oracle_intersect_triangle(const float *v0, const float *v1, const float *v2, const float *o, const float *d, float *out)
{
    pack_isect(isect_triangle(mk(v0[0], v0[1], v0[2]), mk(v1[0], v1[1], v1[2]), mk(v2[0], v2[1], v2[2]), mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])), out);
}
void oracle_intersect_sphere(const float *c, float r, const float *o, const float *d, float *out)
{
    pack_isect(isect_sphere(mk(c[0], c[1], c[2]), r, mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])), out);
}
void oracle_intersect_aab(const float *mn, const float *mx, const float *o, const float *d, float *out)
{
    pack_isect(isect_aab(mk(mn[0], mn[1], mn[2]), mk(mx[0], mx[1], mx[2]), mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])), out);
}
void oracle_intersect_cylinder(const float *base, const float *axis, float r, const float *o, const float *d, float *out)
{
    pack_isect(isect_cylinder(mk(base[0], base[1], base[2]), mk(axis[0], axis[1], axis[2]), r, mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])), out);
}
int32_t oracle_in_rect(const float *p, const float *mn, const float *mx)
{
    return in_rect(mk(p[0], p[1], p[2]), mk(mn[0], mn[1], mn[2]), mk(mx[0], mx[1], mx[2]));
}
uint32_t oracle_xor_shift_32(uint32_t s) { xor_shift_32(&s); return s; }
float oracle_random_between_0_1(uint32_t *s) { return rnd01(s); }
float oracle_random_between(uint32_t *s, float mn, float mx) { return rnd_between(s, mn, mx); }
uint32_t oracle_random_between_u32(uint32_t *s, uint32_t mn, uint32_t mx) { return rnd_between_u32(s, mn, mx); }
void oracle_sample_brdf(uint32_t *s, const float *N, const float *wo, float roughness, const float *mat, float *wi_out, int32_t *is_t)
{
    SampleBRDF r = sample_brdf(s, mk(N[0], N[1], N[2]), mk(wo[0], wo[1], wo[2]), roughness,
                               mk(mat[0], mat[1], mat[2]), mk(mat[3], mat[4], mat[5]), mk(mat[6], mat[7], mat[8]), mat[9]);
    wi_out[0] = r.wi.x; wi_out[1] = r.wi.y; wi_out[2] = r.wi.z; *is_t = r.is_transmission;
}
float oracle_pdf_brdf(const float *N, const float *wi, const float *wo, float roughness, const float *mat)
{
    return pdf_brdf(mk(N[0], N[1], N[2]), mk(wi[0], wi[1], wi[2]), mk(wo[0], wo[1], wo[2]), roughness,
                    mk(mat[0], mat[1], mat[2]), mk(mat[3], mat[4], mat[5]), mk(mat[6], mat[7], mat[8]), mat[9]);
}
void oracle_eval_scattering(const float *N, const float *wi, const float *wo, const float *mat, float roughness, float distance, float *out)
{
    V3 r = eval_scattering(mk(N[0], N[1], N[2]), mk(wi[0], wi[1], wi[2]), mk(wo[0], wo[1], wo[2]),
                           mk(mat[0], mat[1], mat[2]), mk(mat[3], mat[4], mat[5]), mk(mat[6], mat[7], mat[8]), mat[9], roughness, distance);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
uint32_t oracle_sample_random_lights(void *h, uint32_t state)
{
    sample_random_lights_rng(*(OracleScene *)h, &state);
    return state;
}
int64_t oracle_to_fixed(float v) { return to_fixed(v); }

// Explicit ray buffers of BASELINE config 2, generated ONCE here and shared by every side:
//  kind 0, coherent: ray i = a lens-sampled primary of pixel (i % w, i / w) of the params' w x h grid, the
//          camera model of ray.cpp:1215-1237, lens angle from the reference xorshift (random.h) seeded per ray
//          with ort_stream_seed(seed, i, 0);
//  kind 1, incoherent: origin uniform in [lo, hi], direction uniform on the sphere, same per-ray streams.
void oracle_make_rays(int kind, const OrtCamera *cam, const OrtRenderParams *P, const float *lo, const float *hi,
                      uint32_t seed, uint64_t n, float *origins, float *dirs)
{
    for(uint64_t i = 0; i < n; ++i)
    {
        u32 s = ort_stream_seed(seed, (uint32_t)i, (uint32_t)(i >> 32));
        V3 o, d;
        if(kind == 0)
        {
            int x = (int)(i % (uint64_t)P->output_width), y = (int)(i / (uint64_t)P->output_width);
            f32 focal_length = len(sub(cam->p, mk(P->focus_target[0], P->focus_target[1], P->focus_target[2])));
            f32 pixel_x = (2.0f * x / (f32)P->output_width) - 1.0f;
            f32 pixel_y = (2.0f * y / (f32)P->output_height) - 1.0f;
            V3 camera_to_pixel = normalize(sub(add(mul(pixel_x, cam->x_axis), mul(pixel_y, cam->y_axis)), cam->z_axis));
            V3 focal_point = add(cam->p, mul(focal_length, camera_to_pixel));
            f32 random_rad = rnd_between(&s, 0.0f, 2 * PI_32);
            o = sub(add(add(cam->p, mul(P->aperture_radius * cosf(random_rad), cam->x_axis)),
                        mul(P->aperture_radius * sinf(random_rad), cam->y_axis)), mul(P->lens_z_offset, cam->z_axis));
            d = normalize(sub(focal_point, o));
        }
        else
        {
            o.x = lo[0] + (hi[0] - lo[0]) * rnd01(&s);
            o.y = lo[1] + (hi[1] - lo[1]) * rnd01(&s);
            o.z = lo[2] + (hi[2] - lo[2]) * rnd01(&s);
            f32 z = 2.0f * rnd01(&s) - 1.0f;
            f32 phi = 2.0f * PI_32 * rnd01(&s);
            f32 r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
            d = normalize(mk(r * cosf(phi), r * sinf(phi), z));
            if(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f) d = mk(0, 0, 1);
        }
        origins[3 * i] = o.x; origins[3 * i + 1] = o.y; origins[3 * i + 2] = o.z;
        dirs[3 * i] = d.x; dirs[3 * i + 1] = d.y; dirs[3 * i + 2] = d.z;
    }
}

} // extern "C"
