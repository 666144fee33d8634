// scene_flatten.cpp -- octree -> ranked records -> 8-wide quantised BVH (host).
//
// Input is the reference's own scene hand-off (World + BVHOctreeNode root,
// code/ray.h:76-133) exactly as main() leaves it (code/macos_main.mm:334-545).
// Nothing of the octree's topology survives except the ORDER it imposes on the
// leaf records -- the tie-break rank (bvh.h).  The acceleration structure that
// is uploaded is rebuilt from the records: binned-SAH binary BVH, collapsed to
// 8-wide, children placed in octant slots, boxes quantised outward.
#include "scene_flatten.h"
#include "bvh_build.h"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <chrono>
#include <deque>
#include <mutex>
#include <thread>

namespace ort {

namespace {

inline uint32_t payload_size(uint32_t type)
{
    switch(type)
    {
        case ORT_SHAPE_SPHERE:   return (uint32_t)sizeof(OrtSphere);
        case ORT_SHAPE_AAB:      return (uint32_t)sizeof(OrtAAB);
        case ORT_SHAPE_CYLINDER: return (uint32_t)sizeof(OrtCylinder);
        case ORT_SHAPE_TRIANGLE: return (uint32_t)sizeof(OrtTriangle);
        case ORT_SHAPE_CSG:      return (uint32_t)sizeof(OrtCSG);
        default: return 0;
    }
}

inline f3 v3(const ort_v3 &v) { return mk3(v.x, v.y, v.z); }

inline bool node_is_live(const OrtBVHOctreeNode *c)   // the test at ray.cpp:788
{
    return (c->is_leaf && c->push_buffer.used) || c->first_child;
}

inline void grow(float *lo, float *hi, double x, double y, double z)
{
    if(x < lo[0]) lo[0] = (float)x; if(y < lo[1]) lo[1] = (float)y; if(z < lo[2]) lo[2] = (float)z;
    if(x > hi[0]) hi[0] = (float)x; if(y > hi[1]) hi[1] = (float)y; if(z > hi[2]) hi[2] = (float)z;
}

// tight bounds of the region in which the reference's intersector for this
// record can report a hit (before padding)
void prim_bounds(HostPrim &p)
{
    for(int k = 0; k < 3; ++k) { p.lo[k] = FLT_MAX; p.hi[k] = -FLT_MAX; }
    switch(p.kind)
    {
        case PRIM_TRIANGLE:
            grow(p.lo, p.hi, p.a.x, p.a.y, p.a.z);
            grow(p.lo, p.hi, p.b.x, p.b.y, p.b.z);
            grow(p.lo, p.hi, p.c.x, p.c.y, p.c.z);
            break;
        case PRIM_SPHERE:
        {
            double r = fabs((double)p.radius);
            grow(p.lo, p.hi, p.a.x - r, p.a.y - r, p.a.z - r);
            grow(p.lo, p.hi, p.a.x + r, p.a.y + r, p.a.z + r);
        } break;
        case PRIM_AAB:
            grow(p.lo, p.hi, p.a.x, p.a.y, p.a.z);
            grow(p.lo, p.hi, p.b.x, p.b.y, p.b.z);
            break;
        case PRIM_CYLINDER:
        {
            // the intersector works in the frame R*(x - base) and accepts
            // x^2+y^2 <= r^2, 0 <= z <= |axis| there (ray.cpp:294-324); R falls
            // back to identity for axes parallel to Z (ray.cpp:13), whatever
            // their sign -- bound exactly that region, not the "geometric" cylinder
            exact::m3 rot = exact::rotation_matrix_along_z(p.b);
            double len = length(p.b), r = fabs((double)p.radius);
            for(int corner = 0; corner < 8; ++corner)
            {
                double lx = (corner & 1) ? r : -r, ly = (corner & 2) ? r : -r, lz = (corner & 4) ? len : 0.0;
                // world = base + R^T * local
                double wx = p.a.x + rot.r0.x * lx + rot.r1.x * ly + rot.r2.x * lz;
                double wy = p.a.y + rot.r0.y * lx + rot.r1.y * ly + rot.r2.y * lz;
                double wz = p.a.z + rot.r0.z * lx + rot.r1.z * ly + rot.r2.z * lz;
                grow(p.lo, p.hi, wx, wy, wz);
            }
        } break;
    }
}

// a record whose bounds are not finite (NaN / Inf vertices or parameters) cannot be placed in any tree: the
// builders would loop on an infinite extent or cast NaN to an integer.  Rejected with ORT_ERR_ARG instead.
bool prim_is_finite(const HostPrim &p)
{
    auto ok = [](float v) { return v >= -1e30f && v <= 1e30f; };          // false for NaN and +-Inf
    auto ok3 = [&](f3 v) { return ok(v.x) && ok(v.y) && ok(v.z); };
    // the parameters themselves (min / max folds drop NaNs, so the bounds alone would not show them) ...
    if(!ok3(p.a)) return false;
    if((p.kind == PRIM_TRIANGLE || p.kind == PRIM_AAB || p.kind == PRIM_CYLINDER) && !ok3(p.b)) return false;
    if(p.kind == PRIM_TRIANGLE && !ok3(p.c)) return false;
    if((p.kind == PRIM_SPHERE || p.kind == PRIM_CYLINDER) && !ok(p.radius)) return false;
    // ... and the bounds derived from them
    for(int k = 0; k < 3; ++k)
        if(!ok(p.lo[k]) || !ok(p.hi[k])) return false;
    return true;
}

} // namespace

// the per-material constants of DevMaterial, exactly as path.h's pdf_brdf / sample_brdf /
// eval_scattering form them (float arithmetic, this file is compiled with -ffp-contract=off)
static void material_constants(DevMaterial *dm)
{
    f3 Kd = mk3(dm->diffuse[0], dm->diffuse[1], dm->diffuse[2]);
    f3 Ks = mk3(dm->specular[0], dm->specular[1], dm->specular[2]);
    f3 Kt = mk3(dm->transmission[0], dm->transmission[1], dm->transmission[2]);
    float Kd_l = length(Kd), Ks_l = length(Ks), Kt_l = length(Kt);
    float s = Kd_l + Ks_l + Kt_l;
    dm->pd_c = Kd_l / s; dm->ps_c = Ks_l / s; dm->pt_c = Kt_l / s;
    f3 Ed = Kd / ORT_PI_32;
    dm->ed[0] = Ed.x; dm->ed[1] = Ed.y; dm->ed[2] = Ed.z;
    dm->lobes = (length_square(Kd) > 0.0f ? 1u : 0u) | (length_square(Ks) > 0.0f ? 2u : 0u) | (length_square(Kt) > 0.0f ? 4u : 0u);
}

void material_constants_public(DevMaterial *dm) { material_constants(dm); }   // ort_tools.cu (device BSDF harness)

// world-level tables of the flattened scene: materials and the light list
static int fill_world_tables(const OrtWorld *world, FlatScene *out, std::string *err)
{
    // materials (index 0 = "miss", parser.cpp:1187)
    out->materials.resize(world->mat_count ? world->mat_count : 1);
    memset(out->materials.data(), 0, out->materials.size() * sizeof(DevMaterial));
    for(uint32_t i = 0; i < world->mat_count; ++i)
    {
        const OrtMaterial &m = world->materials[i];
        DevMaterial &dm = out->materials[i];
        dm.diffuse[0] = m.diffuse.x; dm.diffuse[1] = m.diffuse.y; dm.diffuse[2] = m.diffuse.z;
        dm.specular[0] = m.specular.x; dm.specular[1] = m.specular.y; dm.specular[2] = m.specular.z;
        dm.transmission[0] = m.transmission.x; dm.transmission[1] = m.transmission.y; dm.transmission[2] = m.transmission.z;
        dm.emit[0] = m.emit_color.x; dm.emit[1] = m.emit_color.y; dm.emit[2] = m.emit_color.z;
        dm.ior = m.ior;
        dm.is_light = m.is_light;
        material_constants(&dm);
    }
    out->info.material_count = world->mat_count;

    // light list: packed (u32 ShapeType, pointer) pairs (push_light, parser.cpp:1144-1182).
    // Only "is this entry a sphere" matters: it decides how many RNG steps
    // sample_random_lights consumes (ray.cpp:542, 562).
    out->light_is_sphere.clear();
    const uint8_t *lb = (const uint8_t *)world->light_push_buffer.base;
    const size_t stride = 4 + sizeof(void *);
    for(size_t consumed = 0; lb && consumed + stride <= world->light_push_buffer.used; consumed += stride)
    {
        uint32_t type; memcpy(&type, lb + consumed, 4);
        out->light_is_sphere.push_back(type == ORT_SHAPE_SPHERE ? 1 : 0);
    }
    if(out->light_is_sphere.size() != world->light_count)
    {
        *err = "light push buffer does not hold light_count entries";
        return ORT_ERR_ARG;
    }
    out->info.light_count = world->light_count;
    return ORT_OK;
}

int collect_records(const OrtWorld *world, const OrtBVHOctreeNode *root,
                    std::vector<HostPrim> *prims, FlatScene *out, std::string *err)
{
    if(!world || !root) { *err = "null world / top_most_node"; return ORT_ERR_ARG; }
    memset(&out->info, 0, sizeof(out->info));

    // full breadth-first walk: the order raycast_bvh pops nodes when no child is
    // culled (queue push at ray.cpp:808, children 0..7, dead children skipped :788)
    struct Item { const OrtBVHOctreeNode *node; uint32_t depth; };
    std::deque<Item> queue;
    queue.push_back(Item{ root, 0u });
    uint32_t rank = 0;
    while(!queue.empty())
    {
        Item it = queue.front(); queue.pop_front();
        const OrtBVHOctreeNode *node = it.node;
        out->info.octree_node_count++;
        if(it.depth > out->info.octree_max_depth) out->info.octree_max_depth = it.depth;
        const uint8_t *base = (const uint8_t *)node->push_buffer.base;
        for(size_t consumed = 0; consumed < node->push_buffer.used; )
        {
            uint32_t type; memcpy(&type, base + consumed, 4);
            uint32_t psize = payload_size(type);
            if(psize == 0 || consumed + 4 + psize > node->push_buffer.used)
            {
                *err = "corrupt leaf push buffer (unknown shape type)";
                return ORT_ERR_ARG;
            }
            const uint8_t *q = base + consumed + 4;
            consumed += 4 + psize;
            HostPrim p; memset(&p, 0, sizeof(p));
            p.rank = rank++;
            bool keep = true;
            switch(type)
            {
                case ORT_SHAPE_SPHERE:
                {
                    OrtSphere s; memcpy(&s, q, sizeof(s));
                    p.kind = PRIM_SPHERE; p.a = v3(s.center); p.radius = s.r; p.mat = s.mat_index;
                    out->info.sphere_count++;
                } break;
                case ORT_SHAPE_AAB:
                {
                    OrtAAB s; memcpy(&s, q, sizeof(s));
                    p.kind = PRIM_AAB; p.a = v3(s.min); p.b = v3(s.max); p.mat = s.mat_index;
                    out->info.box_count++;
                } break;
                case ORT_SHAPE_CYLINDER:
                {
                    OrtCylinder s; memcpy(&s, q, sizeof(s));
                    p.kind = PRIM_CYLINDER; p.a = v3(s.base); p.b = v3(s.axis); p.radius = s.r; p.mat = s.mat_index;
                    out->info.cylinder_count++;
                } break;
                case ORT_SHAPE_TRIANGLE:
                {
                    OrtTriangle s; memcpy(&s, q, sizeof(s));
                    if(!s.mesh || !s.mesh->vertices) { *err = "triangle record without mesh"; return ORT_ERR_ARG; }
                    if(s.i_0 >= s.mesh->vertex_count || s.i_1 >= s.mesh->vertex_count || s.i_2 >= s.mesh->vertex_count)
                    { *err = "triangle record with vertex index out of range"; return ORT_ERR_ARG; }
                    p.kind = PRIM_TRIANGLE;
                    p.a = v3(s.mesh->vertices[s.i_0]);       // ray.cpp:702-704
                    p.b = v3(s.mesh->vertices[s.i_1]);
                    p.c = v3(s.mesh->vertices[s.i_2]);
                    p.mat = s.mesh->mat_index;
                    out->info.triangle_count++;
                } break;
                case ORT_SHAPE_CSG:
                    keep = false;                           // inert record, ray.cpp:718-767
                    out->info.csg_count++;
                    break;
                default:
                    keep = false;
                    break;
            }
            if(keep)
            {
                if(p.mat >= world->mat_count) { *err = "record with material index out of range"; return ORT_ERR_ARG; }
                prim_bounds(p);
                if(!prim_is_finite(p)) { *err = "record with non-finite geometry (NaN / Inf)"; return ORT_ERR_ARG; }
                prims->push_back(p);
            }
        }
        if(node->first_child)
            for(uint32_t ci = 0; ci < 8; ++ci)
            {
                const OrtBVHOctreeNode *child = node->first_child + ci;
                if(node_is_live(child)) queue.push_back(Item{ child, it.depth + 1 });
            }
    }
    out->info.record_count = rank;
    out->info.root_min[0] = root->aabb_min.x; out->info.root_min[1] = root->aabb_min.y; out->info.root_min[2] = root->aabb_min.z;
    out->info.root_max[0] = root->aabb_max.x; out->info.root_max[1] = root->aabb_max.y; out->info.root_max[2] = root->aabb_max.z;

    return fill_world_tables(world, out, err);
}

// ---------------------------------------------------------------------------
// binary BVH, binned SAH
// ---------------------------------------------------------------------------
namespace {

struct Box
{
    float lo[3], hi[3];
    void reset() { for(int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; } }
    void grow(const float *l, const float *h) { for(int k = 0; k < 3; ++k) { if(l[k] < lo[k]) lo[k] = l[k]; if(h[k] > hi[k]) hi[k] = h[k]; } }
    void grow_pt(const float *c) { for(int k = 0; k < 3; ++k) { if(c[k] < lo[k]) lo[k] = c[k]; if(c[k] > hi[k]) hi[k] = c[k]; } }
    double area() const
    {
        double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        if(dx < 0 || dy < 0 || dz < 0) return 0.0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

struct B2Node
{
    Box box;
    uint32_t left, right;     // children (inner)
    uint32_t first, count;    // leaf range in the index permutation (count > 0 = leaf)
    uint32_t span;            // primitives below the node: idx[first .. first + span)
};

struct Builder
{
    const std::vector<HostPrim> &prims;
    const BuildOptions &opt;
    std::vector<uint32_t> idx;
    std::vector<float> cen;        // 3 per prim
    std::vector<B2Node> nodes;

    Builder(const std::vector<HostPrim> &p, const BuildOptions &o) : prims(p), opt(o) {}

    struct Job { uint32_t node, first, count; };

    // Binned-SAH split of one node.  Returns false for a leaf; else fills the two child jobs.
    bool split(const Job &job, std::atomic<uint32_t> &next_node, Job *left, Job *right)
    {
        const int NBINS = 16;
        Box box, cbox; box.reset(); cbox.reset();
        for(uint32_t i = job.first; i < job.first + job.count; ++i)
        {
            const HostPrim &p = prims[idx[i]];
            box.grow(p.lo, p.hi);
            cbox.grow_pt(&cen[3 * idx[i]]);
        }
        B2Node &nd = nodes[job.node];
        nd.box = box; nd.left = nd.right = 0; nd.first = job.first; nd.count = job.count; nd.span = job.count;
        if(job.count <= 1) return false;

        double best_cost = 1e300; int best_axis = -1, best_bin = -1;
        double parent_area = box.area();
        for(int axis = 0; axis < 3; ++axis)
        {
            float c0 = cbox.lo[axis], c1 = cbox.hi[axis];
            if(!(c1 > c0)) continue;
            Box bins[NBINS]; uint32_t cnt[NBINS];
            for(int b = 0; b < NBINS; ++b) { bins[b].reset(); cnt[b] = 0; }
            float scale = (float)NBINS / (c1 - c0);
            for(uint32_t i = job.first; i < job.first + job.count; ++i)
            {
                uint32_t pi = idx[i];
                int b = (int)((cen[3 * pi + axis] - c0) * scale);
                if(b < 0) b = 0; if(b >= NBINS) b = NBINS - 1;
                bins[b].grow(prims[pi].lo, prims[pi].hi); cnt[b]++;
            }
            double la[NBINS]; uint32_t ln[NBINS];
            Box acc; acc.reset(); uint32_t c = 0;
            for(int b = 0; b < NBINS - 1; ++b)
            {
                if(cnt[b]) acc.grow(bins[b].lo, bins[b].hi);
                c += cnt[b]; la[b] = acc.area(); ln[b] = c;
            }
            acc.reset(); c = 0;
            for(int b = NBINS - 1; b > 0; --b)
            {
                if(cnt[b]) acc.grow(bins[b].lo, bins[b].hi);
                c += cnt[b];
                if(ln[b - 1] == 0 || c == 0) continue;
                double cost = la[b - 1] * ln[b - 1] + acc.area() * c;
                if(cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        double leaf_cost = (double)job.count;
        double split_cost = (best_axis >= 0 && parent_area > 0.0) ? opt.traversal_cost + best_cost / parent_area : 1e300;
        if(job.count <= opt.max_leaf && leaf_cost <= split_cost) return false;   // leaf

        uint32_t mid;
        if(best_axis >= 0)
        {
            float c0 = cbox.lo[best_axis], c1 = cbox.hi[best_axis];
            float scale = (float)NBINS / (c1 - c0);
            uint32_t *b = &idx[job.first], *e = b + job.count;
            int axis = best_axis, bin = best_bin;
            uint32_t *m = std::partition(b, e, [&](uint32_t pi)
            {
                int bb = (int)((cen[3 * pi + axis] - c0) * scale);
                if(bb < 0) bb = 0; if(bb >= NBINS) bb = NBINS - 1;
                return bb < bin;
            });
            mid = (uint32_t)(m - &idx[0]);
        }
        else mid = job.first + job.count / 2;          // all centroids coincide: split the range in the middle
        if(mid == job.first || mid == job.first + job.count) mid = job.first + job.count / 2;

        uint32_t l = next_node.fetch_add(2);
        nd.left = l; nd.right = l + 1; nd.count = 0;
        *left = Job{ l, job.first, mid - job.first };
        *right = Job{ l + 1, mid, job.first + job.count - mid };
        return true;
    }

    // Top-down build; subtrees are independent (disjoint index ranges, nodes pre-allocated), so
    // worker threads take them from a shared queue.  The tree does not depend on the schedule.
    void build()
    {
        size_t n = prims.size();
        idx.resize(n); cen.resize(3 * n);
        for(size_t i = 0; i < n; ++i)
        {
            idx[i] = (uint32_t)i;
            for(int k = 0; k < 3; ++k) cen[3 * i + k] = 0.5f * (prims[i].lo[k] + prims[i].hi[k]);
        }
        nodes.assign(2 * n + 1, B2Node());
        std::atomic<uint32_t> next_node(1);
        std::mutex mu;
        std::condition_variable cv;
        std::vector<Job> queue;
        size_t outstanding = 1;                 // jobs queued or being processed
        queue.push_back(Job{ 0u, 0u, (uint32_t)n });
        const uint32_t share_above = 32768;     // subtrees larger than this are offered to other threads
        auto worker = [&]()
        {
            std::vector<Job> local;
            for(;;)
            {
                Job job;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return !queue.empty() || outstanding == 0; });
                    if(queue.empty()) return;
                    job = queue.back(); queue.pop_back();
                }
                local.push_back(job);
                while(!local.empty())
                {
                    Job j = local.back(); local.pop_back();
                    Job l, r;
                    if(!split(j, next_node, &l, &r)) continue;
                    Job small = l.count <= r.count ? l : r, big = l.count <= r.count ? r : l;
                    local.push_back(small);
                    if(big.count > share_above)
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        queue.push_back(big); ++outstanding;
                        cv.notify_one();
                    }
                    else local.push_back(big);
                }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    --outstanding;
                    if(outstanding == 0) cv.notify_all();
                }
            }
        };
        unsigned nt = std::thread::hardware_concurrency();
        if(nt == 0) nt = 1;
        if(nt > 32) nt = 32;
        if(n < 200000) nt = 1;
        std::vector<std::thread> threads;
        for(unsigned i = 1; i < nt; ++i) threads.emplace_back(worker);
        worker();
        for(auto &t : threads) t.join();
        nodes.resize(next_node.load());
    }
};

// per-axis power-of-two grid step: smallest 2^e with 255 * 2^e >= extent
inline int grid_exponent(double extent)
{
    if(!(extent > 0.0)) return -100;
    int e = (int)ceil(log2(extent / 255.0));
    if(e < -100) e = -100;
    if(e > 120) e = 120;
    return e;
}

} // namespace

namespace {

// the 48-byte device record of one primitive (bvh.h); a cylinder also gets its CylinderAux entry
PrimRec make_record(const HostPrim &hp, FlatScene *out)
{
    PrimRec r; memset(&r, 0, sizeof(r));
    r.rank = hp.rank; r.mat = hp.mat;
    r.ax = hp.a.x; r.ay = hp.a.y; r.az = hp.a.z;
    switch(hp.kind)
    {
        case PRIM_TRIANGLE:
            r.bx = hp.b.x; r.by = hp.b.y; r.bz = hp.b.z;
            r.cx = hp.c.x; r.cy = hp.c.y; r.cz = hp.c.z;
            r.kind = PRIM_TRIANGLE;
            break;
        case PRIM_SPHERE:
            r.bx = hp.radius; r.kind = PRIM_SPHERE;
            break;
        case PRIM_AAB:
            r.bx = hp.b.x; r.by = hp.b.y; r.bz = hp.b.z; r.kind = PRIM_AAB;
            break;
        case PRIM_CYLINDER:
        {
            CylinderAux ca; memset(&ca, 0, sizeof(ca));
            exact::m3 rot = exact::rotation_matrix_along_z(hp.b);     // ray.cpp:295
            ca.base[0] = hp.a.x; ca.base[1] = hp.a.y; ca.base[2] = hp.a.z;
            ca.axis_len = length(hp.b);                               // ray.cpp:302
            ca.radius = hp.radius;
            ca.r0[0] = rot.r0.x; ca.r0[1] = rot.r0.y; ca.r0[2] = rot.r0.z;
            ca.r1[0] = rot.r1.x; ca.r1[1] = rot.r1.y; ca.r1[2] = rot.r1.z;
            ca.r2[0] = rot.r2.x; ca.r2[1] = rot.r2.y; ca.r2[2] = rot.r2.z;
            r.bx = hp.b.x; r.by = hp.b.y; r.bz = hp.b.z;
            r.kind = PRIM_CYLINDER | ((uint32_t)out->cylinders.size() << 8);
            out->cylinders.push_back(ca);
        } break;
    }
    return r;
}

// builds one wide tree over `prims` and appends it to out->nodes / prims /
// cylinders; the tree's root is the first node appended
int emit_tree(const std::vector<HostPrim> &prims, const BuildOptions &opt, FlatScene *out,
              uint32_t *depth_out, std::string *err)
{
    *depth_out = 1;
    if(prims.empty())
    {
        // a single empty node: every ray misses
        WideNode n; memset(&n, 0, sizeof(n));
        n.ex = n.ey = n.ez = 127;
        for(int s = 0; s < 8; ++s) { n.qlo_x[s] = n.qlo_y[s] = n.qlo_z[s] = 255; }
        out->nodes.push_back(n);
        return ORT_OK;
    }

    const bool timing = getenv("ORT_TIMING") != 0;
    auto t0 = std::chrono::steady_clock::now();
    // With the optimal collapse the binary tree is built down to single primitives and the
    // dynamic programme below decides which subtrees become leaves.
    BuildOptions bopt = opt;
    if(opt.optimal_collapse) bopt.max_leaf = 1;
    Builder b(prims, bopt);
    b.build();
    auto t1 = std::chrono::steady_clock::now();
    if(timing) fprintf(stderr, "[ort] binned-SAH build of %zu records: %.2f s\n", prims.size(), std::chrono::duration<double>(t1 - t0).count());

    // Optimal collapse to 8-wide (Ylitie, Karras, Laine 2017, section 3): C(n, i) = least SAH cost of
    // representing binary subtree n with at most i child slots of a wide node,
    //   C(n, 1)     = min( A_n * P_n          [a leaf of P_n <= max_leaf primitives],
    //                      A_n * c_node + D(n, 8)  [a wide node of its own] )
    //   C(n, i > 1) = min( D(n, i), C(n, i - 1) ),   D(n, j) = min_{0<k<j} C(left, k) + C(right, j - k)
    // evaluated bottom-up (children carry larger indices than their parent); the emission below follows
    // the minimising choices instead of greedily opening the largest child.
    const size_t NB = b.nodes.size();
    std::vector<float> C;            // C[8 * n + i], i = 1..7
    std::vector<uint8_t> as_leaf, dist_k;     // dist_k[9 * n + j], j = 2..8
    if(opt.optimal_collapse)
    {
        C.assign(8 * NB, 0.f); as_leaf.assign(NB, 0); dist_k.assign(9 * NB, 0);
        const double c_node = opt.wide_node_cost;
        for(size_t n = NB; n-- > 0; )
        {
            const B2Node &nd = b.nodes[n];
            const double A = nd.box.area(), P = (double)nd.span;
            if(nd.count > 0)
            {
                for(int i = 1; i <= 7; ++i) C[8 * n + i] = (float)(A * P);
                as_leaf[n] = 1;
                continue;
            }
            double D[9];
            for(int j = 2; j <= 8; ++j)
            {
                double best = 1e300; int bk = 1;
                for(int k = 1; k < j; ++k)
                {
                    int kl = k > 7 ? 7 : k, kr = (j - k) > 7 ? 7 : (j - k);
                    double v = (double)C[8 * nd.left + kl] + (double)C[8 * nd.right + kr];
                    if(v < best) { best = v; bk = k; }
                }
                D[j] = best; dist_k[9 * n + j] = (uint8_t)bk;
            }
            double c_leaf = nd.span <= opt.max_leaf ? A * P : 1e300;
            double c_int = D[8] + A * c_node;
            as_leaf[n] = c_leaf <= c_int ? 1 : 0;
            C[8 * n + 1] = (float)std::min(c_leaf, c_int);
            for(int i = 2; i <= 7; ++i) C[8 * n + i] = std::min((float)D[i], C[8 * n + i - 1]);
        }
    }
    auto kid_is_leaf = [&](uint32_t id) { return opt.optimal_collapse ? as_leaf[id] != 0 : b.nodes[id].count > 0; };

    struct WItem { uint32_t b2; uint32_t depth; };
    std::deque<WItem> queue;
    size_t next_out = out->nodes.size();     // index of the wide node being filled (BFS order == allocation order)
    out->nodes.push_back(WideNode());
    queue.push_back(WItem{ 0u, 1u });
    uint32_t wide_depth = 0;

    while(!queue.empty())
    {
        WItem it = queue.front(); queue.pop_front();
        size_t self = next_out++;
        if(it.depth > wide_depth) wide_depth = it.depth;

        // gather up to 8 children: by the dynamic programme's choices, or by repeatedly opening the largest inner child
        uint32_t kids[8]; int nk = 0;
        const B2Node &rootn = b.nodes[it.b2];
        if(kid_is_leaf(it.b2)) { kids[nk++] = it.b2; }          // degenerate: the whole (sub)tree is one leaf
        else if(opt.optimal_collapse)
        {
            struct Slot { uint32_t node; int budget; };
            Slot stack[16]; int sp = 0;
            int k = dist_k[9 * it.b2 + 8];
            stack[sp++] = Slot{ rootn.right, 8 - k };
            stack[sp++] = Slot{ rootn.left, k };
            while(sp > 0)
            {
                Slot c = stack[--sp];
                int i = c.budget > 7 ? 7 : c.budget;
                while(i > 1 && C[8 * c.node + i] == C[8 * c.node + i - 1]) --i;
                if(i == 1 || b.nodes[c.node].count > 0) { kids[nk++] = c.node; continue; }
                int kk = dist_k[9 * c.node + i];
                stack[sp++] = Slot{ b.nodes[c.node].right, i - kk };
                stack[sp++] = Slot{ b.nodes[c.node].left, kk };
            }
        }
        else { kids[nk++] = rootn.left; kids[nk++] = rootn.right; }
        for(; !opt.optimal_collapse; )
        {
            if(nk >= 8) break;
            int pick = -1; double pa = -1.0;
            for(int i = 0; i < nk; ++i)
            {
                const B2Node &c = b.nodes[kids[i]];
                if(c.count == 0) { double a = c.box.area(); if(a > pa) { pa = a; pick = i; } }
            }
            if(pick < 0) break;
            const B2Node &c = b.nodes[kids[pick]];
            kids[pick] = c.left;
            kids[nk++] = c.right;
        }

        Box nb; nb.reset();
        for(int i = 0; i < nk; ++i) nb.grow(b.nodes[kids[i]].box.lo, b.nodes[kids[i]].box.hi);

        // octant slots: greedy assignment maximising (centroid - centre) . octant direction
        int slot_of[8]; bool slot_used[8] = { false }; bool kid_done[8] = { false };
        double ctr[3]; for(int k = 0; k < 3; ++k) ctr[k] = 0.5 * ((double)nb.lo[k] + nb.hi[k]);
        for(int round = 0; round < nk; ++round)
        {
            double bestc = -1e300; int bi = -1, bs = -1;
            for(int i = 0; i < nk; ++i)
            {
                if(kid_done[i]) continue;
                const Box &cb = b.nodes[kids[i]].box;
                double dc[3]; for(int k = 0; k < 3; ++k) dc[k] = 0.5 * ((double)cb.lo[k] + cb.hi[k]) - ctr[k];
                for(int s = 0; s < 8; ++s)
                {
                    if(slot_used[s]) continue;
                    double c = ((s & 1) ? dc[0] : -dc[0]) + ((s & 2) ? dc[1] : -dc[1]) + ((s & 4) ? dc[2] : -dc[2]);
                    if(c > bestc) { bestc = c; bi = i; bs = s; }
                }
            }
            slot_of[bi] = bs; slot_used[bs] = true; kid_done[bi] = true;
        }
        int kid_at[8]; for(int s = 0; s < 8; ++s) kid_at[s] = -1;
        for(int i = 0; i < nk; ++i) kid_at[slot_of[i]] = i;

        WideNode n; memset(&n, 0, sizeof(n));
        n.px = nb.lo[0]; n.py = nb.lo[1]; n.pz = nb.lo[2];
        int e[3]; double step[3];
        for(int k = 0; k < 3; ++k) e[k] = grid_exponent((double)nb.hi[k] - nb.lo[k]);
        // quantise outward; if rounding pushes a plane past 255, coarsen the grid
        uint8_t qlo[3][8], qhi[3][8];
        for(int k = 0; k < 3; ++k)
        {
            for(;;)
            {
                step[k] = ldexp(1.0, e[k]);
                bool ok = true;
                for(int s = 0; s < 8 && ok; ++s)
                {
                    if(kid_at[s] < 0) { qlo[k][s] = 255; qhi[k][s] = 0; continue; }
                    const Box &cb = b.nodes[kids[kid_at[s]]].box;
                    double lo = floor(((double)cb.lo[k] - (double)nb.lo[k]) / step[k]);
                    double hi = ceil(((double)cb.hi[k] - (double)nb.lo[k]) / step[k]);
                    if(lo < 0.0) lo = 0.0;
                    if(hi > 255.0) { ok = false; break; }
                    if(hi < lo) hi = lo;
                    qlo[k][s] = (uint8_t)lo; qhi[k][s] = (uint8_t)hi;
                }
                if(ok) break;
                e[k]++;
            }
        }
        n.ex = (uint8_t)(e[0] + 127); n.ey = (uint8_t)(e[1] + 127); n.ez = (uint8_t)(e[2] + 127);
        for(int s = 0; s < 8; ++s)
        {
            n.qlo_x[s] = qlo[0][s]; n.qlo_y[s] = qlo[1][s]; n.qlo_z[s] = qlo[2][s];
            n.qhi_x[s] = qhi[0][s]; n.qhi_y[s] = qhi[1][s]; n.qhi_z[s] = qhi[2][s];
        }

        n.child_base = (uint32_t)out->nodes.size();
        n.prim_base = (uint32_t)out->prims.size();
        uint32_t prim_off = 0;
        bool only_triangles = true;
        for(int s = 0; s < 8; ++s)
        {
            if(kid_at[s] < 0) continue;
            const B2Node &c = b.nodes[kids[kid_at[s]]];
            const uint32_t c_count = kid_is_leaf(kids[kid_at[s]]) ? c.span : 0u;
            if(c_count == 0)
            {
                n.imask |= (uint8_t)(1u << s);
                n.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
                out->nodes.push_back(WideNode());
                queue.push_back(WItem{ kids[kid_at[s]], it.depth + 1 });
            }
            else
            {
                if(c_count > 3 || prim_off + c_count > 24) { *err = "internal: leaf too large"; return ORT_ERR_LIMIT; }
                uint32_t unary = (1u << c_count) - 1u;
                n.meta[s] = (uint8_t)((unary << 5) | prim_off);
                for(uint32_t i = 0; i < c_count; ++i)
                {
                    const HostPrim &hp = prims[b.idx[c.first + i]];
                    if(hp.kind != PRIM_TRIANGLE) only_triangles = false;
                    PrimRec r = make_record(hp, out);
                    out->prims.push_back(r);
                }
                prim_off += c_count;
            }
        }
        if(n.prim_base >= ORT_NODE_MIXED_KINDS) { *err = "too many primitive records"; return ORT_ERR_LIMIT; }
        if(!only_triangles) n.prim_base |= ORT_NODE_MIXED_KINDS;
        out->nodes[self] = n;
    }

    *depth_out = wide_depth;
    if(timing) fprintf(stderr, "[ort] collapse to 8-wide + quantise + emit: %.2f s\n",
                       std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    return ORT_OK;
}

} // namespace

// ---- records and ranks without the octree ---------------------------------------------------------------
namespace {

inline f3 gmin(f3 a, f3 b) { return mk3(ref_min(a.x, b.x), ref_min(a.y, b.y), ref_min(a.z, b.z)); }   // math.h:1085
inline f3 gmax(f3 a, f3 b) { return mk3(ref_max(a.x, b.x), ref_max(a.y, b.y), ref_max(a.z, b.z)); }   // math.h:1097

// centre of get_shape_aabb (ray.cpp:1675-1746): the point push_shape_inside_node sorts into octants
inline f3 box_centre(f3 mn, f3 mx) { return 0.5f * (mn + mx); }
inline f3 centre_of_cylinder(const OrtCylinder &c)
{
    f3 base = v3(c.base), axis = v3(c.axis);
    f3 other = base + axis;
    f3 q = hadamard(axis, axis) / dot(axis, axis);
    f3 e = c.r * (mk3(1, 1, 1) - mk3(sqrtf(q.x), sqrtf(q.y), sqrtf(q.z)));
    return box_centre(gmin(base - e, other - e), gmax(base + e, other + e));
}

// stable LSD radix sort of (key, payload) pairs on the low `bits` bits of the key
void radix_sort_pairs(std::vector<uint64_t> &key, std::vector<uint32_t> &val, int bits)
{
    size_t n = key.size();
    std::vector<uint64_t> k2(n); std::vector<uint32_t> v2(n);
    for(int shift = 0; shift < bits; shift += 11)
    {
        size_t count[2049]; memset(count, 0, sizeof(count));
        for(size_t i = 0; i < n; ++i) count[((key[i] >> shift) & 2047u) + 1]++;
        for(int b = 0; b < 2048; ++b) count[b + 1] += count[b];
        for(size_t i = 0; i < n; ++i)
        {
            size_t d = count[(key[i] >> shift) & 2047u]++;
            k2[d] = key[i]; v2[d] = val[i];
        }
        key.swap(k2); val.swap(v2);
    }
}

} // namespace

int collect_records_from_lists(const OrtWorld *world, const OrtShapeLists *L,
                               std::vector<HostPrim> *prims, FlatScene *out, std::string *err)
{
    if(!world || !L) { *err = "null world / shape lists"; return ORT_ERR_ARG; }
    memset(&out->info, 0, sizeof(out->info));
    // insertion order (macos_main.mm:474-538): triangles mesh by mesh, cylinders, boxes, spheres, CSG
    std::vector<uint64_t> mesh_first(L->mesh_count + 1, 0);
    for(uint32_t m = 0; m < L->mesh_count; ++m)
    {
        const OrtMesh &mesh = L->meshes[m];
        if(mesh.index_count && (!mesh.vertices || !mesh.indices)) { *err = "mesh without vertices / indices"; return ORT_ERR_ARG; }
        mesh_first[m + 1] = mesh_first[m] + mesh.index_count / 3u;
    }
    const uint64_t n_tri = mesh_first[L->mesh_count];
    const uint64_t first_cyl = n_tri, first_box = first_cyl + L->cylinder_count, first_sph = first_box + L->box_count;
    const uint64_t first_csg = first_sph + L->sphere_count, n = first_csg + (L->csg ? 1u : 0u);
    if(n >= 0x7FFFFFFFull) { *err = "too many records"; return ORT_ERR_LIMIT; }

    const f3 root_min = v3(L->root_min), root_max = v3(L->root_max);
    const f3 root_center = 0.5f * (root_min + root_max);            // macos_main.mm:470-471
    const f3 root_half = root_max - root_center;

    // 1. octant path of every record, in insertion order
    std::vector<uint64_t> key(n);
    std::vector<uint32_t> who(n);
    {
        const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for(unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t]()
            {
                for(uint32_t m = 0; m < L->mesh_count; ++m)
                {
                    const OrtMesh &mesh = L->meshes[m];
                    uint64_t cnt = mesh.index_count / 3u;
                    for(uint64_t k = t; k < cnt; k += nt)
                    {
                        f3 a = v3(mesh.vertices[mesh.indices[3 * k]]), b = v3(mesh.vertices[mesh.indices[3 * k + 1]]), c = v3(mesh.vertices[mesh.indices[3 * k + 2]]);
                        uint64_t i = mesh_first[m] + k;
                        key[i] = build::octant_path(build::triangle_centre(a, b, c), root_center, root_half);
                    }
                }
            });
        for(auto &th : pool) th.join();
    }
    for(uint32_t i = 0; i < L->cylinder_count; ++i) key[first_cyl + i] = build::octant_path(centre_of_cylinder(L->cylinders[i]), root_center, root_half);
    for(uint32_t i = 0; i < L->box_count; ++i) key[first_box + i] = build::octant_path(box_centre(v3(L->boxes[i].min), v3(L->boxes[i].max)), root_center, root_half);
    for(uint32_t i = 0; i < L->sphere_count; ++i) key[first_sph + i] = build::octant_path(v3(L->spheres[i].center), root_center, root_half);
    if(L->csg) key[first_csg] = build::octant_path(box_centre(v3(L->csg->aabb_min), v3(L->csg->aabb_max)), root_center, root_half);
    for(uint64_t i = 0; i < n; ++i) who[i] = (uint32_t)i;

    // 2. the node of a record is the shortest prefix of its path no other record shares (capped at 10)
    std::vector<uint64_t> sorted_key(key);
    std::vector<uint32_t> sorted_who(who);
    radix_sort_pairs(sorted_key, sorted_who, 30);
    std::vector<uint8_t> depth(n);
    uint32_t max_depth = 0;
    for(uint64_t p = 0; p < n; ++p)
    {
        int shared = 0;
        if(p > 0) shared = std::max(shared, build::common_levels((uint32_t)sorted_key[p], (uint32_t)sorted_key[p - 1]));
        if(p + 1 < n) shared = std::max(shared, build::common_levels((uint32_t)sorted_key[p], (uint32_t)sorted_key[p + 1]));
        int d = n == 1 ? 0 : std::min(shared + 1, 10);
        depth[sorted_who[p]] = (uint8_t)d;
        if((uint32_t)d > max_depth) max_depth = (uint32_t)d;
    }
    // octree statistics, as collect_records reports them (nodes reachable through live children):
    // every prefix shared by two or more records is an inner node; the others hold records
    uint64_t inner = 0;
    for(int level = 0; level < 10 && n >= 2; ++level)
    {
        int shift = 3 * (10 - level);
        uint64_t run = 1;
        for(uint64_t p = 1; p <= n; ++p)
        {
            bool same = p < n && (level == 0 || (sorted_key[p] >> shift) == (sorted_key[p - 1] >> shift));
            if(same) { ++run; continue; }
            if(run >= 2) ++inner;
            run = 1;
        }
    }
    std::vector<uint64_t>().swap(sorted_key);
    std::vector<uint32_t>().swap(sorted_who);
    // 3. breadth-first node order = (depth, prefix); push-buffer order inside a node = insertion order
    for(uint64_t i = 0; i < n; ++i)
    {
        uint32_t d = depth[i];
        key[i] = ((uint64_t)d << 30) | (d == 0 ? 0u : (key[i] >> (3 * (10 - d))));
    }
    radix_sort_pairs(key, who, 34);
    uint64_t holding = 0;
    for(uint64_t p = 0; p < n; ++p) if(p == 0 || key[p] != key[p - 1]) ++holding;

    // 4. records, in rank order (the order collect_records emits them)
    prims->clear();
    prims->reserve(n);
    for(uint64_t rank = 0; rank < n; ++rank)
    {
        uint64_t i = who[rank];
        HostPrim p; memset(&p, 0, sizeof(p));
        p.rank = (uint32_t)rank;
        if(i < n_tri)
        {
            uint32_t m = (uint32_t)(std::upper_bound(mesh_first.begin(), mesh_first.end(), i) - mesh_first.begin()) - 1u;
            const OrtMesh &mesh = L->meshes[m];
            uint64_t k = i - mesh_first[m];
            uint32_t i0 = mesh.indices[3 * k], i1 = mesh.indices[3 * k + 1], i2 = mesh.indices[3 * k + 2];
            if(i0 >= mesh.vertex_count || i1 >= mesh.vertex_count || i2 >= mesh.vertex_count) { *err = "mesh index out of range"; return ORT_ERR_ARG; }
            p.kind = PRIM_TRIANGLE; p.a = v3(mesh.vertices[i0]); p.b = v3(mesh.vertices[i1]); p.c = v3(mesh.vertices[i2]);
            p.mat = mesh.mat_index;
            out->info.triangle_count++;
        }
        else if(i < first_box)
        {
            const OrtCylinder &s = L->cylinders[i - first_cyl];
            p.kind = PRIM_CYLINDER; p.a = v3(s.base); p.b = v3(s.axis); p.radius = s.r; p.mat = s.mat_index;
            out->info.cylinder_count++;
        }
        else if(i < first_sph)
        {
            const OrtAAB &s = L->boxes[i - first_box];
            p.kind = PRIM_AAB; p.a = v3(s.min); p.b = v3(s.max); p.mat = s.mat_index;
            out->info.box_count++;
        }
        else if(i < first_csg)
        {
            const OrtSphere &s = L->spheres[i - first_sph];
            p.kind = PRIM_SPHERE; p.a = v3(s.center); p.radius = s.r; p.mat = s.mat_index;
            out->info.sphere_count++;
        }
        else { out->info.csg_count++; continue; }       // inert record, ray.cpp:718-767: a rank, nothing to test
        if(p.mat >= world->mat_count) { *err = "record with material index out of range"; return ORT_ERR_ARG; }
        prim_bounds(p);
        if(!prim_is_finite(p)) { *err = "record with non-finite geometry (NaN / Inf)"; return ORT_ERR_ARG; }
        prims->push_back(p);
    }
    out->info.record_count = (uint32_t)n;
    out->info.octree_node_count = (uint32_t)(inner + holding);
    out->info.octree_max_depth = max_depth;
    out->info.root_min[0] = L->root_min.x; out->info.root_min[1] = L->root_min.y; out->info.root_min[2] = L->root_min.z;
    out->info.root_max[0] = L->root_max.x; out->info.root_max[1] = L->root_max.y; out->info.root_max[2] = L->root_max.z;
    return fill_world_tables(world, out, err);
}

// pads the primitive boxes (see bvh.h: conservative culling) and separates the spheres, which get
// a tree of their own, from the rest
static int pad_and_split(std::vector<HostPrim> &prims, const BuildOptions &opt,
                         std::vector<HostPrim> *spheres, std::vector<HostPrim> *shapes, std::vector<HostPrim> *tris, std::string *err,
                         double scene_abs_given = -1.0)
{
    if(opt.max_leaf < 1 || opt.max_leaf > 3) { *err = "max_leaf must be 1..3"; return ORT_ERR_ARG; }
    double scene_abs = scene_abs_given >= 0.0 ? scene_abs_given : 0.0;
    for(size_t i = 0; scene_abs_given < 0.0 && i < prims.size(); ++i)
        for(int k = 0; k < 3; ++k)
        {
            scene_abs = std::max(scene_abs, fabs((double)prims[i].lo[k]));
            scene_abs = std::max(scene_abs, fabs((double)prims[i].hi[k]));
        }
    for(size_t i = 0; i < prims.size(); ++i)
    {
        HostPrim &p = prims[i];
        if(p.kind == PRIM_SPHERE)
        {
            // A ray TANGENT to a sphere (|b^2 - a*c| < 1e-5) is reported by the
            // reference at t = -b/(2a) (ray.cpp:174-183), i.e. HALF WAY to the
            // sphere, outside any static bound.  Spheres therefore live in their
            // own tree that is traversed without clipping to the best hit so far
            // (bvh.h), and their boxes cover every line that can pass the tangent
            // test: radius sqrt(r^2 + 1e-5/a) with a = |dir|^2 >= 1/4.
            double r = fabs((double)p.radius), rr = sqrt(r * r + 4e-5);
            for(int k = 0; k < 3; ++k)
            {
                double c = k == 0 ? p.a.x : (k == 1 ? p.a.y : p.a.z);
                p.lo[k] = (float)(c - rr); p.hi[k] = (float)(c + rr);
            }
        }
        build::pad_box(p.lo, p.hi, (double)opt.pad_rel, (double)opt.pad_scene * scene_abs);
        if(p.kind == PRIM_SPHERE) spheres->push_back(p);
        else if(p.kind == PRIM_TRIANGLE || opt.merge_shapes) tris->push_back(p);
        else shapes->push_back(p);
    }
    std::vector<HostPrim>().swap(prims);
    return ORT_OK;
}

static int finish_flat_scene(FlatScene *out, uint32_t depth, std::string *err)
{
    out->wide_depth = depth;
    // every level pushes at most one pending node group; +2 for the other trees' root entries
    if(out->wide_depth + 2 > ORT_STACK_SIZE)
    {
        *err = "wide BVH deeper than the traversal stack";
        return ORT_ERR_LIMIT;
    }
    // rank -> record index, for kernels that carry (t, rank) only
    out->rank_to_prim.assign(out->info.record_count, 0xFFFFFFFFu);
    for(size_t i = 0; i < out->prims.size(); ++i)
        if(out->prims[i].rank < out->rank_to_prim.size()) out->rank_to_prim[out->prims[i].rank] = (uint32_t)i;
    out->info.bvh_node_count = (uint32_t)out->nodes.size();
    out->info.bvh_node_bytes = (uint32_t)sizeof(WideNode);
    return ORT_OK;
}

int build_wide_bvh(std::vector<HostPrim> &prims, const BuildOptions &opt, FlatScene *out, std::string *err)
{
    out->nodes.clear(); out->prims.clear(); out->cylinders.clear();
    out->wide_depth = 0;
    out->main_root = 0;
    std::vector<HostPrim> spheres, shapes, tris;
    int rc = pad_and_split(prims, opt, &spheres, &shapes, &tris, err);
    if(rc != ORT_OK) return rc;
    out->prims.reserve(spheres.size() + shapes.size() + tris.size());

    // Three trees share the node array, in this order:
    //   [0, main_root)         spheres            -- traversed UNCLIPPED (see above)
    //   [main_root, tri_root)  boxes + cylinders  -- the few analytic shapes; typically the room
    //   [tri_root, ...)        triangles
    // so the kind of a leaf's records follows from the node index, primitive tests of one kind
    // run together, and the shapes (tested first) give an early bound that clips the triangles.
    uint32_t d0 = 0, d1 = 0, d2 = 0;
    if(!spheres.empty())
    {
        rc = emit_tree(spheres, opt, out, &d0, err);
        if(rc != ORT_OK) return rc;
    }
    out->main_root = (uint32_t)out->nodes.size();
    if(!shapes.empty())
    {
        rc = emit_tree(shapes, opt, out, &d1, err);
        if(rc != ORT_OK) return rc;
    }
    out->tri_root = (uint32_t)out->nodes.size();
    rc = emit_tree(tris, opt, out, &d2, err);
    if(rc != ORT_OK) return rc;
    return finish_flat_scene(out, std::max(d0, std::max(d1, d2)), err);
}

// ---- pieces reused by the CUDA execution for the analytic shapes ------------------------------------------
int fill_world_tables_public(const OrtWorld *world, FlatScene *out, std::string *err) { return fill_world_tables(world, out, err); }
int pad_and_split_public(std::vector<HostPrim> &prims, const BuildOptions &opt, double scene_abs,
                         std::vector<HostPrim> *spheres, std::vector<HostPrim> *shapes, std::vector<HostPrim> *tris, std::string *err)
{
    return pad_and_split(prims, opt, spheres, shapes, tris, err, scene_abs);
}
int emit_tree_public(const std::vector<HostPrim> &prims, const BuildOptions &opt, FlatScene *out, uint32_t *depth_out, std::string *err)
{
    return emit_tree(prims, opt, out, depth_out, err);
}
PrimRec make_record_public(const HostPrim &hp, FlatScene *out) { return make_record(hp, out); }

// cylinders, boxes, spheres of the lists as records (rank = index among these, to be replaced), and the
// octant paths of all analytic records, the CSG's last
int collect_analytic(const OrtWorld *world, const OrtShapeLists *L, f3 root_center, f3 root_half,
                     std::vector<HostPrim> *analytic, uint32_t *keys, std::string *err)
{
    analytic->clear();
    uint32_t k = 0;
    for(uint32_t i = 0; i < L->cylinder_count; ++i, ++k)
    {
        const OrtCylinder &s = L->cylinders[i];
        HostPrim p; memset(&p, 0, sizeof(p));
        p.kind = PRIM_CYLINDER; p.a = v3(s.base); p.b = v3(s.axis); p.radius = s.r; p.mat = s.mat_index; p.rank = k;
        keys[k] = build::octant_path(centre_of_cylinder(s), root_center, root_half);
        analytic->push_back(p);
    }
    for(uint32_t i = 0; i < L->box_count; ++i, ++k)
    {
        const OrtAAB &s = L->boxes[i];
        HostPrim p; memset(&p, 0, sizeof(p));
        p.kind = PRIM_AAB; p.a = v3(s.min); p.b = v3(s.max); p.mat = s.mat_index; p.rank = k;
        keys[k] = build::octant_path(box_centre(v3(s.min), v3(s.max)), root_center, root_half);
        analytic->push_back(p);
    }
    for(uint32_t i = 0; i < L->sphere_count; ++i, ++k)
    {
        const OrtSphere &s = L->spheres[i];
        HostPrim p; memset(&p, 0, sizeof(p));
        p.kind = PRIM_SPHERE; p.a = v3(s.center); p.radius = s.r; p.mat = s.mat_index; p.rank = k;
        keys[k] = build::octant_path(v3(s.center), root_center, root_half);
        analytic->push_back(p);
    }
    if(L->csg) keys[k++] = build::octant_path(box_centre(v3(L->csg->aabb_min), v3(L->csg->aabb_max)), root_center, root_half);
    for(size_t i = 0; i < analytic->size(); ++i)
    {
        if((*analytic)[i].mat >= world->mat_count) { *err = "record with material index out of range"; return ORT_ERR_ARG; }
        prim_bounds((*analytic)[i]);
        if(!prim_is_finite((*analytic)[i])) { *err = "record with non-finite geometry (NaN / Inf)"; return ORT_ERR_ARG; }
    }
    return ORT_OK;
}

// ---- the data-parallel builder (bvh_build.h) ---------------------------------------------------------
// Common preparation for its host and CUDA executions: the sphere tree is built here (analytic
// shapes are few); everything else -- triangles, boxes, cylinders: one tree -- is handed over as
// flat arrays in input order.
int prepare_parallel_build(std::vector<HostPrim> &prims, const BuildOptions &opt, FlatScene *out, ParallelBuildInput *in, std::string *err)
{
    out->nodes.clear(); out->prims.clear(); out->cylinders.clear();
    out->wide_depth = 0;
    out->main_root = 0;
    BuildOptions o = opt; o.merge_shapes = true;
    std::vector<HostPrim> spheres, shapes, rest;
    int rc = pad_and_split(prims, o, &spheres, &shapes, &rest, err);
    if(rc != ORT_OK) return rc;
    in->sphere_depth = 0;
    if(!spheres.empty())
    {
        rc = emit_tree(spheres, o, out, &in->sphere_depth, err);
        if(rc != ORT_OK) return rc;
    }
    out->main_root = out->tri_root = (uint32_t)out->nodes.size();
    in->max_leaf = o.max_leaf;
    in->traversal_cost = o.traversal_cost;
    in->node_cost = o.wide_node_cost;
    in->boxes.resize(6 * rest.size());
    in->recs.resize(rest.size());
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for(size_t i = 0; i < rest.size(); ++i)
    {
        for(int k = 0; k < 3; ++k)
        {
            in->boxes[6 * i + k] = rest[i].lo[k]; in->boxes[6 * i + 3 + k] = rest[i].hi[k];
            double c = 0.5 * ((double)rest[i].lo[k] + (double)rest[i].hi[k]);
            lo[k] = std::min(lo[k], c); hi[k] = std::max(hi[k], c);
        }
        in->recs[i] = make_record(rest[i], out);
    }
    for(int k = 0; k < 3; ++k)
    {
        in->scene_lo[k] = rest.empty() ? 0.0 : lo[k];
        double ext = rest.empty() ? 0.0 : hi[k] - lo[k];
        in->scene_scale[k] = ext > 0.0 ? 2097152.0 / ext : 0.0;
    }
    return ORT_OK;
}

// PLOC's weak spot is the top of the tree (it only ever looks at Morton neighbours); once few clusters
// are left, the rest is built top-down with the binned-SAH builder over the cluster boxes -- a few
// thousand items, milliseconds on the host -- and appended to the binary node array.
uint32_t parallel_top_clusters(uint32_t n)
{
    uint32_t k = ORT_PLOC_TOP_CLUSTERS;
    if(const char *e = getenv("ORT_PLOC_TOP")) k = (uint32_t)atoi(e);
    return std::min(k, n / 64u);          // clustering does the bulk of the work whatever the scene size
}

int build_top_tree(const std::vector<uint32_t> &clusters, const BuildOptions &opt, uint32_t max_leaf, float node_cost,
                   build::B2 *nodes, uint32_t *sizes, build::Dp *cost, uint32_t *next_node, uint32_t *root_out, std::string *err)
{
    std::vector<HostPrim> items(clusters.size());
    for(size_t c = 0; c < clusters.size(); ++c)
    {
        memset(&items[c], 0, sizeof(HostPrim));
        for(int k = 0; k < 3; ++k) { items[c].lo[k] = nodes[clusters[c]].lo[k]; items[c].hi[k] = nodes[clusters[c]].hi[k]; }
    }
    BuildOptions o = opt; o.max_leaf = 1;
    Builder tb(items, o);
    tb.build();
    // children are allocated after their parent: walk the nodes backwards
    std::vector<uint32_t> id_of(tb.nodes.size(), 0u);
    for(size_t i = tb.nodes.size(); i-- > 0; )
    {
        const B2Node &nd = tb.nodes[i];
        if(nd.count > 0)
        {
            if(nd.count != 1) { *err = "internal: top tree leaf with several clusters"; return ORT_ERR_LIMIT; }
            id_of[i] = clusters[tb.idx[nd.first]];
        }
        else
        {
            if(nd.left <= i || nd.right <= i) { *err = "internal: top tree node order"; return ORT_ERR_LIMIT; }
            uint32_t id = (*next_node)++;
            build::merge_nodes(id, id_of[nd.left], id_of[nd.right], nodes, sizes, cost, max_leaf, node_cost);
            id_of[i] = id;
        }
    }
    *root_out = id_of[0];
    return ORT_OK;
}

// Host execution of bvh_build.h: plain loops over the same per-element functions the CUDA kernels
// call, in the same order.  Test infrastructure (tests/sim): it pins what the kernels must produce.
int build_wide_bvh_parallel_host(std::vector<HostPrim> &prims, const BuildOptions &opt, uint32_t radius, FlatScene *out, std::string *err)
{
    using namespace build;
    ParallelBuildInput in;
    int rc = prepare_parallel_build(prims, opt, out, &in, err);
    if(rc != ORT_OK) return rc;
    const uint32_t n = (uint32_t)in.recs.size();
    const uint32_t node_offset = (uint32_t)out->nodes.size(), prim_offset = (uint32_t)out->prims.size();
    uint32_t depth = 0;
    if(n == 0)
    {
        WideNode e; memset(&e, 0, sizeof(e));
        e.ex = e.ey = e.ez = 127;
        for(int s = 0; s < 8; ++s) { e.qlo_x[s] = e.qlo_y[s] = e.qlo_z[s] = 255; }
        out->nodes.push_back(e);
        return finish_flat_scene(out, std::max(in.sphere_depth, 1u), err);
    }
    // 1. Morton order
    std::vector<uint64_t> code(n);
    std::vector<uint32_t> order(n);
    for(uint32_t i = 0; i < n; ++i) { code[i] = morton63(&in.boxes[6 * i], &in.boxes[6 * i + 3], in.scene_lo, in.scene_scale); order[i] = i; }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    // 2. PLOC
    std::vector<B2> nodes(2 * (size_t)n);
    std::vector<uint32_t> sizes(2 * (size_t)n, 0u), cluster(n), next_cluster(n), nn(n), fate(n);
    std::vector<Dp> cost(2 * (size_t)n);
    for(uint32_t i = 0; i < n; ++i)
    {
        B2 l; uint32_t src = order[i];
        for(int k = 0; k < 3; ++k) { l.lo[k] = in.boxes[6 * src + k]; l.hi[k] = in.boxes[6 * src + 3 + k]; }
        l.left = B2_LEAF; l.right = src;
        nodes[i] = l; sizes[i] = 1u | B2_LEAF_FLAG; dp_leaf(&cost[i], half_area(l)); cluster[i] = i;
    }
    uint32_t count = n, next_node = n;
    const uint32_t top_k = parallel_top_clusters(n);
    while(count > std::max(1u, top_k))
    {
        for(uint32_t i = 0; i < count; ++i) nn[i] = ploc_nearest(i, count, cluster.data(), nodes.data(), radius);
        for(uint32_t i = 0; i < count; ++i) fate[i] = ploc_fate(i, nn.data());
        uint32_t pos = 0, mid = 0;
        for(uint32_t i = 0; i < count; ++i)
        {
            ploc_apply(i, nn.data(), cluster.data(), fate[i], pos, mid, next_node, nodes.data(), sizes.data(), cost.data(), next_cluster.data(),
                       in.max_leaf, in.node_cost);
            pos += fate[i] != 0u; mid += fate[i] == 2u;
        }
        next_node += mid; count = pos;
        cluster.swap(next_cluster);
    }
    if(count > 1)
    {
        uint32_t top_root = 0;
        std::vector<uint32_t> live(cluster.begin(), cluster.begin() + count);
        rc = build_top_tree(live, opt, in.max_leaf, in.node_cost, nodes.data(), sizes.data(), cost.data(), &next_node, &top_root, err);
        if(rc != ORT_OK) return rc;
        cluster[0] = top_root;
    }
    const uint32_t root = cluster[0];
    // 3. collapse, level by level
    out->nodes.resize(node_offset + 1);
    out->prims.resize(prim_offset + n);
    std::vector<Item> items(1), next_items;
    items[0].b2 = root; items[0].wide = node_offset;
    uint32_t node_count = node_offset + 1, prim_count = prim_offset;
    std::vector<Kids> kids;
    while(!items.empty())
    {
        ++depth;
        kids.resize(items.size());
        uint32_t total_inner = 0, total_prims = 0;
        for(size_t i = 0; i < items.size(); ++i)
        {
            gather_kids(items[i].b2, nodes.data(), sizes.data(), cost.data(), in.max_leaf, &kids[i]);
            total_inner += kids[i].n_inner; total_prims += kids[i].n_prims;
        }
        out->nodes.resize(node_count + total_inner);
        next_items.resize(total_inner);
        uint32_t run_inner = 0, run_prims = 0;
        for(size_t i = 0; i < items.size(); ++i)
        {
            emit_wide(kids[i], items[i].wide, node_count + run_inner, prim_count + run_prims, run_inner,
                      nodes.data(), sizes.data(), in.max_leaf, in.recs.data(), out->nodes.data(), out->prims.data(), next_items.data());
            run_inner += kids[i].n_inner; run_prims += kids[i].n_prims;
        }
        node_count += total_inner; prim_count += total_prims;
        items.swap(next_items);
    }
    if(prim_count != prim_offset + n) { *err = "internal: parallel build lost records"; return ORT_ERR_LIMIT; }
    return finish_flat_scene(out, std::max(in.sphere_depth, depth), err);
}

} // namespace ort
