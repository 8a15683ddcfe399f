// ref_harness.cpp — TEST INFRASTRUCTURE (oracle), never linked into the product.
//
// Links against the UNMODIFIED reference sources compiled in place from /root/reference/Src
// (recipe: oracle/Makefile; outputs only under oracle/_ref/) and exposes them through a small C
// API that tests/ and bench.py's cpu_baseline / --impl reference legs drive with ctypes.
//
// What is called, not restated: RayTracerScene::{AddShape,RayTrace,FindIntersectionWithScene,
// CalculateLightColor}, RMeshShape (OBJ/MTL/PNG load + KdTree build + traversal), RSphere/RPlane/
// RCapsule, every SurfaceMaterial_*, RRay::TestIntersectionWith*, RMath::Barycentric,
// Math::Q_rsqrt, RTexture::Sample, LinearToGamma/MakePixelColor, RayTracerProgram::SetupScene.
//
// What is restated here (the reference bakes 800x800 into these ~25 lines):
//   * the camera-ray generator of ThreadWorker_Render (RayTracerProgram.cpp:133-169) and
//     BufferIndexToCoord (ColorBuffer.h:19-23), with W and H as parameters;
//   * the 10-row task split of UpdateBitmapPixels (RayTracerProgram.cpp:282,320-327), pulled from
//     an atomic counter instead of the mutex queue;
//   * rand(): interposed with the counter RNG of include/rt_rng.h (the .so is linked with
//     -Bsymbolic-functions so the reference objects bind to this definition).
// Private members are reached with `#define private public` (test-only).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <atomic>
#include <chrono>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
#include <unistd.h>
#include <csignal>
#include <assert.h>

#define private public
#define protected public
#include "RayTracerProgram.h"
#include "RayTracerScene.h"
#include "MeshShape.h"
#include "Shapes.h"
#include "SurfaceMaterials.h"
#include "KdTree.h"
#include "Texture.h"
#include "Math.h"
#include "ColorBuffer.h"
#include "Light.h"
#undef private
#undef protected

#include "rt_rng.h"

extern LightData GSceneLights[];                       // RayTracerScene.cpp:14-18
extern "C" const float* ref_unit_vector_table(unsigned int* count);   // ref_math_wrap.cpp

#define NOINSTR __attribute__((no_instrument_function))

// ---------------------------------------------------------------------------------------------
// rand() interposer
// ---------------------------------------------------------------------------------------------
static thread_local uint32_t t_key = 0;
static thread_local uint32_t t_ctr = 0;
static int g_use_libc_rand = 0;

extern "C" NOINSTR int rand(void)
{
    if (g_use_libc_rand)
        return (int)random();              // glibc rand() is random() behind the same lock
    return rt_rand31(t_key, t_ctr++);
}

static NOINSTR inline void rng_begin(uint32_t seed, uint32_t pixel, uint32_t sample)
{
    t_key = rt_rng_key(seed, pixel, sample);
    t_ctr = 0;
}

// ---------------------------------------------------------------------------------------------
// ray counting (only active in the -finstrument-functions twin build of RayTracerScene.cpp)
// ---------------------------------------------------------------------------------------------
static thread_local uint64_t t_rays = 0;
typedef int (*find_fn_t)(const RayTracerScene*, RRay, RayHitResult&);
static void* g_find_addr = nullptr;

extern "C" NOINSTR void __cyg_profile_func_enter(void* fn, void*)
{
    if (fn == g_find_addr)
        t_rays++;
}
extern "C" NOINSTR void __cyg_profile_func_exit(void*, void*) {}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
static RayTracerProgram* g_program = nullptr;
static uint32_t g_table_seed = 0;
static bool g_table_ready = false;

extern "C" int ref_init(void)
{
    if (!g_program)
    {
        g_program = new RayTracerProgram();     // RayTrace dereferences the singleton (RayTracerScene.cpp:34)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wpmf-conversions"
        g_find_addr = (void*)(find_fn_t)(&RayTracerScene::FindIntersectionWithScene);
#pragma GCC diagnostic pop
    }
    return 0;
}

// Fills PseudoRandomUnitVectors with the reference's own InitPseudoRandomUnitVector
// (Math.cpp:24-31) driven by the counter RNG stream (seed, RT_RNG_TABLE_PIXEL, 0).
extern "C" int ref_init_unit_vectors(uint32_t seed)
{
    ref_init();
    if (g_table_ready && g_table_seed == seed)
        return 0;
    int saved = g_use_libc_rand;
    g_use_libc_rand = 0;
    rng_begin(seed, RT_RNG_TABLE_PIXEL, 0);
    RMath::InitPseudoRandomUnitVector();
    g_use_libc_rand = saved;
    g_table_seed = seed;
    g_table_ready = true;
    return 0;
}

extern "C" void ref_set_libc_rand(int on) { g_use_libc_rand = on; }
extern "C" int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

// ---------------------------------------------------------------------------------------------
// scene construction through the reference's public API
// ---------------------------------------------------------------------------------------------
extern "C" void* ref_scene_new(void) { ref_init(); return new RayTracerScene(); }
extern "C" void ref_scene_free(void* s) { delete (RayTracerScene*)s; }

extern "C" void* ref_mat_diffuse(float r, float g, float b) { return new SurfaceMaterial_Diffuse(RVec3(r, g, b)); }
extern "C" void* ref_mat_checker(float r, float g, float b, float size) { return new SurfaceMaterial_DiffuseChecker(RVec3(r, g, b), size); }
extern "C" void* ref_mat_reflective(float r, float g, float b, float fuzz) { return new SurfaceMaterial_Reflective(RVec3(r, g, b), fuzz); }
extern "C" void* ref_mat_emissive(float r, float g, float b) { return new SurfaceMaterial_Emissive(RVec3(r, g, b)); }
extern "C" void* ref_mat_null(void) { return new SurfaceMaterial_Null(); }
extern "C" void* ref_mat_blend(void* a, void* b, float f)
{
    return new SurfaceMaterial_Blend(std::unique_ptr<ISurfaceMaterial>((ISurfaceMaterial*)a),
                                     std::unique_ptr<ISurfaceMaterial>((ISurfaceMaterial*)b), f);
}
extern "C" void* ref_mat_combine(void* a, void* b)
{
    return new SurfaceMaterial_Combine(std::unique_ptr<ISurfaceMaterial>((ISurfaceMaterial*)a),
                                       std::unique_ptr<ISurfaceMaterial>((ISurfaceMaterial*)b));
}

static int add_shape(void* s, std::unique_ptr<RShape> shape, void* mat)
{
    RayTracerScene* scene = (RayTracerScene*)s;
    scene->AddShape(std::move(shape), std::unique_ptr<ISurfaceMaterial>((ISurfaceMaterial*)mat));
    return (int)scene->SceneShapes.size() - 1;
}

extern "C" int ref_add_sphere(void* s, float x, float y, float z, float r, void* mat)
{
    return add_shape(s, RSphere::Create(RVec3(x, y, z), r), mat);
}
extern "C" int ref_add_plane(void* s, float nx, float ny, float nz, float px, float py, float pz, void* mat)
{
    return add_shape(s, RPlane::Create(RVec3(nx, ny, nz), RVec3(px, py, pz)), mat);
}
extern "C" int ref_add_capsule(void* s, float ax, float ay, float az, float bx, float by, float bz, float r, void* mat)
{
    return add_shape(s, RCapsule::Create(RVec3(ax, ay, az), RVec3(bx, by, bz), r), mat);
}
extern "C" int ref_add_triangle(void* s, const float* p, void* mat)
{
    return add_shape(s, RTriangle::Create(RVec3(p), RVec3(p + 3), RVec3(p + 6)), mat);
}
extern "C" int ref_add_mesh(void* s, const char* obj_path, void* mat)
{
    return add_shape(s, RMeshShape::Create(obj_path), mat);
}

// The reference's hard-coded scene (RayTracerProgram.cpp:467-552).  SetupScene opens
// "Data/unitychan.obj" relative to the cwd, so run it from `data_parent` (a directory holding Data/).
extern "C" void* ref_default_scene(const char* data_parent)
{
    ref_init();
    if (g_program->Scene.SceneShapes.empty())
    {
        char cwd[4096];
        if (!getcwd(cwd, sizeof cwd)) return nullptr;
        if (chdir(data_parent) != 0) return nullptr;
        g_program->SetupScene();
        if (chdir(cwd) != 0) return nullptr;
    }
    return &g_program->Scene;
}

extern "C" int ref_num_shapes(void* s) { return (int)((RayTracerScene*)s)->SceneShapes.size(); }

// out[0..5] = bounds min/max, returns HasCullingBounds
extern "C" int ref_shape_bounds(void* s, int i, float* out)
{
    RShape* sh = ((RayTracerScene*)s)->SceneShapes[i].get();
    const RAabb& b = sh->GetBounds();
    out[0] = b.pMin.x; out[1] = b.pMin.y; out[2] = b.pMin.z;
    out[3] = b.pMax.x; out[4] = b.pMax.y; out[5] = b.pMax.z;
    return sh->HasCullingBounds() ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// mesh internals (to pin the product's host-side OBJ/MTL/PNG loader and BVH builder)
// ---------------------------------------------------------------------------------------------
static RMeshShape* mesh_of(void* s, int i)
{
    return dynamic_cast<RMeshShape*>(((RayTracerScene*)s)->SceneShapes[i].get());
}

static int count_nodes(const KdNode* n) { return n ? 1 + count_nodes(n->Left.get()) + count_nodes(n->Right.get()) : 0; }
static int depth_of(const KdNode* n) { return n ? 1 + std::max(depth_of(n->Left.get()), depth_of(n->Right.get())) : 0; }

// out = {points, texcoords, normals, triangles, Textures.size(), bvh nodes, bvh depth}
extern "C" int ref_mesh_counts(void* s, int shape, int* out)
{
    RMeshShape* m = mesh_of(s, shape);
    if (!m) return -1;
    out[0] = (int)m->Points.size();
    out[1] = (int)m->Texcoords.size();
    out[2] = (int)m->Normals.size();
    out[3] = (int)m->PointIndices.size() / 3;
    out[4] = (int)m->Textures.size();
    const KdNode* root = m->Spatial ? m->Spatial->RootNode.get() : nullptr;
    out[5] = count_nodes(root);
    out[6] = depth_of(root);
    return 0;
}

extern "C" int ref_mesh_dump(void* s, int shape, float* points, float* texcoords, float* normals,
                             int* pidx, int* tidx, int* nidx, int* matid)
{
    RMeshShape* m = mesh_of(s, shape);
    if (!m) return -1;
    memcpy(points, m->Points.data(), m->Points.size() * sizeof(RVec3));
    memcpy(texcoords, m->Texcoords.data(), m->Texcoords.size() * sizeof(RVec3));
    memcpy(normals, m->Normals.data(), m->Normals.size() * sizeof(RVec3));
    memcpy(pidx, m->PointIndices.data(), m->PointIndices.size() * sizeof(int));
    memcpy(tidx, m->TexcoordIndices.data(), m->TexcoordIndices.size() * sizeof(int));
    memcpy(nidx, m->NormalIndices.data(), m->NormalIndices.size() * sizeof(int));
    memcpy(matid, m->PolyMaterialId.data(), m->PolyMaterialId.size() * sizeof(int));
    return 0;
}

// wh = {width, height}; {0,0} for a null entry
extern "C" int ref_mesh_texture_info(void* s, int shape, int tex, int* wh)
{
    RMeshShape* m = mesh_of(s, shape);
    if (!m || tex < 0 || tex >= (int)m->Textures.size()) return -1;
    RTexture* t = m->Textures[tex].get();
    wh[0] = t ? t->Width : 0;
    wh[1] = t ? t->Height : 0;
    return 0;
}

extern "C" int ref_mesh_texture_pixels(void* s, int shape, int tex, float* out)
{
    RMeshShape* m = mesh_of(s, shape);
    if (!m || tex < 0 || tex >= (int)m->Textures.size() || !m->Textures[tex]) return -1;
    RTexture* t = m->Textures[tex].get();
    memcpy(out, t->Pixels.data(), t->Pixels.size() * sizeof(RVec4));
    return 0;
}

struct FlatOut { float* bounds; int* escape; int* tri; int* verts; int next; };

static void flatten(const KdNode* n, FlatOut& o)
{
    int i = o.next++;
    o.bounds[6 * i + 0] = n->Bounds.pMin.x; o.bounds[6 * i + 1] = n->Bounds.pMin.y; o.bounds[6 * i + 2] = n->Bounds.pMin.z;
    o.bounds[6 * i + 3] = n->Bounds.pMax.x; o.bounds[6 * i + 4] = n->Bounds.pMax.y; o.bounds[6 * i + 5] = n->Bounds.pMax.z;
    bool leaf = !n->Left && !n->Right;
    o.tri[i] = leaf ? n->Triangle.Index : -1;
    o.verts[3 * i + 0] = leaf ? n->Triangle.p0 : -1;
    o.verts[3 * i + 1] = leaf ? n->Triangle.p1 : -1;
    o.verts[3 * i + 2] = leaf ? n->Triangle.p2 : -1;
    if (n->Left) flatten(n->Left.get(), o);
    if (n->Right) flatten(n->Right.get(), o);
    o.escape[i] = o.next;
}

// Pre-order dump of the reference's KdNode tree: per node 6 bounds floats, the escape index,
// the leaf's TriangleData::Index (-1 for inner nodes) and its three point indices.
extern "C" int ref_mesh_bvh_dump(void* s, int shape, float* bounds, int* escape, int* tri, int* verts)
{
    RMeshShape* m = mesh_of(s, shape);
    if (!m || !m->Spatial || !m->Spatial->RootNode) return -1;
    FlatOut o = { bounds, escape, tri, verts, 0 };
    flatten(m->Spatial->RootNode.get(), o);
    return o.next;
}

// ---------------------------------------------------------------------------------------------
// camera (restated with parametric W x H) — RayTracerProgram.cpp:133-165, ColorBuffer.h:19-23
// ---------------------------------------------------------------------------------------------
struct Camera
{
    int W, H;
    RVec3 ViewPoint;
    float DirZ, RayDistance, Aspect;
};

static NOINSTR Camera make_camera(int W, int H)
{
    Camera c;
    c.W = W; c.H = H;
    c.ViewPoint = RVec3(0, 0, 7.0f);
    c.DirZ = -0.5f;
    c.RayDistance = 1000.0f;
    c.Aspect = (float)W / (float)H;
    return c;
}

static NOINSTR inline void pixel_base_dir(const Camera& c, int PixelIndex, float& dx, float& dy)
{
    int x = PixelIndex % c.W;
    int y = PixelIndex / c.W;
    dx = -(float)(x - c.W / 2) / (c.W * 2) * c.Aspect;
    dy = -(float)(y - c.H / 2) / (c.H * 2);
}

static NOINSTR inline RRay centre_ray(const Camera& c, int PixelIndex)
{
    float dx, dy;
    pixel_base_dir(c, PixelIndex, dx, dy);
    RVec3 Dir(dx, dy, c.DirZ);
    return RRay(c.ViewPoint, Dir.GetNormalizedVec3(), c.RayDistance);
}

// sub-sample i of the ENABLE_ANTIALIASING branch; consumes two Random() draws
static NOINSTR inline RRay jittered_ray(const Camera& c, int PixelIndex, int i)
{
    float dx, dy;
    pixel_base_dir(c, PixelIndex, dx, dy);
    const float inv_pixel_radius = 1.0f / (c.W * 4);
    const float ox[4] = { 0.0f, inv_pixel_radius, 0.0f, inv_pixel_radius };
    const float oy[4] = { 0.0f, 0.0f, inv_pixel_radius, inv_pixel_radius };
    const float offset_radius = inv_pixel_radius * 0.5f;
    float offset_x = ox[i];
    float offset_y = oy[i];
    offset_x += (RMath::Random() - 0.5f) * offset_radius;
    offset_y += (RMath::Random() - 0.5f) * offset_radius;
    RVec3 Dir(dx + offset_x, dy + offset_y, c.DirZ);
    return RRay(c.ViewPoint, Dir.GetNormalizedVec3(), c.RayDistance);
}

// ---------------------------------------------------------------------------------------------
// primary-hit dump (T0 parity): shape index, triangle id, Distance
// ---------------------------------------------------------------------------------------------
struct TraverseCount { uint64_t nodes, tris; };

// Counting twin of KdNode::TestRayIntersection (KdTree.cpp:128-195); the result is cross-checked
// against the real call below.
static bool counting_traverse(const KdNode* n, RRay& TestRay, const RVec3 Points[], int* tri, TraverseCount& tc)
{
    tc.nodes++;
    if (!TestRay.TestIntersectionWithAabb(n->Bounds)) return false;
    bool leaf = true, res = false;
    if (n->Left) { res |= counting_traverse(n->Left.get(), TestRay, Points, tri, tc); leaf = false; }
    if (n->Right) { res |= counting_traverse(n->Right.get(), TestRay, Points, tri, tc); leaf = false; }
    if (leaf)
    {
        const RVec3 TriPoints[] = { Points[n->Triangle.p0], Points[n->Triangle.p1], Points[n->Triangle.p2] };
        RayHitResult hr;
        tc.tris++;
        if (TestRay.TestIntersectionWithTriangle(TriPoints, &hr))
        {
            TestRay.Distance = hr.Distance;
            *tri = n->Triangle.Index;
            return true;
        }
        return false;
    }
    return res;
}

// Restates the shape loop of FindIntersectionWithScene (RayTracerScene.cpp:99-125) only to learn
// WHICH triangle each mesh reported; returns the shape index and checks it, with Distance, against
// the real FindIntersectionWithScene.  counters: [0]=slab tests, [1]=triangle tests, [2]=mismatches.
static int find_with_ids(const RayTracerScene* scene, const RRay& ray, RayHitResult& out, int* tri_out, uint64_t* counters)
{
    RRay TestRay = ray;
    int HitShape = -1, HitTri = -1, Index = 0;
    for (auto& Shape : scene->SceneShapes)
    {
        bool enter = !Shape->HasCullingBounds();
        if (!enter) { counters[0]++; enter = TestRay.TestIntersectionWithAabb(Shape->GetBounds()); }
        if (enter)
        {
            int tri = -1;
            RMeshShape* mesh = dynamic_cast<RMeshShape*>(Shape.get());
            if (mesh && mesh->Spatial && mesh->Spatial->RootNode)
            {
                int tri_real = -1, tri_cnt = -1;
                mesh->Spatial->TestRayIntersection(TestRay, mesh->Points.data(), nullptr, &tri_real);
                RRay Copy = TestRay;
                TraverseCount tc = { 0, 0 };
                counting_traverse(mesh->Spatial->RootNode.get(), Copy, mesh->Points.data(), &tri_cnt, tc);
                counters[0] += tc.nodes; counters[1] += tc.tris;
                if (tri_cnt != tri_real) counters[2]++;
                tri = tri_real;
            }
            bool hit = Shape->TestRayIntersection(TestRay, &out);
            if (hit) { TestRay.Distance = out.Distance; HitShape = Index; HitTri = tri; }
        }
        Index++;
    }
    *tri_out = HitTri;
    return HitShape;
}

// hit: 11 floats per pixel = HitPosition, HitNormal, Distance, SampledColor, SampledAlpha (may be NULL)
extern "C" int ref_trace_primary(void* s, int W, int H, int start, int end, int* shape, int* tri,
                                 float* dist, float* hit, uint64_t* counters)
{
    const RayTracerScene* scene = (RayTracerScene*)s;
    Camera cam = make_camera(W, H);
    uint64_t cnt[3] = { 0, 0, 0 };
    for (int p = start; p <= end; p++)
    {
        RRay ray = centre_ray(cam, p);
        RayHitResult r, r2;
        int t = -1;
        int sh = find_with_ids(scene, ray, r, &t, cnt);
        int sh2 = scene->FindIntersectionWithScene(ray, r2);
        if (sh != sh2 || (sh >= 0 && memcmp(&r.Distance, &r2.Distance, 4) != 0)) cnt[2]++;
        int k = p - start;
        shape[k] = sh2;
        tri[k] = sh2 >= 0 ? t : -1;
        dist[k] = sh2 >= 0 ? r2.Distance : 0.0f;
        if (hit)
        {
            float* h = hit + 11 * (size_t)k;
            if (sh2 >= 0)
            {
                h[0] = r2.HitPosition.x; h[1] = r2.HitPosition.y; h[2] = r2.HitPosition.z;
                h[3] = r2.HitNormal.x; h[4] = r2.HitNormal.y; h[5] = r2.HitNormal.z;
                h[6] = r2.Distance;
                h[7] = r2.SampledColor.x; h[8] = r2.SampledColor.y; h[9] = r2.SampledColor.z;
                h[10] = r2.SampledAlpha;
            }
            else memset(h, 0, 11 * sizeof(float));
        }
    }
    if (counters) { counters[0] = cnt[0]; counters[1] = cnt[1]; counters[2] = cnt[2]; }
    return 0;
}

// Arbitrary rays (origin, direction, distance = 7 floats each) through FindIntersectionWithScene.
extern "C" int ref_trace_rays(void* s, const float* rays, int n, int* shape, int* tri, float* hit11)
{
    const RayTracerScene* scene = (RayTracerScene*)s;
    uint64_t cnt[3] = { 0, 0, 0 };
    for (int k = 0; k < n; k++)
    {
        const float* q = rays + 7 * (size_t)k;
        RRay ray(RVec3(q[0], q[1], q[2]), RVec3(q[3], q[4], q[5]), q[6]);
        RayHitResult r, r2;
        int t = -1;
        find_with_ids(scene, ray, r, &t, cnt);
        int sh = scene->FindIntersectionWithScene(ray, r2);
        shape[k] = sh;
        tri[k] = sh >= 0 ? t : -1;
        float* h = hit11 + 11 * (size_t)k;
        memset(h, 0, 11 * sizeof(float));
        if (sh >= 0)
        {
            h[0] = r2.HitPosition.x; h[1] = r2.HitPosition.y; h[2] = r2.HitPosition.z;
            h[3] = r2.HitNormal.x; h[4] = r2.HitNormal.y; h[5] = r2.HitNormal.z;
            h[6] = r2.Distance;
            h[7] = r2.SampledColor.x; h[8] = r2.SampledColor.y; h[9] = r2.SampledColor.z;
            h[10] = r2.SampledAlpha;
        }
    }
    return (int)cnt[2];
}

// ---------------------------------------------------------------------------------------------
// rendering: the per-pixel body of ThreadWorker_Render with parametric W x H
// ---------------------------------------------------------------------------------------------
enum { MODE_PATH = 0, MODE_PREVIEW = 1, MODE_WHITTED = 2 };

struct RenderJob
{
    const RayTracerScene* scene;
    Camera cam;
    int start, end, mode, max_bounce, pass_begin, pass_count, antialias;
    uint32_t seed;
    float* accum;          // W*H*4: rgb sum + Num, indexed by absolute pixel
    uint32_t* display;     // W*H ARGB or NULL
    std::atomic<int> next_row;
    std::atomic<uint64_t> rays, shadow;
};

// Whitted config: nearest hit, then the reference's (otherwise dead) CalculateLightColor per
// light with the texture colour as surface colour; sky on miss via RayTrace's own miss branch.
static RVec3 whitted(const RayTracerScene* scene, const RRay& ray, uint64_t& shadow)
{
    RayHitResult r;
    int sh = scene->FindIntersectionWithScene(ray, r);
    if (sh == -1)
    {
        // RayTrace would repeat the same (missing) query; call it for the sky colour only and
        // do not let the twin build count that query twice.
        uint64_t saved = t_rays;
        RVec3 sky = scene->RayTrace(ray, 1, RenderOption());
        t_rays = saved;
        return sky;
    }
    RVec3 c = RVec3::Zero();
    const int n_lights = 1;   // GSceneLights has one entry (RayTracerScene.cpp:14-18)
    for (int i = 0; i < n_lights; i++)
    {
        c += scene->CalculateLightColor(&GSceneLights[i], r, r.SampledColor);
        shadow++;
    }
    return c;
}

static void render_pixel(RenderJob& job, int PixelIndex, uint64_t& shadow)
{
    RenderOption opt;
    opt.UseBaseColor = (job.mode == MODE_PREVIEW);
    float* a = job.accum + 4 * (size_t)PixelIndex;
    RVec3 sum(a[0], a[1], a[2]);
    int num = (int)a[3];
    RVec3 c_last = RVec3::Zero();
    for (int pass = job.pass_begin; pass < job.pass_begin + job.pass_count; pass++)
    {
        RVec3 c = RVec3::Zero();
        if (job.antialias)
        {
            for (int i = 0; i < 4; i++)
            {
                rng_begin(job.seed, (uint32_t)PixelIndex, (uint32_t)(pass * 4 + i));
                RRay ray = jittered_ray(job.cam, PixelIndex, i);
                c += (job.mode == MODE_WHITTED) ? whitted(job.scene, ray, shadow)
                                                : job.scene->RayTrace(ray, job.max_bounce, opt);
            }
            c /= 4.0f;
        }
        else
        {
            rng_begin(job.seed, (uint32_t)PixelIndex, (uint32_t)pass);
            RRay ray = centre_ray(job.cam, PixelIndex);
            c = (job.mode == MODE_WHITTED) ? whitted(job.scene, ray, shadow)
                                           : job.scene->RayTrace(ray, job.max_bounce, opt);
        }
        sum += c;      // AccumulatePixel::AddPixel, RayTracerProgram.cpp:57-61
        num++;
        c_last = c;
    }
    a[0] = sum.x; a[1] = sum.y; a[2] = sum.z; a[3] = (float)num;
    if (job.display)
    {
        // preview: MakePixelColor(LinearToGamma(c)) (:179); else GetGammaSpacePixel (:68-71,:185)
        RVec3 lin = (job.mode == MODE_PREVIEW) ? c_last : sum / (float)num;
        job.display[PixelIndex] = MakePixelColor(LinearToGamma(lin));
    }
}

static void render_worker(RenderJob* job)
{
    const int NumTaskRows = 10;      // RayTracerProgram.cpp:282
    const int W = job->cam.W;
    uint64_t shadow = 0;
    t_rays = 0;
    int first_row = job->start / W, last_row = job->end / W;
    for (;;)
    {
        int row = job->next_row.fetch_add(NumTaskRows);
        if (row > last_row) break;
        int s = std::max(job->start, row * W);
        int e = std::min(job->end, (row + NumTaskRows) * W - 1);
        for (int p = s; p <= e; p++) render_pixel(*job, p, shadow);
    }
    (void)first_row;
    job->rays += t_rays;
    job->shadow += shadow;
}

// Renders pixels [start,end] for `pass_count` passes on `nthreads` threads, adding into accum.
// out_stats: [0] = seconds, [1] = FindIntersectionWithScene calls (twin build only, else 0),
// [2] = shadow queries.
extern "C" int ref_render(void* s, int W, int H, int start, int end, int mode, int max_bounce,
                          int pass_begin, int pass_count, int antialias, uint32_t seed, int nthreads,
                          float* accum, uint32_t* display, double* out_stats)
{
    RenderJob job;
    job.scene = (RayTracerScene*)s;
    job.cam = make_camera(W, H);
    job.start = start; job.end = end; job.mode = mode; job.max_bounce = max_bounce;
    job.pass_begin = pass_begin; job.pass_count = pass_count; job.antialias = antialias;
    job.seed = seed; job.accum = accum; job.display = display;
    job.next_row = (start / W);
    job.rays = 0; job.shadow = 0;
    if (nthreads < 1) nthreads = 1;
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) render_worker(&job);
    else
    {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; i++) th.emplace_back(render_worker, &job);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (out_stats)
    {
        out_stats[0] = std::chrono::duration<double>(t1 - t0).count();
        out_stats[1] = (double)job.rays.load();
        out_stats[2] = (double)job.shadow.load();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// primitive known-answer entry points (tier-1 parity tests)
// ---------------------------------------------------------------------------------------------
// rays: 7 floats (origin, dir, distance); boxes: 6 floats; out: 1 = accepted
extern "C" void ref_kat_aabb(const float* rays, const float* boxes, int n, int* out, float* tmin)
{
    for (int i = 0; i < n; i++)
    {
        const float* q = rays + 7 * (size_t)i; const float* b = boxes + 6 * (size_t)i;
        RRay ray(RVec3(q[0], q[1], q[2]), RVec3(q[3], q[4], q[5]), q[6]);
        RAabb box; box.pMin = RVec3(b[0], b[1], b[2]); box.pMax = RVec3(b[3], b[4], b[5]);
        float t = 0.0f;
        out[i] = ray.TestIntersectionWithAabb(box, &t) ? 1 : 0;
        tmin[i] = out[i] ? t : 0.0f;
    }
}

// tris: 9 floats; out7: HitPosition, HitNormal, Distance
extern "C" void ref_kat_triangle(const float* rays, const float* tris, int n, int* out, float* out7)
{
    for (int i = 0; i < n; i++)
    {
        const float* q = rays + 7 * (size_t)i; const float* t = tris + 9 * (size_t)i;
        RRay ray(RVec3(q[0], q[1], q[2]), RVec3(q[3], q[4], q[5]), q[6]);
        RVec3 P[3] = { RVec3(t), RVec3(t + 3), RVec3(t + 6) };
        RayHitResult r;
        out[i] = ray.TestIntersectionWithTriangle(P, &r) ? 1 : 0;
        float* o = out7 + 7 * (size_t)i;
        memset(o, 0, 7 * sizeof(float));
        if (out[i])
        {
            o[0] = r.HitPosition.x; o[1] = r.HitPosition.y; o[2] = r.HitPosition.z;
            o[3] = r.HitNormal.x; o[4] = r.HitNormal.y; o[5] = r.HitNormal.z; o[6] = r.Distance;
        }
    }
}

extern "C" void ref_kat_sphere(const float* rays, const float* spheres, int n, int* out, float* out7)
{
    for (int i = 0; i < n; i++)
    {
        const float* q = rays + 7 * (size_t)i; const float* sp = spheres + 4 * (size_t)i;
        RRay ray(RVec3(q[0], q[1], q[2]), RVec3(q[3], q[4], q[5]), q[6]);
        RayHitResult r;
        out[i] = ray.TestIntersectionWithSphere(RVec3(sp), sp[3], &r) ? 1 : 0;
        float* o = out7 + 7 * (size_t)i;
        memset(o, 0, 7 * sizeof(float));
        if (out[i])
        {
            o[0] = r.HitPosition.x; o[1] = r.HitPosition.y; o[2] = r.HitPosition.z;
            o[3] = r.HitNormal.x; o[4] = r.HitNormal.y; o[5] = r.HitNormal.z; o[6] = r.Distance;
        }
    }
}

extern "C" void ref_kat_plane(const float* rays, const float* planes, int n, int* out, float* out7)
{
    for (int i = 0; i < n; i++)
    {
        const float* q = rays + 7 * (size_t)i; const float* pl = planes + 6 * (size_t)i;
        RRay ray(RVec3(q[0], q[1], q[2]), RVec3(q[3], q[4], q[5]), q[6]);
        RayHitResult r;
        out[i] = ray.TestIntersectionWithPlane(RVec3(pl), RVec3(pl + 3), &r) ? 1 : 0;
        float* o = out7 + 7 * (size_t)i;
        memset(o, 0, 7 * sizeof(float));
        if (out[i])
        {
            o[0] = r.HitPosition.x; o[1] = r.HitPosition.y; o[2] = r.HitPosition.z;
            o[3] = r.HitNormal.x; o[4] = r.HitNormal.y; o[5] = r.HitNormal.z; o[6] = r.Distance;
        }
    }
}

// capsules: start(3), end(3), radius
extern "C" void ref_kat_capsule(const float* rays, const float* caps, int n, int* out, float* out7)
{
    for (int i = 0; i < n; i++)
    {
        const float* q = rays + 7 * (size_t)i; const float* cp = caps + 7 * (size_t)i;
        RRay ray(RVec3(q[0], q[1], q[2]), RVec3(q[3], q[4], q[5]), q[6]);
        RCapsule cap(RVec3(cp), RVec3(cp + 3), cp[6]);
        RayHitResult r;
        out[i] = cap.TestRayIntersection(ray, &r) ? 1 : 0;
        float* o = out7 + 7 * (size_t)i;
        memset(o, 0, 7 * sizeof(float));
        if (out[i])
        {
            o[0] = r.HitPosition.x; o[1] = r.HitPosition.y; o[2] = r.HitPosition.z;
            o[3] = r.HitNormal.x; o[4] = r.HitNormal.y; o[5] = r.HitNormal.z; o[6] = r.Distance;
        }
    }
}

extern "C" void ref_kat_qrsqrt(const float* x, int n, float* out)
{
    for (int i = 0; i < n; i++) out[i] = Math::Q_rsqrt(x[i]);
}

// in: p,a,b,c (12 floats) ; out: u,v,w
extern "C" void ref_kat_barycentric(const float* in, int n, float* out)
{
    for (int i = 0; i < n; i++)
    {
        const float* q = in + 12 * (size_t)i;
        RMath::Barycentric(RVec3(q), RVec3(q + 3), RVec3(q + 6), RVec3(q + 9), out[3 * i], out[3 * i + 1], out[3 * i + 2]);
    }
}

extern "C" int ref_kat_texture_sample(void* s, int shape, int tex, const float* uv, int n, float* out4)
{
    RMeshShape* m = mesh_of(s, shape);
    if (!m || tex < 0 || tex >= (int)m->Textures.size() || !m->Textures[tex]) return -1;
    for (int i = 0; i < n; i++)
    {
        RVec4 c = m->Textures[tex]->Sample(uv[2 * i], uv[2 * i + 1]);
        out4[4 * i] = c.x; out4[4 * i + 1] = c.y; out4[4 * i + 2] = c.z; out4[4 * i + 3] = c.w;
    }
    return 0;
}

// linear rgb (3 floats) -> ARGB via MakePixelColor(LinearToGamma(.)) (ColorBuffer.h:81-109)
extern "C" void ref_kat_display(const float* rgb, int n, uint32_t* out)
{
    for (int i = 0; i < n; i++) out[i] = MakePixelColor(LinearToGamma(RVec3(rgb + 3 * (size_t)i)));
}

extern "C" void ref_light0(float* out7)
{
    out7[0] = (float)GSceneLights[0].Type;
    out7[1] = GSceneLights[0].PositionOrDirection.x; out7[2] = GSceneLights[0].PositionOrDirection.y; out7[3] = GSceneLights[0].PositionOrDirection.z;
    out7[4] = GSceneLights[0].Color.x; out7[5] = GSceneLights[0].Color.y; out7[6] = GSceneLights[0].Color.z;
}
