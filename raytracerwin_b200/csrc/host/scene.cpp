// scene.cpp — host scene objects and their flattening into the POD rt_scene_desc that
// rt_gpu_upload_scene consumes.  Mirrors the construction-time behaviour of the reference's
// RayTracerScene (RayTracerScene.cpp:20-29), shapes (Shapes.h:46-123) and materials
// (SurfaceMaterials.cpp:15-18,40-51,92-96,127-130,145-151,163-167).
#include "rt_host.hpp"
#include "rt_rng.h"

#include <float.h>
#include <math.h>
#include <string.h>

#include <thread>

namespace rtb200 {

// ---- RAabb -----------------------------------------------------------------------------------
RAabb::RAabb() : pMin(FLT_MAX, FLT_MAX, FLT_MAX), pMax(-FLT_MAX, -FLT_MAX, -FLT_MAX) {}

void RAabb::Expand(const RVec3& p)
{
    if (p.x < pMin.x) pMin.x = p.x;
    if (p.y < pMin.y) pMin.y = p.y;
    if (p.z < pMin.z) pMin.z = p.z;
    if (p.x > pMax.x) pMax.x = p.x;
    if (p.y > pMax.y) pMax.y = p.y;
    if (p.z > pMax.z) pMax.z = p.z;
}

void RAabb::ExpandBySphere(const RVec3& c, float r)
{
    if (c.x - r < pMin.x) pMin.x = c.x - r;
    if (c.y - r < pMin.y) pMin.y = c.y - r;
    if (c.z - r < pMin.z) pMin.z = c.z - r;
    if (c.x + r > pMax.x) pMax.x = c.x + r;
    if (c.y + r > pMax.y) pMax.y = c.y + r;
    if (c.z + r > pMax.z) pMax.z = c.z + r;
}

// ---- materials -------------------------------------------------------------------------------
static int push_material(std::vector<rt_material>& out, int type, const RVec3& rgb, float scalar, int a, int b)
{
    rt_material m;
    m.type = type; m.child_a = a; m.child_b = b;
    m.rgb[0] = rgb.x; m.rgb[1] = rgb.y; m.rgb[2] = rgb.z;
    m.scalar = scalar;
    out.push_back(m);
    return (int)out.size() - 1;
}

int SurfaceMaterial_Diffuse::Flatten(std::vector<rt_material>& out) const
{
    return push_material(out, RT_MAT_DIFFUSE, Albedo, 0.0f, -1, -1);
}

SurfaceMaterial_DiffuseChecker::SurfaceMaterial_DiffuseChecker(const RVec3& InAlbedo, float InPatternSize)
    : SurfaceMaterial_Diffuse(InAlbedo)
{
    // a zero pattern size falls back to 1 (SurfaceMaterials.cpp:43-50)
    ReciprocalPatternSize = fabsf(InPatternSize) < FLT_EPSILON ? 1.0f : 1.0f / InPatternSize;
}

int SurfaceMaterial_DiffuseChecker::Flatten(std::vector<rt_material>& out) const
{
    return push_material(out, RT_MAT_CHECKER, Albedo, ReciprocalPatternSize, -1, -1);
}

int SurfaceMaterial_Reflective::Flatten(std::vector<rt_material>& out) const
{
    return push_material(out, RT_MAT_REFLECTIVE, Albedo, Fuzziness, -1, -1);
}

int SurfaceMaterial_Emissive::Flatten(std::vector<rt_material>& out) const
{
    return push_material(out, RT_MAT_EMISSIVE, Color, 0.0f, -1, -1);
}

SurfaceMaterial_Blend::SurfaceMaterial_Blend(std::unique_ptr<ISurfaceMaterial> InMaterialA, std::unique_ptr<ISurfaceMaterial> InMaterialB, float InBlendFactor)
    : BlendMaterialA(std::move(InMaterialA)), BlendMaterialB(std::move(InMaterialB))
{
    // clamped to [0,1] at construction (SurfaceMaterials.cpp:148)
    BlendFactor = InBlendFactor < 0.0f ? 0.0f : (InBlendFactor > 1.0f ? 1.0f : InBlendFactor);
}

int SurfaceMaterial_Blend::Flatten(std::vector<rt_material>& out) const
{
    int a = BlendMaterialA ? BlendMaterialA->Flatten(out) : -1;
    int b = BlendMaterialB ? BlendMaterialB->Flatten(out) : -1;
    return push_material(out, RT_MAT_BLEND, RVec3(0, 0, 0), BlendFactor, a, b);
}

int SurfaceMaterial_Combine::Flatten(std::vector<rt_material>& out) const
{
    int a = MaterialA ? MaterialA->Flatten(out) : -1;
    int b = MaterialB ? MaterialB->Flatten(out) : -1;
    return push_material(out, RT_MAT_COMBINE, RVec3(0, 0, 0), 0.0f, a, b);
}

int SurfaceMaterial_Null::Flatten(std::vector<rt_material>& out) const
{
    return push_material(out, RT_MAT_NULL, RVec3(0, 0, 0), 0.0f, -1, -1);
}

// ---- shapes ----------------------------------------------------------------------------------
void RShape::FlattenCommon(rt_shape& out, int type) const
{
    memset(&out, 0, sizeof out);
    out.type = type;
    out.material = -1;
    out.mesh = -1;
    out.has_bounds = HasCullingBounds() ? 1 : 0;
    out.bounds_min[0] = Aabb.pMin.x; out.bounds_min[1] = Aabb.pMin.y; out.bounds_min[2] = Aabb.pMin.z;
    out.bounds_max[0] = Aabb.pMax.x; out.bounds_max[1] = Aabb.pMax.y; out.bounds_max[2] = Aabb.pMax.z;
}

static void put3(float* d, const RVec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

RSphere::RSphere(const RVec3& InCenter, float InRadius) : Center(InCenter), Radius(InRadius)
{
    Aabb.ExpandBySphere(Center, Radius);
}

void RSphere::Flatten(rt_shape& out) const
{
    FlattenCommon(out, RT_SHAPE_SPHERE);
    put3(out.a, Center);
    out.radius = Radius;
}

void RPlane::Flatten(rt_shape& out) const
{
    FlattenCommon(out, RT_SHAPE_PLANE);
    put3(out.a, Normal);
    put3(out.b, Point);
}

RCapsule::RCapsule(const RVec3& InStart, const RVec3& InEnd, float InRadius) : Start(InStart), End(InEnd), Radius(InRadius)
{
    Aabb.ExpandBySphere(Start, Radius);
    Aabb.ExpandBySphere(End, Radius);
}

void RCapsule::Flatten(rt_shape& out) const
{
    FlattenCommon(out, RT_SHAPE_CAPSULE);
    put3(out.a, Start);
    put3(out.b, End);
    out.radius = Radius;
}

RTriangle::RTriangle(const RVec3& p0, const RVec3& p1, const RVec3& p2)
{
    Points[0] = p0; Points[1] = p1; Points[2] = p2;
    Aabb.Expand(p0); Aabb.Expand(p1); Aabb.Expand(p2);
}

void RTriangle::Flatten(rt_shape& out) const
{
    FlattenCommon(out, RT_SHAPE_TRIANGLE);
    put3(out.a, Points[0]);
    put3(out.b, Points[1]);
    put3(out.c, Points[2]);
}

// ---- unit-vector table -------------------------------------------------------------------------
// Entry i = RMath::RandomUnitVector() (Math.h:34-40) fed with draws 2i and 2i+1 of the table
// stream: t1 = 2*PI*U0, t2 = acosf(1 - 2*U1), v = (sin t1 sin t2, cos t1 sin t2, cos t2), with the
// reference's PI = 3.1415926f (MathHelper.h:14) and libm single-precision functions.
void GenerateUnitVectors(uint32_t seed, uint32_t count, float* out)
{
    const uint32_t key = rt_rng_key(seed, RT_RNG_TABLE_PIXEL, 0);
    const float PI_REF = 3.1415926f;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (count < 65536) nt = 1;
    auto work = [=](uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; i++)
        {
            float u0 = rt_random01(key, 2 * i);
            float u1 = rt_random01(key, 2 * i + 1);
            float t1 = 2.0f * PI_REF * u0;
            float t2 = acosf(1.0f - 2.0f * u1);
            float sin_t2 = sinf(t2);
            out[3 * (size_t)i + 0] = sinf(t1) * sin_t2;
            out[3 * (size_t)i + 1] = cosf(t1) * sin_t2;
            out[3 * (size_t)i + 2] = cosf(t2);
        }
    };
    if (nt == 1) { work(0, count); return; }
    std::vector<std::thread> th;
    uint32_t chunk = (count + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++)
    {
        uint32_t lo = t * chunk, hi = lo + chunk < count ? lo + chunk : count;
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
}

// ---- scene -------------------------------------------------------------------------------------
RayTracerScene::RayTracerScene()
{
    // GSceneLights, RayTracerScene.cpp:14-18: one white point light below the origin
    LightData l;
    l.Type = RT_LIGHT_POINT;
    l.PositionOrDirection = RVec3(0.0f, -4.5f, 0.0f);
    l.Color = RVec3(1, 1, 1);
    Lights.push_back(l);
    memset(&Desc, 0, sizeof Desc);
}

void RayTracerScene::AddShape(std::unique_ptr<RShape> Shape, std::unique_ptr<ISurfaceMaterial> SurfaceMaterial)
{
    Shape->SetSurfaceMaterial(std::move(SurfaceMaterial));
    SceneShapes.push_back(std::move(Shape));
    Dirty = true;
}

void RayTracerScene::SetUnitVectors(uint32_t seed, uint32_t count)
{
    if (count == 0) count = 0xFFFFFF;      // MaxUnitVectorNums, Math.cpp:17
    UnitVectors.resize(3 * (size_t)count);
    GenerateUnitVectors(seed, count, UnitVectors.data());
    Dirty = true;
}

const rt_scene_desc& RayTracerScene::Flatten()
{
    if (!Dirty) return Desc;
    FlatShapes.clear(); FlatMaterials.clear(); FlatMeshes.clear(); FlatLights.clear();
    for (auto& sh : SceneShapes)
    {
        rt_shape r;
        sh->Flatten(r);
        if (ISurfaceMaterial* m = sh->GetSurfaceMaterial()) r.material = m->Flatten(FlatMaterials);
        if (RMeshShape* mesh = sh->AsMesh())
        {
            const FlatMesh& f = mesh->GetFlat();
            rt_mesh fm;
            fm.nodes = f.nodes.data(); fm.num_nodes = (int32_t)f.nodes.size();
            fm.tris = f.tris.data(); fm.num_tris = (int32_t)f.tris.size();
            fm.shade = f.shade.data();
            fm.textures = f.textures.data(); fm.num_textures = (int32_t)f.textures.size();
            r.mesh = (int32_t)FlatMeshes.size();
            FlatMeshes.push_back(fm);
        }
        FlatShapes.push_back(r);
    }
    for (auto& l : Lights)
    {
        rt_light fl;
        fl.type = l.Type;
        put3(fl.pos_or_dir, l.PositionOrDirection);
        put3(fl.color, l.Color);
        FlatLights.push_back(fl);
    }
    memset(&Desc, 0, sizeof Desc);
    Desc.abi_version = RT_GPU_ABI_VERSION;
    Desc.shapes = FlatShapes.data(); Desc.num_shapes = (int32_t)FlatShapes.size();
    Desc.materials = FlatMaterials.data(); Desc.num_materials = (int32_t)FlatMaterials.size();
    Desc.meshes = FlatMeshes.data(); Desc.num_meshes = (int32_t)FlatMeshes.size();
    Desc.lights = FlatLights.data(); Desc.num_lights = (int32_t)FlatLights.size();
    Desc.unit_vectors = UnitVectors.empty() ? nullptr : UnitVectors.data();
    Desc.num_unit_vectors = (uint32_t)(UnitVectors.size() / 3);
    // camera constants of ThreadWorker_Render (RayTracerProgram.cpp:133,164-165)
    Desc.eye[0] = 0.0f; Desc.eye[1] = 0.0f; Desc.eye[2] = 7.0f;
    Desc.dir_z = -0.5f;
    Desc.ray_distance = 1000.0f;
    Desc.bounce_offset = 0.0001f;          // BounceRayStartOffset, SurfaceMaterials.cpp:13
    Dirty = false;
    return Desc;
}

} // namespace rtb200
