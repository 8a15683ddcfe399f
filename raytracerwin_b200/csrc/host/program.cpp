// program.cpp — headless counterpart of the reference's program shell: the hard-coded default
// scene (RayTracerProgram::SetupScene, RayTracerProgram.cpp:467-552) and the progressive render
// driver (UpdateBitmapPixels, :270-422).  Where the reference pushes 80 ten-row tasks per pass
// into the ThreadTaskQueue (:294-302,:320-327) and waits, this driver issues one
// rt_gpu_render_tile per pass over the whole frame; the preview pass, the per-pass progress line
// and the final PNG are kept.
#include "rt_host.hpp"

#include <stdio.h>
#include <chrono>

namespace rtb200 {

template <typename T, typename... Args>
static std::unique_ptr<T> Make(Args&&... args) { return std::unique_ptr<T>(new T(std::forward<Args>(args)...)); }

void RayTracerProgram::SetupScene(const std::string& DataDir)
{
    // four spheres
    Scene.AddShape(RSphere::Create(RVec3(1.5f, 2.5f, -2.0f), 0.9f),
        Make<SurfaceMaterial_Blend>(Make<SurfaceMaterial_Reflective>(), Make<SurfaceMaterial_Diffuse>(RVec3(1.0f, 0.5f, 0.1f)), 0.5f));
    Scene.AddShape(RSphere::Create(RVec3(-1.5f, -0.5f, -3.0f), 0.5f),
        Make<SurfaceMaterial_Diffuse>(RVec3(0.1f, 1.0f, 0.2f)));
    Scene.AddShape(RSphere::Create(RVec3(0.8f, -1.5f, -1.0f), 0.5f),
        Make<SurfaceMaterial_Blend>(Make<SurfaceMaterial_Reflective>(), Make<SurfaceMaterial_Diffuse>(RVec3(0.5f, 0.0f, 0.2f)), 0.5f));
    {
        // gold sphere: emissive at half albedo.  The reference writes RVec3(0.95,0.75,0.1) * 0.5f.
        const RVec3 gold(0.95f, 0.75f, 0.1f);
        Scene.AddShape(RSphere::Create(RVec3(2.8f, -1.2f, -4.0f), 1.5f),
            Make<SurfaceMaterial_Combine>(
                Make<SurfaceMaterial_Blend>(Make<SurfaceMaterial_Reflective>(gold), Make<SurfaceMaterial_Diffuse>(gold), 0.5f),
                Make<SurfaceMaterial_Emissive>(RVec3(gold.x * 0.5f, gold.y * 0.5f, gold.z * 0.5f))));
    }
    // capsule
    Scene.AddShape(RCapsule::Create(RVec3(-1.5f, -1.5f, -1.5f), RVec3(-2.0f, -1.5f, 0.0f), 0.5f),
        Make<SurfaceMaterial_Blend>(Make<SurfaceMaterial_Reflective>(RVec3(0.8f, 0.75f, 0.6f), 0.2f),
                                    Make<SurfaceMaterial_Diffuse>(RVec3(0.25f, 0.75f, 0.6f)), 0.2f));
    // checkered ground plane
    Scene.AddShape(RPlane::Create(RVec3(0.0f, 1.0f, 0.0f), RVec3(0.0f, -2.0f, 0.0f)),
        Make<SurfaceMaterial_Blend>(Make<SurfaceMaterial_Reflective>(RVec3(1, 1, 1), 0.1f),
                                    Make<SurfaceMaterial_DiffuseChecker>(), 0.5f));
    // the textured character mesh goes last
    std::string dir = DataDir;
    if (!dir.empty() && dir.back() != '/') dir += '/';
    Scene.AddShape(RMeshShape::Create(dir + "unitychan.obj"),
        Make<SurfaceMaterial_Blend>(Make<SurfaceMaterial_Reflective>(RVec3(1, 1, 1), 0.2f),
                                    Make<SurfaceMaterial_Diffuse>(RVec3(1.0f, 1.0f, 1.0f)), 1.0f));
}

int RayTracerProgram::Run(int Device, int Width, int Height, int Passes, int MaxBounceTimes, uint32_t Seed,
                          const std::string& PngPath, double* OutSeconds, uint64_t* OutRays, std::string* Error)
{
    rt_gpu_ctx* ctx = nullptr;
    int rc = rt_gpu_create(Device, &ctx);
    if (rc != RT_OK) { if (Error) *Error = rt_gpu_last_error(nullptr); return rc; }
    auto fail = [&](int code) { if (Error) *Error = rt_gpu_last_error(ctx); rt_gpu_destroy(ctx); return code; };

    if ((rc = rt_gpu_upload_scene(ctx, &Scene.Flatten())) != RT_OK) return fail(rc);

    rt_render_params p = {};
    p.width = Width; p.height = Height;
    p.start = 0; p.end = Width * Height - 1;
    p.max_bounce = MaxBounceTimes;
    p.antialias = 1;
    p.seed = Seed;
    p.traverse = RT_TRAVERSE_CULLED;

    // base-colour preview pass (RayTracerProgram.cpp:291-312)
    p.mode = RT_MODE_PREVIEW; p.pass_begin = 0; p.pass_count = 1;
    if ((rc = rt_gpu_reset_accum(ctx, Width, Height)) != RT_OK) return fail(rc);
    if ((rc = rt_gpu_render_tile(ctx, &p)) != RT_OK) return fail(rc);
    if ((rc = rt_gpu_synchronize(ctx)) != RT_OK) return fail(rc);
    // (no reset here: like the reference, the preview pass writes bitcolor only and leaves accuBuffer alone)
    rt_gpu_reset_counters(ctx);

    p.mode = RT_MODE_PATH;
    auto t0 = std::chrono::steady_clock::now();
    auto last = t0;
    for (int Sample = 0; Sample < Passes; Sample++)
    {
        p.pass_begin = Sample; p.pass_count = 1;
        if ((rc = rt_gpu_render_tile(ctx, &p)) != RT_OK) return fail(rc);
        if ((rc = rt_gpu_synchronize(ctx)) != RT_OK) return fail(rc);
        auto now = std::chrono::steady_clock::now();
        int elapsed = (int)std::chrono::duration_cast<std::chrono::milliseconds>(now - t0).count();
        int frame = (int)std::chrono::duration_cast<std::chrono::milliseconds>(now - last).count();
        int remaining = elapsed / (Sample + 1) * (Passes - Sample - 1);
        last = now;
        printf("RayTracer - S: [%d/%d] | T: [%dms / %dms] | F: [%dms]\n", Sample + 1, Passes, elapsed, remaining, frame);
    }
    double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (OutSeconds) *OutSeconds = seconds;

    rt_counters c = {};
    if ((rc = rt_gpu_readback(ctx, RT_READ_COUNTERS_U64, &c, sizeof c)) != RT_OK) return fail(rc);
    if (OutRays) *OutRays = c.rays;

    if (!PngPath.empty())
    {
        std::vector<uint32_t> argb((size_t)Width * Height);
        if ((rc = rt_gpu_readback(ctx, RT_READ_DISPLAY_ARGB8, argb.data(), argb.size() * 4)) != RT_OK) return fail(rc);
        if (!WritePngARGB(PngPath, argb.data(), Width, Height))
        {
            if (Error) *Error = "Failed to save to " + PngPath;
            rt_gpu_destroy(ctx);
            return RT_ERR_INVALID;
        }
        printf("Image saved as %s\n", PngPath.c_str());
    }
    rt_gpu_destroy(ctx);
    return RT_OK;
}

} // namespace rtb200
