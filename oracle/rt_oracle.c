/* rt_oracle.c — TEST INFRASTRUCTURE.  Plain-C CPU restatement of the reference's per-pixel
 * ray/scene hot path, operating on the same flattened rt_scene_desc / rt_render_params the CUDA
 * path consumes (include/rt_gpu.h).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product never does.
 *
 * PARITY STATUS: the reference ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md §4), so this restatement is pinned against the reference ITSELF: the unmodified
 * sources compiled into oracle/_ref/libref_oracle.so (oracle/Makefile, oracle/ref_harness.cpp).
 * tests/test_oracle_vs_ref.py requires bit-identical primary hits, hit records, accumulated
 * colours and ray counts on every config, and tests/golden/ holds vectors generated from
 * that compiled reference (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/Src).
 * Compile with -ffp-contract=off: one rounding per operation, like the reference's Linux build.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rt_gpu.h"
#include "rt_rng.h"

typedef struct { float x, y, z; } v3;
typedef struct { v3 o, d; float dist; } ray_t;                 /* RRay, RRay.h:31-37 */
typedef struct { v3 pos, nrm; float dist; v3 color; float alpha; } hit_t;   /* RayHitResult, RRay.h:13-29 */
typedef struct { uint32_t key, n; } rng_t;                      /* position in the rand() stream */

typedef struct {
    uint64_t rays, camera_rays, shadow_rays, node_tests, tri_tests, mesh_walks;
} cnt_t;

#define FLT_EQUAL_ZERO(a) (fabsf(a) < FLT_EPSILON)              /* MathHelper.h:12 */
#define PI_REF 3.1415926f                                       /* MathHelper.h:14 */

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mulf(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }             /* RVector.h:207 */
static inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } /* :213 */
static inline float magnitude(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }        /* :158 */
static inline float fmin_ref(float a, float b) { return (a < b) ? a : b; }                      /* Math::Min, MathHelper.h:38 */
static inline float fmax_ref(float a, float b) { return (a > b) ? a : b; }                      /* Math::Max, MathHelper.h:35 */
static inline v3 ld3(const float* p) { return V(p[0], p[1], p[2]); }

/* RVec3::GetNormalizedVec3, RVector.h:169-183: short vectors are returned unchanged */
static inline v3 normalized(v3 a)
{
    float sqr_mag = a.x * a.x + a.y * a.y + a.z * a.z;
    if (!FLT_EQUAL_ZERO(sqr_mag)) {
        float one_over_mag = 1.0f / sqrtf(sqr_mag);
        return V(a.x * one_over_mag, a.y * one_over_mag, a.z * one_over_mag);
    }
    return a;
}

/* Math::Q_rsqrt, MathHelper.cpp:26-38 */
static inline float q_rsqrt(float number)
{
    const float x2 = number * 0.5F;
    union { float f; uint32_t i; } conv;
    conv.f = number;
    conv.i = 0x5f3759df - (conv.i >> 1);
    conv.f *= (1.5F - (x2 * conv.f * conv.f));
    return conv.f;
}

/* RVec3::GetNormalizedVec3_Fast, RVector.h:185-199 */
static inline v3 normalized_fast(v3 a)
{
    float sqr_mag = a.x * a.x + a.y * a.y + a.z * a.z;
    if (!FLT_EQUAL_ZERO(sqr_mag)) {
        float one_over_mag = q_rsqrt(sqr_mag);
        return V(a.x * one_over_mag, a.y * one_over_mag, a.z * one_over_mag);
    }
    return a;
}

/* RVec3::Reflect, RVector.h:218-221: *this - normal * 2.0f * Dot(*this, normal) */
static inline v3 reflect(v3 v, v3 n) { return sub(v, mulf(mulf(n, 2.0f), dot(v, n))); }

static inline float rng_random(rng_t* r) { return rt_random01(r->key, r->n++); }   /* RMath::Random, Math.h:17-20 */
static inline int32_t rng_rand(rng_t* r) { return rt_rand31(r->key, r->n++); }

/* ---- primitives --------------------------------------------------------------------------- */

/* RRay::TestIntersectionWithAabb, RRay.cpp:89-136 — a LINE test: accept iff tmax > tmin */
static int slab_test(const ray_t* r, const float* bmin, const float* bmax, float* t)
{
    float tmin = -FLT_MAX, tmax = FLT_MAX;
    if (!FLT_EQUAL_ZERO(r->d.x)) {
        float inv = 1.0f / r->d.x;
        float t1 = (bmin[0] - r->o.x) * inv, t2 = (bmax[0] - r->o.x) * inv;
        tmin = fmax_ref(tmin, fmin_ref(t1, t2)); tmax = fmin_ref(tmax, fmax_ref(t1, t2));
    }
    if (!FLT_EQUAL_ZERO(r->d.y)) {
        float inv = 1.0f / r->d.y;
        float t1 = (bmin[1] - r->o.y) * inv, t2 = (bmax[1] - r->o.y) * inv;
        tmin = fmax_ref(tmin, fmin_ref(t1, t2)); tmax = fmin_ref(tmax, fmax_ref(t1, t2));
    }
    if (!FLT_EQUAL_ZERO(r->d.z)) {
        float inv = 1.0f / r->d.z;
        float t1 = (bmin[2] - r->o.z) * inv, t2 = (bmax[2] - r->o.z) * inv;
        tmin = fmax_ref(tmin, fmin_ref(t1, t2)); tmax = fmin_ref(tmax, fmax_ref(t1, t2));
    }
    if (tmax > tmin) { if (t) *t = tmin; return 1; }
    return 0;
}

/* RRay::TestIntersectionWithTriangleAndFaceNormal, RRay.cpp:147-213.  Writes pos/nrm/dist only. */
static int triangle_test(const ray_t* r, v3 p0, v3 p1, v3 p2, v3 n, v3* pos, v3* nrm, float* dist)
{
    v3 end = add(r->o, mulf(r->d, r->dist));
    float d0 = dot(n, r->o);
    float d1 = dot(n, p0);
    float d2 = d0 - d1;
    if (d2 < 0) return 0;
    if (dot(end, n) - d1 > 0) return 0;
    v3 l = sub(end, r->o);
    float d3 = dot(n, l);
    if (FLT_EQUAL_ZERO(d3)) return 0;
    float df = -(d2 / d3);
    v3 cp = add(r->o, mulf(l, df));
    const v3 P[3] = { p0, p1, p2 };
    for (int i = 0; i < 3; i++) {
        v3 edge = sub(P[(i + 1) % 3], P[i]);
        v3 edge_normal = cross(edge, n);
        if (dot(edge_normal, sub(cp, P[i])) > 0) return 0;
    }
    *pos = cp; *nrm = n; *dist = magnitude(mulf(l, df));
    return 1;
}

/* RRay::TestIntersectionWithSphere, RRay.cpp:25-64 */
static int sphere_test(const ray_t* r, v3 c, float radius, v3* pos, v3* nrm, float* dist)
{
    float dx = r->d.x * r->dist, dy = r->d.y * r->dist, dz = r->d.z * r->dist;
    float _a = dx * dx + dy * dy + dz * dz;
    float _b = 2 * dx * (r->o.x - c.x) + 2 * dy * (r->o.y - c.y) + 2 * dz * (r->o.z - c.z);
    float _c = c.x * c.x + c.y * c.y + c.z * c.z + r->o.x * r->o.x + r->o.y * r->o.y + r->o.z * r->o.z +
               -2 * (c.x * r->o.x + c.y * r->o.y + c.z * r->o.z) - radius * radius;
    float d = _b * _b - 4 * _a * _c;
    if (d >= 0) {
        float t = (-_b - sqrtf(d)) / (_a * 2);
        if (t <= 0) return 0;
        v3 hp = V(r->o.x + t * dx, r->o.y + t * dy, r->o.z + t * dz);
        float dd = magnitude(sub(hp, r->o));
        if (dd > r->dist) return 0;
        *pos = hp; *nrm = normalized(sub(hp, c)); *dist = dd;
        return 1;
    }
    return 0;
}

/* RRay::TestIntersectionWithPlane, RRay.cpp:66-87 */
static int plane_test(const ray_t* r, v3 n, v3 p, v3* pos, v3* nrm, float* dist)
{
    float denom = dot(n, r->d);
    if (fabsf(denom) > 1e-6) {
        v3 p0l0 = sub(p, r->o);
        float t = dot(p0l0, n) / denom;
        if (t >= 0 && t < r->dist) {
            *pos = add(r->o, mulf(r->d, t)); *nrm = n; *dist = t;
            return 1;
        }
    }
    return 0;
}

/* RCapsule::TestRayCylinderIntersection, Shapes.cpp:65-125 (no comparison with ray length) */
static int cylinder_test(const ray_t* r, v3 start, v3 endp, float radius, v3* pos, v3* nrm, float* dist)
{
    v3 d = sub(endp, start);
    v3 m = sub(r->o, start);
    float dd = dot(d, d), nd = dot(r->d, d), mn = dot(m, r->d), md = dot(m, d), mm = dot(m, m);
    if (dot(sub(r->o, start), sub(endp, start)) < 0 && dot(r->d, sub(endp, start)) < 0) return 0;
    if (dot(sub(r->o, endp), sub(start, endp)) < 0 && dot(r->d, sub(start, endp)) < 0) return 0;
    float a = dd - nd * nd;
    float b = dd * mn - nd * md;
    float c = dd * (mm - radius * radius) - md * md;
    if (fabs(a) < FLT_EPSILON) return 0;
    if ((b * b - a * c) < 0) return 0;
    float r_t = (-b - sqrtf(b * b - a * c)) / a;
    if (r_t < 0) return 0;
    v3 v = add(r->o, mulf(r->d, r_t));
    if (dot(sub(v, start), sub(endp, start)) < 0) return 0;
    if (dot(sub(v, endp), sub(start, endp)) < 0) return 0;
    *dist = r_t;
    *pos = add(r->o, mulf(r->d, r_t));
    v3 side = cross(sub(endp, start), sub(*pos, start));
    *nrm = normalized(cross(side, sub(endp, start)));
    return 1;
}

/* RMath::Barycentric, Math.cpp:56-68 */
static void barycentric(v3 p, v3 a, v3 b, v3 c, float* u, float* v, float* w)
{
    v3 v0 = sub(b, a), v1 = sub(c, a), v2 = sub(p, a);
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    *v = (d11 * d20 - d01 * d21) / denom;
    *w = (d00 * d21 - d01 * d20) / denom;
    *u = 1.0f - *v - *w;
}

static inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }   /* Math::Lerp, MathHelper.h:40 */

/* (int)f with f = NaN or out of range is undefined in the reference; pin it to the x86
 * cvttss2si result (INT_MIN), then clamp the texel index into the image (documented divergence,
 * the reference reads out of bounds there). */
static inline int to_int_ref(float f)
{
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int)f;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* one RVec4 texel: stored as floats, or as the PNG's 8-bit codes + the host's table (rt_gpu.h rt_texture) —
 * RTexture::LoadTexturePNG's per-texel conversion, Texture.cpp:119-151 */
static inline void texel(const rt_texture* t, int x, int y, float out[4])
{
    const size_t i = (size_t)y * t->width + x;
    if (t->rgba) { memcpy(out, t->rgba + 4 * i, 4 * sizeof(float)); return; }
    const uint8_t* p = t->texels8 + i * (size_t)t->channels;
    out[0] = t->lut[p[0]]; out[1] = t->lut[p[1]]; out[2] = t->lut[p[2]];
    out[3] = t->channels == 4 ? t->lut[256 + p[3]] : 1.0f;
}

/* RTexture::Sample, Texture.cpp:23-57 */
static void texture_sample(const rt_texture* t, float u, float v, float out[4])
{
    float cu = u - floorf(u), cv = v - floorf(v);
    float fx = cu * (t->width - 1), fy = cv * (t->height - 1);
    int x0 = to_int_ref(floorf(fx)), y0 = to_int_ref(floorf(fy));
    int x1 = to_int_ref(ceilf(fx)), y1 = to_int_ref(ceilf(fy));
    float dx = fx - x0, dy = fy - y0;
    int cx0 = clampi(x0, 0, t->width - 1), cx1 = clampi(x1, 0, t->width - 1);
    int cy0 = clampi(y0, 0, t->height - 1), cy1 = clampi(y1, 0, t->height - 1);
    float p00[4], p01[4], p10[4], p11[4];
    texel(t, cx0, cy0, p00); texel(t, cx1, cy0, p01); texel(t, cx0, cy1, p10); texel(t, cx1, cy1, p11);
    for (int k = 0; k < 4; k++)
        out[k] = lerpf(lerpf(p00[k], p01[k], dx), lerpf(p10[k], p11[k], dx), dy);
}

/* ---- mesh ---------------------------------------------------------------------------------- */

/* KdNode::TestRayIntersection, KdTree.cpp:128-195, on the pre-order array: entering a node whose
 * slab test passes continues at i+1 (Left, then Right by pre-order), everything else jumps to
 * `escape`.  TestRay.Distance shrinks at every accepted leaf (:176) and the last accepted leaf
 * wins.  Returns the leaf-order triangle slot or -1. */
static int bvh_traverse(const rt_mesh* m, ray_t* ray, v3* pos, v3* nrm, float* dist, cnt_t* c)
{
    int best = -1;
    int i = 0;
    const int n = m->num_nodes;
    while (i < n) {
        const rt_bvh_node* node = &m->nodes[i];
        c->node_tests++;
        if (!slab_test(ray, node->bmin, node->bmax, NULL)) { i = node->escape; continue; }
        if (node->tri < 0) { i = i + 1; continue; }
        const rt_tri* t = &m->tris[node->tri];
        v3 hp, hn; float hd;
        c->tri_tests++;
        if (triangle_test(ray, ld3(t->p0), ld3(t->p1), ld3(t->p2), ld3(t->n), &hp, &hn, &hd)) {
            ray->dist = hd;
            *pos = hp; *nrm = hn; *dist = hd;
            best = node->tri;
        }
        i = node->escape;
    }
    return best;
}

/* RMeshShape::TestRayIntersection, MeshShape.cpp:280-331.  On a hit the whole record is
 * replaced (KdTree.cpp:178-181 assigns a fresh RayHitResult: colour 1, alpha 1). */
static int mesh_test(const rt_mesh* m, const ray_t* in, hit_t* out, int* tri_index, cnt_t* c)
{
    ray_t ray = *in;                              /* KdTree::TestRayIntersection copies the ray, KdTree.cpp:229 */
    v3 pos, nrm; float dist;
    if (m->num_nodes > 0) c->mesh_walks++;
    int slot = bvh_traverse(m, &ray, &pos, &nrm, &dist, c);
    if (slot < 0) return 0;
    const rt_tri* t = &m->tris[slot];
    if (tri_index) *tri_index = t->index;
    if (out) {
        out->pos = pos; out->dist = dist;
        out->color = V(1.0f, 1.0f, 1.0f); out->alpha = 1.0f;
        float u, v, w;
        barycentric(pos, ld3(t->p0), ld3(t->p1), ld3(t->p2), &u, &v, &w);
        const rt_shade* s = &m->shade[t->index];
        v3 nn = add(add(mulf(ld3(s->n0), u), mulf(ld3(s->n1), v)), mulf(ld3(s->n2), w));
        out->nrm = normalized_fast(nn);
        if (s->texture >= 0) {
            /* t0*u + t1*v + t2*w, then Sample(x, 1 - y)  (MeshShape.cpp:316-324) */
            float tx = s->uv0[0] * u + s->uv1[0] * v + s->uv2[0] * w;
            float ty = s->uv0[1] * u + s->uv1[1] * v + s->uv2[1] * w;
            float rgba[4];
            texture_sample(&m->textures[s->texture], tx, 1.0f - ty, rgba);
            out->color = V(rgba[0], rgba[1], rgba[2]);
            out->alpha = rgba[3];
        }
    }
    return 1;
}

/* ---- scene --------------------------------------------------------------------------------- */

/* one shape's TestRayIntersection; writes into *out exactly the fields the reference writes */
static int shape_test(const rt_scene_desc* sc, const rt_shape* sh, const ray_t* ray, hit_t* out, int* tri, cnt_t* c)
{
    v3 pos, nrm; float dist;
    switch (sh->type) {
    case RT_SHAPE_SPHERE:                                        /* Shapes.cpp:18-21 */
        if (!sphere_test(ray, ld3(sh->a), sh->radius, &pos, &nrm, &dist)) return 0;
        if (out) { out->pos = pos; out->nrm = nrm; out->dist = dist; }
        return 1;
    case RT_SHAPE_PLANE:                                         /* Shapes.cpp:23-26 */
        if (!plane_test(ray, ld3(sh->a), ld3(sh->b), &pos, &nrm, &dist)) return 0;
        if (out) { out->pos = pos; out->nrm = nrm; out->dist = dist; }
        return 1;
    case RT_SHAPE_TRIANGLE: {                                    /* Shapes.cpp:127-130 */
        v3 p0 = ld3(sh->a), p1 = ld3(sh->b), p2 = ld3(sh->c);
        v3 n = normalized(cross(sub(p1, p0), sub(p2, p0)));      /* RRay.cpp:138-145 */
        if (!triangle_test(ray, p0, p1, p2, n, &pos, &nrm, &dist)) return 0;
        if (out) { out->pos = pos; out->nrm = nrm; out->dist = dist; }
        return 1;
    }
    case RT_SHAPE_CAPSULE: {                                     /* Shapes.cpp:34-63 */
        if (cylinder_test(ray, ld3(sh->a), ld3(sh->b), sh->radius, &pos, &nrm, &dist)) {
            if (out) { out->dist = dist; out->pos = pos; out->nrm = nrm; }
            return 1;
        }
        v3 p1, n1, p2, n2; float d1 = 0.0f, d2 = 0.0f;
        int b1 = sphere_test(ray, ld3(sh->a), sh->radius, &p1, &n1, &d1);
        int b2 = sphere_test(ray, ld3(sh->b), sh->radius, &p2, &n2, &d2);
        if (out && (b1 || b2)) {
            int first = (b1 && b2) ? (d1 < d2) : b1;
            /* whole-struct assignment from a fresh RayHitResult: colour/alpha reset to 1 */
            out->pos = first ? p1 : p2; out->nrm = first ? n1 : n2; out->dist = first ? d1 : d2;
            out->color = V(1.0f, 1.0f, 1.0f); out->alpha = 1.0f;
        }
        return b1 || b2;
    }
    case RT_SHAPE_MESH:
        if (sh->mesh < 0) return 0;
        return mesh_test(&sc->meshes[sh->mesh], ray, out, tri, c);
    }
    return 0;
}

/* RayTracerScene::FindIntersectionWithScene, RayTracerScene.cpp:99-125 */
static int find_intersection(const rt_scene_desc* sc, ray_t test, hit_t* out, int* tri_out, cnt_t* c)
{
    int hit_shape = -1;
    c->rays++;
    for (int i = 0; i < sc->num_shapes; i++) {
        const rt_shape* sh = &sc->shapes[i];
        int enter = !sh->has_bounds;
        if (!enter) { c->node_tests++; enter = slab_test(&test, sh->bounds_min, sh->bounds_max, NULL); }
        if (enter) {
            int tri = -1;
            if (shape_test(sc, sh, &test, out, &tri, c)) {
                test.dist = out->dist;
                hit_shape = i;
                if (tri_out) *tri_out = tri;
            }
        }
    }
    return hit_shape;
}

/* shadow query of CalculateLightColor, RayTracerScene.cpp:152-164: any accepted hit */
static int occluded(const rt_scene_desc* sc, const ray_t* shadow, cnt_t* c)
{
    c->rays++; c->shadow_rays++;
    for (int i = 0; i < sc->num_shapes; i++) {
        const rt_shape* sh = &sc->shapes[i];
        int enter = !sh->has_bounds;
        if (!enter) { c->node_tests++; enter = slab_test(shadow, sh->bounds_min, sh->bounds_max, NULL); }
        if (enter && shape_test(sc, sh, shadow, NULL, NULL, c)) return 1;
    }
    return 0;
}

/* RayTracerScene::CalculateLightColor, RayTracerScene.cpp:127-175 */
static v3 light_color(const rt_scene_desc* sc, const rt_light* l, const hit_t* h, v3 surface, cnt_t* c)
{
    v3 ldir = ld3(l->pos_or_dir);
    float dist = 0.0f;
    if (l->type == RT_LIGHT_POINT) {
        v3 lp = ld3(l->pos_or_dir);
        ldir = normalized(sub(lp, h->pos));
        dist = magnitude(sub(h->pos, lp));
    } else if (l->type == RT_LIGHT_DIRECTIONAL) {
        dist = 1000.0f;
    }
    ray_t shadow = { add(h->pos, mulf(ldir, sc->bounce_offset)), ldir, dist };
    if (occluded(sc, &shadow, c)) return V(0, 0, 0);
    float ldp = fmax_ref(0.0f, dot(h->nrm, ldir));
    return mulf(surface, ldp);
}

/* ---- materials ----------------------------------------------------------------------------- */
typedef struct { v3 att, emi; } bounce_t;                      /* ViewRayBounceResult */

/* RMath::RandomUnitVector, Math.h:34-40 */
static v3 random_unit_vector(rng_t* rng)
{
    float t1 = 2.0f * PI_REF * rng_random(rng);
    float t2 = acosf(1.0f - 2.0f * rng_random(rng));
    float sin_t2 = sinf(t2);
    return V(sinf(t1) * sin_t2, cosf(t1) * sin_t2, cosf(t2));
}

/* RMath::RandomHemisphereDirection, Math.cpp:42-54, with the table index drawn from rand()
 * (oracle/ref_math_wrap.cpp) instead of the shared cursor of Math.cpp:33-40 */
static v3 random_hemisphere(const rt_scene_desc* sc, v3 n, rng_t* rng)
{
    uint32_t idx = (uint32_t)rng_rand(rng) % sc->num_unit_vectors;
    v3 v = ld3(sc->unit_vectors + 3 * (size_t)idx);
    if (dot(v, n) > 0.0f) return v;
    return reflect(v, n);
}

/* SurfaceMaterial_DiffuseChecker::IsBrighterArea, SurfaceMaterials.cpp:66-90 */
static int checker_bright(v3 p, float recip)
{
    int r = 0;
    float fx = p.x * recip, fy = p.y * recip, fz = p.z * recip;
    if (fx - floorf(fx) > 0.5f) r = !r;
    if (fz - floorf(fz) > 0.5f) r = !r;
    if (fy - floorf(fy) > 0.5f) r = !r;
    return r;
}

/* ISurfaceMaterial::BounceViewRay for the seven classes, SurfaceMaterials.cpp:20-187 */
static bounce_t bounce(const rt_scene_desc* sc, int node, const ray_t* in, const hit_t* h, ray_t* out, rng_t* rng)
{
    bounce_t r = { V(0, 0, 0), V(0, 0, 0) };
    if (node < 0) return r;
    const rt_material* m = &sc->materials[node];
    switch (m->type) {
    case RT_MAT_DIFFUSE:
    case RT_MAT_CHECKER: {
        float factor = 1.0f;
        if (m->type == RT_MAT_CHECKER) factor = checker_bright(h->pos, m->scalar) ? 1.0f : 0.5f;   /* :55 */
        float remaining = in->dist - h->dist;                     /* :23 */
        v3 dir = random_hemisphere(sc, h->nrm, rng);              /* :26 */
        out->o = add(h->pos, mulf(dir, sc->bounce_offset)); out->d = dir; out->dist = remaining;   /* :27 */
        float dp = fmax_ref(0.0f, dot(h->nrm, dir));              /* :30 */
        r.att = mulf(ld3(m->rgb), dp);                            /* :32 */
        if (m->type == RT_MAT_CHECKER) r.att = mulf(r.att, factor);   /* :57 */
        return r;
    }
    case RT_MAT_REFLECTIVE: {                                     /* :98-125 */
        float remaining = in->dist - h->dist;
        v3 nd = reflect(in->d, h->nrm);
        if (m->scalar > 0.0f) {
            nd = add(nd, mulf(random_unit_vector(rng), m->scalar));
            nd = normalized(nd);
        }
        out->o = add(h->pos, mulf(nd, sc->bounce_offset)); out->d = nd; out->dist = remaining;
        r.att = ld3(m->rgb);
        return r;
    }
    case RT_MAT_EMISSIVE:                                         /* :132-138 */
        *out = *in;
        r.emi = ld3(m->rgb);
        return r;
    case RT_MAT_BLEND:                                            /* :153-156 */
        return rng_random(rng) > m->scalar ? bounce(sc, m->child_a, in, h, out, rng)
                                           : bounce(sc, m->child_b, in, h, out, rng);
    case RT_MAT_COMBINE: {                                        /* :169-172 */
        /* `A->Bounce(..) + B->Bounce(..)`: operand order is unspecified in C++11; the compiled
         * reference (g++ 13, -O2) evaluates B first and A last, so the outgoing ray and the
         * later RNG draws are A's.  Pinned by tests/test_oracle_vs_ref.py. */
        bounce_t rb = bounce(sc, m->child_b, in, h, out, rng);
        bounce_t ra = bounce(sc, m->child_a, in, h, out, rng);
        r.att = add(ra.att, rb.att); r.emi = add(ra.emi, rb.emi);
        return r;
    }
    case RT_MAT_NULL: {                                           /* :179-187 */
        float remaining = in->dist - h->dist;
        out->o = add(h->pos, mulf(in->d, sc->bounce_offset)); out->d = in->d; out->dist = remaining;
        r.att = V(1, 1, 1);
        return r;
    }
    }
    return r;
}

/* ISurfaceMaterial::PreviewColor, SurfaceMaterials.cpp:35-38,60-64,122-125,140-143,158-161,174-177,189-192 */
static v3 preview_color(const rt_scene_desc* sc, int node, const hit_t* h, rng_t* rng)
{
    if (node < 0) return V(0, 0, 0);
    const rt_material* m = &sc->materials[node];
    switch (m->type) {
    case RT_MAT_DIFFUSE:
        return mulf(ld3(m->rgb), dot(h->nrm, V(0, 1, 0)) * 0.5f + 0.5f);
    case RT_MAT_CHECKER: {
        float factor = checker_bright(h->pos, m->scalar) ? 1.0f : 0.5f;
        return mulf(mulf(ld3(m->rgb), dot(h->nrm, V(0, 1, 0)) * 0.5f + 0.5f), factor);
    }
    case RT_MAT_REFLECTIVE: return ld3(m->rgb);
    case RT_MAT_EMISSIVE: return ld3(m->rgb);
    case RT_MAT_BLEND:
        return rng_random(rng) > m->scalar ? preview_color(sc, m->child_a, h, rng) : preview_color(sc, m->child_b, h, rng);
    case RT_MAT_COMBINE: {
        v3 b = preview_color(sc, m->child_b, h, rng);     /* same operand order as BounceViewRay */
        v3 a = preview_color(sc, m->child_a, h, rng);
        return add(a, b);
    }
    case RT_MAT_NULL: return V(0, 0, 0);
    }
    return V(0, 0, 0);
}

static inline int is_non_zero(v3 a) { return !FLT_EQUAL_ZERO(a.x) && !FLT_EQUAL_ZERO(a.y) && !FLT_EQUAL_ZERO(a.z); }  /* RVector.h:142-145 */

static inline v3 sky_color(v3 d)                                  /* RayTracerScene.cpp:92-93 */
{
    float t = 0.5f * (d.y + 1.0f);
    return add(mulf(V(1.0f, 1.0f, 1.0f), 1.0f - t), mulf(V(0.5f, 0.7f, 1.0f), t));
}

/* RayTracerScene::RayTrace, RayTracerScene.cpp:31-97 */
static v3 ray_trace(const rt_scene_desc* sc, const ray_t* in, int max_bounce, int preview, rng_t* rng, cnt_t* c)
{
    if (max_bounce == 0) return V(0, 0, 0);
    v3 final = V(0, 0, 0);
    hit_t h = { {0,0,0}, {0,0,0}, 0.0f, {1.0f, 1.0f, 1.0f}, 1.0f };
    int shape = find_intersection(sc, *in, &h, NULL, c);
    if (shape != -1) {
        int mat = sc->shapes[shape].material;
        if (preview) {
            if (mat >= 0) final = add(final, mulv(preview_color(sc, mat, &h, rng), h.color));
        } else if (mat >= 0) {
            ray_t out = { {0,0,0}, {0,0,0}, 0.0f };
            bounce_t b = bounce(sc, mat, in, &h, &out, rng);
            if (rng_random(rng) <= h.alpha) {
                if (is_non_zero(b.att))
                    final = add(final, mulv(mulv(b.att, ray_trace(sc, &out, max_bounce - 1, preview, rng, c)), h.color));
                final = add(final, b.emi);
            } else {
                float remaining = in->dist - h.dist;
                ray_t pass = { add(h.pos, mulf(in->d, sc->bounce_offset)), in->d, remaining };
                final = add(final, ray_trace(sc, &pass, max_bounce - 1, preview, rng, c));
            }
        }
    } else {
        return sky_color(in->d);
    }
    return final;
}

/* Whitted config, composed exactly as oracle/ref_harness.cpp composes it from the
 * reference's functions: nearest hit, then Σ_lights CalculateLightColor(light, hit, SampledColor);
 * sky on a miss. */
static v3 whitted(const rt_scene_desc* sc, const ray_t* in, cnt_t* c)
{
    hit_t h = { {0,0,0}, {0,0,0}, 0.0f, {1.0f, 1.0f, 1.0f}, 1.0f };
    int shape = find_intersection(sc, *in, &h, NULL, c);
    if (shape == -1) return sky_color(in->d);
    v3 col = V(0, 0, 0);
    for (int i = 0; i < sc->num_lights; i++) col = add(col, light_color(sc, &sc->lights[i], &h, h.color, c));
    return col;
}

/* ---- camera + per-pixel worker --------------------------------------------------------------- */

/* ThreadWorker_Render's ray generator, RayTracerProgram.cpp:133-165, W/H parametric */
static void pixel_base(const rt_render_params* p, int pixel, float* dx, float* dy)
{
    int x = pixel % p->width, y = pixel / p->width;                /* ColorBuffer.h:19-23 */
    float aspect = (float)p->width / (float)p->height;
    *dx = -(float)(x - p->width / 2) / (p->width * 2) * aspect;
    *dy = -(float)(y - p->height / 2) / (p->height * 2);
}

static ray_t camera_ray(const rt_scene_desc* sc, const rt_render_params* p, int pixel, int sub, rng_t* rng)
{
    float dx, dy;
    pixel_base(p, pixel, &dx, &dy);
    float ox = 0.0f, oy = 0.0f;
    if (sub >= 0) {
        const float inv_pixel_radius = 1.0f / (p->width * 4);
        const float offset_radius = inv_pixel_radius * 0.5f;
        ox = (sub & 1) ? inv_pixel_radius : 0.0f;
        oy = (sub & 2) ? inv_pixel_radius : 0.0f;
        ox += (rng_random(rng) - 0.5f) * offset_radius;
        oy += (rng_random(rng) - 0.5f) * offset_radius;
    }
    ray_t r = { ld3(sc->eye), normalized(V(dx + ox, dy + oy, sc->dir_z)), sc->ray_distance };
    return r;
}

/* LinearToGamma + MakePixelColor, ColorBuffer.h:81-109 (non-OSX ARGB packing) */
static uint32_t make_pixel(v3 lin)
{
    const float e = 1.0f / 2.2f;
    v3 g = V(powf(lin.x, e), powf(lin.y, e), powf(lin.z, e));
    int r = (int)(fmin_ref(fmax_ref(g.x, 0.0f), 1.0f) * 255);
    int gg = (int)(fmin_ref(fmax_ref(g.y, 0.0f), 1.0f) * 255);
    int b = (int)(fmin_ref(fmax_ref(g.z, 0.0f), 1.0f) * 255);
    return (255u << 24) | ((uint32_t)(r & 255) << 16) | ((uint32_t)(gg & 255) << 8) | (uint32_t)(b & 255);
}

static int owns_pixel(const rt_render_params* p, int pixel)
{
    if (p->tile_count <= 1 || p->tile_size <= 0) return 1;
    int x = pixel % p->width, y = pixel / p->width;
    int tiles_x = (p->width + p->tile_size - 1) / p->tile_size;
    int tile = (y / p->tile_size) * tiles_x + x / p->tile_size;
    return tile % p->tile_count == p->tile_rank;
}

typedef struct {
    const rt_scene_desc* sc; const rt_render_params* p;
    float* accum; uint32_t* display; int32_t* ids; float* pdist; float* preview;
    int next_row; pthread_mutex_t lock; cnt_t total;
} job_t;

static void render_pixel(job_t* j, int pixel, cnt_t* c)
{
    const rt_scene_desc* sc = j->sc; const rt_render_params* p = j->p;
    if (p->mode == RT_MODE_PRIMARY) {
        rng_t rng = { 0, 0 };
        ray_t r = camera_ray(sc, p, pixel, -1, &rng);
        hit_t h = { {0,0,0}, {0,0,0}, 0.0f, {1.0f, 1.0f, 1.0f}, 1.0f };
        int tri = -1;
        c->camera_rays++;
        int shape = find_intersection(sc, r, &h, &tri, c);
        if (j->ids) { j->ids[2 * (size_t)pixel] = shape; j->ids[2 * (size_t)pixel + 1] = shape >= 0 ? tri : -1; }
        if (j->pdist) j->pdist[pixel] = shape >= 0 ? h.dist : 0.0f;
        return;
    }
    float* a = j->accum + 4 * (size_t)pixel;
    v3 sum = V(a[0], a[1], a[2]);
    int num = (int)a[3];
    v3 last = V(0, 0, 0);
    for (int pass = p->pass_begin; pass < p->pass_begin + p->pass_count; pass++) {
        v3 col = V(0, 0, 0);
        if (p->antialias) {
            for (int i = 0; i < 4; i++) {
                rng_t rng = { rt_rng_key(p->seed, (uint32_t)pixel, (uint32_t)(pass * 4 + i)), 0 };
                ray_t r = camera_ray(sc, p, pixel, i, &rng);
                c->camera_rays++;
                v3 s = p->mode == RT_MODE_WHITTED ? whitted(sc, &r, c)
                                                  : ray_trace(sc, &r, p->max_bounce, p->mode == RT_MODE_PREVIEW, &rng, c);
                col = add(col, s);
            }
            col = V(col.x / 4.0f, col.y / 4.0f, col.z / 4.0f);     /* RayTracerProgram.cpp:169 */
        } else {
            rng_t rng = { rt_rng_key(p->seed, (uint32_t)pixel, (uint32_t)pass), 0 };
            ray_t r = camera_ray(sc, p, pixel, -1, &rng);
            c->camera_rays++;
            col = p->mode == RT_MODE_WHITTED ? whitted(sc, &r, c)
                                             : ray_trace(sc, &r, p->max_bounce, p->mode == RT_MODE_PREVIEW, &rng, c);
        }
        sum = add(sum, col); num++;                                 /* AccumulatePixel::AddPixel, :57-61 */
        last = col;
    }
    if (p->mode == RT_MODE_PREVIEW) {
        /* UseBaseColor: bitcolor only, accuBuffer is left alone (RayTracerProgram.cpp:175-180) */
        if (j->preview) { float* q = j->preview + 4 * (size_t)pixel; q[0] = last.x; q[1] = last.y; q[2] = last.z; q[3] = 1.0f; }
        if (j->display) j->display[pixel] = make_pixel(last);
        return;
    }
    a[0] = sum.x; a[1] = sum.y; a[2] = sum.z; a[3] = (float)num;
    if (j->display) {
        float fn = (float)num;
        j->display[pixel] = make_pixel(V(sum.x / fn, sum.y / fn, sum.z / fn));               /* :68-71, :185 */
    }
}

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    const rt_render_params* p = j->p;
    cnt_t c; memset(&c, 0, sizeof c);
    const int rows_per_task = 10;                                   /* NumTaskRows, RayTracerProgram.cpp:282 */
    const int last_row = p->end / p->width;
    for (;;) {
        int row = __atomic_fetch_add(&j->next_row, rows_per_task, __ATOMIC_RELAXED);
        if (row > last_row) break;
        int s = row * p->width; if (s < p->start) s = p->start;
        int e = (row + rows_per_task) * p->width - 1; if (e > p->end) e = p->end;
        for (int px = s; px <= e; px++) if (owns_pixel(p, px)) render_pixel(j, px, &c);
    }
    pthread_mutex_lock(&j->lock);
    j->total.rays += c.rays; j->total.camera_rays += c.camera_rays; j->total.shadow_rays += c.shadow_rays;
    j->total.node_tests += c.node_tests; j->total.tri_tests += c.tri_tests; j->total.mesh_walks += c.mesh_walks;
    pthread_mutex_unlock(&j->lock);
    return NULL;
}

/* CPU counterpart of rt_gpu_render_tile + readback.  accum is width*height*4 floats and is
 * ADDED to (pass it zeroed for a fresh frame); display / ids / pdist / counters / preview may be NULL.
 * RT_MODE_PREVIEW leaves accum alone (like the reference) and writes the pass colour to preview
 * (width*height*4 floats, the counterpart of RT_READ_PREVIEW_RGBA_F32). */
int rt_oracle_render_ex(const rt_scene_desc* sc, const rt_render_params* p, int nthreads,
                        float* accum, uint32_t* display, int32_t* ids, float* pdist, rt_counters* counters, float* preview);

int rt_oracle_render(const rt_scene_desc* sc, const rt_render_params* p, int nthreads,
                     float* accum, uint32_t* display, int32_t* ids, float* pdist, rt_counters* counters)
{
    return rt_oracle_render_ex(sc, p, nthreads, accum, display, ids, pdist, counters, NULL);
}

int rt_oracle_render_ex(const rt_scene_desc* sc, const rt_render_params* p, int nthreads,
                        float* accum, uint32_t* display, int32_t* ids, float* pdist, rt_counters* counters, float* preview)
{
    if (!sc || !p || p->width <= 0 || p->height <= 0 || p->start < 0 || p->end >= p->width * p->height) return RT_ERR_INVALID;
    job_t j; memset(&j, 0, sizeof j);
    j.sc = sc; j.p = p; j.accum = accum; j.display = display; j.ids = ids; j.pdist = pdist; j.preview = preview;
    j.next_row = p->start / p->width;
    pthread_mutex_init(&j.lock, NULL);
    if (nthreads <= 1) worker(&j);
    else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
        for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, &j);
        for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&j.lock);
    if (counters) {
        memset(counters, 0, sizeof *counters);
        counters->rays = j.total.rays; counters->camera_rays = j.total.camera_rays; counters->shadow_rays = j.total.shadow_rays;
        counters->node_tests = j.total.node_tests; counters->tri_tests = j.total.tri_tests;
        counters->node_visits = j.total.node_tests; counters->tri_visits = j.total.tri_tests;
        counters->mesh_walks = j.total.mesh_walks;
    }
    return RT_OK;
}

/* arbitrary rays (7 floats: origin, direction, distance) through find_intersection.
 * hit11 = pos, nrm, dist, color, alpha (zeros on a miss) */
int rt_oracle_trace_rays(const rt_scene_desc* sc, const float* rays, int n, int32_t* shape, int32_t* tri, float* hit11)
{
    cnt_t c; memset(&c, 0, sizeof c);
    for (int k = 0; k < n; k++) {
        const float* q = rays + 7 * (size_t)k;
        ray_t r = { V(q[0], q[1], q[2]), V(q[3], q[4], q[5]), q[6] };
        hit_t h = { {0,0,0}, {0,0,0}, 0.0f, {1.0f, 1.0f, 1.0f}, 1.0f };
        int t = -1;
        int s = find_intersection(sc, r, &h, &t, &c);
        shape[k] = s; tri[k] = s >= 0 ? t : -1;
        float* o = hit11 + 11 * (size_t)k;
        memset(o, 0, 11 * sizeof(float));
        if (s >= 0) {
            o[0] = h.pos.x; o[1] = h.pos.y; o[2] = h.pos.z; o[3] = h.nrm.x; o[4] = h.nrm.y; o[5] = h.nrm.z;
            o[6] = h.dist; o[7] = h.color.x; o[8] = h.color.y; o[9] = h.color.z; o[10] = h.alpha;
        }
    }
    return RT_OK;
}

/* ---- primitive known-answer entry points (same signatures as the ref_kat_* of the harness) ---- */
void rt_oracle_kat_aabb(const float* rays, const float* boxes, int n, int* out, float* tmin)
{
    for (int i = 0; i < n; i++) {
        const float* q = rays + 7 * (size_t)i; const float* b = boxes + 6 * (size_t)i;
        ray_t r = { V(q[0], q[1], q[2]), V(q[3], q[4], q[5]), q[6] };
        float t = 0.0f;
        out[i] = slab_test(&r, b, b + 3, &t);
        tmin[i] = out[i] ? t : 0.0f;
    }
}

static void store7(float* o, int hit, v3 pos, v3 nrm, float dist)
{
    memset(o, 0, 7 * sizeof(float));
    if (hit) { o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = nrm.x; o[4] = nrm.y; o[5] = nrm.z; o[6] = dist; }
}

void rt_oracle_kat_triangle(const float* rays, const float* tris, int n, int* out, float* out7)
{
    for (int i = 0; i < n; i++) {
        const float* q = rays + 7 * (size_t)i; const float* t = tris + 9 * (size_t)i;
        ray_t r = { V(q[0], q[1], q[2]), V(q[3], q[4], q[5]), q[6] };
        v3 p0 = ld3(t), p1 = ld3(t + 3), p2 = ld3(t + 6);
        v3 nn = normalized(cross(sub(p1, p0), sub(p2, p0)));
        v3 pos = {0,0,0}, nrm = {0,0,0}; float dist = 0;
        out[i] = triangle_test(&r, p0, p1, p2, nn, &pos, &nrm, &dist);
        store7(out7 + 7 * (size_t)i, out[i], pos, nrm, dist);
    }
}

void rt_oracle_kat_sphere(const float* rays, const float* spheres, int n, int* out, float* out7)
{
    for (int i = 0; i < n; i++) {
        const float* q = rays + 7 * (size_t)i; const float* s = spheres + 4 * (size_t)i;
        ray_t r = { V(q[0], q[1], q[2]), V(q[3], q[4], q[5]), q[6] };
        v3 pos = {0,0,0}, nrm = {0,0,0}; float dist = 0;
        out[i] = sphere_test(&r, ld3(s), s[3], &pos, &nrm, &dist);
        store7(out7 + 7 * (size_t)i, out[i], pos, nrm, dist);
    }
}

void rt_oracle_kat_plane(const float* rays, const float* planes, int n, int* out, float* out7)
{
    for (int i = 0; i < n; i++) {
        const float* q = rays + 7 * (size_t)i; const float* s = planes + 6 * (size_t)i;
        ray_t r = { V(q[0], q[1], q[2]), V(q[3], q[4], q[5]), q[6] };
        v3 pos = {0,0,0}, nrm = {0,0,0}; float dist = 0;
        out[i] = plane_test(&r, ld3(s), ld3(s + 3), &pos, &nrm, &dist);
        store7(out7 + 7 * (size_t)i, out[i], pos, nrm, dist);
    }
}

void rt_oracle_kat_capsule(const float* rays, const float* caps, int n, int* out, float* out7)
{
    rt_scene_desc sc; memset(&sc, 0, sizeof sc);
    cnt_t c; memset(&c, 0, sizeof c);
    for (int i = 0; i < n; i++) {
        const float* q = rays + 7 * (size_t)i; const float* s = caps + 7 * (size_t)i;
        ray_t r = { V(q[0], q[1], q[2]), V(q[3], q[4], q[5]), q[6] };
        rt_shape sh; memset(&sh, 0, sizeof sh);
        sh.type = RT_SHAPE_CAPSULE; memcpy(sh.a, s, 12); memcpy(sh.b, s + 3, 12); sh.radius = s[6];
        hit_t h = { {0,0,0}, {0,0,0}, 0.0f, {1.0f, 1.0f, 1.0f}, 1.0f };
        out[i] = shape_test(&sc, &sh, &r, &h, NULL, &c);
        store7(out7 + 7 * (size_t)i, out[i], h.pos, h.nrm, h.dist);
    }
}

void rt_oracle_kat_qrsqrt(const float* x, int n, float* out) { for (int i = 0; i < n; i++) out[i] = q_rsqrt(x[i]); }

void rt_oracle_kat_barycentric(const float* in, int n, float* out)
{
    for (int i = 0; i < n; i++) {
        const float* q = in + 12 * (size_t)i;
        barycentric(ld3(q), ld3(q + 3), ld3(q + 6), ld3(q + 9), &out[3 * i], &out[3 * i + 1], &out[3 * i + 2]);
    }
}

void rt_oracle_kat_texture_sample(const float* rgba, int width, int height, const float* uv, int n, float* out4)
{
    rt_texture t = { rgba, width, height, NULL, 0, NULL };
    for (int i = 0; i < n; i++) texture_sample(&t, uv[2 * i], uv[2 * i + 1], out4 + 4 * (size_t)i);
}

void rt_oracle_kat_display(const float* rgb, int n, uint32_t* out)
{
    for (int i = 0; i < n; i++) out[i] = make_pixel(ld3(rgb + 3 * (size_t)i));
}
