"""ctypes bindings for the two CHECKERS (test infrastructure — never imported by the product):

  RefOracle   oracle/_ref/libref_oracle.so — the UNMODIFIED reference compiled from /root/reference
              (oracle/Makefile `make ref`); `counting=True` loads the ray-counting twin.
  PortOracle  oracle/librt_oracle.so — the plain-C restatement (oracle/rt_oracle.c), which consumes
              the same rt_scene_desc / rt_render_params as the CUDA path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libref_oracle.so")
REF_COUNT_LIB = os.path.join(HERE, "_ref", "libref_oracle_count.so")
PORT_LIB = os.path.join(HERE, "librt_oracle.so")

MODE_PATH, MODE_PREVIEW, MODE_WHITTED = 0, 1, 2

VP, I, F, U32 = C.c_void_p, C.c_int, C.c_float, C.c_uint32


def ref_available():
    return os.path.exists(REF_LIB)


def port_available():
    return os.path.exists(PORT_LIB)


class RefOracle:
    """The reference's own CPU implementation behind oracle/ref_harness.cpp."""

    def __init__(self, counting=False):
        path = REF_COUNT_LIB if counting else REF_LIB
        L = self.L = C.CDLL(path)
        L.ref_scene_new.restype = VP
        L.ref_scene_free.argtypes = [VP]
        for n, a in (("diffuse", [F] * 3), ("checker", [F] * 4), ("reflective", [F] * 4), ("emissive", [F] * 3),
                     ("null", []), ("blend", [VP, VP, F]), ("combine", [VP, VP])):
            fn = getattr(L, "ref_mat_" + n)
            fn.restype, fn.argtypes = VP, a
        L.ref_add_sphere.argtypes = [VP] + [F] * 4 + [VP]
        L.ref_add_plane.argtypes = [VP] + [F] * 6 + [VP]
        L.ref_add_capsule.argtypes = [VP] + [F] * 7 + [VP]
        L.ref_add_triangle.argtypes = [VP, VP, VP]
        L.ref_add_mesh.argtypes = [VP, C.c_char_p, VP]
        L.ref_default_scene.restype = VP
        L.ref_default_scene.argtypes = [C.c_char_p]
        L.ref_num_shapes.argtypes = [VP]
        L.ref_shape_bounds.argtypes = [VP, I, VP]
        L.ref_mesh_counts.argtypes = [VP, I, VP]
        L.ref_mesh_dump.argtypes = [VP, I] + [VP] * 7
        L.ref_mesh_texture_info.argtypes = [VP, I, I, VP]
        L.ref_mesh_texture_pixels.argtypes = [VP, I, I, VP]
        L.ref_mesh_bvh_dump.argtypes = [VP, I] + [VP] * 4
        L.ref_trace_primary.argtypes = [VP, I, I, I, I] + [VP] * 5
        L.ref_trace_rays.argtypes = [VP, VP, I, VP, VP, VP]
        L.ref_render.argtypes = [VP] + [I] * 9 + [U32, I, VP, VP, VP]
        L.ref_init_unit_vectors.argtypes = [U32]
        L.ref_unit_vector_table.restype = C.POINTER(C.c_float)
        L.ref_unit_vector_table.argtypes = [C.POINTER(C.c_uint)]
        L.ref_set_libc_rand.argtypes = [I]
        L.ref_kat_texture_sample.argtypes = [VP, I, I, VP, I, VP]
        L.ref_init()

    # -- scene construction through the reference's API
    def _mat(self, spec):
        L = self.L
        if spec is None:
            return None
        k = spec[0]
        if k == "diffuse":
            return L.ref_mat_diffuse(*map(float, spec[1]))
        if k == "checker":
            return L.ref_mat_checker(*map(float, spec[1]), float(spec[2]))
        if k == "reflective":
            return L.ref_mat_reflective(*map(float, spec[1]), float(spec[2]))
        if k == "emissive":
            return L.ref_mat_emissive(*map(float, spec[1]))
        if k == "blend":
            return L.ref_mat_blend(self._mat(spec[1]), self._mat(spec[2]), float(spec[3]))
        if k == "combine":
            return L.ref_mat_combine(self._mat(spec[1]), self._mat(spec[2]))
        if k == "null":
            return L.ref_mat_null()
        raise ValueError(k)

    def build_scene(self, shapes):
        L = self.L
        s = L.ref_scene_new()
        for sp in shapes:
            k, m = sp[0], self._mat(sp[-1])
            if k == "sphere":
                L.ref_add_sphere(s, *map(float, sp[1]), float(sp[2]), m)
            elif k == "plane":
                L.ref_add_plane(s, *map(float, sp[1]), *map(float, sp[2]), m)
            elif k == "capsule":
                L.ref_add_capsule(s, *map(float, sp[1]), *map(float, sp[2]), float(sp[3]), m)
            elif k == "triangle":
                p = np.asarray([x for v in sp[1:4] for x in v], np.float32)
                L.ref_add_triangle(s, p.ctypes.data, m)
            elif k == "mesh":
                L.ref_add_mesh(s, os.fsencode(sp[1]), m)
            else:
                raise ValueError(k)
        return s

    def default_scene(self, data_parent):
        """RayTracerProgram::SetupScene run from `data_parent` (the directory that holds Data/)."""
        s = self.L.ref_default_scene(os.fsencode(data_parent))
        if not s:
            raise RuntimeError("ref_default_scene failed")
        return s

    def free_scene(self, s):
        self.L.ref_scene_free(s)

    def init_unit_vectors(self, seed=0):
        self.L.ref_init_unit_vectors(seed)

    def unit_vector_table(self):
        n = C.c_uint()
        p = self.L.ref_unit_vector_table(C.byref(n))
        return np.ctypeslib.as_array(p, (n.value, 3))

    def set_libc_rand(self, on):
        self.L.ref_set_libc_rand(1 if on else 0)

    def hardware_threads(self):
        return self.L.ref_hardware_threads()

    # -- dumps
    def shape_bounds(self, s, i):
        out = np.zeros(6, np.float32)
        has = self.L.ref_shape_bounds(s, i, out.ctypes.data)
        return out, has

    def mesh_counts(self, s, shape):
        out = (C.c_int * 7)()
        if self.L.ref_mesh_counts(s, shape, out) != 0:
            raise RuntimeError("not a mesh")
        return list(out)

    def mesh_dump(self, s, shape):
        npts, ntex, nnrm, ntri, _, _, _ = self.mesh_counts(s, shape)
        d = dict(points=np.zeros((npts, 3), np.float32), texcoords=np.zeros((ntex, 3), np.float32),
                 normals=np.zeros((nnrm, 3), np.float32), pidx=np.zeros(3 * ntri, np.int32),
                 tidx=np.zeros(3 * ntri, np.int32), nidx=np.zeros(3 * ntri, np.int32), matid=np.zeros(ntri, np.int32))
        self.L.ref_mesh_dump(s, shape, *[d[k].ctypes.data for k in ("points", "texcoords", "normals", "pidx", "tidx", "nidx", "matid")])
        return d

    def mesh_texture(self, s, shape, slot):
        wh = (C.c_int * 2)()
        if self.L.ref_mesh_texture_info(s, shape, slot, wh) != 0 or wh[0] == 0:
            return None
        px = np.zeros((wh[1], wh[0], 4), np.float32)
        self.L.ref_mesh_texture_pixels(s, shape, slot, px.ctypes.data)
        return px

    def mesh_bvh(self, s, shape):
        n = self.mesh_counts(s, shape)[5]
        bounds = np.zeros((n, 6), np.float32)
        escape = np.zeros(n, np.int32)
        tri = np.zeros(n, np.int32)
        verts = np.zeros((n, 3), np.int32)
        got = self.L.ref_mesh_bvh_dump(s, shape, bounds.ctypes.data, escape.ctypes.data, tri.ctypes.data, verts.ctypes.data)
        assert got == n
        return bounds, escape, tri, verts

    # -- tracing
    def trace_primary(self, s, W, H, start=0, end=None, want_hit=False):
        end = W * H - 1 if end is None else end
        n = end - start + 1
        shape = np.zeros(n, np.int32)
        tri = np.zeros(n, np.int32)
        dist = np.zeros(n, np.float32)
        hit = np.zeros((n, 11), np.float32) if want_hit else None
        cnt = np.zeros(3, np.uint64)
        self.L.ref_trace_primary(s, W, H, start, end, shape.ctypes.data, tri.ctypes.data, dist.ctypes.data,
                                 None if hit is None else hit.ctypes.data, cnt.ctypes.data)
        return dict(shape=shape, tri=tri, dist=dist, hit=hit, node_tests=int(cnt[0]), tri_tests=int(cnt[1]), mismatches=int(cnt[2]))

    def trace_rays(self, s, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        n = len(rays)
        shape = np.zeros(n, np.int32)
        tri = np.zeros(n, np.int32)
        hit = np.zeros((n, 11), np.float32)
        mism = self.L.ref_trace_rays(s, rays.ctypes.data, n, shape.ctypes.data, tri.ctypes.data, hit.ctypes.data)
        assert mism == 0
        return shape, tri, hit

    def render(self, s, W, H, mode=MODE_PATH, max_bounce=10, pass_begin=0, pass_count=1, antialias=1, seed=0,
               nthreads=1, start=0, end=None, accum=None, want_display=False):
        end = W * H - 1 if end is None else end
        accum = np.zeros((H, W, 4), np.float32) if accum is None else accum
        display = np.zeros((H, W), np.uint32) if want_display else None
        stats = np.zeros(3, np.float64)
        self.L.ref_render(s, W, H, start, end, mode, max_bounce, pass_begin, pass_count, antialias, seed, nthreads,
                          accum.ctypes.data, None if display is None else display.ctypes.data, stats.ctypes.data)
        return dict(accum=accum, display=display, seconds=float(stats[0]), rays=int(stats[1]), shadow_rays=int(stats[2]))

    # -- primitive known-answer tests
    def _kat7(self, fn, rays, prim, width):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        prim = np.ascontiguousarray(prim, np.float32).reshape(-1, width)
        n = len(rays)
        out = np.zeros(n, np.int32)
        o7 = np.zeros((n, 7), np.float32)
        fn.argtypes = [VP, VP, I, VP, VP]
        fn(rays.ctypes.data, prim.ctypes.data, n, out.ctypes.data, o7.ctypes.data)
        return out, o7

    def kat_aabb(self, rays, boxes):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        out = np.zeros(len(rays), np.int32)
        tmin = np.zeros(len(rays), np.float32)
        self.L.ref_kat_aabb.argtypes = [VP, VP, I, VP, VP]
        self.L.ref_kat_aabb(rays.ctypes.data, boxes.ctypes.data, len(rays), out.ctypes.data, tmin.ctypes.data)
        return out, tmin

    def kat_triangle(self, rays, tris):
        return self._kat7(self.L.ref_kat_triangle, rays, tris, 9)

    def kat_sphere(self, rays, spheres):
        return self._kat7(self.L.ref_kat_sphere, rays, spheres, 4)

    def kat_plane(self, rays, planes):
        return self._kat7(self.L.ref_kat_plane, rays, planes, 6)

    def kat_capsule(self, rays, caps):
        return self._kat7(self.L.ref_kat_capsule, rays, caps, 7)

    def kat_qrsqrt(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(x)
        self.L.ref_kat_qrsqrt.argtypes = [VP, I, VP]
        self.L.ref_kat_qrsqrt(x.ctypes.data, x.size, out.ctypes.data)
        return out

    def kat_barycentric(self, pabc):
        pabc = np.ascontiguousarray(pabc, np.float32).reshape(-1, 12)
        out = np.zeros((len(pabc), 3), np.float32)
        self.L.ref_kat_barycentric.argtypes = [VP, I, VP]
        self.L.ref_kat_barycentric(pabc.ctypes.data, len(pabc), out.ctypes.data)
        return out

    def kat_texture_sample(self, s, shape, slot, uv):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((len(uv), 4), np.float32)
        rc = self.L.ref_kat_texture_sample(s, shape, slot, uv.ctypes.data, len(uv), out.ctypes.data)
        assert rc == 0
        return out

    def kat_display(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
        out = np.zeros(len(rgb), np.uint32)
        self.L.ref_kat_display.argtypes = [VP, I, VP]
        self.L.ref_kat_display(rgb.ctypes.data, len(rgb), out.ctypes.data)
        return out

    def light0(self):
        out = np.zeros(7, np.float32)
        self.L.ref_light0.argtypes = [VP]
        self.L.ref_light0(out.ctypes.data)
        return out


class PortOracle:
    """oracle/rt_oracle.c: same inputs as the CUDA path (rt_scene_desc*, rt_render_params*)."""

    def __init__(self):
        L = self.L = C.CDLL(PORT_LIB)
        L.rt_oracle_render_ex.argtypes = [VP, VP, I, VP, VP, VP, VP, VP, VP]
        L.rt_oracle_trace_rays.argtypes = [VP, VP, I, VP, VP, VP]

    def render(self, desc, params, nthreads=1, accum=None, want_display=False, want_primary=False):
        """desc: POINTER(rt_scene_desc); params: rt_render_params."""
        from raytracerwin_b200._abi import rt_counters
        W, H = params.width, params.height
        accum = np.zeros((H, W, 4), np.float32) if accum is None else accum
        display = np.zeros((H, W), np.uint32) if want_display else None
        ids = np.full((H, W, 2), -1, np.int32) if want_primary else None
        dist = np.zeros((H, W), np.float32) if want_primary else None
        # RT_MODE_PREVIEW (1) leaves accum alone, like the reference; the pass colour comes back as "preview"
        preview = np.zeros((H, W, 4), np.float32) if params.mode == 1 else None
        cnt = rt_counters()
        ptr = lambda a: None if a is None else a.ctypes.data
        rc = self.L.rt_oracle_render_ex(C.cast(desc, VP), C.cast(C.pointer(params), VP), nthreads, ptr(accum), ptr(display),
                                        ptr(ids), ptr(dist), C.cast(C.pointer(cnt), VP), ptr(preview))
        if rc != 0:
            raise RuntimeError(f"rt_oracle_render failed: {rc}")
        return dict(accum=accum, display=display, ids=ids, dist=dist, counters=cnt.as_dict(), preview=preview)

    def trace_rays(self, desc, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        n = len(rays)
        shape = np.zeros(n, np.int32)
        tri = np.zeros(n, np.int32)
        hit = np.zeros((n, 11), np.float32)
        self.L.rt_oracle_trace_rays(C.cast(desc, VP), rays.ctypes.data, n, shape.ctypes.data, tri.ctypes.data, hit.ctypes.data)
        return shape, tri, hit

    def _kat7(self, name, rays, prim, width):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        prim = np.ascontiguousarray(prim, np.float32).reshape(-1, width)
        n = len(rays)
        out = np.zeros(n, np.int32)
        o7 = np.zeros((n, 7), np.float32)
        fn = getattr(self.L, name)
        fn.argtypes = [VP, VP, I, VP, VP]
        fn(rays.ctypes.data, prim.ctypes.data, n, out.ctypes.data, o7.ctypes.data)
        return out, o7

    def kat_aabb(self, rays, boxes):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        out = np.zeros(len(rays), np.int32)
        tmin = np.zeros(len(rays), np.float32)
        self.L.rt_oracle_kat_aabb.argtypes = [VP, VP, I, VP, VP]
        self.L.rt_oracle_kat_aabb(rays.ctypes.data, boxes.ctypes.data, len(rays), out.ctypes.data, tmin.ctypes.data)
        return out, tmin

    def kat_triangle(self, rays, tris):
        return self._kat7("rt_oracle_kat_triangle", rays, tris, 9)

    def kat_sphere(self, rays, spheres):
        return self._kat7("rt_oracle_kat_sphere", rays, spheres, 4)

    def kat_plane(self, rays, planes):
        return self._kat7("rt_oracle_kat_plane", rays, planes, 6)

    def kat_capsule(self, rays, caps):
        return self._kat7("rt_oracle_kat_capsule", rays, caps, 7)

    def kat_qrsqrt(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(x)
        self.L.rt_oracle_kat_qrsqrt.argtypes = [VP, I, VP]
        self.L.rt_oracle_kat_qrsqrt(x.ctypes.data, x.size, out.ctypes.data)
        return out

    def kat_barycentric(self, pabc):
        pabc = np.ascontiguousarray(pabc, np.float32).reshape(-1, 12)
        out = np.zeros((len(pabc), 3), np.float32)
        self.L.rt_oracle_kat_barycentric.argtypes = [VP, I, VP]
        self.L.rt_oracle_kat_barycentric(pabc.ctypes.data, len(pabc), out.ctypes.data)
        return out

    def kat_texture_sample(self, rgba, uv):
        rgba = np.ascontiguousarray(rgba, np.float32)
        h, w = rgba.shape[:2]
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((len(uv), 4), np.float32)
        self.L.rt_oracle_kat_texture_sample.argtypes = [VP, I, I, VP, I, VP]
        self.L.rt_oracle_kat_texture_sample(rgba.ctypes.data, w, h, uv.ctypes.data, len(uv), out.ctypes.data)
        return out

    def kat_display(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
        out = np.zeros(len(rgb), np.uint32)
        self.L.rt_oracle_kat_display.argtypes = [VP, I, VP]
        self.L.rt_oracle_kat_display(rgb.ctypes.data, len(rgb), out.ctypes.data)
        return out
