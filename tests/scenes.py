"""Scene / config definitions shared by the tests, bench.py and the golden generator.

Specs are plain tuples understood by raytracerwin_b200.Scene and oracle.bindings.RefOracle alike
(the reference side builds them through RayTracerScene::AddShape and the SurfaceMaterial_* classes).
Configs follow SURVEY.md §8(d).
"""
import os

WHITE = (1.0, 1.0, 1.0)


def c1_torusknot(data):
    """C1: TorusKnot alone, Diffuse(1,1,1); Whitted primary + one shadow ray to GSceneLights[0]."""
    return [("mesh", os.path.join(data, "TorusKnot.obj"), ("diffuse", WHITE))]


def c2_monkey(data):
    """C2: BlenderMonkey, Reflective(0.8) fuzz 0 over a reflective ground plane (RayTracerProgram.cpp:512)."""
    return [("plane", (0.0, 1.0, 0.0), (0.0, -2.0, 0.0), ("reflective", (0.8, 0.8, 0.8), 0.0)),
            ("mesh", os.path.join(data, "BlenderMonkey.obj"), ("reflective", (0.8, 0.8, 0.8), 0.0))]


def c2_monkey_null(data):
    """C2 'refractive' stand-in: SurfaceMaterial_Null pass-through (the reference has no refraction)."""
    return [("plane", (0.0, 1.0, 0.0), (0.0, -2.0, 0.0), ("reflective", (0.8, 0.8, 0.8), 0.0)),
            ("mesh", os.path.join(data, "BlenderMonkey.obj"),
             ("combine", ("null",), ("emissive", (0.05, 0.02, 0.0))))]


def c3_unitychan(data):
    """C3/C4: unitychan as shipped (RayTracerProgram.cpp:546-551): Blend(Reflective(1;0.2), Diffuse(1), 1.0)."""
    return [("mesh", os.path.join(data, "unitychan.obj"),
             ("blend", ("reflective", WHITE, 0.2), ("diffuse", WHITE), 1.0))]


def default_scene(data):
    """RayTracerProgram::SetupScene (RayTracerProgram.cpp:467-552) as a spec list."""
    gold = (0.95, 0.75, 0.1)
    half = tuple(__import__("numpy").float32(g) * __import__("numpy").float32(0.5) for g in gold)
    return [
        ("sphere", (1.5, 2.5, -2.0), 0.9, ("blend", ("reflective", WHITE, 0.0), ("diffuse", (1.0, 0.5, 0.1)), 0.5)),
        ("sphere", (-1.5, -0.5, -3.0), 0.5, ("diffuse", (0.1, 1.0, 0.2))),
        ("sphere", (0.8, -1.5, -1.0), 0.5, ("blend", ("reflective", WHITE, 0.0), ("diffuse", (0.5, 0.0, 0.2)), 0.5)),
        ("sphere", (2.8, -1.2, -4.0), 1.5,
         ("combine", ("blend", ("reflective", gold, 0.0), ("diffuse", gold), 0.5), ("emissive", half))),
        ("capsule", (-1.5, -1.5, -1.5), (-2.0, -1.5, 0.0), 0.5,
         ("blend", ("reflective", (0.8, 0.75, 0.6), 0.2), ("diffuse", (0.25, 0.75, 0.6)), 0.2)),
        ("plane", (0.0, 1.0, 0.0), (0.0, -2.0, 0.0),
         ("blend", ("reflective", WHITE, 0.1), ("checker", WHITE, 5.0), 0.5)),
        ("mesh", os.path.join(data, "unitychan.obj"), ("blend", ("reflective", WHITE, 0.2), ("diffuse", WHITE), 1.0)),
    ]


def deterministic_mix(data):
    """Every deterministic material class on analytic shapes + a mesh (no RNG-dependent geometry)."""
    return [
        ("sphere", (1.2, 0.8, -1.0), 0.7, ("reflective", (0.9, 0.8, 0.7), 0.0)),
        ("sphere", (-1.4, 0.2, -0.5), 0.6, ("combine", ("reflective", (0.5, 0.5, 0.5), 0.0), ("emissive", (0.3, 0.1, 0.05)))),
        ("capsule", (-0.5, -1.2, 0.5), (0.8, -1.4, 1.0), 0.3, ("emissive", (0.8, 0.9, 0.4))),
        ("triangle", (-2.5, -1.0, -2.0), (2.5, -1.0, -2.0), (0.0, 2.5, -2.5), ("reflective", (0.7, 0.7, 0.9), 0.0)),
        ("plane", (0.0, 1.0, 0.0), (0.0, -2.0, 0.0), ("reflective", (0.8, 0.8, 0.8), 0.0)),
        ("sphere", (0.0, -0.3, 2.5), 0.4, ("null",)),
        ("mesh", os.path.join(data, "TorusKnot.obj"), ("reflective", (0.8, 0.6, 0.4), 0.0)),
    ]
