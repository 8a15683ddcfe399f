"""CPU study for the next round: how coherent are BOUNCE rays (round 1) if the queue is reordered?
A packet walk needs |union of its 32 rays' node sets| steps; the lane-per-walk kernel needs about
sum/18 (18 of 32 lanes per instruction, ncu).  Packets win when union < ~1.75 x the mean set size.
Node sets are computed with numpy on the flattened tree of this repo's host builder, for two bounds of the
culled walk: 'exact' (line test only, upper bound) and 'culled' (interval also clipped to [0, final hit
distance], lower bound).  No GPU, no product code path:   python tools/coherence_study.py [W] [H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import raytracerwin_b200 as rt
import scenes
from oracle.bindings import PortOracle

W = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
data = os.path.join(ROOT, "assets", "_ref", "Data")
sc = rt.Scene(scenes.c3_unitychan(data))
port = PortOracle()
nodes, tris, _ = sc.flat_mesh(0)
bmin = np.stack([nodes["bmin"][:, k] for k in range(3)], 1).astype(np.float32)
bmax = np.stack([nodes["bmax"][:, k] for k in range(3)], 1).astype(np.float32)
escape = nodes["escape"].astype(np.int64); leaf = nodes["tri"] >= 0
n = len(nodes)


def camera_rays(x0, y0, w, h):
    ys, xs = np.mgrid[y0:y0 + h, x0:x0 + w]
    dx = -(xs - W // 2).astype(np.float32) / np.float32(2 * W) * np.float32(W / H)
    dy = -(ys - H // 2).astype(np.float32) / np.float32(2 * H)
    d = np.stack([dx, dy, np.full_like(dx, -0.5)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((len(d), 7), np.float32); r[:, :3] = (0, 0, 7); r[:, 3:6] = d; r[:, 6] = 1000.0
    return r, xs.reshape(-1), ys.reshape(-1)


def node_sets(rays, dist_hi):
    """visited (ray, node) pairs of the cursor walk; dist_hi = None: line test only"""
    o, d = rays[:, :3].astype(np.float32), rays[:, 3:6].astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = (1.0 / d).astype(np.float32)
    cur = np.zeros(len(rays), np.int64)
    pairs_r, pairs_n, pairs_t = [], [], []
    ids = np.arange(len(rays))
    step = 0
    while True:
        act = cur < n
        if not act.any():
            break
        a = ids[act]; c = cur[a]
        t1 = (bmin[c] - o[a]) * inv[a]; t2 = (bmax[c] - o[a]) * inv[a]
        tmin = np.minimum(t1, t2).max(1); tmax = np.maximum(t1, t2).min(1)
        enter = tmax > tmin
        if dist_hi is not None:
            enter &= ~(tmax < -1e-3) & ~(tmin > dist_hi[a] * 1.0078125 + 1e-3)
        pairs_r.append(a); pairs_n.append(c); pairs_t.append(np.full(len(a), step, np.int64))
        cur[a] = np.where(enter & ~leaf[c], c + 1, escape[c])
        step += 1
    return np.concatenate(pairs_r), np.concatenate(pairs_n), np.concatenate(pairs_t)


def line_stats(order, pr, pn, pt, nrays, label):
    """lane-per-walk kernel in lock step (no refill): distinct 128-byte lines (4 nodes) a warp requests per step"""
    rank = np.empty(nrays, np.int64); rank[order] = np.arange(nrays)
    warp = rank[pr] // 32
    steps = int(pt.max()) + 1
    ws = warp * steps + pt
    lanes = np.bincount(ws)
    lines = np.bincount(np.unique(ws * (n // 4 + 1) + pn // 4) // (n // 4 + 1), minlength=len(lanes))
    nz = lanes > 0
    print(f"  {label:34s} lines per warp step {lines[nz].mean():5.1f} for {lanes[nz].mean():5.1f} active lanes  ({lines[nz].sum() / lanes[nz].sum():.2f} lines per node fetch)")


def packet_stats(order, pr, pn, nrays, label):
    rank = np.empty(nrays, np.int64); rank[order] = np.arange(nrays)
    pk = rank[pr] // 32
    single = np.bincount(pr, minlength=nrays)
    key = pk * n + pn
    union = np.bincount(np.unique(key) // n, minlength=(nrays + 31) // 32)
    full = nrays // 32
    print(f"  {label:34s} mean set {single.mean():7.1f}  mean union {union[:full].mean():8.1f}  union/set {union[:full].mean() / single.mean():5.2f}"
          f"  lanes per step {single.sum() / union.sum():5.1f}")


def morton(q):
    def part(v):
        v = v.astype(np.int64) & 1023
        v = (v | (v << 16)) & 0x30000FF; v = (v | (v << 8)) & 0x300F00F; v = (v | (v << 4)) & 0x30C30C3; v = (v | (v << 2)) & 0x9249249
        return v
    return part(q[:, 0]) | (part(q[:, 1]) << 1) | (part(q[:, 2]) << 2)


# a window of 8x4 pixel blocks over the figure
win = 384 if W >= 3840 else 192
cam, xs, ys = camera_rays(W // 2 - win // 2, H // 2 - win // 2 - H // 10, win, win)
block = (ys // 4) * (W // 8) + xs // 8
lane = (ys % 4) * 8 + xs % 8
order_blocks = np.lexsort((lane, block))
shape, tri, hit = port.trace_rays(sc.desc, cam)
print(f"{W}x{H}: {len(cam)} camera rays in a {win}x{win} window, {int((shape >= 0).sum())} hit the mesh; tree of {n} nodes")
for name, dh in (("exact", None), ("culled", np.where(shape >= 0, hit[:, 6], 1000.0).astype(np.float32))):
    pr, pn, pt = node_sets(cam, dh)
    print(f" camera rays, {name} walk")
    packet_stats(order_blocks, pr, pn, len(cam), "8x4 pixel blocks (round 0 today)")
    packet_stats(np.random.default_rng(0).permutation(len(cam)), pr, pn, len(cam), "random order")

# bounce rays: from the hit points into a random hemisphere direction (Diffuse, SurfaceMaterials.cpp:20-40)
h = shape >= 0
rng = np.random.default_rng(1)
v = rng.normal(size=(int(h.sum()), 3)).astype(np.float32); v /= np.linalg.norm(v, axis=1, keepdims=True)
nrm = hit[h, 3:6]; v = np.where((v * nrm).sum(1, keepdims=True) < 0, -v, v)
b = np.zeros((len(v), 7), np.float32); b[:, :3] = hit[h, :3] + nrm * 1e-3; b[:, 3:6] = v; b[:, 6] = 1000.0
bshape, _, bhit = port.trace_rays(sc.desc, b)
lo, hi = bmin[0], bmax[0]
cell = np.clip(((b[:, :3] - lo) / (hi - lo) * 1023).astype(np.int64), 0, 1023)
octant = (v[:, 0] > 0).astype(np.int64) | ((v[:, 1] > 0).astype(np.int64) << 1) | ((v[:, 2] > 0).astype(np.int64) << 2)
orders = {"queue order (same pixel blocks)": np.argsort(np.argsort(order_blocks)[h], kind="stable"),
          "sorted by origin cell (Morton)": np.argsort(morton(cell), kind="stable"),
          "direction octant, then origin cell": np.lexsort((morton(cell), octant)),
          "random order": rng.permutation(len(b))}
print(f"{len(b)} bounce rays, {int((bshape >= 0).sum())} hit again")
for name, dh in (("exact", None), ("culled", np.where(bshape >= 0, bhit[:, 6], 1000.0).astype(np.float32))):
    pr, pn, pt = node_sets(b, dh)
    print(f" bounce rays, {name} walk")
    for label, order in orders.items():
        packet_stats(order, pr, pn, len(b), label)
    for label, order in orders.items():
        line_stats(order, pr, pn, pt, len(b), label)
