// bvh_build.cpp — host build of the reference's "KdTree" (an AABB BVH, one triangle per leaf),
// emitted directly as the pre-order, escape-threaded node array the device traverses.
//
// The tree SHAPE must equal the reference's (KdNode::Build, KdTree.cpp:37-126) because the
// traversal order, and with it which of two near-equal hits survives, depends on it:
//   * node bounds  = min/max over the three corners of every triangle in the node (:42-47);
//   * one triangle -> leaf (:50-55);
//   * split axis   = widest axis of the node bounds with the reference's tie rules (:10-35):
//                    x only if strictly wider than y AND z, y only if strictly wider than z;
//   * split value  = mean over the node's triangles of the centroid (v0+v1+v2)/3, summed in
//                    list order in fp32 and divided by the count (:57-66);
//   * a triangle goes left iff centroid[axis] < mean[axis], list order preserved (:72-105);
//   * if one side would take everything, first half left / rest right (:108-113);
//   * children are visited Left then Right (:115-125 build, :138-148 traversal).
// Instead of per-node std::vector copies this works on one index array with a stable two-way
// split through a scratch buffer, precomputes each centroid once (same fp32 expression, so same
// bits), and appends nodes in visiting order so that left child = i+1 and the escape index is
// known when the subtree is finished.
#include "rt_host.hpp"

#include <float.h>
#include <math.h>

namespace rtb200 {

namespace {

struct Builder
{
    const RVec3* P;
    const int* Idx;                 // 3 per triangle
    std::vector<RVec3> centroid;
    std::vector<int> order;         // triangle ids, node ranges are contiguous
    std::vector<int> scratch;
    std::vector<rt_bvh_node>* nodes;
    std::vector<rt_tri>* tris;
    int max_depth = 0;

    void emit_leaf(int node, int t)
    {
        const RVec3& p0 = P[Idx[3 * t]];
        const RVec3& p1 = P[Idx[3 * t + 1]];
        const RVec3& p2 = P[Idx[3 * t + 2]];
        rt_tri r;
        r.p0[0] = p0.x; r.p0[1] = p0.y; r.p0[2] = p0.z; r.index = t;
        r.p1[0] = p1.x; r.p1[1] = p1.y; r.p1[2] = p1.z; r.pad0 = 0.0f;
        r.p2[0] = p2.x; r.p2[1] = p2.y; r.p2[2] = p2.z; r.pad1 = 0.0f;
        // face normal as RRay::TestIntersectionWithTriangle derives it per test (RRay.cpp:138-145):
        // cross(p1-p0, p2-p0), normalised only if |.|^2 >= FLT_EPSILON (RVector.h:169-183)
        float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
        float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
        float nx = ay * bz - az * by;
        float ny = az * bx - ax * bz;
        float nz = ax * by - ay * bx;
        float sqr_mag = nx * nx + ny * ny + nz * nz;
        if (!(fabsf(sqr_mag) < FLT_EPSILON))
        {
            float one_over_mag = 1.0f / sqrtf(sqr_mag);
            nx *= one_over_mag; ny *= one_over_mag; nz *= one_over_mag;
        }
        r.n[0] = nx; r.n[1] = ny; r.n[2] = nz; r.pad2 = 0.0f;
        (*nodes)[node].tri = (int32_t)tris->size();
        tris->push_back(r);
    }

    void build(int begin, int end, int depth)
    {
        if (depth > max_depth) max_depth = depth;
        const int node = (int)nodes->size();
        nodes->push_back(rt_bvh_node());
        const int count = end - begin;

        float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        for (int i = begin; i < end; i++)
        {
            const int t = order[i];
            for (int k = 0; k < 3; k++)
            {
                const RVec3& p = P[Idx[3 * t + k]];
                if (p.x < mn[0]) mn[0] = p.x;
                if (p.y < mn[1]) mn[1] = p.y;
                if (p.z < mn[2]) mn[2] = p.z;
                if (p.x > mx[0]) mx[0] = p.x;
                if (p.y > mx[1]) mx[1] = p.y;
                if (p.z > mx[2]) mx[2] = p.z;
            }
        }
        {
            rt_bvh_node& n = (*nodes)[node];
            n.bmin[0] = mn[0]; n.bmin[1] = mn[1]; n.bmin[2] = mn[2];
            n.bmax[0] = mx[0]; n.bmax[1] = mx[1]; n.bmax[2] = mx[2];
            n.tri = -1;
        }

        if (count == 1)
        {
            emit_leaf(node, order[begin]);
            (*nodes)[node].escape = (int32_t)nodes->size();
            return;
        }

        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        for (int i = begin; i < end; i++)
        {
            const RVec3& c = centroid[order[i]];
            sx += c.x; sy += c.y; sz += c.z;
        }
        const float fcount = (float)count;
        const float mean[3] = { sx / fcount, sy / fcount, sz / fcount };

        const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
        int axis;
        if (ex > ey) axis = (ex > ez) ? 0 : 2;
        else axis = (ey > ez) ? 1 : 2;

        int nl = 0, nr = 0;
        for (int i = begin; i < end; i++)
        {
            const int t = order[i];
            const RVec3& c = centroid[t];
            const float v = axis == 0 ? c.x : (axis == 1 ? c.y : c.z);
            if (v < mean[axis]) order[begin + nl++] = t;     // nl <= i - begin: never overwrites unread entries
            else scratch[nr++] = t;
        }
        int mid;
        if (nl == count || nr == count)
        {
            // everything fell on one side: the list is still in its original order
            if (nr == count) for (int i = 0; i < nr; i++) order[begin + i] = scratch[i];
            mid = begin + count / 2;
        }
        else
        {
            for (int i = 0; i < nr; i++) order[begin + nl + i] = scratch[i];
            mid = begin + nl;
        }

        build(begin, mid, depth + 1);
        build(mid, end, depth + 1);
        (*nodes)[node].escape = (int32_t)nodes->size();
    }
};

} // namespace

int BuildFlatBvh(const RVec3* Points, const int* Indices, int NumTriangles,
                 std::vector<rt_bvh_node>& nodes, std::vector<rt_tri>& tris)
{
    nodes.clear();
    tris.clear();
    if (NumTriangles <= 0) return 0;
    Builder b;
    b.P = Points; b.Idx = Indices; b.nodes = &nodes; b.tris = &tris;
    b.centroid.resize(NumTriangles);
    b.order.resize(NumTriangles);
    b.scratch.resize(NumTriangles);
    for (int t = 0; t < NumTriangles; t++)
    {
        const RVec3& v0 = Points[Indices[3 * t]];
        const RVec3& v1 = Points[Indices[3 * t + 1]];
        const RVec3& v2 = Points[Indices[3 * t + 2]];
        // (v0 + v1 + v2) / 3.0f, component-wise, left to right
        b.centroid[t] = RVec3(((v0.x + v1.x) + v2.x) / 3.0f, ((v0.y + v1.y) + v2.y) / 3.0f, ((v0.z + v1.z) + v2.z) / 3.0f);
        b.order[t] = t;
    }
    nodes.reserve((size_t)2 * NumTriangles);
    tris.reserve(NumTriangles);
    b.build(0, NumTriangles, 1);
    return b.max_depth;
}

} // namespace rtb200
