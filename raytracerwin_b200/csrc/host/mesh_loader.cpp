// mesh_loader.cpp — OBJ / MTL loading for RMeshShape and the flattening of a mesh into
// device-ready records.  Accepts the same subset of the formats as the reference's loader
// (RMeshShape::RMeshShape, MeshShape.cpp:65-278) and resolves them the same way, because the
// triangle numbering (TriangleData::Index), the material-id -> texture mapping and the v-flip
// conventions are all observable in the rendered image:
//   * keyword = text before the first space (:23-35); recognised: v, vt, vn, f, usemtl;
//   * `f` lines are split on single spaces (:50-62; a trailing space adds no token), 3 tokens make
//     one triangle, 4 make two (0,1,2)(0,2,3) (:133-144), any other count is ignored;
//   * each corner token is v/vt/vn, 1-based (:151-158);
//   * `usemtl NAME` switches the current material id, ids are assigned in first-use order (:168-183);
//   * the .mtl file is the .obj path with its first ".obj" replaced by ".mtl" (:203-207); only
//     `newmtl` and `map_Kd` are read; texture paths are relative to the .mtl, with the escaped
//     backslash pairs the exporter writes turned into '/' (:256-264);
//   * missing files are reported and leave an empty mesh / untextured material, never abort.
// Difference from the reference, on purpose: corner indices are validated.  A `v//vn` or `v`
// token parses to index -1 in the reference and is then used to index a vector (UB); here the
// mesh is rejected with an error text instead.
#include "rt_host.hpp"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <iterator>
#include <thread>
#include <fstream>

namespace rtb200 {

namespace {

// Splits on every single `delim`; empty pieces are kept except a final empty one.
std::vector<std::string> SplitKeepEmpty(const std::string& s, char delim)
{
    std::vector<std::string> out;
    size_t start = 0;
    while (start < s.size())
    {
        size_t p = s.find(delim, start);
        if (p == std::string::npos) { out.push_back(s.substr(start)); return out; }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}

// n-th '/'-separated integer of a corner token, 0 when absent/unparsable (the reference's
// stream extraction yields 0 on failure and keeps it for the remaining fields).
int NthIndex(const std::string& tok, int n)
{
    const char* p = tok.c_str();
    int value = 0;
    for (int i = 0; i <= n; i++)
    {
        while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') p++;
        char* end = nullptr;
        long v = strtol(p, &end, 10);
        if (end == p) return 0;
        value = (int)v;
        p = end;
        while (*p == ' ' || *p == '\t' || *p == '\r') p++;
        if (*p) p++;            // the separator character
    }
    return value;
}

// whitespace-separated floats after the keyword; missing values stay 0
int ParseFloats(const std::string& line, size_t from, float* out, int want)
{
    const char* p = line.c_str() + from;
    int got = 0;
    while (got < want)
    {
        char* end = nullptr;
        float v = strtof(p, &end);
        if (end == p) break;
        out[got++] = v;
        p = end;
    }
    return got;
}

std::string SecondWord(const std::string& line)
{
    size_t i = 0;
    while (i < line.size() && !isspace((unsigned char)line[i])) i++;
    while (i < line.size() && isspace((unsigned char)line[i])) i++;
    size_t j = i;
    while (j < line.size() && !isspace((unsigned char)line[j])) j++;
    return line.substr(i, j - i);
}

bool OpenWithParentFallback(const std::string& name, std::ifstream& f, std::string& resolved)
{
    resolved = name;
    f.open(resolved.c_str());
    for (int i = 0; i < 2 && !f.is_open(); i++)
    {
        resolved = std::string("../") + resolved;
        f.clear();
        f.open(resolved.c_str());
    }
    return f.is_open();
}

// One line-aligned piece of an OBJ file, parsed on its own (see RMeshShape::RMeshShape).
struct ObjChunk
{
    std::vector<RVec3> points, texcoords, normals;
    std::vector<int> pidx, tidx, nidx;
    std::vector<int> face_event;                 // per triangle: index into material_names, -1 = inherited
    std::vector<std::string> material_names;     // the chunk's usemtl lines, in order
};

// The line handlers of MeshShape.cpp:96-184 (v / vt / vn / f with quad split / usemtl).
void ParseObjChunk(const char* b, const char* e, ObjChunk& out)
{
    int current_event = -1;
    std::string Line;
    while (b < e)
    {
        const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
        const char* le = nl ? nl : e;
        Line.assign(b, le);
        b = nl ? nl + 1 : e;
        const size_t sp = Line.find(' ');
        const std::string key = sp == std::string::npos ? Line : Line.substr(0, sp);
        if (key == "v")
        {
            float p[3] = { 0, 0, 0 };
            ParseFloats(Line, 1, p, 3);
            out.points.push_back(RVec3(p));
        }
        else if (key == "vt")
        {
            float t[3] = { 0, 0, 0 };
            ParseFloats(Line, 2, t, 2);
            out.texcoords.push_back(RVec3(t[0], t[1], 0.0f));
        }
        else if (key == "vn")
        {
            float n[3] = { 0, 0, 0 };
            ParseFloats(Line, 2, n, 3);
            out.normals.push_back(RVec3(n));
        }
        else if (key == "f")
        {
            std::vector<std::string> tok = SplitKeepEmpty(Line, ' ');
            const int corners = (int)tok.size() - 1;
            static const int tri_order[3] = { 0, 1, 2 };
            static const int quad_order[6] = { 0, 1, 2, 0, 2, 3 };
            const int* order = corners == 3 ? tri_order : (corners == 4 ? quad_order : nullptr);
            const int n = corners == 3 ? 3 : 6;
            if (order)
            {
                for (int i = 0; i < n; i++)
                {
                    const std::string& c = tok[order[i] + 1];
                    out.pidx.push_back(NthIndex(c, 0) - 1);
                    out.tidx.push_back(NthIndex(c, 1) - 1);
                    out.nidx.push_back(NthIndex(c, 2) - 1);
                    if (i % 3 == 0) out.face_event.push_back(current_event);
                }
            }
        }
        else if (key == "usemtl")
        {
            std::vector<std::string> tok = SplitKeepEmpty(Line, ' ');
            out.material_names.push_back(tok.size() > 1 ? tok[1] : std::string());
            current_event = (int)out.material_names.size() - 1;
        }
    }
}

} // namespace

RMeshShape::RMeshShape(const std::string& Filename)
{
    std::ifstream in;
    std::string MeshFilename;
    if (!OpenWithParentFallback(Filename, in, MeshFilename))
    {
        ErrorText = "Error - RMeshShape: Unable to open " + Filename;
        printf("%s!\n", ErrorText.c_str());
        return;
    }

    // The text is parsed in line-aligned chunks on all host threads (the reference's
    // stringstream-per-token loop, MeshShape.cpp:96-184, is most of a big scene's load time) and the chunks
    // are merged in file order, which reproduces the sequential result exactly: every line is handled by
    // the same code whichever thread sees it, and the only state that crosses lines — the current
    // material (MeshShape.cpp:160-184) and the order-dependent bounds — is resolved during the merge.
    std::vector<std::string> MaterialNameList;
    int CurrentMaterialIdx = -1;
    std::string Line;
    {
        std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        unsigned nthreads = std::thread::hardware_concurrency();
        if (nthreads < 1) nthreads = 1;
        if (nthreads > 64) nthreads = 64;
        // (RT_OBJ_CHUNK_MIN lowers the size below which a file is parsed in one piece: tests force the chunked path)
        const size_t chunk_min = getenv("RT_OBJ_CHUNK_MIN") ? (size_t)atoll(getenv("RT_OBJ_CHUNK_MIN")) : (size_t)(4u << 20);
        if (text.size() < chunk_min) nthreads = 1;
        else if (nthreads < 2 && getenv("RT_OBJ_CHUNK_MIN")) nthreads = 3;
        std::vector<size_t> cut(nthreads + 1, text.size());
        cut[0] = 0;
        for (unsigned k = 1; k < nthreads; k++)
        {
            size_t pos = text.size() / nthreads * k;
            if (pos < cut[k - 1]) pos = cut[k - 1];
            const size_t nl = text.find('\n', pos);
            cut[k] = nl == std::string::npos ? text.size() : nl + 1;
        }
        std::vector<ObjChunk> chunks(nthreads);
        auto work = [&](unsigned k) { ParseObjChunk(text.data() + cut[k], text.data() + cut[k + 1], chunks[k]); };
        if (nthreads == 1) work(0);
        else
        {
            std::vector<std::thread> th;
            for (unsigned k = 0; k < nthreads; k++) th.emplace_back(work, k);
            for (auto& t : th) t.join();
        }
        size_t np = 0, nt = 0, nn = 0, nc = 0, nf = 0;
        for (const ObjChunk& c : chunks) { np += c.points.size(); nt += c.texcoords.size(); nn += c.normals.size(); nc += c.pidx.size(); nf += c.face_event.size(); }
        Points.reserve(np); Texcoords.reserve(nt); Normals.reserve(nn);
        PointIndices.reserve(nc); TexcoordIndices.reserve(nc); NormalIndices.reserve(nc); PolyMaterialId.reserve(nf);
        for (const ObjChunk& c : chunks)
        {
            for (const RVec3& v : c.points) { Points.push_back(v); Aabb.Expand(v); }
            Texcoords.insert(Texcoords.end(), c.texcoords.begin(), c.texcoords.end());
            Normals.insert(Normals.end(), c.normals.begin(), c.normals.end());
            PointIndices.insert(PointIndices.end(), c.pidx.begin(), c.pidx.end());
            TexcoordIndices.insert(TexcoordIndices.end(), c.tidx.begin(), c.tidx.end());
            NormalIndices.insert(NormalIndices.end(), c.nidx.begin(), c.nidx.end());
            // usemtl events of the chunk, in order, against the global name list
            std::vector<int> resolved(c.material_names.size());
            for (size_t e = 0; e < c.material_names.size(); e++)
            {
                const std::string& name = c.material_names[e];
                auto it = std::find(MaterialNameList.begin(), MaterialNameList.end(), name);
                if (it == MaterialNameList.end())
                {
                    MaterialNameList.push_back(name);
                    resolved[e] = (int)MaterialNameList.size() - 1;
                }
                else resolved[e] = (int)(it - MaterialNameList.begin());
            }
            for (int ev : c.face_event) PolyMaterialId.push_back(ev < 0 ? CurrentMaterialIdx : resolved[ev]);
            if (!resolved.empty()) CurrentMaterialIdx = resolved.back();
        }
    }
    in.close();

    // validate what the reference would index blindly
    const int ntri = (int)PointIndices.size() / 3;
    for (size_t i = 0; i < PointIndices.size(); i++)
    {
        if (PointIndices[i] < 0 || PointIndices[i] >= (int)Points.size() ||
            NormalIndices[i] < 0 || NormalIndices[i] >= (int)Normals.size() ||
            TexcoordIndices[i] < 0 || TexcoordIndices[i] >= (int)Texcoords.size())
        {
            char msg[256];
            snprintf(msg, sizeof msg, "RMeshShape: %s: face corner %zu has a missing or out-of-range v/vt/vn index (%d/%d/%d)",
                     Filename.c_str(), i, PointIndices[i] + 1, TexcoordIndices[i] + 1, NormalIndices[i] + 1);
            ErrorText = msg;
            printf("%s\n", msg);
            Points.clear(); PointIndices.clear(); TexcoordIndices.clear(); NormalIndices.clear(); PolyMaterialId.clear();
            return;
        }
    }

    // materials
    std::string MaterialFilename = MeshFilename;
    size_t ext = MaterialFilename.find(".obj");
    if (ext != std::string::npos)
    {
        MaterialFilename.replace(ext, 4, ".mtl");
        std::ifstream mtl(MaterialFilename.c_str());
        if (mtl.is_open())
        {
            std::string BasePath;
            size_t slash = MaterialFilename.find_last_of("\\/");
            if (slash != std::string::npos) BasePath = MaterialFilename.substr(0, slash + 1);
            Textures.resize(PolyMaterialId.size());      // sized by triangle count, indexed by material id (:220)
            CurrentMaterialIdx = -1;
            std::vector<std::pair<int, std::string> > TextureJobs;      // (material id, path) in file order
            while (std::getline(mtl, Line))
            {
                const size_t sp = Line.find(' ');
                const std::string key = sp == std::string::npos ? Line : Line.substr(0, sp);
                if (key == "newmtl")
                {
                    const std::string name = SecondWord(Line);
                    auto it = std::find(MaterialNameList.begin(), MaterialNameList.end(), name);
                    CurrentMaterialIdx = it == MaterialNameList.end() ? -1 : (int)(it - MaterialNameList.begin());
                }
                else if (key == "map_Kd")
                {
                    if (CurrentMaterialIdx == -1) continue;       // material not used by the mesh
                    std::string TexturePath = BasePath + SecondWord(Line);
                    size_t bs;
                    while ((bs = TexturePath.find("\\\\")) != std::string::npos) TexturePath.replace(bs, 2, "/");
                    if (CurrentMaterialIdx < (int)Textures.size())
                        TextureJobs.emplace_back(CurrentMaterialIdx, TexturePath);
                }
            }
            // The reference decodes the PNGs one after another while it reads the .mtl; the decodes are
            // independent, so they run on separate threads here and are assigned in file order afterwards
            // (a material named twice keeps its last texture, as there).
            std::vector<std::unique_ptr<RTexture> > Decoded(TextureJobs.size());
            {
                unsigned nthreads = std::thread::hardware_concurrency();
                if (nthreads < 1) nthreads = 1;
                if (nthreads > TextureJobs.size()) nthreads = (unsigned)TextureJobs.size();
                std::vector<std::thread> th;
                for (unsigned t = 0; t < nthreads; t++)
                    th.emplace_back([&, t]() {
                        for (size_t k = t; k < TextureJobs.size(); k += nthreads)
                            Decoded[k] = RTexture::LoadTexturePNG(TextureJobs[k].second);
                    });
                for (auto& x : th) x.join();
            }
            for (size_t k = 0; k < TextureJobs.size(); k++)
                Textures[TextureJobs[k].first] = std::move(Decoded[k]);
        }
    }

    (void)ntri;
    BuildSpatial();
    Loaded = !BuildFailed;
}

RMeshShape::RMeshShape(const float* points, int num_points, const float* normals, int num_normals,
                       const float* texcoords, int num_texcoords, const int32_t* pidx, const int32_t* nidx,
                       const int32_t* tidx, int num_tris)
{
    for (int i = 0; i < num_points; i++) { Points.push_back(RVec3(points + 3 * i)); Aabb.Expand(Points.back()); }
    for (int i = 0; i < num_normals; i++) Normals.push_back(RVec3(normals + 3 * i));
    for (int i = 0; i < num_texcoords; i++) Texcoords.push_back(RVec3(texcoords[2 * i], texcoords[2 * i + 1], 0.0f));
    PointIndices.assign(pidx, pidx + 3 * (size_t)num_tris);
    for (int i = 0; i < 3 * num_tris; i++)
        if (pidx[i] < 0 || pidx[i] >= num_points) { ErrorText = "RMeshShape: point index out of range"; PointIndices.clear(); Points.clear(); return; }
    if (normals && nidx)
    {
        NormalIndices.assign(nidx, nidx + 3 * (size_t)num_tris);
        for (int i = 0; i < 3 * num_tris; i++)
            if (nidx[i] < 0 || nidx[i] >= num_normals) { ErrorText = "RMeshShape: normal index out of range"; PointIndices.clear(); Points.clear(); return; }
    }
    else
    {
        // no vertex normals: give every corner its triangle's geometric normal
        Normals.clear();
        NormalIndices.resize(3 * (size_t)num_tris);
        for (int t = 0; t < num_tris; t++)
        {
            const RVec3& p0 = Points[pidx[3 * t]]; const RVec3& p1 = Points[pidx[3 * t + 1]]; const RVec3& p2 = Points[pidx[3 * t + 2]];
            float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
            float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
            Normals.push_back(RVec3(ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx));
            NormalIndices[3 * t] = NormalIndices[3 * t + 1] = NormalIndices[3 * t + 2] = t;
        }
    }
    if (texcoords && tidx)
    {
        TexcoordIndices.assign(tidx, tidx + 3 * (size_t)num_tris);
        for (int i = 0; i < 3 * num_tris; i++)
            if (tidx[i] < 0 || tidx[i] >= num_texcoords) { ErrorText = "RMeshShape: texcoord index out of range"; PointIndices.clear(); Points.clear(); return; }
    }
    else
    {
        Texcoords.assign(1, RVec3(0, 0, 0));
        TexcoordIndices.assign(3 * (size_t)num_tris, 0);
    }
    PolyMaterialId.assign(num_tris, -1);
    BuildSpatial();
    Loaded = !BuildFailed;
}

static rt_gpu_ctx* g_device_builder = nullptr;
void SetDeviceBvhBuilder(rt_gpu_ctx* ctx) { g_device_builder = ctx; }

void RMeshShape::BuildSpatial()
{
    const int ntri = (int)PointIndices.size() / 3;
    if (g_device_builder && ntri > 0)
    {
        // KdTree::Build on the GPU (csrc/rt_bvh_build.cu): identical arrays; an error leaves the mesh unloaded
        Flat.nodes.resize(2 * (size_t)ntri - 1);
        Flat.tris.resize((size_t)ntri);
        int32_t depth = 0;
        const int rc = rt_gpu_build_bvh(g_device_builder, &Points[0].x, (int32_t)Points.size(), PointIndices.data(), ntri,
                                        Flat.nodes.data(), Flat.tris.data(), &depth, nullptr);
        if (rc != RT_OK)
        {
            ErrorText = std::string("device BVH build failed: ") + rt_gpu_last_error(g_device_builder);
            Flat.nodes.clear(); Flat.tris.clear();
            BuildFailed = true;
            return;
        }
        Flat.depth = depth;
    }
    else
        Flat.depth = BuildFlatBvh(Points.data(), PointIndices.data(), ntri, Flat.nodes, Flat.tris);

    // material id -> compact texture slot (only ids that actually loaded a texture)
    std::vector<int> slot_of(Textures.size(), -1);
    Flat.textures.clear();
    for (size_t m = 0; m < Textures.size(); m++)
    {
        if (Textures[m])
        {
            slot_of[m] = (int)Flat.textures.size();
            rt_texture t;
            t.rgba = nullptr;                       // 8-bit texels + table: expanded on the device
            t.width = Textures[m]->Width;
            t.height = Textures[m]->Height;
            t.texels8 = Textures[m]->Pixels8.data();
            t.channels = Textures[m]->Channels;
            t.lut = Textures[m]->Lut;
            Flat.textures.push_back(t);
        }
    }

    Flat.shade.resize(ntri);
    for (int t = 0; t < ntri; t++)
    {
        rt_shade& s = Flat.shade[t];
        const RVec3* n[3]; const RVec3* uv[3];
        for (int k = 0; k < 3; k++)
        {
            n[k] = &Normals[NormalIndices[3 * t + k]];
            uv[k] = &Texcoords[TexcoordIndices[3 * t + k]];
        }
        s.n0[0] = n[0]->x; s.n0[1] = n[0]->y; s.n0[2] = n[0]->z;
        s.n1[0] = n[1]->x; s.n1[1] = n[1]->y; s.n1[2] = n[1]->z;
        s.n2[0] = n[2]->x; s.n2[1] = n[2]->y; s.n2[2] = n[2]->z;
        s.uv0[0] = uv[0]->x; s.uv0[1] = uv[0]->y;
        s.uv1[0] = uv[1]->x; s.uv1[1] = uv[1]->y;
        s.uv2[0] = uv[2]->x; s.uv2[1] = uv[2]->y;
        // MeshShape.cpp:310-314: MaterialId != -1 && MaterialId < Textures.size() && Textures[id]
        const int id = PolyMaterialId[t];
        s.texture = (id != -1 && id < (int)Textures.size()) ? slot_of[id] : -1;
    }
}

void RMeshShape::Flatten(rt_shape& out) const
{
    FlattenCommon(out, RT_SHAPE_MESH);
}

} // namespace rtb200
