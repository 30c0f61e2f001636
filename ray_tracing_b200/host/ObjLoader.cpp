// ObjLoader.cpp — Wavefront OBJ -> Mesh (vertices / triangles / normals), the asset-ingestion step in front of the BVH
// builder.  The reference relies on Unity's importer for Assets/Graphics/*.obj ("v//vn" triangles,
// "v/vt/vn" triangles and quads); this loader produces what that importer hands to RayComputeManager.CreateAllMeshData
// (RayComputeManager.cs:218: Mesh.vertices / Mesh.triangles / Mesh.normals):
//   * one mesh vertex per distinct (position, normal) pair, in order of first use;
//   * polygons fan-triangulated (a quad a b c d -> a b c, a c d);
//   * Unity's handedness conversion: X negated on positions and normals, triangle winding reversed, so that the front face
//     keeps satisfying  dot(dir, cross(B-A, C-A)) < 0  (RayCommon.hlsl:192-195,206);
//   * files without normals get area-weighted vertex normals.
#include "RayComputeManager.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace Seb {

// returns "" on success
std::string LoadObj(const char* path, Mesh& mesh, bool unityHandedness)
{
    FILE* f = fopen(path, "rb");
    if (!f) return std::string("cannot open ") + path;
    std::vector<Vector3> pos, nrm;
    std::map<std::pair<int, int>, int> lookup;
    mesh.vertices.clear(); mesh.normals.clear(); mesh.triangles.clear();
    std::vector<int> normalIndexOfVertex;
    bool missingNormals = false;
    char line[1024];
    int lineNo = 0;
    auto vertexFor = [&](int vi, int ni) -> int {
        const std::pair<int, int> key(vi, ni);
        auto it = lookup.find(key);
        if (it != lookup.end()) return it->second;
        const int id = (int)mesh.vertices.size();
        Vector3 p = pos[vi];
        Vector3 n = ni >= 0 ? nrm[ni] : Vector3{0, 0, 0};
        if (unityHandedness) { p.x = -p.x; n.x = -n.x; }
        mesh.vertices.push_back(p); mesh.normals.push_back(n); normalIndexOfVertex.push_back(ni);
        lookup[key] = id;
        return id;
    };
    while (fgets(line, sizeof(line), f))
    {
        lineNo++;
        if (line[0] == 'v' && line[1] == ' ')
        {
            Vector3 v; if (sscanf(line + 2, "%f %f %f", &v.x, &v.y, &v.z) != 3) { fclose(f); return "bad vertex at line " + std::to_string(lineNo); }
            pos.push_back(v);
        }
        else if (line[0] == 'v' && line[1] == 'n' && line[2] == ' ')
        {
            Vector3 v; if (sscanf(line + 3, "%f %f %f", &v.x, &v.y, &v.z) != 3) { fclose(f); return "bad normal at line " + std::to_string(lineNo); }
            nrm.push_back(v);
        }
        else if (line[0] == 'f' && line[1] == ' ')
        {
            int corner[64]; int nc = 0;
            char* p = line + 2;
            while (*p && nc < 64)
            {
                while (*p == ' ' || *p == '\t') p++;
                if (*p == '\n' || *p == '\r' || *p == 0) break;
                char* end;
                long vi = strtol(p, &end, 10);
                if (end == p) { fclose(f); return "bad face at line " + std::to_string(lineNo); }
                long ni = 0; bool hasN = false;
                p = end;
                if (*p == '/')
                {
                    p++;
                    if (*p != '/') strtol(p, &end, 10), p = end;          // texture index (unused)
                    if (*p == '/') { p++; ni = strtol(p, &end, 10); hasN = end != p; p = end; }
                }
                if (vi < 0) vi = (long)pos.size() + vi + 1;
                if (ni < 0) ni = (long)nrm.size() + ni + 1;
                if (vi < 1 || vi > (long)pos.size() || (hasN && (ni < 1 || ni > (long)nrm.size()))) { fclose(f); return "index out of range at line " + std::to_string(lineNo); }
                if (!hasN) missingNormals = true;
                corner[nc++] = vertexFor((int)vi - 1, hasN ? (int)ni - 1 : -1);
            }
            if (nc < 3) { fclose(f); return "face with fewer than 3 corners at line " + std::to_string(lineNo); }
            for (int k = 1; k + 1 < nc; k++)
            {
                if (unityHandedness) { mesh.triangles.push_back(corner[0]); mesh.triangles.push_back(corner[k + 1]); mesh.triangles.push_back(corner[k]); }
                else { mesh.triangles.push_back(corner[0]); mesh.triangles.push_back(corner[k]); mesh.triangles.push_back(corner[k + 1]); }
            }
        }
    }
    fclose(f);
    if (mesh.triangles.empty()) return "no faces";
    if (missingNormals)
    {
        std::vector<double> acc(mesh.vertices.size() * 3, 0.0);
        for (size_t t = 0; t + 2 < mesh.triangles.size(); t += 3)
        {
            const Vector3 &a = mesh.vertices[mesh.triangles[t]], &b = mesh.vertices[mesh.triangles[t + 1]], &c = mesh.vertices[mesh.triangles[t + 2]];
            const double e1[3] = {b.x - a.x, b.y - a.y, b.z - a.z}, e2[3] = {c.x - a.x, c.y - a.y, c.z - a.z};
            const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            for (int k = 0; k < 3; k++) for (int a2 = 0; a2 < 3; a2++) acc[3 * mesh.triangles[t + k] + a2] += n[a2];
        }
        for (size_t v = 0; v < mesh.vertices.size(); v++)
        {
            if (normalIndexOfVertex[v] >= 0) continue;
            const double l = std::sqrt(acc[3 * v] * acc[3 * v] + acc[3 * v + 1] * acc[3 * v + 1] + acc[3 * v + 2] * acc[3 * v + 2]);
            mesh.normals[v] = l > 0 ? Vector3{(float)(acc[3 * v] / l), (float)(acc[3 * v + 1] / l), (float)(acc[3 * v + 2] / l)} : Vector3{0, 1, 0};
        }
    }
    return "";
}

} // namespace Seb

// ---- flat C entry points (Python / FFI) ---------------------------------------------------------------------------------------
namespace { thread_local std::string g_objErr; thread_local Seb::Mesh g_objMesh; }

extern "C" {

/* Parses the file; returns 0 and the element counts, or a negative code (rthObjLastError gives the text). */
int rthObjLoad(const char* path, int unityHandedness, int* vertexCount, int* indexCount)
{
    g_objErr = Seb::LoadObj(path, g_objMesh, unityHandedness != 0);
    if (!g_objErr.empty()) return RT_E_INVALID;
    *vertexCount = (int)g_objMesh.vertices.size(); *indexCount = (int)g_objMesh.triangles.size();
    return RT_OK;
}
/* Copies the mesh parsed by the last rthObjLoad on this thread. */
int rthObjCopy(float* vertices, float* normals, int* indices)
{
    memcpy(vertices, g_objMesh.vertices.data(), g_objMesh.vertices.size() * sizeof(Seb::Vector3));
    memcpy(normals, g_objMesh.normals.data(), g_objMesh.normals.size() * sizeof(Seb::Vector3));
    memcpy(indices, g_objMesh.triangles.data(), g_objMesh.triangles.size() * sizeof(int));
    return RT_OK;
}
const char* rthObjLastError(void) { return g_objErr.c_str(); }

} // extern "C"
