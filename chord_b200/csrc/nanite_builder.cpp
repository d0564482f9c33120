// nanite_builder.cpp — meshlet-DAG builder, the hot loops of chord_b200/nanite_builder.py in C++ so that
// multi-million-triangle meshes (Bistro: 3.8 M) build in minutes.  Offline tool, host only (the reference runs NaniteBuilder::build
// at asset-import time, source/asset/nanite_builder.cpp:432-921); the per-frame path never calls it.  Own code: where the reference
// calls third-party libraries this file restates the published algorithm instead of linking them:
//
//   reference (nanite_builder.cpp)                         here
//   meshopt_buildMeshlets, 255 verts / 128 tris (:443-486)  greedy region growing over triangle adjacency, Morton-ordered seeds
//   METIS_PartGraphKway, groups of <= 4 meshlets (:584-726) greedy grouping by shared boundary edges
//   meshopt_simplifyWithAttributes + LockBorder (:806-841)  quadric edge collapse onto existing vertices, group border locked
//   accept if <= 80 % of the triangles remain (:840)         same
//   parent error = simplify error + max child (:851)         same; centre = AABB centre of the simplified vertices (:857-866)
//   meshopt_computeMeshletBounds (cone)                      normal cone with meshopt's apex construction (Ritter spheres)
//
// Output: GPUGLTFMeshlet records + meshletData ([remap | triangles] per meshlet, asset_gltf_helper.cpp:522-548) + the per-meshlet
// (error, centre, parentError, parentCentre) from which chord_b200/nanite_builder.py forms cluster groups and the BVH.
// C ABI, no CUDA; built by chord_b200/build.py into chord_b200/libchord_builder.so.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <unordered_map>
#include <vector>

#include "../../include/chord_abi_types.h"

namespace {

constexpr uint32_t kMaxV = CHORD_MESHLET_MAX_VERTICES, kMaxT = CHORD_MESHLET_MAX_TRIANGLES, kGroupMax = CHORD_CLUSTER_GROUP_MERGE_MAX;
constexpr double kSimplifyTarget = 0.5, kMinReduce = 0.8;

struct Tri { uint32_t v[3]; };
struct V3 { double x, y, z; };
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Meshlet {
    std::vector<Tri> tris;            // global vertex ids
    uint32_t lod; float error; float center[3]; float parentError; float parentCenter[3]; bool hasParent;
};

static inline uint64_t edge_key(uint32_t a, uint32_t b) { return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a; }

static inline uint32_t part1by2(uint32_t x) { x &= 0x3FF; x = (x | (x << 16)) & 0x30000FF; x = (x | (x << 8)) & 0x300F00F; x = (x | (x << 4)) & 0x30C30C3; x = (x | (x << 2)) & 0x9249249; return x; }

// ---- greedy region growing: clusters of <= 128 triangles / <= 255 vertices (meshopt_buildMeshlets stand-in) ----------------
static void build_meshlets(const float* pos, const std::vector<Tri>& tris, std::vector<std::vector<uint32_t>>& out)
{
    const uint32_t n = (uint32_t)tris.size();
    out.clear();
    if (!n) return;
    // triangle adjacency across shared edges: sort (edge, triangle) pairs
    std::vector<std::pair<uint64_t, uint32_t>> et; et.reserve((size_t)n * 3);
    for (uint32_t t = 0; t < n; t++) for (int k = 0; k < 3; k++) et.emplace_back(edge_key(tris[t].v[k], tris[t].v[(k + 1) % 3]), t);
    std::sort(et.begin(), et.end());
    std::vector<uint32_t> adjStart(n + 1, 0), adj;
    {
        std::vector<std::pair<uint32_t, uint32_t>> links;
        for (size_t i = 0; i < et.size();) {
            size_t j = i; while (j < et.size() && et[j].first == et[i].first) j++;
            for (size_t a = i; a < j; a++) for (size_t b = i; b < j; b++) if (a != b && et[a].second != et[b].second) links.emplace_back(et[a].second, et[b].second);
            i = j;
        }
        for (auto& l : links) adjStart[l.first + 1]++;
        for (uint32_t t = 0; t < n; t++) adjStart[t + 1] += adjStart[t];
        adj.resize(links.size());
        std::vector<uint32_t> fill(adjStart.begin(), adjStart.end() - 1);
        for (auto& l : links) adj[fill[l.first]++] = l.second;
    }
    // Morton order of the centroids: seeds (and restarts when a frontier runs dry) walk this order, so neighbours in the list
    // are neighbours in space (meshopt uses a kd-tree for the same purpose)
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    std::vector<float> cent((size_t)n * 3);
    for (uint32_t t = 0; t < n; t++) for (int c = 0; c < 3; c++) {
        const float v = (pos[3 * (size_t)tris[t].v[0] + c] + pos[3 * (size_t)tris[t].v[1] + c] + pos[3 * (size_t)tris[t].v[2] + c]) / 3.0f;
        cent[3 * (size_t)t + c] = v; mn[c] = std::min(mn[c], v); mx[c] = std::max(mx[c], v);
    }
    std::vector<std::pair<uint32_t, uint32_t>> order(n);
    for (uint32_t t = 0; t < n; t++) {
        uint32_t q[3];
        for (int c = 0; c < 3; c++) { const float e = mx[c] - mn[c]; q[c] = e > 0 ? (uint32_t)std::min(1023.0f, (cent[3 * (size_t)t + c] - mn[c]) / e * 1023.0f) : 0u; }
        order[t] = { part1by2(q[0]) | (part1by2(q[1]) << 1) | (part1by2(q[2]) << 2), t };
    }
    std::sort(order.begin(), order.end());
    std::vector<uint32_t> rank(n);
    for (uint32_t i = 0; i < n; i++) rank[order[i].second] = i;
    std::vector<uint8_t> used(n, 0);
    // dense local vertex numbering (the per-group calls see a few hundred vertices of a multi-million-vertex mesh)
    std::vector<uint32_t> uniq; uniq.reserve((size_t)n * 3);
    for (auto& t : tris) for (int k = 0; k < 3; k++) uniq.push_back(t.v[k]);
    std::sort(uniq.begin(), uniq.end()); uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    std::vector<Tri> ltris(n);
    for (uint32_t t = 0; t < n; t++) for (int k = 0; k < 3; k++) ltris[t].v[k] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), tris[t].v[k]) - uniq.begin());
    std::vector<uint32_t> stamp(uniq.size(), 0xFFFFFFFFu);
    uint32_t cursor = 0;
    std::vector<uint32_t> frontier; size_t head;
    while (true) {
        while (cursor < n && used[order[cursor].second]) cursor++;
        if (cursor >= n) break;
        const uint32_t cid = (uint32_t)out.size();
        std::vector<uint32_t> cluster; cluster.reserve(kMaxT);
        uint32_t nVerts = 0;
        frontier.clear(); head = 0; frontier.push_back(order[cursor].second);
        uint32_t local = cursor;                       // restart position in Morton order, near the cluster
        while (cluster.size() < kMaxT) {
            if (head >= frontier.size()) {
                if (cluster.empty()) break;
                // frontier ran dry before the meshlet is full: continue from the nearest unused triangle in Morton order
                uint32_t a = local; while (a < n && used[order[a].second]) a++;
                if (a >= n || a > local + 4096u) break;               // nothing close: close the meshlet
                local = a;
                frontier.push_back(order[a].second);
            }
            const uint32_t t = frontier[head++];
            if (used[t]) continue;
            uint32_t add = 0;
            for (int k = 0; k < 3; k++) if (stamp[ltris[t].v[k]] != cid) { bool dup = false; for (int j = 0; j < k; j++) dup |= ltris[t].v[j] == ltris[t].v[k]; if (!dup) add++; }
            if (nVerts + add > kMaxV) { if (head >= frontier.size()) break; continue; }
            used[t] = 1; cluster.push_back(t);
            for (int k = 0; k < 3; k++) if (stamp[ltris[t].v[k]] != cid) { stamp[ltris[t].v[k]] = cid; nVerts++; }
            for (uint32_t a = adjStart[t]; a < adjStart[t + 1]; a++) if (!used[adj[a]]) frontier.push_back(adj[a]);
        }
        if (cluster.empty()) { used[order[cursor].second] = 1; continue; }     // (cannot happen: a single triangle always fits)
        out.push_back(std::move(cluster));
    }
}

// ---- groups of <= 4 meshlets sharing the most boundary edges (METIS stand-in) ---------------------------------------------
static void group_meshlets(const std::vector<Meshlet>& ms, std::vector<std::vector<uint32_t>>& groups)
{
    const uint32_t n = (uint32_t)ms.size();
    std::vector<std::pair<uint64_t, uint32_t>> em;
    for (uint32_t m = 0; m < n; m++) for (auto& t : ms[m].tris) for (int k = 0; k < 3; k++) em.emplace_back(edge_key(t.v[k], t.v[(k + 1) % 3]), m);
    std::sort(em.begin(), em.end());
    em.erase(std::unique(em.begin(), em.end()), em.end());            // one entry per (edge, meshlet)
    std::vector<std::unordered_map<uint32_t, uint32_t>> link(n);
    for (size_t i = 0; i < em.size();) {
        size_t j = i; while (j < em.size() && em[j].first == em[i].first) j++;
        for (size_t a = i; a < j; a++) for (size_t b = i; b < j; b++) if (a != b) link[em[a].second][em[b].second]++;
        i = j;
    }
    std::vector<uint8_t> taken(n, 0);
    groups.clear();
    for (uint32_t m = 0; m < n; m++) {
        if (taken[m]) continue;
        taken[m] = 1;
        std::vector<uint32_t> g{m};
        while (g.size() < kGroupMax) {
            std::unordered_map<uint32_t, uint32_t> cand;
            for (uint32_t i : g) for (auto& kv : link[i]) if (!taken[kv.first]) cand[kv.first] += kv.second;
            if (cand.empty()) break;
            uint32_t best = 0xFFFFFFFFu, bw = 0;
            for (auto& kv : cand) if (kv.second > bw || (kv.second == bw && kv.first < best)) { best = kv.first; bw = kv.second; }
            taken[best] = 1; g.push_back(best);
        }
        groups.push_back(std::move(g));
    }
}

// ---- quadric-error edge collapse onto existing vertices, open border locked (meshopt_SimplifyLockBorder stand-in) ----------
struct Quadric { double a[10]; double w; };          // symmetric 4x4 (xx xy xz xw yy yz yw zz zw ww) + weight
static inline void q_add(Quadric& q, const Quadric& o) { for (int i = 0; i < 10; i++) q.a[i] += o.a[i]; q.w += o.w; }
static inline double q_eval(const Quadric& q, V3 p)
{
    const double x = p.x, y = p.y, z = p.z;
    return q.a[0] * x * x + 2 * q.a[1] * x * y + 2 * q.a[2] * x * z + 2 * q.a[3] * x + q.a[4] * y * y + 2 * q.a[5] * y * z + 2 * q.a[6] * y + q.a[7] * z * z + 2 * q.a[8] * z + q.a[9];
}

static void simplify(const float* pos, const std::vector<Tri>& in, size_t target, std::vector<Tri>& out, float& errOut)
{
    // local vertex numbering
    std::unordered_map<uint32_t, uint32_t> toLocal; std::vector<uint32_t> toGlobal;
    std::vector<Tri> tris(in.size());
    for (size_t i = 0; i < in.size(); i++) for (int k = 0; k < 3; k++) {
        auto it = toLocal.find(in[i].v[k]);
        if (it == toLocal.end()) { it = toLocal.emplace(in[i].v[k], (uint32_t)toGlobal.size()).first; toGlobal.push_back(in[i].v[k]); }
        tris[i].v[k] = it->second;
    }
    const uint32_t nv = (uint32_t)toGlobal.size(), nt = (uint32_t)tris.size();
    std::vector<V3> p(nv);
    for (uint32_t v = 0; v < nv; v++) p[v] = { pos[3 * (size_t)toGlobal[v]], pos[3 * (size_t)toGlobal[v] + 1], pos[3 * (size_t)toGlobal[v] + 2] };
    std::unordered_map<uint64_t, uint32_t> edgeCount;
    for (auto& t : tris) for (int k = 0; k < 3; k++) edgeCount[edge_key(t.v[k], t.v[(k + 1) % 3])]++;
    std::vector<uint8_t> locked(nv, 0);
    for (auto& kv : edgeCount) if (kv.second == 1) { locked[(uint32_t)(kv.first >> 32)] = 1; locked[(uint32_t)kv.first] = 1; }
    std::vector<Quadric> Q(nv); memset(Q.data(), 0, sizeof(Quadric) * nv);
    for (auto& t : tris) {
        const V3 n0 = cross(sub(p[t.v[1]], p[t.v[0]]), sub(p[t.v[2]], p[t.v[0]]));
        const double area = std::sqrt(dot(n0, n0));
        if (area == 0) continue;
        const V3 n = { n0.x / area, n0.y / area, n0.z / area };
        const double d = -dot(n, p[t.v[0]]), wgt = std::sqrt(area);      // meshoptimizer's quadricFromTriangle weight: error scales linearly
        Quadric k; const double pl[4] = { n.x, n.y, n.z, d };
        int o = 0; for (int r = 0; r < 4; r++) for (int c = r; c < 4; c++) k.a[o++] = pl[r] * pl[c] * wgt;
        k.w = wgt;
        for (int j = 0; j < 3; j++) q_add(Q[t.v[j]], k);
    }
    std::vector<std::vector<uint32_t>> vt(nv);
    for (uint32_t i = 0; i < nt; i++) for (int k = 0; k < 3; k++) vt[tris[i].v[k]].push_back(i);
    std::vector<uint8_t> alive(nt, 1);
    std::vector<uint32_t> remap(nv); for (uint32_t v = 0; v < nv; v++) remap[v] = v;
    auto find = [&](uint32_t v) { while (remap[v] != v) v = remap[v]; return v; };
    auto cost = [&](uint32_t src, uint32_t dst) {
        Quadric q = Q[src]; q_add(q, Q[dst]);
        return q.w > 0 ? std::fabs(q_eval(q, p[dst])) / q.w : 0.0;       // weighted mean squared distance (meshoptimizer's quadricError)
    };
    struct Cand { double c; uint32_t src, dst; bool operator<(const Cand& o) const { return c > o.c || (c == o.c && (src > o.src || (src == o.src && dst > o.dst))); } };
    std::priority_queue<Cand> heap;
    for (auto& kv : edgeCount) {
        const uint32_t a = (uint32_t)(kv.first >> 32), b = (uint32_t)kv.first;
        if (!locked[a]) heap.push({cost(a, b), a, b});
        if (!locked[b]) heap.push({cost(b, a), b, a});
    }
    size_t nAlive = nt; double maxErr = 0;
    std::vector<uint32_t> nb;
    while (!heap.empty() && nAlive > target) {
        const Cand c = heap.top(); heap.pop();
        if (remap[c.src] != c.src || find(c.dst) != c.dst || c.src == c.dst) continue;
        bool flip = false;
        for (uint32_t ti : vt[c.src]) {                                   // reject collapses that flip a surviving triangle
            if (!alive[ti]) continue;
            uint32_t t[3]; bool hasDst = false;
            for (int k = 0; k < 3; k++) { t[k] = find(tris[ti].v[k]); hasDst |= t[k] == c.dst; }
            if (hasDst) continue;
            V3 o[3], w[3];
            for (int k = 0; k < 3; k++) { o[k] = p[t[k]]; w[k] = p[t[k] == c.src ? c.dst : t[k]]; }
            const V3 n0 = cross(sub(o[1], o[0]), sub(o[2], o[0])), n1 = cross(sub(w[1], w[0]), sub(w[2], w[0]));
            if (dot(n0, n1) <= 1e-12 * std::max(dot(n0, n0), 1e-30)) { flip = true; break; }
        }
        if (flip) continue;
        for (uint32_t ti : vt[c.src]) {
            if (!alive[ti]) continue;
            bool hasDst = false;
            for (int k = 0; k < 3; k++) hasDst |= find(tris[ti].v[k]) == c.dst;
            if (hasDst) { alive[ti] = 0; nAlive--; } else vt[c.dst].push_back(ti);
        }
        remap[c.src] = c.dst;
        q_add(Q[c.dst], Q[c.src]);
        maxErr = std::max(maxErr, c.c);
        nb.clear();
        for (uint32_t ti : vt[c.dst]) if (alive[ti]) for (int k = 0; k < 3; k++) { const uint32_t v = find(tris[ti].v[k]); if (v != c.dst) nb.push_back(v); }
        std::sort(nb.begin(), nb.end()); nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
        for (uint32_t v : nb) {
            if (!locked[v]) heap.push({cost(v, c.dst), v, c.dst});
            if (!locked[c.dst]) heap.push({cost(c.dst, v), c.dst, v});
        }
    }
    out.clear();
    for (uint32_t i = 0; i < nt; i++) if (alive[i]) {
        const uint32_t a = find(tris[i].v[0]), b = find(tris[i].v[1]), c = find(tris[i].v[2]);
        if (a != b && b != c && a != c) out.push_back(Tri{{toGlobal[a], toGlobal[b], toGlobal[c]}});
    }
    errOut = (float)std::sqrt(maxErr);               // root of the mean squared distance: a length in object units
}

// ---- meshopt_computeClusterBounds normal cone (meshopt_clusterizer.cpp:712-850), float32 like the original --------------
static void bounding_sphere(const std::vector<float>& pts, float center[3], float& radius)
{
    const size_t n = pts.size() / 3;
    size_t pmin[3] = {0, 0, 0}, pmax[3] = {0, 0, 0};
    for (size_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) { if (pts[3 * i + a] < pts[3 * pmin[a] + a]) pmin[a] = i; if (pts[3 * i + a] > pts[3 * pmax[a] + a]) pmax[a] = i; }
    float d2[3]; int k = 0;
    for (int a = 0; a < 3; a++) { d2[a] = 0; for (int c = 0; c < 3; c++) { const float d = pts[3 * pmax[a] + c] - pts[3 * pmin[a] + c]; d2[a] += d * d; } }
    for (int a = 0; a < 3; a++) if (d2[a] > d2[k]) k = a;
    for (int c = 0; c < 3; c++) center[c] = (pts[3 * pmin[k] + c] + pts[3 * pmax[k] + c]) / 2.0f;
    radius = d2[k] > 0 ? std::sqrt(d2[k]) / 2.0f : 0.0f;
    for (size_t i = 0; i < n; i++) {
        float dd = 0; for (int c = 0; c < 3; c++) { const float d = pts[3 * i + c] - center[c]; dd += d * d; }
        if (dd > radius * radius) {
            const float d = std::sqrt(dd), kk = 0.5f + (radius / d) / 2.0f;
            for (int c = 0; c < 3; c++) center[c] = center[c] * kk + pts[3 * i + c] * (1.0f - kk);
            radius = (radius + d) / 2.0f;
        }
    }
}

static void cone(const float* pos, const std::vector<Tri>& tris, float axis[3], float& cutoff, float apex[3])
{
    std::vector<float> normals, corners, p0s;
    for (auto& t : tris) {
        const float* a = pos + 3 * (size_t)t.v[0]; const float* b = pos + 3 * (size_t)t.v[1]; const float* c = pos + 3 * (size_t)t.v[2];
        const float e1[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, e2[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
        const float n[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
        const float area = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (!(area > 0)) continue;
        for (int k = 0; k < 3; k++) normals.push_back(n[k] / area);
        for (int k = 0; k < 3; k++) corners.push_back(a[k]); for (int k = 0; k < 3; k++) corners.push_back(b[k]); for (int k = 0; k < 3; k++) corners.push_back(c[k]);
        for (int k = 0; k < 3; k++) p0s.push_back(a[k]);
    }
    axis[0] = axis[1] = axis[2] = 0; apex[0] = apex[1] = apex[2] = 0; cutoff = 0;
    if (normals.empty()) return;                                         // degenerate cluster
    float center[3], r, ax[3], ar;
    bounding_sphere(corners, center, r);
    bounding_sphere(normals, ax, ar);
    const float la = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    const float inv = la == 0 ? 0.0f : 1.0f / la;
    for (int k = 0; k < 3; k++) ax[k] *= inv;
    float mindp = 1.0f;
    const size_t n = normals.size() / 3;
    for (size_t i = 0; i < n; i++) mindp = std::min(mindp, normals[3 * i] * ax[0] + normals[3 * i + 1] * ax[1] + normals[3 * i + 2] * ax[2]);
    if (mindp <= 0.1f) { cutoff = 1.0f; return; }                        // wider than ~168 degrees: never culls
    float maxt = 0;
    for (size_t i = 0; i < n; i++) {
        const float dp = normals[3 * i] * ax[0] + normals[3 * i + 1] * ax[1] + normals[3 * i + 2] * ax[2];
        const float cn = (center[0] - p0s[3 * i]) * normals[3 * i] + (center[1] - p0s[3 * i + 1]) * normals[3 * i + 1] + (center[2] - p0s[3 * i + 2]) * normals[3 * i + 2];
        maxt = std::max(maxt, cn / dp);
    }
    for (int k = 0; k < 3; k++) { axis[k] = ax[k]; apex[k] = center[k] - ax[k] * maxt; }
    cutoff = std::sqrt(1.0f - mindp * mindp);
}

}  // namespace

struct chord_builder_result {
    std::vector<chord_GPUGLTFMeshlet> meshlets;
    std::vector<uint32_t> data;
    std::vector<float> error, parentError, center, parentCenter;
    std::vector<uint32_t> lodMeshlets, lodTriangles;
};

extern "C" {

int chord_builder_build(const float* positions, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount, uint32_t maxLods,
                        chord_builder_result** out)
{
    if (!positions || !indices || !out || triangleCount == 0 || maxLods == 0) return 1;
    for (size_t i = 0; i < (size_t)triangleCount * 3; i++) if (indices[i] >= vertexCount) return 1;
    std::vector<Tri> tris(triangleCount);
    for (uint32_t t = 0; t < triangleCount; t++) tris[t] = Tri{{indices[3 * (size_t)t], indices[3 * (size_t)t + 1], indices[3 * (size_t)t + 2]}};
    std::vector<Meshlet> all, cur;
    auto make = [&](const std::vector<Tri>& src, uint32_t lod, float error, const float center[3], std::vector<Meshlet>& dst) {
        std::vector<std::vector<uint32_t>> clusters;
        build_meshlets(positions, src, clusters);
        for (auto& c : clusters) {
            Meshlet m; m.lod = lod; m.error = error; m.hasParent = false; m.parentError = FLT_MAX;
            for (int k = 0; k < 3; k++) { m.center[k] = center[k]; m.parentCenter[k] = center[k]; }
            m.tris.reserve(c.size());
            for (uint32_t t : c) m.tris.push_back(src[t]);
            dst.push_back(std::move(m));
        }
    };
    const float zero[3] = {0, 0, 0};
    make(tris, 0, -1.0f, zero, cur);
    auto* res = new chord_builder_result();
    for (uint32_t lod = 0; lod + 1 < maxLods; lod++) {
        size_t nt = 0; for (auto& m : cur) nt += m.tris.size();
        res->lodMeshlets.push_back((uint32_t)cur.size()); res->lodTriangles.push_back((uint32_t)nt);
        if (cur.size() <= 1) break;
        std::vector<std::vector<uint32_t>> groups;
        group_meshlets(cur, groups);
        std::vector<std::vector<Meshlet>> next(groups.size());
        // groups are independent: simplify + re-cluster them in parallel
#pragma omp parallel for schedule(dynamic, 8)
        for (long gi = 0; gi < (long)groups.size(); gi++) {
            const auto& g = groups[gi];
            std::vector<Tri> gt;
            for (uint32_t i : g) gt.insert(gt.end(), cur[i].tris.begin(), cur[i].tris.end());
            std::vector<Tri> simp; float err = 0;
            simplify(positions, gt, (size_t)((double)gt.size() * kSimplifyTarget), simp, err);
            if (simp.empty() || !((double)simp.size() < (double)gt.size() * kMinReduce)) continue;
            float passed = 0.0f;
            for (uint32_t i : g) passed = std::max(passed, cur[i].error);
            const float cerr = err + passed;
            float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (auto& t : simp) for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) { const float v = positions[3 * (size_t)t.v[k] + c]; mn[c] = std::min(mn[c], v); mx[c] = std::max(mx[c], v); }
            float center[3]; for (int c = 0; c < 3; c++) center[c] = 0.5f * (mx[c] + mn[c]);
            for (uint32_t i : g) { cur[i].hasParent = true; cur[i].parentError = cerr; for (int c = 0; c < 3; c++) cur[i].parentCenter[c] = center[c]; }
            make(simp, lod + 1, cerr, center, next[gi]);
        }
        for (auto& m : cur) all.push_back(std::move(m));
        cur.clear();
        for (auto& v : next) for (auto& m : v) cur.push_back(std::move(m));
        if (cur.empty()) break;
    }
    if (!cur.empty() && res->lodMeshlets.size() < maxLods && (res->lodMeshlets.empty() || all.empty() || all.back().lod != cur.front().lod)) {
        size_t nt = 0; for (auto& m : cur) nt += m.tris.size();
        if (res->lodMeshlets.size() == (size_t)cur.front().lod) { res->lodMeshlets.push_back((uint32_t)cur.size()); res->lodTriangles.push_back((uint32_t)nt); }
    }
    for (auto& m : cur) all.push_back(std::move(m));
    // ---- pack (asset_gltf_helper.cpp:522-548): [remap u32 x V][triangle u32 x T] per meshlet; bounds and cone per meshlet
    const size_t nm = all.size();
    res->meshlets.resize(nm); res->error.resize(nm); res->parentError.resize(nm); res->center.resize(nm * 3); res->parentCenter.resize(nm * 3);
    std::vector<uint32_t> offsets(nm + 1, 0);
    std::vector<std::vector<uint32_t>> packed(nm);
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)nm; i++) {
        const Meshlet& m = all[i];
        std::vector<uint32_t> vs;
        for (auto& t : m.tris) for (int k = 0; k < 3; k++) vs.push_back(t.v[k]);
        std::sort(vs.begin(), vs.end()); vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
        std::vector<uint32_t>& d = packed[i];
        d = vs;
        for (auto& t : m.tris) {
            uint32_t l[3];
            for (int k = 0; k < 3; k++) l[k] = (uint32_t)(std::lower_bound(vs.begin(), vs.end(), t.v[k]) - vs.begin());
            d.push_back(l[0] | (l[1] << 8) | (l[2] << 16));
        }
        chord_GPUGLTFMeshlet& r = res->meshlets[i];
        memset(&r, 0, sizeof(r));
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t v : vs) for (int c = 0; c < 3; c++) { mn[c] = std::min(mn[c], positions[3 * (size_t)v + c]); mx[c] = std::max(mx[c], positions[3 * (size_t)v + c]); }
        r.posMin = chord_float3{mn[0], mn[1], mn[2]}; r.posMax = chord_float3{mx[0], mx[1], mx[2]};
        r.vertexTriangleCount = ((uint32_t)vs.size() & 0xFFu) | (((uint32_t)m.tris.size() & 0xFFu) << 8);
        float axis[3], apex[3], cutoff;
        cone(positions, m.tris, axis, cutoff, apex);
        r.coneAxis = chord_float3{axis[0], axis[1], axis[2]}; r.coneCutOff = cutoff; r.coneApex = chord_float3{apex[0], apex[1], apex[2]};
        r.lod = m.lod;
        res->error[i] = m.error; res->parentError[i] = m.hasParent ? m.parentError : FLT_MAX;
        for (int c = 0; c < 3; c++) { res->center[3 * i + c] = m.center[c]; res->parentCenter[3 * i + c] = m.hasParent ? m.parentCenter[c] : m.center[c]; }
    }
    for (size_t i = 0; i < nm; i++) offsets[i + 1] = offsets[i] + (uint32_t)packed[i].size();
    res->data.resize(offsets[nm]);
    for (size_t i = 0; i < nm; i++) { res->meshlets[i].dataOffset = offsets[i]; std::copy(packed[i].begin(), packed[i].end(), res->data.begin() + offsets[i]); }
    *out = res;
    return 0;
}

uint32_t chord_builder_meshlet_count(const chord_builder_result* r) { return r ? (uint32_t)r->meshlets.size() : 0; }
uint32_t chord_builder_data_count(const chord_builder_result* r) { return r ? (uint32_t)r->data.size() : 0; }
uint32_t chord_builder_lod_count(const chord_builder_result* r) { return r ? (uint32_t)r->lodMeshlets.size() : 0; }

int chord_builder_copy(const chord_builder_result* r, chord_GPUGLTFMeshlet* meshlets, uint32_t* data, float* error, float* center, float* parentError, float* parentCenter,
                       uint32_t* lodMeshlets, uint32_t* lodTriangles)
{
    if (!r) return 1;
    if (meshlets) memcpy(meshlets, r->meshlets.data(), sizeof(chord_GPUGLTFMeshlet) * r->meshlets.size());
    if (data) memcpy(data, r->data.data(), 4 * r->data.size());
    if (error) memcpy(error, r->error.data(), 4 * r->error.size());
    if (center) memcpy(center, r->center.data(), 4 * r->center.size());
    if (parentError) memcpy(parentError, r->parentError.data(), 4 * r->parentError.size());
    if (parentCenter) memcpy(parentCenter, r->parentCenter.data(), 4 * r->parentCenter.size());
    if (lodMeshlets) memcpy(lodMeshlets, r->lodMeshlets.data(), 4 * r->lodMeshlets.size());
    if (lodTriangles) memcpy(lodTriangles, r->lodTriangles.data(), 4 * r->lodTriangles.size());
    return 0;
}

void chord_builder_free(chord_builder_result* r) { delete r; }

}  // extern "C"
