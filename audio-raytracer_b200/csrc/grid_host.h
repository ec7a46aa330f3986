// grid_host.h -- host-side construction of the uniform collider grid (scene_dev.cuh: GridDesc).
//
// Runs at scene upload on the caller's structs (ColliderSphereStruct / ColliderAABBStruct /
// ColliderOBBStruct, Assets/C# Scripts/DataTypes/Collider Structs/*.cs). Every collider is entered
// into all cells its CONSERVATIVE bounds overlap:
//   sphere  C +- (R + mS)          mS covers the cancellation error of the reference's quadratic
//                                  (b*b - 4ac in FP32 can report a graze up to ~3e-7*|oc|^2/R outside)
//   AABB    [min, max] +- m        (RT:286-287 min/max)
//   OBB     C +- (e * 1.001 + m)   e = world extents of the rotated box, for BOTH rotation senses (|R|h and
//                                  |R^T|h), so it covers RT:314-320, PM:294-300 and the double-inverted
//                                  PM:172-179 alike; never larger than the bounding sphere
// with m = 1e-3 + 1e-4 * D, D = 2 * scene diagonal + 33 >= any |origin - centre| the kernels can see (frames whose
// listener is farther than 1.5 diagonals + 32 from the scene centre use the brute-force kernels). A collider that
// the exact FP32 test can possibly report is therefore listed in every cell the ray visits near it.
#pragma once
#include "scene_dev.cuh"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace art {

namespace gridimpl {
struct Box { float lo[3], hi[3]; float r; };   // r: sphere radius (spheres only)
}  // namespace gridimpl

struct HostGrid {
    GridDesc d{};                       // device pointers left null
    std::vector<gridimpl::Box> bS, bA, bO;   // un-inflated bounds per collider (grid_params)
    std::vector<uint2> cells;
    std::vector<uint16_t> entries;
    std::vector<uint2> rangeO;          // per OBB: the cell range it is listed in (packed 8 bits per axis)
    std::vector<float> boxLo, boxHi;    // float4 per collider, canonical order S | A | O: the conservative bounds the cells
                                        // were filled from (inflated by m, spheres by mS as well) -- input of the target fans
    float diag = 0.0f;                  // un-inflated scene diagonal
    float cx = 0, cy = 0, cz = 0;       // scene centre
    float listenerRange = 0.0f;         // ray origins farther than this from the centre -> brute force
    float margin = 0.0f;                // m: inflation of every collider's bounds
    bool ok = false;
    const char* why = "";
};

namespace gridimpl {

inline float h2f(uint16_t h)
{
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, mag = h & 0x7FFFu, bits;
    if (mag >= 0x7C00u) bits = sign | 0x7F800000u | ((mag & 0x3FFu) << 13);
    else if (mag >= 0x0400u) bits = sign | ((mag << 13) + ((127u - 15u) << 23));
    else if (mag == 0) bits = sign;
    else { float f = (float)mag * 5.9604644775390625e-08f; memcpy(&bits, &f, 4); bits |= sign; }
    float out; memcpy(&out, &bits, 4); return out;
}

}  // namespace gridimpl

// Part 1 (cheap, O(colliders)): collider bounds, grid dimensions, margins, the inflated boxes (input of the target fans and of
// the device-side cell fill, k5_grid_build.cu) and the OBB cell ranges. cellScale: cell edge = cellScale * cbrt(volume / colliders).
// g.ok is set when the scene can be gridded; the cell lists are NOT built here.
inline void grid_params(const std::vector<uint16_t>& rawS, const std::vector<uint16_t>& rawA, const std::vector<uint16_t>& rawO,
                        float cellScale, HostGrid& g)
{
    using namespace gridimpl;
    g.ok = false; g.cells.clear(); g.entries.clear(); g.rangeO.clear(); g.boxLo.clear(); g.boxHi.clear();
    std::vector<Box>& bS = g.bS; std::vector<Box>& bA = g.bA; std::vector<Box>& bO = g.bO;
    const size_t ns = rawS.size() / 8, na = rawA.size() / 10, no = rawO.size() / 13;
    if (ns + na + no == 0) { g.why = "empty scene"; return; }
    if (ns > 65535 || na > 65535 || no > 65535) { g.why = "more than 65535 colliders of one type"; return; }

    bS.assign(ns, Box{}); bA.assign(na, Box{}); bO.assign(no, Box{});
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    auto grow = [&](const Box& b) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
    };
    bool finite = true;
    for (size_t i = 0; i < ns; i++) {
        const uint16_t* w = &rawS[8 * i];
        const float R = std::fabs(h2f(w[3]));
        Box b; b.r = R;
        for (int k = 0; k < 3; k++) { const float c = h2f(w[k]); b.lo[k] = c - R; b.hi[k] = c + R; finite = finite && std::isfinite(b.lo[k]) && std::isfinite(b.hi[k]); }
        bS[i] = b; grow(b);
    }
    for (size_t i = 0; i < na; i++) {
        const uint16_t* w = &rawA[10 * i];
        Box b; b.r = 0;
        for (int k = 0; k < 3; k++) {
            const float c = h2f(w[k]), h = h2f(w[3 + k]);
            b.lo[k] = std::min(c - h, c + h); b.hi[k] = std::max(c - h, c + h);
            finite = finite && std::isfinite(b.lo[k]) && std::isfinite(b.hi[k]);
        }
        bA[i] = b; grow(b);
    }
    for (size_t i = 0; i < no; i++) {
        const uint16_t* w = &rawO[13 * i];
        const float hx = std::fabs(h2f(w[3])), hy = std::fabs(h2f(w[4])), hz = std::fabs(h2f(w[5]));
        // Rotation getter (DataTypes/halfQuaternion.cs:34-46): w from x,y,z, then normalise
        float qx = h2f(w[6]), qy = h2f(w[7]), qz = h2f(w[8]);
        const float w2 = 1.0f - (qx * qx + qy * qy + qz * qz);
        float qw = w2 > 0.0f ? std::sqrt(w2) : 0.0f;
        const float qn = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
        Box b; b.r = 0;
        const float rs = std::sqrt(hx * hx + hy * hy + hz * hz);
        float e[3] = { rs, rs, rs };
        if (qn > 1e-6f && std::isfinite(qn)) {
            qx /= qn; qy /= qn; qz /= qn; qw /= qn;
            const float R[3][3] = {
                { 1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw) },
                { 2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw) },
                { 2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy) } };
            const float h[3] = { hx, hy, hz };
            for (int k = 0; k < 3; k++) {
                // the box as RT:314-320 / PM:294-300 see it (world -> local by q) and as PM:172-179 sees it
                // (world -> local by inverse(q)): world extents |R^T| h and |R| h; cover both
                const float e1 = std::fabs(R[k][0]) * h[0] + std::fabs(R[k][1]) * h[1] + std::fabs(R[k][2]) * h[2];
                const float e2 = std::fabs(R[0][k]) * h[0] + std::fabs(R[1][k]) * h[1] + std::fabs(R[2][k]) * h[2];
                e[k] = std::min(rs, std::max(e1, e2));
            }
        }
        for (int k = 0; k < 3; k++) {
            const float c = h2f(w[k]);
            const float ek = e[k] * 1.001f + 1e-4f * rs;
            b.lo[k] = c - ek; b.hi[k] = c + ek;
            finite = finite && std::isfinite(b.lo[k]) && std::isfinite(b.hi[k]);
        }
        bO[i] = b; grow(b);
    }
    if (!finite) { g.why = "non-finite collider bounds"; return; }

    const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    g.diag = std::sqrt(ex * ex + ey * ey + ez * ez);
    g.cx = 0.5f * (lo[0] + hi[0]); g.cy = 0.5f * (lo[1] + hi[1]); g.cz = 0.5f * (lo[2] + hi[2]);
    g.listenerRange = 1.5f * g.diag + 32.0f;
    // error scale: |oc| can reach the listener range plus half the diagonal
    const float D = 2.0f * g.diag + 33.0f;
    const float m = 1e-3f + 1e-4f * D;
    g.margin = m;
    for (int k = 0; k < 3; k++) { lo[k] -= 2.0f * m; hi[k] += 2.0f * m; }

    const size_t nc = ns + na + no;
    const double vol = std::max(1e-9, (double)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]));
    float cs = cellScale * (float)std::cbrt(vol / (double)nc);
    if (!(cs > 0.0f)) cs = 1.0f;
    int dim[3];
    float csz[3];
    for (int k = 0; k < 3; k++) {
        int n = (int)std::ceil((hi[k] - lo[k]) / cs);
        n = std::max(1, std::min(n, 160));
        dim[k] = n;
        csz[k] = (hi[k] - lo[k]) / (float)n;
        if (!(csz[k] > 0.0f)) { csz[k] = 1.0f; dim[k] = 1; }
    }
    GridDesc& d = g.d;
    d.g0x = lo[0]; d.g0y = lo[1]; d.g0z = lo[2];
    d.g1x = hi[0]; d.g1y = hi[1]; d.g1z = hi[2];
    d.csx = csz[0]; d.csy = csz[1]; d.csz = csz[2];
    d.icx = 1.0f / csz[0]; d.icy = 1.0f / csz[1]; d.icz = 1.0f / csz[2];
    d.nx = dim[0]; d.ny = dim[1]; d.nz = dim[2];
    d.errScale = D;
    d.cells = nullptr; d.entries = nullptr; d.rangeO = nullptr;
    auto range = [&](const Box& b, float extra, int i0[3], int i1[3]) {
        for (int k = 0; k < 3; k++) {
            const float a = (b.lo[k] - (m + extra) - lo[k]) / csz[k], z = (b.hi[k] + (m + extra) - lo[k]) / csz[k];
            i0[k] = std::max(0, std::min(dim[k] - 1, (int)std::floor(a)));
            i1[k] = std::max(0, std::min(dim[k] - 1, (int)std::floor(z)));
        }
    };
    auto sphereExtra = [&](const Box& b) { return std::min(D, 1e-6f * D * D / std::max(b.r, 1e-3f)); };
    g.boxLo.reserve(4 * nc + 4); g.boxHi.reserve(4 * nc + 4);
    auto pushBoxes = [&](const std::vector<Box>& bs, int type) {
        for (const Box& b : bs) {
            const float infl = m + (type == 0 ? sphereExtra(b) : 0.0f);
            for (int k = 0; k < 3; k++) { g.boxLo.push_back(b.lo[k] - infl); g.boxHi.push_back(b.hi[k] + infl); }
            g.boxLo.push_back(0.0f); g.boxHi.push_back(0.0f);
        }
    };
    pushBoxes(bS, 0); pushBoxes(bA, 1); pushBoxes(bO, 2);
    g.rangeO.resize(no + 1);
    for (size_t i = 0; i < no; i++) {
        int i0[3], i1[3];
        range(bO[i], 0.0f, i0, i1);
        g.rangeO[i] = make_uint2((uint32_t)(i0[0] | (i0[1] << 8) | (i0[2] << 16)), (uint32_t)(i1[0] | (i1[1] << 8) | (i1[2] << 16)));
    }
    g.ok = true;
    g.why = "";
}

// Part 2, host version (art_grid_build_host: inspection, tests): the cell lists. The library itself fills them on the device.
inline void build_grid(const std::vector<uint16_t>& rawS, const std::vector<uint16_t>& rawA, const std::vector<uint16_t>& rawO,
                       float cellScale, HostGrid& g)
{
    using namespace gridimpl;
    grid_params(rawS, rawA, rawO, cellScale, g);
    if (!g.ok) return;
    g.ok = false;
    const std::vector<Box>& bS = g.bS; const std::vector<Box>& bA = g.bA; const std::vector<Box>& bO = g.bO;
    GridDesc& d = g.d;
    const int dim[3] = { d.nx, d.ny, d.nz };
    const float lo[3] = { d.g0x, d.g0y, d.g0z }, csz[3] = { d.csx, d.csy, d.csz };
    const float m = g.margin, D = d.errScale;
    const size_t nCells = (size_t)dim[0] * dim[1] * dim[2];
    auto range = [&](const Box& b, float extra, int i0[3], int i1[3]) {
        for (int k = 0; k < 3; k++) {
            const float a = (b.lo[k] - (m + extra) - lo[k]) / csz[k], z = (b.hi[k] + (m + extra) - lo[k]) / csz[k];
            i0[k] = std::max(0, std::min(dim[k] - 1, (int)std::floor(a)));
            i1[k] = std::max(0, std::min(dim[k] - 1, (int)std::floor(z)));
        }
    };
    auto sphereExtra = [&](const Box& b) { return std::min(D, 1e-6f * D * D / std::max(b.r, 1e-3f)); };

    // pass 1: counts per (cell, type)
    std::vector<uint32_t> cnt(nCells * 3, 0);
    auto visit = [&](const std::vector<Box>& bs, int type, auto&& fn) {
        for (size_t i = 0; i < bs.size(); i++) {
            int i0[3], i1[3];
            range(bs[i], type == 0 ? sphereExtra(bs[i]) : 0.0f, i0, i1);
            for (int z = i0[2]; z <= i1[2]; z++)
                for (int y = i0[1]; y <= i1[1]; y++)
                    for (int x = i0[0]; x <= i1[0]; x++) fn(((size_t)z * dim[1] + y) * dim[0] + x, type, (uint16_t)i);
        }
    };
    size_t total = 0;
    auto count = [&](size_t cell, int type, uint16_t) { cnt[cell * 3 + type]++; total++; };
    visit(bS, 0, count); visit(bA, 1, count); visit(bO, 2, count);
    if (total > ((size_t)1 << 28)) { g.why = "grid too large"; return; }
    g.cells.resize(nCells);
    std::vector<uint32_t> cursor(nCells * 3);
    uint32_t off = 0;
    for (size_t c = 0; c < nCells; c++) {
        const uint32_t s = cnt[c * 3], a = cnt[c * 3 + 1], o = cnt[c * 3 + 2];
        if (s > (uint32_t)kGridMaxS || a > (uint32_t)kGridMaxA || o > (uint32_t)kGridMaxO) { g.why = "cell list too long"; return; }
        g.cells[c] = make_uint2(off, s | (a << 10) | (o << 21));
        cursor[c * 3] = off; cursor[c * 3 + 1] = off + s; cursor[c * 3 + 2] = off + s + a;
        off += s + a + o;
    }
    g.entries.resize((size_t)off + 8);
    d.nEntries = (int)off;
    auto fill = [&](size_t cell, int type, uint16_t id) { g.entries[cursor[cell * 3 + type]++] = id; };
    visit(bS, 0, fill); visit(bA, 1, fill); visit(bO, 2, fill);   // ascending canonical index inside each list
    g.ok = true;
    g.why = "";
}

}  // namespace art
