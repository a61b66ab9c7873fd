// cq_assemble.h — host half of the mesh upload: where each of the caller's parts (cq_mesh_part) goes in the two triangle
// sets' input arrays (partitionEntities + the per-entity loops of TriangleMeshSet.rebuild, CollisionQuery.swift:331-417,
// 886-900) and which H2D copies put it there.  Host-only and header-only (no CUDA types) so that tests/ can compile and
// check it without a GPU; cq_api.cu / cq_build.cu execute the plan.
//
// Nothing is re-laid-out on the host.  The raw arrays (packed xyz positions, part-local u32 indices) go to the device
// as they are — large parts straight from the caller's memory, runs of small parts through one staging buffer so that
// a world of thousands of small parts still costs a handful of copies — and two kernels (k_expand_verts / k_expand_tris in
// cq_build.cu) turn them into float4 vertices tagged with the part, set-global indices, and per-triangle layer / part
// words, checking the index range on the way.  The earlier host-side assembly built 280 MB of staging arrays for the
// 10 M-triangle terrain (0.6 s with push_back, 0.25 s exactly sized: the page faults of fresh memory) — most of a 1.0 s
// cq_world_create around 5 ms of build kernels.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/cq.h"

namespace cq {

struct PartRow { // one part of a set, as the expansion kernels see it (device copy: 32 B rows)
    int32_t vertLo;   // first vertex of the part in the set's vertex arrays
    int32_t nVerts;
    int32_t triStart; // first input triangle of the part (before the degenerate filter)
    int32_t nTris;
    uint32_t layer;
    int32_t part;     // index into the caller's part array
    int32_t pad[2];
};

struct UploadCopy {
    int kind;          // 0: positions (unit = 3 floats), 1: indices (unit = 3 u32)
    size_t dstUnit;    // first vertex / triangle of the set the copy lands on
    size_t nUnits;
    const void *src;   // caller memory, or nullptr: `stagedOffset` units into stagedPos / stagedIdx
    size_t stagedOffset;
};

struct SetPlan { // one TriangleMeshSet before the degenerate filter
    size_t nVerts = 0, nTris = 0;
    std::vector<PartRow> rows;       // the set's parts in the caller's order
    std::vector<UploadCopy> copies;  // in destination order per kind
    std::vector<float> stagedPos;    // 3 per vertex: the small parts' positions, runs back to back
    std::vector<uint32_t> stagedIdx; // 3 per triangle
    std::vector<int> partTriStart;   // rows[k].triStart, then the end (cq_world_create turns it into filtered ranges)
    std::vector<int> partOfSet;      // rows[k].part
};

struct PartPlacement {
    int set;            // 0 static, 1 dynamic
    int vertLo, vertHi; // range in the set's vertex arrays
};

struct PlanError {
    int code = CQ_OK; // CQ_ERR_INVALID
    int part = -1;
    int kind = 0; // 1 = invalid arrays, 3 = more than INT32_MAX vertices / indices in a set
};

// Parts below this many vertices (triangles) are staged; larger ones are copied from the caller's arrays directly.
#define CQ_STAGE_LIMIT 16384

// parts -> plans[0] (static: no body or bodyType == .static) and plans[1] (dynamic).  Parts keep the caller's order
// within their set (the oracle fixes entity order = array order); a triangle is three consecutive indices, a trailing
// partial triple is ignored (`while tri + 2 < count`, CollisionQuery.swift:376).
inline PlanError plan_upload(const cq_mesh_part *parts, int nParts, SetPlan plans[2], std::vector<PartPlacement> &place,
                             size_t stageLimit = CQ_STAGE_LIMIT) {
    PlanError err;
    place.assign((size_t)std::max(nParts, 0), PartPlacement{0, 0, 0});
    for (int s = 0; s < 2; s++) plans[s] = SetPlan();
    for (int p = 0; p < nParts; p++) {
        const cq_mesh_part &mp = parts[p];
        if (mp.n_verts < 0 || mp.n_indices < 0 || (mp.n_verts > 0 && !mp.positions_xyz) || (mp.n_indices > 0 && !mp.indices)) {
            err.code = CQ_ERR_INVALID, err.part = p, err.kind = 1;
            return err;
        }
        const int s = mp.is_dynamic ? 1 : 0;
        SetPlan &P = plans[s];
        const size_t nv = (size_t)mp.n_verts, nt = (size_t)(mp.n_indices / 3);
        if (P.nVerts + nv > (size_t)INT32_MAX || P.nTris + nt > (size_t)INT32_MAX / 3) {
            err.code = CQ_ERR_INVALID, err.part = p, err.kind = 3;
            return err;
        }
        place[p] = PartPlacement{s, (int)P.nVerts, (int)(P.nVerts + nv)};
        P.rows.push_back(PartRow{(int32_t)P.nVerts, (int32_t)nv, (int32_t)P.nTris, (int32_t)nt, mp.layer, p, {0, 0}});
        P.partTriStart.push_back((int)P.nTris);
        P.partOfSet.push_back(p);
        for (int kind = 0; kind < 2; kind++) {
            const size_t units = kind == 0 ? nv : nt, dst = kind == 0 ? P.nVerts : P.nTris;
            if (!units) continue;
            const void *src = kind == 0 ? (const void *)mp.positions_xyz : (const void *)mp.indices;
            if (units >= stageLimit) {
                P.copies.push_back(UploadCopy{kind, dst, units, src, 0});
                continue;
            }
            size_t at;
            if (kind == 0) {
                at = P.stagedPos.size() / 3;
                P.stagedPos.insert(P.stagedPos.end(), mp.positions_xyz, mp.positions_xyz + 3 * units);
            } else {
                at = P.stagedIdx.size() / 3;
                P.stagedIdx.insert(P.stagedIdx.end(), mp.indices, mp.indices + 3 * units);
            }
            // extend the previous copy of this kind when it is staged and ends exactly where this one starts
            UploadCopy *prev = nullptr;
            for (size_t k = P.copies.size(); k-- > 0;)
                if (P.copies[k].kind == kind) {
                    prev = &P.copies[k];
                    break;
                }
            if (prev && !prev->src && prev->dstUnit + prev->nUnits == dst && prev->stagedOffset + prev->nUnits == at)
                prev->nUnits += units;
            else
                P.copies.push_back(UploadCopy{kind, dst, units, nullptr, at});
        }
        P.nVerts += nv, P.nTris += nt;
    }
    for (int s = 0; s < 2; s++) plans[s].partTriStart.push_back((int)plans[s].nTris);
    return err;
}

// Row of the part that owns vertex / input triangle `key` (0 <= key < total): the last row whose start is <= key.  Rows of
// empty parts share their start with the next row, so the last such row is never empty.  Shared by the expansion kernels
// (cq_build.cu) and the host (error message, tests).
#ifdef __CUDACC__
__host__ __device__
#endif
inline int part_row_of(const PartRow *rows, int nRows, int key, bool byTriangle) {
    int lo = 0, hi = nRows; // first row with start > key
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int start = byTriangle ? rows[mid].triStart : rows[mid].vertLo;
        if (start <= key) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

} // namespace cq
