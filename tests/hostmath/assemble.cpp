// Test harness (CPU only): exposes swift-game-engine_b200/csrc/cq_assemble.h — the host half of the mesh upload — to pytest.
// asm_execute performs the plan's copies with memcpy (what cq_build.cu does with cudaMemcpyAsync) into the raw arrays the
// expansion kernels read.
#include "../../swift-game-engine_b200/csrc/cq_assemble.h"

struct Handle {
    cq::SetPlan plans[2];
    std::vector<cq::PartPlacement> place;
};

extern "C" {

// sizes[4s..4s+3] = nVerts, nTris, nRows, nCopies of set s; err[0..2] = code, part, kind
void *asm_plan(const cq_mesh_part *parts, int n_parts, long long stage_limit, long long *sizes, long long *err) {
    Handle *h = new Handle();
    cq::PlanError e = cq::plan_upload(parts, n_parts, h->plans, h->place, (size_t)stage_limit);
    err[0] = e.code, err[1] = e.part, err[2] = e.kind;
    for (int s = 0; s < 2; s++) {
        sizes[4 * s] = (long long)h->plans[s].nVerts, sizes[4 * s + 1] = (long long)h->plans[s].nTris;
        sizes[4 * s + 2] = (long long)h->plans[s].rows.size(), sizes[4 * s + 3] = (long long)h->plans[s].copies.size();
    }
    return h;
}

// raw_pos: 3*nVerts floats, raw_idx: 3*nTris u32 (both pre-filled with a poison pattern by the caller: every unit must be
// written exactly once), rows: 8 ints per row, covered: per-unit write counts (nVerts + nTris)
void asm_execute(void *hv, int s, float *raw_pos, uint32_t *raw_idx, int32_t *rows, int *part_tri_start, int *part_of_set,
                 int *covered, int *direct_copies) {
    Handle *h = (Handle *)hv;
    cq::SetPlan &P = h->plans[s];
    *direct_copies = 0;
    for (const cq::UploadCopy &c : P.copies) {
        if (c.src) (*direct_copies)++;
        if (c.kind == 0) {
            const float *src = c.src ? (const float *)c.src : P.stagedPos.data() + 3 * c.stagedOffset;
            memcpy(raw_pos + 3 * c.dstUnit, src, sizeof(float) * 3 * c.nUnits);
            for (size_t k = 0; k < c.nUnits; k++) covered[c.dstUnit + k]++;
        } else {
            const uint32_t *src = c.src ? (const uint32_t *)c.src : P.stagedIdx.data() + 3 * c.stagedOffset;
            memcpy(raw_idx + 3 * c.dstUnit, src, sizeof(uint32_t) * 3 * c.nUnits);
            for (size_t k = 0; k < c.nUnits; k++) covered[P.nVerts + c.dstUnit + k]++;
        }
    }
    static_assert(sizeof(cq::PartRow) == 32, "PartRow is 8 ints");
    memcpy(rows, P.rows.data(), sizeof(cq::PartRow) * P.rows.size());
    for (size_t k = 0; k < P.partTriStart.size(); k++) part_tri_start[k] = P.partTriStart[k];
    for (size_t k = 0; k < P.partOfSet.size(); k++) part_of_set[k] = P.partOfSet[k];
}

void asm_placement(void *hv, int part, int *out3) {
    Handle *h = (Handle *)hv;
    out3[0] = h->place[part].set, out3[1] = h->place[part].vertLo, out3[2] = h->place[part].vertHi;
}

void asm_rows_of(void *hv, int s, const int *keys, int n, int by_triangle, int *out) {
    Handle *h = (Handle *)hv;
    for (int i = 0; i < n; i++)
        out[i] = cq::part_row_of(h->plans[s].rows.data(), (int)h->plans[s].rows.size(), keys[i], by_triangle != 0);
}

void asm_free(void *h) { delete (Handle *)h; }
}
