// CollisionQueryService.hpp — C++ host-side mirror of the reference's `final class CollisionQueryService`
// (Game/SceneServices.swift:33-207) and of the streaming active set that drives it (`ActiveChunkSystem`,
// Game/Systems.swift:2354-2396; `WorldPosition.fromWorld`, Game/Components.swift:55-69).  Host policy only, exactly as in
// the reference: it owns the query object and decides on every fixed step between a full rebuild (cq_world_create) and
// a transform refit (cq_world_update_transforms).  Header-only; the Python twin (swift-game-engine_b200/service.py) makes
// the same decisions on the same inputs and tests/ holds both to one scenario.
//
//   * active set changed, markDirty(), no query yet ......................... rebuild   (:52-60)
//   * entity count changed, mesh.dirty, no snapshot, body type or collides changed,
//     vertex or index count changed ......................................... rebuild   (:95-163 "structuralChange")
//   * squared delta of translation, rotation (quaternion vector) or scale > 1e-6
//     ........................................................................ updateStatic/DynamicTransforms (:66-75)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <vector>

#include "CollisionQuery.hpp"

namespace cqhost {

enum class BodyType { None, Static, Kinematic, Dynamic }; // PhysicsBodyComponent.bodyType, None = no body (Components.swift:549-598)

// One entity with a TransformComponent and a StaticMeshComponent, as the service reads it from the World.
struct MeshEntity {
    uint32_t id = 0;
    Float3 translation{0, 0, 0};
    float rotation[4] = {0, 0, 0, 1}; // simd_quatf vector (ix, iy, iz, r)
    Float3 scale{1, 1, 1};
    const float *positions = nullptr; // collisionMesh ?? mesh: n_verts * 3
    int32_t n_verts = 0;
    const uint32_t *indices = nullptr;
    int32_t n_indices = 0;
    BodyType bodyType = BodyType::None;
    bool collides = true;
    bool dirty = false; // StaticMeshComponent.dirty: cleared by the service when it re-snapshots (:186-189)
    uint32_t layer = 1;
    SurfaceMaterial material;
};

// TransformComponent.modelMatrix = T * (R(quat) * S), column-major like simd's matrix_float4x4 (Components.swift:26-44)
inline void modelMatrix(const MeshEntity &e, float out[16]) {
    const float x = e.rotation[0], y = e.rotation[1], z = e.rotation[2], w = e.rotation[3];
    const float r[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
                           {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
                           {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
    const float s[3] = {e.scale.x, e.scale.y, e.scale.z};
    for (int c = 0; c < 3; c++) {
        for (int row = 0; row < 3; row++) out[4 * c + row] = r[row][c] * s[c];
        out[4 * c + 3] = 0.0f;
    }
    out[12] = e.translation.x, out[13] = e.translation.y, out[14] = e.translation.z, out[15] = 1.0f;
}

inline cq_mesh_part makePart(const MeshEntity &e) {
    cq_mesh_part p = {};
    p.positions_xyz = e.positions, p.indices = e.indices, p.n_verts = e.n_verts, p.n_indices = e.n_indices;
    modelMatrix(e, p.model);
    p.layer = e.layer, p.mu_s = e.material.muS, p.mu_k = e.material.muK, p.flatten_ground = e.material.flattenGround ? 1 : 0;
    p.is_dynamic = (e.bodyType != BodyType::None && e.bodyType != BodyType::Static) ? 1 : 0; // partitionEntities, CollisionQuery.swift:886-900
    p.entity_id = e.id;
    return p;
}

// Query: anything constructible from std::vector<cq_mesh_part> with updateTransforms(ids, models16) — CollisionQuery by
// default; tests substitute a recorder.
template <class Query = CollisionQuery> class CollisionQueryServiceT {
  public:
    enum class Action { None, Refit, Rebuild }; // what the last update() did (instrumentation, not in the reference)
    using ActiveSet = std::optional<std::set<uint32_t>>;

    Query *query() const { return query_.get(); }
    Action lastAction() const { return last_; }
    void markDirty() { dirty_ = true; }

    void rebuild(std::vector<MeshEntity> &entities, const ActiveSet &active = std::nullopt) { // :45-50
        std::vector<cq_mesh_part> parts;
        for (MeshEntity *e : filter(entities, active)) parts.push_back(makePart(*e));
        query_.reset(); // one world at a time on the device
        query_ = std::make_unique<Query>(parts);
        dirty_ = false;
        lastActive_ = active;
        refreshCache(entities, active);
        last_ = Action::Rebuild;
    }

    void update(std::vector<MeshEntity> &entities, const ActiveSet &active = std::nullopt) { // :52-77
        if (active != lastActive_ || dirty_ || !query_) return rebuild(entities, active);
        std::vector<MeshEntity *> st, dy;
        if (changes(entities, active, st, dy)) return rebuild(entities, active);
        last_ = Action::None;
        for (auto *group : {&st, &dy}) {
            if (group->empty()) continue;
            std::vector<uint32_t> ids;
            std::vector<float> models;
            for (MeshEntity *e : *group) {
                ids.push_back(e->id);
                float m[16];
                modelMatrix(*e, m);
                models.insert(models.end(), m, m + 16);
            }
            query_->updateTransforms(ids, models); // updateStaticTransforms / updateDynamicTransforms: one C entry point
            last_ = Action::Refit;
        }
        refreshCache(entities, active);
    }

  private:
    struct Snapshot { // StaticMeshSnapshot (:79-87)
        Float3 translation;
        float rotation[4];
        Float3 scale;
        int32_t vertexCount, indexCount;
        BodyType bodyType;
        bool collides;
    };

    static std::vector<MeshEntity *> filter(std::vector<MeshEntity> &entities, const ActiveSet &active) { // :196-206
        std::map<uint32_t, MeshEntity *> byId; // deterministic order = ascending id (the reference: Dictionary order)
        for (MeshEntity &e : entities) {
            if (active && !active->count(e.id)) continue;
            if (e.collides) byId[e.id] = &e;
        }
        std::vector<MeshEntity *> out;
        for (auto &kv : byId) out.push_back(kv.second);
        return out;
    }

    void refreshCache(std::vector<MeshEntity> &entities, const ActiveSet &active) { // :171-194
        cache_.clear();
        for (MeshEntity *e : filter(entities, active)) {
            Snapshot s{e->translation, {e->rotation[0], e->rotation[1], e->rotation[2], e->rotation[3]}, e->scale, e->n_verts,
                       e->n_indices, e->bodyType, e->collides};
            cache_[e->id] = s;
            e->dirty = false;
        }
    }

    static float sq(float a, float b, float c) { return (a * a + b * b) + c * c; }

    // staticMeshChanges (:95-169): true = structural change; otherwise the entities whose transform moved, per set
    bool changes(std::vector<MeshEntity> &entities, const ActiveSet &active, std::vector<MeshEntity *> &st,
                 std::vector<MeshEntity *> &dy) {
        const std::vector<MeshEntity *> ents = filter(entities, active);
        if (ents.size() != cache_.size()) return true;
        const float eps = 1e-6f;
        for (MeshEntity *e : ents) {
            if (e->dirty) return true;
            auto it = cache_.find(e->id);
            if (it == cache_.end()) return true;
            const Snapshot &s = it->second;
            if (s.bodyType != e->bodyType || s.collides != e->collides) return true;
            const float dr[4] = {e->rotation[0] - s.rotation[0], e->rotation[1] - s.rotation[1], e->rotation[2] - s.rotation[2],
                                 e->rotation[3] - s.rotation[3]};
            const bool moved =
                sq(e->translation.x - s.translation.x, e->translation.y - s.translation.y, e->translation.z - s.translation.z) > eps ||
                ((dr[0] * dr[0] + dr[1] * dr[1]) + dr[2] * dr[2]) + dr[3] * dr[3] > eps ||
                sq(e->scale.x - s.scale.x, e->scale.y - s.scale.y, e->scale.z - s.scale.z) > eps;
            if (moved) (e->bodyType == BodyType::None || e->bodyType == BodyType::Static ? st : dy).push_back(e);
            if (e->n_verts != s.vertexCount || e->n_indices != s.indexCount) return true;
        }
        return false;
    }

    std::unique_ptr<Query> query_;
    bool dirty_ = true;
    std::map<uint32_t, Snapshot> cache_;
    ActiveSet lastActive_;
    Action last_ = Action::None;
};

using CollisionQueryService = CollisionQueryServiceT<CollisionQuery>;

// ---------------------------------------------------------------- streaming active set
constexpr double kChunkSize = 512.0; // WorldPosition.chunkSize (Components.swift:55)

// WorldPosition.fromWorld (Components.swift:58-69): per axis chunk = floor((v + 256) / 512)
inline void worldToChunk(const double world[3], int64_t chunk[3]) {
    for (int k = 0; k < 3; k++) chunk[k] = (int64_t)std::floor((world[k] + kChunkSize * 0.5) / kChunkSize);
}

// ActiveChunkSystem.fixedUpdate (Systems.swift:2354-2396): the ids of the entities whose chunk lies within radiusChunks
// (Chebyshev distance; ActiveChunkComponent.radiusChunks default 2) of the player's chunk.  Feed the result to
// CollisionQueryService::update — a changed set is what triggers the rebuild there (SceneServices.swift:55-58).
inline std::set<uint32_t> activeEntityIDs(const double playerWorld[3], const std::vector<MeshEntity> &entities, int radiusChunks = 2) {
    int64_t centre[3];
    worldToChunk(playerWorld, centre);
    const int64_t radius = radiusChunks > 0 ? radiusChunks : 0;
    std::set<uint32_t> out;
    for (const MeshEntity &e : entities) {
        const double w[3] = {e.translation.x, e.translation.y, e.translation.z};
        int64_t c[3];
        worldToChunk(w, c);
        int64_t d = 0;
        for (int k = 0; k < 3; k++) d = std::max<int64_t>(d, c[k] > centre[k] ? c[k] - centre[k] : centre[k] - c[k]);
        if (d <= radius) out.insert(e.id);
    }
    return out;
}

} // namespace cqhost
