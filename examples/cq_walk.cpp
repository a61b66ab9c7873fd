// cq_walk.cpp — the drop-in boundary used from compiled host code, end to end: load a *.static.json asset with the
// library's loader (StaticMeshLoader.loadStaticMeshAsset, StaticMeshLoader.swift:30), flatten it into parts the way
// DemoScene places an asset (collision hulls if the asset has them, else the render mesh) plus a ground plane, build the
// world, and walk a crowd of characters through KinematicMoveStopSystem steps with the reference's own method names
// (cpp/CollisionQuery.hpp).  Prints what a game loop would read back.
//
//   g++ -std=c++17 -O2 examples/cq_walk.cpp -Lswift-game-engine_b200/csrc -lcq -Wl,-rpath,$PWD/swift-game-engine_b200/csrc -o cq_walk
//   ./cq_walk Game/ornate_mirror.static.json [characters=4096] [steps=120]
//
// Exit codes: 0 ok, 2 usage, 3 asset could not be loaded, 4 no usable CUDA device / world creation failed (there is no CPU
// fallback: the message is cq_last_error()).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../swift-game-engine_b200/cpp/CollisionQuery.hpp"

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <asset.static.json> [characters] [steps]\n", argv[0]);
        return 2;
    }
    const int n = argc > 2 ? std::atoi(argv[2]) : 4096, steps = argc > 3 ? std::atoi(argv[3]) : 120;
    cq_static_mesh_asset *asset = nullptr;
    if (cq_static_mesh_load(argv[1], &asset) != CQ_OK) {
        std::fprintf(stderr, "%s\n", cq_last_error());
        return 3;
    }
    // ground plane 80 x 80 at y = -3 (DemoScene.swift:554-566) + every part of the asset at its own transform
    static const float groundPos[12] = {-40, -3, -40, 40, -3, -40, 40, -3, 40, -40, -3, 40};
    static const uint32_t groundIdx[6] = {0, 2, 1, 0, 3, 2};
    std::vector<cq_mesh_part> parts;
    cq_mesh_part ground = {};
    ground.positions_xyz = groundPos, ground.indices = groundIdx, ground.n_verts = 4, ground.n_indices = 6;
    for (int k = 0; k < 16; k++) ground.model[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    ground.layer = 1, ground.mu_s = 0.8f, ground.mu_k = 0.6f, ground.entity_id = 0;
    parts.push_back(ground);
    int triangles = 2;
    for (int p = 0; p < cq_static_mesh_part_count(asset); p++) {
        const int hulls = cq_static_mesh_hull_count(asset, p);
        for (int h = hulls ? 0 : -1; h < hulls; h++) { // hulls when the asset ships them, else the render mesh (hull = -1)
            cq_mesh_part part = {};
            cq_static_mesh_geometry(asset, p, h, &part.positions_xyz, &part.n_verts, &part.indices, &part.n_indices);
            cq_static_mesh_part_transform(asset, p, part.model);
            part.layer = 1u << 4, part.mu_s = 0.6f, part.mu_k = 0.5f, part.entity_id = (uint32_t)parts.size();
            parts.push_back(part);
            triangles += part.n_indices / 3;
            if (!hulls) break;
        }
    }
    std::printf("%s: %d part(s), %d triangles with the ground plane\n", argv[1], (int)parts.size() - 1, triangles);
    try {
        cqhost::CollisionQuery query(parts); // the library copies the arrays; the asset may be freed afterwards
        cq_static_mesh_free(asset);
        // one query the way a gameplay system asks it
        if (auto hit = query.capsuleCastGround({0.0f, 5.0f, 0.0f}, {0.0f, -20.0f, 0.0f}, 1.5f, 1.0f, 0.5f))
            std::printf("ground below (0, 5, 0): toi %.4f, triangle %d, normal (%.3f, %.3f, %.3f)\n", hit->toi, hit->triangleIndex,
                        hit->triangleNormal.x, hit->triangleNormal.y, hit->triangleNormal.z);
        // a crowd on a ring, walking inwards
        cq_controller_params params;
        cq_controller_params_default(&params);
        std::vector<cq_character_state> crowd((size_t)n);
        for (int i = 0; i < n; i++) {
            const float a = 6.2831853f * (float)i / (float)n, r = 12.0f + 6.0f * (float)(i % 7) / 7.0f;
            const float pos[3] = {r * std::cos(a), 2.0f, r * std::sin(a)}, vel[3] = {-4.0f * std::cos(a), 0.0f, -4.0f * std::sin(a)};
            cq_character_state_init(&crowd[(size_t)i], pos, vel);
        }
        const float gravity[3] = {0.0f, -98.0f, 0.0f};
        const auto t0 = std::chrono::steady_clock::now();
        for (int s = 0; s < steps; s++)
            if (!query.moveAndSlide(crowd.data(), n, params, 1.0f / 60.0f, gravity)) {
                std::fprintf(stderr, "%s\n", cq_last_error());
                return 4;
            }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        int grounded = 0;
        for (const auto &c : crowd) grounded += c.grounded ? 1 : 0;
        std::printf("%d characters x %d steps: %.2f ms per step (host buffers in and out), %d grounded, first at (%.3f, %.3f, %.3f)\n",
                    n, steps, ms / steps, grounded, crowd[0].position[0], crowd[0].position[1], crowd[0].position[2]);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 4;
    }
    return 0;
}
