// cq_multi_gpu.cpp — all GPUs of one box from compiled host code, no Python, no PyTorch: what a Swift / C++ engine host
// would do where the reference builds its one CollisionQuery per scene (SceneServices.swift:45-50).
//
//   1. cq_group_create_local(n)          one process drives n GPUs (NCCL communicators inside libcq.so)
//   2. cq_world_create_multi             the world (a procedural terrain here) replicated on every GPU
//   3. cq_multi_capsule_cast_batch       a host batch of blocking sweeps sharded over the replicas; checked against the
//                                        same batch on GPU 0 alone, byte for byte
//   4. device-resident shards + cq_group_gather_records_local: every GPU sweeps its contiguous range of the batch, then
//      one NCCL all-gather puts all hit records on every GPU (the one collective of the path); timed with
//      CUDA events, max over the GPUs, with and without the collective
//   5. cq_multi_move_and_slide_batch     characters walking over the terrain, sharded the same way
//
//   nvcc -std=c++17 -O2 examples/cq_multi_gpu.cpp -Lswift-game-engine_b200/csrc -lcq \
//        -Xlinker -rpath -Xlinker $PWD/swift-game-engine_b200/csrc -o cq_multi_gpu
//   ./cq_multi_gpu [gpus=all] [cells=1000] [sweeps=4194304] [steps=5]
//
// Exit codes: 0 ok, 2 usage, 4 no usable CUDA device / library error (cq_last_error() is printed), 5 results differ.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/cq.h"

#define CHECK(x)                                                                 \
    do {                                                                         \
        if ((x) != CQ_OK) {                                                      \
            std::fprintf(stderr, "%s failed: %s\n", #x, cq_last_error());        \
            return 4;                                                            \
        }                                                                        \
    } while (0)
#define CUDA(x)                                                                  \
    do {                                                                         \
        cudaError_t e_ = (x);                                                    \
        if (e_ != cudaSuccess) {                                                 \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));        \
            return 4;                                                            \
        }                                                                        \
    } while (0)

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static float frand() { // xorshift64*, uniform in [0, 1)
    g_rng ^= g_rng >> 12, g_rng ^= g_rng << 25, g_rng ^= g_rng >> 27;
    return (float)((g_rng * 0x2545F4914F6CDD1Dull) >> 40) * (1.0f / 16777216.0f);
}
static float height(float x, float z) { return 3.0f * std::sin(x * 0.05f) * std::cos(z * 0.04f) + 1.5f * std::sin(x * 0.013f + z * 0.021f); }

int main(int argc, char **argv) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
        std::fprintf(stderr, "no usable CUDA device (there is no CPU fallback)\n");
        return 4;
    }
    const int gpus = argc > 1 && std::atoi(argv[1]) > 0 ? std::atoi(argv[1]) : visible;
    const int cells = argc > 2 ? std::atoi(argv[2]) : 1000;
    const int64_t nSweeps = argc > 3 ? std::atoll(argv[3]) : (1 << 22);
    const int steps = argc > 4 ? std::atoi(argv[4]) : 5;
    if (gpus > visible || cells < 2 || nSweeps < gpus) {
        std::fprintf(stderr, "usage: %s [gpus<=%d] [cells>=2] [sweeps] [steps]\n", argv[0], visible);
        return 2;
    }
    // terrain: cells x cells grid of 2 m cells, two triangles per cell
    const float cell = 2.0f, half = 0.5f * cells * cell;
    std::vector<float> pos((size_t)(cells + 1) * (cells + 1) * 3);
    std::vector<uint32_t> idx((size_t)cells * cells * 6);
    for (int j = 0; j <= cells; j++)
        for (int i = 0; i <= cells; i++) {
            float *p = &pos[((size_t)j * (cells + 1) + i) * 3];
            p[0] = -half + i * cell, p[2] = -half + j * cell, p[1] = height(p[0], p[2]);
        }
    for (int j = 0; j < cells; j++)
        for (int i = 0; i < cells; i++) {
            const uint32_t a = j * (cells + 1) + i, b = a + 1, c = a + cells + 1, d = c + 1;
            uint32_t *t = &idx[((size_t)j * cells + i) * 6];
            t[0] = a, t[1] = c, t[2] = d, t[3] = a, t[4] = d, t[5] = b;
        }
    cq_mesh_part part = {};
    part.positions_xyz = pos.data(), part.indices = idx.data(), part.n_verts = (int32_t)(pos.size() / 3), part.n_indices = (int32_t)idx.size();
    for (int k = 0; k < 16; k++) part.model[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    part.layer = 1, part.mu_s = 0.8f, part.mu_k = 0.6f, part.entity_id = 0;

    cq_group *group = nullptr;
    CHECK(cq_group_create_local(gpus, nullptr, &group));
    cq_multi_world *mw = nullptr;
    CHECK(cq_world_create_multi(group, &part, 1, nullptr, &mw));
    cq_world_info info;
    cq_world_get_info(cq_multi_world_replica(mw, 0), &info);
    std::printf("%d GPU(s), %d triangles per replica, LBVH build %.2f ms, reference-order build %.0f ms (host)\n", gpus,
                info.n_static_triangles, info.build_ms, info.ref_order_ms);

    // blocking sweeps of a human-scale capsule standing on the terrain
    std::vector<cq_capsule_cast> q((size_t)nSweeps);
    for (auto &c : q) {
        const float x = (frand() * 2 - 1) * (half - 4), z = (frand() * 2 - 1) * (half - 4), a = frand() * 6.2831853f, l = 0.05f + 0.45f * frand();
        c.from[0] = x, c.from[2] = z, c.from[1] = height(x, z) + 0.9f + 0.02f + 0.48f * frand();
        c.delta[0] = std::cos(a) * l, c.delta[2] = std::sin(a) * l, c.delta[1] = -0.3f * frand();
        c.radius = 0.4f, c.half_height = 0.5f, c.mask = CQ_LAYER_ALL, c.min_normal_y = 0.0f;
    }
    // 3. the sharded host call against GPU 0 alone (a sample of the batch keeps the single-GPU leg short)
    const int32_t nCheck = (int32_t)std::min<int64_t>(nSweeps, 1 << 20);
    std::vector<cq_cast_hit> multi((size_t)nCheck), single((size_t)nCheck);
    CHECK(cq_multi_capsule_cast_batch(mw, q.data(), nCheck, CQ_CAST_BLOCKING, multi.data()));
    CUDA(cudaSetDevice(cq_group_device(group, 0)));
    CHECK(cq_capsule_cast_batch(cq_multi_world_replica(mw, 0), q.data(), nCheck, CQ_CAST_BLOCKING, single.data()));
    if (std::memcmp(multi.data(), single.data(), sizeof(cq_cast_hit) * (size_t)nCheck) != 0) {
        std::fprintf(stderr, "sharded sweeps differ from the single-GPU result\n");
        return 5;
    }
    int64_t hits = 0;
    for (const auto &h : multi) hits += h.triangle_index >= 0;
    std::printf("sharded host batch of %d sweeps == single-GPU result byte for byte (%.1f%% hit)\n", nCheck, 100.0 * hits / nCheck);

    // 4. device-resident shards + the NCCL gather
    std::vector<cq_capsule_cast *> dQ(gpus);
    std::vector<cq_cast_hit *> dLocal(gpus), dAll(gpus);
    std::vector<cudaStream_t> st(gpus);
    std::vector<cudaEvent_t> e0(gpus), e1(gpus), e2(gpus);
    std::vector<const void *> locPtr(gpus);
    std::vector<void *> allPtr(gpus), stPtr(gpus);
    for (int i = 0; i < gpus; i++) {
        int64_t lo, hi;
        cq_shard_range(nSweeps, gpus, i, &lo, &hi);
        CUDA(cudaSetDevice(cq_group_device(group, i)));
        CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        CUDA(cudaEventCreate(&e0[i]));
        CUDA(cudaEventCreate(&e1[i]));
        CUDA(cudaEventCreate(&e2[i]));
        CUDA(cudaMalloc((void **)&dQ[i], sizeof(cq_capsule_cast) * (size_t)(hi - lo)));
        CUDA(cudaMalloc((void **)&dLocal[i], sizeof(cq_cast_hit) * (size_t)(hi - lo)));
        CUDA(cudaMalloc((void **)&dAll[i], sizeof(cq_cast_hit) * (size_t)nSweeps));
        CUDA(cudaMemcpy(dQ[i], q.data() + lo, sizeof(cq_capsule_cast) * (size_t)(hi - lo), cudaMemcpyHostToDevice));
        locPtr[i] = dLocal[i], allPtr[i] = dAll[i], stPtr[i] = st[i];
    }
    float castMs = 0, totalMs = 0;
    for (int rep = 0; rep < steps + 3; rep++) { // 3 warm-up passes
        for (int i = 0; i < gpus; i++) {
            int64_t lo, hi;
            cq_shard_range(nSweeps, gpus, i, &lo, &hi);
            CUDA(cudaSetDevice(cq_group_device(group, i)));
            CUDA(cudaEventRecord(e0[i], st[i]));
            CHECK(cq_capsule_cast_device(cq_multi_world_replica(mw, i), dQ[i], (int32_t)(hi - lo), CQ_CAST_BLOCKING, dLocal[i], st[i]));
            CUDA(cudaEventRecord(e1[i], st[i]));
        }
        CHECK(cq_group_gather_records_local(group, locPtr.data(), nSweeps, sizeof(cq_cast_hit), allPtr.data(), stPtr.data()));
        float worstCast = 0, worstTotal = 0;
        for (int i = 0; i < gpus; i++) {
            CUDA(cudaSetDevice(cq_group_device(group, i)));
            CUDA(cudaEventRecord(e2[i], st[i]));
            CUDA(cudaStreamSynchronize(st[i]));
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, e0[i], e1[i]);
            cudaEventElapsedTime(&b, e0[i], e2[i]);
            worstCast = std::max(worstCast, a), worstTotal = std::max(worstTotal, b);
        }
        if (rep >= 3) castMs += worstCast, totalMs += worstTotal;
    }
    castMs /= steps, totalMs /= steps;
    // every GPU now holds every record: compare GPU (gpus-1)'s copy with the host result of step 3
    std::vector<cq_cast_hit> back((size_t)nCheck);
    CUDA(cudaSetDevice(cq_group_device(group, gpus - 1)));
    CUDA(cudaMemcpy(back.data(), dAll[gpus - 1], sizeof(cq_cast_hit) * (size_t)nCheck, cudaMemcpyDeviceToHost));
    if (std::memcmp(back.data(), multi.data(), sizeof(cq_cast_hit) * (size_t)nCheck) != 0) {
        std::fprintf(stderr, "gathered records differ from the host batch result\n");
        return 5;
    }
    std::printf("device-resident: %lld sweeps over %d GPU(s): %.3f ms sweeps only (%.2f G sweeps/s), %.3f ms with the NCCL gather of %.0f MB "
                "onto every GPU (%.2f G sweeps/s); gathered records == host batch result\n",
                (long long)nSweeps, gpus, castMs, nSweeps / castMs * 1e-6, totalMs, nSweeps * sizeof(cq_cast_hit) * 1e-6,
                nSweeps / totalMs * 1e-6);

    // 5. characters walking over the terrain, one fixed step per call, sharded host batch
    const int32_t nChars = 1 << 18;
    std::vector<cq_character_state> chars((size_t)nChars);
    for (auto &c : chars) {
        const float x = (frand() * 2 - 1) * (half - 10), z = (frand() * 2 - 1) * (half - 10), a = frand() * 6.2831853f, v = 4.5f * frand();
        const float p[3] = {x, height(x, z) + 0.9f + 0.2f, z}, vel[3] = {std::cos(a) * v, 0.0f, std::sin(a) * v};
        cq_character_state_init(&c, p, vel);
    }
    cq_controller_params params;
    cq_controller_params_default(&params);
    params.radius = 0.4f, params.half_height = 0.5f, params.skin_width = 0.08f;
    const float gravity[3] = {0.0f, -98.0f, 0.0f};
    for (int s = 0; s < 4; s++) CHECK(cq_multi_move_and_slide_batch(mw, chars.data(), nChars, &params, 1.0f / 60.0f, gravity, CQ_MAS_APPLY_GRAVITY));
    int64_t grounded = 0;
    for (const auto &c : chars) grounded += c.grounded;
    std::printf("%d characters x 4 fixed steps over the replicas: %.1f%% grounded\n", nChars, 100.0 * grounded / nChars);

    for (int i = 0; i < gpus; i++) {
        cudaSetDevice(cq_group_device(group, i));
        cudaFree(dQ[i]), cudaFree(dLocal[i]), cudaFree(dAll[i]);
        cudaStreamDestroy(st[i]);
    }
    cq_multi_world_destroy(mw);
    cq_group_destroy(group);
    return 0;
}
