// The reference's crate-level doc example (src/lib.rs:24-47) written against the C++ mirror.
// Built by tests/test_cpp_mirror.py; exits 0 when the scan finds the one expected pair.
#include <cstdio>

#include "../broadphase/layer.hpp"

using namespace broadphase;

int main() {
    try {
        auto layer = LayerBuilder().with_min_depth(0).build<Index64_3D, uint64_t>();
        Bounds<3> system_bounds{{-10.f, -10.f, -10.f}, {10.f, 10.f, 10.f}};
        std::vector<std::pair<Bounds<3>, uint64_t>> objects = {
            {Bounds<3>{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}}, 7},
            {Bounds<3>{{0.5f, 0.5f, 0.5f}, {1.5f, 1.5f, 1.5f}}, 9},
            {Bounds<3>{{-9.f, -9.f, -9.f}, {-8.f, -8.f, -8.f}}, 11},
        };
        layer.clear();
        layer.extend(system_bounds, objects.begin(), objects.end());
        auto pairs = layer.scan();
        std::printf("records=%zu pairs=%zu\n", layer.len(), pairs.size());
        for (const auto &p : pairs) std::printf("(%llu, %llu)\n", (unsigned long long)p.first, (unsigned long long)p.second);
        // (the view borrows the layer's result buffer: it is only valid until the next call, like the reference's &Vec)
        const bool scan_ok = pairs.size() == 1 && pairs[0].first == 9 && pairs[0].second == 7;
        // a box query around the far object (Layer::test_box, src/layer.rs:293-311)
        auto hit = layer.test_box(system_bounds, Bounds<3>{{-9.5f, -9.5f, -9.5f}, {-8.5f, -8.5f, -8.5f}});
        std::printf("test_box=%zu", hit.size());
        for (auto id : hit) std::printf(" %llu", (unsigned long long)id);
        std::printf("\n");
        const bool query_ok = hit.size() == 1 && hit[0] == 11;
        // Clone / PartialEq (src/layer.rs:576-617): reported, not part of the exit status
        try {
            auto copy = layer.clone();
            const bool same = copy == layer;
            copy.clear();
            std::printf("clone_equal=%d cleared_clone_differs=%d\n", same ? 1 : 0, copy != layer ? 1 : 0);
        } catch (const Error &e) {
            std::printf("clone: %s\n", e.what());
        }
        // the sharded frame with a world of one rank (the same C++ path that runs over NVLink with more): an empty frame
        DistFrame<Index64_3D> dist(0, 1, -1, 1 << 16, 1 << 16);
        const auto none = dist.frame(system_bounds, nullptr, nullptr, 0);
        std::printf("dist frame pairs=%zu\n", none.second);
        return (scan_ok && query_ok && none.second == 0) ? 0 : 1;
    } catch (const Error &e) {
        std::printf("error: %s\n", e.what());
        return e.status == BP_ERR_CUDA ? 77 : 2; // 77: no GPU here
    }
}
