// localization_headless.cpp -- headless equivalent of localization_node.cpp:108-140 (iterate): pairs one command
// with one measurement message per tick (FIFO by step index) and drives a Filter through the C++ host mirror.
// Input (stdin, text): first line "<filter> <steps>", then per step "fwd ang k id r b ... ".
// Output: per step "timestep x y yaw M", then the final covariance trace.  Used by tests/test_gpu_cpp_host.py.
#include <cstdio>
#include <iostream>
#include <string>

#include "filter.hpp"

int main() {
    std::string choice;
    int steps = 0;
    if (!(std::cin >> choice >> steps)) return 2;
    try {
        auto filter = slam_host::make_filter(choice);
        filter->setCapacity(50, 16);
        filter->readParams(slam_host::default_params());
        filter->init(0.f, 0.f, 0.f);
        for (int t = 0; t < steps; ++t) {
            slam_host::Command cmd;
            int k = 0;
            std::cin >> cmd.fwd >> cmd.ang >> k;
            std::vector<float> meas((size_t)3 * k);
            for (auto& v : meas) std::cin >> v;
            filter->update(cmd, meas);
            const auto x = filter->getStateVector();
            std::printf("%d %.17g %.17g %.17g %d\n", filter->timestep(), x[0], x[1], x[2], filter->M());
        }
        const auto P = filter->covariance();
        const auto x = filter->getStateVector();
        const size_t n = x.size() + ((choice == "ukf_slam" || choice == "ukf_loc") ? 1 : 0);
        double tr = 0;
        for (size_t i = 0; i < n; ++i) tr += P[i * n + i];
        std::printf("trace %.17g ids", tr);
        for (int id : filter->lm_IDs) std::printf(" %d", id);
        std::printf("\n");
    } catch (const std::runtime_error& e) {
        std::fprintf(stderr, "runtime_error: %s\n", e.what());
        return 1;
    }
    return 0;
}
