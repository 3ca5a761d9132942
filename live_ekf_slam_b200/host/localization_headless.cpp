// localization_headless.cpp -- headless equivalent of localization_node.cpp: readParams() (:28-88: params.yaml, the
// `filter:` switch, dt), initCallback (:90-106), and iterate() (:108-140: one queued command paired FIFO with one queued
// measurement message per timer tick, update(), publishState()) driving a Filter through the C++ host mirror.
//
//   localization_headless <params.yaml> [timer_period|default]      (argv[1] of the reference node is the timer period, :180-184)
//   localization_headless                                            legacy form: the filter choice comes on stdin, default params
//
// stdin (text):  ["map" N  id x y ...]   optional /truth/landmarks message (localization_node.cpp trueMapCallback)
//                <steps>                 (legacy form: "<filter> <steps>")
//                per step: fwd ang k  id r b ...
// stdout: per tick "timestep x y yaw M" from the PUBLISHED state message, then "trace <tr P> ids ...", and for the UKF
// "X <count> <sum>" (the sigma-point field of UKFState).  Used by tests/test_gpu_cpp_host.py.
#include <cstdio>
#include <iostream>
#include <queue>
#include <string>

#include "filter.hpp"

using namespace slam_host;

int main(int argc, char** argv) {
    try {
        std::string choice;
        float dt = 0.05f, x0 = 0.f, y0 = 0.f, yaw0 = 0.f;
        std::unique_ptr<Filter> filter;
        if (argc > 1) {
            const YamlNode config = load_yaml_subset(argv[1]);                 // :29-30
            choice = config["filter"].as_string();                             // :33
            filter = make_filter(choice);                                      // :34-45
            filter->setCapacity(50, choice == "ukf_loc" ? 14 : 16);      // (the localisation-only kernel holds at most 14 detections per message)
            filter->readParams(config);                                        // :47
            dt = config["dt"].as_float();                                      // :86
            if (argc > 2 && std::string(argv[2]) != "default") dt = std::stof(argv[2]);   // :180-184
            if (config.has("init_pose")) {
                x0 = config["init_pose"]["x"].as_float(); y0 = config["init_pose"]["y"].as_float(); yaw0 = config["init_pose"]["yaw"].as_float();
            }
        }
        std::string tok;
        if (!(std::cin >> tok)) return 2;
        std::vector<float> map;
        if (tok == "map") {
            int n = 0;
            std::cin >> n;
            map.resize((size_t)3 * n);
            for (auto& v : map) std::cin >> v;
            std::cin >> tok;
        }
        int steps = 0;
        if (argc > 1) steps = std::stoi(tok);
        else {
            choice = tok;
            std::cin >> steps;
            filter = make_filter(choice);
            filter->setCapacity(50, choice == "ukf_loc" ? 14 : 16);      // (the localisation-only kernel holds at most 14 detections per message)
            filter->readParams(default_params());
        }
        // setupStatePublisher (:186-189): the sink stands in for the ROS topic
        const bool is_ukf = choice == "ukf_slam" || choice == "ukf_loc";
        filter->setupStatePublisher([&](const Filter& f) {
            if (auto* e = dynamic_cast<const EKF*>(&f))
                std::printf("%d %.9g %.9g %.9g %d\n", e->stateMsg.timestep, e->stateMsg.x_v, e->stateMsg.y_v, e->stateMsg.yaw_v, e->stateMsg.M);
            else if (auto* u = dynamic_cast<const UKF*>(&f))
                std::printf("%d %.9g %.9g %.9g %d\n", u->stateMsg.timestep, u->stateMsg.x_v, u->stateMsg.y_v, u->stateMsg.yaw_v, u->stateMsg.M);
            else if (auto* nf = dynamic_cast<const NaiveFilter*>(&f))
                std::printf("%d %.9g %.9g %.9g 0\n", nf->stateMsg.timestep, nf->stateMsg.x_v, nf->stateMsg.y_v, nf->stateMsg.yaw_v);
        });
        if (!map.empty()) { if (auto* u = dynamic_cast<UKF*>(filter.get())) u->setMap(map); filter->map = map; }   // trueMapCallback
        filter->init(x0, y0, yaw0);                                            // initCallback, :90-106
        std::queue<Command> cmdQueue;                                          // :14-15
        std::queue<std::vector<float>> lmMeasQueue;
        for (int t = 0; t < steps; ++t) {
            Command cmd;
            int k = 0;
            std::cin >> cmd.fwd >> cmd.ang >> k;
            std::vector<float> meas((size_t)3 * k);
            for (auto& v : meas) std::cin >> v;
            cmdQueue.push(cmd); lmMeasQueue.push(meas);                        // cmdCallback / lmMeasCallback, :142-150
            // iterate(), :108-140
            if (!filter->isInit || cmdQueue.empty() || lmMeasQueue.empty()) continue;
            const Command c = cmdQueue.front(); cmdQueue.pop();
            const std::vector<float> m = lmMeasQueue.front(); lmMeasQueue.pop();
            filter->update(c, m);
            filter->publishState();
        }
        const auto P = filter->covariance();
        const auto x = filter->rawState();
        const size_t n = x.size();
        double tr = 0;
        for (size_t i = 0; i < n; ++i) tr += P[i * n + i];
        std::printf("trace %.17g ids", tr);
        for (int id : filter->lm_IDs) std::printf(" %d", id);
        std::printf("\n");
        if (is_ukf) {
            auto* u = dynamic_cast<UKF*>(filter.get());
            double sum = 0;
            for (float v : u->stateMsg.X) sum += v;
            std::printf("X %zu %.9g\n", u->stateMsg.X.size(), sum);
        }
        std::fprintf(stderr, "filter %s dt %g topic %s\n", choice.c_str(), dt, filter->stateTopic().c_str());
    } catch (const std::runtime_error& e) {
        std::fprintf(stderr, "runtime_error: %s\n", e.what());
        return 1;
    }
    return 0;
}
