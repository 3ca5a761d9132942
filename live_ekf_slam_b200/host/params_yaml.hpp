// params_yaml.hpp -- the subset of YAML that base_pkg/config/params.yaml uses (nested block mappings of scalars,
// '#' comments, optional quotes), read without yaml-cpp (absent from this image), and the keys the filter node reads:
//   localization_node.cpp:28-47,84-87 (`filter`, `dt`), Filter::readCommonParams filter.h:105-121 (noise profiles,
//   constraints.measurements), plus `init_pose` (params.yaml:19-22: the pose sim_node publishes on /truth/init_veh_pose,
//   localization_node.cpp:90-106) and the simulator's constraints (sim_node.py:219-220,239-241).
#pragma once

#include <cstdlib>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/slam_filter.h"

namespace slam_host {

struct YamlNode {
    std::string scalar;
    std::map<std::string, YamlNode> children;
    bool has(const std::string& k) const { return children.count(k) != 0; }
    const YamlNode& operator[](const std::string& k) const {
        auto it = children.find(k);
        if (it == children.end()) throw std::runtime_error("params.yaml: missing key '" + k + "'");   // yaml-cpp: InvalidNode on as<>()
        return it->second;
    }
    double as_double() const {
        char* end = nullptr;
        const double v = std::strtod(scalar.c_str(), &end);
        if (scalar.empty() || end == scalar.c_str()) throw std::runtime_error("params.yaml: '" + scalar + "' is not a number");
        return v;
    }
    float as_float() const { return (float)as_double(); }
    int as_int() const { return (int)as_double(); }
    bool as_bool() const {
        if (scalar == "true" || scalar == "True" || scalar == "TRUE" || scalar == "yes" || scalar == "on") return true;
        if (scalar == "false" || scalar == "False" || scalar == "FALSE" || scalar == "no" || scalar == "off") return false;
        throw std::runtime_error("params.yaml: '" + scalar + "' is not a boolean");
    }
    const std::string& as_string() const { return scalar; }
};

inline std::string yaml_trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// block mappings only: "key: value" / "key:" + deeper-indented children
inline YamlNode load_yaml_subset(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);    // YAML::LoadFile throws BadFile
    YamlNode root;
    std::vector<std::pair<int, YamlNode*>> stack{{-1, &root}};
    std::string line;
    while (std::getline(in, line)) {
        bool quoted = false;                                      // strip comments outside quotes
        for (size_t i = 0; i < line.size(); ++i) {
            if (line[i] == '"' || line[i] == '\'') quoted = !quoted;
            if (line[i] == '#' && !quoted && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) { line.erase(i); break; }
        }
        if (yaml_trim(line).empty()) continue;
        const int indent = (int)line.find_first_not_of(' ');
        const std::string body = yaml_trim(line);
        const size_t colon = body.find(':');
        if (colon == std::string::npos) continue;                 // sequences / flow style do not occur in params.yaml keys we read
        std::string key = yaml_trim(body.substr(0, colon)), val = yaml_trim(body.substr(colon + 1));
        if (val.size() >= 2 && ((val.front() == '"' && val.back() == '"') || (val.front() == '\'' && val.back() == '\''))) val = val.substr(1, val.size() - 2);
        while (stack.back().first >= indent) stack.pop_back();
        YamlNode& node = stack.back().second->children[key];
        node.scalar = val;
        stack.push_back({indent, &node});
    }
    return root;
}

// Filter::readCommonParams, filter.h:105-121: yaml values go into slam_params unchanged; the V/W mix-up of :116-117 is
// reproduced inside the library when compat_noise_bug != 0 (the default; an optional top-level key can switch it off)
inline slam_params read_common_params(const YamlNode& config) {
    slam_params p{};
    p.v_d = config["process_noise"]["mean"]["v_d"].as_float();
    p.v_th = config["process_noise"]["mean"]["v_th"].as_float();
    p.V_00 = config["process_noise"]["cov"]["V_00"].as_double();
    p.V_11 = config["process_noise"]["cov"]["V_11"].as_double();
    p.w_r = config["sensing_noise"]["mean"]["w_r"].as_float();
    p.w_b = config["sensing_noise"]["mean"]["w_b"].as_float();
    p.W_00 = config["sensing_noise"]["cov"]["W_00"].as_double();
    p.W_11 = config["sensing_noise"]["cov"]["W_11"].as_double();
    p.landmark_id_is_known = config["constraints"]["measurements"]["landmark_id_is_known"].as_bool() ? 1 : 0;
    p.min_landmark_separation = config["constraints"]["measurements"]["min_landmark_separation"].as_float();
    p.compat_noise_bug = config.has("compat_noise_bug") ? (config["compat_noise_bug"].as_bool() ? 1 : 0) : 1;
    // the simulator's constraints (only slam_sim_* reads them)
    p.d_max = config["constraints"]["commands"]["d_max"].as_double();
    p.th_max = config["constraints"]["commands"]["th_max"].as_double();
    p.range_max = config["constraints"]["vision"]["range_max"].as_double();
    p.fov_min = config["constraints"]["vision"]["fov_min"].as_double();
    p.fov_max = config["constraints"]["vision"]["fov_max"].as_double();
    return p;
}

}  // namespace slam_host
