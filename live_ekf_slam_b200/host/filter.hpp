// filter.hpp -- C++ host mirror of the reference's Filter plugin interface
// (ekf_ws/src/localization_pkg/include/localization_pkg/filter.h:54-223) over the C-ABI of include/slam_filter.h.
//
// Same virtual names and call order as the reference; ROS types are replaced by plain data:
//   base_pkg::Command::ConstPtr            -> Command{fwd, ang}            (Command.msg:3-5, float32)
//   std_msgs::Float32MultiArray::ConstPtr  -> const std::vector<float>&    ([id, r, b]*, sim_node.py:245-250)
//   YAML::Node                             -> slam_params                  (the keys readCommonParams reads)
// Errors: where the reference throws std::runtime_error (filter.h:5,74,76; localization_node.cpp:44) so do these.
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/slam_filter.h"
#include "params_yaml.hpp"

namespace slam_host {

enum class FilterChoice { NOT_SET = 0, EKF_SLAM, UKF_LOC, UKF_SLAM, POSE_GRAPH_SLAM, NAIVE_COMMAND_PROPAGATION };  // filter.h:44-51

struct Command { float fwd = 0.f, ang = 0.f; };

// The state messages publishState() fills, field for field in wire order (ROS Header dropped):
//   base_pkg/msg/EKFState.msg:2-13, UKFState.msg:2-17 (X_pred is declared there but never filled: ukf.cpp:89,98,101),
//   NaiveState.msg:2-7.  float32 on the wire, like the reference's implicit double -> float casts (ekf.cpp:197-217).
struct EKFState {
    int32_t timestep = 0;
    float x_v = 0.f, y_v = 0.f, yaw_v = 0.f;
    int32_t M = 0;
    std::vector<float> landmarks;   // [id, x, y] * M  (ekf.cpp:203-207: 3M entries although the .msg comment says 2M)
    std::vector<float> P;           // (3+2M)^2 row by row (ekf.cpp:211-216)
};
struct UKFState {
    int32_t timestep = 0;
    float x_v = 0.f, y_v = 0.f, yaw_v = 0.f;
    int32_t M = 0;
    std::vector<float> landmarks;   // [id, x, y] * M  (ukf.cpp:74-80)
    std::vector<float> P;           // (4+2M)^2 row by row (ukf.cpp:82-88)
    std::vector<float> X;           // sigma points, one after the other (ukf.cpp:91-99)
    std::vector<float> X_pred;      // always empty (ukf.cpp:101 is commented out)
};
struct NaiveState {
    int32_t timestep = 0;
    float x_v = 0.f, y_v = 0.f, yaw_v = 0.f;
};

// BP/config/params.yaml defaults of the keys the hot path reads (:27-52)
inline slam_params default_params() {
    slam_params p{};
    p.v_d = 0.f; p.v_th = 0.f; p.w_r = 0.f; p.w_b = 0.f;
    p.V_00 = 0.01; p.V_11 = 0.001; p.W_00 = 0.01; p.W_11 = 0.01;
    p.landmark_id_is_known = 1; p.min_landmark_separation = 0.1f; p.compat_noise_bug = 1;
    p.d_max = 0.1; p.th_max = 0.0546; p.range_max = 3.0; p.fov_min = -1.57; p.fov_max = 1.57;
    return p;
}

class Filter {  // filter.h:54-145
public:
    FilterChoice type = FilterChoice::NOT_SET;
    bool isInit = false;                    // filter.h:68
    std::vector<float> map;                 // filter.h:69
    std::vector<int> lm_IDs;                // filter.h:70 (refreshed after every update)
    FilterChoice filter_to_compare = FilterChoice::NOT_SET;

    virtual ~Filter() { if (h_) slam_destroy(h_); }

    virtual void readParams(const slam_params& config) {    // filter.h:59,105-121
        params_ = config;
        if (h_) { slam_destroy(h_); h_ = nullptr; }
        if (slam_create(kind(), &params_, 1, max_landmarks_, max_meas_, device_, &h_) != 0)
            throw std::runtime_error(slam_last_error(nullptr));
    }
    virtual void readParams(const YamlNode& config) {        // filter.h:59 with the parsed params.yaml (YAML::Node in the reference)
        readParams(read_common_params(config));
    }
    virtual void init(float x_0, float y_0, float yaw_0) {   // filter.h:60
        need();
        check(slam_init(h_, x_0, y_0, yaw_0));
        isInit = true;
    }
    virtual void update(const Command& cmdMsg, const std::vector<float>& lmMeasMsg) {   // filter.h:61
        need();
        const int k = (int)(lmMeasMsg.size() / 3);           // ekf.cpp:65
        std::vector<float> buf((size_t)3 * max_meas_, 0.f);
        for (int i = 0; i < 3 * k && i < 3 * max_meas_; ++i) buf[i] = lmMeasMsg[i];
        check(slam_step(h_, &cmdMsg.fwd, &cmdMsg.ang, 0, buf.data(), &k));
        int st = 0;
        check(slam_get_status(h_, 0, &st));
        if (st & SLAM_STATUS_SAME_STEP_REMATCH) throw std::runtime_error("index >= 0 && index < size()");  // filter.h:5 + ekf.cpp:115
        if (st & SLAM_STATUS_MEAS_OVERFLOW) throw std::runtime_error("more detections in one message than max_meas");
        refresh_ids();
    }
    virtual void updateNaiveVehPoseEstimate(const std::vector<double>&, const std::vector<int>&) {          // filter.h:74
        throw std::runtime_error("updateNaiveVehPoseEstimate is not defined for this filter.");
    }
    virtual std::vector<double> getStateVector() {           // filter.h:76, ekf.cpp:182-185 (x, y, yaw, landmarks...)
        need();
        std::vector<double> x((size_t)4 + 2 * max_landmarks_);
        int n = 0;
        check(slam_get_state_vector(h_, 0, x.data(), &n));
        x.resize(n);
        return x;
    }
    std::vector<double> covariance() {                       // P_t, row-major n x n (ekf.cpp:211-217 order)
        need();
        const size_t nm = (size_t)4 + 2 * max_landmarks_;
        std::vector<double> P(nm * nm);
        int n = 0;
        check(slam_get_cov(h_, 0, P.data(), &n));
        P.resize((size_t)n * n);
        return P;
    }
    // filter.h:64-65: the reference advertises a ROS topic here; the headless equivalent registers the sink that receives
    // every published state (topic name kept for the shim: /state/ekf, /state/ukf, /state/naive)
    virtual void setupStatePublisher(std::function<void(const Filter&)> sink = nullptr) { statePub_ = std::move(sink); }
    virtual void publishState() = 0;                         // filter.h:66
    virtual std::string stateTopic() const = 0;
    std::vector<double> rawState() {                         // x_t as stored: EKF (x,y,yaw,lm..), UKF (x,y,cos,sin,lm..)
        need();
        std::vector<double> x((size_t)4 + 2 * max_landmarks_);
        int n = 0;
        check(slam_get_state(h_, 0, x.data(), &n));
        x.resize(n);
        return x;
    }
    int timestep() { need(); int t = 0; check(slam_get_timestep(h_, 0, &t)); return t; }
    int M() { need(); int m = 0; check(slam_get_num_landmarks(h_, 0, &m)); return m; }
    void setCapacity(int max_landmarks, int max_meas, int device = 0) { max_landmarks_ = max_landmarks; max_meas_ = max_meas; device_ = device; }

protected:
    virtual int kind() const = 0;
    void need() const { if (!h_) throw std::runtime_error("readParams must be called before the filter is used."); }
    void check(int rc) const { if (rc != 0) throw std::runtime_error(slam_last_error(h_)); }
    void refresh_ids() {
        std::vector<int> ids((size_t)max_landmarks_);
        int m = 0;
        check(slam_get_landmark_ids(h_, 0, ids.data(), &m));
        ids.resize(m);
        lm_IDs = ids;
    }
    template <class Msg>
    void fill_common(Msg& m, int base) {                     // the fields EKFState and UKFState share
        const std::vector<double> x = rawState();
        refresh_ids();
        m.timestep = timestep();
        m.x_v = (float)x[0]; m.y_v = (float)x[1];
        m.M = (int32_t)lm_IDs.size();
        m.landmarks.clear();
        for (int i = 0; i < m.M; ++i) {
            m.landmarks.push_back((float)lm_IDs[i]);
            m.landmarks.push_back((float)x[base + 2 * i]);
            m.landmarks.push_back((float)x[base + 2 * i + 1]);
        }
        const std::vector<double> P = covariance();
        m.P.assign(P.begin(), P.end());                      // double -> float, row by row
    }
    slam_handle_t h_ = nullptr;
    slam_params params_{};
    int max_landmarks_ = 50, max_meas_ = 16, device_ = 0;
    std::function<void(const Filter&)> statePub_;
};

class EKF : public Filter {   // filter.h:148-174
public:
    EKF() { type = FilterChoice::EKF_SLAM; }
    EKFState stateMsg;                                       // last published message
    std::string stateTopic() const override { return "/state/ekf"; }          // ekf.cpp:189
    void publishState() override {                           // ekf.cpp:192-220
        fill_common(stateMsg, 3);
        stateMsg.yaw_v = (float)rawState()[2];
        if (statePub_) statePub_(*this);
    }
protected:
    int kind() const override { return SLAM_EKF_SLAM; }
};

class UKF : public Filter {   // filter.h:177-223; `type` may be overridden to UKF_LOC before readParams (localization_node.cpp:36-38)
public:
    UKF() { type = FilterChoice::UKF_SLAM; }
    UKFState stateMsg;
    std::string stateTopic() const override { return "/state/ukf"; }          // ukf.cpp:57
    void publishState() override {                           // ukf.cpp:60-104
        fill_common(stateMsg, 4);
        const std::vector<double> x = rawState();
        stateMsg.yaw_v = (float)std::remainder(std::atan2(x[3], x[2]), 2 * 3.14159265358979323846);   // ukf.cpp:71
        const size_t nm = (size_t)4 + 2 * max_landmarks_;
        std::vector<double> X(nm * (2 * nm + 1));
        int n = 0;
        check(slam_get_sigma_points(h_, 0, X.data(), &n));
        stateMsg.X.assign(X.begin(), X.begin() + (size_t)n * (2 * n + 1));    // point by point, ukf.cpp:93-97
        stateMsg.X_pred.clear();
        if (statePub_) statePub_(*this);
    }
    // the /truth/landmarks message, float32 [id, x, y]*, kept in Filter::map (filter.h:68); read by the localisation-only mode
    void setMap(const std::vector<float>& landmarks) {
        map = landmarks;
        if (type == FilterChoice::UKF_LOC) { need(); check(slam_set_map(h_, map.data(), (int)(map.size() / 3))); }
    }
protected:
    int kind() const override { return type == FilterChoice::UKF_LOC ? SLAM_UKF_LOC : SLAM_UKF_SLAM; }
};

class NaiveFilter : public Filter {   // filter.h:325-370
public:
    NaiveFilter() { type = FilterChoice::NAIVE_COMMAND_PROPAGATION; }
    NaiveState stateMsg;
    std::string stateTopic() const override { return "/state/naive"; }        // filter.h:355
    void publishState() override {                           // filter.h:358-368
        const std::vector<double> x = rawState();
        stateMsg.timestep = timestep();
        stateMsg.x_v = (float)x[0]; stateMsg.y_v = (float)x[1]; stateMsg.yaw_v = (float)x[2];
        if (statePub_) statePub_(*this);
    }
protected:
    int kind() const override { return SLAM_NAIVE; }
};

// localization_node.cpp:33-45
inline std::unique_ptr<Filter> make_filter(const std::string& filter_choice_str) {
    if (filter_choice_str == "ekf_slam") return std::make_unique<EKF>();
    if (filter_choice_str == "ukf_slam") return std::make_unique<UKF>();
    if (filter_choice_str == "ukf_loc") {
        auto f = std::make_unique<UKF>();
        f->type = FilterChoice::UKF_LOC;                   // localization_node.cpp:38: override the default of UKF_SLAM
        return f;
    }
    if (filter_choice_str == "pose_graph")
        throw std::runtime_error("filter '" + filter_choice_str + "' is outside the B200 hot path (SURVEY.md 8f)");
    throw std::runtime_error("Invalid filter choice in params.yaml.");
}

}  // namespace slam_host
