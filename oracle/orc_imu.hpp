// ORACLE (test infrastructure, NOT product code): CPU restatement of the IMU step in FRONT of the hot path:
// ImuProcess::UndistortPcl (/root/reference/src/IMU_Processing.cpp:755-958) -- forward propagation of the state and its 18x18
// covariance over the IMU samples of a scan, prediction to the scan end, and the backward per-point motion compensation.
// LiDAR-only flow (lidar_meas.is_lidar_end == true, no camera frames).  Quirks kept on purpose ("replicate, don't fix"):
//   * the scan is time-sorted first (:785; std::sort on curvature -- order of equal stamps unspecified there, STABLE here and
//     in the CUDA path), pcl_end_time comes from the LAST point of the UNSORTED cloud (:786);
//   * a point belongs to the last IMU interval whose head offset is < its time (the backward walk :921-957); points at
//     t <= 0 stay untouched;
//   * the first point of the sorted cloud is compensated again by every earlier interval the walk still visits
//     (the `if (it_pcl == begin) break` at :954 does not advance the iterator).
// PARITY UNPINNED (Eigen expression order restated left to right; sin/cos are the arithmetic-only versions of orc_math.hpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "orc_lio.hpp"

namespace orc {

struct ImuSample { double t, gyr[3], acc[3]; };
struct Pose6D { double offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]; };

struct ImuOracle {
    // ImuProcess members (IMU_Processing.cpp:50-66, set_* :108-140)
    double cov_gyr[3] = {0.1, 0.1, 0.1}, cov_acc[3] = {0.1, 0.1, 0.1}, cov_bias_gyr[3] = {0.1, 0.1, 0.1}, cov_bias_acc[3] = {0.1, 0.1, 0.1};
    double mean_acc_norm = 9.81;
    double lid_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, lid_T[3] = {0, 0, 0};
    double acc_s_last[3] = {0, 0, 0}, angvel_last[3] = {0, 0, 0};
    ImuSample last_imu{};
    double last_lidar_end_time = -1.0, last_update_time = 0.0;
    std::vector<Pose6D> IMUpose;

    static Pose6D set_pose6d(double t, const double* a, const double* g, const double* v, const double* p, const double* R) {
        Pose6D k;
        k.offset_time = t;
        for (int i = 0; i < 3; ++i) { k.acc[i] = a[i]; k.gyr[i] = g[i]; k.vel[i] = v[i]; k.pos[i] = p[i]; }
        for (int i = 0; i < 9; ++i) k.rot[i] = R[i];
        return k;
    }

    // pts: [n][4] x, y, z, curvature (ms since lidar_beg_time), float like pcl::PointXYZINormal; in place: sorted + compensated
    void undistort_pcl(State& st, const std::vector<ImuSample>& meas_imu, float* pts, int n, double lidar_beg_time) {
        std::vector<ImuSample> v_imu;
        v_imu.push_back(last_imu);
        v_imu.insert(v_imu.end(), meas_imu.begin(), meas_imu.end());
        const double imu_end_time = v_imu.back().t;
        const double pcl_beg_time = std::max(lidar_beg_time, last_update_time);
        const double pcl_end_time = lidar_beg_time + (double)pts[4 * (size_t)(n - 1) + 3] / double(1000);   // unsorted .back()
        {   // sort(points, time_list)
            std::vector<int> order(n);
            for (int i = 0; i < n; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pts[4 * (size_t)a + 3] < pts[4 * (size_t)b + 3]; });
            std::vector<float> tmp(pts, pts + 4 * (size_t)n);
            for (int i = 0; i < n; ++i)
                for (int c = 0; c < 4; ++c) pts[4 * (size_t)i + c] = tmp[4 * (size_t)order[i] + c];
        }
        last_update_time = pcl_end_time;
        IMUpose.clear();
        IMUpose.push_back(set_pose6d(0.0, acc_s_last, angvel_last, st.vel, st.pos, st.rot));
        double acc_imu[3] = {acc_s_last[0], acc_s_last[1], acc_s_last[2]}, angvel_avr[3] = {angvel_last[0], angvel_last[1], angvel_last[2]}, acc_avr[3];
        double vel_imu[3] = {st.vel[0], st.vel[1], st.vel[2]}, pos_imu[3] = {st.pos[0], st.pos[1], st.pos[2]}, R_imu[9];
        for (int i = 0; i < 9; ++i) R_imu[i] = st.rot[i];
        double dt = 0;
        for (size_t k = 0; k + 1 < v_imu.size(); ++k) {
            const ImuSample &head = v_imu[k], &tail = v_imu[k + 1];
            if (tail.t < last_lidar_end_time) continue;
            for (int i = 0; i < 3; ++i) { angvel_avr[i] = 0.5 * (head.gyr[i] + tail.gyr[i]); acc_avr[i] = 0.5 * (head.acc[i] + tail.acc[i]); }
            for (int i = 0; i < 3; ++i) { angvel_avr[i] = angvel_avr[i] - st.bg[i]; acc_avr[i] = acc_avr[i] * 9.81 / mean_acc_norm - st.ba[i]; }
            dt = (head.t < last_lidar_end_time) ? tail.t - last_lidar_end_time : tail.t - head.t;
            double Exp_f[9], En[9], Ask[9], RA[9];
            so3_exp_dt(angvel_avr, dt, Exp_f);
            skew(acc_avr, Ask);
            std::vector<double> Fx(324, 0.0), cw(324, 0.0);
            for (int i = 0; i < 18; ++i) Fx[i * 18 + i] = 1.0;
            so3_exp_dt(angvel_avr, -dt, En);
            mat3_mul(R_imu, Ask, RA);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    Fx[a * 18 + b] = En[a * 3 + b];
                    Fx[a * 18 + 9 + b] = (a == b) ? -1.0 * dt : -0.0 * dt;            // -Eye3d * dt
                    Fx[(3 + a) * 18 + 6 + b] = (a == b) ? 1.0 * dt : 0.0 * dt;         //  Eye3d * dt
                    Fx[(6 + a) * 18 + b] = -RA[a * 3 + b] * dt;                       // -R_imu * acc_avr_skew * dt
                    Fx[(6 + a) * 18 + 12 + b] = -R_imu[a * 3 + b] * dt;               // -R_imu * dt
                    Fx[(6 + a) * 18 + 15 + b] = (a == b) ? 1.0 * dt : 0.0 * dt;        //  Eye3d * dt
                }
            double RC[9], RCR[9], Cd[9] = {cov_acc[0], 0, 0, 0, cov_acc[1], 0, 0, 0, cov_acc[2]};
            mat3_mul(R_imu, Cd, RC);
            mat3_mul_bt(RC, R_imu, RCR);
            for (int a = 0; a < 3; ++a) {
                cw[a * 18 + a] = cov_gyr[a] * dt * dt;
                for (int b = 0; b < 3; ++b) cw[(6 + a) * 18 + 6 + b] = RCR[a * 3 + b] * dt * dt;
                cw[(9 + a) * 18 + 9 + a] = cov_bias_gyr[a] * dt * dt;
                cw[(12 + a) * 18 + 12 + a] = cov_bias_acc[a] * dt * dt;
            }
            std::vector<double> T(324), nc(324);
            for (int i = 0; i < 18; ++i)
                for (int j = 0; j < 18; ++j) {
                    double s = 0.0;
                    for (int q = 0; q < 18; ++q) s = s + Fx[i * 18 + q] * st.cov[q * 18 + j];
                    T[i * 18 + j] = s;
                }
            for (int i = 0; i < 18; ++i)
                for (int j = 0; j < 18; ++j) {
                    double s = 0.0;
                    for (int q = 0; q < 18; ++q) s = s + T[i * 18 + q] * Fx[j * 18 + q];
                    nc[i * 18 + j] = s + cw[i * 18 + j];
                }
            for (int i = 0; i < 324; ++i) st.cov[i] = nc[i];
            double Rn[9];
            mat3_mul(R_imu, Exp_f, Rn);
            for (int i = 0; i < 9; ++i) R_imu[i] = Rn[i];
            double Ra[3];
            mat3_vec(R_imu, acc_avr, Ra);
            for (int i = 0; i < 3; ++i) acc_imu[i] = Ra[i] + st.grav[i];
            for (int i = 0; i < 3; ++i) pos_imu[i] = (pos_imu[i] + vel_imu[i] * dt) + 0.5 * acc_imu[i] * dt * dt;
            for (int i = 0; i < 3; ++i) vel_imu[i] = vel_imu[i] + acc_imu[i] * dt;
            for (int i = 0; i < 3; ++i) { angvel_last[i] = angvel_avr[i]; acc_s_last[i] = acc_imu[i]; }
            IMUpose.push_back(set_pose6d(tail.t - pcl_beg_time, acc_imu, angvel_avr, vel_imu, pos_imu, R_imu));
        }
        {   // prediction at the frame end (:881-896)
            double note;
            if (imu_end_time > pcl_beg_time) { note = pcl_end_time > imu_end_time ? 1.0 : -1.0; dt = note * (pcl_end_time - imu_end_time); }
            else { note = pcl_end_time > pcl_beg_time ? 1.0 : -1.0; dt = note * (pcl_end_time - pcl_beg_time); }
            double w[3] = {note * angvel_avr[0], note * angvel_avr[1], note * angvel_avr[2]}, E[9], Rn[9];
            for (int i = 0; i < 3; ++i) st.vel[i] = vel_imu[i] + note * acc_imu[i] * dt;
            so3_exp_dt(w, dt, E);
            mat3_mul(R_imu, E, Rn);
            for (int i = 0; i < 9; ++i) st.rot[i] = Rn[i];
            for (int i = 0; i < 3; ++i) st.pos[i] = (pos_imu[i] + note * vel_imu[i] * dt) + note * 0.5 * acc_imu[i] * dt * dt;
        }
        last_imu = v_imu.back();
        last_lidar_end_time = pcl_end_time;
        if (n < 1) return;
        // backward compensation (:914-957), literal walk
        int it = n - 1;
        for (int kp = (int)IMUpose.size() - 1; kp >= 1; --kp) {
            const Pose6D& head = IMUpose[kp - 1];
            for (; (double)pts[4 * (size_t)it + 3] / double(1000) > head.offset_time; --it) {
                compensate(st, head, pts + 4 * (size_t)it);
                if (it == 0) break;
            }
        }
    }
    void compensate(const State& st, const Pose6D& head, float* p) const {
        const double dt = (double)p[3] / double(1000) - head.offset_time;
        double E[9], R_i[9];
        so3_exp_dt(head.gyr, dt, E);
        mat3_mul(head.rot, E, R_i);
        double T_ei[3];
        for (int i = 0; i < 3; ++i) T_ei[i] = ((head.pos[i] + head.vel[i] * dt) + 0.5 * head.acc[i] * dt * dt) - st.pos[i];
        const double P_i[3] = {(double)p[0], (double)p[1], (double)p[2]};
        double a[3], b[3], c[3], d[3];
        mat3_vec(lid_R, P_i, a);
        for (int i = 0; i < 3; ++i) a[i] = a[i] + lid_T[i];
        mat3_vec(R_i, a, b);
        for (int i = 0; i < 3; ++i) b[i] = b[i] + T_ei[i];
        mat3_tvec(st.rot, b, c);
        for (int i = 0; i < 3; ++i) c[i] = c[i] - lid_T[i];
        mat3_tvec(lid_R, c, d);
        p[0] = (float)d[0]; p[1] = (float)d[1]; p[2] = (float)d[2];
    }
};

}  // namespace orc
