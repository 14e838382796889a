// immesh_b200 -- IMU step in front of the hot path: the device bodies of ImuProcess::UndistortPcl
// (/root/reference/src/IMU_Processing.cpp:755-958).  Split of the work:
//   host   (imu_capi.cu)   : time stamps only -- which IMU intervals are live, their dt / offset, the scan-end dt and sign
//   imu_forward_step        : one IMU interval: bias / gravity-scale correction, 18x18 covariance propagation F P F^T + Q,
//                             attitude / velocity / position integration, IMUpose record (block-cooperative, tid / nthreads)
//   imu_predict_end         : state at the scan end (:881-896)
//   imu_compensate          : one point moved into the scan-end frame (:930-950), thread per point
// Same arithmetic order as oracle/orc_imu.hpp (left to right, no FMA): bit-identical results.
#pragma once
#include "hd_math.cuh"

namespace immesh {

#define IM_POSE_DOUBLES 22   // offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]  (Pose6D, common_lib.h)

struct ImuParams {
    double cov_gyr[3], cov_acc[3], cov_bias_gyr[3], cov_bias_acc[3];
    double mean_acc_norm;
    double lid_R[9], lid_T[3];
};
struct ImuStep {          // one live interval (head, tail) of v_imu
    double gyr_avg[3];    // 0.5 * (head + tail), raw
    double acc_avg[3];
    double dt;            // tail - head, or tail - last_lidar_end_time for the first live interval
    double offs_t;        // tail.stamp - pcl_beg_time
};
// running quantities of the propagation that are not part of the 348-double state: [acc_imu 3 | angvel_avr 3 | vel 3 | pos 3 | R 9]
#define IM_IMU_RUN 21

IM_HDN inline void imu_write_pose(double* pose, double t, const double* run) {
    pose[0] = t;
    for (int i = 0; i < 3; ++i) { pose[1 + i] = run[i]; pose[4 + i] = run[3 + i]; pose[7 + i] = run[6 + i]; pose[10 + i] = run[9 + i]; }
    for (int i = 0; i < 9; ++i) pose[13 + i] = run[12 + i];
}
// state: the 348-double StatesGroup image (rot 0, pos 9, vel 12, bg 15, ba 18, grav 21, cov 24).  Fx, T: 324-double scratch.
IM_HDN inline void imu_forward_step(const ImuParams& P, double* state, double* run, const ImuStep& s, double* pose_out, double* Fx, double* T, int tid, int nthreads) {
    double* cov = state + 24;
    const double dt = s.dt;
    if (tid == 0) {
        double angvel[3], acc_avr[3];
        for (int i = 0; i < 3; ++i) { angvel[i] = s.gyr_avg[i] - state[15 + i]; acc_avr[i] = s.acc_avg[i] * 9.81 / P.mean_acc_norm - state[18 + i]; }
        double En[9], Ask[9], RA[9];
        const double* R = run + 12;
        so3_exp_dt(angvel, -dt, En);
        skew3(acc_avr, Ask);
        m3_mul(R, Ask, RA);
        for (int i = 0; i < 324; ++i) Fx[i] = 0.0;
        for (int i = 0; i < 18; ++i) Fx[i * 18 + i] = 1.0;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                Fx[a * 18 + b] = En[a * 3 + b];
                Fx[a * 18 + 9 + b] = (a == b) ? -1.0 * dt : -0.0 * dt;
                Fx[(3 + a) * 18 + 6 + b] = (a == b) ? 1.0 * dt : 0.0 * dt;
                Fx[(6 + a) * 18 + b] = -RA[a * 3 + b] * dt;
                Fx[(6 + a) * 18 + 12 + b] = -R[a * 3 + b] * dt;
                Fx[(6 + a) * 18 + 15 + b] = (a == b) ? 1.0 * dt : 0.0 * dt;
            }
        for (int i = 0; i < 3; ++i) { run[3 + i] = angvel[i]; T[i] = acc_avr[i]; }   // T[0..2]: acc_avr handed to the last phase (T is rewritten below only after the barrier)
    }
    IM_SYNCBLOCK();
    double acc_keep[3] = {T[0], T[1], T[2]};
    IM_SYNCBLOCK();
    for (int idx = tid; idx < 324; idx += nthreads) {
        const int i = idx / 18, j = idx % 18;
        double sum = 0.0;
        for (int k = 0; k < 18; ++k) sum = sum + Fx[i * 18 + k] * cov[k * 18 + j];
        T[idx] = sum;
    }
    IM_SYNCBLOCK();
    {
        // cov_w: diag(cov_gyr) dt^2 at (0,0), R diag(cov_acc) R^T dt^2 at (6,6), diag(cov_bias_gyr) dt^2 at (9,9), diag(cov_bias_acc) dt^2 at (12,12)
        const double* R = run + 12;
        for (int idx = tid; idx < 324; idx += nthreads) {
            const int i = idx / 18, j = idx % 18;
            double sum = 0.0;
            for (int k = 0; k < 18; ++k) sum = sum + T[i * 18 + k] * Fx[j * 18 + k];
            double cw = 0.0;
            if (i < 3 && i == j) cw = P.cov_gyr[i] * dt * dt;
            else if (i >= 6 && i < 9 && j >= 6 && j < 9) {
                const int a = i - 6, b = j - 6;
                // ((R D) R^T)(a,b), products with the zero entries of D included as in the dense evaluation
                double rc[3];
                for (int q = 0; q < 3; ++q) rc[q] = (R[a * 3 + 0] * (q == 0 ? P.cov_acc[0] : 0.0) + R[a * 3 + 1] * (q == 1 ? P.cov_acc[1] : 0.0)) + R[a * 3 + 2] * (q == 2 ? P.cov_acc[2] : 0.0);
                cw = ((rc[0] * R[b * 3 + 0] + rc[1] * R[b * 3 + 1]) + rc[2] * R[b * 3 + 2]) * dt * dt;
            } else if (i >= 9 && i < 12 && i == j) cw = P.cov_bias_gyr[i - 9] * dt * dt;
            else if (i >= 12 && i < 15 && i == j) cw = P.cov_bias_acc[i - 12] * dt * dt;
            cov[idx] = sum + cw;
        }
    }
    IM_SYNCBLOCK();
    if (tid == 0) {
        double* acc_imu = run; double* angvel = run + 3; double* vel = run + 6; double* pos = run + 9; double* R = run + 12;
        double Ef[9], Rn[9], Ra[3];
        so3_exp_dt(angvel, dt, Ef);
        m3_mul(R, Ef, Rn);
        for (int i = 0; i < 9; ++i) R[i] = Rn[i];
        m3_vec(R, acc_keep, Ra);
        for (int i = 0; i < 3; ++i) acc_imu[i] = Ra[i] + state[21 + i];
        for (int i = 0; i < 3; ++i) pos[i] = (pos[i] + vel[i] * dt) + 0.5 * acc_imu[i] * dt * dt;
        for (int i = 0; i < 3; ++i) vel[i] = vel[i] + acc_imu[i] * dt;
        imu_write_pose(pose_out, s.offs_t, run);
    }
    IM_SYNCBLOCK();
}
// state at the scan end: vel_end, rot_end, pos_end  (note = +-1, dt already multiplied by note on the host, :881-896)
IM_HDN inline void imu_predict_end(double* state, const double* run, double note, double dt) {
    const double* acc_imu = run; const double* angvel = run + 3; const double* vel = run + 6; const double* pos = run + 9; const double* R = run + 12;
    const double w[3] = {note * angvel[0], note * angvel[1], note * angvel[2]};
    double E[9], Rn[9];
    for (int i = 0; i < 3; ++i) state[12 + i] = vel[i] + note * acc_imu[i] * dt;
    so3_exp_dt(w, dt, E);
    m3_mul(R, E, Rn);
    for (int i = 0; i < 9; ++i) state[i] = Rn[i];
    for (int i = 0; i < 3; ++i) state[9 + i] = (pos[i] + note * vel[i] * dt) + note * 0.5 * acc_imu[i] * dt * dt;
}
// one application of the backward compensation to point p = (x, y, z, curvature ms) with IMUpose `head`
IM_HDN inline void imu_compensate(const ImuParams& P, const double* state_end, const double* head, float* p) {
    const double dt = (double)p[3] / double(1000) - head[0];
    const double* acc = head + 1; const double* gyr = head + 4; const double* vel = head + 7; const double* pos = head + 10; const double* rot = head + 13;
    double E[9], R_i[9], T_ei[3];
    so3_exp_dt(gyr, dt, E);
    m3_mul(rot, E, R_i);
    for (int i = 0; i < 3; ++i) T_ei[i] = ((pos[i] + vel[i] * dt) + 0.5 * acc[i] * dt * dt) - state_end[9 + i];
    const double P_i[3] = {(double)p[0], (double)p[1], (double)p[2]};
    double a[3], b[3], c[3], d[3];
    m3_vec(P.lid_R, P_i, a);
    for (int i = 0; i < 3; ++i) a[i] = a[i] + P.lid_T[i];
    m3_vec(R_i, a, b);
    for (int i = 0; i < 3; ++i) b[i] = b[i] + T_ei[i];
    for (int i = 0; i < 3; ++i) c[i] = ((state_end[0 * 3 + i] * b[0] + state_end[1 * 3 + i] * b[1]) + state_end[2 * 3 + i] * b[2]) - P.lid_T[i];
    for (int i = 0; i < 3; ++i) d[i] = (P.lid_R[0 * 3 + i] * c[0] + P.lid_R[1 * 3 + i] * c[1]) + P.lid_R[2 * 3 + i] * c[2];
    p[0] = (float)d[0]; p[1] = (float)d[1]; p[2] = (float)d[2];
}
// the point of sorted index s: interval = last head whose offset is < its time; sorted index 0 is re-compensated by every earlier
// head the reference's backward walk still visits (see oracle/orc_imu.hpp)
IM_HDN inline void imu_undistort_point(const ImuParams& P, const double* state_end, const double* poses, int n_pose, float* pts, int s) {
    float* p = pts + 4 * (size_t)s;
    const double t = (double)p[3] / double(1000);
    int j = n_pose - 2;
    while (j >= 0 && !(t > poses[(size_t)j * IM_POSE_DOUBLES])) --j;
    if (j < 0) return;
    imu_compensate(P, state_end, poses + (size_t)j * IM_POSE_DOUBLES, p);
    if (s == 0) {
        for (int q = j - 1; q >= 0; --q)
            if (t > poses[(size_t)q * IM_POSE_DOUBLES]) imu_compensate(P, state_end, poses + (size_t)q * IM_POSE_DOUBLES, p);
    }
}

}  // namespace immesh
