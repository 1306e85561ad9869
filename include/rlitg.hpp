// rlitg.hpp -- C++17 host-side mirror of the reference's C++ seams over the C ABI of librlitg.so (rlitg.h).
//
// The reference's planner is compiled code: `torm::TormIKSolver` (torm/include/torm/torm_ik_solver.h:30-33) and
// `torm::TormOptimizer` (torm/include/torm/torm_optimizer.h:79-100,296-312).  Their KDL / Eigen / MoveIt value types
// do not exist in this build, so the classes below keep the reference's MEMBER NAMES and argument meaning on plain
// structs that read like the KDL ones (Frame{M, p}, Twist{vel, rot}, JntArray(i), Jacobian(r, c)).  Header only;
// needs nothing but librlitg.so (no CUDA toolkit: device memory goes through rlitg_dev_* / rlitg_copy_*).
//
//   rlitg::TormIKSolver         fkSolver / allfkSolver / vikSolver / getJacobian / collisionChecking (batch of one,
//                               FP64 like KDL) + batched overloads on flat std::vector<double>
//   rlitg::Scene                sphere table + occupancy grid handle (RAII)
//   rlitg::TormCost             smoothness operator handle (RAII), getMaxQuadCostInvValue()
//   rlitg::PirlVecEnv           B PIRL environments per step (PirlEnv_Fetch.reset / step), pirl_interpolate()
//   rlitg::TormBatchOptimizer   N trajectories at once: getEndPoseCost, calculate*Increments, getCollisionCost,
//                               calculateFeasibleIncrements / addIncrementsToTrajectory / handleJointLimits (fused in
//                               the update call, as on the device), localOptimizeTSGD
// Errors: every failing C call throws rlitg::Error carrying the code and rlitg_last_error().
#ifndef RLITG_HPP_
#define RLITG_HPP_

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <random>
#include <vector>

#include "rlitg.h"

namespace rlitg {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != RLITG_OK) throw Error(rc, rlitg_last_error());
}

// ---- device buffer (RAII over rlitg_dev_alloc) ---------------------------------------------------------------
template <class T>
class DeviceArray {
  public:
    DeviceArray() = default;
    explicit DeviceArray(size_t n) { resize(n); }
    DeviceArray(const DeviceArray&) = delete;
    DeviceArray& operator=(const DeviceArray&) = delete;
    DeviceArray(DeviceArray&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    DeviceArray& operator=(DeviceArray&& o) noexcept {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~DeviceArray() { release(); }
    void resize(size_t n) {
        if (n == n_) return;
        release();
        void* p = nullptr;
        check(rlitg_dev_alloc(&p, (uint64_t)(n * sizeof(T))));
        p_ = static_cast<T*>(p);
        n_ = n;
    }
    void upload(const T* host, size_t n, void* stream = nullptr) { check(rlitg_copy_h2d(p_, host, n * sizeof(T), stream)); }
    void upload(const std::vector<T>& host, void* stream = nullptr) {
        if (host.size() > n_) resize(host.size());
        upload(host.data(), host.size(), stream);
    }
    void download(T* host, size_t n, void* stream = nullptr) const {
        check(rlitg_copy_d2h(host, p_, n * sizeof(T), stream));
        check(rlitg_stream_sync(stream));
    }
    std::vector<T> download(void* stream = nullptr) const {
        std::vector<T> h(n_);
        if (n_) download(h.data(), n_, stream);
        return h;
    }
    void zero(void* stream = nullptr) { check(rlitg_dev_zero(p_, n_ * sizeof(T), stream)); }
    T* data() const { return p_; }
    size_t size() const { return n_; }

  private:
    void release() {
        if (p_) rlitg_dev_free(p_);
        p_ = nullptr;
        n_ = 0;
    }
    T* p_ = nullptr;
    size_t n_ = 0;
};

// ---- value types that read like the KDL ones -------------------------------------------------------------------
struct Vector {
    double data[3] = {0, 0, 0};
    double& operator[](int i) { return data[i]; }
    double operator[](int i) const { return data[i]; }
    double x() const { return data[0]; }
    double y() const { return data[1]; }
    double z() const { return data[2]; }
    double Norm() const { return std::sqrt(data[0] * data[0] + data[1] * data[1] + data[2] * data[2]); }
};
struct Rotation {  // row-major like KDL::Rotation::data
    double data[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double operator()(int i, int j) const { return data[3 * i + j]; }
    double& operator()(int i, int j) { return data[3 * i + j]; }
    // KDL::Rotation::Quaternion(x, y, z, w) (unnormalised input is used as is)
    static Rotation Quaternion(double x, double y, double z, double w) {
        Rotation R;
        const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        R.data[0] = w2 + x2 - y2 - z2; R.data[1] = 2 * x * y - 2 * w * z; R.data[2] = 2 * x * z + 2 * w * y;
        R.data[3] = 2 * x * y + 2 * w * z; R.data[4] = w2 - x2 + y2 - z2; R.data[5] = 2 * y * z - 2 * w * x;
        R.data[6] = 2 * x * z - 2 * w * y; R.data[7] = 2 * y * z + 2 * w * x; R.data[8] = w2 - x2 - y2 + z2;
        return R;
    }
};
struct Frame {
    Rotation M;
    Vector p;
};
struct Twist {
    Vector vel, rot;
};
struct JntArray {
    std::vector<double> data;
    explicit JntArray(unsigned n = 7) : data(n, 0.0) {}
    double& operator()(unsigned i) { return data[i]; }
    double operator()(unsigned i) const { return data[i]; }
    unsigned rows() const { return (unsigned)data.size(); }
};
struct Jacobian {  // 6 x n, rows 0-2 linear, 3-5 angular (KDL::Jacobian)
    std::vector<double> data;  // row-major [6][n]
    unsigned n;
    explicit Jacobian(unsigned n_ = 7) : data(6 * n_, 0.0), n(n_) {}
    double operator()(unsigned r, unsigned c) const { return data[r * n + c]; }
    unsigned columns() const { return n; }
};

// ---- Scene: sphere table + occupancy grid ----------------------------------------------------------------------
class Scene {
  public:
    Scene() { check(rlitg_scene_create(&h_)); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    ~Scene() {
        if (h_) rlitg_scene_destroy(h_);
    }
    void setSpheres(const std::vector<int32_t>& link, const std::vector<double>& center, const std::vector<double>& radius) {
        check(rlitg_scene_set_spheres(h_, (int)link.size(), link.data(), center.data(), radius.data()));
    }
    // occ [nz][ny][nx]; origin = centre of cell (0,0,0) in the planning frame
    void setOccupancy(const std::vector<uint8_t>& occ, int nx, int ny, int nz, const std::array<double, 3>& origin,
                      double res, double max_prop = 1.0) {
        if (occ.size() != (size_t)nx * ny * nz) throw Error(RLITG_ERR_SHAPE, "occ.size() != nx*ny*nz");
        check(rlitg_scene_set_occupancy(h_, occ.data(), nx, ny, nz, origin.data(), res, max_prop));
    }
    RlitgScene* handle() const { return h_; }

  private:
    RlitgScene* h_ = nullptr;
};

// ---- TormIKSolver ----------------------------------------------------------------------------------------------
// torm/include/torm/torm_ik_solver.h:30-33 (+ collisionChecking :24).  FP64 like KDL.  The single-configuration
// members cost one launch + two small copies each; planners should prefer the batched overloads / the optimiser.
class TormIKSolver {
  public:
    explicit TormIKSolver(const Scene* scene = nullptr, std::array<double, 3> base_off = {-0.086875, 0.0, 0.37743})
        : scene_(scene), base_off_(base_off) {}
    unsigned getDoF() const { return 7; }

    void fkSolver(const JntArray& q_init, Frame& p_in) {
        std::vector<double> pos, rot;
        fkSolver(q_init.data, pos, rot);
        std::memcpy(p_in.p.data, pos.data(), 3 * sizeof(double));
        std::memcpy(p_in.M.data, rot.data(), 9 * sizeof(double));
    }
    void allfkSolver(const JntArray& q_init, std::vector<Frame>& p_in, int n_segments = -1) {
        const unsigned D = q_init.rows();
        stage(q_init.data);
        out_.resize((size_t)(D + 1) * 7);
        check(rlitg_allfk(q_.data(), 1, (int)D, RLITG_F64, 0, out_.data(), nullptr));
        std::vector<double> h((size_t)(D + 1) * 7);
        out_.download(h.data(), h.size());
        const int n = n_segments < 0 ? (int)D + 1 : n_segments;
        p_in.resize(n);
        for (int s = 0; s < n; s++) {
            const double* r = &h[7 * s];
            std::memcpy(p_in[s].p.data, r, 3 * sizeof(double));
            p_in[s].M = Rotation::Quaternion(r[3], r[4], r[5], r[6]);
        }
    }
    void vikSolver(const JntArray& q, const Twist& delta_twist, JntArray& delta_q) {
        const unsigned D = q.rows();
        stage(q.data);
        const double tw[6] = {delta_twist.vel[0], delta_twist.vel[1], delta_twist.vel[2],
                              delta_twist.rot[0], delta_twist.rot[1], delta_twist.rot[2]};
        tw_.resize(6);
        tw_.upload(tw, 6);
        out_.resize(D);
        check(rlitg_pinv_step(q_.data(), tw_.data(), 1, (int)D, RLITG_F64, 1e-5, out_.data(), nullptr));
        delta_q.data.resize(D);
        out_.download(delta_q.data.data(), D);
    }
    void getJacobian(const JntArray& q, Jacobian& jac) {
        std::vector<double> J;
        getJacobian(q.data, J);
        jac = Jacobian(q.rows());
        jac.data = J;
    }
    // TormIKSolver::collisionChecking(values) (:24): sphere-vs-grid approximation of the FCL test
    bool collisionChecking(const std::vector<double>& values) {
        if (!scene_) throw Error(RLITG_ERR_NULL, "TormIKSolver was built without a scene");
        return collisionChecking(values, 7)[0] != 0;
    }

    // ---- batched forms: q flat [B*D] ----
    void fkSolver(const std::vector<double>& q, std::vector<double>& pos, std::vector<double>& rot, int D = 7) {
        const size_t B = q.size() / D;
        stage(q);
        DeviceArray<double> p(3 * B), r(9 * B);
        check(rlitg_fk_jac(q_.data(), (int64_t)B, D, RLITG_F64, 0, p.data(), r.data(), nullptr, nullptr, nullptr));
        pos = p.download();
        rot = r.download();
    }
    void getJacobian(const std::vector<double>& q, std::vector<double>& J, int D = 7) {
        const size_t B = q.size() / D;
        stage(q);
        DeviceArray<double> j(6 * D * B);
        check(rlitg_jacobian(q_.data(), (int64_t)B, D, RLITG_F64, 0, j.data(), nullptr));
        J = j.download();
    }
    std::vector<uint8_t> collisionChecking(const std::vector<double>& q, int D) {
        if (!scene_) throw Error(RLITG_ERR_NULL, "TormIKSolver was built without a scene");
        const size_t B = q.size() / D;
        stage(q);
        DeviceArray<uint8_t> o(B);
        check(rlitg_sphere_grid_collide(q_.data(), (int64_t)B, RLITG_F64, scene_->handle(), base_off_.data(), o.data(),
                                        nullptr));
        return o.download();
    }

  private:
    void stage(const std::vector<double>& q) {
        if (q.size() > q_.size()) q_.resize(q.size());
        q_.upload(q.data(), q.size());
    }
    const Scene* scene_;
    std::array<double, 3> base_off_;
    DeviceArray<double> q_, tw_, out_;
};

// ---- TormCost: smoothness operator -----------------------------------------------------------------------------
class TormCost {  // torm/include/torm/torm_cost.h + the scaling of torm/src/torm_optimizer.cpp:211-236
  public:
    TormCost(int T, double dt, std::array<double, 3> weights = {1.0, 0.0, 0.0}, double ridge = 0.0) : T_(T) {
        check(rlitg_smooth_create(T, dt, weights.data(), ridge, &h_));
    }
    TormCost(const TormCost&) = delete;
    TormCost& operator=(const TormCost&) = delete;
    ~TormCost() {
        if (h_) rlitg_smooth_destroy(h_);
    }
    double getMaxQuadCostInvValue() const {  // before scaling; the reference's `// 96.77624` (:235)
        std::vector<double> band((size_t)(T_ + 10) * 13);
        double scale = 0;
        check(rlitg_smooth_get(h_, band.data(), &scale, nullptr));
        return scale;
    }
    RlitgSmooth* handle() const { return h_; }

  private:
    RlitgSmooth* h_ = nullptr;
    int T_;
};

// ---- parameters (torm/include/torm/torm_parameters.h:53-85, defaults of torm/src/torm_main.cpp:50-92) ----------
struct TormParameters {
    double smoothness_update_weight_ = 30.0;  // the mains divide by T
    double obstacle_update_weight_ = 1.0;
    double jacobian_update_weight_ = 1.0;
    double learning_rate_ = 0.01;
    double rotation_scale_factor_ = 0.17;
    double smoothness_cost_velocity_ = 1.0, smoothness_cost_acceleration_ = 0.0, smoothness_cost_jerk_ = 0.0;
    bool use_velocity_check_ = true, use_collision_check_ = true, use_pseudo_inverse_ = true;
    bool use_stochastic_descent_ = true;  // true in every reference main
    double ridge_factor_ = 0.0;
    double pseudo_inverse_ridge_factor_ = 1e-4;
    double joint_update_limit_ = 0.17;
    double min_clearence_ = 0.001;
    double time_duration_ = 0.1;
    bool affine_axis_quirk = true;
    std::array<double, 3> base_offset = {-0.086875, 0.0, 0.37743};
    std::array<double, 7> joint_min = {-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592};
    std::array<double, 7> joint_max = {1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592};
    std::array<int, 7> continuous = {0, 0, 1, 0, 1, 0, 1};
    std::array<double, 7> vel_limit = {1.256, 1.454, 1.571, 1.521, 1.571, 2.268, 2.268};

    RlitgTormParams to_c(int T) const {
        RlitgTormParams p;
        std::memset(&p, 0, sizeof(p));
        p.T = T; p.D = 7;
        p.dt = time_duration_; p.rot_scale = rotation_scale_factor_; p.jac_weight = jacobian_update_weight_;
        p.clearance = min_clearence_; p.pinv_ridge = pseudo_inverse_ridge_factor_; p.vik_eps = 1e-5;
        p.use_pinv = use_pseudo_inverse_; p.affine_axis_quirk = affine_axis_quirk;
        for (int k = 0; k < 3; k++) p.base_off[k] = base_offset[k];
        for (int k = 0; k < 7; k++) { p.cjoint[k] = continuous[k]; p.vel_limit[k] = vel_limit[k]; }
        return p;
    }
};

// ---- TormBatchOptimizer ----------------------------------------------------------------------------------------
// The per-waypoint loops of torm::TormOptimizer (torm/src/torm_optimizer.cpp:1399-1639) for N candidate
// trajectories at once.  Trajectories live on the device as float32 [N,T,7] (row 0 = start configuration, rows
// 1..T-1 free = group_trajectory_ rows start_idx-1 .. end_idx); the host exchanges them as double like the
// reference's Eigen matrices.  Member names follow the reference; each returns one value PER trajectory.
class TormBatchOptimizer {
  public:
    TormBatchOptimizer(int64_t N, int T, const std::vector<double>& end_poses /*[T*7] x y z qx qy qz qw*/,
                       const TormParameters& parameters = TormParameters(), const Scene* scene = nullptr)
        : N_(N), T_(T), par_(parameters), scene_(scene),
          joint_costs_(T, parameters.time_duration_,
                       {parameters.smoothness_cost_velocity_, parameters.smoothness_cost_acceleration_,
                        parameters.smoothness_cost_jerk_},
                       parameters.ridge_factor_),
          cp_(parameters.to_c(T)) {
        if (end_poses.size() != (size_t)T * 7) throw Error(RLITG_ERR_SHAPE, "end_poses must hold T*7 values");
        std::vector<float> t(end_poses.begin(), end_poses.end());
        targets_.upload(t);
        const size_t inc = (size_t)N * (T - 1) * 7;
        traj_.resize((size_t)N * T * 7);
        best_.resize((size_t)N * T * 7);
        smoothness_increments_.resize(inc);
        collision_increments_.resize(inc);
        endPose_increments_.resize(inc);
        collision_increments_.zero();
        costs_.resize((size_t)N * RLITG_TORM_NCOST);
        alpha_.upload(std::vector<float>((size_t)N, 0.5f));
        lr_.upload(std::vector<float>((size_t)N, (float)parameters.learning_rate_));
        hits_.resize((size_t)N * T);
    }
    int64_t numTrajectories() const { return N_; }
    int numPoints() const { return T_; }

    void setTrajectories(const std::vector<double>& traj /*[N*T*7]*/) {
        if (traj.size() != (size_t)N_ * T_ * 7) throw Error(RLITG_ERR_SHAPE, "traj must hold N*T*7 values");
        std::vector<float> t(traj.begin(), traj.end());
        traj_.upload(t);
    }
    std::vector<double> getTrajectories() const { return widen(traj_.download()); }
    std::vector<double> getBestTrajectories() const { return widen(best_.download()); }

    // getEndPoseCost(grad) (:1399-1425): sum of |dp| + rotation_scale_factor_ * sum of |drot| over the free
    // waypoints; grad = true also fills endPose_increments_ (calculateEndPoseIncrements :1475-1482).
    std::vector<double> getEndPoseCost(bool grad) {
        run(grad ? RLITG_STAGE_POSE_GRAD : RLITG_STAGE_POSE);
        const auto c = costs();
        std::vector<double> out(N_);
        for (int64_t n = 0; n < N_; n++) out[n] = c[n * RLITG_TORM_NCOST + 8] + par_.rotation_scale_factor_ * c[n * RLITG_TORM_NCOST + 9];
        return out;
    }
    void calculateEndPoseIncrements() { run(RLITG_STAGE_POSE_GRAD); }
    void calculateSmoothnessIncrements() { run(RLITG_STAGE_SMOOTH); }
    // performForwardKinematics + calculateCollisionIncrements (:1302-1382, :1484-1553)
    void calculateCollisionIncrements(const int32_t* stochastic_idx_dev = nullptr) { run(RLITG_STAGE_COLL, stochastic_idx_dev); }
    std::vector<double> getCollisionCost() {  // :1448-1463
        run(RLITG_STAGE_COLL);
        return column(5);
    }
    // traj_evaluator::getCost (torm/src/utils/traj_evaluator.cpp:104-132): rows of [pos, rot, vel, acc, jerk, coll,
    // weighted, n_vel_violations, end-pose pos sum, end-pose rot sum]
    std::vector<double> evaluate() {
        run(RLITG_STAGE_EVAL | RLITG_STAGE_POSE | (scene_ ? RLITG_STAGE_COLL : 0));
        return widen(costs());
    }
    // calculateFeasibleIncrements (:1555-1609) + addIncrementsToTrajectory (:1611-1639) + handleJointLimits
    // (:1189-1268), from the increments the last calculate* calls left on the device
    void updateFeasible(int lr_schedule_mode = 3) { update(0, lr_schedule_mode); }
    void updateEndPose(int lr_schedule_mode = 3) { update(1, lr_schedule_mode); }
    void setAlpha(float a) { alpha_.upload(std::vector<float>((size_t)N_, a)); }
    void setLearningRate(float lr) { lr_.upload(std::vector<float>((size_t)N_, lr)); }

    // TormOptimizer::localOptimizeTSGD (:397-598) in lock step over the N trajectories, entirely on the device: per
    // half-iteration the collision-point test (isCurrentTrajectoryCollisionFree :695-717 + the waypoint choice of
    // use_stochastic_descent_ :1502-1515, draws from a host mt19937 seeded with `seed`), one cost call,
    // rlitg_torm_bookkeep (stage choice of lr_schedule_mode_ 1, alpha_ / learning_rate_ schedule, best-trajectory record
    // of :506-560) and the update step.
    // Returns the best cost per trajectory (+inf if never feasible); getBestTrajectories() holds the rows.
    std::vector<double> localOptimizeTSGD(int maxiter, int lr_schedule_mode = 3, unsigned seed = 0) {
        const float inf = std::numeric_limits<float>::infinity();
        best_cost_.upload(std::vector<float>((size_t)N_, inf));
        check(rlitg_copy_d2d(best_.data(), traj_.data(), traj_.size() * sizeof(float), nullptr));
        alpha_.upload(std::vector<float>((size_t)N_, 0.5f));
        lr_.upload(std::vector<float>((size_t)N_, (float)par_.learning_rate_));
        infea_col_.resize((size_t)N_); infea_col_.zero();
        infea_vel_.resize((size_t)N_); infea_vel_.zero();
        stage_idx_.resize((size_t)N_);
        lr_iter_.resize(1);
        const bool check_col = scene_ && par_.use_collision_check_;
        const bool stochastic = scene_ && par_.use_stochastic_descent_;
        std::mt19937 rng(seed);
        std::uniform_real_distribution<float> uni(0.0f, 1.0f);
        std::vector<float> u_host((size_t)N_);
        if (check_col || stochastic) { any_hit_.resize((size_t)N_); stoch_idx_.resize((size_t)N_); u_.resize((size_t)N_); }
        const double lr0 = par_.learning_rate_, lr_max = lr0 * 1000.0, lr_min = lr0 / 100.0;
        const int full = RLITG_STAGE_SMOOTH | RLITG_STAGE_POSE_GRAD | RLITG_STAGE_EVAL | (scene_ ? RLITG_STAGE_COLL : 0);
        for (int it = 0; it < maxiter; it++) {
            if (lr_schedule_mode == 2) {
                const int period = maxiter / 2 > 0 ? maxiter / 2 : 1;
                const double ph = (double)(it % period) / period;
                const float v = (float)(lr_min + 0.5 * (lr_max - lr_min) * (1.0 + std::cos(ph * 3.14159265358979323846)));
                lr_iter_.upload(&v, 1);
            }
            for (int stage = 0; stage < 2; stage++) {
                RlitgTormSchedule sch;
                std::memset(&sch, 0, sizeof(sch));
                sch.lr_schedule_mode = lr_schedule_mode; sch.stage = stage;
                sch.use_collision_check = check_col; sch.use_velocity_check = par_.use_velocity_check_;
                sch.learning_rate = lr0; sch.learning_rate_min = lr_min; sch.learning_rate_max = lr_max;
                sch.hits_per_trajectory = 1;
                if (check_col || stochastic) {
                    if (stochastic) {
                        for (auto& v : u_host) v = uni(rng);
                        u_.upload(u_host);
                    }
                    check(rlitg_torm_collision_point(traj_.data(), N_, T_, RLITG_F32, scene_->handle(), cp_.base_off,
                                                     stochastic ? u_.data() : nullptr, nullptr, any_hit_.data(),
                                                     stochastic ? stoch_idx_.data() : nullptr, nullptr));
                }
                const int32_t* st = stochastic ? stoch_idx_.data() : nullptr;
                const int row_only = stochastic ? RLITG_STAGE_COLL_ROW_ONLY : 0;
                if (lr_schedule_mode == 1) {  // every optimiser picks its stage; the cost kernels skip the other one's work
                    check(rlitg_torm_stage_decide(N_, RLITG_F32, lr_.data(), infea_col_.data(), infea_vel_.data(), &sch,
                                                  stage_idx_.data(), nullptr));
                    run(full | row_only, st, stage_idx_.data());
                } else if (stage == 0) {
                    run(RLITG_STAGE_SMOOTH | RLITG_STAGE_EVAL | (scene_ ? (RLITG_STAGE_COLL | row_only) : 0), st);
                } else {
                    run(RLITG_STAGE_POSE_GRAD | RLITG_STAGE_EVAL);
                }
                check(rlitg_torm_bookkeep(traj_.data(), best_.data(), N_, T_, RLITG_F32, costs_.data(),
                                          check_col ? any_hit_.data() : nullptr, nullptr, alpha_.data(), lr_.data(),
                                          infea_col_.data(), infea_vel_.data(), best_cost_.data(), nullptr,
                                          lr_schedule_mode == 2 ? lr_iter_.data() : nullptr, &sch, nullptr));
                update(stage, lr_schedule_mode, lr_schedule_mode == 1 ? stage_idx_.data() : nullptr);
            }
        }
        const std::vector<float> bc = best_cost_.download();
        return std::vector<double>(bc.begin(), bc.end());
    }

    std::vector<float> costs() const { return costs_.download(); }
    const DeviceArray<float>& smoothnessIncrements() const { return smoothness_increments_; }
    const DeviceArray<float>& collisionIncrements() const { return collision_increments_; }
    const DeviceArray<float>& endPoseIncrements() const { return endPose_increments_; }
    const DeviceArray<float>& trajectories() const { return traj_; }

  private:
    static std::vector<double> widen(const std::vector<float>& v) { return std::vector<double>(v.begin(), v.end()); }
    std::vector<double> column(int k) const {
        const auto c = costs();
        std::vector<double> out(N_);
        for (int64_t n = 0; n < N_; n++) out[n] = c[n * RLITG_TORM_NCOST + k];
        return out;
    }
    void run(int stage_mask, const int32_t* stochastic_idx_dev = nullptr, const int32_t* traj_stage_dev = nullptr) {
        if ((stage_mask & RLITG_STAGE_COLL) && !scene_) throw Error(RLITG_ERR_NULL, "collision stage needs a scene");
        check(rlitg_torm_cost_grad_staged(traj_.data(), targets_.data(), N_, RLITG_F32, scene_ ? scene_->handle() : nullptr,
                                   joint_costs_.handle(), &cp_, stage_mask, traj_stage_dev, stochastic_idx_dev,
                                   smoothness_increments_.data(), collision_increments_.data(),
                                   endPose_increments_.data(), costs_.data(), nullptr, nullptr));
    }
    void update(int stage, int lr_schedule_mode, const int32_t* stage_idx_dev = nullptr) {
        RlitgTormUpdate u;
        std::memset(&u, 0, sizeof(u));
        u.stage = stage;
        u.lr_schedule_mode = lr_schedule_mode;
        u.smoothness_update_weight = par_.smoothness_update_weight_ / T_;
        u.obstacle_update_weight = par_.obstacle_update_weight_;
        u.joint_update_limit = lr_schedule_mode == 0 ? 0.1 : par_.joint_update_limit_;  // ctor override :104-106
        for (int k = 0; k < 7; k++) { u.joint_min[k] = par_.joint_min[k]; u.joint_max[k] = par_.joint_max[k]; }
        if (stage_idx_dev)
            check(rlitg_torm_update_staged(traj_.data(), N_, RLITG_F32, joint_costs_.handle(), &cp_, &u, stage_idx_dev,
                                           smoothness_increments_.data(), collision_increments_.data(),
                                           endPose_increments_.data(), alpha_.data(), lr_.data(), nullptr));
        else
            check(rlitg_torm_update(traj_.data(), N_, RLITG_F32, joint_costs_.handle(), &cp_, &u,
                                    smoothness_increments_.data(), collision_increments_.data(),
                                    endPose_increments_.data(), alpha_.data(), lr_.data(), nullptr));
    }

    int64_t N_;
    int T_;
    TormParameters par_;
    const Scene* scene_;
    TormCost joint_costs_;
    RlitgTormParams cp_;
    DeviceArray<float> targets_, traj_, best_, smoothness_increments_, collision_increments_, endPose_increments_, costs_,
        alpha_, lr_;
    DeviceArray<float> best_cost_, lr_iter_;
    DeviceArray<uint8_t> hits_, infea_col_, infea_vel_, any_hit_;
    DeviceArray<int32_t> stage_idx_, stoch_idx_;
    DeviceArray<float> u_;
};

// ---- PIRL path-wise IK environment --------------------------------------------------------------------------
// Constants of PirlEnv_Fetch.__init__ (pirl/script/envs/PirlEnv.py:118-174) and pirl/launch/pirl.yaml:9-14.
inline RlitgPirlCfg make_pirl_cfg(const std::string& rewardType = "Task+NullSpaceImitation",
                                  const std::string& obsType = "Conf+LinkPoses", int t = 1, int zDim = 0,
                                  bool train_mode = true, bool refineFlag = false) {
    RlitgPirlCfg c;
    std::memset(&c, 0, sizeof(c));
    const double ll[7] = {-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592};
    const double ul[7] = {1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592};
    const int cj[7] = {0, 0, 1, 0, 1, 0, 1};
    for (int j = 0; j < 7; j++) { c.ll[j] = ll[j]; c.ul[j] = ul[j]; c.cjoint[j] = cj[j]; }
    c.n_future = t; c.z_dim = zDim;
    if (obsType == "onlyConf") c.obs_type = 0;
    else if (obsType == "Conf+LinkPoses") c.obs_type = 1;
    else if (obsType == "Conf+EE") c.obs_type = 2;
    else if (obsType == "onlyLinkPoses") c.obs_type = 3;
    else throw Error(RLITG_ERR_ARG, "unknown obsType " + obsType);
    c.train_mode = train_mode; c.refine_flag = refineFlag;
    c.task_suc_r = 10.0; c.dist_suc_r = 1.0; c.rot_suc_r = 1.0; c.imit_suc_r = 0.0;
    if (rewardType == "Task+NullSpaceImitation") { c.task_reward_flag = 1; c.null_reward_flag = 1; }
    else if (rewardType == "Task+Imitation") { c.task_reward_flag = 1; c.imitate_reward_flag = 1; }
    else if (rewardType == "Task") { c.task_reward_flag = 1; }
    else if (rewardType == "Imitation") {
        c.imitate_reward_flag = 1; c.task_suc_r = c.dist_suc_r = c.rot_suc_r = 0.0; c.imit_suc_r = 10.0;
    } else throw Error(RLITG_ERR_ARG, "unknown rewardType " + rewardType);
    c.thr_early_stop = 0.2; c.thr_dist = 0.04; c.thr_rot = 0.104; c.thr_imitate = 0.087; c.zero_task_dist = 0.04 * 2;
    const double dp[4] = {2.0, -65.6, 30.0, 1.0}, rp[4] = {2.0, -5.0, 0.0, 0.0}, ip[4] = {1.0, -3.8, 0.05, 1.0},
                 np_[4] = {1.0, -15, 0.5, 1.0};
    for (int k = 0; k < 4; k++) { c.dist_p[k] = dp[k]; c.rot_p[k] = rp[k]; c.imit_p[k] = ip[k]; c.null_p[k] = np_[k]; }
    c.w_col = 10.0; c.w_jlimit = 1.0; c.w_early_stop = 3.0;
    return c;
}

// B environments stepped by one kernel: PirlEnv_Fetch.reset / step (PirlEnv.py:209-289, 352-392) with a leading
// batch axis.  State lives on the device as float32; observations / rewards / flags come back to the host each step
// (the policy of the C++ caller, PirlInterpolator's TorchScript module, runs on host tensors).
class PirlVecEnv {
  public:
    struct Step {
        std::vector<float> obs;      // [B,O]
        std::vector<float> reward;   // [B]
        std::vector<uint8_t> done;   // [B]
        std::vector<uint8_t> flags;  // [B,8] is_col, is_jlimit, is_early_stop, is_timeout, is_dist_suc, is_rot_suc,
                                     //       is_imitate_suc, is_task_suc
    };
    PirlVecEnv(int64_t B, const std::string& rewardType = "Task+NullSpaceImitation",
               const std::string& obsType = "Conf+LinkPoses", int t = 1, int zDim = 0, bool refineFlag = false,
               const Scene* scene = nullptr, std::array<double, 3> base_off = {-0.086875, 0.0, 0.37743})
        : B_(B), rewardType_(rewardType), obsType_(obsType), t_(t), z_(zDim), refine_(refineFlag), scene_(scene),
          base_off_(base_off), cfg_(make_pirl_cfg(rewardType, obsType, t, zDim, true, refineFlag)) {
        O_ = rlitg_pirl_obs_dim(&cfg_);
        q_.resize((size_t)B * 7); ts_.resize(B); mts_.resize(B); cnt_.resize((size_t)B * 4);
        obs_.resize((size_t)B * O_); rew_.resize(B); done_.resize(B); flags_.resize((size_t)B * 8);
        act_.resize((size_t)B * 7);
        if (zDim > 0) { latent_.resize((size_t)B * zDim); latent_.zero(); }
    }
    int64_t numEnvs() const { return B_; }
    int observationDim() const { return O_; }
    // target_poses [B,Tmax,7] x y z qx qy qz qw; demo_configs [B,Tmax,7] (may be empty); max_timestep [B]
    void setProblems(const std::vector<double>& target_poses, const std::vector<double>& demo_configs,
                     const std::vector<int32_t>& max_timestep, int Tmax) {
        if (target_poses.size() != (size_t)B_ * Tmax * 7 || max_timestep.size() != (size_t)B_ ||
            (!demo_configs.empty() && demo_configs.size() != target_poses.size()))
            throw Error(RLITG_ERR_SHAPE, "setProblems: sizes must be B*Tmax*7 / B");
        Tmax_ = Tmax;
        targets_.upload(std::vector<float>(target_poses.begin(), target_poses.end()));
        demos_host_ = demo_configs;
        if (!demo_configs.empty()) demos_.upload(std::vector<float>(demo_configs.begin(), demo_configs.end()));
        mts_.upload(max_timestep);
    }
    // reset (:352-392): cur_conf = demo_configs[0] with continuous joints refined to (-pi, pi] (:577-590) unless
    // start_conf [B,7] is given; counters and timestep cleared; returns the observation [B,O]
    std::vector<float> reset(const std::vector<double>* start_conf = nullptr, bool train_mode = true) {
        cfg_ = make_pirl_cfg(rewardType_, obsType_, t_, z_, train_mode, refine_);
        std::vector<float> q0((size_t)B_ * 7);
        const double pi = 3.14159265358979323846;
        for (int64_t b = 0; b < B_; b++)
            for (int j = 0; j < 7; j++) {
                double v;
                if (start_conf) v = (*start_conf)[b * 7 + j];
                else {
                    if (demos_host_.empty()) throw Error(RLITG_ERR_NULL, "reset needs start_conf or demo_configs");
                    v = demos_host_[((size_t)b * Tmax_) * 7 + j];
                    if (cfg_.cjoint[j] && std::fabs(v) > pi) {
                        double buf = v - (double)(long long)(v / (2 * pi)) * 2 * pi;
                        if (v > pi && buf > pi) buf -= 2 * pi;
                        if (v < -pi && buf < -pi) buf += 2 * pi;
                        v = buf;
                    }
                }
                q0[b * 7 + j] = (float)v;
            }
        q_.upload(q0);
        ts_.zero(); cnt_.zero();
        check(rlitg_pirl_obs(&cfg_, B_, RLITG_F32, q_.data(), ts_.data(), mts_.data(), targets_.data(), Tmax_,
                             z_ > 0 ? latent_.data() : nullptr, obs_.data(), nullptr));
        return obs_.download();
    }
    Step step(const std::vector<float>& action /*[B,7]*/) {
        if (action.size() != (size_t)B_ * 7) throw Error(RLITG_ERR_SHAPE, "action must hold B*7 values");
        act_.upload(action);
        check(rlitg_pirl_env_step(&cfg_, B_, RLITG_F32, q_.data(), act_.data(), ts_.data(), mts_.data(), targets_.data(),
                                  demos_host_.empty() ? nullptr : demos_.data(), Tmax_, z_ > 0 ? latent_.data() : nullptr,
                                  scene_ ? scene_->handle() : nullptr, base_off_.data(), obs_.data(), rew_.data(),
                                  done_.data(), flags_.data(), cnt_.data(), nullptr, nullptr));
        Step r;
        r.obs = obs_.download(); r.reward = rew_.download(); r.done = done_.download(); r.flags = flags_.download();
        return r;
    }
    std::vector<float> currentConfigurations() const { return q_.download(); }
    std::vector<int32_t> successCounters() const { return cnt_.download(); }  // [B,4] task, dist, rot, imitate

  private:
    int64_t B_;
    std::string rewardType_, obsType_;
    int t_, z_;
    bool refine_;
    const Scene* scene_;
    std::array<double, 3> base_off_;
    RlitgPirlCfg cfg_;
    int O_ = 0, Tmax_ = 0;
    std::vector<double> demos_host_;
    DeviceArray<float> q_, targets_, demos_, latent_, obs_, rew_, act_;
    DeviceArray<int32_t> ts_, mts_, cnt_;
    DeviceArray<uint8_t> done_, flags_;
};

// PirlInterpolator::interpolate (torm/src/pirl_interpolator.cpp:102-140) for B problems at once: observation ->
// policy -> setAction (test mode: add, wrap, clamp, no rejection), one waypoint per step.
// policy(obs [B*O], B, O) -> action [B*7].  Returns the generated configurations [B, n_steps, 7].
template <class Policy>
std::vector<float> pirl_interpolate(PirlVecEnv& env, Policy&& policy, int n_steps,
                                    const std::vector<double>* start_conf = nullptr) {
    const int64_t B = env.numEnvs();
    std::vector<float> out((size_t)B * n_steps * 7);
    std::vector<float> obs = env.reset(start_conf, /*train_mode=*/false);
    for (int k = 0; k < n_steps; k++) {
        const std::vector<float> act = policy(obs, B, env.observationDim());
        PirlVecEnv::Step r = env.step(act);
        obs = std::move(r.obs);
        const std::vector<float> q = env.currentConfigurations();
        for (int64_t b = 0; b < B; b++) std::memcpy(&out[((size_t)b * n_steps + k) * 7], &q[b * 7], 7 * sizeof(float));
    }
    return out;
}

}  // namespace rlitg
#endif  // RLITG_HPP_
