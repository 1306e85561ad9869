// Exercises include/rlitg.hpp (the C++ host mirror of torm::TormIKSolver / torm::TormOptimizer) on a GPU and checks
// it against the CPU oracle (oracle/rlitg_oracle.h -- test infrastructure; linked by tests only).
// Built and run by tests/test_cpp_host.py:  g++ -std=c++17 host_mirror_test.cpp -lrlitg -loracle
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "rlitg.hpp"
extern "C" {
#include "rlitg_oracle.h"
}

static int g_fail = 0;
#define EXPECT(cond, ...)                                              \
    do {                                                               \
        if (!(cond)) {                                                 \
            std::printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond); \
            std::printf(__VA_ARGS__);                                  \
            std::printf("\n");                                         \
            g_fail++;                                                  \
        }                                                              \
    } while (0)

static double max_abs_diff(const std::vector<double>& a, const double* b) {
    double m = 0;
    for (size_t i = 0; i < a.size(); i++) m = std::fmax(m, std::fabs(a[i] - b[i]));
    return m;
}

static const double kLL[7] = {-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592};
static const double kUL[7] = {1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592};

int main() {
    using namespace rlitg;
    int ndev = 0;
    check(rlitg_device_count(&ndev));
    std::printf("devices: %d, library version %d\n", ndev, rlitg_version());
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> G(0.0, 1.0);

    // ---- TormIKSolver, one configuration: pirl/launch/cfg/target_path/hello.yaml:2-3 known answer
    TormIKSolver ik;
    JntArray q(7);
    const double hello[7] = {0.32859072646481297, 0.180483432860646, 1.2922894312536055, 1.4023895537350413,
                             -0.07618293988860358, -1.767936957879053, -1.2634014392835944};
    for (int j = 0; j < 7; j++) q(j) = hello[j];
    Frame f;
    ik.fkSolver(q, f);
    EXPECT(std::fabs(f.p.x() - 0.8) < 1e-9 && std::fabs(f.p.y() - 0.45) < 1e-9 && std::fabs(f.p.z() - 0.25) < 1e-9,
           "p = %.12f %.12f %.12f", f.p.x(), f.p.y(), f.p.z());
    EXPECT(std::fabs(f.M(0, 0) - 1) < 1e-9 && std::fabs(f.M(1, 1) - 1) < 1e-9 && std::fabs(f.M(2, 2) - 1) < 1e-9, "M");

    std::vector<Frame> frames;
    ik.allfkSolver(q, frames);
    EXPECT(frames.size() == 8, "segments %zu", frames.size());
    double e = 0;
    for (int k = 0; k < 3; k++) e = std::fmax(e, std::fabs(frames[7].p[k] - f.p[k]));
    for (int k = 0; k < 9; k++) e = std::fmax(e, std::fabs(frames[7].M.data[k] - f.M.data[k]));
    EXPECT(e < 1e-12, "allfk tip vs fk: %g", e);

    Jacobian jac;
    ik.getJacobian(q, jac);
    std::vector<double> Jo(42);
    orc_jacobian(q.data.data(), 1, 7, Jo.data());
    EXPECT(max_abs_diff(jac.data, Jo.data()) < 1e-12, "jacobian %g", max_abs_diff(jac.data, Jo.data()));
    EXPECT(jac.columns() == 7 && std::fabs(jac(5, 0) - 1.0) < 1e-12, "J(5,0) = %g (shoulder_pan is a z joint)", jac(5, 0));

    Twist tw;
    tw.vel.data[0] = 0.01; tw.vel.data[1] = -0.02; tw.vel.data[2] = 0.005;
    tw.rot.data[0] = 0.03; tw.rot.data[1] = 0.01; tw.rot.data[2] = -0.02;
    JntArray dq(7);
    ik.vikSolver(q, tw, dq);
    const double twv[6] = {0.01, -0.02, 0.005, 0.03, 0.01, -0.02};
    std::vector<double> dqo(7);
    orc_pinv_step(q.data.data(), twv, 1, 7, 1e-5, dqo.data());
    EXPECT(max_abs_diff(dq.data, dqo.data()) < 1e-9, "vik %g", max_abs_diff(dq.data, dqo.data()));

    // ---- batched overloads vs oracle
    const int B = 1000;
    std::vector<double> qs((size_t)B * 7);
    for (int b = 0; b < B; b++)
        for (int j = 0; j < 7; j++) qs[b * 7 + j] = kLL[j] + (kUL[j] - kLL[j]) * U(rng);
    std::vector<double> pos, rot, J;
    ik.fkSolver(qs, pos, rot);
    ik.getJacobian(qs, J);
    std::vector<double> po(3 * B), ro(9 * B), jo(42 * B);
    orc_fk_jac(qs.data(), B, 7, po.data(), ro.data(), jo.data());
    EXPECT(max_abs_diff(pos, po.data()) < 1e-12 && max_abs_diff(rot, ro.data()) < 1e-12 && max_abs_diff(J, jo.data()) < 1e-12,
           "batched fk/jac %g %g %g", max_abs_diff(pos, po.data()), max_abs_diff(rot, ro.data()), max_abs_diff(J, jo.data()));

    // ---- TormCost known answer (torm/src/torm_optimizer.cpp:235 `// 96.77624`)
    {
        TormCost tc(320, 1.0);
        const double s = tc.getMaxQuadCostInvValue();
        EXPECT(std::fabs(s - 96.77624) < 1e-4, "scale %.6f", s);
    }

    // ---- scene: a table-like box in front of the robot
    const int nx = 64, ny = 96, nz = 64;
    const double res = 0.02;
    const std::array<double, 3> origin = {-0.2, -0.96, 0.0};
    std::vector<uint8_t> occ((size_t)nx * ny * nz, 0);
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                const double X = origin[0] + res * x, Y = origin[1] + res * y, Z = origin[2] + res * z;
                if (std::fabs(X - 0.8) <= 0.3 && std::fabs(Y) <= 0.7 && Z <= 0.62) occ[((size_t)z * ny + y) * nx + x] = 1;
            }
    // a small sphere table (link, centre, radius): forearm / wrist / gripper
    std::vector<int32_t> link = {3, 4, 5, 6, 7, 7};
    std::vector<double> center = {0.05, 0, 0, 0.08, 0, 0, 0.06, 0, 0, 0.05, 0, 0, -0.05, 0, 0, 0.03, 0, 0};
    std::vector<double> radius = {0.0625, 0.0625, 0.0625, 0.0625, 0.0625, 0.046875};
    Scene scene;
    scene.setSpheres(link, center, radius);
    scene.setOccupancy(occ, nx, ny, nz, origin, res, 1.0);
    std::vector<float> field((size_t)nx * ny * nz * 4);
    orc_build_field(occ.data(), nx, ny, nz, res, 1.0, field.data());
    OrcScene osc;
    osc.S = (int)link.size(); osc.link = link.data(); osc.center = center.data(); osc.radius = radius.data();
    osc.nx = nx; osc.ny = ny; osc.nz = nz;
    for (int k = 0; k < 3; k++) osc.origin[k] = origin[k];
    osc.res = res; osc.field = field.data();
    const double base_off[3] = {-0.086875, 0.0, 0.37743};
    {
        TormIKSolver ikc(&scene);
        auto hit = ikc.collisionChecking(qs, 7);
        std::vector<uint8_t> ho(B);
        orc_sphere_collide(qs.data(), B, &osc, base_off, ho.data());
        int nd = 0, nh = 0;
        for (int b = 0; b < B; b++) { nd += hit[b] != ho[b]; nh += ho[b]; }
        EXPECT(nd == 0 && nh > 0 && nh < B, "collision flags differ in %d of %d (%d hits)", nd, B, nh);
    }

    // ---- TormBatchOptimizer vs oracle on N noisy copies of a smooth joint path
    const int N = 16, T = 40;
    std::vector<double> path((size_t)T * 7), targets((size_t)T * 7);
    const double q0[7] = {0.2, 0.4, 0.3, 1.2, 0.1, 0.9, 0.2}, q1[7] = {-0.6, 0.1, -0.4, 1.6, 0.5, 0.5, -0.3};
    for (int i = 0; i < T; i++)
        for (int j = 0; j < 7; j++) {
            const double s = (double)i / (T - 1);
            path[i * 7 + j] = q0[j] + (q1[j] - q0[j]) * (3 * s * s - 2 * s * s * s);
        }
    orc_fk_quat(path.data(), T, 7, targets.data());
    for (auto& v : targets) v = std::round(v * 1e6) / 1e6;  // 6-digit poses like the reference's path files
    std::vector<double> traj((size_t)N * T * 7);
    for (int n = 0; n < N; n++)
        for (int i = 0; i < T; i++)
            for (int j = 0; j < 7; j++) traj[((size_t)n * T + i) * 7 + j] = path[i * 7 + j] + (i ? 0.03 * G(rng) : 0.0);
    {   // the device holds float32 trajectories: give the oracle the same values
        for (auto& v : traj) v = (double)(float)v;
        for (auto& v : targets) v = (double)(float)v;
    }
    TormParameters par;
    par.min_clearence_ = 0.3;
    TormBatchOptimizer opt(N, T, targets, par, &scene);
    opt.setTrajectories(traj);

    OrcTormParams op;
    RlitgTormParams cp = par.to_c(T);
    op.T = T; op.D = 7; op.dt = cp.dt; op.rot_scale = cp.rot_scale; op.jac_weight = cp.jac_weight;
    op.clearance = cp.clearance; op.pinv_ridge = cp.pinv_ridge; op.vik_eps = cp.vik_eps; op.use_pinv = cp.use_pinv;
    op.affine_axis_quirk = cp.affine_axis_quirk;
    for (int k = 0; k < 3; k++) op.base_off[k] = cp.base_off[k];
    for (int k = 0; k < 8; k++) { op.cjoint[k] = cp.cjoint[k]; op.vel_limit[k] = cp.vel_limit[k]; }
    std::vector<double> band((size_t)(T + 10) * 13);
    double scale = 0;
    const double w3[3] = {1.0, 0.0, 0.0};
    orc_smooth_setup(T, par.time_duration_, w3, 0.0, band.data(), &scale, nullptr);
    std::vector<double> oc((size_t)N * ORC_NCOST);
    const size_t ninc = (size_t)N * (T - 1) * 7;
    std::vector<double> is(ninc), icl(ninc), ip(ninc);
    orc_torm_cost_grad(traj.data(), targets.data(), &osc, &op, band.data(),
                       ORC_STAGE_POSE_GRAD | ORC_STAGE_SMOOTH | ORC_STAGE_COLL | ORC_STAGE_EVAL, nullptr, N, is.data(),
                       icl.data(), ip.data(), oc.data(), nullptr);

    auto pose_cost = opt.getEndPoseCost(true);
    double worst = 0;
    for (int n = 0; n < N; n++) {
        const double ref = oc[n * ORC_NCOST + 8] + 0.17 * oc[n * ORC_NCOST + 9];
        worst = std::fmax(worst, std::fabs(pose_cost[n] - ref) / ref);
    }
    EXPECT(worst < 1e-4, "getEndPoseCost rel err %g", worst);
    auto coll = opt.getCollisionCost();
    worst = 0;
    double csum = 0;
    for (int n = 0; n < N; n++) {
        worst = std::fmax(worst, std::fabs(coll[n] - oc[n * ORC_NCOST + 5]) / std::fmax(1e-6, oc[n * ORC_NCOST + 5]));
        csum += oc[n * ORC_NCOST + 5];
    }
    EXPECT(worst < 2e-3 && csum > 0, "getCollisionCost rel err %g (sum %g)", worst, csum);
    opt.calculateSmoothnessIncrements();
    {
        auto h = opt.smoothnessIncrements().download();
        double m = 0, mag = 0;
        for (size_t i = 0; i < ninc; i++) { m = std::fmax(m, std::fabs(h[i] - is[i])); mag = std::fmax(mag, std::fabs(is[i])); }
        EXPECT(m < 2e-5 * std::fmax(1.0, mag), "smoothness increments %g (max %g)", m, mag);
    }
    auto ev = opt.evaluate();
    worst = 0;
    for (int n = 0; n < N; n++)
        for (int k : {0, 1, 2, 6}) {
            const double ref = oc[n * ORC_NCOST + k];
            worst = std::fmax(worst, std::fabs(ev[n * RLITG_TORM_NCOST + k] - ref) / std::fmax(1e-9, std::fabs(ref)));
        }
    EXPECT(worst < 1e-4, "evaluate rel err %g", worst);

    // ---- the optimiser loop improves the reported objective
    double before = 0, after = 0;
    for (int n = 0; n < N; n++) before += ev[n * RLITG_TORM_NCOST + 6];
    TormBatchOptimizer free_opt(N, T, targets, par, nullptr);  // free space: every trajectory stays feasible
    free_opt.setTrajectories(traj);
    auto best = free_opt.localOptimizeTSGD(30);
    int nfinite = 0;
    for (int n = 0; n < N; n++)
        if (std::isfinite(best[n])) { after += best[n]; nfinite++; }
    EXPECT(nfinite == N, "%d of %d trajectories feasible", nfinite, N);
    EXPECT(after < 0.5 * before, "objective %g -> %g", before, after);
    auto bt = free_opt.getBestTrajectories();
    e = 0;
    for (int n = 0; n < N; n++)
        for (int j = 0; j < 7; j++) e = std::fmax(e, std::fabs(bt[(size_t)n * T * 7 + j] - traj[(size_t)n * T * 7 + j]));
    EXPECT(e == 0, "start configuration moved by %g", e);
    std::printf("objective %.5f -> %.5f over %d trajectories; kernels launched: %lld\n", before / N, after / N, N,
                (long long)rlitg_launch_count());

    // every lr_schedule_mode_ runs on the device (the scene-bound optimiser: collisions feed the adaptive schedule)
    for (int mode = 0; mode <= 3; mode++) {
        opt.setTrajectories(traj);
        auto bc = opt.localOptimizeTSGD(12, mode);
        int nf = 0;
        for (double v : bc) nf += std::isfinite(v);
        auto after_t = opt.getTrajectories();
        double moved = 0;
        for (size_t i = 0; i < after_t.size(); i++) moved = std::fmax(moved, std::fabs(after_t[i] - traj[i]));
        EXPECT(moved > 1e-3 && moved < 7.0, "mode %d: trajectories moved by %g", mode, moved);
        std::printf("lr_schedule_mode %d: %d of %d trajectories recorded a feasible best\n", mode, nf, N);
    }

    // ---- PirlVecEnv: reset + three steps against the oracle (FP32 device state vs FP64 oracle)
    {
        const int Bp = 96, Tp = 12;
        std::vector<double> demos((size_t)Bp * Tp * 7), tposes((size_t)Bp * Tp * 7);
        for (int b = 0; b < Bp; b++) {
            double cur[7];
            for (int j = 0; j < 7; j++) cur[j] = kLL[j] * 0.5 + (kUL[j] - kLL[j]) * 0.5 * U(rng);
            for (int i = 0; i < Tp; i++)
                for (int j = 0; j < 7; j++) {
                    cur[j] = std::fmin(kUL[j] - 0.05, std::fmax(kLL[j] + 0.05, cur[j] + 0.04 * G(rng)));
                    demos[((size_t)b * Tp + i) * 7 + j] = (double)(float)cur[j];
                }
        }
        orc_fk_quat(demos.data(), (int64_t)Bp * Tp, 7, tposes.data());
        for (auto& v : tposes) v = (double)(float)(std::round(v * 1e6) / 1e6);
        std::vector<int32_t> mts(Bp, Tp - 1);
        PirlVecEnv env(Bp, "Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0, false, &scene);
        env.setProblems(tposes, demos, mts, Tp);
        auto obs0 = env.reset();
        EXPECT(env.observationDim() == 88 && obs0.size() == (size_t)Bp * 88, "obs dim %d", env.observationDim());

        RlitgPirlCfg cc = make_pirl_cfg();
        OrcPirlCfg oc2;
        for (int j = 0; j < 7; j++) { oc2.ll[j] = cc.ll[j]; oc2.ul[j] = cc.ul[j]; oc2.cjoint[j] = cc.cjoint[j]; }
        oc2.n_future = cc.n_future; oc2.z_dim = cc.z_dim; oc2.obs_type = cc.obs_type; oc2.train_mode = cc.train_mode;
        oc2.refine_flag = cc.refine_flag; oc2.task_reward_flag = cc.task_reward_flag;
        oc2.imitate_reward_flag = cc.imitate_reward_flag; oc2.null_reward_flag = cc.null_reward_flag;
        oc2.task_suc_r = cc.task_suc_r; oc2.dist_suc_r = cc.dist_suc_r; oc2.rot_suc_r = cc.rot_suc_r; oc2.imit_suc_r = cc.imit_suc_r;
        oc2.thr_early_stop = cc.thr_early_stop; oc2.thr_dist = cc.thr_dist; oc2.thr_rot = cc.thr_rot;
        oc2.thr_imitate = cc.thr_imitate; oc2.zero_task_dist = cc.zero_task_dist;
        for (int k = 0; k < 4; k++) { oc2.dist_p[k] = cc.dist_p[k]; oc2.rot_p[k] = cc.rot_p[k]; oc2.imit_p[k] = cc.imit_p[k]; oc2.null_p[k] = cc.null_p[k]; }
        oc2.w_col = cc.w_col; oc2.w_jlimit = cc.w_jlimit; oc2.w_early_stop = cc.w_early_stop;

        std::vector<double> oq((size_t)Bp * 7), oobs((size_t)Bp * 88), orew(Bp), oerr((size_t)Bp * 4);
        for (int b = 0; b < Bp; b++)
            for (int j = 0; j < 7; j++) oq[b * 7 + j] = demos[((size_t)b * Tp) * 7 + j];
        std::vector<int32_t> ots(Bp, 0), ocnt((size_t)Bp * 4, 0);
        std::vector<uint8_t> odone(Bp), oflags((size_t)Bp * 8);
        orc_pirl_obs(&oc2, Bp, oq.data(), ots.data(), mts.data(), tposes.data(), Tp, nullptr, oobs.data());
        double m = 0;
        for (size_t i = 0; i < obs0.size(); i++) m = std::fmax(m, std::fabs(obs0[i] - oobs[i]));
        EXPECT(m < 2e-5, "reset observation %g", m);
        int flag_diff = 0, ncol = 0;
        double mobs = 0, mrew = 0;
        for (int k = 0; k < 3; k++) {
            std::vector<float> act((size_t)Bp * 7);
            for (int b = 0; b < Bp; b++)
                for (int j = 0; j < 7; j++)
                    act[b * 7 + j] = (float)(demos[((size_t)b * Tp + k + 1) * 7 + j] - oq[b * 7 + j] + 0.01 * G(rng));
            auto r = env.step(act);
            orc_pirl_env_step(&oc2, Bp, oq.data(), act.data(), ots.data(), mts.data(), tposes.data(), demos.data(), Tp,
                              nullptr, &osc, base_off, oobs.data(), orew.data(), odone.data(), oflags.data(), ocnt.data(),
                              oerr.data());
            const auto qd = env.currentConfigurations();
            for (int b = 0; b < Bp; b++) {
                bool same = r.done[b] == odone[b];
                for (int f = 0; f < 8; f++) same = same && r.flags[b * 8 + f] == oflags[b * 8 + f];
                ncol += oflags[b * 8];
                if (!same) { flag_diff++; continue; }
                mrew = std::fmax(mrew, std::fabs(r.reward[b] - orew[b]));
                for (int o = 0; o < 88; o++) mobs = std::fmax(mobs, std::fabs(r.obs[(size_t)b * 88 + o] - oobs[(size_t)b * 88 + o]));
                for (int j = 0; j < 7; j++) oq[b * 7 + j] = qd[b * 7 + j];  // keep both sides on the same FP32 state
            }
        }
        EXPECT(flag_diff <= 2, "flags differ in %d env-steps", flag_diff);
        // reward carries 2 exp(-5 r) with r = 2 acos|w|: FP32 resolves r to ~5e-4 near 0 -> 1e-2
        EXPECT(mobs < 5e-5 && mrew < 1e-2, "obs %g reward %g", mobs, mrew);
        std::printf("pirl: 3 steps x %d envs, %d collisions, %d threshold ties\n", Bp, ncol, flag_diff);

        // pirl_interpolate with a "follow the demonstration" policy reproduces the demo path
        PirlVecEnv env2(Bp, "Task", "Conf+LinkPoses", 1, 0, false, nullptr);
        env2.setProblems(tposes, demos, mts, Tp);
        int step_k = 0;
        auto policy = [&](const std::vector<float>& obs, int64_t nb, int od) {
            std::vector<float> a((size_t)nb * 7);
            for (int64_t b = 0; b < nb; b++)
                for (int j = 0; j < 7; j++)  // obs[:7] is the current configuration (Conf+LinkPoses)
                    a[b * 7 + j] = (float)demos[((size_t)b * Tp + step_k + 1) * 7 + j] - obs[(size_t)b * od + j];
            step_k++;
            return a;
        };
        auto gen = pirl_interpolate(env2, policy, Tp - 1);
        m = 0;
        for (int b = 0; b < Bp; b++)
            for (int k = 0; k < Tp - 1; k++)
                for (int j = 0; j < 7; j++)
                    m = std::fmax(m, std::fabs(gen[((size_t)b * (Tp - 1) + k) * 7 + j] - demos[((size_t)b * Tp + k + 1) * 7 + j]));
        EXPECT(m < 1e-5, "pirl_interpolate vs demonstration %g", m);
    }

    // ---- error behaviour: C codes surface as exceptions
    bool threw = false;
    try {
        std::vector<double> bad(5, 0.0), p2, r2;
        ik.fkSolver(bad, p2, r2, 5);
    } catch (const Error& err) {
        threw = err.code == RLITG_ERR_SHAPE;
    }
    EXPECT(threw, "D = 5 must throw RLITG_ERR_SHAPE");

    std::printf(g_fail ? "FAILED (%d)\n" : "OK\n", g_fail);
    return g_fail ? 1 : 0;
}
