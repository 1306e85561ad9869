#!/usr/bin/env python
"""CPU study (no GPU needed): where do the FP32 collision increments lose digits?

A NumPy emulation of kernel 3b's per-sphere arithmetic in which every stage can run in float32 or float64
(NumPy rounds after every operation, so float32 here is slightly pessimistic against the kernel's fused
multiply-adds).  The result is compared with the FP64 oracle exactly as tests/test_torm_gpu.py does (row-wise
error relative to the row's largest component).  Stages:
  fk      sphere centres, joint origins / axes from the chain walk
  stencil 7-point velocity / acceleration  (variant "rel": differences against the centre row first)
  grad    projector, curvature, Cartesian gradient
  solve   J J^T + ridge I and the 3 x 3 solve  (variants: adjugate inverse, Cholesky)
  jt      J^T y
Usage: python tools/coll_fp32_study.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

OX = [0.119525, 0.117, 0.219, 0.133, 0.197, 0.1245, 0.1385]
OZ = [0.34858, 0.0599999999999999, 0, 0, 0, 0, 0]
AX = [2, 1, 0, 1, 0, 1, 0]
TIP = 0.16645
RULE_V = np.array([-1.0 / 60, 3.0 / 20, -3.0 / 4, 0.0, 3.0 / 4, -3.0 / 20, 1.0 / 60])
RULE_A = np.array([1.0 / 90, -3.0 / 20, 3.0 / 2, -49.0 / 18, 3.0 / 2, -3.0 / 20, 1.0 / 90])


def chain(q, dt):
    """q [M,7] -> link frames: R [M,8,3,3], p [M,8,3] (segment tips 0..6 and gripper_link 7), in dtype dt"""
    q = q.astype(dt)
    M = q.shape[0]
    R = np.tile(np.eye(3, dtype=dt), (M, 1, 1))
    p = np.zeros((M, 3), dtype=dt)
    Rs, ps = [], []
    for j in range(7):
        p = p + R[:, :, 0] * dt(OX[j])
        if OZ[j] != 0:
            p = p + R[:, :, 2] * dt(OZ[j])
        s, c = np.sin(q[:, j]).astype(dt), np.cos(q[:, j]).astype(dt)
        a = AX[j]
        Rn = R.copy()
        if a == 2:
            Rn[:, :, 0] = c[:, None] * R[:, :, 0] + s[:, None] * R[:, :, 1]
            Rn[:, :, 1] = c[:, None] * R[:, :, 1] - s[:, None] * R[:, :, 0]
        elif a == 1:
            Rn[:, :, 0] = c[:, None] * R[:, :, 0] - s[:, None] * R[:, :, 2]
            Rn[:, :, 2] = s[:, None] * R[:, :, 0] + c[:, None] * R[:, :, 2]
        else:
            Rn[:, :, 1] = c[:, None] * R[:, :, 1] + s[:, None] * R[:, :, 2]
            Rn[:, :, 2] = c[:, None] * R[:, :, 2] - s[:, None] * R[:, :, 1]
        R = Rn
        Rs.append(R.copy()); ps.append(p.copy())
    p = p + R[:, :, 0] * dt(TIP)
    Rs.append(R.copy()); ps.append(p.copy())
    return np.stack(Rs, 1), np.stack(ps, 1)


def emulate(traj, field, origin, res, link, center, radius, dtp, clearance=0.3, ridge=1e-4, quirk=True,
            d_fk=np.float32, d_st=np.float32, d_gr=np.float32, d_so=np.float32, d_jt=np.float32, rel_stencil=False,
            solver="adj", d_fk_st=None, d_fk_jac=None, fixed_point=0):
    from rlitg_b200 import fetch
    N, T, _ = traj.shape
    S = len(link)
    R, p = chain(traj.reshape(-1, 7), d_fk)
    off = np.asarray(fetch.BASE_OFFSET).astype(d_fk)
    c = center.astype(d_fk)
    Rl, pl = R[:, link], p[:, link]                                   # [M,S,3,3], [M,S,3]
    x = (off + pl + np.einsum("msab,sb->msa", Rl, c)).astype(d_fk)    # [M,S,3]
    jo = (p[:, :7] + off).astype(d_fk)                                # [M,7,3]
    ax = np.stack([R[:, j, :, AX[j]] for j in range(7)], 1)           # [M,7,3]
    ja = (ax + jo if quirk else ax).astype(d_fk)
    x = x.reshape(N, T, S, 3); jo = jo.reshape(N, T, 7, 3); ja = ja.reshape(N, T, 7, 3)
    x_st, x_jac = x, x
    if d_fk_st is not None or d_fk_jac is not None:   # separate precisions for the stencil positions / the Jacobian inputs
        def fk_as(dt_):
            R_, p_ = chain(traj.reshape(-1, 7), dt_)
            o_ = np.asarray(fetch.BASE_OFFSET).astype(dt_)
            x_ = (o_ + p_[:, link] + np.einsum("msab,sb->msa", R_[:, link], center.astype(dt_))).reshape(N, T, S, 3)
            jo_ = (p_[:, :7] + o_)
            ax_ = np.stack([R_[:, j, :, AX[j]] for j in range(7)], 1)
            ja_ = (ax_ + jo_ if quirk else ax_)
            return x_, jo_.reshape(N, T, 7, 3), ja_.reshape(N, T, 7, 3)
        x_st = fk_as(d_fk_st or d_fk)[0]
        x_jac, jo, ja = fk_as(d_fk_jac or d_fk)
    # nearest-cell lookup (float32 centres decide the cell, as in the kernel)
    u = np.floor((x.astype(np.float32) - np.float32(origin)) * np.float32(1.0 / res) + np.float32(0.5)).astype(int)
    nz, ny, nx = field.shape[:3]
    inb = (u[..., 0] >= 0) & (u[..., 1] >= 0) & (u[..., 2] >= 0) & (u[..., 0] < nx) & (u[..., 1] < ny) & (u[..., 2] < nz)
    uc = np.clip(u, 0, [nx - 1, ny - 1, nz - 1])
    rec = field[uc[..., 2], uc[..., 1], uc[..., 0]]                    # [N,T,S,4] float32
    ok = inb & (rec[..., 0] < 1e29)
    d = rec[..., 0].astype(d_gr) - radius.astype(d_gr)
    cl = d_gr(clearance)
    pot = np.where(d >= cl, 0, np.where(d >= 0, d_gr(0.5) * ((d - cl) * cl) * (d - cl), -d + d_gr(0.5) * cl))
    pot = np.where(ok, pot, 0).astype(d_gr)
    # stencils over clamped rows
    idx = np.clip(np.arange(T)[:, None] + np.arange(-3, 4)[None], 0, T - 1)   # [T,7]
    xs = x_st[:, idx].astype(d_st)                                          # [N,T,7,S,3]
    inv = d_st(1.0 / dtp)
    if fixed_point:
        # centres stored as 32-bit fixed point (2^-fixed_point metres per unit, from the FP64 walk): the differences against
        # the centre row are exact integer subtractions, rounded to FP32 once (relative to the DIFFERENCE, not the position)
        if os.environ.get("FIXED_RC32"):
            # what the kernel can afford: link origin from the FP64 walk, R c in FP32 (|R c| <= 0.2 m: ~1e-8 m)
            R64, p64 = chain(traj.reshape(-1, 7), np.float64)
            o64 = (np.asarray(fetch.BASE_OFFSET) + p64[:, link])                                   # [M,S,3]
            rc32 = np.einsum("msab,sb->msa", R64[:, link].astype(np.float32), center.astype(np.float32)).astype(np.float32)
            xi_all = (np.rint(o64 * 2.0 ** fixed_point) + np.rint(rc32.astype(np.float64) * 2.0 ** fixed_point)).astype(np.int64)
            xi = xi_all.reshape(N, T, S, 3)[:, idx]
        else:
            xi = np.rint(x_st[:, idx].astype(np.float64) * 2.0 ** fixed_point).astype(np.int64)
        xr = ((xi - xi[:, :, 3:4]).astype(np.float64) * 2.0 ** -fixed_point).astype(d_st)
        vel = sum((inv * d_st(RULE_V[k])) * xr[:, :, k] for k in range(7) if k != 3)
        acc = sum((inv * inv * d_st(RULE_A[k])) * xr[:, :, k] for k in range(7) if k != 3)
    elif rel_stencil:
        xr = xs - xs[:, :, 3:4]
        vel = sum((inv * d_st(RULE_V[k])) * xr[:, :, k] for k in range(7) if k != 3)
        acc = sum((inv * inv * d_st(RULE_A[k])) * xr[:, :, k] for k in range(7) if k != 3)
    else:
        vel = sum((inv * d_st(RULE_V[k])) * xs[:, :, k] for k in range(7) if k != 3)
        acc = sum((inv * inv * d_st(RULE_A[k])) * xs[:, :, k] for k in range(7))
    vel = vel.astype(d_gr); acc = acc.astype(d_gr)
    vm = np.sqrt((vel * vel).sum(-1))
    nv = vel / vm[..., None]
    pg = -rec[..., 1:].astype(d_gr)
    Pa = acc - (nv * acc).sum(-1, keepdims=True) * nv
    Pg = pg - (nv * pg).sum(-1, keepdims=True) * nv
    curv = Pa / (vm * vm)[..., None]
    cg = ((Pg - curv * pot[..., None]) * vm[..., None]).astype(d_so)
    # Jacobian columns
    nanc = np.minimum(link + 1, 7)
    mask = (np.arange(7)[None] < nanc[:, None])                         # [S,7]
    r = x_jac[:, :, :, None, :].astype(d_so) - jo[:, :, None, :, :].astype(d_so)   # [N,T,S,7,3]
    col = np.cross(np.broadcast_to(ja[:, :, None].astype(d_so), r.shape), r) * mask[None, None, :, :, None]
    Mm = np.einsum("ntsja,ntsjb->ntsab", col, col).astype(d_so) + d_so(ridge) * np.eye(3, dtype=d_so)
    if solver == "adj":
        m = Mm.reshape(*Mm.shape[:-2], 9)
        m0, m1, m2, m3, m4, m5, m6, m7, m8 = [m[..., k] for k in range(9)]
        det = m0 * (m4 * m8 - m5 * m7) - m1 * (m3 * m8 - m5 * m6) + m2 * (m3 * m7 - m4 * m6)
        idt = d_so(1) / det
        Mi = np.stack([(m4 * m8 - m5 * m7) * idt, (m2 * m7 - m1 * m8) * idt, (m1 * m5 - m2 * m4) * idt,
                       (m5 * m6 - m3 * m8) * idt, (m0 * m8 - m2 * m6) * idt, (m2 * m3 - m0 * m5) * idt,
                       (m3 * m7 - m4 * m6) * idt, (m1 * m6 - m0 * m7) * idt, (m0 * m4 - m1 * m3) * idt], -1)
        Mi = Mi.reshape(Mm.shape)
        y = np.einsum("ntsab,ntsb->ntsa", Mi, cg).astype(d_so)
    elif solver == "qr":
        # stable route: A = [J^T ; sqrt(ridge) I] = Q R (modified Gram-Schmidt on the 3 columns), J^T (A^T A)^-1 g = Q1 R^-T g
        sq = d_so(np.sqrt(ridge))
        A = np.concatenate([col, np.broadcast_to(sq * np.eye(3, dtype=d_so), col.shape[:3] + (3, 3))], axis=3)  # [N,T,S,10,3]
        a0, a1, a2 = A[..., 0].copy(), A[..., 1].copy(), A[..., 2].copy()
        r00 = np.sqrt((a0 * a0).sum(-1)); q0 = a0 / r00[..., None]
        r01 = (q0 * a1).sum(-1); a1 = a1 - r01[..., None] * q0
        r02 = (q0 * a2).sum(-1); a2 = a2 - r02[..., None] * q0
        r11 = np.sqrt((a1 * a1).sum(-1)); q1 = a1 / r11[..., None]
        r12 = (q1 * a2).sum(-1); a2 = a2 - r12[..., None] * q1
        r22 = np.sqrt((a2 * a2).sum(-1)); q2 = a2 / r22[..., None]
        w0 = cg[..., 0] / r00
        w1 = (cg[..., 1] - r01 * w0) / r11
        w2 = (cg[..., 2] - r02 * w0 - r12 * w1) / r22
        z = (q0[..., :7] * w0[..., None] + q1[..., :7] * w1[..., None] + q2[..., :7] * w2[..., None])
        act = pot >= 1e-4
        inc = -(z.astype(d_jt) * act[..., None]).sum(2)
        return inc[:, 1:].astype(np.float64), act, Mm
    else:  # Cholesky  M = L L^T
        a00, a10, a20, a11, a21, a22 = Mm[..., 0, 0], Mm[..., 1, 0], Mm[..., 2, 0], Mm[..., 1, 1], Mm[..., 2, 1], Mm[..., 2, 2]
        l00 = np.sqrt(a00); l10 = a10 / l00; l20 = a20 / l00
        l11 = np.sqrt(a11 - l10 * l10); l21 = (a21 - l20 * l10) / l11
        l22 = np.sqrt(a22 - l20 * l20 - l21 * l21)
        z0 = cg[..., 0] / l00; z1 = (cg[..., 1] - l10 * z0) / l11; z2 = (cg[..., 2] - l20 * z0 - l21 * z1) / l22
        y2 = z2 / l22; y1 = (z1 - l21 * y2) / l11; y0 = (z0 - l10 * y1 - l20 * y2) / l00
        y = np.stack([y0, y1, y2], -1).astype(d_so)
    contrib = np.einsum("ntsja,ntsa->ntsj", col.astype(d_jt), y.astype(d_jt))   # [N,T,S,7]
    act = pot >= 1e-4
    inc = -(contrib * act[..., None]).sum(2)
    return inc[:, 1:].astype(np.float64), act, Mm


def main():
    from oracle import oracle as O
    from rlitg_b200 import fetch, synth
    f32, f64 = np.float32, np.float64
    for seed, T, N, dtp in ((107, 100, 6, 0.1), (307, 300, 3, 0.1), (44, 37, 5, 0.2)):
        origin, res = (-0.3, -0.96, 0.0), 0.025
        traj = synth.linear_trajectories(N, T, seed=seed, noise=0.03)
        traj = traj.astype(f32).astype(f64)
        targets = np.round(synth.random_ee_paths(1, T, seed=seed + 1)[0], 6)
        occ, _ = synth.random_box_scene((64, 96, 64), origin, res, seed=seed, n_clutter=3)
        link, center, radius = fetch.default_spheres()
        fld = O.build_field(occ, res, 1.0)
        osc = O.Scene(fld, origin, res, link, center, radius)
        op = O.torm_params(T, dt=dtp, clearance=0.3)
        want = O.torm_cost_grad(traj, targets, osc, op, None, O.STAGE_COLL)["inc_coll"]
        field = np.asarray(fld).reshape(occ.shape[0], occ.shape[1], occ.shape[2], 4)

        def stats(name, **kw):
            got, act, Mm = emulate(traj, field, origin, res, link, center, radius, dtp, **kw)
            err = np.abs(got - want).max(-1)
            rr = (err / (np.abs(want).max(-1) + 1e-6)).ravel()
            print("  %-44s med %.1e  >1e-4 %.4f  >1e-3 %.4f  max %.1e" % (name, np.median(rr), (rr > 1e-4).mean(),
                                                                         (rr > 1e-3).mean(), rr.max()))
        print("seed %d T %d N %d dt %.1f" % (seed, T, N, dtp))
        allf64 = dict(d_fk=f64, d_st=f64, d_gr=f64, d_so=f64, d_jt=f64)
        stats("all FP64 (emulation sanity)", **allf64)
        stats("all FP32, adjugate (round-1 kernel)")
        stats("all FP32, adjugate, rel stencil", rel_stencil=True)
        stats("all FP32, cholesky", solver="chol")
        stats("all FP32, cholesky, rel stencil", solver="chol", rel_stencil=True)
        stats("all FP32, QR of [J^T; sqrt(ridge) I]", solver="qr")
        stats("all FP32, QR, rel stencil", solver="qr", rel_stencil=True)
        kw = dict(allf64); kw["d_so"] = f32
        stats("only d_so in FP32 (qr)", solver="qr", **kw)
        kw = dict(allf64); kw["d_fk"] = f32
        stats("only d_fk in FP32 (qr)", solver="qr", **kw)
        kw = dict(d_fk=f64, d_st=f32, d_gr=f32, d_so=f32, d_jt=f32)
        stats("FK FP64, rest FP32 qr+rel", solver="qr", rel_stencil=True, **kw)
        stats("same + fixed-point centres (2^-29 m)", solver="qr", fixed_point=29, **kw)
        os.environ["FIXED_RC32"] = "1"
        stats("same, R c in FP32 (what the kernel does)", solver="qr", fixed_point=29, **kw)
        os.environ.pop("FIXED_RC32")
        stats("FP64 but stencil positions from FP32 FK (qr)", solver="qr", d_fk_st=f32, **allf64)
        stats("FP64 but Jacobian inputs from FP32 FK (qr)", solver="qr", d_fk_jac=f32, **allf64)
        for st in ("d_fk", "d_st", "d_gr", "d_so", "d_jt"):
            kw = dict(allf64); kw[st] = f32
            stats("only %s in FP32 (adj)" % st, **kw)
        kw = dict(allf64); kw["d_so"] = f32
        stats("only d_so in FP32 (chol)", solver="chol", **kw)
        kw = dict(allf64); kw["d_st"] = f32
        stats("only d_st in FP32, rel stencil", rel_stencil=True, **kw)
        kw = dict(d_fk=f64, d_st=f32, d_gr=f32, d_so=f32, d_jt=f32)
        stats("FK FP64, rest FP32 chol+rel", solver="chol", rel_stencil=True, **kw)


if __name__ == "__main__":
    main()
