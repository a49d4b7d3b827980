"""GPU parity for SURVEY 8f-2: dsdtm_pose_optimize[_batch] (Optimizer::PoseOptimization, ref: src/Optimizer.cpp:20-101) against
the oracle's restatement of ceres::Solve for the reference's configuration.

Tolerances (floating point, stated here): final pose within 1e-9 (unit quaternion components and metres; the north_star budget
is 1e-5 rad / 1e-5 m), residual norms within 1e-9, costs within 1e-9 relative, and the DECISIONS -- number of iterations, accepted
steps, stopping rule -- equal. Differences are rounding only: per-lane partial sums + xor-butterfly instead of sequential sums, FMA
contraction, J'J scaled after accumulation, Cholesky substitution instead of the explicit inverse, CUDA's sin/cos/atan/log."""
import numpy as np
import pytest

import helpers as H
import oracle as O

pytestmark = pytest.mark.gpu


def _orc(pr, max_iters=100):
    return O.pose_optimization(pr["normals"], pr["levels"], pr["points_w"], pr["pose_in"], max_iters)


def _problem(case):
    seed, n, noise, outl, lvl, srot, strans = case
    return H.make_ba_problem(seed, n=n, noise=noise, outliers=outl, max_level=lvl, start_rot=srot, start_trans=strans)


def _same(sa, sb):
    return (sb["iterations"], sb["termination"], sb["n_successful"]) == (sa["iterations"], sa["termination"], sa["n_successful"])


@pytest.mark.parametrize("case", H.BA_CASES)
def test_pose_optimize_matches_oracle(ctx, case):
    from dsdtm_b200 import capi
    pr = _problem(case)
    a, ra, sa = _orc(pr)
    b, rb, sb = ctx.pose_optimize(H.ba_obs_records(pr, capi.BA_OBS_DT), pr["pose_in"])
    assert _same(sa, sb), (sa, sb)
    assert sb["n_obs"] == len(pr["levels"])
    assert np.abs(a - b).max() < 1e-9
    assert np.abs(ra - rb).max() < 1e-9
    assert abs(sa["final_cost"] - sb["final_cost"]) <= 1e-9 * max(1.0, sa["final_cost"])
    assert abs(sa["initial_cost"] - sb["initial_cost"]) <= 1e-12 * max(1.0, sa["initial_cost"])


def test_pose_optimize_batch_ragged_equals_single_and_is_deterministic(ctx):
    from dsdtm_b200 import capi
    prs = [_problem(c) for c in H.BA_CASES if c[1] <= 300] + [H.make_ba_problem(50 + i, n=40 + 13 * i, max_level=i % 4) for i in range(9)]
    stride = 304
    obs = np.stack([H.ba_obs_records(p, capi.BA_OBS_DT, stride) for p in prs])
    n_obs = np.array([len(p["levels"]) for p in prs], np.int32)
    poses = np.stack([p["pose_in"] for p in prs])
    out, res, sm = ctx.pose_optimize_batch(obs, n_obs, poses)
    out2, res2, sm2 = ctx.pose_optimize_batch(obs, n_obs, poses)
    valid0 = np.arange(stride)[None, :] < n_obs[:, None]
    assert (out == out2).all() and (res[valid0] == res2[valid0]).all() and (sm == sm2).all()       # run-to-run bit-equal
    for i, p in enumerate(prs):
        b, rb, sb = ctx.pose_optimize(obs[i, :n_obs[i]], p["pose_in"])
        assert (b == out[i]).all() and (rb == res[i, :n_obs[i]]).all() and sb == sm[i]   # a frame's result does not depend on its neighbours
        a, ra, sa = _orc(p)
        assert _same(sa, sm[i]), i
        assert np.abs(a - out[i]).max() < 1e-9 and np.abs(ra - res[i, :n_obs[i]]).max() < 1e-9
    for nf in (2, 3):
        o, r, s = ctx.pose_optimize_batch(obs[:nf], n_obs[:nf], poses[:nf])
        assert (o == out[:nf]).all() and (s == sm[:nf]).all()
    # the sweep kernel (one warp per frame; chosen automatically above one frame per SM) on the same frames: another
    # reduction tree, so equal to rounding, with the same decisions; partial last CTA (21 frames, 4 per CTA)
    ctx.set_option("pose_opt_solo_max", 0)
    try:
        ow, rw, sw = ctx.pose_optimize_batch(obs, n_obs, poses)
        ow2, rw2, sw2 = ctx.pose_optimize_batch(obs, n_obs, poses)
    finally:
        ctx.set_option("pose_opt_solo_max", -1)
    valid = np.arange(stride)[None, :] < n_obs[:, None]                          # columns past n_obs are never written
    assert (ow == ow2).all() and (rw[valid] == rw2[valid]).all() and (sw == sw2).all()
    assert np.abs(ow - out).max() < 1e-10 and np.abs(rw[valid] - res[valid]).max() < 1e-10
    for k in ("iterations", "termination", "n_successful", "n_obs"):
        assert (sw[k] == sm[k]).all(), k
    # without residual norms / with an iteration cap
    o, r, s = ctx.pose_optimize_batch(obs, n_obs, poses, max_iters=2, want_res=False)
    assert r is None and (s["iterations"] <= 2).all()
    for i in (0, 3):
        a, _, sa = _orc(prs[i], 2)
        assert _same(sa, s[i]) and np.abs(a - o[i]).max() < 1e-10


def test_pose_optimize_edge_cases_and_bad_arguments(ctx):
    from dsdtm_b200 import capi
    pr = H.make_ba_problem(21, n=50)
    # no residual block: pose unchanged (exp(log R) round trip), the stopping rule says so
    b, rb, sb = ctx.pose_optimize(np.zeros(0, capi.BA_OBS_DT), pr["pose_in"])
    assert sb["termination"] == capi.BA_NO_RESIDUALS and np.abs(b - pr["pose_in"]).max() < 1e-15 and len(rb) == 0
    # zero iterations
    b, rb, sb = ctx.pose_optimize(H.ba_obs_records(pr, capi.BA_OBS_DT), pr["pose_in"], max_iters=0)
    assert sb["termination"] == capi.BA_NO_CONVERGENCE and sb["iterations"] == 0 and np.abs(b - pr["pose_in"]).max() < 1e-15
    a, ra, _ = _orc(pr, 0)
    assert np.abs(ra - rb).max() < 1e-12
    # the largest frame the call accepts, and one more
    big = H.make_ba_problem(22, n=capi.BA_MAX_OBS, max_level=2, outliers=0.2)
    a, ra, sa = _orc(big)
    b, rb, sb = ctx.pose_optimize(H.ba_obs_records(big, capi.BA_OBS_DT), big["pose_in"])
    assert _same(sa, sb) and np.abs(a - b).max() < 1e-9 and np.abs(ra - rb).max() < 1e-9
    with pytest.raises(capi.DsdtmError):
        ctx.pose_optimize(np.zeros(capi.BA_MAX_OBS + 1, capi.BA_OBS_DT), pr["pose_in"])
    bad = H.ba_obs_records(pr, capi.BA_OBS_DT); bad["level"][3] = 31
    with pytest.raises(capi.DsdtmError):
        ctx.pose_optimize(bad, pr["pose_in"])
    bad["level"][3] = -1
    with pytest.raises(capi.DsdtmError):
        ctx.pose_optimize(bad, pr["pose_in"])
    # a residual that cannot be evaluated: Ceres gives up at iteration zero, parameters untouched
    nan = H.ba_obs_records(pr, capi.BA_OBS_DT); nan["normal"][5, 2] = 0.0
    bn, rn, sn = ctx.pose_optimize(nan, pr["pose_in"])
    an, _, san = O.pose_optimization(nan["normal"], nan["level"], nan["point_w"], pr["pose_in"])
    assert sn["termination"] == capi.BA_FAILURE == san["termination"] and sn["iterations"] == 0 and np.abs(bn - pr["pose_in"]).max() < 1e-15
    assert not np.isfinite(rn[5]) and np.isfinite(np.delete(rn, 5)).all()
    # the context stays usable
    b2, _, sb2 = ctx.pose_optimize(H.ba_obs_records(pr, capi.BA_OBS_DT), pr["pose_in"])
    a2, _, sa2 = _orc(pr)
    assert _same(sa2, sb2) and np.abs(a2 - b2).max() < 1e-9


def test_pose_optimize_full_size_sweep_properties(ctx):
    """BASELINE-size sweep (4096 independent frames x 300 observations): size-independent properties instead of the oracle on
    every frame -- each solve lowers the Cauchy cost, exact data comes back to the truth, replicas of one frame are bit-equal,
    and a sample of frames agrees with the oracle."""
    from dsdtm_b200 import capi
    nf, n = 4096, 300
    base = [H.make_ba_problem(300 + i, n=n, max_level=i % 4, noise=0.0 if i % 4 == 0 else 1e-3, outliers=0.0 if i % 4 == 0 else 0.1) for i in range(16)]
    obs = np.stack([H.ba_obs_records(base[i % 16], capi.BA_OBS_DT) for i in range(nf)])
    n_obs = np.full(nf, n, np.int32)
    poses = np.stack([base[i % 16]["pose_in"] for i in range(nf)])
    out, res, sm = ctx.pose_optimize_batch(obs, n_obs, poses)
    assert (sm["final_cost"] <= sm["initial_cost"]).all() and (sm["termination"] != capi.BA_FAILURE).all()
    for i in range(16):
        assert (out[i::16] == out[i]).all() and (res[i::16] == res[i]).all()          # replicas bit-equal wherever they ran
        a, ra, sa = _orc(base[i])
        assert _same(sa, sm[i]) and np.abs(a - out[i]).max() < 1e-9
        if i % 4 == 0:
            assert np.abs(out[i] - base[i]["truth"]).max() < 1e-6 and res[i].max() < 1e-6
