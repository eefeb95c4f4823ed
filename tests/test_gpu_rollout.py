"""BASELINE.json config 2/3 on the GPU: de-normalisation glue, chained forecast steps (checked per
step against the oracle, teacher-forced and free-running) and ensemble sharding on one device."""
import pytest
import torch

from oracle import pangu_oracle as O
from tests.util import TOL_MODEL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(p, fmt):
    import pangu_pytorch_b200 as pb
    pb.set_operand_dtype(fmt)
    pb.free_workspaces()
    m = pb.PanguModel(device=DEV)
    m.load_state_dict(p, strict=True)
    return m.to(DEV).eval()


def _err(got, ref):
    got, ref = got.detach().double().cpu(), ref.double()
    return float(((got - ref).flatten(2).norm(dim=2) / ref.flatten(2).norm(dim=2)).max())


def test_denorm_fields_matches_normBackData():
    from pangu_pytorch_b200.rollout import denormalize_
    g = torch.Generator().manual_seed(3)
    up, sf = torch.randn(1, 5, 13, 721, 96, generator=g), torch.randn(1, 4, 721, 96, generator=g)
    stats = (torch.randn(4, generator=g), 0.5 + torch.rand(4, generator=g),
             torch.randn(13, 1, 1, 5, generator=g), 0.5 + torch.rand(13, 1, 1, 5, generator=g))
    ru, rs = O.norm_back_data(up, sf, O.output_statistics(stats))
    gu, gs = denormalize_(up.to(DEV).clone(), sf.to(DEV).clone(), stats)
    assert torch.allclose(gu.cpu(), ru, rtol=1e-6, atol=1e-6) and torch.allclose(gs.cpu(), rs, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_three_step_rollout_on_strip(fmt):
    """x_{k+1} = normBack(model(x_k)) on a 96-column strip, 3 chained steps."""
    from pangu_pytorch_b200.rollout import rollout
    p = O.reference_like_weights(seed=0)
    m = _model(p, fmt)
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=1, lat=721, lon=96)
    ref = O.rollout(p, up, sf, stats, maps, ch, steps=3)
    d = lambda t: t.to(DEV)
    dstats = [d(s) for s in stats]
    free = rollout(m, d(up), d(sf), dstats, d(maps), d(ch), steps=3)
    # free running: error may grow, bounded by ~2x the single-step tolerance
    for k, ((gu, gs), (ru, rs)) in enumerate(zip(free, ref)):
        eu, es = _err(gu, ru), _err(gs, rs)
        print(f"rollout {fmt} free-running step {k + 1}: upper {eu:.3e} surface {es:.3e}")
        assert eu < 2 * TOL_MODEL[fmt] and es < 2 * TOL_MODEL[fmt]
    # teacher forced: both fed the oracle's x_k
    for k in range(1, 3):
        iu, is_ = ref[k - 1]
        gu, gs = rollout(m, d(iu), d(is_), dstats, d(maps), d(ch), steps=1)[0]
        eu, es = _err(gu, ref[k][0]), _err(gs, ref[k][1])
        print(f"rollout {fmt} teacher-forced step {k + 1}: upper {eu:.3e} surface {es:.3e}")
        assert eu < TOL_MODEL[fmt] and es < TOL_MODEL[fmt]


def test_ensemble_members_are_independent_and_sharded():
    from pangu_pytorch_b200 import ensemble
    p = O.reference_like_weights(seed=0)
    m = _model(p, "bf16")
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=1, lat=721, lon=96)
    d = lambda t: t.to(DEV)
    args = (m, d(up), d(sf), [d(s) for s in stats], d(maps), d(ch))
    summ = lambda ou, os_: (float(ou.double().mean()), float(os_.double().mean()))
    whole = ensemble.run_ensemble(*args, n_members=4, rank=0, world=1, reduce=summ)
    parts = {}
    for r in range(2):
        parts.update(ensemble.run_ensemble(*args, n_members=4, rank=r, world=2, reduce=summ))
    assert sorted(whole) == sorted(parts) == [0, 1, 2, 3]
    assert all(whole[k] == parts[k] for k in whole)           # bit-identical regardless of sharding
    assert len({whole[k] for k in whole}) == 4                # perturbations differ


def test_weighted_l1_loss_and_output_gradient():
    """models/pangu_sample.py:57-67 on the GPU vs the oracle (and autograd of the oracle for dL/d out)."""
    from pangu_pytorch_b200 import ops
    g = torch.Generator().manual_seed(8)
    lat, lon = 721, 96
    ou, os_ = torch.randn(1, 5, 13, lat, lon, generator=g), torch.randn(1, 4, lat, lon, generator=g)
    stats = (torch.randn(4, generator=g), 0.5 + torch.rand(4, generator=g),
             torch.randn(13, 1, 1, 5, generator=g), 0.5 + torch.rand(13, 1, 1, 5, generator=g))
    ostats = O.output_statistics(stats)
    tu_n, ts_n = torch.randn(1, 5, 13, lat, lon, generator=g), torch.randn(1, 4, lat, lon, generator=g)
    tu, ts = O.norm_back_data(tu_n, ts_n, ostats)                       # physical-unit targets
    ou_r, os_r = ou.clone().requires_grad_(True), os_.clone().requires_grad_(True)
    ref = O.weighted_l1_loss(ou_r, os_r, *O.norm_data(tu, ts, ostats))
    ref.backward()
    d = lambda t: t.to(DEV).contiguous()
    loss, gu, gs = ops.l1_loss(d(ou), d(os_), d(tu), d(ts), d(stats[0]), d(stats[1]), d(stats[2].reshape(13, 5)),
                               d(stats[3].reshape(13, 5)), want_grad=True)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 2e-6 * abs(float(ref))
    # gradient: sign(o - t) * w / N; compare where |o - t| is not at the fp32 noise floor
    for got, want, o, t in ((gu, ou_r.grad, ou, O.norm_data(tu, ts, ostats)[0]), (gs, os_r.grad, os_, O.norm_data(tu, ts, ostats)[1])):
        mask = (o - t).abs() > 1e-4
        assert torch.allclose(got.cpu()[mask], want[mask], rtol=1e-5, atol=0)


@pytest.mark.parametrize("normalised", [True, False])
def test_evaluation_scores_match_oracle(normalised):
    """pangu_scores (latitude-weighted RMSE / ACC per plane, models/pangu_sample.py:236-270) against the oracle's
    restatement of era5_data/score.py; with normalised=True the kernel applies normBackData on the fly."""
    from pangu_pytorch_b200 import ops
    _, _, stats, _, _ = O.synthetic_inputs(seed=1, lat=721, lon=96)
    g = torch.Generator().manual_seed(9)
    ou, os_ = torch.randn(1, 5, 13, 721, 96, generator=g), torch.randn(1, 4, 721, 96, generator=g)
    tu, ts = torch.randn(1, 5, 13, 721, 96, generator=g) * 2 + 1, torch.randn(1, 4, 721, 96, generator=g) * 2 + 1
    ref = O.evaluation_scores(ou.double(), os_.double(), tu.double(), ts.double(), [s.double() for s in stats])
    d = lambda t: t.to(DEV)
    s_mean, s_std = d(stats[0]).reshape(4), d(stats[1]).reshape(4)
    u_mean, u_std = d(stats[2]).reshape(13, 5).contiguous(), d(stats[3]).reshape(13, 5).contiguous()
    if normalised:
        pu, ps = d(ou), d(os_)
    else:
        pu, ps = (d(t.float()) for t in O.norm_back_data(ou, os_, O.output_statistics(stats)))
    got = ops.scores(pu.contiguous(), ps.contiguous(), d(tu), d(ts), s_mean, s_std, u_mean, u_std, normalised=normalised)
    for a, b in zip(got, ref):
        assert tuple(a.shape) == tuple(b.shape)
        assert torch.allclose(a.cpu().double(), b, rtol=2e-5, atol=1e-6)


def test_graphed_rollout_is_bit_identical_to_eager():
    """The CUDA-graph step replays exactly the eager kernels: 3 chained steps must agree bit for bit."""
    from pangu_pytorch_b200.rollout import rollout, rollout_graphed
    p = O.reference_like_weights(seed=0)
    m = _model(p, "bf16")
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=1, lat=721, lon=96)
    d = lambda t: t.to(DEV)
    args = (m, d(up), d(sf), [d(s) for s in stats], d(maps), d(ch))
    eager = rollout(*args, steps=3)
    graphed = rollout_graphed(*args, steps=3)
    for (eu, es), (gu, gs) in zip(eager, graphed):
        assert torch.equal(eu, gu) and torch.equal(es, gs)


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_full_025_seven_day_rollout_against_reference_golden(fmt):
    """BASELINE.json configs[1]: 7 x 24 h free-running rollout at the full 0.25 degree shapes, every step against the
    UNMODIFIED reference's own rollout (tests/golden/rollout7.npz, oracle/make_golden.py --what rollout).  One step
    attenuates an input error by ~0.5, so the free-running bound is 2x the single-step tolerance."""
    from pangu_pytorch_b200.rollout import rollout
    from tests.util import golden, sampled_rel_l2
    gold = golden("rollout7.npz")
    p = O.reference_like_weights(seed=int(gold["weights_seed"]))
    m = _model(p, fmt)
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=int(gold["inputs_seed"]), nontrivial_stats=True)
    d = lambda t: t.to(DEV)
    outs = rollout(m, d(up), d(sf), [d(s) for s in stats], d(maps), d(ch), steps=int(gold["steps"]), keep_on_device=False)
    for k, (gu, gs) in enumerate(outs):
        eu = sampled_rel_l2(gu, gold, f"step{k + 1}.upper")
        es = sampled_rel_l2(gs, gold, f"step{k + 1}.surface")
        vu = (gu[0].double().flatten(1).norm(dim=1) / torch.from_numpy(gold[f"step{k + 1}.upper.var_l2"])).numpy()
        print(f"rollout {fmt} day {k + 1}: sampled rel-L2 upper {eu:.3e} surface {es:.3e}; per-variable norm ratio {vu.min():.4f}..{vu.max():.4f}")
        # free running over seven chained steps: every step adds its own operand-rounding error to a state that already
        # differs, so the bound is 4x the single-step tolerance (measured: bf16 <= 2.6e-2 at days 5-7, fp16 <= 2.8e-3)
        assert eu < 4 * TOL_MODEL[fmt] and es < 4 * TOL_MODEL[fmt]
        assert abs(vu - 1).max() < 2 * TOL_MODEL[fmt]
