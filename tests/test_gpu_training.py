"""Whole-model gradient parity: PanguModel in train mode on the B200 --
forward on the tape, weighted-L1 loss kernel, hand-written backward -- against fp32 autograd through
the CPU oracle (the reference's loss.backward(), models/pangu_sample.py:52-69) with the same weights,
inputs, targets and DropPath draws, full depth on a 96-column longitude strip.

The weighted-L1 loss is not smooth: dL/d(output) = +-w/N flips sign wherever the forward's 16-bit
rounding error exceeds |output - target|, i.e. at ~0.4 % of the grid points, which alone moves the
gradient SEED by ~12 % in L2 -- an artefact of comparing two slightly different forwards, not a backward
error.  The backward is therefore checked with the oracle's own dL/d(output) fed to both sides; the
loss kernel (value and seed) is checked against the oracle separately (tests/test_gpu_rollout.py).

Stated tolerance (per-parameter relative L2 over all 223 gradients, bf16 operands): 16-bit operands
enter every dgrad / wgrad GEMM and the attention backward, so the floor is the operand rounding
accumulated over up to 16 blocks: <= 5e-2 (measured: 2.1e-2 worst; forward tolerance is 2e-2).
"""
import pytest
import torch

from oracle import pangu_oracle as O
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_GRAD = {"bf16": 5e-2, "fp16": 1e-2}
# earth_specific_bias tables (91 % of the parameters): dS = P o (dP - rowsum(P o dP)) cancels a large common-mode part, so
# the rounding of q / k / v / dO is amplified.  fp16 operands -- the DEFAULT of the training path -- are within 4e-2
# (measured <= 1.6e-2); the optional bf16 training format is 7-15 % off and only bounded here, see DESIGN.md 9.
TOL_BIAS_TABLE = {"bf16": 2.5e-1, "fp16": 4e-2}


def _setup(fmt, seed=3):
    import pangu_pytorch_b200 as pb
    pb.set_operand_dtype(fmt)
    pb.set_training_operand_dtype(fmt)          # the training path has its own default (fp16); pin it to the format under test
    pb.free_workspaces()
    p = O.stress_weights(seed=seed, bias_std=0.5)
    model = pb.PanguModel(device=DEV)
    model.load_state_dict(p, strict=True)
    model = model.to(DEV).train()
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=1, lat=721, lon=96)
    g = torch.Generator().manual_seed(7)
    tu = torch.randn(1, 5, 13, 721, 96, generator=g)
    ts = torch.randn(1, 4, 721, 96, generator=g)
    return pb, p, model, (up, sf, stats, maps, ch), (tu, ts)


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_backward_gradients_match_oracle_autograd(fmt):
    pb, p, model, (up, sf, stats, maps, ch), (tu, ts) = _setup(fmt)
    from pangu_pytorch_b200 import training
    torch.manual_seed(5)            # DropPath draws
    dstats = [s.to(DEV) for s in stats]
    out, out_s = model(up.to(DEV), sf.to(DEV), dstats, maps.to(DEV), ch.to(DEV))
    tape = model._tape[1]
    scales = [(t.s1, t.s2) for t in tape.blocks]
    ref_loss, ref, (gu, gs) = O.loss_and_grads(p, up, sf, stats, maps, ch, tu, ts, drop_scales=scales)
    torch.autograd.backward((out, out_s), (gu.to(DEV), gs.to(DEV)))
    torch.cuda.synchronize()
    worst = []
    for name, prm in model.named_parameters():
        assert prm.grad is not None, name
        assert torch.isfinite(prm.grad).all(), name
        worst.append((rel_l2(prm.grad, ref[name]), name))
    worst.sort(reverse=True)
    cats = {}
    for e, n in worst:
        k = n.split("EarthSpecificBlock")[-1].split(".", 1)[-1] if "EarthSpecificBlock" in n else n
        cats[k] = max(cats.get(k, 0.0), e)
    print(f"[{fmt}] worst gradient rel-L2 per parameter kind:", {k: round(v, 4) for k, v in sorted(cats.items(), key=lambda kv: -kv[1])})
    for e, n in worst:
        assert e < (TOL_BIAS_TABLE[fmt] if n.endswith("earth_specific_bias") else TOL_GRAD[fmt]), (n, e)
    training.release_tape(model)


def test_train_step_loss_and_seed():
    """train_step = forward + loss kernel + backward: loss value against the oracle; the seed differs from the
    oracle's only where the forward rounding flips sign(output - target)."""
    pb, p, model, (up, sf, stats, maps, ch), (tu, ts) = _setup("bf16")
    from pangu_pytorch_b200 import training
    torch.manual_seed(5)
    dstats = [s.to(DEV) for s in stats]
    loss = training.train_step(model, up.to(DEV), sf.to(DEV), dstats, maps.to(DEV), ch.to(DEV), tu.to(DEV), ts.to(DEV))
    scales = [(t.s1, t.s2) for t in model._tape[1].blocks]
    ref_loss, ref, _ = O.loss_and_grads(p, up, sf, stats, maps, ch, tu, ts, drop_scales=scales)
    assert abs(float(loss) - float(ref_loss)) < 5e-3 * abs(float(ref_loss))
    # aggregated gradients (weights) still agree closely; only weakly-aggregated ones feel the flipped seeds
    w = dict(model.named_parameters())["layers.EarthSpecificLayer0.blocks.EarthSpecificBlock0.linear.linear1.weight"]
    assert rel_l2(w.grad, ref["layers.EarthSpecificLayer0.blocks.EarthSpecificBlock0.linear.linear1.weight"]) < 5e-2
    training.release_tape(model)


def test_gradients_accumulate_and_frozen_parameters_are_skipped():
    pb, p, model, (up, sf, stats, maps, ch), (tu, ts) = _setup("bf16")
    from pangu_pytorch_b200 import training
    for name, prm in model.named_parameters():
        if "earth_specific_bias" in name or "norm" in name:
            prm.requires_grad_(False)
    for blk in [m for m in model.modules() if hasattr(m, "drop_path")]:
        blk.drop_path.drop_prob = 0.0                           # deterministic: the two steps must agree
    dstats = [s.to(DEV) for s in stats]
    args = (up.to(DEV), sf.to(DEV), dstats, maps.to(DEV), ch.to(DEV), tu.to(DEV), ts.to(DEV))
    training.train_step(model, *args)
    g1 = {n: q.grad.clone() for n, q in model.named_parameters() if q.requires_grad}
    training.train_step(model, *args)
    for n, q in model.named_parameters():
        if not q.requires_grad:
            assert q.grad is None, n
        else:
            assert rel_l2(q.grad, 2 * g1[n]) < 1e-3, n           # .grad accumulates (fp32 atomics: order varies)
    training.release_tape(model)


def test_lora_training_gradients():
    """lora_tune: frozen base, rank-16 adapters on the 67 nn.Linear modules, the two output
    convs trained in full.  Gradients of the adapters and of modules_to_save against oracle autograd through
    ``W + (alpha/r) B A`` (peft semantics, lora_dropout = 0)."""
    pb, p, model, (up, sf, stats, maps, ch), (tu, ts) = _setup("bf16")
    from pangu_pytorch_b200 import lora, training
    lora.add_lora(model, r=16, lora_alpha=16.0, lora_dropout=0.0)
    model.to(DEV).train()
    loras = {n: m for n, m in model.named_modules() if isinstance(m, lora.LoraLinear)}
    assert len(loras) == 67
    g = torch.Generator().manual_seed(11)
    for m in loras.values():                      # peft starts B at zero; give the adapters something to do
        m.B.data.copy_(0.02 * torch.randn(m.B.shape, generator=g))
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    assert len(trainable) == 2 * 67 + 4
    for blk in [m for m in model.modules() if hasattr(m, "drop_path")]:
        blk.drop_path.drop_prob = 0.0
    dstats = [s.to(DEV) for s in stats]
    out, out_s = model(up.to(DEV), sf.to(DEV), dstats, maps.to(DEV), ch.to(DEV))
    # oracle: autograd through the merged weights
    leaves, eff = {}, dict(p)
    for n, m in loras.items():
        a = m.A.detach().cpu().clone().requires_grad_(True)
        b = m.B.detach().cpu().clone().requires_grad_(True)
        leaves[n + ".lora_A.default.weight"], leaves[n + ".lora_B.default.weight"] = a, b
        eff[n + ".weight"] = p[n + ".weight"] + m.scaling * (b @ a)
    for n in ("_output_layer.conv.weight", "_output_layer.conv.bias", "_output_layer.conv_surface.weight",
              "_output_layer.conv_surface.bias"):
        leaves[n] = eff[n] = p[n].clone().requires_grad_(True)
    ou, os_ = O.forward_train(eff, up, sf, stats, maps, ch)
    ou.retain_grad(); os_.retain_grad()
    tun, tsn = O.norm_data(tu, ts, O.output_statistics(stats))
    O.weighted_l1_loss(ou, os_, tun, tsn).backward()
    torch.autograd.backward((out, out_s), (ou.grad.to(DEV), os_.grad.to(DEV)))
    torch.cuda.synchronize()
    named = dict(model.named_parameters())
    worst = sorted(((rel_l2(named[n].grad, leaves[n].grad), n) for n in trainable), reverse=True)
    print("[lora] worst gradient rel-L2:", worst[:4])
    assert worst[0][0] < 5e-2, worst[:4]
    assert all(q.grad is None for n, q in named.items() if n not in trainable)
    # checkpoint round trip: peft-style keys -> merged plain weights == the effective weights the kernels used
    merged = lora.merge_lora_state_dict(lora.peft_state_dict(model))
    n0 = "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock2.attention.linear1"
    assert torch.allclose(merged[n0 + ".weight"].cpu(), eff[n0 + ".weight"].detach(), atol=1e-6)
    training.release_tape(model)


def test_full_025_backward_directional_derivative():
    """Size-independent check of the backward at the FULL 0.25 degree shapes (BASELINE.json configs[3]), where the
    CPU oracle's autograd is out of reach: for random directions v in a few parameter tensors, the central
    difference of the loss kernel's value, (L(theta + eps v) - L(theta - eps v)) / (2 eps), must match <grad, v>.
    The loss is a mean over 1e8 grid points, so the forward's rounding noise averages out of the difference."""
    import pangu_pytorch_b200 as pb
    from pangu_pytorch_b200 import ops, training, engine
    pb.set_operand_dtype("bf16")
    pb.set_training_operand_dtype("bf16")
    pb.free_workspaces()
    torch.manual_seed(0)
    model = pb.PanguModel(device=DEV).to(DEV).train()
    for blk in [m for m in model.modules() if hasattr(m, "drop_path")]:
        blk.drop_path.drop_prob = 0.0
    g = torch.Generator(device=DEV).manual_seed(1)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    up, sf, maps, ch = rn(1, 5, 13, 721, 1440), rn(1, 4, 721, 1440), rn(1, 3, 724, 1440), rn(1, 1, 1, 13, 721, 1440)
    tu, ts = rn(1, 5, 13, 721, 1440), rn(1, 4, 721, 1440)
    stats = [torch.zeros(4, device=DEV), torch.ones(4, device=DEV), torch.zeros(13, 1, 1, 5, device=DEV),
             torch.ones(13, 1, 1, 5, device=DEV)]
    training.train_step(model, up, sf, stats, maps, ch, tu, ts)
    params = dict(model.named_parameters())
    assert all(torch.isfinite(p.grad).all() for p in params.values())

    def loss_at():
        with torch.no_grad():
            ou, os_ = model(up, sf, stats, maps, ch)
            l, _, _ = ops.l1_loss(ou, os_, tu, ts, stats[0], stats[1], stats[2].reshape(13, 5).contiguous(),
                                  stats[3].reshape(13, 5).contiguous())
        return float(l)

    names = ["_output_layer.conv.bias", "layers.EarthSpecificLayer3.blocks.EarthSpecificBlock1.norm2.bias",
             "layers.EarthSpecificLayer2.blocks.EarthSpecificBlock0.linear.linear2.bias",
             "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1.attention.linear1.bias"]
    for i, name in enumerate(names):
        p = params[name]
        v = torch.randn(p.shape, device=DEV, generator=g)
        v /= v.norm()
        analytic = float((p.grad * v).sum())
        eps = 2e-2
        with torch.no_grad():
            p.add_(eps * v); lp = loss_at()
            p.add_(-2 * eps * v); lm = loss_at()
            p.add_(eps * v)
        numeric = (lp - lm) / (2 * eps)
        print(f"directional derivative {name}: analytic {analytic:.4e} numeric {numeric:.4e}")
        assert abs(numeric - analytic) < 0.1 * abs(analytic) + 2e-5, (name, analytic, numeric)
    training.release_tape(model)
    pb.free_workspaces()


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
def test_full_025_gradients_against_reference_golden(fmt):
    """(fp16 = the default operand format of the training path, engine.training_operand_dtype.)  All 223 gradients at the FULL 0.25 degree shapes against the UNMODIFIED reference's own autograd
    (tests/golden/train_grads.npz, written by oracle/make_golden.py --what train: eval-mode DropPath, stress weights,
    0.5 * weighted-MSE loss so that the seed is smooth in the outputs).  The seed is formed here from the GPU
    forward's own outputs, so this is the whole chain -- forward on the tape, seed, backward -- against the reference."""
    import pangu_pytorch_b200 as pb
    from pangu_pytorch_b200 import training
    from tests.util import golden
    gold = golden("train_grads.npz")
    pb.set_operand_dtype(fmt)
    pb.set_training_operand_dtype(fmt)
    pb.free_workspaces()
    p = O.stress_weights(seed=int(gold["weights_seed"]), bias_std=0.5)
    model = pb.PanguModel(device=DEV)
    model.load_state_dict(p, strict=True)
    model = model.to(DEV).train()
    for blk in [m for m in model.modules() if hasattr(m, "drop_path")]:
        blk.drop_path.drop_prob = 0.0
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=int(gold["inputs_seed"]), nontrivial_stats=True)
    g = torch.Generator().manual_seed(int(gold["targets_seed"]))
    tu = torch.randn(1, 5, 13, 721, 1440, generator=g).to(DEV)
    ts = torch.randn(1, 4, 721, 1440, generator=g).to(DEV)
    out, out_s = model(up.to(DEV), sf.to(DEV), [s.to(DEV) for s in stats], maps.to(DEV), ch.to(DEV))
    wu = torch.tensor(O.UPPER_WEIGHTS, device=DEV).view(1, 5, 1, 1, 1)
    ws = torch.tensor(O.SURFACE_WEIGHTS, device=DEV).view(1, 4, 1, 1)
    with torch.no_grad():
        du, ds = out - tu, out_s - ts
        loss = 0.5 * (du * du * wu).mean() + 0.125 * (ds * ds * ws).mean()
        gu, gs = du * wu / du.numel(), 0.25 * ds * ws / ds.numel()
    assert abs(float(loss) - float(gold["loss"])) < 2e-3 * float(gold["loss"])
    torch.autograd.backward((out, out_s), (gu, gs))
    torch.cuda.synchronize()
    worst = []
    for name, prm in model.named_parameters():
        pos = torch.from_numpy(gold[f"grad.{name}.pos"])
        ref = torch.from_numpy(gold[f"grad.{name}.val"]).double()
        mine = prm.grad.reshape(-1)[pos.to(DEV)].double().cpu()
        err = float((mine - ref).norm() / ref.norm().clamp_min(1e-30))
        nrm = float(prm.grad.double().norm()) / float(gold[f"grad.{name}.l2"])
        worst.append((err, name, nrm))
    worst.sort(reverse=True)
    print("[full-size grads vs reference] worst:", [(round(e, 4), n.split("EarthSpecific")[-1], round(r, 4)) for e, n, r in worst[:6]])
    for err, name, nrm in worst:
        tol = TOL_BIAS_TABLE[fmt] if name.endswith("earth_specific_bias") else TOL_GRAD[fmt]
        assert err < tol and abs(nrm - 1.0) < tol, (name, err, nrm)
    training.release_tape(model)
    pb.free_workspaces()
