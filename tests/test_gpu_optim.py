"""GPU: optimiser bookkeeping around the fused clip + Adam kernel (SFM:32-37,158-162) and the per-sample weighted Dice of
the north-star EMA step: parameters trained through host-API autograd ops (cls_proj, MS:721), per-parameter Adam step
counts for parameters that become live late, torch.optim.Adam-compatible state dicts."""
import pytest
import torch

from oracle import segformer_oracle as O
from oracle import synth
from tests import helpers

pytestmark = pytest.mark.gpu


def _wrapper_for(cfg_kwargs, lr=1e-3, wd=5e-5):
    """SegFormerModel wrapper around an arbitrary (small) config with the deterministic synthetic weights."""
    from semisupervisedobjectdetection_b200.configuration import SegformerConfig
    from semisupervisedobjectdetection_b200.modeling import SegformerForSemanticSegmentation
    from semisupervisedobjectdetection_b200.optim import FusedClipAdam
    from semisupervisedobjectdetection_b200.wrapper import SegFormerModel
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w = SegFormerModel(None, lr=lr, weight_decay=wd, scheduler=0.97, device="cuda:0", num_labels=cfg_kwargs.get("num_labels", 1),
                           variant="nvidia/mit-b0")
    ocfg = O.OracleConfig()
    for k, v in cfg_kwargs.items():
        setattr(ocfg, k, v)
    net = SegformerForSemanticSegmentation(SegformerConfig(**cfg_kwargs))
    sd = synth.synth_state_dict(O.state_dict_template(ocfg))
    net.load_state_dict(sd, strict=True)
    net.config.output_hidden_states = True
    w.model = net.to("cuda:0")
    w.optimizer = FusedClipAdam([p for p in w.model.parameters()], lr=lr, weight_decay=wd, betas=(0.5, 0.999), clip_value=1.2,
                                model=w.model)
    return w, ocfg, sd


def test_cls_proj_is_trained_when_cls_tokens_are_in_use():
    """add_cls_token (SFM:92-101) routes sigmoid(cls) through linear_c[i].cls_proj (MS:720-723); the reference's Adam trains
    those Linears.  Their gradients are written by autograd (host-API Linear), not by libsfk: the optimiser must still see
    them as live.  Compared with the oracle's gradient + one clip/Adam step."""
    case = helpers.cases.MODEL_CASES["cls512"]
    w, ocfg, sd = _wrapper_for(dict(case["config"]))
    cls = [synth.uniform01(f"clsopt/cls{s}", (1, ocfg.hidden_sizes[s])) for s in range(4)]
    enc = w.model.segformer.encoder
    for s in range(4):
        enc.cls_token[s] = cls[s].clone().cuda().requires_grad_(True)
    x = synth.uniform01("clsopt/x", (2, 3, 64, 64))
    mask = (synth.uniform01("clsopt/mask", (2, 64, 64)) > 0.7).float()
    loss, _ = w.train_one_epoch(x, mask)

    sd_o = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in sd.items()}
    logits, _, _ = O.forward_logits(x, sd_o, ocfg, None, [c.clone() for c in cls])
    loss_o = O.segmentation_loss(O.predict_probs(logits, (64, 64)), mask.to(torch.int64), "dice")
    loss_o.backward()
    assert abs(float(loss) - float(loss_o)) < 2e-3 * abs(float(loss_o))
    named = dict(w.model.named_parameters())
    for i in range(4):
        for leaf in ("weight", "bias"):
            k = f"decode_head.linear_c.{i}.cls_proj.{leaf}"
            g = sd_o[k].grad
            assert g is not None and float(g.abs().max()) > 0
            want, _, _ = O.clip_adam_l2_step(sd[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, 1e-3, 5e-5)
            got = named[k].detach().cpu()
            assert not torch.equal(got, sd[k]), f"{k} was not updated"
            upd_got, upd_ref = got - sd[k], want - sd[k]
            big = upd_ref.abs() > 0.5e-3
            assert big.any()
            agree = (torch.sign(upd_got[big]) == torch.sign(upd_ref[big])).float().mean()
            assert agree > 0.95, (k, float(agree))


def test_adam_step_count_is_per_parameter():
    """A parameter whose gradient first appears after K steps starts at step 1 (torch.optim.Adam keeps state['step'] per
    parameter): its first update is ~lr, not the ~16x larger step a shared counter would give with beta2 = 0.999."""
    from semisupervisedobjectdetection_b200.optim import FusedClipAdam
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(64, 8, device="cuda"))
    b = torch.nn.Parameter(torch.randn(24, device="cuda"))
    ref_a, ref_b = (torch.nn.Parameter(t.detach().clone()) for t in (a, b))
    lr = 1e-2
    opt = FusedClipAdam([a, b], lr=lr, betas=(0.5, 0.999), weight_decay=5e-5, clip_value=1.2)
    ref = torch.optim.Adam([ref_a, ref_b], lr=lr, betas=(0.5, 0.999), weight_decay=5e-5)
    for it in range(6):
        opt.zero_grad()
        ref.zero_grad(set_to_none=True)
        ga = torch.randn(64, 8, device="cuda") * 0.3
        a.grad.copy_(ga)
        a._sfk_live = True
        ref_a.grad = ga.clone().clamp_(-1.2, 1.2)
        if it >= 4:   # b becomes live only from the fifth step on
            gb = torch.randn(24, device="cuda") * 0.3
            b.grad.copy_(gb)
            b._sfk_live = True
            ref_b.grad = gb.clone().clamp_(-1.2, 1.2)
        opt.step()
        ref.step()
        assert torch.allclose(a.detach(), ref_a.detach(), atol=2e-6, rtol=1e-5), it
        assert torch.allclose(b.detach(), ref_b.detach(), atol=2e-6, rtol=1e-5), it
    assert opt._steps == [6, 2]


def test_optimizer_state_dict_round_trip_matches_torch_layout():
    from semisupervisedobjectdetection_b200.optim import FusedClipAdam
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(16, 8, device="cuda")), torch.nn.Parameter(torch.randn(40, device="cuda"))]
    opt = FusedClipAdam(ps, lr=1e-2, weight_decay=0.0)
    for _ in range(3):
        opt.zero_grad()
        for p in ps:
            p.grad.copy_(torch.randn_like(p))
            p._sfk_live = True
        opt.step()
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sd["state"][0]["step"]) == 3.0 and sd["state"][1]["exp_avg"].shape == (40,)
    # a torch.optim.Adam over same-shaped parameters accepts it as is
    twin = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1e-2, betas=(0.5, 0.999))
    twin.load_state_dict(sd)
    # and a fresh FusedClipAdam continues bit-identically from it
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = FusedClipAdam(qs, lr=1e-2, weight_decay=0.0)
    opt2.load_state_dict(sd)
    g = [torch.randn_like(p) for p in ps]
    for o, params in ((opt, ps), (opt2, qs)):
        o.zero_grad()
        for p, gi in zip(params, g):
            p.grad.copy_(gi)
            p._sfk_live = True
        o.step()
    for p, q in zip(ps, qs):
        assert torch.equal(p.detach(), q.detach())


def test_weighted_dice_gives_rejected_tiles_no_loss_and_no_gradient():
    from semisupervisedobjectdetection_b200.loss import weighted_dice_loss
    p = synth.uniform01("wdice/p", (4, 48, 48)).cuda().requires_grad_(True)
    g = (synth.uniform01("wdice/g", (4, 48, 48)) > 0.5).to(torch.int64).cuda()
    keep = torch.tensor([1.0, 0.0, 1.0, 0.0], device="cuda")
    loss = weighted_dice_loss(p, g, keep)
    loss.backward()
    pc = p.detach().cpu().double().requires_grad_(True)
    gc = g.cpu().double()
    tp = (gc * pc).sum(dim=(1, 2)).abs()
    score = (2 * tp + 1) / (pc.abs().sum(dim=(1, 2)) + gc.sum(dim=(1, 2)) + 1)
    want = (keep.cpu().double() * (1 - score)).sum() / 4
    want.backward()
    assert abs(float(loss) - float(want)) < 1e-6
    assert torch.allclose(p.grad.cpu().double(), pc.grad, atol=1e-10, rtol=1e-4)
    assert float(p.grad[1].abs().max()) == 0.0 and float(p.grad[3].abs().max()) == 0.0


def test_raw_ops_reject_tensors_of_another_device():
    """libsfk launches on the current device; a tensor of another GPU must raise, never launch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from semisupervisedobjectdetection_b200 import _lib, ops
    x = torch.randn(64, 32, device="cuda:1")
    with pytest.raises(_lib.SfkError):
        ops.layernorm_fwd(ops._chk(x, torch.float32, "x"), torch.ones(32, device="cuda:1"), torch.zeros(32, device="cuda:1"))


def test_pipelined_loop_matches_the_sequential_steps():
    """PipelinedEmaLoop (uploads on a copy stream, losses read one step late) gives the same losses and weights as calling
    ema_teacher_student_step on device tensors one batch at a time (eval-mode statistics: no RNG in the step)."""
    from semisupervisedobjectdetection_b200 import teacher_student as TS
    from tests.test_gpu_model import _b0_wrapper

    def make_pair():
        s, _, _ = _b0_wrapper(lr=1e-3)
        t, _, _ = _b0_wrapper(lr=1e-3)
        for p in t.model.parameters():
            p.requires_grad_(False)
        s.model.eval(), t.model.eval()
        return t, s

    batches = []
    for i in range(4):
        batches.append((synth.uniform01(f"pipe{i}/l", (2, 3, 64, 64)).pin_memory(),
                        (synth.uniform01(f"pipe{i}/gt", (2, 64, 64)) > 0.7).float().pin_memory(),
                        synth.uniform01(f"pipe{i}/u", (2, 3, 64, 64)).pin_memory()))
    t1, s1 = make_pair()
    seq = [float(TS.ema_teacher_student_step(t1, s1, b[0].cuda(), b[1].cuda(), b[2].cuda())) for b in batches]
    t2, s2 = make_pair()
    loop = TS.PipelinedEmaLoop(t2, s2)
    piped = list(loop.run(batches))
    assert len(piped) == len(seq)
    for a, b in zip(seq, piped):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (seq, piped)
    # split-K weight gradients are summed with L2 atomics (run-to-run order differs) and Adam turns a sign flip of a ~0 gradient
    # into a full +-lr step, so the weights are compared in the mean, not bit for bit
    for a, b in ((s1, s2), (t1, t2)):
        diff = (a.optimizer.flat_p - b.optimizer.flat_p).abs()
        assert float(diff.mean()) < 1e-6 and float((diff > 1e-4).float().mean()) < 1e-3
    assert list(loop.run([])) == []
