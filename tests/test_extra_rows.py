"""Rows next to the hot path: few-shot cosine losses and the device input pipeline.
CPU part: the oracle restatements against the fixtures made from the reference / cv2 (tests/golden/make_golden_extra.py).
GPU part: the libsfk kernels against the same fixtures and the oracle."""
import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as PO
from oracle import segformer_oracle as O
from oracle import synth
from tests import helpers
from tests.golden import make_golden_extra as G

FS_CASES = {"b6": (6, 6), "b5_b4": (5, 4), "b2": (2, 2)}


def _fs_inputs(tag):
    b1, b2 = FS_CASES[tag]
    c1 = torch.sigmoid((synth.uniform01(f"fs/{tag}/c1", (b1, 1, 512)) - 0.5) * 4)
    c2 = torch.sigmoid((synth.uniform01(f"fs/{tag}/c2", (b2, 1, 512)) - 0.5) * 4)
    return c1, c2, torch.tensor(0.0371), torch.tensor(0.0412)


@pytest.mark.parametrize("tag", list(FS_CASES))
def test_oracle_fewshot_losses_match_reference(tag):
    gold = helpers.load_golden("extra")["fewshot"][tag]
    c1, c2, l1, l2 = _fs_inputs(tag)
    c1.requires_grad_(True), c2.requires_grad_(True)
    total, intra1, intra2, inter = O.fewshot_losses(l1, c1, l2, c2)
    total.backward()
    for got, key in ((total, "total"), (intra1, "intra1"), (intra2, "intra2"), (inter, "inter")):
        assert torch.allclose(got, gold[key], atol=1e-6, rtol=1e-6), key
    assert torch.allclose(c1.grad, gold["g1"], atol=1e-7, rtol=1e-5) and torch.allclose(c2.grad, gold["g2"], atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("name", list(G.PIPE_CASES))
def test_oracle_input_pipeline_matches_cv2_fixtures(name):
    """The numpy restatement of RandomCrop -> flip / rot90 -> /255 -> Resize against what cv2 produced (float tolerance 2e-6:
    cv2's SIMD resize rounds its float products in a different order; masks are exact)."""
    gold = helpers.load_golden("extra")["pipeline"][name]
    tiles, masks, params, crop, osz = G.pipe_inputs(name)
    for b in range(tiles.shape[0]):
        img, m = PO.augment_tile(tiles[b].numpy(), params[b].tolist(), crop, (osz, osz), masks[b].numpy())
        if name == "real":
            assert np.allclose(img[:, ::5, ::7], gold["img_sample"][b].numpy(), atol=2e-6)
            assert np.array_equal(m[::5, ::7], gold["mask_sample"][b].numpy())
            assert abs(img.astype(np.float64).sum() - float(gold["img_sum"][b])) < 1e-2
            assert m.astype(np.float64).sum() == float(gold["mask_sum"][b])
        else:
            assert np.allclose(img, gold["img"][b].numpy(), atol=2e-6)
            assert np.array_equal(m, gold["mask"][b].numpy())


def test_tile_augmenter_draws_legal_parameters():
    from semisupervisedobjectdetection_b200.pipeline import TileAugmenter
    aug = TileAugmenter()
    p = aug.draw_params(4096, 523, 530, torch.Generator().manual_seed(0))
    assert p.dtype == torch.int32 and p.shape == (4096, 4)
    assert int(p[:, 0].min()) >= 0 and int(p[:, 0].max()) <= 23 and int(p[:, 1].min()) >= 0 and int(p[:, 1].max()) <= 30
    ops_seen = set(p[:, 2].tolist())
    assert ops_seen == {0, 1, 2, 3, 4, 5, 6}
    frac_none = float((p[:, 2] == 0).float().mean())
    assert 0.2 < frac_none < 0.3          # OneOf(p=0.75) leaves a quarter of the tiles untouched


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(FS_CASES))
def test_gpu_fewshot_losses(tag):
    from semisupervisedobjectdetection_b200 import fewshot
    gold = helpers.load_golden("extra")["fewshot"][tag]
    c1, c2, l1, l2 = _fs_inputs(tag)
    c1, c2 = c1.cuda().requires_grad_(True), c2.cuda().requires_grad_(True)
    total, intra1, intra2, inter = fewshot.fewshot_losses(l1.cuda(), c1, l2.cuda(), c2)
    total.backward()
    for got, key in ((total, "total"), (intra1, "intra1"), (intra2, "intra2"), (inter, "inter")):
        assert torch.allclose(got.cpu(), gold[key], atol=2e-6, rtol=2e-6), key
    assert torch.allclose(c1.grad.cpu(), gold["g1"], atol=1e-6, rtol=1e-4)
    assert torch.allclose(c2.grad.cpu(), gold["g2"], atol=1e-6, rtol=1e-4)
    assert abs(float(fewshot.compute_similarity(c1.detach()[:min(FS_CASES[tag])], c2.detach()[:min(FS_CASES[tag])])) - float(gold["sim"])) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.PIPE_CASES))
def test_gpu_input_pipeline(name):
    """Device pipeline against the cv2 fixtures (2e-6) and against the oracle on the same tiles; masks bit-exact."""
    from semisupervisedobjectdetection_b200.pipeline import TileAugmenter
    gold = helpers.load_golden("extra")["pipeline"][name]
    tiles, masks, params, crop, osz = G.pipe_inputs(name)
    aug = TileAugmenter(crop=crop, out_hw=(osz, osz))
    img, m = aug(tiles.cuda(), params, masks.cuda())
    img, m = img.cpu(), m.cpu()
    assert img.shape == (tiles.shape[0], 3, osz, osz) and m.shape == (tiles.shape[0], osz, osz)
    for b in range(tiles.shape[0]):
        o_img, o_m = PO.augment_tile(tiles[b].numpy(), params[b].tolist(), crop, (osz, osz), masks[b].numpy())
        assert np.allclose(img[b].numpy(), o_img, atol=1e-6)
        assert np.array_equal(m[b].numpy(), o_m)
    if name == "real":
        assert torch.allclose(img[:, :, ::5, ::7], gold["img_sample"], atol=2e-6)
        assert torch.equal(m[:, ::5, ::7], gold["mask_sample"])
        assert torch.allclose(img.double().sum(dim=(1, 2, 3)), gold["img_sum"], atol=1e-2)
    else:
        assert torch.allclose(img, gold["img"], atol=2e-6) and torch.equal(m, gold["mask"])
    # images only (unlabeled tiles, DL:92-94)
    img2, none = aug(tiles.cuda(), params)
    assert none is None and torch.equal(img2.cpu(), img)


@pytest.mark.gpu
def test_gpu_train_autoencoder_iteration_runs_the_fewshot_path():
    """FS:191-237 end to end on a small cls-capable model: add_cls_token, num_labels = 3 'mse' reconstruction, cosine losses,
    one optimiser step through the summed loss; compared with the oracle's forward on the same weights."""
    from semisupervisedobjectdetection_b200 import fewshot
    from tests.test_gpu_optim import _wrapper_for
    cfg = dict(num_labels=3, hidden_sizes=[32, 64, 160, 512], depths=[1, 1, 1, 1], decoder_hidden_size=64)
    w, ocfg, sd = _wrapper_for(cfg)
    enc = w.model.segformer.encoder
    cls = [synth.uniform01(f"fsrun/cls{s}", (1, ocfg.hidden_sizes[s])) for s in range(4)]
    for s in range(4):
        enc.cls_token[s] = cls[s].clone().cuda().requires_grad_(True)
    x1 = synth.uniform01("fsrun/x1", (4, 3, 64, 64))
    x2 = synth.uniform01("fsrun/x2", (4, 3, 64, 64))
    w.model.train()
    out = fewshot.train_autoencoder_iteration(w, x1, x2)
    total, _, pred1, l1, l2, intra1, intra2, inter = out
    assert pred1.shape == (4, 3, 64, 64) and torch.isfinite(total)
    # oracle forward (eval-mode statistics, SURVEY M7: predict(isEval=False) switches the model to eval)
    sd_o = {k: v.clone() for k, v in sd.items()}
    toks, losses = [], []
    for x in (x1, x2):
        logits, _, cls_list = O.forward_logits(x, sd_o, ocfg, None, [c.clone() for c in cls])
        up = torch.nn.functional.interpolate(logits, size=(64, 64), mode="bilinear", align_corners=False)
        losses.append(O.segmentation_loss(x, up, "mse"))
        toks.append(torch.sigmoid(cls_list[-1]))
    want_total, want_i1, want_i2, want_inter = O.fewshot_losses(losses[0], toks[0], losses[1], toks[1])
    assert abs(float(l1) - float(losses[0])) < 2e-2 * abs(float(losses[0]))
    assert abs(float(inter) - float(want_inter)) < 2e-3 and abs(float(intra1) - float(want_i1)) < 2e-3
    assert abs(float(total) - float(want_total)) < 2e-2 * abs(float(want_total))
    before = w.optimizer.flat_p.clone()
    w.train_from_loss(total)
    assert not torch.equal(before, w.optimizer.flat_p)
