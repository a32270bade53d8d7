"""CPU: host-side mirror of the reference interface (names, signatures, state-dict keys) and the data-parallel plumbing."""
import inspect
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import segformer_oracle as O
from semisupervisedobjectdetection_b200 import modeling
from semisupervisedobjectdetection_b200.configuration import MIT_VARIANTS, SegformerConfig
from semisupervisedobjectdetection_b200.optim import _exec_order_key


def test_state_dict_keys_and_shapes_match_reference_layout():
    for kw in (dict(num_labels=1), dict(num_labels=3, depths=[3, 6, 40, 3], hidden_sizes=[64, 128, 320, 512], decoder_hidden_size=768)):
        model = modeling.SegformerForSemanticSegmentation(SegformerConfig(**kw))
        ocfg = O.OracleConfig()
        for k, v in kw.items():
            setattr(ocfg, k, v)
        tmpl = O.state_dict_template(ocfg)  # pinned to the reference through tests/test_oracle_golden.py
        sd = model.state_dict()
        assert set(sd) == set(tmpl)
        assert all(tuple(sd[k].shape) == tuple(tmpl[k].shape) for k in sd)
    assert len(sd) == 1180  # MiT-b5 state dict has 1180 tensors


def test_variant_table_and_param_count():
    cfg = SegformerConfig.from_variant("nvidia/mit-b5", num_labels=1)
    assert cfg.depths == [3, 6, 40, 3] and cfg.hidden_sizes == [64, 128, 320, 512] and cfg.decoder_hidden_size == 768
    model = modeling.SegformerForSemanticSegmentation(cfg)
    assert sum(p.numel() for p in model.parameters()) == 86170049
    assert set(MIT_VARIANTS) == {"b0", "b1", "b2", "b3", "b4", "b5"}
    with pytest.raises(ValueError):
        SegformerConfig.from_variant("resnet50")


def test_module_signatures_follow_reference():
    sig = inspect.signature
    assert list(sig(modeling.SegformerOverlapPatchEmbeddings.__init__).parameters)[1:] == ["patch_size", "stride", "num_channels", "hidden_size"]
    assert list(sig(modeling.SegformerEfficientSelfAttention.__init__).parameters)[1:] == ["config", "hidden_size", "num_attention_heads", "sequence_reduction_ratio"]
    assert list(sig(modeling.SegformerEfficientSelfAttention.forward).parameters)[1:] == ["hidden_states", "height", "width", "prompt_token_len", "output_attentions"]
    assert list(sig(modeling.SegformerLayer.__init__).parameters)[1:] == ["config", "hidden_size", "num_attention_heads", "drop_path", "sequence_reduction_ratio", "mlp_ratio"]
    assert list(sig(modeling.SegformerLayer.forward).parameters)[1:] == ["hidden_states", "height", "width", "prompt_token_len", "output_attentions"]
    assert list(sig(modeling.SegformerMixFFN.__init__).parameters)[1:] == ["config", "in_features", "hidden_features", "out_features"]
    assert list(sig(modeling.SegformerMLP.__init__).parameters)[1:] == ["config", "input_dim", "cls_dim"]
    assert list(sig(modeling.SegformerDecodeHead.forward).parameters)[1:] == ["encoder_hidden_states", "cls_token"]
    assert list(sig(modeling.SegformerForSemanticSegmentation.forward).parameters)[1:] == ["pixel_values", "labels", "output_attentions", "output_hidden_states", "return_dict"]
    enc = modeling.SegformerEncoder(SegformerConfig())
    assert enc.prompt_tokens == [None] * 4 and enc.cls_token == [None] * 4  # plain Python lists (MS:391-395)
    assert isinstance(enc.block[0], torch.nn.ModuleList)


def test_wrapper_surface():
    from semisupervisedobjectdetection_b200.wrapper import SegFormerModel
    for name in ("predict", "train_one_epoch", "eval_one_epoch", "train_from_loss", "scheduler_step", "frozen_encoder",
                 "unfroze_encoder", "add_prompt_token", "add_cls_token", "show_mask", "train_one_epoch_without_mask",
                 "eval_one_epoch_without_mask"):
        assert callable(getattr(SegFormerModel, name))
    params = list(inspect.signature(SegFormerModel.__init__).parameters)
    assert params[1:8] == ["pretrain_weight", "lr", "weight_decay", "scheduler", "device", "use_dice_loss", "num_labels"]
    assert list(inspect.signature(SegFormerModel.predict).parameters)[1:] == ["img", "mask", "isEval", "use_loss", "output_cls_token"]


def test_dropin_packages_resolve():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "semisupervisedobjectdetection_b200", "dropin")
    sys.path.insert(0, root)
    try:
        for mod in ("segformer_source", "models"):
            for k in [k for k in sys.modules if k == mod or k.startswith(mod + ".")]:
                del sys.modules[k]
        from segformer_source.models.segformer import SegformerConfig as C2, SegformerForSemanticSegmentation as M2  # noqa
        import models.Loss as myLoss
        from models.SegFormerModel import SegFormerModel as SegModel
        assert M2 is modeling.SegformerForSemanticSegmentation
        assert hasattr(myLoss, "SegmentationLoss") and hasattr(myLoss, "diceCoeff") and hasattr(myLoss, "mse")
        assert SegModel.__name__ == "SegFormerModel"
    finally:
        sys.path.remove(root)
        for mod in ("segformer_source", "models"):
            for k in [k for k in sys.modules if k == mod or k.startswith(mod + ".")]:
                del sys.modules[k]


def test_execution_order_key():
    names = ["segformer.encoder.layer_norm.0.weight", "decode_head.classifier.weight", "segformer.encoder.block.1.0.mlp.dense1.weight",
             "segformer.encoder.patch_embeddings.1.proj.weight", "segformer.encoder.block.0.2.layer_norm_1.weight",
             "segformer.encoder.patch_embeddings.0.proj.weight"]
    ordered = sorted(names, key=_exec_order_key)
    assert ordered[0].startswith("segformer.encoder.patch_embeddings.0") and ordered[1].startswith("segformer.encoder.block.0.2")
    assert ordered[2].startswith("segformer.encoder.layer_norm.0") and ordered[3].startswith("segformer.encoder.patch_embeddings.1")
    assert ordered[-1].startswith("decode_head")


def _ddp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semisupervisedobjectdetection_b200.ddp import GradBucketer, broadcast_flat
    n = 10_000
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    b = GradBucketer(flat, bucket_bytes=4096 * 4)
    b.begin_step()
    b.notify(5000)          # buckets entirely behind offset 5000 go first (overlap path)
    launched_early = b._next
    b.finish()
    expect = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, expect)
    p = torch.full((16,), float(rank))
    broadcast_flat(p, 0)
    ok = ok and bool((p == 0).all()) and launched_early == 1 and len(b.buckets) == 3
    out[rank] = ok
    dist.destroy_process_group()


def test_grad_bucketer_allreduce_world2_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


# ---------------------------------------------------------------------------------------------------------------
# GEMM launch planner (host logic of csrc/gemm_sm100.cu, reachable without a GPU through sfk_gemm_plan)
# ---------------------------------------------------------------------------------------------------------------
def _plan(M, N, K, a_mn=0, b_mn=0, out_bf16=1, atomic=0, splits=0, block_n=0):
    import ctypes
    from semisupervisedobjectdetection_b200 import _lib as L
    lib = L.load()
    out = (ctypes.c_int * 8)()
    rc = lib.sfk_gemm_plan(M, N, K, a_mn, b_mn, out_bf16, atomic, splits, block_n, out)
    assert rc == 0, L.last_error()
    keys = ("block_n", "splits", "stages", "grid", "mode", "tmem_cols", "smem", "tiles")
    d = dict(zip(keys, list(out)))
    # mode bits: 1 = CTA pair (cta_group::2), 2 = B-stationary main loop, 4 = TMA-store epilogue
    d["pair"], d["bstat"], d["epi_tma"] = d["mode"] & 1, (d["mode"] >> 1) & 1, (d["mode"] >> 2) & 1
    return d


def test_gemm_plan_invariants_over_the_mit_shapes():
    """Every forward / dgrad / wgrad shape of MiT-b0..b5 at 512x512 (batch 16): legal tile widths, pipeline depth, grids that
    fit the machine, shared memory within the 227 KB budget; K <= 512 non-split GEMMs keep their weight tile resident with
    at least 3 A stages; split-K only with atomics."""
    SMS = 148   # num_sms() falls back to the B200 count without a device
    stages_cfg = [(262144, 64, 256), (65536, 128, 512), (16384, 320, 1280), (4096, 512, 2048),     # b1..b5: (tokens, C, 4C)
                  (262144, 32, 128), (65536, 64, 256), (16384, 160, 640), (4096, 256, 1024)]       # b0
    shapes = []
    for T, C, H4 in stages_cfg:
        shapes += [(T, C, C, 0, 0, 0), (T, H4, C, 0, 0, 0), (T, C, H4, 0, 0, 0),           # q/o, dense1, dense2
                   (T, C, C, 0, 1, 0), (T, C, H4, 0, 1, 0), (T, H4, C, 0, 1, 0),           # dgrads (MN-major B)
                   (C, C, T, 1, 1, 1), (H4, C, T, 1, 1, 1), (C, H4, T, 1, 1, 1)]           # wgrads (split-K, atomics)
    shapes += [(262144, 768, 3072, 0, 0, 0), (262144, 3072, 768, 0, 1, 0), (768, 3072, 262144, 1, 1, 1)]   # literal decode head
    for (M, N, K, a_mn, b_mn, atomic) in shapes:
        p = _plan(M, N, K, a_mn=a_mn, b_mn=b_mn, out_bf16=0 if atomic else 1, atomic=atomic)
        kb = (K + 63) // 64
        assert p["block_n"] % 32 == 0 and 32 <= p["block_n"] <= 256, (M, N, K, p)
        if b_mn:
            assert p["block_n"] % 64 == 0, (M, N, K, p)
        assert 2 <= p["stages"] <= 12 and 1 <= p["splits"] <= kb, (M, N, K, p)
        assert p["splits"] == 1 or atomic, (M, N, K, p)
        assert p["tmem_cols"] == 512 and 2 * p["block_n"] <= 512, (M, N, K, p)
        assert 1 <= p["grid"] <= SMS, (M, N, K, p)
        assert p["smem"] <= 227 * 1024, (M, N, K, p)
        if p["pair"]:
            assert p["grid"] % 2 == 0 and (not b_mn or p["block_n"] % 128 == 0), (M, N, K, p)
        assert p["epi_tma"], (M, N, K, p)   # identity row maps: every mode stores through staged TMA tiles
        if kb <= 2 and not atomic:
            assert p["bstat"] and p["stages"] >= 3, (M, N, K, p)
            # resident weight tile + A ring + 64 KB of epilogue staging (two 4 KB buffers per epilogue warp) fit
            assert p["block_n"] * kb * 128 + p["stages"] * 16384 + 65536 <= p["smem"], (M, N, K, p)
        else:
            assert not p["bstat"], (M, N, K, p)


def test_gemm_plan_split_k_takes_whole_rounds():
    """The wgrad split count follows the rounds model: the stage-3 MixFFN wgrads fit ONE round of the 148 SMs (the old
    'fill 2 x SMs' rule produced 300 units = 3 rounds), and tiny-output wgrads still spread over the machine."""
    for (M, N, K) in [(1280, 320, 16384), (320, 1280, 16384), (320, 320, 16384)]:
        p = _plan(M, N, K, a_mn=1, b_mn=1, out_bf16=0, atomic=1)
        units = p["tiles"] * p["splits"]
        assert units <= 148 and units >= 148 * 0.8, (M, N, K, p)
    p = _plan(64, 64, 262144, a_mn=1, b_mn=1, out_bf16=0, atomic=1)      # one output tile, 4096 k-blocks
    assert p["tiles"] == 1 and 100 <= p["splits"] <= 296, p
    p = _plan(128, 64, 64, a_mn=1, b_mn=1, out_bf16=0, atomic=1)         # nothing to split
    assert p["splits"] == 1, p


def test_gemm_plan_modes():
    """CTA pairs only for the large problems (default option), the B-stationary loop only for K <= 512 non-split GEMMs (with
    a tile width that covers N without padding where one exists), explicit requests honoured."""
    big = _plan(262144, 768, 3072)
    assert big["pair"] == 1 and big["block_n"] == 256 and big["tmem_cols"] == 512 and not big["bstat"], big
    small_k = _plan(262144, 256, 64)
    assert small_k["pair"] == 0 and small_k["bstat"] == 1 and 256 % small_k["block_n"] == 0 and small_k["epi_tma"] == 1, small_k
    mid_k = _plan(16384, 1280, 320)
    assert mid_k["bstat"] == 0 and mid_k["epi_tma"] == 1, mid_k
    forced_ws = _plan(16384, 320, 320, block_n=160)
    assert forced_ws["bstat"] == 1 and forced_ws["block_n"] == 160 and forced_ws["stages"] >= 3, forced_ws
    long_k = _plan(16384, 320, 1280)
    assert long_k["pair"] == 0 and long_k["bstat"] == 0 and long_k["epi_tma"] == 1, long_k
    forced = _plan(4096, 256, 256, block_n=64, splits=0)
    assert forced["block_n"] == 64, forced
    import ctypes
    from semisupervisedobjectdetection_b200 import _lib as L
    out = (ctypes.c_int * 8)()
    assert L.load().sfk_gemm_plan(128, 64, 64, 0, 0, 1, 1, 0, 0, out) != 0     # atomics need an fp32 output
    assert b"atomic" in L.load().sfk_last_error()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores) prints ONE JSON line carrying the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mit_b5_512_teacher_student_train_images_per_sec"
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_variant_names_are_parsed_as_whole_tokens():
    for name, depths in (("nvidia/mit-b5", [3, 6, 40, 3]), ("b0", [2, 2, 2, 2]), ("nvidia/segformer-b2-finetuned-ade-512-512", [3, 4, 6, 3]),
                         ("/ckpt/mit-b1/", [2, 2, 2, 2])):
        assert SegformerConfig.from_variant(name).depths == depths
    for bad in ("foo", "rgb5", "mit-b7", "b0-b1"):
        with pytest.raises(ValueError):
            SegformerConfig.from_variant(bad)


def test_from_pretrained_warns_on_random_init_and_loads_local_checkpoints(tmp_path, monkeypatch):
    """SFM:14-16 starts from the ImageNet-pretrained encoder; offline that is only possible from a local file, and a
    silent random init would change what the reference mains train."""
    cls = modeling.SegformerForSemanticSegmentation
    monkeypatch.delenv("SFK_PRETRAINED_DIR", raising=False)
    with pytest.warns(RuntimeWarning, match="RANDOM"):
        net = cls.from_pretrained("nvidia/mit-b0", num_labels=1)
    sd = {k: torch.full_like(v, 0.5) if v.is_floating_point() else v for k, v in net.state_dict().items()}
    # (a) the reference wrapper's 'model.'-prefixed state dict, classifier of another width skipped
    ref_style = {"model." + k: v for k, v in sd.items()}
    ref_style["model.decode_head.classifier.weight"] = torch.zeros(7, 256, 1, 1)
    path = tmp_path / "wrapper_style.pth"
    torch.save(ref_style, path)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        got = cls.from_pretrained(str(path), ignore_mismatched_sizes=True, num_labels=1, variant="b0")
    assert float(got.segformer.encoder.block[1][0].mlp.dense1.weight.mean()) == 0.5
    assert float(got.decode_head.classifier.weight.abs().max()) < 0.5   # mismatched shape skipped, init kept
    # (b) encoder-only dump in a Hugging Face style snapshot directory found through SFK_PRETRAINED_DIR
    snap = tmp_path / "hub" / "mit-b0"
    snap.mkdir(parents=True)
    torch.save({k[len("segformer."):]: v for k, v in sd.items() if k.startswith("segformer.")}, snap / "pytorch_model.bin")
    monkeypatch.setenv("SFK_PRETRAINED_DIR", str(tmp_path / "hub"))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        got = cls.from_pretrained("nvidia/mit-b0", num_labels=1)
    assert float(got.segformer.encoder.patch_embeddings[0].proj.weight.mean()) == 0.5
    assert float(got.decode_head.linear_fuse.weight.abs().max()) < 0.5       # head not in the dump: random init kept


def test_semantic_segmenter_output_builds_hidden_states_lazily():
    """The wrapper switches output_hidden_states on like the reference (SFM:17) but no script reads the maps: they are built on
    first access, once, under the grad mode of the forward call; attribute / index / tuple access keep working."""
    import torch
    from semisupervisedobjectdetection_b200.modeling import SemanticSegmenterOutput
    calls = []

    def build():
        calls.append(torch.is_grad_enabled())
        return (torch.ones(1), torch.zeros(1))

    with torch.no_grad():
        out = SemanticSegmenterOutput(loss=None, logits=torch.full((1,), 2.0), hidden_states=build)
    assert calls == []                                   # nothing materialised by the forward call
    assert float(out.logits) == 2.0 and out["logits"] is out.logits
    hs = out.hidden_states
    assert calls == [False] and len(hs) == 2             # built once, with grad disabled as in the forward call
    assert out.hidden_states is hs and calls == [False]
    assert out.to_tuple()[0] is out.logits and out.to_tuple()[1] is hs and out[1] is hs
    out.hidden_states = None
    assert out.hidden_states is None and out.to_tuple() == (out.logits,)
    assert SemanticSegmenterOutput(logits=torch.ones(1)).hidden_states is None
