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
