"""Case table shared by the fixture generator (make_golden.py, runs the reference) and the tests (oracle / CUDA path)."""

B0 = dict(num_labels=1)
NO_STOCHASTIC = dict(drop_path_rate=0.0, classifier_dropout_prob=0.0)

MODEL_CASES = {
    # eval mode, BCE-with-valid-mask path (MS:862-866) + Dice (SFM:129)
    "b0_eval": dict(config=dict(B0), batch=2, hw=64, labels="binary"),
    "b0_hw128": dict(config=dict(B0), batch=2, hw=128),
    # prompt tokens shared per stage (2-D) and per layer (3-D), MS:446-453
    "b0_prompt2d": dict(config=dict(B0), batch=2, hw=64, prompts=[(3, 32), (3, 64), (3, 160), (3, 256)]),
    "b0_prompt3d": dict(config=dict(B0), batch=2, hw=64, prompts=[(2, 3, 32), (2, 3, 64), (2, 3, 160), (2, 3, 256)]),
    # cls ("domain") token threading + cls_proj in the head; needs hidden_sizes[-1] == 512 (MS:733)
    "cls512": dict(config=dict(num_labels=1, hidden_sizes=[32, 64, 160, 512], depths=[1, 2, 1, 2], decoder_hidden_size=128),
                   batch=2, hw=64, cls=[1, 1, 1, 1], prompts=[(2, 32), None, (2, 160), None]),
    # multi-class CrossEntropy(ignore_index=255) path (MS:859-861)
    "nl3_ce": dict(config=dict(num_labels=3), batch=2, hw=64, labels="multi"),
    # train mode (BatchNorm batch statistics + running-stat update) with the stochastic parts disabled
    "b0_trainbn": dict(config=dict(num_labels=1, **NO_STOCHASTIC), batch=3, hw=64, train_mode=True),
}

GRAD_KEYS = (
    "patch_embeddings.0.proj.weight",
    "patch_embeddings.2.proj.bias",
    "block.0.0.attention.self.query.weight",
    "block.0.0.attention.self.sr.weight",
    "block.1.0.attention.self.layer_norm.weight",
    "block.2.1.mlp.dwconv.dwconv.weight",
    "block.2.1.mlp.dwconv.dwconv.bias",
    "block.3.1.mlp.dense2.weight",
    "block.3.0.attention.self.value.bias",
    "encoder.layer_norm.3.weight",
    "decode_head.linear_c.1.proj.weight",
    "decode_head.linear_c.3.cls_proj.weight",
    "decode_head.batch_norm.weight",
    "decode_head.classifier.weight",
)


def sample_flat(t, n=2048):
    """Strided sample of a (possibly large) tensor so that fixtures stay small; norms cover the rest."""
    flat = t.reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step].clone()
