"""SegformerConfig mirror (field names and defaults of
segformer_source/models/segformer/configuration_segformer.py:105-154 in the reference), without the HF plumbing."""
from __future__ import annotations

import copy
import warnings

# architecture table from segformer_source/models/segformer/convert_segformer_original_to_pytorch.py:158-178
MIT_VARIANTS = {
    "b0": dict(depths=[2, 2, 2, 2], hidden_sizes=[32, 64, 160, 256], decoder_hidden_size=256),
    "b1": dict(depths=[2, 2, 2, 2], hidden_sizes=[64, 128, 320, 512], decoder_hidden_size=256),
    "b2": dict(depths=[3, 4, 6, 3], hidden_sizes=[64, 128, 320, 512], decoder_hidden_size=768),
    "b3": dict(depths=[3, 4, 18, 3], hidden_sizes=[64, 128, 320, 512], decoder_hidden_size=768),
    "b4": dict(depths=[3, 8, 27, 3], hidden_sizes=[64, 128, 320, 512], decoder_hidden_size=768),
    "b5": dict(depths=[3, 6, 40, 3], hidden_sizes=[64, 128, 320, 512], decoder_hidden_size=768),
}


class SegformerConfig:
    model_type = "segformer"

    def __init__(self, num_channels=3, num_encoder_blocks=4, depths=(2, 2, 2, 2), sr_ratios=(8, 4, 2, 1),
                 hidden_sizes=(32, 64, 160, 256), patch_sizes=(7, 3, 3, 3), strides=(4, 2, 2, 2),
                 num_attention_heads=(1, 2, 5, 8), mlp_ratios=(4, 4, 4, 4), hidden_act="gelu", hidden_dropout_prob=0.0,
                 attention_probs_dropout_prob=0.0, classifier_dropout_prob=0.1, initializer_range=0.02, drop_path_rate=0.1,
                 layer_norm_eps=1e-6, decoder_hidden_size=256, semantic_loss_ignore_index=255, **kwargs):
        if kwargs.get("reshape_last_stage", True) is False:
            warnings.warn("Reshape_last_stage is set to False in this config. This argument is deprecated.", FutureWarning)
        self.num_channels = num_channels
        self.num_encoder_blocks = num_encoder_blocks
        self.depths = list(depths)
        self.sr_ratios = list(sr_ratios)
        self.hidden_sizes = list(hidden_sizes)
        self.patch_sizes = list(patch_sizes)
        self.strides = list(strides)
        self.mlp_ratios = list(mlp_ratios)
        self.num_attention_heads = list(num_attention_heads)
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.classifier_dropout_prob = classifier_dropout_prob
        self.initializer_range = initializer_range
        self.drop_path_rate = drop_path_rate
        self.layer_norm_eps = layer_norm_eps  # kept for API parity; the reference modules ignore it (eps stays 1e-5)
        self.decoder_hidden_size = decoder_hidden_size
        self.reshape_last_stage = kwargs.get("reshape_last_stage", True)
        self.semantic_loss_ignore_index = semantic_loss_ignore_index
        # PretrainedConfig attributes the reference code reads
        self.num_labels = kwargs.get("num_labels", 2)
        self.output_hidden_states = kwargs.get("output_hidden_states", False)
        self.output_attentions = kwargs.get("output_attentions", False)
        self.use_return_dict = kwargs.get("return_dict", True)
        self.problem_type = kwargs.get("problem_type", None)

    @classmethod
    def from_variant(cls, name: str, **overrides) -> "SegformerConfig":
        """`name` like 'nvidia/mit-b5', 'b5', 'nvidia/segformer-b2-finetuned-ade-512-512'."""
        import re
        low = name.lower().rstrip("/")
        # the variant tag is a whole path/dash-separated token: 'mit-b5', 'segformer-b2-finetuned-...', 'b0'
        found = re.findall(r"(?:^|[-/_.])(b[0-5])(?=$|[-/_.])", low)
        key = found[-1] if found and len(set(found)) == 1 else None
        if key is not None and key not in MIT_VARIANTS:
            key = None
        if key is None:
            raise ValueError(f"unknown SegFormer variant in {name!r}; expected one of {sorted(MIT_VARIANTS)}")
        kw = copy.deepcopy(MIT_VARIANTS[key])
        kw.update(overrides)
        return cls(**kw)

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def __repr__(self):
        return f"SegformerConfig({self.to_dict()})"
