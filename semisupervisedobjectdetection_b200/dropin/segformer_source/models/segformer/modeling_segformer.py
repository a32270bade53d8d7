from semisupervisedobjectdetection_b200.modeling import *  # noqa: F401,F403
from semisupervisedobjectdetection_b200.modeling import (  # noqa: F401
    SegformerAttention, SegformerDWConv, SegformerDropPath, SegformerEfficientSelfAttention, SegformerEncoder, SegformerMLP,
    SegformerMixFFN, SegformerOverlapPatchEmbeddings, SegformerSelfOutput, drop_path_scale)
