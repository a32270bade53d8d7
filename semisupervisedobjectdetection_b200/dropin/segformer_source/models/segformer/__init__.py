"""Same public names as segformer_source/models/segformer/__init__.py:25-52 of the reference."""
from semisupervisedobjectdetection_b200.configuration import SegformerConfig  # noqa: F401
from semisupervisedobjectdetection_b200.modeling import (  # noqa: F401
    SegformerDecodeHead,
    SegformerForImageClassification,
    SegformerForSemanticSegmentation,
    SegformerLayer,
    SegformerModel,
    SegformerPreTrainedModel,
)
