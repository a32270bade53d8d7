from semisupervisedobjectdetection_b200.configuration import SegformerConfig  # noqa: F401
