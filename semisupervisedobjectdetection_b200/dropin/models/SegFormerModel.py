from semisupervisedobjectdetection_b200.wrapper import SegFormerModel  # noqa: F401
