from semisupervisedobjectdetection_b200.loss import SegmentationLoss, diceCoeff, mse  # noqa: F401
