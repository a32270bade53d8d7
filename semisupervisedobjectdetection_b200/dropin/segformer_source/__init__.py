"""Drop-in stand-in for the reference's vendored `segformer_source` package (only the SegFormer model is provided)."""
