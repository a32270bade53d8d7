"""Drop-in stand-in for the reference's `models` package (SegFormerModel and Loss)."""
