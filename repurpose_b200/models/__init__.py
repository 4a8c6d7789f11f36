"""Host-side mirror of the reference's `models/` package for the inference hot path
(models/MMCTransformer.py, models/softnms.py, MultiHeadAttention of models/transformer.py)."""
