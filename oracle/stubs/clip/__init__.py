"""Test-infrastructure stub for the `clip` module the reference imports at top level
(models/mdm.py:3).  Never imported by the product path."""
def load(*a, **k):
    raise RuntimeError("clip stub: weights are not available offline")
def tokenize(*a, **k):
    raise RuntimeError("clip stub")
