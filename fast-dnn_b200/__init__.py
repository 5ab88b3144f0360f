"""fast-dnn_b200 — B200-native drop-in for fast-dnn's quantized feed-forward inference path.

The directory name carries a hyphen (it mirrors the reference's ``fast-dnn``); import it with
``importlib.import_module("fast-dnn_b200")`` or through the ``fast_dnn_b200`` alias module at the
repo root.
"""
from . import formats, synth  # noqa: F401
