"""Importable alias for the hyphenated package directory ``fast-dnn_b200/``."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.modules[__name__] = importlib.import_module("fast-dnn_b200")
