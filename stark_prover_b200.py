"""Importable alias for the package directory `stark-prover_b200/` (its name contains a hyphen)."""
import importlib
import sys

_pkg = importlib.import_module("stark-prover_b200")
sys.modules[__name__] = _pkg
