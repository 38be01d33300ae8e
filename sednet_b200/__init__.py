"""Importable alias of the product package.

The product directory is named ``sed-net_b200`` (not a valid Python identifier); this shim makes it importable as
``sednet_b200`` by pointing the package search path at it:  ``import sednet_b200.src.SEDNet``,
``from sednet_b200 import synth, pipeline``.
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "sed-net_b200"))
