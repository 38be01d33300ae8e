"""Drop-in replacements for the hot-path modules of the reference's ``src`` package.

Module and function names mirror the reference so that code written against
``src.SEDNet``, ``src.PointNet``, ``src.mean_shift``, ``src.primitive_forward``, ``src.fitting_utils``,
``src.primitives`` and ``src.segment_utils`` runs unchanged; every function dispatches to the sm_100a kernels in
``libsednet_b200.so`` through the C ABI of ``include/sednet_b200.h`` and raises if the library or a GPU is missing
(there is no CPU fallback).
"""
