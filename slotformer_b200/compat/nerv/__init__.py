"""Minimal stand-in for the un-vendored third-party package ``nerv`` (v0.1.0).

The reference imports ``nerv`` for its trainer base classes and conv helpers
(reference: slotformer/base_slots/models/savi.py:8-9, every configs/*.py:1,
scripts/train.py:12-13).  ``nerv`` is not shipped with the reference and cannot
be installed offline, so this shim provides exactly the symbols the hot-path
callers and the config files touch.  Behaviour is derived from the reference's
call sites; nothing here is on the CUDA hot path.

Install with ``slotformer_b200.compat.install_nerv_shim()`` -- it is a no-op if
a real ``nerv`` is importable.
"""
from . import training, models, utils  # noqa: F401
