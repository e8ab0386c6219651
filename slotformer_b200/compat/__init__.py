"""Compatibility layer so reference-style code (``from nerv.training import
BaseParams``) keeps working without the un-vendored ``nerv`` dependency."""
import importlib
import sys


def install_nerv_shim(force=False):
    """Register ``slotformer_b200.compat.nerv`` as top-level ``nerv``.

    Returns True if the shim was installed, False if a real nerv was found.
    """
    if not force:
        if 'nerv' in sys.modules:
            return getattr(sys.modules['nerv'], '__name__', '') != 'nerv'
        try:
            importlib.import_module('nerv')
            return False
        except ImportError:
            pass
    from . import nerv as shim
    sys.modules['nerv'] = shim
    for sub in ('training', 'models', 'utils'):
        sys.modules[f'nerv.{sub}'] = getattr(shim, sub)
    return True
