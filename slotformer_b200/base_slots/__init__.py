"""Task package surface of reference slotformer/base_slots/__init__.py (build_model only:
datasets and the nerv trainer are out of scope)."""
from .models import build_model  # noqa: F401


def build_dataset(params, val_only=False):
    raise NotImplementedError('datasets are out of scope of the sm_100a engine (SURVEY.md section 2, row 20)')


def build_method(**kwargs):
    raise NotImplementedError('the nerv trainer is out of scope of the sm_100a engine (SURVEY.md section 2, row 14)')
