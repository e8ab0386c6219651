from .rollouter import (Rollouter, SlotRollouter, SingleStepSlotRollouter,  # noqa: F401
                        get_sin_pos_enc, build_pos_enc)
from .slotformer import SlotFormer  # noqa: F401
from .single_step_slotformer import SingleStepSlotFormer  # noqa: F401
from .steve_slotformer import STEVESlotFormer  # noqa: F401


def build_model(params):
    """Same dispatch as reference video_prediction/models/__init__.py:6-36."""
    kwargs = dict(resolution=params.resolution, clip_len=params.input_frames,
                  slot_dict=params.slot_dict, dec_dict=params.dec_dict,
                  rollout_dict=params.rollout_dict, loss_dict=params.loss_dict)
    if params.model == 'SlotFormer':
        return SlotFormer(**kwargs)
    if params.model == 'SingleStepSlotFormer':
        return SingleStepSlotFormer(**kwargs)
    if params.model == 'STEVESlotFormer':
        # rollout only: the dVAE / token decoder are not built (steve_slotformer.py:105-109 is the whole rollout)
        return STEVESlotFormer(dvae_dict=params.dvae_dict, **kwargs)
    raise NotImplementedError(f'{params.model} is not implemented.')
