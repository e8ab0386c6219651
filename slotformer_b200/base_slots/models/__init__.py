from .slot_attention import SlotAttention, SlotAttentionWMask  # noqa: F401
from .savi import StoSAVi  # noqa: F401
from .steve import STEVE  # noqa: F401
from .utils import get_lr, to_rgb_from_tensor, assert_shape, SoftPositionEmbed  # noqa: F401


def build_model(params):
    """Same dispatch as reference base_slots/models/__init__.py:9-34."""
    if params.model == 'StoSAVi':
        return StoSAVi(resolution=params.resolution, clip_len=params.input_frames,
                       slot_dict=params.slot_dict, enc_dict=params.enc_dict,
                       dec_dict=params.dec_dict, pred_dict=params.pred_dict,
                       loss_dict=params.loss_dict)
    if params.model == 'STEVE':
        # slot-extraction half (encoder, SlotAttentionWMask, predictor); the dVAE / SLATE decoder are not built
        return STEVE(resolution=params.resolution, clip_len=params.input_frames, slot_dict=params.slot_dict,
                     dvae_dict=params.dvae_dict, enc_dict=params.enc_dict, dec_dict=params.dec_dict,
                     pred_dict=params.pred_dict, loss_dict=params.loss_dict)
    if params.model == 'dVAE':
        raise NotImplementedError(
            'dVAE: the image tokenizer is a different model, outside the hot-path scope (SURVEY.md section 2, row 12).')
    raise NotImplementedError(f'{params.model} is not implemented.')
