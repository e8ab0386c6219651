from .slot_attention import SlotAttention, SlotAttentionWMask  # noqa: F401
