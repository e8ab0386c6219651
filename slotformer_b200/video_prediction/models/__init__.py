from .rollouter import (Rollouter, SlotRollouter, SingleStepSlotRollouter,  # noqa: F401
                        get_sin_pos_enc, build_pos_enc)
