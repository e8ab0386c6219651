"""PHYRE variant: rollout conditioned on the first frame only (reference
slotformer/video_prediction/models/single_step_slotformer.py:93-140)."""
import torch

from .rollouter import SingleStepSlotRollouter
from .slotformer import SlotFormer


class SingleStepSlotFormer(SlotFormer):

    def _build_loss(self):
        super()._build_loss()
        self.use_cls_loss = False       # a task-success classifier may be attached for PHYRE eval
        self.success_cls = None

    def _build_rollouter(self):
        self.history_len = self.rollout_dict['history_len']     # 1
        self.rollouter = SingleStepSlotRollouter(**self.rollout_dict)

    def classify(self, slots, vid_len=None):
        assert not self.training
        return self.success_cls({'slots': slots, 'vid_len': vid_len})['logits']

    def forward(self, data_dict):
        out = super().forward(data_dict)
        if self.use_cls_loss and self.success_cls is not None:
            slots = torch.cat([out['gt_slots'], out['pred_slots']], dim=1)
            out['logits'] = self.classify(slots, data_dict.get('vid_len', None))
        return out
