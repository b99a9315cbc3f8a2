from dataclasses import dataclass
from typing import Optional

import torch

from ...utils import BaseOutput


@dataclass
class EncoderOutput(BaseOutput):
    latent: torch.Tensor


@dataclass
class DecoderOutput(BaseOutput):
    sample: torch.Tensor
    commit_loss: Optional[torch.FloatTensor] = None
