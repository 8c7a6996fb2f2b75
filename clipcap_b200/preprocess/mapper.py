"""EncoderMapper with the reference's interface (clipcap/preprocess/mapper.py:8-24): item {"data_tensor", "text"} ->
{"embeddings": numpy [N, E], "text"}. `model` is the encode function returned by get_encoder (cc_vit_forward). The
host->device copy goes through a pinned staging buffer on the current stream; the device->host copy of the embeddings is
the only synchronisation point."""
from __future__ import annotations

import torch
from torch.nn import Module


class EncoderMapper:
    def __init__(self, model: Module, device: str = "cuda"):
        self.model = model
        self.device = device
        self._pinned = None

    def __call__(self, item):
        with torch.no_grad():
            data = item["data_tensor"]
            if data.device.type == "cpu" and not data.is_pinned() and torch.cuda.is_available():
                if self._pinned is None or self._pinned.shape != data.shape or self._pinned.dtype != data.dtype:
                    self._pinned = torch.empty_like(data).pin_memory()
                self._pinned.copy_(data)
                data = self._pinned
            features = self.model(data.to(self.device, non_blocking=True))
            embeddings = features.cpu().numpy()
            return {"embeddings": embeddings, "text": item["text"]}
