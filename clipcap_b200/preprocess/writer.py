"""Writer of the reference's preprocess output layout (clipcap/preprocess/writer.py:10-100), so a dataset encoded with
the B200 ViT engine is read by the reference's training dataloader unchanged:

    <out>/encoder_config.yaml                      EncoderConfig.to_dict()
    <out>/embeddings/embeds_<id>.npy               [N, E] array, dtype as the encoder produced it
    <out>/captions/captions_<id>.parquet           one column `caption`, row i belongs to embedding row i

`<id>` is the partition number zero-padded to the digit count of the partition total. Paths go through fsspec (local
paths, s3://, ...), as in the reference. `NumpyWriter` keeps the reference's constructor and call protocol
(`writer(batch)` per encoded batch, `writer.flush()` once per partition)."""
from __future__ import annotations

import io
import math
from typing import List

import fsspec
import yaml

from clipcap_b200.configs import EncoderConfig


def save_config(config: EncoderConfig, output_folder: str) -> None:
    fs, root = fsspec.core.url_to_fs(output_folder)
    fs.makedirs(root, exist_ok=True)
    with fs.open(f"{root}/encoder_config.yaml", "w") as handle:
        yaml.dump(config.to_dict(), handle, default_flow_style=False)


def _partition_tag(partition_id: int, partition_total: int) -> str:
    return str(partition_id).zfill(int(math.log10(partition_total)) + 1)


class NumpyWriter:
    """Accumulates the (embeddings, text) batches of one partition in memory and writes the pair of files on flush()."""

    def __init__(self, partition_id, output_folder, output_partition_count):
        self._fs, root = fsspec.core.url_to_fs(output_folder)
        self._tag = _partition_tag(partition_id, output_partition_count)
        self._embed_path = f"{root}/embeddings/embeds_{self._tag}.npy"
        self._caption_path = f"{root}/captions/captions_{self._tag}.parquet"
        for sub in ("embeddings", "captions"):
            self._fs.makedirs(f"{root}/{sub}", exist_ok=True)
        self._chunks: List = []
        self._texts: List[str] = []

    def __call__(self, batch) -> None:
        self._chunks.append(batch["embeddings"])
        self._texts.extend(batch["text"])

    def flush(self) -> None:
        if not self._chunks or sum(c.shape[0] for c in self._chunks) == 0:
            return
        import numpy as np
        import pandas as pd
        raw = io.BytesIO()
        np.save(raw, np.concatenate(self._chunks))
        with self._fs.open(self._embed_path, "wb") as handle:
            handle.write(raw.getbuffer())
        with self._fs.open(self._caption_path, "wb") as handle:
            pd.DataFrame({"caption": self._texts}).to_parquet(handle)
        self._chunks, self._texts = [], []
