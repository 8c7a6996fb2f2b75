"""The reference's preprocess output format (clipcap/preprocess/writer.py:10-100): `encoder_config.yaml`,
`embeddings/embeds_<id>.npy` ([N, E], dtype as produced by the encoder) and `captions/captions_<id>.parquet` (one
`caption` column), `<id>` zero-padded to the width of the partition count. Paths go through fsspec like the reference."""
from __future__ import annotations

import math
from io import BytesIO

import fsspec
import yaml

from clipcap_b200.encoders.config import EncoderConfig


def save_config(config: EncoderConfig, output_folder: str) -> None:
    fs, output_folder = fsspec.core.url_to_fs(output_folder)
    fs.makedirs(output_folder, exist_ok=True)
    with fs.open(output_folder + "/encoder_config.yaml", "w") as f:
        yaml.dump(config.to_dict(), f, default_flow_style=False)


class OutputSink:
    def __init__(self, output_folder, partition_id, output_partition_count):
        self.fs, output_folder = fsspec.core.url_to_fs(output_folder)
        self.output_folder = output_folder
        self.embed_folder = output_folder + "/embeddings"
        self.captions_folder = output_folder + "/captions"
        self.batch_num = partition_id
        self.oom_partition_count = int(math.log10(output_partition_count)) + 1
        self.fs.makedirs(self.embed_folder, exist_ok=True)
        self.fs.makedirs(self.captions_folder, exist_ok=True)
        self._reset()

    def _reset(self):
        self.embeddings, self.captions, self.batch_count = [], [], 0

    def add(self, sample):
        self.batch_count += sample["embeddings"].shape[0]
        self.embeddings.append(sample["embeddings"])
        self.captions.extend(sample["text"])

    def flush(self):
        if self.batch_count == 0:
            return
        import numpy as np
        import pandas as pd
        batch_num_str = str(self.batch_num).zfill(self.oom_partition_count)
        with self.fs.open(self.embed_folder + "/embeds_" + batch_num_str + ".npy", "wb") as f:
            buf = BytesIO()
            np.save(buf, np.concatenate(self.embeddings))
            f.write(buf.getbuffer())
        df = pd.DataFrame(data=list(zip(self.captions)), columns=["caption"])
        with self.fs.open(self.captions_folder + "/captions_" + batch_num_str + ".parquet", "wb") as f:
            df.to_parquet(f)
        self._reset()


class NumpyWriter:
    def __init__(self, partition_id, output_folder, output_partition_count):
        self.sink = OutputSink(output_folder, partition_id, output_partition_count)

    def __call__(self, batch):
        self.sink.add(batch)

    def flush(self):
        self.sink.flush()
