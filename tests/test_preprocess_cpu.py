"""The preprocess output format (clipcap/preprocess/writer.py) written by clipcap_b200.preprocess and, where the
reference is present, read back against files the reference's own writer produces from the same samples."""
import numpy as np
import pandas as pd
import pytest
import yaml

from clipcap_b200.encoders.config import EncoderConfig
from clipcap_b200.preprocess import NumpyWriter, save_config
from oracle import ref_runner as RR


def _samples():
    rng = np.random.default_rng(0)
    return [{"embeddings": rng.standard_normal((n, 8)).astype(np.float16), "text": [f"caption {i}-{j}" for j in range(n)]}
            for i, n in enumerate((3, 1, 4))]


def test_numpy_writer_layout(tmp_path):
    out = str(tmp_path / "ds")
    save_config(EncoderConfig(encoder_embedding_size=8), out)
    w = NumpyWriter(partition_id=7, output_folder=out, output_partition_count=120)
    for s in _samples():
        w(s)
    w.flush()
    w.flush()  # idempotent on an empty buffer
    emb = np.load(tmp_path / "ds" / "embeddings" / "embeds_007.npy")
    cap = pd.read_parquet(tmp_path / "ds" / "captions" / "captions_007.parquet")
    assert emb.shape == (8, 8) and emb.dtype == np.float16
    assert np.array_equal(emb, np.concatenate([s["embeddings"] for s in _samples()]))
    assert list(cap.columns) == ["caption"] and cap["caption"].tolist() == sum((s["text"] for s in _samples()), [])
    cfg = yaml.safe_load(open(tmp_path / "ds" / "encoder_config.yaml"))
    assert cfg["encoder_embedding_size"] == 8 and cfg["encoder_model_variant"] == "ViT-L/14"


@pytest.mark.skipif(not RR.available(), reason="/root/reference not present (GPU box)")
def test_numpy_writer_matches_reference_writer(tmp_path):
    RR.import_reference()
    from clipcap.preprocess.writer import NumpyWriter as RefWriter
    a, b = str(tmp_path / "ours"), str(tmp_path / "ref")
    for cls, out in ((NumpyWriter, a), (RefWriter, b)):
        w = cls(partition_id=3, output_folder=out, output_partition_count=10)
        for s in _samples():
            w(s)
        w.flush()
    for rel in ("embeddings/embeds_03.npy",):
        assert (tmp_path / "ours" / rel).read_bytes() == (tmp_path / "ref" / rel).read_bytes()
    ours = pd.read_parquet(tmp_path / "ours" / "captions" / "captions_03.parquet")
    ref = pd.read_parquet(tmp_path / "ref" / "captions" / "captions_03.parquet")
    assert ours.equals(ref)
