"""GPU: the CLI drop-ins run end to end with the reference's flags (random-init weights, no checkpoints offline)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sample_c2i_cli(tmp_path, monkeypatch):
    from llamagen_b200.sample import sample_c2i
    monkeypatch.chdir(tmp_path)
    args = sample_c2i.build_parser().parse_args(["--gpt-model", "GPT-B", "--image-size", "256", "--cfg-scale", "4.0",
                                                 "--top-k", "2000", "--seed", "1"])
    sample_c2i.main(args)
    from PIL import Image
    img = Image.open(tmp_path / "sample_c2i.png")
    assert img.size == (4 * 256 + 5 * 2, 2 * 256 + 3 * 2)      # torchvision grid: 8 images, nrow=4, padding 2


def test_sample_t2i_cli_synthetic_features(tmp_path, monkeypatch):
    from llamagen_b200.sample import sample_t2i
    monkeypatch.chdir(tmp_path)
    args = sample_t2i.build_parser().parse_args(["--gpt-model", "GPT-B", "--image-size", "256", "--synthetic-cond"])
    sample_t2i.main(args)
    assert (tmp_path / "sample_t2i.png").stat().st_size > 10000


def test_sample_c2i_ddp_cli_single_rank(tmp_path, monkeypatch):
    from llamagen_b200.sample import sample_c2i_ddp
    monkeypatch.chdir(tmp_path)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = sample_c2i_ddp.build_parser().parse_args(["--gpt-model", "GPT-B", "--image-size", "256", "--image-size-eval", "256",
                                                     "--num-fid-samples", "8", "--per-proc-batch-size", "4", "--sample-dir", "s"])
    sample_c2i_ddp.main(args)
    npz = [f for f in os.listdir(tmp_path / "s") if f.endswith(".npz")]
    assert len(npz) == 1
    arr = np.load(tmp_path / "s" / npz[0])["arr_0"]
    assert arr.shape == (8, 256, 256, 3) and arr.dtype == np.uint8
