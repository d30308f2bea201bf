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


def test_sample_c2i_ddp_cli_resizes_to_eval_size(tmp_path, monkeypatch):
    """--image-size 384 --image-size-eval 256: the bicubic resize of sample_c2i_ddp.py:141-142 runs inside lg_pixels_to_u8."""
    from llamagen_b200.sample import sample_c2i_ddp
    monkeypatch.chdir(tmp_path)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = sample_c2i_ddp.build_parser().parse_args(["--gpt-model", "GPT-B", "--image-size", "384", "--image-size-eval", "256",
                                                     "--num-fid-samples", "4", "--per-proc-batch-size", "2", "--sample-dir", "s"])
    sample_c2i_ddp.main(args)
    npz = [f for f in os.listdir(tmp_path / "s") if f.endswith(".npz")]
    arr = np.load(tmp_path / "s" / npz[0])["arr_0"]
    assert arr.shape == (4, 256, 256, 3) and arr.std() > 1.0


def test_vq_demo_cli_round_trip(tmp_path, monkeypatch):
    """tokenizer/tokenizer_image/vq_demo.py: image -> encode -> decode_code -> <name>_<suffix>.png."""
    from PIL import Image
    from llamagen_b200.sample import vq_demo
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(0)
    Image.fromarray(rng.integers(0, 255, (300, 420, 3), dtype=np.uint8)).save(tmp_path / "in.png")
    args = vq_demo.build_parser().parse_args(["--image-path", str(tmp_path / "in.png"), "--image-size", "256", "--output-dir", "o"])
    out = vq_demo.main(args)
    assert out.endswith("in_tokenizer_image.png")
    assert Image.open(out).size == (256, 256)


@pytest.mark.parametrize("mode", ["synthetic", "feature-dir"])
def test_sample_t2i_ddp_cli_single_rank(tmp_path, monkeypatch, mode):
    """sample_t2i_ddp.py with a 5-row prompt TSV: 6 images (3 per batch, last index past the list), jsonl + captions."""
    import json
    from llamagen_b200.sample import sample_t2i_ddp
    monkeypatch.chdir(tmp_path)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    (tmp_path / "parti.tsv").write_text("Prompt\tCategory\n" + "".join(f"a photo of thing {i}\tx\n" for i in range(5)))
    extra = ["--synthetic-cond"]
    if mode == "feature-dir":
        os.makedirs(tmp_path / "feat")
        rng = np.random.default_rng(0)
        for i in range(5):
            np.save(tmp_path / "feat" / f"{i}.npy", rng.standard_normal((1, 5 + 3 * i, 2048)).astype(np.float32))
        extra = ["--t5-feature-dir", "feat"]
    args = sample_t2i_ddp.build_parser().parse_args(["--gpt-model", "GPT-B", "--image-size", "256", "--prompt-csv", "parti.tsv",
                                                     "--per-proc-batch-size", "3", "--sample-dir", "s"] + extra)
    folder = sample_t2i_ddp.main(args)
    pngs = sorted(os.listdir(os.path.join(folder, "images")))
    assert pngs == [f"{i:06d}.png" for i in range(6)]
    from PIL import Image
    assert Image.open(os.path.join(folder, "images", pngs[0])).size == (256, 256)
    rows = [json.loads(l) for l in open(os.path.join(folder, "result.jsonl"))]
    assert len(rows) == 5 and rows[2]["text"] == "a photo of thing 2" and rows[2]["image_path"].endswith("000002.png")
    assert open(os.path.join(folder, "captions.txt")).read().splitlines()[4] == "a photo of thing 4"
