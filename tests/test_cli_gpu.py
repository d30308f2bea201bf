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


def test_serve_llm_continuous_batching_matches_independent_generate():
    """serve.LLM with iteration-level scheduling (llm_engine.py:511): requests JOIN MID-SEQUENCE of running ones (per-row positions in
    lg_decode_rows) and every request must receive exactly the tokens an INDEPENDENT generate() call with its seed produces.
    The reference point is generate() on max_num_seqs copies of the label (same row count -> same GEMM plan, row 0 draws with
    (seed, step, row 1) exactly like the request's RNG stream), so the comparison is bit-for-bit, sampling included."""
    import torch
    from llamagen_b200 import GPT_models, generate
    from llamagen_b200.serve import LLM, SamplingParams
    torch.manual_seed(0)
    gpt = GPT_models["GPT-B"](vocab_size=16384, block_size=64, num_classes=1000, cls_token_num=1, model_type="c2i")
    gpt = gpt.to("cuda", torch.bfloat16).eval()
    gpt.output.weight.data.normal_(std=0.02)
    S, slots = 64, 8
    sp = SamplingParams(temperature=1.0, top_p=1.0, top_k=2000, max_tokens=S)
    llm = LLM(gpt, cfg_scale=4.0, num_classes=1000, max_num_seqs=slots, seed=11)
    labels = [207, 360, 387, 974, 88, 979, 417, 279, 1, 2, 3]
    # three waves: 3 requests, 17 steps later 5 more (join at depth 17 of the first wave), 30 steps later the rest; 11 requests > 8 slots
    outs = []
    for c in labels[:3]:
        llm.add_request([c], sp)
    for _ in range(17):
        outs += llm.step()
    for c in labels[3:8]:
        llm.add_request([c], sp)
    for _ in range(30):
        outs += llm.step()
    for c in labels[8:]:
        llm.add_request([c], sp)
    while llm.has_unfinished_requests():
        outs += llm.step()
    # wave 1 ends at step 64; all 8 slots were busy, so wave 3 joins at step 65 (wave 2 is then at depth 47) and ends at step 128
    assert llm.steps_run == 2 * S
    got = {int(o.request_id): o.outputs[0].token_ids for o in outs}
    assert sorted(got) == list(range(len(labels))) and all(len(t) == S for t in got.values())
    for rid, c in enumerate(labels):
        ref = generate(gpt, torch.full((slots,), c, device="cuda"), S, cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, seed=11 + rid)
        assert ref[0].cpu().tolist() == got[rid], rid
    # LLM.generate surface (serve/sample_c2i.py:35-67): CFG twins get their conditional request's tokens, sorted by request id
    llm2 = LLM(gpt, cfg_scale=4.0, num_classes=1000, max_num_seqs=4, seed=5)
    prompts = [[c] for c in labels[:6]] + [[1000] for _ in range(6)]
    res = llm2.generate(prompt_token_ids=prompts, sampling_params=sp)
    assert [o.request_id for o in res] == [str(i) for i in range(12)]
    toks = torch.tensor([o.outputs[0].token_ids for o in res])
    assert torch.equal(toks[:6], toks[6:])
    with pytest.raises(ValueError):
        llm2.generate(prompt_token_ids=[[1], [2]], sampling_params=sp)        # cfg on but no null-class twins
    with pytest.raises(ValueError):
        LLM(gpt, cfg_scale=1.0).generate(prompt_token_ids=[[1, 2]])


def test_t2i_feature_files_to_tokens_match_oracle(tmp_path):
    """SURVEY §8 f-3, value check: T5 feature FILES (the fp32 [1, valid_len, dim] .npy layout of language/extract_t5_feature.py:103-108)
    -> cond.load_t5_feature_files -> cond.prepare_condition (batched left-padding) -> generate() must give the greedy tokens
    the ORACLE produces from the reference's own per-prompt front end (sample_t2i.py:92-106, restated inline below).
    fp32 exact mode on the reference-made golden t2i model: token ids bit-exact, logits <= 1e-4."""
    import torch
    from llamagen_b200 import generate
    from llamagen_b200.cond import load_t5_feature_files, prepare_condition
    from oracle import GPTOracle
    from util import build_gpt, load_golden
    g = load_golden("gpt_t2i.pt")
    T, C, S = g["cfg"]["cls_token_num"], g["cfg"]["caption_dim"], g["S"]
    rng = np.random.default_rng(7)
    lens = [5, 37, 120, 64]
    paths = []
    for i, n in enumerate(lens):
        np.save(tmp_path / f"{i}.npy", rng.standard_normal((1, n, C)).astype(np.float32))
        paths.append(str(tmp_path / f"{i}.npy"))
    # --- reference front end (sample_t2i.py:89-106): right-padded features + mask as T5Embedder returns them, then the per-prompt loop
    embs = torch.zeros(len(lens), T, C)
    masks = torch.zeros(len(lens), T)
    for i, n in enumerate(lens):
        embs[i, :n] = torch.from_numpy(np.load(paths[i]))[0]
        masks[i, :n] = 1
    new_masks = torch.flip(masks, dims=[-1])
    new_embs = torch.stack([torch.cat([e[int(m.sum().item()):], e[:int(m.sum().item())]]) for e, m in zip(embs, masks)])
    ref_cond, ref_masks = new_embs * new_masks[:, :, None], new_masks
    ref_t, ref_l = GPTOracle(g["state_dict"], g["cfg"]).generate(ref_cond, S, emb_masks=ref_masks, cfg_scale=4.0, sample_logits=False)
    # --- product front end + engine
    m = build_gpt(g["cfg"], g["state_dict"], torch.float32)
    e2, m2 = load_t5_feature_files(paths, T, C)
    cond, cmask = prepare_condition(e2.cuda(), m2.cuda(), left_padding=True)
    assert torch.equal(cond.cpu(), ref_cond) and torch.equal(cmask.cpu(), ref_masks)
    toks, logits = generate(m, cond, S, emb_masks=cmask, cfg_scale=4.0, sample_logits=False, return_logits=True)
    assert (logits.cpu() - ref_l).abs().max().item() <= 1e-4
    assert torch.equal(toks.cpu(), ref_t)
