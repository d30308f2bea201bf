"""`python -m llamagen_b200.sample.vq_demo` — tokenizer round trip with the flags of
tokenizer/tokenizer_image/vq_demo.py:72-84: image -> centre crop -> VQModel.encode (:59) -> decode_code (:60) ->
bicubic to --image-size + uint8 (:63-64, one kernel here) -> file named <image>_<suffix>.<ext> in --output-dir."""
import argparse
import os

import numpy as np
import torch

from .. import VQ_models
from ..postprocess import to_uint8_nhwc


def center_crop(pil_image, image_size: int):
    """ADM centre crop (dataset/augmentation.py:8-26): halve with a box filter while the short side is at least twice
    the target, bicubic-resize the short side to the target, then cut the central square."""
    from PIL import Image
    w, h = pil_image.size
    while min(w, h) >= 2 * image_size:
        w, h = w // 2, h // 2
        pil_image = pil_image.resize((w, h), resample=Image.BOX)
    scale = image_size / min(w, h)
    pil_image = pil_image.resize((round(w * scale), round(h * scale)), resample=Image.BICUBIC)
    arr = np.array(pil_image)
    top, left = (arr.shape[0] - image_size) // 2, (arr.shape[1] - image_size) // 2
    return Image.fromarray(arr[top:top + image_size, left:left + image_size])


def load_weights(model, path):
    checkpoint = torch.load(path, map_location="cpu", weights_only=False)
    for key in ("ema", "model", "state_dict"):          # vq_demo.py:25-32
        if key in checkpoint:
            model.load_state_dict(checkpoint[key])
            return
    raise Exception("please check model weight")


def main(args):
    from PIL import Image
    torch.manual_seed(args.seed)
    torch.set_grad_enabled(False)
    model = VQ_models[args.vq_model](codebook_size=args.codebook_size, codebook_embed_dim=args.codebook_embed_dim)
    model.to("cuda")
    model.eval()
    if args.vq_ckpt:
        load_weights(model, args.vq_ckpt)
    else:
        print("WARNING: no --vq-ckpt given, using random-init tokenizer weights")
    os.makedirs(args.output_dir, exist_ok=True)
    stem, ext = os.path.splitext(os.path.basename(args.image_path))
    out_path = os.path.join(args.output_dir, f"{stem}_{args.suffix}{ext}" if ext in (".jpg", ".jpeg", ".png") else stem + ext)
    img = center_crop(Image.open(args.image_path).convert("RGB"), args.image_size)
    x = torch.from_numpy(2.0 * (np.array(img) / 255.0) - 1.0).permute(2, 0, 1)[None].float().to("cuda")
    latent, _, [_, _, indices] = model.encode(x)
    output = model.decode_code(indices, latent.shape)
    sample = to_uint8_nhwc(output, size=args.image_size)[0].cpu().numpy()
    Image.fromarray(sample).save(out_path)
    print("Reconstructed image is saved to {}".format(out_path))
    return out_path


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--image-path", type=str, default="assets/example.jpg")
    parser.add_argument("--output-dir", type=str, default="output_vq_demo")
    parser.add_argument("--suffix", type=str, default="tokenizer_image")
    parser.add_argument("--vq-model", type=str, choices=list(VQ_models.keys()), default="VQ-16")
    parser.add_argument("--vq-ckpt", type=str, default=None, help="ckpt path for vq model")
    parser.add_argument("--codebook-size", type=int, default=16384, help="codebook size for vector quantization")
    parser.add_argument("--codebook-embed-dim", type=int, default=8, help="codebook dimension for vector quantization")
    parser.add_argument("--image-size", type=int, choices=[256, 384, 448, 512, 1024], default=512)
    parser.add_argument("--seed", type=int, default=0)
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
