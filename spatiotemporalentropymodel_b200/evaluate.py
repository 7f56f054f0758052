"""Sequence / GOP evaluation driver: the loop of stem/evalSTEM.py:156-231 (evalDataset) over this package's models.

What the reference does per sequence (``evalSTEM.py:186-209``): frame ``index % GOP == 1`` is an I-frame coded with
the image model's real entropy coder (``inferenceI_DVR``, :34-89), every other frame is a P-frame whose latent is
conditioned on the previous *decoded* latent (``inferenceP_DVR``, :92-154); it reports the mean of the per-frame PSNRs
and bpps.  Here the unit of work is one (sequence, GOP) pair: GOPs are independent (``y_conditioned`` is reset by every
I-frame), so they shard round-robin over the ranks of a ``torch.distributed`` job with no data-path collective; the
per-frame (bpp, psnr) rows are combined by one ``all_reduce`` at the end.

Two P-frame modes:
  * ``"estimate"`` - rate from the likelihoods (what ``evalSTEM`` logs as ``estimate_bpp``; the BASELINE metric).  For
    the variants with a spatial context model the whole GOP runs as one batch through ``PFramePipeline``.
  * ``"real"``     - ``compress`` -> strings -> ``decompress`` per frame, bpp from the string lengths, exactly the
    ``inferenceP_DVR`` sequence.
MS-SSIM (``pytorch_msssim``) is not computed (out of scope).
"""
from __future__ import annotations

import argparse
import math
import os
import sys
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .dist import shard_units

Tensor = torch.Tensor


def pad_to_64(x: Tensor) -> Tuple[Tensor, Tuple[int, int, int, int]]:
    """evalSTEM.py:40-54 / :96-109: centred zero padding to a multiple of 64 -> (padded, (left, right, top, bottom))."""
    h, w = x.size(2), x.size(3)
    nh, nw = (h + 63) // 64 * 64, (w + 63) // 64 * 64
    left, top = (nw - w) // 2, (nh - h) // 2
    pad = (left, nw - w - left, top, nh - h - top)
    return F.pad(x, pad, mode="constant", value=0), pad


def crop(x: Tensor, pad: Tuple[int, int, int, int]) -> Tensor:
    left, right, top, bottom = pad
    return x[:, :, top:x.size(2) - bottom, left:x.size(3) - right]


def psnr(a: Tensor, b: Tensor) -> float:
    """evalSTEM.py:29-31"""
    mse = float(F.mse_loss(a, b))
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)


def gop_units(n_frames: Sequence[int], gop: int) -> List[Tuple[int, int, int]]:
    """(sequence, first frame, number of frames) of every GOP, in the reference's frame order."""
    units = []
    for s, n in enumerate(n_frames):
        for f0 in range(0, n, gop):
            units.append((s, f0, min(gop, n - f0)))
    return units


def code_iframe(net, x: Tensor) -> Dict[str, object]:
    """inferenceI_DVR (evalSTEM.py:34-89): real entropy coding of one frame (1, 3, H, W)."""
    xp, pad = pad_to_64(x)
    enc = net.compress(xp)
    dec = net.decompress(enc["strings"], enc["shape"])
    x_hat = crop(dec["x_hat"], pad)
    n_pix = x.size(0) * x.size(2) * x.size(3)
    return {"y_conditioned": dec["y_hat"], "psnr": psnr(x, x_hat),
            "bpp": sum(len(s[0]) for s in enc["strings"]) * 8.0 / n_pix}


def code_pframe_real(net, stem, x: Tensor, y_conditioned: Tensor) -> Dict[str, object]:
    """inferenceP_DVR (evalSTEM.py:92-154) with the real entropy coder."""
    xp, pad = pad_to_64(x)
    y_cur, _ = net.getY(xp)
    enc = stem.compress(y_cur, y_conditioned)
    dec = stem.decompress(enc["strings"], enc["shape"], y_conditioned)
    y_hat = dec["y_hat"] if isinstance(dec, dict) else dec
    x_hat = crop(net.getX(y_hat), pad)
    n_pix = x.size(0) * x.size(2) * x.size(3)
    return {"y_conditioned": y_hat, "psnr": psnr(x, x_hat),
            "bpp": sum(len(s[0]) for s in enc["strings"]) * 8.0 / n_pix}


def as_float(frames: Tensor) -> Tensor:
    """8-bit frames -> what torchvision's ToTensor hands the reference (evalSTEM.py:185): float(v) / 255."""
    return frames.to(torch.float32).div(255.0) if frames.dtype == torch.uint8 else frames


def code_gop(net, stem, frames: Tensor, mode: str = "estimate", all_intra: bool = False) -> List[Tuple[float, float]]:
    """One GOP (T, 3, H, W) on the models' device -> [(bpp, psnr)] per frame (I-frame first). frames: fp32 in [0, 1]
    or uint8 (8-bit samples): the batched P-frame pipeline takes the bytes as they are (a quarter of the upload) and
    evaluates v / 255 on the device, bit-identical to converting first."""
    if mode not in ("estimate", "real"):
        raise ValueError('mode must be "estimate" or "real"')
    T = frames.size(0)
    if all_intra or mode == "real":
        frames = as_float(frames)
    if all_intra:
        return [(o["bpp"], o["psnr"]) for o in (code_iframe(net, frames[t:t + 1]) for t in range(T))]
    out_i = code_iframe(net, as_float(frames[0:1]))
    rows = [(out_i["bpp"], out_i["psnr"])]
    y_cond = out_i["y_conditioned"]
    if T == 1:
        return rows
    if mode == "real":
        for t in range(1, T):
            o = code_pframe_real(net, stem, frames[t:t + 1], y_cond)
            y_cond = o["y_conditioned"]
            rows.append((o["bpp"], o["psnr"]))
        return rows
    from .models import make_pipeline
    H, W = frames.size(2), frames.size(3)
    # one pipeline per (I-frame model, STEM model) pair: GOPs of a repeating shape replay its captured CUDA graph
    # (keyed by the engines, which the models rebuild on load_state_dict / .to(); the pipeline keeps them alive)
    pipes = stem.__dict__.setdefault("_stemb200_pipelines", {})
    key = (id(net.engine()), id(stem.engine()))
    if key not in pipes:
        pipes.clear()
        pipes[key] = make_pipeline(net, stem)
    pipe = pipes[key]
    stats = pipe.run_gop(frames[1:].contiguous(), y_cond, want_outputs=False)["stats"].cpu()
    for t in range(T - 1):
        mse = float(stats[2, t]) / (3 * H * W)
        rows.append((float(stats[0, t] + stats[1, t]) / (H * W), float("inf") if mse == 0 else -10 * math.log10(mse)))
    return rows


def eval_dataset(sequences: Sequence[Tuple[str, Callable[[int, int], Tensor], int]], gop: int,
                 gop_fn: Callable[[Tensor], List[Tuple[float, float]]], rank: int = 0, world: int = 1,
                 device: Optional[torch.device] = None) -> Dict[str, object]:
    """evalDataset (evalSTEM.py:156-231).  sequences: (name, load(first, count) -> (count, 3, H, W) fp32 in [0, 1],
    n_frames); gop_fn codes one GOP and returns its per-frame (bpp, psnr) rows.  GOP u runs on rank u % world; every
    rank returns the same aggregate: PSNR_AVE / BPP_AVE = means over all frames, plus the per-frame table."""
    import torch.distributed as dist
    counts = [n for _, _, n in sequences]
    units = gop_units(counts, gop)
    offsets = [0]
    for n in counts:
        offsets.append(offsets[-1] + n)
    table = torch.zeros((offsets[-1], 2), dtype=torch.float64)
    for u in shard_units(len(units), rank, world):
        s, f0, n = units[u]
        frames = sequences[s][1](f0, n)
        if device is not None:
            frames = frames.to(device)
        rows = gop_fn(frames)
        if len(rows) != n:
            raise RuntimeError(f"gop_fn returned {len(rows)} rows for a GOP of {n} frames")
        table[offsets[s] + f0:offsets[s] + f0 + n] = torch.tensor(rows, dtype=torch.float64)
    if world > 1:
        t = table.to(device) if (device is not None and dist.get_backend() == "nccl") else table
        dist.all_reduce(t, op=dist.ReduceOp.SUM)   # every frame was written by exactly one rank
        table = t.cpu()
    per_seq = {name: {"bpp": float(table[offsets[i]:offsets[i + 1], 0].mean()),
                      "psnr": float(table[offsets[i]:offsets[i + 1], 1].mean())}
               for i, (name, _, _) in enumerate(sequences)}
    return {"BPP_AVE": float(table[:, 0].mean()), "PSNR_AVE": float(table[:, 1].mean()), "frames": table,
            "sequences": per_seq, "n_gops": len(units)}


def png_sequence(path: str, max_index: int) -> Tuple[str, Callable[[int, int], Tensor], int]:
    """f001.png ... as evalSTEM.py:186 reads them (PIL RGB); returned as uint8 CHW, ToTensor's scaling happens on
    the device."""
    import numpy as np
    from PIL import Image

    def load(first: int, count: int) -> Tensor:
        out = []
        for i in range(first + 1, first + count + 1):
            img = np.asarray(Image.open(os.path.join(path, f"f{i:03d}.png")).convert("RGB"), dtype=np.uint8)
            out.append(torch.from_numpy(img.copy()).permute(2, 0, 1))
        return torch.stack(out).contiguous()   # uint8: code_gop / the CUDA pipeline apply ToTensor's v / 255

    return os.path.basename(os.path.normpath(path)), load, max_index


def synthetic_sequence(name: str, n_frames: int, height: int, width: int, seed: int):
    from . import synthetic as S
    frames = S.make_frames(n_frames, height, width, seed=seed)
    return name, (lambda first, count: frames[first:first + count]), n_frames


def main(argv=None) -> int:
    """python -m spatiotemporalentropymodel_b200.evaluate  (torchrun for several GPUs)."""
    ap = argparse.ArgumentParser(description="GOP evaluation of the STEM P-frame codec on B200 (evalSTEM.py driver)")
    ap.add_argument("--variant", default="SpatioTemporalPriorModel_Res")
    ap.add_argument("--checkpoint", help="I-frame model checkpoint ({'state_dict': ...}, evalSTEM.py:300-308)")
    ap.add_argument("--entropy-model-path", help="STEM checkpoint (evalSTEM.py:309-317)")
    ap.add_argument("--dataset-dir", help="directory with one sub-directory of f001.png ... per sequence")
    ap.add_argument("--frames", type=int, default=36, help="frames per sequence (evalSTEM.py:181-184: 36 / 30)")
    ap.add_argument("-gop", "--gop", type=int, default=12)
    ap.add_argument("--mode", default="estimate", choices=["estimate", "real"])
    ap.add_argument("--all-intra", action="store_true")
    ap.add_argument("--synthetic", type=int, default=0, help="use N seeded synthetic 1080p sequences instead of PNGs")
    ap.add_argument("--size", default="1080x1920")
    args = ap.parse_args(argv)
    import torch.distributed as dist
    from . import models as M, synthetic as S
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = M.models["mbt2018"](quality=4)
    net.load_state_dict(torch.load(args.checkpoint, map_location="cpu")["state_dict"] if args.checkpoint
                        else S.make_iframe_state_dict(0))
    stem = getattr(M, args.variant)()
    stem.load_state_dict(torch.load(args.entropy_model_path, map_location="cpu")["state_dict"]
                         if args.entropy_model_path else S.make_stem_state_dict(args.variant, 0))
    net.update(force=True)
    stem.update(force=True)
    net, stem = net.to(dev).eval(), stem.to(dev).eval()
    if args.synthetic:
        h, w = (int(v) for v in args.size.split("x"))
        seqs = [synthetic_sequence(f"synthetic{i}", args.frames, h, w, 100 + i) for i in range(args.synthetic)]
    elif args.dataset_dir:
        seqs = [png_sequence(os.path.join(args.dataset_dir, d), args.frames)
                for d in sorted(os.listdir(args.dataset_dir)) if os.path.isdir(os.path.join(args.dataset_dir, d))]
    else:
        ap.error("give --dataset-dir or --synthetic N")
    with torch.no_grad():
        res = eval_dataset(seqs, args.gop, lambda fr: code_gop(net, stem, fr, args.mode, args.all_intra), rank, world, dev)
    if rank == 0:
        print(f"PSNR_AVE: {res['PSNR_AVE']:.3f}  BPP_AVE: {res['BPP_AVE']:.4f}  ({res['n_gops']} GOPs on {world} GPU(s))")
        for name, v in res["sequences"].items():
            print(f"  {name}: PSNR {v['psnr']:.3f}  bpp {v['bpp']:.4f}")
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
