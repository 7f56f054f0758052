"""Evaluation driver (stem/evalSTEM.py:156-231 restated in spatiotemporalentropymodel_b200.evaluate).
CPU: GOP partition, padding, and the world_size-2 gloo run of eval_dataset against the single-process result.
GPU: both P-frame modes on tiny synthetic sequences against the oracle's evaluation of the same loop."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spatiotemporalentropymodel_b200 import evaluate as EV
from spatiotemporalentropymodel_b200 import synthetic as S


def _seqs():
    return [EV.synthetic_sequence("a", 7, 40, 56, seed=1), EV.synthetic_sequence("b", 5, 40, 56, seed=2),
            EV.synthetic_sequence("c", 3, 40, 56, seed=3)]


def _stub_gop(frames):
    """deterministic stand-in for the codec: rows depend on the frame content and on the position inside the GOP"""
    return [(float(frames[t].mean()) + (0.5 if t == 0 else 0.0), 30.0 + float(frames[t].std()) * 10 + t)
            for t in range(frames.size(0))]


def test_gop_units_and_padding():
    assert EV.gop_units([7, 5, 3], 3) == [(0, 0, 3), (0, 3, 3), (0, 6, 1), (1, 0, 3), (1, 3, 2), (2, 0, 3)]
    x = torch.rand(1, 3, 1080, 1920)
    xp, pad = EV.pad_to_64(x)
    assert xp.shape == (1, 3, 1088, 1920) and pad == (0, 0, 4, 4)          # evalSTEM.py:96-109: centred
    assert torch.equal(EV.crop(xp, pad), x)
    assert float(xp[:, :, :4].abs().max()) == 0.0
    assert EV.psnr(x, x) == float("inf")
    name, load, n = _seqs()[0]
    assert load(2, 3).shape == (3, 3, 40, 56) and n == 7


def test_eval_dataset_single_process():
    res = EV.eval_dataset(_seqs(), 3, _stub_gop)
    assert res["n_gops"] == 6 and res["frames"].shape == (15, 2)
    rows = []
    for _, load, n in _seqs():
        for f0 in range(0, n, 3):
            rows += _stub_gop(load(f0, min(3, n - f0)))
    assert torch.allclose(res["frames"], torch.tensor(rows, dtype=torch.float64))
    assert math.isclose(res["BPP_AVE"], sum(r[0] for r in rows) / 15) and set(res["sequences"]) == {"a", "b", "c"}
    with pytest.raises(RuntimeError):
        EV.eval_dataset(_seqs(), 3, lambda fr: [(0.0, 0.0)])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = EV.eval_dataset(_seqs(), 3, _stub_gop, rank, world)
        q.put((rank, res["frames"], res["PSNR_AVE"], res["BPP_AVE"]))
    finally:
        dist.destroy_process_group()


def test_eval_dataset_two_ranks_match_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = EV.eval_dataset(_seqs(), 3, _stub_gop)
    for _, frames, ps, bp in results:
        assert torch.equal(frames, ref["frames"]) and ps == ref["PSNR_AVE"] and bp == ref["BPP_AVE"]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel", "SpatioTemporalPriorModelWithoutSPM"])
def test_eval_modes_against_oracle_gpu(variant):
    from oracle import stem_oracle as O
    from spatiotemporalentropymodel_b200 import models as M
    dev = torch.device("cuda:0")
    sd_i, sd_s = S.make_iframe_state_dict(0), S.make_stem_state_dict(variant, 0)
    net = M.models["mbt2018"](quality=4)
    net.load_state_dict(sd_i)
    net.update(force=True)
    stem = getattr(M, variant)()
    stem.load_state_dict(sd_s)
    stem.update(force=True)
    net, stem = net.to(dev).eval(), stem.to(dev).eval()
    seqs = [EV.synthetic_sequence("s0", 5, 128, 128, seed=41), EV.synthetic_sequence("s1", 3, 128, 128, seed=42)]
    est = EV.eval_dataset(seqs, 3, lambda fr: EV.code_gop(net, stem, fr, "estimate"), device=dev)
    real = EV.eval_dataset(seqs, 3, lambda fr: EV.code_gop(net, stem, fr, "real"), device=dev)
    intra = EV.eval_dataset(seqs, 3, lambda fr: EV.code_gop(net, stem, fr, "real", all_intra=True), device=dev)
    assert est["n_gops"] == 3 and est["frames"].shape == (8, 2)
    # I-frames (rows 0, 3, 5) are the same real codec in every mode
    for r in (0, 3, 5):
        assert torch.equal(est["frames"][r], real["frames"][r]) and torch.equal(est["frames"][r], intra["frames"][r])
    assert torch.isfinite(est["frames"]).all() and torch.isfinite(real["frames"]).all()
    # oracle evaluation of the first GOP: reference I-frame codec (raster AR scan) + reference P-frame forward
    name, load, _ = seqs[0]
    fr = load(0, 3)
    with torch.no_grad():
        i_ref = O.iframe_ar_code(fr[0:1], sd_i)
        ref = O.gop_forward(fr[1:3], i_ref["y_hat"], sd_i, sd_s, variant)
    i_psnr = -10 * math.log10(float(((i_ref["x_hat"] - fr[0:1]) ** 2).mean()))
    assert abs(float(est["frames"][0, 1]) - i_psnr) < 0.02
    for t, r in enumerate(ref):
        got_bpp, got_psnr = float(est["frames"][1 + t, 0]), float(est["frames"][1 + t, 1])
        # y_conditioned comes from two different AR decoders (a few symbols differ): 1 % / 0.05 dB
        assert abs(got_bpp - float(r["bpp"])) / float(r["bpp"]) < 1e-2, (t, got_bpp, float(r["bpp"]))
        assert abs(got_psnr - float(r["psnr"])) < 0.05, (t, got_psnr, float(r["psnr"]))
