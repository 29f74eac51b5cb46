"""N > 1 data-parallel path on the CPU: world_size 2 over gloo (127.0.0.1).

Exercises the host logic of `a3t_b200.trainer.DataParallelTrainer` — flat parameter/gradient buffers,
round-robin sharding `batch[rank::world]` (abs_task.py:1503-1513), the single all-reduce with the
piggy-backed statistics, gradient = sum_r(loss_r * B_r) / sum_r B_r (trainer.py:583-595 + DDP mean),
clip + Adam + Noam — with the oracle ops standing in for the CUDA kernels (tests only)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from a3t_b200 import graph
from a3t_b200.model import build_model
from a3t_b200.trainer import DataParallelTrainer
from oracle import a3t_oracle as O
from oracle.oracle_backend import OracleBackend

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_tiny.pt")


def _model(fx):
    conf = fx["conf"]
    enc, dec = dict(conf["encoder_conf"]), dict(conf["decoder_conf"])
    for c in (enc, dec):
        c.update(dropout_rate=0.0, positional_dropout_rate=0.0, attention_dropout_rate=0.0)
    m = build_model(enc, dec, conf["model_conf"], vocab_size=fx["vocab"])
    m.load_state_dict(fx["state_dict"], strict=True)
    m.postnet.dropout_rate = 0.0
    return m.train()


def _shard(batch, rank, world):
    keys = ["speech", "text", "masked_position", "speech_mask", "text_mask", "speech_segment_pos", "text_segment_pos"]
    return {k: batch[k][rank::world].contiguous() for k in keys}


def _update(tr):
    """torch restatement of a3t_grad_sqnorm + a3t_adam_step (oracle clip_adam_step, Noam lr)."""
    g = tr.flat_g[:tr.n] / tr.stats[2]
    step = int(tr.step_count) + 1
    lr = O.noam_lr(tr.lr, tr.model_size, tr.warmup, step)
    O.clip_adam_step(tr.flat_p, g, tr.flat_m, tr.flat_v, step, lr, max_norm=tr.max_norm)
    tr.step_count += 1


def _worker(rank, world, port, out, bucket_bytes):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    fx = torch.load(GOLDEN, weights_only=False)
    m = _model(fx)
    tr = DataParallelTrainer(m, ops=OracleBackend(), update_fn=_update, bucket_bytes=bucket_bytes)
    stats = tr.step(_shard(fx["batch"], rank, world)).clone()
    torch.save({"p": tr.flat_p.clone(), "g": tr.flat_g.clone(), "stats": stats, "ranges": tr.exchange_ranges,
                "n": tr.n}, f"{out}.{rank}")
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
@pytest.mark.parametrize("bucket_bytes", [0, 24 << 20, 64 << 10])
def test_two_rank_step_matches_weighted_single_process(tmp_path, bucket_bytes):
    """bucket_bytes = 0 (default): one all-reduce after the backward sweep; 24 MiB: the overlapped code path, the tiny
    model fits one range; 64 KiB: the exchange is cut into several ranges issued during the sweep -- same result,
    ranges tile [0, n + 4) back to front."""
    world = 2
    out = str(tmp_path / "rank")
    mp.spawn(_worker, args=(world, _free_port(), out, bucket_bytes), nprocs=world, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    rg = r0["ranges"]
    assert rg == r1["ranges"] and rg[0][1] == r0["n"] + 4 and rg[-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(rg, rg[1:]))
    assert (len(rg) == 1) if (bucket_bytes == 0 or bucket_bytes >= (1 << 20)) else (len(rg) >= 4)
    # every rank holds the same reduced gradient, statistics and updated parameters
    assert torch.equal(r0["g"], r1["g"]) and torch.equal(r0["p"], r1["p"])
    # expected: per-shard gradients computed independently, combined as sum_r(g_r * B_r) / sum_r B_r
    fx = torch.load(GOLDEN, weights_only=False)
    m = _model(fx)
    names = [n for n, _ in m.named_parameters()]
    acc, wsum, lsum = None, 0.0, 0.0
    for rank in range(world):
        mm = _model(fx)
        P = {n: p.detach().clone() for n, p in mm.named_parameters()}
        P.update({n: b.clone() for n, b in mm.named_buffers()})
        sh = _shard(fx["batch"], rank, world)
        B = sh["speech"].shape[0]
        ops, wc = OracleBackend(), graph.WeightCache()
        loss, _, _, ctx = graph.forward(ops, P, wc, mm.cfg, sh, True, True)
        G = graph.backward(ops, P, wc, mm.cfg, ctx, torch.ones(1))
        flat = torch.cat([G[n].reshape(-1) for n in names]) * B
        acc = flat if acc is None else acc + flat
        wsum += B
        lsum += float(loss) * B
    n = acc.numel()
    err = float((r0["g"][:n] - acc).abs().max())
    assert err <= 2e-5 * float(acc.abs().max()) + 1e-6, err
    assert abs(float(r0["stats"][0]) - lsum) < 1e-4 * max(1.0, abs(lsum)) and float(r0["stats"][2]) == wsum
    # optimizer: same update as the oracle's clip+Adam+Noam on the combined gradient
    p0 = torch.cat([p.detach().reshape(-1) for _, p in m.named_parameters()])
    mbuf, vbuf = torch.zeros(n), torch.zeros(n)
    O.clip_adam_step(p0, acc / wsum, mbuf, vbuf, 1, O.noam_lr(1.0, m.encoder.attention_dim, 4000.0, 1))
    assert torch.allclose(r0["p"], p0, atol=2e-6)
