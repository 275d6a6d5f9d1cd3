"""Data-parallel host logic on CPU with the gloo backend, world_size 2.

The sequence batch is sharded contiguously; every rank computes the gradient of ITS shard's mean NLL (here with the CPU
oracle standing in for the CUDA path, which needs a GPU), the flat gradient is SUM all-reduced and scaled by 1/world.
The result must equal the single-process gradient on the concatenated batch, and `bench.py --impl reference` under a
2-rank launch must print exactly one JSON line (rank 0) while rank 1 exits 0 without work."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import glow_oracle as O
from tests.helpers import ROOT, golden_params, load_golden, small_hparams


def _flat_grad(P, hy, batch):
    Pg = O.clone_params(P, requires_grad=True)
    _, _, loss = O.seq_forward(Pg, hy, batch)
    loss.backward()
    names = sorted(k for k, v in Pg.items() if v.grad is not None)
    return names, torch.cat([Pg[k].grad.reshape(-1) for k in names]), float(loss.detach())


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from lets_face_it_b200.train import allreduce_flat_gradient, shard_batch

    g = load_golden("kat_small_gru")
    hy = O.Hyper.from_hparams(small_hparams("gru"))
    P = golden_params(g)
    full = O.synthetic_batch(hy, 8, 12, seed=21)
    names, grad, loss = _flat_grad(P, hy, shard_batch(full, rank, world))
    allreduce_flat_gradient(grad, world)
    grad.mul_(1.0 / world)
    lt = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        torch.save({"names": names, "grad": grad, "loss": float(lt) / world}, os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_equals_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    g = load_golden("kat_small_gru")
    hy = O.Hyper.from_hparams(small_hparams("gru"))
    names, ref, loss = _flat_grad(golden_params(g), hy, O.synthetic_batch(hy, 8, 12, seed=21))
    assert names == got["names"]
    assert abs(loss - got["loss"]) < 1e-5 * abs(loss)
    err = float((got["grad"].double() - ref.double()).norm() / ref.double().norm())
    assert err < 1e-5, err


def test_shard_batch_rejects_ragged_split():
    from lets_face_it_b200.train import shard_batch

    with pytest.raises(ValueError):
        shard_batch({"p1_face": torch.zeros(5, 3, 2)}, 0, 2)
    s = shard_batch({"p1_face": torch.arange(8.0).view(4, 2, 1)}, 1, 2)
    assert s["p1_face"].shape[0] == 2 and float(s["p1_face"][0, 0, 0]) == 4.0


def test_flow_grad_bucket_is_one_contiguous_range():
    """The bucket reduced while the encoder backward still runs (train.py: flow_grad_range) must cover exactly the
    flow-step weight blocks of the flat layout and nothing whose gradient is finished later (ActNorm / 1x1 conv / encoders)."""
    from lets_face_it_b200.train import flow_grad_range

    blocks, off = {}, 0
    for name, n, k in [("an_bias", 56, 16), ("an_logs", 56, 16), ("inv_l", 3136, 16), ("inv_u", 3136, 16), ("inv_log_s", 56, 16),
                       ("wc", 512 * 1560, 16), ("bc", 512, 16), ("w_ih", 384 * 540, 16), ("b_ih", 384, 16), ("w_hh", 384 * 128, 16),
                       ("b_hh", 384, 16), ("wf", 56 * 128, 16), ("bf", 56, 16), ("lf", 56, 16), ("enc_w_ih.1", 768 * 56, 1),
                       ("enc_w_hh.1", 768 * 256, 1)]:
        blocks[name] = (off, n, k)
        off = (off + n * k + 63) // 64 * 64
    lo, hi = flow_grad_range(blocks)
    assert lo == blocks["wc"][0] and hi == blocks["lf"][0] + 56 * 16
    assert blocks["inv_log_s"][0] + 56 * 16 <= lo and hi <= blocks["enc_w_ih.1"][0]
    bad = dict(blocks)
    bad["enc_w_ih.1"] = (blocks["bc"][0] + 1, 8, 1)
    with pytest.raises(RuntimeError):
        flow_grad_range(bad)


def test_reference_arm_prints_once_under_two_ranks():
    port = 31500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--ref-batch", "2"]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0
