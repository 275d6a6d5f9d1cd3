"""Worker of tests/test_gpu_dp2.py: one rank per GPU (torchrun), NCCL.  Every rank takes its contiguous shard of a seeded
batch, runs one `Trainer.step` (lr = 0: parameters stay put) with the overlapped bucketed all-reduce, and rank 0 compares
the reduced flat gradient x 1/world with the gradient of a single-process step on the whole batch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lets_face_it_b200 import _cabi as cabi  # noqa: E402
from lets_face_it_b200.train import Trainer, shard_batch  # noqa: E402
from tests.helpers import final_hparams  # noqa: E402
from tests.kat import build_kat_model, kat_batch, to_device  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = "cuda:%d" % local
    hp = final_hparams()
    m = build_kat_model(hp, dev)
    m.glow.set_actnorm_init(True)
    m.gemm_mode = cabi.GEMM_BF16X3
    m.train()
    full = kat_batch(hp, 128 * world, 80, seed=77)
    tr = Trainer(m, dropout=False, lr=0.0)
    assert tr.overlap and tr.world == world
    tr.step(to_device(shard_batch(full, rank, world), dev))
    torch.cuda.synchronize()
    got = tr.gflat[:tr.eng.n_theta].clone() / world
    ok = True
    if rank == 0:
        dist_off = Trainer(m, dropout=False, lr=0.0)
        dist_off.world, dist_off.overlap = 1, False      # single-process reference on the whole batch
        dist_off.step(to_device(full, dev))
        torch.cuda.synchronize()
        ref = dist_off.gflat[:tr.eng.n_theta]
        worst = 0.0
        for name, (off, n, k) in tr.eng.blocks.items():
            a, b = got[off:off + n * k].double(), ref[off:off + n * k].double()
            err = float((a - b).norm() / b.norm().clamp_min(1e-30))
            worst = max(worst, err)
            if err > 2e-4:
                ok = False
                print("MISMATCH", name, err)
        print("dp2 worst per-block relative L2 error %.3e" % worst)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
