"""Run under torchrun with 2 ranks (one per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_joint_check.py
Checks SURVEY 8(e): (i) independent mode -- each rank's result equals a single-GPU run of that
rank's sub-batch; (ii) joint mode -- the gathered result of the two ranks equals ONE sampler call
on the whole batch (shared filter, global guidance norm)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from babe_b200 import distributed as bd, edm, sampler  # noqa: E402
from toy_model import ToyDenoiser  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    rank, local, world = bd.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.deterministic = True
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "fit_sampler.npz")))
    y2 = torch.from_numpy(g["y"])
    y = torch.cat((y2, 0.5 * y2.flip(0)), 0).to(dev)             # 4 rows
    B, T = y.shape
    args = sampler.make_args(sample_rate=int(g["sr"]), audio_len=T, T=4, NFFT=int(g["nfft"]), max_iter=10)
    model = ToyDenoiser().to(dev)

    def make(noise_rows):
        s = sampler.BlindSamplerFused(model, edm.EDM(args), args)
        gen = torch.Generator().manual_seed(5)
        # the reference draws the full-batch noise on the host; ranks slice their rows
        s.noise_fn = lambda shape, d: torch.randn((B, T), generator=gen)[noise_rows].to(d)
        return s

    lo, hi = bd.shard_rows(B, rank, world)
    rows = slice(lo, hi)
    # (ii) joint: 2 ranks x 2 rows == 1 call x 4 rows
    s = make(rows)
    s.joint = True
    xj, pj = s.predict_blind_bwe(y[rows].clone())
    xj_all = bd.gather_rows(xj)
    full = make(slice(0, B))
    xf, pf = full.predict_blind_bwe(y.clone())
    e_x, e_p = rel(xj_all, xf), rel(pj, pf)
    # (i) independent: rank result == single-GPU run of the same sub-batch
    s1 = make(rows)
    xi, pi = s1.predict_blind_bwe(y[rows].clone())
    s2 = make(rows)
    xi2, pi2 = s2.predict_blind_bwe(y[rows].clone())
    ok_ind = torch.equal(xi, xi2) and torch.equal(pi, pi2)
    ps = bd.gather_params(pi)
    print(f"rank {rank}: joint vs full-batch x {e_x:.2e} params {e_p:.2e}; independent reproducible {ok_ind}; "
          f"gathered params {tuple(ps.shape)}")
    assert e_x < 1e-4 and e_p < 1e-4 and ok_ind and ps.shape == (world, 2, 5)
    bd.barrier()
    if rank == 0:
        print("multi-GPU joint/independent check OK")
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
