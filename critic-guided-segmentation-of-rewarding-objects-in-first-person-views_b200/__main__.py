"""`python -m cgs_b200 -train|-process|-eval ...`: the sequencing of reference main.py:1540-1566 on the B200 classes.

Same flags as the reference for the hot path (`train_handler.build_parser`).  Data: the reference collects MineRL episodes
(main.py:1272-1359; needs the MineRL download, out of scope); here `--synthetic N` generates the seeded stand-in of
SURVEY.md §8d, or `--data file.npz` loads arrays `X` uint8 [N,64,64,3] and `Y` float [7,N]."""
import sys

import numpy as np
import torch

from . import ops
from .train_handler import Handler, build_parser
from . import synth


def main(argv=None):
    p = build_parser()
    p.add_argument("--synthetic", type=int, default=0, help="train on N seeded synthetic frames (SURVEY.md §8d)")
    p.add_argument("--data", type=str, default="", help=".npz with X uint8 [N,64,64,3] and Y float [7,N]")
    p.add_argument("--precision", default="tf32", choices=["fp32", "tf32"],
                   help="tf32: tensor-core kernels (TF32 / bf16 operands, fp32 accumulate); fp32: exact FFMA kernels")
    args = p.parse_args(argv)
    args.live = not args.frozen          # main.py:1537
    args.inject = not args.noinject      # main.py:1538
    args.name = args.model               # main.py:1539
    if not torch.cuda.is_available():
        raise SystemExit("cgs_b200: no CUDA device (the product has no CPU path)")
    ops.set_precision(args.precision)
    H = Handler(args)
    if args.train:                       # main.py:1549-1550 (load_data -> collect_data needs MineRL: synthetic / .npz instead)
        if args.data:
            d = np.load(args.data)
            X, Y = d["X"], d["Y"]
        else:
            X, Y, _ = synth.synthetic_frames(args.synthetic or 6000, seed=0)
        H.set_data(X, Y)
    if args.cload:                       # main.py:1554-1557
        H.load_models([H.criticname])
    if args.mload:
        H.load_models([H.maskername])
    if args.train:                       # main.py:1558-1564
        if args.critic:
            H.critic_pipe(mode="train")
            H.save_models([H.criticname])
        if args.masker:
            H.segmentation_training()
            H.save_models([H.maskername])
    if args.process:                     # main.py:1569-1570
        H.segment(args.source_imgs)
    return 0


if __name__ == "__main__":
    sys.exit(main())
