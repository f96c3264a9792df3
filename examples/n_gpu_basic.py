"""Equivalent of the reference's `basic/nu1m-basic.py`: variables on the CPU, `w + b` placed on GPU 0 and `w * b`
placed on GPU 1 (manual op-level placement, NU:3-22). Expected output: [[[7, 7], [7, 7]], [[10, 10], [10, 10]]].
With fewer than two GPUs the ops share whatever device exists."""
import torch


def main():
    w = torch.full((2, 2), 2.0)        # /cpu:0, NU:3-5
    b = torch.full((2, 2), 5.0)
    n = torch.cuda.device_count()
    d0 = "cuda:0" if n >= 1 else "cpu"
    d1 = "cuda:1" if n >= 2 else d0
    addwb = w.to(d0) + b.to(d0)        # /gpu:0, NU:8-10
    mutwb = w.to(d1) * b.to(d1)        # /gpu:1, NU:13-15
    print([addwb.cpu().tolist(), mutwb.cpu().tolist()])


if __name__ == "__main__":
    main()
