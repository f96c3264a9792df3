"""Equivalent of the reference's `basic/1u1m-basic.py`: list the local devices, keep the variables on the CPU and
run `w + b`, `w * b` on GPU 0 (1U:6-26). Expected output: the device list, [[7, 7], [7, 7]], [[10, 10], [10, 10]].
Falls back to the CPU when no GPU is visible."""
import torch


def get_available_devices():
    return ["/cpu:0"] + [f"/gpu:{i} ({torch.cuda.get_device_name(i)})" for i in range(torch.cuda.device_count())]


def main():
    print(get_available_devices())                       # 1U:6-11
    w = torch.full((2, 2), 2.0)                          # variables on /cpu:0, 1U:13-15
    b = torch.full((2, 2), 5.0)
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    wd, bd = w.to(dev), b.to(dev)                        # cross-device copy (X7), ops on /gpu:0, 1U:17-19
    addwb, mutwb = wd + bd, wd * bd
    print(addwb.cpu().numpy())
    print(mutwb.cpu().numpy())


if __name__ == "__main__":
    main()
