"""HBM ceilings by access mix on this GPU (torch ops; explains the head kernel's 29 % read / 71 % write mix).
    python bench_micro/hbm_probe.py      -> one JSON line"""
import json
import torch

def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3

N = 1 << 29            # 2 GiB of f32
x = torch.empty(N, device="cuda"); y = torch.empty(N, device="cuda"); z = torch.empty(N // 2, device="cuda")
x.normal_()
res = {"fill_write_only": 4 * N / t(lambda: y.fill_(1.0)) / 1e9,
       "copy_1r_1w": 8 * N / t(lambda: y.copy_(x)) / 1e9,
       "sum_read_only": 4 * N / t(lambda: x.sum()) / 1e9,
       # 2 reads + 5 writes of the same size, the head kernel's DUSty-I mix (8 B in, 20 B out per pixel)
       }
a = torch.empty(2, N // 8, device="cuda").normal_(); o = torch.empty(5, N // 8, device="cuda")
def mix():
    torch.add(a[0], a[1], out=o[0]); o[1].fill_(0.5); o[2].fill_(0.5); o[3].fill_(0.5); o[4].fill_(0.5)
res["mix_2r_5w_separate_kernels"] = 4 * (N // 8) * 7 / t(mix) / 1e9
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
