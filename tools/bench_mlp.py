"""Class-API MLP kernels at the a2c_test.py shape (config 3): 500 -> 6 -> 6 -> 6, batch 65536.  GB/s of X."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ia2c_b200 import _lib  # noqa: E402

lib = _lib.load()
rows, F, O = 65536, 500, 6
P = 6 * F + 6 + 36 + 6 + O * 6 + O
x = torch.nn.functional.one_hot(torch.randint(0, F, (rows,)), F).float().cuda()
params = (torch.randn(P) * 0.1).cuda()
y = torch.empty(rows, O, device="cuda")
h1 = torch.empty(rows, 6, device="cuda")
dy = torch.randn(rows, O, device="cuda")
grad = torch.zeros(P, device="cuda")
ws = torch.empty(lib.ia2c_mlp_backward_workspace(rows, F, O), device="cuda")
st = _lib.stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


xb = rows * F * 4 / 1e9
t = timeit(lambda: _lib.check(lib.ia2c_mlp_forward(_lib.ptr(params), _lib.ptr(x), _lib.ptr(y), _lib.ptr(h1), rows, F, O, 1, 1, st)))
print(f"forward            {t*1e3:8.1f} us   {xb/t*1e3:7.1f} GB/s of X")
t = timeit(lambda: _lib.check(lib.ia2c_mlp_backward(_lib.ptr(params), _lib.ptr(x), _lib.ptr(dy), _lib.ptr(h1), _lib.ptr(grad), None, _lib.ptr(ws),
                                                    rows, F, O, 1, 0, st)))
print(f"backward (B1+B2+B3) {t*1e3:8.1f} us   {xb/t*1e3:7.1f} GB/s of X (read once: h1 saved by the forward)")
idx = x.argmax(-1).contiguous()
t = timeit(lambda: _lib.check(lib.ia2c_mlp_forward_index(_lib.ptr(params), _lib.ptr(idx), _lib.ptr(y), _lib.ptr(h1), rows, F, O, 1, st)))
print(f"forward, index input   {t*1e3:8.1f} us   (8 B of idx + 24 B of h1 + {4*O} B of y per row)")
t = timeit(lambda: _lib.check(lib.ia2c_mlp_backward_index(_lib.ptr(params), _lib.ptr(idx), _lib.ptr(dy), _lib.ptr(h1), _lib.ptr(grad), _lib.ptr(ws),
                                                          rows, F, O, 1, 0, st)))
print(f"backward, index input  {t*1e3:8.1f} us")
