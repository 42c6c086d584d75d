"""dibr_b200_peer_push (csrc/peer_push.cu) on one GPU: the destinations are local buffers, which is
all the kernel knows about them (a peer pointer is a pointer); and PeerGradAllGather in a world of
one rank, which exercises the symmetric-memory setup, the landing layout, the double buffering and
the barrier without a second GPU (tests/test_multi_gpu_gpu.py covers two)."""
import ctypes
import json
import os
import subprocess
import sys

import pytest
import torch

from kaolin_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n16,n_dst,ctas", [(1, 1, 1), (777, 3, 2), (4 * 512 * 5 + 13, 8, 5), (1 << 20, 16, 32)])
def test_peer_push_copies_to_every_destination(n16, n_dst, ctas):
    dev = "cuda"
    src = torch.randint(-2 ** 31, 2 ** 31 - 1, (n16 * 4,), dtype=torch.int32, device=dev)
    off16 = 5
    dst = [torch.full(((n16 + off16 + 3) * 4,), 7, dtype=torch.int32, device=dev) for _ in range(n_dst)]
    arr = (ctypes.c_void_p * n_dst)(*[d.data_ptr() for d in dst])
    st = _lib.lib().dibr_b200_peer_push(ctypes.c_void_p(src.data_ptr()), n16 * 16, arr, n_dst, off16 * 16, ctas,
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 0
    torch.cuda.synchronize()
    for d in dst:
        assert torch.equal(d[off16 * 4:(off16 + n16) * 4], src)
        assert (d[:off16 * 4] == 7).all() and (d[(off16 + n16) * 4:] == 7).all()


def test_peer_push_rejects_bad_arguments():
    lib = _lib.lib()
    buf = torch.zeros(64, dtype=torch.int32, device="cuda")
    arr = (ctypes.c_void_p * 1)(buf.data_ptr())
    p = ctypes.c_void_p(buf.data_ptr())
    assert lib.dibr_b200_peer_push(p, 24, arr, 1, 0, 0, None) == _lib.EINVAL          # not a multiple of 16
    assert lib.dibr_b200_peer_push(p, 32, arr, 1, 8, 0, None) == _lib.EINVAL          # misaligned offset
    assert lib.dibr_b200_peer_push(p, 32, arr, 17, 0, 0, None) == _lib.EINVAL         # too many destinations
    assert lib.dibr_b200_peer_push(None, 32, arr, 1, 0, 0, None) == _lib.EINVAL
    assert lib.dibr_b200_peer_push(p, 0, arr, 1, 0, 0, None) == 0                     # nothing to do


_WORKER = r"""
import json, os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from kaolin_b200.multi_gpu import PeerGradAllGather
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%%d" %% (29900 + os.getpid() %% 90), rank=0, world_size=1,
                        device_id=dev)
out = {}
for engine in ("ce", "sm"):
    try:
        good = True
        for it in range(3):
            g_fvi = torch.randn(4, 320, 3, 2, device=dev)
            g_ff = torch.randn(4, 320, 3, 3, device=dev)
            gather = PeerGradAllGather(4, g_fvi.shape, g_ff.shape, dev, engine=engine)
            gather.hook(g_ff)
            a, b = gather.finish(g_fvi)
            good = good and torch.equal(a, g_fvi) and torch.equal(b, g_ff)
        out[engine] = {"available": True, "ok": bool(good)}
    except Exception as exc:
        out[engine] = {"available": False, "why": ("%%s: %%s" %% (type(exc).__name__, exc))[:300]}
print("PEER1 " + json.dumps(out))
dist.destroy_process_group()
""" % ROOT


def test_peer_all_gather_world_of_one():
    r = subprocess.run([sys.executable, "-c", _WORKER], capture_output=True, text=True, timeout=300, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("PEER1 ")]
    assert r.returncode == 0 and line, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(line[0][6:])
    print("\npeer all-gather, world of one:", res)
    for engine, x in res.items():
        if x["available"]:
            assert x["ok"], engine
