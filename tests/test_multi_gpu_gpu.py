"""Hardware check of the multi-GPU oracle: the view-sharded NCCL result equals
the single-GPU result on the concatenated batch.  Needs >= 2 GPUs (skipped otherwise);
spawns tests/_mgpu_worker.py with torchrun, one process per GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_nccl_equals_single_gpu():
    world = 2
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    # the two ranks write to the same pipe: their lines can interleave, so parse every object that
    # follows a marker instead of whole lines
    dec, results, pos = json.JSONDecoder(), [], 0
    while True:
        pos = r.stdout.find("MGPU_RESULT ", pos)
        if pos < 0:
            break
        obj, end = dec.raw_decode(r.stdout, pos + len("MGPU_RESULT "))
        results.append(obj)
        pos = end
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(results) == world and all(x["ok"] for x in results)
    for x in results:
        for mode in ("overlapped", "pipelined", "chunked"):
            assert x[mode]["images_bit_equal"]
            assert x[mode]["grad_fvi_rel"] <= 1e-5 and x[mode]["grad_ff_rel"] <= 1e-5
        for mode in ("peer_ce", "peer_sm", "peer_mc"):       # all-gather by stores into peer memory, when the box offers it
            if x[mode]["available"]:
                assert x[mode]["equal_to_nccl_all_gather"]
                assert x[mode]["grad_fvi_rel"] <= 1e-5 and x[mode]["grad_ff_rel"] <= 1e-5
            else:
                print("peer-memory all-gather unavailable:", x[mode]["why"])
