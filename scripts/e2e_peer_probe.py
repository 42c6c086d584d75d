"""World-of-one probe of the public-API (e2e) loop with each gradient all-gather transport.

At N = 8 the e2e line of bench.py was 1.5-1.8 ms slower with the peer-memory transports than with
NCCL on another box while `value` was faster; this reproduces bench.py's e2e loop in a world of ONE rank
(the peer path then pushes to itself) so that anything intrinsic to "peer transport + public API +
H2D/D2H on side streams" shows on one GPU, and times the pieces: the step, the staging copy out of
the landing buffer, and the same copy out of ordinary memory."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                       # noqa: E402
from kaolin_b200.multi_gpu import make_grad_all_gather             # noqa: E402
from kaolin_b200.render.mesh import dibr_rasterization             # noqa: E402


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29871", rank=0, world_size=1, device_id=dev)
    B, F, H, W, D, fvz, fvi, fnz, ff = bench.make_scene("c4_shard", 0)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h = [pin(fvz), pin(fvi), pin(ff), pin(fnz)]
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    g_feat = torch.rand((B, H, W, D), device=dev, generator=gen)
    g_soft = torch.rand((B, H, W), device=dev, generator=gen)
    out = {}
    for transport in ("none", "nccl", "peer", "peer_sm", "peer_mc"):
        try:
            NB = 2
            s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            s_cmp = torch.cuda.current_stream(dev)
            dev_in = []
            for _ in range(NB):
                bufs = [torch.empty_like(x, device=dev) for x in h]
                bufs[1].requires_grad_(True); bufs[2].requires_grad_(True)
                dev_in.append(bufs)
            host_out = [(torch.empty((B, F, 3, 2)).pin_memory(), torch.empty((B, F, 3, D)).pin_memory()) for _ in range(NB)]
            dev_out = [tuple(torch.empty(t.shape, device=dev) for t in host_out[k]) for k in range(NB)]
            ev_free, ev_read = [None] * NB, [None] * NB
            state = {"i": 0}
            stage_ev = []

            def step(timed=False):
                i = state["i"]; slot = i % NB
                a_fvz, a_fvi, a_ff, a_fnz = dev_in[slot]
                with torch.cuda.stream(s_h2d):
                    if ev_free[slot] is not None:
                        s_h2d.wait_event(ev_free[slot])
                    with torch.no_grad():
                        a_fvz.copy_(h[0], non_blocking=True); a_fvi.copy_(h[1], non_blocking=True)
                        a_ff.copy_(h[2], non_blocking=True); a_fnz.copy_(h[3], non_blocking=True)
                    ev_up = torch.cuda.Event(); ev_up.record(s_h2d)
                s_cmp.wait_event(ev_up)
                a_fvi.grad = None; a_ff.grad = None
                feat, soft, idx = dibr_rasterization(H, W, a_fvz, a_fvi, a_ff, a_fnz, bench.SIGMAINV, bench.BOXLEN, bench.KNUM)
                if transport == "none":
                    torch.autograd.backward([feat, soft], [g_feat, g_soft])
                    g1, g2 = a_fvi.grad, a_ff.grad
                else:
                    gather, _ = make_grad_all_gather(B, (B, F, 3, 2), (B, F, 3, D), dev, transport=transport)
                    gather.attach(soft)
                    torch.autograd.backward([feat, soft], [g_feat, g_soft])
                    g1, g2 = gather.finish(a_fvi.grad, a_ff.grad)
                if ev_read[slot] is not None:
                    ev_read[slot].synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s_cmp)
                with torch.no_grad():
                    dev_out[slot][0].copy_(g1); dev_out[slot][1].copy_(g2)
                e1.record(s_cmp)
                if timed:
                    stage_ev.append((e0, e1))
                ev_done = torch.cuda.Event(); ev_done.record(s_cmp)
                ev_free[slot] = ev_done
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev_done)
                    host_out[slot][0].copy_(dev_out[slot][0], non_blocking=True)
                    host_out[slot][1].copy_(dev_out[slot][1], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s_d2h)
                ev_read[slot] = ev
                state["i"] = i + 1

            for _ in range(5):
                step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); a.record()
            n = 20
            host_ms = []
            for _ in range(n):
                t1 = time.perf_counter(); step(True); host_ms.append((time.perf_counter() - t1) * 1e3)
            for e in ev_read:
                e.synchronize()
            b.record(); torch.cuda.synchronize()
            out[transport] = {"ms_per_step": a.elapsed_time(b) / n, "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / n,
                              "host_issue_ms_median": sorted(host_ms)[n // 2],
                              "staging_copy_ms": sum(x.elapsed_time(y) for x, y in stage_ev) / len(stage_ev)}
        except Exception as exc:
            out[transport] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        print(transport, out[transport], flush=True)
    print("E2E_PROBE " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
