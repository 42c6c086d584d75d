"""torchrun worker of tests/test_multi_gpu_gpu.py: sharded (NCCL) result == single-GPU result.

Every rank renders the FULL batch on its own GPU (the single-GPU answer), then only its
contiguous view shard, all-gathers the per-view gradients over NCCL with both gather
strategies, and compares: face_idx / soft_mask / features of the shard bit-equal to the
corresponding slice of the full render; gathered gradients within 1e-5 of the full ones
(SURVEY.md §8e "correctness oracle")."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kaolin_b200 import synthetic                                            # noqa: E402
from kaolin_b200.multi_gpu import (ChunkedGradAllGather, OverlappedGradAllGather, PeerGradAllGather,  # noqa: E402
                                   PipelinedGradAllGather, all_gather_view_grads, chunk_ranges, shard_range)
from kaolin_b200.render.mesh import dibr_rasterization                       # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    dist.init_process_group("nccl", device_id=dev)
    B, H, W, D = 4 * world, 256, 256, 3
    fvz, fvi, fnz = synthetic.icosphere_views(B, 4, seed=77)
    ff = synthetic.random_features(B, fvz.shape[1], D, seed=78)
    T = lambda a: torch.from_numpy(a).to(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    g_feat = torch.rand((B, H, W, D), device=dev, generator=gen)
    g_soft = torch.rand((B, H, W), device=dev, generator=gen)

    # single-GPU answer on the concatenated batch
    f_fvi, f_ff = T(fvi).requires_grad_(True), T(ff).requires_grad_(True)
    feat, soft, idx = dibr_rasterization(H, W, T(fvz), f_fvi, f_ff, T(fnz))
    torch.autograd.backward([feat, soft], [g_feat, g_soft])
    full = {"feat": feat.detach(), "soft": soft.detach(), "idx": idx, "g_fvi": f_fvi.grad, "g_ff": f_ff.grad}

    s0, s1 = shard_range(B, rank, world)
    out = {"rank": rank, "world": world, "views": [s0, s1]}
    ok = True
    for mode in ("overlapped", "pipelined", "chunked"):
        l_fvi, l_ff = T(fvi[s0:s1]).requires_grad_(True), T(ff[s0:s1]).requires_grad_(True)
        if mode in ("overlapped", "pipelined"):
            feat, soft, idx = dibr_rasterization(H, W, T(fvz[s0:s1]), l_fvi, l_ff, T(fnz[s0:s1]))
            if mode == "overlapped":
                gather = OverlappedGradAllGather(B).attach(soft)
                torch.autograd.backward([feat, soft], [g_feat[s0:s1], g_soft[s0:s1]])
                g_fvi, g_ff = gather.finish(l_fvi.grad, l_ff.grad)
            else:
                gather = PipelinedGradAllGather(chunks=2).attach(soft)
                torch.autograd.backward([feat, soft], [g_feat[s0:s1], g_soft[s0:s1]])
                g_fvi, g_ff = gather.finish()
            same = (torch.equal(idx, full["idx"][s0:s1]) and torch.equal(soft, full["soft"][s0:s1])
                    and torch.equal(feat, full["feat"][s0:s1]))
        else:
            n = s1 - s0
            gather = ChunkedGradAllGather(n)
            same = True
            for c0, c1 in chunk_ranges(n, 2):
                c_fvi = l_fvi[c0:c1].detach().requires_grad_(True)
                c_ff = l_ff[c0:c1].detach().requires_grad_(True)
                feat, soft, idx = dibr_rasterization(H, W, T(fvz[s0 + c0:s0 + c1]), c_fvi, c_ff,
                                                     T(fnz[s0 + c0:s0 + c1]))
                torch.autograd.backward([feat, soft], [g_feat[s0 + c0:s0 + c1], g_soft[s0 + c0:s0 + c1]])
                gather.submit(c0, c1, [c_fvi.grad, c_ff.grad])
                same = same and torch.equal(idx, full["idx"][s0 + c0:s0 + c1]) \
                    and torch.equal(soft, full["soft"][s0 + c0:s0 + c1])
            g_fvi, g_ff = gather.finish()
        e1, e2 = rel(g_fvi, full["g_fvi"]), rel(g_ff, full["g_ff"])
        out[mode] = {"images_bit_equal": bool(same), "grad_fvi_rel": e1, "grad_ff_rel": e2,
                     "gathered_shape": list(g_fvi.shape)}
        ok = ok and same and e1 <= 1e-5 and e2 <= 1e-5 and tuple(g_fvi.shape) == tuple(full["g_fvi"].shape)
    # all-gather by stores into peer memory (NVLink): three consecutive steps with different data through
    # the persistent double-buffered landing areas; every step bit-equal to the NCCL all-gather of the same
    # local gradients, the first also within 1e-5 of the single-GPU answer
    for engine in ("ce", "sm", "mc"):
        name = "peer_" + engine
        try:
            steps, exact, e1, e2 = [], True, None, None
            for it in range(3):
                scale = 1.0 + 0.5 * it
                l_fvi, l_ff = T(fvi[s0:s1]).requires_grad_(True), T(ff[s0:s1]).requires_grad_(True)
                feat, soft, idx = dibr_rasterization(H, W, T(fvz[s0:s1]), l_fvi, l_ff, T(fnz[s0:s1]))
                gather = PeerGradAllGather(B, l_fvi.shape, l_ff.shape, dev, engine=engine).attach(soft)
                torch.autograd.backward([feat, soft], [g_feat[s0:s1] * scale, g_soft[s0:s1] * scale])
                g_fvi, g_ff = gather.finish(l_fvi.grad, l_ff.grad)
                n_fvi, n_ff = all_gather_view_grads([l_fvi.grad, l_ff.grad], B)
                exact = exact and torch.equal(g_fvi, n_fvi) and torch.equal(g_ff, n_ff)
                if it == 0:
                    e1, e2 = rel(g_fvi, full["g_fvi"]), rel(g_ff, full["g_ff"])
                steps.append(float(g_ff.abs().sum()))
            out[name] = {"available": True, "equal_to_nccl_all_gather": bool(exact), "grad_fvi_rel": e1,
                         "grad_ff_rel": e2, "step_sums": steps}
            ok = ok and exact and e1 <= 1e-5 and e2 <= 1e-5 and steps[0] != steps[1]
        except Exception as exc:     # symmetric memory not available in this container: reported, not a failure
            out[name] = {"available": False, "why": f"{type(exc).__name__}: {exc}"[:400]}
    out["ok"] = bool(ok)
    sys.stdout.write("\nMGPU_RESULT " + json.dumps(out) + "\n")     # one write per rank
    sys.stdout.flush()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
