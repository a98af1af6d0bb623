"""N-rank data-parallel VAE+DFC train step == 1-rank step on the full batch, on real GPUs over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_equivalence.py

Every rank steps its shard (sync-BN statistic sums + gradient all-reduce); rank 0 additionally steps the full batch
alone and compares losses and the (global) gradient buffer.  Prints one JSON line; exit code 1 on mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from icsg3d_b200 import utils
from icsg3d_b200.engine import Dist, VAEEngine


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    Bg = 8
    Bl = Bg // world
    # identical global batch on every rank (same seed), then shard
    M, cond, _ = utils.synthetic_batch(Bg, d=32, seed=11, device=dev)
    eps = torch.randn(Bg, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    eng = VAEEngine(Bl, d=32, seed=3, device=dev, dist=Dist())
    sl = slice(rank * Bl, (rank + 1) * Bl)
    eng.set_inputs(M[sl], cond[sl], eps[sl])
    if eng.peer is not None:
        eng.peer.tick()  # the in-kernel BatchNorm all-reduce needs an epoch >= 1 (normally advanced by _train_body)
    eng.pack_weights()
    from icsg3d_b200 import ops
    eng.pack_inputs()
    eng.encode(True); eng.decode(True); eng.pm_forward(0, True); eng.pm_forward(1, True); eng.losses(); eng.backward()
    eng.dist.all_reduce_sum(eng.vp.grad)
    torch.cuda.synchronize()
    m_dp = eng.metrics_host()
    ok = True
    report = {}
    if rank == 0:
        ref = VAEEngine(Bg, d=32, seed=3, device=dev)
        ref.set_inputs(M, cond, eps)
        ref.pack_weights()
        ref.pack_inputs()
        ref.encode(True); ref.decode(True); ref.pm_forward(0, True); ref.pm_forward(1, True); ref.losses(); ref.backward()
        torch.cuda.synchronize()
        m_1 = ref.metrics_host()
        n = ref.vp.n_trainable
        g_dp, g_1 = eng.vp.grad[:n].double(), ref.vp.grad[:n].double()
        rel = float((g_dp - g_1).norm() / g_1.norm())
        cos = float((g_dp @ g_1) / (g_dp.norm() * g_1.norm()))
        dm = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(m_dp, m_1))
        ok = dm < 1e-3 and cos > 0.995
        report["equivalence"] = {"world": world, "global_batch": Bg, "metrics_dp": m_dp, "metrics_1rank": m_1,
                                 "max_rel_metric_delta": dm, "grad_rel_l2": rel, "grad_cos": cos, "ok": ok}
        print(json.dumps(report["equivalence"]))
    # ---- segmented CUDA-graph replay of the DP step == eager DP step (same kernels, same collectives) ----
    engA = VAEEngine(Bl, d=32, seed=3, device=dev, dist=Dist())
    engB = VAEEngine(Bl, d=32, seed=3, device=dev, dist=Dist())
    for e in (engA, engB):
        e.set_inputs(M[sl], cond[sl], eps[sl])
    engA.capture_train_graph()
    nseg = len(engA._segments)
    for _ in range(3):
        engA.train_step()
        engB.train_step()
    torch.cuda.synchronize()
    dtheta = float((engA.vp.theta - engB.vp.theta).abs().max())
    mA, mB = engA.metrics_host(), engB.metrics_host()
    ok_graph = dtheta == 0.0 and mA == mB
    if rank == 0:
        report["graph_vs_eager"] = {"world": world, "graph_segments": nseg,
                                    "bn_allreduce": "peer-memory kernel" if engA.peer is not None else "nccl",
                                    "theta_max_abs_diff_graph_vs_eager": dtheta, "metrics_graph": mA, "metrics_eager": mB,
                                    "ok": bool(ok and ok_graph)}
        print(json.dumps(report["graph_vs_eager"]))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dp_equivalence_w{world}.json"), "w") as f:
            json.dump(report, f, indent=1)
    ok = ok and ok_graph
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
