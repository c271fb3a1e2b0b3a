"""Not collected by pytest: the host<->device copy ceiling that bounds the `e2e` number of bench.py (host-resident particle arrays,
nk_advance_host).  Every rank moves `gb` GB up and `gb` GB down CONCURRENTLY (two streams, pinned buffers bound to the GPU's NUMA
node like bench.py does) while all other ranks do the same; prints one JSON line with the per-rank and aggregate GB/s.

    python tests/run_pcie_ceiling.py [gb=4]                                      (1 GPU)
    torchrun --nproc-per-node N tests/run_pcie_ceiling.py [gb=4]                 (N GPUs of one box)

With e2e moving 44 B up and 40 B down per particle and step, updates/s <= min(h2d / 44, d2h / 40) per rank."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    gb = float(kv.get("gb", 4))
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    numa = None
    if world > 1:
        import torch.distributed as dist
        from nanokappa_b200.parallel import bind_to_gpu_numa
        numa = bind_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = int(gb * 1e9 / 8)
    up_h = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
    dn_h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    up_d = torch.empty(n, dtype=torch.float64, device="cuda"); dn_d = torch.ones(n, dtype=torch.float64, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d_only", "d2h_only", "both"):
        for rep in range(2):                           # first repetition = warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if mode != "d2h_only":
                with torch.cuda.stream(s_up):
                    e[0].record(); up_d.copy_(up_h, non_blocking=True); e[1].record()
            if mode != "h2d_only":
                with torch.cuda.stream(s_dn):
                    e[2].record(); dn_h.copy_(dn_d, non_blocking=True); e[3].record()
            torch.cuda.synchronize()
        r = {}
        if mode != "d2h_only":
            r["h2d_gbs"] = gb / (e[0].elapsed_time(e[1]) * 1e-3)
        if mode != "h2d_only":
            r["d2h_gbs"] = gb / (e[2].elapsed_time(e[3]) * 1e-3)
        res[mode] = r
    both = res["both"]
    mine = torch.tensor([both["h2d_gbs"], both["d2h_gbs"], res["h2d_only"]["h2d_gbs"], res["d2h_only"]["d2h_gbs"]], device="cuda", dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    if rank == 0:
        h2d = [float(v[0]) for v in allv]; d2h = [float(v[1]) for v in allv]
        ceiling = sum(min(a / 44.0, b / 40.0) * 1e9 for a, b in zip(h2d, d2h))
        print(json.dumps({"n_gpus": world, "gb_each_way_per_rank": gb, "numa_node_rank0": numa,
                          "concurrent_h2d_gbs_per_rank": h2d, "concurrent_d2h_gbs_per_rank": d2h,
                          "h2d_alone_gbs_per_rank": [float(v[2]) for v in allv], "d2h_alone_gbs_per_rank": [float(v[3]) for v in allv],
                          "aggregate_h2d_gbs": sum(h2d), "aggregate_d2h_gbs": sum(d2h),
                          "e2e_ceiling_updates_per_s": ceiling,
                          "note": "ceiling = sum over ranks of min(h2d / 44 B, d2h / 40 B): what nk_advance_host could reach if it did nothing but copy"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
