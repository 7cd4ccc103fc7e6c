#!/usr/bin/env python
"""bench.py -- scene-pairs/sec of the PointContrast hot path (Res16UNet34C + PointInfoNCE, 2.5 cm voxels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--loss nce|hardest] [--workload c1|c0]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one full training iteration on one batch of synthetic scene pairs per rank: coordinate-manager build for
both views (stacked in one pass by default; PCB_PAIR=0: two forward calls), loss, backward, gradient all-reduce (N > 1), fused SGD step.  Prints ONE JSON line (rank 0).

  value    : pairs/s with the batch already resident in HBM, CUDA-event timed, max over ranks.
  e2e      : pairs/s through the public trainer call (`Trainer._train_iter`) with the batch in pinned HOST memory,
             host->device copies and the loss read-back inside the timed region.
  roofline : the dominant kernel (conv_tcgen05_split_kernel: sparse-conv forward / data-gradient) -- algorithmic bytes
              of all its launches in one step / their CUDA-event time, vs the measured HBM peak.
  cpu_baseline : the oracle (ME-0.4.3-algorithm CPU restatement) timed on this box's host cores, bounded sample.

--impl reference times that CPU restatement as the whole measurement (the reference's own arithmetic layer,
MinkowskiEngine 0.4.3, is not in the reference tree and not installable offline -- DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = "scene-pairs/sec Res16UNet34C PointInfoNCE @2.5cm voxel"
WORKLOADS = {   # per-rank batch, synthetic scale -> ~voxels/view
    "c1": dict(batch=4, scale=0.9, desc="BASELINE configs[1]: ~40k voxels/view synthetic ScanNet-shape, batch=4 per GPU"),
    "c0": dict(batch=1, scale=0.24, desc="BASELINE configs[0]: ~4k voxels/view, batch=1"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loss", default="nce", choices=["nce", "hardest"])
    ap.add_argument("--workload", default="c1", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-launch conv profile of one step here")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[2:]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def usable_cores():
    """Host cores this process may really use: affinity mask, capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


# ----------------------------------------------------------------------------------------------- CPU oracle leg
def cpu_oracle_steps(workload, steps, warmup, loss_kind, budget_s):
    """ME-0.4.3-algorithm CPU restatement (the oracle), fp32, torch's default thread count.  Each step is a full training
    step (2x fwd, loss, bwd, SGD) on a BOUNDED SAMPLE of the workload: one scene pair, shrunk (synthetic room scale) so
    that warmup+steps fit `budget_s`; the result is converted to full-size pairs/s by the voxel fraction processed."""
    from oracle import loss_cpu, me_cpu as OR
    from pointcontrast_b200 import synth
    from pointcontrast_b200.config import default_config
    from pointcontrast_b200.model import res16unet
    cores = usable_cores()
    torch.set_num_threads(cores)
    full_scale = WORKLOADS[workload]["scale"]
    full = synth.synth_pair(0, scale=full_scale)
    n_full = len(full["coords0"]) + len(full["coords1"])
    old = res16unet.ME
    res16unet.ME = OR
    try:
        cfg = default_config()
        net = res16unet.Res16UNet34C(3, 32, cfg, D=3).train()
        opt = torch.optim.SGD(net.parameters(), lr=cfg.opt.lr, momentum=cfg.opt.momentum, weight_decay=cfg.opt.weight_decay)
        rng = np.random.default_rng(0)

        def one_step(batch):
            t0 = time.perf_counter()
            opt.zero_grad()
            F = [net(OR.SparseTensor(torch.from_numpy(batch[f"sinput{v}_F"]), coords=torch.from_numpy(batch[f"sinput{v}_C"]))).F
                 for v in "01"]
            pairs = batch["correspondences"]
            if loss_kind == "nce":
                nq = len(np.unique(pairs[:, 0]))
                q, k = loss_cpu.select_positives(pairs, rng.random(nq).astype(np.float32), 4096,
                                                 rng.choice(nq, 4096, replace=False) if nq > 4096 else None)
                loss = loss_cpu.point_nce_loss(F[0], F[1], q, k, 0.4)
            else:
                sel0 = rng.choice(len(F[0]), min(256, len(F[0])), replace=False)
                sel1 = rng.choice(len(F[1]), min(256, len(F[1])), replace=False)
                ps = rng.choice(len(pairs), 1024, replace=False) if len(pairs) > 1024 else None
                a, b = loss_cpu.hardest_contrastive_loss(F[0], F[1], pairs, sel0, sel1, ps)
                loss = a + b
            loss.backward()
            opt.step()
            return time.perf_counter() - t0

        cal = synth.collate_pairs([synth.synth_pair(1, scale=min(0.2, full_scale))])
        one_step(cal)                                           # thread pools, allocator
        t_cal = one_step(cal)
        n_cal = len(cal["sinput0_C"]) + len(cal["sinput1_C"])
        per_step = budget_s / max(1, steps + warmup)
        n_target = per_step / (t_cal / n_cal)
        scale = full_scale * float(np.sqrt(min(1.0, n_target / n_full)))
        scale = max(scale, min(0.2, full_scale))
        batch = synth.collate_pairs([synth.synth_pair(0, scale=scale)])
        n_s = len(batch["sinput0_C"]) + len(batch["sinput1_C"])
        frac = min(1.0, n_s / n_full)
        times = [one_step(batch) for _ in range(warmup + steps)][warmup:]
    finally:
        res16unet.ME = old
    t_s = float(np.mean(times))
    if frac >= 0.999 or n_s <= n_cal:
        t_full, how = t_s / frac, "direct"
    else:       # step time is affine in the voxel count (fixed part: 37.8M-parameter SGD, per-layer overheads): two-point fit
        b = max(0.0, (t_s - t_cal) / (n_s - n_cal))
        a = max(0.0, t_s - b * n_s)
        t_full, how = a + b * n_full, f"affine fit t = {a:.2f} s + {b * 1e3:.3f} ms/voxel from ({n_cal} voxels, {t_cal:.2f} s) and ({n_s}, {t_s:.2f} s)"
    return dict(value=1.0 / t_full, unit="pairs/s", cores=cores, kind="port",
                sample=f"{len(times)} full training steps (2x fwd, loss, bwd, SGD) on one synthetic scene pair of {n_s} voxels "
                       f"({frac:.3f} of a full '{workload}' pair, {n_full} voxels); fp32, {cores} torch threads; "
                       f"scaled to a full pair: {how}"), t_full * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms = cpu_oracle_steps(args.workload, args.steps, args.warmup, args.loss, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "loss": args.loss,
                       "note": "ME-0.4.3-algorithm CPU restatement (oracle); MinkowskiEngine itself is not in the reference tree"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def conv_alg_bytes(rec):
    M = sum(rec["plan"].pair_counts())
    return M * (rec["Cin"] + rec["Cout"]) * 4 + M * 8 + rec["K"] * rec["Cin"] * rec["Cout"] * 4, 2 * M * rec["Cin"] * rec["Cout"]


def run_ours(args):
    import torch.distributed as dist
    from pointcontrast_b200 import _lib, fused, me
    from pointcontrast_b200.config import default_config
    from pointcontrast_b200.data import SyntheticPairLoader
    from pointcontrast_b200.trainer import get_trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    cfg = default_config([f"trainer.batch_size={wl['batch'] * world}", f"misc.num_gpus={world}", "misc.nceT=0.4"])
    loader = SyntheticPairLoader(wl["batch"], scale=wl["scale"], num_batches=2, rank=rank, pin=True)
    torch.manual_seed(0)
    name = "PointNCELossTrainer" if args.loss == "nce" else "HardestContrastiveLossTrainer"
    trainer = get_trainer(name)(cfg, loader)
    host_batches = loader.batches
    keys = ("sinput0_C", "sinput0_F", "sinput1_C", "sinput1_F", "correspondences")
    dev_batches = [{k: (v.to(dev) if k in keys else v) for k, v in b.items()} for b in host_batches]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing
    for i in range(args.warmup):
        trainer.train_step(dev_batches[i % len(dev_batches)])
    sync_all()
    clocks = Clocks(local) if rank == 0 else None
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        loss = trainer.train_step(dev_batches[i % len(dev_batches)])
    e1.record()
    sync_all()
    t_wall1 = time.time()
    launches = _lib.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    pairs_per_step = wl["batch"] * world
    value = pairs_per_step * args.steps / (ms_total / 1e3)

    # ---- end-to-end through the public trainer call, host (pinned) batches
    it = iter(loader)
    for _ in range(2):
        trainer._train_iter(it, None)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        trainer._train_iter(it, None)
    e1.record()
    sync_all()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = pairs_per_step * args.steps / (float(ms2.item()) / 1e3)
    h2d = int(np.mean([sum(b[k].numel() * b[k].element_size() for k in keys) for b in host_batches]))

    # ---- roofline of the dominant kernel: one instrumented step, CUDA events around every conv launch
    roof = None
    if rank == 0:
        me.PROFILE = []
    trainer.train_step(dev_batches[0])          # every rank runs it (the step contains the gradient all-reduce); rank 0 records
    sync_all()
    if rank == 0:
        prof, me.PROFILE = me.PROFILE, None
        peak, peak_src = peaks()
        agg = {}
        for r in prof:
            b, f = conv_alg_bytes(r)
            t = r["ev0"].elapsed_time(r["ev1"])
            tc5 = me.CONV_IMPL == "tcgen05"
            key = (("conv_tcgen05_split_kernel" if tc5 else "conv_mma_kernel") if r["kind"] in ("fwd", "dgrad")
                   else ("wgrad_tcgen05_kernel" if tc5 else "wgrad_mma_kernel")) if r["tc"] else "stem_fp32 (conv_simt / wgrad_stem)"
            a = agg.setdefault(key, dict(bytes=0, flops=0, ms=0.0, launches=0))
            a["bytes"] += b; a["flops"] += f; a["ms"] += t; a["launches"] += 1
            r["bytes"], r["flops"], r["ms"] = b, f, t
        dom = max(agg, key=lambda k: agg[k]["ms"])
        a = agg[dom]
        traffic = None
        try:        # dram__bytes_read+write per launch of this kernel, from the committed ncu pass (profiles/traffic.json)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["kernels"][dom]["dram_bytes_per_launch"]
        except Exception:
            pass
        roof = {"kernel": dom, "bound": "hbm", "achieved": a["bytes"] / (a["ms"] / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": a["bytes"] / (a["ms"] / 1e3) / 1e9 / peak, "traffic": None if fused.PAIR else traffic, "peak_source": peak_src,
                "launches_per_step": a["launches"], "avg_launch_ms": a["ms"] / a["launches"],
                "alg_bytes_per_launch": a["bytes"] / a["launches"], "tensor_tflops": a["flops"] / (a["ms"] / 1e3) / 1e12,
                "share_of_step": a["ms"] / (ms_total / args.steps),
                "traffic_note": "ncu pass of the two-forward schedule (profiles/r1_launch_list.md)" if not fused.PAIR else
                                f"no ncu pass of the stacked schedule yet (twice the rows per launch); the two-forward schedule "
                                f"measured {traffic} dram bytes/launch (profiles/r1_launch_list.md)",
                "other": {k: {"ms": v["ms"], "GB/s": v["bytes"] / (v["ms"] / 1e3) / 1e9, "launches": v["launches"]}
                          for k, v in agg.items() if k != dom}}
        if args.profile_json:
            slim = [{k: v for k, v in r.items() if k not in ("plan", "ev0", "ev1")} for r in prof]
            json.dump({"per_launch": slim, "agg": agg, "ms_per_step": ms_total / args.steps}, open(args.profile_json, "w"), indent=1)

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_oracle_steps(args.workload, 3, 1, args.loss, budget_s=25.0)
        n0 = int(np.mean([len(b["sinput0_C"]) for b in host_batches])); n1 = int(np.mean([len(b["sinput1_C"]) for b in host_batches]))
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (bf16x3-split tensor-core products, fp32 accumulate)", "data": "synthetic",
                "config": {"workload": wl["desc"], "loss": args.loss, "pairs_per_gpu": wl["batch"], "global_batch": pairs_per_step,
                           "voxels_per_view_per_rank": [n0, n1], "parallelism": f"dp{world}",
                           "schedule": "both views stacked in one pass (per-view BatchNorm statistics)" if fused.PAIR
                           else "two forward calls",
                           "l2": "per-step working set (activations + kernel maps, GBs) far exceeds the 126 MB L2; 2 distinct batches cycled",
                           "final_loss": float(loss[0] if isinstance(loss, tuple) else loss)},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "roofline": roof, "cpu_baseline": cb}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
