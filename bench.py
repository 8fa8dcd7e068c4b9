#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native R-YOLOv4 hot path (contract: task prompt ④).

Workload at every N (default): the model/size of BASELINE.json configs[1] — synthetic 800x800, bs=32 PER GPU,
yolov4 / csl / nc=2, random weights (reference init, train.py:28-33) — driven through ONE FULL TRAINING STEP
(train.py:184-202): train-mode forward (batch-statistics BatchNorm), ComputeCSLLoss value + gradient, conv-stack
backward (dgrad + wgrad), one NCCL all-reduce of the flat fp32 gradients when N>1, SGD(momentum .937, nesterov).
It is a superset of configs[1]'s "forward+loss", whose img/s is reported next to it (config.fwd_loss_img_s).
N>1: one rank per GPU, its own 32-image shard of the global batch (weak scaling).
--workload train_v7 runs BASELINE configs[2] (yolov7 / csl / nc=16) the same way.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

--impl reference times the CPU arm: the oracle's torch-CPU restatement of the reference forward +
loss (oracle/model_cpu.py + oracle/hotpath.py, pinned to the reference by tests/) on all host cores, on a
bounded sample (2 of the 32 images per step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(anchors=[[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]],
           angles=[-90, -60, -30, 0, 30, 60])                                   # data/hyp.yaml:2-7
HYP = dict(fl_gamma=0.0, box=0.05, obj=1.0, obj_pw=1.0, cls=0.5, cls_pw=1.0)   # data/hyp.yaml:11-17
S, BS, NC, PER_IMG = 800, 32, 2, 100
# 2*MAC over the convs @800^2: forward; dgrad = forward minus the stem; wgrad = forward
WORKLOADS = {
    "train_v4": dict(ver="yolov4", nc=2, fwd_gflop=218.14, stem_gflop=1.106,
                     name="yolov4/csl/nc2 800x800 full train step (fwd + CSL loss + bwd + SGD)"),
    "train_v7": dict(ver="yolov7", nc=16, fwd_gflop=168.38, stem_gflop=1.106,
                     name="yolov7/csl/nc16 800x800 full train step (fwd + CSL loss + bwd + SGD)"),
}
METRIC, UNIT = "training-step img/s, 800x800, bs=32 per GPU (synthetic, random weights)", "img/s"


def weights_init_normal(m):                                                   # train.py:28-33
    cn = m.__class__.__name__
    if cn.find("Conv2d") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif cn.find("BatchNorm2d") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


def gaussian_label(label, num_class=180, u=0, sig=6.0):
    x = np.arange(-num_class / 2, num_class / 2)
    y = np.exp(-(x - u) ** 2 / (2 * sig ** 2))
    i = int(num_class / 2 - label)
    return np.concatenate([y[i:], y[:i]], axis=0)


def make_targets(seed, bs, nc=NC):
    g = np.random.default_rng(seed)
    rows = []
    for b in range(bs):
        for _ in range(PER_IMG):
            w = g.uniform(0.02, 0.15)
            h = min(w * g.uniform(1, 3), 0.9)
            th = g.uniform(-np.pi / 2, np.pi / 2 - 1e-3)
            rows.append([b, float(g.integers(0, nc)), g.uniform(0.05, 0.95), g.uniform(0.05, 0.95), w, h, th]
                        + list(gaussian_label(th * 180 / np.pi + 90)))
    return torch.tensor(np.array(rows), dtype=torch.float32)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}", "--format=csv,noheader",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        num = lambda s: float(s.split()[0])
        load = [r for r in rows if num(r[3]) > 300] or rows
        reasons = set()
        for r in load:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([num(r[1]) for r in load])), "sm_max_mhz": num(rows[0][2]),
                "power_w_max": max(num(r[3]) for r in rows), "samples": len(rows), "reasons": sorted(reasons)}


class _M:
    def __init__(self, anchors, nc):
        self.anchors, self.nc = anchors, nc
        self._p = torch.nn.Parameter(torch.zeros(1))

    def parameters(self):
        return iter([self._p])


# ------------------------------------------------------------------------------------ CPU arm
def cpu_arm(wl, steps, warmup, sample_imgs=2):
    """Oracle port of the reference training step (forward, CSL loss, autograd backward, SGD) on the host cores.
    Returns img/s and a description of the bounded sample."""
    from oracle import hotpath as hp
    from oracle import model_cpu
    import ryolo_b200 as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    m = R.Yolo(wl["nc"], CFG, "csl", wl["ver"])
    m.apply(weights_init_normal)
    pnames = {k for k, _ in m.named_parameters()}
    sd = {k: v.clone().requires_grad_(k in pnames) for k, v in m.state_dict().items()}
    params = [sd[k] for k in sd if k in pnames]
    opt = torch.optim.SGD(params, lr=0.01, momentum=0.937, nesterov=True)
    img = torch.rand(sample_imgs, 3, S, S)
    tg = make_targets(0, sample_imgs, wl["nc"])
    an = hp.make_anchors(CFG["anchors"])
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        levels, _, stats = model_cpu.forward(sd, img, wl["ver"], "csl", wl["nc"], train=True, decode=False)
        loss, _ = hp.csl_loss(levels, tg, an, wl["nc"], HYP)
        loss.backward()
        opt.step()
        with torch.no_grad():
            for k, v in stats.items():
                sd[k].copy_(v)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    return sample_imgs / float(np.mean(ts)), cores, \
        f"{sample_imgs} of the {BS} images per step, {steps} timed step(s), torch CPU fp32 ({cores} threads)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload if args.workload != "fwd_loss" else "train_v4"]
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    v, cores, sample = cpu_arm(wl, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * 2 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": BS, "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch.distributed as dist
    import ryolo_b200 as R
    from ryolo_b200 import _lib as L
    from ryolo_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L.check(L.lib().ryolo_check_device(local))
    wl = WORKLOADS[args.workload if args.workload != "fwd_loss" else "train_v4"]
    nc = wl["nc"]
    torch.manual_seed(42)
    model = R.Yolo(nc, CFG, "csl", wl["ver"])
    model.apply(weights_init_normal)
    model = model.to(dev).train()
    crit = R.ComputeCSLLoss(model, HYP)
    crit.sync_items = False
    trainer = R.TrainStep(model, crit, lr=0.01, momentum=0.937, nesterov=True)      # train.py:156
    host_imgs = [torch.rand(BS, 3, S, S).pin_memory() for _ in range(2)]
    host_tg = [make_targets(rank * 2 + i, BS, nc).pin_memory() for i in range(2)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    dev_tg = [h.to(dev) for h in host_tg]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(n)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident arm ("value"): full training steps
    def resident(n):
        for i in range(n):
            trainer(dev_imgs[i & 1], dev_tg[i & 1])

    warm = max(args.warmup, 3)
    resident(warm)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.LAUNCHES[0]
    ops.PROFILE = []
    ms_total = timed(resident, args.steps)
    prof, ops.PROFILE = ops.PROFILE, None
    launches = L.LAUNCHES[0] - l0
    clocks = sampler.stop() if sampler else None
    tc_ms = {}
    for tag, a, b in prof:
        tc_ms[tag[0]] = tc_ms.get(tag[0], 0.0) + a.elapsed_time(b) / args.steps

    # ---- forward + loss only (BASELINE configs[1] wording), same model and inputs
    def fwd_loss(n):
        for i in range(n):
            lv = model(dev_imgs[i & 1], training=True)
            crit.value_and_grad(lv, dev_tg[i & 1])

    fwd_loss(2)
    ms_fl = timed(fwd_loss, min(args.steps, 5))
    fl_steps = min(args.steps, 5)

    # ---- end-to-end arm: pinned host -> device copies (prefetched on a side stream) + loss read-back
    copy_stream = torch.cuda.Stream(dev)
    stage_i = [torch.empty_like(dev_imgs[0]) for _ in range(2)]
    stage_t = [torch.empty_like(dev_tg[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    out_host = torch.empty(8, dtype=torch.float32).pin_memory()

    def upload(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            stage_i[s].copy_(host_imgs[s], non_blocking=True)
            stage_t[s].copy_(host_tg[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e(n):
        cur = torch.cuda.current_stream()
        for s in range(2):
            freed[s].record(cur)
        upload(0)
        for i in range(n):
            s = i & 1
            if i + 1 < n:
                upload(i + 1)
            cur.wait_event(ready[s])
            items = trainer(stage_i[s], stage_t[s])
            freed[s].record(cur)
            out_host.copy_(items, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_host[4])

    e2e(2)
    ms_e2e = timed(e2e, args.steps)

    if rank == 0:
        pk, pk_src = peaks()
        value = world * BS * args.steps / ms_total * 1e3
        # dominant kernel = conv_fwd_kernel (forward + dgrad launches, main stream).  The wgrad GEMMs run on a side
        # stream overlapped with dgrad / BN backward, so their per-launch durations include time-slicing and are
        # reported separately.
        gflop = (2 * wl["fwd_gflop"] - wl["stem_gflop"]) * BS          # forward + dgrad (no stem), per step
        tc_total = tc_ms.get("conv", 0.0) + tc_ms.get("dgrad", 0.0)
        tflops = gflop / tc_total                                         # GFLOP / ms == TFLOP/s
        peak = pk["bf16_tflops_sustained"]
        wg_gflop = wl["fwd_gflop"] * BS
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")      # made by tools/conv_traffic.py from an ncu pass
        if wl["ver"] == "yolov4" and os.path.exists(tp):
            tj = json.load(open(tp))
            traffic, traffic_src = tj["dram_bytes_per_launch"], \
                f"profiles/r01_conv_traffic.json: mean DRAM bytes over {tj['launches']} launches (one step, cold cache); " \
                f"algorithmic {tj['algorithmic_bytes_per_launch']:.3g} B/launch"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": BS, "global_batch": BS * world,
                       "targets_per_img": PER_IMG,
                       "parallelism": f"dp{world}" + (" (one NCCL all-reduce of the flat fp32 gradients per step)"
                                                      if world > 1 else ""),
                       "optimizer": "SGD lr .01 momentum .937 nesterov (train.py:156)",
                       "fwd_loss_img_s": world * BS * fl_steps / ms_fl * 1e3,
                       "fwd_loss_ms_per_step": ms_fl / fl_steps,
                       "l2": "inputs (246 MB images + multi-GB activations per step) exceed the 126 MB L2"},
            "e2e": {"value": world * BS * args.steps / ms_e2e * 1e3, "unit": UNIT,
                    "h2d_bytes_per_step": int(host_imgs[0].numel() * 4 + host_tg[0].numel() * 4),
                    "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "conv_fwd_kernel (tcgen05 implicit GEMM: forward + dgrad launches)",
                         "bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": tflops / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": f"{pk_src} bf16_tflops_sustained",
                         "ms_per_step": {k: round(v, 3) for k, v in tc_ms.items()},
                         "algorithmic_gflop_per_step": gflop,
                         "wgrad": {"kernel": "conv_wgrad_kernel (side stream, overlapped)",
                                   "algorithmic_gflop_per_step": wg_gflop,
                                   "achieved": wg_gflop / max(tc_ms.get("wgrad", 0.0), 1e-9)}},
        }
        if world == 1 and not args.no_cpu:
            v, cores, sample = cpu_arm(wl, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="fwd_loss", choices=["fwd_loss", "train_v4", "train_v7"],
                    help="fwd_loss = BASELINE configs[1] (default, the N=1 headline); train_v7 = configs[2] "
                         "(yolov7/csl/nc16 full train step with the gradient all-reduce)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == "__main__":
    main()
