#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W                    # our arm (under torchrun for N > 1), cfg 3, weak scaling
    python bench.py --scaling strong --gpus N ...                    # cfg 4: fixed GLOBAL batch 16 (16/N volumes per GPU)
    python bench.py --config cfg5 --gpus N ...                       # cfg 5: wide model [64..1024], 1 x 4 x 160x192x160 per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU implementation of the same path

Default workload (BASELINE.json configs[2], "cfg 3"): one training step = forward + deep-supervised Dice/focal/boundary loss +
backward (+ gradient all-reduce for N > 1) + AdamW step of the default-architecture enhanced 3D U-Net on a synthetic
batch of 2 volumes per GPU, 4 x 128^3, bf16 compute / fp32 parameters.  value = voxels of all ranks * K / time  [voxels/s].
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "3D U-Net train voxels/s (4ch 128^3) at 1/2/4/8 B200; inference volumes/s"
# conv FLOPs per input voxel (fprop / fprop+dgrad+wgrad), BASELINE.md §3 — analytic = hook-counted on the reference
CONFIGS = {
    "cfg3": {"features": [32, 64, 128, 256, 512], "per_gpu_batch": 2, "size": (128, 128, 128), "fwd_flop": 513215.0,
             "train_flop": 1532476.0,
             "workload": "cfg3: train step, batch 2/GPU, 4x128^3, default arch [32,64,128,256,512], dropout 0.2, "
                         "DeepSupervisionLoss3D(CombinedLoss3D), step = fwd+loss+bwd(+NCCL grad all-reduce)+AdamW"},
    "cfg4": {"features": [32, 64, 128, 256, 512], "global_batch": 16, "size": (128, 128, 128), "fwd_flop": 513215.0,
             "train_flop": 1532476.0,
             "workload": "cfg4: train step, GLOBAL batch 16 (16/N per GPU), 4x128^3, default arch [32,64,128,256,512], dropout "
                         "0.2, DeepSupervisionLoss3D(CombinedLoss3D), step = fwd+loss+bwd(+NCCL grad all-reduce)+AdamW"},
    "cfg5": {"features": [64, 128, 256, 512, 1024], "per_gpu_batch": 1, "size": (160, 192, 160), "fwd_flop": 2037501.0,
             "train_flop": 6098168.0,
             "workload": "cfg5: train step, batch 1/GPU, 4x160x192x160, wide arch [64,128,256,512,1024] (config.py:139), dropout "
                         "0.2, DeepSupervisionLoss3D(CombinedLoss3D), step = fwd+loss+bwd(+NCCL grad all-reduce)+AdamW"},
}
L2_NOTE = "no explicit flush: one step streams >5 GB of activations (>> 126 MB L2)"


def resolve_config(args, world):
    """-> (name, cfg dict with per_gpu_batch resolved, `config` object of the JSON line — identical for both arms)."""
    name = args.config
    if args.scaling == "strong":
        name = "cfg4"
    cfg = dict(CONFIGS[name])
    if "global_batch" in cfg:
        if cfg["global_batch"] % world:
            raise SystemExit("global batch %d does not divide over %d GPUs" % (cfg["global_batch"], world))
        cfg["per_gpu_batch"] = cfg["global_batch"] // world
    line_cfg = {"workload": cfg["workload"], "global_batch": world * cfg["per_gpu_batch"], "parallelism": "dp%d" % world,
                "l2": L2_NOTE}
    return name, cfg, line_cfg


def _peaks():
    p = {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            m = json.load(fh)
        p.update({k: m[k] for k in ("bf16_tflops_sustained", "bf16_tflops", "hbm_gbs") if k in m})
        p["src"] = "measured"
    except Exception:
        pass
    return p


def _ncu_traffic():
    """Per-kernel DRAM traffic of THIS build's dominant kernels: profiles/ncu_traffic.json, regenerated from an
    `ncu --set full` capture by scripts/ncu_traffic.py (dram__bytes_read.sum + dram__bytes_write.sum per launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Polls NVML in-process every
    10 ms (a 5-step timed region lasts ~0.1 s, shorter than nvidia-smi's start-up); falls back to `nvidia-smi -lms`."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def run(self):
        try:
            nv, h = self._nvml_handle()
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                act = lambda bit: "Active" if (rs & bit) else "Not Active"
                self.rows.append([str(sm), str(mx), "0", act(0x8), act(0x40), act(0x20), act(0x4)])
                time.sleep(0.01)
            return
        except Exception:
            pass
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                  "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.proc = p
        for line in p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            if self.stop_flag:
                break
        try:
            p.kill()
        except Exception:
            pass

    def summary(self):
        self.stop_flag = True
        time.sleep(0.05)
        try:
            if self.proc is not None:
                self.proc.kill()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}




# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's OWN PyTorch classes on the host cores — sliced live from /root/reference in
# the build container, or from oracle/_ref/ (materialised by oracle/make_ref.py, shipped with the snapshot) on the GPU box;
# the pinned oracle restatement ("port") only if neither exists.  The only places outside tests/ allowed to execute oracle/.
# ------------------------------------------------------------------------------------------------------------------
def cpu_train_steps(cfg, steps, warmup):
    """Times `steps` training steps of the reference implementation on a BOUNDED sample of the configured workload: ONE
    volume of the per-GPU batch at the configured resolution for the default architecture (1 x 4 x 128^3 = the reference's
    own 128^3 CPU probe of BASELINE.md §2), a 64^3 crop for the wide model.  Throughput is per voxel, so the sample rate is
    directly comparable with the GPU arm's voxels/s."""
    from oracle import ref_slice
    from oracle import unet3d_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    feats = list(cfg["features"])
    size = tuple(cfg["size"]) if feats[0] <= 32 else (64, 64, 64)
    if os.environ.get("B3D_CPU_SAMPLE_SIZE"):
        size = (int(os.environ["B3D_CPU_SAMPLE_SIZE"]),) * 3
    batch = 1
    x, y = O.make_inputs(batch, size[0], size[1], size[2], seed=0)
    if ref_slice.available():
        ns = ref_slice.load()
        torch.manual_seed(0)
        model = ns["UNet3D"](4, 4, features=feats)
        model.train()
        crit = ns["DeepSupervisionLoss3D"]()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
        kind = "reference"

        def step():
            opt.zero_grad()
            loss = crit(model(x), y)
            loss.backward()
            opt.step()
            return float(loss.detach())
    else:
        sd = O.make_state_dict(4, 4, feats, seed=0)
        params = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
        sd.update(params)
        masks = O.make_dropout_masks(batch, feats, 0.2)
        opt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=1e-4)
        kind = "port"

        def step():
            opt.zero_grad()
            main, deep, _ = O.unet_forward(x, sd, feats, training=True, dropout_masks=masks)
            loss = O.deep_supervision_loss(main, deep, y)
            loss.backward()
            opt.step()
            return float(loss.detach())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    vox = batch * size[0] * size[1] * size[2]
    full = tuple(cfg["size"])
    what = "one volume of the per-GPU batch at full resolution" if size == full else "a %dx%dx%d crop of one volume" % size
    return {"value": vox / dt, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": "%s (%d x 4 x %dx%dx%d), same step: fwd + DS loss + bwd + AdamW, fp32, %s classes on the host CPU, "
                      "%d timed steps after %d warm-up" % (what, batch, size[0], size[1], size[2],
                                                           "the reference's own" if kind == "reference" else "oracle-port",
                                                           steps, warmup),
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(args.gpus, 1)
    name, cfg, line_cfg = resolve_config(args, world)
    cb = cpu_train_steps(cfg, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "voxels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": line_cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# Data parallel: SMs left to NCCL's channel CTAs (and NCCL_MAX_CTAS capped to the same number) so an all-reduce running beside
# backward does not stall the persistent one-CTA-per-SM conv grids; 0 = off.  See DESIGN.md section 6.
DP_RESERVED_SMS = "0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-families", action="store_true", help="skip the eager per-kernel-family pass")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import ctypes
    import torch.distributed as dist
    import b3d  # noqa: F401
    import unet3d_b200 as U
    from unet3d_b200 import _lib, ops
    from unet3d_b200.parallel import DataParallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name, cfg, line_cfg = resolve_config(args, world)
    feats, nb, (sd_, sh_, sw_) = list(cfg["features"]), cfg["per_gpu_batch"], cfg["size"]
    vol = sd_ * sh_ * sw_
    # exactly ONE line on stdout: everything else that writes to fd 1 (NCCL prints a version banner there) goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dp_reserved = int(os.environ.get("B3D_DP_RESERVED_SMS", DP_RESERVED_SMS)) if world > 1 else 0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if dp_reserved > 0:   # NCCL's channel CTAs get their own SMs (read by NCCL when the communicator is created)
            os.environ.setdefault("NCCL_MAX_CTAS", str(dp_reserved))
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    lib.b3d_launch_count.restype = ctypes.c_longlong

    torch.manual_seed(0)
    model = U.UNet3D(4, 4, features=feats, dropout_rate=0.2).to(dev)
    model.train()
    crit = U.DeepSupervisionLoss3D()
    net = DataParallel(model, bucket_mb=float(os.environ.get("B3D_BUCKET_MB", "32")), reserved_sms=dp_reserved) if world > 1 else model
    use_graph = not os.environ.get("B3D_NO_GRAPH")
    opt = U.make_adamw(model, lr=1e-4, weight_decay=1e-4, capturable=use_graph)

    g = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.randn(nb, 4, sd_, sh_, sw_, generator=g).pin_memory()
    y_host = torch.randint(0, 4, (nb, sd_, sh_, sw_), generator=g).pin_memory()
    xd, yd = x_host.to(dev), y_host.to(dev)

    def eager_step(xi, yi):
        opt.zero_grad(set_to_none=True)
        out = net(xi)
        loss = crit(out, yi)
        loss.backward()
        opt.step()
        return loss

    step, graphed = eager_step, False
    if use_graph:  # the whole step (fwd + loss + bwd + all-reduce + AdamW) as ONE CUDA graph, replayed per step
        try:
            step = U.GraphedTrainStep(net, crit, opt, xd, yd, warmup=3)
            graphed = True
        except Exception as e:  # noqa: BLE001 - fall back to the eager (still all-CUDA) step
            sys.stderr.write("[bench] CUDA-graph capture failed (%s: %s); running the eager step\n" % (type(e).__name__, e))
            step, graphed = eager_step, False
        if world > 1:   # all ranks must run the same variant (a graphed rank and an eager rank would mismatch collectives)
            flag = torch.tensor([1 if graphed else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                step, graphed = eager_step, False

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step(xd, yd)
    barrier()

    # ---- timed region: device-resident inputs ---------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.rows.clear()   # keep only samples taken under load
    e0.record()
    for _ in range(args.steps):
        step(xd, yd)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.summary() if sampler else None
    vox_per_step = world * nb * vol
    value = vox_per_step * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host inputs, H2D inside the timed region, D2H of the loss ---------------------------------
    # (diagnostic, outside every timed region: what the box's host->device link delivers for this batch from pinned memory —
    # when it is below batch bytes / step time, e2e is bound by the link, not by the GPU work)
    h2d_probe = torch.empty_like(xd)
    barrier()
    e0.record()
    for _ in range(3):
        h2d_probe.copy_(x_host, non_blocking=True)
    e1.record()
    barrier()
    h2d_gbs = 3 * x_host.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del h2d_probe
    if graphed:   # one untimed pass through the pipelined path (its first use sets up the copy stream / pinned-copy machinery)
        step.prefetch(x_host, y_host)
        step.step_prefetched().item()
    barrier()
    e0.record()
    last = 0.0
    if graphed:
        step.prefetch(x_host, y_host)   # batch 0's H2D is inside the timed region
    for it in range(args.steps):
        if graphed:   # every step: H2D of the next batch on a copy stream (pinned host -> staging), replay, loss D2H
            loss_t = step.step_prefetched()
            if it + 1 < args.steps:
                step.prefetch(x_host, y_host)
            last = loss_t.item()
        else:
            xi = x_host.to(dev, non_blocking=True)
            yi = y_host.to(dev, non_blocking=True)
            last = step(xi, yi).item()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": vox_per_step * args.steps / (ms_e2e * 1e-3), "unit": "voxels/s",
           "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e2e / args.steps, "last_loss": last, "h2d_link_gbs_measured": h2d_gbs,
           "pipeline": "H2D of batch i+1 (pinned -> staging, copy stream) overlaps the graph replay of batch i; loss .item() per step"}

    # ---- per-kernel-family CUDA events + launch count: an eager pass of the SAME step after the timed regions (a replayed
    # CUDA graph has no host-side hooks between its kernels); it is not part of `value`.  Every kernel runs alone on one
    # stream here (the weight-gradient side stream is off), so the events around a launch measure that kernel only.
    peaks = _peaks()
    roof, fams, bw_fams, launches, ms_prof = None, {}, {}, 0, None
    if not args.no_families:
        side_saved = (ops.WGRAD_SIDE, ops.WGRAD_STREAM)
        ops.WGRAD_SIDE, ops.WGRAD_STREAM = False, None
        if graphed:
            for _ in range(2):   # the eager allocator pool is cold after the capture: warm it before timing the eager pass
                eager_step(xd, yd)
        ops.PROFILE = []
        ops.PROFILE_BW = []
        l0 = lib.b3d_launch_count()
        barrier()
        # The device must run BEHIND the host here, otherwise the event pair around a small launch also times the host's
        # gap between "record" and "launch" (an eager step costs ~20 ms of host time for ~17 ms of device time, and the family
        # figures used to move by 10 % with the box's host speed).  A ~30 ms spin kernel in front of every step lets the host
        # enqueue the step while the device waits; launches and events then execute back to back.
        spin = int(0.03 * 1.9e9)
        pairs = []
        for _ in range(args.steps):
            torch.cuda._sleep(spin)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            eager_step(xd, yd)
            s1.record()
            pairs.append((s0, s1))
        barrier()
        ms_prof = sum(a.elapsed_time(b) for a, b in pairs)
        launches = lib.b3d_launch_count() - l0
        prof, ops.PROFILE = ops.PROFILE, None
        prof_bw, ops.PROFILE_BW = ops.PROFILE_BW, None
        ops.WGRAD_SIDE, ops.WGRAD_STREAM = side_saved
        fam = {}
        for name, flops, a, b, tag in prof:
            # families by kernel: conv3 = 3x3x3 fprop/dgrad (conv_zs.cu at levels 0-1, conv_igemm.cu below), pointwise = 1x1x1
            # convs + ConvTranspose, wgrad3 = 3x3x3 weight gradients (conv_wg2.cu / conv_wgrad.cu), wgrad_pw
            kind = tag.split(" ")[0] if tag else name
            name = {"conv3": "conv3", "conv1": "pointwise", "convT": "pointwise", "convT_dgrad": "pointwise",
                    "wgrad3": "wgrad3", "wgrad1": "wgrad_pw", "wgradT": "wgrad_pw"}.get(kind, name)
            d = fam.setdefault(name, [0.0, 0.0, 0])
            d[0] += flops; d[1] += a.elapsed_time(b); d[2] += 1
        for name, (fl, tms, cnt) in fam.items():
            fams[name] = {"tflops": fl / (tms * 1e-3) / 1e12 if tms > 0 else None, "ms_per_step": tms / args.steps,
                          "launches_per_step": cnt / args.steps, "gflop_per_launch": fl / max(cnt, 1) / 1e9}
        # bandwidth-bound families: compulsory bytes (every input once + every output once) / CUDA-event time
        bw = {}
        for name, nbytes, a, b, tag in prof_bw:
            if name == "wgrad3_bytes":
                continue
            dd = bw.setdefault(name, [0.0, 0.0, 0])
            dd[0] += nbytes; dd[1] += a.elapsed_time(b); dd[2] += 1
        for name, (nb_, tms, cnt) in bw.items():
            gbs = nb_ / (tms * 1e-3) / 1e9 if tms > 0 else None
            bw_fams[name] = {"gbs": gbs, "frac_of_hbm_peak": (gbs / peaks["hbm_gbs"]) if gbs else None,
                             "ms_per_step": tms / args.steps, "launches_per_step": cnt / args.steps,
                             "compulsory_mb_per_step": nb_ / args.steps / 1e6}
        if "conv3" in fam:
            fl, tms, cnt = fam["conv3"]
            ach = fl / (tms * 1e-3) / 1e12
            tr = _ncu_traffic() or {}
            rep = (tr.get("kernels") or {}).get(tr.get("roofline_kernel", ""), None)
            roof = {"kernel": "zs_kernel / igemm_kernel (3x3x3 conv fprop + dgrad, tcgen05 implicit GEMM)", "bound": "tensor",
                    "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"],
                    # dram__bytes_read+write per launch of the family's largest launch, from the committed ncu capture of this
                    # kernel set (profiles/ncu_traffic.json <- scripts/ncu_traffic.py); null when no capture is committed
                    "traffic": rep["dram_bytes"] if rep else None,
                    "traffic_kernel": ("%s: %.1f MB DRAM vs %.1f MB algorithmic (%s)" % (
                        tr["roofline_kernel"], rep["dram_bytes"] / 1e6, rep["algorithmic_bytes"] / 1e6, tr.get("source", "")))
                    if rep else None,
                    "peak_source": peaks["src"] + " (sustained bf16, kernel timed inside a long step)",
                    "avg_launch_ms": tms / max(cnt, 1), "algorithmic_gflop_per_launch": fl / max(cnt, 1) / 1e9,
                    "flops": "2*voxels*Cin*Cout*27 with the REAL Cin/Cout of every launch (zero-padded channels not counted)",
                    "share_of_step": tms / ms_prof,
                    "measured_in": "eager single-stream pass of the same step inside bench.py (CUDA events around every "
                                   "launch, device kept behind the host by a spin kernel), %.2f ms/step" % (ms_prof / args.steps)}

    # ---- inference (cfg 2): batch 1, eval mode ---------------------------------------------------------------------
    inference = None
    if not args.no_inference:
        model.eval()
        x1 = xd[:1].contiguous()
        infer, inf_graphed = model, False
        if use_graph:
            try:
                infer, inf_graphed = U.GraphedInference(model, x1), True
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("[bench] inference graph capture failed (%s); eager forward\n" % e)
                infer = model
        with torch.no_grad():
            for _ in range(3):
                infer(x1)
            barrier()
            e0.record()
            for _ in range(args.steps):
                infer(x1)
            e1.record()
            barrier()
        ims = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        inference = {"value": world * 1e3 / ims, "unit": "volumes/s", "ms_per_volume": ims,
                     "config": "cfg2: batch 1 per GPU, 4x%dx%dx%d, eval, bf16%s" % (sd_, sh_, sw_, ", CUDA graph" if inf_graphed else ""),
                     "conv_tflops": cfg["fwd_flop"] * vol / (ims * 1e-3) / 1e12}
        if inf_graphed:
            infer.close()
        model.train()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_train_steps(cfg, 2, 1)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": line_cfg, "cuda_graph": graphed,
            "optimizer": type(opt).__name__,
            "conv_tflops_whole_step": cfg["train_flop"] * vox_per_step / world / (ms / args.steps * 1e-3) / 1e12,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernel_families": fams,
            "bandwidth_families": bw_fams, "inference": inference, "cpu_baseline": cpu_baseline,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        barrier()
        if graphed:   # a captured graph holds NCCL work: release it before the communicator goes away
            step.close()
        barrier()
        done = threading.Event()

        def _teardown():
            dist.destroy_process_group()
            done.set()
        th = threading.Thread(target=_teardown, daemon=True)
        th.start()
        if not done.wait(30.0):   # never hang the launcher on communicator teardown
            sys.stderr.write("[bench] rank %d: destroy_process_group did not return in 30 s; exiting\n" % rank)
            sys.stderr.flush()
            os._exit(0)


if __name__ == "__main__":
    main()
