#!/usr/bin/env python
"""bench.py — DLSA rows/sec on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one whole pass of the hot path over the synthetic workload: every Newton pass (fused K1 kernel +
batched Cholesky update; the last pass's H is Sig_inv at the returned coefficients), the one-round combine (+ one NCCL allreduce when N>1) and the
LARS path with AIC/BIC selection.  Workload at N=1 = BASELINE.json configs[1]: simulated logistic regression
n=1e8, p=50, 1024 partitions (40.8 GB of float64 in HBM).  N>1 is weak scaling: every rank holds its own
1e8-row / 1024-partition shard of the same counter-based synthetic stream (global row index keyed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows R] [--p P] [--parts K]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dlsa_rows_per_sec"
UNIT = "rows/s"
FP64_PEAK_FALLBACK_TFLOPS = 37.1  # profiles/r01_microbench_fp64.json (DMMA.8x8x4 == DFMA peak, measured on this pool)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per GPU")
    ap.add_argument("--p", type=int, default=50)
    ap.add_argument("--parts", type=int, default=1024, help="partitions per GPU")
    ap.add_argument("--seed", type=int, default=20210517)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-parts", type=int, default=64, help="partitions in the bounded CPU-baseline sample")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def fp64_peak():
    path = os.path.join(ROOT, "profiles", "r01_microbench_fp64.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return max(v for k, v in d.items() if k.startswith("dmma884_")), "measured (profiles/r01_microbench_fp64.json)"
    except Exception:
        return FP64_PEAK_FALLBACK_TFLOPS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": sorted(hi)[len(hi) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle/ref_port.py: sklearn newton-cg as shipped + pandas
    group-sum + lstsq + LARS) on all host cores, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_port
    n_k = args.rows // args.parts
    K = args.cpu_parts
    cores = os.cpu_count() or 1
    res = ref_port.run_reference_cpu(K=K, n_k=n_k, p=args.p, seed=args.seed, nproc=cores, reps=args.steps,
                                     warmup=args.warmup)
    secs = res["seconds"]
    mean_s = sum(secs) / len(secs)
    val = res["rows"] / mean_s
    sample = (f"{K} of the {args.parts} partitions ({n_k} rows x p={args.p} each) per step; sklearn newton-cg "
              f"(default tol, as shipped) + pandas group-sum + lstsq + LARS; {res['cores']} processes x 1 BLAS thread; "
              "Spark/JVM/Arrow costs excluded")
    import sklearn
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"simulated logistic n={args.rows:.0e}, p={args.p}, {args.parts} partitions (configs[1])",
                      "sample": sample, "sklearn": sklearn.__version__},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


_REAL_STDOUT = None


def quiet_stdout():
    """Library banners (e.g. "NCCL version ...") go to fd 1; keep stdout for the ONE JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the DLSA hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dlsa_b200 import _lib, engine, synth

    n, p, K = args.rows, args.p, args.parts
    row0 = rank * n
    # ---- synthetic shard, generated on the device (law of simulate_logistic, dlsa/models.py:6-40) ----------
    X = torch.empty(n, p, dtype=torch.float64, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    chunk = 10_000_000
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        synth.simulate_logistic_device(m, p, args.seed, dev, row0=row0 + s, out=(X[s:s + m], y[s:s + m]))
    off = synth.equal_partition_offsets(n, K, dev)
    n_total = float(n * world)
    torch.cuda.synchronize()

    def step():
        return engine.dlsa_fit(X, y, off, fit_intercept=False, sample_size=n_total, type="lar")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        fits, comb, path = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    engine.PASS_EVENTS = []
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        fits, comb, path = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    pass_events = engine.PASS_EVENTS
    engine.PASS_EVENTS = None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: the fused pass K1, timed live with CUDA events ------------------
    full = [a.elapsed_time(b) for a, b, is_full in pass_events if is_full]       # all partitions active
    allp = [a.elapsed_time(b) for a, b, _ in pass_events]
    pass_ms = sum(full) / max(len(full), 1)
    bytes_per_row = 8 * (p + 1)
    flops_per_row = p * p + 5 * p
    hbm_peak, hbm_src = measured_peaks()
    f64_peak, f64_src = fp64_peak()
    ach_gbs = n * bytes_per_row / (pass_ms * 1e-3) / 1e9
    ach_tf = n * flops_per_row / (pass_ms * 1e-3) / 1e12
    t_hbm = n * bytes_per_row / (hbm_peak * 1e9)
    t_f64 = n * flops_per_row / (f64_peak * 1e12)
    # The binding roof goes first.  FP64 DMMA/DFMA is the tensor-core path of this kernel (tcgen05 has no f64 kind); its
    # peak is measured on this pool by tools/microbench_fp64.cu (MEASURED_PEAKS.json has no FP64 entry).  The HBM view
    # (against the driver-measured copy bandwidth) is reported beside it.
    hbm_view = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "peak_source": hbm_src}
    f64_view = {"bound": "tensor", "achieved": ach_tf, "peak": f64_peak, "unit": "TFLOP/s", "frac": ach_tf / f64_peak,
                "peak_source": f64_src,
                "note": "FP64 DMMA.8x8x4 and DFMA share one 64 FMA/clk/SM pipe on B200 (37.1 TFLOP/s measured)"}
    binding = f64_view if t_f64 > t_hbm else hbm_view
    roofline = dict(binding)
    roofline.update({"traffic": None, "kernel": "irls_small_kernel (K1 fused Newton pass)" if p <= 56 else
                     "irls_tiled_kernel (K1 fused Newton pass)",
                     "kernel_ms": pass_ms, "kernel_share_of_step": sum(allp) / ms,
                     "algorithmic_bytes_per_row": bytes_per_row, "algorithmic_flops_per_row": flops_per_row,
                     "time_bound_ms": {"hbm": t_hbm * 1e3, "fp64": t_f64 * 1e3},
                     "binding_roof": "fp64" if t_f64 > t_hbm else "hbm",
                     "hbm": hbm_view, "fp64": f64_view})
    traffic_path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.isfile(traffic_path):
        try:
            with open(traffic_path) as f:
                tj = json.load(f)
            if int(tj.get("p", -1)) == p:  # DRAM bytes of one launch from the committed ncu --set full capture
                roofline["traffic"] = tj["dram_bytes_per_launch"] * (n / float(tj["rows"]))
                roofline["traffic_source"] = "profiles/k1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
        except Exception:
            pass

    # ---- e2e: host buffers in, results out, copies inside the timed region --------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            # every rank pins its whole shard (40.8 GB at C2): make sure the box has the host memory for all local
            # ranks before any of them asks for it (an out-of-memory kill would take the whole run down); the decision
            # is taken collectively so that no rank waits at a barrier the others skipped
            barrier()
            mem_ok, mem_msg = 1, ""
            try:
                import psutil
                need = world * (n * (p + 1) * 8) * 1.15
                avail = psutil.virtual_memory().available
                if avail < need:
                    mem_ok = 0
                    mem_msg = f"host memory: {avail / 1e9:.0f} GB available, {need / 1e9:.0f} GB needed to pin {world} shards"
            except ImportError:
                pass
            flag = torch.tensor([mem_ok], dtype=torch.int32, device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                raise MemoryError(mem_msg or "another rank reported too little host memory to pin its shard")
            Xh = torch.empty(n, p, dtype=torch.float64, pin_memory=True)
            yh = torch.empty(n, dtype=torch.float64, pin_memory=True)
            Xh.copy_(X)
            yh.copy_(y)
            offh = off.cpu()
            torch.cuda.synchronize()

            def e2e_step():
                f_, c_, p_ = engine.dlsa_fit_host(Xh, yh, offh, fit_intercept=False, sample_size=n_total, type="lar",
                                                  device=dev, buffers=(X, y))
                res = torch.cat([c_.beta_byOLS, c_.beta_byONESHOT, c_.Sig_inv_sum.view(-1), p_.beta[p_.argmin_BIC],
                                 p_.beta[p_.argmin_AIC], p_.BIC, p_.AIC]).cpu()
                return res

            e2e_step()
            barrier()
            e0.record()
            nsteps = max(1, min(args.steps, 3))
            for _ in range(nsteps):
                res = e2e_step()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item()) / nsteps
            e2e = {"value": n_total / (ems * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": int(Xh.numel() * 8 + yh.numel() * 8 + offh.numel() * 8),
                   "d2h_bytes_per_step": int(res.numel() * 8), "ms_per_step": ems, "steps": nsteps,
                   "api": "dlsa_b200.engine.dlsa_fit_host: pinned host X/y/offsets in, 16 partition blocks whose H2D copies "
                          "overlap the fits of earlier blocks, combined estimates + LARS path read back to the host"}
            del Xh, yh
        except Exception as ex:  # e.g. not enough host memory to pin the shard
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample ---------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import ref_port
            n_k = n // K
            res_cpu = ref_port.run_reference_cpu(K=args.cpu_parts, n_k=n_k, p=p, seed=args.seed, reps=1, warmup=0)
            cpu = {"value": res_cpu["rows"] / res_cpu["seconds"][0], "unit": UNIT, "cores": res_cpu["cores"],
                   "kind": "port",
                   "sample": f"{args.cpu_parts} of {K} partitions ({n_k} rows x p={p}); sklearn newton-cg as shipped "
                             "(default tol) + group-sum + lstsq + LARS; 1 process per core, 1 BLAS thread each"}
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": repr(ex)[:200]}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"simulated logistic n={n:.0e} per GPU, p={p}, {K} partitions per GPU (configs[1])",
                          "rows_total": int(n_total), "partitions_total": K * world, "parallelism": f"partitions x{world}",
                          "newton_passes_per_step": fits.passes, "tol": 1e-10,
                          "warm_start": {"pilot_partitions": fits.pilot_partitions, "pilot_passes": fits.pilot_passes,
                                         "cold_restart": fits.cold_restart,
                                         "note": "pilot block fitted from beta=0, its mean coefficients start the other "
                                                 "partitions (R init.beta); same MLE, same tol; every pass is inside the "
                                                 "timed region"},
                          "l2": "inputs (40.8 GB per pass) are far larger than the 126 MB L2; no flush needed",
                          "seed": args.seed},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "result_check": {"K_eff": comb.K_eff, "path_rows": int(path.beta.shape[0]),
                                "support_BIC": int((path.beta[path.argmin_BIC] != 0).sum().item()),
                                "theta_head": [round(v, 6) for v in comb.beta_byOLS[:3].tolist()]}}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
