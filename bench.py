"""bench.py -- MU iterations/sec of the dense NMF hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--config cfg2|cfg1|cfg4s|cfg5]

A "step" is one `fit(V, beta, tol=-inf, max_iter=ITERS)` pass (the reference's own benchmark protocol,
examples/benchmarks/benchmark.ipynb cell 4: loss evaluations every 10 iterations included) on one
synthetic batch: V = rand(N, C) rounded to bf16-representable values, W0/H0 = |randn|.

  value : ITERS * K * n_gpus / t   with V, W, H resident in HBM (one 65536x4096 shard per GPU)
  e2e   : the same through the public API with HOST (pinned) tensors: the module and V live on the CPU,
          `fit` stages V/W/H through the GPU and copies the factors back, all inside the timed region
  roofline / cpu_baseline : see DESIGN.md section "Measurement"

Multi-GPU (torchrun): rows are sharded, W is replicated, one all-reduce per W update; weak scaling
(every rank owns a full cfg2-sized shard), value = shard-iterations of all ranks per second.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "pytorch-nmf_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

CONFIGS = {
    # name: (N, C, R, beta, description)
    "cfg1": (256, 512, 16, 2.0, "NMF 256x512 rank=16 beta=2 (BASELINE.json configs[0])"),
    "cfg2": (65536, 4096, 64, 1.0, "NMF 4096x65536 (fed as V^T: 65536x4096) rank=64 beta=1 KL (BASELINE.json configs[1])"),
    "cfg4s": (131072, 8192, 128, 1.0, "one 1/8 row shard of NMF 8192x2^20 rank=128 beta=1 (BASELINE.json configs[3])"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc=1590.0, tc_sustained=1400.0, src="fallback")


def ncu_traffic(precision, cfg_name):
    """dram read+write bytes per launch of the fused contraction kernel from the committed ncu capture
    (profiles/r1_ncu_tc_contract_<precision>.txt, taken at cfg2); None when no capture matches."""
    if cfg_name != "cfg2":
        return None
    path = os.path.join(ROOT, "profiles", f"r1_ncu_tc_contract_{precision}.txt")
    try:
        tot = 0.0
        for ln in open(path):
            ln = ln.strip()
            if ln.startswith("dram__bytes_read.sum") or ln.startswith("dram__bytes_write.sum"):
                val, unit = ln.split("=")[1].split()[:2]
                tot += float(val) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
        return tot or None
    except Exception:
        return None


def make_inputs(N, C, R, seed):
    torch.manual_seed(seed)
    V = torch.rand(N, C).bfloat16().float()
    torch.manual_seed(seed + 1)
    W0 = torch.randn(C, R).abs()
    H0 = torch.randn(N, R).abs()
    return V, W0, H0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        top = sm[len(sm) // 2:] if sm else []        # samples under load = upper half
        med = top[len(top) // 2] if top else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def timed_steps(fn, steps, warmup, world):
    """W warm-up steps, then exactly K timed steps bracketed by barrier + synchronize; device time, max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[-1])
    timed_steps.per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def reference_module():
    """The unmodified reference (baseline/_ref, pip-installed from /root/reference) or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "torchnmf")):
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import torchnmf.nmf as rn
        return rn
    return None


def cpu_reference_rate(N, C, R, beta, iters, V=None, W0=None, H0=None, warm=True):
    """it/s of the reference's CPU path (fit(tol=-inf, max_iter=iters)) on this box's host cores."""
    torch.set_num_threads(os.cpu_count())
    try:
        torch.set_flush_denormal(True)       # README.md:101-102 of the reference
    except Exception:
        pass
    if V is None:
        V, W0, H0 = make_inputs(N, C, R, 0)
    rn = reference_module()
    if rn is not None:
        kind = "reference"
        if warm:
            rn.NMF(W=W0, H=H0).fit(V, beta, float("-inf"), 1)
        m = rn.NMF(W=W0, H=H0)
        t0 = time.perf_counter()
        n = m.fit(V, beta, float("-inf"), iters)
        dt = time.perf_counter() - t0
    else:
        from oracle import mu_oracle as orc
        kind = "port"
        if warm:
            orc.fit(V, W0, H0, beta=beta, tol=float("-inf"), max_iter=1)
        t0 = time.perf_counter()
        _, _, n, _ = orc.fit(V, W0, H0, beta=beta, tol=float("-inf"), max_iter=iters)
        dt = time.perf_counter() - t0
    return n / dt, kind, torch.get_num_threads(), dt


def run_reference_arm(a, cfg):
    N, C, R, beta, desc = cfg
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    V, W0, H0 = make_inputs(N, C, R, 0)
    it = a.ref_iters
    for _ in range(max(1, a.warmup)):
        cpu_reference_rate(N, C, R, beta, 1, V, W0, H0, warm=False)
    t0 = time.perf_counter()
    kind, cores = "port", 1
    for _ in range(a.steps):
        _, kind, cores, _ = cpu_reference_rate(N, C, R, beta, it, V, W0, H0, warm=False)
    dt = time.perf_counter() - t0
    rate = a.steps * it / dt
    line = {
        "impl": "reference", "metric": "MU iterations/sec", "value": rate, "unit": "iter/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "N": N, "C": C, "R": R, "beta": beta, "iters_per_step": it,
                   "note": "reference torchnmf CPU path (fit incl. init loss), bounded sample of the same workload"},
        "cpu_baseline": {"value": rate, "unit": "iter/s", "cores": cores, "kind": kind,
                         "sample": f"{a.steps} x fit(tol=-inf, max_iter={it}) on the full {N}x{C} R={R} target"},
        "e2e": {"value": rate, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def _capture_stdout():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the one JSON line is written
    to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--iters", type=int, default=200, help="MU iterations per step (fit max_iter)")
    ap.add_argument("--ref-iters", type=int, default=2, help="MU iterations per reference step")
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    N, C, R, beta, desc = cfg
    if a.impl == "reference":
        run_reference_arm(a, cfg)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    rank, world, local = dist_setup(a.gpus)
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    from torchnmf_b200 import NMF, _capi
    from torchnmf_b200.engine import CudaNmfEngine
    peaks = load_peaks()
    group = dist.group.WORLD if world > 1 else None

    # ---------------- inputs: one full shard per rank (weak scaling), bigger than L2 -----------------
    V_cpu, W0, H0 = make_inputs(N, C, R, 2 * rank)
    if world > 1:                       # W replicas must start identical
        torch.manual_seed(1)
        W0 = torch.randn(C, R).abs()
    V_cpu = V_cpu.pin_memory()
    V_dev = V_cpu.to(dev)
    model = NMF(W=W0, H=H0).to(dev)
    W0d, H0d = W0.to(dev), H0.to(dev)

    def step_resident():
        model.W.data.copy_(W0d); model.H.data.copy_(H0d)
        n = model.fit(V_dev, beta, float("-inf"), a.iters, precision=a.precision, group=group)
        assert n == a.iters

    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    # warm-up outside the sampler, then sample clocks during the timed steps only
    for _ in range(a.warmup):
        step_resident()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = _capi.launch_count()
    ms = timed_steps(step_resident, a.steps, 0, world)
    launches = _capi.launch_count() - l0
    per_step = list(getattr(timed_steps, "per_step", []))
    clocks = sampler.stop() if sampler else None
    precision = model.last_fit_precision
    value = a.iters * a.steps * world / (ms * 1e-3)

    # ---------------- end to end through the public API with host buffers ------------------------------
    e2e = None
    if not a.no_e2e:
        host_model = NMF(W=W0, H=H0)          # CPU-resident module, like the reference's default
        for p in (host_model.W, host_model.H):
            p.data = p.data.pin_memory()
        W0p, H0p = W0.pin_memory(), H0.pin_memory()

        def step_host():
            host_model.W.data.copy_(W0p); host_model.H.data.copy_(H0p)
            n = host_model.fit(V_cpu, beta, float("-inf"), a.iters, precision=a.precision, group=group)
            assert n == a.iters

        for _ in range(max(1, a.warmup // 2)):
            step_host()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        n_loss = 1 + a.iters // 10
        e2e = {"value": a.iters * a.steps * world / dt, "unit": "iter/s",
               "h2d_bytes_per_step": int(V_cpu.nbytes + W0.nbytes + H0.nbytes),
               "d2h_bytes_per_step": int(W0.nbytes + H0.nbytes + 8 * n_loss + 8),
               "ms_per_step": 1e3 * dt / a.steps}

    # ---------------- roofline of the dominant kernel (fused contraction), timed alone with CUDA events ---
    roof = None
    eng = CudaNmfEngine(V_dev, model.W.data, model.H.data, a.precision)
    reps = 10
    times = {}
    for which, nm in ((0, "w"), (1, "h")):
        for _ in range(3):
            eng.contract_only(which, beta)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.contract_only(which, beta)
        e1.record()
        torch.cuda.synchronize()
        times[nm] = e0.elapsed_time(e1) / reps * 1e-3
    eng.close()
    flops_launch = 4.0 * N * C * R if beta == 1 else (2.0 * N * C * R if beta == 2 else 6.0 * N * C * R)
    v_bytes = N * C * (4 if precision == "f32" else 2)
    t_dom = max(times.values())
    which_dom = max(times, key=times.get)
    ach_tf = flops_launch / t_dom / 1e12
    ach_gb = v_bytes / t_dom / 1e9
    tensor = {"achieved": ach_tf, "peak": peaks["tc"], "unit": "TFLOP/s", "frac": ach_tf / peaks["tc"],
              "algorithmic_flops_per_launch": flops_launch,
              "peak_source": f"{peaks['src']} MEASURED_PEAKS.json bf16 burst (f16 has the same tensor peak)"}
    hbm = {"achieved": ach_gb, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach_gb / peaks["hbm"],
           "algorithmic_bytes_per_launch": v_bytes, "peak_source": f"{peaks['src']} MEASURED_PEAKS.json copy bandwidth (burst)"}
    # the binding roofline is the one whose minimum time for this launch is larger (cfg2, R=64: HBM 82 us vs tensor 41 us;
    # at R=128 the two meet); the other one is reported next to it
    hbm_bound = v_bytes / (peaks["hbm"] * 1e9) >= flops_launch / (peaks["tc"] * 1e12) or precision == "f32"
    roof = dict(hbm if hbm_bound else tensor)
    roof.update({"bound": "hbm" if hbm_bound else "tensor", "traffic": ncu_traffic(precision, a.config),
                 "kernel": f"fused {which_dom}-update contraction ({precision})", "kernel_ms": t_dom * 1e3,
                 "kernel_ms_w": times["w"] * 1e3, "kernel_ms_h": times["h"] * 1e3,
                 "tensor" if hbm_bound else "hbm": tensor if hbm_bound else hbm,
                 "step_tensor_frac": (8.0 * N * C * R * a.iters * a.steps / (ms * 1e-3) / 1e12) / peaks["tc_sustained"]
                 if beta == 1 else None})

    # ---------------- CPU baseline (rank 0, N=1 only): the reference's own CPU path on this box ---------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        k = 5 if N * C >= 1 << 26 else 50
        rate, kind, cores, dt = cpu_reference_rate(N, C, R, beta, k, V_cpu, W0, H0)
        cpu = {"value": rate, "unit": "iter/s", "cores": cores, "kind": kind,
               "sample": f"fit(tol=-inf, max_iter={k}) after a 1-iteration warm-up on the full {N}x{C} R={R} target ({dt:.1f}s)"}

    if rank == 0:
        line = {
            "metric": "MU iterations/sec", "value": value, "unit": "iter/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "step_ms": per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"f32": "f32", "f16": "f16", "f16_split": "f16"}.get(precision, precision),
            "data": "synthetic",
            "config": {"workload": desc, "N_per_gpu": N, "C": C, "R": R, "beta": beta, "iters_per_step": a.iters,
                       "precision": precision, "l2": "inputs larger than L2 (V shard >= 512 MiB)" if N * C * 2 > 126e6
                       else "inputs fit in L2", "unit_def": "one MU iteration (W then H update) over one N_per_gpu x C shard; "
                       "loss every 10th iteration included", "parallelism": f"row-shard x{world}" if world > 1 else "single"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
