#!/usr/bin/env python
"""bench.py -- Mamba-2 mixer prefill tokens/s (BASELINE.json metric) on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--seqlen L]

Workload (config.workload): ONE Nanov2-9B Mamba-2 mixer layer (H=128, P=80, G=8, N=128, Q=128, hidden 4480),
bf16 activations and parameters, batch 1, 131072 synthetic tokens, random-init weights by the reference recipe.
N > 1: the same 131072-token sequence sharded contiguously over N ranks (strong scaling) with one
boundary-state all-gather per layer (timeviper_b200/sharded.py).

One step = one pass of the hot path -- causal conv1d+SiLU -> SSD chunked scan (+D skip) -> z-gated grouped
RMSNorm -- over the whole sequence, input (the in_proj output) resident in HBM.  `value` = tokens / step time.
`e2e` = the same metric through the public API `Mamba2MixerPrefill.forward` (in_proj and out_proj on cuBLAS
included) with HOST buffers: pinned hidden_states -> H2D -> forward -> D2H of the output, all inside the
timed region.  `roofline` describes the dominant (slowest) of the three kernels, timed alone with CUDA events.
Inputs (5.9 GB) are far larger than L2 (126 MB), so no explicit L2 flush is needed between iterations.

`--impl reference` times the reference's own CPU implementation of the path (oracle/: the memory-lean
restatement of torch_forward, pinned to the reference by tests/golden) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mamba2_mixer_prefill_tokens_per_s"
UNIT = "tokens/s"
# algorithmic bytes per token (SURVEY.md 8d / DESIGN.md), bf16, 9B dims
BYTES_PER_TOKEN = {"conv1d": 49152, "ssd": 45312, "gated_rmsnorm": 61440}
# per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) from one `ncu --set full` capture of this
# command at seqlen 131072 on 1 GPU (profiles/r01_summary.md); null for any other configuration
NCU_TRAFFIC_128K = {"conv1d": 6.500e9, "ssd": 6.031e9 + 0.111e9, "gated_rmsnorm": 8.028e9}   # ssd = fused + dt/cumsum


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json (burst copy)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_tokens_per_s(sample_tokens, steps=1, warmup=0):
    """Reference CPU path (oracle port of torch_forward, kernel group mapping) on the host cores:
    conv -> SSD -> gated norm from a resident projected input, 9B dims, fp32."""
    from oracle import mamba2_ref as R
    from timeviper_b200.config import Mamba2Config
    cfg = Mamba2Config.nanov2_9b()
    H, P, G, N = cfg.mamba_num_heads, cfg.mamba_head_dim, cfg.n_groups, cfg.ssm_state_size
    torch.manual_seed(1234)
    p = R.nemotron_random_params(cfg.hidden_size, H, P, G, N, nondegenerate=False)
    L = sample_tokens
    proj = torch.randn(1, L, cfg.projection_size) * 0.5

    def step():
        gate, xBC, dt = proj.split([H * P, cfg.conv_dim, H], dim=-1)
        xc, _ = R.causal_conv1d_ref(xBC.transpose(1, 2), p["conv1d.weight"].squeeze(1), p["conv1d.bias"])
        x, Bm, Cm = xc.transpose(1, 2).split([H * P, G * N, G * N], dim=-1)
        y, s = R.ssd_chunked_ref(x.reshape(1, L, H, P), dt, -torch.exp(p["A_log"]), Bm.reshape(1, L, G, N),
                                 Cm.reshape(1, L, G, N), cfg.chunk_size, D=p["D"], dt_bias=p["dt_bias"],
                                 dt_softplus=True)
        return R.gated_rmsnorm_ref(y.reshape(1, L, H * P), p["norm.weight"], None, gate, 1e-5, H * P // G, False)

    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt_s = (time.perf_counter() - t0) / steps
    return L / dt_s, dt_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    sample = args.cpu_sample
    tps, sec = cpu_reference_tokens_per_s(sample, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": max(1, args.steps), "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Nanov2-9B Mamba-2 mixer layer, conv->SSD->gated-norm, batch 1",
                   "sample_tokens": sample, "note": "bounded sample of the 131072-token workload; CPU fp32"},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} tokens of the 9B-dims layer, oracle/mamba2_ref.py (torch_forward "
                                   f"restatement), {cores} torch threads, os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def time_region(fn, steps, dist_on):
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist_on = world > 1
    if dist_on:
        # The one collective on the critical path is a 5.24 MB/rank all-gather: at 4 ranks NCCL's LL128 protocol is ~20 %
        # faster than its default choice at that size (49 vs 62 us, tools/ag_sweep.sh).  Not measured in isolation at 8
        # ranks (in-step 113 us with LL128 vs 95-105 us with the default in an earlier run), so NCCL keeps its own choice
        # there; a user setting always wins.
        if world <= 4:
            os.environ.setdefault("NCCL_PROTO", "LL128")
        # high-priority NCCL stream: the halo all-gather overlaps the shard-wide conv instead of queueing behind it
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    import timeviper_b200 as tv
    from oracle import mamba2_ref as R     # parameter recipe only (nemotron_random_params); not on the timed path

    cfg = tv.Mamba2Config.nanov2_9b()
    Ltot = args.seqlen
    assert Ltot % (world * cfg.chunk_size) == 0, "seqlen must split into whole chunks per rank"
    L = Ltot // world
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)
    p = R.nemotron_random_params(cfg.hidden_size, cfg.mamba_num_heads, cfg.mamba_head_dim, cfg.n_groups,
                                 cfg.ssm_state_size, nondegenerate=False)
    mixer = tv.Mamba2MixerPrefill(cfg).to(torch.bfloat16).to(dev)
    mixer.load_state_dict({k: v.to(torch.bfloat16) for k, v in p.items()}, strict=True)
    mixer.eval()

    # synthetic tokens: the block RMS-normalises its input (modeling_nano.py:941) => unit-variance rows
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    hs_host = torch.empty(1, L, cfg.hidden_size, dtype=torch.bfloat16).pin_memory()
    hs_dev = torch.randn(1, L, cfg.hidden_size, device=dev, generator=g).to(torch.bfloat16)
    hs_host.copy_(hs_dev)
    out_host = torch.empty(1, L, cfg.hidden_size, dtype=torch.bfloat16).pin_memory()
    with torch.no_grad():
        proj = mixer.in_proj(hs_dev)                                    # resident input of the hot path
    family = tv.ssd_kernel_family(torch.bfloat16, cfg.mamba_head_dim, cfg.ssm_state_size, cfg.chunk_size)

    def core():
        with torch.no_grad():
            if dist_on:
                return tv.sharded_scan_core(mixer, proj)[0]
            return mixer.scan_core(proj)

    def e2e():
        with torch.no_grad():
            if dist_on:     # public host-buffer API of the sharded path: copies and compute on three streams
                _, done = tv.sharded_prefill_from_host(mixer, hs_host, out_host)
                torch.cuda.current_stream().wait_event(done)     # the timed region ends after the last D2H copy
            else:   # public host-buffer API: segments streamed over copy/compute/copy streams
                mixer.prefill_from_host(hs_host, out_host, segment_tokens=args.e2e_segment)

    for _ in range(max(3, args.warmup)):
        core()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t0 = time.time()
    ms = time_region(core, args.steps, dist_on)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # per-kernel timing (rank 0, local shard) for the roofline block: each kernel alone, back to back launches,
    # CUDA events on the launching stream.  The board runs into its power cap within ~0.2 s of back-to-back
    # launches of these kernels (SM clock 1965 -> ~1400 MHz, step time +40 %), so every section starts after a
    # short idle gap: the roofline peak is the burst copy bandwidth and is compared with burst kernel times.
    per = {}
    time.sleep(1.0)
    if rank == 0:
        with torch.no_grad():
            H, P, G, N = cfg.mamba_num_heads, cfg.mamba_head_dim, cfg.n_groups, cfg.ssm_state_size
            gate, xBC, dt = proj.split([H * P, cfg.conv_dim, H], dim=-1)
            w, b_ = mixer.conv1d.weight.squeeze(1), mixer.conv1d.bias
            xc = tv.causal_conv1d_fn(xBC.transpose(1, 2), w, b_, activation="silu").transpose(1, 2)
            x, Bm, Cm = torch.split(xc, [H * P, G * N, G * N], dim=-1)
            xv, Bv, Cv = x.view(1, L, H, P), Bm.view(1, L, G, N), Cm.view(1, L, G, N)
            A = -torch.exp(mixer.A_log.float())
            y = tv.mamba_chunk_scan_combined(xv, dt, A, Bv, Cv, cfg.chunk_size, D=mixer.D, dt_bias=mixer.dt_bias,
                                             dt_softplus=True, return_final_states=True)[0].view(1, L, -1)
            fns = {
                "conv1d": lambda: tv.causal_conv1d_fn(xBC.transpose(1, 2), w, b_, activation="silu"),
                "ssd": lambda: tv.mamba_chunk_scan_combined(xv, dt, A, Bv, Cv, cfg.chunk_size, D=mixer.D,
                                                            dt_bias=mixer.dt_bias, dt_softplus=True,
                                                            return_final_states=True),
                "gated_rmsnorm": lambda: tv.rmsnorm_fn(y, mixer.norm.weight, None, z=gate, eps=1e-5,
                                                       group_size=H * P // G, norm_before_gate=False),
            }
            for name, fn in fns.items():
                for _ in range(3):
                    fn()
                per[name] = time_region(fn, max(3, min(args.steps, 10)), False)
                time.sleep(0.5)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        e2e()
    ms_e2e = time_region(e2e, e2e_steps, dist_on)

    if dist_on:
        dist.barrier()

    if rank == 0:
        pk = peaks()
        dom = max(per, key=per.get)
        ach = BYTES_PER_TOKEN[dom] * L / (per[dom] * 1e-3) / 1e9
        kernels = {k: {"ms": v, "gbs": BYTES_PER_TOKEN[k] * L / (v * 1e-3) / 1e9,
                       "frac": BYTES_PER_TOKEN[k] * L / (v * 1e-3) / 1e9 / pk["hbm_gbs"]} for k, v in per.items()}
        path_bytes = sum(BYTES_PER_TOKEN.values())
        cpu_tps, cpu_sec = (None, None)
        cpu_block = None
        if world == 1 and not args.no_cpu_baseline:
            cores = torch.get_num_threads()
            cpu_tps, cpu_sec = cpu_reference_tokens_per_s(args.cpu_sample)
            cpu_block = {"value": cpu_tps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.cpu_sample} tokens of the same layer, oracle/mamba2_ref.py (restatement of "
                                   f"the reference torch_forward), fp32, {cores} torch threads, {cpu_sec:.1f} s"}
        # kernels of libtimeviper_b200.so per step on rank 0.  single GPU: conv, dt cumsum, fused SSD (5 stage kernels
        # in the CUDA-core family), norm.  sharded: + suffix scan and state pass of pass 1 (the cumsum is shared by both
        # passes); ranks > 0 add the 3-row halo conv and the boundary-state fold.
        launches_per_step = {"simt": 7, "tcgen05": 4}[family] + (2 if dist_on else 0)
        line = {
            "metric": METRIC, "value": Ltot / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "Nanov2-9B Mamba-2 mixer layer prefill (conv1d+SiLU -> SSD scan -> gated RMSNorm), "
                                   "batch 1, bf16, random init", "seqlen": Ltot, "tokens_per_gpu": L,
                       "parallelism": f"sp{world}" if dist_on else "single", "ssd_kernel_family": family,
                       "l2": "inputs (5.9 GB) >> L2 (126 MB); no flush needed", "dims": "H128 P80 G8 N128 Q128 hidden4480"},
            "e2e": {"value": Ltot / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "h2d_bytes_per_step": hs_host.numel() * 2 * world, "d2h_bytes_per_step": out_host.numel() * 2 * world,
                    "api": ("Mamba2MixerPrefill.prefill_from_host (H2D / in_proj+conv+SSD+norm+out_proj / D2H pipelined over "
                            "segments, pinned host buffers)" if not dist_on else
                            "sharded_prefill_from_host (H2D / sharded in_proj+conv+SSD+norm+out_proj / D2H on three streams, "
                            "double-buffered: the H2D of step i+1 overlaps the D2H of step i), pinned host buffers")},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / pk["hbm_gbs"],
                         "traffic": NCU_TRAFFIC_128K[dom] if (Ltot == 131072 and world == 1) else None,
                         "traffic_source": "profiles/r01_summary.md (ncu --set full, per launch, bytes)", "peak_source": pk["source"],
                         "algorithmic_bytes_per_token": BYTES_PER_TOKEN[dom]},
            "kernels": kernels,
            "path_roofline": {"bytes_per_token": path_bytes,
                              "frac": path_bytes * Ltot / world / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
        }
        if cpu_block is not None:
            line["cpu_baseline"] = cpu_block
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seqlen", type=int, default=131072)
    ap.add_argument("--cpu-sample", type=int, default=16384, help="tokens in the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-segment", type=int, default=16384, help="tokens per streamed segment in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
