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
# Nanov2-9B Mamba-2 layer dims (timeviper_b200.config.Mamba2Config.nanov2_9b holds the same numbers --
# repeated here so that the CPU reference arm never imports the product package, i.e. never maps its .so)
DIMS_9B = dict(hidden=4480, H=128, P=80, G=8, N=128, Q=128, K=4)
DIMS_9B["conv_dim"] = DIMS_9B["H"] * DIMS_9B["P"] + 2 * DIMS_9B["G"] * DIMS_9B["N"]
DIMS_9B["proj"] = DIMS_9B["H"] * DIMS_9B["P"] + DIMS_9B["conv_dim"] + DIMS_9B["H"]


def ncu_traffic(seqlen, world):
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the path kernels from the newest
    profiles/rNN_ncu_traffic.json, which tools/ncu_traffic.py regenerates from an `ncu --set full` capture of this
    command; None when no capture matches the configuration."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    if d.get("seqlen") != seqlen or d.get("n_gpus") != world:
        return None, os.path.relpath(files[-1], ROOT)
    return d["traffic_bytes_per_launch"], os.path.relpath(files[-1], ROOT)


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
    d = DIMS_9B
    H, P, G, N = d["H"], d["P"], d["G"], d["N"]
    torch.manual_seed(1234)
    p = R.nemotron_random_params(d["hidden"], H, P, G, N, nondegenerate=False)
    L = sample_tokens
    proj = torch.randn(1, L, d["proj"]) * 0.5

    def step():
        gate, xBC, dt = proj.split([H * P, d["conv_dim"], H], dim=-1)
        xc, _ = R.causal_conv1d_ref(xBC.transpose(1, 2), p["conv1d.weight"].squeeze(1), p["conv1d.bias"])
        x, Bm, Cm = xc.transpose(1, 2).split([H * P, G * N, G * N], dim=-1)
        y, s = R.ssd_chunked_ref(x.reshape(1, L, H, P), dt, -torch.exp(p["A_log"]), Bm.reshape(1, L, G, N),
                                 Cm.reshape(1, L, G, N), d["Q"], D=p["D"], dt_bias=p["dt_bias"],
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


def cpu_threads(args):
    """All host cores for the CPU arm: torchrun exports OMP_NUM_THREADS=1, which would otherwise make the arm run on a
    single thread at N > 1 (the round-1 ratios at N = 2/4/8 were void for that reason)."""
    n = args.cpu_threads or os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def reference_config1():
    """BASELINE.json configs[0]: the LITERAL reference -- the unmodified NemotronHMamba2Mixer.forward -> torch_forward
    (modeling_nano.py:671-859) -- small config (hidden 512, 16 heads x 80, G=1, N=128, Q=128), batch 1, 4096 tokens, fp32,
    on the host cores.  Needs the reference package staged under baseline/_ref (oracle/stage_reference.py)."""
    try:
        from oracle import stage_reference
        if not stage_reference.available():
            return {"unavailable": "baseline/_ref/nano not staged on this box"}
        mn, Cfg = stage_reference.load()
        torch.manual_seed(1234)
        hidden, H, P, G, N, Q, L = 512, 16, 80, 1, 128, 128, 4096
        cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G, ssm_state_size=N,
                  mamba_chunk_size=Q, mamba_d_conv=4, num_hidden_layers=2, hybrid_override_pattern="M-",
                  layer_norm_epsilon=1e-5)
        mixer = mn.NemotronHMamba2Mixer(cfg, layer_idx=0).float().eval()
        hs = torch.randn(1, L, hidden)
        best = None
        with torch.no_grad():
            for _ in range(2):
                t0 = time.perf_counter()
                mixer(hs)
                dt_s = time.perf_counter() - t0
                best = dt_s if best is None else min(best, dt_s)
        return {"value": L / best, "unit": UNIT, "seconds": best, "kind": "_ref", "cores": torch.get_num_threads(),
                "workload": "unmodified NemotronHMamba2Mixer.torch_forward, hidden 512, H16 P80 G1 N128 Q128, L=4096, fp32 "
                            "(in_proj and out_proj included)"}
    except Exception as e:      # never let the optional leg take the arm down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads(args)
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
        "same_config_note": "the literal torch_forward needs ~34 GB at the 9B dims (it materialises (b,c,l,s,h,n)); this arm "
                            "times its memory-lean restatement on a bounded token sample of the same layer",
        "config1_literal_reference": reference_config1(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def time_region(fn, steps, dist_on):
    """fn() may return a CUDA event that marks the end of its asynchronous tail (the last D2H copy of the host-buffer API
    runs on a copy stream): the timed region then ends when the tail of the LAST step has completed."""
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tail = None
    for _ in range(steps):
        tail = fn()
    if isinstance(tail, torch.cuda.Event):
        torch.cuda.current_stream().wait_event(tail)
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


def sharded_parity(tv, mixer, proj, rank, world):
    """Self-validation of the sharded path at the timed configuration: rank 0 gathers every rank's projected shard, runs
    the UNSHARDED core on the whole sequence once and compares (a) a strided token sample of every shard's output and
    (b) the last rank's final SSM state with what the sharded run produced.  Relative errors = max|a-b| / max|b|."""
    import torch.distributed as dist
    dev = proj.device
    with torch.no_grad():
        y_sh, ssm_sh = tv.sharded_scan_core(mixer, proj)
        stride = 257
        sample = y_sh[:, ::stride].contiguous()
        shards = [torch.empty_like(proj) for _ in range(world)] if rank == 0 else None
        dist.gather(proj.contiguous(), shards, dst=0)
        samples = [torch.empty_like(sample) for _ in range(world)] if rank == 0 else None
        dist.gather(sample, samples, dst=0)
        ssm_last = ssm_sh.clone()
        dist.broadcast(ssm_last, src=world - 1)
        res = torch.zeros(2, device=dev, dtype=torch.float64)
        if rank == 0:
            full = torch.cat(shards, dim=1)
            del shards
            y_full, ssm_full = mixer.scan_core(full, return_states=True)
            Ls = proj.shape[1]
            scale = float(y_full.float().abs().max())
            y_err = 0.0
            for r in range(world):
                ref = y_full[:, r * Ls:(r + 1) * Ls][:, ::stride].float()
                y_err = max(y_err, float((samples[r].float() - ref).abs().max()) / scale)
            s_err = float((ssm_last - ssm_full).abs().max() / ssm_full.abs().max())
            res[0], res[1] = y_err, s_err
            del full, y_full
        dist.broadcast(res, src=0)
        torch.cuda.empty_cache()
    return float(res[0]), float(res[1])


def bind_to_gpu_numa_node(index):
    """Restrict this process to the CPU cores NVML reports as local to GPU `index` (so that pinned buffers are first-touched
    on that node).  Returns the number of cores, or None when NVML / affinity are unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)]
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return None


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

    # pinned host buffers on the NUMA node of this GPU (first touch by a thread bound to the GPU's cores): with 8 ranks
    # copying at once the PCIe rate otherwise drops to what the cross-socket link gives
    all_cores = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    # synthetic tokens: the block RMS-normalises its input (modeling_nano.py:941) => unit-variance rows
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    hs_host = torch.empty(1, L, cfg.hidden_size, dtype=torch.bfloat16).pin_memory()
    hs_dev = torch.randn(1, L, cfg.hidden_size, device=dev, generator=g).to(torch.bfloat16)
    hs_host.copy_(hs_dev)
    out_host = torch.empty(1, L, cfg.hidden_size, dtype=torch.bfloat16).pin_memory()
    out_host.zero_()                                  # first touch on the GPU's NUMA node
    os.sched_setaffinity(0, all_cores)                # the CPU baseline leg and the Python side use every core again
    with torch.no_grad():
        proj = mixer.in_proj(hs_dev)                                    # resident input of the hot path
    family = tv.ssd_kernel_family(torch.bfloat16, cfg.mamba_head_dim, cfg.ssm_state_size, cfg.chunk_size)

    use_graph = not args.no_graph

    def core():
        with torch.no_grad():
            if dist_on:     # sharded step; one CUDA graph replay per step when the peer (symmetric-memory) exchange is in use
                return (tv.sharded_scan_core_graph(mixer, proj) if use_graph else tv.sharded_scan_core(mixer, proj))[0]
            if use_graph:       # the three kernels + the dt/cumsum pre-kernel replayed as ONE CUDA graph launch
                return mixer.scan_core_graph(proj)
            return mixer.scan_core(proj)

    parity = None
    if dist_on:
        y_rel, state_rel = sharded_parity(tv, mixer, proj, rank, world)
        parity = {"y_rel": y_rel, "state_rel": state_rel, "tol": 2e-2,
                  "what": "sharded vs one unsharded run on rank 0: strided token sample (every 257th) of every shard's "
                          "output, and the last rank's final SSM state; bf16"}
        if not (y_rel < 2e-2 and state_rel < 2e-2):
            raise SystemExit(f"bench.py: sharded result out of tolerance: {parity}")
    # kernels of libtimeviper_b200.so per step on this rank: counted by the library itself around one eager step
    with torch.no_grad():
        n0 = tv.launch_count()
        (tv.sharded_scan_core(mixer, proj) if dist_on else mixer.scan_core(proj))
        launches_per_step = tv.launch_count() - n0

    def e2e():
        with torch.no_grad():
            if dist_on:     # public host-buffer API of the sharded path: copies and compute on three streams
                # consecutive steps are independent sequences: the H2D copy of step i+1 and the D2H copy of step i-1 run
                # beside the compute of step i (three streams, double-buffered); the region ends after the last D2H copy
                _, done = tv.sharded_prefill_from_host(mixer, hs_host, out_host)
                return done
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
                for _ in range(10):     # short kernels after the idle gap: let the clocks come back up before timing
                    fn()
                per[name] = time_region(fn, max(3, min(args.steps, 10)), False)
                time.sleep(0.5)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        e2e()
    torch.cuda.synchronize()
    ms_e2e = time_region(e2e, e2e_steps, dist_on)

    # sustained: the same step back to back for a few seconds (the board settles into its power cap); reported beside
    # the burst figure the headline `value` is
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / max(ms, 1e-3)))
        ms_sus = time_region(core, n_sus, dist_on)
        sustained = {"ms_per_step": ms_sus, "value": Ltot / (ms_sus * 1e-3), "unit": UNIT, "steps": n_sus}
    if dist_on:
        dist.barrier()

    if rank == 0:
        pk = peaks()
        # the roofline block always describes the scan: the dominant kernel of the path (46 % of the step at N = 1) and the one
        # furthest below its roofline; at N > 1 the per-kernel times of the short shard-sized launches are noisier
        dom = "ssd" if "ssd" in per else max(per, key=per.get)
        ach = BYTES_PER_TOKEN[dom] * L / (per[dom] * 1e-3) / 1e9
        kernels = {k: {"ms": v, "gbs": BYTES_PER_TOKEN[k] * L / (v * 1e-3) / 1e9,
                       "frac": BYTES_PER_TOKEN[k] * L / (v * 1e-3) / 1e9 / pk["hbm_gbs"]} for k, v in per.items()}
        path_bytes = sum(BYTES_PER_TOKEN.values())
        cpu_tps, cpu_sec = (None, None)
        cpu_block = None
        if world == 1 and not args.no_cpu_baseline:
            cores = cpu_threads(args)
            cpu_tps, cpu_sec = cpu_reference_tokens_per_s(args.cpu_sample)
            cpu_block = {"value": cpu_tps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.cpu_sample} tokens of the same layer, oracle/mamba2_ref.py (restatement of "
                                   f"the reference torch_forward), fp32, {cores} torch threads, {cpu_sec:.1f} s"}
        traffic, traffic_src = ncu_traffic(Ltot, world)
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
                    "host_cores_bound_to_gpu_numa_node": numa,
                    "api": ("Mamba2MixerPrefill.prefill_from_host (H2D / in_proj+conv+SSD+norm+out_proj / D2H pipelined over "
                            "segments, pinned host buffers)" if not dist_on else
                            "sharded_prefill_from_host (H2D / sharded in_proj+conv+SSD+norm+out_proj / D2H on three streams, "
                            "double-buffered: the H2D of step i+1 overlaps the D2H of step i), pinned host buffers")},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / pk["hbm_gbs"],
                         "traffic": (traffic or {}).get(dom),
                         "traffic_source": traffic_src, "peak_source": pk["source"],
                         "algorithmic_bytes_per_token": BYTES_PER_TOKEN[dom]},
            "kernels": kernels,
            "path_roofline": {"bytes_per_token": path_bytes,
                              "frac": path_bytes * Ltot / world / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
        }
        if cpu_block is not None:
            line["cpu_baseline"] = cpu_block
        if parity is not None:
            line["parity"] = parity
        if sustained is not None:
            line["sustained"] = sustained
        line["config"]["launch"] = ("one CUDA graph replay per step" if use_graph else "eager kernel launches")
        if dist_on:
            from timeviper_b200 import sharded as _sh
            line["config"]["boundary_exchange"] = ("symmetric memory: summary in place, one device barrier, fold reads peers "
                                                   "over NVLink" if _sh._exchanges else "NCCL all-gather")
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


def run_hybrid(args):
    """BASELINE.json configs[3] / [4]: prefill of the 56-layer Nanov2-9B-shaped hybrid LM (random init) over synthetic video
    tokens, Mamba-2 layers on this package's kernels, attention on library SDPA, MLP and projections on cuBLAS, last-token
    lm_head.  A step = one whole prefill from token embeddings resident in HBM to the fp32 logits of the last position.
    --workload hybrid9b      : 5K frames (81,920 tokens), no token drop                                   (configs[3])
    --workload hybrid9b-pdrop: 10K frames (163,840 tokens + 64 text), TransV / pyramid-drop at layers 14/21/30/39 with the
                               reference's default schedule (evaluate.py:167-172)                          (configs[4])
    N > 1: N / S replicas of ONE sample each, every sample sharded over S = --shard GPUs (default 1: N independent replicas,
    no collective on the data path).  With S > 1 the layer loop runs sequence-sharded (Mamba-2 layers: conv halo + boundary
    states; attention layers: K/V all-gather; after a pyramid-drop the survivors are re-balanced with one all-to-all), so
    BASELINE configs[4]'s "batch 4 across 8 GPUs" is `--workload hybrid9b-pdrop --gpus 8 --shard 2`."""
    import timeviper_b200 as tv
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    drop = args.workload == "hybrid9b-pdrop"
    S = max(1, args.shard)
    if world % S:
        raise SystemExit("bench.py: --shard must divide the number of GPUs")
    grp = None
    if S > 1:       # process groups of S consecutive ranks; every rank creates all of them (collective call)
        for g0 in range(0, world, S):
            gnew = dist.new_group(list(range(g0, g0 + S)))
            if g0 <= rank < g0 + S:
                grp = gnew
    text = 64 if drop else 0
    L = args.seqlen if args.seqlen != 131072 else (163840 + text if drop else 81920)
    cfg = tv.Mamba2Config.nanov2_9b_hybrid()
    torch.manual_seed(rank // S)                  # the ranks of one sample hold the same weights
    with torch.device("cuda"):
        model = tv.HybridCausalLM(cfg).to(torch.bfloat16).eval()
    from timeviper_b200.hybrid import shard_bounds
    so = shard_bounds(L, S)
    x = torch.randn(1, so[rank % S + 1] - so[rank % S], cfg.hidden_size, device="cuda").to(torch.bfloat16)   # this rank's shard
    pd = dict(pdrop_type="uni_14_0.8-attn_21_0.6-attn_30_0.4-attn_39_0.2", first_vision_token_position=0,
              num_vision_tokens=L - text, text_prompt_len=text) if drop else None
    run = lambda inp: model(inputs_embeds=inp, pdrop=pd, group=grp)
    share, ev = {}, []

    def pre(m, a):
        e = torch.cuda.Event(enable_timing=True); e.record(); m._e0 = e

    def post(m, a, o):
        e = torch.cuda.Event(enable_timing=True); e.record(); ev.append((m.block_type, m._e0, e))
    for _ in range(max(args.warmup, 1)):
        logits = run(x)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.3)
    n0 = tv.launch_count()
    t0 = time.time()
    ms = time_region(lambda: run(x), args.steps, dist_on)
    clocks = sampler.stop(t0, time.time()) if rank == 0 else None
    launches = tv.launch_count() - n0
    hooks = []
    for layer in model.backbone.layers:
        hooks += [layer.register_forward_pre_hook(pre), layer.register_forward_hook(post)]
    logits = run(x); torch.cuda.synchronize()
    for kind, a, b in ev:
        share[kind] = share.get(kind, 0.0) + a.elapsed_time(b)
    for hk in hooks:
        hk.remove()
    # end to end through the public API with host buffers: pinned embeddings -> H2D -> prefill -> D2H of the logits
    hx = torch.empty(x.shape, dtype=x.dtype).pin_memory(); hx.copy_(x.cpu())
    hl = torch.empty(logits.shape, dtype=logits.dtype).pin_memory()
    k = max(1, min(args.e2e_steps, 3))

    def e2e():
        hl.copy_(run(hx.cuda(non_blocking=True)), non_blocking=True)
    e2e_ms = time_region(e2e, k, dist_on)
    finite = bool(torch.isfinite(logits).all())
    if dist_on:
        fin = torch.tensor([1.0 if finite else 0.0], device="cuda")
        dist.all_reduce(fin, op=dist.ReduceOp.MIN)
        finite = bool(fin.item() > 0)
    if rank == 0:
        pat = cfg.hybrid_override_pattern
        print(json.dumps({
            "metric": "hybrid_prefill_tokens_per_s", "value": (world // S) * L / ms * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "Nanov2-9B-shaped hybrid LM prefill (BASELINE.json configs[%d]): 56 layers " % (4 if drop else 3) +
                                   f"({pat.count('M')} Mamba-2 / {pat.count('*')} attention / {pat.count('-')} MLP), random init, "
                                   "last-token lm_head" + (", TransV / pyramid-drop " + pd["pdrop_type"] if drop else ""),
                       "seqlen": L, "global_batch": world // S,
                       "parallelism": "single" if world == 1 else f"{world // S} replicas x batch 1" + (f", each sequence-sharded over {S} GPUs" if S > 1 else ""),
                       "params_B": round(sum(p.numel() for p in model.parameters()) / 1e9, 2),
                       "l2": "activations (GBs per layer) >> L2 (126 MB); no flush needed"},
            "e2e": {"value": (world // S) * L / e2e_ms * 1e3, "unit": UNIT, "ms_per_step": e2e_ms, "steps": k, "h2d_bytes_per_step": hx.numel() * 2,
                    "d2h_bytes_per_step": hl.numel() * 4, "api": "HybridCausalLM.forward(inputs_embeds[, pdrop]) from pinned host embeddings to host logits"},
            "gpu_launches": launches, "clocks": clocks,
            "layer_time_share_ms": {k2: round(v, 1) for k2, v in sorted(share.items())},
            "finite": finite,
        }))
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
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-segment", type=int, default=16384, help="tokens per streamed segment in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (default: all host cores)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="back-to-back seconds for the sustained step time")
    ap.add_argument("--no-graph", action="store_true", help="N=1: launch the three kernels eagerly instead of one CUDA graph")
    ap.add_argument("--shard", type=int, default=1, help="hybrid workloads: GPUs per sample (sequence-sharded layer loop)")
    ap.add_argument("--workload", default="mixer", choices=["mixer", "hybrid9b", "hybrid9b-pdrop"],
                    help="mixer: the BASELINE metric (default); hybrid9b / hybrid9b-pdrop: BASELINE.json configs[3] / [4], own JSON line")
    args = ap.parse_args()
    if args.workload != "mixer" and args.impl == "ours":
        run_hybrid(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
