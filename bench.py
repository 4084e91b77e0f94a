#!/usr/bin/env python
"""bench.py — headline measurement of the CSN hot path on B200 (contract: task prompt §④).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--heads H]

One step = one pass of the hot path over one batch of synthetic input:
  * CSA training step (BASELINE.json configs[1]): B = 8 query shapes x K = 3 neighbour shapes,
    N = 10 000 points x 256-d, forward (CrossShapeAt.forward, csa_models.py:182-202) + masked CE
    (csa_training.py:94-108) + backward.  metric = shape-pairs/s (B*K pairs per step), whole job.
  * kNN retrieval (configs[2]), reported in the `knn` object of the same line: query shapes scored
    against a 4 000-shape candidate store + top-(K+1); shapes/s.
`value` is device-resident throughput (inputs in HBM before the timed region); `e2e` goes through the
module API with pinned HOST buffers (H2D of every step's features inside the timed region, loss read
back).  With torchrun (N > 1) query shapes are sharded data-parallel, one rank per GPU, parameter
gradients all-reduced over NCCL each step (weak scaling: per-GPU batch fixed).

--impl reference times the reference's own PyTorch implementation of the same step on the host cores: the
unmodified MID-FC/csa_models.py when oracle/_ref holds it (copied there by __graft_entry__.build() in the build
container; git-ignored, travels with the snapshot; kind "reference"), else the oracle port (kind "port").

--config 2 (default) is the headline line (configs[1] + the configs[2] kNN sub-object); --config 4 / 5 print the
MinkowskiNet-head and the N-sweep lines (BASELINE.json configs[3], configs[4]) with the same keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

N_POINTS, D = 10000, 256
CSA_B, CSA_K, N_CLASSES = 8, 3, 15
KNN_CANDIDATES, KNN_TOPK = 4000, 5


def csa_flops_per_query(K: int, h: int, d: int = 256, C: int = 500, N: int = N_POINTS) -> float:
    """Algorithmic fwd FLOPs of one de-duplicated CSA layer call:
    2*N*h*d*[3(K+1)D + (1+2K)(2C + D)]; fwd+bwd = 3x."""
    return 2.0 * N * h * d * (3 * (K + 1) * D + (1 + 2 * K) * (2 * C + D))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_info() -> dict:
    """Physical cores and CPU model of the box (BASELINE.md §3 asks for both beside every CPU number)."""
    model, cores = "unknown", set()
    try:
        phys = core = None
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name") and model == "unknown":
                model = line.split(":", 1)[1].strip()
            elif line.startswith("physical id"):
                phys = line.split(":", 1)[1].strip()
            elif line.startswith("core id"):
                core = line.split(":", 1)[1].strip()
            elif not line.strip():
                if phys is not None and core is not None:
                    cores.add((phys, core))
                phys = core = None
    except OSError:
        pass
    logical = os.cpu_count() or 1
    return {"cpu_model": model, "physical_cores": len(cores) or logical, "logical_cpus": logical}


def load_reference_midfc():
    """The unmodified reference module from oracle/_ref (see __graft_entry__.build), or None."""
    p = ROOT / "oracle" / "_ref" / "MID-FC" / "csa_models.py"
    if not p.exists():
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_csa_models", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.device = torch.device("cpu")
    return mod


def cpu_step_factory(h: int, batch: int, K: int):
    """(step(), kind): one forward + masked CE + backward of the CSA layer on the host, through the reference's own
    module when available (kind "reference"), else through the oracle port (kind "port").  fp32, eval mode."""
    from csn_b200 import synth
    sd = synth.midfc_state(1, h, N_CLASSES)
    x, nb = synth.csa_batch(2, batch, K)
    label = torch.randint(0, N_CLASSES, (batch, N_POINTS), generator=synth.gen(3))
    ref = load_reference_midfc()
    if ref is not None:
        m = ref.get_model("csa", N_CLASSES, h, K).eval()
        m.load_state_dict(sd)
        params = [p for p in m.parameters()]

        def step():
            for p in params:
                p.grad = None
            logits = m(x, "test", nb)
            lg = logits.squeeze(-1).permute(0, 2, 1).contiguous().view(-1, N_CLASSES)   # csa_training.py:94-108
            lb = label.view(-1)
            keep = torch.where(lb > 0)[0]
            loss = torch.nn.functional.cross_entropy(lg[keep], lb[keep])
            loss.backward()
            return float(loss.detach())
        return step, "reference"
    from oracle import csa_oracle as O
    w = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}

    def step():
        for v in w.values():
            v.grad = None
        loss = O.masked_cross_entropy(O.forward_csa(x, nb, w, h), label)
        loss.backward()
        return float(loss.detach())
    return step, "port"


def masked_ce(logits, label):
    """Mean CE over the points whose label is > 0 (MID-FC/csa_training.py:94-108). Written with
    ignore_index instead of the reference's boolean-mask gather: same value and gradient, but no
    data-dependent shapes (no device->host sync inside the step)."""
    return torch.nn.functional.cross_entropy(logits.squeeze(-1), label, ignore_index=0)


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    """The reference's CPU implementation of the same step (same batch, same steps / warm-up) on the host cores;
    rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    hi = host_info()
    cores = hi["physical_cores"]
    torch.set_num_threads(cores)
    h = args.heads
    batch = CSA_B
    step, kind = cpu_step_factory(h, batch, CSA_K)
    t0 = time.perf_counter()
    step()                                   # first warm-up step, also the probe that bounds the run
    t1 = time.perf_counter() - t0
    steps, warmup = args.steps, args.warmup
    note = ""
    if t1 * (steps + warmup) > 240.0:        # a slow host: keep the whole run within a few minutes
        steps = max(1, min(steps, int(200.0 / t1) - 1))
        warmup = 1
        note = f" (host step {t1:.1f} s: bounded to {steps} timed steps)"
    for _ in range(max(0, warmup - 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = batch * CSA_K / dt
    sample = (f"full step: {batch} query shapes x K={CSA_K} neighbours, fwd + masked CE + bwd, fp32, eval, {steps} steps, "
              f"{warmup} warm-up{note}; {hi['cpu_model']}, {cores} physical cores ({hi['logical_cpus']} logical), "
              f"torch {torch.__version__}")
    line = {
        "impl": "reference", "metric": "csa_shape_pairs_per_s_fwd_bwd", "value": value, "unit": "shape-pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MID-FC CSA training step B={CSA_B} K={CSA_K} h={h} N={N_POINTS} D={D} (configs[1])",
                   "heads": h},
        "cpu_baseline": {"value": value, "unit": "shape-pairs/s", "cores": cores, "kind": kind, "sample": sample,
                         "cpu_model": hi["cpu_model"]},
        "e2e": {"value": value, "unit": "shape-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import torch.distributed as dist

    from csn_b200 import _lib as L
    from csn_b200 import knn, midfc, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = args.heads
    pk = peaks()

    # ------------------------------------------------------------------ CSA training step
    model = midfc.get_model("csa", N_CLASSES, h, CSA_K, precision=args.precision).to(dev).eval()
    model.load_state_dict(synth.midfc_state(1, h, N_CLASSES))
    if args.train_mode:
        model.train()
    params = [p for n, p in model.named_parameters() if not n.startswith("fc_1")]
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    # device-resident inputs (two alternating batches > L2 each: 82 MB + 328 MB)
    batches = []
    for _ in range(2):
        nb = torch.relu(torch.randn(CSA_B, CSA_K + 1, D, N_POINTS, 1, device=dev, generator=g))
        x = nb[:, 0].clone()
        lab = torch.randint(0, N_CLASSES, (CSA_B, N_POINTS), device=dev, generator=g)
        batches.append((x, nb, lab))
    flat_grads = None

    def local_step(x, nb, lab):
        for p in params:
            p.grad = None
        if not args.unfused_loss:   # weighted sum + logit conv + masked CE + IoU counters + their backward: csn_csa_head
            loss = model.forward_loss(x, "test", nb, lab)
        else:
            loss = masked_ce(model(x, "test", nb), lab)
        loss.backward()
        return loss

    def exchange():
        if world > 1:  # data-parallel gradient exchange: one flat all-reduce over NCCL
            flat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
            dist.all_reduce(flat)
            flat.div_(world)

    def eager_step(x, nb, lab):
        loss = local_step(x, nb, lab)
        exchange()
        return loss

    # the whole local step (forward, loss, backward: ~120 launches) is captured once per static input set and
    # replayed with one launch; the gradient all-reduce stays outside the graph
    graphs = None
    launch_mode = "eager"
    if not args.no_graph:
        from csn_b200.graphs import GraphedStep
        graphs = [GraphedStep(local_step, *b) for b in batches]
        launch_mode = "CUDA graph replay of the local step (csn_b200.graphs.GraphedStep)"

    step_no = [0]

    def train_step(k):
        if graphs is None:
            return eager_step(*batches[k])
        step_no[0] += 1
        # train mode: the seeds are frozen in the captured graph, the device-resident epoch makes every replay draw new masks
        loss = graphs[k].replay(epoch=step_no[0] if args.train_mode else None)
        exchange()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        train_step(i % 2)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def timed_block():
        """EXACTLY args.steps steps between two events, bracketed by barrier + synchronize; max over ranks."""
        l0 = L.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            train_step(i % 2)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), L.launch_count() - l0

    # the K-step block is repeated until ~1 s has been timed (a 20-step block is 0.06 s: too short for the clock
    # sampler and for the power state to settle); the reported figure is the MEDIAN block
    first_ms, launches = timed_block()
    n_blocks_timed = int(max(1, min(40, -(-1000.0 // max(first_ms, 1e-3)))))
    if world > 1:   # every rank must run the same number of blocks
        nb_t = torch.tensor([n_blocks_timed], device=dev)
        dist.broadcast(nb_t, 0)
        n_blocks_timed = int(nb_t.item())
    block_ms = [first_ms] + [timed_block()[0] for _ in range(n_blocks_timed - 1)]
    if graphs is not None:   # replays do not pass through the C ABI: launches recorded in the graphs x replays
        launches = sum(graphs[i % 2].launches for i in range(args.steps))
    ms_step = sorted(block_ms)[len(block_ms) // 2] / args.steps
    pairs = world * CSA_B * CSA_K
    value = pairs / (ms_step * 1e-3)
    step_flops = 3.0 * CSA_B * csa_flops_per_query(CSA_K, h)
    achieved = step_flops / (ms_step * 1e-3) / 1e12

    # ------------------------------------------------------------------ per-kernel breakdown (outside the timed region)
    # one extra step with CUDA events around every C-ABI launch on the launching stream
    L.profile_begin()
    eager_step(*batches[0])
    prof = L.profile_end()
    # ALGORITHMIC work per entry point (N = 10 000 points, 500-key chunks — not the 10 240-row slots /
    # 512-row chunk tiles the kernels execute; flash recompute of the scores is not counted):
    #   S = B(K+1) projected shapes, nblk = B(2K+1) attention blocks, unit = one 2*N*C*HD contraction per block
    S_, nblk, HD = CSA_B * (CSA_K + 1), CSA_B * (2 * CSA_K + 1), 256 * h
    unit = 2.0 * nblk * N_POINTS * 500 * HD
    proj = 2.0 * N_POINTS * D * HD
    alg_flops = {
        "csn_gemm_colbias": S_ * 3 * proj,                     # Q|K|V projection of every shape (once per step)
        "csn_attn_fwd": 2 * unit,                              # Q K^T + P V
        "csn_gemm_res_ln": nblk * proj,                        # out-projection (+ residual + LayerNorm statistics)
        "csn_gemm_delta": nblk * proj,                         # dO = dZ Wo (+ delta)
        "csn_attn_bwd_dv": unit,                               # P^T dO
        "csn_attn_bwd_dq": unit,                               # dP = dO V^T (dS kernel)
        "csn_gemm": nblk * proj + 2 * unit + 3 * nblk * proj,  # dWo ; dQ = dS K, dK = dS^T Q ; dWq|dWk|dWv
    }
    if "csn_gemm_colbias" not in prof:                         # V not centred: the projection is a plain csn_gemm
        alg_flops["csn_gemm"] += alg_flops.pop("csn_gemm_colbias")
    # algorithmic bytes of the HBM-bound entry points: every operand / result once at its stored width
    rows_blk, rows_slot = nblk * N_POINTS, S_ * N_POINTS
    alg_bytes = {
        "csn_pack_rows": rows_slot * D * (4 + 2),                              # fp32 in, 16-bit out
        "csn_gemm_colbias": rows_slot * (2 * D + 2 * 3 * HD),
        "csn_gemm_res_ln": rows_blk * (2 * HD + 4 * D + 4 * D),                # O + residual in, Z out
        "csn_ln_colsum": S_ * N_POINTS * 4 * D,
        "csn_csa_head": rows_blk * 4 * D + CSA_B * N_POINTS * 4 * D,           # Z in (once), dOutT out
        "csn_ln_bwd": rows_blk * (4 * D + 2 * D) + CSA_B * N_POINTS * 4 * D,   # Z + dOutT in, dZ16 out
        "csn_gemm_delta": rows_blk * (2 * D + 2 * HD + 2 * HD),                # dZ16 + O in, dO out
    }
    kernels = {}
    for name, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"ms": round(d["ms"], 4), "launches": d["launches"]}
        if name in alg_flops:
            e["tflops"] = round(alg_flops[name] / (d["ms"] * 1e-3) / 1e12, 1)
            e["frac_of_sustained_peak"] = round(e["tflops"] / pk["tflops_sustained"], 3)
        if name in alg_bytes:
            e["algorithmic_gbs"] = round(alg_bytes[name] / (d["ms"] * 1e-3) / 1e9, 1)
            e["frac_of_hbm_peak"] = round(e["algorithmic_gbs"] / pk["hbm_gbs"], 3)
        kernels[name] = e
    # dominant kernel = the tensor-core entry point with the largest share of the step (SURVEY §8d classifies every
    # contraction as tensor-bound); its HBM view is reported beside it
    dom_name = max((n for n in prof if n in alg_flops), key=lambda n: prof[n]["ms"])
    dom_d = prof[dom_name]
    dom_tflops = alg_flops[dom_name] / (dom_d["ms"] * 1e-3) / 1e12
    traffic_json = {}
    tpath = ROOT / "profiles" / "r2_traffic.json"
    if tpath.exists():
        traffic_json = json.loads(tpath.read_text())
    compulsory = (CSA_B * (CSA_K + 1) * N_POINTS * D * 4.0      # the step's distinct input shapes, fp32
                  + CSA_B * N_POINTS * 8.0)                      # labels
    roof = {"bound": "tensor", "achieved": dom_tflops, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
            "frac": dom_tflops / pk["tflops_sustained"],
            "kernel": dom_name + " (tensor-core entry point with the largest share of the step; all its launches: "
                                 "algorithmic FLOPs / sum of CUDA-event durations)",
            "launches_per_step": dom_d["launches"], "ms_per_step": round(dom_d["ms"], 4),
            "algorithmic_flops_per_launch": alg_flops[dom_name] / dom_d["launches"],
            "traffic": traffic_json.get(dom_name, {}).get("dram_bytes_per_launch"),
            "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
            "whole_step": {"algorithmic_flops": step_flops, "achieved": achieved, "unit": "TFLOP/s",
                           "frac": achieved / pk["tflops_sustained"],
                           "dram_bytes": traffic_json.get("whole_step", {}).get("dram_bytes"),
                           "compulsory_bytes": compulsory,
                           "note": "dram_bytes: sum over the step's kernels from the committed ncu --set full capture "
                                   "(profiles/); compulsory_bytes: the step's distinct inputs once"}}

    # ------------------------------------------------------------------ e2e: host buffers through the module API
    hx = [torch.empty(CSA_B, D, N_POINTS, 1).pin_memory() for _ in range(2)]
    hn = [torch.empty(CSA_B, CSA_K + 1, D, N_POINTS, 1).pin_memory() for _ in range(2)]
    hl = [torch.empty(CSA_B, N_POINTS, dtype=torch.int64).pin_memory() for _ in range(2)]
    for i in range(2):
        hx[i].copy_(batches[i][0]); hn[i].copy_(batches[i][1]); hl[i].copy_(batches[i][2])
    h2d = hx[0].numel() * 4 + hn[0][:, 1:].numel() * 4 + hl[0].numel() * 8

    # Every step's inputs start in pinned HOST memory and are copied inside the timed region; the copy of step
    # i+1 runs on a side stream while step i computes (double-buffered device staging), the loss is read back
    # to the host every step.
    copy_stream = torch.cuda.Stream(device=dev)
    consumed = [None, None]   # event: the step that read static input set k has finished

    def stage(i):
        k = i % 2
        x, nb, lab = batches[k]   # static input set k of graph k
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])
            x.copy_(hx[k], non_blocking=True)
            for b in range(CSA_B):   # slot 0 (the query itself) is never read by the layer: not copied
                nb[b, 1:].copy_(hn[k][b, 1:], non_blocking=True)
            lab.copy_(hl[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_run(n):
        nxt = stage(0)
        last = 0.0
        for i in range(n):
            ev = nxt
            if i + 1 < n:
                nxt = stage(i + 1)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            loss = train_step(i % 2)
            done = torch.cuda.Event()
            done.record(cur)
            consumed[i % 2] = done
            last = loss.item()   # device -> host read of the step's loss
        return last

    e2e_steps = max(2, min(args.steps, 5))
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = pairs * e2e_steps / t_e2e.item()

    # ------------------------------------------------------------------ e2e from a 16-bit pinned host cache
    # The same host-fed path with the collection kept as fp16 in pinned host memory (converted once, when the cache
    # is built): half the PCIe bytes per step; the device side packs straight from the 16-bit staging buffers
    # (csn_pack_rows_src16).  Reported next to `e2e` (fp32 host buffers, the reference loader's format), not instead.
    e2e_host16 = None
    if not args.no_graph and not args.train_mode and not args.unfused_loss:
        dev16 = [(b[0].half(), b[1].half(), b[2]) for b in batches]
        graphs16 = [GraphedStep(local_step, *b) for b in dev16]
        hx16 = [t[0].cpu().pin_memory() for t in dev16]
        hn16 = [t[1].cpu().pin_memory() for t in dev16]
        consumed16 = [None, None]

        def stage16(i):
            k = i % 2
            x, nb, lab = dev16[k]
            with torch.cuda.stream(copy_stream):
                if consumed16[k] is not None:
                    copy_stream.wait_event(consumed16[k])
                x.copy_(hx16[k], non_blocking=True)
                for b in range(CSA_B):
                    nb[b, 1:].copy_(hn16[k][b, 1:], non_blocking=True)
                lab.copy_(hl[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ev

        def run16(n):
            nxt = stage16(0)
            last = 0.0
            for i in range(n):
                ev = nxt
                if i + 1 < n:
                    nxt = stage16(i + 1)
                cur = torch.cuda.current_stream()
                cur.wait_event(ev)
                loss = graphs16[i % 2].replay()
                exchange()
                done = torch.cuda.Event()
                done.record(cur)
                consumed16[i % 2] = done
                last = loss.item()
            return last

        steps16 = max(2, min(args.steps, 10))
        run16(2)
        barrier()
        t0 = time.perf_counter()
        run16(steps16)
        barrier()
        t16 = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(t16, op=dist.ReduceOp.MAX)
        e2e_host16 = {"value": pairs * steps16 / t16.item(), "unit": "shape-pairs/s",
                      "h2d_bytes_per_step": int(hx16[0].numel() * 2 + hn16[0][:, 1:].numel() * 2 + hl[0].numel() * 8),
                      "d2h_bytes_per_step": 4, "steps": steps16,
                      "note": "features kept as an fp16 pinned host cache (built once); same step, csn_pack_rows_src16"}
        del graphs16, dev16, hx16, hn16

    # ------------------------------------------------------------------ e2e with the GPU-resident feature store
    # (SURVEY §8f-1): the per-point features are constants of the CSA phase, so only shape ids and labels cross PCIe;
    # the step's inputs are gathered by id on the device into the static input sets of the graphs.
    from csn_b200.store import ShardedFeatureStore
    per_rank_store = 96
    n_store = per_rank_store * world
    store = ShardedFeatureStore(n_store, N_POINTS, D, device=dev)      # rank r holds shapes [96 r, 96 (r+1)) of the collection
    gs_ = torch.Generator(device=dev).manual_seed(7 + rank)
    for s0 in range(0, store.hi - store.lo, 16):
        store.feats[s0:s0 + 16] = torch.relu(torch.randn(16, D, N_POINTS, device=dev, generator=gs_))
    store.barrier()   # every shard is loaded before anybody reads a peer's
    store_plane = ("one-sided: device-to-device copies out of the owners' symmetric-memory shards (copy engines over NVLink, no SMs)"
                   if store.one_sided else "two-sided: one batched NCCL point-to-point exchange per step"
                   + (f" (symmetric memory unavailable: {getattr(store, 'symm_error', 'n/a')[:120]})" if world > 1 else ""))
    # ids of EVERY rank's batch, derived from one shared seed (the replicated kNN graph + a DistributedSampler-style
    # permutation): queries are the rank's own shapes, neighbours are arbitrary shapes of the collection, so at N ranks
    # a fraction (N-1)/N of the neighbour blocks is fetched from its owner over NVLink
    gh = torch.Generator().manual_seed(11)
    id_steps = []
    for _ in range(8):
        ids_all = [(torch.randint(0, per_rank_store, (CSA_B,), generator=gh) + r * per_rank_store).tolist() for r in range(world)]
        nbr_all = [torch.randint(0, n_store, (CSA_B, CSA_K), generator=gh).tolist() for r in range(world)]
        id_steps.append((ids_all, nbr_all))
    hl_s = [torch.randint(0, N_CLASSES, (CSA_B, N_POINTS), generator=gh).pin_memory() for _ in range(2)]
    remote_blocks = []

    # one-time check of the data plane: a block fetched from a peer equals what its owner generated (same seed here)
    store_verified = None
    if world > 1:
        ids0, nbr0 = id_steps[0]
        xv, nbv = store.batch(ids0, nbr0)
        torch.cuda.synchronize()
        store_verified = True
        done_check = False
        for b_ in range(CSA_B):
            for k_ in range(CSA_K):
                sid = nbr0[rank][b_][k_]
                o_ = store.owner(sid)
                if o_ == rank or done_check:
                    continue
                go = torch.Generator(device=dev).manual_seed(7 + o_)
                lo_o = store.bounds[o_][0]
                for s0 in range(0, sid - lo_o + 1, 16):   # replay the owner's generator up to the chunk that holds sid
                    chunk = torch.relu(torch.randn(16, D, N_POINTS, device=dev, generator=go))
                want = chunk[(sid - lo_o) % 16]
                store_verified = bool(torch.equal(nbv[b_, k_ + 1].view(D, N_POINTS), want))
                done_check = True
        del xv, nbv

    consumed_s = [None, None]

    def store_stage(i):   # gathers + label copy of step i into static input set i % 2, on the side stream
        k = i % 2
        x, nb, lab = batches[k]
        ids, nbr = id_steps[i % len(id_steps)]
        with torch.cuda.stream(copy_stream):
            if consumed_s[k] is not None:
                copy_stream.wait_event(consumed_s[k])
            store.batch(ids, nbr, out=(x, nb))
            remote_blocks.append(store.last_remote_blocks)
            lab.copy_(hl_s[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def store_run(n):
        nxt = store_stage(0)
        last = 0.0
        for i in range(n):
            ev = nxt
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            loss = train_step(i % 2)
            done = torch.cuda.Event()
            done.record(cur)
            consumed_s[i % 2] = done
            if i + 1 < n:
                nxt = store_stage(i + 1)   # enqueued behind nothing it conflicts with: overlaps step i
            last = loss.item()
        return last

    store_steps = max(2, min(args.steps, 10))
    store_run(2)
    barrier()
    t0 = time.perf_counter()
    store_run(store_steps)
    barrier()
    t_st = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_st, op=dist.ReduceOp.MAX)
    e2e_store = {"value": pairs * store_steps / t_st.item(), "unit": "shape-pairs/s",
                 "h2d_bytes_per_step": int(hl_s[0].numel() * 8 + CSA_B * (CSA_K + 1) * 8), "d2h_bytes_per_step": 4,
                 "steps": store_steps,
                 "nvlink_bytes_per_step_per_rank": int(sum(remote_blocks[-store_steps:]) / max(1, store_steps) * D * N_POINTS * 4),
                 "data_plane": store_plane, "remote_block_verified": store_verified,
                 "note": "collection sharded by shape id over the ranks' HBM (csn_b200.store.ShardedFeatureStore): queries are "
                         "local, neighbour blocks are fetched from their owners on a side stream under the previous step; "
                         "only ids and labels cross PCIe"}
    del store
    clocks = sampler.stop() if rank == 0 else None
    del batches, hx, hn, hl
    graphs = None   # releases the graphs' private memory pools before the 40 GB candidate store is built
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ kNN retrieval
    knn_obj = None
    if not args.no_knn:
        n_c = args.knn_candidates
        q_per_step = 4
        per_rank = (n_c + world - 1) // world
        lo, hi = rank * per_rank, min(n_c, (rank + 1) * per_rank)
        gk = torch.Generator(device=dev).manual_seed(7)
        protos = torch.randn(16, 32, D, device=dev, generator=gk)
        store_rows = torch.empty(n_c * N_POINTS, D, dtype=torch.float16, device=dev)
        store_lo = torch.empty(n_c * N_POINTS, D, dtype=torch.float16, device=dev)   # residuals for the exact band re-score

        def make_shapes(ids):
            gs = torch.Generator(device=dev).manual_seed(1000 + int(ids[0]))
            out = torch.empty(len(ids), N_POINTS, D, device=dev)
            for j, s in enumerate(ids):
                part = torch.randint(0, 32, (N_POINTS,), device=dev, generator=gs)
                jit = torch.randn(32, D, device=dev, generator=gs)
                out[j] = torch.relu(protos[s % 16][part] + 0.15 * jit[part] + 0.5 * torch.randn(N_POINTS, D, device=dev, generator=gs))
            return out

        # each rank normalises its own block of the collection, then the blocks are exchanged
        for s0 in range(lo, hi, 50):
            ids = list(range(s0, min(hi, s0 + 50)))
            st = knn.build_store(make_shapes(ids), exact=True)
            store_rows[s0 * N_POINTS:(s0 + len(ids)) * N_POINTS] = st.rows
            store_lo[s0 * N_POINTS:(s0 + len(ids)) * N_POINTS] = st.rows_lo
        allgather_ms = 0.0
        if world > 1:
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for buf in (store_rows, store_lo):
                if n_c % world == 0:
                    dist.all_gather_into_tensor(buf, buf[lo * N_POINTS:hi * N_POINTS].clone())
                else:
                    for r in range(world):
                        rl, rh = r * per_rank, min(n_c, (r + 1) * per_rank)
                        dist.broadcast(buf[rl * N_POINTS:rh * N_POINTS], src=r)
            a1.record()
            barrier()
            allgather_ms = a0.elapsed_time(a1)
        cstore = knn.ShapeStore(store_rows, [s * N_POINTS for s in range(n_c)], [N_POINTS] * n_c, store_lo)
        # queries: this rank's shapes (sharded by query shape)
        k_steps = max(2, min(args.steps, 3))

        def knn_step(i):
            ids = [(lo + (i * q_per_step + j)) % n_c for j in range(q_per_step)]
            qstore = cstore.subset(ids)
            sc = knn.scores_from_stores(qstore, cstore)
            knn.refine_band(sc, qstore, cstore, KNN_TOPK)   # exact re-score of the top-K boundary band
            return knn.topk_rows(sc, KNN_TOPK)

        knn_step(0)
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(k_steps):
            knn_step(i + 1)
        k1.record()
        barrier()
        tk = torch.tensor([k0.elapsed_time(k1)], device=dev)
        if world > 1:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        kms = tk.item() / k_steps
        # e2e: the query shapes' SSA features start in pinned HOST memory (fp32 rows, as get_all_feats leaves them on the
        # CPU, csa_models.py:282-300); per step: H2D, normalisation into a query store, scoring against the resident
        # candidate store, exact band re-score, top-K, neighbour indices read back to the host
        hq = [make_shapes([(lo + 7 * k_ + j) % n_c for j in range(q_per_step)]).cpu().pin_memory() for k_ in range(2)]

        def knn_e2e_step(i):
            f = hq[i % 2].to(dev, non_blocking=True)
            qs = knn.build_store(f, exact=True)
            sc = knn.scores_from_stores(qs, cstore)
            knn.refine_band(sc, qs, cstore, KNN_TOPK)
            return knn.topk_rows(sc, KNN_TOPK)[1].cpu()

        knn_e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(k_steps):
            knn_e2e_step(i + 1)
        barrier()
        t_ke = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(t_ke, op=dist.ReduceOp.MAX)
        knn_e2e = {"value": world * q_per_step * k_steps / t_ke.item(), "unit": "shapes/s",
                   "h2d_bytes_per_step": int(hq[0].numel() * 4), "d2h_bytes_per_step": q_per_step * KNN_TOPK * 8,
                   "steps": k_steps}
        kflops = 2.0 * q_per_step * n_c * N_POINTS * N_POINTS * D
        kach = kflops / (kms * 1e-3) / 1e12
        knn_obj = {"metric": "knn_retrieval_shapes_per_s", "value": world * q_per_step / (kms * 1e-3), "unit": "shapes/s",
                   "ms_per_step": kms, "steps": k_steps, "e2e": knn_e2e,
                   "config": {"workload": f"{q_per_step} query shapes/rank/step vs {n_c}-shape candidate store, N={N_POINTS}, top-{KNN_TOPK} (configs[2])",
                              "candidate_store_gb": store_rows.numel() * 4 / 1e9, "exact_band_rescore": True, "store_exchange_ms": allgather_ms},
                   "roofline": {"bound": "tensor", "kernel": "knn_score_kernel", "achieved": kach, "peak": pk["tflops_sustained"],
                                "unit": "TFLOP/s", "frac": kach / pk["tflops_sustained"], "traffic": None,
                                "peak_source": pk["source"] + ", sustained bf16 (kernel runs for seconds)"}}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        hi = host_info()
        torch.set_num_threads(hi["physical_cores"])
        cstep, ckind = cpu_step_factory(h, CSA_B, CSA_K)
        cstep()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            cstep()
        cdt = (time.perf_counter() - t0) / n
        cpu = {"value": CSA_B * CSA_K / cdt, "unit": "shape-pairs/s", "cores": hi["physical_cores"], "kind": ckind,
               "cpu_model": hi["cpu_model"],
               "sample": f"the full step ({CSA_B} query shapes x K={CSA_K}), fwd + masked CE + bwd, fp32, eval, 1 warm-up + {n} steps"}
        if knn_obj is not None:
            from oracle import csa_oracle as O
            from csn_b200 import synth as _synth
            nq, nc = 4, 24
            f = _synth.clustered_shapes(5, nq + nc, n_points=N_POINTS, n_categories=4)
            t0 = time.perf_counter()
            O.retrieval_measure(f[:nq], f[nq:])
            kdt = time.perf_counter() - t0
            per_pair = kdt / (nq * nc)
            knn_obj["cpu_baseline"] = {"value": 1.0 / (per_pair * args.knn_candidates), "unit": "shapes/s",
                                       "cores": hi["physical_cores"], "kind": "port", "cpu_model": hi["cpu_model"],
                                       "sample": f"oracle.retrieval_measure on {nq} queries x {nc} candidates of {N_POINTS} points "
                                                 f"({per_pair * 1e3:.1f} ms per pair), extrapolated per pair to {args.knn_candidates} candidates"}

    if rank == 0:
        line = {
            "metric": "csa_shape_pairs_per_s_fwd_bwd", "value": value, "unit": "shape-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision + " operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": f"MID-FC CSA training step B={CSA_B} K={CSA_K} h={h} N={N_POINTS} D={D} per GPU (configs[1])",
                       "heads": h, "parallelism": f"dp{world} over query shapes", "l2": "inputs 410 MB/step > L2, two alternating batches",
                       "dropout": "off (model.eval(): the parity-checkable semantics; training-mode dropout is in the kernels "
                                  "and measured by --train-mode)" if not args.train_mode else "on (model.train(), p = 0.1; fresh masks per replay via csn_set_drop_epoch)",
                       "loss": "module forward + ATen conv / cross-entropy" if args.unfused_loss else
                               "CrossShapeAt.forward_loss (fused head: weighted sum + logit conv + masked CE + IoU counters + backward)",
                       "timed_blocks": n_blocks_timed, "block_ms": [round(b_, 3) for b_ in block_ms],
                       "launch": launch_mode},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "shape-pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "steps": e2e_steps},
            # `e2e` keeps the contract's meaning (pinned HOST feature buffers in the reference loader's fp32 format, PCIe-
            # bound); the two keys below are the same step fed from a 16-bit host cache and from the collection sharded
            # over the GPUs' HBM (the recommended training data path: it scales like the compute)
            "e2e_host16": e2e_host16,
            "e2e_feature_store": e2e_store,
            "roofline": roof,
            "kernels": kernels,
            "cpu_baseline": cpu, "knn": knn_obj,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- configs 4 and 5
def _simple_line(args, metric, value, ms_step, workload, flops, dtype, launches, clocks, e2e, cpu, extra_cfg=None):
    pk = peaks()
    ach = flops / (ms_step * 1e-3) / 1e12
    cfg = {"workload": workload, "parallelism": "1 GPU"}
    cfg.update(extra_cfg or {})
    return {"metric": metric, "value": value, "unit": "shape-pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": {"bound": "tensor", "kernel": "whole step (all launches)", "achieved": ach, "peak": pk["tflops_sustained"],
                         "unit": "TFLOP/s", "frac": ach / pk["tflops_sustained"], "traffic": None,
                         "algorithmic_flops": flops, "peak_source": pk["source"] + ", sustained bf16"},
            "cpu_baseline": cpu}


def _time_steps(step, steps, warmup):
    from csn_b200 import _lib as L
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, L.launch_count() - l0


def run_config4(args) -> None:
    """BASELINE.json configs[3]: MinkowskiNet CSA head (hrnet.py:370-417) on a ragged batch of 8 shapes with
    L_b ~ U[1000, 4000] voxels x 256, K = 3 neighbours, h = 4, d_head = 64, bf16, forward + backward."""
    from csn_b200 import mink, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, K, h = 8, 3, 4
    precision = "bf16" if args.precision == "fp16" and not args.keep_precision else args.precision
    head = mink.CSAHead(256, h, precision=precision).to(dev).eval()
    sd = synth.mink_state(2, h)
    head.load_state_dict(sd, strict=False)
    lens = synth.ragged_lengths(7, B * (K + 1))
    g = synth.gen(8)
    hq = [torch.relu(torch.randn(lens[b], 256, generator=g)).pin_memory() for b in range(B)]
    hk = [[torch.relu(torch.randn(lens[B * (k + 1) + b], 256, generator=g)).pin_memory() for b in range(B)] for k in range(K)]
    q = [t.to(dev).requires_grad_(True) for t in hq]
    keys = [[t.to(dev) for t in row] for row in hk]

    def step(qq=q, kk=keys):
        for p_ in head.parameters():
            p_.grad = None
        out = head(qq, kk)
        loss = sum(o.square().mean() for o in out)
        loss.backward()
        return loss

    sampler = ClockSampler(0)
    sampler.start()
    ms, launches = _time_steps(step, args.steps, args.warmup)
    clocks = sampler.stop()
    # algorithmic FLOPs (full attention, each shape projected once): fwd+bwd = 3 x fwd
    HD = 256
    shapes = lens[:B * (K + 1)]
    pairs = [(b, b) for b in range(B * (K + 1))] + [(b, B * (k + 1) + b) for k in range(K) for b in range(B)]
    fwd = sum(3 * 2.0 * L_ * 256 * HD for L_ in shapes) + sum(4.0 * lens[a] * lens[c] * HD + 2.0 * lens[a] * HD * 256 for a, c in pairs)
    flops = 3.0 * fwd

    def e2e_step():
        qq = [t.to(dev, non_blocking=True).requires_grad_(True) for t in hq]
        kk = [[t.to(dev, non_blocking=True) for t in row] for row in hk]
        return step(qq, kk).item()

    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = max(2, min(args.steps, 5))
    for _ in range(n):
        e2e_step()
    torch.cuda.synchronize()
    e2e_t = (time.perf_counter() - t0) / n
    e2e = {"value": B * K / e2e_t, "unit": "shape-pairs/s", "h2d_bytes_per_step": int(sum(shapes) * 256 * 4), "d2h_bytes_per_step": 4,
           "steps": n}
    cpu = None
    if not args.no_cpu:
        from oracle import csa_oracle as O
        hi = host_info()
        torch.set_num_threads(hi["physical_cores"])
        w = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
        cq = [t.clone().requires_grad_(True) for t in hq[:2]]
        ck = [[t.clone() for t in row[:2]] for row in hk]

        def cstep():
            for v in w.values():
                v.grad = None
            sum(o.square().mean() for o in O.mink_csa_block(cq, ck, w, h)).backward()
        cstep()
        t0 = time.perf_counter()
        cstep()
        cdt = time.perf_counter() - t0
        cpu = {"value": 2 * K / cdt, "unit": "shape-pairs/s", "cores": hi["physical_cores"], "kind": "port", "cpu_model": hi["cpu_model"],
               "sample": f"oracle.mink_csa_block on 2 of the {B} query shapes (x K={K}), fwd+bwd, fp32, 1 warm-up + 1 step"}
    line = _simple_line(args, "mink_csa_head_shape_pairs_per_s_fwd_bwd", B * K / (ms * 1e-3), ms,
                        f"MinkowskiNet CSA head B={B} K={K} h={h} d=64, L_b~U[1000,4000] ({sum(lens[:B])} query points), one ragged batch (configs[3])",
                        flops, precision + " operands, f32 accumulate", launches, clocks, e2e, cpu,
                        {"l2": f"{sum(shapes) * 256 * 4 / 1e6:.0f} MB of features per step; intermediates >> L2"})
    print(json.dumps(line), flush=True)


def run_config5(args) -> None:
    """BASELINE.json configs[4]: MID-FC CSA layer at N = 2k ... 40k points per shape (iters = N / 500; the reference only
    defines N = 10 000), B = 2, K = 3, forward + masked CE + backward as one CUDA graph; one line per N."""
    from csn_b200 import midfc, synth
    from csn_b200.graphs import GraphedStep
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, K, h = 2, 3, args.heads
    for n in (2000, 5000, 10000, 20000, 40000):
        m = midfc.get_model("csa", N_CLASSES, h, K, precision=args.precision).to(dev).eval()
        m.load_state_dict(synth.midfc_state(1, h, N_CLASSES))
        m.attention.iters = n // 500
        x, nb = synth.csa_batch(3, B, K, n_points=n)
        hx, hn = x.pin_memory(), nb.pin_memory()
        x, nb = x.to(dev), nb.to(dev)
        lab = torch.randint(0, N_CLASSES, (B, n), generator=synth.gen(4)).to(dev)
        params = [p_ for k_, p_ in m.named_parameters() if not k_.startswith("fc_1")]

        def step(x_, nb_, lab_):
            for p_ in params:
                p_.grad = None
            loss = m.forward_loss(x_, "test", nb_, lab_)
            loss.backward()
            return loss

        gs = GraphedStep(step, x, nb, lab)
        sampler = ClockSampler(0)
        sampler.start()
        ms, _ = _time_steps(gs.replay, max(args.steps, 20), args.warmup)
        clocks = sampler.stop()
        flops = 3.0 * B * csa_flops_per_query(K, h, N=n)

        def e2e_step():
            x.copy_(hx, non_blocking=True)
            for b in range(B):
                nb[b, 1:].copy_(hn[b, 1:], non_blocking=True)
            return gs.replay().item()

        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            e2e_step()
        torch.cuda.synchronize()
        e2e_t = (time.perf_counter() - t0) / 5
        e2e = {"value": B * K / e2e_t, "unit": "shape-pairs/s", "h2d_bytes_per_step": int(hx.numel() * 4 + hn[:, 1:].numel() * 4),
               "d2h_bytes_per_step": 4, "steps": 5}
        line = _simple_line(args, "csa_shape_pairs_per_s_fwd_bwd", B * K / (ms * 1e-3), ms,
                            f"MID-FC CSA training step B={B} K={K} h={h} N={n} (iters={n // 500}) (configs[4])", flops,
                            args.precision + " operands, f32 accumulate", gs.launches, clocks, e2e, None,
                            {"launch": "CUDA graph replay", "l2": "inputs %.0f MB/step" % ((hx.numel() + hn.numel()) * 4 / 1e6)})
        print(json.dumps(line), flush=True)
        del gs, m


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--heads", type=int, default=1)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--knn-candidates", type=int, default=KNN_CANDIDATES)
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--unfused-loss", action="store_true", help="module forward + ATen conv / cross-entropy instead of the fused head (CrossShapeAt.forward_loss)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--keep-precision", action="store_true", help="config 4: do not switch to bf16 (its stated dtype)")
    ap.add_argument("--train-mode", action="store_true", help="model.train(): dropout p = 0.1 on the attention probabilities and the fc output")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = CSA training step (+ the config-3 kNN sub-object; default), 3 = same line, "
                         "4 = MinkowskiNet CSA head, 5 = N sweep of the MID-FC layer")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4:
        run_config4(args)
    elif args.config == 5:
        run_config5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
