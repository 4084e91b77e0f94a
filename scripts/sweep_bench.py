"""Secondary measurements of BASELINE.json's configs 4 and 5, CUDA-event timed.

  config 4: MinkowskiNet CSA head (mink.CSAHead), batch of 8 ragged shapes L_b ~ U[1000, 4000], K = 1..3
            neighbours, h = 4, d_head = 64, bf16 operands, forward + backward.
  config 5: MID-FC CSA layer at N in {2k, 5k, 10k, 20k, 40k} points (iters = N / 500), B = 2, K = 3, h = 1,
            forward + backward (the reference only defines N = 10 000; other N are an extension).
Prints one JSON object per line; results are copied to profiles/.
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from csn_b200 import midfc, mink, synth  # noqa: E402


def timed(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def config4():
    dev = torch.device("cuda:0")
    head = mink.CSAHead(256, 4, precision="bf16").to(dev).eval()
    head.load_state_dict(synth.mink_state(2, 4), strict=False)
    B = 8
    for K in (1, 2, 3):
        lens = synth.ragged_lengths(7, B * (K + 1))
        g = synth.gen(8)
        q = [torch.relu(torch.randn(lens[b], 256, generator=g)).to(dev).requires_grad_(True) for b in range(B)]
        keys = [[torch.relu(torch.randn(lens[B * (k + 1) + b], 256, generator=g)).to(dev) for b in range(B)] for k in range(K)]

        def step():
            for p in head.parameters():
                p.grad = None
            out = head(q, keys)
            sum(o.square().mean() for o in out).backward()

        ms = timed(step)
        pts = sum(lens[:B])
        print(json.dumps({"config": 4, "workload": f"MinkowskiNet CSA head B={B} K={K} h=4 d=64 bf16, L_b~U[1000,4000] ({pts} query points)",
                          "ms_per_step": round(ms, 3), "shape_pairs_per_s": round(B * K / (ms * 1e-3), 1)}), flush=True)


def config5():
    dev = torch.device("cuda:0")
    B, K, h = 2, 3, 1
    for n in (2000, 5000, 10000, 20000, 40000):
        m = midfc.get_model("csa", 15, h, K).to(dev).eval()
        m.load_state_dict(synth.midfc_state(1, h, 15))
        m.attention.iters = n // 500
        x, nb = synth.csa_batch(3, B, K, n_points=n)
        x, nb = x.to(dev), nb.to(dev)
        params = [p for k, p in m.named_parameters() if not k.startswith("fc_1")]

        def step():
            for p in params:
                p.grad = None
            m.get_csa_feats(x, nb, "test").square().mean().backward()

        ms = timed(step)
        from csn_b200.graphs import GraphedStep
        gs = GraphedStep(lambda a, b: step(), x, nb)     # the same step replayed as one CUDA graph
        ms_g = timed(gs.replay)
        flops = 3.0 * B * 2 * n * h * 256 * (3 * (K + 1) * 256 + (1 + 2 * K) * (2 * 500 + 256))
        print(json.dumps({"config": 5, "workload": f"MID-FC CSA layer fwd+bwd B={B} K={K} h={h} N={n} (iters={n // 500})",
                          "ms_per_step_eager": round(ms, 3), "ms_per_step_graph": round(ms_g, 3),
                          "shape_pairs_per_s": round(B * K / (ms_g * 1e-3), 1),
                          "algorithmic_tflops": round(flops / (ms_g * 1e-3) / 1e12, 1)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["4", "5"]
    if "4" in which:
        config4()
    if "5" in which:
        config5()
