"""One eager CSA training step (config 2: B=8, K=3, h=1, N=10 000) after two warm-up steps — the target of the ncu
captures under profiles/ (launch list and --set full summary).  Usage: python scripts/ncu_step.py [heads]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from csn_b200 import midfc, synth

h = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
model = midfc.get_model("csa", 15, h, 3).to(dev).eval()
model.load_state_dict(synth.midfc_state(1, h, 15))
g = torch.Generator(device=dev).manual_seed(1)
nb = torch.relu(torch.randn(8, 4, 256, 10000, 1, device=dev, generator=g))
x = nb[:, 0].clone()
lab = torch.randint(0, 15, (8, 10000), device=dev, generator=g)
params = [p for n, p in model.named_parameters() if not n.startswith("fc_1")]
# two identical steps (the backward pass runs on autograd's own thread, so NVTX push/pop ranges cannot delimit a step:
# the summaries take the SECOND half of the captured launches)
for i in range(2):
    for p in params:
        p.grad = None
    loss = model.forward_loss(x, "test", nb, lab)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
