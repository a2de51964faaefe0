import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.nn.functional as F
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.synth import synth_batch
dev = torch.device("cuda:0")
def rel(a, c): return float((a.double() - c.double()).abs().max()) / max(float(c.double().abs().max()), 1e-30)
torch.manual_seed(0)
model = ref_trunks.build_trunk("model1", mnn, seed=42).to(dev).eval()
b = synth_batch(300, 99, device=dev)
x0 = (b.x + 0.05 * torch.randn(b.x.shape, device=dev)).detach()
def run():
    x = x0.clone().requires_grad_(True)
    c1 = model.conv1(x, b.edge_index); h1 = torch.relu(c1)
    c2 = model.conv2(h1, b.edge_index); h2 = torch.relu(c2)
    emb = torch.cat([mnn.global_max_pool(h2, b.batch), mnn.global_mean_pool(h2, b.batch)], 1)
    out = model.out(model.fc_g2(torch.relu(model.fc_g1(emb))))
    loss = F.mse_loss(out.view(-1), b.y)
    for t in (c1, h1, c2, h2, emb): t.retain_grad()
    loss.backward()
    r = dict(c1=c1, c2=c2, emb=emb, out=out, g_emb=emb.grad, g_h2=h2.grad, g_c2=c2.grad, g_h1=h1.grad, g_c1=c1.grad, g_x=x.grad)
    r.update({"g_" + k: p.grad.clone() for k, p in model.named_parameters()})
    model.zero_grad()
    return {k: v.detach().clone() for k, v in r.items()}
os.environ["MGS_DISABLE_TC"] = "1"; ff = run()
os.environ.pop("MGS_DISABLE_TC"); tc = run()
for k in ff: print(f"{k:24s} tc vs ffma {rel(tc[k], ff[k]):.2e}")
