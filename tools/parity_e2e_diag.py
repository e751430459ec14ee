"""e2e LoRA-gradient error per tensor for the ViT-B/16 golden case: ours vs fp64 oracle, next to
PyTorch's own bf16 autocast of the oracle on the same GPU (the inherent bf16-operand floor)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import vit_oracle as vo
from tests.golden.make_golden import CASES, synth_inputs
from tests.test_e2e_gpu import build_model

name = os.environ.get("CASE", "vitb16")
cfg, n, c, seed = CASES[name]
n = int(os.environ.get("N", n))
w = vo.synth_weights(cfg, seed)
images, labels = synth_inputs(cfg, n, c, seed + 100)
text = vo.synth_text_features(c, cfg.embed_dim, seed + 200)
ls = 1.0 / 0.07

def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double().flatten(); b = torch.as_tensor(np.asarray(b)).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))

# fp64 truth on the GPU (fast)
def oracle_gpu(dtype, autocast):
    wt = {k: torch.from_numpy(v).cuda().to(dtype).requires_grad_("lora" in k) for k, v in w.items()}
    x = torch.from_numpy(images).cuda().to(dtype); t = torch.from_numpy(text).cuda().to(dtype)
    y = torch.from_numpy(labels).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        feat = vo.vit_forward(x, wt, cfg)
        probs, logits, f = vo.head_forward(feat.float() if autocast else feat, t, ls)
        loss = vo.reference_loss(probs.float() if autocast else probs, y, logits, True)
    loss.backward()
    return {"probs": probs.detach().cpu().numpy(), "loss": float(loss),
            "grads": {k: v.grad.detach().float().cpu().numpy() for k, v in wt.items() if v.requires_grad}}

truth = oracle_gpu(torch.float64, False)
ac = oracle_gpu(torch.float32, True)
m = build_model(cfg, w)
eng = m.model.visual.engine()
eng.forward(torch.from_numpy(images).cuda(), training=True)
head = eng.head(torch.from_numpy(text).cuda(), ls, labels=torch.from_numpy(labels).cuda())
eng.backward_from_head(head)
torch.cuda.synchronize()
names = [k for k in w if "lora" in k]
ours = {k: g.cpu().numpy() for k, g in zip(names, eng.lora_grad_views)}
print(f"case {name} N={n}: probs rel ours {rel(head.probs.cpu().numpy(), truth['probs']):.3e}  autocast {rel(ac['probs'], truth['probs']):.3e}")
print(f"loss ours {float(head.loss_rows.sum()):.8f} autocast {ac['loss']:.8f} truth {truth['loss']:.8f}")
wo = wa = 0
for k in names:
    eo, ea = rel(ours[k], truth["grads"][k]), rel(ac["grads"][k], truth["grads"][k])
    wo, wa = max(wo, eo), max(wa, ea)
    print(f"{k[29:]:45s} ours {eo:.3e}   autocast {ea:.3e}")
print(f"WORST ours {wo:.3e}  autocast {wa:.3e}")
