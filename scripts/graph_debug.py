"""Bisect CUDA-graph capture of the train step on the tiny topology (dev tool)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.svd_unet_oracle import TINY_CONFIG, edm_loss, synthetic_batch
from svd_xtend_b200.unet import UNetSpatioTemporalConditionModel
from svd_xtend_b200.train import ParamArena, FusedAdamW
from oracle.svd_unet_oracle import SVD_CONFIG
FULL = "--full" in sys.argv

dev = "cuda:0"
torch.manual_seed(0)
CFG = SVD_CONFIG if FULL else TINY_CONFIG
unet = UNetSpatioTemporalConditionModel(**CFG).to(dev)
unet.requires_grad_(False)
for n, p in unet.named_parameters():
    if "temporal_transformer_block" in n:
        p.requires_grad_(True)
unet.train()
arena = ParamArena(unet); unet.attach_arena(arena)
opt = FusedAdamW(arena, lr=1e-5); opt.on_updated = lambda: unet.refresh_trainable_operands(shadow_current=True)
b = synthetic_batch(1, 14, 40, 64, seed=1, device=dev) if FULL else synthetic_batch(1, 4, 16, 16, seed=1, device=dev, cross_dim=TINY_CONFIG["cross_attention_dim"])

def fwd():
    with torch.no_grad():
        return unet(b["sample"], b["timestep"], b["encoder_hidden_states"], b["added_time_ids"]).sample
def fwdbwd():
    arena.zero_grad()
    pred = unet(b["sample"], b["timestep"], b["encoder_hidden_states"], b["added_time_ids"]).sample
    loss = edm_loss(pred.float(), b["noisy"], b["latents"], b["sigmas"]); loss.backward(); return loss
def full():
    l = fwdbwd(); opt.step(); return l

for name, fn in (("forward", fwd), ("fwd+bwd", fwdbwd), ("full step", full)):
    for mode in ("global", "thread_local", "relaxed"):
        try:
            for _ in range(2): fn()
            torch.cuda.synchronize()
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s): fn()
            torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode=mode):
                out = fn()
            g.replay(); torch.cuda.synchronize()
            print(f"[graph] {name} mode={mode}: OK", float(out.float().sum()))
            break
        except Exception as e:
            print(f"[graph] {name} mode={mode}: FAILED {type(e).__name__}: {str(e)[:300]}")
            traceback.print_exc(limit=6)
            try: torch.cuda.synchronize()
            except Exception as e2: print("sync after failure:", e2)
