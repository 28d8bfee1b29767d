"""N-rank equivalence check (run under torchrun): the all-reduced gradient on every rank equals the mean of the
single-rank gradients computed with the same weights on each rank's clip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle.svd_unet_oracle import TINY_CONFIG, edm_loss, synthetic_batch
from svd_xtend_b200.unet import UNetSpatioTemporalConditionModel
from svd_xtend_b200.train import ParamArena, GradReducer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
unet = UNetSpatioTemporalConditionModel(**TINY_CONFIG).to(dev)
unet.requires_grad_(False)
for n, p in unet.named_parameters():
    if "temporal_transformer_block" in n:
        p.requires_grad_(True)
unet.train()
arena = ParamArena(unet); unet.attach_arena(arena)
red = GradReducer(arena, bucket_mb=0.5)
unet.grad_hook = lambda ps: red.on_grads_ready(ps) if ps is not None else None

def grads_for(seed, reduce):
    b = synthetic_batch(1, 4, 16, 16, seed=seed, device=dev, cross_dim=TINY_CONFIG["cross_attention_dim"])
    arena.zero_grad()
    unet.grad_hook = (lambda ps: red.on_grads_ready(ps) if ps is not None else None) if reduce else None
    pred = unet(b["sample"], b["timestep"], b["encoder_hidden_states"], b["added_time_ids"]).sample
    edm_loss(pred.float(), b["noisy"], b["latents"], b["sigmas"]).backward()
    if reduce:
        red.finish()
    torch.cuda.synchronize()
    return arena.grad.clone()

g_red = grads_for(100 + rank, True)
singles = [grads_for(100 + r, False) for r in range(world)]
ref = sum(singles) / world
rel = ((g_red - ref).norm() / ref.norm()).item()
others = [torch.empty_like(g_red) for _ in range(world)]
dist.all_gather(others, g_red)
same = all(torch.equal(o, g_red) for o in others)
print(f"[ddp_check] rank {rank}: reduced-vs-mean rel-l2 {rel:.3e}, identical across ranks: {same}, buckets {len(red.buckets)}")
assert rel < 2e-2 and same
dist.destroy_process_group()
