"""N-GPU check of the sharded bank + NCCL all-gather (sdirt_b200.sharding): every rank ends with the full bank, equal to a
single-GPU run of the same points.  Launch with torchrun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from sdirt_b200 import lens_file, sharding
from sdirt_b200.deeplens import PSFNet

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lens = PSFNet(lens_file("rf50mm"), sensor_res=(512, 768), kernel_size=21, device=dev)
lens.numerics = "adaptive"
g = torch.Generator().manual_seed(3)
pts = torch.cat([torch.rand(37, 2, generator=g) * 2 - 1, -(torch.rand(37, 1, generator=g) * 8000 + 300) + 62.25], 1)
L, R = sharding.psf_bank_sharded(lens, pts, ks=21, spp=200000, seed=11, gather=True)
torch.manual_seed(11)
L1, R1 = lens.psf_dp(pts, ks=21, spp=200000)
err = max(float((L - L1).abs().max()), float((R - R1).abs().max()))
t = torch.tensor([err], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: gathered bank {tuple(L.shape)} on every rank, max |sharded - single| = {t.item():.3e}")
assert t.item() < 1e-6
dist.destroy_process_group()
