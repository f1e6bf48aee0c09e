"""Experiment: all-gather of the packed flag bits by copy engines over peer (symmetric) memory -- no SM involved.
torchrun --nproc-per-node N profiles/exp_symm_gather.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
n_words = (1 << 24) // 32
t0 = time.time()
buf = symm.empty(n_words, dtype=torch.int32, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok in %.2f s" % (time.time() - t0), type(hdl).__name__, hdl.rank, hdl.world_size, flush=True)
buf.fill_(rank + 1)
out = torch.zeros(world, n_words, dtype=torch.int32, device=dev)
torch.cuda.synchronize(); dist.barrier()
side = torch.cuda.Stream(priority=-1)
def gather():
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        hdl.barrier()
        for s in range(world):
            r = (rank - s) % world
            out[r].copy_(hdl.get_buffer(r, (n_words,), torch.int32))
        hdl.barrier()
for _ in range(3): gather()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(side):
    a.record()
for _ in range(20): gather()
with torch.cuda.stream(side):
    b.record()
torch.cuda.synchronize()
ok = all(int(out[r][0]) == r + 1 and int(out[r][-1]) == r + 1 for r in range(world))
print(rank, "gather ok", ok, "us per gather", a.elapsed_time(b) / 20 * 1e3, flush=True)
# NCCL for comparison
g = torch.empty(world * n_words, dtype=torch.int32, device=dev)
for _ in range(3): dist.all_gather_into_tensor(g, buf)
torch.cuda.synchronize()
a.record()
for _ in range(20): dist.all_gather_into_tensor(g, buf)
b.record(); torch.cuda.synchronize()
print(rank, "nccl us per gather", a.elapsed_time(b) / 20 * 1e3, flush=True)
dist.destroy_process_group()
