import os, sys, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
nb = [r for r in (rank - 1, rank + 1) if 0 <= r < world]
tok = [torch.zeros(1, device="cuda") for _ in range(4)]
def barrier():
    ops = []
    for k, r in enumerate(nb):
        ops.append(dist.P2POp(dist.isend, tok[2 * k], r))
        ops.append(dist.P2POp(dist.irecv, tok[2 * k + 1], r))
    for q in dist.batch_isend_irecv(ops):
        q.wait()
x = torch.zeros(1 << 20, device="cuda")
for _ in range(20): barrier()
torch.cuda.synchronize(); dist.barrier()
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(200): barrier()
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
if rank == 0: print(f"token barrier: GPU {e0.elapsed_time(e1) / 200 * 1e3:.1f} us each, host enqueue {(t1 - t0) / 200 * 1e6:.1f} us each")
# with a 0.5 ms kernel between barriers (host ahead)
def work():
    for _ in range(8): x.mul_(1.0001)
for _ in range(5): work(); barrier()
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(100): work()
e1.record(); torch.cuda.synchronize(); tw = e0.elapsed_time(e1) / 100
e0.record()
for _ in range(100): work(); barrier()
e1.record(); torch.cuda.synchronize(); tb = e0.elapsed_time(e1) / 100
if rank == 0: print(f"work {tw * 1e3:.1f} us, work + barrier {tb * 1e3:.1f} us -> barrier adds {(tb - tw) * 1e3:.1f} us")
dist.destroy_process_group()
