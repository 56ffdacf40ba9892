import os, time, torch, torch.distributed as dist
dist.init_process_group("nccl")
r = dist.get_rank(); torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
w = dist.get_world_size()
x = torch.zeros(40_000_000 // w, device=dev); out = torch.zeros(40_000_000 // w * w, device=dev)
g = torch.zeros(1_700_000, device=dev); s = torch.zeros(8, dtype=torch.float64, device=dev)
ms_ag = t(lambda: dist.all_gather_into_tensor(out, x)); ms_ar = t(lambda: dist.all_reduce(g)); ms_s = t(lambda: dist.all_reduce(s), 20)
if r == 0: print(f"world {w}: all_gather 160MB {ms_ag:.3f} ms ({0.16*(w-1)/w/ms_ag*1e3:.1f} GB/s per rank) | all_reduce 6.8MB {ms_ar:.3f} ms | all_reduce 64B {ms_s:.3f} ms", flush=True)
dist.destroy_process_group()
