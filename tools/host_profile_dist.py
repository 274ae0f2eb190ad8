"""cProfile of rank 0's host side of one multi-GPU bench step (torchrun)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench as B
nel = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kv, cm, _ = B.build_inputs(nel, net=False)
for _ in range(3):
    B.one_step(kv, cm, None, "fused", 1e-10, False)
dist.barrier(); torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
out = B.one_step(kv, cm, None, "fused", 1e-10, False)
torch.cuda.synchronize()
pr.disable()
ev = out[2]
if dist.get_rank() == 0:
    print("stage ms:", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(3)])
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(60)
    print(s.getvalue()[:12000])
dist.barrier()
dist.destroy_process_group()
