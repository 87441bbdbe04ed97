"""Host-entry throughput with every rank of the box feeding its own GPU from pinned host arrays at the same
time (the ranks share the host's cores, memory and PCIe fabric):  torchrun --nproc-per-node N tools/e2e_multi.py
Compares the flow as float32 with the default (3/4 of each chunk narrowed to binary16 on the CPU)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from davo_b200 import synthetic as S
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 128
pinned = tuple(torch.as_tensor(x).pin_memory().numpy() for x in S.make_inputs(B, 128, 416, seed=3 + rank))
w = S.init_weights(ver)
out = {"n_gpus": world, "host_threads": os.cpu_count()}
for name, env in (("flow_float32", {"DAVO_B200_HOST_FLOW16": "0"}), ("default", {}), ("frac_0.375", {"DAVO_B200_HOST_FLOW16_FRAC": "0.375"})):
    os.environ.update(env)
    s = DAVO(version=ver)
    s.setup_inference(128, 416, "davo", 3, B, device=local)
    s.load_weights(w)
    for _ in range(4):
        s.inference(None, "pose", inputs=pinned)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(25):
        s.inference(None, "pose", inputs=pinned)
    dt = torch.tensor([(time.perf_counter() - t0) / 25], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out[name] = {"pairs_per_s": round(world * 2 * B / float(dt)), "ms": round(float(dt) * 1e3, 3),
                 "h2d_MB_per_rank": round(s.last_host_copy_bytes()[0] / 1e6, 1)}
    s.close()
    for k in env:
        os.environ.pop(k)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
