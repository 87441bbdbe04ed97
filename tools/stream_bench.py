"""BASELINE.json configs[2]: a synthetic KITTI-seq-00-length stream (4541 frames = 4539 samples) sharded by
sample across the ranks (contiguous blocks, padded with the reference's complete_batch_size rule), poses
all-gathered over NCCL once, trajectory composed on rank 0.  Inputs are resident in HBM (a 64-sample
block repeated); each rank runs batches of 128 samples in trajectory mode (only the poses the composition
reads).  Launch:  python -m torch.distributed.run --nproc-per-node N tools/stream_bench.py   (or plain python)
Prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from davo_b200 import synthetic as S, geo_utils, parallel
from davo_b200.davo import DAVO
ver = "v1-decay100k-sharedNN-dilatedPoseNN-cnv6_128-segmask_all-se_flow-abs_flow-fc_tanh"
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, B, H, W = 4539, 128, 128, 416
idx = parallel.padded_indices(N, rank, world)                      # this rank's samples (padded to equal length)
n_local = len(idx)
base = [torch.as_tensor(x).cuda() for x in S.make_inputs(64, H, W, seed=1000)]
sysm = DAVO(version=ver)
sysm.setup_inference(H, W, "davo", 3, B, device=local)
sysm.load_weights(S.init_weights(ver))
if world > 1 and os.environ.get("STREAM_GATHER", "library") == "library":
    sysm.init_comm(rank, world)                                    # davo_allgather_poses; else torch.distributed
poses = torch.empty((n_local, 2, 6), dtype=torch.float32, device="cuda")
def run_stream():
    for b0 in range(0, n_local, B):
        ids = idx[b0:b0 + B]
        sel = torch.as_tensor([i % 64 for i in ids], device="cuda")
        batch = tuple(t.index_select(0, sel) for t in base)      # the batch's samples, gathered on the device
        mode = "trajectory_first" if (rank == 0 and b0 == 0) else "trajectory"
        poses[b0:b0 + len(ids)] = sysm.inference(None, "pose", inputs=batch, as_torch=True, pairs=mode)["pose"]
    return parallel.gather_poses(poses, N, sysm)
for _ in range(2):
    allp = run_stream()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5
for _ in range(reps):
    allp = run_stream()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t0 = time.perf_counter()
    traj = geo_utils.compose_trajectory(allp.cpu().numpy())
    t_host = time.perf_counter() - t0
    print(json.dumps({"workload": "4541-frame stream, sample-sharded x%d, trajectory mode, NCCL all-gather of poses" % world,
                      "n_gpus": world, "frames": int(traj.shape[0]), "ms_per_stream": float(ms.item()),
                      "frames_per_s": 4541 / (float(ms.item()) * 1e-3), "host_composition_ms": 1e3 * t_host}), flush=True)
if world > 1:
    dist.destroy_process_group()
