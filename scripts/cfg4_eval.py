"""BASELINE configs[3]: N synthetic 4 s clips batch-sharded over the ranks, one all-reduce of the metric sums.

    python scripts/cfg4_eval.py --clips 100000                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/cfg4_eval.py --clips 100000                                      # 8 GPUs, same sums

Prints one JSON line on rank 0: clips/s (max-over-ranks device time), the five means and the transform checksum -
identical for every world size (data are generated from global chunk indices)."""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap_ = argparse.ArgumentParser()
ap_.add_argument("--clips", type=int, default=100000)
ap_.add_argument("--chunk", type=int, default=1024)
args = ap_.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
ap = pkg.audioprocessor.AudioProcessor(sampling_rate=16000, n_fft=512, hop_length=160, win_length=512, audio_length=4)
D = pkg.distributed
D.evaluate_synthetic_sharded(ap, min(args.clips, 2 * args.chunk * world), args.chunk)  # warm-up (plans, allocator)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
total = D.evaluate_synthetic_sharded(ap, args.clips, args.chunk)
b.record()
torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    out = pkg.LMAC_metrics.finalize(total[:6])
    print(json.dumps({"workload": f"configs[3]: {args.clips} synthetic 4 s clips, chunks of {args.chunk}, data generated "
                                  "on the device inside the timed region", "n_gpus": world,
                      "clips_per_s": args.clips / float(t), "seconds": float(t), "means": out,
                      "checksum": [float(x) for x in total[6:].cpu()]}))
if world > 1:
    dist.destroy_process_group()
