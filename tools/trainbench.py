"""Time the head's training step (BASELINE cfg 5: Xception OS16 512x512, SyncBN, bf16; 8 images per GPU) with CUDA events.
Single GPU: python tools/trainbench.py [--batch 8] [--steps 10]; multi GPU: torchrun --nproc-per-node N tools/trainbench.py.
Prints one JSON line (rank 0): images/s over all ranks (max-over-ranks device time), kernel launches per step."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--cin', type=int, default=2048)
    ap.add_argument('--cskip', type=int, default=256)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--no-graph', action='store_true', help='launch the kernels one by one instead of replaying the captured CUDA graph')
    ap.add_argument('--profile', action='store_true', help='per-phase timing (forward / backward / all-reduce / update)')
    a = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from dlv3p_b200 import train
    from dlv3p_b200.head import DeepLabHead  # noqa: F401  (package import check)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import head_ref as R          # weight initialisation only (test infrastructure; nothing timed uses it)
    cfg = R.HeadConfig(B=a.batch, H=a.size, W=a.size, OS=16, Cin=a.cin, Cskip=a.cskip, NC=21)
    W = R.make_weights(cfg, 1234)
    tr = train.HeadTrainer(a.batch, a.size, a.size, 16, a.cin, a.cskip, 21, W, device=local, seed=7, graph=not a.no_graph)
    g = torch.Generator(device='cuda').manual_seed(1234 + rank)
    feat = torch.randn(a.batch, cfg.h, cfg.w, a.cin, device='cuda', generator=g).clamp_(min=0).to(torch.bfloat16)
    skip = torch.randn(a.batch, cfg.hs, cfg.ws, a.cskip, device='cuda', generator=g).to(torch.bfloat16)
    labels = torch.randint(0, 21, (a.batch, a.size, a.size), device='cuda', generator=g, dtype=torch.uint8)
    for _ in range(a.warmup):
        tr.train_step(feat, skip, labels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = tr.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        tr.train_step(feat, skip, labels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = (tr.launches - l0) // a.steps
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss = tr.loss()
    phases = None
    if a.profile:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(); tr.forward_backward(feat, skip, labels); ev[1].record(); tr.all_reduce_gradients(); ev[2].record(); tr.apply_gradients(); ev[3].record()
        torch.cuda.synchronize()
        phases = {'forward_backward_ms': ev[0].elapsed_time(ev[1]), 'grad_allreduce_ms': ev[1].elapsed_time(ev[2]), 'update_ms': ev[2].elapsed_time(ev[3])}
    if rank == 0:
        print(json.dumps({'metric': 'images/sec DeepLabV3+ head training step (fwd+loss+bwd+SyncBN+grad all-reduce+SGD)', 'value': a.batch * world * a.steps / (ms / 1e3),
                          'unit': 'images/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'dtype': 'bf16',
                          'config': {'workload': 'cfg5 Xception OS16 %dx%d head train step, %d img/GPU, global batch %d' % (a.size, a.size, a.batch, a.batch * world)},
                          'gpu_launches_per_step': launches, 'cuda_graph': not a.no_graph, 'loss': loss, 'phases': phases, 'exchange': tr.comm_backend()}))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
