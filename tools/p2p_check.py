"""Stand-alone check of the peer-memory allreduce kernel (csrc/comm.cu) at any world size <= 8:
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_check.py
Each rank fills a bf16 buffer with rank-dependent values, reduces buckets of several sizes in place with
pb200_allreduce_p2p (eagerly and from a captured CUDA graph) and compares with NCCL's allreduce of the same data; then
times both. Every wait in the kernel gives up after ~10 s (P2PAllreduce.timeouts()), so a missing peer shows up as an
error line, not as a hung box. Run it under `timeout` all the same. PB200_P2P_MAP=symm selects torch symmetric memory
(cuMem handles) instead of legacy CUDA IPC for the peer mappings."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    from pasero_b200.training import P2PAllreduce
    n = 70_000_000
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    src = (torch.randn(n, device=dev, generator=g) * 0.01).to(torch.bfloat16)
    buf = (P2PAllreduce.symmetric_empty(n, torch.bfloat16, dev) if P2PAllreduce.mapping() == 'symm' else torch.empty_like(src))
    t0 = time.time()
    comm = P2PAllreduce(buf)
    print(f'[{rank}] peer mapping ({P2PAllreduce.mapping()}) done in {time.time() - t0:.2f} s', flush=True)
    ok = True
    for lo, hi in ((0, 8), (8, 8 * 1001), (4096, 4096 + 12_500_000), (0, n)):
        buf.copy_(src)
        ref = src[lo:hi].clone()
        dist.all_reduce(ref)
        comm.allreduce(lo, hi)
        torch.cuda.synchronize()
        err = (buf[lo:hi].float() - ref.float()).abs().max().item()
        untouched = torch.equal(buf[:lo], src[:lo]) and torch.equal(buf[hi:], src[hi:])
        same = [torch.empty_like(buf[lo:hi]) for _ in range(world)]
        dist.all_gather(same, buf[lo:hi].contiguous())
        identical = all(torch.equal(same[0], s) for s in same)
        good = err < 2e-3 and untouched and identical and comm.timeouts() == 0
        ok = ok and good
        if rank == 0:
            print(f'range [{lo}, {hi}): max |p2p - nccl| {err:.3e}, outside untouched {untouched}, replicas identical {identical}, '
                  f'timeouts {comm.timeouts()} -> {"ok" if good else "FAIL"}', flush=True)
    # captured
    buf.copy_(src)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        comm.allreduce(0, n)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        comm.allreduce(0, 12_500_000)
        comm.allreduce(12_500_000, n)
    buf.copy_(src)
    graph.replay()
    ref = src.clone()
    dist.all_reduce(ref)
    torch.cuda.synchronize()
    err = (buf.float() - ref.float()).abs().max().item()
    ok = ok and err < 2e-3 and comm.timeouts() == 0
    if rank == 0:
        print(f'graph replay: max |p2p - nccl| {err:.3e}, timeouts {comm.timeouts()}', flush=True)
    # timing: 140 MB (the config-2 arena) as 6 buckets, p2p graph replay vs NCCL
    for name, fn in (('p2p', graph.replay), ('nccl', lambda: dist.all_reduce(buf))):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            print(f'{name}: {ms * 1e3:.0f} us per 140 MB allreduce = {2 * (world - 1) / world * n * 2 / ms / 1e6:.0f} GB/s bus bandwidth', flush=True)
    if rank == 0:
        print('P2P CHECK OK' if ok else 'P2P CHECK FAILED', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
