"""Raw pinned-memory copy bandwidth with every rank copying at the same time (no kernels): names the resource the
end-to-end path (`HostPoolingPipeline`) shares between the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_probe.py

Prints one JSON line from rank 0: per-GPU and aggregate GB/s for host->device alone, device->host alone and both
directions at once (two streams), each as the max over ranks of the CUDA-event time of 8 x 256 MB copies started
behind a barrier.
"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    nbytes, reps = 256 << 20, 8
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    d_out = torch.ones(nbytes, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(h2d, d2h):
        for it in range(2):                                   # first pass = warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {'what': 'raw pinned copies, all ranks at once', 'n_gpus': world, 'bytes_per_copy': nbytes, 'copies': reps,
           'cpus': os.cpu_count()}
    for name, a, b in (('h2d', True, False), ('d2h', False, True), ('both', True, True)):
        ms = timed(a, b)
        per = nbytes * reps * (int(a) + int(b)) / (ms * 1e-3) / 1e9
        out[name] = {'ms': round(ms, 3), 'gbs_per_gpu': round(per, 2), 'gbs_aggregate': round(per * world, 2)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
