"""Kernel timeline of one data-parallel xDeepFM train step (rank 0), from torch.profiler.
torchrun --nproc-per-node N tools/dp_timeline.py  ->  gpurun_out/dp_timeline_w{N}.txt
A profiler run: the durations are indicative only, never bench values."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from deeptables_b200.deepmodel import DeepModel
    from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn
    conf = bench.make_config('xdeepfm', 0)
    cats = [CategoricalColumn(f'C{i + 1}', 1000000, bench.EMB_DIM) for i in range(bench.F_FIELDS)]
    conts = [ContinuousColumn('input_continuous_all', [f'I{i + 1}' for i in range(bench.N_DENSE)])]
    model = DeepModel('binary', 2, conf, cats, conts, seed=1234)
    model._build_model()
    host = bench.synth_batches(4, 65536, 1000000, 1234 + rank)
    devb = [tuple(t.cuda() for t in hb) for hb in host]
    for s in range(6):
        model.train_step(*devb[s % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for s in range(3):
            model.train_step(*devb[s % 4])
        torch.cuda.synchronize()
    if rank == 0:
        os.makedirs('gpurun_out', exist_ok=True)
        path = f'gpurun_out/dp_trace_w{world}.json'
        prof.export_chrome_trace(path)
        ev = json.load(open(path))['traceEvents']
        ks = [e for e in ev if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
        ks.sort(key=lambda e: e['ts'])
        t0 = ks[0]['ts']
        # one full cycle of a step: from the loss kernel of the second profiled step to the loss kernel of the third
        # (loss -> backward -> exchange -> Adam -> next forward)
        marks = [i for i, e in enumerate(ks) if 'loss_kernel' in e['name']]
        if len(marks) >= 2:
            ks = ks[marks[-2]:marks[-1]]
            t0 = ks[0]['ts']
        with open(f'gpurun_out/dp_timeline_w{world}.txt', 'w') as f:
            f.write(f'# one data-parallel train step on rank 0 of {world} (torch.profiler; us since the step start, duration, '
                    f'stream, kernel); step span {ks[-1]["ts"] + ks[-1]["dur"] - t0:.0f} us\n')
            for e in ks:
                f.write(f"{e['ts'] - t0:10.1f} {e['dur']:9.1f} s{e['args'].get('stream', -1):<4} {e['name'][:90]}\n")
        os.remove(path)
    if world > 1:
        dist.barrier()


if __name__ == '__main__':
    main()
