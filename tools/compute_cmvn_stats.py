#!/usr/bin/env python3
"""compute_cmvn_stats -- the CLI the reference lacks (it only ships the consumer, openeat/utils/cmvn.py).

    python tools/compute_cmvn_stats.py --in_list data/train/format.data --out_cmvn data/train/global_cmvn
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/compute_cmvn_stats.py ...

Reads a wenet ``data.list`` (json lines) or an OpenEAT ``format.data``, shards the utterances over the ranks by
length (no data-path collective), accumulates sum / sum of squares / frame count of the raw 80-bin log-mel
features on each GPU (fp64, in the fbank kernel's epilogue), all-reduces the 161 doubles once (NCCL) and lets
rank 0 write the JSON that ``load_cmvn(path, is_json=True)`` reads (``mean_stat``, ``var_stat``, ``frame_num``).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--in_list', required=True, help='data.list (json lines) or format.data')
    ap.add_argument('--out_cmvn', required=True, help='output JSON stats file')
    ap.add_argument('--mel_bins', type=int, default=80)
    ap.add_argument('--batch_size', type=int, default=256)
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    from openeat_b200.cmvn import compute_cmvn_stats
    from openeat_b200.processor import parse_raw
    from openeat_b200.sharding import shard_by_length

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', init_method='env://', device_id=torch.device('cuda', local))
    with open(args.in_list, encoding='utf-8') as f:
        lines = [l for l in f if l.strip()]
    # shard by file size (a proxy for the duration that needs no decoding), deterministic on every rank
    sizes = []
    for l in lines:
        path = (l.split('"wav"')[1].split('"')[1] if l.lstrip().startswith('{') else l.split('\t')[1].split(':', 1)[1]).split(',')[0]
        try:
            sizes.append(os.path.getsize(path))
        except OSError:
            sizes.append(0)
    mine = shard_by_length(sizes, world)[rank]

    def batches():
        for i in range(0, len(mine), args.batch_size):
            group = [s for s in parse_raw(lines[j] for j in mine[i:i + args.batch_size]) if s['sample_rate'] == 16000]
            yield [s['wav'] for s in group]

    s, q, n = compute_cmvn_stats(batches(), mel_bins=args.mel_bins, out_json=args.out_cmvn)
    if rank == 0:
        print('frames %d  mean[0] %.4f  -> %s' % (n, s[0] / max(n, 1), args.out_cmvn))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
